// gw-b200 fixed-size global aligners (sm_100a): Hirschberg-Myers (what the deprecated create_aligner(max_query, max_target,
// max_alignments, ...) factory and pygenomeworks.CudaAlignerBatch build) and the unbanded Myers aligner.
//
// Behavioural contract = the reference's hirschberg_myers_compute_alignment and callees (cudaaligner/src/
// hirschberg_myers_gpu.cu:240-699) and myers_compute_score_matrix_kernel + myers_backtrace_kernel (myers_gpu.cu:256-442):
// identical paths, i.e. the same query split (midpoint), the same target midpoint among equal score sums (first minimum per
// lane in stride-32 order, then the shuffle tree that prefers the lower lane), the same switch to the full matrix (query slice
// < 63 characters whose (target + 1) x words fit ceil(max_query / 32) x 64 words), the same 64-entry range stack (overflow =
// failed alignment, length 0), the same single-character rule and the same backtrace preference (insertion, deletion, diagonal).
// The implementation is this repo's: one warp per alignment; the bit-vector columns of a score pass live in shared memory
// (one word per lane and 32-word chunk, carries resolved with the ballot adder of myers_kernels.cuh, horizontal delta handed
// from chunk to chunk through a shuffle); the bit patterns of the query and of the reversed query are built once per
// alignment; leaf matrices (<= 2 words x target slice) are walked out of shared memory when they fit.
#pragma once

#include "myers_kernels.cuh"

namespace gwb200
{
namespace galign
{

using myers::WordType;
using myers::kFull;
using myers::kWord;

struct GlobalParams
{
    const char* seqs;            // [n][2][max_len]: query, target (after reverse complement on the host)
    const int32_t* seq_lengths;  // [2n]
    int32_t max_len;             // max(max_query_length, max_target_length)
    int32_t n_alignments;
    int32_t max_query_length, max_target_length;
    int32_t max_result_length;
    int8_t* results;             // [n][max_result_length] path, end -> start
    int32_t* result_lengths;     // [n] (0 = failed)
    // per-alignment workspaces
    WordType* qpat;              // [n][8][pat_stride] forward A,C,T,G then reverse A,C,T,G, one zero word of padding per row
    int32_t pat_stride;          // ceil(max_query / 32) + 1
    int32_t* scores;             // [n][2][max_target + 1] forward / reverse last-row scores
    WordType* leaf_pv;           // [n][leaf_elems]
    WordType* leaf_mv;
    int32_t* leaf_sc;
    int64_t leaf_elems;          // ceil(max_query / 32) * 64 (aligner_global_hirschberg_myers.cpp:36-48)
    WordType* col_ws;            // [n][2][pat_stride] column state when it does not fit shared memory
    int32_t col_smem_words;      // words of pv / mv the dynamic shared memory holds (each)
    int32_t full_myers_threshold; // 63
    int32_t algorithm;           // 0 = Hirschberg-Myers, 1 = unbanded Myers
    unsigned long long* cells;   // executed DP cells (sum over passes of query slice x target slice)
};

constexpr int32_t kStackEntries = 64;

// 32-bit window at bit offset `off` of pattern row `row` (zero beyond the row: the padding word)
__device__ __forceinline__ WordType pat_word(const WordType* row, int32_t idx, int32_t off)
{
    const int32_t w     = idx + (off >> 5);
    const int32_t shift = off & 31;
    return __funnelshift_r(row[w], row[w + 1], shift);
}

// Last-row scores D(query slice, target prefix t), t = 0..T (myers_compute_scores with full_score_matrix = false,
// hirschberg_myers_gpu.cu:240-352). Column state pv / mv in `cpv` / `cmv` (shared or global), one word per lane and chunk.
__device__ void score_pass(const GlobalParams& P, const WordType* pat /* [8][pat_stride] */, const char* target, int32_t T, int32_t qlen,
                           int32_t pat_off, bool reverse, WordType* cpv, WordType* cmv, int32_t* out, int32_t lane)
{
    const int32_t n_words  = (qlen + 31) >> 5;
    const int32_t n_chunks = (n_words + 31) >> 5;
    for (int32_t idx = lane; idx < n_words; idx += 32)
    {
        cpv[idx] = ~WordType(0);
        cmv[idx] = 0;
    }
    if (lane == 0)
        out[0] = qlen;
    __syncwarp();
    int32_t score = qlen;
    const WordType* pbase = pat + (reverse ? 4 : 0) * P.pat_stride;
    for (int32_t t = 1; t <= T; ++t)
    {
        const char tc          = reverse ? target[T - t] : target[t - 1];
        const WordType* row    = pbase + ((tc >> 1) & 3) * P.pat_stride;
        int32_t carry          = (lane == 0) ? 1 : 0;
        int32_t hout_last      = 0;
        for (int32_t c = 0; c < n_chunks; c++)
        {
            const int32_t idx   = c * 32 + lane;
            const int32_t cw    = min(32, n_words - c * 32); // words in this chunk
            const uint32_t mask = cw == 32 ? kFull : ((1u << cw) - 1u);
            if (lane < cw)
            {
                WordType pv       = cpv[idx];
                WordType mv       = cmv[idx];
                const WordType hb = WordType(1) << (idx == n_words - 1 ? qlen - (n_words - 1) * kWord - 1 : kWord - 1);
                const WordType eq = pat_word(row, idx, pat_off);
                const int2 h      = myers::advance_block(mask, lane, hb, eq, pv, mv, carry);
                cpv[idx]          = pv;
                cmv[idx]          = mv;
                hout_last         = h.x;
            }
            __syncwarp();
            // horizontal delta of the chunk's top word enters the next chunk at its lowest word
            const int32_t top = __shfl_sync(kFull, hout_last, cw - 1);
            carry             = (lane == 0) ? top : 0;
            hout_last         = top;
        }
        score += hout_last; // uniform: delta at the highest bit of the last word
        if (lane == 0)
            out[t] = score;
    }
    __syncwarp();
}

// Leaf: the full Myers matrix of a query slice of <= 2 words and its backtrace (hirschberg_myers_compute_path :354-410,
// append_myers_backtrace :145-204). Appends to path (end -> start), returns the number of entries.
__device__ int32_t leaf_path(const GlobalParams& P, const WordType* pat, const char* target, int32_t T, const char* /*query*/, int32_t qlen,
                             int32_t pat_off, WordType* pvm, WordType* mvm, int32_t* scm, int8_t* path, int32_t lane)
{
    const int32_t n_words = (qlen + 31) >> 5; // 1 or 2 (general in the code below)
    const uint32_t mask   = n_words >= 32 ? kFull : ((1u << n_words) - 1u);
    // columns: element (w, t) at w + n_words * t
    if (lane < n_words)
    {
        WordType pv       = ~WordType(0);
        WordType mv       = 0;
        int32_t sc        = min((lane + 1) * kWord, qlen);
        const WordType hb = WordType(1) << (lane == n_words - 1 ? qlen - (n_words - 1) * kWord - 1 : kWord - 1);
        pvm[lane]         = pv;
        mvm[lane]         = mv;
        scm[lane]         = sc;
        for (int32_t t = 1; t <= T; ++t)
        {
            const char tc       = target[t - 1];
            const WordType* row = pat + ((tc >> 1) & 3) * P.pat_stride;
            const WordType eq   = pat_word(row, lane, pat_off);
            const int2 h        = myers::advance_block(mask, lane, hb, eq, pv, mv, lane == 0 ? 1 : 0);
            sc += h.x;
            pvm[lane + n_words * t] = pv;
            mvm[lane + n_words * t] = mv;
            scm[lane + n_words * t] = sc;
        }
    }
    __syncwarp();
    int32_t pos = 0;
    if (lane == 0)
    {
        const WordType last_mask = (qlen & 31) != 0 ? (WordType(1) << (qlen & 31)) - 1 : ~WordType(0);
        auto get = [&](int32_t i, int32_t j) -> int32_t {
            const int32_t w = (i - 1) >> 5;
            WordType m      = (~WordType(1)) << ((i - 1) & 31);
            if (w == n_words - 1)
                m &= last_mask;
            const int32_t o = w + n_words * j;
            return scm[o] - __popc(m & pvm[o]) + __popc(m & mvm[o]);
        };
        int32_t i = qlen, j = T;
        int32_t my = scm[((i - 1) >> 5) + n_words * j];
        while (i > 0 && j > 0)
        {
            const int32_t above = i == 1 ? j : get(i - 1, j);
            const int32_t diag  = i == 1 ? j - 1 : get(i - 1, j - 1);
            const int32_t left  = get(i, j - 1);
            int8_t r;
            if (left + 1 == my)
            {
                r  = myers::st_insertion;
                my = left;
                --j;
            }
            else if (above + 1 == my)
            {
                r  = myers::st_deletion;
                my = above;
                --i;
            }
            else
            {
                r  = (diag == my) ? myers::st_match : myers::st_mismatch;
                my = diag;
                --i;
                --j;
            }
            path[pos++] = r;
        }
        for (; i > 0; --i)
            path[pos++] = myers::st_deletion;
        for (; j > 0; --j)
            path[pos++] = myers::st_insertion;
    }
    pos = __shfl_sync(kFull, pos, 0);
    __syncwarp();
    return pos;
}

// One warp per alignment. Dynamic shared memory: 2 * col_smem_words words (pv, mv column state).
__global__ void __launch_bounds__(32, 16) global_align_kernel(const GlobalParams P)
{
    extern __shared__ __align__(16) WordType s_cols[];
    __shared__ int4 s_stack[kStackEntries];
    const int32_t lane = threadIdx.x;
    const int32_t a    = blockIdx.x;
    if (a >= P.n_alignments)
        return;
    const char* const query  = P.seqs + static_cast<int64_t>(2 * a) * P.max_len;
    const char* const target = P.seqs + static_cast<int64_t>(2 * a + 1) * P.max_len;
    const int32_t qlen_all   = P.seq_lengths[2 * a];
    const int32_t tlen_all   = P.seq_lengths[2 * a + 1];
    int8_t* const path       = P.results + static_cast<int64_t>(a) * P.max_result_length;
    WordType* const pat      = P.qpat + static_cast<int64_t>(a) * 8 * P.pat_stride;
    int32_t* const fwd       = P.scores + static_cast<int64_t>(a) * 2 * (P.max_target_length + 1);
    int32_t* const rev       = fwd + (P.max_target_length + 1);
    WordType* const lpv      = P.leaf_pv + static_cast<int64_t>(a) * P.leaf_elems;
    WordType* const lmv      = P.leaf_mv + static_cast<int64_t>(a) * P.leaf_elems;
    int32_t* const lsc       = P.leaf_sc + static_cast<int64_t>(a) * P.leaf_elems;
    const int32_t nw_all     = (qlen_all + 31) >> 5;
    WordType* cpv            = s_cols;
    WordType* cmv            = s_cols + P.col_smem_words;
    if (nw_all > P.col_smem_words)
    {
        cpv = P.col_ws + static_cast<int64_t>(a) * 2 * P.pat_stride;
        cmv = cpv + P.pat_stride;
    }
    unsigned long long my_cells = 0;

    // query bit patterns, forward and reversed (myers_preprocess, :206-222); character order A, C, T, G via (x >> 1) & 3
    for (int32_t idx = lane; idx <= nw_all; idx += 32)
    {
        WordType f[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0};
        if (idx < nw_all)
        {
            const int32_t off = idx * kWord;
            const int32_t n   = min(qlen_all - off, kWord);
            for (int32_t i = 0; i < n; ++i)
            {
                const char cf = query[off + i];
                const char cr = query[qlen_all - 1 - (off + i)];
                f[0] |= static_cast<WordType>(cf == 'A') << i;
                f[1] |= static_cast<WordType>(cf == 'C') << i;
                f[2] |= static_cast<WordType>(cf == 'T') << i;
                f[3] |= static_cast<WordType>(cf == 'G') << i;
                r[0] |= static_cast<WordType>(cr == 'A') << i;
                r[1] |= static_cast<WordType>(cr == 'C') << i;
                r[2] |= static_cast<WordType>(cr == 'T') << i;
                r[3] |= static_cast<WordType>(cr == 'G') << i;
            }
        }
#pragma unroll
        for (int32_t k = 0; k < 4; k++)
        {
            pat[k * P.pat_stride + idx]       = f[k];
            pat[(4 + k) * P.pat_stride + idx] = r[k];
        }
    }
    __syncwarp();

    int32_t length = 0;
    bool success   = true;
    if (P.algorithm == 1)
    {
        // AlignerGlobalMyers: the whole matrix, then its backtrace (myers_gpu.cu:256-442)
        if (qlen_all == 0 || tlen_all == 0)
        {
            const int8_t v  = qlen_all == 0 ? myers::st_insertion : myers::st_deletion;
            const int32_t n = qlen_all + tlen_all;
            for (int32_t i = lane; i < n; i += 32)
                path[i] = v;
            length = n;
        }
        else
        {
            // the matrices of this mode are sized for the whole problem by the host (leaf_elems >= n_words * (target + 1))
            const int32_t n_words = nw_all;
            const int32_t n_chunks = (n_words + 31) >> 5;
            for (int32_t idx = lane; idx < n_words; idx += 32)
            {
                lpv[idx] = ~WordType(0);
                lmv[idx] = 0;
                lsc[idx] = min((idx + 1) * kWord, qlen_all);
            }
            __syncwarp();
            for (int32_t t = 1; t <= tlen_all; ++t)
            {
                const char tc       = target[t - 1];
                const WordType* row = pat + ((tc >> 1) & 3) * P.pat_stride;
                int32_t carry       = (lane == 0) ? 1 : 0;
                for (int32_t c = 0; c < n_chunks; c++)
                {
                    const int32_t idx   = c * 32 + lane;
                    const int32_t cw    = min(32, n_words - c * 32);
                    const uint32_t mask = cw == 32 ? kFull : ((1u << cw) - 1u);
                    int32_t hout        = 0;
                    if (lane < cw)
                    {
                        WordType pv       = lpv[idx + static_cast<int64_t>(n_words) * (t - 1)];
                        WordType mv       = lmv[idx + static_cast<int64_t>(n_words) * (t - 1)];
                        const WordType hb = WordType(1) << (idx == n_words - 1 ? qlen_all - (n_words - 1) * kWord - 1 : kWord - 1);
                        const int2 h      = myers::advance_block(mask, lane, hb, pat_word(row, idx, 0), pv, mv, carry);
                        hout              = h.x;
                        lpv[idx + static_cast<int64_t>(n_words) * t] = pv;
                        lmv[idx + static_cast<int64_t>(n_words) * t] = mv;
                        lsc[idx + static_cast<int64_t>(n_words) * t] = lsc[idx + static_cast<int64_t>(n_words) * (t - 1)] + h.x;
                    }
                    const int32_t top = __shfl_sync(kFull, hout, cw - 1);
                    carry             = (lane == 0) ? top : 0;
                }
            }
            __syncwarp();
            my_cells += static_cast<unsigned long long>(qlen_all) * tlen_all;
            if (lane == 0)
            {
                const WordType last_mask = (qlen_all & 31) != 0 ? (WordType(1) << (qlen_all & 31)) - 1 : ~WordType(0);
                auto get = [&](int32_t i, int32_t j) -> int32_t {
                    const int32_t w = (i - 1) >> 5;
                    WordType m      = (~WordType(1)) << ((i - 1) & 31);
                    if (w == n_words - 1)
                        m &= last_mask;
                    const int64_t o = w + static_cast<int64_t>(n_words) * j;
                    return lsc[o] - __popc(m & lpv[o]) + __popc(m & lmv[o]);
                };
                int32_t i = qlen_all, j = tlen_all, pos = 0;
                int32_t my = lsc[((i - 1) >> 5) + static_cast<int64_t>(n_words) * j];
                while (i > 0 && j > 0)
                {
                    const int32_t above = i == 1 ? j : get(i - 1, j);
                    const int32_t diag  = i == 1 ? j - 1 : get(i - 1, j - 1);
                    const int32_t left  = get(i, j - 1);
                    int8_t r;
                    if (left + 1 == my)
                    {
                        r  = myers::st_insertion;
                        my = left;
                        --j;
                    }
                    else if (above + 1 == my)
                    {
                        r  = myers::st_deletion;
                        my = above;
                        --i;
                    }
                    else
                    {
                        r  = (diag == my) ? myers::st_match : myers::st_mismatch;
                        my = diag;
                        --i;
                        --j;
                    }
                    path[pos++] = r;
                }
                for (; i > 0; --i)
                    path[pos++] = myers::st_deletion;
                for (; j > 0; --j)
                    path[pos++] = myers::st_insertion;
                length = pos;
            }
            length = __shfl_sync(kFull, length, 0);
        }
    }
    else
    {
        // ---- Hirschberg recursion on an explicit stack (hirschberg_myers, :575-644); ranges are offsets into query / target
        int32_t sp = 0;
        if (lane == 0)
            s_stack[0] = make_int4(0, qlen_all, 0, tlen_all);
        sp = 1;
        __syncwarp();
        while (success && sp > 0)
        {
            const int4 e = s_stack[sp - 1];
            --sp;
            __syncwarp();
            const int32_t ql = e.y - e.x, tl = e.w - e.z;
            if (tl == 0)
            {
                for (int32_t i = lane; i < ql; i += 32)
                    path[length + i] = myers::st_deletion;
                length += ql;
            }
            else if (ql == 0)
            {
                for (int32_t i = lane; i < tl; i += 32)
                    path[length + i] = myers::st_insertion;
                length += tl;
            }
            else if (ql == 1)
            {
                // hirschberg_myers_single_char_warp (:483-519): the last target character equal to the query character is the
                // match, everything else an insertion; without any, the first target character is the mismatch
                const char qc = query[e.x];
                int32_t best  = -1; // largest target offset (within the slice) that matches
                for (int32_t i = lane; i < tl; i += 32)
                    if (target[e.z + i] == qc)
                        best = i;
                for (int32_t d = 16; d > 0; d >>= 1)
                    best = max(best, __shfl_xor_sync(kFull, best, d));
                // path position p <-> target offset tl - 1 - p
                for (int32_t p = lane; p < tl; p += 32)
                {
                    const int32_t off = tl - 1 - p;
                    int8_t v          = myers::st_insertion;
                    if (best >= 0 ? off == best : off == 0)
                        v = best >= 0 ? myers::st_match : myers::st_mismatch;
                    path[length + p] = v;
                }
                length += tl;
            }
            else
            {
                bool leaf = false;
                if (ql < P.full_myers_threshold)
                {
                    const int32_t n_words = (ql + 31) >> 5;
                    if (static_cast<int64_t>(tl + 1) * n_words <= P.leaf_elems)
                    {
                        length += leaf_path(P, pat, target + e.z, tl, query + e.x, ql, e.x, lpv, lmv, lsc, path + length, lane);
                        my_cells += static_cast<unsigned long long>(ql) * tl;
                        leaf = true;
                    }
                }
                if (!leaf)
                {
                    const int32_t qm = e.x + ql / 2;
                    // hirschberg_myers_compute_target_mid_warp (:412-481)
                    score_pass(P, pat, target + e.z, tl, qm - e.x, e.x, false, cpv, cmv, fwd, lane);
                    score_pass(P, pat, target + e.z, tl, e.y - qm, qlen_all - e.y, true, cpv, cmv, rev, lane);
                    my_cells += static_cast<unsigned long long>(ql) * tl;
                    int32_t cur_min = INT32_MAX, mid = 0;
                    for (int32_t t = lane; t <= tl; t += 32)
                    {
                        const int32_t sum = fwd[t] + rev[tl - t];
                        if (sum < cur_min)
                        {
                            cur_min = sum;
                            mid     = t;
                        }
                    }
#pragma unroll
                    for (int32_t i = 16; i > 0; i >>= 1)
                    {
                        const int32_t mv2 = __shfl_down_sync(kFull, cur_min, i);
                        const int32_t mp  = __shfl_down_sync(kFull, mid, i);
                        if (mv2 < cur_min)
                        {
                            cur_min = mv2;
                            mid     = mp;
                        }
                    }
                    mid              = __shfl_sync(kFull, mid, 0);
                    const int32_t tm = e.z + mid;
                    if (sp + 2 > kStackEntries)
                    {
                        success = false;
                    }
                    else
                    {
                        if (lane == 0)
                        {
                            s_stack[sp]     = make_int4(e.x, qm, e.z, tm);
                            s_stack[sp + 1] = make_int4(qm, e.y, tm, e.w);
                        }
                        sp += 2;
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
        }
        if (!success)
            length = 0;
    }
    if (lane == 0)
    {
        P.result_lengths[a] = length;
        if (my_cells)
            atomicAdd(P.cells, my_cells);
    }
}


// ---- AlignerGlobalUkkonen (cudaaligner/src/ukkonen_gpu.cu:62-262, aligner_global_ukkonen.cpp): banded unit-cost NW with the fixed
// band parameter p = 100, int16 scores ("infinity" = 32766), in the reference's slot coordinates (k, l) = ((j - i + p) / 2, i + j)
// because the slot layout is observable (boundary values in never-computed slots, truncating neighbour mapping in the backtrace).
// This repo's shape: one CTA per alignment, one anti-diagonal l per barrier; the last two anti-diagonals live in shared memory, a
// slot is written once (computed value or its initial value) as part of a contiguous int16 row of the band matrix, so the separate
// initialisation pass of the reference (a second sweep over the whole matrix) does not exist.
struct UkkonenParams
{
    const char* seqs;
    const int32_t* seq_lengths;
    int32_t max_len;
    int32_t n_alignments;
    int32_t max_result_length;
    int8_t* results;
    int32_t* result_lengths;
    int16_t* scores;      // [n][matrix_elems]
    int64_t matrix_elems; // ukkonen_max_score_matrix_size (ukkonen_gpu.cu:327-338)
    int32_t p;
    int32_t bw_capacity;  // slots per anti-diagonal the shared memory holds
    unsigned long long* cells;
};

constexpr int32_t kUkkMax = 32766;

__global__ void ukkonen_align_kernel(const UkkonenParams P)
{
    extern __shared__ int16_t s_diag[]; // [3][bw_capacity]
    const int32_t a = blockIdx.x;
    if (a >= P.n_alignments)
        return;
    int32_t m          = P.seq_lengths[2 * a] + 1;
    int32_t n          = P.seq_lengths[2 * a + 1] + 1;
    const char* query  = P.seqs + static_cast<int64_t>(2 * a) * P.max_len;
    const char* target = P.seqs + static_cast<int64_t>(2 * a + 1) * P.max_len;
    int8_t ins = myers::st_insertion, del = myers::st_deletion;
    if (m > n)
    {
        const int32_t t = m;
        m               = n;
        n               = t;
        const char* c   = query;
        query           = target;
        target          = c;
        ins             = myers::st_deletion;
        del             = myers::st_insertion;
    }
    const int32_t p         = P.p;
    const int32_t bw        = (1 + n - m + 2 * p + 1) / 2;
    const int32_t cols      = n + m;
    const int32_t kmax_odd  = (n - m + 2 * p - 1) / 2 + 1;
    const int32_t kmax_even = (n - m + 2 * p) / 2 + 1;
    int16_t* const S        = P.scores + static_cast<int64_t>(a) * P.matrix_elems;
    const int32_t cap       = P.bw_capacity;
    for (int32_t l = 0; l < cols; l++)
    {
        int16_t* const cur      = s_diag + (l % 3) * cap;
        const int16_t* const p1 = s_diag + ((l + 2) % 3) * cap;
        const int16_t* const p2 = s_diag + ((l + 1) % 3) * cap;
        const bool even_type    = ((l + p) & 1) == 0;
        const int32_t kmax      = even_type ? kmax_even : kmax_odd;
        for (int32_t k = threadIdx.x; k < bw; k += blockDim.x)
        {
            const int32_t j = k - (p + l) / 2 + l;
            const int32_t i = l - j;
            int32_t v       = (i == 0) ? j : (j == 0 ? i : kUkkMax);
            if (k < kmax)
            {
                const int32_t dd   = even_type ? 2 * k : 2 * k + 1;
                const int32_t lmin = abs(dd - p);
                const int32_t lmax = dd <= p ? 2 * (m - p + dd) + lmin : 2 * min(m, n - dd + p) + lmin;
                if (lmin + 1 <= l && l < lmax)
                {
                    const int32_t diag = l < 2 ? kUkkMax : p2[k] + (query[i - 1] == target[j - 1] ? 0 : 1);
                    int32_t left, above;
                    if (even_type)
                    {
                        left  = (k < 1 || l < 1) ? kUkkMax : p1[k - 1] + 1;
                        above = l < 1 ? kUkkMax : p1[k] + 1;
                    }
                    else
                    {
                        left  = l < 1 ? kUkkMax : p1[k] + 1;
                        above = (l < 1 || k + 1 >= bw) ? kUkkMax : p1[k + 1] + 1;
                    }
                    v = min(static_cast<int32_t>(static_cast<int16_t>(diag)),
                            min(static_cast<int32_t>(static_cast<int16_t>(left)), static_cast<int32_t>(static_cast<int16_t>(above))));
                }
            }
            cur[k]                             = static_cast<int16_t>(v);
            S[k + static_cast<int64_t>(bw) * l] = static_cast<int16_t>(v);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        auto get = [&](int32_t i, int32_t j) -> int32_t {
            const int32_t k = (j - i + p) / 2; // truncates towards zero (to_band_indices, ukkonen_gpu.cu:54-59)
            const int32_t l = j + i;
            return (k < 0 || k >= bw || l < 0 || l >= cols) ? kUkkMax : S[k + static_cast<int64_t>(bw) * l];
        };
        int8_t* const path = P.results + static_cast<int64_t>(a) * P.max_result_length;
        int32_t i = m - 1, j = n - 1, pos = 0;
        int32_t my = get(i, j);
        while (i > 0 && j > 0)
        {
            const int32_t above = get(i - 1, j), diag = get(i - 1, j - 1), left = get(i, j - 1);
            int8_t r;
            if (left + 1 == my)
            {
                r  = ins;
                my = left;
                --j;
            }
            else if (above + 1 == my)
            {
                r  = del;
                my = above;
                --i;
            }
            else
            {
                r  = (diag == my) ? myers::st_match : myers::st_mismatch;
                my = diag;
                --i;
                --j;
            }
            path[pos++] = r;
        }
        for (; i > 0; --i)
            path[pos++] = del;
        for (; j > 0; --j)
            path[pos++] = ins;
        P.result_lengths[a] = pos;
        atomicAdd(P.cells, static_cast<unsigned long long>(bw) * static_cast<unsigned long long>(cols));
    }
}

} // namespace galign
} // namespace gwb200
