// TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's banded Myers / Ukkonen global aligner.
//
// Nothing in the product path may link, import or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// baseline legs do, and only as a checker.
//
// Parity status: PINNED (tests/test_oracle_aligner.py) against the reference's own known-answer tests:
// Test_AlignerGlobal.cpp:73-157 (CIGAR + edit distance table, MyersBanded rows, incl. the empty-sequence cases),
// Test_ApproximateBandedMyers.cpp:48-172 (bw=7 corner-case CIGARs, bandwidth monotonicity, distance 23), the pygenomeworks
// CIGAR list (pygenomeworks/test/test_cudaaligner_bindings.py) and, on the GPU box, against the unmodified reference
// kernels (oracle/_ref/libgwref.so).
//
// The restatement is literal at the granularity the device code works at: 32-bit words, warp iterations of 32 words,
// per-lane carries (cudaaligner/src/myers_gpu.cu:78-255, 444-1032), the host-side bandwidth clamp
// (cudaaligner/src/aligner_global_myers_banded.cpp:155-258) and result decoding (:376-430).

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace
{

typedef uint32_t WordType;
constexpr int32_t warp_size = 32;
constexpr int32_t word_size = 32;

inline int32_t ceiling_divide(int32_t a, int32_t b) { return (a + b - 1) / b; }

enum AlignmentState : int8_t
{
    match = 0,
    mismatch,
    insertion,
    deletion
};

// device_matrix_view: column-major data[i + n_rows * j] (batched_device_matrices.cuh:61-76)
template <typename T>
struct Mat
{
    int32_t rows = 0, cols = 0;
    std::vector<T> data;
    void reshape(int32_t r, int32_t c)
    {
        rows = r;
        cols = c;
        if (static_cast<int64_t>(data.size()) < static_cast<int64_t>(r) * c)
            data.resize(static_cast<int64_t>(r) * c);
    }
    T& operator()(int32_t i, int32_t j) { return data[i + static_cast<int64_t>(rows) * j]; }
    const T& operator()(int32_t i, int32_t j) const { return data[i + static_cast<int64_t>(rows) * j]; }
};

// One warp iteration: n active lanes (warp_mask = low n bits).
struct Lanes
{
    int32_t n;
    WordType pv[32], mv[32], eq[32];
    int32_t carry_in[32];
};

// warp_add_sync, myers_gpu.cu:104-130: multi-word addition over the active lanes, carry out of the top lane dropped
void warp_add(int32_t n, const WordType* a, const WordType* b, WordType* r)
{
    uint64_t carry = 0;
    for (int32_t l = 0; l < n; l++)
    {
        uint64_t s = static_cast<uint64_t>(a[l]) + static_cast<uint64_t>(b[l]) + carry;
        r[l]       = static_cast<WordType>(s);
        carry      = s >> 32;
    }
}
// warp_leftshift_sync, :78-89
void warp_shl(int32_t n, WordType* v)
{
    WordType in = 0;
    for (int32_t l = 0; l < n; l++)
    {
        WordType out = v[l] >> 31;
        v[l]         = (v[l] << 1) | in;
        in           = out;
    }
}
// warp_rightshift_sync, :91-102
void warp_shr(int32_t n, WordType* v)
{
    for (int32_t l = 0; l < n; l++)
    {
        WordType x = (l + 1 < n) ? (v[l + 1] << 31) : 0;
        v[l]       = (v[l] >> 1) | x;
    }
}

// myers_advance_block / myers_advance_block2, :132-194. carry_out_x[l] for highest_bit[l]; carry_out_y[l] for highest_bit[l] << 1.
void advance_block(Lanes& L, const WordType* highest_bit, int32_t* carry_out_x, int32_t* carry_out_y)
{
    const int32_t n = L.n;
    WordType xv[32], eqm[32], a[32], xh[32], ph[32], mh[32];
    for (int32_t l = 0; l < n; l++)
    {
        xv[l]  = L.eq[l] | L.mv[l];
        eqm[l] = L.eq[l];
        if (L.carry_in[l] < 0)
            eqm[l] |= WordType(1);
        a[l] = eqm[l] & L.pv[l];
    }
    warp_add(n, a, L.pv, xh);
    for (int32_t l = 0; l < n; l++)
    {
        xh[l] = (xh[l] ^ L.pv[l]) | eqm[l];
        ph[l] = L.mv[l] | (~(xh[l] | L.pv[l]));
        mh[l] = L.pv[l] & xh[l];
        carry_out_x[l] = ((ph[l] & highest_bit[l]) == 0 ? 0 : 1) - ((mh[l] & highest_bit[l]) == 0 ? 0 : 1);
        if (carry_out_y)
        {
            const WordType hb2 = highest_bit[l] << 1;
            carry_out_y[l]     = ((ph[l] & hb2) == 0 ? 0 : 1) - ((mh[l] & hb2) == 0 ? 0 : 1);
        }
    }
    warp_shl(n, ph);
    warp_shl(n, mh);
    for (int32_t l = 0; l < n; l++)
    {
        if (L.carry_in[l] < 0)
            mh[l] |= WordType(1);
        if (L.carry_in[l] > 0)
            ph[l] |= WordType(1);
        L.pv[l] = mh[l] | (~(xv[l] | ph[l]));
        L.mv[l] = ph[l] & xv[l];
    }
}

struct Problem
{
    const char* query;
    const char* target;
    int32_t query_size, target_size;
    Mat<WordType> qp; // query patterns [n_words x 4], char order A,C,T,G via (x>>1)&3 (:215, :938-947)
    Mat<WordType> pv, mv;
    Mat<int32_t> score;
};

// get_query_pattern, :210-241
WordType get_query_pattern(const Mat<WordType>& qp, int32_t idx, int32_t query_begin_offset, char x)
{
    const int32_t char_idx   = (x >> 1) & 0x3;
    const int32_t idx_offset = query_begin_offset / word_size;
    const int32_t shift      = query_begin_offset % word_size;
    WordType r               = qp(idx + idx_offset, char_idx);
    if (shift != 0)
    {
        r >>= shift;
        if (idx + idx_offset + 1 < qp.rows)
            r |= qp(idx + idx_offset + 1, char_idx) << (word_size - shift);
    }
    return r;
}

// myers_compute_scores_horizontal_band_impl, :629-674
void horizontal_band(Problem& P, int32_t t_begin, int32_t t_end, int32_t width, int32_t n_words, int32_t pattern_idx_offset)
{
    const int32_t n_warp_iterations = ceiling_divide(n_words, warp_size) * warp_size;
    for (int32_t t = t_begin; t < t_end; ++t)
    {
        int32_t lane0_carry = 1; // worst case for the top border of the band
        for (int32_t base = 0; base < n_warp_iterations; base += warp_size)
        {
            if (base >= n_words)
                break;
            Lanes L;
            const bool full = (base / warp_size) < (n_words / warp_size);
            L.n             = full ? 32 : (n_words % warp_size);
            WordType hb[32];
            for (int32_t l = 0; l < L.n; l++)
            {
                const int32_t idx = base + l;
                L.pv[l]           = P.pv(idx, t - 1);
                L.mv[l]           = P.mv(idx, t - 1);
                hb[l]             = WordType(1) << (idx == (n_words - 1) ? width - (n_words - 1) * word_size - 1 : word_size - 1);
                L.eq[l]           = get_query_pattern(P.qp, idx, pattern_idx_offset, P.target[t - 1]);
                L.carry_in[l]     = (l == 0) ? lane0_carry : 0;
            }
            int32_t cx[32];
            advance_block(L, hb, cx, nullptr);
            for (int32_t l = 0; l < L.n; l++)
            {
                const int32_t idx = base + l;
                P.score(idx, t)   = P.score(idx, t - 1) + cx[l];
                P.pv(idx, t)      = L.pv[l];
                P.mv(idx, t)      = L.mv[l];
            }
            // carry hand-off only when the chunk is a full 32 lanes (:664-667)
            lane0_carry = full ? cx[31] : 0;
        }
    }
}

// myers_compute_scores_diagonal_band_impl, :676-751
void diagonal_band(Problem& P, int32_t t_begin, int32_t t_end, int32_t band_width, int32_t n_words_band, int32_t pattern_idx_offset)
{
    const int32_t n_warp_iterations = ceiling_divide(n_words_band, warp_size) * warp_size;
    for (int32_t t = t_begin; t < t_end; ++t)
    {
        int32_t lane0_carry = 1;
        for (int32_t base = 0; base < n_warp_iterations; base += warp_size)
        {
            if (base >= n_words_band)
                break;
            Lanes L;
            const bool full = (base / warp_size) < (n_words_band / warp_size);
            L.n             = full ? 32 : (n_words_band % warp_size);
            WordType hb[32];
            for (int32_t l = 0; l < L.n; l++)
            {
                L.pv[l] = P.pv(base + l, t - 1);
                L.mv[l] = P.mv(base + l, t - 1);
            }
            warp_shr(L.n, L.pv);
            warp_shr(L.n, L.mv);
            if (full)
            {
                const int32_t idx = base + 31;
                if (idx < n_words_band - 1)
                {
                    L.pv[31] |= P.pv(idx + 1, t - 1) << (word_size - 1);
                    L.mv[31] |= P.mv(idx + 1, t - 1) << (word_size - 1);
                }
            }
            WordType ddb[32];
            for (int32_t l = 0; l < L.n; l++)
            {
                const int32_t idx = base + l;
                L.eq[l]           = get_query_pattern(P.qp, idx, pattern_idx_offset + t - t_begin + 1, P.target[t - 1]);
                hb[l]             = WordType(1) << (idx == (n_words_band - 1) ? band_width - (n_words_band - 1) * word_size - 2 : word_size - 2);
                ddb[l]            = hb[l] << 1;
                if (idx == n_words_band - 1)
                {
                    L.pv[l] |= ddb[l];
                    L.mv[l] &= ~ddb[l];
                }
                L.carry_in[l] = (l == 0) ? lane0_carry : 0;
            }
            int32_t cx[32], cy[32];
            advance_block(L, hb, cx, cy);
            for (int32_t l = 0; l < L.n; l++)
            {
                const int32_t idx        = base + l;
                const int32_t delta_down = ((L.pv[l] & ddb[l]) == 0 ? 0 : 1) - ((L.mv[l] & ddb[l]) == 0 ? 0 : 1);
                P.score(idx, t)          = P.score(idx, t - 1) + cx[l] + delta_down;
                P.pv(idx, t)             = L.pv[l];
                P.mv(idx, t)             = L.mv[l];
            }
            lane0_carry = full ? cy[31] : 0;
        }
    }
}

// myers_compute_scores_edit_dist_banded, :753-846
void compute_scores_banded(Problem& P, int32_t& diagonal_begin, int32_t& diagonal_end, int32_t band_width, int32_t n_words_band, int32_t p)
{
    const int32_t target_size = P.target_size, query_size = P.query_size;
    for (int32_t idx = 0; idx < n_words_band; idx++)
    {
        P.pv(idx, 0)    = ~WordType(0);
        P.mv(idx, 0)    = 0;
        P.score(idx, 0) = std::min((idx + 1) * word_size, band_width);
    }
    if (band_width >= query_size)
    {
        diagonal_begin = target_size + 1;
        diagonal_end   = target_size + 1;
        horizontal_band(P, 1, target_size + 1, query_size, n_words_band, 0);
    }
    else
    {
        const int32_t symmetric_band = (band_width - std::min(1 + 2 * p + std::abs(target_size - query_size), query_size) == 0) ? 1 : 0;
        diagonal_begin               = query_size < target_size ? target_size - query_size + p + 2 : p + 2 + (1 - symmetric_band);
        diagonal_end                 = query_size < target_size ? query_size - p + symmetric_band : query_size - (query_size - target_size) - p + 1;
        horizontal_band(P, 1, diagonal_begin, band_width, n_words_band, 0);
        diagonal_band(P, diagonal_begin, diagonal_end, band_width, n_words_band, 0);
        horizontal_band(P, diagonal_end, target_size + 1, band_width, n_words_band, query_size - band_width);
    }
}

// get_myers_score, :243-255
int32_t get_myers_score(int32_t i, int32_t j, const Problem& P, WordType last_entry_mask)
{
    const int32_t word_idx = (i - 1) / word_size;
    const int32_t bit_idx  = (i - 1) % word_size;
    int32_t s              = P.score(word_idx, j);
    WordType mask          = (~WordType(1)) << bit_idx;
    if (word_idx == P.score.rows - 1)
        mask &= last_entry_mask;
    s -= __builtin_popcount(mask & P.pv(word_idx, j));
    s += __builtin_popcount(mask & P.mv(word_idx, j));
    return s;
}

struct Rle
{
    std::vector<int8_t> path;
    std::vector<int32_t> count;
    int8_t prev_r   = -1;
    int32_t r_count = 0;
    void flush_if_change(int8_t r)
    {
        if (prev_r != r)
        {
            if (prev_r != -1)
            {
                path.push_back(prev_r);
                count.push_back(r_count);
            }
            prev_r  = r;
            r_count = 0;
        }
    }
};

// myers_backtrace_banded, :444-627. Emits the RLE path end -> start (the host reverses it).
void backtrace_banded(const Problem& P, Rle& R, int32_t diagonal_begin, int32_t diagonal_end, int32_t band_width, int32_t target_size,
                      int32_t query_size)
{
    (void)query_size;
    const int32_t out_of_band = INT32_MAX - 1;
    int32_t i                 = band_width;
    int32_t j                 = target_size;
    const WordType last_entry_mask = band_width % word_size != 0 ? (WordType(1) << (band_width % word_size)) - 1 : ~WordType(0);
    const int32_t last_diagonal_score = diagonal_end < 2 ? out_of_band : get_myers_score(1, diagonal_end - 2, P, last_entry_mask) + 2;
    int32_t myscore                   = i > 0 ? P.score((i - 1) / word_size, j) : 0;
    while (j >= diagonal_end)
    {
        int8_t r            = 0;
        const int32_t above = i <= 1 ? (last_diagonal_score + j - diagonal_end) : get_myers_score(i - 1, j, P, last_entry_mask);
        const int32_t diag  = i <= 1 ? (last_diagonal_score + j - 1 - diagonal_end) : get_myers_score(i - 1, j - 1, P, last_entry_mask);
        const int32_t left  = i < 1 ? (last_diagonal_score + j - 1 - diagonal_end) : get_myers_score(i, j - 1, P, last_entry_mask);
        if (left + 1 == myscore)
        {
            r       = insertion;
            myscore = left;
            --j;
        }
        else if (above + 1 == myscore)
        {
            r       = deletion;
            myscore = above;
            --i;
        }
        else
        {
            r       = (diag == myscore ? match : mismatch);
            myscore = diag;
            --i;
            --j;
        }
        R.flush_if_change(r);
        ++R.r_count;
    }
    while (j >= diagonal_begin)
    {
        int8_t r            = 0;
        const int32_t above = i <= 1 ? out_of_band : get_myers_score(i - 1, j, P, last_entry_mask);
        const int32_t diag  = i <= 0 ? j - 1 : get_myers_score(i, j - 1, P, last_entry_mask);
        const int32_t left  = i >= band_width ? out_of_band : get_myers_score(i + 1, j - 1, P, last_entry_mask);
        if (left + 1 == myscore)
        {
            r       = insertion;
            myscore = left;
            ++i;
            --j;
        }
        else if (above + 1 == myscore)
        {
            r       = deletion;
            myscore = above;
            --i;
        }
        else
        {
            r       = (diag == myscore ? match : mismatch);
            myscore = diag;
            --j;
        }
        R.flush_if_change(r);
        ++R.r_count;
    }
    while (i > 0 && j > 0)
    {
        int8_t r            = 0;
        const int32_t above = i == 1 ? j : get_myers_score(i - 1, j, P, last_entry_mask);
        const int32_t diag  = i == 1 ? j - 1 : get_myers_score(i - 1, j - 1, P, last_entry_mask);
        const int32_t left  = i > band_width ? out_of_band : get_myers_score(i, j - 1, P, last_entry_mask);
        if (left + 1 == myscore)
        {
            r       = insertion;
            myscore = left;
            --j;
        }
        else if (above + 1 == myscore)
        {
            r       = deletion;
            myscore = above;
            --i;
        }
        else
        {
            r       = (diag == myscore ? match : mismatch);
            myscore = diag;
            --i;
            --j;
        }
        R.flush_if_change(r);
        ++R.r_count;
    }
    if (i > 0)
    {
        R.flush_if_change(deletion);
        R.r_count += i;
    }
    if (j > 0)
    {
        R.flush_if_change(insertion);
        R.r_count += j;
    }
    if (R.r_count != 0)
    {
        R.path.push_back(R.prev_r);
        R.count.push_back(R.r_count);
    }
}

struct Result
{
    int32_t status     = 1; // cudaaligner::StatusType::uninitialized
    int32_t is_optimal = 0;
    std::vector<int8_t> actions;   // query-start -> query-end order (already reversed like sync_alignments does)
    std::vector<int32_t> runs;
    int64_t cells = 0;
};

// myers_banded_kernel body for one alignment task, :894-1031, then sync_alignments() decoding
// (aligner_global_myers_banded.cpp:402-427).
Result align_one(const char* query, int32_t query_size, const char* target, int32_t target_size, int32_t max_bandwidth,
                 int64_t max_elements_per_matrix)
{
    Result res;
    if (max_bandwidth - 1 < std::abs(target_size - query_size) && query_size != 0 && target_size != 0)
        return res; // path_starts = -1, metadata not optimal -> status stays uninitialized
    if (target_size == 0 || query_size == 0)
    {
        if (query_size == 0 && target_size == 0)
        {
            res.status     = 0;
            res.is_optimal = 1;
            return res;
        }
        res.status     = 0;
        res.is_optimal = 1;
        res.actions.push_back(query_size == 0 ? insertion : deletion);
        res.runs.push_back(query_size + target_size);
        return res;
    }
    Problem P;
    P.query       = query;
    P.target      = target;
    P.query_size  = query_size;
    P.target_size = target_size;
    const int32_t n_words = ceiling_divide(query_size, word_size);
    P.qp.reshape(n_words, 4);
    const char chars[4] = {'A', 'C', 'T', 'G'};
    for (int32_t idx = 0; idx < n_words; idx++)
    {
        for (int32_t c = 0; c < 4; c++)
        {
            // myers_generate_query_pattern, :196-208
            const int32_t offset = idx * word_size;
            const int32_t max_i  = std::min(query_size - offset, word_size);
            WordType r           = 0;
            for (int32_t i = 0; i < max_i; ++i)
                if (chars[c] == query[i + offset])
                    r |= (WordType(1) << i);
            P.qp(idx, c) = r;
        }
    }
    int32_t max_distance_estimate = std::max(1, std::abs(target_size - query_size) + std::min(target_size, query_size) / 20);
    int32_t diagonal_begin = -1, diagonal_end = -1, band_width = 0;
    while (1)
    {
        int32_t p              = std::min(std::min(target_size, query_size), (max_distance_estimate - std::abs(target_size - query_size)) / 2);
        int32_t band_width_new = std::min(1 + 2 * p + std::abs(target_size - query_size), query_size);
        if (band_width_new % word_size == 1 && band_width_new != query_size)
        {
            p += 1;
            band_width_new = std::min(1 + 2 * p + std::abs(target_size - query_size), query_size);
        }
        if (band_width_new > max_bandwidth)
        {
            band_width_new = max_bandwidth;
            p              = (band_width_new - 1 - std::abs(target_size - query_size)) / 2;
        }
        const int32_t n_words_band = ceiling_divide(band_width_new, word_size);
        if (static_cast<int64_t>(n_words_band) * static_cast<int64_t>(target_size + 1) > max_elements_per_matrix)
        {
            band_width = -band_width;
            break;
        }
        band_width = band_width_new;
        P.pv.reshape(n_words_band, target_size + 1);
        P.mv.reshape(n_words_band, target_size + 1);
        P.score.reshape(n_words_band, target_size + 1);
        diagonal_begin = -1;
        diagonal_end   = -1;
        res.cells += static_cast<int64_t>(band_width) * target_size;
        compute_scores_banded(P, diagonal_begin, diagonal_end, band_width, n_words_band, p);
        const int32_t cur_edit_distance = n_words_band > 0 ? P.score(n_words_band - 1, target_size) : target_size;
        if (cur_edit_distance <= max_distance_estimate || band_width == query_size)
            break;
        if (band_width == max_bandwidth)
        {
            band_width = -band_width;
            break;
        }
        max_distance_estimate *= 2;
    }
    if (band_width != 0)
    {
        Rle R;
        backtrace_banded(P, R, diagonal_begin, diagonal_end, std::abs(band_width), target_size, query_size);
        res.is_optimal = band_width > 0 ? 1 : 0;
        if (!R.path.empty())
        {
            res.status = 0;
            res.actions.assign(R.path.rbegin(), R.path.rend());
            res.runs.assign(R.count.rbegin(), R.count.rend());
        }
    }
    return res;
}

} // namespace

extern "C" {

// Host-side bandwidth clamp of FixedBandAligner::add_alignment (aligner_global_myers_banded.cpp:174-178) followed by the
// device task. actions/runs must hold query_length + target_length entries. Returns the number of RLE entries.
int32_t oracle_myers_banded_align(const char* query, int32_t query_length, const char* target, int32_t target_length, int32_t max_bandwidth,
                                  int64_t max_elements_per_matrix, int32_t* status, int32_t* is_optimal, int8_t* actions, int32_t* runs,
                                  int64_t* cells)
{
    if (max_bandwidth > query_length)
        max_bandwidth = (query_length % word_size == 1 ? query_length + 1 : query_length);
    if (max_elements_per_matrix <= 0)
        max_elements_per_matrix = INT64_MAX;
    Result r    = align_one(query, query_length, target, target_length, max_bandwidth, max_elements_per_matrix);
    *status     = r.status;
    *is_optimal = r.is_optimal;
    if (cells)
        *cells = r.cells;
    for (size_t k = 0; k < r.actions.size(); k++)
    {
        actions[k] = r.actions[k];
        runs[k]    = r.runs[k];
    }
    return static_cast<int32_t>(r.actions.size());
}

// Every cell of one band pass (for checking other formulations of the same band, tests/cpp/myers_skew_model.cpp):
// out[(i - 1) + band_width * j] = get_myers_score(i, j) for i = 1 .. band_width, j = 0 .. target_length.
int32_t oracle_myers_band_scores(const char* query, int32_t query_length, const char* target, int32_t target_length, int32_t band_width,
                                 int32_t p, int32_t* out, int32_t* diagonal_begin, int32_t* diagonal_end)
{
    Problem P;
    P.query       = query;
    P.target      = target;
    P.query_size  = query_length;
    P.target_size = target_length;
    const int32_t n_words = ceiling_divide(query_length, word_size);
    P.qp.reshape(n_words, 4);
    const char chars[4] = {'A', 'C', 'T', 'G'};
    for (int32_t idx = 0; idx < n_words; idx++)
        for (int32_t c = 0; c < 4; c++)
        {
            const int32_t offset = idx * word_size;
            const int32_t max_i  = std::min(query_length - offset, word_size);
            WordType r           = 0;
            for (int32_t i = 0; i < max_i; ++i)
                if (chars[c] == query[i + offset])
                    r |= (WordType(1) << i);
            P.qp(idx, c) = r;
        }
    const int32_t n_words_band = ceiling_divide(band_width, word_size);
    P.pv.reshape(n_words_band, target_length + 1);
    P.mv.reshape(n_words_band, target_length + 1);
    P.score.reshape(n_words_band, target_length + 1);
    int32_t db = -1, de = -1;
    compute_scores_banded(P, db, de, band_width, n_words_band, p);
    *diagonal_begin = db;
    *diagonal_end   = de;
    const WordType last_entry_mask = band_width % word_size != 0 ? (WordType(1) << (band_width % word_size)) - 1 : ~WordType(0);
    for (int32_t j = 0; j <= target_length; j++)
        for (int32_t i = 1; i <= band_width; i++)
            out[(i - 1) + static_cast<int64_t>(band_width) * j] = get_myers_score(i, j, P, last_entry_mask);
    return 0;
}

} // extern "C"
