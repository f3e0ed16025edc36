// gw-b200 POA device code, third generation (sm_100a): one warp per window, persistent CTAs, 8 score cells per lane and chunk.
//
// Same behavioural contract as poa_kernels.cuh / poa_kernels_v2.cuh (identical consensus / coverage / MSA / status to the
// reference kernels, cudapoa/src/cudapoa_kernels.cuh:76-542). What changed against v2, and why (profiles/r01_*: 59 % of the
// v2 stall samples were warps parked at CTA barriers, three of four warps idle through traceback / graph update / sort):
//
//   scheduling   one warp owns one window from the first read to the consensus; the grid is persistent (resident CTAs pull
//                window indices from an atomic counter), so every resident warp always has work of its own, no CTA barrier
//                exists anywhere, and batches that are not a multiple of the residency do not pay a wave quantum.
//   DP rows      every lane owns 8 cells of the row per chunk of 256 columns (int16: one 16-byte unit, the rows leave as 8-wide
//                vectors; int32: two units), so that the per-chunk costs are paid once per 256 columns. Chunks are taken left
//                to right with the horizontal carry passed on directly (no two-phase fold). The predecessor rows are read as
//                LDS.128 from a shared-memory ring of the most recent rows (always: the host sizes the pool for it), the
//                value beyond the unit comes from the right neighbour by shuffle (no bank-conflicting scalar load), the
//                substitution test is one xor per four read characters and one predicate-producing logic op per cell,
//                per-row graph metadata is prepared 32 rows ahead by the lanes and handed over as one 16-byte shared-memory
//                record per row.
//   write-back   the finished row is written once into its ring slot and leaves for HBM as ONE bulk asynchronous copy
//                (cp.async.bulk.global.shared::cta, bulk-group completion) issued by lane 0: the lanes issue no global
//                stores in the row loop, and the ring slot is reused only after its bulk read has completed.
//   traceback    score tiles (32 rows x 64 columns), the rows' metadata and the next tile along the predicted path are
//                fetched by bulk asynchronous copies that complete on an mbarrier (cp.async.bulk.shared::cluster.global);
//                the walk itself is v2's speculative warp-uniform walk (reference order cudapoa_nw_banded.cuh:440-534).
#pragma once
#include <type_traits>

#include "poa_kernels.cuh"
#include "poa_kernels_v2.cuh"

namespace gwb200
{
namespace poa
{

// ---- PTX wrappers: bulk asynchronous copies (TMA 1-D), bulk groups, mbarrier, proxy fences ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)), "r"(smem_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait()
{
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load_g2s(void* sdst, const void* gsrc, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t done = 0, spins = 0;
    while (!done)
    {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (++spins > (1u << 22))
            __trap(); // a copy that never completes is a bug: fail loudly instead of hanging the device
    }
}

struct V3Extra
{
    int32_t* work_counter; // next window index (persistent grid), reset by the host before every launch
    int32_t use_bulk;      // 1: rows leave the ring by cp.async.bulk (default); 0: per-lane vector stores (A/B switch)
    int32_t tb_tma;        // 1: traceback tiles by bulk asynchronous copies with prefetch (default); 0: lane loads (A/B switch)
    int32_t wavefront;     // 1: 32-bit score rows as a skewed wavefront (poa_kernels_v4.cuh); 0: dp_rows_v3 (default)
    int32_t max_group;     // chunks per straight-line group of the row loop: 4, 2 or 1 (see dp_rows_v3)
    int32_t row_fence;     // 1: fence.proxy.async in front of every general row that reads rows older than the ring (A/B switch)
};

// Static shared memory of the v3 kernel that is not part of the pool
struct V3Shared
{
    int4 rec[2][32];                  // per-row records of the current and the next 32-row group
    unsigned long long tile_bar[2];   // mbarriers of the two traceback tile buffers
    uint32_t tile_phase[2];           // their next wait parity (the barriers live as long as the CTA); must follow tile_bar
    uint32_t pad_[2];
};

// 1: a chunk whose entering horizontal run dominates all of its cells is decided by one ballot instead of the prefix scan.
// Measured and left off: the scan runs in ~57 % of the chunks of a wide band (right of the alignment path every cell is its left
// neighbour + gap), yet the extra ballot on the carry chain costs more than the scans it saves (C3 1456 vs 1503 windows/s, C2
// 37.3 k vs 38.6 k; profiles/r02_summary.md).
#ifndef GWB200_V3_PURE
#define GWB200_V3_PURE 0
#endif
#ifndef GWB200_V3_CPL32
#define GWB200_V3_CPL32 8 // 4: one 16-byte unit per lane and chunk (the round-2 layout, kept as the A/B partner)
#endif

template <typename ScoreT>
struct V3Cells
{
    // cells per lane and chunk: 8 (int16: one 16-byte unit, int32: two), so that the per-chunk costs -- neighbour shuffle, carry
    // shuffles and ballot, addressing, band tests -- are paid once per 256 columns
    static constexpr int32_t kCPL   = sizeof(ScoreT) == 2 ? 8 : GWB200_V3_CPL32;
    static constexpr int32_t kChunk = 32 * kCPL; // columns per chunk
};

// 16-byte unit <-> CPL int32 values
template <typename ScoreT, int32_t CPL>
__device__ __forceinline__ void unit_load(const ScoreT* p, int32_t* b)
{
    if constexpr (sizeof(ScoreT) == 4)
    {
        const int4 v = *reinterpret_cast<const int4*>(p);
        b[0]         = v.x;
        b[1]         = v.y;
        b[2]         = v.z;
        b[3]         = v.w;
        if constexpr (CPL == 8)
        {
            const int4 u = *reinterpret_cast<const int4*>(p + 4);
            b[4]         = u.x;
            b[5]         = u.y;
            b[6]         = u.z;
            b[7]         = u.w;
        }
    }
    else
    {
        // 8-byte aligned (band starts are multiples of 4 cells)
        const uint2 lo = *reinterpret_cast<const uint2*>(p);
        const uint2 hi = *reinterpret_cast<const uint2*>(p + 4);
        b[0]           = static_cast<int16_t>(lo.x & 0xffffu);
        b[1]           = static_cast<int32_t>(lo.x) >> 16;
        b[2]           = static_cast<int16_t>(lo.y & 0xffffu);
        b[3]           = static_cast<int32_t>(lo.y) >> 16;
        b[4]           = static_cast<int16_t>(hi.x & 0xffffu);
        b[5]           = static_cast<int32_t>(hi.x) >> 16;
        b[6]           = static_cast<int16_t>(hi.y & 0xffffu);
        b[7]           = static_cast<int32_t>(hi.y) >> 16;
    }
}

// unit = {left, s[0], ..., s[CPL-2]} (values already truncated to ScoreT range), 16-byte aligned destination
template <typename ScoreT, int32_t CPL>
__device__ __forceinline__ void unit_store(ScoreT* p, int32_t left, const int32_t (&s)[CPL])
{
    if constexpr (sizeof(ScoreT) == 4)
    {
        *reinterpret_cast<int4*>(p) = make_int4(left, s[0], s[1], s[2]);
        if constexpr (CPL == 8)
            *reinterpret_cast<int4*>(p + 4) = make_int4(s[3], s[4], s[5], s[6]);
    }
    else
    {
        uint4 v;
        v.x = (static_cast<uint32_t>(left) & 0xffffu) | (static_cast<uint32_t>(s[0]) << 16);
        v.y = (static_cast<uint32_t>(s[1]) & 0xffffu) | (static_cast<uint32_t>(s[2]) << 16);
        v.z = (static_cast<uint32_t>(s[3]) & 0xffffu) | (static_cast<uint32_t>(s[4]) << 16);
        v.w = (static_cast<uint32_t>(s[5]) & 0xffffu) | (static_cast<uint32_t>(s[6]) << 16);
        *reinterpret_cast<uint4*>(p) = v;
    }
}

// Closure of s[c] = max(h[c], s[c-1] + gap) over one chunk (CPL consecutive cells per lane) with carry-in `cin` for lane 0:
// the fixpoint of the reference's relaxation loop (cudapoa_nw_banded.cuh:362-390). `left` returns the closed value of the cell
// left of this lane's first cell (lane 0: cin). Fast exit when no lane's first cell is improved by its left neighbour.
template <int32_t CPL>
__device__ __forceinline__ void closure_seq(int32_t (&s)[CPL], int32_t cin, int32_t gap, int32_t lane, int32_t& left)
{
#pragma unroll
    for (int32_t k = 1; k < CPL; k++)
        s[k] = __viaddmax_s32(s[k - 1], gap, s[k]);
    int32_t nb = __shfl_up_sync(kFull, s[CPL - 1], 1);
    if (lane == 0)
        nb = cin;
    left = nb;
    if (__ballot_sync(kFull, nb + gap > s[0]) == 0u)
        return;
    const int32_t gl = CPL * gap;
    int32_t v        = s[CPL - 1] - gl * (lane + 1);
#pragma unroll
    for (int32_t d = 1; d < 32; d <<= 1)
    {
        const int32_t o = __shfl_up_sync(kFull, v, d);
        if (lane >= d)
            v = max(v, o);
    }
    int32_t excl = __shfl_up_sync(kFull, v, 1);
    excl         = (lane == 0) ? cin : max(cin, excl);
    const int32_t L = excl + gl * lane;
#pragma unroll
    for (int32_t k = 0; k < CPL; k++)
        s[k] = max(s[k], L + (k + 1) * gap);
    left = L;
}

// The max-plus prefix scan of one chunk (the grouped row code below needs it for a minority of its chunks)
template <int32_t CPL>
__device__ __forceinline__ int32_t chunk_scan(int32_t (&a)[CPL], const int32_t cin, const int32_t gap, const int32_t lane)
{
    const int32_t gl = CPL * gap;
    int32_t v        = a[CPL - 1] - gl * (lane + 1);
#pragma unroll
    for (int32_t d = 1; d < 32; d <<= 1)
    {
        const int32_t o = __shfl_up_sync(kFull, v, d);
        if (lane >= d)
            v = max(v, o);
    }
    int32_t excl    = __shfl_up_sync(kFull, v, 1);
    excl            = (lane == 0) ? cin : max(cin, excl);
    const int32_t L = excl + gl * lane;
#pragma unroll
    for (int32_t i = 0; i < CPL; i++)
        a[i] = max(a[i], L + (i + 1) * gap);
    return L;
}

// What a group of chunks needs to know about its row (fast rows: at most two predecessors, both in the ring)
template <typename ScoreT>
struct RowCtx
{
    const ScoreT* prow0; // ring rows of the predecessors
    const ScoreT* prow1;
    const uint8_t* rd;   // read + band start of the row
    ScoreT* drow;        // where the lanes write the row (ring slot, or the row in global memory without a ring)
    ScoreT* grow;        // second copy in global memory (lane stores instead of the bulk copy), or nullptr
    int32_t sh0, lim0, sh1, lim1;
    int32_t base, bw, gap, match, mismatch;
    uint32_t base4; // the row's base in every byte
};

// NJ consecutive chunks of one row, starting at chunk c0 (all of them exist: c0 + NJ <= number of chunks). Straight-line code:
// every load of the group is issued before the first dependent instruction needs it. cin: carry into the first chunk (closed
// value of the cell left of it), cleft: what is stored in that cell's place; both are updated for the next group.
template <typename ScoreT, int32_t NJ, bool TWO>
__device__ __forceinline__ void row_group_v3(const RowCtx<ScoreT>& cx, const int32_t c0, const int32_t lane, int32_t& cin, int32_t& cleft)
{
    constexpr int32_t kMin = min_score_of<ScoreT>();
    constexpr int32_t CPL  = V3Cells<ScoreT>::kCPL;
    constexpr int32_t CH   = V3Cells<ScoreT>::kChunk;
    const int32_t gap      = cx.gap;
    int32_t a[NJ][CPL];
    int32_t nbraw[NJ], outv[NJ], s0l0[NJ], leftv[NJ];
    uint32_t need[NJ];
    bool act[NJ];
    {
        // ---- loads: read characters and the predecessors' units of every chunk
        uint32_t w[NJ][CPL / 4];
        int32_t b0[NJ][CPL + 1], b1[TWO ? NJ : 1][CPL + 1];
        bool in0[NJ], in1[TWO ? NJ : 1];
#pragma unroll
        for (int32_t j = 0; j < NJ; j++)
        {
            const int32_t off = (c0 + j) * CH + CPL * lane;
            act[j]            = off < cx.bw;
#pragma unroll
            for (int32_t h = 0; h < CPL / 4; h++)
                w[j][h] = __ldg(reinterpret_cast<const uint32_t*>(cx.rd + off + 4 * h)) ^ cx.base4;
            const int32_t o = cx.sh0 + off;
            in0[j]          = act[j] && o <= cx.lim0;
            // lanes outside the predecessor's band load its first unit instead: their values are discarded by ok0 below
            unit_load<ScoreT, CPL>(cx.prow0 + (in0[j] ? o : 0), b0[j]);
            if (TWO)
            {
                const int32_t o1 = cx.sh1 + off;
                in1[j]           = act[j] && o1 <= cx.lim1;
                unit_load<ScoreT, CPL>(cx.prow1 + (in1[j] ? o1 : 0), b1[j]);
            }
        }
        // ---- the value right of each unit: first value of the right neighbour's unit, or an own load at the band / chunk edge
#pragma unroll
        for (int32_t j = 0; j < NJ; j++)
        {
            const int32_t off = (c0 + j) * CH + CPL * lane;
            {
                const int32_t o   = cx.sh0 + off;
                const int32_t nbv = __shfl_down_sync(kFull, b0[j][0], 1);
                const bool nb_in  = lane < 31 && (off + CPL) < cx.bw && (o + CPL) <= cx.lim0;
                b0[j][CPL]        = nbv;
                if (in0[j] && !nb_in)
                    b0[j][CPL] = cx.prow0[o + CPL];
            }
            if (TWO)
            {
                const int32_t o   = cx.sh1 + off;
                const int32_t nbv = __shfl_down_sync(kFull, b1[j][0], 1);
                const bool nb_in  = lane < 31 && (off + CPL) < cx.bw && (o + CPL) <= cx.lim1;
                b1[j][CPL]        = nbv;
                if (in1[j] && !nb_in)
                    b1[j][CPL] = cx.prow1[o + CPL];
            }
        }
        // ---- candidates (get_scores(), cudapoa_nw_banded.cuh:104-156: the band test is per 4 cells) and the lane-local closure
#pragma unroll
        for (int32_t j = 0; j < NJ; j++)
        {
            const int32_t off = (c0 + j) * CH + CPL * lane;
#pragma unroll
            for (int32_t i = 0; i < CPL; i++)
            {
                // w holds read ^ base in every byte: a zero byte is a match (one logic op with a predicate result per cell)
                const int32_t q  = ((w[j][i >> 2] & (0xffu << (8 * (i & 3)))) == 0u) ? cx.match : cx.mismatch;
                const bool ok0   = (i < 4) ? in0[j] : (in0[j] && (cx.sh0 + off + 4) <= cx.lim0);
                const int32_t t0 = static_cast<ScoreT>(__viaddmax_s32(b0[j][i + 1], gap, b0[j][i] + q));
                int32_t v        = ok0 ? t0 : kMin;
                if (TWO)
                {
                    const bool ok1   = (i < 4) ? in1[j] : (in1[j] && (cx.sh1 + off + 4) <= cx.lim1);
                    const int32_t t1 = static_cast<ScoreT>(__viaddmax_s32(b1[j][i + 1], gap, b1[j][i] + q));
                    v                = ok1 ? max(v, t1) : v;
                }
                a[j][i] = v;
            }
#pragma unroll
            for (int32_t i = 1; i < CPL; i++)
                a[j][i] = __viaddmax_s32(a[j][i - 1], gap, a[j][i]);
        }
    }
    // ---- what the carry resolution needs from the neighbours (independent of the carry: issued for all chunks at once)
#pragma unroll
    for (int32_t j = 0; j < NJ; j++)
    {
        const int32_t last_lane = min(31, (cx.bw - (c0 + j) * CH) / CPL - 1);
        nbraw[j]                = __shfl_up_sync(kFull, a[j][CPL - 1], 1);
        outv[j]                 = __shfl_sync(kFull, a[j][CPL - 1], last_lane);
        s0l0[j]                 = __shfl_sync(kFull, a[j][0], 0);
        need[j]                 = __ballot_sync(kFull, lane > 0 && act[j] && nbraw[j] + gap > a[j][0]);
    }
    // ---- carries, chunk by chunk (uniform control flow)
#pragma unroll
    for (int32_t j = 0; j < NJ; j++)
    {
        leftv[j] = (lane == 0) ? cleft : nbraw[j];
        if (need[j] != 0u || cin + gap > s0l0[j])
        {
            const int32_t last_lane = min(31, (cx.bw - (c0 + j) * CH) / CPL - 1);
            const int32_t gl        = CPL * gap;
            // A horizontal run crosses a lane boundary or enters from the chunk to the left. Right of the alignment path every
            // cell is its left neighbour + gap: the run that enters the chunk dominates all of it (about half of the chunks of
            // a wide band). The locally closed values satisfy a[i] >= a[i-1] + gap, so "every cell of the lane is dominated"
            // is one comparison of the lane's last cell and the whole chunk is decided by one ballot; only the remaining
            // chunks (the one the path crosses) run the max-plus prefix scan over the lanes.
            const int32_t bound = cin + gl * (lane + 1);
            if (GWB200_V3_PURE != 0 && __ballot_sync(kFull, act[j] && a[j][CPL - 1] > bound) == 0u)
            {
                const int32_t base = bound - gl;
                if (lane != 0)
                    leftv[j] = base;
#pragma unroll
                for (int32_t i = 0; i < CPL; i++)
                    a[j][i] = base + (i + 1) * gap;
                outv[j] = cin + gl * (last_lane + 1);
            }
            else
            {
                const int32_t L = chunk_scan<CPL>(a[j], cin, gap, lane);
                if (lane != 0)
                    leftv[j] = L;
                outv[j] = __shfl_sync(kFull, a[j][CPL - 1], last_lane);
            }
        }
        cin   = static_cast<ScoreT>(outv[j]);
        cleft = cin;
    }
    // ---- stores
#pragma unroll
    for (int32_t j = 0; j < NJ; j++)
    {
        const int32_t off = (c0 + j) * CH + CPL * lane;
        if (act[j])
        {
#pragma unroll
            for (int32_t i = 0; i < CPL; i++)
                a[j][i] = static_cast<ScoreT>(a[j][i]);
            unit_store<ScoreT, CPL>(cx.drow + off, static_cast<ScoreT>(leftv[j]), a[j]);
            if (cx.grow != nullptr)
                unit_store<ScoreT, CPL>(cx.grow + off, static_cast<ScoreT>(leftv[j]), a[j]);
        }
    }
}

// The DP rows of needlemanWunschBanded (cudapoa_nw_banded.cuh:266-405) for one warp; everything before and after them (band
// geometry, end cell, traceback) is in nw_banded_v3 below. Rows are written to `scores` (row r at r * stride) exactly once.
template <typename ScoreT, typename SizeT, bool BULK>
__device__ void dp_rows_v3(const Win<SizeT>& g, const int32_t graph_count, const uint8_t* __restrict__ read, const Band<ScoreT>& B,
                           const int32_t band_width, const int32_t max_column, const int32_t gap, const int32_t mismatch, const int32_t match,
                           int4* row_meta, uint8_t* pool, const int32_t pool_bytes, int4* srec, const int32_t max_group, const int32_t row_fence)
{
    constexpr int32_t kMin   = min_score_of<ScoreT>();
    constexpr int32_t CPL    = V3Cells<ScoreT>::kCPL;
    constexpr int32_t CH     = V3Cells<ScoreT>::kChunk;
    constexpr int32_t GC     = 16 / CPL; // chunks per group: 16 cells per lane in flight
    const int32_t lane       = threadIdx.x & 31;
    const int32_t stride     = B.stride;
    ScoreT* const scores     = B.scores;
    const int32_t rowbytes   = stride * static_cast<int32_t>(sizeof(ScoreT));
    const int32_t nchunks    = (band_width + CH - 1) / CH;
    ScoreT* const ring       = reinterpret_cast<ScoreT*>(pool);
    // the host sizes the pool for at least two rows of the widest band (v3_pool_bytes): the rows always go through the ring, so
    // every row access of the fast path is a shared-memory access the compiler can see as one (no generic pointers)
    const int32_t R          = min(pool_bytes / rowbytes, 64);
    if (R < 2)
        __trap(); // only reachable with the development override of the pool size
    constexpr bool use_ring  = true;
    constexpr bool bulk      = BULK;

    // row 0: scores[j] = j * gap (:269-272), also into ring slot 0
    for (int32_t j = lane; j < stride; j += 32)
    {
        const ScoreT v = static_cast<ScoreT>(j * gap);
        scores[j]      = v;
        if (use_ring)
            ring[j] = v;
    }

    // ---- per-row records, prepared one 32-row group ahead (lane k <-> row r0 + 32 + k) and handed over through shared memory
    //   x: band start | base << 16 | min(in-degree, 3) << 24 | "fast row" << 27   (fast: <= 2 predecessors, all in the ring)
    //   y: (band start - band start of predecessor 0) | (last in-band unit offset of predecessor 0) << 16
    //   z: same for predecessor 1
    //   w: ring distance of predecessor 0 | ring distance of predecessor 1 << 8
    // row_meta (global; end-cell search and traceback): {node, pred0 row, pred1 row, base | in-degree << 8 | sink << 16 | band start / 4 << 17}
    auto produce = [&](int32_t row, int32_t node, int32_t base, int32_t pc, int32_t oc, int32_t p0, int32_t p1, int4* dst) {
        const int32_t bs   = B.start(row);
        const int32_t bsp0 = B.start(p0);
        const int32_t bsp1 = B.start(p1);
        const int32_t d0   = row - p0;
        const int32_t d1   = row - p1;
        const bool fast    = use_ring && pc <= 2 && d0 < R && (pc < 2 || d1 < R);
        int4 r;
        r.x = bs | (base << 16) | (min(pc, 3) << 24) | ((fast ? 1 : 0) << 27);
        r.y = (bs - bsp0) | (min(band_width - kCPT, max_column - bsp0) << 16);
        r.z = (bs - bsp1) | (min(band_width - kCPT, max_column - bsp1) << 16);
        r.w = (d0 & 0xff) | ((d1 & 0xff) << 8);
        *dst          = r;
        row_meta[row] = make_int4(node, p0, p1, base | (pc << 8) | ((oc == 0 ? 1 : 0) << 16) | ((bs >> 2) << 17));
    };
    {
        const int32_t row = 1 + lane;
        if (row <= graph_count)
        {
            const int32_t node = g.sorted[row - 1];
            const int32_t pc   = g.in_cnt[node];
            const int32_t p0   = pc > 0 ? static_cast<int32_t>(g.pos[g.in_edge(node, 0)]) + 1 : 0;
            const int32_t p1   = pc > 1 ? static_cast<int32_t>(g.pos[g.in_edge(node, 1)]) + 1 : 0;
            produce(row, node, g.nodes[node], pc, g.out_cnt[node], p0, p1, &srec[lane]);
        }
    }
    __syncwarp();

    int32_t ring_slot = 0;
    ScoreT* rowp      = scores;
    const int32_t off0 = CPL * lane; // local index of this lane's unit in chunk 0

    for (int32_t r0 = 1; r0 <= graph_count; r0 += 32)
    {
        const int32_t nrows  = min(32, graph_count - r0 + 1);
        const bool have_next = r0 + 32 <= graph_count;
        const int32_t nrow   = r0 + 32 + lane;
        const bool nvalid    = have_next && nrow <= graph_count;
        const int4* rec_cur  = srec + (((r0 - 1) >> 5) & 1) * 32;
        int4* rec_nxt        = srec + ((((r0 - 1) >> 5) + 1) & 1) * 32;
        int32_t t_node = 0, t_base = 0, t_pc = 0, t_oc = 1, t_e0 = 0, t_e1 = 0;

        for (int32_t k = 0; k < nrows; k++)
        {
            // next group's metadata: three dependent global loads spread over the row iterations
            if (nvalid && ((0x10100401u >> k) & 1u) != 0u) // k in {0, 10, 20, 28}
            {
                if (k == 0)
                {
                    t_node = g.sorted[nrow - 1];
                }
                else if (k == 10)
                {
                    t_base = g.nodes[t_node];
                    t_pc   = g.in_cnt[t_node];
                    t_oc   = g.out_cnt[t_node];
                    t_e0   = g.in_edge(t_node, 0);
                    t_e1   = g.in_edge(t_node, 1);
                }
                else if (k == 20)
                {
                    t_e0 = t_pc > 0 ? static_cast<int32_t>(g.pos[t_e0]) + 1 : 0; // predecessor rows
                    t_e1 = t_pc > 1 ? static_cast<int32_t>(g.pos[t_e1]) + 1 : 0;
                }
                else if (k == 28)
                {
                    produce(nrow, t_node, t_base, t_pc, t_oc, t_e0, t_e1, &rec_nxt[lane]);
                }
            }
            const int32_t row  = r0 + k;
            const int4 rc      = rec_cur[k];
            const int32_t bs   = rc.x & 0xffff;
            const int32_t base = (rc.x >> 16) & 0xff;
            const int32_t pc   = (rc.x >> 24) & 0x7;
            ring_slot          = (ring_slot + 1 == R) ? 0 : ring_slot + 1;
            rowp += stride;
            ScoreT* const srow = ring + ring_slot * stride;
            ScoreT* const drow = use_ring ? srow : rowp; // where the lanes write the row

            int32_t local0, cin;
            if (rc.x & (1 << 27))
            {
                // ---- fast row: at most two predecessors, both in the shared-memory ring
                int32_t sl0 = ring_slot - (rc.w & 0xff);
                if (sl0 < 0)
                    sl0 += R;
                const ScoreT* prow0 = ring + sl0 * stride;
                const int32_t sh0   = rc.y & 0xffff;
                const int32_t lim0  = rc.y >> 16;
                const ScoreT* prow1 = prow0;
                int32_t sh1 = 0, lim1 = -1;
                if (pc == 2)
                {
                    int32_t sl1 = ring_slot - ((rc.w >> 8) & 0xff);
                    if (sl1 < 0)
                        sl1 += R;
                    prow1 = ring + sl1 * stride;
                    sh1   = rc.z & 0xffff;
                    lim1  = rc.z >> 16;
                }
                // column "-1" / first_element_prev_score (:293-326)
                int32_t first = 0;
                if (pc != 0)
                {
                    if (bs > kCPT && pc == 1)
                    {
                        first = kMin + gap;
                    }
                    else
                    {
                        // local 0 of a predecessor row holds the minimum unless that row's band starts at column 0 (only the
                        // first rows of a graph): no dependent shared-memory loads in the common case
                        int32_t penalty = kMin;
                        if (bs == sh0)
                            penalty = max(kMin, static_cast<int32_t>(prow0[0]));
                        if (pc == 2 && bs == sh1)
                            penalty = max(penalty, static_cast<int32_t>(prow1[0]));
                        first = penalty + gap;
                    }
                }
                local0 = (bs == 0) ? (pc == 0 ? gap : first) : kMin;
                cin    = (pc == 0) ? 0 : first;
                int32_t cleft = local0;
                // Chunks are taken in groups of up to GC (row_group_v3): the candidate values and the lane-local closure of the
                // chunks of a group are independent of each other, so their loads and arithmetic overlap; only the carry between
                // chunks is sequential, and in the common case it is a scalar chain over values that are already known.
                RowCtx<ScoreT> cx;
                cx.prow0 = prow0;
                cx.prow1 = prow1;
                cx.rd    = read + bs;
                cx.drow  = drow;
                cx.grow  = (!bulk && use_ring) ? rowp : nullptr;
                cx.sh0   = sh0;
                cx.lim0  = lim0;
                cx.sh1   = sh1;
                cx.lim1  = lim1;
                cx.base  = base;
                cx.base4 = static_cast<uint32_t>(base) * 0x01010101u;
                cx.bw    = band_width;
                cx.gap   = gap;
                cx.match = match;
                cx.mismatch = mismatch;
                // Few, compact instantiations run as loops: what the row loop executes has to stay resident in the instruction
                // caches with 10+ windows per SM at different places of it. Measured on C3 (1480 windows, profiles/r02_summary.md):
                // groups of 2 chunks 1121 windows/s, of 1 chunk 1012, of 4 chunks 962 (the fully unrolled 8-variant version 828)
                int32_t c0 = 0;
                if (pc == 2)
                {
                    if (max_group >= 2)
                        for (; c0 + 2 <= nchunks; c0 += 2)
                            row_group_v3<ScoreT, 2, true>(cx, c0, lane, cin, cleft);
                    for (; c0 < nchunks; c0++)
                        row_group_v3<ScoreT, 1, true>(cx, c0, lane, cin, cleft);
                }
                else
                {
                    if (GC >= 4 && max_group >= 4)
                        for (; c0 + 4 <= nchunks; c0 += 4)
                            row_group_v3<ScoreT, (GC >= 4 ? 4 : 1), false>(cx, c0, lane, cin, cleft);
                    if (max_group >= 2)
                        for (; c0 + 2 <= nchunks; c0 += 2)
                            row_group_v3<ScoreT, 2, false>(cx, c0, lane, cin, cleft);
                    for (; c0 < nchunks; c0++)
                        row_group_v3<ScoreT, 1, false>(cx, c0, lane, cin, cleft);
                }
            }
            else
            {
                // ---- general row: any number of predecessors, rows that left the ring come from global memory
                const int4 rm         = row_meta[row];
                const int32_t node_id = rm.x;
                const int32_t pcf     = (rm.w >> 8) & 0xff;
                auto pred_index = [&](int32_t p) -> int32_t {
                    return (p == 0) ? rm.y : (p == 1 ? rm.z : static_cast<int32_t>(g.pos[g.in_edge(node_id, p)]) + 1);
                };
                if (bulk)
                {
                    // Rows reach global memory by the bulk copies. A row with three or more predecessors that are all still in
                    // the ring needs none of that; otherwise everything older than the ring is complete once at most
                    // min(R, 4) - 1 groups are pending (the wait makes the writes visible to the waiting thread -- and
                    // invalidates the L1 --, the warp barrier orders the other lanes behind it).
                    int32_t far = 0;
                    for (int32_t p = 0; p < pcf; p++)
                        far |= (row - pred_index(p) >= R) ? 1 : 0;
                    if (far != 0)
                    {
                        if (lane == 0)
                        {
                            if (R >= 5)
                                bulk_wait<3>();
                            else if (R >= 3)
                                bulk_wait<2>();
                            else
                                bulk_wait<1>();
                            if (row_fence != 0)
                                fence_async_all();
                        }
                        __syncwarp();
                    }
                }
                auto pred_row_ptr = [&](int32_t pi) -> const ScoreT* {
                    const int32_t d = row - pi;
                    if (use_ring && d < R)
                    {
                        int32_t sl = ring_slot - d;
                        if (sl < 0)
                            sl += R;
                        return ring + sl * stride;
                    }
                    return scores + static_cast<int64_t>(pi) * stride;
                };
                int32_t first = 0;
                if (pcf != 0)
                {
                    if (bs > kCPT && pcf == 1)
                    {
                        first = kMin + gap;
                    }
                    else
                    {
                        int32_t penalty = kMin;
                        for (int32_t p = 0; p < pcf; p++)
                            penalty = max(penalty, static_cast<int32_t>(pred_row_ptr(pred_index(p))[0]));
                        first = penalty + gap;
                    }
                }
                local0        = (bs == 0) ? (pcf == 0 ? gap : first) : kMin;
                cin           = (pcf == 0) ? 0 : first;
                int32_t cleft = local0;
                const int32_t np = max(pcf, 1);
                for (int32_t c = 0; c < nchunks; c++)
                {
                    const int32_t off = c * CH + off0;
                    const bool active = off < band_width;
                    const int32_t rp  = bs + off;
                    int32_t q[CPL];
#pragma unroll
                    for (int32_t h = 0; h < CPL / 4; h++)
                    {
                        const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(read + rp + 4 * h));
                        q[4 * h + 0]      = (base == static_cast<int32_t>(w0 & 0xff)) ? match : mismatch;
                        q[4 * h + 1]      = (base == static_cast<int32_t>((w0 >> 8) & 0xff)) ? match : mismatch;
                        q[4 * h + 2]      = (base == static_cast<int32_t>((w0 >> 16) & 0xff)) ? match : mismatch;
                        q[4 * h + 3]      = (base == static_cast<int32_t>(w0 >> 24)) ? match : mismatch;
                    }
                    int32_t s[CPL];
#pragma unroll
                    for (int32_t i = 0; i < CPL; i++)
                        s[i] = kMin;
                    for (int32_t p = 0; p < np; p++)
                    {
                        const int32_t pi   = pred_index(p);
                        const int32_t bsp  = B.start(pi);
                        const int32_t lim  = min(band_width - kCPT, max_column - bsp);
                        const int32_t o    = (bs - bsp) + off;
                        const ScoreT* prow = pred_row_ptr(pi);
#pragma unroll
                        for (int32_t h = 0; h < CPL / 4; h++)
                        {
                            const int32_t oh = o + 4 * h;
                            if (active && oh <= lim)
                            {
                                int32_t b0, b1, b2, b3, b4;
                                load5<ScoreT>(prow + oh, b0, b1, b2, b3, b4);
                                const int32_t t0 = static_cast<ScoreT>(max(b0 + q[4 * h + 0], b1 + gap));
                                const int32_t t1 = static_cast<ScoreT>(max(b1 + q[4 * h + 1], b2 + gap));
                                const int32_t t2 = static_cast<ScoreT>(max(b2 + q[4 * h + 2], b3 + gap));
                                const int32_t t3 = static_cast<ScoreT>(max(b3 + q[4 * h + 3], b4 + gap));
                                s[4 * h + 0]     = (p == 0) ? t0 : max(s[4 * h + 0], t0);
                                s[4 * h + 1]     = (p == 0) ? t1 : max(s[4 * h + 1], t1);
                                s[4 * h + 2]     = (p == 0) ? t2 : max(s[4 * h + 2], t2);
                                s[4 * h + 3]     = (p == 0) ? t3 : max(s[4 * h + 3], t3);
                            }
                        }
                    }
                    int32_t left;
                    closure_seq<CPL>(s, cin, gap, lane, left);
#pragma unroll
                    for (int32_t i = 0; i < CPL; i++)
                        s[i] = static_cast<ScoreT>(s[i]);
                    if (lane == 0)
                        left = cleft;
                    if (active)
                    {
                        unit_store<ScoreT, CPL>(drow + off, left, s);
                        if (!bulk && use_ring)
                            unit_store<ScoreT, CPL>(rowp + off, left, s);
                    }
                    const int32_t last_lane = min(31, (band_width - c * CH) / CPL - 1);
                    cin                     = __shfl_sync(kFull, s[CPL - 1], last_lane);
                    cleft                   = cin;
                }
            }
            // last real cell (local band_width) + right padding
            if (lane < 2)
            {
                Vec4<ScoreT> tl;
                tl.x = static_cast<ScoreT>(lane == 0 ? cin : kMin);
                tl.y = static_cast<ScoreT>(kMin);
                tl.z = static_cast<ScoreT>(kMin);
                tl.w = static_cast<ScoreT>(kMin);
                *reinterpret_cast<Vec4<ScoreT>*>(drow + band_width + 4 * lane) = tl;
                if (!bulk && use_ring)
                    *reinterpret_cast<Vec4<ScoreT>*>(rowp + band_width + 4 * lane) = tl;
            }
            if (bulk)
            {
                // generic-proxy writes of the row -> visible to the async proxy; the slot of the NEXT row (last used R rows
                // before it) must have been read by its bulk copy before any lane writes it
                fence_async_smem();
                if (lane == 0)
                {
                    if (R >= 5)
                        bulk_wait_read<3>();
                    else if (R >= 3)
                        bulk_wait_read<1>();
                    else
                        bulk_wait_read<0>();
                }
                __syncwarp();
                if (lane == 0)
                {
                    bulk_store_s2g(rowp, srow, static_cast<uint32_t>(rowbytes));
                    bulk_commit();
                }
            }
            else
            {
                __syncwarp();
            }
        }
    }
    if (bulk)
    {
        if (lane == 0)
        {
            bulk_wait<0>();
            fence_async_all();
        }
    }
    __syncwarp();
}

// End-cell search + traceback of needlemanWunschBanded (cudapoa_nw_banded.cuh:407-549) for one warp, tile refilled by the lanes
// with plain loads (the v2 routine; kept as the A/B partner of traceback_tma below). Returns the alignment length or a code.
template <typename ScoreT, typename SizeT>
__device__ int32_t traceback_plain(const Win<SizeT>& g, const int32_t graph_count, const uint8_t* read, const int32_t read_length,
                                   const Band<ScoreT>& B, SizeT* aln_graph, SizeT* aln_read, const int32_t band_width, const int32_t gap,
                                   const int32_t mismatch, const int32_t match, const int32_t rerun, const bool Adaptive, const int4* row_meta,
                                   uint8_t* pool, unsigned long long* timers, unsigned long long& t_ph__)
{
    constexpr int32_t kMin   = min_score_of<ScoreT>();
    const int32_t lane       = threadIdx.x & 31;
    const int32_t max_column = read_length + 1;
    // ---- end cell: first strict maximum over sink rows at column read_length (cudapoa_nw_banded.cuh:407-426)
    int32_t i = 0;
    {
        int32_t best_s = kMin, best_i = 0;
        for (int32_t idx = 1 + lane; idx <= graph_count; idx += 32)
        {
            const int32_t misc = row_meta[idx].w;
            if ((misc >> 16) & 1)
            {
                const int32_t s = B.get(idx, read_length);
                if (best_s < s)
                {
                    best_s = s;
                    best_i = idx;
                }
            }
        }
#pragma unroll
        for (int32_t d = 16; d >= 1; d >>= 1)
        {
            const int32_t os = __shfl_xor_sync(kFull, best_s, d);
            const int32_t oi = __shfl_xor_sync(kFull, best_i, d);
            if (os > best_s || (os == best_s && oi < best_i))
            {
                best_s = os;
                best_i = oi;
            }
        }
        i = best_i;
    }
    GWB200_TIMER_LAP(1);

    // ---- traceback (cudapoa_nw_banded.cuh:428-549): warp-uniform walk over a shared-memory tile
    ScoreT* tile = reinterpret_cast<ScoreT*>(pool);                                         // [kTileRows][kTileCols]
    int4* tmeta  = reinterpret_cast<int4*>(pool + kTileRows * kTileCols * sizeof(ScoreT)); // [kTileRows]
    uint8_t* tread = reinterpret_cast<uint8_t*>(tmeta + kTileRows);                          // read[J0 - 1 + k], k in [0, 66)
    int8_t* tjump  = reinterpret_cast<int8_t*>(tread + 72);                                  // [5][kTileRows]: 2^m-th pred0 ancestor
    int32_t t_lo = 1, t_hi = 0, J0 = 0; // tile rows [t_lo, t_hi], columns [J0, J0 + kTileCols)

    auto refill = [&](int32_t ri, int32_t rj) {
        __syncwarp();
        t_hi = ri;
        t_lo = max(0, ri - (kTileRows - 1));
        J0   = max(0, rj - (kTileCols - 2)) & ~1;
        {
            const int32_t row = t_lo + lane; // metadata: lane k <-> row t_lo + k
            int32_t up        = -1;          // tile-relative row of the first predecessor, -1 = none inside the tile
            if (row <= t_hi)
            {
                const int4 mm = row >= 1 ? row_meta[row] : make_int4(0, 0, 0, 0);
                tmeta[lane]   = mm;
                if (row >= 1 && mm.y >= t_lo)
                    up = mm.y - t_lo;
            }
            // pointer doubling over the first-predecessor links: tjump[m][r] = 2^m-th ancestor of tile row r (or -1)
            tjump[lane] = static_cast<int8_t>(up);
#pragma unroll
            for (int32_t m = 1; m < 5; m++)
            {
                __syncwarp();
                if (up >= 0)
                    up = tjump[(m - 1) * kTileRows + up];
                tjump[m * kTileRows + lane] = static_cast<int8_t>(up);
            }
        }
        for (int32_t k = lane; k < kTileCols + 2; k += 32)
        {
            const int32_t rp = J0 - 1 + k;
            tread[k]         = (rp >= 0 && rp < read_length) ? __ldg(read + rp) : 0;
        }
        // scores: one row per iteration, 2 columns per lane, band test as get_score() does (:80-102)
        const int32_t nr = t_hi - t_lo;
#pragma unroll 8
        for (int32_t r = 0; r <= nr; r++)
        {
            const int32_t row = t_lo + r;
            const int32_t bs  = B.start(row);
            const int32_t be  = min(bs + band_width, max_column);
            const int32_t J   = J0 + 2 * lane;
            int32_t v0 = kMin, v1 = kMin;
            const ScoreT* rp = B.row_ptr(row);
            if (J >= bs && J <= be)
                v0 = rp[J - bs];
            if (J + 1 >= bs && J + 1 <= be)
                v1 = rp[J + 1 - bs];
            tile[r * kTileCols + 2 * lane]     = static_cast<ScoreT>(v0);
            tile[r * kTileCols + 2 * lane + 1] = static_cast<ScoreT>(v1);
        }
        __syncwarp();
    };
    // score as the reference's get_score(row, column) sees it: tile hit, else the (rare) global-memory path
    auto T = [&](int32_t row, int32_t column) -> int32_t {
        const uint32_t r = static_cast<uint32_t>(row - t_lo);
        const uint32_t c = static_cast<uint32_t>(column - J0);
        if (r <= static_cast<uint32_t>(t_hi - t_lo) && c < static_cast<uint32_t>(kTileCols))
            return tile[r * kTileCols + c];
        return band_get_slow<ScoreT>(B, row, column);
    };

    int32_t aligned_nodes = 0;
    {
        int32_t j      = read_length;
        int32_t prev_i = 0, prev_j = 0;
        int32_t next_node_id    = i > 0 ? row_meta[i].x : 0;
        int32_t loop_count      = 0;
        const int32_t limit     = read_length + graph_count + 2;
        const int32_t threshold = max(1, max_column / 1024);
        const bool check_band   = Adaptive && rerun == 0 && band_width < kMaxAdaptiveBW;
        while (!(i == 0 && j == 0) && loop_count < limit)
        {
            loop_count++;
            if (i < t_lo || i > t_hi || j >= J0 + kTileCols || (j > 0 && j - 1 < J0))
                refill(i, j);
            // ---- speculative run: lane k assumes the previous k steps were all "diagonal through the first predecessor" (the
            // first test of every step, :467-478), finds the row it would stand on by following k first-predecessor links
            // (binary decomposition of k over the tile's pointer-doubling tables) and verifies its own step on the tile; the
            // leading run of successful lanes is exactly what the serial loop would do for those steps, taken at once.
            {
                int32_t pos = i - t_lo; // tile-relative row after `lane` steps, -1 = outside the tile
#pragma unroll
                for (int32_t m = 0; m < 5; m++)
                {
                    if (((lane >> m) & 1) && pos >= 0)
                        pos = tjump[m * kTileRows + pos];
                }
                const int32_t ik = t_lo + pos;
                const int32_t jk = j - lane;
                bool ok          = pos >= 0 && ik >= 1 && jk >= 1 && (jk - 1) >= J0 && (loop_count - 1 + lane) < limit;
                int32_t knode = 0, kup = 0;
                if (ok)
                {
                    const int4 mk = tmeta[pos];
                    knode         = mk.x;
                    kup           = mk.y; // first predecessor row (row 0 for source nodes)
                    ok            = kup >= t_lo;
                    if (check_band && jk > threshold && jk < max_column - threshold)
                    {
                        const int32_t bsk = ((mk.w >> 17) & 0x3fff) << 2;
                        if (jk <= bsk + threshold || jk >= (bsk + band_width - threshold))
                            ok = false; // the serial step below performs the abort
                    }
                    if (ok)
                    {
                        const int32_t cost = ((mk.w & 0xff) == static_cast<int32_t>(tread[jk - J0])) ? match : mismatch;
                        const int32_t sij  = tile[pos * kTileCols + (jk - J0)];
                        const int32_t sd   = tile[(kup - t_lo) * kTileCols + (jk - 1 - J0)];
                        ok                 = sij == sd + cost;
                    }
                }
                // the first step uses next_node_id, which must be the node of row i for the speculation to be the serial behaviour
                const int32_t node0 = __shfl_sync(kFull, knode, 0);
                uint32_t okmask     = __ballot_sync(kFull, ok);
                if (node0 != next_node_id)
                    okmask = 0u;
                const int32_t run = (okmask == kFull) ? 32 : (__ffs(~okmask) - 1);
                if (run > 0)
                {
                    if (lane < run)
                    {
                        aln_graph[aligned_nodes + lane] = static_cast<SizeT>(knode);
                        aln_read[aligned_nodes + lane]  = static_cast<SizeT>(jk - 1);
                    }
                    aligned_nodes += run;
                    loop_count += run - 1; // the loop header already counted one step
                    i = __shfl_sync(kFull, kup, run - 1);
                    j -= run;
                    prev_i = i;
                    prev_j = j;
                    if (i > 0)
                    {
                        const uint32_t r = static_cast<uint32_t>(i - t_lo);
                        next_node_id     = (r <= static_cast<uint32_t>(t_hi - t_lo)) ? tmeta[r].x : row_meta[i].x;
                    }
                    else
                    {
                        next_node_id = 0;
                    }
                    continue;
                }
            }
            const int32_t ti        = i - t_lo;
            const int32_t tj        = j - J0;
            const int32_t scores_ij = tile[ti * kTileCols + tj];
            const int4 m            = tmeta[ti];
            const int32_t row_node  = m.x; // graph[i - 1]
            const int32_t pc_i      = (m.w >> 8) & 0xff;
            bool pred_found         = false;
            if (i != 0 && j != 0)
            {
                if (check_band && j > threshold && j < max_column - threshold)
                {
                    const int32_t bs = ((m.w >> 17) & 0x3fff) << 2; // band start of row i, packed by the DP phase
                    if (j <= bs + threshold)
                    {
                        aligned_nodes = kShiftLeft;
                        break;
                    }
                    if (j >= (bs + band_width - threshold))
                    {
                        aligned_nodes = kShiftRight;
                        break;
                    }
                }
                // the reference uses next_node_id (= graph[prev_i - 1] of the previous step) here; it equals graph[i - 1]
                // whenever the previous step found a predecessor and is stale otherwise -- that behaviour is kept
                const int32_t node_id = next_node_id;
                int32_t nbase = m.w & 0xff, pc = pc_i, pred_i = m.y;
                if (node_id != row_node)
                {
                    nbase  = g.nodes[node_id];
                    pc     = g.in_cnt[node_id];
                    pred_i = (pc == 0) ? 0 : (static_cast<int32_t>(g.pos[g.in_edge(node_id, 0)]) + 1);
                }
                const int32_t match_cost = (nbase == static_cast<int32_t>(tread[tj])) ? match : mismatch; // read[j - 1]
                if (scores_ij == (T(pred_i, j - 1) + match_cost))
                {
                    prev_i     = pred_i;
                    prev_j     = j - 1;
                    pred_found = true;
                }
                else
                {
                    for (int32_t p = 1; p < pc; p++)
                    {
                        pred_i = (p == 1 && node_id == row_node) ? m.z : (static_cast<int32_t>(g.pos[g.in_edge(node_id, p)]) + 1);
                        if (scores_ij == (T(pred_i, j - 1) + match_cost))
                        {
                            prev_i     = pred_i;
                            prev_j     = j - 1;
                            pred_found = true;
                            break;
                        }
                    }
                }
            }
            if (!pred_found && i != 0)
            {
                int32_t pred_i = (pc_i == 0) ? 0 : m.y;
                if (scores_ij == T(pred_i, j) + gap)
                {
                    prev_i     = pred_i;
                    prev_j     = j;
                    pred_found = true;
                }
                else
                {
                    for (int32_t p = 1; p < pc_i; p++)
                    {
                        pred_i = (p == 1) ? m.z : (static_cast<int32_t>(g.pos[g.in_edge(row_node, p)]) + 1);
                        if (scores_ij == T(pred_i, j) + gap)
                        {
                            prev_i     = pred_i;
                            prev_j     = j;
                            pred_found = true;
                            break;
                        }
                    }
                }
            }
            if (!pred_found && scores_ij == T(i, j - 1) + gap)
            {
                prev_i     = i;
                prev_j     = j - 1;
                pred_found = true;
            }
            // next_node_id = graph[prev_i - 1]
            if (prev_i > 0)
            {
                const uint32_t r = static_cast<uint32_t>(prev_i - t_lo);
                next_node_id     = (r <= static_cast<uint32_t>(t_hi - t_lo)) ? tmeta[r].x : row_meta[prev_i].x;
            }
            else
            {
                next_node_id = 0;
            }
            if (lane == 0)
            {
                aln_graph[aligned_nodes] = static_cast<SizeT>((i == prev_i) ? -1 : row_node);
                aln_read[aligned_nodes]  = static_cast<SizeT>((j == prev_j) ? -1 : j - 1);
            }
            aligned_nodes++;
            i = prev_i;
            j = prev_j;
        }
        if (loop_count >= limit)
            aligned_nodes = kNWBacktrackFail;
    }
    __syncwarp();
    return aligned_nodes;
}

// ---- traceback with bulk-asynchronous tiles ---------------------------------------------------------------------------------
// A tile buffer holds 32 rows x 64 columns of scores, the rows' metadata, the pointer-doubling tables of the first-predecessor
// links and the read characters of its columns. Two buffers: the walk runs on one while the next tile along the predicted path
// (rows continue upwards, columns follow the band gradient) is in flight. Every lane copies the column range of one row with
// one cp.async.bulk (16-byte granularity: an int16 row can land up to 4 columns to the right, rowoff[] records where), lane 0
// also copies the 32 metadata records; all complete on the buffer's mbarrier.
template <typename ScoreT>
struct TileBuf
{
    static constexpr int32_t kRows  = 32;
    static constexpr int32_t kCols  = 64;
    static constexpr int32_t kA     = 16 / static_cast<int32_t>(sizeof(ScoreT));      // elements per 16 bytes
    static constexpr int32_t kW     = kCols + (sizeof(ScoreT) == 2 ? 8 : 0);          // row pitch in elements
    static constexpr int32_t kScoresBytes = kRows * kW * static_cast<int32_t>(sizeof(ScoreT));
    static constexpr int32_t kMetaOff     = kScoresBytes;                             // int4 [32]
    static constexpr int32_t kRowOff      = kMetaOff + kRows * 16;                    // int32 [32]
    static constexpr int32_t kJumpOff     = kRowOff + kRows * 4;                      // int8 [5][32]
    static constexpr int32_t kReadOff     = kJumpOff + 5 * kRows;                     // uint8 [80]: read[J0 - 4 + k]
    static constexpr int32_t kBytes       = (kReadOff + 80 + 15) & ~15;
};

template <typename ScoreT, typename SizeT>
__device__ int32_t traceback_tma(const Win<SizeT>& g, const int32_t graph_count, const uint8_t* read, const int32_t read_length,
                                 const Band<ScoreT>& B, SizeT* aln_graph, SizeT* aln_read, const int32_t band_width, const int32_t gap,
                                 const int32_t mismatch, const int32_t match, const int32_t rerun, const bool Adaptive, const int4* row_meta,
                                 uint8_t* pool, unsigned long long* bars, unsigned long long* timers, unsigned long long& t_ph__)
{
    using TB                 = TileBuf<ScoreT>;
    constexpr int32_t kMin   = min_score_of<ScoreT>();
    constexpr int32_t kTR    = TB::kRows;
    constexpr int32_t kTC    = TB::kCols;
    const int32_t lane       = threadIdx.x & 31;
    const int32_t max_column = read_length + 1;
    const int32_t stride     = B.stride;

    // ---- end cell: first strict maximum over sink rows at column read_length (cudapoa_nw_banded.cuh:407-426)
    int32_t i = 0;
    {
        int32_t best_s = kMin, best_i = 0;
        for (int32_t idx = 1 + lane; idx <= graph_count; idx += 32)
        {
            const int32_t misc = row_meta[idx].w;
            if ((misc >> 16) & 1)
            {
                const int32_t s = B.get(idx, read_length);
                if (best_s < s)
                {
                    best_s = s;
                    best_i = idx;
                }
            }
        }
#pragma unroll
        for (int32_t d = 16; d >= 1; d >>= 1)
        {
            const int32_t os = __shfl_xor_sync(kFull, best_s, d);
            const int32_t oi = __shfl_xor_sync(kFull, best_i, d);
            if (os > best_s || (os == best_s && oi < best_i))
            {
                best_s = os;
                best_i = oi;
            }
        }
        i = best_i;
    }
    GWB200_TIMER_LAP(1);

    // ---- tile machinery
    int32_t lo[2] = {1, 1}, hi[2] = {0, 0}, j0[2] = {0, 0}; // rows [lo, hi], columns [j0, j0 + 64) of the two buffers
    uint32_t phase[2];                                       // mbarrier parity to wait for next
    bool inflight[2] = {false, false};
    uint32_t rd_pending = 0;                                 // read characters of the tile in flight (lane l: 4 bytes)
    // the barriers keep their phase across alignments: recover the parity from the barrier word is not possible, so the kernel
    // stores it next to the barriers
    uint32_t* phase_store = reinterpret_cast<uint32_t*>(bars + 2);
    phase[0]              = phase_store[0];
    phase[1]              = phase_store[1];

    auto buf_ptr = [&](int32_t b) -> uint8_t* { return pool + b * TB::kBytes; };

    // start the copies of the tile whose top-right corner is near (ri, rj): rows [max(0, ri - 31), ri], columns from J0
    auto issue = [&](int32_t b, int32_t ri, int32_t J0) {
        uint8_t* base = buf_ptr(b);
        ScoreT* sc    = reinterpret_cast<ScoreT*>(base);
        hi[b]         = ri;
        lo[b]         = max(0, ri - (kTR - 1));
        j0[b]         = J0;
        const int32_t row = lo[b] + lane;
        uint32_t bytes    = 0;
        const ScoreT* src = nullptr;
        ScoreT* dst       = nullptr;
        if (row <= ri)
        {
            const int32_t bs = B.start(row);
            const int32_t x  = J0 - bs;
            int32_t l_src, t_dst;
            if (x >= 0)
            {
                l_src = x & ~(TB::kA - 1);
                t_dst = 0;
            }
            else
            {
                l_src = 0;
                t_dst = (-x + TB::kA - 1) & ~(TB::kA - 1);
            }
            const int32_t nelem = min(stride - l_src, TB::kW - t_dst);
            if (nelem > 0)
            {
                bytes = static_cast<uint32_t>(nelem) * static_cast<uint32_t>(sizeof(ScoreT));
                src   = B.row_ptr(row) + l_src;
                dst   = sc + lane * TB::kW + t_dst;
            }
            // index of column c of this row inside the buffer: rowoff + c
            reinterpret_cast<int32_t*>(base + TB::kRowOff)[lane] = lane * TB::kW + t_dst - l_src - bs;
        }
        // generic-proxy accesses of the buffer's previous contents are ordered before the asynchronous writes
        fence_async_smem();
        __syncwarp();
        const uint32_t meta_bytes = static_cast<uint32_t>(ri - lo[b] + 1) * 16u;
        mbar_arrive_expect_tx(&bars[b], bytes + (lane == 0 ? meta_bytes : 0u));
        if (bytes)
            bulk_load_g2s(dst, src, bytes, &bars[b]);
        if (lane == 0)
            bulk_load_g2s(base + TB::kMetaOff, row_meta + lo[b], meta_bytes, &bars[b]);
        // read characters: lane l < 18 holds read[J0 - 4 + 4l .. +3] until the tile is completed
        rd_pending = 0;
        {
            const int32_t ri4 = J0 - 4 + 4 * lane;
            if (lane < 18 && ri4 >= 0 && ri4 < read_length)
                rd_pending = __ldg(reinterpret_cast<const uint32_t*>(read + ri4));
        }
        inflight[b] = true;
    };

    // wait for the copies of buffer b and finish it: metadata of row 0, first-predecessor jump tables, cells outside the rows'
    // bands (get_score() returns min_score there, :80-102), read characters
    auto complete = [&](int32_t b) {
        mbar_wait(&bars[b], phase[b]);
        phase[b] ^= 1u;
        inflight[b]   = false;
        uint8_t* base = buf_ptr(b);
        ScoreT* sc    = reinterpret_cast<ScoreT*>(base);
        int4* tm      = reinterpret_cast<int4*>(base + TB::kMetaOff);
        int8_t* tj    = reinterpret_cast<int8_t*>(base + TB::kJumpOff);
        uint8_t* trd  = base + TB::kReadOff;
        const int32_t row = lo[b] + lane;
        int32_t up        = -1;
        if (row <= hi[b])
        {
            if (row == 0)
                tm[lane] = make_int4(0, 0, 0, 0);
            const int4 mm = tm[lane];
            if (row >= 1 && mm.y >= lo[b])
                up = mm.y - lo[b];
            // cells of the tile's columns outside [band start, band end] of the row
            const int32_t bs  = (row >= 1) ? (((mm.w >> 17) & 0x3fff) << 2) : 0;
            const int32_t be  = min(bs + band_width, max_column);
            const int32_t ro  = reinterpret_cast<const int32_t*>(base + TB::kRowOff)[lane];
            const int32_t c0  = j0[b];
            const int32_t c1  = j0[b] + kTC; // exclusive
            for (int32_t c = c0; c < min(bs, c1); c++)
                sc[ro + c] = static_cast<ScoreT>(kMin);
            for (int32_t c = max(be + 1, c0); c < c1; c++)
                sc[ro + c] = static_cast<ScoreT>(kMin);
        }
        if (lane < 18)
            reinterpret_cast<uint32_t*>(trd)[lane] = rd_pending;
        tj[lane] = static_cast<int8_t>(up);
#pragma unroll
        for (int32_t m = 1; m < 5; m++)
        {
            __syncwarp();
            if (up >= 0)
                up = tj[(m - 1) * kTR + up];
            tj[m * kTR + lane] = static_cast<int8_t>(up);
        }
        __syncwarp();
    };

    int32_t cur = 0; // buffer the walk runs on
    ScoreT* tile     = nullptr;
    const int4* tmeta = nullptr;
    const int32_t* trow = nullptr;
    const int8_t* tjump = nullptr;
    const uint8_t* tread = nullptr;
    int32_t t_lo = 1, t_hi = 0, J0 = 0;
    auto select = [&](int32_t b) {
        cur           = b;
        uint8_t* base = buf_ptr(b);
        tile          = reinterpret_cast<ScoreT*>(base);
        tmeta         = reinterpret_cast<const int4*>(base + TB::kMetaOff);
        trow          = reinterpret_cast<const int32_t*>(base + TB::kRowOff);
        tjump         = reinterpret_cast<const int8_t*>(base + TB::kJumpOff);
        tread         = base + TB::kReadOff;
        t_lo          = lo[b];
        t_hi          = hi[b];
        J0            = j0[b];
    };
    auto inside = [&](int32_t b, int32_t ri, int32_t rj) -> bool {
        return !(ri < lo[b] || ri > hi[b] || rj >= j0[b] + kTC || (rj > 0 && rj - 1 < j0[b]));
    };
    // leftmost column of a tile that has column rj in its rightmost kA columns (j - 1 >= J0 and j < J0 + 64 hold for j = rj)
    auto tile_j0 = [&](int32_t rj) -> int32_t { return max(0, rj - (kTC - TB::kA)) & ~(TB::kA - 1); };
    // prefetch the tile above the current one along the predicted path
    auto prefetch = [&](int32_t ri, int32_t rj) {
        const int32_t nb = cur ^ 1;
        if (t_lo >= 1)
        {
            const int32_t rows_left = ri - t_lo + 1;
            const int32_t pj        = rj - static_cast<int32_t>(static_cast<float>(rows_left) * B.gradient);
            issue(nb, t_lo - 1, tile_j0(max(pj + 8, 1)));
        }
    };
    auto ensure = [&](int32_t ri, int32_t rj) {
        __syncwarp();
        const int32_t nb = cur ^ 1;
        if (inflight[nb] && inside(nb, ri, rj))
        {
            complete(nb);
            select(nb);
        }
        else
        {
            if (inflight[nb])
            {
                // the prediction missed: drain the copies before the buffer can be used again
                mbar_wait(&bars[nb], phase[nb]);
                phase[nb] ^= 1u;
                inflight[nb] = false;
            }
            issue(cur, ri, tile_j0(rj));
            complete(cur);
            select(cur);
        }
        prefetch(ri, rj);
    };
    // score as the reference's get_score(row, column) sees it: tile hit, else the (rare) global-memory path
    auto T = [&](int32_t row, int32_t column) -> int32_t {
        const uint32_t r = static_cast<uint32_t>(row - t_lo);
        const uint32_t c = static_cast<uint32_t>(column - J0);
        if (r <= static_cast<uint32_t>(t_hi - t_lo) && c < static_cast<uint32_t>(kTC))
            return tile[trow[r] + column];
        return band_get_slow<ScoreT>(B, row, column);
    };

    int32_t aligned_nodes = 0;
    {
        int32_t j      = read_length;
        int32_t prev_i = 0, prev_j = 0;
        int32_t next_node_id    = i > 0 ? row_meta[i].x : 0;
        int32_t loop_count      = 0;
        const int32_t limit     = read_length + graph_count + 2;
        const int32_t threshold = max(1, max_column / 1024);
        const bool check_band   = Adaptive && rerun == 0 && band_width < kMaxAdaptiveBW;
        while (!(i == 0 && j == 0) && loop_count < limit)
        {
            loop_count++;
            if (i < t_lo || i > t_hi || j >= J0 + kTC || (j > 0 && j - 1 < J0))
                ensure(i, j);
            // ---- speculative run (see nw_banded_v2): lane k verifies step k of a run of "diagonal through the first predecessor"
            {
                int32_t pos = i - t_lo;
#pragma unroll
                for (int32_t m = 0; m < 5; m++)
                {
                    if (((lane >> m) & 1) && pos >= 0)
                        pos = tjump[m * kTR + pos];
                }
                const int32_t ik = t_lo + pos;
                const int32_t jk = j - lane;
                bool ok          = pos >= 0 && ik >= 1 && jk >= 1 && (jk - 1) >= J0 && (loop_count - 1 + lane) < limit;
                int32_t knode = 0, kup = 0;
                if (ok)
                {
                    const int4 mk = tmeta[pos];
                    knode         = mk.x;
                    kup           = mk.y;
                    ok            = kup >= t_lo;
                    if (check_band && jk > threshold && jk < max_column - threshold)
                    {
                        const int32_t bsk = ((mk.w >> 17) & 0x3fff) << 2;
                        if (jk <= bsk + threshold || jk >= (bsk + band_width - threshold))
                            ok = false; // the serial step below performs the abort
                    }
                    if (ok)
                    {
                        const int32_t cost = ((mk.w & 0xff) == static_cast<int32_t>(tread[jk - J0 + 3])) ? match : mismatch;
                        const int32_t sij  = tile[trow[pos] + jk];
                        const int32_t sd   = tile[trow[kup - t_lo] + jk - 1];
                        ok                 = sij == sd + cost;
                    }
                }
                const int32_t node0 = __shfl_sync(kFull, knode, 0);
                uint32_t okmask     = __ballot_sync(kFull, ok);
                if (node0 != next_node_id)
                    okmask = 0u;
                const int32_t run = (okmask == kFull) ? 32 : (__ffs(~okmask) - 1);
                if (run > 0)
                {
                    if (lane < run)
                    {
                        aln_graph[aligned_nodes + lane] = static_cast<SizeT>(knode);
                        aln_read[aligned_nodes + lane]  = static_cast<SizeT>(jk - 1);
                    }
                    aligned_nodes += run;
                    loop_count += run - 1;
                    i = __shfl_sync(kFull, kup, run - 1);
                    j -= run;
                    prev_i = i;
                    prev_j = j;
                    if (i > 0)
                    {
                        const uint32_t r = static_cast<uint32_t>(i - t_lo);
                        next_node_id     = (r <= static_cast<uint32_t>(t_hi - t_lo)) ? tmeta[r].x : row_meta[i].x;
                    }
                    else
                    {
                        next_node_id = 0;
                    }
                    continue;
                }
            }
            const int32_t ti        = i - t_lo;
            const int32_t tj        = j - J0;
            const int32_t scores_ij = tile[trow[ti] + j];
            const int4 m            = tmeta[ti];
            const int32_t row_node  = m.x;
            const int32_t pc_i      = (m.w >> 8) & 0xff;
            bool pred_found         = false;
            if (i != 0 && j != 0)
            {
                if (check_band && j > threshold && j < max_column - threshold)
                {
                    const int32_t bs = ((m.w >> 17) & 0x3fff) << 2;
                    if (j <= bs + threshold)
                    {
                        aligned_nodes = kShiftLeft;
                        break;
                    }
                    if (j >= (bs + band_width - threshold))
                    {
                        aligned_nodes = kShiftRight;
                        break;
                    }
                }
                // the reference uses next_node_id (= graph[prev_i - 1] of the previous step) here; kept, see nw_banded_v2
                const int32_t node_id = next_node_id;
                int32_t nbase = m.w & 0xff, pc = pc_i, pred_i = m.y;
                if (node_id != row_node)
                {
                    nbase  = g.nodes[node_id];
                    pc     = g.in_cnt[node_id];
                    pred_i = (pc == 0) ? 0 : (static_cast<int32_t>(g.pos[g.in_edge(node_id, 0)]) + 1);
                }
                const int32_t match_cost = (nbase == static_cast<int32_t>(tread[tj + 3])) ? match : mismatch; // read[j - 1]
                if (scores_ij == (T(pred_i, j - 1) + match_cost))
                {
                    prev_i     = pred_i;
                    prev_j     = j - 1;
                    pred_found = true;
                }
                else
                {
                    for (int32_t p = 1; p < pc; p++)
                    {
                        pred_i = (p == 1 && node_id == row_node) ? m.z : (static_cast<int32_t>(g.pos[g.in_edge(node_id, p)]) + 1);
                        if (scores_ij == (T(pred_i, j - 1) + match_cost))
                        {
                            prev_i     = pred_i;
                            prev_j     = j - 1;
                            pred_found = true;
                            break;
                        }
                    }
                }
            }
            if (!pred_found && i != 0)
            {
                int32_t pred_i = (pc_i == 0) ? 0 : m.y;
                if (scores_ij == T(pred_i, j) + gap)
                {
                    prev_i     = pred_i;
                    prev_j     = j;
                    pred_found = true;
                }
                else
                {
                    for (int32_t p = 1; p < pc_i; p++)
                    {
                        pred_i = (p == 1) ? m.z : (static_cast<int32_t>(g.pos[g.in_edge(row_node, p)]) + 1);
                        if (scores_ij == T(pred_i, j) + gap)
                        {
                            prev_i     = pred_i;
                            prev_j     = j;
                            pred_found = true;
                            break;
                        }
                    }
                }
            }
            if (!pred_found && scores_ij == T(i, j - 1) + gap)
            {
                prev_i     = i;
                prev_j     = j - 1;
                pred_found = true;
            }
            if (prev_i > 0)
            {
                const uint32_t r = static_cast<uint32_t>(prev_i - t_lo);
                next_node_id     = (r <= static_cast<uint32_t>(t_hi - t_lo)) ? tmeta[r].x : row_meta[prev_i].x;
            }
            else
            {
                next_node_id = 0;
            }
            if (lane == 0)
            {
                aln_graph[aligned_nodes] = static_cast<SizeT>((i == prev_i) ? -1 : row_node);
                aln_read[aligned_nodes]  = static_cast<SizeT>((j == prev_j) ? -1 : j - 1);
            }
            aligned_nodes++;
            i = prev_i;
            j = prev_j;
        }
        if (loop_count >= limit)
            aligned_nodes = kNWBacktrackFail;
    }
    // no copy may be in flight when the pool is handed to the next phase
    {
        const int32_t nb = cur ^ 1;
        if (inflight[nb])
        {
            mbar_wait(&bars[nb], phase[nb]);
            phase[nb] ^= 1u;
        }
    }
    __syncwarp();
    if (lane == 0)
    {
        phase_store[0] = phase[0];
        phase_store[1] = phase[1];
    }
    __syncwarp();
    return aligned_nodes;
}

// topologicalSortDeviceUtil (cudapoa_topsort.cuh:45-97): same Kahn FIFO order (the order is observable: rank <-> DP row).
// Staging (all lanes, coalesced): in-degree counters (u8) into the shared-memory pool, and one packed record per node
// {min(out-degree, 3), first child, second child} into a per-window scratch array in global memory (32 bits for 16-bit node ids,
// 64 bits otherwise) that the walk reads sequentially-local (chains run along consecutive node ids: one L1 miss per line).
// Walk (lane 0): one record load + one counter update per edge; the two front entries of the FIFO live in registers (a bubble
// keeps two branches interleaved in the queue), later entries in a 64-entry tagged window in shared memory, the rest is read
// back from sorted[]. Nodes with more than two children read their further children from the adjacency arrays.
template <typename SizeT>
__device__ void topsort_v3(const Win<SizeT>& g, const int32_t node_count, uint8_t* pool, const int32_t pool_bytes, void* scratch)
{
    constexpr bool kSmall = sizeof(SizeT) == 2;
    constexpr int32_t kQ  = 64;
    const int32_t lane    = threadIdx.x & 31;
    int32_t* q_tag        = reinterpret_cast<int32_t*>(pool);
    int32_t* q_node       = q_tag + kQ;
    uint8_t* cnt          = pool + kQ * 8;
    // counters of the nodes that do not fit the pool live in local_cnt (global memory)
    const int32_t n_smem  = max(0, min(node_count, pool_bytes - kQ * 8 - 16));
    uint32_t* rec32       = static_cast<uint32_t*>(scratch);
    uint2* rec64          = static_cast<uint2*>(scratch);

    for (int32_t k = lane; k < kQ; k += 32)
        q_tag[k] = -1;
    __syncwarp();
    int32_t p = 0;
    for (int32_t n0 = 0; n0 < node_count; n0 += 32)
    {
        const int32_t n = n0 + lane;
        int32_t c       = 1;
        if (n < node_count)
        {
            c                = g.in_cnt[n];
            const int32_t oc = g.out_cnt[n];
            const int32_t e0 = oc >= 1 ? static_cast<int32_t>(g.out_edge(n, 0)) : 0;
            const int32_t e1 = oc >= 2 ? static_cast<int32_t>(g.out_edge(n, 1)) : 0;
            if (n < n_smem)
                cnt[n] = static_cast<uint8_t>(c);
            else
                g.local_cnt[n] = static_cast<uint16_t>(c);
            if (kSmall)
                rec32[n] = static_cast<uint32_t>(min(oc, 3)) | (static_cast<uint32_t>(e0 & 0x7fff) << 2) | (static_cast<uint32_t>(e1 & 0x7fff) << 17);
            else
                rec64[n] = make_uint2(static_cast<uint32_t>(e0) | (static_cast<uint32_t>(min(oc, 3)) << 30), static_cast<uint32_t>(e1));
        }
        const bool is_src = n < node_count && c == 0;
        const uint32_t m  = __ballot_sync(kFull, is_src);
        if (is_src)
        {
            const int32_t at = p + __popc(m & ((1u << lane) - 1));
            g.pos[n]         = static_cast<SizeT>(at);
            g.sorted[at]     = static_cast<SizeT>(n);
            if (at < kQ)
            {
                q_node[at] = n;
                q_tag[at]  = at;
            }
        }
        p += __popc(m);
    }
    __syncwarp();
    if (lane == 0)
    {
        int32_t n  = 0;          // next position of sorted[] to process
        int32_t r0 = -1, r1 = -1; // nodes at positions n and n + 1 when known
        auto visit = [&](const int32_t child) {
            int32_t c;
            if (child < n_smem)
            {
                c          = static_cast<int32_t>(cnt[child]) - 1;
                cnt[child] = static_cast<uint8_t>(c);
            }
            else
            {
                c                  = static_cast<int32_t>(g.local_cnt[child]) - 1;
                g.local_cnt[child] = static_cast<uint16_t>(c);
            }
            if (c == 0)
            {
                g.pos[child] = static_cast<SizeT>(p);
                g.sorted[p]  = static_cast<SizeT>(child);
                if (p == n)
                    r0 = child;
                else if (p == n + 1)
                    r1 = child;
                else if (p - kQ < n) // slot p % kQ held position p - kQ, already consumed
                {
                    q_node[p & (kQ - 1)] = child;
                    q_tag[p & (kQ - 1)]  = p;
                }
                p++;
            }
        };
        while (n < p)
        {
            int32_t node = r0;
            if (node < 0)
            {
                const int32_t slot = n & (kQ - 1);
                node               = (q_tag[slot] == n) ? q_node[slot] : static_cast<int32_t>(g.sorted[n]);
            }
            r0 = r1;
            r1 = -1;
            n++;
            int32_t oc2, c0, c1;
            if (kSmall)
            {
                const uint32_t r = rec32[node];
                oc2              = static_cast<int32_t>(r & 3u);
                c0               = static_cast<int32_t>((r >> 2) & 0x7fffu);
                c1               = static_cast<int32_t>(r >> 17);
            }
            else
            {
                const uint2 r = rec64[node];
                oc2           = static_cast<int32_t>(r.x >> 30);
                c0            = static_cast<int32_t>(r.x & 0x3fffffffu);
                c1            = static_cast<int32_t>(r.y);
            }
            if (oc2 >= 1)
                visit(c0);
            if (oc2 >= 2)
                visit(c1);
            if (oc2 == 3)
            {
                const int32_t oc = g.out_cnt[node];
                for (int32_t e = 2; e < oc; e++)
                    visit(static_cast<int32_t>(g.out_edge(node, e)));
            }
        }
    }
    __syncwarp();
}

// needlemanWunschBanded (cudapoa_nw_banded.cuh:177-557) for one warp: band geometry as in nw_banded_v2, rows by dp_rows_v3,
// end cell + traceback by traceback_tma (or traceback_plain when the pool is too small for the tile buffers).
} // namespace poa
} // namespace gwb200
#include "poa_kernels_v4.cuh"
namespace gwb200
{
namespace poa
{

template <typename ScoreT, typename SizeT, bool BULK, bool WAVE>
__device__ int32_t nw_banded_v3(const Win<SizeT>& g, int32_t graph_count, const uint8_t* read, int32_t read_length, ScoreT* scores,
                                float max_buffer_size, SizeT* aln_graph, SizeT* aln_read, int32_t band_width, int32_t gap, int32_t mismatch,
                                int32_t match, int32_t rerun, const bool Adaptive, unsigned long long& cells, int4* row_meta, uint8_t* pool,
                                int32_t pool_bytes, unsigned long long* timers, V3Shared* sh, const int32_t tb_mode, const int32_t wavefront, const int32_t max_group,
                                const int32_t row_fence)
{
    GWB200_TIMER_START();
    const float gradient     = __fdividef(static_cast<float>(read_length + 1), static_cast<float>(graph_count + 1));
    const int32_t max_column = read_length + 1;
    if (Adaptive)
    {
        if (static_cast<double>(gradient) > 1.1)
        {
            int32_t v  = static_cast<int32_t>(max_column * 0.08 * static_cast<double>(gradient));
            band_width = max(band_width, (v + kMinBandWidth - 1) & ~(kMinBandWidth - 1));
        }
        if (static_cast<double>(gradient) < 0.8)
        {
            int32_t v  = static_cast<int32_t>(max_column * 0.1 / static_cast<double>(gradient));
            band_width = max(band_width, (v + kMinBandWidth - 1) & ~(kMinBandWidth - 1));
        }
        band_width = min(band_width, kMaxAdaptiveBW);
        if (band_width == kMaxAdaptiveBW && rerun != 0)
            return rerun;
    }
    int32_t band_shift = band_width / 2;
    if (Adaptive)
    {
        if (rerun == kShiftLeft && band_width <= kMaxAdaptiveBW / 2)
        {
            band_width *= 2;
            band_shift = static_cast<int32_t>(band_shift * 2.5);
        }
        if (rerun == kShiftRight && band_width <= kMaxAdaptiveBW / 2)
        {
            band_width *= 2;
            band_shift = static_cast<int32_t>(band_shift * 1.5);
        }
        const float required = static_cast<float>(graph_count) * static_cast<float>(band_width + kRightPad);
        if (required > max_buffer_size)
            return kNWStorageFail;
    }
    if (threadIdx.x == 0)
        cells += static_cast<unsigned long long>(graph_count) * static_cast<unsigned long long>(band_width);

    Band<ScoreT> B{scores, band_width, band_shift, max_column, band_width + kRightPad, gradient};
    bool done = false;
    if constexpr (WAVE && sizeof(ScoreT) == 4)
    {
        if (wavefront != 0 && B.start(graph_count) < 65536)
            done = dp_rows_v4<ScoreT, SizeT>(g, graph_count, read, B, band_width, max_column, gap, mismatch, match, row_meta, pool, pool_bytes,
                                             sh->rec[0]);
    }
    if (!done)
        dp_rows_v3<ScoreT, SizeT, BULK>(g, graph_count, read, B, band_width, max_column, gap, mismatch, match, row_meta, pool, pool_bytes, sh->rec[0],
                                        max_group, row_fence);
    GWB200_TIMER_LAP(0);
    int32_t result;
    if (tb_mode != 0 && pool_bytes >= TileBuf<ScoreT>::kBytes * 2)
        result = traceback_tma<ScoreT, SizeT>(g, graph_count, read, read_length, B, aln_graph, aln_read, band_width, gap, mismatch, match, rerun,
                                              Adaptive, row_meta, pool, sh->tile_bar, timers, t_ph__);
    else
        result = traceback_plain<ScoreT, SizeT>(g, graph_count, read, read_length, B, aln_graph, aln_read, band_width, gap, mismatch, match, rerun,
                                                Adaptive, row_meta, pool, timers, t_ph__);
    GWB200_TIMER_LAP(2);
    return result;
}

// One window from the backbone to the consensus / MSA (cudapoa_kernels.cuh:200-541) on one warp.
template <typename ScoreT, typename SizeT, bool BULK, bool WAVE>
__device__ void process_window_v3(const DeviceParams& P, const V2Extra& X, const V3Extra& Y, const int32_t w, uint8_t* pool, V3Shared* sh)
{
    const bool MSA      = P.msa != 0;
    const int32_t lane  = threadIdx.x & 31;
    const WindowInfo wi = P.windows[w];
    const int64_t mn    = P.max_nodes;
    unsigned long long* timers = X.timers ? X.timers + static_cast<int64_t>(w) * 8 : nullptr;

    Win<SizeT> g;
    g.max_nodes = P.max_nodes;
    g.nodes     = P.nodes + w * mn;
    g.in_cnt    = P.in_cnt + w * mn;
    g.out_cnt   = P.out_cnt + w * mn;
    g.aln_cnt   = P.aln_cnt + w * mn;
    g.cov       = P.node_cov + w * mn;
    g.local_cnt = P.local_cnt + w * mn;
    g.in_edges  = static_cast<SizeT*>(P.in_edges) + w * mn * kMaxEdges;
    g.out_edges = static_cast<SizeT*>(P.out_edges) + w * mn * kMaxEdges;
    g.aligned   = static_cast<SizeT*>(P.aligned) + w * mn * kMaxAligned;
    g.in_w      = P.in_w + w * mn * kMaxEdges;
    g.sorted    = static_cast<SizeT*>(P.sorted) + w * mn;
    g.pos       = static_cast<SizeT*>(P.pos) + w * mn;

    SizeT* aln_graph = static_cast<SizeT*>(P.aln_graph) + static_cast<int64_t>(w) * P.aln_capacity;
    SizeT* aln_read  = static_cast<SizeT*>(P.aln_read) + static_cast<int64_t>(w) * P.aln_capacity;
    int4* row_meta   = X.row_meta + static_cast<int64_t>(w) * (mn + 1);
    SizeT* rd_node   = static_cast<SizeT*>(X.rd_node) + static_cast<int64_t>(w) * X.rd_capacity;

    const int32_t* seq_lengths = P.seq_lengths + wi.seq_len_offset;
    const uint8_t* sequence    = P.sequences + wi.seq_start;
    const int8_t* base_weights = P.weights + wi.seq_start;
    SizeT* path                = MSA ? static_cast<SizeT*>(P.seq_path) + wi.seq_start : nullptr;

    const float banded_buffer_size = static_cast<float>(P.max_nodes) * static_cast<float>(P.matrix_seq_dim);
    ScoreT* scores;
    if (P.band_mode == bm_full_band)
        scores = static_cast<ScoreT*>(P.scores) + wi.scores_offset * mn;
    else
        scores = static_cast<ScoreT*>(P.scores) + static_cast<int64_t>(banded_buffer_size) * static_cast<int64_t>(w);

    uint8_t* consensus = P.consensus + static_cast<int64_t>(w) * P.max_consensus;
    uint16_t* coverage = P.coverage + static_cast<int64_t>(w) * P.max_consensus;

    const int32_t num_seqs = wi.num_seqs;
    int32_t error          = 0;
    int32_t cons_len       = 0;
    unsigned long long cells = 0;
    int32_t node_count     = 0;

    if (num_seqs <= 0)
    {
        // a group whose reads were all rejected stays in the batch as an empty window (cudapoa_batch.cuh:139-148): nothing to align
        error = st_empty_poa_group;
    }
    else
    {
        // backbone from read 0 (cudapoa_kernels.cuh:200-238), lane-parallel
        node_count = seq_lengths[0];
        for (int32_t n = lane; n < node_count; n += 32)
        {
            g.nodes[n]   = sequence[n];
            g.sorted[n]  = static_cast<SizeT>(n);
            g.pos[n]     = static_cast<SizeT>(n);
            g.aln_cnt[n] = 0;
            g.cov[n]     = 1;
            if (n > 0)
            {
                g.in_edge(n, 0)      = static_cast<SizeT>(n - 1);
                g.w(n, 0)            = static_cast<uint16_t>(base_weights[n - 1] + base_weights[n]);
                g.in_cnt[n]          = 1;
                g.out_edge(n - 1, 0) = static_cast<SizeT>(n);
                g.out_cnt[n - 1]     = 1;
            }
            else
            {
                g.in_cnt[0] = 0;
                g.w(0, 0)   = static_cast<uint16_t>(base_weights[0]);
            }
            if (n == node_count - 1)
                g.out_cnt[n] = 0;
            if (MSA)
                path[n] = static_cast<SizeT>(n);
        }
        __syncwarp();

        for (int32_t s = 1; s < num_seqs; s++)
        {
            const int32_t seq_len = seq_lengths[s];
            const int32_t adv     = (seq_lengths[s - 1] + 3) & ~3;
            sequence += adv;
            base_weights += adv;
            if (MSA)
                path += adv;
            if (node_count >= P.max_nodes)
            {
                error = st_node_count_exceeded_maximum_graph_size;
                break;
            }
            int32_t alen;
            if (P.band_mode != bm_full_band)
            {
                const bool adaptive = (P.band_mode == bm_adaptive_band && P.band_width < kMaxAdaptiveBW);
                int32_t rerun       = 0;
                for (int32_t attempt = 0; attempt < 2; attempt++)
                {
                    alen = nw_banded_v3<ScoreT, SizeT, BULK, WAVE>(g, node_count, sequence, seq_len, scores, banded_buffer_size, aln_graph, aln_read,
                                                             P.band_width, P.gap, P.mismatch, P.match, rerun, adaptive, cells, row_meta, pool,
                                                             X.pool_bytes, timers, sh, Y.tb_tma, Y.wavefront, Y.max_group, Y.row_fence);
                    if (!adaptive || attempt == 1 || !(alen == kShiftLeft || alen == kShiftRight))
                        break;
                    rerun = alen; // rerun with extended and shifted band (cudapoa_kernels.cuh:374-396)
                }
            }
            else
            {
                alen = nw_full<ScoreT, SizeT>(g, node_count, sequence, seq_len, scores, wi.scores_width, aln_graph, aln_read, P.gap, P.mismatch,
                                              P.match, cells);
            }
            if (alen == kNWBacktrackFail)
            {
                error = st_loop_count_exceeded_upper_bound;
                break;
            }
            if (alen == kNWStorageFail)
            {
                error = st_exceeded_adaptive_banded_matrix_size;
                break;
            }
            if (alen < 0)
                alen = 0;
            {
                unsigned long long t_ph__ = timers ? clock64() : 0ull;
                int32_t nc = node_count;
                error      = add_alignment_v2<SizeT>(g, nc, alen, aln_graph, sequence, seq_len, aln_read, base_weights, path, rd_node);
                GWB200_TIMER_LAP(3);
                if (!error)
                {
                    if (P.accurate)
                    {
                        if (lane == 0)
                            racon_topsort(g, nc, P.marks + w * mn, P.check + w * mn,
                                          static_cast<SizeT*>(P.stack) + static_cast<int64_t>(w) * P.stack_capacity, P.stack_capacity);
                        __syncwarp();
                    }
                    else
                    {
                        topsort_v3<SizeT>(g, nc, pool, X.pool_bytes, row_meta); // row_meta is free between traceback and the next rows
                    }
                }
                GWB200_TIMER_LAP(4);
                node_count = nc;
                if (error)
                    break;
            }
        }

        if (!error)
        {
            unsigned long long t_ph__ = timers ? clock64() : 0ull;
            if (MSA)
            {
                error = generate_msa<SizeT>(g, node_count, num_seqs, seq_lengths, static_cast<SizeT*>(P.seq_path) + wi.seq_start,
                                            static_cast<SizeT*>(P.msa_col) + w * mn, P.marks + w * mn, P.check + w * mn,
                                            static_cast<SizeT*>(P.stack) + static_cast<int64_t>(w) * P.stack_capacity, P.stack_capacity,
                                            P.msa_out + static_cast<int64_t>(w) * P.max_seqs * P.max_consensus, P.max_consensus);
            }
            else
            {
                error = generate_consensus<SizeT>(g, node_count, P.cons_scores + w * mn, static_cast<SizeT*>(P.cons_preds) + w * mn, consensus,
                                                  coverage, P.max_consensus, &cons_len);
            }
            GWB200_TIMER_LAP(5);
        }
    }
    __syncwarp();
    if (lane == 0)
    {
        if (error)
        {
            consensus[0] = 0;
            cons_len     = 0;
        }
        P.status[w]        = error;
        P.consensus_len[w] = cons_len;
        P.node_count[w]    = node_count;
        P.cells[w]         = cells;
    }
    __syncwarp();
}

// Persistent grid: one warp per CTA, every CTA pulls window indices from the batch's work counter until none is left.
template <typename ScoreT, typename SizeT, bool BULK, bool WAVE>
__global__ void __launch_bounds__(32, (WAVE ? 8 : (sizeof(ScoreT) == 4 ? 12 : 16))) poa_window_kernel_v3(const DeviceParams P, const V2Extra X, const V3Extra Y)
{
    extern __shared__ __align__(16) uint8_t pool[];
    __shared__ __align__(16) V3Shared sh;
    const int32_t lane = threadIdx.x & 31;
    if (lane == 0)
    {
        mbar_init(&sh.tile_bar[0], 32);
        mbar_init(&sh.tile_bar[1], 32);
        sh.tile_phase[0] = 0;
        sh.tile_phase[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    for (;;)
    {
        int32_t w = 0;
        if (lane == 0)
            w = atomicAdd(Y.work_counter, 1);
        w = __shfl_sync(kFull, w, 0);
        if (w >= P.n_windows)
            break;
        process_window_v3<ScoreT, SizeT, BULK, WAVE>(P, X, Y, w, pool, &sh);
    }
}

} // namespace poa
} // namespace gwb200
