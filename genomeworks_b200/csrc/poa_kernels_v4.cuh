// gw-b200 POA device code, fourth generation of the DP rows (sm_100a): a skewed wavefront over the graph rows.
//
// Included by poa_kernels_v3.cuh (the window kernel, traceback, graph update and sort are v3's); this file only replaces the
// row phase of needlemanWunschBanded (cudapoa_nw_banded.cuh:266-405) for 32-bit scores. Result: the same score matrix in HBM,
// cell for cell, as dp_rows_v3 / the reference recurrence.
//
// Why: the ncu capture of v3 (profiles/r02_poa_v3_4k_ncu.md) shows 1.7 warp instructions per DP cell, spread flat over
// max-plus scans, ballots, neighbour shuffles, carry chains between chunks and ~150 instructions of per-row preamble that a
// 256-column band amortises over two chunks. All of that exists because a row is spread ACROSS the lanes. Here a row lives
// IN one lane:
//
//   mapping      lane = row & 31. At iteration `it` the lane of row r computes the 4 cells of absolute column group
//                A = it - r * S4 (S4 = band_width / 128), i.e. consecutive rows are skewed by S4 groups and a lane is done
//                with row r exactly when row r + 32 starts. The horizontal recurrence is a register chain inside the lane
//                (4 x VIADDMNMX), no scan, no shuffle, no ballot; 32 rows are in flight per warp.
//   predecessors every lane writes its group {left, s0, s1, s2} (the image of the HBM row) to slot it % Wg of its own
//                circular window in shared memory and s3 to a second ring; a successor d rows below finds the group it
//                needs in slot (it - d * S4) % Wg of lane (row - d) & 31: one LDS.128 + one LDS.32 per predecessor and
//                iteration, any predecessor distance up to R = (Wg - 1) / S4 rows, any mix of distances across the lanes.
//                Up to three predecessors are folded with max before the cell update (max over predecessors commutes with
//                the +substitution / +gap terms); rows with more, with a predecessor older than R rows, with row 0 as
//                predecessor or with a predecessor group outside that row's band take a per-lane general path.
//   write-back   every 8 iterations the warp empties the windows cooperatively: the <= 8 pending groups of a lane are <= 128
//                contiguous bytes of its HBM row, so one LDS.128 + STG.128 pair moves the segments of four rows as four full
//                128-byte lines (thread t: row 4j + t / 8, 16-byte chunk t % 8). A per-lane cp.async.bulk was measured out:
//                UBLKCP takes uniform registers, 32 different rows make the compiler emit a 32-step waterfall (~350
//                instructions per hand-over). Rows are therefore complete in HBM at most 8 iterations after they are
//                computed, which is what the rare far-predecessor reads rely on.
//   per row      the lane that starts a row unpacks one 16-byte record prepared 32 rows ahead by all lanes in parallel
//                (band start, base, predecessor distances, validity limit): ~40 instructions per row on one lane.
#pragma once

namespace gwb200
{
namespace poa
{

__device__ __forceinline__ int4 lds128(uint32_t a)
{
    int4 v;
    asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ int32_t lds32(uint32_t a)
{
    int32_t v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, int32_t x, int32_t y, int32_t z, int32_t w)
{
    asm volatile("st.shared.v4.s32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, int32_t x) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }

// Geometry of the wavefront for one alignment
struct WfGeom
{
    int32_t S4;     // skew between consecutive rows, in 4-column groups (band_width / 128)
    int32_t Wg;     // slots per lane window (multiple of 8, >= 32)
    int32_t R;      // largest predecessor distance served from the windows
    uint32_t win4;  // shared address of the group windows: [33][Wg] int4, lane pitch Wg * 16 + 16 (entry 32 = dummy, all kWfLow)
    uint32_t s3r;   // shared address of the s3 rings:     [33][Wg] int32, lane pitch Wg * 4 + 4
    uint32_t l0v;   // shared address of the local-0 values of the last 64 rows
    int32_t pitch4, pitch1;
};

constexpr int32_t kWfLow = -1610612736; // below every reachable score (min_score - 2^29): never wins a max, never overflows with + gap

__device__ __forceinline__ bool wf_geometry(int32_t band_width, uint8_t* pool, int32_t pool_bytes, WfGeom& G)
{
    if (band_width < 256 || (band_width & 127) != 0)
        return false;
    G.S4               = band_width >> 7;
    const int32_t want = max(32, (4 * G.S4 + 1 + 7) & ~7);
    int32_t fit        = ((pool_bytes - 256) / 33 - 20) / 20;
    fit &= ~7;
    G.Wg = min(want, fit);
    if (G.Wg < 32)
        return false;
    G.R = min((G.Wg - 1) / G.S4, 250);
    if (G.R < 2)
        return false;
    G.pitch4 = G.Wg * 16 + 16;
    G.pitch1 = G.Wg * 4 + 4;
    G.win4   = smem_u32(pool);
    G.s3r    = G.win4 + 33 * G.pitch4;
    G.l0v    = G.s3r + 33 * G.pitch1;
    return true;
}

// Everything the general (per-lane) paths need
template <typename ScoreT, typename SizeT>
struct WfCtx
{
    const Win<SizeT>* g;
    const Band<ScoreT>* B;
    WfGeom G;
    int32_t bw, max_column, gap;
};

// Candidate values t0..t3 of one group for a lane whose row is not "fast" at this group: any number of predecessors, each
// one in a window, in HBM (older than R rows; complete there, see the write-back rule), or row 0 (scores[j] = j * gap,
// cudapoa_nw_banded.cuh:269-272); predecessor groups outside that row's band do not contribute (get_scores(), :104-156) and
// when none contributes the candidates are min_score.
template <typename ScoreT, typename SizeT>
__device__ __noinline__ int4 wf_general(const WfCtx<ScoreT, SizeT>& C, int32_t row, int32_t node, int32_t bs, int32_t gx, int32_t sl, int32_t q0,
                                        int32_t q1, int32_t q2, int32_t q3)
{
    constexpr int32_t kMin = min_score_of<ScoreT>();
    const Win<SizeT>& g    = *C.g;
    const int32_t pc       = g.in_cnt[node];
    int32_t M0 = kWfLow, M1 = kWfLow, M2 = kWfLow, M3 = kWfLow, M4 = kWfLow;
    bool any         = false;
    const int32_t np = max(pc, 1);
    for (int32_t p = 0; p < np; p++)
    {
        const int32_t pi  = (pc == 0) ? 0 : static_cast<int32_t>(g.pos[g.in_edge(node, p)]) + 1;
        const int32_t bsp = C.B->start(pi);
        const int32_t lim = min(C.bw - kCPT, C.max_column - bsp);
        const int32_t o   = (bs - bsp) + 4 * gx;
        if (o > lim)
            continue;
        any = true;
        int32_t b0, b1, b2, b3, b4;
        const int32_t d = row - pi;
        if (pi == 0)
        {
            b0 = o * C.gap;
            b1 = b0 + C.gap;
            b2 = b1 + C.gap;
            b3 = b2 + C.gap;
            b4 = b3 + C.gap;
        }
        else if (d <= C.G.R)
        {
            int32_t so = sl - d * C.G.S4;
            if (so < 0)
                so += C.G.Wg;
            const int32_t pl = pi & 31;
            const int4 v     = lds128(C.G.win4 + pl * C.G.pitch4 + (so << 4));
            b0               = v.x;
            b1               = v.y;
            b2               = v.z;
            b3               = v.w;
            b4               = lds32(C.G.s3r + pl * C.G.pitch1 + (so << 2));
        }
        else
        {
            const ScoreT* pp = C.B->scores + static_cast<int64_t>(pi) * C.B->stride + o;
            const int4 v     = __ldcg(reinterpret_cast<const int4*>(pp));
            b0               = v.x;
            b1               = v.y;
            b2               = v.z;
            b3               = v.w;
            b4               = __ldcg(reinterpret_cast<const int32_t*>(pp) + 4);
        }
        M0 = max(M0, b0);
        M1 = max(M1, b1);
        M2 = max(M2, b2);
        M3 = max(M3, b3);
        M4 = max(M4, b4);
    }
    int4 t;
    t.x = any ? __viaddmax_s32(M1, C.gap, M0 + q0) : kMin;
    t.y = any ? __viaddmax_s32(M2, C.gap, M1 + q1) : kMin;
    t.z = any ? __viaddmax_s32(M3, C.gap, M2 + q2) : kMin;
    t.w = any ? __viaddmax_s32(M4, C.gap, M3 + q3) : kMin;
    return t;
}

// first_element_prev_score (cudapoa_nw_banded.cuh:293-326) for a row whose record does not already settle it: the maximum of
// the predecessors' local-0 cells (the last 64 rows' are kept in shared memory), + gap.
template <typename ScoreT, typename SizeT>
__device__ __noinline__ int32_t wf_first(const WfCtx<ScoreT, SizeT>& C, int32_t row, int32_t node, int32_t pc)
{
    constexpr int32_t kMin = min_score_of<ScoreT>();
    const Win<SizeT>& g    = *C.g;
    int32_t penalty        = kMin;
    for (int32_t p = 0; p < pc; p++)
    {
        const int32_t pi = static_cast<int32_t>(g.pos[g.in_edge(node, p)]) + 1;
        int32_t v;
        if (row - pi < 64)
            v = lds32(C.G.l0v + ((pi & 63) << 2));
        else
            v = __ldcg(reinterpret_cast<const int32_t*>(C.B->scores + static_cast<int64_t>(pi) * C.B->stride));
        penalty = max(penalty, v);
    }
    return penalty + C.gap;
}

// The DP rows of one alignment as a wavefront. Returns false (nothing written) when the geometry does not fit the band /
// the shared-memory pool; the caller then runs dp_rows_v3.
template <typename ScoreT, typename SizeT>
__device__ bool dp_rows_v4(const Win<SizeT>& g, const int32_t graph_count, const uint8_t* __restrict__ read, const Band<ScoreT>& B,
                           const int32_t band_width, const int32_t max_column, const int32_t gap, const int32_t mismatch, const int32_t match,
                           int4* row_meta, uint8_t* pool, const int32_t pool_bytes, int4* srec)
{
    static_assert(sizeof(ScoreT) == 4, "wavefront rows are built for 32-bit scores");
    constexpr int32_t kMin = min_score_of<ScoreT>();
    WfCtx<ScoreT, SizeT> C;
    if (!wf_geometry(band_width, pool, pool_bytes, C.G))
        return false;
    C.g          = &g;
    C.B          = &B;
    C.bw         = band_width;
    C.max_column = max_column;
    C.gap        = gap;
    const WfGeom& G      = C.G;
    const int32_t lane   = threadIdx.x & 31;
    const int32_t stride = B.stride;
    ScoreT* const scores = B.scores;
    const int32_t S4     = G.S4;
    const int32_t Wg     = G.Wg;
    const int32_t ngroup = band_width >> 2;
    const uint32_t* const read32 = reinterpret_cast<const uint32_t*>(read);

    // row 0: scores[j] = j * gap (:269-272); read back only through the general path (as values) and by the traceback
    for (int32_t j = lane; j < stride; j += 32)
        scores[j] = static_cast<ScoreT>(j * gap);
    // dummy window (entry 32): what a lane without a second / third predecessor reads
    for (int32_t j = lane; j < Wg; j += 32)
    {
        sts128(G.win4 + 32 * G.pitch4 + (j << 4), kWfLow, kWfLow, kWfLow, kWfLow);
        sts32(G.s3r + 32 * G.pitch1 + (j << 2), kWfLow);
    }

    // ---- per-row records, prepared one 32-row group ahead (lane k <-> row r0 + 32 + k), shared-memory ring of 64
    //   x: band start | base << 16 | min(in-degree, 7) << 24 | "general row" << 27 | "first is min_score + gap" << 28
    //   y: distance to predecessor 0 | 1 << 8 | 2 << 16 (rows)
    //   z: last own group at which every predecessor group is inside its band (may be -1: none)
    //   w: node
    // row_meta (global; end-cell search and traceback): {node, pred0 row, pred1 row, base | in-degree << 8 | sink << 16 | band start / 4 << 17}
    auto produce = [&](int32_t row, int32_t node, int32_t base, int32_t pc, int32_t oc, int32_t p0, int32_t p1, int32_t p2, int4* dst) {
        const int32_t bs = B.start(row);
        bool general     = (pc == 0 || pc > 3);
        bool simple      = true;
        int32_t ginv     = 32767;
        int32_t dd[3]    = {0, 0, 0};
        const int32_t pr[3] = {p0, p1, p2};
#pragma unroll
        for (int32_t k = 0; k < 3; k++)
        {
            if (k < pc)
            {
                const int32_t d = row - pr[k];
                dd[k]           = min(d, 255);
                if (d > G.R)
                    general = true;
                const int32_t bsp = B.start(pr[k]);
                const int32_t lim = min(band_width - kCPT, max_column - bsp);
                ginv              = min(ginv, (lim - (bs - bsp)) >> 2);
                if (bsp == 0)
                    simple = false;
            }
        }
        if (bs > kCPT && pc == 1)
            simple = true;
        if (pc > 3)
            simple = false;
        ginv = max(ginv, -1);
        int4 r;
        r.x           = bs | (base << 16) | (min(pc, 7) << 24) | ((general ? 1 : 0) << 27) | ((simple ? 1 : 0) << 28);
        r.y           = dd[0] | (dd[1] << 8) | (dd[2] << 16);
        r.z           = ginv;
        r.w           = node;
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
            const int32_t p2   = pc > 2 ? static_cast<int32_t>(g.pos[g.in_edge(node, 2)]) + 1 : 0;
            produce(row, node, g.nodes[node], pc, g.out_cnt[node], p0, p1, p2, &srec[lane]);
        }
    }
    __syncwarp();

    // ---- per-lane state of the row in flight
    int32_t cnt    = 0;  // groups left in my row (0 = idle)
    int32_t npend  = 0;  // my groups in the window that have not been written to HBM yet
    int32_t pend0  = 0;  // slot of the first of them
    int32_t goff   = 0;  // and its cell offset in `scores`
    int32_t row    = 0, node = 0, bs = 0, base = 0, pc3 = 0;
    int32_t carry  = 0, stl = 0;
    int32_t rdoff  = 0;  // group index A = it + rdoff
    int32_t i_inv  = -1; // last iteration at which my row is "fast"
    uint32_t pw4_0 = 0, pw4_1 = 0, pw4_2 = 0, ps3_0 = 0, ps3_1 = 0, ps3_2 = 0;
    int32_t ko0 = 0, ko1 = 0, ko2 = 0;
    const uint32_t own4 = G.win4 + lane * G.pitch4;
    const uint32_t own1 = G.s3r + lane * G.pitch1;
    int32_t npass       = 1; // warp-uniform: largest min(in-degree, 3) among the rows in flight

    int32_t it = 0; // iteration counter (uniform); row r computes group it - r * S4
    int32_t sl = 0; // it % Wg

    auto flush = [&]() {
        // the pending groups of a lane are the slots [pend0, pend0 + npend) of the current block of 8 (never wrapping)
        const int32_t meta = (pend0 << 8) | npend;
        const int32_t c    = lane & 7;
#pragma unroll
        for (int32_t j = 0; j < 8; j++)
        {
            const int32_t owner = 4 * j + (lane >> 3);
            const int32_t om    = __shfl_sync(kFull, meta, owner);
            const int32_t og    = __shfl_sync(kFull, goff, owner);
            if (c < (om & 0xff))
            {
                const int4 v = lds128(G.win4 + owner * G.pitch4 + (((om >> 8) + c) << 4));
                *reinterpret_cast<int4*>(scores + og + 4 * c) = v;
            }
        }
        goff += 4 * npend;
        npend = 0;
    };

    // U consecutive iterations of every lane as one straight-line piece: all loads of the U groups (read characters,
    // predecessor groups) are issued before the first dependent instruction, only the horizontal carry runs serially
    // through the 4 * U cells. Legal because what iteration it + k reads was written at iteration it + k - d * S4 <= it - 1
    // for U <= S4; the caller also keeps a chunk inside one block of 8 slots (write-back granularity).
    auto chunk = [&](auto UC) {
        constexpr int32_t U = decltype(UC)::value;
        __syncwarp();
        if (cnt > 0)
        {
            uint32_t w[U];
            int4 v0[U], v1[U], v2[U];
            int32_t m0[U], m1[U], m2[U];
#pragma unroll
            for (int32_t k = 0; k < U; k++)
            {
                w[k]       = (k < cnt) ? __ldg(read32 + (it + k + rdoff)) : 0u;
                int32_t so = sl + k - ko0;
                so += (so >> 31) & Wg;
                v0[k] = lds128(pw4_0 + (so << 4));
                m0[k] = lds32(ps3_0 + (so << 2));
            }
            if (npass >= 2)
            {
#pragma unroll
                for (int32_t k = 0; k < U; k++)
                {
                    int32_t so = sl + k - ko1;
                    so += (so >> 31) & Wg;
                    v1[k] = lds128(pw4_1 + (so << 4));
                    m1[k] = lds32(ps3_1 + (so << 2));
                }
                if (npass >= 3)
                {
#pragma unroll
                    for (int32_t k = 0; k < U; k++)
                    {
                        int32_t so = sl + k - ko2;
                        so += (so >> 31) & Wg;
                        v2[k] = lds128(pw4_2 + (so << 4));
                        m2[k] = lds32(ps3_2 + (so << 2));
                    }
#pragma unroll
                    for (int32_t k = 0; k < U; k++)
                    {
                        v0[k].x = __vimax3_s32(v0[k].x, v1[k].x, v2[k].x);
                        v0[k].y = __vimax3_s32(v0[k].y, v1[k].y, v2[k].y);
                        v0[k].z = __vimax3_s32(v0[k].z, v1[k].z, v2[k].z);
                        v0[k].w = __vimax3_s32(v0[k].w, v1[k].w, v2[k].w);
                        m0[k]   = __vimax3_s32(m0[k], m1[k], m2[k]);
                    }
                }
                else
                {
#pragma unroll
                    for (int32_t k = 0; k < U; k++)
                    {
                        v0[k].x = max(v0[k].x, v1[k].x);
                        v0[k].y = max(v0[k].y, v1[k].y);
                        v0[k].z = max(v0[k].z, v1[k].z);
                        v0[k].w = max(v0[k].w, v1[k].w);
                        m0[k]   = max(m0[k], m1[k]);
                    }
                }
            }
#pragma unroll
            for (int32_t k = 0; k < U; k++)
            {
                if (k < cnt)
                {
                    const int32_t q0 = (static_cast<int32_t>(w[k] & 0xffu) == base) ? match : mismatch;
                    const int32_t q1 = (static_cast<int32_t>((w[k] >> 8) & 0xffu) == base) ? match : mismatch;
                    const int32_t q2 = (static_cast<int32_t>((w[k] >> 16) & 0xffu) == base) ? match : mismatch;
                    const int32_t q3 = (static_cast<int32_t>(w[k] >> 24) == base) ? match : mismatch;
                    int32_t t0, t1, t2, t3;
                    if (it + k <= i_inv)
                    {
                        t0 = __viaddmax_s32(v0[k].y, gap, v0[k].x + q0);
                        t1 = __viaddmax_s32(v0[k].z, gap, v0[k].y + q1);
                        t2 = __viaddmax_s32(v0[k].w, gap, v0[k].z + q2);
                        t3 = __viaddmax_s32(m0[k], gap, v0[k].w + q3);
                    }
                    else
                    {
                        const int4 t = wf_general<ScoreT, SizeT>(C, row, node, bs, it + k + rdoff - (bs >> 2), sl + k, q0, q1, q2, q3);
                        t0           = t.x;
                        t1           = t.y;
                        t2           = t.z;
                        t3           = t.w;
                    }
                    const int32_t s0 = __viaddmax_s32(carry, gap, t0);
                    const int32_t s1 = __viaddmax_s32(s0, gap, t1);
                    const int32_t s2 = __viaddmax_s32(s1, gap, t2);
                    const int32_t s3 = __viaddmax_s32(s2, gap, t3);
                    sts128(own4 + ((sl + k) << 4), stl, s0, s1, s2);
                    sts32(own1 + ((sl + k) << 2), s3);
                    stl   = s3;
                    carry = s3;
                }
            }
            const int32_t done = min(cnt, U);
            if (npend == 0)
                pend0 = sl;
            npend += done;
            cnt -= done;
            if (cnt == 0)
            {
                // last real cell (local band_width) + right padding, straight to HBM (:158-175)
                ScoreT* const tail                 = scores + static_cast<int64_t>(row) * stride + band_width;
                *reinterpret_cast<int4*>(tail)     = make_int4(carry, kMin, kMin, kMin);
                *reinterpret_cast<int4*>(tail + 4) = make_int4(kMin, kMin, kMin, kMin);
            }
        }
        it += U;
        sl += U;
        if (sl == Wg)
            sl = 0;
        if ((sl & 7) == 0)
        {
            __syncwarp();
            flush();
        }
    };
    const int32_t umax = min(4, S4);
    // advance to iteration `until` (or, with until < 0, until no lane has work left)
    auto run_to = [&](int32_t until) {
        for (;;)
        {
            int32_t n = min(umax, 8 - (sl & 7));
            if (until >= 0)
            {
                if (it >= until)
                    break;
                n = min(n, until - it);
            }
            else if (!__any_sync(kFull, cnt > 0))
            {
                break;
            }
            if (n >= 4)
                chunk(std::integral_constant<int32_t, 4>{});
            else if (n == 3)
                chunk(std::integral_constant<int32_t, 3>{});
            else if (n == 2)
                chunk(std::integral_constant<int32_t, 2>{});
            else
                chunk(std::integral_constant<int32_t, 1>{});
        }
    };

    // iteration at which row 1 starts
    {
        const int4 rc1 = srec[0];
        it             = S4 + ((rc1.x & 0xffff) >> 2);
        sl             = it % Wg;
    }
    int32_t t_node = 0, t_base = 0, t_pc = 0, t_oc = 1, t_e0 = 0, t_e1 = 0, t_e2 = 0;
    for (int32_t r = 1; r <= graph_count; r++)
    {
        const int32_t k = (r - 1) & 31;
        // next group's records: three dependent global loads spread over the rows of this group
        {
            const int32_t nrow = ((r - 1) & ~31) + 33 + lane;
            if (nrow <= graph_count)
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
                    t_e2   = g.in_edge(t_node, 2);
                }
                else if (k == 20)
                {
                    t_e0 = t_pc > 0 ? static_cast<int32_t>(g.pos[t_e0]) + 1 : 0; // predecessor rows
                    t_e1 = t_pc > 1 ? static_cast<int32_t>(g.pos[t_e1]) + 1 : 0;
                    t_e2 = t_pc > 2 ? static_cast<int32_t>(g.pos[t_e2]) + 1 : 0;
                }
                else if (k == 28)
                {
                    produce(nrow, t_node, t_base, t_pc, t_oc, t_e0, t_e1, t_e2, &srec[(nrow - 1) & 63]);
                }
            }
        }
        const int4 rc     = srec[(r - 1) & 63];
        const int32_t rbs = rc.x & 0xffff;
        const int32_t I_r = r * S4 + (rbs >> 2);
        run_to(I_r);
        __syncwarp();
        if (lane == (r & 31))
        {
            // ---- my previous row is complete: what is left of it in the window goes out now, then I take over row r
            for (int32_t c = 0; c < npend; c++)
            {
                const int4 v = lds128(own4 + ((pend0 + c) << 4));
                *reinterpret_cast<int4*>(scores + goff + 4 * c) = v;
            }
            npend = 0;
            row   = r;
            node  = rc.w;
            bs    = rbs;
            base  = (rc.x >> 16) & 0xff;
            const int32_t pc   = (rc.x >> 24) & 0x7;
            pc3                = min(pc, 3);
            const bool general = (rc.x >> 27) & 1;
            const int32_t d0 = rc.y & 0xff, d1 = (rc.y >> 8) & 0xff, d2 = (rc.y >> 16) & 0xff;
            i_inv = general ? -1 : it + rc.z;
            {
                const int32_t l0 = (r - d0) & 31;
                pw4_0            = G.win4 + l0 * G.pitch4;
                ps3_0            = G.s3r + l0 * G.pitch1;
                ko0              = d0 * S4;
                const int32_t l1 = (pc >= 2) ? ((r - d1) & 31) : 32;
                pw4_1            = G.win4 + l1 * G.pitch4;
                ps3_1            = G.s3r + l1 * G.pitch1;
                ko1              = (pc >= 2) ? d1 * S4 : 0;
                const int32_t l2 = (pc >= 3) ? ((r - d2) & 31) : 32;
                pw4_2            = G.win4 + l2 * G.pitch4;
                ps3_2            = G.s3r + l2 * G.pitch1;
                ko2              = (pc >= 3) ? d2 * S4 : 0;
            }
            // column "-1" / first_element_prev_score (:293-326)
            int32_t first = 0;
            if (pc != 0)
                first = ((rc.x >> 28) & 1) ? (kMin + gap) : wf_first<ScoreT, SizeT>(C, r, node, g.in_cnt[node]);
            stl   = (bs == 0) ? (pc == 0 ? gap : first) : kMin;
            carry = (pc == 0) ? 0 : first;
            sts32(G.l0v + ((r & 63) << 2), stl);
            rdoff = -r * S4;
            goff  = r * stride;
            cnt   = ngroup;
        }
        __syncwarp();
        npass = __reduce_max_sync(kFull, cnt > 0 ? pc3 : 0);
    }
    run_to(-1);
    __syncwarp();
    // ---- everything that is still in the windows; the traceback reads HBM
    flush();
    __threadfence();
    __syncwarp();
    return true;
}

} // namespace poa
} // namespace gwb200
