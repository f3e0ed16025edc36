// gw-b200 banded Myers, skewed ("systolic") formulation of the score pass -- lane-level core, host + device.
//
// What it computes is the reference's band (myers_compute_scores_edit_dist_banded, cudaaligner/src/myers_gpu.cu:753-846):
// the same edit-distance values D(row, column) for every cell of the band, with the same worst-case assumptions at the band
// edges (horizontal delta +1 above the band's first row, :647-649, :694-696; vertical delta +1 for the row that enters at the
// bottom of a diagonal step, :721-726). How it is laid out is different, because the reference's formulation (lane = word of the
// band, the whole warp on ONE column, the multi-word addition carried across the lanes) is a chain of ~70 dependent
// instructions and three cross-lane operations per column:
//
//   * rows are FIXED to 64-bit blocks of the query (block B = matrix rows 64B .. 64B+63) instead of sliding with the band,
//     so a diagonal step needs no cross-lane shift; the band is a pair of masks per block and column;
//   * the only dependency between blocks of one column is the horizontal delta at the block boundary (Myers' block
//     recurrence: hin in {-1, 0, +1}), so block B can work on column t while block B-1 is already K columns ahead:
//     lane = block (mod the number of blocks in flight), step s = batch of K columns + block index. Per step a lane
//     receives K two-bit deltas and one score from the lane above (two shuffles per K columns) and then runs K columns
//     on registers only;
//   * per block and column it stores pv, mv (64 bits each) and the score of the block's last row, in a layout that makes
//     the stores of one step contiguous across the lanes ([step][16-byte chunk][lane]).
//
// The backtrace reads D(i, j) through score_at() below; its walk (preference order, implicit worst-case values outside the
// band) is the unchanged reference walk of myers_kernels.cuh.
//
// This header is also compiled by g++ into the CPU model test (tests/cpp/myers_skew_model.cpp), which runs exactly these
// functions lane by lane and compares score_at() with the oracle's get_myers_score() for every cell of the band.
#pragma once

#include <cstdint>

#ifdef __CUDACC__
#define GWB_HD __host__ __device__ __forceinline__
#else
#define GWB_HD inline
#endif

namespace gwb200
{
namespace myers
{
namespace skew
{

constexpr int32_t kK       = 8;            // columns per lane step
constexpr int32_t kChunks  = kK + kK / 4;  // 16-byte chunks per (step, lane): K x {pv, mv}, then K scores
constexpr int32_t kMinBand = 128;          // narrower bands (a block could touch both band edges) stay on the classic path
constexpr int32_t kMaxLanes = 32;

GWB_HD int32_t popc64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// bits [0, n) set; n <= 0 -> none, n >= 64 -> all. Two 32-bit halves, no branches.
GWB_HD uint32_t low_mask32(int32_t n)
{
    // n already clamped to [0, 32]
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0), "r"(n));
    return r;
#else
    return n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
#endif
}
GWB_HD uint64_t low_mask(int32_t n)
{
    const int32_t a = n < 0 ? 0 : (n > 32 ? 32 : n);
    const int32_t m = n - 32;
    const int32_t b = m < 0 ? 0 : (m > 32 ? 32 : m);
    return static_cast<uint64_t>(low_mask32(a)) | (static_cast<uint64_t>(low_mask32(b)) << 32);
}

struct Geom
{
    int32_t bw;        // rows of the band
    int32_t q, tsize;  // query / target length
    int32_t db;        // diagonal_begin: first column of the diagonal phase (target_size + 1: none)
    int32_t qmb;       // q - bw: the band's first row in the last phase
    int32_t nbl;       // blocks in flight = lanes used
    int32_t n_batches; // ceil((tsize + 1) / K): columns 0 .. tsize
    int32_t n_steps;   // n_batches + index of the last block
    int32_t last_block;
    // first matrix row (0-based) of the band at column t: 0 in the first phase, one more per column in the diagonal phase,
    // q - bw in the last one (pattern offsets of myers_gpu.cu:826-845)
    GWB_HD int32_t top(int32_t t) const
    {
        int32_t x = t - db + 1;
        x         = x < 0 ? 0 : x;
        return x < qmb ? x : qmb;
    }
};

GWB_HD Geom make_geom(int32_t band_width, int32_t query_size, int32_t target_size, int32_t diagonal_begin)
{
    Geom g;
    g.bw         = band_width < query_size ? band_width : query_size;
    g.q          = query_size;
    g.tsize      = target_size;
    g.db         = band_width >= query_size ? target_size + 1 : diagonal_begin;
    g.qmb        = query_size - g.bw;
    g.last_block = (query_size - 1) / 64;
    // a batch of K columns moves the band by at most K - 1 rows: bw + K - 1 rows, not aligned to blocks
    int32_t nbl  = (g.bw + kK - 2) / 64 + 2;
    if (nbl > g.last_block + 1)
        nbl = g.last_block + 1;
    g.nbl       = nbl;
    g.n_batches = (target_size + 1 + kK - 1) / kK;
    g.n_steps   = g.n_batches + g.last_block;
    return g;
}

// 32-bit words of workspace the records of one pass occupy
GWB_HD int64_t words_needed(const Geom& g) { return static_cast<int64_t>(g.n_steps) * kChunks * g.nbl * 4; }

GWB_HD bool usable(const Geom& g) { return g.bw >= kMinBand && g.nbl <= kMaxLanes; }

// index (in 16-byte units) of chunk `chunk` of the record of block B, column t
GWB_HD int64_t chunk_index(const Geom& g, int32_t B, int32_t t, int32_t chunk)
{
    const int32_t lane = B % g.nbl;
    const int32_t s    = t / kK + B;
    return (static_cast<int64_t>(s) * kChunks + chunk) * g.nbl + lane;
}

struct LaneState
{
    uint64_t pv, mv;
    int32_t S; // D(64B + 63, t): score of the block's last row (valid while that row is inside the band)
    int32_t B; // the block this lane works on
};

GWB_HD void lane_init(LaneState& L, int32_t B)
{
    L.pv = ~0ull; // column 0: D(r, 0) = r + 1; also the worst case for rows that are still below the band
    L.mv = 0ull;
    L.S  = 64 * B + 64;
    L.B  = B;
}

// One column of block L.B (straight-line code: the lanes of a warp are at different places of the band). eq: match bits of the
// block's rows against target[t - 1]; in_pos / in_neg: the horizontal delta of row 64B - 1 (from the lane above) is +1 / -1;
// hold: column 0 of the matrix (eq = 0 and no horizontal delta leave the initial state -- pv all ones -- as it is).
// out_pos / out_neg: horizontal delta of row 64B + 63. L.S follows that delta.
GWB_HD void column(const Geom& g, LaneState& L, int32_t t, uint64_t eq, bool in_pos, bool in_neg, bool hold, uint32_t& out_pos, uint32_t& out_neg)
{
    const int32_t lo     = g.top(t) - 64 * L.B; // first bit of the block inside the band (<= 0: from bit 0)
    const int32_t hi     = lo + g.bw;           // one past the last bit inside the band (>= 64: through bit 63)
    const uint64_t keep  = ~low_mask(lo);       // rows above the band: pv = mv = eq = 0 -> their ph is 1: +1 into the first band row
    const uint64_t below = ~low_mask(hi);       // rows below the band
    // the band's first row takes +1 from above (the block above may already have been retired by its lane)
    const bool pos = !hold && (lo == 0 || (lo < 0 && in_pos));
    const bool neg = !hold && lo < 0 && in_neg;
    uint64_t pv = L.pv & keep, mv = L.mv & keep;
    uint64_t e  = eq & keep;
    const uint64_t xv = e | mv;
    e |= neg ? 1ull : 0ull;
    const uint64_t xh = (((e & pv) + pv) ^ pv) | e;
    uint64_t ph       = mv | ~(xh | pv);
    uint64_t mh       = pv & xh;
    out_pos           = static_cast<uint32_t>(ph >> 63);
    out_neg           = static_cast<uint32_t>(mh >> 63);
    ph                = (ph << 1) | (pos ? 1ull : 0ull);
    mh                = (mh << 1) | (neg ? 1ull : 0ull);
    pv                = mh | ~(xv | ph);
    mv                = ph & xv;
    // rows below the band keep the worst case: the row that enters next finds vertical delta +1 (myers_gpu.cu:721-726)
    L.pv = pv | below;
    L.mv = mv & ~below;
    L.S += static_cast<int32_t>(out_pos) - static_cast<int32_t>(out_neg);
}

// What a lane hands to the lane below after a step
struct Link
{
    uint32_t hbits; // bit k: the horizontal delta of the block's last row in column k of the batch is +1; bit 8 + k: it is -1
    int32_t S0;     // score of the block's last row in the first column of the batch
};

// One step of a lane: batch cb (columns K cb .. K cb + K - 1) of block L.B. eqs[k]: match bits against target[K cb + k - 1].
// rec_pvmv[k] = {pv, mv} and rec_S[k] of the K columns, for the record store. Columns beyond the target run like the others
// (nobody reads them).
GWB_HD Link lane_step(const Geom& g, LaneState& L, int32_t cb, const uint64_t* eqs, Link in, uint64_t (*rec_pvmv)[2], int32_t* rec_S)
{
    Link out;
    out.hbits = 0;
    const int32_t t0 = kK * cb;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int32_t k = 0; k < kK; k++)
    {
        const bool hold = (k == 0) && cb == 0;
        uint32_t op, on;
        column(g, L, t0 + k, hold ? 0ull : eqs[k], ((in.hbits >> k) & 1u) != 0u, ((in.hbits >> (8 + k)) & 1u) != 0u, hold, op, on);
        out.hbits += (op << k) + (on << (8 + k));
        rec_pvmv[k][0] = L.pv;
        rec_pvmv[k][1] = L.mv;
        rec_S[k]       = L.S;
    }
    // The score follows the horizontal delta of the block's last row, which means something only while that row is inside the
    // band. Once per block the row enters the band (at the bottom, during the diagonal phase): from that column on the score
    // is the score of the row above the block -- the lane above is exact there: T = its S0 + its deltas -- plus the block's
    // vertical deltas.
    const int32_t hi_before = (cb == 0 ? 0 : g.top(t0 - 1)) - 64 * L.B + g.bw;
    const int32_t hi_last   = g.top(t0 + kK - 1) - 64 * L.B + g.bw;
    if (hi_before < 64 && hi_last >= 64)
    {
        int32_t T = (L.B == 0) ? t0 : in.S0; // D(64B - 1, t0): matrix row "-1" is D(0, t) = t
        for (int32_t k = 0; k < kK; k++)
        {
            if (k > 0)
                T += (L.B == 0) ? 1 : static_cast<int32_t>((in.hbits >> k) & 1u) - static_cast<int32_t>((in.hbits >> (8 + k)) & 1u);
            if (g.top(t0 + k) - 64 * L.B + g.bw >= 64)
                rec_S[k] = T + popc64(rec_pvmv[k][0]) - popc64(rec_pvmv[k][1]);
        }
        L.S = rec_S[kK - 1];
    }
    out.S0 = rec_S[0];
    return out;
}

// Does block B have a row inside the band in some column of batch cb? (only then anybody reads its records: a lane that has
// moved on to its next block runs a number of batches before the band arrives there)
GWB_HD bool block_in_band(const Geom& g, int32_t B, int32_t cb)
{
    return g.top(kK * cb) <= 64 * B + 63 && g.top(kK * cb + kK - 1) + g.bw - 1 >= 64 * B;
}

// Does the lane leave its block before batch cb? (the block lies above the band from the first column of the batch on)
GWB_HD bool block_retired(const Geom& g, int32_t B, int32_t cb) { return g.top(kK * cb) > 64 * B + 63; }

// D(i, j) for row i (1-based) of the band of column j -- get_myers_score (myers_gpu.cu:243-255) on the block records.
// Loader: pvmv(B, j, pv, mv), S(B, j).
template <typename Loader>
GWB_HD int32_t score_at(const Geom& g, int32_t i, int32_t j, const Loader& ld)
{
    const int32_t top    = g.top(j);
    const int32_t r      = top + i - 1; // matrix row, 0-based
    const int32_t B      = r >> 6;
    const int32_t b      = r & 63;
    const int32_t bottom = top + g.bw - 1;
    uint64_t pv, mv;
    ld.pvmv(B, j, pv, mv);
    if (64 * B + 63 <= bottom)
    {
        // from the block's last row upwards
        const uint64_t m = (b == 63) ? 0ull : (~0ull << (b + 1));
        return ld.S(B, j) - popc64(pv & m) + popc64(mv & m);
    }
    // the block reaches below the band: from the last row of the block above downwards
    const uint64_t m = low_mask(b + 1);
    const int32_t T  = (B == 0) ? j : ld.S(B - 1, j);
    return T + popc64(pv & m) - popc64(mv & m);
}

} // namespace skew
} // namespace myers
} // namespace gwb200
