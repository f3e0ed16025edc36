// gw-b200: host engine + C ABI for the POA path (see include/gwb200.h for the reference interface each entry replaces).
//
// Host-side behaviour follows cudapoa/src/cudapoa_batch.cuh (admission, packing, status decoding) and
// cudapoa/src/batch.cu (BatchConfig derivation, type selection via cudapoa_limits.hpp); memory layout and kernels are
// this repo's own (poa_kernels.cuh). There is no CPU fallback: every entry needs a CUDA device.

#include "../../include/gwb200.h"
#include "common.cuh"
#include "poa_kernels.cuh"
#include "poa_kernels_v2.cuh"
#include "poa_kernels_v3.cuh"

#include <atomic>
#include <cstdlib>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using namespace gwb200;
using namespace gwb200::poa;

namespace
{

inline int32_t align_up(int32_t v, int32_t b) { return (v + b - 1) & ~(b - 1); }
inline int64_t align_up64(int64_t v, int64_t b) { return (v + b - 1) / b * b; }

// cudapoa_limits.hpp:34-59
bool use32bit_score(const gwb200_poa_config& c, int32_t gap, int32_t mismatch, int32_t match)
{
    int32_t upper = c.max_sequence_size * match;
    int32_t lower = c.max_sequence_size * std::max(gap, mismatch) + (c.max_nodes_per_graph - c.max_sequence_size) * gap;
    return (upper > INT16_MAX || (-lower) > (INT16_MAX + 1));
}
bool use32bit_size(const gwb200_poa_config& c)
{
    int32_t m = std::max(c.max_consensus_size, std::max(c.max_nodes_per_graph, c.matrix_sequence_dimension));
    return m > INT16_MAX;
}

struct Carver
{
    uint8_t* base;
    int64_t off = 0;
    template <typename T>
    T* take(int64_t count)
    {
        off      = align_up64(off, 256);
        T* p     = reinterpret_cast<T*>(base + off);
        off += count * static_cast<int64_t>(sizeof(T));
        return p;
    }
};

} // namespace

struct gwb200_poa_batch
{
    int32_t device_id = 0;
    cudaStream_t stream = nullptr;
    int8_t output_mask  = 0;
    gwb200_poa_config cfg{};
    int32_t gap = -8, mismatch = -6, match = 8;
    bool score32 = false, size32 = false, msa = false;
    bool accurate = false; // SPOA_ACCURATE semantics (racon sort after every read)
    int32_t score_bytes = 2, size_bytes = 2;
    int32_t bid = 0;

    int32_t max_poas = 0;
    int32_t poa_count = 0;
    int32_t num_nucleotides_copied = 0;
    int32_t global_sequence_idx = 0;
    int64_t avail_buf_mem = 0, scorebuf_alloc_size = 0;
    int64_t next_scores_offset = 0;
    int64_t seq_capacity = 0; // bytes in the sequences / weights buffers (without slack)

    // pinned host
    uint8_t* h_block = nullptr;
    uint8_t* h_sequences = nullptr;
    int8_t* h_weights = nullptr;
    bool weights_present = false; // false: every read so far had unit base weights (h_weights is not filled, the device array is memset)
    struct CopyJob
    {
        uint8_t* dst;
        const char* src;
        int32_t len;
    };
    std::vector<CopyJob>* deferred = nullptr; // set by add_groups_flat: sequence copies are collected and run on several threads
    int32_t* h_seq_lengths = nullptr;
    WindowInfo* h_windows = nullptr;
    uint8_t* h_consensus = nullptr;
    uint16_t* h_coverage = nullptr;
    int32_t* h_cons_len = nullptr;
    int32_t* h_status = nullptr;
    int32_t* h_node_count = nullptr;
    unsigned long long* h_cells = nullptr;
    uint8_t* h_msa = nullptr;

    // device
    uint8_t* d_block = nullptr;
    bool owns_dblock = true; // false: the block belongs to the caller's allocator (create_batch overload with an allocator)
    DeviceParams P{};
    V2Extra X{};
    V3Extra Y{};
    bool use_v2 = true;
    bool use_v3 = true;      // third-generation kernel (one warp per window, persistent grid); v2/v1 remain for the traceback band
                             // modes, bands wider than 1536 and reads of 64 k bases and more
    int32_t v3_ctas_per_sm = 0;
    bool tb_mode = false;    // static_band_traceback / adaptive_band_traceback
    int32_t trace_bytes = 2; // sizeof(TraceT): 1 unless max_banded_pred_distance > 127 (cudapoa_limits.hpp:56-60)
    int32_t nw_override = 0;
    bool timers_on = false;
    unsigned long long* d_timers = nullptr;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool launched = false;
    bool results_on_host = false;

    static int32_t batches;
};
int32_t gwb200_poa_batch::batches = 0;

namespace
{

struct Sizes
{
    int64_t dev_per_poa = 0, dev_per_matrix = 0, host_per_poa = 0;
    int64_t seq_bytes_per_poa = 0;
    int32_t aln_capacity = 0, stack_capacity = 0;
};

// SPOA_ACCURATE as a library-wide switch (the reference makes it a build flag, cudapoa_kernels.cuh:508-520): batches created while
// it is on re-sort their graphs with racon's topological sort after every read
std::atomic<int> g_spoa_accurate{-1};
bool spoa_accurate()
{
    int v = g_spoa_accurate.load();
    if (v < 0)
    {
        const char* e = std::getenv("GWB200_SPOA_ACCURATE");
        v             = (e && std::atoi(e) != 0) ? 1 : 0;
    }
    return v != 0;
}

Sizes compute_sizes(const gwb200_poa_config& c, int32_t score_bytes, int32_t sz, bool msa, bool tb = false, int32_t trace_bytes = 2)
{
    Sizes s;
    const int64_t mn = c.max_nodes_per_graph;
    const int64_t E  = kMaxEdges;
    s.aln_capacity   = c.max_nodes_per_graph + c.max_sequence_size + 16;
    const bool accurate = spoa_accurate();
    s.stack_capacity = (msa || accurate) ? 4 * c.max_nodes_per_graph + 16 : 0;
    s.seq_bytes_per_poa = static_cast<int64_t>(c.max_sequences_per_poa) * align_up(std::max(c.max_sequence_size, 1), 4);
    int64_t d = 0;
    d += mn * 1;                 // nodes
    d += mn * 2 * 5;             // in_cnt out_cnt aln_cnt cov local_cnt
    d += mn * E * sz * 2;        // in_edges, out_edges
    d += mn * kMaxAligned * sz;  // aligned
    d += mn * E * 2;             // in_w
    d += mn * sz * 2;            // sorted, pos
    d += 2ll * s.aln_capacity * sz;
    d += mn * 4 + mn * sz;       // consensus scratch
    d += s.seq_bytes_per_poa * 2;              // sequences + weights
    d += 4ll * c.max_sequences_per_poa;        // seq lengths
    d += sizeof(WindowInfo);
    d += c.max_consensus_size * 3ll;           // consensus + coverage
    d += 4 * 3 + 8;                            // len, status, node_count, cells
    if (msa)
    {
        d += s.seq_bytes_per_poa * sz; // path
        d += mn * sz + mn * 2;         // msa_col, marks, check
        d += static_cast<int64_t>(s.stack_capacity) * sz;
        d += static_cast<int64_t>(c.max_sequences_per_poa) * c.max_consensus_size;
    }
    else if (accurate)
    {
        d += mn * 2 + static_cast<int64_t>(s.stack_capacity) * sz; // marks, check, stack of the racon sort
    }
    d += (mn + 1) * 16;                                   // v2 row metadata
    d += (align_up(std::max(c.max_sequence_size, 1), 4) + 8ll) * sz; // v2 read -> node map
    d += 64;                                             // phase timers
    d += 256 * 36; // carving alignment slack
    if (tb)
        d += static_cast<int64_t>(c.matrix_sequence_dimension) * c.max_banded_pred_distance * score_bytes + 256; // score ring (allocate_block.hpp:333-334,374)
    s.dev_per_poa    = d;
    // the pooled matrix: scores, or the trace matrix in the traceback modes (allocate_block.hpp:76-84)
    s.dev_per_matrix = static_cast<int64_t>(c.matrix_sequence_dimension) * mn * (tb ? trace_bytes : score_bytes);
    int64_t h        = s.seq_bytes_per_poa * 2 + 4ll * c.max_sequences_per_poa + sizeof(WindowInfo) + c.max_consensus_size * 3ll + 4 * 3 + 8;
    if (msa)
        h += static_cast<int64_t>(c.max_sequences_per_poa) * c.max_consensus_size;
    s.host_per_poa = h + 64;
    return s;
}

// Kernel selection for a batch: warps per window and band chunks per warp.
struct V2Choice
{
    int32_t nw, maxc;
};
V2Choice choose_v2(const gwb200_poa_batch* b)
{
    if (b->tb_mode)
        return {1, 1}; // the traceback-matrix alignment is a one-warp routine (poa_kernels_tb.cuh)
    // warps per window: one per 128-column band chunk, at most 4; chunks per warp bounded by the widest band the mode can
    // reach (adaptive bands grow up to 1536 = 12 chunks)
    const bool adaptive   = b->cfg.band_mode == GWB200_POA_ADAPTIVE_BAND && b->cfg.alignment_band_width < kMaxAdaptiveBW;
    const int32_t nchunks = adaptive ? kMaxAdaptiveBW / 128 : std::max(1, b->cfg.alignment_band_width / 128);
    // measured on B200 (profiles/): up to 2 chunks one warp is fastest (no CTA barriers); wider bands use 4 warps
    int32_t nw = nchunks >= 3 ? 4 : 1;
    if (b->cfg.band_mode == GWB200_POA_FULL_BAND)
        nw = 1;
    if (b->nw_override > 0 && !adaptive && nchunks <= 4)
        nw = b->nw_override; // development switch
    if (b->nw_override > 0 && adaptive)
        return b->nw_override == 1 ? V2Choice{1, 12} : (b->nw_override == 2 ? V2Choice{2, 6} : V2Choice{4, 3}); // development switch
    if (nw == 1 && nchunks == 2)
        return {1, 2};
    if (nchunks > 4 || nw < std::min(nchunks, 4))
        return {4, 3};
    if (nw == 4)
        return {4, 1};
    if (nw == 2)
        return {2, 1};
    return {1, 1};
}

// Dynamic shared memory per CTA of the v2 kernel: the staged read (max_sequence_size + widest band) plus a ring of score rows;
// the traceback tile and the topological-sort staging reuse it. Defaults: 31 KB (32-bit scores; 7 windows per SM for the
// 4-warp kernels, measured best of 5..8 on C3), 24 KB (16-bit scores), 12 KB (one-warp kernels with 16-bit scores, 16 per SM).
// Reads longer than ~10 kb get a pool that holds the staged read plus two rows of the widest band (fewer resident windows,
// which HBM capacity limits anyway at that size).
// Dynamic shared memory per CTA of the v3 kernel. The resident windows of one SM share its 227 KB: the pool is what one CTA gets
// when the batch's capacity (HBM) is spread over the SMs, at most 16 CTAs per SM (registers) -- a batch that HBM limits to a
// few windows per SM gives each of them a deeper ring of score rows. Never less than two traceback tile buffers or two rows
// of the widest band.
int32_t v3_pool_bytes(gwb200_poa_batch* b)
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device_id);
    int smem_sm = 233472;
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, b->device_id);
    int32_t ctas = static_cast<int32_t>(std::min<int64_t>(16, std::max<int64_t>(1, (static_cast<int64_t>(b->max_poas) + sms - 1) / sms)));
    {
        // 32-bit scores run their rows as a wavefront (poa_kernels_v4.cuh): 32 rows in flight per window need windows of
        // 33 x (Wg x 20 + 20) bytes, Wg = 48 slots for the widest adaptive bands -> 6 windows per SM
        const char* e = std::getenv("GWB200_POA_WAVEFRONT");
        if (b->score32 && e && std::atoi(e) != 0)
            ctas = std::min(ctas, 6);
    }
    if (const char* c = std::getenv("GWB200_POA_CTAS_PER_SM")) // development switch
        ctas = std::max(1, std::min(32, std::atoi(c)));
    if (const char* kb = std::getenv("GWB200_POA_POOL_KB")) // development switch
    {
        b->v3_ctas_per_sm = ctas;
        return std::atoi(kb) * 1024;
    }
    const int64_t max_bw  = (b->cfg.band_mode == GWB200_POA_ADAPTIVE_BAND) ? kMaxAdaptiveBW : b->cfg.alignment_band_width;
    const int64_t need    = std::max<int64_t>(2 * (max_bw + 8) * b->score_bytes, 2 * (b->score32 ? TileBuf<int32_t>::kBytes : TileBuf<int16_t>::kBytes));
    const int64_t statics = sizeof(V3Shared) + 64;
    int64_t pool          = smem_sm / ctas - 1024 - statics;
    pool                  = std::min<int64_t>(pool, 96 * 1024) / 256 * 256;
    pool                  = std::max<int64_t>(pool, align_up64(need, 256));
    b->v3_ctas_per_sm     = ctas;
    return static_cast<int32_t>(pool);
}

template <typename ScoreT, typename SizeT>
int32_t v3_action(gwb200_poa_batch* b, int action)
{
    // the wavefront rows (poa_kernels_v4.cuh, 32-bit scores) live in their own instantiation: its code and register budget
    // do not weigh on the default kernel
    auto kfn = poa_window_kernel_v3<ScoreT, SizeT, true, false>;
    if (b->Y.use_bulk == 0)
        kfn = poa_window_kernel_v3<ScoreT, SizeT, false, false>; // rows leave by per-lane vector stores (A/B switch GWB200_POA_BULK=0)
    if constexpr (sizeof(ScoreT) == 4)
    {
        if (b->Y.wavefront != 0)
            kfn = poa_window_kernel_v3<ScoreT, SizeT, true, true>;
    }
    const int32_t smem = b->X.pool_bytes;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, 32, smem) != cudaSuccess)
    {
        cudaGetLastError();
        nb = 0;
    }
    if (action == 1)
        return nb;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device_id);
    const int32_t grid = std::max(1, std::min(b->poa_count, std::max(1, nb) * sms));
    cudaMemsetAsync(b->Y.work_counter, 0, sizeof(int32_t), b->stream);
    kfn<<<grid, 32, smem, b->stream>>>(b->P, b->X, b->Y);
    return 0;
}

int32_t v2_pool_bytes(const gwb200_poa_batch* b)
{
    const int64_t max_bw = (b->cfg.band_mode == GWB200_POA_ADAPTIVE_BAND) ? kMaxAdaptiveBW : b->cfg.alignment_band_width;
    int64_t dflt         = b->score32 ? 31 * 1024 : 24 * 1024;
    if (choose_v2(b).nw == 1)
        dflt = b->score32 ? 24 * 1024 : 12 * 1024; // one-warp kernels (127 registers, __launch_bounds__(32, 16)): up to 16 windows per SM
    const int64_t need   = (b->cfg.max_sequence_size + max_bw + 24) + 2 * (max_bw + 8) * b->score_bytes;
    if (const char* kb = std::getenv("GWB200_POA_POOL_KB")) // development switch
        return std::atoi(kb) * 1024;
    return static_cast<int32_t>(std::max<int64_t>(dflt, align_up64(need, 1024)));
}

// action 0: launch; action 1: return resident CTAs per SM (occupancy) for the chosen kernel
template <typename ScoreT, typename SizeT, int32_t NW, int32_t MAXC>
int32_t v2_action(gwb200_poa_batch* b, int action)
{
    auto kfn           = poa_window_kernel_v2<ScoreT, SizeT, NW, MAXC>;
    const int32_t smem = b->X.pool_bytes;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    // the kernel's residency is sized against the full shared-memory carve-out; do not inherit a device-wide cache preference
    // another library set in this process (the reference does cudaDeviceSetCacheConfig(PreferL1), cudapoa_kernels.cuh:605)
    cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (action == 1)
    {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, 32 * NW, smem) != cudaSuccess)
        {
            cudaGetLastError();
            return 0;
        }
        return nb;
    }
    kfn<<<b->poa_count, 32 * NW, smem, b->stream>>>(b->P, b->X);
    return 0;
}

template <typename ScoreT, typename SizeT>
int32_t typed_action(gwb200_poa_batch* b, int action)
{
    if (b->use_v3)
    {
        if (action == 0)
        {
            b->X.timers = b->timers_on ? b->d_timers : nullptr;
            if (b->timers_on)
                cudaMemsetAsync(b->d_timers, 0, sizeof(unsigned long long) * 8 * b->poa_count, b->stream);
        }
        const int32_t r = v3_action<ScoreT, SizeT>(b, action);
        if (action == 0)
            count_launch();
        return r;
    }
    if (b->use_v2)
    {
        if (action == 0)
        {
            b->X.timers = b->timers_on ? b->d_timers : nullptr;
            if (b->timers_on)
                cudaMemsetAsync(b->d_timers, 0, sizeof(unsigned long long) * 8 * b->poa_count, b->stream);
        }
        const V2Choice ch = choose_v2(b);
        int32_t r;
        if (ch.nw == 1 && ch.maxc == 2)
            r = v2_action<ScoreT, SizeT, 1, 2>(b, action);
        else if (ch.nw == 1 && ch.maxc == 12)
            r = v2_action<ScoreT, SizeT, 1, 12>(b, action);
        else if (ch.nw == 2 && ch.maxc == 6)
            r = v2_action<ScoreT, SizeT, 2, 6>(b, action);
        else if (ch.nw == 4 && ch.maxc == 3)
            r = v2_action<ScoreT, SizeT, 4, 3>(b, action);
        else if (ch.nw == 4)
            r = v2_action<ScoreT, SizeT, 4, 1>(b, action);
        else if (ch.nw == 2)
            r = v2_action<ScoreT, SizeT, 2, 1>(b, action);
        else
            r = v2_action<ScoreT, SizeT, 1, 1>(b, action);
        if (action == 0)
            count_launch();
        return r;
    }
    if (action == 1)
    {
        int nb = 0;
        auto kfn = poa_window_kernel<ScoreT, SizeT, false>;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kfn, 32, 0) != cudaSuccess)
        {
            cudaGetLastError();
            return 0;
        }
        return nb;
    }
    dim3 grid(b->poa_count), block(32);
    if (b->msa)
        poa_window_kernel<ScoreT, SizeT, true><<<grid, block, 0, b->stream>>>(b->P);
    else
        poa_window_kernel<ScoreT, SizeT, false><<<grid, block, 0, b->stream>>>(b->P);
    count_launch();
    return 0;
}

int32_t batch_action(gwb200_poa_batch* b, int action)
{
    if (!b->score32 && !b->size32)
        return typed_action<int16_t, int16_t>(b, action);
    if (b->score32 && !b->size32)
        return typed_action<int32_t, int16_t>(b, action);
    return typed_action<int32_t, int32_t>(b, action);
}

int validate_config(const gwb200_poa_config& c)
{
    if (c.max_sequence_size < 0 || c.max_consensus_size < 0 || c.max_nodes_per_graph < 0 || c.max_sequences_per_poa < 0 ||
        c.alignment_band_width < 0 || c.matrix_sequence_dimension < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "BatchConfig fields cannot be negative.");
    return 0;
}

} // namespace

extern "C" {

int gwb200_poa_init(void) { return GWB200_POA_SUCCESS; }

int gwb200_poa_config_init(gwb200_poa_config* cfg, int32_t max_seq_sz, int32_t max_seq_per_poa, int32_t band_width, int32_t band_mode,
                           float adaptive_storage_factor, float graph_length_factor, int32_t max_pred_dist)
{
    if (!cfg)
        return set_error(GWB200_E_INVALID_ARGUMENT, "cfg is null");
    if (band_mode < 0 || band_mode > GWB200_POA_ADAPTIVE_BAND_TRACEBACK)
        return set_error(GWB200_E_INVALID_ARGUMENT, "unknown band mode");
    // batch.cu:34-71
    const int32_t abw              = align_up(band_width, kMinBandWidth);
    cfg->max_sequence_size         = max_seq_sz;
    cfg->max_consensus_size        = 2 * max_seq_sz;
    cfg->alignment_band_width      = abw;
    cfg->max_sequences_per_poa     = max_seq_per_poa;
    cfg->band_mode                 = band_mode;
    cfg->max_banded_pred_distance  = max_pred_dist > 0 ? max_pred_dist : 2 * abw;
    cfg->max_nodes_per_graph       = align_up(static_cast<int32_t>(graph_length_factor * max_seq_sz), kCPT);
    if (band_mode == GWB200_POA_FULL_BAND)
        cfg->matrix_sequence_dimension = align_up(max_seq_sz, kCPT);
    else if (band_mode == GWB200_POA_STATIC_BAND || band_mode == GWB200_POA_STATIC_BAND_TRACEBACK)
        cfg->matrix_sequence_dimension = align_up(abw + kRightPad, kCPT);
    else
        cfg->matrix_sequence_dimension = align_up(static_cast<int32_t>(adaptive_storage_factor * (abw + kRightPad)), kCPT);
    if (max_seq_sz < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_sequence_size cannot be negative.");
    if (max_seq_per_poa < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_sequences_per_poa cannot be negative.");
    if (band_width < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "alignment_band_width cannot be negative.");
    if (cfg->max_nodes_per_graph < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_nodes_per_graph cannot be negative.");
    if (abw != band_width)
        fprintf(stderr, "Band-width should be multiple of 128. The input was changed from %d to %d\n", band_width, abw);
    return 0;
}

int gwb200_poa_config_init_explicit(gwb200_poa_config* cfg, int32_t max_seq_sz, int32_t max_consensus_sz, int32_t max_nodes_per_poa,
                                    int32_t band_width, int32_t max_seq_per_poa, int32_t matrix_seq_dim, int32_t band_mode, int32_t max_pred_dist)
{
    if (!cfg)
        return set_error(GWB200_E_INVALID_ARGUMENT, "cfg is null");
    if (band_mode < 0 || band_mode > GWB200_POA_ADAPTIVE_BAND_TRACEBACK)
        return set_error(GWB200_E_INVALID_ARGUMENT, "unknown band mode");
    // batch.cu:73-104
    cfg->max_sequence_size         = max_seq_sz;
    cfg->max_consensus_size        = max_consensus_sz;
    cfg->max_nodes_per_graph       = align_up(max_nodes_per_poa, kCPT);
    cfg->matrix_sequence_dimension = align_up(matrix_seq_dim, kCPT);
    cfg->alignment_band_width      = align_up(band_width, kMinBandWidth);
    cfg->max_sequences_per_poa     = max_seq_per_poa;
    cfg->band_mode                 = band_mode;
    cfg->max_banded_pred_distance  = max_pred_dist;
    if (max_seq_sz < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_sequence_size cannot be negative.");
    if (max_consensus_sz < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_consensus_size cannot be negative.");
    if (max_nodes_per_poa < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_nodes_per_graph cannot be negative.");
    if (max_seq_per_poa < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_sequences_per_poa cannot be negative.");
    if (band_width < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "alignment_band_width cannot be negative.");
    if (max_pred_dist < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_banded_pred_distance cannot be negative.");
    if (cfg->max_nodes_per_graph < cfg->max_sequence_size)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_nodes_per_graph should be greater than or equal to max_sequence_size.");
    if (cfg->max_consensus_size < cfg->max_sequence_size)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_consensus_size should be greater than or equal to max_sequence_size.");
    if (cfg->max_sequence_size < cfg->alignment_band_width)
        return set_error(GWB200_E_INVALID_ARGUMENT, "alignment_band_width should not be greater than max_sequence_size.");
    if (cfg->alignment_band_width != band_width)
        fprintf(stderr, "Band-width should be multiple of 128. The input was changed from %d to %d\n", band_width, cfg->alignment_band_width);
    return 0;
}

int gwb200_poa_decode_error(int32_t status, char* message, int32_t message_len, char* hint, int32_t hint_len)
{
    const char* m = nullptr;
    const char* h = "";
    switch (status)
    {
    case GWB200_POA_EXCEEDED_MAXIMUM_POAS:
        m = "Kernel Error: more groups were added than the batch has room for (maximum POAs).";
        h = "Suggestion  : size the batch with more device memory or split the groups over several batches.";
        break;
    case GWB200_POA_EXCEEDED_MAXIMUM_SEQUENCE_SIZE:
        m = "Kernel Error: an input read, or the output consensus/MSA, is longer than the configured maximum.";
        h = "Suggestion  : raise BatchConfig::max_sequence_size / BatchConfig::max_consensus_size.";
        break;
    case GWB200_POA_EXCEEDED_MAXIMUM_SEQUENCES_PER_POA:
        m = "Kernel Error: too many reads in one POA group.";
        h = "Suggestion  : raise BatchConfig::max_sequences_per_poa.";
        break;
    case GWB200_POA_NODE_COUNT_EXCEEDED_MAXIMUM_GRAPH_SIZE:
        m = "Kernel Error: the POA graph grew beyond the maximum number of nodes.";
        h = "Suggestion  : raise BatchConfig::max_nodes_per_graph.";
        break;
    case GWB200_POA_EDGE_COUNT_EXCEEDED_MAXIMUM_GRAPH_SIZE:
        m = "Kernel Error: a node exceeded the maximum number of edges.";
        h = "Suggestion  : the per-node edge limit is a compile-time constant (50).";
        break;
    case GWB200_POA_EXCEEDED_ADAPTIVE_BANDED_MATRIX_SIZE:
        m = "Kernel Error: the score/traceback buffer is too small for the adaptive band of this window.";
        h = "Suggestion  : raise BatchConfig::matrix_sequence_dimension (adaptive_storage_factor).";
        break;
    case GWB200_POA_EXCEEDED_MAXIMUM_PREDECESSOR_DISTANCE:
        m = "Kernel Error: a predecessor lies further back than max_banded_pred_distance allows in traceback mode.";
        h = "Suggestion  : raise BatchConfig::max_banded_pred_distance.";
        break;
    case GWB200_POA_LOOP_COUNT_EXCEEDED_UPPER_BOUND:
        m = "Kernel Error: Needleman-Wunsch traceback did not terminate.";
        h = "Suggestion  : retry with another banding mode.";
        break;
    case GWB200_POA_OUTPUT_TYPE_UNAVAILABLE:
        m = "Kernel Error: the requested output type was not enabled for this batch.";
        h = "Suggestion  : check the consensus/MSA output mask passed to create_batch.";
        break;
    case GWB200_POA_ZERO_WEIGHTED_POA_SEQUENCE:
        m = "Error      : every base weight of the sequence is zero.";
        h = "Suggestion : check the base weights passed with the POA group.";
        break;
    case GWB200_POA_EMPTY_POA_GROUP:
        m = "Error      : no sequence of the POA group could be added.";
        h = "Suggestion : inspect the per-sequence status from add_poa_group.";
        break;
    case GWB200_POA_GENERIC_ERROR:
        m = "Unknown error.";
        break;
    default:
        return set_error(GWB200_E_RUNTIME, "Unknown error type detected.");
    }
    if (message && message_len > 0)
    {
        std::strncpy(message, m, message_len - 1);
        message[message_len - 1] = 0;
    }
    if (hint && hint_len > 0)
    {
        std::strncpy(hint, h, hint_len - 1);
        hint[hint_len - 1] = 0;
    }
    return 0;
}

/* Library-wide SPOA_ACCURATE switch (cudapoa_kernels.cuh:508-520 makes it a build flag): affects batches created afterwards. */
void gwb200_poa_set_spoa_accurate(int32_t on) { g_spoa_accurate.store(on ? 1 : 0); }
int32_t gwb200_poa_get_spoa_accurate(void) { return spoa_accurate() ? 1 : 0; }

int64_t gwb200_poa_estimate_max_poas(const gwb200_poa_config* cfg, int32_t msa_flag, float gpu_memory_usage_quota, int16_t mismatch_score,
                                     int16_t gap_score, int16_t match_score)
{
    if (!cfg)
        return set_error(GWB200_E_INVALID_ARGUMENT, "cfg is null");
    if (int rc = validate_config(*cfg))
        return rc;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
    {
        cudaGetLastError();
        return set_error(GWB200_E_CUDA, "no usable CUDA device: this engine has no CPU fallback");
    }
    const int32_t score_bytes = use32bit_score(*cfg, gap_score, mismatch_score, match_score) ? 4 : 2;
    const int32_t size_bytes  = use32bit_size(*cfg) ? 4 : 2;
    const bool tb             = cfg->band_mode == GWB200_POA_STATIC_BAND_TRACEBACK || cfg->band_mode == GWB200_POA_ADAPTIVE_BAND_TRACEBACK;
    const Sizes sz            = compute_sizes(*cfg, score_bytes, size_bytes, msa_flag != 0, tb, cfg->max_banded_pred_distance > INT8_MAX ? 2 : 1);
    const int64_t per         = sz.dev_per_poa + sz.dev_per_matrix;
    if (per <= 0)
        return 0;
    return static_cast<int64_t>(static_cast<double>(gpu_memory_usage_quota) * static_cast<double>(free_b)) / per;
}

static int poa_batch_create_impl(gwb200_poa_batch** out, int32_t device_id, void* stream, int64_t max_gpu_mem, int8_t output_mask,
                                 const gwb200_poa_config* cfg, int16_t gap_score, int16_t mismatch_score, int16_t match_score, void* ext_block,
                                 int64_t ext_bytes);

int gwb200_poa_batch_create(gwb200_poa_batch** out, int32_t device_id, void* stream, int64_t max_gpu_mem, int8_t output_mask,
                            const gwb200_poa_config* cfg, int16_t gap_score, int16_t mismatch_score, int16_t match_score)
{
    return poa_batch_create_impl(out, device_id, stream, max_gpu_mem, output_mask, cfg, gap_score, mismatch_score, match_score, nullptr, 0);
}

int gwb200_poa_batch_create_in_block(gwb200_poa_batch** out, int32_t device_id, void* stream, void* device_block, int64_t device_block_bytes,
                                     int8_t output_mask, const gwb200_poa_config* cfg, int16_t gap_score, int16_t mismatch_score,
                                     int16_t match_score)
{
    if (!device_block || device_block_bytes <= 0 || (reinterpret_cast<uintptr_t>(device_block) & 255) != 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "device_block has to be a 256-byte aligned device allocation");
    return poa_batch_create_impl(out, device_id, stream, device_block_bytes, output_mask, cfg, gap_score, mismatch_score, match_score, device_block,
                                 device_block_bytes);
}

static int poa_batch_create_impl(gwb200_poa_batch** out, int32_t device_id, void* stream, int64_t max_gpu_mem, int8_t output_mask,
                                 const gwb200_poa_config* cfg, int16_t gap_score, int16_t mismatch_score, int16_t match_score, void* ext_block,
                                 int64_t ext_bytes)
{
    if (!out || !cfg)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (device_id < 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "Device ID has to be non-negative");
    if (max_gpu_mem < -1)
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_gpu_mem has to be either -1 (=all available GPU memory) or greater or equal than 0.");
    if (int rc = validate_config(*cfg))
        return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device_id)
    {
        cudaGetLastError();
        return set_error(GWB200_E_CUDA, "no usable CUDA device: this engine has no CPU fallback");
    }
    DeviceGuard guard(device_id);

    gwb200_poa_batch* b = new gwb200_poa_batch;
    b->device_id        = device_id;
    b->stream           = static_cast<cudaStream_t>(stream);
    b->output_mask      = output_mask;
    b->cfg              = *cfg;
    b->gap              = gap_score;
    b->mismatch         = mismatch_score;
    b->match            = match_score;
    b->score32          = use32bit_score(*cfg, gap_score, mismatch_score, match_score);
    b->size32           = use32bit_size(*cfg);
    b->score_bytes      = b->score32 ? 4 : 2;
    b->size_bytes       = b->size32 ? 4 : 2;
    b->msa              = (output_mask & GWB200_POA_OUTPUT_MSA) != 0;
    b->accurate         = spoa_accurate();
    b->tb_mode          = cfg->band_mode == GWB200_POA_STATIC_BAND_TRACEBACK || cfg->band_mode == GWB200_POA_ADAPTIVE_BAND_TRACEBACK;
    b->trace_bytes      = cfg->max_banded_pred_distance > INT8_MAX ? 2 : 1;
    if (b->tb_mode && cfg->max_banded_pred_distance < 1)
    {
        delete b;
        return set_error(GWB200_E_INVALID_ARGUMENT, "max_banded_pred_distance has to be positive in the traceback band modes");
    }
    b->bid              = gwb200_poa_batch::batches++;
    {
        const char* nwv = std::getenv("GWB200_POA_WARPS"); // development switch: warps per window (1, 2 or 4)
        b->nw_override  = nwv ? std::atoi(nwv) : 0;
    }

    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
    {
        delete b;
        return set_error(GWB200_E_CUDA, "cudaMemGetInfo failed");
    }
    int64_t avail = static_cast<int64_t>(static_cast<double>(free_b) * 0.95);
    if (max_gpu_mem >= 0)
        avail = std::min(avail, max_gpu_mem);
    const int64_t fixed_dev = 4096 * 2 + 256 * 64; // per-batch constant part of the device block
    if (ext_block)
        avail = ext_bytes - fixed_dev; // the caller's block is the budget: everything has to fit inside it
    const Sizes sz = compute_sizes(*cfg, b->score_bytes, b->size_bytes, b->msa, b->tb_mode, b->trace_bytes);
    if (avail < sz.dev_per_poa + (cfg->band_mode == GWB200_POA_FULL_BAND ? 0 : sz.dev_per_matrix) || sz.dev_per_poa + sz.dev_per_matrix <= 0)
    {
        std::string msg = "Requires at least " + std::to_string(sz.dev_per_poa + sz.dev_per_matrix) +
                          " bytes of device memory per CUDAPOA batch to process correctly.";
        delete b;
        return set_error(GWB200_E_RUNTIME, msg.c_str());
    }
    int64_t max_poas = avail / (sz.dev_per_poa + sz.dev_per_matrix);
    max_poas         = std::max<int64_t>(1, std::min<int64_t>(max_poas, 1 << 20));
    b->max_poas      = static_cast<int32_t>(max_poas);
    const int64_t n  = max_poas;
    const int64_t mn = cfg->max_nodes_per_graph;
    const int64_t mc = cfg->max_consensus_size;

    // ---- pinned host block
    {
        int64_t total_h = n * sz.host_per_poa + 4096 * 2 + 4096;
        if (cudaHostAlloc(reinterpret_cast<void**>(&b->h_block), total_h, cudaHostAllocDefault) != cudaSuccess)
        {
            cudaGetLastError();
            delete b;
            return set_error(GWB200_E_BAD_ALLOC, "pinned host allocation failed");
        }
        Carver hc{b->h_block};
        b->seq_capacity  = n * sz.seq_bytes_per_poa;
        b->h_sequences   = hc.take<uint8_t>(b->seq_capacity + 4096);
        b->h_weights     = hc.take<int8_t>(b->seq_capacity + 4096);
        b->h_seq_lengths = hc.take<int32_t>(n * cfg->max_sequences_per_poa);
        b->h_windows     = hc.take<WindowInfo>(n);
        b->h_consensus   = hc.take<uint8_t>(n * mc);
        b->h_coverage    = hc.take<uint16_t>(n * mc);
        b->h_cons_len    = hc.take<int32_t>(n);
        b->h_status      = hc.take<int32_t>(n);
        b->h_node_count  = hc.take<int32_t>(n);
        b->h_cells       = hc.take<unsigned long long>(n);
        if (b->msa)
            b->h_msa = hc.take<uint8_t>(n * cfg->max_sequences_per_poa * mc);
        if (hc.off > total_h + 256 * 16)
        {
            // host_per_poa has 64 B slack per POA; the 256 B carve alignment of 12 arrays is covered by the constant above
        }
        std::memset(b->h_sequences, 0, b->seq_capacity + 4096);
    }
    // ---- device block
    {
        const int64_t fixed   = fixed_dev;
        const int64_t total_d = n * (sz.dev_per_poa + sz.dev_per_matrix) + fixed;
        if (ext_block)
        {
            if (total_d > ext_bytes)
            {
                std::string msg = "Requires at least " + std::to_string(total_d) + " bytes of device memory per CUDAPOA batch to process correctly.";
                cudaFreeHost(b->h_block);
                delete b;
                return set_error(GWB200_E_RUNTIME, msg.c_str());
            }
            b->d_block     = static_cast<uint8_t*>(ext_block);
            b->owns_dblock = false;
        }
        else if (cudaMalloc(reinterpret_cast<void**>(&b->d_block), total_d) != cudaSuccess)
        {
            cudaGetLastError();
            cudaFreeHost(b->h_block);
            delete b;
            return set_error(GWB200_E_BAD_ALLOC, "Out of memory: device allocation for the POA batch failed");
        }
        Carver dc{b->d_block};
        DeviceParams& P = b->P;
        const int64_t S = b->size_bytes;
        P.sequences     = dc.take<uint8_t>(b->seq_capacity + 4096);
        P.weights       = dc.take<int8_t>(b->seq_capacity + 4096);
        P.seq_lengths   = dc.take<int32_t>(n * cfg->max_sequences_per_poa);
        P.windows       = dc.take<WindowInfo>(n);
        P.nodes         = dc.take<uint8_t>(n * mn);
        P.in_cnt        = dc.take<uint16_t>(n * mn);
        P.out_cnt       = dc.take<uint16_t>(n * mn);
        P.aln_cnt       = dc.take<uint16_t>(n * mn);
        P.node_cov      = dc.take<uint16_t>(n * mn);
        P.local_cnt     = dc.take<uint16_t>(n * mn);
        P.in_edges      = dc.take<uint8_t>(n * mn * kMaxEdges * S);
        P.out_edges     = dc.take<uint8_t>(n * mn * kMaxEdges * S);
        P.aligned       = dc.take<uint8_t>(n * mn * kMaxAligned * S);
        P.in_w          = dc.take<uint16_t>(n * mn * kMaxEdges);
        P.sorted        = dc.take<uint8_t>(n * mn * S);
        P.pos           = dc.take<uint8_t>(n * mn * S);
        P.aln_graph     = dc.take<uint8_t>(n * sz.aln_capacity * S);
        P.aln_read      = dc.take<uint8_t>(n * sz.aln_capacity * S);
        P.cons_scores   = dc.take<int32_t>(n * mn);
        P.cons_preds    = dc.take<uint8_t>(n * mn * S);
        P.consensus     = dc.take<uint8_t>(n * mc);
        P.coverage      = dc.take<uint16_t>(n * mc);
        P.consensus_len = dc.take<int32_t>(n);
        P.status        = dc.take<int32_t>(n);
        P.node_count    = dc.take<int32_t>(n);
        P.cells         = dc.take<unsigned long long>(n);
        if (b->msa)
        {
            P.seq_path       = dc.take<uint8_t>((b->seq_capacity + 4096) * S);
            P.msa_col        = dc.take<uint8_t>(n * mn * S);
            P.marks          = dc.take<uint8_t>(n * mn);
            P.check          = dc.take<uint8_t>(n * mn);
            P.stack          = dc.take<uint8_t>(n * sz.stack_capacity * S);
            P.stack_capacity = sz.stack_capacity;
            P.msa_out        = dc.take<uint8_t>(n * cfg->max_sequences_per_poa * mc);
        }
        else if (b->accurate)
        {
            P.marks          = dc.take<uint8_t>(n * mn);
            P.check          = dc.take<uint8_t>(n * mn);
            P.stack          = dc.take<uint8_t>(n * sz.stack_capacity * S);
            P.stack_capacity = sz.stack_capacity;
        }
        b->X.row_meta    = dc.take<int4>(n * (mn + 1));
        b->X.rd_capacity = align_up(std::max(cfg->max_sequence_size, 1), 4) + 8;
        b->X.rd_node     = dc.take<uint8_t>(n * static_cast<int64_t>(b->X.rd_capacity) * S);
        b->d_timers      = dc.take<unsigned long long>(n * 8);
        b->Y.work_counter = dc.take<int32_t>(64);
        b->X.timers      = nullptr;
        b->X.pool_bytes  = v2_pool_bytes(b);
        b->X.tb_scores   = nullptr;
        b->X.tb_trace    = nullptr;
        b->X.tb_height   = cfg->max_banded_pred_distance;
        b->X.tb_trace16  = b->trace_bytes == 2 ? 1 : 0;
        if (b->tb_mode)
            b->X.tb_scores = dc.take<uint8_t>(n * (static_cast<int64_t>(cfg->matrix_sequence_dimension) * cfg->max_banded_pred_distance * b->score_bytes + 256));
        // everything that is left is the score pool (allocate_block.hpp:227-239)
        dc.off                 = align_up64(dc.off, 256);
        P.scores               = b->d_block + dc.off;
        b->scorebuf_alloc_size = total_d - dc.off;
        if (b->tb_mode)
        {
            // in the traceback modes the pooled remainder is the trace matrix (allocate_block.hpp:236-239); cells the walk may
            // read without a prior write (row 0, unreachable boundary cells) are defined as zero
            b->X.tb_trace = P.scores;
            cudaMemsetAsync(P.scores, 0, b->scorebuf_alloc_size, b->stream);
        }
        P.max_nodes            = cfg->max_nodes_per_graph;
        P.matrix_seq_dim       = cfg->matrix_sequence_dimension;
        P.max_consensus        = cfg->max_consensus_size;
        P.max_seqs             = cfg->max_sequences_per_poa;
        P.band_width           = cfg->alignment_band_width;
        P.band_mode            = cfg->band_mode;
        P.gap                  = b->gap;
        P.mismatch             = b->mismatch;
        P.match                = b->match;
        P.msa                  = b->msa ? 1 : 0;
        P.accurate             = b->accurate ? 1 : 0;
        P.aln_capacity         = sz.aln_capacity;
    }
    cudaEventCreate(&b->ev0);
    cudaEventCreate(&b->ev1);
    {
        const char* k = std::getenv("GWB200_POA_KERNEL"); // development A/B switch: "v1" / "v2" select the earlier kernel generations
        b->use_v2     = !(k && std::string(k) == "v1");
        b->use_v3     = !(k && (std::string(k) == "v1" || std::string(k) == "v2"));
        {
            // the v2 kernel packs band starts in 15 bits (x4) and needs its staged read plus a few score rows in shared memory
            if (b->cfg.alignment_band_width > kMaxAdaptiveBW || b->cfg.max_sequence_size >= 65536 || b->X.pool_bytes > 200 * 1024)
                b->use_v2 = false; // first-generation kernel
            if (b->tb_mode)
                b->use_v2 = true; // only the v2 kernel family hosts the traceback-matrix alignment
            // the v3 kernel packs band starts in 16 bits and serves every score-matrix mode; the traceback-matrix modes stay on v2
            if (b->cfg.alignment_band_width > kMaxAdaptiveBW || b->cfg.max_sequence_size >= 65536 || b->tb_mode)
                b->use_v3 = false;
        }
        if (b->use_v3)
        {
            b->X.pool_bytes = v3_pool_bytes(b);
            const char* e   = std::getenv("GWB200_POA_BULK"); // development A/B switches
            b->Y.use_bulk   = (e && std::atoi(e) == 0) ? 0 : 1;
            e               = std::getenv("GWB200_POA_TB_TMA");
            b->Y.tb_tma     = (e && std::atoi(e) == 0) ? 0 : 1;
            e               = std::getenv("GWB200_POA_GROUP"); // development switch: chunks per straight-line group (4, 2, 1)
            // 32-bit scores: 8 cells per lane and chunk, groups of one chunk measured best (C3: 1493 vs 1485 windows/s with two)
            b->Y.max_group  = e ? std::atoi(e) : (b->score32 ? 1 : 2);
            e               = std::getenv("GWB200_POA_ROW_FENCE"); // development A/B switch: 1 = proxy fence in front of every general row
            b->Y.row_fence  = (e && std::atoi(e) != 0) ? 1 : 0;
            e               = std::getenv("GWB200_POA_WAVEFRONT");
            b->Y.wavefront  = (e && std::atoi(e) != 0) ? 1 : 0; // off by default until it beats dp_rows_v3 (development switch)
        }
    }
    gwb200_poa_batch_reset(b);
    *out = b;
    return 0;
}

void gwb200_poa_batch_destroy(gwb200_poa_batch* b)
{
    if (!b)
        return;
    DeviceGuard guard(b->device_id);
    cudaStreamSynchronize(b->stream);
    if (b->ev0)
        cudaEventDestroy(b->ev0);
    if (b->ev1)
        cudaEventDestroy(b->ev1);
    if (b->d_block && b->owns_dblock)
        cudaFree(b->d_block);
    if (b->h_block)
        cudaFreeHost(b->h_block);
    delete b;
}

int gwb200_poa_batch_reset(gwb200_poa_batch* b)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    b->poa_count              = 0;
    b->num_nucleotides_copied = 0;
    b->global_sequence_idx    = 0;
    b->next_scores_offset     = 0;
    b->avail_buf_mem          = b->scorebuf_alloc_size;
    b->weights_present        = false;
    b->launched               = false;
    b->results_on_host        = false;
    return 0;
}

// cudapoa_batch.cuh:103-151 (add_poa_group), :456-472 (add_poa), :475-542 (add_seq_to_poa), :545-570 (reserve_buf)
int gwb200_poa_batch_add_group(gwb200_poa_batch* b, int32_t n, const char* const* seqs, const int8_t* const* weights,
                               const int32_t* lengths, int32_t* per_seq_status, int32_t* n_per_seq)
{
    if (!b || (n > 0 && (!seqs || !lengths)))
        return set_error(GWB200_E_INVALID_ARGUMENT, "null argument");
    if (n_per_seq)
        *n_per_seq = 0;
    if (n <= 0)
        return set_error(GWB200_E_INVALID_ARGUMENT, "empty POA group"); // the reference dereferences end() here
    const gwb200_poa_config& c = b->cfg;
    int32_t max_seq_length     = 0;
    for (int32_t i = 0; i < n; i++)
        max_seq_length = std::max(max_seq_length, lengths[i]);
    // reserve_buf
    {
        const int64_t width = (c.band_mode != GWB200_POA_FULL_BAND) ? c.matrix_sequence_dimension : align_up(max_seq_length + 1 + kCPT, 4);
        const int64_t req   = width * static_cast<int64_t>(c.max_nodes_per_graph) * (b->tb_mode ? b->trace_bytes : b->score_bytes); // cudapoa_batch.cuh:551-553
        if (req > b->avail_buf_mem)
            return GWB200_POA_EXCEEDED_MAXIMUM_POAS;
        b->avail_buf_mem -= req;
    }
    // add_poa
    if (b->poa_count == b->max_poas)
        return GWB200_POA_EXCEEDED_MAXIMUM_POAS;
    WindowInfo wi{};
    wi.num_seqs       = 0;
    wi.seq_len_offset = b->global_sequence_idx;
    wi.seq_start      = b->num_nucleotides_copied;
    wi.scores_width   = 0;
    wi.scores_offset  = b->next_scores_offset;
    WindowInfo* w     = &b->h_windows[b->poa_count];
    *w                = wi;
    b->poa_count++;
    b->results_on_host = false;

    bool poa_empty = true;
    for (int32_t i = 0; i < n; i++)
    {
        const int32_t len = lengths[i];
        const int8_t* wt  = weights ? weights[i] : nullptr;
        int32_t st        = GWB200_POA_SUCCESS;
        if (len > c.max_sequence_size)
        {
            st = GWB200_POA_EXCEEDED_MAXIMUM_SEQUENCE_SIZE;
        }
        else
        {
            if (wt != nullptr)
            {
                bool all_zero = true;
                for (int32_t k = 0; k < len; k++)
                {
                    if (wt[k] < 0)
                        return set_error(GWB200_E_INVALID_ARGUMENT, "Base weights need to be non-negative");
                    if (wt[k] > 0)
                        all_zero = false;
                }
                if (all_zero)
                    st = GWB200_POA_ZERO_WEIGHTED_POA_SEQUENCE;
            }
            if (st == GWB200_POA_SUCCESS)
            {
                const int32_t sw = align_up(len + 1 + kCPT, 4);
                if (sw > w->scores_width)
                {
                    b->next_scores_offset += (sw - w->scores_width);
                    w->scores_width = sw;
                }
                if (w->num_seqs >= c.max_sequences_per_poa)
                {
                    st = GWB200_POA_EXCEEDED_MAXIMUM_SEQUENCES_PER_POA;
                }
                else
                {
                    w->num_seqs++;
                    if (b->deferred)
                        b->deferred->push_back({b->h_sequences + b->num_nucleotides_copied, seqs[i], len});
                    else
                        std::memcpy(b->h_sequences + b->num_nucleotides_copied, seqs[i], len);
                    if (wt == nullptr)
                    {
                        if (b->weights_present)
                            std::memset(b->h_weights + b->num_nucleotides_copied, 1, len);
                    }
                    else
                    {
                        if (!b->weights_present)
                        {
                            // first weighted read of this batch: the reads before it carry unit weights (cudapoa_batch.cuh:519-527)
                            std::memset(b->h_weights, 1, b->num_nucleotides_copied);
                            b->weights_present = true;
                        }
                        std::memcpy(b->h_weights + b->num_nucleotides_copied, wt, len);
                    }
                    b->h_seq_lengths[b->global_sequence_idx] = len;
                    b->num_nucleotides_copied += align_up(len, 4);
                    b->global_sequence_idx++;
                }
            }
        }
        if (st == GWB200_POA_SUCCESS)
            poa_empty = false;
        if (per_seq_status)
            per_seq_status[i] = st;
    }
    if (n_per_seq)
        *n_per_seq = n;
    if (poa_empty)
        return GWB200_POA_EMPTY_POA_GROUP;
    return GWB200_POA_SUCCESS;
}

int gwb200_poa_batch_add_groups_flat(gwb200_poa_batch* b, int32_t n_windows, const int32_t* win_nseq, const int32_t* seq_len,
                                     const char* seq_data, const int8_t* weights, int32_t* n_added)
{
    if (!b || !win_nseq || !seq_len || !seq_data)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null argument");
    if (n_added)
        *n_added = 0;
    std::vector<const char*> seqs;
    std::vector<const int8_t*> wts;
    int64_t off = 0;
    int32_t si  = 0;
    // the sequence bytes are copied into the pinned staging buffer by several threads after the (serial) bookkeeping
    std::vector<gwb200_poa_batch::CopyJob> jobs;
    struct Flush
    {
        gwb200_poa_batch* b;
        std::vector<gwb200_poa_batch::CopyJob>& jobs;
        ~Flush()
        {
            b->deferred = nullptr;
            int64_t total = 0;
            for (const auto& j : jobs)
                total += j.len;
            const int32_t nt = static_cast<int32_t>(std::min<int64_t>(std::min<int64_t>(16, std::max(1u, std::thread::hardware_concurrency())), total / (4 << 20) + 1));
            auto run = [&](size_t lo, size_t hi) {
                for (size_t k = lo; k < hi; k++)
                    std::memcpy(jobs[k].dst, jobs[k].src, jobs[k].len);
            };
            if (nt <= 1)
            {
                run(0, jobs.size());
                return;
            }
            std::vector<std::thread> th;
            const size_t per = (jobs.size() + nt - 1) / nt;
            for (int32_t t = 0; t < nt; t++)
            {
                const size_t lo = std::min(jobs.size(), per * t), hi = std::min(jobs.size(), per * (t + 1));
                if (lo < hi)
                    th.emplace_back(run, lo, hi);
            }
            for (auto& t : th)
                t.join();
        }
    } flush{b, jobs};
    b->deferred = &jobs;
    int last_soft = GWB200_POA_SUCCESS; // empty_poa_group of some window: reported, but the remaining windows are still added
    for (int32_t w = 0; w < n_windows; w++)
    {
        const int32_t ns = win_nseq[w];
        seqs.resize(ns);
        wts.resize(ns);
        int64_t o = off;
        for (int32_t s = 0; s < ns; s++)
        {
            seqs[s] = seq_data + o;
            wts[s]  = weights ? weights + o : nullptr;
            o += seq_len[si + s];
        }
        int rc = gwb200_poa_batch_add_group(b, ns, seqs.data(), weights ? wts.data() : nullptr, seq_len + si, nullptr, nullptr);
        // a group whose reads were all rejected was consumed all the same (it stays in the batch as an empty window,
        // cudapoa_batch.cuh:139-148): the caller's window <-> result mapping must count it
        if (rc != GWB200_POA_SUCCESS && rc != GWB200_POA_EMPTY_POA_GROUP)
            return rc;
        if (rc == GWB200_POA_EMPTY_POA_GROUP)
            last_soft = rc;
        off = o;
        si += ns;
        if (n_added)
            *n_added = w + 1;
    }
    return last_soft;
}

int32_t gwb200_poa_batch_total_poas(const gwb200_poa_batch* b) { return b ? b->poa_count : 0; }
int32_t gwb200_poa_batch_max_poas(const gwb200_poa_batch* b) { return b ? b->max_poas : 0; }
int32_t gwb200_poa_batch_id(const gwb200_poa_batch* b) { return b ? b->bid : -1; }
int32_t gwb200_poa_batch_score_bytes(const gwb200_poa_batch* b) { return b ? b->score_bytes : 0; }

int gwb200_poa_batch_upload(gwb200_poa_batch* b)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    if (b->poa_count == 0)
        return 0;
    DeviceGuard guard(b->device_id);
    // cudapoa_batch.cuh:171-178
    GWB200_CUDA_TRY(cudaMemcpyAsync(const_cast<uint8_t*>(b->P.sequences), b->h_sequences, b->num_nucleotides_copied, cudaMemcpyHostToDevice, b->stream));
    if (b->weights_present)
        GWB200_CUDA_TRY(cudaMemcpyAsync(const_cast<int8_t*>(b->P.weights), b->h_weights, b->num_nucleotides_copied, cudaMemcpyHostToDevice, b->stream));
    else
        GWB200_CUDA_TRY(cudaMemsetAsync(const_cast<int8_t*>(b->P.weights), 1, b->num_nucleotides_copied, b->stream)); // unit weights: no host traffic
    GWB200_CUDA_TRY(cudaMemcpyAsync(const_cast<WindowInfo*>(b->P.windows), b->h_windows, sizeof(WindowInfo) * b->poa_count, cudaMemcpyHostToDevice, b->stream));
    GWB200_CUDA_TRY(cudaMemcpyAsync(const_cast<int32_t*>(b->P.seq_lengths), b->h_seq_lengths, sizeof(int32_t) * b->global_sequence_idx, cudaMemcpyHostToDevice, b->stream));
    return 0;
}

int gwb200_poa_batch_launch(gwb200_poa_batch* b)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    if (b->poa_count == 0)
        return 0;
    DeviceGuard guard(b->device_id);
    b->P.n_windows = b->poa_count;
    GWB200_CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    batch_action(b, 0);
    GWB200_CUDA_TRY(cudaPeekAtLastError());
    GWB200_CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    b->launched        = true;
    b->results_on_host = false;
    return 0;
}

int gwb200_poa_batch_generate(gwb200_poa_batch* b)
{
    if (int rc = gwb200_poa_batch_upload(b))
        return rc;
    return gwb200_poa_batch_launch(b);
}

int gwb200_poa_batch_sync(gwb200_poa_batch* b)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    DeviceGuard guard(b->device_id);
    GWB200_CUDA_TRY(cudaStreamSynchronize(b->stream));
    return 0;
}

static int fetch_status(gwb200_poa_batch* b)
{
    const int64_t n = b->poa_count;
    GWB200_CUDA_TRY(cudaMemcpyAsync(b->h_status, b->P.status, 4 * n, cudaMemcpyDeviceToHost, b->stream));
    GWB200_CUDA_TRY(cudaMemcpyAsync(b->h_cons_len, b->P.consensus_len, 4 * n, cudaMemcpyDeviceToHost, b->stream));
    GWB200_CUDA_TRY(cudaMemcpyAsync(b->h_node_count, b->P.node_count, 4 * n, cudaMemcpyDeviceToHost, b->stream));
    GWB200_CUDA_TRY(cudaMemcpyAsync(b->h_cells, b->P.cells, 8 * n, cudaMemcpyDeviceToHost, b->stream));
    return 0;
}

int gwb200_poa_batch_get_consensus(gwb200_poa_batch* b, char* consensus, uint16_t* coverage, int32_t* lengths, int32_t* status)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    if (!(b->output_mask & GWB200_POA_OUTPUT_CONSENSUS))
        return GWB200_POA_OUTPUT_TYPE_UNAVAILABLE;
    DeviceGuard guard(b->device_id);
    const int64_t n  = b->poa_count;
    const int64_t mc = b->cfg.max_consensus_size;
    if (n == 0)
        return GWB200_POA_SUCCESS;
    if (!b->launched)
        return set_error(GWB200_E_RUNTIME, "get_consensus called before generate_poa");
    // unlike the reference (max_poas rows, cudapoa_batch.cuh:214-223) only the poa_count rows in use are copied
    GWB200_CUDA_TRY(cudaMemcpyAsync(b->h_consensus, b->P.consensus, n * mc, cudaMemcpyDeviceToHost, b->stream));
    GWB200_CUDA_TRY(cudaMemcpyAsync(b->h_coverage, b->P.coverage, n * mc * 2, cudaMemcpyDeviceToHost, b->stream));
    if (int rc = fetch_status(b))
        return rc;
    GWB200_CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->results_on_host = true;
    for (int64_t w = 0; w < n; w++)
    {
        const int32_t st  = b->h_status[w];
        const int32_t len = st == 0 ? b->h_cons_len[w] : 0;
        if (status)
            status[w] = st;
        if (lengths)
            lengths[w] = len;
        if (consensus)
        {
            std::memcpy(consensus + w * mc, b->h_consensus + w * mc, len);
            consensus[w * mc + len] = 0;
        }
        if (coverage)
            std::memcpy(coverage + w * mc, b->h_coverage + w * mc, static_cast<size_t>(len) * 2);
    }
    return GWB200_POA_SUCCESS;
}

int gwb200_poa_batch_get_msa(gwb200_poa_batch* b, char* msa, int32_t* num_rows, int32_t* status)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    if (!(b->output_mask & GWB200_POA_OUTPUT_MSA))
        return GWB200_POA_OUTPUT_TYPE_UNAVAILABLE;
    DeviceGuard guard(b->device_id);
    const int64_t n  = b->poa_count;
    const int64_t mc = b->cfg.max_consensus_size;
    const int64_t ms = b->cfg.max_sequences_per_poa;
    if (n == 0)
        return GWB200_POA_SUCCESS;
    if (!b->launched)
        return set_error(GWB200_E_RUNTIME, "get_msa called before generate_poa");
    GWB200_CUDA_TRY(cudaMemcpyAsync(b->h_msa, b->P.msa_out, n * ms * mc, cudaMemcpyDeviceToHost, b->stream));
    if (int rc = fetch_status(b))
        return rc;
    GWB200_CUDA_TRY(cudaStreamSynchronize(b->stream));
    b->results_on_host = true;
    for (int64_t w = 0; w < n; w++)
    {
        const int32_t st = b->h_status[w];
        if (status)
            status[w] = st;
        const int32_t rows = st == 0 ? b->h_windows[w].num_seqs : 0;
        if (num_rows)
            num_rows[w] = rows;
        if (msa)
        {
            for (int32_t r = 0; r < rows; r++)
            {
                const char* src = reinterpret_cast<const char*>(b->h_msa + (w * ms + r) * mc);
                char* dst       = msa + (w * ms + r) * mc;
                const size_t l  = strnlen(src, mc - 1);
                std::memcpy(dst, src, l);
                dst[l] = 0;
            }
        }
    }
    return GWB200_POA_SUCCESS;
}

int gwb200_poa_batch_get_graphs(gwb200_poa_batch* b, int32_t* node_counts, int32_t* edge_counts, int32_t* status, uint8_t* node_labels,
                                int32_t* edge_src, int32_t* edge_dst, int32_t* edge_weight)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    DeviceGuard guard(b->device_id);
    const int64_t n  = b->poa_count;
    const int64_t mn = b->cfg.max_nodes_per_graph;
    if (n == 0)
        return 0;
    if (!b->launched)
        return set_error(GWB200_E_RUNTIME, "get_graphs called before generate_poa");
    if (int rc = fetch_status(b))
        return rc;
    std::vector<uint16_t> in_cnt(n * mn);
    GWB200_CUDA_TRY(cudaMemcpyAsync(in_cnt.data(), b->P.in_cnt, n * mn * 2, cudaMemcpyDeviceToHost, b->stream));
    GWB200_CUDA_TRY(cudaStreamSynchronize(b->stream));
    const bool fill = edge_src != nullptr;
    std::vector<uint8_t> nodes, edges, wts;
    if (fill)
    {
        nodes.resize(n * mn);
        GWB200_CUDA_TRY(cudaMemcpy(nodes.data(), b->P.nodes, n * mn, cudaMemcpyDeviceToHost));
    }
    int64_t node_off = 0, edge_off = 0;
    std::vector<uint8_t> ebuf(mn * b->size_bytes);
    std::vector<uint16_t> wbuf(mn);
    for (int64_t w = 0; w < n; w++)
    {
        const int32_t st = b->h_status[w];
        if (status)
            status[w] = st;
        const int32_t nc = st == 0 ? b->h_node_count[w] : 0;
        int32_t ec       = 0;
        int32_t max_in   = 0;
        for (int32_t i = 0; i < nc; i++)
        {
            ec += in_cnt[w * mn + i];
            max_in = std::max<int32_t>(max_in, in_cnt[w * mn + i]);
        }
        if (node_counts)
            node_counts[w] = nc;
        if (edge_counts)
            edge_counts[w] = ec;
        if (fill && nc > 0)
        {
            std::memcpy(node_labels + node_off, nodes.data() + w * mn, nc);
            // edges are emitted per sink node in slot order (cudapoa_batch.cuh:376-390); gather slot by slot
            std::vector<int32_t> first(nc + 1, 0);
            for (int32_t i = 0; i < nc; i++)
                first[i + 1] = first[i] + in_cnt[w * mn + i];
            for (int32_t slot = 0; slot < max_in; slot++)
            {
                const uint8_t* dsrc = static_cast<const uint8_t*>(b->P.in_edges) + (w * mn * kMaxEdges + slot * mn) * b->size_bytes;
                GWB200_CUDA_TRY(cudaMemcpy(ebuf.data(), dsrc, nc * b->size_bytes, cudaMemcpyDeviceToHost));
                GWB200_CUDA_TRY(cudaMemcpy(wbuf.data(), b->P.in_w + w * mn * kMaxEdges + slot * mn, nc * 2, cudaMemcpyDeviceToHost));
                for (int32_t i = 0; i < nc; i++)
                {
                    if (slot < in_cnt[w * mn + i])
                    {
                        const int64_t k = edge_off + first[i] + slot;
                        edge_src[k]     = b->size_bytes == 2 ? reinterpret_cast<int16_t*>(ebuf.data())[i] : reinterpret_cast<int32_t*>(ebuf.data())[i];
                        edge_dst[k]     = i;
                        edge_weight[k]  = wbuf[i];
                    }
                }
            }
        }
        node_off += nc;
        edge_off += ec;
    }
    return 0;
}

int64_t gwb200_poa_batch_last_cells(gwb200_poa_batch* b)
{
    if (!b || !b->launched)
        return 0;
    DeviceGuard guard(b->device_id);
    if (!b->results_on_host)
    {
        if (fetch_status(b) != 0)
            return -1;
        cudaStreamSynchronize(b->stream);
    }
    int64_t tot = 0;
    for (int32_t w = 0; w < b->poa_count; w++)
        tot += static_cast<int64_t>(b->h_cells[w]);
    return tot;
}

float gwb200_poa_batch_last_kernel_ms(gwb200_poa_batch* b)
{
    if (!b || !b->launched)
        return 0.f;
    DeviceGuard guard(b->device_id);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, b->ev0, b->ev1) != cudaSuccess)
    {
        cudaGetLastError();
        return -1.f;
    }
    return ms;
}

int32_t gwb200_poa_batch_resident_windows(gwb200_poa_batch* b)
{
    if (!b)
        return 0;
    DeviceGuard guard(b->device_id);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, b->device_id);
    return sms * batch_action(b, 1);
}

int gwb200_poa_batch_enable_timers(gwb200_poa_batch* b, int32_t on)
{
    if (!b)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null batch");
    b->timers_on = on != 0;
    return 0;
}

int gwb200_poa_batch_get_timers(gwb200_poa_batch* b, uint64_t* out8)
{
    if (!b || !out8)
        return set_error(GWB200_E_INVALID_ARGUMENT, "null argument");
    DeviceGuard guard(b->device_id);
    for (int k = 0; k < 8; k++)
        out8[k] = 0;
    if (!b->launched || !b->timers_on || b->poa_count == 0)
        return 0;
    std::vector<unsigned long long> h(static_cast<size_t>(b->poa_count) * 8);
    GWB200_CUDA_TRY(cudaStreamSynchronize(b->stream));
    GWB200_CUDA_TRY(cudaMemcpy(h.data(), b->d_timers, h.size() * 8, cudaMemcpyDeviceToHost));
    for (int32_t w = 0; w < b->poa_count; w++)
        for (int k = 0; k < 8; k++)
            out8[k] += h[static_cast<size_t>(w) * 8 + k];
    return 0;
}

int gwb200_device_fdividef(int32_t n, const float* a, const float* b, float* out)
{
    if (n <= 0)
        return 0;
    float *da = nullptr, *db = nullptr, *dout = nullptr;
    GWB200_CUDA_TRY(cudaMalloc(&da, 4 * n));
    GWB200_CUDA_TRY(cudaMalloc(&db, 4 * n));
    GWB200_CUDA_TRY(cudaMalloc(&dout, 4 * n));
    GWB200_CUDA_TRY(cudaMemcpy(da, a, 4 * n, cudaMemcpyHostToDevice));
    GWB200_CUDA_TRY(cudaMemcpy(db, b, 4 * n, cudaMemcpyHostToDevice));
    fdividef_kernel<<<(n + 255) / 256, 256>>>(n, da, db, dout);
    count_launch();
    GWB200_CUDA_TRY(cudaMemcpy(out, dout, 4 * n, cudaMemcpyDeviceToHost));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return 0;
}

} // extern "C"
