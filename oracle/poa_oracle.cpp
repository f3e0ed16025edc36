// TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference cudapoa algorithms.
//
// Nothing in the product path (genomeworks_b200/, include/) may link, import or call this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as a checker.
//
// Parity status: PINNED. The restatement is checked (tests/test_oracle_poa.py) against the reference's own
// known-answer tests -- Test_CudapoaTopSort.cu:48-58, Test_CudapoaAddAlignment.cu:111-231, Test_CudapoaNW.cu:83-187
// and :444-508, Test_CudapoaGenerateConsensus.cu:83-165, the End2End golden assembly
// (Test_CudapoaBatchEnd2End.cu:39-91 with cudapoa/data/sample-windows.txt -> sample-golden-value.txt) -- and, on the
// GPU box, against the unmodified reference kernels themselves (oracle/_ref/libgwref.so, tests/test_gpu_*.py).
//
// Each function cites the reference file:line it restates (paths relative to /root/reference). The restatement is
// deliberately literal: the score matrix is one flat buffer per window that persists across reads (as the reference's
// per-window slice of scores_d does), including the reference's observable quirks (set_score(-1) offset,
// first_element_prev_score==0 for source nodes, get_scores' band_end = band_width-4, int truncation on store).
//
// The only thing a CPU cannot restate bit-for-bit is the single fast-math float division in
// cudapoa_nw_banded.cuh:207 (div.approx.ftz.f32 under -use_fast_math, cmake/CUDA.cmake:26). It is injected through
// oracle_set_fdiv(): default IEEE a/b; GPU tests install a callback that evaluates __fdividef on the device.

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace
{

constexpr int32_t MAXE           = 50;   // CUDAPOA_MAX_NODE_EDGES       cudapoa_structs.cuh:24
constexpr int32_t MAXA           = 50;   // CUDAPOA_MAX_NODE_ALIGNMENTS  cudapoa_structs.cuh:27
constexpr int32_t CPT            = 4;    // CUDAPOA_CELLS_PER_THREAD
constexpr int32_t MIN_BW         = 128;  // CUDAPOA_MIN_BAND_WIDTH
constexpr int32_t PAD            = 8;    // CUDAPOA_BANDED_MATRIX_RIGHT_PADDING
constexpr int32_t MAX_ADAPTIVE_BW = 1536; // CUDAPOA_MAX_ADAPTIVE_BAND_WIDTH
constexpr int32_t SHIFT_LEFT     = -10;
constexpr int32_t SHIFT_RIGHT    = -11;
constexpr int32_t NW_BACKTRACK_FAILED = -1;
constexpr int32_t NW_ADAPTIVE_STORAGE_FAILED = -2;
constexpr int32_t NW_TRACEBACK_BUFFER_FAILED = -3; // CUDAPOA_KERNEL_NW_TRACEBACK_BUFFER_FAILED cudapoa_structs.cuh:55

// cudapoa.hpp:34-49
enum Status : int32_t
{
    success = 0,
    exceeded_maximum_poas,
    exceeded_maximum_sequence_size,
    exceeded_maximum_sequences_per_poa,
    node_count_exceeded_maximum_graph_size,
    edge_count_exceeded_maximum_graph_size,
    exceeded_adaptive_banded_matrix_size,
    exceeded_maximum_predecessor_distance,
    loop_count_exceeded_upper_bound,
    output_type_unavailable,
    zero_weighted_poa_sequence,
    empty_poa_group,
    generic_error
};

enum BandMode : int32_t
{
    full_band = 0,
    static_band,
    adaptive_band,
    static_band_traceback,
    adaptive_band_traceback
};

typedef float (*fdiv_fn)(float, float);
float ieee_fdiv(float a, float b) { return a / b; }
fdiv_fn g_fdiv = ieee_fdiv;

inline int32_t align_up(int32_t v, int32_t b) { return (v + b - 1) & ~(b - 1); }

// The graph arrays exactly as the reference lays them out per window (cudapoa_kernels.cuh:127-198).
struct Graph
{
    int32_t max_nodes = 0;
    int32_t max_seqs  = 0;
    bool msa          = false;
    std::vector<uint8_t> nodes;
    std::vector<int32_t> in_edges;   // [node*MAXE+e]
    std::vector<uint16_t> in_cnt;
    std::vector<int32_t> out_edges;  // [node*MAXE+e]
    std::vector<uint16_t> out_cnt;
    std::vector<uint16_t> in_w;      // [node*MAXE+e]
    std::vector<int32_t> sorted;     // rank -> node
    std::vector<int32_t> pos;        // node -> rank
    std::vector<int32_t> aligned;    // [node*MAXA+a]
    std::vector<uint16_t> aln_cnt;
    std::vector<uint16_t> coverage;
    std::vector<uint16_t> local_in_cnt;
    // MSA only
    std::vector<int32_t> seq_begin;       // [max_seqs]
    std::vector<uint16_t> out_cov;        // [(node*MAXE+e)*max_seqs + k]
    std::vector<uint16_t> out_cov_cnt;    // [node*MAXE+e]
    int32_t node_count = 0;

    void init(int32_t mn, int32_t ms, bool m)
    {
        max_nodes = mn;
        max_seqs  = ms;
        msa       = m;
        nodes.assign(mn, 0);
        in_edges.assign(static_cast<size_t>(mn) * MAXE, 0);
        in_cnt.assign(mn, 0);
        out_edges.assign(static_cast<size_t>(mn) * MAXE, 0);
        out_cnt.assign(mn, 0);
        in_w.assign(static_cast<size_t>(mn) * MAXE, 0);
        sorted.assign(mn, 0);
        pos.assign(mn, 0);
        aligned.assign(static_cast<size_t>(mn) * MAXA, 0);
        aln_cnt.assign(mn, 0);
        coverage.assign(mn, 0);
        local_in_cnt.assign(mn, 0);
        seq_begin.assign(ms > 0 ? ms : 1, 0);
        if (m)
        {
            out_cov.assign(static_cast<size_t>(mn) * MAXE * ms, 0);
            out_cov_cnt.assign(static_cast<size_t>(mn) * MAXE, 0);
        }
        node_count = 0;
    }
};

// ---------------------------------------------------------------------------------------------------------
// topologicalSortDeviceUtil -- cudapoa_topsort.cuh:45-97
void topsort(Graph& g, int32_t node_count)
{
    int32_t p = 0;
    for (int32_t n = 0; n < node_count; n++)
    {
        g.local_in_cnt[n] = g.in_cnt[n];
        if (g.local_in_cnt[n] == 0)
        {
            g.pos[n]      = p;
            g.sorted[p++] = n;
        }
    }
    for (int32_t n = 0; n < p; n++)
    {
        int32_t node = g.sorted[n];
        for (int32_t e = 0; e < g.out_cnt[node]; e++)
        {
            int32_t out_node = g.out_edges[node * MAXE + e];
            uint16_t c       = g.local_in_cnt[out_node];
            if (--c == 0)
            {
                g.pos[out_node] = p;
                g.sorted[p++]   = out_node;
            }
            g.local_in_cnt[out_node] = c;
        }
    }
}

// raconTopologicalSortDeviceUtil -- cudapoa_topsort.cuh:103-197
void racon_topsort(Graph& g, int32_t node_count)
{
    std::vector<uint8_t> marks(g.max_nodes, 0);
    std::vector<uint8_t> check(g.max_nodes, 1);
    std::vector<int32_t> stack(std::max(g.max_nodes, 1) * 4 + 16, 0); // generous; reference uses max_nodes entries
    int32_t node_idx   = -1;
    int32_t sorted_idx = 0;
    for (int32_t i = 0; i < node_count; i++)
    {
        if (marks[i] != 0)
            continue;
        node_idx++;
        stack[node_idx] = i;
        while (node_idx != -1)
        {
            int32_t node_id = stack[node_idx];
            bool valid      = true;
            if (marks[node_id] != 2)
            {
                for (int32_t e = 0; e < g.in_cnt[node_id]; e++)
                {
                    int32_t b = g.in_edges[node_id * MAXE + e];
                    if (marks[b] != 2)
                    {
                        node_idx++;
                        if (node_idx >= static_cast<int32_t>(stack.size()))
                            stack.resize(stack.size() * 2);
                        stack[node_idx] = b;
                        valid           = false;
                    }
                }
                if (check[node_id])
                {
                    for (int32_t a = 0; a < g.aln_cnt[node_id]; a++)
                    {
                        int32_t aid = g.aligned[node_id * MAXA + a];
                        if (marks[aid] != 2)
                        {
                            node_idx++;
                            if (node_idx >= static_cast<int32_t>(stack.size()))
                                stack.resize(stack.size() * 2);
                            stack[node_idx] = aid;
                            check[aid]      = 0;
                            valid           = false;
                        }
                    }
                }
                if (valid)
                {
                    marks[node_id] = 2;
                    if (check[node_id])
                    {
                        g.sorted[sorted_idx] = node_id;
                        g.pos[node_id]       = sorted_idx;
                        sorted_idx++;
                        for (int32_t a = 0; a < g.aln_cnt[node_id]; a++)
                        {
                            int32_t aid          = g.aligned[node_id * MAXA + a];
                            g.sorted[sorted_idx] = aid;
                            g.pos[aid]           = sorted_idx;
                            sorted_idx++;
                        }
                    }
                }
                else
                {
                    marks[node_id] = 1;
                }
            }
            if (valid)
                node_idx--;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// addAlignmentToGraph -- cudapoa_add_alignment.cuh:65-285
uint8_t add_alignment(Graph& g, int32_t& new_node_count, int32_t node_count, int32_t alignment_length,
                      const int32_t* alignment_graph, const uint8_t* read, const int32_t* alignment_read,
                      const int8_t* base_weights, int32_t s)
{
    int32_t head_node_id = -1;
    int32_t curr_node_id = -1;
    uint16_t prev_weight = 0;
    const uint32_t max_limit = static_cast<uint32_t>(g.max_nodes);
    for (int32_t p = alignment_length - 1; p >= 0; p--)
    {
        int32_t read_pos = alignment_read[p];
        if (read_pos != -1)
        {
            int8_t NODE_WEIGHT    = base_weights[read_pos];
            uint8_t read_base     = read[read_pos];
            int32_t graph_node_id = alignment_graph[p];
            if (graph_node_id == -1)
            {
                curr_node_id = node_count++;
                if (static_cast<uint32_t>(node_count) >= max_limit)
                    return static_cast<uint8_t>(node_count_exceeded_maximum_graph_size);
                g.nodes[curr_node_id]    = read_base;
                g.out_cnt[curr_node_id]  = 0;
                g.in_cnt[curr_node_id]   = 0;
                g.aln_cnt[curr_node_id]  = 0;
                g.coverage[curr_node_id] = 0;
            }
            else
            {
                uint8_t graph_base = g.nodes[graph_node_id];
                if (graph_base == read_base)
                {
                    curr_node_id = graph_node_id;
                }
                else
                {
                    uint16_t num_aligned   = g.aln_cnt[graph_node_id];
                    int32_t aligned_node_id = -1;
                    for (int32_t n = 0; n < num_aligned; n++)
                    {
                        int32_t aid = g.aligned[graph_node_id * MAXA + n];
                        if (g.nodes[aid] == read_base)
                        {
                            aligned_node_id = aid;
                            break;
                        }
                    }
                    if (aligned_node_id != -1)
                    {
                        curr_node_id = aligned_node_id;
                    }
                    else
                    {
                        curr_node_id = node_count++;
                        if (static_cast<uint32_t>(node_count) >= max_limit)
                            return static_cast<uint8_t>(node_count_exceeded_maximum_graph_size);
                        g.nodes[curr_node_id]    = read_base;
                        g.out_cnt[curr_node_id]  = 0;
                        g.in_cnt[curr_node_id]   = 0;
                        g.aln_cnt[curr_node_id]  = 0;
                        g.coverage[curr_node_id] = 0;
                        int32_t new_node_alignments = 0;
                        for (int32_t n = 0; n < num_aligned; n++)
                        {
                            int32_t aid        = g.aligned[graph_node_id * MAXA + n];
                            uint16_t aid_count = g.aln_cnt[aid];
                            g.aligned[aid * MAXA + aid_count] = curr_node_id;
                            g.aln_cnt[aid]                    = aid_count + 1;
                            g.aligned[curr_node_id * MAXA + new_node_alignments] = aid;
                            new_node_alignments++;
                        }
                        g.aligned[graph_node_id * MAXA + num_aligned] = curr_node_id;
                        g.aln_cnt[graph_node_id]                      = num_aligned + 1;
                        g.aligned[curr_node_id * MAXA + new_node_alignments] = graph_node_id;
                        new_node_alignments++;
                        g.aln_cnt[curr_node_id] = static_cast<uint16_t>(new_node_alignments);
                    }
                }
            }
            if (g.msa && read_pos == 0)
                g.seq_begin[s] = curr_node_id;

            if (head_node_id != -1)
            {
                bool edge_exists  = false;
                uint16_t in_count = g.in_cnt[curr_node_id];
                for (int32_t e = 0; e < in_count; e++)
                {
                    if (g.in_edges[curr_node_id * MAXE + e] == head_node_id)
                    {
                        edge_exists = true;
                        g.in_w[curr_node_id * MAXE + e] = static_cast<uint16_t>(g.in_w[curr_node_id * MAXE + e] + (prev_weight + NODE_WEIGHT));
                    }
                }
                if (!edge_exists)
                {
                    g.in_edges[curr_node_id * MAXE + in_count] = head_node_id;
                    g.in_w[curr_node_id * MAXE + in_count]     = static_cast<uint16_t>(prev_weight + NODE_WEIGHT);
                    g.in_cnt[curr_node_id]                     = in_count + 1;
                    uint16_t out_count                         = g.out_cnt[head_node_id];
                    g.out_edges[head_node_id * MAXE + out_count] = curr_node_id;
                    if (g.msa)
                    {
                        g.out_cov_cnt[head_node_id * MAXE + out_count] = 1;
                        g.out_cov[static_cast<size_t>(head_node_id * MAXE + out_count) * g.max_seqs] = static_cast<uint16_t>(s);
                    }
                    g.out_cnt[head_node_id] = out_count + 1;
                    if (out_count + 1 >= MAXE || in_count + 1 >= MAXE)
                        return static_cast<uint8_t>(edge_count_exceeded_maximum_graph_size);
                }
                else if (g.msa)
                {
                    uint16_t out_count = g.out_cnt[head_node_id];
                    for (int32_t e = 0; e < out_count; e++)
                    {
                        if (g.out_edges[head_node_id * MAXE + e] == curr_node_id)
                        {
                            uint16_t c = g.out_cov_cnt[head_node_id * MAXE + e];
                            g.out_cov[static_cast<size_t>(head_node_id * MAXE + e) * g.max_seqs + c] = static_cast<uint16_t>(s);
                            g.out_cov_cnt[head_node_id * MAXE + e] = c + 1;
                            break;
                        }
                    }
                }
            }
            head_node_id = curr_node_id;
            g.coverage[head_node_id]++;
            prev_weight = static_cast<uint16_t>(NODE_WEIGHT);
        }
    }
    new_node_count = node_count;
    return static_cast<uint8_t>(success);
}

// ---------------------------------------------------------------------------------------------------------
// Banded NW -- cudapoa_nw_banded.cuh
template <typename ScoreT>
struct ScoreBuf
{
    std::vector<ScoreT> data;
    int64_t oob_reads  = 0;
    int64_t oob_writes = 0;
    ScoreT rd(int64_t i)
    {
        if (i < 0 || i >= static_cast<int64_t>(data.size()))
        {
            oob_reads++;
            return 0;
        }
        return data[i];
    }
    void wr(int64_t i, int32_t v)
    {
        if (i < 0 || i >= static_cast<int64_t>(data.size()))
        {
            oob_writes++;
            return;
        }
        data[i] = static_cast<ScoreT>(v); // truncation on store, as the device does
    }
};

// get_band_start_for_row -- cudapoa_nw_banded.cuh:67-78
inline int32_t band_start_for_row(int32_t row, float gradient, int32_t bw, int32_t band_shift, int32_t max_column)
{
    float prod             = static_cast<float>(row) * gradient; // single-precision multiply, as on the device
    int32_t diagonal_index = static_cast<int32_t>(prod);
    int32_t start_pos      = std::max(0, diagonal_index - band_shift);
    if (max_column < start_pos + bw)
        start_pos = std::max(0, max_column - bw + CPT);
    start_pos = start_pos - (start_pos % CPT);
    return start_pos;
}

template <typename ScoreT>
struct BandCtx
{
    ScoreBuf<ScoreT>* sb;
    int32_t bw, band_shift, max_column;
    float gradient;
    ScoreT min_score;
    int64_t stride() const { return static_cast<int64_t>(bw + PAD); }
    // get_score_ptr -- :35-43
    int64_t ptr(int32_t row, int32_t column, int32_t band_start) const
    {
        int32_t c = (column == -1) ? 0 : column - band_start;
        return static_cast<int64_t>(c) + static_cast<int64_t>(row) * stride();
    }
    // set_score -- :45-65 (note the column == -1 quirk: offset band_start instead of 0)
    void set_score(int32_t row, int32_t column, int32_t value, int32_t band_start)
    {
        int32_t c = (column == -1) ? band_start : column - band_start;
        sb->wr(static_cast<int64_t>(c) + static_cast<int64_t>(row) * stride(), value);
    }
    // get_score -- :80-102
    ScoreT get_score(int32_t row, int32_t column)
    {
        int32_t bs = band_start_for_row(row, gradient, bw, band_shift, max_column);
        int32_t be = std::min(bs + bw, max_column);
        if ((column > be || column < bs) && column != -1)
            return min_score;
        return sb->rd(ptr(row, column, bs));
    }
    // get_scores -- :104-156
    void get_scores(int32_t row, int32_t column, int32_t gap, const int32_t prof[4], ScoreT out[4])
    {
        int32_t bs = band_start_for_row(row, gradient, bw, band_shift, max_column);
        int32_t be = std::min(bs + bw - CPT, max_column);
        if ((column > be || column < bs) && column != -1)
        {
            out[0] = out[1] = out[2] = out[3] = min_score;
            return;
        }
        int64_t p = ptr(row, column, bs);
        ScoreT a0 = sb->rd(p), a1 = sb->rd(p + 1), a2 = sb->rd(p + 2), a3 = sb->rd(p + 3), n0 = sb->rd(p + 4);
        out[0] = static_cast<ScoreT>(std::max(a0 + prof[0], a1 + gap));
        out[1] = static_cast<ScoreT>(std::max(a1 + prof[1], a2 + gap));
        out[2] = static_cast<ScoreT>(std::max(a2 + prof[2], a3 + gap));
        out[3] = static_cast<ScoreT>(std::max(a3 + prof[3], n0 + gap));
    }
    // initialize_band -- :158-175
    void initialize_band(int32_t row, int32_t band_start)
    {
        int32_t band_end = band_start + bw;
        int32_t bs       = std::max(1, band_start);
        set_score(row, bs, min_score, bs);
        for (int32_t lane = 0; lane < PAD; lane++)
            set_score(row, lane + band_end, min_score, bs);
    }
};

// needlemanWunschBanded -- cudapoa_nw_banded.cuh:177-557
template <typename ScoreT>
int32_t nw_banded(Graph& g, int32_t graph_count, const uint8_t* read, int32_t read_length, ScoreBuf<ScoreT>& sb,
                  float max_buffer_size, int32_t* alignment_graph, int32_t* alignment_read, int32_t band_width,
                  int32_t gap, int32_t mismatch, int32_t match, int32_t rerun, bool adaptive, int64_t* cells_out)
{
    const ScoreT min_score = std::numeric_limits<ScoreT>::min() / 2;
    float gradient         = g_fdiv(static_cast<float>(read_length + 1), static_cast<float>(graph_count + 1));
    int32_t max_column     = read_length + 1;
    if (adaptive)
    {
        if (gradient > 1.1)
            band_width = std::max(band_width, align_up(static_cast<int32_t>(max_column * 0.08 * gradient), MIN_BW));
        if (gradient < 0.8)
            band_width = std::max(band_width, align_up(static_cast<int32_t>(max_column * 0.1 / gradient), MIN_BW));
        band_width = std::min(band_width, MAX_ADAPTIVE_BW);
        if (band_width == MAX_ADAPTIVE_BW && rerun != 0)
            return rerun;
    }
    int32_t band_shift = band_width / 2;
    if (adaptive)
    {
        if (rerun == SHIFT_LEFT && band_width <= MAX_ADAPTIVE_BW / 2)
        {
            band_width *= 2;
            band_shift = static_cast<int32_t>(band_shift * 2.5);
        }
        if (rerun == SHIFT_RIGHT && band_width <= MAX_ADAPTIVE_BW / 2)
        {
            band_width *= 2;
            band_shift = static_cast<int32_t>(band_shift * 1.5);
        }
        float required = static_cast<float>(graph_count) * static_cast<float>(band_width + PAD);
        if (required > max_buffer_size)
            return NW_ADAPTIVE_STORAGE_FAILED;
    }
    if (cells_out)
        *cells_out += static_cast<int64_t>(graph_count) * band_width;

    BandCtx<ScoreT> c{&sb, band_width, band_shift, max_column, gradient, min_score};

    for (int32_t j = 0; j < band_width + PAD; j++)
        sb.wr(j, j * gap);

    for (int32_t graph_pos = 0; graph_pos < graph_count; graph_pos++)
    {
        int32_t node_id    = g.sorted[graph_pos];
        int32_t score_gIdx = graph_pos + 1;
        int32_t band_start = band_start_for_row(score_gIdx, gradient, band_width, band_shift, max_column);
        c.initialize_band(score_gIdx, band_start);

        int32_t first_element_prev_score = 0;
        uint16_t pred_count              = g.in_cnt[node_id];
        int32_t pred_idx                 = 0;
        if (pred_count == 0)
        {
            c.set_score(score_gIdx, -1, gap, band_start);
        }
        else
        {
            pred_idx = g.pos[g.in_edges[node_id * MAXE]] + 1;
            if (band_start > CPT && pred_count == 1)
            {
                first_element_prev_score = min_score + gap;
            }
            else
            {
                int32_t penalty = std::max<int32_t>(min_score, c.get_score(pred_idx, -1));
                for (int32_t p = 0; p < pred_count; p++)
                {
                    int32_t pi = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                    penalty    = std::max<int32_t>(penalty, c.get_score(pi, -1));
                }
                first_element_prev_score = penalty + gap;
            }
            c.set_score(score_gIdx, -1, first_element_prev_score, band_start);
        }
        uint8_t graph_base = g.nodes[node_id];

        for (int32_t chunk = band_start; chunk < band_start + band_width; chunk += 32 * CPT)
        {
            ScoreT cell[32 * CPT];
            for (int32_t lane = 0; lane < 32; lane++)
            {
                int32_t read_pos = chunk + lane * CPT;
                int32_t prof[4];
                for (int32_t k = 0; k < 4; k++)
                {
                    // bytes past the end of the read are whatever follows in the sequence buffer on the device; they
                    // only reach cells with column > read_length, which never influence columns <= read_length.
                    uint8_t rb = (read_pos + k < read_length) ? read[read_pos + k] : 0;
                    prof[k]    = (graph_base == rb) ? match : mismatch;
                }
                ScoreT s[4];
                c.get_scores(pred_idx, read_pos, gap, prof, s);
                for (int32_t p = 1; p < pred_count; p++)
                {
                    int32_t pi = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                    ScoreT t[4];
                    c.get_scores(pi, read_pos, gap, prof, t);
                    for (int32_t k = 0; k < 4; k++)
                        s[k] = std::max(s[k], t[k]);
                }
                for (int32_t k = 0; k < 4; k++)
                    cell[lane * 4 + k] = s[k];
            }
            // horizontal relaxation (:362-390) reaches the unique fixpoint of s[c] = max(s[c], s[c-1] + gap)
            int32_t last = first_element_prev_score;
            for (int32_t k = 0; k < 32 * CPT; k++)
            {
                cell[k] = static_cast<ScoreT>(std::max<int32_t>(last + gap, cell[k]));
                last    = cell[k];
            }
            first_element_prev_score = cell[32 * CPT - 1];
            for (int32_t k = 0; k < 32 * CPT; k++)
            {
                int64_t idx = static_cast<int64_t>(chunk + k + 1 - band_start) + static_cast<int64_t>(score_gIdx) * c.stride();
                sb.wr(idx, cell[k]);
            }
        }
    }

    int32_t aligned_nodes = 0;
    {
        int32_t i      = 0;
        int32_t j      = read_length;
        int32_t mscore = min_score;
        for (int32_t idx = 1; idx <= graph_count; idx++)
        {
            if (g.out_cnt[g.sorted[idx - 1]] == 0)
            {
                int32_t s = c.get_score(idx, j);
                if (mscore < s)
                {
                    mscore = s;
                    i      = idx;
                }
            }
        }
        int32_t prev_i = 0, prev_j = 0;
        int32_t next_node_id = i > 0 ? g.sorted[i - 1] : 0;
        int32_t loop_count   = 0;
        while (!(i == 0 && j == 0) && loop_count < (read_length + graph_count + 2))
        {
            loop_count++;
            int32_t scores_ij = c.get_score(i, j);
            bool pred_found   = false;
            if (i != 0 && j != 0)
            {
                if (adaptive)
                {
                    if (rerun == 0 && band_width < MAX_ADAPTIVE_BW)
                    {
                        int32_t threshold = std::max(1, max_column / 1024);
                        if (j > threshold && j < max_column - threshold)
                        {
                            int32_t bs = band_start_for_row(i, gradient, band_width, band_shift, max_column);
                            if (j <= bs + threshold)
                            {
                                aligned_nodes = SHIFT_LEFT;
                                break;
                            }
                            if (j >= (bs + band_width - threshold))
                            {
                                aligned_nodes = SHIFT_RIGHT;
                                break;
                            }
                        }
                    }
                }
                int32_t node_id    = next_node_id;
                int32_t match_cost = (g.nodes[node_id] == read[j - 1]) ? match : mismatch;
                uint16_t pc        = g.in_cnt[node_id];
                int32_t pred_i     = (pc == 0) ? 0 : (g.pos[g.in_edges[node_id * MAXE]] + 1);
                if (scores_ij == (c.get_score(pred_i, j - 1) + match_cost))
                {
                    prev_i     = pred_i;
                    prev_j     = j - 1;
                    pred_found = true;
                }
                if (!pred_found)
                {
                    for (int32_t p = 1; p < pc; p++)
                    {
                        pred_i = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                        if (scores_ij == (c.get_score(pred_i, j - 1) + match_cost))
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
                int32_t node_id = g.sorted[i - 1];
                uint16_t pc     = g.in_cnt[node_id];
                int32_t pred_i  = (pc == 0) ? 0 : g.pos[g.in_edges[node_id * MAXE]] + 1;
                if (scores_ij == c.get_score(pred_i, j) + gap)
                {
                    prev_i     = pred_i;
                    prev_j     = j;
                    pred_found = true;
                }
                if (!pred_found)
                {
                    for (int32_t p = 1; p < pc; p++)
                    {
                        pred_i = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                        if (scores_ij == c.get_score(pred_i, j) + gap)
                        {
                            prev_i     = pred_i;
                            prev_j     = j;
                            pred_found = true;
                            break;
                        }
                    }
                }
            }
            if (!pred_found && scores_ij == c.get_score(i, j - 1) + gap)
            {
                prev_i     = i;
                prev_j     = j - 1;
                pred_found = true;
            }
            // graph[prev_i - 1] with prev_i == 0 reads one element before the window's sorted_poa slice on the device;
            // the value is only used when the next step still has i != 0, so 0 is a faithful stand-in.
            next_node_id = prev_i > 0 ? g.sorted[prev_i - 1] : 0;

            alignment_graph[aligned_nodes] = (i == prev_i) ? -1 : g.sorted[i - 1];
            alignment_read[aligned_nodes]  = (j == prev_j) ? -1 : j - 1;
            aligned_nodes++;
            i = prev_i;
            j = prev_j;
        }
        if (loop_count >= (read_length + graph_count + 2))
            aligned_nodes = NW_BACKTRACK_FAILED;
    }
    return aligned_nodes;
}

// needlemanWunsch (full band) -- cudapoa_nw.cuh:149-454. Columns beyond read_length inside the last 4-cell group are
// ---------------------------------------------------------------------------------------------------------
// Banded NW with traceback matrix -- cudapoa_nw_tb_banded.cuh. The score matrix is kept only for the last
// `score_matrix_height` rows (row % height), the move of every cell goes to a full-height trace matrix:
// 0 = horizontal, +d = diagonal to the row d above, -d = vertical to the row d above.
template <typename ScoreT, typename TraceT>
struct TbCtx
{
    ScoreBuf<ScoreT>* sb;          // [score_matrix_height][bw + PAD]
    std::vector<TraceT>* tb;       // [max_nodes][bw + PAD], persistent per window like the device buffer (zero-initialised)
    int32_t bw, band_shift, max_column, height;
    float gradient;
    ScoreT min_score;
    int64_t stride() const { return static_cast<int64_t>(bw + PAD); }
    // set_score_tb -- :46-67 (column == -1 quirk: offset band_start)
    void set_score(int32_t row, int32_t column, int32_t value, int32_t band_start)
    {
        int32_t c = (column == -1) ? band_start : column - band_start;
        row       = row % height;
        sb->wr(static_cast<int64_t>(c) + static_cast<int64_t>(row) * stride(), value);
    }
    // get_score_tb -- :111-138
    ScoreT get_score(int32_t row, int32_t column)
    {
        int32_t bs = band_start_for_row(row, gradient, bw, band_shift, max_column);
        int32_t be = std::min(bs + bw, max_column);
        if ((column > be || column < bs) && column != -1)
            return min_score;
        int32_t c = (column == -1) ? 0 : column - bs;
        return sb->rd(static_cast<int64_t>(c) + static_cast<int64_t>(row % height) * stride());
    }
    // initialize_band_tb -- :83-101 (band_end from the unclamped band_start, offsets from max(1, band_start))
    void initialize_band(int32_t row, int32_t band_start)
    {
        int32_t band_end = band_start + bw;
        int32_t bs       = std::max(1, band_start);
        set_score(row, bs, min_score, bs);
        for (int32_t lane = 0; lane < PAD; lane++)
            set_score(row, lane + band_end, min_score, bs);
    }
    void tb_wr(int64_t i, int32_t v)
    {
        if (i >= 0 && i < static_cast<int64_t>(tb->size()))
            (*tb)[i] = static_cast<TraceT>(v); // truncation on store
    }
    int32_t tb_rd(int64_t i) const { return (i >= 0 && i < static_cast<int64_t>(tb->size())) ? static_cast<int32_t>((*tb)[i]) : 0; }
    // get_scores_tb -- :140-262: one predecessor's contribution to four cells and their moves
    void get_scores(int32_t pred_node, int32_t current_node, int32_t column, int32_t gap, const int32_t prof[4], ScoreT score[4], TraceT trace[4])
    {
        int32_t bs = band_start_for_row(pred_node, gradient, bw, band_shift, max_column);
        int32_t be = std::min(bs + bw - CPT, max_column);
        if ((column > be || column < bs) && column != -1)
            return;
        int32_t c  = (column == -1) ? 0 : column - bs;
        int64_t p  = static_cast<int64_t>(c) + static_cast<int64_t>(pred_node % height) * stride();
        ScoreT a[5] = {sb->rd(p), sb->rd(p + 1), sb->rd(p + 2), sb->rd(p + 3), sb->rd(p + 4)};
        for (int32_t k = 0; k < 4; k++)
        {
            const int32_t diag = a[k] + prof[k];
            const int32_t vert = a[k + 1] + gap;
            if (diag >= vert)
            {
                if (diag > score[k])
                {
                    score[k] = static_cast<ScoreT>(diag);
                    trace[k] = static_cast<TraceT>(current_node - pred_node);
                }
            }
            else
            {
                if (vert > score[k])
                {
                    score[k] = static_cast<ScoreT>(vert);
                    trace[k] = static_cast<TraceT>(-(current_node - pred_node));
                }
            }
        }
    }
};

// needlemanWunschBandedTraceback -- cudapoa_nw_tb_banded.cuh:264-643
template <typename ScoreT, typename TraceT>
int32_t nw_banded_tb(Graph& g, int32_t graph_count, const uint8_t* read, int32_t read_length, ScoreBuf<ScoreT>& sb, std::vector<TraceT>& tbuf,
                     float max_buffer_size, int32_t* alignment_graph, int32_t* alignment_read, int32_t band_width, int32_t score_matrix_height,
                     int32_t gap, int32_t mismatch, int32_t match, int32_t rerun, bool adaptive, int64_t* cells_out)
{
    const ScoreT min_score = std::numeric_limits<ScoreT>::min() / 2;
    float gradient         = g_fdiv(static_cast<float>(read_length + 1), static_cast<float>(graph_count + 1));
    int32_t max_column     = read_length + 1;
    int32_t band_shift     = band_width / 2;
    if (adaptive)
    {
        if (rerun == SHIFT_LEFT && band_width <= MAX_ADAPTIVE_BW / 2)
        {
            band_width *= 2;
            band_shift = static_cast<int32_t>(band_shift * 2.5);
        }
        if (rerun == SHIFT_RIGHT && band_width <= MAX_ADAPTIVE_BW / 2)
        {
            band_width *= 2;
            band_shift = static_cast<int32_t>(band_shift * 1.5);
        }
        float required = static_cast<float>(graph_count) * static_cast<float>(band_width + PAD);
        if (required > max_buffer_size)
            return NW_ADAPTIVE_STORAGE_FAILED;
    }
    if (cells_out)
        *cells_out += static_cast<int64_t>(graph_count) * band_width;

    TbCtx<ScoreT, TraceT> c{&sb, &tbuf, band_width, band_shift, max_column, score_matrix_height, gradient, min_score};
    const int64_t stride = c.stride();

    for (int32_t j = 0; j < band_width + PAD; j++)
        c.set_score(0, j, j * gap, 0);

    for (int32_t graph_pos = 0; graph_pos < graph_count; graph_pos++)
    {
        int32_t node_id      = g.sorted[graph_pos];
        int32_t score_gIdx   = graph_pos + 1;
        int32_t band_start   = band_start_for_row(score_gIdx, gradient, band_width, band_shift, max_column);
        int32_t pred_node_id = g.in_edges[node_id * MAXE];
        c.initialize_band(score_gIdx, band_start);

        int32_t first_element_prev_score = 0;
        uint16_t pred_count              = g.in_cnt[node_id];
        int32_t pred_idx                 = 0;
        {
            // vertical boundary, lane 0 (:361-441)
            int32_t penalty;
            if (pred_count == 0)
            {
                sb.wr(static_cast<int64_t>(score_gIdx % score_matrix_height) * stride, gap);
                c.tb_wr(static_cast<int64_t>(score_gIdx) * stride, -score_gIdx);
            }
            else
            {
                const int64_t index = static_cast<int64_t>(score_gIdx) * stride;
                pred_idx            = g.pos[pred_node_id] + 1;
                if ((graph_pos - pred_idx) < score_matrix_height)
                {
                    c.tb_wr(index, -(score_gIdx - pred_idx));
                    if (band_start > CPT && pred_count == 1)
                    {
                        first_element_prev_score = min_score + gap;
                    }
                    else
                    {
                        penalty = std::max<int32_t>(min_score, c.get_score(pred_idx, -1));
                        for (int32_t p = 1; p < pred_count; p++)
                        {
                            int32_t pred_idx_tmp = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                            if ((score_gIdx - pred_idx_tmp) < score_matrix_height)
                            {
                                int32_t score_tmp = c.get_score(pred_idx_tmp, -1);
                                if (penalty < score_tmp)
                                {
                                    penalty = score_tmp;
                                    c.tb_wr(index, -(score_gIdx - pred_idx_tmp));
                                }
                            }
                        }
                        first_element_prev_score = penalty + gap;
                        c.set_score(score_gIdx, -1, first_element_prev_score, band_start);
                    }
                }
                else
                {
                    penalty = min_score;
                    for (int32_t p = 1; p < pred_count; p++)
                    {
                        int32_t pred_idx_tmp = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                        if ((score_gIdx - pred_idx_tmp) < score_matrix_height)
                        {
                            int32_t score_tmp = c.get_score(pred_idx_tmp, -1);
                            if (penalty < score_tmp)
                            {
                                penalty = score_tmp;
                                c.tb_wr(index, -(score_gIdx - pred_idx_tmp));
                            }
                        }
                    }
                    first_element_prev_score = penalty + gap;
                    c.set_score(score_gIdx, -1, first_element_prev_score, band_start);
                }
            }
        }
        const uint8_t graph_base = g.nodes[node_id];

        // chunks of 128 columns; within a chunk the warp's 32 lanes x 4 cells, horizontal closure to its fixpoint (:448-546)
        for (int32_t chunk_start = band_start; chunk_start < band_start + band_width; chunk_start += MIN_BW)
        {
            ScoreT sc[MIN_BW];
            TraceT tr[MIN_BW];
            for (int32_t lane = 0; lane < 32; lane++)
            {
                int32_t read_pos = chunk_start + lane * CPT;
                int32_t prof[4];
                for (int32_t k = 0; k < 4; k++)
                {
                    // bytes past the end of the read only reach cells with column > read_length (see nw_banded)
                    uint8_t rb = (read_pos + k < read_length) ? read[read_pos + k] : 0;
                    prof[k]    = (graph_base == rb) ? match : mismatch;
                }
                ScoreT s4[4] = {min_score, min_score, min_score, min_score};
                TraceT t4[4] = {0, 0, 0, 0}; // uninitialised on the device; only cells no predecessor reaches keep it
                c.get_scores(pred_idx, score_gIdx, read_pos, gap, prof, s4, t4); // predecessor 0: no distance test (as the reference)
                for (int32_t p = 1; p < pred_count; p++)
                {
                    int32_t pred_idx_tmp = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                    if ((score_gIdx - pred_idx_tmp) < score_matrix_height)
                        c.get_scores(pred_idx_tmp, score_gIdx, read_pos, gap, prof, s4, t4);
                }
                for (int32_t k = 0; k < 4; k++)
                {
                    sc[lane * 4 + k] = s4[k];
                    tr[lane * 4 + k] = t4[k];
                }
            }
            // the relaxation loop's fixpoint: left-to-right closure; a cell's move becomes 0 iff the horizontal value is strictly
            // larger than what the predecessors gave (every update in the loop is strict and monotone)
            int32_t left = first_element_prev_score;
            for (int32_t k = 0; k < MIN_BW; k++)
            {
                if (sc[k] < left + gap)
                {
                    sc[k] = static_cast<ScoreT>(left + gap);
                    tr[k] = 0;
                }
                left = sc[k];
            }
            first_element_prev_score = sc[MIN_BW - 1];
            for (int32_t k = 0; k < MIN_BW; k++)
            {
                int64_t local = static_cast<int64_t>(chunk_start + k + 1 - band_start);
                sb.wr(local + static_cast<int64_t>(score_gIdx % score_matrix_height) * stride, sc[k]);
                c.tb_wr(local + static_cast<int64_t>(score_gIdx) * stride, tr[k]);
            }
        }
    }

    // end cell among sinks within the stored score rows (:553-579)
    int32_t aligned_nodes = 0;
    int32_t i = 0, j = read_length;
    int32_t mscore = min_score;
    for (int32_t idx = 1; idx <= graph_count; idx++)
    {
        if (g.out_cnt[g.sorted[idx - 1]] == 0)
        {
            if ((graph_count - idx) < score_matrix_height)
            {
                int32_t s = c.get_score(idx, j);
                if (mscore < s)
                {
                    mscore = s;
                    i      = idx;
                }
            }
        }
    }
    if (i == 0)
    {
        j             = 0;
        aligned_nodes = NW_TRACEBACK_BUFFER_FAILED;
    }
    // traceback over the trace matrix (:581-641)
    int32_t loop_count = 0;
    const int32_t limit = read_length + graph_count + 2;
    while (!(i == 0 && j == 0) && loop_count < limit)
    {
        loop_count++;
        int32_t band_start = band_start_for_row(i, gradient, band_width, band_shift, max_column);
        int32_t trace      = c.tb_rd(static_cast<int64_t>(j - band_start) + static_cast<int64_t>(i) * stride);
        if (trace == 0)
        {
            alignment_graph[aligned_nodes] = -1;
            alignment_read[aligned_nodes]  = j - 1;
            j--;
        }
        else if (trace < 0)
        {
            alignment_graph[aligned_nodes] = g.sorted[i - 1];
            alignment_read[aligned_nodes]  = -1;
            i += trace;
        }
        else
        {
            alignment_graph[aligned_nodes] = g.sorted[i - 1];
            alignment_read[aligned_nodes]  = j - 1;
            i -= trace;
            j--;
            if (adaptive && rerun == 0 && band_width < MAX_ADAPTIVE_BW)
            {
                int32_t threshold = std::max(1, max_column / 1024);
                if (j > threshold && j < max_column - threshold)
                {
                    int32_t bs = band_start_for_row(i, gradient, band_width, band_shift, max_column);
                    if (j <= bs + threshold)
                    {
                        aligned_nodes = SHIFT_LEFT;
                        break;
                    }
                    if (j >= (bs + band_width - threshold))
                    {
                        aligned_nodes = SHIFT_RIGHT;
                        break;
                    }
                }
            }
        }
        aligned_nodes++;
    }
    if (loop_count >= limit)
        aligned_nodes = NW_BACKTRACK_FAILED;
    return aligned_nodes;
}

// computed from out-of-read bytes on the device; they never influence columns <= read_length and are not restated.
template <typename ScoreT>
int32_t nw_full(Graph& g, int32_t graph_count, const uint8_t* read, int32_t read_length, ScoreBuf<ScoreT>& sb,
                int32_t scores_width, int32_t* alignment_graph, int32_t* alignment_read, int32_t gap, int32_t mismatch,
                int32_t match, int64_t* cells_out)
{
    const int32_t type_min = std::numeric_limits<ScoreT>::min();
    const int64_t W        = scores_width;
    if (cells_out)
        *cells_out += static_cast<int64_t>(graph_count) * read_length;
    for (int32_t j = 0; j < read_length + 1; j++)
        sb.wr(j, j * gap);
    for (int32_t graph_pos = 0; graph_pos < graph_count; graph_pos++)
    {
        int32_t node_id = g.sorted[graph_pos];
        int32_t i       = graph_pos + 1;
        uint16_t pc     = g.in_cnt[node_id];
        if (pc == 0)
        {
            sb.wr(i * W, gap);
        }
        else
        {
            int32_t penalty = type_min;
            for (int32_t p = 0; p < pc; p++)
            {
                int32_t pi = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                penalty    = std::max<int32_t>(penalty, sb.rd(pi * W));
            }
            sb.wr(i * W, penalty + gap);
        }
    }
    for (int32_t graph_pos = 0; graph_pos < graph_count; graph_pos++)
    {
        int32_t node_id = g.sorted[graph_pos];
        int32_t i       = graph_pos + 1;
        uint16_t pc     = g.in_cnt[node_id];
        uint8_t base    = g.nodes[node_id];
        int32_t last    = sb.rd(i * W);
        for (int32_t j = 1; j <= read_length; j++)
        {
            int32_t sub  = (base == read[j - 1]) ? match : mismatch;
            int32_t best = 0;
            int32_t np   = std::max<int32_t>(pc, 1);
            for (int32_t p = 0; p < np; p++)
            {
                int32_t pi = (pc == 0) ? 0 : g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                // intermediate results are held in ScoreT on the device (computeScore, cudapoa_nw.cuh:90-123)
                ScoreT cand = static_cast<ScoreT>(std::max<int32_t>(sb.rd(pi * W + j - 1) + sub, sb.rd(pi * W + j) + gap));
                best        = (p == 0) ? cand : std::max<int32_t>(best, cand);
            }
            ScoreT v = static_cast<ScoreT>(std::max<int32_t>(last + gap, best));
            sb.wr(i * W + j, v);
            last = v;
        }
    }
    int32_t aligned_nodes = 0;
    int32_t i = 0, j = read_length;
    int32_t mscore = type_min;
    for (int32_t idx = 1; idx <= graph_count; idx++)
    {
        if (g.out_cnt[g.sorted[idx - 1]] == 0)
        {
            int32_t s = sb.rd(idx * W + j);
            if (mscore < s)
            {
                mscore = s;
                i      = idx;
            }
        }
    }
    int32_t prev_i = 0, prev_j = 0, loop_count = 0;
    while (!(i == 0 && j == 0) && loop_count < (read_length + graph_count + 2))
    {
        loop_count++;
        int32_t scores_ij = sb.rd(i * W + j);
        bool pred_found   = false;
        if (i != 0 && j != 0)
        {
            int32_t node_id    = g.sorted[i - 1];
            int32_t match_cost = (g.nodes[node_id] == read[j - 1]) ? match : mismatch;
            uint16_t pc        = g.in_cnt[node_id];
            int32_t pred_i     = (pc == 0) ? 0 : (g.pos[g.in_edges[node_id * MAXE]] + 1);
            if (scores_ij == (sb.rd(pred_i * W + j - 1) + match_cost))
            {
                prev_i     = pred_i;
                prev_j     = j - 1;
                pred_found = true;
            }
            if (!pred_found)
            {
                for (int32_t p = 1; p < pc; p++)
                {
                    pred_i = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                    if (scores_ij == (sb.rd(pred_i * W + j - 1) + match_cost))
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
            int32_t node_id = g.sorted[i - 1];
            uint16_t pc     = g.in_cnt[node_id];
            int32_t pred_i  = (pc == 0) ? 0 : g.pos[g.in_edges[node_id * MAXE]] + 1;
            if (scores_ij == sb.rd(pred_i * W + j) + gap)
            {
                prev_i     = pred_i;
                prev_j     = j;
                pred_found = true;
            }
            if (!pred_found)
            {
                for (int32_t p = 1; p < pc; p++)
                {
                    pred_i = g.pos[g.in_edges[node_id * MAXE + p]] + 1;
                    if (scores_ij == sb.rd(pred_i * W + j) + gap)
                    {
                        prev_i     = pred_i;
                        prev_j     = j;
                        pred_found = true;
                        break;
                    }
                }
            }
        }
        if (!pred_found && j != 0 && scores_ij == sb.rd(i * W + j - 1) + gap)
        {
            prev_i     = i;
            prev_j     = j - 1;
            pred_found = true;
        }
        alignment_graph[aligned_nodes] = (i == prev_i) ? -1 : g.sorted[i - 1];
        alignment_read[aligned_nodes]  = (j == prev_j) ? -1 : j - 1;
        aligned_nodes++;
        i = prev_i;
        j = prev_j;
    }
    if (loop_count >= (read_length + graph_count + 2))
        aligned_nodes = NW_BACKTRACK_FAILED;
    return aligned_nodes;
}

// ---------------------------------------------------------------------------------------------------------
// branchCompletion -- cudapoa_generate_consensus.cuh:35-119
int32_t branch_completion(Graph& g, int32_t max_score_id_pos, int32_t node_count, std::vector<int32_t>& scores, std::vector<int32_t>& preds)
{
    int32_t node_id = g.sorted[max_score_id_pos];
    uint16_t oe_n   = g.out_cnt[node_id];
    for (int32_t oe = 0; oe < oe_n; oe++)
    {
        int32_t out_node = g.out_edges[node_id * MAXE + oe];
        uint16_t ie_n    = g.in_cnt[out_node];
        for (int32_t ie = 0; ie < ie_n; ie++)
        {
            int32_t id = g.in_edges[out_node * MAXE + ie];
            if (id != node_id)
                scores[id] = -1;
        }
    }
    int32_t max_score = 0, max_score_id = 0;
    for (int32_t gp = max_score_id_pos + 1; gp < node_count; gp++)
    {
        node_id        = g.sorted[gp];
        preds[node_id] = -1;
        int32_t score  = -1;
        uint16_t in_n  = g.in_cnt[node_id];
        for (int32_t e = 0; e < in_n; e++)
        {
            int32_t b = g.in_edges[node_id * MAXE + e];
            if (scores[b] == -1)
                continue;
            int32_t w = static_cast<int32_t>(g.in_w[node_id * MAXE + e]);
            if (score < w || (score == w && scores[preds[node_id]] <= scores[b]))
            {
                score          = w;
                preds[node_id] = b;
            }
        }
        if (preds[node_id] != -1)
            score += scores[preds[node_id]];
        if (max_score <= score)
        {
            max_score    = score;
            max_score_id = node_id;
        }
        scores[node_id] = score;
    }
    return max_score_id;
}

// generateConsensus -- cudapoa_generate_consensus.cuh:141-283. Writes the consensus REVERSED + NUL like the device;
// error protocol consensus[0]=0xFF, consensus[1]=status.
void generate_consensus(Graph& g, int32_t node_count, uint8_t* consensus, uint16_t* coverage, int32_t max_consensus)
{
    std::vector<int32_t> preds(std::max(node_count, 1), -1), scores(std::max(node_count, 1), -1);
    int32_t max_score_id = 0, max_score = -1;
    for (int32_t gp = 0; gp < node_count; gp++)
    {
        int32_t node_id = g.sorted[gp];
        uint16_t in_n   = g.in_cnt[node_id];
        int32_t score   = scores[node_id];
        for (int32_t e = 0; e < in_n; e++)
        {
            int32_t w = static_cast<int32_t>(g.in_w[node_id * MAXE + e]);
            int32_t b = g.in_edges[node_id * MAXE + e];
            if (score < w || (score == w && scores[preds[node_id]] <= scores[b]))
            {
                score          = w;
                preds[node_id] = b;
            }
        }
        if (preds[node_id] != -1)
            score += scores[preds[node_id]];
        if (max_score <= score)
        {
            max_score_id = node_id;
            max_score    = score;
        }
        scores[node_id] = score;
    }
    int32_t loop_count = 0;
    if (g.out_cnt[max_score_id] != 0)
    {
        while (g.out_cnt[max_score_id] != 0 && loop_count < node_count)
        {
            max_score_id = branch_completion(g, g.pos[max_score_id], node_count, scores, preds);
            loop_count++;
        }
    }
    if (loop_count >= node_count)
    {
        consensus[0] = 0xFF;
        consensus[1] = static_cast<uint8_t>(loop_count_exceeded_upper_bound);
        return;
    }
    int32_t cpos = 0, ccount = 0;
    while (preds[max_score_id] != -1)
    {
        consensus[cpos] = g.nodes[max_score_id];
        uint16_t cov    = g.coverage[max_score_id];
        for (int32_t a = 0; a < g.aln_cnt[max_score_id]; a++)
            cov = static_cast<uint16_t>(cov + g.coverage[g.aligned[max_score_id * MAXA + a]]);
        coverage[cpos] = cov;
        max_score_id   = preds[max_score_id];
        cpos           = std::min(cpos + 1, max_consensus - 1);
        ccount++;
    }
    consensus[cpos] = g.nodes[max_score_id];
    uint16_t cov    = g.coverage[max_score_id];
    for (int32_t a = 0; a < g.aln_cnt[max_score_id]; a++)
        cov = static_cast<uint16_t>(cov + g.coverage[g.aligned[max_score_id * MAXA + a]]);
    coverage[cpos] = cov;
    if (ccount >= (max_consensus - 1))
    {
        consensus[0] = 0xFF;
        consensus[1] = static_cast<uint8_t>(exceeded_maximum_sequence_size);
        return;
    }
    cpos++;
    consensus[cpos] = 0;
}

// generateMSAKernel -- cudapoa_generate_msa.cuh:34-227 (racon sort, column assignment, per-read walk)
void generate_msa(Graph& g, int32_t node_count, int32_t num_seqs, uint8_t* consensus, uint8_t* msa, int32_t max_consensus)
{
    racon_topsort(g, node_count);
    std::vector<int32_t> msa_pos_of(std::max(node_count, 1), 0);
    int32_t msa_pos = 0;
    for (int32_t rank = 0; rank < node_count; rank++)
    {
        int32_t node_id     = g.sorted[rank];
        msa_pos_of[node_id] = msa_pos;
        uint16_t ac         = g.aln_cnt[node_id];
        for (int32_t n = 0; n < ac; n++)
            msa_pos_of[g.sorted[++rank]] = msa_pos;
        msa_pos++;
    }
    int32_t msa_length = msa_pos;
    if (msa_length >= max_consensus)
    {
        consensus[0] = 0xFF;
        consensus[1] = static_cast<uint8_t>(exceeded_maximum_sequence_size);
        return;
    }
    for (int32_t s = 0; s < num_seqs; s++)
    {
        uint8_t* row         = msa + static_cast<int64_t>(s) * max_consensus;
        int32_t node_id      = g.seq_begin[s];
        int32_t filled_until = 0;
        while (true)
        {
            int32_t mp = msa_pos_of[node_id];
            row[mp]    = g.nodes[node_id];
            for (int32_t i = filled_until; i < mp; i++)
                row[i] = '-';
            filled_until  = mp + 1;
            bool end_node = true;
            for (int32_t n = 0; n < g.out_cnt[node_id]; n++)
            {
                int32_t to = g.out_edges[node_id * MAXE + n];
                for (int32_t m = 0; m < g.out_cov_cnt[node_id * MAXE + n]; m++)
                {
                    if (g.out_cov[static_cast<size_t>(node_id * MAXE + n) * g.max_seqs + m] == s)
                    {
                        end_node = false;
                        node_id  = to;
                        break;
                    }
                }
                if (!end_node)
                    break;
            }
            if (end_node)
            {
                for (int32_t i = filled_until; i < msa_length; i++)
                    row[i] = '-';
                break;
            }
        }
        row[msa_length] = 0;
    }
}

// cudapoa_limits.hpp:34-59
bool use32bit_score(int32_t max_seq, int32_t max_nodes, int32_t gap, int32_t mismatch, int32_t match)
{
    int32_t upper = max_seq * match;
    int32_t lower = max_seq * std::max(gap, mismatch) + (max_nodes - max_seq) * gap;
    return (upper > INT16_MAX || (-lower) > (INT16_MAX + 1));
}

struct WindowCfg
{
    int32_t max_seq_size, max_consensus, max_nodes, matrix_seq_dim, band_width, max_seqs, band_mode, max_pred_dist;
    int32_t gap, mismatch, match;
    int32_t msa;
};

// generatePOAKernel -- cudapoa_kernels.cuh:76-542, then generateConsensusKernel / generateMSAKernel (:1023-1072).
// Host-side finishing as Batch::get_consensus / get_msa (cudapoa_batch.cuh:229-257, 290-310): reverse, decode status.
template <typename ScoreT>
int32_t run_window(const WindowCfg& cfg, int32_t num_seqs, const int32_t* seq_len, const uint8_t* seq_data, const int8_t* weights_in,
                   char* consensus_out, uint16_t* coverage_out, char* msa_out, int64_t* cells_out, int32_t* node_count_out, Graph* graph_out)
{
    Graph local;
    Graph& g = graph_out ? *graph_out : local;
    g.init(cfg.max_nodes, cfg.max_seqs, cfg.msa != 0);
    std::vector<int64_t> off(num_seqs + 1, 0);
    for (int32_t s = 0; s < num_seqs; s++)
        off[s + 1] = off[s] + seq_len[s];
    std::vector<int8_t> ones;
    const int8_t* weights = weights_in;
    if (!weights)
    {
        ones.assign(off[num_seqs] + 1, 1);
        weights = ones.data();
    }
    std::vector<uint8_t> cons(std::max(cfg.max_consensus, 2) + 1, 0);
    std::vector<uint16_t> cov(std::max(cfg.max_consensus, 2) + 1, 0);
    consensus_out[0] = 0;
    if (node_count_out)
        *node_count_out = 0;

    // backbone (:200-238)
    const uint8_t* seq0 = seq_data;
    int32_t node_count  = seq_len[0];
    if (node_count > cfg.max_nodes)
        return generic_error; // host admission (max_sequence_size <= max_nodes) prevents this
    if (node_count > 0)
    {
        g.nodes[0]    = seq0[0];
        g.sorted[0]   = 0;
        g.in_cnt[0]   = 0;
        g.aln_cnt[0]  = 0;
        g.pos[0]      = 0;
        g.out_cnt[node_count - 1] = 0;
        g.in_w[0]     = static_cast<uint16_t>(weights[0]);
        g.coverage[0] = 1;
        if (cfg.msa)
            g.seq_begin[0] = 0;
        for (int32_t n = 1; n < node_count; n++)
        {
            g.nodes[n]                  = seq0[n];
            g.sorted[n]                 = n;
            g.out_edges[(n - 1) * MAXE] = n;
            g.out_cnt[n - 1]            = 1;
            g.in_edges[n * MAXE]        = n - 1;
            g.in_w[n * MAXE]            = static_cast<uint16_t>(weights[n - 1] + weights[n]);
            g.in_cnt[n]                 = 1;
            g.aln_cnt[n]                = 0;
            g.pos[n]                    = n;
            g.coverage[n]               = 1;
            if (cfg.msa)
            {
                g.out_cov[static_cast<size_t>((n - 1) * MAXE) * g.max_seqs] = 0;
                g.out_cov_cnt[(n - 1) * MAXE]                               = 1;
            }
        }
    }
    cons[0] = 0;

    ScoreBuf<ScoreT> sb;
    int64_t buf_elems;
    if (cfg.band_mode == full_band)
    {
        int32_t maxlen = 0;
        for (int32_t s = 0; s < num_seqs; s++)
            maxlen = std::max(maxlen, seq_len[s]);
        buf_elems = static_cast<int64_t>(cfg.max_nodes) * align_up(maxlen + 1 + CPT, 4); // cudapoa_batch.cuh:502-507
    }
    else
    {
        buf_elems = static_cast<int64_t>(cfg.max_nodes) * cfg.matrix_seq_dim;
    }
    const bool tb_mode = cfg.band_mode == static_band_traceback || cfg.band_mode == adaptive_band_traceback;
    ScoreBuf<ScoreT> sb_tb; // traceback modes: max_pred_dist score rows (allocate_block.hpp:333-334) + the trace matrix
    std::vector<int8_t> trace8;
    std::vector<int16_t> trace16;
    if (tb_mode)
    {
        sb_tb.data.assign(static_cast<int64_t>(cfg.max_pred_dist) * cfg.matrix_seq_dim, 0);
        if (cfg.max_pred_dist > 127)
            trace16.assign(buf_elems, 0);
        else
            trace8.assign(buf_elems, 0);
        buf_elems = 0;
    }
    sb.data.assign(buf_elems, 0);
    const float banded_buffer_size = static_cast<float>(cfg.max_nodes) * static_cast<float>(cfg.matrix_seq_dim);
    int32_t scores_width           = 0;
    for (int32_t s = 0; s < num_seqs; s++)
        scores_width = std::max(scores_width, align_up(seq_len[s] + 1 + CPT, 4));

    std::vector<int32_t> aln_graph(cfg.max_nodes + 16, 0), aln_read(cfg.max_nodes + 16, 0);
    // the alignment can be as long as read_length + graph_count + 2 entries; the device buffers have max_nodes entries
    aln_graph.resize(static_cast<size_t>(cfg.max_nodes) * 2 + cfg.max_seq_size + 16);
    aln_read.resize(aln_graph.size());

    uint8_t err = 0;
    for (int32_t s = 1; s < num_seqs && !err; s++)
    {
        const uint8_t* read = seq_data + off[s];
        const int8_t* w     = weights + off[s];
        int32_t len         = seq_len[s];
        if (node_count >= cfg.max_nodes)
        {
            err = static_cast<uint8_t>(node_count_exceeded_maximum_graph_size);
            break;
        }
        int32_t alen = 0;
        if (cfg.band_mode == adaptive_band && cfg.band_width < MAX_ADAPTIVE_BW)
        {
            alen = nw_banded<ScoreT>(g, node_count, read, len, sb, banded_buffer_size, aln_graph.data(), aln_read.data(), cfg.band_width,
                                     cfg.gap, cfg.mismatch, cfg.match, 0, true, cells_out);
            if (alen == SHIFT_LEFT || alen == SHIFT_RIGHT)
                alen = nw_banded<ScoreT>(g, node_count, read, len, sb, banded_buffer_size, aln_graph.data(), aln_read.data(), cfg.band_width,
                                         cfg.gap, cfg.mismatch, cfg.match, alen, true, cells_out);
        }
        else if (cfg.band_mode == static_band || cfg.band_mode == adaptive_band)
        {
            alen = nw_banded<ScoreT>(g, node_count, read, len, sb, banded_buffer_size, aln_graph.data(), aln_read.data(), cfg.band_width,
                                     cfg.gap, cfg.mismatch, cfg.match, 0, false, cells_out);
        }
        else if (cfg.band_mode == full_band)
        {
            alen = nw_full<ScoreT>(g, node_count, read, len, sb, scores_width, aln_graph.data(), aln_read.data(), cfg.gap, cfg.mismatch,
                                   cfg.match, cells_out);
        }
        else
        {
            // static_band_traceback / adaptive_band_traceback (cudapoa_kernels.cuh:270-346): scores keep max_pred_dist rows,
            // the trace matrix max_nodes rows; TraceT = int8 unless max_banded_pred_distance > 127 (cudapoa_limits.hpp:56-60)
            const bool adaptive_tb = cfg.band_mode == adaptive_band_traceback && cfg.band_width < MAX_ADAPTIVE_BW;
            auto run_tb = [&](int32_t rerun, bool adaptive) {
                if (cfg.max_pred_dist > 127)
                    return nw_banded_tb<ScoreT, int16_t>(g, node_count, read, len, sb_tb, trace16, banded_buffer_size, aln_graph.data(),
                                                          aln_read.data(), cfg.band_width, cfg.max_pred_dist, cfg.gap, cfg.mismatch, cfg.match,
                                                          rerun, adaptive, cells_out);
                return nw_banded_tb<ScoreT, int8_t>(g, node_count, read, len, sb_tb, trace8, banded_buffer_size, aln_graph.data(), aln_read.data(),
                                                     cfg.band_width, cfg.max_pred_dist, cfg.gap, cfg.mismatch, cfg.match, rerun, adaptive, cells_out);
            };
            if (adaptive_tb)
            {
                alen = run_tb(0, true);
                if (alen == SHIFT_LEFT || alen == SHIFT_RIGHT)
                    alen = run_tb(alen, true);
            }
            else
            {
                alen = run_tb(0, false);
            }
            if (alen == NW_TRACEBACK_BUFFER_FAILED)
            {
                err = static_cast<uint8_t>(exceeded_maximum_predecessor_distance);
                break;
            }
        }
        if (alen == NW_BACKTRACK_FAILED)
        {
            err = static_cast<uint8_t>(loop_count_exceeded_upper_bound);
            break;
        }
        if (alen == NW_ADAPTIVE_STORAGE_FAILED)
        {
            err = static_cast<uint8_t>(exceeded_adaptive_banded_matrix_size);
            break;
        }
        if (alen < 0)
        {
            // a second SHIFT code after the rerun is passed to addAlignmentToGraph as a negative length on the device
            // (loop does not execute); restated as a no-op alignment.
            alen = 0;
        }
        int32_t new_count = node_count;
        uint8_t e         = add_alignment(g, new_count, node_count, alen, aln_graph.data(), read, aln_read.data(), w, s);
        if (e != 0)
        {
            err = e;
            break;
        }
        node_count = new_count;
        topsort(g, node_count);
    }
    g.node_count = node_count;
    if (node_count_out)
        *node_count_out = node_count;
    if (err)
        return err;
    if (cfg.msa)
    {
        std::vector<uint8_t> m(static_cast<size_t>(cfg.max_seqs) * cfg.max_consensus + 1, 0);
        generate_msa(g, node_count, num_seqs, cons.data(), m.data(), cfg.max_consensus);
        if (cons[0] == 0xFF)
            return cons[1];
        if (msa_out)
            std::memcpy(msa_out, m.data(), static_cast<size_t>(cfg.max_seqs) * cfg.max_consensus);
        return success;
    }
    generate_consensus(g, node_count, cons.data(), cov.data(), cfg.max_consensus);
    if (cons[0] == 0xFF)
        return cons[1];
    int32_t n = static_cast<int32_t>(std::strlen(reinterpret_cast<char*>(cons.data())));
    for (int32_t k = 0; k < n; k++)
    {
        consensus_out[k] = static_cast<char>(cons[n - 1 - k]);
        if (coverage_out)
            coverage_out[k] = cov[n - 1 - k];
    }
    consensus_out[n] = 0;
    return success;
}

} // namespace

extern "C" {

void oracle_set_fdiv(float (*f)(float, float)) { g_fdiv = f ? f : ieee_fdiv; }

// BatchConfig derivation -- cudapoa/src/batch.cu:34-71. out[0..7] = max_sequence_size, max_consensus_size, max_nodes_per_graph,
// matrix_sequence_dimension, alignment_band_width, max_sequences_per_poa, band_mode, max_banded_pred_distance.
void oracle_poa_batch_config(int32_t max_seq_sz, int32_t max_seq_per_poa, int32_t band_width, int32_t band_mode, float adaptive_storage_factor,
                             float graph_length_factor, int32_t max_pred_dist, int32_t* out)
{
    int32_t abw = align_up(band_width, MIN_BW);
    out[0]      = max_seq_sz;
    out[1]      = 2 * max_seq_sz;
    out[2]      = align_up(static_cast<int32_t>(graph_length_factor * max_seq_sz), CPT);
    if (band_mode == full_band)
        out[3] = align_up(max_seq_sz, CPT);
    else if (band_mode == static_band || band_mode == static_band_traceback)
        out[3] = align_up(abw + PAD, CPT);
    else
        out[3] = align_up(static_cast<int32_t>(adaptive_storage_factor * (abw + PAD)), CPT);
    out[4] = abw;
    out[5] = max_seq_per_poa;
    out[6] = band_mode;
    out[7] = max_pred_dist > 0 ? max_pred_dist : 2 * abw;
}

int32_t oracle_poa_use32bit_score(int32_t max_seq, int32_t max_nodes, int32_t gap, int32_t mismatch, int32_t match)
{
    return use32bit_score(max_seq, max_nodes, gap, mismatch, match) ? 1 : 0;
}

// Full pipeline for a flat list of windows (same flat format as oracle/ref_capi.cu and include/gwb200.h).
// cfg8 = the 8 BatchConfig fields above. Outputs as Batch::get_consensus / get_msa present them.
// cells[n_windows] (optional) receives the executed DP cell count (sum graph_count x band_width incl. reruns).
int32_t oracle_poa_run(int32_t n_windows, const int32_t* win_nseq, const int32_t* seq_len, const char* seq_data, const int8_t* weights,
                       const int32_t* cfg8, int32_t output_msa, int32_t gap, int32_t mismatch, int32_t match,
                       char* consensus, uint16_t* coverage, int32_t* status, char* msa, int64_t* cells, int32_t* node_counts)
{
    WindowCfg cfg{cfg8[0], cfg8[1], cfg8[2], cfg8[3], cfg8[4], cfg8[5], cfg8[6], cfg8[7], gap, mismatch, match, output_msa};
    const bool s32 = use32bit_score(cfg.max_seq_size, cfg.max_nodes, gap, mismatch, match);
    int64_t off    = 0;
    int32_t si     = 0;
    for (int32_t w = 0; w < n_windows; w++)
    {
        const int32_t ns = win_nseq[w];
        char* c          = consensus + static_cast<int64_t>(w) * cfg.max_consensus;
        uint16_t* cv     = coverage ? coverage + static_cast<int64_t>(w) * cfg.max_consensus : nullptr;
        char* m          = msa ? msa + static_cast<int64_t>(w) * cfg.max_seqs * cfg.max_consensus : nullptr;
        int64_t cell     = 0;
        int32_t nc       = 0;
        const uint8_t* sd = reinterpret_cast<const uint8_t*>(seq_data) + off;
        const int8_t* wt  = weights ? weights + off : nullptr;
        int32_t st;
        if (s32)
            st = run_window<int32_t>(cfg, ns, seq_len + si, sd, wt, c, cv, m, &cell, &nc, nullptr);
        else
            st = run_window<int16_t>(cfg, ns, seq_len + si, sd, wt, c, cv, m, &cell, &nc, nullptr);
        status[w] = st;
        if (st != success)
            c[0] = 0;
        if (cells)
            cells[w] = cell;
        if (node_counts)
            node_counts[w] = nc;
        for (int32_t s = 0; s < ns; s++)
            off += seq_len[si + s];
        si += ns;
    }
    return 0;
}

// ---- single-stage entry points used by the known-answer tests (mirror the reference's test-only launchers) ----

// runTopSort -- cudapoa_topsort.cuh:199+ ; adjacency given as out-edge lists (flattened, out_off[n+1]).
void oracle_topsort(int32_t node_count, const int32_t* out_off, const int32_t* out_adj, int32_t* sorted_out)
{
    Graph g;
    g.init(std::max(node_count, 1), 1, false);
    for (int32_t n = 0; n < node_count; n++)
    {
        g.out_cnt[n] = static_cast<uint16_t>(out_off[n + 1] - out_off[n]);
        for (int32_t e = out_off[n]; e < out_off[n + 1]; e++)
        {
            g.out_edges[n * MAXE + (e - out_off[n])] = out_adj[e];
            g.in_cnt[out_adj[e]]++;
        }
    }
    topsort(g, node_count);
    for (int32_t n = 0; n < node_count; n++)
        sorted_out[n] = g.sorted[n];
}

struct OracleGraphHandle
{
    Graph g;
};

// Build a graph from explicit arrays (the layout of the reference tests' BasicGraph / SortedGraph helpers,
// cudapoa/tests/basic_graph.hpp, sorted_graph.hpp): nodes, sorted order, out-edge lists, optional raw in-edge weights.
OracleGraphHandle* oracle_graph_create(int32_t node_count, int32_t max_nodes, const uint8_t* nodes, const int32_t* sorted,
                                       const int32_t* out_off, const int32_t* out_adj, const uint16_t* in_w_raw,
                                       const uint16_t* coverage)
{
    OracleGraphHandle* h = new OracleGraphHandle;
    Graph& g             = h->g;
    g.init(max_nodes, 1, false);
    g.node_count = node_count;
    for (int32_t n = 0; n < node_count; n++)
    {
        g.nodes[n]    = nodes[n];
        g.sorted[n]   = sorted ? sorted[n] : n;
        g.coverage[n] = coverage ? coverage[n] : 0;
    }
    for (int32_t n = 0; n < node_count; n++)
        g.pos[g.sorted[n]] = n;
    // incoming edges are filled in node order over outgoing lists, as the reference helpers do (basic_graph.hpp:80-112)
    for (int32_t n = 0; n < node_count; n++)
    {
        for (int32_t e = out_off[n]; e < out_off[n + 1]; e++)
        {
            int32_t to                         = out_adj[e];
            g.out_edges[n * MAXE + g.out_cnt[n]] = to;
            g.out_cnt[n]++;
            g.in_edges[to * MAXE + g.in_cnt[to]] = n;
            g.in_cnt[to]++;
        }
    }
    // in_w_raw is the raw [node*MAXE+slot] weight array exactly as the reference test harness fills it
    // (Test_CudapoaGenerateConsensus.cu get_incoming_edge_w indexes the slot by FROM-node id; untouched slots are 0)
    if (in_w_raw)
        for (int64_t k = 0; k < static_cast<int64_t>(node_count) * MAXE; k++)
            g.in_w[k] = in_w_raw[k];
    return h;
}

void oracle_graph_destroy(OracleGraphHandle* h) { delete h; }

int32_t oracle_graph_node_count(OracleGraphHandle* h) { return h->g.node_count; }

// returns number of out edges of node n, writes them to out[]
int32_t oracle_graph_out_edges(OracleGraphHandle* h, int32_t n, int32_t* out)
{
    for (int32_t e = 0; e < h->g.out_cnt[n]; e++)
        out[e] = h->g.out_edges[n * MAXE + e];
    return h->g.out_cnt[n];
}

void oracle_graph_set_alignments(OracleGraphHandle* h, int32_t n, int32_t count, const int32_t* aligned)
{
    h->g.aln_cnt[n] = static_cast<uint16_t>(count);
    for (int32_t a = 0; a < count; a++)
        h->g.aligned[n * MAXA + a] = aligned[a];
}

// addAlignment wrapper -- cudapoa_add_alignment.cuh:288+ ; then topsort so that the graph stays usable
int32_t oracle_graph_add_alignment(OracleGraphHandle* h, int32_t alignment_length, const int32_t* alignment_graph, const int32_t* alignment_read,
                                   const uint8_t* read, const int8_t* base_weights)
{
    Graph& g     = h->g;
    int32_t newc = g.node_count;
    uint8_t e    = add_alignment(g, newc, g.node_count, alignment_length, alignment_graph, read, alignment_read, base_weights, 1);
    if (e == 0)
    {
        g.node_count = newc;
    }
    return e;
}

// NW single-stage: mode 0 = full (runNW), 1 = static band (runNWbanded<false>), 2 = adaptive band (with the kernel-level rerun).
// Returns the alignment length or a negative code. int16 scores, like the reference's test launchers.
int32_t oracle_graph_nw(OracleGraphHandle* h, const uint8_t* read, int32_t read_length, int32_t mode, int32_t band_width,
                        int32_t max_nodes, int32_t matrix_seq_dim, int32_t gap, int32_t mismatch, int32_t match,
                        int32_t* alignment_graph, int32_t* alignment_read)
{
    Graph& g = h->g;
    ScoreBuf<int16_t> sb;
    if (mode == 0)
    {
        int32_t W = align_up(read_length + 1 + CPT, 4);
        sb.data.assign(static_cast<size_t>(max_nodes) * W, 0);
        return nw_full<int16_t>(g, g.node_count, read, read_length, sb, W, alignment_graph, alignment_read, gap, mismatch, match, nullptr);
    }
    float bufsz = static_cast<float>(max_nodes) * static_cast<float>(matrix_seq_dim);
    if (mode == static_band_traceback || mode == adaptive_band_traceback)
    {
        // the harness of Test_CudapoaNW.cu:306-442 (runNWbandedTB): one call, rerun = 0, int16 scores and traces,
        // score_matrix_height = BatchConfig::max_banded_pred_distance = 2 x band width (batch.cu:46)
        const int32_t height = 2 * align_up(band_width, MIN_BW);
        sb.data.assign(static_cast<size_t>(height) * matrix_seq_dim, 0);
        std::vector<int16_t> trace(static_cast<size_t>(max_nodes) * matrix_seq_dim, 0);
        return nw_banded_tb<int16_t, int16_t>(g, g.node_count, read, read_length, sb, trace, bufsz, alignment_graph, alignment_read, band_width,
                                              height, gap, mismatch, match, 0, mode == adaptive_band_traceback, nullptr);
    }
    sb.data.assign(static_cast<size_t>(max_nodes) * matrix_seq_dim, 0);
    if (mode == 1)
        return nw_banded<int16_t>(g, g.node_count, read, read_length, sb, bufsz, alignment_graph, alignment_read, band_width, gap, mismatch,
                                  match, 0, false, nullptr);
    int32_t r = nw_banded<int16_t>(g, g.node_count, read, read_length, sb, bufsz, alignment_graph, alignment_read, band_width, gap, mismatch,
                                   match, 0, true, nullptr);
    if (r == SHIFT_LEFT || r == SHIFT_RIGHT)
        r = nw_banded<int16_t>(g, g.node_count, read, read_length, sb, bufsz, alignment_graph, alignment_read, band_width, gap, mismatch, match,
                               r, true, nullptr);
    return r;
}

// generateConsensusTestHost -- cudapoa_generate_consensus.cuh:356-437: returns the consensus as the device wrote it (REVERSED).
int32_t oracle_graph_consensus(OracleGraphHandle* h, int32_t max_consensus, char* consensus_reversed, uint16_t* coverage_reversed)
{
    Graph& g = h->g;
    std::vector<uint8_t> c(max_consensus + 2, 0);
    std::vector<uint16_t> cv(max_consensus + 2, 0);
    generate_consensus(g, g.node_count, c.data(), cv.data(), max_consensus);
    if (c[0] == 0xFF)
        return c[1];
    std::strcpy(consensus_reversed, reinterpret_cast<char*>(c.data()));
    if (coverage_reversed)
        std::memcpy(coverage_reversed, cv.data(), sizeof(uint16_t) * std::strlen(consensus_reversed));
    return 0;
}

} // extern "C"
