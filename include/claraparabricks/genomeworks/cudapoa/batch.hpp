// gw-b200: cudapoa::Batch API -- Entry, Group, BatchConfig, Batch, create_batch with the reference's signatures
// (cudapoa/include/claraparabricks/genomeworks/cudapoa/batch.hpp:41-204). The implementation is a header-only adapter over
// the C ABI of libgwb200.so (include/gwb200.h); all work happens in the sm_100a engine.
#pragma once

#include "cudapoa.hpp"
#include "../utils/allocator.hpp"
#include "../utils/graph.hpp"

#include <cuda_runtime_api.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace claraparabricks
{
namespace genomeworks
{
namespace cudapoa
{

/// One sequence of a POA group: bases, optional per-base weights (nullptr = 1), length.
struct Entry
{
    const char* seq;
    const int8_t* weights;
    int32_t length;
};

/// A POA group (window): the sequences fused into one partial order graph.
typedef std::vector<Entry> Group;

/// Upper limits of a batch; same 8 fields and two constructors as the reference (batch.hpp:60-86, batch.cu:34-104).
struct BatchConfig
{
    int32_t max_sequence_size;
    int32_t max_consensus_size;
    int32_t max_nodes_per_graph;
    int32_t matrix_sequence_dimension;
    int32_t alignment_band_width;
    int32_t max_sequences_per_poa;
    BandMode band_mode;
    int32_t max_banded_pred_distance;

    BatchConfig(int32_t max_seq_sz = 1024, int32_t max_seq_per_poa = 100, int32_t band_width = 256, BandMode banding = BandMode::full_band,
                float adapive_storage_factor = 2.0, float graph_length_factor = 3.0, int32_t max_pred_dist = 0)
    {
        gwb200_poa_config c;
        detail::check(gwb200_poa_config_init(&c, max_seq_sz, max_seq_per_poa, band_width, static_cast<int32_t>(banding), adapive_storage_factor,
                                             graph_length_factor, max_pred_dist));
        assign(c);
    }
    BatchConfig(int32_t max_seq_sz, int32_t max_consensus_sz, int32_t max_nodes_per_poa, int32_t band_width, int32_t max_seq_per_poa,
                int32_t matrix_seq_dim, BandMode banding, int32_t max_pred_dist)
    {
        gwb200_poa_config c;
        detail::check(gwb200_poa_config_init_explicit(&c, max_seq_sz, max_consensus_sz, max_nodes_per_poa, band_width, max_seq_per_poa,
                                                      matrix_seq_dim, static_cast<int32_t>(banding), max_pred_dist));
        assign(c);
    }
    gwb200_poa_config to_c() const
    {
        gwb200_poa_config c;
        c.max_sequence_size         = max_sequence_size;
        c.max_consensus_size        = max_consensus_size;
        c.max_nodes_per_graph       = max_nodes_per_graph;
        c.matrix_sequence_dimension = matrix_sequence_dimension;
        c.alignment_band_width      = alignment_band_width;
        c.max_sequences_per_poa     = max_sequences_per_poa;
        c.band_mode                 = static_cast<int32_t>(band_mode);
        c.max_banded_pred_distance  = max_banded_pred_distance;
        return c;
    }

private:
    void assign(const gwb200_poa_config& c)
    {
        max_sequence_size         = c.max_sequence_size;
        max_consensus_size        = c.max_consensus_size;
        max_nodes_per_graph       = c.max_nodes_per_graph;
        matrix_sequence_dimension = c.matrix_sequence_dimension;
        alignment_band_width      = c.alignment_band_width;
        max_sequences_per_poa     = c.max_sequences_per_poa;
        band_mode                 = static_cast<BandMode>(c.band_mode);
        max_banded_pred_distance  = c.max_banded_pred_distance;
    }
};

/// Batched GPU POA object (batch.hpp:90-162). Not thread safe; one Batch per host thread and stream.
class Batch
{
public:
    virtual ~Batch() = default;
    virtual StatusType add_poa_group(std::vector<StatusType>& per_seq_status, const Group& poa_group) = 0;
    virtual int32_t get_total_poas() const                                                          = 0;
    virtual void generate_poa()                                                                     = 0;
    virtual StatusType get_consensus(std::vector<std::string>& consensus, std::vector<std::vector<uint16_t>>& coverage,
                                     std::vector<genomeworks::cudapoa::StatusType>& output_status)  = 0;
    virtual StatusType get_msa(std::vector<std::vector<std::string>>& msa, std::vector<StatusType>& output_status) = 0;
    virtual void get_graphs(std::vector<DirectedGraph>& graphs, std::vector<StatusType>& output_status)            = 0;
    virtual int32_t batch_id() const                                                                               = 0;
    virtual void reset()                                                                                           = 0;
};

namespace detail
{
class BatchB200 : public Batch
{
public:
    BatchB200(int32_t device_id, cudaStream_t stream, int64_t max_gpu_mem, int8_t output_mask, const BatchConfig& cfg, int16_t gap_score,
              int16_t mismatch_score, int16_t match_score)
        : cfg_(cfg)
    {
#ifdef SPOA_ACCURATE
        gwb200_poa_set_spoa_accurate(1); // the reference's build flag (cudapoa_kernels.cuh:508-520)
#endif
        const gwb200_poa_config c = cfg.to_c();
        check(gwb200_poa_batch_create(&h_, device_id, stream, max_gpu_mem, output_mask, &c, gap_score, mismatch_score, match_score));
    }
    /// The whole batch lives in one block of the caller's allocator (batch.hpp:176-189; allocate_block.hpp:48-100 takes its
    /// device buffer from the allocator the same way). The block goes back to the pool when the batch is destroyed.
    BatchB200(int32_t device_id, cudaStream_t stream, DefaultDeviceAllocator allocator, int64_t block_bytes, int8_t output_mask,
              const BatchConfig& cfg, int16_t gap_score, int16_t mismatch_score, int16_t match_score)
        : cfg_(cfg)
        , allocator_(allocator)
        , block_bytes_(block_bytes)
    {
#ifdef SPOA_ACCURATE
        gwb200_poa_set_spoa_accurate(1); // the reference's build flag (cudapoa_kernels.cuh:508-520)
#endif
        const gwb200_poa_config c = cfg.to_c();
        block_                    = allocator_.allocate(static_cast<std::size_t>(block_bytes_), {stream});
        try
        {
            check(gwb200_poa_batch_create_in_block(&h_, device_id, stream, block_, block_bytes_, output_mask, &c, gap_score, mismatch_score,
                                                   match_score));
        }
        catch (...)
        {
            allocator_.deallocate(block_, static_cast<std::size_t>(block_bytes_));
            throw;
        }
    }
    ~BatchB200() override
    {
        gwb200_poa_batch_destroy(h_);
        if (block_ != nullptr)
            allocator_.deallocate(block_, static_cast<std::size_t>(block_bytes_));
    }
    BatchB200(const BatchB200&) = delete;
    BatchB200& operator=(const BatchB200&) = delete;

    StatusType add_poa_group(std::vector<StatusType>& per_seq_status, const Group& poa_group) override
    {
        const int32_t n = static_cast<int32_t>(poa_group.size());
        std::vector<const char*> seqs(n);
        std::vector<const int8_t*> weights(n);
        std::vector<int32_t> lengths(n), st(n);
        bool any_w = false;
        for (int32_t i = 0; i < n; ++i)
        {
            seqs[i]    = poa_group[i].seq;
            weights[i] = poa_group[i].weights;
            lengths[i] = poa_group[i].length;
            any_w      = any_w || poa_group[i].weights != nullptr;
        }
        int32_t n_st = 0;
        const int rc = check(gwb200_poa_batch_add_group(h_, n, seqs.data(), any_w ? weights.data() : nullptr, lengths.data(), st.data(), &n_st));
        if (rc != exceeded_maximum_poas)
        {
            per_seq_status.clear(); // the reference clears it once the group has been admitted (cudapoa_batch.cuh:123)
            for (int32_t i = 0; i < n_st; ++i)
                per_seq_status.push_back(static_cast<StatusType>(st[i]));
        }
        return static_cast<StatusType>(rc);
    }
    int32_t get_total_poas() const override { return gwb200_poa_batch_total_poas(h_); }
    void generate_poa() override { check(gwb200_poa_batch_generate(h_)); }

    StatusType get_consensus(std::vector<std::string>& consensus, std::vector<std::vector<uint16_t>>& coverage,
                             std::vector<StatusType>& output_status) override
    {
        const int64_t n  = get_total_poas();
        const int64_t mc = cfg_.max_consensus_size;
        std::vector<char> c(static_cast<size_t>(std::max<int64_t>(n, 1) * mc));
        std::vector<uint16_t> cov(c.size());
        std::vector<int32_t> len(std::max<int64_t>(n, 1)), st(std::max<int64_t>(n, 1));
        const int rc = check(gwb200_poa_batch_get_consensus(h_, c.data(), cov.data(), len.data(), st.data()));
        if (rc == output_type_unavailable)
            return output_type_unavailable;
        for (int64_t w = 0; w < n; ++w) // results are appended, like the reference (cudapoa_batch.cuh:229-255)
        {
            output_status.push_back(static_cast<StatusType>(st[w]));
            consensus.emplace_back(c.data() + w * mc, static_cast<size_t>(len[w]));
            coverage.emplace_back(cov.begin() + w * mc, cov.begin() + w * mc + len[w]);
        }
        return success;
    }
    StatusType get_msa(std::vector<std::vector<std::string>>& msa, std::vector<StatusType>& output_status) override
    {
        const int64_t n  = get_total_poas();
        const int64_t mc = cfg_.max_consensus_size;
        const int64_t ms = cfg_.max_sequences_per_poa;
        std::vector<char> m(static_cast<size_t>(std::max<int64_t>(n, 1) * ms * mc));
        std::vector<int32_t> rows(std::max<int64_t>(n, 1)), st(std::max<int64_t>(n, 1));
        const int rc = check(gwb200_poa_batch_get_msa(h_, m.data(), rows.data(), st.data()));
        if (rc == output_type_unavailable)
            return output_type_unavailable;
        for (int64_t w = 0; w < n; ++w)
        {
            msa.emplace_back();
            output_status.push_back(static_cast<StatusType>(st[w]));
            for (int32_t r = 0; r < rows[w]; ++r)
                msa.back().emplace_back(m.data() + (w * ms + r) * mc);
        }
        return success;
    }
    void get_graphs(std::vector<DirectedGraph>& graphs, std::vector<StatusType>& output_status) override
    {
        const int64_t n = get_total_poas();
        std::vector<int32_t> nc(std::max<int64_t>(n, 1)), ec(std::max<int64_t>(n, 1)), st(std::max<int64_t>(n, 1));
        check(gwb200_poa_batch_get_graphs(h_, nc.data(), ec.data(), st.data(), nullptr, nullptr, nullptr, nullptr));
        int64_t tn = 0, te = 0;
        for (int64_t w = 0; w < n; ++w)
        {
            tn += nc[w];
            te += ec[w];
        }
        std::vector<uint8_t> labels(std::max<int64_t>(tn, 1));
        std::vector<int32_t> src(std::max<int64_t>(te, 1)), dst(std::max<int64_t>(te, 1)), wt(std::max<int64_t>(te, 1));
        check(gwb200_poa_batch_get_graphs(h_, nc.data(), ec.data(), st.data(), labels.data(), src.data(), dst.data(), wt.data()));
        graphs.resize(n);
        int64_t no = 0, eo = 0;
        for (int64_t w = 0; w < n; ++w)
        {
            output_status.push_back(static_cast<StatusType>(st[w]));
            DirectedGraph& g = graphs[w];
            for (int32_t k = 0; k < nc[w]; ++k)
                g.set_node_label(k, std::string(1, static_cast<char>(labels[no + k])));
            for (int32_t e = 0; e < ec[w]; ++e)
                g.add_edge(src[eo + e], dst[eo + e], wt[eo + e]);
            no += nc[w];
            eo += ec[w];
        }
    }
    int32_t batch_id() const override { return gwb200_poa_batch_id(h_); }
    void reset() override { check(gwb200_poa_batch_reset(h_)); }

private:
    gwb200_poa_batch* h_ = nullptr;
    BatchConfig cfg_;
    DefaultDeviceAllocator allocator_; // default constructed (no pool) unless the batch was created with an allocator
    char* block_         = nullptr;
    int64_t block_bytes_ = 0;
};
} // namespace detail

/// create_batch with an allocator (batch.hpp:176-189): max_gpu_mem == -1 => the allocator's largest free block.
inline std::unique_ptr<Batch> create_batch(int32_t device_id, cudaStream_t stream, DefaultDeviceAllocator allocator, int64_t max_gpu_mem,
                                           int8_t output_mask, const BatchConfig& batch_size, int16_t gap_score, int16_t mismatch_score,
                                           int16_t match_score)
{
    if (max_gpu_mem < -1)
        throw std::invalid_argument("max_gpu_mem has to be either -1 (=all available GPU memory) or greater or equal than 0.");
    if (max_gpu_mem == -1)
        max_gpu_mem = allocator.get_size_of_largest_free_memory_block();
    return std::unique_ptr<Batch>(
        new detail::BatchB200(device_id, stream, allocator, max_gpu_mem, output_mask, batch_size, gap_score, mismatch_score, match_score));
}

/// create_batch (batch.hpp:191-204); note the reference's argument order: gap, mismatch, match.
inline std::unique_ptr<Batch> create_batch(int32_t device_id, cudaStream_t stream, int64_t max_gpu_mem, int8_t output_mask,
                                           const BatchConfig& batch_size, int16_t gap_score, int16_t mismatch_score, int16_t match_score)
{
    return std::unique_ptr<Batch>(new detail::BatchB200(device_id, stream, max_gpu_mem, output_mask, batch_size, gap_score, mismatch_score, match_score));
}

} // namespace cudapoa
} // namespace genomeworks
} // namespace claraparabricks
