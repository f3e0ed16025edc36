// gw-b200: cudaaligner::Aligner / FixedBandAligner / create_aligner with the reference's signatures
// (cudaaligner/include/claraparabricks/genomeworks/cudaaligner/aligner.hpp:41-219), as a header-only adapter over the
// C ABI of libgwb200.so. Both factory families are served by the banded Myers engine: the FixedBand overloads directly;
// the deprecated (max_query, max_target, max_alignments) overloads with a bandwidth covering the whole query (full Myers,
// exact; the reference uses AlignerGlobalHirschbergMyers there, cudaaligner/src/aligner.cpp:31-74).
#pragma once

#include "alignment.hpp"
#include "../utils/allocator.hpp"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <memory>
#include <string>
#include <vector>

namespace claraparabricks
{
namespace genomeworks
{
namespace cudaaligner
{

/// Device-resident results (aligner.hpp:62-72); pointers are borrowed until reset() / destruction.
struct DeviceAlignmentsPtrs
{
    const int8_t* cigar_operations;
    const int32_t* cigar_runlengths;
    const int32_t* cigar_offsets;
    const uint32_t* metadata; ///< bit 31: is_optimal, bits 26-0: index of the alignment in insertion order
    int64_t total_length;
    int32_t n_alignments;
    static constexpr uint32_t index_mask = (1u << 27) - 1;
};

class Aligner
{
public:
    virtual ~Aligner()                   = default;
    virtual StatusType align_all()       = 0;
    virtual StatusType sync_alignments() = 0;
    virtual StatusType add_alignment(const char* query, int32_t query_length, const char* target, int32_t target_length,
                                     bool reverse_complement_query = false, bool reverse_complement_target = false) = 0;
    virtual const std::vector<std::shared_ptr<Alignment>>& get_alignments() const                                      = 0;
    virtual DeviceAlignmentsPtrs get_alignments_device() const                                                         = 0;
    virtual void reset()                                                                                               = 0;
    virtual void free_temporary_device_buffers()                                                                       = 0;
    virtual int32_t num_alignments() const                                                                             = 0;
    virtual cudaStream_t get_stream() const                                                                            = 0;
    virtual int32_t get_device() const                                                                                 = 0;
    virtual DefaultDeviceAllocator get_device_allocator() const                                                        = 0;
};

class FixedBandAligner : public Aligner
{
public:
    virtual void reset_max_bandwidth(int32_t max_bandwidth) = 0;
    using Aligner::add_alignment;
    virtual StatusType add_alignment(int32_t max_bandwidth, const char* query, int32_t query_length, const char* target, int32_t target_length,
                                     bool reverse_complement_query = false, bool reverse_complement_target = false) = 0;
};

namespace detail
{
class AlignerB200 : public FixedBandAligner
{
public:
    AlignerB200(int32_t max_bandwidth, cudaStream_t stream, int32_t device_id, int64_t max_device_memory, int32_t max_query = -1,
                int32_t max_target = -1, int32_t max_alignments = -1)
        : stream_(stream)
        , device_(device_id)
        , mem_(max_device_memory)
        , max_query_(max_query)
        , max_target_(max_target)
        , max_alignments_(max_alignments)
    {
        check(gwb200_aligner_create(&h_, max_bandwidth, stream, device_id, max_device_memory));
    }
    /// Every device buffer of the aligner is a block of the caller's allocator (aligner.hpp:183,208; cudaaligner/src/aligner.cpp:76-124).
    AlignerB200(int32_t max_bandwidth, cudaStream_t stream, int32_t device_id, DefaultDeviceAllocator allocator, int64_t max_device_memory,
                int32_t max_query = -1, int32_t max_target = -1, int32_t max_alignments = -1)
        : stream_(stream)
        , device_(device_id)
        , mem_(max_device_memory)
        , max_query_(max_query)
        , max_target_(max_target)
        , max_alignments_(max_alignments)
        , allocator_(allocator)
    {
        check(gwb200_aligner_create_with_allocator(&h_, max_bandwidth, stream, device_id, max_device_memory, &AlignerB200::pool_alloc,
                                                   &AlignerB200::pool_free, this));
    }
    ~AlignerB200() override { gwb200_aligner_destroy(h_); }
    AlignerB200(const AlignerB200&) = delete;
    AlignerB200& operator=(const AlignerB200&) = delete;

    StatusType add_alignment(const char* query, int32_t query_length, const char* target, int32_t target_length, bool rc_q = false,
                             bool rc_t = false) override
    {
        return add_alignment(GWB200_ALN_DEFAULT_BANDWIDTH, query, query_length, target, target_length, rc_q, rc_t);
    }
    StatusType add_alignment(int32_t max_bandwidth, const char* query, int32_t query_length, const char* target, int32_t target_length,
                             bool rc_q = false, bool rc_t = false) override
    {
        if (max_query_ >= 0)
        {
            // the deprecated fixed-size factory: AlignerGlobal::add_alignment (cudaaligner/src/aligner_global.cpp:78-141) -- the
            // Alignment object exists from here on (get_alignments() lists it before sync_alignments() fills it in)
            if (query_length < 0 || target_length < 0)
                return StatusType::generic_error;
            if (static_cast<int32_t>(alignments_.size()) >= max_alignments_)
                return StatusType::exceeded_max_alignments;
            if (query_length > max_query_ || target_length > max_target_)
                return StatusType::exceeded_max_length;
        }
        const int rc = check(gwb200_aligner_add_alignment(h_, max_bandwidth, query, query_length, target, target_length, rc_q ? 1 : 0, rc_t ? 1 : 0));
        if (rc == success)
        {
            // the Alignment carries the sequences as they were aligned, i.e. after the reverse complement the flags ask for
            // (the reference builds it from its staging copy, aligner_global_myers_banded.cpp:226-227,413-418)
            if (max_query_ >= 0)
            {
                auto al = std::make_shared<AlignmentB200>(staged(query, query_length, rc_q), staged(target, target_length, rc_t));
                al->set_type(AlignmentType::global_alignment);
                alignments_.push_back(std::move(al));
            }
            else
            {
                pending_.emplace_back(staged(query, query_length, rc_q), staged(target, target_length, rc_t));
            }
        }
        return static_cast<StatusType>(rc);
    }
    static std::string staged(const char* seq, int32_t length, bool reverse_complement)
    {
        std::string out(seq, seq + length);
        if (reverse_complement)
        {
            // genomeutils::reverse_complement (utils/genomeutils.hpp:144-154): A -> T, C -> G, T -> A, G -> C by (c >> 1) & 3
            static const char lookup[4] = {'T', 'G', 'A', 'C'};
            for (int32_t pos = 0; pos < length; ++pos)
                out[pos] = lookup[(static_cast<unsigned char>(seq[length - 1 - pos]) >> 1) & 0x3];
        }
        return out;
    }
    StatusType align_all() override { return static_cast<StatusType>(check(gwb200_aligner_align_all(h_))); }
    StatusType sync_alignments() override
    {
        const int rc = check(gwb200_aligner_sync_alignments(h_));
        auto fetch = [&](int32_t i, AlignmentB200& al) {
            int32_t st = 0, opt = 0, n = 0;
            check(gwb200_aligner_result_info(h_, i, &st, &opt, &n));
            std::vector<int8_t> a(std::max(n, 1));
            std::vector<int32_t> r(std::max(n, 1));
            check(gwb200_aligner_result_runs(h_, i, a.data(), r.data()));
            a.resize(n);
            r.resize(n);
            if (st == success)
                al.set(StatusType::success, opt != 0, std::move(a), std::move(r));
        };
        if (max_query_ >= 0)
        {
            // AlignerGlobal::sync_alignments (aligner_global.cpp:162-190): results are filled into the existing objects
            const int32_t n_results = gwb200_aligner_num_results(h_);
            for (int32_t i = first_unsynced_; i < static_cast<int32_t>(alignments_.size()) && i - first_unsynced_ < n_results; ++i)
            {
                AlignmentB200* al = static_cast<AlignmentB200*>(alignments_[i].get());
                fetch(i - first_unsynced_, *al);
                al->expand();
            }
            if (n_results > 0)
                first_unsynced_ = static_cast<int32_t>(alignments_.size());
            return static_cast<StatusType>(rc);
        }
        if (gwb200_aligner_num_results(h_) == 0 && pending_.empty())
            return static_cast<StatusType>(rc); // nothing was aligned since the last sync: the previous alignments stay
        alignments_.clear();
        for (size_t i = 0; i < pending_.size(); ++i)
        {
            auto al = std::make_shared<AlignmentB200>(std::move(pending_[i].first), std::move(pending_[i].second));
            fetch(static_cast<int32_t>(i), *al);
            alignments_.push_back(std::move(al));
        }
        pending_.clear();
        return static_cast<StatusType>(rc);
    }
    const std::vector<std::shared_ptr<Alignment>>& get_alignments() const override { return alignments_; }
    DeviceAlignmentsPtrs get_alignments_device() const override
    {
        DeviceAlignmentsPtrs p{};
        check(gwb200_aligner_get_alignments_device(h_, &p.cigar_operations, &p.cigar_runlengths, &p.cigar_offsets, &p.metadata, &p.total_length,
                                                   &p.n_alignments));
        return p;
    }
    void reset() override
    {
        check(gwb200_aligner_reset(h_));
        pending_.clear();
        alignments_.clear();
        first_unsynced_ = 0;
    }
    void reset_max_bandwidth(int32_t max_bandwidth) override
    {
        check(gwb200_aligner_reset_max_bandwidth(h_, max_bandwidth));
        pending_.clear();
        alignments_.clear();
    }
    void free_temporary_device_buffers() override { check(gwb200_aligner_free_temporary_device_buffers(h_)); }
    int32_t num_alignments() const override
    {
        return max_query_ >= 0 ? static_cast<int32_t>(alignments_.size()) : gwb200_aligner_num_alignments(h_);
    }
    cudaStream_t get_stream() const override { return stream_; }
    int32_t get_device() const override { return device_; }
    /// The allocator the aligner was created with; an aligner created without one allocates with cudaMalloc directly and
    /// returns a default-constructed allocator (no pool).
    DefaultDeviceAllocator get_device_allocator() const override { return allocator_; }

private:
    static void* pool_alloc(void* user, int64_t bytes)
    {
        AlignerB200* self = static_cast<AlignerB200*>(user);
        try
        {
            return self->allocator_.allocate(static_cast<std::size_t>(bytes), {self->stream_});
        }
        catch (const device_memory_allocation_exception&)
        {
            return nullptr;
        }
    }
    static void pool_free(void* user, void* ptr, int64_t bytes)
    {
        static_cast<AlignerB200*>(user)->allocator_.deallocate(static_cast<char*>(ptr), static_cast<std::size_t>(bytes));
    }

    gwb200_aligner* h_ = nullptr;
    cudaStream_t stream_;
    int32_t device_;
    int64_t mem_;
    int32_t max_query_, max_target_, max_alignments_;
    DefaultDeviceAllocator allocator_;
    int32_t first_unsynced_ = 0; // deprecated factory: first alignment that has no result yet
    std::vector<std::pair<std::string, std::string>> pending_;
    std::vector<std::shared_ptr<Alignment>> alignments_;
};

/// AlignerGlobal (cudaaligner/src/aligner_global.hpp:40-127) on this engine: fixed maximum lengths / count, Hirschberg-Myers
/// (what the reference's deprecated factory builds, cudaaligner/src/aligner.cpp:31-74) or the unbanded Myers aligner.
class AlignerGlobalB200 : public Aligner
{
public:
    AlignerGlobalB200(int32_t algorithm, int32_t max_query_length, int32_t max_target_length, int32_t max_alignments, cudaStream_t stream,
                      int32_t device_id)
        : stream_(stream)
        , device_(device_id)
    {
        check(gwb200_global_aligner_create(&h_, algorithm, max_query_length, max_target_length, max_alignments, stream, device_id, nullptr,
                                           nullptr, nullptr));
    }
    AlignerGlobalB200(int32_t algorithm, int32_t max_query_length, int32_t max_target_length, int32_t max_alignments,
                      DefaultDeviceAllocator allocator, cudaStream_t stream, int32_t device_id)
        : stream_(stream)
        , device_(device_id)
        , allocator_(allocator)
    {
        check(gwb200_global_aligner_create(&h_, algorithm, max_query_length, max_target_length, max_alignments, stream, device_id,
                                           &AlignerGlobalB200::pool_alloc, &AlignerGlobalB200::pool_free, this));
    }
    ~AlignerGlobalB200() override { gwb200_global_aligner_destroy(h_); }
    AlignerGlobalB200(const AlignerGlobalB200&) = delete;
    AlignerGlobalB200& operator=(const AlignerGlobalB200&) = delete;

    StatusType add_alignment(const char* query, int32_t query_length, const char* target, int32_t target_length, bool rc_q = false,
                             bool rc_t = false) override
    {
        const int rc = check(gwb200_global_aligner_add_alignment(h_, query, query_length, target, target_length, rc_q ? 1 : 0, rc_t ? 1 : 0));
        if (rc == success)
        {
            // the Alignment exists from here on (aligner_global.cpp:131-138) and carries the sequences as they are aligned
            auto al = std::make_shared<AlignmentB200>(AlignerB200::staged(query, query_length, rc_q), AlignerB200::staged(target, target_length, rc_t));
            al->set_type(AlignmentType::global_alignment);
            alignments_.push_back(std::move(al));
        }
        return static_cast<StatusType>(rc);
    }
    StatusType align_all() override { return static_cast<StatusType>(check(gwb200_global_aligner_align_all(h_))); }
    StatusType sync_alignments() override
    {
        const int rc = check(gwb200_global_aligner_sync_alignments(h_));
        for (int32_t i = 0; i < static_cast<int32_t>(alignments_.size()); ++i)
        {
            int32_t have = 0, opt = 0, n = 0;
            check(gwb200_global_aligner_result_info(h_, i, &have, &opt, &n));
            if (!have)
                continue; // failed alignment: the object stays as add_alignment made it (aligner_global.cpp:180)
            std::vector<int8_t> st(std::max(n, 1));
            check(gwb200_global_aligner_result_states(h_, i, st.data()));
            st.resize(n);
            std::vector<int8_t> actions;
            std::vector<int32_t> runs;
            for (int32_t k = 0; k < n; ++k)
            {
                if (actions.empty() || actions.back() != st[k])
                {
                    actions.push_back(st[k]);
                    runs.push_back(0);
                }
                ++runs.back();
            }
            AlignmentB200* al = static_cast<AlignmentB200*>(alignments_[i].get());
            al->set(StatusType::success, opt != 0, std::move(actions), std::move(runs));
            al->expand();
        }
        return static_cast<StatusType>(rc);
    }
    const std::vector<std::shared_ptr<Alignment>>& get_alignments() const override { return alignments_; }
    /// AlignerGlobal::get_alignments_device (aligner_global.hpp:78-83): not available for this aligner family
    DeviceAlignmentsPtrs get_alignments_device() const override { throw std::runtime_error("get_alignments_device() not implemented for this aligner"); }
    void reset() override
    {
        check(gwb200_global_aligner_reset(h_));
        alignments_.clear();
    }
    void free_temporary_device_buffers() override {}
    int32_t num_alignments() const override { return static_cast<int32_t>(alignments_.size()); }
    cudaStream_t get_stream() const override { return stream_; }
    int32_t get_device() const override { return device_; }
    DefaultDeviceAllocator get_device_allocator() const override { return allocator_; }

private:
    static void* pool_alloc(void* user, int64_t bytes)
    {
        AlignerGlobalB200* self = static_cast<AlignerGlobalB200*>(user);
        try
        {
            return self->allocator_.allocate(static_cast<std::size_t>(bytes), {self->stream_});
        }
        catch (const device_memory_allocation_exception&)
        {
            return nullptr;
        }
    }
    static void pool_free(void* user, void* ptr, int64_t bytes)
    {
        static_cast<AlignerGlobalB200*>(user)->allocator_.deallocate(static_cast<char*>(ptr), static_cast<std::size_t>(bytes));
    }
    gwb200_global_aligner* h_ = nullptr;
    cudaStream_t stream_;
    int32_t device_;
    DefaultDeviceAllocator allocator_;
    std::vector<std::shared_ptr<Alignment>> alignments_;
};

inline int32_t covering_bandwidth(int32_t max_query_length, int32_t max_target_length)
{
    int32_t bw = std::max(std::max(max_query_length, max_target_length), 2);
    if (bw % 32 == 1)
        ++bw;
    return bw;
}
} // namespace detail

/// Deprecated factory (aligner.hpp:183): fixed maximum lengths / count -> AlignerGlobalHirschbergMyers (cudaaligner/src/aligner.cpp:31-74).
inline std::unique_ptr<Aligner> create_aligner(int32_t max_query_length, int32_t max_target_length, int32_t max_alignments, AlignmentType type,
                                               DefaultDeviceAllocator allocator, cudaStream_t stream, int32_t device_id)
{
    if (type != AlignmentType::global_alignment)
        throw std::runtime_error("Aligner for specified type not implemented yet.");
    return std::unique_ptr<Aligner>(new detail::AlignerGlobalB200(GWB200_GLOBAL_HIRSCHBERG_MYERS, max_query_length, max_target_length, max_alignments,
                                                                  allocator, stream, device_id));
}
/// Deprecated factory (aligner.hpp:196).
inline std::unique_ptr<Aligner> create_aligner(int32_t max_query_length, int32_t max_target_length, int32_t max_alignments, AlignmentType type,
                                               cudaStream_t stream, int32_t device_id, int64_t max_device_memory_allocator_caching_size = -1)
{
    if (type != AlignmentType::global_alignment)
        throw std::runtime_error("Aligner for specified type not implemented yet.");
    if (max_device_memory_allocator_caching_size < -1)
        throw std::invalid_argument("max_device_memory_allocator_caching_size has to be either -1 (=all available GPU memory) or greater or equal than 0.");
    return std::unique_ptr<Aligner>(new detail::AlignerGlobalB200(GWB200_GLOBAL_HIRSCHBERG_MYERS, max_query_length, max_target_length, max_alignments,
                                                                  stream, device_id));
}
/// FixedBand factory with allocator (aligner.hpp:208): max_device_memory == -1 => the allocator's largest free block.
inline std::unique_ptr<FixedBandAligner> create_aligner(AlignmentType type, int32_t max_bandwidth, cudaStream_t stream, int32_t device_id,
                                                        DefaultDeviceAllocator allocator, int64_t max_device_memory)
{
    if (type != AlignmentType::global_alignment)
        throw std::runtime_error("Aligner for specified type not implemented yet.");
    if (max_device_memory < -1)
        throw std::invalid_argument("max_device_memory has to be either -1 (=all available GPU memory) or greater or equal than 0.");
    if (max_device_memory == -1)
        max_device_memory = allocator.get_size_of_largest_free_memory_block();
    return std::unique_ptr<FixedBandAligner>(new detail::AlignerB200(max_bandwidth, stream, device_id, allocator, max_device_memory));
}
/// FixedBand factory (aligner.hpp:219).
inline std::unique_ptr<FixedBandAligner> create_aligner(AlignmentType type, int32_t max_bandwidth, cudaStream_t stream, int32_t device_id,
                                                        int64_t max_device_memory = -1)
{
    if (type != AlignmentType::global_alignment)
        throw std::runtime_error("Aligner for specified type not implemented yet.");
    return std::unique_ptr<FixedBandAligner>(new detail::AlignerB200(max_bandwidth, stream, device_id, max_device_memory));
}

} // namespace cudaaligner
} // namespace genomeworks
} // namespace claraparabricks
