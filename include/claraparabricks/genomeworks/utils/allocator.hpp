// gw-b200: the device allocator that is part of the kept API (create_batch / create_aligner overloads,
// Aligner::get_device_allocator). Member set and semantics of the reference's
// common/base/include/claraparabricks/genomeworks/utils/allocator.hpp:208-358 (CachingDeviceAllocator<T, MemoryResource>,
// DefaultDeviceAllocator, create_default_device_allocator, get_size_of_largest_free_memory_block) and
// utils/device_preallocated_allocator.cuh:48-300 (DevicePreallocatedAllocator: ONE cudaMalloc of the whole pool up front,
// aligned blocks handed out first-fit under a mutex, a freed block waits for the streams it was associated with).
// Copies of an allocator share one pool: two Batch / Aligner objects created from the same allocator carve disjoint blocks
// out of it, and get_size_of_largest_free_memory_block() reflects what the others took.
// Implemented from scratch: the free space is an address-ordered map {offset -> size} with eager coalescing.
#pragma once

#include "cudautils.hpp" // CudaStream, make_cuda_stream, GW_CU_CHECK_ERR come along with the allocator in the reference too

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime_api.h>
#include <exception>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace claraparabricks
{
namespace genomeworks
{

/// Thrown by allocate() when the pool (or the device) cannot serve the request (utils/exceptions.hpp in the reference).
class device_memory_allocation_exception : public std::exception
{
public:
    const char* what() const noexcept override { return "Could not allocate device memory!"; }
};

namespace details
{

/// One device buffer allocated up front, carved into 256-byte aligned blocks.
class DevicePreallocatedAllocator
{
public:
    explicit DevicePreallocatedAllocator(size_t buffer_size)
        : size_(buffer_size / kAlign * kAlign)
    {
        void* p = nullptr;
        if (size_ == 0 || cudaMalloc(&p, size_) != cudaSuccess)
        {
            cudaGetLastError();
            throw device_memory_allocation_exception();
        }
        base_ = static_cast<char*>(p);
        free_.emplace(size_t(0), size_);
    }
    DevicePreallocatedAllocator(const DevicePreallocatedAllocator&) = delete;
    DevicePreallocatedAllocator& operator=(const DevicePreallocatedAllocator&) = delete;
    ~DevicePreallocatedAllocator()
    {
        if (base_)
            cudaFree(base_);
    }

    /// First fit; cudaErrorMemoryAllocation when no free block is large enough. `streams`: work that may still use the block
    /// when it is freed (DeviceFree waits for them).
    cudaError_t DeviceAllocate(void** ptr, size_t bytes, const std::vector<cudaStream_t>& streams)
    {
        const size_t need = (bytes + kAlign - 1) / kAlign * kAlign;
        std::lock_guard<std::mutex> lock(mutex_);
        for (auto it = free_.begin(); it != free_.end(); ++it)
        {
            if (it->second >= need && need > 0)
            {
                const size_t off  = it->first;
                const size_t rest = it->second - need;
                free_.erase(it);
                if (rest > 0)
                    free_.emplace(off + need, rest);
                used_.emplace(off, Used{need, streams});
                *ptr = base_ + off;
                return cudaSuccess;
            }
        }
        *ptr = nullptr;
        return cudaErrorMemoryAllocation;
    }

    cudaError_t DeviceFree(void* ptr)
    {
        if (ptr == nullptr)
            return cudaSuccess;
        std::lock_guard<std::mutex> lock(mutex_);
        const size_t off = static_cast<size_t>(static_cast<char*>(ptr) - base_);
        auto it          = used_.find(off);
        if (it == used_.end())
            return cudaErrorInvalidValue;
        cudaError_t status = cudaSuccess;
        for (cudaStream_t s : it->second.streams)
        {
            const cudaError_t e = cudaStreamSynchronize(s);
            if (e != cudaSuccess)
                status = e;
        }
        size_t begin = off, size = it->second.size;
        used_.erase(it);
        // merge with the free neighbours
        auto next = free_.lower_bound(begin);
        if (next != free_.end() && next->first == begin + size)
        {
            size += next->second;
            next = free_.erase(next);
        }
        if (next != free_.begin())
        {
            auto prev = std::prev(next);
            if (prev->first + prev->second == begin)
            {
                begin = prev->first;
                size += prev->second;
                free_.erase(prev);
            }
        }
        free_.emplace(begin, size);
        return status;
    }

    int64_t get_size_of_largest_free_memory_block() const
    {
        std::lock_guard<std::mutex> lock(mutex_);
        size_t best = 0;
        for (const auto& kv : free_)
            best = kv.second > best ? kv.second : best;
        return static_cast<int64_t>(best);
    }

private:
    static constexpr size_t kAlign = 256;
    struct Used
    {
        size_t size;
        std::vector<cudaStream_t> streams;
    };
    size_t size_ = 0;
    char* base_  = nullptr;
    std::map<size_t, size_t> free_; // offset -> size, address ordered
    std::map<size_t, Used> used_;
    mutable std::mutex mutex_;
};

} // namespace details

/// Plain cudaMalloc / cudaFree allocator (allocator.hpp:78-170); what DefaultDeviceAllocator is when caching is disabled.
template <typename T>
class CudaMallocAllocator
{
public:
    using value_type = T;
    using pointer    = T*;
    CudaMallocAllocator() = default;
    template <typename U>
    CudaMallocAllocator(const CudaMallocAllocator<U>&)
    {
    }
    pointer allocate(std::size_t n, const std::vector<cudaStream_t>& = {})
    {
        void* p = nullptr;
        if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess)
        {
            cudaGetLastError();
            throw device_memory_allocation_exception();
        }
        return static_cast<pointer>(p);
    }
    void deallocate(pointer p, std::size_t) { cudaFree(p); }
    int64_t get_size_of_largest_free_memory_block() const { return cudautils::find_largest_contiguous_device_memory_section(); }
};

/// Allocator over a shared memory resource (allocator.hpp:208-319). Default construction yields a dummy that cannot allocate.
template <typename T, typename MemoryResource>
class CachingDeviceAllocator
{
public:
    using value_type = T;
    using pointer    = T*;

    CachingDeviceAllocator() = default;

    /// \param max_cached_bytes size of the pool this allocator (and every copy of it) hands blocks out of
    /// \param default_stream stream a block is associated with when allocate() names none
    explicit CachingDeviceAllocator(size_t max_cached_bytes, cudaStream_t default_stream = 0)
        : memory_resource_(std::make_shared<MemoryResource>(max_cached_bytes))
        , default_stream_(default_stream)
    {
    }

    CachingDeviceAllocator(const CachingDeviceAllocator&) = default;
    CachingDeviceAllocator(CachingDeviceAllocator&&)      = default;
    CachingDeviceAllocator& operator=(const CachingDeviceAllocator&) = default;
    CachingDeviceAllocator& operator=(CachingDeviceAllocator&&) = default;

    /// rebinding copy: same pool, other value type
    template <typename U>
    CachingDeviceAllocator(const CachingDeviceAllocator<U, MemoryResource>& rhs)
        : memory_resource_(rhs.memory_resource())
        , default_stream_(rhs.default_stream())
    {
    }

    pointer allocate(std::size_t n, const std::vector<cudaStream_t>& streams = {})
    {
        if (!memory_resource_)
            std::abort(); // default-constructed allocator: the reference logs and aborts (allocator.hpp:267-272)
        void* p = nullptr;
        const cudaError_t err =
            memory_resource_->DeviceAllocate(&p, n * sizeof(T), streams.empty() ? std::vector<cudaStream_t>(1, default_stream_) : streams);
        if (err == cudaErrorMemoryAllocation)
            throw device_memory_allocation_exception();
        GW_CU_CHECK_ERR(err);
        return static_cast<pointer>(p);
    }

    void deallocate(pointer p, std::size_t)
    {
        if (!memory_resource_)
            std::abort();
        memory_resource_->DeviceFree(p);
    }

    int64_t get_size_of_largest_free_memory_block() const { return memory_resource_ ? memory_resource_->get_size_of_largest_free_memory_block() : 0; }
    std::shared_ptr<MemoryResource> memory_resource() const { return memory_resource_; }
    cudaStream_t default_stream() const { return default_stream_; }

private:
    std::shared_ptr<MemoryResource> memory_resource_;
    cudaStream_t default_stream_ = 0;
};

/// GW_ENABLE_CACHING_ALLOCATOR is the reference's default build (CMakeLists.txt:41): the pool allocator.
using DefaultDeviceAllocator = CachingDeviceAllocator<char, details::DevicePreallocatedAllocator>;

/// allocator.hpp:331-334
inline int64_t get_size_of_largest_free_memory_block(const DefaultDeviceAllocator& allocator)
{
    return allocator.get_size_of_largest_free_memory_block();
}

/// allocator.hpp:347-358
inline DefaultDeviceAllocator create_default_device_allocator(std::size_t max_caching_size = 2ull * 1024 * 1024 * 1024, cudaStream_t default_stream = 0)
{
    return DefaultDeviceAllocator(max_caching_size, default_stream);
}

} // namespace genomeworks
} // namespace claraparabricks
