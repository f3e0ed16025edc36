// gw-b200: DefaultDeviceAllocator as it appears in the kept API (create_batch / create_aligner overloads,
// Aligner::get_device_allocator). The reference's allocator (common/base/include/.../utils/allocator.hpp:322-358) is a
// caching pool whose only property the hot path consumes is "how much device memory may this object use"
// (get_size_of_largest_free_memory_block, allocate_block.hpp:63-64, aligner.cpp:103-117). This engine carves its own
// arenas, so the allocator here carries exactly that budget.
#pragma once

#include "cudautils.hpp" // CudaStream, make_cuda_stream, GW_CU_CHECK_ERR come along with the allocator in the reference too

#include <cstdint>
#include <cuda_runtime_api.h>

namespace claraparabricks
{
namespace genomeworks
{

class DefaultDeviceAllocator
{
public:
    explicit DefaultDeviceAllocator(int64_t max_bytes = -1, cudaStream_t stream = nullptr)
        : max_bytes_(max_bytes)
        , stream_(stream)
    {
    }
    /// Bytes this allocator may hand out; -1 = whatever the device has free.
    int64_t get_size_of_largest_free_memory_block() const
    {
        if (max_bytes_ >= 0)
            return max_bytes_;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
            return 0;
        return static_cast<int64_t>(free_b);
    }
    cudaStream_t default_stream() const { return stream_; }

private:
    int64_t max_bytes_;
    cudaStream_t stream_;
};

/// allocator.hpp:347-358
inline DefaultDeviceAllocator create_default_device_allocator(int64_t max_caching_size = 2ll * 1024 * 1024 * 1024, cudaStream_t stream = nullptr)
{
    return DefaultDeviceAllocator(max_caching_size, stream);
}
/// allocator.hpp:331-334
inline int64_t get_size_of_largest_free_memory_block(const DefaultDeviceAllocator& a) { return a.get_size_of_largest_free_memory_block(); }

} // namespace genomeworks
} // namespace claraparabricks
