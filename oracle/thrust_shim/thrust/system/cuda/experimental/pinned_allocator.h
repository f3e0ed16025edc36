// TEST INFRASTRUCTURE ONLY (oracle build). Stand-in for the Thrust header that CUDA 12.9 no
// longer ships; the reference includes it from
// common/base/include/claraparabricks/genomeworks/utils/pinned_host_vector.hpp:23.
// Only used when compiling the unmodified reference sources into oracle/_ref/.
#pragma once
#include <thrust/detail/config.h>
#include <cuda_runtime_api.h>
#include <cstddef>
#include <new>
THRUST_NAMESPACE_BEGIN
namespace system
{
namespace cuda
{
namespace experimental
{
template <typename T>
struct pinned_allocator
{
    using value_type      = T;
    using pointer         = T*;
    using const_pointer   = const T*;
    using reference       = T&;
    using const_reference = const T&;
    using size_type       = std::size_t;
    using difference_type = std::ptrdiff_t;
    template <typename U>
    struct rebind
    {
        using other = pinned_allocator<U>;
    };
    pinned_allocator() = default;
    template <typename U>
    pinned_allocator(const pinned_allocator<U>&)
    {
    }
    T* allocate(size_type n, const void* = nullptr)
    {
        T* p = nullptr;
        if (cudaMallocHost(reinterpret_cast<void**>(&p), n * sizeof(T)) != cudaSuccess)
            throw std::bad_alloc();
        return p;
    }
    void deallocate(T* p, size_type) { cudaFreeHost(p); }
    bool operator==(const pinned_allocator&) const { return true; }
    bool operator!=(const pinned_allocator&) const { return false; }
};
} // namespace experimental
} // namespace cuda
} // namespace system
THRUST_NAMESPACE_END
