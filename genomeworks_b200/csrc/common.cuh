// gw-b200: shared host helpers for the C-ABI translation units.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

namespace gwb200
{

// Stores `msg` as the calling thread's last error and returns `code` (a negative GWB200_E_* value).
int set_error(int code, const char* msg);
// Counts one kernel launch issued by this library (gwb200_kernel_launch_count()).
void count_launch(int64_t n = 1);

// Sets the current device for the lifetime of the object and restores the previous one
// (the role of scoped_device_switch, common/base/include/claraparabricks/genomeworks/utils/cudautils.hpp:227-261).
struct DeviceGuard
{
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
        else
            prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

} // namespace gwb200

// CUDA runtime failures are reported to the caller (the reference logs and aborts, common/base/src/cudautils.cpp:75-100).
#define GWB200_CUDA_TRY(expr)                                                                         \
    do                                                                                                \
    {                                                                                                 \
        cudaError_t gwb200_err__ = (expr);                                                            \
        if (gwb200_err__ != cudaSuccess)                                                              \
        {                                                                                             \
            char gwb200_buf__[512];                                                                   \
            snprintf(gwb200_buf__, sizeof(gwb200_buf__), "CUDA error %s at %s:%d (%s)",               \
                     cudaGetErrorString(gwb200_err__), __FILE__, __LINE__, #expr);                    \
            cudaGetLastError();                                                                       \
            return gwb200::set_error(-3, gwb200_buf__);                                               \
        }                                                                                             \
    } while (0)
