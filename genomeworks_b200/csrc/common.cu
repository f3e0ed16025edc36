// gw-b200: shared host helpers (error text, launch counter, version).
#include "../../include/gwb200.h"
#include "common.cuh"

#include <atomic>
#include <string>

namespace gwb200
{
namespace
{
thread_local std::string g_last_error;
std::atomic<int64_t> g_launches{0};
} // namespace

int set_error(int code, const char* msg)
{
    g_last_error = msg ? msg : "";
    return code;
}

void count_launch(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

const char* last_error_cstr() { return g_last_error.c_str(); }
int64_t launches() { return g_launches.load(std::memory_order_relaxed); }

} // namespace gwb200

extern "C" {
const char* gwb200_last_error(void) { return gwb200::last_error_cstr(); }
const char* gwb200_version(void) { return "gw-b200 0.1.0 (sm_100a)"; }
int64_t gwb200_kernel_launch_count(void) { return gwb200::launches(); }
}
