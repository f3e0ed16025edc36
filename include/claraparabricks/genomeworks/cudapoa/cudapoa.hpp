// gw-b200: cudapoa package header -- same names and values as the reference's
// cudapoa/include/claraparabricks/genomeworks/cudapoa/cudapoa.hpp:34-85, implemented over the C ABI (include/gwb200.h).
#pragma once

#include "../../../gwb200.h"

#include <stdexcept>
#include <string>

namespace claraparabricks
{
namespace genomeworks
{
namespace cudapoa
{

/// CUDA POA error type (values cross the device boundary, cudapoa.hpp:34-49)
enum StatusType
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

/// Banding mode of the Needleman-Wunsch stage (cudapoa.hpp:62-69)
enum BandMode
{
    full_band = 0,
    static_band,
    adaptive_band,
    static_band_traceback,
    adaptive_band_traceback
};

/// Output selection bit mask (cudapoa.hpp:75-79)
enum OutputType
{
    consensus = 0x1,
    msa       = 0x1 << 1
};

/// Initialize CUDA POA context.
inline StatusType Init() { return static_cast<StatusType>(gwb200_poa_init()); }

/// Message + hint for a status; throws std::runtime_error for an unknown one (cudapoa/src/cudapoa.cpp:37-92).
inline void decode_error(StatusType error_type, std::string& error_message, std::string& error_hint)
{
    char m[512], h[512];
    if (gwb200_poa_decode_error(static_cast<int32_t>(error_type), m, sizeof(m), h, sizeof(h)) < 0)
        throw std::runtime_error(gwb200_last_error());
    error_message = m;
    error_hint    = h;
}

namespace detail
{
/// Maps the C ABI's negative codes to the exceptions the reference throws.
inline int check(int rc)
{
    if (rc >= 0)
        return rc;
    const std::string msg = gwb200_last_error();
    if (rc == GWB200_E_INVALID_ARGUMENT)
        throw std::invalid_argument(msg);
    throw std::runtime_error(msg);
}
} // namespace detail

} // namespace cudapoa
} // namespace genomeworks
} // namespace claraparabricks
