// gw-b200: cudaaligner package header -- same names and values as the reference's
// cudaaligner/include/claraparabricks/genomeworks/cudaaligner/cudaaligner.hpp:34-68, over the C ABI (include/gwb200.h).
#pragma once

#include "../../../gwb200.h"

#include <cstdint>
#include <stdexcept>
#include <string>

namespace claraparabricks
{
namespace genomeworks
{
namespace cudaaligner
{

enum StatusType
{
    success = 0,
    uninitialized,
    exceeded_max_alignments,
    exceeded_max_length,
    exceeded_max_alignment_difference,
    generic_error
};

enum AlignmentType
{
    global_alignment = 0,
    unset
};

/// The bytes in device results (aligner.hpp:62-72)
enum AlignmentState : int8_t
{
    match = 0,
    mismatch,
    insertion, // absent in query, present in target
    deletion   // present in query, absent in target
};

enum CigarFormat
{
    basic = 0, // symbols I, D, M
    extended   // symbols I, D, X, =
};

inline StatusType Init() { return static_cast<StatusType>(gwb200_aligner_init()); }

namespace detail
{
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

} // namespace cudaaligner
} // namespace genomeworks
} // namespace claraparabricks
