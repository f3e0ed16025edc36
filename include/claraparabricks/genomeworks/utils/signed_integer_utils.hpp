// gw-b200: the small signed-size helpers that callers of the kept API use
// (common/base/include/claraparabricks/genomeworks/utils/signed_integer_utils.hpp:30-58): get_size() returns a container's size as a
// signed integer, throw_on_negative() validates an argument.
#pragma once

#include <cassert>
#include <limits>
#include <stdexcept>
#include <type_traits>

namespace claraparabricks
{
namespace genomeworks
{

/// Size of a container as the signed counterpart of its size_type.
template <class Container>
typename std::make_signed<typename Container::size_type>::type get_size(const Container& c)
{
    typedef typename std::make_signed<typename Container::size_type>::type Signed;
    assert(c.size() <= static_cast<typename Container::size_type>(std::numeric_limits<Signed>::max()));
    return static_cast<Signed>(c.size());
}

/// Size of a container as the requested integer type.
template <class Integer, class Container>
Integer get_size(const Container& c)
{
    assert(c.size() <= static_cast<typename Container::size_type>(std::numeric_limits<Integer>::max()));
    return static_cast<Integer>(c.size());
}

/// Returns x, or throws std::invalid_argument(message) if x is negative.
template <class T>
T throw_on_negative(T x, const char* message)
{
    static_assert(std::is_arithmetic<T>::value, "throw_on_negative expects an arithmetic type.");
    if (x < T(0))
        throw std::invalid_argument(message);
    return x;
}

} // namespace genomeworks
} // namespace claraparabricks
