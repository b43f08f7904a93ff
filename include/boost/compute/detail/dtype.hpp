// Maps C++ scalar types and functor tags onto the integer codes of the C ABI (include/compute_b200.h).
// This replaces type_name<T>() + "-DT=..." JIT options of the reference (type_traits/type_name.hpp:95-99,
// algorithm/detail/radix_sort.hpp:289-311): kernels are compiled ahead of time, the header only picks one.
#ifndef B200_BOOST_COMPUTE_DETAIL_DTYPE_HPP
#define B200_BOOST_COMPUTE_DETAIL_DTYPE_HPP

#include <type_traits>

#include <compute_b200.h>

namespace boost {
namespace compute {
namespace detail {

template<class T, class Enable = void>
struct dtype_of
{
    static const bool supported = false;
};

template<class T>
struct dtype_of<T, typename std::enable_if<std::is_floating_point<T>::value && (sizeof(T) == 4 || sizeof(T) == 8)>::type>
{
    static const bool supported = true;
    static const int value = sizeof(T) == 4 ? BCB_FLOAT : BCB_DOUBLE;
};

template<class T>
struct dtype_of<T, typename std::enable_if<std::is_integral<T>::value && !std::is_same<T, bool>::value>::type>
{
    static const bool supported = true;
    static const int value =
        sizeof(T) == 1 ? (std::is_signed<T>::value ? BCB_CHAR : BCB_UCHAR) :
        sizeof(T) == 2 ? (std::is_signed<T>::value ? BCB_SHORT : BCB_USHORT) :
        sizeof(T) == 4 ? (std::is_signed<T>::value ? BCB_INT : BCB_UINT) :
                         (std::is_signed<T>::value ? BCB_LONG : BCB_ULONG);
};

} // namespace detail

// is_fundamental<T> (type_traits/is_fundamental.hpp:28-55) restricted to the scalar types of the path
template<class T>
struct is_fundamental : std::integral_constant<bool, detail::dtype_of<T>::supported> {};

namespace detail {
// is_radix_sortable<T> (algorithm/detail/radix_sort.hpp:39-47): fundamental and not a vector type
template<class T>
struct is_radix_sortable : is_fundamental<T> {};
} // namespace detail

} // namespace compute
} // namespace boost

#endif
