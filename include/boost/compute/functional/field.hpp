// Comparators that look at ONE scalar field of a record:  f(a.field) < f(b.field)  or  ">".
// The reference turns an arbitrary compare(a, b) -- a BOOST_COMPUTE_FUNCTION, a lambda expression -- into OpenCL C at run
// time and hands it to its merge sort (algorithm/sort.hpp:83-106, stable_sort.hpp:34-50,
// detail/merge_sort_on_gpu.hpp:523-572).  Ahead of time that family is closed; every comparator the reference's own tests
// use is in it (int2_ by .x / .y, a struct by a member, ints by abs(): test_sort.cpp:294-360, test_stable_sort.cpp:41-90,
// test_merge_sort_gpu.cpp:223-380).  Spellings:
//     less_by(&Particle::x), greater_by(&Particle::x)           a member of a struct  (sort_by_x of test_sort.cpp:297)
//     less_by_component<0>(), greater_by_component<1>()         a component of a vector type (a.x < b.x, a.y < b.y)
//     less_by(&Particle::x).abs(), less_abs<int>()              through abs()  (abs_sort of test_merge_sort_gpu.cpp:239)
//     lambda::get<0>(_1) < lambda::get<0>(_2), abs(_1) < abs(_2)  the reference's own lambda spelling (lambda/placeholders.hpp)
// sort(), stable_sort(), detail::merge_sort_on_gpu() and is_sorted() accept them; all three sorts are stable.
#ifndef B200_BOOST_COMPUTE_FUNCTIONAL_FIELD_HPP
#define B200_BOOST_COMPUTE_FUNCTIONAL_FIELD_HPP

#include <cstddef>
#include <type_traits>

#include <compute_b200.h>

#include <boost/compute/detail/dtype.hpp>
#include <boost/compute/types/fundamental.hpp>

namespace boost {
namespace compute {

// what the C ABI needs to know about a field comparator on records of type T
struct field_spec
{
    std::size_t offset;  // bytes from the start of the record
    int dtype;           // bcb_dtype of the field
    int unary;           // BCB_UN_IDENTITY or BCB_UN_ABS
    bool descending;     // ">" instead of "<"
};

// a member of a struct (resolved when the comparator is built)
template<class T, class M>
struct member_compare
{
    static_assert(detail::dtype_of<M>::supported, "field comparators look at a scalar member");
    typedef bool result_type;
    field_spec spec;

    member_compare abs() const { member_compare c = *this; c.spec.unary = BCB_UN_ABS; return c; }
    template<class U> field_spec resolve() const
    {
        // (a member of a base class -- &int2_::x names a member of the vector type's base -- sits at the same offset when
        // the base is the record's first and only base: standard layout guarantees it)
        static_assert(std::is_same<U, T>::value || (std::is_base_of<T, U>::value && std::is_standard_layout<U>::value),
                      "the comparator was built for another record type");
        return spec;
    }
    bool operator()(const T &a, const T &b) const  // the same comparison on the host
    {
        const M x = *reinterpret_cast<const M *>(reinterpret_cast<const char *>(&a) + spec.offset);
        const M y = *reinterpret_cast<const M *>(reinterpret_cast<const char *>(&b) + spec.offset);
        const M fx = (spec.unary == BCB_UN_ABS && x < M(0)) ? static_cast<M>(-x) : x, fy = (spec.unary == BCB_UN_ABS && y < M(0)) ? static_cast<M>(-y) : y;
        return spec.descending ? fx > fy : fx < fy;
    }
};

namespace detail {
template<class T, class M>
inline member_compare<T, M> make_member_compare(M T::*member, bool descending)
{
    static_assert(std::is_standard_layout<T>::value, "field comparators need standard-layout records");
    typename std::aligned_storage<sizeof(T), alignof(T)>::type storage;
    const T *obj = reinterpret_cast<const T *>(&storage);
    member_compare<T, M> c;
    c.spec.offset = static_cast<std::size_t>(reinterpret_cast<const char *>(&(obj->*member)) - reinterpret_cast<const char *>(obj));
    c.spec.dtype = dtype_of<M>::value;
    c.spec.unary = BCB_UN_IDENTITY;
    c.spec.descending = descending;
    return c;
}
} // namespace detail

template<class T, class M> inline member_compare<T, M> less_by(M T::*member) { return detail::make_member_compare(member, false); }
template<class T, class M> inline member_compare<T, M> greater_by(M T::*member) { return detail::make_member_compare(member, true); }

// a component of a vector type, or the scalar itself, resolved against the element type at the call
struct component_compare
{
    typedef bool result_type;
    int index;       // component (0 for a scalar)
    int unary;
    bool descending;

    component_compare abs() const { component_compare c = *this; c.unary = BCB_UN_ABS; return c; }
    template<class U> field_spec resolve() const { return resolve_impl<U>(is_vector_type<U>()); }

private:
    template<class U> field_spec resolve_impl(std::true_type) const
    {
        typedef typename U::scalar_type S;
        field_spec s = { static_cast<std::size_t>(index) * sizeof(S), detail::dtype_of<S>::value, unary, descending };
        return s;
    }
    template<class U> field_spec resolve_impl(std::false_type) const
    {
        static_assert(detail::dtype_of<U>::supported, "component comparators need a scalar or vector element type");
        field_spec s = { 0, detail::dtype_of<U>::value, unary, descending };
        return s;
    }
};

template<int N> inline component_compare less_by_component() { component_compare c = { N, BCB_UN_IDENTITY, false }; return c; }
template<int N> inline component_compare greater_by_component() { component_compare c = { N, BCB_UN_IDENTITY, true }; return c; }
template<class T> inline component_compare less_abs() { component_compare c = { 0, BCB_UN_ABS, false }; return c; }
template<class T> inline component_compare greater_abs() { component_compare c = { 0, BCB_UN_ABS, true }; return c; }

template<class C> struct is_field_compare : std::false_type {};
template<class T, class M> struct is_field_compare<member_compare<T, M> > : std::true_type {};
template<> struct is_field_compare<component_compare> : std::true_type {};

} // namespace compute
} // namespace boost

#endif
