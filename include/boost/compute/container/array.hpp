// array<T, N> (container/array.hpp:48-281 of the reference): a fixed-size device array with the std::array surface.
// Host-side construction takes std::array (the reference takes boost::array; Boost is not a dependency here).
#ifndef B200_BOOST_COMPUTE_CONTAINER_ARRAY_HPP
#define B200_BOOST_COMPUTE_CONTAINER_ARRAY_HPP

#include <array>
#include <cstddef>
#include <stdexcept>

#include <boost/compute/algorithm/copy.hpp>
#include <boost/compute/algorithm/fill.hpp>
#include <boost/compute/buffer.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>
#include <boost/compute/system.hpp>

namespace boost {
namespace compute {

template<class T, std::size_t N>
class array
{
public:
    typedef T value_type;
    typedef std::size_t size_type;
    typedef std::ptrdiff_t difference_type;
    typedef detail::buffer_value<T> reference;
    typedef const detail::buffer_value<T> const_reference;
    typedef buffer_iterator<T> iterator;
    typedef buffer_iterator<T> const_iterator;
    enum { static_size = N };

    explicit array(const context &ctx = system::default_context())
        : m_buffer(ctx, sizeof(T) * N)
    {
    }

    array(const std::array<T, N> &host, command_queue &queue = system::default_queue())
        : m_buffer(queue.get_context(), sizeof(T) * N)
    {
        ::boost::compute::copy(host.begin(), host.end(), begin(), queue);
    }

    array(const array<T, N> &other)
        : m_buffer(other.m_buffer.get_context(), sizeof(T) * N)
    {
        command_queue &queue = system::default_queue();
        ::boost::compute::copy(other.begin(), other.end(), begin(), queue);
        queue.finish();
    }

    array<T, N>& operator=(const array<T, N> &other)
    {
        if(this != &other){
            command_queue &queue = system::default_queue();
            ::boost::compute::copy(other.begin(), other.end(), begin(), queue);
            queue.finish();
        }
        return *this;
    }

    array<T, N>& operator=(const std::array<T, N> &host)
    {
        command_queue &queue = system::default_queue();
        ::boost::compute::copy(host.begin(), host.end(), begin(), queue);
        queue.finish();
        return *this;
    }

    iterator begin() const { return iterator(m_buffer, 0); }
    iterator end() const { return iterator(m_buffer, N); }
    const_iterator cbegin() const { return begin(); }
    const_iterator cend() const { return end(); }

    size_type size() const { return N; }
    bool empty() const { return N == 0; }
    size_type max_size() const { return N; }

    reference operator[](size_type index) { return *(begin() + static_cast<difference_type>(index)); }
    const_reference operator[](size_type index) const { return *(begin() + static_cast<difference_type>(index)); }

    reference at(size_type index)
    {
        if(index >= N){
            throw std::out_of_range("index out of range");
        }
        return operator[](index);
    }

    reference front() { return operator[](0); }
    reference back() { return operator[](N - 1); }

    void fill(const value_type &value, command_queue &queue) { ::boost::compute::fill(begin(), end(), value, queue); }

    void fill(const value_type &value)
    {
        command_queue &queue = system::default_queue();
        fill(value, queue);
        queue.finish();
    }

    const buffer& get_buffer() const { return m_buffer; }

private:
    buffer m_buffer;
};

} // namespace compute
} // namespace boost

#endif
