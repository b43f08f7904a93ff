// vector<T> (container/vector.hpp:103-790 of the reference): a resizable array in device memory, the range
// type every algorithm of the path is called on.  The subset kept here is what the path, its tests and the
// perf harness use: construction from counts / fill values / host ranges / std::vector, size bookkeeping with
// the reference's growth policy (minimum capacity 4, x1.5), element access through buffer_value proxies,
// push_back / resize / assign and begin()/end() buffer_iterators.
#ifndef B200_BOOST_COMPUTE_CONTAINER_VECTOR_HPP
#define B200_BOOST_COMPUTE_CONTAINER_VECTOR_HPP

#include <algorithm>
#include <cstddef>
#include <stdexcept>
#include <vector>

#include <boost/compute/algorithm/copy.hpp>
#include <boost/compute/algorithm/fill.hpp>
#include <boost/compute/buffer.hpp>
#include <boost/compute/detail/default_queue.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>

namespace boost {
namespace compute {

template<class T>
class vector
{
public:
    typedef T value_type;
    typedef std::size_t size_type;
    typedef std::ptrdiff_t difference_type;
    typedef detail::buffer_value<T> reference;
    typedef const detail::buffer_value<T> const_reference;
    typedef buffer_iterator<T> iterator;
    typedef buffer_iterator<T> const_iterator;

    explicit vector(const context &ctx = system::default_context())
        : m_context(ctx), m_size(0)
    {
        allocate(min_capacity);
    }

    explicit vector(size_type count, const context &ctx = system::default_context())
        : m_context(ctx), m_size(count)
    {
        allocate(std::max(count, size_type(min_capacity)));
    }

    vector(size_type count, const T &value, command_queue &queue = system::default_queue())
        : m_context(queue.get_context()), m_size(count)
    {
        allocate(std::max(count, size_type(min_capacity)));
        ::boost::compute::fill(begin(), end(), value, queue);
    }

    template<class InputIterator>
    vector(InputIterator first, InputIterator last, command_queue &queue = system::default_queue(),
           typename std::enable_if<!std::is_integral<InputIterator>::value>::type* = 0)
        : m_context(queue.get_context()), m_size(detail::iterator_range_size(first, last))
    {
        allocate(std::max(m_size, size_type(min_capacity)));
        ::boost::compute::copy(first, last, begin(), queue);
    }

    vector(const std::vector<T> &host, command_queue &queue = system::default_queue())
        : m_context(queue.get_context()), m_size(host.size())
    {
        allocate(std::max(m_size, size_type(min_capacity)));
        if(!host.empty()){
            ::boost::compute::copy(&host[0], &host[0] + host.size(), begin(), queue);
        }
    }

    vector(const vector &other)
        : m_context(other.m_context), m_size(other.m_size)
    {
        allocate(std::max(m_size, size_type(min_capacity)));
        command_queue &queue = system::default_queue();
        ::boost::compute::copy(other.begin(), other.end(), begin(), queue);
        queue.finish();
    }

    vector& operator=(const vector &other)
    {
        if(this != &other){
            command_queue &queue = system::default_queue();
            resize(other.size(), queue);
            ::boost::compute::copy(other.begin(), other.end(), begin(), queue);
            queue.finish();
        }
        return *this;
    }

    iterator begin() const { return iterator(m_data, 0); }
    iterator end() const { return iterator(m_data, m_size); }
    const_iterator cbegin() const { return begin(); }
    const_iterator cend() const { return end(); }

    size_type size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    size_type capacity() const { return m_data.size() / sizeof(T); }
    size_type max_size() const { return static_cast<size_type>(-1) / sizeof(T); }

    void reserve(size_type count, command_queue &queue = system::default_queue())
    {
        if(count > capacity()){
            grow_to(count, queue);
        }
    }

    void resize(size_type count, command_queue &queue = system::default_queue())
    {
        if(count > capacity()){
            grow_to(std::max(count, next_capacity()), queue);
        }
        m_size = count;
    }

    void shrink_to_fit(command_queue &queue = system::default_queue())
    {
        grow_to(std::max(m_size, size_type(min_capacity)), queue);
    }

    reference operator[](size_type index) { return reference(m_data, index * sizeof(T)); }
    const_reference operator[](size_type index) const { return const_reference(m_data, index * sizeof(T)); }

    reference at(size_type index)
    {
        if(index >= m_size){
            throw std::out_of_range("index out of range");
        }
        return operator[](index);
    }

    reference front() { return operator[](0); }
    reference back() { return operator[](m_size - 1); }

    template<class InputIterator>
    void assign(InputIterator first, InputIterator last, command_queue &queue = system::default_queue())
    {
        resize(detail::iterator_range_size(first, last), queue);
        ::boost::compute::copy(first, last, begin(), queue);
    }

    void assign(size_type count, const T &value, command_queue &queue = system::default_queue())
    {
        resize(count, queue);
        ::boost::compute::fill(begin(), end(), value, queue);
    }

    void push_back(const T &value, command_queue &queue = system::default_queue())
    {
        if(m_size == capacity()){
            grow_to(next_capacity(), queue);
        }
        queue.enqueue_write_buffer(m_data, m_size * sizeof(T), sizeof(T), &value);
        ++m_size;
    }

    void pop_back(command_queue & = system::default_queue()) { --m_size; }
    void clear() { m_size = 0; }

    void swap(vector &other)
    {
        std::swap(m_data, other.m_data);
        std::swap(m_size, other.m_size);
        std::swap(m_context, other.m_context);
    }

    const buffer& get_buffer() const { return m_data; }
    context get_context() const { return m_context; }

private:
    enum { min_capacity = 4 };

    size_type next_capacity() const
    {
        const size_type c = capacity();
        return std::max(size_type(min_capacity), c + c / 2 + 1); // x1.5 growth (container/vector.hpp growth policy)
    }

    void allocate(size_type count) { m_data = buffer(m_context, count * sizeof(T)); }

    void grow_to(size_type count, command_queue &queue)
    {
        buffer bigger(m_context, count * sizeof(T));
        const size_type keep = std::min(m_size, count);
        if(keep){
            queue.enqueue_copy_buffer(m_data, bigger, 0, 0, keep * sizeof(T));
            queue.finish(); // the old block is released when m_data is reassigned
        }
        m_data = bigger;
    }

    context m_context;
    buffer m_data;
    size_type m_size;
};

} // namespace compute
} // namespace boost

#endif
