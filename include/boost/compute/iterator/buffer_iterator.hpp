// buffer_iterator<T> (iterator/buffer_iterator.hpp:136-271 of the reference): random-access iterator over
// a device buffer = (buffer, element index).  It does not keep the buffer alive on its own in the reference
// either; here it shares ownership, which is harmless.
#ifndef B200_BOOST_COMPUTE_ITERATOR_BUFFER_ITERATOR_HPP
#define B200_BOOST_COMPUTE_ITERATOR_BUFFER_ITERATOR_HPP

#include <cstddef>
#include <iterator>
#include <type_traits>

#include <boost/compute/buffer.hpp>
#include <boost/compute/command_queue.hpp>

namespace boost {
namespace compute {

class system;

namespace detail {

command_queue& default_queue_ref(); // defined in detail/default_queue.hpp (avoids a cycle with system.hpp)

// buffer_value<T> (detail/buffer_value.hpp:58-69): proxy for one element living in device memory
template<class T>
class buffer_value
{
public:
    buffer_value(const buffer &b, std::size_t byte_offset) : m_buffer(b), m_offset(byte_offset) {}

    operator T() const
    {
        T value;
        default_queue_ref().enqueue_read_buffer(m_buffer, m_offset, sizeof(T), &value);
        return value;
    }

    buffer_value& operator=(const T &value)
    {
        default_queue_ref().enqueue_write_buffer(m_buffer, m_offset, sizeof(T), &value);
        return *this;
    }

    buffer_value& operator=(const buffer_value &other)
    {
        return *this = static_cast<T>(other);
    }

private:
    buffer m_buffer;
    std::size_t m_offset;
};

} // namespace detail

template<class T>
class buffer_iterator
{
public:
    typedef std::random_access_iterator_tag iterator_category;
    typedef T value_type;
    typedef std::ptrdiff_t difference_type;
    typedef T* pointer;
    typedef detail::buffer_value<T> reference;

    buffer_iterator() : m_index(0) {}
    buffer_iterator(const buffer &b, std::size_t index) : m_buffer(b), m_index(index) {}

    const buffer& get_buffer() const { return m_buffer; }
    std::size_t get_index() const { return m_index; }

    // raw device address of the element this iterator points at (what the C ABI consumes)
    T* device_ptr() const { return static_cast<T *>(m_buffer.get()) + m_index; }

    T read(command_queue &queue) const
    {
        T value;
        queue.enqueue_read_buffer(m_buffer, m_index * sizeof(T), sizeof(T), &value);
        return value;
    }

    void write(const T &value, command_queue &queue)
    {
        queue.enqueue_write_buffer(m_buffer, m_index * sizeof(T), sizeof(T), &value);
    }

    reference operator*() const { return reference(m_buffer, m_index * sizeof(T)); }
    reference operator[](difference_type n) const { return reference(m_buffer, (m_index + n) * sizeof(T)); }

    buffer_iterator& operator++() { ++m_index; return *this; }
    buffer_iterator operator++(int) { buffer_iterator t(*this); ++m_index; return t; }
    buffer_iterator& operator--() { --m_index; return *this; }
    buffer_iterator operator--(int) { buffer_iterator t(*this); --m_index; return t; }
    buffer_iterator& operator+=(difference_type n) { m_index += n; return *this; }
    buffer_iterator& operator-=(difference_type n) { m_index -= n; return *this; }
    buffer_iterator operator+(difference_type n) const { return buffer_iterator(m_buffer, m_index + n); }
    buffer_iterator operator-(difference_type n) const { return buffer_iterator(m_buffer, m_index - n); }
    difference_type operator-(const buffer_iterator &o) const
    {
        return static_cast<difference_type>(m_index) - static_cast<difference_type>(o.m_index);
    }

    bool operator==(const buffer_iterator &o) const { return m_buffer.get() == o.m_buffer.get() && m_index == o.m_index; }
    bool operator!=(const buffer_iterator &o) const { return !(*this == o); }
    bool operator<(const buffer_iterator &o) const { return m_index < o.m_index; }
    bool operator>(const buffer_iterator &o) const { return m_index > o.m_index; }
    bool operator<=(const buffer_iterator &o) const { return m_index <= o.m_index; }
    bool operator>=(const buffer_iterator &o) const { return m_index >= o.m_index; }

private:
    buffer m_buffer;
    std::size_t m_index;
};

template<class T>
inline buffer_iterator<T> operator+(std::ptrdiff_t n, const buffer_iterator<T> &it) { return it + n; }

// make_buffer_iterator (iterator/buffer_iterator.hpp:266-271)
template<class T>
inline buffer_iterator<T> make_buffer_iterator(const buffer &b, std::size_t index = 0)
{
    return buffer_iterator<T>(b, index);
}

// is_device_iterator (type_traits/is_device_iterator.hpp)
template<class Iterator>
struct is_device_iterator : std::false_type {};
template<class T>
struct is_device_iterator<buffer_iterator<T> > : std::true_type {};
template<class T>
struct is_device_iterator<const buffer_iterator<T> > : std::true_type {};

namespace detail {
// iterator_range_size (detail/iterator_range_size.hpp)
template<class Iterator>
inline std::size_t iterator_range_size(Iterator first, Iterator last)
{
    return static_cast<std::size_t>(std::distance(first, last));
}
} // namespace detail

} // namespace compute
} // namespace boost

#endif
