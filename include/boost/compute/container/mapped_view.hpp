// mapped_view<T> (container/mapped_view.hpp:43-245 of the reference): a device-iterable view of an existing host
// range.  The reference creates a CL_MEM_USE_HOST_PTR buffer; here the range is registered with the CUDA driver and
// kernels address it in place over PCIe (zero copy).  The host range must outlive the view.
#ifndef B200_BOOST_COMPUTE_CONTAINER_MAPPED_VIEW_HPP
#define B200_BOOST_COMPUTE_CONTAINER_MAPPED_VIEW_HPP

#include <cstddef>

#include <boost/compute/buffer.hpp>
#include <boost/compute/command_queue.hpp>
#include <boost/compute/iterator/buffer_iterator.hpp>
#include <boost/compute/system.hpp>

namespace boost {
namespace compute {

template<class T>
class mapped_view
{
public:
    typedef T value_type;
    typedef std::size_t size_type;
    typedef std::ptrdiff_t difference_type;
    typedef buffer_iterator<T> iterator;
    typedef buffer_iterator<T> const_iterator;

    mapped_view() : m_size(0) {}

    mapped_view(T *host_ptr, size_type n, const context &ctx = system::default_context())
        : m_buffer(n ? buffer::use_host_ptr(ctx, host_ptr, n * sizeof(T)) : buffer()), m_size(n)
    {
    }

    // (the reference maps const ranges read-only; the registration itself does not write)
    mapped_view(const T *host_ptr, size_type n, const context &ctx = system::default_context())
        : m_buffer(n ? buffer::use_host_ptr(ctx, const_cast<T *>(host_ptr), n * sizeof(T)) : buffer()), m_size(n)
    {
    }

    iterator begin() const { return iterator(m_buffer, 0); }
    iterator end() const { return iterator(m_buffer, m_size); }
    const_iterator cbegin() const { return begin(); }
    const_iterator cend() const { return end(); }

    size_type size() const { return m_size; }
    bool empty() const { return m_size == 0; }
    const buffer& get_buffer() const { return m_buffer; }

    // map(): make the device's writes visible to the host (mapped_view.hpp:183-203 enqueues a blocking map);
    // unmap(): hand the range back to the device -- nothing to do for coherent zero-copy memory
    void map(command_queue &queue) { queue.finish(); }
    void unmap(command_queue &) {}

private:
    buffer m_buffer;
    size_type m_size;
};

} // namespace compute
} // namespace boost

#endif
