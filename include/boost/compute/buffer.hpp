// buffer (buffer.hpp:54-191 + memory_object.hpp of the reference): a reference-counted block of device
// memory.  Copies share the allocation (the reference retains the cl_mem); the memory is released when
// the last copy goes away.
#ifndef B200_BOOST_COMPUTE_BUFFER_HPP
#define B200_BOOST_COMPUTE_BUFFER_HPP

#include <cstddef>
#include <memory>

#include <boost/compute/context.hpp>

namespace boost {
namespace compute {

class command_queue;

class buffer
{
public:
    enum mem_flags {
        read_write = (1 << 0),
        read_only = (1 << 2),
        write_only = (1 << 1)
    };

    buffer() {}

    buffer(const context &ctx, std::size_t size, unsigned long long flags = read_write, void * = 0)
        : m_storage(std::make_shared<storage>(ctx, size))
    {
        (void) flags;
    }

    // a buffer that aliases an existing host range (the reference's CL_MEM_USE_HOST_PTR buffers, buffer.hpp:86-98 with
    // a host_ptr): the range is registered for device access and unregistered when the last copy goes away
    static buffer use_host_ptr(const context &ctx, void *host_ptr, std::size_t size)
    {
        buffer b;
        b.m_storage = std::make_shared<storage>(ctx, host_ptr, size);
        return b;
    }

    // device address of the first byte (plays the role of get() returning the cl_mem)
    void* get() const { return m_storage ? m_storage->ptr : 0; }
    std::size_t size() const { return m_storage ? m_storage->size : 0; }
    context get_context() const { return m_storage ? m_storage->ctx : context(); }

    buffer clone(command_queue &queue) const; // defined in command_queue.hpp

    bool operator==(const buffer &other) const { return get() == other.get(); }
    bool operator!=(const buffer &other) const { return get() != other.get(); }

private:
    struct storage
    {
        storage(const context &c, std::size_t bytes) : ctx(c), ptr(0), size(bytes), host(0)
        {
            detail::check(bcb_set_device(c.get_device().id()));
            detail::check(bcb_malloc(&ptr, bytes ? bytes : 1));
        }
        storage(const context &c, void *host_ptr, std::size_t bytes) : ctx(c), ptr(0), size(bytes), host(host_ptr)
        {
            detail::check(bcb_set_device(c.get_device().id()));
            detail::check(bcb_host_register(host_ptr, bytes, &ptr));
        }
        ~storage()
        {
            // frees are synchronous with respect to the device, so queued work that still uses the block is
            // finished first (the reference relies on the OpenCL runtime deferring release the same way)
            if(host){
                bcb_host_unregister(host); // (synchronises the device first)
            }
            else {
                bcb_free(ptr);
            }
        }
        storage(const storage &) = delete;
        storage& operator=(const storage &) = delete;
        context ctx;
        void *ptr;
        std::size_t size;
        void *host; // non-null: ptr is the device alias of this registered host range
    };

    std::shared_ptr<storage> m_storage;
};

} // namespace compute
} // namespace boost

#endif
