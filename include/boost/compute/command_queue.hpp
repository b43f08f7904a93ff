// command_queue (command_queue.hpp:78-1960 of the reference): an in-order queue = one CUDA stream.
// Algorithms enqueue work and return; finish() waits (command_queue.hpp:1564-1572).
#ifndef B200_BOOST_COMPUTE_COMMAND_QUEUE_HPP
#define B200_BOOST_COMPUTE_COMMAND_QUEUE_HPP

#include <cstddef>
#include <memory>

#include <boost/compute/buffer.hpp>
#include <boost/compute/context.hpp>
#include <boost/compute/device.hpp>

namespace boost {
namespace compute {

class command_queue
{
public:
    enum properties {
        enable_profiling = (1 << 1),
        enable_out_of_order_execution = (1 << 0)
    };

    command_queue() {}

    command_queue(const context &ctx, const device &dev, unsigned long long props = 0)
        : m_state(std::make_shared<state>(ctx, dev))
    {
        (void) props;
    }

    // wraps an existing cudaStream_t without owning it (e.g. a framework's stream)
    static command_queue attach(const context &ctx, const device &dev, void *cuda_stream)
    {
        command_queue q;
        q.m_state = std::make_shared<state>(ctx, dev, cuda_stream);
        return q;
    }

    device get_device() const { return m_state ? m_state->dev : device(); }
    context get_context() const { return m_state ? m_state->ctx : context(); }

    // the cudaStream_t handed to the C ABI
    void* get() const { return m_state ? m_state->stream : 0; }

    void finish() { make_current(); detail::check(bcb_stream_synchronize(get())); }
    void flush() {}

    // selects this queue's device for the calling thread; every algorithm calls it before touching the C ABI
    void make_current() const
    {
        if(m_state){
            detail::check(bcb_set_device(m_state->dev.id()));
        }
    }

    // enqueue_write_buffer / read / copy (command_queue.hpp:297-675); offsets and sizes in bytes.
    // The blocking forms of the reference are kept: the call returns when the host memory may be reused.
    void enqueue_write_buffer(const buffer &b, std::size_t offset, std::size_t size, const void *host_ptr)
    {
        make_current();
        detail::check(bcb_memcpy_h2d(get(), static_cast<char *>(b.get()) + offset, host_ptr, size));
        detail::check(bcb_stream_synchronize(get()));
    }
    void enqueue_read_buffer(const buffer &b, std::size_t offset, std::size_t size, void *host_ptr)
    {
        make_current();
        detail::check(bcb_memcpy_d2h(get(), host_ptr, static_cast<const char *>(b.get()) + offset, size));
        detail::check(bcb_stream_synchronize(get()));
    }
    void enqueue_copy_buffer(const buffer &src, const buffer &dst, std::size_t src_offset, std::size_t dst_offset,
                             std::size_t size)
    {
        make_current();
        detail::check(bcb_memcpy_d2d(get(), static_cast<char *>(dst.get()) + dst_offset,
                                     static_cast<const char *>(src.get()) + src_offset, size));
    }

    bool operator==(const command_queue &other) const { return m_state == other.m_state; }
    bool operator!=(const command_queue &other) const { return m_state != other.m_state; }

private:
    struct state
    {
        state(const context &c, const device &d) : ctx(c), dev(d), stream(0), owned(true)
        {
            detail::check(bcb_stream_create(d.id(), &stream));
        }
        state(const context &c, const device &d, void *s) : ctx(c), dev(d), stream(s), owned(false) {}
        ~state()
        {
            bcb_set_device(dev.id());
            if(owned){
                bcb_stream_destroy(stream);
            } else {
                bcb_workspace_release(stream);
            }
        }
        context ctx;
        device dev;
        void *stream;
        bool owned;
    };

    std::shared_ptr<state> m_state;
};

inline buffer buffer::clone(command_queue &queue) const
{
    buffer copy(get_context(), size());
    queue.enqueue_copy_buffer(*this, copy, 0, 0, size());
    return copy;
}

} // namespace compute
} // namespace boost

#endif
