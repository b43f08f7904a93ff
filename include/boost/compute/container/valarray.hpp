// valarray<T> (container/valarray.hpp:30-460 of the reference), the part on the reduce path: construction from a
// host array / size / fill value, size(), operator[], and the reductions sum() / min() / max() (:237-268), which are
// reduce(plus<T>) / min_element / max_element on the underlying device buffer.  The element-wise operator overloads of
// the reference (code-generated OpenCL expressions) are outside this path.
#ifndef B200_BOOST_COMPUTE_CONTAINER_VALARRAY_HPP
#define B200_BOOST_COMPUTE_CONTAINER_VALARRAY_HPP

#include <cstddef>

#include <boost/compute/algorithm/max_element.hpp>
#include <boost/compute/algorithm/min_element.hpp>
#include <boost/compute/algorithm/reduce.hpp>
#include <boost/compute/container/vector.hpp>
#include <boost/compute/functional/operator.hpp>

namespace boost {
namespace compute {

template<class T>
class valarray
{
public:
    typedef T value_type;

    explicit valarray(const context &ctx = system::default_context()) : m_data(ctx) {}
    explicit valarray(size_t size, const context &ctx = system::default_context()) : m_data(size, ctx) {}
    valarray(const T &value, size_t size, const context &ctx = system::default_context()) : m_data(size, ctx)
    {
        ::boost::compute::fill(m_data.begin(), m_data.end(), value, system::default_queue());
    }
    valarray(const T *values, size_t size, const context &ctx = system::default_context()) : m_data(size, ctx)
    {
        ::boost::compute::copy(values, values + size, m_data.begin(), system::default_queue());
    }

    size_t size() const { return m_data.size(); }
    void resize(size_t size) { m_data.resize(size); }
    detail::buffer_value<T> operator[](size_t index) { return m_data[index]; }
    const detail::buffer_value<T> operator[](size_t index) const { return m_data[index]; }

    T sum() const  // container/valarray.hpp:237-246: reduce(begin, end, &result, plus<T>)
    {
        T result = T();
        ::boost::compute::reduce(m_data.begin(), m_data.end(), &result, plus<T>(), system::default_queue());
        return result;
    }
    T (min)() const  // :224-229: *min_element(begin, end)
    {
        command_queue &queue = system::default_queue();
        return (::boost::compute::min_element(m_data.begin(), m_data.end(), queue)).read(queue);
    }
    T (max)() const  // :231-236: *max_element(begin, end)
    {
        command_queue &queue = system::default_queue();
        return (::boost::compute::max_element(m_data.begin(), m_data.end(), queue)).read(queue);
    }

    const vector<T> &data() const { return m_data; }

private:
    vector<T> m_data;
};

} // namespace compute
} // namespace boost

#endif
