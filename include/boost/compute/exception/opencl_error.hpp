// opencl_error (exception/opencl_error.hpp:30-61): the exception every failing runtime call turns into.
// The error code is a cudaError_t value or a BCB_E* code of the C ABI instead of a CL_* constant.
#ifndef B200_BOOST_COMPUTE_EXCEPTION_OPENCL_ERROR_HPP
#define B200_BOOST_COMPUTE_EXCEPTION_OPENCL_ERROR_HPP

#include <exception>
#include <string>

#include <compute_b200.h>

namespace boost {
namespace compute {

class opencl_error : public std::exception
{
public:
    explicit opencl_error(int error) throw()
        : m_error(error), m_error_string(to_string(error))
    {
    }
    ~opencl_error() throw() {}

    int error_code() const throw() { return m_error; }
    std::string error_string() const throw() { return m_error_string; }
    const char* what() const throw() { return m_error_string.c_str(); }

    static std::string to_string(int error)
    {
        const char *s = bcb_error_string(error);
        return s ? std::string(s) : std::string("unknown error");
    }

private:
    int m_error;
    std::string m_error_string;
};

// no_device_found (exception/no_device_found.hpp; thrown by system.hpp:238-241)
class no_device_found : public std::exception
{
public:
    const char* what() const throw() { return "No OpenCL device found"; }
};

namespace detail {
inline void check(int status)
{
    if(status != BCB_SUCCESS){
        throw opencl_error(status);
    }
}
} // namespace detail

} // namespace compute
} // namespace boost

#endif
