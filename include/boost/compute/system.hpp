// system (system.hpp:92-440 of the reference): default device / context / queue and device enumeration.
// Device selection honours BOOST_COMPUTE_DEFAULT_DEVICE (substring of the device name, system.hpp:244-248),
// otherwise device 0 of CUDA_VISIBLE_DEVICES.
#ifndef B200_BOOST_COMPUTE_SYSTEM_HPP
#define B200_BOOST_COMPUTE_SYSTEM_HPP

#include <cstdlib>
#include <string>
#include <vector>

#include <boost/compute/command_queue.hpp>
#include <boost/compute/context.hpp>
#include <boost/compute/device.hpp>

namespace boost {
namespace compute {

class system
{
public:
    static std::size_t device_count()
    {
        int n = 0;
        bcb_device_count(&n);
        return static_cast<std::size_t>(n);
    }

    static std::vector<device> devices()
    {
        std::vector<device> all;
        const std::size_t n = device_count();
        for(std::size_t i = 0; i < n; i++){
            all.push_back(device(static_cast<int>(i)));
        }
        return all;
    }

    static device find_device(const std::string &name)
    {
        const std::vector<device> all = devices();
        for(std::size_t i = 0; i < all.size(); i++){
            if(all[i].name() == name){
                return all[i];
            }
        }
        throw no_device_found();
    }

    static device default_device()
    {
        static device dev = find_default_device();
        return dev;
    }

    static context default_context()
    {
        static context ctx(default_device());
        return ctx;
    }

    static command_queue& default_queue()
    {
        static command_queue queue(default_context(), default_device());
        return queue;
    }

    static void finish() { default_queue().finish(); }

private:
    static device find_default_device()
    {
        const std::vector<device> all = devices();
        if(all.empty()){
            throw no_device_found();
        }
        const char *wanted = std::getenv("BOOST_COMPUTE_DEFAULT_DEVICE");
        if(wanted){
            for(std::size_t i = 0; i < all.size(); i++){
                if(all[i].name().find(wanted) != std::string::npos){
                    return all[i];
                }
            }
        }
        return all[0];
    }
};

} // namespace compute
} // namespace boost

#endif
