// Drop-in acceptance example: the shape of the reference's example/sort_vector.cpp (default queue, host vector
// -> device vector -> sort -> copy back) compiled against this repository's include/ and libcompute_b200.so.
#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <vector>

#include <boost/compute/algorithm/copy.hpp>
#include <boost/compute/algorithm/sort.hpp>
#include <boost/compute/container/vector.hpp>
#include <boost/compute/system.hpp>

namespace compute = boost::compute;

int main()
{
    compute::device gpu = compute::system::default_device();
    std::cout << "device: " << gpu.name() << std::endl;

    std::vector<int> host_vector(10000);
    std::generate(host_vector.begin(), host_vector.end(), rand);

    compute::vector<int> device_vector = host_vector;
    compute::sort(device_vector.begin(), device_vector.end());
    compute::copy(device_vector.begin(), device_vector.end(), host_vector.begin());

    const bool ok = std::is_sorted(host_vector.begin(), host_vector.end());
    std::cout << (ok ? "sorted" : "NOT sorted") << std::endl;
    return ok ? 0 : 1;
}
