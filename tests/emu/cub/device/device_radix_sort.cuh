// tests/emu/cub/device/device_radix_sort.cuh -- TEST INFRASTRUCTURE: host stand-in for the one CUB entry
// point the library uses (a stable LSD radix sort of key/value pairs on bits [begin_bit, end_bit)).
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* d_temp, size_t& temp_bytes, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out, int num,
                                 int begin_bit, int end_bit, cudaStream_t = nullptr) {
        if (d_temp == nullptr) { temp_bytes = 64; return cudaSuccess; }
        const K mask = (end_bit - begin_bit >= (int)(8 * sizeof(K))) ? ~K(0) : (K)(((K(1) << (end_bit - begin_bit)) - 1) << begin_bit);
        std::vector<int> order((size_t)num);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return (keys_in[a] & mask) < (keys_in[b] & mask); });
        for (int i = 0; i < num; ++i) { keys_out[i] = keys_in[order[(size_t)i]]; vals_out[i] = vals_in[order[(size_t)i]]; }
        return cudaSuccess;
    }
};
}  // namespace cub
