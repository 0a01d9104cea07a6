// Host-side helpers shared by the inference drivers.
#pragma once
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "philox.cuh"

namespace ppcseq {

// device allocations released together
struct DevBuf {
    std::vector<void *> ptrs;
    template <typename T>
    int get(T **p, size_t n) {
        *p = nullptr;
        cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(n, 1) * sizeof(T));
        if (e != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return PPCSEQ_ENOMEM; }
        ptrs.push_back((void *)*p);
        return PPCSEQ_OK;
    }
    ~DevBuf() { for (void *p : ptrs) cudaFree(p); }
};

// Philox-backed uniform stream for host-side decisions (tree direction, multinomial acceptance, inits)
struct HostRng {
    uint64_t seed, ctr = 0;
    uint32_t stream;
    uint32_t buf[4];
    int have = 0;
    HostRng(uint64_t s, uint32_t st) : seed(s), stream(st) {}
    double uniform() {                // (0,1)
        if (have < 2) {
            philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), stream, 0x686f7374u, (uint32_t)seed, (uint32_t)(seed >> 32), buf);
            ++ctr; have = 4;
        }
        const uint64_t a = buf[4 - have], b = buf[5 - have];
        have -= 2;
        return ((double)((a << 21) | (b >> 11)) + 0.5) * (1.0 / 9007199254740992.0);
    }
};

}  // namespace ppcseq
