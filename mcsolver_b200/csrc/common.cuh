// Shared host/device helpers of the engine: error plumbing, block reductions, spin maths.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/mcsolver_b200.h"
#include "devmath.cuh"

namespace mcg {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);

// Large per-system buffers (spin planes, Wolff forests) come from the device's stream-ordered memory pool with
// the release threshold lifted, so that back-to-back jobs (one system per MCMainFunction-like call) reuse the
// same pages instead of paying a multi-millisecond cudaMalloc/cudaFree of gigabytes each time.
void *pool_alloc(size_t bytes);
void pool_free(void *p);

#define MCG_CUDA(call)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess)                                                                              \
            throw ::mcg::Error(MCG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                                 std::to_string(__LINE__) + ")");                           \
    } while (0)

#define MCG_REQUIRE(cond, msg)                                          \
    do {                                                                \
        if (!(cond)) throw ::mcg::Error(MCG_ERR_ARG, std::string(msg)); \
    } while (0)

}  // namespace mcg
