// Internal helpers shared by the translation units of libkgb200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/kgb200.h"

namespace kgb {

void set_error(const char* fmt, ...);
int current_device();   // device bound by kgb_init on this thread (-1 if none)
int sm_count();         // SM count of the bound device

inline size_t dtype_size(int dt) {
    switch (dt) {
        case KGB_F64:
        case KGB_I64: return 8;
        case KGB_F32:
        case KGB_I32: return 4;
        case KGB_B8: return 1;
    }
    return 0;
}

#define KGB_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            kgb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                           __LINE__);                                                             \
            return KGB_ECUDA;                                                                     \
        }                                                                                         \
    } while (0)

#define KGB_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            kgb::set_error(__VA_ARGS__);  \
            return KGB_EINVAL;            \
        }                                 \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace kgb
