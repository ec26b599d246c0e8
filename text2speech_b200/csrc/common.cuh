// Host-side plumbing shared by every translation unit: error reporting for the C ABI and
// CUtensorMap construction through the driver entry point (no link-time libcuda dependency, so the
// library loads on a machine without a GPU driver — the symbol/ABI tests run there).
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

namespace wgb {

enum Status : int {
    WGB_OK = 0,
    WGB_ERR_ARGUMENT = 1,   // bad shape / null pointer / unsupported size
    WGB_ERR_CUDA = 2,       // CUDA runtime or driver call failed
    WGB_ERR_DEVICE = 3,     // not an sm_100 device
};

char* error_buffer();   // thread-local, defined in api.cu
int fail(int status, const char* fmt, ...);

#define WGB_CUDA_TRY(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return ::wgb::fail(::wgb::WGB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                         \
    } while (0)

#define WGB_REQUIRE(cond, ...)                                                                      \
    do {                                                                                            \
        if (!(cond)) return ::wgb::fail(::wgb::WGB_ERR_ARGUMENT, __VA_ARGS__);                      \
    } while (0)

#define WGB_LAUNCH_CHECK()  WGB_CUDA_TRY(cudaGetLastError())

// bf16 tensor map with SWIZZLE_128B, zero OOB fill.  dims/strides innermost first; strides in BYTES
// for dims 1..rank-1 (dim 0 is contiguous).  box innermost must be 64 elements (128 B).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                   const uint64_t* strides_bytes, const uint32_t* box);

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

int sm_count();   // SMs of the current device (cached)

// process-wide experiment switches (wgb_set_tuning); unknown keys read as 0
int tuning_get(const char* key);
int tuning_set(const char* key, int value);

}  // namespace wgb
