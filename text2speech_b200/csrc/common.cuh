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

// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// -- as far as SM resources allow -- once every CTA of the kernel before it in the stream has executed
// pdl_launch_dependents() (or exited); it must execute pdl_wait() before touching anything that kernel reads or writes:
// the wait returns when the earlier grid has completed and its memory is visible.  Both are no-ops for a kernel launched
// without the attribute / without a dependent.  The persistent WN kernels trigger at their first instruction, so the next
// kernel's CTAs arrive on an SM the moment this kernel's CTA there exits and run their prologue (barrier init, TMEM
// allocation, tensor-map prefetch, bias staging) while the slowest CTAs of this one are still finishing.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// <<<>>> with the programmatic-dependent-launch attribute when wgb_set_tuning("pdl") is on (default) and the kernel is
// one of those that call pdl_wait() before their first dependent access
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = tuning_get("pdl") ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace wgb
