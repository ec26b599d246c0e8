// Error buffer, device query and CUtensorMap encoding (driver entry point resolved at run time).
#include "common.cuh"

#include <atomic>
#include <cstring>
#include <mutex>

namespace wgb {

static thread_local char g_error[512] = "";

char* error_buffer() { return g_error; }

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return status;
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(WGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(WGB_ERR_ARGUMENT, "tensor base must be 16 B aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = 1;
        if (i > 0) {
            gstr[i - 1] = strides_bytes[i - 1];
            if (gstr[i - 1] % 16 != 0) return fail(WGB_ERR_ARGUMENT, "tensor stride %d not a multiple of 16 B", i);
        }
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                    gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WGB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    return WGB_OK;
}

static std::atomic<int> g_gate_l2_hint{1}, g_res_l2_hint{0}, g_stft_l2_hint{0}, g_fft_mel_warps{12}, g_pdl{1};

int tuning_get(const char* key) {
    if (std::strcmp(key, "gate_l2_hint") == 0) return g_gate_l2_hint.load(std::memory_order_relaxed);
    if (std::strcmp(key, "res_l2_hint") == 0) return g_res_l2_hint.load(std::memory_order_relaxed);
    if (std::strcmp(key, "stft_l2_hint") == 0) return g_stft_l2_hint.load(std::memory_order_relaxed);
    if (std::strcmp(key, "fft_mel_warps") == 0) return g_fft_mel_warps.load(std::memory_order_relaxed);
    if (std::strcmp(key, "pdl") == 0) return g_pdl.load(std::memory_order_relaxed);
    return 0;
}
int tuning_set(const char* key, int value) {
    if (!key) return fail(WGB_ERR_ARGUMENT, "null key");
    if (std::strcmp(key, "gate_l2_hint") == 0) {
        g_gate_l2_hint.store(value, std::memory_order_relaxed);
        return WGB_OK;
    }
    if (std::strcmp(key, "res_l2_hint") == 0) {
        g_res_l2_hint.store(value, std::memory_order_relaxed);
        return WGB_OK;
    }
    if (std::strcmp(key, "stft_l2_hint") == 0) {
        g_stft_l2_hint.store(value, std::memory_order_relaxed);
        return WGB_OK;
    }
    if (std::strcmp(key, "fft_mel_warps") == 0) {
        g_fft_mel_warps.store(value, std::memory_order_relaxed);
        return WGB_OK;
    }
    if (std::strcmp(key, "pdl") == 0) {
        g_pdl.store(value, std::memory_order_relaxed);
        return WGB_OK;
    }
    return fail(WGB_ERR_ARGUMENT, "unknown tuning key '%s'", key);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
        cached = prop.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace wgb
