// FP32 validation mode: CUDA-core kernels that follow the reference op by op (in_layers conv,
// cond conv, gate, res_skip conv, residual/skip add, end conv).  Slow by design — this is the
// <= 1e-5 parity path the BF16 tensor-core kernels are judged against on the GPU itself, and the
// generic GEMM behind the first-cut STFT / mel / upsample ops.
#include "common.cuh"

#include <cuda_bf16.h>

namespace wgb {

// ------------------------------------------------------------------------------------ SGEMM
// C[b][m][n] (+)= sum_k A[b][m + shift][k] * W[n][k] + bias[n]; rows with m + shift outside [0, M)
// read as zero (== the conv's zero padding when a dilated tap is expressed as a row shift).
struct SgemmParams {
    const float* A;
    const float* W;
    const float* bias;
    void* C;
    int M, N, K;
    long long lda, a_batch, ldw, ldc, c_batch;
    int shift, accumulate;
};

constexpr int SG_BM = 128, SG_BN = 64, SG_BK = 16;

template <typename OutT>
__global__ void __launch_bounds__(256) sgemm_nt_kernel(const SgemmParams p) {
    __shared__ __align__(16) float As[SG_BK][SG_BM + 4];
    __shared__ __align__(16) float Ws[SG_BK][SG_BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
    const float* A = p.A + static_cast<long long>(blockIdx.z) * p.a_batch;

    const int a_row = tid >> 1, a_k = (tid & 1) * 8;
    const int w_row = tid >> 2, w_k = (tid & 3) * 4;
    const int am = m0 + a_row, asrc = am + p.shift;
    const bool a_ok = am < p.M && asrc >= 0 && asrc < p.M;
    const bool w_ok = (n0 + w_row) < p.N;
    const float* a_ptr = A + static_cast<long long>(asrc) * p.lda;
    const float* w_ptr = p.W + static_cast<long long>(n0 + w_row) * p.ldw;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += SG_BK) {
        float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, w0 = a0;
        if (a_ok && k0 + a_k < p.K) a0 = *reinterpret_cast<const float4*>(a_ptr + k0 + a_k);
        if (a_ok && k0 + a_k + 4 < p.K) a1 = *reinterpret_cast<const float4*>(a_ptr + k0 + a_k + 4);
        if (w_ok && k0 + w_k < p.K) w0 = *reinterpret_cast<const float4*>(w_ptr + k0 + w_k);
        __syncthreads();
        As[a_k + 0][a_row] = a0.x; As[a_k + 1][a_row] = a0.y; As[a_k + 2][a_row] = a0.z; As[a_k + 3][a_row] = a0.w;
        As[a_k + 4][a_row] = a1.x; As[a_k + 5][a_row] = a1.y; As[a_k + 6][a_row] = a1.z; As[a_k + 7][a_row] = a1.w;
        Ws[w_k + 0][w_row] = w0.x; Ws[w_k + 1][w_row] = w0.y; Ws[w_k + 2][w_row] = w0.z; Ws[w_k + 3][w_row] = w0.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            const float4 x0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 x1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float xa[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const float wa[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], wa[j], acc[i][j]);
        }
    }

    OutT* C = static_cast<OutT*>(p.C) + static_cast<long long>(blockIdx.z) * p.c_batch;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float v = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
            OutT* dst = C + static_cast<long long>(m) * p.ldc + n;
            if constexpr (sizeof(OutT) == 4) {
                if (p.accumulate) v += *dst;
                *dst = v;
            } else {
                *dst = __float2bfloat16_rn(v);
            }
        }
    }
}

int sgemm_nt(const float* A, const float* W, const float* bias, void* C, int out_bf16, int batch, int M, int N, int K,
             long long lda, long long a_batch, long long ldw, long long ldc, long long c_batch, int shift,
             int accumulate, cudaStream_t stream) {
    WGB_REQUIRE(A && W && C, "null pointer");
    WGB_REQUIRE(batch > 0 && M > 0 && N > 0 && K > 0, "bad GEMM shape %d x %d x %d (batch %d)", M, N, K, batch);
    WGB_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldw % 4 == 0 && a_batch % 4 == 0, "K, lda, ldw, a_batch must be multiples of 4");
    WGB_REQUIRE(!(out_bf16 && accumulate), "accumulate needs an fp32 output");
    WGB_REQUIRE(batch <= 65535 && ceil_div(M, SG_BM) <= 65535, "grid too large");
    SgemmParams p{A, W, bias, C, M, N, K, lda, a_batch, ldw, ldc, c_batch, shift, accumulate};
    dim3 grid(ceil_div(N, SG_BN), ceil_div(M, SG_BM), batch);
    if (out_bf16)
        sgemm_nt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
    else
        sgemm_nt_kernel<float><<<grid, 256, 0, stream>>>(p);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ pointwise
// acts = tanh(u[:, :C]) * sigmoid(u[:, C:])   (reference glow.py:33-40), channels-last rows.
__global__ void gate_f32_kernel(const float* __restrict__ u, float* __restrict__ acts, long long rows, int n_ch) {
    const long long total = rows * n_ch;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / n_ch;
        const int c = static_cast<int>(i - r * n_ch);
        const float a = u[r * 2 * n_ch + c], b = u[r * 2 * n_ch + n_ch + c];
        acts[i] = tanhf(a) * (1.f / (1.f + expf(-b)));
    }
}

int gate_f32(const float* u, float* acts, long long rows, int n_ch, cudaStream_t stream) {
    WGB_REQUIRE(u && acts && rows > 0 && n_ch > 0, "bad arguments");
    const long long total = rows * n_ch;
    const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    gate_f32_kernel<<<grid, 256, 0, stream>>>(u, acts, rows, n_ch);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// The reference's module-level function on its own layout: input_a, input_b [B, 2C, T] channels-first,
// out[b, c, t] = tanh((a+b)[b, c, t]) * sigmoid((a+b)[b, C + c, t])   (glow.py:33-40).  t is the fastest index for
// reads and writes alike, so a flat loop over (b, c, t) is fully coalesced.
__global__ void fused_add_tanh_sigmoid_multiply_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ out, int n_ch, int T, long long total) {
    const long long ct = static_cast<long long>(n_ch) * T;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long bi = i / ct;
        const long long rem = i - bi * ct;                  // c * T + t
        const long long it = bi * 2 * ct + rem, is = it + ct;
        const float ut = a[it] + b[it], us = a[is] + b[is];
        out[i] = tanhf(ut) * (1.f / (1.f + expf(-us)));
    }
}

int fused_add_tanh_sigmoid_multiply(const float* a, const float* b, float* out, int batch, int n_ch, int T,
                                    cudaStream_t stream) {
    WGB_REQUIRE(a && b && out && batch > 0 && n_ch > 0 && T > 0, "bad arguments");
    const long long total = static_cast<long long>(batch) * n_ch * T;
    const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    fused_add_tanh_sigmoid_multiply_kernel<<<grid, 256, 0, stream>>>(a, b, out, n_ch, T, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// in-place tanh (act 1) / relu (act 2) with the accurate libm functions: FP32 validation path of the Postnet
__global__ void act_f32_kernel(float* __restrict__ x, long long n, int act) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        x[i] = act == 1 ? tanhf(x[i]) : fmaxf(x[i], 0.f);
}

int act_f32(float* x, long long n, int act, cudaStream_t stream) {
    WGB_REQUIRE(x && n > 0 && (act == 1 || act == 2), "bad arguments");
    const int grid = static_cast<int>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
    act_f32_kernel<<<grid, 256, 0, stream>>>(x, n, act);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// has_res: h += rs[:, :C], skip (+)= rs[:, C:]; else skip (+)= rs (last layer)   (glow.py:165-174)
__global__ void res_skip_f32_kernel(const float* __restrict__ rs, float* __restrict__ h, float* __restrict__ skip,
                                    long long rows, int n_ch, int has_res, int first) {
    const long long total = rows * n_ch;
    const int ld = has_res ? 2 * n_ch : n_ch;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / n_ch;
        const int c = static_cast<int>(i - r * n_ch);
        float s;
        if (has_res) {
            h[i] += rs[r * ld + c];
            s = rs[r * ld + n_ch + c];
        } else {
            s = rs[r * ld + c];
        }
        skip[i] = first ? s : skip[i] + s;
    }
}

int res_skip_f32(const float* rs, float* h, float* skip, long long rows, int n_ch, int has_res, int first,
                 cudaStream_t stream) {
    WGB_REQUIRE(rs && h && skip && rows > 0 && n_ch > 0, "bad arguments");
    const long long total = rows * n_ch;
    const int grid = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
    res_skip_f32_kernel<<<grid, 256, 0, stream>>>(rs, h, skip, rows, n_ch, has_res, first);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace wgb
