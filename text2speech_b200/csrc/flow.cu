// Memory-bound kernels of the flow: noise/state layout changes, WN.start 1x1 conv, invertible 1x1
// conv (forward direction), the FP32-mode end conv + affine coupling, and the upsample im2col.
// The flow state is one fp32 buffer x[B, T, 8] (channels-last == final audio layout, sample 8t+c);
// flow k works on the LAST C_k = 2*n_half channels, so early outputs / noise injections
// (reference glow.py:229-231, :284-289) never move data.
#include "common.cuh"

#include <cuda_bf16.h>

namespace wgb {

static inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 32;
    return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

// ------------------------------------------------------------------------------------ layout
// x[b,t,c] = sigma * z[b,c,t]    (z is host-supplied noise laid out like WaveGlow.forward's output)
__global__ void flow_from_z_kernel(const float* __restrict__ z, float* __restrict__ x, int T, long long total,
                                   float sigma) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long bt = i >> 3;
        const int c = static_cast<int>(i & 7);
        const long long b = bt / T, t = bt - b * T;
        x[i] = sigma * z[(b * 8 + c) * T + t];
    }
}
// z[b,c,t] = x[b,t,c]
__global__ void flow_to_z_kernel(const float* __restrict__ x, float* __restrict__ z, int T, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long t = i % T;
        const long long bc = i / T;
        const long long b = bc >> 3;
        const int c = static_cast<int>(bc & 7);
        z[i] = x[(b * T + t) * 8 + c];
    }
}

int flow_from_z(const float* z, float* x, int batch, int T, float sigma, cudaStream_t stream) {
    WGB_REQUIRE(z && x && batch > 0 && T > 0, "bad arguments");
    const long long total = static_cast<long long>(batch) * T * 8;
    flow_from_z_kernel<<<grid_for(total, 256), 256, 0, stream>>>(z, x, T, total, sigma);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}
int flow_to_z(const float* x, float* z, int batch, int T, cudaStream_t stream) {
    WGB_REQUIRE(z && x && batch > 0 && T > 0, "bad arguments");
    const long long total = static_cast<long long>(batch) * T * 8;
    flow_to_z_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, z, T, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ PCM output
// out[i] = (int16) trunc(x[i] * scale), saturating (waveglow/inference.py:58-62: audio * MAX_WAV_VALUE ->
// astype('int16')).  HBM-bound: 4 B in, 2 B out per sample; 8 samples (32 B in, 16 B out) per thread step.
__global__ void audio_to_int16_kernel(const float* __restrict__ x, short* __restrict__ out, long long n, float scale) {
    const long long n8 = n >> 3;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        short s[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = static_cast<short>(__float2int_rz(fminf(fmaxf(v[j] * scale, -32768.f), 32767.f)));
        reinterpret_cast<uint4*>(out)[i] = *reinterpret_cast<const uint4*>(s);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
        const long long i = (n8 << 3) + threadIdx.x;
        out[i] = static_cast<short>(__float2int_rz(fminf(fmaxf(x[i] * scale, -32768.f), 32767.f)));
    }
}

int audio_to_int16(const float* x, void* out, long long n, float scale, cudaStream_t stream) {
    WGB_REQUIRE(x && out && n > 0, "bad arguments");
    WGB_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0, "buffers must be 16 B aligned");
    audio_to_int16_kernel<<<grid_for((n + 7) / 8, 256), 256, 0, stream>>>(x, static_cast<short*>(out), n, scale);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ convinv (forward)
// x[., 8-C:] <- W x[., 8-C:]   (Invertible1x1Conv.forward, glow.py:100-101); W as [8][8] row-major.
__global__ void flow_mix_kernel(float* __restrict__ x, const float* __restrict__ w, long long rows, int C) {
    __shared__ float sw[64];
    if (threadIdx.x < 64) sw[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const int base = 8 - C;
    for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows;
         r += static_cast<long long>(gridDim.x) * blockDim.x) {
        float v[8], o[8];
        *reinterpret_cast<float4*>(&v[0]) = *reinterpret_cast<const float4*>(x + r * 8);
        *reinterpret_cast<float4*>(&v[4]) = *reinterpret_cast<const float4*>(x + r * 8 + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (i >= base && c >= base) acc = fmaf(sw[(i - base) * 8 + (c - base)], v[c], acc);
            o[i] = i >= base ? acc : v[i];
        }
        *reinterpret_cast<float4*>(x + r * 8) = *reinterpret_cast<const float4*>(&o[0]);
        *reinterpret_cast<float4*>(x + r * 8 + 4) = *reinterpret_cast<const float4*>(&o[4]);
    }
}

int flow_mix(float* x, const float* w, long long rows, int C, cudaStream_t stream) {
    WGB_REQUIRE(x && w && rows > 0 && C >= 2 && C <= 8, "bad arguments");
    flow_mix_kernel<<<grid_for(rows, 256), 256, 0, stream>>>(x, w, rows, C);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ x_stack (first-layer operand)
// Row t of utterance b (pitch out_batch_rows): 64 bf16 =
//   [ 0..11] hi(a0[t-1][0..3]), hi(a0[t][0..3]), hi(a0[t+1][0..3])      a0 = x[., 8-2*n_half .. 8-n_half), zero beyond n_half
//   [12..23] lo parts of the same 12 values (x = hi + lo, both bf16)
//   [24..35] hi parts again (they meet the lo part of the weight)
//   [36..38] 1.0 where tap t-1 / t / t+1 lies inside [0, T) (carries W_in0,tap b_start), [39..41] the same again
//   [42..63] zero
// the K = 64 operand of wgb_tc2_wn_gate_mel0: WN.start folded into in_layers[0] (glow.py:156,160).
__global__ void x_stack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int T, long long rows,
                               long long out_batch_rows, int n_half) {
    pdl_launch_dependents();
    pdl_wait();                       // x comes from the kernel before this one
    const int base = 8 - 2 * n_half;
    for (long long r = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; r < rows;
         r += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = r / T;
        const int t = static_cast<int>(r - b * T);
        __align__(16) __nv_bfloat16 row[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) row[i] = __float2bfloat16_rn(0.f);
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
            const int tt = t + tap - 1;
            const bool in = tt >= 0 && tt < T;
            float xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = 0.f;
            if (in) {
                *reinterpret_cast<float4*>(&xv[0]) = *reinterpret_cast<const float4*>(x + (r + tap - 1) * 8);
                *reinterpret_cast<float4*>(&xv[4]) = *reinterpret_cast<const float4*>(x + (r + tap - 1) * 8 + 4);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) v = (i == base + c && c < n_half) ? xv[i] : v;
                const __nv_bfloat16 hi = __float2bfloat16_rn(v);
                row[tap * 4 + c] = hi;
                row[12 + tap * 4 + c] = __float2bfloat16_rn(v - __bfloat162float(hi));
                row[24 + tap * 4 + c] = hi;
            }
            row[36 + tap] = __float2bfloat16_rn(in ? 1.f : 0.f);
            row[39 + tap] = row[36 + tap];
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (b * out_batch_rows + t) * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = reinterpret_cast<const uint4*>(row)[i];
    }
}

int x_stack(const float* x, void* out, int batch, int T, long long out_batch_rows, int n_half, cudaStream_t stream) {
    WGB_REQUIRE(x && out && batch > 0 && T > 0 && out_batch_rows >= T && n_half >= 1 && n_half <= 4, "bad arguments");
    const long long rows = static_cast<long long>(batch) * T;
    WGB_CUDA_TRY(launch_pdl(x_stack_kernel, grid_for(rows, 128), 128, 0, stream, x, static_cast<__nv_bfloat16*>(out), T, rows,
                            out_batch_rows, n_half));
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ WN.start
// h[r, c] = b[c] + sum_j W[c, j] * x[r, 8 - 2*n_half + j]   (glow.py:156).  HBM-bound on the 1 KB/row
// store: each thread owns 8 fixed channels (weights + bias live in registers for the whole kernel) and
// walks rows, so a row-lane of n_ch/8 threads writes one contiguous n_ch*sizeof(OutT) row per iteration.
template <typename OutT>
__global__ void __launch_bounds__(256) wn_start_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, OutT* __restrict__ h,
                                                       long long rows, int n_ch, int n_half, int T,
                                                       long long h_batch_rows) {
    const int groups = n_ch >> 3;                      // threads per row
    const int lanes = blockDim.x / groups;             // rows per block iteration
    const int g = threadIdx.x % groups, rl = threadIdx.x / groups;
    if (rl >= lanes) return;
    const int c0 = g << 3;
    const int base = 8 - 2 * n_half;
    float wr[8][4], br[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        br[c] = bias[c0 + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) wr[c][j] = j < n_half ? w[(c0 + c) * n_half + j] : 0.f;
    }
    for (long long r = static_cast<long long>(blockIdx.x) * lanes + rl; r < rows;
         r += static_cast<long long>(gridDim.x) * lanes) {
        const float4 lo = *reinterpret_cast<const float4*>(x + r * 8);
        const float4 hi = *reinterpret_cast<const float4*>(x + r * 8 + 4);
        const float xv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        float a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) v = (i == base + j && j < n_half) ? xv[i] : v;
            a[j] = v;
        }
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c)
            o[c] = fmaf(wr[c][3], a[3], fmaf(wr[c][2], a[2], fmaf(wr[c][1], a[1], fmaf(wr[c][0], a[0], br[c]))));
        const long long hb = r / T;
        const long long hr = hb * h_batch_rows + (r - hb * T);           // row of h: utterances h_batch_rows apart
        if constexpr (sizeof(OutT) == 4) {
            float4* d = reinterpret_cast<float4*>(h + hr * n_ch + c0);
            d[0] = make_float4(o[0], o[1], o[2], o[3]);
            d[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(o[4], o[5]), p3 = __floats2bfloat162_rn(o[6], o[7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            pk.z = *reinterpret_cast<uint32_t*>(&p2);
            pk.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(h + hr * n_ch + c0) = pk;
        }
    }
}

int wn_start(const float* x, const float* w, const float* bias, void* h, int out_bf16, long long rows, int n_ch,
             int n_half, int T, long long h_batch_rows, cudaStream_t stream) {
    WGB_REQUIRE(x && w && bias && h && rows > 0, "bad arguments");
    if (T <= 0) { T = static_cast<int>(rows < 0x7fffffffLL ? rows : 0x7fffffffLL); h_batch_rows = T; }   // dense h
    WGB_REQUIRE(h_batch_rows >= T, "h_batch_rows must be >= T");
    WGB_REQUIRE(n_ch % 8 == 0 && n_ch / 8 <= 256 && n_half >= 1 && n_half <= 4,
                "n_ch %% 8 == 0, n_ch <= 2048 and n_half in 1..4 required");
    const int lanes = 256 / (n_ch / 8);
    long long blocks = (rows + lanes - 1) / lanes;
    const int grid = static_cast<int>(blocks < 148LL * 8 ? blocks : 148LL * 8);
    if (out_bf16)
        wn_start_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(x, w, bias, static_cast<__nv_bfloat16*>(h), rows, n_ch,
                                                                 n_half, T, h_batch_rows);
    else
        wn_start_kernel<float><<<grid, 256, 0, stream>>>(x, w, bias, static_cast<float*>(h), rows, n_ch, n_half, T,
                                                         h_batch_rows);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ end from the skip accumulator
// out = sum of the four per-pass slots of skip_acc + b_end  (= WN.end(sum_i skip_i), accumulated by the gate kernels),
// then the affine coupling + W^-1 (infer, glow.py:277-282) or the forward coupling + log_s (:241-246), then -- infer
// only, optional -- WN.start of the next flow (glow.py:156) on the updated row.  Each row is handled by a lane group of
// n_ch/8 = 64 threads: all of them evaluate the (tiny) coupling redundantly from broadcast loads, the first one writes
// x / log_s, and every thread writes its 8 channels of h_next, so the 1 KB row store is contiguous.
template <int NHALF, int DIR>
__global__ void __launch_bounds__(256, 3) end_from_acc_kernel(const float* __restrict__ acc, const float* __restrict__ b_end,
                                                           float* __restrict__ x, const float* __restrict__ w_mix,
                                                           float* __restrict__ log_s, long long rows, int T,
                                                           const float* __restrict__ nw, const float* __restrict__ nbias,
                                                           int next_n_half, __nv_bfloat16* __restrict__ h_next,
                                                           long long h_batch_rows) {
    constexpr int kCh = 512, kGroups = kCh / 8, kLanes = 256 / kGroups;
    constexpr int C = 2 * NHALF, BASE = 8 - C;
    const int g = threadIdx.x % kGroups, rl = threadIdx.x / kGroups;
    const int c0 = g << 3;
    __shared__ float s_wm[64], s_be[8];                 // W^-1 and b_end: broadcast reads, keeps registers for wr/br
    if (threadIdx.x < 64) s_wm[threadIdx.x] = (DIR == 0) ? w_mix[threadIdx.x] : 0.f;
    if (threadIdx.x < 8) s_be[threadIdx.x] = b_end[threadIdx.x];
    __syncthreads();
    float wr[8][4], br[8];
    if (h_next) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            br[c] = nbias[c0 + c];
#pragma unroll
            for (int j = 0; j < 4; ++j) wr[c][j] = j < next_n_half ? nw[(c0 + c) * next_n_half + j] : 0.f;
        }
    }
    const int nb = 8 - 2 * next_n_half;
    for (long long r0 = static_cast<long long>(blockIdx.x) * kLanes; r0 < rows;
         r0 += static_cast<long long>(gridDim.x) * kLanes) {           // block-uniform trip count (barrier inside)
        const bool valid = r0 + rl < rows;
        const long long r = valid ? r0 + rl : rows - 1;
        float out[8];
        {
            const float4 a0 = *reinterpret_cast<const float4*>(acc + r * 8), a1 = *reinterpret_cast<const float4*>(acc + r * 8 + 4);
            out[0] = a0.x; out[1] = a0.y; out[2] = a0.z; out[3] = a0.w;
            out[4] = a1.x; out[5] = a1.y; out[6] = a1.z; out[7] = a1.w;
        }
#pragma unroll
        for (int s = 1; s < 4; ++s) {
            const float* ap = acc + (static_cast<long long>(s) * rows + r) * 8;
            const float4 a0 = *reinterpret_cast<const float4*>(ap), a1 = *reinterpret_cast<const float4*>(ap + 4);
            out[0] += a0.x; out[1] += a0.y; out[2] += a0.z; out[3] += a0.w;
            out[4] += a1.x; out[5] += a1.y; out[6] += a1.z; out[7] += a1.w;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] += s_be[j];
        float xv[8];
        *reinterpret_cast<float4*>(&xv[0]) = *reinterpret_cast<const float4*>(x + r * 8);
        *reinterpret_cast<float4*>(&xv[4]) = *reinterpret_cast<const float4*>(x + r * 8 + 4);
        const long long b = r / T;
        const long long t = r - b * T;
        __syncthreads();          // every thread of the row's lane group has read the old x before its first thread writes
        if constexpr (DIR == 0) {
            float xin[C];
#pragma unroll
            for (int j = 0; j < NHALF; ++j) {
                xin[j] = xv[BASE + j];
                xin[NHALF + j] = (xv[BASE + NHALF + j] - out[j]) * expf(-out[NHALF + j]);
            }
#pragma unroll
            for (int i = 0; i < C; ++i) {
                float a = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) a = fmaf(s_wm[i * 8 + c], xin[c], a);
                xv[BASE + i] = a;
            }
        } else {
#pragma unroll
            for (int j = 0; j < NHALF; ++j) {
                const float ls = out[NHALF + j];
                xv[BASE + NHALF + j] = expf(ls) * xv[BASE + NHALF + j] + out[j];
                if (g == 0 && valid) log_s[(b * NHALF + j) * T + t] = ls;
            }
        }
        if (g == 0 && valid) {
            *reinterpret_cast<float4*>(x + r * 8) = *reinterpret_cast<const float4*>(&xv[0]);
            *reinterpret_cast<float4*>(x + r * 8 + 4) = *reinterpret_cast<const float4*>(&xv[4]);
        }
        if (DIR == 0 && h_next && valid) {
            float a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) v = (i == nb + j) ? xv[i] : v;
                a[j] = v;
            }
            float o[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                o[c] = fmaf(wr[c][3], a[3], fmaf(wr[c][2], a[2], fmaf(wr[c][1], a[1], fmaf(wr[c][0], a[0], br[c]))));
            __nv_bfloat162 p0 = __floats2bfloat162_rn(o[0], o[1]), p1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(o[4], o[5]), p3 = __floats2bfloat162_rn(o[6], o[7]);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&p0);
            pk.y = *reinterpret_cast<uint32_t*>(&p1);
            pk.z = *reinterpret_cast<uint32_t*>(&p2);
            pk.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(h_next + (b * h_batch_rows + t) * kCh + c0) = pk;
        }
    }
}

int end_from_acc(const float* skip_acc, const float* b_end, float* x, const float* w_mix, float* log_s, int batch, int T,
                 int n_half, int direction, const float* next_w_start, const float* next_b_start, int next_n_half,
                 void* h_next, long long h_next_batch_rows, cudaStream_t stream) {
    WGB_REQUIRE(skip_acc && b_end && x && batch > 0 && T > 0, "bad arguments");
    WGB_REQUIRE(n_half >= 1 && n_half <= 4, "n_half must be in 1..4 (got %d)", n_half);
    WGB_REQUIRE(direction == 0 || direction == 1, "direction must be 0 (infer) or 1 (forward)");
    WGB_REQUIRE(direction == 1 ? log_s != nullptr : w_mix != nullptr, "missing log_s / w_mix for this direction");
    if (h_next) {
        WGB_REQUIRE(direction == 0, "the fused WN.start of the next flow exists for the infer direction only");
        WGB_REQUIRE(next_w_start && next_b_start && next_n_half >= 1 && next_n_half <= 4 && h_next_batch_rows >= T,
                    "bad next-flow start arguments");
    } else {
        next_n_half = 1;
    }
    const long long rows = static_cast<long long>(batch) * T;
    const long long blocks = (rows + 3) / 4;
    const int grid = static_cast<int>(blocks < 148LL * 8 ? blocks : 148LL * 8);
    __nv_bfloat16* hn = static_cast<__nv_bfloat16*>(h_next);
#define WGB_EFA_CASE(NH)                                                                                              \
    case NH:                                                                                                          \
        if (direction == 0)                                                                                           \
            end_from_acc_kernel<NH, 0><<<grid, 256, 0, stream>>>(skip_acc, b_end, x, w_mix, log_s, rows, T, next_w_start, \
                                                                 next_b_start, next_n_half, hn, h_next_batch_rows);  \
        else                                                                                                          \
            end_from_acc_kernel<NH, 1><<<grid, 256, 0, stream>>>(skip_acc, b_end, x, w_mix, log_s, rows, T, next_w_start, \
                                                                 next_b_start, next_n_half, hn, h_next_batch_rows);  \
        break;
    switch (n_half) {
        WGB_EFA_CASE(1)
        WGB_EFA_CASE(2)
        WGB_EFA_CASE(3)
        WGB_EFA_CASE(4)
    }
#undef WGB_EFA_CASE
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ FP32 end + coupling
// One warp per row: out = W_end * skip + b_end (glow.py:175), then the affine coupling and, for
// infer, the inverse 1x1 conv (glow.py:277-282); forward writes log_s (glow.py:241-246).
__global__ void end_coupling_f32_kernel(const float* __restrict__ skip, const float* __restrict__ w_end,
                                        const float* __restrict__ b_end, float* __restrict__ x,
                                        const float* __restrict__ w_mix, float* __restrict__ log_s, int batch, int T,
                                        int n_ch, int n_half, int direction) {
    const int lane = threadIdx.x & 31;
    const long long rows = static_cast<long long>(batch) * T;
    const long long warp_id = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long r = warp_id; r < rows; r += n_warps) {
        float out[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) out[j] = 0.f;
        for (int c = lane; c < n_ch; c += 32) {
            const float s = skip[r * n_ch + c];
#pragma unroll
            for (int j = 0; j < 8; ++j) out[j] = fmaf(s, w_end[c * 8 + j], out[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) out[j] += __shfl_xor_sync(0xffffffffu, out[j], o);
            out[j] += b_end[j];
        }
        if (lane == 0) {
            const int C = 2 * n_half, base = 8 - C;
            float xv[8], xin[8];
            for (int i = 0; i < 8; ++i) xv[i] = x[r * 8 + i];
            if (direction == 0) {
                for (int j = 0; j < n_half; ++j) {
                    xin[j] = xv[base + j];
                    xin[n_half + j] = (xv[base + n_half + j] - out[j]) * expf(-out[n_half + j]);
                }
                for (int i = 0; i < C; ++i) {
                    float acc = 0.f;
                    for (int c = 0; c < C; ++c) acc = fmaf(w_mix[i * 8 + c], xin[c], acc);
                    x[r * 8 + base + i] = acc;
                }
            } else {
                const long long b = r / T, t = r - b * T;
                for (int j = 0; j < n_half; ++j) {
                    const float ls = out[n_half + j];
                    x[r * 8 + base + n_half + j] = expf(ls) * xv[base + n_half + j] + out[j];
                    log_s[(b * n_half + j) * T + t] = ls;
                }
            }
        }
    }
}

int end_coupling_f32(const float* skip, const float* w_end, const float* b_end, float* x, const float* w_mix,
                     float* log_s, int batch, int T, int n_ch, int n_half, int direction, cudaStream_t stream) {
    WGB_REQUIRE(skip && w_end && b_end && x && batch > 0 && T > 0, "bad arguments");
    WGB_REQUIRE(n_half >= 1 && n_half <= 4, "n_half must be in 1..4");
    WGB_REQUIRE(direction == 1 ? log_s != nullptr : w_mix != nullptr, "missing log_s / w_mix for this direction");
    const long long rows = static_cast<long long>(batch) * T;
    end_coupling_f32_kernel<<<grid_for(rows * 32, 256), 256, 0, stream>>>(skip, w_end, b_end, x, w_mix, log_s, batch, T,
                                                                         n_ch, n_half, direction);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ upsample im2col
// ConvTranspose1d(80,80,1024,stride 256) (glow.py:183-185) as a GEMM: row q of A holds the four
// frames that touch output samples [256q, 256q+256):  A[b, q, j*ld_tap + c] = mel[b, c, q - j].
template <typename OutT>
__global__ void upsample_im2col_kernel(const float* __restrict__ mel, OutT* __restrict__ a, int n_mel, int F, int taps,
                                       int ld_tap, long long total) {
    const int row_len = taps * ld_tap;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long bq = i / row_len;
        const int k = static_cast<int>(i - bq * row_len);
        const int j = k / ld_tap, c = k - j * ld_tap;
        const long long b = bq / F;
        const int q = static_cast<int>(bq - b * F);
        float v = 0.f;
        if (c < n_mel && q - j >= 0) v = mel[(b * n_mel + c) * F + (q - j)];
        if constexpr (sizeof(OutT) == 4) a[i] = v;
        else a[i] = __float2bfloat16_rn(v);
    }
}

int upsample_im2col(const float* mel, void* a, int out_bf16, int batch, int n_mel, int F, int taps, int ld_tap,
                    cudaStream_t stream) {
    WGB_REQUIRE(mel && a && batch > 0 && n_mel > 0 && F > 0 && taps > 0 && ld_tap >= n_mel, "bad arguments");
    const long long total = static_cast<long long>(batch) * F * taps * ld_tap;
    if (out_bf16)
        upsample_im2col_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, stream>>>(
            mel, static_cast<__nv_bfloat16*>(a), n_mel, F, taps, ld_tap, total);
    else
        upsample_im2col_kernel<float><<<grid_for(total, 256), 256, 0, stream>>>(mel, static_cast<float*>(a), n_mel, F,
                                                                                taps, ld_tap, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// ------------------------------------------------------------------------------------ casts
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}
int cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t stream) {
    WGB_REQUIRE(src && dst && n > 0, "bad arguments");
    cast_bf16_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace wgb
