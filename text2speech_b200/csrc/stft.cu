// Conv-basis STFT / mel / denoiser glue kernels (reference utils/stft.py, utils/layers.py:63-79,
// waveglow/denoiser.py:35-40).  The two dense-basis contractions themselves run through the GEMM
// entry points; everything here is the memory-bound work around them.
//
// Layouts: padded signal ypad[B, ld_pad]; spectrum spec[B, F, 2*cp] channels-last with Re in
// [0, cutoff) and Im in [cp, cp+cutoff) (cp = cutoff rounded up to 4, pad columns are zero because
// the packed basis rows are zero); reference-facing outputs are channels-first [B, cutoff, F].
#include "common.cuh"

#include <cuda_bf16.h>

namespace wgb {

static inline int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = 148LL * 32;
    return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

// reflect-pad L/2 samples each side (stft.py:79-83): ypad[b, i] = y[b, reflect(i - half)]
__global__ void reflect_pad_kernel(const float* __restrict__ y, float* __restrict__ ypad, int N, int half,
                                   long long ld_pad, long long total) {
    const int padded = N + 2 * half;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / ld_pad;
        const int p = static_cast<int>(i - b * ld_pad);
        float v = 0.f;
        if (p < padded) {
            int s = p - half;
            if (s < 0) s = -s;
            if (s >= N) s = 2 * (N - 1) - s;
            v = y[b * N + s];
        }
        ypad[i] = v;
    }
}

int stft_reflect_pad(const float* y, float* ypad, int batch, int N, int half, long long ld_pad, cudaStream_t stream) {
    WGB_REQUIRE(y && ypad && batch > 0 && N > half, "reflect padding needs N > filter_length/2 (N=%d)", N);
    WGB_REQUIRE(ld_pad >= N + 2 * half && ld_pad % 4 == 0, "bad padded stride");
    const long long total = static_cast<long long>(batch) * ld_pad;
    reflect_pad_kernel<<<grid_for(total, 256), 256, 0, stream>>>(y, ypad, N, half, ld_pad, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): operands of the split-bf16 tensor-core GEMMs
__device__ __forceinline__ void split_store(float v, __nv_bfloat16* hi, __nv_bfloat16* lo, long long i) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// range_flag (optional): set to 1 when a sample lies outside [-1, 1] or is NaN -- the two device-wide reductions and host
// syncs of TacotronSTFT.mel_spectrogram's input asserts (layers.py:72-73) folded into this pass
__global__ void reflect_pad_split_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo, int N, int half, long long ld_pad,
                                         long long total, int* __restrict__ range_flag) {
    const int padded = N + 2 * half;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / ld_pad;
        const int p = static_cast<int>(i - b * ld_pad);
        float v = 0.f;
        if (p < padded) {
            int s = p - half;
            if (s < 0) s = -s;
            if (s >= N) s = 2 * (N - 1) - s;
            v = y[b * N + s];
        }
        if (range_flag && !(v >= -1.f && v <= 1.f)) *range_flag = 1;
        split_store(v, hi, lo, i);
    }
}

// 8 outputs per thread: two float4 loads in the interior (16 B aligned when N % 4 == 0 and half % 8 == 0), scalar
// reflected reads at the two edges, one 16 B store per output array
__global__ void reflect_pad_split8_kernel(const float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                                          __nv_bfloat16* __restrict__ lo, int N, int half, long long ld_pad,
                                          long long total8, int* __restrict__ range_flag) {
    const int padded = N + 2 * half;
    const long long per_row = ld_pad >> 3;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total8;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / per_row;
        const int p0 = static_cast<int>(i - b * per_row) << 3;
        const float* row = y + b * N;
        float v[8];
        const int s0 = p0 - half;
        if (s0 >= 0 && s0 + 7 < N) {
            const float4 a = *reinterpret_cast<const float4*>(row + s0), c = *reinterpret_cast<const float4*>(row + s0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int p = p0 + j;
                int sidx = p - half;
                if (sidx < 0) sidx = -sidx;
                if (sidx >= N) sidx = 2 * (N - 1) - sidx;
                v[j] = p < padded ? row[sidx] : 0.f;
            }
        }
        if (range_flag) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 8; ++j) ok = ok && (v[j] >= -1.f && v[j] <= 1.f);
            if (!ok) *range_flag = 1;
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * j] - __low2float(hh), v[2 * j + 1] - __high2float(hh));
            h[j] = *reinterpret_cast<const uint32_t*>(&hh);
            l[j] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint4*>(hi + b * ld_pad + p0) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(lo + b * ld_pad + p0) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

int stft_reflect_pad_split(const float* y, void* hi, void* lo, int batch, int N, int half, long long ld_pad,
                           int* range_flag, cudaStream_t stream) {
    WGB_REQUIRE(y && hi && lo && batch > 0 && N > half, "reflect padding needs N > filter_length/2 (N=%d)", N);
    WGB_REQUIRE(ld_pad >= N + 2 * half && ld_pad % 8 == 0, "bad padded stride");
    if (N % 4 == 0 && half % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
        const long long total8 = static_cast<long long>(batch) * (ld_pad >> 3);
        reflect_pad_split8_kernel<<<grid_for(total8, 256), 256, 0, stream>>>(
            y, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), N, half, ld_pad, total8, range_flag);
        WGB_LAUNCH_CHECK();
        return WGB_OK;
    }
    const long long total = static_cast<long long>(batch) * ld_pad;
    reflect_pad_split_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
        y, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), N, half, ld_pad, total, range_flag);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

__global__ void split_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        split_store(src[i], hi, lo, i);
}

int split_bf16(const float* src, void* hi, void* lo, long long n, cudaStream_t stream) {
    WGB_REQUIRE(src && hi && lo && n > 0, "bad arguments");
    split_bf16_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src, static_cast<__nv_bfloat16*>(hi),
                                                          static_cast<__nv_bfloat16*>(lo), n);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// spec[B,F,2cp] -> magnitude / phase [B,cutoff,F] (stft.py:91-97) and/or channels-last mag_cl[B,F,cp]
__global__ void stft_polar_kernel(const float* __restrict__ spec, float* __restrict__ mag, float* __restrict__ phase,
                                  float* __restrict__ mag_cl, int F, int cutoff, int cp, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        // i enumerates (b, k, f) with f fastest so the channels-first stores coalesce
        const int f = static_cast<int>(i % F);
        const long long bk = i / F;
        const int k = static_cast<int>(bk % cp);
        const long long b = bk / cp;
        const float* s = spec + (b * F + f) * 2 * cp;
        const float re = s[k], im = s[cp + k];
        const float m = sqrtf(re * re + im * im);
        if (mag_cl) mag_cl[(b * F + f) * cp + k] = m;
        if (k < cutoff) {
            if (mag) mag[(b * cutoff + k) * F + f] = m;
            if (phase) phase[(b * cutoff + k) * F + f] = atan2f(im, re);
        }
    }
}

// channels-last only (the mel path): mag_cl[row, k] = |spec[row, k] + i spec[row, cp + k]|, k fastest so that both the
// reads and the writes are contiguous; 4 bins per thread (cp % 4 == 0)
__global__ void stft_mag_cl_kernel(const float* __restrict__ spec, float* __restrict__ mag_cl, int cp, long long total4) {
    const int cp4 = cp >> 2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / cp4;
        const int k4 = static_cast<int>(i - row * cp4);
        const float4 re = reinterpret_cast<const float4*>(spec + row * 2 * cp)[k4];
        const float4 im = reinterpret_cast<const float4*>(spec + row * 2 * cp + cp)[k4];
        reinterpret_cast<float4*>(mag_cl + row * cp)[k4] =
            make_float4(sqrtf(re.x * re.x + im.x * im.x), sqrtf(re.y * re.y + im.y * im.y),
                        sqrtf(re.z * re.z + im.z * im.z), sqrtf(re.w * re.w + im.w * im.w));
    }
}

int stft_polar(const float* spec, float* mag, float* phase, float* mag_cl, int batch, int F, int cutoff, int cp,
               cudaStream_t stream) {
    WGB_REQUIRE(spec && batch > 0 && F > 0 && cutoff > 0 && cp >= cutoff, "bad arguments");
    if (mag_cl && !mag && !phase && cp % 4 == 0) {
        const long long total4 = static_cast<long long>(batch) * F * (cp >> 2);
        stft_mag_cl_kernel<<<grid_for(total4, 256), 256, 0, stream>>>(spec, mag_cl, cp, total4);
        WGB_LAUNCH_CHECK();
        return WGB_OK;
    }
    const long long total = static_cast<long long>(batch) * cp * F;
    stft_polar_kernel<<<grid_for(total, 256), 256, 0, stream>>>(spec, mag, phase, mag_cl, F, cutoff, cp, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// mel_out[b, m, f] = log(max(raw[b, f, m], clip))   (layers.py:77-78, audio_processing.py:70-76)
__global__ void mel_log_kernel(const float* __restrict__ raw, float* __restrict__ out, int F, int n_mel, float clip,
                               long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(i % F);
        const long long bm = i / F;
        const int m = static_cast<int>(bm % n_mel);
        const long long b = bm / n_mel;
        out[i] = logf(fmaxf(raw[(b * F + f) * n_mel + m], clip));
    }
}

int mel_log(const float* raw, float* out, int batch, int F, int n_mel, float clip, cudaStream_t stream) {
    WGB_REQUIRE(raw && out && batch > 0 && F > 0 && n_mel > 0, "bad arguments");
    const long long total = static_cast<long long>(batch) * n_mel * F;
    mel_log_kernel<<<grid_for(total, 256), 256, 0, stream>>>(raw, out, F, n_mel, clip, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// Spectral subtraction without trigonometry (denoiser.py:36-39 + stft.py:102-103):
// mag' = max(mag - bias*strength, 0); Re,Im *= mag'/mag  (cos(atan2(im,re)) = re/mag; mag = 0 -> 0).
__global__ void denoise_scale_kernel(float* __restrict__ spec, const float* __restrict__ bias, float strength,
                                     int cutoff, int cp, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / cp;
        const int k = static_cast<int>(i - row * cp);
        if (k >= cutoff) continue;
        float* s = spec + row * 2 * cp;
        const float re = s[k], im = s[cp + k];
        const float m = sqrtf(re * re + im * im);
        const float m2 = fmaxf(m - bias[k] * strength, 0.f);
        const float g = m > 0.f ? m2 / m : 0.f;
        s[k] = (m > 0.f) ? re * g : m2;       // atan2(0,0) = 0 -> cos = 1, sin = 0 (m2 is 0 here anyway)
        s[cp + k] = im * g;
    }
}

int denoise_scale(float* spec, const float* bias, float strength, long long rows, int cutoff, int cp,
                  cudaStream_t stream) {
    WGB_REQUIRE(spec && bias && rows > 0 && cutoff > 0 && cp >= cutoff, "bad arguments");
    const long long total = rows * cp;
    denoise_scale_kernel<<<grid_for(total, 256), 256, 0, stream>>>(spec, bias, strength, cutoff, cp, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// Griffin-Lim projection (audio_processing.py:64-66: `_, angles = transform(signal)` then
// `inverse(magnitudes, angles)`): keep the phase of spec, impose the target magnitude, no trigonometry:
// Re,Im *= target/|X|; |X| = 0 has phase atan2(0,0) = 0 -> (target, 0).  target is [B,cutoff,F].
__global__ void spec_set_magnitude_kernel(float* __restrict__ spec, const float* __restrict__ target, int F, int cutoff,
                                          int cp, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / cp;                    // row = b * F + f
        const int k = static_cast<int>(i - row * cp);
        if (k >= cutoff) continue;
        const long long b = row / F;
        const int f = static_cast<int>(row - b * F);
        float* s = spec + row * 2 * cp;
        const float re = s[k], im = s[cp + k];
        const float m = sqrtf(re * re + im * im);
        const float tgt = target[(b * cutoff + k) * F + f];
        const float g = m > 0.f ? tgt / m : 0.f;
        s[k] = m > 0.f ? re * g : tgt;
        s[cp + k] = im * g;
    }
}

int spec_set_magnitude(float* spec, const float* target, int batch, int F, int cutoff, int cp, cudaStream_t stream) {
    WGB_REQUIRE(spec && target && batch > 0 && F > 0 && cutoff > 0 && cp >= cutoff, "bad arguments");
    const long long total = static_cast<long long>(batch) * F * cp;
    spec_set_magnitude_kernel<<<grid_for(total, 256), 256, 0, stream>>>(spec, target, F, cutoff, cp, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// Spectral subtraction fused with the hi/lo split that feeds the inverse-basis tensor-core GEMM: spec is read once and
// never written back (denoiser.py:36-39 + stft.py:102-103 + the operand split of wgb_tc_gemm_split3).
__global__ void denoise_scale_split_kernel(const float* __restrict__ spec, const float* __restrict__ bias, float strength,
                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int cutoff,
                                           int cp, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / cp;
        const int k = static_cast<int>(i - row * cp);
        const float* s = spec + row * 2 * cp;
        float re = s[k], im = s[cp + k];
        if (k < cutoff) {
            const float m = sqrtf(re * re + im * im);
            const float m2 = fmaxf(m - bias[k] * strength, 0.f);
            const float g = m > 0.f ? m2 / m : 0.f;
            re = (m > 0.f) ? re * g : m2;
            im = im * g;
        }
        split_store(re, hi, lo, row * 2 * cp + k);
        split_store(im, hi, lo, row * 2 * cp + cp + k);
    }
}

int denoise_scale_split(const float* spec, const float* bias, float strength, void* hi, void* lo, long long rows, int cutoff,
                        int cp, cudaStream_t stream) {
    WGB_REQUIRE(spec && bias && hi && lo && rows > 0 && cutoff > 0 && cp >= cutoff, "bad arguments");
    const long long total = rows * cp;
    denoise_scale_split_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
        spec, bias, strength, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), cutoff, cp, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// (magnitude, phase) [B,cutoff,F] -> spec[B,F,2cp] = [mag cos(phase) | mag sin(phase)]  (stft.py:102-103)
__global__ void stft_recombine_kernel(const float* __restrict__ mag, const float* __restrict__ phase,
                                      float* __restrict__ spec, int F, int cutoff, int cp, long long total) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(i % F);
        const long long bk = i / F;
        const int k = static_cast<int>(bk % cp);
        const long long b = bk / cp;
        float re = 0.f, im = 0.f;
        if (k < cutoff) {
            const float m = mag[(b * cutoff + k) * F + f], ph = phase[(b * cutoff + k) * F + f];
            float sn, cs;
            sincosf(ph, &sn, &cs);
            re = m * cs;
            im = m * sn;
        }
        float* s = spec + (b * F + f) * 2 * cp;
        s[k] = re;
        s[cp + k] = im;
    }
}

int stft_recombine(const float* mag, const float* phase, float* spec, int batch, int F, int cutoff, int cp,
                   cudaStream_t stream) {
    WGB_REQUIRE(mag && phase && spec && batch > 0 && F > 0 && cutoff > 0 && cp >= cutoff, "bad arguments");
    const long long total = static_cast<long long>(batch) * cp * F;
    stft_recombine_kernel<<<grid_for(total, 256), 256, 0, stream>>>(mag, phase, spec, F, cutoff, cp, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// Overlap-add of the per-frame inverse-basis outputs (the conv_transpose1d of stft.py:105-109),
// window-sum normalisation where the envelope exceeds float32 tiny, x L/hop, and the L/2 trim
// (stft.py:111-128).  The envelope is rebuilt per sample exactly like the host loop of
// audio_processing.py:45-47 (frames ascending, float32 running sum of float64 squares).  win_sq == NULL is the
// reference's window=None case: neither the envelope division nor the L/hop scale is applied (both sit inside
// ``if self.window is not None`` at stft.py:111-125).
__global__ void istft_overlap_add_kernel(const float* __restrict__ frames, const double* __restrict__ win_sq,
                                         float* __restrict__ out, int F, int L, int hop, int n_out, long long total) {
    const int half = L / 2;
    const bool windowed = win_sq != nullptr;
    const float scale = windowed ? static_cast<float>(L) / static_cast<float>(hop) : 1.f;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / n_out;
        const int n = static_cast<int>(i - b * n_out) + half;     // position in the untrimmed signal
        const int q = n / hop;
        float acc = 0.f, env = 0.f;
        int f_lo = (n - L + hop) / hop;                            // first frame covering n (ceil((n-L+1)/hop))
        if (n - L + 1 <= 0) f_lo = 0;
        const int f_hi = q < F - 1 ? q : F - 1;
        for (int f = f_lo; f <= f_hi; ++f) {
            const int r = n - f * hop;
            acc += frames[(b * F + f) * L + r];
            if (windowed) env = static_cast<float>(static_cast<double>(env) + win_sq[r]);
        }
        if (env > 1.17549435e-38f) acc /= env;
        out[i] = acc * scale;
    }
}

// 4 consecutive samples per thread (hop % 4 == 0, L % 4 == 0): the same frames cover all four, so each covering frame
// contributes one float4 load and the envelope terms are read once per frame; same summation order as the scalar kernel
__global__ void istft_overlap_add4_kernel(const float* __restrict__ frames, const double* __restrict__ win_sq,
                                          float* __restrict__ out, int F, int L, int hop, int n_out, long long total4) {
    const int half = L / 2;
    const bool windowed = win_sq != nullptr;
    const float scale = windowed ? static_cast<float>(L) / static_cast<float>(hop) : 1.f;
    const int per_row = n_out >> 2;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = i / per_row;
        const int n = (static_cast<int>(i - b * per_row) << 2) + half;      // first of 4 positions, untrimmed signal
        const int q = n / hop;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, env[4] = {0.f, 0.f, 0.f, 0.f};
        int f_lo = (n - L + hop) / hop;
        if (n - L + 1 <= 0) f_lo = 0;
        const int f_hi = q < F - 1 ? q : F - 1;
        // n .. n+3 lie in the same hop interval and the same set of frames (n % 4 == 0, hop % 4 == 0, L % 4 == 0)
        for (int f = f_lo; f <= f_hi; ++f) {
            const int r = n - f * hop;
            const float4 v = *reinterpret_cast<const float4*>(frames + (b * F + f) * L + r);
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
            if (windowed) {
#pragma unroll
                for (int j = 0; j < 4; ++j) env[j] = static_cast<float>(static_cast<double>(env[j]) + win_sq[r + j]);
            }
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = (env[j] > 1.17549435e-38f ? acc[j] / env[j] : acc[j]) * scale;
        *reinterpret_cast<float4*>(out + b * n_out + (n - half)) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

int istft_overlap_add(const float* frames, const double* win_sq, float* out, int batch, int F, int L, int hop,
                      cudaStream_t stream) {
    WGB_REQUIRE(frames && out && batch > 0 && F > 1 && L > 0 && hop > 0, "bad arguments");
    const int n_out = hop * (F - 1);
    if (hop % 4 == 0 && L % 8 == 0 && (reinterpret_cast<uintptr_t>(frames) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        const long long total4 = static_cast<long long>(batch) * (n_out >> 2);
        istft_overlap_add4_kernel<<<grid_for(total4, 256), 256, 0, stream>>>(frames, win_sq, out, F, L, hop, n_out, total4);
        WGB_LAUNCH_CHECK();
        return WGB_OK;
    }
    const long long total = static_cast<long long>(batch) * n_out;
    istft_overlap_add_kernel<<<grid_for(total, 256), 256, 0, stream>>>(frames, win_sq, out, F, L, hop, n_out, total);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace wgb
