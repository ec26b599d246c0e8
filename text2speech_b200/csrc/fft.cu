// Butterfly STFT kernels: TacotronSTFT.mel_spectrogram (utils/layers.py:63-79) and Denoiser.forward
// (waveglow/denoiser.py:35-40 = STFT.transform + spectral subtraction + STFT.inverse, utils/stft.py:71-130) for the stock
// bases.  The reference's forward_basis is window * [cos; -sin](2 pi k n / L) and its inverse_basis the pseudo-inverse of that
// (stft.py:46-60), i.e. a real DFT / inverse real DFT: 10 N log2 N = 0.05 MFLOP per frame as butterflies against 2.1 MFLOP
// as a dense-basis contraction (x3 for fp32-grade accuracy on the bf16 tensor pipe).  These kernels are the HBM-side answer
// to the tensor-core STFT kernels of stft_tc2.cu: the signal is read from DRAM once, nothing but the result is written.
//
// One warp = one frame at a time: 1024 windowed samples -> 512-point complex FFT in registers + two small shared-memory
// exchange buffers (fft_core.cuh) -> real-FFT split by warp shuffles.  Warps are persistent and independent (no block-level
// barrier after the tables are staged); the NEXT frame's samples are loaded into the registers that are dead between stage 1
// and stage 3 of the current one, so no warp ever waits for global memory.
//   mel      warps walk the flat frame axis of the batch a stride of all resident warps apart (an SM always works on
//            consecutive, overlapping frames); |X| -> shared memory -> mel filterbank as 8-bin pieces, every filter summed by
//            one lane in a fixed order -> log(clamp) -> 4-byte stores into out [B, n_mel, F] that merge in L2 with the
//            neighbouring warps' frames
//   denoise  spectral subtraction on the lane's own bins -> inverse split -> the same butterfly code again (a 2-trip loop:
//            one copy in the instruction cache) -> windowed frame; a warp walks a run of consecutive frames and keeps the
//            overlap-add in REGISTERS: the lane's samples 2 (lane + 32 j) + {0,1} of frame r and the samples of frame r + 1
//            that overlap them differ by j -> j - 4 in the same lane (hop = L / 4)
// L = 1024 only (the config's filter length); other shapes, and hand-edited bases, stay on the dense-basis kernels.
#include "common.cuh"
#include "fft_core.cuh"

namespace wgb {

namespace fftk {

using namespace fft;

constexpr int kL = 1024;
constexpr int kHalf = kL / 2;
// mel: one CTA per SM of 12 warps x 168 registers (default; the next frame's samples are prefetched into registers) or of
// 16 warps x 128 registers with a few spills (wgb_set_tuning "fft_mel_warps" = 16, A/B)
constexpr int kDnWarps = 12;              // denoise: the same (the register file is split over 4 schedulers: 3 warps each)
constexpr int kDnThreads = kDnWarps * 32;
constexpr int kMaxMel = 128;
constexpr int kMaxSlots = 16;             // filterbank pieces (of 8 bins) per lane
constexpr int kMagPad = 528;              // |X| buffer: bins 0..512, then zeros that the last pieces may read

__device__ __forceinline__ int reflect_index(int p, int n) {          // np.pad(mode='reflect'), |overhang| < n
    p = p < 0 ? -p : p;
    return p >= n ? 2 * (n - 1) - p : p;
}

// ---- one frame's samples as they sit in memory (reflect-indexed at the ends): lane gets x[2 (lane + 32 j)], x[.. + 1].
// Issued a whole frame ahead of their use (the registers of v are dead between stage 1 and stage 3), so the global-memory
// latency hides behind the previous frame's butterflies.
__device__ __forceinline__ void issue_frame_loads(const float* __restrict__ yb, int n, int start, int lane, float2* nx) {
    const bool interior = start >= 0 && start + kL <= n;
    const bool vec = interior && ((reinterpret_cast<uintptr_t>(yb + start) & 7) == 0);      // warp-uniform
    if (vec) {
#pragma unroll
        for (int j = 0; j < 16; ++j) nx[j] = __ldg(reinterpret_cast<const float2*>(yb + start + 2 * (lane + 32 * j)));
    } else if (interior) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int p = start + 2 * (lane + 32 * j);
            nx[j] = make_float2(__ldg(yb + p), __ldg(yb + p + 1));
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int p = start + 2 * (lane + 32 * j);
            nx[j] = make_float2(__ldg(yb + reflect_index(p, n)), __ldg(yb + reflect_index(p + 1, n)));
        }
    }
}

// samples x window -> v; with CHECK returns true if a sample is outside [-1, 1] or NaN.  Every sample sits in L / hop
// frames: the first and last frame (and every frame when the hop is wider than the central part) check everything, the
// others the 256 samples around their centre.
template <bool CHECK>
__device__ __forceinline__ bool window_frame(const float2* nx, const float2* __restrict__ win2, int lane, bool check_all, cf* v) {
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 x = nx[j];
        if (CHECK) {
            if (check_all || (j >= 6 && j < 10)) bad = bad || !(x.x >= -1.f && x.x <= 1.f) || !(x.y >= -1.f && x.y <= 1.f);
        }
        const float2 w = win2[lane + 32 * j];
        v[j] = cf{x.x * w.x, x.y * w.y};
    }
    return bad;
}

// two barriers per transform; see fft_core.cuh for why none is needed between back-to-back transforms
__device__ __forceinline__ void fft512(int lane, const LaneTw& tw, cf* v, cf* buf_a, cf* buf_b) {
    stage1(lane, tw, v, buf_a);
    __syncwarp();
    stage2(lane, tw, buf_a, buf_b);
    __syncwarp();
    stage3(lane, v, buf_b);
}

__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// partner value for bin k = lane + 32 j: Z[512 - k] from lane (32 - lane) & 31 (register 15 - j there), or this lane's own
// register (16 - j) & 15 when lane = 0
__device__ __forceinline__ cf partner_of(const cf* v, int j, int lane) {
    const cf snd = v[15 - j];
    cf p;
    p.x = __shfl_sync(0xffffffffu, snd.x, (32 - lane) & 31);
    p.y = __shfl_sync(0xffffffffu, snd.y, (32 - lane) & 31);
    if (lane == 0) p = v[(16 - j) & 15];
    return p;
}

// Z (v) -> X in place for the bins k = lane + 32 j with j < J_USED (compile time; the others keep Z: the mel filterbank
// at 22.05 kHz / 8 kHz reads bins 0..371 only); returns X[512] (meaningful in lane 0)
template <int J_USED = 16>
__device__ __forceinline__ float rfft_split(int lane, const LaneTw& tw, cf* v) {
    const float nyq = v[0].x - v[0].y;
    cf x[J_USED];
#pragma unroll
    for (int j = 0; j < J_USED; ++j) x[j] = rfft_bin(v[j], partner_of(v, j, lane), cmul(tw.post, w32(j)));
#pragma unroll
    for (int j = 0; j < J_USED; ++j) v[j] = x[j];
    return nyq;
}

// X (v, nyq) -> conj(Z) in place, ready for the forward FFT (fft_core.cuh: irfft_bin)
__device__ __forceinline__ void irfft_split(int lane, const LaneTw& tw, cf* v, float nyq) {
    cf z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        cf p = partner_of(v, j, lane);
        if (j == 0 && lane == 0) p = cf{nyq, 0.f};
        z[j] = cconj(irfft_bin(v[j], p, cmul(tw.post, w32(j))));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = z[j];
}

// ------------------------------------------------------------------------------------------------ mel
struct MelParams {
    const float* y;            // [B, n]
    const float* window;       // [1024] (the window zero-padded to the filter length; ones for window=None)
    const int* mel_slots;      // [slots_per_lane][32] packed {first bin / 4 | (filter to emit + 1) << 8 | starts a filter << 16}: the
                               // filterbank as 8-bin pieces (4-bin aligned, zero-padded weights); lane l sums the pieces
                               // [q][l], q = 0 .., in order: ALL pieces of the filters it owns, so no two lanes ever add
                               // into the same filter
    int slots_per_lane;
    const float4* mel_w;       // [slots_per_lane][2][32]: the 8 weights of piece [q][l] as two float4, lanes innermost
                               // (conflict-free shared-memory reads)
    float* out;                // [B, n_mel, frames]
    int batch, n, frames, hop, n_mel;
    float clip;
    int* range_flag;           // optional: set to 1 when a sample is outside [-1, 1] or NaN (layers.py:72-73)
};

template <bool CHECK, int kMelWarps, int J_USED>
__global__ void __launch_bounds__(kMelWarps * 32, 1) fft_mel_kernel(const MelParams p) {
    constexpr int kMelThreads = kMelWarps * 32;
    extern __shared__ __align__(16) unsigned char smem[];
    cf* bufs = reinterpret_cast<cf*>(smem);                                         // [kMelWarps][kBufElems]
    float2* win2 = reinterpret_cast<float2*>(bufs + kMelWarps * kBufElems);         // [512]
    float4* mw = reinterpret_cast<float4*>(win2 + kHalf);                           // [slots_per_lane][2][32]
    int* slots = reinterpret_cast<int*>(mw + kMaxSlots * 64);                       // [slots_per_lane][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kHalf; i += kMelThreads) win2[i] = make_float2(p.window[2 * i], p.window[2 * i + 1]);
    for (int i = threadIdx.x; i < 32 * p.slots_per_lane; i += kMelThreads) slots[i] = p.mel_slots[i];
    for (int i = threadIdx.x; i < 64 * p.slots_per_lane; i += kMelThreads) mw[i] = p.mel_w[i];
    __syncthreads();

    LaneTw tw;
    lane_twiddles(lane, tw);
    cf* buf_a = bufs + warp * kBufElems;
    cf* buf_b = buf_a + kBufA;
    float* mag = reinterpret_cast<float*>(buf_a);              // [kMagPad] |X| (zeros past bin 512), then [kMaxMel] filter sums:
    float* acc = mag + kMagPad;                                // buffer A is free once every lane is past stage 2
    bool bad = false;

    // warps walk the frames of the whole batch in flat order, a stride of all resident warps apart: at any moment an SM
    // works on consecutive frames, whose samples overlap (L1 / L2 hits), and no warp ever waits for another
    const int stride = static_cast<int>(gridDim.x) * kMelWarps;
    int b = 0, r = static_cast<int>(blockIdx.x) * kMelWarps + warp;           // (utterance, frame) of this warp's current frame
    while (r >= p.frames) { r -= p.frames; ++b; }
    float2 nx[16];
    if (b < p.batch) issue_frame_loads(p.y + static_cast<size_t>(b) * p.n, p.n, r * p.hop - kHalf, lane, nx);
    while (b < p.batch) {
        cf v[16];
        bad = window_frame<CHECK>(nx, win2, lane, r == 0 || r == p.frames - 1 || p.hop > 256, v) || bad;
        stage1(lane, tw, v, buf_a);
        int bn = b, rn = r + stride;                                                // the frame after this one
        while (rn >= p.frames) { rn -= p.frames; ++bn; }
        if (bn < p.batch)                                                           // v is dead until stage 3: fetch ahead
            issue_frame_loads(p.y + static_cast<size_t>(bn) * p.n, p.n, rn * p.hop - kHalf, lane, nx);
        __syncwarp();
        stage2(lane, tw, buf_a, buf_b);
        __syncwarp();
        stage3(lane, v, buf_b);
        const float nyq = rfft_split<J_USED>(lane, tw, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)                                                // stft.py:94; bins nobody reads stay zero
            mag[lane + 32 * j] = j < J_USED ? sqrt_approx(v[j].x * v[j].x + v[j].y * v[j].y) : 0.f;
        if (lane < kMagPad - kHalf) mag[kHalf + lane] = lane == 0 ? fabsf(nyq) : 0.f;
        __syncwarp();
        {                                                                           // layers.py:77
            const float4* mag4 = reinterpret_cast<const float4*>(mag);
            float sum = 0.f;
            for (int q = 0; q < p.slots_per_lane; ++q) {
                const int slot = slots[q * 32 + lane];
                const float4* a = mag4 + (slot & 255);
                const float4 a0 = a[0], a1 = a[1];
                const float4 w0 = mw[q * 64 + lane], w1 = mw[q * 64 + 32 + lane];
                const float s0 = fmaf(a0.w, w0.w, fmaf(a0.z, w0.z, fmaf(a0.y, w0.y, a0.x * w0.x)));
                const float s1 = fmaf(a1.w, w1.w, fmaf(a1.z, w1.z, fmaf(a1.y, w1.y, a1.x * w1.x)));
                sum = ((slot >> 16) ? 0.f : sum) + (s0 + s1);
                const int emit = ((slot >> 8) & 255) - 1;
                if (emit >= 0) acc[emit] = sum;
            }
        }
        __syncwarp();
        // out [B, n_mel, frames]: a 4-byte store per filter; the neighbouring frames of the same 32-byte sectors come from
        // the neighbouring warps within microseconds and merge in L2 before anything reaches DRAM
        float* ob = p.out + static_cast<size_t>(b) * p.n_mel * p.frames + r;
        for (int m = lane; m < p.n_mel; m += 32) ob[static_cast<size_t>(m) * p.frames] = logf(fmaxf(acc[m], p.clip));   // :78
        __syncwarp();
        b = bn;
        r = rn;
    }
    if (CHECK && bad) atomicOr(p.range_flag, 1);
}

// ------------------------------------------------------------------------------------------------ denoise
struct DenoiseParams {
    const float* y;            // [B, n]
    const float* window;       // [1024]
    const float* bias_spec;    // [513]
    float strength;
    const float* env_tab;      // [16][hop]: window-sum envelope per set of covering frames (stft_tc2.cu), NULL for window=None
    float scale;               // L / hop (stft.py:54, :125)
    float* out;                // [B, 1, hop (frames - 1)]
    int batch, n, frames, hop;
    int run, runs_per_b;       // output blocks per warp, warps per utterance
};

__device__ __forceinline__ void denoise_bin(float& re, float& im, float bias_s) {     // as stft_tc2.cu: denoiser.py:36-38
    const float m2 = fmaf(re, re, im * im);
    float rs;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(m2));
    const float g = fmaxf(fmaf(-bias_s, rs, 1.f), 0.f);
    const bool zero = !(m2 >= 1.17549435e-38f);
    re = zero ? fmaxf(-bias_s, 0.f) : re * g;
    im = zero ? 0.f : im * g;
}

__global__ void __launch_bounds__(kDnThreads, 1) fft_denoise_kernel(const DenoiseParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    cf* bufs = reinterpret_cast<cf*>(smem);
    float2* win2 = reinterpret_cast<float2*>(bufs + kDnWarps * kBufElems);
    float* sbias = reinterpret_cast<float*>(win2 + kHalf);                           // [513] bias_spec * strength

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kHalf; i += kDnThreads) win2[i] = make_float2(p.window[2 * i], p.window[2 * i + 1]);
    for (int i = threadIdx.x; i <= kHalf; i += kDnThreads) sbias[i] = p.bias_spec[i] * p.strength;
    __syncthreads();

    const long long item = static_cast<long long>(blockIdx.x) * kDnWarps + warp;
    if (item >= static_cast<long long>(p.batch) * p.runs_per_b) return;               // whole warp
    const int b = static_cast<int>(item / p.runs_per_b);
    // output block q = samples [q hop, (q+1) hop) of the padded time line = frames q-3 .. q; blocks 2 .. frames survive the
    // L/2 trim at both ends (stft.py:127-128)
    const int q0 = 2 + static_cast<int>(item % p.runs_per_b) * p.run;
    const int q1 = min(q0 + p.run, p.frames + 1);

    LaneTw tw;
    lane_twiddles(lane, tw);
    cf* buf_a = bufs + warp * kBufElems;
    cf* buf_b = buf_a + kBufA;
    const float* yb = p.y + static_cast<size_t>(b) * p.n;
    float* ob = p.out + static_cast<size_t>(b) * p.hop * (p.frames - 1);
    const float frame_scale = 1.f / (512.f * p.scale);          // inverse FFT's 1/512 and the pseudo-inverse's hop / L

    cf pend[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) pend[j] = cf{0.f, 0.f};

    const int r_begin = max(q0 - 3, 0);
    const int r_real_end = min(q1, p.frames);                    // real frames of this run: [r_begin, r_real_end)
    float2 nx[16];
    if (r_begin < r_real_end) issue_frame_loads(yb, p.n, r_begin * p.hop - kHalf, lane, nx);
    for (int r = r_begin; r < q1; ++r) {
        cf v[16];
        if (r < p.frames) {
            window_frame<false>(nx, win2, lane, false, v);
            float nyq = 0.f;
            // the two transforms of a frame share ONE copy of the butterfly code (the instruction cache is the scarce
            // resource of this kernel): pass 0 = forward, pass 1 = inverse by conjugation
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                if (pass == 1) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) denoise_bin(v[j].x, v[j].y, sbias[lane + 32 * j]);
                    float nyq_im = 0.f;
                    denoise_bin(nyq, nyq_im, sbias[kHalf]);
                    irfft_split(lane, tw, v, nyq);
                }
                stage1(lane, tw, v, buf_a);
                if (pass == 0 && r + 1 < r_real_end) issue_frame_loads(yb, p.n, (r + 1) * p.hop - kHalf, lane, nx);
                __syncwarp();
                stage2(lane, tw, buf_a, buf_b);
                __syncwarp();
                stage3(lane, v, buf_b);
                if (pass == 0) nyq = rfft_split(lane, tw, v);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {                       // conj(512 z): x[2m] = Re, x[2m+1] = -Im; times the window
                const float2 w = win2[lane + 32 * j];
                v[j] = cf{v[j].x * frame_scale * w.x, -v[j].y * frame_scale * w.y};
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = cf{0.f, 0.f};
        }
        if (r >= q0) {                                           // block r is complete: frames r-3 .. r, oldest first
            int mask = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) mask |= (r - t >= 0 && r - t < p.frames) ? (1 << t) : 0;
            const float* env = p.env_tab ? p.env_tab + mask * p.hop : nullptr;
            float* dst = ob + static_cast<size_t>(r - 2) * p.hop;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = 2 * (lane + 32 * j);
                float a0 = pend[j].x + v[j].x, a1 = pend[j].y + v[j].y;
                if (env) {                                       // stft.py:113-125
                    const float2 e = __ldg(reinterpret_cast<const float2*>(env + i));
                    if (e.x > 1.17549435e-38f) a0 /= e.x;
                    if (e.y > 1.17549435e-38f) a1 /= e.y;
                    a0 *= p.scale;
                    a1 *= p.scale;
                }
                *reinterpret_cast<float2*>(dst + i) = make_float2(a0, a1);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) pend[j] = cadd(pend[j + 4], v[j + 4]);
#pragma unroll
        for (int j = 8; j < 12; ++j) pend[j] = v[j + 4];
    }
}

}  // namespace fftk

static int check_common(const float* y, const float* window, int batch, int n, int hop) {
    WGB_REQUIRE(y && window, "null pointer");
    WGB_REQUIRE(batch > 0 && hop > 0, "batch (%d) and hop (%d) must be positive", batch, hop);
    WGB_REQUIRE(n > fftk::kHalf, "signal length %d must exceed filter_length / 2 = %d (reflect padding)", n, fftk::kHalf);
    return WGB_OK;
}

int fft_stft_mel(const float* y, const float* window, const void* mel_slots, int slots_per_lane, const float* mel_w,
                 int bins_used, float* out, int batch, int n, int hop, int n_mel, float clip, int* range_flag,
                 cudaStream_t stream) {
    using namespace fftk;
    if (int e = check_common(y, window, batch, n, hop)) return e;
    WGB_REQUIRE(mel_slots && mel_w && out, "null pointer");
    WGB_REQUIRE(n_mel >= 1 && n_mel <= kMaxMel, "n_mel (%d) must be in 1..%d", n_mel, kMaxMel);
    WGB_REQUIRE(slots_per_lane >= 1 && slots_per_lane <= kMaxSlots, "slots_per_lane (%d) must be in 1..%d", slots_per_lane,
                kMaxSlots);
    WGB_REQUIRE(bins_used >= 1 && bins_used <= kHalf + 1, "bins_used (%d) must be in 1..%d", bins_used, kHalf + 1);
    MelParams p{};
    p.y = y; p.window = window; p.mel_slots = static_cast<const int*>(mel_slots); p.slots_per_lane = slots_per_lane;
    p.mel_w = reinterpret_cast<const float4*>(mel_w);
    p.out = out; p.batch = batch; p.n = n; p.frames = n / hop + 1; p.hop = hop; p.n_mel = n_mel; p.clip = clip;
    p.range_flag = range_flag;
    WGB_REQUIRE(static_cast<long long>(batch) * p.frames < 0x40000000LL, "too many frames");
    const int warps = tuning_get("fft_mel_warps") == 16 ? 16 : 12;
    const int smem = warps * kBufElems * 8 + kHalf * 8 + kMaxSlots * (64 * 16 + 32 * 4);
    // bins 0 .. 383 suffice for the usual filterbanks (fmax 8 kHz at 22.05 kHz ends at bin 371): the real-FFT split and the
    // magnitudes of the upper quarter are then compiled out (bin 512, the Nyquist bin, is always computed)
    const bool quarter = bins_used <= 384 && warps == 12;
    void (*kern)(MelParams) =
        warps == 16 ? (range_flag ? fft_mel_kernel<true, 16, 16> : fft_mel_kernel<false, 16, 16>)
        : quarter   ? (range_flag ? fft_mel_kernel<true, 12, 12> : fft_mel_kernel<false, 12, 12>)
                    : (range_flag ? fft_mel_kernel<true, 12, 16> : fft_mel_kernel<false, 12, 16>);
    WGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int per_sm = 1;
    WGB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    const long long resident = static_cast<long long>(sm_count()) * (per_sm > 0 ? per_sm : 1);
    const long long want = (static_cast<long long>(batch) * p.frames + warps - 1) / warps;
    kern<<<static_cast<unsigned>(want < resident ? want : resident), warps * 32, smem, stream>>>(p);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

int fft_denoise(const float* y, const float* window, const float* bias_spec, float strength, const float* env_tab, float* out,
                int batch, int n, int hop, cudaStream_t stream) {
    using namespace fftk;
    if (int e = check_common(y, window, batch, n, hop)) return e;
    WGB_REQUIRE(bias_spec && out, "null pointer");
    WGB_REQUIRE(hop * 4 == kL, "the register overlap-add needs hop = filter_length / 4 (got hop %d)", hop);
    DenoiseParams p{};
    p.y = y; p.window = window; p.bias_spec = bias_spec; p.strength = strength; p.env_tab = env_tab;
    p.scale = static_cast<float>(kL) / static_cast<float>(hop);
    p.out = out; p.batch = batch; p.n = n; p.frames = n / hop + 1; p.hop = hop;
    const int blocks_out = p.frames - 1;                       // hop-sized output blocks per utterance
    if (blocks_out <= 0) return WGB_OK;
    const int smem = kDnWarps * kBufElems * 8 + kHalf * 8 + (kHalf + 4) * 4;
    WGB_CUDA_TRY(cudaFuncSetAttribute(fft_denoise_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // run length: every warp recomputes 3 frames of history, so long runs are cheap per block but leave SMs idle in the last
    // wave; pick the length with the smallest (waves x frames per warp)
    int per_sm = 1;
    WGB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fft_denoise_kernel, kDnThreads, smem));
    const long long resident = static_cast<long long>(sm_count()) * (per_sm > 0 ? per_sm : 1) * kDnWarps;
    int best_run = blocks_out;
    long long best_cost = -1;
    for (int run = 8; run <= 256; ++run) {
        const long long warps = static_cast<long long>(batch) * ceil_div(blocks_out, run);
        const long long waves = (warps + resident - 1) / resident;
        const long long cost = waves * (run + 3);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_run = run; }
    }
    if (best_run > blocks_out) best_run = blocks_out;
    p.run = best_run;
    p.runs_per_b = ceil_div(blocks_out, p.run);
    const long long warps = static_cast<long long>(batch) * p.runs_per_b;
    fft_denoise_kernel<<<static_cast<unsigned>((warps + kDnWarps - 1) / kDnWarps), kDnThreads, smem, stream>>>(p);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace wgb
