// STFT-family GEMMs on CTA pairs (tcgen05.mma.cta_group::2, UMMA M = 256 over two SMs), split-bf16 operands.
//
// The dense-basis contractions of the conv STFT (reference utils/stft.py:85-89 forward, :105-109 inverse) keep fp32-grade
// accuracy on the bf16 tensor pipe by running   C = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T   as ONE K = 3K GEMM:
// K segments [A_hi | A_lo | A_hi] against the packed weight [W_hi | W_hi | W_lo].  What this file adds over the
// one-CTA kernels of wn_tc.cu (kept as the A/B reference):
//
//   * CTA pairs: each CTA loads its own 128 rows and HALF of every 256-row basis tile (16 + 16 KB per K chunk instead
//     of 16 + 32), so the ring is 5-6 stages deep instead of 3-4 and the tensor pipe is fed from half the shared-memory
//     traffic per FLOP;
//   * no padded bins: Im of bin 0 and of bin L/2 are exactly zero in the reference's basis (stft.py:46-51), so the
//     L/2 + 1 = 513 bins are exactly 2 * 512 = 1024 real numbers.  The Re row of the Nyquist bin sits in the Im slot of
//     bin 0: four 256-column passes (Re 128 bins | Im 128 bins) cover the whole spectrum (the one-CTA layout pads 513
//     bins to 640 = five passes), and the inverse GEMM's K is 1024 instead of 1280;
//   * the mel variant runs only the passes that hold a bin with non-zero mel weight (3 of 4 at 22.05 kHz / fmax 8 kHz:
//     bins above 371 carry no weight) -- exact, the skipped magnitudes are multiplied by zero in the reference too;
//   * rows are tiled over ONE flat frame axis for the whole batch: the padded signal of every utterance has a pitch of
//     R hops, so frame r of utterance b is flat row b R + r of a 2-D tensor map whose row stride is the hop (rows
//     overlap in memory exactly like the reference's strided conv reads them); the L/hop - 1 rows per utterance that
//     straddle two utterances are computed and dropped (0.35 % at 10 s) instead of padding every utterance to a whole
//     number of 128-row tiles (4 %).
//
//   EPI_F32      C fp32 [rows, N]                                  (inverse-basis GEMM; generic split-bf16 GEMM)
//   EPI_MEL      |X| -> sparse mel filterbank -> log(clamp)        (all of TacotronSTFT.mel_spectrogram, layers.py:63-79);
//                a bin feeds two adjacent triangular filters and the filter index never decreases, so each row keeps
//                two running sums in registers and emits a filter the moment the table moves past it
//   EPI_DENOISE  max(|X| - bias*strength, 0) e^{j arg X} as the bf16 hi/lo operands of the inverse GEMM
//                                                                   (denoiser.py:36-38 + stft.py:102-103)
//   EPI_OLA      the inverse-basis conv_transpose1d INCLUDING its overlap-add (stft.py:105-128): a tile row is one
//                hop-sized block of the output signal, block q = sum_j frame[q - j] W_j with W_j = the j-th hop-wide
//                slice of the inverse basis -- a tap conv over the frame axis (K = taps x 3L, N = hop), so the sums
//                are complete in TMEM and the epilogue applies the window-sum envelope, the L/hop scale and the L/2
//                trim: the [B, frames, L] intermediate (0.9 GB at 256 x 10 s) and its overlap-add pass do not exist
#include "common.cuh"
#include "ptx.cuh"

namespace wgb {

namespace stft2 {

constexpr int kBlockM = 128;          // rows per CTA (UMMA M = 256 per pair)
constexpr int kBlockN = 256;
constexpr int kHalfN = 128;           // basis rows each CTA loads
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytes = kHalfN * kBlockK * 2;       // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;
constexpr int kThreads = 192;
constexpr int kMaxBins = 512;         // EPI_MEL / EPI_DENOISE: L/2 <= 512 (per-bin table of L/2 + 1 entries in shared memory)

enum Epi { EPI_F32 = 0, EPI_MEL = 1, EPI_DENOISE = 2, EPI_OLA = 3 };

template <int EPI>
struct Smem {
    static constexpr int kStages = 6;
    static constexpr int kExtraOff = kStages * kStageBytes;
    // MEL: the sparse filterbank table, one float4 per bin; DENOISE: bias * strength per bin
    static constexpr int kExtraBytes = EPI == EPI_MEL ? (kMaxBins + 1) * 16 : (EPI == EPI_DENOISE ? (kMaxBins + 1) * 4 : 0);
    static constexpr int kBarOff = kExtraOff + ((kExtraBytes + 15) & ~15);
    static constexpr int kTotal = 1024 + kBarOff + 256;
};

struct Params {
    int n_tiles;                  // 128-row tiles over the flat row axis
    int l2_hint;                  // wgb_set_tuning "stft_l2_hint": bit 0 = basis tiles evict_last
    int rows_total;               // flat rows
    int R, frames;                // flat row = b R + r; it is a real row iff r < frames (R == frames: compact rows);
                                  // outputs are indexed by the compact row b frames + r
    int n_pass, ppi, n_chunks;    // 256-column passes, passes per work item, K chunks (of the 3K split operand) per pass
    int seg_chunks;               // K / 64: chunk kc reads A segment kc / seg_chunks (hi, lo, hi), columns (kc % seg_chunks) * 64
    int taps;                     // EPI_OLA: L / hop; chunk kc belongs to tap kc / (3 seg_chunks) and reads A rows t - tap
                                  // (1 elsewhere: n_chunks = 3 seg_chunks, no shift)
    int out_R;                    // EPI_DENOISE: row pitch per utterance of hi_out / lo_out (>= frames; EPI_OLA's guard rows)
    const float* env_tab;         // EPI_OLA [2^taps][hop]: window-sum envelope per set of covering frames (NULL: window=None)
    float ola_scale;              // EPI_OLA: L / hop (stft.py:125)
    int n_total;                  // EPI_F32: N (row pitch of c_out); EPI_DENOISE: L (row pitch of the hi / lo outputs)
    int cp;                       // EPI_MEL / EPI_DENOISE: L / 2 (bins 0 .. cp - 1 are (Re, Im) pairs, bin cp rides in Im slot 0)
    float* c_out;                 // EPI_F32 [rows, N]; EPI_MEL [B, n_mel, frames]
    __nv_bfloat16* hi_out;        // EPI_DENOISE [B frames, L]: cols 0..cp-1 Re, col cp Re of bin cp, cols cp+1.. Im of bins 1..cp-1
    __nv_bfloat16* lo_out;
    const float* spec_bias;       // EPI_DENOISE [cp + 1]
    float strength;
    const float4* mel_table;      // EPI_MEL [cp + 1]: {first filter index (as float), weight in it, weight in the next, 0};
                                  // over the bins with weight the index never decreases
    int n_mel;
    float mel_clip;
};

// Denoiser.forward on one bin (denoiser.py:36-38, stft.py:102-103): (re, im) -> max(|X| - bias*strength, 0) e^{j arg X}
// without atan2 / cos / sin: both parts are scaled by g = max(|X| - bs, 0) / |X| = max(1 - bs / |X|, 0), with 1 / |X| from
// one rsqrt.approx (relative error 2^-22: the error of g is 2^-22 absolute, i.e. 2^-22 |X| on the output).  |X| = 0
// (or a denormal |X|^2, flushed) keeps the reference's atan2(0, 0) = 0: Re = max(-bs, 0), Im = 0.
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void denoise_bin(float& re, float& im, float bias_s) {
    const float m2 = fmaf(re, re, im * im);
    const float g = fmaxf(fmaf(-bias_s, rsqrt_approx(m2), 1.f), 0.f);
    const bool zero = !(m2 >= 1.17549435e-38f);
    re = zero ? fmaxf(-bias_s, 0.f) : re * g;
    im = zero ? 0.f : im * g;
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
stft_pair_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                 const __grid_constant__ CUtensorMap map_w, const Params p) {
    using SL = Smem<EPI>;
    constexpr int kStages = SL::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::kBarOff);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    uint8_t* s_extra = smem + SL::kExtraOff;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_hi);
        tma_prefetch_desc(&map_lo);
        tma_prefetch_desc(&map_w);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);      // 4 epilogue warps x 2 CTAs (only the leader's copy is used)
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, kTmemCols);
    if (EPI == EPI_MEL && warp >= 2) {          // table in shared memory; first filter index as int bits, -1 = no weight
        float4* tab = reinterpret_cast<float4*>(s_extra);
        for (int i = threadIdx.x - 64; i <= p.cp; i += 128) {
            float4 e = p.mel_table[i];
            e.x = __int_as_float((e.y != 0.f || e.z != 0.f) ? static_cast<int>(e.x) : -1);
            tab[i] = e;
        }
    }
    if (EPI == EPI_DENOISE && warp >= 2) {
        float* bs = reinterpret_cast<float*>(s_extra);
        for (int i = threadIdx.x - 64; i <= p.cp; i += 128) bs[i] = p.spec_bias[i] * p.strength;
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int n_pairs = gridDim.x >> 1;
    const int pair_id = blockIdx.x >> 1;
    const int n_pair_tiles = (p.n_tiles + 1) >> 1;
    const int groups = p.n_pass / p.ppi;
    const int n_items = n_pair_tiles * groups;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int item = pair_id; item < n_items; item += n_pairs) {
                const int tile = 2 * (item / groups) + static_cast<int>(rank);
                // an absent second tile reads rows past the end of the map = zeros
                const int t0 = tile < p.n_tiles ? tile * kBlockM : p.rows_total + 4 * kBlockM;
                for (int pp = 0; pp < p.ppi; ++pp) {
                    const int pass = (item % groups) * p.ppi + pp;
                    for (int kc = 0; kc < p.n_chunks; ++kc) {
                        mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
                        uint8_t* sa = smem + s * kStageBytes;
                        uint8_t* sb = sa + kABytes;
                        if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * kStageBytes);
                        const uint32_t bar = mapa_u32(&full_bar[s], 0);
                        const int tap = kc / (3 * p.seg_chunks);               // EPI_OLA: frame q - tap feeds block q
                        const int rem = kc - tap * 3 * p.seg_chunks;
                        const int seg = rem / p.seg_chunks;                    // 0: A_hi, 1: A_lo, 2: A_hi
                        tma_load_2d_2sm(sa, seg == 1 ? &map_lo : &map_hi, bar, (rem - seg * p.seg_chunks) * kBlockK, t0 - tap);
                        if (p.l2_hint & 1)          // basis tiles (6 MB, re-read by every pair) with evict_last priority
                            tma_load_2d_2sm_hint(sb, &map_w, bar, kc * kBlockK, pass * kBlockN + static_cast<int>(rank) * kHalfN,
                                                 l2_policy_evict_last());
                        else
                            tma_load_2d_2sm(sb, &map_w, bar, kc * kBlockK, pass * kBlockN + static_cast<int>(rank) * kHalfN);
                        if (++s == kStages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(2 * kBlockM, kBlockN);
            int s = 0;
            uint32_t ph = 0, acc_it = 0;
            for (int item = pair_id; item < n_items; item += n_pairs) {
                for (int pp = 0; pp < p.ppi; ++pp, ++acc_it) {
                    const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                    mbar_wait(&tempty_bar[as], aph ^ 1, 200 + as);
                    tc_fence_after_sync();
                    const uint32_t d_tmem = tmem_base + as * kBlockN;
                    for (int kc = 0; kc < p.n_chunks; ++kc) {
                        mbar_wait(&full_bar[s], ph, 300 + s);
                        tc_fence_after_sync();
                        const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
                        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            umma_bf16_ss_2sm(d_tmem, umma_desc_sw128(a_addr + k * kUmmaK * 2),
                                             umma_desc_sw128(b_addr + k * kUmmaK * 2), idesc, (kc | k) != 0);
                        }
                        umma_commit_2sm(&empty_bar[s]);
                        if (++s == kStages) { s = 0; ph ^= 1; }
                    }
                    umma_commit_2sm(&tfull_bar[as]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        for (int item = pair_id; item < n_items; item += n_pairs) {
            const int tile = 2 * (item / groups) + static_cast<int>(rank);
            const int flat = tile * kBlockM + row;
            const int b = flat / p.R;
            const int r = flat - b * p.R;
            // EPI_OLA rows are output blocks: all R = frames + taps - 1 of them exist (the trim is applied below)
            const bool live = tile < p.n_tiles && flat < p.rows_total && (EPI == EPI_OLA || r < p.frames);
            // compact output row (EPI_DENOISE: with the pitch the overlap-add GEMM wants, guard rows between utterances)
            const size_t grow = static_cast<size_t>(b) * (EPI == EPI_DENOISE ? p.out_R : p.frames) + r;

            // EPI_MEL: a bin feeds the filters m0 and m0 + 1, and m0 never decreases as the bins ascend, so a row needs
            // two running sums: a0 (filter `cur`) and a1 (filter `cur + 1`).  When the table moves on to the next filter,
            // filter `cur` is complete and leaves as log(max(., clip)); the state lives in registers across the passes.
            int cur = 0;
            float a0 = 0.f, a1 = 0.f, nyq = 0.f;
            float* mel_out = nullptr;
            if constexpr (EPI == EPI_MEL) mel_out = p.c_out + (static_cast<size_t>(b) * p.n_mel) * p.frames + r;

            for (int pp = 0; pp < p.ppi; ++pp, ++acc_it) {
                const int pass = (item % groups) * p.ppi + pp;
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                mbar_wait(&tfull_bar[as], aph, 400 + as);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kBlockN;

                if constexpr (EPI == EPI_F32) {
                    float* dst = p.c_out + grow * p.n_total + pass * kBlockN;
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        if (live) {
                            float4* d4 = reinterpret_cast<float4*>(dst + ch * 32);
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                    __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                        }
                    }
                } else if constexpr (EPI == EPI_OLA) {
                    // row = hop block q of utterance b (R = frames + taps - 1 blocks per utterance); the frames covering it
                    // are q - j for the taps j with 0 <= q - j < frames: that set picks the envelope row (stft.py:111-121,
                    // audio_processing.py:45-47); blocks inside the trimmed L/2 at both ends are not stored (stft.py:127-128)
                    const int q_lo = p.taps >> 1;
                    const bool keep = live && r >= q_lo && r <= p.frames - 2 + p.taps - q_lo;
                    int mask = 0;
                    for (int j = 0; j < p.taps; ++j) mask |= (r - j >= 0 && r - j < p.frames) ? (1 << j) : 0;
                    const int hop = p.n_total;
                    float* dst = p.c_out + static_cast<size_t>(b) * hop * (p.frames - 1) + static_cast<size_t>(r - q_lo) * hop +
                                 pass * kBlockN;
                    const float* env = p.env_tab ? p.env_tab + static_cast<size_t>(mask) * hop + pass * kBlockN : nullptr;
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        if (keep) {
                            float o[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float a = __uint_as_float(v[j]);
                                if (env) {
                                    const float e = __ldg(env + ch * 32 + j);
                                    if (e > 1.17549435e-38f) a /= e;
                                    a *= p.ola_scale;
                                }
                                o[j] = a;
                            }
                            float4* d4 = reinterpret_cast<float4*>(dst + ch * 32);
#pragma unroll
                            for (int j = 0; j < 8; ++j) d4[j] = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                        }
                    }
                } else if constexpr (EPI == EPI_MEL) {
                    // columns 0..127 = Re of bins 128 pass .. +127, columns 128..255 = the matching Im (bin 0: Re of bin cp)
                    const float4* tab = reinterpret_cast<const float4*>(s_extra);
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t vr[32], vi[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, vr);
                        tmem_ld_32x32b_x32(taddr + 128 + ch * 32, vi);
                        tmem_ld_wait();
                        const int k0 = pass * 128 + ch * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float re = __uint_as_float(vr[j]), im = __uint_as_float(vi[j]);
                            float mag = sqrt_approx(fmaf(re, re, im * im));      // MUFU.SQRT, 2^-23 relative
                            if (k0 + j == 0) {                                   // two real bins: DC here, Nyquist at the end
                                mag = fabsf(re);
                                nyq = fabsf(im);
                            }
                            const float4 e = tab[k0 + j];                        // warp-uniform: broadcast
                            const int m0 = __float_as_int(e.x);
                            if (m0 > cur) {                                      // warp-uniform: the table moved on
                                if (live) mel_out[static_cast<size_t>(cur) * p.frames] = logf(fmaxf(a0, p.mel_clip));
                                a0 = a1;
                                if (m0 != cur + 1) {                             // narrow filters: a step of two or more
                                    if (live) mel_out[static_cast<size_t>(cur + 1) * p.frames] = logf(fmaxf(a1, p.mel_clip));
#pragma unroll 1
                                    for (int m = cur + 2; m < m0; ++m)
                                        if (live) mel_out[static_cast<size_t>(m) * p.frames] = logf(fmaxf(0.f, p.mel_clip));
                                    a0 = 0.f;
                                }
                                a1 = 0.f;
                                cur = m0;
                            }
                            a0 = fmaf(e.y, mag, a0);                             // bins without weight: m0 = -1, e.y = e.z = 0
                            a1 = fmaf(e.z, mag, a1);
                        }
                    }
                    if (pp == p.ppi - 1 && live) {
                        // filters cur, cur + 1 still hold sums; the Nyquist bin (no lower than any other bin's filters)
                        // joins here; filters above got no bin at all
                        const float4 en = tab[p.cp];
                        const int mn = __float_as_int(en.x);
                        for (int m = cur; m < p.n_mel; ++m) {
                            float v = m == cur ? a0 : (m == cur + 1 ? a1 : 0.f);
                            if (m == mn) v = fmaf(en.y, nyq, v);
                            if (m == mn + 1 && mn >= 0) v = fmaf(en.z, nyq, v);
                            mel_out[static_cast<size_t>(m) * p.frames] = logf(fmaxf(v, p.mel_clip));
                        }
                    }
                } else {
                    const int cp = p.cp;
                    const float* bs = reinterpret_cast<const float*>(s_extra);
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t vr[32], vi[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, vr);
                        tmem_ld_32x32b_x32(taddr + 128 + ch * 32, vi);
                        tmem_ld_wait();
                        const int k0 = pass * 128 + ch * 32;
                        uint32_t hr[16], lr[16], hi_[16], li[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float o[4];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                float re = __uint_as_float(vr[j + e]), im = __uint_as_float(vi[j + e]);
                                const int k = k0 + j + e;
                                if (k == 0) {                                    // DC and Nyquist: two real bins
                                    float z0 = 0.f, z1 = 0.f;
                                    denoise_bin(re, z0, bs[0]);
                                    denoise_bin(im, z1, bs[cp]);
                                } else {
                                    denoise_bin(re, im, bs[k]);                  // warp-uniform: broadcast
                                }
                                o[e] = re;
                                o[2 + e] = im;
                            }
                            const __nv_bfloat162 rh = __floats2bfloat162_rn(o[0], o[1]), ih = __floats2bfloat162_rn(o[2], o[3]);
                            const __nv_bfloat162 rl = __floats2bfloat162_rn(o[0] - __low2float(rh), o[1] - __high2float(rh));
                            const __nv_bfloat162 il = __floats2bfloat162_rn(o[2] - __low2float(ih), o[3] - __high2float(ih));
                            hr[j >> 1] = *reinterpret_cast<const uint32_t*>(&rh);
                            lr[j >> 1] = *reinterpret_cast<const uint32_t*>(&rl);
                            hi_[j >> 1] = *reinterpret_cast<const uint32_t*>(&ih);
                            li[j >> 1] = *reinterpret_cast<const uint32_t*>(&il);
                        }
                        if (live) {
                            __nv_bfloat16* hp = p.hi_out + grow * p.n_total + k0;
                            __nv_bfloat16* lp = p.lo_out + grow * p.n_total + k0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                reinterpret_cast<uint4*>(hp)[j] = make_uint4(hr[4 * j], hr[4 * j + 1], hr[4 * j + 2], hr[4 * j + 3]);
                                reinterpret_cast<uint4*>(lp)[j] = make_uint4(lr[4 * j], lr[4 * j + 1], lr[4 * j + 2], lr[4 * j + 3]);
                                reinterpret_cast<uint4*>(hp + cp)[j] = make_uint4(hi_[4 * j], hi_[4 * j + 1], hi_[4 * j + 2], hi_[4 * j + 3]);
                                reinterpret_cast<uint4*>(lp + cp)[j] = make_uint4(li[4 * j], li[4 * j + 1], li[4 * j + 2], li[4 * j + 3]);
                            }
                        }
                    }
                }
                // accumulator stage drained -> hand it back to the leader's MMA thread
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[as], 0));
            }
        }
    }

    // nobody may exit (or free TMEM) while the peer can still signal its barriers / read its operands
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc_2sm(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------ host side

// rows of K bf16 values, `rows` of them, `row_stride` elements apart (row_stride < K: overlapping STFT frames)
static int rows_map(CUtensorMap* m, const void* base, int K, long long rows, long long row_stride) {
    const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(row_stride) * 2};
    const uint32_t box[2] = {kBlockK, kBlockM};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}
static int basis_half_map(CUtensorMap* m, const void* base, int rows, int k) {
    const uint64_t dims[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(k) * 2};
    const uint32_t box[2] = {kBlockK, kHalfN};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}

template <int EPI>
static int launch(const CUtensorMap& hi, const CUtensorMap& lo, const CUtensorMap& w, const Params& p, cudaStream_t stream) {
    constexpr int smem = Smem<EPI>::kTotal;
    static_assert(smem <= 232448, "dynamic shared memory over the 227 KB per-CTA limit");
    auto kern = stft_pair_kernel<EPI>;
    WGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int n_items = ((p.n_tiles + 1) / 2) * (p.n_pass / p.ppi);
    int pairs = sm_count() / 2;
    if (n_items < pairs) pairs = n_items;
    kern<<<2 * pairs, kThreads, smem, stream>>>(hi, lo, w, p);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

// Frames of the padded signals as flat rows: hi / lo bf16 [B, R * hop], frame r of utterance b = flat row b R + r.
static int frame_maps(CUtensorMap* mhi, CUtensorMap* mlo, Params& p, const void* a_hi, const void* a_lo, int batch, int frames,
                      int R, int L, int hop) {
    WGB_REQUIRE(a_hi && a_lo, "null pointer");
    WGB_REQUIRE(batch > 0 && frames > 0 && hop > 0 && hop % 8 == 0, "batch, frames must be positive and hop a multiple of 8");
    WGB_REQUIRE(L > 0 && L % 256 == 0, "filter_length (%d) must be a multiple of 256", L);
    WGB_REQUIRE(static_cast<long long>(R) * hop >= static_cast<long long>(frames - 1) * hop + L,
                "row pitch R (%d hops) must cover the padded signal: (frames - 1) * hop + filter_length", R);
    const long long rows_total = static_cast<long long>(batch) * R;
    WGB_REQUIRE(rows_total < (1ll << 31) - 1024, "too many frames");
    // the last ceil(L / hop) - 1 flat rows would read past the end of the buffer: they are beyond the map (zero fill)
    // and never real rows (r >= frames there)
    const long long rows_in_bounds = (rows_total * hop - L) / hop + 1;
    if (int e = rows_map(mhi, a_hi, L, rows_in_bounds, hop)) return e;
    if (int e = rows_map(mlo, a_lo, L, rows_in_bounds, hop)) return e;
    p.rows_total = static_cast<int>(rows_total);
    p.n_tiles = ceil_div(p.rows_total, kBlockM);
    p.R = R;
    p.frames = frames;
    p.seg_chunks = L / kBlockK;
    p.n_chunks = 3 * p.seg_chunks;
    p.taps = 1;
    p.cp = L / 2;
    return WGB_OK;
}

}  // namespace stft2

// TacotronSTFT.mel_spectrogram (layers.py:63-79) as one kernel.  a_hi / a_lo: reflect-padded signals, bf16 hi / lo parts,
// [B, R * hop]; w3_paired bf16 [L][3L]: the forward basis in the paired order of this file (pass p: Re rows of bins
// 128p..128p+127, then their Im rows, with the Re row of bin L/2 in the Im slot of bin 0), split [hi | hi | lo];
// mel_table [L/2 + 1] float4 {first filter, w_first, w_next, 0} -- over the bins that carry weight the first-filter
// index must not decrease (triangular filters overlapping pairwise; the packer checks it);
// n_pass: passes that hold a bin with non-zero weight; out fp32 [B, n_mel, frames].
int tc2_stft_mel(const void* a_hi, const void* a_lo, const void* w3_paired, const void* mel_table, float* out, int batch,
                 int frames, int R, int L, int hop, int n_pass, int n_mel, float clip, cudaStream_t stream) {
    using namespace stft2;
    WGB_REQUIRE(w3_paired && mel_table && out, "null pointer");
    WGB_REQUIRE(L / 2 <= kMaxBins, "filter_length (%d) above %d: use the one-CTA kernel", L, 2 * kMaxBins);
    WGB_REQUIRE(n_mel >= 1, "n_mel (%d) must be positive", n_mel);
    Params p{};
    p.l2_hint = tuning_get("stft_l2_hint");
    CUtensorMap mhi, mlo, mw;
    if (int e = frame_maps(&mhi, &mlo, p, a_hi, a_lo, batch, frames, R, L, hop)) return e;
    WGB_REQUIRE(n_pass >= 1 && n_pass <= L / kBlockN, "n_pass (%d) must be in 1..%d", n_pass, L / kBlockN);
    p.n_pass = n_pass; p.ppi = n_pass;                    // one CTA pair runs all passes of its tiles (running sums in registers)
    p.c_out = out; p.mel_table = static_cast<const float4*>(mel_table); p.n_mel = n_mel; p.mel_clip = clip;
    if (int e = basis_half_map(&mw, w3_paired, L, 3 * L)) return e;
    return launch<EPI_MEL>(mhi, mlo, mw, p, stream);
}

// Denoiser.forward's transform + spectral subtraction (denoiser.py:36-38): writes the bf16 hi / lo operands
// [B * frames, L] of the inverse-basis GEMM (column order of this file).  bias_spec fp32 [L/2 + 1].
int tc2_stft_denoise(const void* a_hi, const void* a_lo, const void* w3_paired, const float* bias_spec, float strength,
                     void* hi_out, void* lo_out, int batch, int frames, int R, int L, int hop, int out_R, cudaStream_t stream) {
    using namespace stft2;
    WGB_REQUIRE(w3_paired && bias_spec && hi_out && lo_out, "null pointer");
    WGB_REQUIRE(L / 2 <= kMaxBins, "filter_length (%d) above %d: use the one-CTA kernel", L, 2 * kMaxBins);
    Params p{};
    p.l2_hint = tuning_get("stft_l2_hint");
    CUtensorMap mhi, mlo, mw;
    if (int e = frame_maps(&mhi, &mlo, p, a_hi, a_lo, batch, frames, R, L, hop)) return e;
    WGB_REQUIRE(out_R >= frames, "out_R (%d) must be >= frames (%d)", out_R, frames);
    p.n_pass = L / kBlockN; p.ppi = 1; p.n_total = L; p.out_R = out_R;
    p.hi_out = static_cast<__nv_bfloat16*>(hi_out); p.lo_out = static_cast<__nv_bfloat16*>(lo_out);
    p.spec_bias = bias_spec; p.strength = strength;
    if (int e = basis_half_map(&mw, w3_paired, L, 3 * L)) return e;
    return launch<EPI_DENOISE>(mhi, mlo, mw, p, stream);
}

// C[rows, N] fp32 = (A_hi + A_lo)[rows, K] (W_hi + W_lo)^T[N, K] to fp32-grade accuracy; a_hi / a_lo bf16 [rows, K]
// (contiguous rows), w3 bf16 [N][3K] = [W_hi | W_hi | W_lo]; N % 256 == 0, K % 64 == 0.
int tc2_gemm_split3(const void* a_hi, const void* a_lo, const void* w3, float* c, long long rows, int N, int K,
                    cudaStream_t stream) {
    using namespace stft2;
    WGB_REQUIRE(a_hi && a_lo && w3 && c, "null pointer");
    WGB_REQUIRE(rows > 0 && rows < (1ll << 31) - 1024, "bad row count");
    WGB_REQUIRE(N > 0 && N % kBlockN == 0 && K > 0 && K % kBlockK == 0, "N must be a multiple of 256 and K of 64 (N=%d K=%d)", N, K);
    Params p{};
    p.l2_hint = tuning_get("stft_l2_hint");
    p.rows_total = static_cast<int>(rows);
    p.n_tiles = ceil_div(p.rows_total, kBlockM);
    p.R = p.rows_total; p.frames = p.rows_total;          // compact rows
    p.seg_chunks = K / kBlockK; p.n_chunks = 3 * p.seg_chunks; p.taps = 1;
    p.n_pass = N / kBlockN; p.ppi = 1; p.n_total = N;
    p.c_out = c;
    CUtensorMap mhi, mlo, mw;
    if (int e = rows_map(&mhi, a_hi, K, rows, K)) return e;
    if (int e = rows_map(&mlo, a_lo, K, rows, K)) return e;
    if (int e = basis_half_map(&mw, w3, N, 3 * K)) return e;
    return launch<EPI_F32>(mhi, mlo, mw, p, stream);
}

// STFT.inverse's conv_transpose1d + overlap-add + window-sum normalisation + L/hop scale + L/2 trim (stft.py:105-128) as
// ONE GEMM.  s_hi / s_lo: recombined spectra as bf16 hi / lo parts [B, frames + taps - 1, L] (taps = L / hop) whose last
// taps - 1 rows per utterance are ZERO (the frames before the first / after the last one); w_ola bf16 [hop][taps * 3L]:
// tap j's columns = split-bf16 [hi | hi | lo] of the inverse basis rows n = j hop .. (j + 1) hop - 1; env_tab fp32
// [2^taps][hop] = the window-sum envelope for every set of covering frames (bit j set: frame q - j exists), built by the
// host exactly like audio_processing.py:45-47 accumulates it, or NULL for window=None (no division, no scale);
// out fp32 [B, hop * (frames - 1)].  hop % 256 == 0, L % hop == 0, L / hop <= 8.
int tc2_istft_ola(const void* s_hi, const void* s_lo, const void* w_ola, const float* env_tab, float* out, int batch,
                  int frames, int L, int hop, cudaStream_t stream) {
    using namespace stft2;
    WGB_REQUIRE(s_hi && s_lo && w_ola && out, "null pointer");
    WGB_REQUIRE(batch > 0 && frames > 1, "batch must be positive and frames > 1");
    WGB_REQUIRE(L > 0 && L % kBlockK == 0 && hop > 0 && hop % kBlockN == 0 && L % hop == 0 && L / hop <= 8,
                "need hop %% 256 == 0, filter_length %% hop == 0 and at most 8 taps (L=%d hop=%d)", L, hop);
    const int taps = L / hop;
    const long long rows = static_cast<long long>(batch) * (frames + taps - 1);
    WGB_REQUIRE(rows < (1ll << 31) - 1024, "too many frames");
    Params p{};
    p.l2_hint = tuning_get("stft_l2_hint");
    p.rows_total = static_cast<int>(rows);
    p.n_tiles = ceil_div(p.rows_total, kBlockM);
    p.R = frames + taps - 1; p.frames = frames;
    p.seg_chunks = L / kBlockK; p.taps = taps; p.n_chunks = taps * 3 * p.seg_chunks;
    p.n_pass = hop / kBlockN; p.ppi = 1; p.n_total = hop;
    p.c_out = out; p.env_tab = env_tab; p.ola_scale = static_cast<float>(L) / static_cast<float>(hop);
    CUtensorMap mhi, mlo, mw;
    if (int e = rows_map(&mhi, s_hi, L, rows, L)) return e;
    if (int e = rows_map(&mlo, s_lo, L, rows, L)) return e;
    if (int e = basis_half_map(&mw, w_ola, hop, taps * 3 * L)) return e;
    return launch<EPI_OLA>(mhi, mlo, mw, p, stream);
}

}  // namespace wgb
