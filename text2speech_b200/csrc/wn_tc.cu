// WN-layer GEMMs on tcgen05 / TMEM, operands staged by TMA (sm_100a only).
//
// Activations are channels-last bf16 [B, T, C]; a tile is 128 consecutive group steps of one
// utterance (UMMA M = 128, one TMEM lane per time step) and one pass produces 256 output columns
// (UMMA N = 256) into one of two 256-column TMEM accumulator stages, so the epilogue of pass i
// overlaps the MMAs of pass i+1.  K is streamed in 64-element (128 B, SWIZZLE_128B) chunks through a
// 4-stage TMA->mbarrier->UMMA ring.  Warp roles: warp 0 = TMA producer (one elected lane),
// warp 1 = TMEM owner + MMA issuer (one lane), warps 2..5 = epilogue (one TMEM lane quarter each).
//
//   MODE_GATE      in_layers[i] (k=3, dilation d) + cond_layers[i] + bias -> tanh*sigmoid -> acts
//                  (reference glow.py:159-162).  The three taps are three TMA boxes at t0-d, t0, t0+d
//                  (out-of-range rows zero-filled by TMA == the conv's zero padding); the cond 1x1
//                  is 10 more K chunks into the same accumulator, so the "add" costs nothing.
//                  Weight rows are permuted on the host so a pass holds tanh rows c..c+127 in
//                  columns 0..127 and the matching sigmoid rows in columns 128..255.
//   MODE_RES       res half of res_skip_layers[i]: h_out = h_in + W_res acts + b  (glow.py:164-166)
//   MODE_SKIP_END  sum_i W_skip_i acts_i accumulated over all layers in TMEM as ONE K=8*512 GEMM
//                  (same FLOPs as glow.py:171-174, fp32 accumulation, no HBM read-modify-write),
//                  then end 1x1 (glow.py:175), affine coupling and invertible 1x1 conv
//                  (glow.py:277-282 infer / :241-246 forward) in the epilogue, all fp32.
#include "common.cuh"
#include "ptx.cuh"

namespace wgb {

constexpr int kBlockM = 128;
constexpr int kBlockN = 256;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;
constexpr int kBBytes = kBlockN * kBlockK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;
constexpr int kThreads = 192;
constexpr int kNCh = 512;        // WN channels the tensor-core path is specialised for
constexpr int kNCond = 640;      // n_mel_channels * n_group

enum Mode { MODE_GATE = 0, MODE_RES = 1, MODE_SKIP_END = 2, MODE_PLAIN = 3, MODE_STFT_MEL = 4 };
constexpr int kMelMax = 80;          // mel channels the fused STFT->mel epilogue keeps per row (+1 pad column)

struct TcParams {
    int batch, T, tiles_per_b, n_tiles;
    int n_pass, ppi, n_chunks, dilation;
    const float* bias;            // per packed output column (GATE, RES)
    __nv_bfloat16* acts_out;      // GATE   [B,T,512]
    const __nv_bfloat16* h_in;    // RES    [B,T,512]
    __nv_bfloat16* h_out;         // RES    [B,T,512]
    const float* w_end;           // SKIP_END [512][8] fp32 (rows >= 2*n_half zero)
    const float* b_end;           // SKIP_END [8] (skip biases already folded in)
    float* x;                     // SKIP_END flow state [B,T,8] fp32, active channels = last 2*n_half
    const float* w_mix;           // SKIP_END infer: W^-1 as [8][8] row-major (top-left CxC used)
    float* log_s;                 // SKIP_END forward: [B,n_half,T]
    void* c_out;                  // PLAIN  [B,T,N] fp32 (DIR=0) or bf16 (DIR=1); bias may be null
                                  //        DIR=2 (paired Re/Im columns): magnitudes fp32 [B,T,N/2]
                                  //        DIR=3 (paired): spectral subtraction, bf16 hi part [B,T,N] (Re | Im halves)
    void* c_out2;                 // PLAIN  DIR=3: bf16 lo part
    const float* spec_bias;       // PLAIN  DIR=3: denoiser bias spectrum [cutoff]
    float strength;               // PLAIN  DIR=3
    int cutoff;                   // PLAIN  DIR=3: bins < cutoff are real spectrum bins, the rest padding
    const float4* mel_table;      // STFT_MEL [cp]: {first filter index (as float), weight in that filter, weight in the
                                  //          next filter, 0}: every bin lies in at most two adjacent triangular filters
    int n_mel;                    // STFT_MEL (<= 80)
    float mel_clip;               // STFT_MEL log(max(., clip))
    int n_total;                  // PLAIN  N (multiple of 256)
    int seg_chunks, seg_mask;     // PLAIN  K is split in segments of seg_chunks chunks; bit s of seg_mask selects
                                  //        map_a1 (else map_a0) for segment s (split-bf16 hi/lo operands)
    int seg_shift0, seg_dshift;   // PLAIN  segment s reads A rows t + seg_shift0 + s*seg_dshift (conv taps; rows outside
                                  //        [0,T) are zero-filled by TMA = zero padding); 0, 0 for a plain GEMM
    int seg_bstride;              // PLAIN  segment s reads utterance b + s*seg_bstride (segments stacked as [n_seg, B, T, C]); else 0
    int act;                      // PLAIN  DIR 0/1 epilogue: 0 none, 1 tanh, 2 relu
    const void* res;              // PLAIN  DIR 0/1: optional [B,T,N] tensor (fp32 for DIR 0, bf16 for DIR 1) added to the
                                  //        result before the store; may alias c_out (accumulate / residual update in place)
};

// Shared-memory carve-up (offsets from a 1024 B aligned base): [TMA ring][mode extra][barriers].
//   RES:      3 stages + 64 KB h tile (TMA-loaded h_in, updated in place, TMA-stored as h_out) + 2 KB bias
//   SKIP_END: 4 stages + 16 KB W_end^T
//   GATE:     4 stages + 4 KB bias (sigmoid half pre-halved)
//   STFT_MEL: 3 stages + 41.5 KB per-row mel accumulators [128][81] + the sparse mel table [cp] float4
//   others:   4 stages
template <int MODE>
struct SmemLayout {
    static constexpr int kStages = (MODE == MODE_RES || MODE == MODE_STFT_MEL) ? 3 : 4;
    static constexpr int kRing = kStages * kStageBytes;
    static constexpr int kExtraOff = kRing;
    static constexpr int kExtraBytes = MODE == MODE_RES ? kBlockM * kBlockN * 2 + kNCh * 4
                                       : (MODE == MODE_SKIP_END ? kNCh * 8 * 4
                                          : (MODE == MODE_GATE ? 2 * kNCh * 4
                                             : (MODE == MODE_STFT_MEL ? kBlockM * (kMelMax + 1) * 4 + 1024 * 16 : 0)));
    static constexpr int kBarOff = kExtraOff + kExtraBytes;
    static constexpr int kBarBytes = 256;     // full[4] empty[4] tfull[2] tempty[2] hfull hempty tmem_slot
    static constexpr int kTotal = 1024 + kBarOff + kBarBytes;
};

template <int MODE, int NHALF, int DIR>
__global__ void __launch_bounds__(kThreads, 1)
wn_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
             const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_c,
             const TcParams p) {
    using SL = SmemLayout<MODE>;
    constexpr int kStages = SL::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::kBarOff);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* hfull_bar = tempty_bar + 2;       // RES: h_in tile landed in smem
    uint64_t* hempty_bar = hfull_bar + 1;       // RES: h tile buffer free again (TMA store has read it)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hempty_bar + 1);
    uint8_t* s_extra = smem + SL::kExtraOff;
    float* s_wend = reinterpret_cast<float*>(s_extra);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_b);
        if constexpr (MODE == MODE_RES) tma_prefetch_desc(&map_c);
        mbar_init(hfull_bar, 1);
        mbar_init(hempty_bar, 1);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    if (MODE == MODE_SKIP_END && warp >= 2) {
        for (int i = threadIdx.x - 64; i < kNCh * 8; i += 128) s_wend[i] = p.w_end[i];
    }
    if (MODE == MODE_RES && warp >= 2) {
        float* sb = reinterpret_cast<float*>(s_extra + kBlockM * kBlockN * 2);
        for (int i = threadIdx.x - 64; i < kNCh; i += 128) sb[i] = p.bias[i];
    }
    if (MODE == MODE_STFT_MEL && warp >= 2) {
        float4* tab = reinterpret_cast<float4*>(s_extra + kBlockM * (kMelMax + 1) * 4);
        for (int i = threadIdx.x - 64; i < (p.n_total >> 1); i += 128) tab[i] = p.mel_table[i];
    }
    if (MODE == MODE_GATE && warp >= 2) {      // sigmoid(b) = 0.5 tanh(b/2) + 0.5: sigmoid columns carry b/2
        for (int i = threadIdx.x - 64; i < 2 * kNCh; i += 128) s_wend[i] = p.bias[i] * ((i & 128) ? 0.5f : 1.f);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int groups = p.n_pass / p.ppi;
    const int n_items = p.n_tiles * groups;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, hph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int tile = item / groups;
                const int b = tile / p.tiles_per_b;
                const int t0 = (tile % p.tiles_per_b) * kBlockM;
                for (int pp = 0; pp < p.ppi; ++pp) {
                    const int pass = (item % groups) * p.ppi + pp;
                    for (int kc = 0; kc < p.n_chunks; ++kc) {
                        mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
                        uint8_t* sa = smem + s * kStageBytes;
                        uint8_t* sb = sa + kABytes;
                        mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                        if constexpr (MODE == MODE_GATE) {
                            if (kc < 24) {
                                const int tap = kc >> 3;
                                tma_load_3d(sa, &map_a0, &full_bar[s], (kc & 7) * kBlockK,
                                            t0 + (tap - 1) * p.dilation, b);
                            } else {
                                tma_load_3d(sa, &map_a1, &full_bar[s], (kc - 24) * kBlockK, t0, b);
                            }
                        } else if constexpr (MODE == MODE_RES) {
                            tma_load_3d(sa, &map_a0, &full_bar[s], kc * kBlockK, t0, b);
                        } else if constexpr (MODE == MODE_PLAIN || MODE == MODE_STFT_MEL) {
                            const int seg = kc / p.seg_chunks;
                            tma_load_3d(sa, ((p.seg_mask >> seg) & 1) ? &map_a1 : &map_a0, &full_bar[s],
                                        (kc - seg * p.seg_chunks) * kBlockK, t0 + p.seg_shift0 + seg * p.seg_dshift,
                                        b + seg * p.seg_bstride);
                        } else {
                            tma_load_3d(sa, &map_a0, &full_bar[s], (kc & 7) * kBlockK, t0, (kc >> 3) * p.batch + b);
                        }
                        tma_load_2d(sb, &map_b, &full_bar[s], kc * kBlockK, pass * kBlockN);
                        if (++s == kStages) { s = 0; ph ^= 1; }
                    }
                    if constexpr (MODE == MODE_RES) {
                        // h_in[t0:+128, pass*256:+256] -> smem as four 64-column SWIZZLE_128B boxes.  Issued AFTER
                        // this pass's K loads: this thread is in-order, and waiting here for the previous pass's
                        // epilogue must not hold back the operand stream that overlaps that epilogue.
                        mbar_wait(hempty_bar, hph ^ 1, 500);
                        mbar_arrive_expect_tx(hfull_bar, kBlockM * kBlockN * 2);
#pragma unroll
                        for (int j = 0; j < kBlockN / kBlockK; ++j)
                            tma_load_3d(s_extra + j * kABytes, &map_a1, hfull_bar, pass * kBlockN + j * kBlockK, t0, b);
                        hph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(kBlockM, kBlockN);
            int s = 0;
            uint32_t ph = 0;
            uint32_t acc_it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
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
                            umma_bf16_ss(d_tmem, umma_desc_sw128(a_addr + k * kUmmaK * 2),
                                         umma_desc_sw128(b_addr + k * kUmmaK * 2), idesc, (kc | k) != 0);
                        }
                        umma_commit(&empty_bar[s]);
                        if (++s == kStages) { s = 0; ph ^= 1; }
                    }
                    umma_commit(&tfull_bar[as]);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                 // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        uint32_t hph = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int tile = item / groups;
            const int b = tile / p.tiles_per_b;
            const int t = (tile % p.tiles_per_b) * kBlockM + row;
            const bool live = t < p.T;
            const size_t grow = static_cast<size_t>(b) * p.T + t;
            float outv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) outv[j] = 0.f;

            for (int pp = 0; pp < p.ppi; ++pp, ++acc_it) {
                const int pass = (item % groups) * p.ppi + pp;
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                mbar_wait(&tfull_bar[as], aph, 400 + as);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kBlockN;

                if constexpr (MODE == MODE_GATE) {
                    const float4* bt4 = reinterpret_cast<const float4*>(s_wend + pass * kBlockN);
                    const float4* bs4 = bt4 + 32;
                    __nv_bfloat16* dst = p.acts_out + grow * kNCh + pass * 128;
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t vt[32], vs[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, vt);
                        tmem_ld_32x32b_x32(taddr + 128 + ch * 32, vs);
                        tmem_ld_wait();
                        uint32_t packed[16];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bt = bt4[ch * 8 + j], bs = bs4[ch * 8 + j];     // warp-uniform: broadcast LDS.128
                            const float g0 = gate_tanh_sigmoid_h(__uint_as_float(vt[4 * j]) + bt.x,
                                                                 fmaf(__uint_as_float(vs[4 * j]), 0.5f, bs.x));
                            const float g1 = gate_tanh_sigmoid_h(__uint_as_float(vt[4 * j + 1]) + bt.y,
                                                                 fmaf(__uint_as_float(vs[4 * j + 1]), 0.5f, bs.y));
                            const float g2 = gate_tanh_sigmoid_h(__uint_as_float(vt[4 * j + 2]) + bt.z,
                                                                 fmaf(__uint_as_float(vs[4 * j + 2]), 0.5f, bs.z));
                            const float g3 = gate_tanh_sigmoid_h(__uint_as_float(vt[4 * j + 3]) + bt.w,
                                                                 fmaf(__uint_as_float(vs[4 * j + 3]), 0.5f, bs.w));
                            __nv_bfloat162 h01 = __floats2bfloat162_rn(g0, g1), h23 = __floats2bfloat162_rn(g2, g3);
                            packed[2 * j] = *reinterpret_cast<uint32_t*>(&h01);
                            packed[2 * j + 1] = *reinterpret_cast<uint32_t*>(&h23);
                        }
                        if (live) {
                            uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                        }
                    }
                } else if constexpr (MODE == MODE_RES) {
                    // h tile sits in smem in the TMA SWIZZLE_128B layout (four [128 x 64] boxes): row r of box j
                    // is at j*16K + r*128, its 16 B chunk c at ((c ^ (r & 7)) << 4).  Each thread updates its own
                    // row in place; the whole tile then leaves through one TMA store (coalesced, async).
                    const float4* bias4 = reinterpret_cast<const float4*>(s_extra + kBlockM * kBlockN * 2) + pass * (kBlockN / 4);
                    mbar_wait(hfull_bar, hph, 600);
                    hph ^= 1;
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, v);
                        uint8_t* rowp = s_extra + (ch >> 1) * kABytes + row * 128;
                        uint4 old[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            old[i] = *reinterpret_cast<const uint4*>(rowp + ((((ch & 1) * 4 + i) ^ (row & 7)) << 4));
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint32_t* ow = reinterpret_cast<const uint32_t*>(&old[i]);
                            const float4 b0 = bias4[ch * 8 + i * 2], b1 = bias4[ch * 8 + i * 2 + 1];   // broadcast LDS.128
                            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                            uint32_t pk[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const __nv_bfloat162 o2 = *reinterpret_cast<const __nv_bfloat162*>(&ow[j]);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(
                                    __uint_as_float(v[i * 8 + 2 * j]) + bb[2 * j] + __low2float(o2),
                                    __uint_as_float(v[i * 8 + 2 * j + 1]) + bb[2 * j + 1] + __high2float(o2));
                                pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                            }
                            *reinterpret_cast<uint4*>(rowp + ((((ch & 1) * 4 + i) ^ (row & 7)) << 4)) =
                                make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                } else if constexpr (MODE == MODE_STFT_MEL) {
                    // TacotronSTFT.mel_spectrogram in one kernel (stft.py:85-97, layers.py:77-78, audio_processing.py:76):
                    // Re/Im-paired basis rows -> |X| per bin -> scattered into this row's <= 80 mel accumulators in
                    // shared memory (a bin feeds at most two adjacent filters) -> log(max(., clip)) after the last pass
                    float* acc = reinterpret_cast<float*>(s_extra) + row * (kMelMax + 1);     // stride 81: conflict-free
                    const float4* tab = reinterpret_cast<const float4*>(s_extra + kBlockM * (kMelMax + 1) * 4);
                    if (pp == 0) {
                        for (int m = 0; m <= kMelMax; ++m) acc[m] = 0.f;
                    }
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t vr[32], vi[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, vr);
                        tmem_ld_32x32b_x32(taddr + 128 + ch * 32, vi);
                        tmem_ld_wait();
                        const int k0 = pass * 128 + ch * 32;
#pragma unroll 8
                        for (int j = 0; j < 32; ++j) {
                            const float re = __uint_as_float(vr[j]), im = __uint_as_float(vi[j]);
                            const float mag = sqrtf(re * re + im * im);
                            const float4 e = tab[k0 + j];                                      // warp-uniform: broadcast
                            const int m0 = static_cast<int>(e.x);
                            acc[m0] = fmaf(e.y, mag, acc[m0]);
                            acc[m0 + 1] = fmaf(e.z, mag, acc[m0 + 1]);
                        }
                    }
                    if (pp == p.ppi - 1 && live) {
                        float* out = static_cast<float*>(p.c_out) + (static_cast<size_t>(b) * p.n_mel) * p.T + t;
                        for (int m = 0; m < p.n_mel; ++m) out[static_cast<size_t>(m) * p.T] = logf(fmaxf(acc[m], p.mel_clip));
                    }
                } else if constexpr (MODE == MODE_PLAIN && DIR >= 2) {
                    // STFT with Re/Im-paired basis rows: columns 0..127 of the pass are Re of bins 128 pass .. +127,
                    // columns 128..255 the matching Im (stft.py:85-97), so |X| and the Denoiser's spectral
                    // subtraction (denoiser.py:36-38 + stft.py:102-103) happen here instead of in extra passes over HBM
                    const int cp = p.n_total >> 1;
#pragma unroll 1
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t vr[32], vi[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, vr);
                        tmem_ld_32x32b_x32(taddr + 128 + ch * 32, vi);
                        tmem_ld_wait();
                        const int k0 = pass * 128 + ch * 32;
                        if constexpr (DIR == 2) {
                            float m[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float re = __uint_as_float(vr[j]), im = __uint_as_float(vi[j]);
                                m[j] = sqrtf(re * re + im * im);
                            }
                            if (live) {
                                float4* d4 = reinterpret_cast<float4*>(static_cast<float*>(p.c_out) + grow * cp + k0);
#pragma unroll
                                for (int j = 0; j < 8; ++j) d4[j] = make_float4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]);
                            }
                        } else {
                            uint32_t hr[16], lr[16], hi_[16], li[16];
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                float o[4];
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    float re = __uint_as_float(vr[j + e]), im = __uint_as_float(vi[j + e]);
                                    if (k0 + j + e < p.cutoff) {
                                        const float mag = sqrtf(re * re + im * im);
                                        const float m2 = fmaxf(mag - __ldg(p.spec_bias + k0 + j + e) * p.strength, 0.f);
                                        const float g = mag > 0.f ? m2 / mag : 0.f;
                                        re = mag > 0.f ? re * g : m2;      // atan2(0,0) = 0 -> cos = 1, sin = 0
                                        im = im * g;
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
                                __nv_bfloat16* hp = static_cast<__nv_bfloat16*>(p.c_out) + grow * p.n_total + k0;
                                __nv_bfloat16* lp = static_cast<__nv_bfloat16*>(p.c_out2) + grow * p.n_total + k0;
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
                } else if constexpr (MODE == MODE_PLAIN) {
                    const size_t off = grow * p.n_total + pass * kBlockN;
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        float f[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            f[j] = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + pass * kBlockN + ch * 32 + j) : 0.f);
                            if (p.act == 1) f[j] = tanhf(f[j]);
                            else if (p.act == 2) f[j] = fmaxf(f[j], 0.f);
                        }
                        if (p.res && live) {
                            if constexpr (DIR == 0) {
                                const float4* r4 = reinterpret_cast<const float4*>(static_cast<const float*>(p.res) + off + ch * 32);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 r = r4[j];
                                    f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
                                }
                            } else {
                                const uint4* r4 = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.res) + off + ch * 32);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const uint4 r = r4[j];
                                    const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rw[e]);
                                        f[8 * j + 2 * e] += __low2float(h2);
                                        f[8 * j + 2 * e + 1] += __high2float(h2);
                                    }
                                }
                            }
                        }
                        if (live) {
                            if constexpr (DIR == 0) {
                                float4* d4 = reinterpret_cast<float4*>(static_cast<float*>(p.c_out) + off + ch * 32);
#pragma unroll
                                for (int j = 0; j < 8; ++j) d4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                            } else {
                                uint4* d4 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.c_out) + off + ch * 32);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
                                    __nv_bfloat162 h1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
                                    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
                                    __nv_bfloat162 h3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
                                    d4[j] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                       *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
                                }
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        const float4* w4 = reinterpret_cast<const float4*>(s_wend + (pass * kBlockN + ch * 32) * 8);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float a = __uint_as_float(v[j]);
                            const float4 w0 = w4[2 * j], w1 = w4[2 * j + 1];
                            outv[0] = fmaf(a, w0.x, outv[0]);
                            outv[1] = fmaf(a, w0.y, outv[1]);
                            outv[2] = fmaf(a, w0.z, outv[2]);
                            outv[3] = fmaf(a, w0.w, outv[3]);
                            outv[4] = fmaf(a, w1.x, outv[4]);
                            outv[5] = fmaf(a, w1.y, outv[5]);
                            outv[6] = fmaf(a, w1.z, outv[6]);
                            outv[7] = fmaf(a, w1.w, outv[7]);
                        }
                    }
                }
                // accumulator stage drained -> hand it back to the MMA warp
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[as]);
                if constexpr (MODE == MODE_RES) {
                    fence_proxy_async_smem();                       // st.shared -> visible to the TMA store
                    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps
                    if (warp == 2 && lane == 0) {
                        const int t0 = (tile % p.tiles_per_b) * kBlockM;
#pragma unroll
                        for (int j = 0; j < kBlockN / kBlockK; ++j)
                            tma_store_3d(&map_c, s_extra + j * kABytes, pass * kBlockN + j * kBlockK, t0, b);
                        tma_store_commit();
                        tma_store_wait_read0();
                        mbar_arrive(hempty_bar);
                    }
                }
            }

            if constexpr (MODE == MODE_SKIP_END) if (live) {
                constexpr int C = 2 * NHALF, BASE = 8 - C;
                float* xr = p.x + grow * 8;
                float xv[8];
                *reinterpret_cast<float4*>(&xv[0]) = *reinterpret_cast<const float4*>(xr);
                *reinterpret_cast<float4*>(&xv[4]) = *reinterpret_cast<const float4*>(xr + 4);
#pragma unroll
                for (int j = 0; j < 8; ++j) outv[j] += __ldg(p.b_end + j);
                if constexpr (DIR == 0) {                        // infer (glow.py:279-282)
                    float xin[C];
#pragma unroll
                    for (int j = 0; j < NHALF; ++j) {
                        xin[j] = xv[BASE + j];
                        xin[NHALF + j] = (xv[BASE + NHALF + j] - outv[j]) * expf(-outv[NHALF + j]);
                    }
#pragma unroll
                    for (int i = 0; i < C; ++i) {
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < C; ++c) acc = fmaf(__ldg(p.w_mix + i * 8 + c), xin[c], acc);
                        xv[BASE + i] = acc;
                    }
                } else {                                         // forward (glow.py:241-246)
#pragma unroll
                    for (int j = 0; j < NHALF; ++j) {
                        const float ls = outv[NHALF + j];
                        xv[BASE + NHALF + j] = expf(ls) * xv[BASE + NHALF + j] + outv[j];
                        p.log_s[(static_cast<size_t>(b) * NHALF + j) * p.T + t] = ls;
                    }
                }
                *reinterpret_cast<float4*>(xr) = *reinterpret_cast<const float4*>(&xv[0]);
                *reinterpret_cast<float4*>(xr + 4) = *reinterpret_cast<const float4*>(&xv[4]);
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------ host side

static int act_map(CUtensorMap* m, const void* base, int channels, int T, int batches) {
    const uint64_t dims[3] = {static_cast<uint64_t>(channels), static_cast<uint64_t>(T), static_cast<uint64_t>(batches)};
    const uint64_t strides[2] = {static_cast<uint64_t>(channels) * 2, static_cast<uint64_t>(channels) * 2 * T};
    const uint32_t box[3] = {kBlockK, kBlockM, 1};
    return make_tmap_bf16(m, base, 3, dims, strides, box);
}
static int weight_map(CUtensorMap* m, const void* base, int rows, int k) {
    const uint64_t dims[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(k) * 2};
    const uint32_t box[2] = {kBlockK, kBlockN};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}

template <int MODE, int NHALF, int DIR>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& bm, const CUtensorMap& cm, TcParams p,
                  cudaStream_t stream) {
    constexpr int smem = SmemLayout<MODE>::kTotal;
    static_assert(smem <= 232448, "dynamic shared memory over the 227 KB per-CTA limit");
    auto kern = wn_tc_kernel<MODE, NHALF, DIR>;
    WGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int items = p.n_tiles * (p.n_pass / p.ppi);
    const int grid = items < sm_count() ? items : sm_count();
    kern<<<grid, kThreads, smem, stream>>>(a0, a1, bm, cm, p);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

static int fill_common(TcParams& p, int batch, int T) {
    WGB_REQUIRE(batch > 0 && T > 0, "batch (%d) and T (%d) must be positive", batch, T);
    p.batch = batch;
    p.T = T;
    p.tiles_per_b = ceil_div(T, kBlockM);
    p.n_tiles = batch * p.tiles_per_b;
    return WGB_OK;
}

int tc_wn_gate(const void* h, const void* spect, const void* w_packed, const float* bias, void* acts, int batch, int T,
               int dilation, cudaStream_t stream) {
    WGB_REQUIRE(h && spect && w_packed && bias && acts, "null pointer");
    WGB_REQUIRE(dilation >= 1, "dilation must be >= 1");
    TcParams p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = 4; p.ppi = 1; p.n_chunks = (3 * kNCh + kNCond) / kBlockK; p.dilation = dilation;
    p.bias = bias;
    p.acts_out = static_cast<__nv_bfloat16*>(acts);
    CUtensorMap ma0, ma1, mb;
    if (int e = act_map(&ma0, h, kNCh, T, batch)) return e;
    if (int e = act_map(&ma1, spect, kNCond, T, batch)) return e;
    if (int e = weight_map(&mb, w_packed, 2 * kNCh, 3 * kNCh + kNCond)) return e;
    return launch<MODE_GATE, 0, 0>(ma0, ma1, mb, ma0, p, stream);
}

int tc_wn_res(const void* acts, const void* w_res, const float* bias, const void* h_in, void* h_out, int batch, int T,
              cudaStream_t stream) {
    WGB_REQUIRE(acts && w_res && bias && h_in && h_out, "null pointer");
    TcParams p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = 2; p.ppi = 1; p.n_chunks = kNCh / kBlockK;
    p.bias = bias;
    p.h_in = static_cast<const __nv_bfloat16*>(h_in);
    p.h_out = static_cast<__nv_bfloat16*>(h_out);
    CUtensorMap ma0, mhi, mho, mb;
    if (int e = act_map(&ma0, acts, kNCh, T, batch)) return e;
    if (int e = act_map(&mhi, h_in, kNCh, T, batch)) return e;
    if (int e = act_map(&mho, h_out, kNCh, T, batch)) return e;
    if (int e = weight_map(&mb, w_res, kNCh, kNCh)) return e;
    return launch<MODE_RES, 0, 0>(ma0, mhi, mb, mho, p, stream);
}

// C[b,t,n] = sum_k A[b,t,k] W[n,k] + bias[n]; A bf16 [B,T,K] (K % 64 == 0), W bf16 [N,K] (N % 256 == 0).
int tc_gemm_plain(const void* a, const void* w, const float* bias, void* c, int out_bf16, int batch, int T, int N, int K,
                  cudaStream_t stream) {
    WGB_REQUIRE(a && w && c, "null pointer");
    WGB_REQUIRE(N > 0 && N % kBlockN == 0 && K > 0 && K % kBlockK == 0, "N must be a multiple of 256 and K of 64 (N=%d K=%d)", N, K);
    TcParams p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = N / kBlockN; p.ppi = 1; p.n_chunks = K / kBlockK;
    p.bias = bias; p.c_out = c; p.n_total = N;
    p.seg_chunks = p.n_chunks; p.seg_mask = 0;
    CUtensorMap ma0, mb;
    if (int e = act_map(&ma0, a, K, T, batch)) return e;
    if (int e = weight_map(&mb, w, N, K)) return e;
    return out_bf16 ? launch<MODE_PLAIN, 0, 1>(ma0, ma0, mb, ma0, p, stream) : launch<MODE_PLAIN, 0, 0>(ma0, ma0, mb, ma0, p, stream);
}

// conv1d (stride 1, "same" zero padding) as an implicit GEMM: out[b,t,n] = act(bias[n] + sum_tap sum_c W[n][tap*C+c]
// A[b, t + (tap - (taps-1)/2) * dilation, c]).  A bf16 channels-last [B,T,C] (C % 64 == 0), W bf16 [N][taps*C]
// (N % 256 == 0), out fp32 or bf16 [B,T,N].  Every tap is a K segment whose TMA box is shifted in time; rows outside
// the sequence are zero-filled by TMA.  Serves the Tacotron-2 Postnet (tacotron/modules.py:94-137: five k = 5 convs
// with BatchNorm folded into W / bias, tanh between them).
int tc_conv1d_taps(const void* a, const void* w, const float* bias, void* c, int out_bf16, int batch, int T, int N, int C,
                   int taps, int dilation, int act, cudaStream_t stream) {
    WGB_REQUIRE(a && w && c, "null pointer");
    WGB_REQUIRE(N > 0 && N % kBlockN == 0 && C > 0 && C % kBlockK == 0, "N must be a multiple of 256 and C of 64 (N=%d C=%d)", N, C);
    WGB_REQUIRE(taps >= 1 && taps % 2 == 1 && taps <= 31 && dilation >= 1, "taps must be odd (got %d), dilation >= 1", taps);
    WGB_REQUIRE(act >= 0 && act <= 2, "act must be 0 (none), 1 (tanh) or 2 (relu)");
    TcParams p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = N / kBlockN; p.ppi = 1; p.n_chunks = taps * C / kBlockK;
    p.bias = bias; p.c_out = c; p.n_total = N;
    p.seg_chunks = C / kBlockK; p.seg_mask = 0;
    p.seg_shift0 = -((taps - 1) / 2) * dilation; p.seg_dshift = dilation; p.act = act;
    CUtensorMap ma0, mb;
    if (int e = act_map(&ma0, a, C, T, batch)) return e;
    if (int e = weight_map(&mb, w, N, taps * C)) return e;
    return out_bf16 ? launch<MODE_PLAIN, 0, 1>(ma0, ma0, mb, ma0, p, stream) : launch<MODE_PLAIN, 0, 0>(ma0, ma0, mb, ma0, p, stream);
}

// General segmented GEMM behind the backward pass of the WN layers (training direction):
//   out[b,t,n] = act(bias[n] + sum_s sum_c W[n][s*C + c] A_s[b, t + shift0 + s*dshift, c]) + res[b,t,n]
// A_s = a1 when bit s of seg_mask is set, else a0 (both bf16 [B,T,C], C % 64 == 0); W bf16 [N][n_seg*C] (N % 256 == 0);
// out / res fp32 (out_bf16 = 0) or bf16 [B,T,N]; res may be null or alias out.  stacked = 1: a0 is [n_seg, B, T, C] and
// segment s reads plane s (K-concatenation of several layers' tensors without copying them).  Uses: data gradient of the dilated
// conv (taps as shifted segments + residual stream), of res_skip (segments [g_h | g_skip]), of cond (accumulating).
int tc_gemm_seg(const void* a0, const void* a1, int n_seg, int seg_mask, const void* w, const float* bias, const void* res,
                void* c, int out_bf16, int batch, int T, int N, int C, int shift0, int dshift, int act, int stacked,
                cudaStream_t stream) {
    WGB_REQUIRE(a0 && w && c, "null pointer");
    WGB_REQUIRE(n_seg >= 1 && n_seg <= 31, "n_seg must be in 1..31 (got %d)", n_seg);
    WGB_REQUIRE(seg_mask == 0 || a1 != nullptr, "seg_mask selects a1, which is null");
    WGB_REQUIRE(N > 0 && N % kBlockN == 0 && C > 0 && C % kBlockK == 0, "N must be a multiple of 256 and C of 64 (N=%d C=%d)", N, C);
    WGB_REQUIRE(act >= 0 && act <= 2, "act must be 0 (none), 1 (tanh) or 2 (relu)");
    TcParams p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = N / kBlockN; p.ppi = 1; p.n_chunks = n_seg * C / kBlockK;
    p.bias = bias; p.c_out = c; p.n_total = N; p.res = res;
    p.seg_chunks = C / kBlockK; p.seg_mask = seg_mask;
    p.seg_shift0 = shift0; p.seg_dshift = dshift; p.act = act;
    WGB_REQUIRE(!stacked || seg_mask == 0, "stacked segments read a0 only");
    p.seg_bstride = stacked ? batch : 0;                  // a0 is [n_seg, B, T, C]: segment s = plane s
    CUtensorMap ma0, ma1, mb;
    if (int e = act_map(&ma0, a0, C, T, stacked ? batch * n_seg : batch)) return e;
    if (int e = act_map(&ma1, a1 ? a1 : a0, C, T, batch)) return e;
    if (int e = weight_map(&mb, w, N, n_seg * C)) return e;
    return out_bf16 ? launch<MODE_PLAIN, 0, 1>(ma0, ma1, mb, ma0, p, stream) : launch<MODE_PLAIN, 0, 0>(ma0, ma1, mb, ma0, p, stream);
}

// Split-bf16 ("3x bf16") GEMM for fp32-grade accuracy on the tensor cores:
//   C = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T,  A = A_hi + A_lo, W = W_hi + W_lo (each part bf16),
// run as ONE K = 3*K GEMM: K segments [A_hi | A_lo | A_hi] against the packed weight [W_hi | W_hi | W_lo].
// Rows of A may overlap in memory (row_stride < K): that is how STFT frames (hop < filter_length) are read
// straight from the padded signal (reference stft.py:85-89 does the same with a strided conv).
static int gemm_split3(const void* a_hi, const void* a_lo, const void* w3, const float* bias, void* c, int batch, int rows,
                       int N, int K, long long row_stride, long long batch_stride, int epilogue, void* c2,
                       const float* spec_bias, float strength, int cutoff, cudaStream_t stream) {
    WGB_REQUIRE(a_hi && a_lo && w3 && c, "null pointer");
    WGB_REQUIRE(N > 0 && N % kBlockN == 0 && K > 0 && K % kBlockK == 0, "N must be a multiple of 256 and K of 64 (N=%d K=%d)", N, K);
    WGB_REQUIRE(row_stride % 8 == 0 && batch_stride % 8 == 0, "row/batch strides must be multiples of 8 elements (16 B)");
    TcParams p{};
    if (int e = fill_common(p, batch, rows)) return e;
    p.n_pass = N / kBlockN; p.ppi = 1; p.n_chunks = 3 * K / kBlockK;
    p.bias = bias; p.c_out = c; p.n_total = N;
    p.seg_chunks = K / kBlockK; p.seg_mask = 0b010;
    CUtensorMap mhi, mlo, mb;
    const uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows), static_cast<uint64_t>(batch)};
    const uint64_t strides[2] = {static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(batch_stride) * 2};
    const uint32_t box[3] = {kBlockK, kBlockM, 1};
    if (int e = make_tmap_bf16(&mhi, a_hi, 3, dims, strides, box)) return e;
    if (int e = make_tmap_bf16(&mlo, a_lo, 3, dims, strides, box)) return e;
    if (int e = weight_map(&mb, w3, N, 3 * K)) return e;
    p.c_out2 = c2; p.spec_bias = spec_bias; p.strength = strength; p.cutoff = cutoff;
    if (epilogue == 2) return launch<MODE_PLAIN, 0, 2>(mhi, mlo, mb, mhi, p, stream);
    if (epilogue == 3) return launch<MODE_PLAIN, 0, 3>(mhi, mlo, mb, mhi, p, stream);
    return launch<MODE_PLAIN, 0, 0>(mhi, mlo, mb, mhi, p, stream);
}

int tc_gemm_split3(const void* a_hi, const void* a_lo, const void* w3, const float* bias, void* c, int batch, int rows,
                   int N, int K, long long row_stride, long long batch_stride, cudaStream_t stream) {
    return gemm_split3(a_hi, a_lo, w3, bias, c, batch, rows, N, K, row_stride, batch_stride, 0, nullptr, nullptr, 0.f, 0,
                       stream);
}

// STFT with the forward basis packed in Re/Im-paired order (pass p = Re rows of bins 128p..128p+127, then their Im
// rows): |X| straight from the GEMM epilogue, channels-last [B, rows, cp] (stft.py:85-97; the mel path never needs
// Re / Im / phase).
int tc_stft_mag(const void* a_hi, const void* a_lo, const void* w3_paired, void* mag_cl, int batch, int rows, int cp, int K,
                long long row_stride, long long batch_stride, cudaStream_t stream) {
    WGB_REQUIRE(cp > 0 && cp % 128 == 0, "cp (%d) must be a multiple of 128", cp);
    return gemm_split3(a_hi, a_lo, w3_paired, nullptr, mag_cl, batch, rows, 2 * cp, K, row_stride, batch_stride, 2, nullptr,
                       nullptr, 0.f, 0, stream);
}

// TacotronSTFT.mel_spectrogram as ONE kernel: the paired-basis STFT GEMM with the magnitude, the mel filterbank (as a
// sparse per-bin table: every bin lies in at most two adjacent triangular filters) and log(clamp) in the epilogue.
// out fp32 [B, n_mel, rows] (the reference's layout); mel_table [cp] float4 {first filter, w_first, w_next, 0}.
int tc_stft_mel(const void* a_hi, const void* a_lo, const void* w3_paired, const void* mel_table, void* out, int batch,
                int rows, int cp, int K, long long row_stride, long long batch_stride, int n_mel, float clip,
                cudaStream_t stream) {
    WGB_REQUIRE(a_hi && a_lo && w3_paired && mel_table && out, "null pointer");
    WGB_REQUIRE(cp > 0 && cp % 128 == 0 && cp <= 1024, "cp (%d) must be a multiple of 128 and <= 1024", cp);
    WGB_REQUIRE(n_mel >= 1 && n_mel <= kMelMax, "n_mel (%d) must be in 1..%d", n_mel, kMelMax);
    WGB_REQUIRE(K > 0 && K % kBlockK == 0, "K must be a multiple of 64");
    WGB_REQUIRE(row_stride % 8 == 0 && batch_stride % 8 == 0, "row/batch strides must be multiples of 8 elements (16 B)");
    TcParams p{};
    if (int e = fill_common(p, batch, rows)) return e;
    const int N = 2 * cp;
    p.n_pass = N / kBlockN; p.ppi = p.n_pass; p.n_chunks = 3 * K / kBlockK;       // one CTA runs all passes of a tile
    p.c_out = out; p.n_total = N;
    p.seg_chunks = K / kBlockK; p.seg_mask = 0b010;
    p.mel_table = static_cast<const float4*>(mel_table); p.n_mel = n_mel; p.mel_clip = clip;
    CUtensorMap mhi, mlo, mb;
    const uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows), static_cast<uint64_t>(batch)};
    const uint64_t strides[2] = {static_cast<uint64_t>(row_stride) * 2, static_cast<uint64_t>(batch_stride) * 2};
    const uint32_t box[3] = {kBlockK, kBlockM, 1};
    if (int e = make_tmap_bf16(&mhi, a_hi, 3, dims, strides, box)) return e;
    if (int e = make_tmap_bf16(&mlo, a_lo, 3, dims, strides, box)) return e;
    if (int e = weight_map(&mb, w3_paired, N, 3 * K)) return e;
    return launch<MODE_STFT_MEL, 0, 0>(mhi, mlo, mb, mhi, p, stream);
}

// Same GEMM with the Denoiser's spectral subtraction in the epilogue (denoiser.py:36-38 + the cos/sin recombination
// of stft.py:102-103 as a magnitude ratio); writes the bf16 hi / lo operands [B*rows, 2cp] (Re | Im) of the
// inverse-basis GEMM directly.
int tc_stft_denoise(const void* a_hi, const void* a_lo, const void* w3_paired, const float* bias_spec, float strength,
                    void* hi_out, void* lo_out, int batch, int rows, int cutoff, int cp, int K, long long row_stride,
                    long long batch_stride, cudaStream_t stream) {
    WGB_REQUIRE(cp > 0 && cp % 128 == 0 && cutoff > 0 && cutoff <= cp, "cp (%d) must be a multiple of 128 and >= cutoff", cp);
    WGB_REQUIRE(bias_spec && lo_out, "null pointer");
    return gemm_split3(a_hi, a_lo, w3_paired, nullptr, hi_out, batch, rows, 2 * cp, K, row_stride, batch_stride, 3, lo_out,
                       bias_spec, strength, cutoff, stream);
}

int tc_wn_skip_end(const void* acts_all, int n_layers, const void* w_skip, const float* w_end, const float* b_end,
                   float* x, const float* w_mix, float* log_s, int batch, int T, int n_half, int direction,
                   cudaStream_t stream) {
    WGB_REQUIRE(acts_all && w_skip && w_end && b_end && x, "null pointer");
    WGB_REQUIRE(n_layers == 8, "tensor-core skip GEMM is specialised for 8 layers (got %d)", n_layers);
    WGB_REQUIRE(n_half >= 1 && n_half <= 4, "n_half must be in 1..4 (got %d)", n_half);
    WGB_REQUIRE(direction == 0 || direction == 1, "direction must be 0 (infer) or 1 (forward)");
    WGB_REQUIRE(direction == 1 ? log_s != nullptr : w_mix != nullptr, "missing log_s / w_mix for this direction");
    TcParams p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = 2; p.ppi = 2; p.n_chunks = n_layers * kNCh / kBlockK;
    p.w_end = w_end; p.b_end = b_end; p.x = x; p.w_mix = w_mix; p.log_s = log_s;
    CUtensorMap ma0, mb;
    if (int e = act_map(&ma0, acts_all, kNCh, T, batch * n_layers)) return e;
    if (int e = weight_map(&mb, w_skip, kNCh, n_layers * kNCh)) return e;
#define WGB_SKIP_CASE(NH)                                                                            \
    case NH:                                                                                         \
        return direction == 0 ? launch<MODE_SKIP_END, NH, 0>(ma0, ma0, mb, ma0, p, stream)           \
                              : launch<MODE_SKIP_END, NH, 1>(ma0, ma0, mb, ma0, p, stream);
    switch (n_half) {
        WGB_SKIP_CASE(1)
        WGB_SKIP_CASE(2)
        WGB_SKIP_CASE(3)
        WGB_SKIP_CASE(4)
    }
#undef WGB_SKIP_CASE
    return fail(WGB_ERR_ARGUMENT, "unreachable");
}

}  // namespace wgb
