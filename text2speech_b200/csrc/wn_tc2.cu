// WN-layer GEMMs on CTA pairs: tcgen05.mma.cta_group::2 (UMMA M = 256 over two SMs).
//
// Same math as the one-CTA kernels in wn_tc.cu, but the two CTAs of a cluster share every weight tile:
// each CTA TMA-loads its own 128 time rows of the activation operand and only HALF (128 of 256 rows) of
// the weight tile, and the leader CTA issues one M=256 x N=256 MMA that reads both halves.  Per CTA and
// K chunk that is 16 KB + 16 KB instead of 16 KB + 32 KB from L2, and the tensor core reads each weight
// byte from shared memory once per pair instead of once per CTA -- the point on a power-capped part.
//
//   GATE      in_layers[i] (k=3, dilation d) + cond_layers[i] + bias -> tanh*sigmoid -> acts   (glow.py:159-162)
//   GATE_MEL  the same layer with the conditioning path composed with the upsampler: cond_layers[i](upsample(mel))
//             is linear in the 4 x 80 mel values that reach a group step, with a weight that depends only on the
//             step's phase inside its frame (t mod 32), so rows are tiled PHASE-MAJOR (128 consecutive frames at
//             one phase, read through a 4-D tensor map of h) and the conditioning costs K = 320 instead of 640:
//             K = 1856 per layer instead of 2176, and the [B,T,640] cond tensor is never read (glow.py:252-258 + :161)
//   RES       res half of res_skip_layers[i]: h_out = h_in + W_res acts + b                   (glow.py:164-166)
//   SKIP_END  sum_i W_skip_i acts_i as ONE K = 8*512 GEMM, then end 1x1, affine coupling and the
//             invertible 1x1 conv in the fp32 epilogue                     (glow.py:167-175, :277-282 / :241-246)
//
// Protocol (per pair): both CTAs run a TMA producer; all operand loads complete_tx on the LEADER's full
// barrier, which the leader's producer arms with the byte count of both CTAs.  The leader's MMA thread
// commits with .multicast::cluster to the empty / tmem-full barriers of both CTAs; the epilogue warps of
// both CTAs arrive (the peer remotely) on the leader's tmem-empty barrier.  RES additionally stages its
// own 128 x 256 h tile per CTA in two 128-column halves (TMA load -> in-place update in the swizzled layout
// -> TMA store), driven by a seventh warp behind CTA-local barriers: while the epilogue works on one half,
// the other half's store drains and the next pass's h_in streams in, so the HBM latency of the residual
// stream never sits between two epilogues.  RES can also run a third, skinny pass per tile over the same activations:
// N = 16 against the bf16 hi/lo rows of W_end W_skip_i, added into a per-row fp32 accumulator (skip_acc [B*T][8]) --
// the layer's contribution to WN.end's output (glow.py:167-175), so the skip path never re-reads the activations.
#include "common.cuh"
#include "ptx.cuh"

namespace wgb {

namespace tc2 {

constexpr int kBlockM = 128;          // rows per CTA (UMMA M = 256 per pair)
constexpr int kBlockN = 256;
constexpr int kHalfN = 128;           // weight rows each CTA loads
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytes = kHalfN * kBlockK * 2;       // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;
constexpr int kThreads = 192;
constexpr int kThreadsRes = 224;      // + warp 6: h-tile loader / storer
constexpr int kNCh = 512;
constexpr int kNCond = 640;

enum Mode { GATE = 0, RES = 1, SKIP_END = 2, GATE_MEL = 3, GATE_MEL_ACC = 4 };   // _ACC: + skip accumulation in the epilogue
__host__ __device__ constexpr bool is_mel(int mode) { return mode == GATE_MEL || mode == GATE_MEL_ACC; }
constexpr int kPhases = 32;          // group steps per mel frame (hop 256 / n_group 8)
constexpr int kMelK = 320;           // 4 upsample taps x 80 mel channels

// Shared memory (offsets from a 1024 B aligned base): [TMA ring][mode extra][barriers]
//   GATE      6 stages + 4 KB bias (sigmoid half pre-halved)
//   GATE_MEL  6 stages + 4 KB bias + 16 KB W_end W_skip_i (fp32 [512][8]) for the skip accumulation in the epilogue
//   RES       5 stages + 64 KB h tile (the ring depth bounds this latency-bound kernel; the bias is read through L1)
//   SKIP_END  6 stages + 16 KB W_end^T + 10 KB WN.start weights of the next flow
template <int MODE>
struct Smem {
    static constexpr int kStages = MODE == RES ? 5 : 6;
    static constexpr int kExtraOff = kStages * kStageBytes;
    static constexpr int kExtraBytes =
        (MODE == GATE || MODE == GATE_MEL) ? 2 * kNCh * 4 : MODE == GATE_MEL_ACC ? 2 * kNCh * 4 + kNCh * 8 * 4
                                           : (MODE == RES ? kBlockM * kBlockN * 2 : kNCh * 8 * 4 + kNCh * 5 * 4);
    static constexpr int kBarOff = kExtraOff + kExtraBytes;
    static constexpr int kTotal = 1024 + kBarOff + 256;
};

struct Params {
    int batch, T, tiles_per_b, n_tiles;
    int n_pass, ppi, n_chunks, dilation;
    int frames, n_fblk;           // GATE_MEL: frames per tiled sequence (one utterance, or all of them in the padded
                                  // layout), 128-frame blocks per sequence
    int l2_hint;                  // wgb_set_tuning "gate_l2_hint" / "res_l2_hint": bit 0 weights evict_last (all modes);
                                  // GATE_MEL bit 1 h taps evict_first, bit 2 mel_stack evict_last, bit 3 acts stores
                                  // evict_first; RES bit 1 activations evict_first
    int n_tap_chunks;             // GATE_MEL: K chunks of the in_layers part: 24 (three dilated taps of h), or 1 when the
                                  // first layer reads the pre-stacked flow state instead (x_stack, see tc2_wn_gate_mel0)
    int f_pad, f_real;            // GATE_MEL padded layout: frame pitch per utterance (> f_real: guard frames of zeros
                                  // separate the utterances) and real frames per utterance; f_pad = 0: per-utterance tiles
    // GATE_MEL, optional: skip path accumulated in the epilogue.  w_comp fp32 [512][8] = (W_end W_skip_i)^T of THIS
    // layer; skip_acc fp32 [4 passes][rows_total][8]: slot (pass, row) += sum over the pass's 128 channels of
    // acts * w_comp (stored instead of added when skip_first).  One thread owns a slot: deterministic.
    // RES, optional (skip_acc non-NULL): third pass, N = 16 rows of map_x (hi / lo of W_end W_skip_i), result added
    // (stored when skip_first) into skip_acc fp32 [B*T][8]
    const float* w_comp;
    float* skip_acc;
    long long rows_total;
    int skip_first;
    const float* bias;            // GATE [1024] packed order, RES [512]
    __nv_bfloat16* acts_out;      // GATE [B,T,512]
    int seg_mask;                 // RES: bit s set -> K segment s reads the second operand tensor (map_x) instead of map_a0
    int seg_chunks, seg_shift0, seg_dshift;   // RES: K chunk kc reads A columns (kc % seg_chunks) * 64 of rows
                                  //      t + seg_shift0 + (kc / seg_chunks) * seg_dshift (conv taps; plain GEMM: n_chunks, 0, 0)
    __nv_bfloat16* ts_out;        // GATE, training forward (optional): [B,T,1024] = tanh half | sigmoid half, original
                                  //       channel order (what the gate's backward needs; glow.py:33-40 under autograd)
    const float* w_end;           // SKIP_END [512][8] fp32 (rows >= 2*n_half zero)
    const float* b_end;           // SKIP_END [8] (skip biases folded in)
    float* x;                     // SKIP_END flow state [B,T,8]
    const float* w_mix;           // SKIP_END infer: W^-1 [8][8]
    float* log_s;                 // SKIP_END forward: [B,n_half,T]
    // SKIP_END infer, optional: WN.start of the NEXT flow to run (glow.py:156 of flow k-1) fused behind the
    // coupling: h_next[b,t,:] = W[512][n_half_next] x_new[a0 channels of that flow] + b
    const float* next_w_start;
    const float* next_b_start;
    __nv_bfloat16* h_next;
    int next_n_half;
    long long h_next_batch_rows;  // row pitch per utterance of h_next (T, or more in the padded layout)
};

template <int MODE, int NHALF, int DIR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsRes, 1)
pair_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
            const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_c,
            const __grid_constant__ CUtensorMap map_x, const Params p) {
    using SL = Smem<MODE>;
    constexpr int kStages = SL::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::kBarOff);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* hfull_bar = tempty_bar + 2;        // RES [2]: h_in half-tile landed in smem
    uint64_t* hready_bar = hfull_bar + 2;        // RES [2]: half-tile updated in place by the epilogue, ready to store
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(hready_bar + 2);
    uint8_t* s_extra = smem + SL::kExtraOff;
    float* s_f32 = reinterpret_cast<float*>(s_extra);                                    // GATE bias / SKIP_END W_end^T

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    pdl_launch_dependents();          // persistent grid: the next kernel may queue up behind this one right away

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_w);
        if constexpr (MODE == RES || is_mel(MODE)) tma_prefetch_desc(&map_c);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&hfull_bar[i], 1);
            mbar_init(&hready_bar[i], 4);      // the four epilogue warps
        }
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
    if (warp >= 2 && warp < 6) {
        const int i0 = threadIdx.x - 64;
        if constexpr (MODE == GATE || is_mel(MODE)) {
            // packed column c of pass p: tanh row for (c & 255) < 128, else the matching sigmoid row, whose
            // pre-activation is halved (sigmoid(b) = 0.5 tanh(b/2) + 0.5)
            for (int i = i0; i < 2 * kNCh; i += 128) s_f32[i] = p.bias[i] * ((i & 128) ? 0.5f : 1.f);
            if constexpr (MODE == GATE_MEL_ACC) {
                for (int i = i0; i < kNCh * 8; i += 128) s_f32[2 * kNCh + i] = p.w_comp[i];
            }
        } else if constexpr (MODE == RES) {
            // bias stays in global memory (__ldg): the 2 KB would cost the fifth ring stage
        } else {
            for (int i = i0; i < kNCh * 8; i += 128) s_f32[i] = p.w_end[i];
            if (p.h_next) {                       // [512][4] zero-padded weight rows, then [512] bias
                float* s_ws = s_f32 + kNCh * 8;
                for (int i = i0; i < kNCh * 4; i += 128)
                    s_ws[i] = (i & 3) < p.next_n_half ? p.next_w_start[(i >> 2) * p.next_n_half + (i & 3)] : 0.f;
                for (int i = i0; i < kNCh; i += 128) s_ws[kNCh * 4 + i] = p.next_b_start[i];
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                       // everything above touched only weights / biases; activations start here

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
                const bool valid = tile < p.n_tiles;
                const int b = valid ? tile / p.tiles_per_b : 0;
                // first row (GATE_MEL: first frame) of this CTA's tile; an absent second tile reads all-OOB rows = zeros
                const int t0 = valid ? (tile % p.tiles_per_b) * kBlockM : (is_mel(MODE) ? p.frames : p.T) + 4 * kBlockM;
                for (int pp = 0; pp < p.ppi; ++pp) {
                    const int vpass = (item % groups) * p.ppi + pp;
                    // GATE_MEL enumerates (phase, pass) as 128 virtual passes: pass fastest, then the phase
                    const int pass = is_mel(MODE) ? (vpass & 3) : vpass;
                    const int phase = vpass >> 2;
                    for (int kc = 0; kc < p.n_chunks; ++kc) {
                        mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
                        uint8_t* sa = smem + s * kStageBytes;
                        uint8_t* sb = sa + kABytes;
                        const bool skinny = MODE == RES && pass == 2;          // N = 16: 8 weight rows (1 KB) per CTA
                        if (leader) mbar_arrive_expect_tx(&full_bar[s], skinny ? 2 * (kABytes + 1024) : 2 * kStageBytes);
                        const uint32_t bar = mapa_u32(&full_bar[s], 0);
                        if constexpr (MODE == GATE) {
                            if (kc < 24) {
                                const int tap = kc >> 3;
                                tma_load_3d_2sm(sa, &map_a0, bar, (kc & 7) * kBlockK, t0 + (tap - 1) * p.dilation, b);
                            } else {
                                tma_load_3d_2sm(sa, &map_a1, bar, (kc - 24) * kBlockK, t0, b);
                            }
                        } else if constexpr (is_mel(MODE)) {
                            if (kc < p.n_tap_chunks) {
                                // row (f, phase) of the tile needs h at group step 32 f + phase + (tap-1) d = frame
                                // f + (q >> 5), phase q & 31 with q = phase + (tap-1) d (floor / mod, q may be < 0);
                                // frames outside [0, F) are zero-filled by TMA = the conv's zero padding.  With one tap
                                // chunk (pre-stacked first layer) q = phase: the row itself.
                                const int q = phase + (p.n_tap_chunks == 1 ? 0 : ((kc >> 3) - 1) * p.dilation);
                                if (p.l2_hint & 2)
                                    tma_load_4d_2sm_hint(sa, &map_a0, bar, (kc & 7) * kBlockK, q & (kPhases - 1), t0 + (q >> 5), b,
                                                         l2_policy_evict_first());
                                else
                                    tma_load_4d_2sm(sa, &map_a0, bar, (kc & 7) * kBlockK, q & (kPhases - 1), t0 + (q >> 5), b);
                            } else {
                                if (p.l2_hint & 4)
                                    tma_load_3d_2sm_hint(sa, &map_a1, bar, (kc - p.n_tap_chunks) * kBlockK, t0, b, l2_policy_evict_last());
                                else
                                    tma_load_3d_2sm(sa, &map_a1, bar, (kc - p.n_tap_chunks) * kBlockK, t0, b);
                            }
                        } else if constexpr (MODE == RES) {
                            const int seg = kc / p.seg_chunks;
                            if (p.l2_hint & 2)                             // activations are read once here: evict_first
                                tma_load_3d_2sm_hint(sa, ((p.seg_mask >> seg) & 1) ? &map_x : &map_a0, bar,
                                                     (kc - seg * p.seg_chunks) * kBlockK, t0 + p.seg_shift0 + seg * p.seg_dshift, b,
                                                     l2_policy_evict_first());
                            else
                                tma_load_3d_2sm(sa, ((p.seg_mask >> seg) & 1) ? &map_x : &map_a0, bar,
                                                (kc - seg * p.seg_chunks) * kBlockK, t0 + p.seg_shift0 + seg * p.seg_dshift, b);
                        } else {
                            tma_load_3d_2sm(sa, &map_a0, bar, (kc & 7) * kBlockK, t0, (kc >> 3) * p.batch + b);
                        }
                        const int w_row = pass * kBlockN + static_cast<int>(rank) * kHalfN;
                        // weight tiles are re-read by every CTA pair for the whole launch while activations stream through
                        // L2 once: bit 0 of l2_hint gives them evict_last priority (measured: profiles/r02k_l2_hint_ab.json)
                        const bool w_last = (p.l2_hint & 1) != 0;
                        if (is_mel(MODE) && kc >= p.n_tap_chunks) {     // phase-specific composed conditioning weight [32*1024][320]
                            if (w_last)
                                tma_load_2d_2sm_hint(sb, &map_c, bar, (kc - p.n_tap_chunks) * kBlockK, phase * (2 * kNCh) + w_row,
                                                     l2_policy_evict_last());
                            else
                                tma_load_2d_2sm(sb, &map_c, bar, (kc - p.n_tap_chunks) * kBlockK, phase * (2 * kNCh) + w_row);
                        } else if (skinny) {
                            tma_load_2d_2sm(sb, &map_x, bar, kc * kBlockK, static_cast<int>(rank) * 8);
                        } else if (w_last) {
                            tma_load_2d_2sm_hint(sb, &map_w, bar, kc * kBlockK, w_row, l2_policy_evict_last());
                        } else {
                            tma_load_2d_2sm(sb, &map_w, bar, kc * kBlockK, w_row);
                        }
                        if (++s == kStages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader && lane == 0) {
            constexpr uint32_t idesc256 = umma_idesc_bf16_f32(2 * kBlockM, kBlockN);
            constexpr uint32_t idesc16 = umma_idesc_bf16_f32(2 * kBlockM, 16);
            int s = 0;
            uint32_t ph = 0, acc_it = 0;
            for (int item = pair_id; item < n_items; item += n_pairs) {
                const uint32_t idesc = (MODE == RES && item % groups == 2) ? idesc16 : idesc256;
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
    } else if (warp == 6) {
        // ------------------------------------------------------------------ RES: h-tile loader / storer (both CTAs)
        if constexpr (MODE == RES) {
            if (lane == 0) {
                // step k of this CTA = (item pair_id + k*n_pairs, its single pass); half hf = columns hf*128..+127 of
                // the pass = boxes 2hf, 2hf+1 of the 64 KB tile buffer
                auto coords = [&](int item, int& pass, int& t0, int& b) {
                    const int tile = 2 * (item / groups) + static_cast<int>(rank);
                    const bool valid = tile < p.n_tiles;
                    pass = item % groups;
                    b = valid ? tile / p.tiles_per_b : 0;
                    t0 = valid ? (tile % p.tiles_per_b) * kBlockM : p.T + 4 * kBlockM;      // OOB: zeros in, nothing out
                };
                auto load_half = [&](int hf, int pass, int t0, int b) {
                    mbar_arrive_expect_tx(&hfull_bar[hf], kBlockM * kHalfN * 2);
#pragma unroll
                    for (int j = 2 * hf; j < 2 * hf + 2; ++j)
                        tma_load_3d(s_extra + j * kABytes, &map_a1, &hfull_bar[hf], pass * kBlockN + j * kBlockK, t0, b);
                };
                // items whose pass is 2 (the skinny skip pass) have no h tile: step over them
                auto next_h_item = [&](int item) {
                    do item += n_pairs; while (item < n_items && item % groups == 2);
                    return item;
                };
                int pass, t0, b;
                const int first = (pair_id < n_items && pair_id % groups == 2) ? next_h_item(pair_id) : pair_id;
                if (first < n_items) {
                    coords(first, pass, t0, b);
                    load_half(0, pass, t0, b);
                    load_half(1, pass, t0, b);
                }
                uint32_t it = 0;
                for (int item = first; item < n_items; ++it) {
                    coords(item, pass, t0, b);
                    const int nitem = next_h_item(item);
                    const bool has_next = nitem < n_items;
                    int npass = 0, nt0 = 0, nb = 0;
                    if (has_next) {
                        coords(nitem, npass, nt0, nb);
#pragma unroll
                        for (int j = 0; j < kBlockN / kBlockK; ++j)      // pull the next h_in tile into L2 a whole pass early
                            tma_prefetch_3d(&map_a1, npass * kBlockN + j * kBlockK, nt0, nb);
                    }
                    for (int hf = 0; hf < 2; ++hf) {
                        mbar_wait(&hready_bar[hf], it & 1, 700 + hf);
#pragma unroll
                        for (int j = 2 * hf; j < 2 * hf + 2; ++j)
                            tma_store_3d(&map_c, s_extra + j * kABytes, pass * kBlockN + j * kBlockK, t0, b);
                        tma_store_commit();
                        tma_store_wait_read0();                          // smem half free again
                        if (has_next) load_half(hf, npass, nt0, nb);
                    }
                    item = nitem;
                }
                tma_store_wait_all0();                                   // h_out globally visible before the kernel ends
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0, hph = 0;
        for (int item = pair_id; item < n_items; item += n_pairs) {
            const int tile = 2 * (item / groups) + static_cast<int>(rank);
            const bool valid = tile < p.n_tiles;
            const int b = valid ? tile / p.tiles_per_b : 0;
            int t = valid ? (tile % p.tiles_per_b) * kBlockM + row : p.T;
            bool live = t < p.T;
            int bb = b;
            if constexpr (is_mel(MODE)) {                     // phase-major tile: row = frame, t = 32 frame + phase
                live = valid && t < p.frames;
                if (p.f_pad) {                                // padded layout: one frame axis over all utterances
                    bb = t / p.f_pad;
                    t -= bb * p.f_pad;
                    live = live && t < p.f_real;
                }
                t = t * kPhases + ((item % groups) >> 2);
                live = live && t < p.T;                       // partial last frame
            }
            const size_t grow = static_cast<size_t>(bb) * p.T + t;
            float outv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) outv[j] = 0.f;

            for (int pp = 0; pp < p.ppi; ++pp, ++acc_it) {
                const int vpass = (item % groups) * p.ppi + pp;
                const int pass = is_mel(MODE) ? (vpass & 3) : vpass;
                const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
                // GATE_MEL skip accumulation: fetch this thread's slot before waiting for the accumulator
                float sk[8];
                float4 prev0 = make_float4(0.f, 0.f, 0.f, 0.f), prev1 = prev0;
                float* slot = nullptr;
                bool do_skip = false;
                if constexpr (MODE == GATE_MEL_ACC) {
                    do_skip = true;
                    {
#pragma unroll
                        for (int m = 0; m < 8; ++m) sk[m] = 0.f;
                        slot = p.skip_acc + (static_cast<size_t>(pass) * p.rows_total + grow) * 8;
                        if (live && !p.skip_first) {
                            prev0 = *reinterpret_cast<const float4*>(slot);
                            prev1 = *reinterpret_cast<const float4*>(slot + 4);
                        }
                    }
                }
                mbar_wait(&tfull_bar[as], aph, 400 + as);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kBlockN;

                if constexpr (MODE == GATE || is_mel(MODE)) {
                    const float4* bt4 = reinterpret_cast<const float4*>(s_f32 + pass * kBlockN);
                    const float4* bs4 = bt4 + kHalfN / 4;
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
                            const float4 bt = bt4[ch * 8 + j], bs = bs4[ch * 8 + j];      // warp-uniform: broadcast LDS.128
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
                            if constexpr (MODE == GATE_MEL_ACC) {
                                {                                      // 4 channels x [8] composed skip/end weights
                                    const float4* wc = reinterpret_cast<const float4*>(s_f32 + 2 * kNCh) +
                                                       (pass * 128 + ch * 32 + 4 * j) * 2;
                                    const float gv[4] = {g0, g1, g2, g3};
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float4 w0 = wc[2 * e], w1 = wc[2 * e + 1];
                                        sk[0] = fmaf(gv[e], w0.x, sk[0]);
                                        sk[1] = fmaf(gv[e], w0.y, sk[1]);
                                        sk[2] = fmaf(gv[e], w0.z, sk[2]);
                                        sk[3] = fmaf(gv[e], w0.w, sk[3]);
                                        sk[4] = fmaf(gv[e], w1.x, sk[4]);
                                        sk[5] = fmaf(gv[e], w1.y, sk[5]);
                                        sk[6] = fmaf(gv[e], w1.z, sk[6]);
                                        sk[7] = fmaf(gv[e], w1.w, sk[7]);
                                    }
                                }
                            }
                        }
                        if (live) {
                            uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
                            if (is_mel(MODE) && (p.l2_hint & 8)) {           // experiment: streaming stores (evict_first)
                                const uint64_t pol = l2_policy_evict_first();
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    st_global_v4_hint(d4 + j, make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2],
                                                                         packed[4 * j + 3]), pol);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                            }
                        }
                        if constexpr (MODE == GATE) {
                            if (p.ts_out != nullptr && live) {     // training forward: keep tanh and sigmoid for the backward
                                __nv_bfloat16* tdst = p.ts_out + grow * (2 * kNCh) + pass * 128 + ch * 32;
#pragma unroll
                                for (int half = 0; half < 2; ++half) {
                                    uint32_t pk[16];
#pragma unroll
                                    for (int j = 0; j < 16; ++j) {
                                        float v0, v1;
                                        if (half == 0) {
                                            const float4 bt = bt4[ch * 8 + (j >> 1)];
                                            const float b0 = (j & 1) ? bt.z : bt.x, b1 = (j & 1) ? bt.w : bt.y;
                                            v0 = tanh_approx(__uint_as_float(vt[2 * j]) + b0);
                                            v1 = tanh_approx(__uint_as_float(vt[2 * j + 1]) + b1);
                                        } else {
                                            const float4 bs = bs4[ch * 8 + (j >> 1)];
                                            const float b0 = (j & 1) ? bs.z : bs.x, b1 = (j & 1) ? bs.w : bs.y;
                                            v0 = fmaf(tanh_approx(fmaf(__uint_as_float(vs[2 * j]), 0.5f, b0)), 0.5f, 0.5f);
                                            v1 = fmaf(tanh_approx(fmaf(__uint_as_float(vs[2 * j + 1]), 0.5f, b1)), 0.5f, 0.5f);
                                        }
                                        __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                                        pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                                    }
                                    uint4* t4 = reinterpret_cast<uint4*>(tdst + half * kNCh);
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        t4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                                }
                            }
                        }
                    }
                    if constexpr (MODE == GATE_MEL_ACC) {
                        if (live) {
                            *reinterpret_cast<float4*>(slot) =
                                make_float4(prev0.x + sk[0], prev0.y + sk[1], prev0.z + sk[2], prev0.w + sk[3]);
                            *reinterpret_cast<float4*>(slot + 4) =
                                make_float4(prev1.x + sk[4], prev1.y + sk[5], prev1.z + sk[6], prev1.w + sk[7]);
                        }
                    }
                } else if (MODE == RES && pass == 2) {
                    // skinny skip pass: columns 0..7 = hi rows, 8..15 = lo rows of W_end W_skip_i applied to this row
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(taddr, v);
                    tmem_ld_wait();
                    if (live) {
                        float* slot = p.skip_acc + grow * 8;
                        float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
                        if (!p.skip_first) {
                            o0 = *reinterpret_cast<const float4*>(slot);
                            o1 = *reinterpret_cast<const float4*>(slot + 4);
                        }
                        o0.x += __uint_as_float(v[0]) + __uint_as_float(v[8]);
                        o0.y += __uint_as_float(v[1]) + __uint_as_float(v[9]);
                        o0.z += __uint_as_float(v[2]) + __uint_as_float(v[10]);
                        o0.w += __uint_as_float(v[3]) + __uint_as_float(v[11]);
                        o1.x += __uint_as_float(v[4]) + __uint_as_float(v[12]);
                        o1.y += __uint_as_float(v[5]) + __uint_as_float(v[13]);
                        o1.z += __uint_as_float(v[6]) + __uint_as_float(v[14]);
                        o1.w += __uint_as_float(v[7]) + __uint_as_float(v[15]);
                        *reinterpret_cast<float4*>(slot) = o0;
                        *reinterpret_cast<float4*>(slot + 4) = o1;
                    }
                } else if constexpr (MODE == RES) {
                    // h tile sits in smem in the TMA SWIZZLE_128B layout (four [128 x 64] boxes): row r of box j is
                    // at j*16K + r*128, its 16 B chunk c at ((c ^ (r & 7)) << 4).  Each thread updates its own row
                    // in place; the whole tile then leaves through one TMA store (coalesced, asynchronous).
                    const float4* bias4 = reinterpret_cast<const float4*>(p.bias) + pass * (kBlockN / 4);
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        if ((ch & 3) == 0) mbar_wait(&hfull_bar[ch >> 2], hph, 600 + (ch >> 2));
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
                            const float4 b0 = __ldg(bias4 + ch * 8 + i * 2), b1 = __ldg(bias4 + ch * 8 + i * 2 + 1);
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
                        if ((ch & 3) == 3) {                            // half-tile done: hand it to the storer warp
                            fence_proxy_async_smem();                   // st.shared -> visible to the TMA store
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&hready_bar[ch >> 2]);
                        }
                    }
                    hph ^= 1;
                } else {
#pragma unroll 1
                    for (int ch = 0; ch < 8; ++ch) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        const float4* w4 = reinterpret_cast<const float4*>(s_f32 + (pass * kBlockN + ch * 32) * 8);
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
                // accumulator stage drained -> hand it back to the leader's MMA thread
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[as], 0));
            }

            if constexpr (MODE == SKIP_END) if (live) {
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
                if constexpr (DIR == 0) if (p.h_next) {
                    // WN.start of the next flow on the freshly updated row: its a0 = the first next_n_half of its
                    // last 2*next_n_half channels.  Overlaps the MMAs of the next tile (both TMEM stages are free).
                    const int nb = 8 - 2 * p.next_n_half;
                    float a0[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v = (i == nb + j) ? xv[i] : v;
                        a0[j] = v;                                  // weight columns >= next_n_half are zero
                    }
                    const float4* ws4 = reinterpret_cast<const float4*>(s_f32 + kNCh * 8);
                    const float4* bs4 = ws4 + kNCh;
                    uint4* dst = reinterpret_cast<uint4*>(p.h_next + (static_cast<size_t>(b) * p.h_next_batch_rows + t) * kNCh);
#pragma unroll 2
                    for (int c8 = 0; c8 < kNCh / 8; ++c8) {
                        const float4 b0 = bs4[2 * c8], b1 = bs4[2 * c8 + 1];
                        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                        uint32_t pk[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 w0 = ws4[c8 * 8 + 2 * j], w1 = ws4[c8 * 8 + 2 * j + 1];
                            const float v0 = fmaf(w0.w, a0[3], fmaf(w0.z, a0[2], fmaf(w0.y, a0[1], fmaf(w0.x, a0[0], bb[2 * j]))));
                            const float v1 = fmaf(w1.w, a0[3], fmaf(w1.z, a0[2], fmaf(w1.y, a0[1], fmaf(w1.x, a0[0], bb[2 * j + 1]))));
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
                        }
                        dst[c8] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
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

static int act_map(CUtensorMap* m, const void* base, int channels, int T, int batches) {
    const uint64_t dims[3] = {static_cast<uint64_t>(channels), static_cast<uint64_t>(T), static_cast<uint64_t>(batches)};
    const uint64_t strides[2] = {static_cast<uint64_t>(channels) * 2, static_cast<uint64_t>(channels) * 2 * T};
    const uint32_t box[3] = {kBlockK, kBlockM, 1};
    return make_tmap_bf16(m, base, 3, dims, strides, box);
}
static int weight_half_map(CUtensorMap* m, const void* base, int rows, int k) {
    const uint64_t dims[2] = {static_cast<uint64_t>(k), static_cast<uint64_t>(rows)};
    const uint64_t strides[1] = {static_cast<uint64_t>(k) * 2};
    const uint32_t box[2] = {kBlockK, kHalfN};
    return make_tmap_bf16(m, base, 2, dims, strides, box);
}

static int fill_common(Params& p, int batch, int T) {
    WGB_REQUIRE(batch > 0 && T > 0, "batch (%d) and T (%d) must be positive", batch, T);
    p.batch = batch;
    p.T = T;
    p.tiles_per_b = ceil_div(T, kBlockM);
    p.n_tiles = batch * p.tiles_per_b;
    return WGB_OK;
}

template <int MODE, int NHALF, int DIR>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w, const CUtensorMap& c, const Params& p,
                  cudaStream_t stream, const CUtensorMap* x = nullptr) {
    constexpr int smem = Smem<MODE>::kTotal;
    static_assert(smem <= 232448, "dynamic shared memory over the 227 KB per-CTA limit");
    auto kern = pair_kernel<MODE, NHALF, DIR>;
    WGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int n_items = ((p.n_tiles + 1) / 2) * (p.n_pass / p.ppi);
    int pairs = sm_count() / 2;
    if (n_items < pairs) pairs = n_items;
    WGB_CUDA_TRY(launch_pdl(kern, 2 * pairs, MODE == RES ? kThreadsRes : kThreads, smem, stream, a0, a1, w, c, x ? *x : a0, p));
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace tc2

// ts (optional, training forward): bf16 [B,T,1024] receiving tanh | sigmoid of the pre-activations (original order)
int tc2_wn_gate(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts, void* ts, int batch,
                int T, int dilation, cudaStream_t stream) {
    using namespace tc2;
    WGB_REQUIRE(h && cond && w_packed && bias && acts, "null pointer");
    WGB_REQUIRE(dilation >= 1, "dilation must be >= 1");
    Params p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = 4; p.ppi = 1; p.n_chunks = (3 * kNCh + kNCond) / kBlockK; p.dilation = dilation;
    p.l2_hint = tuning_get("gate_l2_hint") & 1;
    p.bias = bias; p.acts_out = static_cast<__nv_bfloat16*>(acts); p.ts_out = static_cast<__nv_bfloat16*>(ts);
    CUtensorMap mh, mc, mw;
    if (int e = act_map(&mh, h, kNCh, T, batch)) return e;
    if (int e = act_map(&mc, cond, kNCond, T, batch)) return e;
    if (int e = weight_half_map(&mw, w_packed, 2 * kNCh, 3 * kNCh + kNCond)) return e;
    return launch<GATE, 0, 0>(mh, mc, mw, mh, p, stream);
}

// Gate layer with the conditioning composed with the upsampler (see GATE_MEL above).  h bf16 [B,T,512] with
// T = 32 * frames; mel_stack bf16 [B, frames, 320] = the four mel frames feeding each frame's group steps (the
// upsample im2col); w_packed bf16 [1024][2176] (only the 1536 in_layers columns are read); w_mel bf16
// [32][1024][320] = per-phase W_cond U_phase in the packed row order; bias fp32 [1024] = b_in + b_cond + W_cond b_up.
static int gate_mel_launch(const void* a_taps, int tap_channels, int n_tap_chunks, const void* mel_stack, const void* w_taps,
                           int w_taps_k, const void* w_mel, const float* bias, void* acts, int batch, int T, int frames_pad,
                           int dilation, const float* w_comp, float* skip_acc, int skip_first, cudaStream_t stream) {
    using namespace tc2;
    WGB_REQUIRE(a_taps && mel_stack && w_taps && w_mel && bias && acts, "null pointer");
    WGB_REQUIRE(dilation >= 1, "dilation must be >= 1");
    WGB_REQUIRE(batch > 0 && T > 0, "batch and T must be positive");
    const int frames = ceil_div(T, kPhases);        // the last frame may be partial (forward() on audio that is not a
                                                    // multiple of 256 samples): its missing rows are guard rows
    WGB_REQUIRE((frames_pad == frames && T % kPhases == 0) || frames_pad >= frames + ceil_div(dilation, kPhases),
                "frames_pad (%d) must equal frames (%d, T a multiple of 32) or leave >= dilation/32 guard frames",
                frames_pad, frames);
    Params p{};
    p.T = T;
    const bool padded = frames_pad > frames;
    // padded layout: h [B, 32*frames_pad, 512] and mel_stack [B, frames_pad, 320] with the guard frames of h zero, tiled
    // as ONE sequence of B*frames_pad frames (the guards are the conv's zero padding between utterances)
    p.batch = padded ? 1 : batch;
    p.frames = padded ? batch * frames_pad : frames;
    p.f_pad = padded ? frames_pad : 0;
    p.f_real = frames;
    p.n_fblk = ceil_div(p.frames, kBlockM);
    p.tiles_per_b = p.n_fblk;                       // "tiles" = 128-frame blocks; each is visited once per phase and pass
    p.n_tiles = p.batch * p.n_fblk;
    p.n_tap_chunks = n_tap_chunks;
    p.l2_hint = tuning_get("gate_l2_hint");
    p.n_pass = 4 * kPhases; p.ppi = 1; p.n_chunks = n_tap_chunks + kMelK / kBlockK; p.dilation = dilation;
    p.bias = bias; p.acts_out = static_cast<__nv_bfloat16*>(acts);
    WGB_REQUIRE((w_comp == nullptr) == (skip_acc == nullptr), "w_comp and skip_acc go together");
    p.w_comp = w_comp; p.skip_acc = skip_acc; p.skip_first = skip_first;
    p.rows_total = static_cast<long long>(batch) * T;
    CUtensorMap mh, mm, mw, mv;
    {
        const uint64_t row_bytes = static_cast<uint64_t>(tap_channels) * 2;
        const uint64_t seq_rows = static_cast<uint64_t>(kPhases) * (padded ? frames_pad : frames);
        const uint64_t dims[4] = {static_cast<uint64_t>(tap_channels), kPhases, static_cast<uint64_t>(p.frames),
                                  static_cast<uint64_t>(p.batch)};
        const uint64_t strides[3] = {row_bytes, row_bytes * kPhases, row_bytes * seq_rows};
        const uint32_t box[4] = {kBlockK, 1, kBlockM, 1};
        if (int e = make_tmap_bf16(&mh, a_taps, 4, dims, strides, box)) return e;
    }
    if (int e = act_map(&mm, mel_stack, kMelK, p.frames, p.batch)) return e;
    if (int e = weight_half_map(&mw, w_taps, 2 * kNCh, w_taps_k)) return e;
    if (int e = weight_half_map(&mv, w_mel, kPhases * 2 * kNCh, kMelK)) return e;
    return skip_acc ? launch<GATE_MEL_ACC, 0, 0>(mh, mm, mw, mv, p, stream) : launch<GATE_MEL, 0, 0>(mh, mm, mw, mv, p, stream);
}

int tc2_wn_gate_mel(const void* h, const void* mel_stack, const void* w_packed, const void* w_mel, const float* bias,
                    void* acts, int batch, int T, int frames_pad, int dilation, const float* w_comp, float* skip_acc,
                    int skip_first, cudaStream_t stream) {
    using namespace tc2;
    return gate_mel_launch(h, kNCh, 3 * kNCh / kBlockK, mel_stack, w_packed, 3 * kNCh + kNCond, w_mel, bias, acts, batch, T,
                           frames_pad, dilation, w_comp, skip_acc, skip_first, stream);
}

// First WN layer with WN.start folded into its in_layers conv (both linear, glow.py:156 + :160, dilation 1):
// in_layers[0](start(a0))[t] = sum_tap (W_in0,tap W_start) a0[t+tap-1] + sum_tap (W_in0,tap b_start) [t+tap-1 in range].
// x_stack bf16 [B, 32*frames_pad, 64] holds, per group step, the hi/lo split of the three a0 taps and the in-range
// indicators (wgb_x_stack); w0 bf16 [1024][64] the matching composed weights (packing.py:pack_gate0).  The layer then
// costs K = 64 + 320 instead of 1536 + 320.
int tc2_wn_gate_mel0(const void* x_stack, const void* mel_stack, const void* w0, const void* w_mel, const float* bias,
                     void* acts, int batch, int T, int frames_pad, const float* w_comp, float* skip_acc, int skip_first,
                     cudaStream_t stream) {
    using namespace tc2;
    return gate_mel_launch(x_stack, kBlockK, 1, mel_stack, w0, kBlockK, w_mel, bias, acts, batch, T, frames_pad, 1, w_comp,
                           skip_acc, skip_first, stream);
}

static int h_map(CUtensorMap* m, const void* base, int T, int batch, long long batch_rows) {
    const uint64_t dims[3] = {tc2::kNCh, static_cast<uint64_t>(T), static_cast<uint64_t>(batch)};
    const uint64_t strides[2] = {tc2::kNCh * 2, static_cast<uint64_t>(tc2::kNCh) * 2 * static_cast<uint64_t>(batch_rows)};
    const uint32_t box[3] = {tc2::kBlockK, tc2::kBlockM, 1};
    return make_tmap_bf16(m, base, 3, dims, strides, box);
}

// h_batch_rows: row pitch per utterance of h_in / h_out (T for dense buffers; more in the padded layout, whose guard
// rows are never loaded nor stored because the tensor maps end at row T)
int tc2_wn_res(const void* acts, const void* w_res, const float* bias, const void* h_in, void* h_out, int batch, int T,
               long long h_batch_rows, const void* w16_layer, float* skip_acc, int skip_first, cudaStream_t stream) {
    using namespace tc2;
    WGB_REQUIRE(acts && w_res && bias && h_in && h_out, "null pointer");
    WGB_REQUIRE(h_batch_rows >= T, "h_batch_rows (%lld) must be >= T (%d)", h_batch_rows, T);
    WGB_REQUIRE((w16_layer == nullptr) == (skip_acc == nullptr), "w16_layer and skip_acc go together");
    Params p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = skip_acc ? 3 : 2; p.ppi = 1; p.n_chunks = kNCh / kBlockK;
    p.l2_hint = tuning_get("res_l2_hint");
    p.seg_chunks = p.n_chunks;
    p.bias = bias;
    p.skip_acc = skip_acc; p.skip_first = skip_first;
    CUtensorMap ma, mhi, mho, mw, mx;
    if (skip_acc) {                      // [16 rows][512] bf16: hi rows 0..7, lo rows 8..15; each CTA loads 8 rows per chunk
        const uint64_t dims[2] = {kNCh, 16};
        const uint64_t strides[1] = {kNCh * 2};
        const uint32_t box[2] = {kBlockK, 8};
        if (int e = make_tmap_bf16(&mx, w16_layer, 2, dims, strides, box)) return e;
    }
    if (int e = act_map(&ma, acts, kNCh, T, batch)) return e;
    if (int e = h_map(&mhi, h_in, T, batch, h_batch_rows)) return e;
    if (int e = h_map(&mho, h_out, T, batch, h_batch_rows)) return e;
    if (int e = weight_half_map(&mw, w_res, kNCh, kNCh)) return e;
    return launch<RES, 0, 0>(ma, mhi, mw, mho, p, stream, skip_acc ? &mx : nullptr);
}

// Residual update fed by a segmented K instead of a 1x1 (training direction: data gradients, glow.py:159-166 under
// autograd):  h_out[b,t,:] = h_in[b,t,:] + bias + sum_s W[:, s*C : (s+1)*C] A_s[b, t + shift0 + s*dshift, :]
// A_s = a1 if bit s of seg_mask else a0, bf16 [B,T,C] (C % 64 == 0); w bf16 [512][n_seg*C]; bias fp32 [512];
// h_in / h_out bf16 [B, h_batch_rows, 512] (may alias).  Dilated taps = shifted segments of one tensor; the
// res_skip data gradient = segments [g_h | g_skip] of two tensors.
int tc2_wn_res_seg(const void* a0, const void* a1, int n_seg, int seg_mask, const void* w, const float* bias,
                   const void* h_in, void* h_out, int batch, int T, long long h_batch_rows, int C, int shift0, int dshift,
                   cudaStream_t stream) {
    using namespace tc2;
    WGB_REQUIRE(a0 && w && bias && h_in && h_out, "null pointer");
    WGB_REQUIRE(seg_mask == 0 || a1 != nullptr, "seg_mask selects a1, which is null");
    WGB_REQUIRE(h_batch_rows >= T, "h_batch_rows (%lld) must be >= T (%d)", h_batch_rows, T);
    WGB_REQUIRE(C > 0 && C % kBlockK == 0 && n_seg >= 1 && n_seg <= 31, "bad segment shape (C=%d n_seg=%d)", C, n_seg);
    Params p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = 2; p.ppi = 1; p.n_chunks = n_seg * C / kBlockK;
    p.l2_hint = tuning_get("res_l2_hint");
    p.seg_chunks = C / kBlockK; p.seg_shift0 = shift0; p.seg_dshift = dshift; p.seg_mask = seg_mask;
    p.bias = bias;
    CUtensorMap ma, ma1, mhi, mho, mw;
    if (int e = act_map(&ma, a0, C, T, batch)) return e;
    if (int e = act_map(&ma1, a1 ? a1 : a0, C, T, batch)) return e;
    if (int e = h_map(&mhi, h_in, T, batch, h_batch_rows)) return e;
    if (int e = h_map(&mho, h_out, T, batch, h_batch_rows)) return e;
    if (int e = weight_half_map(&mw, w, kNCh, n_seg * C)) return e;
    return launch<RES, 0, 0>(ma, mhi, mw, mho, p, stream, &ma1);
}

int tc2_wn_res_taps(const void* a, const void* w, const float* bias, const void* h_in, void* h_out, int batch, int T,
                    long long h_batch_rows, int C, int taps, int dilation, cudaStream_t stream) {
    WGB_REQUIRE(taps >= 1 && taps % 2 == 1 && dilation >= 1, "taps must be odd (got %d), dilation >= 1", taps);
    return tc2_wn_res_seg(a, nullptr, taps, 0, w, bias, h_in, h_out, batch, T, h_batch_rows, C, -((taps - 1) / 2) * dilation,
                          dilation, stream);
}

int tc2_wn_skip_end(const void* acts_all, int n_layers, const void* w_skip, const float* w_end, const float* b_end,
                    float* x, const float* w_mix, float* log_s, int batch, int T, int n_half, int direction,
                    const float* next_w_start, const float* next_b_start, int next_n_half, void* h_next,
                    long long h_next_batch_rows, cudaStream_t stream) {
    using namespace tc2;
    WGB_REQUIRE(acts_all && w_skip && w_end && b_end && x, "null pointer");
    WGB_REQUIRE(n_layers == 8, "tensor-core skip GEMM is specialised for 8 layers (got %d)", n_layers);
    WGB_REQUIRE(n_half >= 1 && n_half <= 4, "n_half must be in 1..4 (got %d)", n_half);
    WGB_REQUIRE(direction == 0 || direction == 1, "direction must be 0 (infer) or 1 (forward)");
    WGB_REQUIRE(direction == 1 ? log_s != nullptr : w_mix != nullptr, "missing log_s / w_mix for this direction");
    Params p{};
    if (int e = fill_common(p, batch, T)) return e;
    p.n_pass = 2; p.ppi = 2; p.n_chunks = n_layers * kNCh / kBlockK;
    p.w_end = w_end; p.b_end = b_end; p.x = x; p.w_mix = w_mix; p.log_s = log_s;
    if (h_next) {
        WGB_REQUIRE(direction == 0, "the fused WN.start of the next flow exists for the infer direction only");
        WGB_REQUIRE(next_w_start && next_b_start && next_n_half >= 1 && next_n_half <= 4, "bad next-flow start arguments");
        WGB_REQUIRE(h_next_batch_rows >= T, "h_next_batch_rows must be >= T");
        p.next_w_start = next_w_start; p.next_b_start = next_b_start; p.next_n_half = next_n_half;
        p.h_next = static_cast<__nv_bfloat16*>(h_next);
        p.h_next_batch_rows = h_next_batch_rows;
    }
    CUtensorMap ma, mw;
    if (int e = act_map(&ma, acts_all, kNCh, T, batch * n_layers)) return e;
    if (int e = weight_half_map(&mw, w_skip, kNCh, n_layers * kNCh)) return e;
#define WGB_SKIP_CASE(NH)                                                                      \
    case NH:                                                                                   \
        return direction == 0 ? launch<SKIP_END, NH, 0>(ma, ma, mw, ma, p, stream)             \
                              : launch<SKIP_END, NH, 1>(ma, ma, mw, ma, p, stream);
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
