// Skip path + WN.end as ONE skinny GEMM on tcgen05 (HBM-bound), then the affine coupling in the epilogue.
//
// The reference sums the skip halves of the eight res_skip_layers into `output` and applies the 1x1 `end` conv
// (glow.py:167-175).  Both are linear and nothing sits between them, so
//     end(sum_i W_skip_i acts_i + b_skip_i) = sum_i (W_end W_skip_i) acts_i + (W_end sum_i b_skip_i + b_end):
// a [2 n_half <= 8] x [8*512] weight.  The composed weight is split into bf16 hi + lo parts (rows 0..7 / 8..15 of
// one N = 16 operand, summed in the epilogue), so it is accurate to ~2^-17 while the GEMM stays a single UMMA
// N = 16 sweep over the stored gated activations.  Per group step that is 2*4096*16 FLOP on the tensor pipe
// instead of 2*4096*512 + 2*512*8, and the kernel runs at the speed the 8 KB/step of acts_all stream from HBM.
// The coupling, the invertible 1x1 conv (infer) / log_s (forward) and optionally WN.start of the next flow follow
// in the fp32 epilogue exactly as in the K = 4096 x N = 512 kernels (wn_tc.cu / wn_tc2.cu), which remain as the
// un-composed variants (WGB_SKIP_KERNEL=pair|single).
#include "common.cuh"
#include "ptx.cuh"

namespace wgb {

namespace skip16 {

constexpr int kBlockM = 128;
constexpr int kBlockN = 16;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kStages = 8;
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytes = kBlockN * kBlockK * 2;      // 2 KB
constexpr int kStageBytes = kABytes + kBBytes;      // 18 KB (both parts 1024 B aligned)
constexpr int kTmemCols = 32;
constexpr int kThreads = 192;
constexpr int kNCh = 512;
constexpr int kStartOff = kStages * kStageBytes;    // next-flow WN.start: [512][4] weights + [512] bias, fp32
constexpr int kStartBytes = kNCh * 5 * 4;
constexpr int kBarOff = kStartOff + kStartBytes;
constexpr int kSmemTotal = 1024 + kBarOff + 256;

struct Params {
    int batch, T, tiles_per_b, n_tiles, n_chunks;
    const float* b_end;           // [8], skip biases folded through W_end
    const float* skip_acc;        // optional [B*T][8]: contributions of the layers accumulated by wgb_tc2_wn_res
    const float* next_w_mix;      // forward, optional: W of the NEXT flow's Invertible1x1Conv ([8][8], top-left C'xC'),
                                  // applied to the updated row before its fused WN.start (glow.py:233 of flow k+1)
    float* x;                     // flow state [B,T,8]
    const float* w_mix;           // infer: W^-1 [8][8]
    float* log_s;                 // forward: [B,n_half,T]
    const float* next_w_start;    // optional fused WN.start of the next flow (infer)
    const float* next_b_start;
    __nv_bfloat16* h_next;
    int next_n_half;
    long long h_next_batch_rows;  // row pitch per utterance of h_next (T, or more in the padded layout)
};

template <int NHALF, int DIR>
__global__ void __launch_bounds__(kThreads, 1)
skip16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* s_ws = reinterpret_cast<float*>(smem + kStartOff);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    pdl_launch_dependents();          // the next kernel may queue up behind this persistent grid right away

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_w);
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
    if (warp >= 2 && p.h_next) {                  // [512][4] zero-padded weight rows, then [512] bias
        const int i0 = threadIdx.x - 64;
        for (int i = i0; i < kNCh * 4; i += 128)
            s_ws[i] = (i & 3) < p.next_n_half ? p.next_w_start[(i >> 2) * p.next_n_half + (i & 3)] : 0.f;
        for (int i = i0; i < kNCh; i += 128) s_ws[kNCh * 4 + i] = p.next_b_start[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                       // everything above touched only weights; activations and the flow state start here

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int b = tile / p.tiles_per_b;
                const int t0 = (tile % p.tiles_per_b) * kBlockM;
                for (int kc = 0; kc < p.n_chunks; ++kc) {
                    mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
                    uint8_t* sa = smem + s * kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                    tma_load_3d(sa, &map_a, &full_bar[s], (kc & 7) * kBlockK, t0, (kc >> 3) * p.batch + b);   // layer kc>>3
                    tma_load_2d(sa + kABytes, &map_w, &full_bar[s], kc * kBlockK, 0);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(kBlockM, kBlockN);
            int s = 0;
            uint32_t ph = 0, acc_it = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++acc_it) {
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
                        umma_bf16_ss(d_tmem, umma_desc_sw128(a_addr + k * kUmmaK * 2), umma_desc_sw128(b_addr + k * kUmmaK * 2),
                                     idesc, (kc | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++acc_it) {
            const int b = tile / p.tiles_per_b;
            const int t = (tile % p.tiles_per_b) * kBlockM + row;
            const bool live = t < p.T;
            const size_t grow = static_cast<size_t>(b) * p.T + t;
            const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
            mbar_wait(&tfull_bar[as], aph, 400 + as);
            tc_fence_after_sync();
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kBlockN, v);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (!live) continue;

            float outv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) outv[j] = __uint_as_float(v[j]) + __uint_as_float(v[8 + j]) + __ldg(p.b_end + j);   // hi + lo
            if (p.skip_acc) {
                const float4 s0 = *reinterpret_cast<const float4*>(p.skip_acc + grow * 8);
                const float4 s1 = *reinterpret_cast<const float4*>(p.skip_acc + grow * 8 + 4);
                outv[0] += s0.x; outv[1] += s0.y; outv[2] += s0.z; outv[3] += s0.w;
                outv[4] += s1.x; outv[5] += s1.y; outv[6] += s1.z; outv[7] += s1.w;
            }
            constexpr int C = 2 * NHALF, BASE = 8 - C;
            float* xr = p.x + grow * 8;
            float xv[8];
            *reinterpret_cast<float4*>(&xv[0]) = *reinterpret_cast<const float4*>(xr);
            *reinterpret_cast<float4*>(&xv[4]) = *reinterpret_cast<const float4*>(xr + 4);
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
            if constexpr (DIR == 1) {
                if (p.next_w_mix) {                          // the next flow's 1x1 conv on its last C' = 2 n_half' channels
                    const int cn = 2 * p.next_n_half, bn = 8 - cn;
                    float xo[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float acc = 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (i >= bn && c >= bn) acc = fmaf(__ldg(p.next_w_mix + (i - bn) * 8 + (c - bn)), xv[c], acc);
                        xo[i] = i >= bn ? acc : xv[i];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) xv[i] = xo[i];
                }
            }
            *reinterpret_cast<float4*>(xr) = *reinterpret_cast<const float4*>(&xv[0]);
            *reinterpret_cast<float4*>(xr + 4) = *reinterpret_cast<const float4*>(&xv[4]);
            if (p.h_next) {
                // WN.start of the next flow (glow.py:156) on the freshly updated row
                const int nb = 8 - 2 * p.next_n_half;
                float a0[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) a = (i == nb + j) ? xv[i] : a;
                    a0[j] = a;                                  // weight columns >= next_n_half are zero
                }
                const float4* ws4 = reinterpret_cast<const float4*>(s_ws);
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

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

template <int NHALF, int DIR>
static int launch(const CUtensorMap& ma, const CUtensorMap& mw, const Params& p, cudaStream_t stream) {
    static_assert(kSmemTotal <= 232448, "dynamic shared memory over the 227 KB per-CTA limit");
    auto kern = skip16_kernel<NHALF, DIR>;
    WGB_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
    WGB_CUDA_TRY(launch_pdl(kern, grid, kThreads, kSmemTotal, stream, ma, mw, p));
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace skip16

int tc_wn_skip16_end(const void* acts_all, int n_layers, const void* w16, const float* b_end, float* x, const float* w_mix,
                     float* log_s, int batch, int T, int n_half, int direction, const float* next_w_start,
                     const float* next_b_start, int next_n_half, void* h_next, long long h_next_batch_rows,
                     const float* skip_acc, const float* next_w_mix, cudaStream_t stream) {
    using namespace skip16;
    WGB_REQUIRE(acts_all && w16 && b_end && x, "null pointer");
    WGB_REQUIRE(n_layers >= 1 && batch > 0 && T > 0, "bad shape");
    WGB_REQUIRE(n_half >= 1 && n_half <= 4, "n_half must be in 1..4 (got %d)", n_half);
    WGB_REQUIRE(direction == 0 || direction == 1, "direction must be 0 (infer) or 1 (forward)");
    WGB_REQUIRE(direction == 1 ? log_s != nullptr : w_mix != nullptr, "missing log_s / w_mix for this direction");
    Params p{};
    p.batch = batch; p.T = T;
    p.tiles_per_b = ceil_div(T, kBlockM);
    p.n_tiles = batch * p.tiles_per_b;
    p.n_chunks = n_layers * kNCh / kBlockK;
    p.b_end = b_end; p.x = x; p.w_mix = w_mix; p.log_s = log_s; p.skip_acc = skip_acc;
    WGB_REQUIRE(!next_w_mix || (direction == 1 && h_next), "next_w_mix goes with direction 1 and a fused next-flow start");
    p.next_w_mix = next_w_mix;
    if (h_next) {
        WGB_REQUIRE(next_w_start && next_b_start && next_n_half >= 1 && next_n_half <= 4, "bad next-flow start arguments");
        WGB_REQUIRE(direction == 0 || next_w_mix, "forward: the next flow's 1x1 conv must run before its WN.start (pass next_w_mix)");
        WGB_REQUIRE(h_next_batch_rows >= T, "h_next_batch_rows must be >= T");
        p.next_w_start = next_w_start; p.next_b_start = next_b_start; p.next_n_half = next_n_half;
        p.h_next = static_cast<__nv_bfloat16*>(h_next);
        p.h_next_batch_rows = h_next_batch_rows;
    }
    CUtensorMap ma, mw;
    {
        const uint64_t dims[3] = {kNCh, static_cast<uint64_t>(T), static_cast<uint64_t>(batch) * n_layers};
        const uint64_t strides[2] = {kNCh * 2, static_cast<uint64_t>(kNCh) * 2 * T};
        const uint32_t box[3] = {kBlockK, kBlockM, 1};
        if (int e = make_tmap_bf16(&ma, acts_all, 3, dims, strides, box)) return e;
    }
    {
        const uint64_t k = static_cast<uint64_t>(n_layers) * kNCh;
        const uint64_t dims[2] = {k, kBlockN};
        const uint64_t strides[1] = {k * 2};
        const uint32_t box[2] = {kBlockK, kBlockN};
        if (int e = make_tmap_bf16(&mw, w16, 2, dims, strides, box)) return e;
    }
#define WGB_SKIP16_CASE(NH)                                                                  \
    case NH:                                                                                 \
        return direction == 0 ? launch<NH, 0>(ma, mw, p, stream) : launch<NH, 1>(ma, mw, p, stream);
    switch (n_half) {
        WGB_SKIP16_CASE(1)
        WGB_SKIP16_CASE(2)
        WGB_SKIP16_CASE(3)
        WGB_SKIP16_CASE(4)
    }
#undef WGB_SKIP16_CASE
    return fail(WGB_ERR_ARGUMENT, "unreachable");
}

}  // namespace wgb
