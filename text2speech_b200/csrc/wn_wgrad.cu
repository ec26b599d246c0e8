// Weight gradients of the WN convolutions on tcgen05 (training direction, SURVEY section 8(f)2;
// what autograd computes for glow.py:141-152 convs inside waveglow/train.py:116-124).
//
//   dW[tap][m][n] = sum_{b,t} G[b, t, m] * X[b, t + (tap - (taps-1)/2) * dilation, n]
//
// G (gradient w.r.t. the conv output) and X (the conv input) are the channels-last bf16 activations the forward /
// backward kernels already hold, so the contraction runs over the ROW index of both operands: both are "MN-major"
// UMMA operands.  A TMA box of {64 channels, 64 rows} lands in shared memory as 64 rows of 128 B (SWIZZLE_128B),
// which is exactly the canonical MN-major SW128 layout: 8-row groups 1024 B apart (SBO), the next 64-channel block
// one box (8 KB) further (LBO).  No transposed copy of any activation is ever made.  Rows shifted out of
// [0, T) by the tap offset are zero-filled by TMA (= the conv's zero padding); rows >= T of the last chunk of an
// utterance are zero in G as well, so ragged T costs nothing.
//
// Tiling: one item = (m tile of 256, n tile of 256, tap, K split); K = all (utterance, 64-row chunk) pairs, split so
// that the grid fills the SMs; partial sums leave through fp32 atomics (red.global.add) into dW (a few MB, L2 resident).
// A 256 x 256 tile is two UMMA M = 128 accumulators (all 512 TMEM columns) fed by the same B chunk: 64 KB of operands
// per 2 x (128 x 256 x 64) MACs = 128 B/clk/SM from L2 instead of the 192 B/clk of a 128 x 256 tile, which is what
// bounds a single-CTA kernel here (the items are ~250 chunks long, so the un-overlapped epilogue is noise).
// Warp roles as in wn_tc.cu: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 epilogue.
#include "common.cuh"
#include "ptx.cuh"

namespace wgb {
namespace wgrad {

constexpr int kM = 256;                         // two UMMA M = 128 halves
constexpr int kN = 256;
constexpr int kRows = 64;                       // K chunk = 64 time steps
constexpr int kBox = 64 * kRows * 2;            // one {64 ch, 64 rows} box = 8 KB
constexpr int kABytes = (kM / 64) * kBox;       // 32 KB
constexpr int kBBytes = (kN / 64) * kBox;       // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStages = 3;
constexpr int kThreads = 192;
constexpr int kSmem = 1024 + kStages * kStageBytes + 256;

struct Params {
    int batch, T, chunks_per_b, chunks_total, splits, chunks_per_split;
    int m_tiles, n_tiles, taps, dilation;
    int ca, cb;
    float* out;                                 // [taps][ca][cb] fp32
    int vec_ok;                                 // out 16 B aligned and cb % 4 == 0: vector reds allowed
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MN-major SW128 operand: LBO = distance between 64-element blocks along M/N, SBO = distance between 8-row K groups
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int i = 0; i < kStages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&tfull_bar[0], 1);
        mbar_init(&tempty_bar[0], 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles = p.m_tiles * p.n_tiles * p.taps;
    const int n_items = tiles * p.splits;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int split = item / tiles;              // neighbouring CTAs share the chunk range (L2 reuse)
                const int tile = item % tiles;
                const int m_tile = tile % p.m_tiles;
                const int n_tile = (tile / p.m_tiles) % p.n_tiles;
                const int tap = tile / (p.m_tiles * p.n_tiles);
                const int shift = (tap - (p.taps - 1) / 2) * p.dilation;
                // skinny operands (ca = 64: the hi/lo stack) load and multiply only the 64-channel boxes that exist
                const int a_boxes = min(kM / 64, (p.ca - m_tile * kM + 63) / 64);
                const int c_begin = split * p.chunks_per_split;
                const int c_end = min(c_begin + p.chunks_per_split, p.chunks_total);
                for (int c = c_begin; c < c_end; ++c) {
                    const int b = c / p.chunks_per_b;
                    const int t0 = (c % p.chunks_per_b) * kRows;
                    mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
                    uint8_t* sa = smem + s * kStageBytes;
                    uint8_t* sb = sa + kABytes;
                    mbar_arrive_expect_tx(&full_bar[s], a_boxes * kBox + kBBytes);
#pragma unroll
                    for (int j = 0; j < kM / 64; ++j)
                        if (j < a_boxes) tma_load_3d(sa + j * kBox, &map_a, &full_bar[s], m_tile * kM + j * 64, t0, b);
#pragma unroll
                    for (int j = 0; j < kN / 64; ++j)
                        tma_load_3d(sb + j * kBox, &map_b, &full_bar[s], n_tile * kN + j * 64, t0 + shift, b);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // kind::f16, D fp32, A/B bf16, both MN-major (bits 15, 16)
            constexpr uint32_t idesc = umma_idesc_bf16_f32(128, kN) | (1u << 15) | (1u << 16);
            int s = 0;
            uint32_t ph = 0, acc_it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
                const int split = item / tiles;
                const int halves = (p.ca - ((item % tiles) % p.m_tiles) * kM) > 128 ? 2 : 1;
                const int c_begin = split * p.chunks_per_split;
                const int c_end = min(c_begin + p.chunks_per_split, p.chunks_total);
                mbar_wait(&tempty_bar[0], (acc_it & 1) ^ 1, 200);
                tc_fence_after_sync();
                for (int c = c_begin; c < c_end; ++c) {
                    mbar_wait(&full_bar[s], ph, 300 + s);
                    tc_fence_after_sync();
                    const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
                    const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
                    for (int k = 0; k < kRows / 16; ++k) {     // 16 rows = two 8-row groups = 2048 B per UMMA
                        const uint64_t db = umma_desc_sw128_mn(b_addr + k * 2048, kBox);
#pragma unroll
                        for (int half = 0; half < 2; ++half)   // channels half*128 .. +127 of the m tile -> accumulator half
                            if (half < halves) umma_bf16_ss(tmem_base + half * kN, umma_desc_sw128_mn(a_addr + half * 2 * kBox + k * 2048, kBox),
                                         db, idesc, (c > c_begin) || (k != 0));
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[0]);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++acc_it) {
            const int split = item / tiles;
            const int tile = item % tiles;
            const int m_tile = tile % p.m_tiles;
            const int n_tile = (tile / p.m_tiles) % p.n_tiles;
            const int tap = tile / (p.m_tiles * p.n_tiles);
            const bool has_work = split * p.chunks_per_split < p.chunks_total;
            mbar_wait(&tfull_bar[0], acc_it & 1, 400);
            tc_fence_after_sync();
            const int n_left = p.cb - n_tile * kN;
            const int halves = (p.ca - m_tile * kM) > 128 ? 2 : 1;
#pragma unroll 1
            for (int half = 0; half < halves; ++half) {
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * kN;
                const int m = m_tile * kM + half * 128 + row;
                float* dst = p.out + (static_cast<size_t>(tap) * p.ca + m) * p.cb + n_tile * kN;
#pragma unroll 1
                for (int ch = 0; ch < 8; ++ch) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(taddr + ch * 32, v);
                    tmem_ld_wait();
                    if (has_work && m < p.ca) {
                        if (ch * 32 + 32 <= n_left && p.vec_ok) {         // 8 x red.v4 instead of 32 scalar reds
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                red_add_v4(dst + ch * 32 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                           __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (ch * 32 + j < n_left) atomicAdd(dst + ch * 32 + j, __uint_as_float(v[j]));
                        }
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[0]);
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace wgrad

// dW[tap][m][n] (+)= sum_{b,t} g[b,t,m] x[b, t + (tap - (taps-1)/2)*dilation, n];  g bf16 [B,T,ca] (ca % 64 == 0),
// x bf16 [B,T,cb] (cb % 8 == 0), dw fp32 [taps][ca][cb].  accumulate = 0 clears dw first (on the stream).
int tc_wgrad(const void* g, const void* x, float* dw, int batch, int T, int ca, int cb, int taps, int dilation,
             int accumulate, cudaStream_t stream) {
    using namespace wgrad;
    WGB_REQUIRE(g && x && dw, "null pointer");
    WGB_REQUIRE(batch > 0 && T > 0, "batch (%d) and T (%d) must be positive", batch, T);
    WGB_REQUIRE(ca > 0 && ca % 64 == 0 && cb > 0 && cb % 8 == 0, "ca must be a multiple of 64 and cb of 8 (ca=%d cb=%d)", ca, cb);
    WGB_REQUIRE(taps >= 1 && taps % 2 == 1 && dilation >= 1, "taps must be odd (got %d), dilation >= 1", taps);
    Params p{};
    p.batch = batch; p.T = T;
    p.chunks_per_b = ceil_div(T, kRows);
    p.chunks_total = batch * p.chunks_per_b;
    p.m_tiles = ceil_div(ca, kM);
    p.n_tiles = ceil_div(cb, kN);
    p.taps = taps; p.dilation = dilation; p.ca = ca; p.cb = cb; p.out = dw;
    const int tiles = p.m_tiles * p.n_tiles * taps;
    // K splits so that one wave of items fills the SMs (every extra split costs a full tile of atomics), but at least
    // 16 chunks per item
    int splits = sm_count() / tiles;
    splits = splits < 1 ? 1 : splits;
    p.vec_ok = (reinterpret_cast<uintptr_t>(dw) % 16 == 0 && cb % 4 == 0) ? 1 : 0;
    const int max_splits = p.chunks_total / 16 > 0 ? p.chunks_total / 16 : 1;
    if (splits > max_splits) splits = max_splits;
    p.chunks_per_split = ceil_div(p.chunks_total, splits);
    p.splits = ceil_div(p.chunks_total, p.chunks_per_split);
    if (!accumulate)
        WGB_CUDA_TRY(cudaMemsetAsync(dw, 0, static_cast<size_t>(taps) * ca * cb * sizeof(float), stream));
    CUtensorMap ma, mb;
    {
        const uint64_t dims[3] = {static_cast<uint64_t>(ca), static_cast<uint64_t>(T), static_cast<uint64_t>(batch)};
        const uint64_t strides[2] = {static_cast<uint64_t>(ca) * 2, static_cast<uint64_t>(ca) * 2 * T};
        const uint32_t box[3] = {64, kRows, 1};
        if (int e = make_tmap_bf16(&ma, g, 3, dims, strides, box)) return e;
    }
    {
        const uint64_t dims[3] = {static_cast<uint64_t>(cb), static_cast<uint64_t>(T), static_cast<uint64_t>(batch)};
        const uint64_t strides[2] = {static_cast<uint64_t>(cb) * 2, static_cast<uint64_t>(cb) * 2 * T};
        const uint32_t box[3] = {64, kRows, 1};
        if (int e = make_tmap_bf16(&mb, x, 3, dims, strides, box)) return e;
    }
    WGB_CUDA_TRY(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    const int items = tiles * p.splits;
    const int grid = items < sm_count() ? items : sm_count();
    wgrad_kernel<<<grid, kThreads, kSmem, stream>>>(ma, mb, p);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace wgb
