// WN gate GEMM on CTA pairs: tcgen05.mma.cta_group::2 (UMMA M = 256 over two SMs).
//
// Same math as MODE_GATE in wn_tc.cu (in_layers k=3 dilated + cond 1x1 + bias -> tanh*sigmoid, reference
// glow.py:159-162), but two CTAs of a cluster share every weight tile: each CTA TMA-loads its own 128
// time rows of the activation operand and only HALF (128 of 256 rows) of the weight tile, and the
// leader CTA issues one M=256 x N=256 MMA that reads both halves.  Per CTA and K chunk that is
// 16 KB + 16 KB instead of 16 KB + 32 KB from L2, and the tensor core reads each weight byte from
// shared memory once per pair instead of once per CTA -- the point on a power-capped part.
//
// Protocol (per pair): both CTAs run a TMA producer; all loads complete_tx on the LEADER's full barrier,
// which the leader's producer arms with the byte count of both CTAs.  The leader's MMA thread commits
// with .multicast::cluster to the empty / tmem-full barriers of both CTAs; the epilogue warps of both
// CTAs arrive (the peer remotely) on the leader's tmem-empty barrier.
#include "common.cuh"
#include "ptx.cuh"

namespace wgb {

namespace tc2 {

constexpr int kBlockM = 128;          // rows per CTA (UMMA M = 256 per pair)
constexpr int kBlockN = 256;
constexpr int kHalfN = 128;           // weight rows each CTA loads
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kStages = 6;
constexpr int kABytes = kBlockM * kBlockK * 2;      // 16 KB
constexpr int kBBytes = kHalfN * kBlockK * 2;       // 16 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kTmemCols = 512;
constexpr int kThreads = 192;
constexpr int kNCh = 512;
constexpr int kNCond = 640;
constexpr int kBiasOff = kStages * kStageBytes;     // fp32 [1024]: tanh biases as is, sigmoid biases pre-halved
constexpr int kBiasBytes = 2 * kNCh * 4;
constexpr int kBarOff = kBiasOff + kBiasBytes;
constexpr int kSmemTotal = 1024 + kBarOff + 256;

struct Params {
    int batch, T, tiles_per_b, n_tiles;
    int n_pass, n_chunks, dilation;
    const float* bias;
    __nv_bfloat16* acts_out;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared variable of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// TMA loads whose completion is signalled on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all prior tcgen05 ops of this thread are done) on the barrier at this smem offset in both CTAs
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gate_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_cond,
            const __grid_constant__ CUtensorMap map_w, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOff);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_h);
        tma_prefetch_desc(&map_cond);
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
    float* s_bias = reinterpret_cast<float*>(smem + kBiasOff);
    if (warp >= 2) {
        // packed column c of pass p: tanh row for (c & 255) < 128, else the matching sigmoid row, whose
        // pre-activation is halved (sigmoid(b) = 0.5 tanh(b/2) + 0.5)
        for (int i = threadIdx.x - 64; i < 2 * kNCh; i += 128) s_bias[i] = p.bias[i] * ((i & 128) ? 0.5f : 1.f);
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int n_pairs = gridDim.x >> 1;
    const int pair_id = blockIdx.x >> 1;
    const int n_pair_tiles = (p.n_tiles + 1) >> 1;
    const int n_items = n_pair_tiles * p.n_pass;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int item = pair_id; item < n_items; item += n_pairs) {
                const int tile = 2 * (item / p.n_pass) + static_cast<int>(rank);
                const int pass = item % p.n_pass;
                const bool valid = tile < p.n_tiles;
                const int b = valid ? tile / p.tiles_per_b : 0;
                const int t0 = valid ? (tile % p.tiles_per_b) * kBlockM : p.T + 4 * kBlockM;   // all rows OOB -> zeros
                for (int kc = 0; kc < p.n_chunks; ++kc) {
                    mbar_wait(&empty_bar[s], ph ^ 1, 100 + s);
                    uint8_t* sa = smem + s * kStageBytes;
                    uint8_t* sb = sa + kABytes;
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * kStageBytes);
                    const uint32_t bar = mapa_u32(&full_bar[s], 0);
                    if (kc < 24) {
                        const int tap = kc >> 3;
                        tma_load_3d_2sm(sa, &map_h, bar, (kc & 7) * kBlockK, t0 + (tap - 1) * p.dilation, b);
                    } else {
                        tma_load_3d_2sm(sa, &map_cond, bar, (kc - 24) * kBlockK, t0, b);
                    }
                    tma_load_2d_2sm(sb, &map_w, bar, kc * kBlockK, pass * kBlockN + static_cast<int>(rank) * kHalfN);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16_f32(2 * kBlockM, kBlockN);
            int s = 0;
            uint32_t ph = 0, acc_it = 0;
            for (int item = pair_id; item < n_items; item += n_pairs, ++acc_it) {
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
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5, both CTAs)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        for (int item = pair_id; item < n_items; item += n_pairs, ++acc_it) {
            const int tile = 2 * (item / p.n_pass) + static_cast<int>(rank);
            const int pass = item % p.n_pass;
            const bool valid = tile < p.n_tiles;
            const int b = valid ? tile / p.tiles_per_b : 0;
            const int t = valid ? (tile % p.tiles_per_b) * kBlockM + row : p.T;
            const bool live = t < p.T;
            const size_t grow = static_cast<size_t>(b) * p.T + t;
            const uint32_t as = acc_it & 1, aph = (acc_it >> 1) & 1;
            mbar_wait(&tfull_bar[as], aph, 400 + as);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * kBlockN;
            const float4* bt4 = reinterpret_cast<const float4*>(s_bias + pass * kBlockN);
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
                }
                if (live) {
                    uint4* d4 = reinterpret_cast<uint4*>(dst + ch * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty_bar[as], 0));
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

}  // namespace tc2

int tc2_wn_gate(const void* h, const void* cond, const void* w_packed, const float* bias, void* acts, int batch, int T,
                int dilation, cudaStream_t stream) {
    using namespace tc2;
    WGB_REQUIRE(h && cond && w_packed && bias && acts, "null pointer");
    WGB_REQUIRE(batch > 0 && T > 0 && dilation >= 1, "bad shape");
    Params p{};
    p.batch = batch; p.T = T;
    p.tiles_per_b = ceil_div(T, kBlockM);
    p.n_tiles = batch * p.tiles_per_b;
    p.n_pass = 4; p.n_chunks = (3 * kNCh + kNCond) / kBlockK; p.dilation = dilation;
    p.bias = bias; p.acts_out = static_cast<__nv_bfloat16*>(acts);

    CUtensorMap mh, mc, mw;
    {
        const uint64_t dims[3] = {kNCh, static_cast<uint64_t>(T), static_cast<uint64_t>(batch)};
        const uint64_t strides[2] = {kNCh * 2, static_cast<uint64_t>(kNCh) * 2 * T};
        const uint32_t box[3] = {kBlockK, kBlockM, 1};
        if (int e = make_tmap_bf16(&mh, h, 3, dims, strides, box)) return e;
    }
    {
        const uint64_t dims[3] = {kNCond, static_cast<uint64_t>(T), static_cast<uint64_t>(batch)};
        const uint64_t strides[2] = {kNCond * 2, static_cast<uint64_t>(kNCond) * 2 * T};
        const uint32_t box[3] = {kBlockK, kBlockM, 1};
        if (int e = make_tmap_bf16(&mc, cond, 3, dims, strides, box)) return e;
    }
    {
        const uint64_t k = 3 * kNCh + kNCond;
        const uint64_t dims[2] = {k, 2 * kNCh};
        const uint64_t strides[1] = {k * 2};
        const uint32_t box[2] = {kBlockK, kHalfN};
        if (int e = make_tmap_bf16(&mw, w_packed, 2, dims, strides, box)) return e;
    }
    WGB_CUDA_TRY(cudaFuncSetAttribute(gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal));
    const int n_items = ((p.n_tiles + 1) / 2) * p.n_pass;
    int pairs = sm_count() / 2;
    if (n_items < pairs) pairs = n_items;
    gate_kernel<<<2 * pairs, kThreads, kSmemTotal, stream>>>(mh, mc, mw, p);
    WGB_LAUNCH_CHECK();
    return WGB_OK;
}

}  // namespace wgb
