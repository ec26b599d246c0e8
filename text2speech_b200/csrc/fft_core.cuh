// 1024-point real FFT per warp (512-point complex FFT, 16 complex values per lane) -- the butterfly form of the reference's
// conv-basis STFT (utils/stft.py:46-60: forward_basis = [cos; -sin](2 pi k n / L) * window, inverse_basis = pinv of it).
//
// Written so that the SAME code compiles for the device and for the host: on the host a "warp" is a loop over 32 lanes
// between the stage boundaries (tests/fft_host_check.cpp pins the index maps, twiddles and the real-FFT split against a
// naive DFT without a GPU).
//
//   512 = 8 x 8 x 8:   n = n0 + 8 n1 + 64 n2,   k = k2 + 8 k1 + 64 k0
//   stage 1  A[n0,n1,k2] = sum_n2 z[n] W8^(n2 k2)                       lane owns (n0,n1) = lane + 32u, u = 0,1
//            * W64^(n1 k2)                                      -> smem A [n0 + 8 n1 + 72 k2]
//   stage 2  B[n0,k1,k2] = sum_n1 A' W8^(n1 k1)                         lane owns (n0,k2) = lane + 32u
//            * W512^(n0 (k2 + 8 k1))                            -> smem B [k2 + 8 k1 + 66 n0]
//   stage 3  Z[k] = sum_n0 B' W8^(n0 k0)                                lane owns (k2,k1) = lane + 32u
//   A lane starts with z[lane + 32 j] and ends with Z[lane + 32 j], j = 0..15 (j = u + 2 n2 on the way in, u + 2 k0 on
//   the way out), so consecutive lanes touch consecutive addresses of the signal, and the partner bin 512 - k of the
//   real-FFT split lives in lane (32 - lane) & 31: one shuffle per value, no third exchange.
//   The row pitches 72 and 66 make all four shared-memory passes conflict-free for 8-byte accesses.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define WGB_FFT_HD __host__ __device__ __forceinline__
#else
#define WGB_FFT_HD inline
#endif

namespace wgb {
namespace fft {

struct cf {
    float x, y;
};

constexpr int kBufA = 576;              // cf per warp, first exchange (stage 1 -> 2): 8 rows of pitch 72
constexpr int kBufB = 528;              // second exchange (stage 2 -> 3): 8 rows of pitch 66
constexpr int kBufElems = kBufA + kBufB;
constexpr float kSqrtHalf = 0.70710678118654752440f;

WGB_FFT_HD cf cmul(cf a, cf b) { return cf{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
WGB_FFT_HD cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
WGB_FFT_HD cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }
WGB_FFT_HD cf cconj(cf a) { return cf{a.x, -a.y}; }
WGB_FFT_HD cf mul_neg_i(cf a) { return cf{a.y, -a.x}; }          // a * (-i)

// e^{-2 pi i num / den}
WGB_FFT_HD cf twiddle(int num, int den) {
#if defined(__CUDA_ARCH__)
    float s, c;
    sincospif(2.0f * static_cast<float>(num) / static_cast<float>(den), &s, &c);      // num / den is exact in fp32 here
    return cf{c, -s};
#else
    const double a = 2.0 * 3.14159265358979323846 * static_cast<double>(num) / static_cast<double>(den);
    return cf{static_cast<float>(cos(a)), static_cast<float>(-sin(a))};
#endif
}

// forward 8-point DFT in place, natural order in and out: a[k] <- sum_n a[n] e^{-2 pi i n k / 8}
WGB_FFT_HD void dft8(cf* a) {
    const cf b0 = cadd(a[0], a[4]), b1 = csub(a[0], a[4]);
    const cf b2 = cadd(a[2], a[6]), b3 = mul_neg_i(csub(a[2], a[6]));
    const cf b4 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    const cf b6 = cadd(a[3], a[7]), b7 = mul_neg_i(csub(a[3], a[7]));
    const cf e0 = cadd(b0, b2), e2 = csub(b0, b2), e1 = cadd(b1, b3), e3 = csub(b1, b3);      // DFT4 of the even samples
    const cf o0 = cadd(b4, b6), o2 = csub(b4, b6), o1 = cadd(b5, b7), o3 = csub(b5, b7);      // DFT4 of the odd samples
    const cf t1 = cf{(o1.x + o1.y) * kSqrtHalf, (o1.y - o1.x) * kSqrtHalf};                   // o1 * (1 - i) / sqrt 2
    const cf t2 = mul_neg_i(o2);
    const cf t3 = cf{(o3.y - o3.x) * kSqrtHalf, -(o3.x + o3.y) * kSqrtHalf};                  // o3 * (-1 - i) / sqrt 2
    a[0] = cadd(e0, o0);
    a[4] = csub(e0, o0);
    a[1] = cadd(e1, t1);
    a[5] = csub(e1, t1);
    a[2] = cadd(e2, t2);
    a[6] = csub(e2, t2);
    a[3] = cadd(e3, t3);
    a[7] = csub(e3, t3);
}

// Per-lane twiddles (fixed for the life of a warp: they live in registers)
struct LaneTw {
    cf s1[8];      // W64^((lane >> 3) k2), k2 = 0..7; the u = 1 butterfly multiplies by W16^k2 on top (n1 = (lane>>3) + 4u)
    cf s2[8];      // W512^((lane & 7) ((lane >> 3) + 8 k1)), k1 = 0..7; u = 1 multiplies by s2u on top (k2 = (lane>>3) + 4u)
    cf s2u;        // W128^(lane & 7)
    cf post;       // W1024^lane: the real-FFT split's e^{-2 pi i k / 1024} at k = lane + 32 j is post * W32^j
};

WGB_FFT_HD void lane_twiddles(int lane, LaneTw& t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        t.s1[k] = twiddle((lane >> 3) * k, 64);
        t.s2[k] = twiddle((lane & 7) * ((lane >> 3) + 8 * k), 512);
    }
    t.s2u = twiddle(lane & 7, 128);
    t.post = twiddle(lane, 1024);
}

// W16^k, k = 0..7 and W32^j, j = 0..15 as literals (folded into the instruction stream)
WGB_FFT_HD cf w16(int k) {
    constexpr float c[8] = {1.f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f,
                            0.f, -0.38268343236508977173f, -0.70710678118654752440f, -0.92387953251128675613f};
    constexpr float s[8] = {0.f, 0.38268343236508977173f, 0.70710678118654752440f, 0.92387953251128675613f,
                            1.f, 0.92387953251128675613f, 0.70710678118654752440f, 0.38268343236508977173f};
    return cf{c[k], -s[k]};
}
WGB_FFT_HD cf w32(int j) {
    constexpr float c[16] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f,
                             0.f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                             -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f, -0.98078528040323044913f};
    constexpr float s[16] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                             0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f,
                             1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                             0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f};
    return cf{c[j], -s[j]};
}

// ---- the three stages.  v[16]: this lane's values; buf_a / buf_b: the warp's two exchange buffers (kBufA, kBufB).  A
// warp-wide barrier goes between stage1 / stage2 and between stage2 / stage3; with two buffers nothing else is needed between
// back-to-back transforms (a buffer is rewritten only after a barrier that every lane reaches after its last read of it).
WGB_FFT_HD void stage1(int lane, const LaneTw& t, const cf* v, cf* buf) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        cf a[8];
#pragma unroll
        for (int n2 = 0; n2 < 8; ++n2) a[n2] = v[u + 2 * n2];
        dft8(a);
        buf[lane + 32 * u] = a[0];
#pragma unroll
        for (int k2 = 1; k2 < 8; ++k2) {
            cf x = cmul(a[k2], t.s1[k2]);
            if (u) x = cmul(x, w16(k2));
            buf[lane + 32 * u + 72 * k2] = x;
        }
    }
}
WGB_FFT_HD void stage2(int lane, const LaneTw& t, const cf* buf_a, cf* buf_b) {
    const int n0 = lane & 7;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int k2 = (lane >> 3) + 4 * u;
        cf a[8];
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) a[n1] = buf_a[n0 + 8 * n1 + 72 * k2];
        dft8(a);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            cf x = cmul(a[k1], t.s2[k1]);
            if (u) x = cmul(x, t.s2u);
            buf_b[k2 + 8 * k1 + 66 * n0] = x;
        }
    }
}
WGB_FFT_HD void stage3(int lane, cf* v, const cf* buf) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        cf a[8];
#pragma unroll
        for (int n0 = 0; n0 < 8; ++n0) a[n0] = buf[lane + 32 * u + 66 * n0];
        dft8(a);
#pragma unroll
        for (int k0 = 0; k0 < 8; ++k0) v[u + 2 * k0] = a[k0];
    }
}

// ---- real-FFT split.  Z = FFT512(z), z[m] = x[2m] + i x[2m+1].  With P[j] = Z[512 - k] for k = lane + 32 j (the value
// this lane receives from lane (32 - lane) & 31; for k = 0 the "partner" is Z[0] itself):
//   X[k] = (Z[k] + conj P)/2 - (i/2) e^{-2 pi i k/1024} (Z[k] - conj P),    X[512] = Re Z[0] - Im Z[0]   (real)
WGB_FFT_HD cf rfft_bin(cf zk, cf partner, cf tw) {
    const cf pc = cconj(partner);
    const cf e = cf{0.5f * (zk.x + pc.x), 0.5f * (zk.y + pc.y)};
    const cf d = cf{0.5f * (zk.x - pc.x), 0.5f * (zk.y - pc.y)};
    const cf o = cmul(mul_neg_i(d), tw);                    // -i d e^{-2 pi i k/1024}
    return cadd(e, o);
}
// and back: Z[k] = (X[k] + conj Q)/2 + (i/2) e^{+2 pi i k/1024} (X[k] - conj Q) with Q = X[512 - k]; feeding conj(Z) to the
// forward FFT gives conj(512 z): the caller reads x[2m] = Re, x[2m+1] = -Im and scales by 1/512.
WGB_FFT_HD cf irfft_bin(cf xk, cf partner, cf tw) {
    const cf qc = cconj(partner);
    const cf e = cf{0.5f * (xk.x + qc.x), 0.5f * (xk.y + qc.y)};
    const cf d = cf{0.5f * (xk.x - qc.x), 0.5f * (xk.y - qc.y)};
    const cf id = cf{-d.y, d.x};                            // i d
    return cadd(e, cmul(id, cconj(tw)));
}

}  // namespace fft
}  // namespace wgb
