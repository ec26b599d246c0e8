// Host check of text2speech_b200/csrc/fft_core.cuh: the same stage functions the CUDA kernels run, with a warp emulated as
// a loop over 32 lanes between the barriers, against a naive double-precision DFT.   g++ -O1 -I text2speech_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "fft_core.cuh"

using namespace wgb::fft;

static void fft512_warp(cf v[32][16]) {
    static cf buf_a[kBufA], buf_b[kBufB];
    LaneTw tw[32];
    for (int l = 0; l < 32; ++l) lane_twiddles(l, tw[l]);
    for (int i = 0; i < kBufA; ++i) buf_a[i] = cf{NAN, NAN};        // every element a later stage reads must be written
    for (int i = 0; i < kBufB; ++i) buf_b[i] = cf{NAN, NAN};
    for (int l = 0; l < 32; ++l) stage1(l, tw[l], v[l], buf_a);
    for (int l = 0; l < 32; ++l) stage2(l, tw[l], buf_a, buf_b);
    for (int l = 0; l < 32; ++l) stage3(l, v[l], buf_b);
}

int main() {
    const double PI = 3.14159265358979323846;
    srand(7);
    int fails = 0;
    // ---- 512-point complex FFT
    std::vector<double> zr(512), zi(512);
    cf v[32][16];
    for (int n = 0; n < 512; ++n) {
        zr[n] = rand() / (double)RAND_MAX - 0.5;
        zi[n] = rand() / (double)RAND_MAX - 0.5;
        v[n & 31][n >> 5] = cf{(float)zr[n], (float)zi[n]};
    }
    fft512_warp(v);
    double worst = 0, norm = 0;
    for (int k = 0; k < 512; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < 512; ++n) {
            const double a = -2 * PI * ((long long)n * k % 512) / 512.0;
            sr += zr[n] * cos(a) - zi[n] * sin(a);
            si += zr[n] * sin(a) + zi[n] * cos(a);
        }
        const cf g = v[k & 31][k >> 5];
        worst = fmax(worst, hypot(g.x - sr, g.y - si));
        norm = fmax(norm, hypot(sr, si));
    }
    printf("fft512: max |err| = %.3e (max |Z| = %.2f)\n", worst, norm);
    if (!(worst < 2e-5 * norm)) ++fails;

    // ---- 1024-point real FFT through the split, and back
    std::vector<double> x(1024);
    for (int n = 0; n < 1024; ++n) x[n] = rand() / (double)RAND_MAX - 0.5;
    for (int m = 0; m < 512; ++m) v[m & 31][m >> 5] = cf{(float)x[2 * m], (float)x[2 * m + 1]};
    fft512_warp(v);
    cf X[32][16];
    float x_nyq = 0.f;
    LaneTw tw[32];
    for (int l = 0; l < 32; ++l) lane_twiddles(l, tw[l]);
    for (int l = 0; l < 32; ++l)
        for (int j = 0; j < 16; ++j) {
            const int pl = (32 - l) & 31;
            const cf partner = l == 0 ? v[0][(16 - j) & 15] : v[pl][15 - j];
            X[l][j] = rfft_bin(v[l][j], partner, cmul(tw[l].post, w32(j)));
        }
    x_nyq = v[0][0].x - v[0][0].y;
    worst = 0; norm = 0;
    for (int k = 0; k <= 512; ++k) {
        double sr = 0, si = 0;
        for (int n = 0; n < 1024; ++n) {
            const double a = -2 * PI * ((long long)n * k % 1024) / 1024.0;
            sr += x[n] * cos(a);
            si += x[n] * sin(a);
        }
        const cf g = k < 512 ? X[k & 31][k >> 5] : cf{x_nyq, 0.f};
        worst = fmax(worst, hypot(g.x - sr, g.y - si));
        norm = fmax(norm, hypot(sr, si));
    }
    printf("rfft1024: max |err| = %.3e (max |X| = %.2f)\n", worst, norm);
    if (!(worst < 2e-5 * norm)) ++fails;

    cf Z[32][16];
    for (int l = 0; l < 32; ++l)
        for (int j = 0; j < 16; ++j) {
            const int pl = (32 - l) & 31;
            cf partner = l == 0 ? X[0][(16 - j) & 15] : X[pl][15 - j];
            if (l == 0 && j == 0) partner = cf{x_nyq, 0.f};
            Z[l][j] = cconj(irfft_bin(X[l][j], partner, cmul(tw[l].post, w32(j))));
        }
    fft512_warp(Z);
    worst = 0;
    for (int m = 0; m < 512; ++m) {
        const cf g = Z[m & 31][m >> 5];
        worst = fmax(worst, fabs(g.x / 512.0 - x[2 * m]));
        worst = fmax(worst, fabs(-g.y / 512.0 - x[2 * m + 1]));
    }
    printf("irfft(rfft(x)): max |err| = %.3e\n", worst);
    if (!(worst < 1e-6)) ++fails;
    printf(fails ? "FAIL\n" : "OK\n");
    return fails;
}
