// CPU emulation of the 16-thread FFT group from ctucopy_b200/csrc/ctu_fft.cuh: the same
// __host__ __device__ code, run "thread by thread" with plain arrays standing in for shared
// memory.  Checks the index algebra of the four-step 256-point FFT, the real split and
// the inverse against a naive O(n^2) DFT.  Test infrastructure only.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../ctucopy_b200/csrc/ctu_fft.cuh"
using namespace ctu;

template <class T> int run(double tol) {
    const double PI = 3.14159265358979323846;
    std::vector<cpx<T>> tw256(256), twsplit(129), twinv(129);
    for (int k1 = 0; k1 < 16; k1++) for (int c = 0; c < 16; c++) {
        double a = -2 * PI * (c * k1) / 256.0; tw256[k1 * 16 + c] = mk<T>((T)cos(a), (T)sin(a)); }
    for (int k = 0; k <= 128; k++) {
        double th = 2 * PI * k / 512.0;
        twsplit[k] = mk<T>((T)(-sin(th) / 2), (T)(-cos(th) / 2));
        twinv[k] = mk<T>((T)cos(th), (T)sin(th));
    }
    std::vector<double> x(512, 0.0);
    srand(3);
    for (int i = 0; i < 400; i++) x[i] = (rand() / (double)RAND_MAX - 0.5) * 2000;
    // thread-private registers
    static cpx<T> reg[16][16]; static cpx<T> lo[16][8], hi[16][8], mid[16];
    std::vector<cpx<T>> xch(16 * XPAD), zlin(256);
    for (int c = 0; c < 16; c++) for (int n1 = 0; n1 < 16; n1++) { int n = 16 * n1 + c; reg[c][n1] = mk<T>((T)x[2 * n], (T)x[2 * n + 1]); }
    for (int c = 0; c < 16; c++) fft256_pass1(reg[c], c, tw256.data(), xch.data());
    for (int c = 0; c < 16; c++) fft256_pass2(reg[c], c, xch.data());
    for (int c = 0; c < 16; c++) fft256_store_linear(reg[c], c, zlin.data());
    for (int c = 0; c < 16; c++) rfft_split(zlin.data(), c, twsplit.data(), lo[c], hi[c], mid[c]);
    std::vector<double> re(257), im(257);
    for (int k = 0; k <= 256; k++) { re[k] = im[k] = 0; for (int n = 0; n < 512; n++) { double a = -2 * PI * ((n * k) % 512) / 512.0; re[k] += x[n] * cos(a); im[k] += x[n] * sin(a); } }
    double emax = 0, ref = 0;
    for (int k = 0; k <= 256; k++) ref = fmax(ref, hypot(re[k], im[k]));
    for (int c = 0; c < 16; c++) for (int j = 0; j < 8; j++) {
        int k = c + 16 * j;
        emax = fmax(emax, hypot(lo[c][j].x - re[k], lo[c][j].y - im[k]));
        emax = fmax(emax, hypot(hi[c][j].x - re[256 - k], hi[c][j].y - im[256 - k]));
    }
    emax = fmax(emax, hypot(mid[0].x - re[128], mid[0].y - im[128]));
    printf("forward: max err %.3g (rel %.3g)\n", emax, emax / ref);
    int bad = (emax / ref > tol);
    // inverse
    for (int c = 0; c < 16; c++) irfft_presplit(zlin.data(), c, twinv.data(), lo[c], hi[c], mid[c]);
    for (int c = 0; c < 16; c++) fft256_load_column(reg[c], c, zlin.data());
    for (int c = 0; c < 16; c++) fft256_pass1(reg[c], c, tw256.data(), xch.data());
    for (int c = 0; c < 16; c++) fft256_pass2(reg[c], c, xch.data());
    double e2 = 0;
    for (int c = 0; c < 16; c++) for (int k2 = 0; k2 < 16; k2++) {
        int n = c + 16 * k2;
        e2 = fmax(e2, fabs(reg[c][k2].x / 512.0 - x[2 * n]));
        e2 = fmax(e2, fabs(-reg[c][k2].y / 512.0 - x[2 * n + 1]));
    }
    printf("inverse: max err %.3g (rel %.3g)\n", e2, e2 / 1000.0);
    bad |= (e2 / 1000.0 > tol);
    // ---- the variants the kernels use now: on-the-fly twiddles in pass 1, and the split / inverse pre-split
    // with the partner thread's registers instead of shared memory (shuffles are emulated by reading the
    // partner's registers directly)
    static cpx<T> r2[16][16], lo2[16][8], hi2[16][8], mid2[16];
    for (int c = 0; c < 16; c++) for (int n1 = 0; n1 < 16; n1++) { int n = 16 * n1 + c; r2[c][n1] = mk<T>((T)x[2 * n], (T)x[2 * n + 1]); }
    for (int c = 0; c < 16; c++) fft256_pass1_rec(r2[c], c, tw256.data(), xch.data());
    for (int c = 0; c < 16; c++) fft256_pass2(r2[c], c, xch.data());
    for (int c = 0; c < 16; c++) {
        cpx<T> Zp[8];
        for (int j = 0; j < 8; j++) Zp[j] = r2[(16 - c) & 15][15 - j];
        rfft_split_pairs(r2[c], Zp, c, twsplit.data(), lo2[c], hi2[c], mid2[c]);
    }
    double e3 = 0;
    for (int c = 0; c < 16; c++) for (int j = 0; j < 8; j++) {
        int k = c + 16 * j;
        e3 = fmax(e3, hypot(lo2[c][j].x - re[k], lo2[c][j].y - im[k]));
        e3 = fmax(e3, hypot(hi2[c][j].x - re[256 - k], hi2[c][j].y - im[256 - k]));
    }
    e3 = fmax(e3, hypot(mid2[0].x - re[128], mid2[0].y - im[128]));
    printf("forward (rec twiddles + register split): max err %.3g (rel %.3g)\n", e3, e3 / ref);
    bad |= (e3 / ref > 2 * tol);
    static cpx<T> zn[16][8], z128[16];
    for (int c = 0; c < 16; c++) irfft_presplit_local(r2[c], zn[c], z128[c], c, twinv.data(), lo2[c], hi2[c], mid2[c]);
    for (int c = 0; c < 16; c++) irfft_presplit_place(r2[c], c, zn[c], zn[(16 - c) & 15], z128[c]);
    for (int c = 0; c < 16; c++) fft256_pass1_rec(r2[c], c, tw256.data(), xch.data());
    for (int c = 0; c < 16; c++) fft256_pass2(r2[c], c, xch.data());
    double e4 = 0;
    for (int c = 0; c < 16; c++) for (int k2 = 0; k2 < 16; k2++) {
        int n = c + 16 * k2;
        e4 = fmax(e4, fabs(r2[c][k2].x / 512.0 - x[2 * n]));
        e4 = fmax(e4, fabs(-r2[c][k2].y / 512.0 - x[2 * n + 1]));
    }
    printf("inverse (register pre-split): max err %.3g (rel %.3g)\n", e4, e4 / 1000.0);
    bad |= (e4 / 1000.0 > 2 * tol);
    // ---- the persistent PCM -> spectrum kernel's form: pass-1 twiddles handed in as six thread constants (must be
    // bit-identical to fft256_pass1_rec) and the split twiddles formed from twsplit[c] times compile-time constants
    static cpx<T> r3[16][16], lo3[16][8], hi3[16][8], mid3[16];
    std::vector<cpx<T>> xch3(16 * XPAD);
    for (int c = 0; c < 16; c++) for (int n1 = 0; n1 < 16; n1++) { int n = 16 * n1 + c; r3[c][n1] = mk<T>((T)x[2 * n], (T)x[2 * n + 1]); r2[c][n1] = r3[c][n1]; }
    for (int c = 0; c < 16; c++) {
        cpx<T> tw[6] = {tw256[16 + c], tw256[32 + c], tw256[48 + c], tw256[64 + c], tw256[128 + c], tw256[192 + c]};
        fft256_pass1_reg(r3[c], c, tw, xch3.data());
        fft256_pass1_rec(r2[c], c, tw256.data(), xch.data());
    }
    for (int i = 0; i < 16 * XPAD; i++) if (i % XPAD < 16) bad |= (xch[i].x != xch3[i].x || xch[i].y != xch3[i].y);
    for (int c = 0; c < 16; c++) fft256_pass2(r3[c], c, xch3.data());
    for (int c = 0; c < 16; c++) {
        cpx<T> Zp[8];
        for (int j = 0; j < 8; j++) Zp[j] = r3[(16 - c) & 15][15 - j];
        rfft_split_pairs_rec(r3[c], Zp, c, twsplit[c], lo3[c], hi3[c], mid3[c]);
    }
    double e5 = 0;
    for (int c = 0; c < 16; c++) for (int j = 0; j < 8; j++) {
        int k = c + 16 * j;
        e5 = fmax(e5, hypot(lo3[c][j].x - re[k], lo3[c][j].y - im[k]));
        e5 = fmax(e5, hypot(hi3[c][j].x - re[256 - k], hi3[c][j].y - im[256 - k]));
    }
    e5 = fmax(e5, hypot(mid3[0].x - re[128], mid3[0].y - im[128]));
    printf("forward (register twiddles + composed split twiddles): max err %.3g (rel %.3g)\n", e5, e5 / ref);
    bad |= (e5 / ref > 2 * tol);
    return bad;
}
// the 8-thread group of the 256-point front end (ctu_frames256.cuh): 128 complex points = 16 x 8, thread g holds z[8 n1 + g]
template <class T> int run256(double tol) {
    const double PI = 3.14159265358979323846;
    const int XP = 9;
    std::vector<cpx<T>> tw128(128), twsplit(129);
    for (int k1 = 0; k1 < 16; k1++) for (int g = 0; g < 8; g++) {
        double a = -2 * PI * (g * k1) / 128.0; tw128[k1 * 8 + g] = mk<T>((T)cos(a), (T)sin(a)); }
    for (int k = 0; k <= 128; k++) { double th = 2 * PI * k / 256.0; twsplit[k] = mk<T>((T)(-sin(th) / 2), (T)(-cos(th) / 2)); }
    std::vector<double> x(256, 0.0);
    srand(5);
    for (int i = 0; i < 200; i++) x[i] = (rand() / (double)RAND_MAX - 0.5) * 2000;
    static cpx<T> reg[8][16], b0[8][8], b1[8][8];
    std::vector<cpx<T>> xch(16 * XP + 8);
    for (int g = 0; g < 8; g++) for (int n1 = 0; n1 < 16; n1++) { int n = 8 * n1 + g; reg[g][n1] = mk<T>((T)x[2 * n], (T)x[2 * n + 1]); }
    for (int g = 0; g < 8; g++) fft128_pass1(reg[g], g, tw128.data(), xch.data(), XP);
    for (int g = 0; g < 8; g++) fft128_pass2(b0[g], b1[g], g, xch.data(), XP);
    for (int g = 0; g < 8; g++) for (int k2 = 0; k2 < 8; k2++) { xch[g + 16 * k2] = b0[g][k2]; xch[g + 8 + 16 * k2] = b1[g][k2]; }
    double emax = 0, ref = 0;
    for (int k = 0; k <= 128; k++) {
        double re = 0, im = 0;
        for (int n = 0; n < 256; n++) { double a = -2 * PI * ((n * k) % 256) / 256.0; re += x[n] * cos(a); im += x[n] * sin(a); }
        ref = fmax(ref, hypot(re, im));
        const cpx<T> X = rfft256_bin(xch.data(), twsplit.data(), k);
        emax = fmax(emax, hypot(X.x - re, X.y - im));
    }
    printf("256-point forward (8 threads x 16 points): max err %.3g (rel %.3g)\n", emax, emax / ref);
    return emax / ref > tol;
}
int main() {
    int bad = run<float>(2e-6) | run<double>(1e-14) | run256<float>(2e-6) | run256<double>(1e-14);
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
