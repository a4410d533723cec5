// 512-point real FFT as a 256-point complex FFT (16 x 16 four-step) carried by a GROUP of
// 16 threads, each holding 16 complex points in registers.  One shared-memory transpose
// between the two radix-16 passes, one for the real-input split.  The same code is used
// for float (feature path) and double (Burg detector path) and for the inverse transform
// (conjugate trick).  Everything here is __host__ __device__ so tests/emu can run the
// identical index algebra on the CPU, thread by thread.
//
// Conventions (match FFTW's R2HC / HC2R as the reference uses them, src/io/in.cc:229,388,
// src/io/out.cc:391,425): forward X[k] = sum_n x[n] e^{-2 pi i nk/512}; inverse
// unnormalised.
#ifndef CTU_FFT_CUH
#define CTU_FFT_CUH

#if defined(__CUDACC__)
#define CTU_HD __host__ __device__ __forceinline__
#else
#define CTU_HD inline
#endif

namespace ctu {

constexpr int NFFT = 512;   // real transform length handled by the fast path
constexpr int NC = 256;     // complex length
constexpr int NBIN = 257;   // NFFT/2 + 1
constexpr int SPITCH = 260; // floats per spectrum row in HBM (512-point path): 16-byte aligned rows, three zero pad columns
constexpr int GROUP = 16;   // threads per frame
constexpr int XPAD = 17;    // row pitch (in complex elements) of the 16x16 exchange tile

template <class T> struct cpx { T x, y; };

template <class T> CTU_HD cpx<T> mk(T x, T y) { cpx<T> r; r.x = x; r.y = y; return r; }

// ---- Blackwell packed FP32 (PTX add/sub/mul/fma.rn.f32x2, sm_100+) -----------------------------------------------------
// A complex float is a register pair, and FADD2 / FMUL2 / FFMA2 work on such pairs in ONE issue slot, with the swap of the
// halves (.LO_HI), the sign of one half (.NP) and the broadcast of a scalar (.F32) as operand modifiers: a complex add is
// one instruction instead of two, a complex multiply two instead of four, and the multiplication by -i in front of an
// add costs nothing (cuobjdump: `FADD2 R16, R6.F32x2.HI_LO, R10.F32x2.LO_HI.NP`).  The FFT kernels are bound by issue
// slots (ncu, profiles/r02_ncu_mfcc_exten.txt: 70 % of the slots, 47 % of the FP32 pipe), and their arithmetic is
// complex arithmetic.  Every lane is rounded exactly like the scalar instruction (IEEE round-to-nearest, no flush), and
// the multiply keeps the contraction nvcc chooses for the scalar form: x = fma(ax, bx, -(ay by)), y = fma(ay, bx, ax by).
// -DCTU_NO_F32X2 compiles the scalar form (A/B runs; the host build for tests/emu always takes it).
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000) && !defined(CTU_NO_F32X2)
#define CTU_F32X2 1
__device__ __forceinline__ unsigned long long f2pk(float x, float y) { unsigned long long u; asm("mov.b64 %0, {%1, %2};" : "=l"(u) : "f"(x), "f"(y)); return u; }
__device__ __forceinline__ cpx<float> f2up(unsigned long long u) { cpx<float> c; asm("mov.b64 {%0, %1}, %2;" : "=f"(c.x), "=f"(c.y) : "l"(u)); return c; }
__device__ __forceinline__ cpx<float> f2add(cpx<float> a, cpx<float> b) {
    unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2pk(a.x, a.y)), "l"(f2pk(b.x, b.y))); return f2up(r);
}
__device__ __forceinline__ cpx<float> f2sub(cpx<float> a, cpx<float> b) {
    unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2pk(a.x, a.y)), "l"(f2pk(b.x, b.y))); return f2up(r);
}
__device__ __forceinline__ cpx<float> f2mul(cpx<float> a, cpx<float> b) {
    unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2pk(a.x, a.y)), "l"(f2pk(b.x, b.y))); return f2up(r);
}
__device__ __forceinline__ cpx<float> f2fma(cpx<float> a, cpx<float> b, cpx<float> d) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2pk(a.x, a.y)), "l"(f2pk(b.x, b.y)), "l"(f2pk(d.x, d.y)));
    return f2up(r);
}
#else
#define CTU_F32X2 0
#endif

template <class T> CTU_HD cpx<T> operator+(cpx<T> a, cpx<T> b) {
#if CTU_F32X2
    if constexpr (sizeof(T) == 4) return f2add(a, b);
#endif
    return mk<T>(a.x + b.x, a.y + b.y);
}
template <class T> CTU_HD cpx<T> operator-(cpx<T> a, cpx<T> b) {
#if CTU_F32X2
    if constexpr (sizeof(T) == 4) return f2sub(a, b);
#endif
    return mk<T>(a.x - b.x, a.y - b.y);
}
template <class T> CTU_HD cpx<T> cmul(cpx<T> a, cpx<T> b) {
#if CTU_F32X2
    if constexpr (sizeof(T) == 4) {
        const cpx<T> t = f2mul(mk<T>(a.y, a.x), mk<T>(b.y, b.y));          // (ay by, ax by)
        return f2fma(a, mk<T>(b.x, b.x), mk<T>(-t.x, t.y));
    }
#endif
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * s, s real
template <class T> CTU_HD cpx<T> cscale(cpx<T> a, T s) {
#if CTU_F32X2
    if constexpr (sizeof(T) == 4) return f2mul(a, mk<T>(s, s));
#endif
    return mk<T>(a.x * s, a.y * s);
}
// element by element: (ax bx, ay by) -- two real samples times their two window values
template <class T> CTU_HD cpx<T> pmul(cpx<T> a, cpx<T> b) {
#if CTU_F32X2
    if constexpr (sizeof(T) == 4) return f2mul(a, b);
#endif
    return mk<T>(a.x * b.x, a.y * b.y);
}
// element by element: (fma(ax, bx, dx), fma(ay, by, dy)) -- two taps of a dot product whose partial sums are a register pair
template <class T> CTU_HD cpx<T> pfma(cpx<T> a, cpx<T> b, cpx<T> d) {
#if CTU_F32X2
    if constexpr (sizeof(T) == 4) return f2fma(a, b, d);
#endif
#if defined(__CUDA_ARCH__)
    return mk<T>(fma(a.x, b.x, d.x), fma(a.y, b.y, d.y));
#else
    return mk<T>(a.x * b.x + d.x, a.y * b.y + d.y);
#endif
}
template <class T> CTU_HD cpx<T> conj(cpx<T> a) { return mk<T>(a.x, -a.y); }
template <class T> CTU_HD cpx<T> mul_mi(cpx<T> a) { return mk<T>(a.y, -a.x); }   // a * (-i)
// a * (1 - i) r and a * (-1 - i) r, r = sqrt(1/2): W8^1 and W8^3.  ((ax + ay) r, (ay - ax) r) and ((ay - ax) r, -(ax + ay) r)
template <class T> CTU_HD cpx<T> mul_w8_1(cpx<T> a) { return cscale(a + mul_mi(a), (T)0.70710678118654752440); }
template <class T> CTU_HD cpx<T> mul_w8_3(cpx<T> a) { return cscale(mul_mi(a) - a, (T)0.70710678118654752440); }

// forward 4-point DFT, natural order in and out
template <class T> CTU_HD void dft4(cpx<T> &a0, cpx<T> &a1, cpx<T> &a2, cpx<T> &a3) {
    cpx<T> s0 = a0 + a2, d0 = a0 - a2, s1 = a1 + a3, d1 = mul_mi(a1 - a3);
    a0 = s0 + s1; a2 = s0 - s1; a1 = d0 + d1; a3 = d0 - d1;
}

// forward 16-point DFT in registers (4 x 4), natural order in and out
template <class T> CTU_HD void dft16(cpx<T> (&a)[16]) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);
    // a[4*k1 + n2] *= W16^(n2*k1)
    a[5] = cmul(a[5], mk<T>(c1, -s1));                       // 1
    a[6] = mul_w8_1(a[6]);                                    // 2: (r2,-r2)
    a[7] = cmul(a[7], mk<T>(s1, -c1));                       // 3
    a[9] = mul_w8_1(a[9]);                                    // 2
    a[10] = mul_mi(a[10]);                                   // 4
    a[11] = mul_w8_3(a[11]);                                  // 6: (-r2,-r2)
    a[13] = cmul(a[13], mk<T>(s1, -c1));                     // 3
    a[14] = mul_w8_3(a[14]);                                  // 6
    a[15] = cmul(a[15], mk<T>(-c1, s1));                     // 9
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
    // a[4*k1 + k2] now holds X[k1 + 4*k2]: transpose the 4x4 register tile
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = i + 1; j < 4; j++) { cpx<T> t = a[4 * i + j]; a[4 * i + j] = a[4 * j + i]; a[4 * j + i] = t; }
}

// ---- 256-point complex FFT over a group of 16 threads ----------------------------------
// pass 1: thread c holds z[16*n1 + c], n1 = 0..15.  Transforms over n1, applies the
//         inter-pass twiddle W256^(c*k1) and stores A[k1][c] into the exchange tile.
// tw256 : [16][16] with tw256[k1*16 + c] = W256^(c*k1)
template <class T>
CTU_HD void fft256_pass1(cpx<T> (&a)[16], int c, const cpx<T> *tw256, cpx<T> *xch) {
    dft16(a);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
        cpx<T> v = a[k1];
        if (k1 > 0) v = cmul(v, tw256[k1 * 16 + c]);
        xch[k1 * XPAD + c] = v;
    }
}
// pass 1 with most inter-pass twiddles formed on the fly: only W^c, W^2c, W^3c, W^4c, W^8c and
// W^12c (W = e^{-2 pi i/256}) are loaded, the other nine are products of two of them.  Trades
// nine shared-memory loads per thread for 36 flops; the shared-memory pipe is the scarcer one.
template <class T>
CTU_HD void fft256_pass1_rec(cpx<T> (&a)[16], int c, const cpx<T> *tw256, cpx<T> *xch) {
    dft16(a);
    const cpx<T> w1 = tw256[16 + c], w2 = tw256[32 + c], w3 = tw256[48 + c];
    xch[c] = a[0];
    xch[1 * XPAD + c] = cmul(a[1], w1);
    xch[2 * XPAD + c] = cmul(a[2], w2);
    xch[3 * XPAD + c] = cmul(a[3], w3);
#pragma unroll
    for (int q = 1; q < 4; q++) {
        const cpx<T> wq = tw256[64 * q + c];            // W^(4q c)
        xch[(4 * q) * XPAD + c] = cmul(a[4 * q], wq);
        xch[(4 * q + 1) * XPAD + c] = cmul(a[4 * q + 1], cmul(wq, w1));
        xch[(4 * q + 2) * XPAD + c] = cmul(a[4 * q + 2], cmul(wq, w2));
        xch[(4 * q + 3) * XPAD + c] = cmul(a[4 * q + 3], cmul(wq, w3));
    }
}
// the same with the six loaded twiddles handed in: they depend on the thread (c) only, so a persistent kernel loads them
// once and keeps them in registers.  tw[0..2] = W^c, W^2c, W^3c; tw[3..5] = W^4c, W^8c, W^12c.  Bit-identical to
// fft256_pass1_rec.
template <class T>
CTU_HD void fft256_pass1_reg(cpx<T> (&a)[16], int c, const cpx<T> (&tw)[6], cpx<T> *xch) {
    dft16(a);
    xch[c] = a[0];
    xch[1 * XPAD + c] = cmul(a[1], tw[0]);
    xch[2 * XPAD + c] = cmul(a[2], tw[1]);
    xch[3 * XPAD + c] = cmul(a[3], tw[2]);
#pragma unroll
    for (int q = 1; q < 4; q++) {
        const cpx<T> wq = tw[2 + q];
        xch[(4 * q) * XPAD + c] = cmul(a[4 * q], wq);
        xch[(4 * q + 1) * XPAD + c] = cmul(a[4 * q + 1], cmul(wq, tw[0]));
        xch[(4 * q + 2) * XPAD + c] = cmul(a[4 * q + 2], cmul(wq, tw[1]));
        xch[(4 * q + 3) * XPAD + c] = cmul(a[4 * q + 3], cmul(wq, tw[2]));
    }
}
// pass 2 (after a group-wide sync): thread c now plays k1 = c; loads A[c][n2], transforms
// over n2; on return a[k2] = Z[c + 16*k2].
template <class T>
CTU_HD void fft256_pass2(cpx<T> (&a)[16], int c, const cpx<T> *xch) {
#pragma unroll
    for (int n2 = 0; n2 < 16; n2++) a[n2] = xch[c * XPAD + n2];
    dft16(a);
}

// forward 8-point DFT in registers, natural order in and out
template <class T> CTU_HD void dft8(cpx<T> (&a)[8]) {
    cpx<T> e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    o1 = mul_w8_1(o1);                                               // W8^1 = (r2, -r2)
    o2 = mul_mi(o2);                                                 // W8^2 = -i
    o3 = mul_w8_3(o3);                                               // W8^3 = (-r2, -r2)
    a[0] = e0 + o0; a[4] = e0 - o0;
    a[1] = e1 + o1; a[5] = e1 - o1;
    a[2] = e2 + o2; a[6] = e2 - o2;
    a[3] = e3 + o3; a[7] = e3 - o3;
}

// ---- 128-point complex FFT over a group of 8 threads (256-point real frames, ctu_frames256.cuh) -------------------
// 128 = 16 x 8: thread g holds z[8 n1 + g], n1 = 0..15.
// pass 1: 16-point DFT over n1, times W128^(g k1) (tw128[k1 * 8 + g]), to the exchange tile at [k1][g] (row pitch xp)
template <class T>
CTU_HD void fft128_pass1(cpx<T> (&a)[16], int g, const cpx<T> *tw128, cpx<T> *xch, int xp) {
    dft16(a);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) xch[k1 * xp + g] = (k1 > 0 && g > 0) ? cmul(a[k1], tw128[k1 * 8 + g]) : a[k1];
}
// pass 2 (after a group-wide sync): thread g plays k1 = g and k1 = g + 8: two 8-point DFTs over the 8 threads' values; on
// return b0[k2] = Z[g + 16 k2], b1[k2] = Z[g + 8 + 16 k2]
template <class T>
CTU_HD void fft128_pass2(cpx<T> (&b0)[8], cpx<T> (&b1)[8], int g, const cpx<T> *xch, int xp) {
#pragma unroll
    for (int n2 = 0; n2 < 8; n2++) { b0[n2] = xch[g * xp + n2]; b1[n2] = xch[(g + 8) * xp + n2]; }
    dft8(b0);
    dft8(b1);
}
// bin k (0..128) of the 256-point real transform from the linear Z[0..127]; twsplit[k] = -i/2 e^{-2 pi i k / 256}
template <class T>
CTU_HD cpx<T> rfft256_bin(const cpx<T> *zlin, const cpx<T> *twsplit, int k) {
    const cpx<T> A = zlin[k == 128 ? 0 : k], B = conj(zlin[k == 0 ? 0 : 128 - k]);
    return cscale(A + B, (T)0.5) + cmul(twsplit[k], A - B);
}

// ---- real-input split ---------------------------------------------------------------------
// After pass 2 every thread stores Z linearly (zlin[c + 16*k2] = a[k2]); after a sync,
// thread c owns the bin pairs (k, 256-k) for k = c + 16*j, j = 0..7.
// twsplit[k] = -i/2 * e^{-2 pi i k/512}, k = 0..128
// Returns X[k] in lo[j] and X[256-k] in hi[j].  For k == 0: lo = X[0], hi = X[256] (both
// real).  For k == 128 (c == 0, j == 8 does not exist; k = 128 arises for c == 0 only via
// the dedicated slot `mid`): mid = X[128].
template <class T>
CTU_HD void rfft_split(const cpx<T> *zlin, int c, const cpx<T> *twsplit, cpx<T> (&lo)[8], cpx<T> (&hi)[8], cpx<T> &mid) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int k = c + 16 * j;
        cpx<T> A = zlin[k];
        if (k == 0) {
            lo[j] = mk<T>(A.x + A.y, (T)0);
            hi[j] = mk<T>(A.x - A.y, (T)0);
        } else {
            cpx<T> B = conj(zlin[NC - k]);
            cpx<T> E = cscale(A + B, (T)0.5);
            cpx<T> Tt = cmul(twsplit[k], A - B);
            lo[j] = E + Tt;
            hi[j] = conj(E - Tt);
        }
    }
    mid = conj(zlin[128]);   // only meaningful for c == 0
}

// ---- real-input split and its inverse WITHOUT shared memory ---------------------------------
// After pass 2 thread c holds Z[c + 16 k2] in a[k2].  X[k] for k = c + 16 j (j < 8) needs Z[k] = a[j]
// and Z[256-k], which thread (16-c)%16 holds at index 15-j (thread 0 pairs with itself at index
// 16-j).  The kernels fetch Zp[j] = partner's a[15-j] with register shuffles (ctu_kernels.cuh);
// the arithmetic lives here so that tests/emu can run it thread by thread.  Same maths as rfft_split.
template <class T>
CTU_HD void rfft_split_pairs(const cpx<T> (&a)[16], const cpx<T> (&Zp)[8], int c, const cpx<T> *twsplit, cpx<T> (&lo)[8], cpx<T> (&hi)[8],
                             cpx<T> &mid) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        cpx<T> Z = Zp[j];
        if (c == 0 && j >= 1) Z = a[j >= 1 ? 16 - j : 0];
        const cpx<T> A = a[j];
        if (j == 0 && c == 0) {
            lo[j] = mk<T>(A.x + A.y, (T)0);
            hi[j] = mk<T>(A.x - A.y, (T)0);
        } else {
            const cpx<T> B = conj(Z);
            const cpx<T> E = cscale(A + B, (T)0.5);
            const cpx<T> Tt = cmul(twsplit[c + 16 * j], A - B);
            lo[j] = E + Tt;
            hi[j] = conj(E - Tt);
        }
    }
    mid = conj(a[8]);   // X[128], meaningful for c == 0
}

// The split twiddles of thread c are twsplit[c + 16 j] = twsplit[c] * e^{-i pi j/16}: one thread-constant value (kept in
// a register by the persistent kernels) times eight compile-time constants, instead of eight shared-memory loads per
// frame.  The product is rounded once more than the table value (<= 1.5 ulp instead of 0.5 ulp on the twiddle).
template <class T>
CTU_HD void rfft_split_pairs_rec(const cpx<T> (&a)[16], const cpx<T> (&Zp)[8], int c, cpx<T> ts_c, cpx<T> (&lo)[8], cpx<T> (&hi)[8],
                                 cpx<T> &mid) {
    // cos(pi j/16), -sin(pi j/16), j = 0..7
    const T rc[8] = {(T)1.0, (T)0.98078528040323044913, (T)0.92387953251128675613, (T)0.83146961230254523708,
                     (T)0.70710678118654752440, (T)0.55557023301960222474, (T)0.38268343236508977173, (T)0.19509032201612826785};
    const T rs[8] = {(T)0.0, (T)-0.19509032201612826785, (T)-0.38268343236508977173, (T)-0.55557023301960222474,
                     (T)-0.70710678118654752440, (T)-0.83146961230254523708, (T)-0.92387953251128675613, (T)-0.98078528040323044913};
#pragma unroll
    for (int j = 0; j < 8; j++) {
        cpx<T> Z = Zp[j];
        if (c == 0 && j >= 1) Z = a[j >= 1 ? 16 - j : 0];
        const cpx<T> A = a[j];
        if (j == 0 && c == 0) {
            lo[j] = mk<T>(A.x + A.y, (T)0);
            hi[j] = mk<T>(A.x - A.y, (T)0);
        } else {
            const cpx<T> B = conj(Z);
            const cpx<T> E = cscale(A + B, (T)0.5);
            const cpx<T> tw = (j == 0) ? ts_c : cmul(ts_c, mk<T>(rc[j], rs[j]));
            const cpx<T> Tt = cmul(tw, A - B);
            lo[j] = E + Tt;
            hi[j] = conj(E - Tt);
        }
    }
    mid = conj(a[8]);   // X[128], meaningful for c == 0
}

// inverse, step 1 (thread-local): from the half-complex bins of thread c (lo[j] = X[k], hi[j] = X[256-k],
// k = c + 16 j; mid = X[128] on thread 0) the values conj(Zc[k]) -> a[j] and conj(Zc[256-k]) -> zn[j],
// and conj(Zc[128]) -> z128.  Same maths as irfft_presplit.
template <class T>
CTU_HD void irfft_presplit_local(cpx<T> (&a)[16], cpx<T> (&zn)[8], cpx<T> &z128, int c, const cpx<T> *twinv, const cpx<T> (&lo)[8],
                                 const cpx<T> (&hi)[8], cpx<T> mid) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j == 0 && c == 0) {
            a[0] = conj(mk<T>(lo[0].x + hi[0].x, lo[0].x - hi[0].x));
            zn[0] = mk<T>((T)0, (T)0);
        } else {
            const cpx<T> S = lo[j] + conj(hi[j]);
            const cpx<T> U = cmul(lo[j] - conj(hi[j]), twinv[c + 16 * j]);
            a[j] = conj(S) - mk<T>(U.y, U.x);                   // conj(Zc[k]) = conj(S + i U) = (Sx - Uy, -Sy - Ux)
            zn[j] = S + mul_mi(U);                              // conj(Zc[256-k]) = conj(conj(S) + i conj(U)) = (Sx + Uy, Sy - Ux)
        }
    }
    // Zc[128] pairs with itself (thread 0): S = 2 Re(mid), U = 2i Im(mid) * twinv[128]
    const cpx<T> Um = cmul(mk<T>((T)0, (T)2 * mid.y), twinv[128]);
    z128 = conj(mk<T>((T)2 * mid.x - Um.y, Um.x));
}
// inverse, step 2: the column a[n1] = conj(Zc[16 n1 + c]) the first inverse pass wants.  n1 < 8 is local
// (step 1); a[15-r] is the partner's zn[r] (thread (16-c)%16), thread 0 takes its own zn[r+1] and z128.
template <class T>
CTU_HD void irfft_presplit_place(cpx<T> (&a)[16], int c, const cpx<T> (&zn)[8], const cpx<T> (&znp)[8], cpx<T> z128) {
#pragma unroll
    for (int r = 0; r < 8; r++) a[15 - r] = (c == 0) ? ((r == 7) ? z128 : zn[r < 7 ? r + 1 : 0]) : znp[r];
}

// ---- inverse: half-complex spectrum -> 512 real samples (unnormalised) ------------------
// Thread c provides X[k] (lo[j]) and X[256-k] (hi[j]) for k = c + 16*j and, for c == 0,
// X[128] in mid.  Writes conj(Zc) into zlin; after a sync run pass1 (loading column c of
// zlin), pass2, and read time samples x[2n] = Re, x[2n+1] = -Im of a[k2], n = c + 16*k2.
// twinv[k] = e^{+2 pi i k/512}, k = 0..128
template <class T>
CTU_HD void irfft_presplit(cpx<T> *zlin, int c, const cpx<T> *twinv, const cpx<T> (&lo)[8], const cpx<T> (&hi)[8], cpx<T> mid) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        int k = c + 16 * j;
        if (k == 0) {
            // X[0], X[256] real:  Zc[0] = (X0 + X256) + i (X0 - X256)
            zlin[0] = conj(mk<T>(lo[j].x + hi[j].x, lo[j].x - hi[j].x));
        } else {
            cpx<T> S = lo[j] + conj(hi[j]);
            cpx<T> U = cmul(lo[j] - conj(hi[j]), twinv[k]);
            // Zc[k] = S + i U ; Zc[256-k] = conj(S) + i conj(U)
            zlin[k] = conj(S) - mk<T>(U.y, U.x);                // conj(Zc[k])
            zlin[NC - k] = S + mul_mi(U);                       // conj(Zc[256-k])
        }
    }
    if (c == 0) {
        // k = 128 pairs with itself: Zc[128] = 2*conj(X[128]) ... derive from the general
        // form with X[k] = X[256-k] = mid:  S = mid + conj(mid), U = (mid - conj(mid)) * i
        cpx<T> S = mk<T>((T)2 * mid.x, (T)0);
        cpx<T> U = cmul(mk<T>((T)0, (T)2 * mid.y), twinv[128]);
        zlin[128] = conj(mk<T>(S.x - U.y, S.y + U.x));
    }
}

template <class T>
CTU_HD void fft256_load_column(cpx<T> (&a)[16], int c, const cpx<T> *zlin) {
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) a[n1] = zlin[16 * n1 + c];
}

template <class T>
CTU_HD void fft256_store_linear(const cpx<T> (&a)[16], int c, cpx<T> *zlin) {
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) zlin[c + 16 * k2] = a[k2];
}

}  // namespace ctu
#endif
