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
template <class T> CTU_HD cpx<T> operator+(cpx<T> a, cpx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <class T> CTU_HD cpx<T> operator-(cpx<T> a, cpx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <class T> CTU_HD cpx<T> cmul(cpx<T> a, cpx<T> b) { return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <class T> CTU_HD cpx<T> conj(cpx<T> a) { return mk<T>(a.x, -a.y); }
template <class T> CTU_HD cpx<T> mul_mi(cpx<T> a) { return mk<T>(a.y, -a.x); }   // a * (-i)

// forward 4-point DFT, natural order in and out
template <class T> CTU_HD void dft4(cpx<T> &a0, cpx<T> &a1, cpx<T> &a2, cpx<T> &a3) {
    cpx<T> s0 = a0 + a2, d0 = a0 - a2, s1 = a1 + a3, d1 = mul_mi(a1 - a3);
    a0 = s0 + s1; a2 = s0 - s1; a1 = d0 + d1; a3 = d0 - d1;
}

// forward 16-point DFT in registers (4 x 4), natural order in and out
template <class T> CTU_HD void dft16(cpx<T> (&a)[16]) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, r2 = (T)0.70710678118654752440;
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);
    // a[4*k1 + n2] *= W16^(n2*k1)
    a[5] = cmul(a[5], mk<T>(c1, -s1));                       // 1
    a[6] = mk<T>((a[6].x + a[6].y) * r2, (a[6].y - a[6].x) * r2);   // 2: (r2,-r2)
    a[7] = cmul(a[7], mk<T>(s1, -c1));                       // 3
    a[9] = mk<T>((a[9].x + a[9].y) * r2, (a[9].y - a[9].x) * r2);   // 2
    a[10] = mul_mi(a[10]);                                   // 4
    a[11] = mk<T>((a[11].y - a[11].x) * r2, -(a[11].x + a[11].y) * r2);  // 6: (-r2,-r2)
    a[13] = cmul(a[13], mk<T>(s1, -c1));                     // 3
    a[14] = mk<T>((a[14].y - a[14].x) * r2, -(a[14].x + a[14].y) * r2);  // 6
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
    const T r2 = (T)0.70710678118654752440;
    cpx<T> e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    dft4(e0, e1, e2, e3);
    dft4(o0, o1, o2, o3);
    o1 = mk<T>((o1.x + o1.y) * r2, (o1.y - o1.x) * r2);              // W8^1 = (r2, -r2)
    o2 = mul_mi(o2);                                                 // W8^2 = -i
    o3 = mk<T>((o3.y - o3.x) * r2, -(o3.x + o3.y) * r2);             // W8^3 = (-r2, -r2)
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
    return mk<T>((T)0.5 * (A.x + B.x), (T)0.5 * (A.y + B.y)) + cmul(twsplit[k], A - B);
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
            cpx<T> E = mk<T>((T)0.5 * (A.x + B.x), (T)0.5 * (A.y + B.y));
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
            const cpx<T> E = mk<T>((T)0.5 * (A.x + B.x), (T)0.5 * (A.y + B.y));
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
            const cpx<T> E = mk<T>((T)0.5 * (A.x + B.x), (T)0.5 * (A.y + B.y));
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
            a[j] = conj(mk<T>(S.x - U.y, S.y + U.x));           // conj(Zc[k])
            zn[j] = conj(mk<T>(S.x + U.y, -S.y + U.x));         // conj(Zc[256-k])
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
            cpx<T> zk = mk<T>(S.x - U.y, S.y + U.x);
            cpx<T> zn = mk<T>(S.x + U.y, -S.y + U.x);
            zlin[k] = conj(zk);
            zlin[NC - k] = conj(zn);
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
