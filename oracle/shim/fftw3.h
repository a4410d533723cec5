/*
 * fftw3.h -- TEST INFRASTRUCTURE ONLY (oracle build), not part of the product.
 *
 * Minimal stand-in for the six FFTW3 entry points and three r2r transform kinds that
 * CtuCopy 4.0.2 uses (FFTW itself is not installed in this image and there is no
 * network). It lets the UNMODIFIED reference sources under /root/reference/src compile
 * and link (see oracle/build_ref.sh). Call sites in the reference:
 *   forward R2HC   src/io/in.cc:229,388
 *   inverse HC2R   src/io/out.cc:391,425  src/nr/nr.cc:199,291  src/vad/vad.cc:176,232
 *   DCT-II REDFT10 src/fea/fea_trap.cc:49,103
 *
 * Conventions follow the published FFTW3 manual ("The Halfcomplex-format DFT",
 * "1d Real-even DFTs"):
 *   R2HC : Y[k] = sum_j x[j] e^{-2 pi i jk/n};  out = r0, r1, ..., r_{n/2}, i_{(n+1)/2-1}, ..., i_1
 *   HC2R : unnormalised inverse of the above (input may be destroyed)
 *   REDFT10: Y[k] = 2 sum_j x[j] cos(pi (j+1/2) k / n)
 * All arithmetic is double precision. Power-of-two sizes use a radix-2 complex FFT of
 * half length with tabulated twiddles; every other size falls back to the O(n^2) sum.
 */
#ifndef CTU_ORACLE_FFTW3_SHIM_H
#define CTU_ORACLE_FFTW3_SHIM_H

#include <stdlib.h>
#include <math.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { FFTW_R2HC = 0, FFTW_HC2R = 1, FFTW_REDFT10 = 5 } fftw_r2r_kind;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

typedef struct ctu_shim_plan_s {
    int n;
    fftw_r2r_kind kind;
    double *in, *out;
    int pow2;        /* n is a power of two >= 4 */
    double *cs;      /* twiddles: cos/sin(2 pi k / n), k < n/2  (pow2) */
    int *rev;        /* bit reversal for the n/2-point complex FFT */
    double *wr, *wi; /* scratch for the half-length complex FFT */
} *fftw_plan;

static inline void *fftw_malloc(size_t n) { return malloc(n); }
static inline void fftw_free(void *p) { free(p); }

static inline fftw_plan fftw_plan_r2r_1d(int n, double *in, double *out, fftw_r2r_kind kind,
                                         unsigned flags) {
    (void)flags;
    fftw_plan p = (fftw_plan)calloc(1, sizeof(struct ctu_shim_plan_s));
    const double pi = 3.14159265358979323846264338327950288;
    p->n = n; p->kind = kind; p->in = in; p->out = out;
    p->pow2 = (n >= 4) && ((n & (n - 1)) == 0) && (kind != FFTW_REDFT10);
    if (p->pow2) {
        int h = n / 2, bits = 0;
        while ((1 << bits) < h) bits++;
        p->cs = (double *)malloc(sizeof(double) * n);
        for (int k = 0; k < h; k++) {
            p->cs[2 * k] = cos(2.0 * pi * k / n);
            p->cs[2 * k + 1] = sin(2.0 * pi * k / n);
        }
        p->rev = (int *)malloc(sizeof(int) * h);
        for (int i = 0; i < h; i++) {
            int r = 0;
            for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            p->rev[i] = r;
        }
        p->wr = (double *)malloc(sizeof(double) * h);
        p->wi = (double *)malloc(sizeof(double) * h);
    }
    return p;
}

/* in-place radix-2 DIT on bit-reversed data, h points, sign = -1 forward / +1 inverse.
 * twiddle e^{sign 2 pi i k/h} = cs[2*(2k)] + sign*i*cs[2*(2k)+1] (table is for n = 2h) */
static inline void ctu_shim_cfft(fftw_plan p, int sign) {
    int h = p->n / 2;
    double *xr = p->wr, *xi = p->wi;
    for (int len = 2; len <= h; len <<= 1) {
        int half = len >> 1, step = h / len;
        for (int i = 0; i < h; i += len) {
            for (int j = 0; j < half; j++) {
                double c = p->cs[2 * (2 * j * step)], s = sign * p->cs[2 * (2 * j * step) + 1];
                double ur = xr[i + j], ui = xi[i + j];
                double vr = xr[i + j + half] * c - xi[i + j + half] * s;
                double vi = xr[i + j + half] * s + xi[i + j + half] * c;
                xr[i + j] = ur + vr; xi[i + j] = ui + vi;
                xr[i + j + half] = ur - vr; xi[i + j + half] = ui - vi;
            }
        }
    }
}

static inline void fftw_execute(const fftw_plan p) {
    const double pi = 3.14159265358979323846264338327950288;
    int n = p->n;
    double *in = p->in, *out = p->out;
    if (p->kind == FFTW_REDFT10) {
        for (int k = 0; k < n; k++) {
            double s = 0.0;
            for (int j = 0; j < n; j++) s += in[j] * cos(pi * (j + 0.5) * k / n);
            out[k] = 2.0 * s;
        }
        return;
    }
    if (!p->pow2) {
        if (p->kind == FFTW_R2HC) {
            for (int k = 0; k <= n / 2; k++) {
                double re = 0, im = 0;
                for (int j = 0; j < n; j++) {
                    double a = 2.0 * pi * ((long long)j * k % n) / n;
                    re += in[j] * cos(a); im -= in[j] * sin(a);
                }
                out[k] = re;
                if (k > 0 && k < n - k) out[n - k] = im;
            }
        } else {
            for (int j = 0; j < n; j++) {
                double s = in[0];
                for (int k = 1; k < n - k; k++) {
                    double a = 2.0 * pi * ((long long)j * k % n) / n;
                    s += 2.0 * (in[k] * cos(a) - in[n - k] * sin(a));
                }
                if (n % 2 == 0) s += in[n / 2] * ((j & 1) ? -1.0 : 1.0);
                out[j] = s;
            }
        }
        return;
    }
    int h = n / 2;
    if (p->kind == FFTW_R2HC) {
        /* pack z[m] = x[2m] + i x[2m+1], FFT_h, split */
        for (int m = 0; m < h; m++) { p->wr[p->rev[m]] = in[2 * m]; p->wi[p->rev[m]] = in[2 * m + 1]; }
        ctu_shim_cfft(p, -1);
        double *zr = p->wr, *zi = p->wi;
        out[0] = zr[0] + zi[0];
        out[h] = zr[0] - zi[0];
        for (int k = 1; k < h; k++) {
            double ar = zr[k], ai = zi[k], br = zr[h - k], bi = -zi[h - k]; /* conj Z[h-k] */
            double er = 0.5 * (ar + br), ei = 0.5 * (ai + bi);               /* even part  */
            double dr = 0.5 * (ar - br), di = 0.5 * (ai - bi);               /* (Z-conj)/2 */
            double orr = di, oi = -dr;                                       /* /i         */
            double c = p->cs[2 * k], s = -p->cs[2 * k + 1];                 /* e^{-2pi i k/n} */
            double tr = orr * c - oi * s, ti = orr * s + oi * c;
            double re = er + tr, im = ei + ti;
            out[k] = re;
            out[n - k] = im;
        }
    } else { /* HC2R, unnormalised: x[j] = sum_k X[k] e^{+2 pi i jk/n} */
        /* E[k] = (X[k] + conj X[h-k]) ; O[k] = (X[k] - conj X[h-k]) e^{+2 pi i k/n};
           z[k] = E[k] + i O[k];  inverse FFT_h gives x[2m] + i x[2m+1] */
        for (int k = 0; k < h; k++) {
            double xr_, xi_, yr_, yi_;
            if (k == 0) { xr_ = in[0]; xi_ = 0.0; yr_ = in[h]; yi_ = 0.0; }
            else { xr_ = in[k]; xi_ = in[n - k]; yr_ = in[h - k]; yi_ = -in[n - (h - k)]; }
            /* conj X[h-k] */
            double er = xr_ + yr_, ei = xi_ + yi_;
            double dr = xr_ - yr_, di = xi_ - yi_;
            double c = p->cs[2 * k], s = p->cs[2 * k + 1];
            double orr = dr * c - di * s, oi = dr * s + di * c;
            double zr_ = er - oi, zi_ = ei + orr;
            p->wr[p->rev[k]] = zr_; p->wi[p->rev[k]] = zi_;
        }
        ctu_shim_cfft(p, +1);
        for (int m = 0; m < h; m++) { out[2 * m] = p->wr[m]; out[2 * m + 1] = p->wi[m]; }
    }
}

static inline void fftw_destroy_plan(fftw_plan p) {
    if (!p) return;
    free(p->cs); free(p->rev); free(p->wr); free(p->wi); free(p);
}

#ifdef __cplusplus
}
#endif
#endif
