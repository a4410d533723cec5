// Cross-frame kernels of the CtuCopy hot path (sm_100a):
//   K2/K3 k_nr_scan      extended spectral subtraction and the VAD-driven hwss / fwss / 2fwss
//                        recursions: one thread per (utterance, bin|band), sequential over
//                        frames, batched over all utterances (src/nr/nr.cc:86-140, 212-261,
//                        331-369, 397-442).  HBM-bound: 4 B in + 4 B out per element.
//   K4    k_burg         per frame, fp64: forward FFT -> (|X|^a, phase) -> unnormalised
//                        inverse FFT -> [Hann] -> Burg lattice -> LPC cepstrum
//                        (src/nr/nr.cc:281-292, src/vad/vad.cc:222-237, src/vdet/Burg.h:49-152)
//   K5    k_cepdet       per utterance: adaptive-threshold cepstral detector
//                        (src/vdet/CepstralDet.h:134-194)
//   K6    k_vad_*        VAD module: criterion, threshold state machines, majority filter,
//                        drop compaction (src/vad/vad.cc:96-107, 220-294, 329-625, 692-745;
//                        src/vad/vad.h:126-175)
//   K9    k_synth        (|X|enh, phase of X) -> inverse FFT -> overlap-add in frame order ->
//                        floor(x/correction) -> clip -> int16 (src/io/out.cc:346-451)
#ifndef CTU_NR_KERNELS_CUH
#define CTU_NR_KERNELS_CUH

#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>

#include "ctu_internal.h"
#include "ctu_kernels.cuh"
#include "ctu_any64.cuh"

namespace ctu {

constexpr int BURG_MAXC = 16;     // cepstral coefficients kept per frame (pitch of d_ceps)
enum { BURG_SRC_NR = 0, BURG_SRC_VAD = 1 };

struct NrParams {
    int mode;            // CtuNrMode
    float a, b, p;
    double ad, pd;
    int a_kind;          // 1: a == 1, 2: a == 2, 0: general
    int initsegs;
};

struct SynthParams {
    double correction;   // max over offsets of the summed overlapping Hamming windows
    int hh;              // frames before a tile that still overlap its first sample
};

struct BurgParams {
    int window, wshift, remove_dc;
    double preem;
    int fb_power;        // spectrum handed to NR is power (else magnitude)
    int a_kind; double a;  // expansion applied before the detector (hwss/fwss); 2fwss: none
    int expand;          // 0 for 2fwss
    int ncoef_nr;        // fea_ncepcoefs (src/nr/nr.cc:266-270)
    int ncoef_vad;       // vad_lpc_coefs
    int ninit; double P, Q;   // detector options <- (nr_initsegs, nr_p, nr_q)
    int use_spec_gain;   // VAD source: post-NR spectrum present (gain from d_spec) else own spectrum
    // FFT sizes other than 512: the general kernel k_burg_any (nfft == 0: the specialised k_burg)
    int nfft, log2m;
    const double2 *any_tw, *any_ts;
};

struct VadParams {
    int cri, thr, drop;
    int energy_db;
    int latency;         // rows between a feature row and the spectrum frame the VAD sees
    int order;           // majority filter length
    int cep_n;           // length of the cepstral vector (lpc: vad_lpc_coefs, fea: feature dim)
    int nbins;           // spectrum bins per frame (wfft/2 + 1)
    int has_E;           // last column of the feature rows is the _E column
    int fea_skip;        // fea criterion: WRITER column holding the reference's internal element 0 (left out of the
                         // distance, src/vad/vad.cc:262-272), or -1 when that element is not written at all
    int cep_init; double cep_p;
    double abs_thr;
    int perc_init; double perc_thr;
    int adapt_init; double adapt_q, adapt_za;
    int dyn_init; double dyn_perc, dyn_min, qmaxinc, qmaxdec, qmindec, qmininc;
};

// ------------------------------------------------------------------------------------------
static inline int build_nr_params(const ctu_config &c, int nr_mode, int vad_src, bool signal_out, int nb, NrParams &N, SynthParams &S,
                                  BurgParams &B, VadParams &V, std::string &err) {
    std::memset(&N, 0, sizeof(N)); std::memset(&S, 0, sizeof(S)); std::memset(&B, 0, sizeof(B)); std::memset(&V, 0, sizeof(V));
    N.mode = nr_mode;
    N.a = (float)c.nr_a; N.b = (float)c.nr_b; N.p = (float)c.nr_p; N.ad = c.nr_a; N.pd = c.nr_p;
    N.a_kind = (c.nr_a == 1.0) ? 1 : (c.nr_a == 2.0) ? 2 : 0;
    N.initsegs = c.nr_initsegs;
    // OLA correction (src/io/out.cc:346-372)
    {
        const int s = c.wshift, w = c.window;
        double pi = 2. * asin(1.), corr = 0.;
        for (int i = 0; i < s; i++) {
            int x = i; double y = 0.;
            while (x < w) { y += 0.54 - (1 - 0.54) * cos(2 * pi * (double)x / (w - 1.)); x += s; }
            if (y > corr) corr = y;
        }
        S.correction = corr;
        S.hh = (w + s - 1) / s - 1;
    }
    B.window = c.window; B.wshift = c.wshift; B.remove_dc = c.remove_dc; B.preem = (double)c.preem;
    B.fb_power = c.fb_power;
    B.a = c.nr_a; B.a_kind = N.a_kind; B.expand = (nr_mode != NR_2FWSS);
    B.ncoef_nr = c.fea_ncepcoefs; B.ncoef_vad = c.vad_lpc_coefs;
    B.ninit = c.nr_initsegs; B.P = c.nr_p; B.Q = c.nr_q;
    B.use_spec_gain = (nr_mode != NR_NONE && c.nr_when == 0);
    if (nr_mode >= NR_HWSS && vad_src == VADSRC_BURG && (c.fea_ncepcoefs > BURG_MAXC || c.fea_ncepcoefs < 2)) {
        err = "CTU: Burg detector supports 2..16 cepstral coefficients"; return CTU_ERR_UNSUPPORTED;
    }
    V.energy_db = c.vad_energy_db;
    V.order = c.vad_filter_order;
    V.cep_init = c.vad_cepdist_init; V.cep_p = c.vad_cepdist_p;
    V.abs_thr = c.vad_absolute_thr;
    V.perc_init = c.vad_perc_init; V.perc_thr = c.vad_perc_thr;
    V.adapt_init = c.vad_adapt_init; V.adapt_q = c.vad_adapt_q; V.adapt_za = c.vad_adapt_za;
    V.dyn_init = c.vad_dyn_init; V.dyn_perc = c.vad_dyn_perc; V.dyn_min = c.vad_dyn_min;
    V.qmaxinc = c.vad_dyn_qmaxinc; V.qmaxdec = c.vad_dyn_qmaxdec; V.qmindec = c.vad_dyn_qmindec; V.qmininc = c.vad_dyn_qmininc;
    // latency of the feature chain as BATCH::save_frame sees it (src/io/batch.cc:172-204)
    V.latency = 0;
    std::string kind(c.fea_kind);
    if (kind != "trapdct" && kind != "lpa" && c.fea_delta) {      // deltas, or one stage used as the -fea_trap window
        int wins[3] = {c.d_win, c.a_win, c.t_win};
        for (int k = 0; k < c.n_order; k++) V.latency += wins[k];
    }
    if (kind == "trapdct") V.latency = (c.fea_trapdct_traplen + 1) / 2 - 1;
    (void)signal_out; (void)nb;
    return CTU_OK;
}

// ------------------------------------------------------------------------------------------
// K2 / K3: noise-reduction scans
// ------------------------------------------------------------------------------------------
// One thread per (utterance, bin), frames walked in blocks of SCAN_UNROLL: all loads of a
// block are issued before the (sequentially dependent) recursion touches them and all
// stores after it, so each thread keeps SCAN_UNROLL x 4 B of reads in flight.  MODE / AKIND
// are compile-time so the hot loop carries only the state it needs (register count decides
// how many bytes per SM are in flight, and this kernel is HBM-bound).
constexpr int SCAN_UNROLL = 16;

struct ScanState { float Navg, Yavg, Nravg; double Nd, Yd; };

// one frame of the recursion for one bin
template <int MODE, int AKIND>
__device__ __forceinline__ float nr_step(const NrParams &N, ScanState &S, float xi, int t, uint8_t flag) {
    const float p = N.p, q = 1.f - N.p;
    if (MODE == NR_EXTEN) {
        if (AKIND == 0) {                   // general exponent: fp64, as the reference writes it
            double Hd = S.Nd / pow(pow(S.Nd, N.ad) + pow(S.Yd, N.ad), 1. / N.ad);
            double xd = (double)xi, Nn = Hd * xd;
            S.Nd = N.pd * S.Nd + (1 - N.pd) * Nn;
            S.Yd = fabs(xd - S.Nd);
            return (float)(xd - Nn);
        }
        // H = Navg/(Navg+Yavg) (a=1) or Navg/hypot(Navg,Yavg) (a=2); the output X-H*X is formed
        // as X*(1-H) with 1-H written without cancellation
        float H, omH;
        if (AKIND == 1) {
            const float r = __frcp_rn(S.Navg + S.Yavg);
            H = S.Navg * r; omH = S.Yavg * r;
        } else {
            const float h2 = fmaf(S.Navg, S.Navg, S.Yavg * S.Yavg);
            const float hh = sqrtf(h2);
            H = __fdiv_rn(S.Navg, hh);
            omH = __fdiv_rn(S.Yavg * S.Yavg, hh * (hh + S.Navg));
        }
        const float Nn = H * xi;
        S.Navg = fmaf(p, S.Navg, q * Nn);
        S.Yavg = fabsf(xi - S.Navg);
        return xi * omH;
    }
    // hwss decrements its counter before use, fwss / 2fwss after (src/nr/nr.cc:226, 367, 440)
    const int ninit = (MODE == NR_HWSS) ? N.initsegs - (t + 1) : N.initsegs - t;
    const bool upd = (flag == 0) || ninit > 0;
    if (MODE == NR_2FWSS) {
        if (upd) S.Navg = fmaf(p, S.Navg, q * xi);
        xi = fabsf(xi - S.Navg);
        if (upd) S.Nravg = fmaf(p, S.Nravg, q * xi);
        return fabsf(xi - S.Nravg);
    }
    if (AKIND == 2) xi = xi * xi;
    else if (AKIND == 0) xi = powf(xi, N.a);
    if (upd) S.Navg = fmaf(p, S.Navg, q * xi);
    xi = xi - N.b * S.Navg;
    if (MODE == NR_HWSS) { if (xi < 0.f) xi = 0.f; }
    else if (xi < 0.f) xi = -xi;
    if (AKIND == 2) xi = sqrtf(xi);
    else if (AKIND == 0) xi = powf(xi, 1.f / N.a);
    return xi;
}

// SIZE: row length known at compile time (257 spectrum bins) so that the SCAN_UNROLL loads of
// a block share one base register with immediate offsets; 0 = runtime (band domain)
template <int MODE, int AKIND, int SIZE>
__global__ void __launch_bounds__(256)
k_nr_scan(const __grid_constant__ NrParams N, const int *__restrict__ nframes, const int64_t *__restrict__ row_off, int u0, int n_utts,
          int size_rt, float *X, const uint8_t *__restrict__ flags) {
    const int size = SIZE ? SIZE : size_rt;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)n_utts * size) return;
    const int u = u0 + (int)(gid / size), bin = (int)(gid % size);
    const int T = nframes[u];
    float *x = X + row_off[u] * size + bin;
    const uint8_t *fl = (MODE != NR_EXTEN) ? flags + row_off[u] : nullptr;
    // exten: smoothed noise Navg / smoothed speech Yavg start at 0.95 / 0.05 (src/nr/nr.cc:86-93);
    // *ss: the noise estimate of a standalone file starts at 0 (src/nr/nr.cc:212-222)
    ScanState S;
    S.Navg = (MODE == NR_EXTEN) ? 0.95f : 0.f; S.Yavg = 0.05f; S.Nravg = 0.f; S.Nd = 0.95; S.Yd = 0.05;
    int t0 = 0;
    for (; t0 + SCAN_UNROLL <= T; t0 += SCAN_UNROLL) {
        float v[SCAN_UNROLL];
        uint8_t f[SCAN_UNROLL];
        float *xr = x + (int64_t)t0 * size;
#pragma unroll
        for (int j = 0; j < SCAN_UNROLL; j++) {
            v[j] = xr[j * size];
            f[j] = (MODE != NR_EXTEN) ? fl[t0 + j] : 0;
        }
#pragma unroll
        for (int j = 0; j < SCAN_UNROLL; j++) v[j] = nr_step<MODE, AKIND>(N, S, v[j], t0 + j, f[j]);
#pragma unroll
        for (int j = 0; j < SCAN_UNROLL; j++) xr[j * size] = v[j];
    }
    for (; t0 < T; t0++) {
        float *xr = x + (int64_t)t0 * size;
        *xr = nr_step<MODE, AKIND>(N, S, *xr, t0, (MODE != NR_EXTEN) ? fl[t0] : 0);
    }
}

template <int MODE, int SIZE>
static inline void launch_nr_scan_a(const NrParams &N, unsigned grid, cudaStream_t s, const int *d_nframes, const int64_t *d_row_off, int u0,
                                    int n, int size, float *X, const uint8_t *flags) {
    if (N.a_kind == 1 || MODE == NR_2FWSS) k_nr_scan<MODE, 1, SIZE><<<grid, 256, 0, s>>>(N, d_nframes, d_row_off, u0, n, size, X, flags);
    else if (N.a_kind == 2) k_nr_scan<MODE, 2, SIZE><<<grid, 256, 0, s>>>(N, d_nframes, d_row_off, u0, n, size, X, flags);
    else k_nr_scan<MODE, 0, SIZE><<<grid, 256, 0, s>>>(N, d_nframes, d_row_off, u0, n, size, X, flags);
}

template <int SIZE>
static inline void launch_nr_scan_m(const NrParams &N, unsigned grid, cudaStream_t s, const int *d_nframes, const int64_t *d_row_off, int u0,
                                    int n, int size, float *X, const uint8_t *flags) {
    switch (N.mode) {
        case NR_EXTEN: launch_nr_scan_a<NR_EXTEN, SIZE>(N, grid, s, d_nframes, d_row_off, u0, n, size, X, flags); break;
        case NR_HWSS: launch_nr_scan_a<NR_HWSS, SIZE>(N, grid, s, d_nframes, d_row_off, u0, n, size, X, flags); break;
        case NR_FWSS: launch_nr_scan_a<NR_FWSS, SIZE>(N, grid, s, d_nframes, d_row_off, u0, n, size, X, flags); break;
        default: launch_nr_scan_a<NR_2FWSS, SIZE>(N, grid, s, d_nframes, d_row_off, u0, n, size, X, flags); break;
    }
}

static inline int launch_nr_scan(const NrParams &N, const int *d_nframes, const int64_t *d_row_off, int u0, int u1, int size, float *X,
                                 const uint8_t *flags, cudaStream_t s, LaunchCtx *lc, std::string &err) {
    int64_t n = (int64_t)(u1 - u0) * size;
    if (n <= 0) return CTU_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    lc->begin("k_nr_scan", s);
    if (size == NBIN) launch_nr_scan_m<NBIN>(N, grid, s, d_nframes, d_row_off, u0, u1 - u0, size, X, flags);
    else launch_nr_scan_m<0>(N, grid, s, d_nframes, d_row_off, u0, u1 - u0, size, X, flags);
    lc->end(s);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e) + " (k_nr_scan)"; return CTU_ERR_CUDA; }
    return CTU_OK;
}

// ------------------------------------------------------------------------------------------
// shared front end in double precision for one frame held by a 16-thread group
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double group_sum16d(double v) {
    const unsigned m = 0xffffu << (threadIdx.x & 16);
    v += __shfl_xor_sync(m, v, 8);
    v += __shfl_xor_sync(m, v, 4);
    v += __shfl_xor_sync(m, v, 2);
    v += __shfl_xor_sync(m, v, 1);
    return v;
}

// ------------------------------------------------------------------------------------------
// K4: Burg cepstrum per frame (fp64 throughout so that detector decisions are reproducible)
// 128 threads = 8 frames per pass, 4 passes per 32-frame tile.
//   * the tile's PCM is staged once into shared memory as int16 (16-byte loads);
//   * forward FFT -> per-bin gain (|X|^a or the post-NR magnitude, over |X|) -> inverse FFT;
//     the time signal reuses the exchange tile's memory;
//   * lattice: sample i = c*CH + j lives in thread c's registers.  All updates are
//     unconditional; the elements the reference no longer reads (i < ik) are driven to exact
//     zeros instead of being masked: thread 0 keeps ef[0] = 0 and takes `below` = 0, which
//     makes eb[ik-1] come out 0 by itself, and ef[ik] is cleared after stage ik (a select
//     chain over the first 16 registers, so every register index stays static);
//   * the predictor coefficients are distributed (thread i holds a_i; one shuffle per stage).
// CH: samples per thread; EXACT: window == 16*CH (no tail masking in the energy sums).
// ------------------------------------------------------------------------------------------
constexpr int BURG_THREADS = 128;
constexpr int BURG_GROUPS = BURG_THREADS / GROUP;

__device__ __forceinline__ double shfl16d(double v, int src) {
    const unsigned m = 0xffffu << (threadIdx.x & 16);
    return __shfl_sync(m, v, src, 16);
}

template <int CH, bool EXACT, int MINB>
__global__ void __launch_bounds__(BURG_THREADS, MINB)
k_burg(const __grid_constant__ BurgParams B, int src_mode, BatchDesc bd, const int16_t *__restrict__ pcm, const float *__restrict__ spec,
       double *__restrict__ ceps, const double2 *__restrict__ g_tw256, const double2 *__restrict__ g_twsplit,
       const double2 *__restrict__ g_twinv, const double *__restrict__ g_win, const double *__restrict__ g_hann) {
    extern __shared__ __align__(16) double smd[];
    const int tid = threadIdx.x;
    const int w = EXACT ? 16 * CH : B.window, s = B.wshift;
    cpx<double> *sTw = reinterpret_cast<cpx<double> *>(smd);            // 256
    cpx<double> *sTs = sTw + 256;                                      // 129 (+1 pad)
    cpx<double> *sTi = sTs + 130;                                      // 129 (+1 pad)
    cpx<double> *sX = sTi + 130;                                       // BURG_GROUPS * 16*17 (also the time signal)
    double *sWin = reinterpret_cast<double *>(sX + BURG_GROUPS * XPAD * 16);   // 512: analysis window
    double *sHann = sWin + NFFT;                                       // 512: detector's Hann (NR source)
    int16_t *sPcm = reinterpret_cast<int16_t *>(sHann + NFFT);         // 8 + (TILE_F-1)*s + w + 1 + 8
    const int2 tile = bd.tiles[blockIdx.x];
    const int u = tile.x, t0 = tile.y;
    const int nf = min(TILE_F, bd.nframes[u] - t0);
    const int64_t row0 = bd.row_off[u] + t0;
    // samples [first-1, first + nsamp): the one before the tile feeds the first pre-emphasis
    const int nsamp = (nf - 1) * s + w + 1;
    const bool at_start = (t0 == 0);
    const int16_t *src = pcm + bd.pcm_off[u] + (int64_t)t0 * s - 1;
    const int phase = (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 1);      // same 16-byte phase in shared memory
    int16_t *dpcm = sPcm + phase;
    for (int k0 = tid * 8 - phase; k0 < nsamp; k0 += BURG_THREADS * 8) {
        if (k0 >= (at_start ? 1 : 0) && k0 + 8 <= nsamp) {
            *reinterpret_cast<int4 *>(dpcm + k0) = __ldg(reinterpret_cast<const int4 *>(src + k0));
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = k0 + j;
                if (k >= 0 && k < nsamp) dpcm[k] = (k == 0 && at_start) ? (int16_t)0 : src[k];
            }
        }
    }
    for (int i = tid; i < 256; i += BURG_THREADS) sTw[i] = mk<double>(g_tw256[i].x, g_tw256[i].y);
    for (int i = tid; i < 129; i += BURG_THREADS) { sTs[i] = mk<double>(g_twsplit[i].x, g_twsplit[i].y); sTi[i] = mk<double>(g_twinv[i].x, g_twinv[i].y); }
    for (int i = tid; i < NFFT; i += BURG_THREADS) {
        sWin[i] = (i < w) ? g_win[i] : 0.0;
        sHann[i] = (i < w) ? ((src_mode == BURG_SRC_NR) ? g_hann[i] : 1.0) : 0.0;
    }
    __syncthreads();
    const int c = tid & (GROUP - 1), grp = tid / GROUP;
    cpx<double> *xch = sX + grp * (XPAD * 16);
    double *xt = reinterpret_cast<double *>(xch);
    const int ncoef = (src_mode == BURG_SRC_NR) ? B.ncoef_nr : B.ncoef_vad;
    const double inv_w = 1.0 / (double)w;
#pragma unroll 1
    for (int pass = 0; pass < TILE_F / BURG_GROUPS; pass++) {
        const int f = pass * BURG_GROUPS + grp;
        const bool active = f < nf;
        {
            cpx<double> a[16];
            cpx<double> lo[8], hi[8], mid;
            if (active) {
                const int16_t *x = dpcm + f * s + 1;                  // x[-1] is the sample before the frame
                double sum = 0;
#pragma unroll
                for (int n1 = 0; n1 < 16; n1++) {
                    const int i0 = 32 * n1 + 2 * c;
                    double y0 = 0, y1 = 0;
                    if (i0 < w) {                                     // w is even for every supported window
                        const double xm = (double)x[i0 - 1], x0 = (double)x[i0], x1 = (double)x[i0 + 1];
                        y0 = sWin[i0] * (x0 - B.preem * xm);
                        y1 = sWin[i0 + 1] * (x1 - B.preem * x0);      // sWin is 0 beyond the window
                    }
                    a[n1] = mk<double>(y0, y1);
                    sum += y0 + y1;
                }
                if (B.remove_dc) {
                    const double mean = group_sum16d(sum) * inv_w;
#pragma unroll
                    for (int n1 = 0; n1 < 16; n1++) {
                        const int i0 = 32 * n1 + 2 * c;
                        if (i0 < w) a[n1].x -= mean;
                        if (i0 + 1 < w) a[n1].y -= mean;
                    }
                }
                fft256_pass1(a, c, sTw, xch);
            }
            __syncwarp();
            if (active) {
                fft256_pass2(a, c, xch);
                rfft_split_shfl(a, c, sTs, lo, hi, mid);
                // (|X|^a or the post-NR spectrum) with the phase of X: every bin is scaled by
                // E/|X| -- what Xa*cos(phi), Xa*sin(phi) amount to (src/nr/nr.cc:281-292,
                // src/vad/vad.cc:222-233) -- with the reference's conventions for bin 0
                // (phase 0, src/io/in.cc:398), the Nyquist bin (real) and atan(0/0) = -pi/2
                const float *srow = (src_mode == BURG_SRC_VAD && B.use_spec_gain) ? spec + (row0 + f) * NBIN : nullptr;
                const bool expand = (src_mode == BURG_SRC_NR && B.expand);
                auto scale_bin = [&](cpx<double> X, int k) -> cpx<double> {
                    double m2 = X.x * X.x + X.y * X.y;
                    const bool edge = (k == 0 || k == NC);
                    if (k == 0) { if (B.remove_dc) m2 = 1e-10; X = mk<double>(sqrt(m2), 0.0); }
                    const double rm = (m2 > 0.0) ? rsqrt(m2) : 0.0;          // 1/|X|
                    const double m = m2 * rm;                                // |X|
                    double E, g;
                    if (srow) { E = (double)srow[k]; g = E * rm; }
                    else {
                        // E = Xa^a with Xa = |X|^2 (fb_power) or |X|;  g = E/|X| without the division
                        // where the exponents are small integers
                        const int ak = expand ? B.a_kind : 1;
                        if (ak == 0) { E = pow(B.fb_power ? m2 : m, B.a); g = E * rm; }
                        else if (B.fb_power) { E = (ak == 2) ? m2 * m2 : m2; g = (ak == 2) ? m2 * m : m; }
                        else { E = (ak == 2) ? m2 : m; g = (ak == 2) ? m : 1.0; }
                    }
                    if (m2 == 0.0) return edge ? mk<double>(E, 0.0) : mk<double>(0.0, -E);
                    return mk<double>(X.x * g, edge ? 0.0 : X.y * g);
                };
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int k = c + 16 * j;
                    lo[j] = scale_bin(lo[j], k);
                    hi[j] = scale_bin(hi[j], NC - k);
                }
                mid = scale_bin(mid, 128);
                irfft_presplit_shfl(a, c, sTi, lo, hi, mid);
            }
            __syncwarp();                                             // pass-2 reads of the exchange tile are done
            if (active) fft256_pass1(a, c, sTw, xch);
            __syncwarp();
            if (active) fft256_pass2(a, c, xch);
            __syncwarp();                                             // xch is re-used for the time signal
            if (active) {
#pragma unroll
                for (int k2 = 0; k2 < 16; k2++) {
                    const int n = c + 16 * k2;
                    xt[2 * n] = a[k2].x;
                    xt[2 * n + 1] = -a[k2].y;
                }
            }
            __syncwarp();
        }
        if (active) {
            // ---- Burg lattice (src/vdet/Burg.h:49-95) on the first w samples ------------------
            double ef[CH], eb[CH];
            double en = 0;
#pragma unroll
            for (int j = 0; j < CH; j++) {
                const int i = c * CH + j;
                double v = (EXACT || i < w) ? xt[i < NFFT ? i : 0] * sHann[i < NFFT ? i : 0] : 0.0;
                if (!EXACT && i >= w) v = 0.0;
                ef[j] = eb[j] = v;
                en += v * v;
            }
            double alpha = group_sum16d(en) * inv_w;
            if (c == 0) ef[0] = 0.0;                                  // never read by the reference
            double a_c = (c == 0) ? 1.0 : 0.0, aa_c = a_c;            // thread i holds a_i
            // the stage loop stays rolled: unrolled 15 times the kernel was 14 k instructions and stalled on instruction
            // fetch (30.4 -> 27.5 ms on 4 M frames)
#pragma unroll 1
            for (int ik = 1; ik < ncoef; ik++) {
                {
                    double below = shfl16d(eb[CH - 1], (c + 15) & 15);
                    if (c == 0) below = 0.0;
                    // three independent chains per parity: the sums are latency-bound otherwise
                    double nu[2] = {0, 0}, df[2] = {0, 0}, db[2] = {0, 0};
#pragma unroll
                    for (int j = 0; j < CH; j++) {
                        const double pv = (j > 0) ? eb[j > 0 ? j - 1 : 0] : below;
                        if (EXACT || c * CH + j < w) {
                            df[j & 1] = fma(ef[j], ef[j], df[j & 1]);
                            db[j & 1] = fma(pv, pv, db[j & 1]);
                            nu[j & 1] = fma(ef[j], pv, nu[j & 1]);
                        }
                    }
                    const double num = group_sum16d(nu[0] + nu[1]) * 2.0;
                    const double den = group_sum16d((df[0] + df[1]) + (db[0] + db[1]));
                    const double rc = -num / den;
                    alpha *= 1 - rc * rc;
#pragma unroll
                    for (int j = CH - 1; j >= 0; j--) {
                        const double pv = (j > 0) ? eb[j > 0 ? j - 1 : 0] : below;
                        const double e0 = ef[j];
                        ef[j] = e0 + rc * pv;
                        eb[j] = pv + rc * e0;
                    }
#pragma unroll
                    for (int j = 1; j < BURG_MAXC && j < CH; j++) if (c == 0 && j == ik) ef[j] = 0.0;
                    // a_i = aa_i + rc * aa_{ik-i} (0 < i < ik), a_ik = rc
                    const double other = shfl16d(aa_c, (ik - c) & 15);
                    if (c == ik) a_c = rc;
                    else if (c >= 1 && c < ik) a_c = aa_c + rc * other;
                    aa_c = a_c;
                }
            }
            // LPC -> cepstrum (src/vdet/Burg.h:141-152), thread 0 of the group
            double av[BURG_MAXC];
#pragma unroll
            for (int k = 0; k < BURG_MAXC; k++) av[k] = shfl16d(a_c, k);
            if (c == 0) {
                double cc[BURG_MAXC];
                double *o = ceps + (row0 + f) * BURG_MAXC;
#pragma unroll
                for (int n = 1; n < BURG_MAXC; n++) {
                    double sum = 0;
#pragma unroll
                    for (int k = 1; k < BURG_MAXC; k++) if (k < n) sum += (double)(n - k) * cc[(n - k) & (BURG_MAXC - 1)] * av[k];
                    cc[n] = -av[n] - sum / n;
                    if (n < ncoef) o[n] = cc[n];
                }
                o[0] = log(alpha);
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// K4g: the Burg front end for FFT sizes other than 512 (fwss / hwss / 2fwss and the LPC cepstral-distance criterion at
// 8 kHz, 22-48 kHz).  Same arithmetic as k_burg, one WARP per frame, everything in shared memory:
//   frame -> FFT -> per-bin gain (|X|^a or the post-NR magnitude, on the phase of X) -> unnormalised inverse -> first
//   `window` samples [x Hann for the NR detector] -> Burg lattice (src/vdet/Burg.h:49-95) -> cepstrum (:141-152).
// The lattice keeps ef (aliasing the time signal) and eb in shared memory; a stage updates them in chunks of 32 samples
// from the END of the frame, so that eb[i-1] is still the old value when sample i is updated.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ANY64_THREADS)
k_burg_any(const __grid_constant__ BurgParams B, int src_mode, BatchDesc bd, AnyTables64 tb, const int16_t *__restrict__ pcm,
           const float *__restrict__ spec, double *__restrict__ ceps, const double *__restrict__ g_hann) {
    extern __shared__ __align__(16) double smd[];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const int nfft = tb.nfft, M = nfft >> 1, nbins = M + 1;
    cpx<double> *z = reinterpret_cast<cpx<double> *>(smd + (size_t)wv * (6 * M + 4));   // M complex = nfft reals
    cpx<double> *Y = z + M;                                                             // M + 1 complex (+ pad)
    double *eb = reinterpret_cast<double *>(Y + M + 2);                                 // nfft reals
    double *ef = reinterpret_cast<double *>(z);
    const int2 tile = bd.tiles[blockIdx.x];
    const int u = tile.x, t0 = tile.y;
    const int nf = min(TILE_F, bd.nframes[u] - t0);
    const int64_t row0 = bd.row_off[u] + t0;
    const int w = B.window, s = B.wshift;
    const int ncoef = (src_mode == BURG_SRC_NR) ? B.ncoef_nr : B.ncoef_vad;
    for (int f = wv; f < nf; f += ANY64_THREADS / 32) {
        any64_analysis(z, tb, pcm + bd.pcm_off[u] + (int64_t)(t0 + f) * s, (t0 + f) == 0, w, B.preem, B.remove_dc, lane);
        // (|X|^a or the post-NR spectrum) with the phase of X -- see k_burg for the conventions (bin 0: phase 0 and the fixed
        // 1e-10 floor under -remove_dc, src/io/in.cc:390-398; Nyquist real; atan(0/0) = -pi/2)
        const float *srow = (src_mode == BURG_SRC_VAD && B.use_spec_gain) ? spec + (row0 + f) * nbins : nullptr;
        const bool expand = (src_mode == BURG_SRC_NR && B.expand);
        for (int k = lane; k <= M; k += 32) {
            cpx<double> X = any64_bin(z, tb, k);
            double m2 = X.x * X.x + X.y * X.y;
            const bool edge = (k == 0 || k == M);
            if (k == 0) { if (B.remove_dc) m2 = 1e-10; X = mk<double>(sqrt(m2), 0.0); }
            const double rm = (m2 > 0.0) ? rsqrt(m2) : 0.0;
            const double m = m2 * rm;
            double E, g;
            if (srow) { E = (double)srow[k]; g = E * rm; }
            else {
                const int ak = expand ? B.a_kind : 1;
                if (ak == 0) { E = pow(B.fb_power ? m2 : m, B.a); g = E * rm; }
                else if (B.fb_power) { E = (ak == 2) ? m2 * m2 : m2; g = (ak == 2) ? m2 * m : m; }
                else { E = (ak == 2) ? m2 : m; g = (ak == 2) ? m : 1.0; }
            }
            Y[k] = (m2 == 0.0) ? (edge ? mk<double>(E, 0.0) : mk<double>(0.0, -E)) : mk<double>(X.x * g, edge ? 0.0 : X.y * g);
        }
        __syncwarp();
        any64_inverse(z, Y, tb, lane);
        // ---- Burg lattice on the first w samples ------------------------------------------------------------------
        double en = 0.0;
        for (int i = lane; i < w; i += 32) {
            const double v = ef[i] * ((src_mode == BURG_SRC_NR) ? g_hann[i] : 1.0);
            ef[i] = v; eb[i] = v;
            en += v * v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) en += __shfl_xor_sync(0xffffffffu, en, o);
        double alpha = en / (double)w;
        __syncwarp();
        double a_c = (lane == 0) ? 1.0 : 0.0, aa_c = a_c;              // lane i holds a_i
#pragma unroll 1
        for (int ik = 1; ik < ncoef; ik++) {
            double num = 0.0, den = 0.0;
            for (int i = ik + lane; i < w; i += 32) {
                const double e1 = ef[i], e2 = eb[i - 1];
                num = fma(e1, e2, num);
                den += e1 * e1 + e2 * e2;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
            const double rc = -(2.0 * num) / den;
            alpha *= 1 - rc * rc;
            for (int base = (w - 1) & ~31; base >= 0; base -= 32) {
                const int i = base + lane;
                const bool ok = i >= 1 && i < w;
                double e0 = 0.0, pv = 0.0;
                if (ok) { e0 = ef[i]; pv = eb[i - 1]; }
                __syncwarp();
                if (ok) { ef[i] = e0 + rc * pv; eb[i] = pv + rc * e0; }
                __syncwarp();
            }
            const double other = __shfl_sync(0xffffffffu, aa_c, (ik - lane) & 31);
            if (lane == ik) a_c = rc;
            else if (lane >= 1 && lane < ik) a_c = aa_c + rc * other;
            aa_c = a_c;
        }
        double av[BURG_MAXC];
#pragma unroll
        for (int k = 0; k < BURG_MAXC; k++) av[k] = __shfl_sync(0xffffffffu, a_c, k);
        if (lane == 0) {
            double cc[BURG_MAXC];
            double *o = ceps + (row0 + f) * BURG_MAXC;
#pragma unroll
            for (int n = 1; n < BURG_MAXC; n++) {
                double sum = 0;
#pragma unroll
                for (int k = 1; k < BURG_MAXC; k++) if (k < n) sum += (double)(n - k) * cc[(n - k) & (BURG_MAXC - 1)] * av[k];
                cc[n] = -av[n] - sum / n;
                if (n < ncoef) o[n] = cc[n];
            }
            o[0] = log(alpha);
        }
        __syncwarp();
    }
}

static inline size_t burg_smem_bytes(int w, int s) {
    return sizeof(double) * (2 * (256 + 130 + 130 + BURG_GROUPS * XPAD * 16) + 2 * NFFT) + sizeof(int16_t) * (size_t)(8 + (TILE_F - 1) * s + w + 1 + 8 + 8);
}

static inline int launch_burg(const BurgParams &B, int src_mode, const BatchDesc &bd, int64_t ntiles, const int16_t *pcm, const float *spec,
                              double *ceps, const double2 *tw, const double2 *ts, const double2 *ti, const double *win, const double *hann,
                              cudaStream_t s, LaunchCtx *lc, std::string &err) {
    if (ntiles <= 0) return CTU_OK;
    if (B.nfft) {                                        // FFT sizes other than 512: the general kernel
        const int M = B.nfft / 2;
        const size_t bytes_any = (size_t)(ANY64_THREADS / 32) * (6 * M + 4) * sizeof(double);
        AnyTables64 tb{B.any_tw, B.any_ts, win, B.nfft, B.log2m};
        cudaError_t e2 = cudaFuncSetAttribute(k_burg_any, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes_any);
        lc->begin("k_burg_any", s);
        if (e2 == cudaSuccess) k_burg_any<<<(unsigned)ntiles, ANY64_THREADS, bytes_any, s>>>(B, src_mode, bd, tb, pcm, spec, ceps, hann);
        lc->end(s);
        if (e2 == cudaSuccess) e2 = cudaGetLastError();
        if (e2 != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e2) + " (k_burg_any)"; return CTU_ERR_CUDA; }
        return CTU_OK;
    }
    if (B.window & 1) { err = "CTU: the Burg detector path needs an even window length"; return CTU_ERR_UNSUPPORTED; }
    size_t bytes = burg_smem_bytes(B.window, B.wshift);
    if (bytes > 227 * 1024) { err = "CTU: window shift too large for the Burg detector kernel"; return CTU_ERR_UNSUPPORTED; }
    cudaError_t e;
    lc->begin("k_burg", s);
#define CTU_BURG_LAUNCH(CH, EX, MB)                                                                                    \
    e = cudaFuncSetAttribute(k_burg<CH, EX, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);             \
    if (e == cudaSuccess) k_burg<CH, EX, MB><<<(unsigned)ntiles, BURG_THREADS, bytes, s>>>(B, src_mode, bd, pcm, spec, ceps, tw, ts, ti, win, hann)
    // two CTAs per SM: capping registers for a third one spills the lattice state and gains nothing (measured)
    if (B.window == 400) { CTU_BURG_LAUNCH(25, true, 2); }
    else if (B.window == 512) { CTU_BURG_LAUNCH(32, true, 2); }
    else if (B.window <= 400) { CTU_BURG_LAUNCH(25, false, 2); }
    else { CTU_BURG_LAUNCH(32, false, 2); }
#undef CTU_BURG_LAUNCH
    lc->end(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e) + " (k_burg)"; return CTU_ERR_CUDA; }
    return CTU_OK;
}

// ------------------------------------------------------------------------------------------
// K5: cepstral detector (src/vdet/CepstralDet.h:134-194), sequential per utterance
// ------------------------------------------------------------------------------------------
// One WARP per utterance: the warp stages blocks of 32 frames of cepstra in shared memory (coalesced, all loads
// in flight at once) and lane 0 runs the sequential state machine out of shared memory.  With one thread per
// utterance every frame waited for its own HBM / L2 round trip (2 ms per launch whatever the batch size, which
// dominated the chunked host path).
constexpr int CEPDET_WARPS = 4;
__global__ void __launch_bounds__(32 * CEPDET_WARPS)
k_cepdet(const __grid_constant__ BurgParams B, const int *__restrict__ nframes, const int64_t *__restrict__ row_off, int u0,
         int n_utts, const double *__restrict__ ceps, uint8_t *__restrict__ flags) {
    __shared__ double stage[CEPDET_WARPS][32 * BURG_MAXC];
    __shared__ uint8_t sflag[CEPDET_WARPS][32];
    const int wv = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * CEPDET_WARPS + wv;
    if (i >= n_utts) return;
    const int u = u0 + i;
    const int T = nframes[u];
    const int nc = B.ncoef_nr;
    const double *cp0 = ceps + row_off[u] * BURG_MAXC;
    uint8_t *fl = flags + row_off[u];
    double c0[BURG_MAXC];
    double dMean = 0, dMean2 = 0, dVar = 0, thr = 0;
    for (int tb = 0; tb < T; tb += 32) {
        const int nb = min(32, T - tb);
        for (int k = lane; k < nb * BURG_MAXC; k += 32) stage[wv][k] = cp0[(int64_t)tb * BURG_MAXC + k];
        __syncwarp();
        if (lane == 0) {
            for (int j = 0; j < nb; j++) {
                const int t = tb + j;
                const double *cp = stage[wv] + j * BURG_MAXC;
                bool res = false;
                if (t == 0) {
                    for (int k = 0; k < nc; k++) c0[k] = cp[k];
                } else {
                    if (t == 1) for (int k = 0; k < nc; k++) c0[k] = (c0[k] + cp[k]) / 2.0;
                    double sum = 0;
                    for (int k = 1; k < nc; k++) { double d = cp[k] - c0[k]; sum += d * d; }
                    const double dist = 4.3429 * sqrt(2 * sum);
                    if (t == 1) { dMean = dist; dMean2 = dist * dist; thr = dMean; }
                    else {
                        res = (t > B.ninit) && (dist >= thr);
                        if (!res) {
                            for (int k = 0; k < nc; k++) c0[k] = B.P * c0[k] + (1 - B.P) * cp[k];
                            dMean = B.Q * dMean + (1 - B.Q) * dist;
                            dMean2 = B.Q * dMean2 + (1 - B.Q) * dist * dist;
                            dVar = dMean2 - dMean * dMean;
                            thr = dMean + 2.0 * sqrt(dVar);
                        }
                    }
                }
                sflag[wv][j] = res ? 1 : 0;
            }
        }
        __syncwarp();
        if (lane < nb) fl[tb + lane] = sflag[wv][lane];
        __syncwarp();
    }
}

static inline int launch_cepdet(const BurgParams &B, const int *d_nframes, const int64_t *d_row_off, int u0, int u1, const double *ceps,
                                uint8_t *flags, cudaStream_t s, LaunchCtx *lc, std::string &err) {
    int n = u1 - u0;
    if (n <= 0) return CTU_OK;
    lc->begin("k_cepdet", s);
    k_cepdet<<<(n + CEPDET_WARPS - 1) / CEPDET_WARPS, 32 * CEPDET_WARPS, 0, s>>>(B, d_nframes, d_row_off, u0, n, ceps, flags);
    lc->end(s);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e) + " (k_cepdet)"; return CTU_ERR_CUDA; }
    return CTU_OK;
}

// ------------------------------------------------------------------------------------------
// K6: VAD module
// ------------------------------------------------------------------------------------------
// energy criterion: one warp per frame over the (post-NR) spectrum
__global__ void k_vad_energy(const __grid_constant__ VadParams V, const float *__restrict__ spec, int64_t row0, int64_t nrows,
                             double *__restrict__ cri) {
    const int64_t r = row0 + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= row0 + nrows) return;
    const int lane = threadIdx.x & 31;
    const float *x = spec + r * V.nbins;
    double e = 0;
    for (int k = lane; k < V.nbins; k += 32) { double v = (double)x[k]; e += v * v; }
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if (lane == 0) cri[r] = V.energy_db ? 10.0 * log10(DBL_MIN + e) : e;
}

// thresholds + background update + majority filter (src/vad/vad.cc:249-276, 300-420, src/vad/vad.h:126-175).
// One WARP per utterance, like k_cepdet: the warp stages the criterion inputs of 32 frames in shared memory (coalesced, all
// loads in flight at once), lane 0 runs the sequential state machines out of shared memory -- same operations in the same
// order as before -- and the majority filter runs one row per lane.  With one thread per utterance every frame waited for
// its own memory round trip: 1-3 ms per launch whatever the batch size, which the chunked host path paid per chunk.
constexpr int VADSCAN_WARPS = 4;
__global__ void __launch_bounds__(VADSCAN_WARPS * 32)
k_vad_scan(const __grid_constant__ VadParams V, const int *__restrict__ nframes, const int64_t *__restrict__ row_off, int u0,
           int n_utts, const double *__restrict__ cri_frame, const double *__restrict__ ceps,
           const double *__restrict__ fea, int fea_dim /* row pitch */, uint8_t *__restrict__ vad0_tmp, uint8_t *__restrict__ vad_out,
           uint8_t *__restrict__ keep) {
    extern __shared__ __align__(16) double sm_vs[];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const int i = blockIdx.x * VADSCAN_WARPS + wv;
    if (i >= n_utts) return;
    const int u = u0 + i;
    const int T = nframes[u];
    const int64_t R0 = row_off[u];
    const int cn = (V.cri == VCRI_ENERGY) ? 1 : V.cep_n;
    double *sc = sm_vs + (size_t)wv * (32 * cn + 64 + 4);      // [32][cn] staged criterion inputs
    double *c0 = sc + 32 * cn;                                // [64] running background (cepstral criteria)
    uint8_t *sv = reinterpret_cast<uint8_t *>(c0 + 64);       // [32] decisions of the block
    double s_min = 0, s_max = 0, mean = 0, mean2 = 0, var = 0, dmin = 0, dmax = 0, dyn = 0;
    for (int base = 0; base < T; base += 32) {
        const int n = min(32, T - base);
        if (V.cri == VCRI_ENERGY) {
            if (lane < n) sc[lane] = cri_frame[R0 + min(base + lane + V.latency, T - 1)];
        } else {
            for (int idx = lane; idx < n * cn; idx += 32) {
                const int j = idx / cn, k = idx - j * cn;
                const int64_t fidx = R0 + min(base + j + V.latency, T - 1);   // spectrum frame this row is paired with
                sc[idx] = (V.cri == VCRI_CEPDIST_LPC) ? ceps[fidx * BURG_MAXC + k] : fea[(R0 + base + j) * fea_dim + k];
            }
        }
        __syncwarp();
        if (lane == 0) {
            for (int j = 0; j < n; j++) {
                const int r = base + j;
                const double *ci_row = sc + j * cn;
                double c;
                if (V.cri == VCRI_ENERGY) {
                    c = ci_row[0];
                } else {
                    // cepstral distance to the running background c0 (src/vad/vad.cc:249-276)
                    double sum = 0;
                    if (r == 0) {
                        for (int k = 0; k < cn; k++) c0[k] = ci_row[k];
                        c = 0.0;
                    } else {
                        for (int k = 0; k < cn; k++) {
                            const double ci = ci_row[k];
                            if (r == 1) c0[k] = (c0[k] + ci) / 2.0;
                            if ((V.cri == VCRI_CEPDIST_LPC) ? (k >= 1) : (k != V.fea_skip)) { double d = ci - c0[k]; sum += d * d; }
                        }
                        c = 4.3429 * sqrt(2 * sum);
                    }
                }
                bool v;
                if (V.thr == VTHR_ABSOLUTE) {
                    v = c >= V.abs_thr;
                } else if (V.thr == VTHR_PERC) {
                    if (r == 0 || (double)r < (double)V.perc_init) { s_min = c; s_max = c; }
                    else { s_min = (c < s_min) ? c : s_min; s_max = (c > s_max) ? c : s_max; }
                    double th = s_min + (V.perc_thr / 100.0) * (s_max - s_min);
                    v = c >= th;
                } else if (V.thr == VTHR_ADAPT) {
                    if (r == 0) { mean = c; mean2 = c * c; var = 0.0; v = false; }
                    else {
                        double th = mean + V.adapt_za * sqrt(var);
                        if ((c < th) || (r <= V.adapt_init)) {
                            mean = V.adapt_q * mean + (1.0 - V.adapt_q) * c;
                            mean2 = V.adapt_q * mean2 + (1.0 - V.adapt_q) * c * c;
                            var = mean2 - mean * mean;
                            v = false;
                        } else v = true;
                    }
                } else {
                    const int i0 = max(1, V.dyn_init);
                    if (r < i0) { dmax = c; dmin = c; dyn = 0.0; v = false; }
                    else if (r == i0) {
                        dmax = fmax(dmax, c) + V.dyn_min / 10.0;
                        dmin = fmin(dmin, c) - V.dyn_min / 10.0;
                        dyn = dmax - dmin; v = false;
                    } else {
                        if (dmax < c) dmax = V.qmaxinc * dmax + (1.0 - V.qmaxinc) * c; else dmax = V.qmaxdec * dmax + (1.0 - V.qmaxdec) * c;
                        if (dmin > c) dmin = V.qmindec * dmin + (1.0 - V.qmindec) * c; else dmin = V.qmininc * dmin + (1.0 - V.qmininc) * c;
                        dyn = dmax - dmin;
                        double th = dmin + (V.dyn_perc / 100.0) * dyn;
                        v = (c > th) && (dyn > V.dyn_min);
                    }
                }
                sv[j] = v ? 1 : 0;
                if (V.cri != VCRI_ENERGY && !(v && r > V.cep_init)) {
                    for (int k = 0; k < cn; k++) c0[k] = V.cep_p * c0[k] + (1.0 - V.cep_p) * ci_row[k];
                }
            }
        }
        __syncwarp();
        if (lane < n) vad0_tmp[R0 + base + lane] = sv[lane];
        __syncwarp();
    }
    // majority vote over `order` decisions centred on the row; zeros beyond both ends
    const int h = (V.order - 1) / 2;
    for (int r = lane; r < T; r += 32) {
        int sum = 0;
        for (int q = r + h - V.order + 1; q <= r + h; q++)
            if (q >= 0 && q < T) sum += vad0_tmp[R0 + q];
        bool dec = ((double)sum / (double)V.order) >= 0.5;
        if (vad_out) vad_out[R0 + r] = dec ? 1 : 0;
        keep[R0 + r] = (dec || !V.drop) ? 1 : 0;
    }
}

// drop mode: compact the kept rows of each utterance towards its first row (one CTA per
// utterance, chunks of 32 rows staged through shared memory so in-place moves are safe)
__global__ void __launch_bounds__(256)
k_vad_compact(const int *__restrict__ nframes, const int64_t *__restrict__ row_off, int u0, const uint8_t *__restrict__ keep, float *fea,
              int dim, int *__restrict__ rows_out) {
    extern __shared__ __align__(16) float smc[];
    __shared__ int s_pos[33];
    const int u = u0 + blockIdx.x;
    const int T = nframes[u];
    const int64_t R0 = row_off[u];
    int written = 0;
    for (int base = 0; base < T; base += 32) {
        const int n = min(32, T - base);
        for (int i = threadIdx.x; i < n * dim; i += blockDim.x) smc[i] = fea[(R0 + base) * dim + i];
        if (threadIdx.x == 0) {
            int pos = 0;
            for (int j = 0; j < n; j++) { s_pos[j] = keep[R0 + base + j] ? pos++ : -1; }
            s_pos[32] = pos;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n * dim; i += blockDim.x) {
            int j = i / dim, col = i - j * dim;
            int pos = s_pos[j];
            if (pos >= 0) fea[(R0 + written + pos) * dim + col] = smc[i];
        }
        written += s_pos[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) rows_out[u] = written;
}

// ------------------------------------------------------------------------------------------
// K9: synthesis.  One CTA = one tile of 32 output hops of one utterance (plus the frames
// before the tile that still overlap it).  Every frame's time signal goes to its own
// shared-memory slot; the overlap-add then sums the slots in frame order, exactly the order
// in which the reference accumulates (src/io/out.cc:427-429), so the result is deterministic.
// ------------------------------------------------------------------------------------------
constexpr int SYN_THREADS = 256;
constexpr int SYN_GROUPS = SYN_THREADS / GROUP;

// frames a synthesis tile holds: one per group, a single pass.  A tile of the plan's synthesis
// tile list covers SYN_FRAMES - hh new hops, so that together with the hh frames before it
// that still overlap its first sample every group is busy.  (Two passes per CTA needed 140 KB
// of shared memory = one CTA per SM; one pass needs 91 KB = two.)
constexpr int SYN_FRAMES = SYN_GROUPS;

// WT / ST: window and shift known at compile time (0 = runtime) -- turns the divisions of the
// overlap-add index arithmetic into shifts / multiplies and prunes the zero padding.
// one synthesis tile: utterance, hops [t0, t0+nf), frames [tfirst, t0+nf) to transform
struct SynTile { int u, t0, T, nf, tfirst, nfr; TileMeta st; };
__device__ __forceinline__ SynTile load_syn_tile(const BatchDesc &bd, int tile, int tile_frames, int hh, int s) {
    SynTile m;
    const int2 t = bd.tiles[tile];
    m.u = t.x; m.t0 = t.y;
    m.T = bd.nframes[m.u];
    m.nf = min(tile_frames, m.T - m.t0);
    m.tfirst = max(m.t0 - hh, 0);
    m.nfr = m.t0 + m.nf - m.tfirst;
    m.st.u = m.u; m.st.t0 = m.tfirst; m.st.nf = m.nfr; m.st.row0 = bd.row_off[m.u] + m.tfirst;
    m.st.g0 = bd.pcm_off[m.u] + (int64_t)m.tfirst * s;
    return m;
}

template <int WT, int ST>
__global__ void __launch_bounds__(SYN_THREADS, 2)
k_synth(const __grid_constant__ SynthParams S, int window, int wshift, float preem, int remove_dc, BatchDesc bd, int tile_frames, int ntiles,
        const int64_t *__restrict__ osamp_off, const int16_t *__restrict__ pcm, const float *__restrict__ spec, int16_t *__restrict__ out,
        const float2 *__restrict__ g_tw256, const float2 *__restrict__ g_twsplit, const float2 *__restrict__ g_twinv,
        const float *__restrict__ g_win) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int w = WT ? WT : window, s = ST ? ST : wshift, hh = S.hh;
    // carve-up
    cpx<float> *sTw = reinterpret_cast<cpx<float> *>(sm);                 // 256
    cpx<float> *sTs = sTw + 256;                                         // 130
    cpx<float> *sTi = sTs + 130;                                         // 130
    cpx<float> *sX = sTi + 130;                                          // SYN_GROUPS * 16*17
    float *sW = reinterpret_cast<float *>(sX + SYN_GROUPS * XPAD * 16);  // w (padded to 512)
    float *sD = sW + NFFT;                                               // (SYN_FRAMES-1)*s + w  (padded to 8)
    float *sYt = sD + (((SYN_FRAMES - 1) * s + w + 7) & ~7);             // SYN_FRAMES * w
    int16_t *raw = reinterpret_cast<int16_t *>(sYt + SYN_FRAMES * w);    // prefetch buffer, 8-sample chunks
    // persistent CTAs (two per SM): the next tile's PCM arrives by cp.async while this one is synthesised
    int tile = blockIdx.x;
    if (tile >= ntiles) return;
    SynTile cur = load_syn_tile(bd, tile, tile_frames, hh, s);
    int edge = 0;
    prefetch_pcm<SYN_THREADS>(raw, pcm, cur.st, (cur.nfr - 1) * s + w + 1, edge);
    for (int i = tid; i < 256; i += SYN_THREADS) sTw[i] = mk<float>(g_tw256[i].x, g_tw256[i].y);
    for (int i = tid; i < 129; i += SYN_THREADS) { sTs[i] = mk<float>(g_twsplit[i].x, g_twsplit[i].y); sTi[i] = mk<float>(g_twinv[i].x, g_twinv[i].y); }
    for (int i = tid; i < w; i += SYN_THREADS) sW[i] = g_win[i];
#pragma unroll 1
    for (; tile < ntiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    SynTile nxt = cur;
    if (next < ntiles) nxt = load_syn_tile(bd, next, tile_frames, hh, s);
    const int u = cur.u, t0 = cur.t0, T = cur.T, nf = cur.nf, tfirst = cur.tfirst, nfr = cur.nfr;
    finish_pcm<SYN_THREADS>(raw, sD, pcm, cur.st, (nfr - 1) * s + w + 1, edge, preem);
    __syncthreads();
    if (next < ntiles) prefetch_pcm<SYN_THREADS>(raw, pcm, nxt.st, (nxt.nfr - 1) * s + w + 1, edge);
    const int c = tid & (GROUP - 1), grp = tid / GROUP;
    cpx<float> *xch = sX + grp * (XPAD * 16);
    const float inv_w = 1.0f / (float)w;
    {
        const int f = grp;
        const bool active = f < nfr;
        if (active) {
            cpx<float> a[16], lo[8], hi[8], mid;
            float Alo[8], Ahi[8], Amid = 0.f;
            // enhanced magnitudes of this frame: issued first so that their HBM latency hides
            // behind the forward transform
            const float *srow = spec + (bd.row_off[u] + tfirst + f) * NBIN;
#pragma unroll
            for (int j = 0; j < 8; j++) { Alo[j] = __ldg(srow + c + 16 * j); Ahi[j] = __ldg(srow + NC - c - 16 * j); }
            Amid = __ldg(srow + 128);
            const float *d = sD + f * s;
            float sum = 0.f;
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) {
                int i0 = 32 * n1 + 2 * c;
                float y0 = (i0 < w) ? sW[i0] * d[i0] : 0.f;
                float y1 = (i0 + 1 < w) ? sW[i0 + 1] * d[i0 + 1] : 0.f;
                a[n1] = mk<float>(y0, y1);
                sum += y0 + y1;
            }
            if (remove_dc) {
                float mean = group_sum16(sum) * inv_w;
#pragma unroll
                for (int n1 = 0; n1 < 16; n1++) {
                    int i0 = 32 * n1 + 2 * c;
                    if (i0 < w) a[n1].x -= mean;
                    if (i0 + 1 < w) a[n1].y -= mean;
                }
            }
            const unsigned hm = 0xffffu << (tid & 16);
            fft256_pass1_rec(a, c, sTw, xch);
            __syncwarp(hm);
            fft256_pass2(a, c, xch);
            rfft_split_shfl(a, c, sTs, lo, hi, mid);
            // enhanced magnitude with the ORIGINAL phase: scale X by |X|enh / (|X| nfft);
            // bin 0 has phase 0, the Nyquist bin is always written non-negative
            // (src/io/out.cc:417-424)
            const float invn = 1.0f / (float)NFFT;
            auto scale_bin = [&](cpx<float> X, float A, bool edge) -> cpx<float> {
                A *= invn;
                if (edge) return mk<float>(A, 0.f);
                float m2 = X.x * X.x + X.y * X.y;
                if (m2 == 0.f) return mk<float>(0.f, -A);
                float g = A * rsqrtf(m2);
                return mk<float>(X.x * g, X.y * g);
            };
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const bool edge = (c == 0 && j == 0);           // k == 0 pairs with k == 256
                lo[j] = scale_bin(lo[j], Alo[j], edge);
                hi[j] = scale_bin(hi[j], Ahi[j], edge);
            }
            mid = scale_bin(mid, Amid, false);
            irfft_presplit_shfl(a, c, sTi, lo, hi, mid);
            __syncwarp(hm);                                     // pass-2 reads of the exchange tile are done
            fft256_pass1_rec(a, c, sTw, xch);
            __syncwarp(hm);
            fft256_pass2(a, c, xch);
            float *yt = sYt + f * w;
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                int n = c + 16 * k2;
                if (2 * n + 1 < w) *reinterpret_cast<float2 *>(yt + 2 * n) = make_float2(a[k2].x, -a[k2].y);
                else if (2 * n < w) yt[2 * n] = a[k2].x;
            }
        }
    }
    __syncthreads();
    // overlap-add in frame order (fp64 accumulator like the reference's cbuffer), quantise:
    // floor(x / correction), clip to +-32767 (src/io/out.cc:427-451)
    const bool last = (t0 + nf == T);
    const int nout = nf * s + (last ? (w - s) : 0);
    int16_t *o = out + osamp_off[u] + (int64_t)t0 * s;
    const int base = (t0 - tfirst) * s;                     // position of the tile's first sample inside the synthesised span
    const double inv_corr = 1.0 / S.correction;           // floor(x * (1/c)) differs from floor(x / c) only within 1e-16 of an integer
    for (int i = tid; i < nout; i += SYN_THREADS) {
        const int pos = base + i;                           // sample index relative to frame `tfirst`
        const int fa = (pos < w) ? 0 : (pos - w) / s + 1;   // first frame with fa*s + w > pos
        const int fb = min(pos / s, nfr - 1);
        double acc = 0.0;
        for (int f = fa; f <= fb; f++) acc += (double)sYt[f * w + (pos - f * s)];
        int v = (int)floor(acc * inv_corr);
        v = max(-32767, min(32767, v));
        o[i] = (int16_t)v;
    }
    __syncthreads();                                      // the slots and sD are re-used by the next tile
    cur = nxt;
    }
}

static inline size_t synth_smem_bytes(int w, int s) {
    size_t fl = 2 * (256 + 130 + 130 + SYN_GROUPS * XPAD * 16) + NFFT + (((SYN_FRAMES - 1) * s + w + 7) & ~7) + (size_t)SYN_FRAMES * w;
    return fl * sizeof(float) + (size_t)(((SYN_FRAMES - 1) * s + w + 1 + 8 + 7) / 8) * 16;
}

// tiles: the plan's synthesis tile list, tile_frames = SYN_FRAMES - hh hops per tile
static inline int launch_synth(const SynthParams &S, const FrameParams &F, const BatchDesc &bd, int tile_frames, int64_t ntiles,
                               const int64_t *d_osamp_off, const int16_t *pcm, const float *spec, int16_t *out, const float2 *tw,
                               const float2 *ts, const float2 *ti, const float *win, cudaStream_t s, LaunchCtx *lc, std::string &err) {
    if (ntiles <= 0) return CTU_OK;
    size_t bytes = synth_smem_bytes(F.window, F.wshift);
    if (bytes > 227 * 1024 || tile_frames < 1) { err = "CTU: window/shift combination needs too much shared memory for synthesis"; return CTU_ERR_UNSUPPORTED; }
    cudaError_t e;
    int per_sm = 1, num_sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    lc->begin("k_synth", s);
#define CTU_SYNTH_LAUNCH(WT, ST)                                                                                                   \
    e = cudaFuncSetAttribute(k_synth<WT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);                          \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_synth<WT, ST>, SYN_THREADS, bytes);         \
    if (e == cudaSuccess)                                                                                                          \
        k_synth<WT, ST><<<(unsigned)std::min<int64_t>(ntiles, (int64_t)std::max(per_sm, 1) * num_sms), SYN_THREADS, bytes, s>>>(  \
            S, F.window, F.wshift, F.preem, F.remove_dc, bd, tile_frames, (int)ntiles, d_osamp_off, pcm, spec, out, tw, ts, ti, win)
    if (F.window == 512 && F.wshift == 256) { CTU_SYNTH_LAUNCH(512, 256); }
    else if (F.window == 400 && F.wshift == 160) { CTU_SYNTH_LAUNCH(400, 160); }
    else { CTU_SYNTH_LAUNCH(0, 0); }
#undef CTU_SYNTH_LAUNCH
    lc->end(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e) + " (k_synth)"; return CTU_ERR_CUDA; }
    return CTU_OK;
}

// K9c: synthesis from a STORED complex spectrum.  k_frames2 can write X (float2 per bin) next to |X|; the synthesis then
// needs no second forward transform of the frame -- no PCM staging, no window, no forward FFT -- only the enhanced
// magnitude on the stored phase, the inverse transform and the overlap-add.  Same arithmetic as k_synth from scale_bin on
// (the stored X is the very value k_synth would recompute), so the waveforms are bit-identical; it trades 2 x 2056 B of
// HBM traffic per frame (write + read of X) for about 40 % of the synthesis kernel's instructions.
template <int WT, int ST>
__global__ void __launch_bounds__(SYN_THREADS, 3)
k_synth_c(const __grid_constant__ SynthParams S, int window, int wshift, BatchDesc bd, int tile_frames, int ntiles,
          const int64_t *__restrict__ osamp_off, const float2 *__restrict__ cspec, const float *__restrict__ spec, int16_t *__restrict__ out,
          const float2 *__restrict__ g_tw256, const float2 *__restrict__ g_twinv) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    const int w = WT ? WT : window, s = ST ? ST : wshift, hh = S.hh;
    cpx<float> *sTw = reinterpret_cast<cpx<float> *>(sm);                 // 256
    cpx<float> *sTi = sTw + 256;                                         // 130
    cpx<float> *sX = sTi + 130;                                          // SYN_GROUPS * 16*17
    float *sYt = reinterpret_cast<float *>(sX + SYN_GROUPS * XPAD * 16); // SYN_FRAMES * w
    for (int i = tid; i < 256; i += SYN_THREADS) sTw[i] = mk<float>(g_tw256[i].x, g_tw256[i].y);
    for (int i = tid; i < 129; i += SYN_THREADS) sTi[i] = mk<float>(g_twinv[i].x, g_twinv[i].y);
    __syncthreads();
    const int c = tid & (GROUP - 1), grp = tid / GROUP;
    const unsigned hm = 0xffffu << (tid & 16);
    cpx<float> *xch = sX + grp * (XPAD * 16);
    // (fetching a tile's inputs one tile ahead, before the overlap-add of the current one, was tried: 11.0 ms at two CTAs
    // per SM against 10.8 ms with three and no prefetch -- the kernel is bound by the transform and the overlap-add)
    struct SynMeta { int u, t0, T, nf, tfirst, nfr; };
    auto meta_of = [&](int tile) {
        SynMeta m;
        const int2 tl = bd.tiles[tile];
        m.u = tl.x; m.t0 = tl.y; m.T = bd.nframes[m.u];
        m.nf = min(tile_frames, m.T - m.t0); m.tfirst = max(m.t0 - hh, 0); m.nfr = m.t0 + m.nf - m.tfirst;
        return m;
    };
    cpx<float> lo[8], hi[8], mid;
    float Alo[8], Ahi[8], Amid = 0.f;
    auto fetch = [&](const SynMeta &m) {
        if (grp < m.nfr) {
            const int64_t row = bd.row_off[m.u] + m.tfirst + grp;
            const float *srow = spec + row * NBIN;
            const float2 *xrow = cspec + row * NBIN;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                Alo[j] = __ldg(srow + c + 16 * j); Ahi[j] = __ldg(srow + NC - c - 16 * j);
                const float2 xl = __ldg(xrow + c + 16 * j), xh = __ldg(xrow + NC - c - 16 * j);
                lo[j] = mk<float>(xl.x, xl.y); hi[j] = mk<float>(xh.x, xh.y);
            }
            Amid = __ldg(srow + 128);
            const float2 xm = __ldg(xrow + 128);
            mid = mk<float>(xm.x, xm.y);
        }
    };
    int tile = blockIdx.x;
    if (tile >= ntiles) return;
#pragma unroll 1
    for (; tile < ntiles; tile += gridDim.x) {
        const SynMeta cur = meta_of(tile);
        fetch(cur);
        const int u = cur.u, t0 = cur.t0, T = cur.T, nf = cur.nf, tfirst = cur.tfirst, nfr = cur.nfr;
        if (grp < nfr) {
            const int f = grp;
            cpx<float> a[16];
            // enhanced magnitude with the ORIGINAL phase: scale X by |X|enh / (|X| nfft); bin 0 has phase 0, the Nyquist
            // bin is always written non-negative (src/io/out.cc:417-424)
            const float invn = 1.0f / (float)NFFT;
            auto scale_bin = [&](cpx<float> X, float A, bool edge) -> cpx<float> {
                A *= invn;
                if (edge) return mk<float>(A, 0.f);
                float m2 = X.x * X.x + X.y * X.y;
                if (m2 == 0.f) return mk<float>(0.f, -A);
                float g = A * rsqrtf(m2);
                return mk<float>(X.x * g, X.y * g);
            };
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const bool edge = (c == 0 && j == 0);               // k == 0 pairs with k == 256
                lo[j] = scale_bin(lo[j], Alo[j], edge);
                hi[j] = scale_bin(hi[j], Ahi[j], edge);
            }
            mid = scale_bin(mid, Amid, false);
            irfft_presplit_shfl(a, c, sTi, lo, hi, mid);
            fft256_pass1_rec(a, c, sTw, xch);
            __syncwarp(hm);
            fft256_pass2(a, c, xch);
            __syncwarp(hm);
            float *yt = sYt + f * w;
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                int n = c + 16 * k2;
                if (2 * n + 1 < w) *reinterpret_cast<float2 *>(yt + 2 * n) = make_float2(a[k2].x, -a[k2].y);
                else if (2 * n < w) yt[2 * n] = a[k2].x;
            }
        }
        __syncthreads();
        // overlap-add in frame order (fp64 accumulator like the reference's cbuffer), floor(x / correction), clip
        // (src/io/out.cc:427-451)
        const bool last = (t0 + nf == T);
        const int nout = nf * s + (last ? (w - s) : 0);
        int16_t *o = out + osamp_off[u] + (int64_t)t0 * s;
        const int base = (t0 - tfirst) * s;
        const double inv_corr = 1.0 / S.correction;
        if (!((w | s) & 3) && !(reinterpret_cast<uintptr_t>(o) & 7)) {
            // four samples per thread where window and shift are multiples of four (16-byte slot reads, one 8-byte store)
            for (int i = 4 * tid; i < nout; i += 4 * SYN_THREADS) {
                const int pos = base + i;
                const int fa = (pos < w) ? 0 : (pos - w) / s + 1;
                const int fb = min(pos / s, nfr - 1);
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int f = fa; f <= fb; f++) {
                    const float4 y = *reinterpret_cast<const float4 *>(sYt + f * w + (pos - f * s));
                    a0 += (double)y.x; a1 += (double)y.y; a2 += (double)y.z; a3 += (double)y.w;
                }
                int v0 = (int)floor(a0 * inv_corr), v1 = (int)floor(a1 * inv_corr), v2 = (int)floor(a2 * inv_corr), v3 = (int)floor(a3 * inv_corr);
                v0 = max(-32767, min(32767, v0)); v1 = max(-32767, min(32767, v1));
                v2 = max(-32767, min(32767, v2)); v3 = max(-32767, min(32767, v3));
                uint2 pk;
                pk.x = (unsigned)(uint16_t)(int16_t)v0 | ((unsigned)(uint16_t)(int16_t)v1 << 16);
                pk.y = (unsigned)(uint16_t)(int16_t)v2 | ((unsigned)(uint16_t)(int16_t)v3 << 16);
                *reinterpret_cast<uint2 *>(o + i) = pk;
            }
        } else if (!((w | s) & 1) && !(reinterpret_cast<uintptr_t>(o) & 3)) {
            // two samples per thread: with an even window and shift an even-aligned pair has the same contributing frames,
            // so the index arithmetic is shared, the slots are read 8 bytes at a time and the pair leaves as one 32-bit store
            // (the overlap-add was ~45 % of this kernel's instructions)
            for (int i = 2 * tid; i < nout; i += 2 * SYN_THREADS) {
                const int pos = base + i;
                const int fa = (pos < w) ? 0 : (pos - w) / s + 1;
                const int fb = min(pos / s, nfr - 1);
                double a0 = 0.0, a1 = 0.0;
                for (int f = fa; f <= fb; f++) {
                    const float2 y = *reinterpret_cast<const float2 *>(sYt + f * w + (pos - f * s));
                    a0 += (double)y.x; a1 += (double)y.y;
                }
                int v0 = (int)floor(a0 * inv_corr), v1 = (int)floor(a1 * inv_corr);
                v0 = max(-32767, min(32767, v0)); v1 = max(-32767, min(32767, v1));
                *reinterpret_cast<unsigned *>(o + i) = (unsigned)(uint16_t)(int16_t)v0 | ((unsigned)(uint16_t)(int16_t)v1 << 16);
            }
        } else
        for (int i = tid; i < nout; i += SYN_THREADS) {
            const int pos = base + i;
            const int fa = (pos < w) ? 0 : (pos - w) / s + 1;
            const int fb = min(pos / s, nfr - 1);
            double acc = 0.0;
            for (int f = fa; f <= fb; f++) acc += (double)sYt[f * w + (pos - f * s)];
            int v = (int)floor(acc * inv_corr);
            v = max(-32767, min(32767, v));
            o[i] = (int16_t)v;
        }
        __syncthreads();                                      // the slots are re-used by the next tile
    }
}

static inline int launch_synth_c(const SynthParams &S, const FrameParams &F, const BatchDesc &bd, int tile_frames, int64_t ntiles,
                                 const int64_t *d_osamp_off, const float2 *cspec, const float *spec, int16_t *out, const float2 *tw,
                                 const float2 *ti, cudaStream_t s, LaunchCtx *lc, std::string &err) {
    if (ntiles <= 0) return CTU_OK;
    if (F.window & 1) { err = "CTU: synthesis needs an even window length"; return CTU_ERR_UNSUPPORTED; }
    const size_t bytes = sizeof(float) * (2 * (256 + 130 + SYN_GROUPS * XPAD * 16) + (size_t)SYN_FRAMES * F.window);
    if (bytes > 227 * 1024 || tile_frames < 1) { err = "CTU: window/shift combination needs too much shared memory for synthesis"; return CTU_ERR_UNSUPPORTED; }
    cudaError_t e;
    int per_sm = 1, num_sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    lc->begin("k_synth_c", s);
#define CTU_SYNTHC_LAUNCH(WT, ST)                                                                                                    \
    e = cudaFuncSetAttribute(k_synth_c<WT, ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);                          \
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_synth_c<WT, ST>, SYN_THREADS, bytes);         \
    if (e == cudaSuccess)                                                                                                            \
        k_synth_c<WT, ST><<<(unsigned)std::min<int64_t>(ntiles, (int64_t)std::max(per_sm, 1) * num_sms), SYN_THREADS, bytes, s>>>(  \
            S, F.window, F.wshift, bd, tile_frames, (int)ntiles, d_osamp_off, cspec, spec, out, tw, ti)
    if (F.window == 512 && F.wshift == 256) { CTU_SYNTHC_LAUNCH(512, 256); }
    else if (F.window == 400 && F.wshift == 160) { CTU_SYNTHC_LAUNCH(400, 160); }
    else { CTU_SYNTHC_LAUNCH(0, 0); }
#undef CTU_SYNTHC_LAUNCH
    lc->end(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e) + " (k_synth_c)"; return CTU_ERR_CUDA; }
    return CTU_OK;
}

static inline int launch_vad_module(VadParams V, const BurgParams &B, const BatchDesc &bd32, int64_t nt32, const int *d_nframes,
                                    const int64_t *d_row_off, int u0, int u1, int64_t row0, int64_t nrows, const int16_t *d_pcm,
                                    const float *d_spec, float *d_fea, const double *d_fea64, int fea_dim, double *d_ceps, double *d_cri,
                                    uint8_t *d_vad0,
                                    uint8_t *d_vadout, uint8_t *d_keep, int *d_rows, const double2 *tw, const double2 *ts,
                                    const double2 *ti, const double *win, cudaStream_t s, LaunchCtx *lc, std::string &err) {
    const int n = u1 - u0;
    if (n <= 0) return CTU_OK;
    cudaError_t e = cudaSuccess;
    if (V.cri == VCRI_ENERGY) {
        if (nrows > 0) {
            lc->begin("k_vad_energy", s);
            k_vad_energy<<<(unsigned)((nrows + 7) / 8), 256, 0, s>>>(V, d_spec, row0, nrows, d_cri);
            lc->end(s);
            e = cudaGetLastError();
        }
        V.cep_n = 0;
    } else if (V.cri == VCRI_CEPDIST_LPC) {
        V.cep_n = B.ncoef_vad;
        if (V.cep_n > BURG_MAXC || V.cep_n < 2) { err = "CTU: -vad_lpc_coefs must be 2..16"; return CTU_ERR_UNSUPPORTED; }
        int st = launch_burg(B, BURG_SRC_VAD, bd32, nt32, d_pcm, d_spec, d_ceps, tw, ts, ti, win, nullptr, s, lc, err);
        if (st) return st;
    } else {
        V.cep_n = fea_dim - V.has_E;      // the reference's vector does not hold the energy
        if (!d_fea64) { err = "CTU: internal: fp64 feature matrix missing for the cepstral-distance VAD"; return CTU_ERR_CONFIG; }
    }
    if (V.cep_n > 64) { err = "CTU: cepstral-distance VAD supports vectors of up to 64 values"; return CTU_ERR_UNSUPPORTED; }
    if (e == cudaSuccess) {
        lc->begin("k_vad_scan", s);
        const int cn_s = (V.cri == VCRI_ENERGY) ? 1 : V.cep_n;
        const size_t vs_bytes = (size_t)VADSCAN_WARPS * (32 * cn_s + 64 + 4) * sizeof(double);
        e = cudaFuncSetAttribute(k_vad_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs_bytes);
        k_vad_scan<<<(n + VADSCAN_WARPS - 1) / VADSCAN_WARPS, VADSCAN_WARPS * 32, vs_bytes, s>>>(V, d_nframes, d_row_off, u0, n, d_cri, d_ceps, d_fea64, fea_dim, d_vad0, d_vadout, d_keep);
        lc->end(s);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && V.drop) {
        size_t bytes = (size_t)32 * fea_dim * sizeof(float);
        e = cudaFuncSetAttribute(k_vad_compact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(bytes, 1024));
        if (e == cudaSuccess) {
            lc->begin("k_vad_compact", s);
            k_vad_compact<<<n, 256, bytes, s>>>(d_nframes, d_row_off, u0, d_keep, d_fea, fea_dim, d_rows);
            lc->end(s);
            e = cudaGetLastError();
        }
    }
    if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e) + " (vad module)"; return CTU_ERR_CUDA; }
    return CTU_OK;
}

}  // namespace ctu
#endif
