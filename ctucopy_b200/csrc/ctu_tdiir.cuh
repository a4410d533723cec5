// K10: -fea_kind td-iir-mfcc, compiled in its own translation unit (ctu_tdiir.cu).
//
// Replaces rawIN::compute_td_iir_mfcc (src/io/in.cc:281-303) and the td-iir branch of rawIN::get_frame (:317-340):
//   every sample goes through 24 fourth-order IIR band filters (second canonical form, coefficients from the `-filters`
//   file: b0..b4 | input gain | a1..a4), the filtered sample is multiplied by the Hamming weight of its POSITION IN THE
//   CIRCULAR BUFFER (n % window, not its position inside a frame), a frame's band energy is the sum of the squares of the
//   `window` values in the buffer, then log((w*w*E/w)/weight), bands in reverse order, 13-point DCT.
// fp64 throughout.  The filter state starts at zero for every utterance
// (= the first file of a reference list whose heap came back zeroed).
//
// Two kernels:
//   k_tdiir_filter   one thread per (utterance, band): the recurrence is sequential in time, the parallelism is
//                    utterances x 24 bands (240 000 threads for BASELINE's 10 000 utterances).  A CTA of 96 threads takes
//                    4 utterances; each of its three warps stages the samples of the two utterances its band threads belong
//                    to in a buffer of its own, in chunks, converted to fp64 once (coalesced loads; the band threads of an
//                    utterance then read the same address: a broadcast) -- no CTA barrier per chunk.  Each thread accumulates the
//                    windowed energy over segments of g = gcd(window, shift) samples and stores one double per segment:
//                    a frame is window / g whole segments, consecutive frames are shift / g segments apart.
//   k_tdiir_frames   one thread per frame: adds the frame's segments per band, log, DCT, writes the row in writer order
//                    (c1..c12, c0; src/io/out.cc:189-197).
// Algorithmic bytes per frame: 2 * shift (PCM) + 52 (row); the segment sums add 2 x 192 * shift / g bytes of workspace
// traffic.  The bound is the FP64 pipe: 11 DFMA / DMUL per sample and band.
#ifndef CTU_TDIIR_CUH
#define CTU_TDIIR_CUH

#include <cstdint>
#include <string>

#include "../../include/ctucopy_b200.h"
#include "ctu_kernels.cuh"

namespace ctu {

constexpr int TDIIR_BANDS = 24;
constexpr int TDIIR_NCEP = 13;
constexpr int TDIIR_UTTS = 4;                              // utterances per CTA
constexpr int TDIIR_THREADS = TDIIR_UTTS * TDIIR_BANDS;    // 96
constexpr int TDIIR_CHUNK = 512;                           // most samples staged per utterance and step
static_assert(TDIIR_THREADS == 96 && TDIIR_UTTS == TDIIR_THREADS / 32 + 1, "warp w holds band threads of utterances w and w + 1");

struct TdiirParams {
    int window, wshift, seg;       // seg = gcd(window, wshift)
    int chunk, run;                // samples staged per step (a multiple or a divisor of seg, <= TDIIR_CHUNK); run = min(seg, chunk)
    int spf, sps;                  // segments per frame / per shift
    double weight;                 // (double)(float) -weight_of_td_iir_mfcc_bank
    const double *coefs;           // [24][10]
    const double *win;             // [window] Hamming, fp64
    const double *dct;             // [13][24]: cos(pi * ((2k-1) i % 96) / 48), k = 1..24 (src/io/in.cc:236-237, 330-333)
};

// first segment of utterance u in the workspace: sum over v < u of ((T_v - 1) * sps + spf)
__host__ __device__ inline int64_t tdiir_seg_base(const TdiirParams &P, int64_t row_off_u, int u) {
    return row_off_u * P.sps + (int64_t)u * (P.spf - P.sps);
}

#ifdef CTU_TDIIR_IMPL

// Eight resident CTAs per SM (80 registers); 9 and 10 (72 / 64 registers) measured the same, 38.1-38.3 ms.
// (Loading the next chunk's samples to registers BEFORE the current chunk is filtered, so that the filter loop hides their
// latency, was measured too: 40.9 ms against 38.1 ms -- the 24 extra live registers cost more than the round trip.)
//
// WARP-PRIVATE staging: a warp holds the band threads of TWO utterances (warp w of the CTA: the last 24 - 8 w bands of
// utterance w and the first 8 + 8 w of utterance w + 1), stages those two utterances' samples in its own buffer and never
// meets the other warps again -- __syncwarp() instead of two CTA barriers per chunk.  With the CTA-wide staging 26 % of the
// stall samples sat at those barriers (profiles/r02_ncu_tdiir_full.txt): the three warps of a CTA run on different
// sub-partitions whose FP64 pipes are not equally loaded, and every chunk made the faster ones wait.  Costs 6 instead of 4
// utterance-chunks of conversions per CTA and chunk (a percent of the filter's work) and 24.7 instead of 16.5 KB.
constexpr int TDIIR_WU = 2;                                // utterances a warp touches
__global__ void __launch_bounds__(TDIIR_THREADS, 8)
k_tdiir_filter(const __grid_constant__ TdiirParams P, const int64_t *__restrict__ pcm_off, const int *__restrict__ nframes,
               const int64_t *__restrict__ row_off, int u0, int n_utts, const int16_t *__restrict__ pcm, double *__restrict__ S) {
    extern __shared__ __align__(16) double smd[];
    double *sWin = smd;                                                        // window
    __shared__ int64_t sOff[TDIIR_UTTS];
    __shared__ int sN[TDIIR_UTTS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, lu = tid / TDIIR_BANDS, b = tid % TDIIR_BANDS;
    const int L = P.chunk, run = P.run;
    const int LP = L + 2;                                  // row pitch: the rows of a warp's two utterances must not start in the same bank
    double *sX = sWin + P.window + wid * (TDIIR_WU * LP);  // [TDIIR_WU][LP]: this warp's samples as doubles
    for (int i = tid; i < P.window; i += TDIIR_THREADS) sWin[i] = P.win[i];
    if (tid < TDIIR_UTTS) {
        const int i = blockIdx.x * TDIIR_UTTS + tid;
        int n = 0;
        int64_t off = 0;
        if (i < n_utts) {
            const int T = nframes[u0 + i];
            n = T > 0 ? (T - 1) * P.wshift + P.window : 0;                     // samples the frames of this utterance cover
            off = pcm_off[u0 + i];
        }
        sN[tid] = n; sOff[tid] = off;
    }
    __syncthreads();                                       // the only CTA barrier
    const int myN = sN[lu];
    const int maxN = max(sN[wid], sN[wid + 1]);            // warp w: utterances w, w + 1 (TDIIR_UTTS = warps + 1)
    const int ui = blockIdx.x * TDIIR_UTTS + lu;
    const int u = u0 + min(ui, n_utts - 1);
    double *dst = S + tdiir_seg_base(P, row_off[u], u) * TDIIR_BANDS + b;
    const double *cf = P.coefs + b * 10;
    // the input gain g is folded into the numerator: with v' = v / g the recurrence reads v' = x - sum a_i s'_i,
    // y = (g b0) v' + sum (g b_i) s'_i -- one multiplication per sample and band fewer (the products g b_i are rounded once:
    // 1e-16 relative, like the order of the feedback sum below)
    const double g = cf[5];
    const double b0 = g * cf[0], b1 = g * cf[1], b2 = g * cf[2], b3 = g * cf[3], b4 = g * cf[4], a1 = cf[6], a2 = cf[7], a3 = cf[8], a4 = cf[9];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;       // state(ff, 0..3) / g: oldest .. newest (src/io/in.cc:289-297)
    double acc = 0.0;
    int wi = 0, gi = 0;                                   // n % window, n % seg at the start of a run
    const double *xs = sX + (lu - wid) * LP;
    // a chunk's samples: TDIIR_WU x PER independent 2-byte loads per lane, all in flight at once; converted once per
    // utterance and warp instead of once per band (I2F.F64 is a quarter-rate instruction)
    constexpr int PER = TDIIR_CHUNK / 32;
    for (int base = 0; base < maxN; base += L) {
#pragma unroll
        for (int k = 0; k < TDIIR_WU; k++) {                // half an utterance-chunk at a time: 8 live values (32 or 16 spill)
            const int16_t *src = pcm + sOff[wid + k] + base;
            const int nk = min(L, sN[wid + k] - base);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int16_t v16[PER / 2];
#pragma unroll
                for (int j = 0; j < PER / 2; j++) {
                    const int i = lane + (h * (PER / 2) + j) * 32;
                    v16[j] = (i < nk) ? src[i] : (int16_t)0;
                }
#pragma unroll
                for (int j = 0; j < PER / 2; j++) {
                    const int i = lane + (h * (PER / 2) + j) * 32;
                    if (i < L) sX[k * LP + i] = (double)v16[j];
                }
            }
        }
        __syncwarp();
        const int n = min(L, myN - base);                 // a multiple of `run` (<= 0: this lane's utterance has ended)
        // runs of `run` samples: a run lies inside one segment and never wraps around the window (run | seg | window), so
        // the sample and window pointers just advance
        for (int i0 = 0; i0 < n; i0 += run) {
            const double *xp = xs + i0, *wp = sWin + wi;
#pragma unroll 4
            for (int i = 0; i < run; i++) {
                // the feedback terms are taken oldest first, so that only ONE multiply-add waits for the newest state (the
                // reference subtracts a1 s3 first, src/io/in.cc:288-290: the same sum, rounded in another order -- 1e-16
                // relative in fp64, against a tolerance of 1e-3 in the log domain)
                double v = xp[i];
                v -= a4 * s0;
                v -= a3 * s1;
                v -= a2 * s2;
                v -= a1 * s3;
                double y = b0 * v;
                y += b4 * s0;
                y += b3 * s1;
                y += b2 * s2;
                y += b1 * s3;
                s0 = s1; s1 = s2; s2 = s3; s3 = v;
                y *= wp[i];
                acc = fma(y, y, acc);
            }
            wi += run; if (wi == P.window) wi = 0;
            gi += run;
            if (gi == P.seg) {
                gi = 0;
                if (ui < n_utts) *dst = acc;
                dst += TDIIR_BANDS;
                acc = 0.0;
            }
        }
        __syncwarp();                                      // the buffer is re-staged by the next iteration
    }
}

__global__ void __launch_bounds__(64)
k_tdiir_frames(const __grid_constant__ TdiirParams P, BatchDesc bd, const double *__restrict__ S, float *__restrict__ out, int out_stride) {
    __shared__ double sD[TDIIR_NCEP * TDIIR_BANDS];
    for (int i = threadIdx.x; i < TDIIR_NCEP * TDIIR_BANDS; i += 64) sD[i] = P.dct[i];
    __syncthreads();
    const int2 tile = bd.tiles[blockIdx.x];
    const int u = tile.x, t = tile.y + threadIdx.x;
    if (t >= bd.nframes[u]) return;
    const double *seg = S + (tdiir_seg_base(P, bd.row_off[u], u) + (int64_t)t * P.sps) * TDIIR_BANDS;
    double E[TDIIR_BANDS];
    const double ww = (double)(P.window * P.window), w1 = (double)P.window;
#pragma unroll
    for (int ff = 0; ff < TDIIR_BANDS; ff++) {
        double e = 0.0;
        for (int q = 0; q < P.spf; q++) e += seg[q * TDIIR_BANDS + ff];
        E[TDIIR_BANDS - 1 - ff] = log((ww * e / w1) / P.weight);               // src/io/in.cc:320-326
    }
    const double normcoef = sqrt(2.0 / TDIIR_BANDS);
    float *o = out + (bd.row_off[u] + t) * out_stride;
#pragma unroll 1
    for (int i = 0; i < TDIIR_NCEP; i++) {
        double c = 0.0;
#pragma unroll
        for (int k = 0; k < TDIIR_BANDS; k++) c += E[k] * sD[i * TDIIR_BANDS + k];
        o[i == 0 ? TDIIR_NCEP - 1 : i - 1] = (float)(c * normcoef);            // writer order: c1..c12, c0
    }
}

int launch_tdiir(const TdiirParams &P, const BatchDesc &bd64, int64_t nt64, const int64_t *d_pcm_off, const int *d_nframes,
                 const int64_t *d_row_off, int u0, int u1, const int16_t *pcm, double *S, float *out, int out_stride, cudaStream_t s,
                 LaunchCtx *lc, std::string &err) {
    const int n = u1 - u0;
    if (n <= 0 || nt64 <= 0) return CTU_OK;
    const size_t bytes = sizeof(double) * ((size_t)P.window + (size_t)(TDIIR_THREADS / 32) * TDIIR_WU * (P.chunk + 2));
    lc->begin("k_tdiir_filter", s);
    k_tdiir_filter<<<(unsigned)((n + TDIIR_UTTS - 1) / TDIIR_UTTS), TDIIR_THREADS, bytes, s>>>(P, d_pcm_off, d_nframes, d_row_off, u0, n, pcm, S);
    lc->end(s);
    lc->begin("k_tdiir_frames", s);
    k_tdiir_frames<<<(unsigned)nt64, 64, 0, s>>>(P, bd64, S, out, out_stride);
    lc->end(s);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("CUDA: ") + cudaGetErrorString(e) + " (k_tdiir)"; return CTU_ERR_CUDA; }
    return CTU_OK;
}

#else
int launch_tdiir(const TdiirParams &P, const BatchDesc &bd64, int64_t nt64, const int64_t *d_pcm_off, const int *d_nframes,
                 const int64_t *d_row_off, int u0, int u1, const int16_t *pcm, double *S, float *out, int out_stride, cudaStream_t s,
                 LaunchCtx *lc, std::string &err);
#endif

}  // namespace ctu
#endif
