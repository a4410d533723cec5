// K1g: the frame kernel for FFT sizes other than 512 (8 kHz telephone speech: 256 points;
// 22-48 kHz audio: 1024 / 2048 points).  Same chain, same arithmetic conventions as k_frames
// (rawIN::get_frame src/io/in.cc:305-419, FB::project_frame src/fea/fb.cc:72-86, specFEA /
// logspecFEA / dctcFEA src/fea/fea_impl.cc:37-131), written for generality rather than speed:
// one WARP per frame, the real transform as an in-place radix-2 complex FFT of half the length
// in shared memory, filter-bank weights and twiddles read from global memory (no size limit of
// the kernel-parameter block).  Every BASELINE configuration is 16 kHz / 512 points and takes
// the specialised kernels; this one exists so that other sampling rates work at all.
#ifndef CTU_FRAMES_ANY_CUH
#define CTU_FRAMES_ANY_CUH

#include "ctu_kernels.cuh"

namespace ctu {

constexpr int ANY_THREADS = 256;              // 8 warps = 8 frames in flight
constexpr int ANY_TILE = 16;                  // frames per CTA (the 16-frame tile list)
constexpr int ANY_MAX_NFFT = 2048;
constexpr int ANY_DC1_MAX = 18;               // frames a sample can be part of (+1), -remove_dc1

struct AnyTables {
    const float2 *tw;        // e^{-2 pi i k / M}, k < M/2           (M = nfft/2)
    const float2 *twsplit;   // -i/2 e^{-2 pi i k / nfft}, k <= M
    const float *win;        // analysis window [window]
    const float *fbw;        // packed filter-bank taps (scaled like FrameParams::w), bands 16-byte aligned
    const int4 *bands;       // [nb] {lo, ntaps, woff, 0}
    int nfft, log2m;
};

__host__ __device__ inline size_t any_smem_floats_per_warp(int nfft) { return (size_t)nfft /* M complex */ + (nfft / 2 + 4) /* bins */ + MAXB + 4; }

// -remove_dc1 (src/io/in.cc:343-350): mean of the ring at frame t, m_t = mean(raw frame t) - sum_d m_{t-d} (w - d s) / w
// (the overlap with frame t-d has already lost m_{t-d}).  One warp per utterance, frames in order, fp64.
__global__ void __launch_bounds__(128)
k_dc1_means(const int *__restrict__ nframes, const int64_t *__restrict__ row_off, const int64_t *__restrict__ pcm_off, int u0, int n_utts,
            int w, int s, const int16_t *__restrict__ pcm, double *__restrict__ m) {
    const int u = u0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (u >= u0 + n_utts) return;
    const int lane = threadIdx.x & 31;
    const int T = nframes[u];
    const int16_t *x = pcm + pcm_off[u];
    double *mu = m + row_off[u];
    for (int t = 0; t < T; t++) {
        int acc = 0;
        for (int i = lane; i < w; i += 32) acc += (int)x[(int64_t)t * s + i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
            double v = (double)acc / (double)w;
            for (int d = 1; d * s < w && d <= t; d++) v -= mu[t - d] * (double)(w - d * s) / (double)w;
            mu[t] = v;
        }
        __syncwarp();
    }
}

template <int SRC, int DST, int KIND>
__global__ void __launch_bounds__(ANY_THREADS)
k_frames_any(const __grid_constant__ FrameParams P, BatchDesc bd, AnyTables tb, const int16_t *__restrict__ pcm,
             const float *__restrict__ src, float *__restrict__ dst) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const int nfft = tb.nfft, M = nfft >> 1, nbins = M + 1;
    float *base = sm + (size_t)wv * any_smem_floats_per_warp(nfft);
    cpx<float> *z = reinterpret_cast<cpx<float> *>(base);              // M complex
    float *row = base + nfft;                                          // nbins (+pad)
    float *sY = row + (M + 4);                                         // MAXB (+pad)
    const int2 tile = bd.tiles[blockIdx.x];
    const int u = tile.x, t0 = tile.y;
    const int nf = min(ANY_TILE, bd.nframes[u] - t0);
    const int64_t row0 = bd.row_off[u] + t0;
    const int w = P.window, s = P.wshift, nb = P.nb;
    constexpr bool WANT_LOG = (DST == DST_FEA) && (KIND == KIND_DCTC || KIND == KIND_LOGSPEC || KIND == KIND_TRAPLOG);
    for (int f = wv; f < nf; f += ANY_THREADS / 32) {
        // ---- A: frame -> spectrum row ------------------------------------------------------------
        if (SRC == SRC_PCM) {
            const int16_t *x = pcm + bd.pcm_off[u] + (int64_t)(t0 + f) * s;
            const float *dn = P.dither ? P.dither + bd.pcm_off[u] + (int64_t)(t0 + f) * s : nullptr;
            const bool at_start = (t0 + f) == 0;
            float *y = reinterpret_cast<float *>(z);                   // the frame as nfft reals, natural order first
            float sum = 0.f;
            // -remove_dc1: frame t subtracts the mean of the whole ring from the ring, so a sample that is in frames
            // t-d .. t has lost m[t-d] + .. + m[t] by the time frame t reads it (m from k_dc1_means)
            float md[ANY_DC1_MAX];
            const int nd = P.dc1 ? min(ANY_DC1_MAX, w / s + 1) : 0;
            for (int d = 0; d < nd; d++) md[d] = (t0 + f - d >= 0) ? (float)P.dc1[row0 + f - d] : 0.f;
            auto ring_off = [&](int i, int d0) {        // what sample i of the frame has lost, counting frames t-d0, t-d0-1, ...
                float c = 0.f;
                for (int d = d0; d < nd; d++) if (i + (d - d0) * s < w) c += md[d];
                return c;
            };
            for (int i = lane; i < nfft; i += 32) {
                float v = 0.f;
                if (i < w) {
                    float xi = (float)x[i];
                    float xp = (i == 0 && at_start) ? 0.f : (float)x[i - 1];
                    if (dn) { xi += dn[i]; if (!(i == 0 && at_start)) xp += dn[i - 1]; }
                    if (P.dc1) {
                        xi -= ring_off(i, 0);
                        // the sample before the frame was remembered at the end of frame t-1 (src/io/in.cc:384)
                        if (i > 0) xp -= ring_off(i - 1, 0);
                        else if (!at_start) xp -= ring_off(s - 1, 1);
                    }
                    v = tb.win[i] * fmaf(-P.preem, xp, xi);
                }
                y[i] = v;
                sum += v;
            }
            if (P.remove_dc) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float mean = sum / (float)w;                    // of the windowed frame, subtracted inside the window only
                __syncwarp();
                for (int i = lane; i < w; i += 32) y[i] -= mean;
            }
            __syncwarp();
            // bit-reversal permutation of the M complex points (z[n] = y[2n] + i y[2n+1])
            for (int n = lane; n < M; n += 32) {
                const int r = (int)(__brev((unsigned)n) >> (32 - tb.log2m));
                if (r > n) { const cpx<float> t = z[n]; z[n] = z[r]; z[r] = t; }
            }
            __syncwarp();
            // radix-2 decimation in time
            for (int len = 2, shift = tb.log2m - 1; len <= M; len <<= 1, shift--) {
                const int half = len >> 1;
                for (int b = lane; b < (M >> 1); b += 32) {
                    const int j0 = b & (half - 1), i0 = ((b - j0) << 1) + j0, i1 = i0 + half;
                    const float2 tw = __ldg(tb.tw + ((size_t)j0 << shift));
                    const cpx<float> t = cmul(z[i1], mk<float>(tw.x, tw.y));
                    const cpx<float> a = z[i0];
                    z[i0] = a + t;
                    z[i1] = a - t;
                }
                __syncwarp();
            }
            // real-input split, power / magnitude
            for (int k = lane; k <= M; k += 32) {
                const cpx<float> A = z[k == M ? 0 : k], B = conj(z[k == 0 ? 0 : M - k]);
                const float2 ts = __ldg(tb.twsplit + k);
                const cpx<float> X = mk<float>(0.5f * (A.x + B.x), 0.5f * (A.y + B.y)) + cmul(mk<float>(ts.x, ts.y), A - B);
                float p = X.x * X.x + X.y * X.y;
                if (k == 0 && P.remove_dc) p = 1e-10f;                // fixed floor (src/io/in.cc:390)
                row[k] = P.take_sqrt ? sqrtf(p) : p;
            }
        } else if (SRC == SRC_SPEC) {
            const float *g = src + (row0 + f) * nbins;
            for (int k = lane; k < nbins; k += 32) row[k] = g[k];
        } else {
            const float *g = src + (row0 + f) * nb;
            for (int b = lane; b < nb; b += 32) { float v = g[b]; if (WANT_LOG) v = logf(v); sY[b] = v; }
        }
        __syncwarp();
        if (DST == DST_SPEC) {
            float *g = dst + (row0 + f) * nbins;
            for (int k = lane; k < nbins; k += 32) g[k] = row[k];
            __syncwarp();
            continue;
        }
        if (SRC != SRC_FB && (P.energy_mode == EN_NR || P.energy_mode == EN_IN)) {
            const float e = half_spectrum_energy(row, nbins, (P.energy_mode == EN_NR) || P.take_sqrt);
            if (lane == 0) P.energy[row0 + f] = e;
        }
        // ---- B: filter bank, one band per lane -----------------------------------------------------
        const bool late_log = (KIND == KIND_LOGSPEC && P.energy_mode == EN_BANDS && DST == DST_FEA);
        if (SRC != SRC_FB) {
            for (int b = lane; b < nb; b += 32) {
                const int4 bs = __ldg(tb.bands + b);
                const float *r = row + bs.x;
                const float *wq = tb.fbw + bs.z;
                float acc = 0.f;
                for (int k = 0; k < bs.y; k++) acc = fmaf(r[k], __ldg(wq + k), acc);   // the reference's summation order
                const bool lg = WANT_LOG && !late_log;
                float yv;
                if (P.inld) { yv = powf(acc, 0.33f) * P.inld_scale; if (lg) yv = logf(yv); }
                else yv = lg ? logf(acc) + P.log_offset : acc * P.lin_scale;
                sY[b] = yv;
            }
            for (int b = nb + lane; b < P.nbp; b += 32) sY[b] = 0.f;
            __syncwarp();
        }
        if (DST == DST_FB) {
            float *g = dst + (row0 + f) * nb;
            for (int b = lane; b < nb; b += 32) g[b] = sY[b];
            __syncwarp();
            continue;
        }
        // ---- C: features ---------------------------------------------------------------------------
        float *g = dst + (row0 + f) * P.out_stride;
        if (KIND == KIND_DCTC) {
            for (int i = lane; i < P.nrows; i += 32) {
                const float *m = P.m2 + i * P.nbp;
                float acc = 0.f;
                for (int k = 0; k < nb; k++) acc = fmaf(sY[k], m[k], acc);
                g[i] = acc;
            }
        } else {
            if (P.energy_mode == EN_BANDS && lane == 0) {
                double acc = 0.5 * (double)sY[0] * (double)sY[0];
                for (int b = 1; b < nb - 1; b++) acc += (double)sY[b] * (double)sY[b];
                acc += 0.5 * (double)sY[nb - 1] * (double)sY[nb - 1];
                P.energy[row0 + f] = (float)log(acc * 2.0);
            }
            for (int b = lane; b < nb; b += 32) g[b] = late_log ? logf(sY[b]) : sY[b];
        }
        __syncwarp();
    }
}

}  // namespace ctu
#endif
