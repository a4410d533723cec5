// K9g: synthesis for FFT sizes other than 512 (enhanced waveforms at 8 kHz: 256 points; 22-48 kHz: 1024 / 2048).
// Same chain as k_synth (sigOUT::save_frame, src/io/out.cc:374-451): the frame's own spectrum gives the phase, the
// noise-reduced magnitude replaces its modulus, inverse real FFT, overlap-add in frame order, floor(x / correction),
// clip.  Written for generality like k_frames_any -- one WARP per frame, radix-2 in shared memory -- and in fp64
// throughout: the specialised kernel reaches +-1 LSB with an fp32 16x16 transform, a 10-stage fp32 radix-2 would not
// stay there for loud frames.  Two kernels: frames -> time-domain segments in a workspace [frames x window] (fp64),
// then one thread per output sample adds the overlapping segments in frame order like the reference's ring.
#ifndef CTU_SYNTH_ANY_CUH
#define CTU_SYNTH_ANY_CUH

#include "ctu_frames_any.cuh"

namespace ctu {

struct AnyTables64 {
    const double2 *tw;       // e^{-2 pi i k / M}, k < M/2           (M = nfft/2)
    const double2 *twsplit;  // -i/2 e^{-2 pi i k / nfft}, k <= M
    const double *win;       // analysis window [window]
    int nfft, log2m;
};

// in-place radix-2 decimation-in-time FFT of M complex points held by one warp in shared memory
__device__ __forceinline__ void warp_fft_radix2(cpx<double> *z, int M, int log2m, const double2 *__restrict__ tw, int lane) {
    for (int n = lane; n < M; n += 32) {
        const int r = (int)(__brev((unsigned)n) >> (32 - log2m));
        if (r > n) { const cpx<double> t = z[n]; z[n] = z[r]; z[r] = t; }
    }
    __syncwarp();
    for (int len = 2, shift = log2m - 1; len <= M; len <<= 1, shift--) {
        const int half = len >> 1;
        for (int b = lane; b < (M >> 1); b += 32) {
            const int j0 = b & (half - 1), i0 = ((b - j0) << 1) + j0, i1 = i0 + half;
            const double2 w = __ldg(tw + ((size_t)j0 << shift));
            const cpx<double> t = cmul(z[i1], mk<double>(w.x, w.y));
            const cpx<double> a = z[i0];
            z[i0] = a + t;
            z[i1] = a - t;
        }
        __syncwarp();
    }
}

constexpr int SYNANY_THREADS = 128;           // 4 warps: 4 x (4M + 4) doubles of shared memory, 131 KB at 2048 points

__global__ void __launch_bounds__(SYNANY_THREADS)
k_synth_frames_any(int window, int wshift, double preem, int remove_dc, BatchDesc bd, AnyTables64 tb, const int16_t *__restrict__ pcm,
                   const float *__restrict__ spec, double *__restrict__ yt) {
    extern __shared__ __align__(16) double smd[];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    const int nfft = tb.nfft, M = nfft >> 1, nbins = M + 1;
    cpx<double> *z = reinterpret_cast<cpx<double> *>(smd + (size_t)wv * (2 * M + 2 * (M + 2)));   // M complex
    cpx<double> *Y = z + M;                                                                        // M + 1 complex (+ pad)
    const int2 tile = bd.tiles[blockIdx.x];
    const int u = tile.x, t0 = tile.y;
    const int nf = min(ANY_TILE, bd.nframes[u] - t0);
    const int64_t row0 = bd.row_off[u] + t0;
    const int w = window, s = wshift;
    for (int f = wv; f < nf; f += SYNANY_THREADS / 32) {
        // ---- analysis: the frame exactly as rawIN::get_frame builds it (src/io/in.cc:362-388) ---------------------
        const int16_t *x = pcm + bd.pcm_off[u] + (int64_t)(t0 + f) * s;
        const bool at_start = (t0 + f) == 0;
        double *y = reinterpret_cast<double *>(z);
        double sum = 0.0;
        for (int i = lane; i < nfft; i += 32) {
            double v = 0.0;
            if (i < w) {
                const double xi = (double)x[i];
                const double xp = (i == 0 && at_start) ? 0.0 : (double)x[i - 1];
                v = tb.win[i] * (xi - preem * xp);
            }
            y[i] = v;
            sum += v;
        }
        if (remove_dc) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const double mean = sum / (double)w;
            __syncwarp();
            for (int i = lane; i < w; i += 32) y[i] -= mean;
        }
        __syncwarp();
        warp_fft_radix2(z, M, tb.log2m, tb.tw, lane);
        // ---- real-input split; enhanced magnitude on the original phase, scaled by 1/nfft (src/io/out.cc:417-424) ---
        const float *A = spec + (row0 + f) * nbins;
        const double invn = 1.0 / (double)nfft;
        for (int k = lane; k <= M; k += 32) {
            const cpx<double> a = z[k == M ? 0 : k], b = conj(z[k == 0 ? 0 : M - k]);
            const double2 ts = __ldg(tb.twsplit + k);
            const cpx<double> X = mk<double>(0.5 * (a.x + b.x), 0.5 * (a.y + b.y)) + cmul(mk<double>(ts.x, ts.y), a - b);
            const double mag = (double)__ldg(A + k) * invn;
            cpx<double> Yk;
            if (k == 0 || k == M) Yk = mk<double>(mag, 0.0);            // bin 0: phase 0; Nyquist: written non-negative
            else {
                const double m2 = X.x * X.x + X.y * X.y;
                if (m2 == 0.0) Yk = mk<double>(0.0, -mag);
                else { const double g = mag / sqrt(m2); Yk = mk<double>(X.x * g, X.y * g); }
            }
            Y[k] = Yk;
        }
        __syncwarp();
        // ---- inverse: Z[k] = E[k] + i O[k], E = (Y[k] + conj Y[M-k]) / 2, O = e^{+2 pi i k/nfft} (Y[k] - conj Y[M-k]) / 2;
        //      z[n] = 2 sum_k Z[k] e^{+2 pi i k n / M} = 2 conj(FFT_M(conj Z))[n]; y[2n] = Re z[n], y[2n+1] = Im z[n] ----------
        for (int k = lane; k < M; k += 32) {
            const cpx<double> a = Y[k], b = conj(Y[M - k]);
            const double2 ts = __ldg(tb.twsplit + k);                  // (-sin/2, -cos/2)  ->  e^{+i th}/2 = (-ts.y, -ts.x)
            const cpx<double> E = mk<double>(0.5 * (a.x + b.x), 0.5 * (a.y + b.y));
            const cpx<double> O = cmul(mk<double>(-ts.y, -ts.x), a - b);
            z[k] = mk<double>(E.x - O.y, -(E.y + O.x));                // conj(Z[k])
        }
        __syncwarp();
        warp_fft_radix2(z, M, tb.log2m, tb.tw, lane);
        double *o = yt + (row0 + f) * w;
        for (int n = lane; n < M; n += 32) {
            const cpx<double> v = z[n];
            if (2 * n < w) o[2 * n] = 2.0 * v.x;
            if (2 * n + 1 < w) o[2 * n + 1] = -2.0 * v.y;
        }
        __syncwarp();
    }
}

// overlap-add in frame order with an fp64 accumulator (the reference's cbuffer), floor(x / correction), clip to +-32767
// (src/io/out.cc:427-451).  One CTA per 16-hop tile of the output (the last tile of a file also writes the window's tail).
__global__ void __launch_bounds__(256)
k_ola_any(int window, int wshift, double inv_corr, BatchDesc bd, const int64_t *__restrict__ osamp_off, const double *__restrict__ yt,
          int16_t *__restrict__ out) {
    const int2 tile = bd.tiles[blockIdx.x];
    const int u = tile.x, t0 = tile.y;
    const int T = bd.nframes[u];
    const int nf = min(ANY_TILE, T - t0);
    const int w = window, s = wshift;
    const int nout = nf * s + ((t0 + nf == T) ? (w - s) : 0);
    const double *y = yt + bd.row_off[u] * w;
    int16_t *o = out + osamp_off[u] + (int64_t)t0 * s;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        const int64_t pos = (int64_t)t0 * s + i;
        const int64_t fa = (pos < w) ? 0 : (pos - w) / s + 1;          // first frame with fa*s + w > pos
        const int64_t fb = min(pos / s, (int64_t)T - 1);
        double acc = 0.0;
        for (int64_t f = fa; f <= fb; f++) acc += y[f * w + (pos - f * s)];
        int v = (int)floor(acc * inv_corr);
        v = max(-32767, min(32767, v));
        o[i] = (int16_t)v;
    }
}

}  // namespace ctu
#endif
