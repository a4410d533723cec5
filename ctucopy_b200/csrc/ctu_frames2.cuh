// K1 (second generation): the fused frame kernel with a GROUP-LOCAL pipeline.
//
// A CTA of 128 threads = 8 groups of 16 threads works on a tile of F2_TILE consecutive frames
// of one utterance.  The CTA synchronises once, after staging the tile's pre-emphasised
// samples and the (small) tables in shared memory; from then on each group takes a frame
// through the whole chain on its own, synchronising only inside its half warp:
//   window, DC removal -> 256-point complex FFT (16 points per thread, ctu_fft.cuh) -> real
//   split -> power / magnitude row in the group's exchange area -> filter bank (bands dealt
//   to the 16 lanes, widest first, taps summed in the reference's order) -> log -> DCT +
//   lifter rows dealt to the lanes -> the frame's feature row straight to HBM.
// No [frames x 257] spectrum tile, no CTA-wide barrier between phases: 37 KB of shared memory
// per CTA instead of 113 KB, six resident CTAs (24 warps) per SM instead of two (16 warps).
// Replaces rawIN::get_frame (src/io/in.cc:305-419), FB::project_frame (src/fea/fb.cc:72-86),
// specFEA / logspecFEA / dctcFEA (src/fea/fea_impl.cc:37-131).
#ifndef CTU_FRAMES2_CUH
#define CTU_FRAMES2_CUH

#include "ctu_kernels.cuh"

namespace ctu {

constexpr int F2_THREADS = 128;
constexpr int F2_GROUPS = F2_THREADS / GROUP;   // 8 frames in flight per CTA
constexpr int F2_TILE = 16;                     // frames per CTA: two passes
constexpr int F2_ROWF = 2 * XPAD * 16 + 16;     // floats between group areas: 544 used + 16, so that the two groups of a warp
                                                // sit half a bank row apart and their 32-bit accesses share one wavefront
constexpr int F2_YOFF = 260;                    // band values of the frame inside that area (16-byte aligned, after the 257 bins)

// tables in global memory (built once per handle), copied to shared memory by every CTA
struct Tables2 {
    const float2 *tw256, *twsplit;
    const float *win;
    const float *fbw;        // packed filter-bank taps, every band starts 16-byte aligned
    const int4 *slots;       // [nslots] {band, lo, ntaps, woff}; band = -1: empty slot.  Slot k is served by lane k % 16
    const float *m2;         // [nrows][m2_pitch] second-stage matrix (pitch odd: conflict-free across lanes)
    int ntaps_total, nslots, m2_pitch, m2_rows;
};

struct Smem2 {
    int oX, oTw, oTs, oW, oFbw, oSlots, oM2, oD, total;      // float offsets
};
__host__ __device__ inline Smem2 smem2_layout(int window, int wshift, int ntaps_total, int nslots, int m2_pitch, int m2_rows) {
    Smem2 L;
    int o = 0;
    L.oX = o; o += F2_GROUPS * F2_ROWF;
    L.oTw = o; o += 256 * 2;
    L.oTs = o; o += 130 * 2;
    L.oW = o; o += (window + 3) & ~3;
    L.oFbw = o; o += (ntaps_total + 3) & ~3;
    L.oSlots = o; o += nslots * 4;
    L.oM2 = o; o += (m2_rows * m2_pitch + 3) & ~3;
    L.oD = o; o += ((F2_TILE - 1) * wshift + window + 3) & ~3;
    L.total = o;
    return L;
}

template <int SRC, int DST, int KIND, int WT>
__global__ void __launch_bounds__(F2_THREADS, 6)
k_frames2(const __grid_constant__ FrameParams P, BatchDesc bd, Tables2 tb, const int16_t *__restrict__ pcm,
          const float *__restrict__ src, float *__restrict__ dst) {
    extern __shared__ __align__(16) float sm[];
    const Smem2 L = smem2_layout(P.window, P.wshift, tb.ntaps_total, tb.nslots, tb.m2_pitch, tb.m2_rows);
    const int tid = threadIdx.x;
    const int2 tile = bd.tiles[blockIdx.x];
    const int u = tile.x, t0 = tile.y;
    const int nf = min(F2_TILE, bd.nframes[u] - t0);
    const int64_t row0 = bd.row_off[u] + t0;
    const int w = WT ? WT : P.window, s = P.wshift;
    float *sD = sm + L.oD, *sW = sm + L.oW, *sFbw = sm + L.oFbw, *sM2 = sm + L.oM2;
    int4 *sSlots = reinterpret_cast<int4 *>(sm + L.oSlots);
    cpx<float> *sTw = reinterpret_cast<cpx<float> *>(sm + L.oTw);
    cpx<float> *sTs = reinterpret_cast<cpx<float> *>(sm + L.oTs);

    // ---- stage (the only CTA-wide phase) -------------------------------------------------------
    if (SRC == SRC_PCM) {
        stage_preem<F2_THREADS>(sD, pcm + bd.pcm_off[u] + (int64_t)t0 * s, (nf - 1) * s + w, t0 == 0, P.preem);
        for (int i = tid; i < w; i += F2_THREADS) sW[i] = tb.win[i];
        for (int i = tid; i < 256; i += F2_THREADS) sTw[i] = mk<float>(tb.tw256[i].x, tb.tw256[i].y);
        for (int i = tid; i < 129; i += F2_THREADS) sTs[i] = mk<float>(tb.twsplit[i].x, tb.twsplit[i].y);
    }
    if (DST != DST_SPEC) {
        if (SRC != SRC_FB) {
            for (int i = tid; i < tb.ntaps_total; i += F2_THREADS) sFbw[i] = tb.fbw[i];
            for (int i = tid; i < tb.nslots; i += F2_THREADS) sSlots[i] = tb.slots[i];
        }
        if (DST == DST_FEA && KIND == KIND_DCTC)
            for (int i = tid; i < tb.m2_rows * tb.m2_pitch; i += F2_THREADS) sM2[i] = tb.m2[i];
    }
    __syncthreads();

    const int c = tid & (GROUP - 1), grp = tid / GROUP;
    const unsigned hm = 0xffffu << (tid & 16);                // the two groups of a warp run independently
    float *area = sm + L.oX + grp * F2_ROWF;                  // exchange tile, then spectrum row + band values
    cpx<float> *xch = reinterpret_cast<cpx<float> *>(area);
    float *sPr = area, *sY = area + F2_YOFF;
    const float inv_w = 1.0f / (float)w;
    const int nb = P.nb;
    constexpr bool WANT_LOG = (DST == DST_FEA) && (KIND == KIND_DCTC || KIND == KIND_LOGSPEC || KIND == KIND_TRAPLOG);
#pragma unroll 1
    for (int f = grp; f < nf; f += F2_GROUPS) {
        // ---- A: frame -> spectrum row ----------------------------------------------------------
        if (SRC == SRC_PCM) {
            cpx<float> a[16];
            const float *d = sD + f * s;
            float sum = 0.f;
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) {
                const int i0 = 32 * n1 + 2 * c;
                float y0 = (i0 < w) ? sW[i0] * d[i0] : 0.f;
                float y1 = (i0 + 1 < w) ? sW[i0 + 1] * d[i0 + 1] : 0.f;
                a[n1] = mk<float>(y0, y1);
                sum += y0 + y1;
            }
            if (P.remove_dc) {
                // mean of the WINDOWED frame, subtracted from the window's samples only (src/io/in.cc:375-382)
                const float mean = group_sum16(sum) * inv_w;
#pragma unroll
                for (int n1 = 0; n1 < 16; n1++) {
                    const int i0 = 32 * n1 + 2 * c;
                    if (i0 < w) a[n1].x -= mean;
                    if (i0 + 1 < w) a[n1].y -= mean;
                }
            }
            fft256_pass1_rec(a, c, sTw, xch);
            __syncwarp(hm);
            fft256_pass2(a, c, xch);
            cpx<float> lo[8], hi[8], mid;
            rfft_split_shfl(a, c, sTs, lo, hi, mid);
            __syncwarp(hm);                                     // every lane holds its bins: the area can be overwritten
            // PCM -> spectrum: the row goes from the registers straight to HBM (each store covers 64 contiguous bytes)
            float *grow = (DST == DST_SPEC) ? dst + (row0 + f) * NBIN : sPr;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int k = c + 16 * j;
                float pl = lo[j].x * lo[j].x + lo[j].y * lo[j].y;
                float ph = hi[j].x * hi[j].x + hi[j].y * hi[j].y;
                if (k == 0 && P.remove_dc) pl = 1e-10f;       // fixed floor (src/io/in.cc:390)
                if (P.take_sqrt) { pl = sqrtf(pl); ph = sqrtf(ph); }
                grow[k] = pl;
                grow[NC - k] = ph;
            }
            if (c == 0) {
                const float pm = mid.x * mid.x + mid.y * mid.y;
                grow[128] = P.take_sqrt ? sqrtf(pm) : pm;
            }
            if (DST == DST_SPEC) continue;
        } else if (SRC == SRC_SPEC) {
            const float *g = src + (row0 + f) * NBIN;
#pragma unroll
            for (int j = 0; j < 17; j++) {
                const int k = c + 16 * j;
                if (k < NBIN) sPr[k] = g[k];
            }
        } else {  // SRC_FB: band values (true scale, post ^0.33)
            const float *g = src + (row0 + f) * nb;
            for (int b = c; b < nb; b += GROUP) {
                float v = g[b];
                if (WANT_LOG) v = logf(v);
                sY[b] = v;
            }
        }
        __syncwarp(hm);
        if (DST == DST_SPEC) {
            float *g = dst + (row0 + f) * NBIN;
#pragma unroll
            for (int j = 0; j < 17; j++) {
                const int k = c + 16 * j;
                if (k < NBIN) g[k] = sPr[k];
            }
            __syncwarp(hm);
            continue;
        }
        // ---- B: filter bank, bands dealt to the lanes (src/fea/fb.cc:72-86) ----------------------
        if (SRC != SRC_FB) {
            for (int sl = c; sl < tb.nslots; sl += GROUP) {
                const int4 bs = sSlots[sl];                   // {band, lo, ntaps, woff}
                if (bs.x < 0) continue;
                const float *r = sPr + bs.y;
                const float *wq = sFbw + bs.w;
                const int n = bs.z, n4 = n & ~3;
                float acc = 0.f;
                int k = 0;
                for (; k < n4; k += 4) {                      // same summation order as the reference loop
                    const float4 w4 = *reinterpret_cast<const float4 *>(wq + k);
                    acc = fmaf(r[k], w4.x, acc);
                    acc = fmaf(r[k + 1], w4.y, acc);
                    acc = fmaf(r[k + 2], w4.z, acc);
                    acc = fmaf(r[k + 3], w4.w, acc);
                }
                for (; k < n; k++) acc = fmaf(r[k], wq[k], acc);
                float y;
                if (P.inld) {
                    y = powf(acc, 0.33f) * P.inld_scale;
                    if (WANT_LOG) y = logf(y);
                } else {
                    y = WANT_LOG ? logf(acc) + P.log_offset : acc * P.lin_scale;
                }
                sY[bs.x] = y;
            }
            if (c < P.nbp - nb) sY[nb + c] = 0.f;             // pad columns read by the 4-wide loop below
            __syncwarp(hm);
        }
        // ---- C: feature transform, rows dealt to the lanes ---------------------------------------
        float *g = dst + (row0 + f) * P.out_stride;
        if (DST == DST_FB || KIND != KIND_DCTC) {
            for (int b = c; b < nb; b += GROUP) g[b] = sY[b];
        } else {
            // c[i] = sum_k ln Y_k m2[i][k] (src/fea/fea_impl.cc:81-131; norm, lifter and writer order folded in on the host)
            const int nbp = P.nbp;
            for (int i = c; i < P.nrows; i += GROUP) {
                const float *m = sM2 + i * tb.m2_pitch;
                float acc = 0.f;
                for (int k = 0; k < nbp; k += 4) {            // pad taps are zero, pad band slots hold finite leftovers * 0
                    const float4 y4 = *reinterpret_cast<const float4 *>(sY + k);
                    acc = fmaf(y4.x, m[k], acc);
                    acc = fmaf(y4.y, m[k + 1], acc);
                    acc = fmaf(y4.z, m[k + 2], acc);
                    acc = fmaf(y4.w, m[k + 3], acc);
                }
                g[i] = acc;
            }
        }
        __syncwarp(hm);
    }
}

}  // namespace ctu
#endif
