// Host <-> device copy ceiling of a box, measured the way the end-to-end entry point copies: chunked cudaMemcpyAsync
// of a 3.2 GB input (H2D) and a 1.56 GB result (D2H) per "step" over separate streams, with N processes (one per GPU)
// running at the same time.  Answers whose limit the flat 1 -> 8 GPU end-to-end curve is: the box's or the
// code's.  Not part of the product; built by tools/copy_probe.sh with nvcc.
//
//   copy_probe <gpu> <start_epoch_s> <seconds> <alloc> <chunk_mb> <dirs>
//     alloc: pinned | wc (write-combined upload buffer) | reg (malloc + cudaHostRegister) | huge (2 MB huge pages via
//            mmap(MAP_HUGETLB) + cudaHostRegister; falls back to THP-advised memory when no huge pages are reserved)
//     dirs : both | h2d | d2h
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static double now() { timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + 1e-6 * tv.tv_usec; }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); return 2; } } while (0)

static void *alloc_host(const std::string &kind, size_t bytes, bool upload, std::string &how) {
    void *p = nullptr;
    if (kind == "pinned" || (kind == "wc" && !upload)) {
        if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
        how = "cudaHostAlloc";
    } else if (kind == "wc") {
        if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) return nullptr;
        how = "cudaHostAlloc(WriteCombined)";
    } else if (kind == "reg") {
        if (posix_memalign(&p, 1 << 21, bytes)) return nullptr;
        memset(p, 1, bytes);
        if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) != cudaSuccess) return nullptr;
        how = "posix_memalign + cudaHostRegister";
    } else {  // huge
        const size_t hb = (bytes + (1 << 21) - 1) & ~(size_t)((1 << 21) - 1);
        p = mmap(nullptr, hb, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
        how = "mmap(MAP_HUGETLB) + cudaHostRegister";
        if (p == MAP_FAILED) {
            p = mmap(nullptr, hb, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (p == MAP_FAILED) return nullptr;
            madvise(p, hb, MADV_HUGEPAGE);
            how = "mmap + MADV_HUGEPAGE + cudaHostRegister (no reserved huge pages)";
        }
        memset(p, 1, hb);
        if (cudaHostRegister(p, hb, cudaHostRegisterDefault) != cudaSuccess) return nullptr;
    }
    return p;
}

int main(int argc, char **argv) {
    if (argc < 7) { fprintf(stderr, "usage: copy_probe gpu start_epoch seconds alloc chunk_mb dirs\n"); return 1; }
    const int gpu = atoi(argv[1]);
    const double start = atof(argv[2]), secs = atof(argv[3]);
    const std::string kind = argv[4], dirs = argv[6];
    const size_t chunk = (size_t)atoi(argv[5]) << 20;
    const size_t in_bytes = (size_t)10000 * 160000 * 2, out_bytes = (size_t)9980000 * 39 * 4;
    CK(cudaSetDevice(gpu));
    std::string how_in, how_out;
    char *h_in = (char *)alloc_host(kind, in_bytes, true, how_in), *h_out = (char *)alloc_host(kind, out_bytes, false, how_out);
    if (!h_in || !h_out) { fprintf(stderr, "host allocation (%s) failed\n", kind.c_str()); return 2; }
    char *d_in, *d_out;
    CK(cudaMalloc(&d_in, in_bytes)); CK(cudaMalloc(&d_out, out_bytes));
    cudaStream_t s_in, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));
    const bool do_in = dirs != "d2h", do_out = dirs != "h2d";
    // warm-up step, then wait for the common start time so that all ranks copy at once
    if (do_in) CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, s_in));
    if (do_out) CK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s_out));
    CK(cudaDeviceSynchronize());
    while (now() < start) usleep(200);
    const double t0 = now();
    int steps = 0;
    // the output is 0.49 of the input: chunks of both directions are issued in that proportion, like the pipeline does
    const size_t ochunk = (size_t)((double)chunk * out_bytes / in_bytes) & ~(size_t)255;
    while (now() - t0 < secs) {
        size_t oi = 0, oo = 0;
        while ((do_in && oi < in_bytes) || (do_out && oo < out_bytes)) {
            if (do_in && oi < in_bytes) { size_t n = std::min(chunk, in_bytes - oi); CK(cudaMemcpyAsync(d_in + oi, h_in + oi, n, cudaMemcpyHostToDevice, s_in)); oi += n; }
            if (do_out && oo < out_bytes) { size_t n = std::min(ochunk, out_bytes - oo); CK(cudaMemcpyAsync(h_out + oo, d_out + oo, n, cudaMemcpyDeviceToHost, s_out)); oo += n; }
        }
        CK(cudaDeviceSynchronize());
        steps++;
    }
    const double dt = now() - t0;
    const double gb_in = do_in ? steps * (double)in_bytes / 1e9 : 0, gb_out = do_out ? steps * (double)out_bytes / 1e9 : 0;
    printf("{\"concurrent\": %d, \"gpu\": %d, \"alloc\": \"%s\", \"how\": \"%s\", \"dirs\": \"%s\", \"chunk_mb\": %d, \"steps\": %d, \"seconds\": %.3f, \"ms_per_step\": %.2f, "
           "\"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"total_gbs\": %.2f}\n",
           argc > 7 ? atoi(argv[7]) : 1, gpu, kind.c_str(), how_in.c_str(), dirs.c_str(), atoi(argv[5]), steps, dt, 1000 * dt / steps, gb_in / dt, gb_out / dt, (gb_in + gb_out) / dt);
    return 0;
}
