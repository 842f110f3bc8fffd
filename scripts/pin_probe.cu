// Where does the first-call time go?  Times cudaHostRegister of a fresh 29 GB malloc under several preparations.
// Build: nvcc -O2 -o gpurun_out/pin_probe scripts/pin_probe.cu -lpthread   Run on the GPU box: gpurun_out/pin_probe [GB]
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static void cat(const char *p) { FILE *f = fopen(p, "r"); if (!f) { printf("%s: n/a\n", p); return; } char b[256]; if (fgets(b, 256, f)) printf("%s: %s", p, b); fclose(f); }
static void touch(char *q, size_t bytes, int nt)
{
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([=] { size_t lo = bytes / nt * t, hi = t == nt - 1 ? bytes : bytes / nt * (t + 1); for (size_t o = lo; o < hi; o += 4096) q[o] = 0; });
    for (auto &t : th) t.join();
}
int main(int argc, char **argv)
{
    size_t bytes = (size_t) (argc > 1 ? atof(argv[1]) : 29.0) * (1ull << 30);
    cat("/sys/kernel/mm/transparent_hugepage/enabled"); cat("/sys/kernel/mm/transparent_hugepage/defrag");
    cudaFree(0);
    double t0, t1, t2;
    {   // 1: plain
        char *p = (char *) malloc(bytes); t0 = now(); cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault); t1 = now();
        printf("1 fresh malloc, register whole: %.2f s (%s)\n", t1 - t0, cudaGetErrorString(e)); cudaHostUnregister(p);
        t0 = now(); e = cudaHostRegister(p, bytes, cudaHostRegisterDefault); t1 = now();
        printf("1b same memory again (pages resident): %.2f s\n", t1 - t0); cudaHostUnregister(p); free(p);
    }
    {   // 2: parallel touch then register
        char *p = (char *) malloc(bytes); t0 = now(); touch(p, bytes, 16); t1 = now(); cudaHostRegister(p, bytes, cudaHostRegisterDefault); t2 = now();
        printf("2 touch x16: %.2f s, then register: %.2f s\n", t1 - t0, t2 - t1); cudaHostUnregister(p); free(p);
    }
    {   // 3: registration of disjoint chunks from 16 threads
        char *p = (char *) malloc(bytes); int nt = 16; size_t ch = (bytes / nt) & ~(size_t) 4095; char *a = (char *) (((size_t) p + 4095) & ~(size_t) 4095);
        t0 = now(); std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back([=] { cudaSetDevice(0); cudaHostRegister(a + ch * t, ch, cudaHostRegisterDefault); });
        for (auto &t : th) t.join(); t1 = now();
        printf("3 register 16 chunks from 16 threads: %.2f s\n", t1 - t0);
        for (int t = 0; t < nt; t++) cudaHostUnregister(a + ch * t); free(p);
    }
    {   // 4: huge pages by madvise, then register
        char *p = (char *) malloc(bytes); char *a = (char *) (((size_t) p + (2u << 20) - 1) & ~(size_t) ((2u << 20) - 1)); size_t len = (bytes - (a - p)) & ~(size_t) ((2u << 20) - 1);
        int r = madvise(a, len, MADV_HUGEPAGE); t0 = now(); cudaHostRegister(p, bytes, cudaHostRegisterDefault); t1 = now();
        printf("4 madvise(HUGEPAGE)=%d, register whole: %.2f s\n", r, t1 - t0); cudaHostUnregister(p); free(p);
    }
    {   // 5: driver-allocated pinned memory and a staged copy into fresh pageable memory
        char *h; t0 = now(); cudaError_t e = cudaMallocHost((void **) &h, (size_t) 1 << 30); t1 = now(); printf("5 cudaMallocHost 1 GB: %.2f s (%s)\n", t1 - t0, cudaGetErrorString(e));
        char *p = (char *) malloc(bytes); t0 = now();
        { int nt = 16; std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back([=] { size_t lo = bytes / nt * t, hi = t == nt - 1 ? bytes : bytes / nt * (t + 1);
              for (size_t o = lo; o < hi; o += (64u << 20)) { size_t n = hi - o < (64u << 20) ? hi - o : (64u << 20); memcpy(p + o, h + (size_t) (t % 15) * (64u << 20), n); } }); for (auto &t : th) t.join(); }
        t1 = now(); printf("5b memcpy pinned -> fresh pageable, 16 threads, %.1f GB: %.2f s\n", bytes / 1e9, t1 - t0);
        t0 = now();
        { int nt = 16; std::vector<std::thread> th; for (int t = 0; t < nt; t++) th.emplace_back([=] { size_t lo = bytes / nt * t, hi = t == nt - 1 ? bytes : bytes / nt * (t + 1);
              for (size_t o = lo; o < hi; o += (64u << 20)) { size_t n = hi - o < (64u << 20) ? hi - o : (64u << 20); memcpy(p + o, h, n); } }); for (auto &t : th) t.join(); }
        t1 = now(); printf("5c same copy again (pages resident): %.2f s\n", t1 - t0);
        free(p); cudaFreeHost(h);
    }
    return 0;
}
