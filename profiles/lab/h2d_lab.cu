// h2d_lab.cu - where does the host->device ceiling of N concurrent ranks come from?
//
// One process per GPU copies a 781 MB batch (64 frames of 2048 x 1489 float32, what lfd_submit moves per step) from
// host staging to its GPU with ONE cudaMemcpyAsync per iteration, for a fixed wall-clock window that starts at a
// common epoch, and prints its own GB/s.  Variants of how the staging is allocated / placed:
//   a  cudaMallocHost                                 (what lfd_create does)
//   b  mmap(MAP_HUGETLB) (fallback: THP madvise) + cudaHostRegister
//   c  cudaHostAlloc(cudaHostAllocWriteCombined)
//   d  cudaMallocHost, allocated and first-touched while pinned to the LOW half of the CPUs
//   e  cudaMallocHost, allocated and first-touched while pinned to the HIGH half of the CPUs
//   f  like a, but TWO staging buffers / streams in flight per process (does one copy engine stream saturate the link?)
// usage: h2d_lab <device> <variant> <start_epoch_s> <seconds>      (profiles/lab/h2d_lab.sh drives it)
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a profiles/lab/h2d_lab.cu -o profiles/lab/h2d_lab
#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/time.h>
#include <unistd.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static double now() { struct timeval tv; gettimeofday(&tv, 0); return tv.tv_sec + 1e-6 * tv.tv_usec; }

static void pin_half(int high)
{
    const int n = (int)sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set; CPU_ZERO(&set);
    for (int c = 0; c < n; c++) if ((c >= n / 2) == (high != 0)) CPU_SET(c, &set);
    sched_setaffinity(0, sizeof set, &set);
}

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: h2d_lab <device> <variant a-f> <start_epoch_s> <seconds>\n"); return 1; }
    const int dev = atoi(argv[1]);
    const char var = argv[2][0];
    const double start = atof(argv[3]), secs = atof(argv[4]);
    const size_t bytes = (size_t)64 * 2048 * 1489 * 4;
    CK(cudaSetDevice(dev));
    const int nbuf = var == 'f' ? 2 : 1;
    void* host[2] = {nullptr, nullptr};
    void* devp[2] = {nullptr, nullptr};
    cudaStream_t st[2];
    const char* how = "";
    if (var == 'd') pin_half(0);
    if (var == 'e') pin_half(1);
    for (int b = 0; b < nbuf; b++) {
        switch (var) {
        case 'b': {
            const size_t huge = (bytes + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1);
            void* p = mmap(nullptr, huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
            how = "MAP_HUGETLB";
            if (p == MAP_FAILED) {
                p = mmap(nullptr, huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                if (p == MAP_FAILED) { perror("mmap"); return 2; }
                madvise(p, huge, MADV_HUGEPAGE);
                how = "THP(madvise)";
            }
            memset(p, 1, huge);
            CK(cudaHostRegister(p, huge, cudaHostRegisterDefault));
            host[b] = p;
            break;
        }
        case 'c': CK(cudaHostAlloc(&host[b], bytes, cudaHostAllocWriteCombined)); memset(host[b], 1, bytes); how = "write-combined"; break;
        default: CK(cudaMallocHost(&host[b], bytes)); memset(host[b], 1, bytes); how = var == 'd' ? "low-half first touch" : var == 'e' ? "high-half first touch" : var == 'f' ? "two buffers in flight" : "cudaMallocHost"; break;
        }
        CK(cudaMalloc(&devp[b], bytes));
        CK(cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking));
    }
    for (int b = 0; b < nbuf; b++) CK(cudaMemcpyAsync(devp[b], host[b], bytes, cudaMemcpyHostToDevice, st[b]));     // warm-up
    CK(cudaDeviceSynchronize());
    while (now() < start) usleep(200);
    const double t0 = now();
    long copies = 0;
    while (now() - t0 < secs) {
        for (int b = 0; b < nbuf; b++) CK(cudaMemcpyAsync(devp[b], host[b], bytes, cudaMemcpyHostToDevice, st[b]));
        for (int b = 0; b < nbuf; b++) CK(cudaStreamSynchronize(st[b]));
        copies += nbuf;
    }
    const double dt = now() - t0;
    printf("{\"device\": %d, \"variant\": \"%c\", \"how\": \"%s\", \"gbs\": %.2f, \"copies\": %ld, \"seconds\": %.2f}\n", dev, var, how,
           copies * (double)bytes / dt / 1e9, copies, dt);
    return 0;
}
