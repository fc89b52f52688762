// How long page-locking host memory takes on this box, by method: cudaMallocHost, and cudaHostRegister of memory backed by
// transparent huge pages (madvise) or by ordinary pages; plus the H2D rate out of each.  A diagnosis for the read-block ring
// of csrc/files.cpp (PinnedPool), not a bench value.   nvcc -O2 -o tools/pin_timing tools/pin_timing.cu
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <sys/mman.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static double h2d_gbps(const void *h, void *d, size_t n)
{
    cudaMemcpy(d, h, n, cudaMemcpyHostToDevice);
    const double t0 = now();
    for (int i = 0; i < 4; ++i)
        cudaMemcpy(d, h, n, cudaMemcpyHostToDevice);
    return 4.0 * n / (now() - t0) / 1e9;
}

int main()
{
    if (FILE *f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r"))
    {
        char line[128] = {0};
        if (fgets(line, sizeof line, f))
            printf("{\"thp_enabled\": \"%.*s\"}\n", (int)strcspn(line, "\n"), line);
        fclose(f);
    }
    cudaFree(0);
    void *d = nullptr;
    cudaMalloc(&d, 256u << 20);
    for (size_t mb : {65, 256})
    {
        const size_t n = mb << 20;
        for (int rep = 0; rep < 2; ++rep)
        {
            double t0 = now();
            void  *p  = nullptr;
            cudaMallocHost(&p, n);
            const double t_malloc = now() - t0;
            const double bw0      = h2d_gbps(p, d, n);
            t0                    = now();
            cudaFreeHost(p);
            const double t_free = now() - t0;
            printf("{\"method\": \"cudaMallocHost\", \"MiB\": %zu, \"s\": %.4f, \"free_s\": %.4f, \"h2d_GBps\": %.1f}\n", mb, t_malloc, t_free, bw0);
            for (int huge = 1; huge >= 0; --huge)
            {
                t0      = now();
                void *q = mmap(nullptr, n + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
                char *a = (char *)(((uintptr_t)q + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1));
                if (huge)
                    madvise(a, n, MADV_HUGEPAGE);
                for (size_t o = 0; o < n; o += 4096)
                    a[o] = 1;
                const double t_touch = now() - t0;
                t0                   = now();
                const cudaError_t e  = cudaHostRegister(a, n, cudaHostRegisterDefault);
                const double t_reg   = now() - t0;
                const double bw      = e == cudaSuccess ? h2d_gbps(a, d, n) : 0;
                t0                   = now();
                if (e == cudaSuccess)
                    cudaHostUnregister(a);
                const double t_unreg = now() - t0;
                munmap(q, n + (2u << 20));
                printf("{\"method\": \"%s + cudaHostRegister\", \"MiB\": %zu, \"touch_s\": %.4f, \"register_s\": %.4f, \"unregister_s\": %.4f, \"h2d_GBps\": %.1f, \"err\": %d}\n",
                       huge ? "mmap + MADV_HUGEPAGE" : "mmap (4 KiB pages)", mb, t_touch, t_reg, t_unreg, bw, (int)e);
            }
        }
    }
    return 0;
}
