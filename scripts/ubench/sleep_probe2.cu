// Developer micro-probe 2: does nanosleep keep its duration (a) next to busy warps, (b) when OTHER mbarriers of the CTA
// complete phases, (c) inside the try_wait + nanosleep back-off loop the kernel uses?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t gtime() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__global__ void probe(int mode, uint32_t arg, uint64_t* out) {
    __shared__ uint64_t bar[4];
    __shared__ uint32_t buf[1024];
    const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&bar[0]);
    const uint32_t b1 = (uint32_t)__cvta_generic_to_shared(&bar[1]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b1), "r"(1));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    volatile __shared__ int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (warp == 0) {
        uint64_t t0 = gtime();
        const int iters = 500;
        uint32_t polls = 0;
        for (int i = 0; i < iters; i++) {
            if (mode <= 1) asm volatile("nanosleep.u32 %0;" ::"r"(arg));
            else {
                // the kernel's loop: wait on bar0 (never completes) for at most ~arg*8 ns of back-off
                uint32_t ns = 32;
                for (int k = 0; k < 8; k++) { if (try_wait(b0, 0)) break; __nanosleep(ns); ns = min(ns * 2, arg); polls++; }
            }
        }
        uint64_t t1 = gtime();
        if (lane == 0) { out[0] = (t1 - t0) / iters; out[1] = polls / iters; }
        __syncwarp();
        if (lane == 0) stop = 1;
    } else if (mode >= 1) {
        // busy warps: ALU + LDS, and one of them completes phases of ANOTHER barrier all the time
        uint32_t x = threadIdx.x;
        while (!stop) {
            for (int k = 0; k < 64; k++) { x = x * 1664525u + buf[(x >> 8) & 1023]; }
            if (warp == 1 && lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
        }
        if (x == 12345) buf[0] = x;
    }
}
int main() {
    uint64_t* d; cudaMalloc(&d, 16);
    for (int mode = 0; mode < 3; mode++)
        for (uint32_t arg : {64u, 512u, 4096u}) {
            probe<<<1, 1024>>>(mode, arg, d);
            uint64_t h[2] = {0, 0}; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("mode=%s arg=%u: %llu ns per iteration (polls %llu)\n", mode == 0 ? "nanosleep alone" : mode == 1 ? "nanosleep + 31 busy warps + phase flips" : "try_wait/back-off loop + busy", arg,
                   (unsigned long long)h[0], (unsigned long long)h[1]);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
