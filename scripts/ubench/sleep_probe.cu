// Developer micro-probe: how long do nanosleep / mbarrier.try_wait(+hint) really suspend a warp on this GPU,
// alone and while other warps of the CTA keep arriving on other mbarriers?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t gtime() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void probe(int mode, uint32_t arg, int noisy, uint64_t* out) {
    __shared__ uint64_t bar[4];
    const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&bar[0]);
    const uint32_t b1 = (uint32_t)__cvta_generic_to_shared(&bar[1]);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b0), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b1), "r"(1));
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        uint64_t t0 = gtime();
        const int iters = 2000;
        for (int i = 0; i < iters; i++) {
            if (mode == 0) {
                asm volatile("nanosleep.u32 %0;" ::"r"(arg));
            } else if (mode == 1) {
                uint32_t ok;
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(b0), "r"(0u), "r"(arg) : "memory");
            } else {
                uint32_t ok;
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(b0), "r"(0u) : "memory");
            }
        }
        uint64_t t1 = gtime();
        if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
    } else if (noisy && warp == 1) {
        // keep arriving on ANOTHER barrier
        for (int i = 0; i < 200000; i++) {
            if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b1) : "memory");
            __syncwarp();
        }
    }
}
int main() {
    uint64_t* d; cudaMalloc(&d, 8);
    for (int noisy = 0; noisy < 2; noisy++)
        for (int mode = 0; mode < 3; mode++)
            for (uint32_t arg : {100u, 1000u, 10000u, 100000u}) {
                probe<<<1, 64>>>(mode, arg, noisy, d);
                uint64_t h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("noisy=%d mode=%s arg=%u: %llu ns per call\n", noisy, mode == 0 ? "nanosleep" : mode == 1 ? "try_wait+hint" : "try_wait", arg, (unsigned long long)h);
                if (mode == 2) break;
            }
    return 0;
}
