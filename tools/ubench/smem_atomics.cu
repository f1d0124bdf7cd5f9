// Micro-benchmarks that decide the scatter/aggregate design: throughput of shared-memory atomics
// (returning / non-returning, add / or / cas) and random LDS/STS on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/smem_atomics smem_atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(uint32_t* out, int iters, uint32_t nbins) {
    extern __shared__ uint32_t s[];
    for (uint32_t i = threadIdx.x; i < nbins * 2; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t x = mix(threadIdx.x * 2654435761u + blockIdx.x);
    for (int it = 0; it < iters; ++it) {
        uint32_t idx[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { x = x * 1664525u + 1013904223u; idx[e] = (x >> 8) % nbins; }
        if (MODE == 0) {          // returning add
#pragma unroll
            for (int e = 0; e < 16; ++e) acc += atomicAdd(&s[idx[e]], 1u);
        } else if (MODE == 1) {   // non-returning add
#pragma unroll
            for (int e = 0; e < 16; ++e) atomicAdd(&s[idx[e]], 1u);
        } else if (MODE == 2) {   // non-returning or
#pragma unroll
            for (int e = 0; e < 16; ++e) atomicOr(&s[idx[e]], 1u << (x & 31));
        } else if (MODE == 3) {   // random LDS.32
#pragma unroll
            for (int e = 0; e < 16; ++e) acc += s[idx[e]];
        } else if (MODE == 4) {   // random STS.32
#pragma unroll
            for (int e = 0; e < 16; ++e) s[idx[e]] = x;
        } else if (MODE == 5) {   // random LDS.64
#pragma unroll
            for (int e = 0; e < 16; ++e) acc += (uint32_t)((unsigned long long*)s)[idx[e]];
        } else if (MODE == 6) {   // 64-bit CAS (returning)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc += (uint32_t)atomicCAS(&((unsigned long long*)s)[idx[e]], 0ULL, (unsigned long long)x);
        } else if (MODE == 7) {   // no memory op: ALU only baseline
#pragma unroll
            for (int e = 0; e < 16; ++e) acc += idx[e];
        } else if (MODE == 8) {   // random STS.64
#pragma unroll
            for (int e = 0; e < 16; ++e) ((unsigned long long*)s)[idx[e]] = x;
        } else if (MODE == 9) {   // match_any on a 12-bit value + popc rank
#pragma unroll
            for (int e = 0; e < 16; ++e) { uint32_t m = __match_any_sync(0xffffffffu, idx[e]); acc += __popc(m); }
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE>
void run(const char* name, uint32_t nbins, int threads) {
    uint32_t* d; cudaMalloc(&d, 4);
    const int iters = 200;
    size_t smem = (size_t)nbins * 8;
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    bench<MODE><<<148, threads, smem>>>(d, 10, nbins);
    cudaEventRecord(a);
    bench<MODE><<<148, threads, smem>>>(d, iters, nbins);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = 148.0 * threads * iters * 16;
    double clk = 1.965e9;
    printf("%-28s bins=%6u thr=%4d  %8.3f ms  %7.2f Gop/s  %6.3f cyc/op/SM  err=%s\n", name, nbins, threads, ms,
           ops / ms / 1e6, ms * 1e-3 * clk / (ops / 148.0), cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    for (int thr : {1024, 512}) {
        for (uint32_t nb : {4096u, 16384u}) {
            run<7>("alu only", nb, thr);
            run<0>("atoms add returning", nb, thr);
            run<1>("atoms add no-return", nb, thr);
            run<2>("atoms or no-return", nb, thr);
            run<3>("lds.32 random", nb, thr);
            run<4>("sts.32 random", nb, thr);
            run<5>("lds.64 random", nb, thr);
            run<8>("sts.64 random", nb, thr);
            run<6>("atoms cas.64 returning", nb, thr);
            run<9>("match_any", nb, thr);
        }
    }
    return 0;
}
