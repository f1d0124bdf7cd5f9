// Micro-benchmark: cost of 8-byte global stores as a function of the run length (consecutive records that
// go to the same bucket region from one CTA tile).  R = 1: every lane hits its own 32-byte sector (direct
// scatter from registers); R = 4: 32-byte runs (4096 buckets, 16 K tile); R = 16: 128-byte runs (1024 buckets).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// tile = 16384 records per CTA iteration; B buckets; each (tile, CTA, bucket) owns R = 16384 / B consecutive slots
template <int R, bool ALIGNED>
__global__ void __launch_bounds__(512, 1) bench(unsigned long long* out, uint64_t cap, int tiles) {
    constexpr uint32_t B = 16384 / R;
    uint32_t x = mix(threadIdx.x * 2654435761u + blockIdx.x * 97u);
    for (int t = 0; t < tiles; ++t) {
        uint64_t tile_slot = ((uint64_t)t * gridDim.x + blockIdx.x) * R;
#pragma unroll 8
        for (int e = 0; e < 32; ++e) {
            x = x * 1664525u + 1013904223u;
            const uint32_t i = e * 512 + threadIdx.x;
            uint32_t b = (i / R) % B, r = i % R;
            b = (b * 2654435761u) >> (32 - __builtin_ctz(B));          // scramble the bucket order
            const uint64_t mis = ALIGNED ? 0 : (b & 3);              // unaligned: runs straddle sector boundaries
            out[(uint64_t)b * cap + tile_slot + r + mis] = ((unsigned long long)x << 32) | b;
        }
    }
}

template <int R, bool ALIGNED>
void run(unsigned long long* d, int tiles) {
    const uint64_t cap = ((uint64_t)tiles * 148 * R + 64);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a); bench<R, ALIGNED><<<148, 512>>>(d, cap, tiles); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
    }
    printf("run length %3d records (%4d B) %s: %7.3f ms  %7.1f GB/s  (%s)\n", R, R * 8, ALIGNED ? "aligned  " : "unaligned", best,
           148.0 * tiles * 16384 * 8 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int tiles = 206;                       // per CTA, as in C2 (500 M records / 148 / 16384)
    unsigned long long* d;
    const size_t bytes = (size_t)16384 * ((uint64_t)tiles * 148 + 80) * 8;
    if (cudaMalloc(&d, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    printf("buffer %.2f GB, %.0f M records\n", bytes / 1e9, 148.0 * tiles * 16384 / 1e6);
    run<1, true>(d, tiles);
    run<2, true>(d, tiles);
    run<4, true>(d, tiles);
    run<4, false>(d, tiles);
    run<8, true>(d, tiles);
    run<8, false>(d, tiles);
    run<16, true>(d, tiles);
    run<16, false>(d, tiles);
    run<32, true>(d, tiles);
    run<32, false>(d, tiles);
    run<64, true>(d, tiles);
    run<128, true>(d, tiles);
    return 0;
}
