// grmkm_result.cuh -- kernels over the finished matrix (kmers[U], matrix[W][U] row-major with a row pitch >= U, genome
// g at bit 63 - (g & 63) of word row g >> 6, kover/utils.py:144-154) and over Ray Surveyor TSV text.
//
//   k_checksum     order-independent digest of the columns (parity checks across GPU counts, bench.py parity_check)
//   k_sum_rows     KmerRuleClassifications.sum_rows (bin/kover/core/kover/learning/common/rules.py:201-267 with
//                  popcount.pyx:76-95): per column, popcount of the words under a row mask
//   k_tsv_pack     from_tsv's bit packer (dataset/create.py:241-271 + utils.py:133-156): fixed-width TSV rows ->
//                  matrix words
//   k_bit_rows / k_gram   Ray Surveyor's similarity (Gram) matrix: shared k-mers per genome pair
#pragma once
#include "grmkm_kernels.cuh"

namespace grmkm {

// digest of one column; the checksum is the wrapping sum of (d, fmix64(d ^ kChkB)) over the columns
constexpr unsigned long long kChkA = 0x9E3779B97F4A7C15ULL, kChkW = 0xC2B2AE3D27D4EB4FULL, kChkB = 0xA5A5A5A5A5A5A5A5ULL;

__global__ void __launch_bounds__(256)
k_checksum(const unsigned long long* __restrict__ kmers, const unsigned long long* __restrict__ matrix, unsigned long long U,
           unsigned long long pitch, uint32_t W, unsigned long long* __restrict__ out /* [2], zeroed */) {
    unsigned long long s0 = 0, s1 = 0;
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < U;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long d = fmix64(kmers[j] + kChkA);
        for (uint32_t w = 0; w < W; ++w) d = fmix64(d ^ (matrix[(unsigned long long)w * pitch + j] + (w + 1) * kChkW));
        s0 += d;
        s1 += fmix64(d ^ kChkB);
    }
    const unsigned long long t0 = block_sum_u64(s0);
    __syncthreads();
    const unsigned long long t1 = block_sum_u64(s1);
    if (threadIdx.x == 0) { atomicAdd(out, t0); atomicAdd(out + 1, t1); }
}

// out[j] = sum_w popcount(matrix[w][j] & mask[w])  -- the learner's hot loop on the resident matrix
__global__ void __launch_bounds__(256)
k_sum_rows(const unsigned long long* __restrict__ matrix, unsigned long long U, unsigned long long pitch, uint32_t W,
           const unsigned long long* __restrict__ mask, uint32_t* __restrict__ out) {
    __shared__ unsigned long long s_mask[512];
    for (uint32_t w = threadIdx.x; w < W; w += blockDim.x) s_mask[w] = mask[w];
    __syncthreads();
    for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < U;
         j += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t s = 0;
        for (uint32_t w = 0; w < W; ++w) {
            const unsigned long long m = s_mask[w];
            if (m) s += (uint32_t)__popcll(__ldcs(matrix + (unsigned long long)w * pitch + j) & m);
        }
        out[j] = s;
    }
}

// One warp per TSV row: row j = "<k-mer>\t<c_0>\t<c_1>...\n" (row_width bytes); matrix row g takes TSV column sel[g].
// Lanes read 32 cells at a time, a ballot turns them into half a word.
__global__ void __launch_bounds__(256)
k_tsv_pack(const uint8_t* __restrict__ body, unsigned long long n_rows, uint32_t row_width, uint32_t k, uint32_t G,
           const uint32_t* __restrict__ sel, unsigned long long j0, unsigned long long U,
           unsigned long long* __restrict__ matrix, unsigned int* __restrict__ bad) {
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    const uint32_t W = (G + 63) / 64;
    for (unsigned long long j = warp; j < n_rows; j += n_warps) {
        const uint8_t* row = body + j * row_width;
        for (uint32_t w = 0; w < W; ++w) {
            uint32_t half[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t g = w * 64 + h * 32 + lane;
                uint32_t c = '0';
                if (g < G) c = row[k + 1 + 2 * sel[g]];
                if (c != '0' && c != '1') *bad = 1u;                    // create.py:121-137: the matrix must be binary
                half[h] = __brev(__ballot_sync(0xffffffffu, c == '1'));   // genome 32h + l at bit 31 - l
            }
            if (lane == 0) matrix[(unsigned long long)w * U + j0 + j] = ((unsigned long long)half[0] << 32) | half[1];
        }
    }
}

// ---- Gram matrix (Ray Surveyor's similarity matrix): gram[a][b] = number of columns present in genomes a and b.
// The column-major words are first turned into per-genome bit rows (64 x 64 bit transposes), then every pair of
// genomes is a popcount(AND) stream over U / 64 words.
__global__ void __launch_bounds__(256)
k_bit_rows(const unsigned long long* __restrict__ matrix, unsigned long long U, unsigned long long pitch, uint32_t W, unsigned long long UW /* ceil(U / 64) */,
           unsigned long long* __restrict__ rows /* [W * 64][UW] */) {
    // one warp per (word row w, block of 64 columns): lane l holds columns 2l and 2l + 1
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long t = warp; t < (unsigned long long)W * UW; t += n_warps) {
        const uint32_t w = (uint32_t)(t / UW);
        const unsigned long long cb = t % UW;
        const unsigned long long j = cb * 64 + 2 * lane;
        const unsigned long long c0 = j < U ? matrix[(unsigned long long)w * pitch + j] : 0ULL;
        const unsigned long long c1 = j + 1 < U ? matrix[(unsigned long long)w * pitch + j + 1] : 0ULL;
        // genome r of this word row (bit 63 - r of a column word) -> one 64-bit row word: column cb * 64 + q at bit q
        for (uint32_t r = 0; r < 64; ++r) {
            const uint32_t b0 = (uint32_t)(c0 >> (63 - r)) & 1u, b1 = (uint32_t)(c1 >> (63 - r)) & 1u;
            const uint32_t lo = __ballot_sync(0xffffffffu, b0), hi = __ballot_sync(0xffffffffu, b1);
            if (lane == 0) {
                // interleave: column 2l -> bit 2l, column 2l + 1 -> bit 2l + 1
                unsigned long long x = lo, y = hi;
                x = (x | (x << 16)) & 0x0000FFFF0000FFFFULL; x = (x | (x << 8)) & 0x00FF00FF00FF00FFULL;
                x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0FULL; x = (x | (x << 2)) & 0x3333333333333333ULL;
                x = (x | (x << 1)) & 0x5555555555555555ULL;
                y = (y | (y << 16)) & 0x0000FFFF0000FFFFULL; y = (y | (y << 8)) & 0x00FF00FF00FF00FFULL;
                y = (y | (y << 4)) & 0x0F0F0F0F0F0F0F0FULL; y = (y | (y << 2)) & 0x3333333333333333ULL;
                y = (y | (y << 1)) & 0x5555555555555555ULL;
                rows[((unsigned long long)w * 64 + r) * UW + cb] = x | (y << 1);
            }
        }
    }
}

// one CTA per genome pair (a <= b): popcount(row_a & row_b) over UW words
__global__ void __launch_bounds__(256)
k_gram(const unsigned long long* __restrict__ rows, unsigned long long UW, uint32_t G, unsigned long long* __restrict__ gram) {
    const unsigned long long n_pairs = (unsigned long long)G * (G + 1) / 2;
    for (unsigned long long pidx = blockIdx.x; pidx < n_pairs; pidx += gridDim.x) {
        // pair index -> (a, b), a <= b, row-major over the upper triangle
        uint32_t a = 0;
        unsigned long long rem = pidx;
        while (rem >= (unsigned long long)(G - a)) { rem -= G - a; ++a; }
        const uint32_t b = a + (uint32_t)rem;
        const unsigned long long* ra = rows + (unsigned long long)a * UW;
        const unsigned long long* rb = rows + (unsigned long long)b * UW;
        unsigned long long s = 0;
        for (unsigned long long i = threadIdx.x; i < UW; i += blockDim.x) s += (unsigned long long)__popcll(ra[i] & rb[i]);
        __syncthreads();
        const unsigned long long t = block_sum_u64(s);
        if (threadIdx.x == 0) { gram[(unsigned long long)a * G + b] = t; gram[(unsigned long long)b * G + a] = t; }
    }
}

}  // namespace grmkm
