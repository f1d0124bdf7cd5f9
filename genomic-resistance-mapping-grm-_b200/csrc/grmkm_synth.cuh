// grmkm_synth.cuh -- synthetic bacterial-scale FASTA generator (bench/test utility).
// Spec: SURVEY.md section 8d / BASELINE.md section 4; the numpy mirror in synth.py produces identical bytes.
// Layout table (u64 words, host-built by synth.py):
//   [0] magic 'GRMSYN01'  [1] seed  [2] n_genomes  [3] core_len  [4] island_len  [5] n_contigs (C)
//   [6] line_width  [7] total_bytes  [8] n_islands (NI)  [9] genome block stride (words)  [10..15] reserved
//   then per genome a block:
//     [0] genome id  [1] byte offset in dst (16-aligned)  [2] fasta bytes  [3] n_present islands  [4] genome length
//     [5 .. 5+C]            contig bounds in genome coordinates (C+1)
//     [.. +C+1]             contig byte offsets relative to the genome's FASTA start (C+1)
//     [.. +C]               header length in bytes
//     [.. +4C]              header text, 32 bytes per contig
//     [.. +NI]              indices of present islands (first n_present valid)
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

namespace grmkm {

constexpr uint64_t kSynthMagic = 0x31304E59534D5247ULL;  // "GRMSYN01" little-endian
constexpr int kSynthHeaderWords = 16;

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t synth_h(uint64_t seed, uint64_t a, uint64_t b) {
    return splitmix64(seed ^ (a * 0x9E3779B97F4A7C15ULL) ^ (b * 0xC2B2AE3D27D4EB4FULL));
}

// base (0..3 = A C G T) of genome g at core position p
__device__ __forceinline__ uint32_t synth_core_base(uint64_t seed, uint64_t g, uint64_t p) {
    uint32_t v = (uint32_t)(synth_h(seed, 0, p) & 3);
    if (synth_h(seed, 1, p) < 184467440737095516ULL) {            // shared variant site (1 %)
        const uint32_t e = 1 + (uint32_t)(synth_h(seed, 3, p) % 7);
        if ((synth_h(seed, 4, (g << 32) + p) >> (64 - e)) == 0)   // allele frequency 2^-e
            v = (v + 1 + (uint32_t)(synth_h(seed, 2, p) % 3)) & 3;
    }
    if (synth_h(seed, 5, (g << 32) + p) < 1844674407370955ULL)    // private SNP (1e-4)
        v = (v + 1 + (uint32_t)(synth_h(seed, 8, (g << 32) + p) % 3)) & 3;
    return v;
}

__global__ void k_synth_fasta(const uint64_t* __restrict__ lay, uint8_t* __restrict__ dst, uint64_t dst_bytes) {
    const uint64_t seed = lay[1], G = lay[2], Lc = lay[3], LI = lay[4], C = lay[5], LW = lay[6], total = lay[7];
    const uint64_t NI = lay[8], stride = lay[9];
    (void)NI;
    const uint64_t n = total < dst_bytes ? total : dst_bytes;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        // genome by byte offset
        uint64_t lo = 0, hi = G - 1;
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) >> 1;
            if (lay[kSynthHeaderWords + mid * stride + 1] <= i) lo = mid; else hi = mid - 1;
        }
        const uint64_t* gb = lay + kSynthHeaderWords + lo * stride;
        const uint64_t gid = gb[0], rel = i - gb[1], flen = gb[2];
        if (rel >= flen) { dst[i] = '\n'; continue; }      // padding between genomes
        const uint64_t* bounds = gb + 5;
        const uint64_t* coff = bounds + (C + 1);
        const uint64_t* hlen = coff + (C + 1);
        const uint64_t* htxt = hlen + C;
        const uint64_t* isl = htxt + 4 * C;
        uint64_t cl = 0, ch = C - 1;
        while (cl < ch) {
            const uint64_t mid = (cl + ch + 1) >> 1;
            if (coff[mid] <= rel) cl = mid; else ch = mid - 1;
        }
        const uint64_t j = cl, q0 = rel - coff[j];
        if (q0 < hlen[j]) { dst[i] = (uint8_t)(htxt[4 * j + (q0 >> 3)] >> (8 * (q0 & 7))); continue; }
        const uint64_t q = q0 - hlen[j], L = bounds[j + 1] - bounds[j];
        const uint64_t line = q / (LW + 1), col = q % (LW + 1), idx = line * LW + col;
        if (col == LW || idx >= L) { dst[i] = '\n'; continue; }
        const bool rc = (j & 1) != 0;
        const uint64_t x = rc ? bounds[j + 1] - 1 - idx : bounds[j] + idx;
        uint32_t v;
        if (x < Lc) v = synth_core_base(seed, gid, x);
        else {
            const uint64_t t = (x - Lc) / LI, qq = (x - Lc) % LI;
            v = (uint32_t)(synth_h(seed, 6, (isl[t] << 32) + qq) & 3);
        }
        if (rc) v = 3 - v;
        dst[i] = "ACGT"[v];
    }
}

// ---- read sets (BASELINE.json configs[3]): FASTQ records of constant width per genome
// Layout: header word [10] = 1, [11] = read length, [12] = substitution threshold (error rate x 2^64), [13] = reads per
// genome; per genome block (stride [9]):
//   [0] genome id  [1] byte offset in dst  [2] FASTQ bytes  [3] n_present islands  [4] genome length  [5] record width
//   [6] header prefix length  [7] digits of the read number  [8..9] header prefix "@g<id>_r" (16 bytes)  [10 ..] islands
// Record j = prefix, j as [7] decimal digits, LF, read, LF, '+', LF, read-length 'I', LF.  Read j of genome g starts at
// h(10, g<<32 | j) mod (L - len + 1), strand h(11, .) & 1, base t substituted when h(12, (g<<32 | j) * 1024 + t) is below
// the threshold -- the same function as synth.genome_reads_fastq_fixed.
__global__ void k_synth_fastq(const uint64_t* __restrict__ lay, uint8_t* __restrict__ dst, uint64_t dst_bytes) {
    const uint64_t seed = lay[1], G = lay[2], Lc = lay[3], LI = lay[4], total = lay[7], stride = lay[9];
    const uint64_t RL = lay[11], thr = lay[12];
    const uint64_t n = total < dst_bytes ? total : dst_bytes;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = G - 1;
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) >> 1;
            if (lay[kSynthHeaderWords + mid * stride + 1] <= i) lo = mid; else hi = mid - 1;
        }
        const uint64_t* gb = lay + kSynthHeaderWords + lo * stride;
        const uint64_t gid = gb[0], rel = i - gb[1], flen = gb[2], L = gb[4], recw = gb[5], pl = gb[6], D = gb[7];
        if (rel >= flen) { dst[i] = '\n'; continue; }      // padding between genomes
        const uint64_t j = rel / recw, q = rel % recw;
        uint8_t out;
        if (q < pl) out = (uint8_t)(gb[8 + (q >> 3)] >> (8 * (q & 7)));
        else if (q < pl + D) {
            uint64_t v = j;
            for (uint64_t d = pl + D - 1; d > q; --d) v /= 10;
            out = (uint8_t)('0' + v % 10);
        } else if (q == pl + D || q == pl + D + 1 + RL || q == pl + D + 3 + RL || q == recw - 1) out = '\n';
        else if (q == pl + D + 2 + RL) out = '+';
        else if (q > pl + D + 3 + RL) out = 'I';
        else {
            const uint64_t t = q - (pl + D + 1), gj = (gid << 32) + j;
            const uint64_t start = synth_h(seed, 10, gj) % (L - RL + 1);
            const bool rc = (synth_h(seed, 11, gj) & 1) != 0;
            const uint64_t x = rc ? start + RL - 1 - t : start + t;
            uint32_t v;
            if (x < Lc) v = synth_core_base(seed, gid, x);
            else {
                const uint64_t it = (x - Lc) / LI, qq = (x - Lc) % LI;
                v = (uint32_t)(synth_h(seed, 6, (gb[10 + it] << 32) + qq) & 3);
            }
            if (rc) v = 3 - v;
            const uint64_t cell = gj * 1024 + t;
            if (synth_h(seed, 12, cell) < thr) v = (v + 1 + (uint32_t)(synth_h(seed, 13, cell) % 3)) & 3;
            out = "ACGT"[v];
        }
        dst[i] = out;
    }
}

inline bool synth_launch(const uint8_t* host_layout, uint64_t layout_bytes, const uint8_t* dev_layout, uint8_t* dev_dst,
                         uint64_t dst_bytes, cudaStream_t st, std::string& msg) {
    if (layout_bytes < kSynthHeaderWords * 8) { msg = "synth layout too small"; return false; }
    const uint64_t* h = reinterpret_cast<const uint64_t*>(host_layout);
    if (h[0] != kSynthMagic) { msg = "synth layout: bad magic"; return false; }
    if (h[2] == 0 || (kSynthHeaderWords + h[2] * h[9]) * 8 > layout_bytes) { msg = "synth layout: truncated"; return false; }
    if (h[7] > dst_bytes) { msg = "synth destination too small"; return false; }
    if (h[10] == 1) {
        if (h[11] == 0 || h[11] > 1024) { msg = "synth layout: read length must be 1..1024"; return false; }
        k_synth_fastq<<<148 * 16, 256, 0, st>>>(reinterpret_cast<const uint64_t*>(dev_layout), dev_dst, dst_bytes);
    } else {
        k_synth_fasta<<<148 * 16, 256, 0, st>>>(reinterpret_cast<const uint64_t*>(dev_layout), dev_dst, dst_bytes);
    }
    return true;
}

}  // namespace grmkm
