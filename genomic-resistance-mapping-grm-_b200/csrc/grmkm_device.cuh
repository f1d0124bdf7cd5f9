// grmkm_device.cuh -- device-side building blocks shared by the kernels.
// Semantics implemented here are SURVEY.md Appendix E (E2-E8); the reference
// call sites they stand in for are cited in include/grmkm.h.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace grmkm {

constexpr int kParseThreads = 256;
constexpr int kChunksPerThread = 4;    // 16-byte chunks per thread: 64 contiguous bytes
constexpr int kTileBytes = kParseThreads * kChunksPerThread * 16;   // 16384 input bytes per parse tile
constexpr int kExtractThreads = 256;
#ifndef GRMKM_AGG_THREADS
#define GRMKM_AGG_THREADS 1024
#endif
constexpr int kAggThreads = GRMKM_AGG_THREADS;      // 1024: one aggregate CTA per SM; 512: two, each with half the table
constexpr int kAggCtasPerSm = 1024 / kAggThreads;
constexpr int kAggMaxSlots = 16 * kAggThreads;
constexpr int kMaxProbe = 96;
constexpr uint64_t kEmptyKey = ~0ULL;
constexpr int kCursorStride = 1;       // spacing of the bucket cursors (32 = 256 B apart was measured: no gain)

// device scalars (u64 each)
enum Scalar : int {
    S_STREAM_LEN = 0,   // entries in the packed base stream
    S_N_RECORDS,        // FASTA/FASTQ records seen
    S_N_WINDOWS,        // k-mer windows partitioned (N)
    S_U_NEEDED,         // columns the aggregate kernel wanted to emit
    S_N_DISTINCT,       // distinct k-mers before the singleton filter
    S_N_SPLITS,         // sub-range splits
    S_N_SOLID,          // records surviving the abundance filter
    S_WORK,             // local-sort overflow flag
    S_OVERFLOW,         // a bucket region of the over-provisioned scatter was too small
    S_STREAM_TOTAL,     // packed-stream entries summed over the batches of a build
    S_N_UNITS,          // units (super-k-mers) scattered
    S_WU_NEEDED,        // distinct (unit, 64-genome block) entries the dedupe produced
    S_N_WIDE,           // wide records the expansion produced
    S_WIDE_OVERFLOW,    // a region of the over-provisioned expansion was too small
    S_COUNT
};

struct FileDesc {
    const uint8_t* ptr;   // device pointer to the file's bytes
    uint64_t len;
    uint64_t tile_begin;  // first tile index of this file
    uint32_t row;         // genome row
    uint32_t kind;        // GRMKM_FASTA / GRMKM_FASTQ
};

// ---- invertible 64-bit mixer (murmur3 fmix64 and its inverse) ---------------
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t h) {
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33; return h;
}
__host__ __device__ __forceinline__ uint64_t unfmix64(uint64_t h) {
    h ^= h >> 33; h *= 0x9cb4b2f8129337dbULL;
    h ^= h >> 33; h *= 0x4f74430c22a54005ULL;
    h ^= h >> 33; return h;
}

// ---- multiplicative hash: a bijection on u64 that costs one 64-bit multiply ----------------
// bucket = top bits of the product, table key = the rest; the k-mer comes back with the inverse.
constexpr uint64_t kHashMul = 0x9E3779B97F4A7C15ULL;
constexpr uint64_t inv_odd64(uint64_t a) {
    uint64_t x = a;                      // Newton: doubles the number of correct low bits per step
    for (int i = 0; i < 6; ++i) x *= 2 - a * x;
    return x;
}
constexpr uint64_t kHashInv = inv_odd64(kHashMul);
static_assert(kHashMul * kHashInv == 1ULL, "kHashInv is not the inverse of kHashMul");
__host__ __device__ __forceinline__ uint64_t khash(uint64_t x) { return x * kHashMul; }
__host__ __device__ __forceinline__ uint64_t kunhash(uint64_t h) { return h * kHashInv; }

// ---- parse transducer summaries ---------------------------------------------
// A tile (or 16-byte chunk) of text is summarised as a function of the parser
// state s in {0..3} at its first byte: end state e[s] and number of stream
// entries c[s] it emits.  FASTA uses states 0 = header line, 1 = sequence line;
// FASTQ uses the line number mod 4.  Composition is associative, so the state
// and stream position of every tile come out of one scan.
struct Sum {
    uint32_t e;           // 4 x 2 bits: e[s] at bits 2s
    uint32_t c0, c1, c2, c3;
};
__device__ __forceinline__ uint32_t sum_end(const Sum& a, uint32_t s) { return (a.e >> (2 * s)) & 3u; }
__device__ __forceinline__ uint32_t sum_cnt(const Sum& a, uint32_t s) {
    return s == 0 ? a.c0 : (s == 1 ? a.c1 : (s == 2 ? a.c2 : a.c3));
}
__device__ __forceinline__ Sum sum_identity() { Sum r; r.e = 0xE4u; r.c0 = r.c1 = r.c2 = r.c3 = 0; return r; }
// A happens first, then B
__device__ __forceinline__ Sum sum_combine(const Sum& A, const Sum& B) {
    uint32_t e0 = sum_end(A, 0), e1 = sum_end(A, 1), e2 = sum_end(A, 2), e3 = sum_end(A, 3);
    Sum r;
    r.e = sum_end(B, e0) | (sum_end(B, e1) << 2) | (sum_end(B, e2) << 4) | (sum_end(B, e3) << 6);
    r.c0 = A.c0 + sum_cnt(B, e0);
    r.c1 = A.c1 + sum_cnt(B, e1);
    r.c2 = A.c2 + sum_cnt(B, e2);
    r.c3 = A.c3 + sum_cnt(B, e3);
    return r;
}
// FASTQ: every summary is a ROTATION of the four states (the state is the line number mod 4 and a piece of text moves
// it on by its newlines), so the composition is "rotate B's counts by A's newlines and add" -- ~16 instructions
// instead of ~45 for the general map.  e of a rotation by n: states (n, n+1, n+2, n+3) mod 4 at bits 0-1, 2-3, 4-5, 6-7.
__device__ __forceinline__ Sum sum_combine_rot(const Sum& A, const Sum& B) {
    const uint32_t n = A.e & 3u;
    uint32_t b0 = B.c0, b1 = B.c1, b2 = B.c2, b3 = B.c3;             // c'[s] = B.c[(s + n) & 3]
    if (n & 1u) { const uint32_t t = b0; b0 = b1; b1 = b2; b2 = b3; b3 = t; }
    if (n & 2u) { uint32_t t = b0; b0 = b2; b2 = t; t = b1; b1 = b3; b3 = t; }
    Sum r;
    r.e = (0x934E39E4u >> (8u * ((n + B.e) & 3u))) & 0xFFu;
    r.c0 = A.c0 + b0; r.c1 = A.c1 + b1; r.c2 = A.c2 + b2; r.c3 = A.c3 + b3;
    return r;
}
template <bool ROT>
__device__ __forceinline__ Sum sum_comb(const Sum& A, const Sum& B) { return ROT ? sum_combine_rot(A, B) : sum_combine(A, B); }
__device__ __forceinline__ Sum sum_shfl_up(const Sum& a, int d) {
    Sum r;
    r.e = __shfl_up_sync(0xffffffffu, a.e, d);
    r.c0 = __shfl_up_sync(0xffffffffu, a.c0, d);
    r.c1 = __shfl_up_sync(0xffffffffu, a.c1, d);
    r.c2 = __shfl_up_sync(0xffffffffu, a.c2, d);
    r.c3 = __shfl_up_sync(0xffffffffu, a.c3, d);
    return r;
}

// Ordered block scan over kParseThreads threads.  excl = fold of all earlier threads, total = fold of the whole
// block.  smem must hold 2 x (blockDim / 32) + 1 Sums: the warp totals are scanned by the first warp (a serial fold in
// every thread was a quarter of the FASTQ parse's instructions).  ROT: every summary is a rotation (FASTQ).
template <bool ROT>
__device__ __forceinline__ void block_scan_sum(const Sum& mine, Sum& excl, Sum& total, Sum* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    Sum inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Sum o = sum_shfl_up(inc, d);
        if (lane >= d) inc = sum_comb<ROT>(o, inc);
    }
    Sum prev = sum_shfl_up(inc, 1);
    if (lane == 0) prev = sum_identity();
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        Sum wi = lane < nwarp ? smem[lane] : sum_identity();
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Sum o = sum_shfl_up(wi, d);
            if (lane >= d) wi = sum_comb<ROT>(o, wi);
        }
        Sum wp = sum_shfl_up(wi, 1);
        if (lane == 0) wp = sum_identity();
        if (lane < nwarp) smem[nwarp + lane] = wp;
        if (lane == nwarp - 1) smem[2 * nwarp] = wi;
    }
    __syncthreads();
    excl = sum_comb<ROT>(smem[nwarp + warp], prev);
    total = smem[2 * nwarp];
    __syncthreads();
}

// ---- byte classes -------------------------------------------------------------
__device__ __forceinline__ bool is_acgt(uint32_t c) {
    uint32_t u = c & 0xDFu;  // upper-case
    return u == 'A' || u == 'C' || u == 'G' || u == 'T';
}

struct Chunk16 {
    uint32_t w[4];
    __device__ __forceinline__ uint32_t byte(int i) const { return (w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }
};

// 16 bytes of a file at offset off (off is a multiple of 16 relative to a
// 16-byte aligned base); bytes beyond len read as 0.
__device__ __forceinline__ Chunk16 load_chunk(const uint8_t* base, uint64_t off, uint64_t len) {
    Chunk16 c;
    if (off + 16 <= len) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + off));
        c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
    } else {
        c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0;
        for (int i = 0; i < 16; ++i)
            if (off + i < len) c.w[i >> 2] |= (uint32_t)base[off + i] << (8 * (i & 3));
    }
    return c;
}

// FASTA states
constexpr uint32_t ST_HDR = 0, ST_SEQ = 1;

// summary of one 16-byte chunk (E2 / E3).  prev = byte before the chunk.
template <int KIND>
__device__ __forceinline__ Sum chunk_summary(const Chunk16& ch, uint32_t prev, uint64_t pos0, uint64_t len,
                                             uint64_t hdr0) {
    Sum r;
    if (KIND == 0) {
        uint32_t t = 0, c_head = 0, c_rest = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint64_t pos = pos0 + i;
            const uint32_t c = ch.byte(i);
            if (pos < len && pos >= hdr0) {
                const bool ls = (pos == hdr0) || (prev == '\n');
                const bool emits = (c != '\n' && c != '\r');
                if (ls) {
                    if (c == '>') { t = 1; c_rest++; }
                    else { t = 2; c_rest += emits; }
                } else if (t == 0) c_head += emits;
                else if (t == 2) c_rest += emits;
            }
            prev = c;
        }
        r.e = t ? (t - 1) * 0x55u : 0xE4u;
        r.c0 = c_rest; r.c1 = c_rest + c_head; r.c2 = 0; r.c3 = 0;
    } else {
        uint32_t nl = 0, a = 0, b = 0;  // a/b: four byte counters indexed by (relative line & 3)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const uint64_t pos = pos0 + i;
            const uint32_t c = ch.byte(i);
            if (pos < len && pos >= hdr0) {
                const bool ls = (pos == hdr0) || (prev == '\n');
                if (c == '\n') nl++;
                else {
                    const uint32_t sh = 8 * (nl & 3);
                    if (c != '\r') a += 1u << sh;
                    if (ls) b += 1u << sh;
                }
            }
            prev = c;
        }
        // incoming line m: sequence bytes are those with (m + r) & 3 == 1, header starts (m + r) & 3 == 0
        r.e = ((0 + nl) & 3) | (((1 + nl) & 3) << 2) | (((2 + nl) & 3) << 4) | (((3 + nl) & 3) << 6);
        r.c0 = ((a >> 8) & 0xFF) + (b & 0xFF);
        r.c1 = (a & 0xFF) + ((b >> 24) & 0xFF);
        r.c2 = ((a >> 24) & 0xFF) + ((b >> 16) & 0xFF);
        r.c3 = ((a >> 16) & 0xFF) + ((b >> 8) & 0xFF);
    }
    return r;
}

// ---- compact FASTA path ------------------------------------------------------------------------
// Inside one parse tile the FASTA transducer fits one u32: bits 0-14 entries emitted after the first line start
// (c_rest), bits 15-29 sequence bytes before it (c_head, emitted only when the chunk starts inside a sequence
// line), bits 30-31 type of the last line start (0 none, 1 header, 2 seq).
constexpr uint32_t kFaH = 15, kFaT = 30, kFaMask = (1u << kFaH) - 1u;
static_assert(kTileBytes <= (int)kFaMask, "the compact FASTA summary counts a tile's entries in 15 bits");
__device__ __forceinline__ uint32_t fa_combine(uint32_t A, uint32_t B) {
    const uint32_t ta = A >> kFaT, tb = B >> kFaT, hb = (B >> kFaH) & kFaMask;
    uint32_t r = (A & ((1u << kFaT) - 1u)) + (B & kFaMask);
    r += (ta == 0) ? (hb << kFaH) : 0u;
    r += (ta == 2) ? hb : 0u;
    return r | ((tb ? tb : ta) << kFaT);
}
__device__ __forceinline__ Sum fa_to_sum(uint32_t a) {
    const uint32_t t = a >> kFaT, c_rest = a & kFaMask, c_head = (a >> kFaH) & kFaMask;
    Sum r; r.e = t ? (t - 1) * 0x55u : 0xE4u; r.c0 = c_rest; r.c1 = c_rest + c_head; r.c2 = 0; r.c3 = 0;
    return r;
}
// ordered scan of compact FASTA summaries over a block of up to 32 warps; s_w holds 2 x warps + 1 words
__device__ __forceinline__ void block_scan_fa(uint32_t mine, uint32_t& excl, uint32_t& total, uint32_t* s_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc = fa_combine(o, inc);
    }
    uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = 0;
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    // warp totals -> exclusive warp prefixes + block total, by the first warp (a serial fold in every thread was a
    // tenth of k_pack's instructions); s_w[nwarp + w] = prefix of warp w, s_w[2 * nwarp] = total
    if (warp == 0) {
        uint32_t v = lane < nwarp ? s_w[lane] : 0u, wi = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi = fa_combine(o, wi);
        }
        uint32_t wp = __shfl_up_sync(0xffffffffu, wi, 1);
        if (lane == 0) wp = 0;
        if (lane < nwarp) s_w[nwarp + lane] = wp;
        if (lane == nwarp - 1) s_w[2 * nwarp] = wi;
    }
    __syncthreads();
    excl = fa_combine(s_w[nwarp + warp], prev);
    total = s_w[2 * nwarp];
    __syncthreads();
}

// One pass over a 16-byte FASTA chunk: compact summary and (EMIT) the packed entries, kept apart
// for the bytes before the chunk's first line start (head) and after it (rest).
struct FaChunk {
    uint32_t sum;
    uint32_t head_c, head_v, rest_c, rest_v, nrec;   // 2-bit codes / validity bits, entry j at bit 2j / j
};
template <bool EMIT>
__device__ __forceinline__ FaChunk fa_chunk(const Chunk16& ch, uint32_t prev, uint64_t pos0, uint64_t len, uint64_t hdr0) {
    // live bytes: [max(hdr0, pos0), min(len, pos0 + 16))
    const uint32_t lo = hdr0 > pos0 ? (uint32_t)min((uint64_t)16, hdr0 - pos0) : 0u;
    const uint32_t hi = len > pos0 ? (uint32_t)min((uint64_t)16, len - pos0) : 0u;
    const uint32_t live = hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
    const uint32_t force_ls = (hdr0 >= pos0 && hdr0 < pos0 + 16) ? (1u << (uint32_t)(hdr0 - pos0)) : 0u;
    uint32_t t = 0, hn = 0, rn = 0;
    FaChunk r; r.head_c = r.head_v = r.rest_c = r.rest_v = r.nrec = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t c = ch.byte(i);
        if ((live >> i) & 1u) {
            const bool ls = ((force_ls >> i) & 1u) || (prev == '\n');
            const bool emits = (c != '\n' && c != '\r');
            const bool ok = is_acgt(c);
            const uint32_t code = ok ? ((c >> 1) & 3u) : 0u;
            if (ls) t = (c == '>') ? 1u : 2u;
            if (ls && c == '>') { rn++; if (EMIT) r.nrec++; }            // record break: one invalid entry
            else if (emits && t != 1u) {
                if (t == 0) {
                    if (EMIT) { r.head_c |= code << (2 * hn); r.head_v |= (uint32_t)ok << hn; }
                    hn++;
                } else {
                    if (EMIT) { r.rest_c |= code << (2 * rn); r.rest_v |= (uint32_t)ok << rn; }
                    rn++;
                }
            }
        }
        prev = c;
    }
    r.sum = rn | (hn << kFaH) | (t << kFaT);
    return r;
}

// ---- SWAR fast path for FASTA text --------------------------------------------------------------
// A 16-byte chunk whose only bytes below 0x40 are '\n' (no '>', '\r', digits, blanks ...) is classified four
// bytes at a time: newline mask, ACGT validity mask and 2-bit codes come out of a handful of word operations
// instead of a 16-iteration byte loop.  Everything else (header lines, CR-LF text, the first / last bytes of a
// file) goes through fa_chunk().
__device__ __forceinline__ uint32_t swar_zero_bytes(uint32_t v) {      // 0x80 in every byte of v that is zero (exact)
    return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v) & 0x80808080u;
}
__device__ __forceinline__ uint32_t swar_movemask(uint32_t m) {         // bits 7, 15, 23, 31 -> bits 0..3
    return (m * 0x00204081u) >> 28;
}

// Line structure first, bases second.  The ALU pipe is the bound (LOP3 / SHF / PRMT issue every other cycle, IMAD goes to
// the FMA pipe), so the tests are phrased with as few of those as possible.
struct FaLines { uint32_t nl; bool fast; uint32_t nlb[4]; };   // nl: 16-bit newline mask; fast: see fa_lines; nlb: 0xFF in every byte below 0x40

__device__ __forceinline__ uint32_t prmt_sign(uint32_t x) {              // 0xFF in every byte of x whose bit 7 is set
    uint32_t d;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(d) : "r"(x));
    return d;
}

__device__ __forceinline__ bool fa_fast_ok(uint64_t pos0, uint64_t len, uint64_t hdr0) {
    return pos0 > hdr0 && pos0 + 16 <= len;          // every byte live, no forced line start inside
}

// fast <=> every byte of the chunk is live and its only bytes below 0x40 are '\n' (no '>', '\r', digits, blanks ...):
// then nl is the newline mask and the chunk is handled four bytes per instruction.  Everything else (header lines,
// CR-LF text, the first / last bytes of a file) goes through the byte loop fa_chunk().
__device__ __forceinline__ FaLines fa_lines(const Chunk16& ch, uint64_t pos0, uint64_t len, uint64_t hdr0) {
    FaLines r; r.nl = 0;
    uint32_t bad = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t x = ch.w[w];
        const uint32_t low = ~(x | (x * 2u)) & 0x80808080u;              // bit 7 of every byte below 0x40
        r.nlb[w] = prmt_sign(low);
        bad |= (x ^ 0x0A0A0A0Au) & r.nlb[w];                             // ... that is not a newline
        r.nl |= swar_movemask(low) << (4 * w);
    }
    r.fast = bad == 0 && fa_fast_ok(pos0, len, hdr0);
    return r;
}

// validity mask (ACGT, either case) and 2-bit codes (position j at bits 2j) of a chunk
__device__ __forceinline__ void fa_bases(const Chunk16& ch, uint32_t& valid, uint32_t& codes) {
    valid = codes = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t x = ch.w[w];
        const uint32_t c = (x >> 1) & 0x03030303u;                      // A0 C1 T2 G3
        codes |= ((c * 0x01041040u) >> 24) << (8 * w);
        const uint32_t t = c | (c >> 4);                                // nibble pairs c0|c1<<4 in byte 0, c2|c3<<4 in byte 2
        const uint32_t sel = __byte_perm(t, 0u, 0x4420);                // selector nibbles c0, c1, c2, c3
        const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);      // 'A' 'C' 'T' 'G' by code
        valid |= swar_movemask(swar_zero_bytes((x & 0xDFDFDFDFu) ^ expect)) << (4 * w);
    }
}

// The same for a chunk that fa_lines() found fast (its only bytes below 0x40 are newlines, nlb marks them): nearly
// every such chunk of a genome is all ACGT, so the per-byte zero tests and the four mask gathers are only done when
// some byte that is not a newline differs from the letter its code stands for.  (The validity bits at newline
// positions are never looked at: fa_chunk_parts squeezes those positions out.)
__device__ __forceinline__ void fa_bases_fast(const Chunk16& ch, const uint32_t (&nlb)[4], uint32_t& valid, uint32_t& codes) {
    codes = 0;
    uint32_t d[4], any = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        const uint32_t x = ch.w[w];
        const uint32_t c = (x >> 1) & 0x03030303u;                      // A0 C1 T2 G3
        codes |= ((c * 0x01041040u) >> 24) << (8 * w);
        const uint32_t t = c | (c >> 4);
        const uint32_t sel = __byte_perm(t, 0u, 0x4420);
        const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);      // 'A' 'C' 'T' 'G' by code
        d[w] = (x & 0xDFDFDFDFu) ^ expect;
        any |= d[w] & ~nlb[w];
    }
    valid = 0xFFFFu;
    if (any) {
        valid = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) valid |= swar_movemask(swar_zero_bytes(d[w])) << (4 * w);
    }
}

// ---- SWAR path for FASTQ text ---------------------------------------------------------------------
// A live chunk whose only control characters are '\n' is classified without a byte loop.  The line index of every
// byte relative to the chunk's first byte (mod 4) comes from two prefix parities of the newline mask (bit 0: newlines
// before the byte; bit 1: newlines that arrived while bit 0 was set), the four per-state entry counts are popcounts,
// and the entries of the chunk for a known state are the SWAR base classification restricted to the sequence line's
// bytes plus one break entry per header-line start.
struct FqChunk { uint32_t nl, e0, e1, ls; };
__device__ __forceinline__ uint32_t prefix_parity16(uint32_t v) {       // bit i = parity of bits 0..i
    v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8;
    return v;
}
__device__ __forceinline__ FqChunk fq_lines(uint32_t nl, bool first_is_line_start) {
    FqChunk f;
    f.nl = nl;
    f.e0 = (prefix_parity16(nl) << 1) & 0xFFFFu;
    f.e1 = (prefix_parity16(nl & f.e0) << 1) & 0xFFFFu;
    f.ls = ((nl << 1) | (first_is_line_start ? 1u : 0u)) & 0xFFFFu;
    return f;
}
__device__ __forceinline__ uint32_t fq_line_mask(const FqChunk& f, uint32_t r) {     // bytes of relative line r (mod 4)
    return ((r & 1u) ? f.e0 : ~f.e0) & ((r & 2u) ? f.e1 : ~f.e1) & 0xFFFFu;
}
__device__ __forceinline__ Sum fq_summary(const FqChunk& f) {
    const uint32_t nn = ~f.nl & 0xFFFFu;
    uint32_t a[4], b[4];
#pragma unroll
    for (uint32_t r = 0; r < 4; ++r) {
        const uint32_t m = fq_line_mask(f, r) & nn;
        a[r] = (uint32_t)__popc(m);                 // bytes that would be sequence entries if line r is a sequence line
        b[r] = (uint32_t)__popc(m & f.ls);          // line starts that would be record breaks if line r is a header line
    }
    const uint32_t n = (uint32_t)__popc(f.nl);
    Sum s;
    s.e = ((0 + n) & 3) | (((1 + n) & 3) << 2) | (((2 + n) & 3) << 4) | (((3 + n) & 3) << 6);
    s.c0 = a[1] + b[0]; s.c1 = a[0] + b[3]; s.c2 = a[3] + b[2]; s.c3 = a[2] + b[1];
    return s;
}
// positions of the chunk's control characters and of its newlines (16-bit masks)
__device__ __forceinline__ void fq_control(const Chunk16& ch, uint32_t& ctrl, uint32_t& nl) {
    ctrl = nl = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        ctrl |= swar_movemask(swar_zero_bytes(ch.w[w] & 0xE0E0E0E0u)) << (4 * w);
        nl |= swar_movemask(swar_zero_bytes(ch.w[w] ^ 0x0A0A0A0Au)) << (4 * w);
    }
}

// Entries of one chunk, kept apart for the bytes before its first line start (head: emitted only when the chunk
// starts inside a sequence line) and after it (rest).  Compacted: entry j at bits 2j / j.
struct FaParts {
    uint32_t hc, rc;        // head / rest codes
    uint32_t hv_rv;         // head validity (low 16) | rest validity << 16
    uint32_t meta;          // hn | rn << 8 | t << 16 | nrec << 20   (t: type of the last line start, 0 none 1 header 2 seq)
    __device__ __forceinline__ uint32_t hn() const { return meta & 0xFFu; }
    __device__ __forceinline__ uint32_t rn() const { return (meta >> 8) & 0xFFu; }
    __device__ __forceinline__ uint32_t t() const { return (meta >> 16) & 0xFu; }
    __device__ __forceinline__ uint32_t nrec() const { return meta >> 20; }
    __device__ __forceinline__ uint32_t sum() const { return rn() | (hn() << kFaH) | (t() << kFaT); }
};

// the byte-loop versions stay out of line: they are rare, and inlining four copies of them costs ~200 registers
__device__ __noinline__ FaParts fa_chunk_parts_slow(const Chunk16& ch, uint32_t prev, uint64_t pos0, uint64_t len, uint64_t hdr0) {
    const FaChunk fc = fa_chunk<true>(ch, prev, pos0, len, hdr0);
    FaParts r;
    r.hc = fc.head_c; r.rc = fc.rest_c; r.hv_rv = fc.head_v | (fc.rest_v << 16);
    r.meta = ((fc.sum >> kFaH) & kFaMask) | ((fc.sum & kFaMask) << 8) | ((fc.sum >> kFaT) << 16) | (fc.nrec << 20);
    return r;
}

// entries of one chunk
__device__ __forceinline__ FaParts fa_chunk_parts(const Chunk16& ch, const FaLines& ln, uint32_t prev, uint64_t pos0, uint64_t len,
                                                  uint64_t hdr0) {
    FaParts r;
    if (ln.fast) {
        uint32_t valid, codes;
        fa_bases_fast(ch, ln.nlb, valid, codes);
        const uint32_t ls = ((ln.nl << 1) | (prev == '\n')) & 0xFFFFu;
        if (ls == 0) {
            const uint32_t hn = 16u - (ln.nl >> 15);           // a newline can only sit at position 15
            r.hc = hn == 16 ? codes : (codes & 0x3FFFFFFFu);
            r.rc = 0; r.hv_rv = hn == 16 ? valid : (valid & 0x7FFFu); r.meta = hn;
            return r;
        }
        const uint32_t q = __ffs(ls) - 1;
        const uint32_t hn = q ? q - 1 : 0u;                    // the byte before the line start is its newline
        r.hc = codes & ((1u << (2 * hn)) - 1u);
        const uint32_t hv = valid & ((1u << hn) - 1u);
        uint32_t rc = codes >> (2 * q), rv = valid >> q, rnl = ln.nl >> q, rn = 16u - q;
        while (rnl) {                                          // further newlines: lines shorter than the chunk
            const uint32_t pz = 31u - __clz(rnl);
            const uint32_t lo_c = (1u << (2 * pz)) - 1u, lo_v = (1u << pz) - 1u;
            rc = (rc & lo_c) | ((rc >> 2) & ~lo_c);
            rv = (rv & lo_v) | ((rv >> 1) & ~lo_v);
            rnl ^= 1u << pz;
            --rn;
        }
        r.rc = rc; r.hv_rv = hv | (rv << 16); r.meta = hn | (rn << 8) | (2u << 16);
        return r;
    }
    return fa_chunk_parts_slow(ch, prev, pos0, len, hdr0);
}

// per-thread accumulator of up to 64 packed entries
struct Acc64 {
    unsigned long long clo, chi, v;
    uint32_t n;
    __device__ __forceinline__ void init() { clo = chi = v = 0; n = 0; }
    // codes: 2*cnt significant bits, vbits: cnt significant bits, upper bits zero
    __device__ __forceinline__ void append(uint32_t codes, uint32_t vbits, uint32_t cnt) {
        if (cnt == 0) return;
        const uint32_t s = 2 * n;
        if (s < 64) { clo |= (unsigned long long)codes << s; if (s > 32) chi |= (unsigned long long)codes >> (64 - s); }
        else chi |= (unsigned long long)codes << (s - 64);
        v |= (unsigned long long)vbits << n;
        n += cnt;
    }
};

// exclusive scan of one u32 per thread over a 1024-thread block; s_warp must hold 33 words
__device__ __forceinline__ uint32_t block_excl_scan_1024(uint32_t v, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        s_warp[lane] = wi - w;          // exclusive warp offsets
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    total = s_warp[32];
    const uint32_t r = s_warp[warp] + inc - v;
    __syncthreads();
    return r;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}

// reverse the order of the 32 2-bit groups of x
__device__ __forceinline__ uint64_t rev2(uint64_t x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

}  // namespace grmkm
