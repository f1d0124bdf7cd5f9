// grmkm_units.cuh -- the super-k-mer ("unit") path of the k-mer matrix build.
//
// The genomes Kover learns from belong to one species: most of every genome is shared with the others.  A
// *unit* is a maximal run of consecutive valid k-mers of one record that have the same minimizer (smallest
// strand-neutral m-mer hash), at most lmax k-mers long, stored as its L + k - 1 bases in the strand with the
// smaller value.  Unit boundaries depend on the sequence only, so the same stretch of DNA yields the same
// units in every genome, whatever its offset, contig layout or strand.  The pipeline is then
//
//   k_unit_bounds     packed stream -> per position "a unit starts here" / "a valid k-mer ends here" bits
//   k_units_scatter   units (16 bytes each, ~11 k-mers at k = 31) -> buckets by a hash of the unit's content
//   k_units_dedupe    per bucket: shared-memory table keyed by (content, 64-genome block) -> one presence word
//   k_units_expand    every DISTINCT (unit, block) -> its k-mers -> records [hash, presence word] by hash range
//   k_aggregate_cols  (grmkm_kernels.cuh, wide input) merges the words of equal k-mers into columns
//
// so the per-occurrence work is one table insert per ~11 k-mers, and the per-k-mer work is done once per
// distinct unit instead of once per genome.  Unit boundaries only affect how much is shared, never the
// result: every valid k-mer window lies in exactly one unit (checked through stats.n_windows).
// Reference semantics are unchanged (multidsk / dsk2kover, kmer_count.py:28-37, kmer_pack.py:28-36).
#pragma once
#include "grmkm_kernels.cuh"

namespace grmkm {

// ---- geometry -------------------------------------------------------------------------------
// unit = L k-mers (1 <= L <= lmax), nb = L + k - 1 <= 53 bases.
//   lo = bases 0..28 (entry j at bits 2j) | (L - 1) << 58
//   hi = bases 29..52                     | row << 48
struct UnitGeom { uint32_t k, m, w, lmax; };

__host__ __device__ inline UnitGeom unit_geom(uint32_t k) {
    UnitGeom g;
    g.k = k;
    g.lmax = 54 - k < 32 ? 54 - k : 32;
    const uint32_t mmin = k < 5 ? k : 5;
    uint32_t wmax = k - mmin + 1;
    if (wmax > g.lmax) wmax = g.lmax;
    g.w = wmax >= 21 ? 21 : wmax >= 13 ? 13 : wmax >= 7 ? 7 : wmax >= 4 ? 4 : wmax >= 2 ? 2 : 1;
    g.m = k - g.w + 1;       // k-mer = w consecutive m-mers
    return g;
}

constexpr unsigned long long kUnitLoMask = (1ULL << 58) - 1;
constexpr unsigned long long kUnitHiMask = (1ULL << 48) - 1;
constexpr unsigned long long kUnitEmptyLo = ~0ULL;     // bit 63 of lo is always 0 (L - 1 <= 31)
constexpr unsigned long long kUnitEmptyHi = ~0ULL;     // block field 0xFFFF never occurs
constexpr uint32_t kMmerSeed = 0x5bd1e995u, kMmerMul = 0x9E3779B1u;

__device__ __forceinline__ unsigned long long rev2_64(unsigned long long x) {
    x = __brevll(x);
    return ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
}

// hash of a unit's content (lo, hi without the row), 32-bit arithmetic only: the scatter takes the bucket from
// the top bits (mulhi), the dedupe table takes its slot from a second multiply
__device__ __forceinline__ uint32_t unit_hash32(uint32_t lo0, uint32_t lo1, uint32_t hi0, uint32_t hi1) {
    uint32_t h = (lo0 * 0x9E3779B1u) ^ (lo1 * 0x85EBCA77u) ^ (hi0 * 0xC2B2AE3Du) ^ (hi1 * 0x27D4EB2Fu);
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    return h ^ (h >> 13);
}

// ---- k_unit_bounds: per group of 32 stream entries, start mask and valid-k-mer mask ---------
// Bit e of .y: the k-mer ENDING at entry e of the group is valid.  Bit e of .x: it is valid and it starts a
// run (the previous k-mer is invalid or has another minimizer).  W = m-mers per k-mer (compile time).
//
// The kernel is bound by the integer ALU pipe (LOP3 / SHF / VIMNMX / ISETP issue every other cycle per scheduler;
// IMAD goes to the FMA pipe), so the work is arranged to need few of those per position (profiles/r01_v3: 16.5):
//  * a thread takes TWO adjacent groups, so neighbouring windows share their m-mer hashes (64 + W instead of
//    2 x (32 + W)), and both groups go through the sliding minimum as the two halves of 16x2 SIMD words;
//  * the m-mer needs no mask: it sits in the low 2m bits of a funnel shift, and multiplying by (mul << (32 - 2m))
//    discards everything above them; strand neutrality comes from adding the two strands' m-mers BEFORE that
//    multiply (two IMADs);
//  * a run ends at k-mer e when an m-mer equal to the window minimum leaves on the left or enters on the right:
//    with inner = min of the W - 1 m-mers both windows share, that is  min(h[e], h[e + W]) <= inner  -- one
//    __vibmin_u16x2, whose predicate outputs are the answer for both groups.  It covers every change of the
//    minimum, reads the same from either strand, and bounds a run by W <= lmax.
// masks of the groups g and g + 1 from their code / validity words and those of group g - 1
template <int W>
__device__ __forceinline__ void unit_bounds_pair(unsigned long long cp, unsigned long long c0, unsigned long long c1, uint32_t vp, uint32_t v0,
                                                 uint32_t v1, uint32_t k, uint32_t m, uint2& ma, uint2& mb) {
    // valid k-mers: smear the invalid entries over the k - 1 entries that follow them
    unsigned long long sa = ~((unsigned long long)vp | ((unsigned long long)v0 << 32));
    unsigned long long sb = ~((unsigned long long)v0 | ((unsigned long long)v1 << 32));
    for (uint32_t c = 1; c < k;) { const uint32_t d = c < k - c ? c : k - c; sa |= sa << d; sb |= sb << d; c += d; }
    const uint32_t vka = ~(uint32_t)(sa >> 32), vkb = ~(uint32_t)(sb >> 32);
    ma = mb = make_uint2(0u, 0u);
    if ((vka | vkb) == 0) return;
    const uint32_t vka_before = (uint32_t)((sa >> 31) & 1ULL) ^ 1u;             // the k-mer ending at the last entry of group g - 1
    // 96-entry window [group g - 1 | g | g + 1]: entry q at bits 2q of x[]
    const uint32_t x[6] = {(uint32_t)cp, (uint32_t)(cp >> 32), (uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)c1, (uint32_t)(c1 >> 32)};
    // forward strand, first base most significant: digit p of r[] = entry 95 - p
    const uint32_t r[7] = {rev2_32(x[5]), rev2_32(x[4]), rev2_32(x[3]), rev2_32(x[2]), rev2_32(x[1]), rev2_32(x[0]), 0u};
    // reverse complement: digit j of y[] = complement of entry j + 32 - k
    const uint32_t S0 = 2 * (32 - k), sw = S0 >> 5, sb5 = S0 & 31;
    uint32_t y[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const uint32_t lo = i + sw < 6 ? x[i + sw] : 0u, hi = i + sw + 1 < 6 ? x[i + sw + 1] : 0u;
        y[i] = __funnelshift_r(lo, hi, sb5) ^ 0xAAAAAAAAu;
    }
    const uint32_t mul = kMmerMul << (32 - 2 * m);      // keeps the low 2m bits of its operand only
    const uint32_t seed = kMmerSeed * mul;
    // P[t] = 16-bit hashes of the m-mers ending at window entries 32 - W + t (low half: group g) and 64 - W + t (high half)
    constexpr int N = 32 + W;
    uint32_t P[N];
    {
        uint32_t h[64 + W];
#pragma unroll
        for (int t = 0; t < 64 + W; ++t) {
            const int fs = 2 * (63 + W - t), rs = 2 * t;
            const uint32_t fw = __funnelshift_r(r[fs >> 5], r[(fs >> 5) + 1], fs & 31);      // m-mer in the low 2m bits
            const uint32_t rc = __funnelshift_r(y[rs >> 5], y[(rs >> 5) + 1], rs & 31);
            h[t] = fw * mul + (rc * mul + seed);        // = (fw + rc + seed) * mul: the same for an m-mer and its reverse complement
        }
#pragma unroll
        for (int t = 0; t < N; ++t) P[t] = __byte_perm(h[t], h[t + 32], 0x7632);
    }
    // inner(e) = min P[e + 1 .. e + W - 1]: prefix / suffix minima over blocks of V = W - 1 starting at index 1
    constexpr int V = W > 1 ? W - 1 : 1;
    uint32_t changed_a = 0, changed_b = 0;
    if (W == 1) { changed_a = changed_b = 0xFFFFFFFFu; }        // every k-mer is its own m-mer
    else {
        constexpr int HI = 30 + W;                      // last index a window touches
        uint32_t pre[HI + 1], suf[HI + 1];
#pragma unroll
        for (int i = 1; i <= HI; ++i) pre[i] = ((i - 1) % V == 0) ? P[i] : __vminu2(pre[i - 1], P[i]);
#pragma unroll
        for (int i = HI; i >= 1; --i) suf[i] = ((i - 1) % V == V - 1 || i == HI) ? P[i] : __vminu2(suf[i + 1], P[i]);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
            const uint32_t inner = __vminu2(suf[e + 1], pre[e + V]);
            bool hi, lo;
            __vibmin_u16x2(__vminu2(P[e], P[e + W]), inner, &hi, &lo);      // hi / lo = (ends <= inner) per half
            changed_a |= (uint32_t)lo << e;
            changed_b |= (uint32_t)hi << e;
        }
    }
    const uint32_t before_a = (vka << 1) | vka_before, before_b = (vkb << 1) | (vka >> 31);
    ma = make_uint2(vka & (~before_a | changed_a), vka);
    mb = make_uint2(vkb & (~before_b | changed_b), vkb);
}

// Grid-stride over the group pairs with the NEXT pair's words fetched before the current pair is worked on: at 91
// registers only 3 - 4 warps per scheduler are resident, and with one pair per thread a fifth of the kernel's stall
// samples sat on the first loads (profiles/r02_v8 source page: line of `cp = codes[g - 1]`).
template <int W>
__global__ void __launch_bounds__(256)
k_unit_bounds(const unsigned long long* __restrict__ codes, const uint32_t* __restrict__ valid,
              const uint64_t* __restrict__ scalars, uint32_t k, uint32_t m, uint2* __restrict__ masks) {
    const uint64_t n_groups = (scalars[S_STREAM_LEN] + 31) >> 5;
    const uint64_t stride = (uint64_t)gridDim.x * 512;
    uint64_t g = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 2;                // groups g and g + 1
    unsigned long long c0 = 0, c1 = 0, cp = 0;
    uint32_t v0 = 0, v1 = 0, vp = 0;
    auto fetch = [&](uint64_t q, unsigned long long& a0, unsigned long long& a1, unsigned long long& ap, uint32_t& b0, uint32_t& b1, uint32_t& bp) {
        a0 = a1 = ap = 0ULL; b0 = b1 = bp = 0u;
        if (q < n_groups) {
            a0 = codes[q]; b0 = valid[q];
            if (q + 1 < n_groups) { a1 = codes[q + 1]; b1 = valid[q + 1]; }
            if (q > 0) { ap = codes[q - 1]; bp = valid[q - 1]; }
        }
    };
    fetch(g, c0, c1, cp, v0, v1, vp);
    while (g < n_groups) {
        unsigned long long n0, n1, np;
        uint32_t w0, w1, wp;
        fetch(g + stride, n0, n1, np, w0, w1, wp);
        uint2 ma, mb;
        unit_bounds_pair<W>(cp, c0, c1, vp, v0, v1, k, m, ma, mb);
        masks[g] = ma;
        if (g + 1 < n_groups) masks[g + 1] = mb;
        g += stride; c0 = n0; c1 = n1; cp = np; v0 = w0; v1 = w1; vp = wp;
    }
}

// ---- k_units_scatter: units -> content-hash buckets ------------------------------------------
// One tile = 1024 groups (32768 stream entries) per CTA iteration.  The unit starts of the tile are compacted
// into a list (block scan of the start-mask popcounts), so every thread handles the same number of units
// whatever their distribution over the groups; a round takes up to kUsMax units per thread (more units =
// more rounds, so any input is handled with bounded shared memory).  Round: histogram with returning
// shared-memory atomics (the rank inside the bucket's run) -> scan + ONE global reservation per (round,
// bucket) -> units to their sorted slot -> copy out in bucket runs.
constexpr int kUsThreads = 1024;
constexpr int kUsMax = 4;
constexpr int kUsStage = kUsThreads * kUsMax;
constexpr int kUsTileGroups = kUsThreads;
constexpr int kUsTileEntries = kUsTileGroups * 32;
constexpr int kUsMaxBuckets = 1024;

struct UnitScatterParams {
    const unsigned long long* codes;
    const uint2* masks;
    const uint64_t* scalars;
    const uint64_t* file_stream_start;
    const FileDesc* files;
    const uint32_t* tile_file;          // file of the first entry of every tile
    uint32_t n_files;
    uint32_t k, lmax, n_buckets;
    unsigned long long* cursors;        // [n_buckets]; COUNT: histogram
    uint4* units;
    unsigned long long cap;             // units per bucket region, 0 = exact offsets
    unsigned long long dump;            // index of the dump area (kUsStage units)
    unsigned long long* overflow;
    unsigned long long* n_windows;
};

constexpr int kUsCodeWords = kUsTileGroups + 3;          // staged 64-bit code words: the group before the tile .. two after it
constexpr int kUsCode32 = 2 * kUsCodeWords + 6;          // as 32-bit words, padded for the 5-word window reads

__host__ __device__ inline size_t unit_scatter_smem(uint32_t MB) {
    return (size_t)kUsStage * 16 + (size_t)MB * 8 + (size_t)kUsCode32 * 4 * 2 + (size_t)MB * 8 +
           (size_t)(kUsTileGroups + 1) * 4 + (size_t)kUsStage * 2 + (size_t)kUsTileEntries * 2 + (size_t)kUsTileGroups * 2;
}

// 2 * nb bits starting at entry a of a staged 2-bit array, as four 32-bit words (not yet masked)
__device__ __forceinline__ void unit_window(const uint32_t* __restrict__ s32, uint32_t a, uint32_t (&v)[4]) {
    const uint32_t wi = a >> 4, sh = 2u * (a & 15u);
    const uint32_t t0 = s32[wi], t1 = s32[wi + 1], t2 = s32[wi + 2], t3 = s32[wi + 3], t4 = s32[wi + 4];
    v[0] = __funnelshift_r(t0, t1, sh); v[1] = __funnelshift_r(t1, t2, sh);
    v[2] = __funnelshift_r(t2, t3, sh); v[3] = __funnelshift_r(t3, t4, sh);
}

__global__ void k_stream_tile_files(const uint64_t* __restrict__ scalars, const uint64_t* __restrict__ fss, uint32_t n_files,
                                    uint32_t* __restrict__ tile_file, uint64_t max_tiles, uint64_t tile_entries) {
    const uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= max_tiles) return;
    const uint64_t p0 = tile * tile_entries;
    if (p0 >= scalars[S_STREAM_LEN]) return;
    uint32_t lo = 0, hi = n_files - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (fss[mid] <= p0) lo = mid; else hi = mid - 1;
    }
    tile_file[tile] = lo;
}

template <bool COUNT>
__global__ void __launch_bounds__(kUsThreads, 1)
k_units_scatter(const UnitScatterParams p) {
    extern __shared__ uint4 s_units[];                                   // [kUsStage] sorted units of the round
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_warp[33];
    __shared__ uint4 s_mask[33];                                          // [L] masks of a unit of L k-mers (L + k - 1 bases)
    const uint32_t MB = p.n_buckets;
    unsigned long long* s_delta = reinterpret_cast<unsigned long long*>(s_units + kUsStage);   // [MB]
    uint32_t* s_fw = reinterpret_cast<uint32_t*>(s_delta + MB);          // [kUsCode32] codes, word 0 = the group before the tile
    uint32_t* s_rc = s_fw + kUsCode32;                                   // [kUsCode32] the same entries reversed and complemented
    uint32_t* s_cnt = s_rc + kUsCode32;                                  // [MB]
    uint32_t* s_off = s_cnt + MB;                                        // [MB]
    uint32_t* s_brk = s_off + MB;                                        // [kUsTileGroups + 1] natural breaks per group
    uint16_t* s_b = reinterpret_cast<uint16_t*>(s_brk + kUsTileGroups + 1);   // [kUsStage]
    uint16_t* s_list = s_b + kUsStage;                                   // [kUsTileEntries] tile-local entry where a unit's first k-mer ends
    uint16_t* s_gf = s_list + kUsTileEntries;                            // [kUsTileGroups] genome row of the group (groups with unit starts)
    const uint64_t stream_len = p.scalars[S_STREAM_LEN];
    const uint64_t n_groups = (stream_len + 31) >> 5;
    const uint64_t n_tiles = (n_groups + kUsTileGroups - 1) / kUsTileGroups;
    const uint32_t k = p.k, lmax = p.lmax;
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t T = (uint32_t)kUsCodeWords * 32u;                 // staged entries
    for (uint32_t i = tid; i < MB; i += kUsThreads) s_cnt[i] = 0;
    if (tid < 6) { s_fw[2 * kUsCodeWords + tid] = 0; s_rc[2 * kUsCodeWords + tid] = 0; }
    if (tid < 33) {
        const uint32_t bits = 2u * (tid + k - 1u);                        // <= 106
        uint4 mk;
        mk.x = bits >= 32u ? 0xFFFFFFFFu : ((1u << bits) - 1u);
        mk.y = bits >= 64u ? 0xFFFFFFFFu : (bits > 32u ? ((1u << (bits - 32u)) - 1u) : 0u);
        mk.z = bits >= 96u ? 0xFFFFFFFFu : (bits > 64u ? ((1u << (bits - 64u)) - 1u) : 0u);
        mk.w = bits > 96u ? ((1u << (bits - 96u)) - 1u) : 0u;
        s_mask[tid] = mk;
    }
    unsigned long long my_windows = 0;
    // this thread's share of a tile's inputs: code words tid (and tid + 1024 for the first three threads), one mask
    // ... and the tile's file: its index, where the NEXT file starts in the stream, its genome row (three dependent
    // loads: fetched a tile ahead like the rest, they were 9 % of the kernel's stall samples on the critical path)
    uint32_t n_tf = 0, n_row = 0; unsigned long long n_next = ~0ULL;
    uint32_t pf_tf = blockIdx.x < n_tiles ? p.tile_file[blockIdx.x] : 0u;      // the file index travels TWO tiles ahead
    auto fetch = [&](uint64_t tile, unsigned long long& c0, unsigned long long& c1, uint2& mk) {
        const uint64_t g0 = tile * kUsTileGroups;
        const uint64_t gi = g0 + tid;        // code word i holds group g0 + i - 1
        c0 = (tile < n_tiles && gi >= 1 && gi - 1 < n_groups) ? p.codes[gi - 1] : 0ULL;
        c1 = (tile < n_tiles && tid < 3 && gi + kUsThreads - 1 < n_groups) ? p.codes[gi + kUsThreads - 1] : 0ULL;
        mk = (tile < n_tiles && gi < n_groups) ? p.masks[gi] : make_uint2(0u, 0u);
        if (tile < n_tiles) {
            n_tf = pf_tf;
            n_next = n_tf + 1 < p.n_files ? p.file_stream_start[n_tf + 1] : ~0ULL;
            n_row = p.files[n_tf].row;
            if (tile + gridDim.x < n_tiles) pf_tf = p.tile_file[tile + gridDim.x];
        }
    };
    unsigned long long nc0, nc1; uint2 nmk;
    fetch(blockIdx.x, nc0, nc1, nmk);
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t g0 = tile * kUsTileGroups, g = g0 + tid;
        const uint2 mk = nmk;
        {
            reinterpret_cast<unsigned long long*>(s_fw)[tid] = nc0;
            reinterpret_cast<unsigned long long*>(s_rc)[kUsCodeWords - 1 - tid] = rev2_64(nc0) ^ 0xAAAAAAAAAAAAAAAAULL;
            if (tid < 3) {
                reinterpret_cast<unsigned long long*>(s_fw)[tid + kUsThreads] = nc1;
                reinterpret_cast<unsigned long long*>(s_rc)[kUsCodeWords - 1 - tid - kUsThreads] = rev2_64(nc1) ^ 0xAAAAAAAAAAAAAAAAULL;
            }
        }
        s_brk[tid] = ~mk.y | mk.x;           // after a position: an invalid k-mer or the start of another run
        if (tid == 0) {
            uint2 mn = make_uint2(0u, 0u);
            if (g0 + kUsTileGroups < n_groups) mn = p.masks[g0 + kUsTileGroups];
            s_brk[kUsTileGroups] = ~mn.y | mn.x;
        }
        if (mk.x) {        // files start on group boundaries (grmkm_api.cu: stream_cap), so a group lies in ONE file
            uint32_t row = n_row;
            if (g * 32ULL >= n_next) {       // (a tile that holds the start of another file)
                uint32_t f = n_tf + 1;
                while (f + 1 < p.n_files && p.file_stream_start[f + 1] <= g * 32ULL) ++f;
                row = p.files[f].row;
            }
            s_gf[tid] = (uint16_t)row;
        }
        // compact the unit starts of the tile
        uint32_t n_units;
        {
            uint32_t o = block_excl_scan<kUsThreads>((uint32_t)__popc(mk.x), s_warp, n_units);
            for (uint32_t m = mk.x; m; m &= m - 1u) s_list[o++] = (uint16_t)(32u * tid + (uint32_t)__ffs(m) - 1u);
        }
        __syncthreads();       // the tile's staging is complete
        fetch(tile + gridDim.x, nc0, nc1, nmk);        // the next tile's inputs travel while this one is processed
        for (uint32_t base = 0; base < n_units; base += kUsStage) {
            uint32_t u0[kUsMax], u1[kUsMax], u2[kUsMax], u3[kUsMax], ubr[kUsMax];
#pragma unroll
            for (int j = 0; j < kUsMax; ++j) {
                const uint32_t idx = base + j * kUsThreads + tid;
                ubr[j] = 0xFFFFFFFFu;
                if (idx < n_units) {
                    const uint32_t P = s_list[idx];              // tile-local entry where the first k-mer ends
                    const uint32_t gl = P >> 5, e = P & 31u;
                    // distance to the next break after entry P; only the next lmax <= 32 positions matter, i.e. the low word
                    // of the 64 break bits shifted down by e + 1 (one funnel shift, one 32-bit find-first-set)
                    const uint32_t after = __funnelshift_rc(s_brk[gl], s_brk[gl + 1], e + 1u);      // (clamped: a shift of 32 yields the high word)
                    const uint32_t dist = after ? (uint32_t)__ffs((int)after) : 33u;
                    const uint32_t L = dist < lmax ? dist : lmax;
                    const uint32_t a = P + 32u - (k - 1u);       // first entry of the unit, counted from the group before the tile
                    const uint32_t nb = L + k - 1u;
                    // The unit is stored in the strand whose words (first 16 bases first) compare smaller -- any rule works
                    // that picks the same strand from either side.  Word 0 of both strands decides almost always, so only
                    // the winner's other three words are extracted.
                    const uint4 mk4 = s_mask[L];                 // low 2 * nb bits of the four words
                    const uint32_t ar = T - a - nb;              // the other strand: the same entries reversed and complemented
                    const uint32_t wf = a >> 4, sf = 2u * (a & 15u), wr = ar >> 4, sr = 2u * (ar & 15u);
                    const uint32_t f0 = __funnelshift_r(s_fw[wf], s_fw[wf + 1], sf) & mk4.x;
                    const uint32_t r0 = __funnelshift_r(s_rc[wr], s_rc[wr + 1], sr) & mk4.x;
                    bool use_rc = r0 < f0;
                    if (f0 == r0) {                              // (palindromic starts: compare the rest)
                        uint32_t fv[4], rv[4];
                        unit_window(s_fw, a, fv);
                        unit_window(s_rc, ar, rv);
                        fv[1] &= mk4.y; fv[2] &= mk4.z; fv[3] &= mk4.w; rv[1] &= mk4.y; rv[2] &= mk4.z; rv[3] &= mk4.w;
                        use_rc = rv[1] != fv[1] ? rv[1] < fv[1] : rv[2] != fv[2] ? rv[2] < fv[2] : rv[3] < fv[3];
                    }
                    const uint32_t* const src = use_rc ? s_rc : s_fw;
                    const uint32_t wi = use_rc ? wr : wf, sh = use_rc ? sr : sf;
                    uint32_t v[4];
                    {
                        const uint32_t t1 = src[wi + 1], t2 = src[wi + 2], t3 = src[wi + 3], t4 = src[wi + 4];
                        v[0] = use_rc ? r0 : f0;
                        v[1] = __funnelshift_r(t1, t2, sh) & mk4.y;
                        v[2] = __funnelshift_r(t2, t3, sh) & mk4.z;
                        v[3] = __funnelshift_r(t3, t4, sh) & mk4.w;
                    }
                    // lo = bases 0..28 | (L - 1) << 58;  hi = bases 29..52 | row << 48
                    const uint32_t lo0 = v[0], lo1 = (v[1] & 0x03FFFFFFu) | ((L - 1u) << 26);
                    const uint32_t hi0 = __funnelshift_r(v[1], v[2], 26), hi1 = __funnelshift_r(v[2], v[3], 26);
                    const uint32_t b = __umulhi(unit_hash32(lo0, lo1, hi0, hi1), MB);
                    const uint32_t rank = atomicAdd(&s_cnt[b], 1u);
                    u0[j] = lo0; u1[j] = lo1; u2[j] = hi0; u3[j] = hi1 | ((uint32_t)s_gf[gl] << 16);
                    ubr[j] = b | (rank << 16);
                    my_windows += L;
                }
            }
            __syncthreads();
            // scan the round's histogram, reserve global space (thread t owns bucket t)
            {
                uint32_t cnt = 0;
                if (tid < MB) { cnt = s_cnt[tid]; s_cnt[tid] = 0; }
                unsigned long long gb = 0;
                if (cnt) gb = atomicAdd(&p.cursors[tid], (unsigned long long)cnt);
                if (!COUNT) {
                    uint32_t total;
                    const uint32_t off = block_excl_scan<kUsThreads>(cnt, s_warp, total);
                    if (tid == 0) s_total = total;
                    if (tid < MB) {
                        s_off[tid] = off;
                        if (cnt) {
                            if (p.cap && gb + cnt > (unsigned long long)(tid + 1) * p.cap) {
                                *p.overflow = 1ULL;
                                s_delta[tid] = p.dump;
                            } else {
                                s_delta[tid] = gb - off;
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (!COUNT) {
#pragma unroll
                for (int j = 0; j < kUsMax; ++j) {
                    if (ubr[j] != 0xFFFFFFFFu) {
                        const uint32_t b = ubr[j] & 0xFFFFu;
                        const uint32_t dst = s_off[b] + (ubr[j] >> 16);
                        s_units[dst] = make_uint4(u0[j], u1[j], u2[j], u3[j]);
                        s_b[dst] = (uint16_t)b;
                    }
                }
                __syncthreads();
                const uint32_t total = s_total;
                for (uint32_t i = tid; i < total; i += kUsThreads) p.units[s_delta[s_b[i]] + i] = s_units[i];
                __syncthreads();
            }
        }
        __syncthreads();       // the tile's staging is rewritten by the next tile
    }
    if (!COUNT) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_windows += __shfl_down_sync(0xffffffffu, my_windows, d);
        if ((tid & 31u) == 0 && my_windows) atomicAdd(p.n_windows, my_windows);
    }
}

// ---- k_units_dedupe: per bucket, distinct (unit content, genome group) -> WB presence words --------
// An entry holds the presence of WB x 64 consecutive genomes (WB = 1, 2 or 4: all of them when there are at most
// 256 genomes), so a unit shared by everybody is ONE entry.  Table: slots x (16-byte key + WB words).
constexpr int kUdThreads = 1024;
#ifndef GRMKM_UD_CHECK
#define GRMKM_UD_CHECK 16
#endif
constexpr int kUdCheck = GRMKM_UD_CHECK;                            // batches between "is the table filling up" checkpoints

__host__ __device__ inline uint32_t unit_words_per_entry(uint32_t W) { return W <= 1 ? 1u : W == 2 ? 2u : 4u; }
__host__ __device__ inline uint32_t unit_dedupe_slots(uint32_t WB) { return WB == 1 ? 8192u : WB == 2 ? 6144u : 4096u; }
__host__ __device__ inline size_t unit_dedupe_smem(uint32_t WB) { return (size_t)unit_dedupe_slots(WB) * (16 + 8 * WB); }
__host__ __device__ inline uint32_t unit_record_stride(uint32_t WB) { return WB == 1 ? 2u : 2u * WB; }   // u64 per wide record

struct UnitDedupeParams {
    const uint4* units;
    const unsigned long long* begin;     // [n_buckets]
    const unsigned long long* end;       // [n_buckets]
    uint32_t n_buckets;
    unsigned long long* out;             // entries [lo, hi (row field = genome group), WB words]
    unsigned long long cap;              // entries
    unsigned long long* needed;          // scalar: entries produced (may exceed cap; the host retries)
};

// CNT (abundance builds, -abundance-min > 1): the key is (content, genome ROW) and the word is the number of times
// the unit occurs in that genome, so that the expansion can hand every k-mer its multiplicity (WB = 1).
template <int WB, bool CNT = false>
__global__ void __launch_bounds__(kUdThreads, 1)
k_units_dedupe(const UnitDedupeParams p) {
    extern __shared__ unsigned long long s_tab[];
    __shared__ uint32_t s_distinct;
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_base;
    constexpr uint32_t S = WB == 1 ? 8192u : WB == 2 ? 6144u : 4096u;
    constexpr uint32_t kSoft = S * 6 / 10;             // flush at a checkpoint above this many entries
    constexpr uint32_t kHard = S - kUdThreads - 64;    // between checkpoints: new keys beyond this bypass the table
    static_assert(!CNT || WB == 1, "abundance entries carry one counter word");
    constexpr uint32_t GS = CNT ? 0 : WB == 1 ? 6 : WB == 2 ? 7 : 8;   // genome group = row >> GS
    constexpr uint32_t ES = 2 + WB;                    // u64 per output entry
    unsigned long long* k_lo = s_tab;
    unsigned long long* k_hi = s_tab + S;
    unsigned long long* wd = s_tab + 2 * S;            // [WB][S]
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t per = S / kUdThreads;
    for (uint32_t i = tid; i < S; i += kUdThreads) { k_lo[i] = kUnitEmptyLo; k_hi[i] = kUnitEmptyHi; }
    for (uint32_t i = tid; i < S * WB; i += kUdThreads) wd[i] = 0ULL;
    if (tid == 0) s_distinct = 0;
    __syncthreads();
    // emit the table's entries and clear it (all inserts are complete when this is called)
    auto flush = [&]() {
        uint32_t cnt = 0;
        for (uint32_t j = 0; j < per; ++j) cnt += k_lo[j * kUdThreads + tid] != kUnitEmptyLo;
        uint32_t total;
        uint32_t off = block_excl_scan<kUdThreads>(cnt, s_warp, total);
        if (tid == 0) { s_base = total ? atomicAdd(p.needed, (unsigned long long)total) : 0ULL; s_distinct = 0; }
        __syncthreads();
        const unsigned long long base = s_base;
        for (uint32_t j = 0; j < per; ++j) {
            const uint32_t i = j * kUdThreads + tid;
            const unsigned long long lo = k_lo[i];
            if (lo != kUnitEmptyLo) {
                const unsigned long long o = base + off++;
                if (o < p.cap) {
                    p.out[ES * o] = lo; p.out[ES * o + 1] = k_hi[i];
#pragma unroll
                    for (int w = 0; w < WB; ++w) p.out[ES * o + 2 + w] = wd[w * S + i];
                }
                k_lo[i] = kUnitEmptyLo; k_hi[i] = kUnitEmptyHi;
#pragma unroll
                for (int w = 0; w < WB; ++w) wd[w * S + i] = 0ULL;
            }
        }
        __syncthreads();
    };
    for (uint32_t b = blockIdx.x; b < p.n_buckets; b += gridDim.x) {
        const unsigned long long rb = p.begin[b], re = p.end[b];
        if (re <= rb) continue;
        const unsigned long long n = re - rb;
        const uint4* src = p.units + rb;
        uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
        if (tid < n) nxt = __ldcs(src + tid);
        uint32_t bi = 0;
        for (unsigned long long base = 0; base < n; base += kUdThreads, ++bi) {
            const uint4 u = nxt;
            const bool have = base + tid < n;
            if (base + kUdThreads + tid < n) nxt = __ldcs(src + base + kUdThreads + tid);
            // Checkpoint every kUdCheck batches (one barrier): flush when the table is filling up.  In between, a
            // NEW key that finds the table nearly full goes straight to the output (duplicates are merged by the
            // column aggregate), so no thread ever waits for a flush.
            if (bi % kUdCheck == 0 && __syncthreads_or(tid == 0 && s_distinct > kSoft)) flush();
            // (the warp is re-converged after the probe loop: see agg_insert in grmkm_kernels.cuh)
            const unsigned long long lo = ((unsigned long long)u.y << 32) | u.x;
            const unsigned long long hi = ((unsigned long long)u.w << 32) | u.z;
            const uint32_t row = u.w >> 16;
            const uint32_t grp = row >> GS;
            const unsigned long long hik = (hi & kUnitHiMask) | ((unsigned long long)grp << 48);
            uint32_t slot = __umulhi((unit_hash32(u.x, u.y, u.z, u.w & 0xFFFFu) + grp) * 0x297A2D39u, S);
            bool placed = true;
            if (have) {
                while (true) {
                    unsigned long long l0 = *(volatile unsigned long long*)&k_lo[slot];
                    if (l0 == kUnitEmptyLo) {
                        if (*(volatile uint32_t*)&s_distinct >= kHard) { placed = false; break; }
                        l0 = atomicCAS(&k_lo[slot], kUnitEmptyLo, lo);
                        if (l0 == kUnitEmptyLo) { atomicAdd(&s_distinct, 1u); l0 = lo; }
                    }
                    if (l0 == lo) {
                        unsigned long long h0 = *(volatile unsigned long long*)&k_hi[slot];
                        if (h0 == kUnitEmptyHi) {
                            h0 = atomicCAS(&k_hi[slot], kUnitEmptyHi, hik);
                            if (h0 == kUnitEmptyHi) h0 = hik;
                        }
                        if (h0 == hik) break;
                    }
                    if (++slot == S) slot = 0;
                }
            }
            __syncwarp();
            if (have) {
                // presence bit 63 - (row & 63) of word (row >> 6) % WB: a native 32-bit shared-memory OR on the right
                // half (a 64-bit atomicOr on shared memory compiles to a compare-and-swap loop)
                const uint32_t r6 = row & 63u, wj = (row >> 6) & (uint32_t)(WB - 1);
                if (placed && CNT) atomicAdd(reinterpret_cast<uint32_t*>(&wd[slot]), 1u);       // low half: the count
                else if (placed) atomicOr(reinterpret_cast<uint32_t*>(&wd[wj * S + slot]) + (r6 < 32u ? 1 : 0), 0x80000000u >> (r6 & 31u));
                else {
                    const unsigned long long o = atomicAdd(p.needed, 1ULL);
                    if (o < p.cap) {
                        p.out[ES * o] = lo; p.out[ES * o + 1] = hik;
#pragma unroll
                        for (int w = 0; w < WB; ++w) p.out[ES * o + 2 + w] = CNT ? 1ULL : (uint32_t)w == wj ? 1ULL << (63u - r6) : 0ULL;
                    }
                }
            }
        }
        __syncthreads();
        flush();
    }
}

// ---- k_units_expand: distinct (unit, group) -> wide records [(hash << wbits) | group, WB words (, padding)] by hash range ----
// Same tile machinery as k_scatter (32 hashed k-mers per thread in registers, counting sort in shared memory,
// bucket runs on the way out); a thread expands one weighted unit.  COUNT: histogram only (exact offsets).
struct UnitExpandParams {
    const unsigned long long* wu;        // [n][2 + WB]
    const unsigned long long* n_ptr;     // entries produced by the dedupe
    unsigned long long cap;              // entries the buffer holds
    uint32_t k, bucket_bits, wbits;
    unsigned long long* cursors;         // [B] histogram (COUNT), exact offsets, or region cursors (region_cap != 0)
    unsigned long long* records;         // [total][unit_record_stride(WB)]
    unsigned long long region_cap;       // records per over-provisioned bucket region; 0 = exact offsets from a count pass
    unsigned long long dump;             // first record of the dump area (kStTile records) behind the regions
    unsigned long long* overflow;        // set when a region was too small (the host redoes the expansion with exact offsets)
};

__device__ __forceinline__ void expand_hash_group(const uint32_t (&r)[4], const uint32_t (&y)[4], uint32_t have,
                                                  uint32_t kmask_lo, uint32_t kmask_hi, uint32_t key_bits, uint32_t* s_cnt,
                                                  unsigned long long (&hsh)[kStPerThread]) {
#pragma unroll
    for (int e = 0; e < kStPerThread; ++e) {
        if ((have >> e) & 1u) {
            const int fs = 2 * (31 - e), fw_w = fs >> 5, fw_s = fs & 31;
            const int rs = 2 * e, rc_w = rs >> 5, rc_s = rs & 31;
            const uint32_t fw_lo = __funnelshift_r(r[fw_w], r[fw_w + 1], fw_s) & kmask_lo;
            const uint32_t fw_hi = __funnelshift_r(r[fw_w + 1], r[fw_w + 2], fw_s) & kmask_hi;
            const uint32_t rc_lo = __funnelshift_r(y[rc_w], y[rc_w + 1], rc_s) & kmask_lo;
            const uint32_t rc_hi = __funnelshift_r(y[rc_w + 1], y[rc_w + 2], rc_s) & kmask_hi;
            const uint64_t fw = ((uint64_t)fw_hi << 32) | fw_lo, rc = ((uint64_t)rc_hi << 32) | rc_lo;
            const uint64_t h = khash(fw < rc ? fw : rc);
            atomicAdd(&s_cnt[(uint32_t)(h >> key_bits)], 1u);
            hsh[e] = h;
        }
    }
}

template <bool COUNT, int WB>
__global__ void __launch_bounds__(kStThreads, 1)
k_units_expand(const UnitExpandParams p) {
    constexpr uint32_t ES = 2 + WB;
    extern __shared__ unsigned long long s_dyn[];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_word[WB][kStThreads];
    __shared__ uint16_t s_blk[kStThreads];
    const uint32_t B = 1u << p.bucket_bits;
    unsigned long long* s_delta = s_dyn;
    unsigned long long* s_rec = s_delta + B;
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_rec + kStTile);
    uint32_t* s_off = s_cnt + B;
    uint16_t* s_src = reinterpret_cast<uint16_t*>(s_off + B);
    const unsigned long long n_all = *p.n_ptr;
    const unsigned long long n = n_all < p.cap ? n_all : p.cap;
    const unsigned long long n_tiles = (n + kStThreads - 1) / kStThreads;
    const uint32_t k = p.k;
    const uint32_t kmask_lo = k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    const uint32_t kmask_hi = k <= 16 ? 0u : (k == 32 ? 0xFFFFFFFFu : ((1u << (2 * k - 32)) - 1u));
    const uint32_t key_bits = 64 - p.bucket_bits;
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < B; i += kStThreads) s_cnt[i] = 0;
    __syncthreads();
    for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const unsigned long long i = tile * kStThreads + tid;
        unsigned long long hsh[kStPerThread];
        uint32_t have = 0;
        if (i < n) {
            const unsigned long long lo = p.wu[ES * i], hik = p.wu[ES * i + 1];
            if (!COUNT) {
#pragma unroll
                for (int w = 0; w < WB; ++w) s_word[w][tid] = p.wu[ES * i + 2 + w];
                s_blk[tid] = (uint16_t)(hik >> 48);
            }
            const uint32_t L = (uint32_t)(lo >> 58) + 1u;
            const unsigned long long Vlo = (lo & kUnitLoMask) | (hik << 58), Vhi = (hik & kUnitHiMask) >> 6;
            // window: the unit's entry t at window entry 33 - k + t, so that its first k-mer ends at window entry 32
            const uint32_t sh = 2u * (33u - k);
            unsigned long long Xlo, Xhi;
            if (sh == 64) { Xlo = 0ULL; Xhi = Vlo; }
            else { Xlo = Vlo << sh; Xhi = (Vhi << sh) | (Vlo >> (64 - sh)); }
            const uint32_t x[4] = {(uint32_t)Xlo, (uint32_t)(Xlo >> 32), (uint32_t)Xhi, (uint32_t)(Xhi >> 32)};
            const uint32_t y[4] = {(uint32_t)Vlo ^ 0xAAAAAAAAu, (uint32_t)(Vlo >> 32) ^ 0xAAAAAAAAu,
                                   (uint32_t)Vhi ^ 0xAAAAAAAAu, (uint32_t)(Vhi >> 32) ^ 0xAAAAAAAAu};
            const uint32_t r[4] = {rev2_32(x[3]), rev2_32(x[2]), rev2_32(x[1]), rev2_32(x[0])};
            have = L >= 32 ? 0xFFFFFFFFu : ((1u << L) - 1u);
            expand_hash_group(r, y, have, kmask_lo, kmask_hi, key_bits, s_cnt, hsh);
        }
        __syncthreads();
        {
            uint32_t cnt[kStMaxBins];
            uint32_t sum = 0;
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)kStMaxBins; ++j) {
                const uint32_t bin = j * kStThreads + tid;
                cnt[j] = 0;
                if (bin < B) { cnt[j] = s_cnt[bin]; s_cnt[bin] = 0; sum += cnt[j]; }
            }
            unsigned long long gb[kStMaxBins];
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)kStMaxBins; ++j) {
                const uint32_t bin = j * kStThreads + tid;
                gb[j] = cnt[j] ? atomicAdd(&p.cursors[bin], (unsigned long long)cnt[j]) : 0ULL;
            }
            if (!COUNT) {
                uint32_t total;
                uint32_t off = block_excl_scan<kStThreads>(sum, s_warp, total);
                if (tid == 0) s_total = total;
#pragma unroll
                for (uint32_t j = 0; j < (uint32_t)kStMaxBins; ++j) {
                    const uint32_t bin = j * kStThreads + tid;
                    if (bin < B) {
                        s_off[bin] = off;
                        if (cnt[j]) {
                            if (p.region_cap && gb[j] + cnt[j] > (unsigned long long)(bin + 1) * p.region_cap) {
                                *p.overflow = 1ULL;
                                s_delta[bin] = p.dump;          // the tile's records of this bucket land in the dump area
                            } else {
                                s_delta[bin] = gb[j] - off;
                            }
                        }
                        off += cnt[j];
                    }
                }
            }
        }
        __syncthreads();
        if (!COUNT) {
#pragma unroll
            for (int e = 0; e < kStPerThread; ++e) {
                if ((have >> e) & 1u) {
                    const uint32_t dst = atomicAdd(&s_off[(uint32_t)(hsh[e] >> key_bits)], 1u);
                    s_rec[dst] = hsh[e];
                    s_src[dst] = (uint16_t)tid;
                }
            }
            __syncthreads();
            const uint32_t total = s_total;
            ulonglong2* out = reinterpret_cast<ulonglong2*>(p.records);
            for (uint32_t j = tid; j < total; j += kStThreads) {
                const unsigned long long h = s_rec[j];
                const uint32_t src = s_src[j];
                const unsigned long long o = s_delta[(uint32_t)(h >> key_bits)] + j;
                const unsigned long long key = (h << p.wbits) | s_blk[src];
                if (WB == 1) out[o] = make_ulonglong2(key, s_word[0][src]);
                else {
#pragma unroll
                    for (int q = 0; q < WB; ++q)        // record = WB x 16 bytes: key, WB words, padding
                        out[o * WB + q] = make_ulonglong2(q == 0 ? key : (2 * q <= WB ? s_word[(2 * q - 1) % WB][src] : 0ULL),
                                                          2 * q < WB ? s_word[(2 * q) % WB][src] : 0ULL);
                }
            }
            __syncthreads();      // s_word / s_blk / s_rec are rewritten by the next tile
        }
    }
}

// region ends for the over-provisioned unit regions + total units (one block)
__global__ void __launch_bounds__(1024)
k_finish_unit_regions(const unsigned long long* __restrict__ begin, unsigned long long* __restrict__ end, uint32_t MB,
                      unsigned long long cap, unsigned long long* __restrict__ scalars, int which) {
    unsigned long long v = 0;
    for (uint32_t b = threadIdx.x; b < MB; b += blockDim.x) {
        unsigned long long e = end[b];
        if (cap && e > (unsigned long long)(b + 1) * cap) { e = (unsigned long long)(b + 1) * cap; end[b] = e; }
        v += e - begin[b];
    }
    const unsigned long long sum = block_sum_u64(v);
    if (threadIdx.x == 0) scalars[which] = sum;
}

}  // namespace grmkm
