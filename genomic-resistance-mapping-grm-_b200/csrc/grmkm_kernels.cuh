// grmkm_kernels.cuh -- the sm_100a kernels of the k-mer matrix path.
//
//   parse   : k_first_header, k_tile_tickets, k_pack (one pass: tile summaries resolved by decoupled look-back)
//             FASTA/FASTQ text -> dense 2-bit base stream + validity mask   (multidsk's bank reader)
//   extract : k_extract<COUNT|SCATTER>
//             canonical k-mers -> hash buckets                              (multidsk's partitioning)
//   count   : k_abundance        per-(k-mer, genome) abundance filter       (multidsk -abundance-min)
//   merge   : k_aggregate        per-bucket shared-memory hash aggregation: presence bits of all
//             genomes ORed into 64-genome words + singleton filter         (dsk2kover)
//   order   : k_sort_*           LSD radix sort of the columns by canonical k-mer, k_gather
//   emit    : k_kmer_strings, k_format_tsv                                  (kmer_sequences / Ray TSV)
#pragma once
#include "grmkm_device.cuh"

namespace grmkm {

// ------------------------------------------------------------------------------------------
// parse
// ------------------------------------------------------------------------------------------

// One warp per file: offset of the first header marker ('>' / '@') at a line start, or len.
__global__ void k_first_header(const FileDesc* __restrict__ files, uint32_t n_files, uint64_t* __restrict__ hdr0) {
    const uint32_t f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= n_files) return;
    const FileDesc fd = files[f];
    const uint32_t marker = fd.kind == 0 ? '>' : '@';
    uint64_t found = fd.len;
    for (uint64_t base = 0; base < fd.len; base += 32) {
        const uint64_t pos = base + lane;
        bool hit = false;
        if (pos < fd.len) {
            const uint32_t c = fd.ptr[pos];
            const bool ls = pos == 0 || fd.ptr[pos - 1] == '\n';
            hit = ls && c == marker;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m) { found = base + (__ffs(m) - 1); break; }
    }
    if (lane == 0) hdr0[f] = found;
}

struct TileCtx {
    FileDesc fd;
    uint64_t hdr0;
    uint64_t off;   // byte offset of this thread's first chunk in the file (64 contiguous bytes per thread)
    uint32_t f;
    bool first_tile;
};

// the thread's four chunks and the byte before each of them
struct ThreadText {
    Chunk16 ch[kChunksPerThread];
    uint32_t prev[kChunksPerThread];
};
__device__ __forceinline__ ThreadText load_thread_text(const TileCtx& t) {
    ThreadText x;
#pragma unroll
    for (int c = 0; c < kChunksPerThread; ++c) x.ch[c] = load_chunk(t.fd.ptr, t.off + 16 * c, t.fd.len);
    uint32_t p = __shfl_up_sync(0xffffffffu, x.ch[kChunksPerThread - 1].byte(15), 1);
    if ((threadIdx.x & 31) == 0) p = (t.off > 0 && t.off - 1 < t.fd.len) ? t.fd.ptr[t.off - 1] : (uint32_t)'\n';
    x.prev[0] = p;
#pragma unroll
    for (int c = 1; c < kChunksPerThread; ++c) x.prev[c] = x.ch[c - 1].byte(15);
    return x;
}

// Everything a CTA needs to start on ticket t, in one 64-byte record: ticket -> tile (order) -> file -> descriptor,
// first header and stream start used to be four dependent loads at the head of every CTA.
struct TileTicket {
    FileDesc fd;
    uint64_t hdr0;
    uint64_t stream_start;      // where the file's entries start in the packed stream
    uint32_t tile, f;
    uint64_t pad;
};
static_assert(sizeof(TileTicket) == 64, "one ticket = four 16-byte loads");
__global__ void k_tile_tickets(const FileDesc* __restrict__ files, uint32_t n_files, const uint64_t* __restrict__ hdr0,
                               const uint64_t* __restrict__ fss, const uint32_t* __restrict__ order, uint64_t n_tiles,
                               TileTicket* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    TileTicket k;
    k.tile = order[t];
    uint32_t lo = 0, hi = n_files - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (files[mid].tile_begin <= k.tile) lo = mid; else hi = mid - 1;
    }
    k.f = lo; k.fd = files[lo]; k.hdr0 = hdr0[lo]; k.stream_start = fss[lo]; k.pad = 0;
    out[t] = k;
}

// ---- single-pass parse: decoupled look-back over the tile summaries --------------------------------------
// The parser is a transducer scan (tile summary = state -> (end state, entries emitted), composition associative).
// Instead of summary kernel + scan kernels + pack kernel (the text read twice), k_pack publishes its tile's
// summary as soon as the block scan has produced it, and one warp walks back over the earlier tiles' published
// summaries until it meets a tile whose incoming (state, position) is already resolved.  Tiles take their index
// from a ticket counter, so every tile a CTA waits for has started and publishes without waiting for anybody.
// Every published word validates itself (bit 63; the arrays are zeroed before the launch), so a reader needs one
// round trip per window of 64 tiles and no fences:
//   a0[t] = 1<<63 | e << 48 | c0 << 24 | c1      summary: end states, entries emitted from state 0 / 1
//   a1[t] = 1<<63 | c2 << 24 | c3                (FASTQ only: states 2 and 3)
//   ps[t] = 1<<63 | state << 61 | position       state and stream position at the tile's first byte
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr unsigned long long kPubValid = 1ULL << 63;
__device__ __forceinline__ unsigned long long pub_a0(const Sum& s) {
    return kPubValid | ((unsigned long long)(s.e & 0xFFu) << 48) | ((unsigned long long)(s.c0 & 0xFFFFFFu) << 24) | (s.c1 & 0xFFFFFFu);
}
__device__ __forceinline__ unsigned long long pub_a1(const Sum& s) {
    return kPubValid | ((unsigned long long)(s.c2 & 0xFFFFFFu) << 24) | (s.c3 & 0xFFFFFFu);
}
__device__ __forceinline__ Sum pub_sum(unsigned long long a0, unsigned long long a1) {
    Sum r;
    r.e = (uint32_t)(a0 >> 48) & 0xFFu; r.c0 = (uint32_t)(a0 >> 24) & 0xFFFFFFu; r.c1 = (uint32_t)a0 & 0xFFFFFFu;
    r.c2 = (uint32_t)(a1 >> 24) & 0xFFFFFFu; r.c3 = (uint32_t)a1 & 0xFFFFFFu;
    return r;
}
static_assert(kTileBytes < (1 << 24), "a tile's entry counts are published in 24 bits");
__device__ __forceinline__ Sum sum_shfl_down(const Sum& a, int d) {
    Sum r;
    r.e = __shfl_down_sync(0xffffffffu, a.e, d);
    r.c0 = __shfl_down_sync(0xffffffffu, a.c0, d);
    r.c1 = __shfl_down_sync(0xffffffffu, a.c1, d);
    r.c2 = __shfl_down_sync(0xffffffffu, a.c2, d);
    r.c3 = __shfl_down_sync(0xffffffffu, a.c3, d);
    return r;
}
// ordered fold of one Sum per lane (lane 0 first); the result is valid in every lane
__device__ __forceinline__ Sum warp_fold_sum(Sum s) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Sum o = sum_shfl_down(s, d);
        if (lane + d < 32) s = sum_combine(s, o);              // lane i: fold of lanes [i, i + 2d)
    }
    Sum w;
    w.e = __shfl_sync(0xffffffffu, s.e, 0); w.c0 = __shfl_sync(0xffffffffu, s.c0, 0); w.c1 = __shfl_sync(0xffffffffu, s.c1, 0);
    w.c2 = __shfl_sync(0xffffffffu, s.c2, 0); w.c3 = __shfl_sync(0xffffffffu, s.c3, 0);
    return w;
}
// one full warp; state and stream position at the first byte of `tile`.  Every file starts in state 0 at a stream
// position the host fixes in advance (file_stream_start: the file's length rounded up to whole groups bounds its
// entries, the gap reads as invalid entries, and no window spans two files anyway), so the chain of a tile ends at
// its file's first tile: tiles before it count as published (identity) and resolved (state 0, that position).
// Tickets are dealt round-robin over the files (PackParams::order), so the tile a CTA depends on was started one
// whole round earlier and is normally resolved by the time it is asked.
// The counts folded here cover only tiles that are in flight, far below 2^32 entries.
// kLbWindows x 32 tiles are polled per round trip (lane 31 of window 0 = the nearest tile): the resolved front
// trails the newest published tile by (tiles per microsecond) x (poll latency), and a poll that does not reach it
// costs a whole extra round trip.
constexpr int kLbWindows = 2;
template <int KIND>
__device__ __noinline__ void tile_lookback(uint64_t tile, uint64_t first, uint64_t pos_first, const unsigned long long* a0,
                                           const unsigned long long* a1, const unsigned long long* ps, uint32_t& st, uint64_t& pos) {
    const int lane = threadIdx.x & 31;
    const unsigned long long ident = kPubValid | (0xE4ULL << 48);      // the identity summary, published
    Sum acc = sum_identity();                      // fold of the tiles between the resolved one and `tile`
    long long hi = (long long)tile - 1;            // nearest tile not folded yet
    while (true) {
        unsigned long long w0[kLbWindows], w1[kLbWindows], wp[kLbWindows];
#pragma unroll
        for (int w = 0; w < kLbWindows; ++w) {
            const long long j = hi - 32 * w - 31 + lane;
            w0[w] = ident; w1[w] = kPubValid; wp[w] = kPubValid | pos_first;
            if (j >= (long long)first) { w0[w] = ld_relaxed_u64(a0 + j); wp[w] = ld_relaxed_u64(ps + j); if (KIND == 1) w1[w] = ld_relaxed_u64(a1 + j); }
        }
        bool done = false;
        int folded = 0;                            // windows of this poll that are folded into acc
#pragma unroll
        for (int w = 0; w < kLbWindows; ++w) {
            const bool rdy_l = (w0[w] & w1[w] & kPubValid) != 0;
            const uint32_t rdy = __ballot_sync(0xffffffffu, rdy_l);
            const uint32_t res = __ballot_sync(0xffffffffu, rdy_l && (wp[w] & kPubValid));
            if (w == 0 && (res >> 31)) {
                // the usual case: the tile right before this one is resolved (no fold; a fold costs ~600 instructions)
                const unsigned long long q0 = __shfl_sync(0xffffffffu, w0[0], 31), q1 = __shfl_sync(0xffffffffu, w1[0], 31);
                const unsigned long long r = __shfl_sync(0xffffffffu, wp[0], 31);
                const Sum a = sum_combine(pub_sum(q0, q1), acc);
                const uint32_t st0 = (uint32_t)(r >> 61) & 3u;
                st = sum_end(a, st0);
                pos = (r & ((1ULL << 61) - 1)) + sum_cnt(a, st0);
                return;
            }
            const int top = res ? 31 - __clz(res) : 0;             // nearest resolved tile of the window, if any
            if ((rdy >> top) != (0xFFFFFFFFu >> top)) break;       // a tile this side of it has not published yet: poll again
            acc = sum_combine(warp_fold_sum(lane < top ? sum_identity() : pub_sum(w0[w], w1[w])), acc);
            folded = w + 1;
            if (res) {
                const unsigned long long r = __shfl_sync(0xffffffffu, wp[w], top);
                const uint32_t st0 = (uint32_t)(r >> 61) & 3u;
                st = sum_end(acc, st0);
                pos = (r & ((1ULL << 61) - 1)) + sum_cnt(acc, st0);
                done = true;
                break;
            }
        }
        if (done) return;
        hi -= 32 * folded;
    }
}

// text tile -> packed stream: codes64[g] holds entries 32g..32g+31 (entry j at bits 2j), valid32[g] bit j.
// Every thread turns its 64 bytes into at most 64 entries (register accumulator), the block scan of the
// thread summaries gives each thread its entry offset inside the tile, the look-back gives the tile its state and
// position, and the tile is assembled in shared memory.
struct PackParams {
    const FileDesc* files;
    const uint64_t* hdr0;
    uint64_t n_tiles;
    uint32_t n_files;
    const TileTicket* tickets;          // [n_tiles] ticket -> tile and its file, tickets dealt round-robin over the files
    uint64_t stream_len;                // padded stream length of the batch (-> scalars[S_STREAM_LEN])
    uint32_t* ticket;                   // zeroed before the launch
    unsigned long long* pub_a0;         // [n_tiles] x 3, zeroed before the launch (see tile_lookback)
    unsigned long long* pub_a1;
    unsigned long long* pub_ps;
    unsigned long long* codes;
    uint32_t* valid;
    uint64_t* scalars;
};

template <int KIND>
__global__ void __launch_bounds__(kParseThreads, 1024 / kParseThreads)
k_pack(const PackParams p) {
    constexpr int kGroups = kTileBytes / 32 + 2;
    __shared__ Sum s_w[kParseThreads / 32];
    __shared__ uint32_t s_codes[kGroups * 2 + 4];
    __shared__ uint32_t s_valid[kGroups + 2];
    __shared__ uint32_t s_nrec, s_st;
    __shared__ uint64_t s_pos;
    __shared__ uint4 s_tk[4];                                           // this CTA's TileTicket
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(p.ticket, 1u);
        s_nrec = 0;
        if (ticket == 0) p.scalars[S_STREAM_LEN] = p.stream_len;
        if (ticket < p.n_tiles) {
            const uint4* src = reinterpret_cast<const uint4*>(p.tickets + ticket);
            const uint4 a = src[0], b = src[1], c = src[2], d = src[3];
            s_tk[0] = a; s_tk[1] = b; s_tk[2] = c; s_tk[3] = d;
        } else {
            s_tk[3] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0u);                 // no tile left
        }
    }
    for (int i = threadIdx.x; i < kGroups * 2 + 4; i += blockDim.x) s_codes[i] = 0;
    for (int i = threadIdx.x; i < kGroups + 2; i += blockDim.x) s_valid[i] = 0;
    __syncthreads();
    const TileTicket& tk = *reinterpret_cast<const TileTicket*>(s_tk);
    const uint64_t tile = tk.tile;
    if (tk.tile == 0xFFFFFFFFu) return;
    unsigned long long* __restrict__ codes = p.codes;
    uint32_t* __restrict__ valid = p.valid;
    TileCtx t;
    t.f = tk.f; t.fd = tk.fd; t.hdr0 = tk.hdr0;
    t.first_tile = (tile == t.fd.tile_begin);
    t.off = (tile - t.fd.tile_begin) * (uint64_t)kTileBytes + (uint64_t)threadIdx.x * (16 * kChunksPerThread);
    const uint64_t file_stream_start = tk.stream_start;
    const ThreadText x = load_thread_text(t);
    // publish the tile's summary, resolve its incoming state and position (warp 0), publish those
    auto publish = [&](const Sum& total) {
        if (threadIdx.x == 0) {
            if (KIND == 1) st_relaxed_u64(p.pub_a1 + tile, pub_a1(total));
            st_relaxed_u64(p.pub_a0 + tile, pub_a0(total));
        }
    };
    auto resolve = [&](uint32_t& st_in, uint64_t& tpos) {
        if (threadIdx.x < 32) {
            uint32_t st; uint64_t pos;
            tile_lookback<KIND>(tile, t.fd.tile_begin, file_stream_start, p.pub_a0, p.pub_a1, p.pub_ps, st, pos);
            if (threadIdx.x == 0) {
                st_relaxed_u64(p.pub_ps + tile, kPubValid | ((unsigned long long)st << 61) | pos);
                s_st = st; s_pos = pos;
            }
        }
        __syncthreads();
        st_in = s_st; tpos = s_pos;
    };
    uint32_t st_in; uint64_t tpos;
    Acc64 acc; acc.init();
    uint32_t nrec = 0, local, e_total;
    if (KIND == 0) {
        FaParts part[kChunksPerThread];
        uint32_t mine = 0;
#pragma unroll
        for (int c = 0; c < kChunksPerThread; ++c) {
            const FaLines ln = fa_lines(x.ch[c], t.off + 16 * c, t.fd.len, t.hdr0);
            part[c] = fa_chunk_parts(x.ch[c], ln, x.prev[c], t.off + 16 * c, t.fd.len, t.hdr0);
            mine = fa_combine(mine, part[c].sum());
        }
        uint32_t ex, tot;
        block_scan_fa(mine, ex, tot, reinterpret_cast<uint32_t*>(s_w));
        publish(fa_to_sum(tot));
        // The entries are accumulated while the look-back is still in flight, as if the tile started inside a sequence
        // line; the few threads in front of the tile's first line start redo it when it started inside a header.
        const uint32_t ex_t = ex >> kFaT;
        auto build = [&](bool in_seq) {                              // in_seq: line type at this thread's first byte
            acc.init(); nrec = 0;
#pragma unroll
            for (int c = 0; c < kChunksPerThread; ++c) {
                // head (only inside a sequence line) and rest of the chunk as ONE run of at most 16 entries
                const uint32_t hn = in_seq ? part[c].hn() : 0u;
                const uint32_t cc = (in_seq ? part[c].hc : 0u) | (hn < 16u ? part[c].rc << (2 * hn) : 0u);      // hn = 16: no rest
                const uint32_t vv = (in_seq ? part[c].hv_rv & 0xFFFFu : 0u) | ((part[c].hv_rv >> 16) << hn);
                acc.append(cc, vv, hn + part[c].rn());
                nrec += part[c].nrec();
                if (part[c].t()) in_seq = (part[c].t() == 2);
            }
        };
        build(ex_t ? (ex_t == 2) : true);
        resolve(st_in, tpos);
        const bool tile_seq = (st_in == ST_SEQ);
        if (!tile_seq && ex_t == 0) build(false);
        local = (ex & kFaMask) + (tile_seq ? ((ex >> kFaH) & kFaMask) : 0u);
        e_total = (tot & kFaMask) + (tile_seq ? ((tot >> kFaH) & kFaMask) : 0u);
    } else {
        Sum sums[kChunksPerThread];
        Sum mine = sum_identity(), excl, total;
#pragma unroll
        for (int c = 0; c < kChunksPerThread; ++c) {
            sums[c] = chunk_summary<1>(x.ch[c], x.prev[c], t.off + 16 * c, t.fd.len, t.hdr0);
            mine = sum_combine(mine, sums[c]);
        }
        block_scan_sum(mine, excl, total, s_w);
        publish(total);
        resolve(st_in, tpos);
        uint32_t st = sum_end(excl, st_in);
        local = sum_cnt(excl, st_in);
        e_total = sum_cnt(total, st_in);
#pragma unroll
        for (int c = 0; c < kChunksPerThread; ++c) {
            uint32_t cbits = 0, vbits = 0, n = 0, prev = x.prev[c];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint64_t pos = t.off + 16 * c + i;
                const uint32_t ch = x.ch[c].byte(i);
                if (pos < t.fd.len && pos >= t.hdr0) {
                    const bool ls = (pos == t.hdr0) || (prev == '\n');
                    if (ch == '\n') st = (st + 1) & 3;
                    else if (st == 0) { if (ls) { n++; nrec++; } }
                    else if (st == 1 && ch != '\r') {
                        if (is_acgt(ch)) { cbits |= ((ch >> 1) & 3u) << (2 * n); vbits |= 1u << n; }
                        n++;
                    }
                }
                prev = ch;
            }
            acc.append(cbits, vbits, n);
        }
    }
    if (acc.n) {
        const uint32_t rel = (uint32_t)(tpos & 31) + local;         // entry offset from the tile's first group
        const uint32_t cb = rel * 2, cw = cb >> 5, cs = cb & 31;
        const uint32_t c0 = (uint32_t)acc.clo, c1 = (uint32_t)(acc.clo >> 32), c2 = (uint32_t)acc.chi, c3 = (uint32_t)(acc.chi >> 32);
        const uint32_t w0 = c0 << cs, w1 = __funnelshift_l(c0, c1, cs), w2 = __funnelshift_l(c1, c2, cs),
                       w3 = __funnelshift_l(c2, c3, cs), w4 = cs ? (c3 >> (32 - cs)) : 0u;
        if (w0) atomicOr(&s_codes[cw], w0);
        if (w1) atomicOr(&s_codes[cw + 1], w1);
        if (w2) atomicOr(&s_codes[cw + 2], w2);
        if (w3) atomicOr(&s_codes[cw + 3], w3);
        if (w4) atomicOr(&s_codes[cw + 4], w4);
        const uint32_t vw = rel >> 5, vs = rel & 31;
        const uint32_t v0 = (uint32_t)acc.v, v1 = (uint32_t)(acc.v >> 32);
        const uint32_t u0 = v0 << vs, u1 = __funnelshift_l(v0, v1, vs), u2 = vs ? (v1 >> (32 - vs)) : 0u;
        if (u0) atomicOr(&s_valid[vw], u0);
        if (u1) atomicOr(&s_valid[vw + 1], u1);
        if (u2) atomicOr(&s_valid[vw + 2], u2);
    }
    if (nrec) atomicAdd(&s_nrec, nrec);
    __syncthreads();
    if (e_total) {
        const uint64_t g0 = tpos >> 5;
        const uint32_t first_off = (uint32_t)(tpos & 31);
        const uint32_t ngroups = (first_off + e_total + 31) >> 5;
        for (uint32_t gi = threadIdx.x; gi < ngroups; gi += blockDim.x) {
            const unsigned long long cw = (unsigned long long)s_codes[2 * gi] | ((unsigned long long)s_codes[2 * gi + 1] << 32);
            const uint32_t vw = s_valid[gi];
            const bool partial = (gi == 0 && first_off) || (gi == ngroups - 1 && ((first_off + e_total) & 31));
            if (partial) {
                if (cw) atomicOr(&codes[g0 + gi], cw);
                if (vw) atomicOr(&valid[g0 + gi], vw);
            } else {
                codes[g0 + gi] = cw;
                valid[g0 + gi] = vw;
            }
        }
    }
    if (threadIdx.x == 0) {
        if (s_nrec) atomicAdd((unsigned long long*)&p.scalars[S_N_RECORDS], (unsigned long long)s_nrec);
        if (e_total) atomicAdd((unsigned long long*)&p.scalars[S_STREAM_TOTAL], (unsigned long long)e_total);
    }
}

// ------------------------------------------------------------------------------------------
// extract: canonical k-mers -> hash buckets
// ------------------------------------------------------------------------------------------
struct ExtractParams {
    const unsigned long long* codes;
    const uint32_t* valid;
    const uint64_t* scalars;            // S_STREAM_LEN
    const uint64_t* file_stream_start;  // [n_files + 1]
    const FileDesc* files;
    uint32_t n_files;
    uint32_t k;
    uint32_t bucket_bits;
    uint32_t row_bits;
    unsigned long long* hist;     // [B] (COUNT) / cursors [B] (SCATTER)
    unsigned long long* records;  // SCATTER
    uint32_t dbg;                 // timing experiments only: 1 = no atomics, 2 = no stores
    const unsigned long long* offsets;
};

template <int MODE>  // 0 = count per bucket, 1 = scatter records
__global__ void __launch_bounds__(kExtractThreads)
k_extract(const ExtractParams p) {
    extern __shared__ uint32_t s_hist[];   // COUNT: B counters
    __shared__ uint32_t s_f0;
    const uint32_t B = 1u << p.bucket_bits;
    const uint64_t stream_len = p.scalars[S_STREAM_LEN];
    const uint64_t n_groups = (stream_len + 31) >> 5;
    const uint64_t n_etiles = (n_groups + kExtractThreads - 1) / kExtractThreads;
    const uint32_t k = p.k;
    const uint64_t kmask = k == 32 ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const uint32_t key_bits = 64 - p.bucket_bits;
    const uint64_t key_mask = (1ULL << key_bits) - 1;
    const int lane = threadIdx.x & 31;
    if (MODE == 0) {
        for (uint32_t i = threadIdx.x; i < B; i += blockDim.x) s_hist[i] = 0;
        __syncthreads();
    }
    for (uint64_t et = blockIdx.x; et < n_etiles; et += gridDim.x) {
        const uint64_t g = et * kExtractThreads + threadIdx.x;
        // file of the tile's first position
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint64_t p0 = et * kExtractThreads * 32ULL;
            uint32_t lo = 0, hi = p.n_files - 1;
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (p.file_stream_start[mid] <= p0) lo = mid; else hi = mid - 1;
            }
            s_f0 = lo;
        }
        __syncthreads();
        unsigned long long cur_c = 0; uint32_t cur_v = 0;
        if (g < n_groups) { cur_c = p.codes[g]; cur_v = p.valid[g]; }
        unsigned long long prev_c = __shfl_up_sync(0xffffffffu, cur_c, 1);
        uint32_t prev_v = __shfl_up_sync(0xffffffffu, cur_v, 1);
        if (lane == 0) {
            if (g > 0 && g - 1 < n_groups) { prev_c = p.codes[g - 1]; prev_v = p.valid[g - 1]; }
            else { prev_c = 0; prev_v = 0; }
        }
        if (g >= n_groups || cur_v == 0) continue;
        const uint64_t pos0 = g * 32ULL;
        uint32_t f = s_f0;
        while (f + 1 < p.n_files && p.file_stream_start[f + 1] <= pos0) ++f;
        uint64_t next_start = (f + 1 < p.n_files) ? p.file_stream_start[f + 1] : ~0ULL;
        uint32_t row = p.files[f].row;
        // state after the last entry of the previous group
        const uint64_t le = k == 32 ? prev_c : ((prev_c >> (2 * (32 - k))) & kmask);
        uint64_t rc = le ^ (0xAAAAAAAAAAAAAAAAULL & kmask);
        uint64_t fw = rev2(le) >> (64 - 2 * k);
        uint32_t run = min((uint32_t)__clz(~prev_v), k);
#pragma unroll 4
        for (int e = 0; e < 32; ++e) {
            const uint32_t c = (uint32_t)(cur_c >> (2 * e)) & 3u;
            fw = ((fw << 2) | c) & kmask;
            rc = (rc >> 2) | ((uint64_t)(c ^ 2u) << (2 * (k - 1)));
            if ((cur_v >> e) & 1u) run = min(run + 1, k); else run = 0;
            if (run == k) {
                const uint64_t pos = pos0 + e;
                if (pos >= next_start) {
                    while (f + 1 < p.n_files && p.file_stream_start[f + 1] <= pos) ++f;
                    next_start = (f + 1 < p.n_files) ? p.file_stream_start[f + 1] : ~0ULL;
                    row = p.files[f].row;
                }
                const uint64_t canon = fw < rc ? fw : rc;
                const uint64_t h = khash(canon);
                const uint32_t b = (uint32_t)(h >> key_bits);
                if (MODE == 0) {
                    atomicAdd(&s_hist[b], 1u);
                } else {
                    unsigned long long slot;
                    if (p.dbg & 1) {
                        const unsigned long long o0 = p.offsets[b], o1 = p.offsets[b + 1];
                        slot = o0 + (o1 > o0 ? (h & key_mask) % (o1 - o0) : 0);
                    } else slot = atomicAdd(&p.hist[(size_t)b * kCursorStride], 1ULL);
                    if (!(p.dbg & 2)) p.records[slot] = ((h & key_mask) << p.row_bits) | row;
                }
            }
        }
    }
    if (MODE == 0) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < B; i += blockDim.x) {
            const uint32_t v = s_hist[i];
            if (v) atomicAdd(&p.hist[(size_t)i * kCursorStride], (unsigned long long)v);
        }
    }
}

// ---- staged scatter: one tile of 16384 stream positions per CTA iteration -------------------------
// Phase 1: every thread turns one 32-entry stream group into up to 32 hashed k-mers held in registers.
//   Forward and reverse-complement values are funnel-shift extractions from a 128-bit window of the
//   packed stream (and of its 2-bit-reversed copy), so there is no serial rolling dependency.
//   The tile histogram is built with non-returning shared-memory atomics.
// Phase 2: scan of the tile histogram; ONE global atomic per (tile, non-empty bucket) reserves space.
// Phase 3: counting sort of the tile into shared memory (slot = atomic walk of the bucket's run).
// Phase 4: copy out; records of one bucket leave the SM as contiguous runs.
// Bucket regions are either exact (cursors = prefix sums from a count pass, cap == 0) or
// over-provisioned (region b = [b*cap, (b+1)*cap)); an overflowing bucket raises S_OVERFLOW and its
// records are diverted to a dump area so nothing is corrupted; the host then re-runs the exact path.
constexpr int kStThreads = 512;
constexpr int kStPerThread = 32;                       // one 32-entry stream group per thread
constexpr int kStTile = kStThreads * kStPerThread;     // 16384 positions = 512 groups
constexpr int kStMaxBuckets = 4096;
constexpr int kStMaxBins = kStMaxBuckets / kStThreads; // bins per thread in phase 2

__host__ __device__ inline size_t staged_smem_bytes(uint32_t B) {
    return (size_t)B * (8 + 4 + 4) + (size_t)kStTile * (8 + 2);
}

struct ScatterParams {
    const unsigned long long* codes;
    const uint32_t* valid;
    const uint64_t* scalars;            // S_STREAM_LEN
    const uint64_t* file_stream_start;  // [n_files + 1]
    const FileDesc* files;
    const uint32_t* tile_file;          // file of the first position of every scatter tile
    uint32_t n_files;
    uint32_t k;
    uint32_t bucket_bits;
    uint32_t row_bits;
    unsigned long long* cursors;        // [B]
    unsigned long long* records;
    unsigned long long cap;             // records per bucket region, 0 = exact offsets
    unsigned long long dump;            // index of the dump area (kStTile records)
    unsigned long long* overflow;       // scalar raised when a region is too small
};

// file of the first position of every scatter tile (one thread per tile)
__global__ void k_scatter_tile_files(const uint64_t* __restrict__ scalars, const uint64_t* __restrict__ fss,
                                     uint32_t n_files, uint32_t* __restrict__ tile_file, uint64_t max_tiles) {
    const uint64_t tile = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tile >= max_tiles) return;
    const uint64_t p0 = tile * (uint64_t)kStTile;
    if (p0 >= scalars[S_STREAM_LEN]) return;
    uint32_t lo = 0, hi = n_files - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (fss[mid] <= p0) lo = mid; else hi = mid - 1;
    }
    tile_file[tile] = lo;
}

__device__ __noinline__ uint32_t row_of_position(const uint64_t* __restrict__ fss, const FileDesc* __restrict__ files,
                                                 uint32_t n_files, uint32_t f, uint64_t pos) {
    while (f + 1 < n_files && fss[f + 1] <= pos) ++f;
    return files[f].row;
}

__device__ __forceinline__ uint32_t rev2_32(uint32_t x) {
    x = __brev(x);
    return ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
}

// exclusive scan of one u32 per thread over an NT-thread block; s_warp must hold 33 words
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* s_warp, uint32_t& total) {
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
        const uint32_t w = lane < NT / 32 ? s_warp[lane] : 0u;
        uint32_t wi = w;
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

// phase 1 body: ALL = every window of this group is valid (no per-entry test)
template <bool ALL>
__device__ __forceinline__ uint32_t scatter_hash_group(const uint32_t (&r)[4], const uint32_t (&y)[4], uint32_t n0, uint32_t n1,
                                                       uint32_t kbits, uint32_t kmask_lo, uint32_t kmask_hi, uint32_t key_bits,
                                                       uint32_t* s_cnt, unsigned long long (&hsh)[kStPerThread]) {
    uint32_t have = 0;
#pragma unroll
    for (int e = 0; e < kStPerThread; ++e) {
        if (ALL || (__funnelshift_r(n0, n1, e) & kbits) == 0) {
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
            if (!ALL) have |= 1u << e;
        }
    }
    return ALL ? 0xFFFFFFFFu : have;
}

// phase 3 body: the hash goes to its sorted slot; the row only when the tile spans several genome rows
template <bool ALL, bool ROWS>
__device__ __forceinline__ void scatter_place_group(const unsigned long long (&hsh)[kStPerThread], uint32_t have, uint32_t key_bits,
                                                    uint32_t row0, bool one_row, uint32_t f, uint64_t pos0,
                                                    const ScatterParams& p, uint32_t* s_off, unsigned long long* s_rec,
                                                    uint16_t* s_row) {
#pragma unroll
    for (int e = 0; e < kStPerThread; ++e) {
        if (ALL || ((have >> e) & 1u)) {
            const uint32_t dst = atomicAdd(&s_off[(uint32_t)(hsh[e] >> key_bits)], 1u);   // s_off[b] walks through the bucket's run
            s_rec[dst] = hsh[e];
            if (ROWS) s_row[dst] = (uint16_t)(one_row ? row0 : row_of_position(p.file_stream_start, p.files, p.n_files, f, pos0 + e));
        }
    }
}

template <int KT>   // compile-time k, or 0 = p.k
__global__ void __launch_bounds__(kStThreads, 1)
k_scatter(const ScatterParams p) {
    extern __shared__ unsigned long long s_dyn[];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_warp[33];
    const uint32_t B = 1u << p.bucket_bits;
    unsigned long long* s_delta = s_dyn;                         // [B]   global base - tile offset
    unsigned long long* s_rec = s_delta + B;                     // [kStTile]
    uint32_t* s_cnt = reinterpret_cast<uint32_t*>(s_rec + kStTile);   // [B]
    uint32_t* s_off = s_cnt + B;                                 // [B]
    uint16_t* s_row = reinterpret_cast<uint16_t*>(s_off + B);    // [kStTile] genome rows, only for tiles that span several
    const uint64_t stream_len = p.scalars[S_STREAM_LEN];
    const uint64_t n_groups = (stream_len + 31) >> 5;
    const uint64_t n_tiles = (n_groups + kStThreads - 1) / kStThreads;
    const uint32_t k = KT ? (uint32_t)KT : p.k;
    const uint32_t kmask_lo = k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    const uint32_t kmask_hi = k <= 16 ? 0u : (k == 32 ? 0xFFFFFFFFu : ((1u << (2 * k - 32)) - 1u));
    const uint32_t kbits = k == 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    const uint32_t key_bits = 64 - p.bucket_bits;
    const uint32_t row_bits = p.row_bits;
    for (uint32_t i = threadIdx.x; i < B; i += kStThreads) s_cnt[i] = 0;
    __syncthreads();
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- phase 1: 32 hashed k-mers per thread into registers + tile histogram
        const uint64_t g = tile * kStThreads + threadIdx.x;
        unsigned long long cur_c = 0, prev_c = 0; uint32_t cur_v = 0, prev_v = 0;
        if (g < n_groups) {
            cur_c = p.codes[g]; cur_v = p.valid[g];
            if (g > 0) { prev_c = p.codes[g - 1]; prev_v = p.valid[g - 1]; }
        }
        unsigned long long hsh[kStPerThread];     // kept across the barriers
        uint32_t have = 0, row0 = 0, f = 0;
        bool one_row = true;
        const uint64_t pos0 = g * 32ULL;
        if (cur_v) {
            // 128-bit window: the 32 entries before mine (x[0], x[1]) and my 32 entries (x[2], x[3]); entry j at bits 2j
            const uint32_t x[4] = {(uint32_t)prev_c, (uint32_t)(prev_c >> 32), (uint32_t)cur_c, (uint32_t)(cur_c >> 32)};
            // invalid-entry bits, shifted so that bit e is the first entry of the window of my entry e
            const unsigned long long nvs = (~((unsigned long long)prev_v | ((unsigned long long)cur_v << 32))) >> (33 - k);
            const uint32_t n0 = (uint32_t)nvs, n1 = (uint32_t)(nvs >> 32);
            // reverse-complement source: window >> 2(33-k), complemented (code ^ 2); forward source: 2-bit reversed window
            const uint32_t S0 = 2 * (33 - k);
            unsigned long long Yl, Yh;
            if (S0 == 64) { Yl = cur_c; Yh = 0; }
            else { Yl = (prev_c >> S0) | (cur_c << (64 - S0)); Yh = cur_c >> S0; }
            const uint32_t y[4] = {(uint32_t)Yl ^ 0xAAAAAAAAu, (uint32_t)(Yl >> 32) ^ 0xAAAAAAAAu,
                                   (uint32_t)Yh ^ 0xAAAAAAAAu, (uint32_t)(Yh >> 32) ^ 0xAAAAAAAAu};
            const uint32_t r[4] = {rev2_32(x[3]), rev2_32(x[2]), rev2_32(x[1]), rev2_32(x[0])};
            f = p.tile_file[tile];
            while (f + 1 < p.n_files && p.file_stream_start[f + 1] <= pos0) ++f;
            const uint64_t next_start = (f + 1 < p.n_files) ? p.file_stream_start[f + 1] : ~0ULL;
            row0 = p.files[f].row;
            one_row = pos0 + kStPerThread <= next_start;
            // windows of my entries cover bits [0, 31 + k) of nvs
            const bool all_valid = (n0 | (n1 & ((1u << (k - 1)) - 1u))) == 0;
            if (all_valid) have = scatter_hash_group<true>(r, y, n0, n1, kbits, kmask_lo, kmask_hi, key_bits, s_cnt, hsh);
            else have = scatter_hash_group<false>(r, y, n0, n1, kbits, kmask_lo, kmask_hi, key_bits, s_cnt, hsh);
        }
        __syncthreads();
        // ---- phase 2: scan the tile histogram, reserve global space, clear the histogram.  Thread t owns bins
        // t, t + 512, ... (conflict-free); the tile is laid out thread-major, which is as good as bucket order.
        {
            uint32_t cnt[kStMaxBins];
            uint32_t sum = 0;
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kStMaxBins; ++i) {
                const uint32_t bin = i * kStThreads + threadIdx.x;
                cnt[i] = 0;
                if (bin < B) { cnt[i] = s_cnt[bin]; s_cnt[bin] = 0; sum += cnt[i]; }
            }
            unsigned long long gb[kStMaxBins];
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kStMaxBins; ++i) {      // all reservations in flight together
                const uint32_t bin = i * kStThreads + threadIdx.x;
                gb[i] = cnt[i] ? atomicAdd(&p.cursors[bin], (unsigned long long)cnt[i]) : 0ULL;
            }
            uint32_t total;
            uint32_t off = block_excl_scan<kStThreads>(sum, s_warp, total);
            if (threadIdx.x == 0) s_total = total;
#pragma unroll
            for (uint32_t i = 0; i < (uint32_t)kStMaxBins; ++i) {
                const uint32_t bin = i * kStThreads + threadIdx.x;
                if (bin < B) {
                    s_off[bin] = off;
                    if (cnt[i]) {
                        if (p.cap && gb[i] + cnt[i] > (unsigned long long)(bin + 1) * p.cap) {
                            *p.overflow = 1ULL;           // region too small: divert, the host re-runs the exact path
                            s_delta[bin] = p.dump;
                        } else {
                            s_delta[bin] = gb[i] - off;
                        }
                    }
                    off += cnt[i];
                }
            }
        }
        // does the tile lie inside one file (one genome row)?  Then the row is not staged per record.
        const uint32_t tf = p.tile_file[tile];
        const bool tile_one_row = (tf + 1 >= p.n_files) || p.file_stream_start[tf + 1] >= (tile + 1) * (uint64_t)kStTile;
        const uint32_t tile_row = p.files[tf].row;
        __syncthreads();
        // ---- phase 3: counting sort into shared memory
        if (tile_one_row) {
            if (have == 0xFFFFFFFFu) scatter_place_group<true, false>(hsh, have, key_bits, row0, one_row, f, pos0, p, s_off, s_rec, s_row);
            else if (have) scatter_place_group<false, false>(hsh, have, key_bits, row0, one_row, f, pos0, p, s_off, s_rec, s_row);
        } else if (have) {
            scatter_place_group<false, true>(hsh, have, key_bits, row0, one_row, f, pos0, p, s_off, s_rec, s_row);
        }
        __syncthreads();
        // ---- phase 4: copy out; consecutive threads write consecutive records of a bucket.  Record =
        // (hash << row_bits) | row: the top row_bits bits of the hash are bucket bits and fall off.  No barrier
        // after it: phase 1 of the next tile touches only s_cnt (already cleared) and registers.
        const uint32_t total = s_total;
        if (tile_one_row) {
#pragma unroll 4
            for (uint32_t i = threadIdx.x; i < total; i += kStThreads) {
                const unsigned long long h = s_rec[i];
                p.records[s_delta[(uint32_t)(h >> key_bits)] + i] = (h << row_bits) | tile_row;
            }
        } else {
            for (uint32_t i = threadIdx.x; i < total; i += kStThreads) {
                const unsigned long long h = s_rec[i];
                p.records[s_delta[(uint32_t)(h >> key_bits)] + i] = (h << row_bits) | s_row[i];
            }
        }
    }
}

// sum of one u64 per thread over a block of up to 1024 threads (valid in thread 0); a 64-bit atomicAdd on shared
// memory is a compare-and-swap loop, 1024 of them on one word took 40 us
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v) {
    __shared__ unsigned long long s_part[32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31u) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x + 31) / 32 ? s_part[threadIdx.x] : 0ULL;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_down_sync(0xffffffffu, t, d);
    }
    return t;
}

// After the scatter: end[b] = cursor[b], clamped to the region when regions are over-provisioned (an
// overflowing bucket is re-done by the exact path, but nothing may read past its region meanwhile);
// scalars[which] = number of records = n_windows.  One block.
__global__ void __launch_bounds__(1024)
k_finish_regions(const unsigned long long* __restrict__ begin, unsigned long long* __restrict__ end, uint32_t B,
                 unsigned long long cap, unsigned long long* __restrict__ scalars, int which) {
    unsigned long long v = 0;
    for (uint32_t b = threadIdx.x; b < B; b += blockDim.x) {
        unsigned long long e = end[b];
        if (cap && e > (unsigned long long)(b + 1) * cap) { e = (unsigned long long)(b + 1) * cap; end[b] = e; }
        v += e - begin[b];
    }
    const unsigned long long sum = block_sum_u64(v);
    if (threadIdx.x == 0) scalars[which] = sum;
}

// region starts for the over-provisioned layout: begin[b] = cursors[b] = b * cap
__global__ void k_init_regions(unsigned long long* __restrict__ begin, unsigned long long* __restrict__ cursors,
                               uint32_t B, unsigned long long cap) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) { begin[b] = (unsigned long long)b * cap; cursors[b] = (unsigned long long)b * cap; }
    if (b == B) begin[b] = (unsigned long long)B * cap;
}

// exclusive scan of the bucket histogram -> offsets[B+1]; cursors[b] = offsets[b]
__global__ void __launch_bounds__(1024)
k_bucket_offsets(unsigned long long* __restrict__ hist_cursor, unsigned long long* __restrict__ offsets, uint32_t B,
                 uint64_t* __restrict__ scalars, int total_scalar, uint32_t stride) {
    __shared__ unsigned long long s_part[1024];
    const uint32_t per = (B + 1023) / 1024;
    const uint32_t b0 = threadIdx.x * per;
    unsigned long long sum = 0;
    for (uint32_t i = 0; i < per; ++i) if (b0 + i < B) sum += hist_cursor[(size_t)(b0 + i) * stride];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    // simple Hillis-Steele inclusive scan
    for (int d = 1; d < 1024; d <<= 1) {
        unsigned long long v = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
        __syncthreads();
        s_part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned long long run = s_part[threadIdx.x] - sum;
    for (uint32_t i = 0; i < per; ++i) {
        if (b0 + i < B) {
            const unsigned long long c = hist_cursor[(size_t)(b0 + i) * stride];
            offsets[b0 + i] = run;
            hist_cursor[(size_t)(b0 + i) * stride] = run;
            run += c;
        }
    }
    if (threadIdx.x == 1023) { offsets[B] = s_part[1023]; scalars[total_scalar] = s_part[1023]; }
}

// ------------------------------------------------------------------------------------------
// aggregate: per-bucket shared-memory hash table
// ------------------------------------------------------------------------------------------
struct AggParams {
    const unsigned long long* records;   // (key << row_bits) | row
    const unsigned long long* begin;     // [B] first record of every bucket
    const unsigned long long* end;       // [B] one past its last record
    uint32_t B;
    uint32_t bucket_bits;
    uint32_t row_bits;
    uint32_t n_words;        // words per column handled here
    uint32_t slots;          // table capacity
    uint32_t keep_singletons;
    uint32_t mode;           // 0: final columns (k-mers), 1: partial columns (hash keys), 2: abundance filter
    uint32_t min_abundance;  // mode 2
    unsigned long long* out_keys;    // [cap]
    unsigned long long* out_words;   // [n_words][cap]  (mode 0/1)
    unsigned long long cap;
    unsigned long long* scalars;
    unsigned long long* bucket_out_counts;  // mode 1/2: entries emitted per bucket (atomic)
    unsigned long long* out_records;        // mode 2: filtered records, written at the bucket's own offset
    uint32_t b_begin, b_end;                // bucket range handled by this launch
    // mode 3 (owner-side merge of partial columns): records are refs (word offset << 8 | source)
    const unsigned long long* parts;
    uint32_t src_words[16];
    uint32_t src_woff[16];
};

__device__ __forceinline__ uint32_t slot_of(uint64_t key, uint32_t slots) {
    uint32_t hh = (uint32_t)key ^ (uint32_t)(key >> 32);
    hh *= 0x9E3779B1u;
    return __umulhi(hh, slots);
}

// Table key for MODE 0/1 is the hash key; for MODE 2 it is the whole record (key, row) and the
// single word per slot is an abundance counter.
template <int MODE>
__global__ void __launch_bounds__(kAggThreads, 1)
k_aggregate(const AggParams p) {
    extern __shared__ unsigned long long s_tab[];   // keys[slots] then words[n_words][slots]
    __shared__ uint32_t s_overflow, s_sp, s_cnt, s_kept, s_wr;
    __shared__ uint32_t s_depth[72];
    __shared__ unsigned long long s_idx[72];
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_red[kAggThreads / 32];
    const uint32_t slots = p.slots;
    const uint32_t W = (MODE == 2) ? 1u : p.n_words;
    constexpr bool kFinal = (MODE == 0 || MODE == 3);   // emits filtered k-mer columns
    unsigned long long* keys = s_tab;
    unsigned long long* words = s_tab + slots;
    const uint32_t key_bits = (MODE == 2) ? (64 - p.bucket_bits + p.row_bits) : (64 - p.bucket_bits);
    const uint64_t row_mask = (1ULL << p.row_bits) - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (uint32_t b = p.b_begin + blockIdx.x; b < p.b_end; b += gridDim.x) {
        const unsigned long long rbeg = p.begin[b], rend = p.end[b];
        if (rbeg == rend) continue;
        __syncthreads();
        if (threadIdx.x == 0) { s_sp = 1; s_depth[0] = 0; s_idx[0] = 0; s_wr = 0; }
        __syncthreads();
        while (true) {
            __syncthreads();
            if (s_sp == 0) break;
            const uint32_t depth = s_depth[s_sp - 1];
            const unsigned long long ridx = s_idx[s_sp - 1];
            __syncthreads();
            if (threadIdx.x == 0) { s_sp--; s_overflow = 0; s_cnt = 0; s_kept = 0; }
            for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) keys[i] = kEmptyKey;
            for (uint32_t i = threadIdx.x; i < slots * W; i += blockDim.x) words[i] = 0;
            __syncthreads();
            // ---- stream the bucket's records through the table
            for (unsigned long long r = rbeg + threadIdx.x; r < rend; r += blockDim.x) {
                if (*(volatile uint32_t*)&s_overflow) break;
                const unsigned long long rec = p.records[r];
                unsigned long long key; uint32_t row = 0;
                const unsigned long long* ent = nullptr;
                if (MODE == 2) key = rec;
                else if (MODE == 3) { ent = p.parts + (rec >> 8); key = ent[0] & ((1ULL << key_bits) - 1); }
                else { key = rec >> p.row_bits; row = (uint32_t)(rec & row_mask); }
                // records carry (bucket_bits - row_bits) redundant bucket bits above the key: mask them for the range test
                if (depth && ((key & (key_bits >= 64 ? ~0ULL : ((1ULL << key_bits) - 1))) >> (key_bits - depth)) != ridx) continue;
                uint32_t slot = slot_of(key, slots);
                bool hit = false;
                for (int probe = 0; probe < kMaxProbe; ++probe) {
                    unsigned long long k0 = *(volatile unsigned long long*)&keys[slot];
                    if (k0 == kEmptyKey) k0 = atomicCAS(&keys[slot], kEmptyKey, key);
                    if (k0 == kEmptyKey || k0 == key) { hit = true; break; }
                    slot = slot + 1 == slots ? 0 : slot + 1;
                }
                if (!hit) { s_overflow = 1; break; }
                if (MODE == 2) {
                    atomicAdd((uint32_t*)&words[slot], 1u);
                } else if (MODE == 3) {
                    const uint32_t src = (uint32_t)(rec & 255u), nw = p.src_words[src], wo = p.src_woff[src];
                    for (uint32_t w = 0; w < nw; ++w) {
                        const unsigned long long v = ent[1 + w];
                        if (v) atomicOr(&words[(wo + w) * slots + slot], v);
                    }
                } else {
                    const uint32_t bit = 63u - (row & 63u);      // utils.py:144-154
                    uint32_t* w32 = (uint32_t*)&words[(row >> 6) * slots + slot];
                    atomicOr(&w32[bit >> 5], 1u << (bit & 31u));
                }
            }
            __syncthreads();
            if (s_overflow) {
                // split this key range in two and retry (terminates: a range of one key needs one slot)
                if (threadIdx.x == 0) {
                    s_depth[s_sp] = depth + 1; s_idx[s_sp] = ridx * 2 + 1; s_sp++;
                    s_depth[s_sp] = depth + 1; s_idx[s_sp] = ridx * 2; s_sp++;
                    atomicAdd(&p.scalars[S_N_SPLITS], 1ULL);
                }
                continue;
            }
            // ---- count what this range emits
            uint32_t occ = 0, kept = 0;
            for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) {
                if (keys[i] != kEmptyKey) {
                    occ++;
                    if (MODE == 2) kept += ((uint32_t)words[i] >= p.min_abundance);
                    else if (MODE == 1) kept++;
                    else {
                        uint32_t pc = 0;
                        for (uint32_t w = 0; w < W; ++w) pc += __popcll(words[w * slots + i]);
                        kept += (pc >= 2 || p.keep_singletons);
                    }
                }
            }
            occ = __reduce_add_sync(0xffffffffu, occ);
            kept = __reduce_add_sync(0xffffffffu, kept);
            if (lane == 0) { atomicAdd(&s_cnt, occ); atomicAdd(&s_kept, kept); }
            __syncthreads();
            if (threadIdx.x == 0) {
                if (MODE == 2) {
                    s_base = rbeg + s_wr;           // filtered records stay inside the bucket's own range
                    s_wr += s_kept;
                } else {
                    s_base = atomicAdd(&p.scalars[S_U_NEEDED], (unsigned long long)s_kept);
                    atomicAdd(&p.scalars[S_N_DISTINCT], (unsigned long long)s_cnt);
                    if (MODE == 1) atomicAdd(&p.bucket_out_counts[b], (unsigned long long)s_kept);
                }
                s_cnt = 0;
            }
            __syncthreads();
            // ---- emit
            const unsigned long long base = s_base;
            for (uint32_t i0 = 0; i0 < slots; i0 += blockDim.x) {
                const uint32_t i = i0 + threadIdx.x;
                bool keep = false;
                unsigned long long key = 0;
                if (i < slots) {
                    key = keys[i];
                    if (key != kEmptyKey) {
                        if (MODE == 2) keep = ((uint32_t)words[i] >= p.min_abundance);
                        else if (MODE == 1) keep = true;
                        else {
                            uint32_t pc = 0;
                            for (uint32_t w = 0; w < W; ++w) pc += __popcll(words[w * slots + i]);
                            keep = (pc >= 2 || p.keep_singletons);
                        }
                    }
                }
                const uint32_t m = __ballot_sync(0xffffffffu, keep);
                uint32_t wbase = 0;
                if (lane == 0 && m) wbase = atomicAdd(&s_cnt, __popc(m));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (keep) {
                    const unsigned long long o = base + wbase + __popc(m & lanemask_lt());
                    if (MODE == 2) {
                        p.out_records[o] = key;
                    } else if (o < p.cap) {
                        const unsigned long long h = ((unsigned long long)b << key_bits) | key;
                        p.out_keys[o] = kFinal ? kunhash(h) : h;
                        for (uint32_t w = 0; w < W; ++w) p.out_words[w * p.cap + o] = words[w * slots + i];
                    }
                }
            }
        }
        if (MODE == 2 && threadIdx.x == 0) {
            p.bucket_out_counts[b] = s_wr;
            atomicAdd(&p.scalars[S_N_SOLID], (unsigned long long)s_wr);
        }
    }
    (void)warp; (void)s_red;
}

// ---- column aggregation (modes 0 = final columns, 1 = partial columns, 3 = owner-side merge of partials) ----
// One CTA per hash bucket.  The bucket's records stream through a shared-memory table
//   keys[slots + tail] (u64), w32[2W][slots + tail] (presence half-words), kept[slots + tail] (u8)
// with eight coalesced loads in flight per thread.  Genome row g lives in half-word plane (g >> 5) ^ 1 at bit
// 31 - (g & 31), so that word w = w32[2w+1] : w32[2w] has row g at bit 63 - (g & 63) (kover/utils.py:144-154).
//
// The home slot is MONOTONE in the key (its top bits scaled to the table; the multiplicative hash already
// spreads them uniformly) and probing never wraps (tail slots), so clusters hold disjoint ascending key ranges:
// the table is almost sorted, and a key's exact rank is (kept slots before it) corrected by the inversions
// inside its own cluster.  Columns therefore leave the kernel in ascending hash order per bucket with no
// sort pass.  A bucket that does not fit is split into key sub-ranges (processed in ascending order): one
// counting sweep to size the bucket's output, one emitting sweep.
constexpr int kAggBatch = 8;

struct AggTable {
    unsigned long long* keys;
    uint32_t* w32;
    uint8_t* kept;         // [total rounded up to 4] bit 7: the slot's column is kept; bits 0-6: inversions found by LATER slots
    uint8_t* dec;          // [total] inversions the slot found itself (larger kept keys between its home and itself)
    uint32_t slots;        // home slots
    uint32_t total;        // slots + tail
};

__device__ __forceinline__ uint32_t home_slot(unsigned long long key, uint32_t shift, uint32_t slots) {
    // top 32 bits of the key below the bits that are fixed inside the current (sub-)range; shift = 64 - key_bits + depth
    return __umulhi((uint32_t)((key << shift) >> 32), slots);
}

struct AggParams2 {
    const unsigned long long* records;   // MODE 0/1: (hash << row_bits) | row;  MODE 3: refs (word offset << 8) | source
    const unsigned long long* begin;     // [B]
    const unsigned long long* end;       // [B]
    uint32_t bucket_bits, row_bits, n_words, slots, keep_singletons;
    uint32_t sub_bits;                   // every bucket is aggregated as 2^sub_bits key sub-ranges (virtual buckets), one CTA pass each
    uint32_t wide_words, wide_stride;    // MODE 4/5: presence words per wide record, u64 per record
    uint32_t table_u32;                  // u32 cells per table slot: 2 x n_words presence half-words; MODE 3: one entry reference per source (even)
    unsigned long long* out_keys;        // [cap]   bucket chunks, at bucket_base[b]
    unsigned long long* out_words;       // [n_words][cap]
    unsigned long long cap;
    unsigned long long* scalars;
    unsigned long long* bucket_base;     // [B] where the bucket's chunk starts in out_*
    unsigned long long* bucket_count;    // [B] columns the bucket emitted
    uint32_t b_begin, b_end;             // bucket_base / bucket_count are indexed by (virtual bucket) - (b_begin << sub_bits)
    // ordered emission (final build in hash order): the virtual buckets are dealt by ticket, every bucket publishes its
    // column count and learns its offset by look-back over its predecessors, so the columns land at their final place:
    // out_keys is the result's k-mer array, word row 0 goes to out_row0 (row 0 of the result matrix, whose stride is
    // only known at the end), rows >= 1 to out_words at stride cap
    uint32_t ordered;
    unsigned long long* out_row0;        // nullptr: row 0 goes to out_words like the others
    unsigned long long* pub;             // [virtual buckets] bit 63: inclusive prefix, bit 62: own count; zeroed before the launch
    unsigned int* ticket;                // zeroed before the launch
    // MODE 3: partial columns [hash, words...] of n_src sources, each list ascending by hash; the entries of
    // bucket b in source s are bounds[s * (b_end - b_begin + 1) + (b - b_begin)] .. [.. + 1]
    const unsigned long long* parts;
    const unsigned long long* bounds;
    uint32_t n_src;
    unsigned long long src_off[16];      // first u64 word of the source in parts
    uint32_t src_words[16];
    uint32_t src_woff[16];
};

// Probe loops split warps: a lane that finds its slot at the first probe does not wait for its neighbours by itself,
// and everything after the loop would then run once per fragment (4.2 fragments per warp were measured,
// profiles/r01_v3).  So every lane of the warp runs every iteration of the record loops (inactive ones predicated
// off) and the warp is re-converged with __syncwarp() right after each probe loop.
__device__ __forceinline__ bool agg_insert(unsigned long long* keys, uint32_t& slot, unsigned long long key, bool active) {
    bool ok = active;
    if (active) {
        int probe = 0;
        while (true) {
            unsigned long long k0 = *(volatile unsigned long long*)&keys[slot];
            if (k0 == key) break;
            if (k0 == kEmptyKey) {
                k0 = atomicCAS(&keys[slot], kEmptyKey, key);
                if (k0 == kEmptyKey || k0 == key) break;
            }
            ++slot;
            if (++probe >= kMaxProbe) { ok = false; break; }
        }
    }
    __syncwarp();
    return ok;
}

template <bool FILTER>
__device__ __forceinline__ void agg_stream(const AggParams2& p, const AggTable& t, const unsigned long long* __restrict__ recs,
                                           uint32_t n, uint32_t key_bits, uint32_t depth, unsigned long long ridx,
                                           volatile uint32_t* overflow) {
    const uint32_t row_bits = p.row_bits;
    const uint32_t row_mask = (1u << row_bits) - 1u;
    const uint32_t shift = 64 - key_bits + depth;
    const uint32_t slots = t.slots, total = t.total;
    unsigned long long* const keys = t.keys;
    uint32_t* const w32 = t.w32;
    for (uint32_t base0 = 0; base0 < n; base0 += kAggThreads * kAggBatch) {
        const uint32_t base = base0 + threadIdx.x;
        unsigned long long r[kAggBatch];
#pragma unroll
        for (int j = 0; j < kAggBatch; ++j) {
            const uint32_t idx = base + j * kAggThreads;
            r[j] = idx < n ? __ldcs(recs + idx) : 0ULL;
        }
        if (__any_sync(0xffffffffu, *overflow != 0)) break;
#pragma unroll
        for (int j = 0; j < kAggBatch; ++j) {
            const unsigned long long key = r[j] >> row_bits;      // carries (bucket_bits - row_bits) redundant bucket bits on top
            bool act = base + j * kAggThreads < n;
            if (FILTER) act = act && ((key << (64 - key_bits)) >> (64 - depth)) == ridx;
            uint32_t slot = home_slot(key, shift, slots);
            const bool ok = agg_insert(keys, slot, key, act);
            if (act) {
                if (ok) {
                    const uint32_t row = (uint32_t)r[j] & row_mask;
                    atomicOr(&w32[((row >> 5) ^ 1u) * total + slot], 0x80000000u >> (row & 31u));
                } else *overflow = 1;
            }
        }
    }
}

// MODE 4 / 5: wide records [(hash << row_bits) | genome group, wide_words presence words (, padding)] of
// wide_stride u64 each (the unit path, grmkm_units.cuh); the words of group g are matrix words g * wide_words ..
template <bool FILTER>
__device__ __forceinline__ void agg_stream_wide(const AggParams2& p, const AggTable& t, const unsigned long long* __restrict__ recs,
                                                uint32_t n, uint32_t key_bits, uint32_t depth, unsigned long long ridx,
                                                volatile uint32_t* overflow) {
    const uint32_t wbits = p.row_bits;
    const uint32_t wmask = (1u << wbits) - 1u;
    const uint32_t shift = 64 - key_bits + depth;
    const uint32_t slots = t.slots, total = t.total;
    const uint32_t WB = p.wide_words, stride2 = p.wide_stride / 2;          // record = stride2 x 16 bytes
    unsigned long long* const keys = t.keys;
    uint32_t* const w32 = t.w32;
    const ulonglong2* const rec2 = reinterpret_cast<const ulonglong2*>(recs);
    constexpr int kWideBatch = 4;
    for (uint32_t base0 = 0; base0 < n; base0 += kAggThreads * kWideBatch) {
        const uint32_t base = base0 + threadIdx.x;
        ulonglong2 r[kWideBatch], r2[kWideBatch];                  // [key, word 0], [word 1, word 2]
#pragma unroll
        for (int j = 0; j < kWideBatch; ++j) {
            const uint32_t idx = base + j * kAggThreads;
            r[j] = idx < n ? __ldcs(rec2 + (size_t)idx * stride2) : make_ulonglong2(0ULL, 0ULL);
            r2[j] = (idx < n && WB > 1) ? __ldcs(rec2 + (size_t)idx * stride2 + 1) : make_ulonglong2(0ULL, 0ULL);
        }
        if (__any_sync(0xffffffffu, *overflow != 0)) break;
#pragma unroll
        for (int j = 0; j < kWideBatch; ++j) {
            const uint32_t idx = base + j * kAggThreads;
            const unsigned long long key = r[j].x >> wbits;
            bool act = idx < n;
            if (FILTER) act = act && ((key << (64 - key_bits)) >> (64 - depth)) == ridx;
            uint32_t slot = home_slot(key, shift, slots);
            const bool ok = agg_insert(keys, slot, key, act);
            if (act) {
                if (ok) {
                    const uint32_t w0 = ((uint32_t)r[j].x & wmask) * WB;
                    {
                        const uint32_t vlo = (uint32_t)r[j].y, vhi = (uint32_t)(r[j].y >> 32);
                        if (vlo) atomicOr(&w32[(2 * w0) * total + slot], vlo);
                        if (vhi) atomicOr(&w32[(2 * w0 + 1) * total + slot], vhi);
                    }
                    for (uint32_t w = 1; w < WB; ++w) {
                        const unsigned long long v = w == 1 ? r2[j].x : w == 2 ? r2[j].y : __ldcs(recs + (size_t)idx * p.wide_stride + 1 + w);
                        const uint32_t vlo = (uint32_t)v, vhi = (uint32_t)(v >> 32);
                        if (vlo) atomicOr(&w32[(2 * (w0 + w)) * total + slot], vlo);
                        if (vhi) atomicOr(&w32[(2 * (w0 + w) + 1) * total + slot], vhi);
                    }
                } else *overflow = 1;
            }
        }
    }
}

// MODE 3: the bucket's entries are one contiguous range per source (the lists arrive sorted by hash)
template <bool FILTER>
__device__ __forceinline__ void agg_stream_parts(const AggParams2& p, const AggTable& t, uint32_t b, uint32_t key_bits,
                                                 uint32_t depth, unsigned long long ridx, volatile uint32_t* overflow) {
    const uint32_t shift = 64 - key_bits + depth;
    const unsigned long long key_mask = (1ULL << key_bits) - 1;
    const uint32_t nb1 = p.b_end - p.b_begin + 1;
    for (uint32_t s = 0; s < p.n_src; ++s) {
        const unsigned long long lo = p.bounds[(size_t)s * nb1 + (b - p.b_begin)], hi = p.bounds[(size_t)s * nb1 + (b - p.b_begin) + 1];
        const uint32_t width = 1 + p.src_words[s];
        const unsigned long long* src = p.parts + p.src_off[s];
        for (unsigned long long i0 = lo; i0 < hi; i0 += kAggThreads) {
            if (__any_sync(0xffffffffu, *overflow != 0)) break;
            const unsigned long long i = i0 + threadIdx.x;
            bool act = i < hi;
            const unsigned long long* ent = src + (act ? i : lo) * width;
            const unsigned long long key = ent[0] & key_mask;
            if (FILTER) act = act && (key >> (key_bits - depth)) == ridx;
            uint32_t slot = home_slot(key, shift, t.slots);
            const bool ok = agg_insert(t.keys, slot, key, act);
            if (act && !ok) *overflow = 1;
            // the slot keeps a REFERENCE to the source's entry (a source lists a key once), not its words: with 16 word-rows a
            // slot of words is 137 bytes, a slot of 8 references 41, so the table holds 3.3 x the keys and the owner needs
            // that many fewer bucket passes (N = 8: merge 1.9 ms with words in the table)
            if (act && ok) t.w32[s * t.total + slot] = (uint32_t)i + 1u;
        }
    }
}

// bounds[s][j] = first entry of source s whose hash is >= (b_begin + j) << key_bits  (j = 0 .. nb; j with
// b_begin + j == 2^bucket_bits means "the end")
__global__ void k_merge_bounds(const unsigned long long* __restrict__ parts, uint32_t n_src, uint32_t nb1, uint32_t b_begin,
                               uint32_t bucket_bits, const unsigned long long* __restrict__ src_off,
                               const unsigned long long* __restrict__ src_count, const uint32_t* __restrict__ src_width,
                               unsigned long long* __restrict__ bounds) {
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (uint64_t)n_src * nb1) return;
    const uint32_t s = (uint32_t)(id / nb1), j = (uint32_t)(id % nb1);
    const unsigned long long n = src_count[s];
    const uint64_t bb = (uint64_t)b_begin + j;
    unsigned long long lo = 0, hi = n;
    if (bb >= (1ULL << bucket_bits)) lo = n;
    else {
        const unsigned long long target = bb << (64 - bucket_bits);
        const unsigned long long* src = parts + src_off[s];
        const uint32_t width = src_width[s];
        while (lo < hi) {
            const unsigned long long mid = (lo + hi) >> 1;
            if (src[mid * width] < target) lo = mid + 1; else hi = mid;
        }
    }
    bounds[id] = lo;
}

// Emission.  Slots are handled in chunks of kAggThreads consecutive slots, lane = slot: shared-memory accesses are
// conflict-free, the inversion walks of neighbouring lanes touch neighbouring slots, and the columns leave in
// (almost) consecutive order.
// Pass A: kept flags; kept slots per (chunk, warp) in s_wc.  Returns this thread's occupied count.
constexpr int kAggMaxChunks = (kAggMaxSlots + kMaxProbe + kAggThreads - 1) / kAggThreads;
static_assert(kAggMaxChunks * (kAggThreads / 32) <= kAggThreads, "the scan of the per-(chunk, warp) counts takes one value per thread");
template <int MODE>
__device__ __forceinline__ uint32_t agg_mark(const AggParams2& p, const AggTable& t, uint32_t* s_wc) {
    uint32_t occ = 0;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t c = 0;
    for (uint32_t i0 = 0; i0 < t.total; i0 += kAggThreads, ++c) {
        const uint32_t i = i0 + threadIdx.x;
        uint32_t kf = 0;
        if (i < t.total && t.keys[i] != kEmptyKey) {
            occ++;
            if (MODE == 1 || MODE == 5 || p.keep_singletons) kf = 1;
            else if (MODE == 3) {
                uint32_t pc = 0;
                for (uint32_t s = 0; s < p.n_src; ++s) {
                    const uint32_t e = t.w32[s * t.total + i];
                    if (e) {
                        const unsigned long long* ent = p.parts + p.src_off[s] + (unsigned long long)(e - 1u) * (1u + p.src_words[s]);
                        for (uint32_t w = 0; w < p.src_words[s]; ++w) pc += __popcll(ent[1 + w]);
                    }
                }
                kf = pc >= 2;
            } else {
                uint32_t pc = 0;
                for (uint32_t h = 0; h < 2 * p.n_words; ++h) pc += __popc(t.w32[h * t.total + i]);
                kf = pc >= 2;
            }
        }
        if (i < t.total) t.kept[i] = (uint8_t)(kf << 7);        // bit 7 = kept, bits 0-6 = the rank correction of agg_fix
        const uint32_t bal = __ballot_sync(0xffffffffu, kf != 0);
        if (lane == 0) s_wc[c * (kAggThreads / 32) + warp] = (uint32_t)__popc(bal);
    }
    return occ;
}

// Rank of a kept slot = kept slots before it, corrected by the inversions it takes part in.  The home slot is monotone
// in the key and probing never wraps, so a LATER slot i can only be out of order with the slots [home(key_i), i): a
// larger key at j < i has home(key_j) >= home(key_i) and sits at or after its home.  Every inversion is therefore found
// by its later element with a walk as long as that element's displacement (0.5 slots on average at half load, instead
// of the whole cluster in both directions).
// Pass B1: the later element of an inversion adds one to the earlier element's correction (bits 0-6 of its flag byte;
// at most kMaxProbe - 1 = 95 later slots can reach back to a slot, so the counter never touches bit 7).
__device__ __forceinline__ void agg_fix(const AggTable& t, uint32_t shift) {
    uint32_t* const kept32 = reinterpret_cast<uint32_t*>(t.kept);
    for (uint32_t i = threadIdx.x; i < t.total; i += kAggThreads) {
        if (!(t.kept[i] & 0x80u)) continue;
        const unsigned long long key = t.keys[i];
        uint32_t dec = 0;
        for (uint32_t j = home_slot(key, shift, t.slots); j < i; ++j)
            if (t.keys[j] > key && (t.kept[j] & 0x80u)) { atomicAdd(&kept32[j >> 2], 1u << (8u * (j & 3u))); ++dec; }
        t.dec[i] = (uint8_t)dec;             // the emission does not walk again
    }
}

// Pass B2: emission at out[base + rank], rank = kept slots before + correction - own inversions.  s_wp = exclusive scan of s_wc.
template <int MODE>
__device__ __forceinline__ void agg_emit(const AggParams2& p, const AggTable& t, const uint32_t* s_wp,
                                         unsigned long long base, uint32_t b, uint32_t key_bits, uint32_t shift) {
    const unsigned long long key_mask = (1ULL << key_bits) - 1;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t c = 0;
    for (uint32_t i0 = 0; i0 < t.total; i0 += kAggThreads, ++c) {
        const uint32_t i = i0 + threadIdx.x;
        const uint32_t flag = i < t.total ? t.kept[i] : 0u;
        const bool kf = (flag & 0x80u) != 0;
        const uint32_t bal = __ballot_sync(0xffffffffu, kf);
        if (!kf) continue;
        const unsigned long long key = t.keys[i];
        const uint32_t rank = s_wp[c * (kAggThreads / 32) + warp] + (uint32_t)__popc(bal & ((1u << lane) - 1u)) + (flag & 0x7Fu) - t.dec[i];
        const unsigned long long o = base + rank;
        if (o < p.cap) {
            const unsigned long long h = ((unsigned long long)b << key_bits) | (key & key_mask);
            p.out_keys[o] = (MODE == 1 || MODE == 5) ? h : kunhash(h);
            if (MODE == 3) {
                for (uint32_t s = 0; s < p.n_src; ++s) {
                    const uint32_t e = t.w32[s * t.total + i], nw = p.src_words[s], wo = p.src_woff[s];
                    const unsigned long long* ent = p.parts + p.src_off[s] + (unsigned long long)(e ? e - 1u : 0u) * (1u + nw);
                    for (uint32_t w = 0; w < nw; ++w) p.out_words[(unsigned long long)(wo + w) * p.cap + o] = e ? ent[1 + w] : 0ULL;
                }
            } else {
                for (uint32_t w = 0; w < p.n_words; ++w) {
                    unsigned long long* const row = (w == 0 && p.out_row0) ? p.out_row0 : p.out_words + w * p.cap;
                    row[o] = ((unsigned long long)t.w32[(2 * w + 1) * t.total + i] << 32) | t.w32[2 * w * t.total + i];
                }
            }
        }
    }
}

// Ordered emission: where does (virtual) bucket idx start?  Called by the whole first warp.  The bucket publishes its own
// count, sums its predecessors' words 32 at a time until one of them carries an inclusive prefix, then publishes its
// own inclusive prefix.  The words are self-validating (flag bits + value in one 64-bit store), so relaxed accesses do;
// buckets are dealt by ticket, so every predecessor belongs to a CTA that is already running (no deadlock whatever
// the residency).
constexpr unsigned long long kAggPubInc = 1ULL << 63, kAggPubOwn = 1ULL << 62, kAggPubMask = (1ULL << 62) - 1;
__device__ __forceinline__ unsigned long long agg_reserve_ordered(const AggParams2& p, uint32_t idx, unsigned long long total) {
    const uint32_t lane = threadIdx.x & 31u;
    if (lane == 0) {
        atomicAdd(&p.scalars[S_U_NEEDED], total);
        st_relaxed_u64(p.pub + idx, (idx == 0 ? kAggPubInc : kAggPubOwn) | total);
    }
    if (idx == 0) return 0;
    unsigned long long excl = 0;
    for (long long j = (long long)idx - 1;; j -= 32) {
        const long long q = j - (long long)lane;
        unsigned long long v;
        do { v = q >= 0 ? ld_relaxed_u64(p.pub + q) : kAggPubInc; } while (!__all_sync(0xffffffffu, (v >> 62) != 0));
        const uint32_t inc = __ballot_sync(0xffffffffu, (v & kAggPubInc) != 0);
        unsigned long long val = v & kAggPubMask;
        if (inc && lane > (uint32_t)(__ffs(inc) - 1)) val = 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        excl += val;
        if (inc) break;
    }
    if (lane == 0) st_relaxed_u64(p.pub + idx, kAggPubInc | (excl + total));
    return excl;
}

template <int MODE>
__global__ void __launch_bounds__(kAggThreads, kAggCtasPerSm)
k_aggregate_cols(const AggParams2 p) {
    extern __shared__ unsigned long long s_tab[];
    __shared__ uint32_t s_overflow, s_sp;
    __shared__ uint32_t s_depth[72];
    __shared__ unsigned long long s_idx[72];
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_base;
    AggTable t;
    t.slots = p.slots; t.total = p.slots + kMaxProbe;
    t.keys = s_tab;
    t.w32 = reinterpret_cast<uint32_t*>(s_tab + t.total);
    t.kept = reinterpret_cast<uint8_t*>(t.w32 + (size_t)p.table_u32 * t.total);
    t.dec = t.kept + ((t.total + 3u) & ~3u);
    const uint32_t key_bits = 64 - p.bucket_bits;
    const uint32_t n_chunks = (t.total + kAggThreads - 1) / kAggThreads;     // <= kAggMaxChunks (slots <= 16384)
    __shared__ uint32_t s_wc[kAggMaxChunks * (kAggThreads / 32)];

    // A virtual bucket = bucket b restricted to the key sub-range `sub` of 2^sub_bits: the scatter can then use
    // 2^sub_bits fewer buckets (longer runs per tile) than the table size demands.  The CTAs of one bucket's
    // sub-ranges have adjacent block indices, run at the same time and share the bucket's records through L2.
    const uint32_t sb = p.sub_bits;
    const uint32_t vb_base = p.b_begin << sb;
    __shared__ uint32_t s_vb;
    for (uint32_t it = 0;; ++it) {
        uint32_t vb = vb_base + blockIdx.x + it * gridDim.x;
        if (p.ordered) {
            __syncthreads();
            if (threadIdx.x == 0) s_vb = atomicAdd(p.ticket, 1u);
            __syncthreads();
            vb = vb_base + s_vb;
        }
        if (vb >= (p.b_end << sb)) break;
        const uint32_t b = vb >> sb, sub = vb & ((1u << sb) - 1u);
        uint32_t n = 0;
        const unsigned long long* recs = nullptr;
        if (MODE == 3) {
            const uint32_t nb1 = p.b_end - p.b_begin + 1;
            for (uint32_t s = 0; s < p.n_src; ++s)
                n += (uint32_t)(p.bounds[(size_t)s * nb1 + (b - p.b_begin) + 1] - p.bounds[(size_t)s * nb1 + (b - p.b_begin)]);
        } else {
            const unsigned long long rbeg = p.begin[b], rend = p.end[b];
            n = rbeg < rend ? (uint32_t)(rend - rbeg) : 0u;
            recs = p.records + (MODE >= 4 ? (unsigned long long)p.wide_stride * rbeg : rbeg);
        }
        if (n == 0) {
            if (threadIdx.x == 0) { p.bucket_base[vb - vb_base] = 0; p.bucket_count[vb - vb_base] = 0; }
            // an empty bucket only publishes its (zero) count: its successors look past it
            if (p.ordered && threadIdx.x == 0) st_relaxed_u64(p.pub + (vb - vb_base), (vb == vb_base ? kAggPubInc : kAggPubOwn));
            continue;
        }
        // phase 0: the whole (virtual) bucket in one table; on overflow phase 1 counts over its key sub-ranges and
        // phase 2 emits them in ascending order
        uint32_t phase = 0;
        uint32_t bucket_total = 0, emitted = 0, bucket_occ = 0, splits = 0;
        __syncthreads();
        if (threadIdx.x == 0) { s_sp = 1; s_depth[0] = sb; s_idx[0] = sub; }
        while (true) {
            __syncthreads();
            if (s_sp == 0) {
                if (phase != 1) break;
                __syncthreads();          // everyone has seen the empty stack before it is refilled
                // counting sweep done: reserve the bucket's chunk, then emit
                if (p.ordered) {
                    if (threadIdx.x < 32) {
                        const unsigned long long base = agg_reserve_ordered(p, vb - vb_base, bucket_total);
                        if (threadIdx.x == 0) s_base = base;
                    }
                } else if (threadIdx.x == 0) s_base = atomicAdd(&p.scalars[S_U_NEEDED], (unsigned long long)bucket_total);
                if (threadIdx.x == 0) {
                    s_sp = 0;
                    s_depth[s_sp] = sb + 1; s_idx[s_sp] = 2ULL * sub + 1; s_sp++;
                    s_depth[s_sp] = sb + 1; s_idx[s_sp] = 2ULL * sub; s_sp++;
                }
                phase = 2;
                continue;
            }
            const uint32_t depth = s_depth[s_sp - 1];
            const unsigned long long ridx = s_idx[s_sp - 1];
            __syncthreads();
            if (threadIdx.x == 0) { s_sp--; s_overflow = 0; }
            for (uint32_t i = threadIdx.x; i < t.total; i += kAggThreads) t.keys[i] = kEmptyKey;
            for (uint32_t i = threadIdx.x; i < t.total * (p.table_u32 / 2); i += kAggThreads) reinterpret_cast<unsigned long long*>(t.w32)[i] = 0;
            __syncthreads();
            if (MODE == 3) {
                if (depth == 0) agg_stream_parts<false>(p, t, b, key_bits, 0, 0, &s_overflow);
                else agg_stream_parts<true>(p, t, b, key_bits, depth, ridx, &s_overflow);
            } else if (MODE >= 4) {
                if (depth == 0) agg_stream_wide<false>(p, t, recs, n, key_bits, 0, 0, &s_overflow);
                else agg_stream_wide<true>(p, t, recs, n, key_bits, depth, ridx, &s_overflow);
            } else {
                if (depth == 0) agg_stream<false>(p, t, recs, n, key_bits, 0, 0, &s_overflow);
                else agg_stream<true>(p, t, recs, n, key_bits, depth, ridx, &s_overflow);
            }
            __syncthreads();
            if (s_overflow) {
                // split this key range in two, lower half first (terminates: a range of one key needs one slot)
                if (threadIdx.x == 0) {
                    if (phase == 0) {
                        s_depth[s_sp] = sb + 1; s_idx[s_sp] = 2ULL * sub + 1; s_sp++;
                        s_depth[s_sp] = sb + 1; s_idx[s_sp] = 2ULL * sub; s_sp++;
                    } else {
                        s_depth[s_sp] = depth + 1; s_idx[s_sp] = ridx * 2 + 1; s_sp++;
                        s_depth[s_sp] = depth + 1; s_idx[s_sp] = ridx * 2; s_sp++;
                    }
                }
                if (phase == 0) phase = 1;
                if (phase == 1) splits++;
                continue;
            }
            const uint32_t occ = agg_mark<MODE>(p, t, s_wc);
            __syncthreads();
            uint32_t total_kept, total_occ;
            {
                constexpr uint32_t kWarps = kAggThreads / 32;
                const uint32_t v = threadIdx.x < n_chunks * kWarps ? s_wc[threadIdx.x] : 0u;
                const uint32_t e = block_excl_scan<kAggThreads>(v, s_warp, total_kept);
                if (threadIdx.x < n_chunks * kWarps) s_wc[threadIdx.x] = e;
            }
            if (phase != 2) { block_excl_scan<kAggThreads>(occ, s_warp, total_occ); bucket_occ += total_occ; }
            if (phase == 1) { bucket_total += total_kept; continue; }
            if (phase == 0) {
                bucket_total = total_kept;
                if (p.ordered) {
                    if (threadIdx.x < 32) {
                        const unsigned long long base = agg_reserve_ordered(p, vb - vb_base, total_kept);
                        if (threadIdx.x == 0) s_base = base;
                    }
                } else if (threadIdx.x == 0) s_base = atomicAdd(&p.scalars[S_U_NEEDED], (unsigned long long)total_kept);
            }
            agg_fix(t, 64 - key_bits + depth);
            __syncthreads();
            agg_emit<MODE>(p, t, s_wc, s_base + emitted, b, key_bits, 64 - key_bits + depth);
            emitted += total_kept;
        }
        if (threadIdx.x == 0) {
            p.bucket_base[vb - vb_base] = s_base;
            p.bucket_count[vb - vb_base] = bucket_total;
            atomicAdd(&p.scalars[S_N_DISTINCT], (unsigned long long)bucket_occ);
            if (splits) atomicAdd(&p.scalars[S_N_SPLITS], (unsigned long long)splits);
        }
    }
}

// ordered emission: word rows >= 1 from stride cap (the aggregate's output) to the result's stride U, which only the
// device knows when this is launched (no host round trip between the aggregate and the end of the build)
__global__ void __launch_bounds__(256)
k_move_rows(const unsigned long long* __restrict__ src, unsigned long long cap, unsigned long long* __restrict__ dst,
            const unsigned long long* __restrict__ u_ptr, uint32_t W) {
    const unsigned long long U = *u_ptr;
    if (U > cap) return;                                   // the aggregate overflowed its guess: the host repeats it
    const unsigned long long n = U * (W - 1);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long w = 1 + i / U, j = i - (w - 1) * U;
        dst[w * U + j] = __ldcs(src + w * cap + j);
    }
}

// bucket chunks (arbitrary order in tmp) -> final arrays in bucket order: one CTA per bucket
__global__ void __launch_bounds__(256)
k_gather_buckets(const unsigned long long* __restrict__ tmp_keys, const unsigned long long* __restrict__ tmp_words,
                 unsigned long long tmp_cap, const unsigned long long* __restrict__ bucket_base,
                 const unsigned long long* __restrict__ offsets, uint32_t B, uint32_t W, unsigned long long U,
                 unsigned long long* __restrict__ keys, unsigned long long* __restrict__ words, unsigned long long out_stride) {
    for (uint32_t b = blockIdx.x; b < B; b += gridDim.x) {
        const unsigned long long src = bucket_base[b], dst = offsets[b], n = offsets[b + 1] - dst;
        for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
            keys[dst + i] = tmp_keys[src + i];
            for (uint32_t w = 0; w < W; ++w) words[(unsigned long long)w * out_stride + dst + i] = tmp_words[(unsigned long long)w * tmp_cap + src + i];
        }
    }
    (void)U;
}

// ------------------------------------------------------------------------------------------
// multi-GPU: partial columns out (AoS records for the all-to-all) and owner-side partition
// ------------------------------------------------------------------------------------------
// record i = [hash, word_0 .. word_{W-1}]
// The same gather, fused with the all-to-all: owner d's slice (buckets [first[d], first[d + 1])) is stored straight into
// rank d's receive buffer over NVLink (peer pointers from torch symmetric memory), at the word offset the counts
// exchange assigned to this rank.  No send buffer, no NCCL send / recv; the ranks meet at one barrier afterwards.
struct PeerSlices {
    unsigned long long* dst[16];          // peer d's receive buffer + this rank's word offset in it
    uint32_t first[17];                   // first (virtual) bucket of owner d; first[n] = B
    uint32_t n;
};
__global__ void __launch_bounds__(256)
k_gather_buckets_peers(const unsigned long long* __restrict__ tmp_keys, const unsigned long long* __restrict__ tmp_words,
                       unsigned long long tmp_cap, const unsigned long long* __restrict__ bucket_base,
                       const unsigned long long* __restrict__ offsets, uint32_t B, uint32_t W, const PeerSlices ps) {
    for (uint32_t b = blockIdx.x; b < B; b += gridDim.x) {
        uint32_t d = 0;
        while (d + 1 < ps.n && b >= ps.first[d + 1]) ++d;
        const unsigned long long src = bucket_base[b], d0 = offsets[b] - offsets[ps.first[d]], n = offsets[b + 1] - offsets[b];
        unsigned long long* __restrict__ dst = ps.dst[d] + d0 * (1 + W);
        const unsigned long long cells = n * (1 + W);
        for (unsigned long long c = threadIdx.x; c < cells; c += blockDim.x) {
            const unsigned long long i = c / (1 + W);
            const uint32_t f = (uint32_t)(c - i * (1 + W));
            dst[c] = f == 0 ? tmp_keys[src + i] : tmp_words[(unsigned long long)(f - 1) * tmp_cap + src + i];
        }
    }
}

// bucket chunks (completion order) -> AoS records [hash, word_0 .. word_{W-1}] in bucket order: the gather and the
// export of a partial build in one pass
__global__ void __launch_bounds__(256)
k_gather_buckets_aos(const unsigned long long* __restrict__ tmp_keys, const unsigned long long* __restrict__ tmp_words,
                     unsigned long long tmp_cap, const unsigned long long* __restrict__ bucket_base,
                     const unsigned long long* __restrict__ offsets, uint32_t B, uint32_t W, unsigned long long* __restrict__ dst) {
    for (uint32_t b = blockIdx.x; b < B; b += gridDim.x) {
        const unsigned long long src = bucket_base[b], d0 = offsets[b], n = offsets[b + 1] - d0;
        // one u64 of the AoS output per thread and step: coalesced stores, the loads hit W + 1 runs
        const unsigned long long cells = n * (1 + W);
        for (unsigned long long c = threadIdx.x; c < cells; c += blockDim.x) {
            const unsigned long long i = c / (1 + W);
            const uint32_t f = (uint32_t)(c - i * (1 + W));
            dst[d0 * (1 + W) + c] = f == 0 ? tmp_keys[src + i] : tmp_words[(unsigned long long)(f - 1) * tmp_cap + src + i];
        }
    }
}

// compact per-bucket filtered records (mode 2 leaves them at the bucket's old offset) into new offsets
__global__ void k_compact_records(const unsigned long long* __restrict__ src, const unsigned long long* __restrict__ old_off,
                                  const unsigned long long* __restrict__ new_off, uint32_t B,
                                  unsigned long long* __restrict__ dst) {
    for (uint32_t b = blockIdx.x; b < B; b += gridDim.x) {
        const unsigned long long s = old_off[b], d = new_off[b], n = new_off[b + 1] - d;
        for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) dst[d + i] = src[s + i];
    }
}

// ------------------------------------------------------------------------------------------
// order: LSD radix sort of (k-mer, column index), one warp per contiguous segment
// ------------------------------------------------------------------------------------------
constexpr int kSortSeg = 4096;        // items per warp segment
constexpr int kSortWarps = 8;

__global__ void __launch_bounds__(kSortWarps * 32)
k_sort_hist(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t shift, uint32_t n_seg,
            uint32_t* __restrict__ hist /* [256][n_seg] */) {
    __shared__ uint32_t s_h[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t seg = blockIdx.x * kSortWarps + warp;
    for (int i = lane; i < 256; i += 32) s_h[warp][i] = 0;
    __syncwarp();
    if (seg < n_seg) {
        const uint64_t beg = (uint64_t)seg * kSortSeg, end = min(n, beg + kSortSeg);
        for (uint64_t i = beg + lane; i < end; i += 32) atomicAdd(&s_h[warp][(keys[i] >> shift) & 255u], 1u);
        __syncwarp();
        for (int i = lane; i < 256; i += 32) hist[(uint64_t)i * n_seg + seg] = s_h[warp][i];
    }
}

// exclusive scan of a u32 array (length n, multiple of 4) in place, three launches:
//   k_scan_u32_partial: per-4096-chunk totals; k_scan_u32_mid: scan of the totals (one block);
//   k_scan_u32_final: chunk-local scan + chunk offset.  All accesses are 16-byte coalesced.
constexpr int kScanChunk = 4096;

__global__ void __launch_bounds__(1024)
k_scan_u32_partial(const uint32_t* __restrict__ a, uint64_t n, uint32_t* __restrict__ partial) {
    __shared__ uint32_t s_warp[33];
    const uint64_t i = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * 4;
    uint32_t v = 0;
    if (i < n) { const uint4 q = *reinterpret_cast<const uint4*>(a + i); v = q.x + q.y + q.z + q.w; }
    uint32_t total;
    block_excl_scan_1024(v, s_warp, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
k_scan_u32_mid(uint32_t* __restrict__ partial, uint32_t nb) {
    __shared__ uint32_t s_warp[33];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < nb ? partial[i] : 0;
        uint32_t total;
        const uint32_t ex = block_excl_scan_1024(v, s_warp, total);
        const uint32_t carry = s_carry;
        if (i < nb) partial[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024)
k_scan_u32_final(uint32_t* __restrict__ a, uint64_t n, const uint32_t* __restrict__ partial) {
    __shared__ uint32_t s_warp[33];
    const uint64_t i = (uint64_t)blockIdx.x * kScanChunk + (uint64_t)threadIdx.x * 4;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (i < n) q = *reinterpret_cast<const uint4*>(a + i);
    uint32_t total;
    uint32_t ex = block_excl_scan_1024(q.x + q.y + q.z + q.w, s_warp, total) + partial[blockIdx.x];
    if (i < n) {
        uint4 o;
        o.x = ex; o.y = ex + q.x; o.z = o.y + q.y; o.w = o.z + q.z;
        *reinterpret_cast<uint4*>(a + i) = o;
    }
}

__global__ void __launch_bounds__(kSortWarps * 32)
k_sort_scatter(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ idx_in, uint64_t n,
               uint32_t shift, uint32_t n_seg, const uint32_t* __restrict__ hist,
               unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ idx_out) {
    __shared__ uint32_t s_b[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t seg = blockIdx.x * kSortWarps + warp;
    if (seg >= n_seg) return;
    for (int i = lane; i < 256; i += 32) s_b[warp][i] = hist[(uint64_t)i * n_seg + seg];
    __syncwarp();
    const uint64_t beg = (uint64_t)seg * kSortSeg, end = min(n, beg + kSortSeg);
    for (uint64_t i0 = beg; i0 < end; i0 += 32) {
        const uint64_t i = i0 + lane;
        const bool act = i < end;
        const uint32_t amask = __ballot_sync(0xffffffffu, act);
        if (act) {
            const unsigned long long key = keys_in[i];
            const uint32_t idx = idx_in ? idx_in[i] : (uint32_t)i;
            const uint32_t d = (uint32_t)(key >> shift) & 255u;
            const uint32_t peers = __match_any_sync(amask, d);
            const uint32_t rank = __popc(peers & lanemask_lt());
            const uint32_t pos = s_b[warp][d] + rank;
            __syncwarp(amask);
            if (rank == 0) s_b[warp][d] += __popc(peers);
            __syncwarp(amask);
            keys_out[pos] = key;
            idx_out[pos] = idx;
        }
    }
}

// ---- order, fast path: one MSD partition on the top bits of the k-mer, then a per-partition bitonic
// sort in shared memory fused with the column gather (2 passes over U instead of 8 radix passes).
// Canonical k-mers are denser near 0 (density 2(1-x)), so partitions hold up to ~2x the average;
// the host sizes the partition count for a 4x margin and falls back to the radix sort on overflow.
constexpr int kLocalSortCap = 16384;
constexpr int kLocalSortThreads = 1024;

__global__ void __launch_bounds__(256)
k_msd_count(const unsigned long long* __restrict__ keys, uint64_t U, uint32_t shift, uint32_t P,
            unsigned long long* __restrict__ hist) {
    extern __shared__ uint32_t s_h[];
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < U; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&s_h[(uint32_t)(keys[i] >> shift)], 1u);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
        if (s_h[i]) atomicAdd(&hist[i], (unsigned long long)s_h[i]);
}

__global__ void __launch_bounds__(256)
k_msd_scatter(const unsigned long long* __restrict__ keys, uint64_t U, uint32_t shift,
              unsigned long long* __restrict__ cursors, unsigned long long* __restrict__ out_keys,
              uint32_t* __restrict__ out_idx) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= U) return;
    const unsigned long long key = keys[i];
    const unsigned long long o = atomicAdd(&cursors[shift >= 64 ? 0u : (uint32_t)(key >> shift)], 1ULL);
    out_keys[o] = key;
    out_idx[o] = (uint32_t)i;
}

__global__ void __launch_bounds__(kLocalSortThreads, 1)
k_local_sort_gather(const unsigned long long* __restrict__ pkeys, const uint32_t* __restrict__ pidx,
                    const unsigned long long* __restrict__ offsets, uint32_t P, uint64_t U, uint32_t W,
                    const unsigned long long* __restrict__ uwords, uint64_t ucap,
                    unsigned long long* __restrict__ kmers, unsigned long long* __restrict__ matrix,
                    unsigned long long* __restrict__ scalars) {
    extern __shared__ unsigned long long s_key[];          // [cap] keys, then [cap] u32 indices
    for (uint32_t part = blockIdx.x; part < P; part += gridDim.x) {
        const unsigned long long beg = offsets[part], end = offsets[part + 1];
        const uint32_t n = (uint32_t)(end - beg);
        if (n == 0) continue;
        if (n > (uint32_t)kLocalSortCap) { if (threadIdx.x == 0) scalars[S_WORK] = 1; continue; }
        uint32_t npad = 1;
        while (npad < n) npad <<= 1;
        uint32_t* s_idx = reinterpret_cast<uint32_t*>(s_key + npad);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < npad; i += blockDim.x) {
            s_key[i] = i < n ? pkeys[beg + i] : ~0ULL;
            s_idx[i] = i < n ? pidx[beg + i] : 0u;
        }
        for (uint32_t size = 2; size <= npad; size <<= 1) {
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (uint32_t t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
                    const uint32_t i = 2 * t - (t & (stride - 1));
                    const uint32_t j = i + stride;
                    const unsigned long long a = s_key[i], b = s_key[j];
                    const bool asc = (i & size) == 0;
                    if ((a > b) == asc) {
                        s_key[i] = b; s_key[j] = a;
                        const uint32_t x = s_idx[i]; s_idx[i] = s_idx[j]; s_idx[j] = x;
                    }
                }
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            kmers[beg + i] = s_key[i];
            const uint32_t src = s_idx[i];
            for (uint32_t w = 0; w < W; ++w) matrix[(uint64_t)w * U + beg + i] = uwords[(uint64_t)w * ucap + src];
        }
    }
}

// columns in final order: kmers[j], matrix[w][j] = uwords[w][idx[j]]
__global__ void k_gather(const unsigned long long* __restrict__ sorted_keys, const uint32_t* __restrict__ idx,
                         uint64_t U, uint32_t W, const unsigned long long* __restrict__ uwords, uint64_t ucap,
                         unsigned long long* __restrict__ kmers, unsigned long long* __restrict__ matrix) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= U) return;
    kmers[j] = sorted_keys[j];
    const uint32_t src = idx[j];
    for (uint32_t w = 0; w < W; ++w) matrix[(uint64_t)w * U + j] = uwords[(uint64_t)w * ucap + src];
}

// ------------------------------------------------------------------------------------------
// emit
// ------------------------------------------------------------------------------------------
__global__ void k_kmer_strings(const unsigned long long* __restrict__ kmers, uint64_t j0, uint64_t n_bytes,
                               uint32_t k, char* __restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bytes) return;
    const uint64_t j = i / k; const uint32_t c = (uint32_t)(i % k);
    const uint32_t code = (uint32_t)(kmers[j0 + j] >> (2 * (k - 1 - c))) & 3u;
    dst[i] = "ACTG"[code];
}

// rows [j0, j0 + n_rows) of the TSV body; row width = k + 2G + 1
__global__ void k_format_tsv(const unsigned long long* __restrict__ kmers, const unsigned long long* __restrict__ matrix,
                             uint64_t U, uint32_t G, uint32_t k, uint64_t j0, uint64_t n_bytes, char* __restrict__ dst) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bytes) return;
    const uint32_t roww = k + 2 * G + 1;
    const uint64_t j = j0 + i / roww; const uint32_t c = (uint32_t)(i % roww);
    char out;
    if (c < k) out = "ACTG"[(uint32_t)(kmers[j] >> (2 * (k - 1 - c))) & 3u];
    else if (c == roww - 1) out = '\n';
    else if (((c - k) & 1u) == 0) out = '\t';
    else {
        const uint32_t g = (c - k) >> 1;
        out = ((matrix[(uint64_t)(g >> 6) * U + j] >> (63 - (g & 63))) & 1ULL) ? '1' : '0';
    }
    dst[i] = out;
}

}  // namespace grmkm
