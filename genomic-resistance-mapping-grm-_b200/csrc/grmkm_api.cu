// grmkm_api.cu -- C ABI (include/grmkm.h) and host orchestration of the sm_100a kernels.
// One context = one GPU.  No CPU fallback: without a device grmkm_create fails.
#include "../../include/grmkm.h"
#include "grmkm_kernels.cuh"
#include "grmkm_units.cuh"
#include "grmkm_synth.cuh"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

using namespace grmkm;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct Input {
    uint32_t row = 0;
    uint32_t kind = 0;
    const uint8_t* host = nullptr;
    const uint8_t* dev = nullptr;
    uint64_t len = 0;
    std::vector<uint8_t> owned;
};

constexpr uint64_t kBatchBytes = 32ull << 20;      // text per pipelined H2D batch (GRMKM_BATCH_BYTES overrides, for tests)

enum Stage { T_START = 0, T_H2D, T_PARSE, T_PACK, T_COUNT, T_BOUNDS, T_SCATTER, T_ABUND, T_DEDUPE, T_EXPAND, T_AGG, T_SORT, T_N };

}  // namespace

struct grmkm_ctx {
    grmkm_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    std::vector<Input> inputs;
    uint32_t n_genomes_decl = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    void* h_tab = nullptr;         // page-locked arena for the per-batch host tables (files, stream starts, ticket order)
    size_t h_tab_cap = 0, h_tab_used = 0;
    uint64_t wide_hint = 0;
    // the packed stream and the tile words of the NEXT build are cleared on a side stream while this build's dedupe /
    // expand / aggregate run (they are dead once the scatter has consumed them)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_scattered = nullptr, ev_cleared = nullptr;
    bool pre_pending = false;
    const void *pre_codes = nullptr, *pre_valid = nullptr, *pre_pub = nullptr;
    uint64_t pre_groups = 0, pre_pub_bytes = 0;
    uint64_t pass_groups = 0, pass_pub_bytes = 0;
    uint64_t wide_capped_need = 0;   // an expansion-buffer request of up to this many bytes was capped by free memory
    uint64_t ucap_hint = 0;        // columns of the previous build + headroom (sizes the aggregate's output)        // wide records of the previous build (sizes the expansion's bucket regions)

    // device buffers (grow-only, reused across builds)
    DevBuf in, files, hdr0, tile_file, tile_pub, tile_order, fss, codes, valid, hist,
        offsets, offsets2, bcounts, records, records2, ukeys, uwords, skeys, sidx_a, sidx_b, shist, kmers, matrix,
        scalars, fmt, synth, owner_start, refs, spart, stile_file, bbase, apub, masks, units, ucur, ubeg, wu, wide;
    size_t device_bytes = 0;

    cudaEvent_t ev[T_N]{};
    bool ev_ok = false;

    // H2D pipeline: copy stream, "batch copied" / "staging half free" events
    uint64_t batch_bytes = kBatchBytes;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2]{}, ev_free[2]{};

    // page-locked host copy of the result (grmkm_host_result)
    void* host_res = nullptr;
    size_t host_res_cap = 0;

    // result
    bool built = false;
    uint64_t U = 0;
    uint32_t W = 0, G = 0;
    grmkm_stats stats{};
    grmkm_times times{};

    // multi-GPU partial state
    uint32_t part_ranks = 0;
    std::vector<uint64_t> part_counts;
    uint64_t part_total = 0, part_cap = 0;      // partial columns of the last partial build; stride of its bucket chunks
    uint32_t part_buckets = 0;
    uint32_t part_words = 0;
    uint32_t cur_bucket_bits = 0;
};

namespace {

int fail(grmkm_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CU_TRY(c, call)                                                                                   \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail((c), GRMKM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));           \
    } while (0)

int ensure(grmkm_ctx* c, DevBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return GRMKM_OK;
    if (b.p) { cudaFree(b.p); c->device_bytes -= b.cap; b.p = nullptr; b.cap = 0; }
    size_t want = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return fail(c, GRMKM_E_NOMEM, "cudaMalloc of " + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
    }
    b.cap = want;
    c->device_bytes += want;
    return GRMKM_OK;
}
#define ENSURE(c, buf, bytes) do { int r_ = ensure((c), (buf), (bytes)); if (r_) return r_; } while (0)

void release(grmkm_ctx* c, DevBuf& b) {
    if (b.p) { cudaFree(b.p); c->device_bytes -= b.cap; }
    b.p = nullptr; b.cap = 0;
}

uint32_t ceil_log2(uint64_t x) {
    uint32_t b = 0;
    while ((1ULL << b) < x) ++b;
    return b;
}

struct Launches { uint32_t n = 0; };

int read_file(grmkm_ctx* c, const char* path, std::vector<uint8_t>& out) {
    const size_t L = strlen(path);
    const bool gz = L > 3 && strcmp(path + L - 3, ".gz") == 0;
    if (gz) {
        gzFile f = gzopen(path, "rb");
        if (!f) return fail(c, GRMKM_E_IO, std::string("cannot open ") + path);
        gzbuffer(f, 1 << 20);
        std::vector<uint8_t> buf(1 << 22);
        int n;
        while ((n = gzread(f, buf.data(), (unsigned)buf.size())) > 0) out.insert(out.end(), buf.begin(), buf.begin() + n);
        const bool bad = n < 0;
        gzclose(f);
        if (bad) return fail(c, GRMKM_E_IO, std::string("cannot inflate ") + path);
        return GRMKM_OK;
    }
    FILE* f = fopen(path, "rb");
    if (!f) return fail(c, GRMKM_E_IO, std::string("cannot open ") + path);
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return fail(c, GRMKM_E_IO, std::string("cannot size ") + path); }
    out.resize((size_t)sz);
    size_t got = sz ? fread(out.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) return fail(c, GRMKM_E_IO, std::string("short read on ") + path);
    return GRMKM_OK;
}

struct Plan {
    uint32_t F = 0, G = 0, W = 0;
    uint64_t n_tiles = 0, n_sblk = 0, max_stream = 0, n_groups_max = 0, in_bytes = 0;
    uint32_t bucket_bits = 0, row_bits = 0, slots = 0, sub_bits = 0;
    size_t agg_smem = 0;
};

size_t agg_smem_budget(const grmkm_ctx* c) {
    // leave room for the kernel's static shared memory
    size_t lim = c->smem_optin ? c->smem_optin : 227 * 1024;
    lim = (lim + 1024) / kAggCtasPerSm - 1024;       // every resident CTA also costs 1 KB of system shared memory
    return lim - 4096;
}

// table capacity for W words per column and the bucket count that keeps a bucket's distinct k-mers
// (estimated as 3x the largest genome) at about half of it
// table = (slots + kMaxProbe) x (u64 key + 2W u32 half-words + u8 kept flag / correction + u8 own inversions); slots = home positions
uint32_t table_slots(const grmkm_ctx* c, uint32_t W) {
    const size_t budget = agg_smem_budget(c);
    const size_t total = (budget - 8) / (10 + 8 * (size_t)W);
    if (total < (size_t)kMaxProbe + 256) return 0;
    return (uint32_t)std::min<size_t>(kAggMaxSlots, total - kMaxProbe);
}
size_t table_smem(uint32_t slots, uint32_t W) {
    return (((size_t)slots + kMaxProbe) * (10 + 8 * (size_t)W) + 8 + 15) & ~size_t(15);
}
// Bucket count: the distinct k-mers of a bucket must fit its shared-memory table at a load of about 0.6 by the estimate
// (which is generous: C2 ends at 0.53).
// The pan-genome of the context is estimated from the largest genome (1.9x its text; a bucket that turns out
// too full is split into key sub-ranges by the kernel, so the estimate only costs time, never correctness).
// Fewer buckets = longer runs per scatter tile = fewer store requests, the scatter's bound.
constexpr uint32_t kUnitMaxBucketBits = 11;      // the unit expansion sorts tiles over at most 2^11 hash buckets
bool units_wanted(const grmkm_ctx* c) {
    return c->cfg.min_abundance <= 1 && !(c->cfg.flags & (GRMKM_FLAG_KMER_RECORDS | GRMKM_FLAG_SIMPLE_SCATTER));
}
uint32_t auto_bucket_bits(const grmkm_ctx* c, uint32_t G) {
    std::vector<uint64_t> row_bytes(std::max(G, 1u), 0);
    for (const Input& in : c->inputs) if (in.row < G) row_bytes[in.row] += in.len;
    const uint64_t max_row = *std::max_element(row_bytes.begin(), row_bytes.end());
    const uint32_t slots = table_slots(c, (G + 63) / 64);
    const uint64_t u_est = max_row + max_row * 9 / 10 + 1024;
    const uint64_t per = std::max<uint64_t>(1, (uint64_t)slots * 60 / 100);
    uint32_t bits = std::min(15u, std::max(6u, ceil_log2((u_est + per - 1) / per)));
    if (!units_wanted(c)) return std::max(std::max(1u, ceil_log2(G)), bits);     // k-mer records carry the row below the hash
    // unit path: past 2^11 buckets every further bit doubles the aggregate's passes over the records (key sub-ranges),
    // so a table that the estimate fills to 0.7 is still the better deal
    if (bits > kUnitMaxBucketBits) {
        const uint64_t per7 = std::max<uint64_t>(1, (uint64_t)slots * 70 / 100);
        if (ceil_log2((u_est + per7 - 1) / per7) <= kUnitMaxBucketBits) bits = kUnitMaxBucketBits;
    }
    return bits;
}

// sort columns (ukeys/uwords, n items, stride ucap) by key into kmers/matrix
int sort_and_gather(grmkm_ctx* c, uint64_t U, uint32_t W, uint64_t ucap, uint32_t key_bits_total, bool keep_order,
                    Launches& L) {
    cudaStream_t st = c->stream;
    ENSURE(c, c->kmers, U * 8);
    ENSURE(c, c->matrix, (size_t)U * W * 8);
    if (U == 0) return GRMKM_OK;
    const uint32_t gblocks = (uint32_t)((U + 255) / 256);
    if (keep_order) {
        // identity order: reuse gather with idx = iota produced by a zero-pass "sort"
        ENSURE(c, c->sidx_a, U * 4);
        std::vector<uint32_t> iota;  // small helper path, only used with GRMKM_FLAG_HASH_ORDER
        iota.resize(U);
        for (uint64_t i = 0; i < U; ++i) iota[i] = (uint32_t)i;
        CU_TRY(c, cudaMemcpyAsync(c->sidx_a.p, iota.data(), U * 4, cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaStreamSynchronize(st));
        k_gather<<<gblocks, 256, 0, st>>>((const unsigned long long*)c->ukeys.p, (const uint32_t*)c->sidx_a.p, U, W,
                                          (const unsigned long long*)c->uwords.p, ucap,
                                          (unsigned long long*)c->kmers.p, (unsigned long long*)c->matrix.p);
        L.n++;
        CU_TRY(c, cudaGetLastError());
        return GRMKM_OK;
    }
    // ---- fast path: MSD partition + shared-memory sort fused with the gather
    ENSURE(c, c->skeys, U * 8);
    ENSURE(c, c->sidx_a, U * 4);
    if (U <= 0xFFFFFFFFULL && !(c->cfg.flags & GRMKM_FLAG_RADIX_ORDER)) {
        uint32_t pbits = std::min(key_bits_total, std::min(15u, ceil_log2((U + 4095) / 4096)));
        const uint32_t Pn = 1u << pbits;
        const uint32_t shift = key_bits_total - pbits;
        ENSURE(c, c->hist, (size_t)std::max<uint32_t>(Pn, 1) * 8 * kCursorStride);
        ENSURE(c, c->offsets2, (size_t)(Pn + 1) * 8);
        uint64_t* d_scalars = (uint64_t*)c->scalars.p;
        CU_TRY(c, cudaMemsetAsync(c->hist.p, 0, (size_t)Pn * 8, st));
        CU_TRY(c, cudaMemsetAsync(d_scalars + S_WORK, 0, 8, st));
        const uint32_t cgrid = (uint32_t)std::min<uint64_t>((U + 255) / 256, (uint64_t)c->sm_count * 8);
        if (pbits == 0) {
            const unsigned long long uu = U;   // one partition holds everything
            CU_TRY(c, cudaMemcpyAsync(c->hist.p, &uu, 8, cudaMemcpyHostToDevice, st));
        } else {
            CU_TRY(c, cudaFuncSetAttribute(k_msd_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Pn * 4)));
            k_msd_count<<<cgrid, 256, (size_t)Pn * 4, st>>>((const unsigned long long*)c->ukeys.p, U, shift, Pn,
                                                           (unsigned long long*)c->hist.p);
        }
        k_bucket_offsets<<<1, 1024, 0, st>>>((unsigned long long*)c->hist.p, (unsigned long long*)c->offsets2.p, Pn,
                                             d_scalars, S_N_SOLID, 1);
        k_msd_scatter<<<gblocks, 256, 0, st>>>((const unsigned long long*)c->ukeys.p, U, pbits ? shift : 64,
                                               (unsigned long long*)c->hist.p, (unsigned long long*)c->skeys.p,
                                               (uint32_t*)c->sidx_a.p);
        const size_t lsm = (size_t)kLocalSortCap * 12;
        CU_TRY(c, cudaFuncSetAttribute(k_local_sort_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lsm));
        k_local_sort_gather<<<std::min<uint32_t>(Pn, (uint32_t)c->sm_count), kLocalSortThreads, lsm, st>>>(
            (const unsigned long long*)c->skeys.p, (const uint32_t*)c->sidx_a.p, (const unsigned long long*)c->offsets2.p,
            Pn, U, W, (const unsigned long long*)c->uwords.p, ucap, (unsigned long long*)c->kmers.p,
            (unsigned long long*)c->matrix.p, (unsigned long long*)d_scalars);
        L.n += 4;
        CU_TRY(c, cudaGetLastError());
        uint64_t overflow = 0;
        CU_TRY(c, cudaMemcpyAsync(&overflow, d_scalars + S_WORK, 8, cudaMemcpyDeviceToHost, st));
        CU_TRY(c, cudaStreamSynchronize(st));
        if (!overflow) return GRMKM_OK;
        // a partition did not fit shared memory (heavily skewed k-mer prefixes): radix sort below
    }
    const uint32_t n_seg = (uint32_t)((U + kSortSeg - 1) / kSortSeg);
    const uint32_t sblocks = (n_seg + kSortWarps - 1) / kSortWarps;
    ENSURE(c, c->sidx_b, U * 4);
    ENSURE(c, c->shist, (size_t)256 * n_seg * 4);
    const uint64_t hist_n = (uint64_t)256 * n_seg;
    const uint32_t n_chunks = (uint32_t)((hist_n + kScanChunk - 1) / kScanChunk);
    ENSURE(c, c->spart, (size_t)n_chunks * 4);
    const uint32_t passes = (key_bits_total + 7) / 8;
    unsigned long long* ka = (unsigned long long*)c->ukeys.p;
    unsigned long long* kb = (unsigned long long*)c->skeys.p;
    uint32_t* ia = nullptr;
    uint32_t* ib = (uint32_t*)c->sidx_a.p;
    uint32_t* ispare = (uint32_t*)c->sidx_b.p;
    for (uint32_t pass = 0; pass < passes; ++pass) {
        const uint32_t shift = pass * 8;
        k_sort_hist<<<sblocks, kSortWarps * 32, 0, st>>>(ka, U, shift, n_seg, (uint32_t*)c->shist.p);
        k_scan_u32_partial<<<n_chunks, 1024, 0, st>>>((const uint32_t*)c->shist.p, hist_n, (uint32_t*)c->spart.p);
        k_scan_u32_mid<<<1, 1024, 0, st>>>((uint32_t*)c->spart.p, n_chunks);
        k_scan_u32_final<<<n_chunks, 1024, 0, st>>>((uint32_t*)c->shist.p, hist_n, (const uint32_t*)c->spart.p);
        k_sort_scatter<<<sblocks, kSortWarps * 32, 0, st>>>(ka, ia, U, shift, n_seg, (const uint32_t*)c->shist.p, kb, ib);
        L.n += 5;
        std::swap(ka, kb);
        uint32_t* t = ia ? ia : ispare;
        ia = ib; ib = t;
    }
    CU_TRY(c, cudaGetLastError());
    k_gather<<<gblocks, 256, 0, st>>>(ka, ia, U, W, (const unsigned long long*)c->uwords.p, ucap,
                                      (unsigned long long*)c->kmers.p, (unsigned long long*)c->matrix.p);
    L.n++;
    CU_TRY(c, cudaGetLastError());
    return GRMKM_OK;
}

int check_ctx(const grmkm_ctx* c) { return c ? GRMKM_OK : GRMKM_E_INVALID; }

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int grmkm_abi_version(void) { return GRMKM_ABI_VERSION; }

int grmkm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int grmkm_create(const grmkm_config* cfg, grmkm_ctx** out) {
    if (!cfg || !out) return fail(nullptr, GRMKM_E_INVALID, "null argument");
    *out = nullptr;
    grmkm_config c{};
    memcpy(&c, cfg, std::min<size_t>(sizeof c, cfg->struct_size ? cfg->struct_size : sizeof c));
    if (c.k < 1 || c.k > 32)
        return fail(nullptr, GRMKM_E_UNSUPPORTED_K, "k must be in 1..32 (got " + std::to_string(c.k) + ")");
    if (c.min_abundance == 0) c.min_abundance = 1;
    if (c.input_kind > GRMKM_FASTQ) return fail(nullptr, GRMKM_E_INVALID, "input_kind must be GRMKM_FASTA or GRMKM_FASTQ");
    if (c.bucket_bits && (c.bucket_bits < 4 || c.bucket_bits > 15))
        return fail(nullptr, GRMKM_E_INVALID, "bucket_bits must be 0 (auto) or 4..15");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, GRMKM_E_NO_DEVICE,
                    std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") +
                        "); libgrmkm has no CPU fallback");
    }
    int dev = c.device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    if (dev >= ndev) return fail(nullptr, GRMKM_E_INVALID, "device ordinal out of range");
    if ((e = cudaSetDevice(dev)) != cudaSuccess)
        return fail(nullptr, GRMKM_E_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    grmkm_ctx* x = new (std::nothrow) grmkm_ctx();
    if (!x) return fail(nullptr, GRMKM_E_NOMEM, "out of host memory");
    x->cfg = c;
    x->device = dev;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
        x->sm_count = prop.multiProcessorCount;
        x->smem_optin = prop.sharedMemPerBlockOptin;
    }
    if (const char* bb = getenv("GRMKM_BATCH_BYTES")) { const long long v = atoll(bb); if (v > 0) x->batch_bytes = (uint64_t)v; }
    if (c.stream) x->stream = (cudaStream_t)c.stream;
    else {
        if ((e = cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            delete x;
            return fail(nullptr, GRMKM_E_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
        }
        x->own_stream = true;
    }
    for (int i = 0; i < T_N; ++i) {
        if (cudaEventCreate(&x->ev[i]) != cudaSuccess) { x->ev_ok = false; break; }
        x->ev_ok = true;
    }
    *out = x;
    return GRMKM_OK;
}

void grmkm_destroy(grmkm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);      // the clear for a next build may still be running
    DevBuf* all[] = {&c->in, &c->files, &c->hdr0, &c->tile_file, &c->tile_pub, &c->tile_order,
                     &c->fss, &c->codes, &c->valid, &c->hist, &c->offsets, &c->offsets2,
                     &c->bcounts, &c->records, &c->records2, &c->ukeys, &c->uwords, &c->skeys, &c->sidx_a, &c->sidx_b,
                     &c->shist, &c->kmers, &c->matrix, &c->scalars, &c->fmt, &c->synth, &c->owner_start, &c->refs, &c->spart, &c->stile_file, &c->bbase, &c->apub,
                     &c->masks, &c->units, &c->ucur, &c->ubeg, &c->wu, &c->wide};
    for (DevBuf* b : all) release(c, *b);
    if (c->ev_ok) for (int i = 0; i < T_N; ++i) cudaEventDestroy(c->ev[i]);
    if (c->host_res) cudaFreeHost(c->host_res);
    if (c->aux_stream) {
        cudaEventDestroy(c->ev_scattered); cudaEventDestroy(c->ev_cleared);
        cudaStreamDestroy(c->aux_stream);
    }
    if (c->copy_stream) {
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(c->ev_copied[i]); cudaEventDestroy(c->ev_free[i]); }
        cudaStreamDestroy(c->copy_stream);
    }
    if (c->h_tab) cudaFreeHost(c->h_tab);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* grmkm_last_error(const grmkm_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int grmkm_reset(grmkm_ctx* c) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    c->inputs.clear();
    c->n_genomes_decl = 0;
    c->built = false;
    c->U = 0; c->W = 0; c->G = 0;
    c->part_ranks = 0;
    return GRMKM_OK;
}

int grmkm_add_genome_bytes(grmkm_ctx* c, uint32_t row, const uint8_t* data, uint64_t n) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!data && n) return fail(c, GRMKM_E_INVALID, "null data");
    Input in; in.row = row; in.kind = c->cfg.input_kind; in.host = data; in.len = n;
    c->inputs.push_back(std::move(in));
    c->built = false;
    return GRMKM_OK;
}

int grmkm_add_genome_device(grmkm_ctx* c, uint32_t row, const void* dev_data, uint64_t n) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!dev_data && n) return fail(c, GRMKM_E_INVALID, "null data");
    if ((uintptr_t)dev_data & 15) return fail(c, GRMKM_E_INVALID, "device input must be 16-byte aligned");
    Input in; in.row = row; in.kind = c->cfg.input_kind; in.dev = (const uint8_t*)dev_data; in.len = n;
    c->inputs.push_back(std::move(in));
    c->built = false;
    return GRMKM_OK;
}

int grmkm_add_genomes(grmkm_ctx* c, uint32_t n, const uint32_t* rows, const void* const* data, const uint64_t* lens,
                      int on_device) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (n && (!rows || !data || !lens)) return fail(c, GRMKM_E_INVALID, "null input list");
    c->inputs.reserve(c->inputs.size() + n);
    for (uint32_t i = 0; i < n; ++i) {
        const int r = on_device ? grmkm_add_genome_device(c, rows[i], data[i], lens[i])
                                : grmkm_add_genome_bytes(c, rows[i], (const uint8_t*)data[i], lens[i]);
        if (r) return r;
    }
    return GRMKM_OK;
}

int grmkm_add_genome_files(grmkm_ctx* c, uint32_t row, const char* const* paths, int n_paths) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (n_paths < 0 || (!paths && n_paths)) return fail(c, GRMKM_E_INVALID, "bad path list");
    for (int i = 0; i < n_paths; ++i) {
        Input in; in.row = row; in.kind = c->cfg.input_kind;
        int r = read_file(c, paths[i], in.owned);
        if (r) return r;
        in.len = in.owned.size();
        c->inputs.push_back(std::move(in));
        c->inputs.back().host = c->inputs.back().owned.data();
    }
    c->built = false;
    return GRMKM_OK;
}

int grmkm_set_genome_count(grmkm_ctx* c, uint32_t n) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    c->n_genomes_decl = n;
    return GRMKM_OK;
}

// ------------------------------------------------------------------------------------------------
// the build
// ------------------------------------------------------------------------------------------------
static int build_impl(grmkm_ctx* c, uint32_t mode /*0 final, 1 partial*/, uint32_t n_ranges) {
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    Launches L;
    c->built = false;
    c->stats = grmkm_stats{};
    c->times = grmkm_times{};

    Plan P;
    P.F = (uint32_t)c->inputs.size();
    uint32_t maxrow = 0;
    for (const Input& in : c->inputs) maxrow = std::max(maxrow, in.row + 1);
    P.G = std::max(maxrow, c->n_genomes_decl);
    P.W = (P.G + 63) / 64;
    c->G = P.G; c->W = P.W; c->U = 0;
    c->part_ranks = 0; c->part_total = 0; c->part_words = P.W;
    if (P.G == 0 || P.F == 0) {
        c->built = (mode == 0);
        c->stats.n_genomes = P.G; c->stats.n_words = P.W;
        if (mode == 1) { c->part_ranks = n_ranges; c->part_counts.assign(n_ranges, 0); }
        return GRMKM_OK;
    }
    P.row_bits = std::max(1u, ceil_log2(P.G));
    if (P.row_bits > 15) return fail(c, GRMKM_E_UNSUPPORTED, "more than 32768 genomes in one context");

    // ---- batches.  Host inputs are staged batch by batch on a copy stream (two staging halves), so that the
    // H2D copy of batch i+1 overlaps parse / pack / scatter of batch i; the scatter appends to the bucket
    // regions, so batches need no merge.  Device-resident inputs (and the exact-offset fallback, which needs
    // a count over everything first) run as one batch.
    struct Batch { uint32_t f0, f1; uint64_t bytes, staged, tiles, stream; };
    // stream entries reserved for a file: every byte yields at most one entry; files start on group boundaries
    auto stream_cap = [](uint64_t len) { return ((len + 31) & ~31ULL) + 32; };
    bool any_host = false;
    for (const Input& in : c->inputs) { P.max_stream += stream_cap(in.len); P.in_bytes += in.len; any_host = any_host || !in.dev; }
    auto make_batches = [&](bool pipelined) {
        std::vector<Batch> v;
        const uint64_t target = pipelined ? c->batch_bytes : ~0ULL;
        Batch cur{0, 0, 0, 0, 0, 0};
        for (uint32_t f = 0; f < P.F; ++f) {
            const Input& in = c->inputs[f];
            if (cur.f1 > cur.f0 && cur.bytes + in.len > target) { v.push_back(cur); cur = Batch{f, f, 0, 0, 0, 0}; }
            cur.f1 = f + 1; cur.bytes += in.len; cur.stream += stream_cap(in.len);
            if (!in.dev) cur.staged += (in.len + 15) & ~15ULL;
            cur.tiles += std::max<uint64_t>(1, (in.len + kTileBytes - 1) / kTileBytes);
        }
        v.push_back(cur);
        return v;
    };

    // ---- aggregate table geometry and bucket count
    const uint32_t Wtab = P.W;
    const size_t budget = agg_smem_budget(c);
    P.slots = table_slots(c, Wtab);
    if (P.slots < 256) return fail(c, GRMKM_E_UNSUPPORTED, "too many genomes for the shared-memory column table");
    P.agg_smem = table_smem(P.slots, Wtab);
    // the unit path (below) carries a genome GROUP in its records, not a row, and its expansion sorts tiles over at most
    // 2^11 buckets: more distinct k-mers than 2^11 tables hold are handled as key sub-ranges of the buckets (virtual
    // buckets: one aggregate pass each over the bucket's records, which stay in L2), and beyond 2^14 by the table's own
    // overflow split -- instead of falling off to the per-record scatter
    const bool units_ok = units_wanted(c);
    P.bucket_bits = c->cfg.bucket_bits ? c->cfg.bucket_bits : auto_bucket_bits(c, P.G);
    if (!units_ok) P.bucket_bits = std::max(P.bucket_bits, P.row_bits);
    if (P.bucket_bits > 15) return fail(c, GRMKM_E_UNSUPPORTED, "bucket_bits > 15");
    if (const char* sbv = getenv("GRMKM_SUB_BITS")) P.sub_bits = (uint32_t)std::min(3, std::max(0, atoi(sbv)));
    if (units_ok && P.bucket_bits > kUnitMaxBucketBits) {
        P.sub_bits = std::max(P.sub_bits, std::min(3u, P.bucket_bits - kUnitMaxBucketBits));
        P.bucket_bits = kUnitMaxBucketBits;
    }
    c->cur_bucket_bits = P.bucket_bits;
    const uint32_t B = 1u << P.bucket_bits;

    ENSURE(c, c->scalars, S_COUNT * 8);
    ENSURE(c, c->hist, (size_t)B * 8 * kCursorStride);
    ENSURE(c, c->offsets, (size_t)(B + 1) * 8);
    uint64_t* d_scalars = (uint64_t*)c->scalars.p;
    const FileDesc* d_files = nullptr;

    if (c->ev_ok) cudaEventRecord(c->ev[T_START], st);
    CU_TRY(c, cudaMemsetAsync(c->scalars.p, 0, S_COUNT * 8, st));
    uint64_t h2d = 0;


    // parse + pack of one batch: leaves codes / valid / fss / S_STREAM_LEN of the batch
    auto front = [&](const Batch& bt, const uint8_t* staged_base, bool timed) -> int {
        const uint32_t F = bt.f1 - bt.f0;
        const uint64_t n_tiles = bt.tiles;
        const uint64_t n_groups_max = bt.stream / 32 + 2;
        if (n_tiles > 0x7fffffffULL) return fail(c, GRMKM_E_UNSUPPORTED, "input too large for one build (tile count)");
        ENSURE(c, c->files, F * sizeof(FileDesc));
        ENSURE(c, c->hdr0, F * 8);
        ENSURE(c, c->fss, (F + 1) * 8);
        ENSURE(c, c->tile_file, n_tiles * sizeof(TileTicket));
        ENSURE(c, c->tile_pub, (3 * n_tiles + 1) * 8);    // published tile summaries / states (look-back) + the ticket counter
        ENSURE(c, c->codes, n_groups_max * 8);
        ENSURE(c, c->valid, n_groups_max * 4);
        ENSURE(c, c->tile_order, n_tiles * 4);
        // the memsets go first: the device clears while the host builds this batch's tables -- unless the previous build
        // already cleared the buffers on the side stream (same buffers, at least this much of them)
        bool cleared = false;
        if (c->pre_pending) {
            CU_TRY(c, cudaStreamWaitEvent(st, c->ev_cleared, 0));
            cleared = c->pre_codes == c->codes.p && c->pre_valid == c->valid.p && c->pre_pub == c->tile_pub.p &&
                      c->pre_groups >= n_groups_max && c->pre_pub_bytes >= (3 * n_tiles + 1) * 8;
            c->pre_pending = false;
        }
        if (!cleared) {
            CU_TRY(c, cudaMemsetAsync(c->codes.p, 0, n_groups_max * 8, st));
            CU_TRY(c, cudaMemsetAsync(c->valid.p, 0, n_groups_max * 4, st));
            CU_TRY(c, cudaMemsetAsync(c->tile_pub.p, 0, (3 * n_tiles + 1) * 8, st));
        }
        c->pass_groups = std::max<uint64_t>(c->pass_groups, n_groups_max);
        c->pass_pub_bytes = std::max<uint64_t>(c->pass_pub_bytes, (3 * n_tiles + 1) * 8);
        // host tables in page-locked memory (asynchronous copies without a staging hop); they stay untouched until the
        // build's final synchronize
        const size_t tab_bytes = (F * sizeof(FileDesc) + (F + 1) * 8 + (size_t)n_tiles * 4 + 63) & ~size_t(63);
        if (c->h_tab_used + tab_bytes > c->h_tab_cap) return fail(c, GRMKM_E_NOMEM, "host table arena too small");
        uint8_t* tab = (uint8_t*)c->h_tab + c->h_tab_used;
        c->h_tab_used += tab_bytes;
        FileDesc* fds = (FileDesc*)tab;
        uint64_t* fss = (uint64_t*)(tab + F * sizeof(FileDesc));      // where every file's entries start in the stream
        uint32_t* order = (uint32_t*)(fss + F + 1);                   // ticket -> tile: round-robin over the files
        uint64_t tiles = 0, soff = 0, spos = 0;
        std::vector<std::pair<uint64_t, uint32_t>> by_tiles(F);       // (tile count, file), most tiles first
        for (uint32_t i = 0; i < F; ++i) {
            const Input& in = c->inputs[bt.f0 + i];
            fds[i].len = in.len; fds[i].row = in.row; fds[i].kind = in.kind; fds[i].tile_begin = tiles;
            const uint64_t nt = std::max<uint64_t>(1, (in.len + kTileBytes - 1) / kTileBytes);
            by_tiles[i] = {nt, i};
            tiles += nt;
            fss[i] = spos; spos += stream_cap(in.len);
            if (in.dev) fds[i].ptr = in.dev;
            else { fds[i].ptr = staged_base + soff; soff += (in.len + 15) & ~15ULL; }
        }
        fss[F] = spos;
        std::sort(by_tiles.begin(), by_tiles.end(), [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) {
            return a.first != b.first ? a.first > b.first : a.second < b.second; });
        {
            size_t o = 0, live = F;                                    // files that still have a tile number `l`
            for (uint64_t l = 0; o < n_tiles; ++l) {
                while (live > 0 && by_tiles[live - 1].first <= l) --live;
                for (size_t q = 0; q < live; ++q) order[o++] = (uint32_t)(fds[by_tiles[q].second].tile_begin + l);
            }
        }
        d_files = (const FileDesc*)c->files.p;
        CU_TRY(c, cudaMemcpyAsync(c->files.p, fds, F * sizeof(FileDesc), cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(c->fss.p, fss, (F + 1) * 8, cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaMemcpyAsync(c->tile_order.p, order, n_tiles * 4, cudaMemcpyHostToDevice, st));
        if (timed && c->ev_ok) cudaEventRecord(c->ev[T_H2D], st);
        k_first_header<<<(F * 32 + 255) / 256, 256, 0, st>>>(d_files, F, (uint64_t*)c->hdr0.p);
        k_tile_tickets<<<(uint32_t)((n_tiles + 255) / 256), 256, 0, st>>>(d_files, F, (const uint64_t*)c->hdr0.p, (const uint64_t*)c->fss.p,
                                                                           (const uint32_t*)c->tile_order.p, n_tiles, (TileTicket*)c->tile_file.p);
        L.n += 2;
        CU_TRY(c, cudaGetLastError());
        if (timed && c->ev_ok) cudaEventRecord(c->ev[T_PARSE], st);
        // single pass: summaries are published and resolved by look-back inside k_pack (grmkm_kernels.cuh)
        PackParams pp{};
        pp.files = d_files; pp.hdr0 = (const uint64_t*)c->hdr0.p; pp.n_tiles = n_tiles; pp.n_files = F;
        pp.tickets = (const TileTicket*)c->tile_file.p; pp.ticket = (uint32_t*)((unsigned long long*)c->tile_pub.p + 3 * n_tiles);
        pp.pub_a0 = (unsigned long long*)c->tile_pub.p; pp.pub_a1 = pp.pub_a0 + n_tiles; pp.pub_ps = pp.pub_a0 + 2 * n_tiles;
        pp.stream_len = bt.stream;
        pp.codes = (unsigned long long*)c->codes.p; pp.valid = (uint32_t*)c->valid.p; pp.scalars = d_scalars;
        if (c->cfg.input_kind == GRMKM_FASTA) k_pack<0><<<(uint32_t)n_tiles, kParseThreads, 0, st>>>(pp);
        else k_pack<1><<<(uint32_t)n_tiles, kParseThreads, 0, st>>>(pp);
        L.n++;
        CU_TRY(c, cudaGetLastError());
        if (timed && c->ev_ok) cudaEventRecord(c->ev[T_PACK], st);
        return GRMKM_OK;
    };

    // ---- extract + scatter, (abundance), aggregate.  Pass 0 scatters into over-provisioned bucket regions
    // without a count pass; if a region overflows (heavily skewed k-mer spectrum) pass 1 redoes the
    // scatter with exact offsets from a count pass.
    const bool staged = B <= (uint32_t)kStMaxBuckets && !(c->cfg.flags & GRMKM_FLAG_SIMPLE_SCATTER);
    // ---- unit (super-k-mer) path: contigs and reads without an abundance filter (grmkm_units.cuh)
    const bool use_units = staged && c->cfg.min_abundance <= 1 && !(c->cfg.flags & GRMKM_FLAG_KMER_RECORDS);
    const UnitGeom ug = unit_geom(c->cfg.k);
    const uint32_t WB = unit_words_per_entry(P.W);            // presence words per dedupe entry / wide record
    const uint32_t wbits = std::max(1u, ceil_log2((P.W + WB - 1) / WB));
    const uint32_t ES = 2 + WB, RS = unit_record_stride(WB);
    uint32_t MB = (uint32_t)c->sm_count;
    if (use_units) {
        // distinct (unit, 64-genome block) entries of a bucket should fit the dedupe table about once
        std::vector<uint64_t> row_bytes(std::max(P.G, 1u), 0);
        for (const Input& in : c->inputs) if (in.row < P.G) row_bytes[in.row] += in.len;
        const uint64_t max_row = *std::max_element(row_bytes.begin(), row_bytes.end());
        const uint64_t entries = (max_row * 19 / 10) * 2 / (ug.w + 1) * 13 / 10 * ((P.W + WB - 1) / WB);
        const uint64_t per_wave = (uint64_t)unit_dedupe_slots(WB) * 6 / 10 * c->sm_count;
        const uint64_t waves = std::max<uint64_t>(1, (entries + per_wave - 1) / per_wave);
        MB = (uint32_t)std::min<uint64_t>(kUsMaxBuckets, waves * c->sm_count);
        if (const char* ub = getenv("GRMKM_UNIT_BUCKETS")) MB = (uint32_t)std::min(kUsMaxBuckets, std::max(1, atoi(ub)));
        ENSURE(c, c->ucur, (size_t)MB * 8);
        ENSURE(c, c->ubeg, (size_t)(MB + 1) * 8);
    }
    const bool try_regions = staged && !(c->cfg.flags & GRMKM_FLAG_EXACT_OFFSETS);
    const uint32_t agrid = std::min<uint32_t>(B, (uint32_t)c->sm_count * kAggCtasPerSm);
    const uint32_t VB = B << P.sub_bits;            // virtual buckets of the column aggregate
    const uint32_t vgrid = std::min<uint32_t>(VB, (uint32_t)c->sm_count * kAggCtasPerSm);
    // final build in hash order: the aggregate writes the columns at their final place (no gather pass)
    const bool ordered = mode == 0 && !(c->cfg.flags & GRMKM_FLAG_KMER_ORDER) && !getenv("GRMKM_UNORDERED");
    uint64_t sc[S_COUNT];
    uint64_t ucap = 0;
    for (int pass = try_regions ? 0 : 1; pass < 2; ++pass) {
        const bool regions = (pass == 0);
        c->pass_groups = 0; c->pass_pub_bytes = 0;
        const std::vector<Batch> batches = make_batches(regions && any_host);
        const bool pipelined = batches.size() > 1;
        {   // host table arena for all batches of this pass (nothing of an earlier pass / build is in flight: both end synchronised)
            size_t need = 0;
            for (const Batch& bt : batches)
                need += (((size_t)(bt.f1 - bt.f0) * sizeof(FileDesc) + (bt.f1 - bt.f0 + 1) * 8 + (size_t)bt.tiles * 4 + 63) & ~size_t(63));
            if (need > c->h_tab_cap) {
                CU_TRY(c, cudaStreamSynchronize(st));
                if (c->h_tab) cudaFreeHost(c->h_tab);
                c->h_tab = nullptr; c->h_tab_cap = 0;
                CU_TRY(c, cudaHostAlloc(&c->h_tab, need + need / 4, cudaHostAllocDefault));
                c->h_tab_cap = need + need / 4;
            }
            c->h_tab_used = 0;
        }
        uint64_t cap = 0;
        if (use_units) {
            if (regions) {
                const double est_units = (double)P.max_stream * 2.0 / (ug.w + 1) * 1.15;
                cap = (uint64_t)(est_units / MB * 1.25) + 4096;
                cap = (cap + 15) & ~15ULL;
                ENSURE(c, c->units, ((uint64_t)MB * cap + kUsStage) * 16);
            } else {
                ENSURE(c, c->units, (P.max_stream + kUsStage) * 16);
            }
        } else if (regions) {
            cap = (uint64_t)((double)P.max_stream / B * 1.25) + 2048;
            cap = (cap + 15) & ~15ULL;
            ENSURE(c, c->records, ((uint64_t)B * cap + kStTile) * 8);
        } else {
            ENSURE(c, c->records, P.max_stream * 8);
        }
        uint64_t half_bytes = 0;
        for (const Batch& bt : batches) half_bytes = std::max(half_bytes, bt.staged);
        ENSURE(c, c->in, half_bytes * (pipelined ? 2 : 1));
        if (pipelined && !c->copy_stream) {
            CU_TRY(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                CU_TRY(c, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
                CU_TRY(c, cudaEventCreateWithFlags(&c->ev_free[i], cudaEventDisableTiming));
            }
        }
        {
            const uint64_t zeros[S_COUNT] = {0};
            CU_TRY(c, cudaMemcpyAsync(d_scalars, zeros, S_COUNT * 8, cudaMemcpyHostToDevice, st));
        }
        if (regions) {
            if (use_units)
                k_init_regions<<<(MB + 1 + 255) / 256, 256, 0, st>>>((unsigned long long*)c->ubeg.p,
                                                                     (unsigned long long*)c->ucur.p, MB, cap);
            else
                k_init_regions<<<(B + 1 + 255) / 256, 256, 0, st>>>((unsigned long long*)c->offsets.p,
                                                                    (unsigned long long*)c->hist.p, B, cap);
            L.n++;
        }
        if (pipelined && c->ev_ok)
            for (int e : {T_H2D, T_PARSE, T_PACK, T_COUNT, T_BOUNDS}) cudaEventRecord(c->ev[e], st);   // stages interleave: only "scatter" is timed
        h2d = 0;
        for (size_t bi = 0; bi < batches.size(); ++bi) {
            const Batch& bt = batches[bi];
            uint8_t* half = (uint8_t*)c->in.p + (pipelined ? (bi & 1) * half_bytes : 0);
            cudaStream_t cs = pipelined ? c->copy_stream : st;
            if (pipelined && bi >= 2) CU_TRY(c, cudaStreamWaitEvent(cs, c->ev_free[bi & 1], 0));
            uint64_t soff = 0;
            for (uint32_t f = bt.f0; f < bt.f1; ++f) {
                const Input& in = c->inputs[f];
                if (in.dev) continue;
                if (in.len) { CU_TRY(c, cudaMemcpyAsync(half + soff, in.host, in.len, cudaMemcpyHostToDevice, cs)); h2d += in.len; }
                soff += (in.len + 15) & ~15ULL;
            }
            if (pipelined) {
                CU_TRY(c, cudaEventRecord(c->ev_copied[bi & 1], cs));
                CU_TRY(c, cudaStreamWaitEvent(st, c->ev_copied[bi & 1], 0));
            }
            int fr = front(bt, half, !pipelined);
            if (fr) return fr;
            if (pipelined) CU_TRY(c, cudaEventRecord(c->ev_free[bi & 1], st));     // the staged text has been consumed

            const uint32_t F = bt.f1 - bt.f0;
            const uint64_t n_groups_max = bt.stream / 32 + 2;
            const uint64_t n_stiles = (n_groups_max + (kStTile / 32) - 1) / (kStTile / 32);
            if (use_units) {
                ENSURE(c, c->masks, n_groups_max * 8);
                const uint64_t n_utiles = (n_groups_max + kUsTileGroups - 1) / kUsTileGroups;
                ENSURE(c, c->stile_file, n_utiles * 4);
                if (!pipelined && c->ev_ok) cudaEventRecord(c->ev[T_COUNT], st);
                const uint32_t bgrid = (uint32_t)((n_groups_max + 511) / 512);       // two groups per thread
#define GRMKM_BOUNDS(WW)                                                                                        \
    k_unit_bounds<WW><<<bgrid, 256, 0, st>>>((const unsigned long long*)c->codes.p, (const uint32_t*)c->valid.p, d_scalars, \
                                             ug.k, ug.m, (uint2*)c->masks.p)
                switch (ug.w) {
                    case 21: GRMKM_BOUNDS(21); break;
                    case 13: GRMKM_BOUNDS(13); break;
                    case 7: GRMKM_BOUNDS(7); break;
                    case 4: GRMKM_BOUNDS(4); break;
                    case 2: GRMKM_BOUNDS(2); break;
                    default: GRMKM_BOUNDS(1); break;
                }
#undef GRMKM_BOUNDS
                k_stream_tile_files<<<(uint32_t)((n_utiles + 255) / 256), 256, 0, st>>>(
                    d_scalars, (const uint64_t*)c->fss.p, F, (uint32_t*)c->stile_file.p, n_utiles, (uint64_t)kUsTileGroups * 32);
                L.n += 2;
                CU_TRY(c, cudaGetLastError());
                if (!pipelined && c->ev_ok) cudaEventRecord(c->ev[T_BOUNDS], st);
                UnitScatterParams up{};
                up.codes = (const unsigned long long*)c->codes.p; up.masks = (const uint2*)c->masks.p; up.scalars = d_scalars;
                up.file_stream_start = (const uint64_t*)c->fss.p; up.files = d_files; up.tile_file = (const uint32_t*)c->stile_file.p;
                up.n_files = F; up.k = ug.k; up.lmax = ug.lmax; up.n_buckets = MB;
                up.cursors = (unsigned long long*)c->ucur.p; up.units = (uint4*)c->units.p; up.cap = cap;
                up.dump = (uint64_t)MB * cap; up.overflow = (unsigned long long*)(d_scalars + S_OVERFLOW);
                up.n_windows = (unsigned long long*)(d_scalars + S_N_WINDOWS);
                const size_t usm = unit_scatter_smem(MB);
                const uint32_t ugrid = (uint32_t)std::min<uint64_t>(n_utiles, (uint64_t)c->sm_count);
                if (!regions) {
                    CU_TRY(c, cudaMemsetAsync(c->ucur.p, 0, (size_t)MB * 8, st));
                    CU_TRY(c, cudaFuncSetAttribute(k_units_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usm));
                    k_units_scatter<true><<<ugrid, kUsThreads, usm, st>>>(up);
                    k_bucket_offsets<<<1, 1024, 0, st>>>((unsigned long long*)c->ucur.p, (unsigned long long*)c->ubeg.p, MB,
                                                         d_scalars, S_N_UNITS, 1);
                    L.n += 2;
                }
                CU_TRY(c, cudaFuncSetAttribute(k_units_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usm));
                k_units_scatter<false><<<ugrid, kUsThreads, usm, st>>>(up);
                L.n++;
                CU_TRY(c, cudaGetLastError());
                continue;
            }
            ExtractParams ep{};
            ep.codes = (const unsigned long long*)c->codes.p;
            ep.valid = (const uint32_t*)c->valid.p;
            ep.scalars = d_scalars;
            ep.file_stream_start = (const uint64_t*)c->fss.p;
            ep.files = d_files;
            ep.n_files = F;
            ep.k = c->cfg.k;
            ep.bucket_bits = P.bucket_bits;
            ep.row_bits = P.row_bits;
            ep.hist = (unsigned long long*)c->hist.p;
            ep.records = (unsigned long long*)c->records.p;
            ep.dbg = 0;
            ep.offsets = (const unsigned long long*)c->offsets.p;
            const uint64_t n_etiles_max = (n_groups_max + kExtractThreads - 1) / kExtractThreads;
            const uint32_t egrid = (uint32_t)std::min<uint64_t>(n_etiles_max, (uint64_t)c->sm_count * 8);
            if (!regions) {
                CU_TRY(c, cudaMemsetAsync(c->hist.p, 0, (size_t)B * 8, st));
                const size_t hist_smem = (size_t)B * 4;
                CU_TRY(c, cudaFuncSetAttribute(k_extract<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_smem));
                k_extract<0><<<egrid, kExtractThreads, hist_smem, st>>>(ep);
                k_bucket_offsets<<<1, 1024, 0, st>>>((unsigned long long*)c->hist.p, (unsigned long long*)c->offsets.p, B,
                                                     d_scalars, S_N_WINDOWS, kCursorStride);
                L.n += 2;
            }
            CU_TRY(c, cudaGetLastError());
            if (!pipelined && c->ev_ok) { cudaEventRecord(c->ev[T_COUNT], st); cudaEventRecord(c->ev[T_BOUNDS], st); }
            if (staged) {
                ENSURE(c, c->stile_file, n_stiles * 4);
                ScatterParams sp{};
                sp.codes = ep.codes; sp.valid = ep.valid; sp.scalars = d_scalars; sp.file_stream_start = ep.file_stream_start;
                sp.files = d_files; sp.tile_file = (const uint32_t*)c->stile_file.p; sp.n_files = F; sp.k = c->cfg.k;
                sp.bucket_bits = P.bucket_bits; sp.row_bits = P.row_bits; sp.cursors = (unsigned long long*)c->hist.p;
                sp.records = (unsigned long long*)c->records.p; sp.cap = cap; sp.dump = (uint64_t)B * cap;
                sp.overflow = (unsigned long long*)(d_scalars + S_OVERFLOW);
                k_scatter_tile_files<<<(uint32_t)((n_stiles + 255) / 256), 256, 0, st>>>(d_scalars, sp.file_stream_start, F,
                                                                                      (uint32_t*)c->stile_file.p, n_stiles);
                const size_t ssm = staged_smem_bytes(B);
                const uint32_t sgrid = (uint32_t)std::min<uint64_t>(n_stiles, (uint64_t)c->sm_count);
#define GRMKM_SCATTER(KT)                                                                                       \
    do {                                                                                                        \
        CU_TRY(c, cudaFuncSetAttribute(k_scatter<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));  \
        k_scatter<KT><<<sgrid, kStThreads, ssm, st>>>(sp);                                                      \
    } while (0)
                switch (c->cfg.k) {
                    case 31: GRMKM_SCATTER(31); break;
                    case 21: GRMKM_SCATTER(21); break;
                    case 15: GRMKM_SCATTER(15); break;
                    default: GRMKM_SCATTER(0); break;
                }
#undef GRMKM_SCATTER
                L.n += 2;
            } else {
                k_extract<1><<<egrid, kExtractThreads, 0, st>>>(ep);
                L.n++;
            }
            CU_TRY(c, cudaGetLastError());
        }
        // bucket b = records[begin[b], end[b]): begin = offsets, end = the cursors (clamped to the region)
        if (use_units)
            k_finish_unit_regions<<<1, 1024, 0, st>>>((const unsigned long long*)c->ubeg.p, (unsigned long long*)c->ucur.p, MB,
                                                      cap, (unsigned long long*)d_scalars, S_N_UNITS);
        else
            k_finish_regions<<<1, 1024, 0, st>>>((const unsigned long long*)c->offsets.p, (unsigned long long*)c->hist.p, B, cap,
                                                 (unsigned long long*)d_scalars, S_N_WINDOWS);
        L.n++;
        CU_TRY(c, cudaGetLastError());
        if (c->ev_ok) cudaEventRecord(c->ev[T_SCATTER], st);
        // codes / valid / tile words are dead from here on: clear them for the next build (or the next pass) on the side
        // stream, under the dedupe / expand / aggregate kernels, which leave most of the HBM bandwidth unused
        if (c->pass_groups && !getenv("GRMKM_NO_PRECLEAR")) {
            if (!c->aux_stream) {
                CU_TRY(c, cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
                CU_TRY(c, cudaEventCreateWithFlags(&c->ev_scattered, cudaEventDisableTiming));
                CU_TRY(c, cudaEventCreateWithFlags(&c->ev_cleared, cudaEventDisableTiming));
            }
            CU_TRY(c, cudaEventRecord(c->ev_scattered, st));
            CU_TRY(c, cudaStreamWaitEvent(c->aux_stream, c->ev_scattered, 0));
            CU_TRY(c, cudaMemsetAsync(c->codes.p, 0, c->pass_groups * 8, c->aux_stream));
            CU_TRY(c, cudaMemsetAsync(c->valid.p, 0, c->pass_groups * 4, c->aux_stream));
            CU_TRY(c, cudaMemsetAsync(c->tile_pub.p, 0, c->pass_pub_bytes, c->aux_stream));
            CU_TRY(c, cudaEventRecord(c->ev_cleared, c->aux_stream));
            c->pre_pending = true;
            c->pre_codes = c->codes.p; c->pre_valid = c->valid.p; c->pre_pub = c->tile_pub.p;
            c->pre_groups = c->pass_groups; c->pre_pub_bytes = c->pass_pub_bytes;
        }

        // ---- optional abundance filter (reads: -abundance-min, kmer_count.py:48)
        const unsigned long long* agg_records = (const unsigned long long*)c->records.p;
        const unsigned long long* agg_begin = (const unsigned long long*)c->offsets.p;
        const unsigned long long* agg_end = (const unsigned long long*)c->hist.p;
        // Unit path: dedupe -> expand -> aggregate run back to back without a host round trip.  The dedupe's entry list
        // and the expansion's bucket regions are sized from estimates; the one synchronisation after the aggregate
        // reads what was really needed, and a guess that was too small repeats the round (entry list: larger; regions:
        // exact offsets from a count pass).
        uint64_t wcap = std::min<uint64_t>(P.max_stream, std::max<uint64_t>(1 << 16, P.max_stream / 16));
        bool wide_exact = !try_regions;
        if (getenv("GRMKM_WIDE_EXACT")) wide_exact = true;
        if (const char* e = getenv("GRMKM_WU_CAP")) wcap = std::max<uint64_t>(16, (uint64_t)atoll(e));     // tests: force the retry
        bool again = true, unit_overflow = false;
        for (int round = 0; again; ++round) {
        again = false;
        if (round > 6) return fail(c, GRMKM_E_UNSUPPORTED, "unit path does not converge");
        if (round > 0) {
            const uint64_t zero3[3] = {0, 0, 0};
            CU_TRY(c, cudaMemcpyAsync(d_scalars + S_U_NEEDED, zero3, 3 * 8, cudaMemcpyHostToDevice, st));      // + S_N_DISTINCT, S_N_SPLITS
            CU_TRY(c, cudaMemcpyAsync(d_scalars + S_WU_NEEDED, zero3, 3 * 8, cudaMemcpyHostToDevice, st));     // + S_N_WIDE, S_WIDE_OVERFLOW
            CU_TRY(c, cudaStreamSynchronize(st));          // zero3 lives on this stack frame
        }
        if (use_units) {
            if (c->ev_ok) cudaEventRecord(c->ev[T_ABUND], st);
            const size_t dsm = unit_dedupe_smem(WB), xsm = staged_smem_bytes(B);
            void (*k_dedupe)(UnitDedupeParams) = WB == 1 ? k_units_dedupe<1> : WB == 2 ? k_units_dedupe<2> : k_units_dedupe<4>;
            void (*k_xcount)(UnitExpandParams) = WB == 1 ? k_units_expand<true, 1> : WB == 2 ? k_units_expand<true, 2> : k_units_expand<true, 4>;
            void (*k_xscat)(UnitExpandParams) = WB == 1 ? k_units_expand<false, 1> : WB == 2 ? k_units_expand<false, 2> : k_units_expand<false, 4>;
            CU_TRY(c, cudaFuncSetAttribute(k_dedupe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
            CU_TRY(c, cudaFuncSetAttribute(k_xcount, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
            CU_TRY(c, cudaFuncSetAttribute(k_xscat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xsm));
            // ---- dedupe: distinct (unit, block) entries
            ENSURE(c, c->wu, wcap * ES * 8);
            UnitDedupeParams dp{};
            dp.units = (const uint4*)c->units.p; dp.begin = (const unsigned long long*)c->ubeg.p;
            dp.end = (const unsigned long long*)c->ucur.p; dp.n_buckets = MB; dp.out = (unsigned long long*)c->wu.p;
            dp.cap = wcap; dp.needed = (unsigned long long*)(d_scalars + S_WU_NEEDED);
            k_dedupe<<<std::min<uint32_t>(MB, (uint32_t)c->sm_count), kUdThreads, dsm, st>>>(dp);
            L.n++;
            CU_TRY(c, cudaGetLastError());
            if (c->ev_ok) cudaEventRecord(c->ev[T_DEDUPE], st);
            // ---- expand: every distinct entry -> wide records of its k-mers, by hash bucket
            UnitExpandParams xp{};
            xp.wu = (const unsigned long long*)c->wu.p; xp.n_ptr = (const unsigned long long*)(d_scalars + S_WU_NEEDED);
            xp.cap = wcap; xp.k = c->cfg.k; xp.bucket_bits = P.bucket_bits; xp.wbits = wbits;
            xp.cursors = (unsigned long long*)c->hist.p; xp.records = nullptr;
            xp.overflow = (unsigned long long*)(d_scalars + S_WIDE_OVERFLOW);
            const uint32_t xgrid = (uint32_t)std::min<uint64_t>((wcap + kStThreads - 1) / kStThreads, (uint64_t)c->sm_count);
            if (!wide_exact) {
                // over-provisioned bucket regions, no count pass: the distinct k-mers are estimated as 1.9 x the largest genome
                // plus 5 % of all input (what every further genome adds to a species' pan-genome), 1.3 records per distinct
                // k-mer and genome group, 1.4 x headroom per bucket; the previous build of this context corrects the guess
                std::vector<uint64_t> row_bytes(std::max(P.G, 1u), 0);
                for (const Input& in : c->inputs) if (in.row < P.G) row_bytes[in.row] += in.len;
                const uint64_t max_row = *std::max_element(row_bytes.begin(), row_bytes.end());
                const uint64_t groups = (P.W + WB - 1) / WB;
                uint64_t est = std::max<uint64_t>(c->wide_hint, std::min<uint64_t>(P.in_bytes, max_row * 19 / 10 + P.in_bytes / 20) * 13 / 10 * groups);
                if (const char* wc = getenv("GRMKM_WIDE_EST")) est = (uint64_t)atoll(wc);
                // (rounded BEFORE the capacity test: tested unrounded, the size asked for was always a little more than the
                // size allocated, and cudaMemGetInfo -- a driver round trip that takes anything from 0.1 to 90 ms while the
                // GPU waits for the next launch -- ran in every build)
                uint64_t rcap = std::max<uint64_t>(64, ((uint64_t)((double)est / B * 1.4) + 1024) & ~15ULL);
                const uint64_t wide_need = ((uint64_t)B * rcap + kStTile) * RS * 8;
                if (wide_need > c->wide.cap) {
                    // the buffer has to grow: never beyond a quarter of what is free (a region that turns out too small
                    // only costs the count pass).  Not asked in the steady state -- cudaMemGetInfo is slow.
                    size_t free_b = 0, total_b = 0;
                    if (c->wide.cap && wide_need <= c->wide_capped_need) {
                        // a request this large was already cut down to what the context holds now: keep the buffer
                        rcap = std::min<uint64_t>(rcap, (c->wide.cap / (RS * 8) - kStTile) / B);
                    } else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                        // (what it has already counts as the budget when that is more: no re-allocation every build)
                        const uint64_t have = std::max<uint64_t>(c->wide.cap, free_b / 4);
                        if (wide_need > have) c->wide_capped_need = wide_need;
                        rcap = std::min<uint64_t>(rcap, (have / (RS * 8) - kStTile) / B);
                    }
                }
                rcap = std::max<uint64_t>(64, rcap & ~15ULL);
                ENSURE(c, c->wide, ((uint64_t)B * rcap + kStTile) * RS * 8);
                k_init_regions<<<(B + 1 + 255) / 256, 256, 0, st>>>((unsigned long long*)c->offsets.p, (unsigned long long*)c->hist.p, B, rcap);
                xp.records = (unsigned long long*)c->wide.p; xp.region_cap = rcap; xp.dump = (uint64_t)B * rcap;
                k_xscat<<<xgrid, kStThreads, xsm, st>>>(xp);
                k_finish_regions<<<1, 1024, 0, st>>>((const unsigned long long*)c->offsets.p, (unsigned long long*)c->hist.p, B, rcap,
                                                     (unsigned long long*)d_scalars, S_N_WIDE);
                L.n += 3;
                CU_TRY(c, cudaGetLastError());
                agg_begin = (const unsigned long long*)c->offsets.p;
                agg_end = (const unsigned long long*)c->hist.p;
            } else {
                // exact offsets: count pass, scan, one synchronisation to size the record array
                CU_TRY(c, cudaMemsetAsync(c->hist.p, 0, (size_t)B * 8, st));
                k_xcount<<<xgrid, kStThreads, xsm, st>>>(xp);
                k_bucket_offsets<<<1, 1024, 0, st>>>((unsigned long long*)c->hist.p, (unsigned long long*)c->offsets.p, B,
                                                     d_scalars, S_N_WIDE, kCursorStride);
                L.n += 2;
                CU_TRY(c, cudaGetLastError());
                CU_TRY(c, cudaMemcpyAsync(sc, d_scalars, sizeof sc, cudaMemcpyDeviceToHost, st));
                CU_TRY(c, cudaStreamSynchronize(st));
                if (regions && sc[S_OVERFLOW]) { unit_overflow = true; break; }
                if (sc[S_WU_NEEDED] > wcap) {
                    // the entry count depends on the order the table fills in (bypass, flush points): leave headroom
                    wcap = sc[S_WU_NEEDED] + sc[S_WU_NEEDED] / 4 + 4096;
                    again = true;
                    continue;
                }
                ENSURE(c, c->wide, (sc[S_N_WIDE] + 1) * RS * 8);
                xp.records = (unsigned long long*)c->wide.p;
                k_xscat<<<xgrid, kStThreads, xsm, st>>>(xp);
                L.n++;
                CU_TRY(c, cudaGetLastError());
                agg_begin = (const unsigned long long*)c->offsets.p;
                agg_end = agg_begin + 1;
            }
            if (c->ev_ok) cudaEventRecord(c->ev[T_EXPAND], st);
            agg_records = (const unsigned long long*)c->wide.p;
        }
        if (c->cfg.min_abundance > 1) {
            ENSURE(c, c->records2, c->records.cap);
            ENSURE(c, c->bcounts, (size_t)B * 8);
            ENSURE(c, c->offsets2, (size_t)(B + 1) * 8);
            AggParams ap{};
            ap.records = agg_records; ap.begin = agg_begin; ap.end = agg_end; ap.B = B; ap.bucket_bits = P.bucket_bits;
            ap.row_bits = P.row_bits; ap.n_words = 1;
            ap.slots = (uint32_t)std::min<size_t>(kAggMaxSlots, budget / 16);
            ap.mode = 2; ap.min_abundance = c->cfg.min_abundance; ap.b_begin = 0; ap.b_end = B;
            ap.scalars = (unsigned long long*)d_scalars;
            ap.bucket_out_counts = (unsigned long long*)c->bcounts.p;
            ap.out_records = (unsigned long long*)c->records2.p;
            const size_t sm = (size_t)ap.slots * 16;
            CU_TRY(c, cudaFuncSetAttribute(k_aggregate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            k_aggregate<2><<<agrid, kAggThreads, sm, st>>>(ap);
            k_bucket_offsets<<<1, 1024, 0, st>>>((unsigned long long*)c->bcounts.p, (unsigned long long*)c->offsets2.p, B,
                                                 d_scalars, S_N_SOLID, 1);
            k_compact_records<<<agrid * 4, 256, 0, st>>>((const unsigned long long*)c->records2.p, agg_begin,
                                                         (const unsigned long long*)c->offsets2.p, B,
                                                         (unsigned long long*)c->records.p);
            L.n += 3;
            CU_TRY(c, cudaGetLastError());
            agg_begin = (const unsigned long long*)c->offsets2.p;
            agg_end = agg_begin + 1;
        }
        if (!use_units && c->ev_ok) for (int e : {T_ABUND, T_DEDUPE, T_EXPAND}) cudaEventRecord(c->ev[e], st);

        // ---- aggregate (dsk2kover): retry with a larger output if the first guess was too small
        // columns: three pan-genome estimates (1.9 x the largest genome each) + 1 % of all input, at most a quarter of the
        // windows (N / 4 alone asked for 163 GB of word rows at 1000 genomes); a guess that is too small is retried with
        // the exact number
        {
            uint64_t max_row = 0;
            std::vector<uint64_t> rb(std::max(P.G, 1u), 0);
            for (const Input& in : c->inputs) if (in.row < P.G) { rb[in.row] += in.len; max_row = std::max(max_row, rb[in.row]); }
            const uint64_t guess = 3 * (max_row * 19 / 10) + P.in_bytes / 100;
            ucap = std::min<uint64_t>(P.max_stream, std::max<uint64_t>(1 << 16, std::min<uint64_t>(P.max_stream / 4, std::max<uint64_t>(guess, c->ucap_hint))));
        }
        ucap = std::min<uint64_t>(ucap, 0xFFFFFFFFULL);
        ENSURE(c, c->bbase, (size_t)VB * 8);
        ENSURE(c, c->bcounts, (size_t)VB * 8);
        if (ordered) ENSURE(c, c->apub, (size_t)(VB + 1) * 8);
        for (int attempt = 0; attempt < 2; ++attempt) {
            if (ordered) {
                // the columns land at their final place: k-mers and word row 0 in the result arrays, rows >= 1 at stride ucap
                ENSURE(c, c->kmers, ucap * 8);
                ENSURE(c, c->matrix, (size_t)ucap * P.W * 8);
                if (P.W > 1) ENSURE(c, c->uwords, (size_t)ucap * P.W * 8);
                CU_TRY(c, cudaMemsetAsync(c->apub.p, 0, (size_t)(VB + 1) * 8, st));
            } else {
                ENSURE(c, c->ukeys, ucap * 8);
                ENSURE(c, c->uwords, (size_t)ucap * P.W * 8);
            }
            AggParams2 ap{};
            ap.records = agg_records; ap.begin = agg_begin; ap.end = agg_end; ap.bucket_bits = P.bucket_bits;
            ap.row_bits = use_units ? wbits : P.row_bits; ap.n_words = P.W; ap.slots = P.slots;
            ap.keep_singletons = c->cfg.keep_singletons; ap.sub_bits = P.sub_bits;
            ap.wide_words = WB; ap.wide_stride = RS; ap.table_u32 = 2 * P.W;
            ap.out_keys = (unsigned long long*)c->ukeys.p; ap.out_words = (unsigned long long*)c->uwords.p;
            ap.cap = ucap; ap.scalars = (unsigned long long*)d_scalars;
            ap.bucket_base = (unsigned long long*)c->bbase.p; ap.bucket_count = (unsigned long long*)c->bcounts.p;
            ap.b_begin = 0; ap.b_end = B;
            if (ordered) {
                ap.ordered = 1;
                ap.out_keys = (unsigned long long*)c->kmers.p; ap.out_row0 = (unsigned long long*)c->matrix.p;
                ap.pub = (unsigned long long*)c->apub.p; ap.ticket = (unsigned int*)((unsigned long long*)c->apub.p + VB);
            }
            if (use_units && mode == 0) {
                CU_TRY(c, cudaFuncSetAttribute(k_aggregate_cols<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.agg_smem));
                k_aggregate_cols<4><<<vgrid, kAggThreads, P.agg_smem, st>>>(ap);
            } else if (use_units) {
                CU_TRY(c, cudaFuncSetAttribute(k_aggregate_cols<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.agg_smem));
                k_aggregate_cols<5><<<vgrid, kAggThreads, P.agg_smem, st>>>(ap);
            } else if (mode == 0) {
                CU_TRY(c, cudaFuncSetAttribute(k_aggregate_cols<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.agg_smem));
                k_aggregate_cols<0><<<vgrid, kAggThreads, P.agg_smem, st>>>(ap);
            } else {
                CU_TRY(c, cudaFuncSetAttribute(k_aggregate_cols<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.agg_smem));
                k_aggregate_cols<1><<<vgrid, kAggThreads, P.agg_smem, st>>>(ap);
            }
            L.n++;
            CU_TRY(c, cudaGetLastError());
            if (ordered) {
                // rows >= 1 move from stride ucap to the result's stride U (row 0 and the k-mers are already in place); U is read
                // on the device, so the build ends with this one synchronisation
                if (c->ev_ok) cudaEventRecord(c->ev[T_AGG], st);
                if (P.W > 1) {
                    k_move_rows<<<(uint32_t)c->sm_count * 8, 256, 0, st>>>((const unsigned long long*)c->uwords.p, ucap, (unsigned long long*)c->matrix.p,
                                                                          (const unsigned long long*)(d_scalars + S_U_NEEDED), P.W);
                    L.n++;
                    CU_TRY(c, cudaGetLastError());
                }
            }
            CU_TRY(c, cudaMemcpyAsync(sc, d_scalars, sizeof sc, cudaMemcpyDeviceToHost, st));
            CU_TRY(c, cudaStreamSynchronize(st));
            if (regions && sc[S_OVERFLOW]) break;
            if (use_units && sc[S_WU_NEEDED] > wcap) {
                wcap = sc[S_WU_NEEDED] + sc[S_WU_NEEDED] / 4 + 4096;
                again = true;
                break;
            }
            if (use_units && !wide_exact && sc[S_WIDE_OVERFLOW]) {
                wide_exact = true; again = true; c->stats.n_region_overflows++;
                break;
            }
            if (sc[S_U_NEEDED] <= ucap) break;
            if (attempt == 1 || sc[S_U_NEEDED] > 0xFFFFFFFFULL)
                return fail(c, GRMKM_E_UNSUPPORTED, "more than 2^32 columns in one context");
            ucap = sc[S_U_NEEDED];
            const uint64_t zero3[3] = {0, 0, 0};
            CU_TRY(c, cudaMemcpyAsync(d_scalars + S_U_NEEDED, zero3, 3 * 8, cudaMemcpyHostToDevice, st));
        }
        }   // round
        if (use_units && !unit_overflow && !(regions && sc[S_OVERFLOW])) c->wide_hint = sc[S_N_WIDE];
        if (!(regions && sc[S_OVERFLOW])) c->ucap_hint = sc[S_U_NEEDED] + sc[S_U_NEEDED] / 8;
        if (!(regions && (sc[S_OVERFLOW] || unit_overflow))) break;
        c->stats.n_region_overflows++;
    }
    if (c->ev_ok && !ordered) cudaEventRecord(c->ev[T_AGG], st);
    const uint64_t U = sc[S_U_NEEDED];

    // ---- final order.  Default: ascending hash (bucket order; every bucket chunk is already sorted), which is
    // identical for any GPU count.  GRMKM_FLAG_KMER_ORDER: ascending canonical k-mer (one extra sort).
    std::vector<uint64_t> h_off;
    if (mode == 0 && (c->cfg.flags & GRMKM_FLAG_KMER_ORDER)) {
        int r = sort_and_gather(c, U, P.W, ucap, 2 * c->cfg.k, false, L);
        if (r) return r;
    } else if (ordered) {
        // nothing left to do: k_move_rows ran behind the aggregate
    } else {
        ENSURE(c, c->offsets2, (size_t)(VB + 1) * 8);
        ENSURE(c, c->kmers, U * 8);
        ENSURE(c, c->matrix, (size_t)U * P.W * 8);
        k_bucket_offsets<<<1, 1024, 0, st>>>((unsigned long long*)c->bcounts.p, (unsigned long long*)c->offsets2.p, VB,
                                             d_scalars, S_N_SOLID, 1);
        if (U && mode == 0) {
            k_gather_buckets<<<std::min<uint32_t>(VB, (uint32_t)c->sm_count * 8), 256, 0, st>>>(
                (const unsigned long long*)c->ukeys.p, (const unsigned long long*)c->uwords.p, ucap,
                (const unsigned long long*)c->bbase.p, (const unsigned long long*)c->offsets2.p, VB, P.W, U,
                (unsigned long long*)c->kmers.p, (unsigned long long*)c->matrix.p, U);
            L.n++;
        }                               // partial build: grmkm_export_partials gathers straight into the caller's AoS buffer
        L.n++;
        CU_TRY(c, cudaGetLastError());
        if (mode == 1) {
            h_off.resize(VB + 1);
            CU_TRY(c, cudaMemcpyAsync(h_off.data(), c->offsets2.p, (size_t)(VB + 1) * 8, cudaMemcpyDeviceToHost, st));
        }
    }
    if (c->ev_ok) cudaEventRecord(c->ev[T_SORT], st);
    CU_TRY(c, cudaStreamSynchronize(st));

    c->U = U;
    c->built = (mode == 0);
    if (mode == 1) {
        // partial columns are in bucket order: owner r holds buckets [r*B/P, (r+1)*B/P)
        c->part_ranks = n_ranges;
        c->part_counts.resize(n_ranges);
        for (uint32_t r = 0; r < n_ranges; ++r)
            c->part_counts[r] = h_off[((uint64_t)B * (r + 1) / n_ranges) << P.sub_bits] - h_off[((uint64_t)B * r / n_ranges) << P.sub_bits];
        c->part_total = U; c->part_cap = ucap; c->part_words = P.W; c->part_buckets = VB;
    }
    grmkm_stats& s = c->stats;
    s.n_input_bytes = P.in_bytes;
    s.n_records = sc[S_N_RECORDS];
    s.n_bases = sc[S_STREAM_TOTAL] - sc[S_N_RECORDS];
    s.n_windows = sc[S_N_WINDOWS];
    s.n_kmers = U;
    s.n_distinct = sc[S_N_DISTINCT];
    s.n_words = P.W; s.n_genomes = P.G; s.n_buckets = B;
    s.n_launches = L.n;
    s.h2d_bytes = h2d;
    s.device_bytes = c->device_bytes;
    s.n_splits = sc[S_N_SPLITS];
    s.n_units = use_units ? sc[S_N_UNITS] : 0;
    s.n_unit_entries = use_units ? sc[S_WU_NEEDED] : 0;
    s.n_wide = use_units ? sc[S_N_WIDE] : 0;
    s.n_unit_buckets = use_units ? MB : 0;
    if (c->ev_ok) {
        float* t[] = {&c->times.h2d, &c->times.parse, &c->times.pack, &c->times.count, &c->times.bounds, &c->times.scatter,
                      &c->times.abundance, &c->times.dedupe, &c->times.expand, &c->times.aggregate, &c->times.sort};
        for (int i = 1; i < T_N; ++i) {
            const cudaError_t te = cudaEventElapsedTime(t[i - 1], c->ev[i - 1], c->ev[i]);
            if (te != cudaSuccess) {
                cudaGetLastError();
                if (getenv("GRMKM_DEBUG")) fprintf(stderr, "grmkm: stage event %d: %s\n", i, cudaGetErrorString(te));
                *t[i - 1] = 0.f;
            }
        }
        cudaEventElapsedTime(&c->times.total, c->ev[T_START], c->ev[T_SORT]);
    }
    return GRMKM_OK;
}

int grmkm_build(grmkm_ctx* c) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    return build_impl(c, 0, 1);
}

int grmkm_dims(const grmkm_ctx* c, uint64_t* n_kmers, uint32_t* n_words, uint32_t* n_genomes) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->built) return fail(const_cast<grmkm_ctx*>(c), GRMKM_E_INVALID, "no result: call grmkm_build first");
    if (n_kmers) *n_kmers = c->U;
    if (n_words) *n_words = c->W;
    if (n_genomes) *n_genomes = c->G;
    return GRMKM_OK;
}

int grmkm_get_stats(const grmkm_ctx* c, grmkm_stats* out) {
    if (check_ctx(c) || !out) return GRMKM_E_INVALID;
    *out = c->stats;
    out->device_bytes = c->device_bytes;
    return GRMKM_OK;
}

int grmkm_stage_times(const grmkm_ctx* c, grmkm_times* out) {
    if (check_ctx(c) || !out) return GRMKM_E_INVALID;
    *out = c->times;
    return GRMKM_OK;
}

int grmkm_copy_kmers_packed(grmkm_ctx* c, uint64_t* dst, uint64_t cap) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->built) return fail(c, GRMKM_E_INVALID, "no result: call grmkm_build first");
    if (cap < c->U) return fail(c, GRMKM_E_CAPACITY, "kmers buffer too small");
    if (!c->U) return GRMKM_OK;
    if (!dst) return fail(c, GRMKM_E_INVALID, "null dst");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaMemcpyAsync(dst, c->kmers.p, c->U * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return GRMKM_OK;
}

int grmkm_copy_matrix(grmkm_ctx* c, uint64_t* dst, uint64_t cap) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->built) return fail(c, GRMKM_E_INVALID, "no result: call grmkm_build first");
    const uint64_t n = c->U * c->W;
    if (cap < n) return fail(c, GRMKM_E_CAPACITY, "matrix buffer too small");
    if (!n) return GRMKM_OK;
    if (!dst) return fail(c, GRMKM_E_INVALID, "null dst");
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaMemcpyAsync(dst, c->matrix.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return GRMKM_OK;
}

int grmkm_copy_kmer_strings(grmkm_ctx* c, char* dst, uint64_t cap) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->built) return fail(c, GRMKM_E_INVALID, "no result: call grmkm_build first");
    const uint32_t k = c->cfg.k;
    const uint64_t n = c->U * k;
    if (cap < n) return fail(c, GRMKM_E_CAPACITY, "kmer string buffer too small");
    if (!n) return GRMKM_OK;
    if (!dst) return fail(c, GRMKM_E_INVALID, "null dst");
    CU_TRY(c, cudaSetDevice(c->device));
    const uint64_t rows_per = std::max<uint64_t>(1, (256ULL << 20) / k);
    ENSURE(c, c->fmt, std::min<uint64_t>(n, rows_per * k));
    for (uint64_t j0 = 0; j0 < c->U; j0 += rows_per) {
        const uint64_t rows = std::min(rows_per, c->U - j0), nb = rows * k;
        k_kmer_strings<<<(uint32_t)((nb + 255) / 256), 256, 0, c->stream>>>((const unsigned long long*)c->kmers.p, j0, nb,
                                                                            k, (char*)c->fmt.p);
        CU_TRY(c, cudaGetLastError());
        CU_TRY(c, cudaMemcpyAsync(dst + j0 * k, c->fmt.p, nb, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return GRMKM_OK;
}

int grmkm_format_tsv(grmkm_ctx* c, const char* const* names, char* dst, uint64_t cap, uint64_t* written) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->built) return fail(c, GRMKM_E_INVALID, "no result: call grmkm_build first");
    if (!names && c->G) return fail(c, GRMKM_E_INVALID, "null names");
    const uint32_t k = c->cfg.k, G = c->G;
    uint64_t hdr = 5;
    for (uint32_t g = 0; g < G; ++g) {
        if (!names[g]) return fail(c, GRMKM_E_INVALID, "null name");
        hdr += 1 + strlen(names[g]);
    }
    hdr += 1;
    const uint64_t roww = (uint64_t)k + 2ULL * G + 1;
    const uint64_t need = hdr + roww * c->U;
    if (written) *written = need;
    if (!dst || cap < need) return fail(c, GRMKM_E_CAPACITY, "tsv buffer too small");
    char* p = dst;
    memcpy(p, "kmers", 5); p += 5;
    for (uint32_t g = 0; g < G; ++g) { *p++ = '\t'; size_t l = strlen(names[g]); memcpy(p, names[g], l); p += l; }
    *p++ = '\n';
    if (!c->U) return GRMKM_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    const uint64_t rows_per = std::max<uint64_t>(1, (256ULL << 20) / roww);
    ENSURE(c, c->fmt, std::min<uint64_t>(roww * c->U, rows_per * roww));
    for (uint64_t j0 = 0; j0 < c->U; j0 += rows_per) {
        const uint64_t rows = std::min(rows_per, c->U - j0), nb = rows * roww;
        k_format_tsv<<<(uint32_t)((nb + 255) / 256), 256, 0, c->stream>>>((const unsigned long long*)c->kmers.p,
                                                                          (const unsigned long long*)c->matrix.p, c->U, G,
                                                                          k, j0, nb, (char*)c->fmt.p);
        CU_TRY(c, cudaGetLastError());
        CU_TRY(c, cudaMemcpyAsync(p + j0 * roww, c->fmt.p, nb, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return GRMKM_OK;
}

int grmkm_host_result(grmkm_ctx* c, const uint64_t** kmers, const uint64_t** matrix) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->built) return fail(c, GRMKM_E_INVALID, "no result: call grmkm_build first");
    const size_t nk = (size_t)c->U * 8, nm = (size_t)c->U * c->W * 8;
    CU_TRY(c, cudaSetDevice(c->device));
    if (c->host_res_cap < nk + nm + 16) {
        if (c->host_res) { cudaFreeHost(c->host_res); c->host_res = nullptr; c->host_res_cap = 0; }
        const size_t want = ((nk + nm) * 5 / 4 + (1 << 20)) & ~size_t(4095);
        cudaError_t e = cudaHostAlloc(&c->host_res, want, cudaHostAllocDefault);
        if (e != cudaSuccess) { c->host_res = nullptr; return fail(c, GRMKM_E_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); }
        c->host_res_cap = want;
    }
    uint8_t* h = (uint8_t*)c->host_res;
    if (nk) CU_TRY(c, cudaMemcpyAsync(h, c->kmers.p, nk, cudaMemcpyDeviceToHost, c->stream));
    if (nm) CU_TRY(c, cudaMemcpyAsync(h + nk, c->matrix.p, nm, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    if (kmers) *kmers = (const uint64_t*)h;
    if (matrix) *matrix = (const uint64_t*)(h + nk);
    return GRMKM_OK;
}

int grmkm_device_result(const grmkm_ctx* c, const uint64_t** d_kmers, const uint64_t** d_matrix) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->built) return fail(const_cast<grmkm_ctx*>(c), GRMKM_E_INVALID, "no result: call grmkm_build first");
    if (d_kmers) *d_kmers = (const uint64_t*)c->kmers.p;
    if (d_matrix) *d_matrix = (const uint64_t*)c->matrix.p;
    return GRMKM_OK;
}

int grmkm_synth_fasta_device(grmkm_ctx* c, const void* layout, uint64_t layout_bytes, void* dev_dst, uint64_t dst_bytes) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!layout || !dev_dst) return fail(c, GRMKM_E_INVALID, "null argument");
    CU_TRY(c, cudaSetDevice(c->device));
    ENSURE(c, c->synth, layout_bytes);
    CU_TRY(c, cudaMemcpyAsync(c->synth.p, layout, layout_bytes, cudaMemcpyHostToDevice, c->stream));
    std::string msg;
    if (!synth_launch((const uint8_t*)layout, layout_bytes, (const uint8_t*)c->synth.p, (uint8_t*)dev_dst, dst_bytes,
                      c->stream, msg))
        return fail(c, GRMKM_E_INVALID, msg);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return GRMKM_OK;
}

int grmkm_set_bucket_bits(grmkm_ctx* c, uint32_t bits) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (bits && (bits < 4 || bits > 15)) return fail(c, GRMKM_E_INVALID, "bucket_bits must be 0 (auto) or 4..15");
    c->cfg.bucket_bits = bits;
    return GRMKM_OK;
}

int grmkm_plan_bucket_bits(grmkm_ctx* c, uint32_t* bits) {
    if (check_ctx(c) || !bits) return GRMKM_E_INVALID;
    uint32_t maxrow = 0;
    for (const Input& in : c->inputs) maxrow = std::max(maxrow, in.row + 1);
    *bits = auto_bucket_bits(c, std::max(1u, std::max(maxrow, c->n_genomes_decl)));
    return GRMKM_OK;
}

int grmkm_build_partial(grmkm_ctx* c, uint32_t n_ranks, uint64_t* counts) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (n_ranks < 1 || n_ranks > 16 || !counts) return fail(c, GRMKM_E_INVALID, "n_ranks must be 1..16");
    int r = build_impl(c, 1, n_ranks);
    if (r) return r;
    for (uint32_t i = 0; i < n_ranks; ++i) counts[i] = c->part_counts[i];
    return GRMKM_OK;
}

int grmkm_export_partials(grmkm_ctx* c, void* dev_dst, uint64_t dst_bytes) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->part_ranks) return fail(c, GRMKM_E_INVALID, "no partial result: call grmkm_build_partial first");
    const uint64_t need = c->part_total * (1 + (uint64_t)c->part_words) * 8;
    if (dst_bytes < need) return fail(c, GRMKM_E_CAPACITY, "partial export buffer too small");
    if (!c->part_total) return GRMKM_OK;
    if (!dev_dst) return fail(c, GRMKM_E_INVALID, "null dst");
    CU_TRY(c, cudaSetDevice(c->device));
    k_gather_buckets_aos<<<std::min<uint32_t>(c->part_buckets, (uint32_t)c->sm_count * 8), 256, 0, c->stream>>>(
        (const unsigned long long*)c->ukeys.p, (const unsigned long long*)c->uwords.p, c->part_cap,
        (const unsigned long long*)c->bbase.p, (const unsigned long long*)c->offsets2.p, c->part_buckets, c->part_words,
        (unsigned long long*)dev_dst);
    CU_TRY(c, cudaGetLastError());
    c->stats.n_launches++;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return GRMKM_OK;
}

int grmkm_export_partials_peers(grmkm_ctx* c, uint32_t n_ranks, void* const* peer_dst, const uint64_t* peer_word_off) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (!c->part_ranks) return fail(c, GRMKM_E_INVALID, "no partial result: call grmkm_build_partial first");
    if (n_ranks != c->part_ranks || n_ranks > 16 || !peer_dst || !peer_word_off)
        return fail(c, GRMKM_E_INVALID, "bad peer export arguments");
    if (!c->part_total) return GRMKM_OK;
    CU_TRY(c, cudaSetDevice(c->device));
    PeerSlices ps{};
    ps.n = n_ranks;
    const uint32_t sub = ceil_log2(std::max(1u, c->part_buckets >> c->cur_bucket_bits));      // virtual buckets per bucket
    const uint64_t B = 1ULL << c->cur_bucket_bits;
    for (uint32_t d = 0; d <= n_ranks; ++d) ps.first[d] = (uint32_t)((B * d / n_ranks) << sub);
    for (uint32_t d = 0; d < n_ranks; ++d) {
        if (c->part_counts[d] && !peer_dst[d]) return fail(c, GRMKM_E_INVALID, "null peer buffer");
        ps.dst[d] = (unsigned long long*)peer_dst[d] + peer_word_off[d];
    }
    k_gather_buckets_peers<<<std::min<uint32_t>(c->part_buckets, (uint32_t)c->sm_count * 8), 256, 0, c->stream>>>(
        (const unsigned long long*)c->ukeys.p, (const unsigned long long*)c->uwords.p, c->part_cap,
        (const unsigned long long*)c->bbase.p, (const unsigned long long*)c->offsets2.p, c->part_buckets, c->part_words, ps);
    CU_TRY(c, cudaGetLastError());
    c->stats.n_launches++;
    return GRMKM_OK;             // asynchronous on the context's stream: the caller's barrier follows on the same stream
}

int grmkm_merge_partials(grmkm_ctx* c, const void* dev_parts, uint32_t n_ranks, uint32_t rank, const uint64_t* src_counts,
                         const uint32_t* src_words, uint32_t total_genomes) {
    if (check_ctx(c)) return GRMKM_E_INVALID;
    if (n_ranks < 1 || n_ranks > 16 || rank >= n_ranks || !src_counts || !src_words)
        return fail(c, GRMKM_E_INVALID, "bad merge arguments");
    CU_TRY(c, cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    Launches L;
    AggParams2 ap{};
    uint64_t n_total = 0, words = 0;
    uint32_t W_total = 0;
    std::vector<uint64_t> h_meta(3 * 16, 0);          // src_off[16], src_count[16], src_width[16] (as u64 / u32 below)
    uint32_t h_width[16] = {0};
    for (uint32_t s = 0; s < n_ranks; ++s) {
        ap.src_off[s] = words; ap.src_words[s] = src_words[s]; ap.src_woff[s] = W_total;
        h_meta[s] = words; h_meta[16 + s] = src_counts[s]; h_width[s] = 1 + src_words[s];
        n_total += src_counts[s]; words += src_counts[s] * (1 + (uint64_t)src_words[s]); W_total += src_words[s];
    }
    ap.n_src = n_ranks;
    if (W_total != (total_genomes + 63) / 64) return fail(c, GRMKM_E_INVALID, "source words do not add up to the genome count");
    if (n_total && !dev_parts) return fail(c, GRMKM_E_INVALID, "null parts");
    c->built = false; c->U = 0; c->W = W_total; c->G = total_genomes;
    c->times = grmkm_times{};
    const uint32_t launches_before = c->stats.n_launches;
    if (n_total == 0) { c->built = true; c->stats.n_kmers = 0; return GRMKM_OK; }
    // the merge table keeps one entry reference per source and slot instead of the words (k_aggregate_cols<3>)
    const uint32_t tw = (n_ranks + 1) & ~1u;
    const size_t slot_bytes = 10 + 4 * (size_t)tw;
    const uint32_t slots = (uint32_t)std::min<size_t>(kAggMaxSlots, (agg_smem_budget(c) - 8) / slot_bytes - kMaxProbe);
    const size_t smem = (((size_t)slots + kMaxProbe) * slot_bytes + 8 + 15) & ~size_t(15);
    for (uint32_t s = 0; s < n_ranks; ++s)
        if (src_counts[s] >= 0xFFFFFFFFULL) return fail(c, GRMKM_E_UNSUPPORTED, "more than 2^32 partial columns from one source");
    // every rank sees 1/P of the hash space: size the buckets for n_total * P entries over the full range
    const uint64_t per = std::max<uint64_t>(1, slots / 2);
    const uint32_t mb = std::min(24u, std::max(6u, ceil_log2((n_total * n_ranks + per - 1) / per)));
    const uint64_t Bfull = 1ULL << mb;
    // this owner's slice of the bucket space (the partial columns of owner r lie in hash range [r/P, (r+1)/P))
    const uint32_t b_lo = (uint32_t)(Bfull * rank / n_ranks);
    const uint32_t b_hi = (uint32_t)((Bfull * (rank + 1) + n_ranks - 1) / n_ranks);
    const uint32_t nb = b_hi - b_lo, nb1 = nb + 1;
    ENSURE(c, c->scalars, S_COUNT * 8);
    ENSURE(c, c->offsets2, (size_t)(nb + 1) * 8);
    ENSURE(c, c->bbase, (size_t)nb * 8);
    ENSURE(c, c->bcounts, (size_t)nb * 8);
    ENSURE(c, c->refs, (size_t)n_ranks * nb1 * 8 + 16 * 8 * 3);      // bounds + source tables
    uint64_t ucap = std::min<uint64_t>(n_total, 0xFFFFFFFFULL);
    ENSURE(c, c->ukeys, ucap * 8);
    ENSURE(c, c->uwords, (size_t)ucap * W_total * 8);
    uint64_t* d_scalars = (uint64_t*)c->scalars.p;
    unsigned long long* d_bounds = (unsigned long long*)c->refs.p;
    unsigned long long* d_meta = d_bounds + (size_t)n_ranks * nb1;
    if (c->ev_ok) cudaEventRecord(c->ev[T_START], st);
    CU_TRY(c, cudaMemsetAsync(c->scalars.p, 0, S_COUNT * 8, st));
    for (uint32_t s = 0; s < 16; ++s) h_meta[32 + s] = 0;
    memcpy(&h_meta[32], h_width, sizeof h_width);
    CU_TRY(c, cudaMemcpyAsync(d_meta, h_meta.data(), 48 * 8, cudaMemcpyHostToDevice, st));
    const unsigned long long* parts = (const unsigned long long*)dev_parts;
    const uint64_t n_search = (uint64_t)n_ranks * nb1;
    k_merge_bounds<<<(uint32_t)((n_search + 255) / 256), 256, 0, st>>>(parts, n_ranks, nb1, b_lo, mb, d_meta, d_meta + 16,
                                                                      (const uint32_t*)(d_meta + 32), d_bounds);
    L.n++;
    CU_TRY(c, cudaGetLastError());
    if (c->ev_ok) cudaEventRecord(c->ev[T_SCATTER], st);
    ap.bucket_bits = mb; ap.row_bits = 0; ap.n_words = W_total; ap.slots = slots;
    ap.keep_singletons = c->cfg.keep_singletons; ap.sub_bits = 0; ap.table_u32 = tw;
    ap.out_keys = (unsigned long long*)c->ukeys.p; ap.out_words = (unsigned long long*)c->uwords.p;
    ap.cap = ucap; ap.scalars = (unsigned long long*)d_scalars; ap.parts = parts; ap.bounds = d_bounds;
    ap.bucket_base = (unsigned long long*)c->bbase.p; ap.bucket_count = (unsigned long long*)c->bcounts.p;
    ap.b_begin = b_lo; ap.b_end = b_hi;
    CU_TRY(c, cudaFuncSetAttribute(k_aggregate_cols<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_aggregate_cols<3><<<std::max(1u, std::min<uint32_t>(nb, (uint32_t)c->sm_count * kAggCtasPerSm)), kAggThreads, smem, st>>>(ap);
    L.n++;
    CU_TRY(c, cudaGetLastError());
    uint64_t sc[S_COUNT];
    CU_TRY(c, cudaMemcpyAsync(sc, d_scalars, sizeof sc, cudaMemcpyDeviceToHost, st));
    CU_TRY(c, cudaStreamSynchronize(st));
    if (c->ev_ok) cudaEventRecord(c->ev[T_AGG], st);
    const uint64_t U = sc[S_U_NEEDED];
    if (c->cfg.flags & GRMKM_FLAG_KMER_ORDER) {
        int r = sort_and_gather(c, U, W_total, ucap, 2 * c->cfg.k, false, L);
        if (r) return r;
    } else {
        ENSURE(c, c->kmers, U * 8);
        ENSURE(c, c->matrix, (size_t)U * W_total * 8);
        k_bucket_offsets<<<1, 1024, 0, st>>>((unsigned long long*)c->bcounts.p, (unsigned long long*)c->offsets2.p, nb,
                                             d_scalars, S_N_SOLID, 1);
        if (U) {
            k_gather_buckets<<<std::max(1u, std::min<uint32_t>(nb, (uint32_t)c->sm_count * 8)), 256, 0, st>>>(
                (const unsigned long long*)c->ukeys.p, (const unsigned long long*)c->uwords.p, ucap,
                (const unsigned long long*)c->bbase.p, (const unsigned long long*)c->offsets2.p, nb, W_total, U,
                (unsigned long long*)c->kmers.p, (unsigned long long*)c->matrix.p, U);
        }
        L.n += 2;
        CU_TRY(c, cudaGetLastError());
    }
    if (c->ev_ok) cudaEventRecord(c->ev[T_SORT], st);
    CU_TRY(c, cudaStreamSynchronize(st));
    c->U = U; c->built = true;
    c->stats.n_kmers = U; c->stats.n_distinct = sc[S_N_DISTINCT]; c->stats.n_words = W_total;
    c->stats.n_genomes = total_genomes; c->stats.n_splits += sc[S_N_SPLITS];
    c->stats.n_launches = launches_before + L.n;
    if (c->ev_ok) {
        cudaEventElapsedTime(&c->times.scatter, c->ev[T_START], c->ev[T_SCATTER]);
        cudaEventElapsedTime(&c->times.aggregate, c->ev[T_SCATTER], c->ev[T_AGG]);
        cudaEventElapsedTime(&c->times.sort, c->ev[T_AGG], c->ev[T_SORT]);
        cudaEventElapsedTime(&c->times.total, c->ev[T_START], c->ev[T_SORT]);
    }
    return GRMKM_OK;
}

}  // extern "C"
