/*
 * grmkm.h -- C ABI of libgrmkm.so: the B200-native k-mer matrix builder that
 * replaces the native half of GRM's "genomes -> genome x k-mer matrix" path.
 *
 * The reference has no in-process FFI for this path: it reaches the native
 * code through argv of child processes.  Each entry point
 * below names the reference interface it stands in for; paths are under
 * /root/reference.
 *
 *   multidsk  argv   bin/kover/core/kover/dataset/tools/kmer_count.py:28-37, 44-53
 *   dsk2kover argv   bin/kover/core/kover/dataset/tools/kmer_pack.py:28-36
 *   Ray       argv   src/app.py:1310 + survey.conf grammar src/app.py:3820-3833
 *   dsk       argv   src/app.py:1371-1372
 *
 * Conventions: extern "C", plain pointers and sizes, no exceptions cross the
 * boundary.  Every call returns GRMKM_OK (0) or a negative GRMKM_E_* code; the
 * message is available from grmkm_last_error().  The caller allocates every
 * output buffer and passes its capacity.  A context owns all device memory and
 * is single-owner (not thread safe); several contexts may coexist.  There is
 * NO CPU fallback: without a CUDA device grmkm_create fails with
 * GRMKM_E_NO_DEVICE.
 */
#ifndef GRMKM_H
#define GRMKM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRMKM_ABI_VERSION 10

enum {
    GRMKM_OK = 0,
    GRMKM_E_INVALID = -1,       /* bad argument / bad state                          */
    GRMKM_E_UNSUPPORTED_K = -2, /* k outside 1..32 (reference allows <=128, kover:114) */
    GRMKM_E_NOMEM = -3,         /* host or device allocation failed                  */
    GRMKM_E_CUDA = -4,          /* CUDA runtime error (message has the detail)       */
    GRMKM_E_IO = -5,            /* file could not be read / inflated                 */
    GRMKM_E_CAPACITY = -6,      /* caller buffer too small                           */
    GRMKM_E_NO_DEVICE = -7,     /* no usable CUDA device: there is no CPU fallback   */
    GRMKM_E_UNSUPPORTED = -8    /* e.g. more than 32768 genomes in one context       */
};

enum { GRMKM_FASTA = 0, GRMKM_FASTQ = 1 };

/* cfg.flags */
#define GRMKM_FLAG_COUNTS 32u /* pooled count table (bin/dsk/dsk -file <list> -kmer-size k, src/app.py:1371-1372): every input is
                                 pooled on ONE row, the result has one column per canonical k-mer whose abundance over all inputs
                                 is >= min_abundance, and matrix row 0 holds that abundance (a count, not presence bits) */

typedef struct grmkm_ctx grmkm_ctx;

/*
 * Parameters = the knobs the reference forwards to the native tools:
 *   k               -kmer-size / -kmer-length / "-k"      (kmer_count.py:31, kmer_pack.py:32, app.py:3821)
 *   min_abundance   -abundance-min  (contigs: 1, kmer_count.py:32; reads: user value, :48)
 *   keep_singletons -filter nothing|singleton  (kover:144-147 -> kmer_pack.py:31); Ray mode: 1
 *   input_kind      contigs (.fna, FASTA) or reads (.fastq, FASTQ)  (create.py:278, 399-402)
 */
typedef struct grmkm_config {
    uint32_t struct_size;     /* = sizeof(grmkm_config), for forward compatibility */
    uint32_t k;               /* 1..32 */
    uint32_t min_abundance;   /* >= 1 */
    uint32_t keep_singletons; /* 0: drop k-mers present in exactly one genome */
    uint32_t input_kind;      /* default kind for grmkm_add_genome_* */
    int32_t device;           /* CUDA ordinal; -1 = current device */
    uint32_t bucket_bits;     /* 0 = auto; log2 of the number of hash buckets */
    uint32_t flags;
    void* stream;             /* cudaStream_t to run on; NULL = context-owned stream */
} grmkm_config;

typedef struct grmkm_stats {
    uint64_t n_bases;       /* nucleotide characters in sequence lines             */
    uint64_t n_windows;     /* valid k-mer windows (= records partitioned)         */
    uint64_t n_input_bytes; /* bytes of FASTA/FASTQ text                           */
    uint64_t n_records;     /* FASTA/FASTQ records                                 */
    uint64_t n_kmers;       /* U: columns kept                                     */
    uint64_t n_distinct;    /* distinct canonical k-mers before the singleton filter */
    uint32_t n_words;       /* ceil(G/64)                                          */
    uint32_t n_genomes;     /* G                                                   */
    uint32_t n_buckets;     /* hash buckets used                                   */
    uint32_t n_launches;    /* kernels launched by the last build                  */
    uint64_t h2d_bytes;     /* host->device bytes moved by the last build          */
    uint64_t device_bytes;  /* device memory held by the context                   */
    uint64_t n_splits;      /* bucket sub-range splits (table overflows handled)   */
    uint64_t n_region_overflows; /* builds redone with exact offsets (a bucket region was too small) */
    uint64_t n_units;        /* super-k-mer units scattered                                */
    uint64_t n_unit_entries; /* distinct (unit, 64-genome block) entries after the dedupe  */
    uint64_t n_wide;         /* [hash, presence word] records the distinct units expand to */
    uint32_t n_unit_buckets; /* content-hash buckets of the unit scatter                   */
    uint32_t n_rounds;       /* abundance builds (min_abundance > 1): rounds of genome rows  */
    uint64_t n_solid_records; /* abundance builds: solid (k-mer, round) presence records     */
} grmkm_stats;

/* per-stage device time of the last build, milliseconds (CUDA events on the build stream) */
typedef struct grmkm_times {
    float h2d, parse, pack, count, scatter, abundance, aggregate, sort, total;
    float bounds, dedupe, expand; /* unit path: run boundaries, per-bucket dedupe, expansion to k-mer records */
} grmkm_times;

int grmkm_abi_version(void);
int grmkm_device_count(void);

int grmkm_create(const grmkm_config* cfg, grmkm_ctx** out);
void grmkm_destroy(grmkm_ctx* ctx);
/* ctx may be NULL: message of the last failed grmkm_create on this thread */
const char* grmkm_last_error(const grmkm_ctx* ctx);

/* Forget the inputs (and result) of the previous build; device buffers are kept for reuse. */
int grmkm_reset(grmkm_ctx* ctx);

/*
 * One line of multidsk's "-file" list (create.py:361-362, 482-489): the file(s)
 * of genome `genome_row` (row order = order of genome_identifiers, create.py:334-347).
 * All inputs of one row are pooled.  *_bytes: host memory, borrowed until the
 * build returns.  *_device: device memory on cfg.device, borrowed likewise.
 * *_files: read (and gunzip'ed when the name ends in .gz) by the library.
 */
int grmkm_add_genome_bytes(grmkm_ctx* ctx, uint32_t genome_row, const uint8_t* data, uint64_t n);
int grmkm_add_genome_device(grmkm_ctx* ctx, uint32_t genome_row, const void* dev_data, uint64_t n);
int grmkm_add_genome_files(grmkm_ctx* ctx, uint32_t genome_row, const char* const* paths, int n_paths);
/* The whole "-file" list in one call (the n lines of create.py:361-362 at once): input i belongs to row rows[i],
 * is lens[i] bytes at data[i]; on_device = 0 host memory (as *_bytes), 1 device memory (as *_device). */
int grmkm_add_genomes(grmkm_ctx* ctx, uint32_t n, const uint32_t* rows, const void* const* data, const uint64_t* lens,
                      int on_device);
/* Declare rows without input (trailing empty genomes still get a matrix row). */
int grmkm_set_genome_count(grmkm_ctx* ctx, uint32_t n_genomes);

/*
 * multidsk + dsk2kover in one call: count canonical k-mers per genome, apply
 * min_abundance per genome, merge across genomes, apply the singleton filter,
 * pack 64 genomes per word (bit 63-(g%64) of word g/64, utils.py:144-154).
 * The result stays device-resident until the next reset/build/destroy.
 * min_abundance > 1 (kmer_count.py:48): the genomes are counted a few rows at a time (per-genome abundance
 * counters in the shared-memory tables), the solid presence of every round is kept, and one presence merge
 * over all rounds follows.  GRMKM_FLAG_COUNTS: see the flag.
 */
int grmkm_build(grmkm_ctx* ctx);

int grmkm_dims(const grmkm_ctx* ctx, uint64_t* n_kmers, uint32_t* n_words, uint32_t* n_genomes);
int grmkm_get_stats(const grmkm_ctx* ctx, grmkm_stats* out);
int grmkm_stage_times(const grmkm_ctx* ctx, grmkm_times* out);

/* canonical k-mers as integers (A0 C1 T2 G3, first base most significant); cap in elements.
 * Column order: ascending grmkm_hash64(k-mer) = k-mer * 0x9E3779B97F4A7C15 mod 2^64 (the order the hash
 * partitions come out in, identical for any GPU count; the reference's own column order is the unspecified
 * partition order of DSK). */
int grmkm_copy_kmers_packed(grmkm_ctx* ctx, uint64_t* dst, uint64_t cap);
/* kmer_sequences: U x k bytes, upper-case, no terminator (create.py:216-220, ds.py:84); cap in bytes */
int grmkm_copy_kmer_strings(grmkm_ctx* ctx, char* dst, uint64_t cap);
/* kmer_matrix: row-major n_words x U (create.py:224-230); cap in elements */
int grmkm_copy_matrix(grmkm_ctx* ctx, uint64_t* dst, uint64_t cap);
/*
 * Ray Surveyor KmerMatrix.tsv (consumer grammar create.py:121-137,241-264):
 * "kmers\t<name_1>...\n" then fixed-width rows "<kmer>\t<0|1>...\n".
 * dst may be NULL to query the size (returned in *written with GRMKM_E_CAPACITY).
 */
int grmkm_format_tsv(grmkm_ctx* ctx, const char* const* names, char* dst, uint64_t cap, uint64_t* written);

/*
 * The result in page-locked host memory owned by the context: one device->host copy at PCIe speed (a copy
 * into pageable caller memory runs 10x slower).  kmers[U], matrix[n_words][U] row-major; the pointers stay
 * valid until the next build / reset / destroy of this context.
 */
int grmkm_host_result(grmkm_ctx* ctx, const uint64_t** kmers, const uint64_t** matrix);

/*
 * Kernels over the finished, device-resident result.
 *
 * grmkm_result_checksum: order-independent 128-bit digest of the columns (k-mer + its words); the digests of the
 *   ranks' slices of a multi-GPU build add up (mod 2^64 per lane) to the digest of the one-GPU matrix.
 * grmkm_sum_rows: KmerRuleClassifications.sum_rows (bin/kover/core/kover/learning/common/rules.py:201-267 with
 *   popcount.pyx:76-95): dst[j] = sum over word rows w of popcount(matrix[w][j] & row_mask[w]); row_mask has
 *   n_words words, example (genome row) g at bit 63-(g%64) of word g/64 (build_row_mask, rules.py:209-222).
 * grmkm_gram: G x G matrix of shared columns, dst[a*G+b] = #columns present in genomes a and b (the similarity
 *   matrix Ray Surveyor prints beside the k-mer matrix, src/app.py:1310).
 */
int grmkm_result_checksum(grmkm_ctx* ctx, uint64_t out[2]);
int grmkm_sum_rows(grmkm_ctx* ctx, const uint64_t* row_mask, uint32_t n_mask_words, uint32_t* dst, uint64_t cap);
int grmkm_gram(grmkm_ctx* ctx, uint64_t* dst, uint64_t cap);

/*
 * from_tsv's bit packer on the GPU (dataset/create.py:241-271 with utils.py:133-156): body = n_rows fixed-width
 * rows "<k-mer>\t<c_0>\t<c_1>...\n" of a Ray Surveyor KmerMatrix.tsv (host memory, row_width = k + 2 x
 * n_tsv_cols + 1); matrix row g takes TSV column sel[g]; dst = row-major ceil(n_genomes/64) x n_rows words
 * (host memory).  A cell other than '0' / '1' fails with GRMKM_E_INVALID (create.py:121-137: binary matrix).
 * Needs no build; any context of the device will do.
 */
int grmkm_tsv_pack(grmkm_ctx* ctx, const uint8_t* body, uint64_t n_rows, uint32_t row_width, uint32_t k,
                   uint32_t n_tsv_cols, uint32_t n_genomes, const uint32_t* sel, uint64_t* dst, uint64_t cap);

/* Device pointers of the result for callers that stay on the GPU: kmers[U] and the matrix, whose word row w starts at
 * d_matrix + w * pitch_words (pitch_words >= U: the columns are written at their final place while their number is
 * still unknown, so the rows are spaced by the build's capacity; the host copies above deliver rows U words apart). */
int grmkm_device_result(const grmkm_ctx* ctx, const uint64_t** d_kmers, const uint64_t** d_matrix, uint64_t* pitch_words);

/*
 * Bench/test utility: materialise the synthetic FASTA of BASELINE.md section 4 /
 * SURVEY.md section 8d on the device.  layout = host table produced by
 * genomic-resistance-mapping-grm-_b200/synth.py (see grmkm_synth_genome there).
 */
int grmkm_synth_fasta_device(grmkm_ctx* ctx, const void* layout, uint64_t layout_bytes, void* dev_dst,
                             uint64_t dst_bytes);

/*
 * Multi-GPU (one context per process/GPU).  The exchange is either fused into the export
 * (grmkm_export_partials_peers: NVLink stores into the owners' receive buffers) or done by the host layer with
 * an all-to-all over NCCL on the buffer grmkm_export_partials fills.
 *
 * grmkm_build_partial: local stages only.  Produces partial columns
 * (hashed k-mer, n_local_words words) grouped by owner rank = hash range,
 * owner r holding buckets [r*B/P, (r+1)*B/P).  counts[P] receives the number of
 * partial columns per owner.  Export copies them, owner-major, as
 * u64 records of (1 + n_local_words) words into caller device memory.
 *
 * grmkm_merge_partials: owner side.  parts = device buffer with the received
 * partial columns of all P sources back to back (source-major), src_counts[P]
 * their lengths, src_words[P] the words per source; source s contributes matrix
 * word rows [word_offset_s, word_offset_s + src_words[s]).  Applies the
 * singleton filter on the full popcount and leaves this owner's slice of the
 * result in the context (dims/copy_* as after grmkm_build).
 */
int grmkm_build_partial(grmkm_ctx* ctx, uint32_t n_ranks, uint64_t* counts);
/* All ranks must partition the hash space identically: read this rank's automatic choice for
 * the inputs added so far, agree on the maximum over ranks, set it before grmkm_build_partial (which refuses
 * n_ranks > 1 without it) and keep it set for grmkm_merge_partials: the owners' hash ranges are cut on that grid. */
int grmkm_plan_bucket_bits(grmkm_ctx* ctx, uint32_t* bits);
int grmkm_set_bucket_bits(grmkm_ctx* ctx, uint32_t bits);
int grmkm_export_partials(grmkm_ctx* ctx, void* dev_dst, uint64_t dst_bytes);
/* The export fused with the all-to-all: owner d's slice of the partial columns is stored straight into rank d's
 * receive buffer (peer_dst[d], peer-accessible device memory, e.g. torch symmetric memory) at word offset
 * peer_word_off[d] -- over NVLink, no send buffer and no NCCL send/recv (Ray's message routing, src/app.py:1310).
 * Asynchronous on the context's stream; the ranks synchronise (barrier) before anybody merges. */
int grmkm_export_partials_peers(grmkm_ctx* ctx, uint32_t n_ranks, void* const* peer_dst, const uint64_t* peer_word_off);
/* This context's rank among the exporters (default: its device index).  Only staggers the order in which the export walks
 * the owners -- rank r starts with owner r + 1 -- so that the P exporters store into P different receivers at any moment
 * (the hash-routed message rounds of `mpiexec -n 4 Ray`, src/app.py:1310); the result does not depend on it. */
int grmkm_set_exchange_rank(grmkm_ctx* ctx, uint32_t rank);
int grmkm_merge_partials(grmkm_ctx* ctx, const void* dev_parts, uint32_t n_ranks, uint32_t rank,
                         const uint64_t* src_counts, const uint32_t* src_words, uint32_t total_genomes);

#ifdef __cplusplus
}
#endif
#endif /* GRMKM_H */
