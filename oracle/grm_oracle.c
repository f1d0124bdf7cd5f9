/*
 * grm_oracle.c -- CPU restatement ORACLE of GRM's k-mer matrix path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it
 * may be used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs as the checker or the timed CPU arm, never by the
 * shipped CUDA path (which fails loudly when libgrmkm.so is missing).
 *
 * What it restates (reference = /root/reference):
 *   - multidsk  (bin/kover/core/kover/dataset/tools/kmer_count.py:23-53 is the
 *     only in-tree trace: argv of the missing binary).  Per genome: the
 *     multiset of canonical k-mers of every record, keep count >= abundance-min
 *     (contigs hard-code 1, kmer_count.py:32; reads pass the user value, :48).
 *   - dsk2kover (tools/kmer_pack.py:23-39, binary missing): G-way merge of the
 *     per-genome sorted solid lists, one presence bit per (genome, k-mer),
 *     "-filter singleton" drops k-mers present in exactly one genome,
 *     64 genomes per uint64 word.
 *   - bit layout: bin/kover/core/kover/utils.py:133-156
 *     (_pack_binary_bytes_to_ints): genome row g -> word g/64, bit 63-(g%64).
 *   - Ray Surveyor TSV as required by its only in-tree consumer,
 *     dataset/create.py:121-137,174-175,241-264.
 *
 * PARITY STATUS: the counting arithmetic lives in GATB-core 1.4.2 / DSK /
 * Kover kmer_tools / Ray, none of which are in the reference checkout
 * (.MISSING_LARGE_BLOBS:1-4; extern.txt:5) -> "parity unpinned" for the
 * k-mer semantics (GATB base code A0 C1 T2 G3, canonical = integer min).  The
 * bit layout, dtype rule, metadata ordering and text grammars ARE pinned by
 * golden vectors generated from the reference's own Python code
 * (tests/golden/make_golden.py).
 *
 * Plain C11, no dependencies beyond libc (+ OpenMP when compiled -fopenmp).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GRMO_FASTA 0
#define GRMO_FASTQ 1

typedef struct grmo_result {
    uint64_t n_kmers;   /* U: number of columns                      */
    uint32_t n_words;   /* ceil(G/64)                                 */
    uint32_t n_genomes; /* G                                          */
    uint64_t n_bases;   /* nucleotide characters seen in sequence lines */
    uint64_t n_windows; /* valid k-mer windows over all genomes       */
    uint64_t* kmers;    /* [U] canonical k-mers, ascending            */
    uint64_t* matrix;   /* [n_words][U] row-major                     */
} grmo_result;

/* ---- semantics E4/E5 ------------------------------ */
static inline int base_code(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'T': case 't': return 2;
        case 'G': case 'g': return 3;
        default: return -1;
    }
}

typedef struct kvec { uint64_t* v; size_t n, cap; } kvec;
static int kvec_push(kvec* a, uint64_t x) {
    if (a->n == a->cap) {
        size_t nc = a->cap ? a->cap * 2 : 1 << 16;
        uint64_t* p = (uint64_t*)realloc(a->v, nc * sizeof(uint64_t));
        if (!p) return -1;
        a->v = p; a->cap = nc;
    }
    a->v[a->n++] = x;
    return 0;
}

/* rolling canonical k-mer extractor; state survives across calls for one record */
typedef struct roll { uint64_t fw, rc, mask; int k, run; } roll;
static inline void roll_reset(roll* r) { r->fw = r->rc = 0; r->run = 0; }
static inline int roll_feed(roll* r, uint8_t c, kvec* out, uint64_t* n_bases) {
    int code = base_code(c);
    (*n_bases)++;
    if (code < 0) { roll_reset(r); return 0; }
    r->fw = ((r->fw << 2) | (uint64_t)code) & r->mask;
    r->rc = (r->rc >> 2) | ((uint64_t)(code ^ 2) << (2 * (r->k - 1)));
    if (r->run < r->k) r->run++;
    if (r->run >= r->k) return kvec_push(out, r->fw < r->rc ? r->fw : r->rc);
    return 0;
}

/*
 * All canonical k-mer windows (with multiplicity) of one file, appended to out.
 * FASTA (E2): bytes before the first '>' at a line start are ignored; a line
 * starting with '>' opens a new record; every other line is sequence; '\r'
 * bytes are dropped; empty lines are ignored; no final newline needed.
 * FASTQ (E3): from the first '@' at a line start, 4 lines per record, only
 * line 2 is sequence.
 */
static int extract_file(const uint8_t* d, size_t n, int kind, int k, kvec* out,
                        uint64_t* n_bases) {
    roll r; r.k = k; r.mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    roll_reset(&r);
    size_t i = 0;
    uint8_t first = (kind == GRMO_FASTQ) ? '@' : '>';
    /* skip to first header at a line start */
    int at_ls = 1;
    for (; i < n; i++) {
        if (at_ls && d[i] == first) break;
        at_ls = (d[i] == '\n');
    }
    if (i >= n) return 0;
    if (kind == GRMO_FASTA) {
        int state = 0; /* 0 line start, 1 header, 2 sequence */
        for (; i < n; i++) {
            uint8_t c = d[i];
            if (c == '\n') { state = 0; continue; }
            if (state == 0) {
                if (c == '>') { state = 1; roll_reset(&r); continue; }
                state = 2;
            }
            if (state == 1) continue;
            if (c == '\r') continue;
            if (roll_feed(&r, c, out, n_bases)) return -1;
        }
    } else {
        uint64_t line = 0;
        int at_start = 1;
        for (; i < n; i++) {
            uint8_t c = d[i];
            if (c == '\n') { line++; at_start = 1; continue; }
            if ((line & 3) == 0) { if (at_start) roll_reset(&r); at_start = 0; continue; }
            at_start = 0;
            if ((line & 3) != 1) continue;
            if (c == '\r') continue;
            if (roll_feed(&r, c, out, n_bases)) return -1;
        }
    }
    return 0;
}

/* LSD radix sort of u64 keys, 8 bits per pass, skipping constant digits */
static int radix_sort_u64(uint64_t* a, size_t n) {
    if (n < 2) return 0;
    uint64_t* b = (uint64_t*)malloc(n * sizeof(uint64_t));
    if (!b) return -1;
    uint64_t* src = a; uint64_t* dst = b;
    for (int pass = 0; pass < 8; pass++) {
        size_t cnt[256]; memset(cnt, 0, sizeof cnt);
        int sh = pass * 8;
        for (size_t i = 0; i < n; i++) cnt[(src[i] >> sh) & 255]++;
        int trivial = 0;
        for (int d = 0; d < 256; d++) if (cnt[d] == n) { trivial = 1; break; }
        if (trivial) continue;
        size_t s = 0;
        for (int d = 0; d < 256; d++) { size_t c = cnt[d]; cnt[d] = s; s += c; }
        for (size_t i = 0; i < n; i++) dst[cnt[(src[i] >> sh) & 255]++] = src[i];
        uint64_t* t = src; src = dst; dst = t;
    }
    if (src != a) memcpy(a, src, n * sizeof(uint64_t));
    free(b);
    return 0;
}

/* sort + run-length + abundance filter, in place; returns #solid k-mers */
static size_t solid_inplace(uint64_t* a, size_t n, uint32_t min_ab, uint32_t* counts) {
    size_t w = 0, i = 0;
    while (i < n) {
        size_t j = i + 1;
        while (j < n && a[j] == a[i]) j++;
        uint64_t c = j - i;
        if (c >= min_ab) {
            a[w] = a[i];
            if (counts) counts[w] = c > 0xFFFFFFFFu ? 0xFFFFFFFFu : (uint32_t)c;
            w++;
        }
        i = j;
    }
    return w;
}

/*
 * Per-genome solid canonical k-mers (what multidsk leaves in <genome>.h5).
 * Returns count, fills *out_kmers (malloc'd, ascending) and optional *out_counts.
 */
int64_t grmo_genome_solid(const uint8_t* const* datas, const uint64_t* lens, const int32_t* kinds,
                          int n_files, int k, uint32_t min_abundance,
                          uint64_t** out_kmers, uint32_t** out_counts,
                          uint64_t* n_bases, uint64_t* n_windows) {
    if (k < 1 || k > 32) return -2;
    kvec v = {0, 0, 0};
    uint64_t nb = 0;
    for (int f = 0; f < n_files; f++)
        if (extract_file(datas[f], (size_t)lens[f], kinds[f], k, &v, &nb)) { free(v.v); return -1; }
    if (n_windows) *n_windows = v.n;
    if (n_bases) *n_bases = nb;
    if (radix_sort_u64(v.v, v.n)) { free(v.v); return -1; }
    uint32_t* cnt = NULL;
    if (out_counts) { cnt = (uint32_t*)malloc((v.n ? v.n : 1) * sizeof(uint32_t)); if (!cnt) { free(v.v); return -1; } }
    size_t m = solid_inplace(v.v, v.n, min_abundance ? min_abundance : 1, cnt);
    *out_kmers = v.v;
    if (out_counts) *out_counts = cnt;
    return (int64_t)m;
}

void grmo_free(void* p) { free(p); }

/* ---- dsk2kover restatement: G-way heap merge + bit packing ---------------- */
typedef struct hnode { uint64_t key; uint32_t g; } hnode;
static inline int hless(hnode a, hnode b) { return a.key < b.key || (a.key == b.key && a.g < b.g); }
static void sift_down(hnode* h, size_t n, size_t i) {
    for (;;) {
        size_t l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && hless(h[l], h[m])) m = l;
        if (r < n && hless(h[r], h[m])) m = r;
        if (m == i) return;
        hnode t = h[i]; h[i] = h[m]; h[m] = t; i = m;
    }
}

/* merge restricted to keys in [lo, hi]; appends columns; returns 0 / -1 */
static int merge_range(uint64_t* const* lists, const size_t* begin, const size_t* end, uint32_t G,
                       uint32_t W, int keep_singletons, kvec* out_k, kvec* out_w) {
    hnode* heap = (hnode*)malloc((G ? G : 1) * sizeof(hnode));
    size_t* pos = (size_t*)malloc((G ? G : 1) * sizeof(size_t));
    uint64_t* col = (uint64_t*)malloc(W * sizeof(uint64_t));
    if (!heap || !pos || !col) { free(heap); free(pos); free(col); return -1; }
    size_t hn = 0;
    for (uint32_t g = 0; g < G; g++) {
        pos[g] = begin[g];
        if (pos[g] < end[g]) { heap[hn].key = lists[g][pos[g]]; heap[hn].g = g; hn++; }
    }
    for (size_t i = hn / 2; i-- > 0;) sift_down(heap, hn, i);
    int rc = 0;
    while (hn) {
        uint64_t key = heap[0].key;
        memset(col, 0, W * sizeof(uint64_t));
        uint32_t present = 0;
        while (hn && heap[0].key == key) {
            uint32_t g = heap[0].g;
            col[g >> 6] |= 1ULL << (63 - (g & 63));   /* utils.py:144-154 */
            present++;
            if (++pos[g] < end[g]) heap[0].key = lists[g][pos[g]];
            else heap[0] = heap[--hn];
            if (hn) sift_down(heap, hn, 0);
        }
        if (present >= 2 || keep_singletons) {
            if (kvec_push(out_k, key)) { rc = -1; break; }
            for (uint32_t w = 0; w < W; w++) if (kvec_push(out_w, col[w])) { rc = -1; break; }
            if (rc) break;
        }
    }
    free(heap); free(pos); free(col);
    return rc;
}

static size_t lower_bound(const uint64_t* a, size_t n, uint64_t x) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t m = (lo + hi) / 2; if (a[m] < x) lo = m + 1; else hi = m; }
    return lo;
}

/*
 * Full build.  files f = 0..n_files-1 belong to genome row rows[f] (0..G-1).
 * threads <= 0 -> all available.  Result freed with grmo_result_free.
 */
int grmo_build(const uint8_t* const* datas, const uint64_t* lens, const uint32_t* rows,
               const int32_t* kinds, int n_files, uint32_t n_genomes, int k,
               uint32_t min_abundance, int keep_singletons, int threads, grmo_result** out) {
    if (k < 1 || k > 32 || !out) return -2;
    uint32_t G = n_genomes, W = (G + 63) / 64;
    uint64_t** lists = (uint64_t**)calloc(G ? G : 1, sizeof(uint64_t*));
    size_t* ln = (size_t*)calloc(G ? G : 1, sizeof(size_t));
    uint64_t* gb = (uint64_t*)calloc(G ? G : 1, sizeof(uint64_t));
    uint64_t* gw = (uint64_t*)calloc(G ? G : 1, sizeof(uint64_t));
    int err = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    int T = omp_get_max_threads();
#else
    int T = 1; (void)threads;
#endif
    /* hot loop A: one counting run per genome (multidsk) */
    #pragma omp parallel for schedule(dynamic, 1)
    for (int64_t g = 0; g < (int64_t)G; g++) {
        const uint8_t** fd = (const uint8_t**)malloc((n_files ? n_files : 1) * sizeof(void*));
        uint64_t* fl = (uint64_t*)malloc((n_files ? n_files : 1) * sizeof(uint64_t));
        int32_t* fk = (int32_t*)malloc((n_files ? n_files : 1) * sizeof(int32_t));
        int nf = 0;
        for (int f = 0; f < n_files; f++) if (rows[f] == (uint32_t)g) { fd[nf] = datas[f]; fl[nf] = lens[f]; fk[nf] = kinds[f]; nf++; }
        int64_t m = grmo_genome_solid(fd, fl, fk, nf, k, min_abundance, &lists[g], NULL, &gb[g], &gw[g]);
        if (m < 0) {
            #pragma omp atomic write
            err = 1;
        } else ln[g] = (size_t)m;
        free(fd); free(fl); free(fk);
    }
    grmo_result* R = (grmo_result*)calloc(1, sizeof(grmo_result));
    if (err || !R) goto fail;
    R->n_genomes = G; R->n_words = W;
    for (uint32_t g = 0; g < G; g++) { R->n_bases += gb[g]; R->n_windows += gw[g]; }
    {
        /* hot loop B: G-way merge (dsk2kover), split into T key ranges by the
         * quantiles of the longest list so host threads can share the work */
        uint32_t gl = 0;
        for (uint32_t g = 1; g < G; g++) if (ln[g] > ln[gl]) gl = g;
        int P = (G && ln[gl] >= (size_t)T * 4) ? T : 1;
        uint64_t* split = (uint64_t*)malloc((P + 1) * sizeof(uint64_t));
        kvec* pk = (kvec*)calloc(P, sizeof(kvec));
        kvec* pw = (kvec*)calloc(P, sizeof(kvec));
        for (int p = 1; p < P; p++) split[p] = lists[gl][ln[gl] * (size_t)p / P];
        #pragma omp parallel for schedule(dynamic, 1)
        for (int p = 0; p < P; p++) {
            size_t* b = (size_t*)malloc((G ? G : 1) * sizeof(size_t));
            size_t* e = (size_t*)malloc((G ? G : 1) * sizeof(size_t));
            for (uint32_t g = 0; g < G; g++) {
                b[g] = (p == 0) ? 0 : lower_bound(lists[g], ln[g], split[p]);
                e[g] = (p == P - 1) ? ln[g] : lower_bound(lists[g], ln[g], split[p + 1]);
            }
            if (merge_range(lists, b, e, G, W, keep_singletons, &pk[p], &pw[p])) {
                #pragma omp atomic write
                err = 1;
            }
            free(b); free(e);
        }
        uint64_t U = 0;
        for (int p = 0; p < P; p++) U += pk[p].n;
        R->n_kmers = U;
        R->kmers = (uint64_t*)malloc((U ? U : 1) * sizeof(uint64_t));
        R->matrix = (uint64_t*)calloc((size_t)(U ? U : 1) * (W ? W : 1), sizeof(uint64_t));
        if (!R->kmers || !R->matrix) err = 1;
        uint64_t off = 0;
        for (int p = 0; p < P && !err; p++) {
            memcpy(R->kmers + off, pk[p].v, pk[p].n * sizeof(uint64_t));
            for (size_t c = 0; c < pk[p].n; c++)
                for (uint32_t w = 0; w < W; w++)
                    R->matrix[(size_t)w * U + off + c] = pw[p].v[c * W + w];
            off += pk[p].n;
        }
        for (int p = 0; p < P; p++) { free(pk[p].v); free(pw[p].v); }
        free(pk); free(pw); free(split);
    }
    if (err) goto fail;
    for (uint32_t g = 0; g < G; g++) free(lists[g]);
    free(lists); free(ln); free(gb); free(gw);
    *out = R;
    return 0;
fail:
    for (uint32_t g = 0; g < G; g++) free(lists[g]);
    free(lists); free(ln); free(gb); free(gw);
    if (R) { free(R->kmers); free(R->matrix); free(R); }
    return -1;
}

void grmo_result_free(grmo_result* r) {
    if (!r) return;
    free(r->kmers); free(r->matrix); free(r);
}

/* canonical k-mer integer -> text, GATB order "ACTG", most significant base first (E5) */
void grmo_kmer_string(uint64_t x, int k, char* dst) {
    static const char L[4] = {'A', 'C', 'T', 'G'};
    for (int i = 0; i < k; i++) dst[i] = L[(x >> (2 * (k - 1 - i))) & 3];
}

/*
 * Ray Surveyor style TSV (Appendix B): "kmers\t<name>...\n" then one fixed
 * width row per column: k chars, then "\t0" / "\t1" per genome, "\n".
 * Returns bytes needed; writes only when cap is large enough.
 */
uint64_t grmo_format_tsv(const grmo_result* r, int k, const char* const* names, char* dst, uint64_t cap) {
    uint64_t G = r->n_genomes, U = r->n_kmers;
    uint64_t hdr = 5;
    for (uint64_t g = 0; g < G; g++) hdr += 1 + strlen(names[g]);
    hdr += 1;
    uint64_t roww = (uint64_t)k + 2 * G + 1;
    uint64_t need = hdr + roww * U;
    if (!dst || cap < need) return need;
    char* p = dst;
    memcpy(p, "kmers", 5); p += 5;
    for (uint64_t g = 0; g < G; g++) { *p++ = '\t'; size_t l = strlen(names[g]); memcpy(p, names[g], l); p += l; }
    *p++ = '\n';
    #pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)U; c++) {
        char* q = p + (uint64_t)c * roww;
        grmo_kmer_string(r->kmers[c], k, q); q += k;
        for (uint64_t g = 0; g < G; g++) {
            uint64_t w = r->matrix[(g >> 6) * U + (uint64_t)c];
            *q++ = '\t';
            *q++ = ((w >> (63 - (g & 63))) & 1) ? '1' : '0';
        }
        *q = '\n';
    }
    return need;
}

int grmo_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
