/*
 * oracle/oligo_oracle.c — CPU restatement of kmertools' oligonucleotide-frequency-vector path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (kmertools_b200/, pykmertools/, the C-ABI
 * library) may import, link or execute this file.  It is called only from tests/, from
 * __graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs, as the checker
 * and as the timed CPU stand-in for the reference.
 *
 * The reference is Rust (no rustc/cargo in this image) so it cannot be compiled into oracle/_ref;
 * this file restates its algorithm function by function.  Parity is PINNED: tests/test_oracle.py
 * checks it byte-for-byte against the reference's own golden files (tests/golden/, copied from
 * /root/reference/test_data) and against every unit KAT of the reference's Rust tests.
 *
 * Each function cites the reference file:line it follows (paths relative to the reference root).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* kmer/src/kmer.rs:6-15 (duplicate kmer/src/lib.rs:7-16): byte -> 2-bit code, 4 = ambiguous.
 * Built at load time from the same rule instead of a literal table: bytes 0..3 -> 0..3,
 * A/a->0, C/c->1, G/g->2, T/t/U/u->3, everything else 4. */
static uint8_t NT4[256];
static int nt4_ready = 0;

/* Runs at load time (constructor) so the table is complete before any OpenMP region can race on it:
 * a lazy first-use init inside the parallel batch driver let one thread re-memset the table while
 * another was decoding with it (found in round 1: first oracle call of a process under-counted). */
__attribute__((constructor)) static void nt4_init(void) {
    if (nt4_ready) return;
    memset(NT4, 4, sizeof NT4);
    NT4[0] = 0; NT4[1] = 1; NT4[2] = 2; NT4[3] = 3;
    NT4['A'] = NT4['a'] = 0;
    NT4['C'] = NT4['c'] = 1;
    NT4['G'] = NT4['g'] = 2;
    NT4['T'] = NT4['t'] = 3;
    NT4['U'] = NT4['u'] = 3;
    nt4_ready = 1;
}

uint8_t ktb_oracle_nt4(uint8_t b) { nt4_init(); return NT4[b]; }

/* kmer/src/kmer.rs:43-52 KmerGenerator::rev_comp */
uint64_t ktb_oracle_rev_comp(uint64_t kmer, int k) {
    uint64_t r = 0;
    for (int i = 0; i < k; i++) {
        r <<= 2;
        r |= (kmer & 3) ^ 3;
        kmer >>= 2;
    }
    return r;
}

/* kmer/src/kmer.rs:80-106 KmerGenerator::next, run to exhaustion.  Writes the (f,r) pairs in order,
 * returns how many were emitted.  fout/rout may be NULL (count only).  State follows kmer.rs:30-41. */
uint64_t ktb_oracle_kmers(const uint8_t *seq, uint64_t len, int k, uint64_t *fout, uint64_t *rout) {
    nt4_init();
    uint64_t fval = 0, rval = 0, n = 0;
    uint64_t run = 0;
    const uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    const uint64_t shift = 2 * (uint64_t)(k - 1);
    for (uint64_t pos = 0; pos < len; pos++) {
        uint64_t c = NT4[seq[pos]];
        if (c < 4) {
            fval = ((fval << 2) | c) & mask;
            rval = (rval >> 2) | ((c ^ 3) << shift);
            run += 1;
        } else {
            run = 0; /* fval/rval deliberately not cleared, like the reference */
        }
        if (run == (uint64_t)k) {
            run -= 1;
            if (fout) fout[n] = fval;
            if (rout) rout[n] = rval;
            n++;
        }
    }
    return n;
}

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}

/* kmer/src/kmer.rs:54-73 KmerGenerator::kmer_pos_maps.
 * pos_map[4^k]: canonical code -> rank, 0 in every non-canonical slot (pinned by pos_map_test,
 * kmer.rs:156-176).  pos_to_kmer[count]: rank -> canonical code.  Returns count.
 * The reference collects min(x, rc(x)) into a HashSet and sorts; a sort + unique is the same set. */
uint64_t ktb_oracle_kmer_pos_maps(int k, uint64_t *pos_map, uint64_t *pos_to_kmer) {
    const uint64_t n = 1ULL << (2 * k);
    uint64_t *mins = (uint64_t *)malloc(n * sizeof(uint64_t));
    for (uint64_t x = 0; x < n; x++) {
        uint64_t rc = ktb_oracle_rev_comp(x, k);
        mins[x] = x < rc ? x : rc;
    }
    qsort(mins, n, sizeof(uint64_t), cmp_u64);
    uint64_t count = 0;
    for (uint64_t i = 0; i < n; i++)
        if (i == 0 || mins[i] != mins[i - 1]) mins[count++] = mins[i];
    if (pos_map) memset(pos_map, 0, n * sizeof(uint64_t));
    for (uint64_t pos = 0; pos < count; pos++) {
        if (pos_map) pos_map[mins[pos]] = pos;
        if (pos_to_kmer) pos_to_kmer[pos] = mins[pos];
    }
    free(mins);
    return count;
}

/* Output width: canonical -> count from kmer_pos_maps, raw -> 4^k (composition/src/oligo.rs:232-236) */
uint64_t ktb_oracle_dim(int k, int canonical) {
    if (!canonical) return 1ULL << (2 * k);
    return ktb_oracle_kmer_pos_maps(k, NULL, NULL);
}

/* kmer/src/lib.rs:19-34 numeric_to_kmer: k chars + NUL into out */
void ktb_oracle_numeric_to_kmer(uint64_t kmer, int k, char *out) {
    static const char L[4] = {'A', 'C', 'G', 'T'};
    for (int i = k - 1; i >= 0; i--) {
        out[i] = L[kmer & 3];
        kmer >>= 2;
    }
    out[k] = 0;
}

/* kmer/src/lib.rs:36-50 kmer_to_numeric */
void ktb_oracle_kmer_to_numeric(const char *kmer, uint64_t *f, uint64_t *r) {
    nt4_init();
    size_t k = strlen(kmer);
    uint64_t fval = 0, rval = 0;
    const uint64_t shift = 2 * (uint64_t)(k - 1);
    const uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
    for (size_t i = 0; i < k; i++) {
        uint64_t c = NT4[(uint8_t)kmer[i]];
        fval = ((fval << 2) | c) & mask;
        rval = (rval >> 2) | ((c ^ 3) << shift);
    }
    *f = fval;
    *r = rval;
}

/* norm_mode: 0 = counts, 1 = CLI normalisation (composition/src/oligo.rs:231-259: total += 1 per
 * k-mer in both modes), 2 = pybindings normalisation (pybindings/src/oligo.rs:39-69: raw mode adds
 * 2 to total per k-mer, line 61, so normalised raw vectors sum to 0.5).
 * out has ktb_oracle_dim(k, canonical) doubles.  Returns the number of valid windows. */
uint64_t ktb_oracle_vectorise_one(const uint8_t *seq, uint64_t len, int k, const uint64_t *pos_map,
                                  int canonical, int norm_mode, double *out, uint64_t dim) {
    nt4_init();
    for (uint64_t i = 0; i < dim; i++) out[i] = 0.0;
    double total = 0.0;
    uint64_t fval = 0, rval = 0, run = 0, nk = 0;
    const uint64_t mask = (1ULL << (2 * k)) - 1;
    const uint64_t shift = 2 * (uint64_t)(k - 1);
    const double step = (norm_mode == 2 && !canonical) ? 2.0 : 1.0;
    for (uint64_t pos = 0; pos < len; pos++) {
        uint64_t c = NT4[seq[pos]];
        if (c < 4) {
            fval = ((fval << 2) | c) & mask;
            rval = (rval >> 2) | ((c ^ 3) << shift);
            run += 1;
        } else {
            run = 0;
        }
        if (run == (uint64_t)k) {
            run -= 1;
            if (canonical) {
                uint64_t m = fval < rval ? fval : rval;
                out[pos_map[m]] += 1.0;
            } else {
                out[fval] += 1.0;
            }
            total += step;
            nk++;
        }
    }
    if (norm_mode != 0) {
        const double d = total > 1.0 ? total : 1.0; /* f64::max(1, total) */
        for (uint64_t i = 0; i < dim; i++) out[i] /= d;
    }
    return nk;
}

/* Batch driver = pybindings/src/oligo.rs:77-81 (rayon into_par_iter, order preserving) and
 * composition/src/oligo.rs:126-143.  OpenMP schedule(dynamic) stands in for rayon's work stealing.
 * out is n x dim row-major f64.  totals (optional) receives the valid-window count per sequence.
 * threads <= 0 -> all cores. */
int ktb_oracle_vectorise_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, int k,
                               int canonical, int norm_mode, double *out, uint64_t *totals,
                               int threads) {
    const uint64_t ncodes = 1ULL << (2 * k);
    uint64_t *pos_map = (uint64_t *)malloc(ncodes * sizeof(uint64_t));
    const uint64_t cnt = ktb_oracle_kmer_pos_maps(k, pos_map, NULL);
    const uint64_t dim = canonical ? cnt : ncodes;
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        uint64_t t = ktb_oracle_vectorise_one(bases + offsets[i], offsets[i + 1] - offsets[i], k,
                                              pos_map, canonical, norm_mode, out + (uint64_t)i * dim, dim);
        if (totals) totals[i] = t;
    }
    free(pos_map);
    return 0;
}

/* Timed CPU baseline ("port" of the reference's per-sequence work, including its allocation
 * pattern): every sequence gets a freshly allocated zeroed Vec<f64> (oligo.rs:237), is counted and
 * normalised, and the row stays alive until the whole batch is done (collect::<Vec<_>>), after which
 * everything is dropped.  checksum defeats dead-code elimination and lets the caller sanity-check. */
double ktb_oracle_baseline_batch(const uint8_t *bases, const uint64_t *offsets, uint64_t n, int k,
                                 int canonical, int norm_mode, int threads, int *threads_used) {
    const uint64_t ncodes = 1ULL << (2 * k);
    uint64_t *pos_map = (uint64_t *)malloc(ncodes * sizeof(uint64_t));
    const uint64_t cnt = ktb_oracle_kmer_pos_maps(k, pos_map, NULL);
    const uint64_t dim = canonical ? cnt : ncodes;
    double **rows = (double **)malloc(n * sizeof(double *));
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    if (threads_used) *threads_used = threads;
    double checksum = 0.0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads) reduction(+ : checksum)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        double *row = (double *)malloc(dim * sizeof(double));
        ktb_oracle_vectorise_one(bases + offsets[i], offsets[i + 1] - offsets[i], k, pos_map,
                                 canonical, norm_mode, row, dim);
        rows[i] = row;
        checksum += row[0] + row[dim - 1];
    }
    for (uint64_t i = 0; i < n; i++) free(rows[i]);
    free(rows);
    free(pos_map);
    return checksum;
}

int ktb_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Header row, composition/src/oligo.rs:69-83 / pybindings/src/oligo.rs:85-99: dim labels of k chars
 * each, written back to back into buf (dim*k bytes, no separators, no NUL). */
int ktb_oracle_header(int k, int canonical, char *buf, uint64_t cap) {
    const uint64_t ncodes = 1ULL << (2 * k);
    char tmp[40];
    if (canonical) {
        uint64_t *p2k = (uint64_t *)malloc(ncodes * sizeof(uint64_t));
        uint64_t cnt = ktb_oracle_kmer_pos_maps(k, NULL, p2k);
        if (cap < cnt * (uint64_t)k) { free(p2k); return -1; }
        for (uint64_t j = 0; j < cnt; j++) {
            ktb_oracle_numeric_to_kmer(p2k[j], k, tmp);
            memcpy(buf + j * k, tmp, k);
        }
        free(p2k);
    } else {
        if (cap < ncodes * (uint64_t)k) return -1;
        for (uint64_t j = 0; j < ncodes; j++) {
            ktb_oracle_numeric_to_kmer(j, k, tmp);
            memcpy(buf + j * k, tmp, k);
        }
    }
    return 0;
}

/* Text row, composition/src/oligo.rs:130-143 and :210-214: norm -> "{:.6}" (Rust rounds the exact
 * binary value half-to-even, as glibc's %.6f does), counts -> "{}" of an integral f64 (no ".0");
 * joined by delim, terminated by '\n'.  Returns bytes written (excluding NUL) or -1 if cap is short. */
int64_t ktb_oracle_format_row(const double *row, uint64_t dim, int norm, const char *delim,
                              char *buf, uint64_t cap) {
    uint64_t w = 0;
    const size_t dl = strlen(delim);
    for (uint64_t i = 0; i < dim; i++) {
        if (cap - w < 64 + dl) return -1;
        if (i) { memcpy(buf + w, delim, dl); w += dl; }
        if (norm) w += (uint64_t)snprintf(buf + w, cap - w, "%.6f", row[i]);
        else w += (uint64_t)snprintf(buf + w, cap - w, "%.0f", row[i]);
    }
    if (cap - w < 2) return -1;
    buf[w++] = '\n';
    buf[w] = 0;
    return (int64_t)w;
}

/* Timed CPU baseline of the FILE-level path (composition/src/oligo.rs:126-144): per sequence vectorise_one, then
 * format every value ("{:.6}" / "{}") and join them into the row's String — in parallel over the sequences of the
 * batch (rayon par_iter().map().collect()) — then the rows are written in order with one buffered writer.
 * path == NULL skips the write.  Returns the number of text bytes produced. */
uint64_t ktb_oracle_baseline_text(const uint8_t *bases, const uint64_t *offsets, uint64_t n, int k, int canonical,
                                  int norm, const char *delim, const char *path, int threads, int *threads_used) {
    const uint64_t ncodes = 1ULL << (2 * k);
    uint64_t *pos_map = (uint64_t *)malloc(ncodes * sizeof(uint64_t));
    const uint64_t cnt = ktb_oracle_kmer_pos_maps(k, pos_map, NULL);
    const uint64_t dim = canonical ? cnt : ncodes;
    char **rows = (char **)malloc(n * sizeof(char *));
    uint64_t *lens = (uint64_t *)malloc(n * sizeof(uint64_t));
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    if (threads_used) *threads_used = threads;
    const uint64_t cap = dim * 24 + 64;
#pragma omp parallel for schedule(dynamic, 64) num_threads(threads)
    for (int64_t i = 0; i < (int64_t)n; i++) {
        double *row = (double *)malloc(dim * sizeof(double));
        ktb_oracle_vectorise_one(bases + offsets[i], offsets[i + 1] - offsets[i], k, pos_map, canonical, norm ? 1 : 0,
                                 row, dim);
        char *txt = (char *)malloc(cap);
        const int64_t w = ktb_oracle_format_row(row, dim, norm, delim, txt, cap);
        rows[i] = txt;
        lens[i] = w > 0 ? (uint64_t)w : 0;
        free(row);
    }
    uint64_t total = 0;
    FILE *fo = path ? fopen(path, "wb") : NULL;
    for (uint64_t i = 0; i < n; i++) {
        if (fo) fwrite(rows[i], 1, lens[i], fo);
        total += lens[i];
        free(rows[i]);
    }
    if (fo) fclose(fo);
    free(rows);
    free(lens);
    free(pos_map);
    return total;
}

/* Checker for the GPU's f32 normalisation (kmertools_b200/csrc/kernels.cuh quot_f32): the same
 * three-operation sequence in C — q0 = c*RN(1/d); rem = fma(-q0,d,c); q = fma(rem,RN(1/d),q0) — must
 * equal (float)((double)c/(double)d), the reference's f64 quotient rounded once, for every
 * 0 <= c <= d, dlo <= d <= dhi.  Returns the number of mismatches. */
#include <math.h>
uint64_t ktb_oracle_check_quot_f32(uint32_t dlo, uint32_t dhi, uint32_t cstep) {
    uint64_t bad = 0;
    if (cstep == 0) cstep = 1;
    for (uint32_t d = dlo; d <= dhi; d++) {
        const float df = (float)d;
        const float rinv = 1.0f / df;
        for (uint32_t c = 0; c <= d; c += cstep) {
            const float cf = (float)c;
            const float q0 = cf * rinv;
            const float rem = fmaf(-q0, df, cf);
            const float q = fmaf(rem, rinv, q0);
            const float want = (float)((double)c / (double)d);
            if (q != want) bad++;
        }
    }
    return bad;
}
