/* Minimal C client of libkmertools_b200.so: canonical, L1-normalised k-mer frequency rows of a few sequences.
 *
 *   gcc -O2 -o vectorise examples/vectorise.c -Iinclude -Lkmertools_b200/lib -lkmertools_b200 \
 *       -Wl,-rpath,$PWD/kmertools_b200/lib
 *   ./vectorise 4 ACGTACGTNACGT GATTACA
 *
 * Prints one line per sequence, the same values `kmertools comp oligo -k 4` writes (composition/src/oligo.rs:130-143).
 * Without a GPU it reports KTB_ERR_NODEVICE and exits 3: the library has no CPU path. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmertools_b200.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s K SEQ [SEQ ...]\n", argv[0]);
        return 2;
    }
    const int k = atoi(argv[1]);
    const uint64_t n = (uint64_t)(argc - 2);

    /* sequences back to back + n+1 offsets: the layout every entry point takes */
    uint64_t *offsets = (uint64_t *)calloc(n + 1, sizeof *offsets);
    size_t total = 0;
    for (uint64_t i = 0; i < n; ++i) total += strlen(argv[2 + i]);
    uint8_t *bases = (uint8_t *)malloc(total ? total : 1);
    for (uint64_t i = 0; i < n; ++i) {
        const size_t len = strlen(argv[2 + i]);
        memcpy(bases + offsets[i], argv[2 + i], len);
        offsets[i + 1] = offsets[i] + len;
    }

    ktb_oligo *h = NULL;
    int rc = ktb_oligo_create(k, 0, &h);
    if (rc != KTB_OK) {
        fprintf(stderr, "ktb_oligo_create: %s\n", ktb_last_error());
        return rc == KTB_ERR_NODEVICE ? 3 : 1;
    }
    const uint64_t dim = ktb_oligo_dim(h, 1);
    double *rows = (double *)malloc(n * dim * sizeof *rows);
    uint64_t *totals = (uint64_t *)malloc(n * sizeof *totals);
    rc = ktb_oligo_vectorise(h, bases, offsets, n, /*canonical*/ 1, KTB_NORM_CLI, KTB_OUT_F64, rows, totals);
    if (rc != KTB_OK) {
        fprintf(stderr, "ktb_oligo_vectorise: %s\n", ktb_last_error());
        return 1;
    }
    for (uint64_t i = 0; i < n; ++i) {
        for (uint64_t j = 0; j < dim; ++j) printf("%s%.6f", j ? " " : "", rows[i * dim + j]);
        printf("\n");
        fprintf(stderr, "sequence %llu: %llu valid %d-mers\n", (unsigned long long)i, (unsigned long long)totals[i], k);
    }
    ktb_oligo_destroy(h);
    free(rows); free(totals); free(bases); free(offsets);
    return 0;
}
