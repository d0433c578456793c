/* refio.h -- FASTA and k-mer model readers (see refio.c) */
#ifndef SF_REFIO_H
#define SF_REFIO_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t num_ref;
    char **names;
    char *bases;   /* all contigs back to back */
    int64_t *off;  /* [num_ref + 1] */
} sf_fasta_t;

int sf_fasta_read(const char *path, sf_fasta_t *out, char *err, size_t errcap);
void sf_fasta_free(sf_fasta_t *f);
/* level_mean[4^k] (malloc'd) in file order */
int sf_model_read(const char *path, float **level_mean, uint32_t *kmer_size, char *err, size_t errcap);

#ifdef __cplusplus
}
#endif
#endif
