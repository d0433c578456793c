/* refio.c -- FASTA and k-mer model file readers of the host pipeline.
 *
 * FASTA: what gen_ref() gets from kseq (reference src/genref.c:100-127): the name is the header up
 * to the first white space, the sequence is every following line with white space removed; plain or
 * gzip files.
 * Model: the text format read_model() accepts (reference src/model.c:38-131): optional "#k\t<K>"
 * line, optional header line, then 4^K rows "kmer\tlevel_mean\tlevel_stdv..." taken in file order
 * (the k-mer text is not looked up).
 */
#include <ctype.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "refio.h"

#define MAX_KMER_SIZE 9

int sf_fasta_read(const char *path, sf_fasta_t *out, char *err, size_t errcap)
{
    memset(out, 0, sizeof *out);
    gzFile fp = gzopen(path, "r");
    if (!fp) {
        snprintf(err, errcap, "cannot open %s", path);
        return -1;
    }
    gzbuffer(fp, 1 << 20);
    size_t cap_b = 1 << 20, n_b = 0;
    char *bases = (char *)malloc(cap_b);
    int cap_r = 16, n_r = 0;
    char **names = (char **)malloc(sizeof(char *) * cap_r);
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (cap_r + 1));
    const size_t LINE = 1 << 16;
    char *line = (char *)malloc(LINE);
    int in_header = 0; /* a header line longer than the buffer continues */
    while (gzgets(fp, line, (int)LINE)) {
        size_t len = strlen(line);
        const int complete = len && line[len - 1] == '\n';
        if (in_header) { /* rest of an over-long header: ignore */
            in_header = !complete;
            continue;
        }
        if (line[0] == '>') {
            if (n_r == cap_r) {
                cap_r *= 2;
                names = (char **)realloc(names, sizeof(char *) * cap_r);
                off = (int64_t *)realloc(off, sizeof(int64_t) * (cap_r + 1));
            }
            size_t e = 1;
            while (line[e] && !isspace((unsigned char)line[e]))
                e++;
            names[n_r] = (char *)malloc(e);
            memcpy(names[n_r], line + 1, e - 1);
            names[n_r][e - 1] = 0;
            off[n_r] = (int64_t)n_b;
            n_r++;
            in_header = !complete;
            continue;
        }
        if (n_r == 0)
            continue; /* text before the first record */
        if (n_b + len + 1 > cap_b) {
            while (n_b + len + 1 > cap_b)
                cap_b *= 2;
            bases = (char *)realloc(bases, cap_b);
        }
        for (size_t i = 0; i < len; i++)
            if (!isspace((unsigned char)line[i]))
                bases[n_b++] = line[i];
    }
    free(line);
    gzclose(fp);
    if (n_r == 0) {
        free(bases); free(names); free(off);
        snprintf(err, errcap, "%s holds no FASTA record", path);
        return -1;
    }
    off[n_r] = (int64_t)n_b;
    out->num_ref = n_r;
    out->names = names;
    out->bases = bases;
    out->off = off;
    return 0;
}

void sf_fasta_free(sf_fasta_t *f)
{
    for (int i = 0; i < f->num_ref; i++)
        free(f->names[i]);
    free(f->names);
    free(f->bases);
    free(f->off);
    memset(f, 0, sizeof *f);
}

int sf_model_read(const char *path, float **level_mean, uint32_t *kmer_size, char *err, size_t errcap)
{
    FILE *fp = fopen(path, "r");
    if (!fp) {
        snprintf(err, errcap, "cannot open model file %s", path);
        return -1;
    }
    uint32_t k = MAX_KMER_SIZE;
    uint32_t want = 1u << (2 * k);
    float *lm = (float *)malloc(sizeof(float) * ((size_t)1 << (2 * MAX_KMER_SIZE)));
    char *buf = NULL;
    size_t cap = 0;
    uint32_t n = 0, line_no = 0;
    while (getline(&buf, &cap, fp) != -1) {
        line_no++;
        if (buf[0] == '#' || !strncmp(buf, "kmer\tlevel_mean\tlevel_stdv", 26) || buf[0] == '\n' || buf[0] == '\r') {
            char key[1000];
            int val = 0;
            if (sscanf(buf, "%999s\t%d", key, &val) == 2 && !strcmp(key, "#k")) {
                if (val <= 0 || val > MAX_KMER_SIZE) {
                    snprintf(err, errcap, "k-mer size (#k\t%d) in file %s is invalid (1..%d)", val, path, MAX_KMER_SIZE);
                    goto fail;
                }
                k = (uint32_t)val;
                want = 1u << (2 * k);
            }
            continue;
        }
        char kmer[64];
        float mean, stdv;
        if (sscanf(buf, "%63s\t%f\t%f", kmer, &mean, &stdv) != 3) {
            snprintf(err, errcap, "file %s is corrupted at line %u", path, line_no);
            goto fail;
        }
        if (n >= want) {
            snprintf(err, errcap, "file %s has too many entries: expected %u k-mers", path, want);
            goto fail;
        }
        lm[n++] = mean;
    }
    if (n != want) {
        snprintf(err, errcap, "file %s prematurely ended: expected %u k-mers, found %u", path, want, n);
        goto fail;
    }
    free(buf);
    fclose(fp);
    *level_mean = lm;
    *kmer_size = k;
    return 0;
fail:
    free(buf);
    free(lm);
    fclose(fp);
    return -1;
}
