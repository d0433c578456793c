/* oracle_ref.c -- k-mer-model reference synthesis restated on the CPU (test oracle).
 *
 * Follows reference src/genref.c:23-47 (per-array z-score), 86-241 (gen_ref) and
 * src/ref.h:13-76 (k-mer rank, reverse complement).  FASTA parsing is not part of
 * the oracle: callers hand over the contig strings.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* ref.h:13-26 -- A,C,G,T (either case) -> 0..3; anything else counts as 0 */
static uint32_t base_code(char b)
{
    switch (b) {
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0;
    }
}

/* ref.h:30-41 -- first base is the most significant digit */
static uint32_t kmer_rank(const char *s, int32_t k)
{
    uint32_t r = 0;
    for (int32_t i = 0; i < k; i++)
        r = (r << 2) | base_code(s[i]);
    return r;
}

/* ref.h:45-65 -- unknown bases complement to 'T' */
static char comp(char b)
{
    switch (b) {
    case 'A': case 'a': return 'T';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    default: return b == 'T' || b == 't' ? 'A' : 'T';
    }
}

/* genref.c:23-47 -- fp32 running sums in array order, population stdv */
static void zscore(float *a, int64_t n)
{
    const float cnt = (float)(uint64_t)n;
    float mean = 0.0f;
    for (int64_t j = 0; j < n; j++)
        mean += a[j];
    mean /= cnt;
    float var = 0.0f;
    for (int64_t j = 0; j < n; j++) {
        float d = a[j] - mean;
        var += d * d;
    }
    var /= cnt;
    const float sd = (float)sqrt((double)var);
    for (int64_t j = 0; j < n; j++)
        a[j] = (a[j] - mean) / sd;
}

orc_ref_t *orc_ref_build(int32_t num_ref, const char *const *seqs, const int32_t *seq_lens,
                         const float *level_mean, int32_t kmer_size, uint32_t flags,
                         int32_t query_size)
{
    const int rna = (flags & ORC_RNA) != 0;
    orc_ref_t *r = (orc_ref_t *)calloc(1, sizeof(orc_ref_t));
    r->num_ref = num_ref;
    r->has_reverse = !rna;
    r->ref_lengths = (int32_t *)calloc((size_t)num_ref, sizeof(int32_t));
    r->ref_seq_lengths = (int32_t *)calloc((size_t)num_ref, sizeof(int32_t));
    r->ref_st_offset = (int32_t *)calloc((size_t)num_ref, sizeof(int32_t));
    r->forward = (float **)calloc((size_t)num_ref, sizeof(float *));
    r->reverse = (float **)calloc((size_t)num_ref, sizeof(float *));

    for (int32_t c = 0; c < num_ref; c++) {
        const char *s = seqs[c];
        const int32_t l = seq_lens[c];
        const int32_t nk = l + 1 - kmer_size; /* k-mers in the contig */

        /* genref.c:128-136: DNA and --full-ref use every k-mer, RNA only 1.5*q of them */
        int32_t rl = nk;
        if (rna && !(flags & ORC_REF)) {
            uint32_t cap = (uint32_t)(query_size * 1.5);
            rl = cap > (uint32_t)nk ? nk : (int32_t)cap;
        }
        r->ref_lengths[c] = rl;
        r->ref_seq_lengths[c] = l;
        r->ref_st_offset[c] = 0;
        float *fw = (float *)malloc(sizeof(float) * (size_t)rl);
        r->forward[c] = fw;

        if (!rna) { /* genref.c:157-164 */
            float *rv = (float *)malloc(sizeof(float) * (size_t)rl);
            r->reverse[c] = rv;
            char *rc = (char *)malloc((size_t)l + 1);
            for (int32_t i = 0; i < l; i++)
                rc[i] = comp(s[l - 1 - i]);
            rc[l] = 0;
            for (int32_t j = 0; j < rl; j++) {
                fw[j] = level_mean[kmer_rank(s + j, kmer_size)];
                rv[j] = level_mean[kmer_rank(rc + j, kmer_size)];
            }
            free(rc);
        } else if (flags & ORC_INV) { /* genref.c:166-177: last rl k-mers, written back to front */
            const char *tail = s + l - rl - (kmer_size - 1);
            for (int32_t j = 0; j < rl; j++)
                fw[rl - 1 - j] = level_mean[kmer_rank(tail + j, kmer_size)];
        } else { /* genref.c:184-197 */
            const char *from = s;
            if (!(flags & ORC_END)) {
                r->ref_st_offset[c] = l - rl - (kmer_size - 1);
                from = s + r->ref_st_offset[c];
            }
            for (int32_t j = 0; j < rl; j++)
                fw[j] = level_mean[kmer_rank(from + j, kmer_size)];
        }

        zscore(fw, rl); /* genref.c:210-217 */
        if (!rna)
            zscore(r->reverse[c], rl);
    }
    return r;
}

void orc_ref_free(orc_ref_t *r)
{
    if (!r)
        return;
    for (int32_t c = 0; c < r->num_ref; c++) {
        free(r->forward[c]);
        free(r->reverse[c]);
    }
    free(r->forward);
    free(r->reverse);
    free(r->ref_lengths);
    free(r->ref_seq_lengths);
    free(r->ref_st_offset);
    free(r);
}

int32_t orc_ref_len(const orc_ref_t *r, int32_t i) { return r->ref_lengths[i]; }
int32_t orc_ref_offset(const orc_ref_t *r, int32_t i) { return r->ref_st_offset[i]; }
const float *orc_ref_fwd(const orc_ref_t *r, int32_t i) { return r->forward[i]; }
const float *orc_ref_rev(const orc_ref_t *r, int32_t i) { return r->reverse[i]; }

/* test helper: a reference made of caller-given event arrays (flat, back to back), used to
 * drive orc_align() -- the restatement of dtw_single() -- with arbitrary, e.g. tie-heavy, inputs */
orc_ref_t *orc_ref_from_events(int32_t num_ref, int32_t has_reverse, const int32_t *lens,
                               const float *fwd_flat, const float *rev_flat)
{
    orc_ref_t *r = (orc_ref_t *)calloc(1, sizeof(orc_ref_t));
    r->num_ref = num_ref;
    r->has_reverse = has_reverse;
    r->ref_lengths = (int32_t *)calloc((size_t)num_ref, sizeof(int32_t));
    r->ref_seq_lengths = (int32_t *)calloc((size_t)num_ref, sizeof(int32_t));
    r->ref_st_offset = (int32_t *)calloc((size_t)num_ref, sizeof(int32_t));
    r->forward = (float **)calloc((size_t)num_ref, sizeof(float *));
    r->reverse = (float **)calloc((size_t)num_ref, sizeof(float *));
    size_t at = 0;
    for (int32_t c = 0; c < num_ref; c++) {
        const size_t n = (size_t)lens[c];
        r->ref_lengths[c] = lens[c];
        r->ref_seq_lengths[c] = lens[c];
        r->forward[c] = (float *)malloc(sizeof(float) * n);
        memcpy(r->forward[c], fwd_flat + at, sizeof(float) * n);
        if (has_reverse) {
            r->reverse[c] = (float *)malloc(sizeof(float) * n);
            memcpy(r->reverse[c], rev_flat + at, sizeof(float) * n);
        }
        at += n;
    }
    return r;
}
