/* oracle.h -- CPU restatement of the reference `sigfish dtw` mapping path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by
 * or executed from the product (sigfish_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call it, and only as the checker.
 *
 * Parity status: PINNED.  Every function here is checked in tests/ against the
 * unmodified reference compiled into oracle/_ref/ (binary PAF output and direct
 * calls into libsigfish_ref.so), and against the golden vectors under
 * tests/golden/ that were produced by that reference binary.
 *
 * Plain C99, no contraction (-ffp-contract=off), same evaluation order as the
 * reference (x86-64, FLT_EVAL_METHOD == 0).  Each function cites the reference
 * file:line it restates (paths relative to /root/reference).
 */
#ifndef SIGFISH_ORACLE_H
#define SIGFISH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* option bits -- same values as reference src/sigfish.h:30-39 */
#define ORC_RNA 0x001
#define ORC_DTW 0x002 /* --dtw-std  */
#define ORC_INV 0x004 /* --invert   */
#define ORC_REF 0x010 /* --full-ref */
#define ORC_END 0x020 /* --from-end */
#define ORC_RNA004 0x400 /* oracle-only bit: pore_flag == OPT_PORE_RNA004 (selects the jnn parameter set) */

/* reference src/sigfish.h:57-63 */
typedef struct {
    uint64_t start;
    float length;
    float mean;
    float stdv;
} orc_event_t;

/* ---- events (reference src/events.c:297-577, src/sigfish.c:330-347) ---- */

/* int16 ADC -> pA, fp32, no contraction (sigfish.c:344-347) */
void orc_to_picoamps(const int16_t *raw, int64_t n, float digitisation, float offset,
                     float range, float *pa);
/* events.c:297-307 */
void orc_prefix_sums(const float *x, int64_t n, double *sum, double *sumsq);
/* events.c:319-368 */
void orc_tstat(const double *sum, const double *sumsq, int64_t n, int64_t w, float *t);
/* events.c:375-447 ; peaks has room for n entries; returns the number emitted */
int64_t orc_peaks(const float *t1, const float *t2, int64_t n, int rna, uint64_t *peaks);
/* whole chain: getevents() events.c:557-577 minus the dead MAD trim.
 * *out is malloc'd (caller frees with orc_free); returns the number of events,
 * or -1 when the reference would hit undefined behaviour (no peak found). */
int64_t orc_detect_events(const int16_t *raw, int64_t n, float digitisation, float offset,
                          float range, int rna, orc_event_t **out);

/* ---- automatic query start, -p < 0 (reference src/jnn.c, src/sigfish.c:380-422) ---- */
/* raw-sample index where the poly-A tail ends, or -1 when adaptor / poly-A are not found */
int64_t orc_polya_end_sample(const int16_t *raw, int64_t n, float digitisation, float offset, float range,
                             int rna004);

/* ---- reference synthesis (reference src/genref.c:23-241, src/ref.h:13-76) ---- */
typedef struct {
    int32_t num_ref;
    int32_t has_reverse;      /* DNA: 1, RNA: 0 */
    int32_t *ref_lengths;     /* k-mer count actually aligned against */
    int32_t *ref_seq_lengths; /* contig length in bases */
    int32_t *ref_st_offset;
    float **forward;
    float **reverse;
} orc_ref_t;

orc_ref_t *orc_ref_build(int32_t num_ref, const char *const *seqs, const int32_t *seq_lens,
                         const float *level_mean, int32_t kmer_size, uint32_t flags,
                         int32_t query_size);
/* test helper: reference made of caller-given event arrays (flat, contig after contig) */
orc_ref_t *orc_ref_from_events(int32_t num_ref, int32_t has_reverse, const int32_t *lens,
                               const float *fwd_flat, const float *rev_flat);
void orc_ref_free(orc_ref_t *r);
/* accessors for ctypes */
int32_t orc_ref_len(const orc_ref_t *r, int32_t i);
int32_t orc_ref_offset(const orc_ref_t *r, int32_t i);
const float *orc_ref_fwd(const orc_ref_t *r, int32_t i);
const float *orc_ref_rev(const orc_ref_t *r, int32_t i);

/* ---- DTW (reference src/cdtw.c) ---- */
void orc_subsequence(const float *x, const float *y, int n, int m, float *cost); /* 171-189 */
float orc_std_dtw(const float *x, const float *y, int n, int m, float *cost);     /* 69-94  */
/* cdtw.c:98-167 + 192-227 reduced to what the caller reads: p.py[0] */
int32_t orc_path_start(const float *cost, int n, int m, int end_col);
/* same walk, also returning the full path (row-major order start->end); returns k.
 * px/py must have room for n+m entries. */
int32_t orc_path_full(const float *cost, int n, int m, int end_col, int32_t *px, int32_t *py);

/* ---- per-read mapping (reference src/sigfish.c:424-505, 507-626, 828-985) ---- */
typedef struct {
    int32_t mapped;      /* 0: the read prints nothing (len 0 / ignored) */
    int32_t status;      /* bit0: ignored, bit1: too short */
    int64_t n_events;    /* et.n as produced by event detection */
    int64_t qstart, qend;
    uint64_t start_raw, end_raw; /* sigfish.c:804-805 */
    int32_t rid;
    int32_t pos_st, pos_end;     /* final PAF coordinates (after flip + offset) */
    int32_t raw_pos_st, raw_pos_end; /* in-array coordinates of the winning hit */
    float score, score2;
    int32_t mapq;
    char strand;
} orc_hit_t;

/* query window + z-score (sigfish.c:424-505); events are modified in place.
 * Returns 1 when the read goes on to DTW. */
int orc_window_normalise(orc_event_t *ev, int64_t *n_events, uint32_t flags, int32_t q, int32_t p,
                         int64_t *qstart, int64_t *qend, int32_t *status);

/* dtw_single (sigfish.c:828-985) on an already normalised event table */
void orc_align(const orc_ref_t *ref, const orc_event_t *ev, int64_t qstart, int64_t qend,
               uint32_t flags, orc_hit_t *hit);

/* test helper: orc_align on a bare array of normalised event means */
void orc_align_means(const orc_ref_t *ref, const float *means, int32_t n, uint32_t flags, orc_hit_t *hit);

/* event_single + normalise_single + dtw_single for one read */
void orc_map_read(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation,
                  float offset, float range, uint32_t flags, int32_t q, int32_t p, orc_hit_t *hit);

/* the same, also returning the (window-normalised) event table; *ev_out is freed with orc_free */
void orc_map_read_events(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation, float offset,
                         float range, uint32_t flags, int32_t q, int32_t p, orc_hit_t *hit, orc_event_t **ev_out);

/* paf_str (sigfish.c:628-660) -- returns the number of bytes written (excl. NUL) */
int orc_paf_line(char *buf, size_t cap, const orc_hit_t *hit, const char *read_id,
                 const char *rname, int32_t ref_seq_len, int64_t len_raw_signal);

/* sam_str (sigfish.c:770-794) for the hit orc_align() produced from the same normalised events */
int orc_sam_line(const orc_ref_t *ref, const orc_event_t *ev, uint32_t flags, const orc_hit_t *hit,
                 const char *read_id, const char *rname, char *buf, size_t cap);
/* whole read -> SAM line (0 bytes when the read prints nothing) */
int orc_map_read_sam(const orc_ref_t *ref, const int16_t *raw, int64_t n, float digitisation, float offset,
                     float range, uint32_t flags, int32_t q, int32_t p, const char *read_id,
                     const char *const *rnames, char *buf, size_t cap);

void orc_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
