/* sfgpu.h -- C-ABI of libsfgpu.so, the B200 (sm_100a) implementation of sigfish's `dtw` hot path.
 *
 * This is the drop-in boundary.  sigfish has no plugin ABI; the seam is
 * the dormant accelerator hook set of the reference (paths relative to /root/reference):
 *     SIGFISH_ACC / HAVE_ACC           src/sigfish.h:37, Makefile:34-36
 *     init hook in init_core()         src/sigfish.c:200-204
 *     free hook in free_core()         src/sigfish.c:221-225
 *     staged process_db() + align_db() src/sigfish.c:1003-1040
 * Each entry point below names the reference code it replaces.  INTEGRATION.md shows the
 * reference-side patch that binds them.
 *
 * Conventions: plain C, caller-owned host buffers (borrowed until the call returns), all device
 * and pinned memory owned by the library.  Every function returns 0 on success, a negative
 * SFGPU_E* code otherwise; sfgpu_strerror() gives the text.  There is no CPU fallback: without a
 * usable CUDA device sfgpu_create() fails.  Entry points are called from one host thread per
 * context; several contexts (one per GPU) may be driven from different threads.
 *
 * Environment (read by the library, none needed):
 *     SFGPU_TRACE=1                     wall-clock stamps of the set-up and submit / collect steps on stderr
 *     SFGPU_CHECKPOINTS_PER_READ=N      16 .. 4096 (default 512): wavefront checkpoints a read gets on long reference
 *                                       segments, 1.3 KB each; fewer save device memory and lengthen the
 *                                       start-coordinate pass (128: 0.17 MB per read, -1.7 % on a 1 Mb contig).
 *                                       Results do not depend on it.
 */
#ifndef SFGPU_H
#define SFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFGPU_ABI_VERSION 2

/* option bits: identical to the reference's opt.flag (src/sigfish.h:30-39) */
#define SFGPU_RNA 0x001 /* --rna       */
#define SFGPU_DTW 0x002 /* --dtw-std   */
#define SFGPU_INV 0x004 /* --invert    */
#define SFGPU_REF 0x010 /* --full-ref  */
#define SFGPU_END 0x020 /* --from-end  */
#define SFGPU_SAM 0x100 /* --sam       */

#define SFGPU_OK 0
#define SFGPU_ECUDA (-1)   /* CUDA runtime error (text in sfgpu_strerror) */
#define SFGPU_EARG (-2)    /* bad argument */
#define SFGPU_ENODEV (-3)  /* no usable sm_100 device */
#define SFGPU_ESTATE (-4)  /* call out of order (e.g. submit before set_ref) */
#define SFGPU_ELIMIT (-5)  /* unsupported size (e.g. query_size > 1024) */
#define SFGPU_EDECODE (-6) /* a BLOW5 record could not be decoded on the device: decode the batch on the host and resubmit */

typedef struct sfgpu_ctx sfgpu_ctx;

/* mirrors the fields of opt_t (src/sigfish.h:121-139) that the hot path reads */
typedef struct {
    int32_t device;       /* CUDA ordinal */
    uint32_t flags;       /* SFGPU_* bits */
    int32_t query_size;   /* -q  (opt.query_size)  */
    int32_t prefix_size;  /* -p  (opt.prefix_size); < 0: automatic query start (RNA only: the jnn adaptor /
                             poly-A finders of src/jnn.c and detect_query_start(), src/sigfish.c:380-422) */
    int32_t kmer_size;    /* core->kmer_size */
    int32_t n_slots;      /* batches in flight (double buffering); 0 -> 2 */
    int32_t pore;         /* opt.pore_flag: 0 r9, 1 r10, 2 rna004 (only selects the jnn parameters) */
    int32_t reserved[5];  /* test knobs, keep 0: [0] checkpoint spacing, [1] restart window, [2] warm-up blocks of a piece,
                             [3] 1 = no read pairing, 2 = --dtw-std pairs reads for every 128 < q <= 256 (default: 192 < q),
                             [4] piece length in checkpoint periods (< 0: never split) */
} sfgpu_opt_t;

/* per-read output of the device stages: what normalise_single() leaves in db->qstart/qend
 * (src/sigfish.c:479-480), what aln_to_str() reads from the event table (804-805) and the winning
 * entry of dtw_single()'s candidate list before the coordinate flip (969-970). */
typedef struct {
    int64_t n_events;     /* et.n (a lower bound > qend when the block stopped early) */
    int32_t qstart, qend; /* query window in events */
    int32_t qlen;         /* 0: the read prints nothing (empty, ignored or no peak found) */
    int32_t status;       /* bit0 ignored, bit1 too short, bit2 no peak, bit3 sequential prefix-sum redo,
                             bit4 automatic query start failed (50-event fall-back, db->prefix_fail) */
    uint64_t start_raw;   /* event[qstart].start */
    uint64_t end_raw;     /* event[qend-1].start + length */
    float score, score2;  /* aln[4].score, aln[3].score */
    int32_t rid;          /* contig index, -1 if none */
    int32_t strand;       /* 0 '+', 1 '-' */
    int32_t pos_st;       /* aln[4].pos_st  (in-array, before flip / ref_st_offset) */
    int32_t pos_end;      /* aln[4].pos_end */
} sfgpu_result_t;

typedef struct {
    float h2d_ms, events_ms, dtw_ms, trace_ms, d2h_ms, total_ms;
    double cells;         /* sum over reads of qlen * total reference columns */
    int64_t samples;      /* raw samples uploaded */
    int32_t dtw_launches, other_launches;
    int32_t tasks_per_read;  /* DTW tasks one read was cut into (groups + pieces of long segments) */
    int32_t piece_blocks;    /* piece length the batch used, in blocks of 64 columns (0: segments not split) */
    int32_t redone_pieces;   /* pieces recomputed because their warm-up front did not verify */
    int32_t pad;
} sfgpu_timing_t;

/* number of CUDA devices with compute capability 10.x */
int sfgpu_device_count(void);

/* Replaces the accelerator init hook of init_core() (src/sigfish.c:200-204), after the model has
 * been loaded (143-164).  level_mean[4^kmer_size] is model[i].level_mean. */
int sfgpu_create(sfgpu_ctx **ctx, const sfgpu_opt_t *opt, const float *level_mean);

/* Replaces gen_ref() (src/genref.c:86-241, called at src/sigfish.c:178) minus the FASTA parsing:
 * bases = all contigs concatenated, base_off[i]..base_off[i+1] delimits contig i.  Builds the
 * forward (and, for DNA, reverse-complement) event arrays on the device, z-scored, and keeps them
 * resident.  Fills ref_lengths / ref_seq_lengths / ref_st_offset exactly as genref.c:140-142,191. */
int sfgpu_set_ref(sfgpu_ctx *ctx, int32_t num_ref, const char *bases, const int64_t *base_off,
                  int32_t *ref_lengths, int32_t *ref_seq_lengths, int32_t *ref_st_offset);

/* Replaces work_db(event_single) + work_db(normalise_single) + align_db() of the staged
 * process_db() (src/sigfish.c:1028-1040) for one batch, asynchronously.  signals holds the reads'
 * int16 samples back to back, read i at [sig_off[i], sig_off[i+1]); digitisation/offset/range are
 * the slow5 record fields narrowed to float as event_single() does (335-337).  The buffers may be
 * reused as soon as the call returns. */
int sfgpu_submit(sfgpu_ctx *ctx, int32_t slot, int32_t n_reads, const int16_t *signals,
                 const int64_t *sig_off, const float *digitisation, const float *offset,
                 const float *range);

/* sfgpu_submit with one pointer per read (signals[i] holds n_samples[i] samples), so that a host that keeps
 * its reads in separate buffers, like db->slow5_rec[i]->raw_signal (src/sigfish.h:168), needs no gather
 * copy: the samples go straight into the slot's pinned staging buffer. */
int sfgpu_submit_reads(sfgpu_ctx *ctx, int32_t slot, int32_t n_reads, const int16_t *const *signals,
                       const int64_t *n_samples, const float *digitisation, const float *offset,
                       const float *range);

/* sfgpu_submit_reads for reads that are still BLOW5 records as they lie in the file: replaces, on top of what
 * sfgpu_submit replaces, work_db(parse_single) = slow5_rec_depress_parse() (src/sigfish.c:317-328,1024; slow5lib
 * src/slow5.c:2575-2609,3191-3283; src/slow5_press.c:1085-1133).  records[i] points at record_bytes[i] bytes (what
 * follows the u64 size in the file); record_press / signal_press are the file header's methods (0 none, 1 zlib;
 * 0 none, 1 svb-zd; anything else: SFGPU_ELIMIT).  The caller has read the head of each record: sig_pos[i] is the
 * offset of the raw_signal field in the decompressed record (2 + id length + 4 + 32 + 8), sig_bytes[i] its size in
 * bytes (the len_raw_signal field of an svb-zd file, 2 * samples otherwise), n_samples[i] the sample count.
 * Inflate and signal decoding run on the device; sfgpu_collect() returns SFGPU_EDECODE when a record turns out
 * malformed (the host then decodes the batch itself and resubmits it with sfgpu_submit_reads). */
int sfgpu_submit_records(sfgpu_ctx *ctx, int32_t slot, int32_t n_reads, const uint8_t *const *records,
                         const int64_t *record_bytes, int32_t record_press, int32_t signal_press, const int32_t *sig_pos,
                         const int64_t *sig_bytes, const int64_t *n_samples, const float *digitisation, const float *offset,
                         const float *range);

/* Same as sfgpu_submit but re-runs the device stages on the inputs already resident in the slot
 * (no host->device copy).  Used to measure device-only throughput. */
int sfgpu_resubmit(sfgpu_ctx *ctx, int32_t slot);

/* Waits for the slot's batch and returns its results in read order (src/sigfish.c:1056-1071 keeps
 * input order).  out has room for the n_reads given to sfgpu_submit. */
int sfgpu_collect(sfgpu_ctx *ctx, int32_t slot, sfgpu_result_t *out);

/* --sam (contexts created with SFGPU_SAM), after sfgpu_collect(): the full warping path of every read's
 * winning hit -- what update_aln() keeps as r2qevent_map via subsequence_path() + path_to_map()
 * (src/sigfish.c:530-571, 599-618; src/cdtw.c:98-167, 192-227) -- and the window's event boundaries
 * that r2qevent_map_to_ss() reads (src/sigfish.c:737-742).
 *   moves[move_off[i] .. move_off[i] + n_moves[i])  the path of read i BACKWARDS from the cell
 *       (qlen-1, pos_end): 0 = diagonal (i-1, j-1), 1 = left (i, j-1), 2 = up (i-1, j); it ends on row 0 at
 *       column pos_st.  The caller sizes move_off so that read i has room for qlen + pos_end - pos_st
 *       moves; n_moves[i] = -1 for reads without a hit.
 *   ev_start / ev_len [i * query_size + k]  event qstart + k of read i (start sample, length). */
int sfgpu_collect_paths(sfgpu_ctx *ctx, int32_t slot, const int64_t *move_off, uint8_t *moves,
                        int32_t *n_moves, uint64_t *ev_start, float *ev_len);

/* CUDA-event timings of the slot's last completed batch */
int sfgpu_timing(sfgpu_ctx *ctx, int32_t slot, sfgpu_timing_t *t);

/* Replaces the accelerator free hook of free_core() (src/sigfish.c:221-225) */
void sfgpu_destroy(sfgpu_ctx *ctx);

const char *sfgpu_strerror(const sfgpu_ctx *ctx);

/* ---- inspection entry points used by the parity tests ---- */

/* copies the z-scored event array of (contig rid, strand) to out; returns its length or <0 */
int sfgpu_ref_events(sfgpu_ctx *ctx, int32_t rid, int32_t strand, float *out, int32_t cap);

/* full event table of one read (no early exit): starts/lengths/means as create_events()
 * (src/events.c:479-508) produces them; returns the number of events or <0 */
int64_t sfgpu_event_table(sfgpu_ctx *ctx, const int16_t *signal, int64_t n_samples, float digitisation,
                          float offset, float range, uint64_t *start, float *length, float *mean,
                          int64_t cap);

/* the int16 samples of read i of the slot's last batch as the device holds them (after sfgpu_submit_records: as
 * the device decoded them); returns the sample count or < 0 */
int64_t sfgpu_slot_signal(sfgpu_ctx *ctx, int32_t slot, int32_t read, int16_t *out, int64_t cap);

/* the normalised query of read i of the slot's last batch (src/sigfish.c:857-867); returns qlen */
int sfgpu_query(sfgpu_ctx *ctx, int32_t slot, int32_t read, float *out, int32_t cap);

/* Loads caller-made event arrays instead of synthesising them from bases: contig i occupies
 * events[ev_off[2i] .. ev_off[2i+1]) on '+' and, when has_reverse != 0, events[ev_off[2i+1] ..
 * ev_off[2i+2]) on '-' (same length); with has_reverse == 0 ev_off has num_ref+1 entries.  The
 * arrays are used as they are (no z-score).  Lets the tests drive the DTW kernels with arbitrary
 * (tie-heavy) inputs, as the reference's subsequence()/std_dtw() can be (src/cdtw.c:69-94,171-189). */
int sfgpu_set_ref_events(sfgpu_ctx *ctx, int32_t num_ref, int32_t has_reverse, const float *events,
                         const int64_t *ev_off);

/* Runs only the alignment stages (dtw_single, src/sigfish.c:828-985) on caller-made queries:
 * queries[i*query_size .. +qlen[i]) is read i's query exactly as dtw_single() would build it
 * (z-scored, already reversed for RNA).  Results come back through sfgpu_collect(); the event
 * fields of sfgpu_result_t are zero except qlen/qend. */
int sfgpu_submit_queries(sfgpu_ctx *ctx, int32_t slot, int32_t n_reads, const float *queries,
                         const int32_t *qlen);

/* Number of reads whose (read, segment group) DTW tasks fill the resident warps of the DTW kernel once (at least 1).
 * A sizing hint only: long segments are cut into column pieces per batch, so batches need not be multiples of it
 * (a 512-read batch against a 1 Mb contig runs at 94 % of the rate of a 16 576-read one). */
int32_t sfgpu_wave_reads(const sfgpu_ctx *ctx);

/* reference columns one read is aligned against (all contigs, both strands for DNA) */
int64_t sfgpu_ref_columns(const sfgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
