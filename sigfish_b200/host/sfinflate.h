/* sfinflate.h -- one-shot zlib-stream decoder used for BLOW5 records (see sfinflate.c) */
#ifndef SF_INFLATE_H
#define SF_INFLATE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SF_INFLATE_LITLEN_BITS 10   /* first-level table index bits: small enough to rebuild per 6 KB record */
#define SF_INFLATE_DIST_BITS 8
#define SF_INFLATE_LITLEN_SUB 1536 /* room for second-level tables (checked when they are built) */
#define SF_INFLATE_DIST_SUB 512

/* decode tables; one per thread, zero-initialised (calloc) before the first use */
typedef struct {
    uint32_t litlen[(1 << SF_INFLATE_LITLEN_BITS) + SF_INFLATE_LITLEN_SUB];
    uint32_t dist[(1 << SF_INFLATE_DIST_BITS) + SF_INFLATE_DIST_SUB];
    uint32_t fixed_litlen[1 << SF_INFLATE_LITLEN_BITS];
    uint32_t fixed_dist[1 << SF_INFLATE_DIST_BITS];
    uint32_t precode[1 << 7];
    uint32_t sym_litlen[288], sym_dist[32], sym_precode[19];
    int ready, fixed_ready;
} sf_inflater;

/* Inflates the zlib stream in[0 .. n_in) into out[0 .. cap_out); *n_out = bytes produced.
 * 0 = ok (Adler-32 verified), 1 = cap_out too small (out[0 .. *n_out) is the valid beginning of the output),
 * -1 = corrupt or unsupported stream. */
int sf_zlib_inflate(sf_inflater *d, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap_out, size_t *n_out);

/* The beginning of a stream whose plain text starts with a u16 length: decodes min(cap_out, 2 + that length + tail)
 * bytes without building lookup tables (the head of a BLOW5 record: id length, id, fixed fields).  1 = done, *n_out
 * bytes are valid; 0 = not decoded here (uncommon layout or damage): use sf_zlib_inflate(). */
int sf_zlib_inflate_prefix(sf_inflater *d, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap_out, size_t tail, size_t *n_out);

#ifdef __cplusplus
}
#endif
#endif
