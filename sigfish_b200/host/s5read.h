/* s5read.h -- sequential SLOW5 / BLOW5 reader (see s5read.c) */
#ifndef SF_S5READ_H
#define SF_S5READ_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#include "sfhost.h"

#ifdef __cplusplus
extern "C" {
#endif

/* opens a .slow5 / .blow5 file and reads its header; NULL on error (text in err) */
sf_s5file_t *sf_s5_open(const char *path, char *err, size_t errcap);
void sf_s5_close(sf_s5file_t *f);
/* value of header attribute `attr` for a read group, NULL when absent (slow5_hdr_get) */
const char *sf_s5_hdr_get(const sf_s5file_t *f, const char *attr, uint32_t read_group);
uint32_t sf_s5_num_read_groups(const sf_s5file_t *f);
int sf_s5_is_binary(const sf_s5file_t *f);
const char *sf_s5_error(const sf_s5file_t *f);
/* next raw (still compressed) record into *mem; returns its size in bytes, 0 at end of file, < 0 on
 * error (slow5_get_next_mem).  Not thread safe. */
int64_t sf_s5_get_next_mem(sf_s5file_t *f, char **mem, size_t *cap);
/* binary files are memory mapped when possible: then records can be taken as views into the mapping instead of
 * copies (same cursor as sf_s5_get_next_mem) */
int sf_s5_is_mapped(const sf_s5file_t *f);
int64_t sf_s5_get_next_view(sf_s5file_t *f, const char **ptr);
void sf_s5_prefault(const sf_s5file_t *f, const char *ptr, size_t len);
/* decodes a raw record (slow5_rec_depress_parse).  Thread safe for distinct rec/scratch; modifies
 * mem for ASCII records. */
int sf_s5_parse(const sf_s5file_t *f, char *mem, size_t bytes, sf_rec_t *rec, char **scratch, size_t *scratch_cap);

/* head of a binary record only (read id, scaling, sample count) plus the position and size of its signal field in
 * the decompressed record: what sfgpu_submit_records() needs.  < 0: not applicable or malformed (use sf_s5_parse). */
int sf_s5_parse_head(const sf_s5file_t *f, const char *mem, size_t bytes, sf_rec_t *rec, int32_t *sig_pos, int64_t *sig_bytes,
                     char **scratch, size_t *scratch_cap);
uint8_t sf_s5_record_press(const sf_s5file_t *f); /* 0 none, 1 zlib */
uint8_t sf_s5_signal_press(const sf_s5file_t *f); /* 0 none, 1 svb-zd */

#ifdef __cplusplus
}
#endif
#endif
