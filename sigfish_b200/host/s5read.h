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
/* decodes a raw record (slow5_rec_depress_parse).  Thread safe for distinct rec/scratch; modifies
 * mem for ASCII records. */
int sf_s5_parse(const sf_s5file_t *f, char *mem, size_t bytes, sf_rec_t *rec, char **scratch, size_t *scratch_cap);

#ifdef __cplusplus
}
#endif
#endif
