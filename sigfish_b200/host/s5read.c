/* s5read.c -- sequential SLOW5 (ASCII) / BLOW5 (binary) reader for the `dtw` host pipeline.
 *
 * Implements the part of the SLOW5 format specification (v0.1.0 / v0.2.0) that load_db() and
 * parse_single() of the reference use through slow5lib (src/sigfish.c:274-328:
 * slow5_get_next_mem + slow5_rec_depress_parse) and the two header attributes init_core() looks at
 * (experiment_type, sequencing_kit: src/sigfish.c:22-80).  Written from the format description:
 *   BLOW5: "BLOW5\1", version (3 x u8), record compression (u8: 0 none, 1 zlib, 2 zstd),
 *          number of read groups (u32), signal compression (u8: 0 none, 1 svb-zd; v0.2.0+),
 *          padding up to byte 64, u32 header length, ASCII header, then records
 *          { u64 n_bytes; n_bytes of (compressed) record }, then "5WOLB".
 *   record (after record decompression): u16 id_len, id, u32 read_group, f64 digitisation, f64
 *          offset, f64 range, f64 sampling_rate, u64 len_raw_signal, signal, auxiliary fields
 *          (ignored here).  With svb-zd the u64 is the number of compressed bytes and the signal is
 *          { u32 n_samples; StreamVByte(zigzag(delta)) }.
 *   SLOW5: the same header as text, one TSV row per record, signal as comma separated integers.
 * zstd record compression is not supported (libzstd is not in this image either).
 */
#include <errno.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include "s5read.h"
#include "sfinflate.h"

struct sf_s5file {
    FILE *fp;
    int binary;
    uint8_t ver[3];
    uint8_t record_press; /* 0 none, 1 zlib */
    uint8_t signal_press; /* 0 none, 1 svb-zd */
    uint32_t num_read_groups;
    /* header attributes: "@name\tv0\tv1..." */
    int n_attr;
    char **attr_name;
    char ***attr_val;
    uint32_t *attr_width; /* values allocated for attribute i (the read-group count when its line was read) */
    char *line; /* ASCII line buffer */
    size_t line_cap;
    char errmsg[256];
    /* binary files are mapped: records are handed out as views into the mapping (sf_s5_get_next_view) */
    const uint8_t *map;
    size_t map_len, map_pos;
};

static void set_err(sf_s5file_t *f, const char *msg) { snprintf(f->errmsg, sizeof f->errmsg, "%s", msg); }

static int hdr_add_line(sf_s5file_t *f, char *line)
{
    /* line starts with '@'; split on tabs */
    size_t len = strlen(line);
    while (len && (line[len - 1] == '\n' || line[len - 1] == '\r'))
        line[--len] = 0;
    char *save = NULL;
    char *name = strtok_r(line + 1, "\t", &save);
    if (!name)
        return 0;
    f->attr_name = (char **)realloc(f->attr_name, sizeof(char *) * (f->n_attr + 1));
    f->attr_val = (char ***)realloc(f->attr_val, sizeof(char **) * (f->n_attr + 1));
    f->attr_width = (uint32_t *)realloc(f->attr_width, sizeof(uint32_t) * (f->n_attr + 1));
    f->attr_name[f->n_attr] = strdup(name);
    /* one value per read group; the row keeps its own width, so that a later (malformed) change of the
     * read-group count cannot make readers index past it */
    const uint32_t width = f->num_read_groups ? f->num_read_groups : 1;
    char **vals = (char **)calloc(width, sizeof(char *));
    if (!vals)
        return -1;
    for (uint32_t g = 0; g < width; g++) {
        char *v = strtok_r(NULL, "\t", &save);
        if (!v)
            break;
        vals[g] = strdup(v);
    }
    f->attr_val[f->n_attr] = vals;
    f->attr_width[f->n_attr] = width;
    f->n_attr++;
    return 0;
}

/* "#num_read_groups" of the text header.  BLOW5 carries the count in its binary header as well: that one wins.
 * The count is bounded and frozen once the first '@' row has been sized with it. */
#define SF_S5_MAX_READ_GROUPS 65536u
static void set_read_groups(sf_s5file_t *f, const char *text)
{
    const unsigned long v = strtoul(text, NULL, 10);
    if (f->binary || f->n_attr > 0 || v == 0 || v > SF_S5_MAX_READ_GROUPS)
        return;
    f->num_read_groups = (uint32_t)v;
}

/* parses the text header held in buf (binary files) */
static int parse_text_header(sf_s5file_t *f, char *buf)
{
    char *save = NULL;
    for (char *line = strtok_r(buf, "\n", &save); line; line = strtok_r(NULL, "\n", &save)) {
        if (line[0] == '@') {
            char *copy = strdup(line);
            hdr_add_line(f, copy);
            free(copy);
        } else if (!strncmp(line, "#num_read_groups\t", 17)) {
            set_read_groups(f, line + 17);
        }
    }
    return 0;
}

sf_s5file_t *sf_s5_open(const char *path, char *err, size_t errcap)
{
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        snprintf(err, errcap, "cannot open %s: %s", path, strerror(errno));
        return NULL;
    }
    sf_s5file_t *f = (sf_s5file_t *)calloc(1, sizeof *f);
    f->fp = fp;
    f->num_read_groups = 1;
    unsigned char head[64];
    size_t n = fread(head, 1, 6, fp);
    if (n == 6 && !memcmp(head, "BLOW5\1", 6)) {
        f->binary = 1;
        if (fread(head + 6, 1, 58, fp) != 58) {
            snprintf(err, errcap, "%s: truncated BLOW5 header", path);
            goto fail;
        }
        memcpy(f->ver, head + 6, 3);
        f->record_press = head[9];
        memcpy(&f->num_read_groups, head + 10, 4);
        if (f->num_read_groups == 0 || f->num_read_groups > SF_S5_MAX_READ_GROUPS) {
            snprintf(err, errcap, "%s: implausible number of read groups (%u)", path, f->num_read_groups);
            f->num_read_groups = 1;
            goto fail;
        }
        f->signal_press = (f->ver[0] > 0 || f->ver[1] >= 2) ? head[14] : 0;
        if (f->record_press > 1) {
            snprintf(err, errcap, "%s: record compression method %d is not supported (only none/zlib)", path, f->record_press);
            goto fail;
        }
        if (f->signal_press > 1) {
            snprintf(err, errcap, "%s: signal compression method %d is not supported (only none/svb-zd)", path, f->signal_press);
            goto fail;
        }
        uint32_t hlen = 0;
        if (fread(&hlen, 4, 1, fp) != 1) {
            snprintf(err, errcap, "%s: truncated BLOW5 header", path);
            goto fail;
        }
        char *buf = (char *)malloc((size_t)hlen + 1);
        if (fread(buf, 1, hlen, fp) != hlen) {
            free(buf);
            snprintf(err, errcap, "%s: truncated BLOW5 header", path);
            goto fail;
        }
        buf[hlen] = 0;
        parse_text_header(f, buf);
        free(buf);
        /* map the file; when that is not possible (a pipe, ...) the stdio reader below is used */
        struct stat st;
        const long pos = ftell(fp);
        if (pos > 0 && fstat(fileno(fp), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > pos) {
            void *m = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fileno(fp), 0);
            if (m != MAP_FAILED) {
                f->map = (const uint8_t *)m;
                f->map_len = (size_t)st.st_size;
                f->map_pos = (size_t)pos;
                madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
            }
        }
        return f;
    }
    /* ASCII: header lines start with '#' or '@' */
    rewind(fp);
    f->binary = 0;
    for (;;) {
        int c = fgetc(fp);
        if (c == EOF)
            break;
        ungetc(c, fp);
        if (c != '#' && c != '@')
            break;
        ssize_t len = getline(&f->line, &f->line_cap, fp);
        if (len <= 0)
            break;
        if (f->line[0] == '@')
            hdr_add_line(f, f->line);
        else if (!strncmp(f->line, "#num_read_groups\t", 17))
            set_read_groups(f, f->line + 17);
        else if (!strncmp(f->line, "#slow5_version\t", 15)) {
            unsigned a = 0, b = 0, c2 = 0;
            sscanf(f->line + 15, "%u.%u.%u", &a, &b, &c2);
            f->ver[0] = (uint8_t)a; f->ver[1] = (uint8_t)b; f->ver[2] = (uint8_t)c2;
        }
    }
    if (f->ver[0] == 0 && f->ver[1] == 0 && f->ver[2] == 0 && f->n_attr == 0) {
        snprintf(err, errcap, "%s: neither a BLOW5 nor a SLOW5 file", path);
        goto fail;
    }
    return f;
fail:
    sf_s5_close(f);
    return NULL;
}

void sf_s5_close(sf_s5file_t *f)
{
    if (!f)
        return;
    if (f->map)
        munmap((void *)f->map, f->map_len);
    if (f->fp)
        fclose(f->fp);
    for (int i = 0; i < f->n_attr; i++) {
        free(f->attr_name[i]);
        for (uint32_t g = 0; g < f->attr_width[i]; g++)
            free(f->attr_val[i][g]);
        free(f->attr_val[i]);
    }
    free(f->attr_name);
    free(f->attr_val);
    free(f->attr_width);
    free(f->line);
    free(f);
}

const char *sf_s5_hdr_get(const sf_s5file_t *f, const char *attr, uint32_t read_group)
{
    if (read_group >= f->num_read_groups)
        return NULL;
    for (int i = 0; i < f->n_attr; i++)
        if (!strcmp(f->attr_name[i], attr))
            return read_group < f->attr_width[i] ? f->attr_val[i][read_group] : NULL;
    return NULL;
}

uint32_t sf_s5_num_read_groups(const sf_s5file_t *f) { return f->num_read_groups; }
int sf_s5_is_binary(const sf_s5file_t *f) { return f->binary; }
const char *sf_s5_error(const sf_s5file_t *f) { return f->errmsg; }

static int reserve(char **mem, size_t *cap, size_t need)
{
    if (need <= *cap)
        return 0;
    size_t nc = need + need / 4 + 64;
    char *p = (char *)realloc(*mem, nc);
    if (!p)
        return -1;
    *mem = p;
    *cap = nc;
    return 0;
}

int sf_s5_is_mapped(const sf_s5file_t *f) { return f->map != NULL; }

/* maps the pages of [ptr, ptr + len) of the mapped file into this process now (from the loader thread), so that the
 * threads that copy the records to the GPUs' staging buffers do not take the page faults; best effort */
void sf_s5_prefault(const sf_s5file_t *f, const char *ptr, size_t len)
{
#ifdef MADV_POPULATE_READ
    if (!f->map || !len)
        return;
    const uintptr_t page = 4096, a = (uintptr_t)ptr & ~(page - 1), e = ((uintptr_t)ptr + len + page - 1) & ~(page - 1);
    const uintptr_t lo = (uintptr_t)f->map, hi = ((uintptr_t)f->map + f->map_len + page - 1) & ~(page - 1);
    if (a >= lo && e <= hi)
        (void)madvise((void *)a, (size_t)(e - a), MADV_POPULATE_READ);
#else
    (void)f; (void)ptr; (void)len;
#endif
}

/* next raw record of a mapped binary file as a view into the mapping (valid until sf_s5_close): no copy, the pages
 * are touched by whoever reads the record.  Returns its size, 0 at end of file, < 0 on error. */
int64_t sf_s5_get_next_view(sf_s5file_t *f, const char **ptr)
{
    if (!f->map)
        return -1;
    const size_t left = f->map_len - f->map_pos;
    if (left < 8) {
        if (left == 5 && !memcmp(f->map + f->map_pos, "5WOLB", 5))
            return 0;
        set_err(f, left == 0 ? "BLOW5 end-of-file marker missing (truncated file?)" : "truncated BLOW5 record header");
        return -2;
    }
    uint64_t sz;
    memcpy(&sz, f->map + f->map_pos, 8);
    if (sz == 0 || sz > ((uint64_t)1 << 40)) {
        set_err(f, "corrupt BLOW5 record size");
        return -2;
    }
    if (sz > left - 8) {
        set_err(f, "truncated BLOW5 record");
        return -2;
    }
    *ptr = (const char *)(f->map + f->map_pos + 8);
    f->map_pos += 8 + (size_t)sz;
    return (int64_t)sz;
}

int64_t sf_s5_get_next_mem(sf_s5file_t *f, char **mem, size_t *cap)
{
    if (f->map) { /* keep the two readers on one cursor */
        const char *p = NULL;
        const int64_t n = sf_s5_get_next_view(f, &p);
        if (n <= 0)
            return n;
        if (reserve(mem, cap, (size_t)n)) {
            set_err(f, "out of memory");
            return -2;
        }
        memcpy(*mem, p, (size_t)n);
        return n;
    }
    if (f->binary) {
        uint64_t sz = 0;
        unsigned char b[8];
        size_t n = fread(b, 1, 8, f->fp);
        if (n < 8) {
            if (n == 5 && !memcmp(b, "5WOLB", 5))
                return 0;
            if (n == 0 && feof(f->fp)) {
                set_err(f, "BLOW5 end-of-file marker missing (truncated file?)");
                return -2;
            }
            set_err(f, "truncated BLOW5 record header");
            return -2;
        }
        memcpy(&sz, b, 8);
        if (sz == 0 || sz > ((uint64_t)1 << 40)) {
            set_err(f, "corrupt BLOW5 record size");
            return -2;
        }
        if (reserve(mem, cap, (size_t)sz)) {
            set_err(f, "out of memory");
            return -2;
        }
        if (fread(*mem, 1, (size_t)sz, f->fp) != sz) {
            set_err(f, "truncated BLOW5 record");
            return -2;
        }
        return (int64_t)sz;
    }
    ssize_t len = getline(&f->line, &f->line_cap, f->fp);
    if (len <= 0)
        return 0;
    if (reserve(mem, cap, (size_t)len + 1)) {
        set_err(f, "out of memory");
        return -2;
    }
    memcpy(*mem, f->line, (size_t)len + 1);
    return (int64_t)len;
}

/* the reference keeps the sample count of a read in an int32 (nsample, src/sigfish.c:339): longer signals are
 * rejected here instead of being carried as wrapped sizes */
#define SF_S5_MAX_SAMPLES ((uint64_t)0x7fffffff - 64)

static int rec_reserve_signal(sf_rec_t *r, size_t n)
{
    if ((uint64_t)n > SF_S5_MAX_SAMPLES)
        return -1;
    if (n <= r->cap_signal)
        return 0;
    int16_t *p = (int16_t *)realloc(r->raw_signal, sizeof(int16_t) * (n + 16));
    if (!p)
        return -1;
    r->raw_signal = p;
    r->cap_signal = n + 16;
    return 0;
}

/* StreamVByte + zigzag delta -> int16 (the signal codec of BLOW5 "svb-zd") */
static int svb_zd_decode(const uint8_t *in, size_t n_in, sf_rec_t *r)
{
    if (n_in < 4)
        return -1;
    uint32_t count;
    memcpy(&count, in, 4);
    const uint8_t *key = in + 4;
    const size_t n_key = ((size_t)count + 3) / 4;
    if (4 + n_key > n_in)
        return -1;
    const uint8_t *data = key + n_key;
    const uint8_t *end = in + n_in;
    if (rec_reserve_signal(r, count))
        return -1;
    int32_t prev = 0;
    int16_t *out = r->raw_signal;
    uint32_t i = 0;
    /* four values per key byte, branch-free, while a 4-byte load per value cannot run past the input (a key
     * byte covers at most 16 data bytes, the last load starts at most 12 bytes in and reads 4) */
    static const uint32_t mask[4] = {0xffu, 0xffffu, 0xffffffu, 0xffffffffu};
    for (; i + 4 <= count && data + 16 <= end; i += 4) {
        const unsigned kb = key[i >> 2];
#define SF_SVB_ONE(j)                                                     \
        do {                                                              \
            const unsigned code = (kb >> (2 * (j))) & 3u;                 \
            uint32_t v;                                                   \
            memcpy(&v, data, 4);                                          \
            v &= mask[code];                                              \
            data += code + 1;                                             \
            prev += (int32_t)(v >> 1) ^ -(int32_t)(v & 1);                \
            out[i + (j)] = (int16_t)prev;                                 \
        } while (0)
        SF_SVB_ONE(0); SF_SVB_ONE(1); SF_SVB_ONE(2); SF_SVB_ONE(3);
#undef SF_SVB_ONE
    }
    for (; i < count; i++) {
        const unsigned code = (key[i >> 2] >> ((i & 3) * 2)) & 3u;
        if (data + code + 1 > end)
            return -1;
        uint32_t v = data[0];
        if (code >= 1) v |= (uint32_t)data[1] << 8;
        if (code >= 2) v |= (uint32_t)data[2] << 16;
        if (code >= 3) v |= (uint32_t)data[3] << 24;
        data += code + 1;
        const int32_t d = (int32_t)(v >> 1) ^ -(int32_t)(v & 1);
        prev += d;
        out[i] = (int16_t)prev;
    }
    r->len_raw_signal = count;
    return 0;
}

static int parse_binary(const sf_s5file_t *f, const char *mem, size_t bytes, sf_rec_t *r, char **scratch, size_t *scratch_cap)
{
    const uint8_t *p = (const uint8_t *)mem;
    size_t n = bytes;
    if (f->record_press == 1) {
        /* zlib stream of unknown inflated size: grow until it fits.  The scratch buffer starts with this
         * thread's decode tables (sfinflate.h), the inflated record follows. */
        const size_t hdr = (sizeof(sf_inflater) + 63) & ~(size_t)63;
        size_t cap = *scratch_cap > hdr ? *scratch_cap - hdr : bytes * 4 + 1024;
        int use_zlib = 0;
        for (;;) {
            const int fresh = *scratch_cap == 0;
            if (reserve(scratch, scratch_cap, hdr + cap))
                return -1;
            if (fresh)
                memset(*scratch, 0, hdr);
            uint8_t *out = (uint8_t *)*scratch + hdr;
            const size_t out_cap = *scratch_cap - hdr;
            size_t got = 0;
            int rc;
            if (!use_zlib) {
                rc = sf_zlib_inflate((sf_inflater *)*scratch, (const uint8_t *)mem, bytes, out, out_cap, &got);
                if (rc < 0) { /* let zlib have the last word on streams the fast decoder rejects */
                    use_zlib = 1;
                    continue;
                }
            } else {
                z_stream zs;
                memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, MAX_WBITS) != Z_OK)
                    return -1;
                zs.next_in = (Bytef *)mem;
                zs.avail_in = (uInt)bytes;
                zs.next_out = (Bytef *)out;
                zs.avail_out = (uInt)out_cap;
                const int zrc = inflate(&zs, Z_FINISH);
                got = zs.total_out;
                inflateEnd(&zs);
                if (zrc == Z_STREAM_END)
                    rc = 0;
                else if ((zrc == Z_BUF_ERROR || zrc == Z_OK) && zs.avail_out == 0)
                    rc = 1; /* output full: grow */
                else
                    return -1; /* input exhausted with room left (truncated stream) or corrupt data */
            }
            if (rc == 0) {
                p = out;
                n = got;
                break;
            }
            if (out_cap > ((size_t)1 << 33))
                return -1; /* no record inflates to more than 2^31 samples */
            cap = out_cap * 2; /* output did not fit */
        }
    }
    size_t o = 0;
    uint16_t idl;
    if (n < 2)
        return -1;
    memcpy(&idl, p, 2);
    o = 2;
    if ((size_t)idl + 4 + 32 + 8 > n - o)
        return -1;
    free(r->read_id);
    r->read_id = (char *)malloc((size_t)idl + 1);
    if (!r->read_id)
        return -1;
    memcpy(r->read_id, p + o, idl);
    r->read_id[idl] = 0;
    o += idl;
    o += 4; /* read_group */
    memcpy(&r->digitisation, p + o, 8); o += 8;
    memcpy(&r->offset, p + o, 8); o += 8;
    memcpy(&r->range, p + o, 8); o += 8;
    memcpy(&r->sampling_rate, p + o, 8); o += 8;
    uint64_t len;
    memcpy(&len, p + o, 8); o += 8;
    if (f->signal_press == 0) {
        if (len > (n - o) / 2) /* no arithmetic on len before this: it comes from the file */
            return -1;
        if (rec_reserve_signal(r, (size_t)len))
            return -1;
        if (len) /* an empty read may have no buffer at all */
            memcpy(r->raw_signal, p + o, (size_t)len * 2);
        r->len_raw_signal = len;
    } else {
        if (len > n - o)
            return -1;
        if (svb_zd_decode(p + o, (size_t)len, r))
            return -1;
    }
    return 0;
}

/* Reads only the head of a binary record: read id, scaling, sample count, and where the signal field lies in the
 * decompressed record (the device decodes the rest, sfgpu_submit_records).  A zlib record is inflated just far
 * enough (its first SF_S5_HEAD_MAX output bytes). */
#define SF_S5_HEAD_MAX 384 /* 2 + id (<= 255) + 4 + 32 + 8 + 4 bytes of the svb-zd stream, with room to spare */
int sf_s5_parse_head(const sf_s5file_t *f, const char *mem, size_t bytes, sf_rec_t *r, int32_t *sig_pos, int64_t *sig_bytes,
                     char **scratch, size_t *scratch_cap)
{
    if (!f->binary)
        return -1;
    const uint8_t *p = (const uint8_t *)mem;
    size_t n = bytes;
    int complete = 1; /* p[0 .. n) is the whole record */
    if (f->record_press == 1) {
        const size_t hdr = (sizeof(sf_inflater) + 63) & ~(size_t)63;
        const int fresh = *scratch_cap == 0;
        if (reserve(scratch, scratch_cap, hdr + SF_S5_HEAD_MAX + 16))
            return -1;
        if (fresh)
            memset(*scratch, 0, hdr);
        uint8_t *out = (uint8_t *)*scratch + hdr;
        size_t got = 0;
        /* id length + id + read group, 4 doubles, signal length, the sample count of an svb-zd stream */
        const int rc = sf_zlib_inflate_prefix((sf_inflater *)*scratch, (const uint8_t *)mem, bytes, out, SF_S5_HEAD_MAX, 4 + 32 + 8 + 4, &got)
                           ? 1
                           : sf_zlib_inflate((sf_inflater *)*scratch, (const uint8_t *)mem, bytes, out, SF_S5_HEAD_MAX, &got);
        if (rc < 0)
            return -1; /* the caller falls back to the full decoder (which lets zlib have the last word) */
        complete = rc == 0;
        p = out;
        n = got;
    }
    uint16_t idl;
    if (n < 2)
        return -1;
    memcpy(&idl, p, 2);
    size_t o = 2;
    if ((size_t)idl + 4 + 32 + 8 > n - o)
        return -1;
    free(r->read_id);
    r->read_id = (char *)malloc((size_t)idl + 1);
    if (!r->read_id)
        return -1;
    memcpy(r->read_id, p + o, idl);
    r->read_id[idl] = 0;
    o += idl;
    o += 4; /* read_group */
    memcpy(&r->digitisation, p + o, 8); o += 8;
    memcpy(&r->offset, p + o, 8); o += 8;
    memcpy(&r->range, p + o, 8); o += 8;
    memcpy(&r->sampling_rate, p + o, 8); o += 8;
    uint64_t len;
    memcpy(&len, p + o, 8); o += 8;
    *sig_pos = (int32_t)o;
    if (f->signal_press == 0) {
        if (len > SF_S5_MAX_SAMPLES || (complete && len > (n - o) / 2))
            return -1;
        r->len_raw_signal = len;
        *sig_bytes = (int64_t)len * 2;
    } else {
        if (len > ((uint64_t)1 << 33) || (complete && len > n - o))
            return -1;
        uint32_t count = 0;
        if (len >= 4) {
            if (n - o < 4)
                return -1;
            memcpy(&count, p + o, 4);
        } else if (len != 0) {
            return -1;
        }
        if ((uint64_t)count > SF_S5_MAX_SAMPLES || 4 + ((uint64_t)count + 3) / 4 > (len ? len : 4))
            return -1;
        r->len_raw_signal = count;
        *sig_bytes = (int64_t)len;
    }
    return 0;
}

uint8_t sf_s5_record_press(const sf_s5file_t *f) { return f->record_press; }
uint8_t sf_s5_signal_press(const sf_s5file_t *f) { return f->signal_press; }

static int parse_ascii(char *line, sf_rec_t *r)
{
    /* read_id, read_group, digitisation, offset, range, sampling_rate, len_raw_signal, raw_signal */
    char *col[8];
    char *s = line;
    for (int i = 0; i < 8; i++) {
        col[i] = s;
        char *t = strchr(s, '\t');
        if (!t) {
            if (i < 7)
                return -1;
            size_t l = strlen(s);
            while (l && (s[l - 1] == '\n' || s[l - 1] == '\r'))
                s[--l] = 0;
        } else {
            *t = 0;
            s = t + 1;
        }
    }
    free(r->read_id);
    r->read_id = strdup(col[0]);
    r->digitisation = strtod(col[2], NULL);
    r->offset = strtod(col[3], NULL);
    r->range = strtod(col[4], NULL);
    r->sampling_rate = strtod(col[5], NULL);
    const uint64_t len = strtoull(col[6], NULL, 10);
    /* n samples take at least 2n - 1 characters ("d,d,...,d") */
    if (len > SF_S5_MAX_SAMPLES || len > (strlen(col[7]) + 1) / 2 + 1)
        return -1;
    if (rec_reserve_signal(r, (size_t)len))
        return -1;
    const char *q = col[7];
    uint64_t i = 0;
    while (i < len && *q && *q != '\n') {
        char *e;
        long v = strtol(q, &e, 10);
        if (e == q)
            break;
        r->raw_signal[i++] = (int16_t)v;
        q = (*e == ',') ? e + 1 : e;
    }
    if (i != len)
        return -1;
    r->len_raw_signal = len;
    return 0;
}

int sf_s5_parse(const sf_s5file_t *f, char *mem, size_t bytes, sf_rec_t *rec, char **scratch, size_t *scratch_cap)
{
    if (f->binary)
        return parse_binary(f, mem, bytes, rec, scratch, scratch_cap);
    return parse_ascii(mem, rec);
}
