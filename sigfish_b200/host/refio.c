/* refio.c -- FASTA and k-mer model file readers of the host pipeline.
 *
 * FASTA / FASTQ: what gen_ref() gets from kseq (reference src/genref.c:100-127, src/kseq.h:184-224): the name is
 * the header up to the first white space, the sequence is every following line; plain or gzip files.
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

/* buffered characters of a (possibly gzip-compressed) file */
typedef struct {
    gzFile fp;
    unsigned char *buf;
    int n, pos, eof;
} chars_t;
#define CHARS_BUF (1 << 20)

static int chars_fill(chars_t *r)
{
    if (r->eof)
        return -1;
    r->n = gzread(r->fp, r->buf, CHARS_BUF);
    r->pos = 0;
    if (r->n <= 0) {
        r->n = 0;
        r->eof = 1;
        return -1;
    }
    return 0;
}
static inline int chars_get(chars_t *r)
{
    if (r->pos >= r->n && chars_fill(r))
        return -1;
    return r->buf[r->pos++];
}
static inline int chars_at_end(chars_t *r) { return r->pos >= r->n && chars_fill(r); }

typedef struct {
    char *s;
    size_t n, cap;
} text_t;

static void text_put(text_t *t, const unsigned char *p, size_t n)
{
    if (t->n + n + 1 > t->cap) {
        while (t->n + n + 1 > t->cap)
            t->cap = t->cap ? t->cap * 2 : 1 << 16;
        t->s = (char *)realloc(t->s, t->cap);
    }
    memcpy(t->s + t->n, p, n);
    t->n += n;
}

/* appends the rest of the current line (without its line feed) to t, or skips it when t is NULL; a carriage return
 * that ends what t then holds past `from` is dropped, as long as more than one character stands there (kseq's rule for
 * Windows line ends: it looks at the whole string it is appending to) */
static void chars_line(chars_t *r, text_t *t, size_t from)
{
    if (chars_at_end(r))
        return; /* nothing follows: kseq leaves the string as it is */
    for (;;) {
        if (r->pos >= r->n && chars_fill(r))
            break;
        const unsigned char *b = r->buf + r->pos;
        const unsigned char *nl = (const unsigned char *)memchr(b, '\n', (size_t)(r->n - r->pos));
        const size_t len = nl ? (size_t)(nl - b) : (size_t)(r->n - r->pos);
        if (t)
            text_put(t, b, len);
        r->pos += (int)len + (nl ? 1 : 0);
        if (nl)
            break;
    }
    if (t && t->n - from > 1 && t->s[t->n - 1] == '\r')
        t->n--;
}

/* The records of a FASTA or FASTQ file as the reference's reader yields them to gen_ref() (kseq_read, reference
 * src/kseq.h:184-224, src/genref.c:113): a record starts at the next '>' or '@'; its name ends at the first white
 * space, the rest of that line is a comment; the sequence is every following line up to one that starts with '>',
 * '@' or '+' (line feeds and a trailing carriage return dropped, anything else kept, other white space included);
 * after a '+' line as many quality characters as there were bases are skipped, whatever they are.  A FASTQ record
 * whose quality is missing or shorter than its sequence ends the file, and is not a record. */
int sf_fasta_read(const char *path, sf_fasta_t *out, char *err, size_t errcap)
{
    memset(out, 0, sizeof *out);
    chars_t in;
    memset(&in, 0, sizeof in);
    in.fp = gzopen(path, "r");
    if (!in.fp) {
        snprintf(err, errcap, "cannot open %s", path);
        return -1;
    }
    gzbuffer(in.fp, 1 << 20);
    in.buf = (unsigned char *)malloc(CHARS_BUF);
    text_t bases = {NULL, 0, 0}, name = {NULL, 0, 0}, qual = {NULL, 0, 0};
    int cap_r = 16, n_r = 0;
    char **names = (char **)malloc(sizeof(char *) * cap_r);
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * (cap_r + 1));
    int c, pending = 0; /* pending: the first character of the next header has been taken already */
    for (;;) {
        if (!pending) {
            while ((c = chars_get(&in)) != -1 && c != '>' && c != '@')
                ;
            if (c == -1)
                break;
        }
        pending = 0;
        if (chars_at_end(&in))
            break; /* a header character as the very last byte */
        name.n = 0;
        while ((c = chars_get(&in)) != -1 && !isspace(c)) {
            const unsigned char ch = (unsigned char)c;
            text_put(&name, &ch, 1);
        }
        if (c != -1 && c != '\n')
            chars_line(&in, NULL, 0); /* comment */
        const size_t seq0 = bases.n;
        while ((c = chars_get(&in)) != -1 && c != '>' && c != '+' && c != '@') {
            if (c == '\n')
                continue;
            const unsigned char ch = (unsigned char)c;
            text_put(&bases, &ch, 1);
            chars_line(&in, &bases, seq0);
        }
        if (c == '>' || c == '@')
            pending = 1;
        int bad = 0;
        if (c == '+') {
            const size_t n_seq = bases.n - seq0;
            while ((c = chars_get(&in)) != -1 && c != '\n')
                ;
            bad = c == -1;
            qual.n = 0;
            while (!bad && !chars_at_end(&in)) {
                chars_line(&in, &qual, 0);
                if (qual.n >= n_seq)
                    break;
            }
            bad |= qual.n != n_seq;
        }
        if (bad) { /* the reference's loop ends here without this record */
            bases.n = seq0;
            break;
        }
        if (n_r == cap_r) {
            cap_r *= 2;
            names = (char **)realloc(names, sizeof(char *) * cap_r);
            off = (int64_t *)realloc(off, sizeof(int64_t) * (cap_r + 1));
        }
        names[n_r] = (char *)malloc(name.n + 1);
        memcpy(names[n_r], name.s ? name.s : "", name.n);
        names[n_r][name.n] = 0;
        off[n_r] = (int64_t)seq0;
        n_r++;
    }
    free(in.buf);
    free(name.s);
    free(qual.s);
    gzclose(in.fp);
    if (n_r == 0) {
        free(bases.s); free(names); free(off);
        snprintf(err, errcap, "%s holds no FASTA record", path);
        return -1;
    }
    off[n_r] = (int64_t)bases.n;
    out->num_ref = n_r;
    out->names = names;
    out->bases = bases.s ? bases.s : (char *)calloc(1, 1);
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
