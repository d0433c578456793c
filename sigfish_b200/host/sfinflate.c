/* sfinflate.c -- a fast one-shot decoder for zlib streams (RFC 1950 / 1951).
 *
 * Why: decoding BLOW5 records is what limits the read rate the host can feed to several GPUs (DESIGN.md 6,
 * tools/hostfeed): with zlib's inflate() a 4.5 k-sample record costs ~50 us of inflate + ~30 us of svb-zd on one
 * core.  Records are small, complete in memory and inflate into a buffer we own, so none of zlib's streaming
 * machinery is needed: this decoder keeps a 64-bit bit buffer refilled 8 bytes at a time, resolves a
 * literal/length code with one lookup in a 10-bit table (longer codes through a second-level table) and copies
 * matches a word at a time.  The reference reads records through slow5lib, which calls zlib's inflate
 * (slow5lib src/slow5_press.c); the output here is byte-identical and the Adler-32 trailer is verified the same
 * way.  s5read.c falls back to zlib for any stream this decoder rejects.
 *
 * Return values of sf_zlib_inflate(): 0 = ok, 1 = output buffer too small (the first *n_out bytes of the output are
 * valid: this is how the head of a record is read without inflating the rest), -1 = corrupt / unsupported stream. */
#include "sfinflate.h"

#include <string.h>

#define LITLEN_BITS SF_INFLATE_LITLEN_BITS
#define DIST_BITS SF_INFLATE_DIST_BITS
#define MAX_CODE_LEN 15
#define N_LITLEN 288
#define N_DIST 32
#define N_PRECODE 19

/* table entry: value << 16 | type << 12 | extra << 8 | nbits */
#define T_LENGTH 0u  /* value = length base, extra = extra bits */
#define T_LITERAL 1u /* value = byte */
#define T_END 2u
#define T_SUB 4u     /* value = index of the second-level table, extra = its index bits, nbits = first-level bits */
#define T_INVALID 8u
#define F_LITERAL (T_LITERAL << 12) /* one-bit tests on an entry: the types are distinct bits */
#define F_SPECIAL ((T_END | T_SUB | T_INVALID) << 12)
#define ENTRY(value, type, extra, nbits) (((uint32_t)(value) << 16) | ((uint32_t)(type) << 12) | ((uint32_t)(extra) << 8) | (uint32_t)(nbits))
#define E_VALUE(e) ((e) >> 16)
#define E_TYPE(e) (((e) >> 12) & 15u)
#define E_EXTRA(e) (((e) >> 8) & 15u)
#define E_NBITS(e) ((e) & 255u)

static const uint16_t k_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t k_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t k_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t k_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
static const uint8_t k_precode_order[N_PRECODE] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

/* the low `len` bits of code, reversed (Huffman codes are packed starting from their most significant bit) */
static inline uint32_t reverse_bits(uint32_t code, int len)
{
    static const uint8_t rev4[16] = {0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15};
    const uint32_t r16 = ((uint32_t)rev4[code & 15] << 12) | ((uint32_t)rev4[(code >> 4) & 15] << 8) |
                         ((uint32_t)rev4[(code >> 8) & 15] << 4) | (uint32_t)rev4[(code >> 12) & 15];
    return r16 >> (16 - len);
}

/* what a symbol decodes to, as a table entry without its bit count */
static uint32_t litlen_symbol(int sym)
{
    if (sym < 256)
        return ENTRY(sym, T_LITERAL, 0, 0);
    if (sym == 256)
        return ENTRY(0, T_END, 0, 0);
    if (sym <= 285)
        return ENTRY(k_len_base[sym - 257], T_LENGTH, k_len_extra[sym - 257], 0);
    return ENTRY(0, T_INVALID, 0, 0);
}

static uint32_t dist_symbol(int sym)
{
    if (sym < 30)
        return ENTRY(k_dist_base[sym], T_LENGTH, k_dist_extra[sym], 0);
    return ENTRY(0, T_INVALID, 0, 0);
}

static uint32_t precode_symbol(int sym) { return ENTRY(sym, T_LITERAL, 0, 0); }

/* Canonical Huffman code (RFC 1951 3.2.2) -> two-level lookup table indexed by the next bits of the stream
 * (LSB first).  symbol[i] is the entry of symbol i without its bit count; table has room for
 * (1 << table_bits) + sub_cap entries.  An incomplete code leaves T_INVALID entries (legal only for a single
 * distance code; hitting one while decoding is an error); an over-subscribed code is rejected.  0 or -1. */
static int build_table(const uint8_t *lens, int n_sym, int table_bits, const uint32_t *symbol, uint32_t *table, int sub_cap)
{
    int count[MAX_CODE_LEN + 1] = {0};
    uint32_t next_code[MAX_CODE_LEN + 2];
    for (int i = 0; i < n_sym; i++)
        count[lens[i]]++;
    count[0] = 0;
    uint32_t code = 0;
    int64_t space = 1; /* Kraft sum check */
    for (int l = 1; l <= MAX_CODE_LEN; l++) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        next_code[l] = code;
        space = (space << 1) - count[l];
        if (space < 0)
            return -1;
    }
    const int primary = 1 << table_bits;
    if (space != 0) /* a complete code fills every entry below */
        for (int i = 0; i < primary; i++)
            table[i] = ENTRY(0, T_INVALID, 0, 1);
    /* second-level sizes: for every first-level index, the longest code that starts with it */
    int any_long = 0;
    for (int l = table_bits + 1; l <= MAX_CODE_LEN; l++)
        any_long |= count[l];
    if (any_long) {
        uint8_t sub_len[1 << LITLEN_BITS];
        uint32_t nc[MAX_CODE_LEN + 2];
        memset(sub_len, 0, (size_t)primary);
        memcpy(nc, next_code, sizeof nc);
        for (int i = 0; i < n_sym; i++) {
            const int l = lens[i];
            if (l <= table_bits)
                continue;
            const uint32_t r = reverse_bits(nc[l]++, l);
            const uint32_t idx = r & (uint32_t)(primary - 1);
            if (l > sub_len[idx])
                sub_len[idx] = (uint8_t)l;
        }
        int next_sub = primary;
        for (int idx = 0; idx < primary; idx++) {
            if (!sub_len[idx])
                continue;
            const int sb = sub_len[idx] - table_bits;
            if (next_sub + (1 << sb) > primary + sub_cap)
                return -1;
            table[idx] = ENTRY(next_sub, T_SUB, sb, table_bits);
            for (int k = 0; k < (1 << sb); k++)
                table[next_sub + k] = ENTRY(0, T_INVALID, 0, 1);
            next_sub += 1 << sb;
        }
    }
    for (int i = 0; i < n_sym; i++) {
        const int l = lens[i];
        if (!l)
            continue;
        const uint32_t r = reverse_bits(next_code[l]++, l);
        const uint32_t e = symbol[i] | (uint32_t)l;
        if (l <= table_bits) {
            for (uint32_t k = r; k < (uint32_t)primary; k += 1u << l)
                table[k] = e;
        } else {
            const uint32_t p = table[r & (uint32_t)(primary - 1)];
            const int sb = (int)E_EXTRA(p);
            const uint32_t base = E_VALUE(p);
            for (uint32_t k = r >> table_bits; k < (1u << sb); k += 1u << (l - table_bits))
                table[base + k] = e;
        }
    }
    return 0;
}

static inline uint64_t load64le(const uint8_t *p)
{
    uint64_t v;
    memcpy(&v, p, 8);
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_BIG_ENDIAN__
    v = __builtin_bswap64(v);
#endif
    return v;
}

static uint32_t adler32_of(const uint8_t *p, size_t n)
{
    uint32_t a = 1, b = 0;
    while (n > 0) {
        size_t k = n < 5552 ? n : 5552; /* largest block for which b cannot overflow 32 bits */
        n -= k;
        /* 16 bytes at a time: b grows by 16 a + sum (16 - i) p[i]; the sums do not depend on each other */
        while (k >= 16) {
            uint32_t s1 = 0, s2 = 0;
            for (int i = 0; i < 16; i++) {
                s1 += p[i];
                s2 += (uint32_t)(16 - i) * p[i];
            }
            b += 16 * a + s2;
            a += s1;
            p += 16;
            k -= 16;
        }
        while (k--) {
            a += *p++;
            b += a;
        }
        a %= 65521u;
        b %= 65521u;
    }
    return (b << 16) | a;
}

/* the bit reader: bb holds bc bits (LSB first).  Past the end of the input it is topped up with zero bytes, counted
 * in `over`: they sit above the real bits, a well-formed stream never consumes them, and (bc >> 3) - over is the
 * number of whole real bytes still in the buffer */
#define REFILL()                                                                  \
    do {                                                                          \
        if (in_end - ip >= 8) {                                                   \
            bb |= load64le(ip) << bc;                                             \
            ip += (63 - bc) >> 3;                                                 \
            bc |= 56;                                                             \
        } else {                                                                  \
            while (bc <= 56) {                                                    \
                if (ip < in_end)                                                  \
                    bb |= (uint64_t)*ip++ << bc;                                  \
                else                                                              \
                    over++;                                                       \
                bc += 8;                                                          \
            }                                                                     \
            if (over * 8 > bc)                                                    \
                return -1; /* padding bits were consumed: the stream is truncated */ \
        }                                                                         \
    } while (0)
#define DROP(n) do { bb >>= (n); bc -= (n); } while (0)
#define BITS(n) ((uint32_t)(bb & (((uint64_t)1 << (n)) - 1)))

int sf_zlib_inflate(sf_inflater *d, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap_out, size_t *n_out)
{
    if (n_in < 6)
        return -1;
    /* RFC 1950: CMF, FLG */
    if ((in[0] & 15) != 8 || (in[0] >> 4) > 7 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20))
        return -1;
    if (!d->ready) { /* symbol -> entry maps, once per decoder */
        for (int i = 0; i < N_LITLEN; i++) d->sym_litlen[i] = litlen_symbol(i);
        for (int i = 0; i < N_DIST; i++) d->sym_dist[i] = dist_symbol(i);
        for (int i = 0; i < N_PRECODE; i++) d->sym_precode[i] = precode_symbol(i);
        d->ready = 1;
    }
    const uint8_t *ip = in + 2;
    const uint8_t *in_end = in + n_in;
    uint8_t *op = out;
    uint8_t *const out_end = out + cap_out;
    uint64_t bb = 0;
    unsigned bc = 0;
    size_t over = 0;
    int last;

    do {
        REFILL();
        last = (int)BITS(1);
        const uint32_t btype = (uint32_t)((bb >> 1) & 3);
        DROP(3);
        if (btype == 0) {
            /* stored: skip to the byte boundary, give the whole bytes of the bit buffer back */
            DROP(bc & 7);
            if ((bc >> 3) < over)
                return -1;
            ip -= (bc >> 3) - over;
            bb = 0;
            bc = 0;
            over = 0;
            if (in_end - ip < 4)
                return -1;
            const uint32_t len = (uint32_t)ip[0] | ((uint32_t)ip[1] << 8), nlen = (uint32_t)ip[2] | ((uint32_t)ip[3] << 8);
            ip += 4;
            if ((len ^ nlen) != 0xffffu || (size_t)(in_end - ip) < len)
                return -1;
            if ((size_t)(out_end - op) < len)
                { *n_out = (size_t)(op - out); return 1; }
            memcpy(op, ip, len);
            op += len;
            ip += len;
            continue;
        }
        if (btype == 3)
            return -1;
        if (btype == 1) {
            if (!d->fixed_ready) {
                uint8_t lens[N_LITLEN + N_DIST];
                int i = 0;
                for (; i < 144; i++) lens[i] = 8;
                for (; i < 256; i++) lens[i] = 9;
                for (; i < 280; i++) lens[i] = 7;
                for (; i < 288; i++) lens[i] = 8;
                for (i = 0; i < N_DIST; i++) lens[N_LITLEN + i] = 5;
                if (build_table(lens, N_LITLEN, LITLEN_BITS, d->sym_litlen, d->fixed_litlen, 0) ||
                    build_table(lens + N_LITLEN, N_DIST, DIST_BITS, d->sym_dist, d->fixed_dist, 0))
                    return -1;
                d->fixed_ready = 1;
            }
        } else {
            /* dynamic: code lengths of the two alphabets, themselves Huffman coded (RFC 1951 3.2.7) */
            const int hlit = (int)BITS(5) + 257;
            const int hdist = (int)((bb >> 5) & 31) + 1;
            const int hclen = (int)((bb >> 10) & 15) + 4;
            DROP(14);
            if (hlit > 286 || hdist > 30)
                return -1;
            uint8_t plen[N_PRECODE] = {0};
            for (int i = 0; i < hclen; i++) {
                REFILL();
                plen[k_precode_order[i]] = (uint8_t)BITS(3);
                DROP(3);
            }
            if (build_table(plen, N_PRECODE, 7, d->sym_precode, d->precode, 0))
                return -1;
            uint8_t lens[N_LITLEN + N_DIST];
            int n = 0;
            while (n < hlit + hdist) {
                REFILL();
                const uint32_t e = d->precode[BITS(7)];
                if (E_TYPE(e) != T_LITERAL)
                    return -1;
                DROP(E_NBITS(e));
                const int sym = (int)E_VALUE(e);
                if (sym < 16) {
                    lens[n++] = (uint8_t)sym;
                    continue;
                }
                int rep;
                uint8_t v = 0;
                if (sym == 16) {
                    if (n == 0)
                        return -1;
                    v = lens[n - 1];
                    rep = 3 + (int)BITS(2);
                    DROP(2);
                } else if (sym == 17) {
                    rep = 3 + (int)BITS(3);
                    DROP(3);
                } else {
                    rep = 11 + (int)BITS(7);
                    DROP(7);
                }
                if (n + rep > hlit + hdist)
                    return -1;
                memset(lens + n, v, (size_t)rep);
                n += rep;
            }
            if (lens[256] == 0)
                return -1; /* no end-of-block code */
            uint8_t ll[N_LITLEN] = {0}, dl[N_DIST] = {0};
            memcpy(ll, lens, (size_t)hlit);
            memcpy(dl, lens + hlit, (size_t)hdist);
            if (build_table(ll, N_LITLEN, LITLEN_BITS, d->sym_litlen, d->litlen, SF_INFLATE_LITLEN_SUB) ||
                build_table(dl, N_DIST, DIST_BITS, d->sym_dist, d->dist, SF_INFLATE_DIST_SUB))
                return -1;
        }
        const uint32_t *const lt = btype == 1 ? d->fixed_litlen : d->litlen;
        const uint32_t *const dt = btype == 1 ? d->fixed_dist : d->dist;

        /* Fast loop: while 16 input bytes and a longest match plus a word of output remain, nothing can run
         * out, so there are no bounds tests.  The entry of the next symbol is looked up before the refill (a
         * refill only adds bits above the ones already there). */
        if (in_end - ip >= 16 && out_end - op >= 272) {
            bb |= load64le(ip) << bc;
            ip += (63 - bc) >> 3;
            bc |= 56;
            uint32_t e = lt[BITS(LITLEN_BITS)];
            for (;;) {
                if (e & F_LITERAL) {
                    /* a first-level literal uses <= 10 bits: three of them fit between two refills */
                    bb >>= (uint8_t)e; bc -= (uint8_t)e;
                    *op++ = (uint8_t)(e >> 16);
                    e = lt[BITS(LITLEN_BITS)];
                    if (e & F_LITERAL) {
                        bb >>= (uint8_t)e; bc -= (uint8_t)e;
                        *op++ = (uint8_t)(e >> 16);
                        e = lt[BITS(LITLEN_BITS)];
                        if (e & F_LITERAL) {
                            bb >>= (uint8_t)e; bc -= (uint8_t)e;
                            *op++ = (uint8_t)(e >> 16);
                            e = lt[BITS(LITLEN_BITS)];
                        }
                    }
                } else {
                    if (e & F_SPECIAL) {
                        if (E_TYPE(e) != T_SUB)
                            break; /* end of block or invalid: the careful loop below deals with it */
                        DROP(LITLEN_BITS);
                        e = lt[E_VALUE(e) + BITS(E_EXTRA(e))];
                        if (E_TYPE(e) == T_INVALID)
                            return -1;
                        DROP(E_NBITS(e) - LITLEN_BITS);
                        if (e & F_LITERAL) {
                            *op++ = (uint8_t)(e >> 16);
                            goto fast_next;
                        }
                        if (E_TYPE(e) == T_END) {
                            e = 0;
                            goto block_done;
                        }
                    } else {
                        DROP(E_NBITS(e));
                    }
                    /* a match: <= 15 bits used since the refill, 5 + 15 + 13 more fit */
                    const uint32_t len = E_VALUE(e) + BITS(E_EXTRA(e));
                    DROP(E_EXTRA(e));
                    uint32_t f = dt[BITS(DIST_BITS)];
                    if (E_TYPE(f) == T_SUB) {
                        DROP(DIST_BITS);
                        f = dt[E_VALUE(f) + BITS(E_EXTRA(f))];
                        if (E_TYPE(f) != T_LENGTH)
                            return -1;
                        DROP(E_NBITS(f) - DIST_BITS);
                    } else {
                        if (E_TYPE(f) != T_LENGTH)
                            return -1;
                        DROP(E_NBITS(f));
                    }
                    const uint32_t dist = E_VALUE(f) + BITS(E_EXTRA(f));
                    DROP(E_EXTRA(f));
                    if (dist > (size_t)(op - out))
                        return -1;
                    const uint8_t *src = op - dist;
                    uint8_t *dst = op;
                    op += len;
                    if (dist >= 8) {
                        do {
                            memcpy(dst, src, 8);
                            dst += 8;
                            src += 8;
                        } while (dst < op);
                    } else {
                        do
                            *dst++ = *src++;
                        while (dst < op);
                    }
                fast_next:
                    if (in_end - ip < 16 || out_end - op < 272)
                        break;
                    bb |= load64le(ip) << bc;
                    ip += (63 - bc) >> 3;
                    bc |= 56;
                    e = lt[BITS(LITLEN_BITS)];
                    continue;
                }
                if (in_end - ip < 16 || out_end - op < 272)
                    break;
                bb |= load64le(ip) << bc;
                ip += (63 - bc) >> 3;
                bc |= 56;
            }
        }

        for (;;) {
            REFILL();
            uint32_t e = lt[BITS(LITLEN_BITS)];
            if (E_TYPE(e) == T_SUB) {
                DROP(LITLEN_BITS);
                e = lt[E_VALUE(e) + BITS(E_EXTRA(e))];
                if (E_TYPE(e) == T_INVALID)
                    return -1;
                DROP(E_NBITS(e) - LITLEN_BITS);
            } else {
                DROP(E_NBITS(e));
            }
            if (E_TYPE(e) == T_LITERAL) {
                if (op >= out_end)
                    { *n_out = (size_t)(op - out); return 1; }
                *op++ = (uint8_t)E_VALUE(e);
                /* >= 41 bits are left: a second and a third literal need no refill */
                e = lt[BITS(LITLEN_BITS)];
                if (E_TYPE(e) != T_LITERAL || op >= out_end)
                    continue;
                DROP(E_NBITS(e));
                *op++ = (uint8_t)E_VALUE(e);
                e = lt[BITS(LITLEN_BITS)];
                if (E_TYPE(e) != T_LITERAL || op >= out_end)
                    continue;
                DROP(E_NBITS(e));
                *op++ = (uint8_t)E_VALUE(e);
                continue;
            }
            if (E_TYPE(e) == T_END)
                break;
            if (E_TYPE(e) != T_LENGTH)
                return -1;
            /* <= 15 bits used so far: 5 extra + 15 distance + 13 extra still fit in the 56 */
            uint32_t len = E_VALUE(e) + BITS(E_EXTRA(e));
            DROP(E_EXTRA(e));
            uint32_t f = dt[BITS(DIST_BITS)];
            if (E_TYPE(f) == T_SUB) {
                DROP(DIST_BITS);
                f = dt[E_VALUE(f) + BITS(E_EXTRA(f))];
                if (E_TYPE(f) != T_LENGTH)
                    return -1;
                DROP(E_NBITS(f) - DIST_BITS);
            } else {
                if (E_TYPE(f) != T_LENGTH)
                    return -1;
                DROP(E_NBITS(f));
            }
            const uint32_t dist = E_VALUE(f) + BITS(E_EXTRA(f));
            DROP(E_EXTRA(f));
            if (dist > (size_t)(op - out))
                return -1;
            if (len > (size_t)(out_end - op))
                { *n_out = (size_t)(op - out); return 1; }
            const uint8_t *src = op - dist;
            if (dist >= 8 && (size_t)(out_end - op) >= len + 8) {
                /* whole words; may write up to 7 bytes past the match, inside the buffer */
                uint8_t *dst = op;
                const uint8_t *const stop = op + len;
                do {
                    memcpy(dst, src, 8);
                    dst += 8;
                    src += 8;
                } while (dst < stop);
            } else {
                for (uint32_t i = 0; i < len; i++)
                    op[i] = src[i];
            }
            op += len;
        }
    block_done:;
    } while (!last);

    /* trailer: Adler-32 of the output, big endian, from the next byte boundary */
    DROP(bc & 7);
    if ((bc >> 3) < over)
        return -1;
    ip -= (bc >> 3) - over;
    if (in_end - ip < 4)
        return -1;
    const uint32_t want = ((uint32_t)ip[0] << 24) | ((uint32_t)ip[1] << 16) | ((uint32_t)ip[2] << 8) | (uint32_t)ip[3];
    const size_t n = (size_t)(op - out);
    if (adler32_of(out, n) != want)
        return -1;
    *n_out = n;
    return 0;
}

/* ---- the beginning of a stream only -------------------------------------------------------------------------------
 *
 * sf_s5_parse_head() needs the first ~90 bytes of a record (id, scaling, sample count).  With the decoder above most
 * of that call is spent filling lookup tables that are then asked ~100 times (gprof on tools/hostfeed: build_table
 * 61 %).  For so few symbols a canonical code is decoded faster without tables: codes of one length are consecutive
 * integers, and left-justified to 15 bits the codes grow with their length (RFC 1951 3.2.2), so the length of the next
 * code is the first l with  window < limit[l]  (window = the next 15 bits, MSB first; limit[l] = one past the last
 * code of length l, left-justified) and the symbol is sorted[offset[l] + (window >> (15 - l)) - first[l]].  Building
 * limit / offset / sorted costs one pass over the code lengths.
 *
 * Handles the common layout only (first block dynamic, the requested bytes inside it); anything else -- and any
 * damage -- returns 0 and the caller uses sf_zlib_inflate(), which has the last word on malformed streams. */
typedef struct {
    uint32_t limit[MAX_CODE_LEN + 2]; /* [l]: (first code of length l + count[l]) << (15 - l); [16] = sentinel */
    uint16_t first[MAX_CODE_LEN + 1];
    uint16_t offset[MAX_CODE_LEN + 1];
    uint16_t sorted[N_LITLEN];
    int min_len;
} canon_code;

/* 0, or -1 for an over-subscribed code or one without symbols */
static int canon_build(canon_code *c, const uint8_t *lens, int n_sym)
{
    int count[MAX_CODE_LEN + 1] = {0};
    for (int i = 0; i < n_sym; i++)
        count[lens[i]]++;
    count[0] = 0;
    uint32_t code = 0;
    int off = 0, space = 1;
    c->min_len = 0;
    for (int l = 1; l <= MAX_CODE_LEN; l++) {
        code = (code + (uint32_t)count[l - 1]) << 1;
        space = (space << 1) - count[l];
        if (space < 0)
            return -1;
        c->first[l] = (uint16_t)code;
        c->offset[l] = (uint16_t)off;
        c->limit[l] = (code + (uint32_t)count[l]) << (MAX_CODE_LEN - l);
        off += count[l];
        if (count[l] && !c->min_len)
            c->min_len = l;
    }
    c->limit[MAX_CODE_LEN + 1] = 0xffffffffu;
    if (!c->min_len)
        return -1;
    uint16_t next[MAX_CODE_LEN + 1];
    memcpy(next, c->offset, sizeof next);
    for (int i = 0; i < n_sym; i++)
        if (lens[i])
            c->sorted[next[lens[i]]++] = (uint16_t)i;
    return 0;
}

static const uint8_t k_rev8[256] = {
#define R2(n) (n), (n) + 2 * 64, (n) + 1 * 64, (n) + 3 * 64
#define R4(n) R2(n), R2((n) + 2 * 16), R2((n) + 1 * 16), R2((n) + 3 * 16)
#define R6(n) R4(n), R4((n) + 2 * 4), R4((n) + 1 * 4), R4((n) + 3 * 4)
    R6(0), R6(2), R6(1), R6(3)
#undef R2
#undef R4
#undef R6
};

/* the symbol of the next code in the low bits of bb (>= 15 valid or zero-padded bits); *len = its length, 0 if the
 * bits are no code (incomplete code) */
static inline int canon_decode(const canon_code *c, uint64_t bb, int *len)
{
    const uint32_t w = (((uint32_t)k_rev8[bb & 255] << 8) | k_rev8[(bb >> 8) & 255]) >> 1; /* 15 bits, MSB first */
    int l = c->min_len;
    while (w >= c->limit[l])
        l++;
    if (l > MAX_CODE_LEN) {
        *len = 0;
        return 0;
    }
    *len = l;
    return c->sorted[c->offset[l] + (w >> (MAX_CODE_LEN - l)) - c->first[l]];
}

/* Decodes the first min(cap_out, 2 + u16le(out[0..2)) + tail) bytes of the zlib stream in[0 .. n_in): a u16 length,
 * that many bytes, and `tail` more (the layout of a BLOW5 record: id length, id, fixed fields).  Returns 1 with
 * *n_out = that many bytes, or 0 = not decoded here. */
int sf_zlib_inflate_prefix(sf_inflater *d, const uint8_t *in, size_t n_in, uint8_t *out, size_t cap_out, size_t tail, size_t *n_out)
{
    if (n_in < 6 || cap_out < 2)
        return 0;
    if ((in[0] & 15) != 8 || (in[0] >> 4) > 7 || ((in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20))
        return 0;
    if (!d->ready) {
        for (int i = 0; i < N_LITLEN; i++) d->sym_litlen[i] = litlen_symbol(i);
        for (int i = 0; i < N_DIST; i++) d->sym_dist[i] = dist_symbol(i);
        for (int i = 0; i < N_PRECODE; i++) d->sym_precode[i] = precode_symbol(i);
        d->ready = 1;
    }
    const uint8_t *ip = in + 2;
    const uint8_t *const in_end = in + n_in;
    uint64_t bb = 0;
    unsigned bc = 0;
    /* the same reader as above, except that running out of input is "not decoded here" */
#define PREFILL()                                                 \
    do {                                                          \
        if (in_end - ip >= 8) {                                   \
            bb |= load64le(ip) << bc;                             \
            ip += (63 - bc) >> 3;                                 \
            bc |= 56;                                             \
        } else {                                                  \
            while (bc <= 56 && ip < in_end) {                     \
                bb |= (uint64_t)*ip++ << bc;                      \
                bc += 8;                                          \
            }                                                     \
        }                                                         \
    } while (0)
#define PDROP(n)                                                  \
    do {                                                          \
        if ((unsigned)(n) > bc)                                   \
            return 0;                                             \
        bb >>= (n);                                               \
        bc -= (unsigned)(n);                                      \
    } while (0)
    PREFILL();
    if (((bb >> 1) & 3) != 2)
        return 0; /* stored / fixed first block: rare (tiny or uncompressible records) */
    PDROP(3);
    const int hlit = (int)BITS(5) + 257;
    const int hdist = (int)((bb >> 5) & 31) + 1;
    const int hclen = (int)((bb >> 10) & 15) + 4;
    PDROP(14);
    if (hlit > 286 || hdist > 30)
        return 0;
    uint8_t plen[N_PRECODE] = {0};
    for (int i = 0; i < hclen; i++) {
        PREFILL();
        plen[k_precode_order[i]] = (uint8_t)BITS(3);
        PDROP(3);
    }
    if (build_table(plen, N_PRECODE, 7, d->sym_precode, d->precode, 0))
        return 0;
    uint8_t lens[N_LITLEN + N_DIST];
    int n = 0;
    while (n < hlit + hdist) {
        PREFILL();
        const uint32_t e = d->precode[BITS(7)];
        if (E_TYPE(e) != T_LITERAL)
            return 0;
        PDROP(E_NBITS(e));
        const int sym = (int)E_VALUE(e);
        if (sym < 16) {
            lens[n++] = (uint8_t)sym;
            continue;
        }
        int rep;
        uint8_t v = 0;
        if (sym == 16) {
            if (n == 0)
                return 0;
            v = lens[n - 1];
            rep = 3 + (int)BITS(2);
            PDROP(2);
        } else if (sym == 17) {
            rep = 3 + (int)BITS(3);
            PDROP(3);
        } else {
            rep = 11 + (int)BITS(7);
            PDROP(7);
        }
        if (n + rep > hlit + hdist)
            return 0;
        memset(lens + n, v, (size_t)rep);
        n += rep;
    }
    if (lens[256] == 0)
        return 0;
    canon_code lc, dc;
    if (canon_build(&lc, lens, hlit))
        return 0;
    const int have_dist = canon_build(&dc, lens + hlit, hdist) == 0;

    size_t want = cap_out, got = 0;
    while (got < want) {
        PREFILL();
        int l;
        const int sym = canon_decode(&lc, bb, &l);
        if (!l)
            return 0;
        PDROP(l);
        if (sym < 256) {
            out[got++] = (uint8_t)sym;
        } else {
            if (sym == 256 || sym > 285 || !have_dist)
                return 0; /* the block ends inside the prefix: the general decoder deals with it */
            const uint32_t x = k_len_extra[sym - 257];
            const uint32_t len = k_len_base[sym - 257] + BITS(x);
            PDROP(x);
            const int ds = canon_decode(&dc, bb, &l);
            if (!l || ds > 29)
                return 0;
            PDROP(l);
            const uint32_t y = k_dist_extra[ds];
            const uint32_t dist = k_dist_base[ds] + BITS(y);
            PDROP(y);
            if (dist > got)
                return 0;
            for (uint32_t i = 0; i < len && got < want; i++, got++)
                out[got] = out[got - dist];
        }
        if (want == cap_out && got >= 2) {
            const size_t w = 2 + ((size_t)out[0] | ((size_t)out[1] << 8)) + tail;
            if (w < want)
                want = w;
            if (got > want)
                got = want;
        }
    }
#undef PREFILL
#undef PDROP
    *n_out = got;
    return 1;
}
