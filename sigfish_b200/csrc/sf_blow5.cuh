// sf_blow5.cuh -- BLOW5 records decoded on the device: zlib inflate (RFC 1950 / 1951) and the svb-zd signal codec.
//
// Replaces, for records the host has only read from the file (paths relative to /root/reference):
//   slow5lib/src/slow5.c:2575-2609, 3191-3283   slow5_rec_depress_parse(): record decompression + field parse
//   slow5lib/src/slow5_press.c:1085-1133        zlib inflate of a record / svb-zd (StreamVByte of zigzag deltas)
//   src/sigfish.c:317-328                       parse_single()
// The host keeps the few header bytes it needs for its own output (read id, scaling, sample count: it inflates
// just the head of each record); the whole inflate + signal decode of every record runs here, one warp per
// record, so that the host feeds compressed bytes (about half the size of the int16 samples) at file-reading speed.
//
// Inflate, one warp per record.  Decoding a Huffman stream is a serial chain (the position of a symbol depends on
// the length of the one before), so all 32 lanes run the same bit reader on the same words (broadcast loads, no
// divergence, nothing to exchange) and the warp is used where there is parallel work: building the lookup tables
// of a block (canonical codes assigned with ballots, table entries written by the lane that owns the symbol) and
// copying matches (byte k of a match comes from out[pos - dist + k % dist], which is already final for every k).
// Tables live in shared memory: an 11-bit first-level table for the literal/length code and a 9-bit one for the
// distance code; the codes longer than that are kept in a short list that is searched linearly.  Literals are stored
// by lane 0.  The grid is exactly the resident blocks (7 per SM): with one more block per SM that block ran alone
// after the others had finished (-24 % together with the funnel-shift bit reader).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#define SF_INF_WARPS 4
// first-level table bits.  Measured on 65 536 records of 4.5 k samples (tools/decode_bench.py, one call, same box):
// 10 / 8 bits 39.3 ms, 11 / 9 bits 23.4 ms, 12 / 9 bits 25.1 ms (fewer resident warps): codes of 11 bits are common
// enough in these streams that sending them through the long-code list costs more than the larger table
#ifndef SF_INF_LBITS
#define SF_INF_LBITS 11
#endif
#ifndef SF_INF_DBITS
#define SF_INF_DBITS 9
#endif
#ifndef SF_INF_MINB
#define SF_INF_MINB 1
#endif
#define SF_INF_LONG_L 288
#define SF_INF_LONG_D 32
// per-warp shared memory, in bytes: litlen table, dist table, precode table (u16 each), long-code lists (u32),
// code lengths (litlen at 0, distance at 288), next_code scratch, precode lengths
#define SF_INF_SMEM_WARP (2 * (1 << SF_INF_LBITS) + 2 * (1 << SF_INF_DBITS) + 2 * 128 + 4 * SF_INF_LONG_L + 4 * SF_INF_LONG_D + 320 + 64 + 32)

// decode status of a record
#define SF_REC_OK 0
#define SF_REC_EHEADER 1   // not a zlib stream we decode (method, preset dictionary, check bits)
#define SF_REC_EBLOCK 2    // bad block type / stored-block length
#define SF_REC_ECODE 3     // over-subscribed or unusable Huffman code, bad code-length sequence
#define SF_REC_ESYMBOL 4   // invalid symbol, distance beyond the output, input exhausted
#define SF_REC_EADLER 5    // Adler-32 mismatch
#define SF_REC_ESIZE 6     // inflated record shorter than the signal field the host announced
#define SF_REC_ESIGNAL 7   // svb-zd stream inconsistent with the announced sample count

struct sf_rec_args {
    const uint8_t *rec;        // the batch's records as read from the file, each starting on an 8-byte boundary
    const int64_t *rec_off;    // [n + 1] byte offsets into rec (multiples of 8); record i is rec_bytes[i] long
    const int64_t *rec_bytes;  // [n]
    uint8_t *scratch;          // inflated records
    const int64_t *scr_off;    // [n + 1] byte offsets into scratch; the capacity of record i is the difference
    const int64_t *sig_pos;    // [n] offset of the signal field inside the (inflated) record
    const int64_t *sig_bytes;  // [n] bytes of the signal field: svb-zd stream, or 2 * samples when not compressed
    const int64_t *sig_len;    // [n] samples
    const int64_t *sig_off;    // [n + 1] sample offsets into signal (multiples of 8)
    int16_t *signal;
    int32_t n_reads;
    int32_t record_press;      // 0 none, 1 zlib
    int32_t signal_press;      // 0 none, 1 svb-zd
    int32_t *status;           // [n]
};

__constant__ uint16_t sf_k_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t sf_k_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t sf_k_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t sf_k_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t sf_k_precode_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// The bit reader: a 64-bit window (lo, hi) over two consecutive aligned 32-bit words of the record and a bit position
// bp < 32 inside it; every lane holds the same state.  peek = one funnel shift (32 valid bits at any time), take =
// add to bp and slide the window when it crosses a word: no 64-bit shifts (two instructions each on the GPU) and no
// separate refill step -- the decode loop of a literal is ~20 instructions instead of ~45 with a 64-bit bit buffer,
// and the inflate kernel is instruction-issue bound (ncu: issue active 70 %).
struct sf_bits {
    const uint32_t *in;  // record start (8-byte aligned)
    int n_words;         // words that hold record bytes; words past them read as zero (never consumed by a
                         // well-formed stream: checked at the end).  Records are < 2 GB (checked by the caller)
    int w;               // next word to load (hi = word w - 1, lo = word w - 2)
    uint32_t lo, hi;
    int bp;
};
__device__ __forceinline__ uint32_t sf_bits_word(const sf_bits &b, int w) { return w < b.n_words ? __ldg(b.in + w) : 0u; }
__device__ __forceinline__ void sf_bits_init(sf_bits &b, const uint8_t *rec, int64_t n_in, int skip_bits)
{
    b.in = reinterpret_cast<const uint32_t *>(rec);
    b.n_words = (int)((n_in + 3) >> 2);
    b.lo = sf_bits_word(b, 0);
    b.hi = sf_bits_word(b, 1);
    b.w = 2;
    b.bp = skip_bits;
}
__device__ __forceinline__ uint32_t sf_bits_peek(const sf_bits &b) { return __funnelshift_r(b.lo, b.hi, b.bp); }
__device__ __forceinline__ void sf_bits_drop(sf_bits &b, int n) // n <= 32
{
    b.bp += n;
    if (b.bp >= 32) {
        b.lo = b.hi;
        b.hi = sf_bits_word(b, b.w);
        b.w++;
        b.bp -= 32;
    }
}
__device__ __forceinline__ uint32_t sf_bits_take(sf_bits &b, int n) // n <= 16
{
    const uint32_t v = sf_bits_peek(b) & ((1u << n) - 1u);
    sf_bits_drop(b, n);
    return v;
}
__device__ __forceinline__ void sf_bits_refill(sf_bits &) {} // the window always holds >= 32 bits
__device__ __forceinline__ int64_t sf_bits_used(const sf_bits &b) { return 32ll * (b.w - 2) + b.bp; }
// has the reader run well past the end of the record?  (a damaged stream can decode the zero padding for ever)
__device__ __forceinline__ bool sf_bits_overrun(const sf_bits &b) { return b.w > b.n_words + 4; }

// Canonical Huffman code -> first-level table (entry = symbol << 4 | length, 0 = none) plus a list of the codes
// longer than `pbits` (entry = reversed code | length << 16 | symbol << 20).  lens[] in shared memory; every lane
// calls with the same arguments.  Returns false for an over-subscribed code or too many long codes.
__device__ __forceinline__ bool sf_build_table(const uint8_t *lens, const int n_sym, const int pbits, uint16_t *table,
                                               uint32_t *longs, const int long_cap, int &n_long, uint16_t *nc, const int lane)
{
    const unsigned full = 0xffffffffu;
    // symbols per code length: lane l counts length l
    int cnt = 0;
    if (lane >= 1 && lane <= 15)
        for (int i = 0; i < n_sym; i++)
            cnt += lens[i] == lane;
    // first code of every length and the Kraft sum
    int space = 1, code = 0;
    bool bad = false;
    for (int l = 1; l <= 15; l++) {
        const int c_prev = l > 1 ? __shfl_sync(full, cnt, l - 1) : 0;
        const int c = __shfl_sync(full, cnt, l);
        code = (code + c_prev) << 1;
        if (lane == 0)
            nc[l] = (uint16_t)code;
        space = (space << 1) - c;
        bad |= space < 0;
    }
    if (bad)
        return false;
    for (int i = lane; i < (1 << pbits); i += 32)
        table[i] = 0;
    __syncwarp();
    n_long = 0;
    for (int l = 1; l <= 15; l++) {
        if (__shfl_sync(full, cnt, l) == 0)
            continue;
        int base = nc[l];
        for (int c0 = 0; c0 < n_sym; c0 += 32) {
            const int sym = c0 + lane;
            const bool match = sym < n_sym && lens[sym] == l;
            const unsigned m = __ballot_sync(full, match);
            if (match) {
                const int rank = __popc(m & ((1u << lane) - 1u));
                const uint32_t r = __brev((uint32_t)(base + rank)) >> (32 - l); // codes are packed from their top bit
                if (l <= pbits) {
                    const uint16_t e = (uint16_t)((sym << 4) | l);
                    for (uint32_t k = r; k < (1u << pbits); k += 1u << l)
                        table[k] = e;
                } else if (n_long + rank < long_cap) {
                    longs[n_long + rank] = r | ((uint32_t)l << 16) | ((uint32_t)sym << 20);
                }
            }
            base += __popc(m);
            if (l > pbits)
                n_long += __popc(m);
        }
    }
    __syncwarp();
    return n_long <= long_cap;
}

// next symbol of a code; needs >= 15 bits in the buffer.  Returns -1 when no code matches.
__device__ __forceinline__ int sf_decode_sym(sf_bits &b, const uint16_t *table, const int pbits, const uint32_t *longs, const int n_long)
{
    const uint32_t peek = sf_bits_peek(b);
    const uint16_t e = table[peek & ((1u << pbits) - 1u)];
    if (e) {
        sf_bits_drop(b, e & 15);
        return e >> 4;
    }
    for (int k = 0; k < n_long; k++) {
        const uint32_t L = longs[k];
        const int l = (L >> 16) & 15;
        if ((peek & ((1u << l) - 1u)) == (L & 0xffffu)) {
            sf_bits_drop(b, l);
            return (int)(L >> 20);
        }
    }
    return -1;
}

// Inflates one zlib stream with the whole warp.  out has room for cap bytes; bytes past cap are decoded but not
// stored.  Returns the status; n_out = bytes the stream holds.
__device__ __forceinline__ int sf_inflate_warp(const uint8_t *rec, const int64_t n_in, uint8_t *out, const int64_t cap64, int64_t &n_out,
                                               uint8_t *smem, const int lane)
{
    // positions in the output are 32-bit (a record inflates to < 2 GB: the caller checks the announced sizes, and a
    // stream that runs past 2^31 - 2^17 bytes is refused below): the symbol loop is instruction bound, 64-bit
    // compares and address arithmetic per literal cost a third of it
    const unsigned cap = (unsigned)(cap64 < 0x7ffe0000ll ? cap64 : 0x7ffe0000ll);
    uint16_t *lt = reinterpret_cast<uint16_t *>(smem);
    uint16_t *dt = lt + (1 << SF_INF_LBITS);
    uint16_t *pt = dt + (1 << SF_INF_DBITS);
    uint32_t *longl = reinterpret_cast<uint32_t *>(pt + 128);
    uint32_t *longd = longl + SF_INF_LONG_L;
    uint8_t *lens = reinterpret_cast<uint8_t *>(longd + SF_INF_LONG_D);
    uint16_t *nc = reinterpret_cast<uint16_t *>(lens + 320);
    uint8_t *plen = lens + 320 + 64;
    n_out = 0;
    if (n_in < 6)
        return SF_REC_EHEADER;
    const unsigned cmf = rec[0], flg = rec[1];
    if ((cmf & 15) != 8 || (cmf >> 4) > 7 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20))
        return SF_REC_EHEADER;
    sf_bits b;
    sf_bits_init(b, rec, n_in, 16); // the two header bytes are done
    unsigned op = 0;
    int last;
    do {
        sf_bits_refill(b);
        last = (int)sf_bits_take(b, 1);
        const int btype = (int)sf_bits_take(b, 2);
        if (btype == 3)
            return SF_REC_EBLOCK;
        if (btype == 0) { // stored: to the byte boundary, LEN, ~LEN, bytes
            sf_bits_drop(b, (8 - (b.bp & 7)) & 7);
            const uint32_t len = sf_bits_take(b, 16);
            const uint32_t nlen = sf_bits_take(b, 16);
            if ((len ^ nlen) != 0xffffu)
                return SF_REC_EBLOCK;
            if ((int64_t)len > 4ll * (b.n_words + 2 - b.w) + 16 || op > 0x7ffe0000u)
                return SF_REC_EBLOCK; // more bytes than the record has left
            for (uint32_t k = 0; k < len; k++) {
                sf_bits_refill(b);
                const uint32_t v = sf_bits_take(b, 8);
                if (lane == 0 && op < cap)
                    out[op] = (uint8_t)v;
                op++;
            }
            continue;
        }
        int n_lit = 288, n_dist = 30;
        __syncwarp();
        if (btype == 1) {
            for (int i = lane; i < 320; i += 32)
                lens[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : (i < 288 ? 8 : 5)));
        } else {
            sf_bits_refill(b);
            n_lit = (int)sf_bits_take(b, 5) + 257;
            n_dist = (int)sf_bits_take(b, 5) + 1;
            const int hclen = (int)sf_bits_take(b, 4) + 4;
            if (n_lit > 286 || n_dist > 30)
                return SF_REC_ECODE;
            if (lane < 19)
                plen[lane] = 0;
            for (int i = lane; i < 320; i += 32) // symbols the header does not mention have no code
                lens[i] = 0;
            __syncwarp();
            for (int i = 0; i < hclen; i++) {
                sf_bits_refill(b);
                const uint32_t v = sf_bits_take(b, 3);
                if (lane == 0)
                    plen[sf_k_precode_order[i]] = (uint8_t)v;
            }
            __syncwarp();
            int n_long = 0;
            if (!sf_build_table(plen, 19, 7, pt, longl, 0, n_long, nc, lane))
                return SF_REC_ECODE;
            // the code lengths of both alphabets, run-length coded with the precode (RFC 1951 3.2.7)
            int n = 0, prev = 0;
            while (n < n_lit + n_dist) {
                sf_bits_refill(b);
                if (sf_bits_overrun(b))
                    return SF_REC_ECODE;
                const int sym = sf_decode_sym(b, pt, 7, longl, 0);
                if (sym < 0)
                    return SF_REC_ECODE;
                int rep = 1, v = sym;
                if (sym == 16) {
                    if (n == 0)
                        return SF_REC_ECODE;
                    v = prev;
                    rep = 3 + (int)sf_bits_take(b, 2);
                } else if (sym == 17) {
                    v = 0;
                    rep = 3 + (int)sf_bits_take(b, 3);
                } else if (sym == 18) {
                    v = 0;
                    rep = 11 + (int)sf_bits_take(b, 7);
                }
                if (n + rep > n_lit + n_dist)
                    return SF_REC_ECODE;
                // litlen lengths go to lens[0 .. n_lit), distance lengths to lens[288 .. 288 + n_dist)
                if (lane == 0)
                    for (int k = 0; k < rep; k++) {
                        const int at = n + k;
                        lens[at < n_lit ? at : 288 + (at - n_lit)] = (uint8_t)v;
                    }
                n += rep;
                prev = v;
            }
            __syncwarp();
            if (lens[256] == 0)
                return SF_REC_ECODE; // no end-of-block code
        }
        __syncwarp();
        int nl_long = 0, nd_long = 0;
        if (!sf_build_table(lens, n_lit, SF_INF_LBITS, lt, longl, SF_INF_LONG_L, nl_long, nc, lane) ||
            !sf_build_table(lens + 288, n_dist, SF_INF_DBITS, dt, longd, SF_INF_LONG_D, nd_long, nc, lane))
            return SF_REC_ECODE;

        for (;;) {
            sf_bits_refill(b);
            if (sf_bits_overrun(b))
                return SF_REC_ESYMBOL;
            int sym = sf_decode_sym(b, lt, SF_INF_LBITS, longl, nl_long);
            if (sym < 0)
                return SF_REC_ESYMBOL;
            if (sym < 256) {
                if (lane == 0 && op < cap)
                    out[op] = (uint8_t)sym;
                op++;
                continue;
            }
            if (sym == 256)
                break;
            sym -= 257;
            if (sym >= 29)
                return SF_REC_ESYMBOL;
            const int len = sf_k_len_base[sym] + (int)sf_bits_take(b, sf_k_len_extra[sym]);
            sf_bits_refill(b);
            const int dsym = sf_decode_sym(b, dt, SF_INF_DBITS, longd, nd_long);
            if (dsym < 0 || dsym >= 30)
                return SF_REC_ESYMBOL;
            const unsigned dist = sf_k_dist_base[dsym] + sf_bits_take(b, sf_k_dist_extra[dsym]);
            if (dist > op || op > 0x7ffe0000u)
                return SF_REC_ESYMBOL;
            __syncwarp(); // the literals lane 0 stored are read by the other lanes below
            for (int k = lane; k < len; k += 32)
                if (op + k < cap)
                    out[op + k] = out[op - dist + ((unsigned)k % dist)];
            __syncwarp();
            op += len;
        }
    } while (!last);
    // every bit consumed must have come from the record
    const int64_t used_bits = sf_bits_used(b);
    if (used_bits > 8 * n_in)
        return SF_REC_ESYMBOL;
    n_out = op;
    __syncwarp();
    // Adler-32 trailer (big endian, from the next byte boundary), checked when the whole stream was stored
    const int64_t tb = (used_bits + 7) >> 3;
    if (tb + 4 > n_in)
        return SF_REC_EADLER;
    if (op <= cap) {
        // a = 1 + sum b_i, b = n + sum (n - i) b_i (mod 65521): partial sums per lane
        unsigned long long sa = 0, sb = 0;
        for (unsigned i = lane; i < op; i += 32) {
            const unsigned v = out[i];
            sa += v;
            sb += (unsigned long long)(op - i) * v;
            if ((i >> 5) % 4096 == 4095)
                sb %= 65521u;
        }
        for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        const unsigned a = (unsigned)((1 + sa) % 65521u);
        const unsigned bsum = (unsigned)(((unsigned long long)(op % 65521) + sb) % 65521u);
        const unsigned want = ((unsigned)rec[tb] << 24) | ((unsigned)rec[tb + 1] << 16) | ((unsigned)rec[tb + 2] << 8) | rec[tb + 3];
        if (((bsum << 16) | a) != want)
            return SF_REC_EADLER;
    }
    return SF_REC_OK;
}

__global__ void __launch_bounds__(32 * SF_INF_WARPS, SF_INF_MINB) sf_inflate_kernel(const sf_rec_args a)
{
    extern __shared__ __align__(16) uint8_t sf_inf_smem[];
    const int lane = threadIdx.x & 31;
    uint8_t *smem = sf_inf_smem + (threadIdx.x >> 5) * SF_INF_SMEM_WARP;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < a.n_reads; i += nw) {
        int st = SF_REC_OK;
        if (a.sig_len[i] > 0 || a.rec_bytes[i] > 0) {
            int64_t n_out = 0;
            const int64_t cap = a.scr_off[i + 1] - a.scr_off[i];
            st = sf_inflate_warp(a.rec + a.rec_off[i], a.rec_bytes[i], a.scratch + a.scr_off[i], cap, n_out, smem, lane);
            if (st == SF_REC_OK && n_out < a.sig_pos[i] + a.sig_bytes[i])
                st = SF_REC_ESIZE;
        }
        if (lane == 0)
            a.status[i] = st;
        __syncwarp();
    }
}

// Signal field -> int16 samples, one warp per record.  svb-zd: u32 count, ceil(count / 4) key bytes (2 bits per
// value: bytes - 1), then the values' bytes; value = zigzag(delta to the previous sample) (slow5_press.c:1085-1133).
// Byte offsets and the running sample are two warp scans per 32 values.
__global__ void __launch_bounds__(128) sf_signal_kernel(const sf_rec_args a)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < a.n_reads; i += nw) {
        const int64_t n = a.sig_len[i];
        int16_t *out = a.signal + a.sig_off[i];
        const int64_t padded = a.sig_off[i + 1] - a.sig_off[i];
        int st = a.record_press ? a.status[i] : SF_REC_OK;
        const uint8_t *src = (a.record_press ? a.scratch + a.scr_off[i] : a.rec + a.rec_off[i]) + a.sig_pos[i];
        if (st == SF_REC_OK && n > 0) {
            if (a.signal_press == 0) {
                if (a.sig_bytes[i] < 2 * n)
                    st = SF_REC_ESIGNAL;
                else
                    for (int64_t k = lane; k < n; k += 32)
                        out[k] = (int16_t)((unsigned)src[2 * k] | ((unsigned)src[2 * k + 1] << 8));
            } else {
                const int64_t nb = a.sig_bytes[i];
                const uint32_t count = nb >= 4 ? ((uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24)) : 0u;
                const int64_t n_key = (n + 3) >> 2;
                if ((int64_t)count != n || 4 + n_key > nb) {
                    st = SF_REC_ESIGNAL;
                } else {
                    const uint8_t *key = src + 4, *data = key + n_key;
                    const int64_t n_data = nb - 4 - n_key;
                    int64_t off = 0;    // data bytes consumed
                    uint32_t prev = 0;  // running sample (int32 arithmetic, as the CPU decoder)
                    bool bad = false;
                    for (int64_t base = 0; base < n; base += 32) {
                        const int64_t s = base + lane;
                        const int len = s < n ? (int)((key[s >> 2] >> ((s & 3) * 2)) & 3) + 1 : 0;
                        int incl = len;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int u = __shfl_up_sync(full, incl, o);
                            if (lane >= o) incl += u;
                        }
                        const int64_t at = off + incl - len;
                        uint32_t v = 0;
                        if (len > 0) {
                            if (at + len > n_data) {
                                bad = true;
                            } else {
                                v = data[at];
                                if (len > 1) v |= (uint32_t)data[at + 1] << 8;
                                if (len > 2) v |= (uint32_t)data[at + 2] << 16;
                                if (len > 3) v |= (uint32_t)data[at + 3] << 24;
                            }
                        }
                        uint32_t d = (v >> 1) ^ (0u - (v & 1u)); // zigzag
                        if (len == 0) d = 0;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t u = __shfl_up_sync(full, d, o);
                            if (lane >= o) d += u;
                        }
                        if (s < n)
                            out[s] = (int16_t)(prev + d);
                        prev += __shfl_sync(full, d, 31);
                        off += __shfl_sync(full, incl, 31);
                    }
                    if (__any_sync(full, bad))
                        st = SF_REC_ESIGNAL;
                }
            }
        }
        for (int64_t k = (st == SF_REC_OK ? n : 0) + lane; k < padded; k += 32)
            out[k] = 0;
        if (lane == 0)
            a.status[i] = st;
    }
}
