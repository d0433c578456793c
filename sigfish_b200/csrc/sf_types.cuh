// sf_types.cuh -- device-side data layout shared by the sigfish_b200 kernels.
//
// HBM layout:
//   ref stream  : one float array; every (contig, strand) "segment" is preceded by one
//                 sentinel column holding +INF.  A sentinel column drives every DTW row to
//                 +INF, which is exactly the virtual column -1 of cdtw.c:171-189, so segments
//                 can be streamed back to back through one wavefront with no per-column test.
//   segments    : sf_seg[], in the reference's processing order (contig ascending, '+' then '-';
//                 sigfish.c:870,890,936).
//   groups      : sf_group[], runs of consecutive segments that one warp streams through.
//   queries     : per read q_cap floats (z-scored event means, reversed for RNA), qlen.
#pragma once
#include <cstdint>

#define SF_INF __int_as_float(0x7f800000)

// option bits: same values as the reference (src/sigfish.h:30-39)
#define SF_RNA 0x001
#define SF_DTW 0x002
#define SF_INV 0x004
#define SF_REF 0x010
#define SF_END 0x020
#define SF_SAM 0x100

struct sf_seg {
    int64_t off;    // stream index of the first real column (the sentinel sits at off-1)
    int32_t rlen;   // number of columns (k-mers aligned against)
    int32_t rid;    // contig index
    int32_t strand; // 0 '+', 1 '-'
    int32_t pad;
};

struct sf_group {
    int64_t begin;      // stream index of the sentinel in front of the first segment
    int64_t end;        // one past the last real column of the last segment
    int32_t seg0;       // first segment
    int32_t nseg;       // number of segments
    int32_t ck_every;   // checkpoint period in blocks of 32 macro-steps = 64 columns (0: none)
    int32_t n_ck;       // checkpoints per task in this group
    int64_t ck_prefix;  // checkpoints of earlier groups (per read)
};

// A DTW task: a whole group, or -- for a group that is one long segment -- a *piece* of it: blocks [b0, b1) of the
// group's block grid (64 columns per block; position 0 = the group's sentinel).  A piece with a predecessor starts
// `warm_blocks` blocks early behind a virtual +INF column, and its wavefront at the piece boundary (the "warm
// front") is compared bit for bit with the front its predecessor reached there (a regular checkpoint: piece
// boundaries are multiples of ck_every).  Equal fronts => everything the piece computes from the boundary on is
// the full-matrix recurrence; a piece whose fronts differ is redone from the predecessor's front (sf_verify_kernel,
// FIX instantiations of the DTW kernels).
struct sf_piece {
    int32_t gid;     // group
    int32_t b0, b1;  // own blocks; b1 is ignored for the last piece (it runs to the end of the group)
    int32_t flags;   // bit 0: has a predecessor (warm-up, head edge); bit 1: has a successor (tail edge)
    int32_t widx;    // index of its warm front among the read's warm fronts (flags & 1)
    int32_t sidx;    // index of its group among the split groups (flags != 0)
    int32_t k;       // piece number inside the group
    int32_t pad;
};

// what one (read, piece) task reports
struct sf_taskres {
    float s1;        // best candidate score
    float s2;        // second best candidate score (value only)
    int32_t seg;     // global segment index of the best candidate
    int32_t chunk;   // chunk index inside the segment (tie order)
    int32_t pos;     // end column inside the segment (-1: no finite cell)
    // pieces only: the chunks cut by the piece's two ends are reported apart and joined with the neighbour's part
    // by the merge (first strict minimum over both parts, sigfish.c:891-901)
    float hmin;      // head edge: minimum of the part of chunk `hchunk` inside this piece
    int32_t hpos, hchunk;
    float tmin;      // tail edge
    int32_t tpos, tchunk;
    int32_t pad;
};

// per-read record produced by the event kernel and consumed by DTW + host epilogue
struct sf_readinfo {
    int64_t n_events;   // events seen (exact unless early exit; then >= qend+1)
    int32_t qstart, qend;
    int32_t qlen;       // 0: read produces no output
    int32_t status;     // bit0 ignored, bit1 too short, bit2 no peak (reference UB), bit3 inexact prefix sums handled
    uint64_t start_raw, end_raw;
};

// final per-read hit (device -> host), 32 bytes
struct sf_hit {
    float score, score2;
    int32_t rid, strand;
    int32_t pos_st, pos_end; // in-array coordinates of the winning hit (before flip/offset)
    int32_t seg, pad;
};
