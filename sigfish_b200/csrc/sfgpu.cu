// sfgpu.cu -- libsfgpu.so: the C-ABI of include/sfgpu.h over the sm_100a kernels in this directory.
//
// Host-side responsibilities (everything else is in the kernels):
//   * reference layout: segments in the reference's processing order (sigfish.c:870,890,936),
//     each preceded by a +INF sentinel column, packed into groups that one warp streams through;
//   * batch slots: pinned staging + device buffers + one CUDA stream per slot, so that the copy of
//     batch n+1 overlaps the kernels of batch n (this replaces the pthread fan-out of thread.c);
//   * launch order per batch: H2D -> events -> DTW scores -> merge + start coordinate -> D2H.
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/sfgpu.h"
#include "sf_types.cuh"
#include "sf_events.cuh"
#include "sf_qstart.cuh"
#include "sf_blow5.cuh"
#include "sf_ref.cuh"
#include "sf_dtw.cuh"
#include "sf_pair_inst.cuh"
#include "sf_trace.cuh"
#include "sf_path.cuh"

namespace {

thread_local char g_err[512] = "";

// SFGPU_TRACE=1 in the environment: wall-clock stamps of the set-up steps on stderr (where start-up time goes)
struct sf_tracer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    sf_tracer() : on(getenv("SFGPU_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void operator()(int dev, const char *what) const
    {
        if (!on)
            return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[sfgpu::%.1f ms] gpu %d: %s\n", ms, dev, what);
    }
};
const sf_tracer g_trace;

struct sf_slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // capacities
    int32_t cap_reads = 0;
    int64_t cap_samples = 0;
    // host pinned
    int16_t *h_signal = nullptr;
    int64_t *h_off = nullptr; // [cap_reads + 1] padded offsets, then [cap_reads] lengths
    float *h_scal = nullptr;  // [3][cap_reads]
    float *h_queries = nullptr; // only for sfgpu_submit_queries
    bool queries_only = false;
    sf_readinfo *h_info = nullptr;
    sf_hit *h_hits = nullptr;
    // device
    int16_t *d_signal = nullptr;
    int64_t *d_off = nullptr;
    float *d_scal = nullptr;
    uint64_t *d_ev_start = nullptr;
    float *d_ev_mean = nullptr, *d_ev_len = nullptr;
    float *d_queries = nullptr;
    int64_t *d_polya = nullptr;
    uint64_t *d_win_start = nullptr; // --sam only
    float *d_win_len = nullptr;
    sf_readinfo *d_info = nullptr;
    sf_taskres *d_res = nullptr;
    float *d_ckpt = nullptr;
    sf_hit *d_hits = nullptr;
    unsigned int *d_counter = nullptr;  // [4] task queues of the two DTW kernels and of their redo passes
    int32_t *d_list_full = nullptr, *d_list_other = nullptr, *d_counts = nullptr; // sf_partition_kernel; [2] = fronts that differed
    int32_t *h_counts = nullptr;        // pinned copy of d_counts
    // records decoded on the device (sfgpu_submit_records): the compressed records, their inflated form, per-record
    // layout (rec_off[n+1], rec_bytes[n], scr_off[n+1], sig_pos[n], sig_bytes[n]) and decode status
    uint8_t *h_rec = nullptr, *d_rec = nullptr, *d_scratch = nullptr;
    int64_t *h_rmeta = nullptr, *d_rmeta = nullptr;
    int32_t *h_status = nullptr, *d_status = nullptr;
    size_t cap_rec = 0, cap_scratch = 0, cap_rmeta = 0;
    int64_t rec_total = 0;
    bool records = false;
    int32_t record_press = 0, signal_press = 0;
    // pieces of split segments: warm fronts, first differing piece per (read, split group); sized per batch
    float *d_warm = nullptr;
    int32_t *d_first_bad = nullptr;
    size_t cap_res = 0, cap_warm = 0, cap_bad = 0;
    int level = 0;                      // split level of the batch in flight
    // state
    int32_t n_reads = 0;
    int64_t n_samples = 0; // padded
    int64_t raw_samples = 0;
    bool busy = false, done = false, timed = false;
    sfgpu_timing_t timing;
};

} // namespace

// One way of cutting the groups into DTW tasks: `len` blocks per piece of a long single-segment group (0: every
// group is one task).  Several levels are prepared with the reference; each batch picks the one with the smallest
// expected time for its number of reads (run_stages).
struct sf_level {
    int32_t len = 0;
    std::vector<sf_piece> pieces;
    std::vector<int32_t> order, warm_piece, split_first, split_count;
    int32_t n_pieces = 0, n_warm = 0, n_split = 0;
    double blocks_per_read = 0.0; // executed blocks of one read: own blocks + warm-up + pipeline fill
    double longest = 0.0;         // blocks of the longest task
    sf_piece *d_pieces = nullptr;
    int32_t *d_order = nullptr, *d_warm_piece = nullptr, *d_split_first = nullptr, *d_split_count = nullptr;
};

struct sfgpu_ctx {
    sfgpu_opt_t opt;
    int R = 0, q_cap = 0, ev_cap = 0, ev_cap_a = 0;
    // two full-length reads per warp (sf_dtw_pair_kernel): rows per lane and register of the last row; 0: off
    int R2 = 0, RQ2 = 0;
    int ck_floats = 0;          // floats per wavefront checkpoint (max over the layouts in use)
    int pair_blocks_per_sm = 0;
    int sm_count = 0;
    float *d_level_mean = nullptr;
    // reference
    bool have_ref = false;
    int32_t num_ref = 0, n_seg = 0, n_groups = 0;
    std::vector<sf_seg> segs;
    std::vector<sf_group> groups;
    std::vector<int32_t> seg_group, order, ref_lengths;
    float *d_stream = nullptr;
    int64_t stream_len = 0;
    sf_seg *d_segs = nullptr;
    sf_group *d_groups = nullptr;
    int32_t *d_order = nullptr, *d_seg_group = nullptr;
    int64_t ck_per_read = 0, ref_columns = 0;
    int32_t min_window = 0;
    int32_t ck_min_cols = 2048; // segments longer than this get wavefront checkpoints (reserved[0] overrides)
    std::vector<sf_level> levels; // levels[0]: no splitting
    int32_t warm_blocks = 0;      // warm-up of a piece, in blocks of 64 columns
    int32_t force_level_len = 0;  // test knob (reserved[4]): > 0 piece length in checkpoint periods, < 0 never split
    std::vector<sf_slot> slots;
    int dtw_blocks_per_sm = 0;
    int inflate_blocks_per_sm = 8;
    char err[512];
};

namespace {

int fail(sfgpu_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    if (c)
        memcpy(c->err, g_err, sizeof g_err);
    return code;
}

#define SF_CUDA(c, call)                                                                          \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return fail((c), SFGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,              \
                        cudaGetErrorString(e_));                                                  \
    } while (0)

const int kRows[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 20, 24, 32};

int pick_rows(int q)
{
    const int need = (q + 31) / 32;
    for (int r : kRows)
        if (r >= need)
            return r;
    return 0;
}

template <typename T> void dfree(T *&p)
{
    if (p)
        cudaFree(p);
    p = nullptr;
}
template <typename T> void hfree(T *&p)
{
    if (p)
        cudaFreeHost(p);
    p = nullptr;
}

void slot_free_buffers(sf_slot &s)
{
    hfree(s.h_signal); hfree(s.h_off); hfree(s.h_scal); hfree(s.h_info); hfree(s.h_hits); hfree(s.h_queries);
    dfree(s.d_signal); dfree(s.d_off); dfree(s.d_scal); dfree(s.d_ev_start); dfree(s.d_ev_mean);
    dfree(s.d_ev_len); dfree(s.d_queries); dfree(s.d_info); dfree(s.d_res); dfree(s.d_ckpt);
    dfree(s.d_hits); dfree(s.d_polya); dfree(s.d_win_start); dfree(s.d_win_len);
    dfree(s.d_list_full); dfree(s.d_list_other); dfree(s.d_warm); dfree(s.d_first_bad);
    hfree(s.h_rec); dfree(s.d_rec); dfree(s.d_scratch); hfree(s.h_rmeta); dfree(s.d_rmeta); hfree(s.h_status); dfree(s.d_status);
    s.cap_rec = s.cap_scratch = s.cap_rmeta = 0;
    s.cap_reads = 0;
    s.cap_samples = 0;
    s.cap_res = s.cap_warm = s.cap_bad = 0;
}

// host_signal: the batch arrives as int16 samples (a pinned staging buffer is needed); false when the device
// decodes the records itself -- pinning ~100 MB that nobody writes costs tens of ms per slot
int slot_reserve(sfgpu_ctx *c, sf_slot &s, int32_t n_reads, int64_t n_samples, bool host_signal = true)
{
    const bool grow = n_samples > s.cap_samples || n_reads > s.cap_reads || (host_signal && !s.h_signal);
    if (grow)
        g_trace(c->opt.device, "slot_reserve: growing buffers");
    if (n_samples > s.cap_samples) {
        const int64_t cap = std::max<int64_t>(n_samples + n_samples / 4, 1 << 20);
        hfree(s.h_signal);
        dfree(s.d_signal);
        s.cap_samples = 0;
        SF_CUDA(c, cudaMalloc(&s.d_signal, sizeof(int16_t) * cap));
        s.cap_samples = cap;
    }
    if (host_signal && !s.h_signal)
        SF_CUDA(c, cudaMallocHost(&s.h_signal, sizeof(int16_t) * s.cap_samples));
    if (n_reads > s.cap_reads) {
        int32_t cap = std::max<int32_t>(n_reads + n_reads / 4, 64);
        hfree(s.h_off); hfree(s.h_scal); hfree(s.h_info); hfree(s.h_hits); hfree(s.h_queries);
        dfree(s.d_off); dfree(s.d_scal); dfree(s.d_ev_start); dfree(s.d_ev_mean); dfree(s.d_ev_len);
        dfree(s.d_queries); dfree(s.d_info); dfree(s.d_ckpt); dfree(s.d_hits); dfree(s.d_polya); dfree(s.d_win_start); dfree(s.d_win_len);
        dfree(s.d_list_full); dfree(s.d_list_other);
        s.cap_reads = 0;
        {
            // the per-read device buffers of a slot (the wavefront checkpoints are the bulk: ~0.66 MB per read on
            // a long reference): refuse a batch that cannot fit instead of dying inside cudaMalloc, and give up the
            // 25 % growth margin when only the batch itself fits
            const double per_read = (double)c->ev_cap * 16.0 + (double)c->q_cap * 4.0 * ((c->opt.flags & SFGPU_SAM) ? 4.0 : 1.0) +
                                    (double)c->ck_per_read * c->ck_floats * 4.0 + 256.0;
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                if (per_read * (double)n_reads > 0.95 * (double)free_b)
                    return fail(c, SFGPU_ELIMIT, "a batch of %d reads needs %.1f GB of device memory per slot, %.1f GB are free: "
                                "use a smaller batch (-K)", n_reads, per_read * n_reads / 1e9, free_b / 1e9);
                if (per_read * (double)cap > 0.95 * (double)free_b)
                    cap = std::max<int32_t>(n_reads, 64);
            }
        }
        const size_t n = (size_t)cap;
        SF_CUDA(c, cudaMallocHost(&s.h_off, sizeof(int64_t) * (2 * n + 1)));
        SF_CUDA(c, cudaMallocHost(&s.h_scal, sizeof(float) * 3 * n));
        SF_CUDA(c, cudaMallocHost(&s.h_info, sizeof(sf_readinfo) * n));
        SF_CUDA(c, cudaMallocHost(&s.h_hits, sizeof(sf_hit) * n));
        SF_CUDA(c, cudaMalloc(&s.d_off, sizeof(int64_t) * (2 * n + 1)));
        SF_CUDA(c, cudaMalloc(&s.d_scal, sizeof(float) * 3 * n));
        SF_CUDA(c, cudaMalloc(&s.d_ev_start, sizeof(uint64_t) * n * c->ev_cap));
        SF_CUDA(c, cudaMalloc(&s.d_ev_mean, sizeof(float) * n * c->ev_cap));
        SF_CUDA(c, cudaMalloc(&s.d_ev_len, sizeof(float) * n * c->ev_cap));
        SF_CUDA(c, cudaMalloc(&s.d_queries, sizeof(float) * n * c->q_cap));
        SF_CUDA(c, cudaMalloc(&s.d_info, sizeof(sf_readinfo) * n));
        if (c->ck_per_read > 0)
            SF_CUDA(c, cudaMalloc(&s.d_ckpt, sizeof(float) * n * c->ck_per_read * (size_t)c->ck_floats));
        SF_CUDA(c, cudaMalloc(&s.d_hits, sizeof(sf_hit) * n));
        SF_CUDA(c, cudaMalloc(&s.d_polya, sizeof(int64_t) * n));
        SF_CUDA(c, cudaMalloc(&s.d_list_full, sizeof(int32_t) * n));
        SF_CUDA(c, cudaMalloc(&s.d_list_other, sizeof(int32_t) * n));
        if (c->opt.flags & SFGPU_SAM) {
            SF_CUDA(c, cudaMalloc(&s.d_win_start, sizeof(uint64_t) * n * c->q_cap));
            SF_CUDA(c, cudaMalloc(&s.d_win_len, sizeof(float) * n * c->q_cap));
        }
        s.cap_reads = cap;
    }
    if (grow)
        g_trace(c->opt.device, "slot_reserve: done");
    return SFGPU_OK;
}

// ---- kernel dispatch over the register tile height R and the recurrence variant ----

template <int R, bool STD> int dtw_occupancy(size_t smem)
{
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sf_dtw_score_kernel<R, STD, false>, SF_DTW_THREADS, smem);
    return nb;
}

template <int R, bool STD>
cudaError_t launch_dtw(const sf_dtw_args &a, int grid, size_t smem, cudaStream_t st)
{
    sf_dtw_score_kernel<R, STD><<<grid, SF_DTW_THREADS, smem, st>>>(a);
    return cudaGetLastError();
}

// redo pass of the warp-per-read kernel (pieces whose warm front did not verify); subsequence DTW only
template <int R, bool STD>
cudaError_t launch_dtw_fix(const sf_dtw_args &a, int grid, size_t smem, cudaStream_t st)
{
    if constexpr (!STD) {
        sf_dtw_score_kernel<R, false, true><<<grid, SF_DTW_THREADS, smem, st>>>(a);
        return cudaGetLastError();
    }
    return cudaErrorInvalidValue;
}

template <int R, bool STD> cudaError_t launch_path(const sf_path_args &a, cudaStream_t st)
{
    const int warps = 4;
    const int grid = (a.n_reads + warps - 1) / warps;
    sf_path_kernel<R, STD><<<grid, warps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

// the reads the pair kernel aligned (R2 rows per lane, 16 lanes per read) are traced two per warp
template <int R2, bool STD> cudaError_t launch_trace_pair(const sf_trace_args &a, cudaStream_t st)
{
    const int warps = 4;
    const int units = (a.n_reads + 1) / 2;
    sf_trace_pair_kernel<R2, STD><<<(units + warps - 1) / warps, warps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

template <int R, bool STD> cudaError_t launch_trace(const sf_trace_args &a, cudaStream_t st, int r2)
{
    const int warps = 4;
    const int grid = (a.n_reads + warps - 1) / warps;
    if constexpr (R >= 3 && R <= 8) { // pairing exists for 64 < q <= 256, i.e. R = 3..8 in the warp-per-read layout
        if (r2 > 0) {
            // the paired reads are skipped by the general kernel (its third parameter only says that there are any)
            sf_trace_kernel<R, STD, 16><<<grid, warps * 32, 0, st>>>(a);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess)
                return e;
            if (r2 == 16)
                return launch_trace_pair<16, STD>(a, st);
            if constexpr (!STD) {
                switch (r2) {
                case 5: return launch_trace_pair<5, false>(a, st);
                case 6: return launch_trace_pair<6, false>(a, st);
                case 7: return launch_trace_pair<7, false>(a, st);
                case 8: return launch_trace_pair<8, false>(a, st);
                case 9: return launch_trace_pair<9, false>(a, st);
                case 10: return launch_trace_pair<10, false>(a, st);
                case 11: return launch_trace_pair<11, false>(a, st);
                case 12: return launch_trace_pair<12, false>(a, st);
                case 13: return launch_trace_pair<13, false>(a, st);
                case 14: return launch_trace_pair<14, false>(a, st);
                case 15: return launch_trace_pair<15, false>(a, st);
                }
            }
            return cudaErrorInvalidValue;
        }
    }
    sf_trace_kernel<R, STD, 0><<<grid, warps * 32, 0, st>>>(a);
    return cudaGetLastError();
}

// Calls f(integral_constant<int, R>, bool_constant<STD>) for the instantiated register-tile height equal to
// `rows` (kRows) and the recurrence variant; returns false when `rows` is not instantiated.
template <int... Rs> struct sf_row_list {};
using sf_rows = sf_row_list<1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 20, 24, 32>;

template <typename F, int... Rs> bool dispatch_rows(int rows, bool std_dtw, F &&f, sf_row_list<Rs...>)
{
    return ((rows == Rs && (std_dtw ? (f(std::integral_constant<int, Rs>{}, std::true_type{}), true)
                                    : (f(std::integral_constant<int, Rs>{}, std::false_type{}), true))) || ...);
}
#define SF_DISPATCH_R(R_, STD_, EXPR)                                                             \
    dispatch_rows((R_), (STD_), [&](auto r_tag_, auto std_tag_) {                                 \
        constexpr int R = decltype(r_tag_)::value;                                                \
        constexpr bool STD = decltype(std_tag_)::value;                                           \
        EXPR;                                                                                     \
    }, sf_rows{})

int pick_level(const sfgpu_ctx *c, int n);
int slot_reserve_level(sfgpu_ctx *c, sf_slot &s, int n, const sf_level &lv);

// the device stages of one batch, on the slot's stream
int run_stages(sfgpu_ctx *c, sf_slot &s, bool with_h2d, bool with_events = true)
{
    const int n = s.n_reads;
    const bool std_dtw = (c->opt.flags & SFGPU_DTW) != 0;
    cudaStream_t st = s.stream;
    SF_CUDA(c, cudaEventRecord(s.ev[0], st));
    if (with_h2d && n > 0 && !with_events) {
        SF_CUDA(c, cudaMemcpyAsync(s.d_queries, s.h_queries, sizeof(float) * (size_t)n * c->q_cap, cudaMemcpyHostToDevice, st));
        SF_CUDA(c, cudaMemcpyAsync(s.d_info, s.h_info, sizeof(sf_readinfo) * n, cudaMemcpyHostToDevice, st));
    }
    if (with_h2d && n > 0 && with_events) {
        if (s.records) {
            SF_CUDA(c, cudaMemcpyAsync(s.d_rec, s.h_rec, (size_t)s.rec_total, cudaMemcpyHostToDevice, st));
            SF_CUDA(c, cudaMemcpyAsync(s.d_rmeta, s.h_rmeta, sizeof(int64_t) * (5 * (size_t)n + 2), cudaMemcpyHostToDevice, st));
        } else {
            SF_CUDA(c, cudaMemcpyAsync(s.d_signal, s.h_signal, sizeof(int16_t) * s.n_samples, cudaMemcpyHostToDevice, st));
        }
        SF_CUDA(c, cudaMemcpyAsync(s.d_off, s.h_off, sizeof(int64_t) * (2 * (size_t)n + 1), cudaMemcpyHostToDevice, st));
        SF_CUDA(c, cudaMemcpyAsync(s.d_scal, s.h_scal, sizeof(float) * 3 * (size_t)s.cap_reads, cudaMemcpyHostToDevice, st));
    }
    SF_CUDA(c, cudaEventRecord(s.ev[1], st));
    s.timing.dtw_launches = 0;
    s.timing.other_launches = 0;
    if (with_h2d && n > 0 && with_events && s.records) {
        // BLOW5 records -> int16 samples on the device (sf_blow5.cuh); counted with the event stage
        sf_rec_args ra;
        ra.rec = s.d_rec;
        ra.rec_off = s.d_rmeta;
        ra.rec_bytes = s.d_rmeta + (n + 1);
        ra.scr_off = s.d_rmeta + (2 * (size_t)n + 1);
        ra.sig_pos = s.d_rmeta + (3 * (size_t)n + 2);
        ra.sig_bytes = s.d_rmeta + (4 * (size_t)n + 2);
        ra.scratch = s.d_scratch;
        ra.sig_off = s.d_off;
        ra.sig_len = s.d_off + (n + 1);
        ra.signal = s.d_signal;
        ra.n_reads = n;
        ra.record_press = s.record_press;
        ra.signal_press = s.signal_press;
        ra.status = s.d_status;
        SF_CUDA(c, cudaMemsetAsync(s.d_status, 0, sizeof(int32_t) * (size_t)n, st));
        if (s.record_press) {
            const int blocks = std::max(1, std::min((n + SF_INF_WARPS - 1) / SF_INF_WARPS, c->sm_count * c->inflate_blocks_per_sm));
            sf_inflate_kernel<<<blocks, 32 * SF_INF_WARPS, SF_INF_WARPS * SF_INF_SMEM_WARP, st>>>(ra);
            SF_CUDA(c, cudaGetLastError());
            s.timing.other_launches++;
        }
        sf_signal_kernel<<<std::max(1, std::min((n + 3) / 4, c->sm_count * 16)), 128, 0, st>>>(ra);
        SF_CUDA(c, cudaGetLastError());
        s.timing.other_launches++;
    }
    if (n > 0 && with_events) {
        sf_ev_args ea;
        ea.signal = s.d_signal;
        ea.sig_off = s.d_off;
        ea.sig_len = s.d_off + (n + 1);
        ea.digitisation = s.d_scal;
        ea.offset = s.d_scal + s.cap_reads;
        ea.range = s.d_scal + 2 * (size_t)s.cap_reads;
        ea.n_reads = n;
        ea.flags = c->opt.flags;
        ea.q = c->opt.query_size;
        ea.p = c->opt.prefix_size;
        ea.ev_cap = c->ev_cap;
        ea.ev_start = s.d_ev_start;
        ea.ev_mean = s.d_ev_mean;
        ea.ev_len = s.d_ev_len;
        ea.queries = s.d_queries;
        ea.q_cap = c->q_cap;
        ea.info = s.d_info;
        ea.keep_all = 0;
        ea.polya_end = s.d_polya;
        ea.cap_a = c->ev_cap_a;
        ea.win_start = s.d_win_start;
        ea.win_len = s.d_win_len;
        if (c->opt.prefix_size < 0) {
            sf_qs_args qa;
            qa.signal = ea.signal;
            qa.sig_off = ea.sig_off;
            qa.sig_len = ea.sig_len;
            qa.digitisation = ea.digitisation;
            qa.offset = ea.offset;
            qa.range = ea.range;
            qa.n_reads = n;
            qa.rna004 = c->opt.pore == 2;
            qa.polya_end = s.d_polya;
            sf_qstart_kernel<<<(n + 31) / 32, 32, 0, st>>>(qa);
            SF_CUDA(c, cudaGetLastError());
            s.timing.other_launches++;
        }
        sf_events_kernel<<<(n + SF_EV_READS_PER_BLOCK - 1) / SF_EV_READS_PER_BLOCK, SF_EV_THREADS, sf_events_smem_bytes(), st>>>(ea);
        SF_CUDA(c, cudaGetLastError());
        s.timing.other_launches++;
    }
    SF_CUDA(c, cudaEventRecord(s.ev[2], st));
    if (n > 0) {
        s.level = pick_level(c, n);
        const sf_level &lv = c->levels[s.level];
        int rcl = slot_reserve_level(c, s, n, lv);
        if (rcl)
            return rcl;
        SF_CUDA(c, cudaMemsetAsync(s.d_counter, 0, sizeof(unsigned int) * 4, st));
        SF_CUDA(c, cudaMemsetAsync(s.d_counts + 2, 0, sizeof(int32_t), st));
        if (lv.n_split > 0)
            SF_CUDA(c, cudaMemsetAsync(s.d_first_bad, 0x7f, sizeof(int32_t) * (size_t)n * lv.n_split, st));
        sf_dtw_args da;
        da.stream = c->d_stream;
        da.segs = c->d_segs;
        da.groups = c->d_groups;
        da.pieces = lv.d_pieces;
        da.order = lv.d_order;
        da.n_pieces = lv.n_pieces;
        da.n_reads = n;
        da.queries = s.d_queries;
        da.info = s.d_info;
        da.q_cap = c->q_cap;
        da.res = s.d_res;
        da.ckpt = s.d_ckpt;
        da.ck_per_read = c->ck_per_read;
        da.ck_floats = c->ck_floats;
        da.counter = s.d_counter;
        da.list = nullptr;
        da.n_list = nullptr;
        da.q_full = c->opt.query_size;
        da.warm = s.d_warm;
        da.n_warm = lv.n_warm;
        da.warm_blocks = c->warm_blocks;
        da.first_bad = s.d_first_bad;
        da.n_split = lv.n_split;
        da.split_first = lv.d_split_first;
        da.split_count = lv.d_split_count;
        const size_t smem = sizeof(float) * SF_DTW_WARPS * sf_smem_floats_per_warp(c->R, std_dtw);
        const size_t psmem = sizeof(float) * SF_DTW_WARPS * sf_pair_smem_floats_per_warp(std_dtw);
        const long long n_tasks = (long long)n * lv.n_pieces;
        if (n_tasks > 0x7ffffff0ll)
            return fail(c, SFGPU_ELIMIT, "batch of %d reads x %d tasks per read exceeds the task counter; use a smaller batch", n, lv.n_pieces);
        const long long want = (n_tasks + SF_DTW_WARPS - 1) / SF_DTW_WARPS;
        const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)c->sm_count * c->dtw_blocks_per_sm));
        const long long pwant = ((n_tasks + 1) / 2 + SF_DTW_WARPS - 1) / SF_DTW_WARPS;
        const int pgrid = c->R2 > 0 ? (int)std::max<long long>(1, std::min<long long>(pwant, (long long)c->sm_count * c->pair_blocks_per_sm)) : 0;
        cudaError_t e = cudaErrorInvalidValue;
        sf_dtw_args oa = da, pa = da;
        if (c->R2 > 0) {
            // Reads with exactly q events run two per warp (pair kernel), the others one per warp.  The two
            // persistent kernels share the GPU: the warp-per-read kernel is launched first and releases its
            // dependents as soon as its blocks are resident (griddepcontrol.launch_dependents); the pair kernel
            // is launched with programmatic stream serialisation, so it starts then instead of after the end.
            // Blocks of the first kernel that find the queue empty exit at once and leave their place to pair
            // blocks: a few ragged reads in a batch cost no extra wave.
            sf_partition_kernel<<<1, SF_PART_THREADS, 0, st>>>(s.d_info, n, c->opt.query_size, s.d_list_full, s.d_list_other, s.d_counts);
            SF_CUDA(c, cudaGetLastError());
            s.timing.other_launches++;
            oa.list = s.d_list_other;
            oa.n_list = s.d_counts + 1;
            oa.counter = s.d_counter + 1;
            SF_DISPATCH_R(c->R, std_dtw, (e = launch_dtw<R, STD>(oa, grid, smem, st)));
            SF_CUDA(c, e);
            s.timing.dtw_launches++;
            pa.list = s.d_list_full;
            pa.n_list = s.d_counts;
            sf_pair_op op = {SF_PAIR_LAUNCH, &pa, pgrid, psmem, st, 0, cudaErrorInvalidValue};
            sf_pair_run(c->R2, c->RQ2, std_dtw, op);
            SF_CUDA(c, op.err);
        } else {
            SF_DISPATCH_R(c->R, std_dtw, (e = launch_dtw<R, STD>(oa, grid, smem, st)));
            SF_CUDA(c, e);
        }
        s.timing.dtw_launches++;
        if (lv.n_split > 0) {
            // pieces of split segments: compare every warm front with the checkpoint its predecessor left at the
            // same boundary, then redo what differs (normally nothing: the redo kernels find no flagged item)
            sf_verify_args va;
            va.pieces = lv.d_pieces;
            va.groups = c->d_groups;
            va.warm_piece = lv.d_warm_piece;
            va.info = s.d_info;
            va.warm = s.d_warm;
            va.ckpt = s.d_ckpt;
            va.n_reads = n;
            va.n_warm = lv.n_warm;
            va.n_split = lv.n_split;
            va.ck_floats = c->ck_floats;
            va.ck_per_read = c->ck_per_read;
            va.n_f = sf_ckpt_floats(c->R);
            va.n_f_pair = c->R2 > 0 ? (c->R2 + 2) * SF_PAIR_LANES : va.n_f;
            va.first_bad = s.d_first_bad;
            va.n_bad = s.d_counts + 2;
            const long long items = (long long)n * lv.n_warm;
            const int vgrid = (int)std::max<long long>(1, std::min<long long>((items + 7) / 8, (long long)c->sm_count * 8));
            sf_verify_kernel<<<vgrid, 256, 0, st>>>(va);
            SF_CUDA(c, cudaGetLastError());
            s.timing.other_launches++;
            const long long fwant = ((long long)n * lv.n_split + SF_DTW_WARPS - 1) / SF_DTW_WARPS;
            oa.counter = s.d_counter + 2;
            pa.counter = s.d_counter + 3;
            e = cudaErrorInvalidValue;
            SF_DISPATCH_R(c->R, std_dtw, (e = launch_dtw_fix<R, STD>(
                oa, (int)std::max<long long>(1, std::min<long long>(fwant, (long long)c->sm_count * c->dtw_blocks_per_sm)), smem, st)));
            SF_CUDA(c, e);
            s.timing.dtw_launches++;
            if (c->R2 > 0) {
                sf_pair_op op = {SF_PAIR_LAUNCH_FIX, &pa,
                                 (int)std::max<long long>(1, std::min<long long>(fwant, (long long)c->sm_count * c->pair_blocks_per_sm)),
                                 psmem, st, 0, cudaErrorInvalidValue};
                sf_pair_run(c->R2, c->RQ2, std_dtw, op);
                SF_CUDA(c, op.err);
                s.timing.dtw_launches++;
            }
        }
    }
    SF_CUDA(c, cudaEventRecord(s.ev[3], st));
    if (n > 0) {
        sf_trace_args ta;
        ta.stream = c->d_stream;
        ta.segs = c->d_segs;
        ta.groups = c->d_groups;
        ta.seg_group = c->d_seg_group;
        ta.n_groups = c->n_groups;
        ta.pieces = c->levels[s.level].d_pieces;
        ta.n_pieces = c->levels[s.level].n_pieces;
        ta.n_reads = n;
        ta.queries = s.d_queries;
        ta.info = s.d_info;
        ta.q_cap = c->q_cap;
        ta.res = s.d_res;
        ta.ckpt = s.d_ckpt;
        ta.ck_per_read = c->ck_per_read;
        ta.ck_floats = c->ck_floats;
        ta.list_full = s.d_list_full;
        ta.n_full = s.d_counts;
        ta.hits = s.d_hits;
        ta.min_window = c->min_window;
        cudaError_t e = cudaErrorInvalidValue;
        SF_DISPATCH_R(c->R, std_dtw, (e = launch_trace<R, STD>(ta, st, c->R2)));
        SF_CUDA(c, e);
        s.timing.other_launches += c->R2 > 0 ? 2 : 1;
    }
    SF_CUDA(c, cudaEventRecord(s.ev[4], st));
    if (n > 0) {
        SF_CUDA(c, cudaMemcpyAsync(s.h_info, s.d_info, sizeof(sf_readinfo) * n, cudaMemcpyDeviceToHost, st));
        SF_CUDA(c, cudaMemcpyAsync(s.h_hits, s.d_hits, sizeof(sf_hit) * n, cudaMemcpyDeviceToHost, st));
        SF_CUDA(c, cudaMemcpyAsync(s.h_counts, s.d_counts, sizeof(int32_t) * 3, cudaMemcpyDeviceToHost, st));
        if (s.records && with_h2d)
            SF_CUDA(c, cudaMemcpyAsync(s.h_status, s.d_status, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    SF_CUDA(c, cudaEventRecord(s.ev[5], st));
    s.busy = true;
    s.done = false;
    s.timed = false;
    return SFGPU_OK;
}

int slot_wait(sfgpu_ctx *c, sf_slot &s)
{
    if (!s.busy)
        return SFGPU_OK;
    SF_CUDA(c, cudaStreamSynchronize(s.stream));
    s.busy = false;
    s.done = true;
    if (!s.timed) {
        float ms[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5; i++)
            SF_CUDA(c, cudaEventElapsedTime(&ms[i], s.ev[i], s.ev[i + 1]));
        s.timing.h2d_ms = ms[0];
        s.timing.events_ms = ms[1];
        s.timing.dtw_ms = ms[2];
        s.timing.trace_ms = ms[3];
        s.timing.d2h_ms = ms[4];
        s.timing.total_ms = ms[0] + ms[1] + ms[2] + ms[3] + ms[4];
        double cells = 0.0;
        for (int i = 0; i < s.n_reads; i++)
            cells += (double)s.h_info[i].qlen * (double)c->ref_columns;
        s.timing.cells = cells;
        s.timing.samples = s.raw_samples;
        s.timing.tasks_per_read = s.n_reads > 0 ? c->levels[s.level].n_pieces : 0;
        s.timing.piece_blocks = s.n_reads > 0 ? c->levels[s.level].len : 0;
        s.timing.redone_pieces = s.n_reads > 0 ? s.h_counts[2] : 0;
        s.timed = true;
    }
    return SFGPU_OK;
}

void free_ref(sfgpu_ctx *c)
{
    dfree(c->d_stream); dfree(c->d_segs); dfree(c->d_groups); dfree(c->d_order); dfree(c->d_seg_group);
    for (auto &lv : c->levels) {
        dfree(lv.d_pieces); dfree(lv.d_order); dfree(lv.d_warm_piece); dfree(lv.d_split_first); dfree(lv.d_split_count);
    }
    c->levels.clear();
    c->have_ref = false;
}

// Cuts the groups into tasks with pieces of `len` blocks (a multiple of every split group's checkpoint period, so
// that the front at a piece's end is one of the regular checkpoints).  len == 0: no splitting.
int build_level(sfgpu_ctx *c, int32_t len, sf_level &lv)
{
    lv = sf_level();
    lv.len = len;
    const bool std_dtw = (c->opt.flags & SFGPU_DTW) != 0;
    const double fill = 1.5; // blocks a task spends filling / draining the lane pipeline
    for (int g = 0; g < c->n_groups; g++) {
        const sf_group &grp = c->groups[g];
        const int64_t n_pos = grp.end - grp.begin;
        // blocks whose end still lies inside the group, in checkpoint periods
        const int64_t units = grp.ck_every > 0 ? (n_pos - 1) / ((int64_t)SF_BLOCK_COLS * grp.ck_every) : 0;
        int64_t n_pc = 1;
        if (len > 0 && !std_dtw && grp.nseg == 1 && grp.ck_every > 0 && len % grp.ck_every == 0)
            n_pc = std::max<int64_t>(1, std::min<int64_t>(units, (n_pos / SF_BLOCK_COLS + len / 2) / len));
        const int per = len > 0 && grp.ck_every > 0 ? len / grp.ck_every : 0;
        if (n_pc > 1) {
            lv.split_first.push_back((int32_t)lv.pieces.size());
            lv.split_count.push_back((int32_t)n_pc);
        }
        (void)per;
        for (int64_t k = 0; k < n_pc; k++) {
            sf_piece pc;
            pc.gid = g;
            pc.k = (int32_t)k;
            pc.pad = 0;
            pc.flags = n_pc > 1 ? ((k > 0 ? 1 : 0) | (k + 1 < n_pc ? 2 : 0)) : 0;
            // boundaries: the `units` checkpoint periods spread evenly over the pieces
            pc.b0 = n_pc > 1 ? (int32_t)(k * units / n_pc) * grp.ck_every : 0;
            pc.b1 = n_pc > 1 && k + 1 < n_pc ? (int32_t)((k + 1) * units / n_pc) * grp.ck_every : (int32_t)((n_pos + SF_BLOCK_COLS - 1) / SF_BLOCK_COLS);
            pc.sidx = n_pc > 1 ? (int32_t)lv.split_first.size() - 1 : -1;
            pc.widx = -1;
            if (pc.flags & 1) {
                pc.widx = lv.n_warm++;
                lv.warm_piece.push_back((int32_t)lv.pieces.size());
            }
            const double own = pc.b1 - pc.b0, blocks = own + fill + ((pc.flags & 1) ? std::min(c->warm_blocks, pc.b0) : 0);
            lv.blocks_per_read += blocks;
            lv.longest = std::max(lv.longest, blocks);
            lv.pieces.push_back(pc);
        }
    }
    lv.n_pieces = (int32_t)lv.pieces.size();
    lv.n_split = (int32_t)lv.split_first.size();
    lv.order.resize(lv.n_pieces);
    for (int i = 0; i < lv.n_pieces; i++)
        lv.order[i] = i;
    std::stable_sort(lv.order.begin(), lv.order.end(), [&](int a, int b) {
        return (lv.pieces[a].b1 - lv.pieces[a].b0) > (lv.pieces[b].b1 - lv.pieces[b].b0);
    });
    auto up = [&](auto *&dst, const auto &v) -> int {
        using T = std::remove_reference_t<decltype(*dst)>;
        SF_CUDA(c, cudaMalloc(&dst, sizeof(T) * std::max<size_t>(1, v.size())));
        if (!v.empty())
            SF_CUDA(c, cudaMemcpy(dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
        return SFGPU_OK;
    };
    int rc;
    if ((rc = up(lv.d_pieces, lv.pieces)) || (rc = up(lv.d_order, lv.order)) || (rc = up(lv.d_warm_piece, lv.warm_piece)) ||
        (rc = up(lv.d_split_first, lv.split_first)) || (rc = up(lv.d_split_count, lv.split_count)))
        return rc;
    return SFGPU_OK;
}

// waits for all slots, releases their buffers (group count / checkpoint layout may change) and the
// resident reference
int drop_ref(sfgpu_ctx *c)
{
    for (auto &s : c->slots) {
        int rc = slot_wait(c, s);
        if (rc)
            return rc;
        slot_free_buffers(s);
        s.done = false;
    }
    free_ref(c);
    return SFGPU_OK;
}

// Lays the segments out in the reference's processing order (contig ascending, '+' then '-':
// sigfish.c:870,890,936), one +INF sentinel column in front of each, packs them into groups and
// allocates the stream (filled with +INF).
int layout_ref(sfgpu_ctx *c, int32_t num_ref, const int32_t *rlens, bool has_reverse)
{
    c->segs.clear();
    c->ref_lengths.assign(rlens, rlens + num_ref);
    int64_t pos = 0; // stream cursor
    for (int32_t r = 0; r < num_ref; r++) {
        for (int strand = 0; strand < (has_reverse ? 2 : 1); strand++) {
            sf_seg sg;
            sg.off = pos + 1; // sentinel at pos
            sg.rlen = rlens[r];
            sg.rid = r;
            sg.strand = strand;
            sg.pad = 0;
            c->segs.push_back(sg);
            pos += (int64_t)rlens[r] + 1;
        }
    }
    c->num_ref = num_ref;
    c->n_seg = (int32_t)c->segs.size();
    c->stream_len = pos + 64;
    c->ref_columns = 0;
    for (const auto &sg : c->segs)
        c->ref_columns += sg.rlen;

    // ---- groups: consecutive segments up to a column budget ----
    const int64_t budget = std::max<int64_t>(4096, std::min<int64_t>(65536, c->ref_columns / 64));
    // checkpoint period: the long segments together get ~512 checkpoints per read (0.66 MB), whatever the
    // genome size; never closer than ck_min_cols/4 (512) columns
    int64_t long_cols = 0;
    for (int s = 0; s < c->n_seg; s++)
        if (c->segs[s].rlen > c->ck_min_cols)
            long_cols += c->segs[s].rlen + 1;
    // SFGPU_CHECKPOINTS_PER_READ in the environment (16 .. 4096, default 512) trades memory against the length of the
    // start-coordinate pass: 128 = 0.17 MB per read for 1.7 % of the throughput on the 1 Mb shape, 64 = 0.08 MB for 3.1 %
    // (profiles/r02_checkpoint_density_knob.txt)
    int64_t ck_target = 512;
    if (const char *e = getenv("SFGPU_CHECKPOINTS_PER_READ")) {
        const long v = strtol(e, nullptr, 10);
        if (v >= 16 && v <= 4096)
            ck_target = v;
    }
    const int64_t cols_per = std::max<int64_t>(c->ck_min_cols / 4, long_cols / ck_target);
    c->groups.clear();
    c->seg_group.assign(c->n_seg, 0);
    int s0 = 0;
    while (s0 < c->n_seg) {
        int s1 = s0;
        int64_t cols = 0;
        int32_t longest = 0;
        while (s1 < c->n_seg && (s1 == s0 || cols + c->segs[s1].rlen + 1 <= budget)) {
            cols += c->segs[s1].rlen + 1;
            longest = std::max(longest, c->segs[s1].rlen);
            s1++;
        }
        sf_group g;
        g.begin = c->segs[s0].off - 1;
        g.end = c->segs[s1 - 1].off + c->segs[s1 - 1].rlen;
        g.seg0 = s0;
        g.nseg = s1 - s0;
        g.ck_every = 0;
        g.n_ck = 0;
        g.ck_prefix = 0;
        if (g.end - g.begin > 0x7ffffff0ll)
            return fail(c, SFGPU_ELIMIT, "segment group too long");
        if (longest > c->ck_min_cols) {
            g.ck_every = (int32_t)((cols_per + SF_BLOCK_COLS - 1) / SF_BLOCK_COLS); // in blocks of 64 columns
            g.n_ck = (int32_t)(((g.end - g.begin) / SF_BLOCK_COLS) / g.ck_every);
        }
        for (int s = s0; s < s1; s++)
            c->seg_group[s] = (int32_t)c->groups.size();
        c->groups.push_back(g);
        s0 = s1;
    }
    c->n_groups = (int32_t)c->groups.size();
    c->ck_per_read = 0;
    for (auto &g : c->groups) {
        g.ck_prefix = c->ck_per_read;
        c->ck_per_read += g.n_ck;
    }
    c->order.resize(c->n_groups);
    for (int i = 0; i < c->n_groups; i++)
        c->order[i] = i;
    std::stable_sort(c->order.begin(), c->order.end(), [&](int a, int b) {
        return (c->groups[a].end - c->groups[a].begin) > (c->groups[b].end - c->groups[b].begin);
    });

    SF_CUDA(c, cudaMalloc(&c->d_stream, sizeof(float) * c->stream_len));
    SF_CUDA(c, cudaMalloc(&c->d_segs, sizeof(sf_seg) * c->n_seg));
    SF_CUDA(c, cudaMemcpy(c->d_segs, c->segs.data(), sizeof(sf_seg) * c->n_seg, cudaMemcpyHostToDevice));
    SF_CUDA(c, cudaMalloc(&c->d_groups, sizeof(sf_group) * c->n_groups));
    SF_CUDA(c, cudaMemcpy(c->d_groups, c->groups.data(), sizeof(sf_group) * c->n_groups, cudaMemcpyHostToDevice));
    SF_CUDA(c, cudaMalloc(&c->d_order, sizeof(int32_t) * c->n_groups));
    SF_CUDA(c, cudaMemcpy(c->d_order, c->order.data(), sizeof(int32_t) * c->n_groups, cudaMemcpyHostToDevice));
    SF_CUDA(c, cudaMalloc(&c->d_seg_group, sizeof(int32_t) * c->n_seg));
    SF_CUDA(c, cudaMemcpy(c->d_seg_group, c->seg_group.data(), sizeof(int32_t) * c->n_seg, cudaMemcpyHostToDevice));
    sf_fill_inf_kernel<<<c->sm_count * 4, 256>>>(c->d_stream, (size_t)c->stream_len);
    SF_CUDA(c, cudaGetLastError());

    // ---- task levels: level 0 = one task per group; further levels cut the long single-segment groups into
    //      pieces of len = base * 2^j blocks, base = the smallest multiple of the checkpoint period that is at
    //      least four warm-ups (and two chunks) long ----
    c->levels.clear();
    c->levels.emplace_back();
    int rc0 = build_level(c, 0, c->levels.back());
    if (rc0)
        return rc0;
    int32_t ck_e = 0;
    int64_t longest_blocks = 0;
    for (const auto &g : c->groups)
        if (g.nseg == 1 && g.ck_every > 0) {
            if (ck_e == 0) ck_e = g.ck_every; // every group with checkpoints has the same period
            longest_blocks = std::max<int64_t>(longest_blocks, (g.end - g.begin) / SF_BLOCK_COLS);
        }
    if (ck_e > 0 && !(c->opt.flags & SFGPU_DTW) && c->force_level_len >= 0) {
        // a piece must hold more than two chunks, so that the chunks cut by its two ends are different ones
        const int32_t two_chunks = (2 * c->q_cap + SF_BLOCK_COLS - 1) / SF_BLOCK_COLS + 1;
        const int32_t min_len = std::max(4 * c->warm_blocks, two_chunks);
        int32_t base = ((min_len + ck_e - 1) / ck_e) * ck_e;
        if (c->force_level_len > 0)
            base = std::max(c->force_level_len, (two_chunks + ck_e - 1) / ck_e) * ck_e;
        for (int64_t len = base; 2 * len <= longest_blocks + len / 2 && c->levels.size() < 24; len *= 2) {
            c->levels.emplace_back();
            rc0 = build_level(c, (int32_t)len, c->levels.back());
            if (rc0)
                return rc0;
            if (c->levels.back().n_split == 0) {
                sf_level &lv = c->levels.back();
                dfree(lv.d_pieces); dfree(lv.d_order); dfree(lv.d_warm_piece); dfree(lv.d_split_first); dfree(lv.d_split_count);
                c->levels.pop_back();
                break;
            }
            if (c->force_level_len > 0)
                break;
        }
    }
    SF_CUDA(c, cudaDeviceSynchronize());
    return SFGPU_OK;
}

// the level with the smallest expected DTW time for a batch of n reads: work / task slots + half the longest task
// (the expected idle tail of a persistent grid whose last tasks have that length)
int pick_level(const sfgpu_ctx *c, int n)
{
    if (c->force_level_len > 0 && c->levels.size() > 1)
        return 1;
    const double slots = c->R2 > 0 ? (double)c->sm_count * c->pair_blocks_per_sm * SF_DTW_WARPS * 2
                                   : (double)c->sm_count * c->dtw_blocks_per_sm * SF_DTW_WARPS;
    int best = 0;
    double best_t = 0.0;
    for (size_t i = 0; i < c->levels.size(); i++) {
        const sf_level &lv = c->levels[i];
        const double t = std::max(lv.longest, (double)n * lv.blocks_per_read / slots + 0.5 * lv.longest);
        if (i == 0 || t < best_t) {
            best = (int)i;
            best_t = t;
        }
    }
    return best;
}

// per-batch buffers whose size depends on the split level
int slot_reserve_level(sfgpu_ctx *c, sf_slot &s, int n, const sf_level &lv)
{
    const size_t need_res = (size_t)n * lv.n_pieces;
    if (need_res > s.cap_res) {
        dfree(s.d_res);
        s.cap_res = 0;
        const size_t cap = need_res + need_res / 4;
        SF_CUDA(c, cudaMalloc(&s.d_res, sizeof(sf_taskres) * cap));
        s.cap_res = cap;
    }
    const size_t need_warm = (size_t)n * lv.n_warm * (size_t)c->ck_floats;
    if (need_warm > s.cap_warm) {
        dfree(s.d_warm);
        s.cap_warm = 0;
        const size_t cap = need_warm + need_warm / 4;
        SF_CUDA(c, cudaMalloc(&s.d_warm, sizeof(float) * cap));
        s.cap_warm = cap;
    }
    const size_t need_bad = (size_t)n * lv.n_split;
    if (need_bad > s.cap_bad) {
        dfree(s.d_first_bad);
        s.cap_bad = 0;
        const size_t cap = need_bad + need_bad / 4;
        SF_CUDA(c, cudaMalloc(&s.d_first_bad, sizeof(int32_t) * cap));
        s.cap_bad = cap;
    }
    return SFGPU_OK;
}

// common part of sfgpu_submit / sfgpu_submit_reads: read i is `len(i)` samples at `ptr(i)`
template <typename PtrFn, typename LenFn>
int submit_common(sfgpu_ctx *c, int32_t slot, int32_t n_reads, PtrFn ptr, LenFn len_of,
                         const float *digitisation, const float *offset, const float *range)
{
    if (!c->have_ref)
        return fail(c, SFGPU_ESTATE, "sfgpu_submit before sfgpu_set_ref");
    if (slot < 0 || slot >= (int)c->slots.size() || n_reads < 0)
        return fail(c, SFGPU_EARG, "bad slot or read count");
    if (n_reads > 0 && (!digitisation || !offset || !range))
        return fail(c, SFGPU_EARG, "null batch array");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    g_trace(c->opt.device, "submit: begin");
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    g_trace(c->opt.device, "submit: slot free");
    // every read starts on a 16-byte boundary so that the event kernel can use 16-byte loads
    int64_t padded = 0, raw = 0;
    for (int i = 0; i < n_reads; i++) {
        const int64_t len = len_of(i);
        if (len < 0 || (len > 0 && !ptr(i)))
            return fail(c, SFGPU_EARG, "bad signal pointer / length for read %d", i);
        padded += (len + 7) & ~7ll;
        raw += len;
    }
    rc = slot_reserve(c, s, std::max(n_reads, 1), padded + 8);
    if (rc)
        return rc;
    int64_t cur = 0;
    int64_t *h_len = s.h_off + (n_reads + 1);
    for (int i = 0; i < n_reads; i++) {
        const int64_t len = len_of(i);
        s.h_off[i] = cur;
        h_len[i] = len;
        cur += (len + 7) & ~7ll;
        s.h_scal[i] = digitisation[i];
        s.h_scal[s.cap_reads + i] = offset[i];
        s.h_scal[2 * (size_t)s.cap_reads + i] = range[i];
    }
    s.h_off[n_reads] = cur;
    // gather of the samples into the pinned staging buffer: the only per-sample work of the host on this path.  One
    // core copies ~8 GB/s, which is less than a batch of short-reference work needs (65 k reads, 590 MB, against
    // 140 ms of device time), so large batches are copied by a few threads.
    auto copy_range = [&](int b, int e) {
        for (int i = b; i < e; i++) {
            const int64_t len = h_len[i], at = s.h_off[i];
            if (len > 0)
                memcpy(s.h_signal + at, ptr(i), sizeof(int16_t) * (size_t)len);
            for (int64_t j = at + len; j < s.h_off[i + 1]; j++)
                s.h_signal[j] = 0;
        }
    };
    const int n_thr = (int)std::min<int64_t>(std::min<int64_t>(8, std::max(1u, std::thread::hardware_concurrency() / 2)),
                                             cur / (8 << 20));
    if (n_thr <= 1) {
        copy_range(0, n_reads);
    } else {
        std::vector<std::thread> th;
        int b = 0;
        for (int t = 0; t < n_thr; t++) { // ranges of about equal sample counts
            const int64_t upto = cur * (t + 1) / n_thr;
            int e = b;
            while (e < n_reads && s.h_off[e + 1] <= upto)
                e++;
            if (t + 1 == n_thr)
                e = n_reads;
            th.emplace_back(copy_range, b, e);
            b = e;
        }
        for (auto &t : th)
            t.join();
    }
    s.n_reads = n_reads;
    s.n_samples = cur;
    s.raw_samples = raw;
    s.queries_only = false;
    s.records = false;
    g_trace(c->opt.device, "submit: samples staged in pinned memory");
    rc = run_stages(c, s, true);
    g_trace(c->opt.device, "submit: stages enqueued");
    return rc;
}


} // namespace

extern "C" {

int sfgpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    int ok = 0;
    for (int i = 0; i < n; i++) { // attributes, not cudaGetDeviceProperties: the latter costs ~0.1 s per device
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10)
            ok++;
    }
    return ok;
}

const char *sfgpu_strerror(const sfgpu_ctx *ctx) { return ctx ? ctx->err : g_err; }

int sfgpu_create(sfgpu_ctx **out, const sfgpu_opt_t *opt, const float *level_mean)
{
    if (!out || !opt || !level_mean)
        return fail(nullptr, SFGPU_EARG, "sfgpu_create: null argument");
    *out = nullptr;
    if (opt->query_size <= 0)
        return fail(nullptr, SFGPU_EARG, "query_size must be positive");
    if (opt->prefix_size < 0 && (!(opt->flags & SFGPU_RNA) || (opt->flags & (SFGPU_INV | SFGPU_END))))
        return fail(nullptr, SFGPU_EARG, "prefix_size < 0 (automatic query start) needs --rna and excludes --invert / --from-end");
    if (opt->kmer_size < 1 || opt->kmer_size > 12)
        return fail(nullptr, SFGPU_EARG, "kmer_size out of range");
    const int rows = pick_rows(opt->query_size);
    if (rows == 0)
        return fail(nullptr, SFGPU_ELIMIT, "query_size %d exceeds the supported maximum of 1024 events", opt->query_size);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(nullptr, SFGPU_ENODEV, "no CUDA device (this library has no CPU fallback)");
    if (opt->device < 0 || opt->device >= ndev)
        return fail(nullptr, SFGPU_EARG, "device %d out of range (%d present)", opt->device, ndev);
    int cc_major = 0, n_sm = 0;
    if (cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, opt->device) != cudaSuccess || cc_major != 10 ||
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, opt->device) != cudaSuccess || n_sm <= 0)
        return fail(nullptr, SFGPU_ENODEV, "device %d is not an sm_100 GPU (kernels are built for sm_100a only)", opt->device);

    sfgpu_ctx *c = new (std::nothrow) sfgpu_ctx();
    if (!c)
        return fail(nullptr, SFGPU_EARG, "out of host memory");
    c->err[0] = 0;
    c->opt = *opt;
    c->R = rows;
    c->q_cap = 32 * rows;
    if (opt->prefix_size >= 0) {
        c->ev_cap = opt->prefix_size + opt->query_size + 4;
        c->ev_cap_a = 0;
    } else { // events 0..50+q for the fall-back window, then q more from the detected start
        c->ev_cap_a = 50 + opt->query_size + 4;
        c->ev_cap = c->ev_cap_a + opt->query_size + 4;
    }
    c->sm_count = n_sm;
    // restart distance of the start-coordinate pass: a path that leaves the window through the restart front is
    // followed from an earlier checkpoint (exact either way), so this is a cost knob.  A warping path of q events
    // spans about q / 1.5 reference columns; with 1.25 q instead of round 1's 2 q the pass is 18 % shorter on the 30 kb
    // shape (1.17 -> 0.95 ms per 16 384 reads; q: 0.87 ms) and no synthetic read left its window.
    c->min_window = opt->query_size + opt->query_size / 4;
    if (opt->reserved[0] > 0)
        c->ck_min_cols = std::max(128, opt->reserved[0]);
    if (opt->reserved[1] > 0)
        c->min_window = opt->reserved[1];
    // warm-up of a piece of a split segment: 2q columns, rounded up to whole blocks
    c->warm_blocks = (2 * opt->query_size + SF_BLOCK_COLS) / SF_BLOCK_COLS;
    if (opt->reserved[2] > 0)
        c->warm_blocks = opt->reserved[2];
    c->force_level_len = opt->reserved[4];
    if (c->opt.n_slots <= 0)
        c->opt.n_slots = 2;
    g_trace(opt->device, "sfgpu_create: begin");
    int rc = [&]() -> int {
        SF_CUDA(c, cudaSetDevice(opt->device));
        SF_CUDA(c, cudaFree(0));
        g_trace(opt->device, "sfgpu_create: CUDA context ready");
        const size_t nm = (size_t)1 << (2 * opt->kmer_size);
        SF_CUDA(c, cudaMalloc(&c->d_level_mean, sizeof(float) * nm));
        SF_CUDA(c, cudaMemcpy(c->d_level_mean, level_mean, sizeof(float) * nm, cudaMemcpyHostToDevice));
        c->slots.resize(c->opt.n_slots);
        for (auto &s : c->slots) {
            SF_CUDA(c, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            for (auto &e : s.ev)
                SF_CUDA(c, cudaEventCreate(&e));
            SF_CUDA(c, cudaMalloc(&s.d_counter, sizeof(unsigned int) * 4));
            SF_CUDA(c, cudaMalloc(&s.d_counts, sizeof(int32_t) * 4));
            SF_CUDA(c, cudaMallocHost(&s.h_counts, sizeof(int32_t) * 4));
            memset(s.h_counts, 0, sizeof(int32_t) * 4);
            memset(&s.timing, 0, sizeof s.timing);
        }
        const bool std_dtw = (opt->flags & SFGPU_DTW) != 0;
        const size_t smem = sizeof(float) * SF_DTW_WARPS * sf_smem_floats_per_warp(rows, std_dtw);
        int nb = 0;
        SF_DISPATCH_R(rows, std_dtw, (nb = dtw_occupancy<R, STD>(smem)));
        if (nb <= 0)
            return fail(c, SFGPU_ECUDA, "DTW kernel does not fit on the device (R=%d)", rows);
        c->dtw_blocks_per_sm = nb;
        {
            int ib = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ib, sf_inflate_kernel, 32 * SF_INF_WARPS,
                                                              SF_INF_WARPS * SF_INF_SMEM_WARP) == cudaSuccess && ib > 0)
                c->inflate_blocks_per_sm = ib;
        }
        // the event kernel keeps the prefix sums and statistics of its reads in > 48 KB of dynamic shared memory
        SF_CUDA(c, cudaFuncSetAttribute(sf_events_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf_events_smem_bytes()));
        c->ck_floats = sf_ckpt_floats(rows);
        // Two full-length reads per warp (16 lanes x R2 rows each, R2 = ceil(q / 16)) for every 64 < q <= 256: the
        // per-macro-step overheads are shared by twice the cells, and at most 15 rows per read are padding.  Round 2
        // had only R2 = 16 (useful rate falling as q/256: q = 256 8.68, 250 8.37, 200 6.62, 130 4.22 TCUPS on the 1 Mb
        // shape) and used it for 192 < q <= 256; --dtw-std still has only that instantiation.  reserved[3]: 1 =
        // pairing off, 2 = pairing for every 128 < q <= 256 also with --dtw-std (tests).
        {
            const int q = opt->query_size;
            int r2 = 0;
            if (q > 64 && q <= 256 && opt->reserved[3] != 1) {
                if (!std_dtw)
                    r2 = (q + SF_PAIR_LANES - 1) / SF_PAIR_LANES;
                else if (q > 192 || (q > 128 && opt->reserved[3] == 2))
                    r2 = 16;
            }
            if (r2 > 0 && sf_pair_exists(r2, std_dtw)) {
                c->R2 = r2;
                c->RQ2 = (q - 1) % r2;
                c->ck_floats = std::max(c->ck_floats, (c->R2 + 2) * SF_PAIR_LANES);
                sf_pair_op op = {SF_PAIR_OCCUPANCY, nullptr, 0, sizeof(float) * SF_DTW_WARPS * sf_pair_smem_floats_per_warp(std_dtw),
                                 nullptr, 0, cudaErrorInvalidValue};
                if (!sf_pair_run(c->R2, c->RQ2, std_dtw, op) || op.err != cudaSuccess || op.blocks_per_sm <= 0)
                    return fail(c, SFGPU_ECUDA, "pair DTW kernel does not fit on the device (R2=%d, RQ=%d)", c->R2, c->RQ2);
                c->pair_blocks_per_sm = op.blocks_per_sm;
            }
        }
        return SFGPU_OK;
    }();
    if (rc != SFGPU_OK) {
        char keep[512];
        memcpy(keep, c->err, sizeof keep);
        sfgpu_destroy(c);
        memcpy(g_err, keep, sizeof keep);
        return rc;
    }
    g_trace(opt->device, "sfgpu_create: end (streams, model, occupancy queries)");
    *out = c;
    return SFGPU_OK;
}

void sfgpu_destroy(sfgpu_ctx *c)
{
    if (!c)
        return;
    g_trace(c->opt.device, "sfgpu_destroy: begin");
    cudaSetDevice(c->opt.device);
    for (auto &s : c->slots) {
        if (s.stream)
            cudaStreamSynchronize(s.stream);
        slot_free_buffers(s);
        dfree(s.d_counter);
        dfree(s.d_counts);
        hfree(s.h_counts);
        for (auto &e : s.ev)
            if (e)
                cudaEventDestroy(e);
        if (s.stream)
            cudaStreamDestroy(s.stream);
    }
    free_ref(c);
    dfree(c->d_level_mean);
    g_trace(c->opt.device, "sfgpu_destroy: end");
    delete c;
}

int sfgpu_set_ref(sfgpu_ctx *c, int32_t num_ref, const char *bases, const int64_t *base_off,
                  int32_t *ref_lengths, int32_t *ref_seq_lengths, int32_t *ref_st_offset)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (num_ref <= 0 || !bases || !base_off)
        return fail(c, SFGPU_EARG, "sfgpu_set_ref: empty reference");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    g_trace(c->opt.device, "sfgpu_set_ref: begin");
    int rc = drop_ref(c);
    if (rc)
        return rc;

    const uint32_t flags = c->opt.flags;
    const bool rna = (flags & SFGPU_RNA) != 0;
    const int k = c->opt.kmer_size;
    const int q = c->opt.query_size;
    std::vector<sf_refseg> rsegs;
    std::vector<int32_t> rlens(num_ref);
    for (int32_t r = 0; r < num_ref; r++) {
        const int64_t l64 = base_off[r + 1] - base_off[r];
        if (l64 < k || l64 > 0x7fffff00ll)
            return fail(c, SFGPU_EARG, "contig %d has %lld bases; need between k=%d and 2^31", r, (long long)l64, k);
        const int32_t l = (int32_t)l64;
        const int32_t nk = l + 1 - k;
        // genref.c:128-136: DNA and --full-ref use every k-mer, RNA the 1.5*q nearest one end
        int32_t rl = nk;
        if (rna && !(flags & SFGPU_REF)) {
            const uint32_t cap = (uint32_t)(q * 1.5);
            rl = cap > (uint32_t)nk ? nk : (int32_t)cap;
        }
        int32_t st_off = 0;
        sf_refseg f;
        f.base_off = base_off[r];
        f.seq_len = l;
        f.first = 0;
        f.mode = 0;
        f.pad = 0;
        if (rna) {
            if (flags & SFGPU_INV) {          // genref.c:166-177 (offset left at 0)
                f.first = l - rl - (k - 1);
                f.mode = 2;
            } else if (!(flags & SFGPU_END)) { // genref.c:185-197
                st_off = l - rl - (k - 1);
                f.first = st_off;
            }
        }
        rlens[r] = rl;
        if (ref_lengths) ref_lengths[r] = rl;
        if (ref_seq_lengths) ref_seq_lengths[r] = l;
        if (ref_st_offset) ref_st_offset[r] = st_off;
        rsegs.push_back(f);
        if (!rna) {
            f.mode = 1;
            rsegs.push_back(f);
        }
    }
    rc = layout_ref(c, num_ref, rlens.data(), !rna);
    if (rc)
        return rc;
    g_trace(c->opt.device, "sfgpu_set_ref: layout done");

    const int64_t n_bases = base_off[num_ref];
    uint8_t *d_bases = nullptr;
    sf_refseg *d_rseg = nullptr;
    float2 *d_stats = nullptr;
    rc = [&]() -> int {
        SF_CUDA(c, cudaMalloc(&d_bases, (size_t)n_bases + 16));
        SF_CUDA(c, cudaMemcpy(d_bases, bases, (size_t)n_bases, cudaMemcpyHostToDevice));
        SF_CUDA(c, cudaMalloc(&d_rseg, sizeof(sf_refseg) * c->n_seg));
        SF_CUDA(c, cudaMemcpy(d_rseg, rsegs.data(), sizeof(sf_refseg) * c->n_seg, cudaMemcpyHostToDevice));
        SF_CUDA(c, cudaMalloc(&d_stats, sizeof(float2) * c->n_seg));
        sf_ref_args ra;
        ra.bases = d_bases;
        ra.rseg = d_rseg;
        ra.segs = c->d_segs;
        ra.n_seg = c->n_seg;
        ra.level_mean = c->d_level_mean;
        ra.k = k;
        ra.stream = c->d_stream;
        ra.stats = d_stats;
        int32_t longest = 0;
        for (const auto &sg : c->segs)
            longest = std::max(longest, sg.rlen);
        dim3 grid((unsigned)std::max(1, std::min(c->sm_count * 2, (longest + 255) / 256)),
                  (unsigned)std::min(c->n_seg, 32768));
        sf_ref_fill_kernel<<<grid, 256>>>(ra);
        SF_CUDA(c, cudaGetLastError());
        const int stat_blocks = std::max(1, std::min((c->n_seg + 3) / 4, c->sm_count * 8));
        sf_ref_stats_kernel<<<stat_blocks, 128>>>(ra);
        SF_CUDA(c, cudaGetLastError());
        sf_ref_scale_kernel<<<grid, 256>>>(ra);
        SF_CUDA(c, cudaGetLastError());
        SF_CUDA(c, cudaDeviceSynchronize());
        return SFGPU_OK;
    }();
    dfree(d_bases);
    dfree(d_rseg);
    dfree(d_stats);
    if (rc != SFGPU_OK) {
        free_ref(c);
        return rc;
    }
    c->have_ref = true;
    g_trace(c->opt.device, "sfgpu_set_ref: end (fill, stats, scale)");
    return SFGPU_OK;
}

int sfgpu_set_ref_events(sfgpu_ctx *c, int32_t num_ref, int32_t has_reverse, const float *events,
                         const int64_t *ev_off)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (num_ref <= 0 || !events || !ev_off)
        return fail(c, SFGPU_EARG, "sfgpu_set_ref_events: empty reference");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    int rc = drop_ref(c);
    if (rc)
        return rc;
    const int per = has_reverse ? 2 : 1;
    std::vector<int32_t> rlens(num_ref);
    for (int32_t r = 0; r < num_ref; r++) {
        const int64_t l = ev_off[per * r + 1] - ev_off[per * r];
        if (l <= 0 || l > 0x7fffff00ll || (has_reverse && ev_off[2 * r + 2] - ev_off[2 * r + 1] != l))
            return fail(c, SFGPU_EARG, "bad event array length for contig %d", r);
        rlens[r] = (int32_t)l;
    }
    rc = layout_ref(c, num_ref, rlens.data(), has_reverse != 0);
    if (rc)
        return rc;
    for (int32_t s = 0; s < c->n_seg; s++) {
        const sf_seg &sg = c->segs[s];
        cudaError_t e = cudaMemcpy(c->d_stream + sg.off, events + ev_off[s], sizeof(float) * sg.rlen, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            free_ref(c);
            return fail(c, SFGPU_ECUDA, "cudaMemcpy of reference events: %s", cudaGetErrorString(e));
        }
    }
    c->have_ref = true;
    return SFGPU_OK;
}

int sfgpu_submit(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const int16_t *signals,
                 const int64_t *sig_off, const float *digitisation, const float *offset,
                 const float *range)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (n_reads > 0 && (!signals || !sig_off))
        return fail(c, SFGPU_EARG, "null batch array");
    return submit_common(c, slot, n_reads, [&](int i) { return signals + sig_off[i]; },
                         [&](int i) { return sig_off[i + 1] - sig_off[i]; }, digitisation, offset, range);
}

int sfgpu_submit_reads(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const int16_t *const *signals,
                       const int64_t *n_samples, const float *digitisation, const float *offset,
                       const float *range)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (n_reads > 0 && (!signals || !n_samples))
        return fail(c, SFGPU_EARG, "null batch array");
    return submit_common(c, slot, n_reads, [&](int i) { return signals[i]; }, [&](int i) { return n_samples[i]; },
                         digitisation, offset, range);
}

int sfgpu_submit_records(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const uint8_t *const *records, const int64_t *record_bytes,
                         int32_t record_press, int32_t signal_press, const int32_t *sig_pos, const int64_t *sig_bytes,
                         const int64_t *n_samples, const float *digitisation, const float *offset, const float *range)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (!c->have_ref)
        return fail(c, SFGPU_ESTATE, "sfgpu_submit_records before sfgpu_set_ref");
    if (slot < 0 || slot >= (int)c->slots.size() || n_reads < 0)
        return fail(c, SFGPU_EARG, "bad slot or read count");
    if (n_reads > 0 && (!records || !record_bytes || !sig_pos || !sig_bytes || !n_samples || !digitisation || !offset || !range))
        return fail(c, SFGPU_EARG, "null batch array");
    if (record_press < 0 || record_press > 1 || signal_press < 0 || signal_press > 1)
        return fail(c, SFGPU_ELIMIT, "record / signal compression %d / %d is not decoded on the device (zlib and svb-zd are)",
                    record_press, signal_press);
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    g_trace(c->opt.device, "submit_records: begin");
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    // layout: records on 8-byte boundaries, inflated records with 256 bytes of slack (auxiliary fields behind the
    // signal are inflated too so that the checksum can be verified), samples padded to multiples of 8
    int64_t rec_total = 0, scr_total = 0, padded = 0, raw = 0;
    for (int i = 0; i < n_reads; i++) {
        const int64_t nb = record_bytes[i], ns = n_samples[i];
        if (nb < 0 || ns < 0 || ns > 0x7fffff00ll || sig_pos[i] < 0 || sig_bytes[i] < 0 || (nb > 0 && !records[i]))
            return fail(c, SFGPU_EARG, "bad record %d", i);
        if (nb > 0x0fffffffll || (int64_t)sig_pos[i] + sig_bytes[i] > 0x7ffd0000ll) // the device decoder counts in 32 bits
            return fail(c, SFGPU_ELIMIT, "record %d is too large for the device decoder (%lld bytes compressed)", i, (long long)nb);
        if (!record_press && (int64_t)sig_pos[i] + sig_bytes[i] > nb)
            return fail(c, SFGPU_EARG, "record %d: signal field runs past the record", i);
        rec_total += (nb + 7) & ~7ll;
        if (record_press)
            scr_total += ((int64_t)sig_pos[i] + sig_bytes[i] + 256 + 15) & ~15ll;
        padded += (ns + 7) & ~7ll;
        raw += ns;
    }
    rc = slot_reserve(c, s, std::max(n_reads, 1), padded + 8, false);
    if (rc)
        return rc;
    if ((size_t)rec_total + 16 > s.cap_rec) {
        hfree(s.h_rec); dfree(s.d_rec);
        s.cap_rec = 0;
        const size_t cap = (size_t)rec_total + (size_t)rec_total / 4 + (1 << 20);
        SF_CUDA(c, cudaMallocHost(&s.h_rec, cap));
        SF_CUDA(c, cudaMalloc(&s.d_rec, cap));
        s.cap_rec = cap;
    }
    if ((size_t)scr_total + 16 > s.cap_scratch) {
        dfree(s.d_scratch);
        s.cap_scratch = 0;
        const size_t cap = (size_t)scr_total + (size_t)scr_total / 4 + (1 << 20);
        SF_CUDA(c, cudaMalloc(&s.d_scratch, cap));
        s.cap_scratch = cap;
    }
    if ((size_t)s.cap_reads > s.cap_rmeta) {
        hfree(s.h_rmeta); dfree(s.d_rmeta); hfree(s.h_status); dfree(s.d_status);
        s.cap_rmeta = 0;
        const size_t n = (size_t)s.cap_reads;
        SF_CUDA(c, cudaMallocHost(&s.h_rmeta, sizeof(int64_t) * (5 * n + 2)));
        SF_CUDA(c, cudaMalloc(&s.d_rmeta, sizeof(int64_t) * (5 * n + 2)));
        SF_CUDA(c, cudaMallocHost(&s.h_status, sizeof(int32_t) * n));
        SF_CUDA(c, cudaMalloc(&s.d_status, sizeof(int32_t) * n));
        s.cap_rmeta = n;
    }
    const size_t n = (size_t)n_reads;
    int64_t *rec_off = s.h_rmeta, *rec_len = s.h_rmeta + (n + 1), *scr_off = s.h_rmeta + (2 * n + 1),
            *spos = s.h_rmeta + (3 * n + 2), *sbytes = s.h_rmeta + (4 * n + 2);
    int64_t *h_len = s.h_off + (n_reads + 1);
    int64_t rcur = 0, scur = 0, cur = 0;
    for (int i = 0; i < n_reads; i++) {
        rec_off[i] = rcur;
        rec_len[i] = record_bytes[i];
        scr_off[i] = scur;
        spos[i] = sig_pos[i];
        sbytes[i] = sig_bytes[i];
        s.h_off[i] = cur;
        h_len[i] = n_samples[i];
        rcur += (record_bytes[i] + 7) & ~7ll;
        if (record_press)
            scur += ((int64_t)sig_pos[i] + sig_bytes[i] + 256 + 15) & ~15ll;
        cur += (n_samples[i] + 7) & ~7ll;
        s.h_scal[i] = digitisation[i];
        s.h_scal[s.cap_reads + i] = offset[i];
        s.h_scal[2 * (size_t)s.cap_reads + i] = range[i];
    }
    rec_off[n_reads] = rcur;
    scr_off[n_reads] = scur;
    s.h_off[n_reads] = cur;
    for (int i = 0; i < n_reads; i++) { // the compressed bytes: the only per-record copy the host makes
        if (record_bytes[i] > 0)
            memcpy(s.h_rec + rec_off[i], records[i], (size_t)record_bytes[i]);
        memset(s.h_rec + rec_off[i] + record_bytes[i], 0, (size_t)(rec_off[i + 1] - rec_off[i] - record_bytes[i]));
    }
    s.n_reads = n_reads;
    s.n_samples = cur;
    s.raw_samples = raw;
    s.rec_total = rcur;
    s.queries_only = false;
    s.records = true;
    s.record_press = record_press;
    s.signal_press = signal_press;
    g_trace(c->opt.device, "submit_records: records staged in pinned memory");
    rc = run_stages(c, s, true);
    g_trace(c->opt.device, "submit_records: stages enqueued");
    return rc;
}

int64_t sfgpu_slot_signal(sfgpu_ctx *c, int32_t slot, int32_t read, int16_t *out, int64_t cap)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (slot < 0 || slot >= (int)c->slots.size())
        return fail(c, SFGPU_EARG, "bad slot");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    if (!s.done || s.queries_only || read < 0 || read >= s.n_reads)
        return fail(c, SFGPU_EARG, "no such read in slot %d", slot);
    const int64_t len = s.h_off[s.n_reads + 1 + read];
    if (len > cap || (len > 0 && !out))
        return fail(c, SFGPU_EARG, "buffer too small");
    if (len > 0)
        SF_CUDA(c, cudaMemcpy(out, s.d_signal + s.h_off[read], sizeof(int16_t) * (size_t)len, cudaMemcpyDeviceToHost));
    return len;
}

int sfgpu_submit_queries(sfgpu_ctx *c, int32_t slot, int32_t n_reads, const float *queries,
                         const int32_t *qlen)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (!c->have_ref)
        return fail(c, SFGPU_ESTATE, "sfgpu_submit_queries before a reference was set");
    if (slot < 0 || slot >= (int)c->slots.size() || n_reads < 0)
        return fail(c, SFGPU_EARG, "bad slot or read count");
    if (n_reads > 0 && (!queries || !qlen))
        return fail(c, SFGPU_EARG, "null batch array");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    rc = slot_reserve(c, s, std::max(n_reads, 1), 8);
    if (rc)
        return rc;
    if (!s.h_queries)
        SF_CUDA(c, cudaMallocHost(&s.h_queries, sizeof(float) * (size_t)s.cap_reads * c->q_cap));
    const int q = c->opt.query_size;
    for (int i = 0; i < n_reads; i++) {
        if (qlen[i] < 0 || qlen[i] > q)
            return fail(c, SFGPU_EARG, "qlen[%d]=%d outside [0, query_size]", i, qlen[i]);
        float *dst = s.h_queries + (size_t)i * c->q_cap;
        memcpy(dst, queries + (size_t)i * q, sizeof(float) * qlen[i]);
        for (int j = qlen[i]; j < c->q_cap; j++)
            dst[j] = 0.0f;
        sf_readinfo ri;
        memset(&ri, 0, sizeof ri);
        ri.n_events = qlen[i];
        ri.qend = qlen[i];
        ri.qlen = qlen[i];
        s.h_info[i] = ri;
    }
    s.n_reads = n_reads;
    s.n_samples = 0;
    s.raw_samples = 0;
    s.queries_only = true;
    s.records = false;
    return run_stages(c, s, true, false);
}

int sfgpu_resubmit(sfgpu_ctx *c, int32_t slot)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (slot < 0 || slot >= (int)c->slots.size())
        return fail(c, SFGPU_EARG, "bad slot");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    if (!s.done)
        return fail(c, SFGPU_ESTATE, "sfgpu_resubmit: slot holds no batch");
    return run_stages(c, s, false, !s.queries_only);
}

int sfgpu_collect(sfgpu_ctx *c, int32_t slot, sfgpu_result_t *out)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (slot < 0 || slot >= (int)c->slots.size())
        return fail(c, SFGPU_EARG, "bad slot");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    if (!s.busy && !s.done)
        return fail(c, SFGPU_ESTATE, "sfgpu_collect: nothing was submitted to slot %d", slot);
    g_trace(c->opt.device, "collect: begin");
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    g_trace(c->opt.device, "collect: batch done");
    if (!out && s.n_reads > 0)
        return fail(c, SFGPU_EARG, "null result array");
    if (s.records)
        for (int i = 0; i < s.n_reads; i++)
            if (s.h_status[i] != 0)
                return fail(c, SFGPU_EDECODE, "record %d of the batch could not be decoded on the device (status %d)", i, s.h_status[i]);
    for (int i = 0; i < s.n_reads; i++) {
        const sf_readinfo &ri = s.h_info[i];
        const sf_hit &h = s.h_hits[i];
        sfgpu_result_t &o = out[i];
        o.n_events = ri.n_events;
        o.qstart = ri.qstart;
        o.qend = ri.qend;
        o.qlen = ri.qlen;
        o.status = ri.status & 31; // bit 5 is internal (read went through the pair kernel)
        o.start_raw = ri.start_raw;
        o.end_raw = ri.end_raw;
        o.score = h.score;
        o.score2 = h.score2;
        o.rid = h.rid;
        o.strand = h.strand;
        o.pos_st = h.pos_st;
        o.pos_end = h.pos_end;
    }
    return SFGPU_OK;
}

int sfgpu_collect_paths(sfgpu_ctx *c, int32_t slot, const int64_t *move_off, uint8_t *moves,
                        int32_t *n_moves, uint64_t *ev_start, float *ev_len)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (!(c->opt.flags & SFGPU_SAM))
        return fail(c, SFGPU_ESTATE, "sfgpu_collect_paths needs a context created with SFGPU_SAM");
    if (slot < 0 || slot >= (int)c->slots.size())
        return fail(c, SFGPU_EARG, "bad slot");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    if (!s.done || s.queries_only)
        return fail(c, SFGPU_ESTATE, "sfgpu_collect_paths: slot holds no finished read batch");
    const int n = s.n_reads;
    if (n == 0)
        return SFGPU_OK;
    if (!move_off || !moves || !n_moves || !ev_start || !ev_len)
        return fail(c, SFGPU_EARG, "null output array");
    const bool std_dtw = (c->opt.flags & SFGPU_DTW) != 0;
    const int q = c->opt.query_size;
    // columns each read's direction window needs, and the caller's move capacity
    std::vector<int64_t> dir_off(n);
    int64_t cols = 0;
    for (int i = 0; i < n; i++) {
        const sf_hit &h = s.h_hits[i];
        const int qlen = s.h_info[i].qlen;
        dir_off[i] = -1;
        if (qlen <= 0 || h.seg < 0 || h.pos_st < 0 || h.pos_end < h.pos_st)
            continue;
        const int64_t width = (int64_t)h.pos_end - (std_dtw ? 0 : h.pos_st) + 1;
        const int64_t need = (int64_t)qlen + h.pos_end - h.pos_st;
        if (move_off[i + 1] - move_off[i] < need)
            return fail(c, SFGPU_EARG, "move buffer of read %d holds %lld entries, %lld needed", i,
                        (long long)(move_off[i + 1] - move_off[i]), (long long)need);
        dir_off[i] = cols;
        cols += width;
    }
    const int64_t total_moves = move_off[n];
    unsigned long long *d_dirs = nullptr;
    int64_t *d_dir_off = nullptr, *d_move_off = nullptr;
    uint8_t *d_moves = nullptr;
    int32_t *d_n = nullptr, *d_sc = nullptr;
    std::vector<int32_t> start_col(n);
    std::vector<uint64_t> ws((size_t)n * c->q_cap);
    std::vector<float> wl((size_t)n * c->q_cap);
    rc = [&]() -> int {
        SF_CUDA(c, cudaMalloc(&d_dirs, sizeof(unsigned long long) * 32 * (size_t)std::max<int64_t>(cols, 1)));
        SF_CUDA(c, cudaMalloc(&d_dir_off, sizeof(int64_t) * n));
        SF_CUDA(c, cudaMalloc(&d_move_off, sizeof(int64_t) * (n + 1)));
        SF_CUDA(c, cudaMalloc(&d_moves, (size_t)std::max<int64_t>(total_moves, 1)));
        SF_CUDA(c, cudaMalloc(&d_n, sizeof(int32_t) * n));
        SF_CUDA(c, cudaMalloc(&d_sc, sizeof(int32_t) * n));
        SF_CUDA(c, cudaMemcpyAsync(d_dir_off, dir_off.data(), sizeof(int64_t) * n, cudaMemcpyHostToDevice, s.stream));
        SF_CUDA(c, cudaMemcpyAsync(d_move_off, move_off, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice, s.stream));
        sf_path_args pa;
        pa.stream = c->d_stream;
        pa.segs = c->d_segs;
        pa.queries = s.d_queries;
        pa.info = s.d_info;
        pa.q_cap = c->q_cap;
        pa.hits = s.d_hits;
        pa.n_reads = n;
        pa.dir_off = d_dir_off;
        pa.dirs = d_dirs;
        pa.move_off = d_move_off;
        pa.moves = d_moves;
        pa.n_moves = d_n;
        pa.start_col = d_sc;
        cudaError_t e = cudaErrorInvalidValue;
        g_trace(c->opt.device, "collect_paths: buffers ready");
        SF_DISPATCH_R(c->R, std_dtw, (e = launch_path<R, STD>(pa, s.stream)));
        SF_CUDA(c, e);
        if (g_trace.on) {
            SF_CUDA(c, cudaStreamSynchronize(s.stream));
            g_trace(c->opt.device, "collect_paths: path kernel done");
        }
        if (total_moves > 0)
            SF_CUDA(c, cudaMemcpyAsync(moves, d_moves, (size_t)total_moves, cudaMemcpyDeviceToHost, s.stream));
        SF_CUDA(c, cudaMemcpyAsync(n_moves, d_n, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s.stream));
        SF_CUDA(c, cudaMemcpyAsync(start_col.data(), d_sc, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s.stream));
        SF_CUDA(c, cudaMemcpyAsync(ws.data(), s.d_win_start, sizeof(uint64_t) * ws.size(), cudaMemcpyDeviceToHost, s.stream));
        SF_CUDA(c, cudaMemcpyAsync(wl.data(), s.d_win_len, sizeof(float) * wl.size(), cudaMemcpyDeviceToHost, s.stream));
        SF_CUDA(c, cudaStreamSynchronize(s.stream));
        g_trace(c->opt.device, "collect_paths: copied back");
        return SFGPU_OK;
    }();
    dfree(d_dirs); dfree(d_dir_off); dfree(d_move_off); dfree(d_moves); dfree(d_n); dfree(d_sc);
    if (rc != SFGPU_OK)
        return rc;
    for (int i = 0; i < n; i++) {
        const int qlen = s.h_info[i].qlen;
        for (int k = 0; k < q; k++) {
            ev_start[(size_t)i * q + k] = k < qlen ? ws[(size_t)i * c->q_cap + k] : 0;
            ev_len[(size_t)i * q + k] = k < qlen ? wl[(size_t)i * c->q_cap + k] : 0.0f;
        }
        // the backtrack must end where the start-pointer pass said it would
        if (n_moves[i] >= 0 && start_col[i] != s.h_hits[i].pos_st)
            return fail(c, SFGPU_ECUDA, "internal: path of read %d starts at column %d, start-coordinate pass gave %d", i,
                        start_col[i], s.h_hits[i].pos_st);
    }
    return SFGPU_OK;
}

int sfgpu_timing(sfgpu_ctx *c, int32_t slot, sfgpu_timing_t *t)
{
    if (!c || !t)
        return fail(c, SFGPU_EARG, "null argument");
    if (slot < 0 || slot >= (int)c->slots.size())
        return fail(c, SFGPU_EARG, "bad slot");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    if (!s.done)
        return fail(c, SFGPU_ESTATE, "sfgpu_timing: slot holds no finished batch");
    *t = s.timing;
    return SFGPU_OK;
}

int sfgpu_ref_events(sfgpu_ctx *c, int32_t rid, int32_t strand, float *out, int32_t cap)
{
    if (!c || !c->have_ref)
        return fail(c, SFGPU_ESTATE, "no reference");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    for (const auto &sg : c->segs) {
        if (sg.rid == rid && sg.strand == strand) {
            if (!out || cap < sg.rlen)
                return fail(c, SFGPU_EARG, "buffer too small (%d < %d)", cap, sg.rlen);
            SF_CUDA(c, cudaMemcpy(out, c->d_stream + sg.off, sizeof(float) * sg.rlen, cudaMemcpyDeviceToHost));
            return sg.rlen;
        }
    }
    return fail(c, SFGPU_EARG, "no such (contig, strand)");
}

int64_t sfgpu_event_table(sfgpu_ctx *c, const int16_t *signal, int64_t n_samples, float digitisation,
                          float offset, float range, uint64_t *start, float *length, float *mean,
                          int64_t cap)
{
    if (!c || !signal || n_samples <= 0)
        return fail(c, SFGPU_EARG, "bad argument");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    const int64_t padded = (n_samples + 7) & ~7ll;
    const int32_t ev_cap = (int32_t)std::min<int64_t>(n_samples + 1, 0x7fffffff);
    int16_t *d_sig = nullptr;
    int64_t *d_off = nullptr;
    float *d_scal = nullptr, *d_mean = nullptr, *d_len = nullptr, *d_q = nullptr;
    uint64_t *d_start = nullptr;
    sf_readinfo *d_info = nullptr;
    sf_readinfo info;
    memset(&info, 0, sizeof info);
    int64_t n_ev = 0;
    int rc = [&]() -> int {
        SF_CUDA(c, cudaMalloc(&d_sig, sizeof(int16_t) * (padded + 8)));
        SF_CUDA(c, cudaMemset(d_sig, 0, sizeof(int16_t) * (padded + 8)));
        SF_CUDA(c, cudaMemcpy(d_sig, signal, sizeof(int16_t) * n_samples, cudaMemcpyHostToDevice));
        const int64_t offs[3] = {0, padded, n_samples};
        SF_CUDA(c, cudaMalloc(&d_off, sizeof offs));
        SF_CUDA(c, cudaMemcpy(d_off, offs, sizeof offs, cudaMemcpyHostToDevice));
        const float sc[3] = {digitisation, offset, range};
        SF_CUDA(c, cudaMalloc(&d_scal, sizeof sc));
        SF_CUDA(c, cudaMemcpy(d_scal, sc, sizeof sc, cudaMemcpyHostToDevice));
        SF_CUDA(c, cudaMalloc(&d_start, sizeof(uint64_t) * ev_cap));
        SF_CUDA(c, cudaMalloc(&d_mean, sizeof(float) * ev_cap));
        SF_CUDA(c, cudaMalloc(&d_len, sizeof(float) * ev_cap));
        SF_CUDA(c, cudaMalloc(&d_q, sizeof(float) * c->q_cap));
        SF_CUDA(c, cudaMalloc(&d_info, sizeof(sf_readinfo)));
        sf_ev_args ea;
        ea.signal = d_sig;
        ea.sig_off = d_off;
        ea.sig_len = d_off + 2;
        ea.digitisation = d_scal;
        ea.offset = d_scal + 1;
        ea.range = d_scal + 2;
        ea.n_reads = 1;
        ea.flags = c->opt.flags;
        ea.q = c->opt.query_size;
        ea.p = c->opt.prefix_size;
        ea.ev_cap = ev_cap;
        ea.ev_start = d_start;
        ea.ev_mean = d_mean;
        ea.ev_len = d_len;
        ea.queries = d_q;
        ea.q_cap = c->q_cap;
        ea.info = d_info;
        ea.keep_all = 1;
        ea.polya_end = nullptr;
        ea.cap_a = 0;
        ea.win_start = nullptr;
        ea.win_len = nullptr;
        sf_events_kernel<<<1, SF_EV_THREADS, sf_events_smem_bytes()>>>(ea);
        SF_CUDA(c, cudaGetLastError());
        SF_CUDA(c, cudaDeviceSynchronize());
        SF_CUDA(c, cudaMemcpy(&info, d_info, sizeof info, cudaMemcpyDeviceToHost));
        n_ev = info.n_events;
        const int64_t m = std::min<int64_t>(n_ev, cap);
        if (m > 0) {
            if (start) SF_CUDA(c, cudaMemcpy(start, d_start, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost));
            if (mean) SF_CUDA(c, cudaMemcpy(mean, d_mean, sizeof(float) * m, cudaMemcpyDeviceToHost));
            if (length) SF_CUDA(c, cudaMemcpy(length, d_len, sizeof(float) * m, cudaMemcpyDeviceToHost));
        }
        return SFGPU_OK;
    }();
    dfree(d_sig); dfree(d_off); dfree(d_scal); dfree(d_start); dfree(d_mean); dfree(d_len); dfree(d_q); dfree(d_info);
    if (rc != SFGPU_OK)
        return rc;
    return n_ev;
}

int sfgpu_query(sfgpu_ctx *c, int32_t slot, int32_t read, float *out, int32_t cap)
{
    if (!c)
        return fail(nullptr, SFGPU_EARG, "null context");
    if (slot < 0 || slot >= (int)c->slots.size())
        return fail(c, SFGPU_EARG, "bad slot");
    SF_CUDA(c, cudaSetDevice(c->opt.device));
    sf_slot &s = c->slots[slot];
    int rc = slot_wait(c, s);
    if (rc)
        return rc;
    if (!s.done || read < 0 || read >= s.n_reads)
        return fail(c, SFGPU_EARG, "no such read in slot %d", slot);
    const int qlen = s.h_info[read].qlen;
    if (qlen > cap || (qlen > 0 && !out))
        return fail(c, SFGPU_EARG, "buffer too small");
    if (qlen > 0)
        SF_CUDA(c, cudaMemcpy(out, s.d_queries + (size_t)read * c->q_cap, sizeof(float) * qlen, cudaMemcpyDeviceToHost));
    return qlen;
}

int64_t sfgpu_ref_columns(const sfgpu_ctx *c) { return c ? c->ref_columns : 0; }

int32_t sfgpu_wave_reads(const sfgpu_ctx *c)
{
    if (!c || !c->have_ref || c->n_groups <= 0)
        return 0;
    const int64_t reads = c->R2 > 0 ? (int64_t)c->sm_count * c->pair_blocks_per_sm * SF_DTW_WARPS * 2
                                    : (int64_t)c->sm_count * c->dtw_blocks_per_sm * SF_DTW_WARPS;
    return (int32_t)std::max<int64_t>(1, reads / c->n_groups);
}

} // extern "C"
