// Shared declarations of the seeksv_b200 kernel library (sm_100a only, no fallback paths).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <string>
#include <vector>

#include "../../include/seeksv_b200.h"

#define SVB_SM_COUNT 148  // B200: 148 SMs (grids are sized from the live device attribute, this is the design point)

// ---- BAM wire format (sam/bam.h:128-198 of the reference's vendored headers) -----------------------
enum : uint32_t {
    OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_H = 5, OP_P = 6, OP_EQ = 7, OP_X = 8
};
enum : uint32_t {
    F_PAIRED = 1, F_PROPER = 2, F_UNMAP = 4, F_MUNMAP = 8, F_REVERSE = 16, F_MREVERSE = 32, F_READ1 = 64,
    F_READ2 = 128, F_SECONDARY = 256, F_QCFAIL = 512, F_DUP = 1024
};

struct ProfEntry {
    double ms = 0, bytes = 0;
    int64_t launches = 0;
};

struct svb_ctx {
    int device = 0;
    int sm_count = SVB_SM_COUNT;
    cudaStream_t stream = nullptr;
    // side streams of the pipelined load (svb_bam_from_bgzf): one for the host->device copies, a few for the inflate
    // launches that follow each uploaded slab
    static constexpr int N_AUX = 8;
    cudaStream_t copy_stream = nullptr, aux[N_AUX] = {};
    cudaEvent_t fork_event = nullptr;
    bool gz_tables_ready = false;  // gzip.cu's constant tables are on this device  // orders work handed from `stream` to `copy_stream`
    std::string err;
    bool prof = false;
    std::map<std::string, ProfEntry> prof_acc;
    struct Pending {
        std::string name;
        cudaEvent_t a, b;
        double bytes;
    };
    std::vector<Pending> prof_pending;
    std::vector<cudaEvent_t> prof_events;  // recycled timing events (cudaEventCreate per scope cost more than the kernels it timed)
    cudaEvent_t prof_event()
    {
        cudaEvent_t e = nullptr;
        if (!prof_events.empty()) e = prof_events.back(), prof_events.pop_back();
        else cudaEventCreate(&e);
        return e;
    }
    std::vector<const char *> prof_names;  // storage for svb_prof_read
    void prof_flush();
    // recycled pinned host buffers (cudaHostAlloc is slow; results are produced over and over)
    std::vector<std::pair<char *, uint64_t>> pinned_free;
    char *pinned_get(uint64_t bytes, uint64_t *cap);
    void pinned_put(char *p, uint64_t cap);
    // The two large device buffers of a load (the uncompressed stream and the compressed file image) are kept for the next
    // load of this context instead of going back to the pool: the pool may have split them for smaller requests in the
    // meantime, and growing it again by gigabytes stalls a load for hundreds of milliseconds.
    std::vector<std::pair<uint8_t *, uint64_t>> big_free;
    uint8_t *big_get(uint64_t bytes, uint64_t *cap);  // contents undefined; usable on `stream` (and after its events)
    void big_put(uint8_t *p, uint64_t cap);           // caller has synchronised every stream that used p
    // Mid-size device buffers that a command needs again and again with about the same size (record rows, result texts): kept
    // in the context instead of going back to the stream-ordered pool - the pool splits a big freed block for the next small
    // request and then has to map fresh memory for the big one again (milliseconds per step, measured).
    std::vector<std::pair<uint8_t *, uint64_t>> dev_free;
    uint8_t *dev_get(uint64_t bytes, uint64_t *cap);  // usable on `stream` (and after its events)
    void dev_put(uint8_t *p, uint64_t cap);           // p's last use is ordered before later work on `stream`
    // One grow-only scratch buffer per command family: the sync-free pipelines carve all their temporaries out of it (Bump,
    // prim.cuh) instead of allocating array by array. A command owns it from its first launch to its final read-back.
    uint8_t *ws[2] = {nullptr, nullptr};
    uint64_t ws_cap[2] = {0, 0};
    int ws_reserve(int which, uint64_t bytes);        // (re)allocates when too small; contents undefined
    // small pinned block for the control-block read-backs
    char *ctl_host = nullptr;
    // capacity hints: what the last svb_getclip of this context needed (texts, arenas, queues), so that a repeated call
    // on similar data never overflows its first estimate
    uint64_t hint[16] = {};
    bool walk_attr = false;                           // the walker's shared-memory opt-in is set on this device
    cudaEvent_t join_event = nullptr;                 // side stream -> main stream
    char *read_buf = nullptr;                         // svb_read_gz_device: the text it returned last (pinned pool)
    uint64_t read_cap = 0;
    cudaEvent_t sw_event = nullptr;                   // side stream -> main stream: getclip's sorted chromosome-switch list is there
    // inflate.cu: slot bitmap + per-resident-warp match lists of the speculative inflate kernel (allocated on first use)
    uint8_t *inflate_scratch = nullptr;
    uint32_t inflate_slot_words = 0, inflate_slots_per_sm = 0, inflate_scratch_head = 0;
    ~svb_ctx();
};

struct PinnedBuf {
    char *p = nullptr;
    uint64_t n = 0, cap = 0;
    int reserve(svb_ctx *ctx, uint64_t bytes)
    {
        release(ctx);
        p = ctx->pinned_get(bytes, &cap);
        n = p ? bytes : 0;
        return (p || bytes == 0) ? 0 : SVB_ERR_CUDA;
    }
    void release(svb_ctx *ctx)
    {
        if (p && ctx) ctx->pinned_put(p, cap);
        p = nullptr, n = cap = 0;
    }
};

// Lean per-record rows produced by the record walker (the getsv passes read these, not the raw stream): 32 bytes, written with
// ONE 256-bit store per record. Rows live in per-chunk slots - record k of chunk c is rows[c * R + k] - so the walker needs no
// per-chunk row base (which would cost a counting pass over the stream first) and the same pass can serve getclip and getsv.
// Measured on the B200 with tools/walk_lab.cu: 32-byte rows 0.39 ms, 48-byte rows (three 16-byte stores) 0.47 ms, dense or slotted alike.
struct __align__(32) Row {
    int32_t tid, pos, end;  // end = bam_calend of the linked libbam (M, D, N)
    uint32_t flagq;         // flag | mapq << 16 | hardclip << 24 | no-cigar << 25
    int32_t lqseq, mtid, mpos, isize;
};
struct RowTable {
    Row *row = nullptr;
    uint16_t *roff = nullptr;  // offset of the record inside its chunk (behind the rows, same allocation): CIGARs are re-read from there
    uint32_t R = 0;            // slots per chunk
    uint64_t cap = 0;          // bytes of the allocation (dev_get)
};
#define FLAGQ_HARDCLIP (1u << 24)
#define FLAGQ_NOCIGAR (1u << 25)
// row slots per chunk: first try fits chunks whose records average >= 64 bytes; a chunk cannot hold more records than the
// second (a record is at least 38 bytes)
#define ROWS_R_FIRST 256u
#define ROWS_R_MAX 448u

struct svb_bam {
    svb_ctx *ctx = nullptr;
    const uint8_t *d_data = nullptr;
    uint8_t *d_owned = nullptr;
    uint64_t owned_cap = 0;
    uint64_t nbytes = 0, first = 0;
    int32_t n_ref = 0;
    uint64_t n_rec = 0, rec_bytes = 0;
    // record chain (bam_index.cu): the stream is cut into 16 KiB chunks; guess[c] = offset of the first record that
    // starts at or after the chunk, count[c] = records starting inside it, base[c] = exclusive prefix of count
    uint64_t n_chunks = 0;
    uint32_t chunk_log2 = 14;  // 16 KiB chunks = streaming tiles (SEEKSV_B200_CHUNK_LOG2 changes it for the walkers)
    uint64_t *d_guess = nullptr, *d_base = nullptr, *d_exit = nullptr;
    uint32_t *d_count = nullptr;
    bool counted = false;  // count / base / n_rec / rec_bytes are valid (a verified walk has run)
    bool guessed = false;  // d_guess holds first-record guesses (guess_starts)
    bool whole_file = false;  // built from a complete BAM: the chain must end exactly at the end of the stream
    std::vector<std::string> names;
    std::vector<uint32_t> lens;
    // rows of the records (walk.cu), valid when rows_ready; the getsv passes keep their per-BAM scalars on the device (getsv.cu)
    RowTable rows;
    bool rows_ready = false;
    uint64_t *d_fkey = nullptr;   // per chunk: (tid, pos) of the first record at or after the chunk, tid -1 last (rows_index)
    int32_t *d_scal = nullptr;    // [0] max span, [1] unsorted flag, [2] indexed flag
    bool rows_indexed = false;
    int32_t max_span = 0;  // max(end - pos) over records, for window queries
    int sorted = -1;       // -1 unknown, 0 no, 1 coordinate-sorted
    uint32_t *d_ref_len = nullptr;  // reference lengths on the device (uploaded once per handle)
    uint64_t own_offset = 0;        // range shards: the getsv passes ignore records that start before this offset (halo)
};

// ---- error plumbing ------------------------------------------------------------------------------------
int svb_fail(svb_ctx *ctx, int code, const char *fmt, ...);
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return svb_fail(ctx, SVB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define CKR(call)                      \
    do {                               \
        int r_ = (call);               \
        if (r_ != 0) return r_;        \
    } while (0)

// RAII timer around one kernel launch (CUDA events on the ctx stream, only when profiling is on)
struct ProfScope {
    svb_ctx *ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    const char *name;
    double bytes;
    cudaStream_t st;
    ProfScope(svb_ctx *c, const char *n, double by, cudaStream_t on = nullptr) : ctx(c), name(n), bytes(by), st(on ? on : c->stream)
    {
        if (ctx->prof) {
            a = ctx->prof_event(), b = ctx->prof_event();
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope()
    {
        if (ctx->prof) {
            cudaEventRecord(b, st);
            ctx->prof_pending.push_back({name, a, b, bytes});
        }
    }
};

// device scratch that is released on scope exit (stream-ordered pool allocations)
template <typename T>
struct DevBuf {
    T *p = nullptr;
    cudaStream_t s = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    cudaError_t alloc(size_t count, cudaStream_t st)
    {
        release();
        s = st;
        n = count;
        return cudaMallocAsync((void **)&p, (count ? count : 1) * sizeof(T), st);
    }
    void release()
    {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~DevBuf() { release(); }
    T *steal()
    {
        T *q = p;
        p = nullptr;
        return q;
    }
};

// ---- device helpers -------------------------------------------------------------------------------------
#ifdef __CUDACC__
// 32-bit little-endian load from an arbitrarily aligned address (BAM records are byte-packed). Two aligned
// word loads + funnel shift; the stream buffer is padded so the second word is always readable.
__device__ __forceinline__ uint32_t ldu32(const uint8_t *p)
{
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t sh = ((uint32_t)a & 3u) * 8u;
    uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    uint32_t hi = __ldg(w + 1);
    return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ int32_t ldi32(const uint8_t *p) { return (int32_t)ldu32(p); }

// The 32-byte fixed core of a record (after the 4-byte block_size), decoded.
struct Core {
    int32_t block_size, tid, pos, l_qseq, mtid, mpos, isize;
    uint32_t l_qname, mapq, n_cigar, flag;
};
template <int WI>
__device__ __forceinline__ void core_fields(const uint32_t (&W)[16], uint32_t sh, uint32_t (&f)[9])
{
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = sh ? __funnelshift_r(W[WI + i], W[WI + i + 1], sh) : W[WI + i];
}
__device__ __forceinline__ Core load_core(const uint8_t *p)
{
    // The 36 bytes sit at an arbitrary byte offset. They are fetched with three or four aligned 16-byte loads (one L1
    // wavefront each per lane) instead of ten 4-byte loads: the walkers are uncoalesced by nature - every lane follows its
    // own chain - and were limited by L1 wavefronts, not by DRAM (0.48 ms measured against 0.29 ms of line fetches).
    uintptr_t a = (uintptr_t)p;
    const uint4 *q = (const uint4 *)(a & ~(uintptr_t)15);
    uint32_t in16 = (uint32_t)a & 15u, sh = ((uint32_t)a & 3u) * 8u;
    uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = make_uint4(0, 0, 0, 0);
    if (in16 + 40 > 48) v3 = __ldg(q + 3);
    uint32_t W[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
    uint32_t f[9];
    switch (in16 >> 2) {
    case 0: core_fields<0>(W, sh, f); break;
    case 1: core_fields<1>(W, sh, f); break;
    case 2: core_fields<2>(W, sh, f); break;
    default: core_fields<3>(W, sh, f); break;
    }
    Core c;
    c.block_size = (int32_t)f[0];
    c.tid = (int32_t)f[1];
    c.pos = (int32_t)f[2];
    c.l_qname = f[3] & 0xff;  // l_read_name:8 mapq:8 bin:16
    c.mapq = (f[3] >> 8) & 0xff;
    c.n_cigar = f[4] & 0xffff;  // n_cigar:16 flag:16
    c.flag = f[4] >> 16;
    c.l_qseq = (int32_t)f[5];
    c.mtid = (int32_t)f[6];
    c.mpos = (int32_t)f[7];
    c.isize = (int32_t)f[8];
    return c;
}
__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ uint32_t warp_max(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#endif

// ---- internal entry points (one per .cu) ------------------------------------------------------------------
static constexpr uint64_t BAD_OFFSET = ~0ull;
#define NO_TID INT32_MIN
int index_records(svb_ctx *ctx, svb_bam *bam);  // walk.cu: chunk arrays (guesses are made by the first pass that needs them)
int ensure_guess(svb_ctx *ctx, svb_bam *bam);   // walk.cu: guess_starts once
// walk.cu: verified chain + per-chunk counts + their prefix + rows (one pass of the record walker without the getclip work)
int ensure_counts(svb_ctx *ctx, svb_bam *bam);
int ensure_rows(svb_ctx *ctx, svb_bam *bam);
// a walk left exit[] without matching the guesses: repair guesses with plain walks (synchronises); the walk has to run again
int repair_guesses(svb_ctx *ctx, svb_bam *bam);
static inline uint32_t rows_first(const svb_bam *b) { return std::max(16u, ROWS_R_FIRST >> (14 - b->chunk_log2)); }
static inline uint32_t rows_max(const svb_bam *b) { return (ROWS_R_MAX >> (14 - b->chunk_log2)) + 4; }
int alloc_rows(svb_ctx *ctx, svb_bam *bam, uint32_t R);  // walk.cu
void free_rows(svb_bam *bam);
// the inflate kernel reads ahead of the current bit position: the device copy of the file image is padded by this much
static constexpr uint64_t SVB_INFLATE_PAD = 1024;
struct PinnedBuf;
int gzip_on_device(svb_ctx *ctx, const char *d_text, uint64_t n, PinnedBuf *out);  // gzip.cu
int inflate_launch(svb_ctx *ctx, cudaStream_t s, const uint8_t *d_file, const void *d_blocks, uint32_t n_blocks, uint8_t *d_out, uint32_t *d_err);
int inflate_on_device(svb_ctx *ctx, const uint8_t *d_file, const void *d_blocks, uint32_t n_blocks, uint8_t *d_out,
                      double out_bytes);                             // inflate.cu
