// Interface of the record walker (walk.cu) towards the two commands.
#pragma once
#include "common.cuh"
#include "prim.cuh"

// getclip's share of the walk: queues of record offsets (filled in arbitrary order, sorted later) and, per chunk, what the
// chromosome-switch rule (quirk Q1) needs about the chunk's first and last mapped-branch record
struct ClipQueues {
    uint64_t *clipped;   // records whose first or last CIGAR op is S and that pass the cheap filters: evaluated by clip_eval
    uint64_t *unmapped;  // unmapped-branch records (clip_reads.h:415)
    uint64_t *switches;  // mapped-branch records whose tid differs from the previous one's (flushed and dropped, clip_reads.h:423-438)
    uint32_t clipped_cap, un_cap, sw_cap;
    uint32_t *counters;  // [0] clipped [1] unmapped [2] switches - they keep counting past the capacities (overflow = retry)
    uint64_t *first_mb;  // per chunk: offset of its first mapped-branch record (BAD_OFFSET: none) - judged by clip_first
    int32_t *last_mb_tid;  // per chunk: tid of its last mapped-branch record (NO_TID: none)
    int32_t min_mapq;
};

// One pass of the walker over the whole stream of `bam` on stream s. clip: fill q (its counters must be zero); rows: write
// bam->rows (allocated by the caller, R slots per chunk); flags[0] is set when a chunk has more records than R. `ticket` must be
// zero. Always writes bam->d_count and bam->d_exit.
int launch_walk(svb_ctx *ctx, cudaStream_t s, svb_bam *bam, bool clip, bool rows, const ClipQueues &q, uint32_t *flags,
                unsigned long long *ticket);

// After a walk, on the same stream: exit(c - 1) == guess(c) for every chunk (else ctl[0] |= 1), exclusive prefix of the chunk
// counts -> bam->d_base, number of records -> ctl64[0], byte offset where the chain ends -> ctl64[1].
void launch_chunk_scan(svb_ctx *ctx, cudaStream_t s, svb_bam *bam, const ScanScratch &sc, uint32_t *ctl_bad, uint64_t *ctl64);

// host side of the above once the control words are back: fills n_rec / rec_bytes / counted, checks the end of a whole file
int accept_counts(svb_ctx *ctx, svb_bam *bam, uint64_t n_rec, uint64_t chain_end);

#ifdef __CUDACC__
// Warp-buffered append to a global queue: the walker's lanes meet a queued record every ~50 records, and one global atomic per
// record on ONE counter serialises in L2 (measured: 0.70 ms for 450 k appends against 0.32 ms for the whole pass without them,
// tools/walk_lab.cu). The warp collects entries in shared memory and flushes 32 at a time with one atomic.
static constexpr int WQ_CAP = 64;
struct WarpQueue {
    uint64_t e[WQ_CAP];
    uint32_t n, pad;
};
__device__ __forceinline__ void wq_flush32(WarpQueue &q, uint32_t lane, uint32_t count, uint32_t *counter, uint64_t *dst, uint32_t cap)
{
    uint32_t b0 = 0;
    if (lane == 0) b0 = atomicAdd(counter, count);
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    if (lane < count && b0 + lane < cap) dst[b0 + lane] = q.e[lane];
}
// all 32 lanes call this together; `has` lanes append `o`
__device__ __forceinline__ void wq_push(WarpQueue &q, bool has, uint64_t o, uint32_t lane, uint32_t *counter, uint64_t *dst, uint32_t cap)
{
    const uint32_t m = __ballot_sync(0xffffffffu, has);
    if (!m) return;
    const uint32_t n0 = q.n;
    if (has) q.e[n0 + __popc(m & ((1u << lane) - 1u))] = o;
    __syncwarp();
    uint32_t n1 = n0 + __popc(m);
    if (n1 >= 32) {
        wq_flush32(q, lane, 32, counter, dst, cap);
        const uint64_t keep = lane + 32 < n1 ? q.e[lane + 32] : 0;
        __syncwarp();
        if (lane + 32 < n1) q.e[lane] = keep;
        n1 -= 32;
    }
    __syncwarp();
    if (lane == 0) q.n = n1;
    __syncwarp();
}
__device__ __forceinline__ void wq_drain(WarpQueue &q, uint32_t lane, uint32_t *counter, uint64_t *dst, uint32_t cap)
{
    const uint32_t n = q.n;
    if (n) wq_flush32(q, lane, n, counter, dst, cap);
    __syncwarp();
    if (lane == 0) q.n = 0;
}
#endif
