// The record walker: the ONE kernel that streams the packed BAM records (both commands use it).
//
// Replaces the serial `while (samread(...) >= 0)` chain of the reference (clip_reads.h:410, cluster.cpp:48, getsv.h:472,
// bam2depth.cpp:75): every record starts where the previous one ends, so the chain is serial by construction. Here the
// stream is cut into 16 KiB chunks; each chunk GUESSES the first record start at or after its beginning (strong
// plausibility test on the 36-byte fixed part, two records deep) and is walked as its own chain. exit(c) == guess(c+1)
// for every chunk proves, by induction from the exact header offset, that the walked chains are the true chain - the
// heuristic affects speed only. Wrong guesses are repaired from the predecessor's exit and the pass is run again.
//
// Form of the walk (round 2; the round-1 form was one thread per chunk with scattered 16-byte loads, 0.50 / 0.70 ms per pass
// on C2 plus a 0.21 ms counting pass - tools/walk_lab.cu holds the measurements that led here):
//   * a warp owns 32 chains, one per lane; lanes whose chunk is exhausted take the next chunk from a ticket counter, so the
//     warps stay full until the stream ends (persistent grid, a multiple of the SM count);
//   * per step every live lane names the 32-byte sectors that hold the head of its next record (fixed part, name, CIGAR - the
//     span is predicted from the previous record); the warp fetches them into shared memory with 16-byte cp.async pieces,
//     sixteen lanes per chain, so one instruction touches 4 lines instead of 32 (the scattered form spent ~9 load instructions
//     of 32 L1 wavefronts each per record) and each lane parses its record from its own shared-memory row;
//   * getclip's work (unmapped branch, chromosome switches, soft-clip predicate) and getsv's (one 32-byte row per record into
//     the chunk's row slots) are template options of the same kernel: a command that needs both pays for one pass;
//   * queue appends go through a per-warp shared-memory buffer (one global atomic per 32 entries): one atomic per entry on a
//     single counter was what bounded round 1's clip_walk (0.70 -> 0.32 ms in the lab).
#include <cstdlib>
#include <cstring>

#include "walk.cuh"

__device__ __forceinline__ bool plausible_one(const uint8_t *d, uint64_t n, uint64_t o, int32_t n_ref, uint64_t *next)
{
    if (o + 36 > n) return false;
    int32_t bs = ldi32(d + o);
    if (bs < 33 || o + 4 + (uint64_t)bs > n) return false;
    int32_t tid = ldi32(d + o + 4);
    if (tid < -1 || tid >= n_ref) return false;
    int32_t pos = ldi32(d + o + 8);
    if (pos < -1 || pos >= (1 << 29)) return false;  // BAM coordinates are below 2^29
    uint32_t w = ldu32(d + o + 12);
    uint32_t l_qname = w & 0xff;
    if (l_qname < 2) return false;
    uint32_t w2 = ldu32(d + o + 16);
    uint32_t n_cigar = w2 & 0xffff;
    if ((w2 >> 16) & 0xf000) return false;  // flag bits above 0x800 are not defined
    int32_t l_qseq = ldi32(d + o + 20);
    if (l_qseq < 0) return false;
    int32_t mtid = ldi32(d + o + 24);
    if (mtid < -1 || mtid >= n_ref) return false;
    int32_t mpos = ldi32(d + o + 28);
    if (mpos < -1 || mpos >= (1 << 29)) return false;
    uint64_t need = 32ull + l_qname + 4ull * n_cigar + ((uint64_t)l_qseq + 1) / 2 + (uint64_t)l_qseq;
    if (need > (uint64_t)bs) return false;
    if ((uint64_t)bs - need > 4ull * (uint64_t)l_qseq + 8192) return false;  // aux block of a sane size
    uint8_t c0 = d[o + 36];
    if (c0 < 33 || c0 > 126) return false;           // qname starts with a printable character ...
    if (d[o + 36 + l_qname - 1] != 0) return false;  // ... and is NUL terminated
    *next = o + 4 + (uint64_t)bs;
    return true;
}

// One warp per chunk: lanes test 32 consecutive byte offsets at a time (global loads; the lines are in L1 after the first touch).
// Two other forms were measured on the B200 and were slower than this 0.12 ms (profiles/r2_summary.md): staging the chunk's first
// kilobyte in shared memory (0.15 ms: the kernel is bound by instruction issue - ~850 warp instructions per chunk - not by memory)
// and a cheap per-offset prefilter followed by one-lane full tests (0.32 ms: the serialised tests cost more than they save).
__global__ void __launch_bounds__(256) guess_starts(const uint8_t *__restrict__ d, uint64_t n, uint64_t first, int32_t n_ref,
                                                    uint64_t n_chunks, uint32_t CHUNK_LOG2, uint64_t *__restrict__ guess)
{
    uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (c >= n_chunks) return;
    uint64_t start = c << CHUNK_LOG2;
    if (start <= first) {
        if (lane == 0) guess[c] = first;
        return;
    }
    // search window: 8 chunks, but never less than the 128 KiB the 16 KiB chunks have always had - with the small chunks of a small
    // stream (index_records) a record of a 40 kb read is longer than 8 chunks and no start was found inside it
    uint64_t limit = min(n, start + max((uint64_t)8 << CHUNK_LOG2, (uint64_t)128 << 10));
    uint64_t found = BAD_OFFSET;
    for (uint64_t base = start; base < limit; base += 32) {
        uint64_t o = base + lane, nx = 0, nx2 = 0;
        bool ok = plausible_one(d, n, o, n_ref, &nx);
        if (ok && nx < n) ok = plausible_one(d, n, nx, n_ref, &nx2);  // two records deep
        uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (m) {
            found = base + (__ffs(m) - 1);
            break;
        }
    }
    if (lane == 0) guess[c] = found == BAD_OFFSET ? n : found;
}

// ---- the staged walker ---------------------------------------------------------------------------------------------------------
namespace {
constexpr int WALK_WARPS = 8;     // warps per CTA
constexpr int STAGE_ROW = 272;    // 256 staged bytes per chain + 16: rows stay 16-byte aligned and start 4 banks apart
struct WalkShared {
    uint8_t stage[WALK_WARPS][32][STAGE_ROW];
    WarpQueue q[WALK_WARPS][2];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ uint32_t lds32(const uint8_t *row, uint32_t off)
{
    const uint32_t *w = (const uint32_t *)row + (off >> 2);
    const uint32_t sh = (off & 3u) * 8u;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}
__device__ __forceinline__ void stg256(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h)
                 : "memory");
}

struct WalkArgs {
    const uint8_t *d;
    uint64_t n, n_chunks;
    uint32_t chunk_log2;
    const uint64_t *guess;
    uint32_t *count;
    uint64_t *exit_;
    Row *rows;
    uint16_t *roff;
    uint32_t R;
    uint32_t *flags;  // [0] a chunk has more records than R
    unsigned long long *ticket;
};

template <bool CLIP, bool ROWS>
__global__ void __launch_bounds__(WALK_WARPS * 32) rec_walk(WalkArgs a, ClipQueues q)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    WalkShared &S = *(WalkShared *)smem_raw;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *stage = &S.stage[warp][0][0];
    const uint8_t *my = &S.stage[warp][lane][0];
    WarpQueue &q_clip = S.q[warp][0], &q_unm = S.q[warp][1];
    if (lane == 0) q_clip.n = 0, q_unm.n = 0;
    __syncwarp();
    const uint8_t *__restrict__ d = a.d;
    const uint64_t n = a.n;
    const uintptr_t base_line = (uintptr_t)d >> 7;  // lines are counted from the one that holds d[0]
    const uint32_t d_in = (uint32_t)((uintptr_t)d & 127);
    // pieces of 16 bytes the walk may fetch: from the piece that holds d[0] to the last one that ends inside the buffer's padding
    const uint64_t piece_lo = d_in >> 4, piece_hi = (n + d_in + 64) >> 4;  // [piece_lo, piece_hi)
    uint64_t o = 0, end = 0, c = 0, first_mb = BAD_OFFSET;
    uint32_t cnt = 0, head_pred = 100;
    int32_t prev_tid = NO_TID;
    bool live = false, done = false;
    for (;;) {
        // lanes without a chain take the next chunks
        const uint32_t need = __ballot_sync(0xffffffffu, !live && !done);
        if (need) {
            unsigned long long b0 = 0;
            const int leader = __ffs(need) - 1;
            if ((int)lane == leader) b0 = atomicAdd(a.ticket, (unsigned long long)__popc(need));
            b0 = __shfl_sync(0xffffffffu, b0, leader);
            if (!live && !done) {
                c = b0 + __popc(need & ((1u << lane) - 1u));
                if (c < a.n_chunks) {
                    o = a.guess[c], end = min(n, (c + 1) << a.chunk_log2);
                    live = o < end && o + 36 <= n;  // (a partial tail shorter than a fixed part ends the walk)
                    cnt = 0, prev_tid = NO_TID, first_mb = BAD_OFFSET;
                    if (!live) {
                        a.count[c] = 0, a.exit_[c] = o;
                        if (CLIP) q.first_mb[c] = BAD_OFFSET, q.last_mb_tid[c] = NO_TID;
                    }
                } else
                    done = true;
            }
        }
        if (__all_sync(0xffffffffu, !live)) {
            if (__all_sync(0xffffffffu, done)) break;
            continue;
        }
        // fetch: the 16-byte pieces [p0, p1] of the (up to) two lines that hold the predicted head
        const uint64_t ao = o + d_in;
        const uint32_t in_line = (uint32_t)ao & 127u;
        uint32_t p0 = (in_line >> 5) << 1, p1 = min(((in_line + head_pred - 1) >> 5) << 1 | 1u, 15u);
        {
            const uint64_t lp = (ao >> 7) << 3;  // first piece of the line
            if (lp + p1 >= piece_hi) p1 = piece_hi > lp ? (uint32_t)(piece_hi - lp) - 1 : 0;
            if (lp + p0 < piece_lo) p0 = (uint32_t)(piece_lo - lp);
        }
        const unsigned long long desc = live ? ((unsigned long long)(ao >> 7) << 9 | p0 << 5 | p1 << 1 | 1u) : 0ull;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t src = 2 * j + (lane >> 4), piece = lane & 15;
            const unsigned long long ds = __shfl_sync(0xffffffffu, desc, src);
            const uint32_t lo = (uint32_t)ds;
            if ((lo & 1u) && piece >= ((lo >> 5) & 15u) && piece <= ((lo >> 1) & 15u))
                cp_async16(stage + src * STAGE_ROW + piece * 16, (const uint8_t *)((base_line + (ds >> 9)) << 7) + piece * 16);
        }
        cp_async_wait_all();
        __syncwarp();
        bool push_clip = false, push_unm = false;
        const uint64_t o_rec = o;
        if (live) {
            const uint32_t avail = (p1 + 1) << 4;  // staged bytes of my row, counted from the line start
            uint32_t f[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) f[i] = lds32(my, in_line + 4 * i);  // (in_line + 36 <= avail: head_pred >= 36, o + 36 <= n)
            const int32_t bs = (int32_t)f[0];
            bool stop = false;
            if (bs < 32) o = BAD_OFFSET, stop = true;  // cannot be a record: corrupt chain (or a wrong guess)
            else if (o + 4 + (uint64_t)bs > n) stop = true;  // partial tail (shard cut mid-record)
            else {
                const int32_t tid = (int32_t)f[1], pos = (int32_t)f[2];
                const uint32_t l_qname = f[3] & 0xff, mapq = (f[3] >> 8) & 0xff, n_cigar = f[4] & 0xffff, flag = f[4] >> 16;
                const uint32_t cig = in_line + 36 + l_qname, head_end = cig + 4 * n_cigar;
                const bool in_smem = head_end <= avail;
                const uint8_t *g = d + o + 36 + l_qname;
                uint32_t first_op = 0, last_op = 0;
                if (ROWS) {
                    int32_t rend = pos;
                    uint32_t fq = flag | (mapq << 16);
                    if (n_cigar == 0) fq |= FLAGQ_NOCIGAR;
                    for (uint32_t j = 0; j < n_cigar; ++j) {
                        const uint32_t w = in_smem ? lds32(my, cig + 4 * j) : ldu32(g + 4 * j), op = w & 15;
                        // bam_calend of the linked libbam: M, D, N only ('=' and 'X' do not advance; probed)
                        if (op == OP_M || op == OP_D || op == OP_N) rend += (int32_t)(w >> 4);
                        if ((j == 0 || j + 1 == n_cigar) && op == OP_H) fq |= FLAGQ_HARDCLIP;  // IsHardClip, clip_reads.cpp:247
                        if (j == 0) first_op = w;
                        last_op = w;
                    }
                    if (cnt < a.R) {
                        stg256(&a.rows[c * a.R + cnt], (uint32_t)tid, (uint32_t)pos, (uint32_t)rend, fq, f[5], f[6], f[7], f[8]);
                        a.roff[c * a.R + cnt] = (uint16_t)(o - (c << a.chunk_log2));  // (a record starts inside its chunk: < 16 KiB)
                    } else
                        a.flags[0] = 1;
                } else if (CLIP && n_cigar) {
                    first_op = in_smem ? lds32(my, cig) : ldu32(g);
                    last_op = in_smem ? lds32(my, head_end - 4) : ldu32(g + 4 * (n_cigar - 1));
                }
                if (CLIP) {
                    if (flag & (F_UNMAP | F_MUNMAP)) push_unm = true;  // clip_reads.h:415 - the unmapped branch wins (quirk Q2)
                    else {
                        if (prev_tid == NO_TID) first_mb = o;  // needs the last mapped-branch record of an earlier chunk: clip_first
                        else if (tid != prev_tid) {            // flush + drop (clip_reads.h:423-438)
                            const uint32_t s = atomicAdd(&q.counters[2], 1u);
                            if (s < q.sw_cap) q.switches[s] = o;
                        } else if (n_cigar != 0 && (int32_t)mapq >= q.min_mapq && !(flag & F_DUP)) {
                            // the cheap part of GetSClipReads (clip_reads.cpp:116-118,122)
                            const uint32_t op1 = first_op & 15, op2 = last_op & 15;
                            push_clip = op1 != OP_H && op2 != OP_H && (op1 == OP_S || op2 == OP_S);
                        }
                        prev_tid = tid;
                    }
                }
                ++cnt;
                head_pred = min(36u + l_qname + 4u * n_cigar + 8u, 129u);
                o += 4 + (uint64_t)bs;
                stop = !(o < end && o + 36 <= n);
            }
            if (stop) {
                live = false;
                a.count[c] = cnt, a.exit_[c] = o;
                if (CLIP) q.first_mb[c] = first_mb, q.last_mb_tid[c] = prev_tid;
            }
        }
        if (CLIP) {
            wq_push(q_clip, push_clip, o_rec, lane, &q.counters[0], q.clipped, q.clipped_cap);
            wq_push(q_unm, push_unm, o_rec, lane, &q.counters[1], q.unmapped, q.un_cap);
        }
        __syncwarp();
    }
    if (CLIP) {
        wq_drain(q_clip, lane, &q.counters[0], q.clipped, q.clipped_cap);
        wq_drain(q_unm, lane, &q.counters[1], q.unmapped, q.un_cap);
    }
}
}  // namespace

static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

int launch_walk(svb_ctx *ctx, cudaStream_t s, svb_bam *bam, bool clip, bool rows, const ClipQueues &q, uint32_t *flags,
                unsigned long long *ticket)
{
    CKR(ensure_guess(ctx, bam));
    WalkArgs a{bam->d_data, bam->nbytes, bam->n_chunks, bam->chunk_log2, bam->d_guess, bam->d_count, bam->d_exit,
               bam->rows.row, bam->rows.roff, bam->rows.R, flags, ticket};
    const int shm = (int)sizeof(WalkShared);
    if (!ctx->walk_attr) {
        CK(cudaFuncSetAttribute(rec_walk<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm));
        CK(cudaFuncSetAttribute(rec_walk<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm));
        CK(cudaFuncSetAttribute(rec_walk<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm));
        CK(cudaFuncSetAttribute(rec_walk<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shm));
        ctx->walk_attr = true;
    }
    // persistent grid: two CTAs of 8 warps per SM = 512 chains per SM (more did not help: the pass is bound by the lines it fetches)
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((bam->n_chunks + WALK_WARPS * 32 - 1) / (WALK_WARPS * 32),
                                                                             (uint64_t)ctx->sm_count * 2));
    const double bytes = (double)(bam->nbytes - bam->first);
    if (clip && rows) {
        ProfScope ps(ctx, "rec_walk_fused", bytes);
        rec_walk<true, true><<<grid, WALK_WARPS * 32, shm, s>>>(a, q);
    } else if (clip) {
        ProfScope ps(ctx, "rec_walk_clip", bytes);
        rec_walk<true, false><<<grid, WALK_WARPS * 32, shm, s>>>(a, q);
    } else if (rows) {
        ProfScope ps(ctx, "rec_walk_rows", bytes);
        rec_walk<false, true><<<grid, WALK_WARPS * 32, shm, s>>>(a, q);
    } else {
        ProfScope ps(ctx, "rec_walk_count", bytes);
        rec_walk<false, false><<<grid, WALK_WARPS * 32, shm, s>>>(a, q);
    }
    CK(cudaGetLastError());
    return 0;
}

// ---- chain verification + chunk count prefix ------------------------------------------------------------------------------------
struct ChunkScanOp {
    uint64_t n_chunks, first;
    const uint64_t *guess, *exit_;
    const uint32_t *count;
    uint64_t *base;
    uint32_t *bad;
    uint64_t *ctl64;
    __device__ uint64_t n() const { return n_chunks; }
    __device__ void load(uint64_t c, uint64_t (&v)[1]) const
    {
        v[0] = count[c];
        bool ok = exit_[c] != BAD_OFFSET;
        if (c > 0 && exit_[c - 1] != guess[c]) ok = false;
        if (!ok) atomicOr(bad, 1u);
    }
    __device__ void store(uint64_t c, const uint64_t (&excl)[1], const uint64_t (&v)[1]) const
    {
        base[c] = excl[0];
        if (c + 1 == n_chunks) base[c + 1] = excl[0] + v[0];
    }
    __device__ void total(const uint64_t (&t)[1]) const
    {
        ctl64[0] = t[0];
        ctl64[1] = n_chunks ? exit_[n_chunks - 1] : first;
    }
};

void launch_chunk_scan(svb_ctx *ctx, cudaStream_t s, svb_bam *bam, const ScanScratch &sc, uint32_t *ctl_bad, uint64_t *ctl64)
{
    ChunkScanOp op{bam->n_chunks, bam->first, bam->d_guess, bam->d_exit, bam->d_count, bam->d_base, ctl_bad, ctl64};
    launch_scan<1, 8>(ctx, s, op, sc, bam->n_chunks);
}

int accept_counts(svb_ctx *ctx, svb_bam *bam, uint64_t n_rec, uint64_t chain_end)
{
    bam->n_rec = n_rec;
    bam->rec_bytes = (chain_end >= bam->first && chain_end <= bam->nbytes) ? chain_end - bam->first : 0;
    bam->counted = true;
    if (bam->whole_file && bam->rec_bytes != bam->nbytes - bam->first)
        return svb_fail(ctx, SVB_ERR_FORMAT, "truncated or corrupt BAM: record chain ends %llu bytes early",
                        (unsigned long long)(bam->nbytes - bam->first - bam->rec_bytes));
    return 0;
}

// ---- repair of wrong guesses (never seen on well-formed BAMs with the two-record test, but possible in principle) -------------
__device__ __forceinline__ void walk_chunk(const uint8_t *d, uint64_t n, uint64_t entry, uint64_t chunk_end, uint32_t *count,
                                           uint64_t *exit_)
{
    uint64_t o = entry;
    uint32_t k = 0;
    while (o < chunk_end) {
        if (o + 36 > n) break;  // partial tail (shard cut mid-record)
        int32_t bs = ldi32(d + o);
        if (bs < 32) {
            o = BAD_OFFSET;
            break;
        }
        if (o + 4 + (uint64_t)bs > n) break;
        ++k;
        o += 4 + (uint64_t)bs;
    }
    *count = k;
    *exit_ = o;
}
__global__ void verify_chain(uint64_t n_chunks, const uint64_t *__restrict__ guess, const uint64_t *__restrict__ exit_,
                             uint32_t *__restrict__ bad)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    bool ok = exit_[c] != BAD_OFFSET;
    if (c > 0 && exit_[c - 1] != guess[c]) ok = false;
    if (!ok) atomicOr(bad, 1u);
}
// Repair round: every chunk whose guess differs from its predecessor's exit re-walks from that exit. The first
// mismatching chunk always gets its true entry (its predecessor is correct by induction), so repeating
// verify + repair converges; the number of rounds is the longest run of consecutive wrong chunks (1 in practice).
__global__ void __launch_bounds__(128) repair_chain(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t CHUNK_LOG2,
                                                    uint64_t *guess, uint32_t *count, uint64_t *exit_,
                                                    const uint64_t *__restrict__ exit_prev)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 || c >= n_chunks) return;
    uint64_t entry = exit_prev[c - 1];
    if (entry == BAD_OFFSET || guess[c] == entry) return;
    guess[c] = entry;
    uint64_t end = min(n, (c + 1) << CHUNK_LOG2);
    walk_chunk(d, n, entry, end, &count[c], &exit_[c]);
}

int repair_guesses(svb_ctx *ctx, svb_bam *bam)
{
    cudaStream_t s = ctx->stream;
    const uint64_t n_chunks = bam->n_chunks;
    DevBuf<uint32_t> flag;
    DevBuf<uint64_t> snap;
    CK(flag.alloc(1, s));
    CK(snap.alloc(n_chunks, s));
    uint32_t h = 0;
    for (int round = 0;; ++round) {
        CK(cudaMemsetAsync(flag.p, 0, 4, s));
        verify_chain<<<nblk(n_chunks, 256), 256, 0, s>>>(n_chunks, bam->d_guess, bam->d_exit, flag.p);
        CK(cudaMemcpyAsync(&h, flag.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (!h) break;
        // every round settles at least the first wrong chunk of every run, so n_chunks rounds always suffice for a sound chain; runs
        // longer than a few chunks only occur behind records longer than guess_starts' window (> 128 KiB)
        if ((uint64_t)round >= std::min<uint64_t>(n_chunks + 2, 4096)) return svb_fail(ctx, SVB_ERR_FORMAT, "corrupt BAM record chain (block_size < 32)");
        ProfScope ps(ctx, "repair_chain", 0);
        CK(cudaMemcpyAsync(snap.p, bam->d_exit, n_chunks * 8, cudaMemcpyDeviceToDevice, s));
        repair_chain<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, bam->nbytes, n_chunks, bam->chunk_log2, bam->d_guess, bam->d_count,
                                                         bam->d_exit, snap.p);
    }
    bam->counted = false, bam->rows_ready = false, bam->rows_indexed = false;
    return 0;
}

// ---- handle set-up ------------------------------------------------------------------------------------------------------------------
int index_records(svb_ctx *ctx, svb_bam *bam)
{
    uint64_t n = bam->nbytes, first = bam->first;
    if (first > n) return svb_fail(ctx, SVB_ERR_ARG, "first_record beyond the stream");
    {
        // 16 KiB chunks for anything large; a small stream (the unmapped-branch records the shards of a BAM exchange, a test
        // fixture) gets smaller chunks, so that the walk still has tens of thousands of chains to hide its latency behind
        uint32_t l2 = 14;
        while (l2 > 10 && (n >> l2) < 150000) --l2;
        bam->chunk_log2 = l2;
        const char *e = getenv("SEEKSV_B200_CHUNK_LOG2");
        int v = e ? atoi(e) : 0;
        if (v >= 10 && v <= 14) bam->chunk_log2 = (uint32_t)v;
    }
    const uint32_t CHUNK_LOG2 = bam->chunk_log2;
    uint64_t n_chunks = (n + (1ull << CHUNK_LOG2) - 1) >> CHUNK_LOG2;
    if (n_chunks == 0) n_chunks = 1;
    bam->n_chunks = n_chunks;
    cudaStream_t s = ctx->stream;
    // one allocation for the four chunk arrays
    CK(cudaMallocAsync((void **)&bam->d_guess, n_chunks * 8 * 3 + 8 + n_chunks * 4, s));
    bam->d_base = bam->d_guess + n_chunks;
    bam->d_exit = bam->d_base + n_chunks + 1;
    bam->d_count = (uint32_t *)(bam->d_exit + n_chunks);
    return 0;
}

int ensure_guess(svb_ctx *ctx, svb_bam *bam)
{
    if (bam->guessed) return 0;
    cudaStream_t s = ctx->stream;
    {
        ProfScope ps(ctx, "guess_starts", (double)(bam->nbytes - bam->first));
        guess_starts<<<nblk(bam->n_chunks * 32, 256), 256, 0, s>>>(bam->d_data, bam->nbytes, bam->first, bam->n_ref, bam->n_chunks,
                                                                 bam->chunk_log2, bam->d_guess);
    }
    CK(cudaGetLastError());
    bam->guessed = true;
    return 0;
}

int alloc_rows(svb_ctx *ctx, svb_bam *bam, uint32_t R)
{
    if (bam->rows.row && bam->rows.R == R) return 0;
    free_rows(bam);
    const uint64_t slots = bam->n_chunks * (uint64_t)R;
    uint8_t *p = ctx->dev_get(slots * (sizeof(Row) + 2), &bam->rows.cap);
    if (!p) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of device memory for the record rows", (unsigned long long)(slots * 34));
    bam->rows.row = (Row *)p, bam->rows.roff = (uint16_t *)(p + slots * sizeof(Row)), bam->rows.R = R;
    return 0;
}
void free_rows(svb_bam *bam)
{
    if (bam->rows.row) bam->ctx->dev_put((uint8_t *)bam->rows.row, bam->rows.cap);
    bam->rows.row = nullptr, bam->rows.roff = nullptr, bam->rows.R = 0, bam->rows.cap = 0, bam->rows_ready = false, bam->rows_indexed = false;
}

// A stand-alone pass for handles nobody has walked yet (svb_bam_n_records, the getsv passes of a fresh handle): counts, verified
// chain, prefix and - when asked - rows, with ONE read-back.
static int walk_alone(svb_ctx *ctx, svb_bam *bam, bool rows)
{
    cudaStream_t s = ctx->stream;
    CK(cudaSetDevice(ctx->device));
    const uint32_t R_try[2] = {rows_first(bam), rows_max(bam)};
    int r_i = 0;
    for (int attempt = 0;; ++attempt) {
        if (rows) CKR(alloc_rows(ctx, bam, R_try[r_i]));
        Bump measure(nullptr);
        auto carve = [&](Bump &b, ScanScratch &sc, uint32_t *&flags, unsigned long long *&ticket, uint64_t *&ctl64) {
            sc = scan_scratch(b, bam->n_chunks, 1, 8);
            flags = b.get<uint32_t>(4);
            ticket = b.get<unsigned long long>(1);
            ctl64 = b.get<uint64_t>(2);
        };
        ScanScratch sc;
        uint32_t *flags;
        unsigned long long *ticket;
        uint64_t *ctl64;
        carve(measure, sc, flags, ticket, ctl64);
        CKR(ctx->ws_reserve(1, measure.used));
        Bump real(ctx->ws[1]);
        carve(real, sc, flags, ticket, ctl64);
        CK(cudaMemsetAsync(ctx->ws[1], 0, real.used, s));
        ClipQueues none{};
        CKR(launch_walk(ctx, s, bam, false, rows, none, flags, ticket));
        launch_chunk_scan(ctx, s, bam, sc, flags + 1, ctl64);
        struct {
            uint32_t flags[4];
            uint64_t c64[2];
        } h;
        CK(cudaMemcpyAsync(h.flags, flags, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h.c64, ctl64, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        if (h.flags[1]) {  // a guess was wrong: repair and walk again
            if (attempt >= 3) return svb_fail(ctx, SVB_ERR_FORMAT, "record chain does not verify");
            CKR(repair_guesses(ctx, bam));
            continue;
        }
        if (rows && h.flags[0]) {  // more records in a chunk than row slots: once more with the largest possible slot count
            if (r_i == 1) return svb_fail(ctx, SVB_ERR_FORMAT, "a chunk holds more than %u records", R_try[1]);
            r_i = 1;
            continue;
        }
        CKR(accept_counts(ctx, bam, h.c64[0], h.c64[1]));
        if (rows) bam->rows_ready = true, bam->rows_indexed = false;
        return 0;
    }
}

int ensure_counts(svb_ctx *ctx, svb_bam *bam)
{
    if (bam->counted) return 0;
    return walk_alone(ctx, bam, false);
}
int ensure_rows(svb_ctx *ctx, svb_bam *bam)
{
    if (bam->rows_ready && bam->counted) return 0;
    return walk_alone(ctx, bam, true);
}
