// getsv / somatic device passes over the lean record columns: insert-size statistics
// (CalculateInsertsizeDeviation, cluster.cpp:15-83), discordant-pair support per junction
// (FindDiscordantReadPairs, getsv.cpp:990-1247 with IsConcordant cluster.cpp:136-147) and per-position
// depth inside merged junction windows (main_depth, bam2depth.cpp:17-142, libbam pileup semantics as
// probed in tests/test_oracle_golden.py). One decode pass streams the packed records once
// (decode_records); everything after it reads ~32 B/record of columns instead of ~320 B of record.
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "stream.cuh"

static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

#define FLAGQ_NOCIGAR (1u << 25)
#define PILEUP_MAXCNT 8000

// ---- decode: packed records -> lean columns -----------------------------------------------------------------
__device__ __forceinline__ bool insert_qualifies(uint32_t fq, int32_t isize, int32_t min_mapq)
{
    uint32_t flag = fq & 0xffff;
    if ((int32_t)((fq >> 16) & 0xff) < min_mapq) return false;  // __g_skip_aln in cluster.cpp's TU (quirk Q9)
    if (fq & FLAGQ_HARDCLIP) return false;
    return (flag & F_PAIRED) && (flag & F_PROPER) && !(flag & F_DUP) && isize > 0;
}

// The getsv walker: one thread per 16 KiB chunk follows the verified record chain (guess + per-chunk record base from
// ensure_counts) and writes the lean columns of its records; every record head is fetched once. On the way it gathers
// the chunk's partial sums for CalculateInsertsizeDeviation (cluster.cpp:48-70) at one mapQ threshold.
__global__ void __launch_bounds__(128)
    decode_walk(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t CHUNK_LOG2, const uint64_t *__restrict__ guess,
                const uint64_t *__restrict__ base, LeanRecords L, int32_t stats_mapq, uint32_t *__restrict__ q_cnt,
                uint64_t *__restrict__ q_sum, uint64_t *__restrict__ q_sq, int32_t *__restrict__ scal /* max_span, unsorted, q_max */)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t span = 0, qmax = 0;
    uint32_t unsorted = 0;
    if (c < n_chunks) {
        uint64_t o = guess[c], end = min(n, (c + 1) << CHUNK_LOG2), i = base[c];
        uint32_t qc = 0, pt = 0;
        uint64_t qs = 0, qq = 0;
        int32_t pp = 0;
        bool have_prev = false;
        bool live = o < end && o + 36 <= n;
        Core k;
        if (live) k = load_core(d + o);
        while (live) {
            if (k.block_size < 32 || o + 4 + (uint64_t)k.block_size > n) break;  // (the chain is verified: tail only)
            // software pipeline: the next record's fixed part is requested before this record's CIGAR is waited for
            uint64_t on = o + 4 + (uint64_t)k.block_size;
            bool next_live = on < end && on + 36 <= n;
            Core kn;
            if (next_live) kn = load_core(d + on);
            const uint8_t *cig = d + o + 36 + k.l_qname;
            int32_t rend = k.pos;
            uint32_t fq = k.flag | (k.mapq << 16);
            if (k.n_cigar == 0) fq |= FLAGQ_NOCIGAR;
            for (uint32_t j = 0; j < k.n_cigar; ++j) {
                uint32_t w = ldu32(cig + 4 * j), op = w & 15;
                // bam_calend of the linked libbam: M, D, N only ('=' and 'X' do not advance; probed)
                if (op == OP_M || op == OP_D || op == OP_N) rend += (int32_t)(w >> 4);
                if ((j == 0 || j + 1 == k.n_cigar) && op == OP_H) fq |= FLAGQ_HARDCLIP;  // IsHardClip, clip_reads.cpp:247
            }
            {
                uint4 *row = (uint4 *)&L.rec[i];
                row[0] = make_uint4((uint32_t)k.tid, (uint32_t)k.pos, (uint32_t)rend, fq);
                row[1] = make_uint4((uint32_t)k.l_qseq, (uint32_t)k.mtid, (uint32_t)k.mpos, (uint32_t)k.isize);
                row[2] = make_uint4((uint32_t)o, (uint32_t)(o >> 32), 0u, 0u);
            }
            span = max(span, max(rend - k.pos, 1));
            if (have_prev && (pt > (uint32_t)k.tid || (pt == (uint32_t)k.tid && pp > k.pos))) unsorted = 1;  // tid -1 sorts last
            pt = (uint32_t)k.tid, pp = k.pos, have_prev = true;
            if (stats_mapq >= 0 && insert_qualifies(fq, k.isize, stats_mapq)) {
                ++qc, qs += (uint64_t)k.isize, qq += (uint64_t)k.isize * (uint64_t)k.isize;
                qmax = max(qmax, k.isize);
            }
            ++i;
            o = on, k = kn, live = next_live;
        }
        q_cnt[c] = qc, q_sum[c] = qs, q_sq[c] = qq;
    }
    span = (int32_t)warp_max((uint32_t)span);
    qmax = (int32_t)warp_max((uint32_t)qmax);
    unsorted = warp_max(unsorted);
    if ((threadIdx.x & 31) == 0) {
        if (span > 0) atomicMax(&scal[0], span);
        if (unsorted) atomicOr((uint32_t *)&scal[1], 1u);
        if (qmax > 0) atomicMax(&scal[2], qmax);
    }
}

// The getsv full pass, streaming form (stream.cuh). Each CTA stages a tile with TMA, indexes it in shared memory, gets the
// global index of the tile's first record from a single-pass chained prefix over the tile record counts (decoupled
// look-back: a tile publishes its count, then sums its predecessors' counts until it meets a published running total), and
// writes one 48-byte row per thread - consecutive threads, consecutive rows: fully coalesced stores.
static constexpr uint64_t ST_AGG = 1ull << 62, ST_INC = 2ull << 62, ST_MASK = (1ull << 62) - 1;

__global__ void __launch_bounds__(STREAM_THREADS)
    decode_stream(const uint8_t *__restrict__ d, uint64_t n, uint64_t padded, uint64_t first, int32_t n_ref, uint64_t n_tiles,
                  uint64_t *__restrict__ guess, uint32_t *__restrict__ count, uint64_t *__restrict__ exit_, uint64_t *__restrict__ base_out,
                  unsigned long long *__restrict__ state, LeanRecords L, uint64_t row_cap, int32_t stats_mapq,
                  uint32_t *__restrict__ q_cnt, uint64_t *__restrict__ q_sum, uint64_t *__restrict__ q_sq,
                  int32_t *__restrict__ scal /* max_span, unsorted, q_max, overflow */)
{
    __shared__ StreamShared S;
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&S.full[s], 1);
        fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < STAGES; ++s) {
            uint64_t t = blockIdx.x + (uint64_t)s * gridDim.x;
            if (t < n_tiles) issue_tile(S, s, d, padded, t);
        }
    uint32_t it = 0;
    int32_t span = 0, qmax = 0;
    uint32_t unsorted = 0;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        mbar_wait(&S.full[s], (it / STAGES) & 1);
        const uint64_t tile_abs = t << TILE_LOG2;
        TileWin w{S.stage[s], d + tile_abs, (uint32_t)min((uint64_t)(TILE + HALO), padded - tile_abs)};
        index_tile(S, w, t, n, first, n_ref, d);
        const uint32_t n_rec = S.n_rec;
        if (tid == 0) {
            uint64_t base = 0;
            if (t > 0) {
                atomicExch(&state[t], ST_AGG | n_rec);
                uint64_t p = t - 1, sum = 0;
                for (;;) {
                    unsigned long long v = *(volatile unsigned long long *)&state[p];
                    if ((v >> 62) == 0) continue;  // predecessor not there yet
                    sum += v & ST_MASK;
                    if ((v >> 62) == 2) break;
                    --p;
                }
                base = sum;
            }
            atomicExch(&state[t], ST_INC | (base + n_rec));
            S.base = base;
            S.red[0] = S.red[1] = S.red[2] = 0;
        }
        __syncthreads();
        const uint64_t base = S.base;
        for (uint32_t kb = 0; kb < n_rec; kb += STREAM_THREADS) {
            const uint32_t k = kb + tid;
            uint32_t qc = 0;
            uint64_t qs = 0, qq = 0;
            if (k < n_rec) {
                const uint32_t off = S.rec_off[k];
                const Core c = w.core(off);
                const uint32_t cg = off + 36 + c.l_qname;
                int32_t rend = c.pos;
                uint32_t fq = c.flag | (c.mapq << 16);
                if (c.n_cigar == 0) fq |= FLAGQ_NOCIGAR;
                for (uint32_t j = 0; j < c.n_cigar; ++j) {
                    uint32_t x = w.u32(cg + 4 * j), op = x & 15;
                    // bam_calend of the linked libbam: M, D, N only ('=' and 'X' do not advance; probed)
                    if (op == OP_M || op == OP_D || op == OP_N) rend += (int32_t)(x >> 4);
                    if ((j == 0 || j + 1 == c.n_cigar) && op == OP_H) fq |= FLAGQ_HARDCLIP;  // IsHardClip, clip_reads.cpp:247
                }
                const uint64_t o = tile_abs + off, row = base + k;
                if (row < row_cap) {
                    uint4 *r = (uint4 *)&L.rec[row];
                    r[0] = make_uint4((uint32_t)c.tid, (uint32_t)c.pos, (uint32_t)rend, fq);
                    r[1] = make_uint4((uint32_t)c.l_qseq, (uint32_t)c.mtid, (uint32_t)c.mpos, (uint32_t)c.isize);
                    r[2] = make_uint4((uint32_t)o, (uint32_t)(o >> 32), 0u, 0u);
                } else
                    scal[3] = 1;
                span = max(span, max(rend - c.pos, 1));
                if (k > 0) {  // coordinate order inside the tile (tile boundaries: boundary_order)
                    uint32_t op_ = S.rec_off[k - 1];
                    uint32_t pt = w.u32(op_ + 4);
                    int32_t pp = (int32_t)w.u32(op_ + 8);
                    if (pt > (uint32_t)c.tid || (pt == (uint32_t)c.tid && pp > c.pos)) unsorted = 1;  // tid -1 sorts last
                }
                if (stats_mapq >= 0 && insert_qualifies(fq, c.isize, stats_mapq)) {
                    qc = 1, qs = (uint64_t)c.isize, qq = (uint64_t)c.isize * (uint64_t)c.isize;
                    qmax = max(qmax, c.isize);
                }
            }
            if (stats_mapq >= 0 && __any_sync(0xffffffffu, qc != 0)) {
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) {
                    qc += __shfl_xor_sync(0xffffffffu, qc, sft);
                    qs += __shfl_xor_sync(0xffffffffu, qs, sft);
                    qq += __shfl_xor_sync(0xffffffffu, qq, sft);
                }
                if ((tid & 31) == 0) {
                    atomicAdd(&S.red[0], (unsigned long long)qc);
                    atomicAdd(&S.red[1], (unsigned long long)qs);
                    atomicAdd(&S.red[2], (unsigned long long)qq);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            count[t] = n_rec, exit_[t] = S.exit_, guess[t] = S.entry, base_out[t] = base;
            q_cnt[t] = (uint32_t)S.red[0], q_sum[t] = S.red[1], q_sq[t] = S.red[2];
            uint64_t tn = t + (uint64_t)STAGES * gridDim.x;
            if (tn < n_tiles) {
                fence_proxy_async();
                issue_tile(S, s, d, padded, tn);
            }
        }
        __syncthreads();
    }
    span = (int32_t)warp_max((uint32_t)span);
    qmax = (int32_t)warp_max((uint32_t)qmax);
    unsorted = warp_max(unsorted);
    if ((tid & 31) == 0) {
        if (span > 0) atomicMax(&scal[0], span);
        if (unsorted) atomicOr((uint32_t *)&scal[1], 1u);
        if (qmax > 0) atomicMax(&scal[2], qmax);
    }
}

// coordinate order across chunk boundaries: the last record of a chunk against the first of the next
__global__ void boundary_order(uint64_t n_chunks, const uint64_t *__restrict__ base, LeanRecords L, uint32_t *__restrict__ unsorted)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 || c >= n_chunks) return;
    uint64_t i = base[c];
    if (i == 0 || i >= L.n || base[c + 1] == i) return;
    uint32_t t0 = (uint32_t)L.rec[i - 1].tid, t1 = (uint32_t)L.rec[i].tid;
    if (t0 > t1 || (t0 == t1 && L.rec[i - 1].pos > L.rec[i].pos)) atomicOr(unsorted, 1u);
}

static int free_lean(svb_ctx *ctx, svb_bam *bam)
{
    cudaStream_t s = ctx->stream;
    void *p[4] = {bam->lean.rec, bam->d_q_cnt, bam->d_q_sum, bam->d_q_sq};
    for (void *x : p)
        if (x) cudaFreeAsync(x, s);
    bam->lean.rec = nullptr, bam->d_q_cnt = nullptr, bam->d_q_sum = nullptr, bam->d_q_sq = nullptr;
    return 0;
}

int decode_records(svb_ctx *ctx, svb_bam *bam, int32_t stats_mapq)
{
    if (bam->lean_ready) return 0;
    cudaStream_t s = ctx->stream;
    const uint64_t n_chunks = bam->n_chunks, stream_bytes = bam->nbytes - bam->first;
    LeanRecords &L = bam->lean;
    DevBuf<int32_t> scal;
    CK(scal.alloc(4, s));
    int32_t h[4] = {0, 0, 0, 0};
    bool done = false;
    if (stream_mode(bam)) {
        // one streaming pass: rows are allocated for "a record is at least 96 bytes" (a 50-base read is ~120), the exact
        // count comes out of the pass; shorter records overflow the estimate and the walker path below redoes the pass
        uint64_t cap = stream_bytes / 96 + 1024;
        CK(cudaMallocAsync((void **)&L.rec, cap * sizeof(LeanRec), s));
        CK(cudaMallocAsync((void **)&bam->d_q_cnt, n_chunks * 4, s));
        CK(cudaMallocAsync((void **)&bam->d_q_sum, n_chunks * 8, s));
        CK(cudaMallocAsync((void **)&bam->d_q_sq, n_chunks * 8, s));
        DevBuf<unsigned long long> state;
        DevBuf<uint64_t> exit_;
        CK(state.alloc(n_chunks, s));
        CK(exit_.alloc(n_chunks, s));
        CK(cudaMemsetAsync(state.p, 0, n_chunks * 8, s));
        CK(cudaMemsetAsync(scal.p, 0, 16, s));
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, decode_stream, STREAM_THREADS, 0));
        unsigned grid = (unsigned)std::min<uint64_t>(n_chunks, (uint64_t)std::max(1, per_sm) * ctx->sm_count);
        {
            ProfScope ps(ctx, "decode_stream", (double)stream_bytes);
            decode_stream<<<grid, STREAM_THREADS, 0, s>>>(bam->d_data, bam->nbytes, (bam->nbytes + 15) & ~15ull, bam->first, bam->n_ref,
                                                          n_chunks, bam->d_guess, bam->d_count, exit_.p, bam->d_base, state.p, L, cap,
                                                          stats_mapq, bam->d_q_cnt, bam->d_q_sum, bam->d_q_sq, scal.p);
        }
        bam->guessed = true;
        int ok = 0;
        CKR(verify_or_repair(ctx, bam, exit_.p, &ok));
        CK(cudaMemcpyAsync(h, scal.p, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (ok && !h[3]) {
            CKR(finish_counts(ctx, bam, exit_.p));  // prefix of the per-tile counts, totals, end-of-stream check
            L.n = bam->n_rec;
            if (L.n >= (1ull << 32)) return svb_fail(ctx, SVB_ERR_ARG, "more than 2^32 records in one shard");
            done = true;
        } else
            free_lean(ctx, bam);
    }
    if (!done) {
        CKR(ensure_counts(ctx, bam));
        uint64_t n = bam->n_rec;
        if (n >= (1ull << 32)) return svb_fail(ctx, SVB_ERR_ARG, "more than 2^32 records in one shard");
        size_t cnt = n ? n : 1;
        CK(cudaMallocAsync((void **)&L.rec, cnt * sizeof(LeanRec), s));
        CK(cudaMallocAsync((void **)&bam->d_q_cnt, n_chunks * 4, s));
        CK(cudaMallocAsync((void **)&bam->d_q_sum, n_chunks * 8, s));
        CK(cudaMallocAsync((void **)&bam->d_q_sq, n_chunks * 8, s));
        L.n = n;
        CK(cudaMemsetAsync(scal.p, 0, 16, s));
        {
            ProfScope ps(ctx, "decode_walk", (double)bam->rec_bytes);
            decode_walk<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, bam->nbytes, n_chunks, bam->chunk_log2, bam->d_guess, bam->d_base, L,
                                                          stats_mapq, bam->d_q_cnt, bam->d_q_sum, bam->d_q_sq, scal.p);
        }
        CK(cudaMemcpyAsync(h, scal.p, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    {
        DevBuf<uint32_t> uns;
        CK(uns.alloc(1, s));
        CK(cudaMemsetAsync(uns.p, 0, 4, s));
        boundary_order<<<nblk(n_chunks, 256), 256, 0, s>>>(n_chunks, bam->d_base, L, uns.p);
        uint32_t hu = 0;
        CK(cudaMemcpyAsync(&hu, uns.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (hu) h[1] = 1;
    }
    CK(cudaGetLastError());
    bam->max_span = h[0];
    bam->sorted = h[1] ? 0 : 1;
    bam->q_max = h[2];
    bam->stats_mapq = stats_mapq;
    bam->lean_ready = true;
    return 0;
}

// ---- insert size ------------------------------------------------------------------------------------------------
__global__ void insert_flags(uint64_t n, const LeanRec *__restrict__ rec, int32_t min_mapq, uint32_t *__restrict__ flag)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = insert_qualifies(rec[i].flagq, rec[i].isize, min_mapq) ? 1u : 0u;
}
// pass 1 (mean == INT_MIN): sum of isize; pass 2: sum of (int32)((isize-mean)*(isize-mean))
__global__ void __launch_bounds__(256)
    insert_sums(uint64_t n, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank, const LeanRec *__restrict__ rec,
                uint64_t max_pairs, int pass, int32_t mean, unsigned long long *__restrict__ acc)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long v = 0;
    if (i < n && flag[i] && (uint64_t)rank[i] <= max_pairs) {
        if (pass == 1) v = rec[i].isize;
        else {
            uint32_t dlt = (uint32_t)(rec[i].isize - mean);
            v = (int32_t)(dlt * dlt);  // the reference multiplies two ints (cluster.cpp:77)
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(acc, (unsigned long long)v);
}

__global__ void chunk_totals(uint64_t n_chunks, const uint32_t *__restrict__ q_cnt, const uint64_t *__restrict__ q_sum,
                             const uint64_t *__restrict__ q_sq, unsigned long long *__restrict__ tot)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long a = 0, b = 0, q = 0;
    if (c < n_chunks) a = q_cnt[c], b = q_sum[c], q = q_sq[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0 && a) {
        atomicAdd(&tot[0], a);
        atomicAdd(&tot[1], b);
        atomicAdd(&tot[2], q);
    }
}

extern "C" int svb_insert_stats(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t max_pairs, int64_t out[4])
{
    if (!ctx || !bam || !out) return svb_fail(ctx, SVB_ERR_ARG, "svb_insert_stats: null argument");
    CKR(decode_records(ctx, bam, min_mapq));  // the decode walker gathers per-chunk partial sums for this mapQ on its way
    cudaStream_t s = ctx->stream;
    uint64_t n = bam->n_rec;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (n == 0 || max_pairs <= 0) return 0;
    if (n >= (1ull << 32)) return svb_fail(ctx, SVB_ERR_ARG, "more than 2^32 records in one shard");
    if (bam->stats_mapq == min_mapq && bam->q_max <= 46340) {
        // Fast path: every |isize - mean| stays below sqrt(2^31), so the reference's int products cannot wrap and
        // sum (x - m)^2 = sum x^2 - 2 m sum x + n m^2 holds exactly in 64-bit integers. Valid when all qualifying records
        // are used (total <= -n); otherwise the ordered cut-off needs the per-record path below.
        DevBuf<unsigned long long> tot;
        CK(tot.alloc(3, s));
        CK(cudaMemsetAsync(tot.p, 0, 24, s));
        {
            ProfScope ps(ctx, "insert_stats", (double)bam->n_chunks * 20);
            chunk_totals<<<nblk(bam->n_chunks, 256), 256, 0, s>>>(bam->n_chunks, bam->d_q_cnt, bam->d_q_sum, bam->d_q_sq, tot.p);
        }
        unsigned long long h[3];
        CK(cudaMemcpyAsync(h, tot.p, 24, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (h[0] <= (unsigned long long)max_pairs) {
            if (h[0] == 0) return 0;
            long long mean = (long long)(h[1] / h[0]);
            out[0] = (int64_t)h[0], out[1] = (int64_t)h[1], out[2] = mean;
            out[3] = (int64_t)h[2] - 2 * mean * (int64_t)h[1] + (int64_t)h[0] * mean * mean;
            return 0;
        }
    }
    DevBuf<uint32_t> flag, rank;
    DevBuf<unsigned long long> acc;
    CK(flag.alloc(n, s));
    CK(rank.alloc(n, s));
    CK(acc.alloc(2, s));
    CK(cudaMemsetAsync(acc.p, 0, 16, s));
    {
        ProfScope ps(ctx, "insert_stats", (double)n * 16);
        insert_flags<<<nblk(n, 256), 256, 0, s>>>(n, bam->lean.rec, min_mapq, flag.p);
        CKR(inclusive_scan_u32(ctx, flag.p, rank.p, n));
        insert_sums<<<nblk(n, 256), 256, 0, s>>>(n, flag.p, rank.p, bam->lean.rec, (uint64_t)max_pairs, 1, 0, acc.p);
    }
    uint32_t total = 0;
    unsigned long long sum = 0;
    CK(cudaMemcpyAsync(&total, rank.p + (n - 1), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&sum, acc.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    uint64_t cnt = std::min<uint64_t>(total, (uint64_t)max_pairs);
    if (cnt == 0) return 0;
    int32_t mean = (int32_t)(sum / cnt);  // unsigned long / int, stored to int (cluster.cpp:72)
    {
        ProfScope ps(ctx, "insert_stats", (double)n * 12);
        insert_sums<<<nblk(n, 256), 256, 0, s>>>(n, flag.p, rank.p, bam->lean.rec, (uint64_t)max_pairs, 2, mean, acc.p + 1);
    }
    long long sq = 0;
    CK(cudaMemcpyAsync(&sq, acc.p + 1, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    out[0] = (int64_t)cnt, out[1] = (int64_t)sum, out[2] = mean, out[3] = sq;
    return 0;
}

// ---- discordant read pairs ------------------------------------------------------------------------------------------
// first record index with (tid, pos) >= (T, P); tid -1 sorts last
__device__ __forceinline__ uint64_t lower_bound_tp(const LeanRec *__restrict__ rec, uint64_t n, int32_t T, int64_t P)
{
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t m = (lo + hi) >> 1;
        uint32_t t = (uint32_t)rec[m].tid;
        bool less = t < (uint32_t)T || (t == (uint32_t)T && (int64_t)rec[m].pos < P);
        if (less) lo = m + 1;
        else hi = m;
    }
    return lo;
}

__global__ void __launch_bounds__(128)
    discordant_kernel(LeanRecords L, int32_t max_span, const svb_junction *__restrict__ J, uint64_t n_j,
                      const uint32_t *__restrict__ ref_len, svb_pair_params prm, int32_t *__restrict__ counts)
{
    uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (w >= n_j) return;
    const svb_junction j = J[w];
    const int kCross = 5;  // kCrossLength, getsv.cpp:15
    int32_t min_is = prm.mean_insert - prm.deviation * prm.times, max_is = prm.mean_insert + prm.deviation * prm.times;
    if (min_is < 0) min_is = 0;
    int32_t tid = j.up_tid, mtid = j.down_tid;
    uint32_t n = 0;
    if (tid >= 0 && (j.up_strand == '+' || j.up_strand == '-')) {
        int32_t beg, end;
        if (j.up_strand == '+') end = j.up_pos, beg = end - max_is;
        else beg = j.up_pos - 1 - kCross, end = j.up_pos - 1 + max_is;
        if (beg <= 0) beg = 1;
        if ((uint32_t)end > ref_len[tid]) end = (int32_t)ref_len[tid];  // int vs unsigned compare, getsv.cpp:1060
        // bam_iter_query(idx, tid, beg, end): records on tid with pos < end and calend > beg
        uint64_t lo = lower_bound_tp(L.rec, L.n, tid, (int64_t)beg - max_span);
        uint64_t hi = lower_bound_tp(L.rec, L.n, tid, end);
        for (uint64_t i = lo + lane; i < hi; i += 32) {
            const LeanRec r = L.rec[i];
            uint32_t fq = r.flagq, flag = fq & 0xffff;
            int32_t pos = r.pos;
            int32_t rend = (fq & FLAGQ_NOCIGAR) ? pos + 1 : r.end;
            if (!(rend > beg)) continue;
            if ((int32_t)((fq >> 16) & 0xff) < prm.min_mapq) continue;  // __g_skip_aln, getsv.cpp:1027,1069
            if (fq & FLAGQ_HARDCLIP) continue;
            if (flag & (F_DUP | F_UNMAP | F_MUNMAP)) continue;
            int32_t isz = r.isize;
            bool rev = flag & F_REVERSE, mrev = flag & F_MREVERSE;
            {  // IsConcordant, cluster.cpp:136-147 (its own, unclamped minimum)
                int32_t lo_c = prm.mean_insert - prm.deviation * prm.times;
                bool conc = false;
                if (!rev && mrev && lo_c <= isz && isz <= max_is) conc = true;
                else if (rev && !mrev && isz < 0) {
                    int32_t a = isz < 0 ? -isz : isz;
                    conc = lo_c <= a && a <= max_is;
                }
                if (conc) continue;
            }
            if (mtid == -1 || mtid != r.mtid) continue;
            int32_t lq = r.lqseq, mpos = r.mpos;
            bool hit = false;
            if (j.up_strand == '+' && j.down_strand == '+' && pos + lq <= j.up_pos + kCross && mpos + 1 >= j.down_pos - kCross) {
                if (!rev && mrev) {
                    int32_t isize = j.up_pos - pos + mpos + lq - j.down_pos + 1;
                    if (tid == mtid && j.up_pos > j.down_pos && j.up_pos - j.down_pos + 1 + 2 * lq <= max_is) {
                        while (isize <= max_is) {  // tandem duplication: add whole copies (getsv.cpp:1081-1091)
                            if (isize >= min_is) {
                                hit = true;
                                break;
                            }
                            isize += j.up_pos - j.down_pos + 1;
                        }
                    } else
                        hit = min_is <= isize && isize <= max_is;
                }
            } else if (j.up_strand == '-' && j.down_strand == '+' && rev && mrev && mpos + 1 >= j.down_pos - kCross) {
                int32_t isize = pos + 1 - j.up_pos + 1 + mpos + lq - j.down_pos + 1;
                hit = min_is <= isize && isize <= max_is;
            } else if (j.up_strand == '+' && j.down_strand == '-' && !rev && !mrev && pos + lq <= j.up_pos + kCross &&
                       mpos + lq <= j.down_pos + kCross) {
                int32_t isize = j.up_pos - pos + j.down_pos - (mpos + lq) + 1;
                hit = min_is <= isize && isize <= max_is;
            }
            n += hit;
        }
    }
    n = warp_sum(n);
    if (lane == 0) counts[w] = (int32_t)n;
}

extern "C" int svb_discordant_support(svb_ctx *ctx, svb_bam *bam, const svb_junction *junctions, uint64_t n,
                                      const svb_pair_params *p, int32_t *counts)
{
    if (!ctx || !bam || !p || (n && (!junctions || !counts))) return svb_fail(ctx, SVB_ERR_ARG, "svb_discordant_support: null argument");
    CKR(decode_records(ctx, bam));
    if (n == 0) return 0;
    if (bam->sorted != 1) return svb_fail(ctx, SVB_ERR_UNSORTED, "the BAM is not coordinate-sorted");
    if (bam->lens.size() != (size_t)bam->n_ref) return svb_fail(ctx, SVB_ERR_ARG, "reference lengths not set (svb_bam_set_refs)");
    cudaStream_t s = ctx->stream;
    DevBuf<svb_junction> dj;
    DevBuf<int32_t> dc;
    DevBuf<uint32_t> dl;
    CK(dj.alloc(n, s));
    CK(dc.alloc(n, s));
    CK(dl.alloc(bam->n_ref, s));
    CK(cudaMemcpyAsync(dj.p, junctions, n * sizeof(svb_junction), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dl.p, bam->lens.data(), (size_t)bam->n_ref * 4, cudaMemcpyHostToDevice, s));
    {
        ProfScope ps(ctx, "discordant_support", 0);
        discordant_kernel<<<nblk(n * 32, 128), 128, 0, s>>>(bam->lean, bam->max_span, dj.p, n, dl.p, *p, dc.p);
    }
    CK(cudaMemcpyAsync(counts, dc.p, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return 0;
}

// ---- window depth ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pileup_eligible(int32_t tid, uint32_t fq, int32_t min_mapq)
{
    // bam_plp_push: tid >= 0 and (flag & 0x704) == 0, after read_bam (bam2depth.h:29-35) set UNMAP for low mapQ
    return tid >= 0 && !((fq & 0xffff) & 0x704) && (int32_t)((fq >> 16) & 0xff) >= min_mapq;
}

__global__ void eligible_flags(LeanRecords L, int32_t min_mapq, uint32_t *__restrict__ flag)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < L.n) flag[i] = pileup_eligible(L.rec[i].tid, L.rec[i].flagq, min_mapq) ? 1u : 0u;
}

// Exact serial emulation of bam_plp_push's cap for one chromosome (one thread per hot chromosome): a read is
// refused only if it starts at the same position as the previously accepted read while more than 8000
// buffer nodes are allocated = 2 + accepted reads whose end >= that position (released lazily).
__global__ void cap_serial(LeanRecords L, const uint32_t *__restrict__ hot_tids, uint32_t n_ref, const uint32_t *__restrict__ flag,
                           uint8_t *__restrict__ kept, uint32_t *__restrict__ ring_all, uint32_t ring)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ref || !hot_tids[t]) return;
    uint32_t *hist = ring_all + (uint64_t)t * ring;  // live accepted reads by end position (mod ring)
    for (uint32_t k = 0; k < ring; ++k) hist[k] = 0;
    uint64_t lo = lower_bound_tp(L.rec, L.n, (int32_t)t, INT32_MIN);
    uint64_t hi = lower_bound_tp(L.rec, L.n, (int32_t)t + 1, INT32_MIN);
    int64_t it_pos = -1;  // position of the previously accepted read (iterator position)
    uint32_t live = 0;
    bool any = false;
    for (uint64_t i = lo; i < hi; ++i) {
        if (!flag[i]) continue;
        int32_t pos = L.rec[i].pos, end = L.rec[i].end;
        if (any && pos == it_pos) {
            if (live + 2 > PILEUP_MAXCNT) {
                kept[i] = 0;
                continue;
            }
            if (end > pos) hist[(uint32_t)end % ring]++, live++;
        } else {
            if (any) {  // release nodes whose end <= pos - 1
                int64_t from = it_pos, to = pos;  // ends in [from, to) leave; ends < from left earlier
                if (to - from >= ring) {
                    for (uint32_t k = 0; k < ring; ++k) hist[k] = 0;
                    live = 0;
                } else
                    for (int64_t e = from; e < to; ++e) {
                        uint32_t &h = hist[(uint32_t)e % ring];
                        live -= h, h = 0;
                    }
            }
            it_pos = pos, any = true;
            hist[(uint32_t)end % ring]++, live++;
        }
    }
}

__device__ __forceinline__ uint64_t first_window(const svb_window *__restrict__ W, uint64_t n, int32_t tid, int32_t p)
{
    // first window with (tid, end) >= (tid, p); windows are disjoint and sorted, so ends are sorted too
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t m = (lo + hi) >> 1;
        bool less = W[m].tid < tid || (W[m].tid == tid && W[m].end < p);
        if (less) lo = m + 1;
        else hi = m;
    }
    return lo;
}

// One thread per record: +1/-1 marks of its M segments into the difference array of every window it overlaps.
// kept == nullptr: every pileup-eligible read is kept and the kernel also tests libbam's 8000-read cap bound: if the record
// 7998 places earlier on the same chromosome starts within max_span of this one, more than 8000 buffer nodes are
// possible and the chromosome is flagged for the exact serial emulation (cap_serial), after which the marks are redone.
__global__ void __launch_bounds__(256)
    depth_marks(const uint8_t *__restrict__ d, LeanRecords L, int32_t min_mapq, int32_t max_span, const uint8_t *__restrict__ kept,
                uint32_t *__restrict__ hot_tids, const svb_window *__restrict__ W, const uint64_t *__restrict__ woff, uint64_t n_w,
                int32_t *__restrict__ diff)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L.n) return;
    const LeanRec r = L.rec[i];
    int32_t tid = r.tid;
    if (!pileup_eligible(tid, r.flagq, min_mapq)) return;
    if (kept) {
        if (!kept[i]) return;
    } else if (i >= PILEUP_MAXCNT - 2) {
        uint64_t j = i - (PILEUP_MAXCNT - 2);
        if (L.rec[j].tid == tid && L.rec[j].pos >= r.pos - max_span) hot_tids[tid] = 1;
    }
    int32_t beg1 = r.pos + 1, end1 = r.end;  // 1-based inclusive [beg1, end1]
    if (end1 < beg1) return;
    uint64_t w = first_window(W, n_w, tid, beg1);
    if (w >= n_w || W[w].tid != tid || W[w].begin > end1) return;
    const uint8_t *p = d + r.off;
    uint32_t lq = ldu32(p + 12) & 0xff, nc = ldu32(p + 16) & 0xffff;
    const uint8_t *cig = p + 36 + lq;
    for (; w < n_w && W[w].tid == tid && W[w].begin <= end1; ++w) {
        int32_t wb = W[w].begin, we = W[w].end;
        int32_t *dw = diff + woff[w];
        int32_t x = beg1;
        for (uint32_t j = 0; j < nc && x <= we; ++j) {
            uint32_t c = ldu32(cig + 4 * j), op = c & 15;
            int32_t len = (int32_t)(c >> 4);
            if (op == OP_M) {  // '=' / 'X' are ignored by this libbam's CIGAR walk (probed)
                int32_t lo = max(x, wb), hi = min(x + len - 1, we);
                if (lo <= hi) {
                    atomicAdd(&dw[lo - wb], 1);
                    atomicAdd(&dw[hi + 1 - wb], -1);
                }
                x += len;
            } else if (op == OP_D || op == OP_N)
                x += len;
        }
    }
}

// prefix sum inside each window (one warp per window), diff -> depth, compacted to the output layout
__global__ void __launch_bounds__(128)
    depth_scan(const svb_window *__restrict__ W, const uint64_t *__restrict__ woff, uint64_t n_w, const int32_t *__restrict__ diff,
               int32_t *__restrict__ depth)
{
    uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (w >= n_w) return;
    uint32_t len = (uint32_t)(W[w].end - W[w].begin + 1);
    const int32_t *dw = diff + woff[w];
    int32_t *out = depth + (woff[w] - w);  // diff has one extra slot per window
    int32_t carry = 0;
    for (uint32_t base = 0; base < len; base += 32) {
        uint32_t k = base + lane;
        int32_t v = k < len ? dw[k] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= (uint32_t)o) v += t;
        }
        v += carry;
        if (k < len) out[k] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

__global__ void fill_u8(uint64_t n, const uint32_t *__restrict__ flag, uint8_t *__restrict__ kept)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) kept[i] = (uint8_t)flag[i];
}

extern "C" int svb_window_depth(svb_ctx *ctx, svb_bam *bam, const svb_window *windows, uint64_t n_w, int32_t min_mapq,
                                int32_t *depth_out)
{
    if (!ctx || !bam || (n_w && (!windows || !depth_out))) return svb_fail(ctx, SVB_ERR_ARG, "svb_window_depth: null argument");
    CKR(decode_records(ctx, bam));
    if (n_w == 0) return 0;
    if (bam->sorted != 1) return svb_fail(ctx, SVB_ERR_UNSORTED, "the BAM is not coordinate-sorted");
    cudaStream_t s = ctx->stream;
    uint64_t n = bam->n_rec;
    std::vector<uint64_t> woff(n_w + 1);
    uint64_t tot = 0;
    for (uint64_t w = 0; w < n_w; ++w) {
        if (windows[w].end < windows[w].begin) return svb_fail(ctx, SVB_ERR_ARG, "svb_window_depth: empty window");
        if (w && (windows[w].tid < windows[w - 1].tid ||
                  (windows[w].tid == windows[w - 1].tid && windows[w].begin <= windows[w - 1].end)))
            return svb_fail(ctx, SVB_ERR_ARG, "svb_window_depth: windows must be sorted and disjoint");
        woff[w] = tot;
        tot += (uint64_t)(windows[w].end - windows[w].begin + 1) + 1;
    }
    woff[n_w] = tot;
    uint64_t n_pos = tot - n_w;
    DevBuf<svb_window> dW;
    DevBuf<uint64_t> dOff;
    DevBuf<int32_t> diff, depth;
    DevBuf<uint32_t> flag, hot, ringbuf;
    DevBuf<uint8_t> kept;
    CK(dW.alloc(n_w, s));
    CK(dOff.alloc(n_w + 1, s));
    CK(diff.alloc(tot, s));
    CK(depth.alloc(n_pos, s));
    CK(cudaMemcpyAsync(dW.p, windows, n_w * sizeof(svb_window), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dOff.p, woff.data(), (n_w + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(diff.p, 0, tot * 4, s));
    if (n) {
        uint32_t n_ref = (uint32_t)bam->n_ref;
        CK(hot.alloc(n_ref + 1, s));
        CK(cudaMemsetAsync(hot.p, 0, (n_ref + 1) * 4, s));
        {
            ProfScope ps(ctx, "depth_marks", (double)n * 13);
            depth_marks<<<nblk(n, 256), 256, 0, s>>>(bam->d_data, bam->lean, min_mapq, bam->max_span, nullptr, hot.p, dW.p, dOff.p, n_w,
                                                     diff.p);
        }
        std::vector<uint32_t> hhot(n_ref);
        CK(cudaMemcpyAsync(hhot.data(), hot.p, n_ref * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        bool any_hot = false;
        for (uint32_t t = 0; t < n_ref; ++t) any_hot |= hhot[t] != 0;
        if (any_hot) {  // rare: coverage beyond libbam's pileup cap (quirk Q12) - serial, exact; then the marks are redone
            CK(flag.alloc(n, s));
            CK(kept.alloc(n, s));
            eligible_flags<<<nblk(n, 256), 256, 0, s>>>(bam->lean, min_mapq, flag.p);
            fill_u8<<<nblk(n, 256), 256, 0, s>>>(n, flag.p, kept.p);
            uint32_t ring = 1;
            while (ring < (uint32_t)bam->max_span + 2) ring <<= 1;
            CK(ringbuf.alloc((uint64_t)n_ref * ring, s));
            {
                ProfScope ps(ctx, "pileup_cap_serial", 0);
                cap_serial<<<nblk(n_ref, 32), 32, 0, s>>>(bam->lean, hot.p, n_ref, flag.p, kept.p, ringbuf.p, ring);
            }
            CK(cudaMemsetAsync(diff.p, 0, tot * 4, s));
            ProfScope ps(ctx, "depth_marks", (double)n * 13);
            depth_marks<<<nblk(n, 256), 256, 0, s>>>(bam->d_data, bam->lean, min_mapq, bam->max_span, kept.p, nullptr, dW.p, dOff.p, n_w,
                                                     diff.p);
        }
    }
    {
        ProfScope ps(ctx, "depth_scan", (double)tot * 8);
        depth_scan<<<nblk(n_w * 32, 128), 128, 0, s>>>(dW.p, dOff.p, n_w, diff.p, depth.p);
    }
    CK(cudaMemcpyAsync(depth_out, depth.p, n_pos * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return 0;
}
