// getsv / somatic device passes over the record rows: insert-size statistics
// (CalculateInsertsizeDeviation, cluster.cpp:15-83), discordant-pair support per junction
// (FindDiscordantReadPairs, getsv.cpp:990-1247 with IsConcordant cluster.cpp:136-147) and per-position
// depth inside merged junction windows (main_depth, bam2depth.cpp:17-142, libbam pileup semantics as
// probed in tests/test_oracle_golden.py). The record walker (walk.cu) streams the packed records once and leaves one 32-byte
// row per record in per-chunk slots; everything here reads rows (and, for the few records inside depth windows, their CIGARs).
//
// svb_getsv_passes runs the three passes as one stream-ordered sequence: mean and deviation stay on the device for the
// pair test, the host reads one control block plus the results at the end. The three single-pass entry points are the same
// stages with their own read-back.
#include "walk.cuh"

static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

#define PILEUP_MAXCNT 8000

struct SvCtl {
    unsigned long long tot[3];  // qualifying records, sum of isize, sum of isize^2
    int32_t q_max, mean, dev, need_slow;
    long long sq;               // sum of (isize - mean)^2 (fast path)
    uint32_t any_hot, pad;
};

__device__ __forceinline__ uint64_t row_key(int32_t tid, int32_t pos) { return (uint64_t)(uint32_t)tid << 32 | (uint32_t)(pos ^ 0x80000000); }  // tid -1 sorts last

__device__ __forceinline__ bool insert_qualifies(uint32_t fq, int32_t isize, int32_t min_mapq)
{
    uint32_t flag = fq & 0xffff;
    if ((int32_t)((fq >> 16) & 0xff) < min_mapq) return false;  // __g_skip_aln in cluster.cpp's TU (quirk Q9)
    if (fq & FLAGQ_HARDCLIP) return false;
    return (flag & F_PAIRED) && (flag & F_PROPER) && !(flag & F_DUP) && isize > 0;
}

struct RowsView {
    const Row *row;
    const uint16_t *roff;  // offset of a record inside its chunk
    uint32_t R, chunk_log2;
    const uint32_t *count;
    const uint64_t *base;  // dense index of a chunk's first row (n_chunks + 1 entries)
    const uint64_t *fkey;  // key of the first row at or after the chunk (rows_index)
    uint64_t n_chunks;
    uint64_t own_off;      // range shards: records that start before this stream offset belong to the previous shard (halo)
    __device__ __forceinline__ bool owned(uint64_t c, uint32_t k) const
    {
        if (own_off == 0) return true;
        const uint64_t oc = own_off >> chunk_log2;
        return c > oc || (c == oc && ((c << chunk_log2) + roff[c * R + k]) >= own_off);
    }
};

// ---- one pass over the rows: index (first keys, longest reference span, coordinate order) and insert-size partial sums -------
// one warp per chunk
template <bool INDEX, bool STATS>
__global__ void __launch_bounds__(256)
    rows_pass(RowsView V, uint64_t *__restrict__ fkey, int32_t *__restrict__ scal /* max span, unsorted */, int32_t stats_mapq,
              uint32_t *__restrict__ q_cnt, uint64_t *__restrict__ q_sum, uint64_t *__restrict__ q_sq, SvCtl *ctl)
{
    const uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (c >= V.n_chunks) return;
    const uint32_t cnt = min(V.count[c], V.R);
    const Row *rows = V.row + c * V.R;
    int32_t span = 0, qmax = 0;
    uint32_t unsorted = 0, qc = 0;
    uint64_t qs = 0, qq = 0;
    for (uint32_t k = lane; k < cnt; k += 32) {
        const Row r = rows[k];
        if (INDEX) {
            span = max(span, max(r.end - r.pos, 1));
            uint64_t prev = 0;
            bool have = false;
            if (k > 0) prev = row_key(rows[k - 1].tid, rows[k - 1].pos), have = true;
            else {  // against the last record of the closest earlier chunk that has one
                for (uint64_t j = c; j > 0;) {
                    --j;
                    const uint32_t cj = min(V.count[j], V.R);
                    if (cj) {
                        const Row p = V.row[j * V.R + cj - 1];
                        prev = row_key(p.tid, p.pos), have = true;
                        break;
                    }
                }
            }
            if (have && prev > row_key(r.tid, r.pos)) unsorted = 1;
        }
        if (STATS && insert_qualifies(r.flagq, r.isize, stats_mapq) && V.owned(c, k)) {
            ++qc, qs += (uint64_t)r.isize, qq += (uint64_t)r.isize * (uint64_t)r.isize;
            qmax = max(qmax, r.isize);
        }
    }
    if (INDEX) {
        span = (int32_t)warp_max((uint32_t)span);
        unsorted = warp_max(unsorted);
        if (lane == 0) {
            // (one atomic per chunk on a single word serialises in L2: only chunks that would change the word issue one)
            if (span > *(volatile int32_t *)&scal[0]) atomicMax(&scal[0], span);
            if (unsorted) atomicOr((uint32_t *)&scal[1], 1u);
            uint64_t k0 = ~0ull;  // no record at or after this chunk
            for (uint64_t j = c; j < V.n_chunks; ++j)
                if (V.count[j]) {
                    const Row f = V.row[j * V.R];
                    k0 = row_key(f.tid, f.pos);
                    break;
                }
            fkey[c] = k0;
        }
    }
    if (STATS) {
        qmax = (int32_t)warp_max((uint32_t)qmax);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            qc += __shfl_xor_sync(0xffffffffu, qc, o);
            qs += __shfl_xor_sync(0xffffffffu, qs, o);
            qq += __shfl_xor_sync(0xffffffffu, qq, o);
        }
        if (lane == 0) {
            q_cnt[c] = qc, q_sum[c] = qs, q_sq[c] = qq;  // (summed by insert_finish: no per-chunk atomics on one address)
            if (qmax > *(volatile int32_t *)&ctl->q_max) atomicMax(&ctl->q_max, qmax);
        }
    }
}

// mean and deviation on the device (cluster.cpp:72-80) when the closed form is exact: every qualifying record is used
// (total <= -n) and every |isize - mean| stays below sqrt(2^31), so that the reference's int products cannot wrap and
// sum (x - m)^2 = sum x^2 - 2 m sum x + n m^2 holds in 64-bit integers. Otherwise need_slow asks the host for the ordered path.
__global__ void __launch_bounds__(256)
    insert_totals(SvCtl *ctl, uint64_t n_chunks, const uint32_t *__restrict__ q_cnt, const uint64_t *__restrict__ q_sum, const uint64_t *__restrict__ q_sq)
{
    __shared__ unsigned long long red[3][8];
    unsigned long long a = 0, b = 0, q = 0;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x)
        a += q_cnt[c], b += q_sum[c], q += q_sq[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = a, red[1][threadIdx.x >> 5] = b, red[2][threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x != 0) return;
    a = b = q = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) a += red[0][w], b += red[1][w], q += red[2][w];
    if (a) atomicAdd(&ctl->tot[0], a), atomicAdd(&ctl->tot[1], b), atomicAdd(&ctl->tot[2], q);
}
__global__ void insert_finish(SvCtl *ctl, long long max_pairs)
{
    const unsigned long long n = ctl->tot[0];
    ctl->mean = 0, ctl->dev = 0, ctl->sq = 0, ctl->need_slow = 0;
    if (n == 0 || max_pairs <= 0) return;
    if (n > (unsigned long long)max_pairs || ctl->q_max > 46340) {
        ctl->need_slow = 1;
        return;
    }
    const long long mean = (long long)(ctl->tot[1] / n);
    const long long sq = (long long)ctl->tot[2] - 2 * mean * (long long)ctl->tot[1] + (long long)n * mean * mean;
    ctl->mean = (int32_t)mean, ctl->sq = sq;
    ctl->dev = (int32_t)sqrt((double)sq / (double)(int32_t)n);  // IEEE division and square root: the host's result
}

// ordered path: the first max_pairs qualifying records in file order (cluster.cpp:48-70)
struct QBaseOp {
    uint64_t n_chunks;
    const uint32_t *q_cnt;
    uint64_t *q_base;
    __device__ uint64_t n() const { return n_chunks; }
    __device__ void load(uint64_t c, uint64_t (&v)[1]) const { v[0] = q_cnt[c]; }
    __device__ void store(uint64_t c, const uint64_t (&excl)[1], const uint64_t (&)[1]) const { q_base[c] = excl[0]; }
    __device__ void total(const uint64_t (&)[1]) const {}
};
// pass 1: sum of isize; pass 2: sum of (int32)((isize-mean)*(isize-mean)); one warp per chunk
__global__ void __launch_bounds__(256)
    insert_ordered(RowsView V, int32_t min_mapq, const uint64_t *__restrict__ q_base, uint64_t max_pairs, int pass, int32_t mean,
                   unsigned long long *__restrict__ acc)
{
    const uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (c >= V.n_chunks) return;
    const uint32_t cnt = min(V.count[c], V.R);
    uint64_t run = q_base[c];
    if (run >= max_pairs) return;
    const Row *rows = V.row + c * V.R;
    long long v = 0, n3 = 0, sq3 = 0, big3 = 0;
    for (uint32_t k0 = 0; k0 < cnt; k0 += 32) {
        const uint32_t k = k0 + lane;
        bool q = false;
        int32_t isize = 0;
        if (k < cnt) {
            const Row r = rows[k];
            isize = r.isize;
            q = insert_qualifies(r.flagq, isize, min_mapq) && V.owned(c, k);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, q);
        if (q && run + __popc(m & ((1u << lane) - 1u)) < max_pairs) {
            if (pass == 1) v += isize;
            else if (pass == 3) {  // shard partials: count, sum of squares and "large" count next to the sum
                v += isize;
                n3 += 1, sq3 += (long long)isize * isize, big3 += isize > 46340;
            } else {
                uint32_t dlt = (uint32_t)(isize - mean);
                v += (int32_t)(dlt * dlt);  // the reference multiplies two ints (cluster.cpp:77)
            }
        }
        run += __popc(m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        if (pass == 3) {
            n3 += __shfl_xor_sync(0xffffffffu, n3, o);
            sq3 += __shfl_xor_sync(0xffffffffu, sq3, o);
            big3 += __shfl_xor_sync(0xffffffffu, big3, o);
        }
    }
    if (lane == 0 && v) atomicAdd(acc, (unsigned long long)v);
    if (lane == 0 && pass == 3 && n3) {
        atomicAdd(acc + 1, (unsigned long long)n3);
        atomicAdd(acc + 2, (unsigned long long)sq3);
        if (big3) atomicAdd(acc + 3, (unsigned long long)big3);
    }
}

// ---- discordant read pairs ------------------------------------------------------------------------------------------
// first chunk whose first-record key is >= K (n_chunks when none); rows >= K start in the chunk before it
__device__ __forceinline__ uint64_t chunk_lower_bound(const uint64_t *__restrict__ fkey, uint64_t n_chunks, uint64_t K)
{
    uint64_t lo = 0, hi = n_chunks;
    while (lo < hi) {
        uint64_t m = (lo + hi) >> 1;
        if (fkey[m] < K) lo = m + 1;
        else hi = m;
    }
    return lo;
}

__global__ void __launch_bounds__(128)
    discordant_kernel(RowsView V, const int32_t *__restrict__ scal, const svb_junction *__restrict__ J, uint64_t n_j,
                      const uint32_t *__restrict__ ref_len, svb_pair_params prm, const SvCtl *__restrict__ ctl, int32_t *__restrict__ counts)
{
    uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (w >= n_j) return;
    if (ctl) prm.mean_insert = ctl->mean, prm.deviation = ctl->dev;  // (fused call: the statistics never left the device)
    const int32_t max_span = scal[0];
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
        const int64_t plo = (int64_t)beg - max_span;
        const uint64_t Klo = row_key(tid, (int32_t)max(plo, (int64_t)INT32_MIN)), Khi = row_key(tid, end);
        uint64_t c0 = chunk_lower_bound(V.fkey, V.n_chunks, Klo), c1 = chunk_lower_bound(V.fkey, V.n_chunks, Khi);
        if (c0 > 0) --c0;
        for (uint64_t c = c0; c < c1; ++c) {
            const uint32_t cnt = min(V.count[c], V.R);
            for (uint32_t k = lane; k < cnt; k += 32) {
                const Row r = V.row[c * V.R + k];
                const uint64_t key = row_key(r.tid, r.pos);
                if (key < Klo || key >= Khi || !V.owned(c, k)) continue;
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
    }
    n = warp_sum(n);
    if (lane == 0) counts[w] = (int32_t)n;
}

// ---- window depth ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool pileup_eligible(int32_t tid, uint32_t fq, int32_t min_mapq)
{
    // bam_plp_push: tid >= 0 and (flag & 0x704) == 0, after read_bam (bam2depth.h:29-35) set UNMAP for low mapQ
    return tid >= 0 && !((fq & 0xffff) & 0x704) && (int32_t)((fq >> 16) & 0xff) >= min_mapq;
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

// chunks that hold records which can reach a window: one thread per window, each chunk listed once
__global__ void depth_select(RowsView V, const int32_t *__restrict__ scal, const svb_window *__restrict__ W, uint64_t n_w, uint32_t *__restrict__ pick,
                             uint32_t *__restrict__ list, uint32_t *__restrict__ n_list)
{
    const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_w) return;
    const int32_t max_span = scal[0];
    const int64_t plo = (int64_t)W[w].begin - 1 - max_span;  // 0-based pos of a record that can still reach the window
    const uint64_t Klo = row_key(W[w].tid, (int32_t)max(plo, (int64_t)INT32_MIN)), Khi = row_key(W[w].tid, W[w].end);  // pos < end (1-based end = pos + 1 <= end)
    uint64_t c0 = chunk_lower_bound(V.fkey, V.n_chunks, Klo), c1 = chunk_lower_bound(V.fkey, V.n_chunks, Khi);
    if (c0 > 0) --c0;
    for (uint64_t c = c0; c < c1; ++c)
        if (atomicExch(&pick[c], 1u) == 0) list[atomicAdd(n_list, 1u)] = (uint32_t)c;
}

// One warp per listed chunk, one lane per record: +1/-1 marks of the M segments of every eligible record into the difference
// array of each window it overlaps (the rows give tid / pos / end / flags, roff the place of the record for its CIGAR).
// kept == nullptr: every pileup-eligible read is kept. kept != nullptr: the survivors of libbam's 8000-read cap (cap_serial).
__global__ void __launch_bounds__(128)
    depth_marks(const uint8_t *__restrict__ d, RowsView V, const uint32_t *__restrict__ list, const uint32_t *__restrict__ n_list, int32_t min_mapq,
                const uint8_t *__restrict__ kept, const svb_window *__restrict__ W, const uint64_t *__restrict__ woff, uint64_t n_w,
                int32_t *__restrict__ diff)
{
    const uint32_t lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5, n = *n_list;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += n_warps) {
        const uint64_t c = list[i];
        const uint32_t cnt = min(V.count[c], V.R);
        for (uint32_t k = lane; k < cnt; k += 32) {
            const Row r = V.row[c * V.R + k];
            const int32_t tid = r.tid;
            if (!pileup_eligible(tid, r.flagq, min_mapq) || !V.owned(c, k)) continue;
            if (kept && !kept[c * V.R + k]) continue;
            const int32_t beg1 = r.pos + 1, end1 = r.end;  // 1-based inclusive [beg1, end1]
            if (end1 < beg1) continue;
            uint64_t w = first_window(W, n_w, tid, beg1);
            if (w >= n_w || W[w].tid != tid || W[w].begin > end1) continue;
            const uint8_t *p = d + (c << V.chunk_log2) + V.roff[c * V.R + k];
            const uint32_t lq = ldu32(p + 12) & 0xff, nc = ldu32(p + 16) & 0xffff;
            const uint8_t *cig = p + 36 + lq;
            for (; w < n_w && W[w].tid == tid && W[w].begin <= end1; ++w) {
                int32_t wb = W[w].begin, we = W[w].end;
                int32_t *dw = diff + woff[w];
                int32_t x = beg1;
                for (uint32_t j = 0; j < nc && x <= we; ++j) {
                    uint32_t cw = ldu32(cig + 4 * j), op = cw & 15;
                    int32_t len = (int32_t)(cw >> 4);
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
    }
}

// libbam's pileup cap (quirk Q12) can only bite where the record 7998 places earlier on the same chromosome starts within
// max_span of an eligible record: more than 8000 buffer nodes are possible there and the chromosome is flagged for the exact
// serial emulation (cap_serial). One thread per chunk; a chunk is looked at row by row only when 17 whole chunks in front of it
// (a chunk holds at most 431 records, so 7998 records span more than 18 chunks) lie within max_span positions.
__global__ void __launch_bounds__(128) hot_check(RowsView V, const int32_t *__restrict__ scal, int32_t min_mapq, uint32_t *__restrict__ hot_tids, SvCtl *ctl)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= V.n_chunks || c < 17) return;
    const uint32_t cnt = min(V.count[c], V.R);
    if (!cnt) return;
    const int32_t max_span = scal[0];
    const Row f = V.row[c * V.R];
    const uint64_t kp = V.fkey[c - 17];
    if ((int32_t)(kp >> 32) != f.tid || (int64_t)(int32_t)((uint32_t)kp ^ 0x80000000) < (int64_t)f.pos - max_span) return;
    uint64_t cj = c - 17;  // chunk of row i - 7998, found by walking back over the dense row bases
    for (uint32_t k = 0; k < cnt; ++k) {
        const uint64_t i = V.base[c] + k;
        if (i < PILEUP_MAXCNT - 2) continue;
        const uint64_t j = i - (PILEUP_MAXCNT - 2);
        while (V.base[cj] > j) --cj;
        while (V.base[cj + 1] <= j) ++cj;
        const Row r = V.row[c * V.R + k];
        if (!pileup_eligible(r.tid, r.flagq, min_mapq)) continue;
        const Row q = V.row[cj * V.R + (j - V.base[cj])];
        if (q.tid == r.tid && q.pos >= r.pos - max_span) {
            hot_tids[r.tid] = 1;
            ctl->any_hot = 1;
        }
    }
}

// Exact serial emulation of bam_plp_push's cap for one chromosome (one thread per hot chromosome): a read is
// refused only if it starts at the same position as the previously accepted read while more than 8000
// buffer nodes are allocated = 2 + accepted reads whose end >= that position (released lazily).
__global__ void cap_serial(RowsView V, const uint32_t *__restrict__ hot_tids, uint32_t n_ref, int32_t min_mapq,
                           uint8_t *__restrict__ kept, uint32_t *__restrict__ ring_all, uint32_t ring)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ref || !hot_tids[t]) return;
    uint32_t *hist = ring_all + (uint64_t)t * ring;  // live accepted reads by end position (mod ring)
    for (uint32_t k = 0; k < ring; ++k) hist[k] = 0;
    uint64_t c0 = chunk_lower_bound(V.fkey, V.n_chunks, row_key((int32_t)t, INT32_MIN));
    const uint64_t c1 = chunk_lower_bound(V.fkey, V.n_chunks, row_key((int32_t)t + 1, INT32_MIN));
    if (c0 > 0) --c0;
    int64_t it_pos = -1;  // position of the previously accepted read (iterator position)
    uint32_t live = 0;
    bool any = false;
    for (uint64_t c = c0; c < c1; ++c) {
        const uint32_t cnt = min(V.count[c], V.R);
        for (uint32_t k = 0; k < cnt; ++k) {
            const Row r = V.row[c * V.R + k];
            if (r.tid != (int32_t)t || !pileup_eligible(r.tid, r.flagq, min_mapq)) continue;
            const int32_t pos = r.pos, end = r.end;
            if (any && pos == it_pos) {
                if (live + 2 > PILEUP_MAXCNT) {
                    kept[c * V.R + k] = 0;
                    continue;
                }
                if (end > pos) hist[(uint32_t)end % ring]++, live++;
            } else {
                if (any) {  // release nodes whose end <= pos - 1
                    int64_t from = it_pos, to = pos;  // ends in [from, to) leave; ends < from left earlier
                    if (to - from >= ring) {
                        for (uint32_t k2 = 0; k2 < ring; ++k2) hist[k2] = 0;
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

// ---- host orchestration ---------------------------------------------------------------------------------------------------
namespace {
RowsView view_of(const svb_bam *bam)
{
    return RowsView{bam->rows.row, bam->rows.roff, bam->rows.R, bam->chunk_log2, bam->d_count, bam->d_base, bam->d_fkey, bam->n_chunks, bam->own_offset};
}

struct SvBuffers {
    SvCtl *ctl;
    uint32_t *q_cnt;
    uint64_t *q_sum, *q_sq, *q_base;
    unsigned long long *acc;
    ScanScratch sc_q;
    svb_junction *J;
    int32_t *counts;
    svb_window *W;
    uint64_t *woff;
    int32_t *diff, *depth;
    uint32_t *pick, *list, *n_list;
    uint32_t *hot;
    size_t zero_end;
};
void carve(Bump &b, SvBuffers &B, uint64_t n_chunks, uint64_t n_j, uint64_t n_w, uint64_t diff_len, uint64_t n_pos, int32_t n_ref)
{
    B.ctl = b.get<SvCtl>(1);
    B.acc = b.get<unsigned long long>(4);
    B.sc_q = scan_scratch(b, n_chunks, 1, 8);
    B.pick = b.get<uint32_t>(n_w ? n_chunks : 1);
    B.n_list = b.get<uint32_t>(1);
    B.hot = b.get<uint32_t>((size_t)n_ref + 1);
    B.diff = b.get<int32_t>(diff_len);
    B.zero_end = (b.used + 255) & ~(size_t)255;
    B.q_cnt = b.get<uint32_t>(n_chunks), B.q_sum = b.get<uint64_t>(n_chunks), B.q_sq = b.get<uint64_t>(n_chunks), B.q_base = b.get<uint64_t>(n_chunks);
    B.J = b.get<svb_junction>(n_j), B.counts = b.get<int32_t>(n_j);
    B.W = b.get<svb_window>(n_w), B.woff = b.get<uint64_t>(n_w + 1);
    B.depth = b.get<int32_t>(n_pos);
    B.list = b.get<uint32_t>(n_w ? n_chunks : 1);
}

// rows of the records + their index (first keys, span, order): made once per handle
int prepare_rows(svb_ctx *ctx, svb_bam *bam)
{
    CKR(ensure_rows(ctx, bam));
    cudaStream_t s = ctx->stream;
    if (!bam->d_scal) {
        CK(cudaMallocAsync((void **)&bam->d_scal, 16, s));
        CK(cudaMallocAsync((void **)&bam->d_fkey, bam->n_chunks * 8, s));
    }
    return 0;
}
int upload_ref_lens(svb_ctx *ctx, svb_bam *bam)
{
    if (bam->d_ref_len) return 0;
    if (bam->lens.size() != (size_t)bam->n_ref) return svb_fail(ctx, SVB_ERR_ARG, "reference lengths not set (svb_bam_set_refs)");
    CK(cudaMallocAsync((void **)&bam->d_ref_len, ((size_t)bam->n_ref + 1) * 4, ctx->stream));
    CK(cudaMemcpyAsync(bam->d_ref_len, bam->lens.data(), (size_t)bam->n_ref * 4, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

struct Request {
    bool stats = false;
    int32_t stats_mapq = 0;
    int64_t max_pairs = 0;
    const svb_junction *junctions = nullptr;
    uint64_t n_j = 0;
    svb_pair_params pp{};
    bool pp_from_stats = false;
    const svb_window *windows = nullptr;
    uint64_t n_w = 0;
    int32_t depth_mapq = 0;
    int64_t *stats_out = nullptr;
    int32_t *counts = nullptr, *depth_out = nullptr;
};

int run_passes(svb_ctx *ctx, svb_bam *bam, const Request &rq)
{
    CK(cudaSetDevice(ctx->device));
    CKR(prepare_rows(ctx, bam));
    cudaStream_t s = ctx->stream;
    const uint64_t n_chunks = bam->n_chunks;
    // windows: layout of the difference arrays (one extra slot per window)
    std::vector<uint64_t> woff(rq.n_w + 1);
    uint64_t tot = 0;
    for (uint64_t w = 0; w < rq.n_w; ++w) {
        const svb_window *W = rq.windows;
        if (W[w].end < W[w].begin) return svb_fail(ctx, SVB_ERR_ARG, "svb_window_depth: empty window");
        if (w && (W[w].tid < W[w - 1].tid || (W[w].tid == W[w - 1].tid && W[w].begin <= W[w - 1].end)))
            return svb_fail(ctx, SVB_ERR_ARG, "svb_window_depth: windows must be sorted and disjoint");
        woff[w] = tot;
        tot += (uint64_t)(W[w].end - W[w].begin + 1) + 1;
    }
    woff[rq.n_w] = tot;
    const uint64_t n_pos = tot - rq.n_w;
    if (rq.n_j) CKR(upload_ref_lens(ctx, bam));
    SvBuffers B{};
    {
        Bump measure(nullptr);
        carve(measure, B, n_chunks, rq.n_j, rq.n_w, tot, n_pos, bam->n_ref);
        CKR(ctx->ws_reserve(1, measure.used));
        Bump real(ctx->ws[1]);
        carve(real, B, n_chunks, rq.n_j, rq.n_w, tot, n_pos, bam->n_ref);
    }
    CK(cudaMemsetAsync(ctx->ws[1], 0, B.zero_end, s));
    const bool need_index = !bam->rows_indexed;
    if (need_index) CK(cudaMemsetAsync(bam->d_scal, 0, 16, s));
    RowsView V = view_of(bam);
    const unsigned g_chunks = nblk(n_chunks * 32, 256);
    if (need_index || rq.stats) {
        ProfScope ps(ctx, "rows_pass", (double)bam->n_rec * sizeof(Row));
        if (need_index && rq.stats) rows_pass<true, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, rq.stats_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
        else if (need_index) rows_pass<true, false><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, 0, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
        else rows_pass<false, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, rq.stats_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
        if (rq.stats) {
            insert_totals<<<(unsigned)std::min<uint64_t>(nblk(n_chunks, 256), (uint64_t)ctx->sm_count), 256, 0, s>>>(B.ctl, n_chunks, B.q_cnt, B.q_sum, B.q_sq);
            insert_finish<<<1, 1, 0, s>>>(B.ctl, (long long)rq.max_pairs);
        }
    }
    const bool fused_pairs = rq.n_j && rq.pp_from_stats;
    auto launch_pairs = [&](const SvCtl *from_ctl) -> int {
        CK(cudaMemcpyAsync(B.J, rq.junctions, rq.n_j * sizeof(svb_junction), cudaMemcpyHostToDevice, s));
        ProfScope ps(ctx, "discordant_support", 0);
        discordant_kernel<<<nblk(rq.n_j * 32, 128), 128, 0, s>>>(V, bam->d_scal, B.J, rq.n_j, bam->d_ref_len, rq.pp, from_ctl, B.counts);
        CK(cudaMemcpyAsync(rq.counts, B.counts, rq.n_j * 4, cudaMemcpyDefault, s));  // (the caller's array may live on the device)
        return 0;
    };
    if (rq.n_j) CKR(launch_pairs(fused_pairs ? B.ctl : nullptr));
    if (rq.n_w) {
        CK(cudaMemcpyAsync(B.W, rq.windows, rq.n_w * sizeof(svb_window), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(B.woff, woff.data(), (rq.n_w + 1) * 8, cudaMemcpyHostToDevice, s));
        {
            ProfScope ps(ctx, "depth_marks", 0);
            depth_select<<<nblk(rq.n_w, 128), 128, 0, s>>>(V, bam->d_scal, B.W, rq.n_w, B.pick, B.list, B.n_list);
            depth_marks<<<grid_for(ctx, n_chunks * 32, 128, 8), 128, 0, s>>>(bam->d_data, V, B.list, B.n_list, rq.depth_mapq, nullptr, B.W, B.woff, rq.n_w, B.diff);
            hot_check<<<nblk(n_chunks, 128), 128, 0, s>>>(V, bam->d_scal, rq.depth_mapq, B.hot, B.ctl);
        }
        {
            ProfScope ps(ctx, "depth_scan", (double)tot * 8);
            depth_scan<<<nblk(rq.n_w * 32, 128), 128, 0, s>>>(B.W, B.woff, rq.n_w, B.diff, B.depth);
        }
        CK(cudaMemcpyAsync(rq.depth_out, B.depth, n_pos * 4, cudaMemcpyDefault, s));
    }
    // ---- the one read-back
    struct {
        SvCtl ctl;
        int32_t scal[4];
    } *h = (decltype(h))ctx->ctl_host;
    CK(cudaMemcpyAsync(&h->ctl, B.ctl, sizeof(SvCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->scal, bam->d_scal, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    const SvCtl hc = h->ctl;
    bam->max_span = h->scal[0], bam->sorted = h->scal[1] ? 0 : 1, bam->rows_indexed = true;
    if ((rq.n_j || rq.n_w) && bam->sorted != 1) return svb_fail(ctx, SVB_ERR_UNSORTED, "the BAM is not coordinate-sorted");
    if (rq.stats) {
        int64_t *out = rq.stats_out;
        out[0] = out[1] = out[2] = out[3] = 0;
        bool redo_pairs = false;
        if (hc.need_slow) {
            // ordered path: the -n cut-off falls inside the file, or an insert size is large enough for the reference's int
            // products to wrap - the first max_pairs qualifying records in file order, products truncated like the reference's
            QBaseOp op{n_chunks, B.q_cnt, B.q_base};
            CK(cudaMemsetAsync(B.sc_q.ticket, 0, 4, s));
            CK(cudaMemsetAsync(B.sc_q.state, 0, (size_t)B.sc_q.tiles_cap * 8, s));
            CK(cudaMemsetAsync(B.acc, 0, 32, s));
            {
                ProfScope ps(ctx, "insert_stats", 0);
                launch_scan<1, 8>(ctx, s, op, B.sc_q, n_chunks);
                insert_ordered<<<g_chunks, 256, 0, s>>>(V, rq.stats_mapq, B.q_base, (uint64_t)rq.max_pairs, 1, 0, B.acc);
            }
            unsigned long long sum = 0;
            CK(cudaMemcpyAsync(&sum, B.acc, 8, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            const uint64_t cnt = std::min<uint64_t>(hc.tot[0], (uint64_t)rq.max_pairs);
            if (cnt) {
                const int32_t mean = (int32_t)(sum / cnt);  // unsigned long / int, stored to int (cluster.cpp:72)
                {
                    ProfScope ps(ctx, "insert_stats", 0);
                    insert_ordered<<<g_chunks, 256, 0, s>>>(V, rq.stats_mapq, B.q_base, (uint64_t)rq.max_pairs, 2, mean, B.acc + 1);
                }
                long long sq = 0;
                CK(cudaMemcpyAsync(&sq, B.acc + 1, 8, cudaMemcpyDeviceToHost, s));
                CK(cudaStreamSynchronize(s));
                out[0] = (int64_t)cnt, out[1] = (int64_t)sum, out[2] = mean, out[3] = sq;
            }
            redo_pairs = fused_pairs;
        } else if (hc.tot[0] && rq.max_pairs > 0) {
            out[0] = (int64_t)hc.tot[0], out[1] = (int64_t)hc.tot[1], out[2] = hc.mean, out[3] = hc.sq;
        }
        if (redo_pairs) {  // the device had no statistics when the pair test ran: once more with the host's
            Request r2 = rq;
            r2.pp.mean_insert = (int32_t)out[2];
            r2.pp.deviation = out[0] ? (int32_t)sqrt((double)out[3] / (double)(int32_t)out[0]) : 0;
            CK(cudaMemcpyAsync(B.J, rq.junctions, rq.n_j * sizeof(svb_junction), cudaMemcpyHostToDevice, s));
            discordant_kernel<<<nblk(rq.n_j * 32, 128), 128, 0, s>>>(V, bam->d_scal, B.J, rq.n_j, bam->d_ref_len, r2.pp, nullptr, B.counts);
            CK(cudaMemcpyAsync(rq.counts, B.counts, rq.n_j * 4, cudaMemcpyDefault, s));  // (the caller's array may live on the device)
            CK(cudaStreamSynchronize(s));
        }
    }
    if (rq.n_w && hc.any_hot) {
        // rare: coverage beyond libbam's pileup cap (quirk Q12) - serial, exact; then the marks are redone with the survivors
        DevBuf<uint8_t> kept;
        DevBuf<uint32_t> ringbuf;
        const uint32_t n_ref = (uint32_t)bam->n_ref;
        CK(kept.alloc(n_chunks * bam->rows.R, s));
        CK(cudaMemsetAsync(kept.p, 1, n_chunks * bam->rows.R, s));
        uint32_t ring = 1;
        while (ring < (uint32_t)bam->max_span + 2) ring <<= 1;
        CK(ringbuf.alloc((uint64_t)n_ref * ring, s));
        {
            ProfScope ps(ctx, "pileup_cap_serial", 0);
            cap_serial<<<nblk(n_ref, 32), 32, 0, s>>>(V, B.hot, n_ref, rq.depth_mapq, kept.p, ringbuf.p, ring);
        }
        CK(cudaMemsetAsync(B.diff, 0, tot * 4, s));
        {
            ProfScope ps(ctx, "depth_marks", 0);
            depth_marks<<<grid_for(ctx, n_chunks * 32, 128, 8), 128, 0, s>>>(bam->d_data, V, B.list, B.n_list, rq.depth_mapq, kept.p, B.W, B.woff, rq.n_w, B.diff);
        }
        depth_scan<<<nblk(rq.n_w * 32, 128), 128, 0, s>>>(B.W, B.woff, rq.n_w, B.diff, B.depth);
        CK(cudaMemcpyAsync(rq.depth_out, B.depth, n_pos * 4, cudaMemcpyDefault, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
    }
    return 0;
}
}  // namespace

extern "C" int svb_insert_stats(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t max_pairs, int64_t out[4])
{
    if (!ctx || !bam || !out) return svb_fail(ctx, SVB_ERR_ARG, "svb_insert_stats: null argument");
    Request rq;
    rq.stats = true, rq.stats_mapq = min_mapq, rq.max_pairs = max_pairs, rq.stats_out = out;
    return run_passes(ctx, bam, rq);
}

extern "C" int svb_discordant_support(svb_ctx *ctx, svb_bam *bam, const svb_junction *junctions, uint64_t n,
                                      const svb_pair_params *p, int32_t *counts)
{
    if (!ctx || !bam || !p || (n && (!junctions || !counts))) return svb_fail(ctx, SVB_ERR_ARG, "svb_discordant_support: null argument");
    Request rq;
    rq.junctions = junctions, rq.n_j = n, rq.pp = *p, rq.counts = counts;
    return run_passes(ctx, bam, rq);
}

extern "C" int svb_window_depth(svb_ctx *ctx, svb_bam *bam, const svb_window *windows, uint64_t n_w, int32_t min_mapq,
                                int32_t *depth_out)
{
    if (!ctx || !bam || (n_w && (!windows || !depth_out))) return svb_fail(ctx, SVB_ERR_ARG, "svb_window_depth: null argument");
    Request rq;
    rq.windows = windows, rq.n_w = n_w, rq.depth_mapq = min_mapq, rq.depth_out = depth_out;
    return run_passes(ctx, bam, rq);
}

extern "C" int svb_getsv_passes(svb_ctx *ctx, svb_bam *bam, const svb_getsv_params *p, const svb_junction *junctions, uint64_t n_j,
                                const svb_window *windows, uint64_t n_w, int64_t stats_out[4], int32_t *counts, int32_t *depth_out)
{
    if (!ctx || !bam || !p || !stats_out || (n_j && (!junctions || !counts)) || (n_w && (!windows || !depth_out)))
        return svb_fail(ctx, SVB_ERR_ARG, "svb_getsv_passes: null argument");
    Request rq;
    rq.stats = true, rq.stats_mapq = p->min_mapq, rq.max_pairs = p->max_pairs, rq.stats_out = stats_out;
    rq.junctions = junctions, rq.n_j = n_j, rq.counts = counts, rq.pp_from_stats = true;
    rq.pp.min_mapq = p->min_mapq, rq.pp.times = p->times;
    rq.windows = windows, rq.n_w = n_w, rq.depth_mapq = p->min_mapq, rq.depth_out = depth_out;
    return run_passes(ctx, bam, rq);
}

// ---- shards of one BAM on several GPUs: the statistics in additive pieces ---------------------------------
// out = {records taken, sum of isize, sum of isize^2, records with isize > 46340} over the first `take` qualifying OWN records of
// this shard in file order (take < 0: all of them). The ranks add these up (NCCL) and derive mean and deviation as
// insert_finish does; a non-zero fourth value means the reference's int products could wrap and the exact second pass
// (svb_insert_sq) is needed.
extern "C" int svb_insert_partial(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t take, int64_t out[4])
{
    if (!ctx || !bam || !out) return svb_fail(ctx, SVB_ERR_ARG, "svb_insert_partial: null argument");
    CK(cudaSetDevice(ctx->device));
    CKR(prepare_rows(ctx, bam));
    cudaStream_t s = ctx->stream;
    const uint64_t n_chunks = bam->n_chunks;
    SvBuffers B{};
    {
        Bump measure(nullptr);
        carve(measure, B, n_chunks, 0, 0, 1, 1, bam->n_ref);
        CKR(ctx->ws_reserve(1, measure.used));
        Bump real(ctx->ws[1]);
        carve(real, B, n_chunks, 0, 0, 1, 1, bam->n_ref);
    }
    CK(cudaMemsetAsync(ctx->ws[1], 0, B.zero_end, s));
    const bool need_index = !bam->rows_indexed;
    if (need_index) CK(cudaMemsetAsync(bam->d_scal, 0, 16, s));
    RowsView V = view_of(bam);
    const unsigned g_chunks = nblk(n_chunks * 32, 256);
    {
        ProfScope ps(ctx, "rows_pass", (double)bam->n_rec * sizeof(Row));
        if (need_index) rows_pass<true, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, min_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
        else rows_pass<false, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, min_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
        insert_totals<<<(unsigned)std::min<uint64_t>(nblk(n_chunks, 256), (uint64_t)ctx->sm_count), 256, 0, s>>>(B.ctl, n_chunks, B.q_cnt, B.q_sum, B.q_sq);
    }
    struct {
        SvCtl ctl;
        int32_t scal[4];
    } *h = (decltype(h))ctx->ctl_host;
    CK(cudaMemcpyAsync(&h->ctl, B.ctl, sizeof(SvCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->scal, bam->d_scal, 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    const SvCtl hc = h->ctl;
    bam->max_span = h->scal[0], bam->sorted = h->scal[1] ? 0 : 1, bam->rows_indexed = true;
    if (take < 0 || (uint64_t)take >= hc.tot[0]) {
        out[0] = (int64_t)hc.tot[0], out[1] = (int64_t)hc.tot[1], out[2] = (int64_t)hc.tot[2], out[3] = hc.q_max > 46340 ? 1 : 0;
        return 0;
    }
    QBaseOp op{n_chunks, B.q_cnt, B.q_base};
    launch_scan<1, 8>(ctx, s, op, B.sc_q, n_chunks);
    insert_ordered<<<g_chunks, 256, 0, s>>>(V, min_mapq, B.q_base, (uint64_t)take, 3, 0, B.acc);
    unsigned long long a[4];
    CK(cudaMemcpyAsync(a, B.acc, 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    out[0] = (int64_t)a[1], out[1] = (int64_t)a[0], out[2] = (int64_t)a[2], out[3] = (int64_t)a[3];
    return 0;
}

// {taken, sum, sum of squares, large} for the async form below: from the totals (take < 0) or from the ordered kernel's accumulators
__global__ void partial_out(const SvCtl *__restrict__ ctl, const unsigned long long *__restrict__ acc, int from_acc, long long *__restrict__ out)
{
    if (from_acc) out[0] = (long long)acc[1], out[1] = (long long)acc[0], out[2] = (long long)acc[2], out[3] = (long long)acc[3];
    else out[0] = (long long)ctl->tot[0], out[1] = (long long)ctl->tot[1], out[2] = (long long)ctl->tot[2], out[3] = ctl->q_max > 46340 ? 1 : 0;
}

// svb_insert_partial without the read-back: the four values are written to d_out (DEVICE memory, int64[4]) in stream order, so
// that the ranks can feed them to a collective directly. take >= 0 must not exceed the shard's qualifying count by the caller's
// own bookkeeping (it is clamped by the kernels either way).
extern "C" int svb_insert_partial_async(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t take, int64_t *d_out)
{
    if (!ctx || !bam || !d_out) return svb_fail(ctx, SVB_ERR_ARG, "svb_insert_partial_async: null argument");
    CK(cudaSetDevice(ctx->device));
    CKR(prepare_rows(ctx, bam));
    cudaStream_t s = ctx->stream;
    const uint64_t n_chunks = bam->n_chunks;
    SvBuffers B{};
    {
        Bump measure(nullptr);
        carve(measure, B, n_chunks, 0, 0, 1, 1, bam->n_ref);
        CKR(ctx->ws_reserve(1, measure.used));
        Bump real(ctx->ws[1]);
        carve(real, B, n_chunks, 0, 0, 1, 1, bam->n_ref);
    }
    CK(cudaMemsetAsync(ctx->ws[1], 0, B.zero_end, s));
    const bool need_index = !bam->rows_indexed;
    if (need_index) CK(cudaMemsetAsync(bam->d_scal, 0, 16, s));
    RowsView V = view_of(bam);
    const unsigned g_chunks = nblk(n_chunks * 32, 256);
    {
        ProfScope ps(ctx, "rows_pass", (double)bam->n_rec * sizeof(Row));
        if (need_index) rows_pass<true, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, min_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
        else rows_pass<false, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, min_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
        if (take < 0) insert_totals<<<(unsigned)std::min<uint64_t>(nblk(n_chunks, 256), (uint64_t)ctx->sm_count), 256, 0, s>>>(B.ctl, n_chunks, B.q_cnt, B.q_sum, B.q_sq);
        else {
            QBaseOp op{n_chunks, B.q_cnt, B.q_base};
            launch_scan<1, 8>(ctx, s, op, B.sc_q, n_chunks);
            insert_ordered<<<g_chunks, 256, 0, s>>>(V, min_mapq, B.q_base, (uint64_t)take, 3, 0, B.acc);
        }
        partial_out<<<1, 1, 0, s>>>(B.ctl, B.acc, take >= 0, (long long *)d_out);
    }
    bam->rows_indexed = true;
    CK(cudaGetLastError());
    return 0;
}

// second pass of the wrap-exact path: sum over the same records of (int32)((isize - mean) * (isize - mean)) (cluster.cpp:77)
extern "C" int svb_insert_sq(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t take, int32_t mean, int64_t *sq)
{
    if (!ctx || !bam || !sq) return svb_fail(ctx, SVB_ERR_ARG, "svb_insert_sq: null argument");
    CK(cudaSetDevice(ctx->device));
    CKR(prepare_rows(ctx, bam));
    cudaStream_t s = ctx->stream;
    const uint64_t n_chunks = bam->n_chunks;
    SvBuffers B{};
    {
        Bump measure(nullptr);
        carve(measure, B, n_chunks, 0, 0, 1, 1, bam->n_ref);
        CKR(ctx->ws_reserve(1, measure.used));
        Bump real(ctx->ws[1]);
        carve(real, B, n_chunks, 0, 0, 1, 1, bam->n_ref);
    }
    CK(cudaMemsetAsync(ctx->ws[1], 0, B.zero_end, s));
    const bool need_index = !bam->rows_indexed;
    if (need_index) CK(cudaMemsetAsync(bam->d_scal, 0, 16, s));
    RowsView V = view_of(bam);
    const unsigned g_chunks = nblk(n_chunks * 32, 256);
    if (need_index) rows_pass<true, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, min_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
    else rows_pass<false, true><<<g_chunks, 256, 0, s>>>(V, bam->d_fkey, bam->d_scal, min_mapq, B.q_cnt, B.q_sum, B.q_sq, B.ctl);
    bam->rows_indexed = true;
    QBaseOp op{n_chunks, B.q_cnt, B.q_base};
    launch_scan<1, 8>(ctx, s, op, B.sc_q, n_chunks);
    insert_ordered<<<g_chunks, 256, 0, s>>>(V, min_mapq, B.q_base, take < 0 ? ~0ull : (uint64_t)take, 2, mean, B.acc);
    long long v = 0;
    CK(cudaMemcpyAsync(&v, B.acc, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    *sq = v;
    return 0;
}

// pair support and window depth of this shard's own records with given statistics, one read-back; counts / depth_out may be
// device pointers (the ranks add them up with one NCCL all-reduce)
extern "C" int svb_pairs_depth(svb_ctx *ctx, svb_bam *bam, const svb_pair_params *p, const svb_junction *junctions, uint64_t n_j,
                               const svb_window *windows, uint64_t n_w, int32_t *counts, int32_t *depth_out)
{
    if (!ctx || !bam || !p || (n_j && (!junctions || !counts)) || (n_w && (!windows || !depth_out)))
        return svb_fail(ctx, SVB_ERR_ARG, "svb_pairs_depth: null argument");
    Request rq;
    rq.junctions = junctions, rq.n_j = n_j, rq.pp = *p, rq.counts = counts;
    rq.windows = windows, rq.n_w = n_w, rq.depth_mapq = p->min_mapq, rq.depth_out = depth_out;
    return run_passes(ctx, bam, rq);
}

extern "C" int svb_bam_set_own_offset(svb_bam *bam, uint64_t own_offset)
{
    if (!bam || own_offset > bam->nbytes) return SVB_ERR_ARG;
    bam->own_offset = own_offset;
    return 0;
}
