// getclip on the device: soft-clip candidate scan, breakpoint-key sort, per-key greedy clustering and
// text emission. Replaces the per-record loop of InputBamOutputReads (clip_reads.h:410-440) with its
// callees GetSClipReads / GetSeq / GenerateCigar / InsertSeq / ReadsInfo::ChangeSeqAndQual
// (clip_reads.cpp:57-108,112-192,260-329) and the writer DisplaySClipReadsAndClipFq (clip_reads.h:300-345).
//
// Round 2: the whole command is ONE stream-ordered sequence of launches. Every element count (candidates, segments, clusters,
// text bytes) stays in a device control block; buffers are capacity-bounded and carved from one workspace; kernels loop over
// device-side counts; the host reads the control block once at the end (overflow = run again with the sizes it reports).
// Round 1 synchronised nine times per call and the GPU sat idle between the pieces (profiles/r1_summary.md). Results stay in
// HBM until svb_clusters_text / svb_clusters_gz asks for them.
#include <algorithm>
#include <cstring>
#include <memory>

#include "walk.cuh"

// ---- device control block ------------------------------------------------------------------------------------------------------
struct ClipCtl {
    uint32_t counters[4];  // [0] clipped records [1] unmapped-branch records [2] chromosome switches [3] candidates
    uint32_t flags[4];     // [0] row slots overflowed [1] chain does not verify [2] a read reaches the next range shard's keys from outside its halo [3] spare
    uint64_t c64[2];       // records, end of the chain
    uint32_t n_seg, n_cl, un_skip, un_own;
    uint32_t abort_main, abort_side, tk_cluster, tk_text;  // tickets of the persistent warps of cluster_build / text_write
    uint64_t arena_bytes, clip_bytes, fq_bytes, un1_bytes, un2_bytes, export_bytes;
    uint64_t part_off[65];  // export_partitions: byte offset + 1 of every group that has records (0: none)
};

struct svb_clusters {
    svb_ctx *ctx = nullptr;
    char *d_text[4] = {nullptr, nullptr, nullptr, nullptr};  // device: the four texts (or nothing in gz mode once compressed)
    uint64_t text_len[4] = {0, 0, 0, 0}, d_cap[5] = {0, 0, 0, 0, 0};
    mutable PinnedBuf text[4];  // pinned host copies, made on first request
    mutable bool text_here[4] = {false, false, false, false};
    PinnedBuf gz[4];             // the same four files as gzip images, compressed on the device (svb_getclip_params.gz_outputs)
    bool gz_mode = false;
    uint8_t *d_export = nullptr;  // sharded runs: the raw unmapped-branch records (svb_getclip_params.export_unmapped_records)
    uint64_t export_len = 0;
    uint64_t part_off[65] = {};
    int32_t n_parts = 1;
    mutable PinnedBuf unmapped_records;
    mutable bool export_here = false;
    uint64_t n_clusters = 0, n_candidates = 0;
    void drop_device()
    {
        for (int w = 0; w < 4; ++w)
            if (d_text[w]) ctx->dev_put((uint8_t *)d_text[w], d_cap[w]), d_text[w] = nullptr;
        if (d_export) ctx->dev_put(d_export, d_cap[4]), d_export = nullptr;
    }
    int take(int w, uint64_t bytes)
    {
        uint8_t *p = ctx->dev_get(bytes, &d_cap[w]);
        if (!p) return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of device memory for the results", (unsigned long long)bytes);
        if (w < 4) d_text[w] = (char *)p;
        else d_export = p;
        return 0;
    }
    ~svb_clusters()
    {
        drop_device();
        for (auto &t : text) t.release(ctx);
        for (auto &t : gz) t.release(ctx);
        unmapped_records.release(ctx);
    }
};

// ---- candidate scan -------------------------------------------------------------------------------------
struct CandArrays {
    uint64_t *off;    // byte offset of the record in the stream (= its rank in file order)
    int32_t *tid;     // chromosome
    int32_t *pos;     // key position (1-based breakpoint)
    uint32_t *begin;  // first base of the "left" part inside the read
    uint32_t *ll;     // length of the left part (text before the breakpoint)
    uint32_t *rl;     // length of the right part
    uint8_t *side;    // 0 = '5' (left clipped), 1 = '3' (right clipped)
};

// bam_aux2i(bam_aux_get(b, "XC")) - clip_reads.cpp:126-127,158-159 - as the libbam the reference links behaves (probed with
// oracle/bamtool.c `auxi`, restated in oracle/bamio.py:aux_walk): the walk upper-cases a field's type before looking up its
// size and the size table knows only 'C'/'A' (1), 'S' (2), 'I' (4) in upper case, so a float or double field is NOT skipped -
// its value bytes are parsed as the next tag and an XC behind it is normally missed. Inside a 'B' array the raw sub-type is
// used ('f' known there) and the step is 32-bit int arithmetic. Bytes past the record read as 0 (the library would see stale
// buffer contents: undefined in the reference). Value: int32; non-integer types give 0.
__device__ int32_t aux_xc(const uint8_t *a, uint32_t n)
{
    auto at = [&](uint32_t i) -> uint32_t { return i < n ? a[i] : 0u; };
    uint32_t s = 0;
    while (s < n) {
        const bool hit = at(s) == 'X' && at(s + 1) == 'C';
        s += 2;
        if (hit) {
            const uint32_t ty = at(s);
            ++s;
            switch (ty) {
            case 'c': return (int8_t)at(s);
            case 'C': return (int32_t)at(s);
            case 's': return (int16_t)(at(s) | at(s + 1) << 8);
            case 'S': return (int32_t)(at(s) | at(s + 1) << 8);
            case 'i':
            case 'I': return (int32_t)(at(s) | at(s + 1) << 8 | at(s + 2) << 16 | at(s + 3) << 24);
            default: return 0;
            }
        }
        uint32_t u = at(s);
        if (u >= 'a' && u <= 'z') u -= 32;
        ++s;
        if (u == 'Z' || u == 'H') {
            while (s < n && a[s]) ++s;
            ++s;
        } else if (u == 'B') {
            const uint32_t sub = at(s), cnt = at(s + 1) | at(s + 2) << 8 | at(s + 3) << 16 | at(s + 4) << 24;
            const uint32_t sz = (sub == 'c' || sub == 'C' || sub == 'A') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
            const uint32_t step = 5u + cnt * sz;  // wraps like the library's int product
            if (step >= 0x80000000u) return 0;    // a backwards step leaves the record: undefined in the reference
            s += step;
        } else
            s += (u == 'C' || u == 'A') ? 1 : u == 'S' ? 2 : u == 'I' ? 4 : 0;
    }
    return 0;
}

struct ClipParams {
    int32_t min_mapq, save_low_quality;
    // range shards: only breakpoint keys (tid, pos) in [lo, hi) belong to this shard (whole file: everything)
    int32_t lo_tid, lo_pos, hi_tid, hi_pos;
    __device__ __forceinline__ bool owns(int32_t tid, int32_t pos) const
    {
        const bool ge_lo = tid > lo_tid || (tid == lo_tid && pos >= lo_pos);
        const bool lt_hi = tid < hi_tid || (tid == hi_tid && pos < hi_pos);
        return ge_lo && lt_hi;
    }
};

struct Emit {
    bool e5 = false, e3 = false, halo_miss = false;
    int32_t pos5 = 0, pos3 = 0;
    uint32_t b5 = 0, l5 = 0, r5 = 0, b3 = 0, l3 = 0, r3 = 0;
};

// GetSClipReads (clip_reads.cpp:112-192) for one mapped-branch record that survived the chromosome-switch test.
// Sequence and aux bytes are only touched for soft-clipped reads (~2 % of the records).
__device__ Emit eval_clip(const uint8_t *__restrict__ d, uint64_t o, const Core &k, const ClipParams &P)
{
    Emit E;
    if (k.n_cigar == 0) return E;
    const uint8_t *p = d + o;
    const uint8_t *cig = p + 36 + k.l_qname;
    uint32_t first = ldu32(cig), last = ldu32(cig + 4 * (k.n_cigar - 1));
    uint32_t op1 = first & 15, op2 = last & 15;
    if (op1 == OP_H || op2 == OP_H || (int32_t)k.mapq < P.min_mapq || (k.flag & F_DUP)) return E;  // clip_reads.cpp:118
    bool s1 = op1 == OP_S, s2 = op2 == OP_S;
    if (!s1 && !s2) return E;
    // GenerateCigar's l (clip_reads.cpp:322): M, D, =, N - X is not counted (quirk Q5)
    int32_t reflen = 0, eq_len = 0;
    for (uint32_t j = 0; j < k.n_cigar; ++j) {
        uint32_t w = ldu32(cig + 4 * j), op = w & 15;
        if (op == OP_M || op == OP_D || op == OP_EQ || op == OP_N) reflen += (int32_t)(w >> 4);
        if (op == OP_EQ) eq_len += (int32_t)(w >> 4);
    }
    const uint8_t *aux = cig + 4 * k.n_cigar + (k.l_qseq + 1) / 2 + k.l_qseq;
    int32_t xc = aux_xc(aux, (uint32_t)max((int64_t)0, (int64_t)(p + 4 + k.block_size - aux)));
    uint32_t len1 = first >> 4, len2 = last >> 4;
    if (s1 != s2) {
        if (xc != 0 && !P.save_low_quality) return E;
        if (s1) {
            if ((int64_t)len1 > k.l_qseq) return E;
            E.e5 = true, E.l5 = len1, E.r5 = k.l_qseq - len1;
        } else {
            if ((int64_t)len2 > k.l_qseq) return E;
            E.e3 = true, E.l3 = k.l_qseq - len2, E.r3 = len2;
        }
    } else {
        int64_t mid = (int64_t)k.l_qseq - len1 - len2;
        if (mid < 0 || k.n_cigar < 2) return E;  // (undefined in the reference: a CIGAR that is one S op)
        if (xc != 0 && !P.save_low_quality) {
            if (!(k.flag & F_REVERSE)) E.e5 = true;
            else E.e3 = true;
        } else
            E.e5 = E.e3 = true;
        E.l5 = len1, E.r5 = (uint32_t)mid;               // clip_reads.cpp:152,179
        E.b3 = len1, E.l3 = (uint32_t)mid, E.r3 = len2;  // clip_reads.cpp:154,185
    }
    E.pos5 = k.pos + 1, E.pos3 = k.pos + reflen;
    // Range shards: a '3' key at or beyond this shard's upper bound belongs to the next shard, which sees this record only if it
    // lies in its halo - and the halo comes from the .bai, whose alignment end (bam_calend: M, D, N) does not count the `=`
    // operations that GenerateCigar's length does. A record that ends, by libbam's count, at or before the start of the window in
    // front of the cut is not promised to the next shard: refuse loudly instead of dropping the cluster member (svb_getclip fails).
    if (E.e3 && eq_len > 0 && (k.tid > P.hi_tid || (k.tid == P.hi_tid && E.pos3 >= P.hi_pos)) &&
        k.pos + reflen - eq_len <= ((((P.hi_pos - 1) >> 14) - 1) << 14))
        E.halo_miss = true;
    E.e5 = E.e5 && P.owns(k.tid, E.pos5);
    E.e3 = E.e3 && P.owns(k.tid, E.pos3);
    return E;
}

// The first mapped-branch record of a chunk needs the last one of an earlier chunk (quirk Q1): one thread per chunk
__global__ void __launch_bounds__(128)
    clip_first(const uint8_t *__restrict__ d, uint64_t n_chunks, int32_t prev_tid0, ClipQueues q)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks || q.first_mb[c] == BAD_OFFSET) return;
    int32_t prev = prev_tid0;  // clip_reads.h:407: last_tid starts at 0 (or the previous shard's last tid)
    for (uint64_t j = c; j > 0;) {
        --j;
        if (q.last_mb_tid[j] != NO_TID) {
            prev = q.last_mb_tid[j];
            break;
        }
    }
    uint64_t o = q.first_mb[c];
    Core k = load_core(d + o);
    if (k.tid != prev) {
        uint32_t s = atomicAdd(&q.counters[2], 1u);
        if (s < q.sw_cap) q.switches[s] = o;
    } else if (k.n_cigar != 0 && (int32_t)k.mapq >= q.min_mapq && !(k.flag & F_DUP)) {
        const uint8_t *cig = d + o + 36 + k.l_qname;
        uint32_t op1 = ldu32(cig) & 15, op2 = ldu32(cig + 4 * (k.n_cigar - 1)) & 15;
        if (op1 != OP_H && op2 != OP_H && (op1 == OP_S || op2 == OP_S)) {
            uint32_t s = atomicAdd(&q.counters[0], 1u);
            if (s < q.clipped_cap) q.clipped[s] = o;
        }
    }
}

// the expensive part of GetSClipReads for the queued soft-clipped records: one thread each, one candidate append per warp
__global__ void __launch_bounds__(128)
    clip_eval(const uint8_t *__restrict__ d, ClipQueues q, ClipParams P, CandArrays c, uint32_t cand_cap, ClipCtl *ctl, uint32_t chunk_log2,
              uint32_t *__restrict__ bucket_cnt, uint32_t *__restrict__ arrival)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t n = min(q.counters[0], q.clipped_cap);
    if (blockIdx.x == 0 && threadIdx.x == 0 &&
        (q.counters[0] > q.clipped_cap || q.counters[1] > q.un_cap || q.counters[2] > q.sw_cap))
        ctl->abort_main = 1;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += stride) {
        const uint32_t i = base + lane;
        Emit E;
        uint64_t o = 0;
        int32_t tid = 0;
        if (i < n) {
            o = q.clipped[i];
            Core k = load_core(d + o);
            tid = k.tid;
            E = eval_clip(d, o, k, P);
        }
        if (E.halo_miss) ctl->flags[2] = 1;
        const uint32_t mine = (uint32_t)E.e5 + (uint32_t)E.e3;
        if (mine) {
            // cluster_build reads this record's bases and qualities next: ask L2 for those lines now (the head and the aux block
            // at the record's end have just been read; the middle of the record has not)
            const uint8_t *p = d + o;
            const uint32_t lq = ldu32(p + 12) & 0xff, nc = ldu32(p + 16) & 0xffff;
            const int32_t l = ldi32(p + 20);
            const uintptr_t a0 = (uintptr_t)(p + 36 + lq + 4 * nc) & ~(uintptr_t)127, a1 = (uintptr_t)(p + 36 + lq + 4 * nc + (l + 1) / 2 + l);
            for (uintptr_t a = a0; a < a1; a += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
        uint32_t incl = mine;
#pragma unroll
        for (int s = 1; s < 32; s <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, s);
            if (lane >= (uint32_t)s) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        if (!total) continue;
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(&q.counters[3], total);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        uint32_t s = b0 + incl - mine;
        // the candidates of a record join the bucket of the chunk the record starts in; `arrival` is their (arbitrary) place in it
        const uint32_t a0 = mine ? atomicAdd(&bucket_cnt[o >> chunk_log2], mine) : 0u;
        if (mine && s < cand_cap) arrival[s] = a0;
        if (mine == 2 && s + 1 < cand_cap) arrival[s + 1] = a0 + 1;
        if (E.e5 && s < cand_cap) {
            c.off[s] = o, c.tid[s] = tid, c.pos[s] = E.pos5, c.begin[s] = E.b5, c.ll[s] = E.l5, c.rl[s] = E.r5, c.side[s] = 0;
        }
        s += E.e5;
        if (E.e3 && s < cand_cap) {
            c.off[s] = o, c.tid[s] = tid, c.pos[s] = E.pos3, c.begin[s] = E.b3, c.ll[s] = E.l3, c.rl[s] = E.r3, c.side[s] = 1;
        }
    }
}

// ---- candidates in BAM order without a sort ---------------------------------------------------------------------------------------
// clip_eval appends candidates in arbitrary order; the clustering needs them in file order (InsertSeq is greedy, clip_reads.cpp:260-283),
// then stably by (flush run, side, position). Round 2 first sorted (record offset, index) pairs with four radix passes (68 us for
// 190 k candidates on C2: latency of the passes, not data). The record offset already says which 16 KiB chunk a candidate comes from
// and a chunk holds a handful of them (at most ~900: a record is at least 36 bytes), so file order = chunk order (a prefix sum over
// the per-chunk counts clip_eval left) + the order inside a chunk (every candidate counts the members of its bucket in front of it).
struct BucketScanOp {
    uint64_t n_chunks;
    const uint32_t *cnt;
    uint32_t *base;
    __device__ uint64_t n() const { return n_chunks; }
    __device__ void load(uint64_t c, uint64_t (&v)[1]) const { v[0] = cnt[c]; }
    __device__ void store(uint64_t c, const uint64_t (&excl)[1], const uint64_t (&)[1]) const { base[c] = (uint32_t)excl[0]; }
    __device__ void total(const uint64_t (&)[1]) const {}
};

// bucket members side by side (arbitrary order inside a bucket): grouped[base(chunk) + arrival] = candidate
__global__ void cand_group(const uint32_t *__restrict__ n_ptr, uint32_t cap, uint32_t chunk_log2, const uint64_t *__restrict__ off,
                           const uint32_t *__restrict__ bucket_base, const uint32_t *__restrict__ arrival, uint32_t *__restrict__ grouped, ClipCtl *ctl)
{
    if (*n_ptr > cap) {  // more candidates than slots: the call is repeated with the reported size, nothing below is used
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->abort_main = 1;
        return;
    }
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        grouped[bucket_base[off[i] >> chunk_log2] + arrival[i]] = i;
}

// sort key = flush run (number of chromosome switches before the record) | side | position, written at the candidate's place in
// file order: the stable sort that follows keeps that order inside a breakpoint key
static constexpr uint32_t SW_LINEAR = 1024;
__global__ void __launch_bounds__(256)
    make_keys(const uint32_t *__restrict__ counters, uint32_t cand_cap, uint32_t sw_cap, uint32_t chunk_log2, const uint32_t *__restrict__ grouped,
              const uint32_t *__restrict__ bucket_base, const uint32_t *__restrict__ bucket_cnt, CandArrays c,
              const uint64_t *__restrict__ sw_sorted, uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
    __shared__ uint64_t ssw[SW_LINEAR];
    if (counters[3] > cand_cap) return;
    const uint32_t n = counters[3], n_sw = min(counters[2], sw_cap);
    const bool linear = n_sw <= SW_LINEAR;  // few switches (a coordinate-sorted BAM has one per chromosome): count them directly
    if (linear) {
        for (uint32_t j = threadIdx.x; j < n_sw; j += blockDim.x) ssw[j] = sw_sorted[j];
        __syncthreads();
    }
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t s = grouped[p];
        const uint64_t r = c.off[s];
        const uint32_t sd = c.side[s];
        // place inside the bucket: members with a smaller record offset, and the '5' candidate of the same record (clip_eval's
        // order for a read clipped on both sides)
        const uint64_t ch = r >> chunk_log2;
        const uint32_t b0 = bucket_base[ch], b1 = b0 + bucket_cnt[ch];
        uint32_t before = 0;
        for (uint32_t q = b0; q < b1; ++q) {
            const uint32_t t = grouped[q];
            const uint64_t rt = c.off[t];
            before += (rt < r) || (rt == r && c.side[t] < sd);
        }
        uint32_t run = 0;
        if (linear) {
            for (uint32_t j = 0; j < n_sw; ++j) run += ssw[j] < r;
        } else {
            uint32_t lo = 0, hi = n_sw;  // lower_bound(sw, r): switches with offset < r
            while (lo < hi) {
                uint32_t m = (lo + hi) >> 1;
                if (sw_sorted[m] < r) lo = m + 1;
                else hi = m;
            }
            run = lo;
        }
        key[b0 + before] = ((uint64_t)run << 33) | ((uint64_t)sd << 32) | (uint32_t)(c.pos[s] ^ 0x80000000);
        val[b0 + before] = s;
    }
}

// ---- segments (one per breakpoint key) ---------------------------------------------------------------------------------------
struct SegScanOp {
    const uint32_t *counters;
    uint32_t cand_cap;
    const uint64_t *key;
    uint32_t *start;
    ClipCtl *ctl;
    __device__ uint64_t n() const { return ctl->abort_main ? 0 : min(counters[3], cand_cap); }
    __device__ void load(uint64_t i, uint64_t (&v)[1]) const { v[0] = (i == 0 || key[i] != key[i - 1]) ? 1 : 0; }
    __device__ void store(uint64_t i, const uint64_t (&excl)[1], const uint64_t (&v)[1]) const
    {
        if (v[0]) start[excl[0]] = (uint32_t)i;
    }
    __device__ void total(const uint64_t (&t)[1]) const
    {
        ctl->n_seg = (uint32_t)t[0];
        start[t[0]] = (uint32_t)n();
    }
};

// per segment: arena bytes = members * (longest left + longest right)
struct SegStatsOp {
    const uint32_t *start, *order;
    CandArrays c;
    uint32_t *maxl, *maxr;
    uint64_t *arena_off;
    uint64_t arena_cap;
    ClipCtl *ctl;
    __device__ uint64_t n() const { return ctl->abort_main ? 0 : ctl->n_seg; }
    __device__ void load(uint64_t s, uint64_t (&v)[1]) const
    {
        uint32_t ml = 0, mr = 0;
        for (uint32_t k = start[s]; k < start[s + 1]; ++k) {
            uint32_t x = order[k];
            ml = max(ml, c.ll[x]);
            mr = max(mr, c.rl[x]);
        }
        maxl[s] = ml, maxr[s] = mr;
        // (a key with ONE member needs no arena: nothing is compared or merged, the text is decoded straight from the record)
        const uint32_t members = start[s + 1] - start[s];
        v[0] = members > 1 ? (uint64_t)members * (ml + mr) : 0;
    }
    __device__ void store(uint64_t s, const uint64_t (&excl)[1], const uint64_t (&)[1]) const { arena_off[s] = excl[0]; }
    __device__ void total(const uint64_t (&t)[1]) const
    {
        ctl->arena_bytes = t[0];
        if (t[0] > arena_cap) ctl->abort_main = 1;
    }
};

struct ClusterOut {
    uint32_t *len_l, *len_r, *support;  // indexed by sorted candidate position (seg start + k)
    uint64_t *cig_off;                  // record whose CIGAR the cluster carries
    uint8_t *noqual;
    uint32_t *seg_ncl;
};

// 98 % of the breakpoint keys hold a single soft-clipped read (C2: 89.9 k of 91.5 k): InsertSeq (clip_reads.cpp:260-283) has
// nothing to compare it with, so its cluster IS the read. One thread per such key fills the cluster's fields; its strings are
// never staged - text_write decodes them from the record. (Until the third session of round 2 these keys went through the warp-
// sequential kernel below like the others and were most of its 0.17 ms.)
__global__ void __launch_bounds__(128)
    cluster_singletons(const uint8_t *__restrict__ d, const ClipCtl *__restrict__ ctl, const uint32_t *__restrict__ start,
                       const uint32_t *__restrict__ order, CandArrays c, ClusterOut out)
{
    if (ctl->abort_main) return;
    const uint32_t n_seg = ctl->n_seg;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < n_seg; s += gridDim.x * blockDim.x) {
        const uint32_t a = start[s];
        if (start[s + 1] - a != 1) continue;
        const uint32_t x = order[a];
        const uint64_t rec = c.off[x];
        const uint8_t *p = d + rec;
        const uint32_t w = ldu32(p + 12), w2 = ldu32(p + 16);
        const int32_t l_qseq = ldi32(p + 20);
        const uint8_t *qual = p + 36 + (w & 0xff) + 4 * (w2 & 0xffff) + (l_qseq + 1) / 2;
        out.len_l[a] = c.ll[x], out.len_r[a] = c.rl[x], out.cig_off[a] = rec, out.support[a] = 1;
        out.noqual[a] = l_qseq > 0 && qual[0] == 0xff;
        out.seg_ncl[s] = 1;
    }
}

// One warp per breakpoint key: the reference's sequential greedy InsertSeq (clip_reads.cpp:260-283) in
// BAM order, with lane-parallel string compares and consensus updates. Strings live in a per-segment
// arena: slot k holds [left part right-aligned at column maxl | right part left-aligned at maxl].
__global__ void __launch_bounds__(128)
    cluster_build(const uint8_t *__restrict__ d, ClipCtl *__restrict__ ctl,
                  const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, CandArrays c,
                  const uint32_t *__restrict__ maxl_, const uint32_t *__restrict__ maxr_, const uint64_t *__restrict__ arena_off,
                  char *__restrict__ arena_seq, char *__restrict__ arena_qual, double limit, ClusterOut out)
{
    if (ctl->abort_main) return;
    const uint32_t n_seg = ctl->n_seg, lane = threadIdx.x & 31;
    const uint32_t n_members = start[n_seg];
    // What a member needs before its bases can be fetched is a chain of dependent loads - order -> candidate fields -> record
    // head: three round trips per member when the warp walks the members one by one, and that latency, not occupancy, is what
    // bounded this kernel (0.175 ms at 32, 40 or 56 registers alike). The lanes fetch the chain for 32 consecutive members at
    // once; the sequential greedy loop takes each member's values from the lane that holds them.
    uint32_t m_base = 0xffffffffu, m_x = 0, m_begin = 0, m_ll = 0, m_rl = 0, m_side = 0, m_w = 0, m_w2 = 0;
    int32_t m_lqseq = 0;
    uint64_t m_rec = 0;
    // warps take sixteen keys (~32 members: one fetch of member chains) at a time from a ticket counter: a key with many reads (a real breakpoint) keeps its warp busy for a
    // while, and a fixed assignment left the last warps running alone
    for (;;) {
        uint32_t s0 = 0;
        if (lane == 0) s0 = atomicAdd(&ctl->tk_cluster, 16u);
        s0 = __shfl_sync(0xffffffffu, s0, 0);
        if (s0 >= n_seg) break;
        const uint32_t s1 = min(s0 + 16u, n_seg);
        // the keys' own fields, one key per lane
        uint32_t g_a = 0, g_b = 0, g_maxl = 0, g_maxr = 0;
        uint64_t g_off = 0;
        if (s0 + lane < s1) {
            const uint32_t sg = s0 + lane;
            g_a = start[sg], g_b = start[sg + 1], g_maxl = maxl_[sg], g_maxr = maxr_[sg], g_off = arena_off[sg];
        }
      for (uint32_t s = s0; s < s1; ++s) {
        const uint32_t a = __shfl_sync(0xffffffffu, g_a, s - s0), b = __shfl_sync(0xffffffffu, g_b, s - s0);
        const uint32_t maxl = __shfl_sync(0xffffffffu, g_maxl, s - s0), stride = maxl + __shfl_sync(0xffffffffu, g_maxr, s - s0);
        const uint64_t seg_off = __shfl_sync(0xffffffffu, g_off, s - s0);
        if (b - a == 1) continue;  // (cluster_singletons)
        char *S = arena_seq + seg_off, *Q = arena_qual + seg_off;
        uint32_t ncl = 0;
        for (uint32_t k = a; k < b; ++k) {
            if (m_base == 0xffffffffu || k < m_base || k >= m_base + 32) {  // (warp-uniform) the next 32 members' chains
                m_base = k;
                const uint32_t kk = k + lane;
                if (kk < n_members) {
                    m_x = order[kk];
                    m_rec = c.off[m_x], m_begin = c.begin[m_x], m_ll = c.ll[m_x], m_rl = c.rl[m_x], m_side = c.side[m_x];
                    const uint8_t *hp = d + m_rec;
                    m_w = ldu32(hp + 12), m_w2 = ldu32(hp + 16), m_lqseq = ldi32(hp + 20);
                }
            }
            const int src = (int)(k - m_base);
            const uint32_t x = __shfl_sync(0xffffffffu, m_x, src);
            const uint64_t rec = __shfl_sync(0xffffffffu, m_rec, src);
            const uint32_t begin = __shfl_sync(0xffffffffu, m_begin, src), ll = __shfl_sync(0xffffffffu, m_ll, src);
            const uint32_t rl = __shfl_sync(0xffffffffu, m_rl, src), side = __shfl_sync(0xffffffffu, m_side, src);
            const uint32_t w = __shfl_sync(0xffffffffu, m_w, src), w2 = __shfl_sync(0xffffffffu, m_w2, src);
            const int32_t l_qseq = __shfl_sync(0xffffffffu, m_lqseq, src);
            const uint8_t *p = d + rec;
            (void)x;
            const uint8_t *seq = p + 36 + (w & 0xff) + 4 * (w2 & 0xffff);
            const uint8_t *qual = seq + (l_qseq + 1) / 2;
            const uint8_t q_first = l_qseq > 0 ? qual[0] : 0;
            char *tS = S + (uint64_t)ncl * stride, *tQ = Q + (uint64_t)ncl * stride;  // tentative new cluster
            // GetSeq (clip_reads.cpp:286-306): 4-bit codes -> "=ACMGRSVTWYHKDBN", quality + 33. All loads of (up to) 256 bases are
            // issued before the first store: with one load -> store per loop iteration every 32 bases waited for their own round trip
            // to the record (ncu: 44 % of the kernel's stall samples sat on these two loads).
            const uint32_t n_bases = ll + rl;
            bool noq = false;
            for (uint32_t j0 = 0; j0 < n_bases; j0 += 256) {
                uint8_t sb[8], qb[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t j = j0 + u * 32 + lane;
                    sb[u] = 0, qb[u] = 0;
                    if (j < n_bases) sb[u] = seq[(begin + j) >> 1], qb[u] = qual[begin + j];
                }
                noq = l_qseq > 0 && q_first == 0xff;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t j = j0 + u * 32 + lane;
                    if (j < n_bases) {
                        const uint32_t idx = begin + j, nib = (sb[u] >> ((~idx & 1) << 2)) & 15, col = maxl - ll + j;
                        // "=ACMGRSV" / "TWYHKDBN" packed into two registers (an indexed constant load serialises over distinct indices)
                        const uint64_t tab = (nib & 8) ? 0x4E42444B48595754ull : 0x565352474D43413Dull;
                        tS[col] = (char)(tab >> ((nib & 7) * 8));
                        tQ[col] = noq ? '*' : (char)(qb[u] + 33);
                    }
                }
            }
            noq = l_qseq > 0 && q_first == 0xff;
            __syncwarp();
            int found = -1;
            for (uint32_t cl = 0; cl < ncl; ++cl) {
                const char *cS = S + (uint64_t)cl * stride;
                uint32_t cL = out.len_l[a + cl], cR = out.len_r[a + cl];
                uint32_t n1 = min(ll, cL), n2 = min(rl, cR);
                uint32_t m1 = 0, m2 = 0;
                for (uint32_t j = lane; j < n1; j += 32) m1 += tS[maxl - 1 - j] == cS[maxl - 1 - j];  // CompareStringEndFirst
                for (uint32_t j = lane; j < n2; j += 32) m2 += tS[maxl + j] == cS[maxl + j];          // CompareStringBeginFirst
                m1 = warp_sum(m1);
                m2 = warp_sum(m2);
                // (double)match/len >= limit; len == 0 gives NaN -> false (clip_reads.cpp:204,216)
                bool ok = n1 > 0 && n2 > 0 && (double)m1 / (double)n1 >= limit && (double)m2 / (double)n2 >= limit;
                if (ok) {
                    found = (int)cl;
                    break;
                }
            }
            if (found < 0) {
                if (lane == 0) {
                    out.len_l[a + ncl] = ll, out.len_r[a + ncl] = rl, out.cig_off[a + ncl] = rec, out.support[a + ncl] = 1;
                    out.noqual[a + ncl] = noq;
                }
                ++ncl;
            } else {
                // ReadsInfo::ChangeSeqAndQual (clip_reads.cpp:57-108)
                char *cS = S + (uint64_t)found * stride, *cQ = Q + (uint64_t)found * stride;
                uint32_t cL = out.len_l[a + found], cR = out.len_r[a + found];
                bool cnoq = out.noqual[a + found];
                uint32_t n1 = min(ll, cL), n2 = min(rl, cR);
                if (!noq && !cnoq) {  // (with a missing quality string the reference indexes out of bounds)
                    for (uint32_t j = lane; j < n1; j += 32) {
                        uint32_t col = maxl - 1 - j;
                        if (cQ[col] < tQ[col]) cQ[col] = tQ[col], cS[col] = tS[col];
                    }
                    for (uint32_t j = lane; j < n2; j += 32) {
                        uint32_t col = maxl + j;
                        if (cQ[col] < tQ[col]) cQ[col] = tQ[col], cS[col] = tS[col];
                    }
                }
                if (cL <= ll) {  // extend to the longer left part; right-clipped clusters take the new CIGAR
                    for (uint32_t j = lane; j < ll - cL; j += 32) {
                        uint32_t col = maxl - ll + j;
                        cS[col] = tS[col], cQ[col] = tQ[col];
                    }
                }
                if (cR < rl) {
                    for (uint32_t j = lane; j < rl - cR; j += 32) {
                        uint32_t col = maxl + cR + j;
                        cS[col] = tS[col], cQ[col] = tQ[col];
                    }
                }
                if (lane == 0) {
                    if (cL <= ll) {
                        out.len_l[a + found] = ll;
                        if (side == 1) out.cig_off[a + found] = rec;
                    }
                    if (cR < rl) {
                        out.len_r[a + found] = rl;
                        if (side == 0) out.cig_off[a + found] = rec;
                    }
                    out.support[a + found] += 1;
                }
            }
            __syncwarp();
        }
        if (lane == 0) out.seg_ncl[s] = ncl;
      }
    }
}

// cluster slot -> flat cluster list
struct ClusterScanOp {
    const uint32_t *seg_ncl;
    uint32_t *cl_seg, *cl_slot;
    ClipCtl *ctl;
    __device__ uint64_t n() const { return ctl->abort_main ? 0 : ctl->n_seg; }
    __device__ void load(uint64_t s, uint64_t (&v)[1]) const { v[0] = seg_ncl[s]; }
    __device__ void store(uint64_t s, const uint64_t (&excl)[1], const uint64_t (&v)[1]) const
    {
        for (uint32_t k = 0; k < (uint32_t)v[0]; ++k) cl_seg[excl[0] + k] = (uint32_t)s, cl_slot[excl[0] + k] = k;
    }
    __device__ void total(const uint64_t (&t)[1]) const { ctl->n_cl = (uint32_t)t[0]; }
};

// ---- text emission --------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dec_len(uint32_t v)
{
    uint32_t n = 1;
    while (v >= 10) v /= 10, ++n;
    return n;
}
__device__ __forceinline__ uint32_t dec_len_i(int32_t v) { return v < 0 ? 1 + dec_len((uint32_t)(-(int64_t)v)) : dec_len((uint32_t)v); }
__device__ __forceinline__ char *put_dec(char *o, uint32_t v)
{
    uint32_t n = dec_len(v);
    for (uint32_t i = n; i-- > 0;) o[i] = '0' + v % 10, v /= 10;
    return o + n;
}
__device__ __forceinline__ char *put_dec_i(char *o, int32_t v)
{
    if (v < 0) {
        *o++ = '-';
        return put_dec(o, (uint32_t)(-(int64_t)v));
    }
    return put_dec(o, (uint32_t)v);
}

struct NameTable {
    const char *blob;
    const uint32_t *off;  // n_ref + 1
};

// cigar text of a record without its S/H ops (GenerateCigar + DisplayCigarVector, clip_reads.cpp:309-329,
// clip_reads.h:489-505)
__device__ uint32_t cigar_text(const uint8_t *p, char *o)
{
    uint32_t w = ldu32(p + 12), n_cigar = ldu32(p + 16) & 0xffff;
    const uint8_t *cig = p + 36 + (w & 0xff);
    uint32_t len = 0;
    for (uint32_t j = 0; j < n_cigar; ++j) {
        uint32_t x = ldu32(cig + 4 * j), op = x & 15;
        if (op == OP_H || op == OP_S) continue;
        if (o) {
            char *e = put_dec(o + len, x >> 4);
            *e = "MIDNSHP=X???????"[op];
            len = (uint32_t)(e - o) + 1;
        } else
            len += dec_len(x >> 4) + 1;
    }
    return len;
}

struct TextScanOp {
    const uint32_t *cl_seg, *cl_slot, *start, *order;
    CandArrays c;
    ClusterOut out;
    const uint8_t *d;
    NameTable names;
    uint64_t *clip_off, *fq_off;
    uint64_t clip_cap, fq_cap;
    ClipCtl *ctl;
    __device__ uint64_t n() const { return ctl->abort_main ? 0 : ctl->n_cl; }
    __device__ void load(uint64_t i, uint64_t (&v)[2]) const
    {
        uint32_t s = cl_seg[i], a = start[s], k = a + cl_slot[i];
        uint32_t x = order[a];
        uint32_t L = out.len_l[k], R = out.len_r[k];
        uint32_t qL = out.noqual[k] ? 1 : L, qR = out.noqual[k] ? 1 : R;
        uint32_t name = names.off[c.tid[x] + 1] - names.off[c.tid[x]];
        uint32_t cg = cigar_text(d + out.cig_off[k], nullptr);
        // chr \t pos \t side \t cigar \t aligned \t alignedQ \t clipped \t clippedQ \t support \n
        v[0] = name + 1 + dec_len_i(c.pos[x]) + 1 + 2 + cg + 1 + L + 1 + qL + 1 + R + 1 + qR + 1 + dec_len(out.support[k]) + 1;
        uint32_t cl = c.side[x] == 0 ? L : R, cq = c.side[x] == 0 ? qL : qR;
        v[1] = 1 + cl + 1 + cl + 1 + 2 + cq + 1;  // @seq \n seq \n + \n qual \n
    }
    __device__ void store(uint64_t i, const uint64_t (&excl)[2], const uint64_t (&)[2]) const { clip_off[i] = excl[0], fq_off[i] = excl[1]; }
    __device__ void total(const uint64_t (&t)[2]) const
    {
        ctl->clip_bytes = t[0], ctl->fq_bytes = t[1];
        if (t[0] > clip_cap || t[1] > fq_cap) ctl->abort_main = 1;
    }
};

// One thread per cluster: everything of its clip.gz line that is not a copy of bases or qualities - the head
// "chr \t pos \t side \t cigar \t", the tabs between the four strings, the tail "\t support \n" - and the frame bytes of its FASTQ
// record. In the warp-per-cluster writer below this was lane-0 code: two thirds of that kernel's instructions ran with one
// active lane (ncu: 11.5 threads per instruction).
__global__ void __launch_bounds__(128)
    text_heads(const ClipCtl *__restrict__ ctl, const uint32_t *__restrict__ cl_seg, const uint32_t *__restrict__ cl_slot,
               const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, CandArrays c, ClusterOut out,
               const uint8_t *__restrict__ d, NameTable names, const uint64_t *__restrict__ clip_off, const uint64_t *__restrict__ fq_off,
               char *__restrict__ clip, char *__restrict__ fq)
{
    if (ctl->abort_main) return;
    const uint32_t n_cl = ctl->n_cl;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cl; i += gridDim.x * blockDim.x) {
        const uint32_t s = cl_seg[i], a = start[s], k = a + cl_slot[i], x = order[a];
        const uint32_t L = out.len_l[k], R = out.len_r[k], side = c.side[x];
        const bool noq = out.noqual[k];
        const uint32_t aN = side == 0 ? R : L, cN = side == 0 ? L : R, aQN = noq ? 1 : aN, cQN = noq ? 1 : cN;
        char *q = clip + clip_off[i];
        const int32_t tid = c.tid[x];
        const char *nm = names.blob + names.off[tid];
        const uint32_t nl = names.off[tid + 1] - names.off[tid];
        for (uint32_t j = 0; j < nl; ++j) *q++ = nm[j];
        *q++ = '\t';
        q = put_dec_i(q, c.pos[x]);
        *q++ = '\t';
        *q++ = side == 0 ? '5' : '3';
        *q++ = '\t';
        q += cigar_text(d + out.cig_off[k], q);
        *q++ = '\t';
        q += aN, *q++ = '\t';
        q += aQN, *q++ = '\t';
        q += cN, *q++ = '\t';
        q += cQN, *q++ = '\t';
        q = put_dec(q, out.support[k]);
        *q++ = '\n';
        char *f = fq + fq_off[i];  // FASTQ record named by its own sequence (clip_reads.h:320,339)
        f[0] = '@', f[1 + cN] = '\n', f[2 + 2 * cN] = '\n', f[3 + 2 * cN] = '+', f[4 + 2 * cN] = '\n', f[5 + 2 * cN + cQN] = '\n';
    }
}

// One warp per cluster writes the four strings of its clip.gz line and of its FASTQ record into the places text_heads leaves free
// (the two kernels write disjoint bytes and run side by side on two streams; the head's length follows from the line's size).
// Source of the strings: the cluster's arena slot, or - for a key with a single member - the record itself (4-bit bases ->
// "=ACMGRSVTWYHKDBN", quality + 33; GetSeq, clip_reads.cpp:286-306). The lanes fetch the chains of dependent loads that lead to
// the strings (cluster -> key -> candidate -> record head) for the eight clusters of a ticket side by side.
__device__ __forceinline__ char base_char(uint32_t packed_byte, uint32_t idx)
{
    const uint32_t nib = (packed_byte >> ((~idx & 1) << 2)) & 15;
    const uint64_t tab = (nib & 8) ? 0x4E42444B48595754ull : 0x565352474D43413Dull;  // "TWYHKDBN" : "=ACMGRSV"
    return (char)(tab >> ((nib & 7) * 8));
}

__global__ void __launch_bounds__(128)
    text_write(ClipCtl *__restrict__ ctl, const uint32_t *__restrict__ cl_seg, const uint32_t *__restrict__ cl_slot,
               const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, CandArrays c, ClusterOut out, const uint8_t *__restrict__ d,
               const uint32_t *__restrict__ maxl_, const uint32_t *__restrict__ maxr_, const uint64_t *__restrict__ arena_off,
               const char *__restrict__ arena_seq, const char *__restrict__ arena_qual, const uint64_t *__restrict__ clip_off,
               const uint64_t *__restrict__ fq_off, char *__restrict__ clip, char *__restrict__ fq)
{
    if (ctl->abort_main) return;
    const uint32_t n_cl = ctl->n_cl, lane = threadIdx.x & 31;
    for (;;) {
        uint32_t i0 = 0;
        if (lane == 0) i0 = atomicAdd(&ctl->tk_text, 8u);
        i0 = __shfl_sync(0xffffffffu, i0, 0);
        if (i0 >= n_cl) break;
        const uint32_t cnt = min(8u, n_cl - i0);
        // lane j < cnt: everything cluster i0 + j needs
        uint32_t m_single = 0, m_L = 0, m_R = 0, m_side = 0, m_noq = 0, m_head = 0, m_aAt = 0, m_cAt = 0;
        int32_t m_lqseq = 0;
        uint64_t m_src = 0, m_line = 0, m_fq = 0;
        if (lane < cnt) {
            const uint32_t i = i0 + lane;
            const uint32_t sg = cl_seg[i], a = start[sg], slot = cl_slot[i], k = a + slot, x = order[a];
            m_single = start[sg + 1] - a == 1;
            m_L = out.len_l[k], m_R = out.len_r[k], m_side = c.side[x], m_noq = out.noqual[k];
            const uint32_t aN = m_side == 0 ? m_R : m_L, cN = m_side == 0 ? m_L : m_R;
            const uint32_t aQN = m_noq ? 1 : aN, cQN = m_noq ? 1 : cN;
            m_line = clip_off[i], m_fq = fq_off[i];
            const uint64_t line_end = i + 1 < n_cl ? clip_off[i + 1] : ctl->clip_bytes;
            // line = head | aligned \t alignedQ \t clipped \t clippedQ \t support \n
            m_head = (uint32_t)(line_end - m_line) - (aN + aQN + cN + cQN + 5 + dec_len(out.support[k]));
            // '5': aligned = right part, clipped = left part; '3': aligned = left, clipped = right (clip_reads.h:308-332)
            if (m_single) {  // bases [begin, begin + L) are the left part, [begin + L, begin + L + R) the right part
                const uint64_t rec = out.cig_off[k];
                const uint8_t *p = d + rec;
                const uint32_t w = ldu32(p + 12), w2 = ldu32(p + 16);
                m_lqseq = ldi32(p + 20);
                m_src = rec + 36 + (w & 0xff) + 4 * (w2 & 0xffff);  // the packed bases; the qualities follow them
                const uint32_t begin = c.begin[x];
                m_aAt = m_side == 0 ? begin + m_L : begin, m_cAt = m_side == 0 ? begin : begin + m_L;
            } else {  // slot = [left part right-aligned at column maxl | right part from column maxl]
                const uint32_t maxl = maxl_[sg], stride = maxl + maxr_[sg];
                m_src = arena_off[sg] + (uint64_t)slot * stride;
                m_aAt = m_side == 0 ? maxl : maxl - m_L, m_cAt = m_side == 0 ? maxl - m_L : maxl;
            }
        }
        for (uint32_t j = 0; j < cnt; ++j) {
            const bool single = __shfl_sync(0xffffffffu, m_single, j) != 0, noq = __shfl_sync(0xffffffffu, m_noq, j) != 0;
            const uint32_t L = __shfl_sync(0xffffffffu, m_L, j), R = __shfl_sync(0xffffffffu, m_R, j), side = __shfl_sync(0xffffffffu, m_side, j);
            const uint32_t head = __shfl_sync(0xffffffffu, m_head, j), aAt = __shfl_sync(0xffffffffu, m_aAt, j), cAt = __shfl_sync(0xffffffffu, m_cAt, j);
            const int32_t l_qseq = __shfl_sync(0xffffffffu, m_lqseq, j);
            const uint64_t src = __shfl_sync(0xffffffffu, m_src, j), line_begin = __shfl_sync(0xffffffffu, m_line, j);
            const uint64_t fq_begin = __shfl_sync(0xffffffffu, m_fq, j);
            const uint32_t aN = side == 0 ? R : L, cN = side == 0 ? L : R, aQN = noq ? 1 : aN;
            // bytes of the bases and of the qualities, indexed by base number (record) or by column (arena slot)
            const uint8_t *sB = single ? d + src : (const uint8_t *)arena_seq + src;
            const uint8_t *qB = single ? d + src + (l_qseq + 1) / 2 : (const uint8_t *)arena_qual + src;
            char *qa = clip + line_begin + head, *qaq = qa + aN + 1, *qc = qaq + aQN + 1, *qcq = qc + cN + 1;
            char *f = fq + fq_begin;
            // bases and qualities of a part are fetched together, up to 256 of each per lane-round, before anything is stored: one
            // exposed round trip per part instead of one per 32 bytes and string; the clipped part is stored three times (clip
            // line, FASTQ name, FASTQ sequence) from one fetch
            for (uint32_t j0 = 0; j0 < aN; j0 += 256) {
                uint8_t s8[8], q8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t t = j0 + u * 32 + lane;
                    s8[u] = 0, q8[u] = 0;
                    if (t < aN) {
                        s8[u] = single ? sB[(aAt + t) >> 1] : sB[aAt + t];
                        if (!noq) q8[u] = qB[aAt + t];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t t = j0 + u * 32 + lane;
                    if (t < aN) {
                        qa[t] = single ? base_char(s8[u], aAt + t) : (char)s8[u];
                        if (!noq) qaq[t] = single ? (char)(q8[u] + 33) : (char)q8[u];
                    }
                }
            }
            for (uint32_t j0 = 0; j0 < cN; j0 += 256) {
                uint8_t s8[8], q8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t t = j0 + u * 32 + lane;
                    s8[u] = 0, q8[u] = 0;
                    if (t < cN) {
                        s8[u] = single ? sB[(cAt + t) >> 1] : sB[cAt + t];
                        if (!noq) q8[u] = qB[cAt + t];
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t t = j0 + u * 32 + lane;
                    if (t < cN) {
                        const char ch = single ? base_char(s8[u], cAt + t) : (char)s8[u];
                        qc[t] = ch, f[1 + t] = ch, f[2 + cN + t] = ch;
                        if (!noq) {
                            const char qh = single ? (char)(q8[u] + 33) : (char)q8[u];
                            qcq[t] = qh, f[5 + 2 * cN + t] = qh;
                        }
                    }
                }
            }
            if (noq && lane == 0) qaq[0] = '*', qcq[0] = '*', f[5 + 2 * cN] = '*';
        }
    }
}

// ---- unmapped-branch records: StoreUnmapSeqAndQual (clip_reads.h:172-219) on the device --------------------------
// The reference keeps a std::map<qname, held mate> over the whole file: the first record of a name is held; a later
// record of the same name and the OTHER end emits the pair (read1 to file 1, read2 to file 2) and erases the entry; a
// later record of the same end is ignored. Output order = file order of the completing record. Here: sort the queued
// records by offset, hash the names, stable-sort entries by hash (entries stay in file order inside a group), run the tiny
// per-name automaton with one thread per hash group (real name compares, so hash collisions cannot change the result), then
// emit text. The whole branch runs on a side stream next to the candidate pipeline.
__device__ __forceinline__ uint32_t qname_len(const uint8_t *p) { return ldu32(p + 12) & 0xff; }

// the unmapped list is sorted by offset: records in front of the shard's own region (range shards) only lend their soft clips
__global__ void unmapped_own(const uint32_t *__restrict__ counters, uint32_t un_cap, const uint64_t *__restrict__ sorted, uint64_t halo_bytes,
                             ClipCtl *ctl)
{
    if (counters[1] > un_cap) ctl->abort_side = 1;
    const uint32_t n = min(counters[1], un_cap);
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        if (sorted[m] < halo_bytes) lo = m + 1;
        else hi = m;
    }
    ctl->un_skip = lo, ctl->un_own = n - lo;
}

__global__ void unmapped_hash(const ClipCtl *__restrict__ ctl, const uint64_t *__restrict__ sorted, const uint8_t *__restrict__ d,
                              uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
    if (ctl->abort_side) return;
    const uint32_t n = ctl->un_own;
    const uint64_t *list = sorted + ctl->un_skip;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint8_t *p = d + list[i];
        uint32_t lq = qname_len(p);
        uint64_t h = 0xcbf29ce484222325ull;
        for (uint32_t j = 0; j < lq && p[36 + j]; ++j) h = (h ^ p[36 + j]) * 0x100000001b3ull;
        key[i] = h & 0xffffffffu, val[i] = i;  // 32 bits order the groups; the names themselves decide inside a group
    }
}

__device__ bool same_name(const uint8_t *a, const uint8_t *b)
{
    uint32_t la = qname_len(a), lb = qname_len(b);
    for (uint32_t j = 0;; ++j) {
        uint8_t x = j < la ? a[36 + j] : 0, y = j < lb ? b[36 + j] : 0;
        if (x != y) return false;
        if (x == 0) return true;
    }
}

// one thread per hash group; mate_of[e] = entry that was held when e completed a pair, else 0xffffffff. Up to H names of a
// group are tracked in registers; a group with more distinct names at once (a 64-bit hash collision of more than eight names)
// switches to replaying the group's history for every further entry - exact, only slower.
__global__ void unmapped_pair(const ClipCtl *__restrict__ ctl, const uint64_t *__restrict__ key, const uint32_t *__restrict__ ent,
                               const uint64_t *__restrict__ sorted, const uint8_t *__restrict__ d, uint32_t *__restrict__ mate_of)
{
    if (ctl->abort_side) return;
    const uint32_t n = ctl->un_own;
    const uint64_t *list = sorted + ctl->un_skip;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        if (j > 0 && key[j] == key[j - 1]) continue;  // (keys are the low 32 bits of the name hash)
        const int H = 8;
        uint32_t held[H];
        int nh = 0;
        bool replay = false;
        for (uint32_t k = j; k < n && key[k] == key[j]; ++k) {
            uint32_t e = ent[k];
            const uint8_t *p = d + list[e];
            bool r1 = (ldu32(p + 16) >> 16) & F_READ1;
            if (replay) {  // state of this name from the group's history: held entry or none
                uint32_t h = 0xffffffffu;
                bool h1 = false;
                for (uint32_t k2 = j; k2 < k; ++k2) {
                    const uint8_t *q = d + list[ent[k2]];
                    if (!same_name(p, q)) continue;
                    bool q1 = (ldu32(q + 16) >> 16) & F_READ1;
                    if (h == 0xffffffffu) h = ent[k2], h1 = q1;
                    else if (q1 != h1) h = 0xffffffffu;
                }
                if (h != 0xffffffffu && h1 != r1) mate_of[e] = h;
                continue;
            }
            int hit = -1;
            for (int h = 0; h < nh; ++h)
                if (same_name(p, d + list[held[h]])) {
                    hit = h;
                    break;
                }
            if (hit < 0) {
                if (nh < H) held[nh++] = e;
                else {
                    replay = true;
                    --k;  // this entry again, by replay
                }
            } else {
                const uint8_t *q = d + list[held[hit]];
                bool q1 = (ldu32(q + 16) >> 16) & F_READ1;
                if (q1 != r1) {
                    mate_of[e] = held[hit];
                    held[hit] = held[--nh];
                }  // same end again: neither emitted nor stored
            }
        }
    }
}

// FASTQ record length of one mate: "@name/1\nSEQ\n+\nQUAL\n" (qual "*" when the BAM stores 0xff, empty when l_qseq == 0)
__device__ __forceinline__ uint32_t unmapped_fq_len(const uint8_t *p)
{
    uint32_t lq = qname_len(p), nl = 0;
    while (nl < lq && p[36 + nl]) ++nl;
    int32_t l = ldi32(p + 20);
    uint32_t nc = ldu32(p + 16) & 0xffff;
    const uint8_t *qual = p + 36 + lq + 4 * nc + (l + 1) / 2;
    uint32_t ql = l > 0 ? (qual[0] == 0xff ? 1u : (uint32_t)l) : 0u;
    return 1 + nl + 2 + 1 + (uint32_t)l + 1 + 2 + ql + 1;
}

struct UnSizesOp {
    const uint64_t *sorted;
    const uint32_t *mate_of;
    const uint8_t *d;
    uint64_t *off1, *off2;
    uint64_t cap1, cap2;
    ClipCtl *ctl;
    __device__ uint64_t n() const { return ctl->abort_side ? 0 : ctl->un_own; }
    __device__ void load(uint64_t e, uint64_t (&v)[2]) const
    {
        v[0] = v[1] = 0;
        if (mate_of[e] != 0xffffffffu) {
            const uint64_t *list = sorted + ctl->un_skip;
            const uint8_t *p = d + list[e], *q = d + list[mate_of[e]];
            bool p1 = (ldu32(p + 16) >> 16) & F_READ1;
            v[0] = unmapped_fq_len(p1 ? p : q), v[1] = unmapped_fq_len(p1 ? q : p);
        }
    }
    __device__ void store(uint64_t e, const uint64_t (&excl)[2], const uint64_t (&)[2]) const { off1[e] = excl[0], off2[e] = excl[1]; }
    __device__ void total(const uint64_t (&t)[2]) const
    {
        ctl->un1_bytes = t[0], ctl->un2_bytes = t[1];
        if (t[0] > cap1 || t[1] > cap2) ctl->abort_side = 1;
    }
};

__device__ void write_unmapped_fq(const uint8_t *p, char end, char *o, uint32_t lane)
{
    uint32_t lq = qname_len(p), nl = 0;
    while (nl < lq && p[36 + nl]) ++nl;
    int32_t l = ldi32(p + 20);
    uint32_t nc = ldu32(p + 16) & 0xffff;
    const uint8_t *seq = p + 36 + lq + 4 * nc, *qual = seq + (l + 1) / 2;
    bool noq = l > 0 && qual[0] == 0xff;
    uint32_t ql = l > 0 ? (noq ? 1u : (uint32_t)l) : 0u;
    if (lane == 0) {
        o[0] = '@', o[1 + nl] = '/', o[2 + nl] = end, o[3 + nl] = '\n';
        o[4 + nl + l] = '\n', o[5 + nl + l] = '+', o[6 + nl + l] = '\n', o[7 + nl + l + ql] = '\n';
        if (noq) o[7 + nl + l] = '*';
    }
    for (uint32_t j = lane; j < nl; j += 32) o[1 + j] = (char)p[36 + j];
    for (uint32_t j = lane; j < (uint32_t)l; j += 32) {
        uint32_t nib = (seq[j >> 1] >> ((~j & 1) << 2)) & 15;
        o[4 + nl + j] = "=ACMGRSVTWYHKDBN"[nib];  // GetSeqAndQual, clip_reads.cpp:375-388 (no toupper needed)
        if (!noq) o[7 + nl + l + j] = (char)(qual[j] + 33);
    }
}

__global__ void __launch_bounds__(128)
    unmapped_write(const ClipCtl *__restrict__ ctl, const uint64_t *__restrict__ sorted, const uint32_t *__restrict__ mate_of, const uint8_t *__restrict__ d,
                   const uint64_t *__restrict__ off1, const uint64_t *__restrict__ off2, char *__restrict__ out1,
                   char *__restrict__ out2)
{
    if (ctl->abort_side) return;
    const uint32_t n = ctl->un_own, lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t *list = sorted + ctl->un_skip;
    for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += n_warps) {
        if (mate_of[e] == 0xffffffffu) continue;
        const uint8_t *p = d + list[e], *q = d + list[mate_of[e]];
        bool p1 = (ldu32(p + 16) >> 16) & F_READ1;
        write_unmapped_fq(p1 ? p : q, '1', out1 + off1[e], lane);
        write_unmapped_fq(p1 ? q : p, '2', out2 + off2[e], lane);
    }
}

// ---- shard plumbing: raw records of the unmapped branch, packed in file order ------------------------------------
// group of a record: FNV-1a-64 of its name modulo the number of groups; key = group, value = file-order index
__global__ void export_group(const ClipCtl *__restrict__ ctl, const uint64_t *__restrict__ sorted, const uint8_t *__restrict__ d, uint32_t parts,
                             uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
    if (ctl->abort_side) return;
    const uint32_t n = ctl->un_own;
    const uint64_t *list = sorted + ctl->un_skip;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint8_t *p = d + list[i];
        uint32_t lq = qname_len(p);
        uint64_t h = 0xcbf29ce484222325ull;
        for (uint32_t j = 0; j < lq && p[36 + j]; ++j) h = (h ^ p[36 + j]) * 0x100000001b3ull;
        key[i] = parts > 1 ? h % parts : 0, val[i] = i;
    }
}
struct ExportScanOp {
    const uint64_t *sorted;
    const uint64_t *gkey;   // group of the i-th exported record
    const uint32_t *order;  // its file-order index
    const uint8_t *d;
    uint64_t *dst_off;
    uint64_t cap;
    ClipCtl *ctl;
    __device__ uint64_t n() const { return ctl->abort_side ? 0 : ctl->un_own; }
    __device__ void load(uint64_t i, uint64_t (&v)[1]) const { v[0] = 4ull + ldu32(d + sorted[ctl->un_skip + order[i]]); }
    __device__ void store(uint64_t i, const uint64_t (&excl)[1], const uint64_t (&)[1]) const
    {
        dst_off[i] = excl[0];
        if (i == 0 || gkey[i] != gkey[i - 1]) ctl->part_off[gkey[i] & 63] = excl[0] + 1;
    }
    __device__ void total(const uint64_t (&t)[1]) const
    {
        ctl->export_bytes = t[0];
        if (t[0] > cap) ctl->abort_side = 1;
    }
};
__global__ void record_copy(const ClipCtl *__restrict__ ctl, const uint64_t *__restrict__ sorted, const uint32_t *__restrict__ order,
                            const uint8_t *__restrict__ d, const uint64_t *__restrict__ dst_off, uint8_t *__restrict__ dst)
{
    if (ctl->abort_side) return;
    const uint32_t n = ctl->un_own, lane = threadIdx.x & 31, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint64_t *list = sorted + ctl->un_skip;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += n_warps) {
        const uint8_t *src = d + list[order[w]];
        uint8_t *out = dst + dst_off[w];
        const uint64_t bytes = 4ull + ldu32(src);
        for (uint64_t i = lane; i < bytes; i += 32) out[i] = src[i];
    }
}

// ---- shard plumbing: tid of the last mapped-branch record (what the next shard needs as prev_tid, quirk Q1) -----------
__global__ void last_mapped_walk(const uint8_t *__restrict__ d, uint64_t n_chunks, const uint64_t *__restrict__ guess,
                                 const uint32_t *__restrict__ count, int32_t *__restrict__ chunk_tid, unsigned long long *__restrict__ last_chunk)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t o = guess[c];
    int32_t tid = 0;
    bool any = false;
    for (uint32_t i = 0, k = count[c]; i < k; ++i) {
        const Core core = load_core(d + o);
        if (!(core.flag & (F_UNMAP | F_MUNMAP))) tid = core.tid, any = true;  // clip_reads.h:415,423-438: only these move last_tid
        o += 4 + (uint64_t)(uint32_t)core.block_size;
    }
    chunk_tid[c] = tid;
    if (any) atomicMax(last_chunk, (unsigned long long)c + 1);
}

static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

extern "C" int svb_bam_last_mapped_tid(svb_ctx *ctx, svb_bam *bam, int32_t *has_one, int32_t *tid)
{
    if (!ctx || !bam || !has_one || !tid) return svb_fail(ctx, SVB_ERR_ARG, "svb_bam_last_mapped_tid: null argument");
    CK(cudaSetDevice(ctx->device));
    CKR(ensure_counts(ctx, bam));  // verified chunk table
    cudaStream_t s = ctx->stream;
    DevBuf<int32_t> chunk_tid;
    DevBuf<unsigned long long> last_chunk;
    CK(chunk_tid.alloc(bam->n_chunks, s));
    CK(last_chunk.alloc(1, s));
    CK(cudaMemsetAsync(last_chunk.p, 0, 8, s));
    last_mapped_walk<<<nblk(bam->n_chunks, 128), 128, 0, s>>>(bam->d_data, bam->n_chunks, bam->d_guess, bam->d_count, chunk_tid.p, last_chunk.p);
    unsigned long long lc = 0;
    CK(cudaMemcpyAsync(&lc, last_chunk.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *has_one = lc != 0, *tid = 0;
    if (lc) CK(cudaMemcpy(tid, chunk_tid.p + (lc - 1), 4, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- host orchestration -----------------------------------------------------------------------------------
namespace {
enum Hint { H_CLIPPED, H_UNMAPPED, H_SWITCH, H_CAND, H_ARENA, H_CLIP, H_FQ, H_UN1, H_UN2, H_EXPORT };

struct Caps {
    uint32_t clipped, un, sw, cand;
    uint64_t arena, clip, fq, un1, un2, exp;
};

// everything the pipeline keeps in the workspace
struct ClipBuffers {
    ClipCtl *ctl;
    unsigned long long *ticket;
    ClipQueues q;
    CandArrays c;
    uint64_t *key[2];
    uint32_t *val[2];
    uint64_t *sw_sorted, *un_sorted, *ukey[2];
    uint32_t *uval[2], *mate_of;
    uint32_t *start, *maxl, *maxr;
    uint64_t *arena_off;
    char *arena_seq, *arena_qual;
    ClusterOut co;
    uint32_t *cl_seg, *cl_slot;
    uint64_t *clip_off, *fq_off, *off1, *off2;
    char *nblob;
    uint32_t *noff;
    uint32_t *bucket_cnt, *bucket_base;  // candidates per 16 KiB chunk and their prefix (file order without a sort)
    RadixScratch rs_key, rs_sw, rs_un, rs_hash;
    ScanScratch sc_chunk, sc_bucket, sc_seg, sc_stats, sc_cl, sc_text, sc_un;
    size_t zero_end;  // [0, zero_end) is cleared before the first launch
};

void carve(Bump &b, ClipBuffers &B, const Caps &cap, uint64_t n_chunks, int off_passes, int key_passes, size_t name_bytes, int32_t n_ref, bool pair_mode,
           bool export_mode)
{
    // --- cleared region: control block, tickets, look-back states
    B.ctl = b.get<ClipCtl>(1);
    B.ticket = b.get<unsigned long long>(1);
    B.rs_key = radix_scratch(b, cap.cand, key_passes);
    B.rs_sw = radix_scratch(b, cap.sw, off_passes);
    B.rs_un = radix_scratch(b, cap.un, off_passes);
    B.rs_hash = radix_scratch(b, cap.un, 4);
    B.sc_chunk = scan_scratch(b, n_chunks, 1, 8);
    B.sc_bucket = scan_scratch(b, n_chunks, 1, 8);
    B.bucket_cnt = b.get<uint32_t>(n_chunks + 1);
    B.sc_seg = scan_scratch(b, cap.cand, 1, 8);
    B.sc_stats = scan_scratch(b, cap.cand, 1, 2);
    B.sc_cl = scan_scratch(b, cap.cand, 1, 4);
    B.sc_text = scan_scratch(b, cap.cand, 2, 2);
    B.sc_un = scan_scratch(b, cap.un, 2, 2);
    B.zero_end = (b.used + 255) & ~(size_t)255;
    // --- queues and per-chunk side arrays
    B.q.clipped = b.get<uint64_t>(cap.clipped), B.q.clipped_cap = cap.clipped;
    B.q.unmapped = b.get<uint64_t>(cap.un), B.q.un_cap = cap.un;
    B.q.switches = b.get<uint64_t>(cap.sw), B.q.sw_cap = cap.sw;
    B.q.counters = B.ctl ? B.ctl->counters : nullptr;
    B.q.first_mb = b.get<uint64_t>(n_chunks);
    B.q.last_mb_tid = b.get<int32_t>(n_chunks);
    B.bucket_base = b.get<uint32_t>(n_chunks + 1);
    // --- candidates
    B.c.off = b.get<uint64_t>(cap.cand), B.c.tid = b.get<int32_t>(cap.cand), B.c.pos = b.get<int32_t>(cap.cand);
    B.c.begin = b.get<uint32_t>(cap.cand), B.c.ll = b.get<uint32_t>(cap.cand), B.c.rl = b.get<uint32_t>(cap.cand);
    B.c.side = b.get<uint8_t>(cap.cand);
    for (int i = 0; i < 2; ++i) B.key[i] = b.get<uint64_t>(cap.cand), B.val[i] = b.get<uint32_t>(cap.cand);
    B.sw_sorted = b.get<uint64_t>(cap.sw);
    B.un_sorted = b.get<uint64_t>(cap.un);
    for (int i = 0; i < 2; ++i) B.ukey[i] = b.get<uint64_t>(cap.un), B.uval[i] = b.get<uint32_t>(cap.un);
    B.mate_of = b.get<uint32_t>(pair_mode ? cap.un : 1);
    B.start = b.get<uint32_t>((size_t)cap.cand + 1), B.maxl = b.get<uint32_t>(cap.cand), B.maxr = b.get<uint32_t>(cap.cand);
    B.arena_off = b.get<uint64_t>(cap.cand);
    B.arena_seq = b.get<char>(cap.arena), B.arena_qual = b.get<char>(cap.arena);
    B.co.len_l = b.get<uint32_t>(cap.cand), B.co.len_r = b.get<uint32_t>(cap.cand), B.co.support = b.get<uint32_t>(cap.cand);
    B.co.cig_off = b.get<uint64_t>(cap.cand), B.co.noqual = b.get<uint8_t>(cap.cand), B.co.seg_ncl = b.get<uint32_t>(cap.cand);
    B.cl_seg = b.get<uint32_t>(cap.cand), B.cl_slot = b.get<uint32_t>(cap.cand);
    B.clip_off = b.get<uint64_t>(cap.cand), B.fq_off = b.get<uint64_t>(cap.cand);
    B.off1 = b.get<uint64_t>((pair_mode || export_mode) ? cap.un : 1), B.off2 = b.get<uint64_t>(pair_mode ? cap.un : 1);
    B.nblob = b.get<char>(name_bytes + 1);
    B.noff = b.get<uint32_t>((size_t)n_ref + 1);
}

int bits_of(uint64_t v)
{
    int b = 1;
    while (b < 64 && (v >> b)) ++b;
    return b;
}
}  // namespace

extern "C" int svb_getclip(svb_ctx *ctx, svb_bam *bam, const svb_getclip_params *prm, svb_clusters **out_)
{
    if (!ctx || !bam || !prm || !out_) return svb_fail(ctx, SVB_ERR_ARG, "svb_getclip: null argument");
    CK(cudaSetDevice(ctx->device));
    if (bam->names.size() != (size_t)bam->n_ref) return svb_fail(ctx, SVB_ERR_ARG, "svb_getclip: reference names not set (svb_bam_set_refs)");
    cudaStream_t s = ctx->stream, side = ctx->aux[0];
    svb_clusters *res = new svb_clusters();
    res->ctx = ctx;
    std::unique_ptr<svb_clusters> guard(res);
    const uint64_t n_chunks = bam->n_chunks, stream_bytes = bam->nbytes - bam->first;
    const bool export_mode = prm->export_unmapped_records != 0, pair_mode = !export_mode;
    res->gz_mode = prm->gz_outputs != 0 && !export_mode;
    const bool want_rows = prm->with_rows != 0, unmapped_only = prm->unmapped_only != 0;
    const int n_parts = std::max(1, std::min(64, prm->export_partitions));
    const int off_passes = (bits_of(bam->nbytes) + 7) / 8;  // offsets sort on just the bytes they use
    // key = run << 33 | side << 32 | pos: runs are numbered by chromosome switches (at most sw_cap of them)
    std::string nblob;
    std::vector<uint32_t> noff(bam->n_ref + 1, 0);
    for (int32_t t = 0; t < bam->n_ref; ++t) {
        noff[t] = (uint32_t)nblob.size();
        nblob += bam->names[t];
    }
    noff[bam->n_ref] = (uint32_t)nblob.size();

    // capacities: sized from the stream (a record is at least ~40 bytes; ~2 % of them are soft-clipped) and from what the last
    // call on this context needed; an overflow reports the exact need and the pass is run again
    Caps cap;
    {
        const uint64_t est = stream_bytes / 1600 + 4096;
        auto want32 = [&](Hint h, uint64_t base) { return (uint32_t)std::min<uint64_t>(std::max(base, ctx->hint[h] + ctx->hint[h] / 8 + 1024), 0xfffffff0u); };
        auto want64 = [&](Hint h, uint64_t base) { return std::max(base, ctx->hint[h] + ctx->hint[h] / 8 + 4096); };
        cap.clipped = want32(H_CLIPPED, est), cap.un = want32(H_UNMAPPED, est), cap.sw = want32(H_SWITCH, 1 << 16);
        cap.cand = want32(H_CAND, est);
        cap.arena = want64(H_ARENA, stream_bytes / 48 + (1 << 20));
        cap.clip = want64(H_CLIP, stream_bytes / 40 + (1 << 20)), cap.fq = want64(H_FQ, stream_bytes / 96 + (1 << 20));
        cap.un1 = pair_mode ? want64(H_UN1, stream_bytes / 64 + (1 << 20)) : 1, cap.un2 = pair_mode ? want64(H_UN2, stream_bytes / 64 + (1 << 20)) : 1;
        cap.exp = export_mode ? want64(H_EXPORT, stream_bytes / 32 + (1 << 20)) : 1;
        if (unmapped_only) {
            // a stream of nothing but unmapped-branch records (the shards' exported mates): every record is queued and nearly all of
            // its bytes come back as FASTQ text - the estimates above (and the hints of the shard's own pass) would overflow and cost
            // two or three repeats of the whole call
            cap.un = (uint32_t)std::min<uint64_t>(stream_bytes / 100 + 4096, 0xfffffff0u);
            cap.un1 = cap.un2 = stream_bytes + (1 << 20);
        }
    }
    ClipCtl h{};
    for (int attempt = 0;; ++attempt) {
        if (attempt > 6) return svb_fail(ctx, SVB_ERR_CUDA, "svb_getclip: the pass did not settle");
        const int key_passes = (33 + bits_of(cap.sw) + 7) / 8;
        // ---- buffers
        ClipBuffers B{};
        {
            Bump measure(nullptr);
            carve(measure, B, cap, n_chunks, off_passes, key_passes, nblob.size(), bam->n_ref, pair_mode, export_mode);
            CKR(ctx->ws_reserve(0, measure.used));
            Bump real(ctx->ws[0]);
            carve(real, B, cap, n_chunks, off_passes, key_passes, nblob.size(), bam->n_ref, pair_mode, export_mode);
        }
        B.q.min_mapq = prm->min_mapq;
        res->drop_device();
        CKR(res->take(0, cap.clip));
        CKR(res->take(1, cap.fq));
        CKR(res->take(2, cap.un1));
        CKR(res->take(3, cap.un2));
        if (export_mode) CKR(res->take(4, cap.exp));
        if (want_rows && !bam->rows_ready) CKR(alloc_rows(ctx, bam, rows_first(bam)));
        const bool do_rows = want_rows && !bam->rows_ready;
        CK(cudaMemsetAsync(ctx->ws[0], 0, B.zero_end, s));
        CK(cudaMemcpyAsync(B.nblob, nblob.data(), nblob.size(), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(B.noff, noff.data(), noff.size() * 4, cudaMemcpyHostToDevice, s));
        ClipCtl *ctl = B.ctl;
        ClipParams P{prm->min_mapq, prm->save_low_quality, INT32_MIN, INT32_MIN, INT32_MAX, INT32_MAX};
        if (prm->key_filter) P.lo_tid = prm->key_lo_tid, P.lo_pos = prm->key_lo_pos, P.hi_tid = prm->key_hi_tid, P.hi_pos = prm->key_hi_pos;

        // ---- 1. the walker: every record head once (getclip's scan and, when asked, getsv's rows)
        CKR(launch_walk(ctx, s, bam, true, do_rows, B.q, ctl->flags, B.ticket));
        clip_first<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, n_chunks, prm->prev_tid, B.q);
        launch_chunk_scan(ctx, s, bam, B.sc_chunk, &ctl->flags[1], ctl->c64);
        CK(cudaEventRecord(ctx->fork_event, s));

        // ---- 2. side stream: the switch list, sorted (make_keys numbers the flush runs with it) - first, so that the main stream
        //         never waits for the mates; then the unmapped-branch records - pair mates by name, emit the two FASTQ files (or
        //         export the records)
        CK(cudaStreamWaitEvent(side, ctx->fork_event, 0));
        {
            RadixJob js{{B.q.switches, B.sw_sorted}, {nullptr, nullptr}, &ctl->counters[2], cap.sw, 0, off_passes, B.rs_sw};
            radix_sort(ctx, side, js);
        }
        CK(cudaEventRecord(ctx->sw_event, side));
        {
            ProfScope ps(ctx, "unmapped_pair", 0, side);
            RadixJob ju{{B.q.unmapped, B.un_sorted}, {nullptr, nullptr}, &ctl->counters[1], cap.un, 0, off_passes, B.rs_un};
            radix_sort(ctx, side, ju);
            unmapped_own<<<1, 1, 0, side>>>(ctl->counters, cap.un, B.un_sorted, prm->halo_bytes, ctl);
            if (export_mode) {
                const unsigned g = grid_for(ctx, cap.un, 256, 4);
                export_group<<<g, 256, 0, side>>>(ctl, B.un_sorted, bam->d_data, (uint32_t)n_parts, B.ukey[0], B.uval[0]);
                RadixJob jg{{B.ukey[0], B.ukey[1]}, {B.uval[0], B.uval[1]}, &ctl->un_own, cap.un, 0, 1, B.rs_hash};
                radix_sort(ctx, side, jg);  // stable: file order inside a group
                ExportScanOp op{B.un_sorted, B.ukey[1], B.uval[1], bam->d_data, B.off1, cap.exp, ctl};
                launch_scan<1, 2>(ctx, side, op, B.sc_un, cap.un);
                record_copy<<<grid_for(ctx, (uint64_t)cap.un * 32, 128, 8), 128, 0, side>>>(ctl, B.un_sorted, B.uval[1], bam->d_data, B.off1, res->d_export);
            } else {
                const unsigned g = grid_for(ctx, cap.un, 256, 4);
                unmapped_hash<<<g, 256, 0, side>>>(ctl, B.un_sorted, bam->d_data, B.ukey[0], B.uval[0]);
                RadixJob jh{{B.ukey[0], B.ukey[1]}, {B.uval[0], B.uval[1]}, &ctl->un_own, cap.un, 0, 4, B.rs_hash};
                radix_sort(ctx, side, jh);
                CK(cudaMemsetAsync(B.mate_of, 0xff, (size_t)cap.un * 4, side));
                unmapped_pair<<<g, 256, 0, side>>>(ctl, B.ukey[1], B.uval[1], B.un_sorted, bam->d_data, B.mate_of);
                UnSizesOp op{B.un_sorted, B.mate_of, bam->d_data, B.off1, B.off2, cap.un1, cap.un2, ctl};
                launch_scan<2, 2>(ctx, side, op, B.sc_un, cap.un);
            }
        }
        if (pair_mode) {
            ProfScope ps(ctx, "unmapped_write", 0, side);
            unmapped_write<<<grid_for(ctx, (uint64_t)cap.un * 32, 128, 8), 128, 0, side>>>(ctl, B.un_sorted, B.mate_of, bam->d_data, B.off1, B.off2,
                                                                                         res->d_text[2], res->d_text[3]);
        }
        CK(cudaEventRecord(ctx->join_event, side));

        // ---- 3. main stream: evaluate the queued soft-clipped records, order candidates: BAM order first (stable base), then (run, side, pos)
        if (!unmapped_only) {
        {
            ProfScope ps(ctx, "clip_eval", 0);
            clip_eval<<<grid_for(ctx, cap.clipped, 128, 8), 128, 0, s>>>(bam->d_data, B.q, P, B.c, cap.cand, ctl, bam->chunk_log2, B.bucket_cnt,
                                                                        (uint32_t *)B.key[1]);
        }
        {
            // a both-side-clipped read yields two candidates with the same record offset but different sides, so
            // (key, record, side) is unique and the order is deterministic
            ProfScope ps(ctx, "sort_candidates", 0);
            const unsigned g = grid_for(ctx, cap.cand, 256, 4);
            BucketScanOp opb{n_chunks, B.bucket_cnt, B.bucket_base};
            launch_scan<1, 8>(ctx, s, opb, B.sc_bucket, n_chunks);
            cand_group<<<g, 256, 0, s>>>(&ctl->counters[3], cap.cand, bam->chunk_log2, B.c.off, B.bucket_base, (const uint32_t *)B.key[1], B.val[1], ctl);
            CK(cudaStreamWaitEvent(s, ctx->sw_event, 0));  // (the sorted switch list comes from the side stream)
            make_keys<<<g, 256, 0, s>>>(ctl->counters, cap.cand, cap.sw, bam->chunk_log2, B.val[1], B.bucket_base, B.bucket_cnt, B.c, B.sw_sorted, B.key[0],
                                        B.val[0]);
            RadixJob j2{{B.key[0], B.key[1]}, {B.val[0], B.val[1]}, &ctl->counters[3], cap.cand, 0, key_passes, B.rs_key};
            radix_sort(ctx, s, j2);
        }
        const uint32_t *order = B.val[1];
        // ---- 4. segments, 5. greedy clustering, 6. text
        {
            ProfScope ps(ctx, "segments", 0);
            SegScanOp op1{ctl->counters, cap.cand, B.key[1], B.start, ctl};
            launch_scan<1, 8>(ctx, s, op1, B.sc_seg, cap.cand);
            SegStatsOp op2{B.start, order, B.c, B.maxl, B.maxr, B.arena_off, cap.arena, ctl};
            launch_scan<1, 2>(ctx, s, op2, B.sc_stats, cap.cand);
        }
        {
            ProfScope ps(ctx, "cluster_build", 0);
            // (occupancy is not what bounds this kernel: 32- and 40-register builds with 12 / 16 CTAs per SM ran in the same 0.175 ms)
            cluster_singletons<<<grid_for(ctx, cap.cand, 128, 8), 128, 0, s>>>(bam->d_data, ctl, B.start, order, B.c, B.co);
            cluster_build<<<grid_for(ctx, (uint64_t)cap.cand * 32, 128, 16), 128, 0, s>>>(bam->d_data, ctl, B.start, order, B.c, B.maxl, B.maxr, B.arena_off,
                                                                                        B.arena_seq, B.arena_qual, prm->match_rate, B.co);
        }
        NameTable nt{B.nblob, B.noff};
        {
            ProfScope ps(ctx, "text_write", 0);
            ClusterScanOp op3{B.co.seg_ncl, B.cl_seg, B.cl_slot, ctl};
            launch_scan<1, 4>(ctx, s, op3, B.sc_cl, cap.cand);
            TextScanOp op4{B.cl_seg, B.cl_slot, B.start, order, B.c, B.co, bam->d_data, nt, B.clip_off, B.fq_off, cap.clip, cap.fq, ctl};
            launch_scan<2, 2>(ctx, s, op4, B.sc_text, cap.cand);
            // the heads on the side stream (long idle by now), the strings here: disjoint bytes of the same two texts
            CK(cudaEventRecord(ctx->fork_event, s));
            CK(cudaStreamWaitEvent(side, ctx->fork_event, 0));
            text_heads<<<grid_for(ctx, cap.cand, 128, 8), 128, 0, side>>>(ctl, B.cl_seg, B.cl_slot, B.start, order, B.c, B.co, bam->d_data, nt, B.clip_off,
                                                                         B.fq_off, res->d_text[0], res->d_text[1]);
            CK(cudaEventRecord(ctx->join_event, side));
            text_write<<<grid_for(ctx, (uint64_t)cap.cand * 32, 128, 16), 128, 0, s>>>(ctl, B.cl_seg, B.cl_slot, B.start, order, B.c, B.co, bam->d_data, B.maxl, B.maxr,
                                                                                     B.arena_off, B.arena_seq, B.arena_qual, B.clip_off, B.fq_off,
                                                                                     res->d_text[0], res->d_text[1]);
        }
        }
        // ---- the one read-back (after the side stream's work)
        CK(cudaStreamWaitEvent(s, ctx->join_event, 0));
        CK(cudaMemcpyAsync(ctx->ctl_host, ctl, sizeof(ClipCtl), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(cudaGetLastError());
        memcpy(&h, ctx->ctl_host, sizeof h);
        if (h.flags[1]) {  // a guess was wrong: repair them and walk again
            if (attempt >= 3) return svb_fail(ctx, SVB_ERR_FORMAT, "record chain does not verify");
            CKR(repair_guesses(ctx, bam));
            continue;
        }
        if (h.flags[2])
            return svb_fail(ctx, SVB_ERR_FORMAT,
                            "a soft-clipped read with `=` CIGAR operations reaches a breakpoint key of the next coordinate-range shard from "
                            "outside that shard's halo (the .bai counts M, D, N only): shard by chromosome or run the single-process command");
        bool again = false;
        auto grow32 = [&](uint32_t &c, uint32_t need) {
            if (need > c) c = need + need / 16 + 64, again = true;
        };
        auto grow64 = [&](uint64_t &c, uint64_t need) {
            if (need > c) c = need + need / 16 + 4096, again = true;
        };
        grow32(cap.clipped, h.counters[0]), grow32(cap.un, h.counters[1]), grow32(cap.sw, h.counters[2]), grow32(cap.cand, h.counters[3]);
        // (byte counts are exact when the stage that computes them ran, 0 when an earlier overflow cut the pipeline short)
        grow64(cap.arena, h.arena_bytes), grow64(cap.clip, h.clip_bytes), grow64(cap.fq, h.fq_bytes);
        grow64(cap.un1, h.un1_bytes), grow64(cap.un2, h.un2_bytes), grow64(cap.exp, h.export_bytes);
        if (do_rows) {
            if (h.flags[0]) free_rows(bam);  // more records in a chunk than row slots: the getsv passes make their own rows
            else if (!again) bam->rows_ready = true, bam->rows_indexed = false;
        }
        if (again) continue;
        if (h.abort_main || h.abort_side) return svb_fail(ctx, SVB_ERR_CUDA, "svb_getclip: inconsistent overflow state");
        break;
    }
    if (!bam->counted) CKR(accept_counts(ctx, bam, h.c64[0], h.c64[1]));
    if (!unmapped_only) {  // (the mates' stream is another kind of input: its needs say nothing about the next shard)
        ctx->hint[H_CLIPPED] = h.counters[0], ctx->hint[H_UNMAPPED] = h.counters[1], ctx->hint[H_SWITCH] = h.counters[2], ctx->hint[H_CAND] = h.counters[3];
        ctx->hint[H_ARENA] = h.arena_bytes, ctx->hint[H_CLIP] = h.clip_bytes, ctx->hint[H_FQ] = h.fq_bytes;
        ctx->hint[H_UN1] = h.un1_bytes, ctx->hint[H_UN2] = h.un2_bytes, ctx->hint[H_EXPORT] = h.export_bytes;
    }
    res->n_candidates = h.counters[3], res->n_clusters = h.n_cl;
    res->text_len[0] = h.clip_bytes, res->text_len[1] = h.fq_bytes, res->text_len[2] = h.un1_bytes, res->text_len[3] = h.un2_bytes;
    res->export_len = h.export_bytes, res->n_parts = n_parts;
    {   // group offsets: a group without records starts where the next one does
        uint64_t next = h.export_bytes;
        res->part_off[n_parts] = next;
        for (int p = n_parts - 1; p >= 0; --p) {
            if (h.part_off[p]) next = h.part_off[p] - 1;
            res->part_off[p] = next;
        }
    }
    if (res->gz_mode) {  // the four files leave the device as gzip images (gzip.cu); an empty file is one empty member
        for (int w = 0; w < 4; ++w) CKR(gzip_on_device(ctx, res->d_text[w], res->text_len[w], &res->gz[w]));
        res->drop_device();
        for (int w = 0; w < 4; ++w) res->text_len[w] = 0;
    }
    *out_ = guard.release();
    return 0;
}

extern "C" void svb_clusters_free(svb_clusters *c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    delete c;
}
extern "C" uint64_t svb_clusters_count(const svb_clusters *c) { return c ? c->n_clusters : 0; }
extern "C" uint64_t svb_clusters_candidates(const svb_clusters *c) { return c ? c->n_candidates : 0; }
// gzip.cu alone (tests / other text outputs): host text -> device -> gzip image in a malloc'ed host buffer (svb_free)
extern "C" int svb_gzip_text(svb_ctx *ctx, const void *text, uint64_t n, char **gz, uint64_t *gz_len)
{
    if (!ctx || (!text && n) || !gz || !gz_len) return svb_fail(ctx, SVB_ERR_ARG, "svb_gzip_text: null argument");
    CK(cudaSetDevice(ctx->device));
    DevBuf<char> d;
    CK(d.alloc(n, ctx->stream));
    if (n) CK(cudaMemcpyAsync(d.p, text, n, cudaMemcpyHostToDevice, ctx->stream));
    PinnedBuf out;
    int rc = gzip_on_device(ctx, d.p, n, &out);
    if (rc == 0) {
        *gz = (char *)malloc(out.n ? out.n : 1);
        if (!*gz) rc = svb_fail(ctx, SVB_ERR_IO, "out of memory");
        else memcpy(*gz, out.p, out.n), *gz_len = out.n;
    }
    out.release(ctx);
    return rc;
}

extern "C" int svb_clusters_gz(const svb_clusters *c, int which, const char **data, uint64_t *len)
{
    if (!c || which < 0 || which > 3 || !data || !len) return SVB_ERR_ARG;
    *data = c->gz[which].p ? c->gz[which].p : "";
    *len = c->gz[which].n;
    return 0;
}

// Results stay in HBM until they are asked for: the first request of a text copies it into pinned host memory.
extern "C" int svb_clusters_unmapped_records(const svb_clusters *c, const char **data, uint64_t *len)
{
    if (!c || !data || !len) return SVB_ERR_ARG;
    svb_ctx *ctx = c->ctx;
    if (!c->export_here && c->d_export && c->export_len) {
        CK(cudaSetDevice(ctx->device));
        CKR(c->unmapped_records.reserve(ctx, c->export_len));
        CK(cudaMemcpyAsync(c->unmapped_records.p, c->d_export, c->export_len, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        c->export_here = true;
    }
    *data = c->unmapped_records.p ? c->unmapped_records.p : "";
    *len = c->unmapped_records.n;
    return 0;
}

extern "C" int svb_clusters_text(const svb_clusters *c, int which, const char **data, uint64_t *len)
{
    if (!c || which < 0 || which > 3 || !data || !len) return SVB_ERR_ARG;
    svb_ctx *ctx = c->ctx;
    if (!c->text_here[which] && c->d_text[which] && c->text_len[which]) {
        CK(cudaSetDevice(ctx->device));
        CKR(c->text[which].reserve(ctx, c->text_len[which]));
        CK(cudaMemcpyAsync(c->text[which].p, c->d_text[which], c->text_len[which], cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        c->text_here[which] = true;
    }
    *data = c->text[which].p ? c->text[which].p : "";
    *len = c->text[which].n;
    return 0;
}
// sizes of the four texts without copying them (the results stay in HBM)
extern "C" int svb_clusters_text_len(const svb_clusters *c, int which, uint64_t *len)
{
    if (!c || which < 0 || which > 3 || !len) return SVB_ERR_ARG;
    *len = c->gz_mode ? 0 : c->text_len[which];
    return 0;
}

extern "C" int svb_clusters_export_device(const svb_clusters *c, const void **d_records, uint64_t *len)
{
    if (!c || !d_records || !len) return SVB_ERR_ARG;
    *d_records = c->d_export, *len = c->export_len;
    return 0;
}
extern "C" int svb_clusters_export_parts(const svb_clusters *c, uint64_t *offsets, int32_t n_parts)
{
    if (!c || !offsets || n_parts != c->n_parts) return SVB_ERR_ARG;
    for (int p = 0; p <= n_parts; ++p) offsets[p] = c->part_off[p];
    return 0;
}
