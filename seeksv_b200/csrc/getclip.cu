// getclip on the device: soft-clip candidate scan, breakpoint-key sort, per-key greedy clustering and
// text emission. Replaces the per-record loop of InputBamOutputReads (clip_reads.h:410-440) with its
// callees GetSClipReads / GetSeq / GenerateCigar / InsertSeq / ReadsInfo::ChangeSeqAndQual
// (clip_reads.cpp:57-108,112-192,260-329) and the writer DisplaySClipReadsAndClipFq (clip_reads.h:300-345).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstring>
#include <memory>

#include "common.cuh"
#include "stream.cuh"

struct svb_clusters {
    svb_ctx *ctx = nullptr;
    PinnedBuf text[4];  // pinned host memory: D2H at full PCIe rate, recycled through the ctx pool
    PinnedBuf gz[4];             // the same four files as gzip images, compressed on the device (svb_getclip_params.gz_outputs)
    bool gz_mode = false;
    PinnedBuf unmapped_records;  // sharded runs: the raw unmapped-branch records (svb_getclip_params.export_unmapped_records)
    uint64_t n_clusters = 0, n_candidates = 0;
    ~svb_clusters()
    {
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

struct ScanOut {
    CandArrays c;
    uint32_t cand_cap;
    uint64_t *unmapped;
    uint32_t un_cap;
    uint64_t *switches;
    uint32_t sw_cap;
    uint32_t *counters;  // [0] candidates, [1] unmapped-branch records, [2] chromosome switches, [3] soft-clipped records
    uint64_t *clipped;   // records whose first or last CIGAR op is S and that pass the cheap filters: evaluated by clip_eval
    uint32_t clipped_cap;
    // range shards: only breakpoint keys (tid, pos) in [lo, hi) belong to this shard (whole file: everything)
    int32_t lo_tid, lo_pos, hi_tid, hi_pos;
    __device__ __forceinline__ bool owns(int32_t tid, int32_t pos) const
    {
        const bool ge_lo = tid > lo_tid || (tid == lo_tid && pos >= lo_pos);
        const bool lt_hi = tid < hi_tid || (tid == hi_tid && pos < hi_pos);
        return ge_lo && lt_hi;
    }
};

// The cheap part of GetSClipReads (clip_reads.cpp:116-118,122): first / last CIGAR op and the H / mapQ / DUP filters.
// Only records that pass (about 2 %) are queued for the expensive part (reference length, XC aux scan, emission), so the
// walker's warps do not stall 31 lanes while one lane scans an aux block.
__device__ __forceinline__ void queue_if_clipped(const uint8_t *__restrict__ d, uint64_t o, const Core &k, int32_t min_mapq,
                                                 const ScanOut &out)
{
    if (k.n_cigar == 0 || (int32_t)k.mapq < min_mapq || (k.flag & F_DUP)) return;
    const uint8_t *cig = d + o + 36 + k.l_qname;
    uint32_t op1 = ldu32(cig) & 15, op2 = ldu32(cig + 4 * (k.n_cigar - 1)) & 15;
    if (op1 == OP_H || op2 == OP_H || (op1 != OP_S && op2 != OP_S)) return;
    uint32_t s = atomicAdd(&out.counters[3], 1u);
    if (s < out.clipped_cap) out.clipped[s] = o;
}

// GetSClipReads (clip_reads.cpp:112-192) for one mapped-branch record that survived the chromosome-switch test.
// Sequence and aux bytes are only touched for soft-clipped reads (~2 % of the records).
__device__ void eval_clip(const uint8_t *__restrict__ d, uint64_t o, const Core &k, int32_t min_mapq, int32_t save_low_quality,
                          const ScanOut &out)
{
    if (k.n_cigar == 0) return;
    const uint8_t *p = d + o;
    const uint8_t *cig = p + 36 + k.l_qname;
    uint32_t first = ldu32(cig), last = ldu32(cig + 4 * (k.n_cigar - 1));
    uint32_t op1 = first & 15, op2 = last & 15;
    if (op1 == OP_H || op2 == OP_H || (int32_t)k.mapq < min_mapq || (k.flag & F_DUP)) return;  // clip_reads.cpp:118
    bool s1 = op1 == OP_S, s2 = op2 == OP_S;
    if (!s1 && !s2) return;
    // GenerateCigar's l (clip_reads.cpp:322): M, D, =, N - X is not counted (quirk Q5)
    int32_t reflen = 0;
    for (uint32_t j = 0; j < k.n_cigar; ++j) {
        uint32_t w = ldu32(cig + 4 * j), op = w & 15;
        if (op == OP_M || op == OP_D || op == OP_EQ || op == OP_N) reflen += (int32_t)(w >> 4);
    }
    const uint8_t *aux = cig + 4 * k.n_cigar + (k.l_qseq + 1) / 2 + k.l_qseq;
    int32_t xc = aux_xc(aux, (uint32_t)max((int64_t)0, (int64_t)(p + 4 + k.block_size - aux)));
    uint32_t len1 = first >> 4, len2 = last >> 4;
    bool emit5 = false, emit3 = false;
    uint32_t b5 = 0, l5 = 0, r5 = 0, b3 = 0, l3 = 0, r3 = 0;
    if (s1 != s2) {
        if (xc != 0 && !save_low_quality) return;
        if (s1) {
            if ((int64_t)len1 > k.l_qseq) return;
            emit5 = true, l5 = len1, r5 = k.l_qseq - len1;
        } else {
            if ((int64_t)len2 > k.l_qseq) return;
            emit3 = true, l3 = k.l_qseq - len2, r3 = len2;
        }
    } else {
        int64_t mid = (int64_t)k.l_qseq - len1 - len2;
        if (mid < 0 || k.n_cigar < 2) return;  // (undefined in the reference: a CIGAR that is one S op)
        if (xc != 0 && !save_low_quality) {
            if (!(k.flag & F_REVERSE)) emit5 = true;
            else emit3 = true;
        } else
            emit5 = emit3 = true;
        l5 = len1, r5 = (uint32_t)mid;             // clip_reads.cpp:152,179
        b3 = len1, l3 = (uint32_t)mid, r3 = len2;  // clip_reads.cpp:154,185
    }
    const CandArrays &c = out.c;
    emit5 = emit5 && out.owns(k.tid, k.pos + 1);
    emit3 = emit3 && out.owns(k.tid, k.pos + reflen);
    if (emit5) {
        uint32_t s = atomicAdd(&out.counters[0], 1u);
        if (s < out.cand_cap) {
            c.off[s] = o, c.tid[s] = k.tid, c.pos[s] = k.pos + 1, c.begin[s] = b5, c.ll[s] = l5, c.rl[s] = r5;
            c.side[s] = 0;
        }
    }
    if (emit3) {
        uint32_t s = atomicAdd(&out.counters[0], 1u);
        if (s < out.cand_cap) {
            c.off[s] = o, c.tid[s] = k.tid, c.pos[s] = k.pos + reflen, c.begin[s] = b3, c.ll[s] = l3, c.rl[s] = r3;
            c.side[s] = 1;
        }
    }
}

#define NO_TID INT32_MIN

// The getclip walker: one thread per 16 KiB chunk follows the record chain from the chunk's guessed first record and does
// the whole per-record work of InputBamOutputReads' loop (clip_reads.h:410-440) on the way - unmapped branch (quirk
// Q2), chromosome-switch drop (quirk Q1, against the previous mapped-branch record it has just walked over), soft-clip
// predicate. Each record head is fetched from HBM exactly once. The first mapped-branch record of a chunk needs the
// last one of an earlier chunk and is left to clip_first.
__global__ void __launch_bounds__(128)
    clip_walk(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t CHUNK_LOG2, const uint64_t *__restrict__ guess,
              uint32_t *__restrict__ count, uint64_t *__restrict__ exit_, uint64_t *__restrict__ first_mb,
              int32_t *__restrict__ last_mb_tid, int32_t min_mapq, int32_t save_low_quality, ScanOut out)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t o = guess[c], end = min(n, (c + 1) << CHUNK_LOG2), first = BAD_OFFSET;
    uint32_t cnt = 0;
    int32_t prev_tid = NO_TID;
    bool live = o < end && o + 36 <= n;  // (a partial tail shorter than a fixed part ends the walk)
    Core k;
    if (live) k = load_core(d + o);
    while (live) {
        if (k.block_size < 32) {
            o = BAD_OFFSET;
            break;
        }
        if (o + 4 + (uint64_t)k.block_size > n) break;
        // software pipeline: request the next record's fixed part before this record's CIGAR / aux bytes are waited for
        uint64_t on = o + 4 + (uint64_t)k.block_size;
        bool next_live = on < end && on + 36 <= n;
        Core kn;
        if (next_live) kn = load_core(d + on);
        ++cnt;
        if (k.flag & (F_UNMAP | F_MUNMAP)) {  // clip_reads.h:415 - the unmapped branch wins (quirk Q2)
            uint32_t s = atomicAdd(&out.counters[1], 1u);
            if (s < out.un_cap) out.unmapped[s] = o;
        } else {
            if (prev_tid == NO_TID) first = o;
            else if (k.tid != prev_tid) {  // flush + drop (clip_reads.h:423-438)
                uint32_t s = atomicAdd(&out.counters[2], 1u);
                if (s < out.sw_cap) out.switches[s] = o;
            } else
                queue_if_clipped(d, o, k, min_mapq, out);
            prev_tid = k.tid;
        }
        o = on, k = kn, live = next_live;
    }
    count[c] = cnt, exit_[c] = o, first_mb[c] = first, last_mb_tid[c] = prev_tid;
}

__global__ void __launch_bounds__(128)
    clip_first(const uint8_t *__restrict__ d, uint64_t n_chunks, const uint64_t *__restrict__ first_mb,
               const int32_t *__restrict__ last_mb_tid, int32_t prev_tid0, int32_t min_mapq, int32_t save_low_quality, ScanOut out)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks || first_mb[c] == BAD_OFFSET) return;
    int32_t prev = prev_tid0;  // clip_reads.h:407: last_tid starts at 0 (or the previous shard's last tid)
    for (uint64_t j = c; j > 0;) {
        --j;
        if (last_mb_tid[j] != NO_TID) {
            prev = last_mb_tid[j];
            break;
        }
    }
    uint64_t o = first_mb[c];
    Core k = load_core(d + o);
    if (k.tid != prev) {
        uint32_t s = atomicAdd(&out.counters[2], 1u);
        if (s < out.sw_cap) out.switches[s] = o;
    } else
        queue_if_clipped(d, o, k, min_mapq, out);
}

// The getclip full pass, streaming form (stream.cuh): TMA-staged 16 KiB tiles, one record per thread from shared memory.
// Same outputs as clip_walk (per-tile count / exit / first mapped-branch record / last mapped-branch tid, the unmapped
// list, the chromosome switches and the queue of soft-clipped records), so clip_first / clip_eval / verification follow
// unchanged.
__global__ void __launch_bounds__(STREAM_THREADS)
    clip_stream(const uint8_t *__restrict__ d, uint64_t n, uint64_t padded, uint64_t first, int32_t n_ref, uint64_t n_tiles,
                uint64_t *__restrict__ guess, uint32_t *__restrict__ count, uint64_t *__restrict__ exit_,
                uint64_t *__restrict__ first_mb, int32_t *__restrict__ last_mb_tid, int32_t min_mapq, ScanOut out)
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
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int s = it % STAGES;
        mbar_wait(&S.full[s], (it / STAGES) & 1);
        const uint64_t tile_abs = t << TILE_LOG2;
        TileWin w{S.stage[s], d + tile_abs, (uint32_t)min((uint64_t)(TILE + HALO), padded - tile_abs)};
        index_tile(S, w, t, n, first, n_ref, d);
        const uint32_t n_rec = S.n_rec;
        for (uint32_t kb = 0; kb < n_rec; kb += STREAM_THREADS) {
            const uint32_t k = kb + tid;
            unsigned long long kmin = ~0ull, kmax = 0ull;  // (k << 32 | tid) of this lane's record if it is mapped-branch
            bool have = false;
            if (k < n_rec) {
                const uint32_t off = S.rec_off[k];
                const Core c = w.core(off);
                const uint64_t o = tile_abs + off;
                if (c.flag & (F_UNMAP | F_MUNMAP)) {  // clip_reads.h:415 - the unmapped branch wins (quirk Q2)
                    uint32_t slot = atomicAdd(&out.counters[1], 1u);
                    if (slot < out.un_cap) out.unmapped[slot] = o;
                } else {
                    have = true;
                    kmin = kmax = (unsigned long long)k << 32 | (uint32_t)c.tid;
                    // quirk Q1: tid of the previous mapped-branch record (normally the record just before this one)
                    int32_t prev_tid = NO_TID;
                    for (uint32_t j = k; j > 0;) {
                        --j;
                        uint32_t oj = S.rec_off[j];
                        if (!((w.u32(oj + 16) >> 16) & (F_UNMAP | F_MUNMAP))) {
                            prev_tid = (int32_t)w.u32(oj + 4);
                            break;
                        }
                    }
                    if (prev_tid == NO_TID) {
                        // first mapped-branch record of the tile: clip_first looks into earlier tiles
                    } else if (c.tid != prev_tid) {  // flush + drop (clip_reads.h:423-438)
                        uint32_t slot = atomicAdd(&out.counters[2], 1u);
                        if (slot < out.sw_cap) out.switches[slot] = o;
                    } else if (c.n_cigar != 0 && (int32_t)c.mapq >= min_mapq && !(c.flag & F_DUP)) {
                        // cheap part of GetSClipReads (clip_reads.cpp:116-118,122) from the staged bytes
                        uint32_t cg = off + 36 + c.l_qname;
                        uint32_t op1 = w.u32(cg) & 15, op2 = w.u32(cg + 4 * (c.n_cigar - 1)) & 15;
                        if (op1 != OP_H && op2 != OP_H && (op1 == OP_S || op2 == OP_S)) {
                            uint32_t slot = atomicAdd(&out.counters[3], 1u);
                            if (slot < out.clipped_cap) out.clipped[slot] = o;
                        }
                    }
                }
            }
            // first / last mapped-branch record of the tile: warp reduction, one shared-memory atomic per warp
            if (__any_sync(0xffffffffu, have)) {
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) {
                    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, sft));
                    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, sft));
                }
                if ((tid & 31) == 0) {
                    atomicMin(&S.first_mb, kmin);
                    atomicMax(&S.last_mb, kmax);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            count[t] = n_rec, exit_[t] = S.exit_, guess[t] = S.entry;
            bool have = S.first_mb != ~0ull;
            first_mb[t] = have ? tile_abs + S.rec_off[(uint32_t)(S.first_mb >> 32)] : BAD_OFFSET;
            last_mb_tid[t] = have ? (int32_t)(uint32_t)S.last_mb : NO_TID;
            uint64_t tn = t + (uint64_t)STAGES * gridDim.x;
            if (tn < n_tiles) {
                fence_proxy_async();  // the stage was read through the generic proxy; the bulk copy writes it through the async proxy
                issue_tile(S, s, d, padded, tn);
            }
        }
        __syncthreads();
    }
}

// the expensive part of GetSClipReads for the queued soft-clipped records (one thread each)
__global__ void __launch_bounds__(128)
    clip_eval(const uint8_t *__restrict__ d, int32_t min_mapq, int32_t save_low_quality, ScanOut out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = min(out.counters[3], out.clipped_cap);
    if (i >= n) return;
    uint64_t o = out.clipped[i];
    Core k = load_core(d + o);
    eval_clip(d, o, k, min_mapq, save_low_quality, out);
}

// sort key = flush run (number of chromosome switches before the record) | side | position
__global__ void make_keys(uint32_t n, const uint32_t *__restrict__ order, CandArrays c, const uint64_t *__restrict__ sw,
                          uint32_t n_sw, uint64_t *__restrict__ key)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = order[i];
    uint64_t r = c.off[s];
    uint32_t lo = 0, hi = n_sw;  // lower_bound(sw, r): switches with index < r
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        if (sw[m] < r) lo = m + 1;
        else hi = m;
    }
    key[i] = ((uint64_t)lo << 33) | ((uint64_t)c.side[s] << 32) | (uint32_t)(c.pos[s] ^ 0x80000000);
}

__global__ void iota_u32(uint32_t n, uint32_t *v)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

__global__ void seg_flags(uint32_t n, const uint64_t *__restrict__ key, uint32_t *__restrict__ flag)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

__global__ void seg_starts(uint32_t n, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ segid,
                           uint32_t *__restrict__ start, uint32_t n_seg)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) start[segid[i] - 1] = i;
    if (i == 0) start[n_seg] = n;
}

// per segment: arena bytes = members * (longest left + longest right)
__global__ void seg_stats(uint32_t n_seg, const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, CandArrays c,
                          uint32_t *__restrict__ maxl, uint32_t *__restrict__ maxr, uint64_t *__restrict__ bytes)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    uint32_t ml = 0, mr = 0;
    for (uint32_t k = start[s]; k < start[s + 1]; ++k) {
        uint32_t x = order[k];
        ml = max(ml, c.ll[x]);
        mr = max(mr, c.rl[x]);
    }
    maxl[s] = ml, maxr[s] = mr;
    bytes[s] = (uint64_t)(start[s + 1] - start[s]) * (ml + mr);
    if (s == 0) bytes[n_seg] = 0;
}

struct ClusterOut {
    uint32_t *len_l, *len_r, *support;  // indexed by sorted candidate position (seg start + k)
    uint64_t *cig_off;                  // record whose CIGAR the cluster carries
    uint8_t *noqual;
    uint32_t *seg_ncl;
};

// One warp per breakpoint key: the reference's sequential greedy InsertSeq (clip_reads.cpp:260-283) in
// BAM order, with lane-parallel string compares and consensus updates. Strings live in a per-segment
// arena: slot k holds [left part right-aligned at column maxl | right part left-aligned at maxl].
__global__ void __launch_bounds__(128)
    cluster_build(const uint8_t *__restrict__ d, uint32_t n_seg,
                  const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, CandArrays c,
                  const uint32_t *__restrict__ maxl_, const uint32_t *__restrict__ maxr_, const uint64_t *__restrict__ arena_off,
                  char *__restrict__ arena_seq, char *__restrict__ arena_qual, double limit, ClusterOut out)
{
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (s >= n_seg) return;
    const uint32_t a = start[s], b = start[s + 1];
    const uint32_t maxl = maxl_[s], stride = maxl + maxr_[s];
    char *S = arena_seq + arena_off[s], *Q = arena_qual + arena_off[s];
    uint32_t ncl = 0;
    for (uint32_t k = a; k < b; ++k) {
        const uint32_t x = order[k];
        const uint64_t rec = c.off[x];
        const uint32_t begin = c.begin[x], ll = c.ll[x], rl = c.rl[x], side = c.side[x];
        const uint8_t *p = d + rec;
        uint32_t w = ldu32(p + 12), w2 = ldu32(p + 16);
        int32_t l_qseq = ldi32(p + 20);
        const uint8_t *seq = p + 36 + (w & 0xff) + 4 * (w2 & 0xffff);
        const uint8_t *qual = seq + (l_qseq + 1) / 2;
        const bool noq = l_qseq > 0 && qual[0] == 0xff;
        char *tS = S + (uint64_t)ncl * stride, *tQ = Q + (uint64_t)ncl * stride;  // tentative new cluster
        // GetSeq (clip_reads.cpp:286-306): 4-bit codes -> "=ACMGRSVTWYHKDBN", quality + 33
        for (uint32_t j = lane; j < ll + rl; j += 32) {
            uint32_t idx = begin + j;
            uint32_t nib = (seq[idx >> 1] >> ((~idx & 1) << 2)) & 15;
            uint32_t col = maxl - ll + j;
            tS[col] = "=ACMGRSVTWYHKDBN"[nib];
            tQ[col] = noq ? '*' : (char)(qual[idx] + 33);
        }
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

// cluster slot -> flat cluster list; one thread per segment
__global__ void list_clusters(uint32_t n_seg, const uint32_t *__restrict__ start, const uint32_t *__restrict__ seg_ncl,
                              const uint32_t *__restrict__ cl_base, uint32_t *__restrict__ cl_seg, uint32_t *__restrict__ cl_slot)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_seg) return;
    uint32_t base = cl_base[s] - seg_ncl[s];  // cl_base is the inclusive scan
    for (uint32_t k = 0; k < seg_ncl[s]; ++k) cl_seg[base + k] = s, cl_slot[base + k] = k;
}

__global__ void text_sizes(uint32_t n_cl, const uint32_t *__restrict__ cl_seg, const uint32_t *__restrict__ cl_slot,
                           const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, CandArrays c, ClusterOut out,
                           const uint8_t *__restrict__ d, NameTable names, uint64_t *__restrict__ clip_len,
                           uint64_t *__restrict__ fq_len)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cl) {
        if (i == n_cl) clip_len[i] = 0, fq_len[i] = 0;
        return;
    }
    uint32_t s = cl_seg[i], a = start[s], k = a + cl_slot[i];
    uint32_t x = order[a];
    uint32_t L = out.len_l[k], R = out.len_r[k];
    uint32_t qL = out.noqual[k] ? 1 : L, qR = out.noqual[k] ? 1 : R;
    uint32_t name = names.off[c.tid[x] + 1] - names.off[c.tid[x]];
    uint32_t cg = cigar_text(d + out.cig_off[k], nullptr);
    // chr \t pos \t side \t cigar \t aligned \t alignedQ \t clipped \t clippedQ \t support \n
    clip_len[i] = name + 1 + dec_len_i(c.pos[x]) + 1 + 2 + cg + 1 + L + 1 + qL + 1 + R + 1 + qR + 1 + dec_len(out.support[k]) + 1;
    uint32_t cl = c.side[x] == 0 ? L : R, cq = c.side[x] == 0 ? qL : qR;
    fq_len[i] = 1 + cl + 1 + cl + 1 + 2 + cq + 1;  // @seq \n seq \n + \n qual \n
}

// one warp per cluster writes its clip.gz line and its FASTQ record
__global__ void __launch_bounds__(128)
    text_write(uint32_t n_cl, const uint32_t *__restrict__ cl_seg, const uint32_t *__restrict__ cl_slot,
               const uint32_t *__restrict__ start, const uint32_t *__restrict__ order, CandArrays c, ClusterOut out,
               const uint8_t *__restrict__ d, NameTable names, const uint32_t *__restrict__ maxl_, const uint32_t *__restrict__ maxr_, const uint64_t *__restrict__ arena_off,
               const char *__restrict__ arena_seq, const char *__restrict__ arena_qual, const uint64_t *__restrict__ clip_off,
               const uint64_t *__restrict__ fq_off, char *__restrict__ clip, char *__restrict__ fq)
{
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_cl) return;
    uint32_t s = cl_seg[i], a = start[s], slot = cl_slot[i], k = a + slot;
    uint32_t x = order[a];
    uint32_t maxl = maxl_[s], stride = maxl + maxr_[s];
    const char *S = arena_seq + arena_off[s] + (uint64_t)slot * stride, *Q = arena_qual + arena_off[s] + (uint64_t)slot * stride;
    uint32_t L = out.len_l[k], R = out.len_r[k];
    bool noq = out.noqual[k];
    uint32_t side = c.side[x];
    // '5': aligned = right part, clipped = left part; '3': aligned = left, clipped = right (clip_reads.h:308-332)
    const char *aS = side == 0 ? S + maxl : S + maxl - L, *aQ = side == 0 ? Q + maxl : Q + maxl - L;
    const char *cS = side == 0 ? S + maxl - L : S + maxl, *cQ = side == 0 ? Q + maxl - L : Q + maxl;
    uint32_t aN = side == 0 ? R : L, cN = side == 0 ? L : R;
    uint32_t aQN = noq ? 1 : aN, cQN = noq ? 1 : cN;
    char *o = clip + clip_off[i];
    uint32_t head = 0;
    if (lane == 0) {
        int32_t tid = c.tid[x];
        const char *nm = names.blob + names.off[tid];
        uint32_t nl = names.off[tid + 1] - names.off[tid];
        char *q = o;
        for (uint32_t j = 0; j < nl; ++j) *q++ = nm[j];
        *q++ = '\t';
        q = put_dec_i(q, c.pos[x]);
        *q++ = '\t';
        *q++ = side == 0 ? '5' : '3';
        *q++ = '\t';
        q += cigar_text(d + out.cig_off[k], q);
        *q++ = '\t';
        head = (uint32_t)(q - o);
    }
    head = __shfl_sync(0xffffffffu, head, 0);
    char *q = o + head;
    for (uint32_t j = lane; j < aN; j += 32) q[j] = aS[j];
    q += aN;
    if (lane == 0) *q = '\t';
    ++q;
    for (uint32_t j = lane; j < aQN; j += 32) q[j] = noq ? '*' : aQ[j];
    q += aQN;
    if (lane == 0) *q = '\t';
    ++q;
    for (uint32_t j = lane; j < cN; j += 32) q[j] = cS[j];
    q += cN;
    if (lane == 0) *q = '\t';
    ++q;
    for (uint32_t j = lane; j < cQN; j += 32) q[j] = noq ? '*' : cQ[j];
    q += cQN;
    if (lane == 0) {
        *q++ = '\t';
        q = put_dec(q, out.support[k]);
        *q++ = '\n';
    }
    // FASTQ record named by its own sequence (clip_reads.h:320,339)
    char *f = fq + fq_off[i];
    if (lane == 0) f[0] = '@', f[1 + cN] = '\n', f[2 + 2 * cN] = '\n', f[3 + 2 * cN] = '+', f[4 + 2 * cN] = '\n', f[5 + 2 * cN + cQN] = '\n';
    for (uint32_t j = lane; j < cN; j += 32) f[1 + j] = cS[j], f[2 + cN + j] = cS[j];
    for (uint32_t j = lane; j < cQN; j += 32) f[5 + 2 * cN + j] = noq ? '*' : cQ[j];
}

// ---- unmapped-branch records: StoreUnmapSeqAndQual (clip_reads.h:172-219) on the device --------------------------
// The reference keeps a std::map<qname, held mate> over the whole file: the first record of a name is held; a later
// record of the same name and the OTHER end emits the pair (read1 to file 1, read2 to file 2) and erases the entry; a
// later record of the same end is ignored. Output order = file order of the completing record. Here: hash the
// names, stable-sort entries by hash (entries stay in file order inside a group), run the tiny per-name automaton
// with one thread per hash group (real name compares, so hash collisions cannot change the result), then emit text.
__device__ __forceinline__ uint32_t qname_len(const uint8_t *p) { return ldu32(p + 12) & 0xff; }

__global__ void unmapped_hash(uint32_t n, const uint64_t *__restrict__ list, const uint8_t *__restrict__ d,
                              uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = d + list[i];
    uint32_t lq = qname_len(p);
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint32_t j = 0; j < lq && p[36 + j]; ++j) h = (h ^ p[36 + j]) * 0x100000001b3ull;
    key[i] = h, val[i] = i;
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

// one thread per hash group; mate_of[e] = entry that was held when e completed a pair, else 0xffffffff
__global__ void unmapped_pair(uint32_t n, const uint64_t *__restrict__ key, const uint32_t *__restrict__ ent,
                              const uint64_t *__restrict__ list, const uint8_t *__restrict__ d, uint32_t *__restrict__ mate_of,
                              uint32_t *__restrict__ overflow)
{
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || (j > 0 && key[j] == key[j - 1])) return;
    const int H = 8;
    uint32_t held[H];
    int nh = 0;
    for (uint32_t k = j; k < n && key[k] == key[j]; ++k) {
        uint32_t e = ent[k];
        const uint8_t *p = d + list[e];
        bool r1 = (ldu32(p + 16) >> 16) & F_READ1;
        int hit = -1;
        for (int h = 0; h < nh; ++h)
            if (same_name(p, d + list[held[h]])) {
                hit = h;
                break;
            }
        if (hit < 0) {
            if (nh < H) held[nh++] = e;
            else atomicOr(overflow, 1u);
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

__global__ void unmapped_sizes(uint32_t n, const uint64_t *__restrict__ list, const uint32_t *__restrict__ mate_of,
                               const uint8_t *__restrict__ d, uint64_t *__restrict__ sz1, uint64_t *__restrict__ sz2)
{
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n) return;
    uint64_t a = 0, b = 0;
    if (e < n && mate_of[e] != 0xffffffffu) {
        const uint8_t *p = d + list[e], *q = d + list[mate_of[e]];
        bool p1 = (ldu32(p + 16) >> 16) & F_READ1;
        a = unmapped_fq_len(p1 ? p : q), b = unmapped_fq_len(p1 ? q : p);
    }
    sz1[e] = a, sz2[e] = b;
}

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
    unmapped_write(uint32_t n, const uint64_t *__restrict__ list, const uint32_t *__restrict__ mate_of, const uint8_t *__restrict__ d,
                   const uint64_t *__restrict__ off1, const uint64_t *__restrict__ off2, char *__restrict__ out1,
                   char *__restrict__ out2)
{
    uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n || mate_of[e] == 0xffffffffu) return;
    const uint8_t *p = d + list[e], *q = d + list[mate_of[e]];
    bool p1 = (ldu32(p + 16) >> 16) & F_READ1;
    write_unmapped_fq(p1 ? p : q, '1', out1 + off1[e], lane);
    write_unmapped_fq(p1 ? q : p, '2', out2 + off2[e], lane);
}

// ---- host orchestration -----------------------------------------------------------------------------------
static int sort_u64(svb_ctx *ctx, uint64_t *keys_in, uint64_t *keys_out, uint32_t n, int bits)
{
    size_t tmp = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, keys_in, keys_out, (int)n, 0, bits, ctx->stream));
    DevBuf<uint8_t> t;
    CK(t.alloc(tmp, ctx->stream));
    CK(cub::DeviceRadixSort::SortKeys(t.p, tmp, keys_in, keys_out, (int)n, 0, bits, ctx->stream));
    return 0;
}

static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

__global__ void count_below(uint32_t n, const uint64_t *__restrict__ off, uint64_t limit, uint32_t *__restrict__ count)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && off[i] < limit) atomicAdd(count, 1u);
}

// ---- shard plumbing: raw records of the unmapped branch, packed in file order ------------------------------------
__global__ void record_sizes(uint32_t n, const uint64_t *__restrict__ off, const uint8_t *__restrict__ d, uint64_t *__restrict__ size)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    size[i] = i < n ? 4ull + ldu32(d + off[i]) : 0ull;
}
__global__ void record_copy(uint32_t n, const uint64_t *__restrict__ off, const uint8_t *__restrict__ d, const uint64_t *__restrict__ dst_off,
                            uint8_t *__restrict__ dst)
{
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint8_t *src = d + off[w];
    uint8_t *out = dst + dst_off[w];
    const uint64_t bytes = dst_off[w + 1] - dst_off[w];
    for (uint64_t i = lane; i < bytes; i += 32) out[i] = src[i];
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

extern "C" int svb_getclip(svb_ctx *ctx, svb_bam *bam, const svb_getclip_params *prm, svb_clusters **out_)
{
    if (!ctx || !bam || !prm || !out_) return svb_fail(ctx, SVB_ERR_ARG, "svb_getclip: null argument");
    cudaStream_t s = ctx->stream;
    svb_clusters *res = new svb_clusters();
    res->ctx = ctx;
    std::unique_ptr<svb_clusters> guard(res);
    const uint64_t n_chunks = bam->n_chunks, stream_bytes = bam->nbytes - bam->first;
    int off_bits = 1;
    while (off_bits < 64 && (bam->nbytes >> off_bits)) ++off_bits;  // offsets sort on just the bits they use

    // ---- 1. the walker: every record head once ------------------------------------------------------------------
    DevBuf<uint32_t> counters;
    DevBuf<uint64_t> un_list, sw_list, c_off, exit_, first_mb, clipped;
    DevBuf<uint32_t> c_begin, c_ll, c_rl;
    DevBuf<int32_t> c_tid, c_pos, last_mb_tid;
    DevBuf<uint8_t> c_side;
    CandArrays c;
    uint32_t hc[4] = {0, 0, 0, 0};
    // sized from the stream (a record is at least ~40 bytes; ~2 % of them are soft-clipped), grown once on overflow
    uint64_t est = stream_bytes / 200;
    uint32_t cand_cap = (uint32_t)std::min<uint64_t>(est / 8 + 4096, 0xffffffffu);
    uint32_t un_cap = (uint32_t)std::min<uint64_t>(est / 8 + 4096, 0xffffffffu), sw_cap = 1 << 16;
    CK(counters.alloc(4, s));
    CK(exit_.alloc(n_chunks, s));
    CK(first_mb.alloc(n_chunks, s));
    CK(last_mb_tid.alloc(n_chunks, s));
    int attempt_walk = 0;
    for (int attempt = 0;; ++attempt) {
        CK(c_off.alloc(cand_cap, s));
        CK(c_begin.alloc(cand_cap, s));
        CK(c_ll.alloc(cand_cap, s));
        CK(c_rl.alloc(cand_cap, s));
        CK(c_tid.alloc(cand_cap, s));
        CK(c_pos.alloc(cand_cap, s));
        CK(c_side.alloc(cand_cap, s));
        CK(un_list.alloc(un_cap, s));
        CK(sw_list.alloc(sw_cap, s));
        CK(clipped.alloc(cand_cap, s));
        c = {c_off.p, c_tid.p, c_pos.p, c_begin.p, c_ll.p, c_rl.p, c_side.p};
        ScanOut so{c, cand_cap, un_list.p, un_cap, sw_list.p, sw_cap, counters.p, clipped.p, cand_cap, INT32_MIN, INT32_MIN, INT32_MAX, INT32_MAX};
        if (prm->key_filter) so.lo_tid = prm->key_lo_tid, so.lo_pos = prm->key_lo_pos, so.hi_tid = prm->key_hi_tid, so.hi_pos = prm->key_hi_pos;
        CK(cudaMemsetAsync(counters.p, 0, 16, s));
        if (stream_mode(bam) && attempt_walk == 0) {
            // streaming pass: its own first-record guesses go to bam->d_guess and are verified below like the walker's
            int per_sm = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, clip_stream, STREAM_THREADS, 0));
            unsigned grid = (unsigned)std::min<uint64_t>(n_chunks, (uint64_t)std::max(1, per_sm) * ctx->sm_count);
            ProfScope ps(ctx, "clip_stream", (double)stream_bytes);
            clip_stream<<<grid, STREAM_THREADS, 0, s>>>(bam->d_data, bam->nbytes, (bam->nbytes + 15) & ~15ull, bam->first, bam->n_ref,
                                                        n_chunks, bam->d_guess, bam->d_count, exit_.p, first_mb.p, last_mb_tid.p,
                                                        prm->min_mapq, so);
            clip_first<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, n_chunks, first_mb.p, last_mb_tid.p, prm->prev_tid, prm->min_mapq,
                                                         prm->save_low_quality, so);
            bam->guessed = true;  // d_guess now holds this pass's guesses (verified or repaired below)
        } else {
            CKR(ensure_guess(ctx, bam));
            ProfScope ps(ctx, "clip_walk", (double)stream_bytes);
            clip_walk<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, bam->nbytes, n_chunks, bam->chunk_log2, bam->d_guess, bam->d_count,
                                                        exit_.p, first_mb.p, last_mb_tid.p, prm->min_mapq, prm->save_low_quality, so);
            clip_first<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, n_chunks, first_mb.p, last_mb_tid.p, prm->prev_tid, prm->min_mapq,
                                                         prm->save_low_quality, so);
        }
        {
            // at most cand_cap queued records are evaluated; an overflow of the queue is caught below and the pass repeated
            ProfScope ps(ctx, "clip_eval", 0);
            clip_eval<<<nblk(cand_cap, 128), 128, 0, s>>>(bam->d_data, prm->min_mapq, prm->save_low_quality, so);
        }
        int ok = 0;
        CKR(verify_or_repair(ctx, bam, exit_.p, &ok));  // (synchronises); a failed verification repairs the guesses
        if (!ok) attempt_walk = 1;                      // ... and the pass is repeated by the walker, which starts from them
        CK(cudaMemcpyAsync(hc, counters.p, 16, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        bool fits = hc[0] <= cand_cap && hc[3] <= cand_cap && hc[1] <= un_cap && hc[2] <= sw_cap;
        if (ok && fits) break;
        if (attempt >= 3) return svb_fail(ctx, SVB_ERR_CUDA, "svb_getclip: the record walk did not settle");
        if (ok) {
            cand_cap = std::max(cand_cap, std::max(hc[0], 2 * hc[3]));  // a record yields at most two candidates
            un_cap = std::max(un_cap, hc[1]), sw_cap = std::max(sw_cap, hc[2]);
        }
    }
    if (!bam->counted) CKR(finish_counts(ctx, bam, exit_.p));  // the walker counted the records of every chunk on its way
    const uint32_t n_cand = hc[0], n_sw = hc[2];
    uint32_t n_un = hc[1];
    res->n_candidates = n_cand;

    // ---- 2. unmapped-branch records: pair mates by name on the device, emit the two FASTQ files ---------------------
    DevBuf<char> un_o1, un_o2;
    bool have_unmapped_gz = false, unmapped_copy_pending = false;
    uint64_t un_bytes[2] = {0, 0};
    auto start_unmapped_copy = [&]() -> int {
        if (!unmapped_copy_pending) return 0;
        unmapped_copy_pending = false;
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->fork_event, 0));
        CK(cudaMemcpyAsync(res->text[2].p, un_o1.p, un_bytes[0], cudaMemcpyDeviceToHost, ctx->copy_stream));
        CK(cudaMemcpyAsync(res->text[3].p, un_o2.p, un_bytes[1], cudaMemcpyDeviceToHost, ctx->copy_stream));
        return 0;
    };
    res->gz_mode = prm->gz_outputs != 0 && !prm->export_unmapped_records;
    struct CopyJoin {  // declared after the buffers the copy stream reads: joined before they are released, on every way out
        cudaStream_t c;
        ~CopyJoin() { cudaStreamSynchronize(c); }
    } copy_join{ctx->copy_stream};
    DevBuf<uint32_t> val0, val1, mate_of, ovf;
    DevBuf<uint64_t> un_sorted, ukey0, ukey1, sz1, sz2, off1, off2;
    if (n_un) {
        CK(un_sorted.alloc(n_un, s));
        CKR(sort_u64(ctx, un_list.p, un_sorted.p, n_un, off_bits));
    }
    const uint64_t *un_first = un_sorted.p;  // (the kernels below read un_sorted through this pointer)
    if (n_un && prm->halo_bytes) {
        // range shards: records in front of the shard's own region only lend their soft clips; the list is sorted by offset
        DevBuf<uint32_t> skip;
        CK(skip.alloc(1, s));
        CK(cudaMemsetAsync(skip.p, 0, 4, s));
        count_below<<<nblk(n_un, 256), 256, 0, s>>>(n_un, un_sorted.p, prm->halo_bytes, skip.p);
        uint32_t h_skip = 0;
        CK(cudaMemcpyAsync(&h_skip, skip.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        un_first += h_skip, n_un -= h_skip;
    }
    if (n_un && prm->export_unmapped_records) {  // a shard: hand the records to the merging rank instead of pairing here
        DevBuf<uint64_t> sz, ro;
        CK(sz.alloc(n_un + 1, s));
        CK(ro.alloc(n_un + 1, s));
        record_sizes<<<nblk(n_un + 1, 256), 256, 0, s>>>(n_un, un_first, bam->d_data, sz.p);
        CKR(exclusive_scan_u64(ctx, sz.p, ro.p, n_un + 1));
        uint64_t total = 0;
        CK(cudaMemcpyAsync(&total, ro.p + n_un, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        CK(un_o1.alloc(total, s));
        record_copy<<<nblk((uint64_t)n_un * 32, 128), 128, 0, s>>>(n_un, un_first, bam->d_data, ro.p, (uint8_t *)un_o1.p);
        CKR(res->unmapped_records.reserve(ctx, total));
        CK(cudaMemcpyAsync(res->unmapped_records.p, un_o1.p, total, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    } else if (n_un) {
        CK(val0.alloc(n_un, s));
        CK(val1.alloc(n_un, s));
        CK(ukey0.alloc(n_un, s));
        CK(ukey1.alloc(n_un, s));
        CK(mate_of.alloc(n_un, s));
        CK(ovf.alloc(1, s));
        CK(sz1.alloc(n_un + 1, s));
        CK(sz2.alloc(n_un + 1, s));
        CK(off1.alloc(n_un + 1, s));
        CK(off2.alloc(n_un + 1, s));
        CK(cudaMemsetAsync(mate_of.p, 0xff, (size_t)n_un * 4, s));
        CK(cudaMemsetAsync(ovf.p, 0, 4, s));
        uint64_t tot[2] = {0, 0};
        {
            ProfScope ps(ctx, "unmapped_pair", 0);
            unmapped_hash<<<nblk(n_un, 256), 256, 0, s>>>(n_un, un_first, bam->d_data, ukey0.p, val0.p);
            size_t tmp = 0;
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, ukey0.p, ukey1.p, val0.p, val1.p, (int)n_un, 0, 64, s));
            DevBuf<uint8_t> t;
            CK(t.alloc(tmp, s));
            CK(cub::DeviceRadixSort::SortPairs(t.p, tmp, ukey0.p, ukey1.p, val0.p, val1.p, (int)n_un, 0, 64, s));
            unmapped_pair<<<nblk(n_un, 128), 128, 0, s>>>(n_un, ukey1.p, val1.p, un_first, bam->d_data, mate_of.p, ovf.p);
            unmapped_sizes<<<nblk(n_un + 1, 256), 256, 0, s>>>(n_un, un_first, mate_of.p, bam->d_data, sz1.p, sz2.p);
            CKR(exclusive_scan_u64(ctx, sz1.p, off1.p, n_un + 1));
            CKR(exclusive_scan_u64(ctx, sz2.p, off2.p, n_un + 1));
        }
        uint32_t hovf = 0;
        CK(cudaMemcpyAsync(&tot[0], off1.p + n_un, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&tot[1], off2.p + n_un, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&hovf, ovf.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (hovf) return svb_fail(ctx, SVB_ERR_FORMAT, "more than 8 distinct read names share one 64-bit name hash");
        CK(un_o1.alloc(tot[0], s));
        CK(un_o2.alloc(tot[1], s));
        {
            ProfScope ps(ctx, "unmapped_write", (double)(tot[0] + tot[1]));
            unmapped_write<<<nblk((uint64_t)n_un * 32, 128), 128, 0, s>>>(n_un, un_first, mate_of.p, bam->d_data, off1.p, off2.p,
                                                                          un_o1.p, un_o2.p);
        }
        if (res->gz_mode) {
            CKR(gzip_on_device(ctx, un_o1.p, tot[0], &res->gz[2]));
            CKR(gzip_on_device(ctx, un_o2.p, tot[1], &res->gz[3]));
            have_unmapped_gz = true;
        } else {
        CKR(res->text[2].reserve(ctx, tot[0]));
        CKR(res->text[3].reserve(ctx, tot[1]));
        // the two FASTQ texts travel to the host on the copy stream while the candidate pipeline below runs; the copies are
        // queued after the sorts (start_unmapped_copy): queued here they slowed the sort's many small launches down by 2x
        CK(cudaEventRecord(ctx->fork_event, s));
        unmapped_copy_pending = true, un_bytes[0] = tot[0], un_bytes[1] = tot[1];
        }
    }
    // gzip mode: every file exists, if only as one empty member (what the reference's ogzstream leaves behind as well)
    auto empty_gz = [&](int which) -> int { return res->gz[which].p ? 0 : gzip_on_device(ctx, nullptr, 0, &res->gz[which]); };
    if (res->gz_mode && !have_unmapped_gz) {
        CKR(empty_gz(2));
        CKR(empty_gz(3));
    }

    if (n_cand == 0) {
        CKR(start_unmapped_copy());
        if (res->gz_mode) {
            CKR(empty_gz(0));
            CKR(empty_gz(1));
        }
        *out_ = guard.release();
        return 0;
    }
    if (bam->names.size() != (size_t)bam->n_ref) return svb_fail(ctx, SVB_ERR_ARG, "svb_getclip: reference names not set (svb_bam_set_refs)");

    // ---- 3. order candidates: BAM order first (stable base), then (run, side, pos) -----------------------------
    DevBuf<uint32_t> ord0, ord1, ord2;
    DevBuf<uint64_t> sw_sorted, rec_sorted, key0, key1;
    CK(sw_sorted.alloc(n_sw, s));
    if (n_sw) CKR(sort_u64(ctx, sw_list.p, sw_sorted.p, n_sw, off_bits));
    CK(ord0.alloc(n_cand, s));
    CK(ord1.alloc(n_cand, s));
    CK(ord2.alloc(n_cand, s));
    CK(rec_sorted.alloc(n_cand, s));
    CK(key0.alloc(n_cand, s));
    CK(key1.alloc(n_cand, s));
    iota_u32<<<nblk(n_cand, 256), 256, 0, s>>>(n_cand, ord0.p);
    {
        // a both-side-clipped read yields two candidates with the same record index but different sides, so
        // (key, record) is unique and the two-pass stable sort is deterministic
        ProfScope ps(ctx, "sort_candidates", (double)n_cand * 24);
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c.off, rec_sorted.p, ord0.p, ord1.p, (int)n_cand, 0, off_bits, s));
        DevBuf<uint8_t> t;
        CK(t.alloc(tmp, s));
        CK(cub::DeviceRadixSort::SortPairs(t.p, tmp, c.off, rec_sorted.p, ord0.p, ord1.p, (int)n_cand, 0, off_bits, s));
        make_keys<<<nblk(n_cand, 256), 256, 0, s>>>(n_cand, ord1.p, c, sw_sorted.p, n_sw, key0.p);
        size_t tmp2 = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, key0.p, key1.p, ord1.p, ord2.p, (int)n_cand, 0, 64, s));
        DevBuf<uint8_t> t2;
        CK(t2.alloc(tmp2, s));
        CK(cub::DeviceRadixSort::SortPairs(t2.p, tmp2, key0.p, key1.p, ord1.p, ord2.p, (int)n_cand, 0, 64, s));
    }
    const uint32_t *order = ord2.p;
    CKR(start_unmapped_copy());

    // ---- 4. segments (one per breakpoint key) -------------------------------------------------------------------
    DevBuf<uint32_t> flag, segid;
    CK(flag.alloc(n_cand, s));
    CK(segid.alloc(n_cand, s));
    seg_flags<<<nblk(n_cand, 256), 256, 0, s>>>(n_cand, key1.p, flag.p);
    CKR(inclusive_scan_u32(ctx, flag.p, segid.p, n_cand));
    uint32_t n_seg = 0;
    CK(cudaMemcpyAsync(&n_seg, segid.p + (n_cand - 1), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    DevBuf<uint32_t> start, maxl, maxr;
    DevBuf<uint64_t> seg_bytes, arena_off;
    CK(start.alloc(n_seg + 1, s));
    CK(maxl.alloc(n_seg, s));
    CK(maxr.alloc(n_seg, s));
    CK(seg_bytes.alloc(n_seg + 1, s));
    CK(arena_off.alloc(n_seg + 1, s));
    seg_starts<<<nblk(n_cand, 256), 256, 0, s>>>(n_cand, flag.p, segid.p, start.p, n_seg);
    seg_stats<<<nblk(n_seg, 256), 256, 0, s>>>(n_seg, start.p, order, c, maxl.p, maxr.p, seg_bytes.p);
    CKR(exclusive_scan_u64(ctx, seg_bytes.p, arena_off.p, n_seg + 1));
    uint64_t arena_bytes = 0;
    CK(cudaMemcpyAsync(&arena_bytes, arena_off.p + n_seg, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));

    // ---- 5. greedy clustering -----------------------------------------------------------------------------------
    DevBuf<char> arena_seq, arena_qual;
    DevBuf<uint32_t> len_l, len_r, support, seg_ncl, cl_base;
    DevBuf<uint64_t> cig_off;
    DevBuf<uint8_t> noqual;
    CK(arena_seq.alloc(arena_bytes, s));
    CK(arena_qual.alloc(arena_bytes, s));
    CK(len_l.alloc(n_cand, s));
    CK(len_r.alloc(n_cand, s));
    CK(cig_off.alloc(n_cand, s));
    CK(support.alloc(n_cand, s));
    CK(noqual.alloc(n_cand, s));
    CK(seg_ncl.alloc(n_seg, s));
    CK(cl_base.alloc(n_seg, s));
    ClusterOut co{len_l.p, len_r.p, support.p, cig_off.p, noqual.p, seg_ncl.p};
    {
        ProfScope ps(ctx, "cluster_build", (double)arena_bytes * 2);
        cluster_build<<<nblk((uint64_t)n_seg * 32, 128), 128, 0, s>>>(bam->d_data, n_seg, start.p, order, c, maxl.p,
                                                                      maxr.p, arena_off.p, arena_seq.p, arena_qual.p, prm->match_rate, co);
    }
    CKR(inclusive_scan_u32(ctx, seg_ncl.p, cl_base.p, n_seg));
    uint32_t n_cl = 0;
    CK(cudaMemcpyAsync(&n_cl, cl_base.p + (n_seg - 1), 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    res->n_clusters = n_cl;

    // ---- 6. text ---------------------------------------------------------------------------------------------------
    std::vector<uint32_t> noff(bam->n_ref + 1, 0);
    std::string nblob;
    for (int32_t t = 0; t < bam->n_ref; ++t) {
        noff[t] = (uint32_t)nblob.size();
        nblob += bam->names[t];
    }
    noff[bam->n_ref] = (uint32_t)nblob.size();
    DevBuf<char> d_nblob;
    DevBuf<uint32_t> d_noff, cl_seg, cl_slot;
    DevBuf<uint64_t> clip_len, fq_len, clip_off, fq_off;
    CK(d_nblob.alloc(nblob.size() + 1, s));
    CK(d_noff.alloc(noff.size(), s));
    CK(cudaMemcpyAsync(d_nblob.p, nblob.data(), nblob.size(), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_noff.p, noff.data(), noff.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cl_seg.alloc(n_cl, s));
    CK(cl_slot.alloc(n_cl, s));
    CK(clip_len.alloc(n_cl + 1, s));
    CK(fq_len.alloc(n_cl + 1, s));
    CK(clip_off.alloc(n_cl + 1, s));
    CK(fq_off.alloc(n_cl + 1, s));
    NameTable nt{d_nblob.p, d_noff.p};
    list_clusters<<<nblk(n_seg, 256), 256, 0, s>>>(n_seg, start.p, seg_ncl.p, cl_base.p, cl_seg.p, cl_slot.p);
    text_sizes<<<nblk(n_cl + 1, 256), 256, 0, s>>>(n_cl, cl_seg.p, cl_slot.p, start.p, order, c, co, bam->d_data, nt,
                                                   clip_len.p, fq_len.p);
    CKR(exclusive_scan_u64(ctx, clip_len.p, clip_off.p, n_cl + 1));
    CKR(exclusive_scan_u64(ctx, fq_len.p, fq_off.p, n_cl + 1));
    uint64_t clip_bytes = 0, fq_bytes = 0;
    CK(cudaMemcpyAsync(&clip_bytes, clip_off.p + n_cl, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&fq_bytes, fq_off.p + n_cl, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    DevBuf<char> d_clip, d_fq;
    CK(d_clip.alloc(clip_bytes, s));
    CK(d_fq.alloc(fq_bytes, s));
    {
        ProfScope ps(ctx, "text_write", (double)(clip_bytes + fq_bytes));
        text_write<<<nblk((uint64_t)n_cl * 32, 128), 128, 0, s>>>(n_cl, cl_seg.p, cl_slot.p, start.p, order, c, co, bam->d_data, nt,
                                                                  maxl.p, maxr.p, arena_off.p, arena_seq.p, arena_qual.p, clip_off.p,
                                                                  fq_off.p, d_clip.p, d_fq.p);
    }
    if (res->gz_mode) {
        CKR(gzip_on_device(ctx, d_clip.p, clip_bytes, &res->gz[0]));
        CKR(gzip_on_device(ctx, d_fq.p, fq_bytes, &res->gz[1]));
    } else {
        CKR(res->text[0].reserve(ctx, clip_bytes));
        CKR(res->text[1].reserve(ctx, fq_bytes));
        CK(cudaMemcpyAsync(res->text[0].p, d_clip.p, clip_bytes, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(res->text[1].p, d_fq.p, fq_bytes, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    *out_ = guard.release();
    return 0;
}

extern "C" void svb_clusters_free(svb_clusters *c) { delete c; }
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

extern "C" int svb_clusters_unmapped_records(const svb_clusters *c, const char **data, uint64_t *len)
{
    if (!c || !data || !len) return SVB_ERR_ARG;
    *data = c->unmapped_records.p ? c->unmapped_records.p : "";
    *len = c->unmapped_records.n;
    return 0;
}

extern "C" int svb_clusters_text(const svb_clusters *c, int which, const char **data, uint64_t *len)
{
    if (!c || which < 0 || which > 3 || !data || !len) return SVB_ERR_ARG;
    *data = c->text[which].p ? c->text[which].p : "";
    *len = c->text[which].n;
    return 0;
}
