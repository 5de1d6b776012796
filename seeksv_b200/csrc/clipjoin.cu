// clip_join on the device: the join of P.clip.gz lines with the realigned clip alignments that
// the reference does in the lock-step loop of InputSoftInfoStoreBreakpoint<T> (getsv.h:423-541), GetAlignInfo (getsv.cpp:25-71)
// and the key rules of GetJunction (getsv.cpp:1705-1845). The rules live in clipjoin_core.h (host + device, checked on the CPU
// against the host mirror by tools/clipjoin_sim.cpp); this file is the parallel plumbing around them:
//
//   1. run heads of the lines / block starts of the alignments: one flag per element, compacted by a chained scan;
//   2. the breakers: CJ_CHUNK run boundaries per thread from a guessed entry, verified chunk to chunk (exit(c) == entry(c + 1)),
//      wrong guesses repaired from the predecessor's exit - the BAM record walker's idiom (walk.cu); well-formed input (blocks of
//      equally named alignments in step with the runs) needs no repair round;
//   3. per run: the alignment range behind its set -> member count -> prefix sum -> members in the map's iteration order
//      (insertion sort per run: a handful of alignments), each classified and keyed against the run's head line;
//   4. compaction of the stored candidates in crossing order, then two stable radix sorts (positions, then chromosome ranks +
//      strands) = Junction::operator< order with the crossing order kept inside a key.
// Sync points: one read-back after 2 (run count, verify flag), one after 4 (candidate count), then the candidates.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "clipjoin_core.h"
#include "prim.cuh"

namespace {
struct CjCtl {
    uint32_t n_runs, n_blocks, bad, too_large;
    uint32_t n_cands, pad[3];
};

struct FlagScanOp {  // compaction of the elements whose flag is set: heads[excl] = i
    CjView v;
    uint64_t count;
    int which;  // 0: line runs, 1: alignment blocks
    uint32_t *heads;
    uint32_t *total_out;
    __device__ uint64_t n() const { return count; }
    __device__ void load(uint64_t i, uint64_t (&x)[1]) const { x[0] = which == 0 ? cj_line_starts_run(v, i) : cj_aln_starts_block(v, i); }
    __device__ void store(uint64_t i, const uint64_t (&excl)[1], const uint64_t (&x)[1]) const
    {
        if (x[0]) heads[excl[0]] = (uint32_t)i;
    }
    __device__ void total(const uint64_t (&t)[1]) const { *total_out = (uint32_t)t[0]; }
};

__device__ __forceinline__ void chunk_bounds(uint32_t n_runs, uint64_t c, uint64_t *k0, uint64_t *k1)
{
    *k0 = 1 + c * CJ_CHUNK;
    *k1 = min((uint64_t)n_runs, *k0 + CJ_CHUNK);
}

__global__ void cj_walk(CjView v, const CjCtl *ctl, const uint32_t *__restrict__ run_head, const uint32_t *__restrict__ block_start,
                        uint32_t *__restrict__ breaker, uint64_t *__restrict__ entry, uint64_t *__restrict__ exit_)
{
    const uint32_t n_runs = ctl->n_runs, n_blocks = ctl->n_blocks;
    if (n_runs < 2) return;
    const uint64_t n_chunks = ((uint64_t)n_runs - 1 + CJ_CHUNK - 1) / CJ_CHUNK;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k0, k1;
        chunk_bounds(n_runs, c, &k0, &k1);
        const uint64_t e = c == 0 ? 0 : cj_guess_entry(v, run_head, block_start, n_blocks, k0);
        entry[c] = e;
        exit_[c] = cj_walk_chunk(v, run_head, k0, k1, e, breaker);
    }
}

__global__ void cj_verify(const CjCtl *ctl, const uint64_t *__restrict__ entry, const uint64_t *__restrict__ exit_, uint32_t *bad)
{
    const uint32_t n_runs = ctl->n_runs;
    if (n_runs < 2) return;
    const uint64_t n_chunks = ((uint64_t)n_runs - 1 + CJ_CHUNK - 1) / CJ_CHUNK;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x)
        if (entry[c] != exit_[c - 1]) atomicOr(bad, 1u);
}

// every chunk whose entry differs from its predecessor's exit (as it was before this round) walks again from there; the first
// wrong chunk of a run of wrong chunks always gets its true entry, so the rounds converge (at most one per chunk)
__global__ void cj_repair(CjView v, const CjCtl *ctl, const uint32_t *__restrict__ run_head, uint32_t *__restrict__ breaker,
                          uint64_t *__restrict__ entry, uint64_t *__restrict__ exit_, const uint64_t *__restrict__ exit_prev)
{
    const uint32_t n_runs = ctl->n_runs;
    if (n_runs < 2) return;
    const uint64_t n_chunks = ((uint64_t)n_runs - 1 + CJ_CHUNK - 1) / CJ_CHUNK;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = exit_prev[c - 1];
        if (entry[c] == e) continue;
        uint64_t k0, k1;
        chunk_bounds(n_runs, c, &k0, &k1);
        entry[c] = e;
        exit_[c] = cj_walk_chunk(v, run_head, k0, k1, e, breaker);
    }
}

struct MemberScanOp {  // members of every run's set (before duplicates are dropped) -> slot offsets
    CjView v;
    const CjCtl *ctl;
    const uint32_t *run_head, *breaker;
    uint32_t *slot_off;
    __device__ uint64_t n() const { return ctl->n_runs; }
    __device__ void load(uint64_t k, uint64_t (&x)[1]) const
    {
        uint64_t lo, hi;
        bool hb, all, crossed;
        cj_run_range(v, run_head, breaker, ctl->n_runs, k, &lo, &hi, &hb, &all, &crossed);
        x[0] = crossed ? (uint64_t)hb + cj_count_members(v, lo, hi, all) : 0;
    }
    __device__ void store(uint64_t k, const uint64_t (&excl)[1], const uint64_t (&x)[1]) const
    {
        slot_off[k] = (uint32_t)excl[0];
        if (k + 1 == ctl->n_runs) slot_off[k + 1] = (uint32_t)(excl[0] + x[0]);
    }
    __device__ void total(const uint64_t (&)[1]) const {}
};

// one thread per run: its set in iteration order, every member classified against the head line. Slots whose member stores
// nothing (unmapped alignment, dropped duplicate) stay invalid.
__global__ void cj_members(CjView v, CjCtl *ctl, const uint32_t *__restrict__ run_head, const uint32_t *__restrict__ breaker,
                           const uint32_t *__restrict__ slot_off, uint32_t *__restrict__ members, svb_join_cand *__restrict__ slot_cand,
                           uint8_t *__restrict__ slot_valid)
{
    const uint32_t n_runs = ctl->n_runs;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_runs; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t s0 = slot_off[k], cnt = slot_off[k + 1] - s0;
        if (!cnt) continue;
        if (cnt > CJ_MAX_SET) {
            ctl->too_large = 1;
            continue;
        }
        uint64_t lo, hi;
        bool hb, all, crossed;
        cj_run_range(v, run_head, breaker, n_runs, k, &lo, &hi, &hb, &all, &crossed);
        const uint32_t kept = cj_fill_members(v, run_head, k, lo, hi, hb, all, hb ? breaker[k] : 0u, members + s0);
        for (uint32_t t = 0; t < kept; ++t) {
            svb_join_cand c;
            if (cj_classify(v, run_head[k], members[s0 + t], &c)) slot_cand[s0 + t] = c, slot_valid[s0 + t] = 1;
        }
    }
}

struct CandScanOp {  // stored candidates, compacted in crossing order; sort input of the first (low key) sort
    const CjCtl *ctl;
    const uint32_t *slot_off;  // slot_off[n_runs] = slots in use
    const uint8_t *slot_valid;
    const svb_join_cand *slot_cand;
    svb_join_cand *cand;
    uint64_t *key;
    uint32_t *val;
    uint32_t *n_out;
    __device__ uint64_t n() const { return ctl->too_large ? 0 : slot_off[ctl->n_runs]; }
    __device__ void load(uint64_t i, uint64_t (&x)[1]) const { x[0] = slot_valid[i]; }
    __device__ void store(uint64_t i, const uint64_t (&excl)[1], const uint64_t (&x)[1]) const
    {
        if (!x[0]) return;
        const svb_join_cand c = slot_cand[i];
        cand[excl[0]] = c, key[excl[0]] = cj_key_low(c), val[excl[0]] = (uint32_t)excl[0];
    }
    __device__ void total(const uint64_t (&t)[1]) const { *n_out = (uint32_t)t[0]; }
};

__global__ void cj_high_keys(const uint32_t *__restrict__ n_ptr, const uint32_t *__restrict__ order, const svb_join_cand *__restrict__ cand,
                             uint64_t *__restrict__ key, uint32_t *__restrict__ val)
{
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) key[i] = cj_key_high(cand[order[i]]), val[i] = order[i];
}
__global__ void cj_gather(const uint32_t *__restrict__ n_ptr, const uint32_t *__restrict__ order, const svb_join_cand *__restrict__ cand,
                          svb_join_cand *__restrict__ out)
{
    const uint32_t n = *n_ptr;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = cand[order[i]];
}

struct CjBuffers {
    CjCtl *ctl;
    ScanScratch sc_runs, sc_blocks, sc_members, sc_cands;
    RadixScratch rs_low, rs_high;
    uint8_t *slot_valid;
    size_t zero_end;
    svb_join_line *lines;
    svb_join_aln *alns;
    char *seqs, *names;
    uint32_t *cigars;
    uint32_t *run_head, *block_start, *breaker, *slot_off, *members;
    uint64_t *entry, *exit_, *exit_prev;
    svb_join_cand *slot_cand, *cand, *sorted;
    uint64_t *key[2];
    uint32_t *val[2];
};

void carve(Bump &b, CjBuffers &B, uint64_t n_lines, uint64_t seq_bytes, uint64_t n_alns, uint64_t name_bytes, uint64_t n_words)
{
    const uint64_t slots = n_alns + 1, chunks = n_lines / CJ_CHUNK + 2;
    B.ctl = b.get<CjCtl>(1);
    B.sc_runs = scan_scratch(b, n_lines, 1, 2);
    B.sc_blocks = scan_scratch(b, n_alns, 1, 2);
    B.sc_members = scan_scratch(b, n_lines, 1, 2);
    B.sc_cands = scan_scratch(b, slots, 1, 8);
    B.rs_low = radix_scratch(b, slots, 8);
    B.rs_high = radix_scratch(b, slots, 8);
    B.slot_valid = b.get<uint8_t>(slots);
    B.zero_end = (b.used + 255) & ~(size_t)255;
    B.lines = b.get<svb_join_line>(n_lines), B.alns = b.get<svb_join_aln>(n_alns);
    B.seqs = b.get<char>(seq_bytes), B.names = b.get<char>(name_bytes), B.cigars = b.get<uint32_t>(n_words);
    B.run_head = b.get<uint32_t>(n_lines + 1), B.block_start = b.get<uint32_t>(n_alns + 1);
    B.breaker = b.get<uint32_t>(n_lines + 1), B.slot_off = b.get<uint32_t>(n_lines + 2), B.members = b.get<uint32_t>(slots);
    B.entry = b.get<uint64_t>(chunks), B.exit_ = b.get<uint64_t>(chunks), B.exit_prev = b.get<uint64_t>(chunks);
    B.slot_cand = b.get<svb_join_cand>(slots), B.cand = b.get<svb_join_cand>(slots), B.sorted = b.get<svb_join_cand>(slots);
    for (int i = 0; i < 2; ++i) B.key[i] = b.get<uint64_t>(slots), B.val[i] = b.get<uint32_t>(slots);
}
}  // namespace

extern "C" int svb_clip_join(svb_ctx *ctx, const svb_join_line *lines, uint64_t n_lines, const char *seqs, uint64_t seq_bytes,
                             const svb_join_aln *alns, uint64_t n_alns, const char *names, uint64_t name_bytes, const uint32_t *cigars,
                             uint64_t n_words, svb_join_cand **cands, uint64_t *n_cands)
{
    if (!ctx || !cands || !n_cands || (n_lines && (!lines || !seqs)) || (n_alns && (!alns || !names)) || (n_words && !cigars))
        return svb_fail(ctx, SVB_ERR_ARG, "svb_clip_join: null argument");
    if (n_lines >= (1ull << 27) || n_alns >= (1ull << 27) || seq_bytes >= (1ull << 32) || name_bytes >= (1ull << 32) || n_words >= (1ull << 32))
        return svb_fail(ctx, SVB_ERR_ARG, "svb_clip_join: more than 2^27 lines / alignments or 4 GiB of text");
    *cands = nullptr, *n_cands = 0;
    if (n_lines == 0) {
        *cands = (svb_join_cand *)malloc(sizeof(svb_join_cand));
        return 0;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    CjBuffers B{};
    {
        Bump measure(nullptr);
        carve(measure, B, n_lines, seq_bytes, n_alns, name_bytes, n_words);
        CKR(ctx->ws_reserve(0, measure.used));
        Bump real(ctx->ws[0]);
        carve(real, B, n_lines, seq_bytes, n_alns, name_bytes, n_words);
    }
    CK(cudaMemsetAsync(ctx->ws[0], 0, B.zero_end, s));
    CK(cudaMemcpyAsync(B.lines, lines, n_lines * sizeof(svb_join_line), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(B.seqs, seqs, seq_bytes, cudaMemcpyHostToDevice, s));
    if (n_alns) CK(cudaMemcpyAsync(B.alns, alns, n_alns * sizeof(svb_join_aln), cudaMemcpyHostToDevice, s));
    if (name_bytes) CK(cudaMemcpyAsync(B.names, names, name_bytes, cudaMemcpyHostToDevice, s));
    if (n_words) CK(cudaMemcpyAsync(B.cigars, cigars, n_words * 4, cudaMemcpyHostToDevice, s));
    const CjView v{B.lines, n_lines, B.seqs, B.alns, n_alns, B.names, B.cigars};
    CjCtl *ctl = B.ctl;
    const uint64_t chunk_cap = n_lines / CJ_CHUNK + 1;
    {
        ProfScope ps(ctx, "clip_join_walk", 0);
        FlagScanOp runs{v, n_lines, 0, B.run_head, &ctl->n_runs}, blocks{v, n_alns, 1, B.block_start, &ctl->n_blocks};
        launch_scan<1, 2>(ctx, s, runs, B.sc_runs, n_lines);
        launch_scan<1, 2>(ctx, s, blocks, B.sc_blocks, std::max<uint64_t>(n_alns, 1));
        cj_walk<<<grid_for(ctx, chunk_cap, 64, 8), 64, 0, s>>>(v, ctl, B.run_head, B.block_start, B.breaker, B.entry, B.exit_);
        cj_verify<<<grid_for(ctx, chunk_cap, 128, 8), 128, 0, s>>>(ctl, B.entry, B.exit_, &ctl->bad);
    }
    CjCtl h{};
    CK(cudaMemcpyAsync(ctx->ctl_host, ctl, sizeof(CjCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    memcpy(&h, ctx->ctl_host, sizeof h);
    for (uint64_t round = 0; h.bad; ++round) {
        if (round > chunk_cap + 1) return svb_fail(ctx, SVB_ERR_CUDA, "svb_clip_join: the boundary walk did not settle");
        ProfScope ps(ctx, "clip_join_repair", 0);
        CK(cudaMemcpyAsync(B.exit_prev, B.exit_, chunk_cap * 8, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemsetAsync(&ctl->bad, 0, 4, s));
        cj_repair<<<grid_for(ctx, chunk_cap, 64, 8), 64, 0, s>>>(v, ctl, B.run_head, B.breaker, B.entry, B.exit_, B.exit_prev);
        cj_verify<<<grid_for(ctx, chunk_cap, 128, 8), 128, 0, s>>>(ctl, B.entry, B.exit_, &ctl->bad);
        CK(cudaMemcpyAsync(ctx->ctl_host, ctl, sizeof(CjCtl), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        memcpy(&h, ctx->ctl_host, sizeof h);
    }
    const uint64_t slots = n_alns + 1;
    {
        ProfScope ps(ctx, "clip_join_members", 0);
        MemberScanOp mo{v, ctl, B.run_head, B.breaker, B.slot_off};
        launch_scan<1, 2>(ctx, s, mo, B.sc_members, n_lines);
        cj_members<<<grid_for(ctx, n_lines, 64, 16), 64, 0, s>>>(v, ctl, B.run_head, B.breaker, B.slot_off, B.members, B.slot_cand, B.slot_valid);
    }
    {
        ProfScope ps(ctx, "clip_join_sort", 0);
        CandScanOp co{ctl, B.slot_off, B.slot_valid, B.slot_cand, B.cand, B.key[0], B.val[0], &ctl->n_cands};
        launch_scan<1, 8>(ctx, s, co, B.sc_cands, slots);
        RadixJob j1{{B.key[0], B.key[1]}, {B.val[0], B.val[1]}, &ctl->n_cands, (uint32_t)slots, 0, 8, B.rs_low};
        radix_sort(ctx, s, j1);
        cj_high_keys<<<grid_for(ctx, slots, 256, 4), 256, 0, s>>>(&ctl->n_cands, B.val[1], B.cand, B.key[0], B.val[0]);
        RadixJob j2{{B.key[0], B.key[1]}, {B.val[0], B.val[1]}, &ctl->n_cands, (uint32_t)slots, 0, 8, B.rs_high};
        radix_sort(ctx, s, j2);
        cj_gather<<<grid_for(ctx, slots, 256, 4), 256, 0, s>>>(&ctl->n_cands, B.val[1], B.cand, B.sorted);
    }
    CK(cudaMemcpyAsync(ctx->ctl_host, ctl, sizeof(CjCtl), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    memcpy(&h, ctx->ctl_host, sizeof h);
    if (h.too_large)
        return svb_fail(ctx, SVB_ERR_FORMAT, "svb_clip_join: a run of clip lines has more than %u alignments (the host join handles such input)", CJ_MAX_SET);
    svb_join_cand *out = (svb_join_cand *)malloc(std::max<size_t>(1, (size_t)h.n_cands) * sizeof(svb_join_cand));
    if (!out) return svb_fail(ctx, SVB_ERR_ARG, "svb_clip_join: out of host memory");
    if (h.n_cands) {
        cudaError_t e = cudaMemcpyAsync(out, B.sorted, (size_t)h.n_cands * sizeof(svb_join_cand), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            free(out);
            return svb_fail(ctx, SVB_ERR_CUDA, "svb_clip_join: %s", cudaGetErrorString(e));
        }
    }
    *cands = out, *n_cands = h.n_cands;
    return 0;
}
