// Walker laboratory: times candidate forms of the full-pass record walkers on a synthetic packed-record stream that has the
// shape of the C2 workload (records of ~295 bytes: 36-byte fixed part, name, CIGAR, bases, qualities, aux), without any
// product code. It answers, on the B200 itself, the questions DESIGN.md section 8 left open:
//   * what does the chunk size do (16 KiB chunks put the 32 chains of a warp at the same offset modulo 16 KiB)?
//   * how much do fewer, wider loads per record buy (256-bit loads)?
//   * what does a warp-cooperative form reach that stages each chain's next 128-byte line(s) in shared memory with coalesced
//     16-byte cp.async pieces and parses the record head from there (one or two lines per record instead of 7-9 scattered
//     load instructions of 32 wavefronts each)?
//   * dense rows (need a per-chunk base, i.e. a counting pass first) against per-chunk row slots.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/walk_lab tools/walk_lab.cu && ./tools/walk_lab [GB=2.7]
//
// Every variant reports a checksum of what it computed so that the forms can be compared with each other.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define CK(x)                                                                            \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) {                                                         \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                     \
        }                                                                                \
    } while (0)

static constexpr uint64_t BAD = ~0ull;

struct Row {  // 48 bytes
    int32_t tid, pos, end;
    uint32_t flagq;
    int32_t lqseq, mtid, mpos, isize;
    uint64_t off, pad;
};

struct Out {
    uint32_t *count;        // per chunk
    uint64_t *exit_;        // per chunk
    uint32_t *counters;     // [0] clipped [1] unmapped
    uint64_t *clipped;      // queue
    uint64_t *unmapped;     // queue
    uint32_t cap;
    Row *rows;              // dense (base[]) or sparse (c * R + k)
    const uint64_t *base;   // dense row base per chunk (nullptr: sparse)
    uint32_t R;
    unsigned long long *sum;  // checksum
};

// ---- unaligned access helpers (global) ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldu32(const uint8_t *p)
{
    uintptr_t a = (uintptr_t)p;
    const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
    uint32_t sh = ((uint32_t)a & 3u) * 8u;
    uint32_t lo = __ldg(w);
    if (sh == 0) return lo;
    return __funnelshift_r(lo, __ldg(w + 1), sh);
}

struct Core {
    int32_t block_size, tid, pos, l_qseq, mtid, mpos, isize;
    uint32_t l_qname, mapq, n_cigar, flag;
};
__device__ __forceinline__ Core core_from_words(const uint32_t *f)
{
    Core c;
    c.block_size = (int32_t)f[0], c.tid = (int32_t)f[1], c.pos = (int32_t)f[2];
    c.l_qname = f[3] & 0xff, c.mapq = (f[3] >> 8) & 0xff, c.n_cigar = f[4] & 0xffff, c.flag = f[4] >> 16;
    c.l_qseq = (int32_t)f[5], c.mtid = (int32_t)f[6], c.mpos = (int32_t)f[7], c.isize = (int32_t)f[8];
    return c;
}
template <int WI, int NW>
__device__ __forceinline__ void shift_words(const uint32_t (&W)[NW], uint32_t sh, uint32_t (&f)[9])
{
#pragma unroll
    for (int i = 0; i < 9; ++i) f[i] = sh ? __funnelshift_r(W[WI + i], W[WI + i + 1], sh) : W[WI + i];
}
// V = 0: three or four 16-byte loads (the round-1 form)
__device__ __forceinline__ Core load_core16(const uint8_t *p)
{
    uintptr_t a = (uintptr_t)p;
    const uint4 *q = (const uint4 *)(a & ~(uintptr_t)15);
    uint32_t in16 = (uint32_t)a & 15u, sh = ((uint32_t)a & 3u) * 8u;
    uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = make_uint4(0, 0, 0, 0);
    if (in16 + 40 > 48) v3 = __ldg(q + 3);
    uint32_t W[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
    uint32_t f[9];
    switch (in16 >> 2) {
    case 0: shift_words<0>(W, sh, f); break;
    case 1: shift_words<1>(W, sh, f); break;
    case 2: shift_words<2>(W, sh, f); break;
    default: shift_words<3>(W, sh, f); break;
    }
    return core_from_words(f);
}
// V = 1: two (rarely three) 32-byte loads
__device__ __forceinline__ void ldg256(const void *p, uint32_t (&w)[8])
{
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}
__device__ __forceinline__ Core load_core32(const uint8_t *p)
{
    uintptr_t a = (uintptr_t)p;
    const uint8_t *q = (const uint8_t *)(a & ~(uintptr_t)31);
    uint32_t in32 = (uint32_t)a & 31u, sh = ((uint32_t)a & 3u) * 8u;
    uint32_t A[8], B[8], C[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ldg256(q, A);
    ldg256(q + 32, B);
    if (in32 + 40 > 64) ldg256(q + 64, C);
    uint32_t W[24];
#pragma unroll
    for (int i = 0; i < 8; ++i) W[i] = A[i], W[8 + i] = B[i], W[16 + i] = C[i];
    uint32_t f[9];
    switch (in32 >> 2) {
    case 0: shift_words<0>(W, sh, f); break;
    case 1: shift_words<1>(W, sh, f); break;
    case 2: shift_words<2>(W, sh, f); break;
    case 3: shift_words<3>(W, sh, f); break;
    case 4: shift_words<4>(W, sh, f); break;
    case 5: shift_words<5>(W, sh, f); break;
    case 6: shift_words<6>(W, sh, f); break;
    default: shift_words<7>(W, sh, f); break;
    }
    return core_from_words(f);
}

// ---- the record work (same in every form) ----------------------------------------------------------------------------------
struct Acc {
    uint32_t cnt = 0;
    unsigned long long sum = 0;
};
template <bool CLIP, bool ROWS>
__device__ __forceinline__ void record_work(const Core &k, uint64_t o, uint32_t first_op, uint32_t last_op, int32_t rend, uint32_t hard, uint64_t c,
                                            Acc &acc, const Out &out)
{
    if (CLIP) {
        if (k.flag & 12u) {
            uint32_t s = atomicAdd(&out.counters[1], 1u);
            if (s < out.cap) out.unmapped[s] = o;
        } else if (k.n_cigar && k.mapq >= 1 && !(k.flag & 1024u)) {
            uint32_t op1 = first_op & 15, op2 = last_op & 15;
            if (op1 != 5 && op2 != 5 && (op1 == 4 || op2 == 4)) {
                uint32_t s = atomicAdd(&out.counters[0], 1u);
                if (s < out.cap) out.clipped[s] = o;
            }
        }
    }
    if (ROWS) {
        uint32_t fq = k.flag | (k.mapq << 16) | (hard << 24);
        uint64_t slot = out.base ? out.base[c] + acc.cnt : c * out.R + acc.cnt;
        if (out.base || acc.cnt < out.R) {
            uint4 *row = (uint4 *)&out.rows[slot];
            row[0] = make_uint4((uint32_t)k.tid, (uint32_t)k.pos, (uint32_t)rend, fq);
            row[1] = make_uint4((uint32_t)k.l_qseq, (uint32_t)k.mtid, (uint32_t)k.mpos, (uint32_t)k.isize);
            row[2] = make_uint4((uint32_t)o, (uint32_t)(o >> 32), 0u, 0u);
        }
        acc.sum += (uint32_t)rend + fq;
    }
    acc.sum += (uint32_t)k.pos;
    ++acc.cnt;
}

// ---- form A: one thread per chunk (round-1 form; V selects the load width) --------------------------------------------------
template <int V, bool CLIP, bool ROWS>
__global__ void __launch_bounds__(128) walk_thread(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t chunk_bytes,
                                                   const uint64_t *__restrict__ guess, Out out)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t o = guess[c], end = min(n, (c + 1) * (uint64_t)chunk_bytes);
    Acc acc;
    bool live = o < end && o + 36 <= n;
    Core k;
    if (live) k = V ? load_core32(d + o) : load_core16(d + o);
    while (live) {
        if (k.block_size < 32 || o + 4 + (uint64_t)k.block_size > n) break;
        uint64_t on = o + 4 + (uint64_t)k.block_size;
        bool next_live = on < end && on + 36 <= n;
        Core kn;
        if (next_live) kn = V ? load_core32(d + on) : load_core16(d + on);
        const uint8_t *cig = d + o + 36 + k.l_qname;
        uint32_t first_op = 0, last_op = 0, hard = 0;
        int32_t rend = k.pos;
        if (ROWS) {
            for (uint32_t j = 0; j < k.n_cigar; ++j) {
                uint32_t w = ldu32(cig + 4 * j), op = w & 15;
                if (op == 0 || op == 2 || op == 3) rend += (int32_t)(w >> 4);
                if ((j == 0 || j + 1 == k.n_cigar) && op == 5) hard = 1;
                if (j == 0) first_op = w;
                last_op = w;
            }
        } else if (k.n_cigar) {
            first_op = ldu32(cig), last_op = ldu32(cig + 4 * (k.n_cigar - 1));
        }
        record_work<CLIP, ROWS>(k, o, first_op, last_op, rend, hard, c, acc, out);
        o = on, k = kn, live = next_live;
    }
    out.count[c] = acc.cnt, out.exit_[c] = o;
    for (int s = 16; s > 0; s >>= 1) acc.sum += __shfl_xor_sync(0xffffffffu, acc.sum, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out.sum, acc.sum);
}

// plain chase (walk_count): 4 bytes per record
__global__ void __launch_bounds__(128) walk_count(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t chunk_bytes,
                                                  const uint64_t *__restrict__ guess, Out out)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t o = guess[c], end = min(n, (c + 1) * (uint64_t)chunk_bytes);
    uint32_t k = 0;
    while (o < end && o + 4 <= n) {
        int32_t bs = (int32_t)ldu32(d + o);
        if (bs < 32 || o + 4 + (uint64_t)bs > n) break;
        ++k, o += 4 + (uint64_t)bs;
    }
    out.count[c] = k, out.exit_[c] = o;
}

// ---- form S: warp-cooperative staging ----------------------------------------------------------------------------------------
// A warp owns 32 chains (one per lane). Per step every live lane names the 128-byte line that holds the start of its next record
// (and, when the head is expected to cross into the next line, that one too); the warp fetches the named lines into shared memory
// with 16-byte cp.async pieces - sixteen consecutive lanes cover one chain's 256 bytes, so an instruction touches four lines instead
// of thirty-two - and each lane then parses its record head out of its own shared-memory row. Lanes whose chain has ended take the
// next chunk from a ticket counter, so the warps stay full until the stream is exhausted.
constexpr int ROW_BYTES = 272;  // 256 + 16: rows stay 16-byte aligned, consecutive rows start 4 banks apart
constexpr int S_WARPS = 8;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds32(const uint8_t *row, uint32_t off)
{
    const uint32_t *w = (const uint32_t *)row + (off >> 2);
    uint32_t sh = (off & 3u) * 8u;
    return sh ? __funnelshift_r(w[0], w[1], sh) : w[0];
}

// CLIPM: 0 none, 1 one global atomic per queued record, 2 queued records are collected in a per-warp shared-memory buffer and
//        flushed 32 at a time (one global atomic per 32 records)
// ROWM : 0 none, 48 = three 16-byte stores, 64 = padded rows written as two 32-byte stores, 32 = compact rows (one 32-byte store)
// FETCH: 0 whole 128-byte lines (cp.async 16-byte pieces), 1 only the 32-byte sectors that hold the expected head, 2 one bulk copy
//        (TMA, cp.async.bulk) per chain, completion through one mbarrier per warp
constexpr int WQ = 64;  // per-warp queue entries
struct WarpQ {
    uint64_t e[2][WQ];
    uint32_t n[2];
    uint32_t pad[2];
    unsigned long long mbar;
    unsigned long long pad2;
};
__device__ __forceinline__ void stg256(void *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f, uint32_t g, uint32_t h)
{
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e), "r"(f), "r"(g), "r"(h) : "memory");
}
__device__ __forceinline__ void wq_push(WarpQ &q, int which, bool has, uint64_t o, uint32_t lane, const Out &out)
{
    const uint32_t m = __ballot_sync(0xffffffffu, has);
    if (!m) return;
    const uint32_t n0 = q.n[which];
    if (has) q.e[which][n0 + __popc(m & ((1u << lane) - 1))] = o;
    __syncwarp();
    uint32_t n1 = n0 + __popc(m);
    if (n1 >= 32) {  // flush the oldest 32
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(&out.counters[which], 32u);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        uint64_t *dst = which == 0 ? out.clipped : out.unmapped;
        if (b0 + lane < out.cap) dst[b0 + lane] = q.e[which][lane];
        uint64_t keep = lane + 32 < n1 ? q.e[which][lane + 32] : 0;
        __syncwarp();
        if (lane + 32 < n1) q.e[which][lane] = keep;
        n1 -= 32;
    }
    __syncwarp();
    if (lane == 0) q.n[which] = n1;
    __syncwarp();
}
__device__ __forceinline__ void wq_flush(WarpQ &q, int which, uint32_t lane, const Out &out)
{
    const uint32_t n = q.n[which];
    if (!n) return;
    uint32_t b0 = 0;
    if (lane == 0) b0 = atomicAdd(&out.counters[which], n);
    b0 = __shfl_sync(0xffffffffu, b0, 0);
    uint64_t *dst = which == 0 ? out.clipped : out.unmapped;
    for (uint32_t i = lane; i < n; i += 32)
        if (b0 + i < out.cap) dst[b0 + i] = q.e[which][i];
}

template <int CLIPM, int ROWM, int FETCH>
__global__ void __launch_bounds__(S_WARPS * 32) walk_staged(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t chunk_bytes,
                                                           const uint64_t *__restrict__ guess, Out out, unsigned long long *ticket)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *S = smem + (size_t)warp * 32 * ROW_BYTES;
    uint8_t *my = S + lane * ROW_BYTES;
    WarpQ &wq = *(WarpQ *)(smem + (size_t)S_WARPS * 32 * ROW_BYTES + warp * sizeof(WarpQ));
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&wq.mbar);
    if (lane == 0) {
        wq.n[0] = wq.n[1] = 0;
        if (FETCH == 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    }
    if (FETCH == 2) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t phase = 0;
    const uintptr_t base_line = (uintptr_t)d >> 7;  // lines are counted from the line that holds d[0]
    const uint32_t d_in = (uint32_t)((uintptr_t)d & 127);
    uint64_t o = 0, end = 0, c = 0;
    bool live = false, done = false;
    uint32_t head_pred = 100;
    Acc acc;
    unsigned long long total = 0;
    for (;;) {
        uint32_t need = __ballot_sync(0xffffffffu, !live && !done);
        if (need) {
            unsigned long long b0 = 0;
            int leader = __ffs(need) - 1;
            if ((int)lane == leader) b0 = atomicAdd(ticket, (unsigned long long)__popc(need));
            b0 = __shfl_sync(0xffffffffu, b0, leader);
            if (!live && !done) {
                c = b0 + __popc(need & ((1u << lane) - 1));
                if (c < n_chunks) {
                    o = guess[c], end = min(n, (c + 1) * (uint64_t)chunk_bytes);
                    live = o < end && o + 36 <= n;
                    acc.cnt = 0;
                    if (!live) out.count[c] = 0, out.exit_[c] = o;
                } else
                    done = true;
            }
        }
        if (__all_sync(0xffffffffu, !live)) {
            if (__all_sync(0xffffffffu, done)) break;
            continue;
        }
        const uint64_t ao = o + d_in;  // offset from the start of line base_line
        const uint32_t in_line = (uint32_t)ao & 127u;
        const bool two = in_line + head_pred > 128;
        if (FETCH == 2) {
            const uint32_t bytes = live ? (two ? 256u : 128u) : 0u;
            uint32_t tot = bytes;
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(tot) : "memory");
            __syncwarp();
            if (live) {
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(my);
                const uint8_t *src = (const uint8_t *)((base_line + (ao >> 7)) << 7);
                asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                             "r"(mbar)
                             : "memory");
            }
            uint32_t okw = 0;
            while (!okw) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(okw)
                             : "r"(mbar), "r"(phase)
                             : "memory");
            }
            phase ^= 1;
        } else {
            // descriptor: line index << 9 | first 16-byte piece << 5 | last piece << 1 | live  (FETCH 0: pieces 0..7 / 0..15)
            uint32_t p0 = 0, p1 = two ? 15u : 7u;
            if (FETCH == 1) p0 = (in_line >> 5) << 1, p1 = min(((in_line + head_pred - 1) >> 5) << 1 | 1u, 15u);
            const unsigned long long desc = live ? ((unsigned long long)(ao >> 7) << 9 | p0 << 5 | p1 << 1 | 1u) : 0ull;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t src = 2 * j + (lane >> 4), q = lane & 15;
                const unsigned long long ds = __shfl_sync(0xffffffffu, desc, src);
                const uint32_t lo = (uint32_t)ds;
                if ((lo & 1u) && q >= ((lo >> 5) & 15u) && q <= ((lo >> 1) & 15u))
                    cp_async16(S + src * ROW_BYTES + q * 16, (const uint8_t *)((base_line + (ds >> 9)) << 7) + q * 16);
            }
            cp_async_wait_all();
        }
        __syncwarp();
        bool q_clip = false, q_unm = false;
        const uint64_t o_rec = o;
        if (live) {
            uint32_t avail = two ? 256u : 128u;
            if (FETCH == 1) avail = min((((in_line + head_pred - 1) >> 5) + 1) << 5, 256u);
            uint32_t f[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) f[i] = lds32(my, in_line + 4 * i);
            Core k = core_from_words(f);
            bool ok = k.block_size >= 32 && o + 4 + (uint64_t)k.block_size <= n;
            if (ok) {
                const uint32_t cig = in_line + 36 + k.l_qname, head_end = cig + 4 * k.n_cigar;
                uint32_t first_op = 0, last_op = 0, hard = 0;
                int32_t rend = k.pos;
                const bool in_smem = head_end <= avail;
                const uint8_t *g = d + o + 36 + k.l_qname;
                if (ROWM) {
                    for (uint32_t j = 0; j < k.n_cigar; ++j) {
                        uint32_t w = in_smem ? lds32(my, cig + 4 * j) : ldu32(g + 4 * j), op = w & 15;
                        if (op == 0 || op == 2 || op == 3) rend += (int32_t)(w >> 4);
                        if ((j == 0 || j + 1 == k.n_cigar) && op == 5) hard = 1;
                        if (j == 0) first_op = w;
                        last_op = w;
                    }
                } else if (k.n_cigar) {
                    first_op = in_smem ? lds32(my, cig) : ldu32(g);
                    last_op = in_smem ? lds32(my, head_end - 4) : ldu32(g + 4 * (k.n_cigar - 1));
                }
                if (CLIPM) {
                    if (k.flag & 12u) q_unm = true;
                    else if (k.n_cigar && k.mapq >= 1 && !(k.flag & 1024u)) {
                        uint32_t op1 = first_op & 15, op2 = last_op & 15;
                        q_clip = op1 != 5 && op2 != 5 && (op1 == 4 || op2 == 4);
                    }
                    if (CLIPM == 1) {
                        if (q_unm) {
                            uint32_t s = atomicAdd(&out.counters[1], 1u);
                            if (s < out.cap) out.unmapped[s] = o;
                        }
                        if (q_clip) {
                            uint32_t s = atomicAdd(&out.counters[0], 1u);
                            if (s < out.cap) out.clipped[s] = o;
                        }
                    }
                }
                if (ROWM) {
                    uint32_t fq = k.flag | (k.mapq << 16) | (hard << 24);
                    if (out.base || acc.cnt < out.R) {
                        if (ROWM == 48) {
                            uint64_t slot = out.base ? out.base[c] + acc.cnt : c * out.R + acc.cnt;
                            uint4 *row = (uint4 *)&out.rows[slot];
                            row[0] = make_uint4((uint32_t)k.tid, (uint32_t)k.pos, (uint32_t)rend, fq);
                            row[1] = make_uint4((uint32_t)k.l_qseq, (uint32_t)k.mtid, (uint32_t)k.mpos, (uint32_t)k.isize);
                            row[2] = make_uint4((uint32_t)o, (uint32_t)(o >> 32), 0u, 0u);
                        } else if (ROWM == 64) {
                            uint8_t *row = (uint8_t *)out.rows + (c * out.R + acc.cnt) * 64;
                            stg256(row, (uint32_t)k.tid, (uint32_t)k.pos, (uint32_t)rend, fq, (uint32_t)k.l_qseq, (uint32_t)k.mtid, (uint32_t)k.mpos,
                                   (uint32_t)k.isize);
                            stg256(row + 32, (uint32_t)o, (uint32_t)(o >> 32), 0u, 0u, 0u, 0u, 0u, 0u);
                        } else {
                            uint8_t *row = (uint8_t *)out.rows + (c * out.R + acc.cnt) * 32;
                            stg256(row, (uint32_t)k.tid | (uint32_t)k.mtid << 16, (uint32_t)k.pos, (uint32_t)rend, fq, (uint32_t)k.l_qseq, (uint32_t)o,
                                   (uint32_t)k.mpos, (uint32_t)k.isize);
                        }
                    }
                    acc.sum += (uint32_t)rend + fq;
                }
                acc.sum += (uint32_t)k.pos;
                ++acc.cnt;
                head_pred = min(36u + k.l_qname + 4 * k.n_cigar + 8u, 129u);
                o += 4 + (uint64_t)k.block_size;
            }
            if (!ok || !(o < end && o + 36 <= n)) {
                live = false;
                out.count[c] = acc.cnt, out.exit_[c] = ok ? o : (k.block_size < 32 ? BAD : o);
                total += acc.sum, acc.sum = 0;
            }
        }
        if (CLIPM == 2) {
            wq_push(wq, 0, q_clip, o_rec, lane, out);
            wq_push(wq, 1, q_unm, o_rec, lane, out);
        }
        __syncwarp();
    }
    if (CLIPM == 2) {
        wq_flush(wq, 0, lane, out);
        wq_flush(wq, 1, lane, out);
    }
    total += acc.sum;
    for (int s = 16; s > 0; s >>= 1) total += __shfl_xor_sync(0xffffffffu, total, s);
    if (lane == 0) atomicAdd(out.sum, total);
}

// ---- host ---------------------------------------------------------------------------------------------------------------------
int main(int argc, char **argv)
{
    double gb = argc > 1 ? atof(argv[1]) : 2.7;
    const uint64_t target = (uint64_t)(gb * 1e9);
    std::vector<uint8_t> h(target + (1 << 20), 0);
    std::vector<uint64_t> offs;
    offs.reserve(target / 280);
    std::mt19937_64 rng(20261017);
    uint64_t o = 341;  // (a header in front, like a BAM)
    int32_t pos = 0;
    const uint64_t first = o;
    while (o + 600 < target) {
        uint32_t lq = 18 + rng() % 9, r = rng() % 100, nc = r < 90 ? 1 : r < 98 ? 2 + rng() % 2 : 4 + rng() % 3, l = 150, aux = 40 + rng() % 41;
        uint32_t bs = 32 + lq + 4 * nc + (l + 1) / 2 + l + aux;
        uint32_t flag = (rng() % 100 < 2 ? 8u : 0u) | (rng() % 100 < 1 ? 1024u : 0u) | 1u | (rng() & 1 ? 16u : 32u) | 2u;
        uint32_t w[9] = {bs, 0u, (uint32_t)pos, lq | 60u << 8 | 4681u << 16, nc | flag << 16, l, 0u, (uint32_t)(pos + 350), 500u};
        memcpy(&h[o], w, 36);
        for (uint32_t i = 0; i + 1 < lq; ++i) h[o + 36 + i] = 'A' + (i % 26);
        uint32_t left = l;
        for (uint32_t j = 0; j < nc; ++j) {
            uint32_t op = 0, len = left / (nc - j);
            if (j == 0 && rng() % 100 < 1) op = 4, len = 20;
            else if (j + 1 == nc && nc > 1 && rng() % 100 < 30) op = 4;
            else if (j > 0 && j + 1 < nc) op = (j & 1) ? 2 : 0;
            left -= (op == 2) ? 0 : len;
            uint32_t cw = len << 4 | op;
            memcpy(&h[o + 36 + lq + 4 * j], &cw, 4);
        }
        offs.push_back(o);
        o += 4 + bs;
        pos += 5;
    }
    const uint64_t n = o, n_rec = offs.size();
    fprintf(stderr, "stream %.3f GB, %llu records, %.1f bytes/record\n", n / 1e9, (unsigned long long)n_rec, (double)(n - first) / n_rec);
    uint8_t *d;
    CK(cudaMalloc(&d, n + 4096));
    CK(cudaMemset(d + n, 0, 4096));
    CK(cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice));
    h.clear();
    h.shrink_to_fit();
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    cudaEvent_t ea, eb;
    CK(cudaEventCreate(&ea));
    CK(cudaEventCreate(&eb));
    const uint32_t cap = 1u << 22, R = 160;
    uint32_t *d_counters;
    uint64_t *d_clipped, *d_unm;
    unsigned long long *d_sum, *d_ticket;
    CK(cudaMalloc(&d_counters, 16));
    CK(cudaMalloc(&d_clipped, cap * 8ull));
    CK(cudaMalloc(&d_unm, cap * 8ull));
    CK(cudaMalloc(&d_sum, 8));
    CK(cudaMalloc(&d_ticket, 8));
    Row *d_rows_dense, *d_rows_sparse;
    CK(cudaMalloc(&d_rows_dense, (n_rec + 16) * sizeof(Row)));
    const uint32_t chunk_sizes[] = {16384, 32768};
    uint64_t max_chunks = n / 8192 + 2;
    CK(cudaMalloc(&d_rows_sparse, (n / 16384 + 2) * (uint64_t)R * 64 * 2 + (1 << 20)));
    uint32_t *d_count;
    uint64_t *d_exit, *d_guess, *d_base;
    CK(cudaMalloc(&d_count, max_chunks * 4));
    CK(cudaMalloc(&d_exit, max_chunks * 8));
    CK(cudaMalloc(&d_guess, max_chunks * 8));
    CK(cudaMalloc(&d_base, max_chunks * 8));

    for (uint32_t cb : chunk_sizes) {
        const uint64_t n_chunks = (n + cb - 1) / cb;
        std::vector<uint64_t> guess(n_chunks), base(n_chunks);
        for (uint64_t c = 0; c < n_chunks; ++c) {
            uint64_t start = std::max<uint64_t>(c * (uint64_t)cb, first);
            auto it = std::lower_bound(offs.begin(), offs.end(), start);
            guess[c] = it == offs.end() ? n : *it;
            base[c] = (uint64_t)(it - offs.begin());
        }
        CK(cudaMemcpy(d_guess, guess.data(), n_chunks * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_base, base.data(), n_chunks * 8, cudaMemcpyHostToDevice));
        const uint32_t Rc = (uint32_t)((uint64_t)R * cb / 16384);
        auto run = [&](const char *name, int rows_mode /*0 none 1 dense 2 sparse*/, auto launch) {
            Out out{d_count, d_exit, d_counters, d_clipped, d_unm, cap, rows_mode == 1 ? d_rows_dense : d_rows_sparse,
                    rows_mode == 1 ? d_base : nullptr, Rc, d_sum};
            float best = 1e30f;
            unsigned long long hs = 0;
            uint32_t hc[4];
            for (int rep = 0; rep < 6; ++rep) {
                CK(cudaMemset(d_counters, 0, 16));
                CK(cudaMemset(d_sum, 0, 8));
                CK(cudaMemset(d_ticket, 0, 8));
                CK(cudaEventRecord(ea));
                launch(out);
                CK(cudaPeekAtLastError());
                CK(cudaEventRecord(eb));
                CK(cudaEventSynchronize(eb));
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, ea, eb));
                if (rep >= 2) best = std::min(best, ms);
            }
            CK(cudaMemcpy(&hs, d_sum, 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(hc, d_counters, 16, cudaMemcpyDeviceToHost));
            std::vector<uint32_t> cnt(n_chunks);
            CK(cudaMemcpy(cnt.data(), d_count, n_chunks * 4, cudaMemcpyDeviceToHost));
            uint64_t tot = 0;
            for (uint32_t x : cnt) tot += x;
            printf("{\"variant\": \"%s\", \"chunk\": %u, \"ms\": %.4f, \"stream_GBps\": %.0f, \"records\": %llu, \"clipped\": %u, \"unmapped\": %u, \"sum\": %llu}\n",
                   name, cb, best, n / (best * 1e-3) / 1e9, (unsigned long long)tot, hc[0], hc[1], hs);
            fflush(stdout);
        };
        const unsigned grid_t = (unsigned)((n_chunks + 127) / 128);
        run("count_thread", 0, [&](Out out) { walk_count<<<grid_t, 128>>>(d, n, n_chunks, cb, d_guess, out); });
        if (cb == 16384) run("clip_thread16", 0, [&](Out out) { walk_thread<0, true, false><<<grid_t, 128>>>(d, n, n_chunks, cb, d_guess, out); });
        if (cb == 16384) run("clip_thread32", 0, [&](Out out) { walk_thread<1, true, false><<<grid_t, 128>>>(d, n, n_chunks, cb, d_guess, out); });
        if (cb == 16384) {
            run("rows_dense_thread16", 1, [&](Out out) { walk_thread<0, false, true><<<grid_t, 128>>>(d, n, n_chunks, cb, d_guess, out); });
            run("rows_dense_thread32", 1, [&](Out out) { walk_thread<1, false, true><<<grid_t, 128>>>(d, n, n_chunks, cb, d_guess, out); });
            run("rows_sparse_thread32", 2, [&](Out out) { walk_thread<1, false, true><<<grid_t, 128>>>(d, n, n_chunks, cb, d_guess, out); });
            run("fused_sparse_thread32", 2, [&](Out out) { walk_thread<1, true, true><<<grid_t, 128>>>(d, n, n_chunks, cb, d_guess, out); });
        }
        const size_t shm = (size_t)S_WARPS * 32 * ROW_BYTES + S_WARPS * sizeof(WarpQ);
#define RUN_S(NAME, ROWSMODE, ...)                                                                                               \
    do {                                                                                                                         \
        CK(cudaFuncSetAttribute(walk_staged<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));               \
        char nm[64];                                                                                                             \
        snprintf(nm, sizeof nm, "%s_x%d", NAME, per_sm);                                                                         \
        run(nm, ROWSMODE, [&](Out out) {                                                                                         \
            walk_staged<__VA_ARGS__><<<grid_s, S_WARPS * 32, shm>>>(d, n, n_chunks, cb, d_guess, out, d_ticket);              \
        });                                                                                                                      \
    } while (0)
        for (int per_sm = 2; per_sm <= 3; ++per_sm) {
            const unsigned grid_s = (unsigned)(sms * per_sm);
            if (cb != 16384 && per_sm != 3) continue;
            RUN_S("count_staged", 0, 0, 0, 0);
            RUN_S("count_staged_precise", 0, 0, 0, 1);
            RUN_S("count_staged_bulk", 0, 0, 0, 2);
            RUN_S("clip_staged_atom", 0, 1, 0, 0);
            RUN_S("clip_staged_wq", 0, 2, 0, 0);
            if (cb == 16384) {
                RUN_S("clip_staged_wq_bulk", 0, 2, 0, 2);
                RUN_S("rows48_sparse_staged", 2, 0, 48, 0);
                RUN_S("rows48_dense_staged", 1, 0, 48, 0);
                RUN_S("rows64_sparse_staged", 2, 0, 64, 0);
                RUN_S("rows32_sparse_staged", 2, 0, 32, 0);
                RUN_S("rows32_sparse_staged_bulk", 2, 0, 32, 2);
                RUN_S("fused48_sparse_staged_wq", 2, 2, 48, 0);
                RUN_S("fused32_sparse_staged_wq", 2, 2, 32, 0);
                RUN_S("fused32_sparse_staged_wq_bulk", 2, 2, 32, 2);
            }
        }
    }
    return 0;
}
