// Device-side building blocks of the sync-free pipelines (getclip.cu, getsv.cu): every element count lives in device memory, so
// that a whole command is ONE stream-ordered sequence of launches and the host reads a single small control block at the end.
//
//   * Bump        - carves one context-owned workspace buffer into typed arrays (no cudaMallocAsync per array)
//   * chained_scan - single-pass exclusive prefix sum (decoupled look-back over tile aggregates), element count read on the device;
//                    the operand is a functor, so the producers of the summed values (segment flags, text sizes, ...) are fused in
//   * radix_sort   - stable LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits, ONE persistent kernel per sort with
//                    grid-wide barriers between the passes, element count read on the device; passes whose digit is the same
//                    in every key are skipped. Replaces the cub::DeviceRadixSort calls of round 1 (SURVEY.md section 7 allowed a
//                    library sort "unless ncu says it matters" - it did: many tiny launches and host-side size queries).
#pragma once
#include "common.cuh"

#ifdef __CUDACC__

// ---- workspace carving -----------------------------------------------------------------------------------------------------
struct Bump {
    uint8_t *base;  // nullptr: measuring pass
    size_t used = 0;
    explicit Bump(uint8_t *b) : base(b) {}
    template <typename T>
    T *get(size_t count)
    {
        used = (used + 255) & ~(size_t)255;
        T *p = base ? (T *)(base + used) : nullptr;
        used += (count ? count : 1) * sizeof(T);
        return p;
    }
};

static inline unsigned grid_for(const svb_ctx *ctx, uint64_t cap_items, unsigned items_per_cta, unsigned ctas_per_sm)
{
    uint64_t need = (cap_items + items_per_cta - 1) / items_per_cta;
    uint64_t most = (uint64_t)ctx->sm_count * ctas_per_sm;
    return (unsigned)std::max<uint64_t>(1, std::min(need, most));
}

// ---- chained scan -----------------------------------------------------------------------------------------------------------
// State word of a tile: bits 63-62 = 0 nothing yet, 1 tile aggregate, 2 inclusive prefix; bits 61-0 = value.
#define SCAN_AGG (1ull << 62)
#define SCAN_INC (2ull << 62)
#define SCAN_VAL ((1ull << 62) - 1)
static constexpr int SCAN_THREADS = 256;  // items per thread are a template parameter: 8 for cheap operands, 2 when an operand costs a DRAM round trip

struct ScanScratch {  // one per scan launch; zeroed before the launch (part of the pipeline's sync area)
    unsigned long long *state;  // NV * tiles_cap
    uint32_t *ticket;
    uint32_t tiles_cap;
};
static inline size_t scan_tiles(uint64_t cap_items, int items) { return (size_t)((cap_items + SCAN_THREADS * items - 1) / (SCAN_THREADS * items) + 1); }

// decoupled look-back of one tile, run by a whole warp: 32 predecessor tiles per round
__device__ __forceinline__ uint64_t scan_look_back(unsigned long long *st, uint64_t tile, uint64_t agg, uint32_t lane)
{
    uint64_t base = 0;
    if (tile > 0) {
        if (lane == 0) atomicExch(&st[tile], SCAN_AGG | agg);
        int64_t p = (int64_t)tile - 1;
        for (;;) {
            const int64_t idx = p - (int64_t)lane;
            const unsigned long long sv = idx >= 0 ? *(volatile unsigned long long *)&st[idx] : SCAN_INC;  // in front of tile 0: prefix 0
            const uint32_t flag = (uint32_t)(sv >> 62);
            const uint32_t inc = __ballot_sync(0xffffffffu, flag == 2), zero = __ballot_sync(0xffffffffu, flag == 0);
            const int first = inc ? __ffs(inc) - 1 : 32;
            const uint32_t need = first < 31 ? (2u << first) - 1u : 0xffffffffu;  // the lanes up to the first inclusive prefix
            if (zero & need) continue;  // a predecessor has not published yet: look again
            uint64_t v = ((int)lane <= first) ? (uint64_t)(sv & SCAN_VAL) : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            base += v;
            if (first < 32) break;
            p -= 32;
        }
    }
    if (lane == 0) atomicExch(&st[tile], SCAN_INC | (base + agg));
    return base;
}

// Op interface (all __device__):
//   uint64_t n() const;                                   element count (from device memory)
//   void load(uint64_t i, uint64_t (&v)[NV]) const;       the values of element i
//   void store(uint64_t i, const uint64_t (&excl)[NV], const uint64_t (&v)[NV]) const;   exclusive prefix of element i
//   void total(const uint64_t (&t)[NV]) const;            called once, by one thread, with the grand totals
template <int NV, int SCAN_ITEMS, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) chained_scan(Op op, ScanScratch sc)
{
    constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
    __shared__ uint64_t warp_sum[NV][SCAN_THREADS / 32];
    __shared__ uint64_t tile_base[NV];
    __shared__ uint32_t s_tile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint64_t n = op.n();
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(sc.ticket, 1u);
        __syncthreads();
        const uint64_t tile = s_tile, first = tile * SCAN_TILE;
        if (first >= n) {
            if (tile == 0 && tid == 0) {
                uint64_t z[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) z[k] = 0;
                op.total(z);
            }
            break;
        }
        uint64_t v[SCAN_ITEMS][NV], run[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) run[k] = 0;
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            const uint64_t i = first + (uint64_t)tid * SCAN_ITEMS + j;
#pragma unroll
            for (int k = 0; k < NV; ++k) v[j][k] = 0;
            if (i < n) op.load(i, v[j]);
#pragma unroll
            for (int k = 0; k < NV; ++k) run[k] += v[j][k];
        }
        // exclusive prefix of the per-thread sums inside the block
        uint64_t excl[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            uint64_t x = run[k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t t = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= (uint32_t)o) x += t;
            }
            if (lane == 31) warp_sum[k][wid] = x;
            excl[k] = x - run[k];
        }
        __syncthreads();
        if (wid == 0) {
            uint64_t agg[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const uint64_t w = lane < SCAN_THREADS / 32 ? warp_sum[k][lane] : 0;
                uint64_t x = w;
#pragma unroll
                for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
                    uint64_t t = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= (uint32_t)o) x += t;
                }
                agg[k] = __shfl_sync(0xffffffffu, x, SCAN_THREADS / 32 - 1);
                __syncwarp();
                if (lane < SCAN_THREADS / 32) warp_sum[k][lane] = x - w;  // exclusive over the warps
            }
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const uint64_t base = scan_look_back(sc.state + (size_t)k * sc.tiles_cap, tile, agg[k], lane);
                if (lane == 0) tile_base[k] = base;
                agg[k] += base;
            }
            if (lane == 0 && first + SCAN_TILE >= n) op.total(agg);
        }
        __syncthreads();
        uint64_t pre[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) pre[k] = tile_base[k] + warp_sum[k][wid] + excl[k];
#pragma unroll
        for (int j = 0; j < SCAN_ITEMS; ++j) {
            const uint64_t i = first + (uint64_t)tid * SCAN_ITEMS + j;
            if (i < n) op.store(i, pre, v[j]);
#pragma unroll
            for (int k = 0; k < NV; ++k) pre[k] += v[j][k];
        }
    }
}

template <int NV, int ITEMS, class Op>
static inline void launch_scan(const svb_ctx *ctx, cudaStream_t s, const Op &op, const ScanScratch &sc, uint64_t cap_items)
{
    chained_scan<NV, ITEMS, Op><<<grid_for(ctx, cap_items, SCAN_THREADS * ITEMS, 8), SCAN_THREADS, 0, s>>>(op, sc);
}

// ---- radix sort ---------------------------------------------------------------------------------------------------------------
// ONE persistent kernel per sort, a CTA per SM (all resident), stable, 8-bit digits:
//   phase 0  digit histograms of ALL passes in one sweep over the keys -> digit bases per pass; a pass in which every key has the
//            same digit moves nothing and is skipped without any work;
//   per pass every tile counts its digits, publishes the counts, adds up what the tiles in front of it published (chained
//            look-back on one word per tile and digit: no grid barrier between counting and placing), ranks its keys in input
//            order and scatters them; one grid-wide barrier ends the pass.
// History (profiles/r2_summary.md): one kernel per pass with ~100 small tiles and a serial look-back cost 16 us per pass and a
// dozen launches per sort (0.41 ms per getclip on C2); count | barrier | place | barrier in one kernel still 9-12 us per pass.
static constexpr int RS_THREADS = 512, RS_WARPS = RS_THREADS / 32, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS, RS_MAX_PASSES = 8;

struct RadixScratch {  // zero at launch
    uint32_t *ghist;   // RS_MAX_PASSES * 256 global digit counts
    uint32_t *state;   // tiles_cap * 256 look-back words: bits 31-28 pass tag, bit 27 inclusive, bits 26-0 count
    uint32_t *bar;     // grid barrier counter
    uint32_t tiles_cap;
};
static inline size_t rs_tiles(uint64_t cap_items) { return (size_t)((cap_items + RS_TILE - 1) / RS_TILE + 1); }

struct RadixJob {
    uint64_t *key[2];
    uint32_t *val[2];      // nullptr: keys only
    const uint32_t *n_ptr;  // element count on the device (clamped to cap; below 2^27)
    uint32_t cap;
    int begin_bit, passes;
    RadixScratch sc;
};

// all CTAs of the grid are resident (grid <= SM count, one CTA fits every SM); epochs count up, the counter never resets
__device__ __forceinline__ void grid_barrier(uint32_t *bar, uint32_t &epoch)
{
    __syncthreads();
    ++epoch;
    if (threadIdx.x == 0) {
        const uint32_t target = epoch * gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        while (*(volatile uint32_t *)bar < target) {}
        __threadfence();
    }
    __syncthreads();
}

template <bool PAIRS>
__global__ void __launch_bounds__(RS_THREADS) rs_sort(RadixJob job)
{
    __shared__ uint32_t wc[RS_WARPS][256];  // phase 0: histograms of all passes; then per-warp digit counts / offsets inside a tile
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t gbase[RS_MAX_PASSES][256];
    __shared__ uint32_t s_skip;
    const uint32_t n = min(min(*job.n_ptr, job.cap), (1u << 27) - 1u), tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t T = (n + RS_TILE - 1) / RS_TILE, G = gridDim.x;
    uint32_t epoch = 0;
    // ---- phase 0: all digit histograms
    for (int i = tid; i < RS_MAX_PASSES * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
    if (tid == 0) s_skip = 0;
    __syncthreads();
    for (uint32_t t = blockIdx.x; t < T; t += G) {
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const uint64_t i = (uint64_t)t * RS_TILE + j * RS_THREADS + tid;
            if (i < n) {
                const uint64_t k = __ldcg(job.key[0] + i) >> job.begin_bit;
                for (int p = 0; p < job.passes; ++p) atomicAdd(&wc[p][(k >> (8 * p)) & 255u], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < job.passes * 256; i += RS_THREADS) {
        const uint32_t c = (&wc[0][0])[i];
        if (c) atomicAdd(&job.sc.ghist[i], c);
    }
    grid_barrier(job.sc.bar, epoch);
    // digit bases of every pass (each CTA computes the same): warp w scans pass w
    if ((int)wid < job.passes) {
        uint32_t carry = 0;
        for (int k0 = 0; k0 < 256; k0 += 32) {
            const uint32_t c = __ldcg(&job.sc.ghist[wid * 256 + k0 + lane]);
            if (c == n) atomicOr(&s_skip, 1u << wid);  // (n == 0: every pass is skipped)
            uint32_t x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= (uint32_t)o) x += y;
            }
            gbase[wid][k0 + lane] = carry + x - c;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
    }
    __syncthreads();
    const uint32_t skip = s_skip;
    int par = 0;
    for (int pass = 0; pass < job.passes; ++pass) {
        if (skip >> pass & 1u) continue;  // the order does not change, the buffers do not flip
        const int shift = job.begin_bit + 8 * pass;
        const uint32_t tag = (uint32_t)(pass + 1) << 28;
        const uint64_t *kin = job.key[par];
        uint64_t *kout = job.key[par ^ 1];
        const uint32_t *vin = PAIRS ? job.val[par] : nullptr;
        uint32_t *vout = PAIRS ? job.val[par ^ 1] : nullptr;
        for (uint32_t t = blockIdx.x; t < T; t += G) {
            for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
            __syncthreads();
            uint64_t key[RS_ITEMS];
            uint32_t rank[RS_ITEMS];
            // warp w owns items [w * 256, (w + 1) * 256) of the tile, 32 consecutive ones per step: ranks follow the input order
#pragma unroll
            for (int j = 0; j < RS_ITEMS; ++j) {
                const uint64_t i = (uint64_t)t * RS_TILE + wid * (32 * RS_ITEMS) + j * 32 + lane;
                key[j] = i < n ? __ldcg(kin + i) : ~0ull;
            }
#pragma unroll
            for (int j = 0; j < RS_ITEMS; ++j) {
                const uint64_t i = (uint64_t)t * RS_TILE + wid * (32 * RS_ITEMS) + j * 32 + lane;
                const bool valid = i < n;
                const uint32_t d = valid ? (uint32_t)(key[j] >> shift) & 255u : 256u;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                uint32_t old = 0;
                if (valid) old = wc[wid][d];
                __syncwarp();
                if (valid && (peers & ((1u << lane) - 1u)) == 0) wc[wid][d] = old + __popc(peers);  // the lowest peer updates
                __syncwarp();
                rank[j] = old + __popc(peers & ((1u << lane) - 1u));
            }
            __syncthreads();
            if (tid < 256) {  // digit `tid`: offsets of the warps inside the tile, then the tile's place among the earlier tiles
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < RS_WARPS; ++w) {
                    const uint32_t x = wc[w][tid];
                    wc[w][tid] = run;
                    run += x;
                }
                uint32_t *st = job.sc.state + tid;
                uint32_t before = 0;
                if (t > 0) {
                    // Look back over the tiles in front, sixteen at a time (independent loads in flight: with one tile per CTA all
                    // tiles publish their counts at about the same moment, and a one-by-one walk was ~0.2 us per tile).
                    atomicExch(&st[(size_t)t * 256], tag | run);
                    int64_t p = (int64_t)t - 1;
                    while (p >= 0) {
                        uint32_t sv[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) sv[q] = p - q >= 0 ? *(volatile uint32_t *)&st[(size_t)(p - q) * 256] : (tag | 1u << 27);
                        int q = 0;
                        bool stop = false;
#pragma unroll
                        for (; q < 16; ++q) {
                            if ((sv[q] >> 28) != (tag >> 28)) break;  // not published yet: come back to this tile
                            before += sv[q] & 0x7ffffffu;
                            if (sv[q] >> 27 & 1u) {
                                stop = true;
                                break;
                            }
                        }
                        if (stop) break;
                        p -= q;
                    }
                }
                atomicExch(&st[(size_t)t * 256], tag | 1u << 27 | (before + run));
                dbase[tid] = gbase[pass][tid] + before;
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < RS_ITEMS; ++j) {
                const uint64_t i = (uint64_t)t * RS_TILE + wid * (32 * RS_ITEMS) + j * 32 + lane;
                if (i < n) {
                    const uint32_t d = (uint32_t)(key[j] >> shift) & 255u;
                    const uint32_t dst = dbase[d] + wc[wid][d] + rank[j];
                    kout[dst] = key[j];
                    if (PAIRS) vout[dst] = __ldcg(vin + i);
                }
            }
            __syncthreads();
        }
        par ^= 1;
        grid_barrier(job.sc.bar, epoch);
    }
    if (par == 0) {  // the sorted sequence sits in buffer 0: bring it to buffer 1 (the job's output side)
        for (uint64_t i = (uint64_t)blockIdx.x * RS_THREADS + tid; i < n; i += (uint64_t)G * RS_THREADS) {
            job.key[1][i] = __ldcg(job.key[0] + i);
            if (PAIRS) job.val[1][i] = __ldcg(job.val[0] + i);
        }
    }
}

// Stable sort of key[0] / val[0] by key bits [begin_bit, begin_bit + 8 * passes); the result is left in key[1] / val[1]
// (key[0] / val[0] are overwritten). The scratch must be zero when the kernel starts.
static inline void radix_sort(const svb_ctx *ctx, cudaStream_t s, const RadixJob &job)
{
    const unsigned g = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(rs_tiles(job.cap), (uint64_t)ctx->sm_count));
    if (job.val[0]) rs_sort<true><<<g, RS_THREADS, 0, s>>>(job);
    else rs_sort<false><<<g, RS_THREADS, 0, s>>>(job);
}
static inline RadixScratch radix_scratch(Bump &b, uint64_t cap_items, int passes)
{
    (void)passes;
    RadixScratch sc;
    sc.tiles_cap = (uint32_t)rs_tiles(cap_items);
    sc.bar = b.get<uint32_t>(1);
    sc.ghist = b.get<uint32_t>(RS_MAX_PASSES * 256);
    sc.state = b.get<uint32_t>((size_t)sc.tiles_cap * 256);
    return sc;
}
static inline ScanScratch scan_scratch(Bump &b, uint64_t cap_items, int nv, int items)
{
    ScanScratch sc;
    sc.tiles_cap = (uint32_t)scan_tiles(cap_items, items);
    sc.ticket = b.get<uint32_t>(1);
    sc.state = b.get<unsigned long long>((size_t)nv * sc.tiles_cap);
    return sc;
}

#endif  // __CUDACC__
