// Device-side building blocks of the sync-free pipelines (getclip.cu, getsv.cu): every element count lives in device memory, so
// that a whole command is ONE stream-ordered sequence of launches and the host reads a single small control block at the end.
//
//   * Bump        - carves one context-owned workspace buffer into typed arrays (no cudaMallocAsync per array)
//   * chained_scan - single-pass exclusive prefix sum (decoupled look-back over tile aggregates), element count read on the device;
//                    the operand is a functor, so the producers of the summed values (segment flags, text sizes, ...) are fused in
//   * radix_sort   - stable LSD radix sort of (u64 key, u32 value) pairs, 8-bit digits, one kernel per digit pass with a chained
//                    per-digit look-back (the "onesweep" scheme), element count read on the device; passes whose digit is the same
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
static constexpr unsigned long long SCAN_AGG = 1ull << 62, SCAN_INC = 2ull << 62, SCAN_VAL = (1ull << 62) - 1;
static constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 4, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct ScanScratch {  // one per scan launch; zeroed before the launch (part of the pipeline's sync area)
    unsigned long long *state;  // NV * tiles_cap
    uint32_t *ticket;
    uint32_t tiles_cap;
};
static inline size_t scan_tiles(uint64_t cap_items) { return (size_t)((cap_items + SCAN_TILE - 1) / SCAN_TILE + 1); }

// Op interface (all __device__):
//   uint64_t n() const;                                   element count (from device memory)
//   void load(uint64_t i, uint64_t (&v)[NV]) const;       the values of element i
//   void store(uint64_t i, const uint64_t (&excl)[NV], const uint64_t (&v)[NV]) const;   exclusive prefix of element i
//   void total(const uint64_t (&t)[NV]) const;            called once, by one thread, with the grand totals
template <int NV, class Op>
__global__ void __launch_bounds__(SCAN_THREADS) chained_scan(Op op, ScanScratch sc)
{
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
        if (tid == 0) {
            uint64_t agg[NV];
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                uint64_t a = 0;
                for (int w = 0; w < SCAN_THREADS / 32; ++w) {
                    uint64_t t = warp_sum[k][w];
                    warp_sum[k][w] = a;
                    a += t;
                }
                agg[k] = a;
            }
            // publish, look back, publish the inclusive prefix
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                unsigned long long *st = sc.state + (size_t)k * sc.tiles_cap;
                uint64_t base = 0;
                if (tile > 0) {
                    atomicExch(&st[tile], SCAN_AGG | agg[k]);
                    uint64_t p = tile - 1;
                    for (;;) {
                        unsigned long long s = *(volatile unsigned long long *)&st[p];
                        if ((s >> 62) == 0) continue;
                        base += s & SCAN_VAL;
                        if ((s >> 62) == 2) break;
                        --p;
                    }
                }
                atomicExch(&st[tile], SCAN_INC | (base + agg[k]));
                tile_base[k] = base;
                agg[k] += base;
            }
            if (first + SCAN_TILE >= n) op.total(agg);
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

template <int NV, class Op>
static inline void launch_scan(const svb_ctx *ctx, cudaStream_t s, const Op &op, const ScanScratch &sc, uint64_t cap_items)
{
    chained_scan<NV, Op><<<grid_for(ctx, cap_items, SCAN_TILE, 4), SCAN_THREADS, 0, s>>>(op, sc);
}

// ---- radix sort ---------------------------------------------------------------------------------------------------------------
static constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS, RS_MAX_PASSES = 8;

struct RadixScratch {  // zeroed before the sort is launched
    uint32_t *hist;      // RS_MAX_PASSES * 256 digit counts, turned into exclusive digit bases by rs_prefix
    uint32_t *skip;      // bit p set: pass p leaves the order unchanged
    uint32_t *state;     // passes * tiles_cap * 256 look-back words (bits 31-30 flag, 29-0 count)
    uint32_t *ticket;    // one per pass
    uint32_t tiles_cap;
};
static inline size_t rs_tiles(uint64_t cap_items) { return (size_t)((cap_items + RS_TILE - 1) / RS_TILE + 1); }

struct RadixJob {
    uint64_t *key[2];
    uint32_t *val[2];      // nullptr: keys only
    const uint32_t *n_ptr;  // element count on the device (clamped to cap)
    uint32_t cap;
    int begin_bit, passes;
    RadixScratch sc;
};
__device__ __forceinline__ uint32_t rs_count(const RadixJob &j) { return min(*j.n_ptr, j.cap); }

static __global__ void __launch_bounds__(RS_THREADS) rs_hist(RadixJob job)
{
    __shared__ uint32_t h[RS_MAX_PASSES][256];
    for (int i = threadIdx.x; i < RS_MAX_PASSES * 256; i += RS_THREADS) (&h[0][0])[i] = 0;
    __syncthreads();
    const uint32_t n = rs_count(job);
    for (uint64_t i = (uint64_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (uint64_t)gridDim.x * RS_THREADS) {
        const uint64_t k = job.key[0][i] >> job.begin_bit;
        for (int p = 0; p < job.passes; ++p) atomicAdd(&h[p][(k >> (8 * p)) & 255], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < job.passes * 256; i += RS_THREADS) {
        const uint32_t c = (&h[0][0])[i];
        if (c) atomicAdd(&job.sc.hist[i], c);
    }
}

// one CTA: digit counts -> exclusive digit bases; a pass in which one digit holds every key is marked as skipped
static __global__ void __launch_bounds__(256) rs_prefix(RadixJob job)
{
    __shared__ uint32_t sh[256];
    const uint32_t n = rs_count(job), d = threadIdx.x;
    for (int p = 0; p < job.passes; ++p) {
        const uint32_t c = job.sc.hist[p * 256 + d];
        sh[d] = c;
        __syncthreads();
        uint32_t base = 0;
        for (uint32_t k = 0; k < d; ++k) base += sh[k];
        job.sc.hist[p * 256 + d] = base;
        if (c == n) atomicOr(job.sc.skip, 1u << p);  // (n == 0: every pass is skipped)
        __syncthreads();
    }
}

template <bool PAIRS>
__global__ void __launch_bounds__(RS_THREADS) rs_pass(RadixJob job, int pass)
{
    __shared__ uint32_t wc[RS_THREADS / 32][256];  // per-warp digit counts, then per-warp digit offsets inside the tile
    __shared__ uint32_t dbase[256];                // where this tile's keys of digit d start in the output
    __shared__ uint32_t s_tile;
    const uint32_t skip = *job.sc.skip;
    if (skip >> pass & 1u) return;
    const int par = __popc(~skip & ((1u << pass) - 1u)) & 1;  // buffers alternate over the passes that run
    const uint64_t *__restrict__ kin = job.key[par];
    uint64_t *__restrict__ kout = job.key[par ^ 1];
    const uint32_t *__restrict__ vin = PAIRS ? job.val[par] : nullptr;
    uint32_t *__restrict__ vout = PAIRS ? job.val[par ^ 1] : nullptr;
    const uint32_t n = rs_count(job), tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int shift = job.begin_bit + 8 * pass;
    uint32_t *state = job.sc.state + (size_t)pass * job.sc.tiles_cap * 256;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(&job.sc.ticket[pass], 1u);
        for (int i = tid; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&wc[0][0])[i] = 0;
        __syncthreads();
        const uint32_t tile = s_tile;
        const uint64_t first = (uint64_t)tile * RS_TILE;
        if (first >= n) break;
        uint64_t key[RS_ITEMS];
        uint32_t rank[RS_ITEMS];
        // warp w owns items [w * 256, (w + 1) * 256) of the tile, 32 consecutive ones per step: ranks follow the input order
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const uint64_t i = first + wid * (32 * RS_ITEMS) + j * 32 + lane;
            const bool valid = i < n;
            key[j] = valid ? kin[i] : ~0ull;
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
        {   // digit `tid`: offsets of the warps inside the tile, then the tile's place among the earlier tiles (look-back)
            const uint32_t d = tid;
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; ++w) {
                const uint32_t t = wc[w][d];
                wc[w][d] = run;
                run += t;
            }
            uint32_t before = 0;
            if (tile > 0) {
                atomicExch(&state[(size_t)tile * 256 + d], 1u << 30 | run);
                uint32_t p = tile - 1;
                for (;;) {
                    const uint32_t s = *(volatile uint32_t *)&state[(size_t)p * 256 + d];
                    if ((s >> 30) == 0) continue;
                    before += s & 0x3fffffffu;
                    if ((s >> 30) == 2) break;
                    --p;
                }
            }
            atomicExch(&state[(size_t)tile * 256 + d], 2u << 30 | (before + run));
            dbase[d] = job.sc.hist[pass * 256 + d] + before;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const uint64_t i = first + wid * (32 * RS_ITEMS) + j * 32 + lane;
            if (i < n) {
                const uint32_t d = (uint32_t)(key[j] >> shift) & 255u;
                const uint32_t dst = dbase[d] + wc[wid][d] + rank[j];
                kout[dst] = key[j];
                if (PAIRS) vout[dst] = vin[i];
            }
        }
    }
}

// the sorted sequence ends in buffer [number of passes that ran] & 1; bring it to buffer 1 (the job's output side)
template <bool PAIRS>
__global__ void rs_finish(RadixJob job)
{
    const uint32_t skip = *job.sc.skip;
    const int par = __popc(~skip & ((1u << job.passes) - 1u)) & 1;
    if (par == 1) return;
    const uint32_t n = rs_count(job);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        job.key[1][i] = job.key[0][i];
        if (PAIRS) job.val[1][i] = job.val[0][i];
    }
}

// Stable sort of key[0] / val[0] by key bits [begin_bit, begin_bit + 8 * passes); the result is left in key[1] / val[1].
// The scratch must be zero when the first kernel runs.
static inline void radix_sort(const svb_ctx *ctx, cudaStream_t s, const RadixJob &job)
{
    const unsigned g = grid_for(ctx, job.cap, RS_TILE, 2);
    rs_hist<<<g, RS_THREADS, 0, s>>>(job);
    rs_prefix<<<1, 256, 0, s>>>(job);
    for (int p = 0; p < job.passes; ++p) {
        if (job.val[0]) rs_pass<true><<<g, RS_THREADS, 0, s>>>(job, p);
        else rs_pass<false><<<g, RS_THREADS, 0, s>>>(job, p);
    }
    if (job.val[0]) rs_finish<true><<<g, 256, 0, s>>>(job);
    else rs_finish<false><<<g, 256, 0, s>>>(job);
}
static inline RadixScratch radix_scratch(Bump &b, uint64_t cap_items, int passes)
{
    RadixScratch sc;
    sc.tiles_cap = (uint32_t)rs_tiles(cap_items);
    sc.hist = b.get<uint32_t>(RS_MAX_PASSES * 256);
    sc.skip = b.get<uint32_t>(1);
    sc.ticket = b.get<uint32_t>(RS_MAX_PASSES);
    sc.state = b.get<uint32_t>((size_t)passes * sc.tiles_cap * 256);
    return sc;
}
static inline ScanScratch scan_scratch(Bump &b, uint64_t cap_items, int nv)
{
    ScanScratch sc;
    sc.tiles_cap = (uint32_t)scan_tiles(cap_items);
    sc.ticket = b.get<uint32_t>(1);
    sc.state = b.get<unsigned long long>((size_t)nv * sc.tiles_cap);
    return sc;
}

#endif  // __CUDACC__
