// Streaming tile pipeline shared by the two full-pass kernels (getclip: clip_stream, getsv: decode_stream).
//
// The packed record stream is cut into 16 KiB tiles. A persistent CTA (grid = a multiple of the SM count) pulls its tiles
// - plus a 1 KiB halo so that the head of a record that starts near the end of a tile is on chip too - into a small ring of
// shared-memory stages with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), so HBM is read in full,
// coalesced lines exactly once per pass. In shared memory warp 0 finds the first record of the tile (32 candidate
// offsets per step, two-record plausibility test), lane 0 walks the chain (~50 hops at shared-memory latency) and the
// whole CTA then parses one record per thread from shared memory. The first-record guess is verified after the kernel
// exactly as for the chunk walkers (exit(t) == guess(t+1), bam_index.cu); bytes beyond the staged window (very long
// read names / CIGARs) are fetched from global memory by the accessor.
#pragma once
#include "common.cuh"

static constexpr uint32_t TILE_LOG2 = 14, TILE = 1u << TILE_LOG2, HALO = 1024, STAGE_BYTES = TILE + HALO + 16;
static constexpr int STAGES = 2, STREAM_THREADS = 128, MAX_TILE_RECS = TILE / 36 + 1;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one staged tile: bytes [0, avail) of the window are in shared memory, anything beyond is read from global memory
struct TileWin {
    const uint8_t *sm;  // 16-byte aligned
    const uint8_t *g;   // global address of the window start
    uint32_t avail;
    __device__ __forceinline__ uint32_t u32(uint32_t off) const
    {
        if (off + 4 <= avail) {
            const uint32_t *p = (const uint32_t *)(sm + (off & ~3u));
            uint32_t sh = (off & 3u) * 8u, lo = p[0];
            if (sh == 0) return lo;
            return __funnelshift_r(lo, p[1], sh);  // (the stage is padded: p[1] is always inside the buffer)
        }
        return ldu32(g + off);
    }
    __device__ __forceinline__ uint8_t u8(uint32_t off) const { return off < avail ? sm[off] : g[off]; }
    __device__ __forceinline__ Core core(uint32_t off) const
    {
        Core c;
        if (off + 40 <= avail) {
            const uint4 *q = (const uint4 *)(sm + (off & ~15u));
            uint32_t in16 = off & 15u, sh = (off & 3u) * 8u;
            uint4 v0 = q[0], v1 = q[1], v2 = q[2], v3 = make_uint4(0, 0, 0, 0);
            if (in16 + 40 > 48) v3 = q[3];
            uint32_t W[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
            uint32_t f[9];
            switch (in16 >> 2) {
            case 0: core_fields<0>(W, sh, f); break;
            case 1: core_fields<1>(W, sh, f); break;
            case 2: core_fields<2>(W, sh, f); break;
            default: core_fields<3>(W, sh, f); break;
            }
            c.block_size = (int32_t)f[0], c.tid = (int32_t)f[1], c.pos = (int32_t)f[2];
            c.l_qname = f[3] & 0xff, c.mapq = (f[3] >> 8) & 0xff;
            c.n_cigar = f[4] & 0xffff, c.flag = f[4] >> 16;
            c.l_qseq = (int32_t)f[5], c.mtid = (int32_t)f[6], c.mpos = (int32_t)f[7], c.isize = (int32_t)f[8];
            return c;
        }
        return load_core(g + off);
    }
};

// the two-record plausibility test of bam_index.cu, on a staged window (off relative to the window, abs = stream offset)
__device__ __forceinline__ bool plausible_win(const TileWin &w, uint64_t win_abs, uint64_t n, uint32_t off, int32_t n_ref, uint32_t *next)
{
    uint64_t o = win_abs + off;
    if (o + 36 > n) return false;
    int32_t bs = (int32_t)w.u32(off);
    if (bs < 33 || o + 4 + (uint64_t)bs > n) return false;
    int32_t tid = (int32_t)w.u32(off + 4);
    if (tid < -1 || tid >= n_ref) return false;
    int32_t pos = (int32_t)w.u32(off + 8);
    if (pos < -1 || pos >= (1 << 29)) return false;
    uint32_t x = w.u32(off + 12), l_qname = x & 0xff;
    if (l_qname < 2) return false;
    uint32_t x2 = w.u32(off + 16), n_cigar = x2 & 0xffff;
    if ((x2 >> 16) & 0xf000) return false;
    int32_t l_qseq = (int32_t)w.u32(off + 20);
    if (l_qseq < 0) return false;
    int32_t mtid = (int32_t)w.u32(off + 24);
    if (mtid < -1 || mtid >= n_ref) return false;
    int32_t mpos = (int32_t)w.u32(off + 28);
    if (mpos < -1 || mpos >= (1 << 29)) return false;
    uint64_t need = 32ull + l_qname + 4ull * n_cigar + ((uint64_t)l_qseq + 1) / 2 + (uint64_t)l_qseq;
    if (need > (uint64_t)bs) return false;
    if ((uint64_t)bs - need > 4ull * (uint64_t)l_qseq + 8192) return false;
    uint8_t c0 = w.u8(off + 36);
    if (c0 < 33 || c0 > 126) return false;
    if (w.u8(off + 36 + l_qname - 1) != 0) return false;
    *next = off + 4 + (uint32_t)bs;
    return true;
}

struct StreamShared {
    alignas(16) uint8_t stage[STAGES][STAGE_BYTES];
    alignas(8) uint64_t full[STAGES];
    uint32_t rec_off[MAX_TILE_RECS];  // record starts of the current tile, relative to the tile start
    uint32_t n_rec;                   // records that start inside the tile
    uint64_t entry, exit_;            // absolute stream offsets
    unsigned long long first_mb, last_mb;  // getclip: (k << 32 | tid) of the first / last mapped-branch record
    uint64_t base;                    // getsv: global index of the tile's first record
    unsigned long long red[4];
};

// Producer side: thread 0 issues the bulk copy of tile `t` into stage `s`.
__device__ __forceinline__ void issue_tile(StreamShared &S, int s, const uint8_t *d, uint64_t padded_bytes, uint64_t t)
{
    uint64_t start = t << TILE_LOG2;
    uint32_t bytes = (uint32_t)min((uint64_t)(TILE + HALO), padded_bytes - start);  // padded_bytes is a multiple of 16
    mbar_expect_tx(&S.full[s], bytes);
    tma_load_1d(S.stage[s], d + start, bytes, &S.full[s]);
}

// Consumer side, steps common to both passes: find the entry (warp 0), walk the chain (lane 0), publish rec_off / n_rec /
// entry / exit in shared memory. Must be called by all threads of the CTA; ends with a __syncthreads().
__device__ __forceinline__ void index_tile(StreamShared &S, const TileWin &w, uint64_t t, uint64_t n, uint64_t first, int32_t n_ref,
                                           const uint8_t *d)
{
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint64_t tile_abs = t << TILE_LOG2;
    if (wid == 0) {
        uint64_t entry;
        if (tile_abs + TILE <= first) entry = first;  // header-only tile
        else if (tile_abs <= first) entry = first;     // the tile that holds the first record: exact
        else {
            entry = BAD_OFFSET;
            uint32_t limit = w.avail > 40 ? w.avail - 40 : 0;
            for (uint32_t base = 0; base < limit; base += 32) {
                uint32_t off = base + lane, nx = 0, nx2 = 0;
                bool ok = off < limit && plausible_win(w, tile_abs, n, off, n_ref, &nx);
                if (ok && tile_abs + nx < n) ok = plausible_win(w, tile_abs, n, nx, n_ref, &nx2);
                uint32_t m = __ballot_sync(0xffffffffu, ok);
                if (m) {
                    entry = tile_abs + base + (__ffs(m) - 1);
                    break;
                }
            }
            if (entry == BAD_OFFSET) {  // no record head in the window (a record longer than a tile): search on in global memory
                uint64_t lim = min(n, tile_abs + (uint64_t)8 * TILE);
                TileWin gw{w.sm, w.g, 0};
                for (uint64_t base = tile_abs + limit; base < lim && entry == BAD_OFFSET; base += 32) {
                    uint32_t off = (uint32_t)(base - tile_abs) + lane, nx = 0, nx2 = 0;
                    bool ok = plausible_win(gw, tile_abs, n, off, n_ref, &nx);
                    if (ok && tile_abs + nx < n) ok = plausible_win(gw, tile_abs, n, nx, n_ref, &nx2);
                    uint32_t m = __ballot_sync(0xffffffffu, ok);
                    if (m) entry = base + (__ffs(m) - 1);
                }
                if (entry == BAD_OFFSET) entry = n;
            }
        }
        if (lane == 0) {
            // chain walk at shared-memory latency
            uint64_t end = min(n, tile_abs + TILE), o = entry;
            uint32_t k = 0;
            while (o < end) {
                if (o + 4 > n) break;
                int32_t bs = (int32_t)w.u32((uint32_t)(o - tile_abs));
                if (bs < 32) {
                    o = BAD_OFFSET;
                    break;
                }
                if (o + 4 + (uint64_t)bs > n) break;
                S.rec_off[k++] = (uint32_t)(o - tile_abs);
                o += 4 + (uint64_t)bs;
            }
            S.n_rec = k, S.entry = entry, S.exit_ = o;
            S.first_mb = ~0ull, S.last_mb = 0ull;
        }
    }
    __syncthreads();
}
