// BGZF inflate on the device: one warp per BGZF block (RFC 1951 DEFLATE: stored, fixed and dynamic Huffman blocks).
//
// Replaces the bgzf.o / zlib inflate of the reference's libbam (sam/bgzf.h:34-134; `bam_read1` -> `bgzf_read` ->
// `inflate_block`), which SURVEY.md section 0 measures at ~75 % of getclip's run time on the host. BGZF blocks are
// independent deflate streams of at most 64 KiB of output, so a whole BAM offers tens of thousands of blocks to
// decode concurrently. Inside a warp lane 0 owns the bit reader and the Huffman decode (serial by nature) and turns the
// bit stream into batches of up to 32 tokens (a literal byte or a (length, distance) match); the 32 lanes then place
// the batch with one prefix sum and copy in parallel: one token per lane for literals and short matches whose source
// lies before the batch, the whole warp per token for long or batch-dependent matches (in stream order). BAM data is
// match-dominated (overlapping reads: ~96 % of the bytes of the C2 workload come from matches of mean length 8), so
// the per-token warp-wide work is what had to be amortised. Decode tables live in shared memory (7 KB per warp):
// a 10-bit single-lookup table for literal/length codes and an 8-bit one for distance codes whose entries already
// carry the base value and the number of extra bits, and a canonical (count / sorted-symbol) fallback for the rare
// longer codes.
#include "common.cuh"
#include <cstddef>

namespace {

constexpr int LIT_FAST = 10, DIST_FAST = 8, BLOCKS_PER_CTA = 4, MAX_BATCH = 32, COOP_LEN = 16, RING = 128;

struct WarpTables {
    // entries: bits 0-3 code length, 4-7 number of extra bits, bit 8 literal, bit 9 end of block, bit 10 invalid symbol,
    // bits 16-31 value (literal byte, length base or distance base); 0 = not a short code
    uint32_t lit_fast[1 << LIT_FAST];
    uint32_t dist_fast[1 << DIST_FAST];
    // (ring and tok double as the 288 x u16 scratch for canonical codes while a table is being built)
    uint32_t ring[RING + 2];  // the next RING words of the compressed stream (refilled by all lanes before each batch)
    uint32_t tok[MAX_BATCH];  // token batch: bit 31 literal (byte in bits 0-7), else length in bits 0-8 and distance in bits 9-24
    uint16_t lit_sym[288], dist_sym[32];  // symbols sorted by (code length, symbol) for the canonical slow path
    uint16_t lit_count[16], dist_count[16];
    uint8_t lens[320];   // code lengths: literal/length alphabet followed by the distance alphabet
    uint8_t cl_fast[128];  // code-length alphabet: (len << 5) | symbol, 7-bit lookup
};
static_assert(offsetof(WarpTables, tok) + sizeof(uint32_t) * MAX_BATCH - offsetof(WarpTables, ring) >= 288 * sizeof(uint16_t),
              "ring + tok double as the canonical-code scratch of build_table");


__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Bit reader in lane 0's registers: three consecutive 32-bit words of the input and a bit position inside the first;
// the 32 bits that follow the position are one funnel shift away, consuming bits is one add and - every 32 bits - a
// register rotation whose load is not needed before another 32 bits have been consumed (latency off the critical path).
// The file image is padded past its end, so reading ahead is safe.
struct BitReader {
    const uint32_t *wp;  // address of `lo`
    uint32_t lo, hi, nx, bp;
    __device__ void init(const uint8_t *p)
    {
        wp = (const uint32_t *)((uintptr_t)p & ~(uintptr_t)3);
        bp = (uint32_t)((uintptr_t)p & 3) * 8;
        lo = __ldg(wp), hi = __ldg(wp + 1), nx = __ldg(wp + 2);
    }
    __device__ void init_at(const uint32_t *base, uint32_t bitpos)
    {
        wp = base + (bitpos >> 5), bp = bitpos & 31;
        lo = __ldg(wp), hi = __ldg(wp + 1), nx = __ldg(wp + 2);
    }
    __device__ uint32_t bit_offset(const uint32_t *base) const { return (uint32_t)(wp - base) * 32 + bp; }
    __device__ __forceinline__ uint32_t window() const { return __funnelshift_r(lo, hi, bp); }  // next 32 bits
    __device__ __forceinline__ void consume(uint32_t n)  // n <= 32
    {
        bp += n;
        if (bp >= 32) {
            bp -= 32;
            lo = hi, hi = nx;
            ++wp;
            nx = __ldg(wp + 2);
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return window() & ((1u << n) - 1); }  // n < 32
    __device__ __forceinline__ uint32_t bits(uint32_t n)
    {
        uint32_t v = peek(n);
        consume(n);
        return v;
    }
    __device__ void align_byte() { consume((8 - (bp & 7)) & 7); }
    __device__ const uint8_t *byte_ptr() const { return (const uint8_t *)wp + (bp >> 3); }  // after align_byte()
};

constexpr uint32_t E_LITERAL = 1u << 8, E_END = 1u << 9, E_INVALID = 1u << 10;

// table entry of a symbol (without its code length, which the caller ORs into bits 0-3)
__device__ __forceinline__ uint32_t lit_entry(int s)
{
    if (s < 256) return E_LITERAL | (uint32_t)s << 16;
    if (s == 256) return E_END;
    if (s > 285) return E_INVALID;
    return (uint32_t)c_len_extra[s - 257] << 4 | (uint32_t)c_len_base[s - 257] << 16;
}
__device__ __forceinline__ uint32_t dist_entry(int s)
{
    if (s > 29) return E_INVALID;
    return (uint32_t)c_dist_extra[s] << 4 | (uint32_t)c_dist_base[s] << 16;
}

// Build the decode tables of one alphabet from its code lengths (canonical Huffman, RFC 1951 3.2.2).
// The group leader assigns codes serially (<= 288 symbols); all lanes of the group fill the single-lookup table.
template <bool IS_LIT, int GROUP>
__device__ void build_table(const uint8_t *lens, int n, uint32_t *fast, int fast_bits, uint16_t *sym_sorted, uint16_t *count,
                            uint16_t *code, uint32_t lane, uint32_t gm)
{
    for (int i = lane; i < (1 << fast_bits); i += GROUP) fast[i] = 0;
    if (lane < 16) count[lane] = 0;
    __syncwarp(gm);
    if (lane == 0) {
        for (int s = 0; s < n; ++s) count[lens[s]]++;
        count[0] = 0;
        uint16_t next_code[16], offs[16];
        uint32_t c = 0;
        offs[1] = 0;
        next_code[0] = 0;
        for (int b = 1; b <= 15; ++b) {
            c = (c + count[b - 1]) << 1;
            next_code[b] = (uint16_t)c;
            if (b < 15) offs[b + 1] = offs[b] + count[b];
        }
        for (int s = 0; s < n; ++s) {
            int l = lens[s];
            if (l) {
                code[s] = next_code[l]++;
                sym_sorted[offs[l]++] = (uint16_t)s;
            }
        }
    }
    __syncwarp(gm);
    for (int s = lane; s < n; s += GROUP) {
        int l = lens[s];
        if (l && l <= fast_bits) {
            uint32_t rev = __brev((uint32_t)code[s]) >> (32 - l);  // codes are sent MSB first, bits are read LSB first
            uint32_t e = (uint32_t)l | (IS_LIT ? lit_entry(s) : dist_entry(s));
            for (uint32_t k = rev; k < (1u << fast_bits); k += 1u << l) fast[k] = e;
        }
    }
    __syncwarp(gm);
}

// canonical decode for codes longer than the lookup width, on a 32-bit window; returns the symbol and its code length
__device__ int slow_decode(uint32_t w, const uint16_t *count, const uint16_t *sym_sorted, uint32_t *code_len)
{
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)(w & 1);
        w >>= 1;
        int c = count[len];
        if (code - c < first) {
            *code_len = (uint32_t)len;
            return sym_sorted[index + (code - first)];
        }
        index += c, first += c;
        first <<= 1, code <<= 1;
    }
    return -1;
}

// one symbol as a table entry (bits 0-3 = its code length); E_INVALID for an undecodable code
template <bool IS_LIT>
__device__ __forceinline__ uint32_t decode_entry(uint32_t w, const uint32_t *fast, int fast_bits, const uint16_t *count,
                                                 const uint16_t *sym_sorted)
{
    uint32_t e = fast[w & ((1u << fast_bits) - 1)];
    if (e) return e;
    uint32_t l = 0;
    int s = slow_decode(w, count, sym_sorted, &l);
    if (s < 0) return E_INVALID;
    return l | (IS_LIT ? lit_entry(s) : dist_entry(s));
}

}  // namespace

struct InflateBlock {
    uint64_t coff, uoff;
    uint32_t clen, ulen;
};

// Two BGZF blocks per warp, 16 lanes each ("group"): the serial Huffman decode is one lane's work, so with one block per
// warp ~80 % of the issued instructions had a single active lane. Two decoder lanes that execute the same instruction
// stream halve that. Each group runs a small state machine (block header / token batches / done); the warp meets at the
// top of every round so that groups in the same state execute their section converged.
template <int GROUP>
__global__ void __launch_bounds__(BLOCKS_PER_CTA * GROUP, 8)
    inflate_bgzf(const uint8_t *__restrict__ file, const InflateBlock *__restrict__ blocks, uint32_t n_blocks, uint8_t *__restrict__ out,
                 uint32_t *__restrict__ error)
{
    constexpr int GROUPS = 32 / GROUP, BATCH = GROUP;
    __shared__ WarpTables tables[BLOCKS_PER_CTA];
    enum { S_HEADER, S_TOKENS, S_DONE };
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t glane = lane & (GROUP - 1), grp = lane / GROUP, leader = grp * GROUP;
    const uint32_t gm = (uint32_t)((1ull << GROUP) - 1ull) << leader;  // this group's lanes
    const uint32_t b = blockIdx.x * BLOCKS_PER_CTA + wid * GROUPS + grp;
    WarpTables &T = tables[wid * GROUPS + grp];
    int state = b < n_blocks ? S_HEADER : S_DONE;
    InflateBlock blk = {0, 0, 0, 0};
    if (state != S_DONE) blk = blocks[b];
    uint8_t *dst = out + blk.uoff;
    const uint32_t *wbase = (const uint32_t *)((uintptr_t)(file + blk.coff) & ~(uintptr_t)3);  // bit positions count from here
    BitReader br;
    br.init(file + blk.coff);  // (only the group leader's copy is used)
    uint32_t pos = 0, final_block = 0, bitpos = 0;
    bool bad = false;
    for (;;) {
        __syncwarp();
        if (state == S_HEADER) {
            uint32_t hdr = 0;
            if (glane == 0) hdr = br.bits(3);
            hdr = __shfl_sync(gm, hdr, leader);
            final_block = hdr & 1;
            const uint32_t type = hdr >> 1;
            if (type == 0) {  // stored
                uint32_t len = 0;
                const uint8_t *src = nullptr;
                if (glane == 0) {
                    br.align_byte();
                    len = br.bits(16);
                    br.consume(16);  // NLEN
                    src = br.byte_ptr();
                    br.init(src + len);
                }
                len = __shfl_sync(gm, len, leader);
                src = (const uint8_t *)__shfl_sync(gm, (unsigned long long)src, leader);
                if (pos + len > blk.ulen) bad = true;
                else {
                    for (uint32_t i = glane; i < len; i += GROUP) dst[pos + i] = src[i];
                    pos += len;
                }
                state = (bad || final_block) ? S_DONE : S_HEADER;
            } else if (type == 1 || type == 2) {
                int n_lit = 288, n_dist = 30, ok = 1;
                if (type == 1) {  // fixed Huffman codes (RFC 1951 3.2.6)
                    for (int i = glane; i < 288; i += GROUP) T.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                    for (int i = glane; i < 32; i += GROUP) T.lens[288 + i] = i < 30 ? 5 : 0;
                } else {  // dynamic: code lengths are themselves Huffman coded (3.2.7), decoded serially by the leader
                    if (glane == 0) {
                        n_lit = (int)br.bits(5) + 257;
                        n_dist = (int)br.bits(5) + 1;
                        int n_cl = (int)br.bits(4) + 4;
                        uint8_t cl[19];
                        for (int i = 0; i < 19; ++i) cl[i] = 0;
                        for (int i = 0; i < n_cl; ++i) cl[c_cl_order[i]] = (uint8_t)br.bits(3);
                        // 7-bit lookup for the code-length alphabet
                        for (int i = 0; i < 128; ++i) T.cl_fast[i] = 0;
                        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, next[8];
                        for (int i = 0; i < 19; ++i) cnt[cl[i]]++;
                        cnt[0] = 0;
                        uint32_t c = 0;
                        next[0] = 0;
                        for (int l = 1; l < 8; ++l) {
                            c = (c + cnt[l - 1]) << 1;
                            next[l] = c;
                        }
                        for (int sy = 0; sy < 19; ++sy) {
                            int l = cl[sy];
                            if (!l) continue;
                            uint32_t rev = __brev(next[l]++) >> (32 - l);
                            for (uint32_t k = rev; k < 128; k += 1u << l) T.cl_fast[k] = (uint8_t)(l << 5 | sy);
                        }
                        int i = 0, total = n_lit + n_dist;
                        if (n_lit > 286 || n_dist > 30) ok = 0;
                        while (ok && i < total) {
                            uint8_t e = T.cl_fast[br.peek(7)];
                            if (!e) {
                                ok = 0;
                                break;
                            }
                            br.consume(e >> 5);
                            int sy = e & 31;
                            if (sy < 16) T.lens[i++] = (uint8_t)sy;
                            else {
                                int rep, v = 0;
                                if (sy == 16) {
                                    if (i == 0) {
                                        ok = 0;
                                        break;
                                    }
                                    v = T.lens[i - 1];
                                    rep = 3 + (int)br.bits(2);
                                } else if (sy == 17) rep = 3 + (int)br.bits(3);
                                else rep = 11 + (int)br.bits(7);
                                if (i + rep > total) {
                                    ok = 0;
                                    break;
                                }
                                while (rep--) T.lens[i++] = (uint8_t)v;
                            }
                        }
                    }
                    ok = __shfl_sync(gm, ok, leader);
                    n_lit = __shfl_sync(gm, n_lit, leader);
                    n_dist = __shfl_sync(gm, n_dist, leader);
                    __syncwarp(gm);
                    if (ok) {
                        // the distance lengths follow the literal/length lengths directly: move them to their own slot
                        uint8_t v[32 / GROUP];
#pragma unroll
                        for (int j = 0; j < 32 / GROUP; ++j) {
                            int i = (int)glane + j * GROUP;
                            v[j] = i < n_dist ? T.lens[n_lit + i] : 0;
                        }
                        __syncwarp(gm);
#pragma unroll
                        for (int j = 0; j < 32 / GROUP; ++j) T.lens[288 + glane + j * GROUP] = v[j];
                        for (int i = n_lit + (int)glane; i < 288; i += GROUP) T.lens[i] = 0;
                    }
                }
                if (!ok) bad = true, state = S_DONE;
                else {
                    __syncwarp(gm);
                    build_table<true, GROUP>(T.lens, type == 1 ? 288 : n_lit, T.lit_fast, LIT_FAST, T.lit_sym, T.lit_count, (uint16_t *)T.ring, glane,
                                      gm);
                    build_table<false, GROUP>(T.lens + 288, n_dist, T.dist_fast, DIST_FAST, T.dist_sym, T.dist_count, (uint16_t *)T.ring, glane, gm);
                    if (glane == 0) bitpos = br.bit_offset(wbase);
                    bitpos = __shfl_sync(gm, bitpos, leader);
                    state = S_TOKENS;
                }
            } else
                bad = true, state = S_DONE;
        }
        // Groups in the token state run the batch converged: every warp-level primitive below names all of their lanes
        // (mt), so the two decoding leaders execute the same instructions side by side. (The header section above uses
        // per-group masks and may run the groups one after the other; it is rare.)
        const uint32_t mt = __ballot_sync(0xffffffffu, state == S_TOKENS);
        if (state == S_TOKENS) {
            // One batch: the leader turns bits into up to GROUP tokens, the group places and copies them. The bit reader is
            // just a bit position here: the words come from a shared-memory ring that the group refills (coalesced) first.
            {   // a batch consumes at most GROUP * 48 bits; 3 * GROUP words are fetched
                const uint32_t cw = bitpos >> 5;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint32_t idx = cw + glane + GROUP * j, v = __ldg(wbase + idx);
                    T.ring[idx & (RING - 1)] = v;
                    if ((idx & (RING - 1)) < 2) T.ring[RING + (idx & (RING - 1))] = v;  // mirror: three consecutive words never wrap
                }
            }
            __syncwarp(mt);
            uint32_t ntok = 0;
            int status = 0;  // 0 = more to come, 1 = end of block, -1 = corrupt
            if (glane == 0) {
                uint32_t bp = bitpos;
                while (ntok < BATCH) {
                    // 64 bits of the stream at bp: a literal/length code with its extra bits (<= 20) and a distance code with
                    // its extra bits (<= 28) both fit, so one look at the ring serves the whole token
                    const uint32_t *rp = T.ring + ((bp >> 5) & (RING - 1));
                    const uint32_t r0 = rp[0], r1 = rp[1], r2 = rp[2];
                    const uint32_t w = __funnelshift_r(r0, r1, bp), whi = __funnelshift_r(r1, r2, bp);
                    const uint32_t e = decode_entry<true>(w, T.lit_fast, LIT_FAST, T.lit_count, T.lit_sym);
                    const uint32_t l = e & 15;
                    uint32_t tok;
                    if (e & E_LITERAL) {
                        tok = 0x80000000u | (e >> 16);
                        bp += l;
                    } else if (e & (E_END | E_INVALID)) {
                        bp += l;
                        status = (e & E_END) ? 1 : -1;
                        break;
                    } else {
                        const uint32_t x = (e >> 4) & 15, used = l + x;
                        const uint32_t len = (e >> 16) + ((w >> l) & ((1u << x) - 1));  // l + x <= 20 bits of the window
                        const uint32_t w2 = __funnelshift_r(w, whi, used);
                        const uint32_t d = decode_entry<false>(w2, T.dist_fast, DIST_FAST, T.dist_count, T.dist_sym);
                        if (d & E_INVALID) {
                            status = -1;
                            break;
                        }
                        const uint32_t l2 = d & 15, x2 = (d >> 4) & 15;
                        const uint32_t dist = (d >> 16) + ((w2 >> l2) & ((1u << x2) - 1));  // l2 + x2 <= 28
                        bp += used + l2 + x2;
                        tok = len | dist << 9;
                    }
                    T.tok[ntok++] = tok;
                }
                bitpos = bp;
            }
            ntok = __shfl_sync(mt, ntok, leader);
            status = __shfl_sync(mt, status, leader);
            bitpos = __shfl_sync(mt, bitpos, leader);
            __syncwarp(mt);
            const uint32_t tok = glane < ntok ? T.tok[glane] : 0u;
            const bool is_lit = tok >> 31;
            const uint32_t n = is_lit ? 1u : (tok & 511u), dist = (tok >> 9) & 0xffffu;
            uint32_t incl = n;  // inclusive prefix sum of the output sizes
#pragma unroll
            for (int d = 1; d < GROUP; d <<= 1) {
                uint32_t v = __shfl_up_sync(mt, incl, d, GROUP);
                if (glane >= (uint32_t)d) incl += v;
            }
            const uint32_t total = __shfl_sync(mt, incl, leader + GROUP - 1), off = incl - n, o = pos + off;
            const bool is_match = !is_lit && n != 0;
            const bool fail = status < 0 || pos + total > blk.ulen || (__ballot_sync(mt, is_match && dist > o) & gm) != 0;
            // matches whose source ends before this batch's output are independent of the other tokens
            const bool coop = !fail && is_match && (n > COOP_LEN || dist < off + n);
            if (!fail) {
                if (is_lit) dst[o] = (uint8_t)tok;
                else if (is_match && !coop) {  // n <= COOP_LEN: all loads are issued before the first store waits for one
                    const uint8_t *src = dst + o - dist;
                    uint8_t v[COOP_LEN];
#pragma unroll
                    for (uint32_t k = 0; k < COOP_LEN; ++k)
                        if (k < n) v[k] = src[k];
#pragma unroll
                    for (uint32_t k = 0; k < COOP_LEN; ++k)
                        if (k < n) dst[o + k] = v[k];
                }
            }
            uint32_t pending = __ballot_sync(mt, coop) & gm;
            __syncwarp(mt);  // the stores above are visible to the lanes that copy below
            while (__any_sync(mt, pending != 0)) {  // long or batch-dependent matches: whole group per token, in stream order
                const bool act = pending != 0;
                const int t = act ? __ffs(pending) - 1 : (int)leader;
                pending &= pending - 1;
                const uint32_t o_t = __shfl_sync(mt, o, t), n_s = __shfl_sync(mt, n, t), d_t = __shfl_sync(mt, dist, t);
                const uint32_t n_t = act ? n_s : 0u;
                const uint8_t *src = dst + o_t - d_t;
                if (d_t >= n_t) {
                    for (uint32_t i = glane; i < n_t; i += GROUP) dst[o_t + i] = src[i];
                } else {  // overlapping match: the last d_t bytes repeat
                    for (uint32_t i = glane; i < n_t; i += GROUP) dst[o_t + i] = src[i % d_t];
                }
                __syncwarp(mt);
            }
            if (fail) bad = true, state = S_DONE;
            else {
                pos += total;
                if (status == 1) {
                    if (glane == 0) br.init_at(wbase, bitpos);  // back to the register reader for the next block header
                    state = final_block ? S_DONE : S_HEADER;
                }
            }
        }
        if (__all_sync(0xffffffffu, state == S_DONE)) break;
    }
    if (b < n_blocks && (bad || pos != blk.ulen) && glane == 0) atomicOr(error, 1u);
}

// BGZF blocks per warp: 1 (default) or 2 (SEEKSV_B200_INFLATE_GROUP=16, two 16-lane groups per warp; measured slower on
// C2: fewer warps per SM leave the dependent shared-memory lookups of the decode loop exposed)
static void launch_inflate(cudaStream_t s, const uint8_t *d_file, const void *d_blocks, uint32_t n_blocks, uint8_t *d_out, uint32_t *d_err)
{
    static const int group = [] {
        const char *e = getenv("SEEKSV_B200_INFLATE_GROUP");
        return e && atoi(e) == 16 ? 16 : 32;
    }();
    const uint32_t grid = (n_blocks + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA;
    if (group == 16) inflate_bgzf<16><<<grid, BLOCKS_PER_CTA * 16, 0, s>>>(d_file, (const InflateBlock *)d_blocks, n_blocks, d_out, d_err);
    else inflate_bgzf<32><<<grid, BLOCKS_PER_CTA * 32, 0, s>>>(d_file, (const InflateBlock *)d_blocks, n_blocks, d_out, d_err);
}

// asynchronous launch over a range of blocks; *d_err is OR-ed with 1 when a block is corrupt
int inflate_launch(cudaStream_t s, const uint8_t *d_file, const void *d_blocks, uint32_t n_blocks, uint8_t *d_out, uint32_t *d_err)
{
    if (!n_blocks) return 0;
    launch_inflate(s, d_file, d_blocks, n_blocks, d_out, d_err);
    return cudaGetLastError() == cudaSuccess ? 0 : SVB_ERR_CUDA;
}

// synchronous wrapper: file image + block table already on the device
int inflate_on_device(svb_ctx *ctx, const uint8_t *d_file, const void *d_blocks, uint32_t n_blocks, uint8_t *d_out, double out_bytes)
{
    cudaStream_t s = ctx->stream;
    DevBuf<uint32_t> err;
    CK(err.alloc(1, s));
    CK(cudaMemsetAsync(err.p, 0, 4, s));
    if (n_blocks) {
        ProfScope ps(ctx, "inflate_bgzf", out_bytes);
        launch_inflate(s, d_file, d_blocks, n_blocks, d_out, err.p);
    }
    uint32_t h = 0;
    CK(cudaMemcpyAsync(&h, err.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (h) return svb_fail(ctx, SVB_ERR_FORMAT, "BGZF inflate failed (corrupt deflate stream)");
    return 0;
}
