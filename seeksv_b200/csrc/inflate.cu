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
#include <cstring>
#include <mutex>

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
    const uint32_t end = (uint32_t)((uintptr_t)(file + blk.coff) & 3) * 8 + blk.clen * 8;      // end of the deflate payload
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
                uint32_t len = 0, ok = 0;
                const uint8_t *src = nullptr;
                if (glane == 0) {
                    br.align_byte();
                    len = br.bits(16);
                    const uint32_t nlen = br.bits(16);
                    src = br.byte_ptr();
                    ok = (len ^ nlen) == 0xffffu && br.bit_offset(wbase) + len * 8 <= end;  // inside the payload
                    if (ok) br.init(src + len);
                }
                len = __shfl_sync(gm, len, leader), ok = __shfl_sync(gm, ok, leader);
                src = (const uint8_t *)__shfl_sync(gm, (unsigned long long)src, leader);
                if (!ok || pos + len > blk.ulen) bad = true;
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
                    if (bitpos >= end) bad = true, state = S_DONE;  // the header ran out of the payload
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
            const bool fail = status < 0 || bitpos > end || pos + total > blk.ulen || (__ballot_sync(mt, is_match && dist > o) & gm) != 0;
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

// =====================================================================================================================
// Default kernel (round 2): SPECULATIVE decode - all 32 lanes of the warp decode Huffman codes of the same deflate block.
//
// The serial form above keeps 31 lanes idle while lane 0 decodes (80 % of its issued instructions had one active thread).
// Huffman streams resynchronise: a decoder started at a wrong bit position falls onto true code boundaries after a few
// tokens. So the bit range of a deflate block is cut into 32 equal sub-ranges and every lane decodes its own, lane 0 from the
// true start, the others from a guess. Then rounds: lane i+1 restarts at the bit position where lane i LEFT its sub-range;
// a lane whose entry did not change keeps its result. Lane 0 is exact, so after round r lanes 0..r are exact (induction) and
// the fixed point is the serial decode whatever the guesses did - they only decide how many rounds it takes (on the C2
// workload 97 % of the blocks are final after ONE repeat; lock-step work 0.14 of the serial token count,
// tools/spec_decode_sim.cpp). The passes only count (bytes, matches per lane); a prefix sum over the lanes then gives
// every lane its place in a per-warp token list (global scratch, one slot per resident warp) and the last pass emits the tokens
// (a literal byte or a (length, distance) match, 4 bytes each). The copy phase takes the list in stream order, up to 32 tokens
// / 1 KiB of output at a time, and assembles that piece of output in SHARED memory: literals and the matches whose source lies
// in front of the piece go in parallel (one per lane, whole warp for long ones), the matches that read this piece's own output
// follow one at a time through shared memory (a ~30-cycle round trip instead of a trip to L2 - that loop was 60 % of the first
// version's time, profiles/r2_inflate.md); the piece is then written out coalesced.
//
// Tables: single lookup for codes up to 10 (literal/length) / 8 (distance) bits, SECOND-LEVEL tables behind the root for the
// longer ones (replaces the 15-step canonical loop; tools/huff2_proto.cpp), built warp-parallel: symbol ranks by
// __match_any_sync, sub-table widths from the cumulative code space. Sizes follow zlib's `enough`: 288 symbols / root 10 / 15
// bits need <= 1334 entries, 32 / 8 / 15 <= 402 (complete codes; incomplete ones are refused exactly where zlib refuses them).
namespace {
constexpr int LIT_ROOT = 10, DIST_ROOT = 8, LIT_CAP = 1344, DIST_CAP = 416, SPEC_WARPS = 4, MIN_SPAN_BITS = 64;
constexpr uint32_t TOK_CAP = 20480;  // tokens per emission segment (a lane's sub-range holds <= 2^19 / 32 = 16384 tokens)
constexpr uint32_t SPAN = 1024;      // output bytes assembled in shared memory per token batch
constexpr uint32_t T_SUB = 1u << 5, T_LIT = 1u << 6, T_END = 1u << 7, T_BAD = 1u << 14;
// table entry: bits 0-4 bits to consume, bit 5 pointer to a sub-table (bits 8-11 its index width, bits 16-31 its offset),
// bit 6 literal, bit 7 end of block, bits 8-11 number of extra bits, bit 14 invalid symbol, bits 16-31 value; 0 = no such code

struct SpecMem {
    uint32_t lit[LIT_CAP];
    uint32_t dist[DIST_CAP];
    union {
        struct {                   // table building
            uint8_t lens[320];     // literal/length code lengths, the distance alphabet from [288]
            uint8_t cl_fast[128];  // code-length alphabet: (len << 5) | symbol, 7-bit lookup
            uint16_t rank[288];    // rank of a symbol among the symbols of its code length
        };
        alignas(16) uint8_t span[SPAN + 16];  // copy phase: the output bytes of the current token batch (from index address & 3)
        uint32_t ring[8 * 32];               // decode passes: the lanes' input rings
    };
    uint32_t cnt[16], first[16], cum[16];
};

__device__ __forceinline__ uint32_t spec_lit_entry(int s)
{
    if (s < 256) return T_LIT | (uint32_t)s << 16;
    if (s == 256) return T_END;
    if (s > 285) return T_BAD;
    return (uint32_t)c_len_extra[s - 257] << 8 | (uint32_t)c_len_base[s - 257] << 16;
}
__device__ __forceinline__ uint32_t spec_dist_entry(int s)
{
    if (s > 29) return T_BAD;
    return (uint32_t)c_dist_extra[s] << 8 | (uint32_t)c_dist_base[s] << 16;
}

// Canonical Huffman tables (RFC 1951 3.2.2) of one alphabet, built by the whole warp. false = over-subscribed, incomplete
// (zlib inftrees.c: an incomplete code is accepted only when its longest code has one bit) or too large for the table.
template <bool IS_LIT>
__device__ bool spec_build(SpecMem &M, const uint8_t *lens, int n, uint32_t *T, uint32_t lane)
{
    constexpr int R = IS_LIT ? LIT_ROOT : DIST_ROOT;
    constexpr uint32_t CAP = IS_LIT ? LIT_CAP : DIST_CAP;
    for (uint32_t i = lane; i < (1u << R); i += 32) T[i] = 0;
    if (lane < 16) M.cnt[lane] = 0;
    __syncwarp();
    for (int s0 = 0; s0 < n; s0 += 32) {  // counts per length + rank of every symbol inside its length
        const int s = s0 + (int)lane;
        const uint32_t l = s < n ? lens[s] : 0u;
        const uint32_t peers = __match_any_sync(0xffffffffu, l);
        const uint32_t r = __popc(peers & ((1u << lane) - 1u)), base = M.cnt[l];
        if (s < n) M.rank[s] = (uint16_t)(base + r);
        __syncwarp();
        if (r == 0) M.cnt[l] = base + __popc(peers);
        __syncwarp();
    }
    if (lane == 0) {  // first code per length and cumulative code space in units of 2^-15
        uint32_t cum = 0, maxl = 0;
        for (int l = 1; l <= 15; ++l) {
            M.first[l] = cum >> (15 - l);
            const uint32_t c = M.cnt[l];
            if (c) maxl = l;
            cum += c << (15 - l);
            M.cum[l] = cum;
        }
        M.first[0] = cum, M.cum[0] = maxl;
    }
    __syncwarp();
    const uint32_t space = M.first[0], maxl = M.cum[0];
    if (space > 32768u || (space < 32768u && maxl > 1)) return false;
    if (maxl > (uint32_t)R) {
        // every R-bit prefix from the first long code on heads a sub-table as wide as its longest code - the code that ends
        // where the prefix ends (codes ascend with their length)
        uint32_t off = 1u << R;
        for (uint32_t p0 = M.cum[R] >> (15 - R); p0 < (1u << R); p0 += 32) {
            const uint32_t p = p0 + lane;
            const bool have = p < (1u << R);
            uint32_t w = 0;
            if (have) {
                const uint32_t lim = (p + 1) << (15 - R);
                int l = R + 1;
                while (l < 15 && M.cum[l] < lim) ++l;
                w = (uint32_t)(l - R);
            }
            const uint32_t sz = have ? 1u << w : 0u;
            uint32_t incl = sz;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (uint32_t)d) incl += v;
            }
            const uint32_t my = off + incl - sz;
            off += __shfl_sync(0xffffffffu, incl, 31);
            if (have && my + sz <= CAP) T[__brev(p) >> (32 - R)] = (uint32_t)R | T_SUB | w << 8 | my << 16;
        }
        if (off > CAP) return false;
    }
    __syncwarp();
    for (int s = (int)lane; s < n; s += 32) {
        const uint32_t l = lens[s];
        if (!l) continue;
        const uint32_t rev = __brev(M.first[l] + M.rank[s]) >> (32 - l);  // codes are sent MSB first, bits are read LSB first
        const uint32_t e = IS_LIT ? spec_lit_entry(s) : spec_dist_entry(s);
        if (l <= (uint32_t)R) {
            for (uint32_t k = rev; k < (1u << R); k += 1u << l) T[k] = e | l;
        } else {
            const uint32_t root = T[rev & ((1u << R) - 1u)], w = (root >> 8) & 15u, off = root >> 16;
            for (uint32_t k = rev >> R; k < (1u << w); k += 1u << (l - R)) T[off + k] = e | (l - R);
        }
    }
    __syncwarp();
    return true;
}

enum { SP_OK = 0, SP_END = 1, SP_BAD = 2 };
struct Span {
    uint32_t exit, nbytes, ntok;
    int status;  // SP_OK: left the range at `exit`; SP_END: end-of-block code consumed, `exit` behind it; SP_BAD: invalid code
};

// One lane decodes tokens from bit `bp` until its position reaches `bound` (tested between tokens). EMIT: the tokens are appended
// to tk[] (bit 31 literal, byte in bits 0-7; else length in bits 0-8, distance in bits 9-24); otherwise only counted. Bit
// positions count from the word wbase.
// Input words reach the lane through a private 8-word ring in shared memory filled by cp.async: a word is requested five
// word-crossings (~11 tokens) before its first use and NO register waits for it in between. (Read-ahead in registers does not
// work: rotating w3 -> w2 -> ... reads the register of a load in flight at the very next crossing, and the loop spent 15-20 % of
// its time on that scoreboard; profiles/r2_inflate.md.) Layout ring[(word & 7) * 32 + lane]: conflict-free.
__device__ __forceinline__ void ring_request(uint32_t *ring_lane, const uint32_t *wbase, uint32_t word)
{
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(ring_lane + (word & 7u) * 32u);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n\tcp.async.commit_group;" ::"r"(sm), "l"(wbase + word) : "memory");
}
template <bool EMIT>
__device__ __forceinline__ void spec_decode(const uint32_t *__restrict__ wbase, const uint32_t *lit, const uint32_t *dtab, uint32_t bp,
                                            const uint32_t bound, Span &r, uint32_t *tk, uint32_t *ring_lane)
{
    uint32_t wi = bp >> 5, sh = bp & 31u;
    uint32_t w0 = 0, w1 = 0, w2 = 0;  // the window: words wi, wi + 1, wi + 2
    if (bp < bound) {
        w0 = __ldg(wbase + wi), w1 = __ldg(wbase + wi + 1), w2 = __ldg(wbase + wi + 2);
#pragma unroll
        for (uint32_t k = 3; k < 8; ++k) ring_request(ring_lane, wbase, wi + k);  // five groups in flight
    }
    uint32_t nb = 0, nt = 0;
    int status = SP_OK;
    // (no break / early exit: the loop has ONE way out, behind which the compiler reconverges the warp - with early exits the
    // lanes stayed split into groups for the rest of the block and every later shuffle took the divergent slow path)
    while (bp < bound && status == SP_OK) {
        // 64 bits of the stream: a literal/length code with its extra bits (<= 20) and a distance code with its extra bits
        // (<= 28) both fit
        const uint32_t w = __funnelshift_r(w0, w1, sh), whi = __funnelshift_r(w1, w2, sh);
        uint32_t e = lit[w & ((1u << LIT_ROOT) - 1u)];
        uint32_t used = e & 31u;
        if (e & T_SUB) {
            e = lit[(e >> 16) + ((w >> LIT_ROOT) & ((1u << ((e >> 8) & 15u)) - 1u))];
            used = LIT_ROOT + (e & 31u);
        }
        if (e & T_LIT) {
            if (EMIT) tk[nt] = 0x80000000u | (e >> 16);
            nb += 1, nt += 1;
        } else if (e == 0 || (e & (T_END | T_BAD))) {
            status = (e & T_END) ? SP_END : SP_BAD;
            if (!(e & T_END)) used = 0;
        } else {
            const uint32_t x = (e >> 8) & 15u, len = (e >> 16) + ((w >> used) & ((1u << x) - 1u));
            used += x;
            const uint32_t v = __funnelshift_r(w, whi, used);
            uint32_t d = dtab[v & ((1u << DIST_ROOT) - 1u)];
            uint32_t used2 = d & 31u;
            if (d & T_SUB) {
                d = dtab[(d >> 16) + ((v >> DIST_ROOT) & ((1u << ((d >> 8) & 15u)) - 1u))];
                used2 = DIST_ROOT + (d & 31u);
            }
            if (d == 0 || (d & T_BAD)) {
                status = SP_BAD, used = 0;
            } else {
                const uint32_t x2 = (d >> 8) & 15u, dist = (d >> 16) + ((v >> used2) & ((1u << x2) - 1u));
                used += used2 + x2;
                if (EMIT) tk[nt] = len | dist << 9;
                nb += len, nt += 1;
            }
        }
        bp += used, sh += used;
#pragma unroll
        for (int twice = 0; twice < 2; ++twice)  // a token has at most 48 bits: up to two word crossings
            if (sh >= 32) {
                sh -= 32, ++wi;
                ring_request(ring_lane, wbase, wi + 7);  // into the slot of word wi - 1
                asm volatile("cp.async.wait_group 5;" ::: "memory");  // the oldest request - word wi + 2 - has landed
                w0 = w1, w1 = w2, w2 = ring_lane[((wi + 2) & 7u) * 32u];
            }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // nothing of this call may land in the ring later
    r.exit = bp, r.nbytes = nb, r.ntok = nt, r.status = status;
}

// Scratch slots for the token lists: one per RESIDENT warp of this kernel, whatever launch or stream it belongs to. A warp
// claims a free bit (starting at its SM's word, where one is free by construction) and gives it back at the end.
// (Warp-uniform control flow: lane 0 does the atomics, every lane runs the loop. With the loop inside `if (lane == 0)` the warp
// came out of it split into {lane 0} and {the others} and stayed split for the whole block - every shuffle took the divergent
// slow path and every instruction issued twice; profiles/r2_inflate.md.)
__device__ uint32_t slot_claim(uint32_t *bitmap, uint32_t n_words, uint32_t lane)
{
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    uint32_t w = smid % n_words;
    for (;;) {
        uint32_t got = 0xffffffffu;
        if (lane == 0) {
            const uint32_t cur = atomicOr(bitmap + w, 0u);
            if (~cur) {
                const uint32_t bit = 1u << (__ffs(~cur) - 1);
                if (!(atomicOr(bitmap + w, bit) & bit)) got = w * 32 + (__ffs(bit) - 1);
            } else
                w = (w + 1) % n_words;  // (a lost race retries the same word)
        }
        got = __shfl_sync(0xffffffffu, got, 0);
        if (got != 0xffffffffu) return got;
    }
}
}  // namespace

template <int MIN_CTAS>
__global__ void __launch_bounds__(SPEC_WARPS * 32, MIN_CTAS)
    inflate_bgzf_spec(const uint8_t *__restrict__ file, const InflateBlock *__restrict__ blocks, uint32_t n_blocks, uint8_t *__restrict__ out,
                      uint32_t *__restrict__ error, uint32_t *slot_bitmap, uint32_t slot_words, uint32_t slots_per_sm, uint32_t *slot_mem)
{
    __shared__ SpecMem mem[SPEC_WARPS];
    const uint32_t lane = threadIdx.x & 31, full = 0xffffffffu;
    const uint32_t b = blockIdx.x * SPEC_WARPS + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    SpecMem &M = mem[threadIdx.x >> 5];
    const InflateBlock blk = blocks[b];
    uint8_t *dst = out + blk.uoff;
    const uint32_t *wbase = (const uint32_t *)((uintptr_t)(file + blk.coff) & ~(uintptr_t)3);  // bit positions count from here
    const uint32_t end = (uint32_t)((uintptr_t)(file + blk.coff) & 3) * 8 + blk.clen * 8;      // end of the deflate payload
    const uint32_t slot = slot_claim(slot_bitmap, slot_words, lane);
    uint32_t *list = slot_mem + ((size_t)(slot >> 5) * slots_per_sm + (slot & 31)) * TOK_CAP;
    BitReader br;
    br.init(file + blk.coff);  // (lane 0's copy reads the block headers)
    uint32_t pos = 0, bitpos = 0;
    bool bad = blk.ulen > 65536u;
    while (!bad) {
        uint32_t hdr = 0;
        if (lane == 0) hdr = br.bits(3);
        hdr = __shfl_sync(full, hdr, 0);
        const uint32_t final_block = hdr & 1, type = hdr >> 1;
        if (type == 0) {  // stored
            uint32_t len = 0, nlen = 0, at = 0;
            if (lane == 0) {
                br.align_byte();
                len = br.bits(16), nlen = br.bits(16);
                at = br.bit_offset(wbase);
            }
            len = __shfl_sync(full, len, 0), nlen = __shfl_sync(full, nlen, 0), at = __shfl_sync(full, at, 0);
            if ((len ^ nlen) != 0xffffu || at + len * 8 > end || pos + len > blk.ulen) {
                bad = true;
                break;
            }
            const uint8_t *src = (const uint8_t *)wbase + (at >> 3);
            for (uint32_t i = lane; i < len; i += 32) dst[pos + i] = src[i];
            pos += len;
            if (lane == 0) br.init(src + len);
            if (final_block) break;
            continue;
        }
        if (type == 3) {
            bad = true;
            break;
        }
        int n_lit = 288, n_dist = 32, ok = 1;
        if (type == 1) {  // fixed Huffman codes (RFC 1951 3.2.6); 286, 287 and 30, 31 take part in the code but are invalid
            for (int i = lane; i < 288; i += 32) M.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            M.lens[288 + lane] = 5;
        } else {  // dynamic: code lengths are themselves Huffman coded (3.2.7), decoded serially by lane 0
            if (lane == 0) {
                n_lit = (int)br.bits(5) + 257;
                n_dist = (int)br.bits(5) + 1;
                const int n_cl = (int)br.bits(4) + 4;
                uint8_t cl[19];
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < n_cl; ++i) cl[c_cl_order[i]] = (uint8_t)br.bits(3);
                for (int i = 0; i < 128; ++i) M.cl_fast[i] = 0;
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
                    const int l = cl[sy];
                    if (!l) continue;
                    const uint32_t rev = __brev(next[l]++) >> (32 - l);
                    for (uint32_t k = rev; k < 128; k += 1u << l) M.cl_fast[k] = (uint8_t)(l << 5 | sy);
                }
                int i = 0;
                const int total = n_lit + n_dist;
                if (n_lit > 286 || n_dist > 30) ok = 0;
                while (ok && i < total) {
                    const uint8_t e = M.cl_fast[br.peek(7)];
                    if (!e) {
                        ok = 0;
                        break;
                    }
                    br.consume(e >> 5);
                    const int sy = e & 31;
                    if (sy < 16) M.lens[i++] = (uint8_t)sy;
                    else {
                        int rep, v = 0;
                        if (sy == 16) {
                            if (i == 0) {
                                ok = 0;
                                break;
                            }
                            v = M.lens[i - 1];
                            rep = 3 + (int)br.bits(2);
                        } else if (sy == 17) rep = 3 + (int)br.bits(3);
                        else rep = 11 + (int)br.bits(7);
                        if (i + rep > total) {
                            ok = 0;
                            break;
                        }
                        while (rep--) M.lens[i++] = (uint8_t)v;
                    }
                }
                if (ok && M.lens[256] == 0) ok = 0;  // no end-of-block code
            }
            ok = __shfl_sync(full, ok, 0);
            n_lit = __shfl_sync(full, n_lit, 0), n_dist = __shfl_sync(full, n_dist, 0);
            if (!ok) {
                bad = true;
                break;
            }
            __syncwarp();
            const uint8_t dl = (int)lane < n_dist ? M.lens[n_lit + lane] : (uint8_t)0;  // the distance lengths follow directly
            __syncwarp();
            M.lens[288 + lane] = dl;
            for (int i = n_lit + (int)lane; i < 288; i += 32) M.lens[i] = 0;
        }
        __syncwarp();
        if (!spec_build<true>(M, M.lens, n_lit, M.lit, lane) || !spec_build<false>(M, M.lens + 288, n_dist, M.dist, lane)) {
            bad = true;
            break;
        }
        if (lane == 0) bitpos = br.bit_offset(wbase);
        bitpos = __shfl_sync(full, bitpos, 0);
        if (bitpos >= end) {
            bad = true;
            break;
        }
        // ---- speculative passes ----
        const uint32_t start = bitpos, span = max((end - start + 31u) / 32u, (uint32_t)MIN_SPAN_BITS);
        const uint32_t my_hi = min(end, start + (lane + 1) * span);
        uint32_t entry = min(end, start + lane * span);
        bool valid = lane == 0 || entry < end;
        Span sp;
        // (every lane makes every call - a lane that has nothing to decode passes an empty range - so that the warp stays
        // converged: lanes skipping a call ran ahead into the next shuffle and never rejoined the others)
        spec_decode<false>(wbase, M.lit, M.dist, entry, valid ? my_hi : 0u, sp, nullptr, M.ring + lane);
        for (;;) {
            const uint32_t p_exit = __shfl_up_sync(full, sp.exit, 1);
            const int p_status = __shfl_up_sync(full, sp.status, 1);
            const bool p_valid = __shfl_up_sync(full, (int)valid, 1) != 0;
            const bool now_valid = lane == 0 || (p_valid && p_status == SP_OK);
            const bool redo = lane != 0 && now_valid && (!valid || entry != p_exit);
            if (!__any_sync(full, redo || (valid && !now_valid))) break;
            valid = now_valid;
            if (redo) entry = p_exit;
            Span again;
            spec_decode<false>(wbase, M.lit, M.dist, entry, redo ? my_hi : 0u, again, nullptr, M.ring + lane);
            if (redo) sp = again;
        }
        // exactly one valid lane saw the end-of-block code (it invalidates its successors); a bad code on a valid lane is real
        const uint32_t endm = __ballot_sync(full, valid && sp.status == SP_END), badm = __ballot_sync(full, valid && sp.status == SP_BAD);
        if (badm || !endm) {
            bad = true;
            break;
        }
        bitpos = __shfl_sync(full, sp.exit, __ffs(endm) - 1);
        const uint32_t nb = valid ? sp.nbytes : 0u, nt = valid ? sp.ntok : 0u;
        uint32_t ib = nb, it = nt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t vb = __shfl_up_sync(full, ib, d), vt = __shfl_up_sync(full, it, d);
            if (lane >= (uint32_t)d) ib += vb, it += vt;
        }
        const uint32_t total_b = __shfl_sync(full, ib, 31);
        if (total_b > blk.ulen - pos || bitpos > end) {  // (sums of <= 2^19 bits worth of tokens cannot wrap)
            bad = true;
            break;
        }
        // ---- emission + copy, in segments of lanes whose tokens fit the list (one segment unless the block has > TOK_CAP tokens) ----
        bool fail = false;
        for (uint32_t lo = 0; lo < 32;) {
            const uint32_t t0 = lo ? __shfl_sync(full, it, lo - 1) : 0u, b0 = lo ? __shfl_sync(full, ib, lo - 1) : 0u;
            const uint32_t hi = lo + __popc(__ballot_sync(full, lane >= lo && it - t0 <= TOK_CAP));  // (it is monotone: a prefix of the lanes)
            {
                Span again;
                spec_decode<true>(wbase, M.lit, M.dist, entry, valid && lane >= lo && lane < hi ? my_hi : 0u, again, list + (it - nt - t0), M.ring + lane);
            }
            const uint32_t seg_tokens = __shfl_sync(full, it, hi - 1) - t0;
            lo = hi;
            __syncwarp();  // the token list is visible to the whole warp
            uint32_t cur = pos + b0;  // output position of the batch inside the block
            uint32_t tnext = lane < seg_tokens ? list[lane] : 0u;
            for (uint32_t base = 0; base < seg_tokens;) {
                const uint32_t t = tnext;
                const bool have = base + lane < seg_tokens, is_lit = t >> 31;
                const uint32_t n = have ? (is_lit ? 1u : (t & 511u)) : 0u, dist = (t >> 9) & 0xffffu;
                uint32_t incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t v = __shfl_up_sync(full, incl, d);
                    if (lane >= (uint32_t)d) incl += v;
                }
                const uint32_t ntake = __popc(__ballot_sync(full, have && incl <= SPAN));  // >= 1: a token is at most 258 bytes
                const uint32_t total = __shfl_sync(full, incl, ntake - 1);
                base += ntake;
                tnext = base + lane < seg_tokens ? list[base + lane] : 0u;  // (in flight while this batch is assembled)
                const bool mine = lane < ntake, is_match = mine && !is_lit;
                // Span index j <-> global byte g[j], with g 4-byte aligned: the batch starts at index al = address & 3, so that
                // whole words of the span are whole words of the output (word loads of far sources, word stores of the flush).
                uint8_t *g = dst + cur;
                const uint32_t al = (uint32_t)(uintptr_t)g & 3u;
                g -= al;
                const uint32_t o = al + incl - n;
                const int s = (int)o - (int)dist;  // source, as a span index (negative or < al: in front of the batch)
                if (is_match && (int)(cur - al) + s < 0) fail = true;
                if (mine && is_lit) M.span[o] = (uint8_t)t;
                // a source in front of the batch is complete in global memory (earlier batches are written out)
                const bool far = is_match && !fail && s + (int)n <= (int)al;
                // one lane per short match: <= 5 aligned words cover the <= 16 source bytes (reading a few bytes around the source is
                // harmless: they lie in the output buffer or its padding). The loads are issued here and used after the long matches.
                const bool far_short = far && n <= COOP_LEN;
                const uint32_t sa = (uint32_t)(uintptr_t)(g + s) & 3u;
                uint32_t w[5] = {0, 0, 0, 0, 0};
                if (far_short) {
                    const uint32_t *wp = (const uint32_t *)((uintptr_t)(g + s) & ~(uintptr_t)3);
                    const uint32_t need = sa + n;
#pragma unroll
                    for (int k = 0; k < 5; ++k)
                        if ((uint32_t)k * 4 < need) w[k] = wp[k];
                }
                uint32_t far_long = __ballot_sync(full, far && n > COOP_LEN);
                while (far_long) {  // whole warp per long match; nothing to wait for between them
                    const int tl = __ffs(far_long) - 1;
                    far_long &= far_long - 1;
                    const uint32_t o_t = __shfl_sync(full, o, tl), n_t = __shfl_sync(full, n, tl);
                    const uint8_t *src = g + __shfl_sync(full, s, tl) + lane;
                    uint8_t *to = M.span + o_t + lane;
#pragma unroll 1
                    for (uint32_t i = lane; i < n_t + lane; i += 32, src += 32, to += 32)  // (uniform trip count, predicated body)
                        if (i < n_t) *to = *src;
                }
                if (far_short) {
                    uint32_t v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = __funnelshift_r(w[k], w[k + 1], sa * 8);
#pragma unroll
                    for (uint32_t k = 0; k < COOP_LEN; ++k)
                        if (k < n) M.span[o + k] = (uint8_t)(v[k >> 2] >> (8 * (k & 3)));
                }
                uint32_t pending = __ballot_sync(full, is_match && !fail && !far);
                __syncwarp();
                while (pending) {  // matches that read this batch's output (or overlap themselves): in stream order, through shared memory
                    const int tl = __ffs(pending) - 1;
                    pending &= pending - 1;
                    const uint32_t on = __shfl_sync(full, o | n << 16, tl), d_t = __shfl_sync(full, dist, tl);
                    const uint32_t o_t = on & 0xffffu, n_t = on >> 16;
                    const int s_t = (int)o_t - (int)d_t;
                    if (s_t >= (int)al && n_t <= 32 && d_t >= n_t) {  // the usual case: source inside the span, one step (uniform branch)
                        if (lane < n_t) M.span[o_t + lane] = M.span[s_t + lane];
                    } else if (d_t >= n_t) {
#pragma unroll 1
                        for (uint32_t i0 = 0; i0 < n_t; i0 += 32) {  // (uniform trip count, predicated body)
                            const int idx = s_t + (int)(i0 + lane);
                            if (i0 + lane < n_t) M.span[o_t + i0 + lane] = idx >= (int)al ? M.span[idx] : g[idx];
                        }
                    } else {  // overlapping match: the last d_t bytes repeat
#pragma unroll 1
                        for (uint32_t i0 = 0; i0 < n_t; i0 += 32) {
                            const int idx = s_t + (int)((i0 + lane) % d_t);
                            if (i0 + lane < n_t) M.span[o_t + i0 + lane] = idx >= (int)al ? M.span[idx] : g[idx];
                        }
                    }
                    __syncwarp();
                }
                {   // write the batch out: whole words, the <= 3 bytes at either end one by one
                    const uint32_t e = al + total;
#pragma unroll 1
                    for (uint32_t m0 = 0; m0 * 4 < e; m0 += 32) {
                        const uint32_t m = m0 + lane, b0 = m * 4, b1 = b0 + 4;
                        if (b0 >= al && b1 <= e) ((uint32_t *)g)[m] = ((const uint32_t *)M.span)[m];
                        else if (b0 < e && b1 > al)
                            for (uint32_t j = max(b0, al); j < min(b1, e); ++j) g[j] = M.span[j];
                    }
                }
                cur += total;
                __syncwarp();  // the next batch reuses the span and may read what was just written
            }
        }
        if (__any_sync(full, fail)) {
            bad = true;
            break;
        }
        pos += total_b;
        if (final_block) break;
        if (lane == 0) br.init_at(wbase, bitpos);  // back to the register reader for the next block header
    }
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicAnd(slot_bitmap + (slot >> 5), ~(1u << (slot & 31)));
        if (bad || pos != blk.ulen) atomicOr(error, 1u);
    }
}


// Kernel selection: the speculative kernel by default; SEEKSV_B200_INFLATE=serial selects the one-decoding-lane form above
// (SEEKSV_B200_INFLATE_GROUP=16: its two-blocks-per-warp variant) - kept for comparison, same results (tests run both).
static int inflate_mode()  // (read per launch, so that one process can compare the kernels)
{
    const char *e = getenv("SEEKSV_B200_INFLATE");
    if (!e || strcmp(e, "serial") != 0) return 0;
    const char *g = getenv("SEEKSV_B200_INFLATE_GROUP");
    return g && atoi(g) == 16 ? 2 : 1;
}

constexpr int SPEC_MIN_CTAS = 6;  // 6 CTAs of 4 warps per SM: <= 85 registers per thread, 6 x 33 KB of shared memory
static int spec_carveout()       // share of the L1 / shared-memory array given to shared memory (percent)
{
    const char *e = getenv("SEEKSV_B200_INFLATE_CARVEOUT");
    return e ? atoi(e) : 100;
}

// the token-list slots of the speculative kernel: bitmap words (one per SM, bits >= slots-per-SM preset) + the lists
static int spec_scratch(svb_ctx *ctx)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (ctx->inflate_scratch) return 0;
    int ctas = 0;
    CK(cudaFuncSetAttribute(inflate_bgzf_spec<SPEC_MIN_CTAS>, cudaFuncAttributePreferredSharedMemoryCarveout, spec_carveout()));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, inflate_bgzf_spec<SPEC_MIN_CTAS>, SPEC_WARPS * 32, 0));
    const uint32_t per_sm = (uint32_t)std::min(32, std::max(1, ctas) * SPEC_WARPS), words = (uint32_t)ctx->sm_count;
    const size_t head = ((size_t)words * 4 + 255) & ~(size_t)255, bytes = head + (size_t)words * per_sm * TOK_CAP * sizeof(uint32_t);
    uint8_t *p = nullptr;
    if (cudaMalloc((void **)&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return svb_fail(ctx, SVB_ERR_CUDA, "cannot allocate %llu bytes of inflate scratch", (unsigned long long)bytes);
    }
    std::vector<uint32_t> init(words, per_sm >= 32 ? 0u : ~0u << per_sm);
    CK(cudaMemcpy(p, init.data(), (size_t)words * 4, cudaMemcpyHostToDevice));
    ctx->inflate_scratch = p, ctx->inflate_slot_words = words, ctx->inflate_slots_per_sm = per_sm, ctx->inflate_scratch_head = (uint32_t)head;
    return 0;
}

static int launch_inflate(svb_ctx *ctx, cudaStream_t s, const uint8_t *d_file, const void *d_blocks, uint32_t n_blocks, uint8_t *d_out,
                          uint32_t *d_err)
{
    const int mode = inflate_mode();
    if (mode == 0) {
        CKR(spec_scratch(ctx));
        const uint32_t grid = (n_blocks + SPEC_WARPS - 1) / SPEC_WARPS;
        inflate_bgzf_spec<SPEC_MIN_CTAS><<<grid, SPEC_WARPS * 32, 0, s>>>(d_file, (const InflateBlock *)d_blocks, n_blocks, d_out, d_err,
                                                                           (uint32_t *)ctx->inflate_scratch, ctx->inflate_slot_words,
                                                                           ctx->inflate_slots_per_sm,
                                                                           (uint32_t *)(ctx->inflate_scratch + ctx->inflate_scratch_head));
        return 0;
    }
    const uint32_t grid = (n_blocks + BLOCKS_PER_CTA - 1) / BLOCKS_PER_CTA;
    if (mode == 2) inflate_bgzf<16><<<grid, BLOCKS_PER_CTA * 16, 0, s>>>(d_file, (const InflateBlock *)d_blocks, n_blocks, d_out, d_err);
    else inflate_bgzf<32><<<grid, BLOCKS_PER_CTA * 32, 0, s>>>(d_file, (const InflateBlock *)d_blocks, n_blocks, d_out, d_err);
    return 0;
}

// asynchronous launch over a range of blocks; *d_err is OR-ed with 1 when a block is corrupt
int inflate_launch(svb_ctx *ctx, cudaStream_t s, const uint8_t *d_file, const void *d_blocks, uint32_t n_blocks, uint8_t *d_out, uint32_t *d_err)
{
    if (!n_blocks) return 0;
    CKR(launch_inflate(ctx, s, d_file, d_blocks, n_blocks, d_out, d_err));
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
        CKR(launch_inflate(ctx, s, d_file, d_blocks, n_blocks, d_out, err.p));
    }
    uint32_t h = 0;
    CK(cudaMemcpyAsync(&h, err.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (h) return svb_fail(ctx, SVB_ERR_FORMAT, "BGZF inflate failed (corrupt deflate stream)");
    return 0;
}
