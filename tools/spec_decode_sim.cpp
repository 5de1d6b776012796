// Host simulation of the speculative 32-lane Huffman decode of csrc/inflate.cu (round 2): every lane of a warp decodes its own
// 1/32 of a deflate block's bit range starting at a GUESSED bit position; Huffman streams resynchronise after a few tokens, so
// most lanes leave their sub-range at a true token boundary. Rounds: lane i+1 restarts at the exit of lane i until nothing
// changes (lane 0 starts at the true position, so by induction the fixed point is the serial decode).
//
//   g++ -O2 -std=c++17 -o spec_decode_sim tools/spec_decode_sim.cpp -lz && ./spec_decode_sim file.bam [n_blocks=300] [lanes=32]
//
// Reports: rounds until the fixed point, lock-step work (sum over rounds of the longest lane's token count) against the serial
// token count, and checks the reassembled output against zlib. Nothing in the product depends on this file.
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

namespace {
const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Code {  // canonical decoder (slow, exact)
    int count[16] = {0};
    std::vector<int> sym;
    bool build(const uint8_t *lens, int n)
    {
        memset(count, 0, sizeof count);
        for (int s = 0; s < n; ++s) count[lens[s]]++;
        count[0] = 0;
        int offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
        sym.assign(n, 0);
        for (int s = 0; s < n; ++s)
            if (lens[s]) sym[offs[lens[s]]++] = s;
        return true;
    }
    int decode(uint64_t w, int &len) const
    {
        int code = 0, first = 0, index = 0;
        for (len = 1; len <= 15; ++len) {
            code |= (int)(w & 1);
            w >>= 1;
            int c = count[len];
            if (code - c < first) return sym[index + (code - first)];
            index += c, first += c, first <<= 1, code <<= 1;
        }
        return -1;
    }
};

struct Bits {
    const uint8_t *p;
    size_t n;
    uint64_t window(uint64_t bp) const
    {
        uint64_t v = 0;
        size_t o = bp >> 3;
        for (int i = 0; i < 8 && o + i < n; ++i) v |= (uint64_t)p[o + i] << (8 * i);
        return v >> (bp & 7);
    }
};

struct Tok {
    uint32_t lit, len, dist;
};
enum { ST_OK, ST_EOB, ST_BAD };
struct Sub {
    uint64_t entry = 0, exit = 0;
    int status = ST_BAD;
    uint64_t ntok = 0, nbytes = 0;
    bool valid = false;
};

// decode tokens from `bp` until the position reaches `bound` (checked at token boundaries), an end-of-block or a bad code
Sub decode_range(const Bits &b, const Code &lc, const Code &dc, uint64_t bp, uint64_t bound, std::vector<Tok> *out)
{
    Sub r;
    r.entry = bp, r.valid = true, r.status = ST_OK;
    while (bp < bound) {
        uint64_t w = b.window(bp);
        int l;
        int s = lc.decode(w, l);
        if (s < 0 || s > 285) {
            r.status = ST_BAD;
            break;
        }
        if (s < 256) {
            bp += l;
            ++r.ntok, ++r.nbytes;
            if (out) out->push_back({(uint32_t)s, 0, 0});
            continue;
        }
        if (s == 256) {
            bp += l;
            r.status = ST_EOB;
            break;
        }
        uint32_t x = kLenExtra[s - 257], len = kLenBase[s - 257] + (uint32_t)((w >> l) & ((1u << x) - 1));
        uint64_t bp2 = bp + l + x;
        uint64_t w2 = b.window(bp2);
        int l2;
        int d = dc.decode(w2, l2);
        if (d < 0 || d > 29) {
            r.status = ST_BAD;
            break;
        }
        uint32_t x2 = kDistExtra[d], dist = kDistBase[d] + (uint32_t)((w2 >> l2) & ((1u << x2) - 1));
        bp = bp2 + l2 + x2;
        ++r.ntok, r.nbytes += len;
        if (out) out->push_back({0, len, dist});
    }
    r.exit = bp;
    return r;
}

struct CopyStats {
    uint64_t batches = 0, toks = 0, matches = 0, serial_old = 0, serial_span = 0, rounds_frontier = 0, rounds_exact = 0, longm = 0, selfov = 0;
    uint64_t dist_hist[8] = {0};  // <4, <16, <64, <256, <1024, <4096, <16384, rest
};
CopyStats g_copy;
uint32_t g_span = 1u << 20;

// token batches of the copy phase: up to 32 tokens, at most `span` output bytes; counts how many tokens would go through a
// one-at-a-time loop under different readiness rules
void copy_stats(const std::vector<Tok> &toks, uint32_t span)
{
    CopyStats &C = g_copy;
    size_t i = 0;
    uint64_t o = 0;
    while (i < toks.size()) {
        size_t j = i;
        uint64_t bytes = 0;
        while (j < toks.size() && j - i < 32) {
            uint32_t n = toks[j].len ? toks[j].len : 1;
            if (j > i && bytes + n > span) break;
            bytes += n, ++j;
        }
        // offsets
        std::vector<uint64_t> off(j - i);
        uint64_t q = o;
        for (size_t k = i; k < j; ++k) off[k - i] = q, q += toks[k].len ? toks[k].len : 1;
        ++C.batches, C.toks += j - i;
        // rule "source ends before the batch" -> everything else serial
        std::vector<char> pend(j - i, 0);
        for (size_t k = i; k < j; ++k) {
            const Tok &t = toks[k];
            if (!t.len) continue;
            ++C.matches;
            uint32_t d = t.dist;
            C.dist_hist[d < 4 ? 0 : d < 16 ? 1 : d < 64 ? 2 : d < 256 ? 3 : d < 1024 ? 4 : d < 4096 ? 5 : d < 16384 ? 6 : 7]++;
            if (t.len > 16) ++C.longm;
            if (t.dist < t.len) ++C.selfov;
            uint64_t se = off[k - i] - t.dist + t.len;
            if (se > o || t.dist < t.len) pend[k - i] = 1, ++C.serial_span;
            if (t.len > 16 || se > o) ++C.serial_old;
        }
        // frontier rounds: ready when the source ends at or before the first pending token's offset (or it is the first pending)
        {
            std::vector<char> p = pend;
            for (;;) {
                size_t f = 0;
                while (f < p.size() && !p[f]) ++f;
                if (f == p.size()) break;
                ++C.rounds_frontier;
                uint64_t F = off[f];
                p[f] = 0;
                for (size_t k = f + 1; k < p.size(); ++k)
                    if (p[k]) {
                        const Tok &t = toks[i + k];
                        if (t.dist >= t.len && off[k] - t.dist + t.len <= F) p[k] = 0;
                    }
            }
        }
        // exact rounds: ready when no still-pending earlier token's destination overlaps the source
        {
            std::vector<char> p = pend;
            for (;;) {
                bool any = false;
                std::vector<char> np = p;
                for (size_t k = 0; k < p.size(); ++k)
                    if (p[k]) {
                        any = true;
                        const Tok &t = toks[i + k];
                        uint64_t s0 = off[k] - t.dist, s1 = s0 + std::min(t.len, t.dist);
                        bool ready = true;
                        for (size_t m = 0; m < k && ready; ++m)
                            if (p[m]) {
                                uint64_t d0 = off[m], d1 = d0 + toks[i + m].len;
                                if (s0 < d1 && d0 < s1) ready = false;
                            }
                        if (ready) np[k] = 0;
                    }
                if (!any) break;
                ++C.rounds_exact;
                p = np;
            }
        }
        o = q, i = j;
    }
}

struct Totals {
    uint64_t blocks = 0, dblocks = 0, serial_tokens = 0, lock_work = 0, emit_work = 0, rounds = 0, max_rounds = 0, bad = 0;
    uint64_t hist[40] = {0};
    uint64_t redecode_lanes = 0;
};

bool inflate_spec(const uint8_t *in, size_t n, int K, std::vector<uint8_t> &out, Totals &T)
{
    Bits b{in, n};
    uint64_t bp = 0;
    const uint64_t end = 8ull * (n - 16);  // (the caller pads by 16 bytes)
    for (;;) {
        uint64_t w = b.window(bp);
        uint32_t final_block = w & 1, type = (w >> 1) & 3;
        bp += 3;
        if (type == 0) {
            bp = (bp + 7) & ~7ull;
            uint32_t len = (uint32_t)(b.window(bp) & 0xffff);
            bp += 32;
            out.insert(out.end(), in + (bp >> 3), in + (bp >> 3) + len);
            bp += 8ull * len;
        } else if (type == 1 || type == 2) {
            uint8_t lens[320] = {0};
            int n_lit = 288, n_dist = 30;
            if (type == 1) {
                for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                for (int i = 0; i < 30; ++i) lens[288 + i] = 5;
            } else {
                auto get = [&](int k) {
                    uint32_t v = (uint32_t)(b.window(bp) & ((1ull << k) - 1));
                    bp += k;
                    return v;
                };
                n_lit = (int)get(5) + 257, n_dist = (int)get(5) + 1;
                int n_cl = (int)get(4) + 4;
                uint8_t cl[19] = {0};
                for (int i = 0; i < n_cl; ++i) cl[kClOrder[i]] = (uint8_t)get(3);
                Code cc;
                cc.build(cl, 19);
                uint8_t all[320] = {0};
                int i = 0, total = n_lit + n_dist;
                while (i < total) {
                    int l;
                    int s = cc.decode(b.window(bp), l);
                    if (s < 0) return false;
                    bp += l;
                    if (s < 16) all[i++] = (uint8_t)s;
                    else {
                        int rep, v = 0;
                        if (s == 16) {
                            if (!i) return false;
                            v = all[i - 1], rep = 3 + (int)get(2);
                        } else if (s == 17) rep = 3 + (int)get(3);
                        else rep = 11 + (int)get(7);
                        if (i + rep > total) return false;
                        while (rep--) all[i++] = (uint8_t)v;
                    }
                }
                memcpy(lens, all, n_lit);
                memcpy(lens + 288, all + n_lit, n_dist);
            }
            Code lc, dc;
            lc.build(lens, n_lit), dc.build(lens + 288, n_dist);
            // ---- the speculative scheme ----
            const uint64_t start = bp, S = (end - start + K - 1) / K;
            std::vector<Sub> sub(K);
            auto bound_of = [&](int i) { return i + 1 < K ? std::min(end, start + (uint64_t)(i + 1) * S) : end; };
            uint64_t work = 0, rounds = 1, mx = 0;
            for (int i = 0; i < K; ++i) {
                uint64_t e = i ? std::min(end, start + (uint64_t)i * S) : start;
                sub[i] = decode_range(b, lc, dc, e, bound_of(i), nullptr);
                mx = std::max(mx, sub[i].ntok);
            }
            work += mx;
            for (;;) {
                bool changed = false;
                mx = 0;
                std::vector<Sub> nxt = sub;
                for (int i = 1; i < K; ++i) {
                    const Sub &p = sub[i - 1];
                    if (!p.valid || p.status != ST_OK) {  // the predecessor ended the block (or failed): nothing to do here
                        if (sub[i].valid) nxt[i] = Sub(), changed = true;
                        continue;
                    }
                    if (sub[i].valid && sub[i].entry == p.exit) continue;
                    nxt[i] = decode_range(b, lc, dc, p.exit, bound_of(i), nullptr);
                    // an exit can lie beyond the next lane's bound as well (a token straddles it) - decode_range then returns at once
                    mx = std::max(mx, nxt[i].ntok);
                    changed = true;
                    ++T.redecode_lanes;
                }
                sub = nxt;
                if (!changed) break;
                work += mx, ++rounds;
            }
            // emission pass (all lanes, true entries)
            std::vector<Tok> toks;
            mx = 0;
            int st = ST_OK;
            for (int i = 0; i < K; ++i) {
                if (!sub[i].valid) continue;
                Sub r = decode_range(b, lc, dc, sub[i].entry, bound_of(i), &toks);
                mx = std::max(mx, r.ntok);
                st = r.status;
                bp = r.exit;
                if (st != ST_OK) break;
            }
            if (st != ST_EOB) return false;
            T.serial_tokens += toks.size(), T.lock_work += work, T.emit_work += mx, T.rounds += rounds, ++T.dblocks;
            T.max_rounds = std::max(T.max_rounds, rounds);
            T.hist[std::min<uint64_t>(rounds, 39)]++;
            copy_stats(toks, g_span);
            for (const Tok &t : toks) {
                if (!t.len) out.push_back((uint8_t)t.lit);
                else {
                    if (t.dist > out.size()) return false;
                    for (uint32_t k = 0; k < t.len; ++k) out.push_back(out[out.size() - t.dist]);
                }
            }
        } else
            return false;
        if (final_block) break;
    }
    ++T.blocks;
    return true;
}
}  // namespace

int main(int argc, char **argv)
{
    if (argc < 2) return fprintf(stderr, "usage: spec_decode_sim file.bam [n_blocks] [lanes]\n"), 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return perror(argv[1]), 1;
    size_t want = argc > 2 ? (size_t)atol(argv[2]) : 300;
    int K = argc > 3 ? atoi(argv[3]) : 32;
    if (argc > 4) g_span = (uint32_t)atoi(argv[4]);
    // read at most the first 256 MiB
    std::vector<uint8_t> raw((size_t)256 << 20);
    raw.resize(fread(raw.data(), 1, raw.size(), f));
    fclose(f);
    std::vector<std::pair<size_t, size_t>> blocks;
    for (size_t o = 0; o + 18 <= raw.size();) {
        size_t bsize = (size_t)(raw[o + 16] | raw[o + 17] << 8) + 1;
        if (o + bsize > raw.size()) break;
        blocks.emplace_back(o, bsize);
        o += bsize;
    }
    std::mt19937_64 rng(1);
    Totals T;
    for (size_t i = 0; i < want && blocks.size() > 2; ++i) {
        auto [o, bsize] = blocks[want >= blocks.size() ? i % blocks.size() : rng() % (blocks.size() - 1)];
        const uint8_t *in = raw.data() + o + 18;
        size_t n = bsize - 18 - 8;
        std::vector<uint8_t> padded(in, in + n);
        padded.resize(n + 16, 0);
        std::vector<uint8_t> out, ref(1 << 16);
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        inflateInit2(&zs, -15);
        zs.next_in = const_cast<Bytef *>(in), zs.avail_in = (uInt)n, zs.next_out = ref.data(), zs.avail_out = (uInt)ref.size();
        int r = inflate(&zs, Z_FINISH);
        ref.resize(zs.total_out);
        inflateEnd(&zs);
        if (r != Z_STREAM_END || !inflate_spec(padded.data(), padded.size(), K, out, T) || out != ref) ++T.bad;
    }
    printf("%llu BGZF blocks (%llu deflate blocks), %llu differ from zlib\n", (unsigned long long)T.blocks, (unsigned long long)T.dblocks,
           (unsigned long long)T.bad);
    printf("serial tokens %llu; lock-step work: sync rounds %llu (%.2f rounds/block, max %llu) + emission %llu = %.3f of serial (ideal 3/%d = %.3f)\n",
           (unsigned long long)T.serial_tokens, (unsigned long long)T.lock_work, (double)T.rounds / T.dblocks, (unsigned long long)T.max_rounds,
           (unsigned long long)T.emit_work, (double)(T.lock_work + T.emit_work) / T.serial_tokens, K, 3.0 / K);
    printf("rounds histogram:");
    for (int i = 1; i < 40; ++i)
        if (T.hist[i]) printf(" %d:%llu", i, (unsigned long long)T.hist[i]);
    printf("\nlanes re-decoded after round 1: %.1f per block\n", (double)T.redecode_lanes / T.dblocks);
    const CopyStats &C = g_copy;
    printf("copy phase (batches of <= 32 tokens, <= %u bytes): %.1f tokens / batch, %.1f matches / batch; one-at-a-time tokens per batch: "
           "round-1 rule (len > 16 or source reaches the batch) %.2f, span rule (source reaches the batch or self-overlap) %.2f; rounds per batch: "
           "frontier %.2f, exact %.2f; matches > 16 bytes %.1f %%, self-overlapping %.1f %%\n",
           g_span, (double)C.toks / C.batches, (double)C.matches / C.batches, (double)C.serial_old / C.batches, (double)C.serial_span / C.batches,
           (double)C.rounds_frontier / C.batches, (double)C.rounds_exact / C.batches, 100.0 * C.longm / C.matches, 100.0 * C.selfov / C.matches);
    printf("distance histogram (<4 <16 <64 <256 <1024 <4096 <16384 rest):");
    for (int i = 0; i < 8; ++i) printf(" %.1f%%", 100.0 * C.dist_hist[i] / C.matches);
    printf("\n");
    return T.bad != 0;
}
