// Host prototype of the decode tables planned for the device inflate kernel (csrc/inflate.cu), validated
// against zlib on the BGZF blocks of a real file:
//
//   g++ -O2 -std=c++17 -o huff2_proto tools/huff2_proto.cpp -lz && ./huff2_proto file.bam [n_blocks=200]
//
// Today lane 0 looks a literal/length code up in a 10-bit table and falls back to a 15-step canonical loop for longer codes
// (3.3 % of the tokens of the C2 workload), then shifts / masks / adds the extra bits of a length, then does the same for the
// distance. The tables below keep the 32-bit entry and the 10-bit / 8-bit root, and add
//   * second-level tables behind the root for longer codes (one extra lookup instead of the loop),
//   * length entries that already hold the FINAL match length when code + extra bits fit the root index (flag E_DONE),
//   * root entries that hold TWO literals when both codes fit the root index (flag E_LIT2).
// Entry layout: bits 0-4 bits to consume, bit 5 E_SUB (bits 16-31 = offset of the sub-table, bits 8-11 = its index width),
// bit 6 E_LIT, bit 7 E_END, bits 8-11 extra-bit count still to read (0 when E_DONE), bit 12 E_LIT2, bit 13 E_DONE, bit 14
// E_BAD, bits 16-31 value (literal | second literal << 8, length base or final length, distance base).
// The program decodes every sampled block with these tables (a scalar loop shaped like the kernel's), compares the output with
// zlib's and reports how many tokens each path served. Nothing in the product depends on this file.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

namespace {

constexpr uint32_t E_SUB = 1u << 5, E_LIT = 1u << 6, E_END = 1u << 7, E_LIT2 = 1u << 12, E_DONE = 1u << 13, E_BAD = 1u << 14;
constexpr int LIT_ROOT = 10, DIST_ROOT = 8;
const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

uint32_t rev(uint32_t code, int len)
{
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
}

uint32_t symbol_entry(bool lit_alphabet, int s)  // without the code length
{
    if (lit_alphabet) {
        if (s < 256) return E_LIT | (uint32_t)s << 16;
        if (s == 256) return E_END;
        if (s > 285) return E_BAD;
        return (uint32_t)kLenExtra[s - 257] << 8 | (uint32_t)kLenBase[s - 257] << 16;
    }
    if (s > 29) return E_BAD;
    return (uint32_t)kDistExtra[s] << 8 | (uint32_t)kDistBase[s] << 16;
}

// root table of `root` bits followed by the sub-tables; returns false for an over-subscribed / unusable code
bool build(const uint8_t *lens, int n, bool lit_alphabet, int root, std::vector<uint32_t> &t)
{
    int count[16] = {0};
    for (int s = 0; s < n; ++s) count[lens[s]]++;
    count[0] = 0;
    uint32_t next[16], code = 0;
    int max_len = 0;
    for (int b = 1; b <= 15; ++b) {
        code = (code + count[b - 1]) << 1;
        next[b] = code;
        if (count[b]) max_len = b;
    }
    std::vector<uint32_t> codes(n, 0);
    for (int s = 0; s < n; ++s)
        if (lens[s]) codes[s] = next[lens[s]]++;
    t.assign((size_t)1 << root, 0);
    // 1. codes that fit the root: replicate over the unused high index bits
    for (int s = 0; s < n; ++s) {
        int l = lens[s];
        if (!l || l > root) continue;
        uint32_t e = symbol_entry(lit_alphabet, s), r = rev(codes[s], l);
        int x = (e >> 8) & 15;
        for (uint32_t k = r; k < (1u << root); k += 1u << l) {
            uint32_t ent = e | (uint32_t)l;
            if (lit_alphabet && !(e & (E_LIT | E_END | E_BAD)) && l + x <= root && x > 0) {
                // the extra bits are part of the index: final length, nothing left to read
                uint32_t extra = (k >> l) & ((1u << x) - 1);
                ent = (uint32_t)(l + x) | E_DONE | ((e >> 16) + extra) << 16;
            }
            t[k] = ent;
        }
    }
    // 2. two literals in one entry when both codes fit the index
    if (lit_alphabet) {
        std::vector<uint32_t> first(t);
        for (uint32_t k = 0; k < (1u << root); ++k) {
            uint32_t e = first[k];
            if (!(e & E_LIT)) continue;
            int l = e & 31;
            uint32_t e2 = first[(k >> l) & ((1u << (root - l)) - 1)];  // the bits behind the first code, zero-extended ...
            int l2 = e2 & 31;
            if ((e2 & E_LIT) && !(e2 & E_LIT2) && l + l2 <= root)      // ... are only trusted when the second code fits entirely
                t[k] = (uint32_t)(l + l2) | E_LIT | E_LIT2 | (((e >> 16) & 0xff) | ((e2 >> 16) & 0xff) << 8) << 16;
        }
    }
    // 3. sub-tables for the codes longer than the root, one per root prefix, sized for the longest code below it
    if (max_len > root) {
        std::vector<int> width((size_t)1 << root, 0);
        for (int s = 0; s < n; ++s)
            if (lens[s] > root) {
                uint32_t p = rev(codes[s], lens[s]) & ((1u << root) - 1);
                if (lens[s] - root > width[p]) width[p] = lens[s] - root;
            }
        for (uint32_t p = 0; p < (1u << root); ++p)
            if (width[p]) {
                if (t[p]) return false;  // a short code and a long one share the prefix: not a prefix code
                uint32_t off = (uint32_t)t.size();
                if (off > 0xffff) return false;
                t[p] = (uint32_t)root | E_SUB | (uint32_t)width[p] << 8 | off << 16;
                t.resize(t.size() + ((size_t)1 << width[p]), 0);
            }
        for (int s = 0; s < n; ++s) {
            int l = lens[s];
            if (l <= root) continue;
            uint32_t r = rev(codes[s], l), p = r & ((1u << root) - 1), off = t[p] >> 16;
            int w = (t[p] >> 8) & 15;
            for (uint32_t k = r >> root; k < (1u << w); k += 1u << (l - root)) t[off + k] = symbol_entry(lit_alphabet, s) | (uint32_t)(l - root);
        }
    }
    return true;
}

struct Bits {
    const uint8_t *p;
    size_t n;
    uint64_t bp = 0;
    uint64_t window() const  // the next 57+ bits
    {
        uint64_t v = 0;
        size_t o = bp >> 3;
        for (int i = 0; i < 8 && o + i < n; ++i) v |= (uint64_t)p[o + i] << (8 * i);
        return v >> (bp & 7);
    }
    uint32_t get(int k)
    {
        uint32_t v = (uint32_t)(window() & ((1ull << k) - 1));
        bp += k;
        return v;
    }
};

struct Stats {
    uint64_t tokens = 0, lookups = 0, lit2 = 0, done = 0, sub = 0, matches = 0, dist_sub = 0, blocks = 0, table_words_max = 0;
};

bool inflate_block(const uint8_t *in, size_t n, std::vector<uint8_t> &out, Stats &st)
{
    Bits b{in, n};
    for (;;) {
        uint32_t final_block = b.get(1), type = b.get(2);
        if (type == 0) {
            b.bp = (b.bp + 7) & ~7ull;
            uint32_t len = b.get(16);
            b.get(16);
            out.insert(out.end(), in + (b.bp >> 3), in + (b.bp >> 3) + len);
            b.bp += 8ull * len;
        } else if (type == 1 || type == 2) {
            uint8_t lens[320] = {0};
            int n_lit = 288, n_dist = 30;
            if (type == 1) {
                for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                for (int i = 0; i < 30; ++i) lens[288 + i] = 5;
            } else {
                n_lit = (int)b.get(5) + 257, n_dist = (int)b.get(5) + 1;
                int n_cl = (int)b.get(4) + 4;
                uint8_t cl[19] = {0};
                for (int i = 0; i < n_cl; ++i) cl[kClOrder[i]] = (uint8_t)b.get(3);
                std::vector<uint32_t> ct;
                if (!build(cl, 19, false, 7, ct)) return false;  // (the 19-symbol alphabet reuses the builder; values unused)
                // map table entries back to symbols: rebuild a plain symbol table for the code-length alphabet
                int count[8] = {0};
                for (int i = 0; i < 19; ++i) count[cl[i]]++;
                count[0] = 0;
                uint32_t next[8], code = 0;
                for (int l = 1; l < 8; ++l) code = (code + count[l - 1]) << 1, next[l] = code;
                uint8_t sym_of[128], len_of[128];
                memset(len_of, 0, sizeof len_of);
                for (int s = 0; s < 19; ++s)
                    if (cl[s]) {
                        uint32_t r = rev(next[cl[s]]++, cl[s]);
                        for (uint32_t k = r; k < 128; k += 1u << cl[s]) sym_of[k] = (uint8_t)s, len_of[k] = cl[s];
                    }
                uint8_t all[320] = {0};
                int i = 0, total = n_lit + n_dist;
                while (i < total) {
                    uint32_t k = (uint32_t)(b.window() & 127);
                    if (!len_of[k]) return false;
                    b.bp += len_of[k];
                    int s = sym_of[k];
                    if (s < 16) all[i++] = (uint8_t)s;
                    else {
                        int rep, v = 0;
                        if (s == 16) {
                            if (!i) return false;
                            v = all[i - 1], rep = 3 + (int)b.get(2);
                        } else if (s == 17) rep = 3 + (int)b.get(3);
                        else rep = 11 + (int)b.get(7);
                        if (i + rep > total) return false;
                        while (rep--) all[i++] = (uint8_t)v;
                    }
                }
                memcpy(lens, all, n_lit);
                memcpy(lens + 288, all + n_lit, n_dist);
            }
            std::vector<uint32_t> lt, dt;
            if (!build(lens, n_lit, true, LIT_ROOT, lt) || !build(lens + 288, n_dist, false, DIST_ROOT, dt)) return false;
            if (lt.size() + dt.size() > st.table_words_max) st.table_words_max = lt.size() + dt.size();
            for (;;) {
                // one look at the stream per token: 57 bits cover a literal/length code with extra bits (<= 20) and a distance
                // code with extra bits (<= 28)
                uint64_t w = b.window();
                uint32_t e = lt[w & ((1u << LIT_ROOT) - 1)];
                ++st.lookups;
                uint32_t used = 0;
                if (e & E_SUB) {
                    used = LIT_ROOT;
                    e = lt[(e >> 16) + ((w >> LIT_ROOT) & ((1u << ((e >> 8) & 15)) - 1))];
                    ++st.sub, ++st.lookups;
                }
                if (!e || (e & E_BAD)) return false;
                used += e & 31;
                if (e & E_LIT) {
                    out.push_back((uint8_t)(e >> 16));
                    ++st.tokens;
                    if (e & E_LIT2) out.push_back((uint8_t)(e >> 24)), ++st.tokens, ++st.lit2;
                    b.bp += used;
                    continue;
                }
                if (e & E_END) {
                    b.bp += used;
                    break;
                }
                uint32_t len = e >> 16;
                if (e & E_DONE) ++st.done;
                else {
                    uint32_t x = (e >> 8) & 15;
                    len += (uint32_t)((w >> used) & ((1u << x) - 1));
                    used += x;
                }
                uint64_t w2 = w >> used;
                uint32_t d = dt[w2 & ((1u << DIST_ROOT) - 1)], used2 = 0;
                ++st.lookups;
                if (d & E_SUB) {
                    used2 = DIST_ROOT;
                    d = dt[(d >> 16) + ((w2 >> DIST_ROOT) & ((1u << ((d >> 8) & 15)) - 1))];
                    ++st.dist_sub, ++st.lookups;
                }
                if (!d || (d & E_BAD)) return false;
                used2 += d & 31;
                uint32_t x2 = (d >> 8) & 15, dist = (d >> 16) + (uint32_t)((w2 >> used2) & ((1u << x2) - 1));
                used2 += x2;
                b.bp += used + used2;
                if (dist > out.size()) return false;
                for (uint32_t k = 0; k < len; ++k) out.push_back(out[out.size() - dist]);
                ++st.tokens, ++st.matches;
            }
        } else
            return false;
        if (final_block) break;
    }
    ++st.blocks;
    return true;
}

}  // namespace

int main(int argc, char **argv)
{
    if (argc < 2) return fprintf(stderr, "usage: huff2_proto file.bam [n_blocks]\n"), 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return perror(argv[1]), 1;
    std::vector<uint8_t> raw;
    uint8_t buf[1 << 16];
    size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + k);
    fclose(f);
    std::vector<std::pair<size_t, size_t>> blocks;
    for (size_t o = 0; o + 18 <= raw.size();) {
        size_t bsize = (size_t)(raw[o + 16] | raw[o + 17] << 8) + 1;
        blocks.emplace_back(o, bsize);
        o += bsize;
    }
    size_t want = argc > 2 ? (size_t)atol(argv[2]) : 200;
    std::mt19937_64 rng(1);
    Stats st;
    size_t bad = 0;
    for (size_t i = 0; i < want && blocks.size() > 2; ++i) {
        auto [o, bsize] = blocks[1 + rng() % (blocks.size() - 2)];
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
        if (r != Z_STREAM_END || !inflate_block(padded.data(), padded.size(), out, st) || out != ref) ++bad;
    }
    printf("%zu blocks decoded, %zu differ from zlib\n", want, bad);
    printf("tokens %llu: %.2f table lookups per token; %.1f %% of the literals came in pairs; %.1f %% of the matches had their length "
           "complete in the entry; second-level lookups: %.2f %% of the tokens (literal/length), %.2f %% of the matches (distance); "
           "largest table pair %llu words\n",
           (unsigned long long)st.tokens, (double)st.lookups / st.tokens, 200.0 * st.lit2 / (double)(st.tokens - st.matches),
           100.0 * st.done / (double)st.matches, 100.0 * st.sub / (double)st.tokens, 100.0 * st.dist_sub / (double)st.matches,
           (unsigned long long)st.table_words_max);
    return bad != 0;
}
