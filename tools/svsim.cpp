// svsim - seeded synthetic-BAM generator for the seeksv_b200 benchmarks and large parity runs
// (simulated reads with planted SVs over a synthetic reference; there is no network
// and no aligner at this scale, so records are emitted ALREADY ALIGNED: split reads become soft clips,
// pairs spanning a junction become discordant mates).
//
//   svsim --out P [--genome chr21:46709983[,name:len...]] [--cov 30] [--nsv 500] [--seed N] [--threads T]
//         [--sample tumor|normal] [--virus] [--max-records N] [--readlen 150] [--level 1]
//
// writes P.bam (BGZF), P.bam.bai, P.fa (reference), P.truth.tsv (planted junctions).
// Test/bench infrastructure; not part of the product library.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

struct Rng {  // splitmix64: cheap, seedable per read pair so that the output does not depend on the thread count
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    double gauss()
    {
        double u = uni(), v = uni();
        if (u < 1e-300) u = 1e-300;
        return sqrt(-2.0 * log(u)) * cos(6.283185307179586 * v);
    }
};

struct Seg {  // piece of the reference in donor order
    int tid;
    int64_t beg, end;  // 0-based half open on the reference
    bool rev;
};

struct Contig {
    std::string name;
    int64_t len;
    std::string seq;
};

static char comp(char c)
{
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 'N';
    }
}

struct Rec {
    int32_t tid, pos;
    uint64_t order;  // tie-break for a deterministic sort
    uint64_t off;       // into the owning thread's arena
    uint32_t len;
    uint32_t arena;
};

static void put32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back(x & 0xff), v.push_back((x >> 8) & 0xff), v.push_back((x >> 16) & 0xff), v.push_back(x >> 24);
}
static int reg2bin(int beg, int end)
{
    --end;
    if (beg >> 14 == end >> 14) return 4681 + (beg >> 14);
    if (beg >> 17 == end >> 17) return 585 + (beg >> 17);
    if (beg >> 20 == end >> 20) return 73 + (beg >> 20);
    if (beg >> 23 == end >> 23) return 9 + (beg >> 23);
    if (beg >> 26 == end >> 26) return 1 + (beg >> 26);
    return 0;
}

struct Piece {
    int tid;
    int64_t rpos;
    int len;
    bool rev;
    int qb, qe;
};

struct Opts {
    std::string out = "svsim";
    std::string genome = "chr21:46709983";
    double cov = 30;
    int nsv = 500, readlen = 150, threads = 0, level = 1;
    uint64_t seed = 20261017, max_records = 0;
    bool normal = false, virus = false;
    double isize_mu = 500, isize_sd = 25;
};

struct Donor {
    std::vector<Seg> segs;
    std::vector<int64_t> start;  // donor offset of each segment
    int64_t len = 0;
    void finish()
    {
        start.clear();
        len = 0;
        for (auto &s : segs) start.push_back(len), len += s.end - s.beg;
    }
    void pieces(int64_t off, int n, std::vector<Piece> &out) const
    {
        out.clear();
        size_t i = std::upper_bound(start.begin(), start.end(), off) - start.begin() - 1;
        int q = 0;
        int64_t o = off - start[i];
        while (q < n) {
            const Seg &s = segs[i];
            int m = (int)std::min<int64_t>(n - q, (s.end - s.beg) - o);
            out.push_back(s.rev ? Piece{s.tid, s.end - o - m, m, true, q, q + m} : Piece{s.tid, s.beg + o, m, false, q, q + m});
            q += m, ++i, o = 0;
        }
    }
};

struct Mate {
    int tid, pos, reflen, lclip, rclip;
    bool rev, unmapped = false, first = false;
    std::string seq;  // reference-forward orientation as stored in the BAM
    std::vector<Piece> pcs;
    int main = 0;
    bool donor_fwd;
};

struct Gen {
    const Opts &o;
    const std::vector<Contig> &ctg;
    Gen(const Opts &o_, const std::vector<Contig> &c) : o(o_), ctg(c) {}

    void fetch(const Donor &d, int64_t off, int n, std::string &s, std::vector<Piece> &pcs) const
    {
        d.pieces(off, n, pcs);
        s.resize(n);
        for (auto &p : pcs) {
            const std::string &r = ctg[p.tid].seq;
            if (!p.rev) memcpy(&s[p.qb], &r[p.rpos], p.len);
            else
                for (int i = 0; i < p.len; ++i) s[p.qb + i] = comp(r[p.rpos + p.len - 1 - i]);
        }
    }

    static void revcomp(std::string &s)
    {
        std::reverse(s.begin(), s.end());
        for (char &c : s) c = comp(c);
    }

    void emit(std::vector<uint8_t> &arena, std::vector<Rec> &recs, uint32_t arena_id, uint64_t order, const char *name, uint32_t flag,
              int tid, int pos, int mapq, const std::vector<uint32_t> &cigar, int mtid, int mpos, int isize, const std::string &seq,
              const std::string &qual, int nm) const
    {
        uint64_t off = arena.size();
        int end = pos;
        for (uint32_t c : cigar)
            if ((c & 15) == 0 || (c & 15) == 2 || (c & 15) == 3) end += c >> 4;
        if (end == pos) end = pos + 1;
        uint32_t lq = (uint32_t)strlen(name) + 1, l = (uint32_t)seq.size();
        // aux: NM:C, AS:C, XS:C, MD:Z<reflen>
        char md[16];
        int mdl = snprintf(md, sizeof md, "%d", end - pos);
        uint32_t auxl = 4 + 4 + 4 + 3 + mdl + 1;
        uint32_t bs = 32 + lq + 4 * (uint32_t)cigar.size() + (l + 1) / 2 + l + auxl;
        put32(arena, bs);
        put32(arena, (uint32_t)tid);
        put32(arena, (uint32_t)pos);
        put32(arena, (uint32_t)(tid >= 0 ? reg2bin(pos, end) : 4680) << 16 | (uint32_t)mapq << 8 | lq);
        put32(arena, flag << 16 | (uint32_t)cigar.size());
        put32(arena, l);
        put32(arena, (uint32_t)mtid);
        put32(arena, (uint32_t)mpos);
        put32(arena, (uint32_t)isize);
        arena.insert(arena.end(), name, name + lq);
        for (uint32_t c : cigar) put32(arena, c);
        auto nib = [](char c) -> uint8_t { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 15; };
        for (uint32_t i = 0; i < l; i += 2) arena.push_back(nib(seq[i]) << 4 | (i + 1 < l ? nib(seq[i + 1]) : 0));
        for (uint32_t i = 0; i < l; ++i) arena.push_back((uint8_t)(qual[i] - 33));
        const uint8_t a1[] = {'N', 'M', 'C', (uint8_t)nm, 'A', 'S', 'C', (uint8_t)std::max(0, (int)l - 5 * nm), 'X', 'S', 'C', 0, 'M', 'D', 'Z'};
        arena.insert(arena.end(), a1, a1 + sizeof a1);
        arena.insert(arena.end(), md, md + mdl + 1);
        recs.push_back(Rec{tid, pos, order, off, (uint32_t)(arena.size() - off), arena_id});
    }

    // one read pair sampled from a donor haplotype
    void pair(const Donor &d, int hap, uint64_t k, std::vector<uint8_t> &arena, std::vector<Rec> &recs, uint32_t arena_id) const
    {
        Rng r(o.seed * 0x100000001b3ull + (uint64_t)hap * 0x9e3779b97f4a7c15ull + k * 2654435761ull + 12345);
        const int L = o.readlen;
        int isz = std::max(L + 10, (int)lround(o.isize_mu + o.isize_sd * r.gauss()));
        if (d.len <= isz + 1) return;
        int64_t s = (int64_t)r.below((uint64_t)(d.len - isz));
        char name[48];
        snprintf(name, sizeof name, "h%d_%llu", hap, (unsigned long long)k);
        Mate m[2];
        for (int e = 0; e < 2; ++e) {
            int64_t st = e == 0 ? s : s + isz - L;
            Mate &x = m[e];
            x.donor_fwd = e == 0;
            fetch(d, st, L, x.seq, x.pcs);
            x.main = 0;
            for (size_t i = 1; i < x.pcs.size(); ++i)
                if (x.pcs[i].len > x.pcs[x.main].len) x.main = (int)i;
            const Piece &p = x.pcs[x.main];
            x.tid = p.tid, x.pos = (int)p.rpos, x.reflen = p.len;
            x.lclip = p.qb, x.rclip = L - p.qe;
            if (p.rev) {
                revcomp(x.seq);
                std::swap(x.lclip, x.rclip);
            }
            x.rev = (e == 1) ^ p.rev;
        }
        bool flip = r.uni() < 0.5;
        m[flip ? 1 : 0].first = true;
        double u = r.uni();
        bool dup = u < 0.01;
        if (u >= 0.01 && u < 0.02) m[r.uni() < 0.5 ? 0 : 1].unmapped = true;
        int lowq = (u >= 0.02 && u < 0.03) ? (int)r.below(20) : -1;
        // per-mate decoration shared between the two records through `m`
        struct Out {
            std::vector<uint32_t> cigar;
            std::string seq, qual;
            int pos, reflen, nm;
        } out[2];
        for (int e = 0; e < 2; ++e) {
            Mate &x = m[e];
            Out &w = out[e];
            w.seq = x.seq;
            w.qual.assign(L, 'I');
            for (int i = 0; i < L; ++i) {
                double q = r.uni();
                if (q > 0.93) w.qual[i] = (char)(33 + 2 + r.below(39));
                else if (q > 0.80) w.qual[i] = 'H';
            }
            for (int i = 1; i < 5; ++i)
                if (r.uni() < 0.5) w.qual[L - i] = 'D';
            int lclip = x.lclip, rclip = x.rclip, n = x.reflen;
            w.pos = x.pos;
            if (lclip == 0 && rclip == 0 && r.uni() < 0.01) {  // background soft clip of random bases
                int kk = 3 + (int)r.below(28);
                if (r.uni() < 0.5) {
                    lclip = kk;
                    for (int i = 0; i < kk; ++i) w.seq[i] = "ACGT"[r.below(4)];
                    w.pos += kk;
                } else {
                    rclip = kk;
                    for (int i = 0; i < kk; ++i) w.seq[L - 1 - i] = "ACGT"[r.below(4)];
                }
                n -= kk;
            }
            w.nm = 0;
            int nerr = (int)(r.uni() < 0.25) + (int)(r.uni() < 0.05);
            for (int i = 0; i < nerr; ++i) w.seq[r.below(L)] = "ACGT"[r.below(4)], ++w.nm;
            w.reflen = n;
            double v = r.uni();
            if (n > 60 && v < 0.01) {
                int a = 10 + (int)r.below(n - 30);
                if (r.uni() < 0.5) {
                    int kk = 1 + (int)r.below(3);
                    if (lclip) w.cigar.push_back((uint32_t)lclip << 4 | 4);
                    w.cigar.push_back((uint32_t)a << 4), w.cigar.push_back((uint32_t)kk << 4 | 1), w.cigar.push_back((uint32_t)(n - a - kk) << 4);
                    w.reflen = n - kk;
                } else {
                    int kk = 1 + (int)r.below(5);
                    if (lclip) w.cigar.push_back((uint32_t)lclip << 4 | 4);
                    w.cigar.push_back((uint32_t)a << 4), w.cigar.push_back((uint32_t)kk << 4 | 2), w.cigar.push_back((uint32_t)(n - a) << 4);
                    w.reflen = n + kk;
                }
                if (rclip) w.cigar.push_back((uint32_t)rclip << 4 | 4);
            } else {
                if (lclip) w.cigar.push_back((uint32_t)lclip << 4 | 4);
                w.cigar.push_back((uint32_t)n << 4);
                if (rclip) w.cigar.push_back((uint32_t)rclip << 4 | 4);
            }
            x.pos = w.pos, x.reflen = w.reflen;
        }
        for (int e = 0; e < 2; ++e) {
            Mate &x = m[e], &y = m[1 - e];
            Out &w = out[e];
            uint32_t flag = 1 | (x.first ? 64 : 128);
            if (x.rev) flag |= 16;
            if (y.rev) flag |= 32;
            if (dup) flag |= 1024;
            int mapq = (lowq >= 0 && e == 0) ? lowq : 60;
            int tid = x.tid, pos = x.pos, mtid = y.tid, mpos = y.pos;
            std::vector<uint32_t> cigar = w.cigar;
            std::string seq = w.seq;
            if (x.unmapped) {
                flag = (flag | 4) & ~16u;
                tid = mtid, pos = mpos, mapq = 0;
                cigar.clear();
                if (x.rev) revcomp(seq);
            }
            if (y.unmapped) {
                flag = (flag | 8) & ~32u;
                mtid = tid, mpos = pos;
            }
            int isize = 0;
            if (!x.unmapped && !y.unmapped && tid == mtid) {
                int left = std::min(pos, mpos), right = std::max(pos + x.reflen, mpos + y.reflen);
                isize = right - left;
                if (pos > mpos || (pos == mpos && !x.first)) isize = -isize;
                bool fr = (!x.rev && y.rev && pos <= mpos) || (x.rev && !y.rev && mpos <= pos);
                if (fr && std::abs(isize) < 1000) flag |= 2;
            }
            uint64_t order = ((uint64_t)hap << 62) | (k << 2) | (uint64_t)e << 1;
            emit(arena, recs, arena_id, order, name, flag, tid, pos, mapq, cigar, mtid, mpos, isize, seq, w.qual, w.nm);
            // hard-clipped supplementary alignment of the second piece of a split read
            if (x.pcs.size() > 1 && !x.unmapped && r.uni() < 0.3) {
                int best = -1;
                for (size_t i = 0; i < x.pcs.size(); ++i)
                    if ((int)i != x.main && (best < 0 || x.pcs[i].len > x.pcs[best].len)) best = (int)i;
                const Piece &p = x.pcs[best];
                if (p.len >= 20) {
                    std::string full = x.seq;
                    if (x.pcs[x.main].rev) revcomp(full);  // back to donor-forward
                    std::string piece = full.substr(p.qb, p.len);
                    int hl = p.qb, hr = L - p.qe;
                    if (p.rev) {
                        revcomp(piece);
                        std::swap(hl, hr);
                    }
                    uint32_t sflag = (flag & ~(2u | 16u)) | 2048;
                    if ((!x.donor_fwd) ^ p.rev) sflag |= 16;
                    std::vector<uint32_t> cg;
                    if (hl) cg.push_back((uint32_t)hl << 4 | 5);
                    cg.push_back((uint32_t)p.len << 4);
                    if (hr) cg.push_back((uint32_t)hr << 4 | 5);
                    emit(arena, recs, arena_id, order | 1, name, sflag, p.tid, (int)p.rpos, 60, cg, mtid, mpos, 0, piece, w.qual.substr(0, p.len), 0);
                }
            }
        }
    }
};

struct Truth {
    std::string type;
    int tid1;
    int64_t p1;
    int tid2;
    int64_t p2;
};

// donor haplotype of one contig: events on a regular grid so that they never overlap
static void plant(const Opts &o, const std::vector<Contig> &ctg, int tid, int nsv, Rng &r, bool germline_only, Donor &d,
                  std::vector<Truth> &truth, std::vector<std::pair<int, Seg>> &inserts)
{
    int64_t L = ctg[tid].len;
    d.segs.clear();
    if (nsv <= 0) {
        d.segs.push_back(Seg{tid, 0, L, false});
        d.finish();
        return;
    }
    int64_t slot = L / nsv;
    int64_t cur = 0;
    struct Move {
        int64_t at;
        Seg seg;
    };
    std::vector<Move> moves;
    std::vector<Seg> base;
    for (int i = 0; i < nsv; ++i) {
        int64_t s0 = (int64_t)i * slot + slot / 4;
        int64_t maxlen = std::min<int64_t>(slot / 2, 10000);
        double u = r.uni();
        bool germ = r.uni() < 0.2;  // 20 % of the events are also in the normal sample
        int64_t len = 50 + (int64_t)r.below((uint64_t)std::max<int64_t>(1, maxlen - 50));
        bool skip = germline_only && !germ;
        if (skip) continue;
        if (u < 0.6) {  // deletion
            base.push_back(Seg{tid, cur, s0, false});
            cur = s0 + len;
            truth.push_back(Truth{"DEL", tid, s0, tid, s0 + len});
        } else if (u < 0.84) {  // inversion
            len = std::max<int64_t>(len, 200);
            base.push_back(Seg{tid, cur, s0, false});
            base.push_back(Seg{tid, s0, s0 + len, true});
            cur = s0 + len;
            truth.push_back(Truth{"INV", tid, s0, tid, s0 + len});
        } else {  // segment moved >= 1 Mb away (or to the far half of a short contig)
            len = std::max<int64_t>(len, 500);
            base.push_back(Seg{tid, cur, s0, false});
            cur = s0 + len;
            int64_t far = (s0 + L / 2) % L;
            far = (far / slot) * slot + (3 * slot) / 4 + 7;  // lands in the quiet last quarter of some slot
            if (far >= L) far = L - 1;
            moves.push_back(Move{far, Seg{tid, s0, s0 + len, false}});
            truth.push_back(Truth{"MOVE", tid, s0, tid, far});
        }
    }
    base.push_back(Seg{tid, cur, L, false});
    // splice the moved segments in at their target coordinates
    std::sort(moves.begin(), moves.end(), [](const Move &a, const Move &b) { return a.at < b.at; });
    size_t mi = 0;
    for (const Seg &s : base) {
        Seg rest = s;
        while (mi < moves.size() && !rest.rev && moves[mi].at >= rest.beg && moves[mi].at < rest.end) {
            d.segs.push_back(Seg{tid, rest.beg, moves[mi].at, false});
            d.segs.push_back(moves[mi].seg);
            rest.beg = moves[mi].at;
            ++mi;
        }
        d.segs.push_back(rest);
    }
    // externally supplied insertions (virus integrations): (after donor segment of this tid at ref pos, segment)
    for (auto &ins : inserts) {
        if (ins.first != tid) continue;
        (void)o;
    }
    std::vector<Seg> clean;
    for (auto &s : d.segs)
        if (s.end > s.beg) clean.push_back(s);
    d.segs.swap(clean);
    d.finish();
}

static bool deflate_block(const uint8_t *src, uint32_t n, int level, std::vector<uint8_t> &out)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    out.resize(18 + deflateBound(&zs, n) + 8);
    zs.next_in = (Bytef *)src, zs.avail_in = n;
    zs.next_out = out.data() + 18, zs.avail_out = (uInt)out.size() - 26;
    int rc = deflate(&zs, Z_FINISH);
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    uint32_t clen = (uint32_t)zs.total_out, bsize = clen + 26;
    const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out.data(), hdr, 16);
    out[16] = (bsize - 1) & 0xff, out[17] = (bsize - 1) >> 8;
    uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), src, n);
    uint8_t *t = out.data() + 18 + clen;
    for (int i = 0; i < 4; ++i) t[i] = (crc >> (8 * i)) & 0xff, t[4 + i] = (n >> (8 * i)) & 0xff;
    out.resize(bsize);
    return true;
}

int main(int argc, char **argv)
{
    Opts o;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--out") o.out = val();
        else if (a == "--genome") o.genome = val();
        else if (a == "--cov") o.cov = atof(val().c_str());
        else if (a == "--nsv") o.nsv = atoi(val().c_str());
        else if (a == "--seed") o.seed = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--threads") o.threads = atoi(val().c_str());
        else if (a == "--readlen") o.readlen = atoi(val().c_str());
        else if (a == "--level") o.level = atoi(val().c_str());
        else if (a == "--max-records") o.max_records = strtoull(val().c_str(), nullptr, 10);
        else if (a == "--sample") o.normal = val() == "normal";
        else if (a == "--virus") o.virus = true;
        else {
            fprintf(stderr, "svsim: unknown option %s\n", a.c_str());
            return 2;
        }
    }
    if (o.threads <= 0) o.threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<Contig> ctg;
    {
        std::string g = o.genome;
        if (o.virus) g += ",HBV:3215,HPV16:7906";
        size_t a = 0;
        while (a < g.size()) {
            size_t b = g.find(',', a);
            if (b == std::string::npos) b = g.size();
            std::string item = g.substr(a, b - a);
            size_t c = item.find(':');
            ctg.push_back(Contig{item.substr(0, c), atoll(item.c_str() + c + 1), ""});
            a = b + 1;
        }
    }
    // reference bases (same genome for tumour and normal: seeded by o.seed only)
    {
        std::vector<std::thread> th;
        for (size_t t = 0; t < ctg.size(); ++t) ctg[t].seq.resize(ctg[t].len);
        const int64_t BLK = 1 << 20;
        std::vector<std::pair<int, int64_t>> blocks;
        for (size_t t = 0; t < ctg.size(); ++t)
            for (int64_t b = 0; b < ctg[t].len; b += BLK) blocks.emplace_back((int)t, b);
        std::atomic<size_t> next(0);
        for (int t = 0; t < o.threads; ++t)
            th.emplace_back([&]() {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= blocks.size()) return;
                    Rng r(o.seed ^ (0xabcdefull + i * 7919));
                    int64_t b = blocks[i].second, e = std::min(ctg[blocks[i].first].len, b + BLK);
                    std::string &s = ctg[blocks[i].first].seq;
                    for (int64_t p = b; p < e;) {
                        uint64_t x = r.next();
                        for (int k = 0; k < 32 && p < e; ++k, ++p) s[p] = "ACGT"[(x >> (2 * k)) & 3];
                    }
                }
            });
        for (auto &t : th) t.join();
    }
    // donor haplotypes
    size_t n_human = ctg.size() - (o.virus ? 2 : 0);
    int64_t human_len = 0;
    for (size_t t = 0; t < n_human; ++t) human_len += ctg[t].len;
    std::vector<Donor> donors(ctg.size()), plain(ctg.size());
    std::vector<Truth> truth;
    std::vector<std::pair<int, Seg>> none;
    Rng pr(o.seed + 77);
    for (size_t t = 0; t < ctg.size(); ++t) {
        plain[t].segs.push_back(Seg{(int)t, 0, ctg[t].len, false});
        plain[t].finish();
        int n = t < n_human ? (int)llround((double)o.nsv * ctg[t].len / human_len) : 0;
        plant(o, ctg, (int)t, n, pr, o.normal, donors[t], truth, none);
    }
    if (o.virus && !o.normal) {  // 40 integrations: a virus fragment spliced into the donor of the first contig
        Donor &d = donors[0];
        Rng vr(o.seed + 99);
        std::vector<Seg> segs;
        int64_t step = ctg[0].len / 41, nextat = step;
        int made = 0;
        for (const Seg &s : d.segs) {
            Seg rest = s;
            while (made < 40 && !rest.rev && nextat > rest.beg + 1000 && nextat < rest.end - 1000) {
                int v = (int)(n_human + (made & 1));
                int64_t vb = (int64_t)vr.below((uint64_t)(ctg[v].len - 1200)), vl = 400 + (int64_t)vr.below(700);
                segs.push_back(Seg{rest.tid, rest.beg, nextat, false});
                segs.push_back(Seg{v, vb, vb + vl, false});
                truth.push_back(Truth{"VIRUS", 0, nextat, v, vb});
                rest.beg = nextat;
                nextat += step;
                ++made;
            }
            if (nextat <= rest.beg + 1000) nextat = rest.beg + step;
            segs.push_back(rest);
        }
        d.segs.swap(segs);
        d.finish();
    }
    // read pairs: half the coverage from the unmodified haplotype, half from the donor
    Gen gen(o, ctg);
    struct Job {
        const Donor *d;
        int hap;
        uint64_t k0, k1;
    };
    std::vector<Job> jobs;
    for (size_t t = 0; t < ctg.size(); ++t) {
        double cov = o.cov;
        if (t >= n_human) cov = o.normal ? 20 : 6000;  // virus contigs at very high depth (pileup cap probe, quirk Q12)
        for (int h = 0; h < 2; ++h) {
            const Donor &d = h == 0 ? plain[t] : donors[t];
            uint64_t n = (uint64_t)((double)d.len * (cov / 2) / (2.0 * o.readlen));
            for (uint64_t k = 0; k < n; k += 65536) jobs.push_back(Job{&d, (int)(t * 2 + h), k, std::min(n, k + 65536)});
        }
    }
    std::vector<std::vector<uint8_t>> arenas(o.threads);
    std::vector<std::vector<Rec>> recs(o.threads);
    {
        std::atomic<size_t> next(0);
        std::vector<std::thread> th;
        for (int t = 0; t < o.threads; ++t)
            th.emplace_back([&, t]() {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= jobs.size()) return;
                    for (uint64_t k = jobs[i].k0; k < jobs[i].k1; ++k) gen.pair(*jobs[i].d, jobs[i].hap, k, arenas[t], recs[t], (uint32_t)t);
                }
            });
        for (auto &t : th) t.join();
    }
    std::vector<Rec> all;
    for (auto &v : recs) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end(), [](const Rec &a, const Rec &b) {
        if (a.tid != b.tid) return (uint32_t)a.tid < (uint32_t)b.tid;
        if (a.pos != b.pos) return a.pos < b.pos;
        return a.order < b.order;
    });
    if (o.max_records && all.size() > o.max_records) all.resize(o.max_records);
    // header
    std::vector<uint8_t> hdr;
    std::string text = "@HD\tVN:1.0\tSO:coordinate\n";
    for (auto &c : ctg) text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len) + "\n";
    hdr.insert(hdr.end(), {'B', 'A', 'M', 1});
    put32(hdr, (uint32_t)text.size());
    hdr.insert(hdr.end(), text.begin(), text.end());
    put32(hdr, (uint32_t)ctg.size());
    for (auto &c : ctg) {
        put32(hdr, (uint32_t)c.name.size() + 1);
        hdr.insert(hdr.end(), c.name.begin(), c.name.end());
        hdr.push_back(0);
        put32(hdr, (uint32_t)c.len);
    }
    // uncompressed stream layout -> BGZF blocks of 0xff00 bytes
    uint64_t total = hdr.size();
    std::vector<uint64_t> uoff(all.size() + 1);
    for (size_t i = 0; i < all.size(); ++i) uoff[i] = total, total += all[i].len;
    uoff[all.size()] = total;
    const uint32_t BLK = 0xff00;
    size_t n_blk = (size_t)((total + BLK - 1) / BLK);
    std::vector<std::vector<uint8_t>> comp(n_blk);
    {
        std::atomic<size_t> next(0);
        std::vector<std::thread> th;
        for (int t = 0; t < o.threads; ++t)
            th.emplace_back([&]() {
                std::vector<uint8_t> buf(BLK);
                for (;;) {
                    size_t b = next.fetch_add(1);
                    if (b >= n_blk) return;
                    uint64_t lo = (uint64_t)b * BLK, hi = std::min<uint64_t>(total, lo + BLK);
                    uint64_t p = lo;
                    if (p < hdr.size()) {
                        uint64_t e = std::min<uint64_t>(hi, hdr.size());
                        memcpy(buf.data(), hdr.data() + p, e - p);
                        p = e;
                    }
                    if (p < hi) {
                        size_t i = std::upper_bound(uoff.begin(), uoff.end(), p) - uoff.begin() - 1;
                        while (p < hi) {
                            const Rec &r = all[i];
                            uint64_t ro = p - uoff[i], n = std::min<uint64_t>(r.len - ro, hi - p);
                            memcpy(buf.data() + (p - lo), arenas[r.arena].data() + r.off + ro, n);
                            p += n, ++i;
                        }
                    }
                    deflate_block(buf.data(), (uint32_t)(hi - lo), o.level, comp[b]);
                }
            });
        for (auto &t : th) t.join();
    }
    std::vector<uint64_t> coff(n_blk + 1, 0);
    for (size_t b = 0; b < n_blk; ++b) coff[b + 1] = coff[b] + comp[b].size();
    {
        FILE *f = fopen((o.out + ".bam").c_str(), "wb");
        if (!f) {
            fprintf(stderr, "svsim: cannot write %s.bam\n", o.out.c_str());
            return 1;
        }
        for (auto &c : comp) fwrite(c.data(), 1, c.size(), f);
        static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        fwrite(eof, 1, 28, f);
        fclose(f);
    }
    // BAI (sam/bam.h:630-700: bins + 16 kb linear index), built from the known virtual offsets
    auto voff = [&](uint64_t u) {
        uint64_t b = u / BLK;
        if (b >= n_blk) return coff[n_blk] << 16;  // EOF block
        return coff[b] << 16 | (u - b * BLK);
    };
    {
        FILE *f = fopen((o.out + ".bam.bai").c_str(), "wb");
        std::vector<uint8_t> out;
        out.insert(out.end(), {'B', 'A', 'I', 1});
        put32(out, (uint32_t)ctg.size());
        size_t i = 0;
        for (size_t t = 0; t < ctg.size(); ++t) {
            std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
            std::vector<uint64_t> lin;
            for (; i < all.size() && all[i].tid == (int)t; ++i) {
                const uint8_t *p = arenas[all[i].arena].data() + all[i].off;
                uint32_t w = p[12] | (p[13] << 8) | (p[14] << 16) | ((uint32_t)p[15] << 24);
                uint32_t bin = w >> 16;
                uint64_t vb = voff(uoff[i]), ve = voff(uoff[i + 1]);
                auto &ch = bins[bin];
                if (!ch.empty() && ch.back().second >> 16 == vb >> 16) ch.back().second = ve;
                else ch.emplace_back(vb, ve);
                // reference span from the cigar
                uint32_t lq = w & 0xff, nc = (p[16] | (p[17] << 8));
                int pos = all[i].pos, end = pos;
                for (uint32_t j = 0; j < nc; ++j) {
                    const uint8_t *c = p + 36 + lq + 4 * j;
                    uint32_t x = c[0] | (c[1] << 8) | (c[2] << 16) | ((uint32_t)c[3] << 24);
                    if ((x & 15) == 0 || (x & 15) == 2 || (x & 15) == 3) end += x >> 4;
                }
                if (end == pos) end = pos + 1;
                for (int wdw = pos >> 14; wdw <= (end - 1) >> 14; ++wdw) {
                    if ((size_t)wdw >= lin.size()) lin.resize(wdw + 1, 0);
                    if (lin[wdw] == 0) lin[wdw] = vb;
                }
            }
            put32(out, (uint32_t)bins.size());
            for (auto &kv : bins) {
                put32(out, kv.first);
                put32(out, (uint32_t)kv.second.size());
                for (auto &c : kv.second) {
                    put32(out, (uint32_t)c.first), put32(out, (uint32_t)(c.first >> 32));
                    put32(out, (uint32_t)c.second), put32(out, (uint32_t)(c.second >> 32));
                }
            }
            for (size_t k = 1; k < lin.size(); ++k)
                if (lin[k] == 0) lin[k] = lin[k - 1];  // bam_index_core fills gaps with the previous offset
            put32(out, (uint32_t)lin.size());
            for (uint64_t v : lin) put32(out, (uint32_t)v), put32(out, (uint32_t)(v >> 32));
        }
        fwrite(out.data(), 1, out.size(), f);
        fclose(f);
    }
    {
        FILE *f = fopen((o.out + ".fa").c_str(), "w");
        for (auto &c : ctg) {
            fprintf(f, ">%s\n", c.name.c_str());
            for (int64_t p = 0; p < c.len; p += 60) {
                fwrite(c.seq.data() + p, 1, (size_t)std::min<int64_t>(60, c.len - p), f);
                fputc('\n', f);
            }
        }
        fclose(f);
        f = fopen((o.out + ".truth.tsv").c_str(), "w");
        for (auto &t : truth) fprintf(f, "%s\t%s\t%lld\t%s\t%lld\n", t.type.c_str(), ctg[t.tid1].name.c_str(), (long long)t.p1 + 1, ctg[t.tid2].name.c_str(), (long long)t.p2 + 1);
        fclose(f);
    }
    fprintf(stderr, "svsim: %zu records, %llu uncompressed bytes, %zu planted events -> %s.bam\n", all.size(), (unsigned long long)total,
            truth.size(), o.out.c_str());
    return 0;
}
