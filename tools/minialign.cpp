// minialign - stand-in for the external `bwa mem ref.fa P.clip.fq.gz` realignment step at benchmark scale.
//
// The reference pipeline keeps the realign step external (README.md:30-31, example/seeksv.sh:3) and so does
// seeksv_b200; bundled bwa cannot index a 3 Gbp synthetic genome inside a benchmark run and does not exist on
// the GPU box, so large synthetic configs use this exact-seed ungapped aligner instead. Parity only needs both
// implementations to consume the SAME clip.sam.
//
//   minialign ref.fa clip.fq[.gz] > clip.sam
//
// One SAM record per FASTQ record, in input order: the best ungapped placement (>= 90 % identity) found from
// 20-mer seeds at both ends of the query on either strand; mapQ 60 when the best placement is unique, 0 when
// tied; unmapped (flag 4) when the query is shorter than 20 or nothing is found.
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

static const int K = 20;

static bool read_all_gz(const char *path, std::string &out)
{
    gzFile g = gzopen(path, "rb");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    std::vector<char> buf(1 << 22);
    int n;
    while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), n);
    gzclose(g);
    return n >= 0;
}

static inline int code(char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}
static bool kmer_of(const char *s, uint64_t &k)
{
    k = 0;
    for (int i = 0; i < K; ++i) {
        int c = code(s[i]);
        if (c < 0) return false;
        k = k << 2 | (uint64_t)c;
    }
    return true;
}
static std::string revcomp(const std::string &s)
{
    std::string r(s.rbegin(), s.rend());
    for (char &c : r) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    return r;
}

struct Hit {
    int tid;
    int64_t pos;  // 0-based placement of the (possibly reverse-complemented) query
    bool rev;
    int mism;
};

int main(int argc, char **argv)
{
    if (argc != 3) {
        fprintf(stderr, "usage: minialign ref.fa clip.fq[.gz] > clip.sam\n");
        return 2;
    }
    std::string fa, fq;
    if (!read_all_gz(argv[1], fa) || !read_all_gz(argv[2], fq)) {
        fprintf(stderr, "minialign: cannot read input\n");
        return 1;
    }
    std::vector<std::string> names, seqs;
    {
        size_t p = 0;
        while (p < fa.size()) {
            size_t nl = fa.find('\n', p);
            if (nl == std::string::npos) nl = fa.size();
            if (fa[p] == '>') {
                size_t e = p + 1;
                while (e < nl && !isspace((unsigned char)fa[e])) ++e;
                names.push_back(fa.substr(p + 1, e - p - 1));
                seqs.emplace_back();
            } else if (!seqs.empty())
                seqs.back().append(fa, p, nl - p);
            p = nl + 1;
        }
    }
    std::vector<std::string> qseq, qqual;
    {
        size_t p = 0;
        int line = 0;
        while (p < fq.size()) {
            size_t nl = fq.find('\n', p);
            if (nl == std::string::npos) nl = fq.size();
            if (line % 4 == 1) qseq.push_back(fq.substr(p, nl - p));
            if (line % 4 == 3) qqual.push_back(fq.substr(p, nl - p));
            ++line;
            p = nl + 1;
        }
    }
    // seed table: k-mer -> list of (query, offset in oriented query, strand)
    struct Seed {
        uint32_t q;
        int32_t off;
        bool rev;
    };
    std::unordered_map<uint64_t, std::vector<Seed>> seeds;
    seeds.reserve(qseq.size() * 4);
    for (size_t i = 0; i < qseq.size(); ++i) {
        const std::string &s = qseq[i];
        if ((int)s.size() < K) continue;
        std::string r = revcomp(s);
        uint64_t k;
        int last = (int)s.size() - K;
        if (kmer_of(s.data(), k)) seeds[k].push_back(Seed{(uint32_t)i, 0, false});
        if (last > 0 && kmer_of(s.data() + last, k)) seeds[k].push_back(Seed{(uint32_t)i, last, false});
        if (kmer_of(r.data(), k)) seeds[k].push_back(Seed{(uint32_t)i, 0, true});
        if (last > 0 && kmer_of(r.data() + last, k)) seeds[k].push_back(Seed{(uint32_t)i, last, true});
    }
    // one pass over the genome with a rolling k-mer; candidate placements per query
    std::vector<std::vector<Hit>> hits(qseq.size());
    std::vector<std::string> rcq(qseq.size());
    const uint64_t mask = (1ull << (2 * K)) - 1;
    for (size_t t = 0; t < seqs.size(); ++t) {
        const std::string &g = seqs[t];
        uint64_t k = 0;
        int valid = 0;
        for (size_t p = 0; p < g.size(); ++p) {
            int c = code(g[p]);
            if (c < 0) {
                valid = 0;
                continue;
            }
            k = (k << 2 | (uint64_t)c) & mask;
            if (++valid < K) continue;
            auto it = seeds.find(k);
            if (it == seeds.end()) continue;
            int64_t kpos = (int64_t)p - K + 1;
            for (const Seed &sd : it->second) {
                const std::string &fw = qseq[sd.q];
                if (sd.rev && rcq[sd.q].empty()) rcq[sd.q] = revcomp(fw);
                const std::string &q = sd.rev ? rcq[sd.q] : fw;
                int64_t start = kpos - sd.off;
                if (start < 0 || start + (int64_t)q.size() > (int64_t)g.size()) continue;
                int mism = 0, lim = (int)q.size() / 10;
                for (size_t j = 0; j < q.size() && mism <= lim; ++j) mism += q[j] != g[start + j];
                if (mism > lim) continue;
                bool dup = false;
                for (const Hit &h : hits[sd.q]) dup |= h.tid == (int)t && h.pos == start && h.rev == sd.rev;
                if (!dup) hits[sd.q].push_back(Hit{(int)t, start, sd.rev, mism});
            }
        }
    }
    std::string out;
    out.reserve(fq.size());
    for (size_t t = 0; t < names.size(); ++t) out += "@SQ\tSN:" + names[t] + "\tLN:" + std::to_string(seqs[t].size()) + "\n";
    for (size_t i = 0; i < qseq.size(); ++i) {
        const std::string &s = qseq[i], &ql = i < qqual.size() ? qqual[i] : s;
        const Hit *best = nullptr;
        int ties = 0;
        for (const Hit &h : hits[i]) {
            if (!best || h.mism < best->mism) best = &h, ties = 1;
            else if (h.mism == best->mism) ++ties;
        }
        if (!best) {
            out += s + "\t4\t*\t0\t0\t*\t*\t0\t0\t" + s + "\t" + ql + "\n";
            continue;
        }
        std::string seq = best->rev ? rcq[i] : s, qual = ql;
        if (best->rev) std::reverse(qual.begin(), qual.end());
        out += s + "\t" + (best->rev ? "16" : "0") + "\t" + names[best->tid] + "\t" + std::to_string(best->pos + 1) + "\t" +
               (ties == 1 ? "60" : "0") + "\t" + std::to_string(s.size()) + "M\t*\t0\t0\t" + seq + "\t" + qual + "\tNM:i:" +
               std::to_string(best->mism) + "\n";
    }
    fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
