#include "junction.h"
#include "bamfile.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <list>
#include <map>
#include <sstream>
#include <thread>

namespace svb {

// ---- small helpers ---------------------------------------------------------------------------------------
bool JunctionKey::operator<(const JunctionKey &o) const
{
    // getsv.h:187-225: chromosomes, then strands, then positions
    if (int c = up_chr.compare(o.up_chr)) return c < 0;
    if (int c = down_chr.compare(o.down_chr)) return c < 0;
    if (up_strand != o.up_strand) return up_strand < o.up_strand;
    if (down_strand != o.down_strand) return down_strand < o.down_strand;
    if (up_pos != o.up_pos) return up_pos < o.up_pos;
    return down_pos < o.down_pos;
}

bool ChrRange::operator<(const ChrRange &o) const
{
    if (int c = chr.compare(o.chr)) return c < 0;
    if (begin != o.begin) return begin < o.begin;
    return end < o.end;
}

CigarVec cigar_from_text(const std::string &s)  // ChangeCigarType, getsv.cpp:433-451
{
    CigarVec v;
    int n = 0;
    for (char ch : s) {
        if (isdigit((unsigned char)ch)) n = n * 10 + (ch - '0');
        else {
            v.emplace_back(n, ch);
            n = 0;
        }
    }
    return v;
}

std::string cigar_to_text(const CigarVec &v, int left_clip, int right_clip)  // DisplayCigarVector, clip_reads.h:489-505
{
    std::string s;
    if (left_clip > 0) s += std::to_string(left_clip) + "S";
    for (auto &p : v) s += std::to_string(p.first) + p.second;
    if (right_clip > 0) s += std::to_string(right_clip) + "S";
    return s;
}

std::string reverse_complement(const std::string &s)  // GetReverseComplementSeq, clip_reads.cpp:414-466
{
    std::string r(s.rbegin(), s.rend());
    for (char &c : r) {
        switch (c) {
        case 'A': case 'a': c = 'T'; break;
        case 'T': case 't': c = 'A'; break;
        case 'C': case 'c': c = 'G'; break;
        case 'G': case 'g': c = 'C'; break;
        case 'n': c = 'N'; break;
        default: break;
        }
    }
    return r;
}

double match_rate_from_end(const std::string &a, const std::string &b)  // CompareStringEndFirst, clip_reads.cpp:194-205
{
    int la = (int)a.size(), lb = (int)b.size(), n = std::min(la, lb), m = 0;
    for (int i = 0; i < n; ++i) m += a[la - 1 - i] == b[lb - 1 - i];
    return (double)m / n;  // n == 0 -> NaN, which fails every >= test
}

double match_rate_from_begin(const std::string &a, const std::string &b)  // CompareStringBeginFirst, clip_reads.cpp:207-217
{
    int n = (int)std::min(a.size(), b.size()), m = 0;
    for (int i = 0; i < n; ++i) m += a[i] == b[i];
    return (double)m / n;
}

std::string format_double(double x)  // ostream << double, precision 6
{
    char buf[64];
    snprintf(buf, sizeof buf, "%g", x);
    return buf;
}

static std::vector<std::string> split_ws(const char *b, const char *e)
{
    std::vector<std::string> t;
    while (b < e) {
        while (b < e && isspace((unsigned char)*b)) ++b;
        const char *s = b;
        while (b < e && !isspace((unsigned char)*b)) ++b;
        if (b > s) t.emplace_back(s, b);
    }
    return t;
}

// cut [b, e) at line boundaries into roughly equal parts
static std::vector<std::pair<const char *, const char *>> line_chunks(const char *b, const char *e, int parts)
{
    std::vector<std::pair<const char *, const char *>> out;
    const char *p = b;
    for (int i = 1; i <= parts && p < e; ++i) {
        const char *q = i == parts ? e : b + (e - b) / parts * i;
        if (q < p) q = p;
        if (q < e) {
            const char *nl = (const char *)memchr(q, '\n', e - q);
            q = nl ? nl + 1 : e;
        }
        out.emplace_back(p, q);
        p = q;
    }
    return out;
}

template <typename F>
static void run_parallel(size_t n, int n_threads, F f)
{
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n) return;
            f(i);
        }
    };
    int nt = (int)std::min<size_t>((size_t)std::max(1, n_threads), n);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
}

static int hw_threads(int n) { return n > 0 ? n : (int)std::max(1u, std::thread::hardware_concurrency()); }

// isspace() of the "C" locale (what `fin >> token` skips) without the call per byte: 30 MB of clip text went through it
static inline bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// atoi() of a token without a heap copy (the token is followed by a tab or a newline inside the text, never by a digit)
static inline int token_int(std::string_view t)
{
    if (t.size() > 9) return atoi(std::string(t).c_str());  // (may overflow: leave it to the library)
    size_t i = 0;
    bool neg = false;
    if (i < t.size() && (t[i] == '-' || t[i] == '+')) neg = t[i] == '-', ++i;
    int v = 0;
    for (; i < t.size() && t[i] >= '0' && t[i] <= '9'; ++i) v = v * 10 + (t[i] - '0');
    return neg ? -v : v;
}

std::vector<ClipLine> parse_clip_text(const std::string &text, int n_threads) { return parse_clip_text(text.data(), text.size(), n_threads); }

std::vector<ClipLine> parse_clip_text(const char *text_begin, size_t text_size, int n_threads)
{
    // `fin >> chr >> pos >> orientation >> cigar >> ... >> support; getline(...)` (getsv.h:453-456): whitespace-separated
    // tokens, the rest of the line is dropped. Fields are views into `text`; line chunks are parsed in parallel.
    // One pass per chunk finds every byte below 0x21 eight bytes at a time (all white space is below 0x21; sequence, quality and
    // number characters are not): a tab closes a field, a newline closes the line, anything else - a blank, a carriage return, an
    // empty field among the first nine - sends that line through the generic `>>` rule.
    const char *b = text_begin, *e = b + text_size;
    auto chunks = line_chunks(b, e, text_size > (1u << 20) ? hw_threads(n_threads) : 1);
    std::vector<std::vector<ClipLine>> part(chunks.size());
    run_parallel(chunks.size(), hw_threads(n_threads), [&](size_t ci) {
        const char *const cb = chunks[ci].first, *const ce = chunks[ci].second;
        std::vector<ClipLine> &out = part[ci];
        out.reserve((size_t)(ce - cb) / 200 + 16);
        auto emit = [&](const std::string_view *t) {
            ClipLine c;
            c.chr = t[0], c.pos = token_int(t[1]), c.side = t[2][0], c.cigar = t[3];
            c.aligned_seq = t[4], c.aligned_qual = t[5], c.clipped_seq = t[6], c.clipped_qual = t[7];
            c.support = token_int(t[8]);
            out.push_back(c);
        };
        auto generic = [&](const char *p, const char *nl) {
            std::string_view t[9];
            int nt = 0;
            const char *q = p;
            while (q < nl && nt < 9) {
                while (q < nl && is_space(*q)) ++q;
                const char *s0 = q;
                while (q < nl && !is_space(*q)) ++q;
                if (q > s0) t[nt++] = std::string_view(s0, (size_t)(q - s0));
            }
            if (nt == 9) emit(t);
        };
        std::string_view t[9];
        int nf = 0;
        bool special = false;
        const char *line = cb, *field = cb, *p = cb;
        auto close_field = [&](const char *q) {
            if (nf < 9) {
                if (q == field) special = true;
                t[nf] = std::string_view(field, (size_t)(q - field));
            }
            ++nf, field = q + 1;
        };
        auto close_line = [&](const char *q) {
            close_field(q);
            if (special || nf < 9) generic(line, q);
            else emit(t);
            nf = 0, special = false, line = field = q + 1;
        };
        while (p < ce) {
            const char *q = ce;
            if (ce - p >= 8) {
                uint64_t x;
                memcpy(&x, p, 8);
                const uint64_t m = (x - 0x2121212121212121ull) & ~x & 0x8080808080808080ull;  // (the LOWEST flag is exact)
                if (!m) {
                    p += 8;
                    continue;
                }
                q = p + (__builtin_ctzll(m) >> 3);
            } else {
                for (q = p; q < ce && (unsigned char)*q >= 0x21; ++q) {}
                if (q == ce) break;
            }
            const char c = *q;
            if (c == '\t') close_field(q);
            else if (c == '\n') close_line(q);
            else special = true;
            p = q + 1;
        }
        if (line < ce) {  // last line without a newline
            close_field(ce);
            if (special || nf < 9) generic(line, ce);
            else emit(t);
        }
    });
    std::vector<size_t> at(part.size() + 1, 0);
    for (size_t i = 0; i < part.size(); ++i) at[i + 1] = at[i] + part[i].size();
    std::vector<ClipLine> out(at.back());
    run_parallel(part.size(), hw_threads(n_threads), [&](size_t ci) { std::copy(part[ci].begin(), part[ci].end(), out.begin() + (ptrdiff_t)at[ci]); });
    return out;
}

static inline uint32_t rd32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

bool parse_bam_alignments(AlignmentSet &set, uint64_t o)
{
    const std::vector<uint8_t> &s = set.storage;
    while (o + 36 <= s.size()) {
        uint32_t bs = rd32(&s[o]);
        if (bs < 32 || o + 4 + bs > s.size()) return false;
        Alignment a;
        a.tid = (int32_t)rd32(&s[o + 4]), a.pos = (int32_t)rd32(&s[o + 8]);
        uint32_t w = rd32(&s[o + 12]), w2 = rd32(&s[o + 16]);
        uint32_t lq = w & 0xff, nc = w2 & 0xffff;
        if (32ull + lq + 4ull * nc > bs) return false;  // the CIGAR has to lie inside the record (same test as the device walker)
        a.mapq = (w >> 8) & 0xff, a.flag = w2 >> 16;
        // bam1_qname is a C string (getsv.h:479 compares it as one): it ends at the first NUL, not at l_qname - the records
        // libbam's SAM reader builds from names of 255+ characters have no NUL inside l_qname (bounded by the record here)
        a.qname = std::string_view((const char *)&s[o + 36], strnlen((const char *)&s[o + 36], bs - 32));
        a.cigar_begin = (uint32_t)set.cigar_words.size(), a.cigar_n = nc;
        for (uint32_t j = 0; j < nc; ++j) set.cigar_words.push_back(rd32(&s[o + 36 + lq + 4 * j]));
        set.recs.push_back(a);
        o += 4 + bs;
    }
    return o == s.size();
}

bool parse_sam_alignments(AlignmentSet &set, int n_threads, std::string &err)
{
    const char *b = (const char *)set.storage.data(), *e = b + set.storage.size();
    // header: @SQ lines give the reference names (samopen(fn, "r") needs them: "fail to read the header" otherwise)
    const char *p = b;
    std::map<std::string, int> name2tid;
    while (p < e && *p == '@') {
        const char *nl = (const char *)memchr(p, '\n', e - p);
        if (!nl) nl = e;
        if (nl - p > 3 && p[1] == 'S' && p[2] == 'Q') {
            std::string line(p, nl);
            size_t a = line.find("\tSN:");
            if (a != std::string::npos) {
                size_t z = line.find('\t', a + 4);
                std::string sn = line.substr(a + 4, (z == std::string::npos ? line.size() : z) - a - 4);
                name2tid[sn] = (int)set.ref_names.size();
                set.ref_names.push_back(sn);
            }
        }
        p = nl < e ? nl + 1 : e;
    }
    if (set.ref_names.empty()) {
        err = "[main_samview] fail to read the header.";
        return false;
    }
    auto chunks = line_chunks(p, e, (e - p) > (1 << 20) ? hw_threads(n_threads) : 1);
    struct Part {
        std::vector<Alignment> recs;
        std::vector<uint32_t> cig;
        std::list<std::string> names;  // rebuilt names of 255+ characters (node addresses survive the splice below)
        bool bad = false;
    };
    std::vector<Part> part(chunks.size());
    run_parallel(chunks.size(), hw_threads(n_threads), [&](size_t ci) {
        const char *q = chunks[ci].first, *ce = chunks[ci].second;
        Part &P = part[ci];
        while (q < ce) {
            const char *nl = (const char *)memchr(q, '\n', ce - q);
            if (!nl) nl = ce;
            const char *le = nl;
            if (le > q && le[-1] == '\r') --le;
            if (le > q) {
                std::string_view f[6];
                int nf = 0;
                for (const char *a = q; nf < 6;) {
                    const char *t = (const char *)memchr(a, '\t', le - a);
                    if (!t) t = le;
                    f[nf++] = std::string_view(a, (size_t)(t - a));
                    if (t == le) break;
                    a = t + 1;
                }
                if (nf < 6) {
                    P.bad = true;
                    return;
                }
                Alignment al;
                al.qname = f[0];
                al.flag = sam_flag(f[1].data(), f[1].data() + f[1].size());
                al.tid = -1;
                if (!(f[2].size() == 1 && f[2][0] == '*')) {
                    auto it = name2tid.find(std::string(f[2]));
                    if (it != name2tid.end()) al.tid = it->second;
                }
                al.pos = (int32_t)strtol(f[3].data(), nullptr, 10) - 1;
                al.mapq = (int32_t)strtol(f[4].data(), nullptr, 10);
                al.cigar_begin = (uint32_t)P.cig.size();
                if (!(f[5].size() == 1 && f[5][0] == '*')) {
                    uint32_t num = 0;
                    for (char ch : f[5]) {
                        if (ch >= '0' && ch <= '9') num = num * 10 + (uint32_t)(ch - '0');
                        else {
                            const char *ops = "MIDNSHP=X", *o = strchr(ops, ch);
                            if (!o) {
                                P.bad = true;
                                return;
                            }
                            P.cig.push_back(num << 4 | (uint32_t)(o - ops));
                            num = 0;
                        }
                    }
                }
                else
                    al.flag |= 4;  // libbam's sam_read1: CIGAR "*" makes the record unmapped
                al.cigar_n = (uint32_t)P.cig.size() - al.cigar_begin;
                if (f[0].size() >= 255) {
                    // libbam keeps l_qname in 8 bits: the record holds the first (length + 1) & 0xff characters, no NUL, and
                    // bam1_qname() runs on into the CIGAR words (then sequence, ...) up to the first zero byte. Such a name
                    // never equals a clipped sequence again, which shifts the lock-step join of getsv.h:467-505 by one line.
                    std::string name(f[0].substr(0, (f[0].size() + 1) & 0xff));
                    bool ended = false;
                    for (uint32_t j = al.cigar_begin; j < P.cig.size() && !ended; ++j)
                        for (int k = 0; k < 4 && !ended; ++k) {
                            char c = (char)(P.cig[j] >> (8 * k));
                            if (c) name.push_back(c);
                            else ended = true;
                        }
                    if (!ended) name.push_back('\x01');  // (packed bases / qualities follow: no letter sequence either)
                    P.names.push_back(std::move(name));
                    al.qname = P.names.back();
                }
                P.recs.push_back(al);
            }
            q = nl < ce ? nl + 1 : ce;
        }
    });
    for (auto &P : part) {
        if (P.bad) {
            err = "malformed SAM line";
            return false;
        }
        uint32_t shift = (uint32_t)set.cigar_words.size();
        set.cigar_words.insert(set.cigar_words.end(), P.cig.begin(), P.cig.end());
        set.rebuilt_names.splice(set.rebuilt_names.end(), P.names);
        for (Alignment a : P.recs) {
            a.cigar_begin += shift;
            set.recs.push_back(a);
        }
    }
    return true;
}

// ---- join: clip.gz lines x realigned clipped sequences -> junctions -------------------------------------------
namespace {
struct AlignInfo {  // getsv.h:24-45
    std::string chr;
    int pos = -1, len = -1, lclip = 0, rclip = 0;
    char strand = '*', type = 'n';
    CigarVec cigar;
};

const char *kOps = "MIDNSHP=X";

AlignInfo align_info(const AlignmentSet &set, const Alignment &b)  // GetAlignInfo, getsv.cpp:25-71
{
    const std::vector<std::string> &names = set.ref_names;
    const uint32_t *cg = set.cigar_words.data() + b.cigar_begin;
    AlignInfo a;
    if (b.flag & 4) {
        a.chr = "Exogenous";
        return a;
    }
    a.type = ((b.flag & 256) || b.mapq == 0) ? 'r' : 'u';
    if (b.cigar_n) {
        uint32_t f = cg[0], l = cg[b.cigar_n - 1];
        if ((f & 15) == 4 || (f & 15) == 5) a.lclip = (int)(f >> 4);
        if ((l & 15) == 4 || (l & 15) == 5) a.rclip = (int)(l >> 4);
    }
    a.len = 0;
    for (uint32_t ci = 0; ci < b.cigar_n; ++ci) {  // GenerateCigar, clip_reads.cpp:309-329
        uint32_t c = cg[ci];
        uint32_t op = c & 15;
        if (op == 4 || op == 5) continue;
        if (op == 0 || op == 2 || op == 7 || op == 3) a.len += (int)(c >> 4);
        a.cigar.emplace_back((int)(c >> 4), kOps[op < 9 ? op : 0]);
    }
    a.strand = (b.flag & 16) ? '-' : '+';
    a.chr = (b.tid >= 0 && (size_t)b.tid < names.size()) ? names[b.tid] : std::string();
    a.pos = b.pos + 1;
    return a;
}

bool hard_clipped(const AlignmentSet &set, const Alignment &b)  // IsHardClip, clip_reads.cpp:247-257
{
    const uint32_t *cg = set.cigar_words.data() + b.cigar_begin;
    return b.cigar_n && ((cg[0] & 15) == 5 || (cg[b.cigar_n - 1] & 15) == 5);
}

SeqInfo make_seq(const std::string &s, const CigarVec &c, int lc, int rc, int sup, int uniq)
{
    SeqInfo x;
    x.seq = s, x.cigar = c, x.lclip = lc, x.rclip = rc, x.support = sup, x.uniq = uniq;
    return x;
}

JunctionKey make_key(const std::string &uc, int up, char us, const std::string &dc, int dp, char ds)
{
    JunctionKey k;
    k.up_chr = uc, k.up_pos = up, k.up_strand = us, k.down_chr = dc, k.down_pos = dp, k.down_strand = ds;
    return k;
}

// GetJunction, getsv.cpp:1705-1845: the entry one (line, alignment) pair stores - false: nothing is stored
bool junction_entry(const ClipLine &line, AlignInfo &ai, JunctionKey &key, SeqInfo &up, SeqInfo &down)
{
    int uniq;
    if (ai.type == 'u') uniq = 2;
    else if (ai.type == 'r') uniq = 1;
    else return false;  // 'n': nothing is stored (quirk Q7)
    CigarVec cig = cigar_from_text(std::string(line.cigar));
    const std::string chr(line.chr), clipped_seq(line.clipped_seq), aligned_seq(line.aligned_seq);
    const int pos = line.pos, sup = line.support;
    auto rev = [](CigarVec v) {
        std::reverse(v.begin(), v.end());
        return v;
    };
    if (ai.strand == '+') {
        if (line.side == '5') {
            key = make_key(ai.chr, ai.pos + ai.len - 1, '+', chr, pos, '+');
            up = make_seq(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq);
            down = make_seq(aligned_seq, cig, 0, 0, sup, 0);
        } else if (line.side == '3') {
            key = make_key(chr, pos, '+', ai.chr, ai.pos, '+');
            up = make_seq(aligned_seq, cig, 0, 0, sup, 0);
            down = make_seq(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq);
        } else
            return false;
    } else if (ai.strand == '-') {
        if (line.side == '5') {
            if (std::make_pair(ai.chr, ai.pos) <= std::make_pair(chr, pos)) {
                key = make_key(ai.chr, ai.pos, '-', chr, pos, '+');
                up = make_seq(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq);
                down = make_seq(aligned_seq, cig, 0, 0, sup, 0);
            } else {
                key = make_key(chr, pos, '-', ai.chr, ai.pos, '+');
                ai.cigar = rev(ai.cigar);  // the reference reverses the stored alignment in place
                up = make_seq(reverse_complement(aligned_seq), rev(cig), 0, 0, sup, 0);
                down = make_seq(reverse_complement(clipped_seq), ai.cigar, ai.rclip, ai.lclip, 0, uniq);
            }
        } else if (line.side == '3') {
            int aend = ai.pos + ai.len - 1;
            if (std::make_pair(chr, pos) <= std::make_pair(ai.chr, aend)) {
                key = make_key(chr, pos, '+', ai.chr, aend, '-');
                up = make_seq(aligned_seq, cig, 0, 0, sup, 0);
                down = make_seq(clipped_seq, ai.cigar, ai.lclip, ai.rclip, 0, uniq);
            } else {
                key = make_key(ai.chr, aend, '+', chr, pos, '-');
                ai.cigar = rev(ai.cigar);
                up = make_seq(reverse_complement(clipped_seq), ai.cigar, ai.rclip, ai.lclip, 0, uniq);
                down = make_seq(reverse_complement(aligned_seq), rev(cig), 0, 0, sup, 0);
            }
        } else
            return false;
    } else
        return false;
    return true;
}

// ... and how the map takes it: every entry of the key with the same clip-length signature accumulates (quirk Q8)
void store_junction(const JunctionKey &key, const SeqInfo &up, const SeqInfo &down, JunctionMap &jm)
{
    auto range = jm.equal_range(key);
    bool fresh = true;
    for (auto it = range.first; it != range.second; ++it) {
        JunctionInfo &o = it->second;
        // clip-length signature (getsv.cpp:1817); every matching entry accumulates (quirk Q8)
        if (o.up.rclip == down.lclip && o.down.lclip == up.rclip) {
            o.up.uniq = std::max(o.up.uniq, up.uniq);
            o.down.uniq = std::max(o.down.uniq, down.uniq);
            o.up.support += up.support;
            o.down.support += down.support;
            if (o.micro == -1) o.micro = it->first.up_pos - key.up_pos;
            fresh = false;
        }
    }
    if (fresh) {
        JunctionInfo o;
        o.up = up, o.down = down;
        jm.insert(std::make_pair(key, o));
    }
}

void add_junction(const ClipLine &line, AlignInfo &ai, JunctionMap &jm)
{
    JunctionKey key;
    SeqInfo up, down;
    if (junction_entry(line, ai, key, up, down)) store_junction(key, up, down, jm);
}
}  // namespace

// InputSoftInfoStoreBreakpoint<T>, getsv.h:423-541, with its quirks: only the first line of a run of
// equal clipped sequences is crossed with the alignments, the first alignment of a new run is filed under the
// previous run's sequence, and the trailing loop does not skip hard-clipped alignments.
void join_clips_with_alignments(const std::vector<ClipLine> &lines, const AlignmentSet &set, JunctionMap &jm)
{
    // keys are (sequence the alignment was filed under, (chromosome, position)); views into the clip text, no copies
    typedef std::pair<std::string_view, std::pair<std::string, int>> AlnKey;
    const std::vector<Alignment> &alns = set.recs;
    // the reference's std::map<key, info> of one run - a handful of entries: kept as a sorted vector (insert keeps the FIRST entry of
    // a key, iteration is in key order, exactly as the map's) so that a run costs no node allocations
    struct Found : std::vector<std::pair<AlnKey, AlignInfo>> {
        void insert(std::pair<AlnKey, AlignInfo> &&kv)
        {
            auto it = std::lower_bound(begin(), end(), kv.first, [](const std::pair<AlnKey, AlignInfo> &a, const AlnKey &b) { return a.first < b; });
            if (it != end() && !(kv.first < it->first)) return;
            std::vector<std::pair<AlnKey, AlignInfo>>::insert(it, std::move(kv));
        }
    } found;
    const ClipLine *head = nullptr;  // first line of the current run
    std::string_view current;
    size_t ai = 0;
    auto cross = [&]() {
        if (head)
            for (auto &kv : found) add_junction(*head, kv.second, jm);
    };
    for (const ClipLine &line : lines) {
        if (current.empty() || current == line.clipped_seq) {
            if (!head) head = &line;
            current = line.clipped_seq;
            continue;
        }
        while (ai < alns.size()) {
            const Alignment &b = alns[ai++];
            if (hard_clipped(set, b)) continue;
            if (current == b.qname) {
                AlignInfo info = align_info(set, b);
                found.insert(std::make_pair(AlnKey(current, std::make_pair(info.chr, info.pos)), info));
            } else {
                AlignInfo info = align_info(set, b);
                AlnKey k(current, std::make_pair(info.chr, info.pos));
                cross();
                found.clear();
                found.insert(std::make_pair(k, info));
                head = &line;
                current = line.clipped_seq;
                break;
            }
        }
        // alignment stream exhausted: this line is dropped and the state stays as it is
    }
    while (ai < alns.size()) {
        const Alignment &b = alns[ai++];
        if (current != b.qname) break;
        AlignInfo info = align_info(set, b);
        found.insert(std::make_pair(AlnKey(current, std::make_pair(info.chr, info.pos)), info));
    }
    cross();
}

// ---- device join (svb_clip_join, csrc/clipjoin.cu): what the host keeps ---------------------------------------------------------
// The arrays the device join takes: clipped sequences and read names packed into two blobs, chromosome names replaced by their
// rank in std::string::compare order (the reference keys its maps on the names; "" = tid outside the header, "Exogenous" = unmapped
// alignment, GetAlignInfo getsv.cpp:25-71).
bool pack_join_inputs(const std::vector<ClipLine> &lines, const AlignmentSet &set, JoinArrays &J, int n_threads)
{
    const std::vector<Alignment> &alns = set.recs;
    // Both passes over the lines / alignments touch one short string per element somewhere in tens of megabytes of file text: a
    // cache miss per element (10 ms single-threaded at C2 size - as long as the host join itself), so the elements are cut into
    // shares: per-share chromosome names and byte counts first, then every share fills its part of the arrays.
    const int nt = hw_threads(n_threads);
    const size_t shares = (size_t)std::max(1, std::min(nt, 64));
    auto share = [&](size_t n, size_t k) { return std::make_pair(n * k / shares, n * (k + 1) / shares); };
    std::vector<uint64_t> seq_at(shares + 1, 0), name_at(shares + 1, 0);
    std::vector<std::vector<std::string>> chr_seen(shares);
    run_parallel(shares, nt, [&](size_t k) {
        auto [l0, l1] = share(lines.size(), k);
        uint64_t sb = 0;
        std::string_view last;
        for (size_t i = l0; i < l1; ++i) {
            const ClipLine &l = lines[i];
            sb += l.clipped_seq.size();
            if (i == l0 || l.chr != last) {
                last = l.chr;
                if (std::find(chr_seen[k].begin(), chr_seen[k].end(), l.chr) == chr_seen[k].end()) chr_seen[k].push_back(std::string(l.chr));
            }
        }
        seq_at[k + 1] = sb;
        auto [a0, a1] = share(alns.size(), k);
        uint64_t nb = 0;
        for (size_t j = a0; j < a1; ++j) nb += alns[j].qname.size();
        name_at[k + 1] = nb;
    });
    for (size_t k = 0; k < shares; ++k) seq_at[k + 1] += seq_at[k], name_at[k + 1] += name_at[k];
    // names -> ranks
    std::map<std::string, int32_t> rank;
    rank[""] = 0, rank["Exogenous"] = 0;
    for (const std::string &n : set.ref_names) rank[n] = 0;
    for (const auto &seen : chr_seen)
        for (const std::string &n : seen) rank.emplace(n, 0);
    if (rank.size() >= (1u << 30)) return false;
    J.rank_names.clear();
    for (auto &kv : rank) kv.second = (int32_t)J.rank_names.size(), J.rank_names.push_back(kv.first);
    std::vector<int32_t> tid_rank(set.ref_names.size());
    for (size_t t = 0; t < set.ref_names.size(); ++t) tid_rank[t] = rank[set.ref_names[t]];
    const int32_t r_none = rank[""], r_exo = rank["Exogenous"];
    const uint64_t seq_bytes = seq_at[shares], name_bytes = name_at[shares];
    if (seq_bytes >= (1ull << 32) || name_bytes >= (1ull << 32) || set.cigar_words.size() >= (1ull << 32)) return false;
    J.lines.resize(lines.size()), J.seqs.resize(seq_bytes), J.alns.resize(alns.size()), J.names.resize(name_bytes);
    const std::map<std::string, int32_t> &ranks = rank;  // (read-only from here on: shared by the threads)
    run_parallel(shares, nt, [&](size_t k) {
        {
            auto [l0, l1] = share(lines.size(), k);
            uint64_t o = seq_at[k];
            std::string_view last;
            int32_t last_rank = 0;
            for (size_t i = l0; i < l1; ++i) {
                const ClipLine &l = lines[i];
                if (i == l0 || l.chr != last) last = l.chr, last_rank = ranks.find(std::string(l.chr))->second;
                J.lines[i] = svb_join_line{(uint32_t)o, (uint32_t)l.clipped_seq.size(), last_rank, l.pos, (uint32_t)(unsigned char)l.side};
                memcpy(&J.seqs[o], l.clipped_seq.data(), l.clipped_seq.size());
                o += l.clipped_seq.size();
            }
        }
        auto [a0, a1] = share(alns.size(), k);
        uint64_t o = name_at[k];
        for (size_t j = a0; j < a1; ++j) {
            const Alignment &a = alns[j];
            const int32_t r = (a.flag & 4) ? r_exo : (a.tid >= 0 && (size_t)a.tid < tid_rank.size()) ? tid_rank[a.tid] : r_none;
            J.alns[j] = svb_join_aln{(uint32_t)o, (uint32_t)a.qname.size(), a.flag, a.cigar_begin, a.cigar_n, r, a.pos, a.mapq};
            memcpy(&J.names[o], a.qname.data(), a.qname.size());
            o += a.qname.size();
        }
    });
    return true;
}

// The candidates of the device join, in its (stable Junction::operator<) order, enter the map exactly as GetJunction stores them:
// the order-dependent accumulation only looks at entries of the candidate's own key, and candidates of one key arrive in the order
// in which the reference's loop meets them. Every device key is checked against the host's own rule on the way.
bool accumulate_join_candidates(const std::vector<ClipLine> &lines, const AlignmentSet &set, const JoinArrays &J, const svb_join_cand *cands,
                                uint64_t n, JunctionMap &jm, std::string &err)
{
    for (uint64_t i = 0; i < n; ++i) {
        const svb_join_cand &c = cands[i];
        if (c.line >= lines.size() || c.aln >= set.recs.size()) {
            err = "clip_join: candidate out of range";
            return false;
        }
        AlignInfo ai = align_info(set, set.recs[c.aln]);
        JunctionKey key;
        SeqInfo up, down;
        if (!junction_entry(lines[c.line], ai, key, up, down)) {
            err = "clip_join: the device stored a pair the host rule drops";
            return false;
        }
        if (c.up_rank < 0 || (size_t)c.up_rank >= J.rank_names.size() || c.down_rank < 0 || (size_t)c.down_rank >= J.rank_names.size() ||
            J.rank_names[c.up_rank] != key.up_chr || J.rank_names[c.down_rank] != key.down_chr || c.up_pos != key.up_pos ||
            c.down_pos != key.down_pos || (char)c.up_strand != key.up_strand || (char)c.down_strand != key.down_strand) {
            err = "clip_join: device key differs from the host rule";
            return false;
        }
        store_junction(key, up, down, jm);
    }
    return true;
}

// MergeJunction, getsv.cpp:1325-1482
// ---- getsv -F (FindJunction, process_bwasw.cpp:5-227) -----------------------------------------------------------------------
namespace {
struct HalfRead {  // Alignment, process_bwasw.h:27-80 (the quality strings are never used)
    std::string chr, left, right;
    int pos = 0;
    CigarVec cigar;
    char side = '5', strand = '+';
};

void minus_cigar_right(CigarVec &v, int length)  // MinusCigarRight, clip_reads.cpp:507-546
{
    int total = 0;
    for (auto &e : v)
        if (e.second == 'M' || e.second == 'I') total += e.first;
    if (total <= length) return;
    int left = total - length;
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].second != 'M' && v[i].second != 'I') continue;
        if (v[i].first >= left) {
            v[i].first = left;
            v.resize(i + 1);
            return;
        }
        left -= v[i].first;
    }
}

void add_cigar_left(CigarVec &v, int length)  // AddCigarLeft, clip_reads.cpp:548-558
{
    if (!v.empty() && v[0].second == 'M') v[0].first += length;
    else v.insert(v.begin(), std::make_pair(length, 'M'));
}

SeqInfo seq_info(const std::string &seq, const CigarVec &cigar, int lclip, int rclip, int support, int uniq)
{
    SeqInfo s;
    s.seq = seq, s.cigar = cigar, s.lclip = lclip, s.rclip = rclip, s.support = support, s.uniq = uniq;
    return s;
}

JunctionKey junction_key(const std::string &uc, int up, char us, const std::string &dc, int dp, char ds)
{
    JunctionKey k;
    k.up_chr = uc, k.up_pos = up, k.up_strand = us, k.down_chr = dc, k.down_pos = dp, k.down_strand = ds;
    return k;
}
}  // namespace

void find_junctions(const uint8_t *stream, uint64_t n, uint64_t first, const std::vector<std::string> &ref_names, int min_mapq,
                    JunctionMap &jm)
{
    static const char kBases[] = "=ACMGRSVTWYHKDBN", kOps[] = "MIDNSHP=X";
    std::map<std::string, HalfRead> pending;  // the first record of a read name waits for one that fits it
    auto i32 = [&](uint64_t o) {
        int32_t v;
        memcpy(&v, stream + o, 4);
        return v;
    };
    for (uint64_t o = first; o + 36 <= n;) {
        const int32_t block = i32(o);
        if (block < 32 || o + 4 + (uint64_t)block > n) break;
        const uint64_t rec = o + 4;
        o = rec + (uint64_t)block;
        const int32_t tid = i32(rec), pos0 = i32(rec + 4), l_qseq = i32(rec + 16);
        const uint32_t bmq = (uint32_t)i32(rec + 8), fnc = (uint32_t)i32(rec + 12);
        const uint32_t l_qname = bmq & 0xff, mapq = (bmq >> 8) & 0xff, n_cigar = fnc & 0xffff, flag = fnc >> 16;
        if ((int)mapq < min_mapq || (flag & 4) || n_cigar == 0) continue;  // __g_skip_aln with g_min_mapQ = -w, then FUNMAP
        const uint64_t cig = rec + 32 + l_qname, seq = cig + 4ull * n_cigar;
        if (seq + ((uint64_t)l_qseq + 1) / 2 > o || tid < 0 || (size_t)tid >= ref_names.size()) continue;
        const uint32_t c1 = (uint32_t)i32(cig), c2 = (uint32_t)i32(cig + 4ull * (n_cigar - 1));
        const char op1 = kOps[std::min<uint32_t>(c1 & 15, 8)], op2 = kOps[std::min<uint32_t>(c2 & 15, 8)];
        if (op1 == 'H' || op2 == 'H' || (op1 == 'S' && op2 == 'S') || (op1 == 'M' && op2 == 'M') || (flag & 1024)) continue;
        HalfRead cur;
        int maplen = 0;  // GenerateCigar, clip_reads.cpp:309-329
        for (uint32_t k = 0; k < n_cigar; ++k) {
            const uint32_t c = (uint32_t)i32(cig + 4ull * k), op = c & 15, len = c >> 4;
            if (op == 4 || op == 5) continue;
            if (op == 0 || op == 2 || op == 7 || op == 3) maplen += (int)len;
            cur.cigar.push_back(std::make_pair((int)len, kOps[std::min<uint32_t>(op, 8)]));
        }
        int left_len, right_len;
        if (op1 == 'S') {
            cur.side = '5', left_len = (int)(c1 >> 4), right_len = l_qseq - left_len, cur.pos = pos0 + 1;
        } else {  // (every other shape counts as clipped on the right)
            cur.side = '3', right_len = (int)(c2 >> 4), left_len = l_qseq - right_len, cur.pos = pos0 + maplen;
        }
        cur.strand = (flag & 16) ? '-' : '+';
        cur.chr = ref_names[tid];
        auto base = [&](int i) { return (char)toupper(kBases[(stream[seq + (i >> 1)] >> ((~i & 1) << 2)) & 15]); };
        for (int i = 0; i < left_len && i < l_qseq; ++i) cur.left.push_back(base(i));                      // GetSeq, clip_reads.cpp:286-306
        for (int i = std::max(left_len, 0); i < left_len + right_len && i < l_qseq; ++i) cur.right.push_back(base(i));
        const std::string name((const char *)stream + rec + 32);
        auto it = pending.find(name);
        if (it == pending.end()) {
            pending.insert(std::make_pair(name, cur));
            continue;
        }
        const HalfRead &prev = it->second;
        const bool same_strand_other_side = prev.strand == cur.strand && prev.side != cur.side;
        const bool other_strand_same_side = prev.strand != cur.strand && prev.side == cur.side;
        if (!same_strand_other_side && !other_strand_same_side) continue;  // does not fit: nothing changes
        JunctionKey key;
        SeqInfo up_i, down_i;
        int micro;
        if (same_strand_other_side) {
            const HalfRead &up = prev.side == '5' ? cur : prev, &down = prev.side == '5' ? prev : cur;
            if (up.left.size() >= down.left.size()) {
                micro = (int)(up.left.size() - down.left.size());
                key = junction_key(up.chr, up.pos - micro, '+', down.chr, down.pos, '+');
                CigarVec c = up.cigar;
                minus_cigar_right(c, micro);
                up_i = seq_info(down.left, c, 0, 0, 0, 2), down_i = seq_info(down.right, down.cigar, 0, 0, 1, 2);
            } else {
                micro = 0;
                key = junction_key(up.chr, up.pos, '+', down.chr, down.pos, '+');
                up_i = seq_info(down.left, up.cigar, 0, (int)down.left.size() - (int)up.left.size(), 0, 2);
                down_i = seq_info(down.right, down.cigar, 0, 0, 1, 2);
            }
        } else {
            const bool prev_first = std::make_pair(prev.chr, prev.pos) < std::make_pair(cur.chr, cur.pos);
            const HalfRead &up = prev_first ? prev : cur, &down = prev_first ? cur : prev;
            if (cur.side == '5') {
                if (up.right.size() >= down.left.size()) {
                    micro = (int)(up.right.size() - down.left.size());
                    key = junction_key(up.chr, up.pos, '-', down.chr, down.pos + micro, '+');
                    CigarVec c = down.cigar;
                    add_cigar_left(c, micro);
                    up_i = seq_info(reverse_complement(up.right), up.cigar, 0, 0, 0, 2);
                    down_i = seq_info(reverse_complement(up.left), c, 0, 0, 1, 2);
                } else {
                    micro = 0;
                    key = junction_key(up.chr, up.pos, '-', down.chr, down.pos, '+');
                    up_i = seq_info(down.left, up.cigar, 0, (int)down.left.size() - (int)up.right.size(), 0, 2);
                    down_i = seq_info(down.right, down.cigar, 0, 0, 1, 2);
                }
            } else {
                if (up.left.size() >= down.right.size()) {
                    micro = (int)(up.left.size() - down.right.size());
                    key = junction_key(up.chr, up.pos - micro, '+', down.chr, down.pos, '-');
                    CigarVec c = up.cigar;
                    minus_cigar_right(c, micro);
                    up_i = seq_info(reverse_complement(down.right), c, 0, 0, 0, 2);
                    down_i = seq_info(reverse_complement(down.left), down.cigar, 0, 0, 1, 2);
                } else {
                    micro = 0;
                    key = junction_key(up.chr, up.pos, '+', down.chr, down.pos, '-');
                    up_i = seq_info(up.left, up.cigar, 0, 0, 0, 2);
                    down_i = seq_info(up.right, down.cigar, (int)down.right.size() - (int)up.left.size(), 0, 1, 2);
                }
            }
        }
        auto jit = jm.find(key);
        if (jit == jm.end()) {
            JunctionInfo info;
            info.up = up_i, info.down = down_i, info.micro = micro, info.pairs = 0;
            jm.insert(std::make_pair(key, info));
        } else if (jit->second.up.seq.size() != up_i.seq.size() || jit->second.down.seq.size() != down_i.seq.size())
            jit->second.down.support++;
        pending.erase(it);
    }
}

// ReadBreakpoint (getsv.cpp:1291-1323) reads with `fin >> token`: by whitespace-separated tokens, not by lines. A line that
// starts with '@' is dropped; after the 23rd token the rest of the line is dropped; a token that does not convert puts the
// stream into its fail state and ends the loop. The same stream operators on the same types reproduce all of that.
void read_breakpoints(const std::string &sv_text, JunctionMap &jm)
{
    std::istringstream fin(sv_text);
    std::string up_chr, down_chr, sv_type, up_cigar, down_cigar, up_seq, down_seq, rest;
    int up_pos = 0, up_reads = 0, down_pos = 0, down_reads = 0, micro = 0, pairs = 0, d1, d2, d3, d4, d5, d6;
    char up_strand = '+', down_strand = '+';
    double r1, r2;
    while (fin >> up_chr) {
        if (up_chr[0] == '@') {
            std::getline(fin, rest);
            continue;
        }
        fin >> up_pos >> up_strand >> up_reads >> down_chr >> down_pos >> down_strand >> down_reads >> micro >> pairs >> sv_type >> d1 >>
            d2 >> d3 >> d4 >> d5 >> d6 >> r1 >> r2 >> up_cigar >> down_cigar >> up_seq >> down_seq;
        std::getline(fin, rest);
        // (after a failed conversion the reference still inserts what the variables hold, then its loop ends: same here)
        JunctionKey key;
        key.up_chr = up_chr, key.down_chr = down_chr, key.up_pos = up_pos, key.down_pos = down_pos;
        key.up_strand = up_strand, key.down_strand = down_strand;
        JunctionInfo info;
        info.up.seq = up_seq, info.up.cigar = cigar_from_text(up_cigar), info.up.support = up_reads;
        info.down.seq = down_seq, info.down.cigar = cigar_from_text(down_cigar), info.down.support = down_reads;
        info.micro = micro, info.pairs = pairs;
        jm.insert(std::make_pair(key, info));
    }
}

void merge_junctions(JunctionMap &jm, int reach)
{
    auto it = jm.begin();
    while (it != jm.end()) {
        JunctionInfo &a = it->second;
        const JunctionKey &ka = it->first;
        if (a.up.rclip > 0 || a.up.lclip > 0) {
            ++it;
            continue;
        }
        auto jt = std::next(it);
        bool absorbed = false;  // `it` was folded into a later entry
        while (jt != jm.end() && ka.up_chr == jt->first.up_chr && ka.down_chr == jt->first.down_chr &&
               ka.up_strand == jt->first.up_strand && ka.down_strand == jt->first.down_strand &&
               jt->first.up_pos - ka.up_pos <= reach) {
            JunctionInfo &b = jt->second;
            const JunctionKey &kb = jt->first;
            if (!(std::abs(kb.down_pos - ka.down_pos) <= reach && b.down.lclip == 0)) {
                ++jt;
                continue;
            }
            std::string u1, d1, u2, d2;
            if (a.up.cigar.size() == 1 && b.up.cigar.size() == 1) {
                int mh = kb.up_pos - ka.up_pos;
                if ((ka.up_strand == '+' && b.up.seq.size() < (size_t)(mh + 5)) || (ka.up_strand == '-' && a.up.seq.size() < (size_t)(mh + 5))) {
                    ++jt;
                    continue;
                }
                if (ka.up_strand == '+') {
                    u1 = a.up.seq, d1 = a.down.seq;
                    u2 = b.up.seq.substr(0, b.up.seq.size() - mh);
                    d2 = b.up.seq.substr(b.up.seq.size() - mh) + b.down.seq;
                } else {
                    u1 = a.up.seq.substr(0, a.up.seq.size() - mh);
                    d1 = a.up.seq.substr(a.up.seq.size() - mh) + a.down.seq;
                    u2 = b.up.seq, d2 = b.down.seq;
                }
            } else if (a.down.cigar.size() == 1 && b.down.cigar.size() == 1) {
                int mh = std::abs(kb.down_pos - ka.down_pos);
                if ((ka.up_strand == '+' && a.down.seq.size() < (size_t)(mh + 5)) || (ka.up_strand == '-' && b.down.seq.size() < (size_t)(mh + 5))) {
                    ++jt;
                    continue;
                }
                if (ka.up_strand == '+') {
                    d1 = a.down.seq.substr(mh), d2 = b.down.seq;
                    u1 = a.up.seq + a.down.seq.substr(0, mh), u2 = b.up.seq;
                } else {
                    d1 = a.down.seq, d2 = b.down.seq.substr(mh);
                    u1 = a.up.seq, u2 = b.up.seq + b.down.seq.substr(0, mh);
                }
            }
            if (!(match_rate_from_end(u1, u2) >= 0.85 && match_rate_from_begin(d1, d2) >= 0.85)) {
                ++jt;
                continue;
            }
            a.up.uniq = std::max(a.up.uniq, b.up.uniq);
            a.down.uniq = std::max(a.down.uniq, b.down.uniq);
            if (a.micro == -1 && b.micro == -1) {
                a.up.support += b.up.support;
                a.down.support += b.down.support;
                if ((a.up.support != 0 && b.down.support != 0) || (a.down.support != 0 && b.up.support != 0))
                    a.micro = kb.up_pos - ka.up_pos;
                jt = jm.erase(jt);
            } else if (a.micro != -1 && b.micro == -1) {
                a.up.support += b.up.support;
                a.down.support += b.down.support;
                jt = jm.erase(jt);
            } else if (a.micro == -1 && b.micro != -1) {
                b.up.support += a.up.support;
                b.down.support += a.down.support;
                absorbed = true;
            } else {
                if (a.up.support > b.up.support || a.down.support == b.down.support) {
                    a.up.support += b.up.support;
                    jt = jm.erase(jt);
                } else if (a.up.support == b.up.support || a.down.support > b.down.support) {
                    a.down.support += b.down.support;
                    jt = jm.erase(jt);
                } else if (b.up.support > a.up.support && a.down.support == b.down.support) {
                    b.up.support += a.up.support;
                    absorbed = true;
                } else if (b.down.support > a.down.support && b.up.support == a.up.support) {
                    b.down.support += a.down.support;
                    absorbed = true;
                } else
                    ++jt;
            }
            if (absorbed) break;
        }
        if (absorbed) it = jm.erase(it);
        else ++it;
    }
}

// ---- depth bookkeeping ---------------------------------------------------------------------------------------------
void collect_breaks(const JunctionMap &jm, int flank, PosDepth &pos2depth, RangeDepth &range2depth, JunctionRanges &j2r)
{
    // GetBreak, getsv.cpp:752-789 (unsigned arithmetic on purpose)
    for (auto &kv : jm) {
        const JunctionKey &k = kv.first;
        pos2depth.insert(std::make_pair(std::make_pair(k.up_chr, k.up_pos), 0));
        pos2depth.insert(std::make_pair(std::make_pair(k.down_chr, k.down_pos), 0));
        int l = flank;
        if (k.up_chr == k.down_chr && k.up_strand == k.down_strand) {
            int d = std::abs(k.down_pos - 1 - k.up_pos);
            if (d < flank) l = d;
        }
        FlankRanges fr;
        fr.r[0] = ChrRange{k.up_chr, (unsigned)(k.up_pos - l + 1), (unsigned)k.up_pos};
        fr.r[1] = ChrRange{k.up_chr, (unsigned)(k.up_pos + 1), (unsigned)(k.up_pos + l)};
        fr.r[2] = ChrRange{k.down_chr, (unsigned)(k.down_pos - l), (unsigned)(k.down_pos - 1)};
        fr.r[3] = ChrRange{k.down_chr, (unsigned)k.down_pos, (unsigned)(k.down_pos + l - 1)};
        for (int i = 0; i < 4; ++i) range2depth.insert(std::make_pair(fr.r[i], 0ul));
        j2r.insert(std::make_pair(k, fr));
    }
}

void merge_ranges(const RangeDepth &range2depth, WindowMap &begin2end)
{
    // MergeOverlap, getsv.cpp:804-835
    if (range2depth.empty()) return;  // (the reference stores one uninitialised window here; it is never matched)
    std::string chr;
    unsigned begin = 0, end = 0;
    bool first = true;
    for (auto &kv : range2depth) {
        const ChrRange &r = kv.first;
        if (first) {
            chr = r.chr, begin = r.begin, end = r.end, first = false;
        } else if (chr == r.chr && begin <= r.begin && end + 1 >= r.begin) {
            if (r.end > end) end = r.end;
        } else {
            begin2end.insert(std::make_pair(std::make_pair(chr, (int)begin), (int)end));
            chr = r.chr, begin = r.begin, end = r.end;
        }
    }
    begin2end.insert(std::make_pair(std::make_pair(chr, (int)begin), (int)end));
}

void account_position(const std::string &chr, int p, int depth, const WindowMap &begin2end, PosDepth &pos2depth,
                      RangeDepth &range2depth)
{
    // bam2depth.cpp:82-124 for pileup position pos = p - 1
    auto w = begin2end.upper_bound(std::make_pair(chr, p));
    if (w == begin2end.begin()) return;
    --w;
    if (w->first.first != chr || p > w->second) return;
    auto r = range2depth.upper_bound(ChrRange{chr, (unsigned)(p + 1), (unsigned)(p + 1)});
    if (r == range2depth.begin()) return;  // bam2depth.cpp:102: this `continue` also skips the point depth below
    --r;
    while (r != range2depth.begin()) {
        if (r->first.chr != w->first.first || r->first.begin < (unsigned)w->first.second) break;
        if ((unsigned)p <= r->first.end) r->second += depth;
        --r;
    }
    if (r == range2depth.begin()) {
        if (r->first.chr == w->first.first && r->first.begin >= (unsigned)w->first.second && (unsigned)p <= r->first.end)
            r->second += depth;
    }
    auto q = pos2depth.find(std::make_pair(chr, p));
    if (q != pos2depth.end()) q->second = depth;
}

// ---- output ------------------------------------------------------------------------------------------------------------
const char *kSvHeader =
    "@left_chr\tleft_pos\tleft_strand\tleft_clip_read_NO\tright_chr\tright_pos\tright_strand\tright_clip_read_NO\t"
    "microhomology_length\tabnormal_readpair_NO\tsvtype\tleft_pos_depth\tright_pos_depth\taverage_depth_of_left_pos_5end\t"
    "average_depth_of_left_pos_3end\taverage_depth_of_right_pos_5end\taverage_depth_of_right_pos_3end\t"
    "left_pos_clip_percentage\tright_pos_clip_percentage\tleft_seq_cigar\tright_seq_cigar\tleft_seq\tright_seq\n";

static const char *sv_type(const JunctionKey &k)  // GetSVType, clip_reads.cpp:572-581
{
    if (k.up_chr != k.down_chr) return "CTX";
    if (k.up_strand != k.down_strand) return "INV";
    if (k.up_pos < k.down_pos) return "DEL";
    if (k.up_pos > k.down_pos) return "INS";
    return "Unknown";
}

static double top_base_fraction(const std::string &s)  // CountLargestBaseFrequency, getsv.cpp:1485-1511
{
    int n[5] = {0, 0, 0, 0, 0};
    for (char c : s) {
        switch (c) {
        case 'A': case 'a': ++n[0]; break;
        case 'T': case 't': ++n[1]; break;
        case 'C': case 'c': ++n[2]; break;
        case 'G': case 'g': ++n[3]; break;
        default: ++n[4];
        }
    }
    return *std::max_element(n, n + 5) / (double)(int)s.size();
}

void write_breakpoints(const JunctionMap &jm, const PosDepth &pos2depth, const RangeDepth &range2depth, const JunctionRanges &j2r,
                       const OutputFilters &f, std::string &body, std::string &filtered, std::string &log)
{
    // OutputBreakpoint, getsv.cpp:838-987
    for (auto &kv : jm) {
        const JunctionKey &k = kv.first;
        const JunctionInfo &o = kv.second;
        int updepth = 0, downdepth = 0;
        auto q = pos2depth.find(std::make_pair(k.up_chr, k.up_pos));
        if (q == pos2depth.end()) log += "Error: There is something wrong in upstream position " + k.up_chr + ":" + std::to_string(k.up_pos) + "\n";
        else updepth = q->second + o.down.support;
        q = pos2depth.find(std::make_pair(k.down_chr, k.down_pos));
        if (q == pos2depth.end()) log += "Error: There is something wrong in downstream position " + k.down_chr + ":" + std::to_string(k.down_pos) + "\n";
        else downdepth = q->second + o.up.support;
        int reads = o.up.support + o.down.support;
        double rate1 = updepth == 0 ? 0 : (double)reads / updepth, rate2 = downdepth == 0 ? 0 : (double)reads / downdepth;
        std::string head = k.up_chr + "\t" + std::to_string(k.up_pos) + "\t" + k.up_strand + "\t" + std::to_string(o.up.support) + "\t" +
                           k.down_chr + "\t" + std::to_string(k.down_pos) + "\t" + k.down_strand + "\t" + std::to_string(o.down.support) +
                           "\t" + std::to_string(o.micro) + "\t" + std::to_string(o.pairs) + "\t" + sv_type(k) + "\t" +
                           std::to_string(updepth) + "\t" + std::to_string(downdepth) + "\t";
        std::string tail = format_double(rate1) + "\t" + format_double(rate2) + "\t" + cigar_to_text(o.up.cigar, o.up.lclip, o.up.rclip) + "\t" +
                           cigar_to_text(o.down.cigar, o.down.lclip, o.down.rclip) + "\t" + o.up.seq + "\t" + o.down.seq + "\n";
        const char *why = nullptr;
        if (!(o.up.uniq + o.down.uniq >= 2 || o.pairs > 0)) why = "mappingQ_too_low";
        else if (k.up_chr == k.down_chr && std::abs(k.up_pos - k.down_pos) < f.min_distance) why = "distance_too_near";
        else if (o.micro > f.max_micro) why = "microhomology_len_too_long";
        else if (o.pairs < f.min_pairs) why = "abnormal_read_pair_no_not_pass";
        else if ((o.up.support > 0 && o.down.support > 0 && rate1 < f.frequency && rate2 < f.frequency) ||
                 (o.up.support == 0 && rate2 < f.frequency) || (o.down.support == 0 && rate1 < f.frequency))
            why = "frequency_too_low";
        else if (o.up.support + o.down.support < f.min_clip_sum) why = "total_clipped_reads_NO_not_pass";
        else if (o.pairs == 0) {
            if (o.up.seq.length() < (size_t)(o.up.lclip + o.up.rclip + f.min_seq_len) ||
                o.down.seq.length() < (size_t)(o.down.lclip + o.down.rclip + f.min_seq_len))
                why = "seq_length_too_short";
            else if (o.up.cigar.size() > (size_t)(2 * f.max_indel + 1) || o.down.cigar.size() > (size_t)(2 * f.max_indel + 1))
                why = "seq_with_too_many_indels";
            else if (top_base_fraction(o.up.seq) >= 0.8 || top_base_fraction(o.down.seq) >= 0.8)
                why = "repeat_bases";
        }
        if (why) {
            filtered += std::string(why) + "\t" + head + tail;
            continue;
        }
        unsigned avg[4] = {0, 0, 0, 0};
        auto jr = j2r.find(k);
        const std::string where = k.up_chr + "\t" + std::to_string(k.up_pos) + "\t" + k.up_strand + "\t" + k.down_chr + "\t" +
                                  std::to_string(k.down_pos) + "\t" + k.down_strand + "\n";  // getsv.cpp:942
        if (jr != j2r.end()) {
            for (int i = 0; i < 4; ++i) {
                auto rd = range2depth.find(jr->second.r[i]);
                if (rd == range2depth.end()) log += "Error depth in the vicinity of junction " + where;
                else {
                    unsigned span = rd->first.end - rd->first.begin + 1;
                    avg[i] = span ? (unsigned)(rd->second / span) : 0;  // (span 0 would be a division by zero in the reference)
                }
            }
        } else
            log += "Error depth in the vicinity of junction " + where;
        body += head;
        for (int i = 0; i < 4; ++i) body += std::to_string((int)avg[i]) + "\t";
        body += tail;
    }
}

// ---- somatic ------------------------------------------------------------------------------------------------------------------
namespace {
struct NormalClip {
    std::string left, right;  // seq_left / seq_right of ReadsInfo
    int support;
};
typedef std::multimap<std::pair<std::string, int>, NormalClip> ClipTable;

// Compare, clip_reads.cpp:333-372: seq2 = 3'-clipped part, seq4 = 3'-aligned part
int shifted_compare(const std::string &s1, const std::string &s2, const std::string &s3, const std::string &s4, double rate)
{
    if (s2.length() < 10) return -1;
    size_t pos = s4.find(s2.substr(0, 10));
    if (pos == std::string::npos) return -1;
    std::string s5 = s3 + s4.substr(0, pos), s6 = s4.substr(pos);
    if (match_rate_from_end(s1, s5) >= rate && match_rate_from_begin(s2, s6) >= rate) return (int)pos;
    return -1;
}

int first_match(const ClipTable &t, const std::string &chr, int pos, const std::string &begin_seq, const std::string &end_seq, double rate)
{
    auto r = t.equal_range(std::make_pair(chr, pos));
    for (auto it = r.first; it != r.second; ++it)
        if (match_rate_from_begin(begin_seq, it->second.right) >= rate && match_rate_from_end(end_seq, it->second.left) >= rate)
            return it->second.support;
    return 0;
}
}  // namespace

void somatic_rows(const std::string &normal_clip_text, const std::string &tumor_sv_text, double rate, int offset, int min_len,
                  int mean_insert, std::vector<SomaticRow> &rows, std::string &log)
{
    // ReadsClipReads<T>, somatic.h:40-70
    ClipTable t3, t5;
    for (const ClipLine &c : parse_clip_text(normal_clip_text)) {
        if (c.clipped_seq.length() < (size_t)min_len) continue;
        std::string chr(c.chr), aseq(c.aligned_seq), cseq(c.clipped_seq);
        if (c.side == '3') t3.insert(std::make_pair(std::make_pair(chr, c.pos), NormalClip{aseq, cseq, c.support}));
        else if (c.side == '5') t5.insert(std::make_pair(std::make_pair(chr, c.pos), NormalClip{cseq, aseq, c.support}));
        else log += "Error:The orientation of soft-clipped reads must be 3 or 5 in position " + chr + ":" + std::to_string(c.pos) + "\n";
    }
    auto window_first = [&](const ClipTable &t, const std::string &chr, int lo, int hi, auto &&pred) {
        for (auto it = t.lower_bound(std::make_pair(chr, lo)); it != t.end() && it->first.first == chr && it->first.second <= hi; ++it)
            if (pred(it->second)) return it->second.support;
        return 0;
    };
    const char *p = tumor_sv_text.data(), *e = p + tumor_sv_text.size();
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', e - p);
        if (!nl) nl = e;
        const char *b = p;
        while (b < nl && isspace((unsigned char)*b)) ++b;
        if (b == nl) {
            p = nl < e ? nl + 1 : e;
            continue;
        }
        SomaticRow row;
        if (*b == '@') {
            // `fin >> up_chr; getline(fin, temp); fout << up_chr << temp << ...` (somatic.cpp:59-65)
            row.is_header = true;
            row.prefix = std::string(b, nl) + "\tleft_clip_read_NO_of_control\tright_clip_read_NO_of_control\tabnormal_read_pair_no_of_control\n";
            rows.push_back(row);
            p = nl < e ? nl + 1 : e;
            continue;
        }
        std::vector<std::string> t = split_ws(b, nl);
        p = nl < e ? nl + 1 : e;
        if (t.size() < 23) continue;
        const std::string &up_chr = t[0], &down_chr = t[4], &up_seq = t[21], &down_seq = t[22];
        int up_pos = atoi(t[1].c_str()), up_n = atoi(t[3].c_str()), down_pos = atoi(t[5].c_str()), down_n = atoi(t[7].c_str());
        int micro = atoi(t[8].c_str());
        char us = t[2][0], ds = t[6][0];
        row.key = make_key(up_chr, up_pos, us, down_chr, down_pos, ds);
        int nl_ = 0, nr_ = 0;
        bool written = true, always = false;
        std::string rc_up = reverse_complement(up_seq), rc_down = reverse_complement(down_seq);
        if (us == '+' && ds == '+') {
            if (micro != -1) {
                nr_ = first_match(t5, down_chr, down_pos, down_seq, up_seq, rate);
                if (down_seq.length() >= (size_t)micro)
                    nl_ = first_match(t3, up_chr, up_pos + micro, down_seq.substr(micro), up_seq + down_seq.substr(0, micro), rate);
                always = true;  // somatic.cpp:111 runs the pair query unconditionally
            } else if (up_n == 0) {
                nr_ = first_match(t5, down_chr, down_pos, down_seq, up_seq, rate);
                nl_ = window_first(t3, up_chr, up_pos, up_pos + offset,
                                   [&](const NormalClip &c) { return shifted_compare(c.left, c.right, up_seq, down_seq, rate) != -1; });
            } else if (down_n == 0) {
                nl_ = first_match(t3, up_chr, up_pos, down_seq, up_seq, rate);
                nr_ = window_first(t5, down_chr, down_pos - offset, down_pos,
                                   [&](const NormalClip &c) { return shifted_compare(up_seq, down_seq, c.left, c.right, rate) != -1; });
            } else
                written = false;
        } else if (us == '+' && ds == '-') {
            if (micro != -1) {
                nl_ = first_match(t3, up_chr, up_pos + micro, down_seq.substr(micro), up_seq + down_seq.substr(0, micro), rate);
                nr_ = first_match(t3, down_chr, down_pos, rc_up, rc_down, rate);
            } else if (up_n == 0) {
                nr_ = first_match(t3, down_chr, down_pos, rc_up, rc_down, rate);
                nl_ = window_first(t3, up_chr, up_pos, up_pos + offset,
                                   [&](const NormalClip &c) { return shifted_compare(c.left, c.right, up_seq, down_seq, rate) != -1; });
            } else if (down_n == 0) {
                nl_ = first_match(t3, up_chr, up_pos, down_seq, up_seq, rate);
                nr_ = window_first(t3, down_chr, down_pos, down_pos + offset,
                                   [&](const NormalClip &c) { return shifted_compare(c.left, c.right, rc_down, rc_up, rate) != -1; });
            } else
                written = false;
        } else if (us == '-' && ds == '+') {
            if (micro != -1) {
                nl_ = first_match(t5, up_chr, up_pos, rc_up, rc_down, rate);
                nr_ = first_match(t5, down_chr, down_pos - micro, up_seq.substr(up_seq.length() - micro) + down_seq,
                                  up_seq.substr(0, up_seq.length() - micro), rate);
            } else if (up_n == 0) {
                nr_ = first_match(t5, down_chr, down_pos, down_seq, up_seq, rate);
                nl_ = window_first(t5, up_chr, up_pos - offset, up_pos,
                                   [&](const NormalClip &c) { return shifted_compare(rc_up, rc_down, c.left, c.right, rate) != -1; });
            } else if (down_n == 0) {
                nl_ = first_match(t5, up_chr, up_pos, rc_up, rc_down, rate);
                nr_ = window_first(t5, down_chr, down_pos - offset, down_pos,
                                   [&](const NormalClip &c) { return shifted_compare(up_seq, down_seq, c.left, c.right, rate) != -1; });
            } else
                written = false;
        } else
            written = false;
        if (!written) {
            log += "The tandem repeat length is error in postion: " + up_chr + "\t" + std::to_string(up_pos) + "\n";
            continue;
        }
        row.normal_left = nl_, row.normal_right = nr_;
        row.query_pairs = always || mean_insert != 0;
        // the 23 tumour columns, numbers re-parsed and re-printed as the reference's iostreams do (somatic.cpp:66,112)
        std::string &o = row.prefix;
        o = up_chr + "\t" + std::to_string(up_pos) + "\t" + us + "\t" + std::to_string(up_n) + "\t" + down_chr + "\t" +
            std::to_string(down_pos) + "\t" + ds + "\t" + std::to_string(down_n) + "\t" + std::to_string(micro) + "\t" +
            std::to_string(atoi(t[9].c_str())) + "\t" + t[10];
        for (int i = 11; i <= 16; ++i) o += "\t" + std::to_string(atoi(t[i].c_str()));
        o += "\t" + format_double(strtod(t[17].c_str(), nullptr)) + "\t" + format_double(strtod(t[18].c_str(), nullptr));
        o += "\t" + t[19] + "\t" + t[20] + "\t" + up_seq + "\t" + down_seq;
        rows.push_back(row);
    }
}

}  // namespace svb
