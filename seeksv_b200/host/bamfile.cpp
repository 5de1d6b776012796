#include "bamfile.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <thread>

bool MappedFile::open(const std::string &path, std::string &err)
{
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) {
        err = "cannot open " + path;
        return false;
    }
    this->fd = -1;
    struct stat st;
    if (fstat(fd, &st) != 0) {
        ::close(fd);
        err = "cannot stat " + path;
        return false;
    }
    size = (uint64_t)st.st_size;
    if (size) {
        void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) {
            ::close(fd);
            err = "cannot map " + path;
            return false;
        }
        data = (const uint8_t *)p;
    }
    this->fd = fd;
    return true;
}
MappedFile::~MappedFile()
{
    if (data) munmap((void *)data, size);
    if (fd >= 0) ::close(fd);
}

bool read_file(const std::string &path, std::vector<uint8_t> &out, std::string &err)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        err = "cannot open " + path;
        return false;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != out.size()) {
        err = "short read on " + path;
        return false;
    }
    return true;
}

// BGZF: gzip members with FEXTRA holding the 'BC' sub-field = member size - 1 (sam/bgzf.h:34-60)
// header at f[o]: returns the member size, 0 if this is not a BGZF member header
static uint32_t bgzf_member(const uint8_t *f, uint64_t n, uint64_t o, uint32_t *xlen_out)
{
    if (o + 18 > n || f[o] != 0x1f || f[o + 1] != 0x8b || f[o + 2] != 8 || !(f[o + 3] & 4)) return 0;
    uint32_t xlen = f[o + 10] | (f[o + 11] << 8);
    uint64_t x = o + 12, xe = x + xlen;
    if (xe > n) return 0;
    uint32_t bsize = 0;
    while (x + 4 <= xe) {
        uint32_t slen = f[x + 2] | (f[x + 3] << 8);
        if (f[x] == 'B' && f[x + 1] == 'C' && slen == 2 && x + 6 <= xe) bsize = (f[x + 4] | (f[x + 5] << 8)) + 1;
        x += 4 + slen;
    }
    if (!bsize || o + bsize > n || bsize < 12 + xlen + 8) return 0;
    *xlen_out = xlen;
    return bsize;
}

uint32_t bgzf_block_size(const uint8_t *file, uint64_t n, uint64_t offset)
{
    uint32_t xl;
    return bgzf_member(file, n, offset, &xl);
}

// The member chain is serial (each header gives the next offset). Like the record chain on the device it is cut into
// segments whose first header is GUESSED (magic + BC + a valid successor), walked in parallel, and stitched only where
// each segment's exit equals the next segment's guess; anything else falls back to the serial walk.
static bool walk_members(const uint8_t *f, uint64_t n, uint64_t from, uint64_t until, std::vector<BgzfBlock> &out, uint64_t &exit_)
{
    uint64_t o = from;
    while (o < until) {
        uint32_t xlen, bsize = bgzf_member(f, n, o, &xlen);
        if (!bsize) return false;
        BgzfBlock b;
        b.coff = o + 12 + xlen;
        b.clen = bsize - 12 - xlen - 8;
        const uint8_t *t = f + o + bsize - 4;
        b.ulen = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
        if (b.ulen > 65536) return false;  // BGZF: at most 64 KiB per block (libbam inflates into a fixed 64 KiB buffer)
        b.uoff = 0;
        out.push_back(b);
        o += bsize;
    }
    exit_ = o;
    return true;
}

bool bgzf_scan(const uint8_t *f, uint64_t n, std::vector<BgzfBlock> &blocks, uint64_t &total, std::string &err)
{
    total = 0;
    blocks.clear();
    int nseg = (int)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), n >> 22);  // >= 4 MiB per segment
    bool stitched = false;
    if (nseg > 1) {
        std::vector<uint64_t> guess(nseg, 0), exit_(nseg, 0);
        std::vector<std::vector<BgzfBlock>> part(nseg);
        std::vector<char> ok(nseg, 0);
        std::vector<std::thread> th;
        for (int s = 0; s < nseg; ++s)
            th.emplace_back([&, s]() {
                uint64_t lo = n / nseg * s, hi = s + 1 == nseg ? n : n / nseg * (s + 1);
                uint64_t g = lo;
                if (s > 0) {
                    g = n;
                    for (uint64_t o = lo; o < std::min(n, lo + (1u << 17)); ++o) {
                        uint32_t xl, bs = bgzf_member(f, n, o, &xl), xl2;
                        if (bs && (o + bs == n || bgzf_member(f, n, o + bs, &xl2))) {
                            g = o;
                            break;
                        }
                    }
                }
                guess[s] = g;
                ok[s] = g <= hi && walk_members(f, n, g, hi, part[s], exit_[s]);
            });
        for (auto &t : th) t.join();
        stitched = true;
        for (int s = 0; s < nseg && stitched; ++s) stitched = ok[s] && (s == 0 || exit_[s - 1] == guess[s]);
        if (stitched && exit_[nseg - 1] != n) stitched = false;
        if (stitched)
            for (auto &p : part) blocks.insert(blocks.end(), p.begin(), p.end());
    }
    if (!stitched) {
        blocks.clear();
        uint64_t e = 0;
        if (!walk_members(f, n, 0, n, blocks, e) || e != n) {
            err = "not a BGZF file (bad block header)";
            return false;
        }
    }
    size_t w = 0;
    for (size_t i = 0; i < blocks.size(); ++i) {
        blocks[i].uoff = total;
        total += blocks[i].ulen;
        if (blocks[i].ulen) blocks[w++] = blocks[i];
    }
    blocks.resize(w);
    return true;
}

static bool inflate_block(const uint8_t *src, uint32_t clen, uint8_t *dst, uint32_t ulen);

bool bai_first_offsets(const std::string &bai_path, std::vector<uint64_t> &first_voff, std::string &err)
{
    std::vector<uint8_t> d;
    if (!read_file(bai_path, d, err)) return false;
    auto bad = [&]() {
        err = "malformed index " + bai_path;
        return false;
    };
    if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return bad();
    size_t o = 4;
    auto i32 = [&](int32_t &v) {
        if (o + 4 > d.size()) return false;
        memcpy(&v, &d[o], 4);
        o += 4;
        return true;
    };
    int32_t n_ref = 0;
    if (!i32(n_ref) || n_ref < 0) return bad();
    first_voff.assign((size_t)n_ref, ~0ull);
    for (int32_t t = 0; t < n_ref; ++t) {
        int32_t n_bin = 0;
        if (!i32(n_bin) || n_bin < 0) return bad();
        for (int32_t b = 0; b < n_bin; ++b) {
            int32_t bin_raw = 0, n_chunk = 0;
            if (!i32(bin_raw) || !i32(n_chunk) || n_chunk < 0 || o + 16ull * (uint64_t)n_chunk > d.size()) return bad();
            const uint32_t bin = (uint32_t)bin_raw;
            for (int32_t c = 0; c < n_chunk; ++c) {
                uint64_t beg;
                memcpy(&beg, &d[o + 16 * (size_t)c], 8);
                if (bin != 37450u) first_voff[t] = std::min(first_voff[t], beg);  // 37450: samtools' metadata pseudo-bin
            }
            o += 16 * (size_t)n_chunk;
        }
        int32_t n_intv = 0;
        if (!i32(n_intv) || n_intv < 0 || o + 8ull * (uint64_t)n_intv > d.size()) return bad();
        o += 8 * (size_t)n_intv;
    }
    return true;
}

bool bai_linear_offsets(const std::string &bai_path, std::vector<std::vector<uint64_t>> &linear, std::string &err)
{
    std::vector<uint8_t> d;
    if (!read_file(bai_path, d, err)) return false;
    auto bad = [&]() {
        err = "malformed index " + bai_path;
        return false;
    };
    if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return bad();
    size_t o = 4;
    auto i32 = [&](int32_t &v) {
        if (o + 4 > d.size()) return false;
        memcpy(&v, &d[o], 4);
        o += 4;
        return true;
    };
    int32_t n_ref = 0;
    if (!i32(n_ref) || n_ref < 0) return bad();
    linear.assign((size_t)n_ref, {});
    for (int32_t t = 0; t < n_ref; ++t) {
        int32_t n_bin = 0;
        if (!i32(n_bin) || n_bin < 0) return bad();
        for (int32_t b = 0; b < n_bin; ++b) {
            int32_t bin_raw = 0, n_chunk = 0;
            if (!i32(bin_raw) || !i32(n_chunk) || n_chunk < 0 || o + 16ull * (uint64_t)n_chunk > d.size()) return bad();
            o += 16 * (size_t)n_chunk;
        }
        int32_t n_intv = 0;
        if (!i32(n_intv) || n_intv < 0 || o + 8ull * (uint64_t)n_intv > d.size()) return bad();
        linear[t].resize((size_t)n_intv);
        if (n_intv) memcpy(linear[t].data(), &d[o], 8 * (size_t)n_intv);
        o += 8 * (size_t)n_intv;
    }
    return true;
}

bool bgzf_voffset_distance(const uint8_t *f, uint64_t n, uint64_t v_a, uint64_t v_b, uint64_t &bytes, std::string &err)
{
    uint64_t c = v_a >> 16, cb = v_b >> 16, acc = 0;
    if (v_a > v_b || cb > n) {
        err = "virtual offsets out of order";
        return false;
    }
    while (c < cb) {
        uint32_t xl, bs = bgzf_member(f, n, c, &xl);
        if (!bs) {
            err = "virtual offset does not point at a BGZF block";
            return false;
        }
        const uint8_t *t = f + c + bs - 4;
        acc += t[0] | (t[1] << 8) | (t[2] << 16) | ((uint64_t)t[3] << 24);
        c += bs;
    }
    if (c != cb) {
        err = "virtual offset does not point at a BGZF block";
        return false;
    }
    bytes = acc + (v_b & 0xffff) - (v_a & 0xffff);
    return true;
}

bool bgzf_read_at(const uint8_t *f, uint64_t n, uint64_t voff, uint8_t *dst, uint32_t want, std::string &err)
{
    uint64_t c = voff >> 16, skip = voff & 0xffff;
    uint32_t got = 0;
    std::vector<uint8_t> buf;
    while (got < want) {
        uint32_t xl, bs = c < n ? bgzf_member(f, n, c, &xl) : 0;
        if (!bs) {
            err = "virtual offset does not point at a BGZF block";
            return false;
        }
        const uint8_t *t = f + c + bs - 4;
        uint32_t ulen = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
        buf.resize(ulen);
        if (ulen && !inflate_block(f + c + 12 + xl, bs - xl - 20, buf.data(), ulen)) {
            err = "BGZF inflate failed";
            return false;
        }
        for (uint64_t i = skip; i < ulen && got < want; ++i) dst[got++] = buf[i];
        skip = skip > ulen ? skip - ulen : 0;
        c += bs;
    }
    return true;
}

bool read_bam_header(const uint8_t *f, uint64_t n, BamHeader &h, std::string &err)
{
    std::vector<uint8_t> text;
    std::vector<BgzfBlock> blocks;
    uint64_t o = 0, want = 1u << 20;
    for (;;) {
        while (o < n && text.size() < want) {
            uint32_t xl, bs = bgzf_member(f, n, o, &xl);
            if (!bs || o + bs > n) {
                err = "not a BGZF file (bad block header)";
                return false;
            }
            uint32_t ulen = f[o + bs - 4] | (f[o + bs - 3] << 8) | (f[o + bs - 2] << 16) | ((uint32_t)f[o + bs - 1] << 24);
            size_t at = text.size();
            text.resize(at + ulen);
            if (ulen && !inflate_block(f + o + 12 + xl, bs - xl - 20, text.data() + at, ulen)) {
                err = "BGZF inflate failed";
                return false;
            }
            o += bs;
        }
        std::string e2;
        if (parse_bam_header(text.data(), text.size(), h, e2)) return true;
        if (o >= n) {
            err = e2;
            return false;
        }
        want *= 4;
    }
}

static bool inflate_block(const uint8_t *src, uint32_t clen, uint8_t *dst, uint32_t ulen)
{
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return false;
    zs.next_in = const_cast<Bytef *>(src);
    zs.avail_in = clen;
    zs.next_out = dst;
    zs.avail_out = ulen;
    int r = inflate(&zs, Z_FINISH);
    bool ok = (r == Z_STREAM_END) && zs.total_out == ulen;
    inflateEnd(&zs);
    return ok;
}

bool bgzf_inflate_range(const uint8_t *file, const std::vector<BgzfBlock> &blocks, size_t b0, size_t b1, uint8_t *dst,
                        int n_threads, std::string &err)
{
    if (b0 >= b1) return true;
    std::atomic<size_t> next(b0);
    std::atomic<bool> bad(false);
    uint64_t base = blocks[b0].uoff;
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(8);
            if (i >= b1 || bad.load()) return;
            for (size_t k = i; k < std::min(i + 8, b1); ++k)
                if (!inflate_block(file + blocks[k].coff, blocks[k].clen, dst + (blocks[k].uoff - base), blocks[k].ulen)) bad = true;
        }
    };
    int nt = (int)std::min<size_t>((size_t)std::max(1, n_threads), (b1 - b0 + 7) / 8);
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (bad) {
        err = "BGZF inflate failed";
        return false;
    }
    return true;
}

bool bgzf_inflate_all(const uint8_t *file, uint64_t n, std::vector<uint8_t> &out, int n_threads, std::string &err)
{
    std::vector<BgzfBlock> blocks;
    uint64_t total;
    if (!bgzf_scan(file, n, blocks, total, err)) return false;
    out.resize(total);
    return bgzf_inflate_range(file, blocks, 0, blocks.size(), out.data(), n_threads, err);
}

static inline uint32_t rd32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }

bool parse_bam_header(const uint8_t *d, uint64_t n, BamHeader &h, std::string &err)
{
    if (n < 12 || memcmp(d, "BAM\1", 4) != 0) {
        err = "[main_samview] fail to read the header.";
        return false;
    }
    uint64_t o = 4;
    uint32_t l_text = rd32(d + o);
    o += 4;
    if (o + l_text + 4 > n) {
        err = "BAM header truncated";
        return false;
    }
    h.text.assign((const char *)d + o, l_text);
    o += l_text;
    uint32_t n_ref = rd32(d + o);
    o += 4;
    h.names.clear();
    h.lengths.clear();
    for (uint32_t i = 0; i < n_ref; ++i) {
        if (o + 4 > n) {
            err = "BAM header truncated";
            return false;
        }
        uint32_t l = rd32(d + o);
        o += 4;
        if (o + l + 4 > n) {
            err = "BAM header truncated";
            return false;
        }
        h.names.emplace_back((const char *)d + o, strnlen((const char *)d + o, l));
        o += l;
        h.lengths.push_back(rd32(d + o));
        o += 4;
    }
    h.first_record = o;
    return true;
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

static void put32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back(x & 0xff), v.push_back((x >> 8) & 0xff), v.push_back((x >> 16) & 0xff), v.push_back(x >> 24);
}

// FLAG column as the linked libbam's sam_read1 takes it: a number in any C base (strtol base 0), else flag letters
uint32_t sam_flag(const char *a, const char *b)
{
    if (a < b && *a >= '0' && *a <= '9') return (uint32_t)strtol(a, nullptr, 0);
    uint32_t f = 0;
    for (; a < b; ++a) {
        const char *tab = "pPuUrR12sfd", *o = strchr(tab, *a);
        if (o && *a) f |= 1u << (o - tab);
    }
    return f;
}

bool sam_to_bam_stream(const std::vector<uint8_t> &text, BamHeader &h, std::vector<uint8_t> &out, std::string &err)
{
    static uint8_t nt16[256];
    static bool init = false;
    if (!init) {
        memset(nt16, 15, sizeof nt16);
        const char *tab = "=ACMGRSVTWYHKDBN";
        for (int i = 0; i < 16; ++i) nt16[(uint8_t)tab[i]] = i, nt16[(uint8_t)tolower(tab[i])] = i;
        init = true;
    }
    const char *p = (const char *)text.data(), *e = p + text.size();
    std::map<std::string, int> name2tid;
    h = BamHeader();
    // header lines
    while (p < e && *p == '@') {
        const char *nl = (const char *)memchr(p, '\n', e - p);
        if (!nl) nl = e;
        std::string line(p, nl);
        h.text += line + "\n";
        if (line.compare(0, 3, "@SQ") == 0) {
            std::string sn;
            uint32_t ln = 0;
            size_t a = 0;
            while (a < line.size()) {
                size_t b = line.find('\t', a);
                if (b == std::string::npos) b = line.size();
                if (line.compare(a, 3, "SN:") == 0) sn = line.substr(a + 3, b - a - 3);
                if (line.compare(a, 3, "LN:") == 0) ln = (uint32_t)strtoul(line.c_str() + a + 3, nullptr, 10);
                a = b + 1;
            }
            name2tid[sn] = (int)h.names.size();
            h.names.push_back(sn);
            h.lengths.push_back(ln);
        }
        p = nl < e ? nl + 1 : e;
    }
    if (h.names.empty()) {
        err = "[main_samview] fail to read the header.";
        return false;
    }
    out.clear();
    out.insert(out.end(), {'B', 'A', 'M', 1});
    put32(out, (uint32_t)h.text.size());
    out.insert(out.end(), h.text.begin(), h.text.end());
    put32(out, (uint32_t)h.names.size());
    for (size_t i = 0; i < h.names.size(); ++i) {
        put32(out, (uint32_t)h.names[i].size() + 1);
        out.insert(out.end(), h.names[i].begin(), h.names[i].end());
        out.push_back(0);
        put32(out, h.lengths[i]);
    }
    h.first_record = out.size();
    std::vector<std::pair<const char *, const char *>> f;
    while (p < e) {
        const char *nl = (const char *)memchr(p, '\n', e - p);
        if (!nl) nl = e;
        const char *le = nl;
        if (le > p && le[-1] == '\r') --le;
        if (le == p) {
            p = nl < e ? nl + 1 : e;
            continue;
        }
        f.clear();
        for (const char *a = p;;) {
            const char *b = (const char *)memchr(a, '\t', le - a);
            if (!b) b = le;
            f.emplace_back(a, b);
            if (b == le) break;
            a = b + 1;
        }
        if (f.size() < 11) {
            err = "SAM line with fewer than 11 fields";
            return false;
        }
        auto str = [&](int i) { return std::string(f[i].first, f[i].second); };
        std::string qname = str(0), rname = str(2), rnext = str(6);
        uint32_t flag = sam_flag(f[1].first, f[1].second);
        int32_t pos = (int32_t)strtol(f[3].first, nullptr, 10) - 1, mapq = (int32_t)strtol(f[4].first, nullptr, 10);
        int32_t mpos = (int32_t)strtol(f[7].first, nullptr, 10) - 1, isize = (int32_t)strtol(f[8].first, nullptr, 10);
        int32_t tid = -1, mtid = -1;
        if (rname != "*") {
            auto it = name2tid.find(rname);
            if (it != name2tid.end()) tid = it->second;
        }
        if (rnext == "=") mtid = tid;
        else if (rnext != "*") {
            auto it = name2tid.find(rnext);
            if (it != name2tid.end()) mtid = it->second;
        }
        std::vector<uint32_t> cigar;
        int32_t end = pos;
        if (!(f[5].second - f[5].first == 1 && *f[5].first == '*')) {
            uint32_t num = 0;
            for (const char *c = f[5].first; c < f[5].second; ++c) {
                if (*c >= '0' && *c <= '9') num = num * 10 + (*c - '0');
                else {
                    const char *ops = "MIDNSHP=X", *o = strchr(ops, *c);
                    if (!o) {
                        err = "invalid CIGAR character";
                        return false;
                    }
                    uint32_t op = (uint32_t)(o - ops);
                    cigar.push_back(num << 4 | op);
                    if (op == 0 || op == 2 || op == 3) end += num;
                    num = 0;
                }
            }
        }
        else
            flag |= 4;  // libbam's sam_read1 ("mapped sequence without CIGAR"): a record with CIGAR "*" leaves as unmapped
        if (end == pos) end = pos + 1;
        bool noseq = f[9].second - f[9].first == 1 && *f[9].first == '*';
        int32_t l_qseq = noseq ? 0 : (int32_t)(f[9].second - f[9].first);
        std::vector<uint8_t> aux;
        for (size_t i = 11; i < f.size(); ++i) {
            const char *a = f[i].first, *b = f[i].second;
            if (b - a < 5 || a[2] != ':' || a[4] != ':') continue;
            char ty = a[3];
            const char *v = a + 5;
            if (ty == 'i' || ty == 'I') {
                long long x = strtoll(v, nullptr, 10);
                aux.push_back(a[0]), aux.push_back(a[1]);
                if (x < 0) {
                    if (x >= -127) aux.push_back('c'), aux.push_back((uint8_t)(int8_t)x);
                    else if (x >= -32767) aux.push_back('s'), aux.push_back(x & 0xff), aux.push_back((x >> 8) & 0xff);
                    else aux.push_back('i'), put32(aux, (uint32_t)x);
                } else {
                    if (x <= 255) aux.push_back('C'), aux.push_back((uint8_t)x);
                    else if (x <= 65535) aux.push_back('S'), aux.push_back(x & 0xff), aux.push_back((x >> 8) & 0xff);
                    else aux.push_back('I'), put32(aux, (uint32_t)x);
                }
            } else if (ty == 'A' || ty == 'a' || ty == 'c' || ty == 'C') {
                aux.push_back(a[0]), aux.push_back(a[1]), aux.push_back('A'), aux.push_back(*v);
            } else if (ty == 'Z' || ty == 'H') {
                aux.push_back(a[0]), aux.push_back(a[1]), aux.push_back(ty);
                aux.insert(aux.end(), v, b);
                aux.push_back(0);
            } else if (ty == 'f') {
                float x = strtof(v, nullptr);
                uint32_t u;
                memcpy(&u, &x, 4);
                aux.push_back(a[0]), aux.push_back(a[1]), aux.push_back('f');
                put32(aux, u);
            } else if (ty == 'd') {
                double x = strtod(v, nullptr);
                uint64_t u;
                memcpy(&u, &x, 8);
                aux.push_back(a[0]), aux.push_back(a[1]), aux.push_back('d');
                put32(aux, (uint32_t)u), put32(aux, (uint32_t)(u >> 32));
            } else if (ty == 'B' && v < b) {  // typed array: XB:B:i,1,2,3
                char sub = *v;
                uint32_t n = 0;
                for (const char *c = v; c < b; ++c) n += *c == ',';
                aux.push_back(a[0]), aux.push_back(a[1]), aux.push_back('B'), aux.push_back((uint8_t)sub);
                put32(aux, n);
                const char *c = v + 1;
                for (uint32_t k = 0; k < n && c < b; ++k) {
                    ++c;  // the comma
                    char *next = nullptr;
                    if (sub == 'f') {
                        float x = strtof(c, &next);
                        uint32_t u;
                        memcpy(&u, &x, 4);
                        put32(aux, u);
                    } else {
                        long long x = strtoll(c, &next, 0);
                        if (sub == 'c' || sub == 'C') aux.push_back((uint8_t)x);
                        else if (sub == 's' || sub == 'S') aux.push_back(x & 0xff), aux.push_back((x >> 8) & 0xff);
                        else put32(aux, (uint32_t)x);
                    }
                    c = next && next > c ? next : b;
                }
            }
        }
        // libbam keeps l_qname in 8 bits and copies that many bytes: a name of 255+ characters leaves truncated and without its NUL
        uint32_t l_qname = ((uint32_t)qname.size() + 1) & 0xff;
        uint32_t bs = 32 + l_qname + 4 * (uint32_t)cigar.size() + (l_qseq + 1) / 2 + l_qseq + (uint32_t)aux.size();
        put32(out, bs);
        put32(out, (uint32_t)tid);
        put32(out, (uint32_t)pos);
        put32(out, (uint32_t)reg2bin(pos, end) << 16 | ((uint32_t)mapq & 0xff) << 8 | l_qname);  // (pos -1 gives bin 4680)
        put32(out, flag << 16 | ((uint32_t)cigar.size() & 0xffff));
        put32(out, (uint32_t)l_qseq);
        put32(out, (uint32_t)mtid);
        put32(out, (uint32_t)mpos);
        put32(out, (uint32_t)isize);
        if (l_qname == qname.size() + 1) {
            out.insert(out.end(), qname.begin(), qname.end());
            out.push_back(0);
        } else
            out.insert(out.end(), qname.begin(), qname.begin() + l_qname);
        for (uint32_t c : cigar) put32(out, c);
        for (int32_t i = 0; i < l_qseq; i += 2) {
            uint8_t hi = nt16[(uint8_t)f[9].first[i]], lo = i + 1 < l_qseq ? nt16[(uint8_t)f[9].first[i + 1]] : 0;
            out.push_back(hi << 4 | lo);
        }
        bool noqual = f[10].second - f[10].first == 1 && *f[10].first == '*';
        for (int32_t i = 0; i < l_qseq; ++i) out.push_back(noqual ? 0xff : (uint8_t)(f[10].first[i] - 33));
        out.insert(out.end(), aux.begin(), aux.end());
        p = nl < e ? nl + 1 : e;
    }
    return true;
}

// gzip file written by write_gz_many: every member announces its size in an 'SV' extra sub-field
// Inflate of the members our own writers produce (svb_write_gz, gzip.cu): dynamic-Huffman blocks of LITERALS only. zlib decodes
// such a stream one symbol per table lookup (~120 MB/s per thread: getsv spent 14-31 ms on the 30 MB clip text of C2); here a
// 12-bit table returns up to two literals per lookup. Anything else in the stream (a stored or fixed block, a length symbol, a
// damaged code) makes this return false and the caller falls back to zlib, which also reports the errors.
namespace {
struct LitBits {
    const uint8_t *p, *end;
    uint64_t buf = 0;
    int cnt = 0;
    inline void refill()
    {
        if (end - p >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            buf |= w << cnt;
            const int adv = (63 - cnt) >> 3;
            p += adv, cnt += adv * 8;
        } else
            while (cnt <= 56 && p < end) buf |= (uint64_t)*p++ << cnt, cnt += 8;
    }
    inline uint32_t take(int k)
    {
        const uint32_t v = (uint32_t)(buf & ((1ull << k) - 1));
        buf >>= k, cnt -= k;
        return v;
    }
};

bool inflate_literal_stream(const uint8_t *in, size_t n, uint8_t *out, uint32_t ulen)
{
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    constexpr int TB = 12;
    std::vector<uint32_t> tab(1u << TB), one(1u << TB);
    LitBits b{in, in + n};
    uint8_t *o = out, *const oe = out + ulen;
    for (;;) {
        b.refill();
        if (b.cnt < 17) return false;
        const uint32_t final_block = b.take(1), type = b.take(2);
        if (type != 2) return false;
        const int n_lit = (int)b.take(5) + 257, n_dist = (int)b.take(5) + 1, n_cl = (int)b.take(4) + 4;
        if (n_lit > 286 || n_dist > 30) return false;
        uint8_t cl[19] = {0}, lens[320] = {0};
        for (int i = 0; i < n_cl; ++i) {
            b.refill();
            cl[order[i]] = (uint8_t)b.take(3);
        }
        // code-length alphabet: canonical codes, 7-bit lookup
        uint8_t cl_sym[128], cl_len[128];
        memset(cl_len, 0, sizeof cl_len);
        {
            int cnt[8] = {0};
            for (int i = 0; i < 19; ++i) cnt[cl[i]]++;
            cnt[0] = 0;
            uint32_t next[8], code = 0;
            for (int l = 1; l < 8; ++l) code = (code + cnt[l - 1]) << 1, next[l] = code;
            for (int s = 0; s < 19; ++s)
                if (cl[s]) {
                    uint32_t c = next[cl[s]]++, r = 0;
                    for (int k = 0; k < cl[s]; ++k) r |= ((c >> k) & 1u) << (cl[s] - 1 - k);
                    for (uint32_t k = r; k < 128; k += 1u << cl[s]) cl_sym[k] = (uint8_t)s, cl_len[k] = cl[s];
                }
        }
        for (int i = 0, total = n_lit + n_dist; i < total;) {
            b.refill();
            const uint32_t k = (uint32_t)(b.buf & 127);
            if (!cl_len[k] || b.cnt < cl_len[k] + 7) return false;
            b.take(cl_len[k]);
            const int s = cl_sym[k];
            if (s < 16) lens[i++] = (uint8_t)s;
            else {
                int rep, v = 0;
                if (s == 16) {
                    if (!i) return false;
                    v = lens[i - 1], rep = 3 + (int)b.take(2);
                } else if (s == 17) rep = 3 + (int)b.take(3);
                else rep = 11 + (int)b.take(7);
                if (i + rep > total) return false;
                while (rep--) lens[i++] = (uint8_t)v;
            }
        }
        if (!lens[256]) return false;
        // literal/length alphabet: canonical codes -> single-symbol table, then pairs
        int cnt[16] = {0};
        for (int s = 0; s < n_lit; ++s) cnt[lens[s]]++;
        cnt[0] = 0;
        uint32_t next[16], first_code[16], code = 0;
        int left = 1;
        for (int l = 1; l <= 15; ++l) {
            code = (code + cnt[l - 1]) << 1, next[l] = first_code[l] = code;
            left = (left << 1) - cnt[l];
            if (left < 0) return false;
        }
        uint16_t sorted[288], offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + cnt[l]);
        {
            uint16_t at[16];
            memcpy(at, offs, sizeof at);
            for (int s = 0; s < n_lit; ++s)
                if (lens[s]) sorted[at[lens[s]]++] = (uint16_t)s;
        }
        std::fill(one.begin(), one.end(), 0u);
        for (int s = 0; s < 256 && s < n_lit; ++s) {  // (end of block and length symbols take the slow path)
            const int l = lens[s];
            if (!l) continue;
            const uint32_t c = next[l]++;
            if (l > TB) continue;
            uint32_t r = 0;
            for (int k = 0; k < l; ++k) r |= ((c >> k) & 1u) << (l - 1 - k);
            const uint32_t e = (uint32_t)l | 1u << 4 | (uint32_t)s << 8;
            for (uint32_t k = r; k < (1u << TB); k += 1u << l) one[k] = e;
        }
        for (uint32_t k = 0; k < (1u << TB); ++k) {
            const uint32_t e = one[k];
            uint32_t t = e;
            if (e) {
                const int l1 = (int)(e & 15);
                const uint32_t e2 = one[k >> l1];  // the bits behind the first code, zero-extended: trusted only if the code fits
                if (e2 && l1 + (int)(e2 & 15) <= TB) t = (uint32_t)(l1 + (e2 & 15)) | 2u << 4 | (e & 0xff00u) | ((e2 >> 8) & 0xffu) << 16;
            }
            tab[k] = t;
        }
        for (;;) {  // tokens of this block
            b.refill();
            if (oe - o >= 8 && b.cnt >= 48) {
                for (int rep = 0; rep < 3; ++rep) {  // three lookups (<= 36 bits) per refill
                    const uint32_t e = tab[b.buf & ((1u << TB) - 1)];
                    if (!e) goto slow;
                    o[0] = (uint8_t)(e >> 8), o[1] = (uint8_t)(e >> 16);
                    o += (e >> 4) & 3;
                    b.buf >>= e & 15, b.cnt -= (int)(e & 15);
                }
                continue;
            }
        slow: {
            // one symbol, canonically (long codes, the end of the block, the last bytes of the member)
            if (b.cnt < 1) return false;
            uint32_t c = 0;
            int l = 0, s = -1;
            uint64_t w = b.buf;
            for (l = 1; l <= 15; ++l) {
                c = c << 1 | (uint32_t)(w & 1);
                w >>= 1;
                if (cnt[l] && c >= first_code[l] && c - first_code[l] < (uint32_t)cnt[l]) {
                    s = sorted[offs[l] + (c - first_code[l])];
                    break;
                }
            }
            if (s < 0 || b.cnt < l) return false;
            b.take(l);
            if (s == 256) break;
            if (s > 256 || o >= oe) return false;
            *o++ = (uint8_t)s;
        }
        }
        if (final_block) break;
    }
    return o == oe;
}
}  // namespace

static bool read_indexed_gz(const uint8_t *f, uint64_t n, std::string &out)
{
    struct Member {
        uint64_t off, size, uoff;
        uint32_t ulen;
    };
    std::vector<Member> ms;
    uint64_t o = 0, total = 0;
    while (o < n) {
        if (o + 28 > n || f[o] != 0x1f || f[o + 1] != 0x8b || f[o + 2] != 8 || f[o + 3] != 4) return false;
        if (f[o + 10] != 8 || f[o + 11] != 0 || f[o + 12] != 'S' || f[o + 13] != 'V' || f[o + 14] != 4 || f[o + 15] != 0) return false;
        uint64_t sz = f[o + 16] | (f[o + 17] << 8) | (f[o + 18] << 16) | ((uint64_t)f[o + 19] << 24);
        if (sz < 28 || o + sz > n) return false;
        const uint8_t *t = f + o + sz - 4;
        uint32_t ulen = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
        ms.push_back(Member{o, sz, total, ulen});
        total += ulen;
        o += sz;
    }
    out.resize(total);
    std::atomic<size_t> next(0);
    std::atomic<bool> bad(false);
    const char *mode = getenv("SEEKSV_B200_GZ_READ");  // "zlib": every member through zlib (tests compare the two)
    const bool fast = !(mode && strcmp(mode, "zlib") == 0);
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= ms.size()) return;
            if (fast && ms[i].ulen && ms[i].size > 28) {  // our writers' literal-only members: the fast decoder, checked like zlib checks them
                const uint8_t *m = f + ms[i].off, *t = m + ms[i].size - 8;
                const uint32_t want_crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
                uint8_t *dst = (uint8_t *)&out[ms[i].uoff];
                if (inflate_literal_stream(m + 20, (size_t)ms[i].size - 28, dst, ms[i].ulen) &&
                    (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, ms[i].ulen) == want_crc)
                    continue;
            }
            z_stream zs;
            memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, 15 + 16) != Z_OK) {
                bad = true;
                return;
            }
            zs.next_in = const_cast<Bytef *>(f + ms[i].off), zs.avail_in = (uInt)ms[i].size;
            zs.next_out = (Bytef *)&out[ms[i].uoff], zs.avail_out = ms[i].ulen;
            int r = ms[i].ulen ? inflate(&zs, Z_FINISH) : Z_STREAM_END;
            if (r != Z_STREAM_END || (ms[i].ulen && zs.total_out != ms[i].ulen)) bad = true;
            inflateEnd(&zs);
        }
    };
    int nt = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), ms.size());
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    return !bad;
}

bool read_text_maybe_gz(const std::string &path, std::string &out, std::string &err)
{
    {
        MappedFile mf;
        std::string e2;
        if (mf.open(path, e2) && mf.size >= 28 && read_indexed_gz(mf.data, mf.size, out)) return true;
        out.clear();
    }
    gzFile g = gzopen(path.c_str(), "rb");  // gzread passes plain text through unchanged
    if (!g) {
        err = "Cannot open file " + path;
        return false;
    }
    gzbuffer(g, 1 << 20);
    out.clear();
    std::vector<char> buf(1 << 22);
    int n;
    while ((n = gzread(g, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), n);
    gzclose(g);
    if (n < 0) {
        err = "read error on " + path;
        return false;
    }
    return true;
}

static int gz_level()
{
    // The reference's ogzstream uses zlib's default level; parity is defined on the DECOMPRESSED bytes, and the files are
    // read back exactly once (by bwa and by getsv). Default (-1): the Huffman-only writer below, ~8x faster than zlib
    // level 1 for ~20 % larger files (reads and qualities have few LZ77 matches to offer anyway).
    // SEEKSV_B200_GZ_LEVEL=0..9 selects zlib at that level.
    const char *e = getenv("SEEKSV_B200_GZ_LEVEL");
    if (!e) return -1;
    int l = atoi(e);
    return l < 0 || l > 9 ? -1 : l;
}

// ---- Huffman-only DEFLATE writer (RFC 1951 dynamic blocks with literals only) ------------------------------------------
namespace {

// code lengths (<= limit) of an optimal prefix code; frequencies are flattened until the longest code fits
void huffman_lengths(const uint32_t *freq_in, int n, int limit, uint8_t *len)
{
    std::vector<uint32_t> freq(freq_in, freq_in + n);
    for (;;) {
        struct Node {
            uint64_t w;
            int parent;
        };
        std::vector<int> leaves;
        for (int i = 0; i < n; ++i) {
            len[i] = 0;
            if (freq[i]) leaves.push_back(i);
        }
        if (leaves.empty()) return;
        if (leaves.size() == 1) {
            len[leaves[0]] = 1;
            return;
        }
        std::sort(leaves.begin(), leaves.end(), [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
        const int m = (int)leaves.size();
        std::vector<Node> node(2 * m - 1);
        for (int i = 0; i < m; ++i) node[i] = Node{freq[leaves[i]], -1};
        int qa = 0, qb = m, end = m;  // two-queue merge: leaves [qa, m), internal nodes [qb, end)
        auto take = [&]() {
            if (qa < m && (qb >= end || node[qa].w <= node[qb].w)) return qa++;
            return qb++;
        };
        while (end < 2 * m - 1) {
            int a = take(), b = take();
            node[end] = Node{node[a].w + node[b].w, -1};
            node[a].parent = node[b].parent = end;
            ++end;
        }
        int longest = 0;
        std::vector<int> depth(2 * m - 1, 0);
        for (int i = 2 * m - 3; i >= 0; --i) depth[i] = depth[node[i].parent] + 1;
        for (int i = 0; i < m; ++i) {
            len[leaves[i]] = (uint8_t)std::min(depth[i], 255);
            longest = std::max(longest, depth[i]);
        }
        if (longest <= limit) return;
        for (int i = 0; i < n; ++i)
            if (freq[i]) freq[i] = (freq[i] + 1) / 2;
    }
}

// canonical codes, bit-reversed for LSB-first output
void canonical_codes(const uint8_t *len, int n, uint16_t *code)
{
    uint32_t count[16] = {0}, next[16] = {0};
    for (int i = 0; i < n; ++i) count[len[i]]++;
    count[0] = 0;
    uint32_t c = 0;
    for (int b = 1; b <= 15; ++b) {
        c = (c + count[b - 1]) << 1;
        next[b] = c;
    }
    for (int i = 0; i < n; ++i) {
        uint32_t l = len[i], v = l ? next[l]++ : 0, r = 0;
        for (uint32_t k = 0; k < l; ++k) r |= ((v >> k) & 1u) << (l - 1 - k);
        code[i] = (uint16_t)r;
    }
}

struct BitWriter {
    uint8_t *p;
    uint64_t acc = 0;
    int nb = 0;
    inline void put(uint32_t v, int n)  // n <= 32
    {
        acc |= (uint64_t)v << nb;
        nb += n;
        if (nb >= 32) {
            uint32_t w = (uint32_t)acc;
            memcpy(p, &w, 4);
            p += 4;
            acc >>= 32;
            nb -= 32;
        }
    }
    uint8_t *finish()
    {
        while (nb > 0) {
            *p++ = (uint8_t)acc;
            acc >>= 8;
            nb -= 8;
        }
        nb = 0;
        return p;
    }
};

// fixed, complete code for the code-length alphabet: 13 symbols of 4 bits, 6 of 5 bits (Kraft sum 1)
const uint8_t kClLen[19] = {4, 5, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 4, 4};
const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// one deflate block per 64 KiB piece of the input; returns the end of the output
uint8_t *huffman_deflate(const uint8_t *in, size_t n, uint8_t *out)
{
    const size_t PIECE = 64 << 10;
    BitWriter bw{out};
    uint16_t cl_code[19];
    canonical_codes(kClLen, 19, cl_code);
    size_t a = 0;
    do {
        const size_t b = std::min(n, a + PIECE);
        uint32_t freq[257] = {0}, f1[256] = {0}, f2[256] = {0}, f3[256] = {0};
        {   // four interleaved histograms: runs of one symbol do not serialise on a single counter
            size_t i = a;
            for (; i + 4 <= b; i += 4) freq[in[i]]++, f1[in[i + 1]]++, f2[in[i + 2]]++, f3[in[i + 3]]++;
            for (; i < b; ++i) freq[in[i]]++;
            for (int k = 0; k < 256; ++k) freq[k] += f1[k] + f2[k] + f3[k];
        }
        freq[256] = 1;
        uint8_t len[258];
        huffman_lengths(freq, 257, 15, len);
        if (b == a) len[0] = 1;  // empty piece: end-of-block alone would be a one-symbol code; add an unused partner
        len[257] = 1;            // the single (unused) distance code
        uint16_t code[257];
        canonical_codes(len, 257, code);
        bw.put(b == n ? 1 : 0, 1);
        bw.put(2, 2);    // dynamic Huffman
        bw.put(0, 5);    // HLIT: 257 literal/length codes
        bw.put(0, 5);    // HDIST: 1 distance code
        bw.put(15, 4);   // HCLEN: 19 code-length codes
        for (int i = 0; i < 19; ++i) bw.put(kClLen[kClOrder[i]], 3);
        for (int i = 0; i < 258;) {
            if (len[i] == 0) {
                int r = 1;
                while (i + r < 258 && len[i + r] == 0 && r < 138) ++r;
                if (r >= 11) {
                    bw.put(cl_code[18], kClLen[18]);
                    bw.put(r - 11, 7);
                } else if (r >= 3) {
                    bw.put(cl_code[17], kClLen[17]);
                    bw.put(r - 3, 3);
                } else {
                    r = 1;
                    bw.put(cl_code[0], kClLen[0]);
                }
                i += r;
            } else {
                bw.put(cl_code[len[i]], kClLen[len[i]]);
                ++i;
            }
        }
        uint32_t packed[256];  // code | length << 16
        for (int i = 0; i < 256; ++i) packed[i] = code[i] | (uint32_t)len[i] << 16;
        size_t i = a;
        for (; i + 2 <= b; i += 2) {  // two symbols per flush check: <= 30 bits
            uint32_t e0 = packed[in[i]], e1 = packed[in[i + 1]];
            uint32_t l0 = e0 >> 16;
            bw.put((e0 & 0xffff) | (e1 & 0xffff) << l0, (int)(l0 + (e1 >> 16)));
        }
        if (i < b) bw.put(packed[in[i]] & 0xffff, (int)(packed[in[i]] >> 16));
        bw.put(code[256], len[256]);
        a = b;
    } while (a < n);
    return bw.finish();
}

bool huffman_gz_member(const char *data, size_t n, std::vector<uint8_t> &out)
{
    // worst case of a 15-bit-limited code on a 64 KiB piece stays below 2 bytes per input byte; the scratch is per thread and
    // only the used part is copied out (a vector resize would zero two megabytes per member)
    static thread_local std::unique_ptr<uint8_t[]> scratch;
    static thread_local size_t scratch_cap = 0;
    const size_t need = n * 2 + (n / (64 << 10) + 1) * 512 + 64;
    if (scratch_cap < need) scratch.reset(new uint8_t[need]), scratch_cap = need;
    uint8_t *buf = scratch.get();
    static const uint8_t head[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 3, 8, 0, 'S', 'V', 4, 0, 0, 0};
    memcpy(buf, head, 18);
    uint8_t *e = huffman_deflate((const uint8_t *)data, n, buf + 20);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const Bytef *)data, (uInt)n), isz = (uint32_t)n;
    for (int i = 0; i < 4; ++i) *e++ = (crc >> (8 * i)) & 0xff;
    for (int i = 0; i < 4; ++i) *e++ = (isz >> (8 * i)) & 0xff;
    const uint32_t sz = (uint32_t)(e - buf);
    for (int i = 0; i < 4; ++i) buf[16 + i] = (sz >> (8 * i)) & 0xff;
    out.assign(buf, e);
    return true;
}

}  // namespace

static bool gz_member(const char *data, size_t n, std::vector<uint8_t> &out)
{
    if (gz_level() < 0) return huffman_gz_member(data, n, out);
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (deflateInit2(&zs, gz_level(), Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    // extra sub-field 'SV' (ignored by every gzip reader): total size of this member, so that our own reader can find the
    // member boundaries and inflate the members in parallel - the idea of BGZF's 'BC' field with a 32-bit size
    gz_header head;
    memset(&head, 0, sizeof head);
    static const Bytef extra_proto[8] = {'S', 'V', 4, 0, 0, 0, 0, 0};
    Bytef extra[8];
    memcpy(extra, extra_proto, 8);
    head.extra = extra, head.extra_len = 8, head.os = 3;
    deflateSetHeader(&zs, &head);
    out.resize(deflateBound(&zs, n) + 64);
    zs.next_in = (Bytef *)data;
    zs.avail_in = (uInt)n;
    zs.next_out = out.data();
    zs.avail_out = (uInt)out.size();
    int r = deflate(&zs, Z_FINISH);
    out.resize(zs.total_out);
    deflateEnd(&zs);
    if (r != Z_STREAM_END || out.size() < 20) return false;
    uint32_t sz = (uint32_t)out.size();
    for (int i = 0; i < 4; ++i) out[16 + i] = (sz >> (8 * i)) & 0xff;  // header(10) XLEN(2) 'S' 'V' LEN(2) | size
    return true;
}

// The reference writes through ogzstream (gzstream.C:53-61); here every file is cut into parts that are compressed as
// independent gzip members by one thread pool over ALL files (zcat / gzread concatenate members transparently). The default
// (Huffman-only) writer makes members of 64 KiB of text, like the device writer (csrc/gzip.cu): such a file can be inflated by the
// GPU's BGZF inflate kernel (svb_read_gz_device: getsv reads P.clip.gz that way - in a multi-GPU run the file is rank 0's
// concatenation of every rank's block files, a quarter of a second of host inflate at 8 x C2). With an explicit zlib level the
// members stay at 1 MiB (LZ77 matches do not cross members).
bool write_gz_many(const std::vector<GzJob> &jobs, int n_threads, std::string &err)
{
    const uint64_t PART = gz_level() < 0 ? (64ull << 10) : (1ull << 20);
    struct Part {
        size_t job;
        uint64_t a, b;
    };
    std::vector<Part> parts;
    for (size_t j = 0; j < jobs.size(); ++j) {
        uint64_t n = jobs[j].n;
        if (n == 0) parts.push_back(Part{j, 0, 0});
        for (uint64_t a = 0; a < n; a += PART) parts.push_back(Part{j, a, std::min(n, a + PART)});
    }
    std::vector<std::vector<uint8_t>> comp(parts.size());
    std::unique_ptr<std::atomic<int>[]> ready(new std::atomic<int>[parts.size()]);  // 0 pending, 1 compressed, -1 failed
    for (size_t i = 0; i < parts.size(); ++i) ready[i].store(0, std::memory_order_relaxed);
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= parts.size()) return;
            bool ok = gz_member(jobs[parts[i].job].data + parts[i].a, (size_t)(parts[i].b - parts[i].a), comp[i]);
            ready[i].store(ok ? 1 : -1, std::memory_order_release);
        }
    };
    // one writer per file, appending each member as soon as it is compressed (parts are handed out in file order)
    std::vector<std::string> errs(jobs.size());
    std::vector<size_t> first(jobs.size() + 1, parts.size());
    for (size_t i = parts.size(); i-- > 0;) first[parts[i].job] = i;
    for (size_t j = jobs.size(); j-- > 0;)
        if (first[j] == parts.size()) first[j] = first[j + 1];
    auto write_one = [&](size_t j) {
        FILE *f = fopen(jobs[j].path.c_str(), "wb");
        if (!f) errs[j] = "Cannot open file " + jobs[j].path;
        for (size_t i = first[j]; i < parts.size() && parts[i].job == j; ++i) {
            int r;
            while ((r = ready[i].load(std::memory_order_acquire)) == 0) std::this_thread::sleep_for(std::chrono::microseconds(50));
            if (r < 0 && errs[j].empty()) errs[j] = "gzip compression failed";
            if (f && errs[j].empty() && fwrite(comp[i].data(), 1, comp[i].size(), f) != comp[i].size()) errs[j] = "write error on " + jobs[j].path;
            std::vector<uint8_t>().swap(comp[i]);
        }
        if (f) fclose(f);
    };
    std::vector<std::thread> writers, th;
    for (size_t j = 0; j < jobs.size(); ++j) writers.emplace_back(write_one, j);
    int nt = (int)std::min<size_t>((size_t)std::max(1, n_threads), parts.size());
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    for (auto &t : writers) t.join();
    for (auto &e : errs)
        if (!e.empty()) {
            err = e;
            return false;
        }
    return true;
}

bool gz_on_host()
{
    const char *m = getenv("SEEKSV_B200_GZ");
    return gz_level() >= 0 || (m && !strcmp(m, "host"));
}

bool write_files(const std::vector<GzJob> &jobs, std::string &err)
{
    std::vector<std::string> errs(jobs.size());
    auto write_one = [&](size_t j) {
        FILE *f = fopen(jobs[j].path.c_str(), "wb");
        if (!f) {
            errs[j] = "Cannot open file " + jobs[j].path;
            return;
        }
        if (jobs[j].n && fwrite(jobs[j].data, 1, jobs[j].n, f) != jobs[j].n) errs[j] = "write error on " + jobs[j].path;
        fclose(f);
    };
    std::vector<std::thread> th;
    for (size_t j = 1; j < jobs.size(); ++j) th.emplace_back(write_one, j);
    if (!jobs.empty()) write_one(0);
    for (auto &t : th) t.join();
    for (auto &e : errs)
        if (!e.empty()) {
            err = e;
            return false;
        }
    return true;
}

bool write_gz(const std::string &path, const char *data, uint64_t n, int n_threads, std::string &err)
{
    return write_gz_many({GzJob{path, data, n}}, n_threads, err);
}
