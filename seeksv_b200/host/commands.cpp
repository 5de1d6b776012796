// The seeksv command line (seeksv.cpp:26-457 of the reference) over the seeksv_b200 C ABI: same commands,
// option strings, defaults, positional arguments, file names and exit codes. Everything that walks BAM
// records happens on the GPU (include/seeksv_b200.h); this file only parses arguments, reads/writes the
// text files and runs the per-junction bookkeeping (junction.cpp).
#include <getopt.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <map>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <future>
#include <mutex>
#include <spawn.h>
#include <sys/wait.h>
#include <sys/stat.h>
#include <cerrno>
#include <thread>

#include "../../include/seeksv_b200.h"
#include "bamfile.h"
#include "junction.h"

using namespace svb;

extern char **environ;

namespace {
const char *kVersion = "1.2.3";  // behaviour of the reference version this is a drop-in for (seeksv.cpp:12)

int n_threads()
{
    const char *e = getenv("SEEKSV_B200_THREADS");
    int n = e ? atoi(e) : 0;
    return n > 0 ? n : (int)std::max(1u, std::thread::hardware_concurrency());
}
int device_index()
{
    const char *e = getenv("SEEKSV_B200_DEVICE");
    return e ? atoi(e) : 0;
}

struct Gpu {  // context + error-to-exit plumbing: every failure is "message on stderr, exit status 1" as in the reference
    svb_ctx *ctx = nullptr;
    bool open()
    {
        // One context per device for the life of the process: a program that calls svb_main repeatedly (bench.py, a
        // pipeline driver) keeps the pinned staging pool and the stream-ordered device pool warm between commands. The
        // CLI binary runs one command per process, so for it this is simply "create".
        static std::map<int, svb_ctx *> cached;
        int dev = device_index();
        auto it = cached.find(dev);
        if (it != cached.end()) ctx = it->second;
        else {
            if (svb_ctx_create(dev, &ctx) != 0) {
                std::cerr << "[seeksv_b200] " << svb_last_error(nullptr) << std::endl;
                return false;
            }
            cached[dev] = ctx;
        }
        svb_prof_enable(ctx, getenv("SEEKSV_B200_PROFILE") ? 1 : 0);
        svb_prof_reset(ctx);
        return true;
    }
    ~Gpu()
    {
        if (ctx && getenv("SEEKSV_B200_PROFILE")) {
            const char *names[64];
            double ms[64], bytes[64];
            int64_t launches[64];
            int n = svb_prof_read(ctx, 64, names, ms, launches, bytes);
            for (int i = 0; i < n && i < 64; ++i)
                fprintf(stderr, "[prof] %-24s %9.3f ms %6lld launches %8.1f GB/s\n", names[i], ms[i], (long long)launches[i],
                        ms[i] > 0 ? bytes[i] / ms[i] / 1e6 : 0.0);
        }
    }
};

// ---- BAM residency across commands -----------------------------------------------------
// getclip and getsv of one sample read the same BAM with an external aligner run in between. Inside `seeksv run`
// (several commands in ONE process) a BAM that a command has loaded stays in HBM - uncompressed stream, chunk table and
// all - and the next command that names the same file (same path, size and mtime) takes it over instead of reading,
// uploading and inflating it again. Outside `run` nothing is kept.
struct Resident {
    std::string key;
    svb_bam *bam;
};
bool g_keep_resident = false;
std::vector<Resident> g_resident;  // most recently released last; at most two (tumour + normal)
std::mutex g_resident_mutex;       // (getsv opens its BAM from a helper thread)

std::string bam_key(const std::string &path)
{
    struct stat st;
    if (stat(path.c_str(), &st) != 0) return std::string();
    char *rp = realpath(path.c_str(), nullptr);
    std::string key = std::string(rp ? rp : path.c_str()) + "|" + std::to_string((long long)st.st_size) + "|" +
                      std::to_string((long long)st.st_mtim.tv_sec) + "." + std::to_string((long long)st.st_mtim.tv_nsec);
    free(rp);
    return key;
}

int open_bam(svb_ctx *ctx, const std::string &path, svb_bam **out)
{
    if (g_keep_resident) {
        std::string key = bam_key(path);
        std::lock_guard<std::mutex> lock(g_resident_mutex);
        for (size_t i = 0; i < g_resident.size(); ++i)
            if (!key.empty() && g_resident[i].key == key) {
                *out = g_resident[i].bam;
                g_resident.erase(g_resident.begin() + i);
                return 0;
            }
    }
    return svb_bam_open(ctx, path.c_str(), n_threads(), out);
}

void release_bam(const std::string &path, svb_bam *bam)
{
    if (!bam) return;
    std::string key = g_keep_resident ? bam_key(path) : std::string();
    if (key.empty()) {
        svb_bam_free(bam);
        return;
    }
    std::lock_guard<std::mutex> lock(g_resident_mutex);
    g_resident.push_back(Resident{key, bam});
    while (g_resident.size() > 2) {
        svb_bam_free(g_resident.front().bam);
        g_resident.erase(g_resident.begin());
    }
}

void drop_resident()
{
    for (auto &r : g_resident) svb_bam_free(r.bam);
    g_resident.clear();
}

void usage_top()
{
    std::cerr << "Program: seeksv (a tool for structural variation detection and virus integration detection)" << '\n'
              << "Version: " << kVersion << '\n'
              << "Contact: Kunlong Qiu(290832867@qq.com)\n\n"
              << "Usage: seeksv <command> [options]\n\n"
              << "Command: getclip\tget soft-clipped reads\n"
              << "         getsv  \tget final sv\n"
              << "         somatic\tget somatic sv" << std::endl;
}

// not part of the reference's surface (seeksv.cpp:60-72 is reproduced byte for byte above): shown by `seeksv run` / `seeksv run --help`
void usage_run()
{
    std::cerr << "Usage: seeksv run -- <command> [-- <command> ...]\n\n"
              << "Several commands in one process; BAMs stay resident on the GPU in between. getclip / getsv / somatic segments\n"
              << "take the arguments of the commands of the same name, any other segment is executed as an external command\n"
              << "(e.g. the aligner): seeksv run -- getclip -o P in.bam -- 'bwa mem ref.fa P.clip.fq.gz > P.clip.sam' -- getsv ..." << std::endl;
}

// Usage(argv[0], argv[1], i) of the reference's Call* functions (seeksv.cpp:146,188,193,384,389,432) runs on the argument vector
// SelectStep has already shifted by one (seeksv.cpp:448-452), so a wrong arity prints "Usage: getclip -o [options] ...": kept.
void usage_cmd(const char *prog, const char *command, int i)
{
    switch (i) {
    case 0:
        std::cerr << "Usage: " << prog << " " << command << " [options] <input.sorted.bam>\n\n"
                  << "Options: -t <double>           Threshold of match rate while combining two soft-clipped reads [0.9]\n"
                  << "         -q <int>              Minimum mapping quality of soft-clipped reads [1]\n"
                  << "         -s                    Save the low quality sequence clipped before alignment by bwa.\n"
                  << "         -o <string>           Prefix of output files [output]" << std::endl;
        break;
    case 1:
        std::cerr << "Usage: " << prog << " " << command
                  << " [options] <input clipped sequence bam> <input orignal sorted bam> <soft-clipped reads file(*clip.gz)> <output SVs> "
                     "<output unmaped clipped sequence fastq>\n"
                  << "Options: -F <FILE>             Samfile/Bamfile of connected readthrough reads\n"
                  << "         -t <double>           Threshold of match rate while combining two soft-clipped reads [0.9]\n"
                  << "         -l <int>              Maximum search length to find microhomology[50]\n"
                  << "         -q <int>              Minimum mapping quality of discordant read pair [20]\n"
                  << "         -Q <int>              Minimum mapping quality of clipped sequences [1], if you use bwa samse to align the reads, please set\n"
                  << "                               this flag to 20\n"
                  << "         -w <int>              Minimum mapping quality of connected  readthrough reads [1]\n"
                  << "         -n <int>              Number of segment(read pairs) used to calculate insert size default [5000000], if you "
                     "donot want to use abnormal read pairs to call sv, set this parameter to 0.\n"
                  << "         -b <int>              Minimum number of soft clipping read,it's the sum of left clipped reads and right "
                     "clipped reads [3]\n"
                  << "         -d <int>              Minimum distance between the clipped sequence position and the aligned sequence "
                     "position [50]\n"
                  << "         -D                    Do not calculate depth of the breakpoints and their ajacency regions\n"
                  << "         -e <int>              Minimum number of read pairs which support the junction [0]\n"
                  << "         -f <int>              Minimun mutation frequency(left_pos_clip_percentage >= 0.1 or "
                     "right_pos_clip_percentage >= 0.1) [0.1].\n"
                  << "                               If you set -D, this value is invalid and set to [0].\n"
                  << "         -T <int>              Maximum length of microhomology, SV with microhomology length longer than [50] will "
                     "be filtered\n"
                  << "         -m <int>              Minimum length of up_seq or down_seq near the breakpoint when abnormal_read_pair_no "
                     "== 0 [30]\n"
                  << "         -i <int>              Maximum indel number of up_seq or down_seq near the breakpoint when "
                     "abnormal_read_pair_no == 0 [1]\n"
                  << "         -L <int>              Calculate average depth of default [200] bp upstream or downstream of the breakpoints\n"
                  << std::endl;
        break;
    case 2:
        std::cerr << "Usage: " << prog << " " << command
                  << " [options] <input normal original bam> <input normal soft-clipped reads file(*.clip.gz)> <input tumor SV file> <output "
                     "somatic SV file>\n\n"
                  << "         -t <int>              Threshold of match rate while comparing two soft-clipped reads [0.9]\n"
                  << "         -q <int>              Minimum mapping quality of discordant read pair [20]\n"
                  << "         -l <int>              Maximum search length to find microhomology [30]\n"
                  << "         -m <int>              Minimum length of the clipped sequence  in normal [10]\n"
                  << "         -n <int>              Number of segment(read pairs) used to calculate insert size default [5000000], if you "
                     "donot want to use abnormal read pairs to call sv, set this parameter to 0."
                  << std::endl;
        break;
    }
}

// wall-clock phase log on stderr when SEEKSV_B200_TIMING is set (for profiling the end-to-end path)
struct Phase {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    bool on = getenv("SEEKSV_B200_TIMING") != nullptr;
    void mark(const char *what)
    {
        auto t1 = std::chrono::steady_clock::now();
        if (on) fprintf(stderr, "[time] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

int fail(const std::string &msg)
{
    std::cerr << msg << std::endl;
    return 1;
}

// ---- getclip: CallGetclip (seeksv.cpp:128-155) + InputBamOutputReads (clip_reads.h:363-484) ---------------------
int cmd_getclip(int argc, char **argv)
{
    svb_getclip_params prm = {0.9, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    prm.gz_outputs = gz_on_host() ? 0 : 1;  // default: the four files are compressed on the device
    prm.with_rows = g_keep_resident ? 1 : 0;  // `seeksv run`: a getsv of the same BAM follows - one pass over the records serves both
    std::string prefix = "output";
    int c;
    optind = 1;
    while ((c = getopt(argc, argv, "t:q:o:s")) >= 0) {
        switch (c) {
        case 't': prm.match_rate = atof(optarg); break;
        case 'q': prm.min_mapq = atoi(optarg); break;
        case 's': prm.save_low_quality = 1; break;
        case 'o': prefix = optarg; break;
        }
    }
    if (argc != optind + 1) {
        usage_cmd(argv[0], argc > 1 ? argv[1] : "", 0);
        return 1;
    }
    std::string bamfile = argv[optind];
    Phase ph;
    Gpu g;
    if (!g.open()) return 1;
    ph.mark("getclip: context");
    svb_bam *bam = nullptr;
    if (open_bam(g.ctx, bamfile, &bam) != 0) {
        std::cerr << "[main_samview] fail to open file for reading." << std::endl;
        return fail(std::string("[seeksv_b200] ") + svb_last_error(g.ctx));
    }
    ph.mark("getclip: load BAM to HBM");
    svb_clusters *cl = nullptr;
    int rc = svb_getclip(g.ctx, bam, &prm, &cl);
    ph.mark("getclip: device passes");
    if (rc != 0) {
        svb_bam_free(bam);
        return fail(std::string("[seeksv_b200] ") + svb_last_error(g.ctx));
    }
    release_bam(bamfile, bam);  // (kept on the GPU for the next command inside `seeksv run`, freed otherwise)
    bam = nullptr;
    const char *ext[4] = {".clip.gz", ".clip.fq.gz", ".unmapped_1.fq.gz", ".unmapped_2.fq.gz"};
    int status = 0;
    std::vector<GzJob> jobs;
    for (int i = 0; i < 4; ++i) {
        const char *data;
        uint64_t len;
        if (prm.gz_outputs) svb_clusters_gz(cl, i, &data, &len);
        else svb_clusters_text(cl, i, &data, &len);
        jobs.push_back(GzJob{prefix + ext[i], data, len});
    }
    std::string werr;
    if (!(prm.gz_outputs ? write_files(jobs, werr) : write_gz_many(jobs, n_threads(), werr))) status = fail(werr);
    ph.mark("getclip: gzip + write outputs");
    std::cerr << "[GetSClipReads] finished!" << std::endl;
    svb_clusters_free(cl);
    return status;
}

// clip.bam (BGZF) or SAM text -> alignment list + reference names
// The join of the clip lines with the clip alignments (InputSoftInfoStoreBreakpoint, getsv.h:423-541). `getsv` runs it on the GPU
// (svb_clip_join) - on a context of its own, because the command's main context is busy loading the original BAM on a helper
// thread at that moment (one context, one thread). The host keeps tokenising, the order-dependent accumulation and MergeJunction.
// SEEKSV_B200_DEVICE_JOIN=0 selects the host mirror (join_clips_with_alignments); it also takes over when the device refuses the
// input (a run of lines with thousands of alignments), and it is what svb_plan_getsv - documented as "no GPU work" - uses unless
// SEEKSV_B200_DEVICE_JOIN=1 asks for the device there too.
// getsv's second context: the clip files are read and joined on it while the command's own context loads the BAM on a helper
// thread (one context, one thread at a time). Kept for the life of the process, like the command context.
svb_ctx *side_context(std::string &err)
{
    static std::map<int, svb_ctx *> cached;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    const int dev = device_index();
    auto it = cached.find(dev);
    if (it != cached.end()) return it->second;
    svb_ctx *ctx = nullptr;
    if (svb_ctx_create(dev, &ctx) != 0) {
        err = svb_last_error(nullptr);
        return nullptr;
    }
    cached[dev] = ctx;
    return ctx;
}

bool device_join_enabled(bool by_default)
{
    const char *e = getenv("SEEKSV_B200_DEVICE_JOIN");
    return e ? atoi(e) != 0 : by_default;
}
bool join_clips(const std::vector<ClipLine> &lines, const AlignmentSet &set, JunctionMap &jm, std::string &err, bool device_by_default)
{
    if (!device_join_enabled(device_by_default)) {
        join_clips_with_alignments(lines, set, jm);
        return true;
    }
    svb_ctx *ctx = side_context(err);
    if (!ctx) return false;
    JoinArrays J;
    if (!pack_join_inputs(lines, set, J, n_threads())) {
        join_clips_with_alignments(lines, set, jm);
        return true;
    }
    svb_join_cand *cands = nullptr;
    uint64_t n = 0;
    int rc = svb_clip_join(ctx, J.lines.data(), J.lines.size(), J.seqs.data(), J.seqs.size(), J.alns.data(), J.alns.size(), J.names.data(),
                           J.names.size(), set.cigar_words.data(), set.cigar_words.size(), &cands, &n);
    if (rc == SVB_ERR_FORMAT) {  // refused: a set too large for the per-run sort
        join_clips_with_alignments(lines, set, jm);
        return true;
    }
    if (rc != 0) {
        err = svb_last_error(ctx);
        return false;
    }
    const bool ok = accumulate_join_candidates(lines, set, J, cands, n, jm, err);
    svb_free(cands);
    return ok;
}

bool load_alignments(const std::string &path, AlignmentSet &set, std::string &err)
{
    std::vector<uint8_t> file;
    if (!read_file(path, file, err)) return false;
    if (path.size() >= 4 && path.rfind(".bam") == path.size() - 4) {
        BamHeader h;
        if (!bgzf_inflate_all(file.data(), file.size(), set.storage, n_threads(), err)) return false;
        if (!parse_bam_header(set.storage.data(), set.storage.size(), h, err)) return false;
        set.ref_names = h.names;
        if (!parse_bam_alignments(set, h.first_record)) {
            err = "corrupt alignment records in " + path;
            return false;
        }
        return true;
    }
    set.storage.swap(file);
    return parse_sam_alignments(set, n_threads(), err);
}

struct Win {  // a device window with the chromosome name the host maps are keyed by
    const std::string *chr;
    int begin, end;
    uint64_t off;
};

// begin2end (ordered by chromosome NAME, int begin/end possibly wrapped, quirk Q11) -> windows as the device wants
// them: (tid, begin clamped to >= 1, end clamped to the chromosome), sorted by (tid, begin), disjoint. A window
// whose chromosome is not in the BAM or that is empty after clamping has no covered position.
uint64_t device_windows(const WindowMap &begin2end, const std::map<std::string, int32_t> &tid_of, const std::vector<uint32_t> &lens,
                        std::vector<svb_window> &dw, std::vector<Win> &hw)
{
    std::vector<std::pair<svb_window, const std::string *>> tmp;
    for (auto &kv : begin2end) {
        auto it = tid_of.find(kv.first.first);
        if (it == tid_of.end()) continue;
        int b = std::max(kv.first.second, 1), e = (int)std::min<int64_t>(kv.second, (int64_t)lens[it->second] + 1);
        if (e < b) continue;
        tmp.push_back(std::make_pair(svb_window{it->second, b, e}, &kv.first.first));
    }
    std::sort(tmp.begin(), tmp.end(), [](const auto &a, const auto &b) {
        return a.first.tid != b.first.tid ? a.first.tid < b.first.tid : a.first.begin < b.first.begin;
    });
    for (auto &w : tmp) {
        if (!dw.empty() && dw.back().tid == w.first.tid && w.first.begin <= dw.back().end) {
            dw.back().end = std::max(dw.back().end, w.first.end);
            hw.back().end = dw.back().end;
        } else {
            dw.push_back(w.first);
            hw.push_back(Win{w.second, w.first.begin, w.first.end, 0});
        }
    }
    uint64_t total = 0;
    for (size_t i = 0; i < dw.size(); ++i) {
        hw[i].off = total;
        total += (uint64_t)(dw[i].end - dw[i].begin + 1);
    }
    return total;
}

// Closed form of main_depth's two map walks (bam2depth.cpp:82-124, literal version: svb::account_position). For a
// covered position P the reference (1) finds the LAST merged window whose begin is <= P and requires P <= its end,
// (2) adds depth(P) to every range r on the chromosome with (r.begin, r.end) <= (P+1, P+1), P <= r.end and
// r.begin >= window.begin (unsigned compare), (3) stores depth(P) if P is a junction position - steps (2) and (3) only
// when the range map holds any key <= (chr, P+1, P+1) at all (the `continue` of bam2depth.cpp:102). Turned around per range:
// r collects depth over [r.begin, r.end] plus - only for the degenerate 0/1-length ranges of quirk Q11 - the position
// r.begin - 1, restricted to positions whose window satisfies the unsigned compare. Sums come from prefix sums.
void account_depth(const std::vector<Win> &hw, const std::vector<int32_t> &depth, const WindowMap &begin2end, PosDepth &pos2depth,
                   RangeDepth &range2depth)
{
    struct DevWin {
        int begin, end;
        uint64_t off;
    };
    struct MapWin {
        int begin, end;
        int64_t lo, hi;  // positions for which this entry is "the last window with begin <= P" and P <= end
    };
    std::map<std::string, std::vector<DevWin>> dev;
    for (auto &w : hw) dev[*w.chr].push_back(DevWin{w.begin, w.end, w.off});
    std::vector<uint64_t> prefix(depth.size() + 1, 0);
    for (size_t i = 0; i < depth.size(); ++i) prefix[i + 1] = prefix[i] + (uint64_t)depth[i];
    auto sum = [&](const std::vector<DevWin> &v, int64_t a, int64_t b) -> uint64_t {  // sum of depth over [a, b]
        uint64_t s = 0;
        auto it = std::lower_bound(v.begin(), v.end(), a, [](const DevWin &w, int64_t x) { return (int64_t)w.end < x; });
        for (; it != v.end() && it->begin <= b; ++it) {
            int64_t lo = std::max<int64_t>(a, it->begin), hi = std::min<int64_t>(b, it->end);
            if (lo <= hi) s += prefix[it->off + (uint64_t)(hi - it->begin) + 1] - prefix[it->off + (uint64_t)(lo - it->begin)];
        }
        return s;
    };
    std::map<std::string, std::vector<MapWin>> mw;
    for (auto &kv : begin2end) mw[kv.first.first].push_back(MapWin{kv.first.second, kv.second, 0, -1});
    for (auto &kv : mw) {
        auto &v = kv.second;  // map order: ascending (int) begin
        for (size_t i = 0; i < v.size(); ++i) {
            v[i].lo = std::max<int64_t>(1, v[i].begin);
            v[i].hi = v[i].end;
            if (i + 1 < v.size()) v[i].hi = std::min<int64_t>(v[i].hi, (int64_t)v[i + 1].begin - 1);
        }
    }
    for (auto &kv : range2depth) {
        const ChrRange &r = kv.first;
        auto m = mw.find(r.chr);
        auto d = dev.find(r.chr);
        if (m == mw.end() || d == dev.end()) continue;
        int64_t b = r.begin, e = r.end;  // unsigned values
        uint64_t total = 0;
        auto add = [&](int64_t lo, int64_t hi) {
            if (lo > hi) return;
            for (const MapWin &w : m->second) {
                if (w.hi < lo || w.lo > hi) continue;
                if (!(r.begin >= (unsigned)w.begin)) continue;  // bam2depth.cpp:107, unsigned vs int
                total += sum(d->second, std::max(lo, w.lo), std::min(hi, w.hi));
            }
        };
        if (b <= e && b <= INT32_MAX) add(std::max<int64_t>(b, 1), std::min<int64_t>(e, INT32_MAX));
        if ((e == b - 1 || e == b) && b - 1 >= 1 && b - 1 <= INT32_MAX) add(b - 1, b - 1);
        kv.second += total;
    }
    for (auto &kv : pos2depth) {
        auto m = mw.find(kv.first.first);
        auto d = dev.find(kv.first.first);
        if (m == mw.end() || d == dev.end()) continue;
        int64_t p = kv.first.second;
        // bam2depth.cpp:101-102: no range at or below (chr, P+1, P+1) in the WHOLE map -> `continue` before the point depth
        // is stored (only the first positions of the lexicographically smallest chromosome can be hit)
        if (range2depth.upper_bound(ChrRange{kv.first.first, (unsigned)(p + 1), (unsigned)(p + 1)}) == range2depth.begin()) continue;
        for (const MapWin &w : m->second)
            if (w.lo <= p && p <= w.hi) {
                uint64_t v = sum(d->second, p, p);
                if (v > 0) kv.second = (int)v;  // only covered positions are visited by the pileup
                break;
            }
    }
}

// Multi-GPU runs (python -m seeksv_b200.mgpu getsv / somatic): the ranks run the device passes on their shards of the BAM and
// add the results up (NCCL); rank 0 then runs this command for the host bookkeeping and the output files with
// SEEKSV_B200_SHARD_RESULTS naming a file that holds those results - int32 little endian: 'SVBR', n (records behind the
// statistics), mean, deviation, number of junction counts, the counts, number of depth values, the depth values - in the
// order in which this command asks for them. The BAM is then opened for its header only.
struct ShardResults {
    bool on = false;
    int32_t n = 0, mean = 0, dev = 0;
    std::vector<int32_t> counts, depth;
    bool load()
    {
        const char *path = getenv("SEEKSV_B200_SHARD_RESULTS");
        if (!path || !*path) return true;
        std::ifstream f(path, std::ios::binary);
        std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        if (raw.size() < 24 || memcmp(raw.data(), "SVBR", 4) != 0) return false;
        const int32_t *w = (const int32_t *)raw.data();
        const size_t nw = raw.size() / 4;
        n = w[1], mean = w[2], dev = w[3];
        size_t o = 4;
        const int32_t nc = w[o++];
        if (nc < 0 || o + (size_t)nc + 1 > nw) return false;
        counts.assign(w + o, w + o + nc);
        o += nc;
        const int32_t nd = w[o++];
        if (nd < 0 || o + (size_t)nd > nw) return false;
        depth.assign(w + o, w + o + nd);
        on = true;
        return true;
    }
};
ShardResults g_shard;

// The same hand-over without the file and without a second join: a caller that drives the sharded passes itself registers a
// provider (svb_set_shard_provider); getsv calls it ONCE, after the junction merge, with the junctions and depth windows in the
// order in which it will ask for their results, and goes on with what the provider filled in.
struct ShardProvider {
    svb_shard_provider_fn fn = nullptr;
    void *user = nullptr;
} g_provider;

int open_original(svb_ctx *ctx, const std::string &path, svb_bam **out)
{
    if (g_shard.on) return svb_bam_open_refs(ctx, path.c_str(), nullptr, 0, 0, n_threads(), out);  // header only: no records are loaded
    return open_bam(ctx, path, out);
}

bool insert_size(Gpu &g, svb_bam *bam, const std::string &file, int min_mapq, int pairs_used, int &mean, int &dev)
{
    if (g_shard.on) {
        if (g_shard.n == 0) return true;
        mean = g_shard.mean, dev = g_shard.dev;
        std::cerr << "Bam/sam " << file << "    Mean insert size : " << mean << "\n"
                  << "Mean deviation: " << dev << std::endl;
        return true;
    }
    // CalculateInsertsizeDeviation, cluster.cpp:15-83: integer mean, (int)sqrt of the double mean square
    int64_t st[4];
    if (svb_insert_stats(g.ctx, bam, min_mapq, pairs_used, st) != 0) return false;
    if (st[0] == 0) return true;  // the reference returns early and leaves both at 0
    mean = (int)st[2];
    dev = (int)sqrt((double)st[3] / (double)(int)st[0]);
    std::cerr << "Bam/sam " << file << "    Mean insert size : " << mean << "\n"
              << "Mean deviation: " << dev << std::endl;
    return true;
}

void to_device_junction(const JunctionKey &k, svb_bam *bam, svb_junction &j)
{
    auto tid_of = [&](const std::string &name) {
        for (int32_t t = 0; t < svb_bam_n_ref(bam); ++t)
            if (name == svb_bam_ref_name(bam, t)) return t;  // BamGetTid, cluster.cpp:219-229
        return (int32_t)-1;
    };
    j.up_tid = tid_of(k.up_chr), j.down_tid = tid_of(k.down_chr);
    j.up_pos = k.up_pos, j.down_pos = k.down_pos, j.up_strand = k.up_strand, j.down_strand = k.down_strand;
    j.pad_[0] = j.pad_[1] = 0;
}

// ---- getsv: CallGetsv (seeksv.cpp:157-364) -----------------------------------------------------------------------
int cmd_getsv(int argc, char **argv)
{
    std::string connect_bam, seed_file;
    double frequency = 0.1;
    int c, flank = 50, min_mapq = 20, pairs_used = 5000000, min_clip_sum = 3, min_distance = 50, max_micro = 50, times = 4, min_pairs = 0,
           flank_len = 200, min_seq_len = 30, max_indel = 1, connect_min_mapq = 1;
    bool with_depth = true;
    optind = 1;
    while ((c = getopt(argc, argv, "F:B:t:l:q:Q:w:n:a:b:d:e:m:i:R:f:T:L:rD")) >= 0) {
        switch (c) {
        case 'F': connect_bam = optarg; break;
        case 'w': connect_min_mapq = atoi(optarg); break;
        case 'B': seed_file = optarg; break;
        case 'l': flank = atoi(optarg); break;
        case 'q': min_mapq = atoi(optarg); break;
        case 'n': pairs_used = atoi(optarg); break;
        case 'b': min_clip_sum = atoi(optarg); break;
        case 'd': min_distance = atoi(optarg); break;
        case 'e': min_pairs = atoi(optarg); break;
        case 'm': min_seq_len = atoi(optarg); break;
        case 'i': max_indel = atoi(optarg); break;
        case 'D': with_depth = false; break;
        case 'f': frequency = atof(optarg); break;
        case 'T': max_micro = atoi(optarg); break;
        case 'L': flank_len = atoi(optarg); break;
        default: break;  // -t -Q -a -R -r are parsed and unused, as in the reference
        }
    }
    if (argc != optind + 5 || flank > 90 || flank < 0 || min_seq_len < 0) {
        usage_cmd(argv[0], argc > 1 ? argv[1] : "", 1);
        return 1;
    }

    std::string clip_aln = argv[optind], original_bam = argv[optind + 1], clipfile = argv[optind + 2], sv_file = argv[optind + 3],
                unmapped_file = argv[optind + 4];
    std::string err, clip_text;
    AlignmentSet alns;
    Phase ph;
    // The insert-size pass always needs the original BAM: its load (file -> HBM, inflate on the device) runs on a helper
    // thread while this thread reads and joins the clip files. Errors are still reported in the reference's order.
    Gpu g;
    svb_bam *bam = nullptr;
    struct Prefetch {
        std::future<int> f;
        svb_bam **bam;
        const std::string *path;
        int wait() { return f.valid() ? f.get() : 0; }
        ~Prefetch()
        {
            wait();
            if (*bam) release_bam(*path, *bam), *bam = nullptr;
        }
    } prefetch{{}, &bam, &original_bam};
    bool load_failed = false;
    if (pairs_used >= 100000)
        prefetch.f = std::async(std::launch::async, [&]() -> int {
            if (!g.open()) return 1;
            return open_original(g.ctx, original_bam, &bam) != 0 ? 2 : 0;
        });
    // the two clip files are read side by side (the text inflates while the alignments are parsed); errors in the reference's order
    // P.clip.gz as our getclip writes it (gzip members of <= 64 KiB of text) is inflated on the GPU, on the side context: the host
    // cores are busy staging the BAM just now. Anything else (a reference-made file, host-written members, plain text) goes
    // through the host reader as before. SEEKSV_B200_GZ_READ=host|zlib keeps everything on the host.
    std::string clip_err;
    const char *clip_data = nullptr;
    uint64_t clip_size = 0;
    const bool gpu_command = pairs_used >= 100000 || with_depth;
    std::future<bool> clip_read = std::async(std::launch::async, [&]() {
        const char *mode = getenv("SEEKSV_B200_GZ_READ");
        if (gpu_command && !mode) {
            std::string e2;
            svb_ctx *side = side_context(e2);
            if (side && svb_read_gz_device(side, clipfile.c_str(), &clip_data, &clip_size) == 0) return true;
            clip_data = nullptr, clip_size = 0;
        }
        if (!read_text_maybe_gz(clipfile, clip_text, clip_err)) return false;
        clip_data = clip_text.data(), clip_size = clip_text.size();
        return true;
    });
    if (!load_alignments(clip_aln, alns, err)) {
        clip_read.wait();
        std::cerr << "[main_samview] fail to open file for reading." << std::endl;
        return fail(err);
    }
    ph.mark("getsv: read clip alignments");
    if (!clip_read.get()) return fail(clip_err);
    ph.mark("getsv: read clip.gz");
    JunctionMap jm;
    if (!seed_file.empty()) {  // ReadBreakpoint, seeksv.cpp:215-219
        std::string seed_text;
        std::ifstream sf(seed_file.c_str());
        if (!sf) std::cerr << "Cannot open file " << seed_file << std::endl;  // (the reference goes on without the seeds)
        else {
            seed_text.assign(std::istreambuf_iterator<char>(sf), std::istreambuf_iterator<char>());
            read_breakpoints(seed_text, jm);
        }
        std::cerr << "[ReadBreakpoint] finish" << std::endl;
    }
    if (!connect_bam.empty()) {  // FindJunction, seeksv.cpp:221-225
        std::vector<uint8_t> file, stream;
        BamHeader ch;
        std::string cerr_text;
        bool ok = read_file(connect_bam, file, cerr_text);
        if (ok && connect_bam.size() >= 4 && connect_bam.rfind(".bam") == connect_bam.size() - 4)
            ok = bgzf_inflate_all(file.data(), file.size(), stream, n_threads(), cerr_text) && parse_bam_header(stream.data(), stream.size(), ch, cerr_text);
        else if (ok)
            ok = sam_to_bam_stream(file, ch, stream, cerr_text);
        if (!ok) {
            std::cerr << "[main_samview] fail to open file for reading." << std::endl;
            return fail(cerr_text);
        }
        find_junctions(stream.data(), stream.size(), ch.first_record, ch.names, connect_min_mapq, jm);
        std::cerr << "'FindJunction' finished" << std::endl;
    }
    {
        std::vector<ClipLine> lines = parse_clip_text(clip_data, (size_t)clip_size, n_threads());
        ph.mark("getsv: tokenise clip lines");
        // (`getsv -n 0 -D` has no BAM pass and runs without a GPU: the host mirror joins there)
        if (!join_clips(lines, alns, jm, err, gpu_command)) return fail("[seeksv_b200] " + err);
    }
    std::cerr << "'InputSoftInfoStoreBreakpoint' finished" << std::endl;
    ph.mark("getsv: join");
    merge_junctions(jm, flank);
    ph.mark("getsv: merge junctions");

    auto need_bam = [&]() -> bool {
        int rc = prefetch.wait();
        if (rc == 1 || load_failed) return false;
        if (rc == 0 && bam) return true;
        if (rc == 0) {
            if (!g.ctx && !g.open()) return false;
            rc = open_original(g.ctx, original_bam, &bam) != 0 ? 2 : 0;
        }
        if (rc != 0) {
            load_failed = true;
            std::cerr << "[main_samview] fail to open file for reading." << std::endl;
            std::cerr << "[seeksv_b200] " << svb_last_error(g.ctx) << std::endl;
            return false;
        }
        return true;
    };
    if (g_provider.fn) {  // sharded run: the ranks' passes are driven by the caller, this command keeps the bookkeeping
        if (!need_bam()) return 1;  // (header only)
        std::vector<svb_junction> pj;
        if (pairs_used >= 100000) {
            pj.resize(jm.size());
            size_t k = 0;
            for (auto &kv : jm) to_device_junction(kv.first, bam, pj[k++]);
        }
        std::vector<svb_window> pw;
        uint64_t total = 0;
        if (with_depth) {
            PosDepth p2d;
            RangeDepth r2d;
            WindowMap b2e;
            JunctionRanges jr;
            collect_breaks(jm, flank_len, p2d, r2d, jr);
            merge_ranges(r2d, b2e);
            std::vector<Win> hw;
            std::map<std::string, int32_t> tid_of;
            std::vector<uint32_t> lens;
            for (int32_t t = 0; t < svb_bam_n_ref(bam); ++t) {
                tid_of.insert(std::make_pair(std::string(svb_bam_ref_name(bam, t)), t));
                lens.push_back(svb_bam_ref_len(bam, t));
            }
            total = device_windows(b2e, tid_of, lens, pw, hw);
        }
        g_shard.counts.assign(pj.size(), 0), g_shard.depth.assign(total, 0);
        int64_t st[3] = {0, 0, 0};
        if (g_provider.fn(pj.data(), pj.size(), pw.data(), pw.size(), min_mapq, pairs_used, times, st, g_shard.counts.data(),
                          g_shard.depth.data(), g_provider.user) != 0)
            return fail("[seeksv_b200] the shard provider failed");
        g_shard.n = (int32_t)std::min<int64_t>(st[0], INT32_MAX), g_shard.mean = (int32_t)st[1], g_shard.dev = (int32_t)st[2];
        ph.mark("getsv: sharded passes (provider)");
    }
    int mean = 0, dev = 0;
    if (pairs_used >= 100000) {
        if (!need_bam()) return 1;
        ph.mark("getsv: load BAM to HBM");
        if (!insert_size(g, bam, original_bam, min_mapq, pairs_used, mean, dev)) return fail(svb_last_error(g.ctx));
        std::cerr << "'CalculateInsertsizeDeviation' finished" << std::endl;
        std::vector<svb_junction> dj(jm.size());
        std::vector<int32_t> counts(jm.size(), 0);
        size_t i = 0;
        for (auto &kv : jm) to_device_junction(kv.first, bam, dj[i++]);
        svb_pair_params pp = {min_mapq, mean, dev, times};
        if (g_shard.on) {
            if (g_shard.counts.size() != counts.size()) return fail("[seeksv_b200] SEEKSV_B200_SHARD_RESULTS: junction count mismatch");
            counts = g_shard.counts;
        } else if (svb_discordant_support(g.ctx, bam, dj.data(), dj.size(), &pp, counts.data()) != 0)
            return fail(svb_last_error(g.ctx));
        i = 0;
        for (auto &kv : jm) kv.second.pairs = counts[i++];
        std::cerr << "'FindDiscordantReadPairs' finished" << std::endl;
    } else
        min_pairs = 0;

    PosDepth pos2depth;
    RangeDepth range2depth;
    WindowMap begin2end;
    JunctionRanges j2r;
    if (with_depth) {
        collect_breaks(jm, flank_len, pos2depth, range2depth, j2r);
        merge_ranges(range2depth, begin2end);
        std::cerr << "'MergeOverlap' finished" << std::endl;
        if (!begin2end.empty()) {
            if (!need_bam()) return 1;
            std::vector<svb_window> dw;
            std::vector<Win> hw;
            std::map<std::string, int32_t> tid_of;
            std::vector<uint32_t> lens;
            for (int32_t t = 0; t < svb_bam_n_ref(bam); ++t) {
                tid_of.insert(std::make_pair(std::string(svb_bam_ref_name(bam, t)), t));
                lens.push_back(svb_bam_ref_len(bam, t));
            }
            uint64_t total = device_windows(begin2end, tid_of, lens, dw, hw);
            std::vector<int32_t> depth(total);
            if (g_shard.on) {
                if (g_shard.depth.size() != depth.size()) return fail("[seeksv_b200] SEEKSV_B200_SHARD_RESULTS: depth length mismatch");
                depth = g_shard.depth;
            } else if (svb_window_depth(g.ctx, bam, dw.data(), dw.size(), min_mapq, depth.data()) != 0)
                return fail(svb_last_error(g.ctx));
            // main_depth visits covered positions in BAM order (tid, pos); the range sums are commutative and every
            // position is written once, so the order of the walk does not matter - but keep it anyway
            if (getenv("SEEKSV_B200_LITERAL_DEPTH_WALK")) {  // the reference's own per-position map walks (cross-check)
                for (size_t i = 0; i < hw.size(); ++i)
                    for (int p = hw[i].begin; p <= hw[i].end; ++p) {
                        int d = depth[hw[i].off + (uint64_t)(p - hw[i].begin)];
                        if (d > 0) account_position(*hw[i].chr, p, d, begin2end, pos2depth, range2depth);
                    }
            } else
                account_depth(hw, depth, begin2end, pos2depth, range2depth);
        }
        std::cerr << "'main_depth' finished" << std::endl;
    } else
        frequency = 0;
    ph.mark("getsv: device passes + depth");

    std::string body, filtered, log;
    OutputFilters f;
    f.min_clip_sum = min_clip_sum, f.min_pairs = min_pairs, f.frequency = frequency, f.min_distance = min_distance;
    f.max_micro = max_micro, f.min_seq_len = min_seq_len, f.max_indel = max_indel;
    write_breakpoints(jm, pos2depth, range2depth, j2r, f, body, filtered, log);
    std::cerr << log;
    std::ofstream fout(sv_file.c_str());
    if (!fout) return fail("Cannot open file " + sv_file);
    fout << kSvHeader << body;
    fout.close();
    std::cout << filtered << std::flush;
    // OutputOneendUnmapBreakpoint (getsv.cpp:1252-1288) never has anything to write (quirk Q7): the file is created empty
    std::ofstream fu(unmapped_file.c_str());
    if (!fu) return fail("Cannot open file " + unmapped_file);
    fu.close();
    return 0;  // (the BAM is released by `prefetch`)
}

// ---- somatic: CallSomatic (seeksv.cpp:366-410) ---------------------------------------------------------------------
int cmd_somatic(int argc, char **argv)
{
    int offset = 30, c, min_len = 10, pairs_used = 5000000, min_mapq = 20;
    double rate = 0.9;
    optind = 1;
    while ((c = getopt(argc, argv, "t:q:l:m:n:")) >= 0) {
        switch (c) {
        case 't': rate = atof(optarg); break;
        case 'q': min_mapq = atoi(optarg); break;
        case 'l': offset = atoi(optarg); break;
        case 'm': min_len = atoi(optarg); break;
        case 'n': pairs_used = atoi(optarg); break;
        }
    }
    if (argc != optind + 4) {
        std::cerr << argc << '\t' << optind << std::endl;
        usage_cmd(argv[0], argc > 1 ? argv[1] : "", 2);
        return 1;
    }
    if (offset >= 90 || offset < 0) {
        std::cerr << "Error: value of -l must in range [0, 90) " << std::endl;
        usage_cmd(argv[0], argc > 1 ? argv[1] : "", 2);
        return 1;
    }
    std::string normal_bam = argv[optind], clipfile = argv[optind + 1], tumor_file = argv[optind + 2], out_file = argv[optind + 3];
    std::string clip_text, tumor_text, err;
    if (!read_text_maybe_gz(clipfile, clip_text, err)) return fail(err);
    Gpu g;
    if (!g.open()) return 1;
    svb_bam *bam = nullptr;
    if (open_original(g.ctx, normal_bam, &bam) != 0) {
        std::cerr << "[main_samview] fail to open file for reading." << std::endl;
        return fail(std::string("[seeksv_b200] ") + svb_last_error(g.ctx));
    }
    int mean = 0, dev = 0;
    if (pairs_used >= 100000 && !insert_size(g, bam, normal_bam, min_mapq, pairs_used, mean, dev)) return fail(svb_last_error(g.ctx));
    std::ofstream fout(out_file.c_str());
    if (!fout) return fail("Error: Cannot open output file " + out_file);
    if (!read_text_maybe_gz(tumor_file, tumor_text, err)) return fail("Error: Cannot open output file " + tumor_file);
    std::vector<SomaticRow> rows;
    std::string log;
    somatic_rows(clip_text, tumor_text, rate, offset, min_len, mean, rows, log);
    std::cerr << log;
    std::vector<svb_junction> dj;
    std::vector<size_t> who;
    for (size_t i = 0; i < rows.size(); ++i)
        if (!rows[i].is_header && rows[i].query_pairs) {
            svb_junction j;
            to_device_junction(rows[i].key, bam, j);
            dj.push_back(j);
            who.push_back(i);
        }
    std::vector<int32_t> counts(dj.size(), 0), per_row(rows.size(), 0);
    svb_pair_params pp = {min_mapq, mean, dev, 4};
    if (g_shard.on) {
        if (g_shard.counts.size() != counts.size()) return fail("[seeksv_b200] SEEKSV_B200_SHARD_RESULTS: junction count mismatch");
        counts = g_shard.counts;
    } else if (!dj.empty() && svb_discordant_support(g.ctx, bam, dj.data(), dj.size(), &pp, counts.data()) != 0)
        return fail(svb_last_error(g.ctx));
    for (size_t k = 0; k < who.size(); ++k) per_row[who[k]] = counts[k];
    for (size_t i = 0; i < rows.size(); ++i) {
        if (rows[i].is_header) fout << rows[i].prefix;
        else fout << rows[i].prefix << '\t' << rows[i].normal_left << '\t' << rows[i].normal_right << '\t' << per_row[i] << '\n';
    }
    fout.close();
    release_bam(normal_bam, bam);
    return 0;
}
}  // namespace

extern "C" void svb_set_shard_provider(svb_shard_provider_fn fn, void *user) { g_provider.fn = fn, g_provider.user = user; }

// One rank's part of the multi-GPU getclip outputs. A shard's clip text (coordinate range or whole chromosomes of a sorted BAM) is
// a sequence of (chromosome, side) blocks, and the whole-file text is a concatenation of the shards' blocks: per chromosome the
// '5' blocks of all shards, then the '3' blocks (DisplaySClipReadsAndClipFq flushes per chromosome, clip_reads.h:300-345,423-438).
// So every rank compresses ITS blocks into gzip members of their own files (<prefix>.<b>.clip.gz / .fq.gz) - gzip members
// concatenate - and the merging rank only orders files; no text travels and nothing is compressed twice.
// *blocks: one line "chromosome<TAB>side" per block, malloc'ed (svb_free).
extern "C" int svb_write_range_blocks(const char *part_prefix, const void *clip, uint64_t n_clip, const void *fq, uint64_t n_fq, int n_threads,
                                      char **blocks, uint64_t *blocks_len)
{
    if (!part_prefix || !blocks || !blocks_len || (!clip && n_clip) || (!fq && n_fq)) return SVB_ERR_ARG;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    struct Block {
        const char *chrom;
        size_t chrom_len;
        char side;
        uint64_t c0, c1, lines;
    };
    std::vector<Block> bl;
    const char *c = (const char *)clip, *ce = c + n_clip;
    for (const char *p = c; p < ce;) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(ce - p));
        if (!nl) return SVB_ERR_FORMAT;
        const char *t1 = (const char *)memchr(p, '\t', (size_t)(nl - p));
        const char *t2 = t1 ? (const char *)memchr(t1 + 1, '\t', (size_t)(nl - t1 - 1)) : nullptr;
        if (!t2 || t2 + 1 >= nl) return SVB_ERR_FORMAT;
        const size_t len = (size_t)(t1 - p);
        const char side = t2[1];
        if (bl.empty() || bl.back().side != side || bl.back().chrom_len != len || memcmp(bl.back().chrom, p, len) != 0)
            bl.push_back(Block{p, len, side, (uint64_t)(p - c), 0, 0});
        bl.back().c1 = (uint64_t)(nl + 1 - c), bl.back().lines += 1;
        p = nl + 1;
    }
    // four FASTQ lines per clip line
    std::vector<uint64_t> f_end(bl.size(), 0);
    {
        const char *f = (const char *)fq, *fe = f + n_fq, *q = f;
        for (size_t b = 0; b < bl.size(); ++b) {
            for (uint64_t k = 0; k < 4 * bl[b].lines; ++k) {
                const char *nl = q < fe ? (const char *)memchr(q, '\n', (size_t)(fe - q)) : nullptr;
                if (!nl) return SVB_ERR_FORMAT;
                q = nl + 1;
            }
            f_end[b] = (uint64_t)(q - f);
        }
        if (q != fe) return SVB_ERR_FORMAT;
    }
    std::vector<GzJob> jobs;
    std::string list;
    for (size_t b = 0; b < bl.size(); ++b) {
        const std::string base = std::string(part_prefix) + "." + std::to_string(b);
        jobs.push_back(GzJob{base + ".clip.gz", c + bl[b].c0, bl[b].c1 - bl[b].c0});
        const uint64_t f0 = b ? f_end[b - 1] : 0;
        jobs.push_back(GzJob{base + ".fq.gz", (const char *)fq + f0, f_end[b] - f0});
        list.append(bl[b].chrom, bl[b].chrom_len);
        list += '\t', list += bl[b].side, list += '\n';
    }
    std::string err;
    if (!jobs.empty() && !write_gz_many(jobs, n_threads, err)) return SVB_ERR_IO;
    *blocks = (char *)malloc(list.size() + 1);
    if (!*blocks) return SVB_ERR_IO;
    memcpy(*blocks, list.data(), list.size());
    (*blocks)[list.size()] = 0;
    *blocks_len = list.size();
    return 0;
}

extern "C" int svb_write_gz(const char *path, const void *data, uint64_t n, int n_threads)
{
    if (!path || (!data && n)) return SVB_ERR_ARG;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::string err;
    return write_gz(path, (const char *)data, n, n_threads, err) ? 0 : SVB_ERR_IO;
}

extern "C" int64_t svb_bai_first_offsets(const char *bai_path, uint64_t *first_voff, int64_t cap)
{
    if (!bai_path) return SVB_ERR_ARG;
    std::vector<uint64_t> v;
    std::string err;
    if (!bai_first_offsets(bai_path, v, err)) return SVB_ERR_IO;
    for (size_t i = 0; i < v.size() && (int64_t)i < cap && first_voff; ++i) first_voff[i] = v[i];
    return (int64_t)v.size();
}

extern "C" int64_t svb_bai_linear_offsets(const char *bai_path, int32_t tid, uint64_t *voff, int64_t cap)
{
    if (!bai_path) return SVB_ERR_ARG;
    std::vector<std::vector<uint64_t>> lin;
    std::string err;
    if (!bai_linear_offsets(bai_path, lin, err)) return SVB_ERR_IO;
    if (tid < 0 || (size_t)tid >= lin.size()) return SVB_ERR_ARG;
    for (size_t i = 0; i < lin[tid].size() && (int64_t)i < cap && voff; ++i) voff[i] = lin[tid][i];
    return (int64_t)lin[tid].size();
}

extern "C" int svb_bam_peek_record(const char *bam_path, uint64_t voffset, int32_t *tid, int32_t *pos)
{
    if (!bam_path || !tid || !pos) return SVB_ERR_ARG;
    MappedFile mf;
    std::string err;
    if (!mf.open(bam_path, err)) return SVB_ERR_IO;
    uint8_t head[12];
    if (!bgzf_read_at(mf.data, mf.size, voffset, head, 12, err)) return SVB_ERR_FORMAT;
    memcpy(tid, head + 4, 4);
    memcpy(pos, head + 8, 4);
    return 0;
}

extern "C" int svb_voffset_distance(const char *bam_path, uint64_t v_a, uint64_t v_b, uint64_t *bytes)
{
    if (!bam_path || !bytes) return SVB_ERR_ARG;
    MappedFile mf;
    std::string err;
    if (!mf.open(bam_path, err)) return SVB_ERR_IO;
    return bgzf_voffset_distance(mf.data, mf.size, v_a, v_b, *bytes, err) ? 0 : SVB_ERR_FORMAT;
}

extern "C" int svb_read_gz(const char *path, char **data, uint64_t *n)
{
    if (!path || !data || !n) return SVB_ERR_ARG;
    std::string text, err;
    if (!read_text_maybe_gz(path, text, err)) return SVB_ERR_IO;
    *data = (char *)malloc(text.size() + 1);
    if (!*data) return SVB_ERR_IO;
    memcpy(*data, text.data(), text.size());
    (*data)[text.size()] = 0;
    *n = text.size();
    return 0;
}

extern "C" int svb_sam_to_stream(const char *sam_path, char **stream, uint64_t *nbytes, uint64_t *first_record)
{
    if (!sam_path || !stream || !nbytes || !first_record) return SVB_ERR_ARG;
    std::vector<uint8_t> file, out;
    std::string err;
    if (!read_file(sam_path, file, err)) return SVB_ERR_IO;
    BamHeader h;
    if (!sam_to_bam_stream(file, h, out, err)) return SVB_ERR_FORMAT;
    *stream = (char *)malloc(out.size() + 1);
    if (!*stream) return SVB_ERR_IO;
    memcpy(*stream, out.data(), out.size());
    *nbytes = out.size(), *first_record = h.first_record;
    return 0;
}

extern "C" void svb_free(void *p) { free(p); }

extern "C" int svb_plan_getsv(const char *clip_aln, const char *clip_file, int32_t n_ref, const char *const *ref_names,
                              const uint32_t *ref_lens, int32_t reach, int32_t flank_len, svb_junction **junctions, uint64_t *n_j,
                              svb_window **windows, uint64_t *n_w)
{
    if (!clip_aln || !clip_file || !junctions || !n_j || !windows || !n_w || (n_ref && (!ref_names || !ref_lens))) return SVB_ERR_ARG;
    std::string err, clip_text;
    AlignmentSet alns;
    Phase ph;
    if (!load_alignments(clip_aln, alns, err)) return SVB_ERR_IO;
    ph.mark("plan: read clip alignments");
    if (!read_text_maybe_gz(clip_file, clip_text, err)) return SVB_ERR_IO;
    ph.mark("plan: read clip.gz");
    JunctionMap jm;
    std::vector<ClipLine> lines = parse_clip_text(clip_text, n_threads());
    ph.mark("plan: tokenise");
    if (!join_clips(lines, alns, jm, err, false)) return SVB_ERR_CUDA;
    ph.mark("plan: join");
    merge_junctions(jm, reach);
    ph.mark("plan: merge");
    std::map<std::string, int32_t> tid_of;
    std::vector<uint32_t> lens(ref_lens, ref_lens + n_ref);
    for (int32_t t = 0; t < n_ref; ++t) tid_of.insert(std::make_pair(std::string(ref_names[t]), t));
    *n_j = jm.size();
    *junctions = (svb_junction *)malloc(std::max<size_t>(1, jm.size()) * sizeof(svb_junction));
    size_t i = 0;
    for (auto &kv : jm) {
        svb_junction &j = (*junctions)[i++];
        auto a = tid_of.find(kv.first.up_chr), b = tid_of.find(kv.first.down_chr);
        j.up_tid = a == tid_of.end() ? -1 : a->second, j.down_tid = b == tid_of.end() ? -1 : b->second;
        j.up_pos = kv.first.up_pos, j.down_pos = kv.first.down_pos;
        j.up_strand = kv.first.up_strand, j.down_strand = kv.first.down_strand, j.pad_[0] = j.pad_[1] = 0;
    }
    PosDepth pos2depth;
    RangeDepth range2depth;
    WindowMap begin2end;
    JunctionRanges j2r;
    collect_breaks(jm, flank_len, pos2depth, range2depth, j2r);
    merge_ranges(range2depth, begin2end);
    std::vector<svb_window> dw;
    std::vector<Win> hw;
    device_windows(begin2end, tid_of, lens, dw, hw);
    *n_w = dw.size();
    *windows = (svb_window *)malloc(std::max<size_t>(1, dw.size()) * sizeof(svb_window));
    if (!dw.empty()) memcpy(*windows, dw.data(), dw.size() * sizeof(svb_window));
    return 0;
}

// The junctions `somatic` asks the normal BAM about (the rows of ReadTumorFileAndOutputSomaticInfo that run
// FindDiscordantReadPairs, somatic.cpp:111-409), in the order in which cmd_somatic batches them: for callers that run the
// pair test themselves (sharded runs). mean_insert is the normal BAM's mean insert size (0 when -n < 100000).
extern "C" int svb_plan_somatic(const char *normal_clip_path, const char *tumor_sv_path, double match_rate, int32_t offset, int32_t min_len,
                                int32_t mean_insert, int32_t n_ref, const char *const *ref_names, svb_junction **junctions, uint64_t *n_junctions)
{
    if (!normal_clip_path || !tumor_sv_path || !junctions || !n_junctions || (n_ref && !ref_names)) return SVB_ERR_ARG;
    std::string clip_text, tumor_text, err, log;
    if (!read_text_maybe_gz(normal_clip_path, clip_text, err) || !read_text_maybe_gz(tumor_sv_path, tumor_text, err)) return SVB_ERR_IO;
    std::vector<SomaticRow> rows;
    somatic_rows(clip_text, tumor_text, match_rate, offset, min_len, mean_insert, rows, log);
    auto tid_of = [&](const std::string &name) {
        for (int32_t t = 0; t < n_ref; ++t)
            if (name == ref_names[t]) return t;
        return (int32_t)-1;
    };
    std::vector<svb_junction> dj;
    for (const SomaticRow &r : rows)
        if (!r.is_header && r.query_pairs) {
            svb_junction j;
            j.up_tid = tid_of(r.key.up_chr), j.down_tid = tid_of(r.key.down_chr);
            j.up_pos = r.key.up_pos, j.down_pos = r.key.down_pos, j.up_strand = r.key.up_strand, j.down_strand = r.key.down_strand;
            j.pad_[0] = j.pad_[1] = 0;
            dj.push_back(j);
        }
    *n_junctions = dj.size();
    *junctions = (svb_junction *)malloc(std::max<size_t>(1, dj.size()) * sizeof(svb_junction));
    if (!dj.empty()) memcpy(*junctions, dj.data(), dj.size() * sizeof(svb_junction));
    return 0;
}

// `seeksv run -- <command> [-- <command> ...]`: getclip / getsv / somatic segments run in this process (BAMs stay resident
// between them), every other segment is executed as an external command (one word: through `sh -c`) and waited for.
static int cmd_run(int argc, char **argv)
{
    std::vector<std::vector<std::string>> segs;
    for (int i = 1; i < argc; ++i) {
        if (strcmp(argv[i], "--") == 0) {
            segs.emplace_back();
            continue;
        }
        if (segs.empty()) segs.emplace_back();
        segs.back().push_back(argv[i]);
    }
    segs.erase(std::remove_if(segs.begin(), segs.end(), [](const std::vector<std::string> &v) { return v.empty(); }), segs.end());
    if (segs.empty() || segs[0][0] == "--help" || segs[0][0] == "-h") {
        usage_run();
        return 1;
    }
    g_keep_resident = true;
    int rc = 0;
    for (auto &seg : segs) {
        const std::string &c = seg[0];
        if (c == "getclip" || c == "getsv" || c == "somatic") {
            std::vector<char *> av;
            std::string prog = "seeksv";
            av.push_back(&prog[0]);
            for (auto &a : seg) av.push_back(&a[0]);
            rc = svb_main((int)av.size(), av.data());
        } else {
            std::cout << std::flush;
            std::cerr << std::flush;
            std::vector<std::string> cmd = seg;
            if (cmd.size() == 1) cmd = {"sh", "-c", seg[0]};
            std::vector<char *> av;
            for (auto &a : cmd) av.push_back(&a[0]);
            av.push_back(nullptr);
            pid_t pid = 0;
            if (posix_spawnp(&pid, av[0], nullptr, nullptr, av.data(), environ) != 0) {
                std::cerr << "[seeksv run] cannot execute " << cmd[0] << std::endl;
                rc = 1;
            } else {
                int st = 0;
                while (waitpid(pid, &st, 0) < 0 && errno == EINTR) {}
                rc = WIFEXITED(st) ? WEXITSTATUS(st) : 1;
                if (rc != 0) std::cerr << "[seeksv run] '" << c << "' ended with status " << rc << std::endl;
            }
        }
        if (rc != 0) break;
    }
    drop_resident();
    g_keep_resident = false;
    return rc;
}

extern "C" int svb_main(int argc, char **argv)
{
    // main + SelectStep, seeksv.cpp:26-58,444-457
    if (argc == 1) {
        usage_top();
        return 1;
    }
    g_shard = ShardResults();
    if (!g_shard.load()) {
        std::cerr << "[seeksv_b200] cannot read SEEKSV_B200_SHARD_RESULTS" << std::endl;
        return 1;
    }
    if (g_provider.fn && strcmp(argv[1], "getsv") == 0) g_shard.on = true;  // results arrive through the provider
    if (strcmp(argv[1], "run") == 0) return cmd_run(argc - 1, argv + 1);
    const char *cmds[4] = {"getclip", "getsv", "somatic", "cluster"};
    int i = 0;
    for (; i < 4; ++i)
        if (strcmp(cmds[i], argv[1]) == 0) break;
    if (i == 4) {
        std::cerr << "[seeksv] unrecognized command '" << argv[1] << "'" << std::endl;
        return 1;
    }
    if (argc == 2) {
        usage_cmd("seeksv", argv[1], i);
        return 1;
    }
    switch (i) {
    case 0: return cmd_getclip(argc - 1, argv + 1);
    case 1: return cmd_getsv(argc - 1, argv + 1);
    case 2: return cmd_somatic(argc - 1, argv + 1);
    default: return 0;  // "cluster" is recognised and does nothing (its dispatch is commented out, seeksv.cpp:454)
    }
}
