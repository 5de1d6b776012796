// clipjoin_sim - the device join's rules (seeksv_b200/csrc/clipjoin_core.h) run in serial loops on the CPU and checked against the
// host mirror of the reference's loop (host/junction.cpp: join_clips_with_alignments; InputSoftInfoStoreBreakpoint, getsv.h:423-541).
//
//   clipjoin_sim <clip.sam | clip.bam> <P.clip.gz> [--wrong-guesses | --dump DIR]
//
// Test infrastructure for the build container (no GPU there): the kernels of csrc/clipjoin.cu are index loops around the same
// functions, with chained scans and radix sorts where this file uses std:: algorithms. Prints "OK <runs> <candidates> <entries>
// <repair rounds>" when the two junction maps are identical (keys, sequences, CIGARs, clip lengths, support, uniqueness, order),
// "DIFF ..." and exit code 1 otherwise. --wrong-guesses starts every chunk from a deliberately wrong entry (entry 0), so the
// verify / repair rounds are exercised on well-formed input too. --dump DIR writes the packed input arrays of svb_clip_join
// (lines.bin, seqs.bin, alns.bin, names.bin, cigars.bin) and the sorted candidates (cands.bin) as raw structs: the GPU tests feed
// the arrays to the C ABI and expect exactly these candidates back.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../seeksv_b200/csrc/clipjoin_core.h"
#include "../seeksv_b200/host/bamfile.h"
#include "../seeksv_b200/host/junction.h"

using namespace svb;

static bool load(const std::string &path, AlignmentSet &set, std::string &err)
{
    std::vector<uint8_t> file;
    if (!read_file(path, file, err)) return false;
    if (path.size() >= 4 && path.rfind(".bam") == path.size() - 4) {
        BamHeader h;
        if (!bgzf_inflate_all(file.data(), file.size(), set.storage, 4, err)) return false;
        if (!parse_bam_header(set.storage.data(), set.storage.size(), h, err)) return false;
        set.ref_names = h.names;
        return parse_bam_alignments(set, h.first_record);
    }
    set.storage.swap(file);
    return parse_sam_alignments(set, 4, err);
}

static bool same(const SeqInfo &a, const SeqInfo &b)
{
    return a.seq == b.seq && a.cigar == b.cigar && a.lclip == b.lclip && a.rclip == b.rclip && a.support == b.support && a.uniq == b.uniq;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: clipjoin_sim <clip.sam|clip.bam> <P.clip.gz> [--wrong-guesses]\n");
        return 2;
    }
    const bool wrong = argc > 3 && !strcmp(argv[3], "--wrong-guesses");
    const std::string dump = argc > 4 && !strcmp(argv[3], "--dump") ? argv[4] : "";
    std::string err, text;
    AlignmentSet set;
    if (!load(argv[1], set, err) || !read_text_maybe_gz(argv[2], text, err)) {
        fprintf(stderr, "%s\n", err.c_str());
        return 2;
    }
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const bool timing = getenv("CLIPJOIN_SIM_TIMING") != nullptr;
    double t0 = now();
    std::vector<ClipLine> lines = parse_clip_text(text, 4);
    double t1 = now();
    JunctionMap want, got;
    join_clips_with_alignments(lines, set, want);
    double t2 = now();

    JoinArrays J;
    const bool packed = pack_join_inputs(lines, set, J);
    double t3 = now();
    if (timing) fprintf(stderr, "tokenise %.1f ms, host join %.1f ms, pack %.1f ms\n", t1 - t0, t2 - t1, t3 - t2);
    if (!packed) {
        fprintf(stderr, "inputs too large for the device join\n");
        return 2;
    }
    const CjView v{J.lines.data(), J.lines.size(), J.seqs.data(), J.alns.data(), J.alns.size(), J.names.data(), set.cigar_words.data()};
    const uint64_t n_lines = v.n_lines, m = v.n_alns;
    std::vector<svb_join_cand> cands;
    uint64_t n_runs = 0, rounds = 0;
    if (n_lines) {
        // 1. run heads, block starts
        std::vector<uint32_t> run_head, block_start;
        for (uint64_t i = 0; i < n_lines; ++i)
            if (cj_line_starts_run(v, i)) run_head.push_back((uint32_t)i);
        for (uint64_t j = 0; j < m; ++j)
            if (cj_aln_starts_block(v, j)) block_start.push_back((uint32_t)j);
        n_runs = run_head.size();
        // 2. breakers: chunks from guessed entries, verify, repair
        std::vector<uint32_t> breaker(n_runs + 1, 0);
        const uint64_t n_chunks = n_runs < 2 ? 0 : (n_runs - 1 + CJ_CHUNK - 1) / CJ_CHUNK;
        std::vector<uint64_t> entry(n_chunks), exit_(n_chunks);
        auto bounds = [&](uint64_t c, uint64_t &k0, uint64_t &k1) { k0 = 1 + c * CJ_CHUNK, k1 = std::min<uint64_t>(n_runs, k0 + CJ_CHUNK); };
        for (uint64_t c = 0; c < n_chunks; ++c) {
            uint64_t k0, k1;
            bounds(c, k0, k1);
            uint64_t e = c == 0 ? 0 : cj_guess_entry(v, run_head.data(), block_start.data(), block_start.size(), k0);
            if (wrong && c > 0) e = 0;
            entry[c] = e;
            exit_[c] = cj_walk_chunk(v, run_head.data(), k0, k1, e, breaker.data());
        }
        for (;;) {
            bool bad = false;
            for (uint64_t c = 1; c < n_chunks; ++c) bad |= entry[c] != exit_[c - 1];
            if (!bad) break;
            if (++rounds > n_chunks + 1) {
                printf("DIFF the boundary walk did not settle\n");
                return 1;
            }
            const std::vector<uint64_t> prev = exit_;
            for (uint64_t c = 1; c < n_chunks; ++c) {
                if (entry[c] == prev[c - 1]) continue;
                uint64_t k0, k1;
                bounds(c, k0, k1);
                entry[c] = prev[c - 1];
                exit_[c] = cj_walk_chunk(v, run_head.data(), k0, k1, entry[c], breaker.data());
            }
        }
        // 3. members per run, classified; 4. compaction in crossing order
        std::vector<uint32_t> members(m + 1);
        for (uint64_t k = 0; k < n_runs; ++k) {
            uint64_t lo, hi;
            bool hb, all, crossed;
            cj_run_range(v, run_head.data(), breaker.data(), n_runs, k, &lo, &hi, &hb, &all, &crossed);
            if (!crossed) continue;
            const uint32_t cnt = (uint32_t)hb + cj_count_members(v, lo, hi, all);
            if (cnt > CJ_MAX_SET) {
                printf("SKIP a run with %u alignments\n", cnt);
                return 0;
            }
            const uint32_t kept = cj_fill_members(v, run_head.data(), k, lo, hi, hb, all, hb ? breaker[k] : 0u, members.data());
            for (uint32_t t = 0; t < kept; ++t) {
                svb_join_cand c;
                if (cj_classify(v, run_head[k], members[t], &c)) cands.push_back(c);
            }
        }
        // two stable sorts: positions, then chromosome ranks + strands
        std::stable_sort(cands.begin(), cands.end(), [](const svb_join_cand &a, const svb_join_cand &b) { return cj_key_low(a) < cj_key_low(b); });
        std::stable_sort(cands.begin(), cands.end(), [](const svb_join_cand &a, const svb_join_cand &b) { return cj_key_high(a) < cj_key_high(b); });
    }
    if (!dump.empty()) {
        auto put = [&](const char *name, const void *p, size_t n) {
            FILE *f = fopen((dump + "/" + name).c_str(), "wb");
            if (!f || (n && fwrite(p, 1, n, f) != n)) {
                fprintf(stderr, "cannot write %s/%s\n", dump.c_str(), name);
                exit(2);
            }
            fclose(f);
        };
        put("lines.bin", J.lines.data(), J.lines.size() * sizeof(svb_join_line));
        put("seqs.bin", J.seqs.data(), J.seqs.size());
        put("alns.bin", J.alns.data(), J.alns.size() * sizeof(svb_join_aln));
        put("names.bin", J.names.data(), J.names.size());
        put("cigars.bin", set.cigar_words.data(), set.cigar_words.size() * 4);
        put("cands.bin", cands.data(), cands.size() * sizeof(svb_join_cand));
    }
    double t4 = now();
    const bool accumulated = accumulate_join_candidates(lines, set, J, cands.data(), cands.size(), got, err);
    if (timing) fprintf(stderr, "rules (serial) %.1f ms, accumulate %.1f ms\n", t4 - t3, now() - t4);
    if (!accumulated) {
        printf("DIFF %s\n", err.c_str());
        return 1;
    }
    if (want.size() != got.size()) {
        printf("DIFF %zu entries, the host mirror has %zu\n", got.size(), want.size());
        return 1;
    }
    auto a = want.begin(), b = got.begin();
    for (size_t i = 0; a != want.end(); ++a, ++b, ++i) {
        const bool key_same = !(a->first < b->first) && !(b->first < a->first);
        if (!key_same || !same(a->second.up, b->second.up) || !same(a->second.down, b->second.down) || a->second.micro != b->second.micro) {
            printf("DIFF entry %zu (%s:%d %s:%d)\n", i, a->first.up_chr.c_str(), a->first.up_pos, a->first.down_chr.c_str(), a->first.down_pos);
            return 1;
        }
    }
    printf("OK %llu %zu %zu %llu\n", (unsigned long long)n_runs, cands.size(), got.size(), (unsigned long long)rounds);
    return 0;
}
