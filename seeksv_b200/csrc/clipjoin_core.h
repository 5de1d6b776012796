// clip_join - the per-element rules of the device join of P.clip.gz lines with the realigned clip alignments
// (SURVEY.md section 8(b) item 2b). Replaces the lock-step loop of InputSoftInfoStoreBreakpoint<T> (getsv.h:423-541) with its
// callees GetAlignInfo (getsv.cpp:25-71) and the key rules of GetJunction (getsv.cpp:1705-1845).
//
// Everything in this header is plain C++ on plain arrays and compiles for the host as well as for the device: the kernels of
// clipjoin.cu are thin index loops around these functions, and tools/clipjoin_sim.cpp runs the very same functions in serial
// loops on the CPU, so the rules are checked against the host mirror (host/junction.cpp: join_clips_with_alignments) in the
// build container, where there is no GPU; on the GPU the tests check the kernels' plumbing (scans, sorts, launches) on top.
//
// The reference's loop is a sequential automaton over two streams:
//   lines   - clip.gz lines; RUNS of adjacent lines with the same clipped sequence; only the FIRST line of a run (its head) is
//             ever crossed with alignments (quirk Q6);
//   alns    - alignments in file order; at the first line of run k (k >= 1) the loop consumes alignments: hard-clipped ones are
//             skipped, ones named like run k-1's sequence join run k-1's set, and the first other one - the BREAKER b_k - ends
//             run k-1 and becomes the first member of run k's set, filed under run k-1's sequence. When the alignments run
//             out, the remaining lines are dropped. After the last line a trailing loop takes alignments (hard-clipped ones
//             too) as long as they are named like the last run.
//   set     - a std::map keyed (sequence filed under, (chromosome NAME, position)): iteration in key order, the first
//             insertion of a key wins. Every member is then passed to GetJunction with the run's head line.
// Parallel form: the breaker of a boundary depends on where the previous boundary stopped, so boundaries are walked in chunks
// of CJ_CHUNK from a GUESSED entry (blocks of equally named alignments line up with runs in well-formed input); exit(c) ==
// entry(c + 1) for every chunk proves the walk, wrong guesses are repaired from the predecessor's exit (the idiom of the BAM
// record walker, walk.cu). With the breakers known every run's set is a range of the alignment stream and the rest is
// per-run / per-member work.
#pragma once
#include <stdint.h>

#include "../../include/seeksv_b200.h"

#if defined(__CUDACC__)
#define CJ_HD __host__ __device__ __forceinline__
#else
#define CJ_HD inline
#endif

static constexpr uint32_t CJ_CHUNK = 64;      // run boundaries walked by one thread
static constexpr uint32_t CJ_MAX_SET = 4096;  // members of one run's set that the per-run insertion sort accepts

struct CjView {
    const svb_join_line *lines;
    uint64_t n_lines;
    const char *seqs;
    const svb_join_aln *alns;
    uint64_t n_alns;
    const char *names;
    const uint32_t *cigars;
};

CJ_HD bool cj_bytes_equal(const char *a, uint32_t la, const char *b, uint32_t lb)
{
    if (la != lb) return false;
    for (uint32_t i = 0; i < la; ++i)
        if (a[i] != b[i]) return false;
    return true;
}
// std::string_view::compare: unsigned bytes, then length
CJ_HD int cj_bytes_compare(const char *a, uint32_t la, const char *b, uint32_t lb)
{
    const uint32_t n = la < lb ? la : lb;
    for (uint32_t i = 0; i < n; ++i) {
        const unsigned char x = (unsigned char)a[i], y = (unsigned char)b[i];
        if (x != y) return x < y ? -1 : 1;
    }
    return la < lb ? -1 : la > lb ? 1 : 0;
}

CJ_HD bool cj_line_starts_run(const CjView &v, uint64_t i)
{
    if (i == 0) return true;
    const svb_join_line &a = v.lines[i - 1], &b = v.lines[i];
    return !cj_bytes_equal(v.seqs + a.seq_off, a.seq_len, v.seqs + b.seq_off, b.seq_len);
}
// blocks of equally named alignments (only used for the guesses)
CJ_HD bool cj_aln_starts_block(const CjView &v, uint64_t j)
{
    if (j == 0) return true;
    const svb_join_aln &a = v.alns[j - 1], &b = v.alns[j];
    return !cj_bytes_equal(v.names + a.name_off, a.name_len, v.names + b.name_off, b.name_len);
}
// IsHardClip, clip_reads.cpp:247-257
CJ_HD bool cj_hard_clipped(const CjView &v, uint64_t j)
{
    const svb_join_aln &a = v.alns[j];
    if (!a.n_cigar) return false;
    const uint32_t *cg = v.cigars + a.cigar_off;
    return (cg[0] & 15) == 5 || (cg[a.n_cigar - 1] & 15) == 5;
}
CJ_HD bool cj_named_like_run(const CjView &v, uint64_t j, uint64_t head_line)
{
    const svb_join_aln &a = v.alns[j];
    const svb_join_line &l = v.lines[head_line];
    return cj_bytes_equal(v.names + a.name_off, a.name_len, v.seqs + l.seq_off, l.seq_len);
}

// Entry guess of the chunk that starts at boundary k0 (k0 >= 2): the first unconsumed alignment after the breaker of boundary
// k0 - 1. In well-formed input that breaker is the first alignment of the block named like run k0 - 1, and block j belongs to run
// j; a missing or extra read upstream shifts the block numbers, so the blocks around number k0 - 1 are searched for one that is
// named like run k0 - 1 and whose predecessor is named like run k0 - 2 (two deep, as the BAM walker's guesses). A wrong guess costs
// repair rounds, never correctness.
static constexpr int64_t CJ_GUESS_WINDOW = 512;
CJ_HD uint64_t cj_guess_entry(const CjView &v, const uint32_t *run_head, const uint32_t *block_start, uint64_t n_blocks, uint64_t k0)
{
    const uint64_t kb = k0 - 1;
    if (n_blocks == 0) return v.n_alns;
    const int64_t center = (int64_t)(kb < n_blocks ? kb : n_blocks - 1);
    for (int64_t d = 0; d <= CJ_GUESS_WINDOW; ++d)
        for (int sgn = 0; sgn < (d ? 2 : 1); ++sgn) {
            const int64_t j = sgn ? center - d : center + d;
            if (j < 0 || j >= (int64_t)n_blocks) continue;
            if (!cj_named_like_run(v, block_start[j], run_head[kb])) continue;
            if (j > 0 && kb > 0 && !cj_named_like_run(v, block_start[j - 1], run_head[kb - 1])) continue;
            return (uint64_t)block_start[j] + 1;
        }
    return kb < n_blocks ? (uint64_t)block_start[kb] + 1 : v.n_alns;
}

// One chunk of run boundaries: boundaries k in [k0, k1) (k >= 1), entry = first alignment not yet consumed. breaker[k] = index
// of b_k, or n_alns when the alignments ran out (then every later boundary gets n_alns too). Returns the exit (next
// unconsumed alignment).
CJ_HD uint64_t cj_walk_chunk(const CjView &v, const uint32_t *run_head, uint64_t k0, uint64_t k1, uint64_t entry, uint32_t *breaker)
{
    uint64_t p = entry;
    const uint64_t m = v.n_alns;
    for (uint64_t k = k0; k < k1; ++k) {
        const uint64_t prev_head = run_head[k - 1];
        while (p < m && (cj_hard_clipped(v, p) || cj_named_like_run(v, p, prev_head))) ++p;
        if (p >= m) {
            for (; k < k1; ++k) breaker[k] = (uint32_t)m;
            return m;
        }
        breaker[k] = (uint32_t)p;
        ++p;
    }
    return p;
}

// The alignment range behind run k's set: [lo, hi). Non-last runs take the alignments that are not hard-clipped, the last run
// (the trailing loop, getsv.h:520-538) takes all of them. *crossed = false: the alignments ran out before the run began.
CJ_HD void cj_run_range(const CjView &v, const uint32_t *run_head, const uint32_t *breaker, uint64_t n_runs, uint64_t k, uint64_t *lo, uint64_t *hi,
                        bool *has_breaker, bool *all_kinds, bool *crossed)
{
    const uint64_t m = v.n_alns;
    *has_breaker = false, *all_kinds = false, *crossed = true;
    if (k >= 1) {
        if (breaker[k] >= m) {
            *crossed = false, *lo = *hi = m;
            return;
        }
        *has_breaker = true;
        *lo = (uint64_t)breaker[k] + 1;
    } else
        *lo = 0;
    if (k + 1 < n_runs) {
        *hi = breaker[k + 1];  // (n_alns when the alignments run out inside this run)
    } else {
        uint64_t p = *lo;
        while (p < m && cj_named_like_run(v, p, run_head[k])) ++p;
        *hi = p, *all_kinds = true;
    }
}

// position GetAlignInfo reports (getsv.cpp:25-71): 1-based, -1 for an unmapped alignment (chromosome "Exogenous")
CJ_HD int32_t cj_info_pos(const svb_join_aln &a) { return (a.flag & 4u) ? -1 : a.pos + 1; }

CJ_HD uint32_t cj_count_members(const CjView &v, uint64_t lo, uint64_t hi, bool all_kinds)
{
    if (all_kinds) return (uint32_t)(hi - lo);
    uint32_t n = 0;
    for (uint64_t j = lo; j < hi; ++j) n += !cj_hard_clipped(v, j);
    return n;
}

// Members of run k in the iteration order of the reference's map, duplicates of a key dropped (the first insertion wins):
// the breaker is filed under the PREVIOUS run's sequence, so it goes in front when that sequence compares below this run's and
// behind otherwise; the others are ordered by (chromosome name, position), insertion order = file order. out[] has room for
// has_breaker + count; returns the number kept. Insertion sort: sets are a handful of alignments (CJ_MAX_SET at most).
CJ_HD uint32_t cj_fill_members(const CjView &v, const uint32_t *run_head, uint64_t k, uint64_t lo, uint64_t hi, bool has_breaker, bool all_kinds,
                               uint32_t breaker, uint32_t *out)
{
    uint32_t n = 0;
    uint32_t *o = out + (has_breaker ? 1 : 0);
    for (uint64_t j = lo; j < hi; ++j) {
        if (!all_kinds && cj_hard_clipped(v, j)) continue;
        const svb_join_aln &a = v.alns[j];
        const int32_t ar = a.chr_rank, ap = cj_info_pos(a);
        // place among the ones kept so far; an equal key was inserted earlier and wins
        uint32_t at = n;
        bool dup = false;
        while (at > 0) {
            const svb_join_aln &b = v.alns[o[at - 1]];
            const int32_t br = b.chr_rank, bp = cj_info_pos(b);
            if (br == ar && bp == ap) {
                dup = true;
                break;
            }
            if (br < ar || (br == ar && bp < ap)) break;
            --at;
        }
        if (dup) continue;
        for (uint32_t t = n; t > at; --t) o[t] = o[t - 1];
        o[at] = (uint32_t)j;
        ++n;
    }
    if (!has_breaker) return n;
    const svb_join_line &prev = v.lines[run_head[k - 1]], &cur = v.lines[run_head[k]];
    const bool front = cj_bytes_compare(v.seqs + prev.seq_off, prev.seq_len, v.seqs + cur.seq_off, cur.seq_len) < 0;
    if (front) out[0] = breaker;
    else {
        for (uint32_t t = 0; t < n; ++t) out[t] = out[t + 1];
        out[n] = breaker;
    }
    return n + 1;
}

// GetAlignInfo + the key rules of GetJunction for (head line, alignment). Returns false when nothing is stored: an unmapped
// alignment (type 'n', quirk Q7), a side other than '5' / '3'.
CJ_HD bool cj_classify(const CjView &v, uint32_t line, uint32_t aln, svb_join_cand *c)
{
    const svb_join_aln &a = v.alns[aln];
    const svb_join_line &l = v.lines[line];
    if (a.flag & 4u) return false;
    if (l.side != '5' && l.side != '3') return false;
    int32_t len = 0;  // GenerateCigar's reference length (clip_reads.cpp:309-329): M, D, =, N
    const uint32_t *cg = v.cigars + a.cigar_off;
    for (uint32_t i = 0; i < a.n_cigar; ++i) {
        const uint32_t op = cg[i] & 15;
        if (op == 0 || op == 2 || op == 7 || op == 3) len += (int32_t)(cg[i] >> 4);
    }
    const int32_t ar = a.chr_rank, ap = a.pos + 1, aend = ap + len - 1;
    const int32_t lr = l.chr_rank, lp = l.pos;
    c->line = line, c->aln = aln;
    c->uniq = ((a.flag & 256u) || a.mapq == 0) ? 1 : 2;
    auto key = [&](int32_t ur, int32_t up, char us, int32_t dr, int32_t dp, char ds, uint8_t variant) {
        c->up_rank = ur, c->up_pos = up, c->up_strand = (uint8_t)us, c->down_rank = dr, c->down_pos = dp, c->down_strand = (uint8_t)ds;
        c->variant = variant;
    };
    if (!(a.flag & 16u)) {
        if (l.side == '5') key(ar, aend, '+', lr, lp, '+', 0);
        else key(lr, lp, '+', ar, ap, '+', 1);
    } else if (l.side == '5') {
        if (ar < lr || (ar == lr && ap <= lp)) key(ar, ap, '-', lr, lp, '+', 2);
        else key(lr, lp, '-', ar, ap, '+', 3);
    } else {
        if (lr < ar || (lr == ar && lp <= aend)) key(lr, lp, '+', ar, aend, '-', 4);
        else key(ar, aend, '+', lr, lp, '-', 5);
    }
    return true;
}

// Junction::operator< (getsv.h:187-225) as two radix keys: chromosomes and strands (high), positions (low)
CJ_HD uint64_t cj_key_high(const svb_join_cand &c)
{
    return ((uint64_t)(uint32_t)c.up_rank << 34) | ((uint64_t)((uint32_t)c.down_rank & 0xffffffffu) << 2) | (uint64_t)((c.up_strand == '-') << 1) |
           (uint64_t)(c.down_strand == '-');
}
CJ_HD uint64_t cj_key_low(const svb_join_cand &c)
{
    return ((uint64_t)((uint32_t)c.up_pos ^ 0x80000000u) << 32) | (uint64_t)((uint32_t)c.down_pos ^ 0x80000000u);
}
