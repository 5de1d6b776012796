// Host-side junction bookkeeping of getsv / somatic: the order-dependent std::multimap logic of the
// reference (getsv.h:423-541, getsv.cpp:25-71,752-987,1325-1511,1705-1862; somatic.h:40-70, somatic.cpp:14-427)
// restated. It touches one entry per soft-clip cluster (about 1-2 % of the records, SURVEY.md section 0: < 1 %
// of the reference's run time) and must reproduce multimap iteration order exactly, so it stays on the host;
// everything that touches every BAM record runs on the GPU behind include/seeksv_b200.h.
#pragma once
#include <stdint.h>

#include <list>
#include <map>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include "../../include/seeksv_b200.h"

namespace svb {

typedef std::vector<std::pair<int, char>> CigarVec;

struct SeqInfo {  // getsv.h:48-70
    std::string seq;
    CigarVec cigar;
    int lclip = 0, rclip = 0, support = 0, uniq = 0;
};

struct JunctionKey {  // getsv.h:149-227
    std::string up_chr, down_chr;
    int up_pos = 0, down_pos = 0;
    char up_strand = '+', down_strand = '+';
    bool operator<(const JunctionKey &o) const;
};

struct JunctionInfo {  // OtherInfo, getsv.h:88-107
    SeqInfo up, down;
    int micro = -1, pairs = 0;
};

typedef std::multimap<JunctionKey, JunctionInfo> JunctionMap;

struct ClipLine {  // one line of P.clip.gz (clip_reads.h:308-332); views into the decompressed file text
    std::string_view chr, cigar, aligned_seq, aligned_qual, clipped_seq, clipped_qual;
    int pos = 0, support = 0;
    char side = '5';
};

struct Alignment {  // the fields of a clip.bam / clip.sam record that GetAlignInfo reads (getsv.cpp:25-71)
    std::string_view qname;  // view into the alignment file image / converted stream
    uint32_t flag = 0;
    int32_t tid = -1, pos = 0, mapq = 0;
    uint32_t cigar_begin = 0, cigar_n = 0;  // slice of the shared CIGAR word array
};

struct AlignmentSet {
    std::vector<Alignment> recs;
    std::vector<uint32_t> cigar_words;
    std::vector<std::string> ref_names;
    std::vector<uint8_t> storage;  // what the qname views point into
    std::list<std::string> rebuilt_names;  // ... except the names libbam would have cut (SAM names of 255+ characters)
};

struct ChrRange {  // getsv.h:231-258 - unsigned on purpose (quirk Q11)
    std::string chr;
    unsigned begin, end;
    bool operator<(const ChrRange &o) const;
};

struct FlankRanges {
    ChrRange r[4];
};

std::vector<ClipLine> parse_clip_text(const std::string &text, int n_threads = 0);
std::vector<ClipLine> parse_clip_text(const char *text, size_t size, int n_threads);  // (views point into `text`)
// packed BAM records (after the header) -> alignment list; qname views point into `set.storage`
bool parse_bam_alignments(AlignmentSet &set, uint64_t first_record);
// SAM text (in set.storage) -> alignment list, parsed by n_threads threads
bool parse_sam_alignments(AlignmentSet &set, int n_threads, std::string &err);
CigarVec cigar_from_text(const std::string &s);
std::string cigar_to_text(const CigarVec &v, int left_clip, int right_clip);
std::string reverse_complement(const std::string &s);
double match_rate_from_end(const std::string &a, const std::string &b);
double match_rate_from_begin(const std::string &a, const std::string &b);
std::string format_double(double x);

// getsv -B: the junctions of an earlier output file enter the map before the join (ReadBreakpoint, getsv.cpp:1291-1323)
void read_breakpoints(const std::string &sv_text, JunctionMap &jm);
// getsv -F: junctions from "connected read-through reads" (FindJunction, process_bwasw.cpp:5-227). `stream` holds packed BAM
// records from `first` on (a BAM's uncompressed stream, or SAM text converted by sam_to_bam_stream).
void find_junctions(const uint8_t *stream, uint64_t n, uint64_t first, const std::vector<std::string> &ref_names, int min_mapq,
                    JunctionMap &jm);
void join_clips_with_alignments(const std::vector<ClipLine> &lines, const AlignmentSet &alns, JunctionMap &jm);
void merge_junctions(JunctionMap &jm, int search_length);
// the device form of the join (svb_clip_join): inputs packed for it, and its candidates taken into the map
struct JoinArrays {
    std::vector<svb_join_line> lines;
    std::vector<svb_join_aln> alns;
    std::string seqs, names;
    std::vector<std::string> rank_names;  // rank -> chromosome name
};
bool pack_join_inputs(const std::vector<ClipLine> &lines, const AlignmentSet &alns, JoinArrays &out, int n_threads = 0);
bool accumulate_join_candidates(const std::vector<ClipLine> &lines, const AlignmentSet &alns, const JoinArrays &arrays, const svb_join_cand *cands,
                                uint64_t n, JunctionMap &jm, std::string &err);

typedef std::map<std::pair<std::string, int>, int> PosDepth;        // pos2depth
typedef std::map<ChrRange, unsigned long> RangeDepth;               // range2depth
typedef std::map<std::pair<std::string, int>, int> WindowMap;       // begin2end
typedef std::map<JunctionKey, FlankRanges> JunctionRanges;          // junction2range_pair

void collect_breaks(const JunctionMap &jm, int flank_len, PosDepth &pos2depth, RangeDepth &range2depth, JunctionRanges &j2r);
void merge_ranges(const RangeDepth &range2depth, WindowMap &begin2end);
// One covered position (chr, p 1-based) with its depth: the two map walks of bam2depth.cpp:82-124.
void account_position(const std::string &chr, int p, int depth, const WindowMap &begin2end, PosDepth &pos2depth,
                      RangeDepth &range2depth);

struct OutputFilters {
    int min_clip_sum = 3, min_pairs = 0, min_distance = 50, max_micro = 50, min_seq_len = 30, max_indel = 1;
    double frequency = 0.1;
};
extern const char *kSvHeader;
void write_breakpoints(const JunctionMap &jm, const PosDepth &pos2depth, const RangeDepth &range2depth, const JunctionRanges &j2r,
                       const OutputFilters &f, std::string &sv_body, std::string &filtered_stdout, std::string &log);

// somatic
struct SomaticRow {
    std::string prefix;  // the tumour line re-printed (23 columns)
    JunctionKey key;
    int normal_left = 0, normal_right = 0;
    bool query_pairs = false;  // whether FindDiscordantReadPairs runs for this row
    bool is_header = false;    // '@' line: prefix is the complete output line
};
// Host part of ReadTumorFileAndOutputSomaticInfo (somatic.cpp:14-427): everything except the discordant-pair
// counts, which the caller fills in from one batched svb_discordant_support call.
void somatic_rows(const std::string &normal_clip_text, const std::string &tumor_sv_text, double rate, int offset, int min_len,
                  int mean_insert, std::vector<SomaticRow> &rows, std::string &log);

}  // namespace svb
