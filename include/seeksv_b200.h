/* seeksv_b200 - C ABI of the B200-native seeksv hot path.
 *
 * The reference (qiukunlong/seeksv v1.2.3) has no plugin / FFI interface: its only stable boundary is
 * the process CLI plus file formats, and - inside the process - the seams between its I/O loops and
 * the per-record work. Each entry point below replaces one of those seams;
 * the citation names the reference function whose work it does (paths relative to
 * /root/reference/seeksv/). The CLI surface itself (seeksv.cpp:26-457) is kept by the C++ host layer
 * in seeksv_b200/host/, which is a thin caller of this ABI.
 *
 * Conventions: plain C types only; every function returns 0 on success and a negative svb_status on
 * failure (svb_last_error() gives the message); no exceptions cross the boundary; one svb_ctx per
 * GPU, used from one host thread at a time. There is NO CPU fallback: without a CUDA device
 * svb_ctx_create fails with SVB_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef SEEKSV_B200_H
#define SEEKSV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVB_ABI_VERSION 2

typedef enum svb_status {
    SVB_OK = 0,
    SVB_ERR_NO_DEVICE = -1, /* no CUDA device / wrong architecture */
    SVB_ERR_CUDA = -2,      /* a CUDA runtime call or kernel failed */
    SVB_ERR_FORMAT = -3,    /* input is not a BAM/BGZF/SAM stream this path understands */
    SVB_ERR_ARG = -4,       /* bad argument */
    SVB_ERR_IO = -5,        /* file could not be opened / read / written */
    SVB_ERR_UNSORTED = -6   /* getsv/somatic need a coordinate-sorted BAM (the reference needs its .bai) */
} svb_status;

typedef struct svb_ctx svb_ctx; /* one GPU: device, stream, scratch pools, timers   */
typedef struct svb_bam svb_bam; /* one BAM (or shard of one) resident in HBM        */

/* ---- context ------------------------------------------------------------------------------------ */
int svb_abi_version(void);
int svb_ctx_create(int device, svb_ctx **out);
void svb_ctx_destroy(svb_ctx *ctx);
const char *svb_last_error(const svb_ctx *ctx); /* ctx may be NULL: error of the last failed create */
/* The CUDA stream every launch of this ctx goes to (cudaStream_t), so that callers can bracket calls
 * with their own CUDA events. One stream per context IS the contract (SURVEY.md section 8b asks for "one
 * stream argument"; it is an argument of the context instead of every call): the entry points are
 * stream-ordered sequences that share the context's workspaces, pinned pools and control block, so two
 * calls on one context may not overlap anyway. A caller that wants concurrency gives each concurrent
 * activity its own context on the same device (bench.py's pairing thread, seeksv_b200/mgpu.py's loader
 * thread do); contexts are cheap after the first. */
void *svb_ctx_stream(svb_ctx *ctx);

/* Per-kernel device time accounting (CUDA events on the ctx stream). Enable, run, then read back:
 * names[i] / ms[i] / launches[i] for up to cap kernels; returns the number of distinct kernels. */
void svb_prof_enable(svb_ctx *ctx, int on);
void svb_prof_reset(svb_ctx *ctx);
int svb_prof_read(svb_ctx *ctx, int cap, const char **names, double *ms, int64_t *launches, double *bytes);

/* ---- BAM residency: replaces samopen/samread -> bam_read1 (clip_reads.h:375,410; cluster.cpp:27,48;
 *      getsv.cpp:1063-1067; bam2depth.cpp:57-75), i.e. the L0 libbam reader -------------------------- */

/* `stream` = the UNCOMPRESSED BAM byte stream ("BAM\1" header + packed records) or a contiguous shard
 * of its record section. first_record = byte offset of the first record start inside `stream`
 * (header length for a whole file). n_ref = number of reference sequences in the header.
 * _device: `stream` is a device pointer (>= nbytes + 64 readable bytes, and 16 readable bytes in front of it unless it
 *          is 16-byte aligned), borrowed for
 *          the life of the svb_bam. _host: copied host->device inside the call (pinned or pageable). */
int svb_bam_from_device(svb_ctx *ctx, const void *d_stream, uint64_t nbytes, uint64_t first_record, int32_t n_ref,
                        svb_bam **out);
int svb_bam_from_host(svb_ctx *ctx, const void *h_stream, uint64_t nbytes, uint64_t first_record, int32_t n_ref,
                      svb_bam **out);
/* Whole .bam file image (BGZF) in host memory. Default: the compressed image goes through pinned staging slabs
 * (cudaMemcpyAsync) and every BGZF block is inflated on the device (one warp per block). With
 * SEEKSV_B200_HOST_INFLATE=1 host threads inflate into the pinned slabs instead and the uncompressed bytes are
 * streamed. The two paths produce identical bytes. n_threads <= 0: use all hardware threads. */
int svb_bam_from_bgzf(svb_ctx *ctx, const void *h_file, uint64_t file_bytes, int n_threads, svb_bam **out);
int svb_bam_open(svb_ctx *ctx, const char *path, int n_threads, svb_bam **out); /* .bam, else SAM text */

/* Chromosome shard of an indexed, coordinate-sorted BAM: the records of references [tid_begin, tid_end), cut at exact
 * record boundaries with the .bai (bai_path NULL: bam_path + ".bai"; the index the reference requires for getsv and
 * somatic, seeksv.cpp:272-279, somatic.cpp:47-54, read with bam_index_load there). Only the BGZF blocks of the range are
 * read and inflated. The range that reaches the last reference with records also holds the unplaced reads at the end of
 * the file. The handle carries the full reference dictionary, so tids mean the same in every shard. */
int svb_bam_open_refs(svb_ctx *ctx, const char *bam_path, const char *bai_path, int32_t tid_begin, int32_t tid_end, int n_threads,
                      svb_bam **out);
/* Coordinate-range shard of an indexed BAM: the records between two BGZF virtual offsets that are record boundaries
 * (entries of the .bai's linear index, svb_bai_linear_offsets). v_begin = 0: from the first record; v_end = ~0: to the end
 * of the file. seeksv_b200/sharding.py:plan_range_shards picks the offsets, the halo and the key bounds. */
int svb_bam_open_voffsets(svb_ctx *ctx, const char *bam_path, uint64_t v_begin, uint64_t v_end, int n_threads, svb_bam **out);
/* Host only, for planning range shards: the linear index of one reference (virtual offset of the first record that overlaps
 * each 16 kb window, 0 = none; returns the number of windows), (tid, 0-based pos) of the record at a virtual offset, and
 * the number of uncompressed bytes between two virtual offsets. */
int64_t svb_bai_linear_offsets(const char *bai_path, int32_t tid, uint64_t *voff, int64_t cap);
int svb_bam_peek_record(const char *bam_path, uint64_t voffset, int32_t *tid, int32_t *pos);
int svb_voffset_distance(const char *bam_path, uint64_t v_a, uint64_t v_b, uint64_t *bytes);
/* tid of the last record of the stream that takes getclip's mapped branch (neither FUNMAP nor FMUNMAP set,
 * clip_reads.h:415-438): what the next shard passes as svb_getclip_params.prev_tid. *has_one = 0 if there is none. */
int svb_bam_last_mapped_tid(svb_ctx *ctx, svb_bam *bam, int32_t *has_one, int32_t *tid);
/* Host only: BGZF virtual offset (compressed offset << 16 | offset inside the block) of the first record of every reference
 * according to the .bai, ~0 for references without records. Returns the number of references in the index (writes at most
 * `cap` entries), negative on error. */
int64_t svb_bai_first_offsets(const char *bai_path, uint64_t *first_voff, int64_t cap);
void svb_bam_free(svb_bam *bam);

/* The resident uncompressed stream of a svb_bam (device pointer, byte count, offset of the first record), e.g. to
 * build further svb_bam views over it with svb_bam_from_device. Valid until svb_bam_free(bam). */
int svb_bam_device_stream(const svb_bam *bam, const void **d_stream, uint64_t *nbytes, uint64_t *first_record);

/* Inflate any BGZF image on the device (inflate.cu alone) and copy the result to h_out; *out_len is always set to
 * the uncompressed size, so a first call with h_out = NULL sizes the buffer (diagnostics / tests). */
int svb_inflate_bgzf(svb_ctx *ctx, const void *h_file, uint64_t file_bytes, void *h_out, uint64_t out_cap, uint64_t *out_len);
/* Copy a range of the resident uncompressed stream back to the host (diagnostics / tests). */
int svb_bam_copy_stream(const svb_bam *bam, void *h_dst, uint64_t offset, uint64_t nbytes);

uint64_t svb_bam_n_records(const svb_bam *bam);
uint64_t svb_bam_record_bytes(const svb_bam *bam);  /* sum over records of 4 + block_size */
int32_t svb_bam_n_ref(const svb_bam *bam);
const char *svb_bam_ref_name(const svb_bam *bam, int32_t tid); /* NULL when the header was not parsed */
uint32_t svb_bam_ref_len(const svb_bam *bam, int32_t tid);
/* Attach header names/lengths to a svb_bam built from a raw stream (needed for text output). */
int svb_bam_set_refs(svb_bam *bam, int32_t n_ref, const char *const *names, const uint32_t *lengths);

/* ---- getclip: replaces GetSClipReads + GetSeq + GenerateCigar + InsertSeq +
 *      ReadsInfo::ChangeSeqAndQual (clip_reads.h:120, clip_reads.cpp:57-108,112-192,260-329) and the
 *      unmapped-mate branch of InputBamOutputReads (clip_reads.h:415-420,172-219) -------------------- */
typedef struct svb_getclip_params {
    double match_rate;        /* -t, default 0.9  (seeksv.cpp:131)  */
    int32_t min_mapq;         /* -q, default 1    (seeksv.cpp:130)  */
    int32_t save_low_quality; /* -s, default 0    (seeksv.cpp:133)  */
    /* Sharded runs: tid of the last mapped-branch record BEFORE this shard (quirk Q1); 0 for a whole
     * file (clip_reads.h:407 starts last_tid at 0). */
    int32_t prev_tid;
    /* Sharded runs: mates of the unmapped branch are paired by name across the WHOLE file (clip_reads.h:172-219), so a
     * shard cannot pair on its own. With this flag set the shard skips the pairing (texts 2 and 3 stay empty) and hands
     * back its unmapped-branch records instead (svb_clusters_unmapped_records); the merging rank concatenates the shards'
     * records in file order into one small stream and runs svb_getclip on it for the two FASTQ texts. */
    int32_t export_unmapped_records;
    /* Coordinate-range shards (inside a chromosome): the stream is [context + halo][own records]. Records that start
     * before halo_bytes (stream offset of the first own record) only lend their soft clips: they are not part of the
     * unmapped branch's output. With key_filter set, only breakpoint keys (tid, 1-based pos) in [lo, hi) are clustered
     * here - a key belongs to the shard whose own records start at or before it - so every key is clustered on exactly
     * one shard, with all of its reads, in file order (seeksv_b200/sharding.py plans halo and bounds from the .bai). */
    int32_t key_filter;
    int32_t key_lo_tid, key_lo_pos, key_hi_tid, key_hi_pos;
    uint64_t halo_bytes;
    /* 1: the four outputs are compressed on the device and handed back as gzip file images (svb_clusters_gz) instead of
     * text (svb_clusters_text then returns empty buffers): smaller device->host copy, no deflate work on the host. The
     * images are what svb_write_gz would write for the same text (multi-member, Huffman-only). */
    int32_t gz_outputs;
    /* 1: the walk over the records also leaves getsv's per-record rows in HBM (same pass, +32 bytes written per record), so
     * that svb_insert_stats / svb_discordant_support / svb_window_depth / svb_getsv_passes on the SAME handle do not stream
     * the records a second time (`seeksv run`: getclip and getsv of one BAM in one process). */
    int32_t with_rows;
    /* 1: only the unmapped branch is evaluated (texts 2 and 3); the stream is the concatenation of unmapped-branch records that
     * the shards of one BAM handed over (svb_clusters_export_device), in file order. */
    int32_t unmapped_only;
    /* with export_unmapped_records: the exported records are grouped by FNV-1a-64(read name) % export_partitions (file order
     * inside a group; 0 or 1: one group; at most 64), so that N ranks can exchange them all-to-all and each pair the names of
     * one group - mates share a name, so every pair is decided on exactly one rank. svb_clusters_export_parts gives the byte
     * offsets of the groups. */
    int32_t export_partitions;
} svb_getclip_params;

typedef struct svb_clusters svb_clusters; /* host-resident result of svb_getclip */

int svb_getclip(svb_ctx *ctx, svb_bam *bam, const svb_getclip_params *p, svb_clusters **out);
void svb_clusters_free(svb_clusters *c);
uint64_t svb_clusters_count(const svb_clusters *c);
uint64_t svb_clusters_candidates(const svb_clusters *c); /* soft-clipped reads that entered clustering */
/* Decompressed contents of P.clip.gz / P.clip.fq.gz / P.unmapped_1.fq.gz / P.unmapped_2.fq.gz
 * (DisplaySClipReadsAndClipFq clip_reads.h:300-345, StoreUnmapSeqAndQual clip_reads.h:172-219), as
 * host buffers owned by the result object. which: 0 clip, 1 clip.fq, 2 unmapped_1, 3 unmapped_2.
 * svb_getclip leaves the texts in HBM; the first svb_clusters_text of a text copies it to (pinned) host memory.
 * svb_clusters_text_len gives a text's length without copying it. */
int svb_clusters_text(const svb_clusters *c, int which, const char **data, uint64_t *len);
int svb_clusters_text_len(const svb_clusters *c, int which, uint64_t *len);
/* gz_outputs: the gzip file image of output `which` (write it to P.clip.gz etc. as is). */
int svb_clusters_gz(const svb_clusters *c, int which, const char **data, uint64_t *len);
/* The packed BAM records (block_size + body, file order) of the unmapped branch; only with export_unmapped_records. */
int svb_clusters_unmapped_records(const svb_clusters *c, const char **data, uint64_t *len);
/* The same records where svb_getclip left them: in HBM (device pointer, valid until svb_clusters_free), for device-to-device
 * exchange between ranks; offsets[0 .. n_parts] = byte offsets of the export_partitions groups (offsets[n_parts] = len). */
int svb_clusters_export_device(const svb_clusters *c, const void **d_records, uint64_t *len);
int svb_clusters_export_parts(const svb_clusters *c, uint64_t *offsets, int32_t n_parts);

/* ---- getsv / somatic device passes ----------------------------------------------------------------- */

/* CalculateInsertsizeDeviation (cluster.h:25, cluster.cpp:15-83): over the first max_pairs records (file
 * order) with mapQ >= min_mapq, not hard-clipped, PAIRED & PROPER & !DUP, isize > 0.
 * out[0] = n, out[1] = sum isize, out[2] = mean (= sum / n, 0 if n == 0),
 * out[3] = sum over those records of (int32)((isize-mean)*(isize-mean)) (the reference's int product). */
int svb_insert_stats(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t max_pairs, int64_t out[4]);

/* FindDiscordantReadPairs (getsv.h:403-404, getsv.cpp:990-1247) for a batch of junctions. */
typedef struct svb_junction {
    int32_t up_tid, up_pos;     /* up_pos / down_pos are the 1-based junction coordinates */
    int32_t down_tid, down_pos; /* tid = -1: name not in the BAM header -> count 0 */
    char up_strand, down_strand; /* '+' / '-' */
    char pad_[2];
} svb_junction;
typedef struct svb_pair_params {
    int32_t min_mapq;    /* -q (20)                      */
    int32_t mean_insert; /* from svb_insert_stats         */
    int32_t deviation;
    int32_t times;       /* 4 (seeksv.cpp:161)            */
} svb_pair_params;
int svb_discordant_support(svb_ctx *ctx, svb_bam *bam, const svb_junction *junctions, uint64_t n,
                           const svb_pair_params *p, int32_t *counts /* host, n */);

/* main_depth (bam2depth.h:38, bam2depth.cpp:17-142): depth at every position of a set of disjoint
 * windows [begin, end] (1-based, inclusive, sorted by (tid, begin)); depth = reads accepted by libbam's
 * pileup (flag & 0x704 == 0 after mapQ < min_mapq -> UNMAP, 8000-read cap) with an M base at the
 * position. depth_out holds sum(end - begin + 1) ints, window after window. */
typedef struct svb_window {
    int32_t tid, begin, end;
} svb_window;
int svb_window_depth(svb_ctx *ctx, svb_bam *bam, const svb_window *windows, uint64_t n_windows, int32_t min_mapq,
                     int32_t *depth_out /* host */);

/* The three passes above as ONE stream-ordered sequence with one read-back (what getsv does to the original BAM after
 * the junction list is known, seeksv.cpp:272-300): insert-size statistics with -q / -n, the pair test of every junction with
 * the mean and deviation that never leave the device, depth of every window position. stats_out as svb_insert_stats
 * (deviation = (int)sqrt((double)out[3] / (int)out[0]) as the reference computes it). */
typedef struct svb_getsv_params {
    int32_t min_mapq;   /* -q (20): insert-size records, discordant pairs and pileup */
    int32_t times;      /* 4 (seeksv.cpp:161) */
    int64_t max_pairs;  /* -n (5000000) */
} svb_getsv_params;
int svb_getsv_passes(svb_ctx *ctx, svb_bam *bam, const svb_getsv_params *p, const svb_junction *junctions, uint64_t n_junctions,
                     const svb_window *windows, uint64_t n_windows, int64_t stats_out[4], int32_t *counts /* host */,
                     int32_t *depth_out /* host */);

/* Shards of one BAM on several GPUs: the insert-size statistics in additive pieces, so that the ranks can add
 * them up (NCCL all-reduce / all-gather of a few integers) and derive mean and deviation exactly as one process would.
 * svb_bam_set_own_offset: coordinate-range shards load [context][halo][own records); the getsv passes count own records only.
 * svb_insert_partial: out = {records taken, sum isize, sum isize^2, records with isize > 46340} over the first `take`
 *   qualifying own records in file order (take < 0: all). A non-zero out[3] anywhere means the reference's int products may
 *   wrap: then the exact sum of (int32)((isize - mean)^2) comes from svb_insert_sq with the global mean.
 * svb_pairs_depth: pair support and window depth of the own records with given statistics; counts / depth_out may be DEVICE
 *   pointers (one NCCL all-reduce adds the shards up). */
int svb_bam_set_own_offset(svb_bam *bam, uint64_t own_offset);
int svb_insert_partial(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t take, int64_t out[4]);
int svb_insert_sq(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t take, int32_t mean, int64_t *sq);
/* svb_insert_partial without the read-back: the four values land in d_out (DEVICE memory) in stream order. */
int svb_insert_partial_async(svb_ctx *ctx, svb_bam *bam, int32_t min_mapq, int64_t take, int64_t *d_out);
int svb_pairs_depth(svb_ctx *ctx, svb_bam *bam, const svb_pair_params *p, const svb_junction *junctions, uint64_t n_junctions,
                    const svb_window *windows, uint64_t n_windows, int32_t *counts, int32_t *depth_out);

/* ---- clip_join: P.clip.gz lines x realigned clip alignments -> junction candidates ----
 * Replaces the lock-step loop of InputSoftInfoStoreBreakpoint<T> (getsv.h:423-541) with GetAlignInfo (getsv.cpp:25-71) and the
 * key rules of GetJunction (getsv.cpp:1705-1845) on the device. The caller tokenises the two files (the host layer does:
 * host/junction.cpp parse_clip_text / parse_sam_alignments / parse_bam_alignments) and passes plain arrays:
 *   lines  - one per clip.gz line, file order: its clipped sequence (bytes seqs[seq_off, seq_off + seq_len)), breakpoint
 *            position, side character ('5' / '3') and the RANK of its chromosome name;
 *   alns   - one per alignment, file order: its read name (bytes names[name_off, ...) - the realigner names a read by its
 *            clipped sequence), FLAG, 0-based POS, MAPQ, CIGAR words cigars[cigar_off, cigar_off + n_cigar) and the rank of its
 *            chromosome name: the name of its tid, "" for a tid outside the header, "Exogenous" when FLAG & 4 (GetAlignInfo).
 * Ranks number the distinct names by std::string::compare order (the reference keys its maps on the names), below 2^30.
 * Result: every (head line of a run of equal clipped sequences, member of the run's alignment set) pair that GetJunction
 * stores, with the junction key it stores it under, stably sorted into Junction::operator< order (getsv.h:187-225) - candidates
 * of one key keep the order in which the reference's loop meets them, which is all its order-dependent accumulation (quirk Q8)
 * depends on. The accumulation itself and MergeJunction stay with the caller (host/junction.cpp). variant: 0 '+' side 5,
 * 1 '+' side 3, 2 / 3 '-' side 5 (alignment first / line first), 4 / 5 '-' side 3 (line first / alignment first);
 * uniq: 2 unique, 1 repeat (secondary or MAPQ 0). *cands is malloc'ed (svb_free). Sets of more than 4096 alignments for one
 * run are refused (SVB_ERR_FORMAT): the caller's host join handles them. */
typedef struct svb_join_line {
    uint32_t seq_off, seq_len;
    int32_t chr_rank, pos;
    uint32_t side;
} svb_join_line;
typedef struct svb_join_aln {
    uint32_t name_off, name_len, flag, cigar_off, n_cigar;
    int32_t chr_rank, pos, mapq;
} svb_join_aln;
typedef struct svb_join_cand {
    uint32_t line, aln; /* indices into lines (the run's head) and alns */
    int32_t up_rank, up_pos, down_rank, down_pos;
    uint8_t up_strand, down_strand, variant, uniq;
} svb_join_cand;
int svb_clip_join(svb_ctx *ctx, const svb_join_line *lines, uint64_t n_lines, const char *seqs, uint64_t seq_bytes,
                  const svb_join_aln *alns, uint64_t n_alns, const char *names, uint64_t name_bytes, const uint32_t *cigars,
                  uint64_t n_cigar_words, svb_join_cand **cands, uint64_t *n_cands);

/* Host-side planning step of getsv, exposed so that callers which keep the BAM resident (bench.py, sharded
 * runs) can drive the device passes themselves: joins P.clip.gz with the realigned clip.bam/clip.sam
 * (InputSoftInfoStoreBreakpoint getsv.h:423-541 + GetJunction getsv.cpp:1705), merges junctions (MergeJunction
 * getsv.cpp:1325, reach = -l), and returns the junction list in Junction::operator< order plus the merged
 * depth windows (GetBreak getsv.cpp:752 + MergeOverlap getsv.cpp:804, flank = -L) clamped to the
 * chromosomes and sorted by (tid, begin). Arrays are malloc'ed; release them with svb_free. No GPU work. */
int svb_plan_getsv(const char *clip_alignments_path, const char *clip_file_path, int32_t n_ref, const char *const *ref_names,
                   const uint32_t *ref_lens, int32_t merge_reach, int32_t flank_len, svb_junction **junctions,
                   uint64_t *n_junctions, svb_window **windows, uint64_t *n_windows);
/* The junctions `somatic` asks the NORMAL BAM about (somatic.cpp:111-409), in the order in which the command batches them, for
 * callers that run the pair test themselves (sharded runs). No GPU work; release with svb_free. */
int svb_plan_somatic(const char *normal_clip_path, const char *tumor_sv_path, double match_rate, int32_t offset, int32_t min_len,
                     int32_t mean_insert, int32_t n_ref, const char *const *ref_names, svb_junction **junctions, uint64_t *n_junctions);
void svb_free(void *p);

/* ---- gzip text files (host only) ----
 * Replaces the reference's ogzstream / igzstream (gzstream.h:86-117, gzstream.C:53-114; call sites clip_reads.h:392-395,
 * getsv.h:433, somatic.h:45). svb_write_gz writes `n` bytes as a multi-member gzip file (1 MiB members compressed by a
 * thread pool; every member carries its size in an 'SV' extra sub-field, which ordinary gzip readers skip).
 * svb_read_gz returns the decompressed content of any gzip or plain text file in a malloc'ed buffer (svb_free);
 * files written by svb_write_gz are inflated member-parallel. n_threads <= 0: all host cores. */
int svb_write_gz(const char *path, const void *data, uint64_t n, int n_threads);
/* Multi-GPU getclip (python -m seeksv_b200.mgpu): one rank's clip / clip.fq texts as gzip files per (chromosome, side) block,
 * <part_prefix>.<b>.clip.gz and <part_prefix>.<b>.fq.gz, b = 0, 1, ... in text order. The whole-file outputs are concatenations
 * of the ranks' block files: per chromosome (file order) the '5' blocks of all ranks, then the '3' blocks - the per-chromosome
 * flush of DisplaySClipReadsAndClipFq (clip_reads.h:300-345,423-438). *blocks: "chromosome<TAB>side\n" per block (svb_free). */
int svb_write_range_blocks(const char *part_prefix, const void *clip, uint64_t n_clip, const void *fq, uint64_t n_fq, int n_threads,
                           char **blocks, uint64_t *blocks_len);
/* The same file image made on the device (gzip.cu): text in host memory -> malloc'ed gzip image (svb_free). */
int svb_gzip_text(svb_ctx *ctx, const void *text, uint64_t n, char **gz, uint64_t *gz_len);
int svb_read_gz(const char *path, char **data, uint64_t *n);
/* The same for a file whose members hold at most 64 KiB of text each (what svb_clusters_gz / the getclip command write): the
 * members are inflated on the GPU by the BGZF inflate kernel - `getsv` reads P.clip.gz this way while the host cores stage the
 * BAM. *data points into pinned memory owned by the context and stays valid until the next call on this context (or
 * svb_ctx_destroy); it is NUL-terminated behind *n bytes. SVB_ERR_FORMAT: not such a file - use svb_read_gz. The CRC32 fields of
 * the members are not checked on this path (the deflate streams and the ISIZE fields are). */
int svb_read_gz_device(svb_ctx *ctx, const char *path, const char **data, uint64_t *n);
/* SAM text -> the uncompressed BAM byte stream svb_bam_from_host takes ("BAM\1" header + packed records; host only, malloc'ed,
 * svb_free). This is the conversion svb_bam_open / getsv apply to an input whose name does not end in ".bam"; it replaces
 * samopen(fn, "r") + sam_read1 of the linked libbam (bam_import.o; call sites clip_reads.h:375, getsv.h:445). */
int svb_sam_to_stream(const char *sam_path, char **stream, uint64_t *nbytes, uint64_t *first_record);

/* Sharded getsv (python -m seeksv_b200.mgpu getsv): the caller runs the three BAM passes of getsv on the shards of the BAM and
 * combines them (NCCL); svb_main("getsv") keeps the host bookkeeping. With a provider registered, getsv opens the BAM for its
 * header only and calls the provider once, after MergeJunction (getsv.cpp:1325-1482), with the junctions (for
 * FindDiscordantReadPairs, getsv.cpp:990-1247; none when -n < 100000) and the merged depth windows (main_depth,
 * bam2depth.cpp:17-142; none with -D) in the order in which it asks for their results. The provider fills stats = {records
 * behind the insert-size statistics, mean, deviation} (CalculateInsertsizeDeviation, cluster.cpp:15-83), one count per junction
 * and one depth per window position, and returns 0. NULL unregisters. */
typedef int (*svb_shard_provider_fn)(const svb_junction *junctions, uint64_t n_junctions, const svb_window *windows, uint64_t n_windows,
                                     int32_t min_mapq, int32_t pairs_used, int32_t times, int64_t *stats, int32_t *counts, int32_t *depth,
                                     void *user);
void svb_set_shard_provider(svb_shard_provider_fn fn, void *user);

/* ---- whole commands (what the CLI calls; same arguments as the reference's Call* functions,
 *      seeksv.cpp:128-410). They print the reference's progress lines to stderr and return the
 *      process exit code. ------------------------------------------------------------------------------ */
int svb_main(int argc, char **argv); /* argv as the seeksv binary gets it */

#ifdef __cplusplus
}
#endif
#endif /* SEEKSV_B200_H */
