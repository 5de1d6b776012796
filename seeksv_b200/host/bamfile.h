// Host-side container handling: BGZF block scan + threaded inflate, BAM header parse, SAM text -> packed
// BAM records. Replaces the parts of the reference's prebuilt libbam (sam/libbam.a: bgzf.o, bam.o,
// bam_import.o, sam.o; headers sam/bgzf.h, sam/bam.h, sam/sam.h) that the hot path sits on.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

struct BgzfBlock {  // (layout shared with the device inflate kernel: 24 bytes)
    uint64_t coff;  // offset of the deflate payload in the file image
    uint64_t uoff;  // offset in the uncompressed stream
    uint32_t clen;  // deflate payload length
    uint32_t ulen;  // uncompressed length (ISIZE)
};

struct BamHeader {
    std::string text;
    std::vector<std::string> names;
    std::vector<uint32_t> lengths;
    uint64_t first_record = 0;  // byte offset of the first record in the uncompressed stream
};

struct MappedFile {  // read-only mmap of a whole file (the descriptor stays open for pread)
    const uint8_t *data = nullptr;
    uint64_t size = 0;
    int fd = -1;
    bool open(const std::string &path, std::string &err);
    ~MappedFile();
};
bool read_file(const std::string &path, std::vector<uint8_t> &out, std::string &err);
uint32_t bgzf_block_size(const uint8_t *file, uint64_t n, uint64_t offset);  // whole member at `offset`, 0 if it is not one
bool bgzf_scan(const uint8_t *file, uint64_t n, std::vector<BgzfBlock> &blocks, uint64_t &total, std::string &err);
// inflate blocks [b0, b1) into dst (dst[0] corresponds to blocks[b0].uoff) with n_threads host threads
bool bgzf_inflate_range(const uint8_t *file, const std::vector<BgzfBlock> &blocks, size_t b0, size_t b1, uint8_t *dst,
                        int n_threads, std::string &err);
bool bgzf_inflate_all(const uint8_t *file, uint64_t n, std::vector<uint8_t> &out, int n_threads, std::string &err);
bool parse_bam_header(const uint8_t *data, uint64_t n, BamHeader &h, std::string &err);
// BAM header of a BGZF file image: inflates leading blocks on the host until the header parses
bool read_bam_header(const uint8_t *file, uint64_t n, BamHeader &h, std::string &err);
// .bai (sam/bam.h:498-536 bam_index_*; written by bam_index_build): BGZF virtual offset (coffset << 16 | uoffset) of the
// first record of every reference, ~0 for a reference without records
bool bai_first_offsets(const std::string &bai_path, std::vector<uint64_t> &first_voff, std::string &err);
// the linear index of every reference: virtual offset of the first record that overlaps each 16 kb window (0: none recorded)
bool bai_linear_offsets(const std::string &bai_path, std::vector<std::vector<uint64_t>> &linear, std::string &err);
// uncompressed bytes between two virtual offsets of a BGZF file image (walks the block headers from a to b; a <= b)
bool bgzf_voffset_distance(const uint8_t *file, uint64_t n, uint64_t v_a, uint64_t v_b, uint64_t &bytes, std::string &err);
// `want` uncompressed bytes starting at a virtual offset (host zlib; for peeking at single records)
bool bgzf_read_at(const uint8_t *file, uint64_t n, uint64_t voff, uint8_t *dst, uint32_t want, std::string &err);
// FLAG column of a SAM line the way the linked libbam reads it (decimal / hex / octal number, or flag letters "pPuUrR12sfd")
uint32_t sam_flag(const char *begin, const char *end);
// SAM text (what samopen(fn, "r") reads) -> "BAM\1" header + packed records, byte for byte what libbam's sam_read1 builds
// (tests/test_sam_text.py): CIGAR "*" sets the unmapped flag, integers narrow to the smallest aux type, names of 255+
// characters are cut to (length + 1) & 0xff bytes
bool sam_to_bam_stream(const std::vector<uint8_t> &text, BamHeader &h, std::vector<uint8_t> &stream, std::string &err);
// gzip/plain text file -> bytes (igzstream / ifstream of the reference, gzstream.h)
bool read_text_maybe_gz(const std::string &path, std::string &out, std::string &err);
// write `data` as a gzip file (ogzstream of the reference; multi-member, compressed by n_threads threads)
bool write_gz(const std::string &path, const char *data, uint64_t n, int n_threads, std::string &err);
struct GzJob {
    std::string path;
    const char *data;
    uint64_t n;
};
bool write_gz_many(const std::vector<GzJob> &jobs, int n_threads, std::string &err);
bool gz_on_host();  // SEEKSV_B200_GZ_LEVEL / SEEKSV_B200_GZ=host: the outputs are compressed by host threads, not on the device
// ready-made file images (compressed on the device): one writer thread per file
bool write_files(const std::vector<GzJob> &jobs, std::string &err);
