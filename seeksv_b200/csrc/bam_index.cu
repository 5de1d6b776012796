// Record-boundary resolution for a packed BAM record stream resident in HBM.
//
// Replaces the serial `while (samread(...) >= 0)` chain of the reference (clip_reads.h:410, cluster.cpp:48,
// getsv.h:472): every record starts where the previous one ends, so the chain is serial by construction. Here the
// stream is cut into 16 KiB chunks; each chunk GUESSES the first record start at or after its beginning (strong
// plausibility test on the 36-byte fixed part, two records deep). The per-record kernels of getclip / getsv are
// "walkers": one thread per chunk follows its own part of the chain from the guess and does the record work on the way,
// so every record head is fetched from HBM once per pass (ncu: any touch of a record head costs a whole 128-byte line,
// profiles/r1_summary.md). A walker also reports where it left its chunk; exit(c) == guess(c+1) for every chunk proves,
// by induction from the exact header offset, that the walked chain is the true chain - the heuristic affects speed
// only. Wrong guesses are repaired from the predecessor's exit in parallel rounds and the walker is run again.
#include <cub/device/device_scan.cuh>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

__device__ __forceinline__ bool plausible_one(const uint8_t *d, uint64_t n, uint64_t o, int32_t n_ref, uint64_t *next)
{
    if (o + 36 > n) return false;
    int32_t bs = ldi32(d + o);
    if (bs < 33 || o + 4 + (uint64_t)bs > n) return false;
    int32_t tid = ldi32(d + o + 4);
    if (tid < -1 || tid >= n_ref) return false;
    int32_t pos = ldi32(d + o + 8);
    if (pos < -1 || pos >= (1 << 29)) return false;  // BAM coordinates are below 2^29
    uint32_t w = ldu32(d + o + 12);
    uint32_t l_qname = w & 0xff;
    if (l_qname < 2) return false;
    uint32_t w2 = ldu32(d + o + 16);
    uint32_t n_cigar = w2 & 0xffff;
    if ((w2 >> 16) & 0xf000) return false;  // flag bits above 0x800 are not defined
    int32_t l_qseq = ldi32(d + o + 20);
    if (l_qseq < 0) return false;
    int32_t mtid = ldi32(d + o + 24);
    if (mtid < -1 || mtid >= n_ref) return false;
    int32_t mpos = ldi32(d + o + 28);
    if (mpos < -1 || mpos >= (1 << 29)) return false;
    uint64_t need = 32ull + l_qname + 4ull * n_cigar + ((uint64_t)l_qseq + 1) / 2 + (uint64_t)l_qseq;
    if (need > (uint64_t)bs) return false;
    if ((uint64_t)bs - need > 4ull * (uint64_t)l_qseq + 8192) return false;  // aux block of a sane size
    uint8_t c0 = d[o + 36];
    if (c0 < 33 || c0 > 126) return false;           // qname starts with a printable character ...
    if (d[o + 36 + l_qname - 1] != 0) return false;  // ... and is NUL terminated
    *next = o + 4 + (uint64_t)bs;
    return true;
}

// one warp per chunk: lanes test 32 consecutive byte offsets at a time
__global__ void __launch_bounds__(256) guess_starts(const uint8_t *__restrict__ d, uint64_t n, uint64_t first, int32_t n_ref,
                                                    uint64_t n_chunks, uint32_t CHUNK_LOG2, uint64_t *__restrict__ guess)
{
    uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (c >= n_chunks) return;
    uint64_t start = c << CHUNK_LOG2;
    if (start <= first) {
        if (lane == 0) guess[c] = first;
        return;
    }
    uint64_t limit = min(n, start + ((uint64_t)8 << CHUNK_LOG2));
    uint64_t found = BAD_OFFSET;
    for (uint64_t base = start; base < limit; base += 32) {
        uint64_t o = base + lane, nx = 0, nx2 = 0;
        bool ok = plausible_one(d, n, o, n_ref, &nx);
        if (ok && nx < n) ok = plausible_one(d, n, nx, n_ref, &nx2);  // two records deep
        uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (m) {
            found = base + (__ffs(m) - 1);
            break;
        }
    }
    if (lane == 0) guess[c] = found == BAD_OFFSET ? n : found;
}

// plain chain walk of one chunk (count + exit); the getclip / getsv walkers do the same walk with work attached
__device__ __forceinline__ void walk_chunk(const uint8_t *d, uint64_t n, uint64_t entry, uint64_t chunk_end, uint32_t *count,
                                           uint64_t *exit_)
{
    uint64_t o = entry;
    uint32_t k = 0;
    while (o < chunk_end) {
        if (o + 4 > n) break;  // partial tail (shard cut mid-record)
        int32_t bs = ldi32(d + o);
        if (bs < 32) {  // cannot be a record: corrupt chain (or a wrong guess)
            o = BAD_OFFSET;
            break;
        }
        if (o + 4 + (uint64_t)bs > n) break;  // partial tail
        ++k;
        o += 4 + (uint64_t)bs;
    }
    *count = k;
    *exit_ = o;
}

// one thread per chunk: latency-bound pointer chase, hidden by having every chunk in flight at once
__global__ void __launch_bounds__(128) walk_count(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t CHUNK_LOG2,
                                                  const uint64_t *__restrict__ guess, uint32_t *__restrict__ count,
                                                  uint64_t *__restrict__ exit_)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t end = min(n, (c + 1) << CHUNK_LOG2);
    walk_chunk(d, n, guess[c], end, &count[c], &exit_[c]);
}

__global__ void verify_chain(uint64_t n_chunks, const uint64_t *__restrict__ guess, const uint64_t *__restrict__ exit_,
                             uint32_t *__restrict__ bad)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    bool ok = exit_[c] != BAD_OFFSET;
    if (c > 0 && exit_[c - 1] != guess[c]) ok = false;
    if (!ok) atomicOr(bad, 1u);
}

// Repair round: every chunk whose guess differs from its predecessor's exit re-walks from that exit. The first
// mismatching chunk always gets its true entry (its predecessor is correct by induction), so repeating
// verify + repair converges; the number of rounds is the longest run of consecutive wrong chunks (1 in practice).
__global__ void __launch_bounds__(128) repair_chain(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint32_t CHUNK_LOG2,
                                                    uint64_t *guess, uint32_t *count, uint64_t *exit_,
                                                    const uint64_t *__restrict__ exit_prev)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 || c >= n_chunks) return;
    uint64_t entry = exit_prev[c - 1];
    if (entry == BAD_OFFSET || guess[c] == entry) return;
    guess[c] = entry;
    uint64_t end = min(n, (c + 1) << CHUNK_LOG2);
    walk_chunk(d, n, entry, end, &count[c], &exit_[c]);
}

__global__ void count_to_u64(uint64_t n, const uint32_t *__restrict__ in, uint64_t *__restrict__ out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
    if (i == n) out[i] = 0;
}

int exclusive_scan_u64(svb_ctx *ctx, const uint64_t *in, uint64_t *out, uint64_t n)
{
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int64_t)n, ctx->stream));
    DevBuf<uint8_t> t;
    CK(t.alloc(tmp, ctx->stream));
    CK(cub::DeviceScan::ExclusiveSum(t.p, tmp, in, out, (int64_t)n, ctx->stream));
    return 0;
}

int inclusive_scan_u32(svb_ctx *ctx, const uint32_t *in, uint32_t *out, uint64_t n)
{
    size_t tmp = 0;
    CK(cub::DeviceScan::InclusiveSum(nullptr, tmp, in, out, (int64_t)n, ctx->stream));
    DevBuf<uint8_t> t;
    CK(t.alloc(tmp, ctx->stream));
    CK(cub::DeviceScan::InclusiveSum(t.p, tmp, in, out, (int64_t)n, ctx->stream));
    return 0;
}

static inline unsigned nblk(uint64_t n, unsigned b) { return (unsigned)((n + b - 1) / b); }

// guesses only: cheap (touches ~300 bytes per chunk), done when a svb_bam is created
int index_records(svb_ctx *ctx, svb_bam *bam)
{
    uint64_t n = bam->nbytes, first = bam->first;
    if (first > n) return svb_fail(ctx, SVB_ERR_ARG, "first_record beyond the stream");
    {
        const char *e = getenv("SEEKSV_B200_CHUNK_LOG2");
        int v = e ? atoi(e) : 0;
        if (v >= 10 && v <= 20) bam->chunk_log2 = (uint32_t)v;
    }
    const uint32_t CHUNK_LOG2 = bam->chunk_log2;
    uint64_t n_chunks = (n + (1ull << CHUNK_LOG2) - 1) >> CHUNK_LOG2;
    if (n_chunks == 0) n_chunks = 1;
    bam->n_chunks = n_chunks;
    cudaStream_t s = ctx->stream;
    CK(cudaMallocAsync((void **)&bam->d_guess, n_chunks * 8, s));
    CK(cudaMallocAsync((void **)&bam->d_count, n_chunks * 4, s));
    CK(cudaMallocAsync((void **)&bam->d_base, (n_chunks + 1) * 8, s));
    if (!stream_mode(bam)) CKR(ensure_guess(ctx, bam));  // the streaming passes find their own first records
    return 0;
}

bool stream_mode(const svb_bam *bam)
{
    // Measured on C2 (profiles/r1_summary.md): the TMA-staged streaming passes are correct but 4-5x slower than the chunk
    // walkers (clip_stream 1.85 ms vs clip_walk 0.50 ms) - per tile, finding the first record and walking the chain is
    // serial work for one lane while the CTA's other 127 threads wait, and only ~6 tiles fit in an SM's shared memory at a
    // time; the walkers keep ~1000 chains in flight per SM. The walkers are therefore the default.
    const char *e = getenv("SEEKSV_B200_PASS");
    return e && !strcmp(e, "stream") && bam->chunk_log2 == 14;
}

int ensure_guess(svb_ctx *ctx, svb_bam *bam)
{
    if (bam->guessed) return 0;
    cudaStream_t s = ctx->stream;
    {
        ProfScope ps(ctx, "guess_starts", (double)(bam->nbytes - bam->first));
        guess_starts<<<nblk(bam->n_chunks * 32, 256), 256, 0, s>>>(bam->d_data, bam->nbytes, bam->first, bam->n_ref, bam->n_chunks,
                                                                 bam->chunk_log2, bam->d_guess);
    }
    CK(cudaGetLastError());
    bam->guessed = true;
    return 0;
}

// verify exit_[] (filled by a walker that started from bam->d_guess) and, when a guess was wrong, repair guesses and
// counts with plain walks. *ok = 1: the walker's results stand.
int verify_or_repair(svb_ctx *ctx, svb_bam *bam, const uint64_t *d_exit, int *ok)
{
    cudaStream_t s = ctx->stream;
    uint64_t n_chunks = bam->n_chunks;
    DevBuf<uint32_t> flags;
    CK(flags.alloc(1, s));
    CK(cudaMemsetAsync(flags.p, 0, 4, s));
    verify_chain<<<nblk(n_chunks, 256), 256, 0, s>>>(n_chunks, bam->d_guess, d_exit, flags.p);
    uint32_t h = 0;
    CK(cudaMemcpyAsync(&h, flags.p, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *ok = h == 0;
    if (h == 0) return 0;
    // wrong guess somewhere (never seen on well-formed BAMs with the two-record test, but possible in principle)
    DevBuf<uint64_t> ex, snap;
    CK(ex.alloc(n_chunks, s));
    CK(snap.alloc(n_chunks, s));
    walk_count<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, bam->nbytes, n_chunks, bam->chunk_log2, bam->d_guess, bam->d_count, ex.p);
    for (int round = 0;; ++round) {
        CK(cudaMemsetAsync(flags.p, 0, 4, s));
        verify_chain<<<nblk(n_chunks, 256), 256, 0, s>>>(n_chunks, bam->d_guess, ex.p, flags.p);
        CK(cudaMemcpyAsync(&h, flags.p, 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (!h) break;
        if (round >= 256) return svb_fail(ctx, SVB_ERR_FORMAT, "corrupt BAM record chain (block_size < 32)");
        ProfScope ps(ctx, "repair_chain", 0);
        CK(cudaMemcpyAsync(snap.p, ex.p, n_chunks * 8, cudaMemcpyDeviceToDevice, s));
        repair_chain<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, bam->nbytes, n_chunks, bam->chunk_log2, bam->d_guess, bam->d_count, ex.p,
                                                         snap.p);
    }
    bam->counted = false;
    return 0;
}

// verified chain + per-chunk counts + their prefix (needed by getsv's decode and by svb_bam_n_records)
int ensure_counts(svb_ctx *ctx, svb_bam *bam)
{
    if (bam->counted) return 0;
    CKR(ensure_guess(ctx, bam));
    cudaStream_t s = ctx->stream;
    uint64_t n_chunks = bam->n_chunks;
    DevBuf<uint64_t> ex, cnt64;
    CK(ex.alloc(n_chunks, s));
    CK(cnt64.alloc(n_chunks + 1, s));
    for (int attempt = 0;; ++attempt) {
        {
            ProfScope ps(ctx, "walk_count", (double)(bam->nbytes - bam->first));
            walk_count<<<nblk(n_chunks, 128), 128, 0, s>>>(bam->d_data, bam->nbytes, n_chunks, bam->chunk_log2, bam->d_guess, bam->d_count, ex.p);
        }
        int ok = 0;
        CKR(verify_or_repair(ctx, bam, ex.p, &ok));
        if (ok) break;
        if (attempt >= 2) return svb_fail(ctx, SVB_ERR_FORMAT, "record chain does not verify");
    }
    CKR(finish_counts(ctx, bam, ex.p));
    return 0;
}

// counts[] are valid (from walk_count or from a fused walker): prefix, totals
int finish_counts(svb_ctx *ctx, svb_bam *bam, const uint64_t *d_exit)
{
    cudaStream_t s = ctx->stream;
    uint64_t n_chunks = bam->n_chunks;
    DevBuf<uint64_t> cnt64;
    CK(cnt64.alloc(n_chunks + 1, s));
    count_to_u64<<<nblk(n_chunks + 1, 256), 256, 0, s>>>(n_chunks, bam->d_count, cnt64.p);
    CKR(exclusive_scan_u64(ctx, cnt64.p, bam->d_base, n_chunks + 1));
    uint64_t n_rec = 0, last_exit = 0;
    CK(cudaMemcpyAsync(&n_rec, bam->d_base + n_chunks, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&last_exit, d_exit + (n_chunks - 1), 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    bam->n_rec = n_rec;
    bam->rec_bytes = (last_exit >= bam->first && last_exit <= bam->nbytes) ? last_exit - bam->first : 0;
    bam->counted = true;
    if (bam->whole_file && bam->rec_bytes != bam->nbytes - bam->first)
        return svb_fail(ctx, SVB_ERR_FORMAT, "truncated or corrupt BAM: record chain ends %llu bytes early",
                        (unsigned long long)(bam->nbytes - bam->first - bam->rec_bytes));
    return 0;
}
