// Record-boundary resolution for a packed BAM record stream resident in HBM.
//
// Replaces the serial `while (samread(...) >= 0)` chain of the reference (clip_reads.h:410,
// cluster.cpp:48, getsv.h:472): every record starts where the previous one ends, so the chain is serial
// by construction. Here the stream is cut into fixed chunks; each chunk GUESSES the first record start
// at or after its beginning (strong plausibility test on the 36-byte fixed part, two records deep), walks
// its own part of the chain, and a verification pass checks that every chunk's exit offset equals the
// next chunk's guess. By induction from chunk 0 (whose entry is exact: the header length) a verified
// chain is the true chain - the heuristic only affects speed, never the result. Chunks that fail are
// re-walked from their predecessor's exit in repair rounds until the whole chain verifies.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

static constexpr uint32_t CHUNK_LOG2 = 14;  // 16 KiB chunks: ~50 records of 320 B per walker thread
static constexpr uint64_t CHUNK = 1ull << CHUNK_LOG2;
static constexpr uint64_t BAD = ~0ull;

__device__ __forceinline__ bool plausible_one(const uint8_t *d, uint64_t n, uint64_t o, int32_t n_ref, uint64_t *next)
{
    if (o + 36 > n) return false;
    int32_t bs = ldi32s(d + o);
    if (bs < 33 || o + 4 + (uint64_t)bs > n) return false;
    int32_t tid = ldi32s(d + o + 4);
    if (tid < -1 || tid >= n_ref) return false;
    int32_t pos = ldi32s(d + o + 8);
    if (pos < -1 || pos >= (1 << 29)) return false;  // BAM coordinates are below 2^29
    uint32_t w = ldu32s(d + o + 12);
    uint32_t l_qname = w & 0xff;
    if (l_qname < 2) return false;
    uint32_t w2 = ldu32s(d + o + 16);
    uint32_t n_cigar = w2 & 0xffff;
    if ((w2 >> 16) & 0xf000) return false;  // flag bits above 0x800 are not defined
    int32_t l_qseq = ldi32s(d + o + 20);
    if (l_qseq < 0) return false;
    int32_t mtid = ldi32s(d + o + 24);
    if (mtid < -1 || mtid >= n_ref) return false;
    int32_t mpos = ldi32s(d + o + 28);
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
__global__ void __launch_bounds__(256) guess_starts(const uint8_t *__restrict__ d, uint64_t n, uint64_t first,
                                                    int32_t n_ref, uint64_t n_chunks, uint64_t *__restrict__ guess)
{
    uint64_t c = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (c >= n_chunks) return;
    uint64_t start = c << CHUNK_LOG2;
    if (start <= first) {
        if (lane == 0) guess[c] = first;
        return;
    }
    uint64_t limit = min(n, start + 8 * CHUNK);
    uint64_t found = BAD;
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
    if (lane == 0) guess[c] = found == BAD ? n : found;
}

__device__ __forceinline__ void walk_chunk(const uint8_t *d, uint64_t n, uint64_t entry, uint64_t chunk_end,
                                           uint32_t *count, uint64_t *exit_, uint64_t *rec_off)
{
    uint64_t o = entry;
    uint32_t k = 0;
    while (o < chunk_end) {
        if (o + 4 > n) break;  // partial tail (shard cut mid-record)
        int32_t bs = ldi32s(d + o);
        if (bs < 32) {  // cannot be a record: corrupt chain (or a wrong guess)
            o = BAD;
            break;
        }
        if (o + 4 + (uint64_t)bs > n) break;  // partial tail
        if (rec_off) rec_off[k] = o;
        ++k;
        o += 4 + (uint64_t)bs;
    }
    if (count) *count = k;
    if (exit_) *exit_ = o;
}

// one thread per chunk: latency-bound pointer chase, hidden by having every chunk in flight at once
__global__ void __launch_bounds__(128) walk_count(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks,
                                                  const uint64_t *__restrict__ guess, uint32_t *__restrict__ count,
                                                  uint64_t *__restrict__ exit_)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t end = min(n, (c + 1) << CHUNK_LOG2);
    walk_chunk(d, n, guess[c], end, &count[c], &exit_[c], nullptr);
}

__global__ void verify_chain(uint64_t n_chunks, const uint64_t *__restrict__ guess, const uint64_t *__restrict__ exit_,
                             uint32_t *__restrict__ bad)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    bool ok = exit_[c] != BAD;
    if (c > 0 && exit_[c - 1] != guess[c]) ok = false;
    if (!ok) atomicOr(bad, 1u);
}

// Repair round: every chunk whose guess differs from its predecessor's exit re-walks from that exit. The first
// mismatching chunk always gets its true entry (its predecessor is correct by induction), so repeating
// verify + repair converges; the number of rounds is the longest run of consecutive wrong chunks (1 in practice).
__global__ void __launch_bounds__(128) repair_chain(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks, uint64_t *guess,
                                                    uint32_t *count, uint64_t *exit_, const uint64_t *__restrict__ exit_prev)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 || c >= n_chunks) return;
    uint64_t entry = exit_prev[c - 1];
    if (entry == BAD || guess[c] == entry) return;
    guess[c] = entry;
    uint64_t end = min(n, (c + 1) << CHUNK_LOG2);
    walk_chunk(d, n, entry, end, &count[c], &exit_[c], nullptr);
}

__global__ void __launch_bounds__(128) walk_write(const uint8_t *__restrict__ d, uint64_t n, uint64_t n_chunks,
                                                  const uint64_t *__restrict__ guess, const uint64_t *__restrict__ base,
                                                  uint64_t *__restrict__ rec_off)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t end = min(n, (c + 1) << CHUNK_LOG2);
    walk_chunk(d, n, guess[c], end, nullptr, nullptr, rec_off + base[c]);
}

__global__ void count_to_u64(uint64_t n, const uint32_t *__restrict__ in, uint64_t *__restrict__ out)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
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

int index_records(svb_ctx *ctx, svb_bam *bam)
{
    const uint8_t *d = bam->d_data;
    uint64_t n = bam->nbytes, first = bam->first;
    if (first > n) return svb_fail(ctx, SVB_ERR_ARG, "first_record beyond the stream");
    uint64_t n_chunks = (n + CHUNK - 1) >> CHUNK_LOG2;
    if (n_chunks == 0) n_chunks = 1;
    cudaStream_t s = ctx->stream;
    DevBuf<uint64_t> guess, exit_, cnt64, base;
    DevBuf<uint32_t> count, flags;
    CK(guess.alloc(n_chunks, s));
    CK(exit_.alloc(n_chunks, s));
    CK(count.alloc(n_chunks, s));
    CK(cnt64.alloc(n_chunks + 1, s));
    CK(base.alloc(n_chunks + 1, s));
    CK(flags.alloc(2, s));
    CK(cudaMemsetAsync(flags.p, 0, 8, s));
    {
        ProfScope ps(ctx, "guess_starts", 0);
        uint64_t threads = n_chunks * 32;
        guess_starts<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d, n, first, bam->n_ref, n_chunks, guess.p);
    }
    {
        ProfScope ps(ctx, "walk_count", 0);
        walk_count<<<(unsigned)((n_chunks + 127) / 128), 128, 0, s>>>(d, n, n_chunks, guess.p, count.p, exit_.p);
    }
    verify_chain<<<(unsigned)((n_chunks + 255) / 256), 256, 0, s>>>(n_chunks, guess.p, exit_.p, flags.p);
    uint32_t hflags[2];
    CK(cudaMemcpyAsync(hflags, flags.p, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint64_t round = 0; hflags[0]; ++round) {
        // a wrong guess (never seen on well-formed BAMs with the two-record test, but possible in principle)
        if (round >= 256) return svb_fail(ctx, SVB_ERR_FORMAT, "corrupt BAM record chain (block_size < 32)");
        ProfScope ps(ctx, "repair_chain", 0);
        DevBuf<uint64_t> snap;
        CK(snap.alloc(n_chunks, s));
        CK(cudaMemcpyAsync(snap.p, exit_.p, n_chunks * 8, cudaMemcpyDeviceToDevice, s));
        repair_chain<<<(unsigned)((n_chunks + 127) / 128), 128, 0, s>>>(d, n, n_chunks, guess.p, count.p, exit_.p, snap.p);
        CK(cudaMemsetAsync(flags.p, 0, 8, s));
        verify_chain<<<(unsigned)((n_chunks + 255) / 256), 256, 0, s>>>(n_chunks, guess.p, exit_.p, flags.p);
        CK(cudaMemcpyAsync(hflags, flags.p, 8, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    CK(cudaMemsetAsync(cnt64.p + n_chunks, 0, 8, s));
    count_to_u64<<<(unsigned)((n_chunks + 255) / 256), 256, 0, s>>>(n_chunks, count.p, cnt64.p);
    CKR(exclusive_scan_u64(ctx, cnt64.p, base.p, n_chunks + 1));
    uint64_t n_rec = 0, last_exit = 0;
    CK(cudaMemcpyAsync(&n_rec, base.p + n_chunks, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&last_exit, exit_.p + (n_chunks - 1), 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    bam->n_rec = n_rec;
    bam->rec_bytes = (last_exit >= first && last_exit <= n) ? last_exit - first : 0;
    CK(cudaMallocAsync((void **)&bam->d_rec_off, (n_rec + 1) * sizeof(uint64_t), s));
    {
        ProfScope ps(ctx, "walk_write", 0);
        walk_write<<<(unsigned)((n_chunks + 127) / 128), 128, 0, s>>>(d, n, n_chunks, guess.p, base.p, bam->d_rec_off);
    }
    CK(cudaMemcpyAsync(bam->d_rec_off + n_rec, &last_exit, 8, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}
