// gzip on the device: the four getclip texts leave the GPU as ready-to-write multi-member gzip files.
//
// Replaces the ogzstream writes of the reference (gzstream.C:53-114; call sites clip_reads.h:392-395,308-345) for the CLI path:
// the text is produced on the device anyway, so compressing it there shrinks the device->host copy (~45 %) and takes the
// deflate work off the host cores, which the ranks of a multi-GPU run share. Format = what host/bamfile.cpp:write_gz_many
// writes, with smaller members: every 64 KiB piece of the text is a gzip member of its own - a gzip header with an 'SV' extra
// sub-field (member size), ONE dynamic-Huffman deflate block with literals only (RFC 1951 3.2.7), CRC32 and ISIZE. Any gzip reader
// concatenates the members; members of at most 64 KiB of text are also what the device inflate kernel (inflate.cu, made for BGZF
// blocks) takes, so getsv reads P.clip.gz back through the GPU (svb_read_gz_device) instead of the host cores, which are busy
// staging the BAM at that moment. (Until the third session of round 2 the members were 1 MiB = 16 blocks.)
//
//   gz_hist_crc   one CTA per piece: byte histogram (per-warp shared-memory counters) and the piece's raw CRC-32
//   gz_codes      one warp per piece: length-limited Huffman code (<= 15 bits), canonical codes, the block header bits
//   gz_layout     bit offset of every piece inside its member, member sizes and offsets, member CRCs (GF(2) shifts)
//   gz_encode     one CTA per piece: per-thread slices -> bit lengths -> block scan -> bits OR-ed / stored into the output
#include "common.cuh"
#include <mutex>

namespace {

constexpr uint32_t PIECE = 64u << 10, MEMBER = PIECE, PPM = MEMBER / PIECE, GZ_THREADS = 256, SLICE = PIECE / GZ_THREADS;
constexpr uint32_t HDR_WORDS = 48;  // dynamic block header: <= 17 + 57 + 258 * 5 bits = 1364 bits
constexpr uint32_t GZ_HEAD = 20;    // 10 bytes header + XLEN + 'S' 'V' LEN + 4 bytes member size

__constant__ uint32_t c_crc_table[256];
__constant__ uint32_t c_crc_shift[24][32];  // [j][b]: the CRC register 1 << b after 2^j zero bytes
// fixed, complete code for the code-length alphabet: 13 symbols of 4 bits, 6 of 5 bits (as host/bamfile.cpp)
__constant__ uint8_t c_cl_len[19] = {4, 5, 5, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 4, 4};
__constant__ uint8_t c_cl_order2[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

__device__ __forceinline__ uint32_t crc_shift_pow2(uint32_t r, int j)  // r after 2^j zero bytes
{
    uint32_t out = 0;
#pragma unroll 4
    for (int b = 0; b < 32; ++b)
        if (r >> b & 1u) out ^= c_crc_shift[j][b];
    return out;
}
__device__ uint32_t crc_shift(uint32_t r, uint64_t nbytes)  // r after nbytes zero bytes
{
    for (int j = 0; nbytes && j < 24; ++j, nbytes >>= 1)
        if (nbytes & 1) r = crc_shift_pow2(r, j);
    return r;
}

// ---- 1. histogram + raw CRC of every piece ------------------------------------------------------------------------------
__global__ void __launch_bounds__(GZ_THREADS) gz_hist_crc(const uint8_t *__restrict__ text, uint64_t n, uint32_t *__restrict__ hist,
                                                          uint32_t *__restrict__ crc_raw)
{
    // four counter sets per warp (by lane & 3): the lanes of a warp mostly see the same few symbols at the same time
    __shared__ uint32_t h[GZ_THREADS / 32][4][256];
    __shared__ uint32_t tab[256];
    __shared__ uint32_t part[GZ_THREADS];
    const uint32_t t = threadIdx.x, w = t >> 5, sub = t & 3;
    for (uint32_t i = t; i < (GZ_THREADS / 32) * 4 * 256; i += GZ_THREADS) (&h[0][0][0])[i] = 0;
    tab[t] = c_crc_table[t];
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * PIECE;
    const uint64_t piece_len = base < n ? min((uint64_t)PIECE, n - base) : 0;
    const uint64_t a = min(piece_len, (uint64_t)t * SLICE), b = min(piece_len, (uint64_t)(t + 1) * SLICE);
    uint32_t crc = 0;  // raw: register starts at 0, no final inversion
    for (uint64_t i = a; i < b; ++i) {
        const uint32_t c = text[base + i];
        atomicAdd(&h[w][sub][c], 1u);
        crc = tab[(crc ^ c) & 0xff] ^ (crc >> 8);
    }
    part[t] = crc;
    __syncthreads();
    uint32_t s = 0;
    for (uint32_t k = 0; k < GZ_THREADS / 32; ++k) s += h[k][0][t] + h[k][1][t] + h[k][2][t] + h[k][3][t];
    hist[(uint64_t)blockIdx.x * 257 + t] = s;
    if (t == 0) hist[(uint64_t)blockIdx.x * 257 + 256] = 1;  // end of block
    // raw(A || B) = shift(raw(A), |B|) ^ raw(B)
    if (piece_len == PIECE) {  // all slices full: pairwise tree, level j joins neighbours 2^j slices apart (2^(8+j) bytes each)
        for (int j = 0; (1u << j) < GZ_THREADS; ++j) {
            uint32_t r = 0;
            const bool active = (t & ((2u << j) - 1)) == 0;
            if (active) r = crc_shift_pow2(part[t], 8 + j) ^ part[t + (1u << j)];
            __syncthreads();
            if (active) part[t] = r;
            __syncthreads();
        }
        if (t == 0) crc_raw[blockIdx.x] = part[0];
    } else if (t == 0) {  // the last piece of a text: slice lengths differ
        uint32_t r = 0;
        for (uint32_t k = 0; k < GZ_THREADS; ++k) {
            const uint64_t lo = min(piece_len, (uint64_t)k * SLICE), hi = min(piece_len, (uint64_t)(k + 1) * SLICE);
            if (hi == lo) break;
            r = (hi - lo == SLICE ? crc_shift_pow2(r, 8) : crc_shift(r, hi - lo)) ^ part[k];
        }
        crc_raw[blockIdx.x] = r;
    }
}

// ---- 2. Huffman code of every piece ---------------------------------------------------------------------------------------
struct PieceBits {  // LSB-first bit writer into 32-bit words (header construction)
    uint32_t *w;
    uint64_t acc = 0;
    uint32_t nb = 0, nw = 0;
    __device__ void put(uint32_t v, uint32_t n)
    {
        acc |= (uint64_t)v << nb;
        nb += n;
        if (nb >= 32) {
            w[nw++] = (uint32_t)acc;
            acc >>= 32;
            nb -= 32;
        }
    }
    __device__ uint32_t finish()
    {
        uint32_t bits = nw * 32 + nb;
        if (nb) w[nw++] = (uint32_t)acc;
        return bits;
    }
};

constexpr int CODES_WARPS = 4;
struct CodeScratch {  // per warp
    uint32_t freq[260];
    uint32_t wgt[520];     // node weights, then node depths
    int16_t parent[520];
    uint16_t sym[260];     // used symbols, ascending
    uint16_t order[260];   // used symbols by (frequency, symbol)
};

// One warp per piece: the O(alphabet) loops and the sort are spread over the lanes, the inherently serial parts (two-queue
// merge, canonical code assignment, header bits) run on lane 0 out of shared memory.
__global__ void __launch_bounds__(CODES_WARPS * 32)
    gz_codes(uint32_t n_pieces, uint64_t n, const uint32_t *__restrict__ hist, uint32_t *__restrict__ codes, uint32_t *__restrict__ hdr,
             uint32_t *__restrict__ hdr_bits, uint64_t *__restrict__ piece_bits)
{
    __shared__ CodeScratch scratch[CODES_WARPS];
    __shared__ uint8_t len_s[CODES_WARPS][260];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t p = blockIdx.x * CODES_WARPS + wid;
    if (p >= n_pieces) return;  // (whole warps leave together)
    CodeScratch &C = scratch[wid];
    uint8_t *len = len_s[wid];
    const uint32_t *f = hist + (uint64_t)p * 257;
    for (int i = lane; i < 257; i += 32) C.freq[i] = f[i];
    __syncwarp();
    // length-limited prefix code: Huffman by sorted two-queue merge; frequencies are flattened until the longest code fits
    for (;;) {
        int m = 0;
        for (int base = 0; base < 288; base += 32) {  // used symbols, compacted in ascending order
            const int sy = base + (int)lane;
            const bool used = sy < 257 && C.freq[sy] != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, used);
            if (used) C.sym[m + __popc(bal & ((1u << lane) - 1))] = (uint16_t)sy;
            m += __popc(bal);
        }
        for (int i = lane; i < 258; i += 32) len[i] = 0;
        __syncwarp();
        if (m == 1) {
            if (lane == 0) len[C.sym[0]] = 1;
            __syncwarp();
            break;
        }
        for (int a = lane; a < m; a += 32) {  // rank sort by (frequency, symbol)
            const uint32_t fa = C.freq[C.sym[a]];
            int r = 0;
            for (int b = 0; b < m; ++b) {
                const uint32_t fb = C.freq[C.sym[b]];
                r += (fb < fa) || (fb == fa && b < a);
            }
            C.order[r] = C.sym[a];
        }
        __syncwarp();
        for (int i = lane; i < m; i += 32) C.wgt[i] = C.freq[C.order[i]], C.parent[i] = -1;
        __syncwarp();
        int longest = 0;
        if (lane == 0) {
            int qa = 0, qb = m, end = m;
            while (end < 2 * m - 1) {
                int pick[2];
                for (int k = 0; k < 2; ++k) pick[k] = (qa < m && (qb >= end || C.wgt[qa] <= C.wgt[qb])) ? qa++ : qb++;
                C.wgt[end] = C.wgt[pick[0]] + C.wgt[pick[1]];
                C.parent[end] = -1;
                C.parent[pick[0]] = C.parent[pick[1]] = (int16_t)end;
                ++end;
            }
            // depth of a node = depth of its parent + 1; parents have larger indices: walk down from the root
            for (int i = 2 * m - 2; i >= 0; --i) C.wgt[i] = C.parent[i] < 0 ? 0 : C.wgt[C.parent[i]] + 1;
            for (int i = 0; i < m; ++i) {
                len[C.order[i]] = (uint8_t)min(255u, C.wgt[i]);
                longest = max(longest, (int)C.wgt[i]);
            }
        }
        longest = __shfl_sync(0xffffffffu, longest, 0);
        __syncwarp();
        if (longest <= 15) break;
        for (int i = lane; i < 257; i += 32)
            if (C.freq[i]) C.freq[i] = (C.freq[i] + 1) / 2;
        __syncwarp();
    }
    const uint64_t base = (uint64_t)p * PIECE;
    const bool empty_piece = base >= n;
    if (lane == 0) {
        if (empty_piece) len[0] = 1;  // end-of-block alone would be a one-symbol code: give it an unused partner
        len[257] = 1;                 // the single (unused) distance code
        // canonical codes, bit-reversed for LSB-first output (into C.wgt: code | length << 16)
        uint32_t count[16], next[16];
        for (int i = 0; i < 16; ++i) count[i] = 0;
        for (int i = 0; i < 257; ++i) count[len[i]]++;
        count[0] = 0;
        uint32_t c = 0;
        next[0] = 0;
        for (int b = 1; b <= 15; ++b) {
            c = (c + count[b - 1]) << 1;
            next[b] = c;
        }
        for (int i = 0; i < 257; ++i) {
            const uint32_t l = len[i], v = l ? next[l]++ : 0;
            C.wgt[i] = (l ? __brev(v) >> (32 - l) : 0) | l << 16;
        }
    }
    __syncwarp();
    uint64_t data_bits = 0;
    for (int i = lane; i < 257; i += 32) {
        const uint32_t cl = C.wgt[i];
        codes[(uint64_t)p * 257 + i] = cl;
        data_bits += (uint64_t)f[i] * (cl >> 16);  // (true frequencies, not the flattened ones)
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) data_bits += __shfl_xor_sync(0xffffffffu, data_bits, d);
    if (lane != 0) return;
    // the block header (RFC 1951 3.2.7) with the fixed code-length code
    uint32_t cl_code[19];
    {
        uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nx[8];
        for (int i = 0; i < 19; ++i) cnt[c_cl_len[i]]++;
        uint32_t cc = 0;
        nx[0] = 0;
        for (int b = 1; b < 8; ++b) {
            cc = (cc + cnt[b - 1]) << 1;
            nx[b] = cc;
        }
        for (int i = 0; i < 19; ++i) {
            const uint32_t l = c_cl_len[i];
            cl_code[i] = __brev(nx[l]++) >> (32 - l);
        }
    }
    PieceBits bw;
    bw.w = hdr + (uint64_t)p * HDR_WORDS;
    const bool last_of_member = (p % PPM) == PPM - 1 || p + 1 == n_pieces;
    bw.put(last_of_member ? 1 : 0, 1);
    bw.put(2, 2);   // dynamic Huffman
    bw.put(0, 5);   // HLIT: 257 literal/length codes
    bw.put(0, 5);   // HDIST: 1 distance code
    bw.put(15, 4);  // HCLEN: 19 code-length codes
    for (int i = 0; i < 19; ++i) bw.put(c_cl_len[c_cl_order2[i]], 3);
    for (int i = 0; i < 258;) {
        if (len[i] == 0) {
            int r = 1;
            while (i + r < 258 && len[i + r] == 0 && r < 138) ++r;
            if (r >= 11) {
                bw.put(cl_code[18], c_cl_len[18]);
                bw.put(r - 11, 7);
            } else if (r >= 3) {
                bw.put(cl_code[17], c_cl_len[17]);
                bw.put(r - 3, 3);
            } else {
                r = 1;
                bw.put(cl_code[0], c_cl_len[0]);
            }
            i += r;
        } else {
            bw.put(cl_code[len[i]], c_cl_len[len[i]]);
            ++i;
        }
    }
    const uint32_t hb = bw.finish();
    hdr_bits[p] = hb;
    piece_bits[p] = hb + data_bits;
}

// ---- 3. layout: bit offset of every piece in its member, member sizes / offsets / CRCs -----------------------------------
__global__ void gz_layout(uint32_t n_pieces, uint32_t n_members, uint64_t n, const uint64_t *__restrict__ piece_bits,
                          const uint32_t *__restrict__ crc_raw, uint64_t *__restrict__ piece_bit_off, uint64_t *__restrict__ member_off,
                          uint32_t *__restrict__ member_crc)
{
    for (uint32_t m = threadIdx.x; m < n_members; m += blockDim.x) {
        uint64_t bit = (uint64_t)GZ_HEAD * 8;
        uint32_t raw = 0;
        uint64_t member_len = 0;
        for (uint32_t p = m * PPM; p < min(n_pieces, (m + 1) * PPM); ++p) {
            piece_bit_off[p] = bit;
            bit += piece_bits[p];
            const uint64_t base = (uint64_t)p * PIECE, plen = base < n ? min((uint64_t)PIECE, n - base) : 0;
            raw = crc_shift(raw, plen) ^ crc_raw[p];
            member_len += plen;
        }
        member_off[m + 1] = (bit + 7) / 8 + 8;  // (size for now; prefix below)
        member_crc[m] = ~(raw ^ crc_shift(0xffffffffu, member_len));
    }
    __syncthreads();
    // member sizes -> offsets: every thread adds up a contiguous share, the shares are scanned in shared memory
    __shared__ uint64_t share[1024];
    const uint32_t per = (n_members + blockDim.x - 1) / blockDim.x;
    const uint32_t m0 = min(n_members, threadIdx.x * per), m1 = min(n_members, m0 + per);
    uint64_t mine = 0;
    for (uint32_t m = m0; m < m1; ++m) mine += member_off[m + 1];
    share[threadIdx.x] = mine;
    __syncthreads();
    for (uint32_t d = 1; d < blockDim.x; d <<= 1) {
        const uint64_t v = threadIdx.x >= d ? share[threadIdx.x - d] : 0;
        __syncthreads();
        share[threadIdx.x] += v;
        __syncthreads();
    }
    uint64_t acc = share[threadIdx.x] - mine;
    if (threadIdx.x == 0) member_off[0] = 0;
    for (uint32_t m = m0; m < m1; ++m) {
        acc += member_off[m + 1];
        member_off[m + 1] = acc;
    }
}

// ---- 4. encode ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void or_bits(uint32_t *out, uint64_t bit, uint64_t v, uint32_t nbits)  // nbits <= 32, any alignment
{
    if (!nbits) return;
    const uint64_t w = bit >> 5;
    const uint32_t sh = (uint32_t)(bit & 31);
    const uint64_t x = (v & ((nbits == 64) ? ~0ull : ((1ull << nbits) - 1))) << sh;
    atomicOr(out + w, (uint32_t)x);
    if (sh + nbits > 32) atomicOr(out + w + 1, (uint32_t)(x >> 32));
}

__global__ void __launch_bounds__(GZ_THREADS)
    gz_encode(const uint8_t *__restrict__ text, uint64_t n, uint32_t n_pieces, const uint32_t *__restrict__ codes, const uint32_t *__restrict__ hdr,
              const uint32_t *__restrict__ hdr_bits, const uint64_t *__restrict__ piece_bits, const uint64_t *__restrict__ piece_bit_off,
              const uint64_t *__restrict__ member_off, const uint32_t *__restrict__ member_crc, uint32_t *__restrict__ out)
{
    __shared__ uint32_t code[257];
    __shared__ uint32_t warp_sum[GZ_THREADS / 32];
    const uint32_t p = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5, m = p / PPM;
    for (uint32_t i = t; i < 257; i += GZ_THREADS) code[i] = codes[(uint64_t)p * 257 + i];
    __syncthreads();
    const uint64_t base = (uint64_t)p * PIECE;
    const uint64_t piece_len = base < n ? min((uint64_t)PIECE, n - base) : 0;
    const uint64_t a = min(piece_len, (uint64_t)t * SLICE), b = min(piece_len, (uint64_t)(t + 1) * SLICE);
    const uint32_t t_eob = piece_len ? (uint32_t)((piece_len - 1) / SLICE) : 0;
    // bit length of this thread's slice (+ the end-of-block code on the last slice)
    uint32_t bits = 0;
    for (uint64_t i = a; i < b; ++i) bits += code[text[base + i]] >> 16;
    if (t == t_eob) bits += code[256] >> 16;
    // exclusive scan over the CTA
    uint32_t incl = bits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += v;
    }
    if (lane == 31) warp_sum[w] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t k = 0; k < w; ++k) before += warp_sum[k];
    const uint64_t member_bit = member_off[m] * 8;
    const uint64_t piece_bit = member_bit + piece_bit_off[p];
    uint64_t bit = piece_bit + hdr_bits[p] + before + (incl - bits);
    if (bits) {
        // first and last word of the slice are shared with the neighbours (OR), the words in between are ours alone (store)
        uint64_t acc = 0;
        uint32_t nb = (uint32_t)(bit & 31);
        uint64_t wi = bit >> 5;
        bool first = true;
        auto push = [&](uint32_t cl) {
            acc |= (uint64_t)(cl & 0xffff) << nb;
            nb += cl >> 16;
            if (nb >= 32) {
                if (first) atomicOr(out + wi, (uint32_t)acc), first = false;
                else out[wi] = (uint32_t)acc;
                ++wi;
                acc >>= 32;
                nb -= 32;
            }
        };
        for (uint64_t i = a; i < b; ++i) push(code[text[base + i]]);
        if (t == t_eob) push(code[256]);
        if (nb) atomicOr(out + wi, (uint32_t)acc);
    }
    if (t == 0) {
        const uint32_t hb = hdr_bits[p];
        const uint32_t *h = hdr + (uint64_t)p * HDR_WORDS;
        for (uint32_t k = 0; k * 32 < hb; ++k) or_bits(out, piece_bit + (uint64_t)k * 32, h[k], min(32u, hb - k * 32));
        if (p % PPM == 0) {  // the member's gzip header
            const uint32_t size = (uint32_t)(member_off[m + 1] - member_off[m]);
            const uint8_t head[GZ_HEAD] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 3, 8, 0, 'S', 'V', 4, 0,
                                           (uint8_t)size, (uint8_t)(size >> 8), (uint8_t)(size >> 16), (uint8_t)(size >> 24)};
            for (uint32_t k = 0; k < GZ_HEAD; ++k) or_bits(out, member_bit + 8ull * k, head[k], 8);
        }
        if (p % PPM == PPM - 1 || p + 1 == n_pieces) {  // the member's trailer: CRC32, ISIZE
            const uint64_t end_byte = (piece_bit + piece_bits[p] + 7) / 8;
            const uint64_t first_piece = (uint64_t)m * PPM;
            const uint64_t member_len = min(n, ((uint64_t)p + 1) * PIECE) - min(n, first_piece * PIECE);
            or_bits(out, end_byte * 8, member_crc[m], 32);
            or_bits(out, end_byte * 8 + 32, (uint32_t)member_len, 32);
        }
    }
}

void crc_tables(uint32_t *table, uint32_t (*shift)[32])
{
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1;
        table[i] = c;
    }
    for (int b = 0; b < 32; ++b) {  // one zero byte
        uint32_t r = 1u << b;
        shift[0][b] = table[r & 0xff] ^ (r >> 8);
    }
    for (int j = 1; j < 24; ++j)  // squaring: 2^j zero bytes = twice 2^(j-1)
        for (int b = 0; b < 32; ++b) {
            uint32_t r = shift[j - 1][b], o = 0;
            for (int k = 0; k < 32; ++k)
                if (r >> k & 1u) o ^= shift[j - 1][k];
            shift[j][b] = o;
        }
}

}  // namespace

// d_text[0, n) on the device -> gzip file image in pinned host memory (out). Runs on ctx->stream; synchronises it.
int gzip_on_device(svb_ctx *ctx, const char *d_text, uint64_t n, PinnedBuf *out)
{
    if (!ctx->gz_tables_ready) {  // constant memory is per device: once per context
        static uint32_t table[256], shift[24][32];
        static std::once_flag once;
        std::call_once(once, [] { crc_tables(table, shift); });
        CK(cudaMemcpyToSymbol(c_crc_table, table, sizeof table));
        CK(cudaMemcpyToSymbol(c_crc_shift, shift, sizeof shift));
        ctx->gz_tables_ready = true;
    }
    cudaStream_t s = ctx->stream;
    const uint32_t n_pieces = (uint32_t)std::max<uint64_t>(1, (n + PIECE - 1) / PIECE), n_members = (n_pieces + PPM - 1) / PPM;
    DevBuf<uint32_t> hist, crc_raw, codes, hdr, hdr_bits, member_crc;
    DevBuf<uint64_t> piece_bits, piece_bit_off, member_off;
    CK(hist.alloc((uint64_t)n_pieces * 257, s));
    CK(crc_raw.alloc(n_pieces, s));
    CK(codes.alloc((uint64_t)n_pieces * 257, s));
    CK(hdr.alloc((uint64_t)n_pieces * HDR_WORDS, s));
    CK(hdr_bits.alloc(n_pieces, s));
    CK(piece_bits.alloc(n_pieces, s));
    CK(piece_bit_off.alloc(n_pieces, s));
    CK(member_off.alloc(n_members + 1, s));
    CK(member_crc.alloc(n_members, s));
    {
        ProfScope ps(ctx, "gz_hist_crc", (double)n);
        gz_hist_crc<<<n_pieces, GZ_THREADS, 0, s>>>((const uint8_t *)d_text, n, hist.p, crc_raw.p);
    }
    {
        ProfScope ps(ctx, "gz_codes", 0);
        gz_codes<<<(n_pieces + CODES_WARPS - 1) / CODES_WARPS, CODES_WARPS * 32, 0, s>>>(n_pieces, n, hist.p, codes.p, hdr.p, hdr_bits.p, piece_bits.p);
    }
    {
        ProfScope ps(ctx, "gz_layout", 0);
        gz_layout<<<1, 1024, 0, s>>>(n_pieces, n_members, n, piece_bits.p, crc_raw.p, piece_bit_off.p, member_off.p, member_crc.p);
    }
    uint64_t total = 0;
    CK(cudaMemcpyAsync(&total, member_off.p + n_members, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    DevBuf<uint32_t> gz;
    CK(gz.alloc(total / 4 + 2, s));
    CK(cudaMemsetAsync(gz.p, 0, (total / 4 + 2) * 4, s));
    {
        ProfScope ps(ctx, "gz_encode", (double)n);
        gz_encode<<<n_pieces, GZ_THREADS, 0, s>>>((const uint8_t *)d_text, n, n_pieces, codes.p, hdr.p, hdr_bits.p, piece_bits.p,
                                                  piece_bit_off.p, member_off.p, member_crc.p, gz.p);
    }
    CK(cudaGetLastError());
    CKR(out->reserve(ctx, total));
    CK(cudaMemcpyAsync(out->p, gz.p, total, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return 0;
}
