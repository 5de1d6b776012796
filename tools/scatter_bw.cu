// Micro-benchmark: how fast does this GPU deliver 128-byte lines when they are asked for the way the record walkers ask for them?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scatter_bw tools/scatter_bw.cu && ./scatter_bw [GB=2.7]
//
// The full-pass kernels (walk_count / clip_walk / decode_walk) run one thread per 16 KiB chunk; every thread hops through its
// chunk in record-sized steps (~295 bytes) and reads 4-50 bytes at each stop, so a warp instruction touches 32 lines that lie
// 16 KiB apart and every hop depends on the previous one only through the address. This program measures the line throughput
// (lines touched x 128 bytes / time) of that pattern without any record logic, next to a plain streaming read:
//   stream      : every thread reads consecutive 16-byte words (coalesced), the whole buffer once
//   hop         : one thread per chunk, one 4-byte load every `step` bytes, address independent of the data (all loads of a chain
//                 can be in flight at once) - the ceiling for scattered lines
//   chase       : the same, but the next address comes out of the loaded word (as block_size does): the dependent-load form
//   chase_ahead : the dependent form with the next stop requested before the current one is consumed (the walkers' pipelining)
// for chunk sizes 4 / 8 / 16 KiB. Prints one JSON line per variant. No product code depends on this file.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__global__ void fill_steps(uint32_t *buf, uint64_t n_words, uint32_t step)
{
    // every word holds the hop length that a record's block_size would give: step +- 16 bytes, multiple of 4
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_words) buf[i] = step + (((uint32_t)(i * 2654435761u) >> 27) & ~3u) - 16u;
}

__global__ void k_stream(const uint4 *__restrict__ buf, uint64_t n16, unsigned long long *sink)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (; i < n16; i += stride) {
        uint4 v = __ldg(buf + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

template <int MODE>  // 0 hop, 1 chase, 2 chase with one stop of look-ahead
__global__ void k_walk(const uint8_t *__restrict__ buf, uint64_t n_chunks, uint32_t chunk_log2, uint32_t step, unsigned long long *sink,
                       unsigned long long *stops)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    uint64_t o = c << chunk_log2, end = (c + 1) << chunk_log2;
    uint32_t acc = 0, n = 0;
    if (MODE == 0) {
        for (; o < end; o += step, ++n) acc += __ldg((const uint32_t *)(buf + o));
    } else if (MODE == 1) {
        while (o < end) {
            uint32_t v = __ldg((const uint32_t *)(buf + o));
            acc += v, ++n;
            o += v;
        }
    } else {
        uint32_t v = __ldg((const uint32_t *)(buf + o));
        while (o < end) {
            uint64_t on = o + v;
            uint32_t vn = on < end ? __ldg((const uint32_t *)(buf + on)) : 0;
            // a second, independent load per stop (the walkers read the CIGAR / flag words of the current record here)
            acc += __ldg((const uint32_t *)(buf + o + 36)) + v;
            ++n;
            o = on, v = vn;
        }
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
    atomicAdd(stops, (unsigned long long)n);
}

int main(int argc, char **argv)
{
    double gb = argc > 1 ? atof(argv[1]) : 2.7;
    uint64_t bytes = ((uint64_t)(gb * 1e9) >> 16) << 16;
    const uint32_t step = 296;
    uint8_t *buf;
    unsigned long long *ctr;
    CK(cudaMalloc(&buf, bytes + 4096));
    CK(cudaMalloc(&ctr, 16));
    fill_steps<<<(unsigned)((bytes / 4 + 255) / 256), 256>>>((uint32_t *)buf, bytes / 4 + 1024, step);
    CK(cudaDeviceSynchronize());
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    auto timed = [&](auto launch) {
        float best = 1e30f;
        for (int rep = 0; rep < 7; ++rep) {
            CK(cudaMemset(ctr, 0, 16));
            CK(cudaEventRecord(a));
            launch();
            CK(cudaEventRecord(b));
            CK(cudaEventSynchronize(b));
            float ms;
            CK(cudaEventElapsedTime(&ms, a, b));
            if (rep >= 2 && ms < best) best = ms;
        }
        CK(cudaGetLastError());
        return best;
    };
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float ms = timed([&] { k_stream<<<sms * 16, 512>>>((const uint4 *)buf, bytes / 16, ctr); });
    printf("{\"variant\": \"stream\", \"ms\": %.4f, \"GBps\": %.1f}\n", ms, bytes / (ms * 1e-3) / 1e9);
    for (uint32_t log2c = 12; log2c <= 14; ++log2c) {
        uint64_t n_chunks = bytes >> log2c;
        unsigned grid = (unsigned)((n_chunks + 127) / 128);
        for (int mode = 0; mode < 3; ++mode) {
            ms = timed([&] {
                if (mode == 0) k_walk<0><<<grid, 128>>>(buf, n_chunks, log2c, step, ctr, ctr + 1);
                else if (mode == 1) k_walk<1><<<grid, 128>>>(buf, n_chunks, log2c, step, ctr, ctr + 1);
                else k_walk<2><<<grid, 128>>>(buf, n_chunks, log2c, step, ctr, ctr + 1);
            });
            unsigned long long h[2];
            CK(cudaMemcpy(h, ctr, 16, cudaMemcpyDeviceToHost));
            // lines touched: a stop every ~296 bytes touches 1 line (4-byte load) or up to 2 (the +36 load); report stops and time,
            // and the line throughput under the one-line-per-stop assumption (44.5 % of the buffer at 128-byte lines for step 296)
            double lines = (double)h[1];
            printf("{\"variant\": \"%s\", \"chunk\": %u, \"ms\": %.4f, \"stops\": %llu, \"line_GBps\": %.1f, \"buffer_GBps\": %.1f}\n",
                   mode == 0 ? "hop" : mode == 1 ? "chase" : "chase_ahead", 1u << log2c, ms, h[1], lines * 128.0 / (ms * 1e-3) / 1e9,
                   bytes / (ms * 1e-3) / 1e9);
        }
    }
    return 0;
}
