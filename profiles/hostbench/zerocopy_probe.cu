// How should the 1 479 marker columns of a pinned host matrix [rows x 20000] f32 reach the GPU?
//   A. cudaMemcpyAsync of the full rows (DMA engine), columns picked on the device afterwards
//   B. zero-copy gather: a kernel reads only the marker columns straight from mapped pinned memory
//   C. zero-copy stream: a kernel reads the full rows from mapped pinned memory (16-byte loads)
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/zerocopy_probe zerocopy_probe.cu && /tmp/zerocopy_probe [rows] [device]
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void gather_kernel(const float *__restrict__ x, int64_t ld, int64_t rows, const int32_t *__restrict__ cols,
                              int kept, float *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += n_warps) {
        const float *s = x + r * ld;
        float *d = dst + r * kept;
        for (int j0 = 0; j0 < kept; j0 += 32 * 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = j0 + u * 32 + lane; v[u] = j < kept ? s[cols[j]] : 0.f; }
#pragma unroll
            for (int u = 0; u < 8; ++u) { const int j = j0 + u * 32 + lane; if (j < kept) d[j] = v[u]; }
        }
    }
}

__global__ void stream_kernel(const float4 *__restrict__ x, int64_t n4, float4 *__restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < n4) v[u] = x[i + u * stride];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i + u * stride < n4) dst[i + u * stride] = v[u];
    }
}

int main(int argc, char **argv) {
    const int64_t rows = argc > 1 ? atoll(argv[1]) : 50000, ld = 20000;
    const int dev = argc > 2 ? atoi(argv[2]) : 0, kept = 1479;
    CK(cudaSetDevice(dev));
    float *hx = nullptr;
    CK(cudaHostAlloc(&hx, (size_t)rows * ld * 4, cudaHostAllocDefault));
    {
        const int nt = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t)
            pool.emplace_back([&, t] { for (int64_t r = rows * t / nt; r < rows * (t + 1) / nt; ++r) for (int64_t j = 0; j < ld; ++j) hx[r * ld + j] = (float)((r + j) & 7); });
        for (auto &th : pool) th.join();
    }
    std::vector<int32_t> cols(ld);
    for (int j = 0; j < ld; ++j) cols[j] = j;
    std::mt19937 rng(1);
    std::shuffle(cols.begin(), cols.end(), rng);
    cols.resize(kept);
    std::sort(cols.begin(), cols.end());
    int sectors = 0, lines64 = 0, lines128 = 0;
    for (int j = 0; j < kept; ++j) {
        sectors += j == 0 || cols[j] / 8 != cols[j - 1] / 8;
        lines64 += j == 0 || cols[j] / 16 != cols[j - 1] / 16;
        lines128 += j == 0 || cols[j] / 32 != cols[j - 1] / 32;
    }
    printf("rows %lld: per row %d columns touch %d 32-B sectors (%.1f KB), %d 64-B lines (%.1f KB), %d 128-B lines (%.1f KB) of %.1f KB\n",
           (long long)rows, kept, sectors, sectors * 32 / 1e3, lines64, lines64 * 64 / 1e3, lines128, lines128 * 128 / 1e3, ld * 4 / 1e3);
    int32_t *dcols; float *ddst, *dfull;
    const int64_t chunk = 4096;
    CK(cudaMalloc(&dcols, kept * 4));
    CK(cudaMemcpy(dcols, cols.data(), kept * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ddst, (size_t)rows * kept * 4));
    CK(cudaMalloc(&dfull, (size_t)2 * chunk * ld * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaStream_t st; cudaStreamCreate(&st);
    const double gb = (double)rows * ld * 4 / 1e9;
    for (int rep = 0; rep < 3; ++rep) {
        float ms;
        cudaEventRecord(e0, st);
        for (int64_t r = 0, i = 0; r < rows; r += chunk, ++i)
            cudaMemcpyAsync(dfull + (i & 1) * chunk * ld, hx + r * ld, (size_t)std::min(chunk, rows - r) * ld * 4, cudaMemcpyHostToDevice, st);
        cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("A  DMA full rows:          %8.2f ms  %6.1f GB/s over PCIe (= of matrix)\n", ms, gb / ms * 1e3);
        for (int blocks : {148, 148 * 4, 148 * 16}) {
            cudaEventRecord(e0, st);
            gather_kernel<<<blocks, 256, 0, st>>>(hx, ld, rows, dcols, kept, ddst);
            cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            printf("B  zero-copy gather %5d CTAs: %8.2f ms  %6.1f GB/s of matrix, %5.1f GB/s of 32-B sectors, %5.1f GB/s useful\n", blocks, ms,
                   gb / ms * 1e3, (double)rows * sectors * 32 / 1e6 / ms, (double)rows * kept * 4 / 1e6 / ms);
        }
        const int64_t n4 = std::min<int64_t>(rows, 2 * chunk) * ld / 4;
        cudaEventRecord(e0, st);
        stream_kernel<<<148 * 8, 256, 0, st>>>((const float4 *)hx, n4, (float4 *)dfull);
        cudaEventRecord(e1, st); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf("C  zero-copy stream (16 B): %8.2f ms  %6.1f GB/s\n", ms, (double)n4 * 16 / 1e6 / ms);
    }
    CK(cudaGetLastError());
    return 0;
}
