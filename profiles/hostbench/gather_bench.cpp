// Host-side marker-column gather: how fast can 16 threads pull 1 479 scattered columns out of a 100k x 20k f32 matrix?
//   g++ -O3 -march=native -pthread -o /tmp/gather_bench gather_bench.cpp && /tmp/gather_bench
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>
#include <xmmintrin.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <typename F> static double run(int n_thr, int64_t rows, F f) {
    const double t0 = now();
    std::vector<std::thread> pool;
    for (int t = 0; t < n_thr; ++t) pool.emplace_back(f, rows * t / n_thr, rows * (t + 1) / n_thr);
    for (auto &th : pool) th.join();
    return now() - t0;
}

int main(int argc, char **argv) {
    const int64_t rows = argc > 1 ? atoll(argv[1]) : 100000, ld = 20000;
    const int n_kept = 1479, n_thr = argc > 2 ? atoi(argv[2]) : (int)std::thread::hardware_concurrency();
    float *x = (float *)aligned_alloc(4096, (size_t)rows * ld * 4);
    float *dst = (float *)aligned_alloc(4096, (size_t)rows * n_kept * 4);
    run(n_thr, rows, [&](int64_t a, int64_t b) { for (int64_t r = a; r < b; ++r) for (int64_t j = 0; j < ld; ++j) x[r * ld + j] = (float)(j & 7); });
    memset(dst, 0, (size_t)rows * n_kept * 4);
    std::vector<int32_t> cols(ld);
    for (int j = 0; j < ld; ++j) cols[j] = j;
    std::mt19937 rng(1);
    std::shuffle(cols.begin(), cols.end(), rng);
    cols.resize(n_kept);
    std::sort(cols.begin(), cols.end());
    const double gb = (double)rows * ld * 4 / 1e9;
    for (int rep = 0; rep < 2; ++rep) {
        volatile double sink = 0;
        double t = run(n_thr, rows, [&](int64_t a, int64_t b) { double s = 0; for (int64_t r = a; r < b; ++r) { const float *p = x + r * ld; for (int64_t j = 0; j < ld; j += 16) s += p[j]; } sink = sink + s; });
        printf("threads %d  stream-read (one load per line): %.1f ms  = %.1f GB/s of matrix\n", n_thr, t * 1e3, gb / t);
        t = run(n_thr, rows, [&](int64_t a, int64_t b) {
            for (int64_t r = a; r < b; ++r) { const float *s = x + r * ld; float *d = dst + r * n_kept; for (int j = 0; j < n_kept; ++j) d[j] = s[cols[j]]; } });
        printf("  plain gather:                 %.1f ms  (%.1f GB/s of matrix)\n", t * 1e3, gb / t);
        t = run(n_thr, rows, [&](int64_t a, int64_t b) {
            for (int64_t r = a; r < b; ++r) {
                const float *s = x + r * ld, *nx = s + ld; float *d = dst + r * n_kept;
                for (int j = 0; j < n_kept; ++j) { _mm_prefetch((const char *)(nx + cols[j]), _MM_HINT_T0); d[j] = s[cols[j]]; } } });
        printf("  prefetch next row:            %.1f ms  (%.1f GB/s of matrix)\n", t * 1e3, gb / t);
        t = run(n_thr, rows, [&](int64_t a, int64_t b) {
            for (int64_t r = a; r < b; ++r) {
                const float *s = x + r * ld, *nx = s + 2 * ld; float *d = dst + r * n_kept;
                for (int j = 0; j < n_kept; ++j) { _mm_prefetch((const char *)(nx + cols[j]), _MM_HINT_NTA); d[j] = s[cols[j]]; } } });
        printf("  prefetch 2 rows ahead (NTA):  %.1f ms  (%.1f GB/s of matrix)\n", t * 1e3, gb / t);
        t = run(n_thr, rows, [&](int64_t a, int64_t b) {
            int64_t r = a;
            for (; r + 4 <= b; r += 4) {
                const float *s0 = x + r * ld, *s1 = s0 + ld, *s2 = s1 + ld, *s3 = s2 + ld; float *d0 = dst + r * n_kept, *d1 = d0 + n_kept, *d2 = d1 + n_kept, *d3 = d2 + n_kept;
                for (int j = 0; j < n_kept; ++j) { const int c = cols[j]; d0[j] = s0[c]; d1[j] = s1[c]; d2[j] = s2[c]; d3[j] = s3[c]; } }
            for (; r < b; ++r) { const float *s = x + r * ld; float *d = dst + r * n_kept; for (int j = 0; j < n_kept; ++j) d[j] = s[cols[j]]; } });
        printf("  4 rows interleaved:           %.1f ms  (%.1f GB/s of matrix)\n", t * 1e3, gb / t);
        t = run(n_thr, rows, [&](int64_t a, int64_t b) {
            int64_t r = a;
            for (; r + 8 <= b; r += 8) {
                const float *s0 = x + r * ld; float *d0 = dst + r * n_kept;
                for (int j = 0; j < n_kept; ++j) { const int c = cols[j];
#pragma GCC unroll 8
                    for (int q = 0; q < 8; ++q) d0[q * n_kept + j] = s0[q * ld + c]; } }
            for (; r < b; ++r) { const float *s = x + r * ld; float *d = dst + r * n_kept; for (int j = 0; j < n_kept; ++j) d[j] = s[cols[j]]; } });
        printf("  8 rows interleaved:           %.1f ms  (%.1f GB/s of matrix)\n", t * 1e3, gb / t);
        t = run(n_thr, rows, [&](int64_t a, int64_t b) {
            int64_t r = a;
            for (; r + 4 <= b; r += 4) {
                const float *s0 = x + r * ld, *s1 = s0 + ld, *s2 = s1 + ld, *s3 = s2 + ld; float *d0 = dst + r * n_kept, *d1 = d0 + n_kept, *d2 = d1 + n_kept, *d3 = d2 + n_kept;
                const float *p0 = s0 + 4 * ld;
                for (int j = 0; j < n_kept; ++j) { const int c = cols[j];
                    _mm_prefetch((const char *)(p0 + c), _MM_HINT_T0); _mm_prefetch((const char *)(p0 + ld + c), _MM_HINT_T0);
                    _mm_prefetch((const char *)(p0 + 2 * ld + c), _MM_HINT_T0); _mm_prefetch((const char *)(p0 + 3 * ld + c), _MM_HINT_T0);
                    d0[j] = s0[c]; d1[j] = s1[c]; d2[j] = s2[c]; d3[j] = s3[c]; } }
            for (; r < b; ++r) { const float *s = x + r * ld; float *d = dst + r * n_kept; for (int j = 0; j < n_kept; ++j) d[j] = s[cols[j]]; } });
        printf("  4 rows + prefetch next 4:     %.1f ms  (%.1f GB/s of matrix)\n", t * 1e3, gb / t);
    }
    return 0;
}
