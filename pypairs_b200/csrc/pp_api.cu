// pp_api.cu -- context management and the small C-ABI entry points (include/pypairs_b200.h).
#include <ctype.h>
#include <sched.h>
#include <stdarg.h>

#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <thread>
#include <mutex>

#include <math.h>

#include "pp_common.cuh"

thread_local std::string g_pp_err;

int pp_fail(pp_ctx *ctx, int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_pp_err = buf;
    return code;
}

void *pp_slab(pp_ctx *ctx, int slot, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (ctx->slab_bytes[slot] >= bytes) return ctx->slab[slot];
    cudaStreamSynchronize(ctx->stream);  // nothing may still be using the old slab
    if (ctx->slab[slot]) cudaFree(ctx->slab[slot]);
    ctx->slab[slot] = nullptr;
    ctx->slab_bytes[slot] = 0;
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ctx->slab[slot], want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&ctx->slab[slot], bytes);
        if (e != cudaSuccess) {
            pp_fail(ctx, PP_ENOMEM, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
            return nullptr;
        }
        ctx->slab_bytes[slot] = bytes;
        return ctx->slab[slot];
    }
    ctx->slab_bytes[slot] = want;
    return ctx->slab[slot];
}

int pp_host_threads() {
    if (const char *e = getenv("PP_HOST_THREADS")) return std::max(1, atoi(e));
    unsigned hc = std::thread::hardware_concurrency();
    hc = std::min(hc ? hc : 8u, 32u);
    if (const char *e = getenv("LOCAL_WORLD_SIZE"))  // one process per GPU: share the host cores
        hc = std::max(1u, hc / (unsigned)std::max(1, atoi(e)));
    return (int)std::max(1u, hc);
}

int pp_copy_h2d(pp_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st) {
    constexpr size_t CHUNK = (size_t)32 << 20;
    if (bytes == 0) return PP_OK;
    bool pageable = false;
    if (bytes >= 2 * CHUNK) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, src) == cudaSuccess)
            pageable = at.type == cudaMemoryTypeUnregistered;
        else
            cudaGetLastError();
    }
    if (!pageable) {
        PP_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return PP_OK;
    }
    for (int b = 0; b < 2; ++b) {
        if (!ctx->pin_h2d[b]) PP_CUDA(cudaHostAlloc(&ctx->pin_h2d[b], CHUNK, cudaHostAllocDefault));
        if (!ctx->pin_h2d_done[b]) PP_CUDA(cudaEventCreateWithFlags(&ctx->pin_h2d_done[b], cudaEventDisableTiming));
    }
    const int n_thr = std::max(1, std::min(pp_host_threads(), 8));
    int i = 0;
    for (size_t off = 0; off < bytes; off += CHUNK, ++i) {
        const int b = i & 1;
        const size_t n = std::min(CHUNK, bytes - off);
        // the DMA that last read this buffer (two chunks ago, or in an earlier call) has drained it; an event that was never
        // recorded is complete
        PP_CUDA(cudaEventSynchronize(ctx->pin_h2d_done[b]));
        const uint8_t *s = (const uint8_t *)src + off;
        uint8_t *d = (uint8_t *)ctx->pin_h2d[b];
        std::vector<std::thread> pool;
        try {
            for (int t = 1; t < n_thr; ++t)
                pool.emplace_back([=] { memcpy(d + n * t / n_thr, s + n * t / n_thr, n * (t + 1) / n_thr - n * t / n_thr); });
        } catch (const std::exception &) {  // fewer threads than wished for: the caller's thread copies the rest
        }
        const size_t done_by_pool = n * (pool.size() + 1) / n_thr;  // parts 1 .. pool.size()
        memcpy(d, s, n / n_thr);
        if (pool.size() + 1 < (size_t)n_thr) memcpy(d + done_by_pool, s + done_by_pool, n - done_by_pool);
        for (auto &th : pool) th.join();
        PP_CUDA(cudaMemcpyAsync((uint8_t *)dst + off, d, n, cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaEventRecord(ctx->pin_h2d_done[b], st));
    }
    return PP_OK;
}

PP_API int pp_version(void) { return PP_ABI_VERSION; }

PP_API int pp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static void bind_to_gpu_numa_node(int device);

PP_API int pp_ctx_create(int device, pp_ctx **out) {
    if (!out) return pp_fail(nullptr, PP_EINVAL, "pp_ctx_create: out is NULL");
    *out = nullptr;
    pp_ctx *ctx = nullptr;  // PP_CUDA reports through g_pp_err while ctx is NULL
    int n = 0;
    PP_CUDA(cudaGetDeviceCount(&n));
    if (n == 0) return pp_fail(nullptr, PP_ECUDA, "pp_ctx_create: no CUDA device");
    if (device < 0) PP_CUDA(cudaGetDevice(&device));
    if (device >= n) return pp_fail(nullptr, PP_EINVAL, "pp_ctx_create: device %d of %d", device, n);
    cudaDeviceProp prop;
    PP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return pp_fail(nullptr, PP_ECUDA, "pp_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
    PP_CUDA(cudaSetDevice(device));
    bind_to_gpu_numa_node(device);
    pp_ctx *c = new pp_ctx();
    c->device = device;
    c->n_sm = prop.multiProcessorCount;
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&c->ev[i]);
    if (e == cudaSuccess) {
        // keep freed blocks in the pool: repeated calls re-use them instead of hitting the driver
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t thresh = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);
        }
    }
    if (e != cudaSuccess) {
        delete c;
        return pp_fail(nullptr, PP_ECUDA, "pp_ctx_create: %s", cudaGetErrorString(e));
    }
    *out = c;
    return PP_OK;
}

// One process per GPU on a multi-socket host: run this thread (and the threads it starts: the gather crew, the staging
// copies) on the cores of the GPU's own NUMA node, so that host buffers allocated from now on are local to the GPU's
// PCIe root.  Only under a multi-rank launcher (LOCAL_WORLD_SIZE > 1) or with PP_NUMA_BIND=1; PP_NUMA_BIND=0 switches it
// off.  Best effort: any failure leaves the affinity as it was.
static void bind_to_gpu_numa_node(int device) {
    const char *env = getenv("PP_NUMA_BIND");
    const char *lws = getenv("LOCAL_WORLD_SIZE");
    const bool want = env ? atoi(env) != 0 : (lws && atoi(lws) > 1);
    if (!want) return;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return;
    int node = -1;
    const int got = fscanf(f, "%d", &node);
    fclose(f);
    if (got != 1 || node < 0) return;
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return;
    char list[4096] = {0};
    const bool ok = fgets(list, sizeof(list), f) != nullptr;
    fclose(f);
    if (!ok) return;
    cpu_set_t now, set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(now), &now) != 0) return;
    int n_set = 0;
    char *save = nullptr;
    for (char *tok = strtok_r(list, ",\n", &save); tok; tok = strtok_r(nullptr, ",\n", &save)) {  // "0-15,32-47"
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &now)) {  // never widen what the launcher allowed
                CPU_SET(c, &set);
                ++n_set;
            }
    }
    if (n_set > 0) sched_setaffinity(0, sizeof(set), &set);
}

PP_API void pp_ctx_destroy(pp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) pp_comm_destroy(ctx);
    for (int i = 0; i < 8; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int b = 0; b < 2; ++b)
        if (ctx->pin[b]) cudaFreeHost(ctx->pin[b]);
    if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
    for (int b = 0; b < 2; ++b) {
        if (ctx->pin_h2d[b]) cudaFreeHost(ctx->pin_h2d[b]);
        if (ctx->pin_h2d_done[b]) cudaEventDestroy(ctx->pin_h2d_done[b]);
    }
    for (int i = 0; i < pp_ctx::N_SLABS; ++i)
        if (ctx->slab[i]) cudaFree(ctx->slab[i]);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->zc_stream) cudaStreamDestroy(ctx->zc_stream);
    delete ctx;
}

PP_API const char *pp_last_error(pp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_pp_err.c_str(); }

// ---- result buffers handed to the caller --------------------------------------------------------------------------
// Large variable-size results (the marker lists of pp_sandbag) are returned in PINNED host blocks from a small
// process-wide pool: the device -> host copy lands directly in the buffer the caller receives (no staging copy, no
// page faults of a fresh malloc), and pp_free() puts the block back.  Small results and failures fall back to malloc.
namespace {
struct ResultBlock {
    void *p;
    size_t bytes;
    bool used;
    uint64_t last_use;
};
std::vector<ResultBlock> g_blocks;
std::mutex g_blocks_mutex;
uint64_t g_blocks_clock = 0;
constexpr size_t RESULT_POOL_MAX = (size_t)2 << 30, RESULT_SMALL = 64 << 10;
constexpr size_t RESULT_MAX_BLOCKS = 16;
}  // namespace

void *pp_result_alloc(size_t bytes, bool *pinned) {
    if (pinned) *pinned = false;
    if (bytes == 0) bytes = 1;
    if (bytes >= RESULT_SMALL && bytes <= RESULT_POOL_MAX / 2) {
        std::lock_guard<std::mutex> lock(g_blocks_mutex);
        ResultBlock *best = nullptr;
        for (auto &b : g_blocks)  // best fit among the free blocks that are not absurdly large for this result
            if (!b.used && b.bytes >= bytes && b.bytes <= 4 * bytes && (!best || b.bytes < best->bytes)) best = &b;
        if (!best) {
            size_t total = 0;
            for (auto &b : g_blocks) total += b.bytes;
            const size_t want = (bytes + bytes / 4 + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
            // make room: drop the least recently used free blocks
            while (!g_blocks.empty() && (total + want > RESULT_POOL_MAX || g_blocks.size() >= RESULT_MAX_BLOCKS)) {
                int lru = -1;
                for (int i = 0; i < (int)g_blocks.size(); ++i)
                    if (!g_blocks[i].used && (lru < 0 || g_blocks[i].last_use < g_blocks[lru].last_use)) lru = i;
                if (lru < 0) break;
                cudaFreeHost(g_blocks[lru].p);
                total -= g_blocks[lru].bytes;
                g_blocks.erase(g_blocks.begin() + lru);
            }
            void *p = nullptr;
            if (total + want <= RESULT_POOL_MAX && g_blocks.size() < RESULT_MAX_BLOCKS &&
                cudaHostAlloc(&p, want, cudaHostAllocPortable) == cudaSuccess) {
                g_blocks.push_back(ResultBlock{p, want, false, 0});
                best = &g_blocks.back();
            } else {
                cudaGetLastError();
            }
        }
        if (best) {
            best->used = true;
            best->last_use = ++g_blocks_clock;
            if (pinned) *pinned = true;
            return best->p;
        }
    }
    return malloc(bytes);
}

PP_API void *pp_alloc_result(size_t bytes) { return pp_result_alloc(bytes, nullptr); }

bool pp_is_pinned_host(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

PP_API void pp_free(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lock(g_blocks_mutex);
        for (auto &b : g_blocks)
            if (b.p == p) {
                b.used = false;
                return;
            }
    }
    free(p);
}

PP_API int pp_get_timings(pp_ctx *ctx, pp_timings *out) {
    if (!ctx || !out) return pp_fail(ctx, PP_EINVAL, "pp_get_timings: NULL argument");
    *out = ctx->tm;
    return PP_OK;
}

PP_API int pp_dense_rank(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                         int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int32_t *gene_cols,
                         int64_t n_genes, uint16_t *out_rank, int32_t *out_max_rank) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    if (!x || !cell_rows || !gene_cols || !out_rank) return pp_fail(ctx, PP_EINVAL, "pp_dense_rank: NULL argument");
    if (x_dtype != PP_F32 && x_dtype != PP_F64) return pp_fail(ctx, PP_EINVAL, "pp_dense_rank: bad x_dtype");
    if (n_rows <= 0 || n_cols <= 0 || ld < n_cols || n_cells < 0 || n_genes < 0)
        return pp_fail(ctx, PP_EINVAL, "pp_dense_rank: bad shape");
    for (int64_t i = 0; i < n_cells; ++i)
        if (cell_rows[i] >= n_rows) return pp_fail(ctx, PP_EINVAL, "pp_dense_rank: cell_rows out of range");
    for (int64_t g = 0; g < n_genes; ++g)
        if (gene_cols[g] < 0 || gene_cols[g] >= n_cols) return pp_fail(ctx, PP_EINVAL, "pp_dense_rank: gene_cols out of range");
    if (out_max_rank) *out_max_rank = 0;
    if (n_cells == 0 || n_genes == 0) return PP_OK;
    PP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->launches = 0;
    const size_t esz = x_dtype == PP_F64 ? 8 : 4;
    DevBuf d_x, d_rows, d_cols, d_rank, d_flags;
    const void *dx = x;
    if (!x_is_device) {
        PP_CUDA(d_x.alloc((size_t)n_rows * ld * esz, st));
        PP_CUDA(cudaMemcpyAsync(d_x.p, x, (size_t)n_rows * ld * esz, cudaMemcpyHostToDevice, st));
        dx = d_x.p;
    }
    PP_CUDA(d_rows.alloc(sizeof(int32_t) * n_cells, st));
    PP_CUDA(cudaMemcpyAsync(d_rows.p, cell_rows, sizeof(int32_t) * n_cells, cudaMemcpyHostToDevice, st));
    PP_CUDA(d_cols.alloc(sizeof(int32_t) * n_genes, st));
    PP_CUDA(cudaMemcpyAsync(d_cols.p, gene_cols, sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice, st));
    PP_CUDA(d_rank.alloc(sizeof(uint16_t) * n_cells * n_genes, st));
    PP_CUDA(d_flags.alloc(sizeof(int32_t) * 4, st));
    PP_CUDA(cudaMemsetAsync(d_flags.p, 0, sizeof(int32_t) * 4, st));
    int rc = pp_rank_launch(ctx, dx, x_dtype, ld, d_rows.as<int32_t>(), n_cells, d_cols.as<int32_t>(), n_genes,
                            d_rank.as<uint16_t>(), n_genes, d_flags.as<int32_t>(), gene_cols);
    if (rc) return rc;
    int32_t flags[4];
    PP_CUDA(cudaMemcpyAsync(out_rank, d_rank.p, sizeof(uint16_t) * n_cells * n_genes, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    if (flags[1]) return pp_fail(ctx, PP_ENAN, "pp_dense_rank: NaN in the expression matrix");
    if (out_max_rank) *out_max_rank = flags[0];
    return PP_OK;
}

// ---- label columns of the result frame (host only) -----------------------------------------------
PP_API int pp_score_labels(const double *scores, int32_t n_classes, int64_t n_cells, int32_t idx_g1, int32_t idx_g2m,
                           int32_t *out_max_class, int32_t *out_cc) {
    if (!scores || !out_max_class || n_classes < 0 || n_cells < 0) return pp_fail(nullptr, PP_EINVAL, "pp_score_labels: bad argument");
    const bool cc = idx_g1 >= 0 && idx_g2m >= 0 && out_cc;
    if (cc && (idx_g1 >= n_classes || idx_g2m >= n_classes)) return pp_fail(nullptr, PP_EINVAL, "pp_score_labels: class index out of range");
    for (int64_t i = 0; i < n_cells; ++i) {
        int arg = 0;
        double best = -INFINITY, sum = 0.0;
        for (int k = 0; k < n_classes; ++k) {
            const double v = scores[(int64_t)k * n_cells + i];
            if (v != v) continue;  // NaN is skipped by idxmax and by sum
            sum += v;
            if (v > best) {
                best = v;
                arg = k;
            }
        }
        out_max_class[i] = sum > 0.0 ? arg : n_classes;
        if (cc) {
            const double a = scores[(int64_t)idx_g1 * n_cells + i], b = scores[(int64_t)idx_g2m * n_cells + i];
            const double sa = a != a ? -INFINITY : a, sb = b != b ? -INFINITY : b;
            const double mx = sa > sb ? sa : sb;
            out_cc[i] = (mx == -INFINITY && (a != a) && (b != b)) ? -1 : mx < 0.5 ? 2 : (sb > sa ? 1 : 0);
        }
    }
    return PP_OK;
}

// ---- integer-ALU peak micro-benchmark -----------------------------------------------------------
namespace {
constexpr int PEAK_ITERS = 4096, PEAK_CHAINS = 16;
__global__ void __launch_bounds__(256) alu_peak_kernel(uint32_t *out, uint32_t seed) {
    uint32_t v[PEAK_CHAINS];
    const uint32_t a = seed ^ (threadIdx.x * 2654435761u), b = ~seed + blockIdx.x;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) v[i] = a + i;
    for (int it = 0; it < PEAK_ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) v[i] = lop3_gt(a, v[(i + 1) % PEAK_CHAINS], v[i]) ^ 0u;
#pragma unroll
        for (int i = 0; i < PEAK_CHAINS; ++i) v[i] = lop3_lt(v[i], b, v[(i + 5) % PEAK_CHAINS]);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < PEAK_CHAINS; ++i) r ^= v[i];
    if (r == 0x12345u) out[0] = r;  // keeps the chains alive
}
}  // namespace

PP_API int pp_alu_peak(pp_ctx *ctx, int32_t repeats, double *out_lop3_lane_ops_per_s, int32_t *out_n_sm) {
    if (!ctx || !out_lop3_lane_ops_per_s) return pp_fail(ctx, PP_EINVAL, "pp_alu_peak: NULL argument");
    PP_CUDA(cudaSetDevice(ctx->device));
    DevBuf d;
    PP_CUDA(d.alloc(16, ctx->stream));
    const int blocks = ctx->n_sm * 8;
    double best = 0;
    for (int r = 0; r < repeats + 1; ++r) {
        PP_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
        alu_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(d.as<uint32_t>(), 0x9e3779b9u + r);
        PP_CHECK_LAUNCH();
        PP_CUDA(cudaEventRecord(ctx->ev[7], ctx->stream));
        PP_CUDA(cudaEventSynchronize(ctx->ev[7]));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
        const double ops = (double)blocks * 256 * (double)PEAK_ITERS * PEAK_CHAINS * 2;
        if (r > 0 && ms > 0) best = std::max(best, ops / (ms * 1e-3));
    }
    *out_lop3_lane_ops_per_s = best;
    if (out_n_sm) *out_n_sm = ctx->n_sm;
    return PP_OK;
}
