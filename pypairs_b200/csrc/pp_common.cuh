// pp_common.cuh -- shared declarations of the pypairs_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pypairs_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pypairs_b200 is written for sm_100a (B200) only"
#endif

#define PP_API extern "C" __attribute__((visibility("default")))

// The handful of NCCL declarations the library needs (values fixed by NCCL's ABI since 2.0); the entry points are
// resolved at run time (pp_comm.cu), so neither nccl.h nor libnccl is needed to build.
namespace ppn {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };
}  // namespace ppn

struct pp_ctx {
    int device = 0;
    int n_sm = 0;
    int smem_optin = 0;
    cudaStream_t stream = nullptr;   // compute stream
    cudaStream_t copy_stream = nullptr;  // H2D prefetch stream
    cudaStream_t zc_stream = nullptr;    // zero-copy gather lane of pp_cyclone (created on first use)
    cudaEvent_t ev[8] = {};
    void *pin[2] = {nullptr, nullptr};  // pinned staging buffers of the host-gather path (grown on demand)
    size_t pin_bytes = 0;
    void *pin_h2d[2] = {nullptr, nullptr};  // pinned staging ring of pp_copy_h2d (pageable sources)
    cudaEvent_t pin_h2d_done[2] = {nullptr, nullptr};
    void *pin_out = nullptr;  // pinned staging buffer for results copied device -> host (grown on demand)
    size_t pin_out_bytes = 0;
    // grow-only device slabs for the few large buffers of a call (matrix copy, rank matrix, bit planes, ...): calls
    // are synchronous, so the same slab is safely re-used by the next call and the allocator is out of the timed path
    static constexpr int N_SLABS = 12;
    void *slab[N_SLABS] = {};
    size_t slab_bytes[N_SLABS] = {};
    std::string err;
    pp_timings tm = {};
    int launches = 0;
    void *comm = nullptr;  // ncclComm_t of pp_comm_init (one rank per ctx), NULL = single GPU
    int comm_rank = 0, comm_world = 0;
};

extern thread_local std::string g_pp_err;  // errors raised without a ctx

int pp_fail(pp_ctx *ctx, int code, const char *fmt, ...);
// pointer to at least `bytes` of device memory in slab `slot` (contents undefined); nullptr + error on failure
void *pp_slab(pp_ctx *ctx, int slot, size_t bytes);
// host threads this process may use for staging work (PP_HOST_THREADS, else hardware threads / ranks on the host)
int pp_host_threads();
// host -> device copy on `st`.  Pinned / registered sources go straight to the DMA engine; a large PAGEABLE source is
// copied by a few threads into a ring of pinned staging buffers, chunk by chunk, each chunk's DMA overlapping the
// next chunk's memcpy (a plain cudaMemcpyAsync from pageable memory is staged by the driver at ~11 GB/s).
int pp_copy_h2d(pp_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t st);
// host buffer for a variable-size result the caller releases with pp_free(): pinned (pool) when large, else malloc
void *pp_result_alloc(size_t bytes, bool *pinned);
bool pp_is_pinned_host(const void *p);  // pinned / registered host memory (a DMA target without staging)
enum { SLAB_X = 0, SLAB_RANK = 1, SLAB_PLANES = 2, SLAB_OUT = 3, SLAB_OUT2 = 4, SLAB_CHUNK0 = 5, SLAB_CHUNK1 = 6, SLAB_MISC = 7, SLAB_SORT = 8, SLAB_ZC0 = 9, SLAB_ZC1 = 10 };

#define PP_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return pp_fail(ctx, PP_ECUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,     \
                           cudaGetErrorString(e__));                                               \
    } while (0)

#define PP_CHECK_LAUNCH()                                                                          \
    do {                                                                                           \
        ctx->launches++;                                                                           \
        PP_CUDA(cudaGetLastError());                                                               \
    } while (0)

// No C++ exception may cross the C ABI: host-side allocations (std::vector, std::thread) are guarded at the entry points.
#define PP_GUARD(ctx, call)                                                                         \
    do {                                                                                            \
        try {                                                                                       \
            return (call);                                                                          \
        } catch (const std::bad_alloc &) {                                                          \
            return pp_fail(ctx, PP_ENOMEM, "%s: host allocation failed", __func__);                 \
        } catch (const std::exception &e) {                                                         \
            return pp_fail(ctx, PP_ECUDA, "%s: internal error: %s", __func__, e.what());            \
        }                                                                                           \
    } while (0)

// RAII device buffer on the ctx stream (stream-ordered allocator keeps repeated calls cheap).
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaStream_t s = nullptr;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    cudaError_t alloc(size_t n, cudaStream_t stream) {
        release();
        s = stream;
        bytes = n;
        if (n == 0) n = 16;
        return cudaMallocAsync(&p, n, stream);
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
    }
    ~DevBuf() { release(); }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

// ---- device helpers ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Lanes of the warp that hold the same 8-bit digit as this lane (valid lanes only; an invalid lane gets 0).
// `match.any` does this in one instruction, but the hardware iterates once per DISTINCT value in the warp: with random
// digits that is ~30 rounds, ~600 cycles per call (ncu on the radix passes: half of all stall samples sat on its
// consumer).  Eight ballots, one per digit bit, are independent and pipeline: ~10x faster.
__device__ __forceinline__ uint32_t warp_digit_peers(uint32_t d, bool valid) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
    }
    return valid ? peers : 0u;
}
__device__ __forceinline__ uint32_t lop3_gt(uint32_t a, uint32_t b, uint32_t g) {
    // (a & ~b) | (~(a ^ b) & g) : "a > b" carried from the less significant planes
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xB2;" : "=r"(r) : "r"(a), "r"(b), "r"(g));
    return r;
}
__device__ __forceinline__ uint32_t lop3_lt(uint32_t a, uint32_t b, uint32_t l) {
    // (~a & b) | (~(a ^ b) & l)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x8E;" : "=r"(r) : "r"(a), "r"(b), "r"(l));
    return r;
}
#endif

// Device-resident CSR rows [r0, r0 + nr) of a pp_csr: indptr rebased to 0 as int64, indices / data of the slice.
struct DevCsr {
    DevBuf b_indptr, b_indices, b_data;
    const int64_t *indptr = nullptr;
    const int32_t *indices = nullptr;
    const void *data = nullptr;
    int64_t n_rows = 0, nnz = 0;
    int dtype = PP_F32;
};
int pp_csr_validate(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, const char *who);
int pp_csr_to_device(pp_ctx *ctx, const pp_csr *x, int64_t r0, int64_t nr, DevCsr *out);
int pp_csr_col_nonzero(pp_ctx *ctx, const DevCsr &m, uint8_t *d_nz);
// out[i, colmap[c]] = value for every stored entry (rows[i], c) with colmap[c] >= 0; out must be zero-filled;
// rows[i] < 0 leaves row i zero.
int pp_csr_densify(pp_ctx *ctx, const DevCsr &m, const int32_t *d_rows, int64_t n_out, const int32_t *d_colmap,
                   void *d_out, int64_t out_ld);

// ---- collectives on the ctx stream (pp_comm.cu); a ctx without communicator fails with PP_ENCCL
int pp_allreduce_max(pp_ctx *ctx, void *d_buf, size_t count, int nccl_dtype);  // in place
int pp_allgather_bytes(pp_ctx *ctx, const void *d_send, void *d_recv, size_t bytes_per_rank);  // in place allowed

// ---- internal entry points (one per .cu) --------------------------------------------------------
// Rank transform: for each slot s < n_slots (cell_rows[s] < 0 marks a padding slot: all ranks 0),
// rank[s * rank_ld + g] = dense rank of x[cell_rows[s], gene_cols[g]] among that cell's kept genes.
int pp_rank_launch(pp_ctx *ctx, const void *d_x, int x_dtype, int64_t ld, const int32_t *d_cell_rows,
                   int64_t n_slots, const int32_t *d_gene_cols, int64_t n_genes, uint16_t *d_rank,
                   int64_t rank_ld, int32_t *d_flags /* [0]=max rank (atomicMax), [1]=NaN seen */,
                   const int32_t *h_gene_cols = nullptr /* host copy of gene_cols: lets the launcher pick the streaming
                                                           kernel when the columns are ascending and dense */);
