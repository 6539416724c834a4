// pp_rank.cu -- per-cell dense-rank transform (prep stage of both hot paths).
//
// Both reference kernels only ever compare two genes *within one cell*
// (sandbag.py:179-182 `raw_data[g1,i] > raw_data[g2,i]`; cyclone.py:212-215
// `a != b`, `a > b`).  Replacing each cell's values by their dense rank among
// that cell's kept genes (ties share a rank, +0.0 == -0.0 as in IEEE `==`) keeps
// every such comparison bit-exact while shrinking an element from a float64 to
// <= 16 bits -- which is what lets the sandbag kernel run bit-sliced and the
// cyclone kernel keep a cell tile in shared memory.
//
// One CTA per cell, persistent over cells.  Per cell:
//   1. gather x[row, gene_cols[g]] -> order-preserving unsigned keys (+ u16 gene index)
//   2. LSD radix sort, 8-bit digits, only over the key bytes that actually vary in this
//      cell (count data stored as doubles varies in 2-3 bytes; float32 input in <= 4)
//   3. head flags + scan -> dense rank, scattered back to gene order as u16
// Keys live in a per-CTA global scratch (L2 resident: 32768 genes * 10 B * 2 buffers).
// Count fast path: when every kept value of the cell is an integer in [0, 65535] (raw UMI / read counts, the
// usual sandbag / cyclone input) no sort is needed: a 65 536-bit presence bitmap in shared memory, a prefix
// popcount over its 2048 words, and rank(v) = #present values below v.  One pass over the row from global
// memory; the counts are staged as u16 in shared memory for the look-up pass.
#include <stdarg.h>

#include <algorithm>

#include "pp_common.cuh"

namespace {

constexpr int RANK_THREADS = 512;
constexpr int RANK_WARPS = RANK_THREADS / 32;

__device__ __forceinline__ uint32_t key_of(float v, bool &nan) {
    v = v + 0.0f;  // -0.0 -> +0.0 so that the two zeros tie (IEEE ==)
    nan = nan || (v != v);
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint64_t key_of(double v, bool &nan) {
    v = v + 0.0;
    nan = nan || (v != v);
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

template <typename T> struct KeyOf;
template <> struct KeyOf<float> { typedef uint32_t type; };
template <> struct KeyOf<double> { typedef uint64_t type; };

constexpr int BM_WORDS = 2048;  // 65 536-bit presence bitmap of the count fast path

__device__ __forceinline__ bool is_small_count(float v) { return v >= 0.0f && v < 65536.0f && v == truncf(v); }
__device__ __forceinline__ bool is_small_count(double v) { return v >= 0.0 && v < 65536.0 && v == trunc(v); }

// "few distinct values" path: a cell whose kept genes take at most HT_MAX distinct values (normalised counts, log1p,
// ...) is ranked by sorting only those values: shared-memory hash set -> compaction -> bitonic sort -> binary search
constexpr int HT_SLOTS = 2048, HT_MAX = 1024;
__device__ __forceinline__ uint32_t hash_slot(uint32_t k) { return (k * 0x9E3779B1u) >> 21; }
__device__ __forceinline__ uint32_t hash_slot(uint64_t k) { return (uint32_t)((k * 0x9E3779B97F4A7C15ull) >> 53); }
__device__ __forceinline__ uint32_t smem_cas(uint32_t *p, uint32_t cmp, uint32_t v) { return atomicCAS(p, cmp, v); }
__device__ __forceinline__ uint64_t smem_cas(uint64_t *p, uint64_t cmp, uint64_t v) {
    return (uint64_t)atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)v);
}

struct alignas(16) RankSmem {
    uint32_t bitmap[BM_WORDS];
    uint32_t bm_prefix[BM_WORDS];
    uint32_t whist[RANK_WARPS][256];  // per-warp digit histogram -> running scatter offsets
    uint32_t dig_total[256];
    uint32_t wtot[RANK_WARPS];
    unsigned long long vary[RANK_WARPS];
    unsigned long long vary_all;
    int n_distinct;
};
static_assert(sizeof(uint64_t) * HT_SLOTS <= sizeof(uint32_t) * 2 * BM_WORDS, "hash set overlays bitmap + bm_prefix");
static_assert(sizeof(uint64_t) * HT_MAX <= sizeof(uint32_t) * RANK_WARPS * 256, "sorted values overlay whist");

template <typename T>
__global__ void __launch_bounds__(RANK_THREADS)
rank_kernel(const T *__restrict__ x, int64_t ld, const int32_t *__restrict__ cell_rows, int64_t n_slots,
            const int32_t *__restrict__ gene_cols, int n, uint16_t *__restrict__ rank_out, int64_t rank_ld,
            uint8_t *__restrict__ scratch, size_t scratch_per_cta, int32_t *__restrict__ flags,
            const int32_t *__restrict__ todo, const int32_t *__restrict__ ctl, int32_t *__restrict__ todo2,
            int32_t *__restrict__ ctl2) {
    typedef typename KeyOf<T>::type K;
    __shared__ RankSmem sm;
    extern __shared__ __align__(16) uint16_t vals[];  // [n] staged counts of the fast path
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp-contiguous element ranges (multiples of 32 so every step is a full, aligned warp step)
    const int per_warp = ((n + RANK_WARPS * 32 - 1) / (RANK_WARPS * 32)) * 32;
    const int lo = min(warp * per_warp, n), hi = min(lo + per_warp, n);

    uint8_t *my = scratch + (size_t)blockIdx.x * scratch_per_cta;
    const size_t n_al = ((size_t)n + 15) & ~(size_t)15;
    K *keyA = reinterpret_cast<K *>(my);
    K *keyB = keyA + n_al;
    uint16_t *idxA = reinterpret_cast<uint16_t *>(keyB + n_al);
    uint16_t *idxB = idxA + n_al;

    // ctl == NULL or ctl[0] == 0: every slot; else only the ctl[1] slots listed in `todo` (the cells the streaming
    // kernel below could not rank on its count fast path)
    const bool listed = ctl != nullptr && ctl[0] != 0;
    const int64_t n_work = listed ? (int64_t)ctl[1] : n_slots;
    for (int64_t wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const int64_t slot = listed ? (int64_t)todo[wi] : wi;
        uint16_t *out = rank_out + slot * rank_ld;
        const int row = cell_rows[slot];
        if (row < 0) {  // padding cell: ties in every gene
            for (int g = tid; g < n; g += RANK_THREADS) out[g] = 0;
            continue;
        }
        const T *xr = x + (int64_t)row * ld;
        // ---- 0. count fast path: one pass over the row from global memory (8 independent loads in flight per
        //         thread), values staged as u16 in shared memory for the rank look-up pass
        {
            for (int i = tid; i < BM_WORDS; i += RANK_THREADS) sm.bitmap[i] = 0;
            __syncthreads();
            bool small = true;
            uint32_t m0 = 0, m1 = 0;  // presence bits of the values 0..63 seen by this thread (see rank_stream_kernel)
            for (int g0 = tid; g0 < n; g0 += RANK_THREADS * 8) {
                T v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int g = g0 + j * RANK_THREADS;
                    v[j] = g < n ? xr[gene_cols[g]] : (T)0;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int g = g0 + j * RANK_THREADS;
                    if (g < n) {
                        const bool ok = is_small_count(v[j]);
                        small = small && ok;
                        const uint32_t u = ok ? (uint32_t)v[j] : 0u;
                        vals[g] = (uint16_t)u;
                        const uint32_t bit = 1u << (u & 31);
                        // most cells hold a handful of distinct small counts: they are collected in registers; for the
                        // rest test first, the atomic is then rare
                        if (u < 32) m0 |= bit;
                        else if (u < 64) m1 |= bit;
                        else if (!(*(volatile uint32_t *)&sm.bitmap[u >> 5] & bit)) atomicOr(&sm.bitmap[u >> 5], bit);
                    }
                }
            }
            m0 = __reduce_or_sync(0xffffffffu, m0);
            m1 = __reduce_or_sync(0xffffffffu, m1);
            if (lane == 0) {
                if (m0) atomicOr(&sm.bitmap[0], m0);
                if (m1) atomicOr(&sm.bitmap[1], m1);
            }
            if (__syncthreads_and(small)) {  // CTA-uniform (NaN fails the test and takes the sort path)
                // exclusive prefix popcount over the 2048 words: 4 words per thread, warp scan, warp totals
                uint32_t c[4], t4 = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) { c[j] = __popc(sm.bitmap[tid * 4 + j]); t4 += c[j]; }
                uint32_t incl = t4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) sm.wtot[warp] = incl;
                __syncthreads();
                uint32_t base = 0;
                for (int w = 0; w < warp; ++w) base += sm.wtot[w];
                uint32_t run = base + incl - t4;
#pragma unroll
                for (int j = 0; j < 4; ++j) { sm.bm_prefix[tid * 4 + j] = run; run += c[j]; }
                if (tid == RANK_THREADS - 1) atomicMax(&flags[0], (int)run - 1);  // distinct values - 1
                __syncthreads();
                const uint32_t b0 = sm.bitmap[0], b1 = sm.bitmap[1], p1 = __popc(b0);
                for (int g = tid; g < n; g += RANK_THREADS) {
                    const uint32_t u = vals[g], below = (1u << (u & 31)) - 1u;
                    out[g] = (uint16_t)(u < 32 ? __popc(b0 & below) : u < 64 ? p1 + __popc(b1 & below)
                                                                            : sm.bm_prefix[u >> 5] + __popc(sm.bitmap[u >> 5] & below));
                }
                __syncthreads();  // bitmap / prefix / vals reused by the next cell
                continue;
            }
        }
        // ---- 0b. few distinct values (any real numbers): hash set of the keys, sort the distinct ones, binary search
        {
            K *ht = reinterpret_cast<K *>(&sm.bitmap[0]);        // HT_SLOTS keys, ~K(0) = empty (a NaN pattern)
            K *sorted = reinterpret_cast<K *>(&sm.whist[0][0]);  // HT_MAX keys
            const K EMPTY = ~(K)0;
            for (int i = tid; i < HT_SLOTS; i += RANK_THREADS) ht[i] = EMPTY;
            if (tid == 0) sm.n_distinct = 0;
            __syncthreads();
            bool bad = false;  // NaN, or too many distinct values
            for (int g0 = tid; g0 < n && !bad; g0 += RANK_THREADS * 4) {
                T v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int g = g0 + j * RANK_THREADS;
                    v[j] = g < n ? xr[gene_cols[g]] : (T)0;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (g0 + j * RANK_THREADS >= n || bad) continue;
                    const K k = key_of(v[j], bad);
                    if (bad) break;
                    uint32_t slot = hash_slot(k);
                    for (;;) {
                        const K cur = *(volatile K *)&ht[slot];
                        if (cur == k) break;
                        if (cur == EMPTY) {
                            const K old = smem_cas(&ht[slot], EMPTY, k);
                            if (old == EMPTY) {
                                if (atomicAdd(&sm.n_distinct, 1) >= HT_MAX) bad = true;
                                break;
                            }
                            if (old == k) break;
                        }
                        if (*(volatile int *)&sm.n_distinct > HT_MAX) {  // the set is being abandoned: stop probing
                            bad = true;
                            break;
                        }
                        slot = (slot + 1) & (HT_SLOTS - 1);
                    }
                }
            }
            if (!__syncthreads_or(bad)) {  // CTA-uniform
                const int D = sm.n_distinct;
                int P2 = 1;
                while (P2 < D) P2 <<= 1;
                // compaction of the non-empty slots: 4 slots per thread, warp scan + warp totals
                K mine[4];
                uint32_t cnt = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    mine[j] = ht[tid * 4 + j];
                    cnt += mine[j] != EMPTY;
                }
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) sm.wtot[warp] = incl;
                for (int i = D + tid; i < P2; i += RANK_THREADS) sorted[i] = EMPTY;  // padding sorts to the end
                __syncthreads();
                uint32_t pos = incl - cnt;
                for (int w = 0; w < warp; ++w) pos += sm.wtot[w];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (mine[j] != EMPTY) sorted[pos++] = mine[j];
                __syncthreads();
                // bitonic sort of P2 <= 1024 keys: at most one compare-exchange per thread and step
                for (int size = 2; size <= P2; size <<= 1)
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        if (tid < (P2 >> 1)) {
                            const int i = 2 * tid - (tid & (stride - 1));  // lower index of this thread's pair
                            const bool up = ((i & size) == 0);
                            const K a = sorted[i], b = sorted[i + stride];
                            if ((a > b) == up) {
                                sorted[i] = b;
                                sorted[i + stride] = a;
                            }
                        }
                        __syncthreads();
                    }
                if (tid == 0) atomicMax(&flags[0], D - 1);
                // dense rank = position of the value among the sorted distinct values (row re-read: L1/L2 hits)
                for (int g = tid; g < n; g += RANK_THREADS) {
                    bool dummy = false;
                    const K k = key_of(xr[gene_cols[g]], dummy);
                    int lo_i = 0, hi_i = D - 1;
                    while (lo_i < hi_i) {
                        const int mid = (lo_i + hi_i) >> 1;
                        if (sorted[mid] < k) lo_i = mid + 1; else hi_i = mid;
                    }
                    out[g] = (uint16_t)lo_i;
                }
                __syncthreads();  // shared memory reused by the next cell
                continue;
            }
        }
        // ---- cells that need a full sort go to rank_sort_kernel (keys sorted in shared memory) when the caller set it up
        if (todo2 != nullptr) {
            if (tid == 0) todo2[atomicAdd(&ctl2[1], 1)] = (int32_t)slot;
            continue;
        }
        // ---- 1. gather + key transform, find the key bits that vary
        bool nan = false;
        K first = key_of(xr[gene_cols[0]], nan);
        K vary = 0;
        for (int g = tid; g < n; g += RANK_THREADS) {
            K k = key_of(xr[gene_cols[g]], nan);
            keyA[g] = k;
            idxA[g] = (uint16_t)g;
            vary |= (k ^ first);
        }
        if (nan) flags[1] = 1;
        unsigned long long v64 = (unsigned long long)vary;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v64 |= __shfl_xor_sync(0xffffffffu, v64, o);
        if (lane == 0) sm.vary[warp] = v64;
        __syncthreads();
        if (tid == 0) {
            unsigned long long a = 0;
            for (int w = 0; w < RANK_WARPS; ++w) a |= sm.vary[w];
            sm.vary_all = a;
        }
        __syncthreads();
        const unsigned long long vary_all = sm.vary_all;

        K *kin = keyA, *kout = keyB;
        uint16_t *iin = idxA, *iout = idxB;
        // ---- 2. LSD radix passes over varying bytes only
        for (int byte = 0; byte < (int)sizeof(K); ++byte) {
            if (((vary_all >> (8 * byte)) & 0xffull) == 0) continue;
            const int shift = 8 * byte;
            for (int i = tid; i < RANK_WARPS * 256; i += RANK_THREADS) (&sm.whist[0][0])[i] = 0;
            __syncthreads();
            for (int ib = lo; ib < hi; ib += 32) {
                const int i = ib + lane;
                const bool valid = i < hi;
                uint32_t d = valid ? ((uint32_t)(kin[i] >> shift) & 255u) : (256u + lane);
                uint32_t peers = warp_digit_peers(d, valid);
                if (valid && lane == __ffs(peers) - 1) sm.whist[warp][d] += __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            // exclusive scan: digit-major, warp-minor
            if (tid < 256) {
                uint32_t s = 0;
                for (int w = 0; w < RANK_WARPS; ++w) s += sm.whist[w][tid];
                sm.dig_total[tid] = s;
            }
            __syncthreads();
            if (warp == 0) {  // scan 256 totals with one warp, 8 per lane
                uint32_t v[8], s = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = sm.dig_total[lane * 8 + j]; s += v[j]; }
                uint32_t incl = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                uint32_t run = incl - s;
#pragma unroll
                for (int j = 0; j < 8; ++j) { sm.dig_total[lane * 8 + j] = run; run += v[j]; }
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t run = sm.dig_total[tid];
                for (int w = 0; w < RANK_WARPS; ++w) {
                    uint32_t c = sm.whist[w][tid];
                    sm.whist[w][tid] = run;
                    run += c;
                }
            }
            __syncthreads();
            // stable scatter
            for (int ib = lo; ib < hi; ib += 32) {
                const int i = ib + lane;
                const bool valid = i < hi;
                K k = valid ? kin[i] : (K)0;
                uint16_t ix = valid ? iin[i] : (uint16_t)0;
                uint32_t d = valid ? ((uint32_t)(k >> shift) & 255u) : (256u + lane);
                uint32_t peers = warp_digit_peers(d, valid);
                uint32_t base = valid ? sm.whist[warp][d] : 0u;
                uint32_t pos = base + __popc(peers & ((1u << lane) - 1u));
                if (valid) {
                    kout[pos] = k;
                    iout[pos] = ix;
                }
                __syncwarp();
                if (valid && lane == __ffs(peers) - 1) sm.whist[warp][d] = base + __popc(peers);
                __syncwarp();
            }
            __syncthreads();  // scratch writes visible CTA-wide before the next pass reads them
            K *tk = kin; kin = kout; kout = tk;
            uint16_t *ti = iin; iin = iout; iout = ti;
        }
        // ---- 3. dense rank = number of distinct keys strictly below
        uint32_t cnt = 0;
        for (int ib = lo; ib < hi; ib += 32) {
            const int i = ib + lane;
            bool head = (i < hi) && (i > 0) && (kin[i] != kin[i - 1]);
            cnt += __popc(__ballot_sync(0xffffffffu, head));
        }
        if (lane == 0) sm.wtot[warp] = cnt;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; ++w) base += sm.wtot[w];
        for (int ib = lo; ib < hi; ib += 32) {
            const int i = ib + lane;
            bool head = (i < hi) && (i > 0) && (kin[i] != kin[i - 1]);
            uint32_t b = __ballot_sync(0xffffffffu, head);
            uint32_t r = base + __popc(b & (0xffffffffu >> (31 - lane)));
            if (i < hi) out[iin[i]] = (uint16_t)r;
            base += __popc(b);
        }
        if (warp == RANK_WARPS - 1 && lane == 0) {
            uint32_t total = 0;
            for (int w = 0; w < RANK_WARPS; ++w) total += sm.wtot[w];
            atomicMax(&flags[0], (int)total);
        }
        __syncthreads();  // smem + scratch reused by the next cell
    }
}


// ------------------------------------------------------------------------------------------------
// Full sort in SHARED memory (tie-free / real-valued cells; round 2).  rank_kernel's sort path ping-pongs keys + indices
// through a global scratch in ~40 sequential steps per warp and pass, each waiting for L2 and for a warp-wide digit
// match: 630 us per cfg3 cell.  Here only the KEYS are sorted -- two shared-memory buffers, stable LSD radix over the
// varying bytes, warp-contiguous ranges with per-warp digit histograms, digit peers from eight ballots (`match.any`
// iterates once per distinct digit: half of all stall samples of the first version sat behind it), up to 32 warps so that
// a warp walks few steps --, then the distinct keys are compacted and every gene's dense rank is a branch-free binary
// search in that array, eight look-ups interleaved (the row is re-read from L2).
// (Tried: a bitonic network in place of the radix passes -- no dependent chains, but 120 steps x 32 768 keys are bound by
// shared-memory bandwidth: 250 us per cfg3 cell against 100.)
// Used when 2 n keys fit in shared memory (float32: n <= ~24 000; float64: ~12 000); else rank_kernel sorts.
// ------------------------------------------------------------------------------------------------
template <int NW> struct alignas(16) SortSmem {
    uint32_t whist[NW][256];
    uint32_t dig_total[256];
    uint32_t wtot[NW];
    unsigned long long vary[NW];
    unsigned long long vary_all;
};

// NW warps per CTA: 32 for long rows (a warp walks its range in steps of 32 keys, one after the other: few steps per
// warp), 8 for short ones (small CTAs, many cells in flight per SM)
template <typename T, int NW>
__global__ void __launch_bounds__(NW * 32)
rank_sort_kernel(const T *__restrict__ x, int64_t ld, const int32_t *__restrict__ cell_rows,
                 const int32_t *__restrict__ gene_cols, int n, uint16_t *__restrict__ rank_out, int64_t rank_ld,
                 int32_t *__restrict__ flags, const int32_t *__restrict__ todo2, const int32_t *__restrict__ ctl2) {
    typedef typename KeyOf<T>::type K;
    __shared__ SortSmem<NW> sm;
    extern __shared__ __align__(16) uint8_t sort_dyn[];
    const int n_al = (n + 31) & ~31;
    K *bufA = reinterpret_cast<K *>(sort_dyn), *bufB = bufA + n_al;
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_warp = ((n + NW * 32 - 1) / (NW * 32)) * 32;
    const int lo = min(warp * per_warp, n), hi = min(lo + per_warp, n);
    const int n_work = ctl2[1];
    for (int wi = blockIdx.x; wi < n_work; wi += gridDim.x) {
        const int64_t slot = todo2[wi];
        uint16_t *out = rank_out + slot * rank_ld;
        const T *xr = x + (int64_t)cell_rows[slot] * ld;
        // ---- 1. keys (eight independent loads in flight per thread: the row comes from DRAM) and the key bits that vary
        bool nan = false;
        const K first = key_of(xr[gene_cols[0]], nan);
        K vary = 0;
        for (int g0 = tid; g0 < n; g0 += NT * 8) {
            T v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int g = g0 + j * NT;
                v[j] = g < n ? xr[gene_cols[g]] : (T)0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int g = g0 + j * NT;
                if (g < n) {
                    const K k = key_of(v[j], nan);
                    bufA[g] = k;
                    vary |= (k ^ first);
                }
            }
        }
        if (nan) flags[1] = 1;
        unsigned long long v64 = (unsigned long long)vary;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v64 |= __shfl_xor_sync(0xffffffffu, v64, o);
        if (lane == 0) sm.vary[warp] = v64;
        __syncthreads();
        if (tid == 0) {
            unsigned long long a = 0;
            for (int w = 0; w < NW; ++w) a |= sm.vary[w];
            sm.vary_all = a;
        }
        __syncthreads();
        const unsigned long long vary_all = sm.vary_all;
        K *kin = bufA, *kout = bufB;
        // ---- 2. stable LSD radix passes over the varying bytes
        for (int byte = 0; byte < (int)sizeof(K); ++byte) {
            if (((vary_all >> (8 * byte)) & 0xffull) == 0) continue;
            const int shift = 8 * byte;
            for (int i = tid; i < NW * 256; i += NT) (&sm.whist[0][0])[i] = 0;
            __syncthreads();
            for (int ib = lo; ib < hi; ib += 32) {
                const int i = ib + lane;
                const bool valid = i < hi;
                const uint32_t d = valid ? ((uint32_t)(kin[i] >> shift) & 255u) : 0u;
                const uint32_t peers = warp_digit_peers(d, valid);
                if (valid && lane == __ffs(peers) - 1) sm.whist[warp][d] += __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t t = 0;
                for (int w = 0; w < NW; ++w) t += sm.whist[w][tid];
                sm.dig_total[tid] = t;
            }
            __syncthreads();
            if (warp == 0) {  // exclusive scan of the 256 digit totals, 8 per lane
                uint32_t v[8], t = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = sm.dig_total[lane * 8 + j]; t += v[j]; }
                uint32_t incl = t;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += u;
                }
                uint32_t run = incl - t;
#pragma unroll
                for (int j = 0; j < 8; ++j) { sm.dig_total[lane * 8 + j] = run; run += v[j]; }
            }
            __syncthreads();
            if (tid < 256) {  // digit-major, warp-minor start offsets
                uint32_t run = sm.dig_total[tid];
                for (int w = 0; w < NW; ++w) {
                    const uint32_t c = sm.whist[w][tid];
                    sm.whist[w][tid] = run;
                    run += c;
                }
            }
            __syncthreads();
            for (int ib = lo; ib < hi; ib += 32) {
                const int i = ib + lane;
                const bool valid = i < hi;
                const K k = valid ? kin[i] : (K)0;
                const uint32_t d = valid ? ((uint32_t)(k >> shift) & 255u) : 0u;
                const uint32_t peers = warp_digit_peers(d, valid);
                const uint32_t base = valid ? sm.whist[warp][d] : 0u;
                if (valid) kout[base + __popc(peers & ((1u << lane) - 1u))] = k;
                __syncwarp();
                if (valid && lane == __ffs(peers) - 1) sm.whist[warp][d] = base + __popc(peers);
                __syncwarp();
            }
            __syncthreads();
            K *t = kin; kin = kout; kout = t;
        }
        // ---- 3. the distinct keys, compacted into the other buffer
        uint32_t cnt = 0;
        for (int ib = lo; ib < hi; ib += 32) {
            const int i = ib + lane;
            const bool head = (i < hi) && (i == 0 || kin[i] != kin[i - 1]);
            cnt += __popc(__ballot_sync(0xffffffffu, head));
        }
        if (lane == 0) sm.wtot[warp] = cnt;
        __syncthreads();
        uint32_t base = 0, D = 0;
        for (int w = 0; w < NW; ++w) {
            if (w < warp) base += sm.wtot[w];
            D += sm.wtot[w];
        }
        for (int ib = lo; ib < hi; ib += 32) {
            const int i = ib + lane;
            const bool head = (i < hi) && (i == 0 || kin[i] != kin[i - 1]);
            const uint32_t b = __ballot_sync(0xffffffffu, head);
            if (head) kout[base + __popc(b & ((1u << lane) - 1u))] = kin[i];
            base += __popc(b);
        }
        if (tid == 0) atomicMax(&flags[0], (int)D - 1);
        __syncthreads();
        // ---- 4. dense rank = position of the gene's key among the distinct keys
        int steps = 0;  // ceil(log2(D)) halvings find the lower bound
        while ((1u << steps) < D) ++steps;
        for (int g0 = tid; g0 < n; g0 += NT * 8) {  // eight look-ups interleaved: loads (L2) and probes (shared memory) overlap
            K k[8];
            int pos[8], len[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int g = g0 + j * NT;
                bool dummy = false;
                k[j] = key_of(g < n ? xr[gene_cols[g]] : (T)0, dummy);
                pos[j] = 0;
                len[j] = (int)D;
            }
            for (int st = 0; st <= steps; ++st) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {  // branch-free lower bound: [pos, pos + len) always holds the answer
                    const int half = len[j] >> 1;
                    const bool right = len[j] > 0 && kout[pos[j] + half] < k[j];
                    pos[j] = right ? pos[j] + half + 1 : pos[j];
                    len[j] = right ? len[j] - half - 1 : half;
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int g = g0 + j * NT;
                if (g < n) out[g] = (uint16_t)pos[j];
            }
        }
        __syncthreads();  // buffers reused by the next cell
    }
}

// ------------------------------------------------------------------------------------------------
// Streaming form of the count fast path for DENSE column spans (sandbag: all / almost all genes of a row; the
// compact [cells, U] matrices of the cyclone host pipeline; densified CSR chunks).
//
// The kernel above loads a row, then runs its shared-memory phases, then loads the next row: with ~3 CTAs per SM
// only about a third of the bytes HBM needs in flight are in flight (1.66 TB/s on cfg3).  Here a producer warp streams
// the rows of ALL cells of the CTA through a ring of 16 KB chunks with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier complete_tx), running ahead of the consumers by the depth of the ring -- across cell boundaries -- so the
// loads of cell i+1 overlap the prefix / look-up / store phases of cell i.  Consumers: pass 1 per chunk (count test,
// u16 staging, presence bitmap), then prefix popcount, then look-up with 16-byte stores of eight ranks.
// Cells that are not small counts (any real value, NaN) are appended to a to-do list and ranked by rank_kernel
// afterwards; a probe of a few cells decides up front whether the matrix is count-like at all.
// ------------------------------------------------------------------------------------------------
constexpr int STREAM_CHUNK = 16384;  // bytes per ring stage
constexpr int STREAM_MAX_STAGES = 8;

constexpr int LUT_N = 2048;  // counts below this are ranked through a look-up table
struct alignas(16) StreamSmem {
    uint32_t bitmap[BM_WORDS];
    uint32_t bm_prefix[BM_WORDS];
    uint32_t wtot[32];
    uint64_t full[STREAM_MAX_STAGES], empty[STREAM_MAX_STAGES];
    int bad[2];
    int dirty;               // a count >= LUT_N was seen: bitmap words >= LUT_N / 32 are in use
    alignas(16) uint8_t seen[LUT_N];     // presence of the counts < LUT_N: plain byte stores, no atomics (writing 1 is idempotent)
    alignas(16) uint16_t lut[LUT_N];     // rank of every count < LUT_N of the current cell
};

// is v an integer in [0, 65536)?  *u = that integer.  float: no conversion instructions (XU pipe) -- adding 2^23 leaves
// the integer part in the low mantissa bits, and the bit pattern of a non-negative float orders like the float
// (-0.0 is turned down here and ranked by the general kernel, which ties it with +0.0 all the same).
__device__ __forceinline__ bool small_count(float v, uint32_t &u) {
    const float t = v + 8388608.0f;
    u = __float_as_uint(t) & 0xffffu;
    return (t - 8388608.0f == v) && (__float_as_uint(v) < 0x47800000u);
}
__device__ __forceinline__ bool small_count(double v, uint32_t &u) {
    const bool ok = is_small_count(v);
    u = ok ? (uint32_t)v : 0u;
    return ok;
}

__device__ __forceinline__ void cons_sync(int n_cons) { asm volatile("bar.sync 1, %0;" ::"r"(n_cons) : "memory"); }

template <typename T>
__global__ void probe_counts_kernel(const T *__restrict__ x, int64_t ld, const int32_t *__restrict__ cell_rows,
                                    int64_t n_slots, const int32_t *__restrict__ gene_cols, int n, int32_t *__restrict__ ctl) {
    const int64_t slot = (int64_t)blockIdx.x * n_slots / gridDim.x;
    const int row = cell_rows[slot];
    if (row < 0) return;
    bool ok = true;
    for (int g = threadIdx.x; g < min(n, 2048); g += blockDim.x) ok = ok && is_small_count(x[(int64_t)row * ld + gene_cols[g]]);
    if (!ok) ctl[0] = 0;
}

__global__ void build_inv_kernel(const int32_t *__restrict__ gene_cols, int n, int c_min, int32_t *__restrict__ inv) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n) inv[gene_cols[g] - c_min] = g;
}

template <typename T, int NCONS>
__global__ void __launch_bounds__(NCONS + 32, NCONS == 512 ? 2 : 1)
rank_stream_kernel(const T *__restrict__ x, int64_t ld, const int32_t *__restrict__ cell_rows, int64_t n_slots, int n,
                   int c_min, int span, const int32_t *__restrict__ inv, int n_stages, uint16_t *__restrict__ rank_out,
                   int64_t rank_ld, int32_t *__restrict__ flags, int32_t *__restrict__ todo, int32_t *__restrict__ ctl) {
    constexpr int CHE = STREAM_CHUNK / (int)sizeof(T);  // elements per chunk
    constexpr int PFX = 512, WPT = BM_WORDS / PFX;      // the prefix pass runs on the first 512 consumers, 4 words each
    static_assert(NCONS % 32 == 0 && NCONS >= PFX && NCONS + 32 <= 1024, "consumer warps + one producer warp");
    __shared__ StreamSmem sm;
    extern __shared__ __align__(128) uint8_t dyn[];
    uint8_t *ring = dyn;                                                           // [n_stages][STREAM_CHUNK]
    uint16_t *vals = reinterpret_cast<uint16_t *>(dyn + (size_t)n_stages * STREAM_CHUNK);  // [n] staged counts
    if (ctl[0] == 0) return;  // the probe found real-valued cells: rank_kernel takes everything
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < n_stages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], NCONS / 32);
        }
        sm.bad[0] = sm.bad[1] = 0;
        sm.dirty = 0;
        fence_mbar_init();
    }
    for (int i = tid; i < BM_WORDS; i += NCONS + 32) sm.bitmap[i] = 0;
    for (int i = tid; i < LUT_N / 4; i += NCONS + 32) reinterpret_cast<uint32_t *>(sm.seen)[i] = 0;
    __syncthreads();
    const uintptr_t xb = reinterpret_cast<uintptr_t>(x);

    if (warp == NCONS / 32) {  // ---- producer warp: one lane issues the bulk copies of every chunk of every cell
        if (lane == 0) {
            uint32_t q = 0;
            for (int64_t slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
                const int row = cell_rows[slot];
                if (row < 0) continue;
                const uintptr_t a = xb + ((uintptr_t)row * (uintptr_t)ld + (uintptr_t)c_min) * sizeof(T);
                const uintptr_t a0 = a & ~(uintptr_t)15, e1 = (a + (uintptr_t)span * sizeof(T) + 15) & ~(uintptr_t)15;
                for (uintptr_t off = 0; a0 + off < e1; off += STREAM_CHUNK, ++q) {
                    const int s = (int)(q % (uint32_t)n_stages);
                    const uint32_t use = q / (uint32_t)n_stages;
                    if (use > 0) mbar_wait(&sm.empty[s], (use - 1) & 1);
                    const uint32_t bytes = (uint32_t)min((uintptr_t)STREAM_CHUNK, e1 - a0 - off);
                    mbar_arrive_expect_tx(&sm.full[s], bytes);
                    bulk_g2s(ring + (size_t)s * STREAM_CHUNK, reinterpret_cast<const void *>(a0 + off), bytes, &sm.full[s]);
                }
            }
        }
        return;
    }

    // ---- consumers
    constexpr int EPV = 16 / (int)sizeof(T);  // elements per 16-byte shared-memory load
    uint32_t q = 0;
    int par = 0;
    for (int64_t slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
        uint16_t *out = rank_out + slot * rank_ld;
        const int row = cell_rows[slot];
        if (row < 0) {  // padding cell: ties in every gene
            for (int g = tid; g < n; g += NCONS) out[g] = 0;
            continue;
        }
        const uintptr_t a = xb + ((uintptr_t)row * (uintptr_t)ld + (uintptr_t)c_min) * sizeof(T);
        const uintptr_t a0 = a & ~(uintptr_t)15, e1 = (a + (uintptr_t)span * sizeof(T) + 15) & ~(uintptr_t)15;
        const int shift = (int)((a - a0) / sizeof(T));  // elements of the first chunk before column c_min
        // the look-up pass writes eight ranks per 16-byte store from the first 16-byte boundary of the output row on
        // (`head` scalar ranks before it); the counts are staged at vals[pad + g] so that those groups of eight are
        // 16-byte aligned in shared memory too (one LDS.128 instead of eight 2-byte loads with bank conflicts)
        const int head = min(n, (int)(((16 - (reinterpret_cast<uintptr_t>(out) & 15)) & 15) >> 1));
        const int pad = (8 - head) & 7;
        // ---- pass 1, chunk by chunk as the bulk copies land: count test, u16 staging, presence
        bool small = true;
        int col0 = -shift;  // column (relative to c_min) of element 0 of the chunk
        for (uintptr_t off = 0; a0 + off < e1; off += STREAM_CHUNK, ++q, col0 += CHE) {
            const int s = (int)(q % (uint32_t)n_stages);
            mbar_wait(&sm.full[s], (q / (uint32_t)n_stages) & 1);
            const uint4 *src = reinterpret_cast<const uint4 *>(ring + (size_t)s * STREAM_CHUNK);
#pragma unroll 2
            for (int e4 = tid; e4 < CHE / EPV; e4 += NCONS) {
                const uint4 raw = src[e4];
                T v[EPV];
                if (sizeof(T) == 4) {
                    v[0] = (T)__uint_as_float(raw.x);
                    v[1] = (T)__uint_as_float(raw.y);
                    v[EPV - 2] = (T)__uint_as_float(raw.z);
                    v[EPV - 1] = (T)__uint_as_float(raw.w);
                } else {
                    v[0] = (T)__hiloint2double((int)raw.y, (int)raw.x);
                    v[EPV - 1] = (T)__hiloint2double((int)raw.w, (int)raw.z);
                }
#pragma unroll
                for (int k = 0; k < EPV; ++k) {
                    const int col = col0 + e4 * EPV + k;
                    bool valid = (unsigned)col < (unsigned)span;
                    int g = col;
                    if (inv) {
                        g = valid ? __ldg(inv + col) : -1;
                        valid = g >= 0;
                    }
                    uint32_t u;
                    const bool ok = small_count(v[k], u);
                    small = small && (ok || !valid);
                    // (dummy-slot stores instead of these two branches: 370 instead of 346 us on cfg3)
                    const bool use = valid && ok;
                    if (use) {
                        vals[pad + g] = (uint16_t)u;
                        if (u < LUT_N) sm.seen[u] = 1;
                    }
                    if (use && u >= LUT_N) {  // rare
                        sm.dirty = 1;
                        const uint32_t bit = 1u << (u & 31);
                        if (!(*(volatile uint32_t *)&sm.bitmap[u >> 5] & bit)) atomicOr(&sm.bitmap[u >> 5], bit);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
        if (!small) sm.bad[par] = 1;
        cons_sync(NCONS);  // seen, bitmap, vals, bad complete
        const bool bad = sm.bad[par] != 0;
        const bool dirty = sm.dirty != 0;
        if (tid == 0) sm.bad[par ^ 1] = 0;  // next cell's flag (nobody reads it before the next cons_sync)
        par ^= 1;
        if (!bad) {
            // presence bytes -> bitmap words 0 .. LUT_N/32 - 1 (assigned, not OR-ed: no zeroing between cells)
            if (tid < LUT_N / 32) {
                const uint4 *sp = reinterpret_cast<const uint4 *>(sm.seen + tid * 32);
                uint32_t w = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4 b = sp[h];
                    const uint32_t x4[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j)  // four 0/1 bytes -> four bits: byte i lands on bit 24 + i of the product
                        w |= (((x4[j] & 0x01010101u) * 0x01020408u) >> 24 & 0xfu) << (h * 16 + j * 4);
                }
                sm.bitmap[tid] = w;
            }
            cons_sync(NCONS);
            for (int i = tid; i < LUT_N / 4; i += NCONS) reinterpret_cast<uint32_t *>(sm.seen)[i] = 0;  // for the next cell
            // exclusive prefix popcount over the bitmap words: WPT words per thread, warp scan, warp totals
            uint32_t c[WPT], t4 = 0, incl = 0;
            if (tid < PFX) {
#pragma unroll
                for (int j = 0; j < WPT; ++j) { c[j] = __popc(sm.bitmap[tid * WPT + j]); t4 += c[j]; }
                incl = t4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) sm.wtot[warp] = incl;
            }
            cons_sync(NCONS);
            if (tid < PFX) {
                uint32_t base = 0;
                for (int w = 0; w < warp; ++w) base += sm.wtot[w];
                uint32_t run = base + incl - t4;
#pragma unroll
                for (int j = 0; j < WPT; ++j) { sm.bm_prefix[tid * WPT + j] = run; run += c[j]; }
                if (tid == PFX - 1) atomicMax(&flags[0], (int)run - 1);  // distinct values - 1
            }
            cons_sync(NCONS);
            auto rank_slow = [&](uint32_t u) -> uint32_t {
                return sm.bm_prefix[u >> 5] + __popc(sm.bitmap[u >> 5] & ((1u << (u & 31)) - 1u));
            };
            for (int v = tid; v < LUT_N; v += NCONS) sm.lut[v] = (uint16_t)rank_slow((uint32_t)v);
            cons_sync(NCONS);
            // ---- look-up: one table read per count; eight ranks per 16-byte store
            auto rank_of = [&](uint32_t u) -> uint32_t { return u < LUT_N ? (uint32_t)sm.lut[u] : rank_slow(u); };
            if (tid < head) out[tid] = (uint16_t)rank_of(vals[pad + tid]);
            const int n_vec = (n - head) >> 3;
            const uint4 *v8p = reinterpret_cast<const uint4 *>(vals + pad + head);  // pad + head is 0 or 8
#pragma unroll 2
            for (int v8 = tid; v8 < n_vec; v8 += NCONS) {
                const uint4 u8 = v8p[v8];
                const uint32_t in[4] = {u8.x, u8.y, u8.z, u8.w};
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) w[j] = rank_of(in[j] & 0xffffu) | (rank_of(in[j] >> 16) << 16);
                *reinterpret_cast<uint4 *>(out + head + v8 * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            for (int g = head + n_vec * 8 + tid; g < n; g += NCONS) out[g] = (uint16_t)rank_of(vals[pad + g]);
        } else {
            for (int i = tid; i < LUT_N / 4; i += NCONS) reinterpret_cast<uint32_t *>(sm.seen)[i] = 0;
            if (tid == 0) todo[atomicAdd(&ctl[1], 1)] = (int32_t)slot;
        }
        cons_sync(NCONS);  // everybody is done with bitmap / prefix / lut / vals
        if (dirty) {  // rare: counts >= LUT_N used the upper bitmap words
            for (int i = LUT_N / 32 + tid; i < BM_WORDS; i += NCONS) sm.bitmap[i] = 0;
            if (tid == 0) sm.dirty = 0;
            cons_sync(NCONS);
        }
    }
}

}  // namespace

// Launch configuration of the streaming kernel for rows of `span` columns holding n kept genes: consumers per CTA
// (512 with two CTAs per SM when they fit, else 1024 with one), ring stages, dynamic shared memory.  ok = false: no fit.
struct StreamCfg { bool ok; int n_cons, n_stages, grid_per_sm; size_t dyn; };
static StreamCfg stream_config(pp_ctx *ctx, int64_t n) {
    StreamCfg c = {false, 0, 0, 0, 0};
    const size_t vals = (((size_t)n + 16) * 2 + 127) & ~(size_t)127;  // staged counts (+ alignment pad + a dummy slot)
    const size_t stat = sizeof(StreamSmem);
    const size_t sm_total = 228 * 1024, reserve = 1024;  // per SM; every resident CTA also reserves 1 KB
    const size_t dyn_max1 = (size_t)ctx->smem_optin - stat, dyn_max2 = sm_total / 2 - reserve - stat;
    if (vals + 3 * STREAM_CHUNK <= dyn_max2) {  // two CTAs of 512 consumers per SM, >= 3 stages each
        c.n_cons = 512;
        c.grid_per_sm = 2;
        c.n_stages = (int)std::min<size_t>(STREAM_MAX_STAGES, (dyn_max2 - vals) / STREAM_CHUNK);
    } else if (vals + 2 * STREAM_CHUNK <= dyn_max1) {  // one CTA of 960 consumers, as many stages as fit
        c.n_cons = 960;
        c.grid_per_sm = 1;
        c.n_stages = (int)std::min<size_t>(STREAM_MAX_STAGES, (dyn_max1 - vals) / STREAM_CHUNK);
    } else {
        return c;
    }
    c.dyn = (size_t)c.n_stages * STREAM_CHUNK + vals;
    c.ok = true;
    return c;
}

template <typename T>
static int rank_launch_typed(pp_ctx *ctx, const T *d_x, int64_t ld, const int32_t *d_cell_rows, int64_t n_slots,
                             const int32_t *d_gene_cols, int64_t n_genes, uint16_t *d_rank, int64_t rank_ld,
                             int32_t *d_flags, const int32_t *h_gene_cols) {
    cudaStream_t st = ctx->stream;
    // ---- streaming count path: ascending columns covering at least half of their span, rows of >= 8 KB, 16-byte
    //      aligned matrix base (the bulk copies start at the 16-byte boundary below a row's first kept column)
    DevBuf d_inv, d_todo, d_ctl;
    const int32_t *todo = nullptr, *ctl = nullptr;
    bool stream = h_gene_cols != nullptr && n_genes * (int64_t)sizeof(T) >= 8192 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0;
    int c_min = 0, span = 0;
    if (stream) {
        for (int64_t g = 1; g < n_genes && stream; ++g) stream = h_gene_cols[g] > h_gene_cols[g - 1];
        c_min = h_gene_cols[0];
        span = h_gene_cols[n_genes - 1] - c_min + 1;
        stream = stream && (int64_t)span <= 2 * n_genes;
    }
    const StreamCfg cfg = stream ? stream_config(ctx, n_genes) : StreamCfg{false, 0, 0, 0, 0};
    if (stream && cfg.ok) {
        PP_CUDA(d_ctl.alloc(sizeof(int32_t) * 2, st));
        PP_CUDA(d_todo.alloc(sizeof(int32_t) * n_slots, st));
        const int32_t init[2] = {1, 0};
        PP_CUDA(cudaMemcpyAsync(d_ctl.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
        probe_counts_kernel<T><<<(unsigned)std::min<int64_t>(n_slots, 64), 256, 0, st>>>(
            d_x, ld, d_cell_rows, n_slots, d_gene_cols, (int)n_genes, d_ctl.as<int32_t>());
        PP_CHECK_LAUNCH();
        const int32_t *inv = nullptr;
        if ((int64_t)span != n_genes) {
            PP_CUDA(d_inv.alloc(sizeof(int32_t) * (size_t)span, st));
            PP_CUDA(cudaMemsetAsync(d_inv.p, 0xff, sizeof(int32_t) * (size_t)span, st));
            build_inv_kernel<<<(unsigned)((n_genes + 255) / 256), 256, 0, st>>>(d_gene_cols, (int)n_genes, c_min, d_inv.as<int32_t>());
            PP_CHECK_LAUNCH();
            inv = d_inv.as<int32_t>();
        }
        const int grid = (int)std::min<int64_t>(n_slots, (int64_t)ctx->n_sm * cfg.grid_per_sm);
        if (cfg.n_cons == 512) {
            PP_CUDA(cudaFuncSetAttribute(rank_stream_kernel<T, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dyn));
            rank_stream_kernel<T, 512><<<grid, 512 + 32, cfg.dyn, st>>>(d_x, ld, d_cell_rows, n_slots, (int)n_genes, c_min, span, inv,
                                                                       cfg.n_stages, d_rank, rank_ld, d_flags,
                                                                       d_todo.as<int32_t>(), d_ctl.as<int32_t>());
        } else {
            PP_CUDA(cudaFuncSetAttribute(rank_stream_kernel<T, 960>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dyn));
            rank_stream_kernel<T, 960><<<grid, 960 + 32, cfg.dyn, st>>>(d_x, ld, d_cell_rows, n_slots, (int)n_genes, c_min, span,
                                                                         inv, cfg.n_stages, d_rank, rank_ld, d_flags,
                                                                         d_todo.as<int32_t>(), d_ctl.as<int32_t>());
        }
        PP_CHECK_LAUNCH();
        todo = d_todo.as<int32_t>();
        ctl = d_ctl.as<int32_t>();
    }
    // ---- general kernel: everything (no streaming pass), or the cells the streaming pass handed over; cells that need
    //      a full sort are handed on to rank_sort_kernel when 2 n keys fit in shared memory
    const size_t ksz = sizeof(T);
    const size_t n_al = ((size_t)n_genes + 15) & ~(size_t)15;
    const size_t per_cta = 2 * n_al * ksz + 2 * n_al * 2;
    const size_t dyn = ((size_t)n_genes * 2 + 15) & ~(size_t)15;
    const size_t sort_dyn = 2 * (((size_t)n_genes + 31) & ~(size_t)31) * ksz;
    const bool sort_small = n_genes <= 4096;  // rank_sort_kernel<T, 8> / <T, 32>
    const size_t sort_static = sort_small ? sizeof(SortSmem<8>) : sizeof(SortSmem<32>);
    const bool smem_sort = sort_dyn + sort_static + 1024 <= (size_t)ctx->smem_optin && getenv("PP_NO_SMEM_SORT") == nullptr;
    DevBuf d_todo2, d_ctl2;
    if (smem_sort) {
        PP_CUDA(d_ctl2.alloc(sizeof(int32_t) * 2, st));
        PP_CUDA(d_todo2.alloc(sizeof(int32_t) * n_slots, st));
        PP_CUDA(cudaMemsetAsync(d_ctl2.p, 0, sizeof(int32_t) * 2, st));
    }
    PP_CUDA(cudaFuncSetAttribute(rank_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    // persistent CTAs: exactly as many as are resident at once (a partial second wave would idle half the SMs)
    int per_sm = 1;
    PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rank_kernel<T>, RANK_THREADS, dyn));
    const int64_t resident = (int64_t)ctx->n_sm * std::max(per_sm, 1);
    const int grid = (int)std::min<int64_t>(n_slots, resident);
    uint8_t *scratch = nullptr;
    if (!smem_sort) {  // global ping-pong buffers of rank_kernel's own sort path; stream-ordered reuse
        scratch = (uint8_t *)pp_slab(ctx, SLAB_SORT, per_cta * (size_t)grid);
        if (!scratch) return PP_ENOMEM;
    }
    rank_kernel<T><<<grid, RANK_THREADS, dyn, st>>>(d_x, ld, d_cell_rows, n_slots, d_gene_cols, (int)n_genes, d_rank, rank_ld,
                                                    scratch, per_cta, d_flags, todo, ctl,
                                                    smem_sort ? d_todo2.as<int32_t>() : nullptr,
                                                    smem_sort ? d_ctl2.as<int32_t>() : nullptr);
    PP_CHECK_LAUNCH();
    if (smem_sort) {
        int sort_per_sm = 1;
        if (sort_small) {
            PP_CUDA(cudaFuncSetAttribute(rank_sort_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_dyn));
            PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sort_per_sm, rank_sort_kernel<T, 8>, 256, sort_dyn));
            const int sgrid = (int)std::min<int64_t>(n_slots, (int64_t)ctx->n_sm * std::max(sort_per_sm, 1));
            rank_sort_kernel<T, 8><<<sgrid, 256, sort_dyn, st>>>(d_x, ld, d_cell_rows, d_gene_cols, (int)n_genes, d_rank, rank_ld, d_flags,
                                                                 d_todo2.as<int32_t>(), d_ctl2.as<int32_t>());
        } else {
            PP_CUDA(cudaFuncSetAttribute(rank_sort_kernel<T, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sort_dyn));
            PP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sort_per_sm, rank_sort_kernel<T, 32>, 1024, sort_dyn));
            const int sgrid = (int)std::min<int64_t>(n_slots, (int64_t)ctx->n_sm * std::max(sort_per_sm, 1));
            rank_sort_kernel<T, 32><<<sgrid, 1024, sort_dyn, st>>>(d_x, ld, d_cell_rows, d_gene_cols, (int)n_genes, d_rank, rank_ld,
                                                                   d_flags, d_todo2.as<int32_t>(), d_ctl2.as<int32_t>());
        }
        PP_CHECK_LAUNCH();
    }
    return PP_OK;
}

int pp_rank_launch(pp_ctx *ctx, const void *d_x, int x_dtype, int64_t ld, const int32_t *d_cell_rows,
                   int64_t n_slots, const int32_t *d_gene_cols, int64_t n_genes, uint16_t *d_rank,
                   int64_t rank_ld, int32_t *d_flags, const int32_t *h_gene_cols) {
    if (n_genes > 65535) return pp_fail(ctx, PP_ELIMIT, "more than 65535 kept genes per cell (%lld)", (long long)n_genes);
    if (n_slots == 0 || n_genes == 0) return PP_OK;
    if (x_dtype == PP_F64)
        return rank_launch_typed<double>(ctx, (const double *)d_x, ld, d_cell_rows, n_slots, d_gene_cols, n_genes, d_rank,
                                         rank_ld, d_flags, h_gene_cols);
    return rank_launch_typed<float>(ctx, (const float *)d_x, ld, d_cell_rows, n_slots, d_gene_cols, n_genes, d_rank, rank_ld,
                                    d_flags, h_gene_cols);
}

// ------------------------------------------------------------------------------------------------
// CSR input (scipy.sparse.csr_matrix / AnnData.X): the reference calls .toarray() on the host
// (pypairs/utils.py:278-281); here the rows a chunk needs are densified on the device.
// ------------------------------------------------------------------------------------------------
namespace {

template <typename T>
__global__ void csr_col_nonzero_kernel(const int32_t *__restrict__ indices, const T *__restrict__ data, int64_t nnz,
                                       uint8_t *__restrict__ nz) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x)
        if (data[k] != (T)0) nz[indices[k]] = 1;
}

// one warp per output row
template <typename T>
__global__ void csr_densify_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                   const T *__restrict__ data, const int32_t *__restrict__ rows, int64_t n_out,
                                   const int32_t *__restrict__ colmap, T *__restrict__ out, int64_t out_ld) {
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_out) return;
    const int r = rows[i];
    if (r < 0) return;
    T *o = out + i * out_ld;
    for (int64_t k = indptr[r] + (threadIdx.x & 31); k < indptr[r + 1]; k += 32) {
        const int c = colmap[indices[k]];
        if (c >= 0) o[c] = data[k];
    }
}

// device-resident CSR: bit 0 of *flag = indptr decreases somewhere, bit 1 = a column index is out of range
__global__ void csr_validate_kernel(const void *__restrict__ indptr, int is64, int64_t n_rows,
                                    const int32_t *__restrict__ indices, int64_t p0, int64_t nnz, int64_t n_cols,
                                    int32_t *__restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += stride) {
        const int64_t a = is64 ? ((const int64_t *)indptr)[r] : (int64_t)((const int32_t *)indptr)[r];
        const int64_t b = is64 ? ((const int64_t *)indptr)[r + 1] : (int64_t)((const int32_t *)indptr)[r + 1];
        if (b < a) bad |= 1;
    }
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) {
        const int32_t c = indices[p0 + k];
        if (c < 0 || (int64_t)c >= n_cols) bad |= 2;
    }
    if (bad) atomicOr(flag, bad);
}

__global__ void rebase_indptr_kernel(const void *__restrict__ src, int is64, int64_t r0, int64_t n, int64_t base,
                                     int64_t *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = is64 ? ((const int64_t *)src)[r0 + i] : (int64_t)((const int32_t *)src)[r0 + i];
    dst[i] = v - base;
}

}  // namespace

int pp_csr_validate(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, const char *who) {
    if (!x || !x->indptr) return pp_fail(ctx, PP_EINVAL, "%s: NULL CSR argument", who);
    if (x->data_dtype != PP_F32 && x->data_dtype != PP_F64) return pp_fail(ctx, PP_EINVAL, "%s: CSR data_dtype must be PP_F32/PP_F64", who);
    if (n_rows <= 0 || n_cols <= 0) return pp_fail(ctx, PP_EINVAL, "%s: bad matrix shape", who);
    if (!x->is_device) {
        auto ip = [&](int64_t i) { return x->indptr_is_64 ? ((const int64_t *)x->indptr)[i] : (int64_t)((const int32_t *)x->indptr)[i]; };
        for (int64_t r = 0; r < n_rows; ++r)
            if (ip(r + 1) < ip(r)) return pp_fail(ctx, PP_EINVAL, "%s: CSR indptr decreases at row %lld", who, (long long)r);
        const int64_t nnz = ip(n_rows) - ip(0);
        if (nnz > 0 && (!x->indices || !x->data)) return pp_fail(ctx, PP_EINVAL, "%s: NULL CSR indices / data", who);
        const int32_t *idx = x->indices + ip(0);
        for (int64_t k = 0; k < nnz; ++k)
            if (idx[k] < 0 || idx[k] >= n_cols) return pp_fail(ctx, PP_EINVAL, "%s: CSR column index out of range", who);
    } else {
        // the same checks on the device: a malformed tensor must not turn into out-of-bounds accesses of
        // csr_col_nonzero_kernel / csr_densify_kernel
        PP_CUDA(cudaSetDevice(ctx->device));
        cudaStream_t st = ctx->stream;
        const size_t isz = x->indptr_is_64 ? 8 : 4;
        int64_t e64[2] = {0, 0};
        int32_t e32[2] = {0, 0};
        for (int j = 0; j < 2; ++j) {
            const int64_t row = j == 0 ? 0 : n_rows;
            void *dst = x->indptr_is_64 ? (void *)&e64[j] : (void *)&e32[j];
            PP_CUDA(cudaMemcpyAsync(dst, (const uint8_t *)x->indptr + isz * row, isz, cudaMemcpyDeviceToHost, st));
        }
        PP_CUDA(cudaStreamSynchronize(st));
        if (!x->indptr_is_64) { e64[0] = e32[0]; e64[1] = e32[1]; }
        const int64_t nnz = e64[1] - e64[0];
        if (e64[0] < 0 || nnz < 0) return pp_fail(ctx, PP_EINVAL, "%s: CSR indptr decreases", who);
        if (nnz > 0 && (!x->indices || !x->data)) return pp_fail(ctx, PP_EINVAL, "%s: NULL CSR indices / data", who);
        DevBuf d_flag;
        PP_CUDA(d_flag.alloc(sizeof(int32_t), st));
        PP_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int32_t), st));
        const unsigned grid = (unsigned)std::min<int64_t>((std::max(n_rows, nnz) + 255) / 256, (int64_t)ctx->n_sm * 16);
        csr_validate_kernel<<<std::max(grid, 1u), 256, 0, st>>>(x->indptr, x->indptr_is_64, n_rows, x->indices, e64[0], nnz,
                                                                n_cols, d_flag.as<int32_t>());
        PP_CHECK_LAUNCH();
        int32_t flag = 0;
        PP_CUDA(cudaMemcpyAsync(&flag, d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
        if (flag & 1) return pp_fail(ctx, PP_EINVAL, "%s: CSR indptr decreases", who);
        if (flag & 2) return pp_fail(ctx, PP_EINVAL, "%s: CSR column index out of range", who);
    }
    return PP_OK;
}

int pp_csr_to_device(pp_ctx *ctx, const pp_csr *x, int64_t r0, int64_t nr, DevCsr *out) {
    cudaStream_t st = ctx->stream;
    const size_t esz = x->data_dtype == PP_F64 ? 8 : 4;
    out->dtype = x->data_dtype;
    out->n_rows = nr;
    PP_CUDA(out->b_indptr.alloc(sizeof(int64_t) * (nr + 1), st));
    int64_t p0 = 0, p1 = 0;
    if (!x->is_device) {
        std::vector<int64_t> ip(nr + 1);
        for (int64_t i = 0; i <= nr; ++i)
            ip[i] = x->indptr_is_64 ? ((const int64_t *)x->indptr)[r0 + i] : (int64_t)((const int32_t *)x->indptr)[r0 + i];
        p0 = ip[0];
        p1 = ip[nr];
        for (int64_t i = 0; i <= nr; ++i) ip[i] -= p0;
        PP_CUDA(cudaMemcpyAsync(out->b_indptr.p, ip.data(), sizeof(int64_t) * (nr + 1), cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaStreamSynchronize(st));  // ip is a local
        out->nnz = p1 - p0;
        PP_CUDA(out->b_indices.alloc(sizeof(int32_t) * out->nnz, st));
        PP_CUDA(out->b_data.alloc(esz * out->nnz, st));
        if (out->nnz) {
            PP_CUDA(cudaMemcpyAsync(out->b_indices.p, x->indices + p0, sizeof(int32_t) * out->nnz, cudaMemcpyHostToDevice, st));
            PP_CUDA(cudaMemcpyAsync(out->b_data.p, (const uint8_t *)x->data + esz * p0, esz * out->nnz, cudaMemcpyHostToDevice, st));
        }
        ctx->tm.h2d_bytes += (int64_t)((sizeof(int32_t) + esz) * out->nnz + sizeof(int64_t) * (nr + 1));
        out->indices = out->b_indices.as<int32_t>();
        out->data = out->b_data.p;
    } else {
        const size_t isz = x->indptr_is_64 ? 8 : 4;
        int64_t ends[2] = {0, 0};
        int32_t e32[2] = {0, 0};
        for (int j = 0; j < 2; ++j) {
            const int64_t row = j == 0 ? r0 : r0 + nr;
            if (x->indptr_is_64) {
                PP_CUDA(cudaMemcpyAsync(&ends[j], (const uint8_t *)x->indptr + isz * row, isz, cudaMemcpyDeviceToHost, st));
            } else {
                PP_CUDA(cudaMemcpyAsync(&e32[j], (const uint8_t *)x->indptr + isz * row, isz, cudaMemcpyDeviceToHost, st));
            }
        }
        PP_CUDA(cudaStreamSynchronize(st));
        if (!x->indptr_is_64) { ends[0] = e32[0]; ends[1] = e32[1]; }
        p0 = ends[0];
        p1 = ends[1];
        if (p1 < p0) return pp_fail(ctx, PP_EINVAL, "CSR indptr decreases");
        out->nnz = p1 - p0;
        rebase_indptr_kernel<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, st>>>(x->indptr, x->indptr_is_64, r0, nr + 1, p0,
                                                                          out->b_indptr.as<int64_t>());
        PP_CHECK_LAUNCH();
        out->indices = x->indices + p0;
        out->data = (const uint8_t *)x->data + esz * p0;
    }
    out->indptr = out->b_indptr.as<int64_t>();
    return PP_OK;
}

int pp_csr_col_nonzero(pp_ctx *ctx, const DevCsr &m, uint8_t *d_nz) {
    if (m.nnz == 0) return PP_OK;
    const unsigned grid = (unsigned)std::min<int64_t>((m.nnz + 255) / 256, (int64_t)ctx->n_sm * 16);
    if (m.dtype == PP_F64)
        csr_col_nonzero_kernel<double><<<grid, 256, 0, ctx->stream>>>(m.indices, (const double *)m.data, m.nnz, d_nz);
    else
        csr_col_nonzero_kernel<float><<<grid, 256, 0, ctx->stream>>>(m.indices, (const float *)m.data, m.nnz, d_nz);
    PP_CHECK_LAUNCH();
    return PP_OK;
}

int pp_csr_densify(pp_ctx *ctx, const DevCsr &m, const int32_t *d_rows, int64_t n_out, const int32_t *d_colmap,
                   void *d_out, int64_t out_ld) {
    if (n_out == 0) return PP_OK;
    const unsigned grid = (unsigned)((n_out + 7) / 8);
    if (m.dtype == PP_F64)
        csr_densify_kernel<double><<<grid, 256, 0, ctx->stream>>>(m.indptr, m.indices, (const double *)m.data, d_rows,
                                                                   n_out, d_colmap, (double *)d_out, out_ld);
    else
        csr_densify_kernel<float><<<grid, 256, 0, ctx->stream>>>(m.indptr, m.indices, (const float *)m.data, d_rows, n_out,
                                                                  d_colmap, (float *)d_out, out_ld);
    PP_CHECK_LAUNCH();
    return PP_OK;
}
