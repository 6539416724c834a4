// pp_sandbag.cu -- B200 all-gene-pairs marker kernel (replaces check_pairs,
// pypairs/tools/sandbag.py:160-199, and the np.where scan at :129).
//
// Data layout in HBM
//   rank   u16 [32*W][G]     per-cell dense ranks, cells class-sorted, every class padded
//                            to a multiple of 32 cells with all-tie cells (pp_rank.cu)
//   planes u32 [NB][W][B][64] bit-sliced ranks: gene block nb (64 genes), cell word w (32
//                            cells of ONE class), plane p (bit p of the rank); bit c of the
//                            word belongs to cell 32*w + c.  A (gene block, run of words) is
//                            one contiguous byte range -> one 1-D TMA bulk copy.
// Pair kernel (pair_kernel): CTA tile = 128 a-genes x 64 b-genes, every thread owns an
// 8x4 register tile of gene pairs and walks all cell words.  For one word the two
// predicates of sandbag.py:179-182 are evaluated for 32 cells at once with the
// LSB->MSB recurrences
//        gt' = (a & ~b) | (~(a ^ b) & gt)        lt' = (~a & b) | (~(a ^ b) & lt)
// -- one LOP3 per plane per predicate -- and accumulated per class with POPC.  Ties set
// neither bit, padding cells tie everywhere, exactly the reference's "add to neither".
// Thresholds (sandbag.py:184-190) are applied in-register at every class boundary and the
// if/elif decision (:192-197) at the end; survivors are compacted with warp-aggregated
// atomics and radix-sorted into the row-major order np.where would produce.
#include <algorithm>

#include "pp_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// nz[g] = 1 if any row of column g is != 0  (the reference's ~np.all(data == 0, axis=0), utils.py:368)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
col_nonzero_kernel(const T *__restrict__ x, int64_t ld, int64_t n_rows, int64_t n_cols, uint8_t *__restrict__ nz) {
    const int64_t g = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (g >= n_cols) return;
    bool any = false;
    for (int64_t r0 = (int64_t)blockIdx.y * 64; r0 < n_rows && !any; r0 += (int64_t)gridDim.y * 64) {
        const int64_t r1 = min(r0 + 64, n_rows);
        for (int64_t r = r0; r < r1; ++r) any |= (x[r * ld + g] != (T)0);
    }
    if (any) nz[g] = 1;
}

// ------------------------------------------------------------------------------------------
// bit-slice: rank[32W][G] -> planes[NB][W][B][64]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bitslice_kernel(const uint16_t *__restrict__ rank, int64_t rank_ld, int n_genes, int n_words, int n_planes,
                uint32_t *__restrict__ planes) {
    const int w = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int g = (blockIdx.y * 8 + (threadIdx.x >> 5)) * 32 + lane;  // this lane's gene
    if (g - lane >= n_genes) return;
    uint32_t r[32];
    const uint16_t *src = rank + (int64_t)w * 32 * rank_ld + g;
#pragma unroll
    for (int c = 0; c < 32; ++c) r[c] = (g < n_genes) ? (uint32_t)src[(int64_t)c * rank_ld] : 0u;
    uint32_t *dst = planes + (((int64_t)(g >> 6) * n_words + w) * n_planes) * 64 + (g & 63);
    for (int p = 0; p < n_planes; ++p) {
        uint32_t word = 0;
#pragma unroll
        for (int c = 0; c < 32; ++c) word |= ((r[c] >> p) & 1u) << c;
        dst[(int64_t)p * 64] = word;
    }
}

// ------------------------------------------------------------------------------------------
// pair kernel
// ------------------------------------------------------------------------------------------
#ifndef PP_TILE_A
#define PP_TILE_A 128
#endif
constexpr int TILE_A = PP_TILE_A;     // a-genes per CTA = 1 or 2 gene blocks
constexpr int A_BLOCKS = TILE_A / 64;
constexpr int CTAS_PER_SM = TILE_A == 64 ? 2 : 1;  // 64 x 64 tiles: two CTAs share an SM, one can be in sweep 2
constexpr int TILE_B = 64;   // b-genes per CTA = 1 gene block
constexpr int STAGES = 3;

// a * b + c as an IMAD: keeps the counter updates off the ALU pipe, which the LOP3 stream saturates
__device__ __forceinline__ uint32_t fma_pipe_mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int N> __device__ __forceinline__ void load_words(const uint32_t *p, uint32_t (&w)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
        const uint4 v = *reinterpret_cast<const uint4 *>(p + i);
        w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
    }
}

struct Chunk {  // a run of cell words of one class
    int32_t w_begin;
    int32_t n_words;
    int32_t class_end;  // class index if this chunk closes the class, else -1
    int32_t pad;
};

struct PairParams {
    const uint32_t *planes;
    const Chunk *chunks;
    const int2 *tiles;       // (ta, tb) per CTA
    const int32_t *thr;      // per class, clipped to int32
    const int32_t *max_pos_dn;  // per class n_k - thr_k: num_neg >= thr_k needs num_pos <= this
    unsigned long long *out; // compacted (key << 16 | class)
    unsigned long long *out_count;
    unsigned long long out_cap;
    int n_genes, n_words, n_planes, n_chunks, n_classes, wc_max;
};

// Two sweeps over the cell words of a 128 x 64 gene tile, one kernel.
//  Sweep 1 (all pairs): only num_pos = #(x[g1] > x[g2]) is accumulated -- ONE LOP3 per plane per pair and word -- and
//     thresholded per class (nup, up_idx).  The decision of sandbag.py:192-197 can only emit when nup == 1 or
//     nup == C-1; and since num_pos + num_neg <= n_k, a class can only count as "down" if num_pos <= n_k - thr_k
//     (ncd = number of such classes): emitting needs ncd >= C-1 (first branch) or ncd >= 1 (second).  Pairs passing
//     these necessary conditions are the "candidates" -- well under 1 % of the pairs on expression data.
//  Sweep 2 (candidates only, 512 per round = one per consumer thread): the same operand stream is replayed and
//     num_neg = #(x[g1] < x[g2]) is accumulated for the candidates, thresholded (ndn, dn_idx), and the decision emits.
// So the second predicate of sandbag.py:179-182 is evaluated exactly where it can matter, halving the LOP3 work.
// Per-pair counters are plain 32-bit registers (any class size), the state packs nup | up_idx << 16 (<= 65 534 classes).
//
// TI x TJ = per-thread register tile of gene pairs (4 x 4), consumer thread grid (128/TI) x (64/TJ) = 512 threads.
// Warps 0 .. 15 consume; one extra warp is the TMA producer.  full[s] (1 arrival + transaction bytes) hands a stage to
// the consumers, empty[s] (one arrival per consumer warp) hands it back, so the consumer warps never meet at a CTA-wide
// barrier inside a sweep and drift up to STAGES-1 chunks apart.
constexpr int CAND_PER_ROUND = (TILE_A / 4) * (TILE_B / 4);  // = consumer threads

// BP = number of bit planes when known at compile time (plane loops fully unrolled), 0 = run-time P.n_planes.
template <int TI, int TJ, int BP>
__global__ void __launch_bounds__((TILE_A / TI) * (TILE_B / TJ) + 32, CTAS_PER_SM) pair_kernel(const PairParams P) {
    constexpr int NTJ = TILE_B / TJ;
    constexpr int PAIR_THREADS = (TILE_A / TI) * NTJ;  // consumer threads
    constexpr int CONSUMER_WARPS = PAIR_THREADS / 32;
    static_assert(TJ == 4 && (TI == 4 || TI == 8), "shared loads below are written for these tiles");
    static_assert(PAIR_THREADS == CAND_PER_ROUND, "one candidate per consumer thread and round");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ uint2 cand[CAND_PER_ROUND];  // .x = la | lb << 8 (tile-local genes), .y = nup | up_idx << 16
    __shared__ int s_total;
    constexpr int CHUNK_CACHE = 192;         // chunk descriptors kept in shared memory (global beyond that)
    __shared__ Chunk s_chunks[CHUNK_CACHE];

    const int tid = threadIdx.x;
    const int ti = tid / NTJ, tj = tid % NTJ;
    const int2 tile = P.tiles[blockIdx.x];
    const int B = BP > 0 ? BP : P.n_planes;
    const int blk_words = P.wc_max * B * 64;  // u32 per gene block per stage
    const uint32_t stage_words = (uint32_t)(A_BLOCKS + 1) * blk_words;
    uint32_t *smem = reinterpret_cast<uint32_t *>(smem_raw);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        fence_mbar_init();
        s_total = 0;
    }
    for (int c = tid; c < min(P.n_chunks, CHUNK_CACHE); c += blockDim.x) s_chunks[c] = P.chunks[c];
    __syncthreads();
    auto chunk_at = [&](int c) -> Chunk { return c < CHUNK_CACHE ? s_chunks[c] : P.chunks[c]; };

    const int64_t blk_stride = (int64_t)P.n_words * B * 64;  // u32 per gene block in HBM
    const uint32_t *gA0 = P.planes + (int64_t)(A_BLOCKS * tile.x) * blk_stride;
    const uint32_t *gA1 = gA0 + blk_stride;
    const uint32_t *gB = P.planes + (int64_t)tile.y * blk_stride;

    if (tid >= PAIR_THREADS) {  // ---- producer warp: one elected lane streams the chunk sequence once per sweep
        int q = 0;              // running chunk number over all sweeps (stage = q % STAGES)
        auto stream_sweep = [&]() {
            for (int c = 0; c < P.n_chunks; ++c, ++q) {
                const int s = q % STAGES;
                if (q >= STAGES) mbar_wait(&empty_bar[s], (uint32_t)(q / STAGES - 1) & 1u);  // consumers released it
                const Chunk ch = chunk_at(c);
                uint32_t *dst = smem + (size_t)s * stage_words;
                const uint32_t bytes = (uint32_t)ch.n_words * B * 64 * 4;
                const int64_t off = (int64_t)ch.w_begin * B * 64;
                mbar_arrive_expect_tx(&full_bar[s], (A_BLOCKS + 1) * bytes);
                bulk_g2s(dst, gA0 + off, bytes, &full_bar[s]);
                if (A_BLOCKS == 2) bulk_g2s(dst + blk_words, gA1 + off, bytes, &full_bar[s]);
                bulk_g2s(dst + A_BLOCKS * blk_words, gB + off, bytes, &full_bar[s]);
            }
        };
        if (tid == PAIR_THREADS) stream_sweep();
        __syncthreads();  // (A) sweep 1 is consumed and the candidates are counted
        const int rounds = (s_total + CAND_PER_ROUND - 1) / CAND_PER_ROUND;
        if (tid == PAIR_THREADS)
            for (int r = 0; r < rounds; ++r) stream_sweep();
        return;
    }

    // =============================== sweep 1: num_pos of every pair of the tile
    uint32_t acc[TI][TJ], st[TI][TJ], ncd[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0, st[i][j] = 0, ncd[i][j] = 0;

    const int la = ti * TI;  // first local a-gene (0..127), never straddles a 64-gene block
    const uint32_t a_off = (uint32_t)(la >> 6) * blk_words + (la & 63);
    const uint32_t b_off = (uint32_t)A_BLOCKS * blk_words + tj * TJ;
    int q = 0;

    for (int c = 0; c < P.n_chunks; ++c, ++q) {
        const int s = q % STAGES;
        const Chunk ch = chunk_at(c);
        mbar_wait(&full_bar[s], (uint32_t)(q / STAGES) & 1u);
        const uint32_t *sa = smem + (size_t)s * stage_words + a_off;
        const uint32_t *sb = smem + (size_t)s * stage_words + b_off;
        if (BP > 0) {
            // compile-time plane count: software-pipelined -- the operands of plane p+1 (or of plane 0 of the next
            // word) are requested before the 16 LOP3 of plane p issue, so the LDS latency hides behind them
            uint32_t a[TI], b[TJ];
            load_words<TI>(sa, a);
            load_words<TJ>(sb, b);
            for (int w = 0; w < ch.n_words; ++w) {
                uint32_t gt[TI][TJ];
#pragma unroll
                for (int p = 0; p < (BP > 0 ? BP : 1); ++p) {
                    uint32_t an[TI], bn[TJ];
                    const int nxt = (p + 1 < BP) ? (p + 1) * 64 : BP * 64;  // next plane, or plane 0 of the next word
                    if (p + 1 < BP || w + 1 < ch.n_words) {
                        load_words<TI>(sa + nxt, an);
                        load_words<TJ>(sb + nxt, bn);
                    }
#pragma unroll
                    for (int i = 0; i < TI; ++i)
#pragma unroll
                        for (int j = 0; j < TJ; ++j)
                            gt[i][j] = (p == 0) ? (a[i] & ~b[j]) : lop3_gt(a[i], b[j], gt[i][j]);
#pragma unroll
                    for (int i = 0; i < TI; ++i) a[i] = an[i];
#pragma unroll
                    for (int j = 0; j < TJ; ++j) b[j] = bn[j];
                }
#pragma unroll
                for (int i = 0; i < TI; ++i)
#pragma unroll
                    for (int j = 0; j < TJ; ++j) acc[i][j] = fma_pipe_mad((uint32_t)__popc(gt[i][j]), 1u, acc[i][j]);
                sa += BP * 64;
                sb += BP * 64;
            }
        } else {
            for (int w = 0; w < ch.n_words; ++w) {
                uint32_t gt[TI][TJ];
                {
                    uint32_t a[TI], b[TJ];
                    load_words<TI>(sa, a);
                    load_words<TJ>(sb, b);
#pragma unroll
                    for (int i = 0; i < TI; ++i)
#pragma unroll
                        for (int j = 0; j < TJ; ++j) gt[i][j] = a[i] & ~b[j];
                }
#pragma unroll 4
                for (int p = 1; p < B; ++p) {
                    uint32_t a[TI], b[TJ];
                    load_words<TI>(sa + p * 64, a);
                    load_words<TJ>(sb + p * 64, b);
#pragma unroll
                    for (int i = 0; i < TI; ++i)
#pragma unroll
                        for (int j = 0; j < TJ; ++j) gt[i][j] = lop3_gt(a[i], b[j], gt[i][j]);
                }
#pragma unroll
                for (int i = 0; i < TI; ++i)
#pragma unroll
                    for (int j = 0; j < TJ; ++j) acc[i][j] = fma_pipe_mad((uint32_t)__popc(gt[i][j]), 1u, acc[i][j]);
                sa += B * 64;
                sb += B * 64;
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);  // this warp is done reading stage s
        if (ch.class_end >= 0) {  // CTA-uniform: `num_pos[i] >= thresholds[i]` of sandbag.py:185-187, last class wins
            const uint32_t k = (uint32_t)ch.class_end;
            const int64_t thr = (int64_t)__ldg(P.thr + k), mpd = (int64_t)__ldg(P.max_pos_dn + k);
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j) {
                    if ((int64_t)acc[i][j] >= thr) st[i][j] = ((st[i][j] & 0xffffu) + 1u) | (k << 16);
                    if ((int64_t)acc[i][j] <= mpd) ncd[i][j] += 1u;
                    acc[i][j] = 0;
                }
        }
    }

    // =============================== candidates: pairs whose decision can still emit (nup == 1 or nup == C-1)
    const int ga0 = tile.x * TILE_A + la, gb0 = tile.y * TILE_B + tj * TJ;
    const uint32_t C = (uint32_t)P.n_classes;
    const int lane = tid & 31;
    uint32_t cmask = 0;
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const uint32_t nup = st[i][j] & 0xffffu;
            const bool can = (nup == 1u) ? (ncd[i][j] >= C - 1u) : (nup == C - 1u && ncd[i][j] >= 1u);
            if (ga0 + i < gb0 + j && gb0 + j < P.n_genes && can) cmask |= 1u << (i * TJ + j);
        }
    const int my_first = cmask ? atomicAdd(&s_total, __popc(cmask)) : 0;  // tickets of this thread's candidates
    __syncthreads();  // (A)
    const int total = s_total;
    const int rounds = (total + CAND_PER_ROUND - 1) / CAND_PER_ROUND;

    // =============================== sweep 2: num_neg of the candidates, CAND_PER_ROUND per replay of the stream
    for (int r = 0; r < rounds; ++r) {
        {
            int t = my_first;
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j)
                    if (cmask & (1u << (i * TJ + j))) {
                        if (t / CAND_PER_ROUND == r)
                            cand[t % CAND_PER_ROUND] = make_uint2((uint32_t)(la + i) | ((uint32_t)(tj * TJ + j) << 8), st[i][j]);
                        ++t;
                    }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(PAIR_THREADS) : "memory");  // consumers only: the list is complete
        const bool have = tid < min(CAND_PER_ROUND, total - r * CAND_PER_ROUND);
        const uint2 me = have ? cand[tid] : make_uint2(0u, 0u);
        const uint32_t ca = me.x & 0xffu, cb = me.x >> 8;
        const uint32_t ca_off = (ca >> 6) * blk_words + (ca & 63), cb_off = (uint32_t)A_BLOCKS * blk_words + cb;
        uint32_t neg = 0, ndn = 0, dn_idx = 0;
        for (int c = 0; c < P.n_chunks; ++c, ++q) {
            const int s = q % STAGES;
            const Chunk ch = chunk_at(c);
            mbar_wait(&full_bar[s], (uint32_t)(q / STAGES) & 1u);
            if (have) {
                const uint32_t *sa = smem + (size_t)s * stage_words + ca_off;
                const uint32_t *sb = smem + (size_t)s * stage_words + cb_off;
                for (int w = 0; w < ch.n_words; ++w) {
                    uint32_t lt = 0;
                    for (int p = 0; p < B; ++p) lt = lop3_lt(sa[p * 64], sb[p * 64], lt);
                    neg += __popc(lt);
                    sa += B * 64;
                    sb += B * 64;
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]);
            if (ch.class_end >= 0) {  // `num_neg[i] >= thresholds[i]` of sandbag.py:188-190
                if ((int64_t)neg >= (int64_t)__ldg(P.thr + ch.class_end)) {
                    ++ndn;
                    dn_idx = (uint32_t)ch.class_end;
                }
                neg = 0;
            }
        }
        // decision of sandbag.py:192-197 + warp-aggregated compaction
        const uint32_t nup = me.y & 0xffffu, up_idx = me.y >> 16;
        const unsigned long long ga = (unsigned long long)(tile.x * TILE_A + (int)ca);
        const unsigned long long gb = (unsigned long long)(tile.y * TILE_B + (int)cb);
        unsigned long long key = 0;
        bool emit = false;
        if (have) {
            if (nup == 1u) {
                if (ndn == C - 1u) {
                    emit = true;
                    key = ((ga * P.n_genes + gb) << 16) | up_idx;
                }
            } else if (ndn == 1u) {
                if (nup == C - 1u) {
                    emit = true;
                    key = ((gb * P.n_genes + ga) << 16) | dn_idx;
                }
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, emit);
        if (m) {
            unsigned long long base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(P.out_count, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
            if (emit && pos < P.out_cap) P.out[pos] = key;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(PAIR_THREADS) : "memory");  // cand[] is rewritten by the next round
    }
}

// per-class counters of selected pairs straight from the planes (one warp per probe pair)
__global__ void probe_kernel(const uint32_t *__restrict__ planes, const Chunk *__restrict__ chunks, int n_chunks,
                             int n_words, int n_planes, int n_classes, const int64_t *__restrict__ pairs,
                             int64_t n_probe, int32_t *__restrict__ pos_out, int32_t *__restrict__ neg_out) {
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= n_probe) return;
    const int g1 = (int)pairs[2 * q], g2 = (int)pairs[2 * q + 1];
    const int64_t bs = (int64_t)n_words * n_planes * 64;
    const uint32_t *pa = planes + (int64_t)(g1 >> 6) * bs + (g1 & 63);
    const uint32_t *pb = planes + (int64_t)(g2 >> 6) * bs + (g2 & 63);
    int pos = 0, neg = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const Chunk ch = chunks[c];
        for (int w = ch.w_begin + lane; w < ch.w_begin + ch.n_words; w += 32) {
            uint32_t gt = 0, lt = 0;
            for (int p = 0; p < n_planes; ++p) {
                const uint32_t a = pa[((int64_t)w * n_planes + p) * 64], b = pb[((int64_t)w * n_planes + p) * 64];
                gt = lop3_gt(a, b, gt);
                lt = lop3_lt(a, b, lt);
            }
            pos += __popc(gt);
            neg += __popc(lt);
        }
        if (ch.class_end >= 0) {
            for (int o = 16; o > 0; o >>= 1) {
                pos += __shfl_xor_sync(0xffffffffu, pos, o);
                neg += __shfl_xor_sync(0xffffffffu, neg, o);
            }
            if (lane == 0) {
                pos_out[q * n_classes + ch.class_end] = pos;
                neg_out[q * n_classes + ch.class_end] = neg;
            }
            pos = neg = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// LSD radix sort of the compacted u64 keys (8-bit digits, 3 kernels per pass)
// ------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ITEMS = 16;                        // per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 8192 keys per CTA

__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const unsigned long long *__restrict__ in, int64_t n, int shift, uint32_t *__restrict__ hist,
                 int n_blocks) {
    __shared__ uint32_t h[256];
    for (int i = threadIdx.x; i < 256; i += SORT_THREADS) h[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    for (int it = 0; it < SORT_ITEMS; ++it) {
        const int64_t i = base + it * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(in[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += SORT_THREADS) hist[(int64_t)i * n_blocks + blockIdx.x] = h[i];
}

// exclusive scan of hist[256 * n_blocks] in place (single CTA)
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t *__restrict__ hist, int64_t total) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < total; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = (i < total) ? hist[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t s = wsum[lane], inc2 = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc2, o);
                if (lane >= o) inc2 += t;
            }
            wsum[lane] = inc2 - s;
        }
        __syncthreads();
        const uint32_t c = carry;
        if (i < total) hist[i] = c + wsum[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c + wsum[31] + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const unsigned long long *__restrict__ in, unsigned long long *__restrict__ out, int64_t n,
                    int shift, const uint32_t *__restrict__ hist, int n_blocks) {
    __shared__ uint32_t whist[SORT_WARPS][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < SORT_WARPS * 256; i += SORT_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    // warp-contiguous ranges keep the scatter stable
    const int64_t lo = (int64_t)blockIdx.x * SORT_TILE + (int64_t)warp * (SORT_TILE / SORT_WARPS);
    const int64_t hi = min(lo + SORT_TILE / SORT_WARPS, n);
    for (int64_t ib = lo; ib < hi; ib += 32) {
        const int64_t i = ib + lane;
        const bool valid = i < hi;
        const uint32_t d = valid ? ((uint32_t)(in[i] >> shift) & 255u) : (256u + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs(peers) - 1) whist[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    if (tid < 256) {
        uint32_t run = hist[(int64_t)tid * n_blocks + blockIdx.x];
        for (int w = 0; w < SORT_WARPS; ++w) {
            const uint32_t c = whist[w][tid];
            whist[w][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int64_t ib = lo; ib < hi; ib += 32) {
        const int64_t i = ib + lane;
        const bool valid = i < hi;
        const unsigned long long k = valid ? in[i] : 0ull;
        const uint32_t d = valid ? ((uint32_t)(k >> shift) & 255u) : (256u + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t base = valid ? whist[warp][d] : 0u;
        if (valid) out[base + __popc(peers & ((1u << lane) - 1u))] = k;
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) whist[warp][d] = base + __popc(peers);
        __syncwarp();
    }
}

int radix_sort_u64(pp_ctx *ctx, unsigned long long *d_a, unsigned long long *d_b, int64_t n, int lo_bit,
                   int hi_bit, unsigned long long **sorted) {
    *sorted = d_a;
    if (n <= 1) return PP_OK;
    const int n_blocks = (int)((n + SORT_TILE - 1) / SORT_TILE);
    DevBuf hist;
    PP_CUDA(hist.alloc(sizeof(uint32_t) * 256 * (size_t)n_blocks, ctx->stream));
    unsigned long long *in = d_a, *out = d_b;
    for (int shift = lo_bit; shift < hi_bit; shift += 8) {
        sort_hist_kernel<<<n_blocks, SORT_THREADS, 0, ctx->stream>>>(in, n, shift, hist.as<uint32_t>(), n_blocks);
        PP_CHECK_LAUNCH();
        sort_scan_kernel<<<1, 1024, 0, ctx->stream>>>(hist.as<uint32_t>(), (int64_t)256 * n_blocks);
        PP_CHECK_LAUNCH();
        sort_scatter_kernel<<<n_blocks, SORT_THREADS, 0, ctx->stream>>>(in, out, n, shift, hist.as<uint32_t>(),
                                                                         n_blocks);
        PP_CHECK_LAUNCH();
        std::swap(in, out);
    }
    *sorted = in;
    return PP_OK;
}

int bit_length64(uint64_t v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

}  // namespace

// Shared body of pp_sandbag (dense x, csr == NULL) and pp_sandbag_csr (csr != NULL, x / ld unused).
static int sandbag_impl(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, const pp_csr *csr, int64_t n_rows,
                        int64_t n_cols, int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int64_t *class_sizes,
                        int32_t n_classes, const int32_t *gene_cols, int64_t n_genes, const int64_t *thresholds,
                        int32_t shard, int32_t n_shards, uint64_t **out_keys, int32_t **out_class, int64_t *out_n,
                        const int64_t *probe_pairs, int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg,
                        int32_t drop_unexpressed, int32_t **out_kept_genes, int64_t *out_n_kept) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    if ((!x && !csr) || !cell_rows || !class_sizes || !gene_cols || !thresholds || !out_keys || !out_class || !out_n)
        return pp_fail(ctx, PP_EINVAL, "pp_sandbag: NULL argument");
    if (x_dtype != PP_F32 && x_dtype != PP_F64) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: x_dtype must be PP_F32/PP_F64");
    if (n_rows <= 0 || n_cols <= 0 || (!csr && ld < n_cols)) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: bad matrix shape");
    if (csr) {
        const int vrc = pp_csr_validate(ctx, csr, n_rows, n_cols, "pp_sandbag_csr");
        if (vrc) return vrc;
    }
    if (n_classes < 1) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: n_classes < 1");
    if (n_classes > 65534) return pp_fail(ctx, PP_ELIMIT, "pp_sandbag: more than 65534 classes (%d)", n_classes);
    if (n_shards < 1 || shard < 0 || shard >= n_shards) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: bad shard %d/%d", shard, n_shards);
    if (n_genes < 0 || n_cells < 0) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: negative size");
    *out_keys = nullptr;
    *out_class = nullptr;
    *out_n = 0;
    ctx->tm = pp_timings();
    ctx->launches = 0;
    int64_t total = 0;
    for (int k = 0; k < n_classes; ++k) {
        if (class_sizes[k] <= 0) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: class %d is empty", k);
        total += class_sizes[k];
    }
    if (total != n_cells) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: class_sizes do not sum to n_cells");
    for (int64_t i = 0; i < n_cells; ++i)
        if (cell_rows[i] < 0 || cell_rows[i] >= n_rows) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: cell_rows[%lld] out of range", (long long)i);
    for (int64_t g = 0; g < n_genes; ++g)
        if (gene_cols[g] < 0 || gene_cols[g] >= n_cols) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: gene_cols[%lld] out of range", (long long)g);
    if (drop_unexpressed && (!out_kept_genes || !out_n_kept))
        return pp_fail(ctx, PP_EINVAL, "pp_sandbag: drop_unexpressed needs out_kept_genes / out_n_kept");
    if (out_kept_genes) *out_kept_genes = nullptr;
    if (out_n_kept) *out_n_kept = 0;

    PP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PP_CUDA(cudaEventRecord(ctx->ev[0], st));
    DevBuf d_slots, d_gene, d_flags, d_chunks, d_tiles, d_thr, d_cnt;
    void *s_x = nullptr, *s_rank = nullptr, *s_planes = nullptr, *s_out = nullptr, *s_out2 = nullptr;  // ctx slabs
    const size_t esz = x_dtype == PP_F64 ? 8 : 4;
    const void *dx = x;
    DevCsr dcsr;
    int rc = PP_OK;
    if (csr) {
        rc = pp_csr_to_device(ctx, csr, 0, n_rows, &dcsr);
        if (rc) return rc;
    } else if (!x_is_device) {
        if (!(s_x = pp_slab(ctx, SLAB_X, (size_t)n_rows * ld * esz))) return PP_ENOMEM;
        PP_CUDA(cudaMemcpyAsync(s_x, x, (size_t)n_rows * ld * esz, cudaMemcpyHostToDevice, st));
        ctx->tm.h2d_bytes = (int64_t)((size_t)n_rows * ld * esz);
        dx = s_x;
    }
    PP_CUDA(cudaEventRecord(ctx->ev[1], st));

    // ---- optional: drop genes that are zero in ALL rows of x (utils.py:368, before sample filtering)
    std::vector<int32_t> kept;
    if (drop_unexpressed) {
        DevBuf d_nz;
        PP_CUDA(d_nz.alloc((size_t)n_cols, st));
        PP_CUDA(cudaMemsetAsync(d_nz.p, 0, (size_t)n_cols, st));
        dim3 grid((unsigned)((n_cols + 255) / 256), (unsigned)std::min<int64_t>((n_rows + 63) / 64, 65535));
        if (csr) {
            rc = pp_csr_col_nonzero(ctx, dcsr, d_nz.as<uint8_t>());
            if (rc) return rc;
        } else if (x_dtype == PP_F64) {
            col_nonzero_kernel<double><<<grid, 256, 0, st>>>((const double *)dx, ld, n_rows, n_cols, d_nz.as<uint8_t>());
            PP_CHECK_LAUNCH();
        } else {
            col_nonzero_kernel<float><<<grid, 256, 0, st>>>((const float *)dx, ld, n_rows, n_cols, d_nz.as<uint8_t>());
            PP_CHECK_LAUNCH();
        }
        std::vector<uint8_t> nz((size_t)n_cols);
        PP_CUDA(cudaMemcpyAsync(nz.data(), d_nz.p, (size_t)n_cols, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
        for (int64_t g = 0; g < n_genes; ++g)
            if (nz[gene_cols[g]]) kept.push_back(gene_cols[g]);
        gene_cols = kept.data();
        n_genes = (int64_t)kept.size();
        int32_t *hk = (int32_t *)malloc(sizeof(int32_t) * (size_t)std::max<int64_t>(n_genes, 1));
        if (!hk) return pp_fail(ctx, PP_ENOMEM, "pp_sandbag: host allocation failed");
        std::copy(kept.begin(), kept.end(), hk);
        *out_kept_genes = hk;
        *out_n_kept = n_genes;
    }
    if (n_genes < 2 || n_cells == 0) return PP_OK;  // no pairs
    if (n_genes > 65535) return pp_fail(ctx, PP_ELIMIT, "pp_sandbag: more than 65535 kept genes");

    // ---- host-side layout tables: padded slots, chunks, tiles
    std::vector<int32_t> slots;
    std::vector<int32_t> class_w0(n_classes + 1, 0);
    {
        int64_t src = 0;
        for (int k = 0; k < n_classes; ++k) {
            const int64_t nk = class_sizes[k], wk = (nk + 31) / 32;
            class_w0[k + 1] = class_w0[k] + (int32_t)wk;
            for (int64_t i = 0; i < wk * 32; ++i) slots.push_back(i < nk ? cell_rows[src + i] : -1);
            src += nk;
        }
    }
    const int W = class_w0[n_classes];
    const int64_t n_slots = (int64_t)W * 32;
    const int NB = (int)((n_genes + 63) / 64);
    const int NB_alloc = (NB + 1) & ~1;

    PP_CUDA(d_slots.alloc(sizeof(int32_t) * n_slots, st));
    PP_CUDA(cudaMemcpyAsync(d_slots.p, slots.data(), sizeof(int32_t) * n_slots, cudaMemcpyHostToDevice, st));
    PP_CUDA(d_gene.alloc(sizeof(int32_t) * n_genes, st));
    PP_CUDA(cudaMemcpyAsync(d_gene.p, gene_cols, sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice, st));
    if (!(s_rank = pp_slab(ctx, SLAB_RANK, sizeof(uint16_t) * n_slots * n_genes))) return PP_ENOMEM;
    PP_CUDA(d_flags.alloc(sizeof(int32_t) * 4, st));
    PP_CUDA(cudaMemsetAsync(d_flags.p, 0, sizeof(int32_t) * 4, st));

    if (csr) {
        // densify the kept genes of a run of slots, rank them, next run (<= 256 MB of dense values at a time)
        std::vector<int32_t> colmap((size_t)n_cols, -1), ident((size_t)n_genes);
        for (int64_t g = 0; g < n_genes; ++g) {
            colmap[gene_cols[g]] = (int32_t)g;
            ident[g] = (int32_t)g;
        }
        int64_t chunk = std::max<int64_t>(32, (((int64_t)256 << 20) / (n_genes * (int64_t)esz)) / 32 * 32);
        chunk = std::min(chunk, n_slots);
        std::vector<int32_t> loc((size_t)n_slots);
        for (int64_t sl = 0; sl < n_slots; ++sl) loc[sl] = slots[sl] < 0 ? -1 : (int32_t)(sl % chunk);
        DevBuf d_colmap, d_ident, d_loc, d_dense;
        PP_CUDA(d_colmap.alloc(sizeof(int32_t) * n_cols, st));
        PP_CUDA(cudaMemcpyAsync(d_colmap.p, colmap.data(), sizeof(int32_t) * n_cols, cudaMemcpyHostToDevice, st));
        PP_CUDA(d_ident.alloc(sizeof(int32_t) * n_genes, st));
        PP_CUDA(cudaMemcpyAsync(d_ident.p, ident.data(), sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice, st));
        PP_CUDA(d_loc.alloc(sizeof(int32_t) * n_slots, st));
        PP_CUDA(cudaMemcpyAsync(d_loc.p, loc.data(), sizeof(int32_t) * n_slots, cudaMemcpyHostToDevice, st));
        PP_CUDA(d_dense.alloc((size_t)chunk * n_genes * esz, st));
        for (int64_t s0 = 0; s0 < n_slots; s0 += chunk) {
            const int64_t nl = std::min(chunk, n_slots - s0);
            PP_CUDA(cudaMemsetAsync(d_dense.p, 0, (size_t)nl * n_genes * esz, st));
            rc = pp_csr_densify(ctx, dcsr, d_slots.as<int32_t>() + s0, nl, d_colmap.as<int32_t>(), d_dense.p, n_genes);
            if (rc) return rc;
            rc = pp_rank_launch(ctx, d_dense.p, x_dtype, n_genes, d_loc.as<int32_t>() + s0, nl, d_ident.as<int32_t>(),
                                n_genes, (uint16_t *)s_rank + s0 * n_genes, n_genes, d_flags.as<int32_t>());
            if (rc) return rc;
        }
        PP_CUDA(cudaStreamSynchronize(st));  // colmap / ident / loc are locals
    } else {
        rc = pp_rank_launch(ctx, dx, x_dtype, ld, d_slots.as<int32_t>(), n_slots, d_gene.as<int32_t>(), n_genes,
                            (uint16_t *)s_rank, n_genes, d_flags.as<int32_t>());
        if (rc) return rc;
    }
    int32_t flags[4];
    PP_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaStreamSynchronize(st));
    if (flags[1]) return pp_fail(ctx, PP_ENAN, "pp_sandbag: NaN in the expression matrix");
    const int B = std::max(1, bit_length64((uint64_t)flags[0]));
    ctx->tm.n_planes = B;

    // ---- bit-slice
    const size_t plane_words = (size_t)NB_alloc * W * B * 64;
    if (!(s_planes = pp_slab(ctx, SLAB_PLANES, plane_words * 4))) return PP_ENOMEM;
    PP_CUDA(cudaMemsetAsync(s_planes, 0, plane_words * 4, st));
    {
        dim3 grid(W, (unsigned)((n_genes + 255) / 256));
        bitslice_kernel<<<grid, 256, 0, st>>>((uint16_t *)s_rank, n_genes, (int)n_genes, W, B, (uint32_t *)s_planes);
        PP_CHECK_LAUNCH();
    }
    PP_CUDA(cudaEventRecord(ctx->ev[2], st));

    // ---- chunks (runs of <= wc_max words inside one class)
    const int smem_budget = std::min(ctx->smem_optin, 227 * 1024) - 8192;  // static: barriers + candidate list
    int wc_max = (smem_budget / CTAS_PER_SM - (CTAS_PER_SM > 1 ? 6144 : 0)) / (STAGES * (A_BLOCKS + 1) * B * 64 * 4);
    wc_max = std::max(1, std::min(wc_max, 16));
    const size_t dyn_smem = (size_t)STAGES * (A_BLOCKS + 1) * wc_max * B * 64 * 4;
    std::vector<Chunk> chunks;
    for (int k = 0; k < n_classes; ++k)
        for (int w = class_w0[k]; w < class_w0[k + 1]; w += wc_max) {
            Chunk c;
            c.w_begin = w;
            c.n_words = std::min(wc_max, class_w0[k + 1] - w);
            c.class_end = (w + c.n_words == class_w0[k + 1]) ? k : -1;
            c.pad = 0;
            chunks.push_back(c);
        }
    // ---- tiles of this shard: (ta, tb) with tb >= A_BLOCKS * ta (only those contain pairs g1 < g2)
    const int TA = (int)((n_genes + TILE_A - 1) / TILE_A), TB = NB;
    std::vector<int2> tiles;
    int64_t t_index = 0, n_pairs_shard = 0;
    // Super-tile order: the ~n_sm tiles that run at the same time come from one SUP_A x SUP_B patch of the tile
    // grid, so they share their 2*SUP_A + SUP_B operand blocks through L2 instead of streaming each operand
    // from HBM once per tile (ncu: 11.7 GB of DRAM reads per cfg3 launch in row-major tile order).
    constexpr int SUP_A = 8 * (128 / TILE_A), SUP_B = 18;
    for (int sa = 0; sa < TA; sa += SUP_A)
        for (int sb = A_BLOCKS * sa; sb < TB; sb += SUP_B)
            for (int ta = sa; ta < std::min(sa + SUP_A, TA); ++ta)
                for (int tb = std::max(sb, A_BLOCKS * ta); tb < std::min(sb + SUP_B, TB); ++tb, ++t_index) {
                    if (t_index % n_shards != shard) continue;
                    tiles.push_back(make_int2(ta, tb));
                    const int64_t a0 = (int64_t)ta * TILE_A, a1 = std::min<int64_t>(a0 + TILE_A, n_genes);
                    const int64_t b0 = (int64_t)tb * TILE_B, b1 = std::min<int64_t>(b0 + TILE_B, n_genes);
                    for (int64_t a = a0; a < a1; ++a) n_pairs_shard += std::max<int64_t>(0, b1 - std::max(b0, a + 1));
                }
    ctx->tm.n_units = n_pairs_shard;
    std::vector<int32_t> thr32(n_classes), mpd32(n_classes);
    for (int k = 0; k < n_classes; ++k) {
        thr32[k] = (int32_t)std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, thresholds[k]));
        mpd32[k] = (int32_t)std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, class_sizes[k] - thresholds[k]));
    }

    PP_CUDA(d_chunks.alloc(sizeof(Chunk) * chunks.size(), st));
    PP_CUDA(cudaMemcpyAsync(d_chunks.p, chunks.data(), sizeof(Chunk) * chunks.size(), cudaMemcpyHostToDevice, st));
    PP_CUDA(d_tiles.alloc(sizeof(int2) * tiles.size(), st));
    PP_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, st));
    PP_CUDA(d_thr.alloc(sizeof(int32_t) * n_classes, st));
    PP_CUDA(cudaMemcpyAsync(d_thr.p, thr32.data(), sizeof(int32_t) * n_classes, cudaMemcpyHostToDevice, st));
    DevBuf d_mpd;
    PP_CUDA(d_mpd.alloc(sizeof(int32_t) * n_classes, st));
    PP_CUDA(cudaMemcpyAsync(d_mpd.p, mpd32.data(), sizeof(int32_t) * n_classes, cudaMemcpyHostToDevice, st));
    PP_CUDA(d_cnt.alloc(sizeof(unsigned long long), st));

    unsigned long long cap = (unsigned long long)std::min<int64_t>(std::max<int64_t>(n_pairs_shard, 1),
                                                                     std::max<int64_t>(1 << 22, n_pairs_shard / 16));
    unsigned long long count = 0;
    // 4x4 register tiles, 512 threads (16 warps/SM): measured 75.0 ms on cfg3 vs 87.8 ms for 8x4 / 256 threads
    // plane count known at compile time for 1..16 (fully unrolled plane loops)
    typedef void (*PairFn)(const PairParams);
    static const PairFn pair_fns[17] = {
        pair_kernel<4, 4, 0>,  pair_kernel<4, 4, 1>,  pair_kernel<4, 4, 2>,  pair_kernel<4, 4, 3>,  pair_kernel<4, 4, 4>,
        pair_kernel<4, 4, 5>,  pair_kernel<4, 4, 6>,  pair_kernel<4, 4, 7>,  pair_kernel<4, 4, 8>,  pair_kernel<4, 4, 9>,
        pair_kernel<4, 4, 10>, pair_kernel<4, 4, 11>, pair_kernel<4, 4, 12>, pair_kernel<4, 4, 13>, pair_kernel<4, 4, 14>,
        pair_kernel<4, 4, 15>, pair_kernel<4, 4, 16>};
    const PairFn pair_fn = pair_fns[B <= 16 ? B : 0];
    PP_CUDA(cudaFuncSetAttribute(pair_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!(s_out = pp_slab(ctx, SLAB_OUT, sizeof(unsigned long long) * cap))) return PP_ENOMEM;
        PP_CUDA(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long), st));
        if (!tiles.empty()) {
            PairParams P;
            P.planes = (uint32_t *)s_planes;
            P.chunks = d_chunks.as<Chunk>();
            P.tiles = d_tiles.as<int2>();
            P.thr = d_thr.as<int32_t>();
            P.max_pos_dn = d_mpd.as<int32_t>();
            P.out = (unsigned long long *)s_out;
            P.out_count = d_cnt.as<unsigned long long>();
            P.out_cap = cap;
            P.n_genes = (int)n_genes;
            P.n_words = W;
            P.n_planes = B;
            P.n_chunks = (int)chunks.size();
            P.n_classes = n_classes;
            P.wc_max = wc_max;
            pair_fn<<<(unsigned)tiles.size(), CAND_PER_ROUND + 32, dyn_smem, st>>>(P);
            PP_CHECK_LAUNCH();
        }
        PP_CUDA(cudaEventRecord(ctx->ev[3], st));
        PP_CUDA(cudaMemcpyAsync(&count, d_cnt.p, sizeof(count), cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));

        if (count <= cap) break;
        if (attempt == 1) return pp_fail(ctx, PP_ECUDA, "pp_sandbag: compaction buffer overflow after resize");
        cap = count;  // exact size known now; evaluate once more
    }

    // ---- optional probe counters
    if (n_probe > 0 && probe_pairs && probe_pos && probe_neg) {
        for (int64_t q = 0; q < n_probe; ++q)
            if (probe_pairs[2 * q] < 0 || probe_pairs[2 * q + 1] >= n_genes || probe_pairs[2 * q] >= probe_pairs[2 * q + 1])
                return pp_fail(ctx, PP_EINVAL, "pp_sandbag: probe pair %lld must satisfy 0 <= g1 < g2 < n_genes", (long long)q);
        DevBuf d_pp, d_pos, d_neg;
        const size_t cb = sizeof(int32_t) * n_probe * n_classes;
        PP_CUDA(d_pp.alloc(sizeof(int64_t) * 2 * n_probe, st));
        PP_CUDA(d_pos.alloc(cb, st));
        PP_CUDA(d_neg.alloc(cb, st));
        PP_CUDA(cudaMemcpyAsync(d_pp.p, probe_pairs, sizeof(int64_t) * 2 * n_probe, cudaMemcpyHostToDevice, st));
        probe_kernel<<<(unsigned)((n_probe + 3) / 4), 128, 0, st>>>((uint32_t *)s_planes, d_chunks.as<Chunk>(),
                                                                   (int)chunks.size(), W, B, n_classes,
                                                                   d_pp.as<int64_t>(), n_probe, d_pos.as<int32_t>(),
                                                                   d_neg.as<int32_t>());
        PP_CHECK_LAUNCH();
        PP_CUDA(cudaMemcpyAsync(probe_pos, d_pos.p, cb, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaMemcpyAsync(probe_neg, d_neg.p, cb, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
    }

    // ---- sort into np.where order and return
    uint64_t *hk = nullptr;
    int32_t *hc = nullptr;
    if (count > 0) {
        if (!(s_out2 = pp_slab(ctx, SLAB_OUT2, sizeof(unsigned long long) * count))) return PP_ENOMEM;
        unsigned long long *sorted = nullptr;
        const int hi_bit = 16 + bit_length64((uint64_t)n_genes * (uint64_t)n_genes);
        rc = radix_sort_u64(ctx, (unsigned long long *)s_out, (unsigned long long *)s_out2, (int64_t)count, 16,
                            hi_bit, &sorted);
        if (rc) return rc;
        hk = (uint64_t *)malloc(sizeof(uint64_t) * count);
        hc = (int32_t *)malloc(sizeof(int32_t) * count);
        if (!hk || !hc) {
            free(hk);
            free(hc);
            return pp_fail(ctx, PP_ENOMEM, "pp_sandbag: host allocation of %llu markers failed", count);
        }
        // device -> pinned staging (a pageable destination would be staged by the driver at a fraction of the rate)
        cudaError_t e = cudaSuccess;
        if (ctx->pin_out_bytes < sizeof(uint64_t) * count) {
            if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
            ctx->pin_out = nullptr;
            ctx->pin_out_bytes = 0;
            const size_t want = std::max<size_t>(sizeof(uint64_t) * count * 2, (size_t)8 << 20);
            e = cudaHostAlloc(&ctx->pin_out, want, cudaHostAllocDefault);
            if (e == cudaSuccess) ctx->pin_out_bytes = want;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->pin_out, sorted, sizeof(uint64_t) * count, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            free(hk);
            free(hc);
            return pp_fail(ctx, PP_ECUDA, "pp_sandbag: result copy failed: %s", cudaGetErrorString(e));
        }
        const uint64_t *packed = (const uint64_t *)ctx->pin_out;
        for (unsigned long long i = 0; i < count; ++i) {
            hc[i] = (int32_t)(packed[i] & 0xffffu);
            hk[i] = packed[i] >> 16;
        }
    }
    PP_CUDA(cudaEventRecord(ctx->ev[4], st));
    PP_CUDA(cudaEventSynchronize(ctx->ev[4]));
    cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->tm.prep_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->tm.main_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->tm.post_ms, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[0], ctx->ev[4]);
    ctx->tm.n_launches = ctx->launches;
    *out_keys = hk;
    *out_class = hc;
    *out_n = (int64_t)count;
    return PP_OK;
}

PP_API int pp_sandbag(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                      int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int64_t *class_sizes,
                      int32_t n_classes, const int32_t *gene_cols, int64_t n_genes, const int64_t *thresholds,
                      int32_t shard, int32_t n_shards, uint64_t **out_keys, int32_t **out_class, int64_t *out_n,
                      const int64_t *probe_pairs, int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg,
                      int32_t drop_unexpressed, int32_t **out_kept_genes, int64_t *out_n_kept) {
    if (ctx && !x) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: NULL argument");
    return sandbag_impl(ctx, x, x_dtype, x_is_device, nullptr, n_rows, n_cols, ld, cell_rows, n_cells, class_sizes,
                        n_classes, gene_cols, n_genes, thresholds, shard, n_shards, out_keys, out_class, out_n,
                        probe_pairs, n_probe, probe_pos, probe_neg, drop_unexpressed, out_kept_genes, out_n_kept);
}

PP_API int pp_sandbag_csr(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, const int32_t *cell_rows,
                          int64_t n_cells, const int64_t *class_sizes, int32_t n_classes, const int32_t *gene_cols,
                          int64_t n_genes, const int64_t *thresholds, int32_t shard, int32_t n_shards,
                          uint64_t **out_keys, int32_t **out_class, int64_t *out_n, const int64_t *probe_pairs,
                          int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg, int32_t drop_unexpressed,
                          int32_t **out_kept_genes, int64_t *out_n_kept) {
    if (ctx && !x) return pp_fail(ctx, PP_EINVAL, "pp_sandbag_csr: NULL argument");
    return sandbag_impl(ctx, nullptr, x ? x->data_dtype : PP_F32, x ? x->is_device : 0, x, n_rows, n_cols, n_cols,
                        cell_rows, n_cells, class_sizes, n_classes, gene_cols, n_genes, thresholds, shard, n_shards,
                        out_keys, out_class, out_n, probe_pairs, n_probe, probe_pos, probe_neg, drop_unexpressed,
                        out_kept_genes, out_n_kept);
}
