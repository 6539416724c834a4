// pp_sandbag.cu -- B200 all-gene-pairs marker kernel (replaces check_pairs,
// pypairs/tools/sandbag.py:160-199, and the np.where scan at :129).
//
// Data layout in HBM
//   rank   u16 [32*W][G]     per-cell dense ranks, cells class-sorted, every class padded
//                            to a multiple of 32 cells with all-tie cells (pp_rank.cu)
//   planes u32, per gene block nb (64 gene positions; positions sorted by the planes a gene's
//                            ranks need, widest first): [2*Bn - 1][W][64] -- Bn low planes
//                            (bit p of the rank) and the suffix planes "rank >= 2^j"; bit c of
//                            cell word w belongs to cell 32*w + c (cells of ONE class).  A
//                            (block, plane, run of words) is one contiguous byte range -> one
//                            1-D TMA bulk copy.  With several ranks (pp_sandbag_dist) every rank owns one such
//                            region holding ITS cells (class-sorted and padded per rank), regions are exchanged
//                            with one all-gather, and a chunk names (owner rank, first word, words).
// Pair kernel (pair_kernel): CTA tile = 128 a-positions x 64 b-positions, every thread owns a
// 4x4 register tile of gene pairs and walks all cell words.  For one word the predicate
// x[a] > x[b] (x[a] < x[b] in the second sweep) of sandbag.py:179-182 is evaluated for 32
// cells at once with the LSB->MSB recurrences
//        gt' = (a & ~b) | (~(a ^ b) & gt)        lt' = (~a & b) | (~(a ^ b) & lt)
// -- one LOP3 per plane per predicate -- and accumulated per class with POPC.  Ties set
// neither bit, padding cells tie everywhere, exactly the reference's "add to neither".
// Thresholds (sandbag.py:184-190) are applied in-register at every class boundary and the
// if/elif decision (:192-197) at the end; survivors are compacted with warp-aggregated
// atomics and radix-sorted into the row-major order np.where would produce.
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

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
// orbits[g] = OR over all cell slots of rank[slot][g]: bit_length(orbits[g]) = planes gene g needs
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
col_or_kernel(const uint16_t *__restrict__ rank, int64_t n_slots, int n_genes, uint32_t *__restrict__ orbits) {
    const int g = blockIdx.x * 256 + threadIdx.x;
    if (g >= n_genes) return;
    uint32_t v = 0;
    for (int64_t s = blockIdx.y; s < n_slots; s += gridDim.y) v |= rank[s * n_genes + g];
    if (v) atomicOr(orbits + g, v);
}

// ------------------------------------------------------------------------------------------
// bit-slice: rank[32W][G] -> per gene block nb (64 gene POSITIONS, genes sorted by the planes they need):
//   planes[blk_off[nb] + (q * W + w) * 64 + (position & 63)]
//     q <  Bn            low plane q   (bit q of the rank)
//     q >= Bn            suffix plane S_j, j = q - Bn + 1 (1 <= j < Bn): OR of the planes >= j, i.e. "rank >= 2^j"
// ------------------------------------------------------------------------------------------
// n_words = cell words of THIS rank's cells (rows of `rank`), stride_words = words per plane of a rank's region (the
// same on every rank, >= n_words; the words beyond n_words stay zero = all-tie cells).
__global__ void __launch_bounds__(256)
bitslice_kernel(const uint16_t *__restrict__ rank, int64_t rank_ld, int n_genes, int n_words, int stride_words,
                const int32_t *__restrict__ pos_gene, const int64_t *__restrict__ blk_off,
                const int32_t *__restrict__ blk_bits, uint32_t *__restrict__ planes) {
    const int lane = threadIdx.x & 31;
    const int pos = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 32 + lane;  // this lane's gene position
    if (pos - lane >= n_genes) return;
    const int g = pos < n_genes ? __ldg(pos_gene + pos) : 0;
    const int nb = pos >> 6, bn = __ldg(blk_bits + nb);
    const int64_t ps = (int64_t)stride_words * 64;  // words per plane
    for (int w = blockIdx.y; w < n_words; w += gridDim.y) {
        uint32_t r[32];
        const uint16_t *src = rank + (int64_t)w * 32 * rank_ld + g;
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = (pos < n_genes) ? (uint32_t)src[(int64_t)c * rank_ld] : 0u;
        uint32_t *dst = planes + __ldg(blk_off + nb) + (int64_t)w * 64 + (pos & 63);
        uint32_t suffix = 0;
        for (int p = bn - 1; p >= 0; --p) {
            uint32_t word = 0;
#pragma unroll
            for (int c = 0; c < 32; ++c) word |= ((r[c] >> p) & 1u) << c;
            dst[p * ps] = word;
            suffix |= word;
            if (p >= 1) dst[(bn + p - 1) * ps] = suffix;
        }
    }
}

// ------------------------------------------------------------------------------------------
// pair kernel
// ------------------------------------------------------------------------------------------
constexpr int TILE_A = 128;  // a-gene positions per CTA = 2 gene blocks
constexpr int A_BLOCKS = TILE_A / 64;
constexpr int TILE_B = 64;   // b-gene positions per CTA = 1 gene block
constexpr int STAGES = 3;
constexpr int MAX_PLANES = 16;         // ranks are u16
constexpr int STAGE_PLANE_WORDS = 72;  // planes x cell words per gene block and stage: 3 x 72 x 256 B = 54 KB per stage
__host__ __device__ constexpr int words_per_chunk(int bp) { return bp <= 4 ? 16 : STAGE_PLANE_WORDS / bp; }

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

struct Chunk {  // a run of cell words of one class inside one rank's plane region
    int32_t w_begin;    // first word, local to the owner's region
    int32_t n_words;
    int32_t class_end;  // class index if this chunk closes the class, else -1
    int32_t owner;      // rank whose region holds the words (0 on a single GPU)
};

struct TileDesc {
    int32_t ta, tb;  // a-tile (128 positions) and b-block (64 positions), tb >= 2 * ta
    int32_t bp;      // planes compared: m low planes, plus one "rank >= 2^m" plane when the two sides differ
    int32_t m;       // min(planes of the a-tile, planes of the b-block)
};

struct PairParams {
    const uint32_t *planes;
    const uint32_t *zeros;      // words_per_chunk(1) * 64 zero words: source of planes a block does not have
    const int64_t *blk_off;     // [gene blocks] first word of the block's planes
    const int32_t *blk_bits;    // [gene blocks] Bn = planes the block's genes need (>= 1)
    const int32_t *pos_gene;    // [n_genes] kept-gene index sitting at a position
    const Chunk *chunks;        // chunk tables, one per plane count
    const TileDesc *tiles;      // per CTA
    const int32_t *thr;         // per class, clipped to int32
    const int32_t *max_pos_dn;  // per class n_k - thr_k: num_neg >= thr_k needs num_pos <= this
    unsigned long long *out;    // compacted (key << 16 | class)
    unsigned long long *out_count;
    unsigned long long out_cap;
    int64_t rank_stride;  // u32 words between the plane regions of two ranks
    int n_genes, n_words, n_classes;  // n_words = words per plane of one region
    int32_t chunk_off[MAX_PLANES + 1], chunk_cnt[MAX_PLANES + 1];  // table of plane count bp
};

// Two sweeps over the cell words of a 128 x 64 tile of gene positions, one kernel.
//  Sweep 1 (all pairs): only #(x[a] > x[b]) is accumulated -- ONE LOP3 per plane per pair and word -- and
//     thresholded per class (nup, up_idx).  The decision of sandbag.py:192-197 can only emit when one of the two
//     counts reaches its threshold in exactly 1 or exactly C-1 classes; and since num_pos + num_neg <= n_k, the other
//     count can only reach thr_k if this one is <= n_k - thr_k (ncd = number of such classes): emitting needs
//     ncd >= C-1 resp. ncd >= 1.  Pairs passing these necessary conditions are the "candidates" -- well under 1 %
//     of the pairs on expression data.  The conditions are symmetric in the two counts, so they hold whichever of
//     the two genes is the reference's g1.
//  Sweep 2 (candidates only, 512 per round = one per consumer thread): the same operand stream is replayed and
//     #(x[a] < x[b]) is accumulated for the candidates, thresholded (ndn, dn_idx), and the decision emits.
// So the second predicate of sandbag.py:179-182 is evaluated exactly where it can matter, halving the LOP3 work.
// Per-pair counters are plain 32-bit registers (any class size), the state packs nup | up_idx << 16 (<= 65 534 classes).
//
// Planes.  Gene positions are sorted by the number of planes their ranks need, so a tile compares only
// m = min(planes of its a-side, planes of its b-side) low planes; if the sides differ, the wider side adds its
// "rank >= 2^m" plane (the narrower side has no such bits): in the recurrence that plane simply is plane m.
// BP = planes compared is a template parameter of the tile body (plane loops fully unrolled); the kernel dispatches
// per tile.
//
// 4 x 4 = per-thread register tile of gene pairs, consumer thread grid 32 x 16 = 512 threads.
// Warps 0 .. 15 consume; one extra warp is the TMA producer: per chunk of cell words and per (block, plane) one 1-D
// bulk copy into the stage [block][plane][word][64 genes].  full[s] (1 arrival + transaction bytes) hands a stage to
// the consumers, empty[s] (one arrival per consumer warp) hands it back, so the consumer warps never meet at a
// CTA-wide barrier inside a sweep and drift up to STAGES-1 chunks apart.
constexpr int TI = 4, TJ = 4;
constexpr int NTJ = TILE_B / TJ;
constexpr int PAIR_THREADS = (TILE_A / TI) * NTJ;  // consumer threads
constexpr int CONSUMER_WARPS = PAIR_THREADS / 32;
constexpr int CAND_PER_ROUND = PAIR_THREADS;  // one candidate per consumer thread and round
constexpr int CHUNK_CACHE = 192;              // chunk descriptors kept in shared memory (global beyond that)

// shared state of a pair_kernel CTA (namespace scope so that the tile bodies address it directly)
extern __shared__ __align__(128) uint8_t pk_stage_raw[];  // dynamic: STAGES stages of [block][plane][word][64]
__shared__ __align__(8) uint64_t pk_full_bar[STAGES];
__shared__ __align__(8) uint64_t pk_empty_bar[STAGES];
__shared__ uint2 pk_cand[CAND_PER_ROUND];  // .x = la | lb << 8 (tile-local positions), .y = nup | up_idx << 16
__shared__ int pk_total;
__shared__ Chunk pk_chunks[CHUNK_CACHE];

template <int BP>
__device__ __forceinline__ void tile_body(const PairParams &P, const TileDesc td) {
    constexpr int WC = words_per_chunk(BP);
    constexpr uint32_t BLK_WORDS = (uint32_t)BP * WC * 64;  // u32 per gene block per stage
    constexpr uint32_t STAGE_WORDS = (A_BLOCKS + 1) * BLK_WORDS;
    constexpr int PLANE = WC * 64;  // u32 between two planes of a block inside a stage
    const int tid = threadIdx.x;
    const int ti = tid / NTJ, tj = tid % NTJ;
    const int n_chunks = P.chunk_cnt[BP];
    uint32_t *smem = reinterpret_cast<uint32_t *>(pk_stage_raw);

    for (int c = tid; c < min(n_chunks, CHUNK_CACHE); c += blockDim.x) pk_chunks[c] = P.chunks[P.chunk_off[BP] + c];
    __syncthreads();
    auto chunk_at = [&](int c) -> Chunk { return c < CHUNK_CACHE ? pk_chunks[c] : P.chunks[P.chunk_off[BP] + c]; };

    if (tid >= PAIR_THREADS) {  // ---- producer warp: streams the chunk sequence once per sweep
        const int lane = tid - PAIR_THREADS;
        // copy i = (block i / BP, plane i % BP) of every chunk belongs to lane i % 32
        constexpr int N_COPIES = (A_BLOCKS + 1) * BP, PER_LANE = (N_COPIES + 31) / 32;
        const uint32_t *src[PER_LANE];
        bool real[PER_LANE];
#pragma unroll
        for (int j = 0; j < PER_LANE; ++j) {
            const int i = lane + 32 * j;
            src[j] = P.zeros;
            real[j] = false;
            if (i < N_COPIES) {
                const int blk = i / BP, q = i % BP;
                const int nb = blk < A_BLOCKS ? A_BLOCKS * td.ta + blk : td.tb;
                const int bn = __ldg(P.blk_bits + nb);
                int plane = -1;
                if (q < td.m) {
                    if (q < bn) plane = q;
                } else if (bn > td.m) {
                    plane = bn + td.m - 1;  // suffix plane S_m
                }
                if (plane >= 0) {
                    src[j] = P.planes + __ldg(P.blk_off + nb) + (int64_t)plane * P.n_words * 64;
                    real[j] = true;
                }
            }
        }
        int q = 0;  // running chunk number over all sweeps (stage = q % STAGES)
        auto stream_sweep = [&]() {
            for (int c = 0; c < n_chunks; ++c, ++q) {
                const int s = q % STAGES;
                const Chunk ch = chunk_at(c);
                const uint32_t bytes = (uint32_t)ch.n_words * 64 * 4;
                if (lane == 0) {
                    if (q >= STAGES) mbar_wait(&pk_empty_bar[s], (uint32_t)(q / STAGES - 1) & 1u);  // consumers released it
                    mbar_arrive_expect_tx(&pk_full_bar[s], N_COPIES * bytes);
                }
                __syncwarp();
                uint32_t *dst = smem + (size_t)s * STAGE_WORDS;
#pragma unroll
                for (int j = 0; j < PER_LANE; ++j) {
                    const int i = lane + 32 * j;
                    if (i < N_COPIES)
                        bulk_g2s(dst + i * PLANE,
                                 src[j] + (real[j] ? (int64_t)ch.owner * P.rank_stride + (int64_t)ch.w_begin * 64 : 0), bytes,
                                 &pk_full_bar[s]);
                }
            }
        };
        stream_sweep();
        __syncthreads();  // (A) sweep 1 is consumed and the candidates are counted
        const int rounds = (pk_total + CAND_PER_ROUND - 1) / CAND_PER_ROUND;
        for (int r = 0; r < rounds; ++r) stream_sweep();
        return;
    }

    // =============================== sweep 1: #(a > b) of every pair of the tile
    uint32_t acc[TI][TJ], st[TI][TJ], ncd[TI][TJ];
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) acc[i][j] = 0, st[i][j] = 0, ncd[i][j] = 0;

    const int la = ti * TI;  // first local a-position (0..127), never straddles a 64-gene block
    const uint32_t a_off = (uint32_t)(la >> 6) * BLK_WORDS + (la & 63);
    const uint32_t b_off = (uint32_t)A_BLOCKS * BLK_WORDS + tj * TJ;
    int q = 0;

    for (int c = 0; c < n_chunks; ++c, ++q) {
        const int s = q % STAGES;
        const Chunk ch = chunk_at(c);
        mbar_wait(&pk_full_bar[s], (uint32_t)(q / STAGES) & 1u);
        const uint32_t *sa = smem + (size_t)s * STAGE_WORDS + a_off;
        const uint32_t *sb = smem + (size_t)s * STAGE_WORDS + b_off;
        // software-pipelined: the operands of plane p+1 (or of plane 0 of the next word) are requested before the
        // 16 LOP3 of plane p issue, so the LDS latency hides behind them
        uint32_t a[TI], b[TJ];
        load_words<TI>(sa, a);
        load_words<TJ>(sb, b);
        for (int w = 0; w < ch.n_words; ++w) {
            uint32_t gt[TI][TJ];
#pragma unroll
            for (int p = 0; p < BP; ++p) {
                uint32_t an[TI], bn[TJ];
                const int nxt = (p + 1 < BP) ? (p + 1) * PLANE : 64;  // next plane, or plane 0 of the next word
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
            sa += 64;
            sb += 64;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&pk_empty_bar[s]);  // this warp is done reading stage s
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

    // =============================== candidates: pairs whose decision can still emit
    const int pa0 = td.ta * TILE_A + la, pb0 = td.tb * TILE_B + tj * TJ;
    const uint32_t C = (uint32_t)P.n_classes;
    const int lane = tid & 31;
    uint32_t cmask = 0;
#pragma unroll
    for (int i = 0; i < TI; ++i)
#pragma unroll
        for (int j = 0; j < TJ; ++j) {
            const uint32_t nup = st[i][j] & 0xffffu;
            const bool can = (nup == 1u) ? (ncd[i][j] >= C - 1u) : (nup == C - 1u && ncd[i][j] >= 1u);
            if (pa0 + i < pb0 + j && pb0 + j < P.n_genes && can) cmask |= 1u << (i * TJ + j);
        }
    const int my_first = cmask ? atomicAdd(&pk_total, __popc(cmask)) : 0;  // tickets of this thread's candidates
    __syncthreads();  // (A)
    const int total = pk_total;
    const int rounds = (total + CAND_PER_ROUND - 1) / CAND_PER_ROUND;

    // =============================== sweep 2: #(a < b) of the candidates, CAND_PER_ROUND per replay of the stream
    for (int r = 0; r < rounds; ++r) {
        {
            int t = my_first;
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < TJ; ++j)
                    if (cmask & (1u << (i * TJ + j))) {
                        if (t / CAND_PER_ROUND == r)
                            pk_cand[t % CAND_PER_ROUND] = make_uint2((uint32_t)(la + i) | ((uint32_t)(tj * TJ + j) << 8), st[i][j]);
                        ++t;
                    }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(PAIR_THREADS) : "memory");  // consumers only: the list is complete
        const bool have = tid < min(CAND_PER_ROUND, total - r * CAND_PER_ROUND);
        const uint2 me = have ? pk_cand[tid] : make_uint2(0u, 0u);
        const uint32_t ca = me.x & 0xffu, cb = me.x >> 8;
        const uint32_t ca_off = (ca >> 6) * BLK_WORDS + (ca & 63), cb_off = (uint32_t)A_BLOCKS * BLK_WORDS + cb;
        uint32_t neg = 0, ndn = 0, dn_idx = 0;
        for (int c = 0; c < n_chunks; ++c, ++q) {
            const int s = q % STAGES;
            const Chunk ch = chunk_at(c);
            mbar_wait(&pk_full_bar[s], (uint32_t)(q / STAGES) & 1u);
            if (have) {
                const uint32_t *sa = smem + (size_t)s * STAGE_WORDS + ca_off;
                const uint32_t *sb = smem + (size_t)s * STAGE_WORDS + cb_off;
                for (int w = 0; w < ch.n_words; ++w) {
                    uint32_t lt = 0;
#pragma unroll
                    for (int p = 0; p < BP; ++p) lt = lop3_lt(sa[p * PLANE], sb[p * PLANE], lt);
                    neg += __popc(lt);
                    sa += 64;
                    sb += 64;
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&pk_empty_bar[s]);
            if (ch.class_end >= 0) {  // `num_neg[i] >= thresholds[i]` of sandbag.py:188-190
                if ((int64_t)neg >= (int64_t)__ldg(P.thr + ch.class_end)) {
                    ++ndn;
                    dn_idx = (uint32_t)ch.class_end;
                }
                neg = 0;
            }
        }
        // decision of sandbag.py:192-197 + warp-aggregated compaction.  The reference's g1 is the gene with the
        // smaller index: if that is the b-position, the two counts (and their class tallies) swap roles.
        unsigned long long key = 0;
        bool emit = false;
        if (have) {
            const unsigned long long ga = (unsigned long long)__ldg(P.pos_gene + td.ta * TILE_A + (int)ca);
            const unsigned long long gb = (unsigned long long)__ldg(P.pos_gene + td.tb * TILE_B + (int)cb);
            const bool a_first = ga < gb;
            const unsigned long long g1 = a_first ? ga : gb, g2 = a_first ? gb : ga;
            const uint32_t nup = a_first ? (me.y & 0xffffu) : ndn, up_idx = a_first ? (me.y >> 16) : dn_idx;
            const uint32_t ndn_c = a_first ? ndn : (me.y & 0xffffu), dn_idx_c = a_first ? dn_idx : (me.y >> 16);
            if (nup == 1u) {
                if (ndn_c == C - 1u) {
                    emit = true;
                    key = ((g1 * P.n_genes + g2) << 16) | up_idx;
                }
            } else if (ndn_c == 1u) {
                if (nup == C - 1u) {
                    emit = true;
                    key = ((g2 * P.n_genes + g1) << 16) | dn_idx_c;
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

__global__ void __launch_bounds__(PAIR_THREADS + 32, 1) pair_kernel(const PairParams P) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&pk_full_bar[s], 1);
            mbar_init(&pk_empty_bar[s], CONSUMER_WARPS);
        }
        fence_mbar_init();
        pk_total = 0;
    }
    const TileDesc td = P.tiles[blockIdx.x];
    switch (td.bp) {  // tile_body starts with a CTA barrier that also publishes the mbarrier initialisation
        case 1: tile_body<1>(P, td); break;
        case 2: tile_body<2>(P, td); break;
        case 3: tile_body<3>(P, td); break;
        case 4: tile_body<4>(P, td); break;
        case 5: tile_body<5>(P, td); break;
        case 6: tile_body<6>(P, td); break;
        case 7: tile_body<7>(P, td); break;
        case 8: tile_body<8>(P, td); break;
        case 9: tile_body<9>(P, td); break;
        case 10: tile_body<10>(P, td); break;
        case 11: tile_body<11>(P, td); break;
        case 12: tile_body<12>(P, td); break;
        case 13: tile_body<13>(P, td); break;
        case 14: tile_body<14>(P, td); break;
        case 15: tile_body<15>(P, td); break;
        default: tile_body<16>(P, td); break;
    }
}

// per-class counters of selected pairs straight from the planes (one warp per probe pair)
__global__ void probe_kernel(const uint32_t *__restrict__ planes, const int64_t *__restrict__ blk_off,
                             const int32_t *__restrict__ blk_bits, const int32_t *__restrict__ gene_pos,
                             const Chunk *__restrict__ chunks, int n_chunks, int n_words, int64_t rank_stride,
                             int n_classes, const int64_t *__restrict__ pairs, int64_t n_probe, int32_t *__restrict__ pos_out,
                             int32_t *__restrict__ neg_out) {
    const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= n_probe) return;
    const int p1 = gene_pos[(int)pairs[2 * q]], p2 = gene_pos[(int)pairs[2 * q + 1]];
    const int b1 = blk_bits[p1 >> 6], b2 = blk_bits[p2 >> 6];
    const uint32_t *pa = planes + blk_off[p1 >> 6] + (p1 & 63);
    const uint32_t *pb = planes + blk_off[p2 >> 6] + (p2 & 63);
    const int64_t ps = (int64_t)n_words * 64;
    int pos = 0, neg = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const Chunk ch = chunks[c];
        const int64_t reg = (int64_t)ch.owner * rank_stride;
        for (int w = ch.w_begin + lane; w < ch.w_begin + ch.n_words; w += 32) {
            uint32_t gt = 0, lt = 0;
            for (int p = 0; p < max(b1, b2); ++p) {
                const uint32_t a = p < b1 ? pa[reg + p * ps + (int64_t)w * 64] : 0u;
                const uint32_t b = p < b2 ? pb[reg + p * ps + (int64_t)w * 64] : 0u;
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
        const uint32_t peers = warp_digit_peers(d, valid);
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
        const uint32_t peers = warp_digit_peers(d, valid);
        const uint32_t base = valid ? whist[warp][d] : 0u;
        if (valid) out[base + __popc(peers & ((1u << lane) - 1u))] = k;
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) whist[warp][d] = base + __popc(peers);
        __syncwarp();
    }
}

// (key << 16 | class) -> the two arrays the caller receives
__global__ void split_keys_kernel(const unsigned long long *__restrict__ packed, unsigned long long n,
                                  uint64_t *__restrict__ keys, int32_t *__restrict__ cls) {
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long v = packed[i];
        keys[i] = v >> 16;
        cls[i] = (int32_t)(v & 0xffffu);
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

// Shared body of pp_sandbag / pp_sandbag_csr (csr != NULL: x / ld unused) and their _dist forms.
//
// Cells are laid out in PIECES: piece (q, k) = the cells of class k owned by data rank q, padded to whole 32-cell
// words; the planes of rank q's pieces form region q of the plane buffer.  On one GPU there is one data rank that
// owns everything (and `shard / n_shards` only deals the gene-pair tiles, every shard preparing all cells itself);
// with dist = true the data ranks are the ranks of the ctx's communicator, every rank prepares only its own region,
// one all-gather completes the buffer, and rank r evaluates the tiles t % world == r.
static int sandbag_impl(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, const pp_csr *csr, int64_t n_rows,
                        int64_t n_cols, int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int64_t *class_sizes,
                        int32_t n_classes, const int32_t *gene_cols, int64_t n_genes, const int64_t *thresholds,
                        int32_t shard, int32_t n_shards, bool dist, int32_t gather_to, uint64_t **out_keys,
                        int32_t **out_class, int64_t *out_n, const int64_t *probe_pairs, int64_t n_probe,
                        int32_t *probe_pos, int32_t *probe_neg, int32_t drop_unexpressed, int32_t **out_kept_genes,
                        int64_t *out_n_kept) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    // PP_TRACE=1: host-side time line of the call on stderr (ms since entry at every mark)
    const bool trace = getenv("PP_TRACE") != nullptr;
    const auto t_entry = std::chrono::steady_clock::now();
    std::string trace_line;
    auto mark = [&](const char *what) {
        if (!trace) return;
        char buf[64];
        snprintf(buf, sizeof(buf), " %s=%.3f", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count());
        trace_line += buf;
    };
    if ((!x && !csr) || !cell_rows || !class_sizes || !gene_cols || !thresholds || !out_keys || !out_class || !out_n)
        return pp_fail(ctx, PP_EINVAL, "pp_sandbag: NULL argument");
    if (x_dtype != PP_F32 && x_dtype != PP_F64) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: x_dtype must be PP_F32/PP_F64");
    if (n_rows <= 0 || n_cols <= 0 || (!csr && ld < n_cols)) return pp_fail(ctx, PP_EINVAL, "pp_sandbag: bad matrix shape");
    if (dist) {
        if (!ctx->comm) return pp_fail(ctx, PP_ENCCL, "pp_sandbag_dist: the context has no communicator (pp_comm_init)");
        if (gather_to >= ctx->comm_world) return pp_fail(ctx, PP_EINVAL, "pp_sandbag_dist: gather_to %d of %d ranks", gather_to, ctx->comm_world);
        shard = ctx->comm_rank;
        n_shards = ctx->comm_world;
    }
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

    // ---- data ranks: contiguous row blocks [rb[q], rb[q+1]) holding equal numbers of kept cells
    const int wd = dist ? ctx->comm_world : 1, me = dist ? ctx->comm_rank : 0;
    const bool receive = !dist || gather_to < 0 || gather_to == me;
    std::vector<int64_t> rb((size_t)wd + 1, 0);
    rb[wd] = n_rows;
    if (wd > 1 && n_cells > 0) {  // rb[q] = row of the (n_cells * q / wd)-th kept cell in row order (counting, no sort)
        std::vector<uint8_t> is_kept((size_t)n_rows, 0);
        for (int64_t i = 0; i < n_cells; ++i) is_kept[(size_t)cell_rows[i]] = 1;
        int64_t seen = 0;
        int q = 1;
        for (int64_t r = 0; r < n_rows && q < wd; ++r) {
            if (!is_kept[(size_t)r]) continue;
            while (q < wd && seen == n_cells * q / wd) rb[q++] = r;
            ++seen;
        }
        for (; q < wd; ++q) rb[q] = n_rows;
    }
    auto owner_of = [&](int64_t row) { return (int)(std::upper_bound(rb.begin() + 1, rb.begin() + wd, row) - (rb.begin() + 1)); };
    // pieces: cells of (owner q, class k), class-sorted order kept
    std::vector<int64_t> piece_cnt((size_t)wd * n_classes, 0);
    std::vector<std::vector<int32_t>> my_cells((size_t)n_classes);
    {
        int64_t src = 0;
        for (int k = 0; k < n_classes; ++k) {
            for (int64_t i = 0; i < class_sizes[k]; ++i) {
                const int32_t row = cell_rows[src + i];
                const int q = wd > 1 ? owner_of(row) : 0;
                ++piece_cnt[(size_t)q * n_classes + k];
                if (q == me) my_cells[k].push_back(row);
            }
            src += class_sizes[k];
        }
    }
    std::vector<int32_t> piece_w((size_t)wd * n_classes), piece_w0((size_t)wd * n_classes);
    int W_me = 0, Wl = 1;  // words of my region in use; words per plane of every region
    for (int q = 0; q < wd; ++q) {
        int64_t w = 0;
        for (int k = 0; k < n_classes; ++k) {
            piece_w0[(size_t)q * n_classes + k] = (int32_t)w;
            piece_w[(size_t)q * n_classes + k] = (int32_t)((piece_cnt[(size_t)q * n_classes + k] + 31) / 32);
            w += piece_w[(size_t)q * n_classes + k];
        }
        if (w > (1 << 26)) return pp_fail(ctx, PP_ELIMIT, "pp_sandbag: more than 2^31 cells");
        if (q == me) W_me = (int)w;
        Wl = std::max(Wl, (int)w);
    }
    const int64_t r0 = rb[me], nr_local = rb[me + 1] - rb[me];  // my row block

    mark("layout");
    PP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PP_CUDA(cudaEventRecord(ctx->ev[0], st));
    DevBuf d_slots, d_gene, d_red, d_chunks, d_tiles, d_thr, d_cnt;
    void *s_x = nullptr, *s_rank = nullptr, *s_planes = nullptr, *s_out = nullptr, *s_out2 = nullptr;  // ctx slabs
    const size_t esz = x_dtype == PP_F64 ? 8 : 4;
    const uint8_t *dx_local = nullptr;  // row r0 of the matrix on the device (dense input)
    DevCsr dcsr;                        // rows [r0, r0 + nr_local), rebased to 0 (CSR input)
    int rc = PP_OK;
    if (csr) {
        rc = pp_csr_to_device(ctx, csr, r0, nr_local, &dcsr);
        if (rc) return rc;
    } else if (!x_is_device) {
        if (!(s_x = pp_slab(ctx, SLAB_X, (size_t)std::max<int64_t>(nr_local, 1) * ld * esz))) return PP_ENOMEM;
        if (nr_local > 0) {
            const size_t bytes = ((size_t)(nr_local - 1) * ld + (size_t)n_cols) * esz;  // the last row may end at n_cols
            rc = pp_copy_h2d(ctx, s_x, (const uint8_t *)x + (size_t)r0 * ld * esz, bytes, st);
            if (rc) return rc;
            ctx->tm.h2d_bytes = (int64_t)bytes;
        }
        dx_local = (const uint8_t *)s_x;
    } else {
        dx_local = (const uint8_t *)x + (size_t)r0 * ld * esz;
    }
    PP_CUDA(cudaEventRecord(ctx->ev[1], st));

    // ---- optional: drop genes that are zero in ALL rows of x (utils.py:368, before sample filtering); every data
    //      rank scans its own row block, the flags are max-reduced
    std::vector<int32_t> kept;
    if (drop_unexpressed) {
        DevBuf d_nz;
        PP_CUDA(d_nz.alloc((size_t)n_cols, st));
        PP_CUDA(cudaMemsetAsync(d_nz.p, 0, (size_t)n_cols, st));
        if (csr) {
            rc = pp_csr_col_nonzero(ctx, dcsr, d_nz.as<uint8_t>());
            if (rc) return rc;
        } else if (nr_local > 0) {
            dim3 grid((unsigned)((n_cols + 255) / 256), (unsigned)std::min<int64_t>((nr_local + 63) / 64, 65535));
            if (x_dtype == PP_F64)
                col_nonzero_kernel<double><<<grid, 256, 0, st>>>((const double *)dx_local, ld, nr_local, n_cols, d_nz.as<uint8_t>());
            else
                col_nonzero_kernel<float><<<grid, 256, 0, st>>>((const float *)dx_local, ld, nr_local, n_cols, d_nz.as<uint8_t>());
            PP_CHECK_LAUNCH();
        }
        if (wd > 1) {
            rc = pp_allreduce_max(ctx, d_nz.p, (size_t)n_cols, ppn::ncclUint8);
            if (rc) return rc;
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
    if (n_genes < 2 || n_cells == 0) return PP_OK;  // no pairs (the same decision on every rank)
    if (n_genes > 65535) return pp_fail(ctx, PP_ELIMIT, "pp_sandbag: more than 65535 kept genes");

    // ---- on a helper thread, while this one feeds the device: everything about tiles and chunks that does not depend on
    //      the rank widths (0.4 ms of host work for cfg3, 2 ms for cfg4 -- on the critical path otherwise, at 8 GPUs the
    //      device work before the pair kernel is shorter than that).
    //      Tiles of this shard: (ta, tb) with tb >= A_BLOCKS * ta (only those contain position pairs a < b).
    const int NB_t = (int)((n_genes + 63) / 64);
    const int TA = (int)((n_genes + TILE_A - 1) / TILE_A), TB = NB_t;
    std::vector<TileDesc> tiles;
    std::vector<int64_t> tile_pairs;
    std::vector<Chunk> chunks;
    PairParams P;
    bool tables_failed = false;
    auto build_tables = [&]() {
        try {
            {
                int64_t t_index = 0;
                // Super-tile order: the ~n_sm tiles that run at the same time come from one SUP_A x SUP_B patch of the tile
                // grid, so they share their 2*SUP_A + SUP_B operand blocks through L2 instead of streaming each operand
                // from HBM once per tile (ncu: 11.7 GB of DRAM reads per cfg3 launch in row-major tile order).
                constexpr int SUP_A = 8, SUP_B = 18;
                for (int sa = 0; sa < TA; sa += SUP_A)
                    for (int sb = A_BLOCKS * sa; sb < TB; sb += SUP_B)
                        for (int ta = sa; ta < std::min(sa + SUP_A, TA); ++ta)
                            for (int tb = std::max(sb, A_BLOCKS * ta); tb < std::min(sb + SUP_B, TB); ++tb, ++t_index) {
                                if (t_index % n_shards != shard) continue;
                                TileDesc td;
                                td.ta = ta;
                                td.tb = tb;
                                td.m = td.bp = 0;  // filled in once the rank widths are known
                                tiles.push_back(td);
                                const int64_t a0 = (int64_t)ta * TILE_A, a1 = std::min<int64_t>(a0 + TILE_A, n_genes);
                                const int64_t b0 = (int64_t)tb * TILE_B, b1 = std::min<int64_t>(b0 + TILE_B, n_genes);
                                int64_t np_tile = std::max<int64_t>(0, a1 - a0) * std::max<int64_t>(0, b1 - b0);  // all a < all b
                                if (b0 < a1) {  // the tile touches the diagonal
                                    np_tile = 0;
                                    for (int64_t a = a0; a < a1; ++a) np_tile += std::max<int64_t>(0, b1 - std::max(b0, a + 1));
                                }
                                tile_pairs.push_back(np_tile);
                            }
            }
            // Chunks: for every class, region after region, runs of <= words_per_chunk(bp) words of that class; one table per
            // plane count (all of them: which ones are in use is only known later, and they are small); table 0 (one chunk
            // per piece) serves the probe kernel.
            for (int bp = 0; bp <= MAX_PLANES; ++bp) {
                P.chunk_off[bp] = (int32_t)chunks.size();
                const int wc = bp == 0 ? Wl : words_per_chunk(bp);
                for (int k = 0; k < n_classes; ++k) {
                    for (int q = 0; q < wd; ++q) {
                        const int w0 = piece_w0[(size_t)q * n_classes + k], w1 = w0 + piece_w[(size_t)q * n_classes + k];
                        for (int w = w0; w < w1; w += wc) {
                            Chunk c;
                            c.w_begin = w;
                            c.n_words = std::min(wc, w1 - w);
                            c.class_end = -1;
                            c.owner = q;
                            chunks.push_back(c);
                        }
                    }
                    chunks.back().class_end = k;  // classes are non-empty: the last chunk pushed belongs to class k
                }
                P.chunk_cnt[bp] = (int32_t)chunks.size() - P.chunk_off[bp];
            }
        } catch (const std::exception &) {
            tables_failed = true;
        }
    };
    struct Joiner {  // joined on every exit path
        std::thread t;
        ~Joiner() {
            if (t.joinable()) t.join();
        }
    } table_thread;
    try {
        table_thread.t = std::thread(build_tables);
    } catch (const std::exception &) {
        build_tables();  // no thread to be had: build them here
    }

    // ---- my slots: row (relative to my block) of every cell of my pieces, -1 = padding cell
    const int64_t n_slots = (int64_t)W_me * 32;
    std::vector<int32_t> slots((size_t)std::max<int64_t>(n_slots, 1), -1);
    for (int k = 0; k < n_classes; ++k) {  // dx_local / dcsr start at row r0 in every input form
        const int64_t base = (int64_t)piece_w0[(size_t)me * n_classes + k] * 32;
        for (size_t i = 0; i < my_cells[k].size(); ++i) slots[(size_t)base + i] = (int32_t)(my_cells[k][i] - r0);
    }
    const int NB = (int)((n_genes + 63) / 64);
    const int NB_alloc = (NB + 1) & ~1;

    PP_CUDA(d_slots.alloc(sizeof(int32_t) * slots.size(), st));
    PP_CUDA(cudaMemcpyAsync(d_slots.p, slots.data(), sizeof(int32_t) * slots.size(), cudaMemcpyHostToDevice, st));
    PP_CUDA(d_gene.alloc(sizeof(int32_t) * n_genes, st));
    PP_CUDA(cudaMemcpyAsync(d_gene.p, gene_cols, sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice, st));
    if (!(s_rank = pp_slab(ctx, SLAB_RANK, sizeof(uint16_t) * std::max<int64_t>(n_slots, 1) * n_genes))) return PP_ENOMEM;
    // one buffer for everything that is max-reduced after ranking: [0] largest rank, [1] NaN seen, [4 + g] OR of gene g's ranks
    PP_CUDA(d_red.alloc(sizeof(uint32_t) * (4 + (size_t)n_genes), st));
    PP_CUDA(cudaMemsetAsync(d_red.p, 0, sizeof(uint32_t) * (4 + (size_t)n_genes), st));
    int32_t *d_flags = d_red.as<int32_t>();
    uint32_t *d_orbits = d_red.as<uint32_t>() + 4;

    if (csr && n_slots > 0) {
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
                                n_genes, (uint16_t *)s_rank + s0 * n_genes, n_genes, d_flags, ident.data());
            if (rc) return rc;
        }
        PP_CUDA(cudaStreamSynchronize(st));  // colmap / ident / loc are locals
    } else if (n_slots > 0) {
        rc = pp_rank_launch(ctx, dx_local, x_dtype, ld, d_slots.as<int32_t>(), n_slots, d_gene.as<int32_t>(), n_genes,
                            (uint16_t *)s_rank, n_genes, d_flags, gene_cols);
        if (rc) return rc;
    }
    // ---- planes every gene needs: bit_length(OR of its ranks) -- the max over the data ranks of their ORs has the same
    //      bit length as the OR over all cells
    if (n_slots > 0) {
        dim3 grid((unsigned)((n_genes + 255) / 256), (unsigned)std::min<int64_t>(std::max<int64_t>(n_slots / 128, 1), 128));
        col_or_kernel<<<grid, 256, 0, st>>>((uint16_t *)s_rank, n_slots, (int)n_genes, d_orbits);
        PP_CHECK_LAUNCH();
    }
    if (wd > 1) {
        rc = pp_allreduce_max(ctx, d_red.p, 4 + (size_t)n_genes, ppn::ncclUint32);
        if (rc) return rc;
    }
    std::vector<uint32_t> red(4 + (size_t)n_genes);
    PP_CUDA(cudaMemcpyAsync(red.data(), d_red.p, sizeof(uint32_t) * red.size(), cudaMemcpyDeviceToHost, st));

    PP_CUDA(cudaStreamSynchronize(st));
    mark("ranked");
    if (red[1]) return pp_fail(ctx, PP_ENAN, "pp_sandbag: NaN in the expression matrix");
    const int B = std::max(1, bit_length64((uint64_t)red[0]));
    ctx->tm.n_planes = B;
    const uint32_t *orbits = red.data() + 4;

    // gene positions sorted by the planes they need (widest first, so the expensive tiles start first)
    std::vector<int32_t> pos_gene((size_t)n_genes), gene_pos((size_t)n_genes);
    {
        std::vector<int32_t> start(MAX_PLANES + 2, 0);  // counting sort by descending plane count, stable
        for (int64_t g = 0; g < n_genes; ++g) ++start[MAX_PLANES - bit_length64(orbits[g]) + 1];
        for (int i = 0; i <= MAX_PLANES; ++i) start[i + 1] += start[i];
        for (int64_t g = 0; g < n_genes; ++g) {
            const int32_t p = start[MAX_PLANES - bit_length64(orbits[g])]++;
            pos_gene[p] = (int32_t)g;
            gene_pos[g] = p;
        }
    }
    std::vector<int32_t> blk_bits((size_t)NB_alloc, 1);
    std::vector<int64_t> blk_off((size_t)NB_alloc, 0);
    size_t plane_words = 0;  // words of ONE region
    for (int nb = 0; nb < NB_alloc; ++nb) {
        int bn = 1;
        for (int64_t p = (int64_t)nb * 64; p < std::min<int64_t>((int64_t)nb * 64 + 64, n_genes); ++p)
            bn = std::max(bn, bit_length64(orbits[pos_gene[p]]));
        blk_bits[nb] = bn;
        blk_off[nb] = (int64_t)plane_words;
        plane_words += (size_t)(2 * bn - 1) * Wl * 64;  // bn low planes + suffix planes S_1 .. S_{bn-1}
    }
    const size_t zero_words = (size_t)words_per_chunk(1) * 64;
    if (!(s_planes = pp_slab(ctx, SLAB_PLANES, ((size_t)wd * plane_words + zero_words) * 4))) return PP_ENOMEM;
    uint32_t *my_region = (uint32_t *)s_planes + (size_t)me * plane_words;
    uint32_t *zeros = (uint32_t *)s_planes + (size_t)wd * plane_words;
    PP_CUDA(cudaMemsetAsync(my_region, 0, plane_words * 4, st));  // padding words of my region stay all-tie
    PP_CUDA(cudaMemsetAsync(zeros, 0, zero_words * 4, st));
    DevBuf d_pos_gene, d_gene_pos, d_blk_bits, d_blk_off;
    PP_CUDA(d_pos_gene.alloc(sizeof(int32_t) * n_genes, st));
    PP_CUDA(cudaMemcpyAsync(d_pos_gene.p, pos_gene.data(), sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice, st));
    PP_CUDA(d_blk_bits.alloc(sizeof(int32_t) * NB_alloc, st));
    PP_CUDA(cudaMemcpyAsync(d_blk_bits.p, blk_bits.data(), sizeof(int32_t) * NB_alloc, cudaMemcpyHostToDevice, st));
    PP_CUDA(d_blk_off.alloc(sizeof(int64_t) * NB_alloc, st));
    PP_CUDA(cudaMemcpyAsync(d_blk_off.p, blk_off.data(), sizeof(int64_t) * NB_alloc, cudaMemcpyHostToDevice, st));

    // ---- bit-slice my cells into my region; the other regions arrive with one all-gather over NVLink
    if (W_me > 0) {
        dim3 grid((unsigned)((n_genes + 255) / 256), (unsigned)std::min(W_me, 32768));
        bitslice_kernel<<<grid, 256, 0, st>>>((uint16_t *)s_rank, n_genes, (int)n_genes, W_me, Wl, d_pos_gene.as<int32_t>(),
                                              d_blk_off.as<int64_t>(), d_blk_bits.as<int32_t>(), my_region);
        PP_CHECK_LAUNCH();
    }
    if (wd > 1) {
        rc = pp_allgather_bytes(ctx, my_region, s_planes, plane_words * 4);
        if (rc) return rc;
    }
    PP_CUDA(cudaEventRecord(ctx->ev[2], st));
    mark("bitslice_enqueued");

    // ---- plane counts of this shard's tiles (the tile list itself was made on the helper thread meanwhile)
    if (table_thread.t.joinable()) table_thread.t.join();
    if (tables_failed) return pp_fail(ctx, PP_ENOMEM, "pp_sandbag: host allocation of the tile tables failed");
    mark("tables_joined");
    int64_t n_pairs_shard = 0;
    double plane_pairs = 0.0;  // sum over the shard's gene pairs of the planes compared
    for (size_t t = 0; t < tiles.size(); ++t) {
        TileDesc &td = tiles[t];
        const int ba = std::max(blk_bits[A_BLOCKS * td.ta], blk_bits[A_BLOCKS * td.ta + 1]), bb = blk_bits[td.tb];
        td.m = std::min(ba, bb);
        td.bp = td.m + (ba != bb ? 1 : 0);
        n_pairs_shard += tile_pairs[t];
        plane_pairs += (double)tile_pairs[t] * td.bp;
    }
    ctx->tm.n_units = n_pairs_shard;
    ctx->tm.avg_planes = n_pairs_shard > 0 ? (float)(plane_pairs / (double)n_pairs_shard) : 0.f;

    const size_t dyn_smem = (size_t)STAGES * (A_BLOCKS + 1) * STAGE_PLANE_WORDS * 64 * 4;
    std::vector<int32_t> thr32(n_classes), mpd32(n_classes);
    for (int k = 0; k < n_classes; ++k) {
        thr32[k] = (int32_t)std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, thresholds[k]));
        mpd32[k] = (int32_t)std::max<int64_t>(INT32_MIN, std::min<int64_t>(INT32_MAX, class_sizes[k] - thresholds[k]));
    }

    PP_CUDA(d_chunks.alloc(sizeof(Chunk) * chunks.size(), st));
    PP_CUDA(cudaMemcpyAsync(d_chunks.p, chunks.data(), sizeof(Chunk) * chunks.size(), cudaMemcpyHostToDevice, st));
    PP_CUDA(d_tiles.alloc(sizeof(TileDesc) * std::max<size_t>(tiles.size(), 1), st));
    PP_CUDA(cudaMemcpyAsync(d_tiles.p, tiles.data(), sizeof(TileDesc) * tiles.size(), cudaMemcpyHostToDevice, st));
    PP_CUDA(d_thr.alloc(sizeof(int32_t) * n_classes, st));
    PP_CUDA(cudaMemcpyAsync(d_thr.p, thr32.data(), sizeof(int32_t) * n_classes, cudaMemcpyHostToDevice, st));
    DevBuf d_mpd;
    PP_CUDA(d_mpd.alloc(sizeof(int32_t) * n_classes, st));
    PP_CUDA(cudaMemcpyAsync(d_mpd.p, mpd32.data(), sizeof(int32_t) * n_classes, cudaMemcpyHostToDevice, st));
    PP_CUDA(d_cnt.alloc(sizeof(unsigned long long) * ((size_t)wd + 1), st));  // [0] my count, [1 ..] all ranks' counts

    unsigned long long cap = (unsigned long long)std::min<int64_t>(std::max<int64_t>(n_pairs_shard, 1),
                                                                     std::max<int64_t>(1 << 22, n_pairs_shard / 16));
    unsigned long long count = 0;
    mark("tables_enqueued");
    PP_CUDA(cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (!(s_out = pp_slab(ctx, SLAB_OUT, sizeof(unsigned long long) * cap))) return PP_ENOMEM;
        PP_CUDA(cudaMemsetAsync(d_cnt.p, 0, sizeof(unsigned long long), st));
        if (!tiles.empty()) {
            P.planes = (uint32_t *)s_planes;
            P.zeros = zeros;
            P.blk_off = d_blk_off.as<int64_t>();
            P.blk_bits = d_blk_bits.as<int32_t>();
            P.pos_gene = d_pos_gene.as<int32_t>();
            P.chunks = d_chunks.as<Chunk>();
            P.tiles = d_tiles.as<TileDesc>();
            P.thr = d_thr.as<int32_t>();
            P.max_pos_dn = d_mpd.as<int32_t>();
            P.out = (unsigned long long *)s_out;
            P.out_count = d_cnt.as<unsigned long long>();
            P.out_cap = cap;
            P.rank_stride = (int64_t)plane_words;
            P.n_genes = (int)n_genes;
            P.n_words = Wl;
            P.n_classes = n_classes;
            pair_kernel<<<(unsigned)tiles.size(), PAIR_THREADS + 32, dyn_smem, st>>>(P);
            PP_CHECK_LAUNCH();
        }
        PP_CUDA(cudaEventRecord(ctx->ev[3], st));
        PP_CUDA(cudaMemcpyAsync(&count, d_cnt.p, sizeof(count), cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));

        mark("pairs_done");
        if (count <= cap) break;
        if (attempt == 1) return pp_fail(ctx, PP_ECUDA, "pp_sandbag: compaction buffer overflow after resize");
        cap = count;  // exact size known now; evaluate once more
    }

    // ---- optional probe counters (the planes are complete on every rank)
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
        PP_CUDA(d_gene_pos.alloc(sizeof(int32_t) * n_genes, st));
        PP_CUDA(cudaMemcpyAsync(d_gene_pos.p, gene_pos.data(), sizeof(int32_t) * n_genes, cudaMemcpyHostToDevice, st));
        probe_kernel<<<(unsigned)((n_probe + 3) / 4), 128, 0, st>>>(
            (uint32_t *)s_planes, d_blk_off.as<int64_t>(), d_blk_bits.as<int32_t>(), d_gene_pos.as<int32_t>(),
            d_chunks.as<Chunk>() + P.chunk_off[0], P.chunk_cnt[0], Wl, (int64_t)plane_words, n_classes, d_pp.as<int64_t>(),
            n_probe, d_pos.as<int32_t>(), d_neg.as<int32_t>());
        PP_CHECK_LAUNCH();
        PP_CUDA(cudaMemcpyAsync(probe_pos, d_pos.p, cb, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaMemcpyAsync(probe_neg, d_neg.p, cb, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
    }

    // ---- several ranks: all-gather the compacted keys (counts first, then the lists padded to the longest)
    unsigned long long *d_keys = (unsigned long long *)s_out;  // unsorted keys to sort, `count` of them
    if (wd > 1) {
        unsigned long long *d_counts = d_cnt.as<unsigned long long>() + 1;
        rc = pp_allgather_bytes(ctx, d_cnt.p, d_counts, sizeof(unsigned long long));
        if (rc) return rc;
        std::vector<unsigned long long> counts((size_t)wd);
        PP_CUDA(cudaMemcpyAsync(counts.data(), d_counts, sizeof(unsigned long long) * wd, cudaMemcpyDeviceToHost, st));
        PP_CUDA(cudaStreamSynchronize(st));
        unsigned long long maxc = 0, all = 0;
        for (int q = 0; q < wd; ++q) {
            maxc = std::max(maxc, counts[q]);
            all += counts[q];
        }
        if (all > 0) {
            if (!(s_out2 = pp_slab(ctx, SLAB_OUT2, sizeof(unsigned long long) * std::max<unsigned long long>(maxc * wd, all))))
                return PP_ENOMEM;
            unsigned long long *gbuf = (unsigned long long *)s_out2;
            if (count) PP_CUDA(cudaMemcpyAsync(gbuf + (size_t)me * maxc, s_out, sizeof(unsigned long long) * count, cudaMemcpyDeviceToDevice, st));
            rc = pp_allgather_bytes(ctx, gbuf + (size_t)me * maxc, gbuf, sizeof(unsigned long long) * maxc);
            if (rc) return rc;
            // pack the lists back to back in the compaction slab (grown if the merged list is longer than `cap`;
            // pp_slab drains the stream before it re-allocates, and my keys already sit in gbuf)
            if (receive) {
                if (!(s_out = pp_slab(ctx, SLAB_OUT, sizeof(unsigned long long) * all))) return PP_ENOMEM;
                unsigned long long off = 0;
                for (int q = 0; q < wd; ++q) {
                    if (counts[q])
                        PP_CUDA(cudaMemcpyAsync((unsigned long long *)s_out + off, gbuf + (size_t)q * maxc,
                                                sizeof(unsigned long long) * counts[q], cudaMemcpyDeviceToDevice, st));
                    off += counts[q];
                }
            }
        }
        d_keys = (unsigned long long *)s_out;
        count = receive ? all : 0;
    }

    mark("gathered");
    // ---- sort into np.where order and return
    uint64_t *hk = nullptr;
    int32_t *hc = nullptr;
    if (count > 0) {
        if (!(s_out2 = pp_slab(ctx, SLAB_OUT2, sizeof(unsigned long long) * count))) return PP_ENOMEM;
        unsigned long long *sorted = nullptr;
        const int hi_bit = 16 + bit_length64((uint64_t)n_genes * (uint64_t)n_genes);
        rc = radix_sort_u64(ctx, d_keys, (unsigned long long *)s_out2, (int64_t)count, 16, hi_bit, &sorted);
        if (rc) return rc;
        // split (key << 16 | class) on the device and copy both arrays straight into the buffers the caller receives
        bool pin_k = false, pin_c = false;
        hk = (uint64_t *)pp_result_alloc(sizeof(uint64_t) * count, &pin_k);
        hc = (int32_t *)pp_result_alloc(sizeof(int32_t) * count, &pin_c);
        if (!hk || !hc) {
            pp_free(hk);
            pp_free(hc);
            return pp_fail(ctx, PP_ENOMEM, "pp_sandbag: host allocation of %llu markers failed", count);
        }
        unsigned long long *other = sorted == d_keys ? (unsigned long long *)s_out2 : d_keys;  // the sort's spare buffer
        uint64_t *d_k = (uint64_t *)other;
        void *s_cls = pp_slab(ctx, SLAB_MISC, sizeof(int32_t) * count);
        cudaError_t e = s_cls ? cudaSuccess : cudaErrorMemoryAllocation;
        int32_t *d_c = (int32_t *)s_cls;
        if (e == cudaSuccess) {
            split_keys_kernel<<<(unsigned)std::min<unsigned long long>((count + 255) / 256, 65535u * 16u), 256, 0, st>>>(sorted, count, d_k, d_c);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(hk, d_k, sizeof(uint64_t) * count, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(hc, d_c, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            pp_free(hk);
            pp_free(hc);
            return pp_fail(ctx, PP_ECUDA, "pp_sandbag: result copy failed: %s", cudaGetErrorString(e));
        }
        mark("sorted_copied");
    }
    PP_CUDA(cudaEventRecord(ctx->ev[4], st));
    PP_CUDA(cudaEventSynchronize(ctx->ev[4]));
    cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->tm.prep_ms, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->tm.main_ms, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->tm.post_ms, ctx->ev[3], ctx->ev[4]);
    cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[0], ctx->ev[4]);
    ctx->tm.n_launches = ctx->launches;
    mark("done");
    if (trace) fprintf(stderr, "pp_sandbag[%d/%d]:%s\n", me, wd, trace_line.c_str());
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
    PP_GUARD(ctx, sandbag_impl(ctx, x, x_dtype, x_is_device, nullptr, n_rows, n_cols, ld, cell_rows, n_cells, class_sizes,
                               n_classes, gene_cols, n_genes, thresholds, shard, n_shards, false, -1, out_keys, out_class, out_n,
                               probe_pairs, n_probe, probe_pos, probe_neg, drop_unexpressed, out_kept_genes, out_n_kept));
}

PP_API int pp_sandbag_csr(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, const int32_t *cell_rows,
                          int64_t n_cells, const int64_t *class_sizes, int32_t n_classes, const int32_t *gene_cols,
                          int64_t n_genes, const int64_t *thresholds, int32_t shard, int32_t n_shards,
                          uint64_t **out_keys, int32_t **out_class, int64_t *out_n, const int64_t *probe_pairs,
                          int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg, int32_t drop_unexpressed,
                          int32_t **out_kept_genes, int64_t *out_n_kept) {
    if (ctx && !x) return pp_fail(ctx, PP_EINVAL, "pp_sandbag_csr: NULL argument");
    PP_GUARD(ctx, sandbag_impl(ctx, nullptr, x ? x->data_dtype : PP_F32, x ? x->is_device : 0, x, n_rows, n_cols, n_cols,
                               cell_rows, n_cells, class_sizes, n_classes, gene_cols, n_genes, thresholds, shard,
                               n_shards, false, -1, out_keys, out_class, out_n, probe_pairs, n_probe, probe_pos, probe_neg,
                               drop_unexpressed, out_kept_genes, out_n_kept));
}

PP_API int pp_sandbag_dist(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                           int64_t ld, const int32_t *cell_rows, int64_t n_cells, const int64_t *class_sizes,
                           int32_t n_classes, const int32_t *gene_cols, int64_t n_genes, const int64_t *thresholds,
                           int32_t gather_to, uint64_t **out_keys, int32_t **out_class, int64_t *out_n,
                           const int64_t *probe_pairs, int64_t n_probe, int32_t *probe_pos, int32_t *probe_neg,
                           int32_t drop_unexpressed, int32_t **out_kept_genes, int64_t *out_n_kept) {
    if (ctx && !x) return pp_fail(ctx, PP_EINVAL, "pp_sandbag_dist: NULL argument");
    PP_GUARD(ctx, sandbag_impl(ctx, x, x_dtype, x_is_device, nullptr, n_rows, n_cols, ld, cell_rows, n_cells, class_sizes,
                               n_classes, gene_cols, n_genes, thresholds, 0, 1, true, gather_to, out_keys, out_class, out_n,
                               probe_pairs, n_probe, probe_pos, probe_neg, drop_unexpressed, out_kept_genes, out_n_kept));
}

PP_API int pp_sandbag_csr_dist(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, const int32_t *cell_rows,
                               int64_t n_cells, const int64_t *class_sizes, int32_t n_classes, const int32_t *gene_cols,
                               int64_t n_genes, const int64_t *thresholds, int32_t gather_to, uint64_t **out_keys,
                               int32_t **out_class, int64_t *out_n, const int64_t *probe_pairs, int64_t n_probe,
                               int32_t *probe_pos, int32_t *probe_neg, int32_t drop_unexpressed,
                               int32_t **out_kept_genes, int64_t *out_n_kept) {
    if (ctx && !x) return pp_fail(ctx, PP_EINVAL, "pp_sandbag_csr_dist: NULL argument");
    PP_GUARD(ctx, sandbag_impl(ctx, nullptr, x ? x->data_dtype : PP_F32, x ? x->is_device : 0, x, n_rows, n_cols, n_cols,
                               cell_rows, n_cells, class_sizes, n_classes, gene_cols, n_genes, thresholds, 0, 1, true,
                               gather_to, out_keys, out_class, out_n, probe_pairs, n_probe, probe_pos, probe_neg,
                               drop_unexpressed, out_kept_genes, out_n_kept));
}
