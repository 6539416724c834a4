// pp_cyclone.cu -- B200 permutation-null scoring (replaces get_phase_scores /
// get_sample_score_guv / get_proportion, pypairs/tools/cyclone.py:201-257).
//
// What the reference does per cell and class (cyclone.py:223-249): reseed MT19937 with
// settings.seed, score the observed marker pairs, then `iterations` times shuffle the
// cell's used-gene vector IN PLACE (cumulatively) and re-score.  Because every cell
// reseeds (:226), all cells of a class see the same permutation sequence, a function of
// (seed, n_used, iterations) only.  We therefore build, once per class, the table
//     E[k][slot] = position of the slot's gene under permutation k   k = 0 (identity) .. iterations
// (slots = the class's pairs regrouped into stars, see build_star_layout) and the k-th null
// score of ANY cell is the observed score evaluated through E[k] on the un-permuted vector.
// E is shared by all cells (L2 resident), cells sit in the lanes of a warp so a slot
// offset is a broadcast and a gather of v[gene] is a conflict-free shared-memory row read.
//
// Data layout
//   rank  u16 [n_cells][U]      per-cell dense ranks over the union of all classes' used
//                               genes (pp_rank.cu); order and ties inside any class subset
//                               are preserved, so every `a != b`, `a > b` is bit-exact
//   tile  u32 smem [U][32]      one CTA = 64 cells as two fp16 ranks per lane, or 128 cells as
//                               four 7-bit ranks per lane (count data); global memory beyond
//                               1 680 marker genes
//   E     u32 [1+iters][S_c]    byte offset (128 * union position) of every slot, per class
// Scores: hits_k/total_k < hits_0/total_0 is decided by exact integer cross
// multiplication in 64 bits, NaN rules of get_proportion (:217-218) applied to both
// sides; score = below / iterations as one IEEE double division (:246).
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include "pp_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// MT19937 exactly as Numba drives it (numba/cpython/randomimpl.py:109-171, 454-521,
// 1929-1946; numba_rnd_init = init_genrand).  Host side: the stream is strictly serial and
// tiny (about 1.6e6 draws per class).
// ------------------------------------------------------------------------------------------
struct Mt {
    uint32_t mt[624];
    int idx;
    void seed(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void refill() {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx = 0;
    }
    uint32_t next() {
        if (idx >= 624) refill();
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

void mt_perm_table(uint32_t seed, int n, int iters, uint16_t *out) {
    Mt g;
    g.seed(seed);
    std::vector<uint16_t> a(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) a[i] = (uint16_t)i;
    for (int k = 0; k < iters; ++k) {
        for (int i = n - 1; i > 0; --i) {
            // randint(0, i+1): low bit_length(i) bits of one draw, rejected while >= i+1
            const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)i);
            uint32_t j;
            do { j = g.next() & mask; } while (j > (uint32_t)i);
            std::swap(a[i], a[j]);
        }
        std::copy(a.begin(), a.end(), out + (size_t)k * n);
    }
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 permutation table on the device (our definition; CPU restatement in
// oracle/pp_oracle.c:orc_perm_table_philox must match bit for bit).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t w[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

__global__ void philox_perm_kernel(uint16_t *__restrict__ table, int n, int iters, uint32_t k0, uint32_t k1,
                                   uint32_t class_index) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= iters) return;
    uint16_t *a = table + (size_t)k * n;
    for (int i = 0; i < n; ++i) a[i] = (uint16_t)i;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = n - 1; i > 0; --i) {
        const int d = n - 1 - i;
        if ((d & 3) == 0) philox4x32_10((uint32_t)(d >> 2), (uint32_t)k, class_index, 0x50414952u, k0, k1, w);
        const uint32_t j = __umulhi(w[d & 3], (uint32_t)(i + 1));
        const uint16_t t = a[i];
        a[i] = a[j];
        a[j] = t;
    }
}

// Star layout.  The pairs of a class are regrouped (host, once per call) into "stars": a hub gene and
// all still-uncovered pairs that have it on one fixed side (greedy max-degree cover; the default
// markers need 653 hubs for 8 142 pairs).  The kernel then loads the hub once per star and one value
// per pair instead of two.  The stream of a class is a sequence of 4-slot vectors:
//     [hub o1 o2 o3] [o4 o5 o6 o7] ... [.. hub hub]      (a star; padded with the hub itself)
// every slot is compared against the current hub; hub and padding slots are ties and ties are counted
// on both sides of the (>=, <=) bookkeeping, so they cancel exactly.  Section A holds the stars whose
// hub is the FIRST gene of its pairs, section B those whose hub is the SECOND gene (roles of the two
// counters swap); both are padded to whole 128-slot chunks and every chunk restarts its hub, so a chunk
// is self-contained.  E[k][s] = byte offset (128 * union position) of slot s under permutation k
// (k = 0: identity = observed).
constexpr int CHUNK_SLOTS = 128;                  // 128 slots * 4 B = 512 B = one 16-byte cp.async per lane
constexpr int STAGE_BYTES = 2 * CHUNK_SLOTS * 4;  // double buffer per warp
static_assert(CHUNK_SLOTS <= 255, "MODE_B8 keeps per-chunk counts in bytes");

struct StarLayout {
    std::vector<int32_t> slot_src;     // class-compact gene index of every slot
    std::vector<uint32_t> start_mask;  // per chunk: bit q set = vector q starts a star (slot 4q is a hub)
    int n_chunks_a = 0, n_chunks = 0;
};

StarLayout build_star_layout(const int32_t *pairs, int np, int nu) {
    std::vector<std::vector<int>> adj[2];
    adj[0].resize(nu);
    adj[1].resize(nu);
    for (int i = 0; i < np; ++i) {
        adj[0][pairs[2 * i]].push_back(i);
        adj[1][pairs[2 * i + 1]].push_back(i);
    }
    std::vector<int> deg[2];
    for (int sd = 0; sd < 2; ++sd) {
        deg[sd].resize(nu);
        for (int g = 0; g < nu; ++g) deg[sd][g] = (int)adj[sd][g].size();
    }
    std::vector<char> covered(np, 0);
    struct Star { int hub, side; std::vector<int> others; };
    std::vector<Star> stars;
    // bucket queue over the degree with lazy deletion (degrees only fall): largest degree first, O(pairs + genes)
    int cur = 0;
    for (int sd = 0; sd < 2; ++sd)
        for (int g = 0; g < nu; ++g) cur = std::max(cur, deg[sd][g]);
    std::vector<std::vector<int>> bucket(cur + 1);  // entries: gene * 2 + side
    for (int sd = 1; sd >= 0; --sd)
        for (int g = nu - 1; g >= 0; --g)
            if (deg[sd][g] > 0) bucket[deg[sd][g]].push_back(g * 2 + sd);
    while (cur > 0) {
        if (bucket[cur].empty()) {
            --cur;
            continue;
        }
        const int e = bucket[cur].back();
        bucket[cur].pop_back();
        const int bg = e >> 1, bs = e & 1;
        if (deg[bs][bg] != cur) continue;  // stale entry
        Star st;
        st.hub = bg;
        st.side = bs;
        for (int i : adj[bs][bg]) {
            if (covered[i]) continue;
            covered[i] = 1;
            const int other = pairs[2 * i + (bs == 0 ? 1 : 0)];
            st.others.push_back(other);
            // the pair leaves the other endpoint's opposite-side star
            if (--deg[1 - bs][other] > 0) bucket[deg[1 - bs][other]].push_back(other * 2 + (1 - bs));
        }
        deg[bs][bg] = 0;
        stars.push_back(std::move(st));
    }
    StarLayout L;
    auto pad_to_chunk = [&]() {
        while (L.slot_src.size() % CHUNK_SLOTS) {
            if (L.slot_src.size() % 4 == 0) {
                const size_t s = L.slot_src.size();
                L.start_mask[s / CHUNK_SLOTS] |= 1u << ((s % CHUNK_SLOTS) / 4);
            }
            L.slot_src.push_back(0);  // gene 0 against itself: ties
        }
    };
    auto emit_hub = [&](int hub) {
        const size_t s = L.slot_src.size();
        if (s % CHUNK_SLOTS == 0) L.start_mask.push_back(0u);
        L.start_mask[s / CHUNK_SLOTS] |= 1u << ((s % CHUNK_SLOTS) / 4);
        L.slot_src.push_back(hub);
    };
    for (int side = 0; side < 2; ++side) {
        for (const Star &st : stars) {
            if (st.side != side) continue;
            emit_hub(st.hub);
            for (int o : st.others) {
                if (L.slot_src.size() % CHUNK_SLOTS == 0) emit_hub(st.hub);  // chunks are self-contained
                L.slot_src.push_back(o);
            }
            while (L.slot_src.size() % 4) L.slot_src.push_back(st.hub);
        }
        pad_to_chunk();
        if (side == 0) L.n_chunks_a = (int)(L.slot_src.size() / CHUNK_SLOTS);
    }
    L.n_chunks = (int)(L.slot_src.size() / CHUNK_SLOTS);
    return L;
}

__global__ void build_table_kernel(const uint16_t *__restrict__ perm, const int32_t *__restrict__ slot_src,
                                   const uint16_t *__restrict__ map, int n_used, int n_slots, int k_base,
                                   uint32_t *__restrict__ E) {
    const int k = blockIdx.y;  // row of this pass: 0 = identity, k > 0 = permutation k_base + k (1-based)
    const uint16_t *pk = perm + (size_t)(k > 0 ? k_base + k - 1 : 0) * n_used;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += gridDim.x * blockDim.x) {
        int g = slot_src[s];
        if (k > 0) g = pk[g];
        E[(size_t)k * n_slots + s] = (uint32_t)map[g] * 128u;
    }
}

// ------------------------------------------------------------------------------------------
// score kernel
// ------------------------------------------------------------------------------------------
constexpr int SCORE_THREADS = 512;
constexpr int SCORE_WARPS = SCORE_THREADS / 32;

struct ClassDesc {
    const uint32_t *E;           // [1 + permutations of this pass][n_slots], row 0 = identity
    const uint32_t *start_mask;  // [n_chunks]
    int32_t n_slots, n_chunks, n_chunks_a, pad;
};

struct ScoreParams {
    const uint16_t *rank;  // [n_cells][rank_ld]
    int64_t rank_ld;
    int64_t n_cells;  // cells of this launch
    int64_t out_ld;   // cells of the whole call (row length of scores / obs_*)
    int64_t out_off;  // first cell of this launch within the call
    int n_union;
    int n_classes;
    int iterations, min_iter, min_pairs;
    const ClassDesc *classes;
    const uint32_t *gtile;     // the launch's ranks as [64-cell tiles][n_union][32] words, two 15-bit ranks each (pack_tiles_kernel)
    const uint32_t *gtile_b8;  // as [128-cell tiles][n_union][32] words, four 7-bit ranks each (pack_tiles_b8_kernel)
    const int32_t *class_order;  // classes by decreasing cost: work unit u = (class_order[u / n_tiles], tile u % n_tiles)
    int *unit_counter;           // dynamic work-unit counters (zeroed before the launch): [0] 64-cell tiles, [1] MODE_B8
    int n_slices;                // > 1: the iterations of a (class, tile) are split over n_slices work units
    int slice_len;               // iterations per slice (a multiple of the warps per CTA)
    int n_slices_b8, slice_len_b8;  // the same for the 128-cell tiles of MODE_B8
    const int32_t *max_rank;     // largest dense rank of any cell ranked so far (flags[0] of the rank kernel)
    int k_count;                 // permutations held by the tables of this launch (rows 1 .. k_count of E)
    int accumulate;              // 1: add the partial counts to `below` (slices and / or several table passes)
    int32_t *below;              // accumulate only: [n_classes][out_ld] partial counts, finished by finalize_kernel
    double *scores;     // [n_classes][n_cells]
    int32_t *obs_hits;  // [n_classes][n_cells]
    int32_t *obs_total;
};

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// Three tile representations (template parameter MODE):
//  MODE_F16 (all marker genes of the call fit the shared-memory tile, <= 1 680 genes): ranks are integer-valued
//        fp16 numbers (exact up to 2048), two cells per 32-bit word = 64 cells per tile, tile in shared memory.
//        Per slot two statistics are taken against the hub h, on two different pipes:
//          X: (h >= v) as a 0xFFFF/0 mask per half   -- HSET2.GE (ALU pipe) + one 32-bit integer accumulate
//          Z: (h >  v) as clamp(h - v, 0, 1) in fp16 -- HADD2.SAT + HADD2 (FMA pipe)
//  MODE_B8 (same tile, and no cell ranked so far holds 128 or more distinct values among its marker genes -- count
//        data): ranks are 7-bit integers, FOUR cells per word = 128 cells per tile.  With H1 = h | 0x80808080 and
//        H2 = H1 - 0x01010101, bit 7 of every byte of H1 - v is (h >= v) and of H2 - v is (h > v) (no borrow crosses
//        a byte: every byte stays >= 0); PRMT with sign replication turns the four bits into 0xFF/0 bytes, which are
//        subtracted from a packed accumulator: 2 IADD + 2 PRMT + 1 IADD3 per slot for both statistics of 4 cells.
//  MODE_GLOBAL (up to 32 768 marker genes): ranks are 15-bit integers, two per word, tiles in global memory in the
//        same [gene][lane] layout (pack_tiles_kernel; L1/L2 cached gathers, shared by all CTAs working on the
//        tile); (x >= y) per half is bit 15/31 of (x | 0x80008000) - y.
// All flush their per-chunk counters into 32-bit per-cell counters, so a class may have any number of pairs.
//  MODE_GLOBAL_B8: the 7-bit arithmetic of MODE_B8 on global-memory tiles (one 128 B line per gather now serves 128 cells).
enum { MODE_GLOBAL = 0, MODE_F16 = 1, MODE_B8 = 2, MODE_GLOBAL_B8 = 3 };
constexpr int B8_MAX_RANK = 127;
template <int MODE> struct Mode {
    static constexpr bool BYTES = MODE == MODE_B8 || MODE == MODE_GLOBAL_B8;  // 7-bit ranks, four cells per word
    static constexpr int NC = BYTES ? 4 : 2;                                  // cells per lane
    static constexpr int CELLS = 32 * NC;                                     // cells per tile
    static constexpr bool SMEM = MODE == MODE_F16 || MODE == MODE_B8;         // tile in shared memory
};
constexpr int TILE_CELLS = 64;  // MODE_F16 / MODE_GLOBAL

__device__ __forceinline__ __half2 as_h2(uint32_t x) { return *reinterpret_cast<const __half2 *>(&x); }
__device__ __forceinline__ uint32_t ge_mask(uint32_t x, uint32_t y) { return __hge2_mask(as_h2(x), as_h2(y)); }
__device__ __forceinline__ uint32_t ge_bits(uint32_t x, uint32_t y) {
    return (((x | 0x80008000u) - y) >> 15) & 0x00010001u;
}
__device__ __forceinline__ uint32_t sign_bytes(uint32_t d) {  // 0xFF in every byte whose bit 7 is set, else 0x00
    uint32_t m;
    asm("prmt.b32 %0, %1, 0, 0xba98;" : "=r"(m) : "r"(d));  // selector msb = replicate the sign of the selected byte
    return m;
}
template <int MODE> __device__ __forceinline__ uint32_t tile_load(uint64_t base, uint32_t off) {
    if (Mode<MODE>::SMEM) return lds32((uint32_t)base + off);
    return __ldg(reinterpret_cast<const uint32_t *>(base + off));  // written by an earlier kernel
}

template <int NC> struct RowAcc {  // per-cell counters of one table row: [section A / B][cell of the lane]
    int x[2][NC], z[2][NC];
};

// For the cells packed in this lane add, over the slots of chunks first, first + stride, ... < last of one
// table row, X += (hub >= v) and Z += (hub > v).  Chunks are staged into this warp's private double buffer with
// cp.async; one broadcast LDS.128 fetches four slot offsets.
template <int MODE>
__device__ __forceinline__ void eval_chunks(uint64_t tile_lane, const uint32_t *__restrict__ Erow,
                                            const uint32_t *__restrict__ start_mask, int first, int stride, int last,
                                            uint4 *stage, int lane, int (&x)[Mode<MODE>::NC], int (&z)[Mode<MODE>::NC]) {
    if (first >= last) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(Erow) + lane;  // 32 lanes x 16 B = one chunk
    cp_async16(stage + lane, src + (size_t)first * 32);
    cp_async_commit();
    int buf = 0;
    uint32_t h = 0;
    for (int c = first; c < last; c += stride) {
        const uint32_t sm = __ldg(start_mask + c);
        if (c + stride < last) {
            cp_async16(stage + (buf ^ 1) * 32 + lane, src + (size_t)(c + stride) * 32);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const uint4 *sb = stage + buf * 32;
        if (MODE == MODE_F16) {
            uint32_t X = 0;                               // sum of -(mask), decoded below
            __half2 z0 = __float2half2_rn(0.f), z1 = z0;  // <= 64 each per chunk: exact
#pragma unroll 8
            for (int q = 0; q < 32; ++q) {
                const uint4 v = sb[q];  // broadcast: four slots
                const uint32_t e0 = tile_load<MODE>(tile_lane, v.x), e1 = tile_load<MODE>(tile_lane, v.y);
                const uint32_t e2 = tile_load<MODE>(tile_lane, v.z), e3 = tile_load<MODE>(tile_lane, v.w);
                if ((sm >> q) & 1u) h = e0;  // warp-uniform: this vector opens a star
                const __half2 hh = as_h2(h);
                X -= ge_mask(h, e0);
                X -= ge_mask(h, e1);
                X -= ge_mask(h, e2);
                X -= ge_mask(h, e3);
                z0 = __hadd2(z0, __hsub2_sat(hh, as_h2(e0)));
                z1 = __hadd2(z1, __hsub2_sat(hh, as_h2(e1)));
                z0 = __hadd2(z0, __hsub2_sat(hh, as_h2(e2)));
                z1 = __hadd2(z1, __hsub2_sat(hh, as_h2(e3)));
            }
            // a true low half adds 1 to the low field and -1 to the high field, a true high half adds 1 to the
            // high field  =>  lo16 = n_lo, hi16 = n_hi - n_lo (mod 2^16)
            const int n_lo = (int)(X & 0xffffu);
            x[0] += n_lo;
            x[1] += (int)(((X >> 16) + (uint32_t)n_lo) & 0xffffu);
            const __half2 zz = __hadd2(z0, z1);
            z[0] += __half2int_rz(__low2half(zz));
            z[1] += __half2int_rz(__high2half(zz));
        } else if (Mode<MODE>::BYTES) {
            // a true byte i adds 256^i - 256^(i+1) to the accumulator: X = N - (N << 8) with N the packed byte counts
            // (<= 128 per chunk), so N = X * 0x01010101 (mod 2^32)
            uint32_t X = 0, Z = 0, h1 = 0x80808080u, h2 = 0x7f7f7f7fu;
#pragma unroll(Mode<MODE>::SMEM ? 8 : 4)
            for (int q = 0; q < 32; ++q) {
                const uint4 v = sb[q];
                const uint32_t e0 = tile_load<MODE>(tile_lane, v.x), e1 = tile_load<MODE>(tile_lane, v.y);
                const uint32_t e2 = tile_load<MODE>(tile_lane, v.z), e3 = tile_load<MODE>(tile_lane, v.w);
                if ((sm >> q) & 1u) {
                    h1 = e0 | 0x80808080u;
                    h2 = h1 - 0x01010101u;
                }
                X = X - sign_bytes(h1 - e0) - sign_bytes(h1 - e1);
                X = X - sign_bytes(h1 - e2) - sign_bytes(h1 - e3);
                Z = Z - sign_bytes(h2 - e0) - sign_bytes(h2 - e1);
                Z = Z - sign_bytes(h2 - e2) - sign_bytes(h2 - e3);
            }
            const uint32_t nx = X * 0x01010101u, nz = Z * 0x01010101u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[i] += (int)((nx >> (8 * i)) & 0xffu);
                z[i] += (int)((nz >> (8 * i)) & 0xffu);
            }
        } else {
            uint32_t X = 0, W = 0;  // packed plain counts of (h >= v) and (v >= h); <= 128 per chunk
#pragma unroll 4
            for (int q = 0; q < 32; ++q) {
                const uint4 v = sb[q];
                const uint32_t e0 = tile_load<MODE>(tile_lane, v.x), e1 = tile_load<MODE>(tile_lane, v.y);
                const uint32_t e2 = tile_load<MODE>(tile_lane, v.z), e3 = tile_load<MODE>(tile_lane, v.w);
                if ((sm >> q) & 1u) h = e0;
                X += ge_bits(h, e0) + ge_bits(h, e1) + ge_bits(h, e2) + ge_bits(h, e3);
                W += ge_bits(e0, h) + ge_bits(e1, h) + ge_bits(e2, h) + ge_bits(e3, h);
            }
            x[0] += (int)(X & 0xffffu);
            x[1] += (int)(X >> 16);
            z[0] += CHUNK_SLOTS - (int)(W & 0xffffu);  // (h > v) = !(v >= h)
            z[1] += CHUNK_SLOTS - (int)(W >> 16);
        }
        __syncwarp();  // every lane is done with `buf` before it is refilled
        buf ^= 1;
    }
}

// One table row.  Section A: the hub is the pair's first gene; section B: its second gene.
template <int MODE>
__device__ __forceinline__ RowAcc<Mode<MODE>::NC> eval_row(uint64_t tile_lane, const uint32_t *__restrict__ Erow,
                                                         const ClassDesc &cd, int first, int stride, uint4 *stage,
                                                         int lane) {
    RowAcc<Mode<MODE>::NC> r;
#pragma unroll
    for (int i = 0; i < Mode<MODE>::NC; ++i) r.x[0][i] = r.x[1][i] = r.z[0][i] = r.z[1][i] = 0;
    eval_chunks<MODE>(tile_lane, Erow, cd.start_mask, first, stride, cd.n_chunks_a, stage, lane, r.x[0], r.z[0]);
    int fb = first;  // first chunk of this warp's stride pattern inside section B
    if (stride > 1) {
        while (fb < cd.n_chunks_a) fb += stride;
    } else {
        fb = cd.n_chunks_a;
    }
    eval_chunks<MODE>(tile_lane, Erow, cd.start_mask, fb, stride, cd.n_chunks, stage, lane, r.x[1], r.z[1]);
    return r;
}

// (hits, total) of get_proportion (cyclone.py:206-215) for cell i of a lane.  Over the slots of section A
// (pair = (h, v)) and B (pair = (v, h)), with ties = hub / padding slots and genuine ties:
//   hits  = #(first > second) = Z_A + (S_B - X_B)          ties are inside X_B, so they drop out
//   total = #(first != second) = S - (X_A - Z_A) - (X_B - Z_B)
template <int NC>
__device__ __forceinline__ void row_counts(const RowAcc<NC> &r, int i, int s_all, int s_b, int &hits, int &total) {
    hits = r.z[0][i] + s_b - r.x[1][i];
    total = s_all - r.x[0][i] + r.z[0][i] - r.x[1][i] + r.z[1][i];
}

// MODE_GLOBAL: ranks [cell][rank_ld] u16 -> [tile][gene][lane] words (lane l = cells l and l + 32 of the 64-cell tile).
// One CTA per (tile, 32-gene block): 64 coalesced 64-byte row reads, transposed through shared memory.
__global__ void __launch_bounds__(256) pack_tiles_kernel(const uint16_t *__restrict__ rank, int64_t rank_ld,
                                                         int64_t n_cells, int n_union, uint32_t *__restrict__ out,
                                                         const int32_t *__restrict__ max_rank) {
    __shared__ uint16_t s[TILE_CELLS][33];
    if (__ldg(max_rank) <= B8_MAX_RANK) return;  // the 7-bit kernels will run: nobody reads these tiles
    const int64_t cell0 = (int64_t)blockIdx.x * TILE_CELLS;
    const int g0 = blockIdx.y * 32, gl = threadIdx.x & 31;
    for (int c = threadIdx.x >> 5; c < TILE_CELLS; c += 8)
        s[c][gl] = (cell0 + c < n_cells && g0 + gl < n_union) ? rank[(cell0 + c) * rank_ld + g0 + gl] : (uint16_t)0;
    __syncthreads();
    uint32_t *o = out + ((size_t)blockIdx.x * n_union + g0) * 32;
    for (int g = threadIdx.x >> 5; g < 32 && g0 + g < n_union; g += 8)
        o[g * 32 + gl] = (uint32_t)s[gl][g] | ((uint32_t)s[gl + 32][g] << 16);
}

// MODE_GLOBAL_B8: the same with four cells per word (lane l = cells l, l + 32, l + 64, l + 96 of a 128-cell tile);
// only read when every rank is <= 127.
__global__ void __launch_bounds__(256) pack_tiles_b8_kernel(const uint16_t *__restrict__ rank, int64_t rank_ld,
                                                            int64_t n_cells, int n_union, uint32_t *__restrict__ out,
                                                            const int32_t *__restrict__ max_rank) {
    __shared__ uint8_t s[128][33];
    if (__ldg(max_rank) > B8_MAX_RANK) return;  // the two-cells-per-word kernels will run
    const int64_t cell0 = (int64_t)blockIdx.x * 128;
    const int g0 = blockIdx.y * 32, gl = threadIdx.x & 31;
    for (int c = threadIdx.x >> 5; c < 128; c += 8)
        s[c][gl] = (cell0 + c < n_cells && g0 + gl < n_union) ? (uint8_t)rank[(cell0 + c) * rank_ld + g0 + gl] : (uint8_t)0;
    __syncthreads();
    uint32_t *o = out + ((size_t)blockIdx.x * n_union + g0) * 32;
    for (int g = threadIdx.x >> 5; g < 32 && g0 + g < n_union; g += 8)
        o[g * 32 + gl] = (uint32_t)s[gl][g] | ((uint32_t)s[gl + 32][g] << 8) | ((uint32_t)s[gl + 64][g] << 16) |
                         ((uint32_t)s[gl + 96][g] << 24);
}

// total == iterations always (cyclone.py:238: NaN != -1); score per :243-249
__device__ __forceinline__ double final_score(uint32_t below, int iterations, int min_iter) {
    if (iterations >= min_iter && iterations > 0) return (double)below / (double)iterations;
    return nan("");
}

// (two CTAs per SM for the global-tile modes were measured slower: 28 vs 24 ms, two tiles thrash one L1)
template <int MODE>
__global__ void __launch_bounds__(SCORE_THREADS, 1) score_kernel(const ScoreParams P) {
    constexpr int NC = Mode<MODE>::NC, CELLS = Mode<MODE>::CELLS;
    constexpr bool SMEM = Mode<MODE>::SMEM;
    extern __shared__ __align__(128) uint8_t smem_dyn[];
    __shared__ int s_obs[4 * NC][32];  // observed counters: [x A, z A, x B, z B][cell of the lane]
    __shared__ uint32_t s_below[CELLS];
    __shared__ int s_unit;

    // the 7-bit path runs iff every cell ranked so far has at most 128 distinct values; the fp16 launch that
    // follows it on the stream takes the other case
    constexpr bool BYTES = Mode<MODE>::BYTES;
    {
        const bool small = __ldg(P.max_rank) <= B8_MAX_RANK;
        if (BYTES != small) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // shared-memory modes: [n_union][32] words in shared memory, followed by the per-warp staging buffers;
    // MODE_GLOBAL: the tiles stay in global memory, shared memory only holds the staging buffers
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem_dyn);
    uint4 *stage = reinterpret_cast<uint4 *>(smem_dyn + (SMEM ? (size_t)P.n_union * 128 : 0) + (size_t)warp * STAGE_BYTES);
    uint64_t tile_lane = (uint64_t)(smem_u32(tile) + lane * 4);
    const int64_t n_tiles = (P.n_cells + CELLS - 1) / CELLS;
    const int n_slices = BYTES ? P.n_slices_b8 : P.n_slices;
    const int slice_len = BYTES ? P.slice_len_b8 : P.slice_len;
    int *unit_counter = P.unit_counter + (BYTES ? 1 : 0);

    // Work units = (class, cell tile, iteration slice), handed out dynamically, most expensive class first:
    // equal-cost tiles would quantise the launch into whole waves (1 563 tiles on 148 SMs = 10.56 -> 11), units a
    // third of that size don't.  Few cells (fewer units than SMs) are split further into iteration slices; every
    // slice recomputes the observed row and adds its part of `below` to a global counter.
    const int64_t n_units = n_tiles * P.n_classes * n_slices;
    for (;;) {
        if (tid == 0) s_unit = atomicAdd(unit_counter, 1);
        __syncthreads();
        const int64_t u = s_unit;
        if (u >= n_units) break;
        const int slice = (int)(u % n_slices);
        const int64_t ct = u / n_slices;
        const int cls = P.class_order[ct / n_tiles];
        const int64_t cell0 = (ct % n_tiles) * CELLS;
        const int k_first = 1 + slice * slice_len;
        const int k_last = min(P.k_count, k_first + slice_len - 1);
        // the tile, already in [gene][lane] order (pack_tiles*_kernel): global-memory modes gather from it in place,
        // shared-memory modes copy it with coalesced 16-byte loads (a strided read of the rank rows cost ~100 us per
        // unit, which made small work units expensive)
        const uint32_t *gt = (BYTES ? P.gtile_b8 : P.gtile) + (size_t)(ct % n_tiles) * P.n_union * 32;
        if (!SMEM) {
            tile_lane = (uint64_t)(uintptr_t)(gt + lane);
        } else {
            const uint4 *src4 = reinterpret_cast<const uint4 *>(gt);
            uint4 *dst4 = reinterpret_cast<uint4 *>(tile);
            for (int i = tid; i < P.n_union * 8; i += SCORE_THREADS) {
                uint4 v = __ldg(src4 + i);
                if (!BYTES) {  // two 15-bit integers -> two fp16 numbers (exact: ranks < 2048)
                    uint32_t *w = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 h = __floats2half2_rn((float)(w[j] & 0xffffu), (float)(w[j] >> 16));
                        w[j] = *reinterpret_cast<const uint32_t *>(&h);
                    }
                }
                dst4[i] = v;
            }
        }
        {
            const ClassDesc cd = P.classes[cls];
            const int s_all = cd.n_slots, s_b = (cd.n_chunks - cd.n_chunks_a) * CHUNK_SLOTS;
            if (tid < 32)
                for (int j = 0; j < 4 * NC; ++j) s_obs[j][tid] = 0;
            if (tid < CELLS) s_below[tid] = 0;
            __syncthreads();  // also orders the tile stores before the first gather
            // ---- observed proportion (row 0): chunks dealt to the warps
            {
                const RowAcc<NC> r = eval_row<MODE>(tile_lane, cd.E, cd, warp, SCORE_WARPS, stage, lane);
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    atomicAdd(&s_obs[0 * NC + i][lane], r.x[0][i]);
                    atomicAdd(&s_obs[1 * NC + i][lane], r.z[0][i]);
                    atomicAdd(&s_obs[2 * NC + i][lane], r.x[1][i]);
                    atomicAdd(&s_obs[3 * NC + i][lane], r.z[1][i]);
                }
            }
            __syncthreads();
            int h0[NC], t0[NC];
            bool ok[NC];
            {
                RowAcc<NC> r;
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    r.x[0][i] = s_obs[0 * NC + i][lane];
                    r.z[0][i] = s_obs[1 * NC + i][lane];
                    r.x[1][i] = s_obs[2 * NC + i][lane];
                    r.z[1][i] = s_obs[3 * NC + i][lane];
                }
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    row_counts(r, i, s_all, s_b, h0[i], t0[i]);
                    ok[i] = !(h0[i] < P.min_pairs || t0[i] == 0);  // observed proportion is not NaN
                }
            }
            if (warp == 0 && slice == 0) {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const int64_t cell = cell0 + 32 * i + lane;
                    if (cell < P.n_cells) {
                        if (P.obs_hits) P.obs_hits[(int64_t)cls * P.out_ld + P.out_off + cell] = h0[i];
                        if (P.obs_total) P.obs_total[(int64_t)cls * P.out_ld + P.out_off + cell] = t0[i];
                    }
                }
            }
            // ---- permutation null: iterations dealt round-robin to the warps
            uint32_t below[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) below[i] = 0;
            for (int k = k_first + warp; k <= k_last; k += SCORE_WARPS) {
                const RowAcc<NC> r = eval_row<MODE>(tile_lane, cd.E + (size_t)k * s_all, cd, 0, 1, stage, lane);
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    int h, t;
                    row_counts(r, i, s_all, s_b, h, t);
                    // new < cur with NaN on either side comparing false (cyclone.py:217-218, 239)
                    if (ok[i] && !(h < P.min_pairs || t == 0) && (long long)h * t0[i] < (long long)h0[i] * t) ++below[i];
                }
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) atomicAdd(&s_below[32 * i + lane], below[i]);
            __syncthreads();
            if (tid < CELLS && cell0 + tid < P.n_cells) {
                const int64_t o = (int64_t)cls * P.out_ld + P.out_off + cell0 + tid;
                if (P.accumulate) {
                    if (s_below[tid]) atomicAdd(P.below + o, (int32_t)s_below[tid]);
                } else {
                    P.scores[o] = final_score(s_below[tid], P.iterations, P.min_iter);
                }
            }
            __syncthreads();
        }
    }
}

// Accumulating calls: scores of cells [out_off, out_off + n_cells) of every class from the summed partial counts.
__global__ void finalize_kernel(const ScoreParams P) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_cells * P.n_classes) return;
    const int64_t o = (i / P.n_cells) * P.out_ld + P.out_off + i % P.n_cells;
    P.scores[o] = final_score((uint32_t)P.below[o], P.iterations, P.min_iter);
}

struct PermKey {
    int mode;
    uint64_t seed;
    int cls;
    int n;
    int iters;
    bool operator<(const PermKey &o) const {
        if (mode != o.mode) return mode < o.mode;
        if (seed != o.seed) return seed < o.seed;
        if (cls != o.cls) return cls < o.cls;
        if (n != o.n) return n < o.n;
        return iters < o.iters;
    }
};
// MT19937 tables depend on (seed, n, iters) only and cost ~10 ms of serial host time each: keep the most recently
// used ones, at most PERM_CACHE_BYTES in total (a long-lived process with varying marker sets must not grow for ever).
// Callers hold a shared_ptr, so an entry evicted by another context's call stays alive while it is in use.
typedef std::shared_ptr<const std::vector<uint16_t>> PermTable;
constexpr size_t PERM_CACHE_BYTES = (size_t)256 << 20;
struct PermEntry {
    PermTable table;
    uint64_t last_use;
};
std::map<PermKey, PermEntry> g_perm_cache;
size_t g_perm_bytes = 0;
uint64_t g_perm_clock = 0;
std::mutex g_perm_mutex;  // contexts are single-threaded, but two contexts may run in two threads

PermTable mt_table_cached(uint32_t seed, int n, int iters) {
    const PermKey key{PP_RNG_MT19937, seed, 0, n, iters};
    {
        std::lock_guard<std::mutex> lock(g_perm_mutex);
        auto it = g_perm_cache.find(key);
        if (it != g_perm_cache.end()) {
            it->second.last_use = ++g_perm_clock;
            return it->second.table;
        }
    }
    auto v = std::make_shared<std::vector<uint16_t>>((size_t)n * iters);
    mt_perm_table(seed, n, iters, v->data());
    const size_t bytes = v->size() * sizeof(uint16_t);
    std::lock_guard<std::mutex> lock(g_perm_mutex);
    while (!g_perm_cache.empty() && g_perm_bytes + bytes > PERM_CACHE_BYTES) {  // evict the least recently used
        auto lru = g_perm_cache.begin();
        for (auto it = g_perm_cache.begin(); it != g_perm_cache.end(); ++it)
            if (it->second.last_use < lru->second.last_use) lru = it;
        g_perm_bytes -= lru->second.table->size() * sizeof(uint16_t);
        g_perm_cache.erase(lru);
    }
    if (bytes <= PERM_CACHE_BYTES && g_perm_cache.find(key) == g_perm_cache.end()) {
        g_perm_cache[key] = PermEntry{v, ++g_perm_clock};
        g_perm_bytes += bytes;
    }
    return v;
}

// Worker threads of one pp_cyclone call (spawning 16 threads per chunk cost ~0.5 ms per chunk).  run(fn) executes
// fn(part, n_parts) on every member -- the calling thread is member 0 -- and returns when all parts are done.
class HostCrew {
  public:
    explicit HostCrew(int n) : n_(1) {
        try {  // a thread that cannot be created just makes the crew smaller
            for (int t = 1; t < std::max(1, n); ++t) {
                threads_.emplace_back([this, t] { loop(t); });
                n_ = t + 1;
            }
        } catch (const std::exception &) {
        }
    }
    ~HostCrew() {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
            ++gen_;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void run(const std::function<void(int, int)> &fn) {
        {
            std::lock_guard<std::mutex> l(m_);
            fn_ = &fn;
            pending_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        fn(0, n_);
        std::unique_lock<std::mutex> l(m_);
        done_cv_.wait(l, [&] { return pending_ == 0; });
    }

  private:
    void loop(int t) {
        uint64_t seen = 0;
        for (;;) {
            const std::function<void(int, int)> *f;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                f = fn_;
            }
            (*f)(t, n_);
            std::lock_guard<std::mutex> l(m_);
            if (--pending_ == 0) done_cv_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> threads_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(int, int)> *fn_ = nullptr;
    uint64_t gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};

// Host-side gene filtering (the north-star keeps it on the host): copy the marker-gene columns of rows
// [r0, r0 + nr) of a host matrix into a dense [nr, n_cols_kept] pinned buffer, rows split over the crew.
template <typename T>
void gather_rows(const T *x, int64_t ld, int64_t r0, int64_t nr, const int32_t *cols, int n_kept, T *dst, HostCrew &crew) {
    // four rows at a time: four independent miss streams per thread (profiles/hostbench/gather_bench.cpp on the
    // 16-core box, 100k x 20k f32, 1 479 columns: 65 ms row by row, 49 ms interleaved; a streaming read takes 47 ms)
    const std::function<void(int, int)> work = [=](int part, int n_parts) {
        const int64_t a = nr * part / n_parts, b = nr * (part + 1) / n_parts;
        int64_t r = a;
        for (; r + 4 <= b; r += 4) {
            const T *s0 = x + (r0 + r) * ld, *s1 = s0 + ld, *s2 = s1 + ld, *s3 = s2 + ld;
            T *d0 = dst + r * n_kept, *d1 = d0 + n_kept, *d2 = d1 + n_kept, *d3 = d2 + n_kept;
            for (int j = 0; j < n_kept; ++j) {
                const int64_t c = cols[j];
                d0[j] = s0[c];
                d1[j] = s1[c];
                d2[j] = s2[c];
                d3[j] = s3[c];
            }
        }
        for (; r < b; ++r) {
            const T *src = x + (r0 + r) * ld;
            T *d = dst + r * n_kept;
            for (int j = 0; j < n_kept; ++j) d[j] = src[cols[j]];
        }
    };
    crew.run(work);
}

// Zero-copy lane: the marker columns of rows [r0, r0 + nr) of a MAPPED host matrix -> dense [nr, n_kept] device buffer.
// One warp per row, eight loads in flight per lane; every load is a 32-byte sector read over PCIe, so the kernel is
// bound by the link's small-read rate, not by the SMs (it runs beside the score kernel's persistent CTAs).
constexpr int ZC_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(ZC_WARPS * 32)
zc_gather_kernel(const T *__restrict__ x, int64_t ld, int64_t r0, int64_t nr, const int32_t *__restrict__ cols, int n_kept,
                 T *__restrict__ dst) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * ZC_WARPS + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * ZC_WARPS;
    for (int64_t r = warp; r < nr; r += n_warps) {
        const T *s = x + (r0 + r) * ld;
        T *d = dst + r * n_kept;
        for (int j0 = 0; j0 < n_kept; j0 += 32 * 8) {
            T v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 32 + lane;
                v[u] = j < n_kept ? __ldcs(s + __ldg(cols + j)) : (T)0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 32 + lane;
                if (j < n_kept) d[j] = v[u];
            }
        }
    }
}

int host_threads() { return pp_host_threads(); }

}  // namespace

PP_API int pp_perm_table(pp_ctx *ctx, int32_t rng_mode, uint64_t seed, int32_t class_index, int32_t n,
                         int32_t iterations, int32_t *out_table) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    if (n < 0 || n > 65535 || iterations < 0 || !out_table) return pp_fail(ctx, PP_EINVAL, "pp_perm_table: bad argument");
    const size_t cnt = (size_t)n * iterations;
    if (cnt == 0) return PP_OK;
    auto body = [&]() -> int {
    if (rng_mode == PP_RNG_MT19937) {
        const PermTable t = mt_table_cached((uint32_t)seed, n, iterations);
        for (size_t i = 0; i < cnt; ++i) out_table[i] = (*t)[i];
        return PP_OK;
    }
    if (rng_mode != PP_RNG_PHILOX) return pp_fail(ctx, PP_EINVAL, "pp_perm_table: rng_mode must be MT19937 or PHILOX");
    PP_CUDA(cudaSetDevice(ctx->device));
    DevBuf d;
    PP_CUDA(d.alloc(cnt * 2, ctx->stream));
    philox_perm_kernel<<<(iterations + 63) / 64, 64, 0, ctx->stream>>>(d.as<uint16_t>(), n, iterations, (uint32_t)seed,
                                                                      (uint32_t)(seed >> 32), (uint32_t)class_index);
    PP_CHECK_LAUNCH();
    std::vector<uint16_t> h(cnt);
    PP_CUDA(cudaMemcpyAsync(h.data(), d.p, cnt * 2, cudaMemcpyDeviceToHost, ctx->stream));
    PP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < cnt; ++i) out_table[i] = h[i];
    return PP_OK;
    };
    PP_GUARD(ctx, body());
}

// pp_cyclone_dist: the score columns of all ranks (cyclone.py:151-170) through ONE all-gather on the ctx stream.
// Every rank contributes a block [n_classes][max shard] of doubles (+ the two observed-counter arrays when asked
// for); receiving ranks copy the gathered blocks to the host and lay them out as [n_classes][all cells], shards in
// rank order.  The NaN flag is max-reduced first so that a NaN on one rank fails the call on every rank.
static int cyclone_gather(pp_ctx *ctx, const double *d_scores, const int32_t *d_oh, const int32_t *d_ot, int32_t *d_flags,
                          int64_t n_cells, int32_t n_classes, const int64_t *shard_cells, int32_t gather_to,
                          double *out_scores, int32_t *out_obs_hits, int32_t *out_obs_total, bool want_oh, bool want_ot) {
    cudaStream_t st = ctx->stream;
    const int wd = ctx->comm_world, me = ctx->comm_rank;
    const bool receive = gather_to < 0 || gather_to == me;
    int rc = pp_allreduce_max(ctx, d_flags, 4, ppn::ncclInt32);
    if (rc) return rc;
    int64_t maxn = 0, n_total = 0;
    for (int q = 0; q < wd; ++q) {
        maxn = std::max(maxn, shard_cells[q]);
        n_total += shard_cells[q];
    }
    const size_t sc_blk = sizeof(double) * (size_t)n_classes * maxn, ob_blk = sizeof(int32_t) * (size_t)n_classes * maxn;
    const size_t blk = sc_blk + (want_oh ? ob_blk : 0) + (want_ot ? ob_blk : 0);
    int32_t flags[4] = {0, 0, 0, 0};
    uint8_t *gbuf = nullptr;
    if (blk) {
        if (!(gbuf = (uint8_t *)pp_slab(ctx, SLAB_OUT, blk * wd))) return PP_ENOMEM;
        uint8_t *mine = gbuf + blk * me;
        if (n_cells > 0) {
            PP_CUDA(cudaMemcpy2DAsync(mine, maxn * sizeof(double), d_scores, n_cells * sizeof(double), n_cells * sizeof(double),
                                      n_classes, cudaMemcpyDeviceToDevice, st));
            if (want_oh)
                PP_CUDA(cudaMemcpy2DAsync(mine + sc_blk, maxn * sizeof(int32_t), d_oh, n_cells * sizeof(int32_t),
                                          n_cells * sizeof(int32_t), n_classes, cudaMemcpyDeviceToDevice, st));
            if (want_ot)
                PP_CUDA(cudaMemcpy2DAsync(mine + sc_blk + (want_oh ? ob_blk : 0), maxn * sizeof(int32_t), d_ot,
                                          n_cells * sizeof(int32_t), n_cells * sizeof(int32_t), n_classes,
                                          cudaMemcpyDeviceToDevice, st));
        }
        rc = pp_allgather_bytes(ctx, mine, gbuf, blk);
        if (rc) return rc;
    }
    // receiving ranks: shards laid side by side on the device ([n_classes][all cells] per array), one copy to pinned
    // staging, then into the caller's arrays
    uint8_t *stage = nullptr;
    bool direct = false;
    const size_t sc_all = sizeof(double) * (size_t)n_classes * n_total, ob_all = sizeof(int32_t) * (size_t)n_classes * n_total;
    const size_t all_bytes = sc_all + (want_oh ? ob_all : 0) + (want_ot ? ob_all : 0);
    if (receive && all_bytes) {
        uint8_t *d_full = (uint8_t *)pp_slab(ctx, SLAB_OUT2, all_bytes);
        if (!d_full) return PP_ENOMEM;
        int64_t off = 0;
        for (int q = 0; q < wd; ++q) {
            const uint8_t *b = gbuf + blk * q;
            const int64_t nq = shard_cells[q];
            if (nq > 0) {
                PP_CUDA(cudaMemcpy2DAsync(d_full + sizeof(double) * off, n_total * sizeof(double), b, maxn * sizeof(double),
                                          nq * sizeof(double), n_classes, cudaMemcpyDeviceToDevice, st));
                if (want_oh)
                    PP_CUDA(cudaMemcpy2DAsync(d_full + sc_all + sizeof(int32_t) * off, n_total * sizeof(int32_t), b + sc_blk,
                                              maxn * sizeof(int32_t), nq * sizeof(int32_t), n_classes, cudaMemcpyDeviceToDevice, st));
                if (want_ot)
                    PP_CUDA(cudaMemcpy2DAsync(d_full + sc_all + (want_oh ? ob_all : 0) + sizeof(int32_t) * off,
                                              n_total * sizeof(int32_t), b + sc_blk + (want_oh ? ob_blk : 0), maxn * sizeof(int32_t),
                                              nq * sizeof(int32_t), n_classes, cudaMemcpyDeviceToDevice, st));
            }
            off += nq;
        }
        direct = pp_is_pinned_host(out_scores) && (!want_oh || pp_is_pinned_host(out_obs_hits)) &&
                 (!want_ot || pp_is_pinned_host(out_obs_total));
        if (direct) {  // the caller's arrays are pinned (pp_alloc_result): no staging
            PP_CUDA(cudaMemcpyAsync(out_scores, d_full, sc_all, cudaMemcpyDeviceToHost, st));
            if (want_oh) PP_CUDA(cudaMemcpyAsync(out_obs_hits, d_full + sc_all, ob_all, cudaMemcpyDeviceToHost, st));
            if (want_ot)
                PP_CUDA(cudaMemcpyAsync(out_obs_total, d_full + sc_all + (want_oh ? ob_all : 0), ob_all, cudaMemcpyDeviceToHost, st));
        } else {
            if (ctx->pin_out_bytes < all_bytes) {
                if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
                ctx->pin_out = nullptr;
                ctx->pin_out_bytes = 0;
                const size_t want = std::max<size_t>(all_bytes * 2, (size_t)8 << 20);
                PP_CUDA(cudaHostAlloc(&ctx->pin_out, want, cudaHostAllocDefault));
                ctx->pin_out_bytes = want;
            }
            stage = (uint8_t *)ctx->pin_out;
            PP_CUDA(cudaMemcpyAsync(stage, d_full, all_bytes, cudaMemcpyDeviceToHost, st));
        }
    }
    PP_CUDA(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaEventRecord(ctx->ev[5], st));
    PP_CUDA(cudaStreamSynchronize(st));
    if (flags[1]) return pp_fail(ctx, PP_ENAN, "pp_cyclone: NaN in the expression matrix");
    if (stage) {
        if (!out_scores) return pp_fail(ctx, PP_EINVAL, "pp_cyclone_dist: out_scores is NULL on a receiving rank");
        // fresh pages of a large result: a few threads take the page faults in parallel
        struct Part { void *dst; const uint8_t *src; size_t n; };
        const Part parts[3] = {{out_scores, stage, sc_all},
                               {want_oh ? out_obs_hits : nullptr, stage + sc_all, ob_all},
                               {want_ot ? out_obs_total : nullptr, stage + sc_all + (want_oh ? ob_all : 0), ob_all}};
        for (const Part &pt : parts) {
            if (!pt.dst) continue;
            const int n_thr = pt.n >= ((size_t)4 << 20) ? 4 : 1;
            std::vector<std::thread> pool;
            for (int t = 1; t < n_thr; ++t)
                pool.emplace_back([=] { memcpy((uint8_t *)pt.dst + pt.n * t / n_thr, pt.src + pt.n * t / n_thr, pt.n * (t + 1) / n_thr - pt.n * t / n_thr); });
            memcpy(pt.dst, pt.src, pt.n / n_thr);
            for (auto &th : pool) th.join();
        }
    }
    return PP_OK;
}

// Shared body of pp_cyclone (dense x, csr == NULL), pp_cyclone_csr (csr != NULL, x / ld unused) and their _dist forms
// (shard_cells != NULL: the scores of all ranks are all-gathered, see cyclone_gather).
static int cyclone_impl(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, const pp_csr *csr, int64_t n_rows,
                        int64_t n_cols, int64_t ld, int64_t row_begin, int64_t n_cells, int32_t n_classes,
                        const int32_t *used_idx, const int64_t *used_off, const int32_t *pairs, const int64_t *pair_off,
                        int32_t iterations, int32_t min_iter, int32_t min_pairs, int32_t rng_mode, uint64_t seed,
                        const int32_t *perm_tables, const int64_t *shard_cells, int32_t gather_to, bool want_oh,
                        bool want_ot, double *out_scores, int32_t *out_obs_hits, int32_t *out_obs_total) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    const bool dist = shard_cells != nullptr;
    if (dist) {
        if (!ctx->comm) return pp_fail(ctx, PP_ENCCL, "pp_cyclone_dist: the context has no communicator (pp_comm_init)");
        if (gather_to >= ctx->comm_world) return pp_fail(ctx, PP_EINVAL, "pp_cyclone_dist: gather_to %d of %d ranks", gather_to, ctx->comm_world);
        if (shard_cells[ctx->comm_rank] != n_cells) return pp_fail(ctx, PP_EINVAL, "pp_cyclone_dist: shard_cells[rank] != n_cells");
        for (int q = 0; q < ctx->comm_world; ++q)
            if (shard_cells[q] < 0) return pp_fail(ctx, PP_EINVAL, "pp_cyclone_dist: negative shard size");
    }
    const bool receive = !dist || gather_to < 0 || gather_to == ctx->comm_rank;
    const bool empty_shard = dist && n_cells == 0 && n_rows == 0;  // a rank without cells still joins the collectives
    if ((!x && !csr && !empty_shard) || !used_idx || !used_off || !pairs || !pair_off || (receive && !out_scores))
        return pp_fail(ctx, PP_EINVAL, "pp_cyclone: NULL argument");
    if (x_dtype != PP_F32 && x_dtype != PP_F64) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: x_dtype must be PP_F32/PP_F64");
    if (!empty_shard && (n_rows <= 0 || n_cols <= 0 || (!csr && ld < n_cols)))
        return pp_fail(ctx, PP_EINVAL, "pp_cyclone: bad matrix shape");
    if (csr && !empty_shard) {
        const int vrc = pp_csr_validate(ctx, csr, n_rows, n_cols, "pp_cyclone_csr");
        if (vrc) return vrc;
    }
    if (row_begin < 0 || n_cells < 0 || row_begin + n_cells > n_rows) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: bad row range");
    if (n_classes < 1) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: n_classes < 1");
    if (iterations < 0) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: iterations < 0");
    if (rng_mode < PP_RNG_MT19937 || rng_mode > PP_RNG_TABLE) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: bad rng_mode");
    if (rng_mode == PP_RNG_TABLE && !perm_tables) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: PP_RNG_TABLE needs perm_tables");
    ctx->tm = pp_timings();
    ctx->launches = 0;
    ctx->tm.n_units = n_cells;
    // PP_TRACE=1: host-side time line of the call (ms since entry at every mark), printed at the end
    const bool tl_on = getenv("PP_TRACE") != nullptr;
    const auto tl_entry = std::chrono::steady_clock::now();
    std::string tl_line;
    auto tl_mark = [&](const char *what) {
        if (!tl_on) return;
        char buf[64];
        snprintf(buf, sizeof(buf), " %s=%.3f", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tl_entry).count());
        tl_line += buf;
    };
    if (n_cells == 0 && !dist) return PP_OK;
    if (n_cells == 0) {  // an empty shard still takes part in the collectives
        PP_CUDA(cudaSetDevice(ctx->device));
        DevBuf d_f;
        PP_CUDA(d_f.alloc(sizeof(int32_t) * 4, ctx->stream));
        PP_CUDA(cudaMemsetAsync(d_f.p, 0, sizeof(int32_t) * 4, ctx->stream));
        return cyclone_gather(ctx, nullptr, nullptr, nullptr, d_f.as<int32_t>(), 0, n_classes, shard_cells, gather_to,
                              out_scores, out_obs_hits, out_obs_total, want_oh, want_ot);
    }

    // ---- union of used genes, per-class maps
    std::vector<int32_t> uni;
    for (int c = 0; c < n_classes; ++c) {
        const int64_t u0 = used_off[c], u1 = used_off[c + 1], p0 = pair_off[c], p1 = pair_off[c + 1];
        if (u1 < u0 || p1 < p0) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: offsets must be non-decreasing");
        if (p1 - p0 == 0) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: class %d has no marker pairs", c);
        if (p1 - p0 > (1 << 28)) return pp_fail(ctx, PP_ELIMIT, "pp_cyclone: class %d has more than 2^28 marker pairs", c);
        for (int64_t i = u0; i < u1; ++i) {
            if (used_idx[i] < 0 || used_idx[i] >= n_cols) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: used_idx out of range");
            if (i > u0 && used_idx[i] <= used_idx[i - 1]) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: used_idx must ascend within a class");
            uni.push_back(used_idx[i]);
        }
        for (int64_t i = 2 * p0; i < 2 * p1; ++i)
            if (pairs[i] < 0 || pairs[i] >= u1 - u0) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: pair index out of range in class %d", c);
    }
    std::sort(uni.begin(), uni.end());
    uni.erase(std::unique(uni.begin(), uni.end()), uni.end());
    const int U = (int)uni.size();
    if (U > 32768) return pp_fail(ctx, PP_ELIMIT, "pp_cyclone: %d distinct marker genes (max 32768)", U);
    // shared-memory modes: the whole tile (128 B per gene) + staging buffers fit in shared memory and ranks < 2048
    const bool fast = U <= 2048 && (size_t)U * 128 + (size_t)SCORE_WARPS * STAGE_BYTES + 2048 <=
                                       (size_t)std::min(ctx->smem_optin, 227 * 1024);
    const size_t dyn_smem = (fast ? (size_t)U * 128 : 0) + (size_t)SCORE_WARPS * STAGE_BYTES;

    PP_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    PP_CUDA(cudaEventRecord(ctx->ev[0], st));
    const size_t esz = x_dtype == PP_F64 ? 8 : 4;
    DevBuf d_rows, d_uni, d_flags, d_scores, d_oh, d_ot, d_cd;
    void *s_rank = nullptr;  // ctx slab
    const uint8_t *x0 = csr ? nullptr : (const uint8_t *)x + (size_t)row_begin * ld * esz;
    // Host input is streamed in chunks of one full wave of 64-cell tiles: H2D of chunk i+1 (copy stream)
    // overlaps rank + score of chunk i (compute stream).  When the marker genes are at most a quarter of the
    // columns the host first gathers just those columns into a pinned staging buffer (gene filtering on the
    // host, multi-threaded), so PCIe carries U instead of n_cols values per cell.  Device input is processed
    // in one launch.
    const int64_t wave_cells = (int64_t)ctx->n_sm * TILE_CELLS;
    // (measured: with 16 host threads the gather beats full-row copies 161 vs 195 ms on cfg2; with 8 threads per
    // rank, 4 ranks sharing the host, it loses 402 vs 349 ms -- so it needs >= 12 threads for this process)
    // Three ways for a chunk of cells of a HOST matrix to reach the device (profiles/feed_probe.py, cfg2 = 100k x 20k
    // f32, 1 479 marker genes, library time of the call):
    //   gather    host threads copy the marker columns into pinned staging, DMA of [cells, U]
    //             16 threads: 82 ms (pinned or pageable input); 4 threads: 113 ms
    //   zerocopy  PINNED / registered matrices only: a gather kernel reads the marker columns straight from the mapped
    //             host memory (32-byte sectors over the GPU's own PCIe link: 1 133 sectors = 36 KB instead of 80 KB per
    //             row, no host core involved): 128 ms, whatever the host is doing
    //   copy      DMA of the full rows, columns picked by the rank kernel: 151 ms pinned, 710 ms pageable
    // With 8 ranks on one host (4 threads each, all at once; profiles/r02/feed_probe_n8_4threads.txt): gather 305-330 ms,
    // zerocopy 176-249 ms (the ranks whose GPU sits on the other socket of the matrix are the slow ones), copy 233-352 ms.
    // So: wide matrices go through the gather, except that a pinned matrix is read in place when this process has
    // fewer than 4 host threads or shares the host with three or more other ranks.  Running both lanes side by side
    // (PP_FEED=both, chunks to whichever lane is free) loses on one GPU -- 114 ms with 16 threads --: the two contend
    // for host memory.
    // PP_FEED = gather | zerocopy | both | copy forces a way (tests, experiments).
    const bool narrow = !csr && !x_is_device && (int64_t)U * 4 <= n_cols;
    const char *feed_env = getenv("PP_FEED");
    std::string feed = feed_env ? feed_env : "auto";
    const uint8_t *xz = nullptr;  // device-side address of x0 when the matrix is mapped host memory
    if (narrow && (feed == "auto" || feed == "zerocopy" || feed == "both") && n_cells > 0) {
        cudaPointerAttributes a0, a1;
        const uint8_t *last = x0 + ((size_t)(n_cells - 1) * ld + (size_t)n_cols) * esz - 1;
        if (cudaPointerGetAttributes(&a0, x0) == cudaSuccess && cudaPointerGetAttributes(&a1, last) == cudaSuccess &&
            a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost && a0.devicePointer && a1.devicePointer &&
            (const uint8_t *)a1.devicePointer - (const uint8_t *)a0.devicePointer == last - x0)
            xz = (const uint8_t *)a0.devicePointer;
        else
            cudaGetLastError();
    }
    if (feed == "auto") {
        // ranks of this job on this host (torchrun's LOCAL_WORLD_SIZE): from four on they saturate the host's memory with
        // their gathers (8 ranks x 4 threads, 100k cells each: 305-330 ms per call against 176-249 ms for the zero-copy reads)
        const char *lws = getenv("LOCAL_WORLD_SIZE");
        const int local_ranks = lws ? std::max(1, atoi(lws)) : 1;
        feed = !narrow ? "copy" : (xz && (host_threads() < 4 || local_ranks >= 4)) ? "zerocopy" : "gather";
    }
    if (!narrow) feed = "copy";
    if (!xz && (feed == "zerocopy" || feed == "both")) feed = "gather";  // not mapped memory
    const bool zc_lane = feed == "zerocopy" || feed == "both";
    const bool host_gather = feed == "gather" || feed == "both";
    const bool compact = host_gather || zc_lane;  // the device chunk buffers hold [cells, U]
    const bool pipelined = !csr && !x_is_device && (n_cells > wave_cells || compact);
    const int64_t chunk_cells = (pipelined || csr) ? std::min(wave_cells, n_cells) : n_cells;
    {
        std::vector<int32_t> rows(chunk_cells);
        for (int64_t i = 0; i < chunk_cells; ++i) rows[i] = (int32_t)i;
        PP_CUDA(d_rows.alloc(sizeof(int32_t) * chunk_cells, st));
        PP_CUDA(cudaMemcpyAsync(d_rows.p, rows.data(), sizeof(int32_t) * chunk_cells, cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaStreamSynchronize(st));  // rows is a local
    }
    PP_CUDA(d_uni.alloc(sizeof(int32_t) * U, st));
    PP_CUDA(cudaMemcpyAsync(d_uni.p, uni.data(), sizeof(int32_t) * U, cudaMemcpyHostToDevice, st));
    const int64_t rank_ld = (U + 7) & ~7;
    if (!(s_rank = pp_slab(ctx, SLAB_RANK, sizeof(uint16_t) * n_cells * rank_ld))) return PP_ENOMEM;
    PP_CUDA(d_flags.alloc(sizeof(int32_t) * 4, st));
    PP_CUDA(cudaMemsetAsync(d_flags.p, 0, sizeof(int32_t) * 4, st));
    int rc = PP_OK;
    // One-launch inputs (device matrix, or a host matrix of at most one wave of tiles): copy + rank go first, the
    // host-side table work below then overlaps them (tables are built on the copy stream).
    if (!csr && !pipelined) {
        const void *dx = x0;
        if (!x_is_device) {
            void *s_x = pp_slab(ctx, SLAB_X, (size_t)n_cells * ld * esz);
            if (!s_x) return PP_ENOMEM;
            PP_CUDA(cudaEventRecord(ctx->ev[6], st));
            PP_CUDA(cudaMemcpyAsync(s_x, x0, (size_t)n_cells * ld * esz, cudaMemcpyHostToDevice, st));
            ctx->tm.h2d_bytes += (int64_t)((size_t)n_cells * ld * esz);
            PP_CUDA(cudaEventRecord(ctx->ev[7], st));
            dx = s_x;
        }
        rc = pp_rank_launch(ctx, dx, x_dtype, ld, d_rows.as<int32_t>(), n_cells, d_uni.as<int32_t>(), U,
                            (uint16_t *)s_rank, rank_ld, d_flags.as<int32_t>());
        if (rc) return rc;
    }

    // ---- per-class permutation + pair tables, on the copy stream
    cudaStream_t ts = ctx->copy_stream;
    std::vector<DevBuf> d_T(n_classes), d_perm(n_classes), d_src(n_classes), d_mask(n_classes), d_map(n_classes);
    std::vector<ClassDesc> cds(n_classes);
    const int32_t *ext = perm_tables;
    std::vector<PermTable> held_tables;
    const bool trace = getenv("PP_TRACE") != nullptr;  // host-side phase times on stderr
    auto now_ms = [] {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    for (int c = 0; c < n_classes; ++c) {
        const double t_a = now_ms();
        const int nu = (int)(used_off[c + 1] - used_off[c]);
        const int np = (int)(pair_off[c + 1] - pair_off[c]);
        std::vector<uint16_t> map(nu > 0 ? nu : 1);
        for (int i = 0; i < nu; ++i)
            map[i] = (uint16_t)(std::lower_bound(uni.begin(), uni.end(), used_idx[used_off[c] + i]) - uni.begin());
        const size_t pcnt = (size_t)nu * iterations;
        PP_CUDA(d_perm[c].alloc(pcnt * 2, ts));
        if (rng_mode == PP_RNG_PHILOX) {
            if (pcnt) {
                philox_perm_kernel<<<(iterations + 63) / 64, 64, 0, ts>>>(d_perm[c].as<uint16_t>(), nu, iterations,
                                                                         (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)c);
                PP_CHECK_LAUNCH();
            }
        } else if (rng_mode == PP_RNG_MT19937) {
            held_tables.push_back(mt_table_cached((uint32_t)seed, nu, iterations));  // alive until the copy is done
            if (pcnt) PP_CUDA(cudaMemcpyAsync(d_perm[c].p, held_tables.back()->data(), pcnt * 2, cudaMemcpyHostToDevice, ts));
        } else {
            std::vector<uint16_t> t(pcnt);
            for (size_t i = 0; i < pcnt; ++i) {
                if (ext[i] < 0 || ext[i] >= nu) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: perm_tables entry out of range (class %d)", c);
                t[i] = (uint16_t)ext[i];
            }
            ext += pcnt;
            if (pcnt) PP_CUDA(cudaMemcpyAsync(d_perm[c].p, t.data(), pcnt * 2, cudaMemcpyHostToDevice, ts));
            PP_CUDA(cudaStreamSynchronize(ts));  // t is a local
        }
        const double t_b = now_ms();
        PP_CUDA(d_map[c].alloc(sizeof(uint16_t) * map.size(), ts));
        PP_CUDA(cudaMemcpyAsync(d_map[c].p, map.data(), sizeof(uint16_t) * map.size(), cudaMemcpyHostToDevice, ts));
        PP_CUDA(cudaStreamSynchronize(ts));  // map is a local
        const StarLayout L = build_star_layout(pairs + 2 * pair_off[c], np, nu);
        const int n_slots = (int)L.slot_src.size();
        if (trace)
            fprintf(stderr, "[pp_cyclone] class %d: %d genes, %d pairs -> %d slots; permutations %.1f ms, star layout %.1f ms\n",
                    c, nu, np, n_slots, t_b - t_a, now_ms() - t_b);
        PP_CUDA(d_src[c].alloc(sizeof(int32_t) * n_slots, ts));
        PP_CUDA(cudaMemcpyAsync(d_src[c].p, L.slot_src.data(), sizeof(int32_t) * n_slots, cudaMemcpyHostToDevice, ts));
        PP_CUDA(d_mask[c].alloc(sizeof(uint32_t) * L.n_chunks, ts));
        PP_CUDA(cudaMemcpyAsync(d_mask[c].p, L.start_mask.data(), sizeof(uint32_t) * L.n_chunks, cudaMemcpyHostToDevice, ts));
        PP_CUDA(cudaStreamSynchronize(ts));  // L is a local
        cds[c].E = nullptr;  // allocated below, once the size of a pass is known
        cds[c].start_mask = d_mask[c].as<uint32_t>();
        cds[c].n_slots = n_slots;
        cds[c].n_chunks = L.n_chunks;
        cds[c].n_chunks_a = L.n_chunks_a;
        cds[c].pad = 0;
    }
    // E holds one row per permutation: iterations x slots x 4 B (41 MB for the default markers, 16 GB for 4 M pairs).
    // It is kept under a third of the free device memory by scoring the iterations in passes of k_pass permutations.
    int k_pass = iterations;
    {
        size_t total_slots = 0, free_b = 0, total_b = 0;
        for (int c = 0; c < n_classes; ++c) total_slots += (size_t)cds[c].n_slots;
        size_t budget = (size_t)1 << 30;  // tables up to 1 GB are simply built; beyond that, ask the device
        const char *e = getenv("PP_TABLE_BUDGET_MB");  // tests
        if (e) {
            budget = (size_t)std::max(1, atoi(e)) << 20;
        } else if (total_slots * sizeof(uint32_t) * ((size_t)iterations + 1) > budget) {
            PP_CUDA(cudaMemGetInfo(&free_b, &total_b));
            budget = std::max(budget, free_b / 3);
        }
        const size_t rows = budget / (total_slots * sizeof(uint32_t));
        k_pass = (int)std::max<size_t>(1, std::min<size_t>((size_t)std::max(iterations, 1), rows > 1 ? rows - 1 : 1));
    }
    const int n_passes = iterations > 0 ? (iterations + k_pass - 1) / k_pass : 1;
    for (int c = 0; c < n_classes; ++c) {
        PP_CUDA(d_T[c].alloc(sizeof(uint32_t) * (size_t)(k_pass + 1) * cds[c].n_slots, ts));
        cds[c].E = d_T[c].as<uint32_t>();
    }
    auto build_tables = [&](int k_base, int k_count, cudaStream_t bs) -> int {
        for (int c = 0; c < n_classes; ++c) {
            const int nu = (int)(used_off[c + 1] - used_off[c]);
            dim3 grid((cds[c].n_slots + 255) / 256, k_count + 1);
            build_table_kernel<<<grid, 256, 0, bs>>>(d_perm[c].as<uint16_t>(), d_src[c].as<int32_t>(),
                                                     d_map[c].as<uint16_t>(), nu, cds[c].n_slots, k_base,
                                                     d_T[c].as<uint32_t>());
            PP_CHECK_LAUNCH();
        }
        return PP_OK;
    };
    rc = build_tables(0, std::min(k_pass, iterations), ts);
    if (rc) return rc;
    PP_CUDA(d_cd.alloc(sizeof(ClassDesc) * n_classes, ts));
    PP_CUDA(cudaMemcpyAsync(d_cd.p, cds.data(), sizeof(ClassDesc) * n_classes, cudaMemcpyHostToDevice, ts));
    const size_t out_cnt = (size_t)n_classes * n_cells;
    PP_CUDA(d_scores.alloc(sizeof(double) * out_cnt, st));
    if (want_oh) PP_CUDA(d_oh.alloc(sizeof(int32_t) * out_cnt, st));
    if (want_ot) PP_CUDA(d_ot.alloc(sizeof(int32_t) * out_cnt, st));
    PP_CUDA(cudaEventRecord(ctx->ev[1], ts));  // tables enqueued
    tl_mark("tables_enqueued");
    bool tables_awaited = false;

    ScoreParams P;
    P.rank_ld = rank_ld;
    P.out_ld = n_cells;
    P.n_union = U;
    P.n_classes = n_classes;
    P.iterations = iterations;
    P.min_iter = min_iter;
    P.min_pairs = min_pairs;
    P.classes = d_cd.as<ClassDesc>();
    P.scores = d_scores.as<double>();
    P.obs_hits = want_oh ? d_oh.as<int32_t>() : nullptr;
    P.obs_total = want_ot ? d_ot.as<int32_t>() : nullptr;
    PP_CUDA(cudaFuncSetAttribute(score_kernel<MODE_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    PP_CUDA(cudaFuncSetAttribute(score_kernel<MODE_B8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
    P.max_rank = d_flags.as<int32_t>();
    const int slow_grid = ctx->n_sm * 2;  // persistent CTAs of the global-tile path
    uint32_t *s_tiles = nullptr;          // packed tiles of one launch, 64-cell tiles then 128-cell tiles
    const size_t tiles64 = (size_t)((chunk_cells + TILE_CELLS - 1) / TILE_CELLS), tiles128 = (tiles64 + 1) / 2;
    if (!(s_tiles = (uint32_t *)pp_slab(ctx, SLAB_MISC, (tiles64 + tiles128) * U * 128))) return PP_ENOMEM;
    P.gtile = s_tiles;
    P.gtile_b8 = s_tiles + tiles64 * U * 32;
    // classes by decreasing cost (slots) + the dynamic work-unit counter
    DevBuf d_order, d_counter, d_below;
    {
        std::vector<int32_t> order(n_classes);
        for (int c = 0; c < n_classes; ++c) order[c] = c;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cds[a].n_slots > cds[b].n_slots; });
        PP_CUDA(d_order.alloc(sizeof(int32_t) * n_classes, st));
        PP_CUDA(cudaMemcpyAsync(d_order.p, order.data(), sizeof(int32_t) * n_classes, cudaMemcpyHostToDevice, st));
        PP_CUDA(d_counter.alloc(2 * sizeof(int), st));
        PP_CUDA(cudaStreamSynchronize(st));  // order is a local
    }
    P.class_order = d_order.as<int32_t>();
    P.unit_counter = d_counter.as<int>();
    // Launches with fewer (class, tile) units than ~8 per SM (few cells -- the reference's own use case -- and the chunks
    // of the host pipeline) split each unit's iterations into slices of >= 2 iterations per warp, aiming at ~12 units per
    // SM, so that the last, partly idle wave of units is short.  (A unit costs a tile copy of a few microseconds and one
    // extra evaluation of the observed row.)
    auto slices_for = [&](int64_t nc, int tile_cells, int k_count, int *slice_len) {
        *slice_len = std::max(k_count, 1);
        const int64_t units0 = (nc + tile_cells - 1) / tile_cells * n_classes;
        if (nc <= 0 || units0 >= 8 * (int64_t)ctx->n_sm || k_count < 4 * SCORE_WARPS) return 1;
        const int64_t want = (12 * (int64_t)ctx->n_sm + units0 - 1) / units0;
        const int64_t len = ((k_count + want - 1) / want + SCORE_WARPS - 1) / SCORE_WARPS * SCORE_WARPS;
        *slice_len = (int)std::max<int64_t>(len, 2 * SCORE_WARPS);
        return (k_count + *slice_len - 1) / *slice_len;
    };
    P.k_count = std::min(k_pass, iterations);
    {
        int len;
        const bool sliced = slices_for(n_cells, TILE_CELLS, P.k_count, &len) > 1 ||
                            slices_for(n_cells, Mode<MODE_B8>::CELLS, P.k_count, &len) > 1;
        P.accumulate = (n_passes > 1 || sliced || chunk_cells < n_cells) ? 1 : 0;  // chunked calls: launches of any size
        if (P.accumulate) {
            PP_CUDA(d_below.alloc(sizeof(int32_t) * out_cnt, st));
            PP_CUDA(cudaMemsetAsync(d_below.p, 0, sizeof(int32_t) * out_cnt, st));
        }
    }
    P.below = d_below.as<int32_t>();
    auto launch_score = [&](int64_t nc) {
        if (!tables_awaited) cudaStreamWaitEvent(st, ctx->ev[1], 0);
        tables_awaited = true;
        P.n_slices = slices_for(nc, TILE_CELLS, P.k_count, &P.slice_len);
        P.n_slices_b8 = slices_for(nc, Mode<MODE_B8>::CELLS, P.k_count, &P.slice_len_b8);
        const int64_t n_units = (nc + TILE_CELLS - 1) / TILE_CELLS * n_classes * P.n_slices;
        const int64_t n_units_b8 = (nc + Mode<MODE_B8>::CELLS - 1) / Mode<MODE_B8>::CELLS * n_classes * P.n_slices_b8;
        cudaMemsetAsync(d_counter.p, 0, 2 * sizeof(int), st);
        // count data (<= 128 distinct values per cell) takes the 4-cells-per-word kernel, anything else the 2-cell
        // kernel: both are enqueued, the device-side flag written by the rank kernel lets exactly one of them run
        ctx->launches += 3;
        const unsigned gy = (unsigned)((U + 31) / 32);
        pack_tiles_kernel<<<dim3((unsigned)((nc + 63) / 64), gy), 256, 0, st>>>(P.rank, rank_ld, nc, U, s_tiles, P.max_rank);
        pack_tiles_b8_kernel<<<dim3((unsigned)((nc + 127) / 128), gy), 256, 0, st>>>(
            P.rank, rank_ld, nc, U, s_tiles + tiles64 * U * 32, P.max_rank);
        if (fast) {
            score_kernel<MODE_B8><<<(unsigned)std::min<int64_t>(n_units_b8, ctx->n_sm), SCORE_THREADS, dyn_smem, st>>>(P);
            score_kernel<MODE_F16><<<(unsigned)std::min<int64_t>(n_units, ctx->n_sm), SCORE_THREADS, dyn_smem, st>>>(P);
        } else {
            score_kernel<MODE_GLOBAL_B8><<<(unsigned)std::min<int64_t>(n_units_b8, slow_grid), SCORE_THREADS, dyn_smem, st>>>(P);
            score_kernel<MODE_GLOBAL><<<(unsigned)std::min<int64_t>(n_units, slow_grid), SCORE_THREADS, dyn_smem, st>>>(P);
        }
    };
    // After every cell has been ranked and scored against the first pass of tables: the remaining passes over the
    // resident rank matrix, then the division for accumulating calls.
    auto finish_scoring = [&]() -> int {
        for (int pass = 1; pass < n_passes; ++pass) {
            const int k_base = pass * k_pass;
            P.k_count = std::min(k_pass, iterations - k_base);
            const int brc = build_tables(k_base, P.k_count, st);
            if (brc) return brc;
            for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells) {
                P.rank = (uint16_t *)s_rank + c0 * rank_ld;
                P.n_cells = std::min(chunk_cells, n_cells - c0);
                P.out_off = c0;
                launch_score(P.n_cells);
                PP_CHECK_LAUNCH();
            }
        }
        if (P.accumulate) {
            P.n_cells = n_cells;
            P.out_off = 0;
            finalize_kernel<<<(unsigned)((n_cells * n_classes + 255) / 256), 256, 0, st>>>(P);
            PP_CHECK_LAUNCH();
        }
        return PP_OK;
    };
    float score_ms = 0.f, copy_ms = 0.f;
    if (csr) {
        // CSR input: only the non-zeros of the cell shard are uploaded; every chunk of cells is densified on the
        // device straight into the compact [cells, U] marker-gene layout, then ranked and scored.
        DevCsr dcsr;
        PP_CUDA(cudaEventRecord(ctx->ev[6], st));
        rc = pp_csr_to_device(ctx, csr, row_begin, n_cells, &dcsr);
        if (rc) return rc;
        PP_CUDA(cudaEventRecord(ctx->ev[7], st));
        std::vector<int32_t> colmap((size_t)n_cols, -1), ident((size_t)U), allrows((size_t)n_cells);
        for (int j = 0; j < U; ++j) {
            colmap[uni[j]] = j;
            ident[j] = j;
        }
        for (int64_t i = 0; i < n_cells; ++i) allrows[i] = (int32_t)i;
        DevBuf d_colmap, d_ident, d_allrows;
        void *s_dense = pp_slab(ctx, SLAB_CHUNK0, (size_t)chunk_cells * U * esz);
        if (!s_dense) return PP_ENOMEM;
        PP_CUDA(d_colmap.alloc(sizeof(int32_t) * n_cols, st));
        PP_CUDA(cudaMemcpyAsync(d_colmap.p, colmap.data(), sizeof(int32_t) * n_cols, cudaMemcpyHostToDevice, st));
        PP_CUDA(d_ident.alloc(sizeof(int32_t) * U, st));
        PP_CUDA(cudaMemcpyAsync(d_ident.p, ident.data(), sizeof(int32_t) * U, cudaMemcpyHostToDevice, st));
        PP_CUDA(d_allrows.alloc(sizeof(int32_t) * n_cells, st));
        PP_CUDA(cudaMemcpyAsync(d_allrows.p, allrows.data(), sizeof(int32_t) * n_cells, cudaMemcpyHostToDevice, st));
        PP_CUDA(cudaStreamSynchronize(st));  // colmap / ident / allrows are locals
        PP_CUDA(cudaEventRecord(ctx->ev[3], st));
        for (int64_t c0 = 0; c0 < n_cells; c0 += chunk_cells) {
            const int64_t nc = std::min(chunk_cells, n_cells - c0);
            PP_CUDA(cudaMemsetAsync(s_dense, 0, (size_t)nc * U * esz, st));
            rc = pp_csr_densify(ctx, dcsr, d_allrows.as<int32_t>() + c0, nc, d_colmap.as<int32_t>(), s_dense, U);
            if (rc) return rc;
            rc = pp_rank_launch(ctx, s_dense, x_dtype, U, d_rows.as<int32_t>(), nc, d_ident.as<int32_t>(), U,
                                (uint16_t *)s_rank + c0 * rank_ld, rank_ld, d_flags.as<int32_t>());
            if (rc) return rc;
            P.rank = (uint16_t *)s_rank + c0 * rank_ld;
            P.n_cells = nc;
            P.out_off = c0;
            launch_score(nc);
            PP_CHECK_LAUNCH();
        }
        rc = finish_scoring();
        if (rc) return rc;
        PP_CUDA(cudaEventRecord(ctx->ev[4], st));
        PP_CUDA(cudaStreamSynchronize(st));  // dcsr goes out of scope
    } else if (!pipelined) {  // copied and ranked above
        PP_CUDA(cudaStreamWaitEvent(st, ctx->ev[1], 0));
        PP_CUDA(cudaEventRecord(ctx->ev[3], st));
        P.rank = (uint16_t *)s_rank;
        P.n_cells = n_cells;
        P.out_off = 0;
        launch_score(n_cells);
        PP_CHECK_LAUNCH();
        rc = finish_scoring();
        if (rc) return rc;
        PP_CUDA(cudaEventRecord(ctx->ev[4], st));
    } else {
        // Two ways for a chunk of cells to reach the device, through one double-buffered lane:
        //   gather: host threads copy the marker columns into pinned staging, H2D of [nc, U]
        //   direct: H2D of the full rows [nc, ld], columns picked by the rank kernel
        // Measured on the B200 box (16 host cores, cfg2: 100k x 20k f32, 1 479 marker genes): direct 195 ms, gather
        // 161 ms, alternating both at once 184 ms -- all three move ~40-50 GB/s through host DRAM, which is the bound,
        // so the way that touches the fewest host bytes is used alone.
        struct Events {  // destroyed on every exit path
            std::vector<cudaEvent_t> ev;
            cudaError_t make(cudaEvent_t *e, unsigned flags) {
                const cudaError_t rc = cudaEventCreateWithFlags(e, flags);
                if (rc == cudaSuccess) ev.push_back(*e);
                return rc;
            }
            ~Events() {
                for (cudaEvent_t e : ev) cudaEventDestroy(e);
            }
        } events;
        void *buf[2] = {nullptr, nullptr}, *zbuf[2] = {nullptr, nullptr};  // device chunk buffers of the two lanes (ctx slabs)
        cudaEvent_t copied[2], freed[2], zready[2], zfreed[2], ev_c0, ev_c1;
        std::vector<cudaEvent_t> ev_s;
        DevBuf d_ident;
        const int n_thr = host_threads();
        HostCrew crew(host_gather ? n_thr : 1);
        const bool dma_lane = host_gather || !zc_lane;  // staged / direct copies by the DMA engine
        const int64_t dev_ld = compact ? U : ld;        // row length of the device chunk buffers
        const int32_t *dev_cols = d_uni.as<int32_t>();
        cudaStream_t cs = ctx->copy_stream;
        if (host_gather) {
            const size_t need = (size_t)chunk_cells * U * esz;
            if (ctx->pin_bytes < need) {
                for (int b = 0; b < 2; ++b) {
                    if (ctx->pin[b]) cudaFreeHost(ctx->pin[b]);
                    ctx->pin[b] = nullptr;
                }
                ctx->pin_bytes = 0;
                for (int b = 0; b < 2; ++b) PP_CUDA(cudaHostAlloc(&ctx->pin[b], need, cudaHostAllocDefault));
                ctx->pin_bytes = need;
            }
        }
        if (compact) {
            std::vector<int32_t> ident(U);
            for (int j = 0; j < U; ++j) ident[j] = j;
            PP_CUDA(d_ident.alloc(sizeof(int32_t) * U, st));
            PP_CUDA(cudaMemcpyAsync(d_ident.p, ident.data(), sizeof(int32_t) * U, cudaMemcpyHostToDevice, st));
            PP_CUDA(cudaStreamSynchronize(st));  // ident is a local
            dev_cols = d_ident.as<int32_t>();
        }
        for (int b = 0; b < 2; ++b) {
            if (dma_lane) {
                if (!(buf[b] = pp_slab(ctx, b == 0 ? SLAB_CHUNK0 : SLAB_CHUNK1, (size_t)chunk_cells * dev_ld * esz)))
                    return PP_ENOMEM;
                PP_CUDA(events.make(&copied[b], cudaEventDisableTiming));
                PP_CUDA(events.make(&freed[b], cudaEventDisableTiming));
            }
            if (zc_lane) {
                if (!(zbuf[b] = pp_slab(ctx, b == 0 ? SLAB_ZC0 : SLAB_ZC1, (size_t)chunk_cells * U * esz))) return PP_ENOMEM;
                PP_CUDA(events.make(&zready[b], cudaEventDisableTiming));
                PP_CUDA(events.make(&zfreed[b], cudaEventDisableTiming));
            }
        }
        if (zc_lane && !ctx->zc_stream) PP_CUDA(cudaStreamCreateWithFlags(&ctx->zc_stream, cudaStreamNonBlocking));
        cudaStream_t zs = ctx->zc_stream;
        PP_CUDA(events.make(&ev_c0, cudaEventDefault));
        PP_CUDA(events.make(&ev_c1, cudaEventDefault));
        PP_CUDA(cudaEventRecord(ctx->ev[3], st));
        PP_CUDA(cudaStreamWaitEvent(cs, ctx->ev[3], 0));  // buffers exist before the first copy
        if (zc_lane) PP_CUDA(cudaStreamWaitEvent(zs, ctx->ev[3], 0));
        PP_CUDA(cudaEventRecord(ev_c0, cs));
        int i = 0, n_dma = 0, n_zc = 0;  // chunks so far: all, per lane
        double gather_host_ms = 0.0, wait_host_ms = 0.0;
        int64_t zc_bytes = 0;  // useful bytes the zero-copy lane fetched (the sectors on the bus are ~6x that)
        // rank + score of one chunk on the compute stream (which must already wait for the chunk's data)
        auto enqueue_compute = [&](const void *chunk_dev, int64_t c0, int64_t nc, cudaEvent_t done) -> int {
            int r = pp_rank_launch(ctx, chunk_dev, x_dtype, dev_ld, d_rows.as<int32_t>(), nc, dev_cols, U,
                                   (uint16_t *)s_rank + c0 * rank_ld, rank_ld, d_flags.as<int32_t>());
            if (r) return r;
            PP_CUDA(cudaEventRecord(done, st));  // the chunk buffer may be refilled
            cudaEvent_t e0, e1;
            PP_CUDA(events.make(&e0, cudaEventDefault));
            PP_CUDA(events.make(&e1, cudaEventDefault));
            ev_s.push_back(e0);
            ev_s.push_back(e1);
            P.rank = (uint16_t *)s_rank + c0 * rank_ld;
            P.n_cells = nc;
            P.out_off = c0;
            PP_CUDA(cudaEventRecord(e0, st));
            launch_score(nc);
            PP_CHECK_LAUNCH();
            PP_CUDA(cudaEventRecord(e1, st));
            return PP_OK;
        };
        // Zero-copy chunks whose gather kernel is in flight.  With both lanes active their rank + score are enqueued only
        // once the gather has finished (polled between the host gathers): the compute stream is in-order, and a chunk
        // enqueued early would hold back every DMA chunk that becomes ready before it.
        struct ZcPending { int64_t c0, nc; int b; };
        std::vector<ZcPending> zc_pending;
        auto drain_zc = [&](bool wait_all) -> int {
            while (!zc_pending.empty()) {
                const ZcPending z = zc_pending.front();
                if (wait_all)
                    PP_CUDA(cudaStreamWaitEvent(st, zready[z.b], 0));
                else if (cudaEventQuery(zready[z.b]) != cudaSuccess)
                    break;
                zc_pending.erase(zc_pending.begin());
                const int r = enqueue_compute(zbuf[z.b], z.c0, z.nc, zfreed[z.b]);
                if (r) return r;
            }
            return PP_OK;
        };
        tl_mark("pipeline_start");
        // the first chunks are an eighth, a quarter and a half wave: the device starts after ~1 ms of host work, not ~5
        for (int64_t c0 = 0, nc = 0; c0 < n_cells; c0 += nc, ++i) {
            if (i == 1) tl_mark("chunk0_enqueued");
            nc = i < 3 ? chunk_cells >> (3 - i) : chunk_cells;
            // ... and the last ones shrink again (half of what is left, a quarter wave at least): what remains to be done once
            // the host has gathered its last chunk is that chunk's copy + rank + score -- 12.7 ms after a full wave
            const int64_t left = n_cells - c0;
            if (i >= 3 && left <= chunk_cells + chunk_cells / 2)
                nc = left <= chunk_cells / 4 ? left : std::max(chunk_cells / 4, left / 2);
            nc = std::min({(std::max<int64_t>(nc, 1) + 127) / 128 * 128, chunk_cells, n_cells - c0});
            if (zc_lane && dma_lane) {
                rc = drain_zc(false);
                if (rc) return rc;
            }
            // lane of this chunk: the zero-copy lane whenever one of its two buffers is not being filled -- it costs the
            // host nothing --, else the DMA lane, whose gather keeps this thread busy for the length of a chunk
            const bool use_zc = zc_lane && (!dma_lane || (i > 0 && zc_pending.size() < 2));
            if (use_zc) {
                const int b = n_zc & 1;
                if (n_zc >= 2) PP_CUDA(cudaStreamWaitEvent(zs, zfreed[b], 0));  // the rank kernel that read zbuf[b] is done
                const unsigned grid = (unsigned)std::min<int64_t>((nc + ZC_WARPS - 1) / ZC_WARPS, ctx->n_sm);
                if (x_dtype == PP_F64)
                    zc_gather_kernel<double><<<grid, ZC_WARPS * 32, 0, zs>>>((const double *)xz, ld, c0, nc, d_uni.as<int32_t>(), U, (double *)zbuf[b]);
                else
                    zc_gather_kernel<float><<<grid, ZC_WARPS * 32, 0, zs>>>((const float *)xz, ld, c0, nc, d_uni.as<int32_t>(), U, (float *)zbuf[b]);
                PP_CHECK_LAUNCH();
                PP_CUDA(cudaEventRecord(zready[b], zs));
                zc_bytes += (int64_t)((size_t)nc * U * esz);
                ++n_zc;
                zc_pending.push_back(ZcPending{c0, nc, b});
                if (!dma_lane) {  // only lane: plain stream order
                    rc = drain_zc(true);
                    if (rc) return rc;
                }
            } else {
                const int b = n_dma & 1;
                const void *src = x0 + (size_t)c0 * ld * esz;
                if (host_gather) {
                    const double t_g0 = now_ms();
                    if (n_dma >= 2) PP_CUDA(cudaEventSynchronize(copied[b]));  // H2D two chunks ago has drained pin[b]
                    const double t_g1 = now_ms();
                    if (x_dtype == PP_F64)
                        gather_rows<double>((const double *)x0, ld, c0, nc, uni.data(), U, (double *)ctx->pin[b], crew);
                    else
                        gather_rows<float>((const float *)x0, ld, c0, nc, uni.data(), U, (float *)ctx->pin[b], crew);
                    src = ctx->pin[b];
                    wait_host_ms += t_g1 - t_g0;
                    gather_host_ms += now_ms() - t_g1;
                }
                if (n_dma >= 2) PP_CUDA(cudaStreamWaitEvent(cs, freed[b], 0));  // the rank kernel two chunks ago is done
                PP_CUDA(cudaMemcpyAsync(buf[b], src, (size_t)nc * dev_ld * esz, cudaMemcpyHostToDevice, cs));
                ctx->tm.h2d_bytes += (int64_t)((size_t)nc * dev_ld * esz);
                PP_CUDA(cudaEventRecord(copied[b], cs));
                PP_CUDA(cudaStreamWaitEvent(st, copied[b], 0));
                ++n_dma;
                rc = enqueue_compute(buf[b], c0, nc, freed[b]);
                if (rc) return rc;
            }
        }
        rc = drain_zc(true);
        if (rc) return rc;
        ctx->tm.h2d_bytes += zc_bytes;
        tl_mark("chunks_enqueued");
        PP_CUDA(cudaEventRecord(ev_c1, cs));
        if (trace)
            fprintf(stderr, "[pp_cyclone] %d chunks of %lld cells (%d by DMA%s, %d by the zero-copy gather), %d host threads: gather %.1f ms, "
                    "waiting for staging buffers %.1f ms\n", i, (long long)chunk_cells, n_dma, host_gather ? " after the host gather" : "",
                    n_zc, n_thr, gather_host_ms, wait_host_ms);
        {
            cudaEvent_t e0, e1;
            PP_CUDA(events.make(&e0, cudaEventDefault));
            PP_CUDA(events.make(&e1, cudaEventDefault));
            ev_s.push_back(e0);
            ev_s.push_back(e1);
            PP_CUDA(cudaEventRecord(e0, st));
            rc = finish_scoring();
            if (rc) return rc;
            PP_CUDA(cudaEventRecord(e1, st));
        }
        PP_CUDA(cudaEventRecord(ctx->ev[4], st));
        PP_CUDA(cudaStreamSynchronize(cs));
        if (zc_lane) PP_CUDA(cudaStreamSynchronize(zs));
        PP_CUDA(cudaStreamSynchronize(st));
        tl_mark("scored");
        cudaEventElapsedTime(&copy_ms, ev_c0, ev_c1);
        for (size_t j = 0; j + 1 < ev_s.size(); j += 2) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev_s[j], ev_s[j + 1]);
            score_ms += ms;
        }
    }
    if (dist) {
        rc = cyclone_gather(ctx, d_scores.as<double>(), d_oh.as<int32_t>(), d_ot.as<int32_t>(), d_flags.as<int32_t>(), n_cells,
                            n_classes, shard_cells, gather_to, out_scores, out_obs_hits, out_obs_total, want_oh, want_ot);
        if (rc) return rc;
    } else {
    // results: straight into the caller's arrays when they are pinned (pp_alloc_result), else device -> pinned staging ->
    // the caller's pageable arrays (a direct copy would be staged by the driver at a fraction of the rate)
    const size_t sc_bytes = sizeof(double) * out_cnt, ob_bytes = sizeof(int32_t) * out_cnt;
    const bool direct = pp_is_pinned_host(out_scores) && (!want_oh || pp_is_pinned_host(out_obs_hits)) &&
                        (!want_ot || pp_is_pinned_host(out_obs_total));
    const size_t stage_bytes = (direct ? 0 : sc_bytes + (want_oh ? ob_bytes : 0) + (want_ot ? ob_bytes : 0)) + 16;
    if (ctx->pin_out_bytes < stage_bytes) {
        if (ctx->pin_out) cudaFreeHost(ctx->pin_out);
        ctx->pin_out = nullptr;
        ctx->pin_out_bytes = 0;
        const size_t want = std::max<size_t>(stage_bytes * 2, (size_t)8 << 20);
        PP_CUDA(cudaHostAlloc(&ctx->pin_out, want, cudaHostAllocDefault));
        ctx->pin_out_bytes = want;
    }
    uint8_t *stage_p = (uint8_t *)ctx->pin_out;
    int32_t *flags = (int32_t *)stage_p;
    uint8_t *st_sc = direct ? (uint8_t *)out_scores : stage_p + 16;
    uint8_t *st_oh = direct ? (uint8_t *)out_obs_hits : st_sc + sc_bytes;
    uint8_t *st_ot = direct ? (uint8_t *)out_obs_total : st_oh + (want_oh ? ob_bytes : 0);
    PP_CUDA(cudaMemcpyAsync(st_sc, d_scores.p, sc_bytes, cudaMemcpyDeviceToHost, st));
    if (want_oh) PP_CUDA(cudaMemcpyAsync(st_oh, d_oh.p, ob_bytes, cudaMemcpyDeviceToHost, st));
    if (want_ot) PP_CUDA(cudaMemcpyAsync(st_ot, d_ot.p, ob_bytes, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(int32_t) * 4, cudaMemcpyDeviceToHost, st));
    PP_CUDA(cudaEventRecord(ctx->ev[5], st));
    PP_CUDA(cudaStreamSynchronize(st));
    if (flags[1]) return pp_fail(ctx, PP_ENAN, "pp_cyclone: NaN in the expression matrix");
    if (!direct) {
        memcpy(out_scores, st_sc, sc_bytes);
        if (want_oh) memcpy(out_obs_hits, st_oh, ob_bytes);
        if (want_ot) memcpy(out_obs_total, st_ot, ob_bytes);
    }
    }
    tl_mark("results_on_host");
    if (tl_on) fprintf(stderr, "pp_cyclone:%s\n", tl_line.c_str());
    if (pipelined) {  // copies overlap compute: h2d = copy-stream span, main = sum of the score launches
        ctx->tm.h2d_ms = copy_ms;
        ctx->tm.main_ms = score_ms;
        cudaEventElapsedTime(&ctx->tm.prep_ms, ctx->ev[0], ctx->ev[1]);
    } else {
        float t_prep = 0.f;  // everything before the first score launch: copy, rank, tables
        if (!x_is_device || csr) cudaEventElapsedTime(&ctx->tm.h2d_ms, ctx->ev[6], ctx->ev[7]);
        cudaEventElapsedTime(&t_prep, ctx->ev[0], ctx->ev[3]);
        ctx->tm.prep_ms = t_prep - ctx->tm.h2d_ms;
        cudaEventElapsedTime(&ctx->tm.main_ms, ctx->ev[3], ctx->ev[4]);
    }
    cudaEventElapsedTime(&ctx->tm.post_ms, ctx->ev[4], ctx->ev[5]);
    cudaEventElapsedTime(&ctx->tm.total_ms, ctx->ev[0], ctx->ev[5]);
    ctx->tm.n_launches = ctx->launches;
    return PP_OK;
}

PP_API int pp_cyclone(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                      int64_t ld, int64_t row_begin, int64_t n_cells, int32_t n_classes, const int32_t *used_idx,
                      const int64_t *used_off, const int32_t *pairs, const int64_t *pair_off, int32_t iterations,
                      int32_t min_iter, int32_t min_pairs, int32_t rng_mode, uint64_t seed,
                      const int32_t *perm_tables, double *out_scores, int32_t *out_obs_hits,
                      int32_t *out_obs_total) {
    if (ctx && !x) return pp_fail(ctx, PP_EINVAL, "pp_cyclone: NULL argument");
    PP_GUARD(ctx, cyclone_impl(ctx, x, x_dtype, x_is_device, nullptr, n_rows, n_cols, ld, row_begin, n_cells, n_classes,
                               used_idx, used_off, pairs, pair_off, iterations, min_iter, min_pairs, rng_mode, seed,
                               perm_tables, nullptr, -1, out_obs_hits != nullptr, out_obs_total != nullptr, out_scores,
                               out_obs_hits, out_obs_total));
}

PP_API int pp_cyclone_csr(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, int64_t row_begin,
                          int64_t n_cells, int32_t n_classes, const int32_t *used_idx, const int64_t *used_off,
                          const int32_t *pairs, const int64_t *pair_off, int32_t iterations, int32_t min_iter,
                          int32_t min_pairs, int32_t rng_mode, uint64_t seed, const int32_t *perm_tables,
                          double *out_scores, int32_t *out_obs_hits, int32_t *out_obs_total) {
    if (ctx && !x) return pp_fail(ctx, PP_EINVAL, "pp_cyclone_csr: NULL argument");
    PP_GUARD(ctx, cyclone_impl(ctx, nullptr, x ? x->data_dtype : PP_F32, x ? x->is_device : 0, x, n_rows, n_cols, n_cols,
                               row_begin, n_cells, n_classes, used_idx, used_off, pairs, pair_off, iterations, min_iter,
                               min_pairs, rng_mode, seed, perm_tables, nullptr, -1, out_obs_hits != nullptr,
                               out_obs_total != nullptr, out_scores, out_obs_hits, out_obs_total));
}

// want_observed: the observed-counter arrays must be requested by ALL ranks or by none (they change the size of the
// all-gathered block); a non-receiving rank passes NULL outputs, so the wish travels separately.
PP_API int pp_cyclone_dist(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows, int64_t n_cols,
                           int64_t ld, int64_t row_begin, int64_t n_cells, int32_t n_classes, const int32_t *used_idx,
                           const int64_t *used_off, const int32_t *pairs, const int64_t *pair_off, int32_t iterations,
                           int32_t min_iter, int32_t min_pairs, int32_t rng_mode, uint64_t seed,
                           const int32_t *perm_tables, const int64_t *shard_cells, int32_t gather_to,
                           int32_t want_observed, double *out_scores, int32_t *out_obs_hits, int32_t *out_obs_total) {
    if (ctx && ((!x && n_rows > 0) || !shard_cells)) return pp_fail(ctx, PP_EINVAL, "pp_cyclone_dist: NULL argument");
    PP_GUARD(ctx, cyclone_impl(ctx, x, x_dtype, x_is_device, nullptr, n_rows, n_cols, ld, row_begin, n_cells, n_classes,
                               used_idx, used_off, pairs, pair_off, iterations, min_iter, min_pairs, rng_mode, seed,
                               perm_tables, shard_cells, gather_to, want_observed != 0, want_observed != 0, out_scores,
                               out_obs_hits, out_obs_total));
}

PP_API int pp_cyclone_csr_dist(pp_ctx *ctx, const pp_csr *x, int64_t n_rows, int64_t n_cols, int64_t row_begin,
                               int64_t n_cells, int32_t n_classes, const int32_t *used_idx, const int64_t *used_off,
                               const int32_t *pairs, const int64_t *pair_off, int32_t iterations, int32_t min_iter,
                               int32_t min_pairs, int32_t rng_mode, uint64_t seed, const int32_t *perm_tables,
                               const int64_t *shard_cells, int32_t gather_to, int32_t want_observed, double *out_scores,
                               int32_t *out_obs_hits, int32_t *out_obs_total) {
    if (ctx && (!x || !shard_cells)) return pp_fail(ctx, PP_EINVAL, "pp_cyclone_csr_dist: NULL argument");
    PP_GUARD(ctx, cyclone_impl(ctx, nullptr, x ? x->data_dtype : PP_F32, x ? x->is_device : 0, x, n_rows, n_cols, n_cols,
                               row_begin, n_cells, n_classes, used_idx, used_off, pairs, pair_off, iterations, min_iter,
                               min_pairs, rng_mode, seed, perm_tables, shard_cells, gather_to, want_observed != 0,
                               want_observed != 0, out_scores, out_obs_hits, out_obs_total));
}

PP_API int pp_phase_scores(pp_ctx *ctx, const void *x, int x_dtype, int x_is_device, int64_t n_rows,
                           int64_t n_cols, int64_t ld, int32_t iterations, int32_t min_iter, int32_t min_pairs,
                           const int64_t *pairs, int64_t n_pairs, const uint8_t *used_mask, double *out_scores) {
    if (!ctx) return pp_fail(nullptr, PP_EINVAL, "ctx is NULL");
    if (!pairs || !used_mask) return pp_fail(ctx, PP_EINVAL, "pp_phase_scores: NULL argument");
    auto body = [&]() -> int {
    std::vector<int32_t> used;
    for (int64_t g = 0; g < n_cols; ++g)
        if (used_mask[g]) used.push_back((int32_t)g);
    std::vector<int32_t> p32((size_t)2 * n_pairs);
    for (int64_t i = 0; i < 2 * n_pairs; ++i) {
        if (pairs[i] < 0 || pairs[i] >= (int64_t)used.size()) return pp_fail(ctx, PP_EINVAL, "pp_phase_scores: pair index out of range");
        p32[i] = (int32_t)pairs[i];
    }
    const int64_t uoff[2] = {0, (int64_t)used.size()}, poff[2] = {0, n_pairs};
    return pp_cyclone(ctx, x, x_dtype, x_is_device, n_rows, n_cols, ld, 0, n_rows, 1, used.data(), uoff, p32.data(), poff,
                      iterations, min_iter, min_pairs, PP_RNG_MT19937, 2803, nullptr, out_scores, nullptr, nullptr);
    };
    PP_GUARD(ctx, body());
}
