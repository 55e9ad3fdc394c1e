// knn_batched.cu -- batched-query path on the 5th-gen tensor cores (sm_100a: tcgen05 + TMEM + bulk copies / TMA).
//
// For a batch of queries the scan IS a dense contraction: S = X . Q^T (rows x queries).  One CTA per SM
// owns a fixed group of BN = 256 queries (resident in shared memory, loaded once) and streams its share of
// the collection's row tiles (BM = 128 rows) through a multi-stage copy -> mbarrier -> tcgen05.mma pipeline
// into a double-buffered TMEM accumulator (2 x 256 columns).  Sixteen epilogue warps read the accumulator
// with tcgen05.ld, hand the stage back at once, and NEVER materialise the rows x queries score matrix: a
// branch-free maximum over each thread's 32 scores tells whether its row beats any query's threshold, the
// rare blocks with a candidate are parked in a small per-warp stash and appended after the tile to a
// per-(CTA, query) candidate list in global memory.
//
// Two operand modes.  Default: TILED bf16 mirrors of the rows and of the query batch (kind::f16; every
// (tile, K step) block is 4 KB / 8 KB of contiguous memory in the no-swizzle K-major UMMA layout, moved by
// plain bulk copies) whose 16 aux columns fold the threshold and the row-norm term into the contraction,
// so the accumulator holds D = dot + thr_q - ||x||^2/2 and "candidate" is D > 0.  Fallback / set_path(4):
// the stored f32 rows as tf32 operands (kind::tf32 through TMA tensor maps, 128-byte swizzle, no second
// copy), thresholds applied in the epilogue.
//
// The row tiles are processed in phases; between phases batched_finish_kernel merges the lists of all
// CTAs of a query (radix select) into its best k' keys and publishes the next phase's threshold: the
// k'-th key, or a GUESS (guess_rank.hpp) that the following finish verifies.  In the last phase it
// re-evaluates the best k' candidates in canonical f64 (the same arithmetic as the oracle), sorts them by
// (dist, id) and PROVES -- with the rounding error of the operands as measured while the mirrors were
// built -- that no dropped row can belong to the top k; a query whose proof fails is flagged and answered
// by the single-query scan (knn_scan.cu).  Collections whose proofs fail escalate: wider margin k', then
// BAND mode (every candidate within the error band above the k-th best is kept and reranked).
//
// Stands for the reference's SearchCommand::execute (src/command/types.rs:114-119, empty) when the
// argument carries many queries; nothing of it exists upstream.
#include "knn_batched.cuh"
#include "guess_rank.hpp"

#include <cuda.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "knn_device.cuh"

namespace vrod {

namespace {

constexpr int BM = 128;          // collection rows per tile  (UMMA M, TMEM lanes)
constexpr int BN = 256;          // queries per CTA           (UMMA N, TMEM columns per accumulator stage)
constexpr int KS = 32;           // f32 elements per K slab = one 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32
constexpr int CAP = 1024;        // candidate keys per (CTA, query)
constexpr int kCheckEvery = 4;     // tiles between two list-maintenance points of the epilogue warps
constexpr int PRUNE_AT = CAP - (kCheckEvery + 2) * BM;   // lists longer than this ask for a prune at the next point
constexpr int kEpiWarps = 16;    // four per TMEM lane quadrant, each takes 64 of the 256 columns (two 32-column blocks per tile)
// warp 0 TMA / bulk copies, warp 1 MMA + TMEM alloc, warps 2..17 epilogue, warp 18 second MMA issuer (bf16 mode)
constexpr int kThreads = 64 + 32 * kEpiWarps + 32;
constexpr int kIssuer2 = kThreads / 32 - 1;
constexpr int SLAB_A_BYTES = BM * KS * 4;   // 16 KB: 128 rows x 128 B
// query slab of one CTA: all BN queries (32 KB) alone, its half (16 KB) in a CTA pair
constexpr int SLAB_B_BYTES = BN * KS * 4;   // 32 KB: 256 queries x 128 B
constexpr int MAX_SLABS = 4;                // tf32, dim <= 128: the query group stays resident (128 KB)
// bf16 operand mode ("H"): a K slab is 64 bf16 = the same 128-byte swizzle row, one MMA covers K = 16.  The mirror of
// a row / a query is [kd data columns | 16 aux columns] (kd = dim rounded up to 16): the aux columns fold the query's
// threshold and the row's norm term into the contraction (see build_mirror_kernel), so the accumulator already
// holds the candidate test value.
constexpr int UMMA_K_H = 16;
constexpr int AUX_H = 16;
// The bf16 mirrors are stored TILED, one contiguous block per (tile, K step): a K step is 16 columns = 32 bytes per
// row, and its block holds [rows/8 groups][2 chunks of 8 columns][8 rows][16 bytes] -- the tensor core's plain
// (no-swizzle) K-major form, core matrices of 8 rows x 16 bytes stored contiguously.  A 128-row tile is T blocks of
// 4 KB back to back in global memory, a 256-query group T blocks of 8 KB, so a pipeline stage is ONE contiguous
// cp.async.bulk of up to 36 KB -- no tensor map, no 128-byte box rows cut out of 288-byte matrix rows (round 1: 48 KB
// of shared-memory writes per 128 x 144 tile, 1286 cycles per tile for the TMA ring alone), no padding of the
// 16-column tail slab.
constexpr int KB_A = BM * 32;               // bytes of one K-step block of a row tile
constexpr int KB_B = BN * 32;               // bytes of one K-step block of a query group
constexpr uint32_t kPlainLBO = 128, kPlainSBO = 256;   // chunk-to-chunk and group-to-group strides inside a block

struct BatchedParams {
    const float *sq_norm, *inv_norm;
    uint32_t n, b, nslab, stages;
    uint32_t qgroups, units;         // units = CTAs; unit u serves query group u % qgroups
    uint32_t tile_begin, tile_end;   // this launch (phase) covers row tiles [tile_begin, tile_end)
    const float *thr_init;      // [b] thresholds carried over from the previous phase (nullptr: +inf)
    unsigned long long *cand;   // [grid][BN][CAP]
    int *cnt_out;               // [grid][BN]
    int *qflags;                // [b] |= 1 when a list could not be pruned (query is rescanned)
    int kprime;
    int stream_q;               // 1: the query slabs are streamed with the row slabs (dim > 128: the group does not fit)
    uint32_t ksteps_last;       // MMA K steps in the last slab (the others hold 4)
    // bf16 mode: tiled mirrors (see KB_A / KB_B)
    const unsigned char *rows_t;   // [tiles][T][KB_A]
    const unsigned char *q_t;      // [query groups of the wave][T][KB_B]
    uint32_t T, S;                 // K steps per tile; K steps per pipeline stage
    int debug_nocand;           // VROD_BATCHED_DEBUG=nocand: thresholds start at -inf (timing experiments only)
    int debug_skip;             // timing experiments only: bit 0 = epilogue skips the TMEM reads (noepi), bit 1 = no MMAs issued (nomma)
    long long *dbg;             // VROD_BATCHED_DEBUG set: per-CTA cycle counters [grid][8]
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
// 1-D bulk copy global -> shared (contiguous, multiple of 16 bytes), completion on an mbarrier
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// One lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// Plain (no-swizzle) K-major operand block: low word = start address | LBO, high word = SBO | descriptor version.
// Advancing to another block only ever adds to the start-address field of the LOW word (shared memory is < 256 KB,
// so the 14-bit field cannot carry into the LBO field): the issue loop needs ONE integer add per operand per MMA.
__device__ __forceinline__ uint32_t plain_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((kPlainLBO >> 4) << 16); }
constexpr uint32_t kPlainDescHi = (kPlainSBO >> 4) | (1u << 14);
__device__ __forceinline__ void tc_mma_bf16_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accum), "r"(kPlainDescHi)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;             // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;   // stride byte offset
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
// kind::tf32, f32 accumulate, A and B K-major, N = 256, M = 128 (one CTA) or 256 (CTA pair: 128 rows per CTA)
__host__ __device__ constexpr uint32_t idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// kind::f16 with bf16 A and B, f32 accumulate, both K-major, N = 256, M = 128
__host__ __device__ constexpr uint32_t idesc_bf16() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Deferred appends.  A candidate costs the warp that finds it 1000-2000 cycles (shared-memory atomics, a global store,
// above all COLD code: instruction fetch), and the accumulator stage it reads goes back to the MMA warp only when all
// epilogue warps are done with it: with 8 warps of 128 columns each (~850 busy cycles per tile) every candidate stalled
// the tensor core ~90 cycles -- 20-30 % of a configs[2] batch.  Two changes keep the rare path off the MMA's critical
// path: (1) sixteen epilogue warps of 64 columns each (~400 busy cycles per ~1400-cycle tile), so a late warp has slack;
// (2) the hitting lane only PARKS its 32 scores in a small per-warp stash (32 predicated STS with static register
// indices) and the warp empties the stash after it has handed the stage back.  A warp may then be late by two tile
// times minus twice its own work before the MMA warp notices.
constexpr int kStash = 4;                 // parked blocks per warp; fuller blocks (the dense early phases) go in rounds
struct StashSlot {      // 140 bytes: 16 warps x 4 slots make the control block 11200 bytes, which just leaves four 36 KB
    float score[32];    // stages next to the 72 KB query operand at dim = 128 (227 KB of shared memory; asserted below)
    uint32_t row, colbase;
    float hx;
};
static_assert(sizeof(StashSlot) == 140, "stash slot size");
struct BatchCtl {
    uint64_t full[8], empty[8], tfull[2], tempty[2], qfull;
    int pflag[4];                // per 64-column part: one of its lists is long, prune at the next maintenance point
    uint32_t tmem_base;
    int flag;
    alignas(16) float thr[BN];
    int cnt[BN];
    StashSlot stash[kEpiWarps][kStash];
};
static_assert(9 * KB_B + 4 * 9 * KB_A + sizeof(BatchCtl) <= 227 * 1024,
              "dim 128 (9 K steps): the resident query group, four whole-tile stages and the control block must fit 227 KB");

// ---- epilogue: prune one (CTA, query) list with one warp ------------------------------------------
// Keep the entries <= t where t is the smallest sampled key with at least kprime entries at or below
// it (verified by an exact count), compact them to the front, tighten the threshold.  A safety net for adversarial
// row orders -- it does not run at all on the benchmark workloads -- so it is written for SIZE (rolled loops over
// the list in L2, one copy per kernel): the register-resident version it replaces was 14 KB of SASS, inlined twice,
// in a kernel whose hot loops want the 32 KB instruction cache for themselves.
// H (bf16 mode): ctl->thr[q] is the threshold FOLDED into the contraction (the accumulator holds D = dot + thr - hx and
// the append recovers the surrogate as thr - D), so it must stay what the query mirror carries: the prune then only
// compacts the list and candidates keep arriving at the phase threshold until the next phase tightens it.
template <bool H>
__device__ __noinline__ void prune_list(unsigned long long *cq, int q, BatchCtl *ctl, int kprime, int lane) {
    int c = ctl->cnt[q];
    if (c > CAP) c = CAP;
    if (c <= kprime) return;
    // 32 samples (arrival order is unrelated to value), sorted across the warp
    unsigned long long s = __ldcg(cq + lane);   // c > kprime >= 64 > lane
#pragma unroll 1
    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll 1
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, s, j);
            const bool up = (lane & k2) == 0, low = (lane & j) == 0;
            s = (low == up) ? (s < o ? s : o) : (s < o ? o : s);
        }
    int j = (kprime * 32 + c - 1) / c;
    if (j > 31) j = 31;
    unsigned long long t;
    int count;
#pragma unroll 1
    while (true) {
        t = __shfl_sync(kFull, s, j);
        count = 0;
#pragma unroll 1
        for (int i = lane; i < c; i += 32) count += __ldcg(cq + i) <= t ? 1 : 0;
        count = __reduce_add_sync(kFull, count);
        if (count >= kprime || j == 31) break;
        ++j;
    }
    // No sampled key has kprime entries at or below it: the list is barely longer than kprime (with c = kprime + 1 the
    // largest of 32 samples is the wanted key only ~1 time in 5) or the samples were unlucky.  Nothing can be dropped
    // safely, so the list stays as it is.  (Round 1 gave the query up here -- with k' = 256 and lists that hover just
    // above 256 entries that flagged three quarters of an 8192-query batch for a rescan.)  A list that really
    // overruns CAP later is caught at the append (qflags), so leaving it alone is safe.
    if (count < kprime) return;
    // in-place compaction, 32 entries at a time (a chunk is read by the whole warp before any lane writes, and it is
    // written at or below where it was read)
    int base = 0;
#pragma unroll 1
    for (int i0 = 0; i0 < c; i0 += 32) {
        const unsigned long long key = i0 + lane < c ? __ldcg(cq + i0 + lane) : kKeyMax;
        const bool keep = key <= t;
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) __stcg(cq + base + __popc(m & ((1u << lane) - 1)), key);
        base += __popc(m);
        __syncwarp();
    }
    if (lane == 0) {
        ctl->cnt[q] = base;
        if (!H) ctl->thr[q] = ord2f((uint32_t)(t >> 32));
    }
}

// The candidate test, in the one form both passes use (so they agree bit for bit):
//   L2      dot + thr > hx           (hx = ||x||^2 / 2;  i.e. hx - dot < thr up to one rounding)
//   cosine  fma(dot, hx, thr) > 0    (hx = 1 / ||x||;    i.e. -dot/||x|| < thr up to one rounding)
// The one rounding is part of the guard's error budget (batched_finish_kernel).
template <bool COS>
__device__ __forceinline__ float cand_w(float dot, float thr, float hx) {
    return COS ? fmaf(dot, hx, thr) : (dot + thr);
}
template <bool COS>
__device__ __forceinline__ bool cand_hit(float w, float hx) {
    return COS ? (w > 0.f) : (w > hx);
}

// Hot filter: 32 scores of one thread (its row x 32 query columns) -> max of the test values.  One add (or
// fma) and one max per score, four independent chains; no per-score compare, select or branch.
template <bool COS>
__device__ __forceinline__ float block_max(const uint32_t (&r)[32], const float *thr, float hx, float m) {
    float m0 = m, m1 = m, m2 = m, m3 = m;
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 th = *reinterpret_cast<const float4 *>(thr + j4 * 4);
        m0 = fmaxf(m0, cand_w<COS>(__uint_as_float(r[j4 * 4 + 0]), th.x, hx));
        m1 = fmaxf(m1, cand_w<COS>(__uint_as_float(r[j4 * 4 + 1]), th.y, hx));
        m2 = fmaxf(m2, cand_w<COS>(__uint_as_float(r[j4 * 4 + 2]), th.z, hx));
        m3 = fmaxf(m3, cand_w<COS>(__uint_as_float(r[j4 * 4 + 3]), th.w, hx));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// bf16 mode: the accumulator already holds D = dot + thr_q - ||x||^2/2 (cosine: dot^ + thr_q), so the hot filter is a
// bare maximum over the 32 columns (half a FMNMX3 per score) and "candidate" means D > 0.
__device__ __forceinline__ float block_max_h(const uint32_t (&r)[32]) {
    float m0 = __uint_as_float(r[0]), m1 = __uint_as_float(r[1]), m2 = __uint_as_float(r[2]), m3 = __uint_as_float(r[3]);
#pragma unroll
    for (int j4 = 1; j4 < 8; ++j4) {
        m0 = fmaxf(m0, __uint_as_float(r[j4 * 4 + 0]));
        m1 = fmaxf(m1, __uint_as_float(r[j4 * 4 + 1]));
        m2 = fmaxf(m2, __uint_as_float(r[j4 * 4 + 2]));
        m3 = fmaxf(m3, __uint_as_float(r[j4 * 4 + 3]));
    }
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// ---- rare paths of the epilogue ---------------------------------------------------------------------
// Code size is a first-order cost here: the tile kernel is far larger than the SM's 32 KB instruction cache, so any
// path that a warp enters only now and then is fetched cold (~2000 cycles for a few dozen straight-line
// instructions; round 1 measured 2000-2500 cycles per entry for its first forms).  The rare paths are therefore
// (a) as few instructions as possible, (b) NOT inlined -- one copy per kernel, shared by both halves of the
// ping-pong loop -- and (c) kept off the accumulator's critical path (stash, above).  (Round 1's register-resident
// append -- a 31-select tree per candidate, inlined twice, plus a separate dense-block form -- was 35 KB of the kernel.)

// Empty a warp's stash, one parked block (= one row x 32 query columns) at a time with the lanes across the COLUMNS: lane j
// tests column j and appends it if it passed.  No inner loops, no per-lane scans: the whole function is ~40
// instructions, the same for one parked block or for a dense one.
template <bool COS, bool H>
__device__ __noinline__ void stash_flush(int first, int nst, const StashSlot *slots, BatchCtl *ctl, unsigned long long *cand, int *qflags,
                                         uint32_t qbase, uint32_t b, int lane) {
#pragma unroll 1
    for (int e = first; e < first + nst; ++e) {
        const StashSlot &sl = slots[e & (kStash - 1)];
        const float dot = sl.score[lane];
        const int q = (int)sl.colbase + lane;
        const float thr = ctl->thr[q], hx = sl.hx;
        bool hit;
        if constexpr (H) hit = dot > 0.f;
        else hit = cand_hit<COS>(cand_w<COS>(dot, thr, hx), hx);
        if (hit) {
            const int pos = atomicAdd(&ctl->cnt[q], 1);
            if (pos >= PRUNE_AT) *(volatile int *)&ctl->pflag[q >> 6] = 1;
            const float v = H ? (thr - dot) : (COS ? -(dot * hx) : (hx - dot));
            if (pos < CAP) __stcg(cand + (size_t)q * CAP + pos, make_key(v, sl.row));
            else if (qbase + q < b) atomicOr(qflags + qbase + q, 1);   // cannot happen; checked
        }
    }
    __syncwarp();
}

// One CTA per unit (tcgen05 cta_group::1, M = 128).  (A CTA-pair variant of the tf32 kernel -- cluster of 2, cta_group::2,
// M = 256, half of the query group per CTA -- passed parity in round 1, measured slower and was removed in round 2:
// tools/mma_probe.cu shows one CTA already issues at the tensor core's full rate.)
// H = bf16 operand mode: operands are the bf16 mirrors with the folded aux columns, MMA kind::f16.
// DENSE = the start phase (one tile per CTA, every threshold still above every surrogate): the epilogue writes all 128
// rows of the tile as candidates of all 256 queries, no test, position = row inside the tile, 256-byte coalesced stores.
template <bool COS, bool H, bool DENSE>
__global__ void __launch_bounds__(kThreads, 1) batched_tile_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                   const __grid_constant__ CUtensorMap tmQ,
                                                                   const BatchedParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int SLAB_B = SLAB_B_BYTES;
    constexpr int KSE = KS;   // f32 elements per K slab (tf32 mode)
    // resident mode: [query operand of the group][stages x row operand]; streamed mode: [stages x (row + query operand)]
    // (tf32: 128-byte-swizzled slabs written by TMA; bf16: S K-step blocks of the tiled mirrors per stage)
    const uint32_t stage_bytes = H ? (p.stream_q ? p.S * (KB_A + KB_B) : p.S * KB_A) : (p.stream_q ? (SLAB_A_BYTES + SLAB_B) : SLAB_A_BYTES);
    unsigned char *q_s = smem;
    unsigned char *a_s = p.stream_q ? smem : smem + (H ? (size_t)p.T * KB_B : (size_t)p.nslab * SLAB_B);
    BatchCtl *ctl = reinterpret_cast<BatchCtl *>(a_s + (size_t)p.stages * stage_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (smem_u32(smem) & 1023u) __trap();          // the 128-byte swizzle needs a 1024-byte aligned base
    const uint32_t unit = blockIdx.x;
    const uint32_t g = unit % p.qgroups;           // query group of this unit
    const uint32_t member = unit / p.qgroups;
    const uint32_t cpg = (p.units - g + p.qgroups - 1) / p.qgroups;        // units serving this group
    const uint32_t span = p.tile_end - p.tile_begin;                        // row tiles of the phase
    const uint32_t my_tiles = member < span ? (span - member + cpg - 1) / cpg : 0;
    auto tile_of = [&](uint32_t i) { return p.tile_begin + member + i * cpg; };

    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(&ctl->full[s], 1);             // this CTA's own TMA (local transaction bytes)
            mbar_init(&ctl->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&ctl->tfull[a], 1);
            mbar_init(&ctl->tempty[a], kEpiWarps);
        }
        mbar_init(&ctl->qfull, 1);
        ctl->flag = 0;
        for (int w = 0; w < 4; ++w) ctl->pflag[w] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < BN; i += kThreads) {
        const bool live = g * BN + i < p.b && !(kDbg && p.debug_nocand);
        const float t0 = (live && p.thr_init) ? p.thr_init[g * BN + i] : __int_as_float(0x7f800000);
        ctl->thr[i] = live ? t0 : -__int_as_float(0x7f800000);
        ctl->cnt[i] = 0;
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ctl->tmem_base;
    const int q_row0 = (int)(g * BN);   // first query row of this CTA's group

    if (warp == 0) {
        // ===== producer =====
        if (H) {
            // ===== bf16 mode: contiguous bulk copies of the tiled mirrors (one per stage and operand) =====
            if (lane == 0) {
                const unsigned char *qg = p.q_t + (size_t)g * p.T * KB_B;
                auto copy = [&](unsigned char *dst, const unsigned char *src, uint32_t bytes, uint64_t *bar) {
#pragma unroll 1
                    for (uint32_t o = 0; o < bytes; o += 32768) bulk_load(dst + o, src + o, bytes - o < 32768 ? bytes - o : 32768, bar);
                };
                if (!p.stream_q) {
                    mbar_expect_tx(&ctl->qfull, p.T * KB_B);
                    copy(q_s, qg, p.T * KB_B, &ctl->qfull);
                }
                const uint32_t ns = (p.T + p.S - 1) / p.S;
                uint32_t stage = 0, phase = 0;
                for (uint32_t i = 0; i < my_tiles; ++i) {
                    const unsigned char *ta = p.rows_t + (size_t)tile_of(i) * p.T * KB_A;
                    for (uint32_t st = 0; st < ns; ++st) {
                        const uint32_t nk = p.T - st * p.S < p.S ? p.T - st * p.S : p.S;
                        mbar_wait(&ctl->empty[stage], phase ^ 1);
                        unsigned char *dst = a_s + (size_t)stage * stage_bytes;
                        mbar_expect_tx(&ctl->full[stage], nk * KB_A + (p.stream_q ? nk * KB_B : 0));
                        copy(dst, ta + (size_t)st * p.S * KB_A, nk * KB_A, &ctl->full[stage]);
                        if (p.stream_q) copy(dst + p.S * KB_A, qg + (size_t)st * p.S * KB_B, nk * KB_B, &ctl->full[stage]);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else if (lane == 0) {
            // ===== tf32 mode: 128-byte-wide TMA boxes out of the stored f32 rows =====
            auto load = [&](void *dst, const CUtensorMap *m, int c0, int c1, uint64_t *bar) { tma_load_2d(dst, m, c0, c1, bar); };
            auto arm = [&](uint64_t *bar, uint32_t bytes) { mbar_expect_tx(bar, bytes); };
            if (!p.stream_q) {
                arm(&ctl->qfull, p.nslab * SLAB_B);
                for (uint32_t s = 0; s < p.nslab; ++s) load(q_s + (size_t)s * SLAB_B, &tmQ, (int)(s * KSE), q_row0, &ctl->qfull);
            }
            uint32_t stage = 0, phase = 0;
            long long w_empty = 0;
            for (uint32_t i = 0; i < my_tiles; ++i) {
                const uint32_t tile = tile_of(i);
                for (uint32_t s = 0; s < p.nslab; ++s) {
                    const long long te = (kDbg && p.dbg) ? clock64() : 0;
                    mbar_wait(&ctl->empty[stage], phase ^ 1);
                    if (kDbg && p.dbg) w_empty += clock64() - te;
                    arm(&ctl->full[stage], stage_bytes);
                    load(a_s + (size_t)stage * stage_bytes, &tmX, (int)(s * KSE), (int)(tile * BM), &ctl->full[stage]);
                    if (p.stream_q) load(a_s + (size_t)stage * stage_bytes + SLAB_A_BYTES, &tmQ, (int)(s * KSE), q_row0, &ctl->full[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
            if (kDbg && p.dbg) p.dbg[blockIdx.x * 16 + 9] = w_empty;
        }
    } else if ((warp == 1 || warp == kIssuer2) && H) {
        // ===== bf16 mode MMA issuer.  The whole warp runs the loop converged (waits included) and ONE elected lane
        // issues; the operand descriptors advance by a single add on their low word.  Round 1 issued from inside an
        // `if (lane == 0)` region and rebuilt both 64-bit descriptors from byte addresses for every MMA: ~18 dependent
        // uniform-datapath instructions per tcgen05.mma, 214 cycles per MMA where the tensor core needs 128
        // (tools/mma_probe.cu measures 128.0 with operands, bulk copies and tcgen05.ld all running). =====
        // TWO issuing warps, alternating tiles (warp 1: even tiles, accumulator stage 0; the last warp: odd tiles, stage 1).
        // Between the last MMA of a tile and the first of the next a single issuer has two barrier waits (~90 cycles each
        // even when long complete), a fence and an election to get through, and the tensor core's queue is too shallow to
        // cover them: ~250 idle cycles per 9-MMA tile (moving the waits between the MMAs of the previous tile only moved
        // the bubble).  With two issuers one of them is always past its waits, its first MMA pending at the issue port
        // while the other's tile executes.  Tiles use different accumulator stages and every tcgen05.commit covers the
        // MMAs of its own thread, so the two streams need no ordering between them.
        if (!p.stream_q) mbar_wait(&ctl->qfull, 0);
        const uint32_t ns = (p.T + p.S - 1) / p.S;
        const uint32_t q_lo = plain_desc_lo(smem_u32(q_s));
        const uint32_t a_lo0 = plain_desc_lo(smem_u32(a_s));
        const uint32_t stage_step = stage_bytes >> 4;
        const uint32_t idesc = idesc_bf16();
        // (Only when a stage is a whole tile, ns == 1: an issuer then never asks for a stage fill more than `stages` fills
        // ahead of the ring, which is what keeps a PARITY wait unambiguous.  With several fills per tile the second issuer
        // would start on fill ns while the ring is at fill 0 -- a fresh barrier answers "parity 1 complete" at once -- so
        // there warp 1 issues every tile, as before.)
        const bool dual = ns == 1;
        if (!dual && warp != 1) goto issuer_done;
        {
        const uint32_t first = (dual && warp != 1) ? 1u : 0u, step = dual ? 2u : 1u;
        long long w_tempty = 0, w_full = 0;
        const long long tstart = kDbg ? clock64() : 0;
        for (uint32_t i = first; i < my_tiles; i += step) {
            const uint32_t acc = i & 1, aphase = (i >> 1) & 1;
            const uint32_t d_tmem = tmem + acc * BN;
            long long t0 = (kDbg && p.dbg) ? clock64() : 0;
            mbar_wait(&ctl->tempty[acc], aphase ^ 1);
            if (kDbg && p.dbg) w_tempty += clock64() - t0;
            tc_fence_after();
            uint32_t ks = 0;
            for (uint32_t st = 0; st < ns; ++st) {
                const uint32_t u = i * ns + st;                        // position in the CTA's stream of stage fills
                const uint32_t stage = u % p.stages, phase = (u / p.stages) & 1u;
                const uint32_t nk = p.T - ks < p.S ? p.T - ks : p.S;
                t0 = (kDbg && p.dbg) ? clock64() : 0;
                mbar_wait(&ctl->full[stage], phase);
                if (kDbg && p.dbg) w_full += clock64() - t0;
                tc_fence_after();
                const uint32_t a_lo = a_lo0 + stage * stage_step;
                const uint32_t b_lo = p.stream_q ? a_lo + p.S * (KB_A >> 4) : q_lo + ks * (KB_B >> 4);
                if (elect_one()) {
                    if (!(kDbg && (p.debug_skip & 2))) {
#pragma unroll 1
                        for (uint32_t kk = 0; kk < nk; ++kk)
                            tc_mma_bf16_lo(d_tmem, a_lo + kk * (KB_A >> 4), b_lo + kk * (KB_B >> 4), idesc, (ks + kk) != 0 ? 1u : 0u);
                    }
                    tc_commit(&ctl->empty[stage]);                      // frees the stage when these MMAs have read it
                    if (st + 1 == ns) tc_commit(&ctl->tfull[acc]);      // accumulator complete
                }
                __syncwarp();
                ks += nk;
            }
        }
        if (kDbg && p.dbg && lane == 0 && warp == 1) {
            p.dbg[blockIdx.x * 16 + 5] = w_tempty;
            p.dbg[blockIdx.x * 16 + 6] = w_full;
            p.dbg[blockIdx.x * 16 + 1] = clock64() - tstart;
        }
        }
    issuer_done:;
    } else if (warp == kIssuer2) {
        // (tf32 mode: the second issuer warp of the bf16 mode has nothing to do)
    } else if (warp == 1) {
        // ===== tf32 mode: MMA issuer (one thread) =====
        if (lane == 0) {
            if (!p.stream_q) mbar_wait(&ctl->qfull, 0);
            uint32_t stage = 0, phase = 0;
            long long w_tempty = 0, w_full = 0;
            const long long tstart = kDbg ? clock64() : 0;
            for (uint32_t i = 0; i < my_tiles; ++i) {
                const uint32_t acc = i & 1, aphase = (i >> 1) & 1;
                long long t0 = (kDbg && p.dbg) ? clock64() : 0;
                mbar_wait(&ctl->tempty[acc], aphase ^ 1);
                if (kDbg && p.dbg) w_tempty += clock64() - t0;
                tc_fence_after();
                const uint32_t d_tmem = tmem + acc * BN;
                for (uint32_t s = 0; s < p.nslab; ++s) {
                    t0 = (kDbg && p.dbg) ? clock64() : 0;
                    mbar_wait(&ctl->full[stage], phase);
                    if (kDbg && p.dbg) w_full += clock64() - t0;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(a_s + (size_t)stage * stage_bytes);
                    const uint32_t b_addr = p.stream_q ? a_addr + SLAB_A_BYTES : smem_u32(q_s + (size_t)s * SLAB_B);
                    const uint32_t nk = s + 1 == p.nslab ? p.ksteps_last : 4u;   // 4 MMA K steps of 32 bytes per slab
#pragma unroll
                    for (uint32_t kk = 0; kk < 4; ++kk) {
                        if ((kDbg && (p.debug_skip & 2)) || kk >= nk) break;
                        const uint64_t ad = umma_desc_sw128(a_addr + kk * 32), bd = umma_desc_sw128(b_addr + kk * 32);
                        const uint32_t accum = (s | kk) != 0 ? 1u : 0u;
                        tc_mma_tf32(d_tmem, ad, bd, idesc_tf32(), accum);
                    }
                    tc_commit(&ctl->empty[stage]);   // frees the slab when these MMAs have read it
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                tc_commit(&ctl->tfull[acc]);         // accumulator complete
            }
            if (kDbg && p.dbg) {
                p.dbg[blockIdx.x * 16 + 5] = w_tempty;
                p.dbg[blockIdx.x * 16 + 6] = w_full;
                p.dbg[blockIdx.x * 16 + 1] = clock64() - tstart;
            }
        }
    } else {
        // ===== epilogue warps: TMEM -> registers -> threshold filter -> candidate lists =====
        const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
        const int ew = warp - 2;                     // 0..15
        const int part = ew >> 2;                    // which 64 of the 256 query columns
        unsigned long long *cand = p.cand + (size_t)blockIdx.x * BN * CAP;
        long long w_tfull = 0, w_prune = 0, n_slow = 0, w_park = 0, w_flush = 0, w_bar = 0;
        const long long tstart = kDbg ? clock64() : 0;
        // per-row factor of the NEXT tile is fetched one tile ahead (its global-load latency would otherwise
        // sit on the critical path of every tile)
        // (the raw value is kept and scaled only when used, so nothing waits on the load here)
        auto row_factor = [&](uint32_t r) -> float {
            if (H || r >= p.n) return 0.f;   // bf16 mode: the norm term is inside the contraction
            return COS ? __ldg(p.inv_norm + r) : __ldg(p.sq_norm + r);
        };
        float hx_next = my_tiles ? row_factor(tile_of(0) * BM + quad * 32 + lane) : 0.f;
        StashSlot *slots = ctl->stash[ew];
        const int col0 = part * (BN / 4);
        const float *thr_p = ctl->thr + col0;
        const float ninf = -__int_as_float(0x7f800000);
        for (uint32_t i = 0; i < my_tiles; ++i) {
            const uint32_t tile = tile_of(i);
            const uint32_t acc = i & 1, aphase = (i >> 1) & 1;
            const uint32_t row = tile * BM + quad * 32 + lane;
            const bool rowok = row < p.n;
            const float hx = COS ? hx_next : 0.5f * hx_next;
            if (i + 1 < my_tiles) hx_next = row_factor(tile_of(i + 1) * BM + quad * 32 + lane);
            long long t0 = (kDbg && p.dbg) ? clock64() : 0;
            mbar_wait(&ctl->tfull[acc], aphase);
            if (kDbg && p.dbg) w_tfull += clock64() - t0;
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + acc * BN + col0;
            uint32_t ra[32], rb[32];
            int nst = 0;                                  // parked blocks of this warp in this tile (warp-uniform)
            // A block with candidates.  Usually one lane with one candidate: the lane parks its 32 scores and the warp moves
            // on.  More hitting lanes than free slots (the dense early phases): park in rounds of kStash lanes and empty the
            // stash in between -- that holds the accumulator stage, but only where a CTA has a handful of tiles anyway.
            auto on_hit = [&](const uint32_t (&r)[32], bool hit, int colbase) {
                const unsigned hm = __ballot_sync(kFull, hit);
                const int nh = __popc(hm), rank = __popc(hm & ((1u << lane) - 1u));
                int done = 0;
#pragma unroll 1
                while (done < nh) {
                    if (nst == kStash) {
                        stash_flush<COS, H>(0, nst, slots, ctl, cand, p.qflags, g * BN, p.b, lane);
                        nst = 0;
                    }
                    const int take = min(kStash - nst, nh - done);
                    if (hit && rank >= done && rank < done + take) {
                        StashSlot &e = slots[nst + rank - done];
#pragma unroll
                        for (int j = 0; j < 32; ++j) e.score[j] = __uint_as_float(r[j]);
                        e.row = row;
                        e.colbase = (uint32_t)colbase;
                        e.hx = hx;
                    }
                    nst += take;
                    done += take;
                    __syncwarp();
                }
            };
            if constexpr (DENSE) {
                const int rt = quad * 32 + lane;   // row inside the tile = its position in every list
#pragma unroll 1
                for (int cb = 0; cb < 2; ++cb) {
                    tc_ld32(taddr + cb * 32, ra);
                    tc_wait_ld();
                    unsigned long long *dst = cand + (size_t)(col0 + cb * 32) * CAP + rt;
                    const float *thr_b = thr_p + cb * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float dot = __uint_as_float(ra[j]);
                        const float v = H ? (thr_b[j] - dot) : (COS ? -(dot * hx) : (hx - dot));
                        if (rowok) __stcg(dst + (size_t)j * CAP, make_key(v, row));
                    }
                }
                if (ew == 0 && lane == 0) {
                    const uint32_t left = p.n - tile * BM;
                    ctl->flag = (int)(left < (uint32_t)BM ? left : (uint32_t)BM);   // rows of this tile = entries of every list
                }
            } else if (!(kDbg && (p.debug_skip & 1))) {
                // Hot path: both 32-column blocks of this warp are fetched at once, then a branch-free max filter per block;
                // a block with a candidate (rare) is parked.  Straight-line, a few dozen instructions.
                tc_ld32(taddr, ra);
                tc_ld32(taddr + 32, rb);
                tc_wait_ld();
                // the scores are in registers: the accumulator stage goes back to the MMA warp NOW, before the filter, the
                // parking and the flush -- a warp with candidates is late for its next tile, not for the tensor core
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->tempty[acc]);
                float wa = H ? block_max_h(ra) : block_max<COS>(ra, thr_p, hx, ninf);
                float wb = H ? block_max_h(rb) : block_max<COS>(rb, thr_p + 32, hx, ninf);
                if (kDbg && (p.debug_skip & 4)) {   // ldonly: no scan of the blocks
                    wa = __uint_as_float(ra[0] & ra[31] & 0x80000000u) - 1.f;
                    wb = __uint_as_float(rb[0] & rb[31] & 0x80000000u) - 1.f;
                }
                const bool hita = (H ? wa > 0.f : cand_hit<COS>(wa, hx)) && rowok;
                const bool hitb = (H ? wb > 0.f : cand_hit<COS>(wb, hx)) && rowok;
                if (__builtin_expect(__any_sync(kFull, hita || hitb), 0) && !(kDbg && (p.debug_skip & 16))) {
                    if (kDbg) n_slow++;
                    const long long tp = kDbg ? clock64() : 0;
                    if (__any_sync(kFull, hita)) on_hit(ra, hita, col0);
                    if (__any_sync(kFull, hitb)) on_hit(rb, hitb, col0 + 32);
                    if (kDbg) w_park += clock64() - tp;
                }
            }
            if (DENSE || (kDbg && (p.debug_skip & 1))) {   // (the hot path above has released its stage already)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&ctl->tempty[acc]);
            }

            // the stage is back with the MMA warp: now the parked candidates
            if (nst && !(kDbg && (p.debug_skip & 8))) {
                const long long tf = kDbg ? clock64() : 0;
                __syncwarp();
                stash_flush<COS, H>(0, nst, slots, ctl, cand, p.qflags, g * BN, p.b, lane);
                if (kDbg) w_flush += clock64() - tf;
            }
            // list maintenance every kCheckEvery tiles, per 64-column part: the four warps that append to the lists of a
            // part (one per lane quadrant) meet at the part's own named barrier -- not all sixteen, so a warp that is late
            // after a flush holds up three others, not the whole epilogue.  The lists have room for the appends of the
            // tiles in between (PRUNE_AT).
            if (DENSE || (i + 1) % kCheckEvery != 0) continue;
            const long long tb = kDbg ? clock64() : 0;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
            if (kDbg) w_bar += clock64() - tb;
            if (*(volatile int *)&ctl->pflag[part]) {
                t0 = kDbg ? clock64() : 0;
                for (int q = col0 + quad; q < col0 + BN / 4; q += 4) {
                    if (ctl->cnt[q] > PRUNE_AT / 2) prune_list<H>(cand + (size_t)q * CAP, q, ctl, p.kprime, lane);
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
                if (quad == 0 && lane == 0) ctl->pflag[part] = 0;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
                if (kDbg) w_prune += clock64() - t0;
            }
        }
        if (kDbg && p.dbg && ew == 0 && lane == 0) {
            p.dbg[blockIdx.x * 16 + 7] = clock64() - tstart;
            p.dbg[blockIdx.x * 16 + 0] = n_slow;

            long long tot = 0;
            for (int q = 0; q < BN; ++q) tot += ctl->cnt[q];
            p.dbg[blockIdx.x * 16 + 2] = tot;
            p.dbg[blockIdx.x * 16 + 3] = w_tfull;
            p.dbg[blockIdx.x * 16 + 4] = w_prune;
            p.dbg[blockIdx.x * 16 + 10] = w_park;
            p.dbg[blockIdx.x * 16 + 11] = w_flush;
            p.dbg[blockIdx.x * 16 + 12] = w_bar;

        }
        asm volatile("bar.sync 5, 512;" ::: "memory");   // all epilogue warps (ids 1..4 are the parts' own barriers)
        for (int q = ew * 32 + lane; q < BN; q += 32 * kEpiWarps) {
            const int c = DENSE ? *(volatile int *)&ctl->flag : ctl->cnt[q];
            p.cnt_out[(size_t)blockIdx.x * BN + q] = c < CAP ? c : CAP;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// -------------------------------------------------------------------------------------------------
// bf16 operand mode: mirrors of the rows and of the queries with the folded aux columns
// -------------------------------------------------------------------------------------------------
// Mirror layout per row / query: [kd data columns][16 aux columns], bf16, kd = dim rounded up to 16.
//   rows (A operand)            Euclidean  x_j              | 1 1 1 -h0 -h1 -h2 0...   (h0+h1+h2 = ||x||^2/2 exactly)
//                               cosine     x_j / ||x||      | 1 1 1  0 ...
//   queries (B operand)         Euclidean  q_j              | t0 t1 t2 1 1 1 0...      (t0+t1+t2 = thr_q exactly)
//                               cosine     q_j              | t0 t1 t2 0 ...
// so the contraction over kd + 16 columns is D = dot~ + thr_q - ||x||^2/2 (cosine: dot^ + thr_q): the candidate
// test "surrogate < thr_q" is D > 0 and the epilogue needs neither the thresholds nor the row norms.  An f32 splits
// into three bf16 parts exactly (3 x 8 significant bits); products with 1.0 are exact in the tensor core.
__device__ __forceinline__ float split3(float v, unsigned short (&t)[3]) {
    const __nv_bfloat16 a = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(a);
    const __nv_bfloat16 b = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b);
    const __nv_bfloat16 c = __float2bfloat16_rn(r2);
    t[0] = __bfloat16_as_ushort(a);
    t[1] = __bfloat16_as_ushort(b);
    t[2] = __bfloat16_as_ushort(c);
    return (__bfloat162float(a) + __bfloat162float(b)) + __bfloat162float(c);   // the value the parts stand for
}
__device__ __forceinline__ unsigned short bf16_bits(float v) { return __bfloat16_as_ushort(__float2bfloat16_rn(v)); }
constexpr unsigned short kBf16One = 0x3F80;

// Byte offset of 8-column chunk `chunk` (K step chunk / 2, half chunk % 2) of row `r` inside a tile of R rows whose
// K-step blocks are R * 32 bytes: [K step][r / 8][chunk % 2][r % 8][16 bytes].
__host__ __device__ __forceinline__ size_t tiled_offset(uint32_t r, uint32_t chunk, uint32_t R) {
    return (size_t)(chunk >> 1) * R * 32 + (size_t)(r >> 3) * kPlainSBO + (size_t)(chunk & 1) * kPlainLBO + (size_t)(r & 7) * 16;
}

// Tiles [tile0, tile0 + ntiles) of the shard -> tiled mirror.  A warp converts 8 rows x 4 chunks per step: lane l
// reads 32 bytes (8 floats: one full sector) of row l % 8 and writes the 16-byte bf16 chunk, so each group of 8 lanes
// stores one contiguous 128-byte core matrix.  Rows at and beyond n_valid (the tail of the last tile, rows not
// inserted yet) become zero rows: D = 0 for them, never a candidate.
template <bool COS>
__global__ void __launch_bounds__(256) build_mirror_kernel(const float *__restrict__ rows, const float *__restrict__ sq_norm,
                                                           const float *__restrict__ inv_norm, uint32_t tile0, uint32_t ntiles, uint32_t n_valid,
                                                           uint32_t ld, uint32_t kd, uint32_t T, unsigned char *__restrict__ out,
                                                           unsigned int *__restrict__ stats) {
    const int lane = threadIdx.x & 31, rr = lane & 7, cc = lane >> 3;
    float worst_err = 0.f, worst_len = 0.f;   // max over this lane's rows of ||x~ - x||^2 and ||x~||^2
    const uint32_t wpg = gridDim.x * (blockDim.x >> 5);
    const uint32_t groups = ntiles * (BM / 8);          // 8-row groups to convert
    const uint32_t nchunk = T * 2, aux = kd / 8;
    for (uint32_t gi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); gi < groups; gi += wpg) {
        const uint32_t tile = tile0 + gi / (BM / 8), rt = (gi % (BM / 8)) * 8 + rr;   // row inside the tile
        const uint32_t r = tile * BM + rt;
        const bool live = r < n_valid;
        const float *x = rows + (size_t)r * ld;
        const float scale = (COS && live) ? __ldg(inv_norm + r) : 1.f;
        unsigned short h[3] = {0, 0, 0};
        if (!COS && live) split3(0.5f * __ldg(sq_norm + r), h);
        unsigned char *tbase = out + (size_t)tile * T * KB_A;
        float err2 = 0.f, len2 = 0.f;   // this lane's share of the row: what the rounding to bf16 changed, what is left
        for (uint32_t ch = cc; ch < nchunk; ch += 4) {
            unsigned short v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const uint32_t c0 = ch * 8;
            if (live) {
                if (ch < aux) {
#pragma unroll
                    for (int g4 = 0; g4 < 2; ++g4) {
                        const uint32_t c = c0 + g4 * 4;
                        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c < ld) f = __ldg(reinterpret_cast<const float4 *>(x + c));   // ld % 4 == 0; columns >= dim hold zeros
                        const float e[4] = {f.x * scale, f.y * scale, f.z * scale, f.w * scale};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const __nv_bfloat16 hb = __float2bfloat16_rn(e[u]);
                            const float back = __bfloat162float(hb), d = back - e[u];   // (exact: the two are within a factor 2)
                            v[g4 * 4 + u] = __bfloat16_as_ushort(hb);
                            err2 = fmaf(d, d, err2);
                            len2 = fmaf(back, back, len2);
                        }
                    }
                } else if (ch == aux) {   // aux columns (kd % 16 == 0: an 8-column chunk is all data or all aux)
                    v[0] = v[1] = v[2] = kBf16One;
                    if (!COS) {
                        v[3] = h[0] ^ 0x8000;
                        v[4] = h[1] ^ 0x8000;
                        v[5] = h[2] ^ 0x8000;
                    }
                }
            }
            uint4 w;
            w.x = v[0] | ((uint32_t)v[1] << 16);
            w.y = v[2] | ((uint32_t)v[3] << 16);
            w.z = v[4] | ((uint32_t)v[5] << 16);
            w.w = v[6] | ((uint32_t)v[7] << 16);
            *reinterpret_cast<uint4 *>(tbase + tiled_offset(rt, ch, BM)) = w;
        }
        // the four lanes that share a row (same lane % 8)
        err2 += __shfl_xor_sync(kFull, err2, 8);
        err2 += __shfl_xor_sync(kFull, err2, 16);
        len2 += __shfl_xor_sync(kFull, len2, 8);
        len2 += __shfl_xor_sync(kFull, len2, 16);
        worst_err = fmaxf(worst_err, err2);
        worst_len = fmaxf(worst_len, len2);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        worst_err = fmaxf(worst_err, __shfl_xor_sync(kFull, worst_err, o));
        worst_len = fmaxf(worst_len, __shfl_xor_sync(kFull, worst_len, o));
    }
    if (lane == 0) {   // non-negative floats order like their bit patterns
        atomicMax(stats + 0, __float_as_uint(worst_err));
        atomicMax(stats + 1, __float_as_uint(worst_len));
    }
}

// |x~ . q~ - x . q| for the bf16 mirrors x~, q~ of a row x and a query q, from what the roundings actually changed:
//   x~ . q~ - x . q = x~ . (q~ - q) + (x~ - x) . q    so    <= max||x~|| * ||q~ - q|| + max||x~ - x|| * ||q||
// with the two maxima over the mirrored rows measured by build_mirror_kernel (mirror_stats) and ||q~ - q|| by whoever
// asks.  Rounding to nearest with 8 significant bits changes a component by at most 2^-8 of its magnitude and by ~0.4 of
// that in the root mean square, so on ordinary data this is ~2.2x below the worst case 2^-7 ||x|| ||q|| (every component
// of both operands on a rounding boundary, all errors aligned) that round 1 charged -- and it IS that worst case on
// data built to reach it (tests/test_error_bounds.py).  (The norms are f32 sums: inflated by 1e-5.)
__device__ __forceinline__ float operand_error(const unsigned int *mirror_stats, float q_norm, float dq_norm) {
    const float err = sqrtf(__uint_as_float(mirror_stats[0])) * 1.00001f, len = sqrtf(__uint_as_float(mirror_stats[1])) * 1.00001f;
    return (len * dq_norm + err * q_norm) * 1.00002f;
}

// Queries of one wave -> tiled mirror (groups of BN queries), with the START threshold: no finite k'-th key exists
// yet, so the threshold is a finite cap above every possible surrogate (every row passes) -- an infinite one would
// poison D = dot + thr - hx.
//   Euclidean  v = ||x||^2/2 - dot  <=  M/2 + sqrt(M) ||q||   (M = max ||x||^2 of the shard)
//   cosine     v = -dot/||x||       <=  ||q||
// One warp per query slot of the wave's groups; slots beyond b become zero rows.
template <bool COS>
__global__ void __launch_bounds__(256) prep_queries_kernel(const float *__restrict__ q, uint32_t b, uint32_t slots, uint32_t ld, uint32_t kd, uint32_t T,
                                                           const unsigned int *__restrict__ maxnorm_bits, unsigned char *__restrict__ qt,
                                                           float *__restrict__ gthr, float *__restrict__ qcap, float cap_sign,
                                                           float *__restrict__ qband, float eps_dot, float acc_eps,
                                                           const unsigned int *__restrict__ mirror_stats) {
    const int lane = threadIdx.x & 31;
    const uint32_t qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= slots) return;
    const bool live = qi < b;
    const float *x = q + (size_t)qi * ld;
    float nq = 0.f, dq2 = 0.f;   // ||q||^2 and ||bf16(q) - q||^2
    if (live)
        for (uint32_t c = lane; c < ld; c += 32) {
            const float d = __bfloat162float(__float2bfloat16_rn(x[c])) - x[c];
            nq = fmaf(x[c], x[c], nq);
            dq2 = fmaf(d, d, dq2);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nq += __shfl_xor_sync(kFull, nq, o);
        dq2 += __shfl_xor_sync(kFull, dq2, o);
    }
    const float nqs = sqrtf(nq) * 1.0001f;
    const float M = __uint_as_float(*maxnorm_bits);
    // (cap_sign = -1 only in the VROD_BATCHED_DEBUG timing modes: no row is ever a candidate)
    const float cap = cap_sign * ((COS ? nqs : fmaf(sqrtf(M), nqs, 0.5f * M)) * 1.02f + 1e-30f);
    unsigned short t[3];
    const float thr = split3(cap, t);
    unsigned char *gbase = qt + (size_t)(qi / BN) * T * KB_B;
    const uint32_t rq = qi % BN, aux = kd / 8;
    for (uint32_t ch = lane; ch < T * 2; ch += 32) {
        unsigned short v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (live) {
            if (ch < aux) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint32_t c = ch * 8 + e;
                    v[e] = c < ld ? bf16_bits(x[c]) : 0;
                }
            } else if (ch == aux) {
                v[0] = t[0];
                v[1] = t[1];
                v[2] = t[2];
                if (!COS) v[3] = v[4] = v[5] = kBf16One;
            }
        }
        uint4 w;
        w.x = v[0] | ((uint32_t)v[1] << 16);
        w.y = v[2] | ((uint32_t)v[3] << 16);
        w.z = v[4] | ((uint32_t)v[5] << 16);
        w.w = v[6] | ((uint32_t)v[7] << 16);
        *reinterpret_cast<uint4 *>(gbase + tiled_offset(rq, ch, BN)) = w;
    }
    if (lane == 0 && live) {
        gthr[qi] = thr;
        qcap[qi] = thr;
        if (qband) {
            // twice the error bound of the approximate surrogate (the guard's E in batched_finish_kernel, with |u| <= cap),
            // rounded up generously: it is a width, not a proof -- the proof stays with the guard
            const float Ms = sqrtf(M * 1.0000002f);
            const float Ed = operand_error(mirror_stats, nqs, sqrtf(dq2));
            const float E = COS ? Ed + (eps_dot + 4.2e-7f) * nqs + acc_eps * (2.01f * nqs + thr)
                                : Ed + (eps_dot + 2.4e-7f) * Ms * nqs + 2.4e-7f * (0.5f * M + thr) + acc_eps * (Ms * nqs + 0.5f * M + 2.f * thr);
            qband[qi] = 2.1f * E + 1e-30f;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// finish: one CTA per query -- merge the CTA lists, exact rerank, guard, hits
// -------------------------------------------------------------------------------------------------
struct FinishParams {
    const float4 *rows4;
    const float4 *q4;       // [b][ld4]
    uint32_t n, ld4, b, k;
    uint32_t rank, world;   // shard coordinates (row_id)
    int kprime, cap, water;
    uint32_t qgroups, units;        // list layout of the tile kernel: CTA u = g + m*qgroups serves query group g
    const unsigned long long *cand;
    const int *cnt_in;
    int *qflags;                        // [b] |= 1: the query cannot be proven from its lists (in: tile kernels; out: a merge that overran)
    const unsigned int *maxnorm_bits;   // max ||x||^2 of the shard, f32 bits
    unsigned long long *glist;          // [b][kprime] best keys over the phases so far (in/out)
    int *gcnt;                          // [b]
    float *gthr;                        // [b] out: threshold for the next phase
    int final_phase;                    // 0: only merge + publish glist/gthr; 1: rerank, guard, hits
    int first_phase;                    // 1: glist is empty
    int *status;
    Hit *out;
    unsigned long long *out_ids;   // optional: final [b][k] ids / distances (single-GPU contexts skip the merge kernel)
    float *out_dist;
    double eps_dot;         // error of the approximate dot product relative to ||x|| ||q||: tf32 mode operand truncation + f32 accumulation;
                            // bf16 mode the accumulation only -- the operand roundings are measured (mirror_stats, operand_error)
    const unsigned int *mirror_stats;   // bf16 mode: see operand_error; nullptr in tf32 mode
    // bf16 operand mode (qh != nullptr): the next phase's threshold goes into the query mirror's aux columns
    unsigned char *qt;      // tiled query mirror of the wave (prep_queries_kernel)
    uint32_t ld_h, kd, T;
    const float *qcap;      // [b] the finite start threshold (above every surrogate)
    double acc_eps;         // f32 accumulation noise of the folded contraction, relative to |dot| + |thr| + |hx|
    // BAND mode (keep > kprime): instead of the kprime best approximate keys a query keeps EVERY key within `band` of its
    // k-th best one -- twice the error bound of the approximate surrogate, so the exact top k is among them by
    // construction (tight clusters: whatever the contraction cannot tell apart becomes a candidate, up to `keep` of them)
    int need, keep;         // rank of the key the threshold hangs on (kprime, band mode: k); stride and capacity of glist
    const float *qband;     // [b] band width per query (nullptr: classic mode)
    // GUESSED thresholds (launch_batched_search): the next phase does not filter at the kprime-th best key seen so far but
    // at the guess_rank-th (< kprime), the rank below which the next phase is expected to find its kprime keys.  The finish
    // of a phase that ran under a guess checks that it did (prev_guess): the new kprime-th key must not lie above the
    // threshold the rows were filtered with, or rows between the two were lost and the query is flagged (qflags bit 2).
    int guess_rank, prev_guess;
};

constexpr int kFinCtl = 128;
constexpr int kFinCap = 2048;                       // keys the finish kernel's selection buffer holds
constexpr int kFinPer = kFinCap / kScanThreads;     // keys per thread in block_select
constexpr int kFinHist = 264;                       // ints: 256 bins + selected bin, rank, size + the threshold key

template <bool COS>
__global__ void __launch_bounds__(kScanThreads, 4) batched_finish_kernel(const FinishParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    CandCtl *ctl = reinterpret_cast<CandCtl *>(smem);
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(smem + kFinCtl);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t qi = blockIdx.x;
    const uint32_t g = qi / BN, ql = qi % BN;
    const float4 *q4 = p.q4 + (size_t)qi * p.ld4;
    const uint32_t nlists = (p.units - g + p.qgroups - 1) / p.qgroups;   // CTAs that scanned rows for this query
    auto cta_of = [&](uint32_t l) { return g + l * p.qgroups; };
    if (tid == 0) {
        cand_reset(ctl);
        ctl->overflow = 0;
    }
    if (warp == 1 && p.final_phase) {   // only the last phase reranks
        const double nq = canon_row_sum<2>(q4, q4, (int)p.ld4, lane);
        if (lane == 0) ctl->nq = nq;
    }
    if (warp == 2 && p.final_phase && p.mirror_stats) {   // what rounding the query to bf16 changed (guard)
        const float *qf = reinterpret_cast<const float *>(q4);
        float dq2 = 0.f;
        for (uint32_t c = lane; c < p.ld4 * 4; c += 32) {
            const float d = __bfloat162float(__float2bfloat16_rn(qf[c])) - qf[c];
            dq2 = fmaf(d, d, dq2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dq2 += __shfl_xor_sync(kFull, dq2, o);
        if (lane == 0) ctl->dq2 = dq2;
    }
    // offsets of this query's lists (previous global list first, then one list per CTA of the group):
    // warp 0 loads the counts and scans them 32 at a time
    int *hist = reinterpret_cast<int *>(buf + p.cap);   // [kFinHist] block_select scratch
    const float band = p.qband ? p.qband[qi] : 0.f;
    int *offs = hist + kFinHist;                        // [nlists + 2]
    // (all threads fetch the counts -- one memory latency --, warp 0 turns them into offsets)
    for (uint32_t m = tid; m < nlists; m += kScanThreads) offs[m + 2] = p.cnt_in[(size_t)cta_of(m) * BN + ql];
    __syncthreads();
    if (warp == 0) {
        int carry = p.first_phase ? 0 : p.gcnt[qi];
        if (lane == 0) {
            offs[0] = 0;
            offs[1] = carry;
        }
        for (uint32_t m0 = 0; m0 < nlists; m0 += 32) {
            const uint32_t m = m0 + lane;
            int v = m < nlists ? offs[m + 2] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, v, o);
                if (lane >= o) v += t;
            }
            if (m < nlists) offs[m + 2] = carry + v;
            carry += __shfl_sync(kFull, v, 31);
        }
    }
    __syncthreads();
    const int total = offs[nlists + 1];
    if (total <= p.cap) {
        // common case: everything fits the sort buffer -- every thread fetches its share of ALL lists at once
        // (binary search of the entry's list in offs), so the gather costs one memory latency, then one select
        const int prev = offs[1];
        for (int i = tid; i < total; i += kScanThreads) {
            if (i < prev) {
                buf[i] = p.glist[(size_t)qi * p.keep + i];
            } else {
                uint32_t lo = 0, hi = nlists - 1;   // largest m with offs[m + 1] <= i
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1) >> 1;
                    if (offs[mid + 1] <= i) lo = mid;
                    else hi = mid - 1;
                }
                buf[i] = __ldcg(p.cand + ((size_t)cta_of(lo) * BN + ql) * CAP + (i - offs[lo + 1]));
            }
        }
        __syncthreads();
        if (tid == 0) ctl->cnt = total;
    } else {
        // more entries than the sort buffer holds (phase 0, or a candidate-heavy phase): stream them through the
        // threshold filter in rounds of as many entries as fit next to the kprime survivors
        const int prev = offs[1];
        for (int i = tid; i < prev; i += kScanThreads) buf[i] = p.glist[(size_t)qi * p.keep + i];
        __syncthreads();
        if (tid == 0) ctl->cnt = prev;
        __syncthreads();
        block_select<kFinPer>(ctl, buf, p.need, p.cap, tid, hist, band, p.qband ? p.keep : 0);
        // rounds of as many whole LISTS as the buffer has room for (a list holds at most CAP = cap/2 keys and a
        // select runs whenever the buffer is more than half full, so at least one list always fits).  A warp takes
        // a list at a time, its lanes read consecutive keys (4 independent loads each) -- no per-key search of the
        // owning list.  After the first select the threshold passes only a small part of a round, so most rounds
        // are a plain load + filter.
        static_assert(CAP * 2 <= kFinCap, "a full CTA list must fit the free half of the selection buffer");
        uint32_t m0 = 0;
        bool all_at_once = true;   // once a threshold exists: ALL remaining lists in one round (one memory latency instead of
                                   // one per ~14 lists); if more keys pass than the buffer holds, the round is undone and the
                                   // lists go through in safe portions after all
        while (m0 < nlists) {
            const int held = ctl->cnt;           // uniform: read between two barriers
            const int room = p.cap - held;
            const unsigned long long thr = *(volatile unsigned long long *)&ctl->thrkey;
            __syncthreads();
            uint32_t lo = m0 + 1, hi = nlists;   // largest m1 in (m0, nlists] with offs[m1 + 1] - offs[m0 + 1] <= room
            while (lo < hi) {
                const uint32_t mid = (lo + hi + 1) >> 1;
                if (offs[mid + 1] - offs[m0 + 1] <= room) lo = mid;
                else hi = mid - 1;
            }
            const bool optimistic = all_at_once && thr != kKeyMax && lo < nlists;
            const uint32_t m1 = optimistic ? nlists : lo;
            // a warp takes FOUR lists at a time, two keys of each per lane and step: 8 independent loads in flight per
            // lane (one list at a time left the warp waiting out a memory latency per list: 17 in a row for the 148 lists
            // of a single query group's start phase)
            constexpr int kLists = 4;
            for (uint32_t mb = m0 + warp; mb < m1; mb += kLists * kScanWarps) {
                int n_m[kLists], longest = 0;
                const unsigned long long *src[kLists];
#pragma unroll
                for (int e = 0; e < kLists; ++e) {
                    const uint32_t m = mb + e * kScanWarps;
                    n_m[e] = m < m1 ? offs[m + 2] - offs[m + 1] : 0;
                    src[e] = p.cand + ((size_t)cta_of(m < m1 ? m : mb) * BN + ql) * CAP;
                    longest = n_m[e] > longest ? n_m[e] : longest;
                }
                for (int j0 = lane; j0 < longest; j0 += 64) {
                    unsigned long long keys[kLists][2];
#pragma unroll
                    for (int e = 0; e < kLists; ++e) {
                        keys[e][0] = j0 < n_m[e] ? __ldcg(src[e] + j0) : kKeyMax;
                        keys[e][1] = j0 + 32 < n_m[e] ? __ldcg(src[e] + j0 + 32) : kKeyMax;
                    }
#pragma unroll
                    for (int e = 0; e < kLists; ++e)
#pragma unroll
                        for (int u = 0; u < 2; ++u)
                            if (keys[e][u] < thr) {
                                const int pos = atomicAdd(&ctl->cnt, 1);
                                if (pos < p.cap) buf[pos] = keys[e][u];
                            }
                }
            }
            __syncthreads();
            if (ctl->cnt > p.cap) {              // uniform
                __syncthreads();
                if (tid == 0) {
                    ctl->cnt = held;             // (what the round wrote above `held` is simply forgotten)
                    if (!optimistic) ctl->overflow = 1;   // cannot happen: a safe portion fits by construction
                }
                __syncthreads();
                if (optimistic) {
                    all_at_once = false;
                    continue;
                }
            }
            m0 = m1;
            const bool no_threshold_yet = ctl->thrkey == kKeyMax && ctl->cnt >= p.need;
            if (m0 < nlists && (ctl->cnt > p.cap / 2 || no_threshold_yet)) block_select<kFinPer>(ctl, buf, p.need, p.cap, tid, hist, band, p.qband ? p.keep : 0);   // uniform
        }
    }
    __syncthreads();
    block_select<kFinPer>(ctl, buf, p.need, p.cap, tid, hist, band, p.qband ? p.keep : 0);

    const int ncand = ctl->cnt;
    if (p.prev_guess && tid == 0) {
        // the phase just merged was filtered at a guessed threshold: it holds if at least `need` keys lie at or below it
        const float kth = ctl->thrkey != kKeyMax ? ord2f((uint32_t)(ctl->thrkey >> 32)) : __int_as_float(0x7f800000);
        if (!(kth <= p.gthr[qi])) atomicOr(p.qflags + qi, 2);
    }
    if (!p.final_phase) {
        for (int i = tid; i < ncand; i += kScanThreads) p.glist[(size_t)qi * p.keep + i] = buf[i];
        const bool guess = p.guess_rank > 0 && p.guess_rank < ncand && ctl->thrkey != kKeyMax;   // uniform
        if (guess) {
            // the list is stored; the buffer may now be cut down to the guess_rank best keys to find the guess_rank-th
            __syncthreads();
            block_select<kFinPer>(ctl, buf, p.guess_rank, p.cap, tid, hist);
        }
        if (tid == 0) {
            p.gcnt[qi] = ncand;
            // a merge that could not hold everything it had to keep (band mode: more rows inside the error band than the list
            // has room for) has dropped keys: the flag must outlive this launch, the last phase cannot know
            if (ctl->overflow) atomicOr(p.qflags + qi, 1);
            // (the kept keys are unsorted: the kprime-th one -- or the guess_rank-th -- is the select's threshold key)
            float thr = ctl->thrkey != kKeyMax ? ord2f((uint32_t)(ctl->thrkey >> 32)) : __int_as_float(0x7f800000);
            if (p.qt) {   // bf16 mode: finite thresholds only, folded into the query mirror as three exact parts
                const float cap = p.qcap[qi];
                if (!(thr < cap)) thr = cap;
                unsigned short t[3];
                thr = split3(thr, t);
                // the first three columns of the aux chunk of this query's row in its group's tile
                unsigned short *aux = reinterpret_cast<unsigned short *>(p.qt + (size_t)(qi / BN) * p.T * KB_B + tiled_offset(qi % BN, p.kd / 8, BN));
                aux[0] = t[0];
                aux[1] = t[1];
                aux[2] = t[2];
            }
            p.gthr[qi] = thr;
        }
        return;
    }
    // u = the largest kept approximate key: the select's threshold key, or the tail of the (sorted) short list
    if (tid == 0) ctl->u_val = ctl->thrkey != kKeyMax ? ord2f((uint32_t)(ctl->thrkey >> 32)) : (ncand > 0 ? ord2f((uint32_t)(buf[ncand - 1] >> 32)) : 0.f);
    __syncthreads();
    rerank_candidates<COS>(buf, ncand, p.rows4, q4, (int)p.ld4, ctl->nq, warp, lane);
    __syncthreads();
    {
        int P = 32;
        while (P < ncand) P <<= 1;
        for (int i = ncand + tid; i < P; i += kScanThreads) buf[i] = kKeyMax;
        __syncthreads();
        block_bitonic(buf, P, tid);
    }
    for (int i = tid; i < (int)p.k; i += kScanThreads) {
        Hit h;
        if (i < ncand) {
            h.id = row_id((uint32_t)buf[i], p.rank, p.world);
            h.dist = ord2f((uint32_t)(buf[i] >> 32));
        } else {
            h.id = kKeyMax;
            h.dist = __int_as_float(0x7f800000);
        }
        h.pad = 0;
        p.out[(size_t)qi * p.k + i] = h;
        if (p.out_ids) {
            p.out_ids[(size_t)qi * p.k + i] = h.id;
            p.out_dist[(size_t)qi * p.k + i] = h.dist;
        }
    }
    if (tid == 0) {
        const int qf = p.qflags[qi];   // (this thread's own atomicOr above included)
        int bad = (ctl->overflow | (qf & 1)) ? 1 : 0;
        if (qf & 2) bad |= 16;
        if (p.n > (uint32_t)ncand) {
            const int kk = (int)p.k < ncand ? (int)p.k : ncand;
            // (no candidate at all -- e.g. a query the tensor-core pass could not represent: flagged below, kk < k)
            const float T = kk > 0 ? ord2f((uint32_t)(buf[kk - 1] >> 32)) : __int_as_float(0x7f800000);
            const double u = (double)ctl->u_val;
            const double xn_max = (double)__uint_as_float(*p.maxnorm_bits) * (1.0 + 1.2e-7);
            const double nq = ctl->nq, nqs = __dsqrt_rn(nq);
            float lb;
            if (!(u == u) || fabs(u) > 3.0e38) {
                lb = -1.f;
            } else if constexpr (COS) {
                // v = -dot~ * inv~ ;  |v - (-dot/||x||)| <= (eps_dot + 3*2^-24) * ||q||
                double E = (p.eps_dot + 3.0e-7) * nqs + 1.2e-7 * fabs(u);
                if (p.mirror_stats) E += (double)operand_error(p.mirror_stats, (float)nqs * 1.0000002f, sqrtf(ctl->dq2)) * 1.000001;
                if (p.qt) E += p.acc_eps * (1.01 * nqs + (double)p.qcap[qi] + fabs(u));
                lb = nqs > 0.0 ? __double2float_rd(1.0 + (u - E) / nqs - 1.0e-12) : -1.f;
            } else {
                // v = hx~ - dot~ ;  |v - (||x||^2/2 - dot)| <= eps_dot*||x||max*||q|| + 2^-23*(||x||max^2/2 + |u|)
                double E = (p.eps_dot + 2.4e-7) * __dsqrt_rn(xn_max) * nqs + 2.4e-7 * (0.5 * xn_max + fabs(u));
                if (p.mirror_stats) E += (double)operand_error(p.mirror_stats, (float)nqs * 1.0000002f, sqrtf(ctl->dq2)) * 1.000001;
                if (p.qt) E += p.acc_eps * (__dsqrt_rn(xn_max) * nqs + 0.5 * xn_max + (double)p.qcap[qi] + fabs(u));
                double s = 2.0 * (u - E) + nq;
                s -= 1.0e-12 * (fabs(s) + nq);
                lb = s > 0.0 ? __double2float_rd(__dsqrt_rd(s)) : 0.f;
                if (!(s > 0.0)) lb = -1.f;
            }
            if (!(lb > T)) bad |= 4;
            if (kk < (int)p.k) bad |= 8;
            if (kDbg && bad && qi < 4)
                printf("[finish dbg] q %u: bad %d (1 overflow/flag, 4 bound, 8 short) ncand %d u %.6g lb %.6g T %.6g nq %.6g xn_max %.6g\n", qi, bad,
                       ncand, u, (double)lb, (double)T, nq, xn_max);
        }
        p.status[qi] = bad ? ((bad & 16) ? 2 : 1) : 0;   // 2: a guessed threshold did not hold (the caller counts those)
    }
}

// -------------------------------------------------------------------------------------------------
// host
// -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// 2D tensor [rows][ld] of f32, box = {one 128-byte swizzle row, box_rows}, zero fill out of bounds (tf32 mode)
bool make_map(CUtensorMap *m, const void *base, uint64_t rows, uint32_t ld, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn || rows == 0) return false;
    cuuint64_t dims[2] = {ld, rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)KS, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

long long *g_dbg_buf = nullptr;

}  // namespace

uint32_t mirror_kd(uint32_t dim) { return (dim + 15u) & ~15u; }
uint32_t mirror_ld(uint32_t dim) { return mirror_kd(dim) + AUX_H; }

size_t mirror_bytes(uint64_t rows, uint32_t dim) { return (size_t)((rows + BM - 1) / BM) * (mirror_ld(dim) / UMMA_K_H) * KB_A; }

cudaError_t launch_build_mirror(const ShardView &s, unsigned short *rows_h, uint32_t row0, uint32_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    // whole tiles: the tile that holds row0 is rebuilt from its start; the rows of the last tile beyond s.n become zero rows
    const uint32_t tile0 = row0 / BM, tile1 = (row0 + n + BM - 1) / BM, ntiles = tile1 - tile0;
    const uint32_t want = (ntiles * (BM / 8) + 7) / 8;
    const uint32_t blocks = want < 148u * 16u ? want : 148u * 16u;
    auto fn = s.metric ? build_mirror_kernel<true> : build_mirror_kernel<false>;
    fn<<<blocks, 256, 0, st>>>(s.rows, s.sq_norm, s.inv_norm, tile0, ntiles, s.n, s.ld, mirror_kd(s.dim), mirror_ld(s.dim) / UMMA_K_H,
                               reinterpret_cast<unsigned char *>(rows_h), s.mirror_stats);
    return cudaGetLastError();
}

bool batched_supported(const ShardView &s, uint32_t b, uint32_t k) {
    if (s.n == 0 || b == 0) return false;
    if (s.ld > 4096) return false;
    if (k > 120) return false;                            // k' = 1.5k + 16 (rounded up to 32s) must stay <= 256 (CAP / 4)
    return encode_fn() != nullptr;
}

cudaError_t launch_batched_search(const ShardView &s, const float *d_q, uint32_t b, uint32_t k, int sm_count, void **scratch,
                                  size_t *scratch_bytes, int *status, Hit *out, unsigned long long *out_ids, float *out_dist,
                                  cudaStream_t st, BatchedStats *stats, cudaEvent_t ev_start, cudaEvent_t ev_stop, bool band_mode,
                                  bool guess_mode, bool wide_margin) {
    // operand mode: bf16 mirrors with folded thresholds when the collection has a mirror, else the stored f32 rows as tf32
    const bool H = s.rows_h != nullptr;
    const uint32_t kd = mirror_kd(s.dim), ld_h = mirror_ld(s.dim);
    const uint32_t nslab = (s.ld + KS - 1) / KS;                                               // tf32 mode: 128-byte K slabs
    const uint32_t ksteps_last = (s.ld - (nslab - 1) * KS + UMMA_K - 1) / UMMA_K;
    const uint32_t T = ld_h / UMMA_K_H;                                                        // bf16 mode: K steps per tile
    const uint32_t ntiles = (s.n + BM - 1) / BM;
    uint32_t qgroups = (b + BN - 1) / BN;
    // one CTA per SM; query groups beyond the SM count are handled in waves
    cudaError_t e = cudaSuccess;
    // k' = 1.5 k + 16 rounded up to a multiple of 32: the keys between the k-th and the k'-th are the margin the proof needs
    // (the bf16 surrogate cannot order rows closer than its error bound).  Round 1 used pow2 >= 2k + 16 (256 for k = 100)
    // with the worst-case operand error; with the measured one (operand_error) 160 already proves every query of
    // configs[2] and of the 1M-row sweep, 128 does not (27 of 256 fail at 1M x 128): 192 it is, and 32 for k = 10.
    // Collections whose proofs fail at that margin (distances that concentrate: 1M x 1536) get round 1's wider one
    // (wide_margin: the caller switches per collection) before they fall back to band mode.
    int kprime = ((int)(k + k / 2 + 16) + 31) / 32 * 32;
    if (wide_margin) {
        kprime = 64;
        while (kprime < (int)(2 * k + 16)) kprime <<= 1;
    }
    {   // tuning knob (any value in [k + 8, 1024] is correct: a k' that is too small only makes more proofs fail)
        static const int kp_env = getenv("VROD_BATCHED_KPRIME") ? atoi(getenv("VROD_BATCHED_KPRIME")) : 0;
        if (kp_env >= (int)k + 8 && kp_env <= CAP) kprime = kp_env;
    }
    // |dot~ - dot| <= eps_dot ||x|| ||q||: tf32 truncates both operands to 10 stored mantissa bits (< 2^-10 each,
    // 2^-9 for the product); bf16 keeps 7 stored bits and the mirrors are rounded to nearest (<= 2^-8 each, 2^-7 for
    // the product); plus the f32 accumulation inside an MMA step.  tests/test_error_bounds.py checks both budgets
    // on the CPU, including operands built to sit just under the rounding boundary in every component.
    // bf16 mode: the operand roundings are not charged at their worst case (2^-7) but as measured (operand_error); eps_dot
    // then only covers the accumulation.
    const double eps_dot = H ? (double)ld_h * ldexp(1.0, -22) : ldexp(1.0, -9) * 1.01 + (double)s.ld * ldexp(1.0, -22);

    // band mode (bf16 operand mode only): every key within the error band above the k-th best approximate key is kept, up to
    // kBandKeep per query
    constexpr int kBandKeep = 1024;
    const bool band = band_mode && H && (int)k <= kBandKeep / 2;
    const int keep = band ? kBandKeep : kprime;
    const uint32_t max_units = (uint32_t)sm_count;   // one CTA (= one unit) per SM
    const uint32_t super_tiles = ntiles;

    for (uint32_t g0 = 0; g0 < qgroups; g0 += max_units) {
        const uint32_t groups = qgroups - g0 < max_units ? qgroups - g0 : max_units;
        // the same number of units for every query group: the units of different groups that need the same row
        // tile then run in lockstep and all but one of them hit L2
        uint32_t units = groups * (max_units / groups);
        if ((unsigned long long)groups * super_tiles < units) units = groups * super_tiles;   // no more units than work
        const uint32_t cpg_max = (units + groups - 1) / groups;
        const uint32_t grid = units;
        const uint32_t bq = b - g0 * BN < groups * BN ? b - g0 * BN : groups * BN;   // queries in this wave
        const float *qw = d_q + (size_t)g0 * BN * s.ld;

        // scratch: cand [grid][BN][CAP] u64, cnt [grid][BN] int, qflags [bq] int, maxnorm is part of the shard
        const size_t cand_bytes = (size_t)grid * BN * CAP * sizeof(unsigned long long);
        const size_t cnt_bytes = (size_t)grid * BN * sizeof(int);
        const size_t flag_bytes = (((size_t)bq * sizeof(int)) + 255) & ~(size_t)255;
        const size_t glist_bytes = (size_t)bq * keep * sizeof(unsigned long long);
        const size_t qh_bytes = H ? (size_t)groups * T * KB_B : 0;                    // tiled query mirror of the wave
        const size_t need = cand_bytes + cnt_bytes + 5 * flag_bytes + glist_bytes + qh_bytes + 2048;
        if (*scratch_bytes < need) {
            if (*scratch) cudaFree(*scratch);
            *scratch = nullptr;
            *scratch_bytes = 0;
            e = cudaMalloc(scratch, need);
            if (e != cudaSuccess) return e;
            *scratch_bytes = need;
        }
        unsigned char *base = reinterpret_cast<unsigned char *>(*scratch);
        unsigned long long *cand = reinterpret_cast<unsigned long long *>(base);
        int *cnt = reinterpret_cast<int *>(base + cand_bytes);
        int *qflags = reinterpret_cast<int *>(base + cand_bytes + cnt_bytes);
        int *gcnt = reinterpret_cast<int *>(base + cand_bytes + cnt_bytes + flag_bytes);
        float *gthr = reinterpret_cast<float *>(base + cand_bytes + cnt_bytes + 2 * flag_bytes);
        float *qcap = reinterpret_cast<float *>(base + cand_bytes + cnt_bytes + 3 * flag_bytes);
        float *qband = reinterpret_cast<float *>(base + cand_bytes + cnt_bytes + 4 * flag_bytes);
        unsigned long long *glist = reinterpret_cast<unsigned long long *>(base + cand_bytes + cnt_bytes + 5 * flag_bytes);
        // (kept 1 KB aligned: bulk copies need 16-byte aligned sources)
        unsigned char *qt = base + ((cand_bytes + cnt_bytes + 5 * flag_bytes + glist_bytes + 1023) & ~(size_t)1023);
        e = cudaMemsetAsync(qflags, 0, flag_bytes, st);
        if (e != cudaSuccess) return e;

        CUtensorMap tmX, tmQ;
        if (H) {
            memset(&tmX, 0, sizeof(tmX));   // the bf16 mode moves its operands with plain bulk copies: no tensor maps
            memset(&tmQ, 0, sizeof(tmQ));
            auto prep = s.metric ? prep_queries_kernel<true> : prep_queries_kernel<false>;
            const char *dbg = kDbg ? getenv("VROD_BATCHED_DEBUG") : nullptr;
            const bool nocand = dbg && (strstr(dbg, "nocand") || strstr(dbg, "noepi") || strstr(dbg, "nomma") || strstr(dbg, "ldonly"));
            const uint32_t slots = groups * BN;
            const float acc_eps_f = (float)((double)(ld_h / UMMA_K_H + 2) * ldexp(1.0, -23));
            prep<<<(slots + 7) / 8, 256, 0, st>>>(qw, bq, slots, s.ld, kd, T, s.maxnorm_bits, qt, gthr, qcap, nocand ? -1.f : 1.f,
                                                  band ? qband : nullptr, (float)eps_dot, acc_eps_f, s.mirror_stats);
            if (stats) stats->launches += 1;
        } else if (!make_map(&tmX, s.rows, s.n, s.ld, BM) || !make_map(&tmQ, qw, bq, s.ld, BN)) {
            return cudaErrorInvalidValue;
        }

        BatchedParams p{};
        p.sq_norm = s.sq_norm;
        p.inv_norm = s.inv_norm;
        p.n = s.n;
        p.b = bq;
        p.nslab = nslab;
        p.ksteps_last = ksteps_last;
        p.qgroups = groups;
        p.units = units;
        p.cand = cand;
        p.cnt_out = cnt;
        p.qflags = qflags;
        p.kprime = band ? CAP : kprime;   // band mode: a list may need more than kprime entries, nothing may be pruned away
        if (kDbg) {
            const char *dbg = getenv("VROD_BATCHED_DEBUG");
            p.debug_nocand = dbg && (strcmp(dbg, "nocand") == 0 || strstr(dbg, "noepi") || strstr(dbg, "nomma") || strstr(dbg, "ldonly"));
            p.debug_skip = dbg ? ((strstr(dbg, "noepi") ? 1 : 0) | (strstr(dbg, "nomma") ? 2 : 0) | (strstr(dbg, "ldonly") ? 4 : 0) |
                                  (strstr(dbg, "noflush") ? 8 : 0) | (strstr(dbg, "nopark") ? 16 : 0)) : 0;
            static long long *dbg_buf = nullptr;
            if (dbg && !dbg_buf) cudaMalloc(&dbg_buf, 1024 * 16 * sizeof(long long));
            p.dbg = dbg ? dbg_buf : nullptr;
            if (dbg) g_dbg_buf = dbg_buf;
        }
        const size_t smem_budget = 227 * 1024 - sizeof(BatchCtl);
        size_t resident, stage_bytes;
        if (H) {
            // bf16 mode.  The query group stays resident (T x 8 KB) whenever at least 3 stages of >= 2 K steps fit next
            // to it (dim + 16 <= ~400); a stage then holds S K-step blocks of the row tile -- the whole tile when it
            // fits 4 times (dim = 128: 4 stages x 36 KB next to 72 KB of queries).  Larger dims stream S = 4 K steps
            // of both operands per stage (48 KB).
            p.rows_t = reinterpret_cast<const unsigned char *>(s.rows_h);
            p.q_t = qt;
            p.T = T;
            uint32_t S = 0;
            const size_t qres = (size_t)T * KB_B;
            if (qres + 3 * 2 * KB_A <= smem_budget) {
                for (S = T; S >= 2; --S)
                    if ((smem_budget - qres) / ((size_t)S * KB_A) >= (S == 2 ? 3u : 4u)) break;
            }
            if (S >= 2) {
                p.stream_q = 0;
                resident = qres;
                stage_bytes = (size_t)S * KB_A;
            } else {
                S = T < 4 ? T : 4;
                p.stream_q = 1;
                resident = 0;
                stage_bytes = (size_t)S * (KB_A + KB_B);
            }
            p.S = S;
        } else {
            // tf32 mode, dim <= 128: the query group is resident (nslab x 32 KB) and only row slabs stream; larger dims
            // stream the query slab next to every row slab (48 KB stages, L2-bandwidth bound: DESIGN.md)
            p.stream_q = nslab > (uint32_t)MAX_SLABS ? 1 : 0;
            const size_t slab_b = (size_t)SLAB_B_BYTES;
            resident = p.stream_q ? 0 : (size_t)nslab * slab_b;
            stage_bytes = p.stream_q ? (SLAB_A_BYTES + slab_b) : SLAB_A_BYTES;
        }
        // (dynamic shared memory starts 1024-byte aligned -- the kernel has no static shared memory and traps
        // otherwise -- so no alignment slack is reserved: at dim 128 in tf32 mode that is what makes the 6th stage fit)
        size_t stages = (smem_budget - resident) / stage_bytes;
        {
            static const int st_env = getenv("VROD_BATCHED_STAGES") ? atoi(getenv("VROD_BATCHED_STAGES")) : 0;
            if (st_env >= 2 && (size_t)st_env < stages) stages = (size_t)st_env;
        }
        if (stages > 8) stages = 8;
        if (stages < 2) return cudaErrorInvalidConfiguration;
        p.stages = (uint32_t)stages;
        const size_t smem = resident + stages * stage_bytes + sizeof(BatchCtl);
        typedef void (*TileFn)(const CUtensorMap, const CUtensorMap, const BatchedParams);
        const TileFn tile_fn = H ? (s.metric ? batched_tile_kernel<true, true, false> : batched_tile_kernel<false, true, false>)
                                 : (s.metric ? batched_tile_kernel<true, false, false> : batched_tile_kernel<false, false, false>);
        // the start phase (one tile per CTA, everything passes) has its own epilogue
        const TileFn tile_fn_first = H ? (s.metric ? batched_tile_kernel<true, true, true> : batched_tile_kernel<false, true, true>)
                                       : (s.metric ? batched_tile_kernel<true, false, true> : batched_tile_kernel<false, false, true>);
        e = cudaFuncSetAttribute(tile_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tile_fn_first, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute cattr[1];
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cattr[0].id = cudaLaunchAttributeClusterDimension;
        cattr[0].val.clusterDim.x = 1;   // (the kernels address shared memory through the cluster window: a cluster of one)
        cattr[0].val.clusterDim.y = 1;
        cattr[0].val.clusterDim.z = 1;
        cfg.attrs = cattr;
        cfg.numAttrs = 1;

        FinishParams f{};
        f.rows4 = reinterpret_cast<const float4 *>(s.rows);
        f.q4 = reinterpret_cast<const float4 *>(qw);
        f.n = s.n;
        f.ld4 = s.ld / 4;
        f.b = bq;
        f.k = k;
        f.rank = s.rank;
        f.world = s.world ? s.world : 1;
        f.kprime = kprime;
        f.need = band ? (int)k : kprime;
        f.keep = keep;
        f.qband = band ? qband : nullptr;
        f.cap = kFinCap;
        f.water = f.cap - kScanThreads;
        f.qgroups = groups;
        f.units = units;
        f.cand = cand;
        f.cnt_in = cnt;
        f.qflags = qflags;
        f.maxnorm_bits = s.maxnorm_bits;
        f.glist = glist;
        f.gcnt = gcnt;
        f.gthr = gthr;
        f.status = status + (size_t)g0 * BN;
        f.out = out + (size_t)g0 * BN * k;
        f.out_ids = out_ids ? out_ids + (size_t)g0 * BN * k : nullptr;
        f.out_dist = out_dist ? out_dist + (size_t)g0 * BN * k : nullptr;
        f.eps_dot = eps_dot;
        f.mirror_stats = H ? s.mirror_stats : nullptr;
        f.qt = H ? qt : nullptr;
        f.ld_h = ld_h;
        f.kd = kd;
        f.T = T;
        f.qcap = qcap;
        f.acc_eps = H ? (double)(ld_h / UMMA_K_H + 2) * ldexp(1.0, -23) : 0.0;
        const size_t fsmem = kFinCtl + (size_t)f.cap * sizeof(unsigned long long) + ((size_t)kFinHist + (size_t)cpg_max + 2) * sizeof(int);
        auto fin_fn = s.metric ? batched_finish_kernel<true> : batched_finish_kernel<false>;

        // Phases over the row tiles: 1 tile per CTA first, then each phase 3.5x the rows seen so far (phase
        // boundaries are multiples of the pair size).  Between
        // phases the finish kernel merges all CTA lists of a query into its exact global k'-th threshold, so
        // the candidate rate of a phase is ~k'/rows_seen instead of ~k'/rows_seen_by_one_CTA.
        uint32_t t_begin = 0, t_end = cpg_max < ntiles ? cpg_max : ntiles;
        bool first = true;
        static const bool no_guess_env = getenv("VROD_BATCHED_NO_GUESS") != nullptr;
        const bool guess = guess_mode && !band && !no_guess_env;
        int prev_guess = 0;
        if (ev_start && g0 == 0) cudaEventRecord(ev_start, st);
        while (true) {
            const bool last = t_end >= ntiles;
            // The next phase's extent is fixed now: the finish kernel of this phase publishes the threshold it filters with.
            // 3.5x the rows seen so far per phase at the kprime-th key (measured best of 2.5 / 3.5 / 5 / 7 / 10 at configs[2]: new +
            // carried keys fit the 2048-key select); short scans (few tiles per unit) take bigger steps: there every
            // extra phase costs a tile launch and a finish kernel (~50 us) that the tiles cannot amortise.  Guessed
            // thresholds let through ~kprime * 2 keys per phase whatever its length, so the steps are 8x.
            uint32_t t_next = t_end;
            if (!last) {
                static const double growth_env = getenv("VROD_BATCHED_GROWTH") ? atof(getenv("VROD_BATCHED_GROWTH")) : 0.0;
                const double growth = growth_env > 1.0 ? growth_env : ((guess || (super_tiles / cpg_max) < 256) ? 8.0 : 3.5);
                unsigned long long nxt = (unsigned long long)((double)t_end * growth);
                if (nxt <= t_end) nxt = t_end + 1;
                t_next = nxt >= ntiles ? ntiles : (uint32_t)nxt;
                // a short rest is not worth a phase of its own
                if (guess && ntiles - t_next < t_next / 2) t_next = ntiles;
                // ... and a rest within 64x of the rows seen goes in ONE phase: its guess lets ~2.5 k' keys through, a phase
                // more costs a launch and a finish kernel (1M x 128, 256 queries: 191 -> 179 us; configs[2]: 4 phases, same time)
                static const double rest_env = getenv("VROD_BATCHED_REST") ? atof(getenv("VROD_BATCHED_REST")) : 64.0;
                if (guess && (double)ntiles <= rest_env * (double)t_end) t_next = ntiles;
            }
            p.tile_begin = t_begin;
            p.tile_end = t_end;
            p.thr_init = (first && !H) ? nullptr : gthr;   // bf16 mode starts from the finite caps of prep_queries_kernel
            e = cudaLaunchKernelEx(&cfg, first ? tile_fn_first : tile_fn, tmX, tmQ, p);
            if (e != cudaSuccess) return e;
            f.first_phase = first ? 1 : 0;
            f.final_phase = last ? 1 : 0;
            f.prev_guess = prev_guess;
            f.guess_rank = 0;
            if (guess && !last) {
                const double rows_seen = (double)t_end * BM, rows_next = (double)(t_next < ntiles ? (unsigned long long)t_next * BM : s.n);
                if (rows_seen >= 8.0 * kprime) f.guess_rank = guess_rank(kprime, rows_next / rows_seen);
            }
            prev_guess = f.guess_rank > 0 ? 1 : 0;
            fin_fn<<<bq, kScanThreads, fsmem, st>>>(f);
            e = cudaGetLastError();
            if (e != cudaSuccess) return e;
            if (stats) stats->launches += 2;
            if (kDbg && g_dbg_buf && getenv("VROD_BATCHED_DEBUG")) {
                static long long h[1024 * 16];
                cudaStreamSynchronize(st);
                cudaMemcpy(h, g_dbg_buf, sizeof(long long) * grid * 16, cudaMemcpyDeviceToHost);
                double a[13] = {0};
                for (uint32_t c = 0; c < grid; ++c)
                    for (int jx = 0; jx < 13; ++jx) a[jx] += (double)h[c * 16 + jx] / grid;
                fprintf(stderr, "[batched dbg] tiles [%u,%u) avg/CTA: warp2 slow entries %.0f, mma total %.0f, candidates %.0f | epi wait_tfull %.0f prune %.0f | mma wait_tempty %.0f wait_full %.0f "
                                "| epi total %.0f (warp2: park %.0f flush %.0f part barrier %.0f) | tiles/CTA %u\n",
                        t_begin, t_end, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[10], a[11], a[12], (t_end - t_begin + cpg_max - 1) / cpg_max);
            }
            if (last) break;
            first = false;
            t_begin = t_end;
            t_end = t_next;
        }
        if (ev_stop && g0 + groups >= qgroups) cudaEventRecord(ev_stop, st);
        if (stats) stats->tiles += (uint64_t)ntiles * groups;
    }
    return cudaSuccess;
}

}  // namespace vrod
