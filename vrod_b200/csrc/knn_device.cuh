// knn_device.cuh -- device helpers shared by the kNN scan kernels (sm_100a).
//
// Hot path of vRod's SEARCH command (call site: reference src/command/types.rs:114-119, empty body).
// Nothing here has a reference counterpart; the semantics are DESIGN.md "Search semantics".
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vrod {

constexpr int kScanThreads = 256;  // 8 warps per CTA
constexpr int kScanWarps = kScanThreads / 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr uint64_t kKeyMax = 0xFFFFFFFFFFFFFFFFull;

// Development instrumentation (cycle / globaltimer stamps, "skip the MMA / the epilogue" timing experiments) is
// compiled in only by `make DEBUG_KERNELS=1`: the production kernels carry none of it (every use is guarded by
// this constant, so the branches and the clock reads fold away).
#ifdef VROD_KERNEL_DEBUG
constexpr bool kDbg = true;
#else
constexpr bool kDbg = false;
#endif

// ---- order-preserving float <-> uint32 (so (value, row) packs into one comparable u64) --------
__device__ __forceinline__ uint32_t f2ord(float v) {
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(b);
}
__device__ __forceinline__ uint64_t make_key(float v, uint32_t row) {
    return ((uint64_t)f2ord(v) << 32) | row;
}

// ---- streaming 128-bit load: read-only path, no L1 allocation (each byte is used once) --------
__device__ __forceinline__ float4 ldg_stream(const float4 *p) {
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p));
    return r;
}

// ---- Philox4x32-10 (Salmon et al., SC'11), the library's own statement of the generator -------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
    uint32_t c2 = 0, c3 = 0;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ float word_to_unit(uint32_t w) {
    return (float)((int32_t)(w >> 8) - (1 << 23)) * 0x1p-23f;
}

// ---- multi-row butterfly reduction --------------------------------------------------------------
// Each lane holds N partial sums (one per row slot).  After the call every lane holds the full sum
// of ONE row slot (returned), replicated over the lanes that differ in the bits not used for slot
// selection; which slot a lane ends up with depends only on its lane id (callers compute it once).  N row slots cost N-1 + log2(LPR/N) shuffles instead of
// N*log2(LPR).  ASC = true walks lane offsets 1,2,4,.. (the canonical adjacent-pair order of the
// exact f64 sums); ASC = false walks LPR/2,..,1.
template <typename T, int N, int OFF, int LPR, bool ASC>
struct RowsReduce {
    static __device__ __forceinline__ T run(T *p, int lane) {
        constexpr bool live = ASC ? (OFF < LPR) : (OFF >= 1);
        if constexpr (!live) {
            static_assert(N == 1, "more row slots than lanes per row");
            return p[0];
        } else {
            constexpr int NEXT = ASC ? OFF * 2 : OFF / 2;
            if constexpr (N > 1) {
                constexpr int H = N / 2;
                const bool up = (lane & OFF) != 0;
#pragma unroll
                for (int j = 0; j < H; ++j) {
                    const T send = up ? p[j] : p[j + H];
                    const T keep = up ? p[j + H] : p[j];
                    p[j] = keep + __shfl_xor_sync(kFull, send, OFF);
                }
                return RowsReduce<T, H, NEXT, LPR, ASC>::run(p, lane);
            } else {
                p[0] = p[0] + __shfl_xor_sync(kFull, p[0], OFF);
                return RowsReduce<T, 1, NEXT, LPR, ASC>::run(p, lane);
            }
        }
    }
};

// ---- per-CTA candidate store ----------------------------------------------------------------------
// A CTA keeps the best `kprime` keys it has seen in shared memory.  Warps append keys that beat the
// current threshold; when the store passes its high-water mark a warp raises prune_req and every
// warp joins a block-wide sort at its next iteration boundary.
struct CandCtl {
    unsigned long long thrkey;  // kprime-th best key so far (kKeyMax until kprime keys are held)
    float thr_f;                // its value part, for the cheap per-row test
    int cnt;                    // keys in the store
    int prune_req;
    int overflow;               // store overran its capacity (cannot happen by construction; checked)
    int done_warps;
    int all_done;
    int is_last;
    double nq;                  // canonical sum q_j^2 (cosine rerank)
    float u_val;                // value part of the kprime-th approximate key (guard)
    int ncand;
    unsigned long long thrkey2; // last-CTA merge: the warps' offers for the bound T0
    float dq2;                  // batched finish: ||bf16(q) - q||^2 (guard of the bf16 operand mode)
};

__device__ __forceinline__ void cand_reset(CandCtl *ctl) {
    ctl->thrkey = kKeyMax;
    ctl->thr_f = __int_as_float(0x7f800000);
    ctl->cnt = 0;
    ctl->prune_req = 0;
    ctl->done_warps = 0;
    ctl->all_done = 0;
}

__device__ __forceinline__ void cand_append(CandCtl *ctl, unsigned long long *buf, unsigned long long key, int cap,
                                            int water) {
    if (key < *(volatile unsigned long long *)&ctl->thrkey) {
        const int pos = atomicAdd(&ctl->cnt, 1);
        if (pos < cap) buf[pos] = key;
        else ctl->overflow = 1;
        if (pos >= water) *(volatile int *)&ctl->prune_req = 1;
    }
}

// Block-wide bitonic sort of P (power of two, >= 64) keys in shared memory, ascending.  Caller has synced.
// Pair i of a stage touches elements inside one 64-key block whenever the partner distance j <= 32, and
// pairs [32w, 32w+32) (+ multiples of kScanThreads) belong to warp w, so those stages only need a warp
// barrier; a block barrier is needed only around the stages with j >= 64.
// P <= 64: one warp sorts in registers (two keys per lane, partners by shuffle; the j = 32 exchange is lane-local).  The
// shared-memory network below pays a load-compare-store-barrier round trip of ~250 cycles per stage however few keys it
// holds -- 2.8 us for the ~40 survivors of a merge, again for the k' reranked keys: most of a single query's serial tail.
__device__ __forceinline__ void warp_sort64_regs(unsigned long long &x0, unsigned long long &x1, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 32) {   // k == 64: lanes hold the pair (lane, lane + 32), ascending
                if (x0 > x1) {
                    const unsigned long long t = x0;
                    x0 = x1;
                    x1 = t;
                }
            } else {
                const unsigned long long o0 = __shfl_xor_sync(kFull, x0, j), o1 = __shfl_xor_sync(kFull, x1, j);
                const bool low = (lane & j) == 0;
                const bool asc0 = (lane & k) == 0, asc1 = ((lane + 32) & k) == 0;
                x0 = (low == asc0) ? (x0 < o0 ? x0 : o0) : (x0 < o0 ? o0 : x0);
                x1 = (low == asc1) ? (x1 < o1 ? x1 : o1) : (x1 < o1 ? o1 : x1);
            }
        }
    }
}
__device__ __forceinline__ void warp_sort64(unsigned long long *a, int P, int lane) {
    unsigned long long x0 = a[lane], x1 = P > 32 ? a[lane + 32] : kKeyMax;
    warp_sort64_regs(x0, x1, P, lane);
    a[lane] = x0;
    if (P > 32) a[lane + 32] = x1;
}

__device__ __forceinline__ void block_bitonic(unsigned long long *a, int P, int tid) {
    if (P <= 64) {
        if (tid < 32) warp_sort64(a, P, tid);
        __syncthreads();
        return;
    }
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < (P >> 1); i += kScanThreads) {
                const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1));
                const int hi = lo | j;
                const bool asc = (lo & k) == 0;
                const unsigned long long x = a[lo], y = a[hi];
                if ((x > y) == asc) { a[lo] = y; a[hi] = x; }
            }
            const int next_j = j > 1 ? (j >> 1) : k;   // first distance of the next merge level is k
            if (j >= 64 || next_j >= 64 || (j == 1 && k == P)) __syncthreads();
            else __syncwarp();
        }
    }
}

// n <= kScanThreads keys: every thread ranks its own key against all of them (pairs of keys per 128-bit
// broadcast load) and writes it to its rank -- a sort in two barriers.  The networks above cost a dependent
// load-compare-store (or shuffle-compare) chain per stage, ~2.5 us even for 40 keys; this is ~0.3 us for them.  Caller has
// synced; buf must be 16-byte aligned with one readable slot behind an odd n.
__device__ __forceinline__ void block_ranksort(unsigned long long *buf, int n, int tid) {
    const unsigned long long mine = tid < n ? buf[tid] : kKeyMax;
    int r0 = 0, r1 = 0;
    if (tid < n) {
        const ulonglong2 *b2 = reinterpret_cast<const ulonglong2 *>(buf);
        const int n2 = n >> 1;
        for (int j = 0; j < n2; ++j) {   // (equal keys -- there should be none -- are ordered by position, so every rank is taken once)
            const ulonglong2 a = b2[j];
            r0 += (a.x < mine || (a.x == mine && 2 * j < tid)) ? 1 : 0;
            r1 += (a.y < mine || (a.y == mine && 2 * j + 1 < tid)) ? 1 : 0;
        }
        if (n & 1) r0 += (buf[n - 1] < mine || (buf[n - 1] == mine && n - 1 < tid)) ? 1 : 0;
    }
    __syncthreads();
    if (tid < n) buf[r0 + r1] = mine;
    __syncthreads();
}

// Keep the best kprime keys (sorted ascending in buf[0..cnt)), refresh the threshold.  All threads of
// the CTA call it right after a __syncthreads().
__device__ __forceinline__ void block_prune(CandCtl *ctl, unsigned long long *buf, int kprime, int cap, int tid) {
    int cnt = ctl->cnt;
    if (cnt > cap) cnt = cap;
    if (cnt <= kScanThreads) {   // (uniform) the usual case everywhere but in the middle of a scan
        block_ranksort(buf, cnt, tid);
    } else {
        int P = 64;
        while (P < cnt) P <<= 1;
        for (int i = cnt + tid; i < P; i += kScanThreads) buf[i] = kKeyMax;
        __syncthreads();
        block_bitonic(buf, P, tid);
    }
    if (tid == 0) {
        const int nc = cnt < kprime ? cnt : kprime;
        ctl->cnt = nc;
        if (nc == kprime) {
            ctl->thrkey = buf[kprime - 1];
            ctl->thr_f = ord2f((uint32_t)(buf[kprime - 1] >> 32));
        } else {
            ctl->thrkey = kKeyMax;
            ctl->thr_f = __int_as_float(0x7f800000);
        }
        ctl->prune_req = 0;
        ctl->all_done = (ctl->done_warps == kScanWarps);
    }
    __syncthreads();
}

// Same contract as block_prune EXCEPT that the kept keys are left unsorted: keep the best kprime keys in
// buf[0..cnt), refresh the threshold (the exact kprime-th smallest key).  Radix select instead of a sort: every
// thread holds its <= MAXPER keys in registers; each pass histograms 8 key bits of the keys that still match
// the selected prefix (256 bins = one per thread, shared-memory atomics), warp 0 locates the bin that holds the
// kprime-th key, and the passes stop as soon as that bin holds one key or exactly the keys still needed
// (3-4 passes on real lists, 8 at most).  ~5x fewer instructions than the 1024-key bitonic sort that dominated
// batched_finish_kernel in round 1.  `hist` is 256 + 4 ints of shared memory.  All threads call it after a
// __syncthreads(); cnt <= cap <= MAXPER * kScanThreads.
// BAND mode (keep_max > 0): T is the kprime-th smallest key as before, but everything up to T + band survives (at most
// keep_max keys; more than that raises ctl->overflow) and the threshold becomes T + band: the caller keeps every
// candidate whose approximate key lies within the error band above the k-th best one instead of a fixed number of them.
__device__ __forceinline__ unsigned long long band_key(unsigned long long T, float band) {
    if (T == kKeyMax) return kKeyMax;
    return make_key(__fadd_ru(ord2f((uint32_t)(T >> 32)), band), 0xffffffffu);
}
template <int MAXPER>
__device__ __forceinline__ void block_select(CandCtl *ctl, unsigned long long *buf, int kprime, int cap, int tid, int *hist, float band = 0.f,
                                             int keep_max = 0) {
    int cnt = ctl->cnt;
    if (cnt > cap) cnt = cap;
    if (cnt <= kprime) {   // uniform: nothing to drop (short lists are cheap to sort)
        block_prune(ctl, buf, kprime, cap, tid);
        if (keep_max > 0) {
            if (tid == 0 && ctl->thrkey != kKeyMax) {
                ctl->thrkey = band_key(ctl->thrkey, band);
                ctl->thr_f = ord2f((uint32_t)(ctl->thrkey >> 32));
            }
            __syncthreads();
        }
        return;
    }
    unsigned long long key[MAXPER];
#pragma unroll
    for (int e = 0; e < MAXPER; ++e) {
        const int i = tid + e * kScanThreads;
        key[e] = i < cnt ? buf[i] : kKeyMax;
    }
    const int lane = tid & 31, warp = tid >> 5;
    unsigned long long prefix = 0, pmask = 0;
    int need = kprime;          // rank of the wanted key among the keys that match the prefix (1-based)
    int bucket = cnt;           // keys that match the prefix
    int shift = 56;
    while (true) {
        hist[tid] = 0;
        __syncthreads();
#pragma unroll
        for (int e = 0; e < MAXPER; ++e)
            if (tid + e * kScanThreads < cnt && (key[e] & pmask) == prefix) atomicAdd(&hist[(int)(key[e] >> shift) & 255], 1);
        __syncthreads();
        if (warp == 0) {
            int v[8], sum = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                v[e] = hist[lane * 8 + e];
                sum += v[e];
            }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            const unsigned m = __ballot_sync(kFull, incl >= need);
            if (lane == __ffs(m) - 1) {
                int before = incl - sum, e = 0;
#pragma unroll
                for (int x = 0; x < 7; ++x)
                    if (before + v[e] < need) {
                        before += v[e];
                        ++e;
                    }
                hist[256] = lane * 8 + e;      // selected bin
                hist[257] = need - before;     // rank inside it
                hist[258] = v[e];              // keys inside it
            }
        }
        __syncthreads();
        const int bin = hist[256];
        need = hist[257];
        bucket = hist[258];   // (rewritten only after the next pass's two barriers)
        prefix |= (unsigned long long)bin << shift;
        pmask |= 0xFFull << shift;
        if (bucket == 1 || bucket == need || shift == 0) break;
        shift -= 8;
    }
    // the threshold key T: if the whole bucket is taken (bucket == need, which covers bucket == 1) its largest key;
    // otherwise all 64 bits are fixed and every key of the bucket equals the prefix
    unsigned long long *tkey = reinterpret_cast<unsigned long long *>(hist + 260);   // 8-byte aligned (hist is)
    if (tid == 0) {
        *tkey = 0;
        ctl->cnt = 0;
    }
    __syncthreads();
    if (bucket == need) {
#pragma unroll
        for (int e = 0; e < MAXPER; ++e)
            if (tid + e * kScanThreads < cnt && (key[e] & pmask) == prefix) atomicMax(tkey, key[e]);
    } else if (tid == 0) {
        *tkey = prefix;   // shift == 0 with more equal keys than needed: any `need` of them do
    }
    __syncthreads();
    const unsigned long long T = *tkey;
    if (keep_max > 0) {
        // band mode: everything at or below T + band survives
        const unsigned long long Tb = band_key(T, band);
#pragma unroll
        for (int e = 0; e < MAXPER; ++e)
            if (tid + e * kScanThreads < cnt && key[e] <= Tb) {
                const int pos = atomicAdd(&ctl->cnt, 1);
                if (pos < keep_max) buf[pos] = key[e];
            }
        __syncthreads();
        if (tid == 0) {
            if (ctl->cnt > keep_max) {
                ctl->overflow = 1;   // the band holds more rows than the list: the query cannot be proven from here
                ctl->cnt = keep_max;
            }
            ctl->thrkey = Tb;
            ctl->thr_f = ord2f((uint32_t)(Tb >> 32));
            ctl->prune_req = 0;
        }
        __syncthreads();
        return;
    }
    // compaction: keys below T all survive; keys equal to T fill up to kprime (keys are unique in practice)
#pragma unroll
    for (int e = 0; e < MAXPER; ++e)
        if (tid + e * kScanThreads < cnt && key[e] < T) buf[atomicAdd(&ctl->cnt, 1)] = key[e];
    __syncthreads();
#pragma unroll
    for (int e = 0; e < MAXPER; ++e)
        if (tid + e * kScanThreads < cnt && key[e] == T) {
            const int pos = atomicAdd(&ctl->cnt, 1);
            if (pos < kprime) buf[pos] = key[e];
        }
    __syncthreads();
    if (tid == 0) {
        ctl->cnt = kprime;
        ctl->thrkey = T;
        ctl->thr_f = ord2f((uint32_t)(T >> 32));
        ctl->prune_req = 0;
    }
    __syncthreads();
}

// ---- canonical f64 sums (bit-identical to oracle/knn_oracle.c by construction of the order) --------
// partial[j mod 128] accumulates element j with fma; lane l owns partials 4l..4l+3; the adjacent-pair
// tree is (p0+p1)+(p2+p3) inside the lane, then lanes xor 1,2,4,8,16.
template <bool COS>
__device__ __forceinline__ void canon_accum(const float4 x, const float4 q, double (&p)[4]) {
    if constexpr (COS) {
        p[0] = fma((double)x.x, (double)q.x, p[0]);
        p[1] = fma((double)x.y, (double)q.y, p[1]);
        p[2] = fma((double)x.z, (double)q.z, p[2]);
        p[3] = fma((double)x.w, (double)q.w, p[3]);
    } else {
        const double a = __dsub_rn((double)x.x, (double)q.x), b = __dsub_rn((double)x.y, (double)q.y);
        const double c = __dsub_rn((double)x.z, (double)q.z), d = __dsub_rn((double)x.w, (double)q.w);
        p[0] = fma(a, a, p[0]);
        p[1] = fma(b, b, p[1]);
        p[2] = fma(c, c, p[2]);
        p[3] = fma(d, d, p[3]);
    }
}
__device__ __forceinline__ double canon_lane_fold(const double (&p)[4]) {
    return __dadd_rn(__dadd_rn(p[0], p[1]), __dadd_rn(p[2], p[3]));
}
__device__ __forceinline__ double canon_warp_tree(double v) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) v = __dadd_rn(v, __shfl_xor_sync(kFull, v, off));
    return v;
}

// One warp computes the canonical sum over a row of `ld4` float4 chunks.  MODE 0: SUM (x-q)^2,
// MODE 1: SUM x*q, MODE 2: SUM x*x.
template <int MODE>
__device__ __forceinline__ double canon_row_sum(const float4 *x, const float4 *q, int ld4, int lane) {
    double p[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = lane; c < ld4; c += 32) {
        const float4 xv = __ldg(x + c);
        if constexpr (MODE == 0) canon_accum<false>(xv, __ldg(q + c), p);
        else if constexpr (MODE == 1) canon_accum<true>(xv, __ldg(q + c), p);
        else canon_accum<true>(xv, xv, p);
    }
    return canon_warp_tree(canon_lane_fold(p));
}

// NR rows at once (independent loads and fma chains in flight together); MODE as above.
template <int MODE, int NR>
__device__ __forceinline__ void canon_rows_sum(const float4 *const (&x)[NR], const float4 *q, int ld4, int lane, double (&out)[NR]) {
    double p[NR][4];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) p[r][e] = 0.0;
    for (int c = lane; c < ld4; c += 32) {
        float4 xv[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) xv[r] = __ldg(x[r] + c);
        float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (MODE != 2) qv = __ldg(q + c);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            if constexpr (MODE == 0) canon_accum<false>(xv[r], qv, p[r]);
            else if constexpr (MODE == 1) canon_accum<true>(xv[r], qv, p[r]);
            else canon_accum<true>(xv[r], xv[r], p[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) out[r] = canon_warp_tree(canon_lane_fold(p[r]));
}

__device__ __forceinline__ float canon_l2_dist(double sq) { return __double2float_rn(__dsqrt_rn(sq)) + 0.0f; }
__device__ __forceinline__ float canon_cos_dist(double dot, double nx, double nq) {
    if (nx == 0.0 || nq == 0.0) return 1.0f;
    const double den = __dmul_rn(__dsqrt_rn(nx), __dsqrt_rn(nq));
    return __double2float_rn(__dsub_rn(1.0, __ddiv_rn(dot, den))) + 0.0f;
}

// Exact re-evaluation of the candidates buf[0..ncand) (approximate keys whose low word is the local row) in the
// canonical f64 order; each entry is replaced by its exact (f32 dist, row) key.  One warp takes NR candidates
// at a time so that their row loads are in flight together.
template <bool COS>
__device__ __forceinline__ void rerank_candidates(unsigned long long *buf, int ncand, const float4 *rows4, const float4 *q4,
                                                  int ld4, double nq, int warp, int lane) {
    constexpr int NR = COS ? 2 : 4;
    for (int c0 = warp * NR; c0 < ncand; c0 += kScanWarps * NR) {
        const float4 *x[NR];
        uint32_t row[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int c = c0 + r < ncand ? c0 + r : c0;
            row[r] = (uint32_t)buf[c];
            x[r] = rows4 + (size_t)row[r] * ld4;
        }
        float dist[NR];
        if constexpr (COS) {
            double nx[NR], dot[NR];
            canon_rows_sum<2, NR>(x, q4, ld4, lane, nx);
            canon_rows_sum<1, NR>(x, q4, ld4, lane, dot);
#pragma unroll
            for (int r = 0; r < NR; ++r) dist[r] = canon_cos_dist(dot[r], nx[r], nq);
        } else {
            double sq[NR];
            canon_rows_sum<0, NR>(x, q4, ld4, lane, sq);
#pragma unroll
            for (int r = 0; r < NR; ++r) dist[r] = canon_l2_dist(sq[r]);
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < NR; ++r)
                if (c0 + r < ncand) buf[c0 + r] = make_key(dist[r], row[r]);
        }
    }
}

}  // namespace vrod
