// knn_scan.cu -- single-query exact top-k scan kernels for sm_100a.
//
// The work vRod's SearchCommand::execute would do (reference src/command/types.rs:114-119, empty
// body) over the rows of a Database collection (reference src/database/mod.rs:6-10, TODO).
//
//   fast_scan_kernel   HBM-bound pass over the shard's f32 rows: per-row f32 distance surrogate
//                      (L2: SUM (x-q)^2, cosine: -dot * 1/||x|| with the precomputed inverse norm),
//                      multi-row butterfly reduction, per-CTA candidate store (threshold filter +
//                      shared-memory bitonic prune), last-CTA merge of all CTA lists, exact f64
//                      rerank of the kprime survivors in the canonical order and a guard that
//                      PROVES no dropped row can belong to the top k (else status = 1).
//   exact_scan_kernel  the same scan with every row evaluated in canonical f64; answers the
//                      queries whose guard failed (and path = 2).
//   row_norms / fill_synthetic / pad_queries / merge_hits  small helpers around them.
#include "knn_device.cuh"
#include "knn_scan.cuh"

#include <math.h>

namespace vrod {

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// -------------------------------------------------------------------------------------------------
struct ScanParams {
    const float4 *rows4;
    const float *inv_norm;
    const float4 *q4;
    uint32_t n, ld4;
    uint32_t rank, world;      // shard coordinates: global id of a local row by row_id()
    uint32_t k;
    int kprime, cap, water;
    uint32_t iters;
    unsigned long long *blk_cand;
    unsigned int *ticket;
    int *status;
    unsigned long long *counters;
    const int *only_if;
    unsigned long long *dbg;   // VROD_SCAN_DEBUG: globaltimer stamps [0] first CTA start, [1] last scan-loop end,
                               // [2] last-CTA merge start, [3] merge end, [4] rerank end, [5] kernel end
    Hit *out;
    unsigned long long *out_ids;  // optional: final [k] ids / distances of this query (single-GPU contexts
    float *out_dist;              // skip the merge kernel; sharded contexts with the fused exchange: the GLOBAL answer)
    double eps;
    XchgArgs x;                   // x.windows != nullptr: the last CTA also exchanges and merges the ranks' lists (fused)
};

constexpr int kCtlBytes = 128;
constexpr int kMaxListsPerThread = 5;   // gridDim.x <= 8 CTAs x 148 SMs = 1184 lists over 256 threads
static_assert(sizeof(CandCtl) <= kCtlBytes, "CandCtl must fit its slot");

// Finish protocol of a scanning CTA + cross-CTA merge by the last CTA to finish.
// Returns true in the last CTA, with the grid's best <= kprime keys sorted in buf[0..ctl->cnt).
__device__ __forceinline__ bool finish_and_merge(CandCtl *ctl, unsigned long long *buf, const ScanParams &p, int tid) {
    const int lane = tid & 31;
    if (lane == 0) atomicAdd(&ctl->done_warps, 1);
    // Warps that are done wait here; a straggler joins either to prune or because it is done too.
    while (true) {
        __syncthreads();
        block_prune(ctl, buf, p.kprime, p.cap, tid);
        if (ctl->all_done) break;
    }
    const int cnt = ctl->cnt;
    unsigned long long *mine = p.blk_cand + (size_t)blockIdx.x * p.kprime;
    for (int i = tid; i < p.kprime; i += kScanThreads) __stcg(mine + i, i < cnt ? buf[i] : kKeyMax);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(p.ticket, 1u);
        ctl->is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!ctl->is_last) return false;
    __threadfence();
    if (kDbg && p.dbg && tid == 0) p.dbg[2] = gtimer();

    // ---- last CTA: merge gridDim.x sorted lists with (almost always) ONE round of loads ----
    // Every thread fetches the first kPre keys of its lists, all loads in flight together.  The kprime-th smallest HEAD
    // T0 bounds the global kprime-th key from above (the kprime smallest heads are kprime distinct keys <= T0), so the
    // answer is made of the keys <= T0 -- and a list almost never holds more than kPre of those (the global top kprime
    // spread over hundreds of lists): the prefetched keys are filtered straight from registers, a list whose kPre-th key
    // is still below T0 is read on in a second round.  T0 is found by counting (each head ranked against all heads in
    // shared memory), not by sorting.  Every step here is a dependent round trip of a single CTA while the rest of the
    // GPU waits: round 1 of the build walked the lists by depth -- one key per active list, a barrier and often a sort
    // per level, 3-6 levels, 10-15 us on a scan that streams 1M x 128 in 80 us.
    if (tid == 0) {
        const int ov = ctl->overflow;
        cand_reset(ctl);
        ctl->overflow = ov;
        ctl->thrkey2 = 0;
        ctl->ncand = 0;
        *p.ticket = 0;  // self-reset for the next launch
    }
    const int nlists = gridDim.x;
    constexpr int kPre = 12;
    unsigned long long pre[2][kPre];                 // lists tid and tid + kScanThreads (the usual grid has 296)
    unsigned long long head[kMaxListsPerThread];
#pragma unroll
    for (int li = 0; li < 2; ++li) {
        const int b = tid + li * kScanThreads;
        const unsigned long long *lst = p.blk_cand + (size_t)b * p.kprime;
#pragma unroll
        for (int u = 0; u < kPre; ++u) pre[li][u] = (b < nlists && u < p.kprime) ? __ldcg(lst + u) : kKeyMax;
    }
#pragma unroll
    for (int li = 2; li < kMaxListsPerThread; ++li) {
        const int b = tid + li * kScanThreads;
        head[li] = b < nlists ? __ldcg(p.blk_cand + (size_t)b * p.kprime) : kKeyMax;
    }
    head[0] = pre[0][0];
    head[1] = pre[1][0];
    // the heads, staged in the (still empty) key buffer behind the slots the survivors will need first
    unsigned long long *hs = buf + ((p.cap - nlists - 1) & ~1);   // 16-byte aligned, one pad slot; cap >= kprime + max_grid + 64
#pragma unroll
    for (int li = 0; li < kMaxListsPerThread; ++li) {
        const int b = tid + li * kScanThreads;
        if (b <= nlists) hs[b] = head[li];   // (b == nlists: the pad, kKeyMax)
    }
    __syncthreads();
    // T0.  Any key with at least kprime heads at or below it will do; the tighter, the fewer survivors.  Up to 512 lists:
    // warp w sorts the heads w, w + 8, w + 16, ... (at most 64, in registers) and offers its r-th smallest, r =
    // ceil(kprime / 8); the largest offer has 8 r >= kprime heads at or below it.  It sits a little above the exact
    // kprime-th smallest head (~60 survivors instead of ~36 for kprime = 32, all sorted by rank-and-scatter anyway) and
    // costs a 64-key register sort -- ranking every head against all 296 of them in shared memory cost 4 us of bank
    // wavefronts.
    const int rr = (p.kprime + kScanWarps - 1) / kScanWarps;
    const int share = nlists / kScanWarps;               // heads every warp has at least
    if (nlists <= 64 * kScanWarps && share >= rr && rr <= 32) {
        const int lane = tid & 31, warp = tid >> 5;
        const int i0 = warp + kScanWarps * lane, i1 = warp + kScanWarps * (lane + 32);
        unsigned long long x0 = i0 < nlists ? hs[i0] : kKeyMax, x1 = i1 < nlists ? hs[i1] : kKeyMax;
        warp_sort64_regs(x0, x1, 64, lane);
        const unsigned long long offer = __shfl_sync(kFull, x0, rr - 1);
        if (lane == 0) {
            if (offer == kKeyMax) ctl->ncand = 1;        // a warp with fewer than r real heads: no bound this way
            else atomicMax(&ctl->thrkey2, offer);
        }
        __syncthreads();
        if (tid == 0) ctl->thrkey = ctl->ncand ? kKeyMax : ctl->thrkey2;
    } else if (nlists >= p.kprime) {
        const ulonglong2 *h2 = reinterpret_cast<const ulonglong2 *>(hs);
        const int n2 = (nlists + 1) >> 1;
        // list tid: the thread ranks its own head against all heads (two keys per 128-bit broadcast load, eight in flight)
        {
            const unsigned long long mine = head[0];
            int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
            int j = 0;
            for (; j + 4 <= n2; j += 4) {
                const ulonglong2 a = h2[j], c = h2[j + 1], d = h2[j + 2], e = h2[j + 3];
                r0 += (a.x < mine ? 1 : 0) + (d.x < mine ? 1 : 0);
                r1 += (a.y < mine ? 1 : 0) + (d.y < mine ? 1 : 0);
                r2 += (c.x < mine ? 1 : 0) + (e.x < mine ? 1 : 0);
                r3 += (c.y < mine ? 1 : 0) + (e.y < mine ? 1 : 0);
            }
            for (; j < n2; ++j) {
                const ulonglong2 a = h2[j];
                r0 += a.x < mine ? 1 : 0;
                r1 += a.y < mine ? 1 : 0;
            }
            if (mine != kKeyMax && r0 + r1 + r2 + r3 == p.kprime - 1) ctl->thrkey = mine;
        }
        // lists beyond the first kScanThreads (40 of the usual 296): warp w ranks heads kScanThreads + w, + w + 8, ... with
        // its lanes across the heads array, so that no warp walks the array a second time on its own
        const int lane = tid & 31, warp = tid >> 5;
        for (int b = kScanThreads + warp; b < nlists; b += kScanWarps) {
            const unsigned long long mine = hs[b];
            int r = 0;
            for (int j = lane; j < n2; j += 32) {
                const ulonglong2 a = h2[j];
                r += (a.x < mine ? 1 : 0) + (a.y < mine ? 1 : 0);
            }
            r = __reduce_add_sync(kFull, r);
            if (lane == 0 && mine != kKeyMax && r == p.kprime - 1) ctl->thrkey = mine;
        }
    }
    __syncthreads();
    const unsigned long long t0key = ctl->thrkey;    // kKeyMax: fewer than kprime non-empty lists
    const int ov_before = ctl->overflow;             // (this CTA's own scan; the gather below may raise the flag for its own reasons)
    if (kDbg && p.dbg && tid == 0) p.dbg[6] = gtimer();
    const int room = p.cap - nlists - 2;
    auto keep = [&](unsigned long long key) {
        const int pos = atomicAdd(&ctl->cnt, 1);
        if (pos < room) buf[pos] = key;
        else ctl->overflow = 1;
    };
    // every key <= T0: the qualifying heads and what follows them in their lists
#pragma unroll
    for (int li = 0; li < kMaxListsPerThread; ++li) {
        if (head[li] == kKeyMax || head[li] > t0key) continue;
        keep(head[li]);
        if (t0key == kKeyMax) continue;              // (no bound: the depth walk below reads the lists)
        int from = 1;                                // first key of the list not looked at yet
        if (li < 2) {
            bool more = true;
#pragma unroll
            for (int u = 1; u < kPre; ++u) {
                if (more && pre[li][u] < t0key) keep(pre[li][u]);
                else more = false;
            }
            if (!more) continue;
            from = kPre;
        }
        const unsigned long long *lst = p.blk_cand + (size_t)(tid + li * kScanThreads) * p.kprime;
        for (int j0 = from; j0 < p.kprime; j0 += 16) {   // lists are sorted: stop at the first key >= T0
            unsigned long long key[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) key[u] = j0 + u < p.kprime ? __ldcg(lst + j0 + u) : kKeyMax;
            bool more = true;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (more && key[u] < t0key) keep(key[u]);
                else more = false;
            }
            if (!more) break;
        }
    }
    __syncthreads();
    if (kDbg && p.dbg && tid == 0) p.dbg[7] = gtimer();
    // The depth walk: the lists are read one level at a time and pruned on the way, so that the buffer never has to hold
    // more than a level.  It takes over (a) when there is no finite T0 -- fewer non-empty lists than kprime: large k on a
    // small grid, tiny collections -- and (b) when more keys lie at or below T0 than the buffer holds: T0 is only as
    // tight as the heads are spread, and thousands of rows at exactly the same distance (sparse rows under the cosine
    // metric: half the collection at distance 1, ordered by id alone) put thousands of keys below any head-based bound.
    bool walk = t0key == kKeyMax;
    if (!walk && ctl->cnt > room) {   // (uniform)
        walk = true;
        __syncthreads();
        if (tid == 0) {
            cand_reset(ctl);
            ctl->overflow = ov_before;   // (what the gather could not hold raised the flag; nothing is lost, it starts over)
        }
        __syncthreads();
#pragma unroll
        for (int li = 0; li < kMaxListsPerThread; ++li)
            if (head[li] != kKeyMax) buf[atomicAdd(&ctl->cnt, 1)] = head[li];   // nlists <= cap
        __syncthreads();
        block_prune(ctl, buf, p.kprime, p.cap, tid);
    }
    if (walk) {
        const int water = p.cap - nlists;
        unsigned int active = 0;
#pragma unroll
        for (int li = 0; li < kMaxListsPerThread; ++li)
            if (head[li] != kKeyMax) active |= 1u << li;
        for (int depth = 1; depth < p.kprime; ++depth) {
            int any = 0;
#pragma unroll
            for (int li = 0; li < kMaxListsPerThread; ++li) {
                if (!(active & (1u << li))) continue;
                const unsigned long long key = __ldcg(p.blk_cand + (size_t)(tid + li * kScanThreads) * p.kprime + depth);
                if (key < *(volatile unsigned long long *)&ctl->thrkey) {
                    const int pos = atomicAdd(&ctl->cnt, 1);
                    if (pos < p.cap) buf[pos] = key;
                    else ctl->overflow = 1;
                    any = 1;
                } else {
                    active &= ~(1u << li);
                }
            }
            if (!__syncthreads_or(any)) break;
            if (ctl->cnt > water || ctl->cnt > p.kprime * 2) block_prune(ctl, buf, p.kprime, p.cap, tid);   // uniform: cnt is stable after the barrier
        }
    }
    __syncthreads();
    block_prune(ctl, buf, p.kprime, p.cap, tid);
    return true;
}

// Write the final k hits from sorted exact keys in buf[0..ncand).
__device__ __forceinline__ int exchange_device(const XchgArgs &x, const Hit *local, uint32_t k, unsigned long long *buf,
                                               unsigned long long *out_ids, float *out_dist, int tid);

// `flag` = this rank's guard flag for the query; it travels to the peers in hit[0].pad.  With the fused exchange the final
// arrays are written by exchange_device (the global answer), not here.
__device__ __forceinline__ void write_hits(const unsigned long long *buf, int ncand, const ScanParams &p, int tid, int flag) {
    for (int i = tid; i < (int)p.k; i += kScanThreads) {
        Hit h;
        if (i < ncand) {
            h.id = row_id((uint32_t)buf[i], p.rank, p.world);
            h.dist = ord2f((uint32_t)(buf[i] >> 32));
        } else {
            h.id = kKeyMax;
            h.dist = __int_as_float(0x7f800000);
        }
        h.pad = i == 0 ? (uint32_t)flag : 0u;
        p.out[i] = h;
        if (p.out_ids && !p.x.windows) {
            p.out_ids[i] = h.id;
            p.out_dist[i] = h.dist;
        }
    }
}

// Tail of a scan's last CTA: publish the k hits and the guard flag; in a sharded context with the fused exchange, meet the
// peers and leave the GLOBAL answer and the global flag instead.
__device__ __forceinline__ void publish(const unsigned long long *keys, int ncand, const ScanParams &p, CandCtl *ctl, unsigned long long *buf,
                                        int bad, int tid) {
    write_hits(keys, ncand, p, tid, bad);
    if (p.x.windows) {
        __threadfence();
        __syncthreads();
        bad = exchange_device(p.x, p.out, p.k, buf, p.out_ids, p.out_dist, tid);
    }
    if (tid == 0 && p.status) {
        *p.status = bad ? 1 : 0;
        if (bad && p.counters) atomicAdd(p.counters, 1ull);
    }
    (void)ctl;
}

// -------------------------------------------------------------------------------------------------
// Fast scan.  LPR lanes cooperate on one row (32/LPR rows per 128-bit load instruction), CH float4
// chunks per lane per row (CH == 0: run-time chunk loop), RB row slots in flight per iteration.
// -------------------------------------------------------------------------------------------------
template <bool COS>
__device__ __forceinline__ float chunk_acc(const float4 x, const float4 q, float acc) {
    if constexpr (COS) {
        acc = fmaf(x.x, q.x, acc);
        acc = fmaf(x.y, q.y, acc);
        acc = fmaf(x.z, q.z, acc);
        acc = fmaf(x.w, q.w, acc);
    } else {
        const float a = x.x - q.x, b = x.y - q.y, c = x.z - q.z, d = x.w - q.w;
        acc = fmaf(a, a, acc);
        acc = fmaf(b, b, acc);
        acc = fmaf(c, c, acc);
        acc = fmaf(d, d, acc);
    }
    return acc;
}

template <int LPR, int CH, int RB, bool COS>
__global__ void __launch_bounds__(kScanThreads, 2) fast_scan_kernel(const ScanParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    CandCtl *ctl = reinterpret_cast<CandCtl *>(smem);
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(smem + kCtlBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int RPL = 32 / LPR;    // rows per load instruction
    constexpr int ROWS = RB * RPL;   // rows per warp iteration
    constexpr int REP = LPR / RB;    // lanes holding the same row total after the reduction
    constexpr int CHR = CH > 0 ? CH : 1;
    const int sub = lane / LPR, ls = lane % LPR;
    const int rsel = ls / REP;
    const bool owner = (ls % REP) == 0;

    if (tid == 0) {
        cand_reset(ctl);
        ctl->overflow = 0;
    }
    float4 qv[CHR];
    if constexpr (CH > 0) {
#pragma unroll
        for (int c = 0; c < CH; ++c) qv[c] = __ldg(p.q4 + ls + LPR * c);
    }
    if (kDbg && p.dbg && tid == 0) atomicMin(p.dbg + 0, gtimer());
    __syncthreads();

    const int water = p.water;
    const uint32_t total_warps = gridDim.x * kScanWarps;
    const uint32_t gw = blockIdx.x * kScanWarps + warp;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (uint32_t it = 0; it < p.iters; ++it) {
        if (*(volatile int *)&ctl->prune_req) {
            __syncthreads();
            block_prune(ctl, buf, p.kprime, p.cap, tid);
        }
        const unsigned long long row0 = ((unsigned long long)it * total_warps + gw) * ROWS;
        const unsigned long long myrow = row0 + (unsigned)(rsel * RPL + sub);
        const bool myok = myrow < p.n;
        float part[RB];
        if constexpr (CH > 0) {
            float4 v[RB][CHR];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const unsigned long long row = row0 + (unsigned)(r * RPL + sub);
                const bool ok = row < p.n;
                const float4 *src = p.rows4 + row * p.ld4 + ls;
#pragma unroll
                for (int c = 0; c < CH; ++c) v[r][c] = ok ? ldg_stream(src + LPR * c) : (COS ? zero4 : qv[c]);
            }
            float inv = 1.0f;
            if (COS && myok) inv = __ldg(p.inv_norm + myrow);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                float acc = 0.f;
#pragma unroll
                for (int c = 0; c < CH; ++c) acc = chunk_acc<COS>(v[r][c], qv[c], acc);
                part[r] = acc;
            }
            float tot = RowsReduce<float, RB, LPR / 2, LPR, false>::run(part, lane);
            if constexpr (COS) tot = -tot * inv;
            const float thr = *(volatile float *)&ctl->thr_f;
            if (myok && owner && !(tot > thr)) cand_append(ctl, buf, make_key(tot, (uint32_t)myrow), p.cap, water);
        } else {
            // generic dimension: run-time chunk loop, q from L1
            float inv = 1.0f;
            if (COS && myok) inv = __ldg(p.inv_norm + myrow);
#pragma unroll
            for (int r = 0; r < RB; ++r) part[r] = 0.f;
            for (uint32_t c = ls; c < p.ld4; c += 4 * LPR) {
                float4 v[RB][4];
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t cc = c + u * LPR;
                    q[u] = cc < p.ld4 ? __ldg(p.q4 + cc) : zero4;
                }
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    const unsigned long long row = row0 + (unsigned)(r * RPL + sub);
                    const bool ok = row < p.n;
                    const float4 *src = p.rows4 + row * p.ld4;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t cc = c + u * LPR;
                        v[r][u] = (ok && cc < p.ld4) ? ldg_stream(src + cc) : (COS ? zero4 : q[u]);
                    }
                }
#pragma unroll
                for (int r = 0; r < RB; ++r)
#pragma unroll
                    for (int u = 0; u < 4; ++u) part[r] = chunk_acc<COS>(v[r][u], q[u], part[r]);
            }
            float tot = RowsReduce<float, RB, LPR / 2, LPR, false>::run(part, lane);
            if constexpr (COS) tot = -tot * inv;
            const float thr = *(volatile float *)&ctl->thr_f;
            if (myok && owner && !(tot > thr)) cand_append(ctl, buf, make_key(tot, (uint32_t)myrow), p.cap, water);
        }
    }

    if (kDbg && p.dbg && tid == 0) atomicMax(p.dbg + 1, gtimer());
    if (!finish_and_merge(ctl, buf, p, tid)) return;
    if (kDbg && p.dbg && tid == 0) p.dbg[3] = gtimer();

    // ---- last CTA: exact rerank of the survivors in canonical f64, guard, output ----
    const int ncand = ctl->cnt;
    if (tid == 0) {
        ctl->ncand = ncand;
        ctl->u_val = ncand > 0 ? ord2f((uint32_t)(buf[ncand - 1] >> 32)) : 0.f;
    }
    if (COS && warp == 0) {
        const double nq = canon_row_sum<2>(p.q4, p.q4, (int)p.ld4, lane);
        if (lane == 0) ctl->nq = nq;
    }
    __syncthreads();
    rerank_candidates<COS>(buf, ncand, p.rows4, p.q4, (int)p.ld4, ctl->nq, warp, lane);
    __syncthreads();
    if (kDbg && p.dbg && tid == 0) p.dbg[4] = gtimer();
    if (ncand <= kScanThreads) {
        block_ranksort(buf, ncand, tid);
    } else {
        int P = 32;
        while (P < ncand) P <<= 1;
        for (int i = ncand + tid; i < P; i += kScanThreads) buf[i] = kKeyMax;
        __syncthreads();
        block_bitonic(buf, P, tid);
    }
    if (tid == 0) {
        int bad = ctl->overflow;
        if (p.n > (uint32_t)ncand) {
            // rows were dropped: every dropped row has surrogate >= u_val.  Bound its exact distance
            // from below and require it to be strictly above the k-th exact distance.
            const int kk = (int)p.k < ncand ? (int)p.k : ncand;
            const float T = ord2f((uint32_t)(buf[kk - 1] >> 32));
            double u = (double)ctl->u_val;
            if (!(u == u)) bad = 1;  // NaN surrogate
            if (u > 3.4028234663852886e38) u = 3.4028234663852886e38;
            float lb;
            if constexpr (COS) {
                const double nqs = __dsqrt_rn(ctl->nq);
                if (nqs > 0.0) {
                    const double d = 1.0 + u / nqs - p.eps;
                    lb = __double2float_rd(d - 4.0e-16 * (1.0 + fabs(u / nqs)));
                } else {
                    lb = -1.0f;  // zero query: every distance is 1, the scan cannot rank; rescan exactly
                }
            } else {
                double s = u * (1.0 - p.eps) - (double)p.ld4 * 4.0 * 1.0e-44;
                if (s < 0.0) s = 0.0;
                lb = __double2float_rd(__dsqrt_rd(s));
            }
            if (!(lb > T)) bad = 1;
            if (kk < (int)p.k) bad = 1;
        }
        ctl->is_last = bad;   // (the slot is free again: this IS the last CTA)
    }
    __syncthreads();
    publish(buf, ncand, p, ctl, buf, ctl->is_last, tid);
    if (kDbg && p.dbg && tid == 0) p.dbg[5] = gtimer();
}

// -------------------------------------------------------------------------------------------------
// Exact scan: every row in canonical f64 (warp per row, RB rows in flight).
// -------------------------------------------------------------------------------------------------
template <int RB, bool COS>
__global__ void __launch_bounds__(kScanThreads, 2) exact_scan_kernel(const ScanParams p) {
    if (p.only_if != nullptr && *p.only_if == 0) return;
    extern __shared__ __align__(16) unsigned char smem[];
    CandCtl *ctl = reinterpret_cast<CandCtl *>(smem);
    unsigned long long *buf = reinterpret_cast<unsigned long long *>(smem + kCtlBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        cand_reset(ctl);
        ctl->overflow = 0;
    }
    if (COS && warp == 0) {
        const double nq = canon_row_sum<2>(p.q4, p.q4, (int)p.ld4, lane);
        if (lane == 0) ctl->nq = nq;
    }
    __syncthreads();
    const double nq = COS ? ctl->nq : 0.0;

    // slot owned after the ascending butterfly: bit-reversed low lane bits
    int rsel = 0;
    {
        int n = RB, off = 1;
        while (n > 1) {
            if (lane & off) rsel += n / 2;
            n >>= 1;
            off <<= 1;
        }
    }
    constexpr int REPMASK = ~(RB - 1) & 31;  // lanes with these bits clear own a distinct slot copy
    const bool owner = (lane & REPMASK) == 0;

    const int water = p.water;
    const uint32_t total_warps = gridDim.x * kScanWarps;
    const uint32_t gw = blockIdx.x * kScanWarps + warp;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (uint32_t it = 0; it < p.iters; ++it) {
        if (*(volatile int *)&ctl->prune_req) {
            __syncthreads();
            block_prune(ctl, buf, p.kprime, p.cap, tid);
        }
        const unsigned long long row0 = ((unsigned long long)it * total_warps + gw) * RB;
        double acc[RB][4], accx[COS ? RB : 1][4];
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                acc[r][e] = 0.0;
                if constexpr (COS) accx[r][e] = 0.0;
            }
        for (uint32_t c = lane; c < p.ld4; c += 32) {
            const float4 q = __ldg(p.q4 + c);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const unsigned long long row = row0 + r;
                const float4 x = row < p.n ? ldg_stream(p.rows4 + row * p.ld4 + c) : (COS ? zero4 : q);
                canon_accum<COS>(x, q, acc[r]);
                if constexpr (COS) canon_accum<true>(x, x, accx[r]);
            }
        }
        double part[RB], partx[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            part[r] = canon_lane_fold(acc[r]);
            if constexpr (COS) partx[r] = canon_lane_fold(accx[r]);
        }
        const double tot = RowsReduce<double, RB, 1, 32, true>::run(part, lane);
        float dist;
        if constexpr (COS) {
            const double nx = RowsReduce<double, RB, 1, 32, true>::run(partx, lane);
            dist = canon_cos_dist(tot, nx, nq);
        } else {
            dist = canon_l2_dist(tot);
        }
        const unsigned long long myrow = row0 + (unsigned)rsel;
        const float thr = *(volatile float *)&ctl->thr_f;
        if (myrow < p.n && owner && !(dist > thr)) cand_append(ctl, buf, make_key(dist, (uint32_t)myrow), p.cap, water);
    }

    if (!finish_and_merge(ctl, buf, p, tid)) return;
    publish(buf, ctl->cnt, p, ctl, buf, 0, tid);   // exact keys: nothing to prove, the flag is clear
}

// -------------------------------------------------------------------------------------------------
// helpers
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_norms_kernel(const float4 *rows4, uint32_t row0, uint32_t n, uint32_t ld4,
                                                        float *inv_norm, float *sq_norm, int *flags) {
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = w; r < n; r += nw) {
        const float4 *x = rows4 + (size_t)(row0 + r) * ld4;
        double p[4] = {0.0, 0.0, 0.0, 0.0};
        int bad = 0;
        for (uint32_t c = lane; c < ld4; c += 32) {
            const float4 v = ldg_stream(x + c);
            const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
            if (!(m <= 3.4028234663852886e38f) || v.x != v.x || v.y != v.y || v.z != v.z || v.w != v.w) bad |= 1;
            if (m > 0x1p40f) bad |= 2;
            canon_accum<true>(v, v, p);
        }
        const double nx = canon_warp_tree(canon_lane_fold(p));
        if (nx > 0.0 && (nx < 0x1p-80 || nx > 0x1p100)) bad |= 2;
        bad = __reduce_or_sync(kFull, bad);
        if (lane == 0) {
            inv_norm[row0 + r] = nx > 0.0 ? __double2float_rn(1.0 / sqrt(nx)) : 0.f;
            const float sqf = __double2float_rn(nx);
            sq_norm[row0 + r] = sqf;
            if (bad) atomicOr(flags, bad);
            if (!(bad & 1)) atomicMax(reinterpret_cast<unsigned int *>(flags) + 1, __float_as_uint(sqf));
        }
    }
}

__global__ void __launch_bounds__(256) fill_synthetic_kernel(float4 *rows4, uint32_t row0, uint32_t n, uint32_t dim,
                                                             uint32_t ld4, uint32_t rank, uint32_t world, uint32_t k0,
                                                             uint32_t k1) {
    const unsigned long long total = (unsigned long long)n * ld4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint32_t r = (uint32_t)(t / ld4), c = (uint32_t)(t % ld4);
        const unsigned long long e = row_id(row0 + r, rank, world) * dim + 4ull * c;  // first element of this float4
        float v[4];
        if ((dim & 3u) == 0) {
            const uint4 w = philox4x32_10((uint32_t)(e >> 2), (uint32_t)(e >> 34), k0, k1);
            v[0] = word_to_unit(w.x); v[1] = word_to_unit(w.y); v[2] = word_to_unit(w.z); v[3] = word_to_unit(w.w);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (4 * c + j < dim) {
                    const unsigned long long ee = e + j;
                    const uint4 w = philox4x32_10((uint32_t)(ee >> 2), (uint32_t)(ee >> 34), k0, k1);
                    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
                    v[j] = word_to_unit(ws[ee & 3]);
                } else {
                    v[j] = 0.f;
                }
            }
        }
        rows4[(size_t)(row0 + r) * ld4 + c] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

__global__ void pad_queries_kernel(const float *src, float *dst, uint32_t b, uint32_t dim, uint32_t ld) {
    const uint32_t total = b * ld;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t r = t / ld, c = t % ld;
        dst[t] = c < dim ? src[(size_t)r * dim + c] : 0.f;
    }
}

// Merge g lists of k hits, each sorted by (dist, id), into the k best under the same order.  No sort: every entry finds its
// rank -- its index in its own list plus, by binary search, the number of entries of every other list that precede it
// (ids are unique, so no two entries compare equal) -- and the entries ranked below k go straight to their slot.  Shards
// interleave in id space (block-cyclic sharding), so the order has to be the full (dist, id) one.  `list_of(s)` returns
// list s (global memory, possibly written by a peer: read volatile).  All threads of the CTA; ends with a barrier.
template <typename ListOf>
__device__ __forceinline__ void ranked_merge(ListOf list_of, uint32_t g, uint32_t k, unsigned long long *out_ids, float *out_dist, int tid) {
    for (uint32_t i = tid; i < k; i += kScanThreads) {
        out_ids[i] = kKeyMax;
        out_dist[i] = __int_as_float(0x7f800000);
    }
    __syncthreads();
    for (uint32_t e = tid; e < g * k; e += kScanThreads) {
        const uint32_t s = e / k, j = e % k;
        const Hit *mine = list_of(s) + j;
        const unsigned long long id = *(volatile const unsigned long long *)&mine->id;
        if (id == kKeyMax) continue;   // padding (a shard with fewer than k rows)
        const float dist = *(volatile const float *)&mine->dist;
        uint32_t rank = j;
        for (uint32_t t = 0; t < g && rank < k; ++t) {
            if (t == s) continue;
            const Hit *lt = list_of(t);
            uint32_t lo = 0, hi = k;       // entries of list t before (dist, id): lower bound
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                const unsigned long long oid = *(volatile const unsigned long long *)&lt[mid].id;
                const float od = *(volatile const float *)&lt[mid].dist;
                const bool before = oid != kKeyMax && (od < dist || (od == dist && oid < id));
                if (before) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_ids[rank] = id;
            out_dist[rank] = dist;
        }
    }
    __syncthreads();
}

// One CTA per query: merge the g gathered lists under (dist, id).
__global__ void __launch_bounds__(kScanThreads) merge_hits_kernel(const Hit *lists, uint32_t g, uint32_t b, uint32_t k,
                                                                  unsigned long long *out_ids, float *out_dist) {
    const uint32_t qi = blockIdx.x;
    ranked_merge([&](uint32_t s) { return lists + ((size_t)s * b + qi) * k; }, g, k, out_ids + (size_t)qi * k, out_dist + (size_t)qi * k,
                 threadIdx.x);
}

// -------------------------------------------------------------------------------------------------
// Fused exchange + merge over NVLink peer memory (row-sharded contexts, one process per GPU).
//
// Every rank owns an exchange window (cudaMalloc + CUDA IPC, mapped by all peers at context creation):
//   flags[3][world][kXchgMaxB]  u32 sequence numbers,   data[3][world][kXchgMaxHits] Hit
// One CTA per query: (1) PUSH this rank's k hits of the query into slot [seq&1][rank] of EVERY rank's window with
// plain stores through the peer mappings (NVLink P2P), fence at system scope, then publish flag = seq in every
// window; (2) WAIT until all `world` flags of the query in the LOCAL window carry seq (the peers' pushes);
// (3) MERGE the world lists from the local window under (dist, id) (ranked_merge) and write the final ids / distances.
// This replaces ncclAllGather + merge_hits_kernel (two launches, ~20-30 us of latency at 8 GPUs).  For single-query
// scans it is not even a launch: the last CTA of fast_scan_kernel / exact_scan_kernel runs it right after it has written
// the rank's k hits (exchange_device below), so a sharded search is ONE kernel per rank.
// Searches are collective and sequence numbers advance in lockstep.  The slot of an exchange is seq % 3: a conditional
// exact re-scan (run by all ranks or by none -- the decision is the OR of all ranks' guard flags, which travels in
// hit[0].pad) takes a sequence number whether it runs or not, so a slot is re-used after three numbers, of which at
// least one (a first-pass scan) was really exchanged: a rank pushes into a slot only after it has merged that
// exchange, which needed every peer's push, which every peer issued after merging everything before it -- including
// the exchange that used the slot last.
// root == kXchgAllRanks: all-to-all as above (one process per GPU: every rank returns the global answer).
// root == r (single-process multi-GPU contexts, where the host reads the answer from device r only): the ranks push
// into rank r's window alone and only rank r waits and merges; the host does not issue search s+1 before it has
// read the answer of search s, which is what keeps the two slots sufficient there.
// -------------------------------------------------------------------------------------------------
// One CTA, one query: push / wait / merge as described above.  `local` = this rank's k hits of the query (hit[0].pad carries
// the rank's guard flag), out_* = the query's k output slots.  buf: >= next_pow2(world * k) keys of shared memory.
// Returns (to every thread) the OR of the ranks' guard flags: the GLOBAL "this query needs the exact scan" decision, the
// same on every rank that merges.  Ranks that only push (root mode, rank != root) return 0.
__device__ __forceinline__ int exchange_device(const XchgArgs &x, const Hit *local, uint32_t k, unsigned long long *buf,
                                               unsigned long long *out_ids, float *out_dist, int tid) {
    const uint32_t rank = x.rank, world = x.world, seq = x.seq, qi = x.qi;
    const uint32_t slot = seq % kXchgSlots;
    const size_t flags_bytes = (size_t)kXchgSlots * kXchgMaxWorld * kXchgMaxB * sizeof(unsigned int);
    auto flag_of = [&](unsigned char *w, uint32_t r) {
        return reinterpret_cast<unsigned int *>(w) + ((size_t)slot * kXchgMaxWorld + r) * kXchgMaxB + qi;
    };
    auto data_of = [&](unsigned char *w, uint32_t r) {
        return reinterpret_cast<Hit *>(w + flags_bytes) + ((size_t)slot * kXchgMaxWorld + r) * kXchgMaxHits + (size_t)qi * k;
    };
    const bool all = x.root == kXchgAllRanks;
    if (qi == 0 && tid == 0) *x.err = 0;   // a time-out (>= 2.5 s later) sets it to 1
    // (1) push
    if (all) {
        for (uint32_t i = tid; i < world * k; i += kScanThreads) {
            const uint32_t r = i / k, j = i % k;
            data_of(x.windows[r], rank)[j] = local[j];
        }
    } else {
        Hit *dst = data_of(x.windows[x.root], rank);
        for (uint32_t j = tid; j < k; j += kScanThreads) dst[j] = local[j];
    }
    __threadfence_system();
    __syncthreads();
    if (all) {
        if (tid < (int)world) *(volatile unsigned int *)flag_of(x.windows[tid], rank) = seq;
    } else {
        if (tid == 0) *(volatile unsigned int *)flag_of(x.windows[x.root], rank) = seq;
        if (rank != x.root) return 0;
    }
    // (2) wait for every peer's push of this query into MY window
    unsigned char *mine = x.windows[rank];
    if (tid < (int)world) {
        volatile unsigned int *f = flag_of(mine, tid);
        unsigned long long spins = 0;
        while ((int)(*f - seq) < 0) {
            __nanosleep(64);
            if (++spins > 40000000ull) {   // ~2.5 s: a peer never arrived; report instead of hanging the GPU
                atomicExch(x.err, 1);
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    // (3) merge under (dist, id); the ranks' guard flags ride in hit[0].pad
    int flagged = 0;
    if (tid < (int)world) flagged = (int)*(volatile const uint32_t *)&data_of(mine, tid)->pad;
    flagged = __syncthreads_or(flagged);
    ranked_merge([&](uint32_t s) { return data_of(mine, s); }, world, k, out_ids, out_dist, tid);
    (void)buf;
    return flagged;
}

// The exchange as a kernel of its own, one CTA per query: after the batched (tensor-core) pass, whose b x k local hits come
// out of batched_finish_kernel, and for calls too large for the scan kernels' fused form.
__global__ void __launch_bounds__(kScanThreads) exchange_merge_kernel(XchgArgs x, const Hit *local, uint32_t b, uint32_t k,
                                                                      unsigned long long *out_ids, float *out_dist) {
    extern __shared__ __align__(16) unsigned char smem[];
    x.qi = blockIdx.x;
    exchange_device(x, local + (size_t)x.qi * k, k, reinterpret_cast<unsigned long long *>(smem), out_ids + (size_t)x.qi * k,
                    out_dist + (size_t)x.qi * k, threadIdx.x);
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
int scan_sm_count(int device) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
    return sms;
}

namespace {

struct Variant {
    uint32_t ld4;
    int rows;  // rows per warp iteration
    void (*l2)(const ScanParams);
    void (*cs)(const ScanParams);
};

#define VROD_VARIANT(LD4, LPR, CH, RB) \
    { LD4, RB *(32 / LPR), fast_scan_kernel<LPR, CH, RB, false>, fast_scan_kernel<LPR, CH, RB, true> }

const Variant kVariants[] = {
    VROD_VARIANT(8, 8, 1, 8),      // d = 32
    VROD_VARIANT(16, 16, 1, 8),    // d = 64
    VROD_VARIANT(32, 32, 1, 8),    // d = 128
    VROD_VARIANT(64, 32, 2, 4),    // d = 256
    VROD_VARIANT(96, 32, 3, 4),    // d = 384
    VROD_VARIANT(128, 32, 4, 2),   // d = 512
    VROD_VARIANT(192, 32, 6, 2),   // d = 768
    VROD_VARIANT(256, 32, 8, 1),   // d = 1024
    VROD_VARIANT(384, 32, 12, 1),  // d = 1536
};
const Variant kGeneric = VROD_VARIANT(0, 32, 0, 2);

const Variant &pick_variant(uint32_t ld4) {
    for (const Variant &v : kVariants)
        if (v.ld4 == ld4) return v;
    return kGeneric;
}

constexpr int kExactRBL2 = 4, kExactRBCos = 2;

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

ScanPlan make_scan_plan(const ShardView &s, uint32_t k, int sm_count, bool exact) {
    ScanPlan pl{};
    const uint32_t ld4 = s.ld / 4;
    int rows;
    const void *fn;
    if (exact) {
        rows = s.metric ? kExactRBCos : kExactRBL2;
        fn = s.metric ? (const void *)exact_scan_kernel<kExactRBCos, true> : (const void *)exact_scan_kernel<kExactRBL2, false>;
        pl.kprime = (int)k;
    } else {
        const Variant &v = pick_variant(ld4);
        rows = v.rows;
        fn = s.metric ? (const void *)v.cs : (const void *)v.l2;
        pl.kprime = next_pow2((int)k + 16);
        if (pl.kprime < 32) pl.kprime = 32;
    }
    // capacity first (it sets the shared memory, which sets occupancy), grid second
    int max_grid = sm_count * 8;
    pl.cap = next_pow2(pl.kprime + max_grid + 64);
    if (pl.cap < 2 * pl.kprime) pl.cap = 2 * pl.kprime;
    if (pl.cap < 512) pl.cap = 512;
    pl.smem = kCtlBytes + (size_t)pl.cap * sizeof(unsigned long long);
    int bps = 1;
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, kScanThreads, pl.smem) != cudaSuccess || bps < 1) bps = 1;
    if (bps > 8) bps = 8;
    long long want = ((long long)s.n + (long long)rows * kScanWarps - 1) / ((long long)rows * kScanWarps);
    if (want < 1) want = 1;
    long long grid = (long long)sm_count * bps;
    if (grid > want) grid = want;
    pl.grid = (int)grid;
    // error bound of the f32 pass (DESIGN.md "Guard"): each term carries <= 2 roundings, the per-lane
    // chain is (ld/LPR) fmas long and the butterfly adds log2(LPR) more; 1.5x safety.
    const double u = ldexp(1.0, -24);
    const double chain = (double)((s.ld + 31) / 32 * 4 + 12);
    pl.eps = 1.5 * chain * u + (double)s.ld * ldexp(1.0, -50);  // + slack for the f64 sums' own rounding
    if (s.metric) pl.eps += 4.0 * u;
    return pl;
}

size_t scan_cand_bytes(int sm_count) { return (size_t)sm_count * 8 * 2048 * sizeof(unsigned long long); }

unsigned long long *g_scan_dbg = nullptr;   // set by scan_debug_enable() (development aid)

unsigned long long *scan_debug_enable() {
    if (!g_scan_dbg) {
        cudaMalloc(&g_scan_dbg, 64);
    }
    return g_scan_dbg;
}

static ScanParams make_params(const ShardView &s, const float *q, uint32_t k, const ScanPlan &plan, const ScanScratch &scr,
                              int rows_per_iter) {
    ScanParams p{};
    p.rows4 = reinterpret_cast<const float4 *>(s.rows);
    p.inv_norm = s.inv_norm;
    p.q4 = reinterpret_cast<const float4 *>(q);
    p.n = s.n;
    p.ld4 = s.ld / 4;
    p.rank = s.rank;
    p.world = s.world ? s.world : 1;
    p.k = k;
    p.kprime = plan.kprime;
    p.cap = plan.cap;
    // prune trigger of the scanning phase: early enough that the threshold starts filtering after a
    // few hundred rows, late enough that a CTA sorts only a handful of times per query
    p.water = 4 * plan.kprime > 256 ? 4 * plan.kprime : 256;
    if (p.water > plan.cap - kScanWarps * 32) p.water = plan.cap - kScanWarps * 32;
    const unsigned long long batches = ((unsigned long long)s.n + rows_per_iter - 1) / rows_per_iter;
    const unsigned long long tw = (unsigned long long)plan.grid * kScanWarps;
    p.iters = (uint32_t)((batches + tw - 1) / tw);
    p.blk_cand = scr.blk_cand;
    p.ticket = scr.ticket;
    p.counters = scr.counters;
    p.eps = plan.eps;
    p.dbg = g_scan_dbg;
    return p;
}

cudaError_t launch_fast_scan(const ShardView &s, const float *q, uint32_t k, const ScanPlan &plan, const ScanScratch &scr,
                             int *status, Hit *out, unsigned long long *out_ids, float *out_dist, cudaStream_t st, const XchgArgs *x) {
    const Variant &v = pick_variant(s.ld / 4);
    ScanParams p = make_params(s, q, k, plan, scr, v.rows);
    p.status = status;
    p.out = out;
    p.out_ids = out_ids;
    p.out_dist = out_dist;
    if (x) p.x = *x;
    (s.metric ? v.cs : v.l2)<<<plan.grid, kScanThreads, plan.smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_exact_scan(const ShardView &s, const float *q, uint32_t k, const ScanPlan &plan, const ScanScratch &scr,
                              const int *only_if_flag, Hit *out, unsigned long long *out_ids, float *out_dist,
                              cudaStream_t st, const XchgArgs *x, int *gstatus) {
    ScanParams p = make_params(s, q, k, plan, scr, s.metric ? kExactRBCos : kExactRBL2);
    p.only_if = only_if_flag;
    p.status = x ? gstatus : nullptr;   // exact keys need no proof: alone, the query's flag stays what the f32 pass left
    p.counters = nullptr;
    p.out = out;
    p.out_ids = out_ids;
    p.out_dist = out_dist;
    if (x) p.x = *x;
    if (s.metric) exact_scan_kernel<kExactRBCos, true><<<plan.grid, kScanThreads, plan.smem, st>>>(p);
    else exact_scan_kernel<kExactRBL2, false><<<plan.grid, kScanThreads, plan.smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_row_norms(const float *rows, uint32_t row0, uint32_t n, uint32_t ld, float *inv_norm, float *sq_norm,
                             int *flags, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    long long blocks = ((long long)n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    row_norms_kernel<<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const float4 *>(rows), row0, n, ld / 4, inv_norm,
                                                  sq_norm, flags);
    return cudaGetLastError();
}

cudaError_t launch_fill_synthetic(float *rows, uint32_t row0, uint32_t n, uint32_t dim, uint32_t ld, uint32_t rank, uint32_t world,
                                  uint64_t seed, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    fill_synthetic_kernel<<<148 * 16, 256, 0, st>>>(reinterpret_cast<float4 *>(rows), row0, n, dim, ld / 4, rank, world ? world : 1,
                                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    return cudaGetLastError();
}

cudaError_t launch_pad_queries(const float *src, float *dst, uint32_t b, uint32_t dim, uint32_t ld, cudaStream_t st) {
    if (b == 0) return cudaSuccess;
    pad_queries_kernel<<<64, 256, 0, st>>>(src, dst, b, dim, ld);
    return cudaGetLastError();
}

cudaError_t launch_merge_hits(const Hit *lists, uint32_t g, uint32_t b, uint32_t k, unsigned long long *out_ids,
                              float *out_dist, cudaStream_t st) {
    if (b == 0) return cudaSuccess;
    merge_hits_kernel<<<b, kScanThreads, 0, st>>>(lists, g, b, k, out_ids, out_dist);
    return cudaGetLastError();
}

size_t xchg_window_bytes() {
    return (size_t)kXchgSlots * kXchgMaxWorld * kXchgMaxB * sizeof(unsigned int) + (size_t)kXchgSlots * kXchgMaxWorld * kXchgMaxHits * sizeof(Hit);
}

cudaError_t launch_exchange_merge(const XchgArgs &x, const Hit *local, uint32_t b, uint32_t k, unsigned long long *out_ids,
                                  float *out_dist, cudaStream_t st) {
    if (b == 0) return cudaSuccess;
    const size_t smem = (size_t)next_pow2((int)(x.world * k) < 32 ? 32 : (int)(x.world * k)) * sizeof(unsigned long long);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(exchange_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    exchange_merge_kernel<<<b, kScanThreads, smem, st>>>(x, local, b, k, out_ids, out_dist);
    return cudaGetLastError();
}

}  // namespace vrod
