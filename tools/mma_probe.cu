// mma_probe.cu -- development microbenchmark (not product code): what bounds the tcgen05.mma issue rate of the
// batched tile kernel's inner loop on B200?  One CTA (or CTA pair) per SM issues the kernel's own MMA shape back to
// back on operands that are already resident in shared memory -- no TMA ring, no epilogue -- and then again with
// (a) bulk copies streaming into other shared-memory stages (the TMA ring's write traffic), (b) epilogue warps
// reading the accumulators with tcgen05.ld, (c) both.  Prints SM cycles per MMA instruction for every mode.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e__ = (x);                                                                  \
        if (e__ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((unsigned short)3)
                 : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_mma_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accum) {
    if constexpr (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
            "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum)
            : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// K-major, NO swizzle ("interleaved"): core matrices of 8 rows x 16 bytes stored contiguously (128 B); the two K chunks of
// one K = 16 step are LBO bytes apart, consecutive 8-row groups SBO bytes apart.
__device__ __forceinline__ uint64_t umma_desc_plain(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16, bf16 x bf16 -> f32, K-major both
__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Ctl {
    uint64_t done, cp[4], stop;
    uint32_t tmem_base;
    volatile int finished;
};

constexpr int kStages = 6;           // row-slab stages (16 KB each) the MMAs cycle through
constexpr int kSlabA = 16384;
constexpr int kSlabs = 3;            // query slabs (like d = 128 in bf16 mode: 64 + 64 + 16 columns)

// mode bits: 1 = bulk copies stream into spare stages, 2 = epilogue warps read the accumulators
template <int CG>
__global__ void __launch_bounds__(320, 1) probe_kernel(int N, int tiles, int mode, const unsigned char *src, long long *out, int ksteps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int slab_b = N / CG * 128;                 // bytes of one query slab held by this CTA
    unsigned char *q_s = smem;
    unsigned char *a_s = smem + kSlabs * slab_b;
    unsigned char *spare = a_s + kStages * kSlabA;   // 2 more stages: targets of the bulk copies
    Ctl *ctl = reinterpret_cast<Ctl *>(spare + 2 * kSlabA);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
    // operands: deterministic small bf16 values (not zeros: the datapath should toggle like real data)
    for (int i = tid; i < (kSlabs * slab_b + kStages * kSlabA) / 4; i += blockDim.x) {
        const uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
        const uint32_t lo = 0x3c00u | ((h >> 3) & 0x1ffu) | ((h & 1u) << 15), hi = 0x3c00u | ((h >> 13) & 0x1ffu) | (((h >> 1) & 1u) << 15);
        reinterpret_cast<uint32_t *>(smem)[i] = lo | (hi << 16);
    }
    if (tid == 0) {
        mbar_init(&ctl->done, 1);
        for (int i = 0; i < 4; ++i) mbar_init(&ctl->cp[i], 1);
        ctl->finished = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy operand writes -> async proxy (MMA) reads
    if (warp == 1) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_base)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = ctl->tmem_base;

    if (warp == 0) {
        // bulk copies into the spare stages, as fast as they complete (two in flight), until the MMA warp is done
        if (lane == 0 && (mode & 1)) {
            long long copies = 0;
            uint32_t ph[2] = {0, 0};
            const long long t0 = clock64();
            for (int s = 0; s < 2; ++s) {
                mbar_expect_tx(&ctl->cp[s], kSlabA);
                bulk_g2s(spare + s * kSlabA, src + ((size_t)(blockIdx.x * 2 + s) % 2048) * kSlabA, kSlabA, &ctl->cp[s]);
            }
            while (!ctl->finished) {
                for (int s = 0; s < 2; ++s) {
                    mbar_wait(&ctl->cp[s], ph[s]);
                    ph[s] ^= 1;
                    ++copies;
                    mbar_expect_tx(&ctl->cp[s], kSlabA);
                    bulk_g2s(spare + s * kSlabA, src + ((size_t)(blockIdx.x * 2 + s + copies * 7) % 2048) * kSlabA, kSlabA, &ctl->cp[s]);
                }
            }
            for (int s = 0; s < 2; ++s) mbar_wait(&ctl->cp[s], ph[s]);
            out[blockIdx.x * 4 + 2] = copies;
            out[blockIdx.x * 4 + 3] = clock64() - t0;
        }
    } else if (warp == 1) {
        // layout (mode >> 2): 0 = 128-byte swizzle slabs, 1 = plain blocks [group][chunk][row] (LBO 128, SBO 256),
        //                    2 = plain blocks [chunk][group][row] (LBO rows*16, SBO 128)
        const int layout = mode >> 2;
        if (rank == 0) {
            const uint32_t idesc = idesc_bf16(128 * CG, N);
            const uint32_t nb = N / CG;     // query rows held by this CTA
            uint32_t a_hi, b_hi, lbo_a, lbo_b, a_step, b_step;
            if (layout == 0) { a_hi = b_hi = 0x40004040u; lbo_a = lbo_b = 16; a_step = b_step = 2; }
            else if (layout == 1) { a_hi = b_hi = 0x4010u; lbo_a = lbo_b = 128; a_step = 4096 >> 4; b_step = (nb * 32) >> 4; }
            else { a_hi = b_hi = 0x4008u; lbo_a = 128 * 16; lbo_b = nb * 16; a_step = 4096 >> 4; b_step = (nb * 32) >> 4; }
            const uint32_t a_lo0 = ((smem_u32(a_s) >> 4) & 0x3FFFu) | ((lbo_a >> 4) << 16);
            const uint32_t b_lo0 = ((smem_u32(q_s) >> 4) & 0x3FFFu) | ((lbo_b >> 4) << 16);
            const long long t0 = clock64();
            int stage = 0;
            for (int t = 0; t < tiles; ++t) {
                const uint32_t d_tmem = tmem + (t & 1) * 256;
                if (elect_one()) {
                    for (int ks = 0; ks < ksteps; ++ks) {
                        uint32_t a_lo, b_lo;
                        if (layout == 0) {
                            a_lo = a_lo0 + ((stage + (ks >> 2)) % kStages) * (kSlabA >> 4) + (ks & 3) * 2;
                            b_lo = b_lo0 + (ks >> 2) * (slab_b >> 4) + (ks & 3) * 2;
                        } else {
                            a_lo = a_lo0 + stage * (kSlabA >> 4) + (ks & 3) * a_step + (ks >> 2) * 0;
                            b_lo = b_lo0 + ks * b_step;
                        }
                        tc_mma_lohi<CG>(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, ks != 0);
                    }
                }
                __syncwarp();
                stage = (stage + 1) % kStages;
            }
            if (elect_one()) {
                if constexpr (CG == 2) tc_commit_pair(&ctl->done);
                else tc_commit(&ctl->done);
            }
            __syncwarp();
            mbar_wait(&ctl->done, 0);
            const long long t1 = clock64();
            if (lane == 0) {
                out[blockIdx.x * 4 + 0] = t1 - t0;
                out[blockIdx.x * 4 + 1] = (long long)tiles * ksteps;
                ctl->finished = 1;
            }
        } else if (lane == 0 && CG == 2) {
            mbar_wait(&ctl->done, 0);   // the pair's commit arrives in both CTAs
            ctl->finished = 1;
        }
    } else if (mode & 2) {
        // epilogue warps: read this warp's lane quadrant x 128 columns, over and over, like one tile's filter
        const int quad = warp & 3, half = (warp - 2) >> 2;
        uint32_t ra[32], acc = 0;
        long long loads = 0;
        while (!ctl->finished) {
            for (int cb = 0; cb < 4; ++cb) {
                tc_ld32(tmem + ((uint32_t)(quad * 32) << 16) + half * 128 + cb * 32, ra);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= ra[j];
                ++loads;
            }
        }
        if (acc == 0x12345678u && lane == 0) out[0] = -1;   // keep the loads alive
        (void)loads;
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int CG>
static void run(int grid, int N, int tiles, int mode, int ksteps, const unsigned char *src, long long *d_out, const char *label) {
    const size_t smem = (size_t)kSlabs * (N / CG * 128) + (kStages + 2) * kSlabA + sizeof(Ctl) + 64;
    CK(cudaFuncSetAttribute(probe_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaMemset(d_out, 0, sizeof(long long) * 4 * 148));
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, probe_kernel<CG>, N, tiles, mode, src, d_out, ksteps));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    static long long h[4 * 148];
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    double cyc = 0, n = 0, copies = 0, ccyc = 0;
    int issuers = 0;
    for (int b = 0; b < grid; ++b) {
        if (h[b * 4 + 1] == 0) continue;
        cyc += (double)h[b * 4 + 0];
        n += (double)h[b * 4 + 1];
        ++issuers;
    }
    for (int b = 0; b < grid; ++b) {
        copies += (double)h[b * 4 + 2];
        ccyc += (double)h[b * 4 + 3];
    }
    const double flops = 2.0 * 128 * CG * N * 16 * n;
    printf("%-92s grid %3d  %7.1f cyc/MMA  (%.0f%% of the %d-cycle formula)  %7.1f TFLOP/s  %6.2f ms", label, grid, cyc / n,
           100.0 * (128.0 * N / 256.0) / (cyc / n), (int)(128.0 * N / 256.0), flops / (ms * 1e-3) / 1e12, ms);
    if (mode & 1) printf("  | bulk copies %.1f B/clk/SM", copies * kSlabA / (ccyc > 0 ? ccyc / grid : 1) / grid);
    printf("\n");
    (void)issuers;
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
}

int main(int argc, char **argv) {
    const int tiles = argc > 1 ? atoi(argv[1]) : 4000;
    unsigned char *src = nullptr;
    long long *d_out = nullptr;
    CK(cudaMalloc(&src, (size_t)2048 * kSlabA + kSlabA));
    CK(cudaMemset(src, 0x3c, (size_t)2048 * kSlabA + kSlabA));
    CK(cudaMalloc(&d_out, sizeof(long long) * 4 * 148));
    const char *modes[4] = {"MMAs alone", "+ bulk copies into spare stages", "+ tcgen05.ld by 8 epilogue warps", "+ bulk copies + tcgen05.ld"};
    const char *layouts[3] = {"128B-swizzle slabs", "plain blocks [grp][chunk][row]", "plain blocks [chunk][grp][row]"};
    for (int grid : {1, 148}) {
        for (int layout = 0; layout < 3; ++layout)
            for (int mode : {0, 1, 2, 3}) {
                char label[160];
                snprintf(label, sizeof(label), "cg1 M128 N256 K16, %s, %s", layouts[layout], modes[mode]);
                run<1>(grid, 256, tiles, mode | (layout << 2), 9, src, d_out, label);
            }
        run<1>(grid, 128, tiles, 0, 9, src, d_out, "cg1 M128 N128 K16, 128B-swizzle slabs, MMAs alone");
        const int g2 = grid == 1 ? 2 : 148;
        for (int layout = 0; layout < 2; ++layout)
            for (int mode : {0, 3}) {
                char label[160];
                snprintf(label, sizeof(label), "cg2 M256 N256 K16, %s, %s", layouts[layout], modes[mode]);
                run<2>(g2, 256, tiles, mode | (layout << 2), 9, src, d_out, label);
            }
    }
    printf("done\n");
    return 0;
}
