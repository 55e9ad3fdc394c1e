// tmem_ld_probe.cu -- development microbenchmark (not product code): how fast can the epilogue warps read TMEM?
// nwarps warps (warp w reads lane quadrant w % 4) load `cols` accumulator columns per iteration with one of the
// tcgen05.ld.32x32b shapes and fold them with a max (so the loads stay alive); prints bytes / clock / SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e__)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int X>
__device__ __forceinline__ void tc_ld(uint32_t taddr, uint32_t *r);
template <>
__device__ __forceinline__ void tc_ld<16>(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tc_ld<32>(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tc_ld<64>(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]),
                   "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]),
                   "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]),
                   "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// DEPTH = loads in flight before the wait (1: load, wait, use; 2: next load issued before the current one is used)
template <int X, int DEPTH>
__global__ void __launch_bounds__(256, 1) ld_kernel(int iters, int cols, long long *out) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * cols;   // two warps of a quadrant read different columns
    uint32_t ra[X], rb[X];
    float m = 0.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (DEPTH == 1) {
            for (int c = 0; c < cols; c += X) {
                tc_ld<X>(tmem + c, ra);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < X; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
            }
        } else {
            tc_ld<X>(tmem, ra);
            tc_wait_ld();
            for (int c = 0; c < cols; c += 2 * X) {
                tc_ld<X>(tmem + c + X, rb);
#pragma unroll
                for (int j = 0; j < X; ++j) m = fmaxf(m, __uint_as_float(ra[j]));
                tc_wait_ld();
                if (c + 2 * X < cols) tc_ld<X>(tmem + c + 2 * X, ra);
#pragma unroll
                for (int j = 0; j < X; ++j) m = fmaxf(m, __uint_as_float(rb[j]));
                if (c + 2 * X < cols) tc_wait_ld();
            }
        }
    }
    const long long t1 = clock64();
    if (m == 123.456f) out[1] = 1;
    if (threadIdx.x == 0) out[blockIdx.x * 2] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int X, int DEPTH>
static void run(int nwarps, int cols, long long *d_out) {
    const int iters = 20000;
    ld_kernel<X, DEPTH><<<148, nwarps * 32>>>(iters, cols, d_out);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    const double bytes = (double)iters * cols * 32 * 4 * nwarps;
    printf("32x32b.x%-3d depth %d  %d warps x %3d columns/iter: %7.1f B/clk/SM   (a 128 x 256 f32 accumulator = %5.0f cycles)\n", X, DEPTH, nwarps, cols,
           bytes / (double)h[0], 131072.0 / (bytes / (double)h[0]));
}

int main() {
    long long *d_out;
    CK(cudaMalloc(&d_out, sizeof(long long) * 2 * 148));
    for (int nw : {4, 8}) {
        const int cols = nw == 4 ? 256 : 128;
        run<16, 1>(nw, cols, d_out);
        run<32, 1>(nw, cols, d_out);
        run<64, 1>(nw, cols, d_out);
        run<16, 2>(nw, cols, d_out);
        run<32, 2>(nw, cols, d_out);
        run<64, 2>(nw, cols, d_out);
    }
    printf("done\n");
    return 0;
}
