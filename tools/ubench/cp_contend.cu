// tcgen05.cp / tcgen05.mma cost while other warps load the shared-memory pipe (not product code).
// Warp 0 issues; warps 4..11 run STS.64 (like the FIR's loaders) or LDS/global-store loops until told to stop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cp_contend cp_contend.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t) ((saddr >> 4) & 0x3fff) | ((uint64_t) ((lbo >> 4) & 0x3fff) << 16) | ((uint64_t) ((sbo >> 4) & 0x3fff) << 32) | ((uint64_t) 1 << 46);
}
constexpr int A_CH = 128 * 16 + 32;
template <int N, int sameB, int sameD>
__global__ void __launch_bounds__(384, 1) k(long long* cyc, int hammer, int nrep, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar; __shared__ uint32_t tslot; __shared__ volatile int stop;
    uint8_t* As = smem;                 // 8 chunks
    uint8_t* Bs = smem + 8 * A_CH;      // 16 KB of weights
    uint8_t* Hs = Bs + 16384;           // 64 KB hammer region
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (8 * A_CH + 16384 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tslot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
    asm volatile("fence.proxy.async.shared::cta;"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tslot;
    if (warp == 0) {
        uint32_t el; asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
        const uint64_t sd0 = make_desc(smem_u32(As), A_CH, 128), bd0 = make_desc(smem_u32(Bs), 512, 128);
        const uint32_t idesc = (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
        long long t0 = clock64();
        if (el) {
            for (int r = 0; r < nrep; r += 4) {
                #pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (mode != 1) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(tmem + 448 + u * 8), "l"(sd0 + (uint64_t) ((u * 2 * A_CH) >> 4)) : "memory");
                    if (mode >= 1) {
                        #pragma unroll
                        for (int m = 0; m < 6; ++m)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                         :: "r"(tmem + (uint32_t) (sameD ? 0 : ((u * 6 + m) % (384 / N)) * N)), "r"(tmem + 448 + u * 8), "l"(bd0 + (uint64_t) (sameB ? 0 : (((r + u) * 6 + m) % (16 * 16 / N)) * (N * 4))), "r"(idesc), "r"(1u) : "memory");
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        for (int i = 0; i < (1 << 24); ++i) if (mbar_try_wait(&bar, 0)) break;
        long long t1 = clock64();
        if (lane == 0) { cyc[0] = t1 - t0; stop = 1; }
    } else if (warp >= 4 && hammer) {
        uint32_t off = (uint32_t) ((warp - 4) * 8192 + lane * 8);
        uint2 v = make_uint2(lane, warp);
        unsigned long long cnt = 0;
        while (!stop) {
            #pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (hammer == 1) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(smem_u32(Hs) + off + (uint32_t) (i * 256)), "r"(v.x), "r"(v.y) : "memory");
                else { uint2 w; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "r"(smem_u32(Hs) + off + (uint32_t) (i * 256)) : "memory"); v.x += w.x; }
            }
            ++cnt;
        }
        if (lane == 0) cyc[1 + warp] = (long long) cnt + (v.x & 1);
    }
    asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}
int main() {
    long long* d; cudaMalloc(&d, 16 * 8);
    const size_t smem = 8 * A_CH + 16384 + 65536;
#define RUN(N, SB, SD) for (int mode = 1; mode < 3; ++mode) { \
        cudaFuncSetAttribute(k<N, SB, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
        cudaMemset(d, 0, 16 * 8); const int nrep = 4096; \
        k<N, SB, SD><<<1, 384, smem>>>(d, 0, nrep, mode); \
        cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; } \
        long long h[16]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost); \
        printf("N=%2d %s %s %s : %7.1f cycles per step of 6 MMAs\n", N, SB ? "same B tile " : "varying B   ", SD ? "same D    " : "rotating D", mode == 1 ? "no cp  " : "with cp", (double) h[0] / nrep); }
    RUN(16, 0, 0) RUN(16, 0, 1) RUN(16, 1, 0) RUN(16, 1, 1)
    RUN(32, 0, 0) RUN(32, 0, 1) RUN(32, 1, 0) RUN(32, 1, 1)
    RUN(64, 0, 0) RUN(64, 1, 1)
    return 0;
}
