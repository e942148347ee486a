// tcgen05 probe (not product code): can a *linear* fp16 buffer serve as a Hankel operand through a swizzled K-major
// shared-memory descriptor?  For integer upsampling (48 -> 192 kHz) the FIR is D[l, r] = sum_t A[l, t] * x[R*r + t] with a
// constant weight matrix A and a Hankel operand B[r, t] = x[R*r + t] (R = 128/L samples between columns).  A K-major
// operand with 32/64/128-byte swizzle has rows 32/64/128 bytes apart, so B would be the converted input itself, stored once
// (element m at swz(2m)), and a K step of 16 samples is a start-address advance of 32 bytes -- which walks across rows, i.e.
// the start address leaves the alignment the swizzle pattern repeats on.  This probe asks what the hardware reads then:
//   D = [I16 | 0] * B^T with B read through a descriptor (mode, start = base + off, base_offset variants),
//   buffer element m holds the value m % 2048 (exact in fp16) at swz(2m): D[k, n] names the element that was read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hankel_probe hankel_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// layout: 0 none, 2 = 128-byte swizzle, 4 = 64-byte, 6 = 32-byte (descriptor bits 61-63)
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr >> 4) & 0x3fff);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t) 1 << 46;
    d |= (uint64_t) (base_off & 7) << 49;
    d |= (uint64_t) (layout & 7) << 61;
    return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__host__ __device__ inline uint32_t swz(uint32_t byte, int rowBytes) {            // Swizzle<B,4,3> on the byte address
    const uint32_t bits = rowBytes == 128 ? 7u : rowBytes == 64 ? 3u : rowBytes == 32 ? 1u : 0u;
    return byte ^ (((byte >> 7) & bits) << 4);
}

constexpr int kN = 64;
constexpr int kBufBytes = 32768;

struct Params { int rowBytes; int off; int baseOffMode; int bufShift; };   // bufShift: the buffer itself starts bufShift bytes after a 1024-byte boundary

__global__ void __launch_bounds__(128, 1)
probe_kernel(float* __restrict__ d_out, int* __restrict__ err, Params P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);       // the swizzle patterns repeat on 1024 bytes
    uint8_t* As = smem;                        // A: M = 128 x K = 16, no swizzle: (k/8)*2048 + (r/8)*128 + (r%8)*16 + (k%8)*2
    uint8_t* Bs = smem + 4096;                 // 1024-byte aligned
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 4096 / 2; i += blockDim.x) {
        const int byte = 2 * i;
        const int kc = byte / 2048, rem = byte % 2048, r = (rem / 128) * 8 + (rem % 128) / 16, k = kc * 8 + (rem % 16) / 2;
        reinterpret_cast<__half*>(As)[i] = __float2half((r < 16 && r == k) ? 1.0f : 0.0f);
    }
    // element m of the linear buffer (which starts at Bs + bufShift) lives at swz(address)
    for (int m = threadIdx.x; m < (kBufBytes - 2048) / 2; m += blockDim.x) {
        const uint32_t a = smem_u32(Bs) + (uint32_t) P.bufShift + 2u * (uint32_t) m;
        const uint32_t sa = swz(a, P.rowBytes);
        *reinterpret_cast<__half*>(smem + (sa - smem_u32(smem))) = __float2half((float) (m % 2048));
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t start = smem_u32(Bs) + (uint32_t) P.bufShift + (uint32_t) P.off;
        const uint32_t bo = P.baseOffMode == 0 ? 0u : P.baseOffMode == 1 ? ((start >> 7) & 7u) : (uint32_t) (P.baseOffMode - 2);
        const uint64_t ad = make_desc(smem_u32(As), 2048, 128, 0, 0);
        const uint64_t bd = P.rowBytes == 16 ? make_desc(start, 16, 128, 0, 0)          // no swizzle: rows 16 bytes apart (LBO = K chunk stride = 16: Hankel)
                                             : make_desc(start, 16, 8 * P.rowBytes, P.rowBytes == 128 ? 2 : P.rowBytes == 64 ? 4 : 6, bo);
        umma_f16_ss(tmem, ad, bd, make_idesc(128, kN), 0);
        umma_commit(&bar);
    }
    __syncwarp();
    bool ok = false;
    for (int i = 0; i < (1 << 22) && !ok; ++i) ok = mbar_try_wait(&bar, 0);
    if (!ok) atomicExch(err, 1);
    tc_fence_after();
    if (ok) {
        for (int c0 = 0; c0 < kN; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + (uint32_t) c0 + ((uint32_t) (warp * 32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * kN + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64) : "memory");
}

int main() {
    float* d_out; int* d_err;
    CK(cudaMalloc(&d_out, sizeof(float) * 128 * kN));
    CK(cudaMalloc(&d_err, sizeof(int)));
    const size_t smem = 4096 + kBufBytes + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    std::vector<float> h(128 * kN);
    const int modes[4] = {16, 32, 64, 128};
    for (int mi = 0; mi < 4; ++mi) {
        const int rb = modes[mi];
        for (int shift = 0; shift <= (rb == 16 ? 0 : 512); shift += 512) {
            for (int off = 0; off <= 1024; off += 32) {
                const int nBo = rb == 16 ? 1 : 2;
                for (int bom = 0; bom < nBo; ++bom) {
                    Params P{rb, off, bom, shift};
                    CK(cudaMemset(d_out, 0xff, sizeof(float) * 128 * kN));
                    CK(cudaMemset(d_err, 0, sizeof(int)));
                    probe_kernel<<<1, 128, smem>>>(d_out, d_err, P);
                    CK(cudaDeviceSynchronize());
                    int err; CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(h.data(), d_out, sizeof(float) * 128 * kN, cudaMemcpyDeviceToHost));
                    // Hankel expectation: D[k, n] = element (off/2 + n*rb/2 + k)
                    int bad = 0, firstBadN = -1, firstBadK = -1;
                    for (int n = 0; n < kN; ++n)
                        for (int k = 0; k < 16; ++k) {
                            const int want = (off / 2 + n * rb / 2 + k) % 2048;
                            if ((int) h[k * kN + n] != want) { if (!bad) { firstBadN = n; firstBadK = k; } ++bad; }
                        }
                    printf("rows %3d B  shift %4d  off %4d  base_offset %s : %s", rb, shift, off, bom == 0 ? "0      " : "(a>>7)&7",
                           err ? "TIMEOUT" : bad == 0 ? "HANKEL OK" : "differs");
                    if (bad) {
                        printf(" (%d of %d; first at n=%d k=%d)  got n=0..9,k=0:", bad, kN * 16, firstBadN, firstBadK);
                        for (int n = 0; n < 10; ++n) printf(" %d", (int) h[0 * kN + n]);
                        printf("  want:");
                        for (int n = 0; n < 10; ++n) printf(" %d", (off / 2 + n * rb / 2) % 2048);
                        printf("  | k=8:");
                        for (int n = 0; n < 6; ++n) printf(" %d", (int) h[8 * kN + n]);
                    }
                    printf("\n");
                }
            }
        }
    }
    return 0;
}
