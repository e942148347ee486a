// tcgen05 probe (not product code): validates the shared-memory descriptor / instruction descriptor / TMEM layouts
// that the tensor-core FIR (f9_umma.cu) relies on, and measures tcgen05.mma issue cost for small N.
//   part 1: D[128 x 32] = A[128 x 32] * B[32 x 32]^T (fp16 in, fp32 out), K-major no-swizzle operands with a padded
//           K-chunk stride, as [N=32] then a second [N=16] MMA accumulating into columns 16..31; checked on the host.
//   part 2: cycles per MMA for M=128, N in {16,32,48,64,96,128,256}, A from shared memory (SS) and from TMEM (TS).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
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
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* err) {
    for (int i = 0; i < (1 << 22); ++i) if (mbar_try_wait(bar, parity)) return true;
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major, no swizzle: element (row r, k) at (k/8)*lbo + (r/8)*sbo + (r%8)*16 + (k%8)*2 bytes.
__host__ __device__ inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr >> 4) & 0x3fff);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t) 1 << 46;                       // descriptor version (Blackwell)
    return d;                                      // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
    return (1u << 4) /* D = f32 */ | (0u << 7) /* A = f16 */ | (0u << 10) /* B = f16 */ | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

constexpr int A_CH = 128 * 16 + 32;     // K-chunk block stride of A (bytes): 128 rows x 16 B + pad
constexpr int B_CH = 256 * 16 + 32;     // same for B (up to 256 rows)
constexpr int KCH  = 8;                 // K chunks held (K = 64)

struct Params { int swap_lbo_sbo; int nrep; };

__global__ void __launch_bounds__(384, 1)
probe_kernel(const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img, const __half* __restrict__ a_rm, float* __restrict__ d_out,
             long long* __restrict__ cyc, int* __restrict__ err, Params P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* As = smem;                       // KCH * A_CH
    uint8_t* Bs = smem + KCH * A_CH;          // KCH * B_CH
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < KCH * A_CH / 16; i += blockDim.x) reinterpret_cast<uint4*>(As)[i] = reinterpret_cast<const uint4*>(a_img)[i];
    for (int i = threadIdx.x; i < KCH * B_CH / 16; i += blockDim.x) reinterpret_cast<uint4*>(Bs)[i] = reinterpret_cast<const uint4*>(b_img)[i];
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    uint32_t phase = 0;

    // ---- part 1: correctness
    if (threadIdx.x == 0) {
        for (int ks = 0; ks < 2; ++ks) {                       // K = 32 in two K=16 steps
            uint32_t l = A_CH, s = 128, lb = B_CH;
            uint64_t ad = P.swap_lbo_sbo ? make_desc(smem_u32(As) + ks * 2 * A_CH, s, l) : make_desc(smem_u32(As) + ks * 2 * A_CH, l, s);
            uint64_t bd = P.swap_lbo_sbo ? make_desc(smem_u32(Bs) + ks * 2 * B_CH, s, lb) : make_desc(smem_u32(Bs) + ks * 2 * B_CH, lb, s);
            umma_f16_ss(tmem, ad, bd, make_idesc(128, 32), ks > 0);                 // D[:, 0:32]  = A * B[0:32]^T
            // second product into columns 16..31:  D[:, 16:32] += A * B[32:48]^T (rows 32..47 of the B image)
            uint64_t bd2 = P.swap_lbo_sbo ? make_desc(smem_u32(Bs) + ks * 2 * B_CH + 32 * 16, s, lb) : make_desc(smem_u32(Bs) + ks * 2 * B_CH + 32 * 16, lb, s);
            umma_f16_ss(tmem + 16, ad, bd2, make_idesc(128, 16), 1);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    bool ok = mbar_wait_bounded(&bar, phase, err); phase ^= 1;
    tc_fence_after();
    if (ok) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t) (warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) d_out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- part 2: issue cost.  Each variant: nrep MMAs walking the KCH chunks, one commit, wait.
    const int Ns[7] = {16, 32, 48, 64, 96, 128, 256};
    for (int mode = 0; mode < 4; ++mode) {
        for (int ni = 0; ni < 7; ++ni) {
            const int N = Ns[ni];
            long long t0 = 0, t1 = 0;
            const int warp_u = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0);
            if (warp_u == 0) {
                const uint32_t idesc = make_idesc(128, N);
                const int nacc = (mode >= 2) ? (N <= 32 ? 8 : (N <= 64 ? 4 : (N <= 128 ? 2 : 1))) : 1;
                const uint64_t ad0 = make_desc(smem_u32(As), A_CH, 128);
                const uint64_t bd0 = make_desc(smem_u32(Bs), B_CH, 128);
                uint32_t el;
                asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
                t0 = clock64();
                if (el) {
                    for (int r = 0; r < P.nrep; r += 8) {
                        #pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int ks = u & 3;
                            const uint64_t ad = ad0 + (uint64_t) ((ks * 2 * A_CH) >> 4);
                            const uint64_t bd = bd0 + (uint64_t) ((ks * 2 * B_CH) >> 4);
                            const uint32_t dcol = (uint32_t) ((u % nacc) * N);
                            if ((mode & 1) == 0) umma_f16_ss(tmem + dcol, ad, bd, idesc, 1);
                            else umma_f16_ts(tmem + dcol, tmem + 256 + ks * 8, bd, idesc, 1);
                        }
                    }
                    umma_commit(&bar);
                }
            }
            __syncwarp();
            ok = mbar_wait_bounded(&bar, phase, err); phase ^= 1;
            if (threadIdx.x == 0) { t1 = clock64(); cyc[mode * 7 + ni] = ok ? (t1 - t0) : -1; }
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        }
    }

    // ---- part 3: A operand in TMEM, written (a) by tcgen05.cp from the canonical smem tile, (b) by tcgen05.st from registers
    for (int variant = 0; variant < 2; ++variant) {
        const uint32_t acol = 256 + variant * 32, dcol = 64 + variant * 32;
        if (variant == 1) {
            for (int ks = 0; ks < 2; ++ks) {
                uint32_t w[8];
                const int row = warp * 32 + lane;
                for (int j = 0; j < 8; ++j) {
                    const __half lo = a_rm[row * 32 + ks * 16 + 2 * j], hi = a_rm[row * 32 + ks * 16 + 2 * j + 1];
                    w[j] = (uint32_t) __half_as_ushort(lo) | ((uint32_t) __half_as_ushort(hi) << 16);
                }
                const uint32_t taddr = tmem + acol + ks * 8 + ((uint32_t) (warp * 32) << 16);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                             :: "r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        const int warp_u = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0);
        if (warp_u == 0) {
            uint32_t el;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
            if (el) {
                for (int ks = 0; ks < 2; ++ks) {
                    if (variant == 0) {
                        const uint64_t sd = make_desc(smem_u32(As) + ks * 2 * A_CH, A_CH, 128);
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(tmem + acol + ks * 8), "l"(sd) : "memory");
                    }
                    const uint64_t bd = make_desc(smem_u32(Bs) + ks * 2 * B_CH, B_CH, 128);
                    umma_f16_ts(tmem + dcol, tmem + acol + ks * 8, bd, make_idesc(128, 32), ks > 0);
                    const uint64_t bd2 = make_desc(smem_u32(Bs) + ks * 2 * B_CH + 32 * 16, B_CH, 128);
                    umma_f16_ts(tmem + dcol + 16, tmem + acol + ks * 8, bd2, make_idesc(128, 16), 1);
                }
                umma_commit(&bar);
            }
        }
        __syncwarp();
        ok = mbar_wait_bounded(&bar, phase, err); phase ^= 1;
        tc_fence_after();
        if (ok) {
            uint32_t v[32];
            const uint32_t taddr = tmem + dcol + ((uint32_t) (warp * 32) << 16);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) d_out[(1 + variant) * 128 * 32 + (warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    // ---- part 4: tcgen05.cp throughput (128x256b = 4 KB per op), alone and interleaved 1:3 with N=32 TS MMAs
    for (int variant = 0; variant < 2; ++variant) {
        long long t0 = 0, t1 = 0;
        const int warp_u = __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0);
        if (warp_u == 0) {
            uint32_t el;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(el));
            const uint64_t sd0 = make_desc(smem_u32(As), A_CH, 128);
            const uint64_t bd0 = make_desc(smem_u32(Bs), B_CH, 128);
            const uint32_t idesc = make_idesc(128, 32);
            t0 = clock64();
            if (el) {
                for (int r = 0; r < P.nrep; r += 4) {
                    #pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(tmem + 256 + u * 8), "l"(sd0 + (uint64_t) ((u * 2 * A_CH) >> 4)) : "memory");
                        if (variant == 1) {
                            umma_f16_ts(tmem + 0, tmem + 256 + u * 8, bd0, idesc, 1);
                            umma_f16_ts(tmem + 32, tmem + 256 + u * 8, bd0, idesc, 1);
                            umma_f16_ts(tmem + 64, tmem + 256 + u * 8, bd0, idesc, 1);
                        }
                    }
                }
                umma_commit(&bar);
            }
        }
        __syncwarp();
        ok = mbar_wait_bounded(&bar, phase, err); phase ^= 1;
        if (threadIdx.x == 0) { t1 = clock64(); cyc[28 + variant] = ok ? (t1 - t0) : -1; }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512) : "memory");
}

static void put(std::vector<uint8_t>& img, int ch_stride, int r, int k, __half v) {
    size_t off = (size_t) (k / 8) * ch_stride + (size_t) (r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
    memcpy(&img[off], &v, 2);
}

int main() {
    const int M = 128, K = 32, NB = 48;
    std::vector<float> A(M * K), B(NB * K);
    srand(1);
    for (auto& x : A) x = (float) __half2float(__float2half((rand() % 2001 - 1000) / 1000.0f));
    for (auto& x : B) x = (float) __half2float(__float2half((rand() % 2001 - 1000) / 1000.0f));
    std::vector<uint8_t> a_img(KCH * A_CH, 0), b_img(KCH * B_CH, 0);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) put(a_img, A_CH, r, k, __float2half(A[r * K + k]));
    for (int r = 0; r < NB; ++r) for (int k = 0; k < K; ++k) put(b_img, B_CH, r, k, __float2half(B[r * K + k]));
    std::vector<double> ref(M * 32);
    for (int r = 0; r < M; ++r) for (int n = 0; n < 32; ++n) {
        double s = 0; for (int k = 0; k < K; ++k) s += (double) A[r * K + k] * B[n * K + k];
        if (n >= 16) for (int k = 0; k < K; ++k) s += (double) A[r * K + k] * B[(32 + n - 16) * K + k];
        ref[r * 32 + n] = s;
    }
    uint8_t *da, *db; float* dd; __half* darm; long long* dc; int* de;
    CK(cudaMalloc(&da, a_img.size())); CK(cudaMalloc(&db, b_img.size())); CK(cudaMalloc(&dd, 3 * M * 32 * 4)); CK(cudaMalloc(&darm, M * K * 2));
    { std::vector<__half> h(M * K); for (int i = 0; i < M * K; ++i) h[i] = __float2half(A[i]); CK(cudaMemcpy(darm, h.data(), M * K * 2, cudaMemcpyHostToDevice)); } CK(cudaMalloc(&dc, 32 * 8)); CK(cudaMalloc(&de, 4));
    CK(cudaMemcpy(da, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
    const size_t smem = KCH * A_CH + KCH * B_CH;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    for (int swap = 0; swap < 2; ++swap) {
        CK(cudaMemset(dd, 0, 3 * M * 32 * 4)); CK(cudaMemset(dc, 0, 32 * 8)); CK(cudaMemset(de, 0, 4));
        Params P{swap, 2048};
        probe_kernel<<<1, 128, smem>>>(da, db, darm, dd, dc, de, P);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("swap=%d kernel error: %s\n", swap, cudaGetErrorString(e)); return 1; }
        std::vector<float> D(3 * M * 32); long long cyc[32]; int err;
        CK(cudaMemcpy(D.data(), dd, 3 * M * 32 * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(cyc, dc, sizeof cyc, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&err, de, 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; for (int i = 0; i < M * 32; ++i) maxerr = fmax(maxerr, fabs(D[i] - ref[i]));
        printf("swap_lbo_sbo=%d timeout=%d max|D-ref|=%.3g  D[0,0]=%g ref=%g  D[5,20]=%g ref=%g\n", swap, err, maxerr, D[0], ref[0], D[5 * 32 + 20], ref[5 * 32 + 20]);
        for (int v = 1; v < 3; ++v) { double me = 0; for (int i = 0; i < M * 32; ++i) me = fmax(me, fabs(D[v * M * 32 + i] - ref[i])); printf("  A in TMEM via %s: max|D-ref|=%.3g\n", v == 1 ? "tcgen05.cp" : "tcgen05.st", me); }
        if (swap == 0) {
            printf("  tcgen05.cp 128x256b alone: %.2f cycles/op;  cp + 3 TS MMAs (N=32): %.2f cycles/group\n", (double) cyc[28] / P.nrep, (double) cyc[29] / P.nrep);
            const int Ns[7] = {16, 32, 48, 64, 96, 128, 256};
            const char* names[4] = {"SS same-D", "TS same-D", "SS rot-D", "TS rot-D"};
            for (int mode = 0; mode < 4; ++mode) for (int i = 0; i < 7; ++i)
                printf("  %s N=%3d : %8.2f cycles/MMA (floor N/2 = %d)\n", names[mode], Ns[i], (double) cyc[mode * 7 + i] / P.nrep, Ns[i] / 2);
        }
    }
    return 0;
}
