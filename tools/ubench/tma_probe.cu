// TMA probe (not product code): checks what the TMA-fed FIR loader relies on and measures what a shallow ring delivers.
//   part 1: a rank-2 tensor map whose row stride (p floats) is SMALLER than the box width, i.e. rows overlap in memory
//           (row y = samples y*p .. y*p + 31), 128-byte swizzle, map read from GLOBAL memory, negative / past-the-end x.
//   part 2: 148 CTAs stream [128 rows x 32 samples] boxes (rows p = 320 floats apart, 17 boxes per tile of 128 rows, as the
//           FIR does) through a ring of 2/3/4/8 stages; eight consumer warps read each stage back (4 x LDS.128 per thread).
//           Reports unique bytes per second, with and without a bulk L2 prefetch issued one tile ahead.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
    for (int i = 0; i < (1 << 24); ++i) if (mbar_try_wait(bar, parity)) return true;
    atomicExch(err, 1);
    return false;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------------- part 1
__global__ void check_kernel(const CUtensorMap* __restrict__ map, int x, int y, int rows, float* out, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    uint8_t* tile = (uint8_t*) (((uintptr_t) smem + 1023) & ~(uintptr_t) 1023);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(&bar, rows * 128); tma_load_2d(tile, map, x, y, &bar); }
    mbar_wait(&bar, 0, err);
    for (int i = threadIdx.x; i < rows * 32; i += blockDim.x) {
        const int r = i >> 5, e = i & 31;
        out[i] = *reinterpret_cast<const float*>(tile + r * 128 + (((e >> 2) ^ (r & 7)) << 4) + (e & 3) * 4);   // undo the 128-byte swizzle
    }
}

// ---------------------------------------------------------------------------------------------------- part 2
struct StreamArgs { const CUtensorMap* map; const float* base; int p, tilesPerCta, stagesPerTile, stages, prefetch; long long slabFloats; };
__global__ void __launch_bounds__(288, 1) stream_kernel(StreamArgs A, float* sink, long long* cyc, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[8], empty[8];
    uint8_t* ring = (uint8_t*) (((uintptr_t) smem + 1023) & ~(uintptr_t) 1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < A.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    const long long slab0 = (long long) blockIdx.x * A.slabFloats;        // this CTA's slab (floats from base)
    if (warp == 8) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int t = 0; t < A.tilesPerCta; ++t) {
                const long long row0 = slab0 + (long long) t * 128 * A.p;  // multiple of p: y = row0 / p
                if (A.prefetch && t + 1 < A.tilesPerCta) {
                    const char* nxt = reinterpret_cast<const char*>(A.base + row0 + 128LL * A.p);
                    for (int c = 0; c < 10; ++c) l2_prefetch(nxt + c * (128 * A.p * 4 / 10), 128 * A.p * 4 / 10);
                }
                for (int st = 0; st < A.stagesPerTile; ++st) {
                    if (!mbar_wait(empty + s, ph ^ 1, err)) return;
                    mbar_expect_tx(full + s, 16384);
                    tma_load_2d(ring + s * 16384, A.map, st * 32, (int) (row0 / A.p), full + s);
                    if (++s == A.stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        int s = 0; uint32_t ph = 0; float acc = 0.f;
        const int row = (warp & 3) * 32 + lane, h = warp >> 2;
        for (int i = 0; i < A.tilesPerCta * A.stagesPerTile; ++i) {
            if (!mbar_wait(full + s, ph, err)) return;
            const uint8_t* st = ring + s * 16384 + row * 128;
            #pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(st + (((4 * h + c) ^ (row & 7)) << 4));
                acc += v.x + v.y + v.z + v.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
            if (++s == A.stages) { s = 0; ph ^= 1; }
        }
        if (acc == 123.456f) sink[threadIdx.x] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int onlyCase = argc > 2 && !strcmp(argv[1], "case") ? atoi(argv[2]) : -1;
    const bool doStream = argc > 1 && !strcmp(argv[1], "stream");
    EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**) &encode, cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
    int* derr; CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));

    // ---- part 1
    if (!doStream) {
        const int N = 100000;
        const int NA = N + 60000;                          // allocation is larger than the tensor: overlapping rows reach past D0
        std::vector<float> h(NA); for (int i = 0; i < NA; ++i) h[i] = (float) i + 0.25f;
        float* d; CK(cudaMalloc(&d, NA * 4)); CK(cudaMemcpy(d, h.data(), NA * 4, cudaMemcpyHostToDevice));
        float* dout; CK(cudaMalloc(&dout, 128 * 32 * 4));
        CUtensorMap* dmap; CK(cudaMalloc(&dmap, sizeof(CUtensorMap)));
        CK(cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000));
        struct Case { int p, rows, x, y, m; };            // m: base offset in floats (the map's base is d + m, must be 16-byte aligned)
        const Case cases[] = {{320, 128, 0, 0, 0}, {320, 128, -200, 0, 0}, {320, 128, 7, 3, 0}, {588, 32, 147 + 5, 2, 0}, {4, 32, 9, 11, 0},
                              {320, 128, 99900 - 127 * 320, 0, 0}, {320, 128, 64, 180, 4}, {320, 128, 99990, 0, 0},
                              {320, 128, 4, 3, 0}, {320, 128, 7, 0, 0}, {320, 128, 2, 0, 0}, {320, 128, 1, 0, 0}, {147, 32, 8, 1, 0}};
        int ci = -1;
        for (const Case& c : cases) {
            if (++ci != onlyCase && onlyCase >= 0) continue;
            const long long D0 = N - c.m;
            cuuint64_t dims[2] = {(cuuint64_t) D0, (cuuint64_t) (D0 / c.p + 1)};
            cuuint64_t strides[1] = {(cuuint64_t) c.p * 4};
            cuuint32_t box[2] = {32, (cuuint32_t) c.rows}, es[2] = {1, 1};
            CUtensorMap hm;
            CUresult r = encode(&hm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d + c.m, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("part1 p=%d rows=%d: encode FAILED (%d)\n", c.p, c.rows, (int) r); continue; }
            CK(cudaMemcpy(dmap, &hm, sizeof hm, cudaMemcpyHostToDevice));
            check_kernel<<<1, 256, 128 * 128 + 1024>>>(dmap, c.x, c.y, c.rows, dout, derr);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("part1 kernel error: %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<float> o(c.rows * 32); CK(cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0, zeros = 0;
            for (int rr = 0; rr < c.rows; ++rr) for (int e2 = 0; e2 < 32; ++e2) {
                const long long xx = (long long) c.x + e2, yy = c.y + rr;
                float want = 0.f;
                if (xx >= 0 && xx < D0 && yy < (long long) dims[1]) { const long long idx = c.m + yy * c.p + xx; want = idx < NA ? h[idx] : -1.f; } else ++zeros;
                if (o[rr * 32 + e2] != want && want != -1.f) { if (bad < 3) printf("   mismatch row %d e %d: got %g want %g\n", rr, e2, o[rr * 32 + e2], want); ++bad; }
            }
            int err; CK(cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost));
            printf("part1 p=%d rows=%d x=%d y=%d m=%d: timeout=%d mismatches=%d (zero-filled %d)\n", c.p, c.rows, c.x, c.y, c.m, err, bad, zeros);
        }
    }

    // ---- part 2
    if (doStream) {
        const int p = 320, tilesPerCta = 40, stagesPerTile = 17, nCta = 148;
        const long long slab = (long long) tilesPerCta * 128 * p + 4096;        // floats per CTA
        const long long N = ((slab + p - 1) / p * p) * nCta + 8192;
        float* d; CK(cudaMalloc(&d, N * 4)); CK(cudaMemset(d, 0, N * 4));
        cuuint64_t dims[2] = {(cuuint64_t) N, (cuuint64_t) (N / p)};
        cuuint64_t strides[1] = {(cuuint64_t) p * 4};
        cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
        CUtensorMap hm;
        CUresult r = encode(&hm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("part2 encode failed %d\n", (int) r); return 1; }
        CUtensorMap* dmap; CK(cudaMalloc(&dmap, sizeof hm)); CK(cudaMemcpy(dmap, &hm, sizeof hm, cudaMemcpyHostToDevice));
        float* sink; CK(cudaMalloc(&sink, 4096)); long long* dcyc; CK(cudaMalloc(&dcyc, nCta * 8));
        float* flush; const size_t flushBytes = 256u << 20; CK(cudaMalloc(&flush, flushBytes));
        CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 1024));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        // slab start must be a multiple of p for y = row0 / p: slab = tilesPerCta*128*p + 4096 is not -> round the slab
        const long long slabR = (slab + p - 1) / p * p;
        for (int prefetch = 0; prefetch < 2; ++prefetch) for (int stages : {2, 3, 4, 8}) {
            StreamArgs A{dmap, d, p, tilesPerCta, stagesPerTile, stages, prefetch, slabR};
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                CK(cudaMemset(flush, rep, flushBytes));
                CK(cudaEventRecord(e0));
                stream_kernel<<<nCta, 288, stages * 16384 + 1024>>>(A, sink, dcyc, derr);
                CK(cudaEventRecord(e1));
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("part2 kernel error: %s\n", cudaGetErrorString(e)); return 1; }
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
            }
            int err; CK(cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost));
            std::vector<long long> cyc(nCta); CK(cudaMemcpy(cyc.data(), dcyc, nCta * 8, cudaMemcpyDeviceToHost));
            double avg = 0; for (long long c : cyc) avg += (double) c / nCta;
            const double uniq = (double) nCta * tilesPerCta * 128 * p * 4, fetched = (double) nCta * tilesPerCta * stagesPerTile * 16384.0;
            printf("part2 stages=%d prefetch=%d: %.3f ms  unique %.0f GB/s  fetched %.0f GB/s  %.0f clk per stage  (timeout=%d)\n", stages, prefetch, best,
                   uniq / best * 1e-6, fetched / best * 1e-6, avg / (tilesPerCta * stagesPerTile), err);
        }
    }
    return 0;
}
