// Micro-benchmarks that size the WindowedSinc kernel design (not product code):
//   (1) shared-memory pipe cost (cycles per warp instruction at saturation) of LDS.32/64/128 for address patterns:
//       uniform (broadcast), G distinct 16-byte chunks shared by 32/G lanes, fully distinct consecutive.
//   (2) FFMA vs FFMA2 issue rate with register operands only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_fma lds_fma.cu ; run on one GPU.
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 65536

template <int BYTES>
__global__ void lds_kernel(const int* __restrict__ lane_off, float* out, long long* cycles) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float) i;
    __syncthreads();
    const int off = lane_off[threadIdx.x & 31];          // float index, multiple of BYTES/4
    float acc = 0.f;
    __syncthreads();
    long long t0 = clock64();
    #pragma unroll 8
    for (int it = 0; it < ITER; ++it) {
        const int a = off + ((it & 7) * 64);             // stay inside a small footprint, keep pattern
        if (BYTES == 4) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned) __cvta_generic_to_shared(sm + a))); acc += v; }
        if (BYTES == 8) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned) __cvta_generic_to_shared(sm + a))); acc += v.x + v.y; }
        if (BYTES == 16) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned) __cvta_generic_to_shared(sm + a))); acc += v.x + v.y + v.z + v.w; }
    }
    __syncthreads();
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>   // 0: scalar FFMA, 1: FFMA2
__global__ void fma_kernel(float* out, long long* cycles, float a, float b) {
    float2 acc[16];
    #pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = make_float2(threadIdx.x * 0.001f + j, j * 0.5f);
    float2 x = make_float2(a, a), c = make_float2(b, b * 1.0001f);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
        #pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (MODE == 0) { acc[j].x = fmaf(x.x, c.x, acc[j].x); acc[j].y = fmaf(x.y, c.y, acc[j].y); }
            else acc[j] = __ffma2_rn(x, c, acc[j]);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0.f;
    #pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j].x + acc[j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int BYTES>
void run_lds(const char* name, const int* h_off, int warps) {
    int* d_off; float* d_out; long long* d_cyc;
    cudaMalloc(&d_off, 32 * 4); cudaMalloc(&d_out, 4 * 1024 * 4); cudaMalloc(&d_cyc, 8 * 8);
    cudaMemcpy(d_off, h_off, 32 * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    lds_kernel<BYTES><<<1, 32 * warps, 8192 * 4 + 64>>>(d_off, d_out, d_cyc);
    cudaEventRecord(e0);
    for (int r = 0; r < 20; ++r) lds_kernel<BYTES><<<1, 32 * warps, 8192 * 4 + 64>>>(d_off, d_out, d_cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    { cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) printf("launch error: %s\n", cudaGetErrorString(le)); }
    const double clk = ms * 1e-3 / 20 * 1.965e9;            // SM cycles per launch at the max clock (launch overhead ~3 us included)
    printf("LDS.%-3d %-34s warps=%2d  %.2f cycles per warp-instruction (SM pipe, events)\n", BYTES * 8, name, warps, clk / ((double) ITER * warps));
    cudaFree(d_off); cudaFree(d_out); cudaFree(d_cyc);
}

int main() {
    int off[32];
    const int warps = 16;
    // 32-bit
    for (int l = 0; l < 32; ++l) off[l] = 0;            run_lds<4>("uniform", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = l;            run_lds<4>("32 consecutive", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = l * 2;        run_lds<4>("stride 2 (2-way conflict)", off, warps);
    // 64-bit
    for (int l = 0; l < 32; ++l) off[l] = 0;            run_lds<8>("uniform", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = (l / 16) * 2; run_lds<8>("2 chunks x 16 lanes", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = l * 2;        run_lds<8>("32 consecutive", off, warps);
    // 128-bit
    for (int l = 0; l < 32; ++l) off[l] = 0;            run_lds<16>("uniform", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = (l / 16) * 4; run_lds<16>("2 chunks x 16 lanes (halves)", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = (l / 8) * 4;  run_lds<16>("4 chunks x 8 lanes (quarters)", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = (l % 4) * 4;  run_lds<16>("4 chunks, lane%4", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = (l / 4) * 4;  run_lds<16>("8 chunks x 4 lanes", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = (l % 8) * 4;  run_lds<16>("8 chunks, lane%8", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = (l / 2) * 4;  run_lds<16>("16 chunks x 2 lanes", off, warps);
    for (int l = 0; l < 32; ++l) off[l] = l * 4;        run_lds<16>("32 consecutive chunks", off, warps);

    float* d_out; long long* d_cyc; cudaMalloc(&d_out, 1024 * 4); cudaMalloc(&d_cyc, 64);
    for (int w : {4, 8, 16}) {
        fma_kernel<0><<<1, 32 * w>>>(d_out, d_cyc, 1.0001f, 0.9999f);
        fma_kernel<0><<<1, 32 * w>>>(d_out, d_cyc, 1.0001f, 0.9999f);
        long long c0 = 0; cudaMemcpy(&c0, d_cyc, 8, cudaMemcpyDeviceToHost);
        fma_kernel<1><<<1, 32 * w>>>(d_out, d_cyc, 1.0001f, 0.9999f);
        fma_kernel<1><<<1, 32 * w>>>(d_out, d_cyc, 1.0001f, 0.9999f);
        long long c1 = 0; cudaMemcpy(&c1, d_cyc, 8, cudaMemcpyDeviceToHost);
        const double fmas = (double) ITER * 32 * 32 * w;         // lane-FMAs
        printf("warps=%2d  scalar FFMA: %.1f FMA/clk/SM   FFMA2: %.1f FMA/clk/SM\n", w, fmas / c0, fmas / c1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
