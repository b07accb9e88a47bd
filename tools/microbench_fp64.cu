// microbench_fp64.cu -- FP64 facts behind the design of the local-BA kernel (csrc/ba2.cu), measured on the GPU box:
//   1. DFMA throughput per SM and the latency of a dependent DFMA chain
//   2. DMMA (mma.sync m8n8k4 f64, the FP64 tensor path) throughput and dependent latency
//   3. the Schur block product of one warp task -- sum over 32 landmarks of (Hpl D)(6x3) * Hpl^T(3x6) -- as
//        (a) 108 DFMA per lane with the sums in registers + the butterfly reduce-scatter flush (what ba2.cu does)
//        (b) operands staged through shared memory in fragment order + 24 DMMA per 32 landmarks (K = 96), no flush
//   4. the cost of a shared-memory publish -> barrier -> re-load round trip (what a scalar LDL^T pays per column)
// build + run (on the GPU box): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench_fp64.cu && /tmp/mb
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void dfma_throughput(double* out, int iters) {
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void dfma_latency(double* out, long long* cycles, int iters) {
    double a = threadIdx.x;
    const double b = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
    const long long t1 = clock64();
    out[threadIdx.x] = a;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void dmma_throughput(double* out, int iters) {
    double c[8][2];
    for (int k = 0; k < 8; ++k) c[k][0] = c[k][1] = 0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) dmma(c[k][0], c[k][1], a, b);
    double s = 0;
    for (int k = 0; k < 8; ++k) s += c[k][0] + c[k][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_latency(double* out, long long* cycles, int iters) {
    double c0 = 0, c1 = 0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) dmma(c0, c1, a, b);
    const long long t1 = clock64();
    out[threadIdx.x] = c0 + c1;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

// ---- the Schur block task: every lane owns a landmark with BD (6x3) and H (6x3) in registers (here: synthetic values)
template <int N>
__device__ __forceinline__ double warp_reduce_scatter(double* v, int lane) {
    int off = 16;
#pragma unroll
    for (int n = N / 2; n >= 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const double keep = up ? v[i + n] : v[i];
            const double send = up ? v[i] : v[i + n];
            v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
        }
    }
#pragma unroll
    for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xFFFFFFFFu, v[0], off);
    return v[0];
}

__device__ __forceinline__ void make_operands(int lane, int g, double BD[6][3], double H[6][3]) {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            BD[r][c] = 1e-3 * (lane + 1) + 0.1 * r - 0.01 * c + 1e-4 * g;
            H[r][c] = 2e-3 * (lane + 2) - 0.1 * c + 0.02 * r - 1e-4 * g;
        }
}

// (a) DFMA formulation: groups of 32 landmarks, sums in registers, one flush per task (= every `groups` groups)
__global__ void __launch_bounds__(256) schur_dfma(double* out, long long* cycles, int tasks, int groups) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t0 = clock64();
    double total = 0;
    for (int t = 0; t < tasks; ++t) {
        double acc[48];
#pragma unroll
        for (int i = 0; i < 48; ++i) acc[i] = 0;
        for (int g = 0; g < groups; ++g) {
            double BD[6][3], H[6][3];
            make_operands(lane, g + t, BD, H);
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 6; ++c) acc[r * 6 + c] += BD[r][0] * H[c][0] + BD[r][1] * H[c][1] + BD[r][2] * H[c][2];
        }
        const double lo = warp_reduce_scatter<32>(acc, lane), hi = warp_reduce_scatter<16>(acc + 32, lane);
        total += lo + hi;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    (void)warp;
}

// (b) DMMA formulation: operands through shared memory in fragment order (A[row][k], B[col][k], k = 3 * landmark + component)
__global__ void __launch_bounds__(256) schur_dmma(double* out, long long* cycles, int tasks, int groups) {
    extern __shared__ double s_dyn[];
    double (*sA)[8][97] = reinterpret_cast<double (*)[8][97]>(s_dyn);                     // [warp][row][k] (+1 padding)
    double (*sB)[8][97] = reinterpret_cast<double (*)[8][97]>(s_dyn + 8 * 8 * 97);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 2 * 8 * 8 * 97; i += 256) s_dyn[i] = 0;
    __syncthreads();
    const long long t0 = clock64();
    double total = 0;
    for (int t = 0; t < tasks; ++t) {
        double c0 = 0, c1 = 0;
        for (int g = 0; g < groups; ++g) {
            double BD[6][3], H[6][3];
            make_operands(lane, g + t, BD, H);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 6; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    sA[warp][r][3 * lane + c] = BD[r][c];
                    sB[warp][r][3 * lane + c] = H[r][c];
                }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 24; ++q) dmma(c0, c1, sA[warp][lane >> 2][4 * q + (lane & 3)], sB[warp][lane >> 2][4 * q + (lane & 3)]);
        }
        total += c0 + c1;   // the 6 x 6 block is already summed over the warp: lane l holds C[l / 4][2 (l % 4) + {0, 1}]
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = total;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// publish -> barrier -> re-load round trips
template <bool kCta>
__global__ void __launch_bounds__(256) roundtrip(double* out, long long* cycles, int iters) {
    __shared__ double s_v[2];
    double v = threadIdx.x;
    if (threadIdx.x == 0) s_v[0] = 1.0;
    __syncthreads();
    const long long t0 = clock64();
    if (kCta || threadIdx.x < 32)
        for (int i = 0; i < iters; ++i) {
            const double r = s_v[i & 1];
            v = fma(v, r, 1e-9);
            if (threadIdx.x == 0) s_v[(i + 1) & 1] = 1.0 + 1e-12 * v;
            if (kCta) __syncthreads();
            else __syncwarp();
        }
    const long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, max clock %d MHz\n", p.name, p.multiProcessorCount, clk / 1000);
    double* out;
    long long* cyc;
    CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
    CK(cudaMalloc(&cyc, 8));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    long long h;
    const int sm = p.multiProcessorCount;
    {
        const int iters = 20000, blocks = sm * 4, threads = 512;
        dfma_throughput<<<blocks, threads>>>(out, 100);
        cudaEventRecord(e0);
        dfma_throughput<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * threads * iters * 8;
        printf("DFMA throughput        %8.3f ms  %8.1f GFMA/s = %6.1f FMA/clk/SM (%.1f TFLOP/s)\n", ms, ops / ms / 1e6, ops / (ms * 1e-3) / sm / (clk * 1e3),
               2 * ops / ms / 1e9);
    }
    {
        const int iters = 100000;
        dfma_latency<<<1, 32>>>(out, cyc, iters);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("DFMA dependent chain   %8.1f cycles per operation (1 warp)\n", (double)h / iters);
    }
    {
        const int iters = 5000, blocks = sm * 4, threads = 512;
        dmma_throughput<<<blocks, threads>>>(out, 10);
        cudaEventRecord(e0);
        dmma_throughput<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = (double)blocks * (threads / 32) * iters * 8 * 256;   // m8n8k4 = 256 FMA per warp instruction
        printf("DMMA m8n8k4 throughput %8.3f ms  %8.1f GFMA/s = %6.1f FMA/clk/SM (%.1f TFLOP/s)\n", ms, fmas / ms / 1e6,
               fmas / (ms * 1e-3) / sm / (clk * 1e3), 2 * fmas / ms / 1e9);
    }
    {
        const int iters = 100000;
        dmma_latency<<<1, 32>>>(out, cyc, iters);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("DMMA dependent chain   %8.1f cycles per operation (1 warp)\n", (double)h / iters);
    }
    for (int groups : {1, 5}) {
        const int tasks = 200;
        schur_dfma<<<sm, 256>>>(out, cyc, tasks, groups);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double a = (double)h / tasks;
        CK(cudaFuncSetAttribute(schur_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8 * 8 * 97 * 8));
        schur_dmma<<<sm, 256, 2 * 8 * 8 * 97 * 8>>>(out, cyc, tasks, groups);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        const double b = (double)h / tasks;
        printf("Schur block task, %d group(s) of 32 landmarks per flush, 8 warps per SM: DFMA + butterfly %8.0f cycles, smem + DMMA %8.0f cycles\n",
               groups, a, b);
    }
    {
        const int iters = 20000;
        roundtrip<false><<<1, 256>>>(out, cyc, iters);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("publish -> __syncwarp -> re-load + 1 DFMA     %8.1f cycles per round trip\n", (double)h / iters);
        roundtrip<true><<<1, 256>>>(out, cyc, iters);
        CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
        printf("publish -> __syncthreads -> re-load + 1 DFMA  %8.1f cycles per round trip (8 warps)\n", (double)h / iters);
    }
    return 0;
}
