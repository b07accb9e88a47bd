// microbench.cu -- integer-pipe throughput on the box's B200 (denominators for the matcher's ALU roofline).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu && gpurun_out/microbench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = __popc(a[i]) + a[(i + 1) & 7];           // 1 POPC + 1 IADD
            if (OP == 1) a[i] = (a[i] ^ a[(i + 1) & 7]) + 0x1234567u;    // 1 LOP3 + 1 IADD
            if (OP == 2) a[i] = a[i] + a[(i + 1) & 7];                   // 1 IADD
            if (OP == 3) a[i] = __reduce_min_sync(0xFFFFFFFFu, a[i]) + a[(i + 1) & 7];  // REDUX + IADD
            if (OP == 4) a[i] = __popc(a[i] ^ a[(i + 1) & 7]) + a[(i + 2) & 7];  // LOP3 + POPC + IADD
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int sms) {
    uint32_t* d;
    const int blocks = sms * 2, iters = 4096;
    cudaMalloc(&d, (size_t)blocks * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<OP><<<blocks, 1024>>>(d, iters, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, 1024>>>(d, iters, 2);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 1024 * iters * 8;
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %8.1f Gop/s  = %6.1f ops/clk/SM at %d MHz (max clock)\n", name, ms, ops / ms * 1e-6,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("popc+iadd", p.multiProcessorCount);
    run<1>("lop3+iadd", p.multiProcessorCount);
    run<2>("iadd", p.multiProcessorCount);
    run<3>("redux.min+iadd", p.multiProcessorCount);
    run<4>("xor+popc+iadd", p.multiProcessorCount);
    return 0;
}
