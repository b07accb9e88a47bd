#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <vector>
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BW, int BH>
__global__ void k(const CUtensorMap* __restrict__ tm, int cx, int cy, int cz, uint8_t* out) {
    __shared__ __align__(128) uint8_t s[BW * BH];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(BW * BH)) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(s)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(cx), "r"(cy), "r"(cz), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    }
    for (int i = tid; i < BW * BH; i += blockDim.x) out[i] = s[i];
}
template <int BW, int BH>
int run(EncodeTiled enc, uint8_t* d_img, int w, int h, int pitch, size_t slot_stride, int cap, int cx, int cy, int cz, const std::vector<uint8_t>& img) {
    cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)cap};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)slot_stride};
    cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d encode rc=%d ", BW, BH, (int)r);
    if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    uint8_t* d_out; cudaMalloc(&d_out, BW * BH);
    k<BW, BH><<<1, 128>>>(dm, cx, cy, cz, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 2; }
    std::vector<uint8_t> out(BW * BH); cudaMemcpy(out.data(), d_out, BW * BH, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < BH; ++y) for (int x = 0; x < BW; ++x) {
        int gx = cx + x, gy = cy + y; uint8_t want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? img[(size_t)cz * slot_stride + (size_t)gy * pitch + gx] : 0;
        bad += out[y * BW + x] != want;
    }
    printf("mismatches %d\n", bad);
    return 0;
}
#include <stdlib.h>
int main(int argc, char** argv) {
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) { printf("no entry\n"); return 1; }
    EncodeTiled enc = (EncodeTiled)fn;
    const int w = 640, h = 480, pitch = 640, cap = 3; const size_t ss = 403200;
    std::vector<uint8_t> img(ss * cap); for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t* d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    if (which == 0) run<64, 32>(enc, d, w, h, pitch, ss, cap, 64, 32, 1, img);
    if (which == 1) run<96, 32>(enc, d, w, h, pitch, ss, cap, 312, 115, 1, img);
    if (which == 2) run<64, 50>(enc, d, w, h, pitch, ss, cap, 312, 115, 1, img);
    if (which == 3) run<128, 50>(enc, d, w, h, pitch, ss, cap, 312, 115, 1, img);
    if (which == 4) run<128, 64>(enc, d, w, h, pitch, ss, cap, -8, -5, 0, img);
    if (which == 5) run<96, 64>(enc, d, w, h, pitch, ss, cap, 312, 115, 2, img);
    if (which == 6) run<112, 50>(enc, d, w, h, pitch, ss, cap, 304, 115, 2, img);
    if (which == 7) run<112, 50>(enc, d, w, h, pitch, ss, cap, -16, -5, 1, img);
    if (which == 8) run<112, 50>(enc, d, w, h, pitch, ss, cap, 624, 435, 0, img);
    return 0;
}
