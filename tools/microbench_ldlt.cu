// microbench_ldlt.cu -- the reduced-system solve of csrc/ba2.cu in isolation (cycles of one warp / one CTA for n = 12, 54):
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -o /tmp/mbl tools/microbench_ldlt.cu && /tmp/mbl
#include <cstdio>
#include <vector>
#include "../ygz_slam_b200/csrc/ba2.cu"

namespace ygzb {
namespace {
__global__ void __launch_bounds__(256) solve_bench(const double* S0, const double* b0, int n, int reps, long long* cyc, double* out, int mode) {
    __shared__ double s_S[54 * 54], s_b[54], s_rd[96];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t_f = 0, t_s = 0;
    for (int r = 0; r < reps; ++r) {
        for (int i = tid; i < n * n; i += 256) s_S[i] = S0[i];
        if (tid < n) s_b[tid] = b0[tid];
        __syncthreads();
        const long long t0 = clock64();
        bool ok = true;
        if (mode == 1) {
            if (tid == 0) ok = lane_ldlt_solve<12>(s_S, s_b);
        } else if (n > 24) ok = ldlt6_factor<16, 16, 4, 4, true>(s_S, s_rd, n, tid);
        else if (warp == 0) ok = ldlt6_factor<4, 8, 5, 3, false>(s_S, s_rd, n, lane);
        const long long t1 = clock64();
        if (mode == 0 && warp == 0 && ok) warp_ldlt_subst(s_S, s_b, s_rd, n, lane);
        const long long t2 = clock64();
        __syncthreads();
        if (tid == 0) {
            t_f += t1 - t0;
            t_s += t2 - t1;
        }
    }
    if (tid == 0) {
        cyc[0] = t_f / reps;
        cyc[1] = t_s / reps;
        for (int i = 0; i < n; ++i) out[i] = s_b[i];
    }
}
}  // namespace
}  // namespace ygzb

int main() {
    for (int cfg : {0, 1, 2}) {
        const int n = cfg == 2 ? 54 : 12, mode = cfg == 1 ? 1 : 0;
        std::vector<double> S(n * n), b(n), M(n * n);
        unsigned s = 12345;
        auto rnd = [&] { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.0 - 0.5; };
        for (auto& v : M) v = rnd();
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double a = i == j ? n : 0.0;
                for (int k = 0; k < n; ++k) a += M[i * n + k] * M[j * n + k];
                S[i * n + j] = a;
            }
        for (auto& v : b) v = rnd();
        double *dS, *db, *dout;
        long long* dc;
        cudaMalloc(&dS, S.size() * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dout, n * 8); cudaMalloc(&dc, 16);
        cudaMemcpy(dS, S.data(), S.size() * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
        ygzb::solve_bench<<<1, 256>>>(dS, db, n, 50, dc, dout, mode);
        long long c[2];
        std::vector<double> x(n);
        cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
        cudaMemcpy(x.data(), dout, n * 8, cudaMemcpyDeviceToHost);
        double res = 0;
        for (int i = 0; i < n; ++i) {
            double r = -b[i];
            for (int j = 0; j < n; ++j) r += S[i * n + j] * x[j];
            res = fmax(res, fabs(r));
        }
        printf("%s n = %2d: factorisation %6lld cycles, substitution %6lld cycles, residual %.2e (%s)\n", mode ? "one lane, registers:" : "block LDL^T + warp substitution:", n, c[0], c[1], res, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

// stubs for the host helpers ba2.cu's launcher refers to (not used here)
namespace ygzb {
int check_cuda(ygzb_ctx*, cudaError_t e, const char*) { return e == cudaSuccess ? 0 : -1; }
int set_error(ygzb_ctx*, int code, const char*, ...) { return code; }
void prof_begin(ygzb_ctx*, int) {}
void prof_end(ygzb_ctx*) {}
void* dev_scratch(ygzb_ctx*, int, size_t) { return nullptr; }
}  // namespace ygzb
