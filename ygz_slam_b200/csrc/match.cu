// match.cu -- 256-bit Hamming distance and brute-force (cross-checked) matching.
//
// Replaces:
//   Matcher::DescriptorDistance                        reference src/Algorithm/Matcher.cpp:30-43
//   cv::BFMatcher(cv::NORM_HAMMING, true).match        reference test/test_orb_match.cpp:87-92
//   distance loop of Matcher::CheckFrameDescriptors    reference src/Algorithm/Matcher.cpp:52-59
//
// One thread owns one query descriptor (8 x u32 in registers); the train descriptors stream through
// shared memory in 8 KB chunks and are read as warp-wide broadcasts.  Per (query, train) pair:
// 8 XOR + 8 POPC + adds.  Both match directions come out of ONE pass over the distance matrix:
//   forward  : per-thread min of (dist << 16 | train)   -> first minimum over the train rows
//   backward : warp-reduce (redux.sync min) of (dist << 16 | query) per train row, combined across
//              warps in shared memory and across CTAs with one global atomicMin per (CTA, train row)
//              -> first minimum over the query rows, which is what crossCheck compares against.
// Roofline: 32*(nA+nB) + 8*nA algorithmic bytes per pair against nA*nB*8 POPC -- the kernel is bound
// by the integer POPC pipe, not by HBM; bench.py reports both fractions.
#include <cstdlib>

#include "common.cuh"

namespace ygzb {

namespace {

constexpr int kQueriesPerCta = 128;
constexpr int kChunk = 256;  // train rows per shared-memory chunk (8 KB)
constexpr int kWarps = kQueriesPerCta / 32;

struct MatchArgs {
    const uint8_t* base;       // descriptor sets: set s starts at base + s * set_stride
    size_t set_stride;
    const int32_t* counts;     // rows in set s
    const int32_t* a_sets;     // per pair: query set
    const int32_t* b_sets;     // per pair: train set
    int cap;                   // result stride per pair (max rows per set)
    unsigned* fwd_key;         // [n_pairs][cap]  (dist << 16 | train) or 0xFFFFFFFF
    unsigned* col_key;         // [n_pairs][cap]  (dist << 16 | query), pre-set to 0xFFFFFFFF
};

// 256-bit Hamming distance.  POPC issues at 16 lanes/clk/SM on sm_100 (measured, tools/microbench.cu)
// against 64 for LOP3, so three carry-save adder steps fold the 8 XOR words into 2 weight-1 and 3
// weight-2 words: 5 POPC + 14 LOP3 instead of 8 POPC + 8 LOP3 balances the XU and ALU pipes.
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& sum, uint32_t& carry) {
    sum = a ^ b ^ c;
    carry = (a & b) | (a & c) | (b & c);
}

template <int kPopc = 5>
__device__ __forceinline__ unsigned hamming256(const uint32_t (&q)[8], const uint4 lo, const uint4 hi) {
    const uint32_t x0 = q[0] ^ lo.x, x1 = q[1] ^ lo.y, x2 = q[2] ^ lo.z, x3 = q[3] ^ lo.w;
    const uint32_t x4 = q[4] ^ hi.x, x5 = q[5] ^ hi.y, x6 = q[6] ^ hi.z, x7 = q[7] ^ hi.w;
    uint32_t s0, c0, s1, c1, s2, c2;
    csa(x0, x1, x2, s0, c0);
    csa(x3, x4, x5, s1, c1);
    csa(s0, s1, x6, s2, c2);
    if (kPopc == 4) {
        // one more adder level folds the three weight-2 words into a weight-2 and a weight-4 word: 4 POPC + 16 LOP3
        uint32_t t, f;
        csa(c0, c1, c2, t, f);
        return __popc(s2) + __popc(x7) + 2 * __popc(t) + 4 * __popc(f);
    }
    return __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
}

template <bool kCross, int kPopc>
__global__ void __launch_bounds__(kQueriesPerCta) match_kernel(const MatchArgs a) {
    __shared__ __align__(16) uint4 s_train[kChunk * 2];
    __shared__ unsigned s_col[kCross ? kWarps * kChunk : 1];

    const int pair = blockIdx.y;
    const int sa = a.a_sets[pair], sb = a.b_sets[pair];
    const int nA = a.counts[sa], nB = a.counts[sb];
    const int q0 = blockIdx.x * kQueriesPerCta;
    if (q0 >= nA) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qi = q0 + tid;
    const bool active = qi < nA;

    uint32_t q[8];
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.base + (size_t)sa * a.set_stride) + (size_t)(active ? qi : q0) * 2;
        const uint4 lo = src[0], hi = src[1];
        q[0] = lo.x; q[1] = lo.y; q[2] = lo.z; q[3] = lo.w;
        q[4] = hi.x; q[5] = hi.y; q[6] = hi.z; q[7] = hi.w;
    }
    const uint4* train = reinterpret_cast<const uint4*>(a.base + (size_t)sb * a.set_stride);
    unsigned best = 0xFFFFFFFFu;
    // lanes past the last query shadow query q0 under the largest index: their column keys tie with q0's own lane
    // and lose on the index, so the inner loop needs no select (set sizes stay below 65535)
    const unsigned qkey = active ? (unsigned)qi : 0xFFFFu;

    for (int j0 = 0; j0 < nB; j0 += kChunk) {
        const int nj = min(kChunk, nB - j0);
        __syncthreads();  // previous chunk fully consumed
        for (int i = tid; i < nj * 2; i += kQueriesPerCta) s_train[i] = train[(size_t)j0 * 2 + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < nj; ++j) {
            const unsigned d = hamming256<kPopc>(q, s_train[2 * j], s_train[2 * j + 1]);
            best = min(best, d * 65536u + (unsigned)(j0 + j));
            if (kCross) {
                const unsigned m = __reduce_min_sync(0xFFFFFFFFu, d * 65536u + qkey);
                if (lane == 0) s_col[warp * kChunk + j] = m;
            }
        }
        if (kCross) {
            __syncthreads();
            for (int j = tid; j < nj; j += kQueriesPerCta) {
                unsigned m = s_col[j];
#pragma unroll
                for (int w = 1; w < kWarps; ++w) m = min(m, s_col[w * kChunk + j]);
                atomicMin(&a.col_key[(size_t)pair * a.cap + j0 + j], m);
            }
        }
    }
    if (active) a.fwd_key[(size_t)pair * a.cap + qi] = best;
}

// decode forward keys, apply the cross check, write results packed per pair
__global__ void match_finalize_kernel(const MatchArgs a, int cross_check, const int32_t* __restrict__ q_offsets,
                                      int32_t* __restrict__ train_idx, int32_t* __restrict__ dist) {
    const int pair = blockIdx.y;
    const int nA = a.counts[a.a_sets[pair]];
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nA) return;
    const unsigned k = a.fwd_key[(size_t)pair * a.cap + qi];
    int j = -1, d = -1;
    if (k != 0xFFFFFFFFu) {
        j = (int)(k & 0xFFFFu);
        d = (int)(k >> 16);
        if (cross_check && (int)(a.col_key[(size_t)pair * a.cap + j] & 0xFFFFu) != qi) j = d = -1;
    }
    const size_t o = (size_t)q_offsets[pair] + qi;
    train_idx[o] = j;
    dist[o] = d;
}

__global__ void hamming_pairs_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B,
                                     const int32_t* __restrict__ ia, const int32_t* __restrict__ ib, int n,
                                     int32_t* __restrict__ dist) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint4* pa = reinterpret_cast<const uint4*>(A) + (size_t)ia[k] * 2;
    const uint4* pb = reinterpret_cast<const uint4*>(B) + (size_t)ib[k] * 2;
    const uint4 a0 = pa[0], a1 = pa[1];
    uint32_t q[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    dist[k] = (int32_t)hamming256<5>(q, pb[0], pb[1]);
}

// exclusive scan of counts[sets[i]] over the pairs/items of a call (n <= a few thousand): offsets[n] = total
__global__ void offsets_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ sets, int n,
                               int32_t* __restrict__ offsets) {
    __shared__ int s_carry;
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_carry = 0;
        offsets[0] = 0;
    }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const int v = i < n ? counts[sets[i]] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xFFFFFFFFu, w, o);
                if (lane >= o) w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int total_incl = s_carry + (warp ? s_warp[warp - 1] : 0) + incl;
        if (i < n) offsets[i + 1] = total_incl;
        __syncthreads();
        if (tid == 1023) s_carry = total_incl;
        __syncthreads();
    }
}

}  // namespace

int launch_offsets(ygzb_ctx* ctx, const int32_t* d_counts, const int32_t* d_sets, int n, int32_t* d_offsets) {
    ProfScope ps(ctx, kStagePack);
    offsets_kernel<<<1, 1024, 0, ctx->stream>>>(d_counts, d_sets, n, d_offsets);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

int launch_match(ygzb_ctx* ctx, const uint8_t* d_base, size_t set_stride, const int32_t* d_counts, const int32_t* d_a_sets,
                 const int32_t* d_b_sets, int n_pairs, int cap, int cross_check, unsigned* d_fwd_key, unsigned* d_col_key,
                 const int32_t* d_q_offsets, int32_t* d_train_idx, int32_t* d_dist) {
    if (n_pairs <= 0) return YGZB_OK;
    MatchArgs a{d_base, set_stride, d_counts, d_a_sets, d_b_sets, cap, d_fwd_key, d_col_key};
    if (cross_check)
        YGZB_CUDA(ctx, cudaMemsetAsync(d_col_key, 0xFF, (size_t)n_pairs * cap * sizeof(unsigned), ctx->stream));
    dim3 grid((cap + kQueriesPerCta - 1) / kQueriesPerCta, n_pairs);
    {
        ProfScope ps(ctx, kStageMatch);
        // YGZB_MATCH_POPC=4|5 picks the adder-tree depth of the distance (tuning knob, identical results)
        static const int popc_variant = [] {
            const char* e = getenv("YGZB_MATCH_POPC");
            return e && e[0] == '4' ? 4 : 5;
        }();
        if (cross_check) {
            if (popc_variant == 4) match_kernel<true, 4><<<grid, kQueriesPerCta, 0, ctx->stream>>>(a);
            else match_kernel<true, 5><<<grid, kQueriesPerCta, 0, ctx->stream>>>(a);
        } else {
            if (popc_variant == 4) match_kernel<false, 4><<<grid, kQueriesPerCta, 0, ctx->stream>>>(a);
            else match_kernel<false, 5><<<grid, kQueriesPerCta, 0, ctx->stream>>>(a);
        }
    }
    YGZB_LAUNCHED(ctx);
    ProfScope ps(ctx, kStageMatchFinalize);
    dim3 fgrid((cap + 255) / 256, n_pairs);
    match_finalize_kernel<<<fgrid, 256, 0, ctx->stream>>>(a, cross_check, d_q_offsets, d_train_idx, d_dist);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

int launch_hamming_pairs(ygzb_ctx* ctx, const uint8_t* d_A, const uint8_t* d_B, const int32_t* d_ia, const int32_t* d_ib,
                         int n, int32_t* d_dist) {
    if (n <= 0) return YGZB_OK;
    hamming_pairs_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_A, d_B, d_ia, d_ib, n, d_dist);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

}  // namespace ygzb
