// describe.cu -- IC_Angle + rotated-BRIEF (ORB) descriptor, one warp per feature.
//
// Replaces FeatureDetector::IC_Angle (reference src/Algorithm/FeatureDetector.cpp:509-537),
// FeatureDetector::ComputeOrbDescriptor (:539-578), the per-feature tail of Detect (:431-441) and
// ComputeAngleAndDescriptor (:580-588); cv::fastAtan2 / cvRound are restated from OpenCV.
//
// Parity decisions shared with the oracle (oracle/detector.cpp):
//   * _umax = canonical ORB table (the reference's init loop reads an unfilled vector, :304-322);
//   * taps are LINEAR addresses on the continuous level image (row wrap kept, outside-buffer = 0);
//   * cos/sin: correctly rounded f32 (f64 evaluation rounded once) -- libm's cosf is glibc-version
//     dependent in the last bit;
//   * every f32 operation is an explicit round-to-nearest intrinsic: no FMA contraction.
// Roofline class: L2-resident gather (<= 1 kB per feature); HBM traffic is bounded by the level
// images already read by fast.cu.
#include "common.cuh"

namespace ygzb {

namespace {

// global (not __constant__): the table is copied to shared memory with one coalesced word per thread
__device__ const int8_t g_orb_pattern[1024] = {
#include "orb_pattern.inc"
};

struct LevelView {
    const uint8_t* img;
    int w, h, pitch;
};

__device__ __forceinline__ int tap(const LevelView& v, int lin) {
    if (lin < 0 || lin >= v.w * v.h) return 0;
    const int r = lin / v.w;
    return v.img[(size_t)r * v.pitch + (lin - r * v.w)];
}

// cv::fastAtan2 scalar path (degrees)
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    constexpr float k = (float)(180 / 3.14159265358979323846);
    constexpr float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k, p5 = 0.1555786518463281f * k,
                    p7 = -0.04432655554792128f * k;
    constexpr float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int kPatchR = 19;                 // EDGE_THRESHOLD: rotated pattern taps stay within +-19 px
constexpr int kPatchW = 2 * kPatchR + 1;    // 39
constexpr int kPatchPitch = 48;            // 12 words: a 4-byte aligned span of 44 B always covers the 39 B row

// one warp: stage the (2*19+1)^2 neighbourhood of the centre in shared memory -- every entry is the
// LINEAR-address tap c + dy*w + dx, so row wrap / outside-buffer semantics are preserved -- then
// compute the angle (returned by all lanes) and descriptor byte `lane` from shared memory only.
__device__ __forceinline__ void describe_one(const LevelView& v, int cx, int cy, const float* __restrict__ s_pat,
                                             uint8_t* __restrict__ s_patch, int lane, float* angle_out,
                                             uint8_t* byte_out) {
    const int c = cy * v.w + cx;
    const int n_px = v.w * v.h;
    const int lin_first = c - kPatchR * v.w - kPatchR;   // linear address of the window's first byte
    const int align = lin_first & 3;
    int ctr_off = kPatchR * kPatchPitch + kPatchR;
    // fast path (levels 0-2 of a 640x480 frame away from the buffer ends): rows are 4-byte congruent, so the
    // window is fetched as 39 x 11 aligned words (14 word loads per lane instead of 78 byte loads)
    if (v.pitch == v.w && (v.w & 3) == 0 && lin_first - align >= 0 &&
        lin_first - align + (kPatchW - 1) * v.w + 44 <= n_px) {
        const uint8_t* src = v.img + (lin_first - align);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_patch);
        // lanes 0-10 fetch the 11 words of an even row, lanes 11-21 those of the next row: 20 passes, one
        // address increment per pass
        if (lane < 22) {
            const int half = lane >= 11, k = lane - 11 * half;
            const uint8_t* p = src + (size_t)half * v.w + 4 * k;
            uint32_t* d = dst + half * (kPatchPitch / 4) + k;
            const size_t step = 2 * (size_t)v.w;
#pragma unroll
            for (int t = 0; t < 19; ++t, p += step, d += 2 * (kPatchPitch / 4)) *d = *reinterpret_cast<const uint32_t*>(p);
            if (!half) *d = *reinterpret_cast<const uint32_t*>(p);  // row 38
        }
        ctr_off += align;
    } else {
        for (int r = 0; r < kPatchW; ++r) {
            const int lin0 = lin_first + r * v.w;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cc = lane + 32 * h;
                if (cc < kPatchW) s_patch[r * kPatchPitch + cc] = (uint8_t)tap(v, lin0 + cc);
            }
        }
    }
    __syncwarp();
    const uint8_t* ctr = s_patch + ctr_off;

    // IC_Angle: lane <-> column u = lane - 15 of the radius-15 disc
    constexpr int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int u = lane - 15;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        m10 = u * (int)ctr[u];
        const int au = u < 0 ? -u : u;
#pragma unroll
        for (int vv = 1; vv <= 15; ++vv) {
            if (au <= kUmax[vv]) {
                const int plus = ctr[vv * kPatchPitch + u], minus = ctr[-vv * kPatchPitch + u];
                m01 += vv * (plus - minus);
                m10 += u * (plus + minus);
            }
        }
    }
    m10 = __reduce_add_sync(0xFFFFFFFFu, m10);
    m01 = __reduce_add_sync(0xFFFFFFFFu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    *angle_out = angle;

    // ComputeOrbDescriptor: a = cos, b = sin of angle * (float)(CV_PI/180.f); f64 evaluation on two
    // lanes only (the FP64 pipe is narrow), broadcast by shuffle
    constexpr float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    float cs = 0.f;
    if (lane == 0) cs = (float)cos((double)rad);
    if (lane == 1) cs = (float)sin((double)rad);
    const float a = __shfl_sync(0xFFFFFFFFu, cs, 0), b = __shfl_sync(0xFFFFFFFFu, cs, 1);
    const float4* pat = reinterpret_cast<const float4*>(s_pat) + lane;   // [test k][lane]: conflict-free 128-bit reads
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 pp = pat[32 * k];  // pattern converted to f32 once per CTA (I2F is an XU op)
        const float x0 = pp.x, y0 = pp.y, x1 = pp.z, y1 = pp.w;
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        // |r|, |c| <= 19 for every pattern point (|p| <= 13*sqrt(2) < 19), so the staged window covers it
        const int t0 = ctr[r0 * kPatchPitch + c0], t1 = ctr[r1 * kPatchPitch + c1];
        val |= (t0 < t1) << k;
    }
    *byte_out = (uint8_t)val;
    __syncwarp();
}

// features of the slot store (written by merge_cells_kernel): level coordinates are integers
__global__ void __launch_bounds__(256) describe_store_kernel(const uint8_t* __restrict__ pyr, size_t slot_stride,
                                                             const int32_t* __restrict__ slots, Geometry g,
                                                             const int32_t* __restrict__ count,
                                                             const int16_t* __restrict__ fx, const int16_t* __restrict__ fy,
                                                             const uint8_t* __restrict__ flevel, float* __restrict__ fangle,
                                                             uint8_t* __restrict__ fdesc) {
    __shared__ __align__(16) float s_pat[1024];
    __shared__ __align__(16) uint8_t s_patch[8][kPatchW * kPatchPitch];
    const int slot = slots[blockIdx.y];
    if (blockIdx.x * 8 >= count[slot]) return;  // most CTAs of the capacity-sized grid have no feature
    // descriptor byte `lane` uses tests 8*lane .. 8*lane+7: store test (8*lane + k) at float4 index k*32 + lane
    for (int t = threadIdx.x; t < 1024; t += 256) {
        const int test = t >> 2, comp = t & 3;
        s_pat[(((test & 7) * 32 + (test >> 3)) << 2) + comp] = (float)g_orb_pattern[t];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= count[slot]) return;
    const size_t o = (size_t)slot * g.n_cells + i;
    const int L = flevel[o];
    LevelView v{pyr + (size_t)slot * slot_stride + g.lv[L].off, g.lv[L].w, g.lv[L].h, g.lv[L].pitch};
    float angle;
    uint8_t byte;
    describe_one(v, fx[o], fy[o], s_pat, s_patch[threadIdx.x >> 5], lane, &angle, &byte);
    fdesc[o * 32 + lane] = byte;
    if (lane == 0) fangle[o] = angle;
}

// caller-supplied features (ComputeAngleAndDescriptor): full-res double pixels, centre = cvRound(px / 2^L)
__global__ void __launch_bounds__(256) describe_list_kernel(const uint8_t* __restrict__ pyr, size_t slot_stride, Geometry g,
                                                            const int32_t* __restrict__ slot_of, int total,
                                                            const double* __restrict__ px, const double* __restrict__ py,
                                                            const uint8_t* __restrict__ level, float* __restrict__ angle_out,
                                                            uint8_t* __restrict__ desc_out) {
    __shared__ __align__(16) float s_pat[1024];
    __shared__ __align__(16) uint8_t s_patch[8][kPatchW * kPatchPitch];
    // descriptor byte `lane` uses tests 8*lane .. 8*lane+7: store test (8*lane + k) at float4 index k*32 + lane
    for (int t = threadIdx.x; t < 1024; t += 256) {
        const int test = t >> 2, comp = t & 3;
        s_pat[(((test & 7) * 32 + (test >> 3)) << 2) + comp] = (float)g_orb_pattern[t];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= total) return;
    const int L = level[i];
    const int slot = slot_of[i];
    LevelView v{pyr + (size_t)slot * slot_stride + g.lv[L].off, g.lv[L].w, g.lv[L].h, g.lv[L].pitch};
    const double scale = (double)(1 << L);
    const int cx = __double2int_rn(px[i] / scale), cy = __double2int_rn(py[i] / scale);
    float angle;
    uint8_t byte;
    describe_one(v, cx, cy, s_pat, s_patch[threadIdx.x >> 5], lane, &angle, &byte);
    desc_out[(size_t)i * 32 + lane] = byte;
    if (lane == 0) angle_out[i] = angle;
}

}  // namespace

int launch_describe_store(ygzb_frames* f, int n) {
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    dim3 grid((g.n_cells + 7) / 8, n);
    ProfScope ps(ctx, kStageDescribe);
    describe_store_kernel<<<grid, 256, 0, ctx->stream>>>(f->d_pyr, ctx->slot_stride, f->d_slots, g, f->d_count, f->d_fx,
                                                         f->d_fy, f->d_flevel, f->d_fangle, f->d_fdesc);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

int launch_describe_list(ygzb_frames* f, int n, const int32_t* d_slot_of, int total, const double* d_x, const double* d_y,
                         const uint8_t* d_level, float* d_angle, uint8_t* d_desc) {
    (void)n;
    ygzb_ctx* ctx = f->ctx;
    if (total <= 0) return YGZB_OK;
    ProfScope ps(ctx, kStageDescribe);
    describe_list_kernel<<<(total + 7) / 8, 256, 0, ctx->stream>>>(f->d_pyr, ctx->slot_stride, ctx->geo, d_slot_of, total,
                                                                   d_x, d_y, d_level, d_angle, d_desc);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

}  // namespace ygzb
