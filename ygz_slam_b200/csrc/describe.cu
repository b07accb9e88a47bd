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
constexpr int kPatchPitch = 44;            // 11 words: a 4-byte aligned span of 44 B always covers the 39 B row, and an
                                           // odd word pitch spreads the rows of a column over all 32 banks
constexpr int kFeatPerWarp = 4;            // features per warp and CTA visit (amortises the pattern staging)

// IC_Angle weights.  Lane l < 31 owns row v = l - 15 of the radius-15 disc and reads it as the 8 words covering
// columns -15 .. 16; word k, byte j is column c = 4k + j - 15.  Row |v| is limited to |c| <= umax[|v|], so the
// column weights (c for m10, 1 for the row sum) are zeroed outside: tables [k][|v|] of packed s8x4 / u8x4.
struct IcTables {
    uint32_t wu[8 * 16];
    uint32_t w1[8 * 16];
};

constexpr IcTables make_ic_tables() {
    constexpr int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    IcTables t{};
    for (int k = 0; k < 8; ++k)
        for (int av = 0; av < 16; ++av) {
            uint32_t wu = 0, w1 = 0;
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * k + j - 15;
                const int ac = c < 0 ? -c : c;
                if (ac <= kUmax[av]) {
                    wu |= (uint32_t)(uint8_t)(int8_t)c << (8 * j);
                    w1 |= 1u << (8 * j);
                }
            }
            t.wu[k * 16 + av] = wu;
            t.w1[k * 16 + av] = w1;
        }
    return t;
}

__device__ const IcTables g_ic_tables = make_ic_tables();

struct DescribeSmem {
    float pat[1024];   // [test k of the byte][lane] as float4 (x0, y0, x1, y1)
    uint32_t ic_wu[8 * 16];
    uint32_t ic_w1[8 * 16];
    uint8_t patch[8][kPatchW * kPatchPitch];
};

__device__ __forceinline__ void stage_tables(DescribeSmem& sm) {
    // descriptor byte `lane` uses tests 8*lane .. 8*lane+7: store test (8*lane + k) at float4 index k*32 + lane
    for (int t = threadIdx.x; t < 1024; t += 256) {
        const int test = t >> 2, comp = t & 3;
        sm.pat[(((test & 7) * 32 + (test >> 3)) << 2) + comp] = (float)g_orb_pattern[t];
    }
    if (threadIdx.x < 128) {
        sm.ic_wu[threadIdx.x] = g_ic_tables.wu[threadIdx.x];
        sm.ic_w1[threadIdx.x] = g_ic_tables.w1[threadIdx.x];
    }
}

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}

// cvRound of a small f32 (|x| < 2^22) without the quarter-rate F2I: adding 1.5 * 2^23 leaves the integer, rounded
// to nearest-even, in the low mantissa bits
__device__ __forceinline__ int round_small(float x) { return __float_as_int(__fadd_rn(x, 12582912.f)) - 0x4B400000; }

// one warp: stage the (2*19+1)^2 neighbourhood of the centre in shared memory -- every entry is the
// LINEAR-address tap c + dy*w + dx, so row wrap / outside-buffer semantics are preserved -- then
// compute the angle (returned by all lanes) and descriptor byte `lane` from shared memory only.
__device__ __forceinline__ void describe_one(const LevelView& v, int cx, int cy, const DescribeSmem& sm,
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
    } else if (v.pitch == v.w) {
        // window reaches past the first / last byte of the level buffer (corners of the upper levels that InFrame lets
        // through, SURVEY 8a hazard 4): the level is contiguous, so a linear address is a memory offset -- byte loads
        // with a range check, no division
        for (int r = 0; r < kPatchW; ++r) {
            const int lin0 = lin_first + r * v.w;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int cc = lane + 32 * h, lin = lin0 + cc;
                if (cc < kPatchW) s_patch[r * kPatchPitch + cc] = (lin >= 0 && lin < n_px) ? v.img[lin] : (uint8_t)0;
            }
        }
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

    // IC_Angle: m10 = sum u * I(u, v), m01 = sum v * I(u, v) over the disc.  Lane <-> row v = lane - 15; the row is
    // read as 9 aligned words, shifted to start at column -15, and reduced with two byte dot products per word
    // (integer sums: any order gives the reference's result).
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int vr = lane - 15, av = vr < 0 ? -vr : vr;
        const int o = ctr_off + vr * kPatchPitch - 15;          // byte offset of column -15 of this row
        const uint32_t* w = reinterpret_cast<const uint32_t*>(s_patch) + (o >> 2);
        const int sh = (o & 3) * 8;
        int rowsum = 0;
        uint32_t lo = w[0];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t hi = w[k + 1];
            const uint32_t px = __funnelshift_r(lo, hi, sh);
            m10 = dp4a_us(px, sm.ic_wu[k * 16 + av], m10);
            rowsum = dp4a_us(px, sm.ic_w1[k * 16 + av], rowsum);
            lo = hi;
        }
        m01 = vr * rowsum;
    }
    m10 = __reduce_add_sync(0xFFFFFFFFu, m10);
    m01 = __reduce_add_sync(0xFFFFFFFFu, m01);
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    *angle_out = angle;

    // ComputeOrbDescriptor: a = cos, b = sin of angle * (float)(CV_PI/180.f); one f64 sincos on one lane (the FP64
    // pipe is narrow), broadcast by shuffle
    constexpr float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    float cs_c = 0.f, cs_s = 0.f;
    if (lane == 0) {
        double sd, cd;
        sincos((double)rad, &sd, &cd);
        cs_c = (float)cd;
        cs_s = (float)sd;
    }
    const float a = __shfl_sync(0xFFFFFFFFu, cs_c, 0), b = __shfl_sync(0xFFFFFFFFu, cs_s, 0);
    const float4* pat = reinterpret_cast<const float4*>(sm.pat) + lane;   // [test k][lane]: conflict-free 128-bit reads
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 pp = pat[32 * k];  // pattern converted to f32 once per CTA
        const float x0 = pp.x, y0 = pp.y, x1 = pp.z, y1 = pp.w;
        const int r0 = round_small(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = round_small(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = round_small(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = round_small(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        // |r|, |c| <= 19 for every pattern point (|p| <= 13*sqrt(2) < 19), so the staged window covers it
        const int t0 = ctr[r0 * kPatchPitch + c0], t1 = ctr[r1 * kPatchPitch + c1];
        val |= (t0 < t1) << k;
    }
    *byte_out = (uint8_t)val;
    __syncwarp();
}

// features of the slot store (written by merge_cells_kernel): level coordinates are integers
template <int kFPW>   // features per warp: 4 amortises the table staging over big batches, 1 keeps a small batch (the key-frames of
                      // one tracking round) one feature deep
__global__ void __launch_bounds__(256) describe_store_kernel(const uint8_t* __restrict__ pyr, size_t slot_stride,
                                                             const int32_t* __restrict__ slots, Geometry g,
                                                             const int32_t* __restrict__ count,
                                                             const int16_t* __restrict__ fx, const int16_t* __restrict__ fy,
                                                             const uint8_t* __restrict__ flevel, float* __restrict__ fangle,
                                                             uint8_t* __restrict__ fdesc) {
    __shared__ __align__(16) DescribeSmem sm;
    const int slot = slots[blockIdx.y];
    const int n = count[slot];
    if (blockIdx.x * (8 * kFPW) >= n) return;  // most CTAs of the capacity-sized grid have no feature
    stage_tables(sm);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < kFPW; ++r) {
        const int i = blockIdx.x * (8 * kFPW) + r * 8 + warp;
        if (i >= n) return;
        const size_t o = (size_t)slot * g.n_cells + i;
        const int L = flevel[o];
        LevelView v{pyr + (size_t)slot * slot_stride + g.lv[L].off, g.lv[L].w, g.lv[L].h, g.lv[L].pitch};
        float angle;
        uint8_t byte;
        describe_one(v, fx[o], fy[o], sm, sm.patch[warp], lane, &angle, &byte);
        fdesc[o * 32 + lane] = byte;
        if (lane == 0) fangle[o] = angle;
    }
}

// caller-supplied features (ComputeAngleAndDescriptor): full-res double pixels, centre = cvRound(px / 2^L)
__global__ void __launch_bounds__(256) describe_list_kernel(const uint8_t* __restrict__ pyr, size_t slot_stride, Geometry g,
                                                            const int32_t* __restrict__ slot_of, int total,
                                                            const double* __restrict__ px, const double* __restrict__ py,
                                                            const uint8_t* __restrict__ level, float* __restrict__ angle_out,
                                                            uint8_t* __restrict__ desc_out) {
    __shared__ __align__(16) DescribeSmem sm;
    stage_tables(sm);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < kFeatPerWarp; ++r) {
        const int i = blockIdx.x * (8 * kFeatPerWarp) + r * 8 + warp;
        if (i >= total) return;
        const int L = level[i];
        const int slot = slot_of[i];
        LevelView v{pyr + (size_t)slot * slot_stride + g.lv[L].off, g.lv[L].w, g.lv[L].h, g.lv[L].pitch};
        const double scale = (double)(1 << L);
        const int cx = __double2int_rn(px[i] / scale), cy = __double2int_rn(py[i] / scale);
        float angle;
        uint8_t byte;
        describe_one(v, cx, cy, sm, sm.patch[warp], lane, &angle, &byte);
        desc_out[(size_t)i * 32 + lane] = byte;
        if (lane == 0) angle_out[i] = angle;
    }
}

}  // namespace

int launch_describe_store(ygzb_frames* f, int n) {
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    ProfScope ps(ctx, kStageDescribe);
    if (n <= 16) {   // a few frames: the launch is one partial wave either way, so keep every warp one feature deep
        dim3 grid((g.n_cells + 7) / 8, n);
        describe_store_kernel<1><<<grid, 256, 0, ctx->stream>>>(f->d_pyr, ctx->slot_stride, f->d_slots, g, f->d_count, f->d_fx, f->d_fy, f->d_flevel,
                                                                f->d_fangle, f->d_fdesc);
    } else {
        dim3 grid((g.n_cells + 8 * kFeatPerWarp - 1) / (8 * kFeatPerWarp), n);
        describe_store_kernel<kFeatPerWarp><<<grid, 256, 0, ctx->stream>>>(f->d_pyr, ctx->slot_stride, f->d_slots, g, f->d_count, f->d_fx, f->d_fy,
                                                                           f->d_flevel, f->d_fangle, f->d_fdesc);
    }
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

int launch_describe_list(ygzb_frames* f, int n, const int32_t* d_slot_of, int total, const double* d_x, const double* d_y,
                         const uint8_t* d_level, float* d_angle, uint8_t* d_desc) {
    (void)n;
    ygzb_ctx* ctx = f->ctx;
    if (total <= 0) return YGZB_OK;
    ProfScope ps(ctx, kStageDescribe);
    describe_list_kernel<<<(total + 8 * kFeatPerWarp - 1) / (8 * kFeatPerWarp), 256, 0, ctx->stream>>>(f->d_pyr, ctx->slot_stride, ctx->geo, d_slot_of, total,
                                                                   d_x, d_y, d_level, d_angle, d_desc);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

}  // namespace ygzb
