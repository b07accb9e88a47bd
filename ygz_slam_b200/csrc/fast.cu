// fast.cu -- FAST-10 detect + score + 3x3 non-max + per-grid-cell Shi-Tomasi selection, fused.
//
// Replaces the per-level body of FeatureDetector::Detect (reference
// src/Algorithm/FeatureDetector.cpp:362-427):
//     fast::fast_corner_detect_10_sse2 / fast_corner_score_10 / fast_nonmax_3x3     (:365-381)
//     frame->InFrame(px, 20, L)  (include/ygz/Basic/Frame.h:67-71, double-scaled)   (:386)
//     cell index, _old_features skip, ShiTomasiScore, keep the strictly best per cell (:391-425)
// and FeatureDetector::ShiTomasiScore (:467-507).
//
// One CTA = one 80x40 tile of one pyramid level of one frame; ALL levels of ALL frames of a batch go
// in a single launch (blockIdx.x walks the tiles of every level, blockIdx.y the frames).  The tile
// (+5 px halo) is staged in shared memory once; every later read (16-px ring, 3x3 score
// neighbourhood, 10x10 Shi-Tomasi window) hits shared memory, so HBM sees each pyramid byte once
// (+ halo): algorithmic traffic = 409,600 B per 8-level frame.
//
// Ordering semantics without an ordered corner list: the reference walks nonmax corners level by
// level in raster order and replaces a cell's candidate only on a STRICTLY larger score.  That is
// "max score, ties -> earliest (level, raster)", with one twist: a NaN score on the very first
// candidate of a cell sticks forever (nothing compares greater than NaN), a NaN later never wins.
// Each tile owns whole cells, so it reduces candidates with two 64-bit shared-memory atomics
//     first = min (raster << 32 | score bits)          -> the earliest candidate and its score
//     best  = max (orderable(score) << 32 | ~raster)   -> the best non-NaN score, earliest on ties
// and merge_cells_kernel replays the level order.  The f32 maths is written with the explicit
// round-to-nearest intrinsics (__fmul_rn / __fadd_rn ...), which ptxas never contracts into FMA; the
// file does not depend on -fmad=false (build.py passes that to align.cu, track.cu, initializer.cu).
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace ygzb {

namespace {

constexpr int kSmW = 112;           // smem tile pitch : x in [x0-16, x0+96); the TMA needs a 16-byte aligned start column
constexpr int kSmX = 15;            // byte offset of region column rx = 0 (pixel x0 - 1) inside a smem row
constexpr int kSmH = kTileH + 10;   // 50 rows         : y in [y0-5, y0+45)
constexpr int kScW = kTileW + 2;    // score region 82 x 42 : tile + 1 px ring
constexpr int kScH = kTileH + 2;
constexpr int kScPitch = 84;
constexpr int kMaxTileCells = 512;
constexpr unsigned long long kEmptyFirst = ~0ull;

// TMA descriptors of the levels whose tiles are fetched with cp.async.bulk.tensor (kernel parameter space)
struct TileMaps {
    CUtensorMap m[3];
};

struct DetectArgs {
    const uint8_t* pyr;
    size_t slot_stride;
    const int32_t* slots;
    const uint8_t* occupied;  // [n][n_cells] or null
    unsigned long long* best_key;
    unsigned long long* first_key;
    int32_t* stats;
    Geometry g;
    int tma_levels;  // levels [0, tma_levels) stage their tile through the TMA
};

__device__ __forceinline__ bool has_run10(unsigned m16) {
    unsigned x = m16 | (m16 << 16);
    unsigned r2 = x & (x >> 1);
    unsigned r4 = r2 & (r2 >> 2);
    unsigned r8 = r4 & (r4 >> 4);
    unsigned r10 = r8 & (r2 >> 8);
    return (r10 & 0xFFFFu) != 0;
}

// Bresenham ring of radius 3 in circular order: RDX/RDY below (dx,dy), index 0 = (0,3)
// Cheap necessary condition: a 10-arc contains at least one pixel of EVERY opposite pair, so a bright
// (dark) corner needs a bright (dark) pixel in each of the pairs (0,8) (4,12) (2,10) (6,14).
template <int PITCH>
__device__ __forceinline__ bool fast_quick_test(const uint8_t* __restrict__ c, int b) {
    const int p = *c;
    const int cb = p + b, c_b = p - b;
    // pairs (0,8) and (4,12) evaluated together (five independent loads, one branch)
    const int v0 = c[3 * PITCH], v8 = c[-3 * PITCH], v4 = c[3], v12 = c[-3];
    bool pb = ((v0 > cb) | (v8 > cb)) & ((v4 > cb) | (v12 > cb));
    bool pd = ((v0 < c_b) | (v8 < c_b)) & ((v4 < c_b) | (v12 < c_b));
    if (!(pb | pd)) return false;
    const int v2 = c[2 * PITCH + 2], v10 = c[-2 * PITCH - 2], v6 = c[-2 * PITCH + 2], v14 = c[2 * PITCH - 2];
    pb &= ((v2 > cb) | (v10 > cb)) & ((v6 > cb) | (v14 > cb));
    pd &= ((v2 < c_b) | (v10 < c_b)) & ((v6 < c_b) | (v14 < c_b));
    return pb | pd;
}

// full segment test: >= 10 contiguous ring pixels all > p+b or all < p-b (strict)
template <int PITCH>
__device__ __forceinline__ bool fast_full_test(const uint8_t* __restrict__ c, int b) {
    constexpr int RDX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int RDY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    const int p = *c;
    const int cb = p + b, c_b = p - b;
    unsigned bright = 0, dark = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int v = c[RDY[i] * PITCH + RDX[i]];
        bright |= (unsigned)(v > cb) << i;
        dark |= (unsigned)(v < c_b) << i;
    }
    return has_run10(bright) || has_run10(dark);
}

template <int PITCH>
__device__ __forceinline__ bool fast_is_corner(const uint8_t* __restrict__ c, int b) {
    return fast_quick_test<PITCH>(c, b) && fast_full_test<PITCH>(c, b);
}

// fast_corner_score_10 in closed form: the bisection returns the largest b for which the pixel is
// still a corner = max over the 16 arcs of (min over the arc of the signed difference) - 1.
template <int PITCH>
__device__ __forceinline__ int fast_score(const uint8_t* __restrict__ c) {
    constexpr int RDX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int RDY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    const int p = *c;
    int d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = (int)c[RDY[i] * PITCH + RDX[i]] - p;
    int lo2[16], hi2[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        lo2[i] = min(d[i], d[(i + 1) & 15]);
        hi2[i] = max(d[i], d[(i + 1) & 15]);
    }
    int lo4[16], hi4[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        lo4[i] = min(lo2[i], lo2[(i + 2) & 15]);
        hi4[i] = max(hi2[i], hi2[(i + 2) & 15]);
    }
    int best_bright = -256, best_dark = 256;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int lo8 = min(lo4[i], lo4[(i + 4) & 15]);
        const int hi8 = max(hi4[i], hi4[(i + 4) & 15]);
        best_bright = max(best_bright, min(lo8, lo2[(i + 8) & 15]));  // min over arc i..i+9
        best_dark = min(best_dark, max(hi8, hi2[(i + 8) & 15]));      // max over arc i..i+9
    }
    return max(best_bright, -best_dark) - 1;
}

// The same closed form on packed halves: every ring pixel v becomes the word (255 - v) << 16 | v, so ONE
// VIMNMX.U16x2 takes the running minimum of v (low half) and of 255 - v, i.e. the running maximum of v (high
// half), at once: 16 IMAD to pack + 56 packed min/max instead of 112 scalar ones.  Values stay <= 255, so the
// 16-bit lanes never interact.  Returns the score for ANY pixel (below the threshold = not a corner).
template <int PITCH>
__device__ __forceinline__ int fast_score_packed(const uint8_t* __restrict__ c) {
    constexpr int RDX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
    constexpr int RDY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    const int p = *c;
    uint32_t P[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) P[i] = (uint32_t)c[RDY[i] * PITCH + RDX[i]] * 0xFFFF0001u + 0x00FF0000u;
    uint32_t m2[16], m4[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m2[i] = __vminu2(P[i], P[(i + 1) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) m4[i] = __vminu2(m2[i], m2[(i + 2) & 15]);
    uint32_t best = 0;  // (max over arcs of min v) | (max over arcs of min (255 - v)) << 16
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const uint32_t a0 = __vimin3_u16x2(m4[i], m4[(i + 4) & 15], m2[(i + 8) & 15]);          // arc i .. i+9
        const uint32_t a1 = __vimin3_u16x2(m4[i + 1], m4[(i + 5) & 15], m2[(i + 9) & 15]);
        best = __vimax3_u16x2(best, a0, a1);
    }
    const int bright = (int)(best & 0xFFFFu) - p;       // largest arc minimum of v - p
    const int dark = p + (int)(best >> 16) - 255;       // p - smallest arc maximum of v
    return max(bright, dark) - 1;
}

// FeatureDetector::ShiTomasiScore tail: the three sums over the 8x8 box are integers < 2^24, so their
// f32 values are exact whatever the summation order; the rest follows the reference operation by operation.
__device__ __forceinline__ float shi_tomasi_from_sums(int sxx, int syy, int sxy) {
    const float dXX = (float)sxx * 0.0078125f;  // / (2.0 * 64): exact
    const float dYY = (float)syy * 0.0078125f;
    const float dXY = (float)sxy * 0.0078125f;
    const float s = __fadd_rn(dXX, dYY);
    const float disc = __fsub_rn(__fmul_rn(s, s), __fmul_rn(4.f, __fsub_rn(__fmul_rn(dXX, dYY), __fmul_rn(dXY, dXY))));
    return __fmul_rn(0.5f, __fsub_rn(s, __fsqrt_rn(disc)));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 5) fast_cells_kernel(const DetectArgs a, const CUtensorMap* __restrict__ tmaps) {
    __shared__ __align__(128) uint8_t s_img[kSmH * kSmW];
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ __align__(16) uint8_t s_score[kScH * kScPitch];
    __shared__ uint16_t s_list[kScH * kScW];
    __shared__ unsigned long long s_best[kMaxTileCells];
    __shared__ unsigned long long s_first[kMaxTileCells];
    __shared__ uint16_t s_quick[kScH * kScW];
    __shared__ uint8_t s_lutx[kTileW], s_luty[kTileH];
    __shared__ int s_n, s_ncorner, s_nnonmax, s_ncand, s_nquick;
    uint16_t* const s_cand = s_quick;  // the prefilter list is dead once the scores are written (<= 800 candidates)

    const Geometry& g = a.g;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    int L = 0;
    while (L + 1 < g.n_levels && (int)blockIdx.x >= g.tile_begin[L + 1]) ++L;
    const int t = blockIdx.x - g.tile_begin[L];
    const int ty = (int)div_magic((unsigned)t, g.tiles_x_magic[L]), tx = t - ty * g.tiles_x[L];
    const int x0 = tx * kTileW, y0 = ty * kTileH;
    const LevelGeom lv = g.lv[L];
    const int item = blockIdx.y;
    const uint8_t* __restrict__ img = a.pyr + (size_t)a.slots[item] * a.slot_stride + lv.off;
    const bool selectable = L < g.n_sel_levels;
    const int scale = 1 << L;
    const int cpt_x = g.cpt_x[L], cpt_y = g.cpt_y[L];        // grid cells per tile row / column at this level (0: not selectable)
    const int n_tile_cells = cpt_x * cpt_y;                   // <= kMaxTileCells (checked by build_geometry)

    if (tid == 0) s_n = s_ncorner = s_nnonmax = s_ncand = s_nquick = 0;
    for (int i = tid; i < kScH * kScPitch / 4; i += 256) reinterpret_cast<uint32_t*>(s_score)[i] = 0;
    for (int i = tid; i < n_tile_cells; i += 256) {
        s_best[i] = 0ull;
        s_first[i] = kEmptyFirst;
    }
    // stage tile + halo, zero outside the image (never used: FAST stays 3 px inside, Shi-Tomasi returns 0 when its
    // window touches the border).  Levels with a TMA descriptor fetch the whole 112 x 50 box with ONE
    // cp.async.bulk.tensor issued by one thread (out-of-bounds elements are zero-filled by the hardware); the
    // other levels (narrower than the box) use plain word loads.
    if (L < a.tma_levels) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bar = smem_u32(&s_bar), dst = smem_u32(s_img);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(kSmH * kSmW)) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmaps + L)), "r"(x0 - 16), "r"(y0 - 5), "r"(a.slots[item]), "r"(bar)
                : "memory");
        }
        // every thread waits for the transaction bytes to land (phase 0 of the barrier)
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(&s_bar)) : "memory");
        }
    } else {
        for (int i = tid; i < kSmH * (kSmW / 4); i += 256) {
            const int r = i / (kSmW / 4), k = i - r * (kSmW / 4);
            const int y = y0 - 5 + r, x = x0 - 16 + 4 * k;
            uint32_t v = 0;
            if (y >= 0 && y < lv.h) {
                const uint8_t* row = img + (size_t)y * lv.pitch;
                if (x >= 0 && x + 3 < lv.w) {
                    v = *reinterpret_cast<const uint32_t*>(row + x);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x + j >= 0 && x + j < lv.w) v |= (uint32_t)row[x + j] << (8 * j);
                }
            }
            reinterpret_cast<uint32_t*>(s_img)[i] = v;
        }
    }
    __syncthreads();

    // pass A: cheap necessary condition on tile + 1 px ring; the survivors (10-20 % on textured frames) are
    // compacted so that the exact score (pass B) runs on dense warps.
    // region pixel (rx, ry) = image pixel (x0 - 1 + rx, y0 - 1 + ry); FAST is defined for x in [3, w-3), y in [3, h-3)
    const int rx_lo = max(0, 4 - x0), rx_hi = min(kScW, lv.w - 2 - x0);
    const int ry_lo = max(0, 4 - y0), ry_hi = min(kScH, lv.h - 2 - y0);
    if (g.threshold <= 126) {
        // SIMD-within-a-register prefilter, 4 horizontally adjacent pixels per thread, on the eight EVEN ring positions:
        // a position is flagged when it differs from the centre by more than the threshold (|v - p| > b, sign ignored,
        // so the filter is a superset of the bright and of the dark test).  VABSDIFF4 gives the four byte differences
        // in one instruction, ((x & 0x7f7f7f7f) + (127-b)*0x01010101 | x) & 0x80808080 flags the bytes > b without
        // inter-byte carries.  Survivors are appended with one shared-memory atomic per warp and iteration (four
        // ballots give every flagged byte its slot).
        // A thread tests 8 adjacent pixels (two words) per item, two items per thread cover the 42 x 88 padded region;
        // the 16 result flags stay in a register and are compacted ONCE: warp scan of the per-thread counts, one
        // shared-memory atomic per warp, then each thread writes its (few) survivors.
        const uint32_t kadd = 0x01010101u * (uint32_t)(127 - g.threshold);
        constexpr int kPairs = 11;               // items of 8 pixels per region row (88 >= 82 columns)
        constexpr int W = kSmW / 4;              // words per smem row
#define FLAG(v, ctr) ((((__vabsdiffu4((v), (ctr)) & 0x7f7f7f7fu) + kadd) | __vabsdiffu4((v), (ctr))) & 0x80808080u)
        // five consecutive (circular) flags among the eight even ring positions f0..f7 = ring 0, 2, .. 14: a 10-arc
        // contains five consecutive even positions, so this is necessary for a corner (and implies the classic
        // "one pixel of every opposite pair" test)
#define RUN5(f0, f1, f2, f3, f4, f5, f6, f7, out)                                                         \
    {                                                                                                    \
        const uint32_t t0 = f0 & f1 & f2, t1 = f1 & f2 & f3, t2 = f2 & f3 & f4, t3 = f3 & f4 & f5;       \
        const uint32_t t4 = f4 & f5 & f6, t5 = f5 & f6 & f7, t6 = f6 & f7 & f0, t7 = f7 & f0 & f1;       \
        out = (t0 & t2) | (t1 & t3) | (t2 & t4) | (t3 & t5) | (t4 & t6) | (t5 & t7) | (t6 & t0) | (t7 & t1); \
    }
        unsigned flags16 = 0;   // bit 8*it + j: pixel j of item `it` of this thread survives
        int idx_of[2];
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int item_id = it * 256 + tid;
            const int ry = item_id / kPairs, rx0 = (item_id - ry * kPairs) * 8;
            idx_of[it] = ry * kScW + rx0;
            if (ry >= ry_lo && ry < ry_hi) {
                // word-aligned base of the centre row: pixel rx0 sits at byte offset rx0 + 15 = (rx0 + 12) + 3
                const uint32_t* rowc = reinterpret_cast<const uint32_t*>(s_img + (ry + 4) * kSmW + rx0 + 12);
                const uint32_t* ru3 = rowc - 3 * W;
                const uint32_t* rd3 = rowc + 3 * W;
                const uint32_t* ru2 = rowc - 2 * W;
                const uint32_t* rd2 = rowc + 2 * W;
                const uint32_t c0 = rowc[0], c1 = rowc[1], c2 = rowc[2], c3 = rowc[3];
                const uint32_t a0 = ru3[0], a1 = ru3[1], a2 = ru3[2], b0 = rd3[0], b1 = rd3[1], b2 = rd3[2];
                const uint32_t u0 = ru2[0], u1 = ru2[1], u2 = ru2[2], u3 = ru2[3];
                const uint32_t d0 = rd2[0], d1 = rd2[1], d2 = rd2[2], d3 = rd2[3];
                const uint32_t ctrA = __funnelshift_r(c0, c1, 24), ctrB = __funnelshift_r(c1, c2, 24);
                const uint32_t um1 = __funnelshift_r(u1, u2, 8), dm1 = __funnelshift_r(d1, d2, 8);  // column +2 of A = -2 of B
                uint32_t mA, mB;
                {
                    const uint32_t f0 = FLAG(__funnelshift_r(b0, b1, 24), ctrA);   // ring 0  ( 0,+3)
                    const uint32_t f1 = FLAG(dm1, ctrA);                           // ring 2  (+2,+2)
                    const uint32_t f2 = FLAG(__funnelshift_r(c1, c2, 16), ctrA);   // ring 4  (+3, 0)
                    const uint32_t f3 = FLAG(um1, ctrA);                           // ring 6  (+2,-2)
                    const uint32_t f4 = FLAG(__funnelshift_r(a0, a1, 24), ctrA);   // ring 8  ( 0,-3)
                    const uint32_t f5 = FLAG(__funnelshift_r(u0, u1, 8), ctrA);    // ring 10 (-2,-2)
                    const uint32_t f6 = FLAG(c0, ctrA);                            // ring 12 (-3, 0)
                    const uint32_t f7 = FLAG(__funnelshift_r(d0, d1, 8), ctrA);    // ring 14 (-2,+2)
                    RUN5(f0, f1, f2, f3, f4, f5, f6, f7, mA)
                }
                {
                    const uint32_t f0 = FLAG(__funnelshift_r(b1, b2, 24), ctrB);
                    const uint32_t f1 = FLAG(__funnelshift_r(d2, d3, 8), ctrB);
                    const uint32_t f2 = FLAG(__funnelshift_r(c2, c3, 16), ctrB);
                    const uint32_t f3 = FLAG(__funnelshift_r(u2, u3, 8), ctrB);
                    const uint32_t f4 = FLAG(__funnelshift_r(a1, a2, 24), ctrB);
                    const uint32_t f5 = FLAG(um1, ctrB);
                    const uint32_t f6 = FLAG(c1, ctrB);
                    const uint32_t f7 = FLAG(dm1, ctrB);
                    RUN5(f0, f1, f2, f3, f4, f5, f6, f7, mB)
                }
                // bit 7 of every byte -> one bit per pixel, then drop the columns outside [rx_lo, rx_hi)
                unsigned f8 = ((((mA >> 7) & 0x01010101u) * 0x01020408u) >> 24) | (((((mB >> 7) & 0x01010101u) * 0x01020408u) >> 24) << 4);
                const int lo = min(max(rx_lo - rx0, 0), 8), hi = min(max(rx_hi - rx0, 0), 8);
                f8 &= ((1u << hi) - 1u) & ~((1u << lo) - 1u);
                flags16 |= f8 << (8 * it);
            }
        }
#undef FLAG
#undef RUN5
        const int cnt = __popc(flags16);
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += up;
        }
        const int tot = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (tot) {  // warp-uniform
            int base = 0;
            if (lane == 31) base = atomicAdd(&s_nquick, tot);
            base = __shfl_sync(0xFFFFFFFFu, base, 31);
            int pos = base + incl - cnt;
            while (flags16) {
                const int k = __ffs(flags16) - 1;
                flags16 &= flags16 - 1;
                s_quick[pos++] = (uint16_t)((k & 8 ? idx_of[1] : idx_of[0]) + (k & 7));
            }
        }
    } else {
        // thresholds the byte trick cannot express: exact opposite-pair test, one warp per region row, lanes over the
        // 82 columns in three passes, survivors appended with one shared-memory atomic per warp and pass
        const int warp_ = tid >> 5;
        for (int ry = ry_lo + warp_; ry < ry_hi; ry += 8) {
            const uint8_t* row = s_img + (ry + 4) * kSmW + kSmX;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const int rx = pass * 32 + lane;
                bool cand = false;
                if (rx >= rx_lo && rx < rx_hi) cand = fast_quick_test<kSmW>(row + rx, g.threshold);
                const unsigned m = __ballot_sync(0xFFFFFFFFu, cand);
                if (m) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&s_nquick, __popc(m));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (cand) s_quick[base + __popc(m & lt_mask)] = (uint16_t)(ry * kScW + rx);
                }
            }
        }
    }
    // grid-cell lookup tables of this tile (local cell column / row of every tile pixel)
    if (selectable) {
        // tiles start on cell boundaries (build_geometry), so the local cell of pixel i is (i * 2^L) / cell_size
        if (tid < kTileW) s_lutx[tid] = (uint8_t)div_magic((unsigned)(tid << L), g.cell_magic);
        else if (tid < kTileW + kTileH) s_luty[tid - kTileW] = (uint8_t)div_magic((unsigned)((tid - kTileW) << L), g.cell_magic);
    }
    __syncthreads();

    // pass B: exact score of every survivor; a pixel is a corner iff its score reaches the threshold (the score is the
    // largest threshold at which the segment test still holds).  Corners go into the dense score tile (0 = not a
    // corner; scores are >= threshold > 0) and into the corner list.
    const int n_quick = s_nquick;
    for (int i0 = 0; i0 < n_quick; i0 += 256) {  // block-uniform trip count
        const int i = i0 + tid;
        bool corner = false;
        int idx = 0;
        if (i < n_quick) {
            idx = s_quick[i];
            const int ry = idx / kScW, rx = idx - ry * kScW;
            const int s = fast_score_packed<kSmW>(s_img + (ry + 4) * kSmW + (rx + kSmX));
            if (s >= g.threshold) {
                corner = true;
                s_score[ry * kScPitch + rx] = (uint8_t)s;
            }
        }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, corner);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_n, __popc(bal));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (corner) s_list[base + __popc(bal & lt_mask)] = (uint16_t)idx;
        }
    }
    __syncthreads();
    const int n_corner = s_n;

    // pass C1: 3x3 non-max on the tile interior; survivors that pass InFrame / grid / occupancy become
    // cell candidates (at most one per 2x2 block: neighbours with equal or larger score kill each other)
    const int cx0 = tx * cpt_x, cy0 = ty * cpt_y;  // first grid cell of the tile (tiles start on cell boundaries)
    {
        int my_corner = 0, my_nonmax = 0;
        for (int i = tid; i < n_corner; i += 256) {
            const int idx = s_list[i];
            const int ry = idx / kScW, rx = idx - ry * kScW;
            if (rx < 1 || rx > kTileW || ry < 1 || ry > kTileH) continue;  // ring pixels belong to a neighbour tile
            ++my_corner;
            const uint8_t* sc = s_score + ry * kScPitch + rx;
            const int s = *sc;
            // fast_nonmax_3x3: dropped if any 8-neighbour corner scores >= s
            if (sc[-1] >= s || sc[1] >= s || sc[-kScPitch - 1] >= s || sc[-kScPitch] >= s || sc[-kScPitch + 1] >= s ||
                sc[kScPitch - 1] >= s || sc[kScPitch] >= s || sc[kScPitch + 1] >= s)
                continue;
            ++my_nonmax;
            if (!selectable) continue;
            const int x = x0 - 1 + rx, y = y0 - 1 + ry;
            // Frame::InFrame(Vector2d(x,y), 20, L): x/2^L >= 20 && x/2^L < W-20 (exact in integers)
            if (x < 20 * scale || x >= (g.W - 20) * scale || y < 20 * scale || y >= (g.H - 20) * scale) continue;
            const int gy = cy0 + s_luty[ry - 1], gx = cx0 + s_lutx[rx - 1];
            const int k = gy * g.grid_cols + gx;
            if (k >= g.n_cells) continue;
            if (a.occupied && a.occupied[(size_t)item * g.n_cells + k]) continue;
            s_cand[atomicAdd(&s_ncand, 1)] = (uint16_t)idx;
        }
        // the per-level statistics need one shared-memory atomic per warp, not per corner
        my_corner = __reduce_add_sync(0xFFFFFFFFu, my_corner);
        my_nonmax = __reduce_add_sync(0xFFFFFFFFu, my_nonmax);
        if (lane == 0 && my_corner) {
            atomicAdd(&s_ncorner, my_corner);
            atomicAdd(&s_nnonmax, my_nonmax);
        }
    }
    __syncthreads();

    // pass C2: eight lanes per candidate (four candidates per warp): Shi-Tomasi on the staged tile, one row of the
    // 8x8 window per lane (the integer sums are exact, so the f32 result equals the reference's sequential
    // accumulation whatever the order), then the two cell keys
    {
        const int sub = lane & 7;
        const int n_cand = s_ncand;
        for (int i0 = (tid >> 5) * 4; i0 < n_cand; i0 += 32) {  // warp-uniform trip count
            const int i = i0 + (lane >> 3);
            const bool active = i < n_cand;
            const int idx = active ? s_cand[i] : 0;
            const int ry = idx / kScW, rx = idx - ry * kScW;
            const int x = x0 - 1 + rx, y = y0 - 1 + ry;
            const bool inside = active && !(x - 4 < 1 || x + 4 >= lv.w - 1 || y - 4 < 1 || y + 4 >= lv.h - 1);
            int sxx = 0, syy = 0, sxy = 0;
            if (inside) {
                // window row y - 4 + sub, columns x - 4 .. x + 3
                const uint8_t* q = s_img + (ry + 4 + sub - 4) * kSmW + (rx + kSmX - 4);
                int left = q[-1], mid = q[0];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int right = q[c + 1];
                    const int dx = right - left, dy = (int)q[c + kSmW] - (int)q[c - kSmW];
                    sxx += dx * dx;
                    syy += dy * dy;
                    sxy += dx * dy;
                    left = mid;
                    mid = right;
                }
            }
#pragma unroll
            for (int o = 4; o >= 1; o >>= 1) {
                sxx += __shfl_xor_sync(0xFFFFFFFFu, sxx, o);
                syy += __shfl_xor_sync(0xFFFFFFFFu, syy, o);
                sxy += __shfl_xor_sync(0xFFFFFFFFu, sxy, o);
            }
            if (active && sub == 0) {
                const float score = inside ? shi_tomasi_from_sums(sxx, syy, sxy) : 0.f;
                const int li = (int)s_luty[ry - 1] * cpt_x + (int)s_lutx[rx - 1];
                const unsigned raster = (unsigned)(y * lv.w + x);
                atomicMin(&s_first[li], ((unsigned long long)raster << 32) | __float_as_uint(score));
                if (!isnan(score))
                    atomicMax(&s_best[li], ((unsigned long long)float_orderable(score) << 32) | (0xFFFFFFFFu - raster));
            }
        }
    }
    __syncthreads();

    if (tid == 0) {
        int32_t* st = a.stats + ((size_t)item * g.n_levels + L) * 2;
        if (s_ncorner) atomicAdd(st, s_ncorner);
        if (s_nnonmax) atomicAdd(st + 1, s_nnonmax);
    }
    if (selectable) {
        const size_t base = ((size_t)item * g.n_sel_levels + L) * g.n_cells;
        for (int li = tid; li < n_tile_cells; li += 256) {
            const int gy = cy0 + li / cpt_x, gx = cx0 + li % cpt_x;
            if (gy >= g.grid_rows || gx >= g.grid_cols) continue;
            a.best_key[base + gy * g.grid_cols + gx] = s_best[li];
            a.first_key[base + gy * g.grid_cols + gx] = s_first[li];
        }
    }
}

// replay of the level order per grid cell + ordered compaction into the slot's feature store
__global__ void __launch_bounds__(1024) merge_cells_kernel(const unsigned long long* __restrict__ best_key,
                                                           const unsigned long long* __restrict__ first_key,
                                                           const int32_t* __restrict__ slots, Geometry g, int32_t* count,
                                                           int16_t* fx, int16_t* fy, uint8_t* flevel, float* fscore,
                                                           int32_t* fcell) {
    constexpr int kMaxPerThread = 8;
    __shared__ int s_warp[32];
    const int item = blockIdx.x, tid = threadIdx.x;
    const int per = (g.n_cells + 1023) / 1024;
    const int k0 = tid * per, k1 = min(g.n_cells, k0 + per);
    int wl[kMaxPerThread], wx[kMaxPerThread], wy[kMaxPerThread], wk[kMaxPerThread];
    float ws[kMaxPerThread];
    int nloc = 0;
    for (int k = k0; k < k1; ++k) {
        bool have = false, stuck = false;
        int bl = 0;
        unsigned braster = 0;
        float bscore = 0.f;
        for (int L = 0; L < g.n_sel_levels && !stuck; ++L) {
            const size_t o = ((size_t)item * g.n_sel_levels + L) * g.n_cells + k;
            const unsigned long long fk = first_key[o];
            if (fk == kEmptyFirst) continue;
            const unsigned long long bk = best_key[o];
            if (!have) {
                const float fs = __uint_as_float((unsigned)fk);
                if (isnan(fs)) {  // first candidate of the cell is NaN: it is never replaced
                    have = stuck = true;
                    bl = L;
                    braster = (unsigned)(fk >> 32);
                    bscore = fs;
                } else {
                    have = true;
                    bl = L;
                    braster = 0xFFFFFFFFu - (unsigned)bk;
                    bscore = orderable_float((unsigned)(bk >> 32));
                }
            } else if (bk != 0ull) {
                const float s = orderable_float((unsigned)(bk >> 32));
                if (s > bscore) {
                    bl = L;
                    braster = 0xFFFFFFFFu - (unsigned)bk;
                    bscore = s;
                }
            }
        }
        if (have && nloc < kMaxPerThread) {
            wl[nloc] = bl;
            wy[nloc] = braster / g.lv[bl].w;
            wx[nloc] = braster - wy[nloc] * g.lv[bl].w;
            ws[nloc] = bscore;
            wk[nloc] = k;
            ++nloc;
        }
    }
    // block exclusive scan of nloc
    const int lane = tid & 31, warp = tid >> 5;
    int incl = nloc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xFFFFFFFFu, v, o);
            if (lane >= o) v += u;
        }
        s_warp[lane] = v;
    }
    __syncthreads();
    const int base = (warp ? s_warp[warp - 1] : 0) + incl - nloc;
    const size_t so = (size_t)slots[item] * g.n_cells;
    // cells of one thread are consecutive and its winners were pushed in cell order
    for (int w = 0; w < nloc; ++w) {
        fx[so + base + w] = (int16_t)wx[w];
        fy[so + base + w] = (int16_t)wy[w];
        flevel[so + base + w] = (uint8_t)wl[w];
        fscore[so + base + w] = ws[w];
        fcell[so + base + w] = wk[w];
    }
    if (tid == 1023) count[slots[item]] = base + nloc;
}

// ---- parity/debug: dense score map and nonmax flags of one level, straight from global memory ---
__global__ void fast_debug_score_kernel(const uint8_t* __restrict__ img, LevelGeom lv, int threshold,
                                        uint8_t* __restrict__ score_map) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= lv.w) return;
    int s = 0;
    if (x >= 3 && x < lv.w - 3 && y >= 3 && y < lv.h - 3) {
        // gather the 7x7 neighbourhood into a private pitch-8 window so the templated code is shared
        uint8_t win[7 * 8];
        for (int j = 0; j < 7; ++j)
            for (int i = 0; i < 7; ++i) win[j * 8 + i] = img[(size_t)(y - 3 + j) * lv.pitch + (x - 3 + i)];
        const uint8_t* c = win + 3 * 8 + 3;
        if (fast_is_corner<8>(c, threshold)) s = fast_score<8>(c);
    }
    score_map[(size_t)y * lv.w + x] = (uint8_t)s;
}

__global__ void fast_debug_nonmax_kernel(const uint8_t* __restrict__ score_map, int w, int h,
                                         uint8_t* __restrict__ nonmax_map) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const int s = score_map[(size_t)y * w + x];
    int keep = s > 0;
    for (int dy = -1; dy <= 1 && keep; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if (!dx && !dy) continue;
            const int xx = x + dx, yy = y + dy;
            if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
            if (score_map[(size_t)yy * w + xx] >= s) {
                keep = 0;
                break;
            }
        }
    nonmax_map[(size_t)y * w + x] = (uint8_t)keep;
}

}  // namespace

int launch_detect(ygzb_frames* f, int n, bool have_occupied) {
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    YGZB_CUDA(ctx, cudaMemsetAsync(f->d_stats, 0, (size_t)n * g.n_levels * 2 * sizeof(int32_t), ctx->stream));
    DetectArgs a;
    a.pyr = f->d_pyr;
    a.slot_stride = ctx->slot_stride;
    a.slots = f->d_slots;
    a.occupied = have_occupied ? f->d_occupied : nullptr;
    a.best_key = f->d_best_key;
    a.first_key = f->d_first_key;
    a.stats = f->d_stats;
    a.g = g;
    dim3 grid(g.tile_begin[g.n_levels], n);
    {
        ProfScope ps(ctx, kStageFastCells);
        a.tma_levels = f->tma_levels;
        // 5 CTAs x 31 KB per SM: ask for the large shared-memory carve-out (the tiles live in smem, L1 is barely used)
        static std::once_flag carveout_once;   // contexts of several host threads share the kernel's attribute
        std::call_once(carveout_once, [] {
            cudaFuncSetAttribute(fast_cells_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        });
        fast_cells_kernel<<<grid, 256, 0, ctx->stream>>>(a, reinterpret_cast<const CUtensorMap*>(f->d_tile_maps));
    }
    YGZB_LAUNCHED(ctx);
    ProfScope ps(ctx, kStageMergeCells);
    merge_cells_kernel<<<n, 1024, 0, ctx->stream>>>(f->d_best_key, f->d_first_key, f->d_slots, g, f->d_count, f->d_fx,
                                                    f->d_fy, f->d_flevel, f->d_fscore, f->d_fcell);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

// CUtensorMap descriptors for the FAST tile fetch.  cuTensorMapEncodeTiled is a driver-API entry point; it is resolved
// through the runtime (cudaGetDriverEntryPoint), so the library does not link against libcuda.
int build_tile_maps(ygzb_frames* f) {
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    f->tma_levels = 0;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return YGZB_OK;  // no descriptor support: the kernel keeps its plain-load staging
    }
    const EncodeTiled encode = reinterpret_cast<EncodeTiled>(fn);
    int n = 0;
    for (int L = 0; L < g.n_levels && L < 3; ++L) {
        if (g.lv[L].w < kSmW || g.lv[L].h < kSmH) break;  // the box must fit the level
        const cuuint64_t dims[3] = {(cuuint64_t)g.lv[L].w, (cuuint64_t)g.lv[L].h, (cuuint64_t)f->capacity};
        const cuuint64_t strides[2] = {(cuuint64_t)g.lv[L].pitch, (cuuint64_t)ctx->slot_stride};
        const cuuint32_t box[3] = {(cuuint32_t)kSmW, (cuuint32_t)kSmH, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUtensorMap m;
        const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, f->d_pyr + g.lv[L].off, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) break;
        memcpy(f->tile_maps[L], &m, sizeof(m));
        n = L + 1;
    }
    if (n > 0) {
        // the descriptors live in global memory (written once by the host, read-only afterwards)
        if (!f->d_tile_maps && cudaMalloc(&f->d_tile_maps, sizeof(f->tile_maps)) != cudaSuccess) {
            cudaGetLastError();
            return YGZB_OK;
        }
        if (cudaMemcpy(f->d_tile_maps, f->tile_maps, sizeof(f->tile_maps), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            return YGZB_OK;
        }
    }
    f->tma_levels = n;
    return YGZB_OK;
}

int launch_fast_debug(ygzb_frames* f, int slot, int level, uint8_t* d_score_map, uint8_t* d_nonmax_map) {
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    const LevelGeom lv = g.lv[level];
    const uint8_t* img = f->d_pyr + (size_t)slot * ctx->slot_stride + lv.off;
    dim3 grid((lv.w + 127) / 128, lv.h);
    fast_debug_score_kernel<<<grid, 128, 0, ctx->stream>>>(img, lv, g.threshold, d_score_map);
    YGZB_LAUNCHED(ctx);
    fast_debug_nonmax_kernel<<<grid, 128, 0, ctx->stream>>>(d_score_map, lv.w, lv.h, d_nonmax_map);
    YGZB_LAUNCHED(ctx);
    return YGZB_OK;
}

}  // namespace ygzb
