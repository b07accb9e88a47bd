// api.cu -- extern "C" entry points of libygz_b200.so (include/ygz_b200.h): context, frame slots,
// host<->device marshalling around the stage launchers.  No computation happens on the host.
#include <stdarg.h>

#include <algorithm>
#include <new>

#include "common.cuh"

namespace ygzb {

int set_error(ygzb_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int check_cuda(ygzb_ctx* ctx, cudaError_t e, const char* what) {
    if (e == cudaSuccess) return YGZB_OK;
    return set_error(ctx, YGZB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

void* dev_scratch(ygzb_ctx* ctx, int which, size_t bytes) {
    if (ctx->d_scratch_bytes[which] >= bytes && ctx->d_scratch[which]) return ctx->d_scratch[which];
    if (ctx->d_scratch[which]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFree(ctx->d_scratch[which]);
        ctx->d_scratch[which] = nullptr;
        ctx->d_scratch_bytes[which] = 0;
    }
    const size_t want = std::max(bytes + bytes / 4, (size_t)4096);
    void* p = nullptr;
    if (check_cuda(ctx, cudaMalloc(&p, want), "cudaMalloc(scratch)") != YGZB_OK) return nullptr;
    ctx->d_scratch[which] = p;
    ctx->d_scratch_bytes[which] = want;
    return p;
}

void* host_scratch(ygzb_ctx* ctx, int which, size_t bytes) {
    if (ctx->h_scratch_bytes[which] >= bytes && ctx->h_scratch[which]) return ctx->h_scratch[which];
    if (ctx->h_scratch[which]) {
        cudaStreamSynchronize(ctx->stream);
        cudaFreeHost(ctx->h_scratch[which]);
        ctx->h_scratch[which] = nullptr;
        ctx->h_scratch_bytes[which] = 0;
    }
    const size_t want = std::max(bytes + bytes / 4, (size_t)4096);
    void* p = nullptr;
    if (check_cuda(ctx, cudaMallocHost(&p, want), "cudaMallocHost(scratch)") != YGZB_OK) return nullptr;
    ctx->h_scratch[which] = p;
    ctx->h_scratch_bytes[which] = want;
    return p;
}

void prof_begin(ygzb_ctx* ctx, int stage) {
    auto* v = static_cast<std::vector<ProfRec>*>(ctx->prof);
    if (!v) ctx->prof = v = new std::vector<ProfRec>();
    ProfRec r;
    r.stage = stage;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, ctx->stream);
    v->push_back(r);
}

void prof_end(ygzb_ctx* ctx) {
    auto* v = static_cast<std::vector<ProfRec>*>(ctx->prof);
    if (v && !v->empty()) cudaEventRecord(v->back().b, ctx->stream);
}

namespace {

// gather the per-slot feature store of the n items of a call into packed SoA arrays
__global__ void pack_features_kernel(const int32_t* __restrict__ slots, const int32_t* __restrict__ offsets, int n_cells,
                                     const int32_t* __restrict__ count, const int16_t* __restrict__ fx,
                                     const int16_t* __restrict__ fy, const uint8_t* __restrict__ flevel,
                                     const float* __restrict__ fscore, const float* __restrict__ fangle,
                                     const int32_t* __restrict__ fcell, const uint8_t* __restrict__ fdesc, float* ox,
                                     float* oy, uint8_t* olevel, float* oscore, float* oangle, int32_t* ocell,
                                     uint8_t* odesc) {
    const int item = blockIdx.y, slot = slots[item];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count[slot]) return;
    const size_t s = (size_t)slot * n_cells + i, d = (size_t)offsets[item] + i;
    const int L = flevel[s];
    ox[d] = (float)((int)fx[s] << L);  // Feature::_pixel = level coordinate * 2^level (FeatureDetector.cpp:415-424)
    oy[d] = (float)((int)fy[s] << L);
    olevel[d] = (uint8_t)L;
    oscore[d] = fscore[s];
    oangle[d] = fangle[s];
    ocell[d] = fcell[s];
    const uint4* src = reinterpret_cast<const uint4*>(fdesc) + s * 2;
    uint4* dst = reinterpret_cast<uint4*>(odesc) + d * 2;
    dst[0] = src[0];
    dst[1] = src[1];
}

__global__ void expand_slots_kernel(const int32_t* __restrict__ slots, const int32_t* __restrict__ offsets, int n,
                                    int32_t* __restrict__ slot_of) {
    const int item = blockIdx.y;
    const int i = offsets[item] + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < offsets[item + 1]) slot_of[i] = slots[item];
}

int build_geometry(ygzb_ctx* ctx) {
    const ygzb_params& p = ctx->prm;
    Geometry& g = ctx->geo;
    memset(&g, 0, sizeof(g));
    if (p.n_levels < 1 || p.n_levels > kMaxLevels) return set_error(ctx, YGZB_ERR_INVALID, "n_levels out of range");
    if (p.image_width < 16 || p.image_height < 16 || p.image_width > 8192 || p.image_height > 8192)
        return set_error(ctx, YGZB_ERR_INVALID, "unsupported image size");
    if (p.cell_size < 1) return set_error(ctx, YGZB_ERR_INVALID, "cell_size must be positive");
    // the FAST kernel keeps scores in u8 with 0 = "not a corner" and builds SWAR constants from 127 - threshold
    if (p.fast_threshold < 1 || p.fast_threshold > 254) return set_error(ctx, YGZB_ERR_INVALID, "fast_threshold must be in [1, 254]");
    g.n_levels = p.n_levels;
    g.W = p.image_width;
    g.H = p.image_height;
    g.cell_size = p.cell_size;
    g.grid_cols = (p.image_width + p.cell_size - 1) / p.cell_size;   // ceil(double(w)/cell) FeatureDetector.cpp:336-337
    g.grid_rows = (p.image_height + p.cell_size - 1) / p.cell_size;
    g.n_cells = g.grid_cols * g.grid_rows;
    g.threshold = p.fast_threshold;
    int w = p.image_width, h = p.image_height;
    size_t off = 0;
    g.tile_begin[0] = 0;
    for (int L = 0; L < p.n_levels; ++L) {
        g.lv[L].w = w;
        g.lv[L].h = h;
        g.lv[L].pitch = (w + 15) & ~15;
        g.lv[L].off = (unsigned)off;
        off += (size_t)g.lv[L].pitch * h;
        off = (off + 255) & ~(size_t)255;
        g.tiles_x[L] = (w + kTileW - 1) / kTileW;
        g.tile_begin[L + 1] = g.tile_begin[L] + g.tiles_x[L] * ((h + kTileH - 1) / kTileH);
        w = (w + 1) / 2;
        h = (h + 1) / 2;
    }
    ctx->slot_stride = off;
    // levels whose corners can pass Frame::InFrame(px, 20, L): need a FAST pixel x in [3, w_L-3) with
    // x >= 20 * 2^L (and the same vertically)
    g.n_sel_levels = 0;
    for (int L = 0; L < p.n_levels; ++L) {
        const bool can = (20 << L) < g.lv[L].w - 3 && (20 << L) < g.lv[L].h - 3;
        if (!can) break;
        g.n_sel_levels = L + 1;
    }
    for (int L = 0; L < g.n_sel_levels; ++L) {
        const int sx = kTileW << L, sy = kTileH << L;
        if (sx % p.cell_size || sy % p.cell_size || (sx / p.cell_size) * (sy / p.cell_size) > 512)
            return set_error(ctx, YGZB_ERR_INVALID,
                             "cell_size %d does not tile the %dx%d FAST tile at level %d (supported: divisors of 40 such as 5, 8, 10, 20)",
                             p.cell_size, kTileW, kTileH, L);
    }
    if (g.n_cells > 8 * 1024) return set_error(ctx, YGZB_ERR_INVALID, "grid has too many cells");
    // magic multipliers for the divisions of the FAST kernel (see common.cuh)
    g.cell_magic = (unsigned)(((1ull << 24) + p.cell_size - 1) / p.cell_size);
    for (int L = 0; L < p.n_levels; ++L) {
        g.tiles_x_magic[L] = (unsigned)(((1ull << 24) + g.tiles_x[L] - 1) / g.tiles_x[L]);
        const long long n_tiles = g.tile_begin[L + 1] - g.tile_begin[L];
        if (n_tiles * g.tiles_x[L] >= (1ll << 24)) return set_error(ctx, YGZB_ERR_INVALID, "image too large for the tile index arithmetic");
        if (L < g.n_sel_levels) {
            g.cpt_x[L] = (kTileW << L) / p.cell_size;
            g.cpt_y[L] = (kTileH << L) / p.cell_size;
            if ((long long)(kTileW << L) * p.cell_size >= (1ll << 24)) return set_error(ctx, YGZB_ERR_INVALID, "cell arithmetic out of range");
        }
    }
    return YGZB_OK;
}

template <typename T>
int dalloc(ygzb_ctx* ctx, T** p, size_t count) {
    return check_cuda(ctx, cudaMalloc((void**)p, std::max(count, (size_t)1) * sizeof(T)), "cudaMalloc");
}

}  // namespace
}  // namespace ygzb

using namespace ygzb;

extern "C" {

void ygzb_default_params(ygzb_params* p) {
    p->image_width = 640;
    p->image_height = 480;
    p->n_levels = 3;
    p->cell_size = 10;
    p->fast_threshold = 15;
    p->fx = 520.9f;
    p->fy = 521.0f;
    p->cx = 325.1f;
    p->cy = 249.7f;
}

int ygzb_create(int device, const ygzb_params* p, ygzb_ctx** out) {
    if (!p || !out) return YGZB_ERR_INVALID;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev) return YGZB_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return YGZB_ERR_NO_DEVICE;
    if (prop.major != 10) return YGZB_ERR_NO_DEVICE;  // the kernels are built for sm_100a only
    ygzb_ctx* ctx = new (std::nothrow) ygzb_ctx();
    if (!ctx) return YGZB_ERR_INVALID;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    ctx->prm = *p;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return YGZB_ERR_CUDA;
    }
    const int rc = build_geometry(ctx);
    if (rc != YGZB_OK) {
        // keep the message reachable: hand the context back so ygzb_last_error works, caller destroys it
        *out = ctx;
        return rc;
    }
    *out = ctx;
    return YGZB_OK;
}

void ygzb_destroy(ygzb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 8; ++i)
        if (ctx->d_scratch[i]) cudaFree(ctx->d_scratch[i]);
    for (int i = 0; i < 4; ++i)
        if (ctx->h_scratch[i]) cudaFreeHost(ctx->h_scratch[i]);
    if (ctx->prof) {
        auto* v = static_cast<std::vector<ProfRec>*>(ctx->prof);
        for (ProfRec& r : *v) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        delete v;
    }
    for (cudaEvent_t e : ctx->timer)
        if (e) cudaEventDestroy(e);
    if (ctx->block_ev) cudaEventDestroy(ctx->block_ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* ygzb_last_error(const ygzb_ctx* ctx) { return ctx ? ctx->err : "null context"; }

int ygzb_synchronize(ygzb_ctx* ctx) {
    if (!ctx) return YGZB_ERR_INVALID;
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return YGZB_OK;
}

// waits on a cudaEventBlockingSync event recorded behind everything enqueued so far: the calling thread sleeps instead of
// spinning (cudaStreamSynchronize spins under the default scheduling flags) -- for hosts with fewer CPUs than waiting threads
int ygzb_synchronize_blocking(ygzb_ctx* ctx) {
    if (!ctx) return YGZB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (!ctx->block_ev) YGZB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->block_ev, cudaEventBlockingSync | cudaEventDisableTiming));
    YGZB_CUDA(ctx, cudaEventRecord(ctx->block_ev, ctx->stream));
    YGZB_CUDA(ctx, cudaEventSynchronize(ctx->block_ev));
    return YGZB_OK;
}

int ygzb_profile_enable(ygzb_ctx* ctx, int on) {
    if (!ctx) return YGZB_ERR_INVALID;
    ctx->prof_on = on ? 1 : 0;
    return YGZB_OK;
}

int ygzb_profile_read(ygzb_ctx* ctx, double* ms, int32_t* launches) {
    if (!ctx || !ms || !launches) return YGZB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < kNumStages; ++i) {
        ms[i] = 0.0;
        launches[i] = 0;
    }
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    auto* v = static_cast<std::vector<ProfRec>*>(ctx->prof);
    if (!v) return YGZB_OK;
    for (ProfRec& r : *v) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            ms[r.stage] += t;
            launches[r.stage] += 1;
        }
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    v->clear();
    return YGZB_OK;
}

int ygzb_profile_stage_count(void) { return kNumStages; }
const char* ygzb_profile_stage_name(int i) {
    static const char* names[kNumStages] = {"bgr2gray", "pyrdown", "fast_cells", "merge_cells", "describe", "match",
                                            "match_finalize", "pack", "other", "align2d", "project_align", "sparse_align",
                                            "pose_only", "local_ba", "klt"};
    return (i >= 0 && i < kNumStages) ? names[i] : "";
}

void* ygzb_stream(ygzb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int ygzb_timer_start(ygzb_ctx* ctx) {
    if (!ctx) return YGZB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    if (!ctx->timer[0]) {
        YGZB_CUDA(ctx, cudaEventCreate(&ctx->timer[0]));
        YGZB_CUDA(ctx, cudaEventCreate(&ctx->timer[1]));
    }
    YGZB_CUDA(ctx, cudaEventRecord(ctx->timer[0], ctx->stream));
    return YGZB_OK;
}

int ygzb_timer_stop(ygzb_ctx* ctx, double* ms) {
    if (!ctx || !ms || !ctx->timer[0]) return YGZB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    YGZB_CUDA(ctx, cudaEventRecord(ctx->timer[1], ctx->stream));
    YGZB_CUDA(ctx, cudaEventSynchronize(ctx->timer[1]));
    float t = 0.f;
    YGZB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->timer[0], ctx->timer[1]));
    *ms = t;
    return YGZB_OK;
}
long long ygzb_launch_count(const ygzb_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ygzb_host_alloc(void** ptr, size_t bytes) {
    return cudaMallocHost(ptr, bytes) == cudaSuccess ? YGZB_OK : YGZB_ERR_CUDA;
}
int ygzb_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? YGZB_OK : YGZB_ERR_CUDA; }

int ygzb_grid_dims(const ygzb_ctx* ctx, int* rows, int* cols) {
    if (!ctx) return YGZB_ERR_INVALID;
    if (rows) *rows = ctx->geo.grid_rows;
    if (cols) *cols = ctx->geo.grid_cols;
    return YGZB_OK;
}

// ---- frames -------------------------------------------------------------------------------------
int ygzb_frames_create(ygzb_ctx* ctx, int capacity, ygzb_frames** out) {
    if (!ctx || !out || capacity < 1) return YGZB_ERR_INVALID;
    *out = nullptr;
    cudaSetDevice(ctx->device);
    ygzb_frames* f = new (std::nothrow) ygzb_frames();
    if (!f) return YGZB_ERR_INVALID;
    memset(f, 0, sizeof(*f));
    f->ctx = ctx;
    f->capacity = capacity;
    const Geometry& g = ctx->geo;
    const size_t cap = (size_t)capacity, nc = (size_t)g.n_cells;
    int rc = YGZB_OK;
    auto A = [&](int r) { if (rc == YGZB_OK) rc = r; };
    A(dalloc(ctx, &f->d_pyr, cap * ctx->slot_stride));
    A(dalloc(ctx, &f->d_count, cap));
    A(dalloc(ctx, &f->d_fx, cap * nc));
    A(dalloc(ctx, &f->d_fy, cap * nc));
    A(dalloc(ctx, &f->d_flevel, cap * nc));
    A(dalloc(ctx, &f->d_fscore, cap * nc));
    A(dalloc(ctx, &f->d_fangle, cap * nc));
    A(dalloc(ctx, &f->d_fcell, cap * nc));
    A(dalloc(ctx, &f->d_fdesc, cap * nc * 32));
    A(dalloc(ctx, &f->d_best_key, cap * std::max(g.n_sel_levels, 1) * nc));
    A(dalloc(ctx, &f->d_first_key, cap * std::max(g.n_sel_levels, 1) * nc));
    A(dalloc(ctx, &f->d_stats, cap * g.n_levels * 2));
    A(dalloc(ctx, &f->d_slots, cap));
    A(dalloc(ctx, &f->d_occupied, cap * nc));
    A(dalloc(ctx, &f->d_offsets, cap + 1));
    if (rc == YGZB_OK) rc = check_cuda(ctx, cudaMemsetAsync(f->d_count, 0, cap * sizeof(int32_t), ctx->stream), "memset");
    if (rc == YGZB_OK) rc = check_cuda(ctx, cudaMemsetAsync(f->d_pyr, 0, cap * ctx->slot_stride, ctx->stream), "memset");
    if (rc == YGZB_OK) rc = build_tile_maps(f);
    if (rc != YGZB_OK) {
        ygzb_frames_destroy(f);
        return rc;
    }
    *out = f;
    return YGZB_OK;
}

void ygzb_frames_destroy(ygzb_frames* f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    void* ptrs[] = {f->d_pyr,   f->d_count,    f->d_fx,        f->d_fy,    f->d_flevel, f->d_fscore,   f->d_fangle, f->d_fcell,
                    f->d_fdesc, f->d_best_key, f->d_first_key, f->d_stats, f->d_slots,  f->d_occupied, f->d_offsets};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (f->d_tile_maps) cudaFree(f->d_tile_maps);
    delete f;
}

int ygzb_frames_layout(const ygzb_frames* f, int* lw, int* lh, int* lpitch, size_t* loff, size_t* slot_stride) {
    if (!f) return YGZB_ERR_INVALID;
    const Geometry& g = f->ctx->geo;
    for (int L = 0; L < g.n_levels; ++L) {
        if (lw) lw[L] = g.lv[L].w;
        if (lh) lh[L] = g.lv[L].h;
        if (lpitch) lpitch[L] = g.lv[L].pitch;
        if (loff) loff[L] = g.lv[L].off;
    }
    if (slot_stride) *slot_stride = f->ctx->slot_stride;
    return YGZB_OK;
}

void* ygzb_frames_device_ptr(ygzb_frames* f) { return f ? f->d_pyr : nullptr; }

int ygzb_frames_build_pyramid(ygzb_frames* f, int first, int count) {
    if (!f || first < 0 || count < 0 || first + count > f->capacity) return YGZB_ERR_INVALID;
    cudaSetDevice(f->ctx->device);
    return launch_pyramid(f, first, count, nullptr);
}

int ygzb_frames_copy(ygzb_frames* f, int src_slot, int dst_slot) {
    if (!f || src_slot < 0 || dst_slot < 0 || src_slot >= f->capacity || dst_slot >= f->capacity) return YGZB_ERR_INVALID;
    if (src_slot == dst_slot) return YGZB_OK;
    ygzb_ctx* ctx = f->ctx;
    cudaSetDevice(ctx->device);
    YGZB_CUDA(ctx, cudaMemcpyAsync(f->d_pyr + (size_t)dst_slot * ctx->slot_stride, f->d_pyr + (size_t)src_slot * ctx->slot_stride,
                                   ctx->slot_stride, cudaMemcpyDeviceToDevice, ctx->stream));
    return YGZB_OK;
}

int ygzb_frames_upload(ygzb_frames* f, int first, int count, const uint8_t* host, int channels, size_t frame_stride) {
    if (!f || !host || first < 0 || count < 0 || first + count > f->capacity) return YGZB_ERR_INVALID;
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    cudaSetDevice(ctx->device);
    const size_t row = (size_t)g.lv[0].w * channels;
    if (channels != 1 && channels != 3) return set_error(ctx, YGZB_ERR_INVALID, "channels must be 1 or 3");
    if (frame_stride < row * g.lv[0].h) return set_error(ctx, YGZB_ERR_INVALID, "frame_stride smaller than one image");
    if (count == 0) return YGZB_OK;
    // cudaMemcpyDefault: `host` may also be a device pointer (frames already resident in HBM, unified addressing).
    // A strided batch copy treats one image as a "row", so its pitch is limited (cudaDeviceProp::memPitch, 2^31 - 1):
    // longer strides (a stacked [stream][frame] array of thousands of frames) fall back to one copy per image.
    const bool one_copy = frame_stride <= (size_t)0x7FFFFFFF;
    if (channels == 1) {
        uint8_t* dst = f->d_pyr + (size_t)first * ctx->slot_stride + g.lv[0].off;
        if (g.lv[0].pitch == g.lv[0].w && one_copy) {
            // level 0 of a slot is one contiguous run: a single strided copy moves the whole batch
            YGZB_CUDA(ctx, cudaMemcpy2DAsync(dst, ctx->slot_stride, host, frame_stride, row * g.lv[0].h, count, cudaMemcpyDefault,
                                             ctx->stream));
        } else {
            for (int i = 0; i < count; ++i)
                YGZB_CUDA(ctx, cudaMemcpy2DAsync(dst + (size_t)i * ctx->slot_stride, g.lv[0].pitch, host + (size_t)i * frame_stride,
                                                 row, row, g.lv[0].h, cudaMemcpyDefault, ctx->stream));
        }
        return launch_pyramid(f, first, count, nullptr);
    }
    uint8_t* d_bgr = (uint8_t*)dev_scratch(ctx, 0, (size_t)count * row * g.lv[0].h);
    if (!d_bgr) return YGZB_ERR_CUDA;
    if (one_copy) {
        YGZB_CUDA(ctx, cudaMemcpy2DAsync(d_bgr, row * g.lv[0].h, host, frame_stride, row * g.lv[0].h, count, cudaMemcpyDefault,
                                         ctx->stream));
    } else {
        for (int i = 0; i < count; ++i)
            YGZB_CUDA(ctx, cudaMemcpyAsync(d_bgr + (size_t)i * row * g.lv[0].h, host + (size_t)i * frame_stride, row * g.lv[0].h,
                                           cudaMemcpyDefault, ctx->stream));
    }
    return launch_pyramid(f, first, count, d_bgr);
}

int ygzb_frames_download_level(ygzb_frames* f, int slot, int level, uint8_t* host) {
    if (!f || !host || slot < 0 || slot >= f->capacity || level < 0 || level >= f->ctx->geo.n_levels) return YGZB_ERR_INVALID;
    ygzb_ctx* ctx = f->ctx;
    const LevelGeom& lv = ctx->geo.lv[level];
    cudaSetDevice(ctx->device);
    YGZB_CUDA(ctx, cudaMemcpy2DAsync(host, lv.w, f->d_pyr + (size_t)slot * ctx->slot_stride + lv.off, lv.pitch, lv.w, lv.h,
                                     cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return YGZB_OK;
}

// ---- FeatureDetector ------------------------------------------------------------------------------
static int upload_slots(ygzb_frames* f, const int32_t* slots, int n) {
    ygzb_ctx* ctx = f->ctx;
    if (!slots || n < 1 || n > f->capacity) return set_error(ctx, YGZB_ERR_INVALID, "bad slot list (n=%d, capacity=%d)", n, f->capacity);
    for (int i = 0; i < n; ++i)
        if (slots[i] < 0 || slots[i] >= f->capacity) return set_error(ctx, YGZB_ERR_INVALID, "slot %d out of range", slots[i]);
    // staged through pinned memory so the async copy does not read a caller buffer that may go away
    int32_t* h = (int32_t*)host_scratch(ctx, 0, (size_t)f->capacity * sizeof(int32_t));
    if (!h) return YGZB_ERR_CUDA;
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h, slots, (size_t)n * sizeof(int32_t));
    YGZB_CUDA(ctx, cudaMemcpyAsync(f->d_slots, h, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    return YGZB_OK;
}

int ygzb_detect(ygzb_frames* f, const int32_t* slots, int n, const uint8_t* occupied, ygzb_keypoints* out) {
    if (!f) return YGZB_ERR_INVALID;
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    cudaSetDevice(ctx->device);
    int rc = upload_slots(f, slots, n);
    if (rc != YGZB_OK) return rc;
    if (occupied)
        YGZB_CUDA(ctx, cudaMemcpyAsync(f->d_occupied, occupied, (size_t)n * g.n_cells, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_detect(f, n, occupied != nullptr)) != YGZB_OK) return rc;
    if ((rc = launch_describe_store(f, n)) != YGZB_OK) return rc;
    f->last_n = n;
    if (!out) return YGZB_OK;

    // packed copy-back: offsets first (one sync), then exactly `total` features per array
    if ((rc = launch_offsets(ctx, f->d_count, f->d_slots, n, f->d_offsets)) != YGZB_OK) return rc;
    // array stride rounded up to 4 elements: the descriptor rows behind the five 4-byte arrays are moved as uint4
    const size_t cap = ((size_t)n * g.n_cells + 3) & ~(size_t)3;
    uint8_t* pk = (uint8_t*)dev_scratch(ctx, 1, cap * 56);
    if (!pk) return YGZB_ERR_CUDA;
    float* ox = (float*)pk;
    float* oy = ox + cap;
    float* oscore = oy + cap;
    float* oangle = oscore + cap;
    int32_t* ocell = (int32_t*)(oangle + cap);
    uint8_t* odesc = (uint8_t*)(ocell + cap);
    uint8_t* olevel = odesc + cap * 32;
    dim3 grid((g.n_cells + 255) / 256, n);
    ProfScope ps(ctx, kStagePack);
    pack_features_kernel<<<grid, 256, 0, ctx->stream>>>(f->d_slots, f->d_offsets, g.n_cells, f->d_count, f->d_fx, f->d_fy,
                                                        f->d_flevel, f->d_fscore, f->d_fangle, f->d_fcell, f->d_fdesc, ox, oy,
                                                        olevel, oscore, oangle, ocell, odesc);
    YGZB_LAUNCHED(ctx);
    YGZB_CUDA(ctx, cudaMemcpyAsync(out->offsets, f->d_offsets, (size_t)(n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t total = (size_t)out->offsets[n];
    if ((size_t)out->capacity < total)
        return set_error(ctx, YGZB_ERR_CAPACITY, "ygzb_keypoints.capacity %d < %zu features", out->capacity, total);
    if (total) {
        YGZB_CUDA(ctx, cudaMemcpyAsync(out->x, ox, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaMemcpyAsync(out->y, oy, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaMemcpyAsync(out->level, olevel, total, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaMemcpyAsync(out->score, oscore, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaMemcpyAsync(out->angle, oangle, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaMemcpyAsync(out->desc, odesc, total * 32, cudaMemcpyDeviceToHost, ctx->stream));
        if (out->cell) YGZB_CUDA(ctx, cudaMemcpyAsync(out->cell, ocell, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return YGZB_OK;
}

int ygzb_detect_stats(ygzb_frames* f, int n, int32_t* stats) {
    if (!f || !stats || n < 1 || n > f->capacity) return YGZB_ERR_INVALID;
    ygzb_ctx* ctx = f->ctx;
    cudaSetDevice(ctx->device);
    YGZB_CUDA(ctx, cudaMemcpyAsync(stats, f->d_stats, (size_t)n * ctx->geo.n_levels * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                   ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return YGZB_OK;
}

int ygzb_describe(ygzb_frames* f, const int32_t* slots, int n, const int32_t* offsets, const double* x, const double* y,
                  const uint8_t* level, float* angle, uint8_t* desc) {
    if (!f || !offsets || !x || !y || !level || !angle || !desc) return YGZB_ERR_INVALID;
    ygzb_ctx* ctx = f->ctx;
    cudaSetDevice(ctx->device);
    int rc = upload_slots(f, slots, n);
    if (rc != YGZB_OK) return rc;
    if (n < 0) return YGZB_ERR_INVALID;
    rc = check_offsets(ctx, offsets, n, "offsets");
    if (rc != YGZB_OK) return rc;
    const int total = offsets[n];
    if (total <= 0) return YGZB_OK;
    for (int i = 0; i < total; ++i)
        if (level[i] >= ctx->geo.n_levels) return set_error(ctx, YGZB_ERR_INVALID, "feature %d: level %d out of range", i, level[i]);
    int max_per = 0;
    for (int i = 0; i < n; ++i) max_per = std::max(max_per, offsets[i + 1] - offsets[i]);
    const size_t T = (size_t)total;
    uint8_t* buf = (uint8_t*)dev_scratch(ctx, 2, T * (8 + 8 + 4 + 4 + 32 + 1) + (size_t)(n + 1) * 4 + 64);
    if (!buf) return YGZB_ERR_CUDA;
    double* dx = (double*)buf;
    double* dy = dx + T;
    float* dangle = (float*)(dy + T);
    int32_t* dslot_of = (int32_t*)(dangle + T);
    int32_t* doff = dslot_of + T;
    uint8_t* ddesc = (uint8_t*)(doff + n + 1);
    ddesc = (uint8_t*)(((uintptr_t)ddesc + 15) & ~(uintptr_t)15);
    uint8_t* dlevel = ddesc + T * 32;
    YGZB_CUDA(ctx, cudaMemcpyAsync(dx, x, T * 8, cudaMemcpyHostToDevice, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(dy, y, T * 8, cudaMemcpyHostToDevice, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(dlevel, level, T, cudaMemcpyHostToDevice, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(doff, offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((max_per + 255) / 256, n);
    expand_slots_kernel<<<grid, 256, 0, ctx->stream>>>(f->d_slots, doff, n, dslot_of);
    YGZB_LAUNCHED(ctx);
    if ((rc = launch_describe_list(f, n, dslot_of, total, dx, dy, dlevel, dangle, ddesc)) != YGZB_OK) return rc;
    YGZB_CUDA(ctx, cudaMemcpyAsync(angle, dangle, T * 4, cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(desc, ddesc, T * 32, cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return YGZB_OK;
}

int ygzb_fast_debug(ygzb_frames* f, int slot, int level, int capacity, int16_t* xy, int32_t* scores, int32_t* n_corners,
                    int32_t* nonmax_idx, int32_t* n_nonmax) {
    if (!f || !xy || !scores || !n_corners || !nonmax_idx || !n_nonmax) return YGZB_ERR_INVALID;
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    if (slot < 0 || slot >= f->capacity || level < 0 || level >= g.n_levels) return YGZB_ERR_INVALID;
    cudaSetDevice(ctx->device);
    const LevelGeom& lv = g.lv[level];
    const size_t px = (size_t)lv.w * lv.h;
    uint8_t* d = (uint8_t*)dev_scratch(ctx, 3, 2 * px);
    uint8_t* h = (uint8_t*)host_scratch(ctx, 1, 2 * px);
    if (!d || !h) return YGZB_ERR_CUDA;
    int rc = launch_fast_debug(f, slot, level, d, d + px);
    if (rc != YGZB_OK) return rc;
    YGZB_CUDA(ctx, cudaMemcpyAsync(h, d, 2 * px, cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // enumeration only (no decisions): raster-order corner list and the positions of the survivors in it
    int nc = 0, nn = 0;
    for (int y = 0; y < lv.h; ++y)
        for (int x = 0; x < lv.w; ++x) {
            const size_t i = (size_t)y * lv.w + x;
            if (!h[i]) continue;
            if (nc < capacity) {
                xy[2 * nc] = (int16_t)x;
                xy[2 * nc + 1] = (int16_t)y;
                scores[nc] = h[i];
                if (h[px + i]) {
                    if (nn < capacity) nonmax_idx[nn] = nc;
                    ++nn;
                }
            }
            ++nc;
        }
    *n_corners = nc;
    *n_nonmax = nn;
    return nc > capacity ? set_error(ctx, YGZB_ERR_CAPACITY, "%d corners > capacity %d", nc, capacity) : YGZB_OK;
}

// ---- Matcher: descriptors ---------------------------------------------------------------------------
int ygzb_match_bf(ygzb_ctx* ctx, const uint8_t* A, int nA, const uint8_t* B, int nB, int cross_check, int32_t* train_idx,
                  int32_t* dist) {
    if (!ctx || nA < 0 || nB < 0 || (nA && !A) || (nB && !B) || !train_idx || !dist) return YGZB_ERR_INVALID;
    if (nA > 65535 || nB > 65535) return set_error(ctx, YGZB_ERR_INVALID, "at most 65535 descriptors per set");
    if (nA == 0) return YGZB_OK;
    cudaSetDevice(ctx->device);
    const int cap = std::max(nA, nB);
    const size_t stride = (size_t)cap * 32;
    // [descs A | descs B | counts(2) a_set b_set q_off(2) | fwd | col | idx | dist]
    uint8_t* buf = (uint8_t*)dev_scratch(ctx, 4, 2 * stride + 64 + (size_t)cap * 16);
    int32_t* h = (int32_t*)host_scratch(ctx, 2, 64);
    if (!buf || !h) return YGZB_ERR_CUDA;
    int32_t* meta = (int32_t*)(buf + 2 * stride);
    unsigned* fwd = (unsigned*)(meta + 16);
    unsigned* col = fwd + cap;
    int32_t* didx = (int32_t*)(col + cap);
    int32_t* ddist = didx + cap;
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    h[0] = nA; h[1] = nB; h[2] = 0; h[3] = 1; h[4] = 0; h[5] = nA;
    YGZB_CUDA(ctx, cudaMemcpyAsync(meta, h, 6 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(buf, A, (size_t)nA * 32, cudaMemcpyHostToDevice, ctx->stream));
    if (nB) YGZB_CUDA(ctx, cudaMemcpyAsync(buf + stride, B, (size_t)nB * 32, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_match(ctx, buf, stride, meta, meta + 2, meta + 3, 1, cap, cross_check, fwd, col, meta + 4, didx, ddist);
    if (rc != YGZB_OK) return rc;
    YGZB_CUDA(ctx, cudaMemcpyAsync(train_idx, didx, (size_t)nA * 4, cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(dist, ddist, (size_t)nA * 4, cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return YGZB_OK;
}

int ygzb_match_frames(ygzb_frames* f, const int32_t* a_slots, const int32_t* b_slots, int n_pairs, int cross_check,
                      int32_t* q_offsets, int32_t* train_idx, int32_t* dist, int capacity) {
    if (!f || !a_slots || !b_slots || n_pairs < 1) return YGZB_ERR_INVALID;
    ygzb_ctx* ctx = f->ctx;
    const Geometry& g = ctx->geo;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < n_pairs; ++i)
        if (a_slots[i] < 0 || a_slots[i] >= f->capacity || b_slots[i] < 0 || b_slots[i] >= f->capacity)
            return set_error(ctx, YGZB_ERR_INVALID, "pair %d: slot out of range", i);
    const size_t P = (size_t)n_pairs, cap = (size_t)g.n_cells;
    // [a_sets | b_sets | q_off (P+1) | fwd | col | idx | dist]
    uint8_t* buf = (uint8_t*)dev_scratch(ctx, 5, (3 * P + 1) * 4 + 64 + P * cap * 16);
    int32_t* h = (int32_t*)host_scratch(ctx, 3, 2 * P * 4);
    if (!buf || !h) return YGZB_ERR_CUDA;
    int32_t* d_a = (int32_t*)buf;
    int32_t* d_b = d_a + P;
    int32_t* d_qoff = d_b + P;
    unsigned* fwd = (unsigned*)(((uintptr_t)(d_qoff + P + 1) + 15) & ~(uintptr_t)15);
    unsigned* col = fwd + P * cap;
    int32_t* didx = (int32_t*)(col + P * cap);
    int32_t* ddist = didx + P * cap;
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h, a_slots, P * 4);
    memcpy(h + P, b_slots, P * 4);
    YGZB_CUDA(ctx, cudaMemcpyAsync(d_a, h, 2 * P * 4, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_offsets(ctx, f->d_count, d_a, n_pairs, d_qoff);
    if (rc != YGZB_OK) return rc;
    rc = launch_match(ctx, f->d_fdesc, cap * 32, f->d_count, d_a, d_b, n_pairs, (int)cap, cross_check, fwd, col, d_qoff, didx,
                      ddist);
    if (rc != YGZB_OK) return rc;
    if (!q_offsets) return YGZB_OK;  // results stay on the device (bench "value" leg)
    YGZB_CUDA(ctx, cudaMemcpyAsync(q_offsets, d_qoff, (P + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t total = (size_t)q_offsets[n_pairs];
    if ((size_t)capacity < total) return set_error(ctx, YGZB_ERR_CAPACITY, "match capacity %d < %zu queries", capacity, total);
    if (total && train_idx && dist) {
        YGZB_CUDA(ctx, cudaMemcpyAsync(train_idx, didx, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaMemcpyAsync(dist, ddist, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return YGZB_OK;
}

int ygzb_hamming_pairs(ygzb_ctx* ctx, const uint8_t* A, int nA, const uint8_t* B, int nB, const int32_t* ia, const int32_t* ib,
                       int n, int32_t* dist) {
    if (!ctx || !A || !B || !ia || !ib || !dist || n < 0 || nA < 1 || nB < 1) return YGZB_ERR_INVALID;
    if (n == 0) return YGZB_OK;
    for (int k = 0; k < n; ++k)
        if (ia[k] < 0 || ia[k] >= nA || ib[k] < 0 || ib[k] >= nB) return set_error(ctx, YGZB_ERR_INVALID, "pair %d: index out of range", k);
    cudaSetDevice(ctx->device);
    const size_t sa = ((size_t)nA * 32 + 15) & ~(size_t)15, sb = ((size_t)nB * 32 + 15) & ~(size_t)15;
    uint8_t* buf = (uint8_t*)dev_scratch(ctx, 4, sa + sb + (size_t)n * 12);
    if (!buf) return YGZB_ERR_CUDA;
    int32_t* dia = (int32_t*)(buf + sa + sb);
    int32_t* dib = dia + n;
    int32_t* dd = dib + n;
    YGZB_CUDA(ctx, cudaMemcpyAsync(buf, A, (size_t)nA * 32, cudaMemcpyHostToDevice, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(buf + sa, B, (size_t)nB * 32, cudaMemcpyHostToDevice, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(dia, ia, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    YGZB_CUDA(ctx, cudaMemcpyAsync(dib, ib, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    int rc = launch_hamming_pairs(ctx, buf, buf + sa, dia, dib, n, dd);
    if (rc != YGZB_OK) return rc;
    YGZB_CUDA(ctx, cudaMemcpyAsync(dist, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    YGZB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return YGZB_OK;
}

}  // extern "C"
