// vo_driver.cpp -- native (C++) lock-step tracking loop over the C ABI: the host side of BASELINE config C5.
//
// CALLER code in the shape of the reference's src/Module/VisualOdometry.cpp:38-107 (AddFrame), :281-302
// (TrackRefFrame), src/Module/LocalMapping.cpp:24-140 (TrackLocalMap: FindCandidates / ProjectMapPoints /
// OptimizeCurrent), VisualOdometry.cpp:182-218 + :304-321 (SetKeyframe / NeedNewKeyFrame) and LocalMapping.cpp:149-172
// (LocalBA -> ba::LocalBAG2O).  It is the same loop as ygz_slam_b200/vo.py (which the tests run against the CPU oracle
// backend for end-to-end parity); this file is that loop without the Python interpreter in the way, so that the
// measured tracked-frames/s reflect the device path and not numpy call overhead.  No image or optimisation
// arithmetic happens here: every numeric step is ONE batched C-ABI call over all streams (ygzb_frames_upload,
// ygzb_sparse_align, ygzb_project_align, ygzb_pose_only, ygzb_detect, ygzb_local_ba).
//
// Input-side simplifications are those of vo.py: ground-truth depth initialises the map points of a key-frame
// (test/test_feature_alignment.cpp:72-85 does the same with TUM depth), no BoW / loop closing.
//
// Build: host C++ only (g++), links libygz_b200.so; see ygz_slam_b200/build.py (libygz_vo.so).
#include <algorithm>
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ygz_b200.h"
#include "../csrc/se3.cuh"

namespace {

constexpr double FX = 520.9, FY = 521.0, CX = 325.1, CY = 249.7;   // config/default.yaml:32-35 (as Python floats in vo.py)
constexpr int W = 640, H = 480;
constexpr int kLocalKeyframes = 3;                                 // LocalMapping.local_keyframes (default.yaml:68)
constexpr int kSlotsPerStream = kLocalKeyframes + 2;
constexpr int kMinInliers = 30;                                    // vo.keyframe.min_features (default.yaml:66)
constexpr int kSpeculativeFrames = 3;                              // engine: frames tracked ahead once a key-frame may trigger any time

struct Mat34 {
    double m[12];
};
Mat34 identity() { return Mat34{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}}; }
ygzb::SE3d to_se3(const Mat34& T) { return ygzb::se3_from_mat(T.m); }
Mat34 from_se3(const ygzb::SE3d& s) {
    Mat34 T;
    ygzb::se3_to_mat(s, T.m);
    return T;
}
Mat34 mul(const Mat34& A, const Mat34& B) {  // plain matrix product of [R|t] transforms
    Mat34 C;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) C.m[4 * r + c] = A.m[4 * r] * B.m[c] + A.m[4 * r + 1] * B.m[4 + c] + A.m[4 * r + 2] * B.m[8 + c];
        C.m[4 * r + 3] = A.m[4 * r] * B.m[3] + A.m[4 * r + 1] * B.m[7] + A.m[4 * r + 2] * B.m[11] + A.m[4 * r + 3];
    }
    return C;
}
Mat34 inv(const Mat34& A) {
    Mat34 C;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) C.m[4 * r + c] = A.m[4 * c + r];
        C.m[4 * r + 3] = -(A.m[r] * A.m[3] + A.m[4 + r] * A.m[7] + A.m[8 + r] * A.m[11]);
    }
    return C;
}
void se3_log(const Mat34& T, double out[6]) { ygzb::se3_log(to_se3(T), out); }  // [upsilon; omega]

struct Keyframe {
    int slot = 0, frame_id = 0;
    Mat34 T;
    std::vector<double> px;      // 2n full-res pixels of its features
    std::vector<int> level;
    std::vector<double> depth;   // n
    std::vector<double> pw;      // 3n world points
    long mp0 = 0;                // map point ids are [mp0, mp0 + n)
    std::vector<long> obs_id;    // older map points tracked into this frame ...
    std::vector<double> obs_px;  // ... and their measured pixels
    int n() const { return (int)depth.size(); }
};

struct Stream {
    int slot0 = 0;
    std::deque<Keyframe> keyframes;   // at most kLocalKeyframes + 1, the newest is the reference key-frame
    Mat34 T = identity();
    bool has_pose = false, lost = false, has_ref = false, has_last = false;
    int frames_since_kf = 0;
    long next_mp = 0;
    std::vector<long> last_id;
    std::vector<double> last_px;
    long n_keyframes = 0, n_ba = 0, n_candidates = 0, n_projected = 0, n_inliers = 0;
    long ba_obs = 0, ba_pts = 0, ba_kfs = 0, ba_trials = 0, ba_iters = 0;
    double ba_flops = 0;   // SURVEY 8d model: per LM trial 300 n_obs + sum_j (216 k_j^2 + 108 k_j + 50) + dim^3 / 3
    int first_local() const { return std::max(0, (int)keyframes.size() - kLocalKeyframes); }
};

struct Params {
    int kf_min_frames;
    double kf_min_rot, kf_min_trans;
};

#define CHK(call)                                     \
    do {                                              \
        int _rc = (call);                             \
        if (_rc != YGZB_OK) return _rc;               \
    } while (0)

// wall time per C-ABI stage (printed when YGZ_VO_TIMING is set): where a lock-step frame goes
enum { kTUpload, kTSparse, kTProject, kTPoseOnly, kTDetect, kTLocalBA, kTStages };
std::atomic<long long> g_stage_ns[kTStages];   // summed over the host threads
struct StageTimer {
    int stage;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit StageTimer(int s) : stage(s) {}
    ~StageTimer() {
        g_stage_ns[stage].fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(),
                                    std::memory_order_relaxed);
    }
};
#define TIMED(stage, call)        \
    do {                          \
        StageTimer _t(stage);     \
        CHK(call);                \
    } while (0)

class Driver {
  public:
    // slot layout: [0, S) = staging slots of the incoming frames (contiguous: one upload per lock-step frame), then a ring
    // of kSlotsPerStream - 1 key-frame slots per stream; a frame that becomes a key-frame is copied device-to-device
    Driver(ygzb_ctx* ctx, ygzb_frames* fr, int n_streams, const Params& p) : ctx_(ctx), fr_(fr), S_(n_streams), prm_(p), st_(n_streams) {
        for (int i = 0; i < S_; ++i) st_[i].slot0 = S_ + i * (kSlotsPerStream - 1);
        int rows = 0, cols = 0;
        ygzb_grid_dims(ctx, &rows, &cols);
        n_cells_ = rows * cols;
    }
    std::vector<Stream>& streams() { return st_; }
    // bytes this driver hands to / reads back from the C ABI (host buffers), counted from the arrays of every call
    long long h2d_image_bytes = 0, h2d_other_bytes = 0, d2h_bytes = 0;

    // one lock-step frame: images[i] = grey frame of stream i, depth[i] = its (static) ground-truth depth map
    int add_frames(const uint8_t* const* images, const double* const* depth, int frame_id) {
        cur_slot_.resize(S_);
        for (int i = 0; i < S_; ++i) cur_slot_[i] = i;
        // one strided copy when the caller's frames are equally spaced in host memory (a stacked [stream][frame] array)
        bool strided = S_ > 1;
        const ptrdiff_t stride = S_ > 1 ? images[1] - images[0] : 0;
        for (int i = 1; i + 1 < S_ && strided; ++i) strided = images[i + 1] - images[i] == stride;
        h2d_image_bytes += (long long)S_ * W * H;
        if (strided && stride >= (ptrdiff_t)W * H) {
            TIMED(kTUpload, ygzb_frames_upload(fr_, 0, S_, images[0], 1, (size_t)stride));
        } else {
            for (int i = 0; i < S_; ++i) TIMED(kTUpload, ygzb_frames_upload(fr_, i, 1, images[i], 1, (size_t)W * H));
        }
        std::vector<int> boot, track;
        for (int i = 0; i < S_; ++i) {
            if (st_[i].lost) continue;
            if (!st_[i].has_ref) boot.push_back(i);
            else track.push_back(i);
        }
        if (!boot.empty()) {
            for (int i : boot) {
                st_[i].T = identity();
                st_[i].has_pose = true;
            }
            CHK(make_keyframes(boot, depth, frame_id, true));
        }
        if (!track.empty()) CHK(track_frames(track, depth, frame_id));
        return YGZB_OK;
    }

  private:
    int next_slot(int i) const {
        const Stream& s = st_[i];
        for (int c = 0; c < kSlotsPerStream - 1; ++c) {
            const int slot = s.slot0 + c;
            bool used = false;
            for (int k = s.first_local(); k < (int)s.keyframes.size(); ++k) used |= s.keyframes[k].slot == slot;
            if (!used) return slot;
        }
        return s.slot0;  // cannot happen: the ring has kLocalKeyframes + 1 slots
    }

    // TrackRefFrame + TrackLocalMap + key-frame decision for the streams in idx
    int track_frames(std::vector<int> idx, const double* const* depth, int frame_id) {
        const int n = (int)idx.size();
        // -- Matcher::SparseImageAlignment(ref, cur) with cur._TCW = ref._TCW (VisualOdometry.cpp:281-302)
        std::vector<int32_t> ref_slot(n), cur_slot(n), offs(n + 1, 0), n_meas(n);
        std::vector<double> px, dep, T_ref(12 * (size_t)n), T_cur(12 * (size_t)n);
        for (int j = 0; j < n; ++j) {
            const Keyframe& kf = st_[idx[j]].keyframes.back();
            ref_slot[j] = kf.slot;
            cur_slot[j] = cur_slot_[idx[j]];
            offs[j + 1] = offs[j] + kf.n();
            px.insert(px.end(), kf.px.begin(), kf.px.end());
            dep.insert(dep.end(), kf.depth.begin(), kf.depth.end());
            std::memcpy(&T_ref[12 * (size_t)j], kf.T.m, sizeof(kf.T.m));
        }
        T_cur = T_ref;
        std::vector<uint8_t> has(px.size() / 2, 1);
        h2d_other_bytes += 8ll * n + 4ll * (n + 1) + 25ll * (long long)has.size() + 192ll * n;
        d2h_bytes += 100ll * n;
        TIMED(kTSparse, ygzb_sparse_align(fr_, n, ref_slot.data(), cur_slot.data(), offs.data(), px.data(), dep.data(), has.data(), T_ref.data(),
                              T_cur.data(), 2, 0, 30, 1e-6, n_meas.data(), nullptr));
        std::vector<int> alive;
        std::vector<Mat34> T_of(S_);
        for (int j = 0; j < n; ++j) {
            Mat34 Tc, Tr;
            std::memcpy(Tc.m, &T_cur[12 * (size_t)j], sizeof(Tc.m));
            std::memcpy(Tr.m, &T_ref[12 * (size_t)j], sizeof(Tr.m));
            double lg[6];
            se3_log(mul(Tc, inv(Tr)), lg);   // Matcher.cpp:482-488: motion norm check
            double nrm = 0;
            for (double v : lg) nrm += v * v;
            if (std::sqrt(nrm) <= 0.2) {
                T_of[idx[j]] = Tc;
                alive.push_back(idx[j]);
            } else {
                st_[idx[j]].lost = true;   // the reference keeps the last pose and reports VO_LOST
            }
        }
        idx = alive;
        if (idx.empty()) return YGZB_OK;

        // -- FindCandidates: project the points of the local key-frames, border 20 (LocalMapping.cpp:47-80), then
        //    ProjectMapPoints: Matcher::FindDirectProjection per candidate (:82-111) as ONE batch.  Poses go in relative
        //    to the reference key-frame (I, T_cur * T_ref^-1): GetWarpAffineMatrix is only correct for an identity
        //    reference pose (Matcher.cpp:425-430, quirk kept in the kernel).
        struct Cand { int stream, kf, n; };
        size_t bound = 0;   // every point of every local key-frame can become a candidate
        for (int i : idx)
            for (int k = st_[i].first_local(); k < (int)st_[i].keyframes.size(); ++k) bound += (size_t)st_[i].keyframes[k].n();
        std::vector<Cand> cand(bound);
        std::vector<int32_t> c_ref_slot(bound), c_cur_slot(bound), c_ref_pose(bound), c_cur_pose(bound);
        const Mat34 eye = identity();
        std::vector<double> poses(eye.m, eye.m + 12), c_ref_px(2 * bound), c_ref_depth(bound), c_cur_px(2 * bound);
        std::vector<uint8_t> c_level(bound);
        std::vector<int> job_begin;
        size_t nc_sz = 0;
        for (int i : idx) {
            Stream& s = st_[i];
            const Mat34& T = T_of[i];
            job_begin.push_back((int)nc_sz);
            const int pose_base = (int)(poses.size() / 12);
            for (int k = s.first_local(); k < (int)s.keyframes.size(); ++k) {
                const Mat34 rel = mul(T, inv(s.keyframes[k].T));
                poses.insert(poses.end(), rel.m, rel.m + 12);
            }
            for (int k = s.first_local(); k < (int)s.keyframes.size(); ++k) {
                const Keyframe& kf = s.keyframes[k];
                const int32_t cur = cur_slot_[i], pose_id = pose_base + (k - s.first_local());
                for (int g = 0; g < kf.n(); ++g) {
                    const double* X = &kf.pw[3 * (size_t)g];
                    const double x = T.m[0] * X[0] + T.m[1] * X[1] + T.m[2] * X[2] + T.m[3];
                    const double y = T.m[4] * X[0] + T.m[5] * X[1] + T.m[6] * X[2] + T.m[7];
                    const double z = T.m[8] * X[0] + T.m[9] * X[1] + T.m[10] * X[2] + T.m[11];
                    const double u = FX * x / z + CX, v = FY * y / z + CY;
                    if (!(z > 0 && u >= 20 && u < W - 20 && v >= 20 && v < H - 20)) continue;
                    const size_t c = nc_sz++;
                    cand[c] = {i, k, g};
                    c_ref_slot[c] = kf.slot;
                    c_cur_slot[c] = cur;
                    c_ref_pose[c] = 0;
                    c_cur_pose[c] = pose_id;
                    c_ref_px[2 * c] = kf.px[2 * (size_t)g];
                    c_ref_px[2 * c + 1] = kf.px[2 * (size_t)g + 1];
                    c_ref_depth[c] = kf.depth[g];
                    c_level[c] = (uint8_t)kf.level[g];
                    c_cur_px[2 * c] = u;
                    c_cur_px[2 * c + 1] = v;
                }
            }
            s.n_candidates += (long)nc_sz - job_begin.back();
        }
        job_begin.push_back((int)nc_sz);
        const int nc = (int)nc_sz;
        std::vector<uint8_t> search_level(nc ? nc : 1), ok(nc ? nc : 1);
        h2d_other_bytes += 57ll * nc + 8ll * (long long)poses.size();
        d2h_bytes += 18ll * nc;
        if (nc)
            TIMED(kTProject, ygzb_project_align(fr_, nc, c_ref_slot.data(), c_cur_slot.data(), (int)(poses.size() / 12), poses.data(), c_ref_pose.data(),
                                   c_cur_pose.data(), c_ref_px.data(), c_ref_depth.data(), c_level.data(), c_cur_px.data(),
                                   search_level.data(), ok.data()));

        // -- ba::OptimizeCurrentPoseOnly on the successfully projected points (LocalMapping.cpp:126; BA.cpp:188-264)
        const int m = (int)idx.size();
        std::vector<int32_t> po(m + 1, 0), n_inl(m);
        std::vector<double> pw(3 * (size_t)nc + 3), obs(2 * (size_t)nc + 2), Tp(12 * (size_t)m);
        std::vector<long> obs_id((size_t)nc + 1);
        size_t q_out = 0;
        for (int j = 0; j < m; ++j) {
            Stream& s = st_[idx[j]];
            int cnt = 0;
            for (int c = job_begin[j]; c < job_begin[j + 1]; ++c) {
                if (!ok[c]) continue;
                const Keyframe& kf = s.keyframes[cand[c].kf];
                std::memcpy(&pw[3 * q_out], &kf.pw[3 * (size_t)cand[c].n], 3 * sizeof(double));
                obs[2 * q_out] = c_cur_px[2 * (size_t)c];
                obs[2 * q_out + 1] = c_cur_px[2 * (size_t)c + 1];
                obs_id[q_out] = kf.mp0 + cand[c].n;
                ++q_out;
                ++cnt;
            }
            po[j + 1] = po[j] + cnt;
            s.n_projected += cnt;
            std::memcpy(&Tp[12 * (size_t)j], T_of[idx[j]].m, sizeof(Mat34));
        }
        const size_t tot = (size_t)po[m];
        std::vector<uint8_t> inl(tot ? tot : 1);
        std::vector<double> dep_out(tot ? tot : 1);
        h2d_other_bytes += 4ll * (m + 1) + 40ll * (long long)tot + 96ll * m;
        d2h_bytes += 96ll * m + 9ll * (long long)tot + 4ll * m;
        TIMED(kTPoseOnly, ygzb_pose_only(ctx_, m, po.data(), pw.data(), obs.data(), Tp.data(), inl.data(), dep_out.data(),
                           n_inl.data()));
        std::vector<int> need;
        for (int j = 0; j < m; ++j) {
            Stream& s = st_[idx[j]];
            if (n_inl[j] < kMinInliers) {
                s.lost = true;
                continue;
            }
            s.last_id.clear();
            s.last_px.clear();
            for (int q = po[j]; q < po[j + 1]; ++q)
                if (inl[q]) {
                    s.last_id.push_back(obs_id[q]);
                    s.last_px.push_back(obs[2 * (size_t)q]);
                    s.last_px.push_back(obs[2 * (size_t)q + 1]);
                }
            s.has_last = true;
            std::memcpy(s.T.m, &Tp[12 * (size_t)j], sizeof(Mat34));
            s.frames_since_kf += 1;
            s.n_inliers += n_inl[j];
            // NeedNewKeyFrame (VisualOdometry.cpp:304-321)
            if (s.frames_since_kf < prm_.kf_min_frames) continue;
            double d[6];
            se3_log(mul(s.T, inv(s.keyframes.back().T)), d);
            const double rot = std::sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]), tr = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            if (rot > prm_.kf_min_rot || tr > prm_.kf_min_trans) need.push_back(idx[j]);
        }
        if (!need.empty()) CHK(make_keyframes(need, depth, frame_id, false));
        return YGZB_OK;
    }

    // SetKeyframe: Detect (grid FAST + ORB), depth-initialised map points, local BA (VisualOdometry.cpp:182-218)
    int make_keyframes(const std::vector<int>& idx, const double* const* depth, int frame_id, bool fresh) {
        const int n = (int)idx.size();
        std::vector<int32_t> slots(n), off(n + 1);
        for (int j = 0; j < n; ++j) slots[j] = cur_slot_[idx[j]];
        const size_t cap = (size_t)n * n_cells_;
        kx_.resize(cap); ky_.resize(cap); klevel_.resize(cap); kscore_.resize(cap); kangle_.resize(cap); kdesc_.resize(cap * 32);
        ygzb_keypoints kp{off.data(), kx_.data(), ky_.data(), klevel_.data(), kscore_.data(), kangle_.data(), kdesc_.data(), nullptr, (int)cap};
        TIMED(kTDetect, ygzb_detect(fr_, slots.data(), n, nullptr, &kp));
        h2d_other_bytes += 4ll * n;
        d2h_bytes += 4ll * (n + 1) + 49ll * off[n];
        std::vector<int> ba_jobs;
        for (int j = 0; j < n; ++j) {
            Stream& s = st_[idx[j]];
            Keyframe kf;
            kf.slot = next_slot(idx[j]);   // the frame leaves its staging slot: keep its pyramid in the stream's key-frame ring
            TIMED(kTDetect, ygzb_frames_copy(fr_, slots[j], kf.slot));
            kf.frame_id = frame_id;
            kf.T = s.T;
            const Mat34 Tin = inv(s.T);
            const int cnt = off[j + 1] - off[j];
            kf.px.resize(2 * (size_t)cnt); kf.level.resize(cnt); kf.depth.resize(cnt); kf.pw.resize(3 * (size_t)cnt);
            for (int g = 0; g < cnt; ++g) {
                const double x = kx_[off[j] + g], y = ky_[off[j] + g];
                const double d = depth[idx[j]][(size_t)(int)y * W + (int)x];
                kf.px[2 * (size_t)g] = x;
                kf.px[2 * (size_t)g + 1] = y;
                kf.level[g] = klevel_[off[j] + g];
                kf.depth[g] = d;
                const double pc[3] = {(x - CX) * d / FX, (y - CY) * d / FY, d};
                for (int r = 0; r < 3; ++r)
                    kf.pw[3 * (size_t)g + r] = Tin.m[4 * r] * pc[0] + Tin.m[4 * r + 1] * pc[1] + Tin.m[4 * r + 2] * pc[2] + Tin.m[4 * r + 3];
            }
            kf.mp0 = s.next_mp;
            s.next_mp += cnt;
            if (!fresh && s.has_last) {
                kf.obs_id = s.last_id;
                kf.obs_px = s.last_px;
            }
            s.keyframes.push_back(std::move(kf));
            while ((int)s.keyframes.size() > kLocalKeyframes + 1) s.keyframes.pop_front();
            s.has_ref = true;
            s.frames_since_kf = 0;
            s.n_keyframes += 1;
            if (!fresh && s.keyframes.size() >= 2) ba_jobs.push_back(idx[j]);
        }
        if (!ba_jobs.empty()) CHK(local_ba(ba_jobs));
        return YGZB_OK;
    }

    // LocalMapping::LocalBA -> ba::LocalBAG2O over the local key-frames and the points at least two of them observe
    int local_ba(const std::vector<int>& idx) {
        const int P = (int)idx.size();
        std::vector<int32_t> kf_off(P + 1, 0), pt_off(P + 1, 0), ob_off(P + 1, 0), kf_idx, pt_idx;
        std::vector<double> poses, pts, obs;
        std::vector<uint8_t> fixed;
        struct Ref { int kf, n; };
        std::vector<std::vector<Ref>> owners(P);
        for (int p = 0; p < P; ++p) {
            Stream& s = st_[idx[p]];
            const int k0 = s.first_local(), nk = (int)s.keyframes.size() - k0;
            for (int k = 0; k < nk; ++k) {
                double lg[6];
                se3_log(s.keyframes[k0 + k].T, lg);
                const double g2o[6] = {lg[3], lg[4], lg[5], lg[0], lg[1], lg[2]};   // VertexSE3Sophus: [omega; upsilon]
                poses.insert(poses.end(), g2o, g2o + 6);
                fixed.push_back(k == 0);   // the oldest local key-frame fixes the gauge (key-frame 0 in the reference)
            }
            // observations: a key-frame observes its own points (detected pixel) and the older points tracked into it
            struct Ob { long id; int kf; double u, v; };
            std::vector<Ob> all;
            auto in_local = [&](long id) {
                for (int k = 0; k < nk; ++k) {
                    const Keyframe& kf = s.keyframes[k0 + k];
                    if (id >= kf.mp0 && id < kf.mp0 + kf.n()) return true;
                }
                return false;
            };
            for (int k = 0; k < nk; ++k) {
                const Keyframe& kf = s.keyframes[k0 + k];
                for (int g = 0; g < kf.n(); ++g) all.push_back({kf.mp0 + g, k, kf.px[2 * (size_t)g], kf.px[2 * (size_t)g + 1]});
                for (size_t q = 0; q < kf.obs_id.size(); ++q)
                    if (in_local(kf.obs_id[q])) all.push_back({kf.obs_id[q], k, kf.obs_px[2 * q], kf.obs_px[2 * q + 1]});
            }
            // points seen by a single key-frame do not constrain anything: keep ids with >= 2 observations, numbered in
            // ascending id order (np.unique in vo.py)
            std::vector<long> ids;
            ids.reserve(all.size());
            for (const Ob& o : all) ids.push_back(o.id);
            std::sort(ids.begin(), ids.end());
            std::vector<long> multi;
            for (size_t a = 0; a < ids.size();) {
                size_t b = a;
                while (b < ids.size() && ids[b] == ids[a]) ++b;
                if (b - a >= 2) multi.push_back(ids[a]);
                a = b;
            }
            for (long id : multi) {
                for (int k = 0; k < nk; ++k) {
                    const Keyframe& kf = s.keyframes[k0 + k];
                    if (id >= kf.mp0 && id < kf.mp0 + kf.n()) {
                        const int g = (int)(id - kf.mp0);
                        owners[p].push_back({k0 + k, g});
                        pts.insert(pts.end(), &kf.pw[3 * (size_t)g], &kf.pw[3 * (size_t)g] + 3);
                        break;
                    }
                }
            }
            int n_ob = 0;
            for (const Ob& o : all) {
                const auto it = std::lower_bound(multi.begin(), multi.end(), o.id);
                if (it == multi.end() || *it != o.id) continue;
                kf_idx.push_back(o.kf);
                pt_idx.push_back((int32_t)(it - multi.begin()));
                obs.push_back(o.u);
                obs.push_back(o.v);
                ++n_ob;
            }
            kf_off[p + 1] = kf_off[p] + nk;
            pt_off[p + 1] = pt_off[p] + (int)multi.size();
            ob_off[p + 1] = ob_off[p] + n_ob;
        }
        ygzb_ba_params bp;
        ygzb_default_ba_params(&bp);
        std::vector<uint8_t> outl(obs.size() / 2 + 1);
        std::vector<ygzb_ba_stats> bst(P);
        h2d_other_bytes += 12ll * (P + 1) + 49ll * (long long)fixed.size() + 24ll * (long long)(pts.size() / 3) + 24ll * (long long)kf_idx.size();
        d2h_bytes += 48ll * (long long)fixed.size() + 24ll * (long long)(pts.size() / 3) + (long long)kf_idx.size();
        static const double zero3[3] = {0, 0, 0};
        static const int32_t zero_i = 0;
        TIMED(kTLocalBA, ygzb_local_ba(ctx_, P, kf_off.data(), pt_off.data(), ob_off.data(), poses.data(), fixed.data(), pts.empty() ? const_cast<double*>(zero3) : pts.data(),
                          kf_idx.empty() ? &zero_i : kf_idx.data(), pt_idx.empty() ? &zero_i : pt_idx.data(), obs.empty() ? zero3 : obs.data(), &bp,
                          outl.data(), bst.data()));
        for (int p = 0; p < P; ++p) {
            Stream& s = st_[idx[p]];
            const int k0 = s.first_local(), nk = (int)s.keyframes.size() - k0;
            {   // problem sizes and the FLOP model of SURVEY 8d (for the roofline of the BA kernel)
                const int no = ob_off[p + 1] - ob_off[p], npt = pt_off[p + 1] - pt_off[p];
                std::vector<int> deg(npt, 0);
                for (int o = ob_off[p]; o < ob_off[p + 1]; ++o) deg[pt_idx[o]]++;
                double per_trial = 300.0 * no;
                for (int d : deg) per_trial += 216.0 * d * d + 108.0 * d + 50.0;
                const double dim = 6.0 * (nk - 1);
                per_trial += dim * dim * dim / 3.0;
                s.ba_obs += no; s.ba_pts += npt; s.ba_kfs += nk; s.ba_trials += bst[p].lm_trials; s.ba_iters += bst[p].iters;
                s.ba_flops += per_trial * bst[p].lm_trials;
            }
            for (int k = 0; k < nk; ++k) {
                const double* g = &poses[6 * (size_t)(kf_off[p] + k)];
                const double v[6] = {g[3], g[4], g[5], g[0], g[1], g[2]};
                s.keyframes[k0 + k].T = from_se3(ygzb::se3_exp(v));
            }
            for (size_t q = 0; q < owners[p].size(); ++q) {
                Keyframe& kf = s.keyframes[owners[p][q].kf];
                std::memcpy(&kf.pw[3 * (size_t)owners[p][q].n], &pts[3 * ((size_t)pt_off[p] + q)], 3 * sizeof(double));
            }
            s.T = s.keyframes.back().T;
            s.n_ba += 1;
        }
        return YGZB_OK;
    }

    ygzb_ctx* ctx_;
    ygzb_frames* fr_;
    int S_, n_cells_ = 0;
    Params prm_;
    std::vector<Stream> st_;
    std::vector<int> cur_slot_;
    std::vector<float> kx_, ky_, kscore_, kangle_;
    std::vector<uint8_t> klevel_, kdesc_;
};


// ---- device-resident engine --------------------------------------------------------------------------------------------
// The same caller logic on ygzb_tracker_*: the local map lives on the device, a ROUND enqueues for every stream a window of
// consecutive frames -- up to and including the first frame that may become a key-frame (a frame is tracked against the
// reference key-frame, never against its predecessor: VisualOdometry.cpp:66, so the frames of a window are independent) --
// as ONE fused chain (upload + pyramid, sparse alignment, candidate projection, direct projection, pose-only), reads one
// 128-byte record per frame back, takes the key-frame decisions, and inserts the key-frames of the round (Detect, map
// points, local BA, all on the device) with one more enqueue.  Results are identical to frame-by-frame processing.
struct KfInfo {
    int entry = 0, n = 0, frame_id = 0;
    Mat34 T;
    long mp0 = 0;
};
struct EStream {
    std::deque<KfInfo> kfs;   // at most YGZB_TRACK_RING, the newest is the reference key-frame; the last kLocalKeyframes are local
    Mat34 T = identity();
    bool has_pose = false, lost = false;
    int frames_since_kf = 0, next_frame = 0;
    long next_mp = 0;
    long n_keyframes = 0, n_ba = 0, n_candidates = 0, n_projected = 0, n_inliers = 0;
    long ba_obs = 0, ba_pts = 0, ba_kfs = 0, ba_trials = 0, ba_iters = 0;
    double ba_flops = 0;
};

class Engine {
    struct Win { int stream, first, count, job0; };   // a window in flight: frames [first, first + count) of a stream (job0 < 0: first frame)
  public:
    Engine(ygzb_ctx* ctx, int n_streams, int window, const Params& p) : ctx_(ctx), S_(n_streams), F_(std::max(1, window)), prm_(p), st_(n_streams) {
        // YGZ_VO_BLOCKING_SYNC=1: sleep instead of spinning in the one synchronisation per round (hosts with fewer CPUs than
        // engine threads; bench.py sets it when the threads of all ranks outnumber the CPUs it may use)
        const char* e = std::getenv("YGZ_VO_BLOCKING_SYNC");
        blocking_sync_ = e && std::atoi(e) != 0;
    }
    ~Engine() {
        if (tr_) ygzb_tracker_destroy(tr_);
        if (fr_) ygzb_frames_destroy(fr_);
        if (h_res_) ygzb_host_free(h_res_);
        if (h_kres_) ygzb_host_free(h_kres_);
    }
    int init(const double* const* depth) {
        CHK(ygzb_frames_create(ctx_, S_ * F_ + S_ * YGZB_TRACK_RING, &fr_));
        const double K[4] = {FX, FY, CX, CY};
        CHK(ygzb_tracker_create(fr_, S_, S_ * F_, K, &tr_));
        for (int i = 0; i < S_; ++i) CHK(ygzb_tracker_set_depth(tr_, i, depth[i]));
        void* p = nullptr;
        CHK(ygzb_host_alloc(&p, sizeof(ygzb_track_result) * (size_t)S_ * F_));
        h_res_ = static_cast<ygzb_track_result*>(p);
        CHK(ygzb_host_alloc(&p, sizeof(ygzb_keyframe_result) * (size_t)S_));
        h_kres_ = static_cast<ygzb_keyframe_result*>(p);
        ygzb_default_ba_params(&ba_);
        return YGZB_OK;
    }
    std::vector<EStream>& streams() { return st_; }
    bool all_reached(int frame) const {
        for (const EStream& s : st_)
            if (s.next_frame < frame) return false;
        return true;
    }
    long long h2d_image_bytes = 0, h2d_other_bytes = 0, d2h_bytes = 0;

    // Runs until every stream has reached frame `limit` (no window crosses it) and nothing is in flight.  The loop is
    // software-pipelined around ONE host synchronisation per round:
    //   1. take the decisions of the tracking batch that has just come back (lost streams, key-frame triggers),
    //   2. enqueue the key-frame insertions of this round (Detect, map points, local BA) -- asynchronous,
    //   3. plan the next window of every stream, upload its frames and enqueue its tracking chain right behind; the tracker
    //      runs the uploads and the sparse alignment of that batch on a second CUDA stream, concurrently with the local BA
    //      of step 2 (the alignment works relative to the reference key-frame and needs its features, not its refined pose),
    //   4. synchronise, apply the key-frame results (poses after the BA, feature counts).
    int run_until(const uint8_t* const* images, int n_frames, int limit, double* traj /* this group's [S][n_frames][12] */) {
        for (;;) {
            std::vector<ygzb_keyframe_job> kjobs;
            std::vector<int> kframe;
            // ---- 1. results of the batch in flight
            for (const Win& b : wins_) {
                EStream& s = st_[b.stream];
                if (b.job0 < 0) {   // first frame of the stream: becomes the first key-frame (depth-initialised map)
                    s.T = identity();
                    s.has_pose = true;
                    kjobs.push_back(make_kf_job(b.stream, b.stream * F_, -1));
                    kframe.push_back(b.first);
                    s.next_frame = b.first + 1;
                    continue;
                }
                int t = 0;
                for (; t < b.count; ++t) {
                    const ygzb_track_result& r = h_res_[b.job0 + t];
                    if (!r.aligned) {   // Matcher::SparseImageAlignment returned false (Matcher.cpp:482-488)
                        s.lost = true;
                        break;
                    }
                    s.n_candidates += r.n_candidates;
                    s.n_projected += r.n_projected;
                    if (r.n_inliers < kMinInliers) {
                        s.lost = true;
                        break;
                    }
                    std::memcpy(s.T.m, r.T_cw, sizeof(s.T.m));
                    s.frames_since_kf += 1;
                    s.n_inliers += r.n_inliers;
                    put_pose(traj, b.stream, n_frames, b.first + t, s);
                    // NeedNewKeyFrame (VisualOdometry.cpp:304-321)
                    if (s.frames_since_kf < prm_.kf_min_frames) continue;
                    double d[6];
                    se3_log(mul(s.T, inv(s.kfs.back().T)), d);
                    const double rot = std::sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]), tr = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    if (rot > prm_.kf_min_rot || tr > prm_.kf_min_trans) {
                        kjobs.push_back(make_kf_job(b.stream, b.stream * F_ + t, b.job0 + t));
                        kframe.push_back(b.first + t);
                        ++t;
                        break;   // frames of the window behind the key-frame (speculative ones) are dropped: tracked again next round
                    }
                }
                if (s.lost) {   // the reference keeps the last pose and reports VO_LOST
                    put_pose(traj, b.stream, n_frames, b.first + t, s);
                    s.next_frame = b.first + t + 1;
                } else {
                    s.next_frame = b.first + t;
                }
            }
            wins_.clear();
            // ---- 2. key-frame insertions of this round (asynchronous)
            if (!kjobs.empty()) {
                StageTimer tm(kTLocalBA);
                for (size_t q = 0; q < kjobs.size(); ++q) {   // bookkeeping that does not need the device's answer
                    EStream& s = st_[kjobs[q].stream];
                    KfInfo kf;
                    kf.entry = kjobs[q].entry;
                    kf.frame_id = kframe[q];
                    kf.mp0 = kjobs[q].mp0;
                    kf.T = s.T;
                    s.kfs.push_back(kf);
                    while ((int)s.kfs.size() > YGZB_TRACK_RING) s.kfs.pop_front();
                    s.frames_since_kf = 0;
                    s.n_keyframes += 1;
                }
                CHK(ygzb_tracker_make_keyframes(tr_, (int)kjobs.size(), kjobs.data(), &ba_, h_kres_));
                h2d_other_bytes += (long long)(kjobs.size() * (sizeof(ygzb_keyframe_job) + 4));
                d2h_bytes += (long long)(kjobs.size() * sizeof(ygzb_keyframe_result));
            }
            // ---- 3. next window of every stream: uploads + tracking chain (asynchronous, right behind the key-frames)
            std::vector<ygzb_track_job> jobs;
            for (int i = 0; i < S_; ++i) {
                EStream& s = st_[i];
                if (s.next_frame >= limit) continue;
                if (s.lost) {
                    for (int k = s.next_frame; k < limit; ++k) put_pose(traj, i, n_frames, k, s);
                    s.next_frame = limit;
                    continue;
                }
                // window: up to and including the first frame that may become a key-frame; past that point every frame may, and a
                // few frames are tracked speculatively -- the ones behind a key-frame trigger are dropped and tracked again against
                // the new key-frame next round (results stay those of frame-by-frame processing)
                int w = 1;
                if (!s.kfs.empty()) {
                    const int sure = prm_.kf_min_frames - s.frames_since_kf;
                    w = sure >= 1 ? std::min(F_, sure) : std::min(F_, kSpeculativeFrames);
                }
                w = std::min(w, limit - s.next_frame);
                CHK(ygzb_tracker_upload(tr_, i * F_, w, images[i] + (size_t)s.next_frame * W * H, (size_t)W * H));
                h2d_image_bytes += (long long)w * W * H;
                if (s.kfs.empty()) {
                    wins_.push_back({i, s.next_frame, 1, -1});
                    continue;
                }
                wins_.push_back({i, s.next_frame, w, (int)jobs.size()});
                const int nl = std::min(kLocalKeyframes, (int)s.kfs.size());
                for (int t = 0; t < w; ++t) {
                    ygzb_track_job j{};
                    j.stream = i;
                    j.cur_slot = i * F_ + t;
                    j.n_local = nl;
                    for (int k = 0; k < nl; ++k) j.entry[k] = s.kfs[s.kfs.size() - nl + k].entry;
                    jobs.push_back(j);
                }
            }
            if (!jobs.empty()) {
                StageTimer tm(kTSparse);
                CHK(ygzb_tracker_track(tr_, (int)jobs.size(), jobs.data(), h_res_));
                h2d_other_bytes += (long long)(jobs.size() * sizeof(ygzb_track_job));
                d2h_bytes += (long long)(jobs.size() * sizeof(ygzb_track_result));
            }
            if (kjobs.empty() && wins_.empty()) return YGZB_OK;   // every stream is at `limit`, nothing in flight
            // ---- 4. one synchronisation per round; key-frame results
            {
                StageTimer tm(kTPoseOnly);
                CHK(blocking_sync_ ? ygzb_synchronize_blocking(ctx_) : ygzb_synchronize(ctx_));
            }
            for (size_t q = 0; q < kjobs.size(); ++q) {
                const ygzb_keyframe_job& kj = kjobs[q];
                const ygzb_keyframe_result& r = h_kres_[q];
                EStream& s = st_[kj.stream];
                s.kfs.back().n = r.n_features;
                s.next_mp = kj.mp0 + r.n_features;
                for (int k = 0; k < kj.n_local; ++k) std::memcpy(s.kfs[s.kfs.size() - kj.n_local + k].T.m, r.T_cw[k], sizeof(Mat34));
                s.T = s.kfs.back().T;
                if (kj.run_ba && kj.n_local >= 2) {
                    s.n_ba += 1;
                    s.ba_obs += r.ba_observations; s.ba_pts += r.ba_points; s.ba_kfs += kj.n_local; s.ba_trials += r.ba_trials; s.ba_iters += r.ba_iters;
                    const double kbar = r.ba_points ? (double)r.ba_observations / r.ba_points : 0.0, dim = 6.0 * (kj.n_local - 1);
                    s.ba_flops += r.ba_trials * (300.0 * r.ba_observations + r.ba_points * (216.0 * kbar * kbar + 108.0 * kbar + 50.0) + dim * dim * dim / 3.0);
                }
                put_pose(traj, kj.stream, n_frames, kframe[q], s);
            }
        }
    }

  private:
    void put_pose(double* traj, int stream, int n_frames, int frame, const EStream& s) const {
        if (!traj) return;
        double* out = traj + ((size_t)stream * n_frames + frame) * 12;
        for (int c = 0; c < 12; ++c) out[c] = s.has_pose ? s.T.m[c] : NAN;
    }
    ygzb_keyframe_job make_kf_job(int stream, int frame_slot, int track_job) const {
        const EStream& s = st_[stream];
        ygzb_keyframe_job kj{};
        kj.stream = stream;
        kj.frame_slot = frame_slot;
        kj.track_job = track_job;
        // ring entry: one that none of the key-frames that stay local uses (the oldest of a full ring leaves the map)
        const int keep = std::min(kLocalKeyframes, (int)s.kfs.size());
        for (int e = 0; e < YGZB_TRACK_RING; ++e) {
            bool used = false;
            for (int k = 0; k < keep; ++k) used |= s.kfs[s.kfs.size() - 1 - k].entry == e;
            if (!used) {
                kj.entry = e;
                break;
            }
        }
        kj.kf_slot = S_ * F_ + stream * YGZB_TRACK_RING + kj.entry;
        const int nl = std::min(kLocalKeyframes, (int)s.kfs.size() + 1);
        kj.n_local = nl;
        for (int k = 0; k < nl - 1; ++k) kj.local_entry[k] = s.kfs[s.kfs.size() - (nl - 1) + k].entry;
        kj.local_entry[nl - 1] = kj.entry;
        kj.run_ba = (track_job >= 0 && nl >= 2) ? 1 : 0;
        kj.mp0 = s.next_mp;
        return kj;
    }

    ygzb_ctx* ctx_;
    ygzb_frames* fr_ = nullptr;
    ygzb_tracker* tr_ = nullptr;
    int S_, F_;
    Params prm_;
    std::vector<EStream> st_;
    ygzb_track_result* h_res_ = nullptr;
    ygzb_keyframe_result* h_kres_ = nullptr;
    ygzb_ba_params ba_;
    std::vector<Win> wins_;
    bool blocking_sync_ = false;
};

}  // namespace

extern "C" {

// PER-STAGE PATH: every numeric step is one blocking C-ABI call with host buffers (the reference's call granularity);
// kept as the cross-check of the device-resident engine below (tests/test_vo.py) and as a diagnostic of bench.py.
// Tracks n_streams independent 640x480 grey streams in lock step for n_frames frames.  The streams are split over
// n_threads host threads; thread 0 drives the caller's context, every further thread creates its own context (= its
// own CUDA stream) on the same device with the same parameters, so the kernels of one group overlap the host work and
// the kernels of the others.  The contexts must use the 3-level pyramid of the reference default.
//   images[s] : n_frames * 480 * 640 bytes, depth[s] : 480 * 640 doubles (static ground-truth depth of stream s)
//   traj      : n_streams * n_frames * 12 doubles out (T_cw after every frame; NaN while a stream has no pose)
//   stats     : n_streams * 16 out: lost, keyframes, local BAs, candidates, projected, inliers, BA observations, BA points,
//               BA key-frames, BA LM trials, BA iterations, BA model FLOP, 0...
//   totals    : (may be NULL) 8 out, timed region only, summed over the host threads: kernel launches, image H2D bytes,
//               other H2D bytes, D2H bytes, 0...
//   seconds   : wall time of frames [warm, n_frames) including the final device synchronisation (all threads meet at a
//               barrier before frame `warm` and after the last frame)
//   device_ms : (may be NULL) the same region timed with CUDA events on the caller's context stream: first event after
//               the warm-up barrier, second one after every thread has synchronised its stream
int ygz_vo_run_stages(ygzb_ctx* ctx, int device, const ygzb_params* params, int n_threads, int n_streams, int n_frames,
               const uint8_t* const* images, const double* const* depth, int kf_min_frames, double kf_min_rot, double kf_min_trans,
               int warm, double* traj, int64_t* stats, double* seconds, double* device_ms, int64_t* totals) {
    if (!ctx || !params || n_streams < 1 || n_frames < 1 || !images || !depth || !traj || !stats || !seconds) return YGZB_ERR_INVALID;
    n_threads = std::max(1, std::min(n_threads, n_streams));
    warm = std::max(0, std::min(warm, n_frames - 1));
    for (auto& v : g_stage_ns) v.store(0);
    std::vector<int> rcs(n_threads, YGZB_OK);
    std::vector<std::vector<long long>> tot(n_threads, std::vector<long long>(4, 0));
    std::barrier sync_point(n_threads);
    std::chrono::steady_clock::time_point t_begin, t_end;
    auto worker = [&](int t) {
        const int s0 = (int)((long)n_streams * t / n_threads), s1 = (int)((long)n_streams * (t + 1) / n_threads), ns = s1 - s0;
        ygzb_ctx* my = ctx;
        int rc = YGZB_OK;
        if (t > 0) rc = ygzb_create(device, params, &my);
        ygzb_frames* fr = nullptr;
        if (rc == YGZB_OK) rc = ygzb_frames_create(my, ns * kSlotsPerStream, &fr);
        Driver drv(my, fr, ns, Params{kf_min_frames, kf_min_rot, kf_min_trans});
        std::vector<const uint8_t*> img(ns);
        for (int k = 0; k < n_frames; ++k) {
            if (k == warm) {
                tot[t][0] = -ygzb_launch_count(my);
                drv.h2d_image_bytes = drv.h2d_other_bytes = drv.d2h_bytes = 0;
                if (rc == YGZB_OK) ygzb_synchronize(my);
                sync_point.arrive_and_wait();
                if (t == 0) {
                    t_begin = std::chrono::steady_clock::now();
                    for (auto& v : g_stage_ns) v.store(0);
                    ygzb_timer_start(ctx);   // device time of the timed region: CUDA events on the caller's context stream
                }
            }
            if (rc != YGZB_OK) continue;   // keep meeting the barriers
            for (int s = 0; s < ns; ++s) img[s] = images[s0 + s] + (size_t)k * W * H;
            rc = drv.add_frames(img.data(), depth + s0, k);
            for (int s = 0; s < ns; ++s) {
                double* out = traj + ((size_t)(s0 + s) * n_frames + k) * 12;
                const Stream& st = drv.streams()[s];
                for (int c = 0; c < 12; ++c) out[c] = st.has_pose ? st.T.m[c] : NAN;
            }
        }
        if (rc == YGZB_OK) ygzb_synchronize(my);
        tot[t][0] += ygzb_launch_count(my);
        tot[t][1] = drv.h2d_image_bytes; tot[t][2] = drv.h2d_other_bytes; tot[t][3] = drv.d2h_bytes;
        sync_point.arrive_and_wait();
        if (t == 0) {
            double ms = 0;
            if (ygzb_timer_stop(ctx, &ms) == YGZB_OK && device_ms) *device_ms = ms;   // every thread has synchronised its stream
            t_end = std::chrono::steady_clock::now();
        }
        for (int s = 0; s < ns; ++s) {
            const Stream& st = drv.streams()[s];
            int64_t* o = stats + 16 * (size_t)(s0 + s);
            for (int c = 0; c < 16; ++c) o[c] = 0;
            o[0] = st.lost; o[1] = st.n_keyframes; o[2] = st.n_ba; o[3] = st.n_candidates; o[4] = st.n_projected; o[5] = st.n_inliers;
            o[6] = st.ba_obs; o[7] = st.ba_pts; o[8] = st.ba_kfs; o[9] = st.ba_trials; o[10] = st.ba_iters; o[11] = (int64_t)st.ba_flops;
        }
        if (fr) ygzb_frames_destroy(fr);
        if (t > 0 && my) ygzb_destroy(my);
        rcs[t] = rc;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    *seconds = std::chrono::duration<double>(t_end - t_begin).count();
    if (totals) {
        for (int c = 0; c < 8; ++c) totals[c] = 0;
        for (int t = 0; t < n_threads; ++t)
            for (int c = 0; c < 4; ++c) totals[c] += tot[t][c];
    }
    if (getenv("YGZ_VO_TIMING")) {
        static const char* names[kTStages] = {"upload", "sparse_align", "project_align", "pose_only", "detect", "local_ba"};
        double sum = 0;
        for (int i = 0; i < kTStages; ++i) sum += 1e-9 * g_stage_ns[i].load();
        const int timed = n_frames - warm;
        fprintf(stderr, "[ygz_vo] %d streams on %d host threads x %d timed frames: %.3f ms per lock-step frame; C-ABI time summed over threads %.3f ms\n",
                n_streams, n_threads, timed, 1e3 * *seconds / timed, 1e3 * sum / timed);
        for (int i = 0; i < kTStages; ++i) fprintf(stderr, "[ygz_vo]   %-14s %.3f ms/frame\n", names[i], 1e-6 * g_stage_ns[i].load() / timed);
    }
    for (int rc : rcs)
        if (rc != YGZB_OK) return rc;
    return YGZB_OK;
}


// Device-resident engine (see Engine above).  Same arguments as ygz_vo_run_stages plus
//   window : frames of one stream that may be in flight in one round (1 = one frame at a time, the latency mode);
// frames [0, warm) are processed before the timed region starts (no window crosses frame `warm`).
int ygz_vo_run(ygzb_ctx* ctx, int device, const ygzb_params* params, int n_threads, int n_streams, int n_frames,
               const uint8_t* const* images, const double* const* depth, int kf_min_frames, double kf_min_rot, double kf_min_trans,
               int warm, int window, double* traj, int64_t* stats, double* seconds, double* device_ms, int64_t* totals) {
    if (!ctx || !params || n_streams < 1 || n_frames < 1 || !images || !depth || !traj || !stats || !seconds) return YGZB_ERR_INVALID;
    n_threads = std::max(1, std::min(n_threads, n_streams));
    warm = std::max(0, std::min(warm, n_frames - 1));
    for (auto& v : g_stage_ns) v.store(0);
    std::vector<int> rcs(n_threads, YGZB_OK);
    std::vector<std::vector<long long>> tot(n_threads, std::vector<long long>(4, 0));
    std::barrier sync_point(n_threads);
    std::chrono::steady_clock::time_point t_begin, t_end;
    auto worker = [&](int t) {
        const int s0 = (int)((long)n_streams * t / n_threads), s1 = (int)((long)n_streams * (t + 1) / n_threads), ns = s1 - s0;
        ygzb_ctx* my = ctx;
        int rc = YGZB_OK;
        if (t > 0) rc = ygzb_create(device, params, &my);
        {
            Engine eng(my, ns, window, Params{kf_min_frames, kf_min_rot, kf_min_trans});
            if (rc == YGZB_OK) rc = eng.init(depth + s0);
            double* my_traj = traj + (size_t)s0 * n_frames * 12;
            if (rc == YGZB_OK) rc = eng.run_until(images + s0, n_frames, warm, my_traj);
            if (rc == YGZB_OK) ygzb_synchronize(my);
            tot[t][0] = -ygzb_launch_count(my);
            eng.h2d_image_bytes = eng.h2d_other_bytes = eng.d2h_bytes = 0;
            sync_point.arrive_and_wait();
            if (t == 0) {
                t_begin = std::chrono::steady_clock::now();
                for (auto& v : g_stage_ns) v.store(0);
                ygzb_timer_start(ctx);
            }
            if (rc == YGZB_OK) rc = eng.run_until(images + s0, n_frames, n_frames, my_traj);
            if (rc == YGZB_OK) ygzb_synchronize(my);
            tot[t][0] += ygzb_launch_count(my);
            tot[t][1] = eng.h2d_image_bytes; tot[t][2] = eng.h2d_other_bytes; tot[t][3] = eng.d2h_bytes;
            sync_point.arrive_and_wait();
            if (t == 0) {
                double ms = 0;
                if (ygzb_timer_stop(ctx, &ms) == YGZB_OK && device_ms) *device_ms = ms;
                t_end = std::chrono::steady_clock::now();
            }
            for (int s = 0; s < ns; ++s) {
                const EStream& st = eng.streams()[s];
                int64_t* o = stats + 16 * (size_t)(s0 + s);
                for (int c = 0; c < 16; ++c) o[c] = 0;
                o[0] = st.lost; o[1] = st.n_keyframes; o[2] = st.n_ba; o[3] = st.n_candidates; o[4] = st.n_projected; o[5] = st.n_inliers;
                o[6] = st.ba_obs; o[7] = st.ba_pts; o[8] = st.ba_kfs; o[9] = st.ba_trials; o[10] = st.ba_iters; o[11] = (int64_t)st.ba_flops;
            }
        }   // (the engine releases its tracker and frame slots before the context goes)
        if (t > 0 && my) ygzb_destroy(my);
        rcs[t] = rc;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker, t);
    worker(0);
    for (auto& th : pool) th.join();
    *seconds = std::chrono::duration<double>(t_end - t_begin).count();
    if (totals) {
        for (int c = 0; c < 8; ++c) totals[c] = 0;
        for (int t = 0; t < n_threads; ++t)
            for (int c = 0; c < 4; ++c) totals[c] += tot[t][c];
    }
    if (getenv("YGZ_VO_TIMING")) {
        const int timed = n_frames - warm;
        fprintf(stderr, "[ygz_vo engine] %d streams on %d host threads, window %d, %d timed frames: %.3f ms per frame index; track rounds %.3f ms, "
                        "key-frame rounds %.3f ms (host wall, summed over threads)\n",
                n_streams, n_threads, window, timed, 1e3 * *seconds / timed, 1e-6 * g_stage_ns[kTSparse].load(), 1e-6 * g_stage_ns[kTLocalBA].load());
    }
    for (int rc : rcs)
        if (rc != YGZB_OK) return rc;
    return YGZB_OK;
}

}  // extern "C"
