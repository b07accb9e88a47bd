// oracle/vo_cpu.cpp -- the tracking loop of BASELINE config C5 on the CPU oracle, in C++ (TEST / MEASUREMENT
// INFRASTRUCTURE ONLY, see oracle.h): bench.py's `--impl reference` arm and its cpu_baseline leg time it, and
// tests/test_vo_cpu.py checks that it reproduces the Python loop (ygz_slam_b200/vo.py on oracle/vo_backend.py).
//
// It is CALLER code in the shape of the reference's src/Module/VisualOdometry.cpp:38-107 (AddFrame), :281-302
// (TrackRefFrame), src/Module/LocalMapping.cpp:24-140 (TrackLocalMap: FindCandidates :47-80, ProjectMapPoints :82-111,
// OptimizeCurrent :126), VisualOdometry.cpp:182-218 + :304-321 (SetKeyframe / NeedNewKeyFrame) and
// LocalMapping.cpp:149-172,301-336 (LocalBA -> ba::LocalBAG2O), one stream = one single-threaded sequence exactly as
// the reference runs one sequence on one thread.  Every numeric step is an oracle call (ora_build_pyramid,
// ora_matcher_sparse_alignment, ora_find_direct_projection, ora_pose_only, ora_detect, ora_local_ba_g2o); streams are
// independent, so the only parallelism offered to the CPU arm is one stream per host thread.
//
// Input-side simplifications are those of the GPU loop (ygz_slam_b200/host/vo_driver.cpp): ground-truth depth
// initialises the map points of a key-frame (test/test_feature_alignment.cpp:72-85 does the same with TUM depth),
// no BoW / loop closing.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <thread>
#include <vector>

#include "oracle.h"
#include "se3.h"

namespace {

using ora::SE3;

constexpr double FX = 520.9, FY = 521.0, CX = 325.1, CY = 249.7;   // config/default.yaml:32-35
constexpr int W = 640, H = 480, kLevels = 3;                        // Frame::Option::_pyramid_level (Frame.h:22-24)
constexpr int kLocalKeyframes = 3;                                  // LocalMapping.local_keyframes (default.yaml:68)
constexpr int kMinInliers = 30;                                     // vo.keyframe.min_features (default.yaml:66)

struct Mat34 {
    double m[12];
};
Mat34 identity() { return Mat34{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}}; }
Mat34 mul(const Mat34& A, const Mat34& B) {
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
void se3_log(const Mat34& T, double out[6]) { SE3::from_mat(T.m).log(out); }   // [upsilon; omega]

struct Keyframe {
    std::vector<uint8_t> pyr;    // the key-frame keeps its pyramid (Frame::_pyramid, Frame.h:138)
    int frame_id = 0;
    Mat34 T;
    std::vector<double> px;      // 2n full-res pixels
    std::vector<int32_t> level;
    std::vector<double> depth;
    std::vector<double> pw;      // 3n world points
    long mp0 = 0;
    std::vector<long> obs_id;
    std::vector<double> obs_px;
    int n() const { return (int)depth.size(); }
};

enum { kTPyramid, kTSparse, kTProject, kTPoseOnly, kTDetect, kTLocalBA, kTHost, kTStages };

struct Tick {
    double* acc;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    explicit Tick(double* a) : acc(a) {}
    ~Tick() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

class CpuStream {
  public:
    CpuStream(int kf_min_frames, double kf_min_rot, double kf_min_trans)
        : kf_min_frames_(kf_min_frames), kf_min_rot_(kf_min_rot), kf_min_trans_(kf_min_trans) {
        pyr_bytes_ = ora_pyramid_layout(W, H, kLevels, nullptr, nullptr, nullptr);
        cur_pyr_.resize(pyr_bytes_);
        const int cells = ((W + 9) / 10) * ((H + 9) / 10);
        kx_.resize(cells); ky_.resize(cells); klevel_.resize(cells); kcell_.resize(cells);
        kscore_.resize(cells); kangle_.resize(cells); kdesc_.resize((size_t)cells * 32);
    }

    // VisualOdometry::AddFrame for one grey frame; depth = the (static) ground-truth depth map of the stream
    void add_frame(const uint8_t* gray, const double* depth, int frame_id) {
        {
            Tick t(&sec[kTPyramid]);
            ora_build_pyramid(gray, W, H, kLevels, cur_pyr_.data());   // Frame::InitFrame
        }
        if (lost) return;
        if (keyframes_.empty()) {
            T = identity();
            has_pose = true;
            make_keyframe(depth, frame_id, true);
            return;
        }
        track(depth, frame_id);
    }

    Mat34 T = identity();
    bool has_pose = false, lost = false;
    long n_keyframes = 0, n_ba = 0, n_candidates = 0, n_projected = 0, n_inliers = 0;
    double sec[kTStages] = {0, 0, 0, 0, 0, 0, 0};

  private:
    int first_local() const { return std::max(0, (int)keyframes_.size() - kLocalKeyframes); }

    void track(const double* depth, int frame_id) {
        const ora_camera cam{(float)FX, (float)FY, (float)CX, (float)CY};
        const Keyframe& ref = keyframes_.back();
        // TrackRefFrame: Matcher::SparseImageAlignment(ref, cur) with cur._TCW = ref._TCW (VisualOdometry.cpp:66,281-302)
        Mat34 Tc = ref.T;
        int ok;
        {
            Tick t(&sec[kTSparse]);
            has_.assign(ref.n(), 1);
            ok = ora_matcher_sparse_alignment(ref.pyr.data(), cur_pyr_.data(), W, H, kLevels, &cam, ref.n(), ref.px.data(), ref.depth.data(),
                                              has_.data(), ref.T.m, Tc.m);
        }
        if (!ok) {
            lost = true;   // the reference keeps the last pose and reports VO_LOST
            return;
        }
        // TrackLocalMap / FindCandidates (LocalMapping.cpp:47-80) + ProjectMapPoints (:82-111): per local key-frame one
        // batch of Matcher::FindDirectProjection with poses relative to that key-frame (I, T_cur * T_kf^-1)
        std::vector<double> pw, obs;
        std::vector<long> obs_id;
        const Mat34 eye = identity();
        for (int k = first_local(); k < (int)keyframes_.size(); ++k) {
            const Keyframe& kf = keyframes_[k];
            c_px_.clear(); c_depth_.clear(); c_level_.clear(); c_init_.clear(); c_g_.clear();
            {
                Tick t(&sec[kTHost]);
                for (int g = 0; g < kf.n(); ++g) {
                    const double* X = &kf.pw[3 * (size_t)g];
                    const double x = Tc.m[0] * X[0] + Tc.m[1] * X[1] + Tc.m[2] * X[2] + Tc.m[3];
                    const double y = Tc.m[4] * X[0] + Tc.m[5] * X[1] + Tc.m[6] * X[2] + Tc.m[7];
                    const double z = Tc.m[8] * X[0] + Tc.m[9] * X[1] + Tc.m[10] * X[2] + Tc.m[11];
                    const double u = FX * x / z + CX, v = FY * y / z + CY;
                    if (!(z > 0 && u >= 20 && u < W - 20 && v >= 20 && v < H - 20)) continue;
                    c_g_.push_back(g);
                    c_px_.push_back(kf.px[2 * (size_t)g]);
                    c_px_.push_back(kf.px[2 * (size_t)g + 1]);
                    c_depth_.push_back(kf.depth[g]);
                    c_level_.push_back(kf.level[g]);
                    c_init_.push_back(u);
                    c_init_.push_back(v);
                }
            }
            const int nc = (int)c_g_.size();
            n_candidates += nc;
            if (!nc) continue;
            c_sl_.resize(nc);
            c_ok_.resize(nc);
            const Mat34 rel = mul(Tc, inv(kf.T));
            {
                Tick t(&sec[kTProject]);
                ora_find_direct_projection(kf.pyr.data(), cur_pyr_.data(), W, H, kLevels, &cam, eye.m, rel.m, nc, c_px_.data(), c_depth_.data(),
                                           c_level_.data(), c_init_.data(), c_sl_.data(), c_ok_.data());
            }
            for (int c = 0; c < nc; ++c) {
                if (!c_ok_[c]) continue;
                const int g = c_g_[c];
                pw.insert(pw.end(), &kf.pw[3 * (size_t)g], &kf.pw[3 * (size_t)g] + 3);
                obs.push_back(c_init_[2 * (size_t)c]);
                obs.push_back(c_init_[2 * (size_t)c + 1]);
                obs_id.push_back(kf.mp0 + g);
            }
        }
        const int np = (int)obs_id.size();
        n_projected += np;
        // ba::OptimizeCurrentPoseOnly (LocalMapping.cpp:126; BA.cpp:188-264)
        int n_inl = 0;
        inl_.assign(np ? np : 1, 0);
        dep_out_.assign(np ? np : 1, 0.0);
        if (np) {
            Tick t(&sec[kTPoseOnly]);
            n_inl = ora_pose_only(&cam, np, pw.data(), obs.data(), Tc.m, inl_.data(), dep_out_.data());
        }
        if (n_inl < kMinInliers) {
            lost = true;
            return;
        }
        last_id_.clear();
        last_px_.clear();
        for (int q = 0; q < np; ++q)
            if (inl_[q]) {
                last_id_.push_back(obs_id[q]);
                last_px_.push_back(obs[2 * (size_t)q]);
                last_px_.push_back(obs[2 * (size_t)q + 1]);
            }
        has_last_ = true;
        T = Tc;
        frames_since_kf_ += 1;
        n_inliers += n_inl;
        // NeedNewKeyFrame (VisualOdometry.cpp:304-321)
        if (frames_since_kf_ < kf_min_frames_) return;
        double d[6];
        se3_log(mul(T, inv(keyframes_.back().T)), d);
        const double rot = std::sqrt(d[3] * d[3] + d[4] * d[4] + d[5] * d[5]), tr = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (rot > kf_min_rot_ || tr > kf_min_trans_) make_keyframe(depth, frame_id, false);
    }

    // SetKeyframe: Detect, depth-initialised map points, local BA (VisualOdometry.cpp:182-218)
    void make_keyframe(const double* depth, int frame_id, bool fresh) {
        const ora_detect_params prm{W, H, 10, 15, kLevels};
        ora_features f{0, kx_.data(), ky_.data(), klevel_.data(), kscore_.data(), kangle_.data(), kdesc_.data(), kcell_.data()};
        int cnt;
        {
            Tick t(&sec[kTDetect]);
            cnt = ora_detect(cur_pyr_.data(), &prm, nullptr, &f);
        }
        Keyframe kf;
        kf.pyr = cur_pyr_;
        kf.frame_id = frame_id;
        kf.T = T;
        const Mat34 Tin = inv(T);
        kf.px.resize(2 * (size_t)cnt); kf.level.resize(cnt); kf.depth.resize(cnt); kf.pw.resize(3 * (size_t)cnt);
        for (int g = 0; g < cnt; ++g) {
            // the device path hands the pixel over as f32 (exact: level coordinate * 2^level)
            const double x = kx_[g], y = ky_[g];
            const double d = depth[(size_t)(int)y * W + (int)x];
            kf.px[2 * (size_t)g] = x;
            kf.px[2 * (size_t)g + 1] = y;
            kf.level[g] = klevel_[g];
            kf.depth[g] = d;
            const double pc[3] = {(x - CX) * d / FX, (y - CY) * d / FY, d};
            for (int r = 0; r < 3; ++r)
                kf.pw[3 * (size_t)g + r] = Tin.m[4 * r] * pc[0] + Tin.m[4 * r + 1] * pc[1] + Tin.m[4 * r + 2] * pc[2] + Tin.m[4 * r + 3];
        }
        kf.mp0 = next_mp_;
        next_mp_ += cnt;
        if (!fresh && has_last_) {
            kf.obs_id = last_id_;
            kf.obs_px = last_px_;
        }
        keyframes_.push_back(std::move(kf));
        while ((int)keyframes_.size() > kLocalKeyframes + 1) keyframes_.pop_front();
        frames_since_kf_ = 0;
        n_keyframes += 1;
        if (!fresh && keyframes_.size() >= 2) local_ba();
    }

    // LocalMapping::LocalBA -> ba::LocalBAG2O over the local key-frames and the points at least two of them observe
    void local_ba() {
        const ora_camera cam{(float)FX, (float)FY, (float)CX, (float)CY};
        std::vector<int32_t> kf_idx, pt_idx;
        std::vector<double> poses, pts, obs;
        std::vector<uint8_t> fixed;
        struct Ref { int kf, n; };
        std::vector<Ref> owners;
        const int k0 = first_local(), nk = (int)keyframes_.size() - k0;
        {
            Tick t(&sec[kTHost]);
            for (int k = 0; k < nk; ++k) {
                double lg[6];
                se3_log(keyframes_[k0 + k].T, lg);
                const double g2o[6] = {lg[3], lg[4], lg[5], lg[0], lg[1], lg[2]};   // VertexSE3Sophus: [omega; upsilon]
                poses.insert(poses.end(), g2o, g2o + 6);
                fixed.push_back(k == 0);   // the oldest local key-frame fixes the gauge (key-frame 0 in the reference)
            }
            struct Ob { long id; int kf; double u, v; };
            std::vector<Ob> all;
            auto in_local = [&](long id) {
                for (int k = 0; k < nk; ++k) {
                    const Keyframe& kf = keyframes_[k0 + k];
                    if (id >= kf.mp0 && id < kf.mp0 + kf.n()) return true;
                }
                return false;
            };
            for (int k = 0; k < nk; ++k) {
                const Keyframe& kf = keyframes_[k0 + k];
                for (int g = 0; g < kf.n(); ++g) all.push_back({kf.mp0 + g, k, kf.px[2 * (size_t)g], kf.px[2 * (size_t)g + 1]});
                for (size_t q = 0; q < kf.obs_id.size(); ++q)
                    if (in_local(kf.obs_id[q])) all.push_back({kf.obs_id[q], k, kf.obs_px[2 * q], kf.obs_px[2 * q + 1]});
            }
            std::vector<long> ids;
            ids.reserve(all.size());
            for (const Ob& o : all) ids.push_back(o.id);
            std::sort(ids.begin(), ids.end());
            std::vector<long> multi;   // ids with >= 2 observations, ascending
            for (size_t a = 0; a < ids.size();) {
                size_t b = a;
                while (b < ids.size() && ids[b] == ids[a]) ++b;
                if (b - a >= 2) multi.push_back(ids[a]);
                a = b;
            }
            for (long id : multi)
                for (int k = 0; k < nk; ++k) {
                    const Keyframe& kf = keyframes_[k0 + k];
                    if (id >= kf.mp0 && id < kf.mp0 + kf.n()) {
                        const int g = (int)(id - kf.mp0);
                        owners.push_back({k0 + k, g});
                        pts.insert(pts.end(), &kf.pw[3 * (size_t)g], &kf.pw[3 * (size_t)g] + 3);
                        break;
                    }
                }
            for (const Ob& o : all) {
                const auto it = std::lower_bound(multi.begin(), multi.end(), o.id);
                if (it == multi.end() || *it != o.id) continue;
                kf_idx.push_back(o.kf);
                pt_idx.push_back((int32_t)(it - multi.begin()));
                obs.push_back(o.u);
                obs.push_back(o.v);
            }
        }
        const int n_pt = (int)owners.size(), n_obs = (int)kf_idx.size();
        const ora_ba_params bp{20, 5.991, 5.991, 1e-5, 10};   // BA.cpp:450-452,501-502; g2o Levenberg defaults
        std::vector<uint8_t> outl(n_obs + 1);
        ora_ba_stats st;
        static const double zero3[3] = {0, 0, 0};
        static const int32_t zero_i = 0;
        {
            Tick t(&sec[kTLocalBA]);
            ora_local_ba_g2o(&cam, nk, poses.data(), fixed.data(), n_pt, n_pt ? pts.data() : const_cast<double*>(zero3), n_obs,
                             n_obs ? kf_idx.data() : &zero_i, n_obs ? pt_idx.data() : &zero_i, n_obs ? obs.data() : zero3, &bp, outl.data(), &st);
        }
        for (int k = 0; k < nk; ++k) {
            const double* g = &poses[6 * (size_t)k];
            const double v[6] = {g[3], g[4], g[5], g[0], g[1], g[2]};
            SE3::exp(v).to_mat(keyframes_[k0 + k].T.m);
        }
        for (int q = 0; q < n_pt; ++q) {
            Keyframe& kf = keyframes_[owners[q].kf];
            std::memcpy(&kf.pw[3 * (size_t)owners[q].n], &pts[3 * (size_t)q], 3 * sizeof(double));
        }
        T = keyframes_.back().T;
        n_ba += 1;
    }

    int kf_min_frames_;
    double kf_min_rot_, kf_min_trans_;
    size_t pyr_bytes_ = 0;
    std::vector<uint8_t> cur_pyr_;
    std::deque<Keyframe> keyframes_;
    int frames_since_kf_ = 0;
    long next_mp_ = 0;
    bool has_last_ = false;
    std::vector<long> last_id_;
    std::vector<double> last_px_;
    // scratch
    std::vector<uint8_t> has_, c_ok_, inl_, kdesc_;
    std::vector<double> c_px_, c_depth_, c_init_, dep_out_, kx_, ky_;
    std::vector<int32_t> c_level_, c_sl_, klevel_, kcell_;
    std::vector<int> c_g_;
    std::vector<float> kscore_, kangle_;
};

}  // namespace

// Tracks n_streams independent 640x480 grey streams of n_frames frames each, one stream per task on `threads` host
// threads (a stream is a sequential job, as in the reference).  Frames [0, warm) of every stream run untimed, then all
// threads meet and frames [warm, n_frames) are timed (wall clock until the last stream finishes).
//   images[s] : n_frames * 480 * 640 bytes, depth[s] : 480 * 640 doubles
//   traj      : n_streams * n_frames * 12 doubles out (T_cw after every frame; NaN while a stream has no pose)
//   stats     : n_streams * 8 out: lost, keyframes, local BAs, candidates, projected, inliers, 0, 0
//   stage_sec : kTStages (7) doubles out or NULL: seconds inside pyramid / sparse alignment / direct projection / pose-only /
//               detect / local BA / host bookkeeping, summed over the streams, timed region only
extern "C" int ora_vo_run(int n_streams, int n_frames, const uint8_t* const* images, const double* const* depth, int kf_min_frames,
                          double kf_min_rot, double kf_min_trans, int warm, int threads, double* traj, int64_t* stats, double* seconds,
                          double* stage_sec) {
    if (n_streams < 1 || n_frames < 1 || !images || !depth || !seconds) return -1;
    threads = std::max(1, std::min(threads, n_streams));
    warm = std::max(0, std::min(warm, n_frames - 1));
    std::vector<CpuStream> st;
    st.reserve(n_streams);
    for (int s = 0; s < n_streams; ++s) st.emplace_back(kf_min_frames, kf_min_rot, kf_min_trans);
    auto phase = [&](int k0, int k1) {
        std::atomic<int> next{0};
        auto work = [&]() {
            for (int s = next.fetch_add(1); s < n_streams; s = next.fetch_add(1))
                for (int k = k0; k < k1; ++k) {
                    st[s].add_frame(images[s] + (size_t)k * W * H, depth[s], k);
                    if (traj) {
                        double* out = traj + ((size_t)s * n_frames + k) * 12;
                        for (int c = 0; c < 12; ++c) out[c] = st[s].has_pose ? st[s].T.m[c] : NAN;
                    }
                }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < threads; ++t) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
    };
    phase(0, warm);
    for (auto& s : st)
        for (double& v : s.sec) v = 0;
    const auto t0 = std::chrono::steady_clock::now();
    phase(warm, n_frames);
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int s = 0; s < n_streams; ++s) {
        if (stats) {
            int64_t* o = stats + 8 * (size_t)s;
            o[0] = st[s].lost; o[1] = st[s].n_keyframes; o[2] = st[s].n_ba; o[3] = st[s].n_candidates; o[4] = st[s].n_projected;
            o[5] = st[s].n_inliers; o[6] = o[7] = 0;
        }
        if (stage_sec)
            for (int i = 0; i < kTStages; ++i) stage_sec[i] = (s == 0 ? 0.0 : stage_sec[i]) + st[s].sec[i];
    }
    return 0;
}
