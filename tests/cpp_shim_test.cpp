// Builds against the C++ shim (ygz_slam_b200/host/ygz_b200.hpp) the way a reference caller would:
// test_orb_match shape (Detect on two frames + brute-force match) followed by sparse alignment and
// direct projection.  Input: two raw 640x480 grey frames + depth of frame 1 + the relative pose, written
// by the Python test; prints a few summary numbers that the Python side compares with the oracle.
#include <cmath>
#include <cstdio>
#include <map>
#include <set>
#include <vector>

#include "../ygz_slam_b200/host/ygz_b200.hpp"

using namespace ygz;

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) return 3;
    Frame f1, f2;
    f1._color.create(480, 640, 1);
    f2._color.create(480, 640, 1);
    std::vector<float> depth(640 * 480);
    double T2[12];
    if (fread(f1._color.data, 1, 640 * 480, fp) != 640 * 480 || fread(f2._color.data, 1, 640 * 480, fp) != 640 * 480 ||
        fread(depth.data(), 4, 640 * 480, fp) != 640 * 480 || fread(T2, 8, 12, fp) != 12)
        return 4;
    fclose(fp);
    PinholeCamera cam;
    Frame::SetCamera(&cam);
    f1.InitFrame();
    f2.InitFrame();
    FeatureDetector det;
    det.Detect(&f1);
    det.Detect(&f2);
    std::vector<int> idx, dist;
    Matcher::BruteForceMatch(&f1, &f2, idx, dist);
    int n_match = 0;
    long sum_dist = 0;
    for (size_t i = 0; i < idx.size(); ++i)
        if (idx[i] >= 0) { ++n_match; sum_dist += dist[i]; }
    printf("features %zu %zu matches %d dist_sum %ld\n", f1._features.size(), f2._features.size(), n_match, sum_dist);
    // give the features of frame 1 depth + map points, then align frame 2 against it
    std::vector<MapPoint> mps(f1._features.size());
    for (size_t i = 0; i < f1._features.size(); ++i) {
        Feature* f = f1._features[i];
        f->_depth = depth[(int)f->_pixel[1] * 640 + (int)f->_pixel[0]];
        f->_mappoint = &mps[i];
    }
    Matcher m;
    const bool ok = m.SparseImageAlignment(&f1, &f2);
    double est[12];
    f2._TCW.matrix3x4(est);
    double err = 0;
    for (int i = 0; i < 12; ++i) err = std::fmax(err, std::fabs(est[i] - T2[i]));
    printf("sparse_align ok %d max_abs_pose_diff_to_gt %.6f\n", (int)ok, err);
    int n_ok = 0;
    for (size_t i = 0; i < f1._features.size(); i += 4) {
        Feature* f = f1._features[i];
        const Vector3d pc((f->_pixel[0] - cam.cx()) * f->_depth / cam.fx(), (f->_pixel[1] - cam.cy()) * f->_depth / cam.fy(), f->_depth);
        Vector2d px = cam.World2Pixel(pc, f2._TCW);  // T_ref = I: world == ref camera
        int level = 0;
        n_ok += m.FindDirectProjection(&f1, &f2, f, px, level);
    }
    printf("direct_projection ok %d\n", n_ok);
    // Tracker: KLT from frame 1 to frame 2 (test_LK_tracking.cpp shape)
    Tracker trk;
    trk.SetReference(&f1);
    trk.Track(&f2);
    std::vector<Feature*> tf;
    std::vector<Vector2d> tp;
    trk.GetTrackedPixel(tf, tp);
    printf("klt status %d tracked %zu mean_disparity %.4f\n", (int)trk.Status(), tp.size(), trk.MeanDisparity());

    // test_local_ba.cpp shape: 8 key-frames x 16 points, perturbed poses / points, both BA flavours
    for (int flavour = 0; flavour < 2; ++flavour) {
        const double rot[8][3] = {{0, 0, 0}, {0.1, 0, 0}, {0, 0.1, 0}, {0, 0, 0.1}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        const double tr[8][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0.1, 0, 0}, {0, 0.1, 0}, {0, 0, 0.1}, {0.1, 0.1, 0.1}};
        std::vector<Frame> kfs(8);
        std::vector<MapPoint> pts(16);
        std::vector<Vector3d> truth(16);
        std::map<unsigned long, Frame*> keyframe_of;
        unsigned lcg = 12345u;
        auto noise = [&](double s) { lcg = lcg * 1664525u + 1013904223u; return s * (((lcg >> 8) & 0xFFFF) / 32768.0 - 1.0); };
        std::vector<SE3> T_true(8);
        for (int k = 0; k < 8; ++k) {
            const double v[6] = {tr[k][0], tr[k][1], tr[k][2], rot[k][0], rot[k][1], rot[k][2]};
            T_true[k] = SE3::exp(v);
            kfs[k]._keyframe_id = k;
            kfs[k]._is_keyframe = true;
            keyframe_of[k] = &kfs[k];
        }
        std::vector<std::vector<Feature*>> feats(16, std::vector<Feature*>(8, nullptr));   // [point][key-frame]
        for (int j = 0; j < 16; ++j) {
            truth[j] = Vector3d(j % 2, (j / 2) % 2, 2 + j / 4);
            pts[j]._id = j;
            for (int k = 0; k < 8; ++k) {
                Feature* f = new Feature(cam.World2Pixel(truth[j], T_true[k]));
                kfs[k]._features.push_back(f);
                pts[j]._obs[k] = f;
                feats[j][k] = f;
            }
            pts[j]._pos_world = Vector3d(truth[j][0] + noise(0.05), truth[j][1] + noise(0.05), truth[j][2] + noise(0.05));
        }
        for (int k = 0; k < 8; ++k) {
            double v[6] = {tr[k][0], tr[k][1], tr[k][2], rot[k][0], rot[k][1], rot[k][2]};
            if (k > 0)
                for (int c = 0; c < 6; ++c) v[c] += noise(0.03);
            kfs[k]._TCW = SE3::exp(v);
        }
        auto rms = [&]() {
            double s2 = 0;
            for (int j = 0; j < 16; ++j)
                for (int k = 0; k < 8; ++k) {
                    const Vector2d px = cam.World2Pixel(pts[j]._pos_world, kfs[k]._TCW);
                    const Feature* f = feats[j][k];
                    s2 += (px[0] - f->_pixel[0]) * (px[0] - f->_pixel[0]) + (px[1] - f->_pixel[1]) * (px[1] - f->_pixel[1]);
                }
            return std::sqrt(s2 / (16 * 8));
        };
        std::set<Frame*> lk;
        std::set<MapPoint*> lm;
        for (auto& f : kfs) lk.insert(&f);
        for (auto& m_ : pts) lm.insert(&m_);
        const double before = rms();
        if (flavour == 0) ba::LocalBAG2O(lk, lm, keyframe_of);
        else ba::LocalBA(lk, lm, keyframe_of);
        if (flavour == 1) {
            // ba::OptimizeCurrent / OptimizeCurrentPointOnly on top of the converged map: perturb the last frame's pose and
            // a few points, the refinement must bring the reprojection error back down
            Frame& cur = kfs[7];
            for (int j = 0; j < 16; ++j) {   // the current frame is not (yet) an observer key-frame of its points
                feats[j][7]->_mappoint = &pts[j];
                pts[j]._obs.erase(7);
            }
            double v[6];
            cur._TCW.log(v);
            for (int c = 0; c < 6; ++c) v[c] += noise(0.01);
            cur._TCW = SE3::exp(v);
            const double b2 = rms();
            ba::OptimizeCurrent(&cur, keyframe_of);
            int n_bad = 0;
            for (Feature* f : cur._features) n_bad += f->_bad;
            printf("optimize_current rms_px_before %.4f after %.6f bad %d\n", b2, rms(), n_bad);
            for (int j = 0; j < 16; j += 3) pts[j]._pos_world = Vector3d(pts[j]._pos_world[0] + 0.02, pts[j]._pos_world[1] - 0.02, pts[j]._pos_world[2] + 0.03);
            const double b3 = rms();
            ba::OptimizeCurrentPointOnly(&cur, keyframe_of);
            printf("optimize_point_only rms_px_before %.4f after %.6f\n", b3, rms());
        }
        double T0[12];
        kfs[0]._TCW.matrix3x4(T0);
        printf("local_ba %s rms_px_before %.4f after %.6f kf0_moved %d\n", flavour == 0 ? "g2o" : "ceres", before, rms(),
               (int)(std::fabs(T0[3]) + std::fabs(T0[7]) + std::fabs(T0[11]) > 1e-12));
    }
    {
        // slot ownership: a live frame (e.g. a key-frame kept by Memory) keeps its device pyramid however many frames come
        // and go; the pool reports exhaustion instead of recycling a live slot; the section-8b helpers are reachable
        auto& rt = b200::Runtime::Get();
        const int free0 = rt.FreeSlots();
        Frame* keyframe = new Frame();
        keyframe->_color.create(480, 640, 1);
        std::memcpy(keyframe->_color.data, f1._color.data, 640 * 480);
        keyframe->InitFrame();
        const int kf_slot = keyframe->_slot;
        for (int k = 0; k < 200; ++k) {   // 200 transient frames through a 64-slot pool
            Frame tmp;
            tmp._color.create(480, 640, 1);
            std::memcpy(tmp._color.data, f2._color.data, 640 * 480);
            tmp.InitFrame();
            if (tmp._slot == kf_slot) return 10;
        }
        FeatureDetector det2;
        det2.Detect(keyframe);   // its pyramid is still there
        const size_t n_kf = keyframe->_features.size();
        uint8_t before[32];
        std::memcpy(before, keyframe->_features[0]->_desc, 32);
        det2.ComputeDescriptor(keyframe->_features[0]);
        const bool same_desc = std::memcmp(before, keyframe->_features[0]->_desc, 32) == 0;
        bool exhausted = false;
        std::vector<Frame*> hold;
        try {
            for (int k = 0; k < 100; ++k) {
                hold.push_back(new Frame());
                hold.back()->_color.create(480, 640, 1);
                hold.back()->InitFrame();
            }
        } catch (const b200::Error&) {
            exhausted = true;
        }
        for (Frame* f : hold) delete f;
        delete keyframe;
        const auto J = cvutils::JacobXYZ2Cam(Vector3d(0.1, -0.2, 2.0));
        printf("slots keyframe_features %zu same_as_frame1 %d compute_descriptor_same %d exhausted %d free_restored %d jac %.6f\n", n_kf,
               (int)(n_kf == f1._features.size()), (int)same_desc, (int)exhausted, (int)(rt.FreeSlots() == free0), J(0, 4));
    }
    if (argc > 2) {
        // test_orb_match.cpp:72-75 shape: vocabulary from a DBoW3 binary file, Frame::ComputeBoW, Matcher::SearchByBoW
        ORBVocabulary vocab;
        if (!vocab.loadFromBinaryFile(argv[2])) return 11;
        Frame::SetORBVocabulary(&vocab);
        f1.ComputeBoW();
        f2.ComputeBoW();
        double sum1 = 0;
        for (auto& kv : f1._bow_vec) sum1 += kv.second;
        size_t in_fv = 0;
        for (auto& kv : f1._feature_vec) in_fv += kv.second.size();
        std::map<int, int> bow_matches;
        const int cnt = m.SearchByBoW(&f1, &f2, bow_matches);
        long h = 0;
        for (auto& kv : bow_matches) h = (h * 31 + kv.first * 7 + kv.second) % 1000003;
        printf("bow words %u bow_vec %zu feature_vec_nodes %zu features_in_nodes %zu sum %.12f matches %d map %zu hash %ld\n", vocab.size(),
               f1._bow_vec.size(), f1._feature_vec.size(), in_fv, sum1, cnt, bow_matches.size(), h);
        if (argc > 3) {
            // LocalMapping::CreateNewMapPoints (LocalMapping.cpp:375-571): a third frame with a wider baseline is the new key-frame,
            // frame 1 its neighbour with map points on every second feature (ground-truth depth); matches between two features
            // without a map point are triangulated, matches with a mapped feature are associated
            FILE* f3p = fopen(argv[3], "rb");
            if (!f3p) return 12;
            Frame f3;
            f3._color.create(480, 640, 1);
            double T3[12];
            if (fread(f3._color.data, 1, 640 * 480, f3p) != 640 * 480 || fread(T3, 8, 12, f3p) != 12) return 13;
            fclose(f3p);
            f3.InitFrame();
            det.Detect(&f3);
            f3.ComputeBoW();
            f1._keyframe_id = 1;
            f3._keyframe_id = 3;
            f1._TCW = SE3();
            f3._TCW = SE3::from3x4(T3);
            for (size_t i = 0; i < f1._features.size(); ++i) {
                Feature* f = f1._features[i];
                if (i % 2 == 0) {
                    f->_mappoint = &mps[i];
                    mps[i]._pos_world = cam.Pixel2Camera(f->_pixel, f->_depth);
                    mps[i]._bad = false;
                } else {
                    f->_mappoint = nullptr;
                }
            }
            std::vector<Vector2d> px_saved;   // CreateNewMapPoints refines the neighbour's matched pixels (LocalMapping.cpp:446)
            for (Feature* f : f1._features) px_saved.push_back(f->_pixel);
            LocalMapping lm;
            lm.CreateNewMapPoints(&f3, {&f1});
            double sx = 0, sy = 0, sz = 0;
            for (auto& mp : lm._new_points) {
                sx += mp->_pos_world[0];
                sy += mp->_pos_world[1];
                sz += mp->_pos_world[2];
            }
            printf("create_new_map_points new %d associated %d sum %.9f %.9f %.9f\n", lm._cnt_new_mappoints, lm._cnt_associate_mps, sx, sy, sz);
            for (size_t i = 0; i < f1._features.size(); ++i) {   // (the points of `lm` die with it)
                f1._features[i]->_mappoint = nullptr;
                f1._features[i]->_pixel = px_saved[i];
            }
        }
        Frame::SetORBVocabulary(nullptr);
    }
    {
        // Initializer (RANSAC half) on the brute-force matches of the two frames
        std::vector<Vector2d> q1, q2;
        for (size_t i = 0; i < idx.size(); ++i)
            if (idx[i] >= 0) {
                q1.push_back(f1._features[i]->_pixel);
                q2.push_back(f2._features[idx[i]]->_pixel);
            }
        Initializer init;
        const bool use_h = init.FindModels(q1, q2);
        int nh = 0, nf = 0;
        for (bool v : init._inliers_H) nh += v;
        for (bool v : init._inliers_F) nf += v;
        printf("initializer pairs %zu use_h %d score_h %.3f score_f %.3f inl_h %d inl_f %d f22 %.9e\n", q1.size(), (int)use_h, init._score_H,
               init._score_F, nh, nf, init._F21(2, 2));
        Initializer full;
        const bool ok = full.TryInitialize(q1, q2, &f1, &f2);
        double T21[12];
        full._T21.matrix3x4(T21);
        int n_tri = 0;
        for (bool v : full._inliers) n_tri += v;
        printf("try_initialize ok %d triangulated %d t21 %.9e %.9e %.9e r00 %.9e\n", (int)ok, ok ? n_tri : 0, T21[3], T21[7], T21[11], T21[0]);
    }
    return 0;
}
