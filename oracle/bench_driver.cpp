// oracle/bench_driver.cpp -- multi-threaded driver of the CPU path for bench.py's `--impl reference` arm and the
// cpu_baseline leg (TEST / MEASUREMENT INFRASTRUCTURE ONLY, see oracle.h).  Frame-parallel: every thread runs the
// reference's single-threaded per-frame path (pyramid -> Detect -> cross-checked BF match against the previous
// frame) on its own contiguous chunk of frames, so all host cores are busy without any Python in the timed region.
#include <chrono>
#include <cmath>
#include <thread>
#include <vector>

#include "oracle.h"

namespace {

struct Feat {
    int n = 0;
    std::vector<uint8_t> desc;
};

void run_chunk(const uint8_t* frames, int n_frames, int w, int h, int n_levels, int lo, int hi, long* n_feat) {
    const int cells = (int)std::ceil(w / 10.0) * (int)std::ceil(h / 10.0);
    size_t total = ora_pyramid_layout(w, h, n_levels, nullptr, nullptr, nullptr);
    std::vector<uint8_t> pyr(total);
    std::vector<double> px(cells), py(cells);
    std::vector<int32_t> level(cells), cell(cells), idx(cells), dist(cells);
    std::vector<float> score(cells), angle(cells);
    Feat prev, cur;
    prev.desc.resize((size_t)cells * 32);
    cur.desc.resize((size_t)cells * 32);
    const ora_detect_params prm{w, h, 10, 15, n_levels};
    long nf = 0;
    for (int k = lo; k <= hi; ++k) {  // one extra frame so that every frame of [lo, hi) has its successor
        const uint8_t* g = frames + (size_t)(k % n_frames) * w * h;
        ora_build_pyramid(g, w, h, n_levels, pyr.data());
        ora_features f{0, px.data(), py.data(), level.data(), score.data(), angle.data(), cur.desc.data(), cell.data()};
        cur.n = ora_detect(pyr.data(), &prm, nullptr, &f);
        if (k > lo) ora_match_bf(prev.desc.data(), prev.n, cur.desc.data(), cur.n, 1, idx.data(), dist.data());
        if (k < hi) nf += cur.n;
        std::swap(prev, cur);
    }
    *n_feat = nf;
}

}  // namespace

// processes frames [0, count) (indices wrap modulo n_frames) on `threads` threads; returns elapsed seconds
extern "C" double ora_bench_extract_match(const uint8_t* frames, int n_frames, int w, int h, int n_levels, int count, int threads,
                                          long* n_feat_out) {
    if (threads < 1) threads = 1;
    std::vector<std::thread> pool;
    std::vector<long> nf(threads, 0);
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; ++t) {
        const int lo = (int)((long)count * t / threads), hi = (int)((long)count * (t + 1) / threads);
        if (hi > lo) pool.emplace_back(run_chunk, frames, n_frames, w, h, n_levels, lo, hi, &nf[t]);
    }
    for (auto& th : pool) th.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long total = 0;
    for (long v : nf) total += v;
    if (n_feat_out) *n_feat_out = total;
    return dt;
}
