/* examples/extract_match.c -- the C ABI from plain C, in the shape of the reference's test/test_orb_match.cpp:
 * two 640x480 grey frames -> pyramid -> grid FAST + ORB -> cross-checked brute-force match -> the "good match"
 * filter of test_orb_match.cpp:97-105 on the host.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/extract_match.c -Lygz_slam_b200 -lygz_b200 -Wl,-rpath,$PWD/ygz_slam_b200 -o extract_match
 *   ./extract_match frame1.pgm frame2.pgm        (binary P5, 640x480; without arguments two synthetic frames are used)
 *
 * Needs a B200 (sm_100) at run time: there is no CPU fallback, ygzb_create says so.                                   */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ygz_b200.h"

#define W 640
#define H 480

static int read_pgm(const char* path, unsigned char* dst) {
    FILE* f = fopen(path, "rb");
    int w = 0, h = 0, maxv = 0;
    if (!f) return 0;
    if (fscanf(f, "P5 %d %d %d", &w, &h, &maxv) != 3 || w != W || h != H || maxv != 255) {
        fclose(f);
        return 0;
    }
    fgetc(f); /* the single whitespace after the header */
    if (fread(dst, 1, (size_t)W * H, f) != (size_t)W * H) {
        fclose(f);
        return 0;
    }
    fclose(f);
    return 1;
}

/* deterministic texture: blocks of random grey levels, shifted by (dx, dy) for the second frame */
static void synthetic(unsigned char* dst, int dx, int dy) {
    int x, y;
    for (y = 0; y < H; ++y)
        for (x = 0; x < W; ++x) {
            unsigned u = (unsigned)((x + dx) / 9), v = (unsigned)((y + dy) / 7);
            unsigned hsh = (u * 2654435761u) ^ (v * 40503u + 0x9E3779B9u);
            hsh ^= hsh >> 13;
            hsh *= 0x5bd1e995u;
            dst[y * W + x] = (unsigned char)(40 + (hsh >> 24) % 160);
        }
}

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != YGZB_OK) {                                                        \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ctx ? ygzb_last_error(ctx) : "no usable sm_100 device"); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    ygzb_params prm;
    ygzb_ctx* ctx = NULL;
    ygzb_frames* frames = NULL;
    unsigned char* img = (unsigned char*)malloc((size_t)2 * W * H);
    int rows = 0, cols = 0, cap, i, n_match = 0, n_good = 0, min_dist = 9999;
    ygzb_keypoints kp;
    int32_t slots[2] = {0, 1}, a_slot[1] = {0}, b_slot[1] = {1}, qoff[2];
    int32_t *idx, *dist;

    if (!img) return 1;
    if (argc >= 3) {
        if (!read_pgm(argv[1], img) || !read_pgm(argv[2], img + (size_t)W * H)) {
            fprintf(stderr, "expected two binary 640x480 PGM files\n");
            return 2;
        }
    } else {
        synthetic(img, 0, 0);
        synthetic(img + (size_t)W * H, 5, 3);
    }
    ygzb_default_params(&prm); /* 640x480, 3 levels, cell 10, threshold 15: config/default.yaml + Frame.h:23 */
    CHECK(ygzb_create(0, &prm, &ctx));
    CHECK(ygzb_frames_create(ctx, 2, &frames));
    CHECK(ygzb_frames_upload(frames, 0, 2, img, 1, (size_t)W * H)); /* Frame::InitFrame x 2 */
    CHECK(ygzb_grid_dims(ctx, &rows, &cols));
    cap = 2 * rows * cols;
    memset(&kp, 0, sizeof kp);
    kp.offsets = (int32_t*)malloc(3 * sizeof(int32_t));
    kp.x = (float*)malloc(cap * sizeof(float));
    kp.y = (float*)malloc(cap * sizeof(float));
    kp.level = (uint8_t*)malloc(cap);
    kp.score = (float*)malloc(cap * sizeof(float));
    kp.angle = (float*)malloc(cap * sizeof(float));
    kp.desc = (uint8_t*)malloc((size_t)cap * 32);
    kp.cell = NULL;
    kp.capacity = cap;
    idx = (int32_t*)malloc(cap * sizeof(int32_t));
    dist = (int32_t*)malloc(cap * sizeof(int32_t));
    CHECK(ygzb_detect(frames, slots, 2, NULL, &kp)); /* FeatureDetector::Detect x 2 */
    CHECK(ygzb_match_frames(frames, a_slot, b_slot, 1, 1, qoff, idx, dist, cap)); /* BFMatcher(NORM_HAMMING, true) */
    for (i = 0; i < qoff[1]; ++i)
        if (idx[i] >= 0) {
            ++n_match;
            if (dist[i] < min_dist) min_dist = dist[i];
        }
    /* test_orb_match.cpp:97-105: keep matches below 3 * clamp(min distance, 20, 50) */
    if (min_dist < 20) min_dist = 20;
    if (min_dist > 50) min_dist = 50;
    for (i = 0; i < qoff[1]; ++i)
        if (idx[i] >= 0 && dist[i] < 3 * min_dist) ++n_good;
    printf("keypoints %d + %d, cross-checked matches %d, good matches %d, kernel launches %lld\n", kp.offsets[1], kp.offsets[2] - kp.offsets[1],
           n_match, n_good, ygzb_launch_count(ctx));
    ygzb_frames_destroy(frames);
    ygzb_destroy(ctx);
    return 0;
}
