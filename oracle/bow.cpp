// oracle/bow.cpp -- CPU restatement of the DBoW3 pieces the reference's matching path uses (TEST INFRASTRUCTURE ONLY, see
// oracle.h): the ORBvoc.bin loader, the vocabulary-tree transform behind Frame::ComputeBoW, and Matcher::SearchByBoW.
//   DBoW3::Vocabulary::loadFromBinaryFile   reference thirdparty/DBoW3/src/Vocabulary.cpp:1180-1225
//   DBoW3::Vocabulary::transform (features) reference thirdparty/DBoW3/src/Vocabulary.cpp:706-776
//   DBoW3::Vocabulary::transform (one)      reference thirdparty/DBoW3/src/Vocabulary.cpp:790-832
//   DBoW3::DescManip::distance (binary)     reference thirdparty/DBoW3/src/DescManip.cpp:91-115  (= 256-bit Hamming distance)
//   DBoW3::BowVector::addWeight/normalize   reference thirdparty/DBoW3/src/BowVector.cpp:29-78
//   Frame::ComputeBoW (levelsup = 4)        reference src/Basic/Frame.cpp:190-201
//   Matcher::SearchByBoW                    reference src/Algorithm/Matcher.cpp:196-292, ComputeThreeMaxima :294-336
// Unlike the rest of the oracle, DBoW3 IS vendored in the reference tree, so this file follows real source; it cannot be
// compiled there (it needs OpenCV C++), so the restatement is checked by an independent numpy descent in tests/test_bow.py
// over the reference's own vocab/ORBvoc.bin when that file is present.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "oracle.h"

namespace {

struct Node {
    int parent = 0;
    std::vector<int> children;
    uint8_t desc[32] = {0};
    double weight = 0;
    int word_id = -1;
};

}  // namespace

struct ora_vocab {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<Node> nodes;
    int n_words = 0;
};

namespace {

inline int hamming256(const uint8_t* a, const uint8_t* b) {
    uint64_t x[4], y[4];
    std::memcpy(x, a, 32);
    std::memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) + __builtin_popcountll(x[3] ^ y[3]);
}

// Vocabulary::transform(feature, word_id, weight, nid, levelsup).  `node` is only written at level L - levelsup; the caller in
// the reference leaves it uninitialised when a leaf is shallower than that -- the leaf's own id is reported here instead.
void descend(const ora_vocab& v, const uint8_t* f, int levelsup, int* word, double* weight, int* node) {
    const int nid_level = v.L - levelsup;
    int nid = -1;
    if (nid_level <= 0) nid = 0;
    int final_id = 0, level = 0;
    do {
        ++level;
        const std::vector<int>& ch = v.nodes[final_id].children;
        final_id = ch[0];
        int best = hamming256(f, v.nodes[final_id].desc);
        for (size_t c = 1; c < ch.size(); ++c) {
            const int d = hamming256(f, v.nodes[ch[c]].desc);
            if (d < best) {
                best = d;
                final_id = ch[c];
            }
        }
        if (level == nid_level) nid = final_id;
    } while (!v.nodes[final_id].children.empty());
    if (nid < 0) nid = final_id;
    *word = v.nodes[final_id].word_id;
    *weight = v.nodes[final_id].weight;
    *node = nid;
}

}  // namespace

// The file: u32 nb_nodes (root included), u32 size_node, i32 k, i32 L, i32 scoring, i32 weighting, then records of size_node
// bytes {i32 parent, u8 descriptor[32], f32 weight, u8 is_leaf}.  The reference reads `while (!f.eof())`, so the last record is
// processed twice: node nb_nodes is a copy of node nb_nodes - 1 appended to the same parent (and, if a leaf, one more word).
// It can never win a descent (equal distance, later child, strict `<`); it is reproduced so that node and word counts agree.
extern "C" ora_vocab* ora_vocab_load(const uint8_t* bytes, size_t n_bytes) {
    if (!bytes || n_bytes < 24) return nullptr;
    uint32_t nb_nodes, size_node;
    int32_t hdr[4];
    std::memcpy(&nb_nodes, bytes, 4);
    std::memcpy(&size_node, bytes + 4, 4);
    std::memcpy(hdr, bytes + 8, 16);
    if (size_node < 41 || hdr[0] < 1 || hdr[1] < 1 || nb_nodes < 2) return nullptr;
    const size_t n_rec = (n_bytes - 24) / size_node;
    if (n_rec + 1 != nb_nodes) return nullptr;
    ora_vocab* v = new ora_vocab;
    v->k = hdr[0]; v->L = hdr[1]; v->scoring = hdr[2]; v->weighting = hdr[3];
    v->nodes.resize((size_t)nb_nodes + 1);
    for (size_t nid = 1; nid <= (size_t)nb_nodes; ++nid) {
        const uint8_t* rec = bytes + 24 + std::min(nid - 1, n_rec - 1) * size_node;   // nid == nb_nodes: the eof() repeat
        Node& nd = v->nodes[nid];
        int32_t parent;
        float w;
        std::memcpy(&parent, rec, 4);
        std::memcpy(nd.desc, rec + 4, 32);
        std::memcpy(&w, rec + 36, 4);
        if (parent < 0 || (size_t)parent >= nid) {   // a parent precedes its children in the file
            delete v;
            return nullptr;
        }
        nd.parent = parent;
        nd.weight = w;
        v->nodes[parent].children.push_back((int)nid);
        if (rec[40]) nd.word_id = v->n_words++;
    }
    // every childless node must be a word (the descent stops at children.empty())
    for (size_t nid = 1; nid <= (size_t)nb_nodes; ++nid)
        if (v->nodes[nid].children.empty() && v->nodes[nid].word_id < 0) {
            delete v;
            return nullptr;
        }
    if (v->nodes[0].children.empty()) {
        delete v;
        return nullptr;
    }
    return v;
}

extern "C" void ora_vocab_free(ora_vocab* v) { delete v; }

extern "C" void ora_vocab_info(const ora_vocab* v, int32_t* info /* k, L, scoring, weighting, nodes (root and repeat included), words */) {
    info[0] = v->k; info[1] = v->L; info[2] = v->scoring; info[3] = v->weighting;
    info[4] = (int32_t)v->nodes.size(); info[5] = v->n_words;
}

// Vocabulary::transform(features, BowVector&, FeatureVector&, levelsup).  Per feature: word id, node id at level L - levelsup and
// the word's weight (node = -1 when the word is stopped, weight <= 0: such a feature enters neither vector).  The BowVector
// comes back as (word id ascending, value); returns its size.
extern "C" int ora_bow_transform(const ora_vocab* v, int n, const uint8_t* desc, int levelsup, int32_t* word, int32_t* node, double* weight,
                                 int32_t* bow_word, double* bow_value) {
    std::map<uint32_t, double> bow;
    const bool sum = v->weighting == 0 || v->weighting == 1;   // TF_IDF, TF: addWeight; IDF, BINARY: addIfNotExist
    for (int i = 0; i < n; ++i) {
        int w_id, nid;
        double w;
        descend(*v, desc + 32 * (size_t)i, levelsup, &w_id, &w, &nid);
        word[i] = w_id;
        weight[i] = w;
        node[i] = w > 0 ? nid : -1;
        if (!(w > 0)) continue;
        auto it = bow.lower_bound((uint32_t)w_id);
        if (it != bow.end() && it->first == (uint32_t)w_id) {
            if (sum) it->second += w;
        } else {
            bow.insert(it, {(uint32_t)w_id, w});
        }
    }
    const bool must = v->scoring != 5;           // every scoring object but DOT_PRODUCT normalises
    const bool l2 = v->scoring == 1;             // L2_NORM -> L2, all others L1
    if (sum && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (auto& e : bow) e.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) {
            for (auto& e : bow) norm += std::fabs(e.second);
        } else {
            for (auto& e : bow) norm += e.second * e.second;
            norm = std::sqrt(norm);
        }
        if (norm > 0.0)
            for (auto& e : bow) e.second /= norm;
    }
    int q = 0;
    for (auto& e : bow) {
        bow_word[q] = (int32_t)e.first;
        bow_value[q] = e.second;
        ++q;
    }
    return q;
}

// Matcher::SearchByBoW with the feature vectors given as one node id per feature (index lists of a DBoW3 feature vector are
// ascending feature indices, so scanning key-frame 2 in index order visits a node's features in the reference's order).
// match12[i] = index in key-frame 2 or -1; returns the reference's return value (matches, minus -- with check_orientation --
// the ones outside the three dominant rotation bins, which the reference counts out but does not delete).
extern "C" int ora_search_by_bow(int n1, const uint8_t* desc1, const int32_t* node1, const float* angle1, int n2, const uint8_t* desc2,
                                 const int32_t* node2, const float* angle2, int th_low, float knn_ratio, int check_orientation,
                                 int32_t* match12) {
    constexpr int HISTO_LENGTH = 30;
    std::vector<int> rot_hist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int cnt = 0;
    // the reference walks the common nodes in ascending node id and, inside a node, key-frame 1's features in index order: the
    // histogram's push order only matters through the bin sizes, the matches themselves are order-free
    for (int i = 0; i < n1; ++i) {
        match12[i] = -1;
        if (node1[i] < 0) continue;
        int best1 = 256, best2 = 256, best_idx = -1;
        for (int j = 0; j < n2; ++j) {
            if (node2[j] != node1[i]) continue;
            const int dist = ora_descriptor_distance(desc1 + 32 * (size_t)i, desc2 + 32 * (size_t)j);
            if (dist < best1) {
                best2 = best1;
                best1 = dist;
                best_idx = j;
            } else if (dist < best2) {
                best2 = dist;
            }
        }
        if (best1 < th_low && float(best1) < knn_ratio * float(best2)) {
            match12[i] = best_idx;
            if (check_orientation) {
                float rot = angle1[i] - angle2[best_idx];
                if (rot < 0) rot += 360;
                int bin = (int)std::round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                if (bin >= 0 && bin < HISTO_LENGTH) rot_hist[bin].push_back(best_idx);
            }
            ++cnt;
        }
    }
    if (check_orientation) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)rot_hist[i].size();
            if (s > max1) {
                max3 = max2; max2 = max1; max1 = s;
                ind3 = ind2; ind2 = ind1; ind1 = i;
            } else if (s > max2) {
                max3 = max2; max2 = s;
                ind3 = ind2; ind2 = i;
            } else if (s > max3) {
                max3 = s;
                ind3 = i;
            }
        }
        if (max2 < 0.1f * (float)max1) {
            ind2 = -1;
            ind3 = -1;
        } else if (max3 < 0.1f * (float)max1) {
            ind3 = -1;
        }
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            cnt -= (int)rot_hist[i].size();
        }
    }
    return cnt;
}
