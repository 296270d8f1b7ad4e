// bow_oracle.cpp — CPU restatement of the bag-of-words transform that follows ORB extraction in the reference
// (SURVEY.md §8f rank 3).  TEST INFRASTRUCTURE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline leg may call it.
//
// The algorithm lives in the reference's vendored DBoW2 (Thirdparty/DBoW2/DBoW2); this file restates
//   TemplatedVocabulary::loadFromTextFile   TemplatedVocabulary.h:1338-1423   (node stream → tree, word ids)
//   TemplatedVocabulary::transform (tree)   TemplatedVocabulary.h:1218-1262   (greedy descent, first minimum wins)
//   TemplatedVocabulary::transform (image)  TemplatedVocabulary.h:1127-1194   (BowVector + FeatureVector, levelsup)
//   BowVector::addWeight / addIfNotExist / normalize   BowVector.cpp:32-85
//   FeatureVector::addFeature               FeatureVector.cpp:30-46
//   FORB::distance                          FORB.cpp:81-101
//   L1Scoring::score                        ScoringObject.cpp:23-68
// as called by Frame::ComputeBoW (src/Frame.cc:739-747: transform(desc, bow, feat, 4)).
// Pinned: tests/test_bow_oracle.py compares it with the reference's own DBoW2 sources compiled unmodified
// (oracle/_ref/libref_bow.so) on synthetic vocabularies written in the reference's text format.
//
// Defined where the reference is not: a leaf shallower than the level asked for by `levelsup` leaves the reference's
// NodeId uninitialised (TemplatedVocabulary.h:1149-1155); here it reads 0 (root), like the nid_level <= 0 case.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "orb_oracle.h"

namespace {

struct BowNode {
    uint32_t parent = 0;
    std::vector<uint32_t> children;
    uint8_t desc[32] = {0};
    double weight = 0.0;
    uint32_t word_id = 0;
};
struct BowVocab {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<BowNode> nodes;     // nodes[0] = root
    std::vector<uint32_t> words;    // word id → node id
};

int popcount256(const uint8_t *a, const uint8_t *b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

void add_node(BowVocab &v, uint32_t parent, int is_leaf, const uint8_t *desc, double weight) {
    const uint32_t nid = (uint32_t)v.nodes.size();
    v.nodes.emplace_back();
    BowNode &nd = v.nodes.back();
    nd.parent = parent;
    v.nodes[parent].children.push_back(nid);
    memcpy(v.nodes[nid].desc, desc, 32);
    v.nodes[nid].weight = weight;
    if (is_leaf > 0) {
        v.nodes[nid].word_id = (uint32_t)v.words.size();
        v.words.push_back(nid);
    }
}

// one descriptor down the tree; *nid = node on the way at level L - levelsup
void descend(const BowVocab &v, const uint8_t *f, int levelsup, uint32_t *word, double *weight, uint32_t *nid) {
    const int nid_level = v.L - levelsup;
    *nid = 0;
    uint32_t cur = 0;
    int level = 0;
    do {
        ++level;
        const std::vector<uint32_t> &ch = v.nodes[cur].children;
        cur = ch[0];
        int best = popcount256(f, v.nodes[cur].desc);
        for (size_t i = 1; i < ch.size(); ++i) {
            const int d = popcount256(f, v.nodes[ch[i]].desc);
            if (d < best) { best = d; cur = ch[i]; }
        }
        if (level == nid_level) *nid = cur;
    } while (!v.nodes[cur].children.empty());
    *word = v.nodes[cur].word_id;
    *weight = v.nodes[cur].weight;
}

}  // namespace

extern "C" {

void *orc_vocab_from_nodes(const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc, const double *weight, int n_nodes,
                           int k, int L, int scoring, int weighting) {
    BowVocab *v = new BowVocab();
    v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->nodes.resize(1);
    v->nodes.reserve((size_t)n_nodes + 1);
    for (int i = 0; i < n_nodes; ++i) {
        if (parent[i] < 0 || parent[i] > i) { delete v; return nullptr; }   // a parent must already exist
        add_node(*v, (uint32_t)parent[i], is_leaf[i], desc + (size_t)i * 32, weight[i]);
    }
    return v;
}

void *orc_vocab_load_text(const char *path) {
    std::ifstream f(path);
    if (!f.good()) return nullptr;
    std::string line;
    if (!std::getline(f, line)) return nullptr;
    BowVocab *v = new BowVocab();
    {
        std::stringstream ss(line);
        ss >> v->k >> v->L >> v->scoring >> v->weighting;
        if (ss.fail() || v->k < 0 || v->k > 20 || v->L < 1 || v->L > 10 || v->scoring < 0 || v->scoring > 5 || v->weighting < 0 || v->weighting > 3) {
            delete v;
            return nullptr;
        }
    }
    v->nodes.resize(1);
    while (std::getline(f, line)) {
        if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;   // the reference reads a phantom node here
        std::stringstream ss(line);
        int pid = 0, leaf = 0;
        ss >> pid >> leaf;
        uint8_t d[32] = {0};
        for (int i = 0; i < 32; ++i) { int b = 0; ss >> b; if (!ss.fail()) d[i] = (uint8_t)b; }
        double w = 0.0;
        ss >> w;
        if (pid < 0 || pid >= (int)v->nodes.size()) { delete v; return nullptr; }
        add_node(*v, (uint32_t)pid, leaf, d, w);
    }
    return v;
}

void orc_vocab_free(void *v) { delete (BowVocab *)v; }
int orc_vocab_words(void *v) { return (int)((BowVocab *)v)->words.size(); }
int orc_vocab_nodes(void *v) { return (int)((BowVocab *)v)->nodes.size(); }

int orc_bow_transform(void *vp, const uint8_t *desc, int n, int levelsup, uint32_t *word_id, uint32_t *node_id, uint32_t *bow_ids,
                      double *bow_vals, int *n_bow, uint32_t *fv_nodes, int32_t *fv_off, uint32_t *fv_idx, int *n_fv) {
    const BowVocab &v = *(const BowVocab *)vp;
    std::map<uint32_t, double> bow;
    std::map<uint32_t, std::vector<uint32_t>> fv;
    *n_bow = 0; *n_fv = 0;
    fv_off[0] = 0;
    if (v.nodes.size() <= 1) return 0;
    const bool must = v.scoring != 5;                 // every scoring but DOT_PRODUCT normalises
    const bool l2 = v.scoring == 1;
    const bool tf = v.weighting == 0 || v.weighting == 1;
    for (int i = 0; i < n; ++i) {
        uint32_t w = 0, nid = 0;
        double wt = 0.0;
        descend(v, desc + (size_t)i * 32, levelsup, &w, &wt, &nid);
        if (word_id) word_id[i] = w;
        if (node_id) node_id[i] = nid;
        if (wt > 0) {                                  // stopped words (weight 0) are skipped
            std::map<uint32_t, double>::iterator it = bow.lower_bound(w);
            if (it != bow.end() && it->first == w) { if (tf) it->second += wt; }
            else bow.insert(it, std::make_pair(w, wt));
            fv[nid].push_back((uint32_t)i);
        }
    }
    if (tf && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (auto &e : bow) e.second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (!l2) { for (auto &e : bow) norm += fabs(e.second); }
        else { for (auto &e : bow) norm += e.second * e.second; norm = sqrt(norm); }
        if (norm > 0.0) for (auto &e : bow) e.second /= norm;
    }
    int i = 0;
    for (auto &e : bow) { bow_ids[i] = e.first; bow_vals[i] = e.second; ++i; }
    *n_bow = i;
    int k = 0, o = 0;
    for (auto &e : fv) {
        fv_nodes[k] = e.first;
        fv_off[k] = o;
        for (uint32_t j : e.second) fv_idx[o++] = j;
        ++k;
    }
    fv_off[k] = o;
    *n_fv = k;
    return 0;
}

// ---- cv::undistortPoints as Frame::UndistortKeyPoints / ComputeImageBounds call it (src/Frame.cc:749-811) ----
// OpenCV is not vendored; restated from its documented model (distortion k1 k2 p1 p2 k3 [k4 k5 k6 s1..s4], five fixed-point
// iterations in double precision, then P = K) and pinned bit-for-bit against cv2 4.13.0 in tests/test_undistort.py.
static void undistort_one(double px, double py, double fx, double fy, double cx, double cy, const double *k, float *ox, float *oy) {
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    double x = (px - cx) * ifx, y = (py - cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }      // test for a zero crossing of the rational model
        const double dX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double dY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    const double xx = fx * x + 0.0 * y + cx, yy = 0.0 * x + fy * y + cy, ww = 1.0 / (0.0 * x + 0.0 * y + 1.0);
    *ox = (float)(xx * ww);
    *oy = (float)(yy * ww);
}

void orc_undistort_points(const float *xy, int n, float fx, float fy, float cx, float cy, const float *D, int nd, float *out) {
    if (nd <= 0 || D[0] == 0.0f) { memcpy(out, xy, (size_t)n * 8); return; }
    double k[14] = {0};
    for (int i = 0; i < nd && i < 14; ++i) k[i] = D[i];
    for (int i = 0; i < n; ++i) undistort_one(xy[2 * i], xy[2 * i + 1], fx, fy, cx, cy, k, &out[2 * i], &out[2 * i + 1]);
}

void orc_image_bounds(int cols, int rows, float fx, float fy, float cx, float cy, const float *D, int nd, float *bounds) {
    if (nd > 0 && D[0] != 0.0f) {
        const float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
        float u[8];
        orc_undistort_points(c, 4, fx, fy, cx, cy, D, nd, u);
        bounds[0] = u[0] < u[4] ? u[0] : u[4];      // min(x0, x2)
        bounds[1] = u[2] > u[6] ? u[2] : u[6];      // max(x1, x3)
        bounds[2] = u[1] < u[3] ? u[1] : u[3];      // min(y0, y1)
        bounds[3] = u[5] > u[7] ? u[5] : u[7];      // max(y2, y3)
    } else {
        bounds[0] = 0.f; bounds[1] = (float)cols; bounds[2] = 0.f; bounds[3] = (float)rows;
    }
}

double orc_bow_score_l1(const uint32_t *ids1, const double *v1, int n1, const uint32_t *ids2, const double *v2, int n2) {
    double score = 0;
    int i = 0, j = 0;
    while (i < n1 && j < n2) {
        if (ids1[i] == ids2[j]) { score += fabs(v1[i] - v2[j]) - fabs(v1[i]) - fabs(v2[j]); ++i; ++j; }
        else if (ids1[i] < ids2[j]) ++i;               // lower_bound on a sorted sequence = skip the smaller keys
        else ++j;
    }
    return -score / 2.0;
}

}  // extern "C"
