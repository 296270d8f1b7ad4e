// orb_oracle.cpp — CPU ORACLE for the ORB front-end hot path.  TEST INFRASTRUCTURE, NOT PRODUCT.
//
// Restates the reference algorithm (file:line citations are relative to /root/reference) over scalar
// restatements of the OpenCV primitives it calls.  See orb_oracle.h for the
// rules about who may load this and for how the oracle is pinned (cv2 4.13.0 + oracle/_ref).
//
// Build: g++ -std=c++17 -O2 -ffp-contract=off -fPIC -shared (never -march=native / fast-math): the
// container's libm (cosf/sinf) and libstdc++ (std::sort tie order) are part of the definition.
#include "orb_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <list>
#include <thread>
#include <utility>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
inline int rnd(float v) { return (int)lrintf(v); }   // cvRound(float): round-half-even
inline int rnd(double v) { return (int)lrint(v); }   // cvRound(double)
inline int reflect101(int i, int n) {                // BORDER_REFLECT_101 index map (A2)
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

const int kEdge = 19;        // EDGE_THRESHOLD, ORBextractor.cc:73
const int kHalfPatch = 15;   // HALF_PATCH_SIZE, :72
const int kPatch = 31;       // PATCH_SIZE, :71

const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

// ------------------------------------------------------------------------------------------------
// A1: cv::resize INTER_LINEAR on 8UC1 (fixed point, 11-bit coefficients)
// ------------------------------------------------------------------------------------------------
void resize_linear(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw, int dh,
                   size_t dstep) {
    std::vector<int> xo(dw), yo(dh);
    std::vector<short> xa(2 * dw), ya(2 * dh);
    const double kx = (double)sw / dw, ky = (double)sh / dh;
    for (int d = 0; d < dw; ++d) {
        float f = (float)((d + 0.5) * kx - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sw - 1) { s = sw - 1; f = 0.f; }
        xo[d] = s;
        xa[2 * d] = (short)rnd((1.f - f) * 2048.f);
        xa[2 * d + 1] = (short)rnd(f * 2048.f);
    }
    for (int d = 0; d < dh; ++d) {
        float f = (float)((d + 0.5) * ky - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        yo[d] = s;  // rows are clipped when read; the fraction is kept (OpenCV's vertical pass)
        ya[2 * d] = (short)rnd((1.f - f) * 2048.f);
        ya[2 * d + 1] = (short)rnd(f * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    auto hrow = [&](int sy, std::vector<int> &out) {
        sy = std::min(std::max(sy, 0), sh - 1);
        const uint8_t *S = src + (size_t)sy * sstep;
        for (int d = 0; d < dw; ++d) {
            int s = xo[d], s1 = std::min(s + 1, sw - 1);
            out[d] = S[s] * xa[2 * d] + S[s1] * xa[2 * d + 1];
        }
    };
    int have0 = INT_MIN, have1 = INT_MIN;
    for (int y = 0; y < dh; ++y) {
        int s0 = std::min(std::max(yo[y], 0), sh - 1), s1 = std::min(std::max(yo[y] + 1, 0), sh - 1);
        if (have1 == s0) { r0.swap(r1); std::swap(have0, have1); }
        if (have0 != s0) { hrow(s0, r0); have0 = s0; }
        if (have1 != s1) { hrow(s1, r1); have1 = s1; }
        const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
        uint8_t *D = dst + (size_t)y * dstep;
        for (int d = 0; d < dw; ++d)
            D[d] = (uint8_t)((((b0 * (r0[d] >> 4)) >> 16) + ((b1 * (r1[d] >> 4)) >> 16) + 2) >> 2);
    }
}

// A2: copyMakeBorder(..., b,b,b,b, BORDER_REFLECT_101): dst is (w+2b)×(h+2b)
void border101(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep, int b) {
    for (int y = -b; y < h + b; ++y) {
        const uint8_t *S = src + (size_t)reflect101(y, h) * sstep;
        uint8_t *D = dst + (size_t)(y + b) * dstep;
        for (int x = -b; x < w + b; ++x) D[x + b] = S[reflect101(x, w)];
    }
}

// A3: GaussianBlur 7×7 σ=2 on 8U: separable integer kernel, sum 256, REFLECT_101
void gaussian7(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep) {
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t *S = src + (size_t)y * sstep;
        uint16_t *T = &tmp[(size_t)y * w];
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            if (x >= 3 && x + 3 < w)
                for (int i = 0; i < 7; ++i) acc += k[i] * S[x + i - 3];
            else
                for (int i = 0; i < 7; ++i) acc += k[i] * S[reflect101(x + i - 3, w)];
            T[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t *R[7];
        for (int j = 0; j < 7; ++j) R[j] = &tmp[(size_t)reflect101(y + j - 3, h) * w];
        uint8_t *D = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int j = 0; j < 7; ++j) acc += (uint32_t)k[j] * R[j][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// A4: cv::FAST(roi, kps, t, nonmaxSuppression=true), TYPE_9_16
// ------------------------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// M = max over the 16 arcs of 9 contiguous ring pixels of min(d) (brighter centre) and of min(-d)
inline int fast_arc_measure(const uint8_t *p, const ptrdiff_t *off) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = INT_MIN;
    for (int k = 0; k < 16; ++k) {
        int lo = d[k], hi = d[k];
        for (int j = 1; j < 9; ++j) {
            lo = std::min(lo, d[k + j]);
            hi = std::max(hi, d[k + j]);
        }
        best = std::max(best, std::max(lo, -hi));
    }
    return best;
}

struct FastHit { int x, y, score; };

void fast9_nms(const uint8_t *roi, int w, int h, size_t step, int t, std::vector<FastHit> &out) {
    out.clear();
    if (w < 7 || h < 7) return;
    ptrdiff_t off[16];
    for (int k = 0; k < 16; ++k) off[k] = kRingDy[k] * (ptrdiff_t)step + kRingDx[k];
    const int iw = w - 6, ih = h - 6;  // interior [3,w-3)×[3,h-3)
    // score map with a 1-pixel zero frame so NMS needs no bounds tests
    std::vector<int> sc((size_t)(iw + 2) * (ih + 2), 0);
    for (int y = 0; y < ih; ++y) {
        const uint8_t *row = roi + (size_t)(y + 3) * step + 3;
        int *S = &sc[(size_t)(y + 1) * (iw + 2) + 1];
        for (int x = 0; x < iw; ++x) {
            const uint8_t *p = row + x;
            const int v = p[0];
            // any 9-arc contains one pixel of every opposite pair: cheap necessary conditions
            const int hiT = v + t, loT = v - t;
            bool brighter = true, darker = true;  // ring brighter / darker than the centre
            for (int k = 0; k < 8 && (brighter || darker); k += 2) {
                const int a = p[off[k]], b = p[off[k + 8]];
                brighter = brighter && (a > hiT || b > hiT);
                darker = darker && (a < loT || b < loT);
            }
            if (!brighter && !darker) continue;
            const int m = fast_arc_measure(p, off);
            if (m > t) S[x] = m - 1;
        }
    }
    const int sw = iw + 2;
    for (int y = 0; y < ih; ++y) {
        const int *S = &sc[(size_t)(y + 1) * sw + 1];
        for (int x = 0; x < iw; ++x) {
            const int s = S[x];
            if (s <= 0) continue;  // a score-0 corner (t==0) can never beat its neighbours
            if (s > S[x - 1] && s > S[x + 1] && s > S[x - sw - 1] && s > S[x - sw] &&
                s > S[x - sw + 1] && s > S[x + sw - 1] && s > S[x + sw] && s > S[x + sw + 1])
                out.push_back({x + 3, y + 3, s});
        }
    }
}

// A5: cv::fastAtan2 (degrees), every operation individually rounded to fp32
float fast_atan2(float y, float x) {
    const float k = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k,
                p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ------------------------------------------------------------------------------------------------
// image plane helper
// ------------------------------------------------------------------------------------------------
struct Plane {
    int w = 0, h = 0;             // unpadded size
    size_t step = 0;              // of the padded buffer
    std::vector<uint8_t> buf;     // (w+38)×(h+38), REFLECT_101 border
    const uint8_t *roi() const { return buf.data() + (size_t)kEdge * step + kEdge; }
    uint8_t *roi() { return buf.data() + (size_t)kEdge * step + kEdge; }
    void alloc(int W, int H) {
        w = W; h = H; step = (size_t)W + 2 * kEdge;
        buf.assign(step * (size_t)(H + 2 * kEdge), 0);
    }
};

// ------------------------------------------------------------------------------------------------
// quadtree (ORBextractor.cc:480-779)
// ------------------------------------------------------------------------------------------------
struct QNode {
    int x0, x1, y0, y1;          // UL.x, UR.x, UL.y, BL.y
    std::vector<int> pts;        // candidate indices, insertion order
    bool leaf = false;           // bNoMore
    std::list<QNode>::iterator self;
};

// DivideNode (:480-536)
void split_node(const QNode &n, const std::vector<orc_keypoint> &c, QNode ch[4]) {
    const int hx = (int)std::ceil(static_cast<float>(n.x1 - n.x0) / 2);
    const int hy = (int)std::ceil(static_cast<float>(n.y1 - n.y0) / 2);
    const int xm = n.x0 + hx, ym = n.y0 + hy;
    ch[0].x0 = n.x0; ch[0].x1 = xm;   ch[0].y0 = n.y0; ch[0].y1 = ym;
    ch[1].x0 = xm;   ch[1].x1 = n.x1; ch[1].y0 = n.y0; ch[1].y1 = ym;
    ch[2].x0 = n.x0; ch[2].x1 = xm;   ch[2].y0 = ym;   ch[2].y1 = n.y1;
    ch[3].x0 = xm;   ch[3].x1 = n.x1; ch[3].y0 = ym;   ch[3].y1 = n.y1;
    for (int i : n.pts) {
        const orc_keypoint &k = c[i];
        const int q = (k.x < (float)xm ? 0 : 1) + (k.y < (float)ym ? 0 : 2);
        ch[q].pts.push_back(i);
    }
    for (int q = 0; q < 4; ++q) ch[q].leaf = ch[q].pts.size() == 1;
}

typedef std::pair<int, QNode *> SizedNode;
bool node_less(SizedNode &a, SizedNode &b) {  // compareNodes (:538-553)
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->x0 < b.second->x0;
}

// returns -3 when the region is so tall that the reference would index an empty root vector
int distribute(const std::vector<orc_keypoint> &cand, int minX, int maxX, int minY, int maxY, int N,
               std::vector<orc_keypoint> &out) {
    out.clear();
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
    if (nIni < 1) return cand.empty() ? 0 : -3;
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<QNode> nodes;
    std::vector<QNode *> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        QNode r;
        r.x0 = (int)(hX * static_cast<float>(i));
        r.x1 = (int)(hX * static_cast<float>(i + 1));
        r.y0 = 0;
        r.y1 = maxY - minY;
        nodes.push_back(r);
        roots[i] = &nodes.back();
    }
    for (size_t i = 0; i < cand.size(); ++i) {
        size_t b = (size_t)(cand[i].x / hX);
        if (b >= roots.size()) return -3;  // reference: out-of-bounds write (UB)
        roots[b]->pts.push_back((int)i);
    }
    for (auto it = nodes.begin(); it != nodes.end();) {
        if (it->pts.size() == 1) { it->leaf = true; ++it; }
        else if (it->pts.empty()) it = nodes.erase(it);
        else ++it;
    }

    std::vector<SizedNode> pending;  // children with >1 point created by the last pass
    // pushes the non-empty children to the list front in order 0..3 (:637-676)
    auto adopt = [&](QNode ch[4], int *nExpand) {
        for (int q = 0; q < 4; ++q) {
            if (ch[q].pts.empty()) continue;
            nodes.push_front(std::move(ch[q]));
            if (nodes.front().pts.size() > 1) {
                if (nExpand) ++*nExpand;
                pending.push_back(std::make_pair((int)nodes.front().pts.size(), &nodes.front()));
                nodes.front().self = nodes.begin();
            }
        }
    };

    bool done = false;
    while (!done) {
        int before = (int)nodes.size();
        int nExpand = 0;
        pending.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {
            if (it->leaf) { ++it; continue; }
            QNode ch[4];
            split_node(*it, cand, ch);
            adopt(ch, &nExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == before) {
            done = true;
        } else if ((int)nodes.size() + nExpand * 3 > N) {
            while (!done) {  // "largest first" phase (:685-751)
                before = (int)nodes.size();
                std::vector<SizedNode> work = pending;
                pending.clear();
                std::sort(work.begin(), work.end(), node_less);
                for (int j = (int)work.size() - 1; j >= 0; --j) {
                    QNode ch[4];
                    split_node(*work[j].second, cand, ch);
                    adopt(ch, nullptr);
                    nodes.erase(work[j].second->self);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == before) done = true;
            }
        }
    }
    // best response per node, first maximum wins (:757-776)
    out.reserve(nodes.size());
    for (const QNode &n : nodes) {
        int best = n.pts[0];
        float br = cand[best].response;
        for (size_t k = 1; k < n.pts.size(); ++k)
            if (cand[n.pts[k]].response > br) { best = n.pts[k]; br = cand[best].response; }
        out.push_back(cand[best]);
    }
    return (int)out.size();
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// extractor
// ------------------------------------------------------------------------------------------------
struct orc_extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;
    std::vector<float> sf, inv, sig2, invsig2;
    std::vector<int> quota;
    int umax[kHalfPatch + 1];
    std::vector<Plane> pyr;
    std::vector<std::vector<uint8_t>> blurred;            // unpadded, step = w
    std::vector<std::vector<orc_keypoint>> cand, sel;     // per level taps
};

extern "C" {

orc_extractor *orc_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    if (nlevels < 1 || nfeatures < 0) return nullptr;
    orc_extractor *e = new orc_extractor;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = ini_th; e->minTh = min_th;
    e->scaleFactor = scale_factor;  // the reference member is a double initialised from a float
    // :415-431 — per-level scale tables, fp32 products of a double scaleFactor rounded to float
    e->sf.assign(nlevels, 1.f); e->sig2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        e->sf[i] = (float)(e->sf[i - 1] * e->scaleFactor);
        e->sig2[i] = e->sf[i] * e->sf[i];
    }
    e->inv.resize(nlevels); e->invsig2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) { e->inv[i] = 1.0f / e->sf[i]; e->invsig2[i] = 1.0f / e->sig2[i]; }
    // :435-447 — per-level quotas
    e->quota.assign(nlevels, 0);
    float factor = (float)(1.0f / e->scaleFactor);
    float want = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        e->quota[l] = rnd(want);
        sum += e->quota[l];
        want *= factor;
    }
    e->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // :453-468 — end of each row of the circular patch
    int v, v0;
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= vmax; ++v) e->umax[v] = rnd(std::sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    e->pyr.resize(nlevels); e->blurred.resize(nlevels); e->cand.resize(nlevels); e->sel.resize(nlevels);
    return e;
}

void orc_destroy(orc_extractor *ex) { delete ex; }

void orc_params(const orc_extractor *e, float *sf, float *inv_sf, float *sigma2, float *inv_sigma2,
                int *quota, int *umax16) {
    for (int i = 0; i < e->nlevels; ++i) {
        if (sf) sf[i] = e->sf[i];
        if (inv_sf) inv_sf[i] = e->inv[i];
        if (sigma2) sigma2[i] = e->sig2[i];
        if (inv_sigma2) inv_sigma2[i] = e->invsig2[i];
        if (quota) quota[i] = e->quota[i];
    }
    if (umax16) for (int i = 0; i <= kHalfPatch; ++i) umax16[i] = e->umax[i];
}

// IC_Angle (:76-103) on the padded plane
static float ic_angle(const Plane &P, float px, float py, const int *umax) {
    int m01 = 0, m10 = 0;
    const int step = (int)P.step;
    const uint8_t *c = P.roi() + (ptrdiff_t)rnd(py) * step + rnd(px);
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vs = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int lo = c[u + v * step], up = c[u - v * step];
            vs += lo - up;
            m10 += u * (lo + up);
        }
        m01 += v * vs;
    }
    return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor (:107-146) on the blurred unpadded level
static void rbrief(const uint8_t *img, int step, const orc_keypoint &kp, uint8_t *desc) {
    const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);  // :106
    const float angle = kp.angle * factorPI;
    const float a = std::cos(angle), b = std::sin(angle);  // float overloads = cosf/sinf (H3)
    const uint8_t *c = img + (ptrdiff_t)rnd(kp.y) * step + rnd(kp.x);
    for (int i = 0; i < 32; ++i) {
        int byte = 0;
        for (int j = 0; j < 8; ++j) {
            const int8_t *t = &kPattern[(i * 8 + j) * 4];
            const float x0 = t[0], y0 = t[1], x1 = t[2], y1 = t[3];
            const int v0 = c[rnd(x0 * b + y0 * a) * step + rnd(x0 * a - y0 * b)];
            const int v1 = c[rnd(x1 * b + y1 * a) * step + rnd(x1 * a - y1 * b)];
            byte |= (v0 < v1) << j;
        }
        desc[i] = (uint8_t)byte;
    }
}

int orc_extract(orc_extractor *e, const uint8_t *img, int rows, int cols, size_t step,
                const int32_t *rects, int n_rects, int lap0, int lap1, orc_keypoint *kps,
                uint8_t *desc, int cap, int *n_out, int *mono_index) {
    if (n_out) *n_out = 0;
    if (mono_index) *mono_index = -1;
    if (!img || rows <= 0 || cols <= 0) return -1;  // :1129-1130
    const int L = e->nlevels;

    // ---- ComputePyramid (:1209-1234) ----
    for (int l = 0; l < L; ++l) {
        const float s = e->inv[l];
        const int w = rnd((float)cols * s), h = rnd((float)rows * s);
        Plane &P = e->pyr[l];
        P.alloc(w, h);
        if (l == 0) {
            for (int y = 0; y < h; ++y) memcpy(P.roi() + (size_t)y * P.step, img + (size_t)y * step, w);
        } else {
            const Plane &Q = e->pyr[l - 1];
            resize_linear(Q.roi(), Q.w, Q.h, Q.step, P.roi(), w, h, P.step);
        }
        std::vector<uint8_t> tmp((size_t)w * h);
        for (int y = 0; y < h; ++y) memcpy(&tmp[(size_t)y * w], P.roi() + (size_t)y * P.step, w);
        border101(tmp.data(), w, h, w, P.buf.data(), P.step, kEdge);
    }

    // ---- ComputeKeyPointsOctTree (:781-935) ----
    const float W = 35;
    std::vector<FastHit> hits;
    for (int l = 0; l < L; ++l) {
        const Plane &P = e->pyr[l];
        const int minBX = kEdge - 3, minBY = minBX;
        const int maxBX = P.w - kEdge + 3, maxBY = P.h - kEdge + 3;
        std::vector<orc_keypoint> &acc = e->cand[l];
        acc.clear();
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        const int wCell = nCols > 0 ? (int)std::ceil(width / nCols) : 0;
        const int hCell = nRows > 0 ? (int)std::ceil(height / nRows) : 0;
        const float scale = e->sf[l];
        for (int i = 0; i < nRows; ++i) {
            const float iniY = (float)(minBY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            for (int j = 0; j < nCols; ++j) {
                const float iniX = (float)(minBX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = (float)maxBX;
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                const uint8_t *roi = P.roi() + (ptrdiff_t)y0 * (ptrdiff_t)P.step + x0;
                fast9_nms(roi, cw, ch, P.step, e->iniTh, hits);
                if (hits.empty()) fast9_nms(roi, cw, ch, P.step, e->minTh, hits);
                for (const FastHit &hh : hits) {
                    orc_keypoint k;
                    k.x = (float)hh.x; k.y = (float)hh.y;
                    k.size = 7.f; k.angle = -1.f; k.response = (float)hh.score;
                    k.octave = 0; k.class_id = -1;
                    k.x += j * wCell;
                    k.y += i * hCell;
                    acc.push_back(k);
                }
                // DANI dynamic-area deletion over the whole accumulated list (:871-907)
                for (orc_keypoint &k : acc) {
                    k.x += minBX; k.y += minBY;
                    k.x *= scale; k.y *= scale;
                }
                if (n_rects > 0) {
                    size_t keep = 0;
                    for (size_t q = 0; q < acc.size(); ++q) {
                        const int px = rnd(acc[q].x), py = rnd(acc[q].y);  // Point2f→Point2i
                        bool hit = false;
                        for (int r = 0; r < n_rects && !hit; ++r) {
                            const int32_t *R = rects + 4 * r;
                            hit = R[0] <= px && px < R[0] + R[2] && R[1] <= py && py < R[1] + R[3];
                        }
                        if (!hit) acc[keep++] = acc[q];
                    }
                    acc.resize(keep);
                }
                const float scale_inverse = 1 / scale;
                for (orc_keypoint &k : acc) {
                    k.x *= scale_inverse; k.y *= scale_inverse;
                    k.x -= minBX; k.y -= minBY;
                }
            }
        }
        std::vector<orc_keypoint> &sel = e->sel[l];
        int rc = distribute(acc, minBX, maxBX, minBY, maxBY, e->quota[l], sel);
        if (rc < 0) return rc;
        const int scaledPatch = (int)(kPatch * e->sf[l]);
        for (orc_keypoint &k : sel) {
            k.x += minBX; k.y += minBY;
            k.octave = l;
            k.size = (float)scaledPatch;
        }
    }
    for (int l = 0; l < L; ++l)
        for (orc_keypoint &k : e->sel[l]) k.angle = ic_angle(e->pyr[l], k.x, k.y, e->umax);

    // ---- descriptors + output ordering (:1142-1206) ----
    int total = 0;
    for (int l = 0; l < L; ++l) total += (int)e->sel[l].size();
    if (n_out) *n_out = total;
    if (total > cap) return -2;
    int mono = 0, stereo = total - 1;
    for (int l = 0; l < L; ++l) {
        e->blurred[l].clear();
        std::vector<orc_keypoint> &sel = e->sel[l];
        if (sel.empty()) continue;
        const Plane &P = e->pyr[l];
        std::vector<uint8_t> work((size_t)P.w * P.h);
        for (int y = 0; y < P.h; ++y) memcpy(&work[(size_t)y * P.w], P.roi() + (size_t)y * P.step, P.w);
        e->blurred[l].resize(work.size());
        gaussian7(work.data(), P.w, P.h, P.w, e->blurred[l].data(), P.w);
        const float scale = e->sf[l];
        for (const orc_keypoint &k0 : sel) {
            uint8_t d[32];
            rbrief(e->blurred[l].data(), P.w, k0, d);
            orc_keypoint k = k0;
            if (l != 0) { k.x *= scale; k.y *= scale; }
            int at;
            if (k.x >= lap0 && k.x <= lap1) at = stereo--;
            else at = mono++;
            kps[at] = k;
            memcpy(desc + (size_t)at * 32, d, 32);
        }
    }
    if (mono_index) *mono_index = mono;
    return 0;
}

int orc_level_size(const orc_extractor *e, int level, int *w, int *h) {
    if (level < 0 || level >= e->nlevels) return -1;
    *w = e->pyr[level].w; *h = e->pyr[level].h;
    return 0;
}

int orc_get_level(const orc_extractor *e, int level, int padded, uint8_t *dst, size_t dst_step) {
    if (level < 0 || level >= e->nlevels) return -1;
    const Plane &P = e->pyr[level];
    if (padded) {
        for (int y = 0; y < P.h + 2 * kEdge; ++y)
            memcpy(dst + (size_t)y * dst_step, P.buf.data() + (size_t)y * P.step, P.w + 2 * kEdge);
    } else {
        for (int y = 0; y < P.h; ++y) memcpy(dst + (size_t)y * dst_step, P.roi() + (size_t)y * P.step, P.w);
    }
    return 0;
}

int orc_get_blurred(const orc_extractor *e, int level, uint8_t *dst, size_t dst_step) {
    if (level < 0 || level >= e->nlevels) return -1;
    const Plane &P = e->pyr[level];
    if (e->blurred[level].empty()) return 1;  // level had no keypoints: the reference skips the blur
    for (int y = 0; y < P.h; ++y) memcpy(dst + (size_t)y * dst_step, &e->blurred[level][(size_t)y * P.w], P.w);
    return 0;
}

static int copy_out(const std::vector<orc_keypoint> &v, orc_keypoint *out, int cap) {
    const int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = v[i];
    return n;
}
int orc_get_candidates(const orc_extractor *e, int level, orc_keypoint *out, int cap) {
    if (level < 0 || level >= e->nlevels) return -1;
    return copy_out(e->cand[level], out, cap);
}
int orc_get_selected(const orc_extractor *e, int level, orc_keypoint *out, int cap) {
    if (level < 0 || level >= e->nlevels) return -1;
    return copy_out(e->sel[level], out, cap);
}

// ---- primitives ----
void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t sstep, uint8_t *dst, int dw,
                          int dh, size_t dstep) {
    resize_linear(src, sw, sh, sstep, dst, dw, dh, dstep);
}
void orc_border_reflect101_u8(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst,
                              size_t dstep, int border) {
    border101(src, w, h, sstep, dst, dstep, border);
}
void orc_gaussian7_u8(const uint8_t *src, int w, int h, size_t sstep, uint8_t *dst, size_t dstep) {
    gaussian7(src, w, h, sstep, dst, dstep);
}
int orc_fast9_nms(const uint8_t *roi, int w, int h, size_t step, int threshold, int32_t *xys, int cap) {
    std::vector<FastHit> hits;
    fast9_nms(roi, w, h, step, threshold, hits);
    for (size_t i = 0; i < hits.size() && (int)i < cap; ++i) {
        xys[3 * i] = hits[i].x; xys[3 * i + 1] = hits[i].y; xys[3 * i + 2] = hits[i].score;
    }
    return (int)hits.size();
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
int orc_cvround(float v) { return rnd(v); }
void orc_sincos(float angle_rad, float *s, float *c) { *s = std::sin(angle_rad); *c = std::cos(angle_rad); }
void orc_sincos_array(const float *a, int64_t n, float *s, float *c, int nthreads) {
    // the host libm's sinf/cosf over an array (the definition the device port is swept against)
    auto work = [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; ++i) { s[i] = std::sin(a[i]); c[i] = std::cos(a[i]); } };
    if (nthreads <= 1) { work(0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
    for (auto &t : th) t.join();
}

int orc_distribute(const orc_keypoint *in, int n, int minX, int maxX, int minY, int maxY, int N,
                   orc_keypoint *out, int cap) {
    std::vector<orc_keypoint> c(in, in + n), o;
    int rc = distribute(c, minX, maxX, minY, maxY, N, o);
    if (rc < 0) return rc;
    for (int i = 0; i < rc && i < cap; ++i) out[i] = o[i];
    return rc;
}

void orc_sort_nodes(const int32_t *sizes, const int32_t *ulx, int n, int32_t *perm_out) {
    std::vector<QNode> nodes(n);
    std::vector<SizedNode> v(n);
    for (int i = 0; i < n; ++i) { nodes[i].x0 = ulx[i]; v[i] = std::make_pair((int)sizes[i], &nodes[i]); }
    std::sort(v.begin(), v.end(), node_less);
    for (int i = 0; i < n; ++i) perm_out[i] = (int32_t)(v[i].second - nodes.data());
}

// ---- matcher ----
int orc_descriptor_distance(const uint8_t *a, const uint8_t *b) {
    // ORBmatcher.cc:2054-2070: eight 32-bit words, SWAR bit count
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0x0F0F0F0Fu) * 0x01010101u) >> 24);
    }
    return dist;
}

static inline int ham256(const uint64_t *a, const uint64_t *b) {
    return __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
           __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
}

// A7: per query the two smallest distances, ascending, ties → lower train index first
void orc_knn2(const uint8_t *q, int nq, const uint8_t *db, int64_t ndb, int32_t *idx, int32_t *dist,
              int nthreads) {
    auto work = [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            uint64_t qa[4];
            memcpy(qa, q + (size_t)i * 32, 32);
            int d0 = INT_MAX, d1 = INT_MAX;
            int64_t i0 = -1, i1 = -1;
            for (int64_t j = 0; j < ndb; ++j) {
                uint64_t t[4];
                memcpy(t, db + (size_t)j * 32, 32);
                const int d = ham256(qa, t);
                if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
                else if (d < d1) { d1 = d; i1 = j; }
            }
            idx[2 * i] = (int32_t)i0; idx[2 * i + 1] = (int32_t)i1;
            dist[2 * i] = d0; dist[2 * i + 1] = d1;
        }
    };
    if (nthreads <= 1 || nq < 2) { work(0, nq); return; }
    nthreads = std::min(nthreads, nq);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back(work, (int)((int64_t)nq * t / nthreads), (int)((int64_t)nq * (t + 1) / nthreads));
    for (auto &t : th) t.join();
}

void orc_ratio_test(const int32_t *dist, int nq, double ratio, uint8_t *keep) {
    for (int i = 0; i < nq; ++i) {
        const bool two = dist[2 * i + 1] != INT_MAX;
        // Frame.cc:1085: float distance < float distance * double 0.7
        keep[i] = two && ((double)(float)dist[2 * i] < (double)(float)dist[2 * i + 1] * ratio);
    }
}

void orc_top2_lists(const uint8_t *q, int nq, const uint8_t *db, const int32_t *cand,
                    const int32_t *cand_off, int32_t *best_idx, int32_t *best_dist,
                    int32_t *second_idx, int32_t *second_dist) {
    // ORBmatcher.cc:84-121 shape: strict '<' so the first candidate wins ties; defaults 256 (a distance of 256
    // can therefore never be recorded).  second_idx is the candidate whose octave the reference keeps as
    // bestLevel2 (:110, :117); -1 when no second candidate was recorded.
    for (int i = 0; i < nq; ++i) {
        int b = 256, s = 256, bi = -1, si = -1;
        for (int c = cand_off[i]; c < cand_off[i + 1]; ++c) {
            const int d = orc_descriptor_distance(q + (size_t)i * 32, db + (size_t)cand[c] * 32);
            if (d < b) { s = b; si = bi; b = d; bi = cand[c]; }
            else if (d < s) { s = d; si = cand[c]; }
        }
        best_idx[i] = bi; best_dist[i] = b; second_dist[i] = s;
        if (second_idx) second_idx[i] = si;
    }
}

static void three_maxima(const std::vector<int> *h, int L, int &i1, int &i2, int &i3) {
    // ORBmatcher.cc:2008-2049
    int m1 = 0, m2 = 0, m3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = (int)h[i].size();
        if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
        else if (s > m3) { m3 = s; i3 = i; }
    }
    if (m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
    else if (m3 < 0.1f * (float)m1) { i3 = -1; }
}

void orc_three_maxima(const int32_t *counts, int L, int32_t *ind) {
    std::vector<std::vector<int>> h(L);
    for (int i = 0; i < L; ++i) h[i].assign(counts[i], 0);
    int a = -1, b = -1, c = -1;
    three_maxima(h.data(), L, a, b, c);
    ind[0] = a; ind[1] = b; ind[2] = c;
}

static inline int rot_bin(float a, float b) {
    // ORBmatcher.cc:345-350: rot in degrees, factor 1/30 (quirk Q10), round half away from zero
    const float factor = 1.0f / 30;
    float rot = a - b;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == 30) bin = 0;
    return bin;
}

void orc_rot_hist_filter(const float *angle_a, const float *angle_b, int n, uint8_t *keep) {
    std::vector<int> hist[30];
    for (int i = 0; i < n; ++i) hist[rot_bin(angle_a[i], angle_b[i])].push_back(i);
    int i1 = -1, i2 = -1, i3 = -1;
    three_maxima(hist, 30, i1, i2, i3);
    for (int i = 0; i < n; ++i) keep[i] = 1;
    for (int b = 0; b < 30; ++b) {
        if (b == i1 || b == i2 || b == i3) continue;
        for (int i : hist[b]) keep[i] = 0;
    }
}

int orc_search_init(const uint8_t *d1, const float *ang1, const int32_t *oct1, int n1,
                    const uint8_t *d2, const float *ang2, int n2, const int32_t *cand,
                    const int32_t *cand_off, float nnratio, int check_ori, int32_t *m12) {
    // ORBmatcher.cc:644-759 with the candidate lists (GetFeaturesInArea output) given explicitly
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) m12[i] = -1;
    std::vector<int> hist[30];
    std::vector<int> matchedDist(n2, INT_MAX), m21(n2, -1);
    for (int i1 = 0; i1 < n1; ++i1) {
        if (oct1[i1] > 0) continue;
        if (cand_off[i1] == cand_off[i1 + 1]) continue;
        int best = INT_MAX, best2 = INT_MAX, bi = -1;
        for (int c = cand_off[i1]; c < cand_off[i1 + 1]; ++c) {
            const int i2 = cand[c];
            const int d = orc_descriptor_distance(d1 + (size_t)i1 * 32, d2 + (size_t)i2 * 32);
            if (matchedDist[i2] <= d) continue;
            if (d < best) { best2 = best; best = d; bi = i2; }
            else if (d < best2) best2 = d;
        }
        if (best <= 50 && best < (float)best2 * nnratio) {
            if (m21[bi] >= 0) { m12[m21[bi]] = -1; --nmatches; }
            m12[i1] = bi; m21[bi] = i1; matchedDist[bi] = best; ++nmatches;
            if (check_ori) hist[rot_bin(ang1[i1], ang2[bi])].push_back(i1);
        }
    }
    if (check_ori) {
        int a = -1, b = -1, c = -1;
        three_maxima(hist, 30, a, b, c);
        for (int k = 0; k < 30; ++k) {
            if (k == a || k == b || k == c) continue;
            for (int i1 : hist[k])
                if (m12[i1] >= 0) { m12[i1] = -1; --nmatches; }
        }
    }
    return nmatches;
}


// ---- Frame grid (src/Frame.cc:387-418 AssignFeaturesToGrid, :727-738 PosInGrid, :659-725 GetFeaturesInArea) ----
// 64×48 cells (include/Frame.h:52-53) over [minX,maxX)×[minY,maxY); queries are (x, y, r) triples; candidate
// lists come out in the reference's order: cell column, then cell row, then insertion order inside the cell.
int orc_features_in_area(const float *xy, const int32_t *octave, int n, float minX, float minY, float maxX, float maxY,
                         const float *queries, int nq, int minLevel, int maxLevel, int32_t *cand_off, int32_t *cand, int cap) {
    const int COLS = 64, ROWS = 48;
    const float wInv = static_cast<float>(COLS) / static_cast<float>(maxX - minX);
    const float hInv = static_cast<float>(ROWS) / static_cast<float>(maxY - minY);
    std::vector<std::vector<int>> grid(COLS * ROWS);
    for (int i = 0; i < n; ++i) {
        const int px = (int)std::round((xy[2 * i] - minX) * wInv), py = (int)std::round((xy[2 * i + 1] - minY) * hInv);
        if (px < 0 || px >= COLS || py < 0 || py >= ROWS) continue;
        grid[px * ROWS + py].push_back(i);
    }
    int total = 0;
    cand_off[0] = 0;
    for (int q = 0; q < nq; ++q) {
        const float x = queries[3 * q], y = queries[3 * q + 1], r = queries[3 * q + 2];
        const int x0 = std::max(0, (int)std::floor((x - minX - r) * wInv)), x1 = std::min(COLS - 1, (int)std::ceil((x - minX + r) * wInv));
        const int y0 = std::max(0, (int)std::floor((y - minY - r) * hInv)), y1 = std::min(ROWS - 1, (int)std::ceil((y - minY + r) * hInv));
        if (!(x0 >= COLS || x1 < 0 || y0 >= ROWS || y1 < 0)) {
            const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = x0; ix <= x1; ++ix)
                for (int iy = y0; iy <= y1; ++iy)
                    for (int j : grid[ix * ROWS + iy]) {
                        if (checkLevels) {
                            if (octave[j] < minLevel) continue;
                            if (maxLevel >= 0 && octave[j] > maxLevel) continue;
                        }
                        const float dx = xy[2 * j] - x, dy = xy[2 * j + 1] - y;
                        if (std::fabs(dx) < r && std::fabs(dy) < r) {
                            if (total < cap) cand[total] = j;
                            ++total;
                        }
                    }
        }
        cand_off[q + 1] = total;
    }
    return total;
}

// ---- ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints) (src/ORBmatcher.cc:43-213),
// frames with Nleft == -1 (monocular, rectified stereo, RGB-D) ----
// Frame: xy = mvKeysUn[i].pt, octave, 32-byte descriptors, u_right = mvuRight (NULL: all -1), kp_obs[i] = Observations() of the
// map point already attached to keypoint i (-1: null pointer; NULL: all -1).  Map points (m rows, processed in order):
// proj5 = {mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos, mTrackDepth}, level = mnTrackScaleLevel, flags bit0 =
// mbTrackInView, bit1 = isBad(), n_obs = Observations().  assigned[i] = map point written into mvpMapPoints[i], or -1.
int orc_search_by_projection(const float *xy, const int32_t *octave, const uint8_t *desc, int n, const float *u_right, const int32_t *kp_obs,
                             float minX, float minY, float maxX, float maxY, const float *scale_factors, const float *mp_proj5,
                             const int32_t *mp_level, const uint8_t *mp_flags, const int32_t *mp_obs, const uint8_t *mp_desc, int m,
                             float nnratio, float th, int far_points, float th_far, int32_t *assigned) {
    int nmatches = 0;
    std::vector<int> obs(n, -1);
    if (kp_obs) obs.assign(kp_obs, kp_obs + n);
    for (int i = 0; i < n; ++i) assigned[i] = -1;
    const bool bFactor = th != 1.0;                                            // :47
    std::vector<int32_t> off(2), cand(std::max(n, 1));
    for (int j = 0; j < m; ++j) {
        if (!(mp_flags[j] & 1)) continue;                                      // :52 (mbTrackInViewR is false when Nleft == -1)
        if (far_points && mp_proj5[5 * j + 4] > th_far) continue;              // :55
        if (mp_flags[j] & 2) continue;                                         // :58
        const int lvl = mp_level[j];
        float r = (mp_proj5[5 * j + 3] > 0.998) ? 2.5f : 4.0f;                 // :215-221 (float against the double literal)
        if (bFactor) r *= th;                                                  // :68-69
        const float q[3] = {mp_proj5[5 * j], mp_proj5[5 * j + 1], r * scale_factors[lvl]};
        const int total = orc_features_in_area(xy, octave, n, minX, minY, maxX, maxY, q, 1, lvl - 1, lvl, off.data(), cand.data(), n);
        if (total == 0) continue;                                              // :74
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < total; ++c) {
            const int idx = cand[c];
            if (obs[idx] > 0) continue;                                        // :88-90
            if (u_right && u_right[idx] > 0) {                                 // :92-97
                const float er = std::fabs(mp_proj5[5 * j + 2] - u_right[idx]);
                if (er > r * scale_factors[lvl]) continue;
            }
            const int dist = orc_descriptor_distance(mp_desc + (size_t)j * 32, desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = octave[idx]; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = octave[idx]; bestDist2 = dist; }
        }
        if (bestDist <= 100) {                                                 // :124 TH_HIGH
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
                assigned[bestIdx] = j; obs[bestIdx] = mp_obs[j];               // :130: mvpMapPoints[bestIdx] = pMP
                ++nmatches;
            }
        }
    }
    return nmatches;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:222-425), frames with Nleft == -1 ----
// Both feature vectors as CSR over ascending vocabulary node ids (DBoW2::FeatureVector is a std::map): nodes[nn], off[nn+1], idx[].
// kf_mp[i]: 0 = keyframe feature i holds no map point, 1 = a good one, 2 = a bad one.  The walk is ordered: a frame feature that
// already received a map point is skipped by every later keyframe feature (:281-282).  assigned[i] = keyframe feature whose map
// point was written to vpMapPointMatches[i], or -1 (also after the rotation purge of :404-422).  Returns nmatches.
int orc_search_by_bow(const uint8_t *kf_desc, const float *kf_angle, int n_kf, const uint8_t *kf_mp, const int32_t *kf_nodes, const int32_t *kf_off,
                      const int32_t *kf_idx, int kf_nn, const uint8_t *f_desc, const float *f_angle, int n_f, const int32_t *f_nodes,
                      const int32_t *f_off, const int32_t *f_idx, int f_nn, float nnratio, int check_ori, int32_t *assigned) {
    (void)n_kf;
    for (int i = 0; i < n_f; ++i) assigned[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[30];
    int a = 0, b = 0;
    while (a < kf_nn && b < f_nn) {                                            // :243
        if (kf_nodes[a] == f_nodes[b]) {
            for (int iKF = kf_off[a]; iKF < kf_off[a + 1]; ++iKF) {
                const int realIdxKF = kf_idx[iKF];
                if (kf_mp[realIdxKF] != 1) continue;                           // :256-260 (null pointer or isBad())
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = f_off[b]; iF < f_off[b + 1]; ++iF) {
                    const int realIdxF = f_idx[iF];
                    if (assigned[realIdxF] >= 0) continue;                     // :281-282
                    const int dist = orc_descriptor_distance(kf_desc + (size_t)realIdxKF * 32, f_desc + (size_t)realIdxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                // :331-358; the right-camera half (:360-388) needs bestDist1R <= TH_LOW, which stays 256 when Nleft == -1
                if (bestDist1 <= 50 && (float)bestDist1 < nnratio * (float)bestDist2) {
                    assigned[bestIdxF] = realIdxKF;
                    if (check_ori) rotHist[rot_bin(kf_angle[realIdxKF], f_angle[bestIdxF])].push_back(bestIdxF);
                    ++nmatches;
                }
            }
            ++a; ++b;
        } else if (kf_nodes[a] < f_nodes[b]) {
            while (a < kf_nn && kf_nodes[a] < f_nodes[b]) ++a;                 // lower_bound(Fit->first)
        } else {
            while (b < f_nn && f_nodes[b] < kf_nodes[a]) ++b;
        }
    }
    if (check_ori) {                                                           // :404-422
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, 30, i1, i2, i3);
        for (int i = 0; i < 30; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int j : rotHist[i]) { assigned[j] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// ---- ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, vector<MapPoint*> &vpMatches12) (src/ORBmatcher.cc:760-901) ----
// Differences to the frame variant above: a candidate needs a good map point of its own and must not be matched yet (vbMatched2,
// :819-823), the threshold is bestDist1 < TH_LOW (strict, :843), the result is indexed by the FIRST keyframe's feature.
// matches12[i] = feature of keyframe 2 whose map point ends in vpMatches12[i], or -1.
int orc_search_by_bow_kf(const uint8_t *desc1, const float *angle1, int n1, const uint8_t *mp1, const int32_t *nodes1, const int32_t *off1,
                         const int32_t *idx1, int nn1, const uint8_t *desc2, const float *angle2, int n2, const uint8_t *mp2, const int32_t *nodes2,
                         const int32_t *off2, const int32_t *idx2, int nn2, float nnratio, int check_ori, int32_t *matches12) {
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<char> matched2(std::max(n2, 1), 0);
    int nmatches = 0;
    std::vector<int> rotHist[30];
    int a = 0, b = 0;
    while (a < nn1 && b < nn2) {
        if (nodes1[a] == nodes2[b]) {
            for (int i1 = off1[a]; i1 < off1[a + 1]; ++i1) {
                const int q = idx1[i1];
                if (mp1[q] != 1) continue;                                     // :801-805
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = off2[b]; i2 < off2[b + 1]; ++i2) {
                    const int t = idx2[i2];
                    if (matched2[t] || mp2[t] != 1) continue;                  // :819-826
                    const int dist = orc_descriptor_distance(desc1 + (size_t)q * 32, desc2 + (size_t)t * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = t; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < 50 && (float)bestDist1 < nnratio * (float)bestDist2) {   // :843-845
                    matches12[q] = bestIdx2;
                    matched2[bestIdx2] = 1;
                    if (check_ori) rotHist[rot_bin(angle1[q], angle2[bestIdx2])].push_back(q);
                    ++nmatches;
                }
            }
            ++a; ++b;
        } else if (nodes1[a] < nodes2[b]) {
            while (a < nn1 && nodes1[a] < nodes2[b]) ++a;
        } else {
            while (b < nn2 && nodes2[b] < nodes1[a]) ++b;
        }
    }
    if (check_ori) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, 30, i1, i2, i3);
        for (int i = 0; i < 30; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int j : rotHist[i]) { matches12[j] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// ---- classical rectified-stereo association (slot = Frame::ComputeStereoMatches, src/Frame.cc:813-915) ----
// PARITY UNPINNED: this tree replaced the function's matcher by LightGlue (src/Frame.cc:822-860), so there is no reference
// code to compile for it.  What follows restates the published algorithm of the upstream ORB-SLAM3 function of the same name
// (row bands of ±2·scale around every right keypoint, Hamming search among the band's keypoints within one octave and the
// disparity range, 11×11 SAD refinement over ±5 px on the keypoint's pyramid level with a parabola fit, disparity gate, and
// the median cut with the factor 1.5·1.4), ending in the same bookkeeping as this tree's tail (:862-914).  Level images are
// the unblurred, unpadded pyramid levels of the two extractors (mvImagePyramid).
int orc_stereo_rowband(const orc_extractor *exL, const orc_extractor *exR, const orc_keypoint *kL, const uint8_t *dL, int nL,
                       const orc_keypoint *kR, const uint8_t *dR, int nR, float mbf, float mb, float *uRight, float *depth) {
    for (int i = 0; i < nL; ++i) { uRight[i] = -1.f; depth[i] = -1.f; }
    if (nL == 0 || nR == 0) return 0;
    const int thOrbDist = (100 + 50) / 2;                       // (TH_HIGH + TH_LOW) / 2
    const int nRows = exL->pyr[0].h;
    std::vector<std::vector<int>> rowIdx(nRows);
    for (int iR = 0; iR < nR; ++iR) {
        const float kpY = kR[iR].y;
        const float r = 2.0f * exR->sf[kR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi)
            if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);   // (upstream indexes without the guard; keypoints stay 19 px inside)
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < nL; ++iL) {
        const int levelL = kL[iL].octave;
        const float vL = kL[iL].y, uL = kL[iL].x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<int> &cands = rowIdx[row];
        if (cands.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100, bestIdxR = 0;                       // TH_HIGH
        for (int iR : cands) {
            if (kR[iR].octave < levelL - 1 || kR[iR].octave > levelL + 1) continue;
            const float uR = kR[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orc_descriptor_distance(dL + (size_t)iL * 32, dR + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist >= thOrbDist) continue;
        // sub-pixel refinement by correlation on the keypoint's level
        const float uR0 = kR[bestIdxR].x;
        const float scaleFactor = exL->inv[levelL];
        const float scaleduL = std::round(uL * scaleFactor), scaledvL = std::round(vL * scaleFactor), scaleduR0 = std::round(uR0 * scaleFactor);
        const int w = 5, L = 5;
        const Plane &PL = exL->pyr[levelL], &PR = exR->pyr[levelL];
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
        if (iniu < 0 || endu >= PR.w) continue;
        const int cu = (int)scaleduL, cv = (int)scaledvL, cr = (int)scaleduR0;
        if (cv - w < 0 || cv + w >= PL.h || cu - w < 0 || cu + w >= PL.w || cr - L - w < 0) continue;   // (windows inside the levels; always true for extractor keypoints)
        int bestSad = INT_MAX, bestincR = 0;
        float vDists[2 * 5 + 1];
        for (int incR = -L; incR <= L; ++incR) {
            int sad = 0;
            for (int yy = -w; yy <= w; ++yy) {
                const uint8_t *a = PL.roi() + (size_t)(cv + yy) * PL.step + (cu - w);
                const uint8_t *b = PR.roi() + (size_t)(cv + yy) * PR.step + (cr + incR - w);
                for (int xx = 0; xx <= 2 * w; ++xx) sad += std::abs((int)a[xx] - (int)b[xx]);
            }
            if (sad < bestSad) { bestSad = sad; bestincR = incR; }
            vDists[L + incR] = (float)sad;
        }
        if (bestincR == -L || bestincR == L) continue;
        const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (deltaR < -1 || deltaR > 1) continue;
        float bestuR = exL->sf[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);
        float disparity = uL - bestuR;
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) { disparity = 0.01f; bestuR = uL - 0.01f; }
            depth[iL] = mbf / disparity;
            uRight[iL] = bestuR;
            vDistIdx.push_back(std::make_pair(bestSad, iL));
        }
    }
    if (vDistIdx.empty()) return 0;
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
        --kept;
    }
    return kept;
}

// ---- stereo association tail (src/Frame.cc:862-914) fed by the Hamming kNN + Lowe ratio of :1078-1085 ----
// For every left keypoint i with keep[i]: iR = idx[2i], distance = (float)dist[2i]; disparity gate [0, mbf/mb),
// depth = mbf/disparity (0.01 when disparity <= 0), then the 1.5·median distance cut.  uRight/depth are N_left
// arrays (−1 = no stereo).  NOTE: the reference feeds this tail from LightGlue (score 1−distance); the build pairs
// it with the Hamming kNN instead and both oracle and CUDA restate exactly this composition.
int orc_stereo_tail(const float *uL, const float *uR, int nL, int nR, const int32_t *idx, const int32_t *dist,
                    const uint8_t *keep, float mbf, float mb, float *uRight, float *depth) {
    for (int i = 0; i < nL; ++i) { uRight[i] = -1.f; depth[i] = -1.f; }
    const float minD = 0, maxD = mbf / mb;
    std::vector<std::pair<float, int>> v;
    for (int i = 0; i < nL; ++i) {
        if (!keep[i]) continue;
        const int iR = idx[2 * i];
        if (iR < 0 || iR >= nR) continue;
        float disparity = uL[i] - uR[iR];
        if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0) disparity = 0.01f;
            depth[i] = mbf / disparity;
            uRight[i] = uR[iR];
            v.push_back(std::make_pair((float)dist[2 * i], i));
        }
    }
    if (v.empty()) return 0;
    std::sort(v.begin(), v.end());
    const float median = v[v.size() / 2].first;
    const float thDist = 1.5f * median;
    int kept = (int)v.size();
    for (int i = (int)v.size() - 1; i >= 0; --i) {
        if (v[i].first < thDist) break;
        uRight[v[i].second] = -1;
        depth[v[i].second] = -1;
        --kept;
    }
    return kept;
}

}  // extern "C"
