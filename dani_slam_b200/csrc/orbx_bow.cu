// orbx_bow.cu — B200 (sm_100a) bag-of-words transform of ORB descriptors: the step Frame::ComputeBoW
// (/root/reference/src/Frame.cc:739-747) runs on every keyframe's descriptors with the reference's vendored DBoW2,
//   mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4)       Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1194
// SURVEY.md §8(f) rank 3.  Bit-exact (word ids, node ids, double weights) against oracle/bow_oracle.cpp, which is
// pinned to the reference's own DBoW2 sources compiled unmodified.
//
//   B1 k_bow_descend<GS>   TemplatedVocabulary::transform(feature, …) :1218-1262 — a group of GS lanes per descriptor
//                          walks the tree; lanes take the children of the current node, a packed (distance, position)
//                          min picks the first minimum (strict '<' in the reference), one level per round
//   B2 k_bow_vectors       BowVector::addWeight / addIfNotExist / normalize (BowVector.cpp:32-85) and
//                          FeatureVector::addFeature (FeatureVector.cpp:30-46) — one block per image: the std::map
//                          orders are a sort by (word id, feature index) and by (node id, feature index); weights of a
//                          word are added serially in feature order and the L1 / L2 norm serially in word order, in
//                          double precision without contraction, exactly like the map iteration does
//
// The vocabulary (k-ary tree, 256-bit node descriptors, per-word weights) is uploaded once per handle and stays
// resident in HBM: 1.1 M nodes x 32 B = 35 MB for an ORBvoc-sized tree (k = 10, L = 6), i.e. L2-resident on B200.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "orbx_internal.h"

namespace {

thread_local std::string tl_vocab_error;

#define VCUDA_TRY(v, call)                                                                   \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            (v)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                   \
            cudaGetLastError();                                                              \
            return ORBX_ERR_CUDA;                                                            \
        }                                                                                    \
    } while (0)

struct BowTree {     // device views
    const uint4 *desc;        // 2 x uint4 per node
    const int *childOff;      // CSR over child
    const int *child;
    const double *weight;
    const int *wordId;
    int L;
};

__device__ __forceinline__ int ham256(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// One group of GS lanes per descriptor.  Image b has n[b] descriptors at desc + b*descStride (bytes); outputs are
// indexed b*cap + i.  leaf = leaf node id (weight lookup later), nodeOut = ancestor at level L - levelsup (0 when the
// leaf is shallower or the level is <= 0).
template <int GS>
__global__ void __launch_bounds__(256, 8) k_bow_descend(BowTree t, const uint8_t *__restrict__ desc, size_t descStride, const int *__restrict__ nPer,
                                                     int cap, int levelsup, uint32_t *__restrict__ wordOut, uint32_t *__restrict__ nodeOut,
                                                     int *__restrict__ leafOut) {
    const int b = blockIdx.y;
    const int n = min(nPer[b], cap);
    const int gl = threadIdx.x % GS;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) / GS;
    if (i >= n) return;                       // whole groups leave together (GS divides 32 and blockDim)
    const unsigned gmask = GS == 32 ? 0xffffffffu : (((1u << GS) - 1u) << ((threadIdx.x & 31) / GS * GS));
    const uint4 *f = reinterpret_cast<const uint4 *>(desc + (size_t)b * descStride + (size_t)i * 32);
    const uint4 f0 = f[0], f1 = f[1];
    const int nidLevel = t.L - levelsup;
    int cur = 0, level = 0;
    uint32_t nid = 0;
    int c0 = t.childOff[0], c1 = t.childOff[1];
    while (c1 > c0) {
        ++level;
        uint32_t best = 0xffffffffu;          // (distance << 16) | position among the children
        for (int p = gl; p < c1 - c0; p += GS) {
            const int ch = t.child[c0 + p];
            const uint4 d0 = t.desc[2 * (size_t)ch], d1 = t.desc[2 * (size_t)ch + 1];
            best = min(best, ((uint32_t)ham256(f0, f1, d0, d1) << 16) | (uint32_t)p);
        }
#pragma unroll
        for (int o = GS / 2; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(gmask, best, o, GS));
        cur = t.child[c0 + (int)(best & 0xffffu)];
        if (level == nidLevel) nid = (uint32_t)cur;
        c0 = t.childOff[cur]; c1 = t.childOff[cur + 1];
    }
    if (gl == 0) {
        const size_t o = (size_t)b * cap + i;
        wordOut[o] = (uint32_t)t.wordId[cur];
        nodeOut[o] = nid;
        leafOut[o] = cur;
    }
}

// in-place bitonic sort of P (power of two) 64-bit keys in shared memory
__device__ void bitonic_sort_u64(unsigned long long *a, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

// ordered compaction of the segment heads of the sorted keys a[0..nValid): returns the number of heads; for head h
// (in order) headPos[h] = its position.  warpSum: blockDim/32 + 1 ints of shared scratch.
__device__ int segment_heads(const unsigned long long *a, int nValid, int *headPos, int *warpSum) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    int base = 0;
    for (int c0 = 0; c0 < nValid; c0 += blockDim.x) {
        const int j = c0 + threadIdx.x;
        const bool head = j < nValid && (j == 0 || (uint32_t)(a[j] >> 32) != (uint32_t)(a[j - 1] >> 32));
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) warpSum[warp] = __popc(bal);
        __syncthreads();
        if (warp == 0) {
            int v = lane < nWarps ? warpSum[lane] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            if (lane < nWarps) warpSum[lane] = incl - v;
            if (lane == 31) warpSum[nWarps] = incl;
        }
        __syncthreads();
        if (head) headPos[base + warpSum[warp] + __popc(bal & ((1u << lane) - 1u))] = j;
        base += warpSum[nWarps];
        __syncthreads();
    }
    return base;
}

// One block per image.  Shared memory: P keys (8 B) + P ints (head positions) + scratch.
__global__ void __launch_bounds__(256) k_bow_vectors(BowTree t, const int *__restrict__ nPer, int cap, int P, int tfMode, int normMode,
                                                     const uint32_t *__restrict__ wordIn, const uint32_t *__restrict__ nodeIn, const int *__restrict__ leafIn,
                                                     uint32_t *__restrict__ bowIds, double *__restrict__ bowVals, int *__restrict__ nBow,
                                                     uint32_t *__restrict__ fvNodes, int *__restrict__ fvOff, uint32_t *__restrict__ fvIdx, int *__restrict__ nFv) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem);
    int *headPos = reinterpret_cast<int *>(keys + P);
    int *warpSum = headPos + P;
    __shared__ int sValid;
    __shared__ double sNorm;
    const int b = blockIdx.x;
    const int n = min(nPer[b], cap);
    const size_t o = (size_t)b * cap;
    const unsigned long long SENT = ~0ull;

    // ---- BowVector: sort kept features by (word id, feature index) ----
    if (threadIdx.x == 0) sValid = 0;
    __syncthreads();
    int kept = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long key = SENT;
        if (i < n && t.weight[leafIn[o + i]] > 0.0) { key = ((unsigned long long)wordIn[o + i] << 32) | (unsigned)i; ++kept; }   // stopped words are skipped
        keys[i] = key;
    }
    atomicAdd(&sValid, kept);
    __syncthreads();
    const int nValid = sValid;
    bitonic_sort_u64(keys, P);
    int nb = segment_heads(keys, nValid, headPos, warpSum);
    for (int h = threadIdx.x; h < nb; h += blockDim.x) {
        const int j = headPos[h];
        const uint32_t w = (uint32_t)(keys[j] >> 32);
        const double wt = t.weight[leafIn[o + (uint32_t)keys[j]]];
        double s = wt;                                     // first feature of the word inserts, later ones add (TF modes)
        if (tfMode)
            for (int q = j + 1; q < nValid && (uint32_t)(keys[q] >> 32) == w; ++q) s = __dadd_rn(s, wt);
        bowIds[o + h] = w;
        bowVals[o + h] = s;
    }
    __syncthreads();
    // normalisation exactly as the std::map iteration does it: a serial sum in word order
    if (threadIdx.x == 0) {
        double norm = 0.0;
        if (normMode == 1) { for (int h = 0; h < nb; ++h) norm = __dadd_rn(norm, fabs(bowVals[o + h])); }
        else if (normMode == 2) { for (int h = 0; h < nb; ++h) { const double v = bowVals[o + h]; norm = __dadd_rn(norm, __dmul_rn(v, v)); } norm = sqrt(norm); }
        else if (tfMode && nb > 0) norm = (double)nb;      // no normalisation asked: term frequencies divided by the vector size
        sNorm = norm;
        nBow[b] = nb;
    }
    __syncthreads();
    if (sNorm > 0.0)
        for (int h = threadIdx.x; h < nb; h += blockDim.x) bowVals[o + h] = __ddiv_rn(bowVals[o + h], sNorm);
    __syncthreads();

    // ---- FeatureVector: sort kept features by (node id, feature index) ----
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long key = SENT;
        if (i < n && t.weight[leafIn[o + i]] > 0.0) key = ((unsigned long long)nodeIn[o + i] << 32) | (unsigned)i;
        keys[i] = key;
    }
    __syncthreads();
    bitonic_sort_u64(keys, P);
    const int nf = segment_heads(keys, nValid, headPos, warpSum);
    int *off = fvOff + (size_t)b * (cap + 1);
    for (int h = threadIdx.x; h < nf; h += blockDim.x) {
        fvNodes[o + h] = (uint32_t)(keys[headPos[h]] >> 32);
        off[h] = headPos[h];
    }
    for (int j = threadIdx.x; j < nValid; j += blockDim.x) fvIdx[o + j] = (uint32_t)keys[j];
    if (threadIdx.x == 0) { off[nf] = nValid; nFv[b] = nf; }
}

}  // namespace

struct orbx_vocab {
    int device = 0;
    int k = 0, L = 0, scoring = 0, weighting = 0;
    int nNodes = 0, nWords = 0, maxChildren = 0;
    uint4 *d_desc = nullptr;
    int *d_childOff = nullptr, *d_child = nullptr, *d_wordId = nullptr;
    double *d_weight = nullptr;
    cudaStream_t stream = nullptr;
    // scratch of the host-buffer entry point (one image), grown on demand
    int workCap = 0;
    uint8_t *w_desc = nullptr;
    uint32_t *w_word = nullptr, *w_node = nullptr, *w_bowIds = nullptr, *w_fvNodes = nullptr, *w_fvIdx = nullptr;
    int *w_leaf = nullptr, *w_fvOff = nullptr, *w_counts = nullptr;   // counts: n, nBow, nFv
    double *w_bowVals = nullptr;
    // leaf scratch of the device entry point
    int *b_leaf = nullptr;
    size_t b_leafCap = 0;
    std::mutex mu;          // the reference shares one const vocabulary between threads: calls on a handle serialise
    std::string err;
};

namespace {

struct HostTree {
    std::vector<int> parent;                  // per node (root: 0)
    std::vector<std::vector<int>> children;
    std::vector<uint8_t> desc;                // 32 B per node
    std::vector<double> weight;
    std::vector<int> wordId;
    int nWords = 0;
    HostTree() { parent.push_back(0); children.emplace_back(); desc.assign(32, 0); weight.push_back(0.0); wordId.push_back(0); }
    bool add(int p, int isLeaf, const uint8_t *d, double w) {
        const int nid = (int)parent.size();
        if (p < 0 || p >= nid) return false;
        parent.push_back(p); children.emplace_back(); children[p].push_back(nid);
        desc.insert(desc.end(), d, d + 32); weight.push_back(w);
        wordId.push_back(isLeaf > 0 ? nWords++ : 0);
        return true;
    }
};

orbx_vocab *upload(const HostTree &T, int k, int L, int scoring, int weighting, int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        tl_vocab_error = "orbx_vocab: no such CUDA device (liborbx has no CPU fallback)";
        return nullptr;
    }
    if (L < 1 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) { tl_vocab_error = "orbx_vocab: bad k/L/scoring/weighting"; return nullptr; }
    orbx_vocab *v = new orbx_vocab();
    v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting;
    v->nNodes = (int)T.parent.size(); v->nWords = T.nWords;
    std::vector<int> off(v->nNodes + 1, 0), child;
    child.reserve(v->nNodes);
    for (int i = 0; i < v->nNodes; ++i) {
        off[i] = (int)child.size();
        child.insert(child.end(), T.children[i].begin(), T.children[i].end());
        v->maxChildren = std::max(v->maxChildren, (int)T.children[i].size());
    }
    off[v->nNodes] = (int)child.size();
    if (v->maxChildren >= 65536) { tl_vocab_error = "orbx_vocab: more than 65535 children under one node"; delete v; return nullptr; }
    auto fail = [&](const char *what, cudaError_t e) { tl_vocab_error = std::string(what) + ": " + cudaGetErrorString(e); cudaGetLastError(); orbx_vocab_destroy(v); return (orbx_vocab *)nullptr; };
    cudaError_t e;
    OrbxDeviceGuard dg_(device);
    if ((e = dg_.status) != cudaSuccess) return fail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", e);
    if ((e = cudaMalloc((void **)&v->d_desc, (size_t)v->nNodes * 32)) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&v->d_childOff, (size_t)(v->nNodes + 1) * sizeof(int))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&v->d_child, std::max<size_t>(child.size(), 1) * sizeof(int))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&v->d_wordId, (size_t)v->nNodes * sizeof(int))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMalloc((void **)&v->d_weight, (size_t)v->nNodes * sizeof(double))) != cudaSuccess) return fail("cudaMalloc", e);
    if ((e = cudaMemcpy(v->d_desc, T.desc.data(), (size_t)v->nNodes * 32, cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if ((e = cudaMemcpy(v->d_childOff, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if (!child.empty() && (e = cudaMemcpy(v->d_child, child.data(), child.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if ((e = cudaMemcpy(v->d_wordId, T.wordId.data(), (size_t)v->nNodes * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    if ((e = cudaMemcpy(v->d_weight, T.weight.data(), (size_t)v->nNodes * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) return fail("cudaMemcpy", e);
    return v;
}

int launch_transform(orbx_vocab *v, const uint8_t *d_desc, size_t descStride, const int *d_n, int batch, int cap, int levelsup,
                     uint32_t *d_word, uint32_t *d_node, int *d_leaf, uint32_t *d_bowIds, double *d_bowVals, int *d_nBow,
                     uint32_t *d_fvNodes, int *d_fvOff, uint32_t *d_fvIdx, int *d_nFv) {
    if (batch <= 0 || cap <= 0) return ORBX_OK;
    if (v->nNodes <= 1) {      // empty vocabulary: transform() clears both vectors and returns (TemplatedVocabulary.h:1133-1136)
        VCUDA_TRY(v, cudaMemsetAsync(d_nBow, 0, (size_t)batch * sizeof(int), v->stream));
        VCUDA_TRY(v, cudaMemsetAsync(d_nFv, 0, (size_t)batch * sizeof(int), v->stream));
        VCUDA_TRY(v, cudaMemsetAsync(d_fvOff, 0, (size_t)batch * (cap + 1) * sizeof(int), v->stream));
        return ORBX_OK;
    }
    int P = 32;
    while (P < cap) P <<= 1;
    const size_t smem = (size_t)P * 12 + 64 * sizeof(int);
    if (smem > 200 * 1024) { v->err = "orbx_bow: more than 16384 descriptors per image"; return ORBX_ERR_CAPACITY; }
    BowTree t{v->d_desc, v->d_childOff, v->d_child, v->d_weight, v->d_wordId, v->L};
    const int gs = v->maxChildren <= 8 ? 8 : (v->maxChildren <= 16 ? 16 : 32);
    const int perBlock = 256 / gs;
    dim3 grid((cap + perBlock - 1) / perBlock, batch);
    if (gs == 8) k_bow_descend<8><<<grid, 256, 0, v->stream>>>(t, d_desc, descStride, d_n, cap, levelsup, d_word, d_node, d_leaf);
    else if (gs == 16) k_bow_descend<16><<<grid, 256, 0, v->stream>>>(t, d_desc, descStride, d_n, cap, levelsup, d_word, d_node, d_leaf);
    else k_bow_descend<32><<<grid, 256, 0, v->stream>>>(t, d_desc, descStride, d_n, cap, levelsup, d_word, d_node, d_leaf);
    const int tfMode = (v->weighting == 0 || v->weighting == 1) ? 1 : 0;
    const int normMode = v->scoring == 5 ? 0 : (v->scoring == 1 ? 2 : 1);   // DOT_PRODUCT: none; L2_NORM: L2; everything else L1
    if (smem > 48 * 1024) VCUDA_TRY(v, cudaFuncSetAttribute(k_bow_vectors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bow_vectors<<<batch, 256, smem, v->stream>>>(t, d_n, cap, P, tfMode, normMode, d_word, d_node, d_leaf, d_bowIds, d_bowVals, d_nBow,
                                                  d_fvNodes, d_fvOff, d_fvIdx, d_nFv);
    VCUDA_TRY(v, cudaGetLastError());
    return ORBX_OK;
}

}  // namespace

extern "C" {

orbx_vocab *orbx_vocab_create_from_nodes(const int32_t *parent, const uint8_t *is_leaf, const uint8_t *desc, const double *weight,
                                         int n_nodes, int k, int L, int scoring, int weighting, int device) {
    if (n_nodes < 0 || (n_nodes > 0 && (!parent || !is_leaf || !desc || !weight))) { tl_vocab_error = "orbx_vocab_create_from_nodes: null argument"; return nullptr; }
    HostTree T;
    for (int i = 0; i < n_nodes; ++i)
        if (!T.add(parent[i], is_leaf[i], desc + (size_t)i * 32, weight[i])) { tl_vocab_error = "orbx_vocab_create_from_nodes: parent id not yet defined"; return nullptr; }
    return upload(T, k, L, scoring, weighting, device);
}

orbx_vocab *orbx_vocab_load_text(const char *path, int device) {
    FILE *f = path ? fopen(path, "r") : nullptr;
    if (!f) { tl_vocab_error = "orbx_vocab_load_text: cannot open file"; return nullptr; }
    std::vector<char> line(1 << 12);
    auto readline = [&]() -> bool {          // whole line, whatever its length
        size_t len = 0;
        for (;;) {
            if (!fgets(line.data() + len, (int)(line.size() - len), f)) return len > 0;
            len += strlen(line.data() + len);
            if (len > 0 && line[len - 1] == '\n') return true;
            if (len + 1 >= line.size()) line.resize(line.size() * 2); else return true;   // EOF without newline
        }
    };
    int k = 0, L = 0, n1 = 0, n2 = 0;
    if (!readline() || sscanf(line.data(), "%d %d %d %d", &k, &L, &n1, &n2) != 4 || k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) {
        fclose(f);
        tl_vocab_error = "orbx_vocab_load_text: not a vocabulary text file";    // same acceptance test as TemplatedVocabulary.h:1359
        return nullptr;
    }
    HostTree T;
    while (readline()) {
        char *p = line.data();
        while (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n') ++p;
        if (!*p) continue;                     // blank line (the reference reads a phantom node from it)
        char *end = nullptr;
        const long pid = strtol(p, &end, 10); p = end;
        const long leaf = strtol(p, &end, 10); p = end;
        uint8_t d[32];
        for (int i = 0; i < 32; ++i) { d[i] = (uint8_t)strtol(p, &end, 10); p = end; }
        const double w = strtod(p, &end);
        if (!T.add((int)pid, (int)leaf, d, w)) { fclose(f); tl_vocab_error = "orbx_vocab_load_text: parent id not yet defined"; return nullptr; }
    }
    fclose(f);
    return upload(T, k, L, n1, n2, device);
}

void orbx_vocab_destroy(orbx_vocab *v) {
    if (!v) return;
    OrbxDeviceGuard dg_(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    void *bufs[] = {v->d_desc, v->d_childOff, v->d_child, v->d_wordId, v->d_weight, v->w_desc, v->w_word, v->w_node, v->w_bowIds, v->w_fvNodes,
                    v->w_fvIdx, v->w_leaf, v->w_fvOff, v->w_counts, v->w_bowVals, v->b_leaf};
    for (void *p : bufs) if (p) cudaFree(p);
    delete v;
}

const char *orbx_vocab_last_error(const orbx_vocab *v) { return v ? v->err.c_str() : tl_vocab_error.c_str(); }

int orbx_vocab_info(const orbx_vocab *v, int *k, int *L, int *n_nodes, int *n_words) {
    if (!v) return ORBX_ERR_ARG;
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (n_nodes) *n_nodes = v->nNodes;
    if (n_words) *n_words = v->nWords;
    return ORBX_OK;
}

void *orbx_vocab_stream(orbx_vocab *v) { return v ? (void *)v->stream : nullptr; }
int orbx_vocab_sync(orbx_vocab *v) {
    if (!v) return ORBX_ERR_ARG;
    VCUDA_TRY(v, cudaStreamSynchronize(v->stream));
    return ORBX_OK;
}

int orbx_bow_transform_batch_device(orbx_vocab *v, const uint8_t *d_desc, size_t desc_stride_bytes, const int32_t *d_n, int batch, int cap,
                                    int levelsup, uint32_t *d_word_id, uint32_t *d_node_id, uint32_t *d_bow_ids, double *d_bow_vals,
                                    int32_t *d_n_bow, uint32_t *d_fv_nodes, int32_t *d_fv_off, uint32_t *d_fv_idx, int32_t *d_n_fv) {
    if (!v) return ORBX_ERR_ARG;
    if (batch < 0 || cap < 0 || (batch > 0 && cap > 0 && (!d_desc || !d_n || !d_word_id || !d_node_id || !d_bow_ids || !d_bow_vals || !d_n_bow ||
                                                           !d_fv_nodes || !d_fv_off || !d_fv_idx || !d_n_fv))) {
        v->err = "orbx_bow_transform_batch_device: null argument";
        return ORBX_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(v->mu);
    OrbxDeviceGuard dg_(v->device); VCUDA_TRY(v, dg_.status);
    const size_t need = (size_t)batch * cap;
    if (need > v->b_leafCap) {
        VCUDA_TRY(v, cudaStreamSynchronize(v->stream));
        if (v->b_leaf) cudaFree(v->b_leaf);
        v->b_leaf = nullptr; v->b_leafCap = 0;
        VCUDA_TRY(v, cudaMalloc((void **)&v->b_leaf, need * sizeof(int)));
        v->b_leafCap = need;
    }
    return launch_transform(v, d_desc, desc_stride_bytes, d_n, batch, cap, levelsup, d_word_id, d_node_id, v->b_leaf, d_bow_ids, d_bow_vals, d_n_bow,
                            d_fv_nodes, d_fv_off, d_fv_idx, d_n_fv);
}

int orbx_bow_transform(orbx_vocab *v, const uint8_t *desc, int n, int levelsup, uint32_t *word_id, uint32_t *node_id, uint32_t *bow_ids,
                       double *bow_vals, int32_t *n_bow, uint32_t *fv_nodes, int32_t *fv_off, uint32_t *fv_idx, int32_t *n_fv) {
    if (!v) return ORBX_ERR_ARG;
    if (n < 0 || !n_bow || !n_fv || !fv_off || (n > 0 && (!desc || !bow_ids || !bow_vals || !fv_nodes || !fv_idx))) {
        v->err = "orbx_bow_transform: null argument";
        return ORBX_ERR_ARG;
    }
    *n_bow = 0; *n_fv = 0; fv_off[0] = 0;
    if (n == 0) return ORBX_OK;
    if (n > 16384) { v->err = "orbx_bow: more than 16384 descriptors per image"; return ORBX_ERR_CAPACITY; }
    std::lock_guard<std::mutex> lk(v->mu);
    OrbxDeviceGuard dg_(v->device); VCUDA_TRY(v, dg_.status);
    if (n > v->workCap) {
        VCUDA_TRY(v, cudaStreamSynchronize(v->stream));
        void **bufs[] = {(void **)&v->w_desc, (void **)&v->w_word, (void **)&v->w_node, (void **)&v->w_bowIds, (void **)&v->w_fvNodes,
                         (void **)&v->w_fvIdx, (void **)&v->w_leaf, (void **)&v->w_fvOff, (void **)&v->w_counts, (void **)&v->w_bowVals};
        for (void **p : bufs) { if (*p) cudaFree(*p); *p = nullptr; }
        v->workCap = 0;
        const size_t c = (size_t)n;
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_desc, c * 32));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_word, c * 4));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_node, c * 4));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_bowIds, c * 4));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_fvNodes, c * 4));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_fvIdx, c * 4));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_leaf, c * 4));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_fvOff, (c + 1) * 4));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_counts, 4 * sizeof(int)));
        VCUDA_TRY(v, cudaMalloc((void **)&v->w_bowVals, c * 8));
        v->workCap = n;
    }
    cudaStream_t s = v->stream;
    const int cap = n;                       // one image: rows are sized by this call, not by the largest one seen
    int counts[3] = {n, 0, 0};
    VCUDA_TRY(v, cudaMemcpyAsync(v->w_counts, counts, sizeof(counts), cudaMemcpyHostToDevice, s));
    VCUDA_TRY(v, cudaMemcpyAsync(v->w_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    int rc = launch_transform(v, v->w_desc, 0, v->w_counts, 1, cap, levelsup, v->w_word, v->w_node, v->w_leaf, v->w_bowIds, v->w_bowVals,
                              v->w_counts + 1, v->w_fvNodes, v->w_fvOff, v->w_fvIdx, v->w_counts + 2);
    if (rc) return rc;
    VCUDA_TRY(v, cudaMemcpyAsync(counts, v->w_counts, sizeof(counts), cudaMemcpyDeviceToHost, s));
    VCUDA_TRY(v, cudaStreamSynchronize(s));
    const int nb = counts[1], nf = counts[2];
    if (word_id) VCUDA_TRY(v, cudaMemcpyAsync(word_id, v->w_word, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (node_id) VCUDA_TRY(v, cudaMemcpyAsync(node_id, v->w_node, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    if (nb > 0) {
        VCUDA_TRY(v, cudaMemcpyAsync(bow_ids, v->w_bowIds, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
        VCUDA_TRY(v, cudaMemcpyAsync(bow_vals, v->w_bowVals, (size_t)nb * 8, cudaMemcpyDeviceToHost, s));
    }
    VCUDA_TRY(v, cudaMemcpyAsync(fv_off, v->w_fvOff, (size_t)(nf + 1) * 4, cudaMemcpyDeviceToHost, s));
    if (nf > 0) VCUDA_TRY(v, cudaMemcpyAsync(fv_nodes, v->w_fvNodes, (size_t)nf * 4, cudaMemcpyDeviceToHost, s));
    VCUDA_TRY(v, cudaStreamSynchronize(s));
    const int total = fv_off[nf];
    if (total > 0) {
        VCUDA_TRY(v, cudaMemcpyAsync(fv_idx, v->w_fvIdx, (size_t)total * 4, cudaMemcpyDeviceToHost, s));
        VCUDA_TRY(v, cudaStreamSynchronize(s));
    }
    *n_bow = nb; *n_fv = nf;
    return ORBX_OK;
}

}  // extern "C"
