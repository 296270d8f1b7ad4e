// orbx_match.cu — B200 (sm_100a) Hamming matching inner loops of the reference's ORBmatcher and of
// Frame::ComputeStereoFishEyeMatches (file:line relative to /root/reference):
//   DescriptorDistance            src/ORBmatcher.cc:2054-2070   256-bit XOR + popcount
//   BFMatcher.knnMatch(k=2)       src/Frame.cc:1078             brute-force top-2, ties → lower index
//   Lowe ratio                    src/Frame.cc:1085             (float)d0 < (float)d1 * 0.7 (double)
//   best/second loops             src/ORBmatcher.cc:84-140 etc. explicit candidate lists
//   rotation histogram filter     src/ORBmatcher.cc:345-352, :2008-2049
//
// The brute-force kernel is POPC-pipe bound, not HBM bound (the DB streams once per 1024 queries):
// each thread keeps R query descriptors in registers, the block stages DB rows in shared memory and
// every thread reads them as broadcast uint4; the running top-2 is two packed (dist<<23 | row) keys so
// the update is branch-free min/max and the tie rule (lower train index first) falls out of the key
// order.  Per-chunk partial top-2s are merged by one warp per query with a shuffle reduction.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "orbx_internal.h"

namespace {

#define KNN_THREADS 256
#define KNN_DBT 128          // DB rows staged per shared-memory tile
#define KNN_IDX_BITS 23      // rows per chunk < 2^23 so that dist (≤256, 9 bits) fits above
#define KNN_KEY_NONE 0xffffffffu

__device__ __forceinline__ int ham256(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// 256-bit Hamming distance with a carry-save adder front end: three full adders (2 LOP3 each) fold seven of the
// eight XOR words into two "ones" words and three "twos" words, so a pair costs 5 POPC instead of 8 — the POPC
// pipe (16 lanes/clk/SM) is the binding unit of the brute-force kernel, LOP3 runs on the wider ALU pipe.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int ham256_csa(const uint4 &a0, const uint4 &a1, const uint4 &b0, const uint4 &b1) {
    const uint32_t x0 = a0.x ^ b0.x, x1 = a0.y ^ b0.y, x2 = a0.z ^ b0.z, x3 = a0.w ^ b0.w;
    const uint32_t x4 = a1.x ^ b1.x, x5 = a1.y ^ b1.y, x6 = a1.z ^ b1.z, x7 = a1.w ^ b1.w;
    // each full adder is exactly two LOP3: sum = xor3 (0x96), carry = majority (0xE8)
    const uint32_t s1 = lop3_xor3(x0, x1, x2), c1 = lop3_maj(x0, x1, x2);
    const uint32_t s2 = lop3_xor3(x3, x4, x5), c2 = lop3_maj(x3, x4, x5);
    const uint32_t s3 = lop3_xor3(s1, s2, x6), c3 = lop3_maj(s1, s2, x6);
    return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}

// grid: (nChunks, nQueryTiles).  Thread t of query tile y owns queries y*THREADS*R + r*THREADS + t.
template <int R>
__global__ void __launch_bounds__(KNN_THREADS) k_knn2_partial(const uint4 *__restrict__ q, int nq,
                                                              const uint4 *__restrict__ db, long long ndb,
                                                              long long chunkRows, uint2 *__restrict__ partial) {
    __shared__ uint4 tile[KNN_DBT * 2];
    const int tid = threadIdx.x;
    const long long c0 = (long long)blockIdx.x * chunkRows;
    const long long c1 = min(c0 + chunkRows, ndb);
    const int qbase = blockIdx.y * (KNN_THREADS * R);
    uint4 qa[R], qb[R];
    uint32_t k0[R], k1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int qi = qbase + r * KNN_THREADS + tid;
        if (qi < nq) { qa[r] = q[2 * (long long)qi]; qb[r] = q[2 * (long long)qi + 1]; }
        else { qa[r] = make_uint4(0, 0, 0, 0); qb[r] = qa[r]; }
        k0[r] = KNN_KEY_NONE; k1[r] = KNN_KEY_NONE;
    }
    for (long long t0 = c0; t0 < c1; t0 += KNN_DBT) {
        const int rows = (int)min((long long)KNN_DBT, c1 - t0);
        __syncthreads();
        for (int i = tid; i < rows * 2; i += KNN_THREADS) tile[i] = db[2 * t0 + i];
        __syncthreads();
        uint32_t jkey = (uint32_t)(t0 - c0);
#pragma unroll 4
        for (int j = 0; j < rows; ++j, ++jkey) {
            const uint4 d0 = tile[2 * j], d1 = tile[2 * j + 1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t key = ((uint32_t)ham256_csa(qa[r], qb[r], d0, d1) << KNN_IDX_BITS) + jkey;
                const uint32_t hi = max(key, k0[r]);
                k0[r] = min(key, k0[r]);
                k1[r] = min(k1[r], hi);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int qi = qbase + r * KNN_THREADS + tid;
        if (qi < nq) partial[(long long)blockIdx.x * nq + qi] = make_uint2(k0[r], k1[r]);
    }
}

__device__ __forceinline__ void top2_insert(unsigned long long k, unsigned long long &a, unsigned long long &b) {
    const unsigned long long hi = k > a ? k : a;
    a = k < a ? k : a;
    b = b < hi ? b : hi;
}

// one warp per query: lanes stride over the chunks, shuffle-reduce the two smallest (dist, idx) keys
__global__ void __launch_bounds__(256) k_knn2_merge_partials(const uint2 *__restrict__ partial, int nq, int nChunks,
                                                             long long chunkRows, long long idxBase,
                                                             int32_t *__restrict__ idx, int32_t *__restrict__ dist) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nq) return;
    const unsigned long long NONE = ~0ull;
    unsigned long long a = NONE, b = NONE;
    for (int c = lane; c < nChunks; c += 32) {
        const uint2 p = partial[(long long)c * nq + warp];
        const unsigned long long base = (unsigned long long)(idxBase + (long long)c * chunkRows);
        if (p.x != KNN_KEY_NONE) top2_insert(((unsigned long long)(p.x >> KNN_IDX_BITS) << 32) | (base + (p.x & ((1u << KNN_IDX_BITS) - 1))), a, b);
        if (p.y != KNN_KEY_NONE) top2_insert(((unsigned long long)(p.y >> KNN_IDX_BITS) << 32) | (base + (p.y & ((1u << KNN_IDX_BITS) - 1))), a, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
        top2_insert(oa, a, b);
        top2_insert(ob, a, b);
    }
    if (lane == 0) {
        idx[2 * warp] = a == NONE ? -1 : (int32_t)(a & 0xffffffffu);
        dist[2 * warp] = a == NONE ? INT_MAX : (int32_t)(a >> 32);
        idx[2 * warp + 1] = b == NONE ? -1 : (int32_t)(b & 0xffffffffu);
        dist[2 * warp + 1] = b == NONE ? INT_MAX : (int32_t)(b >> 32);
    }
}

// merge of per-shard (idx, dist) top-2 lists, e.g. after an all-gather across GPUs
// (shard g's lists start shardStride int32 after shard g-1's: 2·nq for two separate arrays, 4·nq for the packed
// {idx[nq×2], dist[nq×2]} records of orbx_knn2_merge_packed_device)
__global__ void __launch_bounds__(256) k_knn2_merge_shards(const int32_t *__restrict__ idxAll, const int32_t *__restrict__ distAll,
                                                           int nShards, int nq, long long shardStride, int32_t *__restrict__ idx,
                                                           int32_t *__restrict__ dist) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nq) return;
    const unsigned long long NONE = ~0ull;
    unsigned long long a = NONE, b = NONE;
    for (int e = lane; e < 2 * nShards; e += 32) {
        const long long o = (long long)(e >> 1) * shardStride + (long long)warp * 2 + (e & 1);
        const int32_t i = idxAll[o], d = distAll[o];
        if (i >= 0) top2_insert(((unsigned long long)(uint32_t)d << 32) | (uint32_t)i, a, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
        top2_insert(oa, a, b);
        top2_insert(ob, a, b);
    }
    if (lane == 0) {
        idx[2 * warp] = a == NONE ? -1 : (int32_t)(a & 0xffffffffu);
        dist[2 * warp] = a == NONE ? INT_MAX : (int32_t)(a >> 32);
        idx[2 * warp + 1] = b == NONE ? -1 : (int32_t)(b & 0xffffffffu);
        dist[2 * warp + 1] = b == NONE ? INT_MAX : (int32_t)(b >> 32);
    }
}

__global__ void k_ratio(const int32_t *dist, int nq, double ratio, uint8_t *keep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int d0 = dist[2 * i], d1 = dist[2 * i + 1];
    keep[i] = (d1 != INT_MAX) && ((double)(float)d0 < __dmul_rn((double)(float)d1, ratio));
}

// best / second-best over explicit candidate lists (the a12 loops, src/ORBmatcher.cc:84-121): one warp per query, lanes stride
// over the list; the two smallest (distance, list position) keys reproduce the scalar loop's strict '<' (first candidate wins
// ties).  The loop starts from bestDist = bestDist2 = 256, so a distance of 256 is never recorded.  secondIdx is the candidate
// whose octave the reference keeps as bestLevel2 (:110, :117).
__device__ __forceinline__ void warp_top2(unsigned long long &a, unsigned long long &b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
        top2_insert(oa, a, b);
        top2_insert(ob, a, b);
    }
}
__global__ void __launch_bounds__(256) k_top2_lists(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ db,
                                                    const int32_t *__restrict__ cand, const int32_t *__restrict__ off, int32_t *bestIdx,
                                                    int32_t *bestDist, int32_t *secondIdx, int32_t *secondDist) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= nq) return;
    const uint4 a0 = q[2 * (long long)i], a1 = q[2 * (long long)i + 1];
    const unsigned long long NONE = ~0ull;
    unsigned long long a = NONE, b = NONE;
    const int c0 = off[i], c1 = off[i + 1];
    for (int c = c0 + lane; c < c1; c += 32) {
        const int t = cand[c];
        const int d = ham256(a0, a1, db[2 * (long long)t], db[2 * (long long)t + 1]);
        if (d < 256) top2_insert(((unsigned long long)(unsigned)d << 32) | (unsigned)(c - c0), a, b);
    }
    warp_top2(a, b);
    if (lane == 0) {
        bestIdx[i] = a == NONE ? -1 : cand[c0 + (int)(a & 0xffffffffu)];
        bestDist[i] = a == NONE ? 256 : (int)(a >> 32);
        secondIdx[i] = b == NONE ? -1 : cand[c0 + (int)(b & 0xffffffffu)];
        secondDist[i] = b == NONE ? 256 : (int)(b >> 32);
    }
}

// rotation histogram + three maxima (one block)
__global__ void __launch_bounds__(256) k_rot_hist(const float *a, const float *b, int n, uint8_t *keep) {
    __shared__ int hist[30];
    __shared__ int sel[3];
    const int tid = threadIdx.x;
    if (tid < 30) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        float rot = __fsub_rn(a[i], b[i]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));  // quirk Q10: degrees × 1/30, half away from zero
        if (bin == 30) bin = 0;
        bin = min(max(bin, 0), 29);
        keep[i] = (uint8_t)bin;
        atomicAdd(&hist[bin], 1);
    }
    __syncthreads();
    if (tid == 0) {  // ComputeThreeMaxima (:2008-2049)
        int m1 = 0, m2 = 0, m3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = hist[i];
            if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
            else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
            else if (s > m3) { m3 = s; i3 = i; }
        }
        if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
        else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
        sel[0] = i1; sel[1] = i2; sel[2] = i3;
    }
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const int bin = keep[i];
        keep[i] = (bin == sel[0] || bin == sel[1] || bin == sel[2]) ? 1 : 0;
    }
}

// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:644-759) with the candidate lists given explicitly.
// Queries must be replayed in order (earlier matches lock train keypoints through vMatchedDistance and can be
// stolen later), so ONE warp walks the queries sequentially; the 32 lanes share the candidate
// scan of the current query (distance + lock test per candidate) and shuffle-reduce the two smallest
// (distance, list position) keys, which reproduces the strict-'<' first-wins order of the scalar loop.
__global__ void __launch_bounds__(32) k_search_init(const uint4 *__restrict__ d1, const float *__restrict__ ang1,
                                                    const int32_t *__restrict__ oct1, int n1, const uint4 *__restrict__ d2,
                                                    const float *__restrict__ ang2, int n2, const int32_t *__restrict__ cand,
                                                    const int32_t *__restrict__ off, float nnratio, int checkOri,
                                                    int32_t *m12, int32_t *m21, int32_t *matchedDist, int8_t *binOf,
                                                    int32_t *nMatchesOut) {
    const int lane = threadIdx.x;
    __shared__ int hist[30];
    for (int i = lane; i < 30; i += 32) hist[i] = 0;
    for (int i = lane; i < n1; i += 32) { m12[i] = -1; binOf[i] = -1; }
    for (int i = lane; i < n2; i += 32) { m21[i] = -1; matchedDist[i] = INT_MAX; }
    __syncwarp();
    int nmatches = 0;
    const unsigned long long NONE = ~0ull;
    for (int i1 = 0; i1 < n1; ++i1) {
        if (oct1[i1] > 0) continue;                       // level1 > 0 (:662-664)
        const int c0 = off[i1], c1 = off[i1 + 1];
        if (c0 == c1) continue;
        const uint4 qa = d1[2 * (long long)i1], qb = d1[2 * (long long)i1 + 1];
        unsigned long long a = NONE, b = NONE;
        for (int c = c0 + lane; c < c1; c += 32) {
            const int i2 = cand[c];
            const int dist = ham256(qa, qb, d2[2 * (long long)i2], d2[2 * (long long)i2 + 1]);
            if (matchedDist[i2] <= dist) continue;        // :685-686
            top2_insert(((unsigned long long)(unsigned)dist << 32) | (unsigned)(c - c0), a, b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long oa = __shfl_xor_sync(0xffffffffu, a, o), ob = __shfl_xor_sync(0xffffffffu, b, o);
            top2_insert(oa, a, b);
            top2_insert(ob, a, b);
        }
        if (a != NONE) {
            const int best = (int)(a >> 32);
            const int best2 = b == NONE ? INT_MAX : (int)(b >> 32);
            const int bestIdx2 = cand[c0 + (int)(a & 0xffffffffu)];
            if (best <= 50 && (float)best < __fmul_rn((float)best2, nnratio)) {   // TH_LOW, fp32 ratio (:698-700)
                const int prev = m21[bestIdx2];
                if (prev >= 0) --nmatches;
                ++nmatches;
                int bin = -1;
                if (checkOri) {
                    float rot = __fsub_rn(ang1[i1], ang2[bestIdx2]);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
                    if (bin == 30) bin = 0;
                    bin = min(max(bin, 0), 29);
                }
                __syncwarp();   // every lane has read m21/ang before lane 0 updates the bookkeeping
                if (lane == 0) {
                    if (prev >= 0) m12[prev] = -1;
                    m12[i1] = bestIdx2;
                    m21[bestIdx2] = i1;
                    matchedDist[bestIdx2] = best;
                    if (checkOri) { binOf[i1] = (int8_t)bin; ++hist[bin]; }
                }
            }
        }
        __syncwarp();
    }
    if (checkOri) {
        __shared__ int sel[3];
        if (lane == 0) {  // ComputeThreeMaxima (:2008-2049); stale (stolen) entries count, as in the reference
            int mx1 = 0, mx2 = 0, mx3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < 30; ++i) {
                const int sz = hist[i];
                if (sz > mx1) { mx3 = mx2; mx2 = mx1; mx1 = sz; i3 = i2; i2 = i1; i1 = i; }
                else if (sz > mx2) { mx3 = mx2; mx2 = sz; i3 = i2; i2 = i; }
                else if (sz > mx3) { mx3 = sz; i3 = i; }
            }
            if ((float)mx2 < __fmul_rn(0.1f, (float)mx1)) { i2 = -1; i3 = -1; }
            else if ((float)mx3 < __fmul_rn(0.1f, (float)mx1)) { i3 = -1; }
            sel[0] = i1; sel[1] = i2; sel[2] = i3;
        }
        __syncwarp();
        int dropped = 0;
        for (int i = lane; i < n1; i += 32) {
            const int bin = binOf[i];
            if (bin >= 0 && bin != sel[0] && bin != sel[1] && bin != sel[2] && m12[i] >= 0) { m12[i] = -1; ++dropped; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dropped += __shfl_xor_sync(0xffffffffu, dropped, o);
        nmatches -= dropped;
    }
    if (lane == 0) *nMatchesOut = nmatches;
}

// ---- Frame grid: AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:387-418, :659-738) ----
#define GRID_COLS 64
#define GRID_ROWS 48
#define GRID_CELLS (GRID_COLS * GRID_ROWS)

__device__ __forceinline__ int grid_cell_of(float x, float y, float minX, float minY, float wInv, float hInv) {
    const int px = (int)roundf(__fmul_rn(__fsub_rn(x, minX), wInv)), py = (int)roundf(__fmul_rn(__fsub_rn(y, minY), hInv));
    if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) return -1;
    return px * GRID_ROWS + py;
}
__global__ void k_grid_count(const float2 *xy, int n, float minX, float minY, float wInv, float hInv, int *cellCnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = grid_cell_of(xy[i].x, xy[i].y, minX, minY, wInv, hInv);
    if (c >= 0) atomicAdd(&cellCnt[c], 1);
}
// single block: exclusive scan of a[0..n) into out[0..n], out[n] = total
__global__ void __launch_bounds__(1024) k_scan_excl(const int *a, int n, int *out) {
    __shared__ int warpSum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? a[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) warpSum[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warpSum[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (threadIdx.x >= o) wi += t;
            }
            warpSum[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const int excl = carry + warpSum[threadIdx.x >> 5] + incl - v;
        if (i < n) out[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}
__global__ void k_grid_fill(const float2 *xy, int n, float minX, float minY, float wInv, float hInv, const int *cellOff, int *cellFill, int *items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = grid_cell_of(xy[i].x, xy[i].y, minX, minY, wInv, hInv);
    if (c >= 0) items[cellOff[c] + atomicAdd(&cellFill[c], 1)] = i;
}
// push_back order inside a cell is ascending keypoint index: sort each (short) cell list
__global__ void k_grid_sort_cells(const int *cellOff, int *items) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= GRID_CELLS) return;
    const int lo = cellOff[c], hi = cellOff[c + 1];
    for (int i = lo + 1; i < hi; ++i) {
        const int v = items[i];
        int j = i - 1;
        while (j >= lo && items[j] > v) { items[j + 1] = items[j]; --j; }
        items[j + 1] = v;
    }
}
// one thread per query; FILL=false counts, FILL=true writes the list (reference order: ix, iy, insertion)
template <bool FILL>
__global__ void k_area_query(const float2 *xy, const int32_t *octave, const int *cellOff, const int *items, float minX, float minY,
                             float wInv, float hInv, const float *queries, int nq, int minLevel, int maxLevel, int *cnt,
                             const int *outOff, int32_t *out, const int32_t *qLevels = nullptr, const uint8_t *qActive = nullptr) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    if (qActive && !qActive[q]) { if (!FILL) cnt[q] = 0; return; }       // callers that skip queries (level1 > 0, points not in view …)
    if (qLevels) { minLevel = qLevels[2 * q]; maxLevel = qLevels[2 * q + 1]; }
    const float x = queries[3 * q], y = queries[3 * q + 1], r = queries[3 * q + 2];
    const int x0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), wInv)));
    const int x1 = min(GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), wInv)));
    const int y0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), hInv)));
    const int y1 = min(GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), hInv)));
    int n = 0;
    int at = FILL ? outOff[q] : 0;
    if (!(x0 >= GRID_COLS || x1 < 0 || y0 >= GRID_ROWS || y1 < 0)) {
        const bool checkLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = x0; ix <= x1; ++ix)
            for (int iy = y0; iy <= y1; ++iy) {
                const int c = ix * GRID_ROWS + iy;
                for (int k = cellOff[c]; k < cellOff[c + 1]; ++k) {
                    const int j = items[k];
                    if (checkLevels) {
                        const int o = octave[j];
                        if (o < minLevel) continue;
                        if (maxLevel >= 0 && o > maxLevel) continue;
                    }
                    const float2 pj = xy[j];
                    if (fabsf(__fsub_rn(pj.x, x)) < r && fabsf(__fsub_rn(pj.y, y)) < r) {
                        if (FILL) out[at++] = j;
                        ++n;
                    }
                }
            }
    }
    if (!FILL) cnt[q] = n;
}

// ---- ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, …) (src/ORBmatcher.cc:43-213), frames with Nleft == -1 ----
// Stage 1 (parallel, one warp per map point): Hamming distance of the map point's descriptor to every keypoint of its
// GetFeaturesInArea list; candidates that fail the right-image check of :92-97 (a per-candidate, order-independent test)
// get the marker 0xffff.
__global__ void __launch_bounds__(256) k_sbp_dist(const uint4 *__restrict__ mpDesc, int m, const uint4 *__restrict__ desc,
                                                  const int32_t *__restrict__ cand, const int32_t *__restrict__ off,
                                                  const float *__restrict__ uRight, const float *__restrict__ projXR,
                                                  const float *__restrict__ radius, uint16_t *__restrict__ distOut) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (j >= m) return;
    const int c0 = off[j], c1 = off[j + 1];
    if (c0 == c1) return;
    const uint4 a0 = mpDesc[2 * (long long)j], a1 = mpDesc[2 * (long long)j + 1];
    const float xr = uRight ? projXR[j] : 0.f, rad = radius[j];
    for (int c = c0 + lane; c < c1; c += 32) {
        const int idx = cand[c];
        int d = ham256(a0, a1, desc[2 * (long long)idx], desc[2 * (long long)idx + 1]);
        if (uRight) {
            const float ur = uRight[idx];
            if (ur > 0.f && fabsf(__fsub_rn(xr, ur)) > rad) d = 0xffff;
        }
        distOut[c] = (uint16_t)d;
    }
}
// Stage 2 (ordered replay, ONE warp): map points are visited in order because an accepted match attaches the map point to its
// keypoint and later map points skip keypoints whose map point has observations (:88-90).  Per map point the lanes scan the
// stored distances, drop occupied keypoints, and shuffle-reduce the two smallest (distance, list position) keys = the scalar
// loop's best / second best; then the level-aware ratio rule of :123-128.  obs[] = Observations() of the map point attached to
// each keypoint (-1: none), updated as matches are accepted.
__global__ void __launch_bounds__(32) k_sbp_replay(int m, const uint8_t *__restrict__ active, const int32_t *__restrict__ mpObs,
                                                   const int32_t *__restrict__ cand, const int32_t *__restrict__ off,
                                                   const uint16_t *__restrict__ dist, const int32_t *__restrict__ octave, int n,
                                                   float nnratio, int32_t *obs, int32_t *assigned, int32_t *nMatchesOut) {
    const int lane = threadIdx.x;
    for (int i = lane; i < n; i += 32) assigned[i] = -1;
    __syncwarp();
    const unsigned long long NONE = ~0ull;
    int nmatches = 0;
    for (int j = 0; j < m; ++j) {
        if (!active[j]) continue;
        const int c0 = off[j], c1 = off[j + 1];
        if (c0 == c1) continue;
        unsigned long long a = NONE, b = NONE;
        for (int c = c0 + lane; c < c1; c += 32) {
            const int d = dist[c];
            if (d >= 256) continue;                               // right-image check failed, or 256 (never below the initial 256)
            if (obs[cand[c]] > 0) continue;
            top2_insert(((unsigned long long)(unsigned)d << 32) | (unsigned)(c - c0), a, b);
        }
        warp_top2(a, b);
        if (a == NONE) continue;
        const int bestDist = (int)(a >> 32), bestIdx = cand[c0 + (int)(a & 0xffffffffu)];
        if (bestDist > 100) continue;                             // TH_HIGH (:124)
        const int bestLevel = octave[bestIdx];
        const int bestDist2 = b == NONE ? 256 : (int)(b >> 32);
        const int bestLevel2 = b == NONE ? -1 : octave[cand[c0 + (int)(b & 0xffffffffu)]];
        const float lim = __fmul_rn(nnratio, (float)bestDist2);
        if (bestLevel == bestLevel2 && (float)bestDist > lim) continue;
        if (bestLevel != bestLevel2 || (float)bestDist <= lim) {
            ++nmatches;
            __syncwarp();
            if (lane == 0) { obs[bestIdx] = mpObs[j]; assigned[bestIdx] = j; }
            __syncwarp();
        }
    }
    if (lane == 0) *nMatchesOut = nmatches;
}
// ---- ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (src/ORBmatcher.cc:222-425), frames with Nleft == -1 ----
// The host has already merged the two feature vectors: query j is keyframe feature qKF[j] (in the walk's order: common nodes
// ascending, then the node's stored order; features without a good map point left out) and its candidates are fIdx[qC0[j] .. qC1[j])
// — the frame's features of the same vocabulary node.  ONE warp replays the queries in order because a frame feature that received
// a map point is skipped by every later query (:281-282); per query the lanes share the candidate scan (occupancy test, Hamming
// distance) and shuffle-reduce the two smallest (distance, list position) keys = the scalar loop's strict-'<' best / second best.
// Then TH_LOW and the fp32 ratio test (:331-333), the rotation histogram (:346-353) and, at the end, the purge of the matches
// outside the three fullest bins (:404-422).  SearchByBoW(KeyFrame*, KeyFrame*, …) (:760-901) is the same walk with the strict threshold
// (thLow = 49) and candidate lists already reduced to the second keyframe's features with a good map point.
__global__ void __launch_bounds__(32) k_search_by_bow(const uint4 *__restrict__ kfDesc, const float *__restrict__ kfAng,
                                                      const uint4 *__restrict__ fDesc, const float *__restrict__ fAng, int nF,
                                                      const int32_t *__restrict__ qKF, const int32_t *__restrict__ qC0,
                                                      const int32_t *__restrict__ qC1, int nQ, const int32_t *__restrict__ fIdx,
                                                      float nnratio, int thLow, int checkOri, int32_t *assigned, int8_t *binOf, int32_t *nMatchesOut) {
    const int lane = threadIdx.x;
    __shared__ int hist[30];
    __shared__ int sel[3];
    for (int i = lane; i < 30; i += 32) hist[i] = 0;
    for (int i = lane; i < nF; i += 32) { assigned[i] = -1; binOf[i] = -1; }
    __syncwarp();
    const unsigned long long NONE = ~0ull;
    int nmatches = 0;
    for (int j = 0; j < nQ; ++j) {
        const int kf = qKF[j], c0 = qC0[j], c1 = qC1[j];
        const uint4 qa = kfDesc[2 * (long long)kf], qb = kfDesc[2 * (long long)kf + 1];
        unsigned long long a = NONE, b = NONE;
        for (int c = c0 + lane; c < c1; c += 32) {
            const int t = fIdx[c];
            if (assigned[t] >= 0) continue;                                   // :281-282
            const int d = ham256(qa, qb, fDesc[2 * (long long)t], fDesc[2 * (long long)t + 1]);
            if (d < 256) top2_insert(((unsigned long long)(unsigned)d << 32) | (unsigned)(c - c0), a, b);
        }
        warp_top2(a, b);
        if (a == NONE) continue;
        const int best = (int)(a >> 32), best2 = b == NONE ? 256 : (int)(b >> 32);
        if (best <= thLow && (float)best < __fmul_rn(nnratio, (float)best2)) {   // TH_LOW (<= 50 for a frame, < 50 between keyframes), fp32 ratio
            const int bestIdx = fIdx[c0 + (int)(a & 0xffffffffu)];
            __syncwarp();
            if (lane == 0) {
                assigned[bestIdx] = kf;
                if (checkOri) {
                    float rot = __fsub_rn(kfAng[kf], fAng[bestIdx]);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
                    if (bin == 30) bin = 0;
                    bin = min(max(bin, 0), 29);
                    binOf[bestIdx] = (int8_t)bin;
                    ++hist[bin];
                }
            }
            __syncwarp();
            ++nmatches;
        }
    }
    if (checkOri) {
        __syncwarp();
        if (lane == 0) {  // ComputeThreeMaxima (:2008-2049)
            int m1 = 0, m2 = 0, m3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < 30; ++i) {
                const int sz = hist[i];
                if (sz > m1) { m3 = m2; m2 = m1; m1 = sz; i3 = i2; i2 = i1; i1 = i; }
                else if (sz > m2) { m3 = m2; m2 = sz; i3 = i2; i2 = i; }
                else if (sz > m3) { m3 = sz; i3 = i; }
            }
            if ((float)m2 < __fmul_rn(0.1f, (float)m1)) { i2 = -1; i3 = -1; }
            else if ((float)m3 < __fmul_rn(0.1f, (float)m1)) { i3 = -1; }
            sel[0] = i1; sel[1] = i2; sel[2] = i3;
        }
        __syncwarp();
        int removed = 0;
        for (int i = lane; i < nF; i += 32) {
            if (assigned[i] < 0) continue;
            const int bin = binOf[i];
            if (bin != sel[0] && bin != sel[1] && bin != sel[2]) { assigned[i] = -1; ++removed; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nmatches -= removed;
    }
    if (lane == 0) *nMatchesOut = nmatches;
}
// vbPrevMatched update of SearchForInitialization (src/ORBmatcher.cc:754-756)
__global__ void k_update_prev(const int32_t *m12, int n1, const float2 *xy2, float2 *prev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n1 && m12[i] >= 0) prev[i] = xy2[m12[i]];
}

// ---- stereo association tail (src/Frame.cc:862-914) over the kNN + ratio matches; one block ----
__global__ void __launch_bounds__(256) k_stereo_tail(const float *uL, const float *uR, int nL, int nR, const int32_t *idx,
                                                     const int32_t *dist, const uint8_t *keep, float mbf, float mb, float *uRight,
                                                     float *depth, int *keptOut) {
    __shared__ int hist[257];
    __shared__ float thDist;
    __shared__ int total, dropped;
    const int tid = threadIdx.x;
    for (int i = tid; i < 257; i += 256) hist[i] = 0;
    if (tid == 0) { total = 0; dropped = 0; }
    __syncthreads();
    const float maxD = __fdiv_rn(mbf, mb);
    for (int i = tid; i < nL; i += 256) {
        float ur = -1.f, dp = -1.f;
        if (keep[i]) {
            const int iR = idx[2 * i];
            if (iR >= 0 && iR < nR) {
                float disparity = __fsub_rn(uL[i], uR[iR]);
                if (disparity >= 0.f && disparity < maxD) {
                    if (disparity <= 0.f) disparity = 0.01f;
                    dp = __fdiv_rn(mbf, disparity);
                    ur = uR[iR];
                    atomicAdd(&hist[min(max(dist[2 * i], 0), 256)], 1);
                    atomicAdd(&total, 1);
                }
            }
        }
        uRight[i] = ur; depth[i] = dp;
    }
    __syncthreads();
    if (tid == 0) {  // median of the sorted distances = element [size/2]
        float th = 3.4e38f;
        if (total > 0) {
            const int k = total / 2;
            int acc = 0, med = 0;
            for (int d = 0; d <= 256; ++d) { acc += hist[d]; if (acc > k) { med = d; break; } }
            th = __fmul_rn(1.5f, (float)med);
        }
        thDist = th;
    }
    __syncthreads();
    for (int i = tid; i < nL; i += 256) {
        if (uRight[i] != -1.f || depth[i] != -1.f) {
            if (!((float)dist[2 * i] < thDist)) { uRight[i] = -1.f; depth[i] = -1.f; atomicAdd(&dropped, 1); }
        }
    }
    __syncthreads();
    if (tid == 0) *keptOut = total - dropped;
}

thread_local std::string tl_merr;

}  // namespace

// Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:749-811): cv::undistortPoints(pts, K, D, R = I, P = K).
// One thread per point; double precision, every operation individually rounded (the library is built -fmad=false),
// five fixed-point iterations like OpenCV's default criteria.  xy points at the first float of the first (x, y) pair;
// consecutive points are strideFloats apart (2 for packed pairs, 7 for cv::KeyPoint records).
struct UndistortK { double fx, fy, cx, cy, k[14]; };
__global__ void k_undistort(const float *__restrict__ xy, int strideIn, int n, UndistortK K, float *__restrict__ out, int strideOut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *k = K.k;
    const double ifx = 1.0 / K.fx, ify = 1.0 / K.fy;
    double x = ((double)xy[(size_t)i * strideIn] - K.cx) * ifx, y = ((double)xy[(size_t)i * strideIn + 1] - K.cy) * ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double dY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    const double xx = K.fx * x + 0.0 * y + K.cx, yy = 0.0 * x + K.fy * y + K.cy, ww = 1.0 / (0.0 * x + 0.0 * y + 1.0);
    out[(size_t)i * strideOut] = (float)(xx * ww);
    out[(size_t)i * strideOut + 1] = (float)(yy * ww);
}

struct orbx_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint2 *d_partial = nullptr; size_t partialCap = 0;
    uint8_t *d_buf = nullptr; size_t bufCap = 0;  // staging for the host-buffer entry points
    uint8_t *d_buf2 = nullptr; size_t buf2Cap = 0; // second arena: candidate lists whose size is only known after a count pass
    int nSM = 148;
    long long tcLaunches = 0;      // tensor-core kNN launches since creation (test / bench evidence)
    bool useTc = true;             // tensor-core kNN (orbx_knn_tc.cu) for large problems; ORBX_KNN_POPC keeps the POPC kernel (same results)
};

cudaError_t orbx_knn2_tc_launch(const uint8_t *d_q, int nq, const uint8_t *d_db, long long ndb, long long chunkRows, int nChunks, uint2 *d_partial,
                                cudaStream_t stream);

namespace {
#define MCUDA_TRY(m, call)                                                                \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            (m)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            return ORBX_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

int stage(orbx_matcher *m, size_t bytes) {
    if (bytes <= m->bufCap && m->d_buf) return ORBX_OK;
    if (m->d_buf) cudaFree(m->d_buf);
    m->d_buf = nullptr; m->bufCap = 0;
    MCUDA_TRY(m, cudaMalloc((void **)&m->d_buf, bytes));
    m->bufCap = bytes;
    return ORBX_OK;
}
int stage2(orbx_matcher *m, size_t bytes) {
    if (bytes <= m->buf2Cap && m->d_buf2) return ORBX_OK;
    if (m->d_buf2) cudaFree(m->d_buf2);
    m->d_buf2 = nullptr; m->buf2Cap = 0;
    bytes += bytes / 2;                                 // list sizes vary from call to call: grow with headroom
    MCUDA_TRY(m, cudaMalloc((void **)&m->d_buf2, bytes));
    m->buf2Cap = bytes;
    return ORBX_OK;
}
inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

// Bump allocator over the staging arena (all blocks 256-byte aligned).
struct Carver {
    uint8_t *p;
    template <typename T> T *take(size_t count) { T *r = reinterpret_cast<T *>(p); p += al256(count * sizeof(T)); return r; }
};

// Frame::AssignFeaturesToGrid on the device (src/Frame.cc:387-418): cellOff[GRID_CELLS + 1] + items[n] in push_back order.
// dCnt / dOff / dFill are GRID_CELLS + 1 ints each and contiguous.
int grid_build(orbx_matcher *m, const float2 *dxy, int n, float minX, float minY, float wInv, float hInv, int *dCnt, int *dOff, int *dFill,
               int *dItems) {
    cudaStream_t s = m->stream;
    MCUDA_TRY(m, cudaMemsetAsync(dCnt, 0, 3 * al256((GRID_CELLS + 1) * 4), s));
    if (n > 0) k_grid_count<<<(n + 255) / 256, 256, 0, s>>>(dxy, n, minX, minY, wInv, hInv, dCnt);
    k_scan_excl<<<1, 1024, 0, s>>>(dCnt, GRID_CELLS, dOff);
    if (n > 0) {
        k_grid_fill<<<(n + 255) / 256, 256, 0, s>>>(dxy, n, minX, minY, wInv, hInv, dOff, dFill, dItems);
        k_grid_sort_cells<<<(GRID_CELLS + 255) / 256, 256, 0, s>>>(dOff, dItems);
    }
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}
}  // namespace

extern "C" {

orbx_matcher *orbx_matcher_create(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
        cudaGetLastError();
        tl_merr = "orbx_matcher_create: no such CUDA device (liborbx has no CPU fallback)";
        return nullptr;
    }
    orbx_matcher *m = new orbx_matcher;
    m->device = device;
    OrbxDeviceGuard dg_(device);
    if (dg_.status != cudaSuccess || cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) {
        tl_merr = "orbx_matcher_create: cannot create a stream";
        delete m;
        return nullptr;
    }
    cudaDeviceGetAttribute(&m->nSM, cudaDevAttrMultiProcessorCount, device);
    m->useTc = getenv("ORBX_KNN_POPC") == nullptr;
    return m;
}

void orbx_matcher_destroy(orbx_matcher *m) {
    if (!m) return;
    OrbxDeviceGuard dg_(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->d_partial) cudaFree(m->d_partial);
    if (m->d_buf) cudaFree(m->d_buf);
    if (m->d_buf2) cudaFree(m->d_buf2);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

long long orbx_debug_knn_tc_launches(const orbx_matcher *m) { return m ? m->tcLaunches : 0; }
const char *orbx_matcher_last_error(const orbx_matcher *m) { return m ? m->err.c_str() : tl_merr.c_str(); }
void *orbx_matcher_stream(orbx_matcher *m) { return m ? (void *)m->stream : nullptr; }
int orbx_matcher_sync(orbx_matcher *m) {
    if (!m) return ORBX_ERR_ARG;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    MCUDA_TRY(m, cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_hamming_knn2_device(orbx_matcher *m, const uint8_t *d_q, int nq, const uint8_t *d_db, int64_t ndb,
                             int64_t idx_base, int32_t *d_idx, int32_t *d_dist) {
    if (!m) return ORBX_ERR_ARG;
    if (nq < 0 || ndb < 0 || (nq > 0 && (!d_q || !d_idx || !d_dist)) || (ndb > 0 && !d_db) ||
        idx_base < 0 || idx_base + ndb > (int64_t)INT_MAX) {
        m->err = "orbx_hamming_knn2_device: bad argument";
        return ORBX_ERR_ARG;
    }
    if (nq == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    if (m->useTc && nq >= 64 && ndb >= 8192) {
        // tensor-core path: CTAs of 128 queries × one DB chunk; about four waves of CTAs, chunk length a multiple of the 256-row MMA tile
        const int qTilesTc = (nq + 127) / 128;
        long long nCh = std::max(1LL, (long long)m->nSM * 4 / qTilesTc);
        long long chunkRowsTc = (ndb + nCh - 1) / nCh;
        chunkRowsTc = std::max(256LL, (chunkRowsTc + 255) / 256 * 256);
        chunkRowsTc = std::min<long long>(chunkRowsTc, ((1LL << 22) - 256) / 256 * 256);      // the kernel's keys hold 22 row bits
        const int nChunksTc = (int)((ndb + chunkRowsTc - 1) / chunkRowsTc);
        const size_t needTc = (size_t)nChunksTc * nq;
        if (needTc > m->partialCap || !m->d_partial) {
            if (m->d_partial) { MCUDA_TRY(m, cudaStreamSynchronize(m->stream)); cudaFree(m->d_partial); }
            m->d_partial = nullptr; m->partialCap = 0;
            MCUDA_TRY(m, cudaMalloc((void **)&m->d_partial, needTc * sizeof(uint2)));
            m->partialCap = needTc;
        }
        MCUDA_TRY(m, orbx_knn2_tc_launch(d_q, nq, d_db, ndb, chunkRowsTc, nChunksTc, m->d_partial, m->stream));
        ++m->tcLaunches;
        k_knn2_merge_partials<<<(nq * 32 + 255) / 256, 256, 0, m->stream>>>(m->d_partial, nq, nChunksTc, chunkRowsTc, idx_base, d_idx, d_dist);
        MCUDA_TRY(m, cudaGetLastError());
        return ORBX_OK;
    }
    // queries per block tile: R·256; pick R so small problems still spread over the SMs
    const int R = nq > 2 * KNN_THREADS ? 4 : (nq > KNN_THREADS ? 2 : 1);
    const int qTiles = (nq + KNN_THREADS * R - 1) / (KNN_THREADS * R);
    // DB chunks: about 4 resident blocks per SM overall, chunk length a multiple of the smem tile
    long long targetChunks = std::max(1LL, (long long)m->nSM * 4 / qTiles);
    long long chunkRows = std::max<long long>(KNN_DBT, (ndb + targetChunks - 1) / targetChunks);
    chunkRows = (chunkRows + KNN_DBT - 1) / KNN_DBT * KNN_DBT;
    chunkRows = std::min<long long>(chunkRows, (1LL << KNN_IDX_BITS) - KNN_DBT);
    const int nChunks = (int)std::max<long long>(1, (ndb + chunkRows - 1) / chunkRows);
    const size_t need = (size_t)nChunks * nq;
    if (need > m->partialCap || !m->d_partial) {
        if (m->d_partial) { MCUDA_TRY(m, cudaStreamSynchronize(m->stream)); cudaFree(m->d_partial); }
        m->d_partial = nullptr; m->partialCap = 0;
        MCUDA_TRY(m, cudaMalloc((void **)&m->d_partial, need * sizeof(uint2)));
        m->partialCap = need;
    }
    dim3 grd(nChunks, qTiles);
    const uint4 *q4 = reinterpret_cast<const uint4 *>(d_q), *db4 = reinterpret_cast<const uint4 *>(d_db);
    if (R == 4) k_knn2_partial<4><<<grd, KNN_THREADS, 0, m->stream>>>(q4, nq, db4, ndb, chunkRows, m->d_partial);
    else if (R == 2) k_knn2_partial<2><<<grd, KNN_THREADS, 0, m->stream>>>(q4, nq, db4, ndb, chunkRows, m->d_partial);
    else k_knn2_partial<1><<<grd, KNN_THREADS, 0, m->stream>>>(q4, nq, db4, ndb, chunkRows, m->d_partial);
    k_knn2_merge_partials<<<(nq * 32 + 255) / 256, 256, 0, m->stream>>>(m->d_partial, nq, nChunks, chunkRows, idx_base, d_idx, d_dist);
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}

int orbx_knn2_merge_device(orbx_matcher *m, const int32_t *d_idx_all, const int32_t *d_dist_all, int n_shards, int nq,
                           int32_t *d_idx, int32_t *d_dist) {
    if (!m || n_shards < 1 || nq < 0 || (nq > 0 && (!d_idx_all || !d_dist_all || !d_idx || !d_dist))) {
        if (m) m->err = "orbx_knn2_merge_device: bad argument";
        return ORBX_ERR_ARG;
    }
    if (nq == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    k_knn2_merge_shards<<<(nq * 32 + 255) / 256, 256, 0, m->stream>>>(d_idx_all, d_dist_all, n_shards, nq, 2LL * nq, d_idx, d_dist);
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}

int orbx_knn2_merge_packed_device(orbx_matcher *m, const int32_t *d_packed_all, int n_shards, int nq, int32_t *d_idx, int32_t *d_dist) {
    if (!m || n_shards < 1 || nq < 0 || (nq > 0 && (!d_packed_all || !d_idx || !d_dist))) {
        if (m) m->err = "orbx_knn2_merge_packed_device: bad argument";
        return ORBX_ERR_ARG;
    }
    if (nq == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    k_knn2_merge_shards<<<(nq * 32 + 255) / 256, 256, 0, m->stream>>>(d_packed_all, d_packed_all + 2LL * nq, n_shards, nq, 4LL * nq, d_idx, d_dist);
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}

int orbx_hamming_knn2(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int64_t ndb, int32_t *idx,
                      int32_t *dist) {
    if (!m) return ORBX_ERR_ARG;
    if (nq < 0 || ndb < 0 || (nq > 0 && (!query || !idx || !dist)) || (ndb > 0 && !train)) {
        m->err = "orbx_hamming_knn2: bad argument";
        return ORBX_ERR_ARG;
    }
    if (nq == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const size_t qB = al256((size_t)nq * 32), dB = al256((size_t)ndb * 32 + 32), oB = al256((size_t)nq * 8);
    int rc = stage(m, qB + dB + 2 * oB);
    if (rc) return rc;
    uint8_t *dq = m->d_buf, *dd = dq + qB;
    int32_t *di = reinterpret_cast<int32_t *>(dd + dB), *ddist = reinterpret_cast<int32_t *>(dd + dB + oB);
    MCUDA_TRY(m, cudaMemcpyAsync(dq, query, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    if (ndb > 0) MCUDA_TRY(m, cudaMemcpyAsync(dd, train, (size_t)ndb * 32, cudaMemcpyHostToDevice, m->stream));
    rc = orbx_hamming_knn2_device(m, dq, nq, dd, ndb, 0, di, ddist);
    if (rc) return rc;
    MCUDA_TRY(m, cudaMemcpyAsync(idx, di, (size_t)nq * 8, cudaMemcpyDeviceToHost, m->stream));
    MCUDA_TRY(m, cudaMemcpyAsync(dist, ddist, (size_t)nq * 8, cudaMemcpyDeviceToHost, m->stream));
    MCUDA_TRY(m, cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_ratio_test_device(orbx_matcher *m, const int32_t *d_dist, int nq, double ratio, uint8_t *d_keep) {
    if (!m) return ORBX_ERR_ARG;
    if (nq <= 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    k_ratio<<<(nq + 255) / 256, 256, 0, m->stream>>>(d_dist, nq, ratio, d_keep);
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}

int orbx_ratio_test(orbx_matcher *m, const int32_t *dist, int nq, double ratio, uint8_t *keep) {
    if (!m) return ORBX_ERR_ARG;
    if (nq < 0 || (nq > 0 && (!dist || !keep))) { m->err = "orbx_ratio_test: bad argument"; return ORBX_ERR_ARG; }
    if (nq == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const size_t dB = al256((size_t)nq * 8);
    int rc = stage(m, dB + al256(nq));
    if (rc) return rc;
    int32_t *dd = (int32_t *)m->d_buf;
    uint8_t *dk = m->d_buf + dB;
    MCUDA_TRY(m, cudaMemcpyAsync(dd, dist, (size_t)nq * 8, cudaMemcpyHostToDevice, m->stream));
    rc = orbx_ratio_test_device(m, dd, nq, ratio, dk);
    if (rc) return rc;
    MCUDA_TRY(m, cudaMemcpyAsync(keep, dk, nq, cudaMemcpyDeviceToHost, m->stream));
    MCUDA_TRY(m, cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_hamming_top2_lists(orbx_matcher *m, const uint8_t *query, int nq, const uint8_t *train, int64_t ndb,
                            const int32_t *cand, const int32_t *cand_off, int32_t *best_idx, int32_t *best_dist,
                            int32_t *second_idx, int32_t *second_dist) {
    if (!m) return ORBX_ERR_ARG;
    if (nq < 0 || ndb < 0 || (nq > 0 && (!query || !cand_off || !best_idx || !best_dist || !second_dist))) {
        m->err = "orbx_hamming_top2_lists: bad argument";
        return ORBX_ERR_ARG;
    }
    if (nq == 0) return ORBX_OK;
    const int nc = cand_off[nq];
    if (nc > 0 && !cand) { m->err = "orbx_hamming_top2_lists: bad argument"; return ORBX_ERR_ARG; }
    for (int i = 0; i < nc; ++i)
        if (cand[i] < 0 || cand[i] >= ndb) { m->err = "orbx_hamming_top2_lists: candidate index out of range"; return ORBX_ERR_ARG; }
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const size_t qB = al256((size_t)nq * 32), dB = al256((size_t)ndb * 32 + 32), cB = al256((size_t)std::max(nc, 1) * 4),
                 oB = al256((size_t)(nq + 1) * 4);
    int rc = stage(m, qB + dB + cB + 5 * oB);
    if (rc) return rc;
    Carver cv{m->d_buf};
    uint8_t *dq = cv.take<uint8_t>((size_t)nq * 32), *dd = cv.take<uint8_t>((size_t)ndb * 32 + 32);
    int32_t *dc = cv.take<int32_t>(std::max(nc, 1)), *doff = cv.take<int32_t>(nq + 1);
    int32_t *dbi = cv.take<int32_t>(nq + 1), *dbd = cv.take<int32_t>(nq + 1), *dsi = cv.take<int32_t>(nq + 1), *dsd = cv.take<int32_t>(nq + 1);
    cudaStream_t s = m->stream;
    MCUDA_TRY(m, cudaMemcpyAsync(dq, query, (size_t)nq * 32, cudaMemcpyHostToDevice, s));
    if (ndb > 0) MCUDA_TRY(m, cudaMemcpyAsync(dd, train, (size_t)ndb * 32, cudaMemcpyHostToDevice, s));
    if (nc > 0) MCUDA_TRY(m, cudaMemcpyAsync(dc, cand, (size_t)nc * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(doff, cand_off, (size_t)(nq + 1) * 4, cudaMemcpyHostToDevice, s));
    k_top2_lists<<<(nq * 32 + 255) / 256, 256, 0, s>>>((const uint4 *)dq, nq, (const uint4 *)dd, dc, doff, dbi, dbd, dsi, dsd);
    MCUDA_TRY(m, cudaGetLastError());
    MCUDA_TRY(m, cudaMemcpyAsync(best_idx, dbi, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(best_dist, dbd, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    if (second_idx) MCUDA_TRY(m, cudaMemcpyAsync(second_idx, dsi, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(second_dist, dsd, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaStreamSynchronize(s));
    return ORBX_OK;
}

int orbx_search_for_initialization(orbx_matcher *m, const uint8_t *desc1, const float *angle1, const int32_t *octave1, int n1,
                                   const uint8_t *desc2, const float *angle2, int n2, const int32_t *cand,
                                   const int32_t *cand_off, float nnratio, int check_orientation, int32_t *matches12,
                                   int32_t *n_matches) {
    if (!m) return ORBX_ERR_ARG;
    if (n1 < 0 || n2 < 0 || (n1 > 0 && (!desc1 || !angle1 || !octave1 || !cand_off || !matches12)) || (n2 > 0 && (!desc2 || !angle2)) || !n_matches) {
        m->err = "orbx_search_for_initialization: bad argument";
        return ORBX_ERR_ARG;
    }
    *n_matches = 0;
    if (n1 == 0) return ORBX_OK;
    const int nc = cand_off[n1];
    for (int i = 0; i < nc; ++i)
        if (cand[i] < 0 || cand[i] >= n2) { m->err = "orbx_search_for_initialization: candidate index out of range"; return ORBX_ERR_ARG; }
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const size_t d1B = al256((size_t)n1 * 32), d2B = al256((size_t)n2 * 32 + 32), a1B = al256((size_t)n1 * 4), a2B = al256((size_t)n2 * 4 + 4),
                 cB = al256((size_t)std::max(nc, 1) * 4), oB = al256((size_t)(n1 + 1) * 4);
    int rc = stage(m, d1B + d2B + 2 * a1B + a2B + cB + oB + a1B /*m12*/ + 2 * a2B /*m21, matchedDist*/ + al256(n1) + 256);
    if (rc) return rc;
    uint8_t *p = m->d_buf;
    uint8_t *dd1 = p; p += d1B;
    uint8_t *dd2 = p; p += d2B;
    float *da1 = (float *)p; p += a1B;
    int32_t *do1 = (int32_t *)p; p += a1B;
    float *da2 = (float *)p; p += a2B;
    int32_t *dc = (int32_t *)p; p += cB;
    int32_t *doff = (int32_t *)p; p += oB;
    int32_t *dm12 = (int32_t *)p; p += a1B;
    int32_t *dm21 = (int32_t *)p; p += a2B;
    int32_t *dmd = (int32_t *)p; p += a2B;
    int8_t *dbin = (int8_t *)p; p += al256(n1);
    int32_t *dn = (int32_t *)p;
    cudaStream_t s = m->stream;
    MCUDA_TRY(m, cudaMemcpyAsync(dd1, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    if (n2 > 0) MCUDA_TRY(m, cudaMemcpyAsync(dd2, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(da1, angle1, (size_t)n1 * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(do1, octave1, (size_t)n1 * 4, cudaMemcpyHostToDevice, s));
    if (n2 > 0) MCUDA_TRY(m, cudaMemcpyAsync(da2, angle2, (size_t)n2 * 4, cudaMemcpyHostToDevice, s));
    if (nc > 0) MCUDA_TRY(m, cudaMemcpyAsync(dc, cand, (size_t)nc * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(doff, cand_off, (size_t)(n1 + 1) * 4, cudaMemcpyHostToDevice, s));
    k_search_init<<<1, 32, 0, s>>>((const uint4 *)dd1, da1, do1, n1, (const uint4 *)dd2, da2, n2, dc, doff, nnratio, check_orientation,
                                   dm12, dm21, dmd, dbin, dn);
    MCUDA_TRY(m, cudaGetLastError());
    MCUDA_TRY(m, cudaMemcpyAsync(matches12, dm12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(n_matches, dn, 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaStreamSynchronize(s));
    return ORBX_OK;
}

int orbx_features_in_area(orbx_matcher *m, const float *keypoints_xy, const int32_t *octave, int n, float min_x, float min_y,
                          float max_x, float max_y, const float *queries_xyr, int nq, int min_level, int max_level,
                          int32_t *cand_off, int32_t *cand, int cap, int32_t *total_out) {
    if (!m) return ORBX_ERR_ARG;
    if (n < 0 || nq < 0 || (n > 0 && (!keypoints_xy || !octave)) || (nq > 0 && !queries_xyr) || !cand_off || !total_out || !(max_x > min_x) || !(max_y > min_y)) {
        m->err = "orbx_features_in_area: bad argument";
        return ORBX_ERR_ARG;
    }
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const float wInv = (float)GRID_COLS / (float)(max_x - min_x), hInv = (float)GRID_ROWS / (float)(max_y - min_y);
    const size_t xyB = al256((size_t)std::max(n, 1) * 8), ocB = al256((size_t)std::max(n, 1) * 4), cellB = al256((GRID_CELLS + 1) * 4),
                 itB = al256((size_t)std::max(n, 1) * 4), qB = al256((size_t)std::max(nq, 1) * 12), qcB = al256((size_t)(nq + 1) * 4);
    int rc = stage(m, xyB + ocB + 3 * cellB + itB + qB + 2 * qcB);
    if (rc) return rc;
    uint8_t *p = m->d_buf;
    float2 *dxy = (float2 *)p; p += xyB;
    int32_t *doc = (int32_t *)p; p += ocB;
    int *dCnt = (int *)p; p += cellB;
    int *dOff = (int *)p; p += cellB;
    int *dFill = (int *)p; p += cellB;
    int *dItems = (int *)p; p += itB;
    float *dq = (float *)p; p += qB;
    int *dqCnt = (int *)p; p += qcB;
    int *dqOff = (int *)p; p += qcB;
    cudaStream_t s = m->stream;
    if (n > 0) {
        MCUDA_TRY(m, cudaMemcpyAsync(dxy, keypoints_xy, (size_t)n * 8, cudaMemcpyHostToDevice, s));
        MCUDA_TRY(m, cudaMemcpyAsync(doc, octave, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    }
    if (nq > 0) MCUDA_TRY(m, cudaMemcpyAsync(dq, queries_xyr, (size_t)nq * 12, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemsetAsync(dCnt, 0, 2 * cellB + cellB, s));   // dCnt, dOff, dFill are contiguous
    if (n > 0) k_grid_count<<<(n + 255) / 256, 256, 0, s>>>(dxy, n, min_x, min_y, wInv, hInv, dCnt);
    k_scan_excl<<<1, 1024, 0, s>>>(dCnt, GRID_CELLS, dOff);
    if (n > 0) {
        k_grid_fill<<<(n + 255) / 256, 256, 0, s>>>(dxy, n, min_x, min_y, wInv, hInv, dOff, dFill, dItems);
        k_grid_sort_cells<<<(GRID_CELLS + 255) / 256, 256, 0, s>>>(dOff, dItems);
    }
    int total = 0;
    if (nq > 0) {
        k_area_query<false><<<(nq + 127) / 128, 128, 0, s>>>(dxy, doc, dOff, dItems, min_x, min_y, wInv, hInv, dq, nq, min_level, max_level, dqCnt, nullptr, nullptr);
        k_scan_excl<<<1, 1024, 0, s>>>(dqCnt, nq, dqOff);
        MCUDA_TRY(m, cudaMemcpyAsync(cand_off, dqOff, (size_t)(nq + 1) * 4, cudaMemcpyDeviceToHost, s));
        MCUDA_TRY(m, cudaStreamSynchronize(s));
        total = cand_off[nq];
    } else {
        cand_off[0] = 0;
    }
    *total_out = total;
    if (total > 0 && cand && cap >= total) {
        // the candidate array lives after the staged inputs: grow the staging buffer if needed (contents are kept
        // by re-running the cheap build when the buffer moves)
        int32_t *dOut = nullptr;
        MCUDA_TRY(m, cudaMalloc((void **)&dOut, (size_t)total * 4));
        k_area_query<true><<<(nq + 127) / 128, 128, 0, s>>>(dxy, doc, dOff, dItems, min_x, min_y, wInv, hInv, dq, nq, min_level, max_level, nullptr, dqOff, dOut);
        cudaError_t e = cudaMemcpyAsync(cand, dOut, (size_t)total * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(dOut);
        if (e != cudaSuccess) { m->err = std::string("orbx_features_in_area: ") + cudaGetErrorString(e); return ORBX_ERR_CUDA; }
    } else if (total > 0 && cand) {
        m->err = "orbx_features_in_area: candidate capacity too small (see total_out)";
        return ORBX_ERR_CAPACITY;
    }
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}

int orbx_stereo_tail(orbx_matcher *m, const float *u_left, const float *u_right, int n_left, int n_right, const int32_t *idx,
                     const int32_t *dist, const uint8_t *keep, float mbf, float mb, float *mvu_right, float *mv_depth, int32_t *n_kept) {
    if (!m) return ORBX_ERR_ARG;
    if (n_left < 0 || n_right < 0 || (n_left > 0 && (!u_left || !idx || !dist || !keep || !mvu_right || !mv_depth)) || (n_right > 0 && !u_right) || !n_kept) {
        m->err = "orbx_stereo_tail: bad argument";
        return ORBX_ERR_ARG;
    }
    *n_kept = 0;
    if (n_left == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const size_t lB = al256((size_t)n_left * 4), rB = al256((size_t)std::max(n_right, 1) * 4), iB = al256((size_t)n_left * 8), kB = al256(n_left);
    int rc = stage(m, 3 * lB + rB + 2 * iB + kB + 256);
    if (rc) return rc;
    uint8_t *p = m->d_buf;
    float *duL = (float *)p; p += lB;
    float *duR = (float *)p; p += rB;
    int32_t *di = (int32_t *)p; p += iB;
    int32_t *dd = (int32_t *)p; p += iB;
    uint8_t *dk = p; p += kB;
    float *dur = (float *)p; p += lB;
    float *ddp = (float *)p; p += lB;
    int *dn = (int *)p;
    cudaStream_t s = m->stream;
    MCUDA_TRY(m, cudaMemcpyAsync(duL, u_left, (size_t)n_left * 4, cudaMemcpyHostToDevice, s));
    if (n_right > 0) MCUDA_TRY(m, cudaMemcpyAsync(duR, u_right, (size_t)n_right * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(di, idx, (size_t)n_left * 8, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dd, dist, (size_t)n_left * 8, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dk, keep, (size_t)n_left, cudaMemcpyHostToDevice, s));
    k_stereo_tail<<<1, 256, 0, s>>>(duL, duR, n_left, n_right, di, dd, dk, mbf, mb, dur, ddp, dn);
    MCUDA_TRY(m, cudaGetLastError());
    MCUDA_TRY(m, cudaMemcpyAsync(mvu_right, dur, (size_t)n_left * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(mv_depth, ddp, (size_t)n_left * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(n_kept, dn, 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaStreamSynchronize(s));
    return ORBX_OK;
}

static UndistortK make_undistort(float fx, float fy, float cx, float cy, const float *D, int nd) {
    UndistortK K;
    K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy;
    for (int i = 0; i < 14; ++i) K.k[i] = i < nd ? (double)D[i] : 0.0;
    return K;
}

int orbx_undistort_points_device(orbx_matcher *m, const float *d_xy, int stride_in, int n, float fx, float fy, float cx, float cy,
                                 const float *dist_coef, int n_coef, float *d_out, int stride_out) {
    if (!m) return ORBX_ERR_ARG;
    if (n < 0 || stride_in < 2 || stride_out < 2 || n_coef < 0 || n_coef > 14 || (n_coef > 0 && !dist_coef) || (n > 0 && (!d_xy || !d_out))) {
        m->err = "orbx_undistort_points_device: bad argument";
        return ORBX_ERR_ARG;
    }
    if (n == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    if (n_coef == 0 || dist_coef[0] == 0.0f) {          // mvKeysUn = mvKeys (src/Frame.cc:751-755)
        if (d_out != d_xy || stride_in != stride_out)
            MCUDA_TRY(m, cudaMemcpy2DAsync(d_out, (size_t)stride_out * 4, d_xy, (size_t)stride_in * 4, 8, n, cudaMemcpyDeviceToDevice, m->stream));
        return ORBX_OK;
    }
    k_undistort<<<(n + 127) / 128, 128, 0, m->stream>>>(d_xy, stride_in, n, make_undistort(fx, fy, cx, cy, dist_coef, n_coef), d_out, stride_out);
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}

int orbx_undistort_keypoints(orbx_matcher *m, const orbx_keypoint *kps, int n, float fx, float fy, float cx, float cy, const float *dist_coef,
                             int n_coef, orbx_keypoint *kps_un) {
    if (!m) return ORBX_ERR_ARG;
    if (n < 0 || n_coef < 0 || n_coef > 14 || (n_coef > 0 && !dist_coef) || (n > 0 && (!kps || !kps_un))) { m->err = "orbx_undistort_keypoints: bad argument"; return ORBX_ERR_ARG; }
    if (n == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const size_t bytes = (size_t)n * sizeof(orbx_keypoint);
    int rc = stage(m, al256(bytes));
    if (rc) return rc;
    float *d = (float *)m->d_buf;
    MCUDA_TRY(m, cudaMemcpyAsync(d, kps, bytes, cudaMemcpyHostToDevice, m->stream));
    rc = orbx_undistort_points_device(m, d, 7, n, fx, fy, cx, cy, dist_coef, n_coef, d, 7);      // in place: only pt changes
    if (rc) return rc;
    MCUDA_TRY(m, cudaMemcpyAsync(kps_un, d, bytes, cudaMemcpyDeviceToHost, m->stream));
    MCUDA_TRY(m, cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_image_bounds(orbx_matcher *m, int cols, int rows, float fx, float fy, float cx, float cy, const float *dist_coef, int n_coef,
                      float *bounds4) {
    if (!m) return ORBX_ERR_ARG;
    if (!bounds4 || n_coef < 0 || n_coef > 14 || (n_coef > 0 && !dist_coef)) { m->err = "orbx_image_bounds: bad argument"; return ORBX_ERR_ARG; }
    if (n_coef == 0 || dist_coef[0] == 0.0f) {          // src/Frame.cc:804-810
        bounds4[0] = 0.f; bounds4[1] = (float)cols; bounds4[2] = 0.f; bounds4[3] = (float)rows;
        return ORBX_OK;
    }
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    int rc = stage(m, 256);
    if (rc) return rc;
    const float c[8] = {0.f, 0.f, (float)cols, 0.f, 0.f, (float)rows, (float)cols, (float)rows};
    float u[8];
    float *d = (float *)m->d_buf;
    MCUDA_TRY(m, cudaMemcpyAsync(d, c, sizeof(c), cudaMemcpyHostToDevice, m->stream));
    rc = orbx_undistort_points_device(m, d, 2, 4, fx, fy, cx, cy, dist_coef, n_coef, d + 8, 2);
    if (rc) return rc;
    MCUDA_TRY(m, cudaMemcpyAsync(u, d + 8, sizeof(u), cudaMemcpyDeviceToHost, m->stream));
    MCUDA_TRY(m, cudaStreamSynchronize(m->stream));
    bounds4[0] = u[0] < u[4] ? u[0] : u[4];
    bounds4[1] = u[2] > u[6] ? u[2] : u[6];
    bounds4[2] = u[1] < u[3] ? u[1] : u[3];
    bounds4[3] = u[5] > u[7] ? u[5] : u[7];
    return ORBX_OK;
}

int orbx_rot_hist_filter_device(orbx_matcher *m, const float *d_a, const float *d_b, int n, uint8_t *d_keep) {
    if (!m) return ORBX_ERR_ARG;
    if (n <= 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    k_rot_hist<<<1, 256, 0, m->stream>>>(d_a, d_b, n, d_keep);
    MCUDA_TRY(m, cudaGetLastError());
    return ORBX_OK;
}

int orbx_rot_hist_filter(orbx_matcher *m, const float *angle_a, const float *angle_b, int n, uint8_t *keep) {
    if (!m) return ORBX_ERR_ARG;
    if (n < 0 || (n > 0 && (!angle_a || !angle_b || !keep))) { m->err = "orbx_rot_hist_filter: bad argument"; return ORBX_ERR_ARG; }
    if (n == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const size_t aB = al256((size_t)n * 4);
    int rc = stage(m, 2 * aB + al256(n));
    if (rc) return rc;
    float *da = (float *)m->d_buf, *db = (float *)(m->d_buf + aB);
    uint8_t *dk = m->d_buf + 2 * aB;
    MCUDA_TRY(m, cudaMemcpyAsync(da, angle_a, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    MCUDA_TRY(m, cudaMemcpyAsync(db, angle_b, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    rc = orbx_rot_hist_filter_device(m, da, db, n, dk);
    if (rc) return rc;
    MCUDA_TRY(m, cudaMemcpyAsync(keep, dk, n, cudaMemcpyDeviceToHost, m->stream));
    MCUDA_TRY(m, cudaStreamSynchronize(m->stream));
    return ORBX_OK;
}

int orbx_search_by_projection(orbx_matcher *m, const orbx_keypoint *keypoints_un, const uint8_t *descriptors, int n, const float *u_right,
                              const int32_t *kp_obs, const float *bounds4, const float *scale_factors, int n_levels, const float *mp_proj5,
                              const int32_t *mp_level, const uint8_t *mp_flags, const int32_t *mp_obs, const uint8_t *mp_desc, int n_mp,
                              float nnratio, float th, int far_points, float th_far, int32_t *assigned, int32_t *n_matches) {
    if (!m) return ORBX_ERR_ARG;
    if (n < 0 || n_mp < 0 || n_levels < 1 || !bounds4 || !scale_factors || !n_matches || (n > 0 && (!keypoints_un || !descriptors || !assigned)) ||
        (n_mp > 0 && (!mp_proj5 || !mp_level || !mp_flags || !mp_obs || !mp_desc)) || !(bounds4[2] > bounds4[0]) || !(bounds4[3] > bounds4[1])) {
        m->err = "orbx_search_by_projection: bad argument";
        return ORBX_ERR_ARG;
    }
    *n_matches = 0;
    for (int i = 0; i < n; ++i) assigned[i] = -1;
    if (n == 0 || n_mp == 0) return ORBX_OK;
    // per map point, on the host (a handful of scalar operations each): the skip tests of :52-59, the search radius of
    // :64-69 / :215-221 and the level band nPredictedLevel-1 … nPredictedLevel of :72
    std::vector<float> q((size_t)n_mp * 3), rad(n_mp), xr(n_mp);
    std::vector<int32_t> lv((size_t)n_mp * 2);
    std::vector<uint8_t> act(n_mp);
    const bool bFactor = th != 1.0;
    for (int j = 0; j < n_mp; ++j) {
        const int L = mp_level[j];
        bool a = (mp_flags[j] & 1) && !(mp_flags[j] & 2) && !(far_points && mp_proj5[5 * j + 4] > th_far);
        if (a && (L < 0 || L >= n_levels)) { m->err = "orbx_search_by_projection: predicted level out of range"; return ORBX_ERR_ARG; }
        float r = (mp_proj5[5 * j + 3] > 0.998) ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        const float rs = a ? r * scale_factors[L] : 0.f;
        q[3 * j] = mp_proj5[5 * j]; q[3 * j + 1] = mp_proj5[5 * j + 1]; q[3 * j + 2] = rs;
        rad[j] = rs; xr[j] = mp_proj5[5 * j + 2];
        lv[2 * j] = L - 1; lv[2 * j + 1] = L;
        act[j] = a;
    }
    std::vector<float> xy((size_t)n * 2);
    std::vector<int32_t> oc(n), obs(n, -1);
    for (int i = 0; i < n; ++i) { xy[2 * i] = keypoints_un[i].x; xy[2 * i + 1] = keypoints_un[i].y; oc[i] = keypoints_un[i].octave; }
    if (kp_obs) obs.assign(kp_obs, kp_obs + n);
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const float minX = bounds4[0], minY = bounds4[1];
    const float wInv = (float)GRID_COLS / (float)(bounds4[2] - bounds4[0]), hInv = (float)GRID_ROWS / (float)(bounds4[3] - bounds4[1]);
    const size_t cellB = al256((GRID_CELLS + 1) * 4);
    int rc = stage(m, al256((size_t)n * 8) + 5 * al256((size_t)n * 4) + al256((size_t)n * 32 + 32) + 3 * cellB + al256((size_t)n_mp * 12) +
                          al256((size_t)n_mp * 8) + 3 * al256((size_t)n_mp * 4) + al256(n_mp) + al256((size_t)n_mp * 32 + 32) +
                          2 * al256((size_t)(n_mp + 1) * 4) + 256);
    if (rc) return rc;
    Carver cv{m->d_buf};
    float2 *dxy = cv.take<float2>(n);
    int32_t *doc = cv.take<int32_t>(n), *dobs = cv.take<int32_t>(n), *dasg = cv.take<int32_t>(n);
    float *dur = cv.take<float>(n);
    int *dItems = cv.take<int>(n);
    uint8_t *ddesc = cv.take<uint8_t>((size_t)n * 32 + 32);
    int *dCnt = (int *)cv.p; cv.p += cellB;
    int *dOff = (int *)cv.p; cv.p += cellB;
    int *dFill = (int *)cv.p; cv.p += cellB;
    float *dq = cv.take<float>((size_t)n_mp * 3);
    int32_t *dlv = cv.take<int32_t>((size_t)n_mp * 2);
    float *drad = cv.take<float>(n_mp), *dxr = cv.take<float>(n_mp);
    int32_t *dmobs = cv.take<int32_t>(n_mp);
    uint8_t *dact = cv.take<uint8_t>(n_mp);
    uint8_t *dmdesc = cv.take<uint8_t>((size_t)n_mp * 32 + 32);
    int *dqCnt = cv.take<int>(n_mp + 1), *dqOff = cv.take<int>(n_mp + 1);
    int32_t *dn = cv.take<int32_t>(1);
    cudaStream_t s = m->stream;
    MCUDA_TRY(m, cudaMemcpyAsync(dxy, xy.data(), (size_t)n * 8, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(doc, oc.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dobs, obs.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    if (u_right) MCUDA_TRY(m, cudaMemcpyAsync(dur, u_right, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(ddesc, descriptors, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dq, q.data(), (size_t)n_mp * 12, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dlv, lv.data(), (size_t)n_mp * 8, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(drad, rad.data(), (size_t)n_mp * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dxr, xr.data(), (size_t)n_mp * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dmobs, mp_obs, (size_t)n_mp * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dact, act.data(), (size_t)n_mp, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dmdesc, mp_desc, (size_t)n_mp * 32, cudaMemcpyHostToDevice, s));
    rc = grid_build(m, dxy, n, minX, minY, wInv, hInv, dCnt, dOff, dFill, dItems);
    if (rc) return rc;
    k_area_query<false><<<(n_mp + 127) / 128, 128, 0, s>>>(dxy, doc, dOff, dItems, minX, minY, wInv, hInv, dq, n_mp, 0, 0, dqCnt, nullptr, nullptr, dlv, dact);
    k_scan_excl<<<1, 1024, 0, s>>>(dqCnt, n_mp, dqOff);
    int total = 0;
    MCUDA_TRY(m, cudaMemcpyAsync(&total, dqOff + n_mp, 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaStreamSynchronize(s));
    if (total > 0) {
        rc = stage2(m, al256((size_t)total * 4) + al256((size_t)total * 2));
        if (rc) return rc;
        int32_t *dcand = (int32_t *)m->d_buf2;
        uint16_t *ddist = (uint16_t *)(m->d_buf2 + al256((size_t)total * 4));
        k_area_query<true><<<(n_mp + 127) / 128, 128, 0, s>>>(dxy, doc, dOff, dItems, minX, minY, wInv, hInv, dq, n_mp, 0, 0, nullptr, dqOff, dcand, dlv, dact);
        k_sbp_dist<<<(n_mp * 32 + 255) / 256, 256, 0, s>>>((const uint4 *)dmdesc, n_mp, (const uint4 *)ddesc, dcand, dqOff, u_right ? dur : nullptr, dxr, drad, ddist);
        k_sbp_replay<<<1, 32, 0, s>>>(n_mp, dact, dmobs, dcand, dqOff, ddist, doc, n, nnratio, dobs, dasg, dn);
        MCUDA_TRY(m, cudaGetLastError());
        MCUDA_TRY(m, cudaMemcpyAsync(assigned, dasg, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        MCUDA_TRY(m, cudaMemcpyAsync(n_matches, dn, 4, cudaMemcpyDeviceToHost, s));
        MCUDA_TRY(m, cudaStreamSynchronize(s));
    }
    return ORBX_OK;
}

// f_mp == nullptr: SearchByBoW(KeyFrame*, Frame&, …), assigned[n_f] by frame feature.  f_mp != nullptr: SearchByBoW(KeyFrame*, KeyFrame*, …):
// candidates need a good map point, strict threshold; `assigned` is still indexed by the second side (the callers below invert it).
static int search_by_bow_impl(orbx_matcher *m, const uint8_t *kf_desc, const float *kf_angle, int n_kf, const uint8_t *kf_mp, const int32_t *kf_nodes,
                       const int32_t *kf_off, const int32_t *kf_idx, int kf_nn, const uint8_t *f_desc, const float *f_angle, int n_f,
                       const uint8_t *f_mp, const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_idx_in, int f_nn, float nnratio,
                       int check_orientation, int32_t *assigned, int32_t *n_matches) {
    if (!m) return ORBX_ERR_ARG;
    if (n_kf < 0 || n_f < 0 || kf_nn < 0 || f_nn < 0 || !n_matches || (n_f > 0 && (!f_desc || !f_angle || !assigned)) ||
        (n_kf > 0 && (!kf_desc || !kf_angle || !kf_mp)) || (kf_nn > 0 && (!kf_nodes || !kf_off || !kf_idx)) ||
        (f_nn > 0 && (!f_nodes || !f_off || !f_idx_in))) {
        m->err = "orbx_search_by_bow: bad argument";
        return ORBX_ERR_ARG;
    }
    *n_matches = 0;
    for (int i = 0; i < n_f; ++i) assigned[i] = -1;
    if (n_kf == 0 || n_f == 0 || kf_nn == 0 || f_nn == 0) return ORBX_OK;
    // the merge walk over the two node-sorted feature vectors (:243-402; lower_bound = skipping ahead), on the host: it only
    // touches node ids.  Queries = keyframe features with a good map point (:256-260), candidates = the frame's list of that node.
    std::vector<int32_t> qKF, qC0, qC1;
    for (int i = 0; i < f_off[f_nn]; ++i)
        if (f_idx_in[i] < 0 || f_idx_in[i] >= n_f) { m->err = "orbx_search_by_bow: feature index of the second side out of range"; return ORBX_ERR_ARG; }
    // between keyframes a candidate needs a good map point of its own (:819-826): drop the others from the lists, order kept
    std::vector<int32_t> fOffK, fIdxK;
    if (f_mp) {
        fOffK.assign(1, 0);
        for (int b = 0; b < f_nn; ++b) {
            for (int i = f_off[b]; i < f_off[b + 1]; ++i)
                if (f_mp[f_idx_in[i]] == 1) fIdxK.push_back(f_idx_in[i]);
            fOffK.push_back((int32_t)fIdxK.size());
        }
        f_off = fOffK.data();
    }
    const int32_t *f_idx = f_mp ? fIdxK.data() : f_idx_in;
    const int nIdxF = f_off[f_nn];
    for (int a = 0, b = 0; a < kf_nn && b < f_nn;) {
        if (kf_nodes[a] == f_nodes[b]) {
            for (int i = kf_off[a]; i < kf_off[a + 1]; ++i) {
                const int kf = kf_idx[i];
                if (kf < 0 || kf >= n_kf) { m->err = "orbx_search_by_bow: keyframe feature index out of range"; return ORBX_ERR_ARG; }
                if (kf_mp[kf] != 1) continue;
                qKF.push_back(kf); qC0.push_back(f_off[b]); qC1.push_back(f_off[b + 1]);
            }
            ++a; ++b;
        } else if (kf_nodes[a] < f_nodes[b]) {
            ++a;
        } else {
            ++b;
        }
    }
    const int nQ = (int)qKF.size();
    if (nQ == 0 || nIdxF == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    int rc = stage(m, al256((size_t)n_kf * 32 + 32) + al256((size_t)n_kf * 4) + al256((size_t)n_f * 32 + 32) + 2 * al256((size_t)n_f * 4) + al256(n_f) +
                          3 * al256((size_t)nQ * 4) + al256((size_t)nIdxF * 4) + 512);
    if (rc) return rc;
    Carver cv{m->d_buf};
    uint8_t *dkd = cv.take<uint8_t>((size_t)n_kf * 32 + 32);
    float *dka = cv.take<float>(n_kf);
    uint8_t *dfd = cv.take<uint8_t>((size_t)n_f * 32 + 32);
    float *dfa = cv.take<float>(n_f);
    int32_t *dasg = cv.take<int32_t>(n_f);
    int8_t *dbin = cv.take<int8_t>(n_f);
    int32_t *dqKF = cv.take<int32_t>(nQ), *dqC0 = cv.take<int32_t>(nQ), *dqC1 = cv.take<int32_t>(nQ);
    int32_t *dfi = cv.take<int32_t>(nIdxF);
    int32_t *dn = cv.take<int32_t>(1);
    cudaStream_t s = m->stream;
    MCUDA_TRY(m, cudaMemcpyAsync(dkd, kf_desc, (size_t)n_kf * 32, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dka, kf_angle, (size_t)n_kf * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dfd, f_desc, (size_t)n_f * 32, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dfa, f_angle, (size_t)n_f * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dqKF, qKF.data(), (size_t)nQ * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dqC0, qC0.data(), (size_t)nQ * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dqC1, qC1.data(), (size_t)nQ * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dfi, f_idx, (size_t)nIdxF * 4, cudaMemcpyHostToDevice, s));
    k_search_by_bow<<<1, 32, 0, s>>>((const uint4 *)dkd, dka, (const uint4 *)dfd, dfa, n_f, dqKF, dqC0, dqC1, nQ, dfi, nnratio, f_mp ? 49 : 50,
                                     check_orientation ? 1 : 0, dasg, dbin, dn);
    MCUDA_TRY(m, cudaGetLastError());
    MCUDA_TRY(m, cudaMemcpyAsync(assigned, dasg, (size_t)n_f * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(n_matches, dn, 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaStreamSynchronize(s));
    return ORBX_OK;
}

int orbx_search_by_bow(orbx_matcher *m, const uint8_t *kf_desc, const float *kf_angle, int n_kf, const uint8_t *kf_mp, const int32_t *kf_nodes,
                       const int32_t *kf_off, const int32_t *kf_idx, int kf_nn, const uint8_t *f_desc, const float *f_angle, int n_f,
                       const int32_t *f_nodes, const int32_t *f_off, const int32_t *f_idx, int f_nn, float nnratio, int check_orientation,
                       int32_t *assigned, int32_t *n_matches) {
    return search_by_bow_impl(m, kf_desc, kf_angle, n_kf, kf_mp, kf_nodes, kf_off, kf_idx, kf_nn, f_desc, f_angle, n_f, nullptr, f_nodes, f_off, f_idx, f_nn,
                              nnratio, check_orientation, assigned, n_matches);
}

int orbx_search_by_bow_keyframes(orbx_matcher *m, const uint8_t *desc1, const float *angle1, int n1, const uint8_t *mp1, const int32_t *nodes1,
                                 const int32_t *off1, const int32_t *idx1, int nn1, const uint8_t *desc2, const float *angle2, int n2,
                                 const uint8_t *mp2, const int32_t *nodes2, const int32_t *off2, const int32_t *idx2, int nn2, float nnratio,
                                 int check_orientation, int32_t *matches12, int32_t *n_matches) {
    if (!m) return ORBX_ERR_ARG;
    if (n1 < 0 || n2 < 0 || !n_matches || (n1 > 0 && !matches12) || (n2 > 0 && !mp2)) { m->err = "orbx_search_by_bow_keyframes: bad argument"; return ORBX_ERR_ARG; }
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    std::vector<int32_t> by2((size_t)std::max(n2, 1), -1);
    const int rc = search_by_bow_impl(m, desc1, angle1, n1, mp1, nodes1, off1, idx1, nn1, desc2, angle2, n2, mp2, nodes2, off2, idx2, nn2, nnratio,
                                      check_orientation, by2.data(), n_matches);
    if (rc) return rc;
    for (int j = 0; j < n2; ++j)
        if (by2[j] >= 0) matches12[by2[j]] = j;                // every feature of keyframe 2 is matched at most once (vbMatched2)
    return ORBX_OK;
}

int orbx_search_for_initialization_frames(orbx_matcher *m, const orbx_keypoint *kps1, const uint8_t *desc1, int n1, const orbx_keypoint *kps2,
                                          const uint8_t *desc2, int n2, const float *bounds4, float *prev_matched_xy, int window_size,
                                          float nnratio, int check_orientation, int32_t *matches12, int32_t *n_matches) {
    if (!m) return ORBX_ERR_ARG;
    if (n1 < 0 || n2 < 0 || !bounds4 || !n_matches || (n1 > 0 && (!kps1 || !desc1 || !prev_matched_xy || !matches12)) || (n2 > 0 && (!kps2 || !desc2)) ||
        !(bounds4[2] > bounds4[0]) || !(bounds4[3] > bounds4[1])) {
        m->err = "orbx_search_for_initialization_frames: bad argument";
        return ORBX_ERR_ARG;
    }
    *n_matches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return ORBX_OK;
    // F2.GetFeaturesInArea(vbPrevMatched[i1].x, vbPrevMatched[i1].y, windowSize, level1, level1) for level1 == 0 (:661-666)
    std::vector<float> q((size_t)n1 * 3), a1(n1), a2(n2), xy2((size_t)n2 * 2);
    std::vector<int32_t> o1(n1), o2(n2);
    std::vector<uint8_t> act(n1);
    for (int i = 0; i < n1; ++i) {
        q[3 * i] = prev_matched_xy[2 * i]; q[3 * i + 1] = prev_matched_xy[2 * i + 1]; q[3 * i + 2] = (float)window_size;
        a1[i] = kps1[i].angle; o1[i] = kps1[i].octave; act[i] = kps1[i].octave <= 0;
    }
    for (int i = 0; i < n2; ++i) { xy2[2 * i] = kps2[i].x; xy2[2 * i + 1] = kps2[i].y; a2[i] = kps2[i].angle; o2[i] = kps2[i].octave; }
    OrbxDeviceGuard dg_(m->device); MCUDA_TRY(m, dg_.status);
    const float minX = bounds4[0], minY = bounds4[1];
    const float wInv = (float)GRID_COLS / (float)(bounds4[2] - bounds4[0]), hInv = (float)GRID_ROWS / (float)(bounds4[3] - bounds4[1]);
    const size_t cellB = al256((GRID_CELLS + 1) * 4);
    int rc = stage(m, al256((size_t)n1 * 32) + al256((size_t)n2 * 32 + 32) + 4 * al256((size_t)n1 * 4) + al256(n1) * 2 + al256((size_t)n1 * 12) +
                          al256((size_t)n1 * 8) + 2 * al256((size_t)(n1 + 1) * 4) + al256((size_t)n2 * 8) + 5 * al256((size_t)n2 * 4) + 3 * cellB + 512);
    if (rc) return rc;
    Carver cv{m->d_buf};
    uint8_t *dd1 = cv.take<uint8_t>((size_t)n1 * 32), *dd2 = cv.take<uint8_t>((size_t)n2 * 32 + 32);
    float *da1 = cv.take<float>(n1);
    int32_t *do1 = cv.take<int32_t>(n1), *dm12 = cv.take<int32_t>(n1);
    int8_t *dbin = cv.take<int8_t>(n1);
    uint8_t *dact = cv.take<uint8_t>(n1);
    float *dq = cv.take<float>((size_t)n1 * 3);
    float2 *dprev = cv.take<float2>(n1);
    int *dqCnt = cv.take<int>(n1 + 1), *dqOff = cv.take<int>(n1 + 1);
    float2 *dxy2 = cv.take<float2>(n2);
    float *da2 = cv.take<float>(n2);
    int32_t *do2 = cv.take<int32_t>(n2), *dm21 = cv.take<int32_t>(n2), *dmd = cv.take<int32_t>(n2);
    int *dItems = cv.take<int>(n2);
    int *dCnt = (int *)cv.p; cv.p += cellB;
    int *dOff = (int *)cv.p; cv.p += cellB;
    int *dFill = (int *)cv.p; cv.p += cellB;
    int32_t *dn = cv.take<int32_t>(1);
    cudaStream_t s = m->stream;
    MCUDA_TRY(m, cudaMemcpyAsync(dd1, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dd2, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(da1, a1.data(), (size_t)n1 * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(do1, o1.data(), (size_t)n1 * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dact, act.data(), (size_t)n1, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dq, q.data(), (size_t)n1 * 12, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dprev, prev_matched_xy, (size_t)n1 * 8, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(dxy2, xy2.data(), (size_t)n2 * 8, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(da2, a2.data(), (size_t)n2 * 4, cudaMemcpyHostToDevice, s));
    MCUDA_TRY(m, cudaMemcpyAsync(do2, o2.data(), (size_t)n2 * 4, cudaMemcpyHostToDevice, s));
    rc = grid_build(m, dxy2, n2, minX, minY, wInv, hInv, dCnt, dOff, dFill, dItems);
    if (rc) return rc;
    k_area_query<false><<<(n1 + 127) / 128, 128, 0, s>>>(dxy2, do2, dOff, dItems, minX, minY, wInv, hInv, dq, n1, 0, 0, dqCnt, nullptr, nullptr, nullptr, dact);
    k_scan_excl<<<1, 1024, 0, s>>>(dqCnt, n1, dqOff);
    int total = 0;
    MCUDA_TRY(m, cudaMemcpyAsync(&total, dqOff + n1, 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaStreamSynchronize(s));
    rc = stage2(m, al256((size_t)std::max(total, 1) * 4));
    if (rc) return rc;
    int32_t *dcand = (int32_t *)m->d_buf2;
    if (total > 0)
        k_area_query<true><<<(n1 + 127) / 128, 128, 0, s>>>(dxy2, do2, dOff, dItems, minX, minY, wInv, hInv, dq, n1, 0, 0, nullptr, dqOff, dcand, nullptr, dact);
    k_search_init<<<1, 32, 0, s>>>((const uint4 *)dd1, da1, do1, n1, (const uint4 *)dd2, da2, n2, dcand, dqOff, nnratio, check_orientation,
                                   dm12, dm21, dmd, dbin, dn);
    k_update_prev<<<(n1 + 255) / 256, 256, 0, s>>>(dm12, n1, dxy2, dprev);
    MCUDA_TRY(m, cudaGetLastError());
    MCUDA_TRY(m, cudaMemcpyAsync(matches12, dm12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(prev_matched_xy, dprev, (size_t)n1 * 8, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaMemcpyAsync(n_matches, dn, 4, cudaMemcpyDeviceToHost, s));
    MCUDA_TRY(m, cudaStreamSynchronize(s));
    return ORBX_OK;
}

int orbx_descriptor_distance(const uint8_t *a, const uint8_t *b) {
    int d = 0;
    for (int i = 0; i < 32; i += 8) {
        unsigned long long x, y;
        __builtin_memcpy(&x, a + i, 8);
        __builtin_memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

}  // extern "C"
