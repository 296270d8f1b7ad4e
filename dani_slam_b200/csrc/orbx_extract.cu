// orbx_extract.cu — B200 (sm_100a) ORB extractor: the ORBextractor::operator() pipeline of
// /root/reference/src/ORBextractor.cc:1125-1207 as batched CUDA kernels behind the C ABI in
// include/orbx.h.  All file:line citations are relative to /root/reference.
//
// Pipeline for a batch of B equally sized frames (one launch per stage, grid.y = frame):
//   K1 k_pyr_level   ×(L-1)  ComputePyramid :1209-1234 — bilinear 8U resize, 11-bit fixed point
//   K2 k_fast_cells          cell loop :781-869 — one warp per 35-px cell, two phases: compass pre-test on every pixel,
//                            exact FAST-9 measure only for the queued (pixel, side) entries; cell-local NMS, iniTh/minTh
//                            retry, row-major ordered candidate list (k_fast_cells_v1_list redoes over-full cells)
//   K3 k_qt_* / k_quadtree   DANI filter :871-907 + DistributeOctTree :555-779 — candidates classified once
//                            into a quadtree histogram, exact list-order emulation on node counts incl.
//                            libstdc++ sort ties, general single-kernel version as fallback
//   K7 k_assemble            output ordering :1157-1204 (mono from the front, lapping from the back)
//   K5 k_blur                GaussianBlur 7×7 σ=2 :1171-1172 — separable integer, REFLECT_101 halo tiles
//   K4+K6 k_orient_desc      IC_Angle :76-103 + computeOrbDescriptor :107-146 — one warp per keypoint
//
// Everything is integer or individually rounded fp32 (no FMA contraction: __f*_rn intrinsics and
// -fmad=false), so results are bit-identical to the CPU oracle.  No tensor cores: no stage is a
// dense contraction.  The 19-px REFLECT_101 border of mvImagePyramid is never read by this path
//  and is materialised lazily on the host by orbx_get_pyramid.
#include <cuda.h>            // CUtensorMap (the encoder itself is fetched through cudaGetDriverEntryPoint: no libcuda link dependency)
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "glibc_sincosf.h"
#include "orbx_internal.h"
#include "stdsort_port.h"

namespace {

const int8_t h_pattern[1024] = {
#include "orb_pattern.inc"
};

// ------------------------------------------------------------------------------------------------
// kernel argument bundle
// ------------------------------------------------------------------------------------------------
struct ExParams {
    const OrbxGeom *g;
    const uint8_t *in0;       // level 0 source (caller's device frames, or the internal copy)
    long long in0Stride;
    int in0Pitch;
    uint8_t *pyr;             // internal pyramid block, B × frameBytes
    uint8_t *blur;            // blurred levels, same layout as pyr
    const OrbxCell *cells;
    const int2 *tabX;         // per level ≥1: {src index, a0 | a1<<16} per destination column
    const int2 *tabY;
    uint32_t *slots;          // per cell candidate slots: x | y<<8 | score<<16 (cell-ROI coords)
    int *cellCnt;             // candidates per cell
    float2 *ptXY;             // per slot: drifted (x,y) relative to the 16-px border
    uint32_t *ptNode;         // per slot: quadtree node position, ORBX_NODE_ERASED when deleted
    float4 *sel;              // per frame selTotal entries: x, y (level coords), response, -
    int *selCnt;              // per frame × level
    OrbxWork *work;           // per frame selTotal entries
    int *workCnt;             // per frame
    orbx_keypoint *kps;
    uint8_t *desc;
    int cap;
    int *nOut;
    int *monoIdx;
    const int8_t *pattern;    // 1024 bytes
};

__device__ __forceinline__ const uint8_t *level_ptr(const ExParams &p, const OrbxGeom &g, int l, int b,
                                                    int &pitch) {
    if (l == 0) {
        pitch = p.in0Pitch;
        return p.in0 + (long long)b * p.in0Stride;
    }
    pitch = g.lv[l].pitch;
    return p.pyr + (long long)b * g.frameBytes + g.lv[l].off;
}

// ------------------------------------------------------------------------------------------------
// K1: bilinear resize (cv::resize INTER_LINEAR 8UC1), separable inside a block:
// phase 1 forms the horizontal sums (S[sx]*a0 + S[sx+1]*a1) >> 4 (they fit 16 bits) once per needed source
// row into shared memory — a destination column keeps its source offset and coefficients in registers;
// phase 2 combines two of those rows per output row, 4 pixels per thread, one 32-bit store.
// ------------------------------------------------------------------------------------------------
// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers: one elected lane issues a box load, the warp waits on the barrier ----
struct OrbxTmaMaps { CUtensorMap m[ORBX_MAX_LEVELS]; };   // one rank-3 map (x bytes, y rows, frame) per pyramid level
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ORBX_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ORBX_MBAR_DONE_%=;\n"
        "bra ORBX_MBAR_WAIT_%=;\n"
        "ORBX_MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// K1 (TMA form): one pyramid level, one WARP per strip of 128 destination columns × PYR2_RS destination rows, no block barrier.
// Lane 0 fetches the strip's source rows with one TMA box load; every lane owns 4 destination columns, whose source bytes all
// lie in a 8-byte window of the staged row (three aligned words, two PRMT), and walks down the destination rows keeping the
// horizontal sums of the two current source rows in registers — each source row is filtered exactly once per strip and the
// intermediate never touches shared memory.  Arithmetic identical to k_pyr_level below.
// ------------------------------------------------------------------------------------------------
#define PYR2_RS 16
struct Pyr2Args {
    uint8_t *dst; long long dstStride; int dp, dw, dh;
    const int2 *tabX, *tabY;         // same tables as k_pyr_level
    int tilesX, nStrips;             // strips = tilesX × ceil(dh / PYR2_RS)
    int boxW, boxH, slotBytes;       // TMA box (source bytes × source rows per strip), shared memory per warp
};
template <int WPB>
__global__ void __launch_bounds__(WPB * 32) k_pyr_level_tma(Pyr2Args a, const __grid_constant__ CUtensorMap srcMap) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int item = blockIdx.x * WPB + warp, b = blockIdx.y;
    if (item >= a.nStrips) return;                       // warps are independent
    const int ty = item / a.tilesX, tx = item - ty * a.tilesX;
    const int x0 = tx * 128, y0 = ty * PYR2_RS;
    uint8_t *T = smem_raw + (size_t)warp * a.slotBytes;
    uint64_t *bar = reinterpret_cast<uint64_t *>(T + a.slotBytes - 8);
    const int c0 = a.tabX[x0].x & ~15;                   // the box starts on a 16-byte boundary of the source row (TMA rule), at or before the strip's first source column
    // lane yy holds the row table entry of destination row y0+yy (PYR2_RS <= 32)
    const int2 myRow = a.tabY[min(y0 + min(lane, PYR2_RS - 1), a.dh - 1)];
    const int r0 = __shfl_sync(0xffffffffu, myRow.x, 0) & 0xffff;     // first source row of the strip
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, (uint32_t)(a.boxW * a.boxH));
        tma_load_3d(T, &srcMap, bar, c0, r0, b);
    }
    // per-lane column constants (while the box is in flight)
    const int gx = x0 + 4 * lane;
    const bool act = gx < a.dw;
    uint32_t coef[4], sel[4];
    int rel0 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int2 t = a.tabX[min(gx + j, a.dw - 1)];
        const int rel = t.x - c0;
        if (j == 0) rel0 = rel;
        const uint32_t d = (uint32_t)(rel - rel0);       // 0..5: byte of the window holding S[s0]; S[s0+1] follows (a1 = 0 where s0 is the last column)
        sel[j] = d | ((d + 1u) << 4);
        coef[j] = (uint32_t)t.y;                         // a0 | a1 << 16
    }
    const uint32_t selWin = 0x3210u + 0x1111u * (uint32_t)(rel0 & 3);
    const uint8_t *Tw = T + (rel0 & ~3);
    const int boxW = a.boxW;
    auto hrow = [&](const uint8_t *rowp, uint32_t (&h)[4]) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(rowp);
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];
        const uint32_t lo = __byte_perm(w0, w1, selWin), hi = __byte_perm(w1, w2, selWin);   // the 8 bytes from S[s0 of column 0]
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __dp2a_lo(coef[j], __byte_perm(lo, hi, sel[j]), 0u) >> 4;   // (S[s0]*a0 + S[s1]*a1) >> 4
    };
    __syncwarp();
    mbar_wait(bar, 0);
    uint32_t hA[4], hB[4];
    int curA = 0;
    const uint8_t *rowB = Tw + boxW;                     // staged row of hB (the box holds one row more than the strip needs)
    hrow(Tw, hA);
    hrow(rowB, hB);
    uint8_t *Dp = a.dst + (long long)b * a.dstStride + (long long)y0 * a.dp + gx;
    const int nRows = min(PYR2_RS, a.dh - y0);
    for (int yy = 0; yy < nRows; ++yy, Dp += a.dp) {
        const uint32_t rows = (uint32_t)__shfl_sync(0xffffffffu, myRow.x, yy), cf = (uint32_t)__shfl_sync(0xffffffffu, myRow.y, yy);
        const int i0 = (int)(rows & 0xffffu) - r0, i1 = (int)(rows >> 16) - r0;
        while (curA < i0) {                              // warp-uniform: slide the two-row window down
#pragma unroll
            for (int j = 0; j < 4; ++j) hA[j] = hB[j];
            ++curA;
            rowB += boxW;
            hrow(rowB, hB);
        }
        if (i1 == i0) {                                  // clamped at the bottom edge (warp-uniform, stays so for the rest of the strip)
#pragma unroll
            for (int j = 0; j < 4; ++j) hB[j] = hA[j];
        }
        const uint32_t b0 = cf & 0xffffu, b1 = cf >> 16;
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (((b0 * hA[j]) >> 16) + ((b1 * hB[j]) >> 16) + 2u) >> 2;
        const uint32_t out = v[0] + v[1] * 256u + v[2] * 65536u + v[3] * 16777216u;   // v <= 255: bytes pack by multiply-add
        if (act) *reinterpret_cast<uint32_t *>(Dp) = out;   // pitch is a multiple of 128: the padding is writable
    }
}

#define PYR_TW 128
#define PYR_TH 32
struct PyrArgs {             // everything by value: no dependent global loads before the pixel loads
    const uint8_t *src; long long srcStride; int sp, sw, sh;
    uint8_t *dst; long long dstStride; int dp, dw, dh;
    const int2 *tabX, *tabY;         // per destination column {source column, a0 | a1<<16}; per row {i0 | i1<<16 (clamped source rows), b0 | b1<<16}
    const int2 *tileX, *tileY;       // per tile column {first staged source column (16-aligned) | 16-byte chunks<<16, rcp of the chunks}; per tile row {first source row, rows}
    int tilesX, srcRows, srcPitch;   // shared-memory source tile: srcRows × srcPitch bytes (pitch multiple of 16)
};
__global__ void __launch_bounds__(256) k_pyr_level(PyrArgs a) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *T = smem_raw;                                                        // source tile
    uint16_t (*H)[PYR_TW] = reinterpret_cast<uint16_t (*)[PYR_TW]>(smem_raw + a.srcRows * a.srcPitch);   // horizontal sums
    const int b = blockIdx.y;
    const int ty = blockIdx.x / a.tilesX, tx = blockIdx.x - ty * a.tilesX;
    const int x0 = tx * PYR_TW, y0 = ty * PYR_TH;
    const int tid = threadIdx.x;
    const uint8_t *S = a.src + (long long)b * a.srcStride;
    const int sp = a.sp, sw = a.sw;
    const int2 tX = a.tileX[tx], tY = a.tileY[ty];       // tile geometry, precomputed on the host
    const int c0a = tX.x & 0xffff, nChunks = tX.x >> 16, r0 = tY.x, nR = tY.y;
    const uint32_t rcpChunks = (uint32_t)tX.y;           // (i * rcp) >> 16 == i / nChunks for i < 32768
    const bool aligned = ((((unsigned long long)S | (unsigned)sp) & 15ull) == 0);
    // phase 0: stage the source tile with 16-byte loads
    {
        const int rowBytes = aligned ? min(sp, (sw + 15) & ~15) : sw;
        for (int i = tid; i < nR * nChunks; i += 256) {
            const int r = (int)(((uint32_t)i * rcpChunks) >> 16), c = i - r * nChunks;
            const int gx = c0a + 16 * c;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gx < rowBytes) {
                const uint8_t *q = S + (long long)(r0 + r) * sp + gx;
                if (aligned) {
                    v = *reinterpret_cast<const uint4 *>(q);
                } else {
                    uint32_t ww[4] = {0, 0, 0, 0};
                    for (int j = 0; j < 16; ++j)
                        if (gx + j < sw) ww[j >> 2] |= (uint32_t)q[j] << (8 * (j & 3));
                    v = make_uint4(ww[0], ww[1], ww[2], ww[3]);
                }
            }
            *reinterpret_cast<uint4 *>(T + r * a.srcPitch + 16 * c) = v;
        }
    }
    __syncthreads();
    // phase 1: thread owns destination column dx, walks the staged source rows
    {
        const int dx = tid & (PYR_TW - 1);
        const int gx = x0 + dx;
        if (gx < a.dw) {
            const int2 t = a.tabX[gx];
            const int s0 = t.x - c0a, s1 = min(t.x + 1, sw - 1) - c0a;
            const uint32_t coef = (uint32_t)t.y;   // a0 | a1<<16
            const uint8_t *col = T + (tid >> 7) * a.srcPitch;
            const int step = 2 * a.srcPitch;
#pragma unroll 4
            for (int r = tid >> 7; r < nR; r += 256 / PYR_TW, col += step) {
                const uint32_t px = (uint32_t)col[s0] | ((uint32_t)col[s1] << 8);
                H[r][dx] = (uint16_t)(__dp2a_lo(coef, px, 0u) >> 4);   // (S[s0]*a0 + S[s1]*a1) >> 4
            }
        }
    }
    __syncthreads();
    // phase 2: thread owns 4 destination columns of four output rows (8 rows apart); all quantities are non-negative
    // (coefficients 0..2048, sums < 2^16), so the arithmetic shifts of the reference are plain unsigned shifts
    {
        const int cx = (tid & 31) * 4;
        const int gx = x0 + cx;
        if (gx < a.dw) {
            const int yy0 = tid >> 5;
            uint8_t *Dp = a.dst + (long long)b * a.dstStride + (long long)(y0 + yy0) * a.dp + gx;
            const long long rowStep = 8ll * a.dp;
            const int2 *tyP = a.tabY + y0 + yy0;
#pragma unroll
            for (int k = 0; k < PYR_TH / 8; ++k, Dp += rowStep) {
                if (y0 + yy0 + 8 * k < a.dh) {
                    const int2 ty2 = tyP[8 * k];
                    const uint32_t rows = (uint32_t)ty2.x, cf = (uint32_t)ty2.y;
                    const int i0 = (int)(rows & 0xffffu) - r0, i1 = (int)(rows >> 16) - r0;
                    const uint32_t b0 = cf & 0xffffu, b1 = cf >> 16;
                    const uint2 u0 = *reinterpret_cast<const uint2 *>(&H[i0][cx]);
                    const uint2 u1 = *reinterpret_cast<const uint2 *>(&H[i1][cx]);
                    const uint32_t v0 = (((b0 * (u0.x & 0xffffu)) >> 16) + ((b1 * (u1.x & 0xffffu)) >> 16) + 2u) >> 2;
                    const uint32_t v1 = (((b0 * (u0.x >> 16)) >> 16) + ((b1 * (u1.x >> 16)) >> 16) + 2u) >> 2;
                    const uint32_t v2 = (((b0 * (u0.y & 0xffffu)) >> 16) + ((b1 * (u1.y & 0xffffu)) >> 16) + 2u) >> 2;
                    const uint32_t v3 = (((b0 * (u0.y >> 16)) >> 16) + ((b1 * (u1.y >> 16)) >> 16) + 2u) >> 2;
                    // v <= 255: bytes are packed by multiply-add (no masks needed)
                    *reinterpret_cast<uint32_t *>(Dp) = v0 + (v1 << 8) + (v2 << 16) + (v3 << 24);  // pitch multiple of 128: padding is writable
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: FAST-9_16 per cell (cv::FAST + NMS on the cell ROI).  One warp per cell.
//
// Two kernels share the staging and the score formulation.  The single-phase one (fast_cell_v1: ORBX_FAST_V1
// cross-check and the fallback for over-full cells) computes the exact FAST score of EVERY interior pixel without
// any data-dependent branch, two pixels per instruction; the production kernel k_fast_cells further down first
// discards the pixels that fail the compass bound and evaluates the exact measure only for the rest.  Both use
// Blackwell's packed 3-input min/max (VIMNMX3.U16x2, the DPX family):
//   M = max( v − min_k max(ring[k..k+8]),  max_k min(ring[k..k+8]) − v )      (k circular over 16)
// which equals OpenCV's max-over-arcs-of-min|diff| (A4) because min_arc(v−r) = v − max_arc(r).
// A window of 9 is max3(max3(r0,r1,r2), max3(r3,r4,r5), max3(r6,r7,r8)): 32 instructions per polarity
// for all 16 arcs.  A lane owns 4 horizontally adjacent pixels (two u16x2 pairs); ring samples are
// carved out of three aligned 32-bit shared-memory words per row with PRMT.
// ------------------------------------------------------------------------------------------------
struct FastSmem {
    int roiPitch, scorePitch;
    int roiOff, scoreOff, listOff, queueOff, total;     // v1 kernel
    int entryOff, maskOff, total2, qCap;                // two-phase kernel (qCap = entries the queue holds)
};
__host__ __device__ inline FastSmem fast_smem_layout(int maxCw, int maxCh, int maxSlotCap) {
    FastSmem s;
    const int G = (maxCw - 6 + 3) / 4;      // 4-pixel groups per interior row
    s.roiPitch = 4 * G + 8;                 // ROI column x lives at byte x+1; a group reads bytes 4g..4g+11
    s.scorePitch = 4 * G + 8;               // interior column c lives at byte c+4
    s.roiOff = 0;
    s.scoreOff = s.roiOff + s.roiPitch * maxCh;
    s.listOff = s.scoreOff + s.scorePitch * (maxCh - 6 + 2);
    s.queueOff = s.listOff + 4 * maxSlotCap;             // u16 per 4-pixel group: groups whose score word is non-zero
    s.total = (s.queueOff + 2 * G * (maxCh - 6) + 15) & ~15;
    // two-phase kernel: the entry queue holds half of the cell's pixels (a third pass the compass test on corner-dense
    // frames); a cell that needs more is handed to the single-phase kernel (k_fast_cells_v1), like the quadtree's deep
    // fallback — shared memory per warp decides how many warps an SM holds
    const int nPix = 4 * G * (maxCh - 6);
    s.qCap = (nPix / 2 + 1) & ~1;
    s.entryOff = (s.listOff + 3) & ~3;
    s.maskOff = s.entryOff + 2 * s.qCap + 4;             // one pass-mask byte per 4-pixel group, padded to 128 groups per lane quartet
    s.total2 = (s.maskOff + G * (maxCh - 6) + 128 + 15) & ~15;
    return s;
}

struct Row3 { uint32_t w0, w1, w2; };
__device__ __forceinline__ Row3 ld_row3(const uint8_t *p) {
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
    Row3 r; r.w0 = q[0]; r.w1 = q[1]; r.w2 = q[2];
    return r;
}
// bytes (B, B+1) of the 12-byte span {w0,w1,w2} as a zero-extended u16x2
template <int B>
__device__ __forceinline__ uint32_t pair_at(const Row3 &r) {
    static_assert(B >= 0 && B <= 10, "pair outside the 12-byte span");
    if constexpr (B <= 2) return __byte_perm(r.w0, 0u, 0x4040u + B + ((B + 1) << 8));
    else if constexpr (B == 3) return __byte_perm(__byte_perm(r.w0, r.w1, 0x5432u), 0u, 0x4241u);
    else if constexpr (B <= 6) return __byte_perm(r.w1, 0u, 0x4040u + (B - 4) + ((B - 3) << 8));
    else if constexpr (B == 7) return __byte_perm(__byte_perm(r.w1, r.w2, 0x5432u), 0u, 0x4241u);
    else return __byte_perm(r.w2, 0u, 0x4040u + (B - 8) + ((B - 7) << 8));
}

// exact FAST measure minus lowTh, clamped at 0, for the pixel pair P (0/1) of a 4-pixel group
template <int P>
__device__ __forceinline__ uint32_t fast_pair_score(const Row3 (&R)[7], uint32_t biasT2) {
    // ring order k=0..15 = (dx,dy): (0,3)(1,3)(2,2)(3,1)(3,0)(3,-1)(2,-2)(1,-3)(0,-3)(-1,-3)(-2,-2)(-3,-1)(-3,0)(-3,1)(-2,2)(-1,3)
    // R[i] is the row dy = i-3; byte index of pixel pair P at horizontal offset dx is 4+2P+dx
    constexpr int C = 4 + 2 * P;
    uint32_t r[16];
    r[0] = pair_at<C + 0>(R[6]);  r[1] = pair_at<C + 1>(R[6]);  r[2] = pair_at<C + 2>(R[5]);  r[3] = pair_at<C + 3>(R[4]);
    r[4] = pair_at<C + 3>(R[3]);  r[5] = pair_at<C + 3>(R[2]);  r[6] = pair_at<C + 2>(R[1]);  r[7] = pair_at<C + 1>(R[0]);
    r[8] = pair_at<C + 0>(R[0]);  r[9] = pair_at<C - 1>(R[0]);  r[10] = pair_at<C - 2>(R[1]); r[11] = pair_at<C - 3>(R[2]);
    r[12] = pair_at<C - 3>(R[3]); r[13] = pair_at<C - 3>(R[4]); r[14] = pair_at<C - 2>(R[5]); r[15] = pair_at<C - 1>(R[6]);
    const uint32_t v2 = pair_at<C>(R[3]);
    uint32_t tmx[16], tmn[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        tmx[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
        tmn[k] = __vimin3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    }
    uint32_t nmx[16], nmn[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        nmx[k] = __vimax3_u16x2(tmx[k], tmx[(k + 3) & 15], tmx[(k + 6) & 15]);   // max of ring[k..k+8]
        nmn[k] = __vimin3_u16x2(tmn[k], tmn[(k + 3) & 15], tmn[(k + 6) & 15]);   // min of ring[k..k+8]
    }
    uint32_t lo = __vimin3_u16x2(nmx[0], nmx[1], nmx[2]), hi = __vimax3_u16x2(nmn[0], nmn[1], nmn[2]);
#pragma unroll
    for (int k = 3; k < 15; k += 2) {
        lo = __vimin3_u16x2(lo, nmx[k], nmx[k + 1]);
        hi = __vimax3_u16x2(hi, nmn[k], nmn[k + 1]);
    }
    lo = __vminu2(lo, nmx[15]);
    hi = __vmaxu2(hi, nmn[15]);
    // Both differences are taken with a +256 bias per half so that no borrow can cross the halves: plain 32-bit
    // adds (which the compiler may place on the FMA pipe as IMAD) replace the packed 16x2 subtractions that would
    // compete with VIMNMX3 for the ALU pipe.  biasT2 = (256 + lowTh) per half.
    const uint32_t A = (v2 + 0x01000100u) - lo, B = (hi + 0x01000100u) - v2;   // M + 256 candidates, each in [1, 511]
    const uint32_t M = __vmaxu2(A, B);
    return __vmaxu2(M, biasT2) - biasT2;                                      // max(M - lowTh, 0) per half
}

// single-phase FAST of one cell by one warp (`base` = the warp's shared memory, fast_smem_layout(...).total bytes)
__device__ void fast_cell_v1(const ExParams &p, int maxSlotCap, int c, int b, uint8_t *base, int lane) {
    const OrbxGeom &g = *p.g;
    const OrbxCell cell = p.cells[c];
    const FastSmem L = fast_smem_layout(g.maxCw, g.maxCh, maxSlotCap);
    uint8_t *roi = base + L.roiOff;
    uint8_t *score = base + L.scoreOff;
    uint32_t *list = reinterpret_cast<uint32_t *>(base + L.listOff);
    uint16_t *queue = reinterpret_cast<uint16_t *>(base + L.queueOff);

    int pitch;
    const uint8_t *img = level_ptr(p, g, cell.level, b, pitch);
    const int cw = cell.cw, ch = cell.ch;
    const int iw = cw - 6, ih = ch - 6;
    int *cntOut = p.cellCnt + (long long)b * g.nCellsTotal + c;
    if (iw <= 0 || ih <= 0) {  // ROI smaller than 7×7: cv::FAST finds nothing
        if (lane == 0) *cntOut = 0;
        return;
    }
    const int rp = L.roiPitch, sp = L.scorePitch;
    // stage the ROI: column x at byte x+1 so that every 4-pixel group is word aligned
    const uint8_t *src = img + (long long)cell.y0 * pitch + cell.x0;
    if (((((unsigned long long)img) | (unsigned)pitch) & 3ull) == 0) {
        // aligned 32-bit loads: shared-memory word k of a row holds image columns x0-1+4k .. x0+2+4k, i.e. the two
        // aligned global words around it funnel-shifted by the (cell-uniform) misalignment; (row, word) items are
        // flattened over the lanes and four items are in flight per lane
        const int mis = (cell.x0 - 1) & 3;
        const uint32_t *gsrc = reinterpret_cast<const uint32_t *>(src - 1 - mis);
        const int pitchW = pitch >> 2;
        const int nW = (cw + 4) >> 2;
        const int items = nW * ch;
        const uint32_t rcpW = (65536u + nW - 1) / nW;   // (i*rcpW)>>16 == i/nW for i < 3449 (cells are ≤ 20 words × 76 rows)
        uint32_t *roi32 = reinterpret_cast<uint32_t *>(roi);
        const int rpW = rp >> 2;
#pragma unroll 4
        for (int i = lane; i < items; i += 32) {
            const int y = (int)(((uint32_t)i * rcpW) >> 16), k = i - y * nW;
            const uint32_t *q = gsrc + (long long)y * pitchW + k;
            const uint32_t a = q[0], bq = q[1];
            roi32[y * rpW + k] = __funnelshift_r(a, bq, 8 * mis);
        }
    } else {
        for (int y = 0; y < ch; ++y)
            for (int x = lane; x < cw; x += 32) roi[y * rp + x + 1] = src[(long long)y * pitch + x];
    }
    // zero the score map (its 1-px frame of zeros = "outside the cell interior counts 0")
    {
        uint32_t *s32 = reinterpret_cast<uint32_t *>(score);
        const int nw = (sp * (ih + 2)) >> 2;
        for (int i = lane; i < nw; i += 32) s32[i] = 0;
    }
    __syncwarp();

    const int G = (iw + 3) >> 2;             // 4-pixel groups per interior row (≤ 19 for cells ≤ 75 px wide)
    const int nGroups = G * ih;              // groups of the cell in row-major order: lane work items
    const uint32_t rcpG = (65536u + G - 1) / G;   // (i*rcpG)>>16 == i/G for i < 3449 (cells are ≤ 19×69 groups)
    const int t = g.lowTh;
    const uint32_t biasT2 = (uint32_t)(256 + t) * 0x10001u;

    // pass 1: score map (value = M - lowTh clamped at 0; real score = value + lowTh - 1)
    int nQ = 0;   // groups that contain at least one corner, in row-major order
    for (int i0 = 0; i0 < nGroups; i0 += 32) {
        const int gi = i0 + lane;
        uint32_t word = 0;
        if (gi < nGroups) {
            const int yi = (int)(((uint32_t)gi * rcpG) >> 16), lg = gi - yi * G;
            const uint8_t *rowp = roi + yi * rp + 4 * lg;  // ROI row (yi+3)+dy = yi + i for i = 0..6
            Row3 R[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) R[i] = ld_row3(rowp + i * rp);
            const uint32_t s0 = fast_pair_score<0>(R, biasT2), s1 = fast_pair_score<1>(R, biasT2);
            const int nValid = min(iw - 4 * lg, 4);
            const uint32_t colMask = nValid >= 4 ? 0xffffffffu : ((1u << (8 * nValid)) - 1u);
            word = __byte_perm(s0, s1, 0x6420u) & colMask;
            *reinterpret_cast<uint32_t *>(score + (yi + 1) * sp + 4 * lg + 4) = word;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, word != 0);
        if (word != 0) queue[nQ + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)gi;
        nQ += __popc(bal);
    }
    __syncwarp();

    // pass 2: cell-local 3×3 strict NMS → row-major ordered list; count survivors above iniTh
    const int iniRel = g.iniTh - t + 1;      // value >= iniRel  ⇔  M > iniTh
    int n = 0, nIni = 0;
    for (int i0 = 0; i0 < nQ; i0 += 32) {        // only the groups with corners: ≈ a quarter of all groups
        const int qi = i0 + lane;
        const int gi = qi < nQ ? (int)queue[qi] : 0;
        const int yi = (int)(((uint32_t)gi * rcpG) >> 16), lg = gi - yi * G;
        uint32_t flags = 0, vals = 0;        // flags: bit i = pixel i of the group survives
        if (qi < nQ) {
            const uint8_t *rowp = score + yi * sp + 4 * lg;   // score rows yi, yi+1 (centre), yi+2
            const Row3 T = ld_row3(rowp), Cn = ld_row3(rowp + sp), Bt = ld_row3(rowp + 2 * sp);
            vals = Cn.w1;
            {
                // pair 0 centre bytes (4,5), pair 1 centre bytes (6,7)
                const uint32_t tL0 = pair_at<3>(T), tM0 = pair_at<4>(T), tR0 = pair_at<5>(T), tM1 = pair_at<6>(T), tR1 = pair_at<7>(T);
                const uint32_t bL0 = pair_at<3>(Bt), bM0 = pair_at<4>(Bt), bR0 = pair_at<5>(Bt), bM1 = pair_at<6>(Bt), bR1 = pair_at<7>(Bt);
                const uint32_t cL0 = pair_at<3>(Cn), cM0 = pair_at<4>(Cn), cR0 = pair_at<5>(Cn), cM1 = pair_at<6>(Cn), cR1 = pair_at<7>(Cn);
                const uint32_t n0 = __vimax3_u16x2(__vimax3_u16x2(tL0, tM0, tR0), __vimax3_u16x2(bL0, bM0, bR0), __vmaxu2(cL0, cR0));
                const uint32_t n1 = __vimax3_u16x2(__vimax3_u16x2(tR0, tM1, tR1), __vimax3_u16x2(bR0, bM1, bR1), __vmaxu2(cR0, cR1));
                // centre > all 8 neighbours ⇔ bit 15 of (centre + 0x7fff - neighbourMax) per half; no borrow crosses halves
                const uint32_t d0 = (cM0 + 0x7fff7fffu) - n0, d1 = (cM1 + 0x7fff7fffu) - n1;
                flags = ((d0 >> 15) & 1u) | ((d0 >> 30) & 2u) | ((d1 >> 13) & 4u) | ((d1 >> 28) & 8u);
            }
        }
        // a group of 4 adjacent pixels holds at most 2 local maxima
        const int cntLane = __popc(flags);
        uint32_t iniFlags = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if ((flags >> i) & 1u) iniFlags |= (uint32_t)((int)((vals >> (8 * i)) & 0xff) >= iniRel) << i;
        const int cntIni = __popc(iniFlags);
        const uint32_t lt = (1u << lane) - 1u;
        const uint32_t b0 = __ballot_sync(0xffffffffu, cntLane & 1), b1 = __ballot_sync(0xffffffffu, cntLane & 2);
        const uint32_t q0 = __ballot_sync(0xffffffffu, cntIni & 1), q1 = __ballot_sync(0xffffffffu, cntIni & 2);
        int at = n + __popc(b0 & lt) + 2 * __popc(b1 & lt);
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if ((flags >> i) & 1u) {
                const uint32_t s = ((vals >> (8 * i)) & 0xff) + (uint32_t)(t - 1);   // FAST score = M - 1
                list[at++] = (uint32_t)(4 * lg + i + 3) | ((uint32_t)(yi + 3) << 8) | (s << 16);
            }
        n += __popc(b0) + 2 * __popc(b1);
        nIni += __popc(q0) + 2 * __popc(q1);
    }
    __syncwarp();
    // retry rule (:843-846): if the iniTh pass is empty after NMS, the minTh pass is the result
    const int th = nIni > 0 ? g.iniTh : g.minTh;
    uint32_t *out = p.slots + (long long)b * g.slotsTotal + cell.slot;
    int m = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        uint32_t e = 0;
        bool keep = false;
        if (i < n) {
            e = list[i];
            keep = (int)(e >> 16) >= th;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (keep) out[m + __popc(bal & ((1u << lane) - 1))] = e;
        m += __popc(bal);
    }
    if (lane == 0) *cntOut = m;
}

// every cell of the cell table (ORBX_FAST_V1=1: cross-check of the two-phase kernel)
template <int WPB>
__global__ void __launch_bounds__(WPB * 32) k_fast_cells_v1(ExParams p, int maxSlotCap) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * WPB + warp, b = blockIdx.y;
    if (c >= p.g->nCellsTotal) return;  // warps are independent: no block-level barrier
    fast_cell_v1(p, maxSlotCap, c, b, smem_raw + (size_t)warp * fast_smem_layout(p.g->maxCw, p.g->maxCh, maxSlotCap).total, lane);
}
// the cells the two-phase kernel could not hold (list of (frame, cell) appended by k_fast_cells); a small fixed grid
// whose warps stride over the list, so that the usual empty list costs one short launch
template <int WPB>
__global__ void __launch_bounds__(WPB * 32) k_fast_cells_v1_list(ExParams p, int maxSlotCap, const int2 *list, const int *count) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *base = smem_raw + (size_t)warp * fast_smem_layout(p.g->maxCw, p.g->maxCh, maxSlotCap).total;
    const int n = *count;
    for (int i = blockIdx.x * WPB + warp; i < n; i += gridDim.x * WPB) {
        const int2 e = list[i];
        fast_cell_v1(p, maxSlotCap, e.y, e.x, base, lane);
        __syncwarp();
    }
}

// ---- two-phase FAST: compass pre-test on every pixel, exact measure only for the pixels that pass ----
// An arc of 9 contiguous ring pixels contains at least one of every opposite pair {k, k+8}, so
//   A = max_arcs min_k (v - r_k) <= min(max(v-r0, v-r8), max(v-r4, v-r12))   (and the same for B with r - v):
// a pixel whose compass bound is <= t on both sides cannot be a corner at threshold t (the high-speed test of the
// FAST paper, here in exact score form).  About a third of the pixels of a corner-dense frame pass; they are
// compacted into a per-cell queue of (pixel, side) entries and the exact measure is evaluated two entries per lane:
// for side B the ring and the centre are complemented (255 - x), which turns B into the same min-of-window-max form
// as A, so one packed u16x2 op sequence serves any mix of sides.
struct FastEntryRing { uint32_t x[16]; uint32_t v; };

// ceil(65536 / n) for n < 128: (i * rcp) >> 16 == i / n for i < 3449 — the per-cell divisions of the FAST kernel as one constant load
struct Rcp16Table { uint32_t v[128]; };
static constexpr Rcp16Table make_rcp16_table() {
    Rcp16Table t{};
    t.v[0] = 0;
    for (int n = 1; n < 128; ++n) t.v[n] = (65536u + n - 1) / n;
    return t;
}
__constant__ Rcp16Table kRcp16 = make_rcp16_table();
__device__ __forceinline__ uint32_t rcp16(int n) { return n < 128 ? kRcp16.v[n] : (65536u + n - 1) / n; }

// inclusive warp prefix sum: the shuffle's own "source lane in range" predicate guards the add (two instructions per step)
__device__ __forceinline__ int warp_inclusive_sum(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
        asm volatile("{ .reg .s32 t; .reg .pred p; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.s32 %0, %0, t; }" : "+r"(v) : "r"(o));
    return v;
}

__device__ __forceinline__ uint32_t fast_window_minmax(const uint32_t (&r)[16]) {
    uint32_t tmx[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) tmx[k] = __vimax3_u16x2(r[k], r[(k + 1) & 15], r[(k + 2) & 15]);
    uint32_t nmx[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) nmx[k] = __vimax3_u16x2(tmx[k], tmx[(k + 3) & 15], tmx[(k + 6) & 15]);   // max of ring[k..k+8]
    uint32_t lo = __vimin3_u16x2(nmx[0], nmx[1], nmx[2]);
#pragma unroll
    for (int k = 3; k < 15; k += 2) lo = __vimin3_u16x2(lo, nmx[k], nmx[k + 1]);
    return __vminu2(lo, nmx[15]);
}

// exact measure for the entries of q[0..nE): writes max(M - t, 0) into the score map.  PUSH: compact the offsets of the
// non-zero results, in order, into q itself (in place: position <= entries consumed) and return their count.
// Entry: bits 0-13 ROI byte offset of the pixel, bit 15 side (0: the darker-ring side A, 1: the brighter-ring side B; a pixel
// that passes the compass test on both sides comes as side A).  In x-space (x = r for side A, 255 - r for side B, same for
// the centre) both sides are "centre minus the smallest 9-window maximum", and the OTHER side's compass bound is
// min(max(x0,x8), max(x4,x12)) - centre: the rare pixel where that also exceeds t gets the other side evaluated too (a
// warp-uniform branch) and keeps the larger measure, which is the two-sided FAST score.
template <bool PUSH>
__device__ __forceinline__ int fast_exact_entries(const uint8_t *roi, uint8_t *score, uint16_t *q, int nE, int rp, uint32_t biasT2, int lane) {
    int nN = 0;
    const int rp2 = 2 * rp, rp3 = 3 * rp;
    const uint32_t kT = 0x80ff80ffu - biasT2;         // (0x8000 - t - 1) per half
    for (int i0 = 0; i0 < nE; i0 += 64) {
        // lane i takes entries i0+i and i0+32+i: the 32 entries one LDS serves are consecutive in row-major order
        // (they span ~3 ROI rows), which keeps shared-memory bank conflicts low
        const int j0 = i0 + lane, j1 = j0 + 32;
        const bool ok0 = j0 < nE, ok1 = j1 < nE;
        const uint32_t e0 = q[ok0 ? j0 : 0], e1 = ok1 ? (uint32_t)q[j1] : e0;
        if (PUSH) __syncwarp();                       // all lanes hold their entries before anyone compacts into q
        const int o0 = (int)(e0 & 0x3fffu), o1 = (int)(e1 & 0x3fffu);
        // x = r*sg + c per half: side A (bit 15 clear) keeps r, side B takes 255 - r
        const int sg0 = (e0 & 0x8000u) ? -1 : 1, sg1 = (e1 & 0x8000u) ? -65536 : 65536;
        const uint32_t cc = ((e0 & 0x8000u) ? 255u : 0u) | ((e1 & 0x8000u) ? (255u << 16) : 0u);
        const uint8_t *p0 = roi + o0, *p1 = roi + o1;
        uint32_t r[16];
#define ORBX_RING(k, dx, dyoff) r[k] = (uint32_t)((int)p1[(dyoff) + (dx)] * sg1 + ((int)p0[(dyoff) + (dx)] * sg0 + (int)cc));
        ORBX_RING(0, 0, rp3)   ORBX_RING(1, 1, rp3)   ORBX_RING(2, 2, rp2)    ORBX_RING(3, 3, rp)
        ORBX_RING(4, 3, 0)     ORBX_RING(5, 3, -rp)   ORBX_RING(6, 2, -rp2)   ORBX_RING(7, 1, -rp3)
        ORBX_RING(8, 0, -rp3)  ORBX_RING(9, -1, -rp3) ORBX_RING(10, -2, -rp2) ORBX_RING(11, -3, -rp)
        ORBX_RING(12, -3, 0)   ORBX_RING(13, -3, rp)  ORBX_RING(14, -2, rp2)  ORBX_RING(15, -1, rp3)
        const uint32_t v2 = (uint32_t)((int)p1[0] * sg1 + ((int)p0[0] * sg0 + (int)cc));
#undef ORBX_RING
        const uint32_t lo = fast_window_minmax(r);
        const uint32_t M = (v2 + 0x01000100u) - lo;              // M_side + 256 per half, in [1, 511]
        uint32_t val2 = __vmaxu2(M, biasT2) - biasT2;            // max(M_side - t, 0)
        // the other side's compass bound: bit 15 of a half ⇔ bound > t
        const uint32_t two = ((__vminu2(__vmaxu2(r[0], r[8]), __vmaxu2(r[4], r[12])) + kT) - v2) & 0x80008000u;
        if (__any_sync(0xffffffffu, two != 0)) {
#pragma unroll
            for (int k = 0; k < 16; ++k) r[k] ^= 0x00ff00ffu;    // 255 - x per half
            const uint32_t lo2 = fast_window_minmax(r);
            const uint32_t M2 = ((v2 ^ 0x00ff00ffu) + 0x01000100u) - lo2;
            const uint32_t alt = (__vmaxu2(M2, biasT2) - biasT2) & ((two >> 15) * 0xffffu);
            val2 = __vmaxu2(val2, alt);
        }
        const uint32_t val0 = val2 & 0xffffu, val1 = val2 >> 16;
        if (ok0) score[o0 - rp2] = (uint8_t)val0;                // score pitch == ROI pitch, two rows up
        if (ok1) score[o1 - rp2] = (uint8_t)val1;
        if (PUSH) {
            const bool k0 = ok0 && val0 != 0, k1 = ok1 && val1 != 0;
            const uint32_t b0 = __ballot_sync(0xffffffffu, k0), b1 = __ballot_sync(0xffffffffu, k1);
            const uint32_t lt = (1u << lane) - 1u;
            const int n0 = __popc(b0);
            if (k0) q[nN + __popc(b0 & lt)] = (uint16_t)o0;
            if (k1) q[nN + n0 + __popc(b1 & lt)] = (uint16_t)o1;
            nN += n0 + __popc(b1);
        }
    }
    return nN;
}

// A launch covers a range of the cell table whose cells need about the same shared memory ("slot" = the per-warp
// footprint for cells up to cw × ch).  The few much taller cells of the small top levels ride along as "tall" cells that
// take tallSlots adjacent slots each (the warps in between stay idle); their blocks come first in the grid so that
// their longer per-cell latency overlaps the bulk of the work instead of forming a tail.
struct FastRange {
    int cellBase, nCells, cw, ch;                 // regular cells: one warp and one slot each
    int tallBase, nTall, tallCw, tallCh;          // tall cells
    int tallSlots, nTallBlocks;
    int2 *denseList; int *denseN;                 // (frame, cell) pairs whose candidates overflow the entry queue
};
// RP: compile-time ROI pitch (0 = take it from the layout at run time); the usual cells (35-44 px) need 44, 48 or 52 bytes per row.
// (The ROI is staged with ordinary loads: a TMA box must start on a 16-byte boundary of the image row — measured on this B200,
// tools/microbench/tma_probe.cu — which cell ROIs do not; the pyramid and blur tiles, whose origin is free, use TMA.)
#ifndef FAST_GPL
#define FAST_GPL 4     // rounds of 32 four-pixel groups per iteration of the compass loop (1 … 4: the per-round counts share one scan word)
#endif
template <int WPB, int RP>
__global__ void __launch_bounds__(WPB * 32, 2048 / (WPB * 32 * 2)) k_fast_cells(ExParams p, const __grid_constant__ FastRange R) {   // ≤ 64 registers: 32 warps per SM

    extern __shared__ __align__(128) uint8_t smem_raw[];
    const OrbxGeom &g = *p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    const size_t slotBytes = (size_t)fast_smem_layout(R.cw, R.ch, 1).total2;
    int c, maxCw, maxCh;
    if ((int)blockIdx.x < R.nTallBlocks) {
        const int perBlock = WPB / R.tallSlots;
        if (warp % R.tallSlots != 0 || warp / R.tallSlots >= perBlock) return;
        const int cl = blockIdx.x * perBlock + warp / R.tallSlots;
        if (cl >= R.nTall) return;
        c = R.tallBase + cl; maxCw = R.tallCw; maxCh = R.tallCh;
    } else {
        const int cl = (blockIdx.x - R.nTallBlocks) * WPB + warp;
        if (cl >= R.nCells) return;  // warps are independent: no block-level barrier below
        c = R.cellBase + cl; maxCw = R.cw; maxCh = R.ch;
    }
    const OrbxCell cell = p.cells[c];
    const FastSmem L = fast_smem_layout(maxCw, maxCh, 1);
    uint8_t *base = smem_raw + (size_t)warp * slotBytes;
    uint8_t *roi = base + L.roiOff;
    uint8_t *score = base + L.scoreOff;
    uint16_t *queue = reinterpret_cast<uint16_t *>(base + L.entryOff);

    int pitch;
    const uint8_t *img = level_ptr(p, g, cell.level, b, pitch);
    const int cw = cell.cw, ch = cell.ch;
    const int iw = cw - 6, ih = ch - 6;
    int *cntOut = p.cellCnt + (long long)b * g.nCellsTotal + c;
    if (iw <= 0 || ih <= 0) {  // ROI smaller than 7×7: cv::FAST finds nothing
        if (lane == 0) *cntOut = 0;
        return;
    }
    const int rp = RP ? RP : L.roiPitch;   // == score pitch
    // stage the ROI: column x at byte x+1 so that every 4-pixel group is word aligned; the same loop zeroes the score map
    // (pixels that never pass the compass test and the 1-px frame — "outside the cell interior counts 0" — must read 0 in the NMS)
    {
    uint32_t *s32 = reinterpret_cast<uint32_t *>(score);
    const uint8_t *src = img + (long long)cell.y0 * pitch + cell.x0;
    if (((((unsigned long long)img) | (unsigned)pitch) & 3ull) == 0) {
        // aligned 32-bit loads: shared-memory word k of a row holds image columns x0-1+4k .. x0+2+4k, i.e. the two
        // aligned global words around it funnel-shifted by the (cell-uniform) misalignment; (row, word) items are
        // flattened over the lanes and four items are in flight per lane
        const int mis = (cell.x0 - 1) & 3;
        const uint32_t *gsrc = reinterpret_cast<const uint32_t *>(src - 1 - mis);
        const int pitchW = pitch >> 2;
        const int nW = (cw + 4) >> 2;
        const int items = nW * ch;
        const uint32_t rcpW = rcp16(nW);              // (i*rcpW)>>16 == i/nW for i < 3449 (cells are ≤ 20 words × 76 rows)
        uint32_t *roi32 = reinterpret_cast<uint32_t *>(roi);
        const int rpW = rp >> 2;
#pragma unroll 4
        for (int i = lane; i < items; i += 32) {
            const int y = (int)(((uint32_t)i * rcpW) >> 16), k = i - y * nW;
            const uint32_t *q = gsrc + (uint32_t)(y * pitchW + k);      // 32-bit word index: one wide multiply-add per address
            const uint32_t a = q[0], bq = q[1];
            roi32[y * rpW + k] = __funnelshift_r(a, bq, 8 * mis);
            if (y < ih + 2) s32[y * rpW + k] = 0;      // the words of the score map the NMS can read (bytes 3 .. iw+4 of a row)
        }
    } else {
        for (int y = 0; y < ch; ++y)
            for (int x = lane; x < cw; x += 32) roi[y * rp + x + 1] = src[(long long)y * pitch + x];
        const int nw = (rp * (ih + 2)) >> 2;
        for (int i = lane; i < nw; i += 32) s32[i] = 0;
    }
    }
    __syncwarp();

    const int G = (iw + 3) >> 2;             // 4-pixel groups per interior row (≤ 19 for cells ≤ 75 px wide)
    const int nGroups = G * ih;              // groups of the cell in row-major order: lane work items
    const uint32_t rcpG = rcp16(G);             // (i*rcpG)>>16 == i/G for i < 3449 (cells are ≤ 19×69 groups)
    const uint32_t rcpRp = (uint32_t)((0x100000000ull + (unsigned)rp - 1) / (unsigned)rp);   // umulhi(o, rcpRp) == o/rp
    const uint32_t lt = (1u << lane) - 1u;
    const int rp3 = 3 * rp;
    const int stepRows = (int)((32u * rcpG) >> 16), stepGroups = 32 - stepRows * G;          // 32 groups further on
    const int stepOff = stepRows * rp + 4 * stepGroups, wrapOff = rp - 4 * G;
    const int nValidLast = iw - 4 * (G - 1);                               // valid pixels of a row's last group (1..4)
    const uint32_t lastMask = nValidLast >= 4 ? 0x80808080u : (0x80808080u & ((1u << (8 * nValidLast)) - 1u));
    uint32_t *out = p.slots + (long long)b * g.slotsTotal + cell.slot;
    int m = 0;
    bool dense = false;
    // the reference runs cv::FAST at iniThFAST and, when that leaves nothing after NMS, again at minThFAST (:826-846)
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int t = attempt == 0 ? g.iniTh : g.minTh;
        if (attempt == 1 && !(g.minTh < g.iniTh)) break;   // a higher threshold cannot un-suppress anything: stays empty
        const uint32_t biasT2 = (uint32_t)(256 + t) * 0x10001u;
        const uint32_t kT = (uint32_t)(0x8000 - t - 1) * 0x10001u;

        // phase 1: compass test of every 4-pixel group; every passing pixel becomes one queue entry, in row-major pixel order.
        // One iteration covers FAST_GPL rounds of 32 consecutive groups (round h: group i0 + 32·h + lane, so the lanes of a load
        // read consecutive words — no bank conflicts); the per-round counts of a lane travel as the bytes of ONE word through ONE
        // warp scan (a round has at most 128 passing pixels, so no byte overflows), and one round of loop bookkeeping serves
        // 128·FAST_GPL pixels.  Lane state (lg, off) of group i0 + lane advances by 32 groups per round: +stepRows rows and
        // +stepGroups groups, with one conditional row wrap.
        int nE = 0;
        {
            const int row0 = (int)(((uint32_t)lane * rcpG) >> 16);
            int lg = lane - row0 * G;
            int off = (row0 + 3) * rp + 4 * lg + 4;    // ROI byte of the group's first pixel
            for (int i0 = 0; i0 < nGroups; i0 += 32 * FAST_GPL) {
                uint32_t pm[FAST_GPL], sb[FAST_GPL];         // bit 8i+7: pixel i passes / is a side-B entry (does not pass on side A)
                int offs[FAST_GPL];
#pragma unroll
                for (int h = 0; h < FAST_GPL; ++h) {
                    offs[h] = off;
                    pm[h] = 0; sb[h] = 0;
                    if (i0 + 32 * h + lane < nGroups) {
                        const uint8_t *cp = roi + off;
                        Row3 Cn = ld_row3(cp - 4), Tp, Bt;
                        Tp.w1 = *reinterpret_cast<const uint32_t *>(cp - rp3);
                        Bt.w1 = *reinterpret_cast<const uint32_t *>(cp + rp3);
                        Tp.w0 = Tp.w2 = Bt.w0 = Bt.w2 = 0;
                        uint32_t XA[2], XB[2];
#pragma unroll
                        for (int P = 0; P < 2; ++P) {
                            const uint32_t v2 = P ? pair_at<6>(Cn) : pair_at<4>(Cn);
                            const uint32_t r0 = P ? pair_at<6>(Bt) : pair_at<4>(Bt), r8 = P ? pair_at<6>(Tp) : pair_at<4>(Tp);
                            const uint32_t r4 = P ? pair_at<9>(Cn) : pair_at<7>(Cn), r12 = P ? pair_at<3>(Cn) : pair_at<1>(Cn);
                            const uint32_t hiMin = __vmaxu2(__vminu2(r0, r8), __vminu2(r4, r12));   // A-side bound: v - hiMin
                            const uint32_t loMax = __vminu2(__vmaxu2(r0, r8), __vmaxu2(r4, r12));   // B-side bound: loMax - v
                            XA[P] = (v2 + kT) - hiMin;       // bit 15 of a half ⇔ bound > t (no borrow crosses the halves)
                            XB[P] = (loMax + kT) - v2;
                        }
                        const uint32_t colMask = lg == G - 1 ? lastMask : 0x80808080u;
                        const uint32_t pA = __byte_perm(XA[0], XA[1], 0x7531u);       // high bytes of the four halves: px0..px3
                        const uint32_t pB = __byte_perm(XB[0], XB[1], 0x7531u);
                        pm[h] = (pA | pB) & colMask;
                        sb[h] = pm[h] & ~pA;
                    }
                    lg += stepGroups; off += stepOff;        // 32 groups further on
                    if (lg >= G) { lg -= G; off += wrapOff; }
                }
                uint32_t anyPm = pm[0];
#pragma unroll
                for (int h = 1; h < FAST_GPL; ++h) anyPm |= pm[h];
                if (__any_sync(0xffffffffu, anyPm != 0)) {
                    uint32_t cnts = 0;                       // byte h: this lane's passing pixels of round h
#pragma unroll
                    for (int h = 0; h < FAST_GPL; ++h) cnts |= (uint32_t)__popc(pm[h]) << (8 * h);
                    const uint32_t incl = (uint32_t)warp_inclusive_sum((int)cnts);
                    const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
                    const int nNew = __dp4a(tot, 0x01010101u, 0u);
                    if (nE + nNew > L.qCap) { dense = true; break; }      // warp-uniform: more corners than the queue holds
                    const uint32_t excl = incl - cnts;
                    int base = nE;
#pragma unroll
                    for (int h = 0; h < FAST_GPL; ++h) {
                        uint16_t *qa = queue + (base + (int)((excl >> (8 * h)) & 0xffu));
                        if (pm[h] & 0x80u) *qa++ = (uint16_t)((uint32_t)offs[h] | ((sb[h] << 8) & 0x8000u));
                        if (pm[h] & 0x8000u) *qa++ = (uint16_t)((uint32_t)(offs[h] + 1) | (sb[h] & 0x8000u));
                        if (pm[h] & 0x800000u) *qa++ = (uint16_t)((uint32_t)(offs[h] + 2) | ((sb[h] >> 8) & 0x8000u));
                        if (pm[h] & 0x80000000u) *qa++ = (uint16_t)((uint32_t)(offs[h] + 3) | ((sb[h] >> 16) & 0x8000u));
                        base += (int)((tot >> (8 * h)) & 0xffu);
                    }
                    nE += nNew;
                }
            }
        }
        if (dense) break;
        __syncwarp();
        // phase 2: exact measure of the entries
        const int nN = fast_exact_entries<true>(roi, score, queue, nE, rp, biasT2, lane);
        __syncwarp();
        // phase 3: cell-local 3×3 strict NMS of the corners, survivors in row-major order straight to the slots
        for (int i0 = 0; i0 < nN; i0 += 32) {
            const int i = i0 + lane;
            bool keep = false;
            uint32_t e = 0;
            if (i < nN) {
                const int o = (int)queue[i];
                const uint8_t *sc = score + o - 2 * rp;
                const uint32_t cv = sc[0];
                const uint32_t n8 = __vimax3_u32(__vimax3_u32(sc[-rp - 1], sc[-rp], sc[-rp + 1]), __vimax3_u32(sc[rp - 1], sc[rp], sc[rp + 1]),
                                                 max((uint32_t)sc[-1], (uint32_t)sc[1]));
                keep = cv > n8;
                const int y = (int)__umulhi((uint32_t)o, rcpRp), x = o - y * rp - 1;     // ROI coordinates
                e = (uint32_t)x | ((uint32_t)y << 8) | ((cv + (uint32_t)(t - 1)) << 16);  // FAST score = M - 1
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) out[m + __popc(bal & lt)] = e;
            m += __popc(bal);
        }
        if (m > 0) break;
        __syncwarp();
    }
    if (lane == 0) {
        if (dense) R.denseList[atomicAdd(R.denseN, 1)] = make_int2(b, c);    // the single-phase kernel redoes this cell
        else *cntOut = m;
    }
}

// ------------------------------------------------------------------------------------------------
// K3: DANI filter + quadtree (DistributeOctTree), one block per (frame, level)
// ------------------------------------------------------------------------------------------------
#define QT_THREADS 128   // default block size; large-nFeatures handles use QT_THREADS_BIG
#define QT_THREADS_BIG 512

// block-wide exclusive scan of a[0..n) in place; returns the total.  All threads must call.
template <int NT>
__device__ int block_exclusive_scan(int *a, int n, int *scratch /* >= NT+1 ints */) {
    const int tid = threadIdx.x;
    const int per = (n + NT - 1) / NT;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    scratch[tid] = sum;
    __syncthreads();
    if (tid < 32) {
        // 256 partial sums: each lane of warp 0 scans 8 of them
        int loc[NT / 32];
        int s = 0;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) { loc[k] = s; s += scratch[tid * (NT / 32) + k]; }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
        }
        const int excl = incl - s;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) scratch[tid * (NT / 32) + k] = excl + loc[k];
        if (tid == 31) scratch[NT] = incl;
    }
    __syncthreads();
    int run = scratch[tid];
    for (int i = lo; i < hi; ++i) { const int v = a[i]; a[i] = run; run += v; }
    const int total = scratch[NT];
    __syncthreads();
    return total;
}

struct QtSmem {
    int boxOff[2], cntOff[2], childCntOff, childPosOff, keptPosOff, pendOff, pendIdxOff, sortOff,
        bestOff, prefixOff, scratchOff, tmp4Off, total;
};
__host__ __device__ inline QtSmem qt_smem_layout(int nodeCap, int maxCellsLevel) {
    QtSmem s;
    int o = 0;
    s.sortOff = o; o += 8 * nodeCap;               // 64-bit sort elements first (alignment)
    s.boxOff[0] = o; o += 8 * nodeCap;             // short4 per node
    s.boxOff[1] = o; o += 8 * nodeCap;
    s.cntOff[0] = o; o += 4 * nodeCap;
    s.cntOff[1] = o; o += 4 * nodeCap;
    s.childCntOff = o; o += 16 * nodeCap;
    s.childPosOff = o; o += 16 * nodeCap;
    s.keptPosOff = o; o += 4 * nodeCap;
    s.pendOff = o; o += 4 * nodeCap;
    s.pendIdxOff = o; o += 4 * nodeCap;
    s.bestOff = o; o += 4 * nodeCap;
    s.prefixOff = o; o += 4 * (maxCellsLevel + 1);
    s.scratchOff = o; o += 4 * (QT_THREADS_BIG + 2);
    o = (o + 15) & ~15;                            // tmp4 doubles as a 64-bit buffer for the rank sort
    s.tmp4Off = o; o += 16 * nodeCap;
    s.total = (o + 15) & ~15;
    return s;
}

// quadrant of a point inside a node (DivideNode :480-536): children n1..n4 = 0..3.
// halfX = ceil(static_cast<float>(UR.x-UL.x)/2) of a non-negative int < 2^24 is exactly (w+1)>>1.
__host__ __device__ __forceinline__ int qt_quadrant(short4 bx, float x, float y) {
    const int hx = (bx.y - bx.x + 1) >> 1, hy = (bx.w - bx.z + 1) >> 1;
    const float xm = (float)(bx.x + hx), ym = (float)(bx.z + hy);
    return (x < xm ? 0 : 1) + (y < ym ? 0 : 2);
}
__host__ __device__ __forceinline__ short4 qt_child_box(short4 bx, int q) {  // box = {x0, x1, y0, y1}
    const int hx = (bx.y - bx.x + 1) >> 1, hy = (bx.w - bx.z + 1) >> 1;
    const short xm = (short)(bx.x + hx), ym = (short)(bx.z + hy);
    short4 c;
    c.x = (q & 1) ? xm : bx.x;
    c.y = (q & 1) ? bx.y : xm;
    c.z = (q & 2) ? ym : bx.z;
    c.w = (q & 2) ? bx.w : ym;
    return c;
}

// Visit every live candidate of the level with 4 loads in flight per thread (the passes are bound by
// L2 latency, not bandwidth): f(index, nodeWord, xy).
template <bool NEED_XY, int NT, class F>
__device__ __forceinline__ void qt_for_points(const uint32_t *ptNode, const float2 *ptXY, int nPts, F f) {
    for (int i0 = threadIdx.x; i0 < nPts; i0 += 4 * NT) {
        uint32_t v[4];
        float2 xy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * NT;
            v[u] = i < nPts ? ptNode[i] : ORBX_NODE_ERASED;
        }
        if (NEED_XY) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * NT;
                xy[u] = i < nPts ? ptXY[i] : make_float2(0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (v[u] != ORBX_NODE_ERASED) f(i0 + u * NT, v[u], xy[u]);
    }
}

// Per-point state lives in two DENSE per-(frame, level) arrays indexed by the candidate's order index
// (prefix over cells in processing order + rank inside the cell — the position it would have in the
// reference's vToDistributeKeys):  ptXY[i] = drifted (x, y);  ptNode[i] = node position (bits 15:0) |
// FAST score (bits 23:16) | quadrant scratch (bits 31:30), or ORBX_NODE_ERASED.
template <int NT>
__global__ void __launch_bounds__(NT) k_quadtree(ExParams p, int nodeCapMax, int maxCellsLevel, const int *onlyFlagged) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const OrbxGeom &g = *p.g;
    const int l = blockIdx.x, b = blockIdx.y;
    if (onlyFlagged && !onlyFlagged[b * g.nlevels + l]) return;   // the histogram kernel already produced this (frame, level)
    const OrbxLevel &LV = g.lv[l];
    const int tid = threadIdx.x;
    const QtSmem S = qt_smem_layout(nodeCapMax, maxCellsLevel);
    orbx_sort::elem_t *sortbuf = reinterpret_cast<orbx_sort::elem_t *>(smem_raw + S.sortOff);
    short4 *box[2] = {reinterpret_cast<short4 *>(smem_raw + S.boxOff[0]), reinterpret_cast<short4 *>(smem_raw + S.boxOff[1])};
    int *cnt[2] = {reinterpret_cast<int *>(smem_raw + S.cntOff[0]), reinterpret_cast<int *>(smem_raw + S.cntOff[1])};
    int *childCnt = reinterpret_cast<int *>(smem_raw + S.childCntOff);
    int *childPos = reinterpret_cast<int *>(smem_raw + S.childPosOff);
    int *keptPos = reinterpret_cast<int *>(smem_raw + S.keptPosOff);
    int *pend = reinterpret_cast<int *>(smem_raw + S.pendOff);
    int *pendIdx = reinterpret_cast<int *>(smem_raw + S.pendIdxOff);
    unsigned *best = reinterpret_cast<unsigned *>(smem_raw + S.bestOff);
    int *prefix = reinterpret_cast<int *>(smem_raw + S.prefixOff);
    int *scratch = reinterpret_cast<int *>(smem_raw + S.scratchOff);
    int *tmp4 = reinterpret_cast<int *>(smem_raw + S.tmp4Off);
    __shared__ int sh_size, sh_C;

    const int nCells = LV.nCells;
    const int N = LV.quota;
    const int *cellCnt = p.cellCnt + (long long)b * g.nCellsTotal + LV.cellBase;
    const OrbxCell *cells = p.cells + LV.cellBase;
    const uint32_t *slots = p.slots + (long long)b * g.slotsTotal;
    float2 *ptXY = p.ptXY + (long long)b * g.slotsTotal + LV.slotBase;
    uint32_t *ptNode = p.ptNode + (long long)b * g.slotsTotal + LV.slotBase;
    float4 *sel = p.sel + (long long)b * g.selTotal + LV.selBase;
    int *selCntOut = p.selCnt + b * g.nlevels + l;

    for (int c = tid; c < nCells; c += NT) prefix[c] = cellCnt[c];
    __syncthreads();
    const int nPts = block_exclusive_scan<NT>(prefix, nCells, scratch);
    if (tid == 0) prefix[nCells] = nPts;
    if (nPts == 0 || LV.nIni < 1) {
        if (tid == 0) *selCntOut = 0;
        return;
    }
    const int nIni = LV.nIni;
    for (int i = tid; i < nIni; i += NT) childCnt[i] = 0;
    __syncthreads();

    // ---- DANI dynamic-area deletion with its fp32 round trip (:871-907) + root binning
    const float sc = LV.sf, inv = __fdiv_rn(1.f, sc);
    const float hX = LV.hX;
    const int nRects = g.nRects;
    for (int i = tid; i < nPts; i += NT) {
        int lo = 0, hi = nCells;  // last cell with prefix[c] <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (prefix[mid] <= i) lo = mid; else hi = mid;
        }
        const OrbxCell cell = cells[lo];
        const uint32_t e = slots[cell.slot + (i - prefix[lo])];
        const int trips = nCells - cell.seq;  // filter passes this cell's keypoints live through
        float x = __fadd_rn((float)(e & 0xff), (float)(cell.cx * LV.wCell));  // pt.x += j*wCell (:863)
        float y = __fadd_rn((float)((e >> 8) & 0xff), (float)(cell.cy * LV.hCell));
        bool erased = false;
        for (int t = 0; t < trips; ++t) {
            const float xs = __fmul_rn(__fadd_rn(x, (float)ORBX_BORDER), sc);
            const float ys = __fmul_rn(__fadd_rn(y, (float)ORBX_BORDER), sc);
            if (nRects > 0) {
                const int px = __float2int_rn(xs), py = __float2int_rn(ys);  // Point2f → Point2i
                for (int r = 0; r < nRects; ++r) {
                    const int rx = g.rects[4 * r], ry = g.rects[4 * r + 1];
                    if (rx <= px && px < rx + g.rects[4 * r + 2] && ry <= py && py < ry + g.rects[4 * r + 3]) {
                        erased = true;
                        break;
                    }
                }
                if (erased) break;
            }
            const float xn = __fsub_rn(__fmul_rn(xs, inv), (float)ORBX_BORDER);
            const float yn = __fsub_rn(__fmul_rn(ys, inv), (float)ORBX_BORDER);
            const bool fixed = (xn == x) && (yn == y);
            x = xn;
            y = yn;
            if (fixed) break;  // later trips reproduce this one exactly
        }
        ptXY[i] = make_float2(x, y);
        if (erased) {
            ptNode[i] = ORBX_NODE_ERASED;
        } else {
            int bin = (int)__fdiv_rn(x, hX);  // vpIniNodes[kp.pt.x/hX] (:584)
            bin = min(max(bin, 0), nIni - 1);
            ptNode[i] = (uint32_t)bin | ((e >> 16) << 16);
            atomicAdd(&childCnt[bin], 1);
        }
    }
    __syncthreads();
    // initial list: non-empty roots in order (:565-601)
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < nIni; ++i) {
            keptPos[i] = -1;
            if (childCnt[i] > 0) {
                short4 bx;
                bx.x = (short)(int)__fmul_rn(hX, (float)i);
                bx.y = (short)(int)__fmul_rn(hX, (float)(i + 1));
                bx.z = 0;
                bx.w = (short)(LV.maxBY - ORBX_BORDER);
                box[0][m] = bx;
                cnt[0][m] = childCnt[i];
                keptPos[i] = m++;
            }
        }
        sh_size = m;
    }
    __syncthreads();
    int size = sh_size;
    if (size != nIni) {  // some root was empty: renumber
        qt_for_points<false, NT>(ptNode, ptXY, nPts, [&](int i, uint32_t v, float2) {
            ptNode[i] = (v & 0x00ff0000u) | (uint32_t)keptPos[v & 0xffffu];
        });
        __syncthreads();
    }

    int cur = 0;
    bool done = (size == 0);
    bool phase2 = false;
    int nPend = 0;
    while (!done) {
        const int prevSize = size;
        const int nxt = cur ^ 1;
        short4 *bxC = box[cur], *bxN = box[nxt];
        int *cnC = cnt[cur], *cnN = cnt[nxt];
        if (!phase2) {
            // ---- full pass: split every node holding more than one point (:616-683)
            for (int i = tid; i < 4 * size; i += NT) childCnt[i] = 0;
            __syncthreads();
            qt_for_points<true, NT>(ptNode, ptXY, nPts, [&](int i, uint32_t v, float2 xy) {
                const uint32_t nd = v & 0xffffu;
                if (cnC[nd] <= 1) return;
                const int q = qt_quadrant(bxC[nd], xy.x, xy.y);
                atomicAdd(&childCnt[4 * nd + q], 1);
                ptNode[i] = (v & 0x00ffffffu) | ((uint32_t)q << 30);
            });
            __syncthreads();
            // creation order = list order × child order; children go to the list front (reversed)
            for (int i = tid; i < 4 * size; i += NT) {
                childPos[i] = childCnt[i] > 0 ? 1 : 0;
            }
            for (int i = tid; i < size; i += NT) keptPos[i] = cnC[i] <= 1 ? 1 : 0;
            __syncthreads();
            const int C = block_exclusive_scan<NT>(childPos, 4 * size, scratch);
            const int nKept = block_exclusive_scan<NT>(keptPos, size, scratch);
            for (int i = tid; i < 4 * size; i += NT) {
                const int n = childCnt[i];
                if (n > 0) {
                    const int pos = C - 1 - childPos[i];
                    bxN[pos] = qt_child_box(bxC[i >> 2], i & 3);
                    cnN[pos] = n;
                    childPos[i] = pos;
                } else {
                    childPos[i] = -1;
                }
            }
            for (int i = tid; i < size; i += NT) {
                if (cnC[i] <= 1) {
                    const int pos = C + keptPos[i];
                    bxN[pos] = bxC[i];
                    cnN[pos] = cnC[i];
                    keptPos[i] = pos;
                } else {
                    keptPos[i] = -1;
                }
            }
            __syncthreads();
            qt_for_points<false, NT>(ptNode, ptXY, nPts, [&](int i, uint32_t v, float2) {
                const uint32_t nd = v & 0xffffu, q = v >> 30;
                ptNode[i] = (v & 0x00ff0000u) | (uint32_t)(cnC[nd] <= 1 ? keptPos[nd] : childPos[4 * nd + q]);
            });
            // expandable children (more than one point) in creation order: scan over (list position, child)
            for (int i = tid; i < 4 * prevSize; i += NT) tmp4[i] = childCnt[i] > 1 ? 1 : 0;
            __syncthreads();
            nPend = block_exclusive_scan<NT>(tmp4, 4 * prevSize, scratch);
            for (int i = tid; i < 4 * prevSize; i += NT)
                if (childCnt[i] > 1) pend[tmp4[i]] = childPos[i];
            __syncthreads();
            size = C + nKept;
            cur = nxt;
            if (size >= N || size == prevSize) done = true;
            else if (size + 3 * nPend > N) phase2 = true;
        } else {
            // ---- "largest first" pass (:685-751): sort pending by (size, UL.x) exactly like std::sort
            for (int i = tid; i < size; i += NT) pendIdx[i] = -1;
            for (int i = tid; i < 4 * nPend; i += NT) childCnt[i] = 0;
            __syncthreads();
            for (int i = tid; i < nPend; i += NT) {
                const int pos = pend[i];
                pendIdx[pos] = i;
                const unsigned long long key = ((unsigned long long)(unsigned)cnC[pos] << 16) | (unsigned short)bxC[pos].x;
                sortbuf[i] = (key << orbx_sort::kPayloadBits) | (unsigned long long)i;
            }
            __syncthreads();
            if (tid == 0) orbx_sort::sort(sortbuf, nPend);   // overlaps with the classification below
            qt_for_points<true, NT>(ptNode, ptXY, nPts, [&](int i, uint32_t v, float2 xy) {
                const uint32_t nd = v & 0xffffu;
                const int pi = pendIdx[nd];
                if (pi < 0) return;
                const int q = qt_quadrant(bxC[nd], xy.x, xy.y);
                atomicAdd(&childCnt[4 * pi + q], 1);
                ptNode[i] = (v & 0x00ffffffu) | ((uint32_t)q << 30);
            });
            __syncthreads();
            // walk the sorted array from the back until the list reaches N nodes (:701-747)
            if (tid == 0) {
                int sz = size, c = 0;
                for (int j = nPend - 1; j >= 0; --j) {
                    const int pi = (int)(sortbuf[j] & ((1ull << orbx_sort::kPayloadBits) - 1));
                    for (int q = 0; q < 4; ++q) {
                        if (childCnt[4 * pi + q] > 0) { childPos[4 * pi + q] = c++; ++sz; }
                        else childPos[4 * pi + q] = -1;
                    }
                    --sz;
                    best[pi] = 1;  // processed marker, indexed by pending index
                    if (sz >= N) {
                        for (int jj = j - 1; jj >= 0; --jj) best[(int)(sortbuf[jj] & ((1ull << orbx_sort::kPayloadBits) - 1))] = 0;
                        break;
                    }
                }
                sh_C = c;
                sh_size = sz;
            }
            __syncthreads();
            const int C = sh_C;
            // survivors of the old list keep their relative order behind the new children
            for (int i = tid; i < size; i += NT) {
                const int pi = pendIdx[i];
                keptPos[i] = (pi >= 0 && best[pi] == 1) ? 0 : 1;
            }
            __syncthreads();
            block_exclusive_scan<NT>(keptPos, size, scratch);
            for (int i = tid; i < size; i += NT) {
                const int pi = pendIdx[i];
                if (pi >= 0 && best[pi] == 1) {
                    keptPos[i] = -1;
                } else {
                    const int pos = C + keptPos[i];
                    bxN[pos] = bxC[i];
                    cnN[pos] = cnC[i];
                    keptPos[i] = pos;
                }
            }
            for (int i = tid; i < 4 * nPend; i += NT) {
                const int pi = i >> 2;
                if (best[pi] == 1 && childCnt[i] > 0) {
                    const int pos = C - 1 - childPos[i];
                    bxN[pos] = qt_child_box(bxC[pend[pi]], i & 3);
                    cnN[pos] = childCnt[i];
                    childPos[i] = pos;
                } else {
                    childPos[i] = -1;
                }
            }
            __syncthreads();
            qt_for_points<false, NT>(ptNode, ptXY, nPts, [&](int i, uint32_t v, float2) {
                const uint32_t nd = v & 0xffffu, q = v >> 30;
                const int pi = pendIdx[nd];
                ptNode[i] = (v & 0x00ff0000u) | (uint32_t)((pi >= 0 && best[pi] == 1) ? childPos[4 * pi + q] : keptPos[nd]);
            });
            __syncthreads();
            // next pending list: expandable children in creation order = processing order × child order
            if (tid == 0) {
                int np = 0;
                for (int j = nPend - 1; j >= 0; --j) {
                    const int pi = (int)(sortbuf[j] & ((1ull << orbx_sort::kPayloadBits) - 1));
                    if (best[pi] != 1) break;
                    for (int q = 0; q < 4; ++q)
                        if (childCnt[4 * pi + q] > 1) pendIdx[np++] = childPos[4 * pi + q];
                }
                for (int i = 0; i < np; ++i) pend[i] = pendIdx[i];
                sh_C = np;
            }
            __syncthreads();
            nPend = sh_C;
            size = sh_size;
            cur = nxt;
            if (size >= N || size == prevSize) done = true;
        }
        __syncthreads();
    }

    // ---- best response per node, first maximum in insertion order wins (:757-776)
    for (int i = tid; i < size; i += NT) best[i] = 0;
    __syncthreads();
    qt_for_points<false, NT>(ptNode, ptXY, nPts, [&](int i, uint32_t v, float2) {
        atomicMax(&best[v & 0xffffu], (((v >> 16) & 0xffu) << 24) | (0xffffffu - (unsigned)i));
    });
    __syncthreads();
    for (int i = tid; i < size; i += NT) {
        const int order = (int)(0xffffffu - (best[i] & 0xffffffu));
        const float2 xy = ptXY[order];
        // :919-923 — add the border back; response = FAST score
        sel[i] = make_float4(__fadd_rn(xy.x, (float)ORBX_BORDER), __fadd_rn(xy.y, (float)ORBX_BORDER), (float)(best[i] >> 24), 0.f);
    }
    if (tid == 0) *selCntOut = size;
}

// ------------------------------------------------------------------------------------------------
// K3 (fast path): the same DistributeOctTree emulation split into throughput-parallel and serial parts.
// The quadtree geometry is data independent (a node's children are its box cut at the ceil-halves), so every
// candidate's path down to depth QT_DMAX is computed ONCE and counted in a per-(frame, level) histogram of the
// depth-QT_DMAX cells; the coarser levels are sums of four.
//   k_qt_prefix    (level, frame)  prefix over the level's cell counts = candidate order index; clears the tables
//   k_qt_classify  warp per cell   DANI round trips (:871-907), root bin (:584), path code, histogram atomics
//   k_qt_nodes     (level, frame)  list emulation on node COUNTS only (reads a child's count from the table);
//                                  std::sort = introsort partitioning + parallel stable rank sort; writes the table
//                                  of final nodes; raises the `deep` flag if a depth-QT_DMAX node must be split
//   k_quadtree     (flagged only)  the general kernel redoes flagged (frame, level) pairs from the raw slots
//   k_qt_attach    warp per cell   every candidate walks its own path to its final node; atomicMax of
//                                  (response, first-in-order) per node (:757-776)
//   k_qt_select    (level, frame)  writes the selected keypoints in list order (:919-923)
// ------------------------------------------------------------------------------------------------
#define QT_DMAX 6
#define QT_TREE 5461                      // Σ_{d=0..6} 4^d nodes per root
#define QT_MAX_INI 12
__host__ __device__ __forceinline__ int qt_off(int nIni, int d) { return nIni * (((1 << (2 * d)) - 1) / 3); }

struct QtTables {               // per-handle device tables of the fast path (indexed by frame, then level)
    unsigned *hist;             // [frame][level][maxIni·QT_TREE] point counts per tree node
    unsigned short *finalPos;   // same shape: list position + 1 of final nodes, 0 elsewhere
    unsigned *best;             // [frame][selTotal] best (response<<24 | 0xffffff-order) per final node
    int *cellPrefix;            // [frame][nCellsTotal] exclusive prefix of the cell counts inside their level
    int *deep;                  // [frame][level] 1 = the general kernel must redo this pair
    const unsigned short *pathLut;   // per level: spread x-path codes per (root, floor x) and spread y-path codes per floor y
    int maxIni;
};

template <int NT>
__global__ void __launch_bounds__(NT) k_qt_prefix(ExParams p, QtTables t, int maxCellsLevel) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    int *prefix = reinterpret_cast<int *>(smem_raw);
    int *scratch = prefix + maxCellsLevel + 1;
    const OrbxGeom &g = *p.g;
    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const OrbxLevel &LV = g.lv[l];
    const int nCells = LV.nCells;
    const int *cellCnt = p.cellCnt + (long long)b * g.nCellsTotal + LV.cellBase;
    for (int c = tid; c < nCells; c += NT) prefix[c] = cellCnt[c];
    __syncthreads();
    block_exclusive_scan<NT>(prefix, nCells, scratch);
    int *out = t.cellPrefix + (long long)b * g.nCellsTotal + LV.cellBase;
    for (int c = tid; c < nCells; c += NT) out[c] = prefix[c];
    const long long tb = ((long long)b * g.nlevels + l) * (long long)t.maxIni * QT_TREE;
    const int treeN = min(LV.nIni, t.maxIni) * QT_TREE;
    for (int i = tid; i < treeN; i += NT) t.hist[tb + i] = 0;
    unsigned *best = t.best + (long long)b * g.selTotal + LV.selBase;
    for (int i = tid; i < LV.selCap; i += NT) best[i] = 0;
    if (tid == 0) t.deep[b * g.nlevels + l] = (LV.nIni > QT_MAX_INI) ? 1 : 0;
}

#define QT_CLS_BLOCKS 8
// grid (QT_CLS_BLOCKS, level, frame): one thread per candidate (grid-stride inside the level), the candidate's cell is
// found by a binary search over the level's cell prefix, so all lanes stay busy whatever the cell occupancies are
__global__ void __launch_bounds__(128) k_qt_classify(ExParams p, QtTables t) {
    const OrbxGeom &g = *p.g;
    const int l = blockIdx.y, b = blockIdx.z;
    const OrbxLevel &LV = g.lv[l];
    const int nCells = LV.nCells;
    if (nCells == 0) return;
    const int *cellCnt = p.cellCnt + (long long)b * g.nCellsTotal + LV.cellBase;
    const int *prefix = t.cellPrefix + (long long)b * g.nCellsTotal + LV.cellBase;
    const OrbxCell *cells = p.cells + LV.cellBase;
    const uint32_t *slots = p.slots + (long long)b * g.slotsTotal;
    float2 *ptXY = p.ptXY + (long long)b * g.slotsTotal + LV.slotBase;
    uint32_t *ptNode = p.ptNode + (long long)b * g.slotsTotal + LV.slotBase;
    const int nIni = LV.nIni;
    const bool tabled = nIni <= QT_MAX_INI;
    unsigned *hist = t.hist + ((long long)b * g.nlevels + l) * (long long)t.maxIni * QT_TREE + qt_off(nIni, QT_DMAX);
    const float sc = LV.sf, inv = __fdiv_rn(1.f, sc);
    const float hX = LV.hX;
    const int nRects = g.nRects;
    const int rootH = LV.maxBY - ORBX_BORDER;
    // one warp per cell (its candidates are consecutive in the order index): the cell record and prefix are read once per warp
    const int lane = threadIdx.x & 31;
    for (int ci = blockIdx.x * 4 + (threadIdx.x >> 5); ci < nCells; ci += QT_CLS_BLOCKS * 4)
    for (int j = lane, cnt = cellCnt[ci]; j < cnt; j += 32) {
        const OrbxCell cell = cells[ci];
        const int i = prefix[ci] + j;
        const uint32_t e = slots[cell.slot + j];
        const int trips = nCells - cell.seq;   // filter passes this cell's keypoints live through (:871-907)
        float x = __fadd_rn((float)(e & 0xff), (float)(cell.cx * LV.wCell));  // pt.x += j*wCell (:863)
        float y = __fadd_rn((float)((e >> 8) & 0xff), (float)(cell.cy * LV.hCell));
        bool erased = false;
        for (int tr = 0; tr < trips; ++tr) {
            const float xs = __fmul_rn(__fadd_rn(x, (float)ORBX_BORDER), sc);
            const float ys = __fmul_rn(__fadd_rn(y, (float)ORBX_BORDER), sc);
            if (nRects > 0) {
                const int px = __float2int_rn(xs), py = __float2int_rn(ys);  // Point2f → Point2i
                for (int r = 0; r < nRects; ++r) {
                    const int rx = g.rects[4 * r], ry = g.rects[4 * r + 1];
                    if (rx <= px && px < rx + g.rects[4 * r + 2] && ry <= py && py < ry + g.rects[4 * r + 3]) {
                        erased = true;
                        break;
                    }
                }
                if (erased) break;
            }
            const float xn = __fsub_rn(__fmul_rn(xs, inv), (float)ORBX_BORDER);
            const float yn = __fsub_rn(__fmul_rn(ys, inv), (float)ORBX_BORDER);
            const bool fixed = (xn == x) && (yn == y);
            x = xn;
            y = yn;
            if (fixed) break;  // later trips reproduce this one exactly
        }
        ptXY[i] = make_float2(x, y);
        if (erased) {
            ptNode[i] = ORBX_NODE_ERASED;
            continue;
        }
        int bin = 0;
        if (nIni > 1) {                       // with a single root the clamp below makes every candidate land in it
            bin = (int)__fdiv_rn(x, hX);      // vpIniNodes[kp.pt.x/hX] (:584)
            bin = min(max(bin, 0), nIni - 1);
        }
        short4 bx;
        bx.x = (short)(int)__fmul_rn(hX, (float)bin);
        bx.y = (short)(int)__fmul_rn(hX, (float)(bin + 1));
        bx.z = 0;
        bx.w = (short)rootH;
        unsigned code = (unsigned)bin;
        if (LV.lutW > 0) {
            // The x and y decisions of the six splits are independent, and every split coordinate is an integer, so
            // (x < xm) == (floor(x) < xm): the path is a table lookup on the integer parts (a coordinate the drift
            // pushed just below 0 takes the leftmost path like 0 does).
            const int xi = min(max((int)floorf(x), 0), LV.lutW - 1), yi = min(max((int)floorf(y), 0), LV.lutH - 1);
            code = code * 4096u + t.pathLut[LV.lutX + bin * LV.lutW + xi] + t.pathLut[LV.lutY + yi];
        } else {
#pragma unroll
            for (int d = 0; d < QT_DMAX; ++d) {
                const int q = qt_quadrant(bx, x, y);
                bx = qt_child_box(bx, q);
                code = code * 4u + (unsigned)q;
            }
        }
        ptNode[i] = (code & 0xffffu) | ((e >> 16) << 16);
        if (tabled) atomicAdd(&hist[code], 1u);
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) k_qt_nodes(ExParams p, QtTables t, int nodeCapMax) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const OrbxGeom &g = *p.g;
    const int l = blockIdx.x, b = blockIdx.y;
    const OrbxLevel &LV = g.lv[l];
    const int tid = threadIdx.x;
    const QtSmem S = qt_smem_layout(nodeCapMax, 0);
    orbx_sort::elem_t *sortbuf = reinterpret_cast<orbx_sort::elem_t *>(smem_raw + S.sortOff);
    short4 *box[2] = {reinterpret_cast<short4 *>(smem_raw + S.boxOff[0]), reinterpret_cast<short4 *>(smem_raw + S.boxOff[1])};
    int *cnt[2] = {reinterpret_cast<int *>(smem_raw + S.cntOff[0]), reinterpret_cast<int *>(smem_raw + S.cntOff[1])};
    int *childCnt = reinterpret_cast<int *>(smem_raw + S.childCntOff);
    int *childPos = reinterpret_cast<int *>(smem_raw + S.childPosOff);
    int *keptPos = reinterpret_cast<int *>(smem_raw + S.keptPosOff);
    int *pend = reinterpret_cast<int *>(smem_raw + S.pendOff);
    int *pendIdx = reinterpret_cast<int *>(smem_raw + S.pendIdxOff);
    unsigned *best = reinterpret_cast<unsigned *>(smem_raw + S.bestOff);
    int *scratch = reinterpret_cast<int *>(smem_raw + S.scratchOff);
    int *tmp4 = reinterpret_cast<int *>(smem_raw + S.tmp4Off);
    unsigned *nid[2] = {reinterpret_cast<unsigned *>(smem_raw + S.total), reinterpret_cast<unsigned *>(smem_raw + S.total + 4 * nodeCapMax)};
    __shared__ int sh_size, sh_C, sh_deep;

    const int N = LV.quota;
    const int nIni = LV.nIni;
    int *selCntOut = p.selCnt + b * g.nlevels + l;
    int *deepOut = t.deep + b * g.nlevels + l;
    if (nIni > QT_MAX_INI || nIni < 1) return;   // flagged by k_qt_prefix: the general kernel handles it
    const long long tb = ((long long)b * g.nlevels + l) * (long long)t.maxIni * QT_TREE;
    unsigned *hist = t.hist + tb;
    unsigned short *finalPos = t.finalPos + tb;
    const float hX = LV.hX;
    const int rootH = LV.maxBY - ORBX_BORDER;
    if (tid == 0) sh_deep = 0;
    __syncthreads();

    // coarser levels = sums of four children (reads bypass L1: the counts were produced by atomics)
    for (int d = QT_DMAX - 1; d >= 0; --d) {
        const int n = nIni << (2 * d), o = qt_off(nIni, d), oc = qt_off(nIni, d + 1);
        for (int i = tid; i < n; i += NT)
            hist[o + i] = __ldcg(&hist[oc + 4 * i]) + __ldcg(&hist[oc + 4 * i + 1]) + __ldcg(&hist[oc + 4 * i + 2]) + __ldcg(&hist[oc + 4 * i + 3]);
        __syncthreads();
    }
    // initial list: non-empty roots in order (:565-601)
    if (tid == 0) {
        int m = 0;
        for (int i = 0; i < nIni; ++i) {
            const int c = (int)__ldcg(&hist[i]);
            if (c > 0) {
                short4 bx;
                bx.x = (short)(int)__fmul_rn(hX, (float)i);
                bx.y = (short)(int)__fmul_rn(hX, (float)(i + 1));
                bx.z = 0;
                bx.w = (short)rootH;
                box[0][m] = bx;
                cnt[0][m] = c;
                nid[0][m] = (unsigned)i;   // depth 0
                ++m;
            }
        }
        sh_size = m;
    }
    __syncthreads();
    int size = sh_size;

    int cur = 0;
    bool done = (size == 0);
    bool phase2 = false;
    int nPend = 0;
    while (!done) {
        const int prevSize = size;
        const int nxt = cur ^ 1;
        short4 *bxC = box[cur], *bxN = box[nxt];
        int *cnC = cnt[cur], *cnN = cnt[nxt];
        unsigned *idC = nid[cur], *idN = nid[nxt];
        if (!phase2) {
            // ---- full pass: split every node holding more than one point (:616-683); child counts from the table
            for (int i = tid; i < 4 * size; i += NT) {
                const int pn = i >> 2;
                int c = 0;
                if (cnC[pn] > 1) {
                    const unsigned id = idC[pn], d = id >> 24, ix = id & 0xffffffu;
                    if (d >= QT_DMAX) sh_deep = 1;
                    else c = (int)__ldcg(&hist[qt_off(nIni, d + 1) + 4 * ix + (i & 3)]);
                }
                childCnt[i] = c;
                childPos[i] = c > 0 ? 1 : 0;
            }
            for (int i = tid; i < size; i += NT) keptPos[i] = cnC[i] <= 1 ? 1 : 0;
            __syncthreads();
            if (sh_deep) break;
            const int C = block_exclusive_scan<NT>(childPos, 4 * size, scratch);
            const int nKept = block_exclusive_scan<NT>(keptPos, size, scratch);
            for (int i = tid; i < 4 * size; i += NT) {
                const int n = childCnt[i];
                if (n > 0) {
                    const int pos = C - 1 - childPos[i];
                    const unsigned id = idC[i >> 2];
                    bxN[pos] = qt_child_box(bxC[i >> 2], i & 3);
                    cnN[pos] = n;
                    idN[pos] = (((id >> 24) + 1u) << 24) | ((id & 0xffffffu) * 4u + (unsigned)(i & 3));
                    childPos[i] = pos;
                } else {
                    childPos[i] = -1;
                }
            }
            for (int i = tid; i < size; i += NT) {
                if (cnC[i] <= 1) {
                    const int pos = C + keptPos[i];
                    bxN[pos] = bxC[i];
                    cnN[pos] = cnC[i];
                    idN[pos] = idC[i];
                }
            }
            // expandable children (more than one point) in creation order: scan over (list position, child)
            for (int i = tid; i < 4 * size; i += NT) tmp4[i] = childCnt[i] > 1 ? 1 : 0;
            __syncthreads();
            nPend = block_exclusive_scan<NT>(tmp4, 4 * size, scratch);
            for (int i = tid; i < 4 * size; i += NT)
                if (childCnt[i] > 1) pend[tmp4[i]] = childPos[i];
            __syncthreads();
            size = C + nKept;
            cur = nxt;
            if (size >= N || size == prevSize) done = true;
            else if (size + 3 * nPend > N) phase2 = true;
        } else {
            // ---- "largest first" pass (:685-751): sort pending by (size, UL.x) exactly like std::sort
            for (int i = tid; i < size; i += NT) pendIdx[i] = -1;
            __syncthreads();
            for (int i = tid; i < nPend; i += NT) {
                const int pos = pend[i];
                pendIdx[pos] = i;
                const unsigned long long key = ((unsigned long long)(unsigned)cnC[pos] << 16) | (unsigned short)bxC[pos].x;
                sortbuf[i] = (key << orbx_sort::kPayloadBits) | (unsigned long long)i;
                const unsigned id = idC[pos], d = id >> 24, ix = id & 0xffffffu;
                if (d >= QT_DMAX) {
                    sh_deep = 1;
                } else {
                    const int o = qt_off(nIni, d + 1) + 4 * ix;
#pragma unroll
                    for (int q = 0; q < 4; ++q) childCnt[4 * i + q] = (int)__ldcg(&hist[o + q]);
                }
            }
            __syncthreads();
            if (sh_deep) break;
            // std::sort = serial introsort partitioning (thread 0) + its final insertion sort, which is a stable sort
            // of the partitioned order and is done here as a parallel rank sort
                    if (tid == 0) orbx_sort::introsort_loop_only(sortbuf, nPend);
            __syncthreads();
                    {
                orbx_sort::elem_t *sorted = reinterpret_cast<orbx_sort::elem_t *>(tmp4);
                for (int i = tid; i < nPend; i += NT) {
                    const orbx_sort::elem_t e = sortbuf[i];
                    const unsigned long long k = e >> orbx_sort::kPayloadBits;
                    int rank = 0;
                    for (int j = 0; j < nPend; ++j) {
                        const unsigned long long kj = sortbuf[j] >> orbx_sort::kPayloadBits;
                        rank += (kj < k) || (kj == k && j < i);
                    }
                    sorted[rank] = e;
                }
                __syncthreads();
                for (int i = tid; i < nPend; i += NT) sortbuf[i] = sorted[i];
                __syncthreads();
            }
                    // walk the sorted array from the back until the list reaches N nodes (:701-747), in parallel: processing
            // step r handles sorted element nPend-1-r; a scan over the children counts gives every step its creation
            // index base and the list size after it, the first step reaching N is the break point
            if (tid == 0) sh_C = nPend - 1;   // break step (all processed unless some step reaches N)
            for (int r = tid; r < nPend; r += NT) {
                const int pi = (int)(sortbuf[nPend - 1 - r] & ((1ull << orbx_sort::kPayloadBits) - 1));
                tmp4[r] = (childCnt[4 * pi] > 0) + (childCnt[4 * pi + 1] > 0) + (childCnt[4 * pi + 2] > 0) + (childCnt[4 * pi + 3] > 0);
            }
            __syncthreads();
            block_exclusive_scan<NT>(tmp4, nPend, scratch);
            for (int r = tid; r < nPend; r += NT) {
                const int pi = (int)(sortbuf[nPend - 1 - r] & ((1ull << orbx_sort::kPayloadBits) - 1));
                const int kk = (childCnt[4 * pi] > 0) + (childCnt[4 * pi + 1] > 0) + (childCnt[4 * pi + 2] > 0) + (childCnt[4 * pi + 3] > 0);
                if (size + tmp4[r] + kk - (r + 1) >= N) atomicMin(&sh_C, r);
            }
            __syncthreads();
            const int rBreak = sh_C;
            __syncthreads();
            for (int r = tid; r < nPend; r += NT) {
                const int pi = (int)(sortbuf[nPend - 1 - r] & ((1ull << orbx_sort::kPayloadBits) - 1));
                const bool proc = r <= rBreak;
                best[pi] = proc ? 1u : 0u;  // processed marker, indexed by pending index
                int c = tmp4[r];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool has = childCnt[4 * pi + q] > 0;
                    childPos[4 * pi + q] = (proc && has) ? c : -1;
                    c += has;
                }
                if (r == rBreak) { sh_C = c; sh_size = size + c - (r + 1); }
            }
            if (tid == 0 && nPend == 0) { sh_C = 0; sh_size = size; }   // nothing left to expand: the list stays as it is
            __syncthreads();
            const int C = sh_C;
            for (int i = tid; i < size; i += NT) {
                const int pi = pendIdx[i];
                keptPos[i] = (pi >= 0 && best[pi] == 1) ? 0 : 1;
            }
            __syncthreads();
            block_exclusive_scan<NT>(keptPos, size, scratch);
            for (int i = tid; i < size; i += NT) {
                const int pi = pendIdx[i];
                if (!(pi >= 0 && best[pi] == 1)) {
                    const int pos = C + keptPos[i];
                    bxN[pos] = bxC[i];
                    cnN[pos] = cnC[i];
                    idN[pos] = idC[i];
                }
            }
            for (int i = tid; i < 4 * nPend; i += NT) {
                const int pi = i >> 2;
                if (best[pi] == 1 && childCnt[i] > 0) {
                    const int pos = C - 1 - childPos[i];
                    const unsigned id = idC[pend[pi]];
                    bxN[pos] = qt_child_box(bxC[pend[pi]], i & 3);
                    cnN[pos] = childCnt[i];
                    idN[pos] = (((id >> 24) + 1u) << 24) | ((id & 0xffffffu) * 4u + (unsigned)(i & 3));
                    childPos[i] = pos;
                } else {
                    childPos[i] = -1;
                }
            }
            __syncthreads();
                    // next pending list: expandable children in creation order (creation index c = C-1-position)
            for (int i = tid; i < C; i += NT) tmp4[i] = 0;
            __syncthreads();
            for (int i = tid; i < 4 * nPend; i += NT)
                if (childPos[i] >= 0 && childCnt[i] > 1) tmp4[C - 1 - childPos[i]] = 1;
            __syncthreads();
            const int nNext = block_exclusive_scan<NT>(tmp4, C, scratch);
            for (int i = tid; i < 4 * nPend; i += NT)
                if (childPos[i] >= 0 && childCnt[i] > 1) pendIdx[tmp4[C - 1 - childPos[i]]] = childPos[i];
            __syncthreads();
            for (int i = tid; i < nNext; i += NT) pend[i] = pendIdx[i];
            if (tid == 0) sh_C = nNext;
            __syncthreads();
            nPend = sh_C;
            size = sh_size;
            cur = nxt;
            if (size >= N || size == prevSize) done = true;
        }
        __syncthreads();
        }
    if (sh_deep) {  // a node deeper than the table would have to be split: hand over to the general kernel
        if (tid == 0) *deepOut = 1;
        return;
    }


    // leaf → final node table for k_qt_attach: a final node at depth d owns 4^(QT_DMAX-d) consecutive leaves (leaves under
    // dropped empty children hold no candidate and are never looked up, so the table needs no clearing)
    {
        const int warp = tid >> 5, lane = tid & 31, nWarps = NT / 32;
        for (int i = warp; i < size; i += nWarps) {
            const unsigned id = nid[cur][i], d = id >> 24, ix = id & 0xffffffu;
            const int span = 1 << (2 * (QT_DMAX - d));
            unsigned short *dst = finalPos + (size_t)ix * span;
            for (int k = lane; k < span; k += 32) dst[k] = (unsigned short)(i + 1);
        }
    }
    if (tid == 0) *selCntOut = size;
}

// grid (QT_CLS_BLOCKS, level, frame): every candidate looks up the final node of its depth-QT_DMAX leaf
__global__ void __launch_bounds__(128) k_qt_attach(ExParams p, QtTables t) {
    const OrbxGeom &g = *p.g;
    const int l = blockIdx.y, b = blockIdx.z;
    if (t.deep[b * g.nlevels + l]) return;           // redone by the general kernel
    const OrbxLevel &LV = g.lv[l];
    const int nCells = LV.nCells;
    if (nCells == 0) return;
    const int nPts = t.cellPrefix[(long long)b * g.nCellsTotal + LV.cellBase + nCells - 1] + p.cellCnt[(long long)b * g.nCellsTotal + LV.cellBase + nCells - 1];
    const uint32_t *ptNode = p.ptNode + (long long)b * g.slotsTotal + LV.slotBase;
    const unsigned short *finalPos = t.finalPos + ((long long)b * g.nlevels + l) * (long long)t.maxIni * QT_TREE;
    unsigned *best = t.best + (long long)b * g.selTotal + LV.selBase;
    for (int i = blockIdx.x * 128 + threadIdx.x; i < nPts; i += QT_CLS_BLOCKS * 128) {
        const uint32_t v = ptNode[i];
        if (v == ORBX_NODE_ERASED) continue;
        const unsigned pos = finalPos[v & 0xffffu];
        // first maximum in insertion order wins: larger key = larger response, then smaller order index
        if (pos) atomicMax(&best[pos - 1], (((v >> 16) & 0xffu) << 24) | (0xffffffu - (unsigned)i));
    }
}

// attach + select in one block per (level, frame): the per-node maxima live in shared memory (no global atomics, one launch less)
template <int NT>
__global__ void __launch_bounds__(NT) k_qt_attach_select(ExParams p, QtTables t) {
    extern __shared__ __align__(16) unsigned s_best[];
    const OrbxGeom &g = *p.g;
    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    if (t.deep[b * g.nlevels + l]) return;           // the general kernel wrote this level's list itself
    const OrbxLevel &LV = g.lv[l];
    const int size = p.selCnt[b * g.nlevels + l];
    for (int i = tid; i < size; i += NT) s_best[i] = 0;
    __syncthreads();
    const int nCells = LV.nCells;
    if (nCells > 0) {
        const int nPts = t.cellPrefix[(long long)b * g.nCellsTotal + LV.cellBase + nCells - 1] + p.cellCnt[(long long)b * g.nCellsTotal + LV.cellBase + nCells - 1];
        const uint32_t *ptNode = p.ptNode + (long long)b * g.slotsTotal + LV.slotBase;
        const unsigned short *finalPos = t.finalPos + ((long long)b * g.nlevels + l) * (long long)t.maxIni * QT_TREE;
        for (int i = tid; i < nPts; i += NT) {
            const uint32_t v = ptNode[i];
            if (v == ORBX_NODE_ERASED) continue;
            const unsigned pos = finalPos[v & 0xffffu];
            // first maximum in insertion order wins: larger key = larger response, then smaller order index
            if (pos) atomicMax(&s_best[pos - 1], (((v >> 16) & 0xffu) << 24) | (0xffffffu - (unsigned)i));
        }
    }
    __syncthreads();
    const float2 *ptXY = p.ptXY + (long long)b * g.slotsTotal + LV.slotBase;
    float4 *sel = p.sel + (long long)b * g.selTotal + LV.selBase;
    for (int i = tid; i < size; i += NT) {
        const unsigned k = s_best[i];
        const float2 xy = ptXY[0xffffffu - (k & 0xffffffu)];
        // :919-923 — add the border back; response = FAST score
        sel[i] = make_float4(__fadd_rn(xy.x, (float)ORBX_BORDER), __fadd_rn(xy.y, (float)ORBX_BORDER), (float)(k >> 24), 0.f);
    }
}

__global__ void __launch_bounds__(128) k_qt_select(ExParams p, QtTables t) {
    const OrbxGeom &g = *p.g;
    const int l = blockIdx.x, b = blockIdx.y;
    if (t.deep[b * g.nlevels + l]) return;           // the general kernel wrote this level's list itself
    const OrbxLevel &LV = g.lv[l];
    const int size = p.selCnt[b * g.nlevels + l];
    const unsigned *best = t.best + (long long)b * g.selTotal + LV.selBase;
    const float2 *ptXY = p.ptXY + (long long)b * g.slotsTotal + LV.slotBase;
    float4 *sel = p.sel + (long long)b * g.selTotal + LV.selBase;
    for (int i = threadIdx.x; i < size; i += 128) {
        const unsigned k = best[i];
        const float2 xy = ptXY[0xffffffu - (k & 0xffffffu)];
        // :919-923 — add the border back; response = FAST score
        sel[i] = make_float4(__fadd_rn(xy.x, (float)ORBX_BORDER), __fadd_rn(xy.y, (float)ORBX_BORDER), (float)(k >> 24), 0.f);
    }
}

// ------------------------------------------------------------------------------------------------
// K7: output ordering (:1157-1204).  One block per frame.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_assemble(ExParams p) {
    const OrbxGeom &g = *p.g;
    const int b = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_cnt[ORBX_MAX_LEVELS + 1];
    __shared__ int s_warp[8];
    __shared__ int s_run;
    if (tid == 0) {
        int t = 0;
        for (int l = 0; l < g.nlevels; ++l) { s_cnt[l] = t; t += p.selCnt[b * g.nlevels + l]; }
        s_cnt[g.nlevels] = t;
        s_run = 0;
        p.nOut[b] = t;
    }
    __syncthreads();
    const int total = s_cnt[g.nlevels];
    const bool fits = total <= p.cap;
    const float4 *sel = p.sel + (long long)b * g.selTotal;
    orbx_keypoint *kps = p.kps + (long long)b * p.cap;
    OrbxWork *work = p.work + (long long)b * g.selTotal;
    const float lap0 = (float)g.lap0, lap1 = (float)g.lap1;
    for (int base = 0; base < total; base += 256) {
        const int i = base + tid;
        int l = 0;
        bool mono = false, valid = i < total;
        float4 s = make_float4(0, 0, 0, 0);
        float x = 0, y = 0;
        if (valid) {
            while (i >= s_cnt[l + 1]) ++l;
            s = sel[g.lv[l].selBase + (i - s_cnt[l])];
            x = s.x; y = s.y;
            if (l != 0) { x = __fmul_rn(x, g.lv[l].sf); y = __fmul_rn(y, g.lv[l].sf); }  // :1188-1190
            mono = !(x >= lap0 && x <= lap1);                                              // :1192
        }
        // rank among mono keypoints in order
        const unsigned bal = __ballot_sync(0xffffffffu, mono);
        const int lane = tid & 31, warp = tid >> 5;
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const int monoRank = before + __popc(bal & ((1u << lane) - 1));
        if (valid) {
            const int out = mono ? monoRank : (total - 1 - (i - monoRank));  // lapping ones fill from the back
            if (fits) {
                orbx_keypoint k;
                k.x = x; k.y = y;
                k.size = (float)g.lv[l].patch_size;
                k.angle = -1.f;
                k.response = s.z;
                k.octave = l;
                k.class_id = -1;
                kps[out] = k;
            }
            OrbxWork w;
            w.level = l;
            w.cx = __float2int_rn(s.x);
            w.cy = __float2int_rn(s.y);
            w.out = fits ? out : -1;
            work[i] = w;
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 8; ++w) t += s_warp[w];
            s_run += t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        p.monoIdx[b] = s_run;
        p.workCnt[b] = fits ? total : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// K5: 7×7 Gaussian blur, σ=2, integer separable kernel, REFLECT_101 on the level.
// The block stages its input tile (+3-row / +16-byte halo) with 16-byte loads into shared memory; a thread
// then owns 4 adjacent pixels (one 32-bit store) and marches down BLUR_RH rows: per input row it reads
// three aligned shared-memory words, forms the four horizontal sums with ten DP4A (the tap words slide over the
// aligned words, nothing is realigned) and keeps the last 7 rows of sums in registers (fully unrolled ring) for
// the vertical pass.
// ------------------------------------------------------------------------------------------------
#define BLUR_TW 256          // pixels per block row-strip (64 threads × 4 px)
#define BLUR_RH 16           // output rows per thread
#define BLUR_STRIPS 2        // row strips per block (blockDim.y)
#define BLUR_TH (BLUR_RH * BLUR_STRIPS)
#define BLUR_SP (BLUR_TW + 32)   // tile pitch: columns x0-16 .. x0+271
struct BlurTile { short level, tx, ty, pad; };
__constant__ uint32_t kBlurTaps[10] = {
    (18u << 8) | (34u << 16) | (48u << 24), 56u | (48u << 8) | (34u << 16) | (18u << 24),                    // pixel 0: words 0, 1
    (18u << 16) | (34u << 24), 48u | (56u << 8) | (48u << 16) | (34u << 24), 18u,                            // pixel 1: words 0, 1, 2
    18u << 24, 34u | (48u << 8) | (56u << 16) | (48u << 24), 34u | (18u << 8),                               // pixel 2: words 0, 1, 2
    18u | (34u << 8) | (48u << 16) | (56u << 24), 48u | (34u << 8) | (18u << 16)};                           // pixel 3: words 1, 2

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// tmaMask bit l: level l is staged with ONE TMA box per tile (maps.m[l]: rows viewed as 16-byte chunks, box 18 chunks × 38 rows; rows
// and chunks outside the image arrive as zeros, the ≤ 3 reflected rows at the top / bottom edge are then copied inside shared memory).
__global__ void __launch_bounds__(64 * BLUR_STRIPS) k_blur(ExParams p, const BlurTile *tiles, const __grid_constant__ OrbxTmaMaps maps, int tmaMask) {
    __shared__ __align__(128) uint8_t tile[(BLUR_TH + 6) * BLUR_SP];
    __shared__ __align__(8) uint64_t tileBar;
    const OrbxGeom &g = *p.g;
    const BlurTile T = tiles[blockIdx.x];
    const int b = blockIdx.y;
    const int l = T.level;
    const OrbxLevel &LV = g.lv[l];
    int pitch;
    const uint8_t *src = level_ptr(p, g, l, b, pitch);
    const int w = LV.w, h = LV.h;
    const int tx0 = T.tx * BLUR_TW, ty0 = T.ty * BLUR_TH;
    const int tid = threadIdx.y * 64 + threadIdx.x;
    const bool aligned = ((((unsigned long long)src | (unsigned)pitch) & 15ull) == 0);
    if ((tmaMask >> l) & 1) {
        if (tid == 0) {
            mbar_init(&tileBar, 1);
            mbar_expect_tx(&tileBar, (BLUR_TH + 6) * BLUR_SP);
            tma_load_4d(tile, &maps.m[l], &tileBar, 0, (tx0 - 16) >> 4, ty0 - 3, b);
        }
        __syncthreads();                       // the barrier is initialised before anyone polls it
        mbar_wait(&tileBar, 0);
        // REFLECT_101 rows above the first / below the last image row (only tiles that touch those edges)
        if (ty0 == 0 || ty0 + BLUR_TH + 3 > h) {
            for (int i = tid; i < (BLUR_TH + 6) * (BLUR_SP / 16); i += 64 * BLUR_STRIPS) {
                const int r = i / (BLUR_SP / 16), c = i - r * (BLUR_SP / 16);
                const int gy = ty0 + r - 3;
                if (gy < 0 || (gy >= h && gy <= h + 2)) {
                    const int sr = reflect101(gy, h) - (ty0 - 3);
                    *reinterpret_cast<uint4 *>(&tile[r * BLUR_SP + 16 * c]) = *reinterpret_cast<const uint4 *>(&tile[sr * BLUR_SP + 16 * c]);
                }
            }
        }
    } else
    // stage rows ty0-3 .. ty0+BLUR_TH+2 (reflected), columns tx0-16 .. tx0+BLUR_TW+15 (clipped to the row)
    {
        const int nChunks = BLUR_SP / 16;
        const int rowBytes = aligned ? min(pitch, (w + 15) & ~15) : w;   // bytes of a row that may be read
        for (int i = tid; i < (BLUR_TH + 6) * nChunks; i += 64 * BLUR_STRIPS) {
            const int r = i / nChunks, c = i - r * nChunks;
            const int gx = tx0 - 16 + 16 * c;
            const int sy = reflect101(min(ty0 + r - 3, h + 2), h);
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gx >= 0 && gx < rowBytes) {
                const uint8_t *q = src + (long long)sy * pitch + gx;
                if (aligned) {
                    v = *reinterpret_cast<const uint4 *>(q);
                } else {
                    uint32_t ww[4] = {0, 0, 0, 0};
                    for (int j = 0; j < 16; ++j)
                        if (gx + j < w) ww[j >> 2] |= (uint32_t)q[j] << (8 * (j & 3));
                    v = make_uint4(ww[0], ww[1], ww[2], ww[3]);
                }
            }
            *reinterpret_cast<uint4 *>(&tile[r * BLUR_SP + 16 * c]) = v;
        }
    }
    __syncthreads();
    // REFLECT_101 at the left / right image edge: patch the three halo columns of every staged row in place, so
    // the main loop reads plain aligned words everywhere
    {
        const bool leftTile = tx0 == 0, rightTile = (w + 2 >= tx0 - 16) && (w < tx0 + BLUR_TW + 16);
        if (leftTile || rightTile) {
            for (int i = tid; i < (BLUR_TH + 6) * 6; i += 64 * BLUR_STRIPS) {
                const int r = i / 6, k = i - r * 6;
                uint8_t *trow8 = &tile[r * BLUR_SP];
                if (k < 3) {
                    if (leftTile) trow8[16 - (k + 1)] = trow8[16 + min(k + 1, w - 1)];               // x = -(k+1) ← x = k+1
                } else if (rightTile) {
                    const int j = k - 3, xd = w + j, xs = max(w - 2 - j, 0);                        // x = w+j ← x = w-2-j
                    const int cd = xd - (tx0 - 16), cs = xs - (tx0 - 16);
                    if (cd >= 0 && cd < BLUR_SP && cs >= 0 && cs < BLUR_SP) trow8[cd] = trow8[cs];
                }
            }
        }
    }
    __syncthreads();
    const int x = tx0 + threadIdx.x * 4;
    const int y0 = ty0 + threadIdx.y * BLUR_RH;
    if (x >= w || y0 >= h) return;
    // tap words (constant bank operands of the DP4As): pixel i = byte 4+i of the 12-byte window, its taps (18,34,48,56,48,34,18)
    // cover bytes 1+i .. 7+i, i.e. they slide over the three aligned words instead of the bytes being realigned
    const uint32_t K0A = kBlurTaps[0], K0B = kBlurTaps[1], K1A = kBlurTaps[2], K1B = kBlurTaps[3], K1C = kBlurTaps[4];
    const uint32_t K2A = kBlurTaps[5], K2B = kBlurTaps[6], K2C = kBlurTaps[7], KLO = kBlurTaps[8], KHI = kBlurTaps[9];
    const int dpitch = LV.pitch;
    unsigned long long dst = (unsigned long long)__cvta_generic_to_global(p.blur + (long long)b * g.frameBytes + LV.off + (long long)y0 * dpitch + x);
    asm volatile("" : "+l"(dst));              // one 64-bit register pair from here on
    // the 12-byte window x-4 .. x+7 of a staged row = three aligned words (tile column 0 = image column tx0-16)
    const uint32_t *trow = reinterpret_cast<const uint32_t *>(tile) + threadIdx.y * BLUR_RH * (BLUR_SP / 4) + ((x - 4 - (tx0 - 16)) >> 2);
    // every horizontal sum carries +128: the vertical taps add up to 256, so the vertical sum arrives with its rounding term
    // 32768 already in place (sums stay below 2^16 horizontally and 2^24 vertically)
    const uint32_t kRound = 128u;
    const int nOut = min(BLUR_RH, h - y0);
    uint32_t ring[7][4];
#pragma unroll
    for (int r = 0; r < BLUR_RH + 6; ++r) {
        const uint32_t w0 = trow[r * (BLUR_SP / 4)], w1 = trow[r * (BLUR_SP / 4) + 1], w2 = trow[r * (BLUR_SP / 4) + 2];
        // pixel i sits at byte 4+i of {w0,w1,w2}; taps are bytes 1+i .. 7+i
        uint32_t *hr = ring[r % 7];
        // the taps slide over the three aligned words instead of the bytes being realigned: 10 DP4A, no PRMT
        hr[0] = __dp4a(w0, K0A, __dp4a(w1, K0B, kRound));
        hr[1] = __dp4a(w0, K1A, __dp4a(w1, K1B, __dp4a(w2, K1C, kRound)));
        hr[2] = __dp4a(w0, K2A, __dp4a(w1, K2B, __dp4a(w2, K2C, kRound)));
        hr[3] = __dp4a(w1, KLO, __dp4a(w2, KHI, kRound));
        if (r >= 6) {
            uint32_t v[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                v[i] = 18u * (ring[(r - 6) % 7][i] + ring[r % 7][i]) + 34u * (ring[(r - 5) % 7][i] + ring[(r - 1) % 7][i]) +
                       48u * (ring[(r - 4) % 7][i] + ring[(r - 2) % 7][i]) + 56u * ring[(r - 3) % 7][i];
            // (sum + 32768) >> 16 is byte 2 of each sum: three PRMT gather the four output bytes
            const uint32_t out = __byte_perm(__byte_perm(v[0], v[1], 0x0062u), __byte_perm(v[2], v[3], 0x0062u), 0x5410u);
            if (r - 6 < nOut) {
                const unsigned long long q = dst + (unsigned long long)(uint32_t)dpitch * (uint32_t)(r - 6);   // one wide multiply-add
                asm volatile("st.global.u32 [%0], %1;" ::"l"(q), "r"(out) : "memory");
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4 + K6: orientation (IC_Angle :76-103, fastAtan2 A5) and rBRIEF (:107-146).  One warp per keypoint.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale),
                p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = (float)2.2204460492503131e-16;
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

template <int WPB>
__global__ void __launch_bounds__(WPB * 32, 2048 / (WPB * 32)) k_orient_desc(ExParams p) {
    const OrbxGeom &g = *p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * WPB + warp, b = blockIdx.y;
    if (i >= p.workCnt[b]) return;
    const OrbxWork wk = p.work[(long long)b * g.selTotal + i];
    if (wk.out < 0) return;
    int pitch;
    const uint8_t *img = level_ptr(p, g, wk.level, b, pitch);
    // moments over the 31-row circular patch: lane = column u+15, lane 31 idle.  The row half-widths are the
    // reference's umax table for HALF_PATCH_SIZE=15 (checked against the ctor maths in orbx_create).
    constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int u = lane - ORBX_HALF_PATCH;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const uint8_t *pr = img + (long long)(wk.cy - ORBX_HALF_PATCH) * pitch + wk.cx + u;
        const int au = u < 0 ? -u : u;
        int colSum = 0;   // Σ val over the rows this column belongs to, and Σ v·val
#pragma unroll
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; ++v, pr += pitch) {
            if (au <= UMAX[v < 0 ? -v : v]) {
                const int val = *pr;
                colSum += val;
                m01 += v * val;
            }
        }
        m10 = u * colSum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    // rBRIEF on the blurred level: lane computes descriptor byte `lane` (tests 8·lane … 8·lane+7)
    const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    const float ca = orbx_libm::cosf_glibc(rad), sa = orbx_libm::sinf_glibc(rad);
    const OrbxLevel &LV = g.lv[wk.level];
    const uint8_t *bl = p.blur + (long long)b * g.frameBytes + LV.off + (long long)wk.cy * LV.pitch + wk.cx;
    const int bp = LV.pitch;
    const int4 *pat4 = reinterpret_cast<const int4 *>(p.pattern) + lane * 2;   // 8 tests × (x0, y0, x1, y1) int8
    const int4 q0 = pat4[0], q1 = pat4[1];
    const int words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    int byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int wv = words[j];
        const float tx0 = (float)(signed char)(wv & 0xff), ty0 = (float)(signed char)((wv >> 8) & 0xff);
        const float tx1 = (float)(signed char)((wv >> 16) & 0xff), ty1 = (float)(signed char)((wv >> 24) & 0xff);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(tx0, sa), __fmul_rn(ty0, ca)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(tx0, ca), __fmul_rn(ty0, sa)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(tx1, sa), __fmul_rn(ty1, ca)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(tx1, ca), __fmul_rn(ty1, sa)));
        const int v0 = bl[r0 * bp + c0], v1 = bl[r1 * bp + c1];
        byte |= (v0 < v1) << j;
    }
    // gather 4 bytes per lane group and store 8 words
    uint32_t word = (uint32_t)byte;
    word |= __shfl_down_sync(0xffffffffu, (uint32_t)byte, 1) << 8;
    word |= __shfl_down_sync(0xffffffffu, (uint32_t)byte, 2) << 16;
    word |= __shfl_down_sync(0xffffffffu, (uint32_t)byte, 3) << 24;
    uint8_t *d = p.desc + ((long long)b * p.cap + wk.out) * 32;
    if ((lane & 3) == 0) reinterpret_cast<uint32_t *>(d)[lane >> 2] = word;
    if (lane == 0) p.kps[(long long)b * p.cap + wk.out].angle = angle;
}

// TMA form of the same kernel: the two patches a keypoint needs — 31 rows of the unblurred level for the moments, 37 rows of the
// blurred level for the rotated pattern (|offset| <= 18) — arrive as two box loads per warp instead of ~1260 scattered byte
// loads through L1; all further reads are shared-memory bytes.  A box starts on a 16-byte boundary of the image row, so the
// patch sits at a per-keypoint byte offset (0..15) inside its box.  Same arithmetic, same results.
#define OD_AW 48     // unblurred box: 15 (alignment slack) + 31 columns, rounded to 16
#define OD_AH 31
#define OD_BW 64     // blurred box: 15 + 37 columns, rounded to 16
#define OD_BH 37
#define OD_SLOT 4096 // per-warp shared memory: [A box 1488][pad to 1536][B box 2368][pad][mbarrier]
template <int WPB>
__global__ void __launch_bounds__(WPB * 32, 2048 / (WPB * 32)) k_orient_desc_tma(ExParams p, const __grid_constant__ OrbxTmaMaps pyrMaps,
                                                                                   const __grid_constant__ OrbxTmaMaps blurMaps) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const OrbxGeom &g = *p.g;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * WPB + warp, b = blockIdx.y;
    if (i >= p.workCnt[b]) return;
    const OrbxWork wk = p.work[(long long)b * g.selTotal + i];
    if (wk.out < 0) return;
    uint8_t *A = smem_raw + (size_t)warp * OD_SLOT, *B = A + 1536;
    uint64_t *bar = reinterpret_cast<uint64_t *>(A + OD_SLOT - 8);
    const int xa = (wk.cx - ORBX_HALF_PATCH) & ~15, xb = (wk.cx - 18) & ~15;
    if (lane == 0) {
        mbar_init(bar, 1);
        mbar_expect_tx(bar, OD_AW * OD_AH + OD_BW * OD_BH);
        tma_load_3d(A, &pyrMaps.m[wk.level], bar, xa, wk.cy - ORBX_HALF_PATCH, b);
        tma_load_3d(B, &blurMaps.m[wk.level], bar, xb, wk.cy - 18, b);
    }
    // pattern words of this lane's descriptor byte (while the boxes are in flight)
    const int4 *pat4 = reinterpret_cast<const int4 *>(p.pattern) + lane * 2;   // 8 tests × (x0, y0, x1, y1) int8
    const int4 q0 = pat4[0], q1 = pat4[1];
    __syncwarp();
    mbar_wait(bar, 0);
    constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    const int u = lane - ORBX_HALF_PATCH;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const uint8_t *pr = A + (wk.cx - ORBX_HALF_PATCH - xa) + lane;      // row v = -15 of column u
        const int au = u < 0 ? -u : u;
        int colSum = 0;
#pragma unroll
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; ++v, pr += OD_AW) {
            if (au <= UMAX[v < 0 ? -v : v]) {
                const int val = *pr;
                colSum += val;
                m01 += v * val;
            }
        }
        m10 = u * colSum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);
    const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);
    const float rad = __fmul_rn(angle, factorPI);
    const float ca = orbx_libm::cosf_glibc(rad), sa = orbx_libm::sinf_glibc(rad);
    const uint8_t *bl = B + 18 * OD_BW + (wk.cx - xb);                      // the keypoint's pixel inside the blurred box
    const int words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    int byte = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int wv = words[j];
        const float tx0 = (float)(signed char)(wv & 0xff), ty0 = (float)(signed char)((wv >> 8) & 0xff);
        const float tx1 = (float)(signed char)((wv >> 16) & 0xff), ty1 = (float)(signed char)((wv >> 24) & 0xff);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(tx0, sa), __fmul_rn(ty0, ca)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(tx0, ca), __fmul_rn(ty0, sa)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(tx1, sa), __fmul_rn(ty1, ca)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(tx1, ca), __fmul_rn(ty1, sa)));
        const int v0 = bl[r0 * OD_BW + c0], v1 = bl[r1 * OD_BW + c1];
        byte |= (v0 < v1) << j;
    }
    uint32_t word = (uint32_t)byte;
    word |= __shfl_down_sync(0xffffffffu, (uint32_t)byte, 1) << 8;
    word |= __shfl_down_sync(0xffffffffu, (uint32_t)byte, 2) << 16;
    word |= __shfl_down_sync(0xffffffffu, (uint32_t)byte, 3) << 24;
    uint8_t *d = p.desc + ((long long)b * p.cap + wk.out) * 32;
    if ((lane & 3) == 0) reinterpret_cast<uint32_t *>(d)[lane >> 2] = word;
    if (lane == 0) p.kps[(long long)b * p.cap + wk.out].angle = angle;
}

// packed host layout (rows of `cols` bytes, frames back to back) → pitched level-0 planes of the internal pyramid
__global__ void k_repitch(const uint8_t *__restrict__ packed, int rows, int cols, uint8_t *__restrict__ dst, long long frameBytes, int pitch) {
    const int b = blockIdx.z, y = blockIdx.y;
    const uint8_t *s = packed + ((long long)b * rows + y) * cols;
    uint8_t *d = dst + (long long)b * frameBytes + (long long)y * pitch;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < cols; x += gridDim.x * blockDim.x) d[x] = s[x];
}

// debug kernels ----------------------------------------------------------------------------------
__global__ void k_dbg_sincos(const float *a, int n, float *s, float *c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { s[i] = orbx_libm::sinf_glibc(a[i]); c[i] = orbx_libm::cosf_glibc(a[i]); }
}
__global__ void k_dbg_atan2(const float *y, const float *x, int n, float *o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = fast_atan2_deg(y[i], x[i]);
}
// same two-phase evaluation as k_quadtree_hist: serial partitioning + stable rank sort (single thread here)
__global__ void k_dbg_sort(orbx_sort::elem_t *a, orbx_sort::elem_t *tmp, int n) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        orbx_sort::introsort_loop_only(a, n);
        for (int i = 0; i < n; ++i) {
            const unsigned long long k = a[i] >> orbx_sort::kPayloadBits;
            int rank = 0;
            for (int j = 0; j < n; ++j) {
                const unsigned long long kj = a[j] >> orbx_sort::kPayloadBits;
                rank += (kj < k) || (kj == k && j < i);
            }
            tmp[rank] = a[i];
        }
        for (int i = 0; i < n; ++i) a[i] = tmp[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Classical rectified-stereo association (slot = Frame::ComputeStereoMatches, src/Frame.cc:813-915).
// This tree replaced the function's matcher by LightGlue, so the algorithm is the restated upstream one (see oracle/
// orb_oracle.cpp: orc_stereo_rowband — parity unpinned): row band of ±2·scale around every right keypoint, Hamming search
// within one octave and the disparity range, 11×11 SAD over ±5 px on the keypoint's pyramid level with a parabola fit,
// disparity gate, median cut at 1.5·1.4·median.  One warp per left keypoint; the band test replaces the per-row index lists
// (candidates are visited in ascending right index, which is the lists' insertion order, so ties resolve the same way).
// ------------------------------------------------------------------------------------------------
struct SmLevels {
    const uint8_t *L[ORBX_MAX_LEVELS], *R[ORBX_MAX_LEVELS];
    int pitchL[ORBX_MAX_LEVELS], pitchR[ORBX_MAX_LEVELS], w[ORBX_MAX_LEVELS], h[ORBX_MAX_LEVELS];
    float sf[ORBX_MAX_LEVELS], inv[ORBX_MAX_LEVELS];
    int nlevels, nRows;
};
// per right keypoint: {first band row, last band row, octave, pt.x bits}
__global__ void k_sm_prep(const orbx_keypoint *kR, int nR, SmLevels lv, int4 *prep) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nR) return;
    const orbx_keypoint k = kR[i];
    const int oct = min(max(k.octave, 0), lv.nlevels - 1);
    const float r = __fmul_rn(2.0f, lv.sf[oct]);
    prep[i] = make_int4((int)floorf(__fsub_rn(k.y, r)), (int)ceilf(__fadd_rn(k.y, r)), k.octave, __float_as_int(k.x));
}
template <int WPB>
__global__ void __launch_bounds__(WPB * 32) k_sm_match(const orbx_keypoint *kL, const uint4 *dL, int nL, const int4 *prep, const uint4 *dR, int nR, SmLevels lv,
                                                       float mbf, float mb, float *uRight, float *depth, int *sadOut) {
    const int iL = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (iL >= nL) return;
    const orbx_keypoint k = kL[iL];
    float ur = -1.f, dp = -1.f;
    int sadBest = -1;
    const int levelL = k.octave;
    const float vL = k.y, uL = k.x;
    const int row = (int)vL;
    const float maxD = __fdiv_rn(mbf, mb);
    const float minU = __fsub_rn(uL, maxD), maxU = uL;      // uL - minD with minD = 0
    const unsigned long long NONE = ~0ull;
    unsigned long long best = NONE;
    if (row >= 0 && row < lv.nRows && !(maxU < 0.f) && levelL >= 0 && levelL < lv.nlevels) {
        const uint4 a0 = dL[2 * (long long)iL], a1 = dL[2 * (long long)iL + 1];
        for (int iR = lane; iR < nR; iR += 32) {
            const int4 c = prep[iR];
            if (row < c.x || row > c.y) continue;                          // not in this keypoint's row band
            if (c.z < levelL - 1 || c.z > levelL + 1) continue;
            const float uR = __int_as_float(c.w);
            if (!(uR >= minU && uR <= maxU)) continue;
            const uint4 b0 = dR[2 * (long long)iR], b1 = dR[2 * (long long)iR + 1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
                          __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            if (d < 100) {                                                 // bestDist starts at TH_HIGH, strict '<'
                const unsigned long long key = ((unsigned long long)(unsigned)d << 32) | (unsigned)iR;
                best = key < best ? key : best;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, best, o);
        best = ob < best ? ob : best;
    }
    if (best != NONE && (int)(best >> 32) < (100 + 50) / 2) {              // thOrbDist
        const int bestIdxR = (int)(best & 0xffffffffu);
        const float uR0 = __int_as_float(prep[bestIdxR].w);
        const float scaleFactor = lv.inv[levelL];
        const int cu = (int)roundf(__fmul_rn(uL, scaleFactor)), cv = (int)roundf(__fmul_rn(vL, scaleFactor)), cr = (int)roundf(__fmul_rn(uR0, scaleFactor));
        const int w = 5, Ls = 5;
        const int wl = lv.w[levelL], hl = lv.h[levelL];
        // iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1; the other tests keep the windows inside the levels
        if (!(cr < 0 || cr + Ls + w + 1 >= wl) && cv - w >= 0 && cv + w < hl && cu - w >= 0 && cu + w < wl && cr - Ls - w >= 0) {
            int acc[11];
#pragma unroll
            for (int i = 0; i < 11; ++i) acc[i] = 0;
            const uint8_t *PL = lv.L[levelL], *PR = lv.R[levelL];
            const int pl = lv.pitchL[levelL], pr = lv.pitchR[levelL];
            for (int p = lane; p < 121; p += 32) {
                const int yy = p / 11, xx = p - yy * 11;
                const int a = PL[(long long)(cv - w + yy) * pl + (cu - w + xx)];
                const uint8_t *q = PR + (long long)(cv - w + yy) * pr + (cr - Ls - w + xx);
#pragma unroll
                for (int i = 0; i < 11; ++i) acc[i] += abs(a - (int)q[i]);
            }
#pragma unroll
            for (int i = 0; i < 11; ++i)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
            int bestSad = INT_MAX, bestinc = 0;
#pragma unroll
            for (int i = 0; i < 11; ++i)
                if (acc[i] < bestSad) { bestSad = acc[i]; bestinc = i - Ls; }
            if (bestinc != -Ls && bestinc != Ls) {
                float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int i = 0; i < 11; ++i) {           // static indexing keeps acc[] in registers
                    if (i == Ls + bestinc - 1) d1 = (float)acc[i];
                    if (i == Ls + bestinc) d2 = (float)acc[i];
                    if (i == Ls + bestinc + 1) d3 = (float)acc[i];
                }
                const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), den);
                if (!(deltaR < -1.f || deltaR > 1.f)) {                    // (a NaN from 0/0 passes both tests, as in the scalar code)
                    float bestuR = __fmul_rn(lv.sf[levelL], __fadd_rn(__fadd_rn((float)cr, (float)bestinc), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= 0.f && disparity < maxD) {
                        if (disparity <= 0.f) { disparity = 0.01f; bestuR = __fsub_rn(uL, 0.01f); }
                        dp = __fdiv_rn(mbf, disparity);
                        ur = bestuR;
                        sadBest = bestSad;
                    }
                }
            }
        }
    }
    if (lane == 0) { uRight[iL] = ur; depth[iL] = dp; sadOut[iL] = sadBest; }
}
// median cut: the reference sorts (SAD, index) pairs and takes element [size/2]; one block, two 256-bin histograms (SAD < 2^16)
__global__ void __launch_bounds__(256) k_sm_tail(const int *sad, int nL, float *uRight, float *depth, int *keptOut) {
    __shared__ int hist[256];
    __shared__ int s_total, s_hi, s_before, s_med, s_dropped;
    const int tid = threadIdx.x;
    hist[tid] = 0;
    if (tid == 0) { s_total = 0; s_dropped = 0; }
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nL; i += 256) {
        const int v = sad[i];
        if (v >= 0) { atomicAdd(&hist[min(v >> 8, 255)], 1); ++mine; }
    }
    atomicAdd(&s_total, mine);
    __syncthreads();
    const int total = s_total;
    if (total == 0) { if (tid == 0) *keptOut = 0; return; }
    const int kth = total / 2;
    if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < 256; ++b) { if (acc + hist[b] > kth) break; acc += hist[b]; }
        s_hi = b; s_before = acc;
    }
    __syncthreads();
    const int hi = s_hi;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nL; i += 256) {
        const int v = sad[i];
        if (v >= 0 && min(v >> 8, 255) == hi) atomicAdd(&hist[v & 255], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = s_before, b = 0;
        for (; b < 256; ++b) { if (acc + hist[b] > kth) break; acc += hist[b]; }
        s_med = (hi << 8) | b;
    }
    __syncthreads();
    const float thDist = __fmul_rn(__fmul_rn(1.5f, 1.4f), (float)s_med);
    int dropped = 0;
    for (int i = tid; i < nL; i += 256) {
        const int v = sad[i];
        if (v >= 0 && !((float)v < thDist)) { uRight[i] = -1.f; depth[i] = -1.f; ++dropped; }
    }
    atomicAdd(&s_dropped, dropped);
    __syncthreads();
    if (tid == 0) *keptOut = total - s_dropped;
}

thread_local std::string tl_error;

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
#define ORBX_MAX_CHUNKS 40
#define ORBX_MAX_SIDE 4
struct orbx_extractor {
    int device = 0;
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
    double scaleFactor = 1.0;
    int maxW = 0, maxH = 0, maxBatch = 0;
    float sf[ORBX_MAX_LEVELS], inv[ORBX_MAX_LEVELS], sig2[ORBX_MAX_LEVELS], invsig2[ORBX_MAX_LEVELS];
    int quota[ORBX_MAX_LEVELS];
    int umax[ORBX_HALF_PATCH + 1];
    cudaStream_t stream = nullptr, sH2D = nullptr, sD2H = nullptr, sSide[ORBX_MAX_SIDE] = {};
    int nSide = 2, nSub = 0, nSteady = 0;     // side streams in use, sub-batches of the device path, steady chunks of the host path
    cudaEvent_t evFork = nullptr, evJoin[ORBX_MAX_SIDE] = {};
    // the blur only needs the pyramid: it runs on a partner stream next to FAST + quadtree (index 0: main stream, 1+i: side i)
    cudaStream_t sAux[ORBX_MAX_SIDE + 1] = {};
    cudaEvent_t evPyrDone[ORBX_MAX_SIDE + 1] = {}, evBlurDone[ORBX_MAX_SIDE + 1] = {};
    bool overlapBlur = true;
    cudaEvent_t evIn[ORBX_MAX_CHUNKS] = {}, evOut[ORBX_MAX_CHUNKS] = {};
    std::string err;
    long long launches = 0;

    // geometry of the current image size
    OrbxGeom geom;
    int curRows = -1, curCols = -1;
    int curLap0 = 0, curLap1 = 0;
    std::vector<int> curRects;
    bool geomDirty = true;
    int h_tabXOff[ORBX_MAX_LEVELS] = {0}, h_tabYOff[ORBX_MAX_LEVELS] = {0}, h_tileXOff[ORBX_MAX_LEVELS] = {0}, h_tileYOff[ORBX_MAX_LEVELS] = {0};
    std::vector<OrbxCell> h_cells;
    std::vector<unsigned short> h_pathLut;   // quadtree path tables of all levels
    unsigned short *d_pathLut = nullptr; size_t pathLutCap = 0;
    std::vector<BlurTile> h_tiles;
    int maxSlotCap = 0, nodeCapMax = 0, maxCellsLevel = 0, maxIni = 1;
    bool useHistQuadtree = true;
    int dbgCalls = 0;
    // developer knobs, read from the environment once in orbx_create
    bool dbgChunks = false, dbgSkipH2D = false, dbgSkipD2H = false;
    std::string chunkPlan;
    bool fastV1 = false;            // ORBX_FAST_V1: single-phase FAST kernel (every pixel gets the exact measure)
    unsigned *d_hist = nullptr; size_t histCap = 0;
    unsigned short *d_finalPos = nullptr; size_t finalPosCap = 0;
    unsigned *d_best = nullptr; size_t bestCap = 0;
    int *d_cellPrefix = nullptr; size_t cellPrefixCap = 0;
    int *d_deep = nullptr;
    int *d_dense = nullptr;        // per run (indexed by its first frame): cells the two-phase FAST kernel handed to the single-phase one
    int2 *d_denseList = nullptr; size_t denseListCap = 0;
    std::vector<int> denseSlots;   // counters the last public call used
    int nSM = 148;
    uint8_t *d_stage = nullptr; size_t stageCap = 0;   // packed H2D staging when the level-0 pitch is padded
    int lastBatch = 0;
    bool lastIn0Internal = true;
    // single-frame (latency) path of orbx_extract: pinned staging both ways, one device block for all outputs, the whole call
    // (copy in, kernels, copy out) captured once per (rows, cols, cap) as a CUDA graph
    uint8_t *h_in = nullptr; size_t hInCap = 0;
    uint8_t *d_single = nullptr, *h_single = nullptr; size_t singleCap = 0;
    cudaGraphExec_t oneExec = nullptr;
    int oneRows = -1, oneCols = -1, oneCap = -1, oneWarm = 0;
    long long oneLaunches = 0;
    bool useGraph = true;
    uint8_t *d_sm = nullptr; size_t smCap = 0;     // scratch of orbx_stereo_matches
    bool useTmaPyr = true;          // ORBX_NO_TMA: stage tiles with ordinary loads (same results; cross-check and fallback)

    // device buffers (sized for maxW × maxH × maxBatch at create)
    OrbxGeom *d_geom = nullptr;
    OrbxCell *d_cells = nullptr; size_t cellsCap = 0;
    BlurTile *d_tiles = nullptr; size_t tilesCap = 0;
    int2 *d_tabX = nullptr, *d_tabY = nullptr; size_t tabXCap = 0, tabYCap = 0;
    int8_t *d_pattern = nullptr;
    uint8_t *d_pyr = nullptr, *d_blur = nullptr; size_t pyrCap = 0;
    uint32_t *d_slots = nullptr; uint32_t *d_ptNode = nullptr; float2 *d_ptXY = nullptr; size_t slotsCap = 0;
    int *d_cellCnt = nullptr; size_t cellCntCap = 0;
    float4 *d_sel = nullptr; OrbxWork *d_work = nullptr; size_t selCap = 0;
    int *d_selCnt = nullptr, *d_workCnt = nullptr;
    // staging for the host-buffer entry points
    orbx_keypoint *d_kps = nullptr; uint8_t *d_desc = nullptr; size_t outCap = 0;
    int *d_nOut = nullptr, *d_mono = nullptr;
    int *h_nOut = nullptr, *h_mono = nullptr;  // pinned
    // optional per-stage timing (bench roofline): events at the 6 stage boundaries
    bool profiling = false, evPending = false;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double stageMs[6] = {0, 0, 0, 0, 0, 0};
    long long stageCalls = 0;
};

namespace {

#define CUDA_TRY(ex, call)                                                                              \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            (ex)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return ORBX_ERR_CUDA;                                                                       \
        }                                                                                               \
    } while (0)

inline int cv_round_f(float v) { return (int)lrintf(v); }

template <class T>
int ensure(orbx_extractor *ex, T *&ptr, size_t &cap, size_t need) {
    if (need <= cap && ptr) return ORBX_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    CUDA_TRY(ex, cudaMalloc((void **)&ptr, need * sizeof(T)));
    cap = need;
    return ORBX_OK;
}

// Level sizes, cell grid and quadtree roots for a rows×cols image — the host restatement of
// :1213-1214 (level sizes), :789-822 (cell grid) and :558-560 (roots), fp32 where the reference is.
int build_geometry(orbx_extractor *ex, int rows, int cols) {
    OrbxGeom &G = ex->geom;
    memset(&G, 0, sizeof(G));
    G.nlevels = ex->nlevels; G.rows = rows; G.cols = cols;
    G.iniTh = ex->iniTh; G.minTh = ex->minTh; G.lowTh = std::min(ex->iniTh, ex->minTh);
    memcpy(G.umax, ex->umax, sizeof(G.umax));
    ex->h_cells.clear();
    ex->h_tiles.clear();
    ex->h_pathLut.clear();
    long long off = 0, slot = 0;
    int cellBase = 0, selBase = 0;
    ex->maxSlotCap = 1; ex->nodeCapMax = 8; ex->maxCellsLevel = 1; ex->maxIni = 1;
    G.maxCw = 7; G.maxCh = 7;
    for (int l = 0; l < ex->nlevels; ++l) {
        OrbxLevel &V = G.lv[l];
        V.sf = ex->sf[l]; V.inv = ex->inv[l]; V.quota = ex->quota[l];
        V.w = cv_round_f((float)cols * V.inv);
        V.h = cv_round_f((float)rows * V.inv);
        if (V.w <= 2 * ORBX_BORDER + 6 || V.h <= 2 * ORBX_BORDER + 6 || V.w > 32000 || V.h > 32000) {
            ex->err = "image too small (or too large) for the pyramid: level " + std::to_string(l) + " is " +
                      std::to_string(V.w) + "x" + std::to_string(V.h);
            return ORBX_ERR_GEOMETRY;
        }
        V.pitch = orbx_align_up(V.w, 128);
        V.off = off;
        off += orbx_align_up_ll((long long)V.pitch * V.h, 256);
        V.patch_size = (int)((float)ORBX_PATCH * V.sf);
        V.maxBX = V.w - ORBX_EDGE + 3;
        V.maxBY = V.h - ORBX_EDGE + 3;
        const float width = (float)(V.maxBX - ORBX_BORDER), height = (float)(V.maxBY - ORBX_BORDER);
        V.nCols = (int)(width / 35.f);
        V.nRows = (int)(height / 35.f);
        V.wCell = V.nCols > 0 ? (int)ceilf(width / (float)V.nCols) : 0;
        V.hCell = V.nRows > 0 ? (int)ceilf(height / (float)V.nRows) : 0;
        // rows/cols that survive the skip tests (:810, :819; quirk Q2) form a prefix
        V.nRowsOK = 0;
        for (int i = 0; i < V.nRows; ++i)
            if (!((float)(ORBX_BORDER + i * V.hCell) >= (float)(V.maxBY - 3))) V.nRowsOK = i + 1; else break;
        V.nColsOK = 0;
        for (int j = 0; j < V.nCols; ++j)
            if (!((float)(ORBX_BORDER + j * V.wCell) >= (float)(V.maxBX - 6))) V.nColsOK = j + 1; else break;
        V.cellBase = cellBase;
        V.nCells = V.nRowsOK * V.nColsOK;
        V.slotCap = ((V.wCell + 1) / 2) * ((V.hCell + 1) / 2);  // strict 3×3 NMS ⇒ ≤ ⌈w/2⌉·⌈h/2⌉ survivors
        V.slotBase = slot;
        for (int i = 0; i < V.nRowsOK; ++i)
            for (int j = 0; j < V.nColsOK; ++j) {
                OrbxCell c;
                const int iniX = ORBX_BORDER + j * V.wCell, iniY = ORBX_BORDER + i * V.hCell;
                const int maxX = std::min(iniX + V.wCell + 6, V.maxBX), maxY = std::min(iniY + V.hCell + 6, V.maxBY);
                c.level = (short)l; c.x0 = (short)iniX; c.y0 = (short)iniY;
                c.cw = (short)(maxX - iniX); c.ch = (short)(maxY - iniY);
                c.cx = (short)j; c.cy = (short)i;
                c.seq = i * V.nColsOK + j;
                c.slot = slot;
                slot += V.slotCap;
                ex->h_cells.push_back(c);
                G.maxCw = std::max(G.maxCw, (int)c.cw);
                G.maxCh = std::max(G.maxCh, (int)c.ch);
            }
        if (V.wCell + 6 > 255 || V.hCell + 6 > 255) { ex->err = "cell larger than 255 px"; return ORBX_ERR_GEOMETRY; }
        if (fast_smem_layout(V.wCell + 6, V.hCell + 6, 1).roiPitch * (V.hCell + 6) > 16383) {   // FAST queue entries hold a 14-bit ROI offset
            ex->err = "cell ROI larger than 16383 bytes";
            return ORBX_ERR_GEOMETRY;
        }
        cellBase += V.nCells;
        ex->maxSlotCap = std::max(ex->maxSlotCap, V.slotCap);
        ex->maxCellsLevel = std::max(ex->maxCellsLevel, V.nCells);
        // quadtree roots (:558-560)
        const int rw = V.maxBX - ORBX_BORDER, rh = V.maxBY - ORBX_BORDER;
        V.nIni = (int)roundf((float)rw / (float)rh);
        if (V.nIni < 1) { ex->err = "image too tall: quadtree would have no root node (reference faults)"; return ORBX_ERR_GEOMETRY; }
        V.hX = (float)rw / (float)V.nIni;
        ex->maxIni = std::max(ex->maxIni, std::min(V.nIni, QT_MAX_INI));
        // quadtree path tables (k_qt_classify): the six split decisions per integer x (per root) and per integer y
        V.lutW = V.lutH = V.lutX = V.lutY = 0;
        if (V.nIni <= QT_MAX_INI && rw > 0 && rh > 0) {
            V.lutW = rw; V.lutH = rh;
            V.lutX = (int)ex->h_pathLut.size();
            for (int bin = 0; bin < V.nIni; ++bin)
                for (int xi = 0; xi < rw; ++xi) {
                    short4 bx;
                    bx.x = (short)(int)(V.hX * (float)bin); bx.y = (short)(int)(V.hX * (float)(bin + 1)); bx.z = 0; bx.w = (short)rh;
                    unsigned code = 0;
                    for (int d = 0; d < QT_DMAX; ++d) { const int q = qt_quadrant(bx, (float)xi, 0.f) & 1; bx = qt_child_box(bx, q); code = code * 4u + (unsigned)q; }
                    ex->h_pathLut.push_back((unsigned short)code);
                }
            V.lutY = (int)ex->h_pathLut.size();
            for (int yi = 0; yi < rh; ++yi) {
                short4 bx;
                bx.x = 0; bx.y = (short)rw; bx.z = 0; bx.w = (short)rh;
                unsigned code = 0;
                for (int d = 0; d < QT_DMAX; ++d) { const int q = qt_quadrant(bx, 0.f, (float)yi) & 2; bx = qt_child_box(bx, q); code = code * 4u + (unsigned)q; }
                ex->h_pathLut.push_back((unsigned short)code);
            }
        }
        V.nodeCap = std::max(4 * V.nIni, V.quota + 4) + 4;
        if (V.nodeCap > 60000) { ex->err = "nfeatures too large (quadtree node index is 16-bit)"; return ORBX_ERR_ARG; }
        ex->nodeCapMax = std::max(ex->nodeCapMax, V.nodeCap);
        V.selBase = selBase;
        V.selCap = V.nodeCap;
        selBase += V.selCap;
        for (int ty = 0; ty < (V.h + BLUR_TH - 1) / BLUR_TH; ++ty)
            for (int tx = 0; tx < (V.w + BLUR_TW - 1) / BLUR_TW; ++tx) {
                BlurTile t; t.level = (short)l; t.tx = (short)tx; t.ty = (short)ty; t.pad = 0;
                ex->h_tiles.push_back(t);
            }
    }
    G.frameBytes = off;
    G.slotsTotal = slot;
    G.nCellsTotal = cellBase;
    G.selTotal = selBase;
    return ORBX_OK;
}

int upload_tables(orbx_extractor *ex) {
    OrbxGeom &G = ex->geom;
    // resize tables: identical arithmetic to cv::resize's coefficient set-up
    std::vector<int2> tx, ty;
    std::vector<int> txo(ORBX_MAX_LEVELS, 0), tyo(ORBX_MAX_LEVELS, 0), tileXo(ORBX_MAX_LEVELS, 0), tileYo(ORBX_MAX_LEVELS, 0);
    for (int l = 1; l < G.nlevels; ++l) {
        const int sw = G.lv[l - 1].w, sh = G.lv[l - 1].h, dw = G.lv[l].w, dh = G.lv[l].h;
        txo[l] = (int)tx.size();
        tyo[l] = (int)ty.size();
        const double kx = (double)sw / dw, ky = (double)sh / dh;
        for (int d = 0; d < dw; ++d) {
            float f = (float)((d + 0.5) * kx - 0.5);
            int s = (int)floorf(f);
            f -= s;
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= sw - 1) { s = sw - 1; f = 0.f; }
            const int a0 = (short)cv_round_f((1.f - f) * 2048.f), a1 = (short)cv_round_f(f * 2048.f);
            tx.push_back(make_int2(s, (a0 & 0xffff) | (a1 << 16)));
        }
        for (int pad = 0; pad < 4; ++pad) tx.push_back(make_int2(0, 0));  // 4-pixel threads may peek past the end
        for (int d = 0; d < dh; ++d) {
            float f = (float)((d + 0.5) * ky - 0.5);
            int s = (int)floorf(f);
            f -= s;
            const int b0 = (short)cv_round_f((1.f - f) * 2048.f), b1 = (short)cv_round_f(f * 2048.f);
            const int i0 = std::min(std::max(s, 0), sh - 1), i1 = std::min(std::max(s + 1, 0), sh - 1);   // rows are clamped, weights are not (A1)
            ty.push_back(make_int2(i0 | (i1 << 16), (b0 & 0xffff) | (b1 << 16)));
        }
        // per-tile geometry of k_pyr_level (what its prologue would otherwise derive from the tables)
        const int tilesX = (dw + PYR_TW - 1) / PYR_TW, tilesY = (dh + PYR_TH - 1) / PYR_TH;
        const int srcRows = (int)ceil((PYR_TH - 1) * (double)sh / dh) + 4;
        const int srcPitch = 16 * ((int)ceil(((PYR_TW - 1) * (double)sw / dw + 18.0) / 16.0) + 1);
        tileXo[l] = (int)tx.size();
        for (int t = 0; t < tilesX; ++t) {
            const int x0 = t * PYR_TW, xLast = std::min(x0 + PYR_TW, dw) - 1;
            const int c0a = tx[txo[l] + x0].x & ~15;
            const int cLast = std::min(tx[txo[l] + xLast].x + 1, sw - 1);
            const int nChunks = std::min((cLast - c0a) / 16 + 1, srcPitch / 16);
            tx.push_back(make_int2(c0a | (nChunks << 16), (65536 + nChunks - 1) / nChunks));
        }
        tileYo[l] = (int)ty.size();
        for (int t = 0; t < tilesY; ++t) {
            const int y0 = t * PYR_TH, yLast = std::min(y0 + PYR_TH, dh) - 1;
            const int r0 = ty[tyo[l] + y0].x & 0xffff;
            const int r1 = (ty[tyo[l] + yLast].x >> 16) & 0xffff;
            ty.push_back(make_int2(r0, std::min(r1 - r0 + 1, srcRows)));
        }
    }
    for (int l = 0; l < ORBX_MAX_LEVELS; ++l) { ex->h_tabXOff[l] = txo[l]; ex->h_tabYOff[l] = tyo[l]; ex->h_tileXOff[l] = tileXo[l]; ex->h_tileYOff[l] = tileYo[l]; }
    int rc;
    if ((rc = ensure(ex, ex->d_tabX, ex->tabXCap, std::max<size_t>(tx.size(), 1)))) return rc;
    if ((rc = ensure(ex, ex->d_tabY, ex->tabYCap, std::max<size_t>(ty.size(), 1)))) return rc;
    if ((rc = ensure(ex, ex->d_cells, ex->cellsCap, std::max<size_t>(ex->h_cells.size(), 1)))) return rc;
    if ((rc = ensure(ex, ex->d_tiles, ex->tilesCap, std::max<size_t>(ex->h_tiles.size(), 1)))) return rc;
    if ((rc = ensure(ex, ex->d_pathLut, ex->pathLutCap, std::max<size_t>(ex->h_pathLut.size(), 1)))) return rc;
    cudaStream_t s = ex->stream;
    if (!tx.empty()) CUDA_TRY(ex, cudaMemcpyAsync(ex->d_tabX, tx.data(), tx.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
    if (!ty.empty()) CUDA_TRY(ex, cudaMemcpyAsync(ex->d_tabY, ty.data(), ty.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
    if (!ex->h_cells.empty())
        CUDA_TRY(ex, cudaMemcpyAsync(ex->d_cells, ex->h_cells.data(), ex->h_cells.size() * sizeof(OrbxCell), cudaMemcpyHostToDevice, s));
    if (!ex->h_pathLut.empty())
        CUDA_TRY(ex, cudaMemcpyAsync(ex->d_pathLut, ex->h_pathLut.data(), ex->h_pathLut.size() * sizeof(unsigned short), cudaMemcpyHostToDevice, s));
    if (!ex->h_tiles.empty())
        CUDA_TRY(ex, cudaMemcpyAsync(ex->d_tiles, ex->h_tiles.data(), ex->h_tiles.size() * sizeof(BlurTile), cudaMemcpyHostToDevice, s));
    CUDA_TRY(ex, cudaStreamSynchronize(s));  // the host vectors above are about to go out of scope
    return ORBX_OK;
}

int ensure_buffers(orbx_extractor *ex, int batch) {
    const OrbxGeom &G = ex->geom;
    int rc;
    const size_t B = (size_t)batch;
    size_t pyrNeed = B * (size_t)G.frameBytes;
    if (pyrNeed > ex->pyrCap || !ex->d_pyr) {
        if (ex->d_pyr) cudaFree(ex->d_pyr);
        if (ex->d_blur) cudaFree(ex->d_blur);
        ex->d_pyr = ex->d_blur = nullptr; ex->pyrCap = 0;
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_pyr, pyrNeed));
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_blur, pyrNeed));
        ex->pyrCap = pyrNeed;
    }
    size_t slotsNeed = B * (size_t)G.slotsTotal;
    if (slotsNeed > ex->slotsCap || !ex->d_slots) {
        if (ex->d_slots) cudaFree(ex->d_slots);
        if (ex->d_ptNode) cudaFree(ex->d_ptNode);
        if (ex->d_ptXY) cudaFree(ex->d_ptXY);
        ex->d_slots = ex->d_ptNode = nullptr; ex->d_ptXY = nullptr; ex->slotsCap = 0;
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_slots, slotsNeed * sizeof(uint32_t)));
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_ptNode, slotsNeed * sizeof(uint32_t)));
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_ptXY, slotsNeed * sizeof(float2)));
        ex->slotsCap = slotsNeed;
    }
    if ((rc = ensure(ex, ex->d_cellCnt, ex->cellCntCap, B * (size_t)std::max(G.nCellsTotal, 1)))) return rc;
    if ((rc = ensure(ex, ex->d_hist, ex->histCap, B * (size_t)G.nlevels * (size_t)ex->maxIni * QT_TREE))) return rc;
    if ((rc = ensure(ex, ex->d_finalPos, ex->finalPosCap, B * (size_t)G.nlevels * (size_t)ex->maxIni * QT_TREE))) return rc;
    if ((rc = ensure(ex, ex->d_best, ex->bestCap, B * (size_t)std::max(G.selTotal, 1)))) return rc;
    if ((rc = ensure(ex, ex->d_cellPrefix, ex->cellPrefixCap, B * (size_t)std::max(G.nCellsTotal, 1)))) return rc;
    if ((rc = ensure(ex, ex->d_denseList, ex->denseListCap, B * (size_t)std::max(G.nCellsTotal, 1)))) return rc;
    size_t selNeed = B * (size_t)std::max(G.selTotal, 1);
    if (selNeed > ex->selCap || !ex->d_sel) {
        if (ex->d_sel) cudaFree(ex->d_sel);
        if (ex->d_work) cudaFree(ex->d_work);
        ex->d_sel = nullptr; ex->d_work = nullptr; ex->selCap = 0;
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_sel, selNeed * sizeof(float4)));
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_work, selNeed * sizeof(OrbxWork)));
        ex->selCap = selNeed;
    }
    return ORBX_OK;
}

int prepare(orbx_extractor *ex, int rows, int cols, const int32_t *rects, int nRects, int lap0, int lap1, int batch) {
    if (rows > ex->maxH || cols > ex->maxW) {
        ex->err = "image larger than the max_width × max_height this handle was created for";
        return ORBX_ERR_ARG;
    }
    if (nRects < 0 || nRects > ORBX_MAX_RECTS || (nRects > 0 && !rects)) {
        ex->err = "n_rects out of range (max " + std::to_string(ORBX_MAX_RECTS) + ")";
        return ORBX_ERR_ARG;
    }
    int rc;                                   // (the public entry points hold an OrbxDeviceGuard on ex->device)
    bool sizeChanged = rows != ex->curRows || cols != ex->curCols;
    if (sizeChanged) {
        if ((rc = build_geometry(ex, rows, cols))) { ex->curRows = ex->curCols = -1; return rc; }
        if ((rc = upload_tables(ex))) { ex->curRows = ex->curCols = -1; return rc; }
        ex->curRows = rows; ex->curCols = cols;
        ex->geomDirty = true;
    }
    std::vector<int> r(rects ? rects : nullptr, rects ? rects + 4 * nRects : nullptr);
    if (ex->geomDirty || lap0 != ex->curLap0 || lap1 != ex->curLap1 || r != ex->curRects) {
        ex->geom.lap0 = lap0; ex->geom.lap1 = lap1; ex->geom.nRects = nRects;
        for (int i = 0; i < 4 * nRects; ++i) ex->geom.rects[i] = rects[i];
        // the previous batch may still be reading d_geom
        CUDA_TRY(ex, cudaStreamSynchronize(ex->stream));
        CUDA_TRY(ex, cudaMemcpyAsync(ex->d_geom, &ex->geom, sizeof(OrbxGeom), cudaMemcpyHostToDevice, ex->stream));
        CUDA_TRY(ex, cudaStreamSynchronize(ex->stream));
        ex->curLap0 = lap0; ex->curLap1 = lap1; ex->curRects = r;
        ex->geomDirty = false;
    }
    return ensure_buffers(ex, batch);
}

// ---- TMA tensor maps (host): cuTensorMapEncodeTiled through the runtime's driver entry point ----
typedef CUresult (*PFN_orbxTmaEncode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_orbxTmaEncode tma_encoder() {
    static PFN_orbxTmaEncode fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
        return (PFN_orbxTmaEncode)f;
    }();
    return fn;
}
// rank-3 u8 map (x bytes, y rows, frame) over `batch` planes of w×h bytes with the given row pitch and plane stride, box = boxW × boxH × 1.
// false when the layout does not meet TMA's rules (16-byte aligned base, pitch and stride; box ≤ 256 per dimension): the caller then
// stages with ordinary loads.
bool tma_encode_level(CUtensorMap *out, const void *base, int w, int h, int batch, long long pitch, long long planeStride, int boxW, int boxH) {
    PFN_orbxTmaEncode fn = tma_encoder();
    if (!fn) return false;
    if (((uintptr_t)base & 15) || (pitch & 15) || (planeStride & 15) || (boxW & 15) || boxW < 16 || boxW > 256 || boxH < 1 || boxH > 256 || w < 1 || h < 1) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)std::max(batch, 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)planeStride};
    const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
    const cuuint32_t es[3] = {1u, 1u, 1u};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// rank-4 u8 map that views every row as chunks of 16 bytes: (byte in chunk, chunk, row, frame).  A box of {16, nChunks, rows, 1} lands in
// shared memory as rows of 16·nChunks bytes — wider than the 256-element limit of one box dimension — and starts on a chunk boundary by
// construction.  Chunks at or beyond ceil(w / 16) and rows outside [0, h) read as zero.
bool tma_encode_chunked(CUtensorMap *out, const void *base, int w, int h, int batch, long long pitch, long long planeStride, int boxChunks, int boxRows) {
    PFN_orbxTmaEncode fn = tma_encoder();
    if (!fn) return false;
    if (((uintptr_t)base & 15) || (pitch & 15) || (planeStride & 15) || boxChunks < 1 || boxChunks > 256 || boxRows < 1 || boxRows > 256 || w < 1 || h < 1) return false;
    if ((long long)((w + 15) / 16) * 16 > pitch) return false;       // the last chunk of a row must lie inside the row's pitch
    const cuuint64_t dims[4] = {16, (cuuint64_t)((w + 15) / 16), (cuuint64_t)h, (cuuint64_t)std::max(batch, 1)};
    const cuuint64_t strides[3] = {16, (cuuint64_t)pitch, (cuuint64_t)planeStride};
    const cuuint32_t box[4] = {16u, (cuuint32_t)boxChunks, (cuuint32_t)boxRows, 1u};
    const cuuint32_t es[4] = {1u, 1u, 1u, 1u};
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int harvest_stage_times(orbx_extractor *ex) {
    if (!ex->evPending) return ORBX_OK;
    CUDA_TRY(ex, cudaEventSynchronize(ex->ev[6]));
    for (int i = 0; i < 6; ++i) {
        float ms = 0.f;
        CUDA_TRY(ex, cudaEventElapsedTime(&ms, ex->ev[i], ex->ev[i + 1]));
        ex->stageMs[i] += ms;
    }
    ++ex->stageCalls;
    ex->evPending = false;
    return ORBX_OK;
}

// enqueue the whole pipeline for `batch` frames whose level 0 lives at (in0, stride, pitch)
// (in0 and the four output pointers already point at frame `first`; internal per-frame buffers are offset here)
int run_pipeline(orbx_extractor *ex, const uint8_t *in0, long long in0Stride, int in0Pitch, int batch,
                 orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int *d_nOut, int *d_mono, int first = 0,
                 cudaStream_t onStream = nullptr) {
    const OrbxGeom &G = ex->geom;
    ExParams P;
    const size_t f = (size_t)first;
    P.g = ex->d_geom;
    P.in0 = in0; P.in0Stride = in0Stride; P.in0Pitch = in0Pitch;
    P.pyr = ex->d_pyr + f * G.frameBytes; P.blur = ex->d_blur + f * G.frameBytes;
    P.cells = ex->d_cells;
    P.tabX = ex->d_tabX; P.tabY = ex->d_tabY;
    P.slots = ex->d_slots + f * G.slotsTotal; P.cellCnt = ex->d_cellCnt + f * G.nCellsTotal;
    P.ptXY = ex->d_ptXY + f * G.slotsTotal; P.ptNode = ex->d_ptNode + f * G.slotsTotal;
    P.sel = ex->d_sel + f * G.selTotal; P.selCnt = ex->d_selCnt + f * G.nlevels;
    P.work = ex->d_work + f * G.selTotal; P.workCnt = ex->d_workCnt + f;
    P.kps = d_kps; P.desc = d_desc; P.cap = cap; P.nOut = d_nOut; P.monoIdx = d_mono;
    P.pattern = ex->d_pattern;
    cudaStream_t s = onStream ? onStream : ex->stream;
    const bool prof = ex->profiling && !onStream;
    if (prof) {
        int rc = harvest_stage_times(ex);
        if (rc) return rc;
        CUDA_TRY(ex, cudaEventRecord(ex->ev[0], s));
    }
    // K1
    for (int l = 1; l < G.nlevels; ++l) {
        // TMA form: one warp per 128 × PYR2_RS strip (needs a 16-byte aligned source; level 0 may be the caller's own buffer)
        if (ex->useTmaPyr) {
            const OrbxLevel &SV = G.lv[l - 1], &DV = G.lv[l];
            Pyr2Args A2;
            A2.dst = P.pyr + DV.off; A2.dstStride = G.frameBytes; A2.dp = DV.pitch; A2.dw = DV.w; A2.dh = DV.h;
            A2.tabX = ex->d_tabX + ex->h_tabXOff[l]; A2.tabY = ex->d_tabY + ex->h_tabYOff[l];
            A2.tilesX = (DV.w + 127) / 128;
            A2.nStrips = A2.tilesX * ((DV.h + PYR2_RS - 1) / PYR2_RS);
            A2.boxW = orbx_align_up((int)ceil(127.0 * SV.w / DV.w) + 1 + 12 + 15, 16);
            A2.boxH = (int)ceil((PYR2_RS - 1) * (double)SV.h / DV.h) + 4;
            A2.slotBytes = orbx_align_up(A2.boxW * A2.boxH + 8, 128);
            CUtensorMap srcMap;
            const bool ok = A2.boxW <= 256 && A2.boxH <= 256 &&
                            tma_encode_level(&srcMap, l == 1 ? P.in0 : P.pyr + SV.off, SV.w, SV.h, batch, l == 1 ? P.in0Pitch : SV.pitch,
                                             l == 1 ? P.in0Stride : G.frameBytes, A2.boxW, A2.boxH);
            if (ok) {
                const int WPB2 = 4;
                const size_t smem2 = (size_t)A2.slotBytes * WPB2;
                if (smem2 > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_pyr_level_tma<WPB2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
                k_pyr_level_tma<WPB2><<<dim3((A2.nStrips + WPB2 - 1) / WPB2, batch), WPB2 * 32, smem2, s>>>(A2, srcMap);
                ++ex->launches;
                continue;
            }
        }
        const int tilesX = (G.lv[l].w + PYR_TW - 1) / PYR_TW, tilesY = (G.lv[l].h + PYR_TH - 1) / PYR_TH;
        const int srcRows = (int)ceil((PYR_TH - 1) * (double)G.lv[l - 1].h / G.lv[l].h) + 4;
        const int srcPitch = 16 * ((int)ceil(((PYR_TW - 1) * (double)G.lv[l - 1].w / G.lv[l].w + 18.0) / 16.0) + 1);
        const size_t smem = (size_t)srcRows * (srcPitch + PYR_TW * sizeof(uint16_t));
        if (smem > 48 * 1024) { ex->err = "scale factor too large for the pyramid kernel"; return ORBX_ERR_ARG; }
        PyrArgs A;
        if (l == 1) { A.src = P.in0; A.srcStride = P.in0Stride; A.sp = P.in0Pitch; }
        else { A.src = P.pyr + G.lv[l - 1].off; A.srcStride = G.frameBytes; A.sp = G.lv[l - 1].pitch; }
        A.sw = G.lv[l - 1].w; A.sh = G.lv[l - 1].h;
        A.dst = P.pyr + G.lv[l].off; A.dstStride = G.frameBytes; A.dp = G.lv[l].pitch; A.dw = G.lv[l].w; A.dh = G.lv[l].h;
        A.tabX = ex->d_tabX + ex->h_tabXOff[l]; A.tabY = ex->d_tabY + ex->h_tabYOff[l];
        A.tileX = ex->d_tabX + ex->h_tileXOff[l]; A.tileY = ex->d_tabY + ex->h_tileYOff[l];
        A.tilesX = tilesX; A.srcRows = srcRows; A.srcPitch = srcPitch;
        k_pyr_level<<<dim3(tilesX * tilesY, batch), 256, smem, s>>>(A);
        ++ex->launches;
    }
    if (prof) CUDA_TRY(ex, cudaEventRecord(ex->ev[1], s));
    // K2
    {
        const int WPB = 4;
        if (ex->fastV1) {
            const FastSmem L = fast_smem_layout(G.maxCw, G.maxCh, ex->maxSlotCap);
            const size_t smem = (size_t)L.total * WPB;
            if (smem > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_fast_cells_v1<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            dim3 grd((G.nCellsTotal + WPB - 1) / WPB, batch);
            if (G.nCellsTotal > 0) {
                k_fast_cells_v1<WPB><<<grd, WPB * 32, smem, s>>>(P, ex->maxSlotCap);
                ++ex->launches;
            }
        } else {
            // dense-cell list of this run: frames [f, f+batch) own the list rows f*nCells.. and the counter at index f
            int2 *denseList = ex->d_denseList + f * (size_t)G.nCellsTotal;
            int *denseN = ex->d_dense + f;
            CUDA_TRY(ex, cudaMemsetAsync(denseN, 0, sizeof(int), s));
            ex->denseSlots.push_back((int)f);
            // Consecutive levels whose cells need about the same shared memory (within 20 %) form a group: the small top
            // levels have much taller cells (2 rows of cells cover the level) and would otherwise set the per-warp
            // footprint, hence the resident warps, for everybody.
            struct Grp { int cellBase, nCells, cw, ch; size_t need; int l0, l1; };
            Grp grp[ORBX_MAX_LEVELS];
            int nGrp = 0;
            for (int l0 = 0; l0 < G.nlevels;) {
                int cwMax = 0, chMax = 0, l1 = l0, nC = 0, cellBase = -1;
                size_t lo = 0, hi = 0;
                for (; l1 < G.nlevels; ++l1) {
                    const OrbxLevel &V = G.lv[l1];
                    if (V.nCells <= 0) continue;
                    const size_t need = (size_t)fast_smem_layout(V.wCell + 6, V.hCell + 6, 1).total2;
                    // (up to kFreeSlot bytes per warp the register file, not shared memory, limits the resident warps: no need to split)
                    const size_t kFreeSlot = 7000;
                    if (nC > 0 && std::max(hi, need) > kFreeSlot && (std::max(hi, need) * 5 > std::min(lo, need) * 6)) break;
                    lo = nC ? std::min(lo, need) : need; hi = std::max(hi, need);
                    cwMax = std::max(cwMax, V.wCell + 6); chMax = std::max(chMax, V.hCell + 6);
                    if (cellBase < 0) cellBase = V.cellBase;
                    nC += V.nCells;
                }
                if (nC > 0) grp[nGrp++] = Grp{cellBase, nC, cwMax, chMax, (size_t)fast_smem_layout(cwMax, chMax, 1).total2, l0, l1};
                l0 = l1;
            }
            auto launch_fast = [&](FastRange R, dim3 grid, size_t smem) -> int {
                const int rpA = fast_smem_layout(R.cw, R.ch, 1).roiPitch, rpB = R.nTall > 0 ? fast_smem_layout(R.tallCw, R.tallCh, 1).roiPitch : rpA;
                const int rpSel = rpA == rpB ? rpA : 0;         // compile-time ROI pitch for the two usual widths
#define ORBX_LAUNCH_FAST(RPV)                                                                                                             \
                do {                                                                                                                      \
                    if (smem > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_fast_cells<WPB, RPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                    k_fast_cells<WPB, RPV><<<grid, WPB * 32, smem, s>>>(P, R);                                                       \
                } while (0)
                if (rpSel == 44) ORBX_LAUNCH_FAST(44);
                else if (rpSel == 48) ORBX_LAUNCH_FAST(48);
                else if (rpSel == 52) ORBX_LAUNCH_FAST(52);
                else ORBX_LAUNCH_FAST(0);
#undef ORBX_LAUNCH_FAST
                ++ex->launches;
                return ORBX_OK;
            };
            // the usual shape is one big group plus a small group of tall cells: one launch, tall cells on several slots
            bool merged = false;
            if (nGrp == 2 && grp[1].nCells * 8 <= grp[0].nCells) {
                const int slots = (int)((grp[1].need + grp[0].need - 1) / grp[0].need);
                if (slots <= WPB) {
                    FastRange R{grp[0].cellBase, grp[0].nCells, grp[0].cw, grp[0].ch, grp[1].cellBase, grp[1].nCells, grp[1].cw, grp[1].ch, slots, 0, denseList, denseN};
                    R.nTallBlocks = (grp[1].nCells + WPB / slots - 1) / (WPB / slots);
                    int rcL = launch_fast(R, dim3(R.nTallBlocks + (R.nCells + WPB - 1) / WPB, batch), grp[0].need * WPB);
                    if (rcL) return rcL;
                    merged = true;
                }
            }
            for (int i = 0; i < nGrp && !merged; ++i) {
                FastRange R{grp[i].cellBase, grp[i].nCells, grp[i].cw, grp[i].ch, 0, 0, grp[i].cw, grp[i].ch, 1, 0, denseList, denseN};
                int rcL = launch_fast(R, dim3((R.nCells + WPB - 1) / WPB, batch), grp[i].need * WPB);
                if (rcL) return rcL;
            }
            // cells with more corner candidates than the two-phase kernel's queue holds go to the single-phase kernel
            if (G.nCellsTotal > 0) {
                const FastSmem L1 = fast_smem_layout(G.maxCw, G.maxCh, ex->maxSlotCap);
                const size_t smem = (size_t)L1.total * WPB;
                if (smem > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_fast_cells_v1_list<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_fast_cells_v1_list<WPB><<<2 * ex->nSM, WPB * 32, smem, s>>>(P, ex->maxSlotCap, denseList, denseN);
                ++ex->launches;
            }
        }
    }
    if (prof) CUDA_TRY(ex, cudaEventRecord(ex->ev[2], s));
    // blur tiles come in by TMA where the level's layout allows it (level 0 may be the caller's own buffer)
    OrbxTmaMaps blurSrcMaps;
    int blurTmaMask = 0;
    memset(&blurSrcMaps, 0, sizeof(blurSrcMaps));
    if (ex->useTmaPyr)
        for (int l = 0; l < G.nlevels; ++l) {
            const OrbxLevel &V = G.lv[l];
            if (tma_encode_chunked(&blurSrcMaps.m[l], l == 0 ? P.in0 : P.pyr + V.off, V.w, V.h, batch, l == 0 ? P.in0Pitch : V.pitch,
                                   l == 0 ? P.in0Stride : G.frameBytes, BLUR_SP / 16, BLUR_TH + 6))
                blurTmaMask |= 1 << l;
        }
    // After FAST the pipeline forks: the latency-bound quadtree chain (+ assemble) goes to the partner stream, which has the
    // HIGHEST stream priority — its blocks are dispatched as soon as they are ready — while the instruction-bound blur (it depends
    // on the pyramid only) stays on this stream and fills the issue slots the quadtree kernels leave idle.  (ORBX_BLUR_ON_AUX: the
    // earlier arrangement, blur on an equal-priority partner stream, where the block scheduler ran the two mostly back to back.)
    int aux = -1;
    cudaStream_t sQ = s;
    if (!prof && ex->overlapBlur) {
        aux = 0;
        for (int i = 0; i < ORBX_MAX_SIDE; ++i) if (onStream && onStream == ex->sSide[i]) aux = 1 + i;
        CUDA_TRY(ex, cudaEventRecord(ex->evPyrDone[aux], s));
        CUDA_TRY(ex, cudaStreamWaitEvent(ex->sAux[aux], ex->evPyrDone[aux], 0));
#ifdef ORBX_BLUR_ON_AUX
        dim3 grdB((unsigned)ex->h_tiles.size(), batch);
        k_blur<<<grdB, dim3(64, BLUR_STRIPS), 0, ex->sAux[aux]>>>(P, ex->d_tiles, blurSrcMaps, blurTmaMask);
        ++ex->launches;
        CUDA_TRY(ex, cudaEventRecord(ex->evBlurDone[aux], ex->sAux[aux]));
#else
        sQ = ex->sAux[aux];
#endif
    }
    // K3
    {
        const QtSmem L = qt_smem_layout(ex->nodeCapMax, ex->maxCellsLevel);
        if (L.total > 200 * 1024) { ex->err = "nfeatures too large for the quadtree kernel's shared memory"; return ORBX_ERR_ARG; }
        dim3 grd(G.nlevels, batch);
        const QtSmem LN = qt_smem_layout(ex->nodeCapMax, 0);
        const size_t smemN = (size_t)LN.total + 8 * (size_t)ex->nodeCapMax;          // + node identities
        const bool useHist = ex->useHistQuadtree && smemN <= 200 * 1024;
        const bool big = ex->nodeCapMax > 600;   // large quotas (4K): few, long blocks → 512 threads each
        QtTables T;
        T.maxIni = ex->maxIni;
        T.hist = ex->d_hist + f * G.nlevels * (size_t)ex->maxIni * QT_TREE;
        T.finalPos = ex->d_finalPos + f * G.nlevels * (size_t)ex->maxIni * QT_TREE;
        T.best = ex->d_best + f * G.selTotal;
        T.cellPrefix = ex->d_cellPrefix + f * G.nCellsTotal;
        T.deep = ex->d_deep + f * G.nlevels;
        T.pathLut = ex->d_pathLut;
        if (useHist) {
            const size_t smemP = (size_t)(ex->maxCellsLevel + 1 + QT_THREADS + 2) * sizeof(int);
            if (smemP > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_qt_prefix<QT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemP));
            k_qt_prefix<QT_THREADS><<<grd, QT_THREADS, smemP, sQ>>>(P, T, ex->maxCellsLevel);
            dim3 grdC(QT_CLS_BLOCKS, G.nlevels, batch);
            k_qt_classify<<<grdC, 128, 0, sQ>>>(P, T);
            if (big) {
                if (smemN > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_qt_nodes<QT_THREADS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemN));
                k_qt_nodes<QT_THREADS_BIG><<<grd, QT_THREADS_BIG, smemN, sQ>>>(P, T, ex->nodeCapMax);
            } else {
                if (smemN > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_qt_nodes<QT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemN));
                k_qt_nodes<QT_THREADS><<<grd, QT_THREADS, smemN, sQ>>>(P, T, ex->nodeCapMax);
            }
            ex->launches += 3;
        }
        // general kernel: everything when the fast path is off, otherwise only the flagged (frame, level) pairs
        if (big) {
            if (L.total > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_quadtree<QT_THREADS_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
            k_quadtree<QT_THREADS_BIG><<<grd, QT_THREADS_BIG, L.total, sQ>>>(P, ex->nodeCapMax, ex->maxCellsLevel, useHist ? T.deep : nullptr);
        } else {
            if (L.total > 48 * 1024) CUDA_TRY(ex, cudaFuncSetAttribute(k_quadtree<QT_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
            k_quadtree<QT_THREADS><<<grd, QT_THREADS, L.total, sQ>>>(P, ex->nodeCapMax, ex->maxCellsLevel, useHist ? T.deep : nullptr);
        }
        ++ex->launches;
        if (useHist) {
            const size_t smemS = (size_t)ex->nodeCapMax * sizeof(unsigned);
            if (smemS <= 48 * 1024) {
                k_qt_attach_select<512><<<grd, 512, smemS, sQ>>>(P, T);
                ++ex->launches;
            } else {
                dim3 grdC(QT_CLS_BLOCKS, G.nlevels, batch);
                k_qt_attach<<<grdC, 128, 0, sQ>>>(P, T);
                k_qt_select<<<grd, 128, 0, sQ>>>(P, T);
                ex->launches += 2;
            }
        } else {
            CUDA_TRY(ex, cudaMemsetAsync(T.deep, 0, (size_t)batch * G.nlevels * sizeof(int), sQ));
        }
    }
    if (prof) CUDA_TRY(ex, cudaEventRecord(ex->ev[3], s));
    // K7
    k_assemble<<<batch, 256, 0, sQ>>>(P);
    ++ex->launches;
    if (prof) CUDA_TRY(ex, cudaEventRecord(ex->ev[4], s));
    // K5
    if (aux < 0) {
        dim3 grd((unsigned)ex->h_tiles.size(), batch);
        k_blur<<<grd, dim3(64, BLUR_STRIPS), 0, s>>>(P, ex->d_tiles, blurSrcMaps, blurTmaMask);
        ++ex->launches;
    } else {
#ifndef ORBX_BLUR_ON_AUX
        CUDA_TRY(ex, cudaEventRecord(ex->evBlurDone[aux], sQ));       // quadtree chain + assemble done
        dim3 grd((unsigned)ex->h_tiles.size(), batch);
        k_blur<<<grd, dim3(64, BLUR_STRIPS), 0, s>>>(P, ex->d_tiles, blurSrcMaps, blurTmaMask);
        ++ex->launches;
#endif
        CUDA_TRY(ex, cudaStreamWaitEvent(s, ex->evBlurDone[aux], 0));
    }
    if (prof) CUDA_TRY(ex, cudaEventRecord(ex->ev[5], s));
    // K4 + K6
    {
        const int WPB = 8;
        dim3 grd((G.selTotal + WPB - 1) / WPB, batch);
        bool tmaOk = ex->useTmaPyr;
        OrbxTmaMaps pyrMaps, blurMaps;
        if (tmaOk) {
            memset(&pyrMaps, 0, sizeof(pyrMaps)); memset(&blurMaps, 0, sizeof(blurMaps));
            for (int l = 0; l < G.nlevels && tmaOk; ++l) {
                const OrbxLevel &V = G.lv[l];
                tmaOk = tma_encode_level(&pyrMaps.m[l], l == 0 ? P.in0 : P.pyr + V.off, V.w, V.h, batch, l == 0 ? P.in0Pitch : V.pitch,
                                         l == 0 ? P.in0Stride : G.frameBytes, OD_AW, OD_AH) &&
                        tma_encode_level(&blurMaps.m[l], P.blur + V.off, V.w, V.h, batch, V.pitch, G.frameBytes, OD_BW, OD_BH);
            }
        }
        if (tmaOk) {
            const size_t smemO = (size_t)OD_SLOT * WPB;
            k_orient_desc_tma<WPB><<<grd, WPB * 32, smemO, s>>>(P, pyrMaps, blurMaps);
        } else {
            k_orient_desc<WPB><<<grd, WPB * 32, 0, s>>>(P);
        }
        ++ex->launches;
    }
    if (prof) { CUDA_TRY(ex, cudaEventRecord(ex->ev[6], s)); ex->evPending = true; }
    CUDA_TRY(ex, cudaGetLastError());
    ex->lastBatch = std::max(ex->lastBatch, first + batch);
    return ORBX_OK;
}

// Device-resident batch: split into sub-batches issued round-robin on two side streams, so the latency-bound
// kernels of one sub-batch (quadtree, orientation gathers) overlap the ALU-bound FAST kernel of the next.
// With per-stage profiling on, everything stays on the main stream (stage times are then serial times).
int run_batch(orbx_extractor *ex, const uint8_t *in0, long long in0Stride, int in0Pitch, int batch, orbx_keypoint *d_kps,
              uint8_t *d_desc, int cap, int *d_nOut, int *d_mono) {
    ex->lastBatch = 0;
    ex->denseSlots.clear();
    // sub-batches of at least ~128 frames (smaller launches lose more to fixed latencies than the overlap gains)
    const int nSub = ex->profiling ? 1 : (ex->nSub > 0 ? (batch >= 2 * ex->nSub ? ex->nSub : 1) : std::min(4, std::max(batch >= 32 ? 2 : 1, batch / 128)));
    if (nSub == 1) return run_pipeline(ex, in0, in0Stride, in0Pitch, batch, d_kps, d_desc, cap, d_nOut, d_mono);
    cudaStream_t *side = ex->sSide;
    const int nSide = ex->nSide;
    CUDA_TRY(ex, cudaEventRecord(ex->evFork, ex->stream));
    for (int i = 0; i < nSide; ++i) CUDA_TRY(ex, cudaStreamWaitEvent(side[i], ex->evFork, 0));
    const int sub = (batch + nSub - 1) / nSub;
    int k = 0;
    for (int c0 = 0; c0 < batch; c0 += sub, ++k) {
        const int cn = std::min(sub, batch - c0);
        int rc = run_pipeline(ex, in0 + (long long)c0 * in0Stride, in0Stride, in0Pitch, cn, d_kps + (size_t)c0 * cap,
                              d_desc + (size_t)c0 * cap * 32, cap, d_nOut + c0, d_mono + c0, c0, side[k % nSide]);
        if (rc) return rc;
    }
    for (int i = 0; i < nSide; ++i) {
        CUDA_TRY(ex, cudaEventRecord(ex->evJoin[i], side[i]));
        CUDA_TRY(ex, cudaStreamWaitEvent(ex->stream, ex->evJoin[i], 0));
    }
    return ORBX_OK;
}

}  // namespace

extern "C" {

int orbx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *orbx_last_error(const orbx_extractor *ex) { return ex ? ex->err.c_str() : tl_error.c_str(); }

orbx_extractor *orbx_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int device,
                            int max_width, int max_height, int max_batch) {
    if (nlevels < 1 || nlevels > ORBX_MAX_LEVELS || nfeatures < 0 || !(scale_factor > 1.0f) || ini_th < 0 ||
        ini_th > 255 || min_th < 0 || min_th > 255 || max_width <= 0 || max_height <= 0 || max_batch <= 0) {
        tl_error = "orbx_create: bad parameter";
        return nullptr;
    }
    int ndev = orbx_device_count();
    if (device < 0 || device >= ndev) {
        tl_error = "orbx_create: no such CUDA device (liborbx has no CPU fallback)";
        return nullptr;
    }
    orbx_extractor *ex = new orbx_extractor;
    ex->device = device; ex->nfeatures = nfeatures; ex->nlevels = nlevels; ex->iniTh = ini_th; ex->minTh = min_th;
    ex->scaleFactor = scale_factor;
    ex->maxW = max_width; ex->maxH = max_height; ex->maxBatch = max_batch;
    // ctor maths of the reference (:415-447)
    ex->sf[0] = 1.f; ex->sig2[0] = 1.f;
    for (int i = 1; i < nlevels; ++i) {
        ex->sf[i] = (float)(ex->sf[i - 1] * ex->scaleFactor);
        ex->sig2[i] = ex->sf[i] * ex->sf[i];
    }
    for (int i = 0; i < nlevels; ++i) { ex->inv[i] = 1.0f / ex->sf[i]; ex->invsig2[i] = 1.0f / ex->sig2[i]; }
    float factor = (float)(1.0f / ex->scaleFactor);
    float want = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        ex->quota[l] = cv_round_f(want);
        sum += ex->quota[l];
        want *= factor;
    }
    ex->quota[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax (:453-468)
    {
        int v, v0;
        const int vmax = (int)floorf(ORBX_HALF_PATCH * sqrtf(2.f) / 2 + 1);
        const int vmin = (int)ceilf(ORBX_HALF_PATCH * sqrtf(2.f) / 2);
        const double hp2 = ORBX_HALF_PATCH * ORBX_HALF_PATCH;
        for (v = 0; v <= vmax; ++v) ex->umax[v] = (int)lrint(sqrt(hp2 - v * v));
        for (v = ORBX_HALF_PATCH, v0 = 0; v >= vmin; --v) {
            while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
            ex->umax[v] = v0;
            ++v0;
        }
    }
    auto fail = [&](const std::string &m) { tl_error = m; orbx_destroy(ex); return (orbx_extractor *)nullptr; };
#define CREATE_TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
    OrbxDeviceGuard dg_(device);
    CREATE_TRY(dg_.status);
    CREATE_TRY(cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&ex->sH2D, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&ex->sD2H, cudaStreamNonBlocking));
    for (int i = 0; i < ORBX_MAX_SIDE; ++i) CREATE_TRY(cudaStreamCreateWithFlags(&ex->sSide[i], cudaStreamNonBlocking));
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ex->nSM = v; }
#ifdef ORBX_DEV_KNOBS   // developer builds only (make DEV=1): schedule knobs and copy-skipping diagnostics never ship
    if (const char *e = getenv("ORBX_NSIDE")) ex->nSide = std::min(ORBX_MAX_SIDE, std::max(1, atoi(e)));
    if (const char *e = getenv("ORBX_NSUB")) ex->nSub = std::min(16, std::max(1, atoi(e)));
    if (const char *e = getenv("ORBX_NSTEADY")) ex->nSteady = std::min(ORBX_MAX_CHUNKS - 2, std::max(1, atoi(e)));
    ex->overlapBlur = getenv("ORBX_SERIAL_BLUR") == nullptr;
    ex->dbgChunks = getenv("ORBX_DEBUG_CHUNKS") != nullptr;
    ex->dbgSkipH2D = getenv("ORBX_DEBUG_SKIP_H2D") != nullptr;
    ex->dbgSkipD2H = getenv("ORBX_DEBUG_SKIP_D2H") != nullptr;
    if (const char *e = getenv("ORBX_CHUNK_PLAN")) ex->chunkPlan = e;
#endif
    CREATE_TRY(cudaEventCreateWithFlags(&ex->evFork, cudaEventDisableTiming));
    for (int i = 0; i < ORBX_MAX_SIDE; ++i) CREATE_TRY(cudaEventCreateWithFlags(&ex->evJoin[i], cudaEventDisableTiming));
    for (int i = 0; i <= ORBX_MAX_SIDE; ++i) {
#ifdef ORBX_BLUR_ON_AUX
        CREATE_TRY(cudaStreamCreateWithFlags(&ex->sAux[i], cudaStreamNonBlocking));
#else
        { int least = 0, greatest = 0;                     // numerically lowest = highest priority
          CREATE_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
          CREATE_TRY(cudaStreamCreateWithPriority(&ex->sAux[i], cudaStreamNonBlocking, greatest)); }
#endif
        CREATE_TRY(cudaEventCreateWithFlags(&ex->evPyrDone[i], cudaEventDisableTiming));
        CREATE_TRY(cudaEventCreateWithFlags(&ex->evBlurDone[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < ORBX_MAX_CHUNKS; ++i) {
        const unsigned evFlags = ex->dbgChunks ? cudaEventDefault : cudaEventDisableTiming;
        CREATE_TRY(cudaEventCreateWithFlags(&ex->evIn[i], evFlags));
        CREATE_TRY(cudaEventCreateWithFlags(&ex->evOut[i], evFlags));
    }
    CREATE_TRY(cudaMalloc((void **)&ex->d_geom, sizeof(OrbxGeom)));
    CREATE_TRY(cudaMalloc((void **)&ex->d_pattern, 1024));
    CREATE_TRY(cudaMemcpy(ex->d_pattern, h_pattern, 1024, cudaMemcpyHostToDevice));
    {
        static const int kUmax15[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        for (int i = 0; i <= ORBX_HALF_PATCH; ++i)
            if (ex->umax[i] != kUmax15[i]) return fail("orbx_create: umax table differs from the compiled-in HALF_PATCH_SIZE=15 table");
    }
    CREATE_TRY(cudaMalloc((void **)&ex->d_selCnt, (size_t)max_batch * ORBX_MAX_LEVELS * sizeof(int)));
    CREATE_TRY(cudaMalloc((void **)&ex->d_workCnt, (size_t)max_batch * sizeof(int)));
    CREATE_TRY(cudaMalloc((void **)&ex->d_deep, (size_t)max_batch * ORBX_MAX_LEVELS * sizeof(int)));
    CREATE_TRY(cudaMalloc((void **)&ex->d_dense, (size_t)max_batch * sizeof(int)));
    CREATE_TRY(cudaMemset(ex->d_dense, 0, (size_t)max_batch * sizeof(int)));
    ex->useHistQuadtree = getenv("ORBX_LEGACY_QUADTREE") == nullptr;
    ex->fastV1 = getenv("ORBX_FAST_V1") != nullptr;
    ex->useGraph = getenv("ORBX_NO_GRAPH") == nullptr;
    ex->useTmaPyr = getenv("ORBX_NO_TMA") == nullptr;     // same results either way; tiles are then staged with ordinary loads      // same kernels either way; the graph only removes launch overhead
    CREATE_TRY(cudaMalloc((void **)&ex->d_nOut, (size_t)max_batch * sizeof(int)));
    CREATE_TRY(cudaMalloc((void **)&ex->d_mono, (size_t)max_batch * sizeof(int)));
    CREATE_TRY(cudaHostAlloc((void **)&ex->h_nOut, (size_t)max_batch * sizeof(int), cudaHostAllocDefault));
    CREATE_TRY(cudaHostAlloc((void **)&ex->h_mono, (size_t)max_batch * sizeof(int), cudaHostAllocDefault));
#undef CREATE_TRY
    // size all image-dependent buffers for the largest image now, so the hot path never allocates
    int rc = build_geometry(ex, max_height, max_width);
    if (rc == ORBX_OK) rc = upload_tables(ex);
    if (rc == ORBX_OK) rc = ensure_buffers(ex, max_batch);
    if (rc != ORBX_OK) return fail("orbx_create: " + ex->err);
    ex->curRows = max_height; ex->curCols = max_width; ex->geomDirty = true;
    return ex;
}

void orbx_destroy(orbx_extractor *ex) {
    if (!ex) return;
    OrbxDeviceGuard dg_(ex->device);
    if (ex->stream) cudaStreamSynchronize(ex->stream);
    if (ex->oneExec) cudaGraphExecDestroy(ex->oneExec);
    if (ex->h_in) cudaFreeHost(ex->h_in);
    if (ex->h_single) cudaFreeHost(ex->h_single);
    if (ex->d_single) cudaFree(ex->d_single);
    if (ex->d_sm) cudaFree(ex->d_sm);
    void *ptrs[] = {ex->d_geom, ex->d_cells, ex->d_tiles, ex->d_pathLut, ex->d_tabX, ex->d_tabY,
                    ex->d_pattern, ex->d_pyr, ex->d_blur, ex->d_slots, ex->d_ptNode, ex->d_ptXY, ex->d_cellCnt,
                    ex->d_sel, ex->d_work, ex->d_selCnt, ex->d_workCnt, ex->d_hist, ex->d_finalPos, ex->d_best, ex->d_cellPrefix, ex->d_deep, ex->d_dense, ex->d_denseList, ex->d_stage, ex->d_kps, ex->d_desc, ex->d_nOut, ex->d_mono};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (ex->h_nOut) cudaFreeHost(ex->h_nOut);
    if (ex->h_mono) cudaFreeHost(ex->h_mono);
    for (int i = 0; i < 7; ++i) if (ex->ev[i]) cudaEventDestroy(ex->ev[i]);
    for (int i = 0; i < ORBX_MAX_CHUNKS; ++i) { if (ex->evIn[i]) cudaEventDestroy(ex->evIn[i]); if (ex->evOut[i]) cudaEventDestroy(ex->evOut[i]); }
    if (ex->evFork) cudaEventDestroy(ex->evFork);
    for (int i = 0; i < ORBX_MAX_SIDE; ++i) if (ex->evJoin[i]) cudaEventDestroy(ex->evJoin[i]);
    for (int i = 0; i <= ORBX_MAX_SIDE; ++i) {
        if (ex->sAux[i]) { cudaStreamSynchronize(ex->sAux[i]); cudaStreamDestroy(ex->sAux[i]); }
        if (ex->evPyrDone[i]) cudaEventDestroy(ex->evPyrDone[i]);
        if (ex->evBlurDone[i]) cudaEventDestroy(ex->evBlurDone[i]);
    }
    for (int i = 0; i < ORBX_MAX_SIDE; ++i) if (ex->sSide[i]) { cudaStreamSynchronize(ex->sSide[i]); cudaStreamDestroy(ex->sSide[i]); }
    if (ex->sH2D) cudaStreamDestroy(ex->sH2D);
    if (ex->sD2H) cudaStreamDestroy(ex->sD2H);
    if (ex->stream) cudaStreamDestroy(ex->stream);
    delete ex;
}

int orbx_params(const orbx_extractor *ex, float *sf, float *inv, float *sig2, float *invsig2, int32_t *quota) {
    if (!ex) return ORBX_ERR_ARG;
    for (int i = 0; i < ex->nlevels; ++i) {
        if (sf) sf[i] = ex->sf[i];
        if (inv) inv[i] = ex->inv[i];
        if (sig2) sig2[i] = ex->sig2[i];
        if (invsig2) invsig2[i] = ex->invsig2[i];
        if (quota) quota[i] = ex->quota[i];
    }
    return ORBX_OK;
}

int orbx_extract_batch_device(orbx_extractor *ex, const uint8_t *d_images, size_t frame_stride, int batch, int rows,
                              int cols, size_t step, const int32_t *rects, int n_rects, int lap0, int lap1,
                              orbx_keypoint *d_kps, uint8_t *d_desc, int cap, int32_t *d_n_out, int32_t *d_mono) {
    if (!ex) return ORBX_ERR_ARG;
    if (!d_images || rows <= 0 || cols <= 0 || batch <= 0) { ex->err = "empty image"; return ORBX_EMPTY; }
    if (batch > ex->maxBatch || !d_kps || !d_desc || !d_n_out || !d_mono || cap <= 0) {
        ex->err = "orbx_extract_batch_device: bad argument (batch > max_batch, null output or cap <= 0)";
        return ORBX_ERR_ARG;
    }
    OrbxDeviceGuard dg_(ex->device);
    int rc = prepare(ex, rows, cols, rects, n_rects, lap0, lap1, batch);
    if (rc) return rc;
    ex->lastIn0Internal = false;
    return run_batch(ex, d_images, (long long)frame_stride, (int)step, batch, d_kps, d_desc, cap, d_n_out, d_mono);
}

// Best effort: nothing of this handle is in flight any more (used before an error return, so that the caller may free or reuse
// the host buffers the asynchronous copies were reading and writing).
static void drain_streams(orbx_extractor *ex) {
    if (ex->sH2D) cudaStreamSynchronize(ex->sH2D);
    for (int i = 0; i < ORBX_MAX_SIDE; ++i) if (ex->sSide[i]) cudaStreamSynchronize(ex->sSide[i]);
    for (int i = 0; i <= ORBX_MAX_SIDE; ++i) if (ex->sAux[i]) cudaStreamSynchronize(ex->sAux[i]);
    if (ex->stream) cudaStreamSynchronize(ex->stream);
    if (ex->sD2H) cudaStreamSynchronize(ex->sD2H);
    cudaGetLastError();
}

// The host-buffer batch call.  copyOnly = the same copies on the same streams with the same chunking but no kernel launches
// (orbx_copy_only_batch: the transfer ceiling of the host path, a measurement aid).
static int host_batch(orbx_extractor *ex, const uint8_t *const *images, int batch, int rows, int cols, size_t step,
                      const int32_t *rects, int n_rects, int lap0, int lap1, orbx_keypoint *kps, uint8_t *desc,
                      int cap, int32_t *n_out, int32_t *mono_index, bool copyOnly) {
    int result = ORBX_OK;
    for (int b0 = 0; b0 < batch; b0 += ex->maxBatch) {
        const int nb = std::min(ex->maxBatch, batch - b0);
        const auto tStart = std::chrono::steady_clock::now();
        ++ex->dbgCalls;
        int rc = prepare(ex, rows, cols, rects, n_rects, lap0, lap1, nb);
        if (rc) return rc;
        const OrbxGeom &G = ex->geom;
        // staging for outputs
        const size_t need = (size_t)nb * cap;
        if (need > ex->outCap || !ex->d_kps) {
            if (ex->d_kps) cudaFree(ex->d_kps);
            if (ex->d_desc) cudaFree(ex->d_desc);
            ex->d_kps = nullptr; ex->d_desc = nullptr; ex->outCap = 0;
            CUDA_TRY(ex, cudaMalloc((void **)&ex->d_kps, need * sizeof(orbx_keypoint)));
            CUDA_TRY(ex, cudaMalloc((void **)&ex->d_desc, need * 32));
            ex->outCap = need;
        }
        // Software pipeline over chunks of the batch: H2D of chunk k+1 and D2H of chunk k-1 overlap the kernels
        // of chunk k (three streams, events between them).  PCIe moves 307 KB in and ~64 KB out per frame, about
        // half of the kernel time at 640x480, so the copies hide completely behind the compute stream.
        // The first two chunks are short (1/32 and 3/32 of the batch) so that the kernels start after a brief copy;
        // the rest is split evenly.
        int chunkLen[ORBX_MAX_CHUNKS];
        int nChunks = 0;
        if (nb >= 256) {
            // plan in 1024ths of the batch (developer knob ORBX_CHUNK_PLAN="32,96,..." overrides it)
            int plan[ORBX_MAX_CHUNKS], nPlan = 0;
            if (!ex->chunkPlan.empty()) {
                for (const char *q = ex->chunkPlan.c_str(); *q && nPlan < ORBX_MAX_CHUNKS;) { plan[nPlan++] = atoi(q); while (*q && *q != ',') ++q; if (*q == ',') ++q; }
            } else {
                // steady chunks of about 110 frames: smaller ones lose launch efficiency, larger ones lengthen the tail after the last
                // copy-in (measured on B200: 2048 frames of 640x480 run at 155 k frames/s in 8 chunks, 161 k in 16-20, 150 k in 32)
                const int nSteady = ex->nSteady > 0 ? ex->nSteady : std::min(ORBX_MAX_CHUNKS - 2, std::max(2, (nb - nb / 8 + 109) / 110));
                plan[nPlan++] = 32; plan[nPlan++] = 96;
                for (int i = 0; i < nSteady; ++i) plan[nPlan++] = (1024 - 128) / nSteady;
            }
            int left = nb;
            for (int i = 0; i < nPlan && left > 0; ++i) {
                const int c = i == nPlan - 1 ? left : std::min(left, std::max(1, (int)((long long)nb * plan[i] / 1024)));
                chunkLen[nChunks++] = c; left -= c;
            }
            if (left > 0) chunkLen[nChunks - 1] += left;
        } else {
            const int parts = nb >= 64 ? 4 : 1;
            int left = nb;
            for (int i = parts; i > 0; --i) { const int c = (left + i - 1) / i; chunkLen[nChunks++] = c; left -= c; }
        }
        cudaStream_t sC = ex->stream, sIn = ex->sH2D, sOut = ex->sD2H;
        // earlier asynchronous work of this handle must be finished before its buffers are refilled
        CUDA_TRY(ex, cudaEventRecord(ex->evOut[0], sC));
        CUDA_TRY(ex, cudaStreamWaitEvent(sIn, ex->evOut[0], 0));
        for (int k = 0, c0 = 0; k < nChunks; c0 += chunkLen[k], ++k) {
            const int cn = chunkLen[k];
            if (cn <= 0) continue;
            bool contiguous = true;
            for (int b = 1; b < cn && contiguous; ++b)
                contiguous = images[b0 + c0 + b] == images[b0 + c0] + (size_t)b * rows * step;
            uint8_t *lvl0 = ex->d_pyr + (size_t)c0 * G.frameBytes + G.lv[0].off;
            if (contiguous && step == (size_t)cols && G.lv[0].pitch == cols) {
                // frames are back to back and rows are dense: one strided copy for the whole chunk
#ifdef ORBX_DEV_KNOBS
                if (!(ex->dbgSkipH2D && ex->dbgCalls > 2))   // developer aid: reuse the frames the first calls copied in
#endif
                CUDA_TRY(ex, cudaMemcpy2DAsync(lvl0, (size_t)G.frameBytes, images[b0 + c0], (size_t)rows * cols, (size_t)rows * cols, cn,
                                               cudaMemcpyHostToDevice, sIn));
            } else if (contiguous && step == (size_t)cols) {
                // dense host frames but a padded device pitch: one flat copy into a staging area, then a re-pitch kernel
                const size_t bytes = (size_t)cn * rows * cols;
                if ((size_t)nb * rows * cols > ex->stageCap || !ex->d_stage) {
                    CUDA_TRY(ex, cudaStreamSynchronize(sIn));
                    if (ex->d_stage) cudaFree(ex->d_stage);
                    ex->d_stage = nullptr; ex->stageCap = 0;
                    CUDA_TRY(ex, cudaMalloc((void **)&ex->d_stage, (size_t)nb * rows * cols));
                    ex->stageCap = (size_t)nb * rows * cols;
                }
                uint8_t *stg = ex->d_stage + (size_t)c0 * rows * cols;
                CUDA_TRY(ex, cudaMemcpyAsync(stg, images[b0 + c0], bytes, cudaMemcpyHostToDevice, sIn));
                k_repitch<<<dim3((cols + 1023) / 1024, rows, cn), 256, 0, sIn>>>(stg, rows, cols, lvl0, G.frameBytes, G.lv[0].pitch);
                ++ex->launches;
            } else {
                for (int b = 0; b < cn; ++b)
                    CUDA_TRY(ex, cudaMemcpy2DAsync(lvl0 + (size_t)b * G.frameBytes, G.lv[0].pitch, images[b0 + c0 + b], step, cols, rows,
                                                   cudaMemcpyHostToDevice, sIn));
            }
            cudaStream_t sK = (nChunks > 1 && !ex->profiling) ? ex->sSide[k % ex->nSide] : sC;   // chunks rotate over the side streams
            CUDA_TRY(ex, cudaEventRecord(ex->evIn[k], sIn));
            CUDA_TRY(ex, cudaStreamWaitEvent(sK, ex->evIn[k], 0));
            ex->lastIn0Internal = true;
            if (c0 == 0) { ex->lastBatch = 0; ex->denseSlots.clear(); }
            if (!copyOnly) {
                rc = run_pipeline(ex, lvl0, G.frameBytes, G.lv[0].pitch, cn, ex->d_kps + (size_t)c0 * cap, ex->d_desc + (size_t)c0 * cap * 32, cap,
                                  ex->d_nOut + c0, ex->d_mono + c0, c0, sK == sC ? nullptr : sK);
                if (rc) return rc;
            }
            CUDA_TRY(ex, cudaEventRecord(ex->evOut[k], sK));
            CUDA_TRY(ex, cudaStreamWaitEvent(sOut, ex->evOut[k], 0));
            CUDA_TRY(ex, cudaMemcpyAsync(ex->h_nOut + c0, ex->d_nOut + c0, cn * sizeof(int), cudaMemcpyDeviceToHost, sOut));
            CUDA_TRY(ex, cudaMemcpyAsync(ex->h_mono + c0, ex->d_mono + c0, cn * sizeof(int), cudaMemcpyDeviceToHost, sOut));
#ifdef ORBX_DEV_KNOBS
            if (!ex->dbgSkipD2H)
#endif
            {
            CUDA_TRY(ex, cudaMemcpyAsync(kps + ((size_t)b0 + c0) * cap, ex->d_kps + (size_t)c0 * cap, (size_t)cn * cap * sizeof(orbx_keypoint),
                                         cudaMemcpyDeviceToHost, sOut));
            CUDA_TRY(ex, cudaMemcpyAsync(desc + ((size_t)b0 + c0) * cap * 32, ex->d_desc + (size_t)c0 * cap * 32, (size_t)cn * cap * 32,
                                         cudaMemcpyDeviceToHost, sOut));
            }
        }
        const auto tEnq = std::chrono::steady_clock::now();
        CUDA_TRY(ex, cudaStreamSynchronize(sOut));
        for (int i = 0; i < ex->nSide; ++i) CUDA_TRY(ex, cudaStreamSynchronize(ex->sSide[i]));
        CUDA_TRY(ex, cudaStreamSynchronize(sC));
        if (ex->dbgChunks && nChunks > 1) {   // developer aid: when each chunk's copy-in and kernels finished
            fprintf(stderr, "[orbx chunks] host enqueue %.2f ms, total %.2f ms;", std::chrono::duration<double, std::milli>(tEnq - tStart).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count());
            for (int k = 1; k < nChunks; ++k) {
                float a = 0, c = 0;
                cudaEventElapsedTime(&a, ex->evIn[0], ex->evIn[k]);
                cudaEventElapsedTime(&c, ex->evIn[0], ex->evOut[k]);
                fprintf(stderr, " %d:in+%.2f,out+%.2f", chunkLen[k], a, c);
            }
            float c0ms = 0; cudaEventElapsedTime(&c0ms, ex->evIn[0], ex->evOut[0]);
            fprintf(stderr, " (chunk0 %d out+%.2f)\n", chunkLen[0], c0ms);
        }
        for (int b = 0; b < nb && !copyOnly; ++b) {
            n_out[b0 + b] = ex->h_nOut[b];
            mono_index[b0 + b] = ex->h_mono[b];
            if (ex->h_nOut[b] > cap) result = ORBX_ERR_CAPACITY;
        }
    }
    if (result == ORBX_ERR_CAPACITY) ex->err = "keypoint capacity too small for at least one frame (see n_out)";
    return result;
}

static int host_batch_checked(orbx_extractor *ex, const uint8_t *const *images, int batch, int rows, int cols, size_t step,
                              const int32_t *rects, int n_rects, int lap0, int lap1, orbx_keypoint *kps, uint8_t *desc,
                              int cap, int32_t *n_out, int32_t *mono_index, bool copyOnly) {
    if (!ex) return ORBX_ERR_ARG;
    if (!images || rows <= 0 || cols <= 0 || batch <= 0) { ex->err = "empty image"; return ORBX_EMPTY; }
    for (int b = 0; b < batch; ++b) if (!images[b]) { ex->err = "empty image"; return ORBX_EMPTY; }
    if (!kps || !desc || !n_out || !mono_index || cap <= 0) { ex->err = "orbx_extract_batch: null output or cap <= 0"; return ORBX_ERR_ARG; }
    OrbxDeviceGuard dg_(ex->device);
    if (dg_.status != cudaSuccess) { ex->err = std::string("cudaSetDevice: ") + cudaGetErrorString(dg_.status); return ORBX_ERR_CUDA; }
    const int rc = host_batch(ex, images, batch, rows, cols, step, rects, n_rects, lap0, lap1, kps, desc, cap, n_out, mono_index, copyOnly);
    if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) drain_streams(ex);   // asynchronous copies may still touch the caller's buffers
    return rc;
}

int orbx_extract_batch(orbx_extractor *ex, const uint8_t *const *images, int batch, int rows, int cols, size_t step,
                       const int32_t *rects, int n_rects, int lap0, int lap1, orbx_keypoint *kps, uint8_t *desc,
                       int cap, int32_t *n_out, int32_t *mono_index) {
    return host_batch_checked(ex, images, batch, rows, cols, step, rects, n_rects, lap0, lap1, kps, desc, cap, n_out, mono_index, false);
}

int orbx_copy_only_batch(orbx_extractor *ex, const uint8_t *const *images, int batch, int rows, int cols, size_t step,
                         orbx_keypoint *kps, uint8_t *desc, int cap) {
    if (!ex || batch <= 0) return ORBX_ERR_ARG;
    std::vector<int32_t> n(batch), m(batch);
    return host_batch_checked(ex, images, batch, rows, cols, step, nullptr, 0, 0, 0, kps, desc, cap, n.data(), m.data(), true);
}

// Layout of the single-frame output block (device and pinned host copies): [n_out, mono_index, pad to 256 B][cap keypoints]
// [cap descriptors]
static inline size_t single_kps_off() { return 256; }
static inline size_t single_desc_off(int cap) { return 256 + (((size_t)cap * sizeof(orbx_keypoint) + 255) & ~(size_t)255); }
static inline size_t single_bytes(int cap) { return single_desc_off(cap) + (size_t)cap * 32; }

// One frame, latency path (what Frame::ExtractORB calls once per image, src/Frame.cc:420-427): the image goes through a pinned
// staging buffer, all outputs come back in ONE copy, there is ONE stream synchronisation, and from the third call with the same
// (rows, cols, cap) the copy-in, the ~17 kernels and the copy-out replay as one CUDA graph.
static int extract_one(orbx_extractor *ex, const uint8_t *image, int rows, int cols, size_t step, const int32_t *rects, int n_rects,
                       int lap0, int lap1, orbx_keypoint *kps, uint8_t *desc, int cap, int *n_out, int *mono_index) {
    int rc = prepare(ex, rows, cols, rects, n_rects, lap0, lap1, 1);
    if (rc) return rc;
    const OrbxGeom &G = ex->geom;
    cudaStream_t s = ex->stream;
    const size_t inBytes = (size_t)rows * cols, outBytes = single_bytes(cap);
    if (inBytes > ex->hInCap || !ex->h_in) {
        CUDA_TRY(ex, cudaStreamSynchronize(s));
        if (ex->h_in) cudaFreeHost(ex->h_in);
        ex->h_in = nullptr; ex->hInCap = 0;
        CUDA_TRY(ex, cudaHostAlloc((void **)&ex->h_in, (size_t)ex->maxW * ex->maxH, cudaHostAllocDefault));
        ex->hInCap = (size_t)ex->maxW * ex->maxH;
    }
    if (outBytes > ex->singleCap || !ex->d_single) {
        CUDA_TRY(ex, cudaStreamSynchronize(s));
        if (ex->oneExec) { cudaGraphExecDestroy(ex->oneExec); ex->oneExec = nullptr; }
        if (ex->d_single) cudaFree(ex->d_single);
    if (ex->d_sm) cudaFree(ex->d_sm);
        if (ex->h_single) cudaFreeHost(ex->h_single);
        ex->d_single = ex->h_single = nullptr; ex->singleCap = 0;
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_single, outBytes));
        CUDA_TRY(ex, cudaHostAlloc((void **)&ex->h_single, outBytes, cudaHostAllocDefault));
        ex->singleCap = outBytes;
    }
    if (rows != ex->oneRows || cols != ex->oneCols || cap != ex->oneCap) {
        if (ex->oneExec) { CUDA_TRY(ex, cudaStreamSynchronize(s)); cudaGraphExecDestroy(ex->oneExec); ex->oneExec = nullptr; }
        ex->oneRows = rows; ex->oneCols = cols; ex->oneCap = cap; ex->oneWarm = 0;
    }
    // the previous call's graph has completed (every call ends with a synchronisation), so the staging buffer is free
    if (step == (size_t)cols) memcpy(ex->h_in, image, inBytes);
    else for (int y = 0; y < rows; ++y) memcpy(ex->h_in + (size_t)y * cols, image + (size_t)y * step, cols);

    uint8_t *lvl0 = ex->d_pyr + G.lv[0].off;
    int *dN = (int *)ex->d_single, *dM = dN + 1;
    orbx_keypoint *dK = (orbx_keypoint *)(ex->d_single + single_kps_off());
    uint8_t *dD = ex->d_single + single_desc_off(cap);
    auto enqueue = [&]() -> int {
        CUDA_TRY(ex, cudaMemcpy2DAsync(lvl0, (size_t)G.lv[0].pitch, ex->h_in, (size_t)cols, (size_t)cols, rows, cudaMemcpyHostToDevice, s));
        ex->lastBatch = 0; ex->denseSlots.clear();
        int r = run_pipeline(ex, lvl0, G.frameBytes, G.lv[0].pitch, 1, dK, dD, cap, dN, dM);
        if (r) return r;
        CUDA_TRY(ex, cudaMemcpyAsync(ex->h_single, ex->d_single, outBytes, cudaMemcpyDeviceToHost, s));
        return ORBX_OK;
    };
    ex->lastIn0Internal = true;
    if (ex->oneExec) {
        CUDA_TRY(ex, cudaGraphLaunch(ex->oneExec, s));
        ex->launches += ex->oneLaunches;
        ex->lastBatch = 1; ex->denseSlots.assign(1, 0);
    } else if (ex->useGraph && !ex->profiling && ex->oneWarm >= 2) {
        // the first two calls ran eagerly (they also set the kernels' shared-memory attributes); capture this one
        const long long before = ex->launches;
        cudaGraph_t graph = nullptr;
        CUDA_TRY(ex, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        rc = enqueue();
        cudaError_t e = cudaStreamEndCapture(s, &graph);
        if (rc == ORBX_OK && e == cudaSuccess && graph && cudaGraphInstantiate(&ex->oneExec, graph, 0) == cudaSuccess) {
            ex->oneLaunches = ex->launches - before;
            CUDA_TRY(ex, cudaGraphLaunch(ex->oneExec, s));
        } else {                                   // capture refused: keep launching eagerly (same kernels, more launch overhead)
            cudaGetLastError();
            ex->oneExec = nullptr; ex->useGraph = false;
            ex->launches = before;
            rc = enqueue();
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        }
        if (graph) cudaGraphDestroy(graph);
    } else {
        rc = enqueue();
        if (rc) return rc;
        ++ex->oneWarm;
    }
    CUDA_TRY(ex, cudaStreamSynchronize(s));
    const int n = ((const int *)ex->h_single)[0], mono = ((const int *)ex->h_single)[1];
    if (n_out) *n_out = n;
    if (mono_index) *mono_index = mono;
    if (n > cap) { ex->err = "keypoint capacity too small (see n_out)"; return ORBX_ERR_CAPACITY; }
    if (n > 0) {
        memcpy(kps, ex->h_single + single_kps_off(), (size_t)n * sizeof(orbx_keypoint));
        memcpy(desc, ex->h_single + single_desc_off(cap), (size_t)n * 32);
    }
    return ORBX_OK;
}

int orbx_extract(orbx_extractor *ex, const uint8_t *image, int rows, int cols, size_t step, const int32_t *rects,
                 int n_rects, int lap0, int lap1, orbx_keypoint *kps, uint8_t *desc, int cap, int *n_out,
                 int *mono_index) {
    if (!ex) return ORBX_ERR_ARG;
    if (n_out) *n_out = 0;
    if (mono_index) *mono_index = -1;
    if (!image || rows <= 0 || cols <= 0) { ex->err = "empty image"; return ORBX_EMPTY; }
    if (!kps || !desc || cap <= 0) { ex->err = "orbx_extract: null output or cap <= 0"; return ORBX_ERR_ARG; }
    OrbxDeviceGuard dg_(ex->device);
    if (dg_.status != cudaSuccess) { ex->err = std::string("cudaSetDevice: ") + cudaGetErrorString(dg_.status); return ORBX_ERR_CUDA; }
    const int rc = extract_one(ex, image, rows, cols, step, rects, n_rects, lap0, lap1, kps, desc, cap, n_out, mono_index);
    if (rc != ORBX_OK && rc != ORBX_ERR_CAPACITY) {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(ex->stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) { cudaGraph_t g = nullptr; cudaStreamEndCapture(ex->stream, &g); if (g) cudaGraphDestroy(g); }
        drain_streams(ex);
    }
    return rc;
}

int orbx_stereo_matches(orbx_extractor *ex_left, orbx_extractor *ex_right, const orbx_keypoint *kps_left, const uint8_t *desc_left, int n_left,
                        const orbx_keypoint *kps_right, const uint8_t *desc_right, int n_right, float mbf, float mb, float *mvu_right, float *mv_depth,
                        int32_t *n_stereo) {
    if (!ex_left || !ex_right) return ORBX_ERR_ARG;
    orbx_extractor *ex = ex_left;
    if (n_left < 0 || n_right < 0 || !n_stereo || (n_left > 0 && (!kps_left || !desc_left || !mvu_right || !mv_depth)) || (n_right > 0 && (!kps_right || !desc_right))) {
        ex->err = "orbx_stereo_matches: bad argument";
        return ORBX_ERR_ARG;
    }
    *n_stereo = 0;
    for (int i = 0; i < n_left; ++i) { mvu_right[i] = -1.f; mv_depth[i] = -1.f; }
    if (n_left == 0 || n_right == 0) return ORBX_OK;
    if (ex_left->device != ex_right->device || ex_left->nlevels != ex_right->nlevels || ex_left->curRows != ex_right->curRows || ex_left->curCols != ex_right->curCols ||
        ex_left->lastBatch < 1 || ex_right->lastBatch < 1 || !ex_left->lastIn0Internal || !ex_right->lastIn0Internal) {
        ex->err = "orbx_stereo_matches: both extractors must hold the pyramid of a host-image call of the same size on the same device";
        return ORBX_ERR_ARG;
    }
    for (int i = 0; i < n_left; ++i) if (kps_left[i].octave < 0 || kps_left[i].octave >= ex->nlevels) { ex->err = "orbx_stereo_matches: octave out of range"; return ORBX_ERR_ARG; }
    for (int i = 0; i < n_right; ++i) if (kps_right[i].octave < 0 || kps_right[i].octave >= ex->nlevels) { ex->err = "orbx_stereo_matches: octave out of range"; return ORBX_ERR_ARG; }
    OrbxDeviceGuard dg_(ex->device);
    CUDA_TRY(ex, dg_.status);
    CUDA_TRY(ex, cudaStreamSynchronize(ex_right->stream));       // its pyramid is read on the left handle's stream
    const size_t a256 = 255;
    const size_t kLB = ((size_t)n_left * sizeof(orbx_keypoint) + a256) & ~a256, kRB = ((size_t)n_right * sizeof(orbx_keypoint) + a256) & ~a256;
    const size_t dLB = ((size_t)n_left * 32 + a256) & ~a256, dRB = ((size_t)n_right * 32 + a256) & ~a256, pB = ((size_t)n_right * 16 + a256) & ~a256;
    const size_t oB = ((size_t)n_left * 4 + a256) & ~a256;
    const size_t need = kLB + kRB + dLB + dRB + pB + 3 * oB + 256;
    if (need > ex->smCap || !ex->d_sm) {
        CUDA_TRY(ex, cudaStreamSynchronize(ex->stream));
        if (ex->d_sm) cudaFree(ex->d_sm);
        ex->d_sm = nullptr; ex->smCap = 0;
        CUDA_TRY(ex, cudaMalloc((void **)&ex->d_sm, need));
        ex->smCap = need;
    }
    uint8_t *p = ex->d_sm;
    orbx_keypoint *dkL = (orbx_keypoint *)p; p += kLB;
    orbx_keypoint *dkR = (orbx_keypoint *)p; p += kRB;
    uint8_t *ddL = p; p += dLB;
    uint8_t *ddR = p; p += dRB;
    int4 *dprep = (int4 *)p; p += pB;
    float *dur = (float *)p; p += oB;
    float *ddp = (float *)p; p += oB;
    int *dsad = (int *)p; p += oB;
    int *dn = (int *)p;
    cudaStream_t s = ex->stream;
    CUDA_TRY(ex, cudaMemcpyAsync(dkL, kps_left, (size_t)n_left * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, s));
    CUDA_TRY(ex, cudaMemcpyAsync(dkR, kps_right, (size_t)n_right * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, s));
    CUDA_TRY(ex, cudaMemcpyAsync(ddL, desc_left, (size_t)n_left * 32, cudaMemcpyHostToDevice, s));
    CUDA_TRY(ex, cudaMemcpyAsync(ddR, desc_right, (size_t)n_right * 32, cudaMemcpyHostToDevice, s));
    SmLevels lv;
    memset(&lv, 0, sizeof(lv));
    const OrbxGeom &GL = ex_left->geom, &GR = ex_right->geom;
    lv.nlevels = ex->nlevels; lv.nRows = GL.lv[0].h;
    for (int l = 0; l < ex->nlevels; ++l) {
        lv.L[l] = ex_left->d_pyr + GL.lv[l].off; lv.R[l] = ex_right->d_pyr + GR.lv[l].off;     // frame 0 of each handle's last call
        lv.pitchL[l] = GL.lv[l].pitch; lv.pitchR[l] = GR.lv[l].pitch; lv.w[l] = GL.lv[l].w; lv.h[l] = GL.lv[l].h;
        lv.sf[l] = ex->sf[l]; lv.inv[l] = ex->inv[l];
    }
    k_sm_prep<<<(n_right + 255) / 256, 256, 0, s>>>(dkR, n_right, lv, dprep);
    k_sm_match<8><<<(n_left + 7) / 8, 256, 0, s>>>(dkL, (const uint4 *)ddL, n_left, dprep, (const uint4 *)ddR, n_right, lv, mbf, mb, dur, ddp, dsad);
    k_sm_tail<<<1, 256, 0, s>>>(dsad, n_left, dur, ddp, dn);
    ex->launches += 3;
    CUDA_TRY(ex, cudaGetLastError());
    CUDA_TRY(ex, cudaMemcpyAsync(mvu_right, dur, (size_t)n_left * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ex, cudaMemcpyAsync(mv_depth, ddp, (size_t)n_left * 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ex, cudaMemcpyAsync(n_stereo, dn, 4, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(ex, cudaStreamSynchronize(s));
    return ORBX_OK;
}

int orbx_sync(orbx_extractor *ex) {
    if (!ex) return ORBX_ERR_ARG;
    OrbxDeviceGuard dg_(ex->device);
    CUDA_TRY(ex, cudaStreamSynchronize(ex->stream));
    return ORBX_OK;
}
void *orbx_stream(orbx_extractor *ex) { return ex ? (void *)ex->stream : nullptr; }

// test hook: how many (frame, level) pairs of the last batch call fell back from the histogram quadtree kernel to
// the general one (deep trees); -1 when the histogram kernel is not in use
int orbx_debug_deep_count(orbx_extractor *ex) {
    if (!ex || !ex->d_deep) return ORBX_ERR_ARG;
    if (!ex->useHistQuadtree) return -1;
    OrbxDeviceGuard dg_(ex->device);
    if (dg_.status != cudaSuccess || cudaStreamSynchronize(ex->stream) != cudaSuccess) return ORBX_ERR_CUDA;
    std::vector<int> f((size_t)std::max(ex->lastBatch, 1) * ex->nlevels);
    if (cudaMemcpy(f.data(), ex->d_deep, f.size() * sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return ORBX_ERR_CUDA;
    int n = 0;
    for (int v : f) n += v != 0;
    return n;
}


// test hook: cells of the last batch call that the two-phase FAST kernel handed to the single-phase kernel
int orbx_debug_dense_count(orbx_extractor *ex) {
    if (!ex || !ex->d_dense) return ORBX_ERR_ARG;
    if (ex->fastV1) return -1;
    OrbxDeviceGuard dg_(ex->device);
    if (dg_.status != cudaSuccess || cudaStreamSynchronize(ex->stream) != cudaSuccess) return ORBX_ERR_CUDA;
    int n = 0;
    for (int slot : ex->denseSlots) {
        int v = 0;
        if (cudaMemcpy(&v, ex->d_dense + slot, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return ORBX_ERR_CUDA;
        n += v;
    }
    return n;
}

int orbx_set_profiling(orbx_extractor *ex, int on) {
    if (!ex) return ORBX_ERR_ARG;
    OrbxDeviceGuard dg_(ex->device); CUDA_TRY(ex, dg_.status);
    if (on && !ex->ev[0])
        for (int i = 0; i < 7; ++i) CUDA_TRY(ex, cudaEventCreate(&ex->ev[i]));
    if (!on) { int rc = harvest_stage_times(ex); if (rc) return rc; }
    ex->profiling = on != 0;
    return ORBX_OK;
}

int orbx_get_stage_ms(orbx_extractor *ex, double *ms6, long long *calls) {
    if (!ex || !ms6) return ORBX_ERR_ARG;
    int rc = harvest_stage_times(ex);
    if (rc) return rc;
    for (int i = 0; i < 6; ++i) ms6[i] = ex->stageMs[i];
    if (calls) *calls = ex->stageCalls;
    return ORBX_OK;
}

int orbx_reset_stage_ms(orbx_extractor *ex) {
    if (!ex) return ORBX_ERR_ARG;
    int rc = harvest_stage_times(ex);
    for (int i = 0; i < 6; ++i) ex->stageMs[i] = 0;
    ex->stageCalls = 0;
    return rc;
}
long long orbx_launch_count(const orbx_extractor *ex) { return ex ? ex->launches : 0; }

int orbx_level_size(const orbx_extractor *ex, int level, int *w, int *h) {
    if (!ex || level < 0 || level >= ex->nlevels || ex->curRows < 0) return ORBX_ERR_ARG;
    if (w) *w = ex->geom.lv[level].w;
    if (h) *h = ex->geom.lv[level].h;
    return ORBX_OK;
}

static int copy_plane(orbx_extractor *ex, const uint8_t *d_src, int pitch, int w, int h, int padded, uint8_t *dst,
                      size_t dst_step) {
    CUDA_TRY(ex, cudaStreamSynchronize(ex->stream));
    if (!padded) {
        CUDA_TRY(ex, cudaMemcpy2D(dst, dst_step, d_src, pitch, w, h, cudaMemcpyDeviceToHost));
        return ORBX_OK;
    }
    // copy the level into the middle of the padded plane, then reflect (BORDER_REFLECT_101, :1224-1230)
    const int E = ORBX_EDGE;
    CUDA_TRY(ex, cudaMemcpy2D(dst + (size_t)E * dst_step + E, dst_step, d_src, pitch, w, h, cudaMemcpyDeviceToHost));
    auto refl = [](int i, int n) { if (n == 1) return 0; while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i; return i; };
    for (int y = 0; y < h; ++y) {
        uint8_t *row = dst + (size_t)(y + E) * dst_step + E;
        for (int x = -E; x < 0; ++x) row[x] = row[refl(x, w)];
        for (int x = w; x < w + E; ++x) row[x] = row[refl(x, w)];
    }
    for (int y = -E; y < h + E; ++y) {
        if (y >= 0 && y < h) continue;
        memcpy(dst + (size_t)(y + E) * dst_step, dst + (size_t)(refl(y, h) + E) * dst_step, w + 2 * E);
    }
    return ORBX_OK;
}

int orbx_get_pyramid(orbx_extractor *ex, int frame, int level, int padded, uint8_t *dst, size_t dst_step) {
    if (!ex || !dst || level < 0 || level >= ex->nlevels || frame < 0 || frame >= ex->lastBatch) return ORBX_ERR_ARG;
    const OrbxGeom &G = ex->geom;
    if (level == 0 && !ex->lastIn0Internal) { ex->err = "level 0 of a device-resident batch is the caller's own buffer"; return ORBX_ERR_ARG; }
    OrbxDeviceGuard dg_(ex->device);
    return copy_plane(ex, ex->d_pyr + (size_t)frame * G.frameBytes + G.lv[level].off, G.lv[level].pitch, G.lv[level].w,
                      G.lv[level].h, padded, dst, dst_step);
}

int orbx_get_blurred(orbx_extractor *ex, int frame, int level, uint8_t *dst, size_t dst_step) {
    if (!ex || !dst || level < 0 || level >= ex->nlevels || frame < 0 || frame >= ex->lastBatch) return ORBX_ERR_ARG;
    const OrbxGeom &G = ex->geom;
    OrbxDeviceGuard dg_(ex->device);
    return copy_plane(ex, ex->d_blur + (size_t)frame * G.frameBytes + G.lv[level].off, G.lv[level].pitch, G.lv[level].w,
                      G.lv[level].h, 0, dst, dst_step);
}

int orbx_get_candidates(orbx_extractor *ex, int frame, int level, orbx_keypoint *out, int cap) {
    if (!ex || level < 0 || level >= ex->nlevels || frame < 0 || frame >= ex->lastBatch) return ORBX_ERR_ARG;
    const OrbxGeom &G = ex->geom;
    const OrbxLevel &V = G.lv[level];
    OrbxDeviceGuard dg_(ex->device);
    CUDA_TRY(ex, cudaStreamSynchronize(ex->stream));
    std::vector<int> cnt(std::max(V.nCells, 1));
    if (V.nCells) CUDA_TRY(ex, cudaMemcpy(cnt.data(), ex->d_cellCnt + (size_t)frame * G.nCellsTotal + V.cellBase, V.nCells * sizeof(int), cudaMemcpyDeviceToHost));
    size_t total = 0;
    for (int c = 0; c < V.nCells; ++c) total += cnt[c];
    std::vector<uint32_t> node(std::max<size_t>(total, 1));
    std::vector<float2> xy(std::max<size_t>(total, 1));
    if (total) {  // dense per-level arrays, indexed by candidate order (see k_quadtree)
        const size_t o = (size_t)frame * G.slotsTotal + V.slotBase;
        CUDA_TRY(ex, cudaMemcpy(node.data(), ex->d_ptNode + o, total * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(ex, cudaMemcpy(xy.data(), ex->d_ptXY + o, total * 8, cudaMemcpyDeviceToHost));
    }
    int n = 0;
    for (size_t s = 0; s < total; ++s) {
        if (node[s] == ORBX_NODE_ERASED) continue;
        if (out && n < cap) {
            orbx_keypoint kp;
            kp.x = xy[s].x; kp.y = xy[s].y; kp.size = 7.f; kp.angle = -1.f;
            kp.response = (float)((node[s] >> 16) & 0xff); kp.octave = 0; kp.class_id = -1;
            out[n] = kp;
        }
        ++n;
    }
    return n;
}

int orbx_get_selected(orbx_extractor *ex, int frame, int level, orbx_keypoint *out, int cap) {
    if (!ex || level < 0 || level >= ex->nlevels || frame < 0 || frame >= ex->lastBatch) return ORBX_ERR_ARG;
    const OrbxGeom &G = ex->geom;
    const OrbxLevel &V = G.lv[level];
    OrbxDeviceGuard dg_(ex->device);
    CUDA_TRY(ex, cudaStreamSynchronize(ex->stream));
    int n = 0;
    CUDA_TRY(ex, cudaMemcpy(&n, ex->d_selCnt + frame * G.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<float4> s(std::max(n, 1));
    if (n) CUDA_TRY(ex, cudaMemcpy(s.data(), ex->d_sel + (size_t)frame * G.selTotal + V.selBase, n * sizeof(float4), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n && i < cap && out; ++i) {
        orbx_keypoint kp;
        kp.x = s[i].x; kp.y = s[i].y; kp.size = (float)V.patch_size; kp.angle = -1.f; kp.response = s[i].z;
        kp.octave = level; kp.class_id = -1;
        out[i] = kp;
    }
    return n;
}

void *orbx_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void *orbx_host_alloc_wc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void orbx_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- test hooks ----
void orbx_debug_sort_nodes(const int32_t *sizes, const int32_t *ulx, int n, int32_t *perm_out) {
    std::vector<orbx_sort::elem_t> a(n);
    for (int i = 0; i < n; ++i)
        a[i] = ((((unsigned long long)(unsigned)sizes[i] << 16) | (unsigned short)ulx[i]) << orbx_sort::kPayloadBits) | (unsigned)i;
    // evaluated the way the quadtree kernel does it: partitioning phase, then a stable sort by key
    orbx_sort::introsort_loop_only(a.data(), n);
    std::stable_sort(a.begin(), a.end(), [](orbx_sort::elem_t x, orbx_sort::elem_t y) { return orbx_sort::less(x, y); });
    for (int i = 0; i < n; ++i) perm_out[i] = (int32_t)(a[i] & ((1ull << orbx_sort::kPayloadBits) - 1));
}

int orbx_debug_sort_nodes_device(int device, const int32_t *sizes, const int32_t *ulx, int n, int32_t *perm_out) {
    OrbxDeviceGuard dg_(device);
    if (dg_.status != cudaSuccess) return ORBX_ERR_CUDA;
    std::vector<orbx_sort::elem_t> a(std::max(n, 1));
    for (int i = 0; i < n; ++i)
        a[i] = ((((unsigned long long)(unsigned)sizes[i] << 16) | (unsigned short)ulx[i]) << orbx_sort::kPayloadBits) | (unsigned)i;
    orbx_sort::elem_t *d = nullptr;
    if (cudaMalloc((void **)&d, a.size() * 16) != cudaSuccess) return ORBX_ERR_CUDA;
    cudaMemcpy(d, a.data(), a.size() * 8, cudaMemcpyHostToDevice);
    k_dbg_sort<<<1, 32>>>(d, d + a.size(), n);
    cudaError_t e = cudaMemcpy(a.data(), d, a.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return ORBX_ERR_CUDA;
    for (int i = 0; i < n; ++i) perm_out[i] = (int32_t)(a[i] & ((1ull << orbx_sort::kPayloadBits) - 1));
    return ORBX_OK;
}

int orbx_debug_sincos_device(int device, const float *angles, int n, float *sin_out, float *cos_out) {
    OrbxDeviceGuard dg_(device);
    if (dg_.status != cudaSuccess) return ORBX_ERR_CUDA;
    float *d = nullptr;
    if (cudaMalloc((void **)&d, (size_t)n * 12) != cudaSuccess) return ORBX_ERR_CUDA;
    cudaMemcpy(d, angles, (size_t)n * 4, cudaMemcpyHostToDevice);
    k_dbg_sincos<<<(n + 255) / 256, 256>>>(d, n, d + n, d + 2 * (size_t)n);
    cudaMemcpy(sin_out, d + n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaMemcpy(cos_out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? ORBX_OK : ORBX_ERR_CUDA;
}

int orbx_debug_atan2_device(int device, const float *y, const float *x, int n, float *deg_out) {
    OrbxDeviceGuard dg_(device);
    if (dg_.status != cudaSuccess) return ORBX_ERR_CUDA;
    float *d = nullptr;
    if (cudaMalloc((void **)&d, (size_t)n * 12) != cudaSuccess) return ORBX_ERR_CUDA;
    cudaMemcpy(d, y, (size_t)n * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d + n, x, (size_t)n * 4, cudaMemcpyHostToDevice);
    k_dbg_atan2<<<(n + 255) / 256, 256>>>(d, d + n, n, d + 2 * (size_t)n);
    cudaError_t e = cudaMemcpy(deg_out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? ORBX_OK : ORBX_ERR_CUDA;
}

}  // extern "C"
