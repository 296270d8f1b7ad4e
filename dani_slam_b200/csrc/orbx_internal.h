// orbx_internal.h — shared declarations of liborbx (not part of the C ABI).
#ifndef ORBX_INTERNAL_H
#define ORBX_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/orbx.h"

#define ORBX_MAX_LEVELS 16
#define ORBX_EDGE 19        // EDGE_THRESHOLD (ORBextractor.cc:73)
#define ORBX_BORDER 16      // minBorderX/Y = EDGE_THRESHOLD-3 (:789-790)
#define ORBX_HALF_PATCH 15  // HALF_PATCH_SIZE (:72)
#define ORBX_PATCH 31       // PATCH_SIZE (:71)
#define ORBX_NODE_ERASED 0xFFFFFFFFu
#define ORBX_MAX_RECTS 64

// Per-level geometry, computed on the host exactly as the reference does (fp32 where it does).
struct OrbxLevel {
    int w, h, pitch;            // unpadded level size, bytes per row in HBM
    long long off;              // byte offset of this level inside a frame's pyramid block
    float sf, inv;              // mvScaleFactor / mvInvScaleFactor
    int quota;                  // mnFeaturesPerLevel
    int patch_size;             // (int)(31*sf)
    int maxBX, maxBY;           // w-16, h-16
    int nCols, nRows, wCell, hCell, nColsOK, nRowsOK;  // cell grid (:794-822)
    int cellBase, nCells;       // cells of this level inside the frame's cell table (OK cells only)
    int slotCap;                // candidate slots per cell
    long long slotBase;         // first slot of this level inside the frame's slot array
    int nIni;                   // quadtree roots (:558)
    float hX;                   // (:560)
    int nodeCap;                // quadtree node capacity
    int lutX, lutY, lutW, lutH; // quadtree path tables: offsets into the u16 table array and their extents (lutW = 0: none)
    int selBase, selCap;        // selected-keypoint list of this level inside the frame's list
};

struct OrbxCell {               // one FAST cell (ROI of the level)
    short level, x0, y0, cw, ch;
    short cx, cy;               // cell column j / row i
    int seq;                    // processing order index inside the level (row-major over OK cells)
    long long slot;             // first slot (inside the frame's slot array)
};

struct OrbxGeom {               // device-side copy of everything the kernels need
    int nlevels, rows, cols;
    int iniTh, minTh, lowTh;
    int nCellsTotal, selTotal;
    long long frameBytes, slotsTotal;
    int maxCw, maxCh;
    int lap0, lap1;
    int nRects;
    int rects[4 * ORBX_MAX_RECTS];
    int umax[ORBX_HALF_PATCH + 1];
    OrbxLevel lv[ORBX_MAX_LEVELS];
};

struct OrbxWork {               // one selected keypoint, handed from k_assemble to k_orient_desc
    int level, cx, cy, out;
};

// Every entry point runs on its handle's device and leaves the calling thread's current device as it found it
// (host applications such as torch keep their own notion of the current device).
struct OrbxDeviceGuard {
    int prev = -1;
    cudaError_t status = cudaSuccess;
    explicit OrbxDeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) status = cudaSetDevice(dev);
        if (prev == dev) prev = -1;                 // nothing to restore
    }
    ~OrbxDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    OrbxDeviceGuard(const OrbxDeviceGuard &) = delete;
    OrbxDeviceGuard &operator=(const OrbxDeviceGuard &) = delete;
};

inline int orbx_align_up(int v, int a) { return (v + a - 1) / a * a; }
inline long long orbx_align_up_ll(long long v, long long a) { return (v + a - 1) / a * a; }

#endif
