// tma_probe — isolates the TMA box load used by the FAST / pyramid kernels: rank-3 u8 tensor (x bytes, y rows, frame), box W×H×1,
// one warp, lane 0 issues cp.async.bulk.tensor.3d, the warp waits on the mbarrier, the result is compared with direct loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu     Run: ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, int boxW, int boxH, uint8_t *out, int variant, const uint8_t *raw, const CUtensorMap *gmap) {
    const CUtensorMap *mp = (variant & 16) ? gmap : &map;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + ((boxW * boxH + 7) & ~7));
    const int lane = threadIdx.x;
    if (variant & 32) {       // convergent: every lane runs this, elect.sync picks the issuer (the CUTLASS pattern)
        if (lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile(
            "{\n"
            ".reg .pred P;\n"
            "elect.sync _|P, 0xffffffff;\n"
            "@P mbarrier.arrive.expect_tx.shared::cta.b64 _, [%5], %6;\n"
            "@P cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
            "}\n" ::"r"(smem_u32(smem)), "l"(mp), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "r"(boxW * boxH) : "memory");
    } else
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        if (variant & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(boxW * boxH) : "memory");
        if (variant & 8) {            // plain bulk copy of boxW*boxH bytes from the start of the tensor: tests the mbarrier machinery alone
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem)), "l"(raw), "r"(boxW * boxH), "r"(smem_u32(bar)) : "memory");
        } else if (variant & 4) {     // rank-2 map
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_u32(smem)), "l"(mp), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
        } else if (variant & 2) {     // no .tile qualifier
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(smem)), "l"(mp), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
        } else {
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(smem)), "l"(mp), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
        }
    }
    __syncwarp();
    asm volatile(
        "{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = lane; i < boxW * boxH; i += 32) out[i] = smem[i];
}

typedef CUresult (*PFN_enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                            CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_canon(const __grid_constant__ CUtensorMap map, uint16_t *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 64 * 32 * 2);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(64 * 32 * 2) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(smem)), "l"(&map), "r"(64), "r"(32), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(smem)[i];
}

int main(int argc, char **argv) {
    if (argc > 1 && atoi(argv[1]) == 100) {
        uint16_t *d = nullptr, *o = nullptr;
        cudaMalloc(&d, 1024 * 256 * 2); cudaMalloc(&o, 64 * 32 * 2);
        cudaMemset(d, 0x11, 1024 * 256 * 2);
        void *f = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
        typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        CUtensorMap map;
        const cuuint64_t dims[2] = {1024, 256}, strides[1] = {2048};
        const cuuint32_t box[2] = {64, 32}, es[2] = {1, 1};
        CUresult r = ((PFN)f)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("canonical fp16 2D swizzle128: encode rc=%d", (int)r);
        k_canon<<<1, 128, 64 * 32 * 2 + 1024>>>(map, o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel: %s\n", cudaGetErrorString(e));
        return 0;
    }
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    const int W = argc > 2 ? atoi(argv[2]) : 752, H = 480, P = 768, B = 2;
    const int l2p = argc > 3 ? atoi(argv[3]) : 0, dtype = argc > 4 ? atoi(argv[4]) : 0, es_ = dtype == 7 ? 4 : 1;   // dtype 0 = UINT8, 7 = FLOAT32
    std::vector<uint8_t> h((size_t)P * H * B + 4096);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t *d = nullptr, *dout = nullptr;
    cudaMalloc(&d, h.size()); cudaMalloc(&dout, 65536);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *f = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f) { printf("no encoder\n"); return 1; }
    PFN_enc enc = (PFN_enc)f;
    struct Case { int boxW, boxH, x, y, z; long long planeStride; int variant; } cases[] = {
        {48, 44, 13, 20, 1, (long long)P * H, 0}, {48, 44, 13, 20, 1, (long long)P * H, 1}, {176, 21, 100, 7, 0, (long long)P * H, 0},
        {64, 57, 700, 440, 1, (long long)P * H, 0}, {48, 44, 13, 20, 1, (long long)P * H + 1264 * 16, 0},
        {48, 44, 13, 20, 1, (long long)P * H, 2}, {48, 44, 13, 20, 0, (long long)P * H, 4}, {48, 44, 0, 0, 0, (long long)P * H, 8}, {48, 44, 0, 0, 0, (long long)P * H, 9}, {48, 44, 13, 20, 0, (long long)P * H, 20}, {48, 44, 13, 20, 1, (long long)P * H, 16}, {48, 44, 13, 20, 1, (long long)P * H, 32}, {48, 44, 13, 20, 1, (long long)P * H, 48},
        {48, 44, 16, 20, 1, (long long)P * H, 0}, {48, 44, 0, 0, 0, (long long)P * H, 0}, {64, 32, 16, 20, 1, (long long)P * H, 0}, {48, 44, 13, 20, 1, (long long)P * H, 64},
        {48, 44, 4, 20, 1, (long long)P * H, 0}, {48, 44, 8, 20, 1, (long long)P * H, 0}, {48, 44, 12, 21, 1, (long long)P * H, 0}, {48, 44, 2, 20, 1, (long long)P * H, 0}};
    int ci = -1;
    for (const Case &c : cases) {
        ++ci;
        if (only >= 0 && ci != only) continue;
        CUtensorMap map;
        const cuuint64_t dims[3] = {(cuuint64_t)(W / es_), (cuuint64_t)H, (cuuint64_t)B}, strides[2] = {(cuuint64_t)P, (cuuint64_t)c.planeStride};
        const cuuint32_t box[3] = {(cuuint32_t)(c.boxW / es_), (cuuint32_t)c.boxH, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&map, (CUtensorMapDataType)dtype, (c.variant & 4) ? 2 : 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         (CUtensorMapL2promotion)((c.variant & 64) ? 2 : l2p), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box %dx%d at (%d,%d,%d) planeStride %lld variant %d: encode rc=%d", c.boxW, c.boxH, c.x, c.y, c.z, c.planeStride, c.variant, (int)r);
        if (r != CUDA_SUCCESS) { printf("\n"); continue; }
        { const unsigned long long *mw = (const unsigned long long *)&map; printf("\n  map:"); for (int i = 0; i < 16; ++i) printf(" %llx", mw[i]); printf("\n"); }
        cudaMemset(dout, 0xAB, 65536);
        CUtensorMap *gmap = nullptr; cudaMalloc(&gmap, 128); cudaMemcpy(gmap, &map, 128, cudaMemcpyHostToDevice);
        k_probe<<<1, 32, c.boxW * c.boxH + 64>>>(map, c.x / es_, c.y, c.z, c.boxW, c.boxH, dout, c.variant, d, gmap);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  KERNEL ERROR: %s\n", cudaGetErrorString(e)); return 2; }
        std::vector<uint8_t> o((size_t)c.boxW * c.boxH);
        cudaMemcpy(o.data(), dout, o.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < c.boxH; ++yy)
            for (int xx = 0; xx < c.boxW; ++xx) {
                const int gx = c.x + xx, gy = c.y + yy;
                const uint8_t want = (gx < W && gy < H) ? h[(size_t)c.z * c.planeStride + (size_t)gy * P + gx] : 0;
                bad += o[(size_t)yy * c.boxW + xx] != want;
            }
        if (c.variant & 8) { bad = 0; for (int i = 0; i < c.boxW * c.boxH; ++i) bad += o[i] != h[i]; }
        printf("  mismatches %d\n", bad);
    }
    return 0;
}
