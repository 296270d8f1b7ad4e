// orbx_knn_tc.cu — brute-force Hamming kNN (k = 2) on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// For bit vectors q, d ∈ {0,1}^256:  hamming(q, d) = |q| + |d| − 2·(q · d).  The dot products of a tile of 128 queries with a
// tile of 256 database rows are ONE 128×256×256 integer GEMM: every bit becomes one unsigned operand byte in shared memory
// (K-major, 128-byte swizzle — the canonical UMMA operand layout).  The order of the 256 terms of a dot product is free, so
// operand word j (j = 0…7) of a 32-bit descriptor word x holds the bits j, j+8, j+16, j+24 IN PLACE: a database byte is
// x & (1 << j) ∈ {0, 2^j} — one AND per four operand bytes, no shifts — and the query byte of the same term is scaled the other
// way, bit << (7 − j), so that every product is 128·(q bit)·(d bit) and the accumulator is exactly 128·(q · d).
// `tcgen05.mma.kind::i8` (unsigned × unsigned) accumulates in int32 in tensor memory,
// and the epilogue reads the accumulators back with `tcgen05.ld`, forms packed (distance << 23 | row) keys and keeps the two
// smallest per query — the same keys, tie rule (lower row first) and per-chunk partial format as the POPC kernel in
// orbx_match.cu, whose merge kernel finishes the job.  Results are bit-identical to the POPC path (tests/test_match_gpu.py).
//
// One CTA = one tile of 128 queries × one chunk of database rows; warp roles (21 warps):
//   warp 0        allocates tensor memory; lane 0 issues the MMAs (8 per database tile: K = 8 × 32 bytes) and commits them
//   warps 1-4     producers: two database rows per thread and tile (32 B each, coalesced, prefetched one tile ahead), expanded to the
//                 swizzled operand tile of the free stage (64 ANDs + 16 16-byte stores per row); the row's popcount goes into the
//                 per-column key base
//   warps 5-20    two epilogue groups of 8 warps — even tiles (accumulator 0) and odd tiles (accumulator 1) — so that the read-out of
//                 one accumulator overlaps the next tile's MMA and the other group's read-out; in a group two threads serve a query
//                 (column halves), 4 × `tcgen05.ld.32x32b.x32` each, double-buffered; per column ONE multiply-add forms a max-ordered
//                 key (2·dot − |d| in the high bits, inverted row below), a max3 tree pre-reduces 32 columns, and the exact top-2
//                 insertion runs only for the groups of columns that can improve the running second best
// Three mbarrier pipelines connect them (shared-memory stage full/empty, accumulator full/empty), two stages each; the key bases
// live in a 4-deep ring of their own.
#include <cuda_runtime.h>
#include <stdint.h>

#include "orbx_internal.h"

namespace {

constexpr int TC_M = 128;             // queries per CTA
constexpr int TC_N = 256;             // database rows per MMA tile
constexpr int TC_KBYTES = 256;        // operand bytes per row (one byte per descriptor bit)
constexpr int TC_PRODUCERS = 128;     // threads (warps 1-4): two database rows each per tile
constexpr int TC_EPILOGUE = 256;      // threads of ONE epilogue group (warps 5-12: even tiles, warps 13-20: odd tiles)
constexpr int TC_THREADS = 32 + TC_PRODUCERS + 2 * TC_EPILOGUE;
constexpr int TC_A_BYTES = TC_M * TC_KBYTES;            // 32 KB: two K-blocks of [128 rows][128 B]
constexpr int TC_B_BYTES = TC_N * TC_KBYTES;            // 64 KB per stage: two K-blocks of [256 rows][128 B]
constexpr int TC_BASE_STAGES = 4;     // key-base ring: written by the producers of tile t, read by its epilogue, reused by tile t + 4
constexpr int TC_SMEM = TC_A_BYTES + 2 * TC_B_BYTES + TC_BASE_STAGES * TC_N * 4 + 4096 /* barriers, tmem address, merge buffer */ + 1024 /* alignment slack */;
constexpr uint32_t TC_ROW_BITS = 22;     // rows of a chunk inside the max-ordered keys of the epilogue (chunks hold < 2^22 - 1 rows)
constexpr uint32_t TC_IDX_BITS = 23;
constexpr uint32_t TC_KEY_NONE = 0xffffffffu;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void bar_wait(uint64_t *b, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_DONE_%=;\n"
        "bra TC_WAIT_%=;\n"
        "TC_DONE_%=:\n"
        "}\n" ::"r"(s32(b)), "r"(parity) : "memory");
}

// shared-memory matrix descriptor of a K-major operand slab [rows][128 B] with the 128-byte swizzle: start address (>> 4), leading
// byte offset unused for swizzled K-major (1), stride byte offset = 8 rows × 128 B = 1024 (>> 4), descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc(uint32_t smemAddr) {
    return (uint64_t)((smemAddr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// One 32-byte descriptor row → 256 operand bytes in the two K-block slabs of a tile (slab k holds operand bytes 128k … 128k+127).
// Operand word 8i + j = the bits j, j+8, j+16, j+24 of descriptor word i, left where they are (DB: byte ∈ {0, 2^j}) or moved to
// bit 7 − j (QUERY: byte ∈ {0, 2^(7−j)}).  rowAddr: shared-window address of the row in slab 0; the 16-byte chunks of a row are
// XOR-swizzled with the row index modulo 8.
template <bool QUERY>
__device__ __forceinline__ void expand_row(uint32_t rowAddr, uint32_t slabStride, uint32_t r, const uint4 &lo, const uint4 &hi) {
    const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    const uint32_t rx = (r & 7u) << 4;
#pragma unroll
    for (int c = 0; c < 16; ++c) {                // chunk c = operand words 4c … 4c+3 = descriptor word c >> 1, j = 4·(c & 1) … +3
        const uint32_t x = w[c >> 1];
        const int j0 = 4 * (c & 1);
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = j0 + k;
            o[k] = QUERY ? ((x >> j) & 0x01010101u) << (7 - j) : x & (0x01010101u << j);
        }
        sts128(rowAddr + (uint32_t)(c >> 3) * slabStride + (((uint32_t)(c & 7) << 4) ^ rx), o[0], o[1], o[2], o[3]);
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t (&v)[32], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ int popc256(const uint4 &lo, const uint4 &hi) {
    return __popc(lo.x) + __popc(lo.y) + __popc(lo.z) + __popc(lo.w) + __popc(hi.x) + __popc(hi.y) + __popc(hi.z) + __popc(hi.w);
}

// grid (nChunks, ceil(nq / 128)); partial[chunk * nq + q] = {best key, second key} with rows relative to the chunk start
__global__ void __launch_bounds__(TC_THREADS, 1) k_knn2_tc(const uint4 *__restrict__ q, int nq, const uint4 *__restrict__ db, long long ndb, long long chunkRows,
                                                           uint2 *__restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // the swizzle atoms need 1024-byte alignment
    uint8_t *sA = smem;                                         // [2 K-blocks][128][128]
    uint8_t *sB = smem + TC_A_BYTES;                            // [2 stages][2 K-blocks][256][128]
    uint32_t *sBase = reinterpret_cast<uint32_t *>(sB + 2 * TC_B_BYTES);      // [4][256]: key base of the column (ring over tiles)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sBase + TC_BASE_STAGES * TC_N);          // full[2], empty[2], tfull[2], tempty[2]
    uint32_t *tmemAddr = reinterpret_cast<uint32_t *>(bars + 8);
    uint2 *sMerge = reinterpret_cast<uint2 *>(tmemAddr + 2);                  // [3][128]: the top-2 of the other (group, column half) threads of a query
    uint64_t *full = bars, *empty = bars + 2, *tfull = bars + 4, *tempty = bars + 6;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.y * TC_M;
    const long long row0 = (long long)blockIdx.x * chunkRows;
    const long long rowsHere = min(chunkRows, ndb - row0);
    const int nTiles = (int)((rowsHere + TC_N - 1) / TC_N);

    // ---- set-up: barriers, tensor memory, the query tile ----
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            bar_init(&full[s], TC_PRODUCERS);
            bar_init(&empty[s], 1);                             // the MMA commit: the tensor core has read the stage
            bar_init(&tfull[s], 1);
            bar_init(&tempty[s], TC_EPILOGUE);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();                                           // converged again after the tid == 0 branch (.sync.aligned below)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmemAddr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int r = tid; r < TC_M; r += TC_THREADS) {              // queries beyond nq are zero rows (their results are not written)
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (q0 + r < nq) { lo = q[2 * (long long)(q0 + r)]; hi = q[2 * (long long)(q0 + r) + 1]; }
        expand_row<true>(s32(sA) + (uint32_t)r * 128u, TC_M * 128, (uint32_t)r, lo, hi);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes → visible to the tensor core's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmemAddr;

    if (warp == 0) {
        // ---- MMA issuer ----
        if (lane == 0) {
            // instruction descriptor (kind::i8): D = s32 (2 << 4), A and B unsigned 8-bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            const uint32_t aBase = s32(sA);
            for (int t = 0; t < nTiles; ++t) {
                const int s = t & 1;
                const uint32_t ph = (uint32_t)(t >> 1) & 1u;
                bar_wait(&tempty[s], ph ^ 1u);                  // the epilogue has drained this accumulator
                bar_wait(&full[s], ph);                         // the producers have filled this stage
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t bBase = s32(sB + s * TC_B_BYTES);
                const uint32_t d = tmem + (uint32_t)(s * TC_N);
#pragma unroll
                for (int k = 0; k < TC_KBYTES / 32; ++k) {      // K = 32 bytes per MMA; 4 steps inside each 128-byte K-block
                    const uint64_t ad = umma_desc(aBase + (uint32_t)(k >> 2) * (TC_M * 128) + (uint32_t)(k & 3) * 32);
                    const uint64_t bd = umma_desc(bBase + (uint32_t)(k >> 2) * (TC_N * 128) + (uint32_t)(k & 3) * 32);
                    const uint32_t acc = k > 0 ? 1u : 0u;
                    asm volatile(
                        "{\n"
                        ".reg .pred p;\n"
                        "setp.ne.b32 p, %4, 0;\n"
                        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
                        "}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                }
                // commits track the completion of everything issued so far: the stage's shared memory is free, the accumulator is full
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[s])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&tfull[s])) : "memory");
            }
        }
    } else if (warp <= 4) {
        // ---- producers ----  (rows r and r + 128 of every tile; the next tile's rows are fetched while the current ones are expanded)
        const int r = tid - 32;                                 // 0 … 127
        uint4 lo[2], hi[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            lo[h] = make_uint4(0, 0, 0, 0); hi[h] = lo[h];
            if (r + 128 * h < rowsHere) { lo[h] = db[2 * (row0 + r + 128 * h)]; hi[h] = db[2 * (row0 + r + 128 * h) + 1]; }
        }
        for (int t = 0; t < nTiles; ++t) {
            const int s = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            uint4 clo[2], chi[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                clo[h] = lo[h]; chi[h] = hi[h];
                const long long nr = (long long)(t + 1) * TC_N + r + 128 * h;
                lo[h] = make_uint4(0, 0, 0, 0); hi[h] = lo[h];
                if (nr < rowsHere) { lo[h] = db[2 * (row0 + nr)]; hi[h] = db[2 * (row0 + nr) + 1]; }
            }
            bar_wait(&empty[s], ph ^ 1u);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t rr = (uint32_t)(r + 128 * h);    // the tile row = accumulator column
                const long long lr = (long long)t * TC_N + rr;  // row inside the chunk
                expand_row<false>(s32(sB) + (uint32_t)(s * TC_B_BYTES) + rr * 128u, TC_N * 128, rr, clo[h], chi[h]);
                // key base of the column: (256 − |d|) above the inverted row (larger key = smaller distance, then smaller row); 0 = no row
                // (ring slot t & 3: its previous user, tile t − 4, was drained before the MMA of tile t − 2 could start, and that MMA's
                // commit is what freed this shared-memory stage)
                sBase[(t & (TC_BASE_STAGES - 1)) * TC_N + rr] =
                    lr < rowsHere ? ((uint32_t)(256 - popc256(clo[h], chi[h])) << TC_ROW_BITS) + (((1u << TC_ROW_BITS) - 1u) - (uint32_t)lr) : 0u;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_arrive(&full[s]);
        }
    } else {
        // ---- epilogue: thread = (query row, column half) ----
        const int et = tid - 32 - TC_PRODUCERS;                 // 0 … 511
        const int grp = et >> 8;                                // 0: even tiles, 1: odd tiles
        const int ew = (et >> 5) & 7;                           // warp inside the group
        const int quarter = warp & 3;                           // the TMEM lane quarter this warp may access is fixed by its id
        const int half = ew >> 2;                               // 0: columns 0-127, 1: columns 128-255 (four consecutive warps cover each quarter once)
        const int m = quarter * 32 + lane;                      // query row = TMEM lane
        int qn = 0;
        if (q0 + m < nq) {
            const uint4 lo = q[2 * (long long)(q0 + m)], hi = q[2 * (long long)(q0 + m) + 1];
            qn = popc256(lo, hi);
        }
        // running top-2 in MAX order over keys  (2·dot − |d| + 256) << 22 | (2^22 − 1 − row)   (real keys are > 0)
        uint32_t a = 0, b = 0;
        for (int t = grp; t < nTiles; t += 2) {
            const int s = grp;                                  // == t & 1
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            bar_wait(&tfull[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t base = s32(sBase) + (uint32_t)((t & (TC_BASE_STAGES - 1)) * TC_N + half * 128) * 4u;      // shared-window address of the key bases
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * TC_N + half * 128);
            // 32 columns per load, two register sets: the load of the next group is in flight while this one is reduced
            uint32_t va[32], vb[32];
            tmem_ld32(va, taddr);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t (&v)[32] = (g & 1) ? vb : va;
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (g + 1 < 4) tmem_ld32((g & 1) ? va : vb, taddr + (uint32_t)((g + 1) * 32));
                // accumulator = 128 · dot; key = dot · 2^23 + base: one multiply-add per column (a column past the chunk end has base 0
                // and dot 0: key 0 never wins)
                constexpr uint32_t kScale = 1u << (TC_ROW_BITS + 1 - 7);
                uint32_t kmax = 0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    uint4 bs;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(bs.x), "=r"(bs.y), "=r"(bs.z), "=r"(bs.w) : "r"(base + (uint32_t)(g * 32 + j) * 4u));
                    v[j] = v[j] * kScale + bs.x; v[j + 1] = v[j + 1] * kScale + bs.y;
                    v[j + 2] = v[j + 2] * kScale + bs.z; v[j + 3] = v[j + 3] * kScale + bs.w;
                    kmax = max(max(kmax, max(v[j], v[j + 1])), max(v[j + 2], v[j + 3]));
                }
                if (kmax > b) {                                 // this group can change the running top-2: exact insertion
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t lo2 = min(v[j], a);
                        a = max(v[j], a);
                        b = max(b, lo2);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(&tempty[s]);
        }
        // merge the two column halves of every query, convert to the (distance << 23 | row) keys of the merge kernel, write the partial result
        const int part = grp * 2 + half;                        // the four threads of a query: part 0 merges
        if (part > 0) sMerge[(part - 1) * TC_M + m] = make_uint2(a, b);
        asm volatile("bar.sync 1, %0;" ::"r"(2 * TC_EPILOGUE) : "memory");   // named barrier: the 512 epilogue threads only
        if (part == 0 && q0 + m < nq) {
#pragma unroll
            for (int o3 = 0; o3 < 3; ++o3) {
                const uint2 o = sMerge[o3 * TC_M + m];
                uint32_t lo2 = min(o.x, a); a = max(o.x, a); b = max(b, lo2);
                lo2 = min(o.y, a); a = max(o.y, a); b = max(b, lo2);
            }
            auto conv = [&](uint32_t k) -> uint32_t {
                if (k == 0) return TC_KEY_NONE;
                const uint32_t sc = k >> TC_ROW_BITS, row = ((1u << TC_ROW_BITS) - 1u) - (k & ((1u << TC_ROW_BITS) - 1u));
                return ((uint32_t)(qn + 256 - (int)sc) << TC_IDX_BITS) | row;          // distance = |q| + |d| − 2·dot = |q| + 256 − score
            };
            partial[(long long)blockIdx.x * nq + (q0 + m)] = make_uint2(conv(a), conv(b));
        }
    }
    // ---- teardown ----
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace

// launches the tensor-core kernel for chunks of `chunkRows` rows; partial as for the POPC kernel.  Returns cudaSuccess or the launch error.
cudaError_t orbx_knn2_tc_launch(const uint8_t *d_q, int nq, const uint8_t *d_db, long long ndb, long long chunkRows, int nChunks, uint2 *d_partial,
                                cudaStream_t stream) {
    // the opt-in to > 48 KB of dynamic shared memory is per device (a process may drive several)
    cudaError_t e = cudaFuncSetAttribute(k_knn2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)nChunks, (unsigned)((nq + TC_M - 1) / TC_M));
    k_knn2_tc<<<grid, TC_THREADS, TC_SMEM, stream>>>(reinterpret_cast<const uint4 *>(d_q), nq, reinterpret_cast<const uint4 *>(d_db), ndb, chunkRows, d_partial);
    return cudaGetLastError();
}
