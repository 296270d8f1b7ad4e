// orbx_knn_tc.cu — brute-force Hamming kNN (k = 2) on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// For bit vectors q, d ∈ {0,1}^256:  hamming(q, d) = |q| + |d| − 2·(q · d) = |q| − Σ_k (2·q_k − 1)·d_k.  The sums of a tile of 128
// queries with a tile of 256 database rows are ONE 128×256×256 integer GEMM with the queries as ±1 and the database rows as 0/1:
// every bit becomes one operand byte in shared memory (K-major, 128-byte swizzle — the canonical UMMA operand layout).  The order
// of the 256 terms is free, so operand word j (j = 0…7) of a 32-bit descriptor word x holds the bits j, j+8, j+16, j+24 IN PLACE:
// a database byte is x & (1 << j) ∈ {0, 2^j} (j = 7: one shift down, 2^6) — one AND per four operand bytes — and the query byte
// of the same term carries the inverse scale, ±2^(6−j) (j = 7: ±1), so that every product is ±64·(d bit) and the accumulator is
// exactly 64·(2·(q · d) − |d|): the database popcount never has to be computed.  `tcgen05.mma.kind::i8` (signed × unsigned)
// accumulates in int32 in tensor memory, and the epilogue reads the accumulators back with `tcgen05.ld`.  Larger accumulator =
// smaller distance: a max3 tree over the raw accumulators of 32 columns decides whether the group can touch the running top-2
// of the query at all; only then are the packed (score, inverted row) keys formed and inserted.  Tie rule (lower row first) and
// per-chunk partial format are those of the POPC kernel in orbx_match.cu, whose merge kernel finishes the job.  Results are
// bit-identical to the POPC path (tests/test_match_gpu.py).
//
// One CTA = one tile of 128 queries × one chunk of database rows; warp roles (13 warps; TC_GROUPS = 2 adds a second epilogue group):
//   warp 0        allocates tensor memory; lane 0 issues the MMAs (8 per database tile: K = 8 × 32 bytes) and commits them
//   warps 1-4     producers: two database rows per thread and tile (32 B each, coalesced, prefetched one tile ahead), expanded to the
//                 swizzled operand tile of the free stage (64 ANDs, 8 shifts and 16 16-byte stores per row)
//   warps 5-12    the epilogue group: two threads serve a query (column halves), 4 × `tcgen05.ld.32x32b.x32` each, double-buffered,
//                 while the tensor core fills the other accumulator stage; 16 max3 per 32 columns on the raw accumulators, and only
//                 an octet whose best score beats the running second best of the thread's query forms its keys ((score << 22) +
//                 inverted row, one multiply-add per column) and runs the exact top-2 insertion.  (One group of 8 warps keeps up
//                 with the tensor core; with two groups — even / odd tiles — the extra spinning warps cost more than they gave.)
// Three mbarrier pipelines connect them (shared-memory stage full/empty, accumulator full/empty), two stages each.
#include <cuda_runtime.h>
#include <stdint.h>

#include "orbx_internal.h"

namespace {

constexpr int TC_M = 128;             // queries per CTA
constexpr int TC_N = 256;             // database rows per MMA tile
constexpr int TC_KBYTES = 256;        // operand bytes per row (one byte per descriptor bit)
constexpr int TC_PRODUCERS = 128;     // threads (warps 1-4): two database rows each per tile
constexpr int TC_EPILOGUE = 256;      // threads of ONE epilogue group (warps 5-12; a second group would be warps 13-20)
constexpr int TC_GROUPS = 1;           // epilogue groups: 1 = every tile by the same 8 warps, 2 = even / odd tiles by 8 warps each
constexpr int TC_THREADS = 32 + TC_PRODUCERS + TC_GROUPS * TC_EPILOGUE;
constexpr int TC_A_BYTES = TC_M * TC_KBYTES;            // 32 KB: two K-blocks of [128 rows][128 B]
constexpr int TC_B_BYTES = TC_N * TC_KBYTES;            // 64 KB per stage: two K-blocks of [256 rows][128 B]
constexpr int TC_SMEM = TC_A_BYTES + 2 * TC_B_BYTES + 4096 /* barriers, tmem address, merge buffer */ + 1024 /* alignment slack */;
constexpr uint32_t TC_ROW_BITS = 22;     // rows of a chunk inside the max-ordered keys of the epilogue (chunks hold < 2^22 - 1 rows)
constexpr uint32_t TC_IDX_BITS = 23;
constexpr uint32_t TC_KEY_NONE = 0xffffffffu;
constexpr int TC_NONE = (int)0x80000000;       // "no candidate" in the signed max-ordered keys of the epilogue (real keys are >= -2^30)

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
// Operand word 8i + j = the bits j, j+8, j+16, j+24 of descriptor word i.  DB: the bits stay where they are, byte ∈ {0, 2^j}
// (j = 7: shifted down once, {0, 2^6}).  QUERY: byte = +s for a set bit, −s for a clear one, s = 2^(6−j) (j = 7: 1); a query
// beyond nq (zero = true) is all-zero bytes.  rowAddr: shared-window address of the row in slab 0; the 16-byte chunks of a row
// are XOR-swizzled with the row index modulo 8.
template <bool QUERY>
__device__ __forceinline__ void expand_row(uint32_t rowAddr, uint32_t slabStride, uint32_t r, const uint4 &lo, const uint4 &hi, bool zero = false) {
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
            if (QUERY) {
                const uint32_t sc = j < 7 ? (1u << (6 - j)) : 1u;
                const uint32_t M = ((x >> j) & 0x01010101u) * 255u;                              // 0xff in the bytes of set bits
                o[k] = zero ? 0u : (M & (0x01010101u * sc)) | (~M & (0x01010101u * (256u - sc)));   // +s | −s as two's-complement bytes
            } else {
                o[k] = j < 7 ? x & (0x01010101u << j) : (x >> 1) & 0x40404040u;
            }
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
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + 2 * TC_B_BYTES);         // full[2], empty[2], tfull[2], tempty[2]
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
        expand_row<true>(s32(sA) + (uint32_t)r * 128u, TC_M * 128, (uint32_t)r, lo, hi, q0 + r >= nq);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes → visible to the tensor core's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmemAddr;

    if (warp == 0) {
        // ---- MMA issuer ----
        if (lane == 0) {
            // instruction descriptor (kind::i8): D = s32 (2 << 4), A signed 8-bit (1 << 7), B unsigned 8-bit, both K-major, N >> 3 at bit 17,
            // M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
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
                expand_row<false>(s32(sB) + (uint32_t)(s * TC_B_BYTES) + rr * 128u, TC_N * 128, rr, clo[h], chi[h]);
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
        // running top-2 in signed MAX order over keys  (score << 22) + (2^22 − 1 − row),  score = 2·dot − |d| = accumulator / 64
        int a = TC_NONE, b = TC_NONE;
        for (int t = grp; t < nTiles; t += TC_GROUPS) {
            const int s = t & 1;
            const uint32_t ph = (uint32_t)(t >> 1) & 1u;
            bar_wait(&tfull[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(s * TC_N + half * 128);
            const long long col0 = (long long)t * TC_N + half * 128;          // chunk row of this thread's first column
            // 32 columns per load, two register sets: the load of the next group is in flight while this one is reduced
            uint32_t va[32], vb[32];
            tmem_ld32(va, taddr);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t (&v)[32] = (g & 1) ? vb : va;
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (g + 1 < 4) tmem_ld32((g & 1) ? va : vb, taddr + (uint32_t)((g + 1) * 32));
                // maxima of the four octets of columns and of the group
                int m8[4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    int mx = max(max((int)v[8 * o], (int)v[8 * o + 1]), (int)v[8 * o + 2]);
                    mx = max(max(mx, (int)v[8 * o + 3]), (int)v[8 * o + 4]);
                    mx = max(max(mx, (int)v[8 * o + 5]), (int)v[8 * o + 6]);
                    m8[o] = max(mx, (int)v[8 * o + 7]);
                }
                const int vmax = max(max(m8[0], m8[1]), max(m8[2], m8[3]));
                // Can a column enter the top-2?  Only with a score above the running second best's: this thread meets its rows in
                // increasing order, so a later row never wins a tie.  thr = the smallest accumulator (64 · score) that qualifies
                // (b = TC_NONE: everything does); b only grows, so a stale thr is merely conservative.
                const int thr = ((b >> TC_ROW_BITS) + 1) * 64;
                if (vmax >= thr) {
                    const long long lr0 = col0 + g * 32;
                    const int nValid = (int)min(32LL, rowsHere - lr0);          // columns past the end of the chunk hold no row
                    const int inv0 = (int)(((1u << TC_ROW_BITS) - 1u) - (uint32_t)lr0);
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        if (m8[o] >= thr) {
#pragma unroll
                            for (int j = 8 * o; j < 8 * o + 8; ++j) {
                                const int key = j < nValid ? (int)v[j] * (1 << (TC_ROW_BITS - 6)) + (inv0 - j) : TC_NONE;
                                const int lo2 = min(key, a);
                                a = max(key, a);
                                b = max(b, lo2);
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(&tempty[s]);
        }
        // merge the four (tile parity, column half) results of every query, convert to the (distance << 23 | row) keys of the merge kernel,
        // write the partial result
        const int part = grp * 2 + half;                        // the threads of a query: part 0 merges
        if (part > 0) sMerge[(part - 1) * TC_M + m] = make_uint2((uint32_t)a, (uint32_t)b);
        asm volatile("bar.sync 1, %0;" ::"r"(TC_GROUPS * TC_EPILOGUE) : "memory");   // named barrier: the epilogue threads only
        if (part == 0 && q0 + m < nq) {
#pragma unroll
            for (int o3 = 0; o3 < 2 * TC_GROUPS - 1; ++o3) {
                const uint2 o = sMerge[o3 * TC_M + m];
                int lo2 = min((int)o.x, a); a = max((int)o.x, a); b = max(b, lo2);
                lo2 = min((int)o.y, a); a = max((int)o.y, a); b = max(b, lo2);
            }
            auto conv = [&](int k) -> uint32_t {
                if (k == TC_NONE) return TC_KEY_NONE;
                const int sc = k >> TC_ROW_BITS;                                        // arithmetic shift: the score, exactly
                const uint32_t row = ((1u << TC_ROW_BITS) - 1u) - ((uint32_t)k & ((1u << TC_ROW_BITS) - 1u));
                return ((uint32_t)(qn - sc) << TC_IDX_BITS) | row;                     // distance = |q| + |d| − 2·dot = |q| − score
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
