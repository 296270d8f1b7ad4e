// Pipe micro-benchmark for the FAST kernel's inner loop (B200): throughput of VIMNMX3.U16x2, HMNMX2, FMNMX3, PRMT,
// POPC alone and mixed, to see which instruction classes share an issue pipe.  Prints ops/clk/SM.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed * (threadIdx.x + 1) + i * 77u; b[i] = seed ^ (i * 0x9e3779b9u + threadIdx.x); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {                      // VIMNMX3.U16x2 only (2 per slot)
                a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]);
                b[i] = __vimin3_u16x2(b[i], a[i], b[(i + 3) & 7]);
            } else if (MODE == 1) {               // HMNMX2 only (2 per slot)
                uint32_t r, s;
                asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a[i]), "r"(b[i]));
                asm volatile("min.f16x2 %0, %1, %2;" : "=r"(s) : "r"(b[i]), "r"(a[(i + 1) & 7]));
                a[i] = r; b[i] = s;
            } else if (MODE == 2) {               // 1 VIMNMX3 + 1 HMNMX2 per slot
                uint32_t s;
                a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]);
                asm volatile("min.f16x2 %0, %1, %2;" : "=r"(s) : "r"(b[i]), "r"(a[(i + 3) & 7]));
                b[i] = s;
            } else if (MODE == 3) {               // PRMT only (2 per slot)
                a[i] = __byte_perm(a[i], b[i], 0x5140);
                b[i] = __byte_perm(b[i], a[(i + 1) & 7], 0x3625);
            } else if (MODE == 4) {               // 1 VIMNMX3 + 1 PRMT per slot
                a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]);
                b[i] = __byte_perm(b[i], a[i], 0x3625);
            } else if (MODE == 5) {               // POPC only (2 per slot)
                a[i] = __popc(a[i] ^ b[i]) + a[i];
                b[i] = __popc(b[i] ^ a[(i + 1) & 7]) + b[i];
            } else if (MODE == 6) {               // FMNMX3 (2 per slot)
                float x = __uint_as_float(a[i]), y = __uint_as_float(b[i]), z = __uint_as_float(a[(i + 1) & 7]), r, s;
                asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x), "f"(y), "f"(z));
                asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(s) : "f"(y), "f"(r), "f"(z));
                a[i] = __float_as_uint(r); b[i] = __float_as_uint(s);
            } else if (MODE == 7) {               // 1 VIMNMX3 + 1 IMAD per slot
                a[i] = __vimax3_u16x2(a[i], b[i], a[(i + 1) & 7]);
                b[i] = b[i] * 3u + a[i];
            } else if (MODE == 8) {               // HFMA2 only (2 per slot) — FMA-pipe reference
                __half2 x = *reinterpret_cast<__half2 *>(&a[i]), y = *reinterpret_cast<__half2 *>(&b[i]);
                x = __hfma2(x, y, x); y = __hfma2(y, x, y);
                a[i] = *reinterpret_cast<uint32_t *>(&x); b[i] = *reinterpret_cast<uint32_t *>(&y);
            } else if (MODE == 9) {               // 1 HMNMX2 + 1 HFMA2 per slot
                uint32_t s;
                asm volatile("min.f16x2 %0, %1, %2;" : "=r"(s) : "r"(b[i]), "r"(a[(i + 3) & 7]));
                __half2 x = *reinterpret_cast<__half2 *>(&a[i]), y = *reinterpret_cast<__half2 *>(&s);
                x = __hfma2(x, y, x);
                a[i] = *reinterpret_cast<uint32_t *>(&x); b[i] = s;
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc ^= a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, uint32_t *d, int nSM, double clkGHz) {
    const int iters = 4096, blocks = nSM * 8;
    k<MODE><<<blocks, 256>>>(d, 64, 1u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 3u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256 * iters * 16;   // 16 instructions of interest per iteration per thread
    printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at %.3f GHz)\n", name, ms, ops / (ms * 1e-3) / nSM / (clkGHz * 1e9), clkGHz);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ghz = clk / 1e6;
    uint32_t *d;
    cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    run<0>("VIMNMX3.U16x2", d, p.multiProcessorCount, ghz);
    run<1>("HMNMX2", d, p.multiProcessorCount, ghz);
    run<2>("VIMNMX3 + HMNMX2", d, p.multiProcessorCount, ghz);
    run<3>("PRMT", d, p.multiProcessorCount, ghz);
    run<4>("VIMNMX3 + PRMT", d, p.multiProcessorCount, ghz);
    run<5>("POPC(+LOP,IADD)", d, p.multiProcessorCount, ghz);
    run<6>("FMNMX3", d, p.multiProcessorCount, ghz);
    run<7>("VIMNMX3 + IMAD", d, p.multiProcessorCount, ghz);
    run<8>("HFMA2", d, p.multiProcessorCount, ghz);
    run<9>("HMNMX2 + HFMA2", d, p.multiProcessorCount, ghz);
    return 0;
}
