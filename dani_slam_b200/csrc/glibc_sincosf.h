// glibc_sincosf.h — bit-exact restatement of glibc 2.39's x86-64 FMA `sinf`/`cosf` for |x| < 120.
//
// Why: the reference computes the rBRIEF rotation with `(float)cos(angle)`, `(float)sin(angle)` on a
// float argument (/root/reference/src/ORBextractor.cc:111-112, resolved to cosf/sinf through
// `using namespace std`, :66).  glibc's sinf/cosf are NOT correctly rounded (≈0.56 ulp), so neither
// CUDA's sinf/cosf nor (float)sin((double)x) reproduce them; a 1-ulp difference can flip a descriptor
// bit.  glibc evaluates a double-precision polynomial after a fast range reduction;
// the IFUNC-selected variant on FMA hosts (`__sinf_fma`/`__cosf_fma`, libm.so.6 .text 0x7e800 /
// 0x7e330 in this image) contracts specific multiply-adds.  The operation order and the fused
// operations below were read from that variant's disassembly, the constants from its
// `__sincosf_table` (.rodata 0xb8120); tests/test_device_math_gpu.py sweeps every float in [0, 2π] against the
// host libm, on the CPU (host build of this header) and on the GPU (device build).
//
// Works identically as host C++ (std::fma) and CUDA device code (__fma_rn & friends): only IEEE
// double mul/add/fma, one float→double and one double→float conversion.
#ifndef ORBX_GLIBC_SINCOSF_H
#define ORBX_GLIBC_SINCOSF_H

#include <stdint.h>
#include <string.h>
#if !defined(__CUDA_ARCH__)
#include <math.h>
#endif

#if defined(__CUDACC__)
#define ORBX_HD __host__ __device__ __forceinline__
#else
#define ORBX_HD static inline
#endif

namespace orbx_libm {

#if defined(__CUDA_ARCH__)
ORBX_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
ORBX_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
ORBX_HD uint32_t fbits(float f) { return __float_as_uint(f); }
ORBX_HD float d2f(double d) { return __double2float_rn(d); }
ORBX_HD int d2i_trunc(double d) { return __double2int_rz(d); }
#else
ORBX_HD double dmul(double a, double b) { volatile double r = a * b; return r; }
ORBX_HD double dfma(double a, double b, double c) { return fma(a, b, c); }
ORBX_HD uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
ORBX_HD float d2f(double d) { return (float)d; }
ORBX_HD int d2i_trunc(double d) { return (int)d; }
#endif

// __sincosf_table[0]; table[1] is the same with c0..c4 negated.
#define ORBX_HPI_INV 0x1.45f306dc9c883p+23  /* 2/pi * 2^24 */
#define ORBX_HPI     0x1.921fb54442d18p+0
#define ORBX_C0      0x1.0000000000000p+0
#define ORBX_C1     -0x1.ffffffd0c621cp-2
#define ORBX_C2      0x1.55553e1068f19p-5
#define ORBX_C3     -0x1.6c087e89a359dp-10
#define ORBX_C4      0x1.99343027bf8c3p-16
#define ORBX_S1     -0x1.555545995a603p-3
#define ORBX_S2      0x1.1107605230bc4p-7
#define ORBX_S3     -0x1.994eb3774cf24p-13

// sine polynomial: x already multiplied by the quadrant sign
ORBX_HD double poly_sin(double x, double x2) {
    const double s1 = dfma(x2, ORBX_S3, ORBX_S2);
    const double x3 = dmul(x2, x);
    const double x7 = dmul(x2, x3);
    const double s = dfma(x3, ORBX_S1, x);
    return dfma(s1, x7, s);
}
// cosine polynomial; neg selects table[1] (all c negated)
ORBX_HD double poly_cos(double x2, bool neg) {
    const double sg = neg ? -1.0 : 1.0;
    const double x4 = dmul(x2, x2);
    const double c1 = dfma(x2, sg * ORBX_C1, sg * ORBX_C0);
    const double c2 = dfma(x2, sg * ORBX_C4, sg * ORBX_C3);
    const double x6 = dmul(x2, x4);
    const double c = dfma(x4, sg * ORBX_C2, c1);
    return dfma(c2, x6, c);
}

ORBX_HD uint32_t abstop12(float f) { return (fbits(f) >> 20) & 0x7ff; }

// reduce_fast: n = round(x * 2/pi) through a 2^24-scaled truncation, x -= n*pi/2 (one fnmadd)
ORBX_HD double reduce_fast(double x, int *np) {
    const double r = dmul(x, ORBX_HPI_INV);
    const int n = (d2i_trunc(r) + 0x800000) >> 24;
    *np = n;
    return dfma(-(double)n, ORBX_HPI, x);
}

// valid for |y| < 120 (abstop12 <= 0x42e); the extractor only feeds [0, 2π].
ORBX_HD float sinf_glibc(float y) {
    const double x = (double)y;
    const uint32_t top = abstop12(y);
    if (top <= 0x3f3) {  // |y| < pi/4
        if (top <= 0x397) return y;  // |y| < 2^-12
        return d2f(poly_sin(x, dmul(x, x)));
    }
    int n;
    const double r = reduce_fast(x, &n);
    const double r2 = dmul(r, r);
    if (n & 1) return d2f(poly_cos(r2, (n & 2) != 0));
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return d2f(poly_sin(dmul(r, sgn), r2));
}

ORBX_HD float cosf_glibc(float y) {
    const double x = (double)y;
    const uint32_t top = abstop12(y);
    if (top <= 0x3f3) {
        if (top <= 0x397) return 1.0f;
        return d2f(poly_cos(dmul(x, x), false));
    }
    int n;
    const double r = reduce_fast(x, &n);
    const double r2 = dmul(r, r);
    if ((n & 1) == 0) return d2f(poly_cos(r2, (n & 2) != 0));
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    return d2f(poly_sin(dmul(r, sgn), r2));
}

}  // namespace orbx_libm
#endif
