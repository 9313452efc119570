// mrc_math.cuh -- scalar device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

__device__ __forceinline__ double m_log10(double x) { return log10(x); }
__device__ __forceinline__ float m_log10(float x) { return __log2f(x) * 0.30102999566398120f; }      // MUFU.LG2
__device__ __forceinline__ double m_exp10(double x) { return exp10(x); }
__device__ __forceinline__ float m_exp10(float x) { return exp2f(x * 3.3219280948873623f); }
__device__ __forceinline__ double m_atan(double x) { return atan(x); }
__device__ __forceinline__ float m_atan(float x) { return atanf(x); }

// pcmfile.py:87-101 with quantize.py:90-111 at 16 bits: x = sign * (2*|c|) / 65535 ; |c| = 32768 -> 0.0 (Q1)
// The division is done as a multiplication by RN(1/65535) with one correction step (Markstein): y0 = RN(q * rcp),
// r = q - 65535 * y0 exactly (FMA), y = RN(y0 + r * rcp) -- equal to the correctly rounded quotient for every one of the
// 32768 magnitudes (checked exhaustively in tests/test_host_logic.py::test_pcm_division_by_reciprocal_is_exact).
template <typename T>
__device__ __forceinline__ T pcm_to_fraction(int c) {
    // every step below is odd in q (round-to-nearest is symmetric), so the sign needs no separate handling
    const int q2 = (c == -32768) ? 0 : 2 * c;
    if (sizeof(T) == 4) return T((float)q2 * (1.0f / 65535.0f));        // fast mode: one float multiply (relative error 6e-8)
    const double q = (double)q2, rcp = 1.0 / 65535.0;
    const double y0 = q * rcp;
    return T(fma(fma(-65535.0, y0, q), rcp, y0));
}

// quantize.py:12-38 magnitude code of |x| with nBits (sign handled by the caller).
// code = trunc(((2^nBits - 1)*|x| + 1) / 2): multiply, add, divide -- three roundings, no FMA (Appendix A).
// The division by 2.0 is a multiplication by 0.5 here: exact in binary floating point (the operand is >= 1, so
// no underflow), hence the same three roundings.  nBits <= 31, so the code fits an int.
__device__ __forceinline__ int quant_mag_code(double ax, int nBits) {
    if (ax >= 1.0) return (int)((1ll << (nBits - 1)) - 1);
    const double full = (double)((1ll << nBits) - 1);
    return __double2int_rz(__dmul_rn(__dadd_rn(__dmul_rn(full, ax), 1.0), 0.5));
}

// quantize.py:114-146: leading zeros of the magnitude code, capped at 2^nScaleBits - 1.
// int(math.log(code, 2)) == 63 - clz(code) for every reachable code (tests/test_oracle_kat.py).
__device__ __forceinline__ int scale_factor_of(double ax, int nScaleBits, int nMantBits) {
    const int cap = (1 << nScaleBits) - 1;
    const int nBits = cap + nMantBits;
    const int code = quant_mag_code(ax, nBits);
    const int top = code > 0 ? 31 - __clz(code) : 0;
    const int lz = (nBits - 2) - top;
    return lz < cap ? lz : cap;
}

// quantize.py:294-322: block-floating-point mantissa of one line.
__device__ __forceinline__ int mantissa_of(double x, int scale, int nScaleBits, int nMantBits) {
    const int cap = (1 << nScaleBits) - 1;
    const int nBits = cap + nMantBits;
    int code = quant_mag_code(fabs(x), nBits);      // x == 0 gives trunc(0.5) = 0 by itself
    if (scale != cap) code >>= (cap - scale);
    return code + ((x < 0.0) ? (1 << (nMantBits - 1)) : 0);
}

// quantize.py:325-357 + :90-111: inverse of mantissa_of.
__device__ __forceinline__ double dequantize_of(int mant, int scale, int nScaleBits, int nMantBits) {
    const int cap = (1 << nScaleBits) - 1;
    const int nBits = cap + nMantBits;
    const int signbit = 1 << (nMantBits - 1);
    const bool neg = mant >= signbit;
    long long mag = neg ? mant - signbit : mant;
    if (scale != cap) {
        const int sh = cap - scale;
        const long long m0 = mag;
        mag <<= sh;
        if (sh > 0 && m0 > 0) mag += 1ll << (sh - 1);
    }
    const double s = neg ? -1.0 : 1.0;
    return __ddiv_rn(__dmul_rn(__dmul_rn(s, (double)mag), 2.0), (double)((1ll << nBits) - 1));
}

// psychoac.py:68-78 for one (masker, line) pair: intensity of a tonal masker at Bark distance dz.
//   s15 = SPL - 15, g = 0.37*max(SPL-40, 0)
__device__ __forceinline__ double masker_intensity(double dz, double s15, double g) {
    const double adz = fabs(dz);
    double e = s15;
    if (adz > 0.5) {
        const double t = __dadd_rn(adz, -0.5);
        e = __dadd_rn(e, __dmul_rn(-27.0, t));
        if (dz > 0.5) e = __dadd_rn(e, __dmul_rn(g, t));
    }
    return exp10(__ddiv_rn(__dadd_rn(e, -96.0), 10.0));
}

// ---- psychoacoustic spreading helpers (double precision in both modes) -------------------------------------
// q / 10 correctly rounded without the division routine: y0 = RN(q * RN(1/10)); r = q - 10*y0 is exact in an FMA;
// y = RN(y0 + r * RN(1/10)) (Markstein's correction step).
__device__ __forceinline__ double div10(double q) {
    const double y0 = q * 0.1;
    const double r = fma(-10.0, y0, q);
    return fma(r, 0.1, y0);
}

// 10^y for y in about [-300, 300], < 1.5 ulp: y = n*log10(2)/64 + r with |r| <= log10(2)/128 (two-FMA reduction
// against a hi/lo split of log10(2)/64), 10^r by a degree-6 Taylor polynomial in r (error < 1e-19 relative),
// 2^(n mod 64 / 64) from a 64-entry table `tab` (shared memory), 2^(n div 64) added to the exponent field.
__device__ __forceinline__ double exp10_tab(double y, const double* __restrict__ tab) {
    const double magic = 6755399441055744.0;                       // 1.5 * 2^52: rint through the add
    const double tn = fma(y, 212.60339807279118, magic);
    const int n = __double2loint(tn);
    const double nd = tn - magic;
    double r = fma(nd, -0.004703593682222618, y);                  // hi part: 38 significant bits, nd*hi exact
    r = fma(nd, -2.7088530630863833e-14, r);
    double p = 0.2069958486968681;
    p = fma(p, r, 0.5393829291955814);
    p = fma(p, r, 1.171255148912267);
    p = fma(p, r, 2.034678592293476);
    p = fma(p, r, 2.650949055239199);
    p = fma(p, r, 2.302585092994046);
    p = p * r;
    const double t = tab[n & 63];
    const double v = fma(t, p, t);
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}

// ---- the same helpers in the precision of the mode (fp64: the code-exact arithmetic above; fp32: the fast mode's
// float arithmetic with the hardware's approximate exponential / logarithm -- MUFU.EX2 / MUFU.LG2) ---------------
__device__ __forceinline__ double rn_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double rn_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double rn_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float rn_add(float a, float b) { return a + b; }
__device__ __forceinline__ float rn_mul(float a, float b) { return a * b; }
__device__ __forceinline__ float rn_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double sp_div10(double q) { return div10(q); }
__device__ __forceinline__ float sp_div10(float q) { return q * 0.1f; }
__device__ __forceinline__ double sp_exp10(double y, const double* __restrict__ tab) { return exp10_tab(y, tab); }
__device__ __forceinline__ float sp_exp10(float y, const double* __restrict__) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * 3.3219280948873623f));
    return r;
}
