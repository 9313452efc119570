// mrc_fft.cuh -- shared-memory FFT used by the analysis kernel (MDCT, Hann spectra) and the decode kernel (IMDCT).
#pragma once
#include "mrc_internal.cuh"

// Position of input element n of a size-2^logn FFT for fft_r4 below: the digits of n reversed, in radix 4 with one
// innermost radix-2 digit when logn is odd (n = t0 + 4 t1 + ... + 4^(m-1) t_(m-1) + 4^m u  ->  u + 2 (t_(m-1) + 4 t_(m-2)
// + ... + 4^(m-1) t0)).
__device__ __forceinline__ int fft_r4_pos(int n, int logn) {
    const int m2 = logn & ~1;                                  // bits of the radix-4 digits
    const unsigned t = (unsigned)n & ((1u << m2) - 1u);
    unsigned b = m2 ? (__brev(t) >> (32 - m2)) : 0u;           // bits reversed ...
    b = ((b & 0xaaaaaaaau) >> 1) | ((b & 0x55555555u) << 1);   // ... and swapped back inside every digit
    return (logn & 1) ? (int)(((unsigned)n >> m2) | (b << 1)) : (int)b;
}

// In-place decimation-in-time FFT of size 2^logn on input placed by fft_r4_pos: one radix-2 stage first when logn is
// odd, then radix-4 stages (half as many barriers and 0.4 x the instructions of radix 2).  `nthr` threads of this group
// (local id lt); a radix-4 stage has n/4 butterflies.  tw[k] = exp(-2*pi*j*k/Ltab), k < Ltab/2, Ltab = 2^logLtab >= n.
// NB > 1: NB independent transforms of that size on the consecutive sub-arrays a + s*n, all stages in lock step.
//
// WL (warp-local stages): butterflies bi = 32m .. 32m+31 of a radix-4 stage with 4h <= 128 read and write exactly the
// elements [128m, 128m+128), in every such stage, and one warp owns them in all stages (lt must be warp-aligned and
// n >= 128) -- so those stages are separated by __syncwarp() only; a block barrier is needed before the first stage
// whose butterflies reach across 128-element spans, and after the last stage.  The arithmetic is the same either way.
template <typename T, int NB = 1, bool WL = false>
__device__ __forceinline__ void fft_r4(cpx<T>* a, int logn, int lt, int nthr, const cpx<T>* __restrict__ tw,
                                       int logLtab) {
    const int n = 1 << logn;
    int h = 1;
    if (logn & 1) {
        if constexpr (WL) {
            // two radix-2 butterflies per thread, laid out so that butterfly group m covers elements [128m, 128m+128)
            for (int bi = lt; bi < NB * (n >> 2); bi += nthr) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int q = ((bi >> 5) << 6) + (bi & 31) + 32 * t;
                    const cpx<T> u = a[2 * q], v = a[2 * q + 1];
                    a[2 * q].x = u.x + v.x;      a[2 * q].y = u.y + v.y;
                    a[2 * q + 1].x = u.x - v.x;  a[2 * q + 1].y = u.y - v.y;
                }
            }
            __syncwarp();
        } else {
            for (int q = lt; q < NB * (n >> 1); q += nthr) {
                const cpx<T> u = a[2 * q], v = a[2 * q + 1];
                a[2 * q].x = u.x + v.x;      a[2 * q].y = u.y + v.y;
                a[2 * q + 1].x = u.x - v.x;  a[2 * q + 1].y = u.y - v.y;
            }
            __syncthreads();
        }
        h = 2;
    }
    const int half_tab = 1 << (logLtab - 1);
    for (; 4 * h <= n; h <<= 2) {
        const int logh = 31 - __clz(h);
        for (int bi = lt; bi < NB * (n >> 2); bi += nthr) {
            const int j = bi & (h - 1);
            const int base = ((bi >> logh) << (logh + 2)) + j;     // sub-array s = bi / (n/4) lands at s*n by itself
            const int k1 = j << (logLtab - logh - 2);          // W_(4h)^j = tw[j * Ltab / (4h)]
            const cpx<T> w1 = tw[k1], w2 = tw[2 * k1];
            const int k3 = 3 * k1;
            cpx<T> w3;
            if (k3 >= half_tab) { w3 = tw[k3 - half_tab]; w3.x = -w3.x; w3.y = -w3.y; }
            else w3 = tw[k3];
            const cpx<T> x0 = a[base], x1 = a[base + h], x2 = a[base + 2 * h], x3 = a[base + 3 * h];
            const T t1x = x1.x * w1.x - x1.y * w1.y, t1y = x1.x * w1.y + x1.y * w1.x;
            const T t2x = x2.x * w2.x - x2.y * w2.y, t2y = x2.x * w2.y + x2.y * w2.x;
            const T t3x = x3.x * w3.x - x3.y * w3.y, t3y = x3.x * w3.y + x3.y * w3.x;
            const T s02x = x0.x + t2x, s02y = x0.y + t2y, d02x = x0.x - t2x, d02y = x0.y - t2y;
            const T s13x = t1x + t3x, s13y = t1y + t3y, d13x = t1x - t3x, d13y = t1y - t3y;
            a[base].x = s02x + s13x;          a[base].y = s02y + s13y;
            a[base + h].x = d02x + d13y;      a[base + h].y = d02y - d13x;          // d02 - j d13
            a[base + 2 * h].x = s02x - s13x;  a[base + 2 * h].y = s02y - s13y;
            a[base + 3 * h].x = d02x - d13y;  a[base + 3 * h].y = d02y + d13x;      // d02 + j d13
        }
        // the next stage (4h) stays inside the warp's 128-element span iff 16h <= 128
        if (WL && 16 * h <= 128 && 16 * h <= n) __syncwarp();
        else __syncthreads();
    }
}

// ---- conflict-free variant for the analysis kernel (power-of-two sizes) ----------------------------------------
// The L1/shared data pipe is what bounds the analysis kernel (profiles/r02p_ncu_full_analysis.csv: 64 % busy, half of
// its wavefronts bank-conflict replays, four fifths of those in the transforms), so the work buffer is SWIZZLED and
// the twiddles come from per-stage tables:
//   * element e lives at fft_swz(e): the low index bits XORed with two higher ones, chosen so that the 8 lanes of a
//     quarter warp (16-byte elements; 16 lanes and 8-byte elements in fp32) hit 8 different bank slots in EVERY access
//     pattern of the transform -- radix-2 stage (lanes vary e1..e3), radix-4 stages with h = 1 (e2..e4), h = 2
//     (e0, e3, e4), h = 4 (e0, e1, e4), h >= 8 (e0..e2) -- and in the consumers that read consecutive elements;
//   * stage h reads W1[j] = W_4h^j and W2[j] = W_4h^2j from two contiguous runs of h entries (consecutive lanes,
//     consecutive entries) and forms W3 = W1 * W2 in registers: the strided reads of one shared root table cost up to 8
//     wavefronts per load, and the FP64 pipe has room for the four extra operations.
template <typename T> __device__ __forceinline__ int fft_swz(int e);
template <> __device__ __forceinline__ int fft_swz<double>(int e) { return e ^ (((e >> 3) & 1) * 3) ^ (((e >> 4) & 1) * 6); }
template <> __device__ __forceinline__ int fft_swz<float>(int e) { return e ^ (((e >> 4) & 1) * 5) ^ (((e >> 5) & 1) * 10); }

// entries (complex) of the stage tables of a 2^logn-point transform: 2h per stage h = h0, 4 h0, ... (4h <= n), where
// h0 = 4 for even logn (the h = 1 stage has unit twiddles) and 2 for odd logn (after the radix-2 stage)
__host__ __device__ constexpr int fft_stage_entries(int logn) {
    int tot = 0;
    for (int h = (logn & 1) ? 2 : 4; 4 * h <= (1 << logn); h <<= 2) tot += 2 * h;
    return tot;
}

// Which input goes where, lane by lane.  Thread index idx (0 .. n-1 over the threads and their trips) handles input
// element place_index(idx); it is stored at fft_swz(fft_r4_pos(.)).  For n >= 512 the map is chosen so that the 8 lanes
// of every quarter warp read 8 consecutive inputs and write 8 different bank slots: the three input
// bits that the digit reversal sends to the lowest three output bits are the lane's low bits XORed with three bits of
// the warp index (a bijection of 0 .. n-1).
template <int LOGN>
__device__ __forceinline__ int fft_place_index(int idx) {
    if constexpr (LOGN < 9) return idx;
    else {
        const int lane = idx & 31, hi = idx >> 5, y = (lane ^ hi) & 7, rest = hi >> 3;
        if constexpr (LOGN == 9) return lane | ((rest & 1) << 5) | (y << 6);                    // output bits 0..2 <- input bits 8, 6, 7
        else if constexpr (LOGN == 10)                                                            //              <- input bits 8, 9, 6
            return lane | ((rest & 1) << 5) | ((y & 1) << 6) | (((rest >> 1) & 1) << 7) | ((y >> 1) << 8);
        else return lane | ((rest & 7) << 5) | (y << 8) | ((rest >> 3) << 11);                    // LOGN >= 11: <- input bits 10, 8, 9
    }
}

// In-place DIT transform of size 2^LOGN on swizzled data placed by fft_r4_pos; `st` = the stage tables (shared
// memory).  Warp-local early stages as in fft_r4<.., WL = true> (lt warp-aligned, n >= 128: the swizzle stays inside
// aligned runs of 16 elements).  Ends with a block barrier.
template <typename T, int LOGN>
__device__ __forceinline__ void fft_sw(cpx<T>* a, int lt, int nthr, const cpx<T>* __restrict__ st) {
    constexpr int n = 1 << LOGN;
    constexpr bool WL = n >= 128;
    int h = 1;
    if constexpr (LOGN & 1) {
        for (int bi = lt; bi < (n >> 2); bi += nthr) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int q = WL ? ((bi >> 5) << 6) + (bi & 31) + 32 * t : bi + (n >> 2) * t;
                const int p0 = fft_swz<T>(2 * q), p1 = p0 ^ 1;       // the swizzle is linear over GF(2)
                const cpx<T> u = a[p0], v = a[p1];
                a[p0].x = u.x + v.x;  a[p0].y = u.y + v.y;
                a[p1].x = u.x - v.x;  a[p1].y = u.y - v.y;
            }
        }
        if (WL) __syncwarp(); else __syncthreads();
        h = 2;
    } else {
        // h = 1: unit twiddles
        for (int bi = lt; bi < (n >> 2); bi += nthr) {
            const int b0 = 4 * bi;
            const int p0 = fft_swz<T>(b0), p1 = p0 ^ 1, p2 = p0 ^ 2, p3 = p0 ^ 3;
            const cpx<T> x0 = a[p0], x1 = a[p1], x2 = a[p2], x3 = a[p3];
            const T s02x = x0.x + x2.x, s02y = x0.y + x2.y, d02x = x0.x - x2.x, d02y = x0.y - x2.y;
            const T s13x = x1.x + x3.x, s13y = x1.y + x3.y, d13x = x1.x - x3.x, d13y = x1.y - x3.y;
            a[p0].x = s02x + s13x;  a[p0].y = s02y + s13y;
            a[p1].x = d02x + d13y;  a[p1].y = d02y - d13x;
            a[p2].x = s02x - s13x;  a[p2].y = s02y - s13y;
            a[p3].x = d02x - d13y;  a[p3].y = d02y + d13x;
        }
        if (WL && 16 <= n) __syncwarp(); else __syncthreads();
        h = 4;
    }
    for (; 4 * h <= n; h <<= 2) {
        const int logh = 31 - __clz(h);
        const int k1 = fft_swz<T>(h), k2 = fft_swz<T>(2 * h), k3 = fft_swz<T>(3 * h);
        for (int bi = lt; bi < (n >> 2); bi += nthr) {
            const int j = bi & (h - 1);
            const int base = ((bi >> logh) << (logh + 2)) + j;
            const cpx<T> w1 = st[j], w2 = st[h + j];
            // leg i sits at base + i h = base ^ (i h) (disjoint bits), and the swizzle is linear over GF(2)
            const int p0 = fft_swz<T>(base), p1 = p0 ^ k1, p2 = p0 ^ k2, p3 = p0 ^ k3;
            const cpx<T> x0 = a[p0], x1 = a[p1], x2 = a[p2], x3 = a[p3];
            const T w3x = w1.x * w2.x - w1.y * w2.y, w3y = w1.x * w2.y + w1.y * w2.x;
            const T t1x = x1.x * w1.x - x1.y * w1.y, t1y = x1.x * w1.y + x1.y * w1.x;
            const T t2x = x2.x * w2.x - x2.y * w2.y, t2y = x2.x * w2.y + x2.y * w2.x;
            const T t3x = x3.x * w3x - x3.y * w3y, t3y = x3.x * w3y + x3.y * w3x;
            const T s02x = x0.x + t2x, s02y = x0.y + t2y, d02x = x0.x - t2x, d02y = x0.y - t2y;
            const T s13x = t1x + t3x, s13y = t1y + t3y, d13x = t1x - t3x, d13y = t1y - t3y;
            a[p0].x = s02x + s13x;  a[p0].y = s02y + s13y;
            a[p1].x = d02x + d13y;  a[p1].y = d02y - d13x;          // d02 - j d13
            a[p2].x = s02x - s13x;  a[p2].y = s02y - s13y;
            a[p3].x = d02x - d13y;  a[p3].y = d02y + d13x;          // d02 + j d13
        }
        st += 2 * h;
        if (WL && 16 * h <= 128 && 16 * h <= n) __syncwarp();
        else __syncthreads();
    }
}

// ---- transform sizes 2^p and 9 * 2^p (block switching: a + b = 1152 gives 288- and 576-point transforms) -------
// n = 9 P, P = 2^p: input index i = 9 i1 + i2 goes to sub-array i2 (a P-point transform over i1), then one radix-9
// pass:  X[k1 + P k2] = sum_{i2} ( W_n^(i2 k1) * Sub_i2[k1] ) * W_9^(i2 k2)   -- in place: column k1 in, column k1 out.
template <int NFFT>
struct FftShape {
    static constexpr bool pow2 = (NFFT & (NFFT - 1)) == 0;
    static constexpr int P = pow2 ? NFFT : NFFT / 9;
    static constexpr int logP = (P == 1) ? 0 : (P == 2) ? 1 : (P == 4) ? 2 : (P == 8) ? 3 : (P == 16) ? 4 : (P == 32) ? 5 :
                                (P == 64) ? 6 : (P == 128) ? 7 : (P == 256) ? 8 : (P == 512) ? 9 : (P == 1024) ? 10 :
                                (P == 2048) ? 11 : -1;
    static_assert(logP >= 0 && (pow2 || P * 9 == NFFT), "FFT size must be 2^p or 9 * 2^p");
};

template <int NFFT>
__device__ __forceinline__ int fft_pos(int i) {
    using S = FftShape<NFFT>;
    if constexpr (S::pow2) return fft_r4_pos(i, S::logP);
    else {
        const int i1 = i / 9, i2 = i - 9 * i1;
        return i2 * S::P + fft_r4_pos(i1, S::logP);
    }
}

// tw9[m] = exp(-2*pi*j*m/n9), n9 a multiple of NFFT; w9[m] = exp(-2*pi*j*m/9).  Ends with a barrier.
template <typename T, int NFFT>
__device__ __forceinline__ void fft_any(cpx<T>* a, int lt, int nthr, const cpx<T>* __restrict__ tw, int logLtab,
                                        const cpx<T>* __restrict__ tw9, int n9, const cpx<T>* w9) {
    using S = FftShape<NFFT>;
    if constexpr (S::pow2) {
        // warp-local early stages: callers give whole warps a warp-aligned lt when the group size is a multiple of 32
        if constexpr (NFFT >= 128) {
            if ((nthr & 31) == 0) fft_r4<T, 1, true>(a, S::logP, lt, nthr, tw, logLtab);
            else fft_r4<T, 1, false>(a, S::logP, lt, nthr, tw, logLtab);
        } else {
            fft_r4<T, 1, false>(a, S::logP, lt, nthr, tw, logLtab);
        }
    } else {
        constexpr int P = S::P;
        fft_r4<T, 9>(a, S::logP, lt, nthr, tw, logLtab);
        // thread (k1, g) computes the outputs k2 = g, g+3, g+6 of column k1; all reads precede all writes
        const int stride = n9 / NFFT;
        const bool act = lt < 3 * P;
        const int k1 = lt / 3, g = lt - 3 * k1;
        cpx<T> o[3];
        if (act) {
            cpx<T> v[9];
#pragma unroll
            for (int i2 = 0; i2 < 9; ++i2) {
                const cpx<T> x = a[i2 * P + k1], w = tw9[i2 * k1 * stride];
                v[i2].x = x.x * w.x - x.y * w.y;
                v[i2].y = x.x * w.y + x.y * w.x;
            }
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const int k2 = g + 3 * q;
                T sx = v[0].x, sy = v[0].y;
#pragma unroll
                for (int i2 = 1; i2 < 9; ++i2) {
                    const cpx<T> w = w9[(i2 * k2) % 9];
                    sx += v[i2].x * w.x - v[i2].y * w.y;
                    sy += v[i2].x * w.y + v[i2].y * w.x;
                }
                o[q].x = sx; o[q].y = sy;
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int q = 0; q < 3; ++q) a[k1 + P * (g + 3 * q)] = o[q];
        }
        __syncthreads();
    }
}
