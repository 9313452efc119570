// mrc_analysis.cu -- K1/K2: framing + PCM conversion + M/S + KBD window + MDCT (N/4-point complex FFT) +
// Hann-windowed FFT + tonal masker detection + Bark-domain spreading + SMR + grant order, fused in one CTA per
// 2L-sample block, everything staged in shared memory (two CTAs per SM: <= 64 registers, ~114 KB each at L = 1024).
//
// How the psychoacoustic part differs from a literal transcription (same results):
//   * masker spreading is the reference's sum re-associated (geometric tails by affine warp scans, plateau sums,
//     one 10**x only for loud maskers below a line); MRC_FLAG_SPREAD_SEQUENTIAL keeps the pair-by-pair loop;
//   * the band SMR is a maximum over lines: lines are ranked by a cheap upper bound and only those that can be the
//     maximum get their complete threshold (warp-cooperatively);
//   * a spectrum is only evaluated in the bands whose SMR OverallSMRs would keep (L/R vs M/S per band);
//   * the grant order of the water-filling is a merge of the per-band runs, not a loop over arg-maxima.
//
// Reference path restated here (file:line in laser55/mrcAudioCodec):
//   pcmfile.py:87-101 + quantize.py:90-111   int16 -> signed fraction (Q1: -32768 -> 0.0)
//   pacfileThem.py:799-802                    block = concat(prior, current)
//   codecThem.py:363-364                      Mid/Side in the time domain
//   window.py:104-121, mdct.py:53-96          KBD window, MDCT with n0=(b+1)/2 and the 2/N factor
//   ms_stereo.py:5-27                         ms_switch per band on the unscaled L/R lines
//   quantize.py:114-146 (codecThem.py:440-454) overall scale factor, lines *= 2^scale
//   psychoac.py:134-173, :31-78               masked threshold (tonal maskers only, Q2 integer frequency step)
//   psychoac.py:176-219, ms_stereo.py:70-81   SMR per band and M/S vs L/R selection
//   bitalloc.py:106-155                       (order of grants only: the allocation itself needs the reservoir)
//
// One CTA = one block, NT = L/2 threads; thread t owns MDCT lines t and t+L/2.  L = (a+b)/2 is a template parameter:
// nMDCTLines for long blocks, 576 for the transition blocks of block switching (a+b = 1024+128), 128 for short ones.
#include <algorithm>
#include <cstdlib>

#include "mrc_internal.cuh"
#include "mrc_math.cuh"
#include "mrc_fft.cuh"
#include "mrc_tma.cuh"

// -DMRC_PHASE_CLOCKS: development build that accumulates, per CTA, the cycles between phase boundaries (read by
// scripts/phase_clocks.py through mrc_debug_phase_clocks); never defined in the shipped library.
#ifdef MRC_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[32];
__device__ unsigned long long g_band_clk[3][32];      // pass 2 per band: cycles, tasks, complete thresholds
#define MRC_CLK(i)                                                                                                  \
    do {                                                                                                            \
        if (threadIdx.x == 0) {                                                                                     \
            const long long t_ = clock64();                                                                         \
            atomicAdd(&g_phase_clk[i], (unsigned long long)(t_ - s_clk_last));                                      \
            s_clk_last = t_;                                                                                        \
        }                                                                                                           \
    } while (0)
// warp-level: cycles of lane 0 between consecutive marks inside spread_line_warp / complete()
#define MRC_WCLK_BEGIN() long long wt_ = clock64()
#define MRC_WSYNC() __syncwarp()        /* so that lane 0's clock sees the slowest lane of the step before */
#define MRC_WCLK(i)                                                                                                 \
    do {                                                                                                            \
        const long long t_ = clock64();                                                                             \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase_clk[i], (unsigned long long)(t_ - wt_));                     \
        wt_ = t_;                                                                                                   \
    } while (0)
#else
#define MRC_CLK(i)
#define MRC_WCLK_BEGIN()
#define MRC_WSYNC()
#define MRC_WCLK(i)
#endif

namespace {

constexpr int MRC_ZLUT = 832;         // cells of 1/32 Bark: Bark(24 kHz) = 24.6
#ifndef MRC_NEAR_LOUD_N
#define MRC_NEAR_LOUD_N 2
#endif
constexpr int MRC_NEAR_LOUD = MRC_NEAR_LOUD_N;     // loud maskers included in the pass-1 bound of a line's threshold
// (any number is exact: it only decides how tight the bound is.  30 min of the bench stream, audio-s/s and 10**x pairs per
// hour: 6 -> 55.8 k / 1.78 G, 4 -> 56.4 k / 1.26 G, 3 -> 56.6 k / 1.00 G, 2 -> 56.8 k / 0.73 G; scripts/lib_variant_bench.py)

// What pass 1 leaves per line for pass 2: an UPPER bound of the line's rho (rounded up to float: it only selects
// candidates) and the line's position among the maskers (m_lo | m_hi << 16), so that the warp that completes a line's
// threshold does not search for it again.
struct LineInfo {
    float ub;
    uint32_t rng;
};

template <typename T>
struct Smem {
    T* sx;          // [2][2L]  time samples L, R -- only when the block arrives as doubles (per-block seam, XIN)
    uint32_t* px;   // [2L]     the block's PCM frames as they lie in the clip: left | right << 16 (otherwise)
    cpx<T>* buf;    // [L]      FFT work buffer
    T* lines;       // [4][L]   MDCT lines L,R,M,S
    T* xi;          // [L]      FFT intensity, later SMR-per-line scratch
    struct LineInfo* li;   // [L] pass 1 -> pass 2: bound and masker range of every line (fp64: aliases xi)
    uint32_t* rng;  // [L]      fp32: the ranges on their own (the bounds are floats already: they stay in xi)
    T* xi4;         // [4][L]   fast mode (fp32, power-of-two L): the intensities of all four spectra, from the spectra of L
                    //          and R by linearity; lies over the block's samples, which are no longer needed by then
    cpx<T>* buf2;   // [L]      fast mode: second FFT work buffer (R next to L), over xi .. ms15
    // masker tables of the current spectrum, in the mode's own precision (Q = L/2 >= number of maskers)
    T* mz;          // [Q]      Bark position
    T* ms15;        // [Q]      SPL - 15
    T* mg;          // [Q]      0.37*max(SPL-40,0)
    T* mc;          // [Q]      10^((SPL-15-96)/10): intensity inside +-0.5 Bark
    T* mU;          // [Q]      quiet maskers at or below i, decayed to z_i (upper slope, -27 dB/Bark)
    T* mS;          // [Q]      maskers at or above i, decayed to z_i (lower slope, -27 dB/Bark); [npk] = 0
    T* mP;          // [Q]      sum of mc below i ([npk] = total); only for the pass-1 bound (npk < Q always)
    T* zb;          // [L]      Bark position of every MDCT line (tb.bark staged: pass 2 reads it at the head of a dependent chain)
    const cpx<T>* stL;  // stage tables of the L-point transform (Hann spectra) and of the L/2-point one (MDCT), power-of-two
    const cpx<T>* stQ;  // L only (mrc_fft.cuh: fft_sw)
    int* pbin;      // [Q]      peak bins; once the masker tables are built the same words hold zlut
    uint16_t* zlut; // [MRC_ZLUT+1] number of maskers with z < g/32 Bark (only when Q ints can hold it)
    int* lcnt;      // [Q+1]    number of loud maskers (g > 0) below index i
    uint16_t* lidx; // [Q]      their indices, ascending
    T* mkey;        // [2][1024] merge buffers of the grant order (alias `lines` when it is large enough)
    uint16_t* mid;  // [2][1024]
    double* etab;   // [64]     2^(j/64)
};

// stage-table entries (complex) a block of L lines keeps in shared memory
__host__ __device__ constexpr int stage_entries_of(int L) {
    if (L & (L - 1)) return 0;
    int lg = 0;
    while ((1 << lg) < L) ++lg;
    return fft_stage_entries(lg) + fft_stage_entries(lg - 1);
}

template <typename T, bool XIN>
__device__ __forceinline__ Smem<T> carve(unsigned char* raw, int L) {
    Smem<T> s;
    const int Q = L / 2;
    T* p = reinterpret_cast<T*>(raw);
    const bool lin = sizeof(T) == 4 && !(L & (L - 1));
    s.xi4 = lin ? p : nullptr;
    if constexpr (XIN) { s.sx = p; s.px = nullptr; p += 4 * L; }
    else { s.sx = nullptr; s.px = reinterpret_cast<uint32_t*>(p); p += lin ? 4 * L : (2 * L * 4) / (int)sizeof(T); }
    s.buf = reinterpret_cast<cpx<T>*>(p); p += 2 * L;
    s.lines = p;         p += 4 * L;
    s.xi = p;            p += L;
    s.buf2 = reinterpret_cast<cpx<T>*>(s.xi);     // L + 2Q = 2L values: xi, mz, ms15
    s.mz = p;            p += Q;
    s.ms15 = p;          p += Q;
    s.mg = p;            p += Q;
    {
        int lg = 0;
        while ((1 << lg) < L) ++lg;
        s.stL = reinterpret_cast<const cpx<T>*>(p);
        s.stQ = s.stL + ((L & (L - 1)) ? 0 : fft_stage_entries(lg));
        p += 2 * stage_entries_of(L);
    }
    s.zb = p;            p += L;
    s.li = reinterpret_cast<LineInfo*>(s.xi);
    s.rng = nullptr;
    if (sizeof(T) == 4) { s.rng = reinterpret_cast<uint32_t*>(p); p += L; }
    // mc, mU, mS, mP (4Q values = the FFT work buffer's 2L) live in the FFT work buffer: it is idle while maskers are
    // spread
    T* r = reinterpret_cast<T*>(s.buf);
    s.mc = r;
    s.mU = r + Q;
    s.mS = r + 2 * Q;
    s.mP = r + 3 * Q;
    double* d = reinterpret_cast<double*>((reinterpret_cast<size_t>(p) + 7) & ~(size_t)7);
    s.etab = d;          d += 64;
    int* ip = reinterpret_cast<int*>(d);
    s.pbin = ip;         ip += Q;
    s.zlut = (Q * 4 >= (MRC_ZLUT + 2) * 2) ? reinterpret_cast<uint16_t*>(s.pbin) : nullptr;
    s.lcnt = ip;         ip += Q + 1;
    s.lidx = reinterpret_cast<uint16_t*>(ip);
    const size_t mbytes = 2048 * sizeof(T) + 2048 * 2;
    if ((size_t)(4 * L) * sizeof(T) >= mbytes) s.mkey = s.lines;
    else {
        const size_t a = (reinterpret_cast<size_t>(s.lidx) + (size_t)Q * 2 + 15) & ~(size_t)15;
        s.mkey = reinterpret_cast<T*>(a);
    }
    s.mid = reinterpret_cast<uint16_t*>(s.mkey + 2048);
#ifdef MRC_DEBUG_ASSERTS
    {   // everything carved must lie inside the dynamic shared memory the launch asked for
        unsigned dyn;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
        const unsigned char* end1 = reinterpret_cast<const unsigned char*>(s.lidx + Q);
        const unsigned char* end2 = reinterpret_cast<const unsigned char*>(s.mid + 2048);
        MRC_ASSERT(end1 <= raw + dyn && end2 <= raw + dyn);
        MRC_ASSERT(reinterpret_cast<const unsigned char*>(s.etab + 64) <= reinterpret_cast<const unsigned char*>(s.pbin));
    }
#endif
    return s;
}

template <typename T>
__device__ __forceinline__ void li_store(const Smem<T>& sm, int k, float ub, uint32_t rng) {
    if constexpr (sizeof(T) == 8) { LineInfo v; v.ub = ub; v.rng = rng; sm.li[k] = v; }
    else { sm.xi[k] = ub; sm.rng[k] = rng; }
}
template <typename T>
__device__ __forceinline__ float li_ub(const Smem<T>& sm, int k) {
    if constexpr (sizeof(T) == 8) return sm.li[k].ub; else return sm.xi[k];
}
template <typename T>
__device__ __forceinline__ uint32_t li_rng(const Smem<T>& sm, int k) {
    if constexpr (sizeof(T) == 8) return sm.li[k].rng; else return sm.rng[k];
}

template <typename T>
__device__ __forceinline__ T warp_max(T v) {
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// time sample n of spectrum c (0 L, 1 R, 2 M, 3 S) -- codecThem.py:363-364 -- from the block's PCM frames (converted here:
// the frames stay packed in shared memory, 8 KB instead of 32 KB of doubles and one 4-byte read per frame) or, on the
// per-block seam, from the doubles the caller handed in
template <typename T, bool XIN>
__device__ __forceinline__ void frame_lr(const Smem<T>& sm, int N, int n, T& l, T& r) {
    if constexpr (XIN) { l = sm.sx[n]; r = sm.sx[N + n]; }
    else {
        const uint32_t w = sm.px[n];
        l = pcm_to_fraction<T>((int)(short)(w & 0xffffu));
        r = pcm_to_fraction<T>((int)(short)(w >> 16));
    }
}
template <typename T>
__device__ __forceinline__ T spec_of(int c, T l, T r) {
    if (c == 0) return l;
    if (c == 1) return r;
    if (c == 2) return (l + r) / T(2);
    return (l - r) / T(2);
}
template <typename T, bool XIN>
__device__ __forceinline__ T tsample(const Smem<T>& sm, int N, int c, int n) {
    if constexpr (XIN) return spec_of<T>(c, sm.sx[n], sm.sx[N + n]);
    else {
        const uint32_t w = sm.px[n];
        if (c == 0) return pcm_to_fraction<T>((int)(short)(w & 0xffffu));
        if (c == 1) return pcm_to_fraction<T>((int)(short)(w >> 16));
        return spec_of<T>(c, pcm_to_fraction<T>((int)(short)(w & 0xffffu)), pcm_to_fraction<T>((int)(short)(w >> 16)));
    }
}
// two neighbouring table values with one load (i even)
__device__ __forceinline__ void ld2(const double* p, int i, double& a, double& b) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p + i));
    a = v.x; b = v.y;
}
__device__ __forceinline__ void ld2(const float* p, int i, float& a, float& b) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p + i));
    a = v.x; b = v.y;
}

// position of line k among the maskers: m_lo = number of maskers with dz > 0.5 (a prefix: z_m ascends),
// m_hi = first masker with dz < -0.5, dz = z_k - z_m as the reference computes it
template <typename T>
__device__ __forceinline__ void masker_range(const Smem<T>& sm, T zk, int npk, int& m_lo, int& m_hi) {
    if (sm.zlut != nullptr) {
        // start from the count table (maskers per 1/32 Bark cell, prefix-summed), then settle with the exact
        // comparisons: a cell holds two or three maskers at most
        int g = (int)((zk - T(0.5)) * T(32.0));
        int m = sm.zlut[g < 0 ? 0 : (g > MRC_ZLUT ? MRC_ZLUT : g)];
        while (m > 0 && !(zk - sm.mz[m - 1] > T(0.5))) --m;
        while (m < npk && zk - sm.mz[m] > T(0.5)) ++m;
        m_lo = m;
        g = (int)((zk + T(0.5)) * T(32.0)) + 1;
        m = sm.zlut[g < 0 ? 0 : (g > MRC_ZLUT ? MRC_ZLUT : g)];
        if (m < m_lo) m = m_lo;
        while (m > m_lo && zk - sm.mz[m - 1] < T(-0.5)) --m;
        while (m < npk && !(zk - sm.mz[m] < T(-0.5))) ++m;
        m_hi = m;
        return;
    }
    int lo = 0, hi = npk;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (zk - sm.mz[mid] > T(0.5)) lo = mid + 1; else hi = mid;
    }
    m_lo = lo;
    hi = npk;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (zk - sm.mz[mid] < T(-0.5)) hi = mid; else lo = mid + 1;
    }
    m_hi = lo;
}

// one loud masker (index m) onto a line more than 0.5 Bark above it: the reference's exponent, unfused
template <typename T>
__device__ __forceinline__ T loud_term(const Smem<T>& sm, T zk, int m) {
    const T t = rn_add(rn_add(zk, -sm.mz[m]), T(-0.5));
    const T e = rn_add(rn_add(sm.ms15[m], rn_mul(T(-27.0), t)), rn_mul(sm.mg[m], t));
    return sp_exp10(sp_div10(rn_add(e, T(-96.0))), sm.etab);
}

// the two geometric tails at line k
template <typename T>
__device__ __forceinline__ T tail_terms(const Smem<T>& sm, T zk, int npk, int m_lo, int m_hi) {
    T a = T(0.0);
    if (m_lo > 0) a += sm.mU[m_lo - 1] * sp_exp10(T(-2.7) * ((zk - sm.mz[m_lo - 1]) - T(0.5)), sm.etab);
    if (m_hi < npk) a += sm.mS[m_hi] * sp_exp10(T(-2.7) * ((sm.mz[m_hi] - zk) - T(0.5)), sm.etab);
    return a;
}

// Masked threshold intensity at MDCT line k from the factorised masker tables (same mathematics as
// psychoac.py:68-78 summed over all maskers, re-associated):
//   maskers more than 0.5 Bark below the line   quiet ones (g = 0): one geometric tail U, decayed from the nearest
//                                               loud ones (slope depends on their level): one 10**x each
//   maskers within +-0.5 Bark                   their plateau intensities, summed directly
//   maskers more than 0.5 Bark above the line   one geometric tail S, decayed from the nearest
// The region of every (masker, line) pair is decided by the reference's own comparisons on dz = z_k - z_m.
//
// spread_line_bound: a LOWER bound of that threshold, cheap enough for every line: only the MRC_NEAR_LOUD nearest
// loud maskers, and the plateau sum as a difference of prefix sums minus its worst-case rounding error.
template <typename T>
__device__ __forceinline__ T spread_line_bound(const Smem<T>& sm, const DevTables<T>& tb, int k, int npk,
                                                    unsigned& n_general, int& m_lo, int& m_hi) {
    const T zk = sm.zb[k];
    masker_range(sm, zk, npk, m_lo, m_hi);
    T a = tb.quiet[k] + tail_terms(sm, zk, npk, m_lo, m_hi);
    const int nl = sm.lcnt[m_lo];
    const int j0 = nl > MRC_NEAR_LOUD ? nl - MRC_NEAR_LOUD : 0;
    for (int j = nl - 1; j >= j0; --j) a += loud_term(sm, zk, sm.lidx[j]);
    n_general += (unsigned)(nl - j0);
    // worst-case rounding of the two prefix sums: 2^-46 of the total in fp64, 4e-5 in fp32
    const T w = (sm.mP[m_hi] - sm.mP[m_lo]) - (sizeof(T) == 8 ? T(1.4210854715202004e-14) : T(4e-5)) * sm.mP[npk];
    return a + fmax(w, T(0.0));
}

// Complete thresholds are evaluated by GROUPS of GW lanes.  GW = 32 (a warp per line) measured best: with two groups of
// 16 to a warp every band of a spectrum fits in one round and two lines share every instruction issued, but the 16-wide
// window of masker_range_grp misses its bound more often (scalar search) and the phase is bound by its slowest band either
// way (profiles/r02_phase_clocks.log: a warp waits 12 k cycles per spectrum for the slowest group after 7.5 k of its own).
constexpr int GW = 32;
struct Grp {
    unsigned mask;      // the group's lanes inside its warp
    int gl;             // this thread's lane inside the group
    int shift;          // the group's first lane
};
__device__ __forceinline__ unsigned grp_ballot(const Grp& g, bool p) {
    return (__ballot_sync(g.mask, p) >> g.shift) & ((GW == 32) ? 0xffffffffu : ((1u << (GW & 31)) - 1u));
}

// The same threshold, complete, evaluated by a group of lanes.  Everything that goes through 10**x is one list of
// items dealt out to the lanes -- item 0 the tail of the maskers more than 0.5 Bark below the line, item 1 the tail of
// those above, items 2.. the loud maskers below -- so that one pass of the exponential covers a typical line; the
// plateau maskers are summed lane-strided, lane 2 adds the threshold in quiet, then a butterfly sum.  All lanes of the
// group return it.
template <typename T>
__device__ __forceinline__ T spread_line_grp(const Smem<T>& sm, const DevTables<T>& tb, int k, int npk, const Grp& g,
                                             unsigned& n_general, unsigned& n_window) {
    MRC_WCLK_BEGIN();
    const T zk = sm.zb[k];
    const T quiet = tb.quiet[k];
    const uint32_t rng = li_rng<T>(sm, k);       // the line's masker range, found in pass 1
    const int m_lo = (int)(rng & 0xffffu), m_hi = (int)(rng >> 16);
    MRC_WCLK(16);
    const int nl = sm.lcnt[m_lo];
    T a = T(0.0);
    for (int base = 0; base < nl + 2; base += GW) {
        const int item = base + g.gl, j = item - 2;
        const bool tail = item < 2;
        const bool valid = tail ? (item == 0 ? m_lo > 0 : m_hi < npk) : j < nl;
        int mi = tail ? (item == 0 ? m_lo - 1 : m_hi) : (int)sm.lidx[j < nl ? j : 0];
        mi = valid ? mi : 0;
        const T mzv = sm.mz[mi];
        const T t = rn_add(item == 1 ? rn_add(mzv, -zk) : rn_add(zk, -mzv), T(-0.5));
        // loud masker: the reference's exponent, unfused (psychoac.py:70-76); tails: -27 dB per Bark beyond the plateau
        const T e = rn_add(rn_add(sm.ms15[mi], rn_mul(T(-27.0), t)), rn_mul(sm.mg[mi], t));
        const T y = tail ? T(-2.7) * t : sp_div10(rn_add(e, T(-96.0)));
        const T coef = tail ? (item == 0 ? sm.mU[mi] : sm.mS[mi]) : T(1.0);
        const T v = coef * sp_exp10(y, sm.etab);
        if (valid) a += v;
    }
    MRC_WCLK(17);
    for (int m = m_lo + g.gl; m < m_hi; m += GW) a += sm.mc[m];
    if (g.gl == 2) a += quiet;
    MRC_WCLK(18);
    if (g.gl == 0) { n_general += (unsigned)nl; n_window += (unsigned)(m_hi - m_lo); }
#pragma unroll
    for (int o = GW / 2; o; o >>= 1) a += __shfl_xor_sync(g.mask, a, o, GW);
    MRC_WCLK(20);
    return a;
}

// what the last warp leaves for the CTA's NEXT block while the current one is finished: the block's place in its clip
// and, when its 2L frames lie inside the clip at a 16-byte aligned address, the frames themselves (one bulk copy into
// `px`, which the current block no longer reads by then)
struct NextBlock {
    int state;               // 0 nothing, 1 place + frames on their way (mbarrier), 2 place only
    int clip, b, nblk_clip;
};

// One block: `it` = its index in this launch (wave-local, or into the geometry's list), it_next = the CTA's next one (-1:
// none).
template <typename T, int L_, bool XIN>
__device__ __forceinline__ void analysis_block(const Smem<T>& sm, const DevTables<T>& tb, const CodecParams& cp,
                                               const ClipMap& cm, const int16_t* __restrict__ pcm,
                                               const double* __restrict__ xin, int g0, int it, int it_next,
                                               const Handoff<T>& ho, const AnalysisTaps<T>& taps,
                                               unsigned long long* peak_counter, unsigned long long* mbar,
                                               NextBlock* nxt, unsigned& mbar_parity) {
    constexpr int L = L_, N = 2 * L, Q = L / 2, NT = Q, nwarp = NT >> 5;
    static_assert(NT % 32 == 0, "whole warps");
    const int nb = tb.nb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool POW2 = FftShape<L>::pow2;
    constexpr int LOGL = FftShape<L>::logP;          // log2 L (power-of-two L only)
    // Fast mode (fp32, power-of-two L): the transforms are linear, so the MDCT lines and the Hann spectra of M and S come
    // from those of L and R -- two MDCTs and two FFTs per joint block instead of four and four.  The rounding of the S
    // spectrum is then relative to L and R instead of to S itself, which is what separates this mode from the exact one.
    constexpr bool LIN = POW2 && sizeof(T) == 4;

    __shared__ int s_clip, s_b, s_nblk_clip;
    __shared__ T s_red[4][32];
    __shared__ T s_smr[4][MRC_BSTRIDE];
    __shared__ int s_scale[4];
    __shared__ unsigned int s_ms;
    __shared__ int s_wcnt[33];
    __shared__ T s_scan[5][32];
    __shared__ T s_band_smr[MRC_BSTRIDE];
    __shared__ unsigned s_best[MRC_BSTRIDE];     // per band: highest pass-1 bound (float bits, low 11 bits = 2047 - line)
    __shared__ int s_npk;
#ifdef MRC_PHASE_CLOCKS
    __shared__ long long s_clk_last;
    if (threadIdx.x == 0) s_clk_last = clock64();
#endif

    const int lb = cm.list ? cm.list[it] : it;   // index inside this wave's hand-off buffers
    const int g = g0 + lb;                    // global block index
    const int pre = nxt->state;               // left by this CTA's previous block (published by the barrier that closed it)
    if (tid == 0) {
        if (pre) { s_clip = nxt->clip; s_b = nxt->b; s_nblk_clip = nxt->nblk_clip; }
        else {
            int lo = 0, hi = cm.n_clips;      // last clip whose first block <= g
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (cm.clip_blk0[mid] <= g) lo = mid; else hi = mid;
            }
            s_clip = lo;
            s_b = g - cm.clip_blk0[lo];
            s_nblk_clip = cm.clip_blk0[lo + 1] - cm.clip_blk0[lo];
        }
        s_ms = 0u;
    }
    __syncthreads();
    const int b = s_b;
    const bool is_flush = (b == s_nblk_clip - 1);
    const bool joint = cp.joint && !(cp.flush_nonjoint && is_flush);
    const int nspec = joint ? 4 : 2;

    // ---- phase 0: frame [ (b-1)L, (b+1)L ) of the clip, zero outside ------------------------------------
    if constexpr (XIN) {
        for (int i = tid; i < 2 * N; i += NT) sm.sx[i] = T(xin[(size_t)g * 2 * N + i]);
    } else if (pre == 1) {
        mbar_wait(mbar, mbar_parity);         // the frames were fetched while the previous block was finished
        mbar_parity ^= 1u;
    } else {
        const long long frames = cm.clip_off[s_clip + 1] - cm.clip_off[s_clip];
        const uint32_t* __restrict__ p32 = reinterpret_cast<const uint32_t*>(pcm) + (cm.clip_off[s_clip] - cm.pcm_frame0);
        // the block's a "prior" samples then its b new ones (pacfileThem.py:799-802)
        const long long s0 = cm.blk_start ? cm.blk_start[g] - tb.a : (long long)(b + cm.blk_base - 1) * L;
        for (int n = tid; n < N; n += NT) {
            const long long s = s0 + n;
            uint32_t w = 0;
            if (s >= 0 && s < frames) w = __ldg(p32 + s);
            sm.px[n] = w;
        }
    }
    __syncthreads();
    MRC_CLK(0);

    // ---- phase 1: KBD window + MDCT, two spectra at a time (each an L/2-point complex FFT) --------------
    // Power-of-two L: swizzled work buffers and per-stage twiddle tables (mrc_fft.cuh, fft_sw).  9 * 2^p lines (the
    // transition blocks of block switching): the root table staged in `xi` (idle until the first intensities are written).
    cpx<T>* const tws = reinterpret_cast<cpx<T>*>(sm.xi);
    const int ntw = 1 << (tb.logLtab - 1);
    if constexpr (!POW2)
        for (int i = tid; i < ntw; i += NT) tws[i] = tb.tw_fft[i];   // published by the barrier before the first FFT
    {
        const int grp = tid / (NT / 2), lt = tid - grp * (NT / 2), gthr = NT / 2;
        const T two_over_n = T(2) / T(N);
        for (int pair = 0; pair < (LIN ? 1 : nspec / 2); ++pair) {
            const int c = pair * 2 + grp;
            cpx<T>* a = sm.buf + grp * Q;
            if constexpr (POW2) {
                // one input index per thread (NT == Q) for BOTH spectra of the pair: the four frames and window values it
                // needs are read once.  With y(i) = window[i] * x[i] extended antiperiodically (y(i +- N) = -y(i)):
                //   re = -y(3Q-1-2n) - y(3Q+2n),  im = y(Q-1-2n) - y(Q+2n)      (mdct.py:64-70 reduced to N/4 points)
                const int n = fft_place_index<LOGL - 1>(tid);
                int iA = 3 * Q + 2 * n, iB = Q - 1 - 2 * n;
                T sA = T(-1), sB = T(1);
                if (iA >= N) { iA -= N; sA = T(1); }
                if (iB < 0) { iB += N; sB = T(-1); }
                const int i1 = 3 * Q - 1 - 2 * n, i4 = Q + 2 * n;
                T l1, r1, lA, rA, lB, rB, l4, r4;
                frame_lr<T, XIN>(sm, N, i1, l1, r1);
                frame_lr<T, XIN>(sm, N, iA, lA, rA);
                frame_lr<T, XIN>(sm, N, iB, lB, rB);
                frame_lr<T, XIN>(sm, N, i4, l4, r4);
                const T k1 = tb.kbd[i1], kA = tb.kbd[iA], kB = tb.kbd[iB], k4 = tb.kbd[i4];
                const cpx<T> w = tb.tw_pre[n];
                const int r = fft_swz<T>(fft_r4_pos(n, LOGL - 1));
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int cc = pair * 2 + h;
                    const T re = fma(sA, kA * spec_of<T>(cc, lA, rA), -(k1 * spec_of<T>(cc, l1, r1)));
                    const T im = fma(sB, kB * spec_of<T>(cc, lB, rB), -(k4 * spec_of<T>(cc, l4, r4)));
                    cpx<T> v;
                    v.x = re * w.x - im * w.y;
                    v.y = re * w.y + im * w.x;
                    sm.buf[h * Q + r] = v;
                }
            } else {
                for (int n = lt; n < Q; n += gthr) {
                    T re, im;
                    // n0 = (b+1)/2 (mdct.py:66) is the standard phase N/4 + 1/2 shifted by rot = (a-b)/4 samples: the
                    // transform of the sequence rotated by rot, wrapped samples negated (the kernel is antiperiodic)
                    auto y = [&](int i) {
                        int j = i + tb.rot;
                        T sg = T(1);
                        if (j >= N) { j -= N; sg = T(-1); }
                        else if (j < 0) { j += N; sg = T(-1); }
                        return sg * (tb.kbd[j] * tsample<T, XIN>(sm, N, c, j));
                    };
                    if (n < Q / 2) {
                        re = -y(3 * Q - 1 - 2 * n) - y(3 * Q + 2 * n);
                        im = y(Q - 1 - 2 * n) - y(Q + 2 * n);
                    } else {
                        re = y(2 * n - Q) - y(3 * Q - 1 - 2 * n);
                        im = -y(Q + 2 * n) - y(5 * Q - 1 - 2 * n);
                    }
                    const cpx<T> w = tb.tw_pre[n];
                    const int r = fft_pos<Q>(n);
                    a[r].x = re * w.x - im * w.y;
                    a[r].y = re * w.y + im * w.x;
                }
            }
            __syncthreads();
            MRC_CLK(13);
            if constexpr (POW2) fft_sw<T, LOGL - 1>(a, lt, gthr, sm.stQ);
            else fft_any<T, Q>(a, lt, gthr, tws, tb.logLtab, tb.tw9, L, tb.w9);
            MRC_CLK(14);
            T* X = sm.lines + c * L;
            for (int k = lt; k < Q; k += gthr) {
                const cpx<T> w = tb.tw_post[k];
                const cpx<T> t = a[POW2 ? fft_swz<T>(k) : k];
                X[2 * k] = two_over_n * (t.x * w.x - t.y * w.y);
                X[L - 1 - 2 * k] = -two_over_n * (t.x * w.y + t.y * w.x);
            }
            __syncthreads();
            MRC_CLK(1);
        }
        if constexpr (LIN) {
            if (nspec == 4) {
                for (int k = tid; k < L; k += NT) {
                    const T l = sm.lines[k], r = sm.lines[L + k];
                    sm.lines[2 * L + k] = (l + r) / T(2);
                    sm.lines[3 * L + k] = (l - r) / T(2);
                }
                __syncthreads();
            }
        }
    }

    if (taps.lines4 != nullptr) {
        T* o = taps.lines4 + (size_t)lb * 4 * cp.Lmax;           // rows of Lmax, zero beyond this geometry's L
        for (int i = tid; i < 4 * cp.Lmax; i += NT) {
            const int c = i / cp.Lmax, k = i - c * cp.Lmax;
            o[i] = (c < nspec && k < L) ? sm.lines[c * L + k] : T(0);
        }
    }

    // ---- phase 2: ms_switch on the unscaled L/R lines (ms_stereo.py:5-27) -------------------------------
    if (joint) {
        for (int bd = warp; bd < nb; bd += nwarp) {
            const int lo = tb.c_band_lo[bd], n = tb.c_band_n[bd];
            T sd = 0, ss = 0;
            for (int i = lane; i < n; i += 32) {
                const T l = sm.lines[lo + i], r = sm.lines[L + lo + i];
                const T l2 = l * l, r2 = r * r;
                sd += fabs(l2 - r2);
                ss += fabs(l2 + r2);
            }
            sd = warp_sum(sd);
            ss = warp_sum(ss);
            if (lane == 0 && sd < T(0.8) * ss) atomicOr(&s_ms, 1u << bd);
        }
    }

    // ---- phase 3: overall scale factors (quantize.py:114-146 with nMantBits=5), lines *= 2^scale --------
    {
        T mx[4] = {0, 0, 0, 0};
        for (int i = tid; i < L; i += NT)
            for (int c = 0; c < nspec; ++c) mx[c] = fmax(mx[c], fabs(sm.lines[c * L + i]));
        for (int c = 0; c < nspec; ++c) {
            const T v = warp_max(mx[c]);
            if (lane == 0) s_red[c][warp] = v;
        }
        __syncthreads();
        if (warp == 0) {
            for (int c = 0; c < nspec; ++c) {
                T v = (lane < nwarp) ? s_red[c][lane] : T(0);
                v = warp_max(v);
                if (lane == 0) s_scale[c] = scale_factor_of((double)v, cp.n_scale_bits, 5);
            }
            if (lane == 0 && nspec == 2) s_scale[2] = s_scale[3] = 0;
        }
        __syncthreads();
        MRC_CLK(2);
        for (int i = tid; i < L; i += NT)
            for (int c = 0; c < nspec; ++c) sm.lines[c * L + i] *= T(1 << s_scale[c]);
        // (no sync needed yet: the next reader of `lines` is after several barriers)
    }

    // ---- phase 4: psychoacoustic model per spectrum ------------------------------------------------------
    const T xi_den = T(N) * T(N) * T(0.375);
    int my_peaks = 0;
    unsigned n_general = 0, n_window = 0, n_loud = 0;
    // OverallSMRs (ms_stereo.py:70-81) keeps, per band, the SMRs of (M, S) where ms_switch is set and of (L, R) where
    // it is not, and drops the other pair.  So a spectrum's SMR is only needed in the bands that select it: the
    // others are not evaluated (all of the spectrum's model when no band selects it).  The stage taps and the
    // reference-order mode evaluate everything.
    const bool all_bands = !joint || cp.spread_seq || taps.smr4 != nullptr || taps.npeaks != nullptr;
    const unsigned band_mask = (nb >= 32) ? 0xffffffffu : ((1u << nb) - 1u);
    if constexpr (LIN) {
        // Hann spectra of L and R side by side (half the threads each), then the intensities of all the spectra at once
        for (int it = 0; it < L / NT; ++it) {
            const int n = fft_place_index<LOGL>(tid + it * NT);
            T h0, h1, l0, r0, l1, r1;
            ld2(tb.hann, 2 * n, h0, h1);
            frame_lr<T, XIN>(sm, N, 2 * n, l0, r0);
            frame_lr<T, XIN>(sm, N, 2 * n + 1, l1, r1);
            const int r = fft_swz<T>(fft_r4_pos(n, LOGL));
            cpx<T> v;
            v.x = h0 * l0; v.y = h1 * l1;
            sm.buf[r] = v;
            v.x = h0 * r0; v.y = h1 * r1;
            sm.buf2[r] = v;
        }
        __syncthreads();
        MRC_CLK(15);
        {
            const int grp = tid / (NT / 2), lt = tid - grp * (NT / 2);
            fft_sw<T, LOGL>(grp ? sm.buf2 : sm.buf, lt, NT / 2, sm.stL);
        }
        MRC_CLK(23);
        // X[k] = E[k] + W^k O[k] for L and R; M = (L + R)/2, S = (L - R)/2;  XI = 4|X|^2 / (N^2 * 3/8)   (psychoac.py:151)
        for (int k = tid; k < L; k += NT) {
            const int pk = fft_swz<T>(k), pc = fft_swz<T>(k ? L - k : 0);
            const cpx<T> w = tb.tw_rfft[k];
            T xr[2], xim[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const cpx<T>* z = h ? sm.buf2 : sm.buf;
                const cpx<T> zk = z[pk], zc = z[pc];
                const T ex = (zk.x + zc.x) * T(0.5), ey = (zk.y - zc.y) * T(0.5);
                const T dx = zk.x - zc.x, dy = zk.y + zc.y;           // D = Zk - conj(Zc)
                const T ox = dy * T(0.5), oy = -dx * T(0.5);          // O = D / (2j)
                xr[h] = ex + (ox * w.x - oy * w.y);
                xim[h] = ey + (ox * w.y + oy * w.x);
            }
            sm.xi4[k] = T(4) * (xr[0] * xr[0] + xim[0] * xim[0]) / xi_den;
            sm.xi4[L + k] = T(4) * (xr[1] * xr[1] + xim[1] * xim[1]) / xi_den;
            if (nspec == 4) {
                const T mr = (xr[0] + xr[1]) * T(0.5), mi = (xim[0] + xim[1]) * T(0.5);
                const T sr = (xr[0] - xr[1]) * T(0.5), si = (xim[0] - xim[1]) * T(0.5);
                sm.xi4[2 * L + k] = T(4) * (mr * mr + mi * mi) / xi_den;
                sm.xi4[3 * L + k] = T(4) * (sr * sr + si * si) / xi_den;
            }
        }
        __syncthreads();
        MRC_CLK(3);
    }
    for (int c = 0; c < nspec; ++c) {
        const unsigned need = all_bands ? band_mask : ((c < 2 ? ~s_ms : s_ms) & band_mask);
        if (need == 0u) {
            if (tid < nb) s_smr[c][tid] = T(0);
            continue;                                    // uniform: s_ms is shared
        }
        if constexpr (!LIN) {
            // a. Hann window, real 2L-point FFT through an L-point complex FFT
            if constexpr (POW2) {
                for (int it = 0; it < L / NT; ++it) {
                    const int n = fft_place_index<LOGL>(tid + it * NT);
                    T h0, h1, x0, x1;
                    ld2(tb.hann, 2 * n, h0, h1);
                    if constexpr (XIN) {
                        x0 = tsample<T, XIN>(sm, N, c, 2 * n);
                        x1 = tsample<T, XIN>(sm, N, c, 2 * n + 1);
                    } else {
                        const uint2 w = *reinterpret_cast<const uint2*>(sm.px + 2 * n);
                        const int la = (int)(short)(w.x & 0xffffu), ra = (int)(short)(w.x >> 16);
                        const int lb2 = (int)(short)(w.y & 0xffffu), rb = (int)(short)(w.y >> 16);
                        if (c == 0) { x0 = pcm_to_fraction<T>(la); x1 = pcm_to_fraction<T>(lb2); }
                        else if (c == 1) { x0 = pcm_to_fraction<T>(ra); x1 = pcm_to_fraction<T>(rb); }
                        else {
                            x0 = spec_of<T>(c, pcm_to_fraction<T>(la), pcm_to_fraction<T>(ra));
                            x1 = spec_of<T>(c, pcm_to_fraction<T>(lb2), pcm_to_fraction<T>(rb));
                        }
                    }
                    cpx<T> v;
                    v.x = h0 * x0;
                    v.y = h1 * x1;
                    sm.buf[fft_swz<T>(fft_r4_pos(n, LOGL))] = v;
                }
            } else {
                for (int i = tid; i < ntw; i += NT) tws[i] = tb.tw_fft[i];   // xi held the previous spectrum's per-line scratch
                for (int n = tid; n < L; n += NT) {
                    const int r = fft_pos<L>(n);
                    sm.buf[r].x = tb.hann[2 * n] * tsample<T, XIN>(sm, N, c, 2 * n);
                    sm.buf[r].y = tb.hann[2 * n + 1] * tsample<T, XIN>(sm, N, c, 2 * n + 1);
                }
            }
            __syncthreads();
            MRC_CLK(15);
            if constexpr (POW2) fft_sw<T, LOGL>(sm.buf, tid, NT, sm.stL);
            else fft_any<T, L>(sm.buf, tid, NT, tws, tb.logLtab, tb.tw9, L, tb.w9);
            MRC_CLK(23);
            // b. X[k] = E[k] + W^k O[k];  XI = 4|X|^2 / (N^2 * 3/8)   (psychoac.py:151)
            for (int k = tid; k < L; k += NT) {
                const cpx<T> zk = sm.buf[POW2 ? fft_swz<T>(k) : k], zc = sm.buf[POW2 ? fft_swz<T>(k ? L - k : 0) : (k ? L - k : 0)];
                const T ex = (zk.x + zc.x) * T(0.5), ey = (zk.y - zc.y) * T(0.5);
                const T dx = zk.x - zc.x, dy = zk.y + zc.y;           // D = Zk - conj(Zc)
                const T ox = dy * T(0.5), oy = -dx * T(0.5);          // O = D / (2j)
                const cpx<T> w = tb.tw_rfft[k];
                const T xr = ex + (ox * w.x - oy * w.y), xim = ey + (ox * w.y + oy * w.x);
                sm.xi[k] = T(4) * (xr * xr + xim * xim) / xi_den;
            }
            __syncthreads();
            MRC_CLK(3);
        }
        const T* const xic = LIN ? sm.xi4 + c * L : sm.xi;       // this spectrum's intensities
        // c. strict local maxima at bins 1 .. L-102, ascending order (psychoac.py:158-170)
        {
            int found = -1;
            for (int t = tid; t < Q; t += NT) {       // NT == Q: one trip
                const int p0 = 2 * t, p1 = p0 + 1;
                if (p0 >= 1 && p0 <= L - 102 && xic[p0] > xic[p0 - 1] && xic[p0] > xic[p0 + 1]) found = p0;
                if (p1 <= L - 102 && xic[p1] > xic[p1 - 1] && xic[p1] > xic[p1 + 1]) found = p1;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, found >= 0);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            __syncthreads();
            if (warp == 0) {
                int v = (lane < nwarp) ? s_wcnt[lane] : 0;
                int incl = v;
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                s_wcnt[lane] = incl - v;
                if (lane == 31) s_npk = incl;
            }
            __syncthreads();
            MRC_ASSERT(s_npk <= Q);
            if (found >= 0) sm.pbin[s_wcnt[warp] + __popc(bal & ((1u << lane) - 1u))] = found;
            __syncthreads();
            MRC_CLK(4);
        }
        const int npk = s_npk;
        // d. masker parameters (psychoac.py:163-165, :37-49), in double in both modes
        {
            const int i = tid;                       // npk <= Q == NT: one masker per thread
            T z = 0, s15 = 0, g = 0, cmid = 0;
            bool loud = false;
            if (i < npk) {
                // products and sums kept unfused (__dmul_rn/__dadd_rn), in the reference's order
                const int p = sm.pbin[i];
                const T x0 = xic[p - 1], x1 = xic[p], x2 = xic[p + 1];
                const T sum = rn_add(rn_add(x0, x1), x2);
                const T spl = fmax(rn_add(T(96.0), rn_mul(T(10.0), m_log10(sum))), T(-30.0));
                const T num = rn_add(rn_add(rn_mul(T(p - 1), x0), rn_mul(T(p), x1)),
                                             rn_mul(T(p + 1), x2));
                const T f = rn_div(rn_mul(T(tb.fstep), num), sum);
                const T fq = rn_div(f, T(7500.0));
                z = rn_add(rn_mul(T(13.0), m_atan(rn_div(rn_mul(T(0.76), f), T(1000.0)))),
                              rn_mul(T(3.5), m_atan(rn_mul(fq, fq))));
                s15 = rn_add(spl, T(-15.0));
                g = rn_mul(T(0.37), fmax(rn_add(spl, T(-40.0)), T(0.0)));
                loud = g > T(0.0);
                n_loud += loud ? 1u : 0u;
                sm.mz[i] = z; sm.ms15[i] = s15; sm.mg[i] = g;
                if (!cp.spread_seq) {
                    cmid = sp_exp10(sp_div10(rn_add(s15, T(-96.0))), sm.etab);
                    sm.mc[i] = cmid;
                }
            }
            if (tid == 0) my_peaks += npk;
            if (!cp.spread_seq) {
                // ordered list of the loud maskers (g > 0: their upper slope depends on their level)
                const unsigned bal = __ballot_sync(0xffffffffu, loud);
                __syncthreads();                     // publishes mz for the neighbour reads below
                // decay between neighbouring maskers: rho(d) = 10^(-2.7 d), the -27 dB/Bark slope of both sides
                T rU = T(0.0), rS = T(0.0);           // rU = rho(z_i - z_{i-1}); rS = rho(z_{i+1} - z_i)
                if (i < npk) {
                    if (i > 0) rU = sp_exp10(T(-2.7) * (z - sm.mz[i - 1]), sm.etab);
                    if (i + 1 < npk) rS = sp_exp10(T(-2.7) * (sm.mz[i + 1] - z), sm.etab);
                }
                // two affine recurrences and a plain sum by warp scans:
                //   U_i = cq_i + rU_i * U_{i-1}   ascending, quiet maskers only     (maps composed with shfl_up)
                //   S_i = c_i  + rS_i * S_{i+1}   descending, all maskers          (maps composed with shfl_down)
                //   P_i = c_i  + P_{i-1}
                T aU = rU, bU = (i < npk && !loud) ? cmid : T(0.0);
                T aS = rS, bS = (i < npk) ? cmid : T(0.0);
                T pP = (i < npk) ? cmid : T(0.0);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const T alU = __shfl_up_sync(0xffffffffu, aU, o), blU = __shfl_up_sync(0xffffffffu, bU, o);
                    const T alS = __shfl_down_sync(0xffffffffu, aS, o), blS = __shfl_down_sync(0xffffffffu, bS, o);
                    const T plP = __shfl_up_sync(0xffffffffu, pP, o);
                    if (lane >= o) {
                        bU = fma(aU, blU, bU); aU = aU * alU;
                        pP += plP;
                    }
                    if (lane + o < 32) { bS = fma(aS, blS, bS); aS = aS * alS; }
                }
                if (lane == 31) { s_scan[0][warp] = aU; s_scan[1][warp] = bU; s_scan[4][warp] = pP; }
                if (lane == 0) { s_scan[2][warp] = aS; s_scan[3][warp] = bS; s_wcnt[warp] = __popc(bal); }
                __syncthreads();
                // every warp combines the per-warp totals itself (lane w holds warp w's): no second barrier, nobody idles
                {
                    const bool in = lane < nwarp;
                    T a1 = in ? s_scan[0][lane] : T(1.0), b1 = in ? s_scan[1][lane] : T(0.0);
                    T a2 = in ? s_scan[2][lane] : T(1.0), b2 = in ? s_scan[3][lane] : T(0.0);
                    T p3 = in ? s_scan[4][lane] : T(0.0);
                    int lc = in ? s_wcnt[lane] : 0;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const T al1 = __shfl_up_sync(0xffffffffu, a1, o), bl1 = __shfl_up_sync(0xffffffffu, b1, o);
                        const T al2 = __shfl_down_sync(0xffffffffu, a2, o), bl2 = __shfl_down_sync(0xffffffffu, b2, o);
                        const T pl3 = __shfl_up_sync(0xffffffffu, p3, o);
                        const int lcl = __shfl_up_sync(0xffffffffu, lc, o);
                        if (lane >= o) {
                            b1 = fma(a1, bl1, b1); a1 = a1 * al1;
                            p3 += pl3;
                            lc += lcl;
                        }
                        if (lane + o < 32) { b2 = fma(a2, bl2, b2); a2 = a2 * al2; }
                    }
                    // carried in: U, P and the loud count from the warps before, S from the warps after
                    const int wb = warp > 0 ? warp - 1 : 0, wa = warp + 1 < 32 ? warp + 1 : 31;
                    const T cU = __shfl_sync(0xffffffffu, b1, wb), cP = __shfl_sync(0xffffffffu, p3, wb);
                    const int cL = __shfl_sync(0xffffffffu, lc, wb);
                    const T cS = __shfl_sync(0xffffffffu, b2, wa);
                    if (warp > 0) { bU = fma(aU, cU, bU); pP += cP; }
                    if (warp + 1 < nwarp) bS = fma(aS, cS, bS);
                    const int lpos = (warp > 0 ? cL : 0) + __popc(bal & ((1u << lane) - 1u));
                    if (i <= npk) sm.lcnt[i] = lpos;     // entry npk = total (threads >= npk are not loud)
                    if (loud) sm.lidx[lpos] = (uint16_t)i;
                }
                if (i < npk) { sm.mU[i] = bU; sm.mS[i] = bS; sm.mP[i + 1] = pP; }   // P: inclusive sum up to i = sum below i+1
                if (tid == 0) { sm.mS[npk] = T(0.0); sm.mP[0] = T(0.0); }
                if (sm.zlut != nullptr) {
                    // count table over Bark cells for masker_range (the peak bins in these words were last read before
                    // the first barrier above): zlut[g] = number of maskers with z < g/32 Bark.  The maskers ascend in z, so
                    // the count is a step function of the cell: masker i raises it to i+1 from its own cell up to the next
                    // masker's -- every thread fills its stretch, no counting and no prefix sum.
                    auto cell_of = [&](T zz) {
                        int cell = (int)(zz * T(32.0)) + 1;               // counted from cell+1 on: z < g/32 for g > z*32
                        return cell < 1 ? 1 : (cell > MRC_ZLUT ? MRC_ZLUT : cell);
                    };
                    if (tid < npk) {
                        const int c0 = cell_of(z), c1 = (tid + 1 < npk) ? cell_of(sm.mz[tid + 1]) : MRC_ZLUT + 1;
                        for (int g = c0; g < c1; ++g) sm.zlut[g] = (uint16_t)(tid + 1);
                        if (tid == 0) for (int g = 0; g < c0; ++g) sm.zlut[g] = 0;
                    }
                    if (npk == 0) for (int g = tid; g <= MRC_ZLUT; g += NT) sm.zlut[g] = 0;
                }
            }
        }
        if (tid < MRC_BSTRIDE) s_best[tid] = 0u;
        __syncthreads();
        MRC_CLK(5);
        // e. masked threshold at the MDCT lines, f. SMR per line (psychoac.py:212-214), band maxima (:215-219)
        const T sc6 = T(6) * T(s_scale[c]);
        auto line_spl = [&](int k) -> T {           // SPL of the (scaled) MDCT line, scale undone
            const T X = sm.lines[c * L + k];
            return fmax(T(96) + T(10) * m_log10((T(2) * (X * X)) / T(T(0.5))), T(-30)) - sc6;
        };
        auto thr_of = [&](T a) -> T { return fmax(T(96) + T(10) * m_log10(T(a)), T(-30)); };
        if (cp.spread_seq) {
            // reference order (psychoac.py:68-78, :168): every masker onto every line, one 10**x per pair
            const int k0 = tid, k1 = tid + Q;
            const T z0 = tb.bark[k0], z1 = tb.bark[k1];
            T a0 = tb.quiet[k0], a1 = tb.quiet[k1];
            for (int m = 0; m < npk; ++m) {
                const T zm = sm.mz[m], s15 = sm.ms15[m], gg = sm.mg[m];
                a0 += masker_intensity(z0 - zm, s15, gg);
                a1 += masker_intensity(z1 - zm, s15, gg);
            }
            sm.xi[k0] = line_spl(k0) - thr_of(a0);
            sm.xi[k1] = line_spl(k1) - thr_of(a1);
            __syncthreads();
            for (int bd = warp; bd < nb; bd += nwarp) {
                const int lo = tb.c_band_lo[bd], n = tb.c_band_n[bd];
                T v = -INFINITY;
                for (int i = lane; i < n; i += 32) v = fmax(v, sm.xi[lo + i]);
                v = warp_max(v);
                if (lane == 0) s_smr[c][bd] = v;
            }
        } else {
            // The band SMR is a maximum over lines, so only lines that can be the maximum need their complete
            // threshold.  With both clamps of psychoac.py:12 folded in, a line's SMR is 10 log10(rho) - 6 scale,
            // rho = max(4 X^2, floor) / max(threshold intensity, floor): monotone in rho.
            //   pass 1 (thread per line): rho from a LOWER bound of the threshold = an upper bound of the line's rho;
            //   pass 2 (warp per band): complete threshold (warp-cooperative) of the band's line with the highest
            //           bound -> its true rho and SMR; then every other line of the band whose bound reaches the best
            //           true rho seen so far (less 1e-9) gets its complete threshold too; lines whose bound stays
            //           below cannot be the maximum.
            // The band SMR is the maximum of the completely evaluated lines' SMRs: exact.
            const T FLOOR = T(2.5118864315095823e-13);        // 10^((-30-96)/10): where SPL() clamps
            auto x2c = [&](int k) -> T {
                const T X = sm.lines[c * L + k];
                return fmax((T(2.0) * (X * X)) / T(0.5), FLOOR);
            };
            {
                // lines of bands that do not select this spectrum are never candidates (ub = -1)
                auto bound_of = [&](int k) {
                    const int bd = tb.line2band[k];
                    float ub = -1.0f;
                    uint32_t rng = 0;
                    if ((need >> bd) & 1u) {
                        int m_lo, m_hi;
                        const T r = T(x2c(k) / fmax(spread_line_bound(sm, tb, k, npk, n_general, m_lo, m_hi), FLOOR));
                        if constexpr (sizeof(T) == 8) ub = __double2float_ru(r); else ub = r;
                        rng = (uint32_t)m_lo | ((uint32_t)m_hi << 16);
                    }
                    // the band's first candidate: (about) the highest bound, by one shared atomic per band and warp.  ANY
                    // line may go first -- pass 2 evaluates every line whose bound reaches the best true value -- so the
                    // low mantissa bits make room for the line index.
                    const unsigned key = ub > 0.0f ? ((__float_as_uint(ub) & ~2047u) | (unsigned)(2047 - k)) : 0u;
                    // a warp's 32 consecutive lines mostly lie in one band: one reduction and one atomic; otherwise (the
                    // narrow bands at the bottom, a warp across a band edge) every lane posts its own
                    if (__all_sync(0xffffffffu, bd == __shfl_sync(0xffffffffu, bd, 0))) {
                        const unsigned top = __reduce_max_sync(0xffffffffu, key);
                        if (lane == 0 && top) atomicMax(&s_best[bd], top);
                    } else if (key) atomicMax(&s_best[bd], key);
                    li_store<T>(sm, k, ub, rng);
                };
                static_assert(L <= 2048, "line index in 11 bits");
                bound_of(tid);
                bound_of(tid + Q);
            }
            __syncthreads();
            MRC_CLK(6);
            const T slack = sizeof(T) == 8 ? T(T(1.0) - 1e-9) : T(T(1.0) - 1e-4);   // bound vs true value: rounding only
            Grp grp;
            grp.shift = lane & ~(GW - 1);
            grp.gl = lane - grp.shift;
            grp.mask = ((GW == 32) ? 0xffffffffu : ((1u << (GW & 31)) - 1u)) << grp.shift;
            auto complete = [&](int k, T& smr, T& rho) {         // whole group; all its lanes get the results
                const T a = spread_line_grp(sm, tb, k, npk, grp, n_general, n_window);
                MRC_WCLK_BEGIN();
                // line_spl(k) - thr_of(a) with the two logarithms side by side (odd lanes the line, even lanes the
                // threshold): same operations on the same operands, half the latency
                const T X = sm.lines[c * L + k];
                const T lg = m_log10((lane & 1) ? (T(2) * (X * X)) / T(T(0.5)) : T(a));
                const T lg_a = __shfl_sync(grp.mask, lg, 0, GW), lg_x = __shfl_sync(grp.mask, lg, 1, GW);
                smr = (fmax(T(96) + T(10) * lg_x, T(-30)) - sc6) - fmax(T(96) + T(10) * lg_a, T(-30));
                rho = T(x2c(k) / fmax(a, FLOOR));
                MRC_WCLK(21);
#ifdef MRC_PHASE_CLOCKS
                if (lane == 0) atomicAdd(&g_phase_clk[22], 1ull);
#endif
            };
            // pass 2, one GROUP of lanes per BAND that selects this spectrum (32 groups: one round of the CTA covers
            // every band; bands are dealt out from the top: their lines see the most loud maskers and take longest):
            // the band's line with the highest bound gets its complete threshold; its true rho is the bar every other
            // line of the band has to reach with its bound to be evaluated as well (rare: the bounds are tight, about
            // one extra line per block).  The two groups of a warp run their bands in lock step where they can and
            // diverge where they must (all their synchronising operations name only the group's lanes).
            {
                const unsigned bl = need & band_mask;
                const int nbl = __popc(bl);
                constexpr int GPW = 32 / GW;                 // groups per warp
                for (int p = warp * GPW + (lane / GW); p < nbl; p += nwarp * GPW) {
                    MRC_WCLK_BEGIN();
#ifdef MRC_PHASE_CLOCKS
                    const long long tb0_ = clock64();
                    unsigned ncomp_ = 0;
#endif
                    const int bd = 31 - (int)__fns(__brev(bl), 0, p + 1);     // p-th needed band from the top
                    const int lo = tb.c_band_lo[bd], n = tb.c_band_n[bd];
                    const unsigned key = s_best[bd];
                    const int kbest = key ? 2047 - (int)(key & 2047u) : lo;
                    T best, rbest;
                    MRC_WCLK(24);
                    complete(kbest, best, rbest);
                    MRC_WCLK(25);
                    for (int base = 0; base < n; base += GW) {
                        const int i = base + grp.gl;
                        const T ub = (i < n) ? T(li_ub<T>(sm, lo + i)) : T(-1);
                        unsigned bal = grp_ballot(grp, i < n && lo + i != kbest && ub >= rbest * slack);
                        while (bal) {
                            const int l = __ffs(bal) - 1;
                            bal &= bal - 1;
                            const T ubl = __shfl_sync(grp.mask, ub, l, GW);
                            if (ubl >= rbest * slack) {          // rbest may have risen since the ballot
                                T smr, rho;
                                complete(lo + base + l, smr, rho);
                                best = fmax(best, smr);
                                rbest = fmax(rbest, rho);
                            }
                        }
                    }
                    if (grp.gl == 0) s_band_smr[bd] = best;
                    MRC_WCLK(26);
#ifdef MRC_PHASE_CLOCKS
                    if (lane == 0) {
                        atomicAdd(&g_band_clk[0][bd], (unsigned long long)(clock64() - tb0_));
                        atomicAdd(&g_band_clk[1][bd], 1ull);
                    }
#endif
                }
            }
#ifdef MRC_PHASE_CLOCKS
            {
                MRC_WCLK_BEGIN();
                __syncthreads();
                MRC_WCLK(27);                        // warp 0's wait for the slowest group
            }
#endif
            __syncthreads();
            MRC_CLK(7);
            MRC_CLK(8);
            if (tid < nb) {
                T v = T(0);                              // bands that do not select this spectrum: value never used
                if ((need >> tid) & 1u) v = s_band_smr[tid];
                s_smr[c][tid] = v;
            }
        }
        __syncthreads();
        MRC_CLK(9);
        if (taps.npeaks != nullptr && tid == 0) taps.npeaks[lb * 4 + c] = npk;
    }
    // The block's samples are no longer needed (every barrier of the loop above lies behind): the last warp's first lane
    // places the CTA's next block and asks the copy engine for its frames.
    if (tid == NT - 32) {
        int st = 0;
        if (!XIN && it_next >= 0) {
            const int gn = g0 + (cm.list ? cm.list[it_next] : it_next);
            int lo = 0, hi = cm.n_clips;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (cm.clip_blk0[mid] <= gn) lo = mid; else hi = mid;
            }
            const int bn = gn - cm.clip_blk0[lo];
            nxt->clip = lo; nxt->b = bn; nxt->nblk_clip = cm.clip_blk0[lo + 1] - cm.clip_blk0[lo];
            const long long frames = cm.clip_off[lo + 1] - cm.clip_off[lo];
            const long long s0 = cm.blk_start ? cm.blk_start[gn] - tb.a : (long long)(bn + cm.blk_base - 1) * L;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(pcm) + (cm.clip_off[lo] - cm.pcm_frame0) + s0;
            st = 2;
            if (s0 >= 0 && s0 + N <= frames && (reinterpret_cast<size_t>(src) & 15) == 0) {
                fence_proxy_async();          // the CTA's reads and writes of this buffer come first
                mbar_expect_tx(mbar, (unsigned)(N * 4));
                bulk_g2s(sm.px, src, (unsigned)(N * 4), mbar);
                st = 1;
            }
        }
        nxt->state = st;
    }
    if (tid == 0) {
        if (peak_counter) atomicAdd(peak_counter, (unsigned long long)my_peaks);
    }
    if (peak_counter) {     // executed-work counters: [1] general (10**x) pairs, [2] plateau adds, [3] loud maskers
        const unsigned g1 = __reduce_add_sync(0xffffffffu, n_general), g2 = __reduce_add_sync(0xffffffffu, n_window),
                       g3 = __reduce_add_sync(0xffffffffu, n_loud);
        if (lane == 0) {
            atomicAdd(peak_counter + 1, (unsigned long long)g1);
            atomicAdd(peak_counter + 2, (unsigned long long)g2);
            atomicAdd(peak_counter + 3, (unsigned long long)g3);
        }
    }
    if (tid == 0) {
        if (taps.npeaks != nullptr) for (int c = nspec; c < 4; ++c) taps.npeaks[lb * 4 + c] = 0;
    }
    if (taps.smr4 != nullptr) {
        for (int i = tid; i < 4 * MRC_BSTRIDE; i += NT) {
            const int c = i / MRC_BSTRIDE, bd = i % MRC_BSTRIDE;
            taps.smr4[(size_t)lb * 4 * MRC_BSTRIDE + i] = (c < nspec && bd < nb) ? s_smr[c][bd] : T(0);
        }
    }

    // ---- phase 5: per band pick (M,S) or (L,R) (ms_stereo.py:70-81; codecThem.py:509-559) ---------------
    const unsigned ms = s_ms;
    {
        T* oA = ho.lines + (size_t)lb * 2 * cp.Lmax;     // [2][L] compact at a stride of 2*Lmax per block
        T* oB = oA + L;
        for (int k = tid; k < L; k += NT) {
            const bool m = (ms >> tb.line2band[k]) & 1u;
            oA[k] = sm.lines[(m ? 2 : 0) * L + k];
            oB[k] = sm.lines[(m ? 3 : 1) * L + k];
        }
        for (int w2 = warp; w2 < 2 * nb; w2 += nwarp) {
            const int ch = w2 / nb, bd = w2 - ch * nb;
            const bool m = (ms >> bd) & 1u;
            const T* src = sm.lines + ((m ? 2 : 0) + ch) * L + tb.c_band_lo[bd];
            const int n = tb.c_band_n[bd];
            T v = 0;
            for (int i = lane; i < n; i += 32) v = fmax(v, fabs(src[i]));
            v = warp_max(v);
            if (lane == 0) {
                ho.bandmax[(size_t)lb * 2 * MRC_BSTRIDE + ch * MRC_BSTRIDE + bd] = v;
                if (ho.smr != nullptr)
                    ho.smr[(size_t)lb * 2 * MRC_BSTRIDE + ch * MRC_BSTRIDE + bd] = s_smr[(m ? 2 : 0) + ch][bd];
            }
        }
        if (tid < 4) ho.ovs[lb * 4 + tid] = (uint8_t)s_scale[tid];
        if (tid == 0) ho.ms[lb] = ms;
    }

    // ---- phase 6: order of the water-filling grants (bitalloc.py:131-149) --------------------------------
    // Each band's SMR trajectory (smr, -12, -6, -6, ...) is independent of the other bands, so the greedy
    // arg-max visits the (band, level) tokens in globally sorted order: key descending, first (lowest) band on
    // ties.  Joint blocks sort 2*nb bands together; non-joint blocks sort each channel on its own.
    // The 2*nb bands give 2*nb runs of 15 keys that are already in order, so the order is a merge: runs padded to 64
    // runs of 16, then 6 rounds of pairwise merging in which every element finds its place by a binary search in
    // the partner run (stable: ties in key go to the lower band, then to the run that came first).  Non-joint
    // blocks keep their two channels in separate halves and skip the last round.  All comparisons are on exactly
    // the values the reference's `smr[i] -= 12.0 / 6.0` updates produce.
    // Fast path: every key is s_b - 6m with m in {0, 2, 3, ..., 15}, so with s_b = 6 q_b + f_b (q_b integer, 0 <= f_b < 6)
    // a token sits on the integer level v = q_b - m and the order is: level descending, then f_b descending, then band
    // ascending.  A token's position = tokens on higher levels + bands of its own level that precede its band -- counted
    // with one 64-bit mask of band ranks per level: no sorting, no searches.  This is the order of the keys in real
    // arithmetic; the reference's keys carry the rounding of its repeated subtractions (< 1e-12), so the two orders can
    // only differ where keys of two bands come closer than that.  In fp64 mode the fast path is therefore used only when
    // all f_b are pairwise at least 1e-9 apart (also across the wrap 0 ~ 6) or belong to bitwise equal SMRs (equal
    // channels: identical key sequences, ordered by band); otherwise the merge below runs (about once in 10^6 blocks).
    __shared__ int s_unsafe;
    {
        __syncthreads();                                 // phase 5 is done reading sm.lines: the buffers below alias it
        MRC_CLK(10);
        constexpr int NLEV = 128, LEV_OFF = 64;
        unsigned long long* const lmask = reinterpret_cast<unsigned long long*>(sm.mkey);      // [2][NLEV]
        int* const lstart = reinterpret_cast<int*>(lmask + 2 * NLEV);                           // [2][NLEV]
        double* const bf = reinterpret_cast<double*>(lstart + 2 * NLEV);                        // [64] f_b
        double* const bs = bf + 64;                                                             // [64] s_b
        int* const bq = reinterpret_cast<int*>(bs + 64);                                        // [64] q_b
        int* const brank = bq + 64;                                                             // [64]
        const int ntot = 2 * nb, per_group = nb * MRC_MAX_LEVELS;
        const double EPS = 1e-9;
        uint16_t* const tokbuf = reinterpret_cast<uint16_t*>(brank + 64);                       // [MRC_TOK_STRIDE]
        if (tid == 0) s_unsafe = 0;
        if (tid < ntot) {
            const int ch = tid / nb, bd = tid - ch * nb;
            const bool m = (ms >> bd) & 1u;
            const double sv = (double)s_smr[(m ? 2 : 0) + ch][bd];
            double qd = floor(sv * (1.0 / 6.0));
            double f = fma(-6.0, qd, sv);
            if (f < 0.0) { qd -= 1.0; f += 6.0; }
            if (f >= 6.0) { qd += 1.0; f -= 6.0; }
            bool bad = !(qd > -48.0 && qd < 62.0) || !(f >= 0.0 && f < 6.0);       // also NaN / infinities
            if (sizeof(T) == 8 && !(f >= EPS && f <= 6.0 - EPS)) bad = true;
            bf[tid] = f; bs[tid] = sv; bq[tid] = bad ? 0 : (int)qd;
            if (bad) s_unsafe = 1;
        }
        __syncthreads();
        // rank of every band among the bands of its group by (f descending, band ascending): one warp per band, the
        // lanes hold the other bands (two each)
        for (int bb = warp; bb < ntot; bb += nwarp) {
            const int grp = joint ? 0 : bb / nb;
            const int b0 = joint ? 0 : grp * nb, b1 = joint ? ntot : b0 + nb;
            const double f = bf[bb], sv = bs[bb];
            int r = 0;
            bool bad = false;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = b0 + lane + 32 * h;
                const bool in = j < b1;
                const double fj = in ? bf[j] : 0.0;
                r += __popc(__ballot_sync(0xffffffffu, in && (fj > f || (fj == f && j < bb))));
                if (sizeof(T) == 8 && in && j != bb && fabs(fj - f) < EPS && !(bs[j] == sv)) bad = true;
            }
            if (lane == 0) brank[bb] = r;
            if (bad) s_unsafe = 1;
        }
        __syncthreads();
        // which bands hold a token on level v: one thread per (group, level) gathers the ranks -- no atomics
        for (int i = tid; i < 2 * NLEV; i += NT) {
            const int grp = i / NLEV, v = i - grp * NLEV - LEV_OFF;
            unsigned long long mk = 0ull;
            const int b0 = joint ? 0 : grp * nb, b1 = joint ? (grp == 0 ? ntot : 0) : b0 + nb;
            for (int j = b0; j < b1; ++j) {
                const int m = bq[j] - v;                 // token of band j on this level: m = 0 or 2 <= m <= 15
                if (m == 0 || (m >= 2 && m <= MRC_MAX_LEVELS)) mk |= 1ull << brank[j];
            }
            lmask[i] = mk;
        }
        __syncthreads();
        if (warp < 2) {                                  // tokens on higher levels, per group: suffix sums over the levels
            int c[4], tot = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) { c[i] = __popcll(lmask[warp * NLEV + lane * 4 + i]); tot += c[i]; }
            int suf = tot;                               // inclusive suffix sum over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_down_sync(0xffffffffu, suf, o);
                if (lane + o < 32) suf += v;
            }
            int above = suf - tot;                       // tokens on the levels of higher lanes
#pragma unroll
            for (int i = 3; i >= 0; --i) { lstart[warp * NLEV + lane * 4 + i] = above; above += c[i]; }
        }
        for (int e = tid; e < MRC_TOK_STRIDE; e += NT) tokbuf[e] = 0xffffu;       // slots past the tokens
        __syncthreads();
        uint16_t* ot = ho.tokens + (size_t)lb * MRC_TOK_STRIDE;
        if (!s_unsafe) {
            for (int e = tid; e < ntot * MRC_MAX_LEVELS; e += NT) {
                const int bb = e / MRC_MAX_LEVELS, l = e - bb * MRC_MAX_LEVELS;
                const int grp = joint ? 0 : bb / nb;
                const int li = grp * NLEV + (bq[bb] - (l ? l + 1 : 0) + LEV_OFF);
                const int pos = lstart[li] + __popcll(lmask[li] & ((1ull << brank[bb]) - 1ull));
                MRC_ASSERT(li >= 0 && li < 2 * NLEV && pos >= 0 && pos < per_group * (joint ? 2 : 1));
                tokbuf[grp * per_group + pos] = (uint16_t)(bb | (l << 8));
            }
            __syncthreads();
            for (int e = tid; e < MRC_TOK_STRIDE / 2; e += NT)
                reinterpret_cast<uint32_t*>(ot)[e] = reinterpret_cast<const uint32_t*>(tokbuf)[e];
        }
        __syncthreads();                                 // (the merge buffers alias the arrays above)
    }
    if (s_unsafe) {
        MRC_CLK(10);
        T* const key0 = sm.mkey;
        T* const key1 = sm.mkey + 1024;
        uint16_t* const id0 = sm.mid;
        uint16_t* const id1 = sm.mid + 1024;
        for (int e = tid; e < 1024; e += NT) {
            const int run = e >> 4, lvl = e & 15;
            // joint: run = band of the 2nb-band list; non-joint: channel 0 in runs 0..31, channel 1 in runs 32..63
            int bb = -1;
            if (joint) { if (run < 2 * nb) bb = run; }
            else { const int ch = run >> 5, bd = run & 31; if (bd < nb) bb = ch * nb + bd; }
            T v = -INFINITY;
            uint16_t id = 0xffffu;
            if (bb >= 0 && lvl < MRC_MAX_LEVELS) {
                const int ch = bb / nb, bd = bb - ch * nb;
                const bool m = (ms >> bd) & 1u;
                v = s_smr[(m ? 2 : 0) + ch][bd];
                if (lvl >= 1) v -= T(12);
                for (int q = 1; q < lvl; ++q) v -= T(6);
                id = (uint16_t)((bb << 4) | lvl);        // band-major: ties in key go to the lower band
            }
            key0[e] = v;
            id0[e] = id;
        }
        __syncthreads();
        MRC_CLK(11);
        int cur = 0;
        const int last_w = joint ? 512 : 256;
        for (int w = 16; w <= last_w; w <<= 1) {
            const T* ks = cur ? key1 : key0;
            const uint16_t* is = cur ? id1 : id0;
            T* kd = cur ? key0 : key1;
            uint16_t* idd = cur ? id0 : id1;
            // number of partner elements that go before an element: a fixed-step search (the predicate holds for a
            // prefix of the sorted partner run), two elements per thread in lock step so that their dependent loads
            // overlap
            auto before = [&](int idx, T v, uint16_t id, bool first) -> bool {
                const T x = ks[idx];
                bool b = x > v;
                if (x == v) {                            // rare: only ties look at the ids
                    const uint16_t xi = is[idx];
                    b = xi < id || (xi == id && !first);
                }
                return b;
            };
            const int logw = 31 - __clz(w);
            for (int e0 = tid; e0 < 1024; e0 += 2 * NT) {
                const int e1 = e0 + NT;
                const bool two = e1 < 1024;
                const int q0 = e0 >> logw, i0 = e0 - (q0 << logw), q1 = two ? (e1 >> logw) : q0, i1 = two ? e1 - (q1 << logw) : i0;
                const T v0 = ks[e0], v1 = ks[two ? e1 : e0];
                const uint16_t d0 = is[e0], d1 = is[two ? e1 : e0];
                const int b0 = (q0 ^ 1) << logw, b1 = (q1 ^ 1) << logw;          // partner runs
                const bool f0 = (q0 & 1) == 0, f1 = (q1 & 1) == 0;              // elements of the first run win ties
                int p0 = 0, p1 = 0;
                for (int step = w >> 1; step >= 1; step >>= 1) {
                    if (before(b0 + p0 + step - 1, v0, d0, f0)) p0 += step;
                    if (before(b1 + p1 + step - 1, v1, d1, f1)) p1 += step;
                }
                if (before(b0 + p0, v0, d0, f0)) ++p0;
                if (before(b1 + p1, v1, d1, f1)) ++p1;
                const int dst0 = ((q0 >> 1) << (logw + 1)) + i0 + p0;
                kd[dst0] = v0;
                idd[dst0] = d0;
                if (two) {
                    const int dst1 = ((q1 >> 1) << (logw + 1)) + i1 + p1;
                    kd[dst1] = v1;
                    idd[dst1] = d1;
                }
            }
            __syncthreads();
            cur ^= 1;
        }
        const uint16_t* fin = cur ? id1 : id0;
        uint16_t* ot = ho.tokens + (size_t)lb * MRC_TOK_STRIDE;
        const int per_group = nb * MRC_MAX_LEVELS;
        for (int i = tid; i < MRC_TOK_STRIDE; i += NT) {
            int src = -1;
            if (joint) { if (i < 2 * per_group) src = i; }
            else if (i < per_group) src = i;
            else if (i < 2 * per_group) src = 512 + (i - per_group);
            uint16_t o = 0xffffu;
            if (src >= 0) { const uint16_t id = fin[src]; o = (uint16_t)((id >> 4) | ((id & 15) << 8)); }
            ot[i] = o;
        }
    }
    MRC_CLK(12);
}

// A CTA takes MRC_BLOCKS_PER_CTA blocks (blockIdx.x, blockIdx.x + gridDim.x, ...: consecutive blocks go to different CTAs,
// so the cheap blocks of a silent passage are spread evenly): the tables that do not depend on the block -- 2^(j/64), the
// Bark positions of the lines, the FFT stage twiddles -- are staged once per CTA, and every block but the first finds its
// frames already in shared memory (one cp.async.bulk issued while the block before is finished).  Measured on the 1 h
// stream (scripts/gpu_bpc.sh, audio-s/s fp64 / fp32): 1 block per CTA 56.4k / 87.0k, 2: 56.6k / 86.9k, 4: 56.2k / 85.9k,
// 16: 54.0k, a grid of fully persistent CTAs 53.6k / 82.5k -- the staging it saves is small, and the kernels of the
// neighbouring waves (cost, pack, reservoir maps, on streams of their own) get their SMs when analysis CTAs retire: long-
// lived CTAs hold them back (pack 6.3 -> 13.1 ms per hour with persistent CTAs).  So: two.  The per-block seam (XIN)
// launches one CTA per block.
#ifndef MRC_BLOCKS_PER_CTA
#define MRC_BLOCKS_PER_CTA 2
#endif
template <typename T, int L_, bool XIN>
__global__ void __launch_bounds__(L_ / 2, (L_ <= 1024) ? (sizeof(T) == 4 ? 3 : 2) : 1)
analysis_kernel(DevTables<T> tb, CodecParams cp, ClipMap cm, const int16_t* __restrict__ pcm,
                const double* __restrict__ xin, int g0, int nblk, Handoff<T> ho, AnalysisTaps<T> taps,
                unsigned long long* peak_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int L = L_, NT = L / 2;
    const int tid = threadIdx.x;
    Smem<T> sm = carve<T, XIN>(smem_raw, L);
    __shared__ __align__(8) unsigned long long s_mbar;
    __shared__ NextBlock s_next;
    if (tid < 64) sm.etab[tid] = tb.exp_tab[tid];
    for (int i = tid; i < L; i += NT) sm.zb[i] = tb.bark[i];
    if constexpr (FftShape<L>::pow2) {
        cpx<T>* dst = const_cast<cpx<T>*>(sm.stL);
        for (int i = tid; i < stage_entries_of(L); i += NT) dst[i] = tb.tw_stage[i];
    }
    if (tid == 0) {
        s_next.state = 0;
        mbar_init(&s_mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    unsigned parity = 0;
    for (int it = blockIdx.x; it < nblk; it += gridDim.x) {
        const int nx = it + (int)gridDim.x;
        analysis_block<T, L_, XIN>(sm, tb, cp, cm, pcm, xin, g0, it, nx < nblk ? nx : -1, ho, taps, peak_counter, &s_mbar,
                                   &s_next, parity);
        __syncthreads();                     // the next block rewrites what this one's last phase still reads
    }
}

}  // namespace

#ifdef MRC_PHASE_CLOCKS
extern "C" int mrc_debug_band_clocks(unsigned long long* out96) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out96, g_band_clk, sizeof g_band_clk) == cudaSuccess ? 0 : -1;
}
extern "C" int mrc_debug_phase_clocks(unsigned long long* out32, int reset) {
    cudaDeviceSynchronize();
    if (out32 && cudaMemcpyFromSymbol(out32, g_phase_clk, sizeof g_phase_clk) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[32] = {0};
        if (cudaMemcpyToSymbol(g_phase_clk, z, sizeof z) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

static int blocks_per_cta() {          // MRC_BLOCKS_PER_CTA in the environment overrides the built-in value (tuning)
    static int n = 0;
    if (n == 0) {
        const char* e = getenv("MRC_BLOCKS_PER_CTA");
        n = e ? atoi(e) : MRC_BLOCKS_PER_CTA;
        if (n < 1) n = 1;
    }
    return n;
}

size_t analysis_smem_bytes(int L, int elem, bool xin) {
    const size_t Q = L / 2;
    size_t merge = (size_t)2048 * elem + 2048 * 2;                     // merge buffers of the grant order
    if ((size_t)(4 * L) * elem >= merge) merge = 0;                    // ... living in `lines`
    const bool lin = elem == 4 && !(L & (L - 1));                      // fast mode: four intensity arrays over the samples
    const size_t samples = (xin || lin) ? (size_t)(4 * L) * elem : (size_t)8 * L;   // values of the seam, or packed PCM frames
    const size_t stage = (size_t)stage_entries_of(L) * 2 * elem + (elem == 8 ? 0 : (size_t)L * 4);
    return samples + (size_t)(8 * L) * elem + 3 * Q * elem + stage + 8 + 64 * 8 + (2 * Q + 1) * 4 + Q * 2 + 32 + merge;
}

template <typename T>
void launch_analysis(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const ClipMap& cm,
                     const int16_t* pcm, const double* xin, int g0, int nblk, Handoff<T> ho, AnalysisTaps<T> taps,
                     unsigned long long* peak_counter) {
    if (nblk <= 0) return;
    const bool x = xin != nullptr;
    const size_t smem = analysis_smem_bytes(tb.L, sizeof(T), x);
#define MRC_LAUNCH_ANALYSIS_X(LL, XX)                                                                            \
    {                                                                                                            \
        cudaFuncSetAttribute(analysis_kernel<T, LL, XX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        const int grid = XX ? nblk : (nblk + blocks_per_cta() - 1) / blocks_per_cta();                            \
        analysis_kernel<T, LL, XX><<<grid, LL / 2, smem, st>>>(tb, cp, cm, pcm, xin, g0, nblk, ho, taps,         \
                                                               peak_counter);                                    \
    }
#define MRC_LAUNCH_ANALYSIS(LL)                                                                                  \
    case LL:                                                                                                     \
        if (x) MRC_LAUNCH_ANALYSIS_X(LL, true) else MRC_LAUNCH_ANALYSIS_X(LL, false)                             \
        break;
    switch (tb.L) {
        MRC_LAUNCH_ANALYSIS(128)       // short blocks (128 + 128)
        MRC_LAUNCH_ANALYSIS(256)
        MRC_LAUNCH_ANALYSIS(512)
        MRC_LAUNCH_ANALYSIS(576)       // transition blocks (1024 + 128, 128 + 1024)
        MRC_LAUNCH_ANALYSIS(1024)
        MRC_LAUNCH_ANALYSIS(2048)
        default: break;
    }
#undef MRC_LAUNCH_ANALYSIS
#undef MRC_LAUNCH_ANALYSIS_X
}

template void launch_analysis<double>(cudaStream_t, const DevTables<double>&, const CodecParams&, const ClipMap&,
                                      const int16_t*, const double*, int, int, Handoff<double>,
                                      AnalysisTaps<double>, unsigned long long*);
template void launch_analysis<float>(cudaStream_t, const DevTables<float>&, const CodecParams&, const ClipMap&,
                                     const int16_t*, const double*, int, int, Handoff<float>, AnalysisTaps<float>,
                                     unsigned long long*);
