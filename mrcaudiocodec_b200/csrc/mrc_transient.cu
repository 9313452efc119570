// mrc_transient.cu -- block switching, step 1: the transient detector of the reference's encode loop
// (SURVEY.md 8 f1).
//   TransientDetector            pacfileThem.py:1021-1056
//   filter                       :1146-1147  20th-order Chebyshev-II high-pass as second-order sections; the host
//                                designs it with the reference's own scipy calls and hands the sections over
//   scipy.signal.sosfilt         direct form II transposed, zero state at the start of EVERY nMDCTLines-sample
//                                block (the reference calls it per block without zi), products and sums unfused
//   decision                     :1192  sum(blkswMem) > 1 or any(blksw == 1)
// Every block is filtered on its own, so the detector is data parallel: one thread per (block, channel) runs the
// 10-section cascade over the block's samples (state in registers, coefficients in the constant bank) and keeps the
// peak of each 128-sample segment; a second kernel compares neighbouring peaks (the previous block's last peak is the
// only thing that crosses a block boundary, :1053) and writes two flags per block.
#include "mrc_internal.cuh"
#include "mrc_math.cuh"

namespace {

template <int NSEC>
__global__ void __launch_bounds__(128)
peaks_kernel(SosParams sp, const int64_t* __restrict__ clip_off, const int32_t* __restrict__ clip_sb0, int n_clips,
             const int16_t* __restrict__ pcm, int L, int nsb_total, double* __restrict__ peaks) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int sb = t >> 1, ch = t & 1;
    if (sb >= nsb_total) return;
    int lo = 0, hi = n_clips;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (clip_sb0[mid] <= sb) lo = mid; else hi = mid;
    }
    const long long frames = clip_off[lo + 1] - clip_off[lo];
    const long long s0 = (long long)(sb - clip_sb0[lo]) * L;
    const int16_t* __restrict__ src = pcm + (clip_off[lo] + s0) * 2 + ch;
    const int nvalid = (int)min((long long)L, frames - s0);      // the last block of a clip is zero padded
    const int nsec = NSEC ? NSEC : sp.n;
    double z0[NSEC ? NSEC : MRC_MAX_SOS], z1[NSEC ? NSEC : MRC_MAX_SOS];
#pragma unroll
    for (int s = 0; s < (NSEC ? NSEC : MRC_MAX_SOS); ++s) { z0[s] = 0.0; z1[s] = 0.0; }
    const int nseg = L / MRC_SHORT;
    double* out = peaks + ((size_t)sb * 2 + ch) * nseg;
    for (int seg = 0; seg < nseg; ++seg) {
        double pk = 0.0;
        for (int i = 0; i < MRC_SHORT; ++i) {
            const int n = seg * MRC_SHORT + i;
            double xc = (n < nvalid) ? pcm_to_fraction<double>((int)__ldg(src + 2 * n)) : 0.0;
#pragma unroll
            for (int s = 0; s < (NSEC ? NSEC : MRC_MAX_SOS); ++s) {
                if (!NSEC && s >= nsec) break;
                const double xn = __dadd_rn(__dmul_rn(sp.c[s][0], xc), z0[s]);
                z0[s] = __dadd_rn(__dsub_rn(__dmul_rn(sp.c[s][1], xc), __dmul_rn(sp.c[s][3], xn)), z1[s]);
                z1[s] = __dsub_rn(__dmul_rn(sp.c[s][2], xc), __dmul_rn(sp.c[s][4], xn));
                xc = xn;
            }
            pk = fmax(pk, fabs(xc));
        }
        out[seg] = pk;
    }
}

// flags[sb]: bit 0 = a transient in the block's first segment (blksw == 1), bit 1 = one in a later segment
// (then sum(blksw) > 1: positions are unique integers)
__global__ void __launch_bounds__(128)
decide_kernel(SosParams sp, const int32_t* __restrict__ clip_sb0, int n_clips, int L, int nsb_total,
              const double* __restrict__ peaks, uint8_t* __restrict__ flags) {
    const int sb = blockIdx.x * blockDim.x + threadIdx.x;
    if (sb >= nsb_total) return;
    int lo = 0, hi = n_clips;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (clip_sb0[mid] <= sb) lo = mid; else hi = mid;
    }
    const bool first = (sb == clip_sb0[lo]);
    const int nseg = L / MRC_SHORT;
    unsigned f = 0u;
    for (int ch = 0; ch < 2; ++ch) {
        const double* P = peaks + ((size_t)sb * 2 + ch) * nseg;
        double mx = 0.0;
        for (int i = 0; i < nseg; ++i) mx = fmax(mx, P[i]);
        if (!(mx > sp.t0)) continue;                               // :1045
        double prev = first ? 0.0 : peaks[((size_t)(sb - 1) * 2 + ch) * nseg + (nseg - 1)];     // P[:,0] = P[:,nSeg] (:1053)
        for (int i = 0; i < nseg; ++i) {
            if (__dmul_rn(P[i], sp.t1) > prev) f |= (i == 0) ? 1u : 2u;                   // :1047
            prev = P[i];
        }
    }
    flags[sb] = (uint8_t)f;
}

}  // namespace

void launch_transient(cudaStream_t st, const SosParams& sp, const int64_t* clip_off, const int32_t* clip_sb0,
                      int n_clips, const int16_t* pcm, int L, int nsb_total, double* peaks, uint8_t* flags) {
    if (nsb_total <= 0) return;
    const int nt = 2 * nsb_total;
    if (sp.n == 10)
        peaks_kernel<10><<<(nt + 127) / 128, 128, 0, st>>>(sp, clip_off, clip_sb0, n_clips, pcm, L, nsb_total, peaks);
    else
        peaks_kernel<0><<<(nt + 127) / 128, 128, 0, st>>>(sp, clip_off, clip_sb0, n_clips, pcm, L, nsb_total, peaks);
    decide_kernel<<<(nsb_total + 127) / 128, 128, 0, st>>>(sp, clip_sb0, n_clips, L, nsb_total, peaks, flags);
}
