// mrc_train.cu -- SURVEY.md 8 f3: the data-parallel part of Huffman table training.
//   huffman_training_script.py:33-66   every block of every training file goes through EncodeNoHuff
//                                      (codecThem.py:234-260) and each channel's compacted mantissa vector through
//   huffman.py:56-71                   calculateFrequencies(table, data)
// calculateFrequencies is a histogram with a quirk: a value never seen before (always a new maximum, because the
// function also creates every smaller key) resets the counts of ALL smaller values to zero when it is the first new
// value of its call (its local `current_max` restarts at -1 on every call).  So the final table holds the counts of
// the data from the last such reset onward.  The host finds that position from the per-call maxima (callmax_kernel)
// and hist_kernel counts from there.
#include "mrc_internal.cuh"

namespace {

// one warp per call = (block, channel): largest transmitted mantissa, -1 when the channel transmits none
__global__ void __launch_bounds__(256)
callmax_kernel(int L, int ncalls, const uint8_t* __restrict__ alloc, const uint16_t* __restrict__ mant,
               const uint8_t* __restrict__ line2band, int32_t* __restrict__ cmax) {
    const int call = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (call >= ncalls) return;
    const uint8_t* a = alloc + (size_t)call * MRC_BSTRIDE;
    const uint16_t* m = mant + (size_t)call * L;
    int v = -1;
    for (int k = lane; k < L; k += 32)
        if (a[line2band[k]]) v = max(v, (int)m[k]);
    v = __reduce_max_sync(0xffffffffu, v);
    if (lane == 0) cmax[call] = v;
}

// counts of the transmitted mantissas of calls >= first_call; in call first_call only from the first mantissa that
// exceeds `thr` on (thr = -2: from the start)
__global__ void __launch_bounds__(256)
hist_kernel(int L, int ncalls, int first_call, int thr, const uint8_t* __restrict__ alloc,
            const uint16_t* __restrict__ mant, const uint8_t* __restrict__ line2band, unsigned long long* __restrict__ hist) {
    const int call = first_call + blockIdx.x;
    if (call >= ncalls) return;
    const uint8_t* a = alloc + (size_t)call * MRC_BSTRIDE;
    const uint16_t* m = mant + (size_t)call * L;
    __shared__ int s_first;
    if (threadIdx.x == 0) s_first = (call == first_call && thr > -2) ? L : 0;
    __syncthreads();
    if (call == first_call && thr > -2) {
        int mine = L;
        for (int k = threadIdx.x; k < L; k += blockDim.x)
            if (a[line2band[k]] && (int)m[k] > thr) { mine = k; break; }
        atomicMin(&s_first, mine);
        __syncthreads();
    }
    const int first = s_first;
    for (int k = first + threadIdx.x; k < L; k += blockDim.x)
        if (a[line2band[k]]) atomicAdd(&hist[m[k]], 1ull);
}

}  // namespace

void launch_callmax(cudaStream_t st, int L, int ncalls, const uint8_t* alloc, const uint16_t* mant,
                    const uint8_t* line2band, int32_t* cmax) {
    if (ncalls <= 0) return;
    callmax_kernel<<<(ncalls + 7) / 8, 256, 0, st>>>(L, ncalls, alloc, mant, line2band, cmax);
}

void launch_hist(cudaStream_t st, int L, int ncalls, int first_call, int thr, const uint8_t* alloc, const uint16_t* mant,
                 const uint8_t* line2band, unsigned long long* hist) {
    if (ncalls - first_call <= 0) return;
    hist_kernel<<<ncalls - first_call, 256, 0, st>>>(L, ncalls, first_call, thr, alloc, mant, line2band, hist);
}
