// mrc_quant.cu -- K3: the per-stream serial stage.  One CTA per clip walks the clip's blocks in order, carrying
// the bit reservoir (one int) from block to block and from channel to channel:
//   bit budget + reservoir      codecThem.py:299-308 (single channel) / :381-396 (joint)
//   water-filling allocation    bitalloc.py:106-155 (grant order was sorted by the analysis kernel)
//   reservoir = int(bitsLeft)   codecThem.py:332 / :503
//   scale factors + mantissas   codecThem.py:336-350 / :509-559, quantize.py:114-146, :294-322
//   Huffman table choice        codecThem.py:136-203 (cost rule incl. Q4), reservoir += bits_saved :224 / :274
//   chunk sizes                 pacfileThem.py:651-707 / :825-880
// The next block's hand-off record is prefetched with cp.async while the current one is processed.
#include "mrc_internal.cuh"
#include "mrc_math.cuh"

namespace {

constexpr int QT = 256;     // threads per CTA

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

struct GroupResult {
    int table[2];
    int mant_bits[2];    // mantissa payload bits as written
    int saved[2];
};

// Water-filling over the presorted grant tokens, by one warp.  Token = band | level<<8 (band in [0, 2nb)).
// A token is granted iff nLines[band] <= bitsLeft at its turn (first grant of a band costs 2*nLines but checks
// nLines only -- Q5); bitsLeft only decreases, so a refused band stays refused, which is the reference's
// exclusion.  Returns the integer number of bits spent.
__device__ __forceinline__ int alloc_scan_warp(const uint16_t* tok, int ntok, const int* nl, int min_nl,
                                               double budget, int* alloc, int lane) {
    if (!(budget > 0.0)) return 0;
    const int L0 = (int)floor(budget);
    int rem = L0;
    int pos = 0;
    while (pos < ntok && rem >= min_nl) {
        const int j = pos + lane;
        const bool valid = j < ntok;
        const int t = valid ? tok[j] : 0;
        const int bb = t & 0xff, lvl = t >> 8;
        const int n = valid ? nl[bb] : 0x3fffffff;
        const bool cand = valid && n <= rem;
        const int cc = cand ? (lvl == 0 ? 2 * n : n) : 0;
        int incl = cc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = incl - cc;
        const unsigned fm = __ballot_sync(0xffffffffu, cand && n > rem - excl);
        if (fm == 0u) {
            if (cand) atomicMax(&alloc[bb], lvl + 2);
            rem -= __shfl_sync(0xffffffffu, incl, 31);
            pos += 32;
        } else {
            const int f = __ffs(fm) - 1;
            if (cand && lane < f) atomicMax(&alloc[bb], lvl + 2);
            rem -= __shfl_sync(0xffffffffu, excl, f);
            pos += f + 1;
        }
    }
    return L0 - rem;
}

template <typename T>
struct Stage {
    T* lines;            // [2][L]
    T* bandmax;          // [2][32]
    uint16_t* tokens;    // [768]
};

template <typename T>
__global__ void __launch_bounds__(QT)
quant_kernel(DevTables<T> tb, CodecParams cp, const HuffDev* __restrict__ huff, ClipMap cm, int c0, int g0,
             Handoff<T> ho, QuantOut qo, const int32_t* __restrict__ reservoir_in, int32_t* reservoir_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = tb.L, nb = tb.nb, nb2 = 2 * nb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int clip = c0 + blockIdx.x;
    const int blk0 = cm.clip_blk0[clip], nblk = cm.clip_blk0[clip + 1] - blk0;

    // shared memory carve-up: two prefetch stages, then per-clip constants
    const size_t stage_bytes = (size_t)2 * L * sizeof(T) + 2 * MRC_BSTRIDE * sizeof(T) + MRC_TOK_STRIDE * 2;
    Stage<T> stg[2];
    for (int s = 0; s < 2; ++s) {
        unsigned char* p = smem_raw + s * stage_bytes;
        stg[s].lines = reinterpret_cast<T*>(p);
        stg[s].bandmax = reinterpret_cast<T*>(p + (size_t)2 * L * sizeof(T));
        stg[s].tokens = reinterpret_cast<uint16_t*>(p + (size_t)2 * L * sizeof(T) + 2 * MRC_BSTRIDE * sizeof(T));
    }
    uint8_t* s_l2b = smem_raw + 2 * stage_bytes;                       // [L]
    __shared__ int s_nl[2 * MRC_BSTRIDE];                              // nLines per band of the 2nb-band list
    __shared__ int s_alloc[2 * MRC_BSTRIDE];
    __shared__ int s_sf[2 * MRC_BSTRIDE];
    __shared__ unsigned int s_acc[2][MRC_N_HUFF_TABLES];               // cost | extra<<16 per channel and table
    __shared__ int s_raw[2];
    __shared__ int s_R, s_spent;
    __shared__ HuffDev s_h;

    for (int i = tid; i < L; i += QT) s_l2b[i] = tb.line2band[i];
    for (int i = tid; i < nb2; i += QT) s_nl[i] = tb.band_n[i % nb];
    for (int i = tid; i < (int)(sizeof(HuffDev) / 4); i += QT)
        reinterpret_cast<uint32_t*>(&s_h)[i] = reinterpret_cast<const uint32_t*>(huff)[i];
    if (tid == 0) s_R = reservoir_in ? reservoir_in[clip] : 0;
    int min_nl = 0x7fffffff;
    for (int i = 0; i < nb; ++i) min_nl = min(min_nl, tb.band_n[i]);

    auto prefetch = [&](int b, int s) {
        const size_t lb = (size_t)(blk0 - g0 + b);
        const unsigned char* gl = reinterpret_cast<const unsigned char*>(ho.lines + lb * 2 * L);
        const unsigned char* gm = reinterpret_cast<const unsigned char*>(ho.bandmax + lb * 2 * MRC_BSTRIDE);
        const unsigned char* gt = reinterpret_cast<const unsigned char*>(ho.tokens + lb * MRC_TOK_STRIDE);
        const int nl16 = (int)(2 * L * sizeof(T) / 16), nm16 = (int)(2 * MRC_BSTRIDE * sizeof(T) / 16),
                  nt16 = MRC_TOK_STRIDE * 2 / 16;
        for (int i = tid; i < nl16; i += QT) cp_async16(reinterpret_cast<unsigned char*>(stg[s].lines) + i * 16, gl + i * 16);
        for (int i = tid; i < nm16; i += QT) cp_async16(reinterpret_cast<unsigned char*>(stg[s].bandmax) + i * 16, gm + i * 16);
        for (int i = tid; i < nt16; i += QT) cp_async16(reinterpret_cast<unsigned char*>(stg[s].tokens) + i * 16, gt + i * 16);
    };

    long long running = cp.header_bytes;     // byte offset of the next chunk inside this clip's .pac
    prefetch(0, 0);
    cp_async_commit();

    const int cap = (1 << cp.n_scale_bits) - 1;
    for (int b = 0; b < nblk; ++b) {
        const int s = b & 1;
        if (b + 1 < nblk) prefetch(b + 1, s ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const size_t lb = (size_t)(blk0 - g0 + b);
        const bool joint = cp.joint && !(cp.flush_nonjoint && b == nblk - 1);
        const int ngroups = joint ? 1 : 2;
        GroupResult res;
        for (int grp = 0; grp < ngroups; ++grp) {
            const int ch0 = joint ? 0 : grp, nch = joint ? 2 : 1;
            const int bb0 = ch0 * nb, nbb = nch * nb;                 // bands of this group in the 2nb list
            const uint16_t* tok = stg[s].tokens + (joint ? 0 : grp * nb * MRC_MAX_LEVELS);
            const int ntok = nbb * MRC_MAX_LEVELS;
            if (tid < nb2) s_alloc[tid] = (tid >= bb0 && tid < bb0 + nbb) ? 0 : s_alloc[tid];
            if (tid < 2 * MRC_N_HUFF_TABLES) (&s_acc[0][0])[tid] = 0u;
            __syncthreads();
            // --- budget and allocation (warp 0) ---
            if (warp == 0) {
                double B;
                const int R = s_R;
                if (joint) {
                    B = cp.budget_joint + (double)R;      // += bitReservoir
                    B -= 1.0;                             // -= blkswBitA
                    B -= 1.0;                             // -= blkswBitB
                } else {
                    B = cp.budget_single + (double)R;     // blksw bits already subtracted, then += bitReservoir
                }
                const int spent = alloc_scan_warp(tok, ntok, s_nl, min_nl, B, s_alloc, lane);
                if (lane == 0) {
                    const double left = B - (double)spent;     // exact: see DESIGN.md "bit budget arithmetic"
                    s_R = (int)left;                           // int() truncates toward zero
                    s_spent = spent;
                }
            }
            __syncthreads();
            // --- per band scale factor (Q6: also for zero-allocation bands, with nMantBits = 0) ---
            if (tid >= bb0 && tid < bb0 + nbb) {
                const int ch = tid / nb, bd = tid - ch * nb;
                const double mx = (double)stg[s].bandmax[ch * MRC_BSTRIDE + bd];
                s_sf[tid] = scale_factor_of(mx, cp.n_scale_bits, s_alloc[tid]);
            }
            if (tid < nch) {
                int raw = 0;
                const int ch = ch0 + tid;
                for (int bd = 0; bd < nb; ++bd) raw += s_alloc[ch * nb + bd] * s_nl[bd];
                s_raw[ch] = raw;
            }
            __syncthreads();
            // --- mantissas + Huffman cost of the four tables ---
            for (int ci = 0; ci < nch; ++ci) {
                const int ch = ch0 + ci;
                unsigned int acc[MRC_N_HUFF_TABLES] = {0u, 0u, 0u, 0u};
                uint16_t* om = qo.mant + (lb * 2 + ch) * L;
                for (int k = tid; k < L; k += QT) {
                    const int bd = s_l2b[k];
                    const int Rb = s_alloc[ch * nb + bd];
                    int m = 0;
                    if (Rb) {
                        m = mantissa_of((double)stg[s].lines[ch * L + k], s_sf[ch * nb + bd], cp.n_scale_bits, Rb);
#pragma unroll
                        for (int t = 0; t < MRC_N_HUFF_TABLES; ++t) {
                            const int len = (m < MRC_HUFF_LUT) ? s_h.len[t][m] : 0;
                            if (len) acc[t] += (unsigned)len + ((m == s_h.esc[t]) ? ((unsigned)Rb << 16) : 0u);
                            else acc[t] += (unsigned)(Rb + s_h.esc_len[t]);
                        }
                    }
                    om[k] = (uint16_t)m;
                }
#pragma unroll
                for (int t = 0; t < MRC_N_HUFF_TABLES; ++t) {
                    unsigned int v = acc[t];
                    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) atomicAdd(&s_acc[ch][t], v);
                }
            }
            __syncthreads();
            // --- table choice, reservoir update (thread 0) ---
            if (tid == 0) {
                int R = s_R;
                for (int ci = 0; ci < nch; ++ci) {
                    const int ch = ch0 + ci;
                    const int raw = s_raw[ch];
                    int best = raw, table = MRC_NO_TABLE, bits = raw;
                    for (int t = 0; t < MRC_N_HUFF_TABLES && !cp.no_huff; ++t) {
                        const int cost = (int)(s_acc[ch][t] & 0xffffu), extra = (int)(s_acc[ch][t] >> 16);
                        if (cost < best) { best = cost; table = t; bits = cost + extra; }
                    }
                    R += raw - best;
                    res.table[ch] = table;
                    res.mant_bits[ch] = bits;
                    res.saved[ch] = raw - best;
                }
                s_R = R;
            }
            __syncthreads();
        }
        // --- block outputs ---
        if (tid < nb2) {
            const int ch = tid / nb, bd = tid - ch * nb;
            qo.alloc[(lb * 2 + ch) * MRC_BSTRIDE + bd] = (uint8_t)s_alloc[tid];
            qo.sf[(lb * 2 + ch) * MRC_BSTRIDE + bd] = (uint8_t)s_sf[tid];
        }
        if (tid == 0) {
            for (int ch = 0; ch < 2; ++ch) {
                int bits = 4 + 1 + 1 + nb * (cp.n_mant_size_bits + cp.n_scale_bits) + res.mant_bits[ch];
                if (joint) bits += (ch == 0) ? (4 * cp.n_scale_bits + nb) : 0;
                else bits += cp.n_scale_bits;
                const int nbytes = (bits + 7) >> 3;
                qo.table[lb * 2 + ch] = (uint8_t)res.table[ch];
                qo.chunk_bytes[lb * 2 + ch] = (uint32_t)nbytes;
                qo.chunk_off[lb * 2 + ch] = running;
                running += 4 + nbytes;
            }
            qo.reservoir[lb] = s_R;
        }
        __syncthreads();     // stage s may be overwritten by the prefetch issued two iterations from now
    }
    if (tid == 0) {
        qo.clip_bytes[clip] = running;
        if (reservoir_out) reservoir_out[clip] = s_R;
    }
    (void)cap;
}

}  // namespace

template <typename T>
void launch_quant(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                  const ClipMap& cm, int c0, int nclips_wave, int g0, Handoff<T> ho, QuantOut qo,
                  const int32_t* reservoir_in, int32_t* reservoir_out) {
    if (nclips_wave <= 0) return;
    const size_t stage_bytes = (size_t)2 * tb.L * sizeof(T) + 2 * MRC_BSTRIDE * sizeof(T) + MRC_TOK_STRIDE * 2;
    const size_t smem = 2 * stage_bytes + tb.L;
    cudaFuncSetAttribute(quant_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    quant_kernel<T><<<nclips_wave, QT, smem, st>>>(tb, cp, huff, cm, c0, g0, ho, qo, reservoir_in, reservoir_out);
}

template void launch_quant<double>(cudaStream_t, const DevTables<double>&, const CodecParams&, const HuffDev*,
                                   const ClipMap&, int, int, int, Handoff<double>, QuantOut, const int32_t*,
                                   int32_t*);
template void launch_quant<float>(cudaStream_t, const DevTables<float>&, const CodecParams&, const HuffDev*,
                                  const ClipMap&, int, int, int, Handoff<float>, QuantOut, const int32_t*, int32_t*);
