// mrc_cost.cu -- K3a: everything about the serial stage that does NOT depend on the bit reservoir, done in
// parallel over blocks so that the serial walk (mrc_chain.cu) is left with a few hundred cycles per block.
//
// The reference allocates bits greedily (bitalloc.py:106-155), then quantises every band at its allocation
// (codecThem.py:336-350 / :509-559, quantize.py:114-146, :294-322), then prices the mantissas under the four
// trained Huffman books (codecThem.py:136-203) and credits the saving to the reservoir (:224 / :274).  Only the
// *stopping point* of the greedy loop depends on the reservoir: the order of the grants was fixed by the analysis
// kernel (tokens).  So for every band and every possible allocation 2..16 this kernel quantises the band and
// prices it under the four books (15 * 2L quantisations per block, integer-exact), and re-orders the result into
// grant order as *prefix sums*: entry j holds, for "every token before j granted", the bits spent per channel, the
// cost of each channel under each book, and the bits actually written (they differ from the cost by quirk Q4: a
// mantissa equal to the escape value is priced at the escape code's length but written as escape code + raw
// bits).  The chain kernel then reads every total at the first refused token with two loads.
//
// One CTA per block, 256 threads; thread <-> (band, level) pair, adjacent threads share a band (broadcast reads).
#include "mrc_internal.cuh"
#include "mrc_math.cuh"

namespace {

constexpr int CT = 256;

struct LutEntry {            // per mantissa value 0..64 (65 = "not a key of any book"), 4 x 16-bit fields (book 0 low)
    unsigned long long key;  // code length where the value is a key of the book, else 0
    unsigned long long nk;   // 0xffff where it is not a key
    unsigned long long esc;  // 0xffff where it is the book's escape value
};

__device__ __forceinline__ unsigned long long splat16(unsigned v) {
    const unsigned long long x = v & 0xffffu;
    return x | (x << 16) | (x << 32) | (x << 48);
}

template <typename T>
__global__ void __launch_bounds__(CT)
cost_kernel(DevTables<T> tb, CodecParams cp, const HuffDev* __restrict__ huff, ClipMap cm, int g0, Handoff<T> ho,
            unsigned char* rec, unsigned char* pw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = tb.L, nb = tb.nb, nb2 = 2 * nb, npair = nb2 * MRC_MAX_LEVELS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* s_lines = reinterpret_cast<double*>(smem_raw);                                   // [2][L]
    unsigned long long* s_c4 = reinterpret_cast<unsigned long long*>(s_lines + 2 * L);       // [npair] cost
    unsigned long long* s_w4 = s_c4 + MRC_NSLOT;                                             // [npair] bits written
    __shared__ double s_bmax[2 * MRC_BSTRIDE];
    __shared__ uint16_t s_tok[MRC_TOK_STRIDE];
    __shared__ LutEntry s_lut[MRC_HUFF_LUT + 1];
    __shared__ unsigned s_tot[MRC_NCHUNK][10];     // per chunk: totals -> exclusive prefixes of the 9 words, local max
    __shared__ int s_joint;
    __shared__ int s_blo[MRC_BSTRIDE], s_bn[MRC_BSTRIDE];
    __shared__ unsigned long long s_esclen4;
    // per allocation level (Rb = level + 2) and mantissa value: what the value costs under the four books (.x) and what
    // is actually written (.y), one byte per book (<= 9 + 16 + 16) -- the inner loop below is one 8-byte load, four
    // byte permutes and four adds per line
    __shared__ uint2 s_lv[MRC_MAX_LEVELS][MRC_HUFF_LUT + 1];
    __shared__ unsigned char s_sf[MRC_NSLOT];              // scale factor of every (band, level) pair
    __shared__ int4 s_piece[2 * MRC_BSTRIDE + 2 * 64];     // (band of the 2nb list, first line in s_lines, lines, -)
    __shared__ int s_npiece;

    const size_t lb = cm.list ? (size_t)cm.list[blockIdx.x] : (size_t)blockIdx.x;
    const int g = g0 + (int)lb;
    if (tid == 0) {
        int lo = 0, hi = cm.n_clips;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cm.clip_blk0[mid] <= g) lo = mid; else hi = mid;
        }
        const int last = cm.clip_blk0[lo + 1] - 1;
        s_joint = cp.joint && !(cp.flush_nonjoint && g == last);
        unsigned long long e4 = 0;
        for (int t = 0; t < MRC_N_HUFF_TABLES; ++t) e4 |= (unsigned long long)(huff->esc_len[t] & 0xffff) << (16 * t);
        s_esclen4 = e4;
    }
    if (tid <= MRC_HUFF_LUT) {
        LutEntry e;
        e.key = 0; e.nk = 0; e.esc = 0;
        for (int t = 0; t < MRC_N_HUFF_TABLES; ++t) {
            const unsigned len = (tid < MRC_HUFF_LUT) ? huff->len[t][tid] : 0u;
            if (len) e.key |= (unsigned long long)len << (16 * t);
            else e.nk |= 0xffffull << (16 * t);
            if (tid < MRC_HUFF_LUT && tid == huff->esc[t]) e.esc |= 0xffffull << (16 * t);
        }
        s_lut[tid] = e;
    }
    if (tid < nb) { s_blo[tid] = tb.band_lo[tid]; s_bn[tid] = tb.band_n[tid]; }
    {
        const T* gl = ho.lines + lb * 2 * cp.Lmax;
        for (int i = tid; i < 2 * L; i += CT) s_lines[i] = (double)gl[i];
        const T* gm = ho.bandmax + lb * 2 * MRC_BSTRIDE;
        if (tid < 2 * MRC_BSTRIDE) s_bmax[tid] = (double)gm[tid];
        const uint16_t* gt = ho.tokens + lb * MRC_TOK_STRIDE;
        for (int i = tid; i < MRC_TOK_STRIDE; i += CT) s_tok[i] = gt[i];
    }
    __syncthreads();
    const bool joint = s_joint != 0;
    for (int q = tid; q < MRC_MAX_LEVELS * (MRC_HUFF_LUT + 1); q += CT) {
        const int lvl = q / (MRC_HUFF_LUT + 1), m = q - lvl * (MRC_HUFF_LUT + 1), Rb = lvl + 2;
        const LutEntry e = s_lut[m];
        const unsigned long long rb4 = splat16((unsigned)Rb);
        const unsigned long long c = e.key + (e.nk & (rb4 + s_esclen4));      // 16-bit fields, each < 256
        const unsigned long long w = c + (e.esc & rb4);
        auto pack8 = [](unsigned long long v) {
            return (unsigned)(v & 0xff) | ((unsigned)((v >> 16) & 0xff) << 8) | ((unsigned)((v >> 32) & 0xff) << 16) |
                   ((unsigned)((v >> 48) & 0xff) << 24);
        };
        s_lv[lvl][m] = make_uint2(pack8(c), pack8(w));
    }
    __syncthreads();

    // ---- phase 1: price every (band, level) ---------------------------------------------------------------
    // Work items are (piece of a band, level): bands wider than CP lines are cut into pieces of at most CP lines (the widest
    // band holds a sixth of all lines: one thread per (band, level) made everybody wait for its 363 iterations); adjacent
    // threads share a piece (broadcast reads of the lines) and the pieces of a band add up in its accumulator.
    constexpr int CP = 32;
    unsigned* const acc = reinterpret_cast<unsigned*>(s_c4);            // [npair][2] cost; s_w4 likewise: bits written
    unsigned* const accw = reinterpret_cast<unsigned*>(s_w4);
    for (int p = tid; p < npair; p += CT) {
        const int bb = p / MRC_MAX_LEVELS, lvl = p - bb * MRC_MAX_LEVELS;
        const int ch = bb >= nb, bd = bb - ch * nb;
        s_sf[p] = (unsigned char)scale_factor_of(s_bmax[ch * MRC_BSTRIDE + bd], cp.n_scale_bits, lvl + 2);
        acc[2 * p] = acc[2 * p + 1] = 0u;
        accw[2 * p] = accw[2 * p + 1] = 0u;
    }
    // Piece list, by one warp: first every full piece (CP lines), then the partial ones from the widest band down -- sizes
    // descend, so the lanes of a warp (two or three neighbouring pieces at 15 levels each) run the same number of lines and
    // every round of the item loop below gives all threads about the same work (in band order a warp mixed 4-line and
    // 32-line pieces, and 30 % of the kernel's stall samples were threads waiting for the longest item).
    if (warp == 0) {
        int nfull[2], part[2], tot_full = 0, tot_part = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int bb = lane + 32 * h;
            const int n = bb < nb2 ? s_bn[bb >= nb ? bb - nb : bb] : 0;
            nfull[h] = n / CP;
            part[h] = (n % CP) ? 1 : 0;
        }
        // exclusive prefix of the full pieces in band order; of the partial ones in reverse band order
        int fo[2], po[2];
        {
            int a = nfull[0], b = nfull[1];
            int ia = a, ib = b;
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb2 = __shfl_up_sync(0xffffffffu, ib, o);
                if (lane >= o) { ia += ta; ib += tb2; }
            }
            const int suma = __shfl_sync(0xffffffffu, ia, 31), sumb = __shfl_sync(0xffffffffu, ib, 31);
            fo[0] = ia - a; fo[1] = suma + ib - b;
            tot_full = suma + sumb;
            int pa = part[0], pb = part[1];
            int ja = pa, jb = pb;                      // inclusive suffix sums (bands above, then this one)
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_down_sync(0xffffffffu, ja, o), tb2 = __shfl_down_sync(0xffffffffu, jb, o);
                if (lane + o < 32) { ja += ta; jb += tb2; }
            }
            const int allb = __shfl_sync(0xffffffffu, jb, 0);
            const int alla = __shfl_sync(0xffffffffu, ja, 0);
            po[1] = jb - pb;                           // partial pieces of higher bands (all in the second half)
            po[0] = allb + ja - pa;
            tot_part = alla + allb;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int bb = lane + 32 * h;
            if (bb >= nb2) continue;
            const int ch = bb >= nb, bd = bb - ch * nb, first = ch * L + s_blo[bd];
            for (int q = 0; q < nfull[h]; ++q) s_piece[fo[h] + q] = make_int4(bb, first + q * CP, CP, 0);
            if (part[h]) s_piece[tot_full + po[h]] = make_int4(bb, first + nfull[h] * CP, s_bn[bd] % CP, 0);
        }
        if (lane == 0) s_npiece = tot_full + tot_part;
    }
    __syncthreads();
    const int nitem = s_npiece * MRC_MAX_LEVELS;
    for (int it = tid; it < nitem; it += CT) {
        const int pc = it / MRC_MAX_LEVELS, lvl = it - pc * MRC_MAX_LEVELS;
        const int4 pi = s_piece[pc];
        const int p = pi.x * MRC_MAX_LEVELS + lvl, Rb = lvl + 2;
        const int sf = s_sf[p];
        const double* x = s_lines + pi.y;
        const uint2* __restrict__ tabl = s_lv[lvl];
        unsigned c01 = 0, c23 = 0, w01 = 0, w23 = 0;        // books 0|1 and 2|3 in 16-bit fields: sums < 363 * 41
        for (int i = 0; i < pi.z; ++i) {
            const int m = mantissa_of(x[i], sf, cp.n_scale_bits, Rb);
            const uint2 e = tabl[m < MRC_HUFF_LUT ? m : MRC_HUFF_LUT];
            c01 += __byte_perm(e.x, 0u, 0x4140);
            c23 += __byte_perm(e.x, 0u, 0x4342);
            w01 += __byte_perm(e.y, 0u, 0x4140);
            w23 += __byte_perm(e.y, 0u, 0x4342);
        }
        atomicAdd(&acc[2 * p], c01);
        atomicAdd(&acc[2 * p + 1], c23);
        atomicAdd(&accw[2 * p], w01);
        atomicAdd(&accw[2 * p + 1], w23);
    }
    __syncthreads();

    // ---- phase 2: grant order; per 32-token chunk: local exclusive prefix sums, totals, local maximum -------
    unsigned char* out = rec + lb * (size_t)MRC_REC_BYTES;
    uint32_t* o_tn = reinterpret_cast<uint32_t*>(out + MRC_REC_TN);
    uint32_t* o_cp = reinterpret_cast<uint32_t*>(out + MRC_REC_CP);
    uint4* o_pc = reinterpret_cast<uint4*>(out + MRC_REC_PC);
    int32_t* o_mx = reinterpret_cast<int32_t*>(out + MRC_REC_MX);
    uint4* o_pw = reinterpret_cast<uint4*>(pw + lb * (size_t)MRC_PW_BYTES);
    const int per_group = nb * MRC_MAX_LEVELS;
    for (int k = warp; k < MRC_NCHUNK; k += CT / 32) {
        const int slot = k * 32 + lane;
        int src = -1;                                   // index into the analysis kernel's sorted token list
        if (joint) { if (slot < 2 * per_group) src = slot; }
        else {
            const int grp = slot >= MRC_GROUP_SLOTS, i = slot - grp * MRC_GROUP_SLOTS;
            if (i < per_group) src = grp * per_group + i;
        }
        uint32_t tn = 0xffffffffu;
        unsigned v[9] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // [0] bits (ch0 | ch1<<16), [1..4] cost, [5..8] written
        int n = 0;
        const bool valid = src >= 0;
        if (valid) {
            const unsigned tok = s_tok[src];
            const int bb = tok & 0xff, lvl = tok >> 8;
            const int ch = bb >= nb, bd = bb - ch * nb;
            n = s_bn[bd];
            tn = tok | ((uint32_t)n << 16);
            v[0] = (unsigned)(lvl == 0 ? 2 * n : n) << (16 * ch);
            const int p = bb * MRC_MAX_LEVELS + lvl;
            const unsigned long long c1 = s_c4[p], w1 = s_w4[p];
            const unsigned long long c0 = lvl ? s_c4[p - 1] : 0ull, w0 = lvl ? s_w4[p - 1] : 0ull;
            unsigned dc[4], dw[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                dc[t] = (unsigned)((c1 >> (16 * t)) & 0xffff) - (unsigned)((c0 >> (16 * t)) & 0xffff);
                dw[t] = (unsigned)((w1 >> (16 * t)) & 0xffff) - (unsigned)((w0 >> (16 * t)) & 0xffff);
            }
            // two books per word: lo + 65536*hi (mod 2^32); a sum decodes field by field while every total < 65536
            v[1 + 2 * ch] = dc[0] + (dc[1] << 16);
            v[2 + 2 * ch] = dc[2] + (dc[3] << 16);
            v[5 + 2 * ch] = dw[0] + (dw[1] << 16);
            v[6 + 2 * ch] = dw[2] + (dw[3] << 16);
        }
        unsigned inc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) inc[i] = v[i];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const unsigned u = __shfl_up_sync(0xffffffffu, inc[i], o);
                if (lane >= o) inc[i] += u;
            }
        }
        o_tn[slot] = tn;
        o_cp[slot] = inc[0] - v[0];                                       // chunk-local; offsets added in phase 3
        o_pc[slot] = make_uint4(inc[1] - v[1], inc[2] - v[2], inc[3] - v[3], inc[4] - v[4]);
        o_pw[slot] = make_uint4(inc[5] - v[5], inc[6] - v[6], inc[7] - v[7], inc[8] - v[8]);
        const unsigned ex0 = inc[0] - v[0];
        const int need = valid ? (int)((ex0 & 0xffffu) + (ex0 >> 16)) + n : (int)0x80000000;
        const int mxl = __reduce_max_sync(0xffffffffu, need);
        if (lane == 31) {
#pragma unroll
            for (int i = 0; i < 9; ++i) s_tot[k][i] = inc[i];
            s_tot[k][9] = (unsigned)mxl;
        }
    }
    __syncthreads();
    // ---- phase 3: exclusive prefix over the chunks (restarting at the second group of a non-joint block) ---
    if (tid < 9) {
        unsigned run = 0;
        for (int k = 0; k < MRC_NCHUNK; ++k) {
            if (k == MRC_GROUP_CHUNKS && !joint) run = 0;
            const unsigned t = s_tot[k][tid];
            s_tot[k][tid] = run;
            run += t;
        }
    }
    __syncthreads();
    if (tid == 32) {
        int mx = (int)0x80000000;
        for (int k = 0; k < MRC_NCHUNK; ++k) {
            if (k == MRC_GROUP_CHUNKS && !joint) mx = (int)0x80000000;
            const int mxl = (int)s_tot[k][9];
            const unsigned sp = s_tot[k][0];
            if (mxl != (int)0x80000000) mx = max(mx, (int)((sp & 0xffffu) + (sp >> 16)) + mxl);
            o_mx[k] = mx;
        }
        for (int k = MRC_NCHUNK; k < 32; ++k) o_mx[k] = 0x7fffffff;
        // the block's own budget and band count travel with the record (they differ between block geometries)
        o_mx[MRC_MX_K] = joint ? cp.k_joint : cp.k_single;
        o_mx[MRC_MX_FRAC] = joint ? cp.frac_joint : cp.frac_single;
        o_mx[MRC_MX_NB] = nb;
        int mn = 0x7fffffff;
        for (int b = 0; b < nb; ++b) mn = min(mn, s_bn[b]);
        o_mx[MRC_MX_MINNL] = mn;
    }
    for (int k = warp; k < MRC_NCHUNK; k += CT / 32) {      // same thread re-reads what it wrote above
        const int slot = k * 32 + lane;
        o_cp[slot] += s_tot[k][0];
        uint4 a = o_pc[slot];
        a.x += s_tot[k][1]; a.y += s_tot[k][2]; a.z += s_tot[k][3]; a.w += s_tot[k][4];
        o_pc[slot] = a;
        uint4 w = o_pw[slot];
        w.x += s_tot[k][5]; w.y += s_tot[k][6]; w.z += s_tot[k][7]; w.w += s_tot[k][8];
        o_pw[slot] = w;
    }
}

}  // namespace

template <typename T>
void launch_cost(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                 const ClipMap& cm, int g0, int nblk, Handoff<T> ho, unsigned char* rec, unsigned char* pw) {
    if (nblk <= 0) return;
    const size_t smem = (size_t)2 * tb.L * sizeof(double) + 2 * MRC_NSLOT * sizeof(unsigned long long);
    cudaFuncSetAttribute(cost_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cost_kernel<T><<<nblk, CT, smem, st>>>(tb, cp, huff, cm, g0, ho, rec, pw);
}

template void launch_cost<double>(cudaStream_t, const DevTables<double>&, const CodecParams&, const HuffDev*,
                                  const ClipMap&, int, int, Handoff<double>, unsigned char*, unsigned char*);
template void launch_cost<float>(cudaStream_t, const DevTables<float>&, const CodecParams&, const HuffDev*,
                                 const ClipMap&, int, int, Handoff<float>, unsigned char*, unsigned char*);
