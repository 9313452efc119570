// mrc_chain.cu -- K3b: the serial walk.  The only state the reference carries from block to block (and, for
// independent channels, from channel to channel) is codingParams.bitReservoir, one int:
//   bit budget + reservoir      codecThem.py:299-308 (single channel) / :381-396 (joint)
//   water-filling allocation    bitalloc.py:106-155 -- here: how far down the presorted grant list the budget reaches
//   reservoir = int(bitsLeft)   codecThem.py:332 / :503
//   Huffman table choice        codecThem.py:136-203 (strict minimum below the raw size, ties -> lowest index)
//   reservoir += bits_saved     codecThem.py:224 / :274
//   chunk sizes                 pacfileThem.py:651-707 / :825-880
// One warp per clip.  Per block it receives the cost kernel's 16.6 KB record through a TMA bulk copy
// (cp.async.bulk + mbarrier, MRC_CHAIN_STAGES deep, so the next blocks are already in shared memory), finds the
// first 32-token chunk the budget cannot fully pay with one ballot over the chunk maxima, resolves that chunk and
// the tail with warp scans, and reads every total (bits spent, cost under each book) off the chunk checkpoint plus
// a warp reduction.  Outputs: which tokens were granted (one 32-bit mask per chunk) and the ChainBlk record.
#include "mrc_internal.cuh"

namespace {

constexpr int MRC_CHAIN_STAGES = 4;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__global__ void __launch_bounds__(32)
chain_kernel(CodecParams cp, ClipMap cm, int c0, int g0, int nblk_wave, int min_nl, ChainIO io,
             const int32_t* __restrict__ reservoir_in, int32_t* __restrict__ reservoir_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[MRC_CHAIN_STAGES];
    const int lane = threadIdx.x;
    const int clip = c0 + blockIdx.x;
    const int blk0 = cm.clip_blk0[clip], nblk_clip = cm.clip_blk0[clip + 1] - blk0;
    const int b_lo = max(blk0, g0) - blk0, b_hi = min(blk0 + nblk_clip, g0 + nblk_wave) - blk0;   // [b_lo, b_hi)
    if (b_hi <= b_lo) return;
    const int nb = cp.nb;
    const int band_hdr_bits = nb * (cp.n_mant_size_bits + cp.n_scale_bits);

    if (lane == 0) {
        for (int s = 0; s < MRC_CHAIN_STAGES; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int b, int s) {          // lane 0 only
        const unsigned char* src = io.rec + (size_t)(blk0 + b - g0) * MRC_REC_BYTES;
        mbar_expect_tx(&s_bar[s], MRC_REC_BYTES);
        bulk_g2s(smem_raw + (size_t)s * MRC_REC_BYTES, src, MRC_REC_BYTES, &s_bar[s]);
    };
    if (lane == 0)
        for (int i = 0; i < MRC_CHAIN_STAGES && b_lo + i < b_hi; ++i) issue(b_lo + i, i);

    int R;
    long long running;
    if (b_lo == 0) {
        R = reservoir_in ? reservoir_in[clip] : 0;
        running = cp.header_bytes;
    } else {
        R = io.clip_res[clip];
        running = io.clip_run[clip];
    }

    for (int b = b_lo; b < b_hi; ++b) {
        const int it = b - b_lo, s = it % MRC_CHAIN_STAGES;
        mbar_wait(&s_bar[s], (unsigned)((it / MRC_CHAIN_STAGES) & 1));
        const unsigned char* st = smem_raw + (size_t)s * MRC_REC_BYTES;
        const uint32_t* tn = reinterpret_cast<const uint32_t*>(st + MRC_REC_TN);
        const uint4* dd = reinterpret_cast<const uint4*>(st + MRC_REC_D);
        const uint32_t* ck = reinterpret_cast<const uint32_t*>(st + MRC_REC_CK);
        const int32_t* mx = reinterpret_cast<const int32_t*>(st + MRC_REC_MX);

        const bool joint = cp.joint && !(cp.flush_nonjoint && b == nblk_clip - 1);
        const int ngroups = joint ? 1 : 2;
        const int nck = joint ? MRC_NCHUNK : MRC_GROUP_CHUNKS;
        unsigned gmask = 0u;                       // lane k: granted tokens of chunk k
        int table[2] = {MRC_NO_TABLE, MRC_NO_TABLE}, wbits[2] = {0, 0};
        for (int grp = 0; grp < ngroups; ++grp) {
            const int k0 = grp * MRC_GROUP_CHUNKS;
            double B;
            if (joint) {
                B = cp.budget_joint + (double)R;      // += bitReservoir
                B -= 1.0;                             // -= blkswBitA
                B -= 1.0;                             // -= blkswBitB
            } else {
                B = cp.budget_single + (double)R;     // blksw bits already subtracted, then += bitReservoir
            }
            unsigned acc[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            int raw0 = 0, raw1 = 0;
            if (B > 0.0) {
                const int B0 = B >= 2147483647.0 ? 0x7fffffff : (int)floor(B);
                const int mxk = (lane < nck) ? mx[k0 + lane] : (int)0x80000000;
                const unsigned fm = __ballot_sync(0xffffffffu, mxk > B0);
                const int ks = fm ? (__ffs(fm) - 1) : nck - 1;       // first chunk with a refusal (or the last chunk)
                if (lane >= k0 && lane < k0 + ks) gmask = 0xffffffffu;
                const uint32_t* c = ck + (k0 + ks) * MRC_CK_WORDS;
                const int cr0 = (int)c[8], cr1 = (int)c[9];      // bits already spent by the chunks before ks
                if (lane == 0) {                                 // checkpoint totals enter through lane 0
                    const uint4 a = *reinterpret_cast<const uint4*>(c), b2 = *reinterpret_cast<const uint4*>(c + 4);
                    acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; acc[3] = a.w;
                    acc[4] = b2.x; acc[5] = b2.y; acc[6] = b2.z; acc[7] = b2.w;
                    raw0 = cr0; raw1 = cr1;
                }
                int rem = B0 - (cr0 + cr1);
                int k = ks;
                unsigned active = 0xffffffffu;
                while (k < nck && rem >= min_nl) {
                    const int slot = (k0 + k) * 32 + lane;
                    const uint32_t t = tn[slot];
                    const bool valid = t != 0xffffffffu;
                    const int n = (int)(t >> 16), lvl = (int)((t >> 8) & 0xff), bb = (int)(t & 0xff);
                    const bool cand = valid && ((active >> lane) & 1u) && n <= rem;
                    const int cc = cand ? (lvl == 0 ? 2 * n : n) : 0;
                    int incl = cc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const int excl = incl - cc;
                    const unsigned f2 = __ballot_sync(0xffffffffu, cand && n > rem - excl);
                    bool grant;
                    int kk = k;
                    if (f2 == 0u) {
                        grant = cand;
                        rem -= __shfl_sync(0xffffffffu, incl, 31);
                        active = 0xffffffffu;
                        ++k;
                    } else {
                        const int f = __ffs(f2) - 1;              // first refusal: that band is out from here on
                        grant = cand && lane < f;
                        rem -= __shfl_sync(0xffffffffu, excl, f);
                        if (f == 31) { active = 0xffffffffu; ++k; }
                        else active &= ~((2u << f) - 1u);         // resume this chunk after lane f
                    }
                    const unsigned gb = __ballot_sync(0xffffffffu, grant);
                    if (lane == k0 + kk) gmask |= gb;
                    if (grant) {
                        const uint4 d = dd[slot];
                        const bool ch = bb >= nb;
                        acc[0] += ch ? 0u : d.x; acc[1] += ch ? 0u : d.y; acc[2] += ch ? 0u : d.z; acc[3] += ch ? 0u : d.w;
                        acc[4] += ch ? d.x : 0u; acc[5] += ch ? d.y : 0u; acc[6] += ch ? d.z : 0u; acc[7] += ch ? d.w : 0u;
                        if (ch) raw1 += cc; else raw0 += cc;
                    }
                }
            }
            unsigned tot[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) tot[i] = __reduce_add_sync(0xffffffffu, acc[i]);
            const int r0 = (int)__reduce_add_sync(0xffffffffu, (unsigned)raw0);
            const int r1 = (int)__reduce_add_sync(0xffffffffu, (unsigned)raw1);
            const double left = B - (double)(r0 + r1);      // exact: both are integers-plus-a-fixed-fraction < 2^53
            R = (int)left;                                  // int() truncates toward zero
            const int ch_lo = joint ? 0 : grp, ch_hi = joint ? 2 : grp + 1;
            for (int ch = ch_lo; ch < ch_hi; ++ch) {
                const int raw = ch ? r1 : r0;
                int best = raw, tb_ = MRC_NO_TABLE, bits = raw;
                if (!cp.no_huff) {
#pragma unroll
                    for (int t = 0; t < MRC_N_HUFF_TABLES; ++t) {
                        const int cost = (int)(tot[ch * 4 + t] & 0xffffu), wb = (int)(tot[ch * 4 + t] >> 16);
                        if (cost < best) { best = cost; tb_ = t; bits = wb; }
                    }
                }
                R += raw - best;                            // bitReservoir += bits_saved
                table[ch] = tb_;
                wbits[ch] = bits;
            }
        }
        // ---- block outputs ----
        const size_t lb = (size_t)(blk0 + b - g0);
        io.gmask[lb * 32 + lane] = gmask;
        if (lane == 0) {
            ChainBlk o;
            for (int ch = 0; ch < 2; ++ch) {
                int bits = 4 + 1 + 1 + band_hdr_bits + wbits[ch];
                if (joint) bits += (ch == 0) ? (4 * cp.n_scale_bits + nb) : 0;
                else bits += cp.n_scale_bits;
                const int nbytes = (bits + 7) >> 3;
                o.chunk_off[ch] = running;
                o.chunk_bytes[ch] = (unsigned)nbytes;
                o.table[ch] = (unsigned char)table[ch];
                running += 4 + nbytes;
            }
            o.reservoir = R;
            o.pad[0] = o.pad[1] = 0;
            io.cblk[lb] = o;
        }
        __syncwarp();
        if (lane == 0 && b + MRC_CHAIN_STAGES < b_hi) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            issue(b + MRC_CHAIN_STAGES, s);
        }
    }
    if (lane == 0) {
        io.clip_res[clip] = R;
        io.clip_run[clip] = running;
        if (b_hi == nblk_clip) {
            io.clip_bytes[clip] = running;
            if (reservoir_out) reservoir_out[clip] = R;
        }
    }
}

}  // namespace

void launch_chain(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0, int nblk,
                  int min_nlines, ChainIO io, const int32_t* reservoir_in, int32_t* reservoir_out) {
    if (nclips <= 0 || nblk <= 0) return;
    const size_t smem = (size_t)MRC_CHAIN_STAGES * MRC_REC_BYTES;
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    chain_kernel<<<nclips, 32, smem, st>>>(cp, cm, c0, g0, nblk, min_nlines, io, reservoir_in, reservoir_out);
}
