// mrc_chain.cu -- K3b/K3c/K3d: the serial walk and its parallel replay.
//
// The only state the reference carries from block to block (and, for independent channels, from channel to
// channel) is codingParams.bitReservoir, one int:
//   bit budget + reservoir      codecThem.py:299-308 (single channel) / :381-396 (joint)
//   water-filling allocation    bitalloc.py:106-155 -- here: how far down the presorted grant list the budget reaches
//   reservoir = int(bitsLeft)   codecThem.py:332 / :503
//   Huffman table choice        codecThem.py:136-203 (strict minimum below the raw size, ties -> lowest index)
//   reservoir += bits_saved     codecThem.py:224 / :274
//   chunk sizes                 pacfileThem.py:651-707 / :825-880
//
// chain_kernel   (serial, one warp per clip): computes nothing but the reservoir sequence.  Per block it receives
//                the cost kernel's 18.5 KB record through a TMA bulk copy (cp.async.bulk + mbarrier, a ring of
//                MRC_CHAIN_STAGES), finds the first refused token with two ballots (over the chunk maxima, then
//                inside that chunk), reads the totals "everything before it granted" with two loads, resolves the
//                short tail after the first refusal with warp scans, and updates the reservoir in integers.
// finish_kernel  (parallel, one warp per block): replays the block from the reservoir the chain recorded and
//                writes what the chain left out: grant masks, table ids, bits written, chunk sizes.
// offsets_kernel (parallel, one warp per clip): chunk sizes -> byte offsets inside the clip's .pac.
#include <algorithm>
#include <cstdio>
#include "mrc_internal.cuh"
#include "mrc_tma.cuh"

namespace {

constexpr int MRC_CHAIN_STAGES = 4;
constexpr unsigned INVALID_TOKEN = 0xffffffffu;

__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ uint4 sub4(uint4 a, uint4 b) { return make_uint4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ uint4 redux4(uint4 a) {
    return make_uint4(__reduce_add_sync(0xffffffffu, a.x), __reduce_add_sync(0xffffffffu, a.y),
                      __reduce_add_sync(0xffffffffu, a.z), __reduce_add_sync(0xffffffffu, a.w));
}

// What one group (a joint block, or one channel of a non-joint block) comes to for budget B0 = k + reservoir.
struct GroupTotals {
    unsigned spent;     // bits spent, channel 0 | channel 1 << 16
    uint4 cost;         // Huffman cost of the granted mantissas, {ch0 b0|b1<<16, ch0 b2|b3<<16, ch1 .., ch1 ..}
    uint4 wbits;        // bits actually written (FULL only)
};

// The greedy allocation over the presorted grant tokens of chunks [k0, k0+nck) (bitalloc.py:131-149): a token is
// granted iff its band's nLines <= bits left at its turn; the first grant of a band costs 2*nLines but checks only
// nLines (Q5); bits left only decrease, so a refused band stays refused, which is the reference's exclusion.
// All lanes return the same totals.  FULL also accumulates the written bits and the grant masks (lane k <-> chunk k).
template <bool FULL>
__device__ __forceinline__ GroupTotals walk_group(const uint32_t* tn, const uint32_t* cpre, const uint4* pc,
                                                  const uint4* pw, const int32_t* mx, int k0, int nck, int B0,
                                                  int min_nl, int lane, unsigned& gmask, unsigned& n_iter) {
    GroupTotals g;
    g.spent = 0u;
    g.cost = make_uint4(0u, 0u, 0u, 0u);
    g.wbits = make_uint4(0u, 0u, 0u, 0u);
    if (B0 <= 0) return g;
    // first refused token: the chunk from the running maxima, the lane from the stored prefix (no scan)
    const int mxk = (lane < nck) ? mx[k0 + lane] : (int)0x80000000;
    const unsigned fm = __ballot_sync(0xffffffffu, mxk > B0);
    int k, f;
    if (fm) {
        k = __ffs(fm) - 1;
        const int slot = (k0 + k) * 32 + lane;
        const uint32_t t = tn[slot];
        const uint32_t c = cpre[slot];
        const int before = (int)((c & 0xffffu) + (c >> 16));
        const unsigned f2 = __ballot_sync(0xffffffffu, t != INVALID_TOKEN && (int)(t >> 16) > B0 - before);
        f = __ffs(f2) - 1;                           // f2 != 0: this chunk's maximum exceeded B0
    } else {
        k = nck - 1;                                 // everything granted: the group's last slot is never a token,
        f = 31;                                      // so its prefix is the group total
    }
    const int jstar = (k0 + k) * 32 + f;
    MRC_ASSERT(k >= 0 && f >= 0 && jstar < MRC_NSLOT);
    g.spent = cpre[jstar];
    g.cost = pc[jstar];
    if (FULL) {
        g.wbits = pw[jstar];
        if (lane >= k0 && lane < k0 + k) gmask = 0xffffffffu;
        if (lane == k0 + k) gmask |= (1u << f) - 1u;
    }
    int rem = B0 - (int)((g.spent & 0xffffu) + (g.spent >> 16));
    // tail: tokens after the first refusal that still fit (few: rem < nLines of the refused band)
    unsigned active = ~((2u << f) - 1u);             // lanes after f
    if (f == 31) { active = 0xffffffffu; ++k; }
    unsigned a_spent = 0u;
    uint4 a_cost = make_uint4(0u, 0u, 0u, 0u), a_w = make_uint4(0u, 0u, 0u, 0u);
    bool any = false;
    while (k < nck && rem >= min_nl) {
        ++n_iter;
        const int slot = (k0 + k) * 32 + lane;
        MRC_ASSERT(slot < MRC_NSLOT);
        const uint32_t t = tn[slot];
        const int n = (int)(t >> 16);
        const bool cand = t != INVALID_TOKEN && ((active >> lane) & 1u) && n <= rem;
        const int cc = cand ? ((t & 0xff00u) ? n : 2 * n) : 0;      // level 0: 2 bits per line
        int incl = cc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = incl - cc;
        const unsigned f2 = __ballot_sync(0xffffffffu, cand && n > rem - excl);
        bool grant;
        const int kk = k;
        if (f2 == 0u) {
            grant = cand;
            rem -= __shfl_sync(0xffffffffu, incl, 31);
            active = 0xffffffffu;
            ++k;
        } else {
            const int ff = __ffs(f2) - 1;            // refused: that band is out from here on
            grant = cand && lane < ff;
            rem -= __shfl_sync(0xffffffffu, excl, ff);
            if (ff == 31) { active = 0xffffffffu; ++k; }
            else active &= ~((2u << ff) - 1u);       // resume this chunk after lane ff
        }
        if (FULL) {
            const unsigned gb = __ballot_sync(0xffffffffu, grant);
            if (lane == k0 + kk) gmask |= gb;
        }
        if (grant) {                                 // this token's own contribution = difference of the prefixes
            any = true;
            a_spent += cpre[slot + 1] - cpre[slot];
            a_cost = add4(a_cost, sub4(pc[slot + 1], pc[slot]));
            if (FULL) a_w = add4(a_w, sub4(pw[slot + 1], pw[slot]));
        }
    }
    if (__any_sync(0xffffffffu, any)) {
        g.spent += __reduce_add_sync(0xffffffffu, a_spent);
        g.cost = add4(g.cost, redux4(a_cost));
        if (FULL) g.wbits = add4(g.wbits, redux4(a_w));
    }
    return g;
}

// reservoir after the group: int(bitsLeft) truncates toward zero, then += bits_saved per channel
__device__ __forceinline__ int reservoir_after(const GroupTotals& g, int B0, int fracpos, int no_huff, int* table,
                                               int* wbits) {
    const int left = B0 - (int)((g.spent & 0xffffu) + (g.spent >> 16));
    int R = left >= 0 ? left : left + fracpos;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const int raw = (int)((g.spent >> (16 * ch)) & 0xffffu);
        const unsigned c01 = ch ? g.cost.z : g.cost.x, c23 = ch ? g.cost.w : g.cost.y;
        const int c[4] = {(int)(c01 & 0xffffu), (int)(c01 >> 16), (int)(c23 & 0xffffu), (int)(c23 >> 16)};
        int best = raw, tb = MRC_NO_TABLE;
        if (!no_huff) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (c[t] < best) { best = c[t]; tb = t; }
        }
        R += raw - best;
        if (table) {
            const unsigned w01 = ch ? g.wbits.z : g.wbits.x, w23 = ch ? g.wbits.w : g.wbits.y;
            const int w[4] = {(int)(w01 & 0xffffu), (int)(w01 >> 16), (int)(w23 & 0xffffu), (int)(w23 >> 16)};
            table[ch] = tb;
            wbits[ch] = tb == MRC_NO_TABLE ? raw : w[tb];
        }
    }
    return R;
}

__global__ void __launch_bounds__(32)
chain_kernel(CodecParams cp, ClipMap cm, int c0, int g0, int nblk_wave, ChainIO io,
             const int32_t* __restrict__ reservoir_in, int32_t* __restrict__ reservoir_out,
             unsigned long long* __restrict__ iter_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[MRC_CHAIN_STAGES];
    const int lane = threadIdx.x;
    const int clip = c0 + blockIdx.x;
    const int blk0 = cm.clip_blk0[clip], nblk_clip = cm.clip_blk0[clip + 1] - blk0;
    const int b_lo = max(blk0, g0) - blk0, b_hi = min(blk0 + nblk_clip, g0 + nblk_wave) - blk0;   // [b_lo, b_hi)
    if (b_hi <= b_lo) return;

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

    int R = (b_lo == 0) ? (reservoir_in ? reservoir_in[clip] : 0) : io.clip_res[clip];
    unsigned n_iter = 0, dummy = 0;

    for (int b = b_lo; b < b_hi; ++b) {
        const int it = b - b_lo, s = it % MRC_CHAIN_STAGES;
        mbar_wait(&s_bar[s], (unsigned)((it / MRC_CHAIN_STAGES) & 1));
        const unsigned char* st = smem_raw + (size_t)s * MRC_REC_BYTES;
        const uint32_t* tn = reinterpret_cast<const uint32_t*>(st + MRC_REC_TN);
        const uint32_t* cpre = reinterpret_cast<const uint32_t*>(st + MRC_REC_CP);
        const uint4* pc = reinterpret_cast<const uint4*>(st + MRC_REC_PC);
        const int32_t* mx = reinterpret_cast<const int32_t*>(st + MRC_REC_MX);
        const bool joint = cp.joint && !(cp.flush_nonjoint && b == nblk_clip - 1);
        const int R0 = R;
        int R1 = R;
        const int K = mx[MRC_MX_K], frac = mx[MRC_MX_FRAC];      // this block's budget (its geometry's)
        const int min_nl = mx[MRC_MX_MINNL];
        if (joint) {
            const int B0 = K + R;
            const GroupTotals g = walk_group<false>(tn, cpre, pc, nullptr, mx, 0, MRC_NCHUNK, B0, min_nl, lane, dummy, n_iter);
            R = reservoir_after(g, B0, frac, cp.no_huff, nullptr, nullptr);
        } else {
            int B0 = K + R;
            GroupTotals g = walk_group<false>(tn, cpre, pc, nullptr, mx, 0, MRC_GROUP_CHUNKS, B0, min_nl, lane, dummy, n_iter);
            R = R1 = reservoir_after(g, B0, frac, cp.no_huff, nullptr, nullptr);
            B0 = K + R;
            g = walk_group<false>(tn, cpre, pc, nullptr, mx, MRC_GROUP_CHUNKS, MRC_GROUP_CHUNKS, B0, min_nl, lane, dummy, n_iter);
            R = reservoir_after(g, B0, frac, cp.no_huff, nullptr, nullptr);
        }
        if (lane == 0) io.rsv[blk0 + b - g0] = make_int4(R0, R1, R, 0);
        __syncwarp();                               // every lane is done reading stage s (generic-proxy reads need
        if (lane == 0 && b + MRC_CHAIN_STAGES < b_hi)    // no proxy fence before the async-proxy overwrite)
            issue(b + MRC_CHAIN_STAGES, s);
    }
    if (lane == 0) {
        if (iter_counter) atomicAdd(iter_counter, (unsigned long long)n_iter);
        io.clip_res[clip] = R;
        if (b_hi == nblk_clip && reservoir_out) reservoir_out[clip] = R;
    }
}

// ---- single-stream fast path -------------------------------------------------------------------------------
// When a wave holds long stretches of one clip the serial walk above is the critical path of the whole encode.
// table_kernel (parallel, one CTA per block) then tabulates the block's reservoir map R_in -> R_out for every
// R_in in [r_lo, r_lo + ntab): one thread per R_in walks the grant tokens serially from the first token the
// smallest budget refuses (everything before it is granted for every larger budget as well).  r_lo = -(largest
// band + 1, rounded up): int(bitsLeft) cannot go lower (Q5 overspends by at most nLines), and savings push R_in up,
// not down.  Two more entries give the closed form for "every token granted" (silence: R grows without bound).
// chain_table_kernel then needs one shared-memory load per block; budgets outside the table and not all-granting
// (rare: the blocks right after a Huffman table wins big) take the complete walk on the global record.
constexpr int TAB_THREADS = 256;

__global__ void __launch_bounds__(TAB_THREADS)
table_kernel(CodecParams cp, ClipMap cm, int g0, ChainIO io, int r_lo, int ntab, int tabw, int* __restrict__ tab) {
    __shared__ uint2 s_nc[MRC_NSLOT];            // per token: lines of its band (never granted: INT_MAX), bits it costs
    __shared__ uint32_t s_dsp[MRC_NSLOT];
    __shared__ uint4 s_dpc[MRC_NSLOT];
    __shared__ int s_mx[32];
    __shared__ int s_joint;
    const int tid = threadIdx.x;
    const size_t lb = blockIdx.x;
    const int g = g0 + (int)lb;
    const unsigned char* rec = io.rec + lb * (size_t)MRC_REC_BYTES;
    const uint32_t* tn = reinterpret_cast<const uint32_t*>(rec + MRC_REC_TN);
    const uint32_t* cpre = reinterpret_cast<const uint32_t*>(rec + MRC_REC_CP);
    const uint4* pc = reinterpret_cast<const uint4*>(rec + MRC_REC_PC);
    const int32_t* mx = reinterpret_cast<const int32_t*>(rec + MRC_REC_MX);
    if (tid == 0) {
        int lo = 0, hi = cm.n_clips;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cm.clip_blk0[mid] <= g) lo = mid; else hi = mid;
        }
        s_joint = cp.joint && !(cp.flush_nonjoint && g == cm.clip_blk0[lo + 1] - 1);
    }
    if (tid < 32) s_mx[tid] = mx[tid];
    for (int j = tid; j < MRC_NSLOT; j += TAB_THREADS) {
        const uint32_t t = tn[j];
        const unsigned n = t >> 16;
        s_nc[j] = (t == INVALID_TOKEN) ? make_uint2(0x7fffffffu, 0u) : make_uint2(n, (t & 0xff00u) ? n : 2u * n);
        const int j1 = (j + 1 < MRC_NSLOT) ? j + 1 : j;          // the last slot is never a token
        s_dsp[j] = cpre[j1] - cpre[j];
        s_dpc[j] = sub4(pc[j1], pc[j]);
    }
    __syncthreads();
    const bool joint = s_joint != 0;
    const int ngroups = joint ? 1 : 2, nck = joint ? MRC_NCHUNK : MRC_GROUP_CHUNKS;
    const int K = s_mx[MRC_MX_K], frac = s_mx[MRC_MX_FRAC], min_nl = s_mx[MRC_MX_MINNL];
    int* out = tab + lb * (size_t)(2 * tabw);
    for (int grp = 0; grp < 2; ++grp) {
        int* o = out + grp * tabw;
        if (grp >= ngroups) {                    // unused second table of a joint block
            for (int i = tid; i < tabw; i += TAB_THREADS) o[i] = 0;
            continue;
        }
        const int k0 = grp * MRC_GROUP_CHUNKS;
        const int jend = (k0 + nck) * 32;
        for (int base = 0; base < ntab; base += TAB_THREADS) {
            const int idx = base + tid;
            const int B0 = K + r_lo + idx;
            // first chunk the smallest budget of this warp (lane 0's) cannot fully pay: everything before it is granted
            // for all 32 budgets, and the prefix sums at its start stand for that
            const int b0min = __shfl_sync(0xffffffffu, B0, 0);
            const unsigned over = __ballot_sync(0xffffffffu, (tid & 31) < nck && s_mx[k0 + (tid & 31)] > b0min);
            const int klo = over ? __ffs(over) - 1 : nck - 1;
            const int j0 = (k0 + klo) * 32;
            const unsigned sp0 = cpre[j0];
            const uint4 pc0 = pc[j0];
            GroupTotals gt;
            gt.spent = sp0; gt.cost = pc0; gt.wbits = make_uint4(0u, 0u, 0u, 0u);
            int rem = B0 - (int)((sp0 & 0xffffu) + (sp0 >> 16));
            for (int jb = j0; jb < jend; jb += 16) {
                if (__all_sync(0xffffffffu, rem < min_nl)) break;
                // which of the 16 tokens this budget grants: one load, one compare, one subtraction and one mask bit each;
                // the sums of what was granted -- few tokens past the first refusal -- are added afterwards
                unsigned granted = 0u;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint2 nc = s_nc[jb + i];
                    if ((int)nc.x <= rem) {
                        rem -= (int)nc.y;
                        granted |= 1u << i;
                    }
                }
                while (granted) {
                    const int j = jb + __ffs(granted) - 1;
                    granted &= granted - 1;
                    gt.spent += s_dsp[j];
                    gt.cost = add4(gt.cost, s_dpc[j]);
                }
            }
            if (idx < ntab) {
                if (B0 <= 0) { gt.spent = 0u; gt.cost = make_uint4(0u, 0u, 0u, 0u); }
                o[idx] = reservoir_after(gt, B0, frac, cp.no_huff, nullptr, nullptr);
            }
        }
        if (tid == 0) {                          // every token granted: R_out = B0 + o[ntab] once B0 >= o[ntab+1]
            GroupTotals gt;
            gt.spent = cpre[jend - 1]; gt.cost = pc[jend - 1]; gt.wbits = make_uint4(0u, 0u, 0u, 0u);
            const int big = 1 << 28;
            o[ntab] = reservoir_after(gt, big, frac, cp.no_huff, nullptr, nullptr) - big;
            o[ntab + 1] = s_mx[k0 + nck - 1];
            for (int i = ntab + 2; i < tabw; ++i) o[i] = 0;
        }
    }
}

constexpr int TAB_STAGES = 6;

__global__ void __launch_bounds__(32)
chain_table_kernel(CodecParams cp, ClipMap cm, int c0, int g0, int nblk_wave, ChainIO io, int r_lo,
                   int ntab, int tabw, const int* __restrict__ tab, const int32_t* __restrict__ reservoir_in,
                   int32_t* __restrict__ reservoir_out, unsigned long long* __restrict__ iter_counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[TAB_STAGES];
    const int lane = threadIdx.x;
    const int clip = c0 + blockIdx.x;
    const int blk0 = cm.clip_blk0[clip], nblk_clip = cm.clip_blk0[clip + 1] - blk0;
    const int b_lo = max(blk0, g0) - blk0, b_hi = min(blk0 + nblk_clip, g0 + nblk_wave) - blk0;
    if (b_hi <= b_lo) return;
    // a stage = the block's table (for the common case) followed by its full record (for the complete walk)
    const unsigned tbytes = (unsigned)(2 * tabw * 4), sbytes = tbytes + MRC_REC_BYTES;
    if (lane == 0) {
        for (int s = 0; s < TAB_STAGES; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int b, int s) {          // lane 0 only
        unsigned char* dst = smem_raw + (size_t)s * sbytes;
        mbar_expect_tx(&s_bar[s], sbytes);
        bulk_g2s(dst, tab + (size_t)(blk0 + b - g0) * (2 * tabw), tbytes, &s_bar[s]);
        bulk_g2s(dst + tbytes, io.rec + (size_t)(blk0 + b - g0) * MRC_REC_BYTES, MRC_REC_BYTES, &s_bar[s]);
    };
    if (lane == 0)
        for (int i = 0; i < TAB_STAGES && b_lo + i < b_hi; ++i) issue(b_lo + i, i);
    int R = (b_lo == 0) ? (reservoir_in ? reservoir_in[clip] : 0) : io.clip_res[clip];
    unsigned n_slow = 0, dummy = 0;
    int s = 0;
    unsigned parity = 0;
    for (int b = b_lo; b < b_hi; ++b) {
        mbar_wait(&s_bar[s], parity);
        const unsigned char* stg = smem_raw + (size_t)s * sbytes;
        const int* T0 = reinterpret_cast<const int*>(stg);
        const bool joint = cp.joint && !(cp.flush_nonjoint && b == nblk_clip - 1);
        const int32_t* mxs = reinterpret_cast<const int32_t*>(stg + tbytes + MRC_REC_MX);
        const int K = mxs[MRC_MX_K], frac = mxs[MRC_MX_FRAC], min_nl = mxs[MRC_MX_MINNL];
        const int R0 = R;
        int R1 = R;
        const int ngroups = joint ? 1 : 2;
        for (int grp = 0; grp < ngroups; ++grp) {
            const int* T = T0 + grp * tabw;
            const int B0 = K + R;
            const int idx = R - r_lo;
            if ((unsigned)idx < (unsigned)ntab) R = T[idx];
            else if (B0 >= T[ntab + 1]) R = B0 + T[ntab];
            else {                               // outside the table and not all-granting: the complete walk
                const unsigned char* rec = stg + tbytes;
                const GroupTotals gt = walk_group<false>(
                    reinterpret_cast<const uint32_t*>(rec + MRC_REC_TN), reinterpret_cast<const uint32_t*>(rec + MRC_REC_CP),
                    reinterpret_cast<const uint4*>(rec + MRC_REC_PC), nullptr, reinterpret_cast<const int32_t*>(rec + MRC_REC_MX),
                    grp * MRC_GROUP_CHUNKS, joint ? MRC_NCHUNK : MRC_GROUP_CHUNKS, B0, min_nl, lane, dummy, dummy);
                R = reservoir_after(gt, B0, frac, cp.no_huff, nullptr, nullptr);
                ++n_slow;
            }
            if (grp == 0) R1 = R;
        }
        if (lane == 0) io.rsv[blk0 + b - g0] = make_int4(R0, R1, R, 0);
        __syncwarp();
        if (lane == 0 && b + TAB_STAGES < b_hi) issue(b + TAB_STAGES, s);
        if (++s == TAB_STAGES) { s = 0; parity ^= 1u; }
    }
    if (lane == 0) {
        if (iter_counter) atomicAdd(iter_counter, (unsigned long long)n_slow);
        io.clip_res[clip] = R;
        if (b_hi == nblk_clip && reservoir_out) reservoir_out[clip] = R;
    }
}

// ---- segment-composed reservoir maps ---------------------------------------------------------------------------
// The per-block maps R_in -> R_out compose.  segment_kernel (parallel: one CTA per segment of `S` consecutive blocks of
// one clip, one thread per R_in of the tabulated range) chases every R_in through the segment's blocks: a block's
// table while the running value stays inside the range, the closed form "every token granted" above it, and for the
// rare value that is neither a complete walk of that block's record, done by the whole warp once per distinct value
// (the maps contract: after a block or two the thousand trajectories of a segment have merged into a handful; where
// they have not -- digital silence right at the start of a segment shifts every value by the same amount -- the
// values past the sixteenth distinct one are given up: entry RIN_NONE).  It also composes the closed forms:
// R_out = R_in + delta for R_in >= theta, when every block of the segment grants everything along the way (silence:
// the reservoir grows by thousands of bits per block, far outside any table), and lists up to sixteen distinct results
// that lie outside the tabulated range ("exits").
// extras_kernel (parallel: one warp per exit) follows each exit through the next segments, block by block, until it
// is back inside the range, and leaves (R_in -> R_out) pairs with the segments it crosses: these are exactly the
// out-of-range values the real reservoir can enter those segments with.
// chain_seg_kernel (serial, one warp per clip) then takes ONE step per segment -- table entry, closed form or pair --
// and walks block by block only through segments it can enter no other way (those that straddle two clips, or follow a
// wave boundary inside a silent passage).  It records the reservoir at the start of every segment it stepped over;
// expand_kernel (parallel, one warp per segment) replays those segments block by block for finish_kernel.
constexpr int SEG_THREADS = 384;
constexpr int SEG_EPT = 6;                       // table entries per thread: ntab <= 2304
constexpr int RIN_NONE = (int)0x80000000;
constexpr int SEG_MAX_WALKS = 16;                 // distinct out-of-table values a warp follows per block and group
constexpr int SEG_EXITS = 16;                    // distinct out-of-range results kept per segment
constexpr int SEGX_W = 3 * SEG_EXITS;            // per segment: exits, pair inputs, pair outputs
constexpr int EXTRA_HOPS = 6;

struct BlkStep { int R0, R1, R; };

// One block, every lane with the same R: group by group through table / closed form / complete walk.
__device__ __forceinline__ BlkStep step_block(const CodecParams& cp, const unsigned char* __restrict__ rec,
                                              const int* __restrict__ T0, bool joint, int R, int r_lo, int ntab,
                                              int tabw, int lane, unsigned& n_slow) {
    const int32_t* mx = reinterpret_cast<const int32_t*>(rec + MRC_REC_MX);
    const int K = __ldg(mx + MRC_MX_K);
    BlkStep o;
    o.R0 = R;
    o.R1 = R;
    const int ngroups = joint ? 1 : 2;
    unsigned dummy = 0;
    for (int grp = 0; grp < ngroups; ++grp) {
        const int* T = T0 + grp * tabw;
        const int B0 = K + R, idx = R - r_lo;
        if ((unsigned)idx < (unsigned)ntab) R = __ldg(T + idx);
        else if (B0 >= __ldg(T + ntab + 1)) R = B0 + __ldg(T + ntab);
        else {
            const GroupTotals gt = walk_group<false>(
                reinterpret_cast<const uint32_t*>(rec + MRC_REC_TN), reinterpret_cast<const uint32_t*>(rec + MRC_REC_CP),
                reinterpret_cast<const uint4*>(rec + MRC_REC_PC), nullptr, mx, grp * MRC_GROUP_CHUNKS,
                joint ? MRC_NCHUNK : MRC_GROUP_CHUNKS, B0, __ldg(mx + MRC_MX_MINNL), lane, dummy, dummy);
            R = reservoir_after(gt, B0, __ldg(mx + MRC_MX_FRAC), cp.no_huff, nullptr, nullptr);
            ++n_slow;
        }
        if (grp == 0) o.R1 = R;
    }
    o.R = R;
    return o;
}

// clip that holds global block g
__device__ __forceinline__ int clip_of_block(const ClipMap& cm, int g) {
    int lo = 0, hi = cm.n_clips;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cm.clip_blk0[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// insert v into a small set of slots (RIN_NONE = empty); returns the slot or -1 when the set is full
__device__ __forceinline__ int slot_insert(int* slots, int n, int v) {
    for (int i = 0; i < n; ++i) {
        const int old = atomicCAS(slots + i, RIN_NONE, v);
        if (old == RIN_NONE || old == v) return i;
    }
    return -1;
}

// comp: [nseg][segw] ints: entries 0..ntab-1 the composed map (RIN_NONE: not followed), [ntab] delta, [ntab+1] theta
// (INT_MAX: no closed form), [ntab+2] 1 if the segment lies inside one clip (else the row is not to be used).
// segx: [nseg][SEGX_W], reset here; rin[seg] is reset to RIN_NONE.
__global__ void __launch_bounds__(SEG_THREADS)
segment_kernel(CodecParams cp, ClipMap cm, int g0, int nblk_wave, int S, ChainIO io, int r_lo, int ntab, int tabw,
               const int* __restrict__ tab, int segw, int* __restrict__ comp, int* __restrict__ segx,
               int* __restrict__ rin) {
    __shared__ int s_pure, s_flush_lb;
    const int tid = threadIdx.x, lane = tid & 31;
    const int seg = blockIdx.x;
    const int lb0 = seg * S, lb1 = min(lb0 + S, nblk_wave);
    int* row = comp + (size_t)seg * segw;
    int* sx = segx + (size_t)seg * SEGX_W;
    if (tid < SEGX_W) sx[tid] = RIN_NONE;
    if (tid == 0) {
        rin[seg] = RIN_NONE;
        const int ga = g0 + lb0, gb = g0 + lb1 - 1;
        const int c = clip_of_block(cm, ga);
        s_pure = gb < cm.clip_blk0[c + 1];
        // wave-local index of the clip's last block (the non-joint Close() flush block) if it falls in this segment
        s_flush_lb = (cp.flush_nonjoint && gb == cm.clip_blk0[c + 1] - 1) ? lb1 - 1 : -1;
    }
    __syncthreads();
    if (!s_pure) {
        if (tid == 0) { row[ntab] = 0; row[ntab + 1] = 0x7fffffff; row[ntab + 2] = 0; }
        return;
    }
    int R[SEG_EPT];
    bool act[SEG_EPT];
#pragma unroll
    for (int e = 0; e < SEG_EPT; ++e) {
        const int idx = tid + e * SEG_THREADS;
        act[e] = idx < ntab;
        R[e] = r_lo + (act[e] ? idx : 0);
    }
    long long P = 0, theta = -(1ll << 40);           // thread 0: the composed closed form
    unsigned dummy = 0;
    for (int lb = lb0; lb < lb1; ++lb) {
        const unsigned char* rec = io.rec + (size_t)lb * MRC_REC_BYTES;
        const int32_t* mx = reinterpret_cast<const int32_t*>(rec + MRC_REC_MX);
        const int K = __ldg(mx + MRC_MX_K), frac = __ldg(mx + MRC_MX_FRAC), min_nl = __ldg(mx + MRC_MX_MINNL);
        const bool joint = cp.joint && lb != s_flush_lb;
        const int ngroups = joint ? 1 : 2, nck = joint ? MRC_NCHUNK : MRC_GROUP_CHUNKS;
        for (int grp = 0; grp < ngroups; ++grp) {
            const int* T = tab + (size_t)lb * (2 * tabw) + grp * tabw;
            const int c_all = __ldg(T + ntab), thr = __ldg(T + ntab + 1);
            if (tid == 0) {
                theta = max(theta, (long long)thr - K - P);
                P += (long long)K + c_all;
            }
            bool need[SEG_EPT];
#pragma unroll
            for (int e = 0; e < SEG_EPT; ++e) {
                const int idx = R[e] - r_lo, B0 = K + R[e];
                need[e] = false;
                if (R[e] == RIN_NONE) continue;
                if ((unsigned)idx < (unsigned)ntab) R[e] = __ldg(T + idx);
                else if (B0 >= thr) R[e] = B0 + c_all;
                else need[e] = act[e];
            }
            for (int walks = 0;; ++walks) {          // warp-uniform: one complete walk per distinct value of the warp
                unsigned msel = 0u;                  // lanes of the first entry that still needs a walk, and its value
                int Rsel = 0;
#pragma unroll
                for (int e = SEG_EPT - 1; e >= 0; --e) {
                    const unsigned me = __ballot_sync(0xffffffffu, need[e]);
                    if (me) { msel = me; Rsel = R[e]; }
                }
                if (!msel) break;
                if (walks >= SEG_MAX_WALKS) {        // too many distinct values: these trajectories are given up
#pragma unroll
                    for (int e = 0; e < SEG_EPT; ++e)
                        if (need[e]) { R[e] = RIN_NONE; need[e] = false; }
                    break;
                }
                const int Rl = __shfl_sync(0xffffffffu, Rsel, __ffs(msel) - 1);
                const GroupTotals gt = walk_group<false>(
                    reinterpret_cast<const uint32_t*>(rec + MRC_REC_TN), reinterpret_cast<const uint32_t*>(rec + MRC_REC_CP),
                    reinterpret_cast<const uint4*>(rec + MRC_REC_PC), nullptr, mx, grp * MRC_GROUP_CHUNKS, nck, K + Rl,
                    min_nl, lane, dummy, dummy);
                const int Rn = reservoir_after(gt, K + Rl, frac, cp.no_huff, nullptr, nullptr);
#pragma unroll
                for (int e = 0; e < SEG_EPT; ++e)
                    if (need[e] && R[e] == Rl) { R[e] = Rn; need[e] = false; }
            }
        }
    }
    MRC_ASSERT(ntab + 3 <= segw && ntab <= SEG_EPT * SEG_THREADS);
#pragma unroll
    for (int e = 0; e < SEG_EPT; ++e) {
        if (act[e]) row[tid + e * SEG_THREADS] = R[e];
        // results outside the tabulated range: what the next segment can be entered with (distinct values only)
        const bool out = act[e] && R[e] != RIN_NONE && (unsigned)(R[e] - r_lo) >= (unsigned)ntab;
        const unsigned grp_mask = __match_any_sync(0xffffffffu, out ? R[e] : RIN_NONE);
        if (out && (__ffs(grp_mask) - 1) == lane) slot_insert(sx, SEG_EXITS, R[e]);
    }
    if (tid == 0) {
        const bool ok = theta < 0x7fffffffll && P > -0x7fffffffll && P < 0x7fffffffll;
        row[ntab] = ok ? (int)P : 0;
        row[ntab + 1] = ok ? (int)max(theta, -0x7fffffffll) : 0x7fffffff;
        row[ntab + 2] = 1;
    }
}

// one warp per (segment, exit): follow the exit value through the following segments of the same clip
__global__ void __launch_bounds__(SEG_EXITS * 32)
extras_kernel(CodecParams cp, ClipMap cm, int g0, int nblk_wave, int S, int nseg, ChainIO io, int r_lo, int ntab,
              int tabw, const int* __restrict__ tab, int segw, const int* __restrict__ comp, int* segx) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int seg0 = blockIdx.x;
    int R = segx[(size_t)seg0 * SEGX_W + w];         // written by segment_kernel (an earlier launch)
    if (R == RIN_NONE) return;
    const int clip = clip_of_block(cm, g0 + seg0 * S);
    const int clip_end = cm.clip_blk0[clip + 1] - g0;                  // wave-local end of the clip
    unsigned n_slow = 0;
    for (int hop = 1; hop <= EXTRA_HOPS; ++hop) {
        const int seg = seg0 + hop;
        if (seg >= nseg) return;
        const int lb0 = seg * S, lb1 = min(lb0 + S, nblk_wave);
        if (lb1 > clip_end) return;                  // the next segment is not wholly inside this clip
        const int* row = comp + (size_t)seg * segw;
        int Rn;
        if (R >= row[ntab + 1]) Rn = R + row[ntab];  // closed form: nothing to leave behind
        else {
            const int flush_lb = (cp.flush_nonjoint && lb1 == clip_end) ? lb1 - 1 : -1;
            Rn = R;
            for (int lb = lb0; lb < lb1; ++lb)
                Rn = step_block(cp, io.rec + (size_t)lb * MRC_REC_BYTES, tab + (size_t)lb * (2 * tabw),
                                cp.joint && lb != flush_lb, Rn, r_lo, ntab, tabw, lane, n_slow).R;
            if (lane == 0) {
                int* sx = segx + (size_t)seg * SEGX_W;
                const int slot = slot_insert(sx + SEG_EXITS, SEG_EXITS, R);
                if (slot >= 0) sx[2 * SEG_EXITS + slot] = Rn;    // same input -> same output: concurrent writers agree
            }
        }
        R = Rn;
        if ((unsigned)(R - r_lo) < (unsigned)ntab) return;            // back inside the range: the tables take over
    }
}

// The serial pass.  Everything it reads arrives through TMA bulk copies: the composed rows of the next CS_STAGES
// segments sit in a ring of shared-memory stages (a step is then a shared-memory look-up, not a trip to L2), and when a
// segment has to be walked block by block the per-block tables of CS_FB consecutive blocks are fetched with ONE bulk copy
// (they are contiguous) while the lanes fetch the blocks' budgets in parallel.  Only a complete walk (a value outside a
// block's table that does not grant everything) reads its record from global memory.
constexpr int CS_STAGES = 10;
constexpr int CS_FB = 16;

__global__ void __launch_bounds__(32)
chain_seg_kernel(CodecParams cp, ClipMap cm, int c0, int g0, int nblk_wave, int S, ChainIO io, int r_lo, int ntab,
                 int tabw, const int* __restrict__ tab, int segw, const int* __restrict__ comp,
                 const int* __restrict__ segx, int* __restrict__ rin, int fb_blocks,
                 const int32_t* __restrict__ reservoir_in, int32_t* __restrict__ reservoir_out,
                 unsigned long long* __restrict__ iter_counter, int* __restrict__ bound_rec,
                 const int* __restrict__ bound_ref) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[CS_STAGES + 1];
    const int lane = threadIdx.x;
    const int clip = c0 + blockIdx.x;
    const int blk0 = cm.clip_blk0[clip], nblk_clip = cm.clip_blk0[clip + 1] - blk0;
    const int b_lo = max(blk0, g0) - blk0, b_hi = min(blk0 + nblk_clip, g0 + nblk_wave) - blk0;
    if (b_hi <= b_lo) return;
    const int lb_lo = blk0 + b_lo - g0, lb_hi = blk0 + b_hi - g0;      // wave-local
    const int flush_lb = (cp.flush_nonjoint && b_hi == nblk_clip) ? lb_hi - 1 : -1;
    // segments that lie wholly inside this clip's blocks of the wave: [seg_first, seg_lim)
    const int seg_first = (lb_lo + S - 1) / S;
    const int seg_lim = (lb_hi == nblk_wave) ? (nblk_wave + S - 1) / S : lb_hi / S;
    const unsigned row_bytes = (unsigned)segw * 4u, sx_bytes = (unsigned)SEGX_W * 4u, stage_bytes = row_bytes + sx_bytes;
    const unsigned tb_bytes = (unsigned)(2 * tabw) * 4u;               // both groups' tables of one block
    unsigned char* const fb = smem_raw + (size_t)CS_STAGES * stage_bytes;
    if (lane == 0) {
        for (int i = 0; i <= CS_STAGES; ++i) mbar_init(&s_bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int seg) {               // lane 0 only
        const int st = (seg - seg_first) % CS_STAGES;
        unsigned char* dst = smem_raw + (size_t)st * stage_bytes;
        mbar_expect_tx(&s_bar[st], stage_bytes);
        bulk_g2s(dst, comp + (size_t)seg * segw, row_bytes, &s_bar[st]);
        bulk_g2s(dst + row_bytes, segx + (size_t)seg * SEGX_W, sx_bytes, &s_bar[st]);
    };
    if (lane == 0)
        for (int i = 0; i < CS_STAGES && seg_first + i < seg_lim; ++i) issue(seg_first + i);
    int R = (b_lo == 0) ? (reservoir_in ? reservoir_in[clip] : 0) : io.clip_res[clip];
    unsigned n_slow = 0, n_blk = 0, n_pair = 0, n_closed = 0, n_esc = 0, n_nopair = 0, n_impure = 0, fb_parity = 0, dummy = 0;
    int lb = lb_lo;
    const int nseg_all = (nblk_wave + S - 1) / S;
    bool merged = false;
    while (lb < lb_hi) {
        int end;
        const int seg = lb / S;
        // A shard that walks ahead of its reservoir (bound_rec: from a guessed value) leaves the reservoir at every segment
        // boundary; the walk from the true value (bound_ref) stops where it meets that trajectory: from a common value on
        // the two walks are the same, and everything the guessed one wrote beyond that point stands.
        if (lb % S == 0) {
            if (bound_ref && bound_ref[seg] == R) { merged = true; break; }
            if (bound_rec && lane == 0) bound_rec[seg] = R;
        }
        if (lb % S == 0 && seg >= seg_first && seg < seg_lim) {
            const int k = seg - seg_first, st = k % CS_STAGES;
            mbar_wait(&s_bar[st], (unsigned)((k / CS_STAGES) & 1));
            const int* row = reinterpret_cast<const int*>(smem_raw + (size_t)st * stage_bytes);
            const int* sx = row + segw;
            const int seg_end = min(lb + S, nblk_wave);
            const int idx = R - r_lo;
            int Rn = RIN_NONE;
            const bool pure = row[ntab + 2] != 0;
            if (pure) {
                if ((unsigned)idx < (unsigned)ntab) Rn = row[idx];
                else if (R >= row[ntab + 1]) { Rn = R + row[ntab]; ++n_closed; }
                else {
                    const int v = (lane < SEG_EXITS) ? sx[SEG_EXITS + lane] : RIN_NONE;       // one pair per lane
                    const unsigned hit = __ballot_sync(0xffffffffu, v == R);
                    if (hit) { Rn = sx[2 * SEG_EXITS + __ffs(hit) - 1]; ++n_pair; }
                }
            }
            __syncwarp();                            // every lane is done with the stage before it is refilled
            if (lane == 0 && seg + CS_STAGES < seg_lim) issue(seg + CS_STAGES);
            if (Rn != RIN_NONE) {
                if (lane == 0) rin[seg] = R;
                R = Rn;
                lb = seg_end;
                continue;
            }
            if (!pure) ++n_impure;
            else if ((unsigned)idx < (unsigned)ntab) ++n_esc;
            else ++n_nopair;
            if (lane == 0) rin[seg] = RIN_NONE;      // (a walk from a guessed reservoir may have stepped over it)
            end = seg_end;
        } else {
            end = min((seg + 1) * S, lb_hi);         // a piece of a segment shared with another clip or wave edge
        }
        // block by block through [lb, end), fb_blocks at a time
        for (int q0 = lb; q0 < end; q0 += fb_blocks) {
            const int nq = min(fb_blocks, end - q0);
            if (lane == 0) {
                mbar_expect_tx(&s_bar[CS_STAGES], (unsigned)nq * tb_bytes);
                bulk_g2s(fb, tab + (size_t)q0 * (2 * tabw), (unsigned)nq * tb_bytes, &s_bar[CS_STAGES]);
            }
            int hK = 0, hF = 0, hM = 0;              // lane i: budget, fraction flag, narrowest band of block q0 + i
            if (lane < nq) {
                const int32_t* mx = reinterpret_cast<const int32_t*>(io.rec + (size_t)(q0 + lane) * MRC_REC_BYTES + MRC_REC_MX);
                hK = __ldg(mx + MRC_MX_K); hF = __ldg(mx + MRC_MX_FRAC); hM = __ldg(mx + MRC_MX_MINNL);
            }
            mbar_wait(&s_bar[CS_STAGES], fb_parity);
            fb_parity ^= 1u;
            for (int i = 0; i < nq; ++i) {
                const int q = q0 + i;
                const int K = __shfl_sync(0xffffffffu, hK, i);
                const bool joint = cp.joint && q != flush_lb;
                const int* T0 = reinterpret_cast<const int*>(fb + (size_t)i * tb_bytes);
                const int R0 = R;
                int R1 = R;
                for (int grp = 0; grp < (joint ? 1 : 2); ++grp) {
                    const int* T = T0 + grp * tabw;
                    const int B0 = K + R, idx = R - r_lo;
                    if ((unsigned)idx < (unsigned)ntab) R = T[idx];
                    else if (B0 >= T[ntab + 1]) R = B0 + T[ntab];
                    else {
                        const unsigned char* rec = io.rec + (size_t)q * MRC_REC_BYTES;
                        const GroupTotals gt = walk_group<false>(
                            reinterpret_cast<const uint32_t*>(rec + MRC_REC_TN), reinterpret_cast<const uint32_t*>(rec + MRC_REC_CP),
                            reinterpret_cast<const uint4*>(rec + MRC_REC_PC), nullptr, reinterpret_cast<const int32_t*>(rec + MRC_REC_MX),
                            grp * MRC_GROUP_CHUNKS, joint ? MRC_NCHUNK : MRC_GROUP_CHUNKS, B0, __shfl_sync(0xffffffffu, hM, i),
                            lane, dummy, dummy);
                        R = reservoir_after(gt, B0, __shfl_sync(0xffffffffu, hF, i), cp.no_huff, nullptr, nullptr);
                        ++n_slow;
                    }
                    if (grp == 0) R1 = R;
                }
                if (lane == 0) io.rsv[q] = make_int4(R0, R1, R, 0);
            }
            n_blk += (unsigned)nq;
            __syncwarp();                            // before the next batch overwrites the tables
        }
        lb = end;
    }
    if (merged) {
        // rows still on their way into this CTA's shared memory must land before it is given up
        const int sg0 = max(lb / S, seg_first), sg1 = min(seg_lim, lb / S + CS_STAGES);
        for (int sg = sg0; sg < sg1; ++sg) {
            const int k = sg - seg_first;
            mbar_wait(&s_bar[k % CS_STAGES], (unsigned)((k / CS_STAGES) & 1));
        }
    }
    if (lane == 0) {
        if (iter_counter) {
            atomicAdd(iter_counter, (unsigned long long)n_slow);           // complete walks taken by the serial pass
            atomicAdd(iter_counter + 1, (unsigned long long)n_blk);        // blocks it stepped through one by one
            atomicAdd(iter_counter + 2, (unsigned long long)n_pair);       // segments entered through an (exit -> result) pair
            atomicAdd(iter_counter + 3, (unsigned long long)n_closed);     // segments stepped by the closed form
            atomicAdd(iter_counter + 4, (unsigned long long)n_esc);        // entered inside the range, trajectory not followed
            atomicAdd(iter_counter + 5, (unsigned long long)n_nopair);     // entered outside the range, value not anticipated
            atomicAdd(iter_counter + 6, (unsigned long long)n_impure);
        }
        if (merged) R = bound_ref[nseg_all];          // where the guessed walk ended
        if (bound_rec) bound_rec[nseg_all] = R;
        io.clip_res[clip] = R;
        if (b_hi == nblk_clip && reservoir_out) reservoir_out[clip] = R;
        if (bound_ref && iter_counter) atomicAdd(iter_counter + 7, (unsigned long long)(merged ? lb / S + 1 : 0));
    }
}

constexpr int EXP_WARPS = 4;

__global__ void __launch_bounds__(EXP_WARPS * 32)
expand_kernel(CodecParams cp, ClipMap cm, int g0, int nblk_wave, int S, int nseg, ChainIO io, int r_lo, int ntab,
              int tabw, const int* __restrict__ tab, const int* __restrict__ rin) {
    const int lane = threadIdx.x & 31;
    const int seg = blockIdx.x * EXP_WARPS + (threadIdx.x >> 5);
    if (seg >= nseg) return;
    int R = rin[seg];
    if (R == RIN_NONE) return;                       // the serial pass went through this segment block by block
    const int lb0 = seg * S, lb1 = min(lb0 + S, nblk_wave);
    int flush_lb = -1;
    if (cp.flush_nonjoint) {                         // the segment lies inside one clip: is its last block that clip's last?
        const int gb = g0 + lb1 - 1;
        if (gb == cm.clip_blk0[clip_of_block(cm, gb) + 1] - 1) flush_lb = lb1 - 1;
    }
    unsigned n_slow = 0;
    for (int lb = lb0; lb < lb1; ++lb) {
        const BlkStep o = step_block(cp, io.rec + (size_t)lb * MRC_REC_BYTES, tab + (size_t)lb * (2 * tabw),
                                     cp.joint && lb != flush_lb, R, r_lo, ntab, tabw, lane, n_slow);
        if (lane == 0) io.rsv[lb] = make_int4(o.R0, o.R1, o.R, 0);
        R = o.R;
    }
}

constexpr int FIN_WARPS = 8;

__global__ void __launch_bounds__(FIN_WARPS * 32)
finish_kernel(CodecParams cp, ClipMap cm, int g0, int nblk, ChainIO io) {
    const int lane = threadIdx.x & 31;
    const int lb = blockIdx.x * FIN_WARPS + (threadIdx.x >> 5);
    if (lb >= nblk) return;
    const int g = g0 + lb;
    int lo = 0, hi = cm.n_clips;                     // clip of this block (every lane: uniform, cached loads)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (cm.clip_blk0[mid] <= g) lo = mid; else hi = mid;
    }
    const bool joint = cp.joint && !(cp.flush_nonjoint && g == cm.clip_blk0[lo + 1] - 1);
    const unsigned char* rec = io.rec + (size_t)lb * MRC_REC_BYTES;
    const uint32_t* tn = reinterpret_cast<const uint32_t*>(rec + MRC_REC_TN);
    const uint32_t* cpre = reinterpret_cast<const uint32_t*>(rec + MRC_REC_CP);
    const uint4* pc = reinterpret_cast<const uint4*>(rec + MRC_REC_PC);
    const int32_t* mx = reinterpret_cast<const int32_t*>(rec + MRC_REC_MX);
    const uint4* pw = reinterpret_cast<const uint4*>(io.pw + (size_t)lb * MRC_PW_BYTES);
    const int4 rs = io.rsv[lb];
    unsigned gmask = 0u, n_iter = 0u;
    int table[2] = {MRC_NO_TABLE, MRC_NO_TABLE}, wbits[2] = {0, 0};
    const int K = mx[MRC_MX_K], frac = mx[MRC_MX_FRAC], min_nl = mx[MRC_MX_MINNL];
    if (joint) {
        const int B0 = K + rs.x;
        const GroupTotals gt = walk_group<true>(tn, cpre, pc, pw, mx, 0, MRC_NCHUNK, B0, min_nl, lane, gmask, n_iter);
        reservoir_after(gt, B0, frac, cp.no_huff, table, wbits);
    } else {
        int t2[2], w2[2];
        int B0 = K + rs.x;
        GroupTotals gt = walk_group<true>(tn, cpre, pc, pw, mx, 0, MRC_GROUP_CHUNKS, B0, min_nl, lane, gmask, n_iter);
        reservoir_after(gt, B0, frac, cp.no_huff, t2, w2);
        table[0] = t2[0]; wbits[0] = w2[0];
        B0 = K + rs.y;
        gt = walk_group<true>(tn, cpre, pc, pw, mx, MRC_GROUP_CHUNKS, MRC_GROUP_CHUNKS, B0, min_nl, lane, gmask, n_iter);
        reservoir_after(gt, B0, frac, cp.no_huff, t2, w2);
        table[1] = t2[1]; wbits[1] = w2[1];
    }
    io.gmask[(size_t)lb * 32 + lane] = gmask;
    if (lane == 0) {
        const int nb = mx[MRC_MX_NB];
        ChainBlk o;
        for (int ch = 0; ch < 2; ++ch) {
            int bits = 4 + 1 + 1 + nb * (cp.n_mant_size_bits + cp.n_scale_bits) + wbits[ch];
            if (joint) bits += (ch == 0) ? (4 * cp.n_scale_bits + nb) : 0;
            else bits += cp.n_scale_bits;
            o.chunk_off[ch] = 0;                     // filled by offsets_kernel
            o.chunk_bytes[ch] = (unsigned)((bits + 7) >> 3);
            o.table[ch] = (unsigned char)table[ch];
        }
        o.reservoir = rs.z;
        o.pad[0] = o.pad[1] = 0;
        io.cblk[lb] = o;
    }
}

// one warp per clip: running byte offset over the clip's blocks inside this wave (carried across waves in clip_run)
__global__ void __launch_bounds__(32)
offsets_kernel(CodecParams cp, ClipMap cm, int c0, int g0, int nblk_wave, ChainIO io) {
    const int lane = threadIdx.x;
    const int clip = c0 + blockIdx.x;
    const int blk0 = cm.clip_blk0[clip], nblk_clip = cm.clip_blk0[clip + 1] - blk0;
    const int b_lo = max(blk0, g0) - blk0, b_hi = min(blk0 + nblk_clip, g0 + nblk_wave) - blk0;
    if (b_hi <= b_lo) return;
    long long running = (b_lo == 0) ? (long long)cp.header_bytes : io.clip_run[clip];
    for (int base = b_lo; base < b_hi; base += 32) {
        const int b = base + lane;
        const bool in = b < b_hi;
        ChainBlk* cb = io.cblk + (blk0 + b - g0);
        const unsigned s0 = in ? cb->chunk_bytes[0] : 0u, s1 = in ? cb->chunk_bytes[1] : 0u;
        const long long sz = in ? (long long)(8u + s0 + s1) : 0ll;
        long long incl = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (in) {
            const long long off = running + incl - sz;
            cb->chunk_off[0] = off;
            cb->chunk_off[1] = off + 4 + s0;
        }
        running += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) {
        io.clip_run[clip] = running;
        if (b_hi == nblk_clip) io.clip_bytes[clip] = running;
    }
}

}  // namespace

void launch_chain(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0, int nblk,
                  ChainIO io, const int32_t* reservoir_in, int32_t* reservoir_out,
                  unsigned long long* iter_counter) {
    if (nclips <= 0 || nblk <= 0) return;
    const size_t smem = (size_t)MRC_CHAIN_STAGES * MRC_REC_BYTES;
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    chain_kernel<<<nclips, 32, smem, st>>>(cp, cm, c0, g0, nblk, io, reservoir_in, reservoir_out,
                                           iter_counter);
}

void launch_finish(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk, ChainIO io) {
    if (nblk <= 0) return;
    finish_kernel<<<(nblk + FIN_WARPS - 1) / FIN_WARPS, FIN_WARPS * 32, 0, st>>>(cp, cm, g0, nblk, io);
}

void launch_offsets(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0, int nblk,
                    ChainIO io) {
    if (nclips <= 0 || nblk <= 0) return;
    offsets_kernel<<<nclips, 32, 0, st>>>(cp, cm, c0, g0, nblk, io);
}

void launch_table(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk,
                  ChainIO io, int r_lo, int ntab, int tabw, int* tab) {
    if (nblk <= 0) return;
    table_kernel<<<nblk, TAB_THREADS, 0, st>>>(cp, cm, g0, io, r_lo, ntab, tabw, tab);
}

void launch_chain_table(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0,
                        int nblk, ChainIO io, int r_lo, int ntab, int tabw, const int* tab,
                        const int32_t* reservoir_in, int32_t* reservoir_out, unsigned long long* iter_counter) {
    if (nclips <= 0 || nblk <= 0) return;
    const size_t smem = (size_t)TAB_STAGES * ((size_t)2 * tabw * 4 + MRC_REC_BYTES);
    cudaFuncSetAttribute(chain_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    chain_table_kernel<<<nclips, 32, smem, st>>>(cp, cm, c0, g0, nblk, io, r_lo, ntab, tabw, tab,
                                                 reservoir_in, reservoir_out, iter_counter);
}

void launch_segments(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk, int S, ChainIO io,
                     int r_lo, int ntab, int tabw, const int* tab, int segw, int* comp, int* segx, int* rin) {
    if (nblk <= 0) return;
    const int nseg = (nblk + S - 1) / S;
    segment_kernel<<<nseg, SEG_THREADS, 0, st>>>(cp, cm, g0, nblk, S, io, r_lo, ntab, tabw, tab, segw, comp, segx, rin);
    extras_kernel<<<nseg, SEG_EXITS * 32, 0, st>>>(cp, cm, g0, nblk, S, nseg, io, r_lo, ntab, tabw, tab, segw, comp, segx);
}

void launch_chain_seg(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0, int nblk,
                      int S, ChainIO io, int r_lo, int ntab, int tabw, const int* tab, int segw, const int* comp,
                      const int* segx, int* rin, const int32_t* reservoir_in, int32_t* reservoir_out,
                      unsigned long long* iter_counter, int* bound_rec, const int* bound_ref) {
    if (nclips <= 0 || nblk <= 0) return;
    // shared memory: the ring of composed rows + as many per-block tables as fit (at most CS_FB, at least one)
    const size_t ring = (size_t)CS_STAGES * ((size_t)segw * 4 + SEGX_W * 4), tb = (size_t)2 * tabw * 4;
    int fbn = (int)std::min<size_t>(CS_FB, (200 * 1024 - ring) / tb);
    if (fbn < 1) fbn = 1;
    const size_t smem = ring + (size_t)fbn * tb;
    cudaFuncSetAttribute(chain_seg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    chain_seg_kernel<<<nclips, 32, smem, st>>>(cp, cm, c0, g0, nblk, S, io, r_lo, ntab, tabw, tab, segw, comp, segx, rin, fbn,
                                               reservoir_in, reservoir_out, iter_counter, bound_rec, bound_ref);
}

void launch_expand(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk, int S, ChainIO io, int r_lo,
                   int ntab, int tabw, const int* tab, const int* rin) {
    if (nblk <= 0) return;
    const int nseg = (nblk + S - 1) / S;
    expand_kernel<<<(nseg + EXP_WARPS - 1) / EXP_WARPS, EXP_WARPS * 32, 0, st>>>(cp, cm, g0, nblk, S, nseg, io, r_lo, ntab,
                                                                                 tabw, tab, rin);
}

int segment_max_ntab() { return SEG_THREADS * SEG_EPT; }
int segment_aux_width() { return SEGX_W; }
