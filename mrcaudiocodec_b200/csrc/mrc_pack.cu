// mrc_pack.cu -- K4: quantisation at the chosen allocation + bit-packing of the channel chunks, one CTA per
// (block, channel).
//   allocation                    the chain kernel's grant masks over the presorted tokens (bitalloc.py:106-155)
//   scale factors + mantissas     codecThem.py:336-350 / :509-559, quantize.py:114-146, :294-322
//   chunk layout and size         pacfileThem.py:651-789 (non-joint) / :825-970 (joint); SURVEY.md Appendix B
//   MSB-first bit writer          bitpack.py:36-101
//   file header                   pacfileThem.py:586-613 (numSamples quirk Q9)
// Code lengths of the L mantissas go through a block-wide exclusive scan; every symbol is then OR-ed into a
// zeroed shared-memory bit buffer at its own offset and the finished chunk is streamed out with its <L nBytes
// prefix.
#include "mrc_internal.cuh"
#include "mrc_math.cuh"

namespace {

constexpr int PT = 256;

template <typename T>
__global__ void __launch_bounds__(PT, 8)
pack_kernel(DevTables<T> tb, CodecParams cp, const HuffDev* __restrict__ huff, ClipMap cm, int g0, Handoff<T> ho,
            ChainIO io, PackTaps taps, const int64_t* __restrict__ clip_base, uint8_t* __restrict__ out,
            long long out_cap, const uint8_t* __restrict__ header_template, int* overflow_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* bitbuf = reinterpret_cast<uint32_t*>(smem_raw);
    const int L = cp.L, nb = cp.nb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t lb = cm.list ? (size_t)cm.list[blockIdx.x >> 1] : (size_t)(blockIdx.x >> 1);
    const int ch = blockIdx.x & 1;
    const int g = g0 + (int)lb;
    const uint8_t* __restrict__ line2band = tb.line2band;
    const int* __restrict__ band_lo = tb.band_lo;

    __shared__ int s_clip, s_b, s_nblk;
    __shared__ int s_alloc[MRC_BSTRIDE], s_sf[MRC_BSTRIDE];
    __shared__ int s_wsum[PT / 32];
    __shared__ HuffDev s_h;
    if (tid == 0) {
        int lo = 0, hi = cm.n_clips;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cm.clip_blk0[mid] <= g) lo = mid; else hi = mid;
        }
        s_clip = lo;
        s_b = g - cm.clip_blk0[lo];
        s_nblk = cm.clip_blk0[lo + 1] - cm.clip_blk0[lo];
    }
    if (tid < MRC_BSTRIDE) s_alloc[tid] = 0;
    for (int i = tid; i < (int)(sizeof(HuffDev) / 4); i += PT)
        reinterpret_cast<uint32_t*>(&s_h)[i] = reinterpret_cast<const uint32_t*>(huff)[i];
    const ChainBlk cb = io.cblk[lb];
    const int nbytes = (int)cb.chunk_bytes[ch];
    const int nwords = (nbytes + 3) / 4 + 2;
    for (int i = tid; i < nwords; i += PT) bitbuf[i] = 0u;
    __syncthreads();

    // allocation of this channel's bands = 2 + highest granted level (tokens of a band are granted in level order)
    {
        const uint32_t* tn = reinterpret_cast<const uint32_t*>(io.rec + lb * (size_t)MRC_REC_BYTES + MRC_REC_TN);
        const uint32_t* gm = io.gmask + lb * 32;
        for (int j = tid; j < MRC_NSLOT; j += PT) {
            const uint32_t t = tn[j];
            if (t == 0xffffffffu || !((gm[j >> 5] >> (j & 31)) & 1u)) continue;
            const int bb = (int)(t & 0xff), lvl = (int)((t >> 8) & 0xff);
            if ((bb >= nb) == (ch != 0)) atomicMax(&s_alloc[bb - ch * nb], lvl + 2);
        }
    }
    __syncthreads();
    if (tid < nb)
        s_sf[tid] = scale_factor_of((double)ho.bandmax[(lb * 2 + ch) * MRC_BSTRIDE + tid], cp.n_scale_bits, s_alloc[tid]);
    __syncthreads();
    if (taps.mant != nullptr)
        for (int k = L + tid; k < cp.Lmax; k += PT) taps.mant[(lb * 2 + ch) * cp.Lmax + k] = 0;
    if (taps.alloc != nullptr && tid < MRC_BSTRIDE) {
        taps.alloc[(lb * 2 + ch) * MRC_BSTRIDE + tid] = (uint8_t)(tid < nb ? s_alloc[tid] : 0);
        taps.sf[(lb * 2 + ch) * MRC_BSTRIDE + tid] = (uint8_t)(tid < nb ? s_sf[tid] : 0);
    }

    const bool joint = cp.joint && !(cp.flush_nonjoint && s_b == s_nblk - 1);
    const int table = cb.table[ch];
    const int band_hdr = cp.n_mant_size_bits + cp.n_scale_bits;
    int hdr_bits = 4 + 1 + 1;
    if (joint) hdr_bits += (ch == 0) ? 4 * cp.n_scale_bits + nb : 0;
    else hdr_bits += cp.n_scale_bits;

    auto put = [&](int pos, uint32_t v, int n) {       // n <= 32 bits of v at bit position pos, MSB first
        if (n <= 0) return;
        const int w = pos >> 5, off = pos & 31;
        MRC_ASSERT(pos >= 0 && n <= 32 && w + 1 < nwords);
        const unsigned long long x = (unsigned long long)v << (64 - off - n);
        const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
        if (hi) atomicOr(&bitbuf[w], hi);
        if (lo) atomicOr(&bitbuf[w + 1], lo);
    };

    // per-line symbols: LPT consecutive lines per thread
    const int LPT = (L + PT - 1) / PT;           // 1 .. 8 (L = 576 -> 3: threads past the last line idle)
    const T* __restrict__ lines = ho.lines + lb * 2 * cp.Lmax + (size_t)ch * L;
    uint32_t sym[8];
    int slen[8];
    int local = 0;
    for (int i = 0; i < LPT; ++i) {
        const int k = tid * LPT + i;
        sym[i] = 0u;
        slen[i] = 0;
        if (k >= L) continue;
        const int bd = line2band[k];
        const int Rb = s_alloc[bd];
        int n = 0, m = 0;
        uint32_t v = 0;
        if (Rb) {
            m = mantissa_of((double)lines[k], s_sf[bd], cp.n_scale_bits, Rb);
            if (table == MRC_NO_TABLE) { n = Rb; v = (uint32_t)m; }
            else {
                const int len = (m < MRC_HUFF_LUT) ? s_h.len[table][m] : 0;
                if (len && m != s_h.esc[table]) { n = len; v = s_h.code[table][m]; }
                else { n = s_h.esc_len[table] + Rb; v = ((uint32_t)s_h.esc_code[table] << Rb) | (uint32_t)m; }
            }
        }
        if (taps.mant != nullptr) taps.mant[(lb * 2 + ch) * cp.Lmax + k] = (uint16_t)m;
        sym[i] = v;
        slen[i] = n;
        local += n;
    }
    // block-wide exclusive scan of `local`
    int incl = local;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += s_wsum[w];
    int pos = woff + incl - local;               // mantissa bits before this thread's first line
    for (int i = 0; i < LPT; ++i) {
        const int k = tid * LPT + i;
        if (k >= L) break;
        const int bd = line2band[k];
        const int p = hdr_bits + band_hdr * (bd + 1) + pos;
        if (k == band_lo[bd]) {                   // first line of a band also writes the band header
            const int a = s_alloc[bd];
            put(p - band_hdr, (uint32_t)(((a ? a - 1 : 0) << cp.n_scale_bits) | s_sf[bd]), band_hdr);
        }
        put(p, sym[i], slen[i]);
        pos += slen[i];
    }
    if (tid == 0) {
        put(0, (uint32_t)table, 4);               // huffTable(4)
        put(4, (uint32_t)(tb.geom & 3), 2);       // blkswA, blkswB: 1 = that window half is short (pacfileThem.py:720-721)
        int p = 6;
        if (joint) {
            if (ch == 0) {
                for (int i = 0; i < 4; ++i) { put(p, ho.ovs[lb * 4 + i], cp.n_scale_bits); p += cp.n_scale_bits; }
                const uint32_t ms = ho.ms[lb];
                for (int bd = 0; bd < nb; ++bd) put(p + bd, (ms >> bd) & 1u, 1);
            }
        } else {
            put(p, ho.ovs[lb * 4 + ch], cp.n_scale_bits);
        }
    }
    __syncthreads();

    if (out == nullptr) return;                   // taps only
    const long long dst_off = clip_base[s_clip] + cb.chunk_off[ch];
    if (dst_off + 4 + nbytes > out_cap) {         // never write past the caller's buffer; the host reports NOSPACE
        if (tid == 0) *overflow_flag = 1;
        return;
    }
    uint8_t* dst = out + dst_off;
    if (tid < 4) dst[tid] = (uint8_t)((uint32_t)nbytes >> (8 * tid));       // <L nBytes
    for (int i = tid; i < nbytes; i += PT) dst[4 + i] = (uint8_t)(bitbuf[i >> 2] >> (24 - 8 * (i & 3)));

    if (s_b == 0 && ch == 0 && cp.header_bytes > 0) {   // first chunk of a clip also writes the file header (a shard that
                                                        // does not start the stream has header_bytes = 0)
        uint8_t* h = out + clip_base[s_clip];
        for (int i = tid; i < cp.header_bytes; i += PT) h[i] = header_template[i];
        __syncthreads();
        if (tid == 0) {
            long long ns = cm.clip_off[s_clip + 1] - cm.clip_off[s_clip];
            if (ns % cp.Lmax == 0) ns += cp.Lmax; // Q9: bumped only when already a multiple
            for (int i = 0; i < 4; ++i) h[10 + i] = (uint8_t)((unsigned long long)ns >> (8 * i));
        }
    }
}

__global__ void clip_scan_kernel(const int64_t* __restrict__ clip_bytes, int64_t* __restrict__ clip_base, int c0,
                                 int n, int64_t* running) {
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = *running;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const long long v = (i < n) ? clip_bytes[c0 + i] : 0;
        long long incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        long long woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_w[w];
        const long long carry = s_carry;
        if (i < n) clip_base[c0 + i] = carry + woff + incl - v;
        __syncthreads();
        if (tid == blockDim.x - 1) s_carry = carry + woff + incl;
        __syncthreads();
    }
    if (tid == 0) { *running = s_carry; clip_base[c0 + n] = s_carry; }
}

}  // namespace

void launch_clip_scan(cudaStream_t st, const int64_t* clip_bytes, int64_t* clip_base, int c0, int n,
                      int64_t* running) {
    clip_scan_kernel<<<1, 256, 0, st>>>(clip_bytes, clip_base, c0, n, running);
}

template <typename T>
void launch_pack(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                 const ClipMap& cm, int g0, int nblk, Handoff<T> ho, ChainIO io, PackTaps taps,
                 const int64_t* clip_base, uint8_t* out, long long out_cap, const uint8_t* header_template,
                 int* overflow_flag) {
    if (nblk <= 0) return;
    const size_t smem = (size_t)cp.Lmax * 25 / 8 + 512;
    pack_kernel<T><<<2 * nblk, PT, smem, st>>>(tb, cp, huff, cm, g0, ho, io, taps, clip_base, out, out_cap,
                                               header_template, overflow_flag);
}

template void launch_pack<double>(cudaStream_t, const DevTables<double>&, const CodecParams&, const HuffDev*,
                                  const ClipMap&, int, int, Handoff<double>, ChainIO, PackTaps, const int64_t*,
                                  uint8_t*, long long, const uint8_t*, int*);
template void launch_pack<float>(cudaStream_t, const DevTables<float>&, const CodecParams&, const HuffDev*,
                                 const ClipMap&, int, int, Handoff<float>, ChainIO, PackTaps, const int64_t*,
                                 uint8_t*, long long, const uint8_t*, int*);
