// mrc_decode.cu -- K5: the decode mirror path.
//   chunk parse                     pacfileThem.py:161-319 (ReadDataBlock) / :321-585 (JointReadDataBlock);
//                                   bit reader bitpack.py:104-170; prefix-code walk :445-480 (here a 9-bit LUT)
//   dequantise                      codecThem.py:30-63 / :65-134 with quantize.py:325-357, :90-111
//   M/S reconstruct                 ms_stereo.py:33-49
//   IMDCT + window                  mdct.py:98-122 (here an L/2-point complex FFT DCT-IV), window.py:104-121
//   overlap-add, first block drop   pacfileThem.py:312-314, :575-580, :1175-1177, :178-185
//   PCM conversion                  pcmfile.py:164-174 with quantize.py:61-87 at 16 bits
// One CTA per block pair (both channels), NT = L/2 threads, L = (a+b)/2 of the pair's geometry (block switching:
// the host reads a and b from the chunk headers and launches each geometry over its own list of pairs).  A malformed
// chunk (reads past its nBytes, block-size bits that differ from the pair's, unknown table id) raises the error flag
// and decodes as silence instead of reading on.
#include "mrc_decode.cuh"
#include "mrc_math.cuh"
#include "mrc_fft.cuh"

namespace {

struct BitReader {
    const uint32_t* w;       // big-endian words in shared memory
    int pos, nbits;
    bool bad;
    __device__ __forceinline__ uint32_t peek(int n) const {     // n in 1..32
        const int i = pos >> 5, off = pos & 31;
        const unsigned long long x = ((unsigned long long)w[i] << 32) | w[i + 1];
        return (uint32_t)((x << off) >> (64 - n));
    }
    __device__ __forceinline__ uint32_t peek_at(int at, int n) const {     // n in 1..32 bits starting at bit `at`
        const int i = at >> 5, off = at & 31;
        const unsigned long long x = ((unsigned long long)w[i] << 32) | w[i + 1];
        return (uint32_t)((x << off) >> (64 - n));
    }
    __device__ __forceinline__ uint32_t read(int n) {
        if (n <= 0) return 0;
        if (pos + n > nbits) { bad = true; pos = nbits; return 0; }
        const uint32_t v = peek(n);
        pos += n;
        return v;
    }
};

// shared-memory working set of one pair
template <typename T>
struct DSmem {
    uint32_t* cw;        // [2][cwords]  chunk payloads as big-endian words
    int* mant;           // [2][L]
    T* lines;            // [2][L]       dequantised MDCT lines (after M/S reconstruct: L, R)
    cpx<T>* buf;         // [L]          two L/2-point FFTs
    T* v;                // [2][L]       DCT-IV outputs
};

template <typename T>
__device__ __forceinline__ DSmem<T> dcarve(unsigned char* raw, int L, int cwords) {
    DSmem<T> s;
    T* p = reinterpret_cast<T*>(raw);
    s.lines = p;  p += 2 * L;
    s.buf = reinterpret_cast<cpx<T>*>(p);  p += 2 * L;
    s.v = p;      p += 2 * L;
    s.mant = reinterpret_cast<int*>(p);
    s.cw = reinterpret_cast<uint32_t*>(s.mant + 2 * L);
    (void)cwords;
    return s;
}

// dequantise + rescale + M/S + IMDCT + window; ints in shared memory.  Writes y[2][2L] to global.
template <typename T, int L_>
__device__ __forceinline__ void synthesize(const DevTables<T>& tb, const CodecParams& cp, DSmem<T>& sm, bool joint,
                                           const int* s_alloc, const int* s_sf, const int* s_ovs, unsigned ms,
                                           T* __restrict__ yout) {
    constexpr int L = L_, Q = L / 2, NT = Q;
    const int tid = threadIdx.x, nb = tb.nb;
    // dequantise (codecThem.py:44-52 / :96-115) -- arithmetic in double in both precisions (a handful of ops)
    for (int i = tid; i < 2 * L; i += NT) {
        const int ch = i / L, k = i - ch * L;
        const int bd = tb.line2band[k];
        const int Rb = s_alloc[ch * MRC_BSTRIDE + bd];
        double x = 0.0;
        if (Rb) {
            x = dequantize_of(sm.mant[i], s_sf[ch * MRC_BSTRIDE + bd], cp.n_scale_bits, Rb);
            const int sc = joint ? s_ovs[((ms >> bd) & 1u) ? 2 + ch : ch] : s_ovs[ch];
            x = x / (double)(1 << sc);
        }
        sm.lines[i] = T(x);
    }
    __syncthreads();
    if (joint) {      // ms_stereo.py:33-49
        for (int k = tid; k < L; k += NT) {
            if ((ms >> tb.line2band[k]) & 1u) {
                const T m = sm.lines[k], s = sm.lines[L + k];
                sm.lines[k] = m + s;
                sm.lines[L + k] = m - s;
            }
        }
        __syncthreads();
    }
    // DCT-IV of both channels: t[n] = (X[2n] + j X[L-1-2n]) * pre[n]; FFT; c = T*post; v[2k]=Re, v[L-1-2k]=-Im
    {
        const int grp = tid / (NT / 2), lt = tid - grp * (NT / 2), gthr = NT / 2;
        const T* X = sm.lines + grp * L;
        cpx<T>* a = sm.buf + grp * Q;
        // Power-of-two L: the analysis kernel's conflict-free transform (mrc_fft.cuh: swizzled work buffer, per-stage
        // twiddle tables, lane-mapped input placement); the stage tables of the L/2-point transform are staged in `v`
        // (written only after the transform).  9 * 2^p lines (transition blocks): the root table, as before.
        if constexpr (FftShape<L>::pow2) {
            constexpr int LOGQ = FftShape<L>::logP - 1;
            cpx<T>* const st = reinterpret_cast<cpx<T>*>(sm.v);
            for (int i = tid; i < fft_stage_entries(LOGQ); i += NT) st[i] = tb.tw_stage[fft_stage_entries(LOGQ + 1) + i];
            for (int i = lt; i < Q; i += gthr) {
                const int n = fft_place_index<LOGQ>(i);
                const T re = X[2 * n], im = X[L - 1 - 2 * n];
                const cpx<T> w = tb.tw_pre[n];
                cpx<T> t;
                t.x = re * w.x - im * w.y;
                t.y = re * w.y + im * w.x;
                a[fft_swz<T>(fft_r4_pos(n, LOGQ))] = t;
            }
            __syncthreads();
            fft_sw<T, LOGQ>(a, lt, gthr, st);
        } else {
            cpx<T>* const tws = reinterpret_cast<cpx<T>*>(sm.v);
            for (int i = tid; i < (1 << (tb.logLtab - 1)); i += NT) tws[i] = tb.tw_fft[i];
            for (int n = lt; n < Q; n += gthr) {
                const T re = X[2 * n], im = X[L - 1 - 2 * n];
                const cpx<T> w = tb.tw_pre[n];
                const int r = fft_pos<Q>(n);
                a[r].x = re * w.x - im * w.y;
                a[r].y = re * w.y + im * w.x;
            }
            __syncthreads();
            fft_any<T, Q>(a, lt, gthr, tws, tb.logLtab, tb.tw9, L, tb.w9);
        }
        T* v = sm.v + grp * L;
        for (int k = lt; k < Q; k += gthr) {
            const cpx<T> w = tb.tw_post[k];
            const cpx<T> t = a[FftShape<L>::pow2 ? fft_swz<T>(k) : k];
            v[2 * k] = t.x * w.x - t.y * w.y;
            v[L - 1 - 2 * k] = -(t.x * w.y + t.y * w.x);
        }
        __syncthreads();
    }
    // unfold (x[n] = 2 v_ext[n + L/2]) and window.  With a != b the phase n0 = (b+1)/2 (mdct.py:103) is the standard
    // one shifted by rot = (a-b)/4 samples: output n is the standard output n - rot, negated where it wraps.
    for (int i = tid; i < 4 * L; i += NT) {
        const int ch = i / (2 * L), n = i - ch * 2 * L;
        const T* v = sm.v + ch * L;
        int m = n;
        T sg = T(2);
        if constexpr (!FftShape<L>::pow2) {
            m = n - tb.rot;
            if (m < 0) { m += 2 * L; sg = T(-2); }
            else if (m >= 2 * L) { m -= 2 * L; sg = T(-2); }
        }
        T x;
        if (m < Q) x = v[m + Q];
        else if (m < 3 * Q) x = -v[3 * Q - 1 - m];
        else x = -v[m - 3 * Q];
        yout[i] = (sg * x) * tb.kbd[n];
    }
}

// ---- parse: ONE WARP PER CHUNK ---------------------------------------------------------------------------------
// The chunk grammar is serial (every field's position depends on the fields before it), so a chunk is one warp's work
// whatever is done; what can be chosen is how many chunks are in flight.  Parsing inside the synthesis CTA (one CTA of
// L/2 threads per pair, two of its sixteen warps parsing while fourteen wait) kept six chunks in flight per SM and the
// whole kernel waited on that latency (profiles/r02t: 69 % of all stall samples at the barrier after the parse).  Here a
// CTA is eight independent warps, each with its chunk staged in its own slice of shared memory: ~56 chunks in flight per
// SM.  Output per pair: mantissa codes (u16), bit allocation and scale factor per band, overall scales, ms_switch and
// flags -- 4.2 KB that the synthesis kernel reads back coalesced.
constexpr int PARSE_WARPS = 8;

struct ParseGeo {
    int nb, geom, L, Lmax, n_scale_bits, n_mant_size_bits, joint, flush_nonjoint;
    uint16_t band_lo[MRC_BSTRIDE], band_n[MRC_BSTRIDE];
};

__global__ void __launch_bounds__(PARSE_WARPS * 32)
parse_kernel(ParseGeo pg, const HuffDev* __restrict__ huff, const HuffDecDev* __restrict__ hdec, DecodeMap dm,
             const uint8_t* __restrict__ pac, int p0, int npairs, ParseOut po, int* error_flag, int cwords) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint16_t s_hlut[MRC_N_HUFF_TABLES][1 << MRC_HUFF_PEEK];     // the 9-bit decode tables (4 KB)
    __shared__ int s_esc[MRC_N_HUFF_TABLES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = pg.nb;
    if (tid < MRC_N_HUFF_TABLES) s_esc[tid] = huff->esc[tid];
    for (int i = tid; i < (MRC_N_HUFF_TABLES << MRC_HUFF_PEEK) / 2; i += PARSE_WARPS * 32)
        reinterpret_cast<uint32_t*>(&s_hlut[0][0])[i] = reinterpret_cast<const uint32_t*>(&hdec->lut[0][0])[i];
    __syncthreads();
    const int ci = blockIdx.x * PARSE_WARPS + warp;            // chunk of this warp: (list entry, channel)
    if (ci >= 2 * npairs) return;
    const int li = ci >> 1, ch = ci & 1;
    const int lp = dm.list ? dm.list[li] : li, p = p0 + lp;
    bool joint;
    {
        int lo = 0, hi = dm.n_clips;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (dm.clip_pair0[mid] <= p) lo = mid; else hi = mid;
        }
        const int last = dm.clip_pair0[lo + 1] - 1;
        joint = pg.joint && !(pg.flush_nonjoint && p == last);
    }
    // stage the chunk payload as big-endian words: aligned 32-bit loads, realigned by the payload's byte offset
    uint32_t* cw = reinterpret_cast<uint32_t*>(smem_raw) + warp * cwords;
    const int nbytes = min((int)dm.chunk_len[2 * p + ch], (cwords - 2) * 4);
    {
        const uint8_t* src = pac + dm.chunk_pos[2 * p + ch];
        const int mis = (int)(reinterpret_cast<size_t>(src) & 3);
        const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src - mis);
        for (int wi = lane; wi < cwords; wi += 32) {
            const int valid = nbytes - 4 * wi;                 // payload bytes from this word on
            uint32_t w = 0;
            if (valid > 0) {
                const uint32_t lo = __ldg(src32 + wi);
                const uint32_t hi = (mis && valid + mis > 4) ? __ldg(src32 + wi + 1) : 0u;
                w = __byte_perm(__funnelshift_r(lo, hi, 8 * mis), 0u, 0x0123);
                if (valid < 4) w &= 0xffffffffu << (8 * (4 - valid));
            }
            cw[wi] = w;
        }
    }
    __syncwarp();
    BitReader br;
    br.w = cw;
    br.pos = 0;
    br.nbits = 8 * nbytes;
    br.bad = false;
    const int table = (int)br.read(4);
    const int swA = (int)br.read(1), swB = (int)br.read(1);
    if (((swA << 1) | swB) != pg.geom) br.bad = true;          // both chunks of a pair carry the pair's geometry
    if (table != MRC_NO_TABLE && table >= MRC_N_HUFF_TABLES) br.bad = true;
    uint8_t* o_ovs = po.ovs + (size_t)lp * 4;
    if (joint) {
        if (ch == 0) {
            for (int i = 0; i < 4; ++i) { const int v = (int)br.read(pg.n_scale_bits); if (lane == 0) o_ovs[i] = (uint8_t)v; }
            unsigned ms = 0;
            for (int bd = 0; bd < nb; ++bd) ms |= br.read(1) << bd;
            if (lane == 0) po.ms[lp] = ms;
        }
    } else {
        const int v = (int)br.read(pg.n_scale_bits);
        if (lane == 0) { o_ovs[ch] = (uint8_t)v; if (ch == 0) po.ms[lp] = 0u; }
    }
    uint16_t* mant = po.mant + ((size_t)lp * 2 + ch) * pg.Lmax;
    uint8_t* o_alloc = po.alloc + ((size_t)lp * 2 + ch) * MRC_BSTRIDE;
    uint8_t* o_sf = po.sf + ((size_t)lp * 2 + ch) * MRC_BSTRIDE;
    int bd = 0;
    for (; bd < nb && !br.bad; ++bd) {
        int ba = (int)br.read(pg.n_mant_size_bits);
        if (ba) ba += 1;
        const int sfv = (int)br.read(pg.n_scale_bits);
        if (lane == 0) { o_alloc[bd] = (uint8_t)ba; o_sf[bd] = (uint8_t)sfv; }
        if (!ba) continue;
        const int lo = pg.band_lo[bd], n = pg.band_n[bd];
        if (table == MRC_NO_TABLE) {
            // raw mantissas: fixed width, one per lane and trip
            if (br.pos + n * ba > br.nbits) { br.bad = true; break; }
            for (int i = lane; i < n; i += 32) mant[lo + i] = (uint16_t)br.peek_at(br.pos + i * ba, ba);
            br.pos += n * ba;
        } else {
            // Huffman-coded mantissas, 32 bit positions at a time: every lane decodes the code that WOULD start at its
            // bit, the lanes that really are code starts are found by pointer doubling along "next start" (five steps),
            // and their symbols land at their rank
            const uint16_t* lut = s_hlut[table];
            const int esc = s_esc[table];
            int done = 0;
            while (done < n) {
                const int pi = br.pos + lane;
                int tot = 0, val = 0, len = 0;
                bool ok = false;
                if (pi < br.nbits) {
                    const uint32_t e = lut[br.peek_at(pi, MRC_HUFF_PEEK)];
                    len = e >> 8; val = e & 0xff;
                    if (len != 0 && pi + len <= br.nbits) {
                        tot = len;
                        ok = true;
                        if (val == esc) {                    // escape code + ba raw bits
                            if (pi + len + ba <= br.nbits) tot = len + ba; else ok = false;
                        }
                    }
                }
                int J = ok ? min(lane + tot, 32) : 32;       // where the next code starts; 32 = past this window
                unsigned R = 1u;                             // lanes that are code starts: the chain from lane 0
#pragma unroll
                for (int st = 0; st < 5; ++st) {
                    const bool on = (R >> lane) & 1u;
                    R |= __reduce_or_sync(0xffffffffu, (on && J < 32) ? (1u << J) : 0u);
                    const int Jn = __shfl_sync(0xffffffffu, J, J & 31);
                    J = (J < 32) ? Jn : 32;
                }
                const bool on = (R >> lane) & 1u;
                const int rank = __popc(R & ((1u << lane) - 1u));
                const int want = n - done, cnt = min(__popc(R), want);
                if (__ballot_sync(0xffffffffu, on && rank < want && !ok)) { br.bad = true; break; }
                if (on && rank < want)
                    mant[lo + done + rank] = (uint16_t)((val == esc) ? (int)br.peek_at(pi + len, ba) : val);
                const unsigned lastm = __ballot_sync(0xffffffffu, on && rank == cnt - 1);
                br.pos = __shfl_sync(0xffffffffu, pi + tot, __ffs(lastm) - 1);
                done += cnt;
            }
        }
    }
    if (lane == 0) {
        for (; bd < nb; ++bd) o_alloc[bd] = 0;               // fields a malformed chunk never reached
        po.flags[(size_t)lp * 2 + ch] = (uint8_t)((br.bad ? 1 : 0) | (joint ? 2 : 0));
        if (br.bad) atomicExch(error_flag, 1);
    }
}

// ---- synthesis of the parsed pairs: one CTA per pair ----------------------------------------------------------
template <typename T, int L_>
__global__ void __launch_bounds__(L_ / 2)
decode_kernel(DevTables<T> tb, CodecParams cp, DecodeMap dm, ParseOut po, T* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int L = L_, NT = L / 2;
    const int tid = threadIdx.x, nb = tb.nb;
    DSmem<T> sm = dcarve<T>(smem_raw, L, 0);
    __shared__ int s_alloc[2 * MRC_BSTRIDE], s_sf[2 * MRC_BSTRIDE], s_ovs[4];
    const int lp = dm.list ? dm.list[blockIdx.x] : (int)blockIdx.x;
    const unsigned f0 = po.flags[(size_t)lp * 2], f1 = po.flags[(size_t)lp * 2 + 1];
    const bool bad = ((f0 | f1) & 1u) != 0, joint = (f0 & 2u) != 0;
    // a malformed chunk decodes its pair as silence (the error flag is already raised)
    if (tid < 2 * nb) {
        const int ch = tid / nb, bd = tid - ch * nb;
        s_alloc[ch * MRC_BSTRIDE + bd] = bad ? 0 : (int)po.alloc[((size_t)lp * 2 + ch) * MRC_BSTRIDE + bd];
        s_sf[ch * MRC_BSTRIDE + bd] = (int)po.sf[((size_t)lp * 2 + ch) * MRC_BSTRIDE + bd];
    }
    if (tid < 4) s_ovs[tid] = bad ? 0 : (int)po.ovs[(size_t)lp * 4 + tid];
    const unsigned ms = (bad || !joint) ? 0u : po.ms[lp];
    {
        const uint32_t* m32 = reinterpret_cast<const uint32_t*>(po.mant + (size_t)lp * 2 * cp.Lmax);
        for (int i = tid; i < L; i += NT) {                  // two codes per load; channel stride Lmax
            const int ch = i / (L / 2), k2 = i - ch * (L / 2);
            const uint32_t w = m32[ch * (cp.Lmax / 2) + k2];
            sm.mant[ch * L + 2 * k2] = (int)(w & 0xffffu);
            sm.mant[ch * L + 2 * k2 + 1] = (int)(w >> 16);
        }
    }
    __syncthreads();
    synthesize<T, L>(tb, cp, sm, joint, s_alloc, s_sf, s_ovs, ms, y + (size_t)lp * 4 * cp.Lmax);
}

template <typename T, int L_>
__global__ void __launch_bounds__(L_ / 2)
decode_ints_kernel(DevTables<T> tb, CodecParams cp, int joint, const int32_t* __restrict__ sf,
                   const int32_t* __restrict__ alloc, const int32_t* __restrict__ mant,
                   const int32_t* __restrict__ ovs, const int32_t* __restrict__ ms, T* __restrict__ y) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int L = L_, NT = L / 2;
    const int tid = threadIdx.x, nb = tb.nb, lp = blockIdx.x;
    DSmem<T> sm = dcarve<T>(smem_raw, L, 0);
    __shared__ int s_alloc[2 * MRC_BSTRIDE], s_sf[2 * MRC_BSTRIDE], s_ovs[4];
    __shared__ unsigned s_ms;
    if (tid == 0) s_ms = 0u;
    __syncthreads();
    if (tid < 2 * nb) {
        const int ch = tid / nb, bd = tid - ch * nb;
        s_alloc[ch * MRC_BSTRIDE + bd] = alloc[(size_t)lp * 2 * nb + tid];
        s_sf[ch * MRC_BSTRIDE + bd] = sf[(size_t)lp * 2 * nb + tid];
    }
    if (tid < 4) s_ovs[tid] = ovs[(size_t)lp * 4 + tid];
    if (tid < nb && ms[(size_t)lp * nb + tid]) atomicOr(&s_ms, 1u << tid);
    for (int i = tid; i < 2 * L; i += NT) sm.mant[i] = mant[(size_t)lp * 2 * L + i];
    __syncthreads();
    synthesize<T, L>(tb, cp, sm, joint != 0, s_alloc, s_sf, s_ovs, joint ? s_ms : 0u, y + (size_t)lp * 4 * L);
}

// pcmfile.py:164-174: sign/magnitude, |x|>=1 -> 32767 else trunc((65535|x|+1)/2)
__device__ __forceinline__ int fraction_to_pcm(double x) {
    const double ax = fabs(x);
    const int q = (int)quant_mag_code(ax, 16);
    return (x < 0.0) ? -q : q;
}

template <typename T>
__global__ void __launch_bounds__(256)
ola_kernel(CodecParams cp, DecodeMap dm, int p0, const T* __restrict__ y, const int64_t* __restrict__ clip_frame_off,
           int16_t* __restrict__ pcm) {
    const int L = cp.Lmax, lp = blockIdx.x, p = p0 + lp, tid = threadIdx.x;
    __shared__ int s_clip, s_last;
    if (tid == 0) {
        int lo = 0, hi = dm.n_clips;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (dm.clip_pair0[mid] <= p) lo = mid; else hi = mid;
        }
        s_clip = lo;
        s_last = dm.clip_pair0[lo + 1] - 1;
    }
    __syncthreads();
    const int j = p - dm.clip_pair0[s_clip];
    const bool has_next = p < s_last;
    // this pair's window halves (a, b) and the next pair's; the saved tail (b samples) meets the next head (a' = b)
    int a = L, b = L, n2 = 2 * L;
    long long pos = (long long)j * L;
    if (dm.pair_geom) {
        const int q = dm.pair_geom[p];
        a = (q & 2) ? MRC_SHORT : L;
        b = (q & 1) ? MRC_SHORT : L;
        pos = dm.pair_pos[p];
        if (has_next) { const int q2 = dm.pair_geom[p + 1]; n2 = b + ((q2 & 1) ? MRC_SHORT : L); }
    }
    const int n1 = a + b;
    const T* cur = y + (size_t)lp * 4 * L;            // [2][a+b] of pair p
    const T* nxt = cur + 4 * L;                       // pair p+1 (same wave: waves hold whole clips)
    uint32_t* out = reinterpret_cast<uint32_t*>(pcm) + clip_frame_off[s_clip] + pos;
    for (int n = tid; n < b; n += blockDim.x) {
        double l = (double)cur[a + n], r = (double)cur[n1 + a + n];
        if (has_next) {       // np.add(overlapAndAdd, decoded[:a]): saved tail first, then the new head
            l = l + (double)nxt[n];
            r = r + (double)nxt[n2 + n];
        }
        const int cl = fraction_to_pcm(l), cr = fraction_to_pcm(r);
        out[n] = ((uint32_t)(uint16_t)(int16_t)cl) | ((uint32_t)(uint16_t)(int16_t)cr << 16);
    }
}

size_t decode_smem_bytes(int L, int elem, int cwords) {
    return (size_t)6 * L * elem + (size_t)2 * L * 4 + (size_t)2 * cwords * 4;
}

int chunk_words(const CodecParams& cp) {
    const int bits = 6 + 4 * cp.n_scale_bits + cp.nb + cp.nb * (cp.n_mant_size_bits + cp.n_scale_bits) + cp.L * 25;
    return (bits + 31) / 32 + 2;
}

}  // namespace

// layout of the parse output of one wave inside one allocation (every array 256-byte aligned)
static size_t po_align(size_t x) { return (x + 255) & ~(size_t)255; }
size_t parse_out_bytes(int npairs, int Lmax) {
    const size_t n = (size_t)npairs;
    return po_align(n * 2 * Lmax * 2) + 2 * po_align(n * 2 * MRC_BSTRIDE) + po_align(n * 4) + po_align(n * 4) + po_align(n * 2);
}
ParseOut parse_out_carve(void* base, int npairs, int Lmax) {
    const size_t n = (size_t)npairs;
    unsigned char* p = static_cast<unsigned char*>(base);
    ParseOut po;
    po.mant = reinterpret_cast<uint16_t*>(p);  p += po_align(n * 2 * Lmax * 2);
    po.alloc = p;                              p += po_align(n * 2 * MRC_BSTRIDE);
    po.sf = p;                                 p += po_align(n * 2 * MRC_BSTRIDE);
    po.ovs = p;                                p += po_align(n * 4);
    po.ms = reinterpret_cast<uint32_t*>(p);    p += po_align(n * 4);
    po.flags = p;
    return po;
}

template <typename T>
void launch_decode(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                   const HuffDecDev* hdec, const DecodeMap& dm, const uint8_t* pac, int p0, int npairs, T* y,
                   int* error_flag, const ParseOut& po) {
    if (npairs <= 0) return;
    const int cw = chunk_words(cp);
    ParseGeo pg;
    pg.nb = tb.nb; pg.geom = tb.geom; pg.L = tb.L; pg.Lmax = cp.Lmax;
    pg.n_scale_bits = cp.n_scale_bits; pg.n_mant_size_bits = cp.n_mant_size_bits;
    pg.joint = cp.joint; pg.flush_nonjoint = cp.flush_nonjoint;
    for (int b = 0; b < MRC_BSTRIDE; ++b) { pg.band_lo[b] = tb.c_band_lo[b]; pg.band_n[b] = tb.c_band_n[b]; }
    const size_t psmem = (size_t)PARSE_WARPS * cw * 4;
    cudaFuncSetAttribute(parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
    parse_kernel<<<(2 * npairs + PARSE_WARPS - 1) / PARSE_WARPS, PARSE_WARPS * 32, psmem, st>>>(
        pg, huff, hdec, dm, pac, p0, npairs, po, error_flag, cw);
    const size_t smem = decode_smem_bytes(tb.L, sizeof(T), 0);
#define MRC_LAUNCH_DEC(LL)                                                                                     \
    case LL:                                                                                                   \
        cudaFuncSetAttribute(decode_kernel<T, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        decode_kernel<T, LL><<<npairs, LL / 2, smem, st>>>(tb, cp, dm, po, y);                                  \
        break;
    switch (tb.L) {
        MRC_LAUNCH_DEC(128)
        MRC_LAUNCH_DEC(256)
        MRC_LAUNCH_DEC(512)
        MRC_LAUNCH_DEC(576)
        MRC_LAUNCH_DEC(1024)
        MRC_LAUNCH_DEC(2048)
        default: break;
    }
#undef MRC_LAUNCH_DEC
}

template <typename T>
void launch_decode_ints(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, int joint,
                        const int32_t* sf, const int32_t* alloc, const int32_t* mant, const int32_t* ovs,
                        const int32_t* ms, int npairs, T* y) {
    if (npairs <= 0) return;
    const size_t smem = decode_smem_bytes(tb.L, sizeof(T), 0);
#define MRC_LAUNCH_DECI(LL)                                                                                    \
    case LL:                                                                                                   \
        cudaFuncSetAttribute(decode_ints_kernel<T, LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        decode_ints_kernel<T, LL><<<npairs, LL / 2, smem, st>>>(tb, cp, joint, sf, alloc, mant, ovs, ms, y);     \
        break;
    switch (tb.L) {
        MRC_LAUNCH_DECI(128)
        MRC_LAUNCH_DECI(256)
        MRC_LAUNCH_DECI(512)
        MRC_LAUNCH_DECI(576)
        MRC_LAUNCH_DECI(1024)
        MRC_LAUNCH_DECI(2048)
        default: break;
    }
#undef MRC_LAUNCH_DECI
}

template <typename T>
void launch_ola(cudaStream_t st, const CodecParams& cp, const DecodeMap& dm, int p0, int npairs, const T* y,
                const int64_t* clip_frame_off, int16_t* pcm) {
    if (npairs <= 0) return;
    ola_kernel<T><<<npairs, 256, 0, st>>>(cp, dm, p0, y, clip_frame_off, pcm);
}

#define MRC_INST(T)                                                                                            \
    template void launch_decode<T>(cudaStream_t, const DevTables<T>&, const CodecParams&, const HuffDev*,       \
                                   const HuffDecDev*, const DecodeMap&, const uint8_t*, int, int, T*, int*,     \
                                   const ParseOut&);                                                            \
    template void launch_decode_ints<T>(cudaStream_t, const DevTables<T>&, const CodecParams&, int,             \
                                        const int32_t*, const int32_t*, const int32_t*, const int32_t*,         \
                                        const int32_t*, int, T*);                                               \
    template void launch_ola<T>(cudaStream_t, const CodecParams&, const DecodeMap&, int, int, const T*,         \
                                const int64_t*, int16_t*);
MRC_INST(double)
MRC_INST(float)
