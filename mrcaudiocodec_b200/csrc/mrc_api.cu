// mrc_api.cu -- host side of libmrc.so: context, tables, wave scheduling and the extern "C" entry points
// declared in include/mrc.h.  No CPU implementation of any codec stage lives here: every stage is a kernel.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>
#include <string>
#include <vector>

#include "mrc_internal.cuh"
#include "mrc_decode.cuh"

namespace {

std::string g_create_error;

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

struct TablesDev {
    Buf kbd, hann, tw_pre, tw_post, tw_fft, tw_rfft, tw9, tw_stage, bark, quiet, bark_d, quiet_d, exp_tab;
};

// Everything that depends on the block geometry (window halves a, b): long blocks always; the transition and short
// blocks of block switching once mrc_set_switch_tables has been called.
struct GeoDev {
    bool set = false;
    int a = 0, b = 0, L = 0, nb = 0;
    TablesDev td, tf;                 // double and float copies
    Buf band_lo, band_n, line2band;
    DevTables<double> tbd;
    DevTables<float> tbf;
    CodecParams cp;
    std::vector<int> h_band_lo, h_band_n;
};

}  // namespace

struct mrc_ctx {
    mrc_config cfg;
    int L = 0, logL = 0, nb = 0;
    bool tables_set = false;
    cudaStream_t stream = nullptr;
    std::string err;

    GeoDev geo[MRC_N_GEO];            // [MRC_GEO_LONG] is the context's own geometry (a = b = n_mdct_lines)
    DevTables<double>& tbd = geo[0].tbd;
    DevTables<float>& tbf = geo[0].tbf;
    CodecParams& cp = geo[0].cp;
    std::vector<int>& h_band_n = geo[0].h_band_n;
    Buf huff, header;
    HuffDev h_huff;
    // block switching
    bool switch_set = false;          // mrc_set_switch_tables called: all four geometries are available
    SosParams sos;
    Buf sb0, peaks, flags, blk_start, blk_geom, blk_list;
    uint8_t h_header[4 + 18 + 4 + 2 * MRC_MAX_BANDS];

    // scratch (grow only)
    Buf clip_off, clip_blk0, clip_bytes, clip_base, clip_res, clip_run, running, overflow, peakctr, res_in, res_out, bound;
    struct WaveSet { Buf lines, bandmax, tokens, ovs, ms, rec, pw, rsv, gmask, cblk, tab, comp, segx, rin; } sets[3];
    Buf q_alloc, q_sf, q_mant;
    cudaStream_t stream2 = nullptr, stream3 = nullptr;    // analysis stream, H2D copy stream
    cudaStream_t stream4 = nullptr;                       // D2H copy stream (bitstream of finished waves)
    cudaStream_t stream5 = nullptr;                       // composition of the reservoir maps (forks off stream2)
    int64_t* h_prog = nullptr;                            // pinned: per wave, how far the output is final
    unsigned char* h_idx = nullptr;                       // pinned: decode's chunk index of the wave being issued (3 slots)
    size_t h_idx_slot = 0;                                // bytes per slot
    cudaEvent_t h_idx_ev[3] = {nullptr, nullptr, nullptr};
    int h_prog_cap = 0;
    bool no_tables = false;          // MRC_FLAG_NO_CHAIN_TABLES
    int tab_min_blocks = 512;        // blocks per clip in a wave from which the reservoir maps are tabulated
    int seg_blocks = 32;             // blocks per composed reservoir map (0: walk the per-block maps one by one)
    std::vector<cudaEvent_t> evpool;
    Buf tap_lines, tap_smr, tap_npk;
    Buf pcm_dev, out_dev, xin_dev;
    Buf dec[20];

    double ms[8] = {0};
    int64_t counters[8] = {0};
    cudaEvent_t ev[8] = {nullptr};
};

namespace {

int fail(mrc_ctx* c, int code, const char* fmt, const char* detail = "") {
    char tmp[512];
    snprintf(tmp, sizeof tmp, fmt, detail);
    if (c) c->err = tmp; else g_create_error = tmp;
    return code;
}

#define CK(call)                                                                             \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) return fail(ctx, MRC_E_CUDA, "CUDA error: %s", cudaGetErrorString(e_)); \
    } while (0)

cudaError_t ensure(Buf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return cudaSuccess;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max<size_t>(bytes, 256);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e == cudaSuccess) b.cap = want;
    return e;
}

void release(Buf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

template <typename T>
cudaError_t upload(Buf& b, const std::vector<T>& v, cudaStream_t st) {
    cudaError_t e = ensure(b, v.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st);
}

template <typename T>
cudaError_t upload_tables(mrc_ctx* c, GeoDev& g, TablesDev& d, DevTables<T>& tb, const double* window,
                          const double* hann_w, const double* bark_t, const double* quiet_t) {
    const int L = g.L, N = 2 * L;
    std::vector<T> kbd(N), hann(N), bark(L), quiet(L);
    for (int i = 0; i < N; ++i) { kbd[i] = (T)window[i]; hann[i] = (T)hann_w[i]; }
    for (int i = 0; i < L; ++i) { bark[i] = (T)bark_t[i]; quiet[i] = (T)quiet_t[i]; }
    // twiddles: angles evaluated in double by the host libm
    int logLtab = 0;                                       // largest power of two dividing L: L itself, or L/9
    while (!((L >> logLtab) & 1)) ++logLtab;
    const int Ltab = 1 << logLtab;
    const bool pow2 = (Ltab == L);
    std::vector<cpx<T>> pre(L / 2), post(L / 2), fft(std::max(Ltab / 2, 1)), rfft(L), tw9(pow2 ? 1 : L);
    const double pi = 3.14159265358979323846;
    for (int n = 0; n < L / 2; ++n) {
        const double a = -pi * (4.0 * n + 1.0) / (4.0 * L);
        pre[n].x = (T)cos(a); pre[n].y = (T)sin(a);
        const double b = -pi * n / (double)L;
        post[n].x = (T)cos(b); post[n].y = (T)sin(b);
    }
    for (int n = 0; n < Ltab / 2; ++n) {
        const double f = -2.0 * pi * n / (double)Ltab;
        fft[n].x = (T)cos(f); fft[n].y = (T)sin(f);
    }
    if (!pow2)
        for (int m = 0; m < L; ++m) {
            const double f = -2.0 * pi * m / (double)L;
            tw9[m].x = (T)cos(f); tw9[m].y = (T)sin(f);
        }
    for (int m = 0; m < 9; ++m) {
        const double f = -2.0 * pi * m / 9.0;
        tb.w9[m].x = (T)cos(f); tb.w9[m].y = (T)sin(f);
    }
    for (int k = 0; k < L; ++k) {
        const double a = -2.0 * pi * k / (double)N;
        rfft[k].x = (T)cos(a); rfft[k].y = (T)sin(a);
    }
    cudaError_t e;
    if ((e = upload(d.kbd, kbd, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.hann, hann, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.bark, bark, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.quiet, quiet, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.tw_pre, pre, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.tw_post, post, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.tw_fft, fft, c->stream)) != cudaSuccess) return e;
    // per-stage twiddles of the L-point (Hann spectra) and L/2-point (MDCT) transforms, power-of-two L: stage h holds
    // W1[j] = exp(-2 pi i j / 4h) and W2[j] = exp(-2 pi i 2j / 4h), j < h (mrc_fft.cuh: fft_stage_entries, fft_sw)
    std::vector<cpx<T>> stage;
    if (pow2) {
        for (int lg = logLtab; lg >= logLtab - 1; --lg)
            for (int h = (lg & 1) ? 2 : 4; 4 * h <= (1 << lg); h <<= 2)
                for (int w = 1; w <= 2; ++w)
                    for (int j = 0; j < h; ++j) {
                        const double f = -2.0 * pi * (double)(w * j) / (double)(4 * h);
                        cpx<T> v; v.x = (T)cos(f); v.y = (T)sin(f);
                        stage.push_back(v);
                    }
    }
    if (stage.empty()) stage.resize(1);
    if ((e = upload(d.tw_stage, stage, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.tw_rfft, rfft, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.tw9, tw9, c->stream)) != cudaSuccess) return e;
    std::vector<double> bark_d(bark_t, bark_t + L), quiet_d(quiet_t, quiet_t + L), etab(64);
    for (int j = 0; j < 64; ++j) etab[j] = exp2(j / 64.0);       // correctly rounded by glibc
    if ((e = upload(d.bark_d, bark_d, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.quiet_d, quiet_d, c->stream)) != cudaSuccess) return e;
    if ((e = upload(d.exp_tab, etab, c->stream)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;   // host vectors die at scope exit
    tb.bark_d = (const double*)d.bark_d.p; tb.quiet_d = (const double*)d.quiet_d.p;
    tb.exp_tab = (const double*)d.exp_tab.p;
    tb.L = L; tb.logL = pow2 ? logLtab : 0; tb.nb = g.nb; tb.sample_rate = c->cfg.sample_rate;
    tb.fstep = c->cfg.sample_rate / N;
    tb.a = g.a; tb.b = g.b; tb.geom = (g.a != c->L ? 2 : 0) | (g.b != c->L ? 1 : 0);
    tb.rot = (g.a - g.b) / 4;
    tb.logLtab = logLtab;
    tb.tw9 = (const cpx<T>*)d.tw9.p;
    tb.kbd = (const T*)d.kbd.p; tb.hann = (const T*)d.hann.p;
    tb.tw_pre = (const cpx<T>*)d.tw_pre.p; tb.tw_post = (const cpx<T>*)d.tw_post.p;
    tb.tw_fft = (const cpx<T>*)d.tw_fft.p; tb.tw_rfft = (const cpx<T>*)d.tw_rfft.p;
    tb.tw_stage = (const cpx<T>*)d.tw_stage.p;
    tb.bark = (const T*)d.bark.p; tb.quiet = (const T*)d.quiet.p;
    tb.band_lo = (const int*)g.band_lo.p; tb.band_n = (const int*)g.band_n.p;
    tb.line2band = (const uint8_t*)g.line2band.p;
    memset(tb.c_band_lo, 0, sizeof tb.c_band_lo); memset(tb.c_band_n, 0, sizeof tb.c_band_n);
    for (int b = 0; b < g.nb; ++b) { tb.c_band_lo[b] = (uint16_t)g.h_band_lo[b]; tb.c_band_n[b] = (uint16_t)g.h_band_n[b]; }
    return cudaSuccess;
}

// band table of one geometry: lines per band, line -> band
int set_geo_bands(mrc_ctx* ctx, GeoDev& g, const int32_t* band_nlines, int n_bands) {
    if (n_bands < 1 || n_bands > MRC_CODED_BANDS)
        return fail(ctx, MRC_E_INVALID, "n_bands out of range (1..25: the reference's tables have 25 or 9 bands)");
    g.nb = n_bands;
    g.h_band_lo.assign(n_bands, 0);
    g.h_band_n.assign(n_bands, 0);
    int acc = 0;
    for (int b = 0; b < n_bands; ++b) {
        if (band_nlines[b] < 1)
            return fail(ctx, MRC_E_INVALID, "empty scale factor band (the reference crashes on these too)");
        g.h_band_lo[b] = acc;
        g.h_band_n[b] = band_nlines[b];
        acc += band_nlines[b];
    }
    if (acc != g.L) return fail(ctx, MRC_E_INVALID, "band_nlines must sum to the block's MDCT lines");
    std::vector<uint8_t> l2b(g.L);
    for (int b = 0; b < n_bands; ++b)
        for (int i = 0; i < g.h_band_n[b]; ++i) l2b[g.h_band_lo[b] + i] = (uint8_t)b;
    CK(upload(g.band_lo, g.h_band_lo, ctx->stream));
    CK(upload(g.band_n, g.h_band_n, ctx->stream));
    CK(upload(g.line2band, l2b, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MRC_OK;
}

// bit budgets of one geometry up to the point where the reservoir is added, in the reference's order of operations
void set_geo_budgets(mrc_ctx* ctx, GeoDev& g) {
    const mrc_config& c = ctx->cfg;
    const int halfN = g.L, nb = g.nb;
    CodecParams& cp = g.cp;
    {   // codecThem.py:381-391
        double B = c.target_bits_per_sample * halfN;
        B -= c.n_scale_bits * nb;
        B -= c.n_mant_size_bits * nb;
        B += B;
        B -= nb;
        B -= c.n_scale_bits * 4;
        cp.budget_joint = B;
    }
    {   // codecThem.py:299-306
        double B = c.target_bits_per_sample * halfN;
        B -= c.n_scale_bits * (nb + 1);
        B -= c.n_mant_size_bits * nb;
        B -= 1;
        B -= 1;
        cp.budget_single = B;
    }
    {   // integer form: bitBudget = k + frac + bitReservoir (the joint path's two blksw bits folded into k)
        const double bj = cp.budget_joint - 1.0 - 1.0, bs = cp.budget_single;
        cp.k_joint = (int)floor(bj);   cp.frac_joint = (bj - floor(bj)) > 0.0 ? 1 : 0;
        cp.k_single = (int)floor(bs);  cp.frac_single = (bs - floor(bs)) > 0.0 ? 1 : 0;
    }
    cp.L = g.L; cp.nb = nb; cp.Lmax = ctx->L;
    cp.n_scale_bits = c.n_scale_bits; cp.n_mant_size_bits = c.n_mant_size_bits;
    cp.max_mant_bits = std::min(16, 1 << c.n_mant_size_bits);
    cp.joint = c.joint; cp.flush_nonjoint = 1; cp.no_huff = 0;
    cp.spread_seq = (c.flags & MRC_FLAG_SPREAD_SEQUENTIAL) ? 1 : 0;
    cp.header_bytes = 26 + 2 * ctx->geo[0].nb;
}

template <typename T> DevTables<T>& tb_of(mrc_ctx* ctx, int q);
template <> DevTables<double>& tb_of<double>(mrc_ctx* ctx, int q) { return ctx->geo[q].tbd; }
template <> DevTables<float>& tb_of<float>(mrc_ctx* ctx, int q) { return ctx->geo[q].tbf; }

struct EncodeJob {
    // inputs (exactly one of d_pcm / d_xin)
    const int16_t* d_pcm = nullptr;
    const int16_t* h_pcm = nullptr;        // when set: frames are uploaded into d_pcm wave by wave on the copy stream
    const double* d_xin = nullptr;
    const int64_t* h_clip_off = nullptr;   // [n_clips+1]
    int n_clips = 0;
    bool flush_nonjoint = true;
    int joint = 1;
    int no_huff = 0;
    const int32_t* h_res_in = nullptr;     // [n_clips] or null
    int32_t* h_res_out = nullptr;
    // packed output (device), optional
    uint8_t* d_out = nullptr;
    int64_t out_cap = 0;
    int64_t* h_clip_byte_off = nullptr;    // [n_clips+1]
    // when set: the bytes of finished waves are copied to h_out (host) on the D2H stream while later waves run;
    // *h_copied = how many leading bytes of the output have been queued (the caller copies the rest and syncs stream4)
    uint8_t* h_out = nullptr;
    int64_t h_out_cap = 0;
    int64_t* h_copied = nullptr;
    // taps to host, optional
    double* t_lines = nullptr; int32_t* t_ovs = nullptr; int32_t* t_ms = nullptr; double* t_smr = nullptr;
    int32_t* t_npk = nullptr;
    int32_t* t_alloc = nullptr; int32_t* t_sf = nullptr; int32_t* t_mant = nullptr; int32_t* t_table = nullptr;
    int32_t* t_res = nullptr; int32_t* t_cbytes = nullptr;
    int32_t* t_res_mid = nullptr;          // reservoir between the two channels of a non-joint block (after channel 0)
    bool need_quant = true;
    bool dev_qtaps = false;                // leave alloc / mantissa taps of the (single) wave in q_alloc / q_mant
    int geom = MRC_GEO_LONG;               // d_xin jobs: the geometry of every block (per-block seam with a, b given)
    bool switching = false;                // d_pcm jobs: transient detector + look-ahead decide each block's geometry
    // one stream sharded by block range (mrc_encode_shard): a single clip whose h_clip_off spans the WHOLE stream, of
    // which this job encodes shard_blocks blocks starting at block shard_first (+ the flush block when flush_nonjoint);
    // d_pcm starts at stream frame pcm_frame0.  The reservoir comes from / goes to the neighbouring shards through
    // `exchange`, called once before the serial pass and once right after it.
    bool shard = false;
    int64_t shard_first = 0, shard_blocks = 0, pcm_frame0 = 0;
    bool shard_header = true;
    mrc_reservoir_exchange exchange = nullptr;
    void* exchange_user = nullptr;
};

// ---- block switching: which nMDCTLines-frame blocks are written as eight short blocks --------------------------
// Runs the detector over all clips (PCM already on the device), brings the two flags per block back and lays out
// the written blocks of every clip: the block at frames [kL, (k+1)L) is short iff it holds a transient after its
// first 128 samples or the NEXT block holds one in its first 128 samples (pacfileThem.py:1192; the last block of a
// clip has no look-ahead); a = the previous written block's b; the Close() flush block has b = nMDCTLines.
int plan_switched_blocks(mrc_ctx* ctx, const int16_t* d_pcm, const int64_t* h_clip_off, int nc, cudaStream_t st,
                         std::vector<int32_t>& blk0, std::vector<int64_t>& bstart, std::vector<uint8_t>& bgeom,
                         std::vector<uint8_t>* flags_out) {
    const int L = ctx->L, S = MRC_SHORT, nseg = L / S;
    std::vector<int32_t> sb0(nc + 1);
    long long nsb = 0;
    for (int c = 0; c < nc; ++c) {
        const long long fr = h_clip_off[c + 1] - h_clip_off[c];
        if (fr < 0) return fail(ctx, MRC_E_INVALID, "clip_frame_offsets must be non-decreasing");
        sb0[c] = (int32_t)nsb;
        nsb += (fr + L - 1) / L;
        if (nsb > 0x0fff0000ll) return fail(ctx, MRC_E_INVALID, "too many blocks in one call");
    }
    sb0[nc] = (int32_t)nsb;
    std::vector<uint8_t> flags((size_t)nsb);
    if (nsb > 0) {
        std::vector<int64_t> coff(h_clip_off, h_clip_off + nc + 1);
        CK(upload(ctx->clip_off, coff, st));
        CK(upload(ctx->sb0, sb0, st));
        CK(ensure(ctx->peaks, (size_t)nsb * 2 * nseg * 8));
        CK(ensure(ctx->flags, (size_t)nsb));
        CK(cudaEventRecord(ctx->ev[1], st));
        launch_transient(st, ctx->sos, (const int64_t*)ctx->clip_off.p, (const int32_t*)ctx->sb0.p, nc, d_pcm, L,
                         (int)nsb, (double*)ctx->peaks.p, (uint8_t*)ctx->flags.p);
        CK(cudaEventRecord(ctx->ev[2], st));
        CK(cudaMemcpyAsync(flags.data(), ctx->flags.p, (size_t)nsb, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        float t = 0;
        cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2]);
        ctx->ms[3] = t;
    }
    if (flags_out) *flags_out = flags;
    blk0.assign(nc + 1, 0);
    bstart.clear();
    bgeom.clear();
    for (int c = 0; c < nc; ++c) {
        blk0[c] = (int32_t)bstart.size();
        const int n = sb0[c + 1] - sb0[c];
        const uint8_t* f = flags.data() + sb0[c];
        bool a_short = false;
        for (int k = 0; k < n; ++k) {
            const bool is_short = (f[k] & 2) || (k + 1 < n && (f[k + 1] & 1));
            if (is_short) {
                for (int i = 0; i < nseg; ++i) {
                    bstart.push_back((int64_t)k * L + (int64_t)i * S);
                    bgeom.push_back((uint8_t)((a_short ? 2 : 0) | 1));
                    a_short = true;
                }
            } else {
                bstart.push_back((int64_t)k * L);
                bgeom.push_back((uint8_t)(a_short ? 2 : 0));
                a_short = false;
            }
        }
        bstart.push_back((int64_t)n * L);                     // Close(): nMDCTLines zeros, non-joint
        bgeom.push_back((uint8_t)(a_short ? 2 : 0));
        if (bstart.size() > 0x7fff0000ull) return fail(ctx, MRC_E_INVALID, "too many blocks in one call");
    }
    blk0[nc] = (int32_t)bstart.size();
    return MRC_OK;
}

constexpr int WAVE_BLOCKS = 1 << 14;   // blocks per wave: ~0.9 GB of hand-off per buffer set in fp64; the serial walk of
                                       // the last wave is the un-overlapped tail of a call, so waves are kept short
constexpr int NSETS = 3;               // buffer sets: waves w+1 and w+2 can be analysed while wave w is chained and packed
                                       // (the serial walk of some waves takes longer than an analysis: with two sets
                                       // the analysis stream stalled on the set that walk still held)

cudaEvent_t pool_event(mrc_ctx* ctx, size_t i) {
    while (ctx->evpool.size() <= i) {
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        ctx->evpool.push_back(e);
    }
    return ctx->evpool[i];
}

// Waves are ranges of consecutive blocks (clips may straddle them: the chain kernel carries reservoir and byte
// offset per clip from wave to wave).  Stream `stream2` runs analysis + cost of wave w+1 while `stream` runs the
// chain walk, the clip-offset scan and quantise+pack of wave w.
template <typename T>
int run_encode_t(mrc_ctx* ctx, const EncodeJob& job) {
    const int L = ctx->L, nc = job.n_clips;
    // taps of a per-block-seam job are sized by that block's geometry
    const int nb = ctx->geo[job.geom].nb, Lt = ctx->geo[job.geom].L;
    cudaStream_t st = ctx->stream, st2 = ctx->stream2;
    // ---- block map ----
    std::vector<int32_t> blk0(nc + 1);
    std::vector<int64_t> bstart;
    std::vector<uint8_t> bgeom;
    int64_t uploaded = 0;              // frames of job.h_pcm already queued for upload
    if (job.switching) {
        // the detector reads every clip before the first block can be laid out: upload all PCM first
        if (job.h_pcm) {
            const int64_t frames = job.h_clip_off[nc];
            if (frames > 0)
                CK(cudaMemcpyAsync((int16_t*)job.d_pcm, job.h_pcm, (size_t)frames * 4, cudaMemcpyHostToDevice, st));
            uploaded = frames;
        }
        const int rc = plan_switched_blocks(ctx, job.d_pcm, job.h_clip_off, nc, st, blk0, bstart, bgeom, nullptr);
        if (rc != MRC_OK) return rc;
    } else {
        long long tot = 0;
        for (int c = 0; c < nc; ++c) {
            const long long fr = job.h_clip_off[c + 1] - job.h_clip_off[c];
            if (fr < 0) return fail(ctx, MRC_E_INVALID, "clip_frame_offsets must be non-decreasing");
            blk0[c] = (int32_t)tot;
            if (job.shard) tot += job.shard_blocks + (job.flush_nonjoint ? 1 : 0);
            else tot += job.d_xin ? fr / (2 * Lt) : (fr + L - 1) / L + (job.flush_nonjoint ? 1 : 0);
            if (tot > 0x7fff0000ll) return fail(ctx, MRC_E_INVALID, "too many blocks in one call");
        }
        blk0[nc] = (int32_t)tot;
    }
    const int nblk_total = blk0[nc];
    std::vector<int64_t> coff(job.h_clip_off, job.h_clip_off + nc + 1);
    CK(upload(ctx->clip_off, coff, st));
    CK(upload(ctx->clip_blk0, blk0, st));
    CK(ensure(ctx->clip_bytes, (size_t)(nc + 1) * 8));
    CK(ensure(ctx->clip_base, (size_t)(nc + 2) * 8));
    CK(ensure(ctx->clip_res, (size_t)(nc + 1) * 4));
    CK(ensure(ctx->clip_run, (size_t)(nc + 1) * 8));
    CK(ensure(ctx->running, 8));
    CK(ensure(ctx->overflow, 4));
    CK(ensure(ctx->peakctr, 128));
    CK(cudaMemsetAsync(ctx->clip_bytes.p, 0, (size_t)(nc + 1) * 8, st));
    CK(cudaMemsetAsync(ctx->clip_base.p, 0, (size_t)(nc + 2) * 8, st));
    CK(cudaMemsetAsync(ctx->running.p, 0, 8, st));
    CK(cudaMemsetAsync(ctx->overflow.p, 0, 4, st));
    CK(cudaMemsetAsync(ctx->peakctr.p, 0, 128, st));
    const int32_t* d_res_in = nullptr;
    int32_t* d_res_out = nullptr;
    if (job.exchange) {
        CK(ensure(ctx->res_in, (size_t)nc * 4));
        CK(ensure(ctx->res_out, (size_t)nc * 4));
        d_res_in = (const int32_t*)ctx->res_in.p;
        d_res_out = (int32_t*)ctx->res_out.p;
    }
    if (job.h_res_in) {
        CK(ensure(ctx->res_in, (size_t)nc * 4));
        CK(cudaMemcpyAsync(ctx->res_in.p, job.h_res_in, (size_t)nc * 4, cudaMemcpyHostToDevice, st));
        d_res_in = (const int32_t*)ctx->res_in.p;
    }
    if (job.h_res_out) {
        CK(ensure(ctx->res_out, (size_t)nc * 4));
        d_res_out = (int32_t*)ctx->res_out.p;
    }
    ClipMap cm;
    cm.clip_off = (const int64_t*)ctx->clip_off.p;
    cm.clip_blk0 = (const int32_t*)ctx->clip_blk0.p;
    cm.n_clips = nc;
    cm.blk_start = nullptr; cm.blk_geom = nullptr; cm.list = nullptr;
    cm.pcm_frame0 = job.shard ? job.pcm_frame0 : 0;
    cm.blk_base = job.shard ? (int)job.shard_first : 0;
    // geometries in play, their launch parameters, and (block switching) the per-wave block lists of each
    const int q_lo = job.switching ? 0 : job.geom, q_hi = job.switching ? MRC_N_GEO - 1 : job.geom;
    CodecParams cpq[MRC_N_GEO];
    int max_nl = 0;
    for (int q = q_lo; q <= q_hi; ++q) {
        cpq[q] = ctx->geo[q].cp;
        cpq[q].joint = job.joint;
        cpq[q].flush_nonjoint = job.flush_nonjoint ? 1 : 0;
        cpq[q].no_huff = job.no_huff;
        if (job.shard && !job.shard_header) cpq[q].header_bytes = 0;       // the stream's first shard writes the file header
        for (int v : ctx->geo[q].h_band_n) max_nl = std::max(max_nl, v);
    }
    const CodecParams& cp = cpq[q_lo];     // for the kernels that read only geometry-independent fields
    // Wave boundaries.  Full waves hold WAVE_BLOCKS blocks; the first ones ramp up (the first analysis starts after a
    // small PCM upload) and the last ones ramp down (what is left un-overlapped at the end of a call is the serial walk
    // and the packing of the last wave: 1024 blocks instead of up to 16384).  A histogram / taps job of one wave stays
    // one wave.
    std::vector<int> wave_g0;              // [nwaves+1]
    {
        std::vector<int> sizes, tail;
        int rem = nblk_total, tailsum = 0;
        if (nblk_total > 4096 && !job.dev_qtaps && !job.shard) {    // a shard is one wave: its tables stay resident while it
                                                                    // waits for the reservoir of the shard before it
            for (int sz = 1024; sz <= WAVE_BLOCKS / 2 && tailsum + sz <= rem / 2; sz *= 2) { tail.push_back(sz); tailsum += sz; }
            int body = rem - tailsum;
            for (int sz = 2048; sz < WAVE_BLOCKS && sz <= body / 4; sz *= 2) { sizes.push_back(sz); body -= sz; }
            while (body > 0) { const int sz = std::min(body, WAVE_BLOCKS); sizes.push_back(sz); body -= sz; }
            for (size_t i = tail.size(); i-- > 0;) sizes.push_back(tail[i]);
        } else if (nblk_total > 0) {
            sizes.push_back(nblk_total);
        }
        wave_g0.push_back(0);
        for (int sz : sizes) wave_g0.push_back(wave_g0.back() + sz);
    }
    const int nwaves = (int)wave_g0.size() - 1;
    std::vector<int32_t> lists;            // [wave][geometry] wave-local indices, back to back
    std::vector<int> list_off((size_t)nwaves * MRC_N_GEO + 1, 0);
    if (job.switching) {
        CK(upload(ctx->blk_start, bstart, st));
        CK(upload(ctx->blk_geom, bgeom, st));
        cm.blk_start = (const int64_t*)ctx->blk_start.p;
        cm.blk_geom = (const uint8_t*)ctx->blk_geom.p;
        lists.reserve((size_t)nblk_total);
        for (int w = 0; w < nwaves; ++w) {
            const int g0 = wave_g0[w], nblk = wave_g0[w + 1] - g0;
            for (int q = 0; q < MRC_N_GEO; ++q) {
                list_off[(size_t)w * MRC_N_GEO + q] = (int)lists.size();
                for (int i = 0; i < nblk; ++i)
                    if (bgeom[(size_t)g0 + i] == q) lists.push_back(i);
            }
        }
        list_off[(size_t)nwaves * MRC_N_GEO] = (int)lists.size();
        CK(upload(ctx->blk_list, lists, st));
    }
    // reservoir-map tables of the single-stream fast path: R_in in [r_lo, r_lo + ntab)
    // [-128, 640) for long calls, [-128, 1024) for shorter ones: reservoirs outside the table take the complete walk
    // (exact either way).  The bench stream's reservoir has its median near 500 and 11 % of the blocks in [640, 1024):
    // tabulating those costs 2 ms per hour on the analysis stream and takes 12 ms off the serial walk, which pays when
    // the walk is what a call waits for -- single streams of less than about 45 minutes, where there are too few
    // waves to hide it (profiles/r01zz_chain_table_range.log: 600 s stream 26.8 k -> 32.8 k audio-s/s, 1 h stream
    // 43.6 k -> 42.7 k).  Reservoirs below -128 are possible (down to -(largest band + 1)) but rare.
    int r_lo = -std::min(128, (max_nl + 1 + 31) / 32 * 32), r_hi = 1024;
    if (const char* e = getenv("MRC_CHAIN_TABLE_LO")) r_lo = -std::max(32, (atoi(e) + 31) / 32 * 32);     // tuning knobs: any
    if (const char* e = getenv("MRC_CHAIN_TABLE_HI")) r_hi = std::max(32, (atoi(e) + 31) / 32 * 32);      // range is exact
    const int ntab = -r_lo + r_hi, tabw = (ntab + 2 + 3) / 4 * 4, segw = tabw + 4;
    const int seg_S = (ntab <= segment_max_ntab()) ? ctx->seg_blocks : 0;

    // ---- buffer sets ----
    const bool want_atap = job.t_lines || job.t_smr || job.t_npk;
    const bool want_qtap = job.t_alloc || job.t_sf || job.t_mant || job.dev_qtaps;
    const bool any_tap = want_atap || job.t_alloc || job.t_sf || job.t_mant || job.t_ovs || job.t_ms || job.t_table || job.t_res ||
                         job.t_cbytes || job.t_res_mid;
    size_t W = 1;
    for (int w = 0; w < nwaves; ++w) W = std::max<size_t>(W, (size_t)(wave_g0[w + 1] - wave_g0[w]));
    const int nsets = std::max(1, std::min(nwaves, NSETS));
    Handoff<T> ho[NSETS];
    ChainIO io[NSETS];
    for (int s = 0; s < nsets; ++s) {
        mrc_ctx::WaveSet& ws = ctx->sets[s];
        CK(ensure(ws.lines, W * 2 * L * sizeof(T)));
        CK(ensure(ws.bandmax, W * 2 * MRC_BSTRIDE * sizeof(T)));
        CK(ensure(ws.tokens, W * MRC_TOK_STRIDE * 2));
        CK(ensure(ws.ovs, W * 4));
        CK(ensure(ws.ms, W * 4));
        CK(ensure(ws.rec, W * MRC_REC_BYTES));
        CK(ensure(ws.pw, W * MRC_PW_BYTES));
        CK(ensure(ws.rsv, W * sizeof(int4)));
        CK(ensure(ws.gmask, W * 32 * 4));
        CK(ensure(ws.cblk, W * sizeof(ChainBlk)));
        ho[s].lines = (T*)ws.lines.p; ho[s].bandmax = (T*)ws.bandmax.p; ho[s].smr = nullptr;
        ho[s].tokens = (uint16_t*)ws.tokens.p; ho[s].ovs = (uint8_t*)ws.ovs.p; ho[s].ms = (uint32_t*)ws.ms.p;
        io[s].rec = (const unsigned char*)ws.rec.p; io[s].pw = (const unsigned char*)ws.pw.p;
        io[s].rsv = (int4*)ws.rsv.p; io[s].gmask = (uint32_t*)ws.gmask.p;
        io[s].cblk = (ChainBlk*)ws.cblk.p;
        io[s].clip_res = (int32_t*)ctx->clip_res.p; io[s].clip_run = (int64_t*)ctx->clip_run.p;
        io[s].clip_bytes = (int64_t*)ctx->clip_bytes.p;
    }
    PackTaps ptaps;
    ptaps.alloc = ptaps.sf = nullptr; ptaps.mant = nullptr;
    if (want_qtap) {
        CK(ensure(ctx->q_alloc, W * 2 * MRC_BSTRIDE));
        CK(ensure(ctx->q_sf, W * 2 * MRC_BSTRIDE));
        CK(ensure(ctx->q_mant, W * 2 * L * 2));
        ptaps.alloc = (uint8_t*)ctx->q_alloc.p; ptaps.sf = (uint8_t*)ctx->q_sf.p; ptaps.mant = (uint16_t*)ctx->q_mant.p;
    }
    AnalysisTaps<T> taps;
    taps.lines4 = nullptr; taps.smr4 = nullptr; taps.npeaks = nullptr;
    if (want_atap) {
        CK(ensure(ctx->tap_lines, W * 4 * L * sizeof(T)));
        CK(ensure(ctx->tap_smr, W * 4 * MRC_BSTRIDE * sizeof(T)));
        CK(ensure(ctx->tap_npk, W * 4 * 4));
        taps.lines4 = (T*)ctx->tap_lines.p; taps.smr4 = (T*)ctx->tap_smr.p; taps.npeaks = (int32_t*)ctx->tap_npk.p;
    }

    // events per wave: 0 analysis start, 1 analysis end, 2 cost end, 3 chain start, 4 chain end, 5 pack end,
    // 6 PCM of the wave uploaded (7 = that upload queued), 8-9 around the bitstream copy drained after the wave
    // 10-11 around the composition of the reservoir maps on its own stream
    constexpr int EPW = 12;
    for (int i = 0; i < nwaves * EPW; ++i) pool_event(ctx, (size_t)i);
    auto ev = [&](int w, int k) { return ctx->evpool[(size_t)w * EPW + k]; };
    // everything queued on `st` so far (tables of this call, PCM upload) must precede the first analysis
    CK(cudaEventRecord(ctx->ev[0], st));
    CK(cudaStreamWaitEvent(st2, ctx->ev[0], 0));

    int launches = 0;
    int64_t copied = 0;                // leading output bytes already queued for the host
    std::vector<char> wave_ended(nwaves, 0), wave_h2d(nwaves, 0), wave_d2h(nwaves, 0);
    const bool stream_out = job.h_out != nullptr && job.d_out != nullptr && job.need_quant;
    if (stream_out && ctx->h_prog_cap < 2 * nwaves) {
        if (ctx->h_prog) cudaFreeHost(ctx->h_prog);
        ctx->h_prog = nullptr;
        ctx->h_prog_cap = 0;
        CK(cudaMallocHost((void**)&ctx->h_prog, (size_t)(2 * nwaves + 64) * 8));
        ctx->h_prog_cap = 2 * nwaves + 64;
    }
    // Output bytes are final up to the end of the last block of a finished wave (clips are laid out in order, and a
    // clip's base is known once its predecessors are complete).  After wave v+2 has been queued the host waits for
    // wave v's pack, reads that position and queues the copy; the kernels of the later waves hide it.
    auto drain = [&](int v) -> int {
        CK(cudaEventSynchronize(ev(v, 5)));
        int64_t done = ctx->h_prog[2 * v] + (wave_ended[v] ? 0 : ctx->h_prog[2 * v + 1]);
        done = std::min(done, std::min(job.h_out_cap, job.out_cap));
        if (done > copied) {
            CK(cudaEventRecord(ev(v, 8), ctx->stream4));
            CK(cudaMemcpyAsync(job.h_out + copied, job.d_out + copied, (size_t)(done - copied), cudaMemcpyDeviceToHost,
                               ctx->stream4));
            CK(cudaEventRecord(ev(v, 9), ctx->stream4));
            wave_d2h[v] = 1;
            copied = done;
        }
        return MRC_OK;
    };
    std::vector<unsigned char> hb;     // host bounce buffer for taps
    int c_lo = 0;
    cudaStream_t st3 = ctx->stream3;
    if (job.h_pcm) CK(cudaStreamWaitEvent(st3, ctx->ev[0], 0));
    for (int w = 0; w < nwaves; ++w) {
        const int s = w % nsets;
        const int g0 = wave_g0[w], nblk = wave_g0[w + 1] - g0;
        while (blk0[c_lo + 1] <= g0) ++c_lo;                       // clip holding block g0
        int c_hi = c_lo;
        while (blk0[c_hi + 1] < g0 + nblk) ++c_hi;                 // clip holding block g0+nblk-1
        const int nfin = (blk0[c_hi + 1] <= g0 + nblk) ? c_hi - c_lo + 1 : c_hi - c_lo;   // clips ending in this wave
        if (job.h_pcm && !job.switching) {
            // frames this wave reads: up to the end of its last block (or of that clip); everything before was needed
            // by earlier blocks, so the upload front only moves forward.  The copy stream runs ahead of the kernels.
            const long long b_last = (long long)(g0 + nblk - 1) - blk0[c_hi];
            const int64_t fr_hi = job.h_clip_off[c_hi + 1] - job.h_clip_off[c_hi];
            const int64_t need = job.h_clip_off[c_hi] + std::min<int64_t>(fr_hi, (b_last + 1) * (int64_t)L);
            if (need > uploaded) {
                CK(cudaEventRecord(ev(w, 7), st3));
                CK(cudaMemcpyAsync((int16_t*)job.d_pcm + uploaded * 2, job.h_pcm + uploaded * 2,
                                   (size_t)(need - uploaded) * 4, cudaMemcpyHostToDevice, st3));
                uploaded = need;
                wave_h2d[w] = 1;
            }
            CK(cudaEventRecord(ev(w, 6), st3));
            CK(cudaStreamWaitEvent(st2, ev(w, 6), 0));
        }
        // ---- stream 2: analysis + cost (needs buffer set s free: pack of wave w-nsets done) ----
        if (w >= nsets) CK(cudaStreamWaitEvent(st2, ev(w - nsets, 5), 0));
        CK(cudaEventRecord(ev(w, 0), st2));
        // one launch per geometry present in the wave (a single one unless block switching is on)
        auto for_geos = [&](auto&& fn) {
            for (int q = q_lo; q <= q_hi; ++q) {
                ClipMap cmq = cm;
                int n = nblk;
                if (job.switching) {
                    const int o = list_off[(size_t)w * MRC_N_GEO + q];
                    n = list_off[(size_t)w * MRC_N_GEO + q + 1] - o;
                    cmq.list = (const int32_t*)ctx->blk_list.p + o;
                }
                if (n > 0) { fn(q, cmq, n); ++launches; }
            }
        };
        // long stretches of one clip in the wave: the serial walk is the critical path, tabulate the reservoir maps
        const bool use_tab = job.need_quant && !ctx->no_tables && nblk / (c_hi - c_lo + 1) >= ctx->tab_min_blocks;
        bool seg_forked = false;
        for_geos([&](int q, const ClipMap& cmq, int n) {
            launch_analysis<T>(st2, tb_of<T>(ctx, q), cpq[q], cmq, job.d_pcm, job.d_xin, g0, n, ho[s], taps,
                               (unsigned long long*)ctx->peakctr.p);
        });
        CK(cudaEventRecord(ev(w, 1), st2));
        if (job.need_quant)
            for_geos([&](int q, const ClipMap& cmq, int n) {
                launch_cost<T>(st2, tb_of<T>(ctx, q), cpq[q], (const HuffDev*)ctx->huff.p, cmq, g0, n, ho[s],
                               (unsigned char*)ctx->sets[s].rec.p, (unsigned char*)ctx->sets[s].pw.p);
            });
        if (use_tab) {
            // The reservoir maps (per block, then composed over segments) only feed the serial pass: they run on a stream of
            // their own that forks off after the cost kernel, so that the next wave's analysis does not queue behind them
            // (the composition is a short kernel of few CTAs followed by a latency-bound one).
            CK(ensure(ctx->sets[s].tab, W * 2 * (size_t)tabw * 4));
            CK(cudaEventRecord(ev(w, 10), st2));
            CK(cudaStreamWaitEvent(ctx->stream5, ev(w, 10), 0));
            launch_table(ctx->stream5, cp, cm, g0, nblk, io[s], r_lo, ntab, tabw, (int*)ctx->sets[s].tab.p);
            ++launches;
            if (seg_S > 0) {
                const size_t nseg = (W + seg_S - 1) / seg_S;
                CK(ensure(ctx->sets[s].comp, nseg * (size_t)segw * 4));
                CK(ensure(ctx->sets[s].rin, nseg * 4));
                CK(ensure(ctx->sets[s].segx, nseg * (size_t)segment_aux_width() * 4));
                launch_segments(ctx->stream5, cp, cm, g0, nblk, seg_S, io[s], r_lo, ntab, tabw, (const int*)ctx->sets[s].tab.p,
                                segw, (int*)ctx->sets[s].comp.p, (int*)ctx->sets[s].segx.p, (int*)ctx->sets[s].rin.p);
                launches += 2;
            }
            CK(cudaEventRecord(ev(w, 11), ctx->stream5));
            seg_forked = true;
        }
        CK(cudaEventRecord(ev(w, 2), st2));
        // ---- main stream: chain -> clip offsets -> quantise + pack ----
        CK(cudaStreamWaitEvent(st, ev(w, 2), 0));
        if (seg_forked) CK(cudaStreamWaitEvent(st, ev(w, 11), 0));
        int32_t shard_res = 0;
        // A shard that is not the stream's first walks AHEAD of its reservoir: the serial pass runs from a guessed value
        // while the shards before it are still at work, leaving the reservoir at every segment boundary; when the true value
        // arrives, a second pass walks from it only until it meets that trajectory (the maps contract: a segment or two),
        // and the reservoir for the next shard is known without another pass over the whole shard.  Exact: from a common
        // value on, the two walks are the same walk.
        const bool ahead = job.exchange && job.shard && !job.shard_header && job.need_quant && use_tab && seg_S > 0 && nc == 1 &&
                           nwaves == 1 && !getenv("MRC_SHARD_NO_SPECULATION");
        int* d_bound = nullptr;
        if (ahead) {
            const size_t nseg = ((size_t)nblk + seg_S - 1) / seg_S;
            CK(ensure(ctx->bound, (nseg + 1) * 4));
            d_bound = (int*)ctx->bound.p;
            const int32_t guess = 512;
            CK(cudaMemcpyAsync(ctx->res_in.p, &guess, 4, cudaMemcpyHostToDevice, st));
            launch_chain_seg(st, cp, cm, c_lo, c_hi - c_lo + 1, g0, nblk, seg_S, io[s], r_lo, ntab, tabw,
                             (const int*)ctx->sets[s].tab.p, segw, (const int*)ctx->sets[s].comp.p,
                             (const int*)ctx->sets[s].segx.p, (int*)ctx->sets[s].rin.p, d_res_in, d_res_out,
                             (unsigned long long*)ctx->peakctr.p + 4, d_bound, nullptr);
            ++launches;
        }
        if (job.exchange) {
            // everything that does not depend on the reservoir is done (or running); now wait for the shard before us
            CK(cudaStreamSynchronize(st));
            if (job.exchange(job.exchange_user, 0, &shard_res) != 0)
                return fail(ctx, MRC_E_STATE, "reservoir exchange callback failed (receive)");
            CK(cudaMemcpyAsync(ctx->res_in.p, &shard_res, 4, cudaMemcpyHostToDevice, st));
        }
        CK(cudaEventRecord(ev(w, 3), st));
        if (job.need_quant) {
            if (ahead) {
                if (shard_res != 512) {
                    launch_chain_seg(st, cp, cm, c_lo, c_hi - c_lo + 1, g0, nblk, seg_S, io[s], r_lo, ntab, tabw,
                                     (const int*)ctx->sets[s].tab.p, segw, (const int*)ctx->sets[s].comp.p,
                                     (const int*)ctx->sets[s].segx.p, (int*)ctx->sets[s].rin.p, d_res_in, d_res_out,
                                     (unsigned long long*)ctx->peakctr.p + 4, nullptr, d_bound);
                    ++launches;
                }
            } else if (use_tab && seg_S > 0) {
                launch_chain_seg(st, cp, cm, c_lo, c_hi - c_lo + 1, g0, nblk, seg_S, io[s], r_lo, ntab, tabw,
                                 (const int*)ctx->sets[s].tab.p, segw, (const int*)ctx->sets[s].comp.p,
                                 (const int*)ctx->sets[s].segx.p, (int*)ctx->sets[s].rin.p, d_res_in, d_res_out,
                                 (unsigned long long*)ctx->peakctr.p + 4);
                launches += 2;
            } else if (use_tab)
                launch_chain_table(st, cp, cm, c_lo, c_hi - c_lo + 1, g0, nblk, io[s], r_lo, ntab, tabw,
                                   (const int*)ctx->sets[s].tab.p, d_res_in, d_res_out,
                                   (unsigned long long*)ctx->peakctr.p + 4);
            else
                launch_chain(st, cp, cm, c_lo, c_hi - c_lo + 1, g0, nblk, io[s], d_res_in, d_res_out,
                             (unsigned long long*)ctx->peakctr.p + 4);
            ++launches;
        }
        if (job.exchange) {
            // hand the reservoir on before anything else: the next shard's serial pass waits for nothing but this
            CK(cudaMemcpyAsync(&shard_res, ctx->res_out.p, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (job.exchange(job.exchange_user, 1, &shard_res) != 0)
                return fail(ctx, MRC_E_STATE, "reservoir exchange callback failed (send)");
        }
        if (job.need_quant && use_tab && seg_S > 0) {
            launch_expand(st, cp, cm, g0, nblk, seg_S, io[s], r_lo, ntab, tabw, (const int*)ctx->sets[s].tab.p,
                          (const int*)ctx->sets[s].rin.p);
        }
        CK(cudaEventRecord(ev(w, 4), st));
        if (job.need_quant) {
            launch_finish(st, cp, cm, g0, nblk, io[s]);
            launch_offsets(st, cp, cm, c_lo, c_hi - c_lo + 1, g0, nblk, io[s]);
            launches += 2;
        }
        if (job.need_quant && (job.d_out || want_qtap)) {
            if (nfin > 0) {
                launch_clip_scan(st, io[s].clip_bytes, (int64_t*)ctx->clip_base.p, c_lo, nfin, (int64_t*)ctx->running.p);
                ++launches;
            }
            for_geos([&](int q, const ClipMap& cmq, int n) {
                launch_pack<T>(st, tb_of<T>(ctx, q), cpq[q], (const HuffDev*)ctx->huff.p, cmq, g0, n, ho[s], io[s], ptaps,
                               (const int64_t*)ctx->clip_base.p, job.d_out, job.out_cap, (const uint8_t*)ctx->header.p,
                               (int*)ctx->overflow.p);
            });
        }
        if (stream_out) {
            const bool ended = blk0[c_hi + 1] <= g0 + nblk;      // the wave's last clip ends with the wave
            wave_ended[w] = ended ? 1 : 0;
            CK(cudaMemcpyAsync(&ctx->h_prog[2 * w], (const int64_t*)ctx->clip_base.p + (ended ? c_hi + 1 : c_hi), 8,
                               cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&ctx->h_prog[2 * w + 1], (const int64_t*)ctx->clip_run.p + c_hi, 8,
                               cudaMemcpyDeviceToHost, st));
        }
        CK(cudaEventRecord(ev(w, 5), st));
        CK(cudaGetLastError());
        // the host waits for the pack of a wave two behind the one it has just queued: with three buffer sets the
        // device never runs out of queued work while the host sleeps on that event
        if (stream_out && w >= 2) { const int rc = drain(w - 2); if (rc != MRC_OK) return rc; }
        if (!any_tap) continue;
        // ---- taps of this wave to the host (parity runs only; synchronous) ----
        auto fetch = [&](const void* dsrc, size_t bytes) -> cudaError_t {
            hb.resize(bytes);
            cudaError_t e = cudaMemcpyAsync(hb.data(), dsrc, bytes, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) return e;
            return cudaStreamSynchronize(st);
        };
        if (job.t_lines) {          // device rows hold Lmax lines; the caller's hold the geometry's Lt
            CK(fetch(taps.lines4, (size_t)nblk * 4 * L * sizeof(T)));
            const T* src = (const T*)hb.data();
            double* d = job.t_lines + (size_t)g0 * 4 * Lt;
            for (size_t r = 0; r < (size_t)nblk * 4; ++r)
                for (int i = 0; i < Lt; ++i) d[r * Lt + i] = (double)src[r * L + i];
        }
        if (job.t_smr) {
            CK(fetch(taps.smr4, (size_t)nblk * 4 * MRC_BSTRIDE * sizeof(T)));
            const T* src = (const T*)hb.data();
            for (int b = 0; b < nblk; ++b)
                for (int c = 0; c < 4; ++c)
                    for (int k = 0; k < nb; ++k)
                        job.t_smr[((size_t)(g0 + b) * 4 + c) * nb + k] = (double)src[((size_t)b * 4 + c) * MRC_BSTRIDE + k];
        }
        if (job.t_npk) {
            CK(fetch(taps.npeaks, (size_t)nblk * 16));
            memcpy(job.t_npk + (size_t)g0 * 4, hb.data(), (size_t)nblk * 16);
        }
        if (job.t_ovs) {
            CK(fetch(ho[s].ovs, (size_t)nblk * 4));
            for (size_t i = 0; i < (size_t)nblk * 4; ++i) job.t_ovs[(size_t)g0 * 4 + i] = hb[i];
        }
        if (job.t_ms) {
            CK(fetch(ho[s].ms, (size_t)nblk * 4));
            const uint32_t* src = (const uint32_t*)hb.data();
            for (int b = 0; b < nblk; ++b)
                for (int k = 0; k < nb; ++k) job.t_ms[(size_t)(g0 + b) * nb + k] = (src[b] >> k) & 1u;
        }
        if (job.t_alloc || job.t_sf) {
            for (int which = 0; which < 2; ++which) {
                int32_t* dst = which ? job.t_sf : job.t_alloc;
                if (!dst) continue;
                CK(fetch(which ? ptaps.sf : ptaps.alloc, (size_t)nblk * 2 * MRC_BSTRIDE));
                for (int b = 0; b < nblk; ++b)
                    for (int c = 0; c < 2; ++c)
                        for (int k = 0; k < nb; ++k)
                            dst[((size_t)(g0 + b) * 2 + c) * nb + k] = hb[((size_t)b * 2 + c) * MRC_BSTRIDE + k];
            }
        }
        if (job.t_mant) {
            CK(fetch(ptaps.mant, (size_t)nblk * 2 * L * 2));
            const uint16_t* src = (const uint16_t*)hb.data();
            int32_t* d = job.t_mant + (size_t)g0 * 2 * Lt;
            for (size_t r = 0; r < (size_t)nblk * 2; ++r)
                for (int i = 0; i < Lt; ++i) d[r * Lt + i] = src[r * L + i];
        }
        if (job.t_res_mid) {
            CK(fetch(io[s].rsv, (size_t)nblk * sizeof(int4)));
            const int4* src = (const int4*)hb.data();
            for (int b = 0; b < nblk; ++b) job.t_res_mid[g0 + b] = src[b].y;
        }
        if (job.t_table || job.t_res || job.t_cbytes) {
            CK(fetch(io[s].cblk, (size_t)nblk * sizeof(ChainBlk)));
            const ChainBlk* src = (const ChainBlk*)hb.data();
            for (int b = 0; b < nblk; ++b) {
                if (job.t_table) { job.t_table[(size_t)(g0 + b) * 2] = src[b].table[0]; job.t_table[(size_t)(g0 + b) * 2 + 1] = src[b].table[1]; }
                if (job.t_res) job.t_res[g0 + b] = src[b].reservoir;
                if (job.t_cbytes) { job.t_cbytes[(size_t)(g0 + b) * 2] = (int32_t)src[b].chunk_bytes[0]; job.t_cbytes[(size_t)(g0 + b) * 2 + 1] = (int32_t)src[b].chunk_bytes[1]; }
            }
        }
    }
    // the analysis stream has nothing outstanding that the main stream does not already wait for (event 2 of the
    // last wave), so synchronising the main stream ends the call.
    unsigned long long pk[16] = {0};
    CK(cudaMemcpyAsync(pk, ctx->peakctr.p, 128, cudaMemcpyDeviceToHost, st));
    if (job.h_res_out) CK(cudaMemcpyAsync(job.h_res_out, d_res_out, (size_t)nc * 4, cudaMemcpyDeviceToHost, st));
    int ovf = 0;
    if (job.d_out) {
        CK(cudaMemcpyAsync(job.h_clip_byte_off, ctx->clip_base.p, (size_t)(nc + 1) * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(&ovf, ctx->overflow.p, 4, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaStreamSynchronize(st2));
    if (stream_out) CK(cudaStreamSynchronize(ctx->stream4));
    float t_an = 0, t_cost = 0, t_chain = 0, t_pk = 0, t_h2d = 0, t_d2h = 0;
    for (int w = 0; w < nwaves; ++w) {
        float t;
        if (wave_h2d[w]) { CK(cudaEventElapsedTime(&t, ev(w, 7), ev(w, 6))); t_h2d += t; }
        if (wave_d2h[w]) { CK(cudaEventElapsedTime(&t, ev(w, 8), ev(w, 9))); t_d2h += t; }
        CK(cudaEventElapsedTime(&t, ev(w, 0), ev(w, 1))); t_an += t;
        CK(cudaEventElapsedTime(&t, ev(w, 1), ev(w, 2))); t_cost += t;
        CK(cudaEventElapsedTime(&t, ev(w, 3), ev(w, 4))); t_chain += t;
        CK(cudaEventElapsedTime(&t, ev(w, 4), ev(w, 5))); t_pk += t;
    }
    if (getenv("MRC_TIMELINE")) {      // development aid: when did each wave's stages start and end (ms since the first)
        fprintf(stderr, "serial pass: %llu complete walks, %llu blocks stepped one by one, %llu segments through pairs, %llu by "
                "closed form; walked through: %llu not followed, %llu not anticipated, %llu straddling clips\n", pk[4], pk[5], pk[6], pk[7],
                pk[8], pk[9], pk[10]);
        if (job.shard) fprintf(stderr, "shard walked ahead of its reservoir: met the guessed trajectory after %llu segment(s) "
                               "(0: not met, or not walked ahead)\n", pk[11]);
        for (int w = 0; w < nwaves; ++w) {
            float t[6];
            for (int k = 0; k < 6; ++k) cudaEventElapsedTime(&t[k], ev(0, 0), ev(w, k));
            fprintf(stderr, "wave %2d blocks %6d  analysis %7.2f-%7.2f  cost+table -%7.2f | chain %7.2f-%7.2f  pack -%7.2f\n", w,
                    wave_g0[w + 1] - wave_g0[w], t[0], t[1], t[2], t[3], t[4], t[5]);
        }
    }
    ctx->ms[0] = t_an; ctx->ms[1] = t_chain; ctx->ms[2] = t_pk; ctx->ms[7] = t_cost;
    ctx->ms[4] = t_h2d; ctx->ms[5] = t_d2h;      // copies on the copy streams (they overlap the kernels)
    if (!job.switching) ctx->ms[3] = 0;
    else launches += 2;                // the two transient detector kernels
    ctx->counters[0] = launches;
    ctx->counters[2] = nblk_total;
    ctx->counters[4] = nwaves;
    ctx->counters[1] = (int64_t)pk[0];
    ctx->counters[3] = (int64_t)pk[4];      // complete walks taken by the serial pass (pk[5]: blocks it stepped one by one)
    ctx->counters[5] = (int64_t)pk[1]; ctx->counters[6] = (int64_t)pk[2]; ctx->counters[7] = (int64_t)pk[3];
    if (job.d_out && nblk_total == 0) for (int c = 0; c <= nc; ++c) job.h_clip_byte_off[c] = 0;
    if (job.h_copied) *job.h_copied = ovf ? 0 : copied;
    if (ovf) return fail(ctx, MRC_E_NOSPACE, "output buffer too small for the encoded batch");
    return MRC_OK;
}

int run_encode(mrc_ctx* ctx, const EncodeJob& job) {
    if (!ctx->tables_set) return fail(ctx, MRC_E_STATE, "mrc_set_tables has not been called");
    if ((job.switching || job.geom != MRC_GEO_LONG) && !ctx->switch_set)
        return fail(ctx, MRC_E_STATE, "block switching needs mrc_set_switch_tables");
    if (job.switching && !job.joint)
        return fail(ctx, MRC_E_INVALID, "block switching follows the reference's loop, which is the joint flow: create the context with joint = 1");
    const int rc = (ctx->cfg.precision == MRC_PRECISION_FP32) ? run_encode_t<float>(ctx, job) : run_encode_t<double>(ctx, job);
    if (rc != MRC_OK) {
        // an early exit may leave work queued that still reads the caller's PCM or writes the caller's output:
        // nothing of this call may be in flight once it has returned (the error message is already set)
        cudaStreamSynchronize(ctx->stream3);
        cudaStreamSynchronize(ctx->stream2);
        cudaStreamSynchronize(ctx->stream5);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->stream4);
    }
    return rc;
}

int64_t worst_case_bytes(const mrc_ctx* ctx, const int64_t* off, int nc) {
    const int L = ctx->L;
    // every mantissa escaped at 16 bits: (9+16) bits per line plus chunk headers
    const int64_t per_chunk = 4 + (6 + 4 * ctx->cfg.n_scale_bits + ctx->nb + ctx->nb * 8 + (int64_t)L * 25 + 7) / 8;
    int64_t tot = 0;
    for (int c = 0; c < nc; ++c) {
        const int64_t fr = off[c + 1] - off[c];
        tot += (int64_t)sizeof(ctx->h_header) + ((fr + L - 1) / L + 1) * 2 * per_chunk;
        if (ctx->cfg.flags & MRC_FLAG_BLOCK_SWITCHING)      // eight short blocks instead of one: 16 chunk headers
            tot += ((fr + L - 1) / L + 1) * 16 * (int64_t)(4 + (6 + 4 * ctx->cfg.n_scale_bits + 9 + 9 * 8 + 7) / 8 + 1);
    }
    return tot;
}

int64_t nominal_bytes(const mrc_ctx* ctx, const int64_t* off, int nc) {
    const int L = ctx->L;
    int64_t tot = 0;
    for (int c = 0; c < nc; ++c) {
        const int64_t fr = off[c + 1] - off[c];
        const int64_t nblk = (fr + L - 1) / L + 1;
        tot += 128 + nblk * (int64_t)(2.0 * ctx->cfg.target_bits_per_sample * L / 8.0 + 64);
        if (ctx->cfg.flags & MRC_FLAG_BLOCK_SWITCHING) tot += nblk * 160;
    }
    return tot;
}

// mrc_encode_shard: the caller's exchange callback behind a relay that remembers what was received and whether the
// result has been handed on, so that a second attempt of the same shard replays instead of exchanging again
struct ShardRelay {
    mrc_reservoir_exchange fn;
    void* user;
    int32_t r_in;
    bool got_in, gave_out;
};

int32_t shard_relay_fn(void* u, int32_t have_result, int32_t* r) {
    ShardRelay* q = (ShardRelay*)u;
    if (!have_result) {
        if (!q->got_in) {
            const int32_t rc = q->fn(q->user, 0, &q->r_in);
            if (rc != 0) return rc;
            q->got_in = true;
        }
        *r = q->r_in;
        return 0;
    }
    if (q->gave_out) return 0;
    const int32_t rc = q->fn(q->user, 1, r);
    if (rc == 0) q->gave_out = true;
    return rc;
}

}  // namespace

// ================================================================================================================
extern "C" {

int32_t mrc_version(void) { return MRC_VERSION; }

const char* mrc_last_error(const mrc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t mrc_create(const mrc_config* cfg, mrc_ctx** out) {
    mrc_ctx* ctx = nullptr;
    if (!cfg || !out) return fail(nullptr, MRC_E_INVALID, "null argument");
    *out = nullptr;
    int logL = 0;
    while ((1 << logL) < cfg->n_mdct_lines) ++logL;
    if ((1 << logL) != cfg->n_mdct_lines || logL < 7 || logL > 11)
        return fail(nullptr, MRC_E_INVALID, "n_mdct_lines must be 128, 256, 512, 1024 or 2048");
    if (cfg->n_scale_bits < 1 || cfg->n_scale_bits > 4 || cfg->n_mant_size_bits < 4 || cfg->n_mant_size_bits > 5)
        return fail(nullptr, MRC_E_INVALID, "n_scale_bits must be in 1..4 and n_mant_size_bits 4 or 5 (16-bit mantissa cap)");
    if (cfg->precision != MRC_PRECISION_FP64 && cfg->precision != MRC_PRECISION_FP32)
        return fail(nullptr, MRC_E_INVALID, "unknown precision");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, MRC_E_CUDA, "no usable CUDA device (%s); libmrc has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "bad device ordinal");
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess)
        return fail(nullptr, MRC_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    ctx = new mrc_ctx();
    ctx->cfg = *cfg;
    ctx->L = cfg->n_mdct_lines;
    ctx->logL = logL;
    // The main stream carries the serial reservoir walk: highest priority, so that its one CTA is placed as soon as
    // an SM can take it instead of queueing behind the next wave's thousands of analysis CTAs (stream2, lowest).
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, MRC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    if ((e = cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, prio_lo)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, MRC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream3, cudaStreamNonBlocking)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        cudaStreamDestroy(ctx->stream2);
        delete ctx;
        return fail(nullptr, MRC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream4, cudaStreamNonBlocking)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        cudaStreamDestroy(ctx->stream2);
        cudaStreamDestroy(ctx->stream3);
        delete ctx;
        return fail(nullptr, MRC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    if ((e = cudaStreamCreateWithPriority(&ctx->stream5, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        cudaStreamDestroy(ctx->stream2);
        cudaStreamDestroy(ctx->stream3);
        cudaStreamDestroy(ctx->stream4);
        delete ctx;
        return fail(nullptr, MRC_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    for (auto& ev : ctx->ev) cudaEventCreate(&ev);
    *out = ctx;
    return MRC_OK;
}

int32_t mrc_destroy(mrc_ctx* ctx) {
    if (!ctx) return MRC_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    for (GeoDev& g : ctx->geo) {
        for (TablesDev* d : {&g.td, &g.tf}) {
            Buf* tb[] = {&d->kbd, &d->hann, &d->tw_pre, &d->tw_post, &d->tw_fft, &d->tw_rfft, &d->tw9, &d->tw_stage, &d->bark, &d->quiet,
                         &d->bark_d, &d->quiet_d, &d->exp_tab};
            for (Buf* b : tb) release(*b);
        }
        Buf* gb[] = {&g.band_lo, &g.band_n, &g.line2band};
        for (Buf* b : gb) release(*b);
    }
    Buf* all[] = {&ctx->huff, &ctx->header, &ctx->clip_off, &ctx->clip_blk0, &ctx->clip_bytes,
                  &ctx->clip_base, &ctx->running, &ctx->overflow, &ctx->peakctr, &ctx->res_in, &ctx->res_out, &ctx->bound,
                  &ctx->clip_res, &ctx->clip_run, &ctx->q_alloc, &ctx->q_sf, &ctx->q_mant,
                  &ctx->tap_lines, &ctx->tap_smr, &ctx->tap_npk, &ctx->pcm_dev, &ctx->out_dev, &ctx->xin_dev,
                  &ctx->sb0, &ctx->peaks, &ctx->flags, &ctx->blk_start, &ctx->blk_geom, &ctx->blk_list};
    for (Buf* b : all) release(*b);
    for (auto& b : ctx->dec) release(b);
    for (auto& ws : ctx->sets) {
        Buf* wb[] = {&ws.lines, &ws.bandmax, &ws.tokens, &ws.ovs, &ws.ms, &ws.rec, &ws.pw, &ws.rsv, &ws.gmask, &ws.cblk, &ws.tab,
                     &ws.comp, &ws.segx, &ws.rin};
        for (Buf* b : wb) release(*b);
    }
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->evpool) if (ev) cudaEventDestroy(ev);
    cudaStreamSynchronize(ctx->stream2);
    cudaStreamDestroy(ctx->stream2);
    cudaStreamSynchronize(ctx->stream3);
    cudaStreamDestroy(ctx->stream3);
    cudaStreamSynchronize(ctx->stream4);
    cudaStreamDestroy(ctx->stream4);
    cudaStreamSynchronize(ctx->stream5);
    cudaStreamDestroy(ctx->stream5);
    if (ctx->h_prog) cudaFreeHost(ctx->h_prog);
    if (ctx->h_idx) cudaFreeHost(ctx->h_idx);
    for (auto& e : ctx->h_idx_ev)
        if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MRC_OK;
}

int32_t mrc_set_tables(mrc_ctx* ctx, const mrc_tables* t) {
    if (!ctx || !t) return MRC_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    if (t->n_bands < 1 || t->n_bands > MRC_CODED_BANDS)
        return fail(ctx, MRC_E_INVALID, "n_bands out of range (1..25: the reference's tables have 25 or 9 bands)");
    if (t->n_huff_tables != MRC_N_HUFF_TABLES) return fail(ctx, MRC_E_INVALID, "exactly four Huffman tables expected");
    const int L = ctx->L;
    GeoDev& g0 = ctx->geo[MRC_GEO_LONG];
    g0.a = g0.b = g0.L = L;
    {
        const int rc = set_geo_bands(ctx, g0, t->band_nlines, t->n_bands);
        if (rc != MRC_OK) return rc;
    }
    ctx->nb = t->n_bands;
    // Huffman LUTs
    HuffDev& h = ctx->h_huff;
    memset(&h, 0, sizeof h);
    for (int tb = 0; tb < MRC_N_HUFF_TABLES; ++tb) {
        for (int v = 0; v < MRC_HUFF_LUT; ++v) {
            h.len[tb][v] = t->huff_len[tb * MRC_HUFF_LUT + v];
            h.code[tb][v] = t->huff_code[tb * MRC_HUFF_LUT + v];
        }
        const int esc = t->huff_escape[tb];
        if (esc < 0 || esc >= MRC_HUFF_LUT || h.len[tb][esc] == 0)
            return fail(ctx, MRC_E_INVALID, "escape value must be a key of its table");
        // Everything downstream is sized for the trained books' code lengths: the pack kernel's bit buffer and
        // worst_case_bytes (escape + 16 raw bits <= 25 bits per line), the cost kernel's 8-bit per-line prices and
        // 16-bit totals, the decoder's 9-bit look-up.  Reject anything longer here instead of mis-encoding later.
        for (int v = 0; v < MRC_HUFF_LUT; ++v) {
            if (h.len[tb][v] > MRC_HUFF_PEEK)
                return fail(ctx, MRC_E_INVALID, "Huffman code longer than 9 bits (the code books' limit in this library)");
            if (h.len[tb][v] && (h.code[tb][v] >> h.len[tb][v]) != 0)
                return fail(ctx, MRC_E_INVALID, "Huffman code has bits above its length");
        }
        if ((int64_t)L * (16 + h.len[tb][esc]) >= 65536)
            return fail(ctx, MRC_E_INVALID, "n_mdct_lines * (16 + escape code length) must stay below 65536");
        h.esc[tb] = esc;
        h.esc_len[tb] = h.len[tb][esc];
        h.esc_code[tb] = h.code[tb][esc];
    }
    CK(ensure(ctx->huff, sizeof(HuffDev)));
    release(ctx->dec[0]);            // D_HDEC: the decoder's look-up table is rebuilt from the new books on next use
    CK(cudaMemcpyAsync(ctx->huff.p, &h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(upload_tables<double>(ctx, g0, g0.td, g0.tbd, t->kbd_window, t->hann_window, t->bark, t->quiet_intensity));
    CK(upload_tables<float>(ctx, g0, g0.tf, g0.tbf, t->kbd_window, t->hann_window, t->bark, t->quiet_intensity));
    set_geo_budgets(ctx, g0);
    g0.set = true;
    ctx->switch_set = false;         // geometries 1..3 were derived for the previous tables
    const mrc_config& c = ctx->cfg;
    ctx->no_tables = (c.flags & MRC_FLAG_NO_CHAIN_TABLES) != 0;
    if (const char* e = getenv("MRC_CHAIN_TABLE_MIN_BLOCKS")) ctx->tab_min_blocks = std::max(1, atoi(e));   // test knob
    if (const char* e = getenv("MRC_CHAIN_SEGMENT_BLOCKS")) ctx->seg_blocks = std::max(0, atoi(e));         // test knob
    if (ctx->cp.max_mant_bits != 16) return fail(ctx, MRC_E_INVALID, "only a 16-bit mantissa cap is supported");
    // .pac header template (pacfileThem.py:592-613); numSamples is patched per clip by the pack kernel
    uint8_t* hd = ctx->h_header;
    memset(hd, 0, sizeof ctx->h_header);
    memcpy(hd, "PAC ", 4);
    auto put32 = [&](int o, uint32_t v) { for (int i = 0; i < 4; ++i) hd[o + i] = (uint8_t)(v >> (8 * i)); };
    auto put16 = [&](int o, uint32_t v) { for (int i = 0; i < 2; ++i) hd[o + i] = (uint8_t)(v >> (8 * i)); };
    put32(4, (uint32_t)c.sample_rate); put16(8, 2); put32(10, 0); put32(14, (uint32_t)L);
    put16(18, (uint32_t)c.n_scale_bits); put16(20, (uint32_t)c.n_mant_size_bits); put32(22, (uint32_t)t->n_bands);
    for (int b = 0; b < t->n_bands; ++b) put16(26 + 2 * b, (uint32_t)t->band_nlines[b]);
    CK(ensure(ctx->header, sizeof ctx->h_header));
    CK(cudaMemcpyAsync(ctx->header.p, hd, sizeof ctx->h_header, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->tables_set = true;
    return MRC_OK;
}

int32_t mrc_host_alloc(void** p, int64_t bytes) {
    if (!p || bytes < 0) return MRC_E_INVALID;
    return cudaMallocHost(p, (size_t)std::max<int64_t>(bytes, 1)) == cudaSuccess ? MRC_OK : MRC_E_CUDA;
}
int32_t mrc_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? MRC_OK : MRC_E_CUDA; }

int32_t mrc_encode_batch_device(mrc_ctx* ctx, const int16_t* d_pcm, const int64_t* clip_frame_offsets,
                                int32_t n_clips, uint8_t* d_out, int64_t out_cap, int64_t* clip_byte_offsets) {
    if (!ctx) return MRC_E_INVALID;
    if (!clip_frame_offsets || !clip_byte_offsets || n_clips < 0 || (!d_out && out_cap > 0))
        return fail(ctx, MRC_E_INVALID, "bad argument");
    cudaSetDevice(ctx->cfg.device);
    EncodeJob job;
    job.d_pcm = d_pcm; job.h_clip_off = clip_frame_offsets; job.n_clips = n_clips;
    job.joint = ctx->cfg.joint; job.flush_nonjoint = true;
    job.switching = (ctx->cfg.flags & MRC_FLAG_BLOCK_SWITCHING) != 0;
    job.d_out = d_out; job.out_cap = out_cap; job.h_clip_byte_off = clip_byte_offsets;
    cudaEventRecord(ctx->ev[4], ctx->stream);
    const int rc = run_encode(ctx, job);
    cudaEventRecord(ctx->ev[5], ctx->stream);
    cudaEventSynchronize(ctx->ev[5]);
    float t = 0;
    cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]);
    ctx->ms[6] = t;
    return rc;
}

int32_t mrc_encode_batch(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                         uint8_t* out, int64_t out_cap, int64_t* clip_byte_offsets) {
    if (!ctx) return MRC_E_INVALID;
    if (!clip_frame_offsets || !clip_byte_offsets || n_clips < 0) return fail(ctx, MRC_E_INVALID, "bad argument");
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    const int64_t frames = clip_frame_offsets[n_clips] - clip_frame_offsets[0];
    if (clip_frame_offsets[0] != 0) return fail(ctx, MRC_E_INVALID, "clip_frame_offsets[0] must be 0");
    CK(ensure(ctx->pcm_dev, (size_t)std::max<int64_t>(frames, 1) * 4));
    // device staging for the bitstream: nominal size with head-room, never more than the worst case
    int64_t cap = std::min(worst_case_bytes(ctx, clip_frame_offsets, n_clips),
                           std::max<int64_t>(2 * nominal_bytes(ctx, clip_frame_offsets, n_clips), out_cap));
    int64_t copied = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        copied = 0;
        CK(ensure(ctx->out_dev, (size_t)cap));
        CK(cudaEventRecord(ctx->ev[4], st));
        EncodeJob job;
        job.d_pcm = (const int16_t*)ctx->pcm_dev.p; job.h_clip_off = clip_frame_offsets; job.n_clips = n_clips;
        job.h_pcm = frames > 0 ? pcm : nullptr;             // uploaded wave by wave, overlapped with the kernels
        job.joint = ctx->cfg.joint; job.flush_nonjoint = true;
        job.switching = (ctx->cfg.flags & MRC_FLAG_BLOCK_SWITCHING) != 0;
        job.d_out = (uint8_t*)ctx->out_dev.p; job.out_cap = cap; job.h_clip_byte_off = clip_byte_offsets;
        job.h_out = out; job.h_out_cap = out ? out_cap : 0; job.h_copied = &copied;
        const int rc = run_encode(ctx, job);
        cudaStreamSynchronize(ctx->stream4);                // whatever was queued for `out` has landed (or failed with rc)
        if (rc == MRC_E_NOSPACE && attempt == 0) {          // staging too small: retry once at the worst case
            cap = worst_case_bytes(ctx, clip_frame_offsets, n_clips);
            continue;
        }
        if (rc != MRC_OK) return rc;
        break;
    }
    const int64_t total = clip_byte_offsets[n_clips];
    if (total > out_cap) return fail(ctx, MRC_E_NOSPACE, "output buffer too small (clip_byte_offsets holds the sizes)");
    CK(cudaEventRecord(ctx->ev[7], st));
    if (total > copied)
        CK(cudaMemcpyAsync(out + copied, (const uint8_t*)ctx->out_dev.p + copied, (size_t)(total - copied),
                           cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[5], st));
    CK(cudaStreamSynchronize(st));
    float t = 0;
    cudaEventElapsedTime(&t, ctx->ev[7], ctx->ev[5]); ctx->ms[5] += t;      // + the tail the waves had not drained
    cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]); ctx->ms[6] = t;
    return MRC_OK;
}

static int32_t encode_shard_impl(mrc_ctx* ctx, bool on_device, const int16_t* pcm, int64_t pcm_frame0, int64_t pcm_frames,
                                 int64_t total_frames, int64_t first_block, int64_t n_blocks, int32_t is_first,
                                 int32_t is_last, uint8_t* out, int64_t out_cap, int64_t* out_bytes,
                                 mrc_reservoir_exchange exchange, void* user) {
    if (!ctx) return MRC_E_INVALID;
    if (!out_bytes || !exchange || pcm_frames < 0 || total_frames < 0 || first_block < 0 || n_blocks < 0 || (!pcm && pcm_frames > 0))
        return fail(ctx, MRC_E_INVALID, "bad argument");
    if (ctx->cfg.flags & MRC_FLAG_BLOCK_SWITCHING)
        return fail(ctx, MRC_E_INVALID, "sharding a stream by block range is built for long blocks only");
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    const int L = ctx->L;
    const int64_t nblk_stream = (total_frames + L - 1) / L;
    if (first_block + n_blocks > nblk_stream) return fail(ctx, MRC_E_INVALID, "shard runs past the end of the stream");
    if (is_last && first_block + n_blocks != nblk_stream) return fail(ctx, MRC_E_INVALID, "the last shard must end with the stream");
    const int64_t need_lo = std::max<int64_t>(first_block - 1, 0) * L;
    const int64_t need_hi = std::min<int64_t>((first_block + n_blocks) * L, total_frames);
    if (n_blocks > 0 && (pcm_frame0 > need_lo || pcm_frame0 + pcm_frames < need_hi))
        return fail(ctx, MRC_E_INVALID, "pcm does not cover the shard's blocks and their halo");
    *out_bytes = 0;
    if (n_blocks + (is_last ? 1 : 0) == 0) {
        // nothing to encode: pass the reservoir through; a first shard without blocks still owes the file header
        int32_t r = 0;
        if (exchange(user, 0, &r) != 0 || exchange(user, 1, &r) != 0)
            return fail(ctx, MRC_E_STATE, "reservoir exchange callback failed");
        if (is_first) {
            const int hb = ctx->cp.header_bytes;
            *out_bytes = hb;
            if (hb > out_cap || !out) return fail(ctx, MRC_E_NOSPACE, "output buffer too small (out_bytes holds the size)");
            uint8_t hd[sizeof ctx->h_header];
            memcpy(hd, ctx->h_header, sizeof hd);
            int64_t ns = total_frames;
            if (ns % L == 0) ns += L;                // Q9 (pacfileThem.py:595-597)
            for (int i = 0; i < 4; ++i) hd[10 + i] = (uint8_t)((uint64_t)ns >> (8 * i));
            if (on_device) CK(cudaMemcpy(out, hd, (size_t)hb, cudaMemcpyHostToDevice));
            else memcpy(out, hd, (size_t)hb);
        }
        return MRC_OK;
    }
    if (!on_device) {
        CK(ensure(ctx->pcm_dev, (size_t)std::max<int64_t>(pcm_frames, 1) * 4));
        if (pcm_frames > 0)
            CK(cudaMemcpyAsync(ctx->pcm_dev.p, pcm, (size_t)pcm_frames * 4, cudaMemcpyHostToDevice, st));
    }
    const int64_t off[2] = {0, total_frames};
    int64_t boff[2] = {0, 0};
    const int64_t shard_off[2] = {0, (n_blocks + (is_last ? 1 : 0)) * (int64_t)L};      // for the size estimates only
    int64_t cap = std::min(worst_case_bytes(ctx, shard_off, 1), std::max<int64_t>(2 * nominal_bytes(ctx, shard_off, 1), out_cap));
    ShardRelay relay = {exchange, user, 0, false, false};     // a retry (staging too small) must not serve the neighbours twice
    if (on_device) cap = out_cap;                              // the caller's device buffer is the staging
    for (int attempt = 0; attempt < (on_device ? 1 : 2); ++attempt) {
        if (!on_device) CK(ensure(ctx->out_dev, (size_t)cap));
        CK(cudaEventRecord(ctx->ev[4], st));
        EncodeJob job;
        job.d_pcm = on_device ? pcm : (const int16_t*)ctx->pcm_dev.p; job.h_clip_off = off; job.n_clips = 1;
        job.joint = ctx->cfg.joint; job.flush_nonjoint = is_last != 0;
        job.shard = true; job.shard_first = first_block; job.shard_blocks = n_blocks; job.pcm_frame0 = pcm_frame0;
        job.shard_header = is_first != 0;
        job.exchange = shard_relay_fn; job.exchange_user = &relay;
        job.d_out = on_device ? out : (uint8_t*)ctx->out_dev.p; job.out_cap = cap; job.h_clip_byte_off = boff;
        const int rc = run_encode(ctx, job);
        if (rc == MRC_E_NOSPACE && attempt == 0 && !on_device) { cap = worst_case_bytes(ctx, shard_off, 1); continue; }
        if (rc != MRC_OK) return rc;
        break;
    }
    *out_bytes = boff[1];
    if (boff[1] > out_cap || !out) return fail(ctx, MRC_E_NOSPACE, "output buffer too small (out_bytes holds the size)");
    if (!on_device && boff[1] > 0) CK(cudaMemcpyAsync(out, ctx->out_dev.p, (size_t)boff[1], cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(ctx->ev[5], st));
    CK(cudaStreamSynchronize(st));
    float t = 0;
    cudaEventElapsedTime(&t, ctx->ev[4], ctx->ev[5]); ctx->ms[6] = t;
    return MRC_OK;
}

int32_t mrc_encode_shard(mrc_ctx* ctx, const int16_t* pcm, int64_t pcm_frame0, int64_t pcm_frames, int64_t total_frames,
                         int64_t first_block, int64_t n_blocks, int32_t is_first, int32_t is_last, uint8_t* out,
                         int64_t out_cap, int64_t* out_bytes, mrc_reservoir_exchange exchange, void* user) {
    return encode_shard_impl(ctx, false, pcm, pcm_frame0, pcm_frames, total_frames, first_block, n_blocks, is_first, is_last,
                             out, out_cap, out_bytes, exchange, user);
}

int32_t mrc_encode_shard_device(mrc_ctx* ctx, const int16_t* d_pcm, int64_t pcm_frame0, int64_t pcm_frames,
                                int64_t total_frames, int64_t first_block, int64_t n_blocks, int32_t is_first,
                                int32_t is_last, uint8_t* d_out, int64_t out_cap, int64_t* out_bytes,
                                mrc_reservoir_exchange exchange, void* user) {
    return encode_shard_impl(ctx, true, d_pcm, pcm_frame0, pcm_frames, total_frames, first_block, n_blocks, is_first, is_last,
                             d_out, out_cap, out_bytes, exchange, user);
}

int32_t mrc_stage_analysis(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                           double* mdct_lines, int32_t* overall_scale, int32_t* ms_switch, double* smr,
                           int32_t* n_peaks) {
    if (!ctx || !clip_frame_offsets || n_clips < 0) return MRC_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    const int64_t frames = clip_frame_offsets[n_clips];
    CK(ensure(ctx->pcm_dev, (size_t)std::max<int64_t>(frames, 1) * 4));
    if (frames > 0) CK(cudaMemcpyAsync(ctx->pcm_dev.p, pcm, (size_t)frames * 4, cudaMemcpyHostToDevice, ctx->stream));
    EncodeJob job;       // the whole-clip taps always follow the long-block flow (block counts are fixed up front)
    job.d_pcm = (const int16_t*)ctx->pcm_dev.p; job.h_clip_off = clip_frame_offsets; job.n_clips = n_clips;
    job.joint = ctx->cfg.joint; job.flush_nonjoint = true; job.need_quant = false;
    job.t_lines = mdct_lines; job.t_ovs = overall_scale; job.t_ms = ms_switch; job.t_smr = smr; job.t_npk = n_peaks;
    return run_encode(ctx, job);
}

int32_t mrc_stage_alloc_quant(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets,
                              int32_t n_clips, int32_t* bit_alloc, int32_t* scale_factor, int32_t* mantissa,
                              int32_t* huff_table, int32_t* reservoir, int32_t* chunk_bytes) {
    if (!ctx || !clip_frame_offsets || n_clips < 0) return MRC_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    const int64_t frames = clip_frame_offsets[n_clips];
    CK(ensure(ctx->pcm_dev, (size_t)std::max<int64_t>(frames, 1) * 4));
    if (frames > 0) CK(cudaMemcpyAsync(ctx->pcm_dev.p, pcm, (size_t)frames * 4, cudaMemcpyHostToDevice, ctx->stream));
    EncodeJob job;
    job.d_pcm = (const int16_t*)ctx->pcm_dev.p; job.h_clip_off = clip_frame_offsets; job.n_clips = n_clips;
    job.joint = ctx->cfg.joint; job.flush_nonjoint = true;
    job.t_alloc = bit_alloc; job.t_sf = scale_factor; job.t_mant = mantissa; job.t_table = huff_table;
    job.t_res = reservoir; job.t_cbytes = chunk_bytes;
    return run_encode(ctx, job);
}

int32_t mrc_encode_block_ab(mrc_ctx* ctx, const double* data, int32_t a, int32_t b, int32_t joint, int32_t* reservoir,
                            int32_t* scale_factor, int32_t* bit_alloc, int32_t* mantissa, int32_t* overall_scale,
                            int32_t* ms_switch, int32_t* huff_table, int32_t* chunk_bytes) {
    if (!ctx || !data || !reservoir) return MRC_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    const int Lc = ctx->L;
    if ((a != Lc && a != MRC_SHORT) || (b != Lc && b != MRC_SHORT))
        return fail(ctx, MRC_E_INVALID, "window halves a and b must each be n_mdct_lines or 128");
    const int q = (a != Lc ? 2 : 0) | (b != Lc ? 1 : 0);
    const int N = a + b;
    const bool mono = (joint & 4) != 0;       // one channel (codingParams.nChannels = 1): run it as both channels of an
    if (mono && (joint & 1))                  // independent-channel block and keep channel 0 and the reservoir after it
        return fail(ctx, MRC_E_INVALID, "a single channel has no joint (M/S) flow");
    CK(ensure(ctx->xin_dev, (size_t)2 * N * 8));
    CK(cudaMemcpyAsync(ctx->xin_dev.p, data, (size_t)(mono ? 1 : 2) * N * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (mono)
        CK(cudaMemcpyAsync((double*)ctx->xin_dev.p + N, data, (size_t)N * 8, cudaMemcpyHostToDevice, ctx->stream));
    const int64_t off[2] = {0, N};
    int32_t res_in = *reservoir, res_out = 0;
    if (mono) {
        const int nbq = ctx->geo[q].nb, Lq = ctx->geo[q].L;
        std::vector<int32_t> sf2(2 * nbq), ba2(2 * nbq), mant2(2 * (size_t)Lq);
        int32_t ovs4[4], ht2[2], cb2[2], mid = 0;
        EncodeJob mj;
        mj.d_xin = (const double*)ctx->xin_dev.p; mj.h_clip_off = off; mj.n_clips = 1;
        mj.geom = q;
        mj.joint = 0; mj.no_huff = (joint & 2) ? 1 : 0; mj.flush_nonjoint = false;
        mj.h_res_in = &res_in; mj.h_res_out = &res_out;
        mj.t_alloc = ba2.data(); mj.t_sf = sf2.data(); mj.t_mant = mant2.data(); mj.t_table = ht2;
        mj.t_cbytes = cb2; mj.t_ovs = ovs4; mj.t_res_mid = &mid;
        const int rc = run_encode(ctx, mj);
        if (rc != MRC_OK) return rc;
        if (scale_factor) memcpy(scale_factor, sf2.data(), (size_t)nbq * 4);
        if (bit_alloc) memcpy(bit_alloc, ba2.data(), (size_t)nbq * 4);
        if (mantissa) memcpy(mantissa, mant2.data(), (size_t)Lq * 4);
        if (overall_scale) overall_scale[0] = ovs4[0];
        if (huff_table) huff_table[0] = ht2[0];
        if (chunk_bytes) chunk_bytes[0] = cb2[0];
        *reservoir = mid;
        return MRC_OK;
    }
    EncodeJob job;
    job.d_xin = (const double*)ctx->xin_dev.p; job.h_clip_off = off; job.n_clips = 1;
    job.geom = q;
    job.joint = (joint & 1) ? 1 : 0; job.no_huff = (joint & 2) ? 1 : 0; job.flush_nonjoint = false;
    job.h_res_in = &res_in; job.h_res_out = &res_out;
    job.t_alloc = bit_alloc; job.t_sf = scale_factor; job.t_mant = mantissa; job.t_table = huff_table;
    job.t_cbytes = chunk_bytes; job.t_ovs = overall_scale; job.t_ms = ms_switch;
    const int rc = run_encode(ctx, job);
    if (rc == MRC_OK) *reservoir = res_out;
    return rc;
}

int32_t mrc_encode_block(mrc_ctx* ctx, const double* data, int32_t joint, int32_t* reservoir,
                         int32_t* scale_factor, int32_t* bit_alloc, int32_t* mantissa, int32_t* overall_scale,
                         int32_t* ms_switch, int32_t* huff_table, int32_t* chunk_bytes) {
    if (!ctx) return MRC_E_INVALID;
    return mrc_encode_block_ab(ctx, data, ctx->L, ctx->L, joint, reservoir, scale_factor, bit_alloc, mantissa,
                               overall_scale, ms_switch, huff_table, chunk_bytes);
}

int32_t mrc_set_switch_tables(mrc_ctx* ctx, const mrc_block_tables* t3, const double* sos, int32_t n_sections,
                              double t0, double t1) {
    if (!ctx || !t3 || !sos) return MRC_E_INVALID;
    if (!ctx->tables_set) return fail(ctx, MRC_E_STATE, "mrc_set_tables has not been called");
    if (n_sections < 1 || n_sections > MRC_MAX_SOS) return fail(ctx, MRC_E_INVALID, "1..16 second-order sections");
    if (ctx->L < 4 * MRC_SHORT) return fail(ctx, MRC_E_INVALID, "block switching needs n_mdct_lines >= 512");
    if (ctx->L != 1024) return fail(ctx, MRC_E_INVALID, "block switching is built for n_mdct_lines = 1024 (transition blocks of 576 lines)");
    cudaSetDevice(ctx->cfg.device);
    const int Lc = ctx->L;
    const int want[3][2] = {{Lc, MRC_SHORT}, {MRC_SHORT, Lc}, {MRC_SHORT, MRC_SHORT}};
    for (int i = 0; i < 3; ++i) {
        const mrc_block_tables& t = t3[i];
        if (t.a != want[i][0] || t.b != want[i][1])
            return fail(ctx, MRC_E_INVALID, "block tables must come in the order (L,128), (128,L), (128,128)");
        GeoDev& g = ctx->geo[i + 1];
        g.a = t.a; g.b = t.b; g.L = (t.a + t.b) / 2;
        const int rc = set_geo_bands(ctx, g, t.band_nlines, t.n_bands);
        if (rc != MRC_OK) return rc;
        CK(upload_tables<double>(ctx, g, g.td, g.tbd, t.window, t.hann_window, t.bark, t.quiet_intensity));
        CK(upload_tables<float>(ctx, g, g.tf, g.tbf, t.window, t.hann_window, t.bark, t.quiet_intensity));
        set_geo_budgets(ctx, g);
        g.set = true;
    }
    ctx->sos.n = n_sections;
    ctx->sos.t0 = t0; ctx->sos.t1 = t1;
    for (int s2 = 0; s2 < n_sections; ++s2) {
        const double* r = sos + 6 * s2;
        if (r[3] != 1.0) return fail(ctx, MRC_E_INVALID, "second-order sections must be normalised (a0 = 1)");
        ctx->sos.c[s2][0] = r[0]; ctx->sos.c[s2][1] = r[1]; ctx->sos.c[s2][2] = r[2];
        ctx->sos.c[s2][3] = r[4]; ctx->sos.c[s2][4] = r[5];
    }
    ctx->switch_set = true;
    return MRC_OK;
}

int32_t mrc_detect_transients(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                              uint8_t* flags, int32_t* block_ab, int32_t block_cap, int32_t* clip_block_offsets) {
    if (!ctx || !clip_frame_offsets || n_clips < 0) return MRC_E_INVALID;
    if (!ctx->switch_set) return fail(ctx, MRC_E_STATE, "block switching needs mrc_set_switch_tables");
    cudaSetDevice(ctx->cfg.device);
    const int64_t frames = clip_frame_offsets[n_clips];
    CK(ensure(ctx->pcm_dev, (size_t)std::max<int64_t>(frames, 1) * 4));
    if (frames > 0) CK(cudaMemcpyAsync(ctx->pcm_dev.p, pcm, (size_t)frames * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int32_t> blk0;
    std::vector<int64_t> bstart;
    std::vector<uint8_t> bgeom, fl;
    const int rc = plan_switched_blocks(ctx, (const int16_t*)ctx->pcm_dev.p, clip_frame_offsets, n_clips, ctx->stream,
                                        blk0, bstart, bgeom, &fl);
    if (rc != MRC_OK) return rc;
    if (flags) memcpy(flags, fl.data(), fl.size());
    if (clip_block_offsets) memcpy(clip_block_offsets, blk0.data(), (size_t)(n_clips + 1) * 4);
    if (block_ab) {
        if ((int64_t)bgeom.size() > block_cap) return fail(ctx, MRC_E_NOSPACE, "block_ab too small (clip_block_offsets holds the counts)");
        for (size_t i = 0; i < bgeom.size(); ++i) {
            block_ab[2 * i] = (bgeom[i] & 2) ? MRC_SHORT : ctx->L;
            block_ab[2 * i + 1] = (bgeom[i] & 1) ? MRC_SHORT : ctx->L;
        }
    }
    return MRC_OK;
}

int32_t mrc_mantissa_histogram(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                               int32_t prior_max, int64_t* hist, int32_t* max_out, int32_t* reset_out) {
    if (!ctx) return MRC_E_INVALID;
    if (!clip_frame_offsets || n_clips < 0 || !hist || !max_out || !reset_out) return fail(ctx, MRC_E_INVALID, "bad argument");
    if (ctx->cfg.joint) return fail(ctx, MRC_E_INVALID, "the training front end encodes independent channels: create the context with joint = 0");
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t st = ctx->stream;
    const int L = ctx->L;
    const int64_t frames = clip_frame_offsets[n_clips];
    int64_t nblk = 0;
    for (int c = 0; c < n_clips; ++c) nblk += (clip_frame_offsets[c + 1] - clip_frame_offsets[c] + L - 1) / L;
    if (nblk > WAVE_BLOCKS) return fail(ctx, MRC_E_INVALID, "too many blocks in one histogram call (split the corpus)");
    memset(hist, 0, 65536 * sizeof(int64_t));
    *max_out = prior_max;
    *reset_out = 0;
    if (nblk == 0) return MRC_OK;
    CK(ensure(ctx->pcm_dev, (size_t)std::max<int64_t>(frames, 1) * 4));
    CK(cudaMemcpyAsync(ctx->pcm_dev.p, pcm, (size_t)frames * 4, cudaMemcpyHostToDevice, st));
    EncodeJob job;
    job.d_pcm = (const int16_t*)ctx->pcm_dev.p; job.h_clip_off = clip_frame_offsets; job.n_clips = n_clips;
    job.joint = 0; job.no_huff = 1; job.flush_nonjoint = false;       // the script counts the blocks of its own loop only
    job.dev_qtaps = true;
    const int rc = run_encode(ctx, job);
    if (rc != MRC_OK) return rc;
    const int ncalls = (int)nblk * 2;
    const size_t cmax_bytes = ((size_t)ncalls * 4 + 15) & ~(size_t)15;
    CK(ensure(ctx->dec[15], cmax_bytes + 65536 * 8));
    int32_t* d_cmax = (int32_t*)ctx->dec[15].p;
    unsigned long long* d_hist = (unsigned long long*)((char*)ctx->dec[15].p + cmax_bytes);
    launch_callmax(st, L, ncalls, (const uint8_t*)ctx->q_alloc.p, (const uint16_t*)ctx->q_mant.p,
                   (const uint8_t*)ctx->geo[0].line2band.p, d_cmax);
    std::vector<int32_t> cmax(ncalls);
    CK(cudaMemcpyAsync(cmax.data(), d_cmax, (size_t)ncalls * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // the last call whose maximum is a new record: its first record-breaking mantissa resets all smaller counts
    int gmax = prior_max, first_call = 0, thr = -2;
    for (int c = 0; c < ncalls; ++c)
        if (cmax[c] > gmax) { first_call = c; thr = gmax; gmax = cmax[c]; *reset_out = 1; }
    CK(cudaMemsetAsync(d_hist, 0, 65536 * 8, st));
    launch_hist(st, L, ncalls, first_call, thr, (const uint8_t*)ctx->q_alloc.p, (const uint16_t*)ctx->q_mant.p,
                (const uint8_t*)ctx->geo[0].line2band.p, d_hist);
    CK(cudaMemcpyAsync(hist, d_hist, 65536 * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    *max_out = gmax;
    return MRC_OK;
}

int32_t mrc_measure_peaks(mrc_ctx* ctx, double* out4) {
    if (!ctx || !out4) return MRC_E_INVALID;
    cudaSetDevice(ctx->cfg.device);
    return measure_peaks(ctx->stream, out4) == 0 ? MRC_OK : fail(ctx, MRC_E_CUDA, "peak micro-benchmark failed");
}

int32_t mrc_last_timing(const mrc_ctx* ctx, double* ms8, int64_t* counters8) {
    if (!ctx) return MRC_E_INVALID;
    if (ms8) memcpy(ms8, ctx->ms, sizeof ctx->ms);
    if (counters8) memcpy(counters8, ctx->counters, sizeof ctx->counters);
    return MRC_OK;
}

}  // extern "C"

#include "mrc_api_decode.inc"
