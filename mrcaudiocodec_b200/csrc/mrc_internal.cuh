// mrc_internal.cuh -- shared declarations for the libmrc.so kernels (sm_100a).
//
// Data layout in HBM (per encode wave of nblk blocks, L = n_mdct_lines, NB = n_bands <= 25):
//   analysis kernel -> hand-off      lines   [nblk][2][L]   real  selected (M|L, S|R per band), scaled by 2^overall
//                                    bandmax [nblk][2][32]  real  max |line| per band of the selected lines
//                                    tokens  [nblk][768]    u16   water-filling grant order (band | level<<8)
//                                    ovs     [nblk][4]      u8    overall scale factors L,R,M,S
//                                    ms      [nblk]         u32   ms_switch bit mask
//   cost kernel -> chain / finish    rec     [nblk][MRC_REC_BYTES]  per grant token (in grant order): band/level/
//                                    nLines and, as prefix sums over the order, bits spent and Huffman cost under
//                                    the four books; per 32-token chunk the running maximum of (bits spent + nLines)
//                                    pw      [nblk][MRC_PW_BYTES]   the same prefix sums for the bits actually written
//   table kernel -> chain            tab     [nblk][2][tabw] i32    reservoir map R_in -> R_out (single-stream path)
//   chain kernel -> finish           rsv     [nblk] int4            reservoir before group 0 / group 1 / after
//   finish, offsets -> pack          gmask   [nblk][32] u32  granted tokens per 32-token chunk
//                                    cblk    [nblk] ChainBlk  table ids, chunk sizes and offsets, reservoir
//   pack kernel (taps only)          alloc,sf [nblk][2][32] u8 ; mant [nblk][2][L] u16
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mrc.h"

// -DMRC_DEBUG_ASSERTS (make debug -> libmrc_debug.so, run by scripts/debug_asserts.py): index checks at the places
// where a wrong size would write outside a buffer -- the carving of the analysis kernel's shared memory, the token
// slots of the serial walk, the pack kernel's bit buffer, the reservoir-map rows.  compute-sanitizer is closed on the
// GPU pool this was developed on, so this build is the memory-safety net; it is never the shipped library.
#ifdef MRC_DEBUG_ASSERTS
#include <cstdio>
#define MRC_ASSERT(cond)                                                                                            \
    do {                                                                                                            \
        if (!(cond)) {                                                                                              \
            printf("MRC_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                                               \
            __trap();                                                                                               \
        }                                                                                                           \
    } while (0)
#else
#define MRC_ASSERT(cond)
#endif

#define MRC_TOK_STRIDE 768      // >= 2*25*15 grant tokens per block
#define MRC_MAX_LEVELS 15       // grants per band: 0->2, then +1 up to 16 bits
#define MRC_BSTRIDE 32          // band stride in per-band arrays
#define MRC_CODED_BANDS 25      // most bands per channel the grant-token machinery holds (25 * 15 <= MRC_GROUP_SLOTS)

// ---- per-block record the cost kernel hands to the chain kernel (one TMA bulk copy per block) ---------------
// Token slots: a joint block sorts the 2*nb bands of both channels into one grant order (slots 0 .. 2*nb*15-1);
// a non-joint block has one order per channel: channel 0 in slots 0.., channel 1 in slots MRC_GROUP_SLOTS...
// Unused slots hold 0xffffffff.
#define MRC_NSLOT 768
#define MRC_GROUP_SLOTS 384
#define MRC_NCHUNK (MRC_NSLOT / 32)
#define MRC_GROUP_CHUNKS (MRC_GROUP_SLOTS / 32)
#define MRC_REC_TN 0                                   // u32 [768]   token (band | level<<8) | nLines<<16
#define MRC_REC_CP (MRC_NSLOT * 4)                     // u32 [768]   bits spent if every earlier token of the group is
                                                       //             granted: channel 0 | channel 1 << 16
#define MRC_REC_PC (MRC_REC_CP + MRC_NSLOT * 4)        // uint4 [768] Huffman cost of the four books under the same
                                                       //             assumption: {ch0 b0|b1<<16, ch0 b2|b3<<16, ch1 .., ch1 ..}
                                                       //             (sums of 16-bit fields carried mod 2^32)
#define MRC_REC_MX (MRC_REC_PC + MRC_NSLOT * 16)       // i32 [32]    per 32-token chunk: running max of (bits spent + nLines)
#define MRC_REC_BYTES (MRC_REC_MX + 32 * 4)            // 18560, a multiple of 16
// words MRC_NCHUNK.. of the MX array carry the block's own geometry to the serial kernels (block switching: the
// budget and the band count differ between long, transition and short blocks)
#define MRC_MX_K 24                                    // integer part of the bit budget before the reservoir
#define MRC_MX_FRAC 25                                 // 1 if the budget has a fractional part
#define MRC_MX_NB 26                                   // scale factor bands of the block
#define MRC_MX_MINNL 27                                // lines of its narrowest band: no grant fits below that
#define MRC_PW_BYTES (MRC_NSLOT * 16)                  // uint4 [768] like MRC_REC_PC for the bits actually written (Q4)

struct ChainBlk {                // what the chain kernel decides per block (32 bytes)
    long long chunk_off[2];      // byte offset of each channel chunk (its <L prefix) inside the clip's .pac
    unsigned int chunk_bytes[2]; // payload bytes of each channel chunk
    int reservoir;               // codingParams.bitReservoir after the block
    unsigned char table[2];      // Huffman table id per channel (15 = none)
    unsigned char pad[2];
};

template <typename T> struct cpx { T x, y; };

// Block geometries: window halves a (previous block) and b (this block), each nMDCTLines or 128.
// geometry id = 2*(a short) + (b short) = the two blksw bits of the chunk header (pacfileThem.py:720-721).
#define MRC_GEO_LONG 0
#define MRC_GEO_START 1                          // (L, 128)
#define MRC_GEO_STOP 2                           // (128, L)
#define MRC_GEO_SHORT 3                          // (128, 128)
#define MRC_N_GEO 4
#define MRC_SHORT 128

template <typename T>
struct DevTables {
    int L, logL, nb, sample_rate, fstep;         // L = (a+b)/2 lines; fstep = sample_rate // (a+b)   (Q2)
    int a, b, geom;
    int rot;                                     // (a-b)/4: MDCT phase n0 = (b+1)/2 as a rotation of the standard one
    int logLtab;                                 // tw_fft holds exp(-2*pi*j*k/2^logLtab), k < 2^(logLtab-1)
    const cpx<T>* tw9;                           // [L]    exp(-2*pi*j*m/L) for the radix-9 pass (L = 9 * 2^p only)
    cpx<T> w9[9];                                //        exp(-2*pi*j*m/9)
    const T* kbd;                                // [2L]   TransitionWindow(a, b)
    const T* hann;                               // [2L]
    const cpx<T>* tw_pre;                        // [L/2]  exp(-j*pi*(4n+1)/(4L))
    const cpx<T>* tw_post;                       // [L/2]  exp(-j*pi*k/L)
    const cpx<T>* tw_fft;                        // [L/2]  exp(-2*pi*j*k/L)
    const cpx<T>* tw_stage;                      // per-stage twiddles of the L- and L/2-point transforms (mrc_fft.cuh, power-of-two L)
    const cpx<T>* tw_rfft;                       // [L]    exp(-2*pi*j*k/(2L))
    const T* bark;                               // [L]
    const T* quiet;                              // [L]
    const double* bark_d;                        // [L]   the same two tables in double (spreading is always fp64)
    const double* quiet_d;                       // [L]
    const double* exp_tab;                       // [64]  2^(j/64)
    const int* band_lo;                          // [nb]
    const int* band_n;                           // [nb]
    const uint8_t* line2band;                    // [L]
    // the band table by value as well (kernel parameters live in the constant bank: a warp-uniform lookup costs a few
    // cycles instead of a dependent trip to L2) -- used by the analysis kernel's per-band loops
    uint16_t c_band_lo[MRC_BSTRIDE], c_band_n[MRC_BSTRIDE];
};

struct HuffDev {
    uint8_t len[MRC_N_HUFF_TABLES][MRC_HUFF_LUT + 3];
    uint16_t code[MRC_N_HUFF_TABLES][MRC_HUFF_LUT + 3];
    int32_t esc[MRC_N_HUFF_TABLES];
    int32_t esc_len[MRC_N_HUFF_TABLES];
    int32_t esc_code[MRC_N_HUFF_TABLES];
};

struct ClipMap {
    const int64_t* clip_off;     // [n_clips+1] frame offsets into pcm            (device)
    const int32_t* clip_blk0;    // [n_clips+1] first global block of every clip   (device)
    int n_clips;
    // block switching: blocks differ in size, so position and geometry are per block (null otherwise)
    const int64_t* blk_start;    // [nblk_total] first frame of the block's new samples, relative to its clip
    const uint8_t* blk_geom;     // [nblk_total] MRC_GEO_*
    // the kernels of one geometry run over a list of wave-local block indices (null: all blocks of the wave)
    const int32_t* list;
    // one stream sharded by block range (single-clip jobs only; both 0 otherwise): the job's first block is block
    // blk_base of the stream, and the PCM buffer starts at frame pcm_frame0 of the stream (the N/2-sample halo of the
    // shard's first block included); clip_off still spans the WHOLE stream, so samples past its end read as zero
    long long pcm_frame0;
    int blk_base;
};

template <typename T>
struct Handoff {
    T* lines;
    T* bandmax;
    T* smr;
    uint16_t* tokens;
    uint8_t* ovs;
    uint32_t* ms;
};

template <typename T>
struct AnalysisTaps {            // all nullable
    T* lines4;                   // [nblk][4][L]  unscaled
    T* smr4;                     // [nblk][4][32]
    int32_t* npeaks;             // [nblk][4]
};

struct ChainIO {
    const unsigned char* rec;    // [nblk][MRC_REC_BYTES]   (wave-local)
    const unsigned char* pw;     // [nblk][MRC_PW_BYTES]    (wave-local)
    int4* rsv;                   // [nblk] reservoir before group 0, before group 1, after the block, - (wave-local)
    uint32_t* gmask;             // [nblk][32]              (wave-local)
    ChainBlk* cblk;              // [nblk]                  (wave-local)
    int32_t* clip_res;           // [n_clips] reservoir carried from wave to wave
    int64_t* clip_run;           // [n_clips] running byte offset inside the clip's .pac, carried likewise
    int64_t* clip_bytes;         // [n_clips] final .pac size, written when the clip's last block is done
};

struct PackTaps {                // all nullable (parity taps)
    uint8_t* alloc;              // [nblk][2][32]
    uint8_t* sf;                 // [nblk][2][32]
    uint16_t* mant;              // [nblk][2][L]
};

struct CodecParams {
    int L, nb, n_scale_bits, n_mant_size_bits, max_mant_bits, joint;
    int Lmax;                    // nMDCTLines: per-block stride of the hand-off buffers (L of a short block is smaller)
    int no_huff;                 // 1: EncodeNoHuff (codecThem.py:234-260): table 15, no reservoir credit
    int flush_nonjoint;          // 1: the last block of every clip is the non-joint Close() flush block (Q10)
    int spread_seq;              // 1: masker spreading summed pair by pair in the reference's order (psychoac.py:168)
    double budget_joint;         // value of bitBudget just before `+= bitReservoir` (codecThem.py:381-391)
    double budget_single;        // value of bitBudget just before `+= bitReservoir` (codecThem.py:299-306)
    // the same budgets as integers: bitBudget = k + frac + bitReservoir with 0 <= frac < 1 (the blksw bits of the
    // joint path, subtracted after the reservoir is added, are folded into k_joint); frac_* = (frac > 0)
    int k_joint, k_single, frac_joint, frac_single;
    int header_bytes;            // .pac file header size
};

// ---- kernel launchers (defined in the .cu files) -------------------------------------------------------------
template <typename T>
void launch_analysis(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const ClipMap& cm,
                     const int16_t* pcm, const double* xin, int g0, int nblk, Handoff<T> ho, AnalysisTaps<T> taps,
                     unsigned long long* peak_counter);

// cost of every (band, level) under the four books, re-ordered into grant order as prefix sums
template <typename T>
void launch_cost(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                 const ClipMap& cm, int g0, int nblk, Handoff<T> ho, unsigned char* rec, unsigned char* pw);

// serial: one warp per clip c0 .. c0+nclips-1 walks the clip's blocks that lie in [g0, g0+nblk) and records the
// reservoir each block starts with
void launch_chain(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0, int nblk,
                  ChainIO io, const int32_t* reservoir_in, int32_t* reservoir_out,
                  unsigned long long* iter_counter);
// single-stream fast path: tabulate every block's reservoir map for R_in in [r_lo, r_lo+ntab) (parallel), then walk
// the tables (serial); tab is [nblk][2][tabw] ints, tabw >= ntab + 2 and even
void launch_table(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk,
                  ChainIO io, int r_lo, int ntab, int tabw, int* tab);
void launch_chain_table(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0,
                        int nblk, ChainIO io, int r_lo, int ntab, int tabw, const int* tab,
                        const int32_t* reservoir_in, int32_t* reservoir_out, unsigned long long* iter_counter);
// single-stream fast path, second stage: compose the per-block maps over segments of S consecutive blocks and follow
// the values that leave the tabulated range through the next segments (parallel), then one serial step per segment and
// a parallel replay of the stepped-over segments into io.rsv.
// comp is [ceil(nblk/S)][segw] ints, segw >= ntab + 3; segx is [ceil(nblk/S)][segment_aux_width()]; rin [ceil(nblk/S)].
void launch_segments(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk, int S, ChainIO io,
                     int r_lo, int ntab, int tabw, const int* tab, int segw, int* comp, int* segx, int* rin);
void launch_chain_seg(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0, int nblk,
                      int S, ChainIO io, int r_lo, int ntab, int tabw, const int* tab, int segw, const int* comp,
                      const int* segx, int* rin, const int32_t* reservoir_in, int32_t* reservoir_out,
                      unsigned long long* iter_counter, int* bound_rec = nullptr, const int* bound_ref = nullptr);
// bound_rec [nseg + 1]: the walk leaves the reservoir at every segment boundary and, last, its result (a shard walking
// ahead from a guessed reservoir); bound_ref: the walk from the true reservoir stops where it meets that trajectory
// the parallel replay of the stepped-over segments (after launch_chain_seg; nothing the next shard waits for)
void launch_expand(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk, int S, ChainIO io, int r_lo,
                   int ntab, int tabw, const int* tab, const int* rin);
int segment_max_ntab();
int segment_aux_width();
// parallel: one warp per block replays the block from its recorded reservoir: grant masks, table ids, chunk sizes
void launch_finish(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int g0, int nblk,
                   ChainIO io);
// parallel: one warp per clip turns the chunk sizes into byte offsets inside the clip's .pac
void launch_offsets(cudaStream_t st, const CodecParams& cp, const ClipMap& cm, int c0, int nclips, int g0, int nblk,
                    ChainIO io);

template <typename T>
void launch_pack(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                 const ClipMap& cm, int g0, int nblk, Handoff<T> ho, ChainIO io, PackTaps taps,
                 const int64_t* clip_base, uint8_t* out, long long out_cap, const uint8_t* header_template,
                 int* overflow_flag);

// clip_base[c0+i] = *running + sum_{j<i} clip_bytes[c0+j]; *running += sum  (one CTA)
void launch_clip_scan(cudaStream_t st, const int64_t* clip_bytes, int64_t* clip_base, int c0, int n,
                      int64_t* running);

size_t analysis_smem_bytes(int L, int elem, bool xin);

// mrc_transient.cu -- block switching: per nMDCTLines-frame block, flags bit 0 = transient in the first 128 samples,
// bit 1 = transient later in the block (pacfileThem.py:1021-1056)
#define MRC_MAX_SOS 16
struct SosParams {
    int n;                       // second-order sections
    double t0, t1;               // thresholds T[0], T[1] (pacfileThem.py:1154)
    double c[MRC_MAX_SOS][5];    // b0, b1, b2, a1, a2 (a0 = 1)
};
void launch_transient(cudaStream_t st, const SosParams& sp, const int64_t* clip_off, const int32_t* clip_sb0,
                      int n_clips, const int16_t* pcm, int L, int nsb_total, double* peaks, uint8_t* flags);

// mrc_train.cu
void launch_callmax(cudaStream_t st, int L, int ncalls, const uint8_t* alloc, const uint16_t* mant,
                    const uint8_t* line2band, int32_t* cmax);
void launch_hist(cudaStream_t st, int L, int ncalls, int first_call, int thr, const uint8_t* alloc, const uint16_t* mant,
                 const uint8_t* line2band, unsigned long long* hist);

// mrc_peaks.cu
int measure_peaks(cudaStream_t st, double* out);
