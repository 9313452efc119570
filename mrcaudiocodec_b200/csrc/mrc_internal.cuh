// mrc_internal.cuh -- shared declarations for the libmrc.so kernels (sm_100a).
//
// Data layout in HBM (per encode wave of nblk blocks, L = n_mdct_lines, NB = n_bands <= 32):
//   analysis kernel -> hand-off      lines   [nblk][2][L]   real  selected (M|L, S|R per band), scaled by 2^overall
//                                    bandmax [nblk][2][32]  real  max |line| per band of the selected lines
//                                    smr     [nblk][2][32]  real  selected SMRs (tap / token source)
//                                    tokens  [nblk][768]    u16   water-filling grant order (band | level<<8)
//                                    ovs     [nblk][4]      u8    overall scale factors L,R,M,S
//                                    ms      [nblk]         u32   ms_switch bit mask
//   alloc/quantise kernel -> pack    alloc,sf [nblk][2][32] u8 ; table [nblk][2] u8 ; mant [nblk][2][L] u16 ;
//                                    chunk_bytes [nblk][2] u32 ; chunk_off [nblk][2] i64 ; reservoir [nblk] i32
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mrc.h"

#define MRC_TOK_STRIDE 768      // >= 2*25*15 grant tokens per block
#define MRC_MAX_LEVELS 15       // grants per band: 0->2, then +1 up to 16 bits
#define MRC_BSTRIDE 32          // band stride in per-band arrays

template <typename T> struct cpx { T x, y; };

template <typename T>
struct DevTables {
    int L, logL, nb, sample_rate, fstep;         // fstep = sample_rate // (2L)   (Q2)
    const T* kbd;                                // [2L]
    const T* hann;                               // [2L]
    const cpx<T>* tw_pre;                        // [L/2]  exp(-j*pi*(4n+1)/(4L))
    const cpx<T>* tw_post;                       // [L/2]  exp(-j*pi*k/L)
    const cpx<T>* tw_fft;                        // [L/2]  exp(-2*pi*j*k/L)
    const cpx<T>* tw_rfft;                       // [L]    exp(-2*pi*j*k/(2L))
    const T* bark;                               // [L]
    const T* quiet;                              // [L]
    const int* band_lo;                          // [nb]
    const int* band_n;                           // [nb]
    const uint8_t* line2band;                    // [L]
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

struct QuantOut {
    uint8_t* alloc;
    uint8_t* sf;
    uint8_t* table;
    uint16_t* mant;
    uint32_t* chunk_bytes;
    int64_t* chunk_off;
    int32_t* reservoir;
    int64_t* clip_bytes;         // [n_clips]
};

struct CodecParams {
    int L, nb, n_scale_bits, n_mant_size_bits, max_mant_bits, joint;
    int no_huff;                 // 1: EncodeNoHuff (codecThem.py:234-260): table 15, no reservoir credit
    int flush_nonjoint;          // 1: the last block of every clip is the non-joint Close() flush block (Q10)
    double budget_joint;         // value of bitBudget just before `+= bitReservoir` (codecThem.py:381-391)
    double budget_single;        // value of bitBudget just before `+= bitReservoir` (codecThem.py:299-306)
    int header_bytes;            // .pac file header size
};

// ---- kernel launchers (defined in the .cu files) -------------------------------------------------------------
template <typename T>
void launch_analysis(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const ClipMap& cm,
                     const int16_t* pcm, const double* xin, int g0, int nblk, Handoff<T> ho, AnalysisTaps<T> taps,
                     unsigned long long* peak_counter);

template <typename T>
void launch_quant(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                  const ClipMap& cm, int c0, int nclips_wave, int g0, Handoff<T> ho, QuantOut qo,
                  const int32_t* reservoir_in, int32_t* reservoir_out);

void launch_pack(cudaStream_t st, const CodecParams& cp, const HuffDev* huff, const int* band_lo, const int* band_n,
                 const uint8_t* line2band, const ClipMap& cm, int g0, int nblk, QuantOut qo, const uint8_t* ovs,
                 const uint32_t* ms, const int64_t* clip_base, uint8_t* out, long long out_cap,
                 const uint8_t* header_template, int* overflow_flag);

// clip_base[c0+i] = *running + sum_{j<i} clip_bytes[c0+j]; *running += sum  (one CTA)
void launch_clip_scan(cudaStream_t st, const int64_t* clip_bytes, int64_t* clip_base, int c0, int n,
                      int64_t* running);

size_t analysis_smem_bytes(int L, int elem);

// mrc_peaks.cu
int measure_peaks(cudaStream_t st, double* out);
