// mrc_decode.cuh -- declarations of the decode kernels (defined in mrc_decode.cu).
#pragma once
#include "mrc_internal.cuh"

#define MRC_HUFF_PEEK 9          // longest code in the four trained books (Appendix E)

struct HuffDecDev {
    uint16_t lut[MRC_N_HUFF_TABLES][1 << MRC_HUFF_PEEK];   // value | len<<8, indexed by the next 9 bits
};

struct DecodeMap {
    const int64_t* chunk_pos;    // [2*npairs] byte offset of the chunk payload (after its <L prefix) in `pac`
    const uint32_t* chunk_len;   // [2*npairs]
    const int32_t* clip_pair0;   // [n_clips+1] first global pair of every clip
    int n_clips;
    // block switching (null / unused for streams of long blocks only)
    const uint8_t* pair_geom;    // [npairs] MRC_GEO_* from the chunk headers
    const int64_t* pair_pos;     // [npairs] first output frame (relative to the clip) of tail(pair) + head(next pair)
    const int32_t* list;         // wave-local pairs of the geometry being launched (null: all)
};

// what the parse kernel (one warp per chunk) hands to the synthesis kernel (one CTA per pair), wave-local pair index
struct ParseOut {
    uint16_t* mant;              // [npairs][2][Lmax]  mantissa codes
    uint8_t* alloc;              // [npairs][2][32]    bits per mantissa, per band (0 = band not coded)
    uint8_t* sf;                 // [npairs][2][32]    scale factor per band
    uint8_t* ovs;                // [npairs][4]        overall scale factors
    uint32_t* ms;                // [npairs]           ms_switch mask
    uint8_t* flags;              // [npairs][2]        bit 0: malformed chunk, bit 1: joint pair
};
size_t parse_out_bytes(int npairs, int Lmax);
ParseOut parse_out_carve(void* base, int npairs, int Lmax);

// parse (one warp per chunk), then dequantise + M/S + IMDCT + window (one CTA per pair) for pairs [p0, p0+npairs):
// writes y [npairs][2][2L] (head, tail)
template <typename T>
void launch_decode(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                   const HuffDecDev* hdec, const DecodeMap& dm, const uint8_t* pac, int p0, int npairs, T* y,
                   int* error_flag, const ParseOut& po);

// same synthesis from explicit integers (per-block seam): ints are [npairs][2][..] like mrc_decode_block
template <typename T>
void launch_decode_ints(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, int joint,
                        const int32_t* sf, const int32_t* alloc, const int32_t* mant, const int32_t* ovs,
                        const int32_t* ms, int npairs, T* y);

// overlap-add + PCM conversion: frame block j of a clip = tail(y_j) + head(y_{j+1}) (last: tail only)
template <typename T>
void launch_ola(cudaStream_t st, const CodecParams& cp, const DecodeMap& dm, int p0, int npairs, const T* y,
                const int64_t* clip_frame_off, int16_t* pcm);
