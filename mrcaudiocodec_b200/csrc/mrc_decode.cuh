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

// parse + dequantise + M/S + IMDCT + window for pairs [p0, p0+npairs): writes y [npairs][2][2L] (head, tail)
template <typename T>
void launch_decode(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, const HuffDev* huff,
                   const HuffDecDev* hdec, const DecodeMap& dm, const uint8_t* pac, int p0, int npairs, T* y,
                   int* error_flag);

// same synthesis from explicit integers (per-block seam): ints are [npairs][2][..] like mrc_decode_block
template <typename T>
void launch_decode_ints(cudaStream_t st, const DevTables<T>& tb, const CodecParams& cp, int joint,
                        const int32_t* sf, const int32_t* alloc, const int32_t* mant, const int32_t* ovs,
                        const int32_t* ms, int npairs, T* y);

// overlap-add + PCM conversion: frame block j of a clip = tail(y_j) + head(y_{j+1}) (last: tail only)
template <typename T>
void launch_ola(cudaStream_t st, const CodecParams& cp, const DecodeMap& dm, int p0, int npairs, const T* y,
                const int64_t* clip_frame_off, int16_t* pcm);
