/*
 * mrc.h -- C ABI of libmrc.so: the B200-native encode/decode hot path of the MRC ("Music 422 PAC")
 * perceptual audio codec, a drop-in behind the reference's own seam.
 *
 * Reference interfaces replaced (file:line in laser55/mrcAudioCodec):
 *   per-block seam     PACFile.Encode / JointEncode / Decode / JointDecode   pacfileThem.py:987-1019
 *                      -> codecThem.Encode :205-231, JointEncode :262-278, Decode :30-63, JointDecode :65-134
 *   whole-file loop    audiofile.py:24-38 driving PACFile.WriteFileHeader :586-619, WriteDataBlock :622-790,
 *                      JointWriteDataBlock :793-972, Close :973-984, ReadFileHeader :130-158,
 *                      ReadDataBlock :161-319, JointReadDataBlock :321-585 and the PCM sample conversion of
 *                      PCMFile.ReadDataBlock pcmfile.py:68-102 / WriteDataBlock :156-185
 *   block switching    the `__main__` encode loop pacfileThem.py:1142-1215 with TransientDetector :1021-1056
 *                      (MRC_FLAG_BLOCK_SWITCHING, mrc_set_switch_tables, mrc_detect_transients, *_block_ab)
 *
 * Conventions: every entry point returns 0 on success or a negative MRC_E_* code; mrc_last_error(ctx) gives a
 * human-readable message.  No exceptions cross the boundary.  There is no CPU fallback: without a CUDA device
 * mrc_create fails with MRC_E_CUDA.  A context is bound to one device and one CUDA stream; calls on one context
 * must be serialised by the caller.  The library never keeps host pointers beyond the call that received them.
 * All multi-byte values in .pac data are little-endian, bits are MSB-first (bitpack.py:36-101).
 */
#ifndef MRC_H_
#define MRC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRC_VERSION 100

#define MRC_OK 0
#define MRC_E_INVALID (-1)   /* bad argument / unsupported configuration            */
#define MRC_E_CUDA (-2)      /* CUDA error or no device                             */
#define MRC_E_NOSPACE (-3)   /* output buffer too small (needed size is reported)   */
#define MRC_E_FORMAT (-4)    /* malformed .pac input                                */
#define MRC_E_STATE (-5)     /* tables not set, etc.                                */

#define MRC_MAX_BANDS 32
#define MRC_HUFF_LUT 65      /* mantissa values 0..64 can have a code; anything else is an escape (Appendix E) */
#define MRC_N_HUFF_TABLES 4
#define MRC_NO_TABLE 15

#define MRC_PRECISION_FP64 0 /* code-exact mode                                      */
#define MRC_PRECISION_FP32 1 /* fast mode (MDCT/SMR within 1e-5 relative)            */

/* Masker spreading (psychoac.py:68-78, :168) summed pair by pair in the reference's order: one 10**x per
 * (masker, line) pair.  Default (flag clear) is the factorised evaluation of the same sum. */
#define MRC_FLAG_SPREAD_SEQUENTIAL 1
/* Never tabulate the per-block reservoir maps (the single-stream fast path of the serial walk); results are the
 * same either way, the flag exists for cross-checking and timing. */
#define MRC_FLAG_NO_CHAIN_TABLES 2
/* Block switching: mrc_encode_batch follows the reference's `__main__` loop (pacfileThem.py
 * :1142-1215) instead of the plain per-block loop: a transient detector with one block of look-ahead decides, per
 * n_mdct_lines-frame block, between one long block and eight 128-sample short blocks.  Needs
 * mrc_set_switch_tables, joint = 1 and n_mdct_lines = 1024.  Decoding never needs the flag: the block sizes are in
 * every chunk header. */
#define MRC_FLAG_BLOCK_SWITCHING 4

typedef struct mrc_ctx mrc_ctx;

/* Mirrors the attribute bag the reference fills at pacfileThem.py:1105-1121. */
typedef struct mrc_config {
    int32_t device;                 /* CUDA device ordinal                                            */
    int32_t sample_rate;            /* codingParams.sampleRate (int, pcmfile.py:46)                   */
    int32_t n_mdct_lines;           /* codingParams.nMDCTLines: 128 .. 2048, a power of two           */
    int32_t n_scale_bits;           /* codingParams.nScaleBits  (4)                                   */
    int32_t n_mant_size_bits;       /* codingParams.nMantSizeBits (4)                                 */
    int32_t joint;                  /* 1: JointWriteDataBlock flow (M/S), 0: WriteDataBlock flow      */
    int32_t precision;              /* MRC_PRECISION_*                                                */
    int32_t flags;                  /* MRC_FLAG_*                                                     */
    double target_bits_per_sample;  /* codingParams.targetBitsPerSample                               */
    int64_t reserved1[4];
} mrc_config;

/* Data-independent tables, computed by the host exactly as the reference computes them (numpy), uploaded once.
 * kbd_window = TransitionWindow(ones, L, L) window.py:104-121; hann_window window.py:36-42;
 * bark = Bark(MDCTFreq) psychoac.py:29,143; quiet_intensity = Intensity(Thresh(MDCTFreq)) psychoac.py:155;
 * band_nlines = AssignMDCTLinesFromFreqLimits psychoac.py:86-105; huff_* from training_data/*_table.pkl in
 * alphabetical table order (percussive, silence, speech, tonal). */
typedef struct mrc_tables {
    int32_t n_bands;
    int32_t n_huff_tables;                 /* must be MRC_N_HUFF_TABLES                                  */
    const int32_t* band_nlines;            /* [n_bands]                                                  */
    const double* kbd_window;              /* [2*n_mdct_lines]                                           */
    const double* hann_window;             /* [2*n_mdct_lines]                                           */
    const double* bark;                    /* [n_mdct_lines]                                             */
    const double* quiet_intensity;         /* [n_mdct_lines]                                             */
    const int32_t* huff_escape;            /* [n_huff_tables] escape value of each table                 */
    const uint8_t* huff_len;               /* [n_huff_tables][MRC_HUFF_LUT] code length, 0 = not a key   */
    const uint16_t* huff_code;             /* [n_huff_tables][MRC_HUFF_LUT] code bits (right aligned)    */
} mrc_tables;

/* One block geometry of block switching: window halves a (the previous block's size) and b (this block's), each
 * n_mdct_lines or 128.  window = TransitionWindow(ones, a, b) window.py:104-121; hann_window window.py:36-42 at
 * a+b points; bark / quiet_intensity at the (a+b)/2 line frequencies; band_nlines = AssignMDCTLinesFromFreqLimits
 * ((a+b)/2, sampleRate, the 9-band short table) pacfileThem.py:213-219. */
typedef struct mrc_block_tables {
    int32_t a, b;
    int32_t n_bands;
    const int32_t* band_nlines;            /* [n_bands], sums to (a+b)/2                                 */
    const double* window;                  /* [a+b]                                                      */
    const double* hann_window;             /* [a+b]                                                      */
    const double* bark;                    /* [(a+b)/2]                                                  */
    const double* quiet_intensity;         /* [(a+b)/2]                                                  */
} mrc_block_tables;

int32_t mrc_version(void);
const char* mrc_last_error(const mrc_ctx* ctx);   /* ctx may be NULL: message of the last failed mrc_create */

int32_t mrc_create(const mrc_config* cfg, mrc_ctx** out);
int32_t mrc_destroy(mrc_ctx* ctx);
int32_t mrc_set_tables(mrc_ctx* ctx, const mrc_tables* t);

/* Block switching tables, after mrc_set_tables: t3 = the geometries (L,128), (128,L), (128,128) in this order;
 * sos = the transient detector's high-pass as n_sections rows of b0 b1 b2 a0 a1 a2 with a0 = 1 (the reference
 * designs it with scipy.signal.cheby2(20, 40, 9000./sampleRate, 'high') + tf2sos, pacfileThem.py:1146-1147);
 * t0, t1 = the thresholds T (pacfileThem.py:1154: 0.1, 0.075). */
int32_t mrc_set_switch_tables(mrc_ctx* ctx, const mrc_block_tables* t3, const double* sos, int32_t n_sections,
                              double t0, double t1);

/* Pinned host memory for PCM / bitstream buffers (optional; pageable memory works but copies slower). */
int32_t mrc_host_alloc(void** p, int64_t bytes);
int32_t mrc_host_free(void* p);

/* ---- whole-file batch encode: replaces the audiofile.py:24-38 loop over PCMFile.ReadDataBlock +
 * PACFile.(Joint)WriteDataBlock + Close for n_clips independent files ------------------------------------
 * pcm                : interleaved 16-bit stereo frames of all clips back to back (host memory)
 * clip_frame_offsets : [n_clips+1] frame offsets into pcm
 * out / out_cap      : receives the n_clips .pac files back to back (header + block chunks + flush block)
 * clip_byte_offsets  : [n_clips+1] byte offsets into out (written even on MRC_E_NOSPACE, so the caller can
 *                      size the buffer: clip_byte_offsets[n_clips] is the total)                           */
int32_t mrc_encode_batch(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                         uint8_t* out, int64_t out_cap, int64_t* clip_byte_offsets);

/* Same, with pcm and out in DEVICE memory of ctx's device (clip offset arrays stay on the host).  Used to time
 * the kernels with inputs resident in HBM.  Runs on the context's stream and synchronises before returning. */
int32_t mrc_encode_batch_device(mrc_ctx* ctx, const int16_t* d_pcm, const int64_t* clip_frame_offsets,
                                int32_t n_clips, uint8_t* d_out, int64_t out_cap, int64_t* clip_byte_offsets);

/* ---- whole-file batch decode: replaces the loop over PACFile.(Joint)ReadDataBlock + PCMFile.WriteDataBlock --
 * pac / clip_byte_offsets : n_clips .pac files back to back (host memory)
 * pcm_out / pcm_cap_frames: receives interleaved 16-bit frames; per clip (#block pairs) * n_mdct_lines frames
 *                           (pairs 1..B overlap-added + the saved tail, first block dropped: pacfileThem.py
 *                           :1175-1177, :178-185)
 * clip_frame_offsets      : [n_clips+1] frame offsets into pcm_out (output)
 * The last block pair of every file is read as a non-joint pair (Close() writes it so, Q10); the others follow
 * ctx's `joint` flag.                                                                                       */
int32_t mrc_decode_batch(mrc_ctx* ctx, const uint8_t* pac, const int64_t* clip_byte_offsets, int32_t n_clips,
                         int16_t* pcm_out, int64_t pcm_cap_frames, int64_t* clip_frame_offsets);
int32_t mrc_decode_batch_device(mrc_ctx* ctx, const uint8_t* d_pac, const uint8_t* h_pac,
                                const int64_t* clip_byte_offsets, int32_t n_clips, int16_t* d_pcm_out,
                                int64_t pcm_cap_frames, int64_t* clip_frame_offsets);

/* ---- one stream sharded by block range across GPUs (one context per GPU) ----------------------------------------
 * The blocks of a stream are independent except for two things a block inherits from the one before it: the previous
 * n_mdct_lines frames (pacfileThem.py:799-802 -- a halo the shard reads from the PCM itself) and
 * codingParams.bitReservoir (codecThem.py:391, :503, :274 -- one int).  A shard encodes blocks
 * [first_block, first_block + n_blocks) of the stream's ceil(total_frames / n_mdct_lines) blocks: transform,
 * psychoacoustics, bit prices and the composed reservoir maps first (all independent of the reservoir), then it calls
 * exchange(user, 0, &r) to RECEIVE the reservoir the previous shard ended with (the first shard is given 0), runs the
 * serial pass, calls exchange(user, 1, &r) to HAND ON its own final reservoir, and only then packs its chunks.  The
 * shards' outputs concatenated in order are byte-identical to mrc_encode_batch of the whole stream; the first shard
 * writes the file header (is_first), the last one the Close() flush block (is_last).
 * pcm : host frames [pcm_frame0, pcm_frame0 + pcm_frames) of the stream; must cover
 *       [max(first_block - 1, 0) * L, min((first_block + n_blocks) * L, total_frames)).
 * The callbacks return 0 on success.  Long blocks only (no MRC_FLAG_BLOCK_SWITCHING). */
typedef int32_t (*mrc_reservoir_exchange)(void* user, int32_t have_result, int32_t* reservoir);
int32_t mrc_encode_shard(mrc_ctx* ctx, const int16_t* pcm, int64_t pcm_frame0, int64_t pcm_frames, int64_t total_frames,
                         int64_t first_block, int64_t n_blocks, int32_t is_first, int32_t is_last, uint8_t* out,
                         int64_t out_cap, int64_t* out_bytes, mrc_reservoir_exchange exchange, void* user);
/* Same with pcm and out in DEVICE memory of ctx's device (kernel-only timing; no retry when out is too small). */
int32_t mrc_encode_shard_device(mrc_ctx* ctx, const int16_t* d_pcm, int64_t pcm_frame0, int64_t pcm_frames,
                                int64_t total_frames, int64_t first_block, int64_t n_blocks, int32_t is_first,
                                int32_t is_last, uint8_t* d_out, int64_t out_cap, int64_t* out_bytes,
                                mrc_reservoir_exchange exchange, void* user);

/* ---- per-block seam (compat layer; explicit bit reservoir in/out) -----------------------------------------
 * mrc_encode_block = codecThem.Encode (joint=0) / JointEncode (joint=1) on one block; joint|2 skips the Huffman
 * stage (EncodeNoHuff, codecThem.py:234-260: table 15, no reservoir credit); joint|4 (with joint bit 0 clear) is
 * codecThem.Encode with codingParams.nChannels = 1 (:216 loops over the channels): data is ONE channel [a+b], every
 * output holds one channel's worth ([n_bands], [n_mdct_lines], [1]), ms_switch is not written, and the reservoir
 * returned is the one after that channel.
 * data             : [2][2*n_mdct_lines] float64 signed fractions (prior block, current block) per channel
 * reservoir        : in/out codingParams.bitReservoir
 * scale_factor,bit_alloc : [2][n_bands]; mantissa: [2][n_mdct_lines] aligned to MDCT lines (0 where the band has
 * no bits); overall_scale: [4] (joint: L,R,M,S; non-joint: ch0,ch1,-,-); ms_switch: [n_bands]; huff_table: [2];
 * chunk_bytes: [2] size of each channel chunk as (Joint)WriteDataBlock would write it.                      */
int32_t mrc_encode_block(mrc_ctx* ctx, const double* data, int32_t joint, int32_t* reservoir,
                         int32_t* scale_factor, int32_t* bit_alloc, int32_t* mantissa, int32_t* overall_scale,
                         int32_t* ms_switch, int32_t* huff_table, int32_t* chunk_bytes);
/* The same seam for any block geometry of block switching (codingParams.a, codingParams.b): data is
 * [2][a+b], scale_factor / bit_alloc / ms_switch have that geometry's band count (9 unless a = b = n_mdct_lines),
 * mantissa is [2][(a+b)/2].  Needs mrc_set_switch_tables unless a = b = n_mdct_lines. */
int32_t mrc_encode_block_ab(mrc_ctx* ctx, const double* data, int32_t a, int32_t b, int32_t joint, int32_t* reservoir,
                            int32_t* scale_factor, int32_t* bit_alloc, int32_t* mantissa, int32_t* overall_scale,
                            int32_t* ms_switch, int32_t* huff_table, int32_t* chunk_bytes);
int32_t mrc_decode_block_ab(mrc_ctx* ctx, int32_t a, int32_t b, int32_t joint, const int32_t* scale_factor,
                            const int32_t* bit_alloc, const int32_t* mantissa, const int32_t* overall_scale,
                            const int32_t* ms_switch, double* data_out);

/* TransientDetector (pacfileThem.py:1021-1056) over whole clips plus the look-ahead decision (:1192), as
 * mrc_encode_batch applies them under MRC_FLAG_BLOCK_SWITCHING.  All outputs may be NULL.
 * flags [sum over clips of ceil(frames/L)]: bit 0 = a transient in the block's first 128 samples, bit 1 = one later
 * in the block.  block_ab [block_cap][2]: (a, b) of every block the encoder writes, flush blocks included;
 * clip_block_offsets [n_clips+1]: first written block of every clip (written even on MRC_E_NOSPACE). */
int32_t mrc_detect_transients(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                              uint8_t* flags, int32_t* block_ab, int32_t block_cap, int32_t* clip_block_offsets);

/* mrc_decode_block = codecThem.Decode x2 (joint=0) / JointDecode (joint=1): returns the windowed IMDCT output
 * [2][2*n_mdct_lines] (before overlap-add), like the reference functions. */
int32_t mrc_decode_block(mrc_ctx* ctx, int32_t joint, const int32_t* scale_factor, const int32_t* bit_alloc,
                         const int32_t* mantissa, const int32_t* overall_scale, const int32_t* ms_switch,
                         double* data_out);

/* ---- stage taps for parity tests (run the encode front end on whole clips and return intermediates) -------
 * All output pointers may be NULL.  n_blocks_total = sum over clips of (ceil(frames/L) + 1).
 * mdct_lines [n_blocks_total][4][L]  UNSCALED lines of L,R,M,S (M,S zero for non-joint blocks)
 * overall_scale [n_blocks_total][4], ms_switch [n_blocks_total][n_bands], smr [n_blocks_total][4][n_bands],
 * n_peaks [n_blocks_total][4] tonal maskers found per spectrum.                                              */
int32_t mrc_stage_analysis(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                           double* mdct_lines, int32_t* overall_scale, int32_t* ms_switch, double* smr,
                           int32_t* n_peaks);
/* bit_alloc,scale_factor [n_blocks_total][2][n_bands]; mantissa [n_blocks_total][2][L]; huff_table
 * [n_blocks_total][2]; reservoir [n_blocks_total] (codingParams.bitReservoir after the block);
 * chunk_bytes [n_blocks_total][2].                                                                           */
int32_t mrc_stage_alloc_quant(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets,
                              int32_t n_clips, int32_t* bit_alloc, int32_t* scale_factor, int32_t* mantissa,
                              int32_t* huff_table, int32_t* reservoir, int32_t* chunk_bytes);

/* ---- Huffman table training front end (huffman_training_script.py:33-66) ----------------------
 * Encodes the clips block by block like the script's loop -- independent channels (the context must have joint = 0),
 * EncodeNoHuff (codecThem.py:234-260), reservoir starting at 0 per file, no flush block -- and runs every channel's
 * compacted mantissa vector through calculateFrequencies (huffman.py:56-71), including that function's reset quirk:
 * the first never-seen (hence record-high) value of a call zeroes the counts of all smaller values.
 * prior_max : largest mantissa value counted by earlier calls (-1 for the first call)
 * hist      : [65536] counts accumulated in THIS call since its last reset (or since its start if it had none)
 * max_out   : largest value seen so far (prior_max included);  reset_out : 1 if a reset happened in this call (the
 *             caller then replaces its running table by hist instead of adding to it)
 * At most 16384 blocks per call.                                                                                */
int32_t mrc_mantissa_histogram(mrc_ctx* ctx, const int16_t* pcm, const int64_t* clip_frame_offsets, int32_t n_clips,
                               int32_t prior_max, int64_t* hist, int32_t* max_out, int32_t* reset_out);

/* ---- instrumentation ------------------------------------------------------------------------------------- */
/* Device time (ms, CUDA events on the stream each kernel is launched on, summed over waves) of the stages of the
 * last encode/decode call: [0] analysis kernels, [1] chain (serial reservoir walk) kernels (after a decode: host milliseconds of the header checks and the <L nBytes> chain walk), [2] clip-offset scan +
 * quantise/pack kernels, [3] decode kernels (after an encode with block switching: the transient detector kernels),
 * [4] H2D, [5] D2H (encode: sums of the wave-by-wave copies on the copy streams, which overlap the kernels; decode: 0, not timed apart), [6] whole call on the main stream, [7] cost kernels.
 * Analysis+cost of wave w+1 overlap chain+pack of wave w, so [0]+[7]+[1]+[2] can exceed [6].
 * counters [0] kernel launches, [1] sum over spectra of tonal maskers, [2] blocks, [4] waves, and the work the
 * factorised spreading executed: [5] general (masker, line) pairs priced with one 10**x each, [6] plateau
 * additions, [7] loud maskers (g > 0). */
int32_t mrc_last_timing(const mrc_ctx* ctx, double* ms8, int64_t* counters8);
/* Micro-benchmarks of this GPU's pipes, for the roofline denominators MEASURED_PEAKS.json does not carry:
 * out4[0] FP64 FMA TFLOP/s, [1] FP32 FMA TFLOP/s, [2] MUFU.EX2 Gop/s, [3] device copy GB/s (read+write). */
int32_t mrc_measure_peaks(mrc_ctx* ctx, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* MRC_H_ */
