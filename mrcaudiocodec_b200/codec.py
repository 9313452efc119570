"""Python host side of the B200 codec: a thin object around one libmrc context (one GPU, one stream).

Batch API (what is timed): encode_batch / decode_batch replace the reference's whole-file loops
(audiofile.py:24-38 over pcmfile.py + pacfileThem.py).  The per-block seam with the reference's own function
names lives in codec_gpu.py."""
import ctypes as C

import numpy as np

from . import _lib
from .tables import Tables, BlockTables, transient_sos, TRANSIENT_THRESHOLDS, N_SHORT


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Codec(object):
    def __init__(self, sample_rate=48000, n_mdct_lines=1024, n_scale_bits=4, n_mant_size_bits=4,
                 target_bits_per_sample=128000. / 48000., joint=True, precision="fp64", device=0,
                 band_limits=None, spreading="factorised", chain_tables=True, block_switching=False,
                 switch_tables=None, transient_sos_sections=None, window="kbd"):
        """block_switching=True: encode_batch follows the reference's `__main__` loop (transient detector, one
        block of look-ahead, eight 128-sample short blocks around transients; pacfileThem.py:1142-1215).
        switch_tables=True (implied by block_switching) only loads the extra block geometries, which is what
        decoding a switched stream and the per-block seam with a != b need.  transient_sos_sections overrides the
        detector's filter (default: designed with scipy exactly like the reference).
        window: "kbd" (the reference codec's KBD alpha=4 window) or "sine" (window.py:10-25, SineWindow) for both the
        MDCT analysis and the IMDCT synthesis; a stream must be decoded with the window it was encoded with (the .pac
        header does not record it)."""
        self.lib = _lib.load()
        self.L = int(n_mdct_lines)
        self.sample_rate = int(sample_rate)
        self.joint = bool(joint)
        self.precision = precision
        cfg = _lib.MrcConfig()
        cfg.device = int(device)
        cfg.sample_rate = self.sample_rate
        cfg.n_mdct_lines = self.L
        cfg.n_scale_bits = int(n_scale_bits)
        cfg.n_mant_size_bits = int(n_mant_size_bits)
        cfg.joint = 1 if joint else 0
        cfg.precision = {"fp64": _lib.PRECISION_FP64, "fp32": _lib.PRECISION_FP32}[precision]
        cfg.flags = {"factorised": 0, "sequential": _lib.FLAG_SPREAD_SEQUENTIAL}[spreading]
        if not chain_tables:
            cfg.flags |= _lib.FLAG_NO_CHAIN_TABLES
        if block_switching:
            cfg.flags |= _lib.FLAG_BLOCK_SWITCHING
        self.block_switching = bool(block_switching)
        cfg.target_bits_per_sample = float(target_bits_per_sample)
        self._ctx = C.c_void_p()
        rc = self.lib.mrc_create(C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            raise _lib.MrcError(rc, self.lib.mrc_last_error(None).decode())
        if window != "kbd" and (block_switching or switch_tables):
            raise ValueError("block switching uses the reference's KBD transition windows (window.py:104-121)")
        self.tables = Tables(self.L, self.sample_rate, band_limits, window)
        t = _lib.MrcTables()
        T = self.tables
        t.n_bands = T.n_bands
        t.n_huff_tables = len(T.huff_escape)
        t.band_nlines = T.band_nlines.ctypes.data_as(_lib.c_i32p)
        t.kbd_window = T.kbd.ctypes.data_as(_lib.c_f64p)
        t.hann_window = T.hann.ctypes.data_as(_lib.c_f64p)
        t.bark = T.bark.ctypes.data_as(_lib.c_f64p)
        t.quiet_intensity = T.quiet.ctypes.data_as(_lib.c_f64p)
        t.huff_escape = T.huff_escape.ctypes.data_as(_lib.c_i32p)
        t.huff_len = T.huff_len.ctypes.data_as(_lib.c_u8p)
        t.huff_code = T.huff_code.ctypes.data_as(_lib.c_u16p)
        self._check(self.lib.mrc_set_tables(self._ctx, C.byref(t)))
        self.n_bands = T.n_bands
        self.block_tables = {(self.L, self.L): None}
        if block_switching or switch_tables:
            self._set_switch_tables(transient_sos_sections)

    def _set_switch_tables(self, sos=None):
        L = self.L
        geos = [(L, N_SHORT), (N_SHORT, L), (N_SHORT, N_SHORT)]
        arr = (_lib.MrcBlockTables * 3)()
        for i, (a, b) in enumerate(geos):
            bt = BlockTables(a, b, L, self.sample_rate)
            self.block_tables[(a, b)] = bt
            arr[i].a, arr[i].b, arr[i].n_bands = a, b, bt.n_bands
            arr[i].band_nlines = bt.band_nlines.ctypes.data_as(_lib.c_i32p)
            arr[i].window = bt.window.ctypes.data_as(_lib.c_f64p)
            arr[i].hann_window = bt.hann.ctypes.data_as(_lib.c_f64p)
            arr[i].bark = bt.bark.ctypes.data_as(_lib.c_f64p)
            arr[i].quiet_intensity = bt.quiet.ctypes.data_as(_lib.c_f64p)
        self.sos = np.ascontiguousarray(transient_sos(self.sample_rate) if sos is None else sos, dtype=np.float64)
        self._check(self.lib.mrc_set_switch_tables(self._ctx, arr, _ptr(self.sos), int(self.sos.shape[0]),
                                                   float(TRANSIENT_THRESHOLDS[0]), float(TRANSIENT_THRESHOLDS[1])))

    def geometry(self, a, b):
        """(n_lines, n_bands, band_nlines, band_lower) of the block geometry (a, b)."""
        bt = self.block_tables.get((int(a), int(b)))
        if bt is None:
            if (int(a), int(b)) != (self.L, self.L):
                raise ValueError("block geometry (%d, %d) not loaded: create the Codec with switch_tables=True" % (a, b))
            return self.L, self.n_bands, self.tables.band_nlines, self.tables.band_lower
        return bt.n_lines, bt.n_bands, bt.band_nlines, bt.band_lower

    def detect_transients(self, clips):
        """TransientDetector + look-ahead decision over whole clips: (flags per n_mdct_lines-frame block, list of
        per-clip [(a, b), ...] of the blocks the encoder writes, flush block included)."""
        pcm, off = self._concat(clips)
        nsb = int(np.sum((np.diff(off) + self.L - 1) // self.L))
        flags = np.zeros(max(nsb, 1), np.uint8)
        cap = 8 * nsb + len(clips) + 1
        ab = np.zeros((cap, 2), np.int32)
        boff = np.zeros(len(clips) + 1, np.int32)
        self._check(self.lib.mrc_detect_transients(self._ctx, _ptr(pcm), _ptr(off), len(clips), _ptr(flags), _ptr(ab),
                                                   cap, _ptr(boff)))
        return flags[:nsb], [[tuple(int(v) for v in r) for r in ab[boff[i]:boff[i + 1]]] for i in range(len(clips))]

    # ------------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise _lib.MrcError(rc, self.lib.mrc_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.mrc_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def n_blocks(self, frames):
        """block pairs per clip: ceil(frames/L) data blocks + the Close() flush block."""
        return (int(frames) + self.L - 1) // self.L + 1

    @staticmethod
    def _concat(clips):
        # frames x 2, nothing else: a mono array reshaped to (-1, 2) would be encoded as if its even and odd samples were
        # the two channels.  One channel goes through the per-block seam (codec_gpu, nChannels = 1)
        for i, c in enumerate(clips):
            sh = np.shape(c)
            if np.size(c) and (len(sh) != 2 or sh[1] != 2):
                raise ValueError("clip %d has shape %r: the whole-file entry points take int16 [frames, 2] (stereo)"
                                 % (i, tuple(sh)))
        clips = [np.ascontiguousarray(c, dtype=np.int16).reshape(-1, 2) for c in clips]
        off = np.zeros(len(clips) + 1, dtype=np.int64)
        off[1:] = np.cumsum([c.shape[0] for c in clips])
        pcm = np.concatenate(clips, axis=0) if clips else np.zeros((0, 2), np.int16)
        return np.ascontiguousarray(pcm), off

    def nominal_capacity(self, frame_offsets):
        fr = np.diff(frame_offsets)
        nblk = (fr + self.L - 1) // self.L + 1
        return int(np.sum(256 + nblk * (2 * 1.5 * self.L * 16 // 8 // 4 + 256 + (256 if self.block_switching else 0))))

    # ---- batch encode ----------------------------------------------------------------------------
    def encode_batch(self, pcm, frame_offsets, out=None):
        """pcm: int16 [frames,2] of all clips back to back (host), frame_offsets int64 [n+1].
        Returns (out uint8 array, byte_offsets int64 [n+1])."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        frame_offsets = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        n = len(frame_offsets) - 1
        boff = np.zeros(n + 1, dtype=np.int64)
        if out is None:
            out = np.empty(self.nominal_capacity(frame_offsets), dtype=np.uint8)
        rc = self.lib.mrc_encode_batch(self._ctx, _ptr(pcm), _ptr(frame_offsets), n, _ptr(out), out.size, _ptr(boff))
        if rc == _lib.MRC_E_NOSPACE and boff[n] > out.size:
            out = np.empty(int(boff[n]), dtype=np.uint8)
            rc = self.lib.mrc_encode_batch(self._ctx, _ptr(pcm), _ptr(frame_offsets), n, _ptr(out), out.size,
                                           _ptr(boff))
        self._check(rc)
        return out, boff

    def encode_clips(self, clips):
        """list of int16 [frames,2] arrays -> list of .pac byte strings."""
        pcm, off = self._concat(clips)
        out, boff = self.encode_batch(pcm, off)
        return [out[boff[i]:boff[i + 1]].tobytes() for i in range(len(clips))]

    def encode_batch_device(self, d_pcm_ptr, frame_offsets, d_out_ptr, out_cap):
        """device-resident variant: raw device pointers (e.g. torch tensor .data_ptr())."""
        frame_offsets = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        n = len(frame_offsets) - 1
        boff = np.zeros(n + 1, dtype=np.int64)
        self._check(self.lib.mrc_encode_batch_device(self._ctx, C.c_void_p(d_pcm_ptr), _ptr(frame_offsets), n,
                                                     C.c_void_p(d_out_ptr), int(out_cap), _ptr(boff)))
        return boff

    # ---- one stream sharded by block range (SURVEY.md 8e "within one stream") --------------------
    def shard_pcm_range(self, total_frames, first_block, n_blocks):
        """frames [lo, hi) of the stream a shard needs: its blocks plus the n_mdct_lines-frame halo before them."""
        lo = max(int(first_block) - 1, 0) * self.L
        hi = min((int(first_block) + int(n_blocks)) * self.L, int(total_frames))
        return lo, max(hi, lo)

    def encode_shard(self, pcm, pcm_frame0, total_frames, first_block, n_blocks, is_first, is_last, recv_reservoir,
                     send_reservoir, out=None, device_ptrs=None):
        """Blocks [first_block, first_block + n_blocks) of a stream of total_frames frames.  pcm: int16 [frames, 2]
        starting at stream frame pcm_frame0 (see shard_pcm_range).  recv_reservoir() -> int is called once everything
        that does not depend on the reservoir is done and must return the reservoir the previous shard ended with (0
        for the first shard); send_reservoir(r) is called with this shard's final reservoir right after the serial
        pass, before the chunks are packed.  Returns the shard's bytes: concatenated in order the shards are the
        .pac file mrc_encode_batch writes for the whole stream."""
        if device_ptrs is None:
            pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1, 2)
        err = []

        def _cb(user, have_result, r):
            try:
                if have_result:
                    send_reservoir(int(r[0]))
                else:
                    r[0] = int(recv_reservoir())
                return 0
            except Exception as e:         # no exception may cross the C boundary
                err.append(e)
                return 1
        cb = _lib.RESERVOIR_EXCHANGE(_cb)
        nbytes = C.c_int64(0)
        if device_ptrs is not None:
            # (d_pcm pointer, frames it holds, d_out pointer, capacity): kernel-only timing, returns the byte count
            d_pcm, n_fr, d_out, cap = device_ptrs
            rc = self.lib.mrc_encode_shard_device(self._ctx, C.c_void_p(d_pcm), int(pcm_frame0), int(n_fr),
                                                  int(total_frames), int(first_block), int(n_blocks),
                                                  1 if is_first else 0, 1 if is_last else 0, C.c_void_p(d_out), int(cap),
                                                  C.byref(nbytes), cb, None)
            if err:
                raise err[0]
            self._check(rc)
            return int(nbytes.value)
        if out is None:
            out = np.empty(256 + (int(n_blocks) + 1) * (int(2 * 1.5 * self.L * 16 // 8 // 4) + 256), dtype=np.uint8)
        rc = self.lib.mrc_encode_shard(self._ctx, _ptr(pcm), int(pcm_frame0), pcm.shape[0], int(total_frames),
                                       int(first_block), int(n_blocks), 1 if is_first else 0, 1 if is_last else 0,
                                       _ptr(out), out.size, C.byref(nbytes), cb, None)
        if err:
            raise err[0]
        self._check(rc)
        return out[:nbytes.value]

    # ---- batch decode ----------------------------------------------------------------------------
    def decode_batch(self, pac, byte_offsets, pcm_out=None):
        pac = np.ascontiguousarray(pac, dtype=np.uint8)
        byte_offsets = np.ascontiguousarray(byte_offsets, dtype=np.int64)
        n = len(byte_offsets) - 1
        foff = np.zeros(n + 1, dtype=np.int64)
        if pcm_out is None:
            # a chunk pair is at least ~2*(4+30) bytes, but size exactly: ask the library (NOSPACE reports sizes)
            rc = self.lib.mrc_decode_batch(self._ctx, _ptr(pac), _ptr(byte_offsets), n, None, 0, _ptr(foff))
            if rc not in (0, _lib.MRC_E_NOSPACE):
                self._check(rc)
            pcm_out = np.empty((int(foff[n]), 2), dtype=np.int16)
        self._check(self.lib.mrc_decode_batch(self._ctx, _ptr(pac), _ptr(byte_offsets), n, _ptr(pcm_out),
                                              pcm_out.shape[0], _ptr(foff)))
        return pcm_out, foff

    def decode_clips(self, blobs):
        arrs = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
        off = np.zeros(len(arrs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([a.size for a in arrs])
        pac = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
        pcm, foff = self.decode_batch(pac, off)
        return [pcm[foff[i]:foff[i + 1]].copy() for i in range(len(blobs))]

    def decode_batch_device(self, d_pac_ptr, h_pac, byte_offsets, d_pcm_ptr, pcm_cap_frames):
        byte_offsets = np.ascontiguousarray(byte_offsets, dtype=np.int64)
        n = len(byte_offsets) - 1
        foff = np.zeros(n + 1, dtype=np.int64)
        self._check(self.lib.mrc_decode_batch_device(self._ctx, C.c_void_p(d_pac_ptr), _ptr(h_pac), _ptr(byte_offsets),
                                                     n, C.c_void_p(d_pcm_ptr), int(pcm_cap_frames), _ptr(foff)))
        return foff

    # ---- stage taps --------------------------------------------------------------------------------
    def stage_analysis(self, clips):
        pcm, off = self._concat(clips)
        nb = sum(self.n_blocks(c) for c in np.diff(off))
        lines = np.zeros((nb, 4, self.L), np.float64)
        ovs = np.zeros((nb, 4), np.int32)
        ms = np.zeros((nb, self.n_bands), np.int32)
        smr = np.zeros((nb, 4, self.n_bands), np.float64)
        npk = np.zeros((nb, 4), np.int32)
        self._check(self.lib.mrc_stage_analysis(self._ctx, _ptr(pcm), _ptr(off), len(clips), _ptr(lines), _ptr(ovs),
                                                _ptr(ms), _ptr(smr), _ptr(npk)))
        return dict(mdct=lines, overallScale=ovs, ms_switch=ms, smr=smr, n_peaks=npk)

    def stage_alloc_quant(self, clips):
        pcm, off = self._concat(clips)
        nb = sum(self.n_blocks(c) for c in np.diff(off))
        ba = np.zeros((nb, 2, self.n_bands), np.int32)
        sf = np.zeros((nb, 2, self.n_bands), np.int32)
        mant = np.zeros((nb, 2, self.L), np.int32)
        ht = np.zeros((nb, 2), np.int32)
        res = np.zeros(nb, np.int32)
        cb = np.zeros((nb, 2), np.int32)
        self._check(self.lib.mrc_stage_alloc_quant(self._ctx, _ptr(pcm), _ptr(off), len(clips), _ptr(ba), _ptr(sf),
                                                   _ptr(mant), _ptr(ht), _ptr(res), _ptr(cb)))
        return dict(bitAlloc=ba, scaleFactor=sf, mantissa=mant, huffTable=ht, reservoir=res, chunkBytes=cb)

    def stage_reservoir(self, pcm, frame_offsets):
        """The serial stage's decisions only (no mantissa taps: cheap enough for hour-long streams): per block the
        Huffman table ids, codingParams.bitReservoir after the block and the two chunk sizes."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        off = np.ascontiguousarray(frame_offsets, dtype=np.int64)
        nb = sum(self.n_blocks(c) for c in np.diff(off))
        ht = np.zeros((nb, 2), np.int32)
        res = np.zeros(nb, np.int32)
        cb = np.zeros((nb, 2), np.int32)
        self._check(self.lib.mrc_stage_alloc_quant(self._ctx, _ptr(pcm), _ptr(off), len(off) - 1, None, None, None,
                                                   _ptr(ht), _ptr(res), _ptr(cb)))
        return dict(huffTable=ht, reservoir=res, chunkBytes=cb)

    # ---- per-block seam ------------------------------------------------------------------------------
    def encode_block(self, data, joint, reservoir, a=None, b=None):
        """data: float64 [2, a+b] (a = b = L by default).  Returns dict + new reservoir."""
        a = self.L if a is None else int(a)
        b = self.L if b is None else int(b)
        nl, nbands, _, _ = self.geometry(a, b)
        if int(joint) & 4:                      # one channel (nChannels = 1): outputs hold channel 0 only
            data = np.ascontiguousarray(data, dtype=np.float64).reshape(1, a + b)
        else:
            data = np.ascontiguousarray(data, dtype=np.float64).reshape(2, a + b)
        res = np.array([int(reservoir)], dtype=np.int32)
        sf = np.zeros((2, nbands), np.int32)
        ba = np.zeros((2, nbands), np.int32)
        mant = np.zeros((2, nl), np.int32)
        ovs = np.zeros(4, np.int32)
        ms = np.zeros(nbands, np.int32)
        ht = np.zeros(2, np.int32)
        cb = np.zeros(2, np.int32)
        self._check(self.lib.mrc_encode_block_ab(self._ctx, _ptr(data), a, b, int(joint), _ptr(res), _ptr(sf),
                                                 _ptr(ba), _ptr(mant), _ptr(ovs), _ptr(ms), _ptr(ht), _ptr(cb)))
        return dict(scaleFactor=sf, bitAlloc=ba, mantissa=mant, overallScale=ovs, ms_switch=ms, huffTable=ht,
                    chunkBytes=cb), int(res[0])

    def decode_block(self, joint, scaleFactor, bitAlloc, mantissa, overallScale, ms_switch=None, a=None, b=None):
        a = self.L if a is None else int(a)
        b = self.L if b is None else int(b)
        nl, nbands, _, _ = self.geometry(a, b)
        sf = np.ascontiguousarray(scaleFactor, dtype=np.int32).reshape(2, nbands)
        ba = np.ascontiguousarray(bitAlloc, dtype=np.int32).reshape(2, nbands)
        mant = np.ascontiguousarray(np.asarray(mantissa, dtype=np.int32)[:, :nl], dtype=np.int32).reshape(2, nl)
        ovs = np.zeros(4, np.int32)
        o = np.asarray(overallScale, dtype=np.int32).ravel()
        ovs[:o.size] = o
        ms = np.zeros(nbands, np.int32) if ms_switch is None else \
            np.ascontiguousarray(ms_switch, dtype=np.int32)
        out = np.zeros((2, a + b), np.float64)
        self._check(self.lib.mrc_decode_block_ab(self._ctx, a, b, 1 if joint else 0, _ptr(sf), _ptr(ba), _ptr(mant),
                                                 _ptr(ovs), _ptr(ms), _ptr(out)))
        return out

    # ---- Huffman table training front end ------------------------------------------
    def mantissa_histogram(self, clips, prior_max=-1):
        """calculateFrequencies over the EncodeNoHuff mantissas of `clips` (at most 16384 blocks per call).
        Returns (hist int64 [65536] since the last reset inside the call, largest value so far, reset flag)."""
        pcm, off = self._concat(clips)
        hist = np.zeros(65536, np.int64)
        mx = np.zeros(1, np.int32)
        rs = np.zeros(1, np.int32)
        self._check(self.lib.mrc_mantissa_histogram(self._ctx, _ptr(pcm), _ptr(off), len(clips), int(prior_max),
                                                    _ptr(hist), _ptr(mx), _ptr(rs)))
        return hist, int(mx[0]), bool(rs[0])

    def measure_peaks(self):
        out = np.zeros(4, np.float64)
        self._check(self.lib.mrc_measure_peaks(self._ctx, _ptr(out)))
        return dict(fp64_tflops=float(out[0]), fp32_tflops=float(out[1]), mufu_gops=float(out[2]),
                    copy_gbs=float(out[3]))

    def last_timing(self):
        ms = np.zeros(8, np.float64)
        cnt = np.zeros(8, np.int64)
        self.lib.mrc_last_timing(self._ctx, _ptr(ms), _ptr(cnt))
        return dict(analysis_ms=ms[0], chain_ms=ms[1], pack_ms=ms[2], decode_ms=ms[3], h2d_ms=ms[4], d2h_ms=ms[5],
                    total_ms=ms[6], cost_ms=ms[7], launches=int(cnt[0]), maskers=int(cnt[1]), blocks=int(cnt[2]),
                    waves=int(cnt[4]), chain_iters=int(cnt[3]), general_pairs=int(cnt[5]), window_adds=int(cnt[6]), loud_maskers=int(cnt[7]))
