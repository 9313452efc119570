"""
make_golden.py -- generate tests/golden/*.npz by running the UNMODIFIED reference (through oracle/ref_shim.py)
on seeded synthetic PCM.  Run in the build container only (needs /root/reference); the fixtures are committed.

Each fixture holds: the input PCM, the coding parameters, the reference's .pac bytes, every integer the
reference's seam returned per block (overall scales, ms_switch, bit allocations, scale factors, line-aligned
mantissas, Huffman table ids, bit reservoir after the block), the reference-decoded PCM, and -- for the first
N_FLOAT_BLOCKS blocks -- the float64 outputs of the reference's MDCT() and CalcSMRs() calls (recorded by
wrapping those two names inside the shimmed codecThem module).

usage: python oracle/make_golden.py [name ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import ref_shim  # noqa: E402
import ref_driver  # noqa: E402
from mrcaudiocodec_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
N_FLOAT_BLOCKS = 10

CASES = {
    # name: (seed, seconds, sampleRate, joint, kbps per channel or None -> reference default 2.86 b/sample)
    "joint48k_128": (1, 1.0, 48000, True, 128),
    "joint48k_64": (2, 1.0, 48000, True, 64),
    "indep48k_128": (3, 0.5, 48000, False, 128),
    "joint44k_default": (4, 0.5, 44100, True, None),
    "indep48k_64": (5, 0.5, 48000, False, 64),
    # real material: cuts from the reference's own training WAVs (44.1 kHz stereo), the reference's default rate
    "wav_harps44k": (("wav", "tonal/harps.wav", 3.0), 0.6, 44100, True, None),
    "wav_speech44k": (("wav", "speech/spfg53_1.wav", 1.5), 0.6, 44100, True, None),
}


def read_wav_cut(rel, start_s, seconds):
    """int16 [n, 2] cut of one of /root/reference/training_data/*/*.wav (stdlib wave: plain RIFF parsing)."""
    import wave
    w = wave.open(os.path.join("/root/reference/training_data", rel), "rb")
    sr, nch = w.getframerate(), w.getnchannels()
    assert w.getsampwidth() == 2
    w.setpos(int(start_s * sr))
    x = np.frombuffer(w.readframes(int(seconds * sr)), dtype="<i2").reshape(-1, nch)
    w.close()
    if nch == 1:
        x = np.repeat(x, 2, axis=1)
    return np.ascontiguousarray(x[:, :2]), sr


def line_aligned(mant, bitAlloc, nLines, table, codes_escape=None):
    """compacted mantissa array (or code strings) -> int32[1024] aligned to MDCT lines."""
    out = np.zeros(int(np.sum(nLines)), np.int32)
    lo = 0
    i = 0
    for b, n in enumerate(nLines):
        n = int(n)
        if bitAlloc[b]:
            for j in range(n):
                v = mant[i]
                if isinstance(v, str):
                    v = codes_escape(v)
                out[lo + j] = int(v)
                i += 1
        lo += n
    return out


def run_case(name):
    seed, seconds, sr, joint, kbps = CASES[name]
    if isinstance(seed, tuple):
        pcm, wsr = read_wav_cut(seed[1], seed[2], seconds)
        assert wsr == sr, (wsr, sr)
    else:
        pcm = synth.synth_short(seed, seconds, sr)
    tbps = 2.86 if kbps is None else kbps * 1000.0 / sr
    m = ref_shim.load()
    codec = m["codecThem"]
    rec = {"mdct": [], "smr": []}
    oM, oS = codec.MDCT, codec.CalcSMRs

    def wM(*a, **k):
        r = oM(*a, **k)
        rec["mdct"].append(np.array(r, copy=True))
        return r

    def wS(*a, **k):
        r = oS(*a, **k)
        rec["smr"].append(np.array(r, copy=True))
        return r
    codec.MDCT, codec.CalcSMRs = wM, wS
    t0 = time.time()
    try:
        blob, blocks = ref_driver.ref_encode(pcm, sampleRate=sr, joint=joint, targetBitsPerSample=tbps)
    finally:
        codec.MDCT, codec.CalcSMRs = oM, oS
    t1 = time.time()
    dec = ref_driver.ref_decode(blob, joint=joint)
    t2 = time.time()

    import json
    with open(os.path.join(HERE, "mrc_oracle", "huffman_tables.json")) as fh:
        tabs = json.load(fh)["tables"]
    nB = len(blocks)
    nBands = len(blocks[0]["bitAlloc"][0])
    nLines = np.frombuffer(blob[26:26 + 2 * nBands], dtype='<u2').astype(np.int64)
    halfN = int(nLines.sum())
    ovs = np.zeros((nB, 4), np.int32)
    ms = np.zeros((nB, nBands), np.int32)
    ba = np.zeros((nB, 2, nBands), np.int32)
    sf = np.zeros((nB, 2, nBands), np.int32)
    mant = np.zeros((nB, 2, halfN), np.int32)
    ht = np.zeros((nB, 2), np.int32)
    res = np.zeros(nB, np.int64)
    isj = np.zeros(nB, np.int32)
    for i, b in enumerate(blocks):
        isj[i] = 1 if b["joint"] else 0
        o = b["overallScale"]
        ovs[i, :len(o)] = o
        if b["ms_switch"] is not None:
            ms[i] = b["ms_switch"]
        for c in range(2):
            ba[i, c] = b["bitAlloc"][c]
            sf[i, c] = b["scaleFactor"][c]
            ht[i, c] = b["huffTable"][c]
            t = b["huffTable"][c]
            if t == 15:
                esc = None
            else:
                rev = {v: int(k) for k, v in tabs[t]["codes"].items()}
                escc = tabs[t]["codes"][str(tabs[t]["escape"])]

                def esc(s, rev=rev, escc=escc):
                    p = s.split("/")
                    return int(p[1]) if p[0] == escc else rev[p[0]]
            mant[i, c] = line_aligned(b["mantissa"][c], b["bitAlloc"][c], nLines, t, esc)
        res[i] = b["reservoir"]
    # float taps: joint block = 4 MDCT + 4 CalcSMRs, non-joint block = 2 + 2, in call order
    nf = min(N_FLOAT_BLOCKS, nB)
    per = [4 if b["joint"] else 2 for b in blocks]
    mdct_f = np.zeros((nf, 4, halfN))
    smr_f = np.zeros((nf, 4, nBands))
    k = 0
    for i in range(nf):
        for c in range(per[i]):
            mdct_f[i, c] = rec["mdct"][k + c][:halfN]
            smr_f[i, c] = rec["smr"][k + c]
        k += per[i]
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pcm=pcm, sampleRate=sr, joint=int(joint), tbps=tbps,
                        pac=np.frombuffer(blob, np.uint8), overallScale=ovs, ms_switch=ms, bitAlloc=ba,
                        scaleFactor=sf, mantissa=mant, huffTable=ht, reservoir=res, isJoint=isj,
                        decoded=dec, mdct=mdct_f, smr=smr_f, nLines=nLines)
    tabs_used = sorted(set(ht.ravel().tolist()))
    print("%-18s blocks %3d  pac %6d B  tables %s  reservoir[min,max]=[%d,%d]  encode %.1fs decode %.1fs" %
          (name, nB, len(blob), tabs_used, res.min(), res.max(), t1 - t0, t2 - t1))


SWITCHED_CASES = {
    # name: (generator, seed, seconds, sampleRate, kbps per channel or None -> reference default 2.86 b/sample)
    "switched48k_128": ("percussive", 11, 1.0, 48000, 128),
    "switched44k_default": ("short", 12, 0.8, 44100, None),
}


def run_switched_case(name):
    """Block switching: the reference's own TransientDetector / JointWriteDataBlock / Close driven
    by its `__main__` loop (ref_driver.ref_encode_switched).  Stored: the canonical stream (last block written,
    Close with b = nMDCTLines), the stream exactly as the shipped loop writes it (last block dropped, Q11), the
    (a, b) of every written block, the detector's output per PCM block, the filter sections used (their last bits
    depend on the LAPACK build behind tf2sos), and the reference decoder's PCM for the canonical stream."""
    from scipy import signal
    gen, seed, seconds, sr, kbps = SWITCHED_CASES[name]
    pcm = synth.synth_percussive(seed, seconds, sr) if gen == "percussive" else synth.synth_short(seed, seconds, sr)
    tbps = 2.86 if kbps is None else kbps * 1000.0 / sr
    t0 = time.time()
    blob, geom, dets = ref_driver.ref_encode_switched(pcm, sampleRate=sr, targetBitsPerSample=tbps)
    blob_shipped, geom_shipped, _ = ref_driver.ref_encode_switched(pcm, sampleRate=sr, targetBitsPerSample=tbps,
                                                                   drop_last=True)
    t1 = time.time()
    dec = ref_driver.ref_decode(blob, joint=True)
    b_, a_ = signal.cheby2(20, 40, 9000. / sr, 'high')
    sos = signal.tf2sos(b_, a_)
    flags = np.zeros(len(dets), np.uint8)
    for i, d in enumerate(dets):
        flags[i] = (1 if np.any(d == 1) else 0) | (2 if np.any(d > 1) else 0)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), pcm=pcm, sampleRate=sr, tbps=tbps,
                        pac=np.frombuffer(blob, np.uint8), pac_shipped=np.frombuffer(blob_shipped, np.uint8),
                        geom=np.array(geom, np.int32), geom_shipped=np.array(geom_shipped, np.int32), flags=flags,
                        sos=sos, decoded=dec)
    ns = sum(1 for g in geom if g[1] == 128)
    print("%-20s pcm blocks %3d  written %3d (%d short)  pac %6d B  encode %.1fs" %
          (name, len(dets), len(geom), ns, len(blob), t1 - t0))


if __name__ == "__main__":
    for n in (sys.argv[1:] or (list(CASES) + list(SWITCHED_CASES))):
        if n in SWITCHED_CASES:
            run_switched_case(n)
        else:
            run_case(n)
