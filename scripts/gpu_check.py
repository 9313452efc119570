"""Diagnostic parity run on a GPU box: golden fixtures vs libmrc stage taps and bytes.  Prints mismatch counts."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import Codec  # noqa: E402

CASES = ["joint48k_128", "joint48k_64", "indep48k_128", "joint44k_default", "indep48k_64"]


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp64"
    for name in CASES:
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        sr, joint, tbps = int(g["sampleRate"]), bool(g["joint"]), float(g["tbps"])
        t0 = time.time()
        c = Codec(sample_rate=sr, joint=joint, target_bits_per_sample=tbps, precision=prec)
        t1 = time.time()
        a = c.stage_analysis([g["pcm"]])
        nB = g["reservoir"].shape[0]
        nf = g["mdct"].shape[0]
        isj = g["isJoint"]
        print("== %s  (create %.2fs) blocks %d" % (name, t1 - t0, nB))
        for i in range(nf):
            ns = 4 if isj[i] else 2
            ref = g["mdct"][i, :ns]
            got = a["mdct"][i, :ns]
            err = np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)
            serr = np.abs(a["smr"][i, :ns] - g["smr"][i, :ns]).max()
            if i < 3 or err > 1e-12 or serr > 1e-6:
                print("  blk %d mdct rel err %.2e  smr abs err %.2e  peaks %s" % (i, err, serr, a["n_peaks"][i]))
        ovs_bad = 0
        for i in range(nB):
            ns = 4 if isj[i] else 2
            ovs_bad += int(np.any(a["overallScale"][i, :ns] != g["overallScale"][i, :ns]))
        ms_bad = int(np.sum(np.any(a["ms_switch"] != g["ms_switch"], axis=1) & (isj == 1)))
        print("  overallScale mismatching blocks %d, ms_switch mismatching blocks %d" % (ovs_bad, ms_bad))
        q = c.stage_alloc_quant([g["pcm"]])
        for k in ("bitAlloc", "scaleFactor", "mantissa", "huffTable", "reservoir"):
            ref = g[k]
            got = q[k]
            bad = np.nonzero(np.any((got != ref).reshape(nB, -1), axis=1))[0]
            print("  %-12s mismatching blocks %d %s" % (k, bad.size, bad[:8].tolist()))
            if k == "reservoir" and bad.size:
                print("     got", got[bad[:6]].tolist(), "ref", ref[bad[:6]].tolist())
        blob = c.encode_clips([g["pcm"]])[0]
        ref = g["pac"].tobytes()
        same = blob == ref
        print("  bytes: got %d ref %d identical %s" % (len(blob), len(ref), same))
        if not same:
            n = min(len(blob), len(ref))
            d = np.nonzero(np.frombuffer(blob[:n], np.uint8) != np.frombuffer(ref[:n], np.uint8))[0]
            print("     first differing byte offsets", d[:10].tolist())
        print("  timing", c.last_timing())
        c.close()


if __name__ == "__main__":
    main()
