"""dev: compare the segment-composed chain against the per-block table walk and the plain walk on switched streams"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mrcaudiocodec_b200 import Codec, synth, pacfile

def enc(clips, seg, tables=True, switching=True):
    if MIN < 5:
        os.environ["MRC_CHAIN_TABLE_MIN_BLOCKS"] = "1"
    os.environ["MRC_CHAIN_SEGMENT_BLOCKS"] = str(seg)
    c = Codec(block_switching=switching, chain_tables=tables)
    b = c.encode_clips(clips)
    t = c.last_timing()
    c.close()
    return b, t

MIN = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
perc = np.concatenate([synth.synth_percussive(200 + i, 30.0) for i in range(int(2 * MIN))], axis=0)
stream = synth.synth_clip(3, 60 * MIN, threads=8, fast=True)
for name, clips, sw in (("perc", [perc], True), ("perc+stream", [perc, stream], True), ("stream-long", [stream], False)):
    ref, _ = enc(clips, 0, tables=False, switching=sw)
    for seg in (0, 5, 32, 32):
        got, t = enc(clips, seg, switching=sw)
        for ci, (a, b) in enumerate(zip(ref, got)):
            if a == b:
                print(name, "seg", seg, "clip", ci, "OK", len(a), "slow walks", t["chain_iters"])
                continue
            ia, ib = pacfile.chunk_index(a), pacfile.chunk_index(b)
            k = next(i for i in range(min(len(ia), len(ib))) if ia[i] != ib[i] or a[ia[i][0]:ia[i][0] + ia[i][1]] != b[ib[i][0]:ib[i][0] + ib[i][1]])
            print(name, "seg", seg, "clip", ci, "DIFF first differing chunk", k, "block", k // 2, "of", len(ia) // 2,
                  "sizes", ia[k], ib[k], "geom bits", (a[ia[k][0]] >> 2) & 3)
            # geometry of the neighbourhood
            print("   geoms around:", [(a[ia[2 * j][0]] >> 2) & 3 for j in range(max(0, k // 2 - 6), min(len(ia) // 2, k // 2 + 4))])
