"""Shared-memory wavefronts per source line of one kernel from an .ncu-rep (cuda,sass view): where the L1/shared data
pipe goes.  usage: src_smem.py rep nblocks [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; nblk = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[2]
ix = {n: i for i, n in enumerate(hdr)}
W, I, E, G = ix["L1 Wavefronts Shared"], ix["L1 Wavefronts Shared Ideal"], ix["L1 Wavefronts Shared Excessive"], ix["L2 Theoretical Sectors Global"]
agg = collections.OrderedDict()
for r in rows[3:]:
    if r and r[0].isdigit():
        try:
            k = (int(r[0]), r[1]); a = agg.setdefault(k, [0, 0, 0, 0])
            for j, c in enumerate((W, I, E, G)):
                a[j] += int(r[c] or 0)
        except Exception:
            pass
tw = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values()); tg = sum(v[3] for v in agg.values())
print("shared wavefronts per block %.0f (ideal %.0f), global sectors per block %.0f" % (tw / nblk, ti / nblk, tg / nblk))
print(" line  wavefronts/blk  ideal/blk  gsect/blk  source")
for (ln, src), v in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][3]))[:top]:
    print("%5d %10.0f %10.0f %10.0f  %s" % (ln, v[0] / nblk, v[1] / nblk, v[3] / nblk, src[:110]))
