"""Executed warp instructions per source line of one kernel from an .ncu-rep.  usage: src_instr.py rep nblocks [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; nblk = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = collections.OrderedDict()
for r in rows[3:]:
    if r and r[0].isdigit():
        try:
            k = (int(r[0]), r[1]); a = agg.get(k, [0, 0]); a[0] += int(r[4]); a[1] += int(r[7]); agg[k] = a
        except Exception:
            pass
ex = sum(v[1] for v in agg.values()); tot = sum(v[0] for v in agg.values())
print("warp instructions per block %.0f" % (ex / nblk))
cum = 0
for (ln, src), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    cum += e
    print("%5d %8.0f/blk %5.2f%% (cum %5.1f%%) smp %5.2f%%  %s" % (ln, e / nblk, 100.0 * e / ex, 100.0 * cum / ex, 100.0 * s / max(tot, 1), src[:100]))
