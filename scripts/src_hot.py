"""Per-source-line stall samples of one kernel from an .ncu-rep (cuda,sass view).  usage: src_hot.py rep [top]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = collections.OrderedDict()
for r in rows[3:]:
    if r and r[0].isdigit():
        try:
            k = (int(r[0]), r[1]); a = agg.get(k, [0, 0]); a[0] += int(r[4]); a[1] += int(r[7]); agg[k] = a
        except Exception:
            pass
tot = sum(v[0] for v in agg.values()); ex = sum(v[1] for v in agg.values())
print("total samples", tot, "warp instructions", ex)
for (ln, src), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d %6.2f%% %12d  %s" % (ln, 100.0 * s / max(tot, 1), e, src[:120]))
