"""Stall samples / executed warp instructions of analysis_kernel aggregated by code region. usage: src_regions.py rep nblocks"""
import bisect, collections, csv, subprocess, sys
rep = sys.argv[1]; nblk = float(sys.argv[2]) if len(sys.argv) > 2 else 5626.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
lines = []
for r in rows[3:]:
    if r and r[0].isdigit():
        try: lines.append((int(r[0]), int(r[4]), int(r[7])))
        except Exception: pass
tot = sum(l[1] for l in lines); ex = sum(l[2] for l in lines)
src = open('mrcaudiocodec_b200/csrc/mrc_analysis.cu').read().split('\n')
keys = ["fft_dit(", "warp_max(", "tsample(", "masker_range(", "loud_term(", "tail_terms(", "spread_line_bound(", "spread_line_warp(",
        "analysis_kernel(", "phase 0", "phase 1", "phase 2", "phase 3", "phase 4", "// a. Hann", "// b. X[k]", "// c. strict", "// d. masker",
        "count table over Bark", "// e. masked", "pass 1 (thread", "// pass 2a", "// pass 2b", "phase 5", "phase 6"]
marks = []
for i, l in enumerate(src, 1):
    for k in keys:
        if k in l and (l.strip().startswith("//") or "__device__" in l or "__global__" in l or "analysis_kernel(" in l or "for (int sg" in l):
            marks.append((i, k)); break
marks.sort(); bounds = [m[0] for m in marks]
agg = collections.OrderedDict()
for ln, smp, e in lines:
    i = bisect.bisect_right(bounds, ln) - 1
    name = marks[i][1] if i >= 0 else 'top(math helpers)'
    a = agg.setdefault(name, [0, 0]); a[0] += smp; a[1] += e
print("total: %d samples, %.0f k warp-instr per block" % (tot, ex / nblk / 1000))
for k, (s_, e) in agg.items():
    print("%-26s samples %5.1f%%  instr %5.1f%%  (%5.1f k instr/block)" % (k, 100 * s_ / tot, 100 * e / ex, e / nblk / 1000))
