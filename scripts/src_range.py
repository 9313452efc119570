"""Stall samples and executed warp instructions of source-line ranges of mrc_analysis.cu (ranges given as a:b ...)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; nblk = float(sys.argv[2]); ranges = [tuple(map(int, a.split(':'))) for a in sys.argv[3:]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
src = open('mrcaudiocodec_b200/csrc/mrc_analysis.cu').read().split('\n')
tot = ex = 0
per = collections.defaultdict(lambda: [0, 0])
other = [0, 0]
for r in rows[3:]:
    if r and r[0].isdigit():
        try:
            ln, s, e = int(r[0]), int(r[4]), int(r[7])
        except Exception:
            continue
        tot += s; ex += e
        # only lines whose text matches this file (inlined headers carry their own numbering)
        mine = 0 < ln <= len(src) and src[ln - 1].strip()[:40] == r[1].strip()[:40]
        hit = False
        if mine:
            for a, b in ranges:
                if a <= ln <= b:
                    per[(a, b)][0] += s; per[(a, b)][1] += e; hit = True; break
        if not hit:
            key = 'headers' if not mine else 'rest'
            per[key][0] += s; per[key][1] += e
print("samples %d, warp instr per block %.0f" % (tot, ex / nblk))
for k, (s, e) in per.items():
    print("%-14s samples %5.1f%%   instr %5.1f%% (%6.0f/blk)" % (str(k), 100.0 * s / tot, 100.0 * e / ex, e / nblk))
