#!/bin/bash
# SASS evidence for profiles/: the full listing of the hot kernels (fp64, L=1024) and a mnemonic histogram per kernel.
# usage: scripts/sass_listing.sh <tag>
TAG=${1:-r01}
SO=mrcaudiocodec_b200/libmrc.so
cuobjdump -sass $SO > /tmp/libmrc.sass
python - "$TAG" <<'PY'
import re, sys, collections
tag = sys.argv[1]
txt = open('/tmp/libmrc.sass').read().split('\n')
funcs = []
cur = None
for ln in txt:
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = [m.group(1), []]; funcs.append(cur)
    elif cur is not None:
        cur[1].append(ln)
def short(name):
    for k in ("analysis_kernelIdLi1024ELb0", "analysis_kernelIfLi1024ELb0", "cost_kernelId", "chain_table_kernel", "chain_seg_kernel", "segment_kernel", "extras_kernel", "expand_kernel", "table_kernel", "chain_kernel", "parse_kernel",
              "finish_kernel", "offsets_kernel", "pack_kernelId", "decode_kernelIdLi1024", "ola_kernelId", "clip_scan_kernel", "analysis_kernelIdLi576", "analysis_kernelIdLi128", "peaks_kernelILi10", "decide_kernel"):
        if k in name: return k
    return None
hist_lines = ["kernel,instructions," + ",".join(["DFMA", "DADD", "DMUL", "DSETP", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "REDUX", "VOTE", "ATOMS", "UBLKCP", "SYNCS", "HMMA", "UTC"])]
keep = {}
for name, body in funcs:
    s = short(name)
    if s is None: continue
    ops = collections.Counter()
    n = 0
    for ln in body:
        m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
        if m:
            n += 1
            ops[m.group(1).split('.')[0]] += 1
    cols = ["DFMA", "DADD", "DMUL", "DSETP", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "REDUX", "VOTE", "ATOMS", "UBLKCP", "SYNCS", "HMMA"]
    utc = sum(v for k, v in ops.items() if k.startswith("UTC"))
    hist_lines.append("%s,%d,%s,%d" % (s, n, ",".join(str(ops[c]) for c in cols), utc))
    if s in ("analysis_kernelIdLi1024ELb0", "chain_seg_kernel", "chain_kernel"):
        keep[s] = body
open('profiles/%s_sass_mnemonics.csv' % tag, 'w').write("\n".join(hist_lines) + "\n")
import os
for s, body in (keep.items() if os.environ.get("SASS_FULL_LISTINGS") else ()):     # ~350 KB each: only on request
    with open('profiles/%s_sass_%s.txt' % (tag, s), 'w') as fh:
        for ln in body:
            m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*/\*', ln)
            if m: fh.write("%s  %s\n" % (m.group(1), m.group(2).rstrip(' ;')))
print("\n".join(hist_lines))
PY
