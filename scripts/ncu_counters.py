#!/usr/bin/env python
"""Turn ncu captures (read here, no GPU needed) into the per-block counters bench.py's roofline uses.

  python scripts/ncu_counters.py --full gpurun_out/X_prof.ncu-rep --wave gpurun_out/X_wave.csv --out profiles/analysis_counters.json

--full : `ncu --set full` capture holding one analysis_kernel launch of a full wave (B200_PROFILING.md recipe).  From
         it: executed FP64 flops (2 x DFMA + DADD + DMUL thread instructions), FP64 pipe and issue-slot utilisation,
         DRAM bytes -- all divided by the launch's blocks (grid size).
--wave : `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of the
         same command: DRAM bytes of every kernel of one full wave, for the pipeline-traffic figure.
Nothing here is a bench value: the counters are per-block properties of (kernel, workload) that bench.py multiplies
by its own live timings."""
import argparse
import csv
import json
import re
import subprocess


def raw_rows(rep, kernel_regex=None):
    cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
    if kernel_regex:
        cmd += ["--kernel-name", "regex:" + kernel_regex]
    raw = subprocess.run(cmd, stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v, unit):
    m = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1.0)


def grid_blocks(s):
    return int(re.findall(r"\d+", s)[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", required=True)
    ap.add_argument("--wave", default="")
    ap.add_argument("--out", required=True)
    ap.add_argument("--kernel", default="analysis_kernel<double")
    ap.add_argument("--tag", default="")
    ap.add_argument("--precision", default="fp64", help="key of --out the counters are stored under (fp64 | fp32)")
    ap.add_argument("--blocks-per-cta", type=int, default=2,
                    help="blocks one analysis CTA takes (MRC_BLOCKS_PER_CTA of the build): blocks of a launch = grid x this")
    a = ap.parse_args()
    hdr, units, rows = raw_rows(a.full, a.kernel.split("<")[0])
    idx = {h: i for i, h in enumerate(hdr)}
    rows = [x for x in rows if a.kernel in x[idx["Kernel Name"]]]
    r = max(rows, key=lambda x: grid_blocks(x[idx["Grid Size"]]))       # the largest launch captured
    g = lambda name: float(r[idx[name]])
    blocks = grid_blocks(r[idx["Grid Size"]]) * a.blocks_per_cta
    cyc = g("sm__cycles_elapsed.max")
    per_cycle = lambda op: g("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op)
    dfma, dadd, dmul = per_cycle("dfma") * cyc, per_cycle("dadd") * cyc, per_cycle("dmul") * cyc
    ffma, fadd, fmul = per_cycle("ffma") * cyc, per_cycle("fadd") * cyc, per_cycle("fmul") * cyc
    flops = 2.0 * dfma + dadd + dmul
    flops32 = 2.0 * ffma + fadd + fmul
    dram = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
        to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
    out = {
        "kernel": r[idx["Kernel Name"]][:80], "source": a.full.replace("gpurun_out/", "profiles/ (summary of) "), "tag": a.tag,
        "launch_blocks": blocks, "blocks_per_cta": a.blocks_per_cta, "launch_ms_under_ncu": g("gpu__time_duration.sum"),
        "fp64_thread_inst_per_block": {"dfma": dfma / blocks, "dadd": dadd / blocks, "dmul": dmul / blocks},
        "fp64_flop_per_block": flops / blocks,
        "fp32_thread_inst_per_block": {"ffma": ffma / blocks, "fadd": fadd / blocks, "fmul": fmul / blocks},
        "fp32_flop_per_block": flops32 / blocks,
        "pipe_fma_active_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active")
        if "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active" in idx else None,
        "warp_inst_per_block": g("smsp__inst_executed.sum") / blocks,
        "pipe_fp64_active_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "smem_bank_conflicts_per_block": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / blocks,
        "dram_bytes_per_block": dram / blocks,
        "registers": int(g("launch__registers_per_thread")),
    }
    if a.wave:
        # launch list: "ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size",
        # "Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
        per = {}
        with open(a.wave) as fh:
            rd = csv.reader(l for l in fh if l.startswith('"'))
            h = next(rd)
            ix = {n: i for i, n in enumerate(h)}
            for row in rd:
                k = (int(row[ix["ID"]]), row[ix["Kernel Name"]].split("(")[0], grid_blocks(row[ix["Grid Size"]]))
                per.setdefault(k, {})[row[ix["Metric Name"]]] = to_bytes(row[ix["Metric Value"]].replace(",", ""),
                                                                           row[ix["Metric Unit"]])
        # the wave whose analysis launch has the most blocks; its kernels are the launches between that analysis launch and
        # the next one on the list (ncu serialises the streams: a wave's launches are contiguous up to interleaving with
        # its neighbour, so kernels are attributed by name, one launch of each per wave, taking the largest grid)
        best = {}
        for (i, name, grid), m in per.items():
            b = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
            t = m.get("gpu__time_duration.sum", 0.0)
            short = name.split("::")[-1].split("<")[0]
            if short not in ("analysis_kernel", "cost_kernel", "table_kernel", "segment_kernel", "extras_kernel",
                             "chain_seg_kernel", "chain_kernel", "expand_kernel", "finish_kernel", "offsets_kernel",
                             "clip_scan_kernel", "pack_kernel"):
                continue                      # e.g. the pipe micro-benchmarks of mrc_measure_peaks
            if short not in best or grid > best[short]["grid"]:
                best[short] = {"grid": grid, "dram_bytes": b, "time_ns": t}
        wave_blocks = best["analysis_kernel"]["grid"] * a.blocks_per_cta if "analysis_kernel" in best else blocks
        out["wave"] = {"blocks": wave_blocks, "kernels": best,
                       "dram_bytes_per_block": sum(v["dram_bytes"] for v in best.values()) / wave_blocks}
    import os
    allp = {}
    if os.path.exists(a.out):
        try:
            allp = json.load(open(a.out))
        except Exception:
            allp = {}
    allp[a.precision] = out
    with open(a.out, "w") as fh:
        json.dump(allp, fh, indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
