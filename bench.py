#!/usr/bin/env python
"""bench.py -- encoded audio-seconds per second on B200 (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # our arm (libmrc.so on the GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores

A step = one pass of the encode hot path (PCM frames -> .pac bytes) over one synthetic stream per GPU.
Workload at N=1: BASELINE.json configs[1], a 1 h 48 kHz stereo 16-bit stream (tones + noise + transients + silent
and -70 dBFS seconds), joint M/S, 128 kb/s/ch, fp64 code-exact mode.  N>1: every rank encodes its own 1 h stream
(weak scaling, no data-path collective), then one NCCL all-gather of the per-rank bitstream lengths.
`value` is measured with the PCM already in HBM and the bitstream left in HBM; `e2e` goes through the public host
API (mrc_encode_batch via mrcaudiocodec_b200.Codec.encode_batch) with pinned host buffers, H2D and D2H inside the
timed region.  The input (691 MB) is larger than L2 (126 MB), so nothing is cache-resident between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 48000
TBPS = 128000.0 / 48000.0
METRIC = "encoded audio-seconds/sec, 48 kHz stereo 128 kb/s/ch"
NCU_ANALYSIS_DRAM_BYTES_PER_BLOCK = 12489.0      # (23.42 + 46.84) MB / 5626 blocks, profiles/r01zz_ncu_full_summary.csv


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seconds", type=float, default=3600.0, help="stream length per GPU")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--spreading", default="factorised", choices=["factorised", "sequential"],
                    help="masker spreading: factorised (default) or pair by pair in the reference's order")
    ap.add_argument("--no-sequential-sample", action="store_true",
                    help="skip the short pair-by-pair run that is reported as roofline_sequential")
    ap.add_argument("--workload", default="stream", choices=["stream", "batch"],
                    help="stream: one clip of --seconds per GPU (BASELINE configs[1]); batch: the same audio cut into "
                         "independent clips of --clip-seconds (configs[3] shape)")
    ap.add_argument("--clip-seconds", type=float, default=30.0)
    ap.add_argument("--block-switching", action="store_true",
                    help="encode with the reference's transient detector / look-ahead loop (SURVEY 8 f1): eight "
                         "128-sample short blocks around transients instead of long blocks only")
    ap.add_argument("--decode", action="store_true", help="also time the decode mirror path on the encoded stream")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-seconds", type=float, default=1.5)
    return ap.parse_args()


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops(n_blocks_joint, n_blocks_flush, maskers, L=1024):
    """SURVEY.md §8d, reference formulation: per spectrum MDCT (N + 2N + 5N log2 N + 2N) + Hann FFT
    (N + 2.5 N log2 N + 1.5 N) + ~75 kFLOP for log10/quantise/MS, plus 40 FLOP per (masker, line) pair."""
    N = 2 * L
    lg = np.log2(N)
    per_spec = (N + 2 * N + 5 * N * lg + 2 * N) + (N + 2.5 * N * lg + 1.5 * N) + 75000.0
    spectra = 4 * n_blocks_joint + 2 * n_blocks_flush
    return spectra * per_spec + 40.0 * L * maskers


def run_reference(args, rank, world):
    """The reference algorithm (oracle port -- the Python-2 reference itself cannot run here) on the host
    cores: every worker process encodes its own bounded sample of the 1 h workload."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import multiprocessing as mp
    from mrcaudiocodec_b200 import synth
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = max(1, min(cores, 64))
    sample_s = args.cpu_sample_seconds
    pcm = synth.synth_clip(0, max(sample_s * cores, 10.0), fast=True)
    n = int(sample_s * SR)
    segs = [pcm[i * n:(i + 1) * n] for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(args.warmup, 0) and 1):
            pool.map(_oracle_encode, [s[:SR // 4] for s in segs])
        t0 = time.time()
        for _ in range(args.steps):
            pool.map(_oracle_encode, segs)
        dt = time.time() - t0
    audio = args.steps * cores * sample_s
    v = audio / dt
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "1 h synthetic 48 kHz stereo stream, joint M/S, 128 kb/s/ch, fp64 (configs[1]); "
                                   "each step = %d independent %.1f s samples of it, one per host core" % (cores, sample_s)},
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "port",
                             "sample": "%d x %.1f s segments per step, %d steps (oracle/mrc_oracle, numpy)" %
                                       (cores, sample_s, args.steps)},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def _oracle_encode(pcm):
    import mrc_oracle as o
    blob, _ = o.driver.encode_pcm(pcm, joint=True)
    return len(blob)


def cpu_baseline(pcm, seconds):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mrc_oracle as o
    n = int(seconds * SR)
    o.driver.encode_pcm(pcm[:2048], joint=True)          # warm the table caches
    t0 = time.time()
    o.driver.encode_pcm(pcm[:n], joint=True)
    dt = time.time() - t0
    return {"value": seconds / dt, "unit": "audio-s/s", "cores": 1, "kind": "port",
            "sample": "first %.1f s of the same stream, oracle/mrc_oracle (numpy restatement of the reference), "
                      "1 thread, %.1f s of CPU time" % (seconds, dt)}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mrcaudiocodec_b200 import Codec, synth
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    seconds = args.seconds
    threads = max(1, min(16, (os.cpu_count() or 8) // max(world, 1)))
    pcm = synth.synth_clip(rank, seconds, threads=threads, fast=True)            # [frames, 2] int16
    frames = pcm.shape[0]
    if args.workload == "batch":
        cf = int(round(args.clip_seconds * SR))
        off = np.unique(np.append(np.arange(0, frames, cf), frames)).astype(np.int64)
    else:
        off = np.array([0, frames], dtype=np.int64)
    n_clips = len(off) - 1
    codec = Codec(device=local, precision=args.precision, spreading=args.spreading,
                  block_switching=args.block_switching)
    L = codec.L
    nblk = int(sum(codec.n_blocks(f) for f in np.diff(off)))

    # device-resident buffers (torch is plumbing: device memory + NCCL)
    d_pcm = torch.from_numpy(pcm).to(dev)
    cap = int(2.0 * TBPS * 2 * frames / 8) + (1 << 20)
    d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
    h_pcm = torch.from_numpy(pcm).pin_memory()
    h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    h_pcm_np, h_out_np = h_pcm.numpy(), h_out.numpy()
    lens = torch.zeros(1, dtype=torch.int64, device=dev)
    gathered = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)] if world > 1 else None

    last_boff = [None]

    def step_device():
        boff = codec.encode_batch_device(d_pcm.data_ptr(), off, d_out.data_ptr(), cap)
        if world > 1:                    # the path's only collective: per-shard bitstream lengths -> offsets
            lens[0] = int(boff[-1])
            dist.all_gather(gathered, lens)
        return int(boff[-1])

    def step_e2e():
        out, boff = codec.encode_batch(h_pcm_np, off, out=h_out_np)
        last_boff[0] = boff
        if world > 1:
            lens[0] = int(boff[-1])
            dist.all_gather(gathered, lens)
        return int(boff[-1])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        dev_ms, launches, maskers, stage = 0.0, 0, 0, np.zeros(4)
        extra = {}
        t0 = time.perf_counter()
        for _ in range(args.steps):
            nbytes = fn()
            t = codec.last_timing()
            dev_ms += t["total_ms"]
            launches += t["launches"]
            maskers = t["maskers"]
            work = {k: t[k] for k in ("general_pairs", "window_adds", "loud_maskers", "waves", "chain_iters")}
            stage += np.array([t["analysis_ms"], t["cost_ms"], t["chain_ms"], t["pack_ms"]])
            extra["blocks_written"] = t["blocks"]
            extra["transient_ms"] = t["decode_ms"]
        barrier()
        wall = time.perf_counter() - t0
        clocks = sampler.stop()
        tt = torch.tensor([wall, dev_ms / 1000.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return dict(wall=float(tt[0]), dev=float(tt[1]), launches=launches, maskers=maskers, nbytes=nbytes, work=work,
                    stage_ms=(stage / args.steps).tolist(), clocks=clocks, extra=extra)

    peaks = codec.measure_peaks()
    r_dev = timed(step_device)
    r_e2e = timed(step_e2e)

    audio_total = world * seconds * args.steps
    value = audio_total / r_dev["wall"]
    e2e_value = audio_total / r_e2e["wall"]

    if rank == 0:
        mp_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(mp_path):
            try:
                hbm_peak = float(json.load(open(mp_path))["hbm_gbs"])
                hbm_src = "MEASURED_PEAKS.json"
            except Exception:
                pass
        an_ms = r_dev["stage_ms"][0]
        flops = algorithmic_flops(nblk - n_clips, n_clips, r_dev["maskers"], L)
        is64 = args.precision == "fp64"
        peak_tf = peaks["fp64_tflops"] if is64 else peaks["fp32_tflops"]
        ach_tf = flops / (an_ms * 1e-3) / 1e12
        alg_bytes = nblk * (4096 + 2 * L * (8 if is64 else 4) * 1 + 64 * 2 * (8 if is64 else 4) + 1536 + 8)
        line = {
            "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * r_dev["wall"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if is64 else "f32", "data": "synthetic",
            "config": {"workload": ("%.0f s synthetic 48 kHz stereo 16-bit stream per GPU (BASELINE configs[1] = 1 h), "
                                    "joint M/S, 128 kb/s/ch, %s mode, long blocks N=2048" % (seconds, args.precision))
                       if args.workload == "stream" else
                       ("%d independent clips of %.0f s (%.0f s of synthetic 48 kHz stereo audio per GPU, BASELINE "
                        "configs[3] shape), joint M/S, 128 kb/s/ch, %s mode" % (n_clips, args.clip_seconds, seconds,
                                                                                 args.precision)),
                       "l2": "input %.0f MB per step > 126 MB L2, no flush needed" % (frames * 4 / 1e6),
                       "blocks_per_step": nblk, "bitstream_bytes": r_dev["nbytes"]},
            "device_ms_per_step": 1000.0 * r_dev["dev"] / args.steps,
            "stage_ms_per_step": {"analysis": r_dev["stage_ms"][0], "cost": r_dev["stage_ms"][1],
                                  "chain": r_dev["stage_ms"][2], "pack": r_dev["stage_ms"][3],
                                  "note": "per-kernel sums; analysis+cost of wave w+1 overlap chain+pack of wave w"},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(frames * 4),
                    "d2h_bytes_per_step": int(r_e2e["nbytes"]), "ms_per_step": 1000.0 * r_e2e["wall"] / args.steps},
            "gpu_launches": int(r_dev["launches"]),
            "clocks": r_dev["clocks"],
            "roofline": {"kernel": "analysis_kernel (MDCT + psychoacoustics, fused)",
                         "bound": "fp64" if is64 else "fp32", "achieved": ach_tf, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf else None,
                         # dram__bytes_read + dram__bytes_write of one analysis launch, from the committed ncu capture
                         # (profiles/r01zz_ncu_full_summary.csv: 70.3 MB for 5626 blocks), scaled to this run's launch size
                         "traffic": NCU_ANALYSIS_DRAM_BYTES_PER_BLOCK * nblk / max(r_dev["work"]["waves"], 1),
                         "peak_source": "mrc_measure_peaks (FMA chain on this GPU, this run)",
                         "work": "SURVEY 8d reference formulation (40 FLOP per masker-line pair), %d maskers measured; "
                                 "the %s evaluation executes the work listed under executed_work" %
                                 (r_dev["maskers"], args.spreading),
                         "launches_per_step": r_dev["work"]["waves"],
                         "avg_launch_ms": an_ms / max(r_dev["work"]["waves"], 1), "ms_per_step": an_ms},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (an_ms * 1e-3) / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": alg_bytes / (an_ms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": hbm_src,
                             "traffic": NCU_ANALYSIS_DRAM_BYTES_PER_BLOCK * nblk / max(r_dev["work"]["waves"], 1)},
            "pipe_peaks": peaks,
            "executed_work": r_dev["work"],
        }
        if args.block_switching:
            line["config"]["workload"] += (", BLOCK SWITCHING on (transient detector + look-ahead, short blocks of 128: "
                                           "%d blocks written for %d blocks of 1024 frames)" %
                                           (r_dev["extra"]["blocks_written"], nblk))
            line["stage_ms_per_step"]["transient_detector"] = r_dev["extra"]["transient_ms"]
            line["roofline"]["work"] += "; work formula evaluated on the long-block equivalent of the stream"
        if args.spreading == "factorised" and not args.no_sequential_sample and not args.block_switching:
            # the same analysis kernel summing the maskers pair by pair in the reference's order, on a bounded
            # sample of the same stream: this is the kernel SURVEY 8d's 40-FLOP-per-pair work formula describes
            sample_s = min(seconds, 600.0)
            nfr = int(sample_s * SR)
            cs = Codec(device=local, precision=args.precision, spreading="sequential")
            soff = np.array([0, nfr], dtype=np.int64)
            for _ in range(2):
                cs.encode_batch_device(d_pcm.data_ptr(), soff, d_out.data_ptr(), cap)
            ts = cs.last_timing()
            sflops = algorithmic_flops(cs.n_blocks(nfr) - 1, 1, ts["maskers"], L)
            s_tf = sflops / (ts["analysis_ms"] * 1e-3) / 1e12
            line["roofline_sequential"] = {"kernel": "analysis_kernel, MRC_FLAG_SPREAD_SEQUENTIAL", "bound": line["roofline"]["bound"],
                                           "achieved": s_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": s_tf / peak_tf,
                                           "sample": "first %.0f s of the stream, 1 launch" % sample_s,
                                           "avg_launch_ms": ts["analysis_ms"]}
            cs.close()
        if args.decode:
            # the mirror path (rows a14-a17): .pac bytes in pinned host memory -> int16 PCM in host memory
            nbytes = int(last_boff[0][-1])
            pac = h_out_np[:nbytes]
            h_dec = torch.empty(((nblk + 8) * L, 2), dtype=torch.int16).pin_memory()
            pcm_out = h_dec.numpy()
            for _ in range(2):
                codec.decode_batch(pac, last_boff[0], pcm_out=pcm_out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                codec.decode_batch(pac, last_boff[0], pcm_out=pcm_out)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            td = codec.last_timing()
            line["decode"] = {"e2e_value": seconds / dt, "unit": "audio-s/s", "ms_per_step": 1000.0 * dt,
                              "kernel_ms": td["decode_ms"], "h2d_bytes_per_step": nbytes,
                              "d2h_bytes_per_step": int(pcm_out.nbytes)}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(pcm, args.cpu_sample_seconds)
        print(json.dumps(line))
    codec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
