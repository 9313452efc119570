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

roofline: the dominant kernel is analysis_kernel (fused MDCT + psychoacoustic model), an FP64 kernel.  `achieved` =
the FP64 flops the kernel EXECUTES per block (2 x DFMA + DADD + DMUL thread instructions of one full-wave launch,
counted by ncu and committed as profiles/analysis_counters.json by scripts/ncu_counters.py) x the blocks of this run
/ the kernel's launch durations measured live with CUDA events on its stream; `peak` = this GPU's FP64 FMA rate
measured in the same run (mrc_measure_peaks).  The reference-formulation work (40 flops per masker-line
pair) is reported beside it as `work_reduction_vs_reference`, not as throughput.
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
ALGORITHMIC_BYTES_PER_BLOCK = 4096 + 683 + 8       # PCM in + bitstream + length prefixes
COUNTERS = os.path.join(ROOT, "profiles", "analysis_counters.json")


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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1 GPUs: weak = every rank encodes its own stream of --seconds (the driver's contract); strong = "
                         "ONE stream of --seconds sharded by block range across the ranks (N/2-sample halo from the PCM, "
                         "the reservoir handed rank to rank as one int32; mrcaudiocodec_b200.dist.encode_stream_sharded)")
    ap.add_argument("--no-decode", action="store_true", help="skip the decode mirror path (timed by default)")
    ap.add_argument("--no-music", action="store_true",
                    help="skip the second encode measurement on dense-masker music-like material (synth_music)")
    ap.add_argument("--music-seconds", type=float, default=600.0)
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
                                   "each step = %d independent %.1f s samples of it, one per host core; per-HOST figure: "
                                   "the arm uses all cores of the box whatever --gpus says (%d)" %
                                   (cores, sample_s, args.gpus)},
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
    strong = args.scaling == "strong" and world > 1
    codec = Codec(device=local, precision=args.precision, spreading=args.spreading,
                  block_switching=args.block_switching)
    L = codec.L
    relay = None
    if strong:
        # one stream for the whole job: this rank synthesises only the frames of its block range and their halo
        from mrcaudiocodec_b200 import dist as mdist
        if os.environ.get("MRC_SHARD_RELAY", "gloo") == "gloo":
            relay = dist.new_group(backend="gloo")     # the reservoir hand-off (one int, host to host) as a CPU message
        total_frames = int(round(seconds * SR))
        nblk_stream = (total_frames + L - 1) // L
        blk_lo, blk_hi = mdist.stream_shard_range(nblk_stream, rank, world)
        f_lo, f_hi = codec.shard_pcm_range(total_frames, blk_lo, blk_hi - blk_lo)
        pcm = synth.synth_range(0, f_lo, f_hi, seconds, threads=threads, fast=True)
        frames = pcm.shape[0]
    else:
        pcm = synth.synth_clip(rank, seconds, threads=threads, fast=True)            # [frames, 2] int16
        frames = pcm.shape[0]
    if args.workload == "batch":
        cf = int(round(args.clip_seconds * SR))
        off = np.unique(np.append(np.arange(0, frames, cf), frames)).astype(np.int64)
    else:
        off = np.array([0, frames], dtype=np.int64)
    n_clips = len(off) - 1
    nblk = int(sum(codec.n_blocks(f) for f in np.diff(off)))
    if strong:
        nblk = blk_hi - blk_lo + (1 if rank == world - 1 else 0)

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
        if strong:
            n, _ = mdist.encode_stream_sharded(codec, None, f_lo, total_frames, device=dev, relay_group=relay,
                                               device_ptrs=(d_pcm.data_ptr(), frames, d_out.data_ptr(), cap))
            return int(n)
        boff = codec.encode_batch_device(d_pcm.data_ptr(), off, d_out.data_ptr(), cap)
        if world > 1:                    # the path's only collective: per-shard bitstream lengths -> offsets
            lens[0] = int(boff[-1])
            dist.all_gather(gathered, lens)
        return int(boff[-1])

    def step_e2e():
        if strong:
            blob, _ = mdist.encode_stream_sharded(codec, h_pcm_np, f_lo, total_frames, device=dev, out=h_out_np,
                                                  relay_group=relay)
            return int(blob.size)
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

    audio_total = (1 if strong else world) * seconds * args.steps
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
        is64 = args.precision == "fp64"
        peak_tf = peaks["fp64_tflops"] if is64 else peaks["fp32_tflops"]
        nwaves = max(r_dev["work"]["waves"], 1)
        # executed work of the dominant kernel: per-block counters from this round's ncu capture of one full-wave launch
        cnt = None
        if os.path.exists(COUNTERS):
            try:
                cnt = json.load(open(COUNTERS)).get(args.precision)
            except Exception:
                cnt = None
        ref_flops = algorithmic_flops(nblk - n_clips, n_clips, r_dev["maskers"], L)
        roof = {"kernel": "analysis_kernel (MDCT + psychoacoustics, fused)", "bound": "fp64" if is64 else "fp32",
                "peak": peak_tf, "unit": "TFLOP/s", "peak_source": "mrc_measure_peaks (FMA chain on this GPU, this run)",
                "launches_per_step": nwaves, "avg_launch_ms": an_ms / nwaves, "ms_per_step": an_ms}
        if cnt:
            flop_key = "fp64_flop_per_block" if is64 else "fp32_flop_per_block"
            inst = cnt["fp64_thread_inst_per_block" if is64 else "fp32_thread_inst_per_block"]
            slots = sum(inst.values())                     # an ADD or a MUL takes the pipe slot an FMA would
            ex_flops = cnt[flop_key] * nblk
            ach_tf = ex_flops / (an_ms * 1e-3) / 1e12
            roof.update({
                "achieved": ach_tf, "frac": ach_tf / peak_tf if peak_tf else None,
                "frac_pipe_slots": (slots * nblk / (an_ms * 1e-3)) / (peak_tf * 1e12 / 2.0) if peak_tf else None,
                "traffic": cnt["dram_bytes_per_block"] * nblk / nwaves,
                "work": "FLOPs the kernel executes: %.0f per block (2 x FMA + ADD + MUL thread instructions of one "
                        "%d-block launch, ncu, %s) x %d blocks / the kernel's launch durations measured live inside the step "
                        "(where it shares the SMs with the cost / reservoir-map / pack kernels of the neighbouring waves); "
                        "frac_pipe_slots counts FMA, ADD and MUL as one pipe slot each" %
                        (cnt[flop_key], cnt["launch_blocks"], cnt.get("source", "profiles/"), nblk),
                "ncu_pipe_active_pct": cnt.get("pipe_fp64_active_pct" if is64 else "pipe_fma_active_pct"),
                "ncu_issue_active_pct": cnt.get("issue_active_pct"),
                "ncu_launch_ms_alone": cnt.get("launch_ms_under_ncu"),
                "work_reduction_vs_reference": ref_flops / ex_flops,
                "reference_formulation_flops_per_step": ref_flops})
        else:
            roof.update({"achieved": None, "frac": None, "traffic": None,
                         "work": "profiles/analysis_counters.json missing: no executed-flop count for this build"})
        alg_bytes = ALGORITHMIC_BYTES_PER_BLOCK * nblk
        line = {
            "metric": METRIC, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * r_dev["wall"] / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64" if is64 else "f32",
            "data": "synthetic",
            "config": {"workload": ("ONE %.0f s synthetic 48 kHz stereo 16-bit stream (BASELINE configs[1] = 1 h) sharded by "
                                    "block range over %d GPUs: N/2-sample halo read from the PCM, reservoir handed rank to "
                                    "rank (one int32 per boundary, NCCL send/recv), joint M/S, 128 kb/s/ch, %s mode; "
                                    "per-rank figures below are rank 0's" % (seconds, world, args.precision)) if strong else
                       ("%.0f s synthetic 48 kHz stereo 16-bit stream per GPU (BASELINE configs[1] = 1 h), "
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
                                  "note": "per-kernel sums; analysis+cost(+reservoir maps) of wave w+1 overlap chain+pack of wave w"},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(frames * 4),
                    "d2h_bytes_per_step": int(r_e2e["nbytes"]), "ms_per_step": 1000.0 * r_e2e["wall"] / args.steps},
            "gpu_launches": int(r_dev["launches"]),
            "clocks": r_dev["clocks"],
            "roofline": roof,
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (1e-3 * 1000.0 * r_dev["dev"] / args.steps) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (1e-3 * 1000.0 * r_dev["dev"] / args.steps) / 1e9 / hbm_peak,
                             "peak_source": hbm_src,
                             "work": "algorithmic bytes (%d per block: PCM in + bitstream + prefixes) / whole step" %
                                     ALGORITHMIC_BYTES_PER_BLOCK},
            "pipe_peaks": peaks,
            "executed_work": r_dev["work"],
        }
        if cnt and cnt.get("wave"):
            wv = cnt["wave"]
            line["pipeline_traffic"] = {"dram_bytes_per_block": wv["dram_bytes_per_block"],
                                        "algorithmic_bytes_per_block": ALGORITHMIC_BYTES_PER_BLOCK,
                                        "ratio": wv["dram_bytes_per_block"] / ALGORITHMIC_BYTES_PER_BLOCK,
                                        "gbs_at_this_step_rate": wv["dram_bytes_per_block"] * nblk /
                                        (1e-3 * 1000.0 * r_dev["dev"] / args.steps) / 1e9,
                                        "source": "ncu dram__bytes of every kernel of one %d-block wave (%s)" %
                                                  (wv["blocks"], wv.get("source") or cnt.get("source", "profiles/"))}
        if args.block_switching:
            line["config"]["workload"] += (", BLOCK SWITCHING on (transient detector + look-ahead, short blocks of 128: "
                                           "%d blocks written for %d blocks of 1024 frames)" %
                                           (r_dev["extra"]["blocks_written"], nblk))
            line["stage_ms_per_step"]["transient_detector"] = r_dev["extra"]["transient_ms"]
        solo = world == 1                # the side measurements below belong to the N = 1 line (cpu_baseline: "rank 0 at N=1 only")
        if args.spreading == "factorised" and not args.no_sequential_sample and not args.block_switching and solo:
            # the same analysis kernel summing the maskers pair by pair in the reference's order, on a bounded
            # sample of the same stream: the kernel SURVEY 8d's 40-flop-per-pair work formula describes (here the formula
            # IS roughly what runs, so the reference-formulation flop count is used)
            sample_s = min(seconds, 600.0)
            nfr = int(sample_s * SR)
            cs = Codec(device=local, precision=args.precision, spreading="sequential")
            soff = np.array([0, nfr], dtype=np.int64)
            for _ in range(2):
                cs.encode_batch_device(d_pcm.data_ptr(), soff, d_out.data_ptr(), cap)
            ts = cs.last_timing()
            sflops = algorithmic_flops(cs.n_blocks(nfr) - 1, 1, ts["maskers"], L)
            s_tf = sflops / (ts["analysis_ms"] * 1e-3) / 1e12
            line["roofline_sequential"] = {"kernel": "analysis_kernel, MRC_FLAG_SPREAD_SEQUENTIAL", "bound": roof["bound"],
                                           "achieved": s_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": s_tf / peak_tf,
                                           "work": "SURVEY 8d formula, 40 flops per (masker, line) pair",
                                           "sample": "first %.0f s of the stream, 1 launch" % sample_s,
                                           "avg_launch_ms": ts["analysis_ms"]}
            cs.close()
        if not args.no_decode and solo:
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
                              "kernel_ms": td["decode_ms"], "host_chain_walk_ms": td["chain_ms"],
                              "h2d_bytes_per_step": nbytes,
                              "d2h_bytes_per_step": int(pcm_out.nbytes),
                              "note": "decode of this rank's stream through Codec.decode_batch, host buffers, 3 steps"}
            del h_dec
        if not args.no_music and args.workload == "stream" and not args.block_switching and solo:
            # the same encode on dense-masker material (synth_music: chords of harmonic notes, most of the spectrum above
            # 40 dB SPL): the band-maximum search prunes far less there, so this is the unfriendly end of the input range
            ms_ = min(args.music_seconds, seconds)
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as ex:
                parts = list(ex.map(lambda i: synth.synth_music(100 + i, 30.0), range(max(1, int(ms_ // 30)))))
            mus = np.concatenate(parts, axis=0)
            moff = np.array([0, mus.shape[0]], dtype=np.int64)
            d_mus = torch.from_numpy(mus).to(dev)
            for _ in range(2):
                codec.encode_batch_device(d_mus.data_ptr(), moff, d_out.data_ptr(), cap)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                codec.encode_batch_device(d_mus.data_ptr(), moff, d_out.data_ptr(), cap)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            tm = codec.last_timing()
            line["music"] = {"value": (mus.shape[0] / SR) / dt, "unit": "audio-s/s", "seconds": mus.shape[0] / SR,
                             "ms_per_step": 1000.0 * dt, "analysis_ms": tm["analysis_ms"],
                             "maskers_per_block": tm["maskers"] / max(tm["blocks"], 1),
                             "general_pairs_per_block": tm["general_pairs"] / max(tm["blocks"], 1),
                             "note": "device-resident encode of synth_music material (dense loud maskers), same codec"}
            del d_mus
        if not args.no_cpu_baseline and solo:
            line["cpu_baseline"] = cpu_baseline(pcm, args.cpu_sample_seconds)
        print(json.dumps(line))
    codec.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
