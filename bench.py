#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: MB/s ingested end to end, TF-IDF (benchmarks/tf-idf-dampr.py)
on synthetic Zipf text, through the Dampr DSL on the B200 engine.

    python bench.py --gpus 1 --steps K --warmup W            # this engine
    python bench.py --impl reference --gpus 1 ...             # CPU port of the reference's runner

A step = one complete run of the TF-IDF graph (len + tokenise/count + fold + cross + sink_tsv) over
the whole corpus (10 GB, V = 1e6, split evenly over the ranks: strong scaling).
  value  whole-job MB/s with the text resident in HBM when the step starts (max over ranks).
  e2e    the same job from page-locked host memory: host->device copies of the text and the
         device->host fetch of the result table are inside the timed region.
  roofline  the tokenise+combine kernel (dominant): algorithmic bytes = 1 B per input byte
         (SURVEY §8(d)), CUDA-event duration measured inside the timed steps, peak = measured HBM
         copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  oracle/cpu_runner.py (multi-process Python port of the reference's runner) on a
         bounded prefix of the same corpus, all host cores.
"""
import argparse
import json
import math
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

RX = re.compile(r"[^\w]+")
MB = 1e6


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gb", type=float, default=float(os.environ.get("DAMPR_BENCH_GB", "10")),
                    help="corpus size in GB (10 = the BASELINE config)")
    ap.add_argument("--vocab", type=int, default=1000000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        threading.Thread.__init__(self)
        self.daemon = True
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        """In-process NVML sampling (spawning nvidia-smi every 200 ms stalls CUDA calls of this process
        for tens of ms); nvidia-smi is the fallback when pynvml is missing."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            R = pynvml
            bits = [("hw_slowdown", R.nvmlClocksThrottleReasonHwSlowdown),
                    ("hw_thermal_slowdown", R.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", R.nvmlClocksThrottleReasonSwThermalSlowdown),
                    ("sw_power_cap", R.nvmlClocksThrottleReasonSwPowerCap)]
            while not self.stop_flag.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                rs = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                row = [str(self.gpu), str(sm), str(mx), "", hex(rs)]
                row += ["Active" if rs & b else "Not Active" for _n, b in bits]
                self.samples.append(row)
                # NVML queries and CUDA calls share a driver lock: sample sparsely (2 Hz) so the
                # sampler itself does not stall the allocation calls of the timed steps
                self.stop_flag.wait(0.5)
            return
        except Exception:
            pass
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [float(s[1]) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for nm, v in zip(names, s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def tfidf_job(Dampr, source_dataset, out_dir):
    """benchmarks/tf-idf-dampr.py:9-21 verbatim, on a given input dataset."""
    docs = Dampr.read_input(source_dataset)
    doc_freq = docs.flat_map(lambda x: set(RX.split(x.lower()))).count(reduce_buffer=float("inf"))
    idf = doc_freq.cross_right(docs.len(),
                               lambda df, total: (df[0], df[1], math.log(1 + (float(total) / df[1]))),
                               memory=True)
    return idf.sink_tsv(out_dir).run()


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def cpu_baseline(sample_bytes_arr, n_procs):
    """Time the oracle CPU port on a prefix of the corpus (rank 0 only)."""
    from oracle import cpu_runner
    tmp = tempfile.mkdtemp(prefix="dampr_cpu_")
    path = os.path.join(tmp, "sample.txt")
    try:
        with open(path, "wb") as f:
            f.write(sample_bytes_arr.tobytes())
        with open(path, "rb") as f:  # warm the page cache like benchmarks/run.sh:3
            while f.read(1 << 24):
                pass
        sec, n_terms, n_lines = cpu_runner.timed_tfidf(path, os.path.join(tmp, "idfs"), n_procs)
        return sec, n_terms, n_lines
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cut_at_line(arr, nbytes):
    """Largest prefix <= nbytes that ends at a newline."""
    nbytes = min(nbytes, len(arr))
    if nbytes == len(arr):
        return arr
    tail = arr[max(0, nbytes - 65536):nbytes]
    idx = np.flatnonzero(tail == 10)
    cut = nbytes - (len(tail) - 1 - int(idx[-1])) if len(idx) else nbytes
    return arr[:cut]


def main():
    args = parse_args()
    rank, world, local = dist_env()
    if world != max(1, args.gpus) and world > 1:
        args.gpus = world
    total_bytes = int(args.gb * 1e9)
    shard_bytes = total_bytes // world
    n_lines = int(shard_bytes / 99.94)  # mean synthetic line: 99.94 B at V = 1e6 (measured)

    from dampr_b200 import device as dev
    from dampr_b200 import synth, settings

    if args.impl == "reference" and rank != 0:
        return 0

    settings.device = local
    ctx_err = None
    try:
        ctx = dev.Ctx(local)
    except dev.DeviceError as e:
        ctx, ctx_err = None, e
    if ctx is None:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback): %s" % ctx_err)

    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- synthetic corpus: generated on the device, mirrored into page-locked host memory -------------
    vocab = synth.make_vocab(args.vocab)
    cdf = synth.make_cdf(args.vocab)
    tb = ctx.synth_text(1234 + rank, n_lines, vocab[0], vocab[1], cdf)
    nbytes = tb.n
    pin = dev.PinnedBuffer(nbytes)
    ctx.check(ctx.lib.dampr_textbuf_download(ctx.h, tb.h, 0, pin.array.ctypes.data, nbytes))
    host_text = pin.array[:nbytes]

    ncores = os.cpu_count() or 1
    workload = "benchmarks/tf-idf-dampr.py on %.2f GB synthetic Zipf(1.1) text, V=%d, %d GPU(s)" % (
        nbytes * world / 1e9, args.vocab, world)

    if args.impl == "reference":
        # CPU port of the reference's runner, bounded sample per step (a few seconds of CPU work)
        per_core = 6e6
        sample = cut_at_line(host_text, int(min(nbytes, per_core * ncores * 3)))
        tb.free()
        times = []
        for i in range(args.warmup + args.steps):
            sec, n_terms, n_lines_s = cpu_baseline(sample, ncores)
            if i >= args.warmup:
                times.append(sec)
        val = len(sample) * len(times) / sum(times) / MB
        line = {"impl": "reference", "metric": "MB/s ingested end-to-end, TF-IDF 10 GB synthetic text, 1/2/4/8 GPU", "value": val,
                "unit": "MB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": workload, "sample": "%d-byte line-aligned prefix per step" % len(sample)},
                "cpu_baseline": {"value": val, "unit": "MB/s", "cores": ncores, "kind": "port",
                                 "sample": "%d-byte prefix of the corpus, %d processes (oracle/cpu_runner.py)" % (
                                     len(sample), ncores)},
                "e2e": {"value": val, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    from dampr_b200 import Dampr
    from dampr_b200 import runner as runner_mod
    from dampr_b200.plan import MemoryText, DeviceText

    out_root = tempfile.mkdtemp(prefix="dampr_bench_r%d_" % rank)

    def barrier():
        ctx.sync()
        if use_dist:
            dist.barrier()

    def step(resident, i):
        out_dir = os.path.join(out_root, "idfs_%s_%d" % ("dev" if resident else "host", i))
        src = DeviceText(tb) if resident else MemoryText(host_text)
        tfidf_job(Dampr, src, out_dir)
        return out_dir

    def max_over_ranks(x):
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up immediately before each timed loop (no host-side pause in between) -------------------
    sampler = ClockSampler(local)
    if not os.environ.get("DAMPR_BENCH_NOSAMPLER"):
        sampler.start()
    verbose = bool(os.environ.get("DAMPR_BENCH_VERBOSE"))
    for i in range(args.warmup):
        ts = time.perf_counter()
        shutil.rmtree(step(True, -1 - i), ignore_errors=True)
        if verbose:
            print("warmup resident %d: %.1f ms %s" % (i, 1e3 * (time.perf_counter() - ts),
                                                      [round(ms, 1) for _s, ms in runner_mod.LAST_STATS.ms]), file=sys.stderr)

    # ---- timed: device-resident input ("value") ------------------------------------------------------------
    runner_ctx = runner_mod.get_ctx(local)
    barrier()
    runner_ctx.timings_reset()
    l0 = runner_ctx.launches()
    t0 = time.perf_counter()
    dev_steps = []
    for i in range(args.steps):
        ts = time.perf_counter()
        d = step(True, i)
        dev_steps.append([round(1e3 * (time.perf_counter() - ts), 1)] +
                         [round(ms, 1) for _s, ms in runner_mod.LAST_STATS.ms])
    barrier()
    t_dev = max_over_ranks(time.perf_counter() - t0)
    launches = runner_ctx.launches() - l0
    ktimes = runner_ctx.timings()
    for i in range(args.warmup):
        ts = time.perf_counter()
        shutil.rmtree(step(False, -1 - i), ignore_errors=True)
        if verbose:
            print("warmup e2e %d: %.1f ms %s" % (i, 1e3 * (time.perf_counter() - ts),
                                                 [round(ms, 1) for _s, ms in runner_mod.LAST_STATS.ms]), file=sys.stderr)

    # ---- timed: host-resident input ("e2e") ----------------------------------------------------------------
    barrier()
    t0 = time.perf_counter()
    e2e_steps = []
    for i in range(args.steps):
        ts = time.perf_counter()
        d = step(False, i)
        e2e_steps.append([round(1e3 * (time.perf_counter() - ts), 1)] +
                         [round(ms, 1) for _s, ms in runner_mod.LAST_STATS.ms])
    barrier()
    t_e2e = max_over_ranks(time.perf_counter() - t0)
    stats = runner_mod.LAST_STATS.stages if runner_mod.LAST_STATS else []
    sampler.stop_flag.set()
    if sampler.is_alive():
        sampler.join(timeout=5)
    if os.environ.get("DAMPR_BENCH_PROFILE") and rank == 0:
        # development aid: where the host time of a device-resident step goes (after the timed loops)
        import cProfile
        import io
        import pstats
        pr = cProfile.Profile()
        pr.enable()
        for i in range(5):
            step(True, 100 + i)
        pr.disable()
        for key in ("cumulative", "tottime"):
            sio = io.StringIO()
            pstats.Stats(pr, stream=sio).sort_stats(key).print_stats(35)
            print(sio.getvalue()[:7000], file=sys.stderr)
    elif os.environ.get("DAMPR_BENCH_PROFILE"):
        for i in range(5):
            step(True, 100 + i)

    # result size fetched from the device per step
    n_terms = 0
    try:
        for fn in os.listdir(d):
            with open(os.path.join(d, fn)) as f:
                n_terms += sum(1 for _ in f)
    except Exception:
        pass

    total = nbytes * world
    value = total * args.steps / t_dev / MB
    e2e = total * args.steps / t_e2e / MB
    tc = [ms for name, ms in ktimes if name == "text_count"]
    peak, peak_kind = measured_peak()
    roof = None
    if tc:
        per_launch_ms = sum(tc) / len(tc)
        per_launch_bytes = nbytes * args.steps / len(tc)
        ach = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9
        traffic, traffic_src, issue = None, None, None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_text_traffic.json")) as f:
                tj = json.load(f)
            # dram__bytes_read+write of one ncu --set full capture of this kernel, scaled from the
            # captured launch (250 MB of the same corpus) to this launch's bytes
            traffic = tj["dram_bytes_per_input_byte"] * per_launch_bytes
            traffic_src = "ncu capture scaled per input byte: " + tj["source"]
            # what actually bounds this kernel: instruction issue. smsp__inst_executed of the same capture
            # per input byte against 4 schedulers x SMs x SM clock (one warp instruction per cycle each)
            wipb = tj.get("warp_instructions_per_input_byte")
            clk = (sampler.summary().get("sm_mhz") or 0) * 1e6
            if wipb and clk:
                issue_peak = 4 * ctx.num_sms() * clk / wipb / 1e9
                issue = {"bound": "issue", "warp_instructions_per_byte": wipb, "peak": issue_peak, "unit": "GB/s",
                         "frac": ach / issue_peak}
        except Exception:
            pass
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "text_count2_kernel<NONWORD_LOWER_SET> (csrc/text2.cu)",
                "algorithmic_bytes_per_launch": per_launch_bytes, "ms_per_launch": per_launch_ms,
                "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)", "issue_roofline": issue}

    line = {"metric": "MB/s ingested end-to-end, TF-IDF 10 GB synthetic text, 1/2/4/8 GPU", "value": value, "unit": "MB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_dev / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "bytes_per_rank": nbytes, "lines_per_rank": n_lines,
                       "l2": "inputs (%.1f GB per rank) are larger than L2" % (nbytes / 1e9),
                       "resident_step_ms_total_then_stages": dev_steps,
                       "stages": [[s.split("`")[1][:40] if "`" in s else s, h] for s, h, _d in stats],
                       "e2e_stage_ms": [round(ms, 2) for _s, ms in (runner_mod.LAST_STATS.ms if runner_mod.LAST_STATS else [])]},
            "e2e": {"value": e2e, "unit": "MB/s", "h2d_bytes_per_step": nbytes * world,
                    # result table per term: 16-byte key string + 8-byte count (single GPU; the synthetic
                    # corpus has no hashed tokens), plus the 8-byte code after the exchange (multi GPU)
                    "d2h_bytes_per_step": int(n_terms) * (24 if world == 1 else 32),
                    "ms_per_step": 1e3 * t_e2e / args.steps,
                    "step_ms": e2e_steps},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roof}

    if rank == 0 and not args.no_cpu_baseline:
        sample = cut_at_line(host_text, int(min(nbytes, 6e6 * ncores * 3)))
        cpu_baseline(sample[:min(len(sample), 4 << 20)], ncores)  # warm-up (imports, fork)
        sec, _t, _l = cpu_baseline(sample, ncores)
        line["cpu_baseline"] = {"value": len(sample) / sec / MB, "unit": "MB/s", "cores": ncores, "kind": "port",
                                "sample": "%d-byte line-aligned prefix of the corpus, %d processes "
                                          "(oracle/cpu_runner.py, a leaner port of the reference's runner)" % (
                                              len(sample), ncores)}
    shutil.rmtree(out_root, ignore_errors=True)
    if rank == 0:
        print(json.dumps(line))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
