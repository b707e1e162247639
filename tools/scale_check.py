"""Full-size runs of the other BASELINE configs with size-independent checks (development aid; the
numbers land in profiles/).  python tools/scale_check.py [config2] [config4] [config5]

config2  a_group_by(k).sum() on 10 GB of 16-byte records (6.25e8), K = 1e7, through the DSL
         (host columns -> device partition + sort + segmented reduce -> host frame)
config4  group_by external sort on 50 GB of records with the device arena capped at 16 GB (spill path)
config5  joins at 20 GB x 2 GB: device sort of both sides + merge-join ranges, and the broadcast hash
         build + probe, on device-resident synthetic records (C-ABI level)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from dampr_b200 import Dampr, settings
from dampr_b200 import device as dev
from dampr_b200 import runner as runner_mod
from dampr_b200.inputs import ArrayKVInput

GOLD = np.uint64(0x9E3779B97F4A7C15)


def host_kv(n, n_keys, seed=42):
    """Same distribution as gen.kv, built in blocks (numpy only)."""
    from dampr_b200 import synth  # noqa: F401
    keys = np.empty(n, dtype=np.uint64)
    vals = np.empty(n, dtype=np.int64)
    rng = np.random.default_rng(seed)
    B = 1 << 26
    for lo in range(0, n, B):
        hi = min(n, lo + B)
        with np.errstate(over="ignore"):
            keys[lo:hi] = rng.integers(0, n_keys, size=hi - lo, dtype=np.uint64) * GOLD
        vals[lo:hi] = rng.integers(-1000, 1000, size=hi - lo)
    return keys, vals


def stage_summary():
    return [(s.split("`")[1][:30] if "`" in s else s, how, round(ms, 1))
            for (s, how, _d), (_s2, ms) in zip(runner_mod.LAST_STATS.stages, runner_mod.LAST_STATS.ms)]


def config2(out):
    n = 625_000_000
    t0 = time.time()
    keys, vals = host_kv(n, 10_000_000)
    tgen = time.time() - t0
    total = int(vals.sum())
    ctx = runner_mod.get_ctx()
    t0 = time.time()
    res = Dampr.read_input(ArrayKVInput(keys, vals)).a_group_by(lambda x: x[0], lambda x: x[1]).sum().run()
    cold = time.time() - t0    # first run: device pool and staging ring are allocated here
    del res
    ctx.timings_reset()
    t0 = time.time()
    res = Dampr.read_input(ArrayKVInput(keys, vals)).a_group_by(lambda x: x[0], lambda x: x[1]).sum().run()
    wall = time.time() - t0
    fr = res.datasets
    rk, rv = fr.cols[0], fr.cols[1]
    assert int(np.asarray(rv).sum()) == total, "sum of group sums != sum of values"
    assert len(np.unique(rk)) == len(rk), "a key appears in two groups"
    kt = {}
    for name, ms in ctx.timings():
        kt[name] = kt.get(name, 0.0) + ms
    dev_ms = sum(v for k, v in kt.items() if k in ("part_hist", "part_scatter", "leaf_sort", "seg_reduce", "misc"))
    out["config2"] = {"records": n, "groups": int(len(rk)), "wall_s": round(wall, 3), "cold_wall_s": round(cold, 3),
                      "gen_s": round(tgen, 1),
                      "MB_per_s_end_to_end": round(16 * n / wall / 1e6, 1), "kernel_ms": {k: round(v, 2) for k, v in kt.items()},
                      "device_partition_sort_reduce_ms": round(dev_ms, 2),
                      "roofline_16N_plus_16G_GBps": round((16 * n + 16 * len(rk)) / dev_ms / 1e6, 1),
                      "stages": stage_summary()}
    del keys, vals


def config4(out):
    n = 3_125_000_000
    settings.device_arena_bytes = 16 << 30
    if os.environ.get("DAMPR_HOST_THREADS"):
        dev.set_option("host_threads", int(os.environ["DAMPR_HOST_THREADS"]))
    try:
        t0 = time.time()
        runner_mod.get_ctx()   # the CUDA context exists before the job (a long-running engine's state)
        tctx = time.time() - t0
        t0 = time.time()
        keys, vals = host_kv(n, n // 4, seed=7)
        tgen = time.time() - t0
        total = int(vals.sum())
        from dampr_b200 import plan
        t0 = time.time()
        plan._may_overflow(vals)
        t_ovf = time.time() - t0
        runs = []
        for label in ("cold", "second job (host run buffer kept; its page-locking finishes in the background)",
                      "third job (host run buffer kept and page-locked)"):
            t0 = time.time()
            res = Dampr.read_input(ArrayKVInput(keys, vals)).group_by(lambda x: x[0], lambda x: x[1]) \
                .reduce(lambda k, it: sum(it)).run()
            wall = time.time() - t0
            fr = res.datasets
            rk, rv = fr.cols[0], fr.cols[1]
            assert int(np.asarray(rv).sum()) == total
            # key uniqueness without a host sort of ~8e8 keys (np.unique took ten minutes here): every key is
            # k * GOLD for some k < n/4, so k = key * GOLD^-1 must hit every slot at most once
            inv_gold = np.uint64(pow(0x9E3779B97F4A7C15, -1, 1 << 64))
            with np.errstate(over="ignore"):
                kk = np.asarray(rk).view(np.uint64) * inv_gold
            assert int(kk.max()) < n // 4
            seen = np.zeros(n // 4, dtype=np.bool_)
            seen[kk] = True
            assert int(seen.sum()) == len(rk)
            runs.append({"run": label, "wall_s": round(wall, 2), "MB_per_s_end_to_end": round(16 * n / wall / 1e6, 1),
                         "groups": int(len(rk)), "spill": getattr(runner_mod.LAST_STATS, "spill", None),
                         "stages": stage_summary()})
            del res, fr, rk, rv, kk, seen
        out["config4"] = {"records": n, "arena_bytes": 16 << 30, "gen_s": round(tgen, 1), "ctx_create_s": round(tctx, 2),
                          "overflow_check_s_included_in_wall": round(t_ovf, 2),
                          "wall_s": runs[0]["wall_s"], "MB_per_s_end_to_end": runs[0]["MB_per_s_end_to_end"],
                          "runs": runs}
    finally:
        settings.device_arena_bytes = None


def config5(out):
    ctx = runner_mod.get_ctx()
    nl, nr = 1_250_000_000, 125_000_000
    L = ctx.synth_kv(1, nl, 2 * nr)       # ~50 % of the left keys have a partner
    R = ctx.synth_kv(2, nr, 4 * nr)
    ctx.sync()
    # broadcast hash build + probe (build side must be unique: reduce it first)
    ctx.timings_reset()
    t0 = time.time()
    Ru = R.sort_reduce(dev.OP_FIRST, dev.KEY_MIX)
    nbuild = len(Ru)
    vals, hit = Ru.hash_probe(L)
    ctx.sync()
    wall_probe = time.time() - t0
    kt = {}
    for name, ms in ctx.timings():
        kt[name] = kt.get(name, 0.0) + ms
    hits = int(hit.sum())
    vals.free()
    # merge join: sort both sides, ranges per left key
    ctx.timings_reset()
    t0 = time.time()
    L.sort(dev.KEY_MIX)
    Ru.sort(dev.KEY_MIX)
    rows = L.join_ranges(Ru, dev.KEY_MIX)
    ctx.sync()
    wall_join = time.time() - t0
    kt2 = {}
    for name, ms in ctx.timings():
        kt2[name] = kt2.get(name, 0.0) + ms
    matched_rows = rows[rows[:, 3] > rows[:, 2]]
    matched_records = int((matched_rows[:, 1] - matched_rows[:, 0]).sum())
    assert matched_records == hits, (matched_records, hits)  # both joins agree on the matching left records
    out["config5"] = {"left": nl, "right": nr, "build_unique": nbuild, "left_records_with_partner": hits,
                      "probe_wall_s": round(wall_probe, 3), "probe_kernel_ms": {k: round(v, 2) for k, v in kt.items()},
                      "merge_join_wall_s": round(wall_join, 3), "merge_join_kernel_ms": {k: round(v, 2) for k, v in kt2.items()},
                      "left_groups": int(len(rows))}
    L.free()
    R.free()
    Ru.free()


if __name__ == "__main__":
    which = sys.argv[1:] or ["config2", "config4", "config5"]
    out = {}
    for w in which:
        t0 = time.time()
        {"config2": config2, "config4": config4, "config5": config5}[w](out)
        print(w, "done in %.1fs" % (time.time() - t0), json.dumps(out[w])[:1500], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/scale_check.json", "w") as f:
        json.dump(out, f, indent=1)
