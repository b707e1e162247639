"""Where the wall time of a host-resident kv job goes (development aid).
python tools/kv_e2e.py [million_records] [n_keys]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from dampr_b200 import Dampr
from dampr_b200 import device as dev
from dampr_b200 import runner as runner_mod
from dampr_b200.inputs import ArrayKVInput

n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 200_000_000
K = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
rng = np.random.default_rng(42)
with np.errstate(over="ignore"):
    keys = rng.integers(0, K, size=n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)
vals = rng.integers(-1000, 1000, size=n)
ctx = runner_mod.get_ctx()


def t(label, f, reps=3):
    best = None
    for _ in range(reps):
        t0 = time.time()
        r = f()
        ctx.sync()
        dt = time.time() - t0
        best = dt if best is None else min(best, dt)
    print("%-44s %8.1f ms  (%.1f GB/s on 16 B/record)" % (label, best * 1e3, 16 * n / best / 1e9))
    return r


def up():
    kv = ctx.kv_from_columns(keys, vals)
    kv.free()


t("upload columns (pageable numpy)", up)
kv = ctx.kv_from_columns(keys, vals)
t("download columns (pageable numpy)", lambda: kv.columns())
recs = np.empty((n, 2), dtype=np.uint64)
recs[:, 0] = keys
recs[:, 1] = vals.view(np.uint64)
kv.free()


def upr():
    k2 = ctx.kv_from_records(recs)
    k2.free()


t("upload records (pageable numpy)", upr)


def job():
    res = Dampr.read_input(ArrayKVInput(keys, vals)).a_group_by(lambda x: x[0], lambda x: x[1]).sum().run()
    return res


res = t("DSL a_group_by(k).sum() end to end", job)
print([(s.split("`")[1][:30] if "`" in s else s[:40], how[:60], round(ms, 1))
       for (s, how, _d), (_s2, ms) in zip(runner_mod.LAST_STATS.stages, runner_mod.LAST_STATS.ms)])


def job2():
    return Dampr.read_input(ArrayKVInput(keys, vals)).group_by(lambda x: x[0], lambda x: x[1]) \
        .reduce(lambda k, it: sum(it)).run()


t("DSL group_by(k).reduce(sum) end to end", job2)


def job3():
    return Dampr.read_input(ArrayKVInput(keys, vals)).sort_by(lambda x: x[0]).run()


t("DSL sort_by(k) end to end", job3, reps=2)
print([(s.split("`")[1][:30] if "`" in s else s[:40], how[:60], round(ms, 1))
       for (s, how, _d), (_s2, ms) in zip(runner_mod.LAST_STATS.stages, runner_mod.LAST_STATS.ms)])
