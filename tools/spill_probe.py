"""Where the time of the spill path goes (development aid): config-4-shaped run at 1/5 scale.
   python tools/spill_probe.py [million_records] [arena_GB]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dampr_b200 import Dampr, settings
from dampr_b200 import runner as runner_mod
from dampr_b200.inputs import ArrayKVInput


def main():
    n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 625_000_000
    settings.device_arena_bytes = int(float(sys.argv[2]) * (1 << 30)) if len(sys.argv) > 2 else (3276 << 20)
    ctx = runner_mod.get_ctx()
    if os.environ.get("DAMPR_HOST_THREADS"):
        from dampr_b200 import device as dev
        dev.set_option("host_threads", int(os.environ["DAMPR_HOST_THREADS"]))
    kv = ctx.synth_kv(7, n, n // 4)
    keys, vals = kv.columns()
    kv.free()
    vals = vals.view(np.int64)
    total = int(vals.sum())
    for rep in range(3):
        t0 = time.time()
        res = Dampr.read_input(ArrayKVInput(keys, vals)).group_by(lambda x: x[0], lambda x: x[1]).reduce(lambda k, it: sum(it)).run()
        wall = time.time() - t0
        fr = res.datasets
        ok = int(np.asarray(fr.cols[1]).sum()) == total
        print(json.dumps({"records": n, "wall_s": round(wall, 2), "MB_per_s": round(16 * n / wall / 1e6, 1), "ok": ok,
                          "spill": getattr(runner_mod.LAST_STATS, "spill", None),
                          "stage_ms": [round(ms) for _s, ms in runner_mod.LAST_STATS.ms]}), flush=True)
        del res, fr


if __name__ == "__main__":
    main()
