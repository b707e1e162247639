"""Sweep of the planned leaf segment size (development aid): python tools/kv_sweep.py [million_records]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kv_bench
from dampr_b200 import device as dev

n = int(float(sys.argv[1]) * 1e6) if len(sys.argv) > 1 else 100_000_000
ctx = dev.Ctx(0)
for target in (2600, 1300, 650, 325, 3900):
    dev.set_option("kv_leaf_target", target)
    for nk, label in ((n, "K=N"), (10_000_000, "K=1e7")):
        r = kv_bench.sort_case(ctx, n, nk, label="%s leaf_target=%d" % (label, target))
        print(json.dumps({k: r[k] for k in ("label", "ms", "kernels_ms", "frac_of_hbm_peak", "ok")}), flush=True)
dev.set_option("kv_leaf_target", 2600)
ctx.close()
