"""cProfile of HBM-resident TF-IDF steps (development aid): python tools/profile_step.py [gb]"""
import cProfile
import io
import os
import pstats
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dampr_b200 import Dampr, synth
from dampr_b200 import runner as runner_mod
from dampr_b200.plan import DeviceText

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
ctx = runner_mod.get_ctx()
V = 1_000_000
vocab, cdf = synth.make_vocab(V), synth.make_cdf(V)
tb = ctx.synth_text(1234, int(gb * 1e9 / 99.94), vocab[0], vocab[1], cdf)
out_root = tempfile.mkdtemp(prefix="dampr_prof_")


def step(i):
    bench.tfidf_job(Dampr, DeviceText(tb), os.path.join(out_root, "o%d" % i))


for i in range(3):
    step(i)
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for i in range(3, 8):
    step(i)
pr.disable()
print("5 steps: %.1f ms/step" % ((time.perf_counter() - t0) * 200))
print([(s[:40], round(ms, 1)) for s, ms in runner_mod.LAST_STATS.ms])
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print(s.getvalue()[:6000])
shutil.rmtree(out_root, ignore_errors=True)
