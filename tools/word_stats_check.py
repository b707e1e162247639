"""examples/word-stats.py at scale on synthetic text resident in HBM, with size-independent checks
(development aid).  python tools/word_stats_check.py [gb]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from dampr_b200 import Dampr, synth
from dampr_b200 import runner as runner_mod
from dampr_b200.plan import DeviceText

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
ctx = runner_mod.get_ctx()
V = 1_000_000
vocab, cdf = synth.make_vocab(V), synth.make_cdf(V)
tb = ctx.synth_text(1234, int(gb * 1e9 / 99.94), vocab[0], vocab[1], cdf)


def job():
    words = Dampr.read_input(DeviceText(tb)).flat_map(lambda line: line.split())
    top_words = words.count(lambda x: x).sort_by(lambda wc: -wc[1])
    total_count = top_words.fold_by(key=lambda word: 1, value=lambda x: x[1], binop=lambda x, y: x + y)
    word_lengths = top_words.fold_by(lambda tc: len(tc[0]), value=lambda tc: tc[1], binop=lambda x, y: x + y) \
        .sort_by(lambda cl: cl[0])
    avg = word_lengths.map(lambda wl: wl[0] * wl[1]).a_group_by(lambda x: 1).sum() \
        .join(total_count).reduce(lambda awl, tc: next(awl)[1] / float(next(tc)[1]))
    return Dampr.run(total_count, top_words, word_lengths, avg, name="word-stats")


for rep in range(3):
    t0 = time.time()
    tc, tw, wl, awl = job()
    wall = time.time() - t0
    stats = runner_mod.LAST_STATS
    print("run %d: %.1f ms (%.1f GB/s)" % (rep, wall * 1e3, tb.n / wall / 1e9))
if os.environ.get("PROFILE"):
    import cProfile, pstats, io
    pr = cProfile.Profile()
    pr.enable()
    job()
    pr.disable()
    sio = io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats("tottime").print_stats(25)
    print(sio.getvalue()[:5000])
for (s, how, _d), (_s2, ms) in zip(stats.stages, stats.ms):
    print("  %-44s %-70s %8.1f ms" % ((s.split("`")[1] if "`" in s else s)[:44], how[:70], ms))
total = tc.read(1)[0][1]
top = tw.read(5)
hist = dict(wl.read())
avgv = awl.read(1)[0][1]
counts = [c for _w, c in tw.read()]
assert counts == sorted(counts, reverse=True), "top_words not sorted by -count"
assert sum(counts) == total, "sum of word counts != total"
assert sum(hist.values()) == total, "length histogram does not add up"
assert abs(avgv - sum(k * v for k, v in hist.items()) / float(total)) < 1e-12
print("total words", total, "distinct", len(counts), "avg len", avgv, "top", top[:3])
print("ok")
