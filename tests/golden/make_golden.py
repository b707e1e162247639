"""Generate golden vectors by running the REAL reference (Refefer/Dampr, /root/reference) in the dev
container.  The reference cannot travel to the GPU box, so its outputs on the deterministic
synthetic inputs of oracle/gen.py are committed here as tests/golden/*.json.

    python tests/golden/make_golden.py

Inputs are not stored: they are regenerated from (generator, seed, size) recorded in each fixture.
"""
import json
import math
import os
import re
import sys
import tempfile
import unittest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DAMPR_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import dampr as ref  # the reference package (REF precedes the repo's shim on sys.path)
assert os.path.realpath(os.path.dirname(ref.__file__)).startswith(os.path.realpath(REF)), ref.__file__
from dampr import Dampr, Dataset
from dampr.dataset import Chunker
from oracle import gen

import numpy as np


class KVSlice(Dataset):
    def __init__(self, keys, vals, lo, hi):
        self.keys, self.vals, self.lo, self.hi = keys, vals, lo, hi

    def read(self):
        for i in range(self.lo, self.hi):
            yield i, (int(self.keys[i]), int(self.vals[i]))


class KVChunks(Chunker):
    def __init__(self, keys, vals, n_chunks=8):
        self.keys, self.vals, self.n_chunks = keys, vals, n_chunks

    def chunks(self):
        n = len(self.keys)
        step = max(1, -(-n // self.n_chunks))
        for lo in range(0, n, step):
            yield KVSlice(self.keys, self.vals, lo, min(n, lo + step))


def write(name, obj):
    path = os.path.join(HERE, name)
    with open(path, "w") as f:
        json.dump(obj, f, separators=(",", ":"), sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


def text_fixture(name, data, spec):
    tmp = tempfile.mkdtemp(prefix="golden_")
    path = os.path.join(tmp, "corpus.txt")
    with open(path, "wb") as f:
        f.write(data)
    RX = re.compile(r"[^\w]+")
    chunk = max(64, len(data) // 8 + 1)  # integral chunk size: the reference is exact
    docs = Dampr.text(path, chunk)
    # examples/wc.py
    wc = docs.flat_map(lambda x: x.split()) \
        .fold_by(lambda x: x, value=lambda x: 1, binop=lambda x, y: x + y) \
        .sort_by(lambda x: -x[1])
    wc_rows = sorted(wc.run("golden-wc").read())
    # benchmarks/tf-idf-dampr.py
    doc_freq = docs.flat_map(lambda x: set(RX.split(x.lower()))).count(reduce_buffer=float("inf"))
    idf = doc_freq.cross_right(docs.len(),
                               lambda df, total: (df[0], df[1], math.log(1 + (float(total) / df[1]))),
                               memory=True)
    sink = os.path.join(tmp, "idfs")
    idf.sink_tsv(sink).run("golden-tfidf")
    lines = []
    for fn in sorted(os.listdir(sink)):
        with open(os.path.join(sink, fn)) as f:
            lines.extend(l.rstrip("\n") for l in f)
    n_lines = docs.len().read()[0]
    write(name, {"spec": spec, "bytes": len(data), "n_lines": n_lines, "wc": wc_rows, "tfidf_lines": sorted(lines)})


def kv_fixture():
    keys, vals = gen.kv(42, 20000, 700)
    src = Dampr.read_input(KVChunks(keys, vals))
    agg = sorted(src.a_group_by(lambda x: x[0], lambda x: x[1]).sum().read())
    cnt = sorted(src.count(lambda x: x[0]).read())
    grp = sorted(src.group_by(lambda x: x[0], lambda x: x[1]).reduce(lambda k, it: sum(it)).read())
    mx = sorted(src.a_group_by(lambda x: x[0], lambda x: x[1]).reduce(max).read())
    mean = sorted(src.mean(lambda x: x[0], lambda x: x[1]).read())
    srt = src.map(lambda x: x[1]).sort_by(lambda v: v).read()
    lk, lv = gen.kv(1, 6000, 900)
    rk, rv = gen.kv(2, 800, 1800)
    L = Dampr.read_input(KVChunks(lk, lv)).group_by(lambda x: x[0], lambda x: x[1])
    R = Dampr.read_input(KVChunks(rk, rv)).group_by(lambda x: x[0], lambda x: x[1])
    inner = sorted((k, sorted(l), sorted(r)) for k, (l, r) in L.join(R).reduce(lambda l, r: (list(l), list(r))).read())
    left = sorted((k, sorted(l), sorted(r)) for k, (l, r) in L.join(R).left_reduce(lambda l, r: (list(l), list(r))).read())
    small = Dampr.read_input(KVChunks(rk, rv)).map(lambda x: x[0])
    big = Dampr.read_input(KVChunks(lk, lv))
    probe = sorted(small.cross_set(big, lambda b, table: (b[0], b[0] in table), agg=set).read())
    write("kv.json", {
        "spec": {"agg": "gen.kv(42,20000,700)", "join": "gen.kv(1,6000,900) x gen.kv(2,800,1800)"},
        "sum": agg, "count": cnt, "group_sum": grp, "max": mx, "mean": mean, "sorted_vals_head": sorted(srt)[:50],
        "sorted_vals_len": len(srt), "reference_output_is_sorted": srt == sorted(srt),
        "sorted_vals_sha256": __import__("hashlib").sha256(json.dumps(sorted(srt)).encode()).hexdigest(),
        "inner": inner, "left_len": len(left), "left_nomatch": sum(1 for _k, _l, r in left if not r),
        "probe_true": sum(1 for _k, hit in probe if hit), "probe_len": len(probe)})


def reference_suite():
    """The reference's own tests with the removed assertEquals aliased; URL test skipped."""
    unittest.TestCase.assertEquals = unittest.TestCase.assertEqual
    sys.path.insert(0, REF)
    import tests.test_dampr as t
    suite = unittest.TestSuite()
    for nm in unittest.TestLoader().getTestCaseNames(t.DamprTest):
        if nm != "test_read_url":
            suite.addTest(t.DamprTest(nm))
    res = unittest.TextTestRunner(verbosity=0).run(suite)
    write("reference_suite.json", {"run": res.testsRun, "failures": len(res.failures), "errors": len(res.errors)})


if __name__ == "__main__":
    vocab = gen.make_vocab(2000)
    cdf = gen.make_cdf(2000)
    text_fixture("text_zipf.json", gen.text(1234, 3000, vocab=vocab, cdf=cdf),
                 {"gen": "gen.text(1234, 3000, vocab=make_vocab(2000), cdf=make_cdf(2000))"})
    text_fixture("text_dirty.json", gen.dirty_text(7, 1500, 2000), {"gen": "gen.dirty_text(7, 1500, 2000)"})
    kv_fixture()
    if "--suite" in sys.argv:
        reference_suite()
