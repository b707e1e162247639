"""BASELINE.json's full sizes through size-independent properties (no oracle run is feasible there):
10 GB of text (configs[1]/[3] shape) and 6.25e8 kv records (configs[2] shape), generated on the device."""
import numpy as np
import pytest

from dampr_b200 import device as dev
from dampr_b200 import synth

pytestmark = pytest.mark.gpu


def _count(ctx, tb, lo, hi, mode, log2=21):
    tab = ctx.table(log2)
    try:
        tab.count(tb, lo, hi, mode)
        st = tab.stats()
        assert st["flags"] == 0
        codes, counts, _reps = tab.fetch()
        order = np.argsort(codes)
        return st, codes[order], counts[order].astype(np.int64)
    finally:
        tab.free()


@pytest.mark.parametrize("mode", [dev.TOK_NONWORD_LOWER_SET, dev.TOK_WS], ids=["tfidf", "wc"])
def test_text_10gb_properties(ctx, mode):
    free, _total = ctx.mem_info()
    if free < 40 << 30:
        pytest.skip("needs 40 GB of free device memory")
    V = 1_000_000
    n_lines = int(10e9 / 99.94)
    vocab, cdf = synth.make_vocab(V), synth.make_cdf(V)
    tb = ctx.synth_text(1234, n_lines, vocab[0], vocab[1], cdf)
    try:
        st, codes, counts = _count(ctx, tb, 0, tb.n, mode)
        assert st["lines"] == n_lines
        assert int(counts.sum()) == st["folded"] and len(codes) == st["entries"]
        assert len(np.unique(codes)) == len(codes)
        if mode == dev.TOK_NONWORD_LOWER_SET:
            assert int(counts.max()) <= n_lines      # a document frequency cannot exceed the line count
            assert st["folded"] <= st["raw"]
        else:
            assert st["folded"] == st["raw"]
        # linearity: owner ranges that tile the text count every line exactly once, so the tables of the
        # parts add up to the table of the whole (the multi-GPU sharding rule), wherever the cuts fall
        cuts = [0, (tb.n // 3) // 4096 * 4096, (tb.n // 2 + 12345) // 4096 * 4096, tb.n]
        parts_c, parts_n, lines = [], [], 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            s2, c2, n2 = _count(ctx, tb, a, b, mode)
            lines += s2["lines"]
            parts_c.append(c2)
            parts_n.append(n2)
        u, inv = np.unique(np.concatenate(parts_c), return_inverse=True)
        tot = np.zeros(len(u), dtype=np.int64)
        np.add.at(tot, inv, np.concatenate(parts_n))
        acc = {"codes": u, "counts": tot}
        assert lines == n_lines
        assert np.array_equal(acc["codes"], codes) and np.array_equal(acc["counts"], counts)
    finally:
        tb.free()


def test_kv_10gb_sort_and_fold_properties(ctx):
    free, _total = ctx.mem_info()
    if free < 60 << 30:
        pytest.skip("needs 60 GB of free device memory")
    n, K = 625_000_000, 10_000_000
    kv = ctx.synth_kv(42, n, K)
    try:
        # two independent routes to the per-key sums: the fused shared-memory hash aggregate ...
        a = kv.sort_reduce(dev.OP_SUM_I64, dev.KEY_MIX)
        assert len(kv) == 0          # sort_reduce consumes its input (include/dampr_b200.h)
        ak, av = a.columns()
        a.free()
    finally:
        kv.free()
    # the total is also what ONE group over the same values gives (values do not depend on the key count)
    one = ctx.synth_kv(42, n, 1)
    try:
        t = one.sort_reduce(dev.OP_SUM_I64, dev.KEY_MIX)
        _tk, tv = t.columns()
        t.free()
    finally:
        one.free()
    assert len(tv) == 1 and int(av.view(np.int64).sum()) == int(tv.view(np.int64)[0])
    kv = ctx.synth_kv(42, n, K)
    try:
        # ... and a full stable sort followed by the segmented reduce of the sorted records, which keeps them
        kv.sort(dev.KEY_MIX)
        offs = kv.group_offsets()
        b = kv.reduce_by_key(dev.OP_SUM_I64)
        assert len(kv) == n
        bk, bv = b.columns()
        b.free()
        assert len(offs) - 1 == len(bk) == len(ak) <= K
        assert int(offs[-1]) == n and bool(np.all(np.diff(offs.astype(np.int64)) > 0))
        oa, ob = np.argsort(ak), np.argsort(bk)
        assert np.array_equal(ak[oa], bk[ob]) and np.array_equal(av[oa], bv[ob])
        assert len(np.unique(ak)) == len(ak)
        # counts per key add up to n and equal the group sizes of the sorted run
        c = kv.reduce_by_key(dev.OP_COUNT)
        ck, cv = c.columns()
        c.free()
        assert int(cv.view(np.int64).sum()) == n
    finally:
        kv.free()
