"""GPU parity: partition + sort + segmented reduce + join + probe (csrc/kv.cu, ops.cu) vs oracle."""
import numpy as np
import pytest

from dampr_b200 import device as dev
from oracle import gen, refsem

pytestmark = pytest.mark.gpu


def xf_key(keys, xf):
    k = keys.astype(np.uint64)
    if xf == dev.KEY_I64:
        return k ^ np.uint64(1 << 63)
    if xf == dev.KEY_F64:
        neg = (k >> np.uint64(63)).astype(bool)
        return np.where(neg, ~k, k | np.uint64(1 << 63))
    return k


def mix64(x):
    x = x.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def check_sort(ctx, keys, vals, xf):
    kv = ctx.kv_from_columns(keys, vals)
    kv.sort(xf)
    k2, v2 = kv.columns()
    sk = mix64(keys.view(np.uint64)) if xf == dev.KEY_MIX else xf_key(keys.view(np.uint64), xf)
    order = np.argsort(sk, kind="stable")
    assert np.array_equal(k2, keys.view(np.uint64)[order])
    assert np.array_equal(v2, vals.view(np.uint64)[order])  # stable: ties keep input order
    kv.free()


@pytest.mark.parametrize("n", [0, 1, 2, 31, 1000, 4096, 4097, 50000, 300000])
@pytest.mark.parametrize("xf", [dev.KEY_RAW, dev.KEY_MIX, dev.KEY_I64])
def test_sort_sizes(ctx, n, xf):
    keys, vals = gen.kv(42, n, max(1, n // 3))
    check_sort(ctx, keys, vals, xf)


def test_sort_large_two_levels(ctx):
    keys, vals = gen.kv(7, 6_000_000, 5_000_000)
    check_sort(ctx, keys, vals, dev.KEY_MIX)
    check_sort(ctx, keys, vals, dev.KEY_RAW)


def test_sort_skew_and_duplicates(ctx):
    rng = np.random.default_rng(3)
    n = 400000
    keys = rng.zipf(1.3, size=n).astype(np.uint64)  # heavy hitters
    vals = np.arange(n, dtype=np.int64)
    for xf in (dev.KEY_MIX, dev.KEY_RAW):
        check_sort(ctx, keys, vals, xf)
    keys[:] = 17  # a single key
    check_sort(ctx, keys, vals, dev.KEY_MIX)
    keys = np.arange(n, dtype=np.uint64)[::-1].copy()  # dense small ints, reversed
    check_sort(ctx, keys, vals, dev.KEY_RAW)


def test_sort_signed_and_float(ctx):
    rng = np.random.default_rng(9)
    n = 100000
    ik = rng.integers(-10**12, 10**12, size=n).astype(np.int64)
    vals = np.arange(n, dtype=np.int64)
    kv = ctx.kv_from_columns(ik.view(np.uint64), vals)
    kv.sort(dev.KEY_I64)
    k2, _ = kv.columns()
    assert np.array_equal(k2.view(np.int64), np.sort(ik, kind="stable"))
    fk = rng.standard_normal(n) * 1e6
    kv = ctx.kv_from_columns(fk.view(np.uint64), vals)
    kv.sort(dev.KEY_F64)
    k2, _ = kv.columns()
    assert np.array_equal(k2.view(np.float64), np.sort(fk, kind="stable"))


@pytest.mark.parametrize("n,nk", [(0, 1), (1, 1), (5000, 10), (5000, 5000), (200000, 1000), (3_000_000, 400_000)])
def test_sort_reduce_sum_count(ctx, n, nk):
    keys, vals = gen.kv(11, n, nk)
    kv = ctx.kv_from_columns(keys, vals)
    out = kv.sort_reduce(dev.OP_SUM_I64, dev.KEY_MIX)
    k2, v2 = out.columns()
    got = dict(zip(k2.tolist(), v2.view(np.int64).tolist()))
    assert len(got) == len(k2)
    assert got == refsem.group_sum(keys, vals)
    kv2 = ctx.kv_from_columns(keys, vals)
    out = kv2.sort_reduce(dev.OP_COUNT, dev.KEY_RAW)
    k2, v2 = out.columns()
    assert np.array_equal(k2, np.sort(k2))  # RAW transform returns groups in key order
    assert dict(zip(k2.tolist(), v2.tolist())) == refsem.group_count(keys)


def test_reduce_ops(ctx):
    keys, vals = gen.kv(5, 100000, 300)
    for op, f in ((dev.OP_MIN_I64, min), (dev.OP_MAX_I64, max), (dev.OP_FIRST, lambda a, b: a),
                  (dev.OP_LAST, lambda a, b: b)):
        kv = ctx.kv_from_columns(keys, vals)
        out = kv.sort_reduce(op, dev.KEY_MIX)
        k2, v2 = out.columns()
        assert dict(zip(k2.tolist(), v2.view(np.int64).tolist())) == refsem.group_fold(keys, vals, f)


def test_reduce_float_sum_tolerance(ctx):
    """fp tolerance: |x - fsum| <= 4 n 2^-53 sum|v|."""
    import math
    rng = np.random.default_rng(1)
    n = 200000
    keys = rng.integers(0, 50, size=n).astype(np.uint64)
    vals = rng.standard_normal(n) * 1e3
    kv = ctx.kv_from_columns(keys, vals.view(np.uint64))
    out = kv.sort_reduce(dev.OP_SUM_F64, dev.KEY_MIX)
    k2, v2 = out.columns()
    got = dict(zip(k2.tolist(), v2.view(np.float64).tolist()))
    for k in range(50):
        sel = vals[keys == k]
        exact = math.fsum(sel.tolist())
        tol = 4 * len(sel) * 2.0 ** -53 * float(np.abs(sel).sum())
        assert abs(got[k] - exact) <= tol


def test_reduce_consumes_or_keeps_its_input_as_documented(ctx):
    """dampr_kv_sort_reduce consumes its input; dampr_kv_reduce_by_key leaves the sorted run alone, also
    when segments are larger than a leaf (heavy keys) and the sizes force one or two partition levels."""
    for n, nk in ((300_000, 40), (3_000_000, 1500)):
        keys, vals = gen.kv(31, n, nk)
        kv = ctx.kv_from_columns(keys, vals)
        out = kv.sort_reduce(dev.OP_SUM_I64, dev.KEY_RAW)
        assert len(kv) == 0
        k1, v1 = out.columns()
        assert dict(zip(k1.tolist(), v1.view(np.int64).tolist())) == refsem.group_sum(keys, vals)
        kv = ctx.kv_from_columns(keys, vals).sort(dev.KEY_RAW)
        before = kv.records().copy()
        for op, exp in ((dev.OP_SUM_I64, refsem.group_sum(keys, vals)), (dev.OP_COUNT, refsem.group_count(keys))):
            red = kv.reduce_by_key(op)
            rk, rv = red.columns()
            assert dict(zip(rk.tolist(), rv.view(np.int64).tolist())) == exp
            assert len(kv) == n and np.array_equal(kv.records(), before)


def test_heavy_single_key_reduce(ctx):
    n = 1_000_000
    keys = np.full(n, 12345, dtype=np.uint64)
    keys[::1000] = 99
    vals = np.ones(n, dtype=np.int64)
    kv = ctx.kv_from_columns(keys, vals)
    out = kv.sort_reduce(dev.OP_SUM_I64, dev.KEY_MIX)
    k2, v2 = out.columns()
    assert dict(zip(k2.tolist(), v2.tolist())) == {12345: n - 1000, 99: 1000}


def test_group_offsets_and_merge(ctx):
    keys, vals = gen.kv(21, 50000, 700)
    kv = ctx.kv_from_columns(keys, vals).sort(dev.KEY_RAW)
    offs = kv.group_offsets()
    k2, _ = kv.columns()
    heads = np.flatnonzero(np.concatenate(([True], k2[1:] != k2[:-1])))
    assert np.array_equal(offs[:-1], heads) and offs[-1] == len(k2)
    # k-way merge of sorted runs == stable sort of the concatenation; with reduce == group sums
    runs = []
    allk, allv = [], []
    for s in range(5):
        k, v = gen.kv(100 + s, 20000, 900)
        runs.append(ctx.kv_from_columns(k, v).sort(dev.KEY_RAW))
        allk.append(k)
        allv.append(v)
    merged = dev.kv_merge(ctx, runs, dev.KEY_RAW)
    mk, _ = merged.columns()
    assert np.array_equal(mk, np.sort(np.concatenate(allk)))
    red = dev.kv_merge(ctx, runs, dev.KEY_RAW, dev.OP_SUM_I64)
    rk, rv = red.columns()
    assert dict(zip(rk.tolist(), rv.view(np.int64).tolist())) == refsem.group_sum(np.concatenate(allk), np.concatenate(allv))


def test_join_ranges(ctx):
    lk, lv = gen.kv(1, 30000, 2000)
    rk, rv = gen.kv(2, 3000, 4000)
    L = ctx.kv_from_columns(lk, lv).sort(dev.KEY_MIX)
    R = ctx.kv_from_columns(rk, rv).sort(dev.KEY_MIX)
    rows = L.join_ranges(R, dev.KEY_MIX)
    Lk, Lv = L.columns()
    Rk, Rv = R.columns()
    got_inner, got_left = {}, {}
    for lb, le, rb, re_ in rows.tolist():
        k = int(Lk[lb])
        pair = (Lv[lb:le].view(np.int64).tolist(), Rv[rb:re_].view(np.int64).tolist())
        got_left[k] = pair
        if re_ > rb:
            got_inner[k] = pair
    assert got_inner == refsem.inner_join(lk, lv, rk, rv)
    assert got_left == refsem.left_join(lk, lv, rk, rv)


def test_hash_probe(ctx):
    bk = np.unique(gen.kv(3, 20000, 50000)[0])
    bv = np.arange(len(bk), dtype=np.int64)
    pk, pv = gen.kv(4, 100000, 50000)
    B = ctx.kv_from_columns(bk, bv)
    P = ctx.kv_from_columns(pk, pv)
    vals, hit = B.hash_probe(P)
    k2, v2 = vals.columns()
    table = dict(zip(bk.tolist(), bv.tolist()))
    exp_hit = np.array([k in table for k in pk.tolist()], dtype=np.uint8)
    assert np.array_equal(hit, exp_hit)
    exp_val = np.array([table.get(k, 0) for k in pk.tolist()], dtype=np.int64)
    assert np.array_equal(v2.view(np.int64), exp_val)
    assert np.array_equal(k2, pk)


def test_partition_by_owner(ctx):
    keys, vals = gen.kv(8, 100000, 30000)
    kv = ctx.kv_from_columns(keys, vals)
    for nd in (1, 2, 3, 8):
        out, counts = kv.partition_by_owner(nd)
        k2, v2 = out.columns()
        owner = (mix64(keys) % np.uint64(nd)).astype(np.int64)
        order = np.argsort(owner, kind="stable")
        assert np.array_equal(k2, keys[order]) and np.array_equal(v2.view(np.int64), vals[order])
        assert np.array_equal(counts, np.bincount(owner, minlength=nd).astype(np.uint64))


def test_synth_kv_matches_numpy(ctx):
    kv = ctx.kv(10000)
    ctx.check(ctx.lib.dampr_synth_kv(ctx.h, kv.h, 42, 10000, 777))
    k2, v2 = kv.columns()
    ek, ev = gen.kv(42, 10000, 777)
    assert np.array_equal(k2, ek) and np.array_equal(v2.view(np.int64), ev)
