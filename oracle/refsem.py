"""CPU restatement of the reference semantics on the hot path — TEST INFRASTRUCTURE ONLY.

Written from SURVEY §8(a)/Appendix A, not copied from the reference.  Each function cites the
reference lines it restates.  Pinned against the real reference by tests/golden/make_golden.py
(run in the dev container where /root/reference is importable) -> tests/golden/*.json.
"""
import math
import re
from collections import Counter

import numpy as np

RX = re.compile(r"[^\w]+")  # benchmarks/tf-idf-dampr.py:11


def text_lines(data):
    """Lines of a text file as Dampr.text yields them: (char offset, line without its newline).

    Restates TextInput.chunks (dampr/inputs.py:48-56) + TextLineDataset.read
    (dampr/dataset.py:458-476) under the intended "every line exactly once" semantics: the file is
    opened in text mode (universal newlines: '\\r\\n' and lone '\\r' end lines too), a last line
    without terminator is still a line, an empty file has no lines.
    """
    s = data.decode("utf-8")
    s = s.replace("\r\n", "\n").replace("\r", "\n")
    pos = 0
    out = []
    for line in s.split("\n"):
        out.append((pos, line))
        pos += len(line) + 1
    if out and out[-1][1] == "":
        out.pop()  # text after the final '\n' is empty: not a line
    return out


def wc_counts(data):
    """examples/wc.py:11-13: flat_map(x.split()) -> fold_by(identity, 1, +)."""
    c = Counter()
    for _, line in text_lines(data):
        c.update(line.split())
    return c


def docfreq(data):
    """benchmarks/tf-idf-dampr.py:12-15: flat_map(set(RX.split(x.lower()))).count(); also
    benchmarks/baseline.py:15-18.  Returns (Counter term -> document frequency, n_lines)."""
    c = Counter()
    n = 0
    for _, line in text_lines(data):
        c.update(set(RX.split(line.lower())))
        n += 1
    return c, n


def termfreq_nonset(data):
    """Same tokeniser without set(): every token (and every '' re.split yields) counts."""
    c = Counter()
    for _, line in text_lines(data):
        c.update(RX.split(line.lower()))
    return c


def tfidf_rows(data):
    """Rows the tf-idf script sinks (tf-idf-dampr.py:17-21): (word, df, log(1 + total/df))."""
    c, total = docfreq(data)
    return [(w, df, math.log(1 + (float(total) / df))) for w, df in c.items()]


def tfidf_sink_lines(data):
    """sink_tsv formatting (dampr/dampr.py:521-529): '\\t'.join(str(p) for p in row)."""
    return sorted(u"\t".join(str(p) for p in row) for row in tfidf_rows(data))


# ---- kv records (configs 2, 4, 5) ----------------------------------------------------------
def group_sum(keys, vals):
    """a_group_by(k, v).sum() (dampr/dampr.py:386-404, 701-708): dict key -> exact integer sum."""
    order = np.argsort(keys, kind="stable")
    k = keys[order]
    v = vals[order]
    if len(k) == 0:
        return {}
    heads = np.flatnonzero(np.concatenate(([True], k[1:] != k[:-1])))
    sums = np.add.reduceat(v.astype(np.int64), heads)
    return dict(zip(k[heads].tolist(), sums.tolist()))


def group_count(keys):
    u, c = np.unique(keys, return_counts=True)
    return dict(zip(u.tolist(), c.tolist()))


def group_fold(keys, vals, binop):
    """Generic left fold per key in input order (ARReduce._reduce, dampr/dampr.py:678-683)."""
    acc = {}
    for k, v in zip(keys.tolist(), vals.tolist()):
        acc[k] = binop(acc[k], v) if k in acc else v
    return acc


def group_values(keys, vals):
    """group_by(k, v): key -> list of values in input order (stable sort, dataset.py:162-164)."""
    out = {}
    for k, v in zip(keys.tolist(), vals.tolist()):
        out.setdefault(k, []).append(v)
    return out


def inner_join(lk, lv, rk, rv):
    """InnerJoin.reduce (dampr/base.py:264-283): key -> (left values, right values), keys on both."""
    L = group_values(lk, lv)
    R = group_values(rk, rv)
    return {k: (L[k], R[k]) for k in L if k in R}


def left_join(lk, lv, rk, rv):
    """LeftJoin.reduce (dampr/base.py:295-315): every left key, [] when the right side lacks it."""
    L = group_values(lk, lv)
    R = group_values(rk, rv)
    return {k: (L[k], R.get(k, [])) for k in L}
