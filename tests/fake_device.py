"""numpy stand-in for the device (dampr_b200.device.Ctx / KV): TEST INFRASTRUCTURE ONLY.

It lets the CPU suite (-m "not gpu") exercise the host logic that sits above the C-ABI — the runner's
host-map + shuffle plumbing, key codecs, joins, sinks, the frame lowering, the spill logic — without a
GPU. The product never imports it and has no CPU fallback; every operation here restates, with numpy,
what the corresponding entry point of include/dampr_b200.h is documented to do.
"""
import numpy as np


class FakeKV(object):
    """numpy stand-in for device.KV: enough of the interface for the host logic of spill.py."""

    def __init__(self, ctx, capacity):
        self.ctx = ctx
        self.rec = np.zeros((capacity, 2), dtype=np.uint64)
        self.n = 0

    def __len__(self):
        return self.n

    def upload_columns(self, off, keys, vals):
        k = np.asarray(keys).view(np.uint64)
        self.rec[off:off + len(k), 0] = k
        self.rec[off:off + len(k), 1] = np.asarray(vals).view(np.uint64)
        self.n = max(self.n, off + len(k))

    def upload(self, off, recs, count=None):
        c = len(recs) if count is None else count
        self.rec[off:off + c] = recs[:c]
        self.n = max(self.n, off + c)

    @staticmethod
    def _order(keys, xform):
        from dampr_b200 import spill
        from dampr_b200 import device as dev
        if xform == dev.KEY_MIX:
            x = keys.copy()
            with np.errstate(over="ignore"):
                x ^= x >> np.uint64(30); x *= np.uint64(0xBF58476D1CE4E5B9)
                x ^= x >> np.uint64(27); x *= np.uint64(0x94D049BB133111EB)
                x ^= x >> np.uint64(31)
            return x
        return spill._order_domain(keys, xform)

    def sort(self, xform):
        o = np.argsort(self._order(self.rec[:self.n, 0], xform), kind="stable")
        self.rec[:self.n] = self.rec[:self.n][o]
        return self

    def records(self):
        return self.rec[:self.n].copy()

    def columns(self):
        return self.rec[:self.n, 0].copy(), self.rec[:self.n, 1].copy()

    def records_into(self, out, off=0):
        out[:] = self.rec[off:off + len(out)]
        return out

    def columns_into(self, keys, vals):
        keys[:] = self.rec[:self.n, 0]
        vals[:] = self.rec[:self.n, 1]

    def partition_by_owner(self, nb):
        own = self._order(self.rec[:self.n, 0], 1) % np.uint64(nb)
        o = np.argsort(own, kind="stable")
        out = FakeKV(self.ctx, max(1, self.n))
        out.rec[:self.n] = self.rec[:self.n][o]
        out.n = self.n
        return out, np.bincount(own.astype(np.int64), minlength=nb).astype(np.uint64)

    def sort_reduce(self, op, xform, sorted_run=False):
        from dampr_b200 import device as dev
        self.sort(xform)
        k, v = self.columns()
        heads = np.flatnonzero(np.concatenate(([True], k[1:] != k[:-1]))) if len(k) else np.zeros(0, dtype=np.int64)
        out = FakeKV(self.ctx, max(1, len(heads)))
        vi, vf = v.view(np.int64), v.view(np.float64)
        ends = np.concatenate((heads[1:], [len(k)])).astype(np.int64)
        if len(heads) == 0:
            r = vi[:0]
        elif op == dev.OP_SUM_I64:
            r = np.add.reduceat(vi, heads)
        elif op == dev.OP_COUNT:
            r = ends - heads
        elif op == dev.OP_MIN_I64:
            r = np.minimum.reduceat(vi, heads)
        elif op == dev.OP_MAX_I64:
            r = np.maximum.reduceat(vi, heads)
        elif op == dev.OP_SUM_F64:
            r = np.array([np.cumsum(vf[a:b])[-1] for a, b in zip(heads, ends)]).view(np.int64)   # left fold
        elif op == dev.OP_MIN_F64:
            r = np.minimum.reduceat(vf, heads).view(np.int64)
        elif op == dev.OP_MAX_F64:
            r = np.maximum.reduceat(vf, heads).view(np.int64)
        elif op == dev.OP_LAST:
            r = vi[ends - 1]
        else:
            r = vi[heads]                     # FIRST (stable sort keeps input order)
        out.rec[:len(heads), 0] = k[heads]
        out.rec[:len(heads), 1] = np.asarray(r, dtype=np.int64).view(np.uint64)
        out.n = len(heads)
        self.n = 0
        return out

    def reduce_by_key(self, op):
        """dampr_kv_reduce_by_key: fold adjacent equal keys of a key-sorted kv; the input is left alone"""
        from dampr_b200 import device as dev
        k, v = self.columns()
        heads = np.flatnonzero(np.concatenate(([True], k[1:] != k[:-1]))) if len(k) else np.zeros(0, dtype=np.int64)
        assert len(set(k[heads].tolist())) == len(heads), "reduce_by_key: equal keys are not adjacent"
        tmp = FakeKV(self.ctx, max(1, self.n))
        tmp.rec[:self.n] = self.rec[:self.n]
        tmp.n = self.n
        # group order must stay the input's: reduce in RAW order, then restore the order of first appearance
        red = tmp.sort_reduce(op, dev.KEY_RAW)
        rk, rv = red.columns()
        pos = {int(key): i for i, key in enumerate(k[heads].tolist())}
        order = np.argsort(np.array([pos[int(x)] for x in rk.tolist()], dtype=np.int64), kind="stable")
        out = FakeKV(self.ctx, max(1, len(rk)))
        out.rec[:len(rk), 0] = rk[order]
        out.rec[:len(rk), 1] = rv[order]
        out.n = len(rk)
        return out

    def free(self):
        pass

    # ---- the rest of the KV interface the runner and the planner use ------------------------------
    def join_ranges(self, right, xform):
        """rows[g] = (left_begin, left_end, right_begin, right_end) per LEFT key group; both sides sorted."""
        lk = self.rec[:self.n, 0]
        rk = right.rec[:right.n, 0]
        if len(lk) == 0:
            return np.zeros((0, 4), dtype=np.uint64)
        heads = np.flatnonzero(np.concatenate(([True], lk[1:] != lk[:-1])))
        ends = np.concatenate((heads[1:], [len(lk)]))
        ko = self._order(lk[heads], xform)
        ro = self._order(rk, xform)
        rb = np.searchsorted(ro, ko, side="left")
        re_ = np.searchsorted(ro, ko, side="right")
        return np.stack((heads, ends, rb, re_), axis=1).astype(np.uint64)

    def hash_probe(self, probe):
        """self = build side (unique keys): (values aligned with probe, hit flags)."""
        bk, bv = self.columns()
        o = np.argsort(bk)
        bk, bv = bk[o], bv[o]
        pk = probe.rec[:probe.n, 0]
        pos = np.searchsorted(bk, pk)
        pos = np.minimum(pos, max(0, len(bk) - 1))
        hit = (bk[pos] == pk) if len(bk) else np.zeros(len(pk), dtype=bool)
        out = FakeKV(self.ctx, max(1, len(pk)))
        out.rec[:len(pk), 0] = pk
        out.rec[:len(pk), 1] = np.where(hit, bv[pos] if len(bk) else 0, 0)
        out.n = len(pk)
        return out, hit.astype(np.uint8)

    def hash_join(self, probe):
        """dampr_kv_hash_join: the probe records that found a partner, and (key, build value), in probe order"""
        vals, hit = self.hash_probe(probe)
        m = hit.astype(bool)
        a, b = FakeKV(self.ctx, max(1, int(m.sum()))), FakeKV(self.ctx, max(1, int(m.sum())))
        a.rec[:int(m.sum())] = probe.rec[:probe.n][m]
        b.rec[:int(m.sum())] = vals.rec[:vals.n][m]
        a.n = b.n = int(m.sum())
        return a, b

    def group_offsets(self):
        k = self.rec[:self.n, 0]
        heads = np.flatnonzero(np.concatenate(([True], k[1:] != k[:-1]))) if len(k) else np.zeros(0, dtype=np.int64)
        return np.concatenate((heads, [len(k)])).astype(np.uint64)


class FakeCtx(object):
    def kv(self, capacity):
        return FakeKV(self, capacity)

    def kv_merge(self, runs, xform, op=-1):
        """dampr_kv_merge: the runs MUST be key-sorted under xform (checked here: the device kernel relies on
        it); the result is the stable sort of their concatenation in run order, optionally folded."""
        from dampr_b200 import device as dev
        recs = []
        for r in runs:
            o = FakeKV._order(r.rec[:r.n, 0], xform)
            assert np.all(o[1:] >= o[:-1]), "kv_merge: a run is not sorted under the key transform"
            recs.append(r.rec[:r.n])
        cat = FakeKV(self, max(1, sum(len(x) for x in recs)))
        if recs:
            allr = np.concatenate(recs)
            cat.rec[:len(allr)] = allr
            cat.n = len(allr)
        if op < 0:
            return cat.sort(xform)
        return cat.sort_reduce(op, xform)

    def kv_merge_ranges(self, kv, offsets, xform, op=-1):
        views = []
        offs = [int(x) for x in offsets]
        for a, b in zip(offs[:-1], offs[1:]):
            v = FakeKV(self, max(1, b - a))
            v.rec[:b - a] = kv.rec[a:b]
            v.n = b - a
            views.append(v)
        return self.kv_merge(views, xform, op)

    def sync(self):
        pass

    def mem_info(self):
        return (1 << 30, 1 << 30)

    h = 1   # "open" for runner.get_ctx

    def kv_from_columns(self, keys, vals=None):
        kv = FakeKV(self, max(1, len(keys)))
        kv.upload_columns(0, keys, vals if vals is not None else np.zeros(len(keys), dtype=np.int64))
        return kv

    def kv_from_records(self, recs):
        recs = np.asarray(recs).reshape(-1, 2)
        kv = FakeKV(self, max(1, len(recs)))
        kv.upload(0, recs, len(recs))
        return kv

    def close(self):
        pass

    def textbuf(self, capacity):
        from dampr_b200.plan import NotLowerable
        raise NotLowerable("the numpy stand-in has no tokeniser: text stages run as host maps")

    table = textbuf



# ---- text side: a stand-in tokeniser, so that plan.TextScan's plumbing (layout of several files in one
# buffer, chunked uploads, owner ranges, flag handling, table growth, result fetch) runs on CPU too. The
# token rules restate the reference's lambdas (examples/wc.py:12, benchmarks/tf-idf-dampr.py:12-14) the way
# oracle/refsem.py does; the device kernels are checked against the oracle by the -m gpu suites, not here.
import re as _re

_RX = _re.compile(r"[^\w]+")


class FakeTextBuf(object):
    def __init__(self, ctx, capacity):
        self.ctx = ctx
        self.capacity = int(capacity)
        self.buf = np.full(self.capacity + 64, 10, dtype=np.uint8)   # '\n' padding
        self.n = 0

    def set_length(self, n):
        assert n <= self.capacity
        self.n = int(n)
        self.buf[self.n:] = 10

    def upload(self, off, host, length):
        self.buf[off:off + length] = np.frombuffer(memoryview(host), dtype=np.uint8)[:length] \
            if not isinstance(host, np.ndarray) else host.reshape(-1).view(np.uint8)[:length]

    def upload_file(self, off, path, file_off, length):
        with open(path, "rb") as f:
            f.seek(file_off)
            data = f.read(length)
        assert len(data) == length
        self.buf[off:off + length] = np.frombuffer(data, dtype=np.uint8)

    def download(self, off, length):
        return self.buf[off:off + length].copy()

    def free(self):
        pass


class FakeTable(object):
    def __init__(self, ctx, log2):
        self.ctx = ctx
        self.capacity = 1 << int(log2)
        self.clear()

    def clear(self):
        self.counts, self.first = {}, {}
        self.fb = []
        self.st = {"lines": 0, "empty": 0, "folded": 0, "flags": 0, "hashed": 0, "raw": 0, "fallback": 0}

    def free(self):
        pass

    def fallback_lines(self):
        return np.array(self.fb, dtype=np.uint64)

    def count(self, tb, lo, hi, mode, cr_is_data=False):
        from dampr_b200 import device as dev
        self.cr_is_data = cr_is_data
        data = tb.buf[:tb.n].tobytes()
        pos = 0
        while pos < len(data):
            end = data.find(b"\n", pos)
            nxt = len(data) if end < 0 else end + 1
            line = data[pos:len(data) if end < 0 else end]
            if lo <= pos < hi:
                self._line(line, pos, mode, dev)
            pos = nxt
        if len(self.counts) * 2 > self.capacity:
            self.st["flags"] |= dev.TF_TABLEFULL if len(self.counts) >= self.capacity else 0

    def _line(self, line, off, mode, dev):
        st = self.st
        nonascii = any(b >= 0x80 for b in line)
        cr = (b"\r" in line) and not getattr(self, "cr_is_data", False)
        if mode != dev.TOK_WS and (nonascii or cr):
            # per-line fallback of the [^\w]+ tokenisers: the line contributes nothing here and is handed back
            if len(line) < (1 << 16) and len(self.fb) < (1 << 16):
                self.fb.append((off << 16) | len(line))
                st["fallback"] += 1
            else:
                st["flags"] |= dev.TF_NONASCII
            return
        st["lines"] += 1
        if nonascii:
            st["flags"] |= dev.TF_NONASCII
            return
        text = line.decode("ascii")
        if "\r" in text and mode == dev.TOK_WS:
            st["flags"] |= dev.TF_CR          # str.split mode flags it scan-wide; the host ignores it for the token counts
        if mode == dev.TOK_WS:
            toks, limit = text.split(), 9
        else:
            toks, limit = _RX.split(text.lower()), 12
        st["raw"] += sum(1 for t in toks if t != "")
        if mode == dev.TOK_NONWORD_LOWER_SET:
            toks = set(toks)
        for t in toks:
            if t == "":
                st["empty"] += 1
                continue
            st["folded"] += 1
            if len(t) > limit or "\0" in t:
                st["hashed"] += 1
            if t not in self.counts:
                low = text if mode == dev.TOK_WS else text.lower()
                self.first[t] = off + low.find(t)
            self.counts[t] = self.counts.get(t, 0) + 1

    def stats(self):
        d = dict(self.st)
        d["entries"] = len(self.counts)
        return d

    def verify(self, tb, lo, hi, mode):
        pass

    def fetch_words(self, tb, mode, width=32, with_codes=True):
        toks = list(self.counts)
        words = np.array([t.encode("ascii")[:width] for t in toks], dtype="S%d" % width) if toks \
            else np.zeros(0, dtype="S%d" % width)
        counts = np.array([self.counts[t] for t in toks], dtype=np.uint64)
        if not with_codes:
            return words, counts, None, None
        limit = 9 if mode == 0 else 12
        codes = np.array([(1 << 63 | i) if (len(t) > limit or "\0" in t) else i + 1 for i, t in enumerate(toks)], dtype=np.uint64)
        reps = np.array([(self.first[t] << 20) | len(t) for t in toks], dtype=np.uint64)
        return words, counts, codes, reps


class FakeTextCtx(FakeCtx):
    """FakeCtx plus the text entry points."""

    def textbuf(self, capacity):
        return FakeTextBuf(self, capacity)

    def table(self, log2):
        return FakeTable(self, log2)

    def sync_copy_stream(self):
        pass


class FakePinned(object):
    def __init__(self, nbytes):
        self.array = np.zeros(int(nbytes), dtype=np.uint8)
