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

    def partition_by_owner(self, nb):
        own = self._order(self.rec[:self.n, 0], 1) % np.uint64(nb)
        o = np.argsort(own, kind="stable")
        out = FakeKV(self.ctx, max(1, self.n))
        out.rec[:self.n] = self.rec[:self.n][o]
        out.n = self.n
        return out, np.bincount(own.astype(np.int64), minlength=nb).astype(np.uint64)

    def sort_reduce(self, op, xform):
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

    def group_offsets(self):
        k = self.rec[:self.n, 0]
        heads = np.flatnonzero(np.concatenate(([True], k[1:] != k[:-1]))) if len(k) else np.zeros(0, dtype=np.int64)
        return np.concatenate((heads, [len(k)])).astype(np.uint64)


class FakeCtx(object):
    def kv(self, capacity):
        return FakeKV(self, capacity)

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

