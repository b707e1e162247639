"""Out-of-core shuffle for kv records that do not fit the device arena (BASELINE config 4).

Reference behaviour replaced: MaxMemoryWriter / CSDatasetWriter spill a sorted run per partition
whenever the worker's RSS grew by settings.max_memory_per_worker (dataset.py:190-262,
memory.py:72-113), then ReduceStageRunner merges the runs of one partition at a time
(stagerunner.py:269-282, dataset.py:571-579).

Here the spill trigger is the device arena (settings.device_arena_bytes, default 70 % of HBM):
  pass 1  stream the input in arena-sized batches; each batch is SORTED (and folded, when the stage has a
          combiner) on the device and cut into P key ranges of the sort order; every range slice is copied
          device -> host as one sorted run of its bucket (the reference's "one sorted run per partition per
          spill", dataset.py:162-164, 236-253);
  pass 2  one bucket at a time: host -> device, k-way merge of its runs (+ segmented reduce) in one read and
          one write (csrc/merge.cu), results streamed back.  A key lives in exactly one bucket, so buckets are
          independent, exactly like the reference's reduce partitions.
PCIe traffic: 2 x 16 B per record each way; the device never holds more than one batch / bucket.
Grouping order (mixed-key order inside a bucket, buckets in owner order) is unobservable in results
(SURVEY "Result-order contract").

sort_by needs a globally ordered result, so external_sort buckets by KEY RANGE instead: splitters are
quantiles of a key sample, every batch is sorted on the device and cut at the splitters into sorted
runs (the reference's sorted run per spill, dataset.py:162-164), and each range is then sorted once
more on the device by the k-way merge of its runs in batch order -- the merge is stable, so equal keys keep their
input order exactly like heapq.merge over runs (dataset.py:571-579).
"""
import numpy as np

from . import device as dev
from . import settings


def arena_bytes(ctx):
    if settings.device_arena_bytes:
        return int(settings.device_arena_bytes)
    try:
        free, total = ctx.mem_info()
        return int(0.7 * total)
    except Exception:
        return 96 << 30


RECORD_FOOTPRINT = 48  # bytes of device memory per record during a sort: data + ping-pong + output


def needs_spill(ctx, n_records):
    return n_records * RECORD_FOOTPRINT > arena_bytes(ctx)


def _upload_chunks(ctx, kchunks, vchunks):
    """One device kv from column chunks, uploaded back to back (no host-side concatenation)."""
    kv = ctx.kv(max(1, sum(len(k) for k in kchunks)))
    off = 0
    for k, v in zip(kchunks, vchunks):
        kv.upload_columns(off, k, v)
        off += len(k)
    ctx.sync()
    return kv


def _upload_runs(ctx, runs):
    """One device kv from (n_i, 2) uint64 record runs, uploaded back to back in run order."""
    kv = ctx.kv(max(1, sum(len(r) for r in runs)))
    off = 0
    for r in runs:
        kv.upload(off, r, len(r))
        off += len(r)
    ctx.sync()
    return kv


def _mix64(x):
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def _lazy_cuts(keys, splitters, dom):
    """np.searchsorted(dom(keys), splitters, side="left") for keys already sorted under `dom`, evaluating the
    order domain only at the probed elements (a batch holds hundreds of millions of keys)."""
    n = len(keys)
    out = np.empty(len(splitters), dtype=np.int64)
    lo0 = 0
    for j, sp in enumerate(np.asarray(splitters, dtype=np.uint64).tolist()):
        lo, hi = lo0, n
        while lo < hi:
            mid = (lo + hi) >> 1
            if int(dom(keys[mid:mid + 1])[0]) < sp:
                lo = mid + 1
            else:
                hi = mid
        out[j] = lo
        lo0 = lo
    return out


def external_group(ctx, chunk_iter, n_records, op=None, xform=dev.KEY_MIX):
    """chunk_iter yields (keys uint64[], vals 8-byte[]) column chunks. Returns a list of
    (keys, vals) numpy result pieces: key-sorted (under `xform`) inside each piece; with `op` one
    record per key. Also returns stats {"buckets", "batches", "spilled_bytes"}.

    pass 1  every arena-sized batch is sorted on the device — folded as well when `op` is given, which is
            the reference's map-side combine before a spill (ReducedWriter.flush, dataset.py:107-117) and
            bounds what a hot key can put into one bucket — and cut into P key RANGES of the sort order
            (equal ranges of the mixed key under KEY_MIX, where keys are uniform): every bucket receives one
            SORTED RUN per batch (SortedWriter, dataset.py:162-164);
    pass 2  one bucket at a time: its runs go back to the device and are k-way merged (+ folded) in one read
            and one write (dampr_kv_merge_ranges, csrc/merge.cu) — MergeDataset.read (dataset.py:571-579)
            + PartialReduceCombiner (base.py:393-402) — instead of being sorted again."""
    arena = arena_bytes(ctx)
    per_batch = max(1 << 16, arena // RECORD_FOOTPRINT)
    n_buckets = max(2, int(np.ceil(1.3 * n_records / float(per_batch))))
    mixed = xform == dev.KEY_MIX
    if not mixed:
        raise ValueError("external_group buckets by ranges of the mixed key; use external_sort for ordered keys")
    # bucket b holds mix64(key) in [b * 2^64 / P, (b + 1) * 2^64 / P)
    splitters = np.array([(b << 64) // n_buckets for b in range(1, n_buckets)], dtype=np.uint64)
    buckets = [[] for _ in range(n_buckets)]
    stats = {"buckets": n_buckets, "batches": 0, "spilled_bytes": 0, "arena_bytes": arena,
             "seconds": {"upload": 0.0, "sort": 0.0, "download": 0.0, "cut": 0.0, "merge_upload": 0.0, "merge": 0.0,
                         "merge_download": 0.0}}
    sec = stats["seconds"]
    op2 = dev.OP_SUM_I64 if op == dev.OP_COUNT else op   # partial counts are added up
    import time as _time

    def flush_batch(kchunks, vchunks):
        t0 = _time.perf_counter()
        kv = _upload_chunks(ctx, kchunks, vchunks)
        t1 = _time.perf_counter()
        try:
            if op is None:
                kv.sort(xform)
                ctx.sync()
                t2 = _time.perf_counter()
                recs = kv.records()
            else:
                red = kv.sort_reduce(op, xform, sorted_run=True)
                ctx.sync()
                t2 = _time.perf_counter()
                try:
                    recs = red.records()
                finally:
                    red.free()
        finally:
            kv.free()
        t3 = _time.perf_counter()
        sec["upload"] += t1 - t0
        sec["sort"] += t2 - t1
        sec["download"] += t3 - t2
        cuts = _lazy_cuts(recs[:, 0], splitters, _mix64)
        sec["cut"] += _time.perf_counter() - t3
        edges = [0] + cuts.tolist() + [len(recs)]
        for b in range(n_buckets):
            if edges[b + 1] > edges[b]:
                buckets[b].append(recs[edges[b]:edges[b + 1]])
        stats["batches"] += 1
        stats["spilled_bytes"] += recs.nbytes

    pend_k, pend_v, pend_n = [], [], 0
    for keys, vals in chunk_iter:
        pos = 0
        while pos < len(keys):
            take = min(len(keys) - pos, per_batch - pend_n)
            pend_k.append(keys[pos:pos + take])
            pend_v.append(np.asarray(vals[pos:pos + take]).view(np.uint64))
            pend_n += take
            pos += take
            if pend_n >= per_batch:
                flush_batch(pend_k, pend_v)
                pend_k, pend_v, pend_n = [], [], 0
    if pend_n:
        flush_batch(pend_k, pend_v)

    out = []
    for b in range(n_buckets):
        runs, buckets[b] = buckets[b], None
        if not runs:
            continue
        if len(runs) == 1:   # one sorted (folded) run: nothing to merge
            out.append((runs[0][:, 0].copy(), runs[0][:, 1].copy()))
            continue
        out.extend(_merge_runs(ctx, runs, xform, -1 if op is None else op2, per_batch, sec))
    return out, stats


def _merge_runs(ctx, runs, xform, op, per_batch, sec=None):
    """[(keys, vals)] — the k-way merge (+ fold, op >= 0) of sorted host runs on the device. Runs that do not
    fit the arena together (a skewed bucket) are merged in key-range slices: the runs are cut at sampled
    splitters so that every slice fits, and the slices come out in key order."""
    total = sum(len(r) for r in runs)
    if total <= per_batch:
        import time as _time
        t0 = _time.perf_counter()
        kv = _upload_runs(ctx, runs)
        t1 = _time.perf_counter()
        try:
            offs = np.concatenate(([0], np.cumsum([len(r) for r in runs]))).astype(np.uint64)
            m = ctx.kv_merge_ranges(kv, offs, xform, op)
            ctx.sync()
        finally:
            kv.free()
        t2 = _time.perf_counter()
        try:
            k, v = m.columns()
        finally:
            m.free()
        if sec is not None:
            sec["merge_upload"] += t1 - t0
            sec["merge"] += t2 - t1
            sec["merge_download"] += _time.perf_counter() - t2
        return [(k, v)]
    dom = (lambda k: _mix64(k)) if xform == dev.KEY_MIX else (lambda k: _order_domain(k, xform))
    rng = np.random.default_rng(len(runs))
    samp = np.unique(np.concatenate([dom(r[rng.integers(0, len(r), size=min(len(r), 8192)), 0]) for r in runs]))
    parts = max(2, int(np.ceil(2.0 * total / per_batch)))
    q = np.maximum(1, (np.arange(1, parts) * len(samp)) // parts)
    spl = np.unique(samp[q]) if len(samp) > 1 else np.zeros(0, dtype=np.uint64)
    if len(spl) == 0:
        # a single key fills the bucket: fold / concatenate run by run (runs in order = the stable result)
        if op < 0:
            return [(r[:, 0].copy(), r[:, 1].copy()) for r in runs]
        acc = None
        for r in runs:
            kv = ctx.kv_from_records(r)
            red = kv.reduce_by_key(op)
            kv.free()
            piece = red.records()
            red.free()
            if acc is None:
                acc = piece
            else:
                kv2 = ctx.kv_from_records(np.concatenate((acc, piece)))
                red2 = kv2.reduce_by_key(dev.OP_SUM_I64 if op == dev.OP_COUNT else op)
                kv2.free()
                acc = red2.records()
                red2.free()
        return [(acc[:, 0].copy(), acc[:, 1].copy())]
    cuts = [np.concatenate(([0], _lazy_cuts(r[:, 0], spl, dom), [len(r)])) for r in runs]
    out = []
    for j in range(len(spl) + 1):
        sub = [r[c[j]:c[j + 1]] for r, c in zip(runs, cuts) if c[j + 1] > c[j]]
        if not sub:
            continue
        if sum(len(x) for x in sub) > per_batch and len(spl) + 1 > 1 and \
                sum(len(x) for x in sub) < total:
            out.extend(_merge_runs(ctx, sub, xform, op, per_batch))
        elif len(sub) == 1:
            out.append((sub[0][:, 0].copy(), sub[0][:, 1].copy()))
        else:
            kv = _upload_runs(ctx, sub)
            try:
                offs = np.concatenate(([0], np.cumsum([len(x) for x in sub]))).astype(np.uint64)
                m = ctx.kv_merge_ranges(kv, offs, xform, op)
            finally:
                kv.free()
            try:
                out.append(m.columns())
            finally:
                m.free()
    return out


def _order_domain(keys_u64, xform):
    """uint64 view of the keys whose unsigned order is the sort order under `xform`."""
    if xform == dev.KEY_RAW:
        return keys_u64
    if xform == dev.KEY_I64:
        return keys_u64 ^ np.uint64(1 << 63)
    if xform == dev.KEY_F64:
        neg = (keys_u64 >> np.uint64(63)).astype(bool)
        return np.where(neg, ~keys_u64, keys_u64 ^ np.uint64(1 << 63))
    raise ValueError("external_sort needs an order-preserving key transform")


def external_sort(ctx, chunk_iter, n_records, xform, sample_keys, _depth=0):
    """Globally key-ordered (stable) sort of records that do not fit the device arena. chunk_iter yields
    (keys uint64[], vals 8-byte[]) chunks in input order; sample_keys is a random sample of the keys.
    Returns (pieces, stats): pieces = [(keys, vals)] in ascending key order."""
    arena = arena_bytes(ctx)
    per_batch = max(1 << 16, arena // RECORD_FOOTPRINT)
    n_buckets = max(2, int(np.ceil(1.3 * n_records / float(per_batch))))
    samp = np.unique(_order_domain(np.asarray(sample_keys).view(np.uint64), xform))
    # a splitter is never the smallest sampled key, so the range below the first splitter is not empty
    q = np.maximum(1, (np.arange(1, n_buckets) * len(samp)) // n_buckets)
    splitters = np.unique(samp[q]) if len(samp) > 1 else np.zeros(0, dtype=np.uint64)
    nb = len(splitters) + 1
    buckets = [[] for _ in range(nb)]
    sizes = [0] * nb
    stats = {"buckets": nb, "batches": 0, "spilled_bytes": 0, "arena_bytes": arena}

    def flush_batch(kchunks, vchunks):
        kv = _upload_chunks(ctx, kchunks, vchunks)
        try:
            kv.sort(xform)
            recs = kv.records()
        finally:
            kv.free()
        cuts = _lazy_cuts(recs[:, 0], splitters, lambda k: _order_domain(k, xform))
        edges = [0] + cuts.tolist() + [len(recs)]
        for b in range(nb):
            if edges[b + 1] > edges[b]:
                buckets[b].append(recs[edges[b]:edges[b + 1]])
                sizes[b] += edges[b + 1] - edges[b]
        stats["batches"] += 1
        stats["spilled_bytes"] += recs.nbytes

    pend_k, pend_v, pend_n = [], [], 0
    for keys, vals in chunk_iter:
        pos = 0
        while pos < len(keys):
            take = min(len(keys) - pos, per_batch - pend_n)
            pend_k.append(np.asarray(keys[pos:pos + take]).view(np.uint64))
            pend_v.append(np.asarray(vals[pos:pos + take]).view(np.uint64))
            pend_n += take
            pos += take
            if pend_n >= per_batch:
                flush_batch(pend_k, pend_v)
                pend_k, pend_v, pend_n = [], [], 0
    if pend_n:
        flush_batch(pend_k, pend_v)

    out = []
    for b in range(nb):
        runs, buckets[b] = buckets[b], None
        if not runs:
            continue
        if len(runs) == 1:  # one sorted run: already in order
            out.append((runs[0][:, 0].copy(), runs[0][:, 1].copy()))
            continue
        if sizes[b] > per_batch:
            lo = min(int(_order_domain(r[:1, 0], xform)[0]) for r in runs)
            hi = max(int(_order_domain(r[-1:, 0], xform)[0]) for r in runs)
            if lo == hi:  # a single heavy key: runs in batch order are the stable result
                for r in runs:
                    out.append((r[:, 0].copy(), r[:, 1].copy()))
                continue
            if _depth < 16:  # skewed range: split it again on a sample of its own keys
                rng = np.random.default_rng(b + 1)
                sub = [r[rng.integers(0, len(r), size=min(len(r), 4096)), 0] for r in runs]
                sub.append(np.concatenate([r[:1, 0] for r in runs] + [r[-1:, 0] for r in runs]))
                pieces, st = external_sort(ctx, ((r[:, 0], r[:, 1]) for r in runs), sizes[b], xform,
                                           np.concatenate(sub), _depth + 1)
                out.extend(pieces)
                stats["batches"] += st["batches"]
                stats["spilled_bytes"] += st["spilled_bytes"]
                continue
        # sorted runs in batch order -> k-way merge (stable: ties keep the batch order, like heapq.merge)
        out.extend(_merge_runs(ctx, runs, xform, -1, max(per_batch, sizes[b])))
    return out, stats
