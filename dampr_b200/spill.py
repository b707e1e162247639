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
PCIe traffic: 2 x 16 B per record each way. The device holds the batch / bucket being worked on and the NEXT one,
which a host thread uploads meanwhile (_Upload / _pipelined: its own stream, staging ring and copy threads), so
both PCIe directions are busy; the runs live in one process-wide host buffer (_RunArena) that is kept between
jobs and page-locked in the background once a job has touched it; results are written in place (_OutCols).
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
# ... and while the next batch is uploaded next to the one being sorted (pass 1 / pass 2 overlap)
RECORD_FOOTPRINT_OVERLAPPED = 64


def needs_spill(ctx, n_records):
    return n_records * RECORD_FOOTPRINT > arena_bytes(ctx)


def _batches(chunk_iter, per_batch):
    """Regroup (keys, vals) column chunks into batches of at most per_batch records: lists of (k, v) slices."""
    pend, pend_n = [], 0
    for keys, vals in chunk_iter:
        pos = 0
        while pos < len(keys):
            take = min(len(keys) - pos, per_batch - pend_n)
            pend.append((np.asarray(keys[pos:pos + take]).view(np.uint64), np.asarray(vals[pos:pos + take]).view(np.uint64)))
            pend_n += take
            pos += take
            if pend_n >= per_batch:
                yield pend
                pend, pend_n = [], 0
    if pend_n:
        yield pend


class _Upload(object):
    """One device kv being filled by a host thread while the caller keeps the device busy with the previous
    batch: PCIe is full duplex and the staged copies release the GIL, so the upload of batch i+1 runs next to the
    sort and the download of batch i. The kv is allocated by the CALLER's thread (the device block pool is not
    thread-safe); the thread only stages and enqueues copies on the copy stream (its own page-locked ring)."""

    def __init__(self, ctx, parts, columns):
        import threading
        self.ctx = ctx
        self.n = sum(len(p[0]) if columns else len(p) for p in parts)
        self.kv = ctx.kv(max(1, self.n))
        self.err = None
        self.seconds = 0.0

        def work():
            import time as _time
            t0 = _time.perf_counter()
            try:
                off = 0
                for p in parts:
                    if columns:
                        self.kv.upload_columns(off, p[0], p[1])
                        off += len(p[0])
                    else:
                        self.kv.upload(off, p, len(p))
                        off += len(p)
            except BaseException as e:   # re-raised by wait()
                self.err = e
            self.seconds = _time.perf_counter() - t0
        self.th = threading.Thread(target=work, name="dampr-spill-upload")
        self.th.start()

    def wait(self):
        """The filled kv (all copies complete)."""
        self.th.join()
        if self.err is not None:
            self.kv.free()
            raise self.err
        sync = getattr(self.ctx, "sync_copy_stream", None) or self.ctx.sync
        sync()
        return self.kv


def _upload_runs(ctx, runs):
    """One device kv from (n_i, 2) uint64 record runs, uploaded back to back in run order."""
    kv = ctx.kv(max(1, sum(len(r) for r in runs)))
    off = 0
    for r in runs:
        kv.upload(off, r, len(r))
        off += len(r)
    ctx.sync()
    return kv


def _pipelined(ctx, items, columns):
    """(tag, filled kv, upload seconds) for every (tag, parts) of `items` — parts = a list of column chunks
    (columns=True) or of record runs, or None for an item that needs no device copy (kv is None then) — the
    upload of the next item overlapping whatever the consumer does with the current one."""
    it = iter(items)
    cur = next(it, None)
    up = _Upload(ctx, cur[1], columns) if cur is not None and cur[1] is not None else None
    while cur is not None:
        kv = up.wait() if up is not None else None
        sec = up.seconds if up is not None else 0.0
        nxt = next(it, None)
        try:
            up = _Upload(ctx, nxt[1], columns) if nxt is not None and nxt[1] is not None else None
        except BaseException:
            if kv is not None:
                kv.free()
            raise
        try:
            yield cur[0], kv, sec
        except BaseException:
            if up is not None:
                try:
                    up.wait().free()
                except BaseException:
                    pass
            raise
        cur = nxt


_HOST_ARENA = {"buf": None, "leased": False, "chunks": [], "reg": 0, "thread": None}


def _register_prefix(buf, upto):
    """Page-lock buf[reg:upto] (background thread, after the job that touched those pages) as ONE region: a copy
    may not straddle two registrations (the library sends such a copy through its staging ring instead)."""
    done = _HOST_ARENA["reg"]
    if done < upto and _HOST_ARENA["buf"] is buf:
        if dev.host_register(buf[done:upto]):
            _HOST_ARENA["chunks"].append((done, upto))
            _HOST_ARENA["reg"] = upto


def _drop_cached_arena():
    th = _HOST_ARENA["thread"]
    buf = _HOST_ARENA["buf"]
    _HOST_ARENA["buf"] = None        # stops the registration loop
    if th is not None:
        th.join()
    if buf is not None:
        for lo, hi in _HOST_ARENA["chunks"]:
            dev.host_unregister(buf[lo:hi])
    _HOST_ARENA.update({"chunks": [], "reg": 0, "thread": None})


class _RunArena(object):
    """Host memory for the sorted runs one job spills (pass 1 writes them back to back, pass 2 reads them). The
    buffer is leased from a process-wide cache (settings.host_spill_cache_bytes): a second job finds its pages
    already mapped — no page faults while the downloads land, no unmapping of tens of GB at the end — and, once
    the background registration that starts when the first job ends is through, PAGE-LOCKED: the run downloads
    of pass 1 and the run uploads of pass 2 then go straight over PCIe, without the staging ring and its host
    copies (registering touched pages runs at ~15 GB/s here, untouched ones at ~4 GB/s, which is why the first
    job does not wait for it). A nested lease (external_sort splitting a skewed range of runs that live in the
    arena) gets a buffer of its own."""

    def __init__(self, n_records, ctx=None):
        n_records = max(1, int(n_records))
        self.cached = False
        self.real_device = ctx is not None and hasattr(ctx, "lib")
        if not _HOST_ARENA["leased"]:
            buf = _HOST_ARENA["buf"]
            if buf is None or len(buf) < n_records:
                _drop_cached_arena()       # (unregisters) before the new one is mapped
                buf = np.empty((n_records, 2), dtype=np.uint64)
                if buf.nbytes <= int(settings.host_spill_cache_bytes or 0):
                    _HOST_ARENA["buf"] = buf
            elif _HOST_ARENA["thread"] is not None:
                _HOST_ARENA["thread"].join()
                _HOST_ARENA["thread"] = None
            self.cached = _HOST_ARENA["buf"] is buf
            _HOST_ARENA["leased"] = self.cached
        else:
            buf = np.empty((n_records, 2), dtype=np.uint64)
        self.buf, self.pos = buf, 0
        self.reg = _HOST_ARENA["reg"] if self.cached else 0

    def take(self, kv):
        """Download a device kv as the next run; returns the (n, 2) view."""
        n = len(kv)
        if self.pos + n > len(self.buf):
            raise RuntimeError("spill arena overflow: %d + %d > %d" % (self.pos, n, len(self.buf)))
        out = self.buf[self.pos:self.pos + n]
        if n:
            cut = self.reg - self.pos
            if 0 < cut < n:    # the run straddles the end of the page-locked prefix: one copy per kind of memory
                kv.records_into(out[:cut], 0)
                kv.records_into(out[cut:], cut)
            else:
                kv.records_into(out)
        self.pos += n
        return out

    def release(self):
        if self.cached:
            _HOST_ARENA["leased"] = False
            if self.real_device and self.pos > _HOST_ARENA["reg"] and _HOST_ARENA["buf"] is self.buf:
                import threading
                th = threading.Thread(target=_register_prefix, args=(self.buf, self.pos), name="dampr-spill-register")
                _HOST_ARENA["thread"] = th
                th.start()
        self.buf = None


def release_host_arena():
    """Return the cached run buffer to the OS."""
    if not _HOST_ARENA["leased"]:
        _drop_cached_arena()


import atexit as _atexit  # noqa: E402

_atexit.register(release_host_arena)


class _OutCols(object):
    """Result columns of a spill pipeline, written in place: one allocation sized for the worst case (untouched
    pages cost nothing), filled piece by piece straight from the device downloads — no concatenation at the end."""

    def __init__(self, cap):
        self.k = np.empty(max(1, cap), dtype=np.uint64)
        self.v = np.empty(max(1, cap), dtype=np.uint64)
        self.n = 0

    def take_kv(self, kv):
        g = len(kv)
        if g:
            kv.columns_into(self.k[self.n:self.n + g], self.v[self.n:self.n + g])
            self.n += g

    def take_records(self, recs):
        g = len(recs)
        self.k[self.n:self.n + g] = recs[:, 0]
        self.v[self.n:self.n + g] = recs[:, 1]
        self.n += g

    def pieces(self):
        return [(self.k[:self.n], self.v[:self.n])]


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
            + PartialReduceCombiner (base.py:393-402) — instead of being sorted again.
    In both passes the upload of the next batch / bucket runs on a host thread next to the device work and the
    download of the current one (_pipelined): PCIe carries both directions at once."""
    import time as _time
    arena = arena_bytes(ctx)
    per_batch = max(1 << 16, arena // RECORD_FOOTPRINT_OVERLAPPED)
    n_buckets = max(2, int(np.ceil(1.3 * n_records / float(per_batch))))
    mixed = xform == dev.KEY_MIX
    if not mixed:
        raise ValueError("external_group buckets by ranges of the mixed key; use external_sort for ordered keys")
    # bucket b holds mix64(key) in [b * 2^64 / P, (b + 1) * 2^64 / P)
    splitters = np.array([(b << 64) // n_buckets for b in range(1, n_buckets)], dtype=np.uint64)
    buckets = [[] for _ in range(n_buckets)]
    stats = {"buckets": n_buckets, "batches": 0, "spilled_bytes": 0, "arena_bytes": arena,
             "seconds": {"upload_thread": 0.0, "sort_download": 0.0, "cut": 0.0, "pass1_wall": 0.0,
                         "merge_upload_thread": 0.0, "merge_download": 0.0, "pass2_wall": 0.0}}
    sec = stats["seconds"]
    op2 = dev.OP_SUM_I64 if op == dev.OP_COUNT else op   # partial counts are added up
    t_pass = _time.perf_counter()
    spilled = 0
    arena_host = _RunArena(n_records, ctx)
    try:
        return _external_group_passes(ctx, chunk_iter, per_batch, n_buckets, splitters, buckets, stats, op, op2, xform,
                                      arena_host)
    finally:
        arena_host.release()


def _external_group_passes(ctx, chunk_iter, per_batch, n_buckets, splitters, buckets, stats, op, op2, xform, arena_host):
    import time as _time
    sec = stats["seconds"]
    t_pass = _time.perf_counter()
    spilled = 0
    for _tag, kv, up_s in _pipelined(ctx, ((None, b) for b in _batches(chunk_iter, per_batch)), True):
        sec["upload_thread"] += up_s
        t1 = _time.perf_counter()
        try:
            if op is None:
                kv.sort(xform)
                recs = arena_host.take(kv)
            else:
                red = kv.sort_reduce(op, xform, sorted_run=True)
                try:
                    recs = arena_host.take(red)
                finally:
                    red.free()
        finally:
            kv.free()
        t3 = _time.perf_counter()
        sec["sort_download"] += t3 - t1
        cuts = _lazy_cuts(recs[:, 0], splitters, _mix64)
        sec["cut"] += _time.perf_counter() - t3
        edges = [0] + cuts.tolist() + [len(recs)]
        for b in range(n_buckets):
            if edges[b + 1] > edges[b]:
                buckets[b].append(recs[edges[b]:edges[b + 1]])
        stats["batches"] += 1
        stats["spilled_bytes"] += recs.nbytes
        spilled += len(recs)
    sec["pass1_wall"] = _time.perf_counter() - t_pass

    t_pass = _time.perf_counter()
    out = _OutCols(spilled)
    _merge_buckets(ctx, buckets, xform, -1 if op is None else op2, per_batch, out, sec)
    sec["pass2_wall"] = _time.perf_counter() - t_pass
    return out.pieces(), stats


def _merge_buckets(ctx, buckets, xform, op, per_batch, out, sec=None):
    """Pass 2: the runs of every bucket, in bucket order, merged (+ folded, op >= 0) on the device into `out`.
    Buckets whose runs fit the arena together are pipelined (the next bucket's runs upload while this one is
    merged and downloaded); a bucket of one run is already final; an oversized (skewed) bucket is cut into key
    range slices (_merge_runs)."""
    import time as _time

    def items():
        for b in range(len(buckets)):
            runs, buckets[b] = buckets[b], None
            if not runs:
                continue
            total = sum(len(r) for r in runs)
            if len(runs) > 1 and total <= per_batch:
                yield ("merge", runs), runs
            else:
                yield ("host", runs), None

    for (kind, runs), kv, up_s in _pipelined(ctx, items(), False):
        if kind == "merge":
            t1 = _time.perf_counter()
            try:
                offs = np.concatenate(([0], np.cumsum([len(r) for r in runs]))).astype(np.uint64)
                m = ctx.kv_merge_ranges(kv, offs, xform, op)
            finally:
                kv.free()
            try:
                out.take_kv(m)
            finally:
                m.free()
            if sec is not None:
                sec["merge_upload_thread"] += up_s
                sec["merge_download"] += _time.perf_counter() - t1
        elif len(runs) == 1:   # one sorted (folded) run: nothing to merge
            out.take_records(runs[0])
        else:
            for k, v in _merge_runs(ctx, runs, xform, op, per_batch):
                _take_cols(out, k, v)


def _take_cols(out, k, v):
    g = len(k)
    out.k[out.n:out.n + g] = k
    out.v[out.n:out.n + g] = np.asarray(v).view(np.uint64)
    out.n += g


def _merge_runs(ctx, runs, xform, op, per_batch, sec=None):
    """[(keys, vals)] — the k-way merge (+ fold, op >= 0) of sorted host runs on the device. Runs that do not
    fit the arena together (a skewed bucket) are merged in key-range slices: the runs are cut at sampled
    splitters so that every slice fits, and the slices come out in key order."""
    total = sum(len(r) for r in runs)
    if total <= per_batch:
        kv = _upload_runs(ctx, runs)
        try:
            offs = np.concatenate(([0], np.cumsum([len(r) for r in runs]))).astype(np.uint64)
            m = ctx.kv_merge_ranges(kv, offs, xform, op)
        finally:
            kv.free()
        try:
            k, v = m.columns()
        finally:
            m.free()
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
    per_batch = max(1 << 16, arena // RECORD_FOOTPRINT_OVERLAPPED)
    n_buckets = max(2, int(np.ceil(1.3 * n_records / float(per_batch))))
    samp = np.unique(_order_domain(np.asarray(sample_keys).view(np.uint64), xform))
    # a splitter is never the smallest sampled key, so the range below the first splitter is not empty
    q = np.maximum(1, (np.arange(1, n_buckets) * len(samp)) // n_buckets)
    splitters = np.unique(samp[q]) if len(samp) > 1 else np.zeros(0, dtype=np.uint64)
    nb = len(splitters) + 1
    buckets = [[] for _ in range(nb)]
    sizes = [0] * nb
    stats = {"buckets": nb, "batches": 0, "spilled_bytes": 0, "arena_bytes": arena}
    arena_host = _RunArena(n_records, ctx)
    try:
        return _external_sort_passes(ctx, chunk_iter, per_batch, nb, splitters, buckets, sizes, stats, xform, _depth,
                                     arena_host)
    finally:
        arena_host.release()


def _external_sort_passes(ctx, chunk_iter, per_batch, nb, splitters, buckets, sizes, stats, xform, _depth, arena_host):
    for _tag, kv, _up_s in _pipelined(ctx, ((None, bt) for bt in _batches(chunk_iter, per_batch)), True):
        try:
            kv.sort(xform)
            recs = arena_host.take(kv)
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
