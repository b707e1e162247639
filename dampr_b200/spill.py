"""Out-of-core shuffle for kv records that do not fit the device arena (BASELINE config 4).

Reference behaviour replaced: MaxMemoryWriter / CSDatasetWriter spill a sorted run per partition
whenever the worker's RSS grew by settings.max_memory_per_worker (dataset.py:190-262,
memory.py:72-113), then ReduceStageRunner merges the runs of one partition at a time
(stagerunner.py:269-282, dataset.py:571-579).

Here the spill trigger is the device arena (settings.device_arena_bytes, default 70 % of HBM):
  pass 1  stream the input in arena-sized batches; each batch is partitioned ON THE DEVICE into P
          host buckets (owner = mix64(key) % P, dampr_kv_partition_by_owner: destination-contiguous),
          and every bucket slice is copied device -> host into that bucket's spill list
          (the reference's "one run per partition per spill");
  pass 2  one bucket at a time: host -> device, partition + sort (+ segmented reduce) on the
          device, results streamed back.  A key lives in exactly one bucket, so buckets are
          independent, exactly like the reference's reduce partitions.
PCIe traffic: 2 x 16 B per record each way; the device never holds more than one batch / bucket.
Grouping order (mixed-key order inside a bucket, buckets in owner order) is unobservable in results
(SURVEY "Result-order contract").

sort_by needs a globally ordered result, so external_sort buckets by KEY RANGE instead: splitters are
quantiles of a key sample, every batch is sorted on the device and cut at the splitters into sorted
runs (the reference's sorted run per spill, dataset.py:162-164), and each range is then sorted once
more on the device from its runs in batch order -- the device sort is stable, so equal keys keep their
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


def external_group(ctx, chunk_iter, n_records, op=None, xform=dev.KEY_MIX):
    """chunk_iter yields (keys uint64[], vals 8-byte[]) column chunks. Returns a list of
    (keys, vals) numpy result pieces: key-sorted (under `xform`) inside each piece; with `op` one
    record per key. Also returns stats {"buckets", "batches", "spilled_bytes"}."""
    arena = arena_bytes(ctx)
    per_batch = max(1 << 16, arena // RECORD_FOOTPRINT)
    n_buckets = max(2, int(np.ceil(1.3 * n_records / float(per_batch))))
    buckets = [[] for _ in range(n_buckets)]
    stats = {"buckets": n_buckets, "batches": 0, "spilled_bytes": 0, "arena_bytes": arena}

    def flush_batch(kchunks, vchunks):
        kv = _upload_chunks(ctx, kchunks, vchunks)
        try:
            parts, counts = kv.partition_by_owner(n_buckets)
        finally:
            kv.free()
        try:
            recs = parts.records()  # (n, 2) uint64, bucket-contiguous
        finally:
            parts.free()
        off = 0
        for b, c in enumerate(counts.tolist()):
            if c:
                buckets[b].append(recs[off:off + c])
                off += c
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
        if not buckets[b]:
            continue
        kv = _upload_runs(ctx, buckets[b])
        buckets[b] = None
        try:
            if op is None:
                kv.sort(xform)
                k, v = kv.columns()
            else:
                red = kv.sort_reduce(op, xform)
                try:
                    k, v = red.columns()
                finally:
                    red.free()
        finally:
            kv.free()
        out.append((k, v))
    return out, stats


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
        cuts = np.searchsorted(_order_domain(recs[:, 0], xform), splitters, side="left")
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
        kv = _upload_runs(ctx, runs)
        try:
            kv.sort(xform)
            k, v = kv.columns()
        finally:
            kv.free()
        out.append((k, v))
    return out, stats
