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


def external_group(ctx, chunk_iter, n_records, op=None, xform=dev.KEY_MIX):
    """chunk_iter yields (keys uint64[], vals 8-byte[]) column chunks. Returns a list of
    (keys, vals) numpy result pieces: key-sorted (under `xform`) inside each piece; with `op` one
    record per key. Also returns stats {"buckets", "batches", "spilled_bytes"}."""
    arena = arena_bytes(ctx)
    per_batch = max(1 << 16, arena // RECORD_FOOTPRINT)
    n_buckets = max(2, int(np.ceil(1.3 * n_records / float(per_batch))))
    buckets = [[] for _ in range(n_buckets)]
    stats = {"buckets": n_buckets, "batches": 0, "spilled_bytes": 0, "arena_bytes": arena}

    def flush_batch(keys, vals):
        kv = ctx.kv_from_columns(keys, vals)
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
                flush_batch(np.concatenate(pend_k), np.concatenate(pend_v))
                pend_k, pend_v, pend_n = [], [], 0
    if pend_n:
        flush_batch(np.concatenate(pend_k), np.concatenate(pend_v))

    out = []
    for b in range(n_buckets):
        if not buckets[b]:
            continue
        recs = np.concatenate(buckets[b]) if len(buckets[b]) > 1 else buckets[b][0]
        buckets[b] = None
        kv = ctx.kv_from_records(recs)
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
