"""CPU port of the reference's runner for the hot path — TEST / BASELINE INFRASTRUCTURE ONLY.

A compact multi-process restatement of how Refefer/Dampr executes the TF-IDF and word-count
graphs, used (a) as the `cpu_baseline` / `--impl reference` arm of bench.py on
the GPU box, where /root/reference does not exist, and (b) as a second oracle.  It follows the
reference's algorithm stage by stage, in pure Python like the reference:

  map      forked workers pull byte-range chunks from a queue (StageRunner.run, stagerunner.py:15-43),
           read lines with the chunk-ownership rule (TextLineDataset.read, dataset.py:458-476), apply
           the user's tokeniser per line, fold (token, 1) into a dict combiner
           (ReducedWriter.add_record, dataset.py:100-105), then sort the dict by key
           (SortedWriter, dataset.py:162-164), hash-partition it into 91 partitions
           (DefaultShuffler.shuffle, base.py:416-433) and ship each partition as a gzip'd pickle blob
           (MemGZipDataset, dataset.py:528-547);
  reduce   one job per partition: heapq.merge of the workers' sorted runs (MergeDataset.read,
           dataset.py:571-579), groupby, left fold (ARReduce._reduce, dampr.py:678-683);
  finish   IDF per term and tab-separated sink lines (tf-idf-dampr.py:17-21, dampr.py:521-529).

It is NOT the reference and is a little leaner (no nested generator per fused map, no RSS polling),
so it errs on the side of a faster CPU baseline.  DESIGN.md records both timed side by side in the
dev container.
"""
import gzip
import heapq
import io
import itertools
import math
import multiprocessing as mp
import os
import pickle
import re
import time
from operator import itemgetter

N_PARTITIONS = 91           # settings.py:11
BATCH = 1000                # settings.py:20
RX = re.compile(r"[^\w]+")  # tf-idf-dampr.py:11


def _lines(path, start, end):
    """Lines whose first byte lies in [start, end) (every line exactly once)."""
    with open(path, "rb") as f:
        if start > 0:
            f.seek(start - 1)
            if f.read(1) != b"\n":
                f.readline()
        pos = f.tell() if start > 0 else 0
        if start == 0:
            f.seek(0)
        while pos < end:
            line = f.readline()
            if not line:
                break
            pos += len(line)
            yield line.decode("utf-8").rstrip("\n")


def _dump_partitions(items):
    """sorted (k, v) list -> {partition: gzip'd pickle blob} like DefaultShuffler + MemGZipDataset."""
    bufs = [[] for _ in range(N_PARTITIONS)]
    for kv in items:
        bufs[hash(kv[0]) % N_PARTITIONS].append(kv)
    out = {}
    for p, rows in enumerate(bufs):
        if not rows:
            continue
        raw = io.BytesIO()
        with gzip.GzipFile(fileobj=raw, mode="wb", compresslevel=1) as g:
            for i in range(0, len(rows), BATCH):
                pickle.dump(rows[i:i + BATCH], g, pickle.HIGHEST_PROTOCOL)
        out[p] = raw.getvalue()
    return out


def _load_blob(blob):
    with gzip.GzipFile(fileobj=io.BytesIO(blob)) as g:
        try:
            while True:
                for kv in pickle.load(g):
                    yield kv
        except EOFError:
            return


def _map_worker(path, mode, in_q, out_q):
    combiner = {}
    n_lines = 0
    get = combiner.get
    while True:
        job = in_q.get()
        if job is None:
            break
        start, end = job
        for line in _lines(path, start, end):
            n_lines += 1
            toks = set(RX.split(line.lower())) if mode == "tfidf" else line.split()
            for t in toks:
                c = get(t)
                combiner[t] = 1 if c is None else 1 + c
    items = sorted(combiner.items(), key=itemgetter(0))
    out_q.put((n_lines, _dump_partitions(items)))


def _reduce_worker(in_q, out_q):
    out = []
    while True:
        job = in_q.get()
        if job is None:
            break
        _p, blobs = job
        runs = [_load_blob(b) for b in blobs]
        merged = heapq.merge(*runs, key=itemgetter(0)) if len(runs) > 1 else runs[0]
        for k, grp in itertools.groupby(merged, key=itemgetter(0)):
            vs = (kv[1] for kv in grp)
            acc = next(vs)
            for v in vs:
                acc = acc + v
            out.append((k, acc))
    out_q.put(out)


def _run_pool(target, args, jobs, n_procs):
    in_q, out_q = mp.Queue(), mp.Queue()
    for j in jobs:
        in_q.put(j)
    procs = []
    for _ in range(n_procs):
        p = mp.Process(target=target, args=args + (in_q, out_q))
        p.start()
        in_q.put(None)
        procs.append(p)
    results = [out_q.get() for _ in range(n_procs)]
    for p in procs:
        p.join()
    return results


def count_tokens(path, mode="tfidf", n_procs=None, chunk_size=None):
    """[(token, count)], n_lines — map/shuffle/reduce as the reference runs `.count()` on text."""
    n_procs = n_procs or os.cpu_count()
    size = os.path.getsize(path)
    if chunk_size is None:
        chunk_size = size // n_procs + 1 if mode == "tfidf" else 16 * 1024 ** 2
    jobs = [(s, min(size, s + chunk_size)) for s in range(0, size, chunk_size)]
    ctx_results = _run_pool(_map_worker, (path, mode), jobs, n_procs)
    n_lines = sum(r[0] for r in ctx_results)
    by_part = {}
    for _n, parts in ctx_results:
        for p, blob in parts.items():
            by_part.setdefault(p, []).append(blob)
    red = _run_pool(_reduce_worker, (), sorted(by_part.items()), n_procs)
    rows = [kv for part in red for kv in part]
    return rows, n_lines


def tfidf(path, out_dir, n_procs=None):
    """benchmarks/tf-idf-dampr.py end to end: returns (n_terms, n_lines)."""
    rows, total = count_tokens(path, "tfidf", n_procs)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "part-0"), "w", encoding="utf-8") as f:
        for w, df in rows:
            print(u"\t".join(str(p) for p in (w, df, math.log(1 + (float(total) / df)))), file=f)
    return len(rows), total


def wc(path, n_procs=None):
    rows, _n = count_tokens(path, "wc", n_procs)
    rows.sort(key=lambda kv: -kv[1])
    return rows


def timed_tfidf(path, out_dir, n_procs=None):
    t0 = time.time()
    n_terms, n_lines = tfidf(path, out_dir, n_procs)
    return time.time() - t0, n_terms, n_lines
