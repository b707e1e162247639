"""Host map over input chunks in forked worker processes.

A map stage whose lambdas cannot be lowered runs as CPython, exactly like the reference's
MapStageRunner (stagerunner.py:54-129: forked workers pull chunk jobs, `Map.stream` applies the user
functions, ReducedWriter folds equal keys in a dict when the stage has a combiner, dataset.py:84-117).
What differs is what happens to the output: the workers hand their (key, value) lists back to the
parent, which feeds them to the device shuffle instead of sorted gzip'd pickle runs.

Order: every worker takes a CONTIGUOUS range of chunks and results are concatenated in worker order, so
the record order equals the sequential one (the engine defines `first()` as first by input offset).
With a combiner the workers fold `binop(acc, v)` left to right per key; the partials are folded again in
order by the caller (associative binops only, as in the reference).
"""
import multiprocessing
import os
import pickle
import traceback
import warnings


def sequential(mapper, chunks, supp):
    keys, vals = [], []
    ka, va = keys.append, vals.append
    for ch in chunks:
        for k, v in mapper.map(ch, *supp):
            ka(k)
            va(v)
    return keys, vals


def _fold(mapper, chunks, supp, binop):
    acc = {}
    for ch in chunks:
        for k, v in mapper.map(ch, *supp):
            if k in acc:
                acc[k] = binop(acc[k], v)
            else:
                acc[k] = v
    return list(acc.keys()), list(acc.values())


def _worker(conn, mapper, chunks, supp, binop):
    code = 0
    try:
        try:
            out = _fold(mapper, chunks, supp, binop) if binop is not None else sequential(mapper, chunks, supp)
            try:
                payload = pickle.dumps(("ok", out), protocol=pickle.HIGHEST_PROTOCOL)
            except Exception as pe:   # records that cannot cross a process boundary: the parent maps in-process
                payload = pickle.dumps(("unpicklable", "%s: %s" % (type(pe).__name__, pe)))
        except BaseException as e:   # the parent re-raises: a failing lambda must not hang the run
            text = "%s: %s\n%s" % (type(e).__name__, e, traceback.format_exc())
            try:
                payload = pickle.dumps(("exc", (e, text)), protocol=pickle.HIGHEST_PROTOCOL)
            except Exception:
                payload = pickle.dumps(("err", text))
            code = 1
        conn.send_bytes(payload)
        conn.close()
    finally:
        os._exit(code)   # never run the parent's atexit handlers (CUDA runtime, thread pools) in a child


def parallel(mapper, chunks, supp, binop=None, processes=None):
    """(keys, vals) of mapper over chunks, mapped by up to `processes` forked workers. binop folds equal
    keys inside each worker (map-side combine); keys must then be hashable, as in the reference."""
    chunks = list(chunks)
    nproc = max(1, min(int(processes or os.cpu_count() or 1), len(chunks)))
    if nproc <= 1:
        return _fold(mapper, chunks, supp, binop) if binop is not None else sequential(mapper, chunks, supp)
    ctx = multiprocessing.get_context("fork")   # lambdas and open datasets are inherited, not pickled
    bounds = [len(chunks) * i // nproc for i in range(nproc + 1)]
    procs = []
    for i in range(nproc):
        recv, send = ctx.Pipe(duplex=False)
        p = ctx.Process(target=_worker, args=(send, mapper, chunks[bounds[i]:bounds[i + 1]], supp, binop))
        p.daemon = True
        with warnings.catch_warnings():
            # CPython 3.12 warns about fork() in a process that has threads (the copy pool, CUDA's own):
            # the children only run the user's map functions and leave through os._exit
            warnings.simplefilter("ignore", DeprecationWarning)
            p.start()
        send.close()
        procs.append((p, recv))
    keys, vals, err, local = [], [], None, False
    for p, recv in procs:
        try:
            tag, out = pickle.loads(recv.recv_bytes())
        except EOFError:
            tag, out = "err", "worker %d died without a result (exit code %s)" % (p.pid, p.exitcode)
        recv.close()
        p.join()
        if tag == "ok" and err is None:
            keys.extend(out[0])
            vals.extend(out[1])
        elif tag == "unpicklable":
            local = True
        elif tag != "ok" and err is None:
            err = out
    if err is not None:
        if isinstance(err, tuple):   # the user's own exception, with the worker's traceback as its cause
            raise err[0] from RuntimeError("raised in a host map worker:\n%s" % err[1])
        raise RuntimeError("host map worker failed: %s" % err)
    if local:   # some records (lambdas, generators, open files ...) only live in one process: map here instead
        return _fold(mapper, chunks, supp, binop) if binop is not None else sequential(mapper, chunks, supp)
    return keys, vals


def parallel_reduce(reducer, ds, processes=None):
    """reducer.reduce over a grouped RecordsDataset (equal keys adjacent), the groups split over forked
    workers at group boundaries — the reference's ReduceStageRunner runs one reduce job per partition in
    its process pool (stagerunner.py:269-282); any split between two groups is such a partitioning.
    Results come back in group order."""
    from .datasets import RecordsDataset
    keys, values = ds.keys, ds.values
    n = len(keys)
    nproc = max(1, min(int(processes or os.cpu_count() or 1), n))
    cuts = [0]
    for i in range(1, nproc):
        t = max(cuts[-1], n * i // nproc)
        while 0 < t < n and keys[t] == keys[t - 1]:
            t += 1
        if t > cuts[-1] and t < n:
            cuts.append(t)
    cuts.append(n)
    if len(cuts) <= 2:
        ks, vs = [], []
        for k, v in reducer.reduce(ds):
            ks.append(k)
            vs.append(v)
        return ks, vs

    class _Sub(object):   # Mapper-like adapter so that _worker / sequential can drive the reducer
        def map(self, part):
            return reducer.reduce(part)

    def sub(a, b):
        codes = ds.codes[a:b] if ds.codes is not None else None
        return RecordsDataset(keys[a:b], values[a:b], codes, ds.codec)
    parts = [sub(a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    return parallel(_Sub(), parts, (), None, len(parts))
