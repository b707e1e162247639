"""Multi-GPU parity check (run under torchrun on a GPU box):
   python -m torch.distributed.run --nproc-per-node 2 tools/mgpu_check.py
Every rank runs the same DSL script; text files are sharded by byte range, kv inputs by record range,
results live on their owner rank and are gathered here only to compare with the oracle."""
import os
import re
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from dampr_b200 import Dampr, settings
from dampr_b200 import runner as runner_mod
from dampr_b200.inputs import ArrayKVInput
from oracle import gen, refsem

RX = re.compile(r"[^\w]+")


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    settings.device = local
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    data = gen.text(4321, 60000, V=20000) + gen.dirty_text(5, 2000, 1500)
    path = os.path.join(tempfile.gettempdir(), "dampr_mgpu_corpus.txt")
    if rank == 0:
        with open(path, "wb") as f:
            f.write(data)
    dist.barrier()
    # tf-idf document frequencies: every rank owns a slice of the terms
    docs = Dampr.text(path)
    mine = docs.flat_map(lambda x: set(RX.split(x.lower()))).count().read()
    assert any("device text tokenise+combine" in how for _s, how, _d in runner_mod.LAST_STATS.stages)
    n_lines = docs.len().read()[0]
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    merged = {}
    for p in parts:
        for k, v in p:
            assert k not in merged, "term %r owned by two ranks" % (k,)
            merged[k] = v
    exp, exp_lines = refsem.docfreq(data)
    assert merged == dict(exp), "document frequencies differ from the oracle"
    assert n_lines == exp_lines
    # the whole tf-idf job: every rank sinks the terms it owns into its own part files
    import math
    import shutil
    out_dir = os.path.join(tempfile.gettempdir(), "dampr_mgpu_idf")
    if rank == 0:
        shutil.rmtree(out_dir, ignore_errors=True)
    dist.barrier()
    doc_freq = docs.flat_map(lambda x: set(RX.split(x.lower()))).count(reduce_buffer=float("inf"))
    doc_freq.cross_right(docs.len(), lambda df, total: (df[0], df[1], math.log(1 + (float(total) / df[1]))),
                         memory=True).sink_tsv(out_dir).run()
    dist.barrier()
    if rank == 0:
        lines = []
        for fn in sorted(os.listdir(out_dir)):
            with open(os.path.join(out_dir, fn)) as f:
                lines.extend(l.rstrip("\n") for l in f)
        assert sorted(lines) == sorted(refsem.tfidf_sink_lines(data)), "tf-idf sink lines differ from the oracle"
    # word count (str.split, long tokens -> hashed codes + cross-rank string exchange)
    mine = docs.flat_map(lambda x: x.split()).count().read()
    dist.all_gather_object(parts, mine)
    merged = {}
    for p in parts:
        for k, v in p:
            assert k not in merged
            merged[k] = v
    assert merged == dict(refsem.wc_counts(data)), "word counts differ from the oracle"
    # kv fold
    keys, vals = gen.kv(77, 1_000_000, 90_000)
    mine = Dampr.read_input(ArrayKVInput(keys, vals)).a_group_by(lambda x: x[0], lambda x: x[1]).sum().read()
    dist.all_gather_object(parts, mine)
    merged = {}
    for p in parts:
        for k, v in p:
            assert k not in merged
            merged[k] = v
    assert merged == refsem.group_sum(keys, vals), "kv sums differ from the oracle"
    # reduce-side joins (BASELINE config 5): both sides exchanged by the owner of the key (one Splitter for both
    # sides, base.py:264-283), per-side folds joined on the owner; product against a unique right side
    import itertools
    lk, lv = gen.kv(1, 300_000, 20_000)
    rk, rv = gen.kv(2, 40_000, 30_000)
    lk, rk = lk.view(np.int64), rk.view(np.int64)
    G = lambda ks, vs: Dampr.read_input(ArrayKVInput(ks, vs)).group_by(lambda x: x[0], lambda x: x[1])
    inner = refsem.inner_join(lk, lv, rk, rv)
    left = refsem.left_join(lk, lv, rk, rv)

    def gathered(mine):
        dist.all_gather_object(parts, list(mine))
        merged = {}
        for p in parts:
            for k, v in p:
                assert k not in merged, "join key %r on two ranks" % (k,)
                merged[k] = v
        return merged
    got = gathered(G(lk, lv).join(G(rk, rv)).reduce(lambda l, r: (sum(l), len(list(r)))).read())
    assert any("device join" in how and "all-to-all" in how for _s, how, _d in runner_mod.LAST_STATS.stages), \
        runner_mod.LAST_STATS.stages
    assert got == {k: (sum(a), len(b)) for k, (a, b) in inner.items()}, "inner join folds differ from the oracle"
    got = gathered(G(lk, lv).join(G(rk, rv)).left_reduce(lambda l, r: (sum(l), sum(r))).read())
    assert got == {k: (sum(a), sum(b)) for k, (a, b) in left.items()}, "left join folds differ from the oracle"
    uk, first = np.unique(rk, return_index=True)
    uv = rv[first]
    rows = G(lk, lv).join(G(uk, uv)).reduce(lambda l, r: itertools.product(l, r), many=True).read()
    assert any("device join" in how and "exchanged" in how for _s, how, _d in runner_mod.LAST_STATS.stages), \
        runner_mod.LAST_STATS.stages
    dist.all_gather_object(parts, list(rows))
    table = dict(zip(uk.tolist(), uv.tolist()))
    exp = sorted((int(k), (int(v), table[int(k)])) for k, v in zip(lk.tolist(), lv.tolist()) if int(k) in table)
    assert sorted(x for p in parts for x in p) == exp, "product join rows differ from the oracle"
    dist.barrier()
    if rank == 0:
        print("mgpu_check ok: world=%d terms=%d lines=%d" % (world, len(exp), exp_lines))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
