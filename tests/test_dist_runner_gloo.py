"""CPU: the runner under torch.distributed (gloo, world_size 2) with the numpy stand-in for the device
(tests/fake_device.py): kv folds exchanged by key owner and finished with the run merge, len() of an
owner-partitioned result (summed over the ranks, then REPLICATED: the host fold that follows it is allowed and
its sink is written by rank 0 alone), cross_right against that replicated total, rank-numbered sink parts, and
the loud refusal of a host stage that would need records from another rank."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, math
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, os.path.join(%(root)r, "tests"))
    import numpy as np, torch, torch.distributed as tdist
    from fake_device import FakeCtx
    from dampr_b200 import Dampr, settings, dist as D
    from dampr_b200 import runner as runner_mod
    from dampr_b200.inputs import ArrayKVInput
    from oracle import gen, refsem

    tdist.init_process_group("gloo")
    rank, world = D.world()
    runner_mod._CTX = {settings.device: FakeCtx()}
    out_root = os.environ["DAMPR_TEST_OUT"]

    def host_shuffle(ctx, kv, header=None):
        # dist.shuffle_kv with host buffers (the product does all of this in one C-ABI call over NCCL)
        parts, counts = kv.partition_by_owner(world)
        recv_counts = D.exchange_counts(counts)
        total = int(recv_counts.sum())
        send = torch.from_numpy(np.ascontiguousarray(parts.records()).view(np.uint8).reshape(-1))
        recv = torch.empty(total * 16, dtype=torch.uint8)
        D.all_to_all_bytes(send, counts, recv, recv_counts)
        out = ctx.kv_from_records(recv.numpy().view(np.uint64).reshape(-1, 2))
        heads = np.asarray(D.all_gather_objects(list(header)), dtype=np.int64) if header is not None else None
        return out, np.concatenate(([0], np.cumsum(recv_counts))).astype(np.uint64), heads
    D.shuffle_kv = host_shuffle

    def gathered(mine):
        parts = [None] * world
        tdist.all_gather_object(parts, list(mine))
        merged = {}
        for p in parts:
            for k, v in p:
                assert k not in merged, "key %%r owned by two ranks" %% (k,)
                merged[k] = v
        return merged, [len(p) for p in parts]

    keys, vals = gen.kv(77, 200000, 9000)
    src = Dampr.read_input(ArrayKVInput(keys, vals))
    sums = src.a_group_by(lambda x: x[0], lambda x: x[1]).sum()
    merged, sizes = gathered(sums.read())
    assert merged == refsem.group_sum(keys, vals)
    assert all(s > 0 for s in sizes)
    assert any("owner fold" in how for _s, how, _d in runner_mod.LAST_STATS.stages)
    for kind, exp in (("min", refsem.group_min(keys, vals) if hasattr(refsem, "group_min") else None),):
        if exp is not None:
            m, _ = gathered(src.a_group_by(lambda x: x[0], lambda x: x[1]).reduce(min).read())
            assert m == exp

    # len() of the owner-partitioned result: the global number of groups on EVERY rank
    total = sums.len().read()
    assert total == [len(merged)], (rank, total)

    # cross_right against the replicated total + rank-numbered sink parts
    out_dir = os.path.join(out_root, "sink")
    sums.cross_right(sums.len(), lambda kv, n: (kv[0], kv[1], n), memory=True).sink_tsv(out_dir).run()
    tdist.barrier()
    lines = []
    for fn in sorted(os.listdir(out_dir)):
        with open(os.path.join(out_dir, fn)) as f:
            lines.extend(l.rstrip("\\n") for l in f)
    exp_lines = sorted("%%s\\t%%s\\t%%s" %% (k, v, len(merged)) for k, v in merged.items())
    assert sorted(lines) == exp_lines, (len(lines), len(exp_lines))

    # a replicated result is written once (rank 0), not once per rank
    len_dir = os.path.join(out_root, "len")
    sums.len().sink(len_dir).run()
    tdist.barrier()
    got = []
    for fn in sorted(os.listdir(len_dir)):
        with open(os.path.join(len_dir, fn)) as f:
            got.extend(l.strip() for l in f if l.strip())
    assert got == [str(len(merged))], got

    # reduce-side joins: both sides exchanged by key owner, folds joined on the owner; product with a unique right
    import itertools
    lk, lv = gen.kv(1, 30000, 2000)
    rk, rv = gen.kv(2, 4000, 3000)
    lk, rk = lk.view(np.int64), rk.view(np.int64)
    G = lambda ks, vs: Dampr.read_input(ArrayKVInput(ks, vs)).group_by(lambda x: x[0], lambda x: x[1])
    inner = refsem.inner_join(lk, lv, rk, rv)
    left = refsem.left_join(lk, lv, rk, rv)
    got, _ = gathered(G(lk, lv).join(G(rk, rv)).reduce(lambda l, r: (sum(l), len(list(r)))).read())
    assert any("device join" in how and "all-to-all" in how for _s, how, _d in runner_mod.LAST_STATS.stages), runner_mod.LAST_STATS.stages
    assert got == {k: (sum(a), len(b)) for k, (a, b) in inner.items()}
    got, _ = gathered(G(lk, lv).join(G(rk, rv)).left_reduce(lambda l, r: (sum(l), sum(r))).read())
    assert got == {k: (sum(a), sum(b)) for k, (a, b) in left.items()}
    uk, first = np.unique(rk, return_index=True)
    uv = rv[first]
    rows = G(lk, lv).join(G(uk, uv)).reduce(lambda l, r: itertools.product(l, r), many=True).read()
    assert any("device join" in how and "exchanged" in how for _s, how, _d in runner_mod.LAST_STATS.stages), \
        runner_mod.LAST_STATS.stages
    allrows = [None] * world
    tdist.all_gather_object(allrows, list(rows))
    table = dict(zip(uk.tolist(), uv.tolist()))
    exp = sorted((int(k), (int(v), table[int(k)])) for k, v in zip(lk.tolist(), lv.tolist()) if int(k) in table)
    assert sorted(x for p in allrows for x in p) == exp

    # a host stage that needs the whole input refuses loudly instead of folding this rank's view
    try:
        Dampr.memory(list(range(100))).group_by(lambda x: x %% 7).reduce(lambda k, it: sorted(it)).read()
    except runner_mod.DistributedUnsupported:
        pass
    else:
        raise AssertionError("host reduce ran under torch.distributed")
    tdist.barrier()
    tdist.destroy_process_group()
    open(os.path.join(out_root, "rank%%d.ok" %% rank), "w").write("ok")
""")


def test_gloo_world2_runner(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, DAMPR_TEST_OUT=str(tmp_path))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-6000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()
