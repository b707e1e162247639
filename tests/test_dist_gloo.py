"""CPU: the N>1 plumbing (dampr_b200/dist.py) with the gloo backend, world_size 2: counts exchange,
variable-size record all-to-all, integer all-reduce, object gather."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from dampr_b200 import dist as D
    dist.init_process_group("gloo")
    rank, n = D.world()
    assert D.active() and n == 2
    rng = np.random.default_rng(100 + rank)
    keys = rng.integers(0, 1000, size=5000).astype(np.uint64)
    vals = np.ones(5000, dtype=np.uint64)
    owner = (keys %% np.uint64(n)).astype(np.int64)
    order = np.argsort(owner, kind="stable")
    recs = np.stack([keys[order], vals[order]], axis=1)
    counts = np.bincount(owner, minlength=n)
    recv_counts = D.exchange_counts(counts)
    send_t = torch.from_numpy(recs.view(np.uint8).reshape(-1))
    recv_t = torch.empty(int(recv_counts.sum()) * 16, dtype=torch.uint8)
    D.all_to_all_bytes(send_t, counts, recv_t, recv_counts)
    got = recv_t.numpy().view(np.uint64).reshape(-1, 2)
    assert (got[:, 0] %% np.uint64(n) == rank).all()
    tot = D.all_reduce_sum_int([len(got), int(got[:, 1].sum())])
    assert tot == [10000, 10000], tot
    objs = D.all_gather_objects({"rank": rank})
    assert [o["rank"] for o in objs] == [0, 1]
    # the union of both ranks' keys, grouped on their owners, has the right totals
    allk = np.concatenate([np.random.default_rng(100 + r).integers(0, 1000, size=5000) for r in range(n)])
    mine = allk[allk %% n == rank]
    u, c = np.unique(mine, return_counts=True)
    gu, gc = np.unique(got[:, 0], return_counts=True)
    assert np.array_equal(u.astype(np.uint64), gu) and np.array_equal(c, gc)
    dist.destroy_process_group()
    open(os.path.join(os.environ["DAMPR_TEST_OUT"], "rank%%d.ok" %% rank), "w").write("ok")
""")


def test_gloo_world2_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    env = dict(os.environ, DAMPR_TEST_OUT=str(tmp_path))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()
