"""GPU parity at world size > 1: torchrun over min(2, device_count) GPUs runs tools/mgpu_check.py (document
frequencies, the whole tf-idf job with its sink, str.split word counts, a kv fold: every result gathered and
compared with the oracle) and a short bench.py whose own parity check must pass. Skipped on one GPU."""
import json
import os
import socket
import subprocess
import sys

import pytest

from dampr_b200 import device as dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _torchrun(n, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), script] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def _world():
    n = dev.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (found %d)" % n)
    return 2


def test_mgpu_check_world2():
    n = _world()
    r = _torchrun(n, os.path.join(ROOT, "tools", "mgpu_check.py"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    assert "mgpu_check ok: world=%d" % n in r.stdout


def test_bench_parity_world2():
    n = _world()
    r = _torchrun(n, os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--steps", "1", "--warmup", "1", "--gb", "0.5",
                  "--no-cpu-baseline")
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == n and line["parity"]["equal"] is True and line["parity"]["ranks"] == n
    assert line["parity"]["full_size"]["host_vs_resident_equal"] is True
