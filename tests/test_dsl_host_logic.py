"""The DSL cases of tests/test_gpu_dsl.py (the reference suite's literals and the engine's own cases) run
once more on CPU against tests/fake_device.py: what is exercised here is everything ABOVE the C-ABI — graph
building, stage fusion, host maps, key codecs, the shuffle plumbing of the runner, joins, sinks, the frame
lowering — with numpy standing in for the kernels. Text stages have no stand-in tokeniser and run as host
maps. The GPU versions of the same cases (-m gpu) are the parity tests proper."""
import inspect

import pytest

import test_gpu_dsl as G
from fake_device import FakeCtx

from dampr_b200 import Dampr, settings
from dampr_b200 import plan
from dampr_b200 import runner as runner_mod

CASES = sorted(name for name, f in vars(G).items() if name.startswith("test_") and callable(f))


@pytest.mark.parametrize("name", CASES)
def test_dsl_case_on_the_numpy_device(name, monkeypatch, tmp_path):
    fake = FakeCtx()
    monkeypatch.setattr(runner_mod, "_CTX", {settings.device: fake})
    monkeypatch.setattr(plan, "_BUFFERS", {})
    fn = getattr(G, name)
    kwargs = {}
    for arg in inspect.signature(fn).parameters:
        if arg == "items":
            kwargs[arg] = Dampr.memory(list(range(10, 20)), partitions=2)
        elif arg == "ctx":
            kwargs[arg] = fake
        elif arg == "tmp_path":
            kwargs[arg] = tmp_path
        else:
            pytest.skip("fixture %s is not available on CPU" % arg)
    fn(**kwargs)
