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


import test_gpu_workloads as W

WORKLOADS = ["test_kv_larger_vs_oracle", "test_kv_workloads_match_reference_golden", "test_map_side_join_lowered_to_hash_probe",
             "test_non_lowerable_map_runs_in_forked_workers", "test_non_lowerable_reduce_runs_in_forked_workers",
             "test_non_lowerable_text_falls_back_to_host_map", "test_shim_package_runs_reference_style_script",
             "test_sort_by_over_binary_records", "test_spill_path_with_capped_arena", "test_text_workloads_larger_vs_oracle",
             "test_columnar_join_idioms"]


@pytest.mark.parametrize("name", WORKLOADS)
def test_workload_on_the_numpy_device(name, monkeypatch, tmp_path):
    """The BASELINE-shaped workloads that do not need the device tokeniser: kv folds / sorts / joins / spill on
    the stand-in, text workloads through the host-map path (compared with the oracle like on the GPU)."""
    fake = FakeCtx()
    monkeypatch.setattr(runner_mod, "_CTX", {settings.device: fake})
    monkeypatch.setattr(plan, "_BUFFERS", {})
    fn = getattr(W, name)
    kwargs = {a: (fake if a == "ctx" else tmp_path) for a in inspect.signature(fn).parameters}
    fn(**kwargs)


@pytest.mark.parametrize("fixture,maker", W.test_text_workloads_match_reference_golden.pytestmark[0].args[1],
                         ids=["text_zipf", "text_dirty"])
def test_reference_golden_text_workloads_through_the_host_map_path(fixture, maker, monkeypatch, tmp_path):
    """wc.py / tf-idf-dampr.py on the golden corpora give the REAL reference's recorded outputs also when
    every text stage runs as a host map (the path non-lowerable pipelines take)."""
    fake = FakeCtx()
    monkeypatch.setattr(runner_mod, "_CTX", {settings.device: fake})
    monkeypatch.setattr(plan, "_BUFFERS", {})
    fix = W.load(fixture)
    data = maker()
    p = tmp_path / "corpus.txt"
    p.write_bytes(data)
    rows = W.run_wc(str(p))
    assert sorted(rows) == [tuple(r) for r in fix["wc"]]
    assert [c for _w, c in rows] == sorted((c for _w, c in rows), reverse=True)
    assert W.run_tfidf(str(p), str(tmp_path / "idfs"), len(data) / 8 + 1) == fix["tfidf_lines"]
    assert Dampr.text(str(p)).len().read() == [fix["n_lines"]]


from fake_device import FakeTextCtx, FakePinned

TEXT_WORKLOADS = ["test_text_workloads_larger_vs_oracle", "test_gzip_text_inputs_are_lowered",
                  "test_lines_with_many_distinct_tokens_stay_on_the_device", "test_non_lowerable_text_falls_back_to_host_map",
                  "test_one_cr_line_in_a_large_text_stays_on_the_device"]


def _text_fake(monkeypatch):
    fake = FakeTextCtx()
    monkeypatch.setattr(runner_mod, "_CTX", {settings.device: fake})
    monkeypatch.setattr(plan, "_BUFFERS", {})
    monkeypatch.setattr(plan, "_pinned_ring", lambda nslots, slot_bytes: [FakePinned(slot_bytes) for _ in range(nslots)])
    return fake


@pytest.mark.parametrize("name", TEXT_WORKLOADS)
def test_text_workload_through_the_scan_plumbing(name, monkeypatch, tmp_path):
    """plan.TextScan (files laid out in one buffer, chunked uploads, owner ranges, flags, fallbacks, the result
    frame) with a stand-in tokeniser: the lowered text path minus the kernel."""
    fake = _text_fake(monkeypatch)
    fn = getattr(W, name)
    fn(**{a: (fake if a == "ctx" else tmp_path) for a in inspect.signature(fn).parameters})


@pytest.mark.parametrize("fixture,maker", W.test_text_workloads_match_reference_golden.pytestmark[0].args[1],
                         ids=["text_zipf", "text_dirty"])
def test_reference_golden_text_workloads_through_the_scan_plumbing(fixture, maker, monkeypatch, tmp_path):
    """The same golden outputs of the REAL reference, this time through the lowered text path's host side
    (scan plumbing, line-count by-product, relabel, memoised cross, native sink)."""
    W.test_text_workloads_match_reference_golden(_text_fake(monkeypatch), tmp_path, fixture, maker)
