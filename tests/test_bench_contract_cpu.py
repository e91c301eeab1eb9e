"""The bench contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the keys the
driver reads (its CPU step replaced by an instant stand-in -- the real one is tens of seconds per step and is exercised by
the driver), and the product arm refuses to produce a number without a CUDA device (no CPU fallback)."""
import io
import json
import os
import subprocess
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line_has_the_contract_keys(monkeypatch):
    import bench
    calls = []

    def fake_step_factory(sample_frames, threads):
        assert sample_frames == bench.CPU_SAMPLE_FRAMES >= 4     # a [previous, first] pattern that is not degenerate
        assert threads == (os.cpu_count() or 1)                  # all the host threads it can use

        def step():
            calls.append(1)
            return 0.5
        return step

    buf = io.StringIO()
    monkeypatch.setattr(bench, "cpu_reference_step", fake_step_factory)
    monkeypatch.setattr(bench, "_OUT", buf, raising=False)
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference(SimpleNamespace(gpus=1, steps=5, warmup=3))
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "frames/s"
    assert line["metric"].startswith("stylized frames/sec")
    # the run is bounded: at most one warm-up and two timed CPU steps whatever K / W ask for, and the line says so
    assert len(calls) == 3 and line["steps"] == 2 and line["warmup"] == 1
    fps = bench.CPU_SAMPLE_FRAMES / (bench.STEPS_DDIM * 0.5)
    assert abs(line["value"] - fps) < 1e-12 and abs(line["ms_per_step"] - 500.0) < 1e-9
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["vs_baseline"] is None and line["config"]["workload"].startswith("SD-v1.5 three-branch")


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CUDA device" in out.stderr
