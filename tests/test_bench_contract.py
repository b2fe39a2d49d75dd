"""bench.py's driver-facing contract on a box without a GPU: the reference arm's JSON line (alone and
under torchrun, where rank 0 alone prints), and the product arm refusing to run without CUDA (no CPU
fallback).  The reference arm steps the unmodified reference from baseline/_ref when
__graft_entry__.build() has vendored it there, else the oracle's torch-op port -- both CPU paths."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
SMALL = ["--envs", "2048", "--steps", "2", "--warmup", "1", "--no-ref-cuda"]


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def _check_reference_line(d, n):
    assert d["impl"] == "reference" and "unavailable" not in d
    assert d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == n and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert abs(d["value"] - 2048 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    assert d["config"]["envs_per_gpu"] == 2048 and d["config"]["num_agents"] == 3 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", *SMALL], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 1)


def test_reference_arm_under_torchrun_prints_once():
    """launched like the driver launches N > 1: every rank exits 0, rank 0 alone runs and prints"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", BENCH, "--impl", "reference",
                        "--gpus", "2", *SMALL], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 2)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without CUDA")
def test_product_arm_refuses_without_cuda():
    r = subprocess.run([sys.executable, BENCH, "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert not _json_lines(r.stdout)
    assert "CUDA" in (r.stderr + r.stdout)
