"""The reference arm of bench.py (`--impl reference`) runs without a GPU: check the JSON line the driver parses.  Needs the compiled
reference (oracle/_ref); a 6000-read set keeps it to a few seconds.  CPU only."""
import json
import os
import subprocess
import sys
import pytest

from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "6000", "--steps", "1", "--warmup", "1"] + list(args),
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip().splitlines()


def test_reference_arm_line():
    if not oracle.ref_binary("minimap2-nd"):
        pytest.skip("oracle/_ref not built")
    out = _run({})
    assert len(out) == 1
    d = json.loads(out[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "overlaps/sec" and d["unit"] == "overlaps/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb, e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert d["work"]["records"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    out = _run({"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert out == []
