"""bench.py contract checks that need no GPU: the reference arm prints exactly one JSON line with the keys
the driver reads, and non-zero ranks of a torchrun launch stay silent."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BENCH = os.path.join(ROOT, "oracle", "_ref", "ref_bench_c128_g384")

pytestmark = pytest.mark.skipif(not os.path.exists(REF_BENCH), reason="oracle/_ref not built (python oracle/build_ref.py)")


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--streams", "256", *args],
                          capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run({}, "--steps", "3", "--warmup", "3")
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "10ms frames/sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 3 and d["warmup"] == 3 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "256 streams" in cb["sample"]
    assert d["config"]["streams_per_gpu"] == 256 and "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "2", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""
