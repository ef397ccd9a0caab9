"""CPU-only checks of bench.py: the reference arm runs (it is the one bench leg that needs no GPU),
prints one well-formed JSON line, and the interval/traffic helpers behave."""
import importlib.util
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_union_of_overlapping_kernel_intervals():
    b = _bench()
    assert b.union_ms([]) == 0
    assert b.union_ms([(0, 10), (5, 20), (30, 40)]) == 30
    assert b.union_ms([(3, 4), (0, 10)]) == 10


def test_traffic_is_read_from_committed_ncu_summaries():
    b = _bench()
    tr, src = b.ncu_traffic()
    assert tr is None or (6.0e10 < tr < 8.0e10 and "profiles/" in src)     # ~ one 64 GiB batch


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--file-mib", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GiB/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_nonzero_rank_is_silent():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
