"""`bench.py --impl reference` (the driver's reference arm) runs without a GPU: one JSON line with the contract's keys, timed on
the reference's own compiled sources when oracle/_ref holds them (kind "reference"), else on the oracle port (kind "port")."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env, timeout=240, cwd=ROOT)
    assert r.returncode == 0
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 1.0
    assert d["metric"].startswith("frames/sec ORB extract+match") and d["steps"] == 5
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    have_ref = all(os.path.exists(os.path.join(ROOT, "oracle", "_ref", f)) for f in ("liborbextractor_ref.so", "liborbmatcher_ref.so"))
    assert cb["kind"] == ("reference" if have_ref else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
