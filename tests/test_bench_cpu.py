"""CPU checks of bench.py: the reference arm (`--impl reference`) prints the contract's JSON line, uses every host
core even under torchrun's OMP_NUM_THREADS=1, stays silent on ranks > 0, and its linear-algebra leg computes the
same energy and gradient as the oracle (so the CPU baseline times the right work).  The GPU arm needs a B200 and
is exercised by the driver / `tools/gpu_call_*.sh`."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_line_and_threads():
    out = _run({"OMP_NUM_THREADS": "1"}, "--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "2048")
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("gp_energy+grad evals/sec")
    assert d["unit"] == "evals/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["extrapolated"] is True
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and "sample" in cb
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the arm must still use the host's cores
    assert cb["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--impl", "reference", "--gpus", "2", "--steps", "1",
               "--warmup", "0", "--size", "2048")
    assert out.strip() == ""


def test_cpu_leg_computes_the_oracle_energy_and_gradient():
    import bench
    import gplite_b200 as g
    from oracle import gplite_oracle as o
    n, d = 300, bench.D_FULL
    _, energy, grad = bench.cpu_linalg_eval(n, d)
    X, y = bench.synth(n, d)
    gp = g.gp_init(cfs=g.cf_matern52(lscale=bench.LSCALE, magn=bench.MAGN), lik=g.lik_gaussian(sigma=bench.SIGMA))
    eo, go = o.energy_grad_gaussian_full(gp, X, y, jitter=bench.JITTER, include_prior=False)
    assert abs(energy - eo) < 1e-9 * abs(eo)
    assert np.max(np.abs(grad - go)) < 1e-7 * np.max(np.abs(go))
