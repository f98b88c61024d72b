"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm prints
one JSON line with the agreed keys (the b200 arm needs a device and is exercised on the GPU box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "16", *extra], capture_output=True, text=True, check=True, cwd=ROOT).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    return json.loads(lines[0])


def test_reference_arm_json_line():
    d = _run()
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_device():
    try:
        import torch
        if torch.cuda.is_available():
            return
    except ImportError:
        pass
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, cwd=ROOT)
    assert p.returncode != 0 and "no CPU path" in (p.stderr + p.stdout)


def test_numa_binding_is_a_no_op_without_a_device():
    """bench.py binds each rank to its GPU's NUMA node when N > 1; without a device (or without sysfs topology) it
    must leave the process where it is and report None."""
    import os
    import bench
    before = os.sched_getaffinity(0)
    assert bench.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before
