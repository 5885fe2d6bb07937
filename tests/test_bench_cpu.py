"""CPU: bench.py's reference arm must give the reference's OpenMP code every CPU
of the box (round 1's arm ran on ONE thread: libgomp, loaded with OMP_PROC_BIND
set, had pinned the process before the affinity mask was read), and both arms
must describe the same workload."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _clean_cpu_count():
    out = subprocess.run([sys.executable, "-c", "import os; print(len(os.sched_getaffinity(0)))"],
                         capture_output=True, text=True, check=True,
                         env={k: v for k, v in os.environ.items() if not k.startswith("OMP_")})
    return int(out.stdout)


def test_reference_arm_uses_every_host_cpu(ref):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    ncpu = _clean_cpu_count()
    # the hostile environment: torchrun's OMP_NUM_THREADS=1, binding requested, and a
    # library that drags libgomp in before bench.py is even imported
    env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="close", OMP_PLACES="cores")
    code = ("import torch, json, sys; sys.path.insert(0, %r); import bench; "
            "dt, thr, N = bench.cpu_reference_cycles(3, 5, 2, 0, 2); "
            "print(json.dumps({'threads': thr, 'N': N, 'host': bench.HOST_CPUS, 'dt': dt}))" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    r = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert r["N"] == 33 and r["dt"] > 0
    assert r["threads"] == ncpu, r
    if ncpu > 1:
        assert r["threads"] > 1


def test_both_arms_print_the_same_config():
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 2, 4, 8):
        a, b = bench.workload_config(n), bench.workload_config(n)
        assert a == b and "workload" in a and "model" not in a
    assert "513^3" in bench.workload_config(1)["workload"]
    assert "4097x513x513" in bench.workload_config(8)["workload"]
