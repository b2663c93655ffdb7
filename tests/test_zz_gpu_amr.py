"""GPU parity of the hanging-node / adaptive path against the CPU oracle.

The device kernels of csrc/device/kernels_constraints.cu implement E^T A E on the assembled objects; the same algorithm is
verified on the CPU by tests/test_oracle_amr.py.  First green on a B200 in the round-1 driver run (GPUTEST_r01: all seven
cases passed), so the cases are plain (strict) tests now.  Each runs in its own process with a timeout.
"""
import json
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def run_case(*args, env=None):
    import os
    r = subprocess.run([sys.executable, str(H.ROOT / "tests" / "amr_gpu_case.py"), *map(str, args)], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, **(env or {})))
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "{}"
    print(line)
    print(r.stderr[-2000:])
    assert r.returncode == 0, (r.returncode, line, r.stderr[-2000:])
    return json.loads(line)


@pytest.mark.parametrize("dim,deg,rounds", [(2, 1, 3), (2, 2, 3), (3, 1, 2), (3, 2, 1)])
def test_hanging_node_mesh_matches_oracle(dim, deg, rounds):
    res = run_case("static", dim, deg, rounds)
    assert res["ok"] and res["hanging_p"] > 0


def test_adaptive_driver_matches_the_oracle_mirror():
    res = run_case("driver")
    assert res["ok"] and max(res["cells_per_step"]) > 256


@pytest.mark.parametrize("fp32", ["0", "1"])
def test_chebyshev_with_fp32_inner_passes_matches_oracle(fp32):
    """Opt-in PE_CHEB_FP32=1 (kernels_solver.cu::k_spmv_bsr_cheb_f32): same fields, same iteration count as the FP64 polynomial."""
    res = run_case("chebfp32", env={"PE_CHEB_FP32": fp32})
    assert res["ok"] and res["env"] == fp32
