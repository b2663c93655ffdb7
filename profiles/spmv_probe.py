"""SpMV micro-benchmark: assembled displacement (A) and pressure (J) matrices of the 3D Q1/Q1
config at the given refinement, x ~ uniform(-1, 1) from a fixed seed, `reps` timed repetitions after a
warm-up, CUDA events on the library's stream.  Also the command profiled with ncu (profiles/README.md).
usage: python profiles/spmv_probe.py <refine> [reps]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 7
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
capi, fss = H.capi, H.fss
inp = capi.InputData(text=H.make_input(dim=3, refine=refine, degree_u=1))
mesh = fss.make_mesh(inp)
dev = capi.create_device_backend(0)
fss.upload_problem(dev, inp, mesh)
dev.pressure_set_uniform(inp.p_init)
dev.displacement_assemble()
dev.assemble_jacobian(inp.time_step)
st = dev.stats()
rng = np.random.default_rng(1234)
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
out = {"refine": refine, "reps": reps, "peak_gbs": peak}
for name, which, n, nbytes in (("A_u", capi.MAT_ELASTICITY, st["n_dofs_u"], st["spmv_bytes_u"]), ("J_p", capi.MAT_JACOBIAN, st["n_dofs_p"], st["spmv_bytes_p"])):
    x = rng.uniform(-1, 1, n)
    capi.device_spmv(dev, which, x, reps=20)
    ms, _ = capi.device_spmv(dev, which, x, reps=reps)
    out[name] = {"rows": n, "algorithmic_bytes": nbytes, "ms": ms, "gbs": nbytes / ms / 1e6, "frac_of_measured_peak": nbytes / ms / 1e6 / peak,
                 "frac_of_8TBs": nbytes / ms / 1e6 / 8000.0}
print(json.dumps(out))
dev.close()
