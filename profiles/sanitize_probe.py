"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): 3D Q1/Q1 4^3 cells, initial state + one time step,
default paths (block-CSR + persistent CG kernel).  usage: compute-sanitizer --tool memcheck python profiles/sanitize_probe.py"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import helpers as H  # noqa: E402

inp = H.capi.InputData(text=H.make_input(dim=3, refine=2, degree_u=1))
mesh = H.fss.make_mesh(inp)
dev = H.capi.create_device_backend(0)
H.fss.upload_problem(dev, inp, mesh)
print(H.fss.initialize(dev, inp))
print({k: v for k, v in H.fss.time_step(dev, inp).items() if k != "residual_history"})
dev.close()
print("sanitize_probe: done")
