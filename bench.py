#!/usr/bin/env python
"""bench.py — fixed-stress time steps per second on the 3D Q1/Q1 ~6M-DoF config (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (one process per GPU, torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle) on the host cores

A "step" is one pass of PoroElasticProblem::run's time-loop body (lib/include/PoroelasticityFSS.h:328-407):
inner pressure iterations, displacement assemble+solve, strain projection, convergence check; AMR, stresses
and VTK excluded.  Workload at every N: 3D unit-cube hex mesh, refine 7 (128^3 cells; 6,440,067
displacement + 2,146,689 pressure DoFs), cell-partitioned across the N ranks => strong scaling.

Printed JSON (one line, rank 0): value = steps/s with all state resident in HBM (CUDA events on the library's
stream, max over ranks); e2e = the same through the C-ABI with HOST buffers (pinned host -> device copy of the
step's state and device -> host copy of p and u inside the timed region); roofline = the displacement-matrix
CSR SpMV (dominant kernel) from per-launch CUDA events recorded inside the timed region; cpu_baseline = the CPU
oracle on a bounded sample (see `sample`).
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "fixed_stress_time_steps_per_second"
UNIT = "steps/s"


def input_text(refine, precond, cheb_degree, eig_ratio, max_its, cells=None, size=None):
    H = importlib.import_module("poroelasticity-dealii_b200").inputs
    extra = (f"  set Preconditioner = {precond}\n  set Chebyshev degree = {cheb_degree}\n"
             f"  set Chebyshev eigenvalue ratio = {eig_ratio}\n  set CG max iterations = {max_its}\n")
    text = H.make_input(dim=3, refine=refine, degree_u=1, extra_gpu=extra, cells=cells)
    if size:
        text = text.replace("set Domain size              = 10, 10, 10", "set Domain size              = " + ", ".join(str(x) for x in size))
    return text


def weak_cells(n_gpus, per_axis=143):
    """BASELINE config 5: one 143^3-cell block (~12 M DoFs) per GPU, stacked along z so that the contiguous cell
    ranges of the lexicographic partition are exactly the blocks."""
    return [per_axis, per_axis, per_axis * n_gpus], [10, 10, 10 * n_gpus]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md).  nvidia-smi needs
    about a second to deliver its first line, so the sampler is started before the warm-up steps and only the samples
    that arrive inside [mark_begin, mark_end] are reported (all samples under load if the window is shorter than
    the sampling period)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.rows, self.proc, self.t0, self.t1 = device, [], None, None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
        window = "timed region"
        if len(inside) < 2:  # timed region shorter than the sampling period: fall back to the whole loaded interval
            inside = [r for t, r in self.rows if self.t0 is None or t >= self.t0 - 5.0]
            window = "warm-up + timed region (timed region shorter than two sampling periods)"
        num = lambda x: x.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in inside if r and num(r[0])]
        mx = [float(r[1]) for r in inside if len(r) > 1 and num(r[1])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in inside if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in inside if len(r) > 2 and num(r[2])]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window}


def cpu_sample(refine, threads, golden, cg_sample_its=(2, 5, 5)):
    """Bounded CPU sample of the same workload with the oracle (the reference's SSOR-CG algorithm).

    Builds the full-size systems, then times a few iterations of each solver and one call of each assembly
    operator, and extrapolates one time step with the oracle's own iteration counts recorded offline by
    tests/golden/make_oracle_counts.py (a full oracle step at 128^3 takes tens of minutes)."""
    sys.path.insert(0, str(ROOT / "tests"))  # the oracle binding lives with the tests; only the CPU legs import it
    import helpers as H
    capi, fss = H.capi, H.fss
    lib = H.load_oracle()
    thr = lib.po_set_threads(threads)
    inp = capi.InputData(text=input_text(refine, 1, 4, 30, 1000))
    mesh = fss.make_mesh(inp)
    b = H.create_oracle_backend()
    prm = inp.params()
    t0 = time.perf_counter()
    fss.upload_problem(b, inp, mesh, prm)
    t_setup = time.perf_counter() - t0
    dt = inp.time_step

    def timed(fn):
        t = time.perf_counter()
        try:
            fn()
        except capi.BackendError as e:  # the capped CG runs end in NoConvergence by construction
            if e.status != capi.PE_ERR_NO_CONVERGENCE:
                raise
        return time.perf_counter() - t

    its_u, its_p, its_m = cg_sample_its
    b.pressure_set_uniform(inp.p_init)
    t_asm_u_first = timed(b.displacement_assemble)      # matrix + rhs (once per run)
    prm.cg_max_iterations = its_u
    b.set_params(prm)
    t_u = timed(b.displacement_solve) / (its_u + 1)     # +1: the initial residual vmult and SSOR apply
    t_asm_u = timed(b.displacement_assemble)            # rhs only (every FSS iteration)
    b.project_assemble_matrix()
    t_proj_rhs = timed(lambda: b.project_assemble_rhs(fss.VOLUMETRIC_COMPONENTS[3]))
    prm.cg_max_iterations = its_m
    b.set_params(prm)
    t_m = timed(lambda: b.project_solve(0)) / (its_m + 1)
    b.volumetric_strain_from_projection([0, 3, 5], True)
    b.pressure_begin_step(); b.pressure_zero_update(); b.update_volumetric_strain()
    t_res = timed(lambda: b.assemble_residual(dt))
    t_jac = timed(lambda: b.assemble_jacobian(dt))
    prm.cg_max_iterations = its_p
    b.set_params(prm)
    t_p = timed(b.pressure_solve) / (its_p + 1)
    b.close()
    g = golden
    per_step = (g["pressure_iterations"] * (t_res + t_jac) + g["cg_its_pressure"] * t_p + t_asm_u + g["cg_its_displacement"] * t_u +
                t_proj_rhs + g["cg_its_projection"] * t_m + t_res)
    detail = {"setup_s": round(t_setup, 2), "s_per_cg_it_u": t_u, "s_per_cg_it_p": t_p, "s_per_cg_it_proj": t_m, "s_residual": t_res,
              "s_jacobian": t_jac, "s_u_rhs": t_asm_u, "s_u_matrix_and_rhs": t_asm_u_first, "s_proj_rhs": t_proj_rhs, "counts": g, "est_s_per_step": per_step}
    return 1.0 / per_step, thr, detail


def golden_counts(refine):
    """Oracle iteration counts per time step (mean over the recorded steps)."""
    p = ROOT / "tests" / "golden" / f"oracle_counts_r{refine}.json"
    if p.exists():
        d = json.loads(p.read_text())
        steps = d["steps"]
        keys = ("pressure_iterations", "cg_its_pressure", "cg_its_displacement", "cg_its_projection")
        return {k: float(np.mean([s[k] for s in steps])) for k in keys} | {"source": p.name, "extrapolated": False}
    # no record at this size: scale the SSOR-CG counts of the largest recorded mesh with 2^(levels) (kappa ~ h^-1 at best);
    # flagged in the output
    for r in range(refine - 1, 3, -1):
        q = ROOT / "tests" / "golden" / f"oracle_counts_r{r}.json"
        if q.exists():
            g = golden_counts(r)
            f = 2.0 ** (refine - r)
            return {"pressure_iterations": g["pressure_iterations"], "cg_its_pressure": g["cg_its_pressure"] * f,
                    "cg_its_displacement": g["cg_its_displacement"] * f, "cg_its_projection": g["cg_its_projection"],
                    "source": q.name + f" x{f:g} (extrapolated)", "extrapolated": True}
    raise RuntimeError("no oracle count record under tests/golden")


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port; deal.II cannot be built here) on the host cores."""
    if rank != 0:
        return
    g = golden_counts(args.refine)
    vals = []
    detail = None
    for _ in range(max(1, min(args.steps, 2))):  # each "step" is one bounded sample; two are enough for a stable number
        v, thr, detail = cpu_sample(args.refine, 0, g)
        vals.append(v)
    v = float(np.median(vals))
    sample = ("oracle (SSOR-CG, reference settings) at full size: %d/%d/%d CG iterations of the u/p/projection solvers and one call of each "
              "assembly operator timed, extrapolated to one step with the oracle's recorded iteration counts (%s)" % (2, 5, 5, g["source"]))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "detail": detail}
    print(json.dumps(line), flush=True)


def workload_config(args):
    if args.workload == "c5":
        cells, size = weak_cells(args.gpus)
        nn = (cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1)
        return {"workload": f"3D weak scaling: one 143^3-cell block per GPU stacked along z ({cells[0]}x{cells[1]}x{cells[2]} cells, {3 * nn} u + {nn} p DoFs), "
                            "Q1/Q1, shipped input.data properties; BASELINE.json configs[4]", "parallelism": f"z-slabs x{args.gpus}",
                "l2_policy": "inputs larger than L2, no flush", "preconditioner": "chebyshev-jacobi" if args.precond == 1 else "jacobi",
                "chebyshev_degree": args.cheb_degree, "cg_max_iterations": args.max_its}
    n = 2 ** args.refine
    return {"workload": f"3D unit-cube hex mesh, Q1 displacement / Q1 pressure, refine {args.refine} ({n}^3 cells, {3 * (n + 1) ** 3} u + {(n + 1) ** 3} p DoFs), "
                        "shipped input.data properties, dt=60, rollers on all faces, well source; BASELINE.json configs[3]",
            "refine": args.refine, "parallelism": f"cell-partitioned x{args.gpus}", "l2_policy": "inputs larger than L2 (6.3 GB matrix vs 126 MB L2), no flush",
            "preconditioner": "chebyshev-jacobi" if args.precond == 1 else "jacobi", "chebyshev_degree": args.cheb_degree,
            "cg_max_iterations": args.max_its}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--refine", type=int, default=7)
    ap.add_argument("--precond", type=int, default=0)
    ap.add_argument("--cheb-degree", type=int, default=4)
    ap.add_argument("--eig-ratio", type=float, default=30.0)
    ap.add_argument("--max-its", type=int, default=4000)
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"], help="c4: 128^3 strong scaling (headline); c5: 143^3 cells per GPU, weak scaling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import __graft_entry__ as G
    G.build()
    pkg = importlib.import_module("poroelasticity-dealii_b200")
    capi = pkg.capi
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        box = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        nccl_id = box[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cells, size = weak_cells(world) if args.workload == "c5" else (None, None)
    inp = capi.InputData(text=input_text(args.refine, args.precond, args.cheb_degree, args.eig_ratio, args.max_its, cells=cells, size=size))
    prob = capi.Problem(inp, device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    t0 = time.perf_counter()
    prob.initialize(verbose=False)
    t_init = time.perf_counter() - t0
    be = prob.backend
    lib = be.lib
    stream = torch.cuda.ExternalStream(lib.pe_stream(be.ctx), device=torch.device("cuda", local))

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        prob.step()
    f64 = C.POINTER(C.c_double)
    n_p, n_u = be.n_p, be.n_u
    # state that defines a time step (fields + warm starts), kept on pinned host memory so the e2e leg can replay
    # exactly the steps of the timed region through host buffers
    state_ids = {"p": (capi.VEC_P, n_p), "u": (capi.VEC_U, n_u), "ev": (capi.VEC_VOL_STRAIN, n_p), "ev0": (capi.VEC_VOL_STRAIN0, n_p),
                 "exx": (capi.VEC_STRAIN0 + 0, n_p), "eyy": (capi.VEC_STRAIN0 + 3, n_p), "ezz": (capi.VEC_STRAIN0 + 5, n_p)}
    host = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k, (_, n) in state_ids.items()}

    def ptr(t):
        return C.cast(t.data_ptr(), f64)

    def download_state():
        for k, (which, n) in state_ids.items():
            be._ck(lib.pe_get_vector(be.ctx, which, ptr(host[k]), n), "get_vector")

    def upload_state():
        for k, (which, n) in state_ids.items():
            be._ck(lib.pe_set_vector(be.ctx, which, ptr(host[k]), n), "set_vector")

    download_state()  # snapshot (not timed)
    # ---- timed region: K steps, state resident in HBM
    be.reset_stats()
    lib.pe_set_profiling(be.ctx, 1)
    barrier()
    sampler.mark_begin()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    reports = [prob.step() for _ in range(args.steps)]
    e1.record(stream)
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    stats = be.stats()
    lib.pe_set_profiling(be.ctx, 0)

    # ---- e2e: the SAME K steps again through the C-ABI with host buffers: every step uploads its input state from
    # pinned host memory (H2D) and downloads the resulting state (D2H) inside the timed region
    e2e = None
    if not args.no_e2e:
        barrier()
        t0 = time.perf_counter()
        e2e_reports = []
        for _ in range(args.steps):
            upload_state()
            e2e_reports.append(prob.step())
            download_state()
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        nbytes = int(sum(n for _, n in state_ids.values()) * 8)
        e2e = {"value": args.steps / float(t_e2e.item()), "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "cg_its_displacement": [r["cg_its_displacement"] for r in e2e_reports]}

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = peaks.get("hbm_gbs", 6650.0)
        spmv_ms = stats["spmv_ms_u"] / max(1, stats["spmv_timed_u"])
        achieved = stats["spmv_bytes_u"] / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
        spmv_ms_p = stats["spmv_ms_p"] / max(1, stats["spmv_timed_p"])
        value = args.steps / (ms_total * 1e-3)
        traffic = None
        tr = ROOT / "profiles" / "traffic_r1.json"
        if tr.exists() and world == 1 and args.refine == 7 and args.workload == "c4":
            key = "k_spmv_bsr<3> C4 (128^3 cells, 1 GPU)" if stats["bsr_block_size"] == 3 else "k_spmv<32> C4 (CSR, PE_FORMAT=csr)"
            traffic = json.loads(tr.read_text()).get(key, {}).get("traffic")
        bsr = int(stats["bsr_block_size"])
        fmt = f"block-CSR {bsr}x{bsr}" if bsr else "CSR"
        if stats["pcg_iterations_u"] > 0:
            kernel_name = (f"k_pcg (persistent Jacobi-CG kernel; its SpMV+dot phase on the {fmt} displacement matrix, "
                           "incl. the in-kernel halo send/wait and the grid/peer reduction that ends the phase)")
            timing_source = "in-kernel %globaltimer of CTA 0 around every SpMV phase of the timed region (a persistent kernel has no per-pass launches to bracket with CUDA events)"
        else:
            kernel_name = (f"k_spmv_bsr<{bsr},*>" if bsr else "k_spmv<32,*>") + f" ({fmt} SpMV of the displacement matrix with fused d.h / residual / Chebyshev epilogues)"
            timing_source = "CUDA events around every launch on the library's stream, inside the timed region"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if args.workload == "c5" else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args), "clocks": clocks, "gpu_launches": int(stats["kernel_launches"]),
            "e2e": e2e,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "timing_source": timing_source,
                         "achieved": achieved, "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                         "traffic": traffic, "traffic_source": "profiles/traffic_r1.json (ncu --set full capture of this kernel on this workload)" if traffic else None,
                         "algorithmic_bytes_per_launch": stats["spmv_bytes_u"], "matrix_format": fmt, "avg_launch_ms": spmv_ms,
                         "csr_equivalent_gbs": ((stats["nnz_u"] * 12.0 + stats["n_dofs_u"] * 20.0) / (spmv_ms * 1e-3) / 1e9) if spmv_ms > 0 else None,
                         "launches_timed": int(stats["spmv_timed_u"]),
                         "pressure_spmv": {"avg_launch_ms": spmv_ms_p, "achieved": (stats["spmv_bytes_p"] / (spmv_ms_p * 1e-3) / 1e9) if spmv_ms_p > 0 else None,
                                           "launches_timed": int(stats["spmv_timed_p"])},
                         "spmv_share_of_step": (stats["spmv_ms_u"] + stats["spmv_ms_p"]) / ms_total if ms_total > 0 else None},
            "iterations_per_step": {"pressure_inner": float(np.mean([r["pressure_iterations"] for r in reports])),
                                    "cg_pressure": float(np.mean([r["cg_its_pressure"] for r in reports])),
                                    "cg_displacement": float(np.mean([r["cg_its_displacement"] for r in reports])),
                                    "cg_projection": float(np.mean([r["cg_its_projection"] for r in reports])),
                                    "fss": float(np.mean([r["fss_iterations"] for r in reports])),
                                    "cg_displacement_per_step": [r["cg_its_displacement"] for r in reports],
                                    "matrix_passes_u": stats["spmv_launches_u"] / args.steps, "matrix_passes_p": stats["spmv_launches_p"] / args.steps},
            "init_s": t_init, "setup_ms": stats["setup_ms"],
        }
        if world == 1 and not args.no_cpu_baseline and args.workload == "c4":
            try:
                g = golden_counts(args.refine)
                v, thr, detail = cpu_sample(args.refine, 0, g)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": thr, "kind": "port",
                                        "sample": "oracle at full size: 2/5/5 CG iterations of the u/p/projection SSOR-CG solvers and one call of each assembly "
                                                  f"operator timed, extrapolated to one step with the oracle's recorded iteration counts ({g['source']})",
                                        "detail": detail}
            except Exception as exc:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
        print(json.dumps(line), flush=True)
    prob.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
