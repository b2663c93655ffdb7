#!/usr/bin/env python
"""bench.py — fixed-stress time steps per second on the 3D Q1/Q1 ~6M-DoF config (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (one process per GPU, torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle) on the host cores

A "step" is one pass of PoroElasticProblem::run's time-loop body (lib/include/PoroelasticityFSS.h:328-407):
inner pressure iterations, displacement assemble+solve, strain projection, convergence check; AMR, stresses
and VTK excluded.  Default workload (c4) at every N: 3D unit-cube hex mesh, refine 7 (128^3 cells; 6,440,067
displacement + 2,146,689 pressure DoFs), cell-partitioned across the N ranks => strong scaling.  Other workloads:
c3 (refine 6), c5 (weak scaling, 143^3 cells per GPU), c2 (2D consolidation, 512^2 cells, top traction).

STEP WINDOW.  The physics is transient (CG iterations per step fall as the pressure pulse decays), so the window is
pinned: after the W warm-up steps the state of the initialised problem (FSS:310-317) is restored and the timed region
runs TIME STEPS 1..K.  Every run, every N and both arms therefore measure the same K steps; the CPU arm uses the
oracle's own recorded iteration counts of exactly these steps (tests/golden/oracle_counts_r*.json).

Printed JSON (one line, rank 0): value = steps/s with all state resident in HBM (CUDA events on the library's
stream, max over ranks); e2e = the same K steps through the C-ABI with HOST buffers (pinned host -> device copy of
the step's state and device -> host copy of the result inside the timed region); roofline = the displacement-matrix
pass (dominant kernel); cpu_baseline = the CPU oracle on a bounded sample (see `sample`); parity = fields of this very
run against the recorded oracle run (field samples at 4096 lattice nodes + norms); the run FAILS (rc 3) above 1e-8;
parity.reference_run = the same samples against the reference's own sources run on a deal.II API shim (C3: 4 steps, C4: 2).
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
PARITY_TOL = 1e-8


def golden_tag(args):
    """Name of the recorded oracle run of a workload under tests/golden (None: the CPU oracle cannot hold it, no record)."""
    return {"c2": f"c2_r{args.refine}", "c3": f"r{args.refine}", "c4": f"r{args.refine}"}.get(args.workload)


def input_text(args, world):
    H = importlib.import_module("poroelasticity-dealii_b200").inputs
    extra = (f"  set Preconditioner = {args.precond}\n  set Chebyshev degree = {args.cheb_degree}\n"
             f"  set Chebyshev eigenvalue ratio = {args.eig_ratio}\n  set CG max iterations = {args.max_its}\n")
    if args.workload == "c2":  # BASELINE configs[1]: undrained top load (DS:249-277), rollers elsewhere
        return H.make_input(dim=2, refine=args.refine, degree_u=1, dirichlet=([0, 1, 2], [0, 0, 1], [0.0, 0.0, 0.0]),
                            neumann=([3], [1], [-1e6]), extra_gpu=extra)
    if args.workload == "c5":
        cells, size = weak_cells(world)
        text = H.make_input(dim=3, refine=args.refine, degree_u=1, extra_gpu=extra, cells=cells)
        return text.replace("set Domain size              = 10, 10, 10", "set Domain size              = " + ", ".join(str(x) for x in size))
    return H.make_input(dim=3, refine=args.refine, degree_u=1, extra_gpu=extra)


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


# ---- the recorded oracle run (tests/golden/make_oracle_counts.py) -------------------------------------------------------
COUNT_KEYS = ("pressure_iterations", "cg_its_pressure", "cg_its_displacement", "cg_its_projection")


def golden_record(tag):
    p = ROOT / "tests" / "golden" / f"oracle_counts_{tag}.json"
    return json.loads(p.read_text()) if tag and p.exists() else None


def golden_counts(tag, steps):
    """Oracle iteration counts of TIME STEPS 1..steps (the bench's pinned window), per step.  Steps beyond the end of the
    record repeat its last step (flagged: the counts only fall as the transient decays, so this cannot flatter the GPU)."""
    rec = golden_record(tag)
    if rec is None:  # no record at this size: scale the largest recorded mesh with 2^levels (kappa ~ h^-1 at best); flagged
        stem, refine = tag.rsplit("r", 1)
        for r in range(int(refine) - 1, 3, -1):
            if golden_record(f"{stem}r{r}") is not None:
                g = golden_counts(f"{stem}r{r}", steps)
                f = 2.0 ** (int(refine) - r)
                per = [{k: (s[k] * f if k.startswith("cg_its_p") or k == "cg_its_displacement" else s[k]) for k in COUNT_KEYS} for s in g["per_step"]]
                return {"per_step": per, "recorded_steps": 0, "source": g["source"] + f" x{f:g} (extrapolated in mesh size)", "extrapolated": True}
        raise RuntimeError(f"no oracle count record for {tag} under tests/golden")
    rs = rec["steps"]
    per = [{k: rs[min(i, len(rs) - 1)][k] for k in COUNT_KEYS} for i in range(steps)]
    return {"per_step": per, "recorded_steps": len(rs), "source": f"oracle_counts_{tag}.json steps 1..{min(steps, len(rs))}"
            + ("" if steps <= len(rs) else f", steps {len(rs) + 1}..{steps} repeat step {len(rs)}"), "extrapolated": steps > len(rs)}


def golden_fields(tag):
    p = ROOT / "tests" / "golden" / f"oracle_fields_{tag}.npz"
    return np.load(p) if tag and p.exists() else None


def oracle_setup(args, threads):
    """Loads the CPU oracle (built for THIS host's CPU) and assembles the workload's systems."""
    sys.path.insert(0, str(ROOT / "tests"))  # the oracle binding lives with the tests; only the CPU legs import it
    import helpers as H
    lib = H.load_oracle()
    thr = lib.po_set_threads(threads)  # explicit: torchrun exports OMP_NUM_THREADS=1
    inp = H.capi.InputData(text=input_text(args, 1))
    mesh = H.fss.make_mesh(inp)
    b = H.create_oracle_backend()
    prm = inp.params()
    t0 = time.perf_counter()
    H.fss.upload_problem(b, inp, mesh, prm)
    return H, b, inp, prm, thr, time.perf_counter() - t0


def cpu_sample(args, threads, counts, cg_sample_its=((1, 3), (2, 6), (2, 6))):
    """Bounded CPU sample of the same workload with the oracle (the reference's SSOR-CG algorithm): builds the full-size
    systems, times each solver twice from a zero start — capped at `lo` and at `hi` iterations — and one call of each assembly
    operator, and prices TIME STEPS 1..K with the oracle's own recorded per-step iteration counts (a full oracle step at 128^3
    takes ~15 minutes).  Cost per CG iteration = (t_hi - t_lo) / (hi - lo); what a solve costs besides its iterations (the
    first preconditioner application, norms) = t_lo - lo * that.  (Round 1 divided one short run by its + 1, which prices the
    start-up like an iteration and came out 20-25 % low against a full step.)"""
    H, b, inp, prm, thr, t_setup = oracle_setup(args, threads)
    capi, fss = H.capi, H.fss
    dt = inp.time_step

    def timed(fn):
        t = time.perf_counter()
        try:
            fn()
        except capi.BackendError as e:  # the capped CG runs end in NoConvergence by construction
            if e.status != capi.PE_ERR_NO_CONVERGENCE:
                raise
        return time.perf_counter() - t

    def solver_cost(solve, reset, lo_hi):
        lo, hi = lo_hi
        t = []
        for cap in (lo, hi):
            prm.cg_max_iterations = cap
            b.set_params(prm)
            reset()
            t.append(timed(solve))
        per_it = (t[1] - t[0]) / (hi - lo)
        return per_it, max(0.0, t[0] - lo * per_it)

    b.pressure_set_uniform(inp.p_init)
    t_asm_u_first = timed(b.displacement_assemble)      # matrix + rhs (once per run)
    zero_u, zero_p = np.zeros(b.n_u), np.zeros(b.n_p)
    t_u, f_u = solver_cost(b.displacement_solve, lambda: b.set_vector(capi.VEC_U, zero_u), cg_sample_its[0])
    t_asm_u = timed(b.displacement_assemble)            # rhs only (every FSS iteration)
    b.project_assemble_matrix()
    t_proj_rhs = timed(lambda: b.project_assemble_rhs(fss.VOLUMETRIC_COMPONENTS[inp.dim]))
    t_m, f_m = solver_cost(lambda: b.project_solve(0), lambda: b.set_vector(capi.VEC_STRAIN0, zero_p), cg_sample_its[2])
    b.volumetric_strain_from_projection([fss.TENSOR_TO_ENTRY[inp.dim][c] for c in fss.VOLUMETRIC_COMPONENTS[inp.dim]], True)
    b.pressure_begin_step(); b.pressure_zero_update(); b.update_volumetric_strain()
    t_res = timed(lambda: b.assemble_residual(dt))
    t_jac = timed(lambda: b.assemble_jacobian(dt))
    t_p, f_p = solver_cost(b.pressure_solve, b.pressure_zero_update, cg_sample_its[1])
    b.close()

    def price(g):
        n_p_solves = max(0, g["pressure_iterations"] - 1)  # the pass that finds the residual converged does not solve (FSS:366-371)
        return (g["pressure_iterations"] * t_res + n_p_solves * (t_jac + f_p) + g["cg_its_pressure"] * t_p + t_asm_u + f_u + g["cg_its_displacement"] * t_u +
                t_proj_rhs + inp.dim * f_m + g["cg_its_projection"] * t_m + t_res)

    per_step = [price(g) for g in counts["per_step"]]
    mean = float(np.mean(per_step))
    detail = {"setup_s": round(t_setup, 2), "s_per_cg_it_u": t_u, "s_per_cg_it_p": t_p, "s_per_cg_it_proj": t_m, "s_per_solve_u": f_u, "s_per_solve_p": f_p,
              "s_per_solve_proj": f_m, "cg_sample_its_lo_hi": list(map(list, cg_sample_its)), "s_residual": t_res,
              "s_jacobian": t_jac, "s_u_rhs": t_asm_u, "s_u_matrix_and_rhs": t_asm_u_first, "s_proj_rhs": t_proj_rhs,
              "counts_source": counts["source"], "counts_extrapolated": counts["extrapolated"],
              "mean_counts": {k: float(np.mean([g[k] for g in counts["per_step"]])) for k in COUNT_KEYS},
              "est_s_per_step": mean, "est_s_first_step": per_step[0], "est_s_last_step": per_step[-1]}
    return 1.0 / mean, thr, detail, price


def extrapolation_check(args, threads):
    """One FULL oracle time step at C3 (64^3) next to the sampled estimate of the same step at the same size: says how far the
    few-iteration sample is from a real run."""
    a3 = argparse.Namespace(**vars(args))
    a3.workload, a3.refine = "c3", 6
    counts = golden_counts("r6", 1)
    _, thr, detail, _ = cpu_sample(a3, threads, counts)
    H, b, inp, prm, thr, _ = oracle_setup(a3, threads)
    prm.cg_max_iterations = 4000
    b.set_params(prm)
    t0 = time.perf_counter()
    H.fss.initialize(b, inp)
    t_init = time.perf_counter() - t0
    t0 = time.perf_counter()
    rep = H.fss.time_step(b, inp)
    actual = time.perf_counter() - t0
    b.close()
    return {"workload": "c3 (64^3), time step 1", "estimated_s": detail["est_s_per_step"], "actual_s": actual, "est_over_actual": detail["est_s_per_step"] / actual,
            "actual_counts": {k: rep[k] for k in COUNT_KEYS}, "recorded_counts": counts["per_step"][0], "oracle_init_s": t_init, "cores": thr}


def sample_text(counts):
    return ("oracle (SSOR-CG, reference settings) at full size: the u/p/projection solvers timed at 1 and 3 / 2 and 6 / 2 and 6 CG iterations from a "
            "zero start (cost per iteration = the difference) and one call of each assembly operator; time steps 1..K (the GPU arm's window) priced "
            "with the oracle's recorded per-step iteration counts (%s); the oracle reproduces the reference's own sources run on a deal.II API shim "
            "(oracle/_ref/fss_ref) iteration count for iteration count, at full size for the 64^3 configuration (tests/test_reference_run.py)" % counts["source"])


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU algorithm (oracle port; deal.II cannot be built here) on the host cores."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if golden_tag(args) is None:
        print(json.dumps({"impl": "reference", "unavailable": f"no recorded oracle run for workload {args.workload} (a full CPU step at this size takes hours)"}), flush=True)
        return
    counts = golden_counts(golden_tag(args), args.steps)
    # one bounded sample (about a minute at 128^3 on 16 cores: full-size assembly once, ten CG iterations of each solver); the arm is
    # launched at every N of a scaling run and measures the same thing each time
    v, thr, detail, _ = cpu_sample(args, cores, counts)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample_text(counts)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "detail": detail}
    if not args.no_extrapolation_check and world == 1:  # one full oracle step at 64^3 (~2.5 min): once per scaling run is enough
        try:
            line["extrapolation_check"] = extrapolation_check(args, cores)
        except Exception as exc:
            line["extrapolation_check"] = {"failed": str(exc)}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    common = {"preconditioner": "chebyshev-jacobi" if args.precond == 1 else "jacobi", "chebyshev_degree": args.cheb_degree if args.precond == 1 else None,
              "chebyshev_inner_passes": ("fp32 copy of the matrix values (preconditioner only)" if os.environ.get("PE_CHEB_FP32", "1") != "0" else "fp64") if args.precond == 1 else None,
              "cg_max_iterations": args.max_its, "l2_policy": "inputs larger than L2 (GBs of matrix vs 126 MB L2), no flush",
              "step_window": f"time steps 1..{args.steps} of the run (state of the initialised problem restored after the {args.warmup} warm-up steps)"}
    if args.workload == "c5":
        cells, size = weak_cells(world)
        nn = (cells[0] + 1) * (cells[1] + 1) * (cells[2] + 1)
        return {"workload": f"3D weak scaling: one 143^3-cell block per GPU stacked along z ({cells[0]}x{cells[1]}x{cells[2]} cells, {3 * nn} u + {nn} p DoFs), "
                            "Q1/Q1, shipped input.data properties; BASELINE.json configs[4]", "parallelism": f"z-slabs x{world}", **common}
    if args.workload == "c2":
        n = 2 ** args.refine
        return {"workload": f"2D consolidation on a uniformly refined square, Q1/Q1, refine {args.refine} ({n}^2 cells, {2 * (n + 1) ** 2} u + {(n + 1) ** 2} p DoFs), "
                            "top traction -1e6 (Neumann), rollers elsewhere, shipped properties; BASELINE.json configs[1]", "refine": args.refine,
                "parallelism": f"cell-partitioned x{world}", **common}
    n = 2 ** args.refine
    return {"workload": f"3D unit-cube hex mesh, Q1 displacement / Q1 pressure, refine {args.refine} ({n}^3 cells, {3 * (n + 1) ** 3} u + {(n + 1) ** 3} p DoFs), "
                        f"shipped input.data properties, dt=60, rollers on all faces, well source; BASELINE.json configs[{3 if args.refine == 7 else 2}]",
            "refine": args.refine, "parallelism": f"cell-partitioned x{world}", **common}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--refine", type=int, default=None)
    ap.add_argument("--precond", type=int, default=1, help="0 Jacobi, 1 Chebyshev-Jacobi polynomial (default)")
    ap.add_argument("--cheb-degree", type=int, default=4)
    ap.add_argument("--eig-ratio", type=float, default=30.0)
    ap.add_argument("--max-its", type=int, default=4000)
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4", "c5"],
                    help="c4: 128^3 strong scaling (headline); c3: 64^3; c5: 143^3 cells per GPU, weak scaling; c2: 2D 512^2 consolidation")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extrapolation-check", action="store_true")
    args = ap.parse_args()
    if args.refine is None:
        args.refine = {"c2": 9, "c3": 6, "c4": 7, "c5": 7}[args.workload]
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

    def allreduce_max(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    inp = capi.InputData(text=input_text(args, world))
    dim = inp.dim
    prob = capi.Problem(inp, device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    t0 = time.perf_counter()
    prob.initialize(verbose=False)
    t_init = allreduce_max(time.perf_counter() - t0)
    be = prob.backend
    lib = be.lib
    stream = torch.cuda.ExternalStream(lib.pe_stream(be.ctx), device=torch.device("cuda", local))

    f64 = C.POINTER(C.c_double)
    n_p, n_u = be.n_p, be.n_u
    # state that defines a time step (fields + warm starts of the three solvers), kept on pinned host memory
    vol_entries = {2: [0, 2], 3: [0, 3, 5]}[dim]
    state_ids = {"p": (capi.VEC_P, n_p), "u": (capi.VEC_U, n_u), "ev": (capi.VEC_VOL_STRAIN, n_p), "ev0": (capi.VEC_VOL_STRAIN0, n_p)}
    for e in vol_entries:
        state_ids[f"e{e}"] = (capi.VEC_STRAIN0 + e, n_p)
    host = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k, (_, n) in state_ids.items()}
    init_state = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k, (_, n) in state_ids.items()}

    def ptr(t):
        return C.cast(t.data_ptr(), f64)

    def download_state(dst):
        for k, (which, n) in state_ids.items():
            be._ck(lib.pe_get_vector(be.ctx, which, ptr(dst[k]), n), "get_vector")

    def upload_state(src):
        for k, (which, n) in state_ids.items():
            be._ck(lib.pe_set_vector(be.ctx, which, ptr(src[k]), n), "set_vector")

    # ---- parity record: sample dofs this rank owns
    gold_f = golden_fields(golden_tag(args))
    gold_c = golden_record(golden_tag(args))
    sample = None
    if gold_f is not None:
        gp, gu = prob.global_ids(capi.FIELD_PRESSURE), prob.global_ids(capi.FIELD_DISPLACEMENT)

        def locate(gids, wanted):
            order = np.argsort(gids)
            pos = np.searchsorted(gids[order], wanted)
            pos = np.minimum(pos, len(gids) - 1)
            hit = gids[order][pos] == wanted
            return np.nonzero(hit)[0], order[pos[hit]]

        sp_i, sp_l = locate(gp, gold_f["p_dof"])
        su_i, su_l = locate(gu, gold_f["u_dof"])
        sample = {"p_i": sp_i, "p_l": sp_l, "u_i": su_i, "u_l": su_l}

    def sample_fields(state):
        """(indices into the golden sample, values) of the sample nodes this rank owns + local sums for the norms"""
        p, u = state["p"].numpy(), state["u"].numpy()
        out = {"p_sq": float(p @ p), "p_sum": float(p.sum()), "u_sq": float(u @ u)}
        if sample is not None:
            out["p_i"], out["p_v"] = sample["p_i"], p[sample["p_l"]].copy()
            out["u_i"], out["u_v"] = sample["u_i"], np.stack([u[sample["u_l"] + a] for a in range(dim)], axis=1)
        return out

    download_state(init_state)  # the initialised problem (not timed)
    sampler = ClockSampler(local)
    sampler.start()
    warm_reports = [prob.step() for _ in range(args.warmup)]
    upload_state(init_state)    # back to time step 0: the timed region is steps 1..K whatever W was
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
    ms_total = allreduce_max(e0.elapsed_time(e1))
    stats = be.stats()
    lib.pe_set_profiling(be.ctx, 0)
    download_state(host)
    final_resident = sample_fields(host)

    # ---- e2e: the SAME K steps again through the C-ABI with host buffers: every step uploads its input state from
    # pinned host memory (H2D) and downloads the resulting state (D2H) inside the timed region; the downloaded fields of
    # every step feed the parity record
    e2e, per_step_fields = None, []
    if not args.no_e2e:
        for k in state_ids:
            host[k].copy_(init_state[k])
        barrier()
        t0 = time.perf_counter()
        e2e_reports = []
        for _ in range(args.steps):
            upload_state(host)
            e2e_reports.append(prob.step())
            download_state(host)
            per_step_fields.append(sample_fields(host))
        barrier()
        t_e2e = allreduce_max(time.perf_counter() - t0)
        nbytes = int(sum(n for _, n in state_ids.values()) * 8)
        e2e = {"value": args.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "cg_its_displacement": [r["cg_its_displacement"] for r in e2e_reports]}
    else:
        per_step_fields = [None] * (args.steps - 1) + [final_resident]

    # ---- parity against the recorded oracle run (gathered on rank 0)
    gathered = [per_step_fields]
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(per_step_fields, gathered, dst=0)
    parity = None
    rc = 0
    if rank == 0:
        if gold_f is None or gold_c is None:
            parity = {"checked": False, "reason": "no recorded oracle run for this workload (CPU oracle cannot hold it / not recorded)"}
        else:
            n_rec = min(len(gold_c["steps"]), gold_f["p"].shape[0] - 1)
            worst, per, n_cmp = 0.0, [], 0
            for s in range(args.steps):
                if s + 1 > n_rec or gathered[0][s] is None:
                    continue
                P, U = np.full(gold_f["p"].shape[1], np.nan), np.full(gold_f["u"].shape[1:], np.nan)
                p_sq = p_sum = u_sq = 0.0
                for r in range(world):
                    f = gathered[r][s]
                    P[f["p_i"]] = f["p_v"]
                    U[f["u_i"]] = f["u_v"]
                    p_sq, p_sum, u_sq = p_sq + f["p_sq"], p_sum + f["p_sum"], u_sq + f["u_sq"]
                gp_, gu_, gc = gold_f["p"][s + 1], gold_f["u"][s + 1], gold_c["steps"][s]
                errs = {"p_samples_rel_l2": float(np.linalg.norm(P - gp_) / np.linalg.norm(gp_)),
                        "u_samples_rel_l2": float(np.linalg.norm(U - gu_) / np.linalg.norm(gu_)),
                        "p_l2_rel": abs(np.sqrt(p_sq) - gc["p_l2"]) / gc["p_l2"], "p_sum_rel": abs(p_sum - gc["p_sum"]) / abs(gc["p_sum"]),
                        "u_l2_rel": abs(np.sqrt(u_sq) - gc["u_l2"]) / gc["u_l2"]}
                errs = {k: (v if np.isfinite(v) else float("inf")) for k, v in errs.items()}
                per.append({"step": s + 1, **errs})
                worst = max(worst, *errs.values())
                n_cmp += 1
            # the same configuration run by the reference's own code (oracle/_ref/fss_ref on a deal.II API shim; C3: 4 time steps,
            # C4: 2), at the same sample dofs — informative only: the oracle record above differs from it by 1e-13
            ref_run = None
            try:
                rp = ROOT / "tests" / "golden" / {"r6": "reference_run_q1_c3_r6.json", "r7": "reference_run_q1_c4_r7_2steps.json"}.get(golden_tag(args) or "", "none")
                if rp.exists():
                    rr = json.loads(rp.read_text())
                    devs = []
                    for s in range(min(args.steps, rr["n_steps"])):
                        if gathered[0][s] is None:
                            continue
                        P, U = np.full(gold_f["p"].shape[1], np.nan), np.full(gold_f["u"].shape[1:], np.nan)
                        for r in range(world):
                            P[gathered[r][s]["p_i"]] = gathered[r][s]["p_v"]
                            U[gathered[r][s]["u_i"]] = gathered[r][s]["u_v"]
                        rp_, ru_ = np.array(rr["steps"][s]["p_samples"]), np.array(rr["steps"][s]["u_samples"])
                        devs += [float(np.linalg.norm(P - rp_) / np.linalg.norm(rp_)), float(np.linalg.norm(U - ru_) / np.linalg.norm(ru_))]
                    if devs:
                        ref_run = {"max_rel": max(devs), "steps_compared": len(devs) // 2, "against": f"tests/golden/{rp.name} (the reference's own sources, unmodified, run on a deal.II API shim)"}
            except Exception as exc:  # never let the extra comparison take the line down
                ref_run = {"failed": str(exc)}
            parity = {"checked": n_cmp > 0, "parity_max_rel": worst if n_cmp else None, "tolerance": PARITY_TOL, "steps_compared": n_cmp, "reference_run": ref_run,
                      "samples": int(gold_f["p"].shape[1]), "against": f"tests/golden/oracle_counts_{golden_tag(args)}.json + oracle_fields_{golden_tag(args)}.npz (CPU oracle, SSOR-CG)",
                      "per_step": per if len(per) <= 4 else per[:2] + per[-2:], "ok": bool(n_cmp == 0 or worst <= PARITY_TOL)}
            if n_cmp and worst > PARITY_TOL:
                rc = 3

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak = peaks.get("hbm_gbs", 6650.0)
        spmv_ms = stats["spmv_ms_u"] / max(1, stats["spmv_timed_u"])
        achieved = stats["spmv_bytes_u"] / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else None
        spmv_ms_p = stats["spmv_ms_p"] / max(1, stats["spmv_timed_p"])
        inner_ms = stats["inner_ms_u"] / max(1, stats["inner_passes_u"])
        value = args.steps / (ms_total * 1e-3)
        traffic, ncu_dram_peak = None, None
        tr = ROOT / "profiles" / "traffic_r2.json"
        if tr.exists():
            cap = json.loads(tr.read_text()).get("k_spmv_sell<3,double> C4 (128^3 cells, 1 GPU)", {})
            if world == 1 and args.workload == "c4":
                traffic = cap.get("traffic")
            if cap.get("dram_gbs") and cap.get("dram_pct_of_ncu_peak"):  # what ncu calls 100 % DRAM throughput on this part (a read-only stream
                ncu_dram_peak = cap["dram_gbs"] / (cap["dram_pct_of_ncu_peak"] / 100.0)  # runs above the read+write copy the measured peak is)
        bsr = int(stats["bsr_block_size"])
        sell_fmt = bool(stats["sell_format_u"])
        fmt = (f"sliced block-ELL {bsr}x{bsr} (TMA-fed)" if sell_fmt else (f"block-CSR {bsr}x{bsr}" if bsr else "CSR"))
        if stats["pcg_iterations_u"] > 0:
            kernel_name = (f"k_pcg2 (persistent single-reduction CG kernel): its FP64 pass w = A z + three dot products on the {fmt} displacement matrix, "
                           "incl. the halo wait and the one grid/peer reduction that ends the pass")
            timing_source = "in-kernel %globaltimer of CTA 0 around every FP64 pass of the timed region (a persistent kernel has no per-pass launches to bracket with CUDA events)"
        else:
            kernel_name = ("k_spmv_sell<3,double,*>" if sell_fmt else (f"k_spmv_bsr<{bsr},*>" if bsr else "k_spmv<32,*>")) + f" ({fmt} SpMV of the displacement matrix with fused epilogues)"
            timing_source = "CUDA events around every launch on the library's stream, inside the timed region"
        cg_u = [r["cg_its_displacement"] for r in reports]
        # what a step spends outside the displacement solve is small; ms per displacement-CG iteration makes rounds comparable
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak" if args.workload == "c5" else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clocks, "gpu_launches": int(stats["kernel_launches"]),
            "e2e": e2e, "parity": parity,
            "ms_per_displacement_cg_iteration": (ms_total / max(1, sum(cg_u))),
            "displacement_solve_ms_per_cg_iteration": (stats["pcg_ms_u"] / max(1, stats["pcg_iterations_u"])) if stats["pcg_iterations_u"] else None,
            "roofline": {"bound": "hbm", "kernel": kernel_name, "timing_source": timing_source,
                         "achieved": achieved, "peak": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None,
                         "frac_of_ncu_dram_peak": (achieved / ncu_dram_peak) if (achieved and ncu_dram_peak) else None, "ncu_dram_peak_gbs": ncu_dram_peak,
                         "traffic": traffic, "traffic_source": "profiles/traffic_r2.json (ncu --set full capture of this matrix pass as a stand-alone kernel on this workload)" if traffic else None,
                         "algorithmic_bytes_per_launch": stats["spmv_bytes_u"], "matrix_format": fmt, "avg_launch_ms": spmv_ms,
                         "csr_equivalent_gbs": ((stats["nnz_u"] * 12.0 + stats["n_dofs_u"] * 20.0) / (spmv_ms * 1e-3) / 1e9) if spmv_ms > 0 else None,
                         "launches_timed": int(stats["spmv_timed_u"]),
                         "preconditioner_pass": {"avg_ms": inner_ms, "algorithmic_bytes": stats["inner_bytes_u"], "passes_timed": int(stats["inner_passes_u"]),
                                                 "achieved": (stats["inner_bytes_u"] / (inner_ms * 1e-3) / 1e9) if inner_ms > 0 else None,
                                                 "frac": (stats["inner_bytes_u"] / (inner_ms * 1e-3) / 1e9 / peak) if inner_ms > 0 else None} if stats["inner_passes_u"] else None,
                         "pressure_spmv": {"avg_launch_ms": spmv_ms_p, "achieved": (stats["spmv_bytes_p"] / (spmv_ms_p * 1e-3) / 1e9) if spmv_ms_p > 0 else None,
                                           "launches_timed": int(stats["spmv_timed_p"])},
                         "phase_ms_per_step": {"fp64_passes": stats["spmv_ms_u"] / args.steps, "preconditioner_passes": stats["inner_ms_u"] / args.steps,
                                               "vector_updates": stats["update_ms_u"] / args.steps, "reductions": stats["reduce_ms_u"] / args.steps,
                                               "displacement_solves": stats["pcg_ms_u"] / args.steps, "pressure_and_projection_solves": stats["pcg_ms_p"] / args.steps,
                                               "pressure_and_projection_phases": dict(zip(("cg_pass_streams", "inner_passes", "updates", "reductions", "wait_barrier_inner", "wait_barrier_cg",
                                                                                           "wait_peer_mailboxes", "wait_barrier_update", "n_cg_passes", "n_inner_passes"),
                                                                                          [v / args.steps for v in stats["phase_ms_p"]])),
                                               "waits_inside_the_phases": {"barrier_behind_inner_passes": stats["wait_inner_ms_u"] / args.steps,
                                                                           "barrier_behind_cg_pass": stats["wait_cg_ms_u"] / args.steps,
                                                                           "peer_mailboxes": stats["wait_peer_ms_u"] / args.steps,
                                                                           "barrier_behind_updates": stats["wait_update_ms_u"] / args.steps}},
                         "spmv_share_of_step": (stats["spmv_ms_u"] + stats["inner_ms_u"] + stats["spmv_ms_p"]) / ms_total if ms_total > 0 else None},
            "iterations_per_step": {"pressure_inner": float(np.mean([r["pressure_iterations"] for r in reports])),
                                    "cg_pressure": float(np.mean([r["cg_its_pressure"] for r in reports])),
                                    "cg_displacement": float(np.mean(cg_u)),
                                    "cg_projection": float(np.mean([r["cg_its_projection"] for r in reports])),
                                    "fss": float(np.mean([r["fss_iterations"] for r in reports])),
                                    "cg_displacement_per_step": cg_u,
                                    "warmup_equals_timed": [r["cg_its_displacement"] for r in warm_reports] == cg_u[:len(warm_reports)] if len(warm_reports) <= len(cg_u) else None,
                                    "matrix_passes_u": stats["spmv_launches_u"] / args.steps, "matrix_passes_p": stats["spmv_launches_p"] / args.steps,
                                    "fp64_pass_equivalents_u": (stats["spmv_timed_u"] + stats["inner_passes_u"] * (stats["inner_bytes_u"] / stats["spmv_bytes_u"] if stats["spmv_bytes_u"] else 1.0)) / args.steps},
            "init_s": t_init, "setup_ms": stats["setup_ms"],
        }
        if world == 1 and not args.no_cpu_baseline and golden_tag(args) is not None:
            try:
                counts = golden_counts(golden_tag(args), args.steps)
                v, thr, detail, _ = cpu_sample(args, os.cpu_count() or 1, counts)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample_text(counts), "detail": detail}
            except Exception as exc:  # the baseline must never take the GPU number down with it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
        print(json.dumps(line), flush=True)
    prob.close()
    if world > 1:
        rc_t = torch.tensor([rc], device="cuda")
        dist.broadcast(rc_t, src=0)
        rc = int(rc_t.item())
        dist.destroy_process_group()
    if rc:
        sys.stderr.write(f"bench.py: field parity against the recorded oracle run exceeded {PARITY_TOL:g}\n")
        sys.exit(rc)


if __name__ == "__main__":
    main()
