"""Cell partition + halo plan: pure host logic, checked serially and with a 2-rank gloo job on CPU."""
import os
import sys

import numpy as np
import pytest

import helpers as H

capi = H.capi
C = __import__("ctypes")


class Part:
    def __init__(self, mesh, dp, du, rank, nranks):
        lib = capi.load_host()
        self.h = lib.peh_partition(mesh.h, dp.h, du.h, rank, nranks)
        assert self.h, lib.peh_last_error()
        v = capi.PartView()
        lib.peh_part_view_get(self.h, C.byref(v))
        self.mesh = capi.Mesh(v.mesh)
        self.cell_global = capi._np(v.cell_global, self.mesh.n_cells, np.int64)
        self.n_owned_cells = v.n_owned_cells
        self.field = []
        for f, d in ((0, dp), (1, du)):
            F = v.field[f]
            nn = F.n_neighbors
            send_ptr = capi._np(F.send_ptr, nn + 1, np.int64)
            self.field.append(dict(
                n_owned=F.n_owned, n_local=F.n_local,
                cell_dofs=capi._np(F.cell_dofs, self.mesh.n_cells * d.n_loc, np.int32).reshape(-1, d.n_loc),
                l2g=capi._np(F.local_to_global, F.n_local, np.int64),
                neigh=capi._np(F.neighbor_rank, nn, np.int32), send_ptr=send_ptr,
                send_idx=capi._np(F.send_idx, int(send_ptr[-1]) if nn else 0, np.int32),
                recv_ptr=capi._np(F.recv_ptr, nn + 1, np.int64)))

    def __del__(self):
        capi.load_host().peh_part_destroy(self.h)


def scrambled_box(dim, refine, seed=7):
    """A structured box whose cells come in random order — the situation of a mesh file written by a mesh generator."""
    m = capi.mesh_rectangle(dim, [10.0] * dim, refine)
    rng = np.random.default_rng(seed)
    m.permute_cells(rng.permutation(m.arrays.n_cells))
    return m


def meshes():
    sfc = scrambled_box(3, 2)
    sfc.reorder_sfc()
    yield "scrambled 4^3 in space-filling-curve order", sfc, 1
    gm = capi.mesh_read_msh(H.ROOT / "tests" / "golden" / "distorted_hex4.msh", 3)
    gm.reorder_sfc()
    yield "distorted Gmsh hexes in space-filling-curve order", gm, 2
    yield "morton 4^3", capi.mesh_rectangle(3, [10, 10, 10], 2), 1
    yield "lexi 5x3x4 Q2", capi.mesh_subdivided(3, [10, 10, 10], [5, 3, 4]), 2
    yield "2d 8x8 Q2", capi.mesh_rectangle(2, [10, 10], 3), 2


@pytest.mark.parametrize("balanced", [False, True])
@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
def test_partition_invariants(nranks, balanced, monkeypatch):
    monkeypatch.setenv("PE_BALANCED_OWNERSHIP", "1" if balanced else "0")
    for name, mesh, deg in meshes():
        dim = mesh.arrays.dim
        dp, du = capi.HostDofs(mesh, 1, 1), capi.HostDofs(mesh, deg, dim)
        parts = [Part(mesh, dp, du, r, nranks) for r in range(nranks)]
        for f, d in ((0, dp), (1, du)):
            owned = [p.field[f]["l2g"][: p.field[f]["n_owned"]] for p in parts]
            allo = np.concatenate(owned)
            assert len(allo) == d.n_dofs and len(np.unique(allo)) == d.n_dofs, name  # disjoint cover
            owner = np.empty(d.n_dofs, int)
            for r, o in enumerate(owned):
                owner[o] = r
            for r, p in enumerate(parts):  # owned = [interior ascending | boundary ascending]; interior rows touch no ghost
                F = p.field[f]
                o = owned[r]
                touches_ghost = np.zeros(F["n_local"], bool)
                cd = F["cell_dofs"]
                ghost_cells = (cd >= F["n_owned"]).any(axis=1)
                touches_ghost[np.unique(cd[ghost_cells])] = True
                tb = touches_ghost[: F["n_owned"]]
                n_int = int((~tb).sum())
                assert not tb[:n_int].any() and tb[n_int:].all(), name
                assert np.all(np.diff(o[:n_int]) > 0) and np.all(np.diff(o[n_int:]) > 0)
            for r, p in enumerate(parts):
                F = p.field[f]
                # local cell dofs map back to the global numbering
                assert np.array_equal(F["l2g"][F["cell_dofs"]], d.cell_dofs[p.cell_global]), name
                # every cell that touches an owned dof is local (owned rows assemble without communication)
                touch = np.where((owner[d.cell_dofs] == r).any(axis=1))[0]
                assert set(touch) <= set(p.cell_global.tolist())
                # ghosts are grouped by owner in neighbour order and match the receive plan
                ghosts = F["l2g"][F["n_owned"]:]
                for k, q in enumerate(F["neigh"]):
                    seg = ghosts[F["recv_ptr"][k]: F["recv_ptr"][k + 1]]
                    assert np.all(owner[seg] == q)
                    # what q sends to r is exactly this segment, in this order
                    Fq = parts[q].field[f]
                    kk = list(Fq["neigh"]).index(r)
                    sent = Fq["l2g"][Fq["send_idx"][Fq["send_ptr"][kk]: Fq["send_ptr"][kk + 1]]]
                    assert np.array_equal(sent, seg), (name, r, q)
                assert F["recv_ptr"][-1] == len(ghosts) if len(F["neigh"]) else len(ghosts) == 0
        # boundary faces of the local meshes are domain-boundary faces only
        nb = sum(len(p.mesh.bface_cell) for p in parts)
        assert nb >= len(mesh.arrays.bface_cell)


@pytest.mark.parametrize("dim,refine", [(2, 4), (3, 3)])
def test_space_filling_curve_order_makes_contiguous_ranges_compact(dim, refine):
    """partitioned unstructured meshes.  mesh.hpp::reorder_cells_sfc must (i) only permute cells (same
    geometry, same boundary faces), (ii) recover the locality of the Morton order a scrambled mesh has lost: the halo of
    the partition shrinks from 'nearly everything' to within 1.5x of the structured Morton partition."""
    ref = capi.mesh_rectangle(dim, [10.0] * dim, refine)
    bad = scrambled_box(dim, refine)
    fixed = scrambled_box(dim, refine)
    perm = fixed.reorder_sfc()
    a, b = bad.arrays, fixed.arrays
    assert sorted(perm.tolist()) == list(range(a.n_cells))
    assert np.array_equal(b.cell_vertices, a.cell_vertices[perm])
    faces = lambda m: sorted((tuple(sorted(m.cell_vertices[c][[v for v in range(1 << dim) if ((v >> (f // 2)) & 1) == f % 2]])), i)
                             for c, f, i in zip(m.bface_cell, m.bface_local, m.bface_id))
    assert faces(a) == faces(b) == faces(ref.arrays)
    ctr = b.xyz[b.cell_vertices].mean(axis=1)
    assert np.abs(np.diff(ctr, axis=0)).max(axis=1).mean() < 2.0 * 10.0 / 2 ** refine  # consecutive cells are neighbours on average
    nranks = 4
    ghosts = {}
    for name, mesh in (("morton", ref), ("scrambled", bad), ("sfc", fixed)):
        dp, du = capi.HostDofs(mesh, 1, 1), capi.HostDofs(mesh, 1, dim)
        parts = [Part(mesh, dp, du, r, nranks) for r in range(nranks)]
        ghosts[name] = sum(p.field[1]["n_local"] - p.field[1]["n_owned"] for p in parts)
    assert ghosts["sfc"] <= 1.5 * ghosts["morton"]
    assert ghosts["scrambled"] > 2 * ghosts["sfc"]


def test_balanced_ownership_evens_out_the_rows(monkeypatch):
    """PE_BALANCED_OWNERSHIP=1: interface nodes are dealt out among the ranks that touch them instead of all going to the
    lowest rank; whole nodes move (block rows stay intact) and the spread of owned rows over 8 octants shrinks."""
    mesh = capi.mesh_rectangle(3, [10, 10, 10], 4)  # 16^3 cells, 8 octants of 8^3
    dp, du = capi.HostDofs(mesh, 1, 1), capi.HostDofs(mesh, 1, 3)
    spread = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("PE_BALANCED_OWNERSHIP", mode)
        parts = [Part(mesh, dp, du, r, 8) for r in range(8)]
        rows = np.array([p.field[1]["n_owned"] for p in parts])
        assert rows.sum() == du.n_dofs and (rows % 3 == 0).all()
        for p in parts:
            o = p.field[1]["l2g"][: p.field[1]["n_owned"]]
            assert (o.reshape(-1, 3) // 3 == (o[::3] // 3)[:, None]).all() or set((o // 3).tolist()) == set((o[::3] // 3).tolist())
        spread[mode] = rows.max() / rows.mean()
    assert spread["0"] > 1.15          # 9^3 nodes on the corner octant against a mean of 17^3 / 8: 1.19 at this small size
    assert spread["1"] < 1.08
    assert spread["1"] < spread["0"]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(H.ROOT / "tests"))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mesh = capi.mesh_rectangle(3, [10, 10, 10], 2)
        dp, du = capi.HostDofs(mesh, 1, 1), capi.HostDofs(mesh, 1, 3)
        part = Part(mesh, dp, du, rank, world)
        # a global SPD test matrix on the u pattern: graph Laplacian + I, identical on every rank
        import scipy.sparse as sp
        rows = np.repeat(du.cell_dofs, du.n_loc, axis=1).ravel()
        cols = np.tile(du.cell_dofs, (1, du.n_loc)).ravel()
        A = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(du.n_dofs, du.n_dofs))
        A.data[:] = 1.0 + (A.indices % 7) * 0.125
        xg = np.cos(np.arange(du.n_dofs) * 0.37)
        F = part.field[1]
        x = np.zeros(F["n_local"])
        x[: F["n_owned"]] = xg[F["l2g"][: F["n_owned"]]]
        # halo exchange following the plan (what pe_halo_exchange does with ncclSend/ncclRecv)
        reqs, bufs = [], []
        for k, qk in enumerate(F["neigh"]):
            s = torch.from_numpy(x[F["send_idx"][F["send_ptr"][k]: F["send_ptr"][k + 1]]].copy())
            r = torch.zeros(int(F["recv_ptr"][k + 1] - F["recv_ptr"][k]), dtype=torch.float64)
            reqs += [dist.isend(s, int(qk)), dist.irecv(r, int(qk))]
            bufs.append((k, r))
        for rq in reqs:
            rq.wait()
        for k, r in bufs:
            x[F["n_owned"] + F["recv_ptr"][k]: F["n_owned"] + F["recv_ptr"][k + 1]] = r.numpy()
        ok_halo = np.array_equal(x, xg[F["l2g"]])
        # owned rows of A only reference local columns; local SpMV + allreduced dot match the global ones
        g2l = -np.ones(du.n_dofs, int)
        g2l[F["l2g"]] = np.arange(F["n_local"])
        Aown = A[F["l2g"][: F["n_owned"]]]
        ok_cols = np.all(g2l[Aown.indices] >= 0)
        y = Aown @ xg  # == local rows times local columns
        yl = sp.csr_matrix((Aown.data, g2l[Aown.indices], Aown.indptr), shape=(F["n_owned"], F["n_local"])) @ x
        dot = torch.tensor([float(yl @ x[: F["n_owned"]])], dtype=torch.float64)
        dist.all_reduce(dot)
        ok_dot = abs(dot.item() - xg @ (A @ xg)) <= 1e-9 * abs(dot.item())
        q.put((rank, bool(ok_halo), bool(ok_cols), bool(np.allclose(y, yl)), bool(ok_dot)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("balanced", ["0", "1"])
def test_halo_plan_with_two_gloo_ranks(balanced, monkeypatch):
    import torch.multiprocessing as mp
    monkeypatch.setenv("PE_BALANCED_OWNERSHIP", balanced)  # inherited by the spawned ranks
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + int(balanced)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r


class StructuredPart(Part):
    """peh_partition_structured: the same view, built from the lattice alone"""

    def __init__(self, dim, size, cells, morton, rank, nranks):
        lib = capi.load_host()
        s = np.asarray(size, dtype=np.float64)
        n = np.asarray(list(cells) + [1] * (3 - len(cells)), dtype=np.int32)
        self.h = lib.peh_partition_structured(dim, capi._p(s, C.c_double), capi._p(n, C.c_int32), int(morton), rank, nranks)
        assert self.h, lib.peh_last_error()
        v = capi.PartView()
        lib.peh_part_view_get(self.h, C.byref(v))
        self.mesh = capi.Mesh(v.mesh)
        self.cell_global = capi._np(v.cell_global, self.mesh.n_cells, np.int64)
        self.n_owned_cells = v.n_owned_cells
        self.field = []
        for f, n_loc in ((0, 1 << dim), (1, (1 << dim) * dim)):
            F = v.field[f]
            nn = F.n_neighbors
            send_ptr = capi._np(F.send_ptr, nn + 1, np.int64)
            self.field.append(dict(
                n_owned=F.n_owned, n_local=F.n_local,
                cell_dofs=capi._np(F.cell_dofs, self.mesh.n_cells * n_loc, np.int32).reshape(-1, n_loc),
                l2g=capi._np(F.local_to_global, F.n_local, np.int64),
                neigh=capi._np(F.neighbor_rank, nn, np.int32), send_ptr=send_ptr,
                send_idx=capi._np(F.send_idx, int(send_ptr[-1]) if nn else 0, np.int32),
                recv_ptr=capi._np(F.recv_ptr, nn + 1, np.int64)))


@pytest.mark.parametrize("dim,cells,morton", [(3, [4, 4, 4], True), (3, [8, 8, 8], True), (3, [5, 3, 4], False), (3, [7, 5, 6], False),
                                              (2, [8, 8], True), (2, [6, 4], False), (3, [3, 3, 12], False)])
@pytest.mark.parametrize("nranks", [2, 3, 4, 8])
@pytest.mark.parametrize("balanced", [False, True])
def test_structured_part_equals_the_general_partition(dim, cells, morton, nranks, balanced, monkeypatch):
    """The per-rank structured builder (no global mesh / dof maps on the rank) reproduces make_part() array by array: local
    sub-mesh, cell lists, local numbering [owned interior | owned boundary | ghosts by owner], global ids, send / receive plans."""
    monkeypatch.setenv("PE_BALANCED_OWNERSHIP", "1" if balanced else "0")
    size = [10.0, 7.0, 13.0][:dim]
    if morton:
        level = int(np.log2(cells[0]))
        m = capi.mesh_rectangle(dim, size, level)
    else:
        m = capi.mesh_subdivided(dim, size, cells)
    dp, du = capi.HostDofs(m, 1, 1), capi.HostDofs(m, 1, dim)
    for rank in range(nranks):
        A = Part(m, dp, du, rank, nranks)
        B = StructuredPart(dim, size, cells, morton, rank, nranks)
        assert A.n_owned_cells == B.n_owned_cells
        assert np.array_equal(A.cell_global, B.cell_global)
        for name in ("xyz", "cell_vertices", "bface_cell", "bface_local", "bface_id"):
            assert np.array_equal(getattr(A.mesh, name), getattr(B.mesh, name)), name
        for f in range(2):
            for k in A.field[f]:
                assert np.array_equal(np.asarray(A.field[f][k]), np.asarray(B.field[f][k])), (f, k, rank)
