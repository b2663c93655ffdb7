"""Writes tests/golden/square10.msh: a 10x10 quad mesh of [-5,5]^2 in Gmsh 2.2 ASCII with the same
conventions as the reference's shipped domain.msh (domain.geo:22-28): physical lines 0 bottom, 1 right,
2 top, 3 left; physical surface 7; counter-clockwise quads; O(1e-12) jitter on interior nodes like a
real Gmsh run produces.  Run:  python tests/golden/make_msh.py"""
from pathlib import Path

import numpy as np

def square10():
    n = 10
    rng = np.random.default_rng(7)
    xs = np.linspace(-5, 5, n + 1)
    ids = {}
    nodes = []
    for j in range(n + 1):
        for i in range(n + 1):
            jit = rng.uniform(-1e-12, 1e-12, 2) if 0 < i < n and 0 < j < n else np.zeros(2)
            ids[(i, j)] = len(nodes) + 1
            nodes.append((xs[i] + jit[0], xs[j] + jit[1]))
    elems = []
    for i in range(n):
        elems.append((1, 0, 1, ids[(i, 0)], ids[(i + 1, 0)]))       # bottom
    for j in range(n):
        elems.append((1, 1, 2, ids[(n, j)], ids[(n, j + 1)]))       # right
    for i in range(n):
        elems.append((1, 2, 3, ids[(n - i, n)], ids[(n - i - 1, n)]))  # top
    for j in range(n):
        elems.append((1, 3, 4, ids[(0, n - j)], ids[(0, n - j - 1)]))  # left
    for j in range(n):
        for i in range(n):
            elems.append((3, 7, 6, ids[(i, j)], ids[(i + 1, j)], ids[(i + 1, j + 1)], ids[(i, j + 1)]))
    out = ["$MeshFormat", "2.2 0 8", "$EndMeshFormat", "$Nodes", str(len(nodes))]
    out += [f"{k + 1} {x:.16g} {y:.16g} 0" for k, (x, y) in enumerate(nodes)]
    out += ["$EndNodes", "$Elements", str(len(elems))]
    for k, e in enumerate(elems):
        out.append(f"{k + 1} {e[0]} 2 {e[1]} {e[2]} " + " ".join(str(v) for v in e[3:]))
    out += ["$EndElements", ""]
    Path(__file__).with_name("square10.msh").write_text("\n".join(out))




# ---- distorted meshes (non-affine cells): exercise the general J^-T / det J path, the Gmsh hex reader and boundary ids
def write_msh(path, nodes, elems):
    """path: a file name next to this script, or an absolute path"""
    out = ["$MeshFormat", "2.2 0 8", "$EndMeshFormat", "$Nodes", str(len(nodes))]
    out += [f"{k + 1} " + " ".join(f"{c:.16g}" for c in p) for k, p in enumerate(nodes)]
    out += ["$EndNodes", "$Elements", str(len(elems))]
    for k, e in enumerate(elems):
        out.append(f"{k + 1} {e[0]} 2 {e[1]} {e[2]} " + " ".join(str(v) for v in e[3:]))
    out += ["$EndElements", ""]
    target = Path(path) if Path(path).is_absolute() else Path(__file__).with_name(path)
    target.write_text("\n".join(out))


def distorted_quad(n=8, amp=0.2, seed=11):
    rng = np.random.default_rng(seed)
    h = 10.0 / n
    idx = lambda i, j: j * (n + 1) + i + 1
    nodes = []
    for j in range(n + 1):
        for i in range(n + 1):
            d = rng.uniform(-amp * h, amp * h, 2) if 0 < i < n and 0 < j < n else np.zeros(2)
            nodes.append((-5 + i * h + d[0], -5 + j * h + d[1], 0.0))
    elems = []
    # boundary ids in the colorize convention: 0/1 = x-min/x-max, 2/3 = y-min/y-max
    for j in range(n):
        elems.append((1, 0, 1, idx(0, j), idx(0, j + 1)))
        elems.append((1, 1, 2, idx(n, j), idx(n, j + 1)))
    for i in range(n):
        elems.append((1, 2, 3, idx(i, 0), idx(i + 1, 0)))
        elems.append((1, 3, 4, idx(i, n), idx(i + 1, n)))
    for j in range(n):
        for i in range(n):
            elems.append((3, 7, 6, idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)))
    write_msh(f"distorted_quad{n}.msh", nodes, elems)


def distorted_hex(n=4, amp=0.15, seed=13, path=None, shuffle=False):
    rng = np.random.default_rng(seed)
    h = 10.0 / n
    idx = lambda i, j, k: (k * (n + 1) + j) * (n + 1) + i + 1
    nodes = []
    for k in range(n + 1):
        for j in range(n + 1):
            for i in range(n + 1):
                inside = 0 < i < n and 0 < j < n and 0 < k < n
                d = rng.uniform(-amp * h, amp * h, 3) if inside else np.zeros(3)
                nodes.append((-5 + i * h + d[0], -5 + j * h + d[1], -5 + k * h + d[2]))
    elems = []
    for a in range(n):
        for b in range(n):
            elems.append((3, 0, 1, idx(0, a, b), idx(0, a + 1, b), idx(0, a + 1, b + 1), idx(0, a, b + 1)))
            elems.append((3, 1, 2, idx(n, a, b), idx(n, a + 1, b), idx(n, a + 1, b + 1), idx(n, a, b + 1)))
            elems.append((3, 2, 3, idx(a, 0, b), idx(a + 1, 0, b), idx(a + 1, 0, b + 1), idx(a, 0, b + 1)))
            elems.append((3, 3, 4, idx(a, n, b), idx(a + 1, n, b), idx(a + 1, n, b + 1), idx(a, n, b + 1)))
            elems.append((3, 4, 5, idx(a, b, 0), idx(a + 1, b, 0), idx(a + 1, b + 1, 0), idx(a, b + 1, 0)))
            elems.append((3, 5, 6, idx(a, b, n), idx(a + 1, b, n), idx(a + 1, b + 1, n), idx(a, b + 1, n)))
    for k in range(n):
        for j in range(n):
            for i in range(n):  # gmsh hexahedron: bottom face counter-clockwise, then the top face
                elems.append((5, 7, 6, idx(i, j, k), idx(i + 1, j, k), idx(i + 1, j + 1, k), idx(i, j + 1, k),
                              idx(i, j, k + 1), idx(i + 1, j, k + 1), idx(i + 1, j + 1, k + 1), idx(i, j + 1, k + 1)))
    if shuffle:  # a mesh generator does not write cells in lattice order: shuffle the hexahedra (boundary quads stay in front)
        nb = 6 * n * n
        perm = np.random.default_rng(seed + 1).permutation(len(elems) - nb)
        elems = elems[:nb] + [elems[nb + q] for q in perm]
    write_msh(path or f"distorted_hex{n}.msh", nodes, elems)


if __name__ == "__main__":
    square10()
    distorted_quad()
    distorted_hex()
