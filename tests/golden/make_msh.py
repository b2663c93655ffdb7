"""Writes tests/golden/square10.msh: a 10x10 quad mesh of [-5,5]^2 in Gmsh 2.2 ASCII with the same
conventions as the reference's shipped domain.msh (domain.geo:22-28): physical lines 0 bottom, 1 right,
2 top, 3 left; physical surface 7; counter-clockwise quads; O(1e-12) jitter on interior nodes like a
real Gmsh run produces.  Run:  python tests/golden/make_msh.py"""
from pathlib import Path

import numpy as np

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
