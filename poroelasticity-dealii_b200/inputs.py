"""input.data texts: the reference's shipped file and the synthetic benchmark configurations.

Product-side helper (no oracle involved): bench.py and the tests build their parameter files here; the text goes
through the same parser as a file on disk (csrc/host/input_data.hpp, InputDataPoroel.h:77-222)."""

SHIPPED_INPUT = """
subsection Mesh
  set Dimensions               = 2
  set Domain size              = 10, 10
  set Initial refinement level = 4
  set Max refinement level     = 6
end
subsection In situ
  set Displacement boundary labels     = 0, 1, 2, 3
  set Displacement boundary components = 0, 0, 1, 1
  set Displacement boundary values     = 0, -1e-5, 0, -1e-5
  set Initial pressure                 = 10e6
  set Stress boundary components       =
  set Stress boundary labels           =
  set Stress boundary values           =
end
subsection Properties
  set Young modulus         = 1.4e10
  set Biot coefficient      = 0.9
  set Bulk density          = 2700
  set Fluid compressibility = 5.8e-10   # 1.45e-10 for water
  set Permeability          = 10          # in mDa
  set Poisson ratio         = 0.3
  set Porosity              = 0.3
  set Viscosity             = 1e-3
  set Well radius           = 1
  set Flow rate             = 1e-5
end
subsection Solver
  set Time step  = 60
  set Time max   = 1e3
end
"""


def make_input(dim=2, refine=4, degree_u=2, extra_gpu="", cells=None, neumann=None, dirichlet=None, refine_every=0):
    """input.data text with the shipped properties (input.data:24-35), extended to 3D as SURVEY §8d says.  Uniform mesh unless
    `refine_every` (or a `Refine every` line in extra_gpu) says otherwise: the parser's default is the reference's every-5th-step
    refinement (FSS:333), which the uniform-mesh parity cases and the benchmarks switch off."""
    size = ", ".join(["10"] * dim)
    if dirichlet is None:
        labels = ", ".join(str(i) for i in range(2 * dim))
        comps = ", ".join(str(i // 2) for i in range(2 * dim))
        vals = ", ".join("0" if i % 2 == 0 else "-1e-5" for i in range(2 * dim))
    else:
        labels, comps, vals = (", ".join(str(x) for x in col) for col in dirichlet)
    nl, nc, nv = (", ".join(str(x) for x in col) for col in neumann) if neumann else ("", "", "")
    cells_line = f"  set Cells per axis = {', '.join(str(c) for c in cells)}\n" if cells else ""
    if "Refine every" not in extra_gpu:
        cells_line += f"  set Refine every = {refine_every}\n"
    return f"""
subsection Mesh
  set Dimensions               = {dim}
  set Domain size              = {size}
  set Initial refinement level = {refine}
end
subsection In situ
  set Displacement boundary labels     = {labels}
  set Displacement boundary components = {comps}
  set Displacement boundary values     = {vals}
  set Initial pressure                 = 10e6
  set Stress boundary components       = {nc}
  set Stress boundary labels           = {nl}
  set Stress boundary values           = {nv}
end
subsection Properties
  set Young modulus         = 1.4e10
  set Biot coefficient      = 0.9
  set Bulk density          = 2700
  set Fluid compressibility = 5.8e-10
  set Permeability          = 10
  set Poisson ratio         = 0.3
  set Porosity              = 0.3
  set Viscosity             = 1e-3
  set Well radius           = 1
  set Flow rate             = 1e-5
end
subsection Solver
  set Time step  = 60
  set Time max   = 1e3
end
subsection GPU
  set Displacement FE degree = {degree_u}
{cells_line}{extra_gpu}
end
"""
