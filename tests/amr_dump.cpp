// amr_dump.cpp — writes everything four refinement passes produce on a distorted box (indicators, line table, flags via the
// executed mesh, transferred values, hanging-node constraints) to a file, byte for byte; tests/test_amr.py runs it with different
// thread counts and compares the files.   usage: amr_dump <dim> <base level> <out file>
#include <cstdio>
#include <cstring>
#include "../poroelasticity-dealii_b200/csrc/host/amr.hpp"
int main(int argc, char** argv) {
  int dim = atoi(argv[1]), base = atoi(argv[2]);
  double size[3] = {10, 7, 5};
  mesh::Mesh m0 = mesh::create_hyper_rectangle(dim, size, base);
  // distort vertices deterministically
  for (size_t i = 0; i < m0.xyz.size(); ++i) m0.xyz[i] += 0.02 * std::sin(1.7 * i + 0.3);
  amr::Forest F = amr::Forest::from_mesh(m0, base);
  FILE* f = fopen(argv[3], "wb");
  for (int pass = 0; pass < 4; ++pass) {
    mesh::Mesh am = F.active_mesh();
    dofs::NodeMaps mp;
    dofs::DofMap dp = dofs::distribute_dofs(am, 1, 1, &mp);
    std::vector<double> p(dp.n_dofs), vv(F.n_vertices(), 0.0);
    for (int64_t c = 0; c < am.n_cells(); ++c)
      for (int k = 0; k < (1 << dim); ++k) {
        int32_t v = am.cell_vertices[c * (1 << dim) + k];
        double x = F.xyz[(int64_t)v * dim], y = F.xyz[(int64_t)v * dim + 1];
        double val = std::exp(-0.3 * ((x - 1) * (x - 1) + (y + 0.5) * (y + 0.5))) + 0.01 * x;
        p[dp.cell_dofs[c * (1 << dim) + k]] = val; vv[v] = val;
      }
    const double* in[1] = {p.data()};
    F.store_vertex_values(am, dp, 1, in);
    auto eta = amr::kelly_estimate(F, vv);
    fwrite(eta.data(), sizeof(float), eta.size(), f);
    amr::mark_fixed_fraction(F, eta, 0.6, 0.4, base, base + 3);
    auto T = F.line_table();
    fwrite(T.ptr.data(), 4, T.ptr.size(), f); fwrite(T.members.data(), 4, T.members.size(), f); fwrite(T.half.data(), 4, T.half.size(), f);
    auto r = F.execute();
    fwrite(&r, sizeof r, 1, f);
    mesh::Mesh nm = F.active_mesh();
    fwrite(nm.cell_vertices.data(), 4, nm.cell_vertices.size(), f); fwrite(nm.xyz.data(), 8, nm.xyz.size(), f);
    fwrite(nm.bface_cell.data(), 4, nm.bface_cell.size(), f); fwrite(nm.bface_id.data(), 4, nm.bface_id.size(), f);
    dofs::NodeMaps m2; dofs::DofMap d2 = dofs::distribute_dofs(nm, 1, 1, &m2);
    std::vector<double> out(d2.n_dofs); double* o[1] = {out.data()};
    F.fetch_vertex_values(nm, d2, 1, o);
    fwrite(out.data(), 8, out.size(), f);
    dofs::ConstraintTable t; t.init(d2.n_dofs); amr::hanging_node_constraints(F, nm, d2, m2, t); t.close();
    std::vector<int32_t> ld, ed; std::vector<int64_t> ep; std::vector<double> ew, lg; t.flatten(ld, ep, ed, ew, lg);
    fwrite(ld.data(), 4, ld.size(), f); fwrite(ed.data(), 4, ed.size(), f); fwrite(ew.data(), 8, ew.size(), f);
    printf("pass %d: %lld cells, coarsened %d refined %d, %zu hanging\n", pass, (long long)nm.n_cells(), r.first, r.second, ld.size());
  }
  fclose(f);
  return 0;
}
