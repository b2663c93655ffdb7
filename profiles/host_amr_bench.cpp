// host_amr_bench.cpp — timing of the host side of one adaptive-refinement pass (csrc/host/amr.hpp), no GPU involved.
//   g++ -O2 -fopenmp -std=c++17 -o /tmp/host_amr_bench profiles/host_amr_bench.cpp && /tmp/host_amr_bench 6
// Builds a 3D box of 2^base cells per axis, refines twice around the well axis (1.35 M active cells for base = 6), then
// times the pieces PoroElasticProblem::refine_mesh runs on the host: transfer store, Kelly estimator, fixed-fraction
// marking, prepare + execute, extraction of the active mesh.  Output of this round: profiles/host_amr_bench.txt.
#include <chrono>
#include <cstdio>
#include "../poroelasticity-dealii_b200/csrc/host/amr.hpp"
using clk = std::chrono::steady_clock;
static double since(clk::time_point t) { return std::chrono::duration<double>(clk::now() - t).count(); }
int main(int argc, char** argv) {
  int base = argc > 1 ? atoi(argv[1]) : 6, rounds = 2;
  double size[3] = {10, 10, 10};
  mesh::Mesh m0 = mesh::create_hyper_rectangle(3, size, base);
  amr::Forest F = amr::Forest::from_mesh(m0, base);
  for (int r = 0; r < rounds; ++r) {
    auto act = F.active_cells();
    for (int32_t c : act) {
      double x = 0, y = 0;
      for (int k = 0; k < 8; ++k) { x += F.xyz[(int64_t)F.cells[c].v[k] * 3]; y += F.xyz[(int64_t)F.cells[c].v[k] * 3 + 1]; }
      F.cells[c].refine_flag = std::hypot(x / 8, y / 8) < 2.5 / (r + 1);
    }
    F.execute();
  }
  // now a typical AMR pass on the big mesh: transfer payload + flags from an estimator
  mesh::Mesh am = F.active_mesh();
  dofs::DofMap dp = dofs::distribute_dofs(am, 1, 1);
  std::vector<double> p(dp.n_dofs, 1.0);
  const double* in[3] = {p.data(), p.data(), p.data()};
  auto t = clk::now(); F.store_vertex_values(am, dp, 3, in); printf("store %.2f\n", since(t));
  std::vector<double> vv(F.n_vertices());
  for (int64_t v = 0; v < F.n_vertices(); ++v) vv[v] = std::sin(F.xyz[v * 3]) * std::cos(F.xyz[v * 3 + 1]);
  t = clk::now(); auto eta = amr::kelly_estimate(F, vv); printf("kelly %.2f\n", since(t));
  t = clk::now(); amr::mark_fixed_fraction(F, eta, 0.6, 0.4, base, base + 3); printf("mark %.2f\n", since(t));
  int nr = 0, nc = 0; for (auto& c : F.cells) { nr += c.refine_flag; nc += c.coarsen_flag; }
  printf("flags: refine %d coarsen %d of %lld\n", nr, nc, (long long)am.n_cells());
  t = clk::now(); auto T = F.line_table(); printf("line_table %.2f (%lld lines)\n", since(t), (long long)T.n_lines());
  t = clk::now(); F.prepare(); printf("prepare %.2f\n", since(t));
  t = clk::now(); auto r = F.execute(); printf("execute %.2f (coarsened %d refined %d)\n", since(t), r.first, r.second);
  t = clk::now(); am = F.active_mesh(); printf("active_mesh %.2f (%lld cells)\n", since(t), (long long)am.n_cells());
  // setup_dofs() of the new mesh (problem.hpp): numbering of both handlers, hanging-node + Dirichlet tables, flattening
  dofs::NodeMaps mp, mu;
  t = clk::now(); dofs::DofMap np_ = dofs::distribute_dofs(am, 1, 1, &mp); dofs::DofMap nu = dofs::distribute_dofs(am, 1, 3, &mu);
  printf("distribute_dofs p+u %.2f (%lld + %lld dofs)\n", since(t), (long long)np_.n_dofs, (long long)nu.n_dofs);
  std::vector<int32_t> ld, ed; std::vector<int64_t> ep; std::vector<double> ew, lg;
  t = clk::now();
  dofs::ConstraintTable tp; tp.init(np_.n_dofs); amr::hanging_node_constraints(F, am, np_, mp, tp); tp.close(); tp.flatten(ld, ep, ed, ew, lg);
  printf("constraints p %.2f (%zu lines)\n", since(t), ld.size());
  t = clk::now();
  dofs::ConstraintTable tu; tu.init(nu.n_dofs); amr::hanging_node_constraints(F, am, nu, mu, tu);
  dofs::add_dirichlet(tu, am, nu, {0, 1, 2, 3, 4, 5}, {0, 0, 1, 1, 2, 2}, {0, -1e-5, 0, -1e-5, 0, -1e-5});
  tu.close(); tu.flatten(ld, ep, ed, ew, lg);
  printf("constraints u %.2f (%zu lines)\n", since(t), ld.size());
  return 0;
}
