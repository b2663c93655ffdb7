// ref_main.cpp — the driver the reference repository names but does not contain (code/CMakeLists.txt:8, source/Runner.cpp; the
// commented-out main at PoroelasticityFSS.h:504-537 shows its shape): read the parameter file named on the command line and call
// PoroElasticProblem<dim>::run().  Everything it calls is the reference's own, unmodified code, compiled from /root/reference by
// oracle/Makefile (target _ref/fss_ref) against the deal.II API shim in oracle/dealii_shim (NOT deal.II).  TEST INFRASTRUCTURE:
// its output is recorded once by tests/golden/make_reference_run.py and pins the oracle; nothing in the product links it.
#include <PoroelasticityFSS.h>
#include <parse_command_line.h>

int main(int argc, char** argv) {
  try {
    const std::string file = parse_command_line::parse_command_line(argc, argv);
    input_data::InputDataPoroel data;
    data.read_input_file(file);
    if (data.dim == 2) {
      poro_elastisity::PoroElasticProblem<2> problem(data);
      problem.run();
    } else if (data.dim == 3) {
      poro_elastisity::PoroElasticProblem<3> problem(data);
      problem.run();
    } else {
      std::cerr << "Dimensions must be 2 or 3" << std::endl;
      return 2;
    }
  } catch (std::exception& exc) {
    std::cerr << "Exception on processing: " << exc.what() << std::endl;
    return 1;
  }
  return 0;
}
