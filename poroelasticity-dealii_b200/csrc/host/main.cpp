// main.cpp — the driver the reference never shipped (code/CMakeLists.txt:8 names a missing
// source/Runner.cpp; the stale code/Makefile:128-153 a missing fss-poroel.cc).
//   fss-poroel <input.data>
// argv handling follows parse_command_line (lib/include/parse_command_line.h:5-27); the try/catch
// shape follows the commented-out main at lib/include/PoroelasticityFSS.h:504-537.
// Multi-GPU: start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK set; rank 0 writes the
// NCCL unique id to $PE_NCCL_ID_FILE (default ./.pe_nccl_id), the others read it.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>

#include "problem.hpp"

namespace parse_command_line {
// Behaviour of parse_command_line (PCL:5-27): without arguments print a hint and exit(1); otherwise echo every
// argument on its own line and take the first one as the input file name.
std::string parse_command_line(int argc, char* const* argv) {
  if (argc < 2) {
    std::cout << "specify the file name" << std::endl;
    std::exit(1);
  }
  for (int i = 1; i < argc; ++i) std::cout << argv[i] << std::endl;
  return std::string(argv[1]);
}
}  // namespace parse_command_line

static int env_int(const char* k, int def) {
  const char* v = std::getenv(k);
  return v ? std::atoi(v) : def;
}

int main(int argc, char** argv) {
  try {
    std::string input_file_name = parse_command_line::parse_command_line(argc, argv);
    const int rank = env_int("RANK", 0), nranks = env_int("WORLD_SIZE", 1), device = env_int("LOCAL_RANK", 0);
    input_data::InputDataPoroel data;
    data.read_input_file(input_file_name, /*echo=*/rank == 0);
    unsigned char id[128] = {0};
    size_t id_bytes = 0;
    if (nranks > 1) {
      const char* f = std::getenv("PE_NCCL_ID_FILE");
      std::string path = f ? f : "./.pe_nccl_id";
      if (rank == 0) {
        id_bytes = sizeof id;
        if (pe_nccl_unique_id(id, &id_bytes) != 0) throw std::runtime_error("pe_nccl_unique_id failed");
        std::ofstream o(path + ".tmp", std::ios::binary);
        o.write((const char*)id, (std::streamsize)id_bytes);
        o.close();
        std::rename((path + ".tmp").c_str(), path.c_str());
      } else {
        for (int tries = 0; tries < 600; ++tries) {
          std::ifstream i(path, std::ios::binary);
          if (i) {
            i.read((char*)id, sizeof id);
            id_bytes = (size_t)i.gcount();
            if (id_bytes > 0) break;
          }
          std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
        if (!id_bytes) throw std::runtime_error("timed out waiting for the NCCL id file");
      }
    }
    poro_elastisity::PoroElasticProblem problem(data, device, rank, nranks, id, id_bytes);
    problem.run(/*verbose=*/true);
  } catch (std::exception& exc) {
    std::cerr << std::endl << std::endl << "----------------------------------------------------" << std::endl;
    std::cerr << "Exception on processing: " << std::endl << exc.what() << std::endl << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  } catch (...) {
    std::cerr << std::endl << std::endl << "----------------------------------------------------" << std::endl;
    std::cerr << "Unknown exception!" << std::endl << "Aborting!" << std::endl
              << "----------------------------------------------------" << std::endl;
    return 1;
  }
  return 0;
}
