// Exercises the C++ mirror of the reference's call sites (csrc/jne_host.hpp) exactly as INTEGRATION.md section 3 shows it:
// calculate_eigenvalues_parallel with a sender callback, then calculate_eigenvalues for one seed.
// Prints "seed ev0 ev1 ..." lines (hex floats) that tests/test_host_cpp.py compares with the Python binding.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "jne_host.hpp"

int main(int argc, char** argv) {
  const uint32_t dim = argc > 1 ? std::atoi(argv[1]) : 2, steps = argc > 2 ? std::atoi(argv[2]) : 103;
  const uint32_t n = argc > 3 ? std::atoi(argv[3]) : 5;
  const int model = argc > 4 ? std::atoi(argv[4]) : 0;
  try {
    jne::Engine gpu({0});
    std::vector<uint32_t> seeds(n);
    for (uint32_t i = 0; i < n; ++i) seeds[i] = i + 1;
    std::map<uint32_t, std::vector<double>> got;
    jne::calculate_eigenvalues_parallel(gpu, dim, steps, seeds, jne::Model((uint8_t)model),
        [&](uint32_t seed, const double* ev, int p) { got[seed].assign(ev, ev + p); }, /*quiet*/ true, /*chunk*/ 2);
    for (auto& kv : got) {
      std::printf("%u", kv.first);
      for (double v : kv.second) std::printf(" %a", v);
      std::printf("\n");
    }
    const std::vector<double> one = jne::calculate_eigenvalues(gpu, dim, steps, 3, jne::Model((uint8_t)model));
    std::printf("single");
    for (double v : one) std::printf(" %a", v);
    std::printf("\n");
    try { jne::Model bad(7); } catch (const jne::Error& e) { std::printf("error %d\n", e.status); }
  } catch (const jne::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 3;
  }
  return 0;
}
