// The reference's README / benches/sample.rs:79-99 example on the B200 engine: N(3, I) in `dim` dimensions, DiagNutsSettings
// with num_tune = 1000 and maxdepth = 3, start at 3.5, 1000 post-warmup draws per chain - here for `nchains` chains at once.
//   g++ -std=c++17 -Iinclude examples/sample_normal.cpp -Lnuts_rs_b200 -lnuts_b200 -Wl,-rpath,$PWD/nuts_rs_b200 -o sample_normal
// Exit code 77 = no sm_100 device (the library has no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nuts_b200.hpp"

int main(int argc, char** argv) {
  const uint64_t nchains = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 4;
  const uint64_t dim = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 10;
  const uint64_t num_tune = argc > 3 ? std::strtoull(argv[3], nullptr, 10) : 1000;
  const uint64_t num_draws = argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 1000;
  try {
    nuts_b200::DiagNutsSettings settings;
    settings.num_tune = num_tune;
    settings.maxdepth = 3;  // small value just for testing... (benches/sample.rs:84)
    nuts_b200::CudaMath math = nuts_b200::CudaMath::normal(nchains, dim, 3.0);
    nuts_b200::Chains chains(math, settings, /*seed=*/42);
    const std::vector<double> init(nchains * dim, 3.5);
    for (int32_t st : chains.set_position(init))
      if (st != 0) {
        std::fprintf(stderr, "bad initial point\n");
        return 1;
      }
    nuts_b200::Draws tune = chains.draw(num_tune);
    nuts_b200::Draws draws = chains.draw(num_draws);
    double sum = 0.0, sumsq = 0.0, checksum = 0.0;
    uint64_t steps = 0, divergences = 0, tuning_flags = 0;
    for (uint64_t t = 0; t < num_draws; ++t)
      for (uint64_t c = 0; c < nchains; ++c) {
        const double* x = draws.position(c, t);
        for (uint64_t i = 0; i < dim; ++i) {
          sum += x[i];
          sumsq += (x[i] - 3.0) * (x[i] - 3.0);
          checksum += x[i] * (double)(1 + (i + 3 * c + 7 * t) % 11);
        }
        steps += draws.n_steps[draws.at(c, t)];
        divergences += draws.diverging[draws.at(c, t)];
        tuning_flags += draws.tuning[draws.at(c, t)];
      }
    const double n = (double)(num_draws * nchains * dim);
    std::printf("chains %llu dim %llu draws %llu mean %.6f var %.6f leapfrogs %llu divergences %llu tuning_flags %llu direct %d checksum %.17g\n",
                (unsigned long long)nchains, (unsigned long long)dim, (unsigned long long)num_draws, sum / n, sumsq / n,
                (unsigned long long)steps, (unsigned long long)divergences, (unsigned long long)tuning_flags, (int)chains.last_draw_direct(),
                checksum);
    const bool ok = std::fabs(sum / n - 3.0) < 0.1 && std::fabs(sumsq / n - 1.0) < 0.15 && divergences == 0 && tuning_flags == 0;
    return ok ? 0 : 2;
  } catch (const nuts_b200::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return e.code == NUTS_ERR_NO_DEVICE ? 77 : 1;
  }
}
