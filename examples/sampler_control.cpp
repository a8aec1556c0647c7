// The host-side control flow of the reference's Sampler (src/sampler.rs:1253-1552) on the B200 engine: chains start with the
// init-retry loop, run in the background in batches, are paused / resumed, report progress - and the run is checkpointed in the
// middle of the warm-up and resumed in a second set of chains, which must reproduce the remaining draws bit for bit.
//   g++ -std=c++17 -pthread -Iinclude examples/sampler_control.cpp -Lnuts_rs_b200 -lnuts_b200 -Wl,-rpath,$PWD/nuts_rs_b200 -o sampler_control
// Exit code 77 = no sm_100 device (the library has no CPU fallback).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "nuts_b200.hpp"

int main() {
  const uint64_t nchains = 32, dim = 50;
  try {
    nuts_b200::DiagNutsSettings settings;
    settings.num_tune = 100;
    settings.num_draws = 200;
    settings.maxdepth = 6;
    nuts_b200::CudaMath math = nuts_b200::CudaMath::normal(nchains, dim, 3.0);
    nuts_b200::Chains chains(math, settings, /*seed=*/7);
    // chain 5 is handed the mode itself twice (zero gradient => NutsError::BadInitGrad) before it gets a valid point
    int bad_calls = 0;
    auto init = [&](uint64_t chain, double* x) {
      const bool bad = chain == 5 && bad_calls < 2;
      if (bad) ++bad_calls;
      for (uint64_t i = 0; i < dim; ++i) x[i] = bad ? 3.0 : 3.0 + 0.5 * std::sin(0.37 * (double)(chain * dim + i + 1));
    };
    for (int32_t st : chains.set_position_with_retries(init))
      if (st != 0) return std::fprintf(stderr, "chain without a valid initial point\n"), 1;
    if (bad_calls != 2) return std::fprintf(stderr, "retry loop did not run (%d)\n", bad_calls), 1;

    std::vector<double> all((settings.num_tune + settings.num_draws) * nchains * dim);
    nuts_b200::Sampler sampler(chains, settings, [&](const nuts_b200::Draws& d, uint64_t first) {
      std::memcpy(all.data() + first * nchains * dim, d.data(), d.n_draws() * nchains * dim * sizeof(double));
    }, /*batch=*/25);
    sampler.pause();
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    const nuts_b200::Progress p1 = sampler.progress();
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    const nuts_b200::Progress p2 = sampler.progress();
    if (!p2.paused || p1.finished_draws != p2.finished_draws || p2.finished) return std::fprintf(stderr, "pause did not hold\n"), 1;
    // paused between two batches: take a checkpoint here
    nuts_b200::ChainState ckpt = chains.checkpoint();
    const uint64_t at = p2.finished_draws;
    sampler.resume();
    sampler.wait();
    const nuts_b200::Progress p3 = sampler.progress();
    if (!p3.finished || p3.finished_draws != settings.num_tune + settings.num_draws || p3.tuning) return std::fprintf(stderr, "run did not finish\n"), 1;

    // resume from the checkpoint in a second sampler: the remaining draws must be identical
    nuts_b200::CudaMath math2 = nuts_b200::CudaMath::normal(nchains, dim, 3.0);
    nuts_b200::Chains chains2(math2, settings, /*seed=*/7);
    chains2.restore(ckpt);
    const uint64_t rest = settings.num_tune + settings.num_draws - at;
    nuts_b200::Draws again = chains2.draw(rest);
    const bool same = rest == 0 || std::memcmp(again.data(), all.data() + at * nchains * dim, rest * nchains * dim * sizeof(double)) == 0;
    double mean = 0.0;
    const size_t tail = settings.num_draws * nchains * dim;
    for (size_t k = all.size() - tail; k < all.size(); ++k) mean += all[k];
    mean /= (double)tail;
    std::printf("paused_at %llu leapfrogs %llu divergences %llu resumed_identical %d mean %.4f\n", (unsigned long long)at,
                (unsigned long long)p3.leapfrogs, (unsigned long long)p3.divergences, (int)same, mean);
    return (same && std::fabs(mean - 3.0) < 0.05) ? 0 : 2;
  } catch (const nuts_b200::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return e.code == NUTS_ERR_NO_DEVICE ? 77 : 1;
  }
}
