// TEST INFRASTRUCTURE (oracle/_ref) -- not part of the shipped product.
//
// Link-time replacement for /root/reference/euler/common/random.cc:22-28.
// The reference seeds a thread_local std::default_random_engine with time(0)
// and offers no seed API, so "fixed seed" parity is only definable against a
// build in which the same engine + distribution can be seeded.  This file
// keeps the reference's declaration (euler/common/random.h: ThreadLocalRandom)
// and its engine/distribution types and only adds ref_seed()/ref_draws().
#include <cstdint>
#include <random>

#include "euler/common/random.h"

namespace {
thread_local std::default_random_engine g_engine(1);
thread_local std::uniform_real_distribution<double> g_uniform(0., 1.);
thread_local uint64_t g_draws = 0;
}  // namespace

namespace euler {
namespace common {
double ThreadLocalRandom() {
  ++g_draws;
  return g_uniform(g_engine);
}
}  // namespace common
}  // namespace euler

extern "C" void ref_seed(uint64_t seed) {
  g_engine.seed(static_cast<std::default_random_engine::result_type>(seed));
  g_uniform.reset();
  g_draws = 0;
}

extern "C" uint64_t ref_draws() { return g_draws; }

extern "C" double ref_uniform() { return euler::common::ThreadLocalRandom(); }
