// oracle/ref_rng_hook.h — TEST INFRASTRUCTURE.
// Force-included (-include) in front of every reference translation unit when
// building oracle/_ref/libref_oracle.so.  The reference draws its randoms from
// `static std::mt19937 RNGS[32]` seeded by std::random_device
// (global.hpp:14,42-53), which makes castRay non-reproducible.  Without
// touching the sources, the name `mt19937` is redirected to an engine whose
// words come from the harness (scripted list, Philox sample stream, or a real
// free-running Mersenne twister), so the REAL castRay/Material/BVH code can be
// replayed on exactly the uniforms the GPU path consumes.
#pragma once
#include <cstdint>
#include <random>

extern "C" uint32_t b2pt_oracle_next_u32();  // defined in ref_harness.cpp

namespace std {
typedef mt19937 b2pt_real_mt19937;
struct b2pt_hooked_engine {
    typedef uint32_t result_type;
    b2pt_hooked_engine() {}
    explicit b2pt_hooked_engine(unsigned) {}
    static constexpr result_type min() { return 0u; }
    static constexpr result_type max() { return 0xFFFFFFFFu; }
    result_type operator()() { return b2pt_oracle_next_u32(); }
};
}  // namespace std
#define mt19937 b2pt_hooked_engine
