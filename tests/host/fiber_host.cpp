// tests/host/fiber_host.cpp -- TEST INFRASTRUCTURE: exercises stan_b200/cpp/b200/fiber.hpp on its own (no GPU, no
// reference headers): many fibers in one thread, yields from deep recursion, exceptions thrown and caught inside
// fibers while other fibers are suspended inside try blocks, floating-point state carried across switches.
#include <b200/fiber.hpp>

#include <cmath>
#include <cstdio>
#include <memory>
#include <stdexcept>
#include <vector>

using b200::fiber;

// a recursive "tree builder" that yields at every leaf, like build_tree -> evolve
static long build(int depth, long& leaves, int id) {
  if (depth == 0) {
    ++leaves;
    fiber::yield();
    if ((leaves + id) % 97 == 0)
      throw std::domain_error("leaf rejected");
    return 1;
  }
  volatile char pad[512];   // some stack per frame
  pad[0] = static_cast<char>(depth);
  long n = 0;
  try {
    n += build(depth - 1, leaves, id);
    n += build(depth - 1, leaves, id);   // yields happen inside this try block in other fibers meanwhile
  } catch (const std::domain_error&) {
    n += 1000000;                        // handled here, then the tree goes on
  }
  return n + pad[0] - depth;
}

extern "C" int fiber_selftest(int n_fibers, int depth, long* total_leaves, long* total_value, double* fp_check) {
  std::vector<long> leaves(n_fibers, 0), value(n_fibers, 0);
  std::vector<double> fp(n_fibers, 0.0);
  std::vector<std::unique_ptr<fiber>> fibers;
  for (int i = 0; i < n_fibers; ++i)
    fibers.emplace_back(new fiber([&, i] {
      double acc = 0.0;
      for (int rep = 0; rep < 3; ++rep) {
        value[i] += build(depth, leaves[i], i);
        acc += std::sqrt(static_cast<double>(i + rep + 1));   // fp registers / control words survive the switches
        fiber::yield();
      }
      fp[i] = acc;
      if (i % 5 == 0)
        throw std::runtime_error("a chain that fails at its very end");   // swallowed at the bottom of the fiber
    }));
  long sweeps = 0;
  for (;;) {
    int live = 0;
    for (auto& f : fibers) {
      if (f->done())
        continue;
      f->resume();
      live += !f->done();
    }
    ++sweeps;
    if (!live)
      break;
  }
  *total_leaves = 0;
  *total_value = 0;
  *fp_check = 0;
  for (int i = 0; i < n_fibers; ++i) {
    *total_leaves += leaves[i];
    *total_value += value[i];
    *fp_check += fp[i];
  }
  return static_cast<int>(sweeps);
}
