// TEST INFRASTRUCTURE: compiles include/clarabel_b200.hpp and drives it like the reference's examples/rust/example_qp.rs
// (tests/basic_qp.rs data).  Run by tests/test_host_setup_cpu.py under the CUDA-runtime stand-in, where only the host
// side (construction, ABI plumbing, solution unpacking) can be exercised; on a GPU box it prints the solution.
#include <cstdio>

#include "../../include/clarabel_b200.hpp"

int main() {
  using namespace cb200;
  // basic_qp.rs:8-43: P = [4 1; 1 2] (triu), q = [1, 1], A = [-1 -1; -1 0; 0 -1; 1 1; 1 0; 0 1], b = [-1 0 0 1 0.7 0.7]
  const uint64_t Pp[] = {0, 1, 3}, Pi[] = {0, 0, 1};
  const double Px[] = {4., 1., 2.}, q[] = {1., 1.};
  const uint64_t Ap[] = {0, 4, 8}, Ai[] = {0, 1, 3, 4, 0, 2, 3, 5};
  const double Ax[] = {-1., -1., 1., 1., -1., -1., 1., 1.}, b[] = {-1., 0., 0., 1., 0.7, 0.7};
  CscMatrix P{2, 2, Pp, Pi, Px}, A{6, 2, Ap, Ai, Ax};
  try {
    DefaultSolver solver(P, q, A, b, {NonnegativeConeT(3), NonnegativeConeT(3)});
    const LinearSolverInfo li = solver.linear_solver_info();
    std::printf("created: kkt nnzA=%llu nnzL=%llu name=%s\n", (unsigned long long)li.nnzA, (unsigned long long)li.nnzL, li.name);
    solver.solve();
    std::printf("status=%d iterations=%u x=[%g, %g]\n", (int)solver.solution.status, solver.solution.iterations,
                solver.solution.x[0], solver.solution.x[1]);
    // a problem with every cone kind goes through the same constructor
    const uint64_t m2 = 2 + 3 + 3 + 3 + 3 + 3 + 3;
    std::vector<uint64_t> Ap2 = {0, 0, 0}, Ai2;
    std::vector<double> Ax2, b2(m2, 0.0);
    CscMatrix A2{m2, 2, Ap2.data(), Ai2.data(), Ax2.data()};
    DefaultSolver s2(P, q, A2, b2.data(), {ZeroConeT(2), NonnegativeConeT(3), SecondOrderConeT(3), PSDTriangleConeT(2),
                                          ExponentialConeT(), PowerConeT(0.3), GenPowerConeT({0.5, 0.5}, 1)});
    std::printf("mixed cones: ok\n");
  } catch (const SolverError& e) {
    std::printf("SolverError: %s\n", e.what());
    return 2;
  }
  return 0;
}
