"""CPU oracle for the dense-tableau simplex pivot loop — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement of the reference algorithm
(Toptachamann/Linear_Programming_Solver, `src/main/java/lpsolver/`), written so that the
CUDA path in `linear_programming_solver_b200/` can be checked against it.  It is *not* part
of the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import or execute anything in here.  The product package never
imports it and has no CPU fallback.

Tiers
-----
* Tier D (`arith.Dec15`)  — exact reference arithmetic: `java.math.BigDecimal` with
  `MathContext(15, HALF_UP)` restated with Python's `decimal` (both implement the General
  Decimal Arithmetic spec: exact result, then one rounding to 15 digits HALF_UP).
* Tier F (`arith.F64`, and the C twin in `tier_f.c`) — the same control flow in IEEE
  binary64 with separate multiply and subtract roundings (no FMA) and true division.  The
  GPU kernels are required to be bit-identical to this tier.

Parity pinning: the reference is Java and there is no JVM in the build container, so
`oracle/_ref` (a build of the reference itself) does not exist.  The restatement is pinned
against every known-answer vector the reference's own Spock specs hold for this path
(`LPStateSpec.groovy`, `LPSolverSpec.groovy`), the recorded pivot sequences in
`logs/lp_solver.log`, and the `io_files/input.txt` fixture — see `tests/test_oracle_golden.py`.
15th-digit rounding behaviour is exercised by the reference tests only through 6-decimal
results, so HALF_UP-vs-other at the 15th digit is pinned by the decimal spec, not by a
reference vector.
"""
