"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A numpy/scipy (and, in `oracle/c/`, plain-C) restatement of the reference's
algorithm for the cubed-sphere assembly / explicit-RK hot path of
GridapGeosciences.jl v0.7.2 (`/root/reference`, pure Julia on Gridap 0.20.9).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import anything from this package, and only as the checker / the CPU baseline.
The product (`gridapgeosciences.jl_b200/`) never imports it.

PARITY STATUS
  * pinned against the reference's own known-answer tests (see tests/test_oracle_*.py):
    metric == J J^T, g g^-1 == I, panel independence (MetricFieldsTests.jl:93-122),
    forward/inverse map identity (ForwardInverseMapTests.jl:20-40), corners on the sphere
    (CubedSphereTests.jl:60-69), area = 4 pi r^2 (SurfaceArea.jl:25-38), intrinsic == extrinsic
    (AmbientLaplaceBeltramiTests.jl:28-30), L2 convergence slope >= p (LaplaceBeltrami.jl:101-105),
    and the wave / shallow-water end-state regression numbers
    (TransientWaveEquation.jl:128-129, TransientShallowWater.jl:208-209).
  * PARITY UNPINNED (no reference test, no runnable reference: Julia is absent and Gridap is not
    vendored): element-matrix entries, CSR pattern, DOF numbering, RT basis/sign conventions.
    These follow our restatement of Gridap 0.20 conventions and are flagged
    where they are implemented.
"""
