# baseline/run_reference.jl -- NOT RUN IN THIS REPOSITORY'S BUILD ENVIRONMENT (no Julia, no network).
#
# Times the reference's own CPU path for bench.py's headline metric (cells/s assembled, residual + Jacobian, Laplace-Beltrami,
# CG-Q2, quadrature degree 18) so that a real Gridap number can be put into BASELINE.md next to the CPU port that
# `bench.py --impl reference` times.  Serial:  julia --project=... baseline/run_reference.jl 6
# MPI:  mpiexec -n N julia --project=... baseline/run_reference.jl 6   (with the distributed model constructor of
# test/Laplacian/mpi/LaplaceBeltramiTests.jl:11-14)
using Gridap, Gridap.FESpaces, GridapGeosciences

nref = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 6      # 2^8 = 256 per panel edge is the headline size; 6 keeps the run short
p, degree = 2, 18
model = AtlasDiscreteModel(CubedSphereMesh(1.0), nref; manifold_style = IntrinsicManifold())
Ω = Triangulation(model); dΩ = Measure(Ω, degree)
V = TestFESpace(model, ReferenceFE(lagrangian, Float64, p); conformity = :H1); U = TrialFESpace(V)
inv_metric_cf, meas_cf = InvMetricCellField(Ω), MeasureCellField(Ω)
a(u, v) = ∫((∇(v) ⋅ (inv_metric_cf ⋅ ∇(u))) * meas_cf)dΩ
uh = FEFunction(U, randn(num_free_dofs(U)))
assem = SparseMatrixAssembler(U, V)
A = assemble_matrix(a, assem, U, V)                        # warm-up (compilation) + pattern
b = assemble_vector(v -> a(uh, v), assem, V)
nsteps = 3
t = @elapsed for _ in 1:nsteps
  assemble_matrix!(a, A, assem, U, V)
  assemble_vector!(v -> a(uh, v), b, assem, V)
end
println("cells/s assembled (residual+Jacobian): ", num_cells(model) * nsteps / t, "  (", num_cells(model), " cells, ", Threads.nthreads(), " thread)")
