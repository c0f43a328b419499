# tools/dump_reference.jl -- NOT RUN IN THIS REPOSITORY'S BUILD ENVIRONMENT (no Julia, no network).
#
# Wherever Julia 1.10 + GridapGeosciences v0.7.2 (Gridap 0.20.9) are installed, this script dumps from the REAL reference the
# quantities that no reference test pins (SURVEY.md §8c): DOF numbering, CSC pattern and values of the Laplace-Beltrami
# matrix, RT signs, a few element matrices.  The arrays use the same names as tests/golden/make_golden.py, so that
# tests/test_golden.py then checks the CUDA path against Gridap itself instead of against the CPU restatement.
#
#   julia --project=/path/to/GridapGeosciences.jl tools/dump_reference.jl tests/golden/reference_laplace_beltrami_n4.npz
using Gridap, Gridap.FESpaces, Gridap.Geometry, Gridap.CellData, GridapGeosciences, SparseArrays, NPZ

out = length(ARGS) >= 1 ? ARGS[1] : "reference_laplace_beltrami_n4.npz"
nref, radius = 2, 1.0                                 # n = 2^nref = 4 cells per panel edge
model = AtlasDiscreteModel(CubedSphereMesh(radius), nref; manifold_style = IntrinsicManifold())
Ω = Triangulation(model)
inv_metric_cf, meas_cf = InvMetricCellField(Ω), MeasureCellField(Ω)
d = Dict{String,Any}()
grid = get_grid(model)
d["cell_nodes"] = permutedims(reduce(hcat, [collect(Int64, c) for c in get_cell_node_ids(grid)])) .- 1
for (p, degree) in ((1, 12), (2, 18))
  dΩ = Measure(Ω, degree)
  V = TestFESpace(model, ReferenceFE(lagrangian, Float64, p); conformity = :H1)
  U = TrialFESpace(V)
  a(u, v) = ∫((∇(v) ⋅ (inv_metric_cf ⋅ ∇(u))) * meas_cf)dΩ          # test/Laplacian/LaplaceBeltrami.jl:47
  ids = get_cell_dof_ids(V)
  d["q$(p)_cell_dofs"] = permutedims(reduce(hcat, [collect(Int64, c) for c in ids])) .- 1
  du, dv = get_trial_fe_basis(U), get_fe_basis(V)
  Ke = collect(get_array(a(du, dv)))                                 # element matrices, Gridap local order
  d["q$(p)_Ke_first8"] = permutedims(cat(Ke[1:8]...; dims = 3), (3, 1, 2))
  A = assemble_matrix(a, U, V)                                       # SparseMatrixCSC; symmetric pattern => CSR == CSC
  d["q$(p)_indptr"] = A.colptr .- 1
  d["q$(p)_indices"] = A.rowval .- 1
  d["q$(p)_data"] = A.nzval
end
npzwrite(out, d)
println("wrote ", out)
