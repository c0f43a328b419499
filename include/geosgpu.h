/*
 * geosgpu.h -- C ABI of libgeosgpu.so: the B200 (sm_100a) drop-in for the data-parallel hot path of
 * GridapGeosciences.jl v0.7.2: cell-wise quadrature -> element matrix/vector -> sparse assembly of
 * manifold FE forms on the cubed sphere, and the explicit Runge-Kutta stage residual / mass solve.
 *
 * Every entry point below names the reference interface it replaces (paths relative to the
 * reference checkout).  The reference reaches that arithmetic through Gridap's `Assembler` and
 * `LinearSolver` abstract types (src/ODEs/ODEs.jl:33-39 imports `collect_cell_vector,
 * collect_cell_matrix, assemble_vector_add!, allocate_matrix, assemble_matrix_add!, assemble_matrix!,
 * assemble_vector!`); INTEGRATION.md shows the Julia `ccall` shim that binds these symbols.
 *
 * Conventions
 *   - plain C: opaque handles, pointers and sizes only.  No exceptions cross the boundary: every call
 *     returns GG_OK (0) or a negative error code, with a message available from gg_last_error().
 *   - the library owns all device memory behind the handles; host arrays passed IN are copied before
 *     the call returns; outputs are written into caller-allocated HOST buffers unless the function
 *     name says `_device`.
 *   - ids crossing the ABI are 1-based Int32 (Gridap `Table{Int32}`), Dirichlet DOFs are negative;
 *     CSR export is Int64 (SparseMatrixCSC{Float64,Int}), 0- or 1-based on request.
 *   - all floating point data is FP64.
 *   - one gg_ctx per process = one GPU (one process per GPU; a second gg_init on the same device fails until the first
 *     context is finalised).  Multi-GPU runs exchange halos and dot-product operands through CUDA-IPC peer windows
 *     (gg_comm_*), stored by the solver kernels themselves; no NCCL call in the data path.  Calls on one context are serialised by
 *     the caller; work is enqueued on the context's stream and the call returns after the stream
 *     has been synchronised unless stated otherwise.
 *   - there is NO CPU fallback: if no CUDA device is usable gg_init fails with GG_ERR_CUDA.
 */
#ifndef GEOSGPU_H
#define GEOSGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_OK 0
#define GG_ERR_INVALID (-1)     /* bad argument */
#define GG_ERR_CUDA (-2)        /* CUDA runtime error (message has the CUDA string) */
#define GG_ERR_UNSUPPORTED (-3) /* valid request outside what the kernels implement (never a silent fallback) */
#define GG_ERR_NOMEM (-4)
#define GG_ERR_NOCONV (-5)      /* iterative solver did not reach the requested tolerance */
#define GG_ERR_COMM (-6)        /* a wait on a peer window timed out (lost rank): results of the call are invalid */

typedef struct gg_ctx gg_ctx;
typedef struct gg_mesh gg_mesh;
typedef struct gg_space gg_space;
typedef struct gg_csr gg_csr;
typedef struct gg_vec gg_vec;
typedef struct gg_solver gg_solver;
typedef struct gg_rk gg_rk;
typedef struct gg_plan gg_plan;
typedef struct gg_comm gg_comm;

/* ---- context ------------------------------------------------------------------------------- */
int gg_init(int device_id, gg_ctx** out);
void gg_finalize(gg_ctx* ctx);
/* message of the last failing call on this thread (ctx may be NULL) */
const char* gg_last_error(gg_ctx* ctx);
/* cudaStream_t of the context (as void*), so that callers can order their own work / NCCL calls */
void* gg_stream(gg_ctx* ctx);
int gg_synchronize(gg_ctx* ctx);
/* number of kernel launches issued by this context so far (bench.py's `gpu_launches`) */
int64_t gg_launch_count(gg_ctx* ctx);
/* async != 0: assembly / SpMV entry points enqueue their kernels on gg_stream() and return without synchronising
 * (the RK stage loop and bench.py's device-resident timing); results are complete after gg_synchronize(). */
int gg_set_async(gg_ctx* ctx, int enabled);
/* Opt-in: the fused Laplace-Beltrami assembly (gg_assemble_matrix / gg_assemble_matrix_and_vector, intrinsic form, Q1 / Q2)
 * computes one element matrix per GEOMETRY CLASS -- cells with the same chart box; every (column, row) pair of the cubed sphere
 * occurs on all 6 panels -- instead of one per cell; the CSR fill and the residual read the class matrices.  Results are bit
 * for bit those of the per-cell path and are recomputed on every call.  Off by default (also GG_LB_CLASSES=1). */
int gg_set_geometry_classes(gg_ctx* ctx, int enabled);
/* measurement hooks: per-kernel CUDA-event timing on the context's stream (names: "k_lb_fast",
 * "k_lb_action_fast", "k_gather_chunks", "k_gather_dofs"), and an FP64-FMA throughput probe (the roofline
 * denominator of the FP64-bound element kernels; MEASURED_PEAKS.json has no FP64 figure). */
int gg_profile_enable(gg_ctx* ctx, int enabled);
int gg_profile_query(gg_ctx* ctx, const char* kernel_name, int64_t* count, double* total_ms);
int gg_measure_fp64_peak(gg_ctx* ctx, double* tflops);

/* ---- atlas mesh ------------------------------------------------------------------------------
 * Replaces AtlasDiscreteModel(CubedSphereMesh(radius), nref) -- src/Geometry/AtlasDiscreteModels.jl:222-229,
 * src/Geometry/AtlasGrids.jl:227-304, src/Geometry/CubedSphereMesh.jl:83-116.
 *   dim        2 (surface) or 3 (shell, src/Geometry/CubedSphereWithThicknessMesh.jl: n x n cells per panel, n_layers cells
 *              through the shell, 0 = n as the reference's isotropic refinement; hexahedra ordered gamma-fastest; CG Q1/Q2 and DG
 *              spaces and the Laplace-Beltrami form are implemented on it)
 *   n          cells per panel edge (the reference only offers n = 2^nref; any n >= 1 is accepted here)
 *   ordering   0 = one-shot parent-major, x-fastest inside a panel (AtlasGrids.jl:254-280)
 *              1 = successive `refine` calls (src/Geometry/AtlasDiscreteModels.jl:248-262,286-344): recursive, the 4 children
 *                  of a parent stay together in Gridap order BL,BR,TL,TR  (n = 2^l)
 *              2 = uniformly refined p4est forest (src/Distributed/AtlasOctreeDistributedDiscreteModels.jl:164-208): Morton
 *                  order inside each of the 6 trees -- the same cell order as 1  (n = 2^l).  dim 3: the p8est order, octant
 *                  coordinates (x, y, z) = (gamma, alpha, beta), x bit lowest (n_layers = n = 2^l); equal-count chunks of
 *                  this order (gg_partition_range) are p4est's partition (:286-294)
 * Cells [first_cell, first_cell + n_local) of the global panel-major order are instantiated (pass 0 and
 * -1 for all); a rank of the linear partition `div((i-1)*P, Nc)+1`
 * (src/Distributed/AtlasDistributedDiscreteModels.jl:31) uses its own slice, see gg_partition_range(). */
int gg_mesh_cubed_sphere(gg_ctx* ctx, int dim, int n, int n_layers, double radius, double thickness,
                         int ordering, gg_mesh** out);
/* Connectivity supplied by the caller (the Julia host owns model construction): replaces the
 * AtlasGrid fields `cell_chart_coords` / `cell_to_chart` (AtlasGrids.jl:141-182) and
 * get_cell_node_ids (AtlasGrids.jl:383-384).  cell_node_ids: [n_cells][4], 1-based.
 * cell_chart_coords: [n_cells][4][2] in Gridap vertex order BL,BR,TL,TR.  Cells must be axis-aligned
 * rectangles in chart space (true for every cubed-sphere atlas, serial or p4est). */
int gg_mesh_from_arrays(gg_ctx* ctx, int dim, int64_t n_cells, int64_t n_nodes, const int32_t* cell_to_chart,
                        const double* cell_chart_coords, const int32_t* cell_node_ids, double radius,
                        double thickness, gg_mesh** out);
int gg_mesh_info(gg_mesh* m, int64_t* n_cells, int64_t* n_nodes, int64_t* n_edges);
int gg_mesh_get_cell_nodes(gg_mesh* m, int32_t* out /* [n_cells][4], 1-based */);
int gg_mesh_get_cell_to_chart(gg_mesh* m, int32_t* out /* [n_cells], 1-based panel */);
int gg_mesh_get_chart_coords(gg_mesh* m, double* out /* [n_cells][4][2] */);
void gg_mesh_destroy(gg_mesh* m);
/* Linear partition of the reference: [first,last) 0-based cell range owned by `part` (0-based) of `nparts`. */
int gg_partition_range(int64_t n_cells, int nparts, int part, int64_t* first, int64_t* last);
/* Sub-mesh of an existing mesh: cells listed in `cells` (0-based, ascending), e.g. owned + ghost cells of a rank
 * (AtlasDistributedDiscreteModels.jl:34-66).  Node ids are kept global. */
int gg_mesh_restrict(gg_mesh* m, int64_t n_sub, const int64_t* cells, gg_mesh** out);
/* ghost cells of the owned range [first,last): cells outside it sharing a vertex with a cell inside (:34-49).
 * Pass out=NULL to query the count. */
int gg_mesh_ghost_cells(gg_mesh* m, int64_t first, int64_t last, int64_t* n_ghost, int64_t* out);

/* ---- partitioned models (multi-GPU) ------------------------------------------------------------------
 * Replace AtlasDiscreteModel(ranks, mesh, nref) (src/Distributed/AtlasDistributedDiscreteModels.jl:20-113): linear cell
 * partition `div((i-1)*P, Nc)+1` (:31), ghost layer = cells sharing a vertex with an owned cell (:32-49), and
 * PartitionedArrays' PRange / `consistent!` on DOF vectors (src/ODEs/StageOperators.jl:109,134; src/ODEs/DAE.jl:291,314).
 * One process per GPU.  The plan is a pure function of (n, nparts) computed redundantly by every rank ON THE HOST (no GPU
 * and no communication needed); the only thing ranks exchange at set-up are the 64-byte CUDA-IPC handles of their peer
 * windows.  At run time halo values and dot-product operands are stored straight into the peers' HBM over NVLink by the
 * solver kernels (no NCCL call inside a step).  DOF ownership: the rank owning the lowest-numbered cell around the DOF. */
int gg_plan_create(int n, double radius, int nparts, int part, gg_plan** out);
/* The same for the 3-D shell, AtlasOctreeDistributedDiscreteModel(ranks, CubedSphereWithThicknessMesh(r, t), l)
 * (src/Distributed/AtlasOctreeDistributedDiscreteModels.jl:164-243,286-294): cells in the order `ordering` of gg_mesh_cubed_sphere
 * (2 = p8est Morton), equal-count chunks of that order (p4est's partition of its space-filling curve), vertex-adjacent ghost
 * layer; hexahedral CG / DG / RT spaces are exchanged through the same peer windows. */
int gg_plan_create_shell(int n, int n_layers, double radius, double thickness, int ordering, int nparts, int part, gg_plan** out);
/* Opt-in alternative to the reference's cut of the 2-D sphere: "panel strips" -- the in-panel cell index (cells are panel-major)
 * is cut like the linear partition, so rank r owns the same rows of cells on all six panels.  Every rank then holds the six
 * congruent cells of a geometry class, whose condensed mass operators the RK solvers store once (the linear partition leaves a
 * rank 3, 2 or 1 of them as the slice shrinks, which is what the per-iteration cell phase of the shallow-water solves loses under
 * weak scaling).  Owned cells are not one range: gg_plan_info returns the smallest owned cell and one past the largest,
 * gg_plan_owned_cells the count.  Global results do not depend on the partition. */
int gg_plan_create_strips(int n, double radius, int nparts, int part, gg_plan** out);
int gg_plan_owned_cells(gg_plan* p, int64_t* n_owned);
void gg_plan_destroy(gg_plan* p);
int gg_plan_info(gg_plan* p, int64_t* n_cells_global, int64_t* first_owned, int64_t* last_owned, int64_t* n_local);
int gg_plan_local_cells(gg_plan* p, int64_t* out /* [n_local] global cell ids (0-based), ascending: owned + ghost */);
/* spaces are identified by (kind, order) of gg_space_builtin; slot_capacity = max local DOFs over the ranks */
int gg_plan_space_info(gg_plan* p, int kind, int order, int64_t* n_global, int64_t* n_local, int64_t* n_owned,
                       int32_t* n_neighbors, int64_t* slot_capacity);
int gg_plan_space_maps(gg_plan* p, int kind, int order, int64_t* l2g /* [n_local] 0-based global DOF */,
                       int32_t* owner /* [n_local] owner rank */);
/* k-th neighbour: DOFs sent to it (local ids, and the ids they have over there) and received from it, ascending global id */
int gg_plan_space_neighbor(gg_plan* p, int kind, int order, int k, int32_t* peer, int64_t* n_send, int64_t* n_recv,
                           int32_t* send_idx, int32_t* send_remote_idx, int32_t* recv_idx);
/* peer window of this rank: header (flags, reduction operands) + an exchange slot of slot_doubles doubles */
int gg_comm_create(gg_ctx* ctx, int rank, int world, int64_t slot_doubles, gg_comm** out);
int gg_comm_handle(gg_comm* c, uint8_t* out64 /* cudaIpcMemHandle_t */);
int gg_comm_connect(gg_comm* c, const uint8_t* handles /* [world][64], rank order */);
/* epoch of the first timed-out wait (0 = none): a lost peer is an error code, not a hung GPU */
int gg_comm_error(gg_comm* c, int64_t* error_epoch);
void gg_comm_destroy(gg_comm* c);
/* the rank's sub-model (owned + ghost cells).  Spaces, vectors and RK steppers built on it are partitioned: local
 * numbering = the global builtin numbering restricted to the local cells; gg_norm_sq / gg_integrate count owned cells only
 * (sum the results over the ranks).  The plan (and comm) must outlive the mesh. */
int gg_mesh_partitioned(gg_ctx* ctx, gg_plan* p, gg_comm* comm, gg_mesh** out);
int gg_space_global_dofs(gg_space* s, int64_t* l2g /* 1-based global DOF ids */, int32_t* owner);
/* consistent!(v): owner -> ghost update through the peer windows (collective: every rank must call it); GG_ERR_COMM after a
 * peer-window timeout */
int gg_vec_consistent(gg_space* s, gg_vec* v);

/* ---- FE spaces -------------------------------------------------------------------------------
 * Replace TestFESpace(model, ReferenceFE(lagrangian|raviart_thomas, Float64, order); conformity, constraint)
 * (call sites: test/Laplacian/LaplaceBeltrami.jl:34-35, test/Transient/TransientWaveEquation.jl:40-47). */
#define GG_SPACE_CG 0 /* lagrangian, conformity=:H1 */
#define GG_SPACE_DG 1 /* lagrangian, conformity=:L2 */
#define GG_SPACE_RT 2 /* raviart_thomas, conformity=:HDiv */
/* ReferenceFE(lagrangian, VectorValue{2,Float64}, p), conformity=:H1 on the intrinsic model: DOFs are contravariant components
 * in the chart of the entity's master cell; cells on the other side of a panel interface see them through the 2x2
 * change-of-basis blocks of src/FESpaces/GradConformingFESpaces.jl:60-113 (gg_space_change_of_basis) */
#define GG_SPACE_CGV 3
#define GG_CONSTRAINT_NONE 0
#define GG_CONSTRAINT_ZEROMEAN 1 /* constraint=:zeromean: last free DOF fixed (Gridap ZeroMeanFESpace) */
/* raviart_thomas on the 3-D shell with dirichlet_tags=["bottom_boundary","top_boundary"] and zero Dirichlet value: no DOFs on
 * the boundary faces (test/Tutorials/LinearBoussinesq.jl:55-57) */
#define GG_CONSTRAINT_NOFLUX 2
/* Built-in numbering = our restatement of Gridap's (parity unpinned). */
int gg_space_builtin(gg_mesh* m, int kind, int order, int constraint, gg_space** out);
/* Numbering owned by the caller (exact Gridap numbering when called from Julia):
 * cell_dof_ids [n_cells][ndofs_loc] 1-based, negative = Dirichlet; cell_dof_signs (RT) may be NULL (= all +1). */
int gg_space_from_arrays(gg_mesh* m, int kind, int order, int64_t n_free, const int32_t* cell_dof_ids,
                         const int8_t* cell_dof_signs, gg_space** out);
int gg_space_info(gg_space* s, int64_t* n_free, int32_t* ndofs_loc, int64_t* n_dirichlet);
int gg_space_get_cell_dofs(gg_space* s, int32_t* out /* [n_cells][ndofs_loc], 1-based */);
int gg_space_get_cell_signs(gg_space* s, int8_t* out /* [n_cells][ndofs_loc] */);
/* GG_SPACE_CGV: the 2x2 blocks coeffs = pinvJ(J_slave) . J_master^T of every (cell, node), identity away from panel interfaces
 * (GenerateChangeOfBasisMatrixMap, src/FESpaces/GradConformingFESpaces.jl:60-113) */
int gg_space_change_of_basis(gg_space* s, double* out /* [n_cells][n_nodes][2][2] */);
void gg_space_destroy(gg_space* s);

/* ---- device vectors --------------------------------------------------------------------------
 * Stand in for the Julia `Vector{Float64}` of free DOF values (and for per-quadrature-point data). */
int gg_vec_create(gg_ctx* ctx, int64_t n, gg_vec** out);
int gg_vec_set(gg_vec* v, const double* host);  /* H2D */
int gg_vec_get(gg_vec* v, double* host);        /* D2H */
int gg_vec_fill(gg_vec* v, double value);
int64_t gg_vec_size(gg_vec* v);
double* gg_vec_device_ptr(gg_vec* v);
void gg_vec_destroy(gg_vec* v);

/* ---- sparsity pattern and CSR matrix ------------------------------------------------------------
 * Replaces allocate_matrix / symbolic loop of SparseMatrixAssembler (call sites
 * src/ODEs/ODEOpsFromTFEOps.jl:154-160,385-408; src/ODEs/DAE.jl:161-162).  Built ONCE on the device
 * (radix sort + run-length of the (row, col) pairs of all cells); also builds the slot -> contributing
 * (cell, local pair) lists used by the atomic-free numeric phase.  Columns are sorted inside each row. */
int gg_pattern(gg_space* test, gg_space* trial, gg_csr** out);
/* DG pattern with facet-neighbour coupling (SkeletonTriangulation terms): every row couples its cell with the four cells across
 * its facets (test/Tutorials/AdvectionUpwinding.jl:124-131 assembles jac = a + b + s on it for the implicit march) */
int gg_pattern_skeleton(gg_space* dg_space, gg_csr** out);
/* values of jac(u, du, v) = a(du,v) + b(du,v) + s(du,v), the DG upwind advection operator (volume + skeleton terms), for the
 * ambient wind given at gg_quadrature_points [n_cells][Q][3] and gg_skeleton_points [n_cells][4][nqf][3] */
int gg_assemble_advection(gg_space* dg_space, int32_t degree, const double* velocity_cells, const double* velocity_facets, gg_csr* A);
int gg_csr_info(gg_csr* A, int64_t* n_rows, int64_t* n_cols, int64_t* nnz);
/* index_base 0 or 1.  Any of rowptr/colind/vals may be NULL to skip it. */
int gg_csr_export(gg_csr* A, int64_t* rowptr, int64_t* colind, double* vals, int index_base);
/* The same matrix as SparseMatrixCSC{Float64,Int}(m, n, colptr, rowval, nzval) -- what allocate_matrix / assemble_matrix!
 * hand back in the reference (src/ODEs/ODEOpsFromTFEOps.jl:154-160,385-408) -- for ANY test x trial pair (rectangular and
 * non-symmetric blocks such as GG_FORM_DIV, GG_FORM_VECTOR_DIV or the advection operator included): colptr [n_cols+1],
 * rowval [nnz] with rows ascending inside each column, nzval [nnz].  The structure is built at the first call; later calls
 * with colptr = rowval = NULL only permute the current values on the device and copy them out. */
int gg_csr_export_csc(gg_csr* A, int64_t* colptr, int64_t* rowval, double* nzval, int index_base);
double* gg_csr_device_values(gg_csr* A);
/* y = alpha * A x + beta * y on the device */
int gg_csr_spmv(gg_csr* A, double alpha, gg_vec* x, double beta, gg_vec* y);
void gg_csr_destroy(gg_csr* A);

/* ---- forms -----------------------------------------------------------------------------------
 * Gridap forms are Julia closures interpreted as lazy operation trees; the library implements the forms of
 * the reference's drivers as hand-written kernels selected by id (the documented deviation). */
typedef enum {
  /* a(u,v) = int grad(v).(ginv.grad(u)) sqrtg   -- test/Laplacian/LaplaceBeltrami.jl:47 */
  GG_FORM_LAPLACE_BELTRAMI = 1,
  /* a(u,v) = int grad_G(v).grad_G(u) dG on the extrinsic model (2x3 Jacobian, pseudo-determinant,
   * pseudo-inverse) -- test/AmbientModel/AmbientLaplaceBeltrami.jl:44, src/Geometry/AtlasGrids.jl:374-376 */
  GG_FORM_LAPLACE_BELTRAMI_EXTRINSIC = 2,
  /* a(u,v) = int u v sqrtg - int grad(v).(ginv.grad(u)) sqrtg -- test/Laplacian/Helmholtz.jl:43 */
  GG_FORM_HELMHOLTZ = 3,
  /* m(p,q) = int q p sqrtg -- scalar block of test/Transient/TransientWaveEquation.jl:63 */
  GG_FORM_MASS_SCALAR = 4,
  /* m(u,v) = int (v.(g u)) / sqrtg -- RT block of test/Transient/TransientWaveEquation.jl:63 */
  GG_FORM_MASS_RT = 5,
  /* b(u,q) = int q div(u)  (rows: scalar test, cols: RT trial) -- test/Transient/TransientWaveEquation.jl:64 */
  GG_FORM_DIV = 6,
  /* l(v) = int f v sqrtg, f given at the quadrature points -- test/Laplacian/LaplaceBeltrami.jl:53 */
  GG_FORM_SOURCE = 7,
  /* vector H1: a(u,v) = int (u.(g v)) sqrtg and l(v) = int (f.(g v)) sqrtg with the contravariant field f given at the
   * quadrature points, qdata [n_cells][Q][2] (test/Projection/L2_projection_Lagrangian_vector.jl:41-42) */
  GG_FORM_MASS_VECTOR = 8,
  GG_FORM_SOURCE_VECTOR = 9,
  /* 3-D shell: l(v) = int_G ((ginv grad(u_h)).n_G) v sqrtg dG over the bottom and top boundaries, u_h = args->u -- the boundary
   * term of the integration by parts in the 3-D run of test/Laplacian/LaplaceBeltrami.jl:58-64 (u_h = interpolate(f, U)) */
  GG_FORM_SHELL_BOUNDARY_FLUX = 10,
  /* vector H1: a(u,v) = int (D1(u) D1(v) - D2(u) D2(v)) / sqrtg with D1(u) = div(sqrtg u) ("divg_product_rule") and
   * D2(u) = div(R g u) = -curl(g u) ("divergenceRgu") -- test/SurfaceStokes/VectorLaplacian2D.jl:38-50,71-73 and the viscous
   * block biform_curl of test/SurfaceStokes/SurfaceStokes_TaylorHood.jl:100-102 (nu through args->scale) */
  GG_FORM_VECTOR_LAPLACIAN = 11,
  /* b(u,q) = int q D1(u) (rows: Lagrangian test, columns: vector H1 trial) -- biform_p of
   * test/SurfaceStokes/SurfaceStokes_TaylorHood.jl:106; biform_u's pressure term is minus its transpose (:104) */
  GG_FORM_VECTOR_DIV = 12
} gg_form;

typedef struct {
  gg_space* test;
  gg_space* trial;   /* NULL for linear forms */
  int32_t degree;    /* Measure(Omega, degree): tensor Gauss-Legendre with ceil((degree+1)/2) points per direction */
  gg_vec* u;         /* DOF vector of the trial space for residual / action forms (free values), or NULL */
  gg_vec* qdata;     /* per-quadrature-point coefficient [n_cells][Q] (x-fastest inside the cell), or NULL */
  double dirichlet;  /* value of the (single, zeromean) Dirichlet DOF when gathering u */
  double scale;      /* result is multiplied by this (1.0 if 0 is passed) */
} gg_form_args;

/* assemble_matrix! / assemble_matrix_add! (src/ODEs/ODEs.jl:36-38): numeric phase into the pattern of A. */
int gg_assemble_matrix(gg_form form, const gg_form_args* args, gg_csr* A, int add);
/* assemble_vector! / assemble_vector_add! (src/ODEs/ODEs.jl:35,39): b (size n_free of the test space).
 * For bilinear forms this is the residual / matrix-free action b_i = a(u_h, phi_i) with u_h = args->u. */
int gg_assemble_vector(gg_form form, const gg_form_args* args, gg_vec* b, int add);
/* Jacobian and residual in ONE pass over the cells (shared geometry): A = a(.,.), b_i = a(u_h, phi_i). */
int gg_assemble_matrix_and_vector(gg_form form, const gg_form_args* args, gg_csr* A, gg_vec* b);
/* parity hook: dense element matrices of cells [first, first+n) -> host out[n][m_test][m_trial] */
int gg_element_matrices(gg_form form, const gg_form_args* args, int64_t first_cell, int64_t n, double* out);
/* parity hook: element vectors of cells [first, first+n) -> host out[n][m_test] */
int gg_element_vectors(gg_form form, const gg_form_args* args, int64_t first_cell, int64_t n, double* out);
/* chart / ambient coordinates of the quadrature points: out_chart [n_cells][Q][2], out_xyz [n_cells][Q][3]
 * (AmbientMapCellField, src/CellData/CellFields.jl:207-212); either may be NULL */
int gg_quadrature_points(gg_mesh* m, int32_t degree, double* out_chart, double* out_xyz);
/* ambient coordinates of the facet quadrature points (Gauss-Legendre, nq = ceil((degree+1)/2) per facet) of the four sides
 * of every cell, xyz [n_cells][4][nq][3] (local edges in Gridap order, points from the first to the second vertex):
 * AmbientMapCellField on SkeletonTriangulation(model) (src/CellData/CellFields.jl:207-260) */
int gg_skeleton_points(gg_mesh* m, int32_t degree, int32_t* n_pts, double* xyz);
/* sum(int f sqrtg dOmega) with f at the quadrature points (NULL = 1: the surface area of
 * test/Geometry/SurfaceArea.jl:19-24) */
int gg_integrate(gg_mesh* m, int32_t degree, gg_vec* qdata, double* out);

/* ---- geometry parity hooks (device evaluation of src/Fields/CubedSphere{Map,Metric,InvMetric}.jl) ---- */
/* pts [n][2] chart points of panel `panel` (1-based); outputs (any may be NULL):
 * xyz [n][3], jac [n][2][3] (J[i][k] = d phi_k / d x_i), metric [n][3] (g11,g12,g22), inv_metric [n][3], sqrtg [n] */
int gg_eval_geometry(gg_ctx* ctx, int panel, double radius, int64_t n, const double* pts, double* xyz,
                     double* jac, double* metric, double* inv_metric, double* sqrtg);

/* ---- linear solver (mass solves of the RK stages) ------------------------------------------------
 * Stands in for LUSolver() / CGSolver(JacobiLinearSolver(); rtol=1e-12) (test/Tutorials/ShallowWater.jl:175-176)
 * at the `Algebra.solve!(x, ::LinearSolver, op, cache)` seam (src/ODEs/StageOperators.jl:55-88). */
/* On a partitioned space with connected peer windows (gg_mesh_partitioned + gg_comm_connect) the solve is collective: dot products
 * over the owned rows are all-reduced and the owner -> ghost update runs through peer memory inside the kernel; x comes back
 * consistent (owned + ghost values).  GG_ERR_COMM after a peer-window time-out. */
int gg_solver_pcg_create(gg_csr* A, double rtol, int max_iter, gg_solver** out);
/* numerical_setup!: refresh the Jacobi preconditioner after A's values changed */
int gg_solver_update(gg_solver* s);
int gg_solver_solve(gg_solver* s, gg_vec* b, gg_vec* x, int32_t* iters, double* rel_res);
void gg_solver_destroy(gg_solver* s);

/* ---- interpolation and norms (initial conditions / error measures of the transient drivers) ---------
 * Replace `interpolate(f o ambient_map_cf, P)` for Lagrangian spaces and, for Raviart-Thomas spaces,
 * `interpolate(meas_cf*((pinvJ o transpose o grad(ambient_map_cf)) . (vX o ambient_map_cf)), U)`
 * (test/Transient/TransientShallowWater.jl:88-95, src/Helpers/Operators.jl:16-24): the caller evaluates its
 * ambient function at the points returned by gg_space_interp_points(); the pull-back sqrtg * pinvJ(J) . v,
 * the inverse contravariant Piola map and the reference moments run on the device.
 *   n_pts   points per cell (Lagrangian: the ndofs_loc nodes; RT_k: 4 (k+2) facet + (k+2)^2 interior moment points; hexahedral
 *           RT_k on the shell: 6 (k+2)^2 face + interior moment points)
 *   chart   [n_cells][n_pts][dim] or NULL;  xyz [n_cells][n_pts][3] or NULL */
int gg_space_interp_points(gg_space* s, int32_t* n_pts, double* chart, double* xyz);
/* vals: Lagrangian [n_cells][n_pts] scalars; RT [n_cells][n_pts][3] ambient vectors.  out: free DOF values. */
int gg_space_interpolate(gg_space* s, const double* vals, gg_vec* out);
/* squared norms used by the drivers' error measures (test/Transient/TransientWaveEquation.jl:112-117):
 * Lagrangian: sum(int e^2 sqrtg);  RT: sum(int e.(g e) / sqrtg), quadrature degree `degree`. */
int gg_norm_sq(gg_space* s, gg_vec* e, int32_t degree, double* out);
/* vector H1 spaces: sum(int grad(e) : (ginv . grad(e)) sqrtg), the gradient part of eu_h1
 * (test/SurfaceStokes/SurfaceStokes_TaylorHood.jl:122, VectorLaplacian2D.jl:91) */
int gg_h1_seminorm_sq(gg_space* s, gg_vec* e, int32_t degree, double* out);

/* Delta_s f of an AMBIENT function f at the quadrature points, closed form (use_automatic_differentiation=false branch of
 * Δs(f, Ω), src/CellData/SurfaceDiffOps.jl:26-65,78-82; Hessian of the map src/Fields/CubedSphereMap.jl:97-123, metric
 * derivatives CubedSphereMetric.jl:52-71, CubedSphereInvMetric.jl:37-56).  The caller evaluates grad_X f [n_cells][Q][3] and
 * hess_X f [n_cells][Q][3][3] at the ambient points of gg_quadrature_points; out has n_cells * Q entries and can be passed
 * as `qdata` of GG_FORM_SOURCE (the manufactured right-hand side of test/Laplacian/LaplaceBeltrami.jl:41-53). */
int gg_surface_laplacian(gg_mesh* m, int32_t degree, const double* grad_f, const double* hess_f, gg_vec* out);
/* Contravariant components of the surface gradient, ginv . (J grad_X f): out [n_cells][Q][2]
 * (∇s(f, Ω), src/CellData/SurfaceDiffOps.jl:107-115,127-133) */
int gg_surface_gradient(gg_mesh* m, int32_t degree, const double* grad_f, gg_vec* out);
/* Surface divergence of an ambient vector field, 1/sqrtg div(sqrtg ginv J F): F [n_cells][Q][3],
 * grad_F [n_cells][Q][3][3] with grad_F[l][k] = dF_k/dX_l (Gridap's convention); out [n_cells][Q]
 * (divs(f, Ω), src/CellData/SurfaceDiffOps.jl:228-259) */
int gg_surface_divergence(gg_mesh* m, int32_t degree, const double* F, const double* grad_F, gg_vec* out);
/* sum(int (u_h - f)^2 sqrtg) with f at the quadrature points (NULL = 0): the error measure of
 * test/Laplacian/LaplaceBeltrami.jl:82-83; `dirichlet` is the value of the constrained DOF of a zeromean space.
 * Raviart-Thomas spaces: f holds the two CONTRAVARIANT components of the exact field per point and the measure is
 * sum(int (u_h/sqrtg - f).g.(u_h/sqrtg - f) sqrtg), the velocity error of test/Laplacian/HodgeLaplacian_scalar.jl:104-105 */
int gg_l2_error_sq(gg_space* s, gg_vec* u, double dirichlet, gg_vec* fq, int32_t degree, double* out);
/* (int u_h d xi) / (int d xi) in the CHART measure: the shift Gridap's ZeroMeanFESpace subtracts after a solve
 * (FEFunction(::ZeroMeanFESpace, ...): c = -(vol_i . x) / vol), used by test/Laplacian/LaplaceBeltrami.jl:71 */
int gg_chart_mean(gg_space* s, gg_vec* u, double dirichlet, double* out);

/* ---- explicit Runge-Kutta drivers ----------------------------------------------------------------
 * GG_PROBLEM_WAVE: replaces Gridap.ODEs `ode_march!` for EXRungeKutta on the linear-wave operator
 * (test/Transient/TransientWaveEquation.jl:63-77; constant mass).
 * GG_PROBLEM_SHALLOW_WATER: replaces the DAE-aware `ode_march!` of src/ODEs/RungeKuttaEX.jl:48-134 on the rotating
 * shallow-water operator of test/Transient/TransientShallowWater.jl:113-154 (prognostic u in RT_k, p in DG_k;
 * diagnostic q in CG_{k+1}, F in RT_k, Phi in DG_k): a diagnostic solve (src/ODEs/DAE.jl:241-274) before every
 * stage and after the update, then slope_i = -M^-1 r (src/ODEs/StageOperators.jl:55-88).
 * The whole step runs on the device without host synchronisation. */
#define GG_PROBLEM_WAVE 1
#define GG_PROBLEM_SHALLOW_WATER 2
/* DG upwind advection of a scalar, volume + skeleton terms (test/Tutorials/AdvectionUpwinding.jl:107-137), Q = DG-Q_order,
 * marched with the explicit tableau (BASELINE configs[2]); the state is [p] alone (n_u = 0) */
#define GG_PROBLEM_ADVECTION 3
/* Thermal shallow water (test/Tutorials/ThermalShallowWater.jl:170-252): state [u; p; B] with B in DG_k (gg_rk_info reports
 * n_p = 2 x the DG size), diagnostics [q; F; Phi; b] (gg_rk_diag_info reports n_Phi = 2 x the DG size); same DAE stepper and
 * q0 handling as GG_PROBLEM_SHALLOW_WATER, plus the buoyancy cell terms and the skeleton terms u_s1, u_s2, B_s1, B_s2
 * (upwinding_sign threshold 1e-4).  `gravity` and `topography` are not used (B carries gravity). */
#define GG_PROBLEM_THERMAL_SHALLOW_WATER 4
/* Linearised Boussinesq equations on the 3-D shell (test/Tutorials/LinearBoussinesq.jl:112-134), RT_k x DG_k x DG_k on hexahedra
 * with no flux through the bottom and top boundaries: state [u; p; b] (gg_rk_info reports n_p = 2 x the DG size), constant mass,
 * marched with the explicit tableau.  `coriolis` holds the rotation rate omega at the quadrature points [n_cells][nq^3] (gamma
 * fastest), `sound_speed` c and `buoyancy_frequency` N the two constants (1 if 0).  The mesh must be a builtin 3-D shell. */
#define GG_PROBLEM_BOUSSINESQ 5
#define GG_Q0_ALIAS 0         /* q0 aliases q, as in the reference (RungeKuttaEX.jl:69-75 pass one array twice) */
#define GG_Q0_START_OF_STEP 1 /* q0 = potential vorticity at the beginning of the step */
typedef struct {
  int32_t problem;     /* GG_PROBLEM_* */
  int32_t order;       /* p_fe: RT_p x DG_p */
  int32_t degree;      /* quadrature degree */
  int32_t n_stages;
  const double* A;     /* [n_stages][n_stages] row-major, strictly lower triangular */
  const double* b;     /* [n_stages] */
  const double* c;     /* [n_stages] */
  double dt;
  double rtol;         /* PCG tolerance of the mass solves (1e-12 if 0) */
  /* shallow water only */
  double gravity;           /* g_r of the Bernoulli potential */
  double tau;               /* upwinding parameter of res_u; negative: dt/2 (TransientShallowWater.jl:150) */
  int32_t q0_mode;          /* GG_Q0_* */
  const double* coriolis;   /* host [n_cells][Q]: f at the quadrature points (cor_cf) */
  const double* topography; /* host [n_cells][Q]: b at the quadrature points, or NULL (= 0) */
  /* DG advection only: the ambient wind beta_x (3 components) at the cell quadrature points [n_cells][Q][3]
   * (gg_quadrature_points) and at the facet quadrature points of every cell side [n_cells][4][nq][3] (gg_skeleton_points) */
  const double* velocity_cells;
  const double* velocity_facets;
  /* linear Boussinesq only */
  double sound_speed;
  double buoyancy_frequency;
} gg_rk_args;
int gg_rk_create(gg_mesh* m, const gg_rk_args* args, gg_rk** out);
int gg_rk_info(gg_rk* rk, int64_t* n_u, int64_t* n_p);
int gg_rk_set_state(gg_rk* rk, const double* x /* [n_u + n_p] host */);
int gg_rk_get_state(gg_rk* rk, double* x);
double* gg_rk_state_device_ptr(gg_rk* rk);
/* returns GG_ERR_NOCONV if a mass / diagnostic solve of the march stopped at max_iter, GG_ERR_COMM after a peer-window timeout */
int gg_rk_step(gg_rk* rk, int32_t n_steps);
/* matrix-free stage residual r = res(x) (parity hook), host in/out.  Shallow water: the diagnostics are solved
 * at x first and q0 aliases q. */
int gg_rk_residual(gg_rk* rk, const double* x, double* r);
int gg_rk_stats(gg_rk* rk, int64_t* pcg_iters_total, int64_t* solves_total);
/* measurement hook: time (ns, summed over all solves so far) the matrix-free PCG of mass solver `which` (0: RT mass,
 * 1: potential-vorticity mass) spent in its three phases -- out[0] cell-wise operator application, out[1] DOF gather +
 * dot product, out[2] vector updates + block preconditioner -- and out[3] the iterations they cover (read with
 * %globaltimer inside the persistent kernel).  Zeros when that solver runs on the assembled matrix. */
int gg_rk_solver_phase_ns(gg_rk* rk, int which, double* out);
/* shallow water: sizes of the diagnostic blocks y = [q; F; Phi] */
int gg_rk_diag_info(gg_rk* rk, int64_t* n_q, int64_t* n_F, int64_t* n_Phi);
/* diagnostics of the last solve (after gg_rk_step: those of the final state), host out [n_q + n_F + n_Phi] */
int gg_rk_get_diagnostics(gg_rk* rk, double* y);
/* parity hooks: y = diagnostics(x) (src/ODEs/DAE.jl:241-274); r = res_x(x, y, q0) + mass(0) (StageOperators.jl:24-38),
 * all host arrays; q0 has n_q entries */
int gg_rk_diagnostics(gg_rk* rk, const double* x, double* y);
int gg_rk_swe_residual(gg_rk* rk, const double* x, const double* y, const double* q0, double* r);
void gg_rk_destroy(gg_rk* rk);

#ifdef __cplusplus
}
#endif
#endif /* GEOSGPU_H */
