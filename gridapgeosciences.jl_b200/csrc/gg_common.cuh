// Internal declarations shared by the translation units of libgeosgpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <map>
#include <memory>
#include <stdexcept>

#include "../../include/geosgpu.h"

namespace gg {

// ---- error handling: C++ exceptions inside, codes at the ABI ---------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string& m);
#define GG_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      throw gg::Error(GG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                                       ":" + std::to_string(__LINE__) + ")");                           \
  } while (0)
#define GG_REQUIRE(cond, code, msg)              \
  do {                                           \
    if (!(cond)) throw gg::Error((code), (msg)); \
  } while (0)
// wraps the body of every extern "C" entry point
#define GG_API_BEGIN try {
#define GG_API_END                         \
  }                                        \
  catch (const gg::Error& e) {             \
    gg::set_last_error(e.what());          \
    return e.code;                         \
  }                                        \
  catch (const std::bad_alloc&) {          \
    gg::set_last_error("host out of memory"); \
    return GG_ERR_NOMEM;                   \
  }                                        \
  catch (const std::exception& e) {        \
    gg::set_last_error(e.what());          \
    return GG_ERR_INVALID;                 \
  }                                        \
  return GG_OK;

// ---- device buffer ------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  explicit DevBuf(size_t n_) { alloc(n_); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  void alloc(size_t n_) {
    release();
    n = n_;
    if (n) {
      cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
      if (e != cudaSuccess) { p = nullptr; n = 0; throw Error(GG_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    }
  }
  void upload(const T* h, size_t cnt, cudaStream_t s) {
    if (cnt != n) alloc(cnt);
    if (cnt) GG_CUDA(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void download(T* h, cudaStream_t s) const {
    if (n) GG_CUDA(cudaMemcpyAsync(h, p, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    GG_CUDA(cudaStreamSynchronize(s));
  }
};

}  // namespace gg

// ---- handle types (opaque at the ABI) ---------------------------------------------------------------
namespace gg { struct FormTab; struct TilePlan; }
struct gg_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // second stream + events for overlapping the HBM-bound CSR fill with the FP64-bound element kernel
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_batch[8] = {nullptr}, ev_join = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  // device tabulations of reference bases at quadrature points, keyed by (kind, order, nq)
  std::map<uint64_t, std::shared_ptr<gg::FormTab>> tabs;
  // scratch: entry-major element matrix / vector buffers, grown on demand and reused across calls
  gg::DevBuf<double> ebuf, vbuf;
  // (order, nq) currently resident in the __constant__ tables of the fast scalar kernels
  int const_order = -1, const_nq = -1;
  // tables resident in the __constant__ blocks of the matrix-free PCG (gg_mf.cu): nq of the weights, (order, nq) keys
  int64_t mf_key_w = -1, mf_key_rt = -1, mf_key_sc = -1;
  gg::DevBuf<double> swe_lane_tab;   // the same tables as one block in global memory (gg_swe_lanes.cuh)
  int64_t swe_sf_key = -1;   // (order, nq) resident in the tables of the sum-factorised shallow-water kernels (gg_swe_sf.cuh)
  int hex_order = -1;    // order of the hexahedral index tables resident in gg_forms.cu's __constant__ block
  int64_t swe_key = -1;  // (order, nq) of the shallow-water tabulations resident in gg_swe.cu's __constant__ block
  int64_t ebe_key = -1;  // (order, nq) of the pair tables resident in gg_ebe.cu's __constant__ block
  // when set, assembly / solver entry points enqueue their work and return without synchronising the stream
  bool async = false;
  bool geometry_classes = false;   // Laplace-Beltrami assembly computes one element matrix per geometry class (gg_set_geometry_classes)
  // optional per-kernel CUDA-event timing (gg_profile_*): name -> (events, count)
  bool profile = false;
  struct ProfEntry { std::string name; cudaEvent_t a, b; };
  std::vector<ProfEntry> prof;
};

namespace gg {
// RAII event pair around one launch when profiling is on
struct ProfScope {
  gg_ctx* ctx; cudaEvent_t a = nullptr, b = nullptr; const char* name;
  ProfScope(gg_ctx* c, const char* n) : ctx(c), name(n) {
    if (ctx->profile) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream); }
  }
  ~ProfScope() {
    if (a) { cudaEventRecord(b, ctx->stream); ctx->prof.push_back({name, a, b}); }
  }
};
inline void finish_call(gg_ctx* ctx) {
  if (!ctx->async) {
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) throw Error(GG_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
  }
}
}  // namespace gg

// Reference-element tables for one (order, nq) pair, host copies; see gg_refel.cu
struct gg_tables;
struct gg_plan;
struct gg_comm;

struct gg_mesh {
  gg_ctx* ctx = nullptr;
  int dim = 2;
  int n = 0;  // cells per panel edge if built by gg_mesh_cubed_sphere, else 0
  double radius = 1.0, thickness = 0.0;
  int64_t n_cells = 0, n_nodes = 0, n_edges = 0;
  // host copies (0-based)
  std::vector<int32_t> cell_nodes;     // [n_cells][4]
  std::vector<int32_t> cell_to_chart;  // [n_cells] 0-based panel
  std::vector<double> chart_coords;    // [n_cells][4][2]
  std::vector<int32_t> cell_edges;     // [n_cells][4] edge ids by first encounter
  std::vector<int8_t> cell_edge_flip;  // [n_cells][4]
  std::vector<int32_t> edge_cells;     // [n_edges][2] cells around the edge in increasing id (-1 if none)
  // device: per-cell axis-aligned chart box (SoA) and panel
  gg::DevBuf<double> d_x0, d_y0, d_hx, d_hy;
  gg::DevBuf<int32_t> d_chart;
  // distinct chart intervals [lo, lo+h] of the cells in x and in y (cells of one mesh column / row share theirs),
  // the interval id of every cell, and per quadrature size nq the tangents of the interval's abscissae
  int64_t n_xint = 0, n_yint = 0;
  gg::DevBuf<double> d_xlo, d_xh, d_ylo, d_yh;
  gg::DevBuf<int32_t> d_ix, d_iy;
  std::vector<int32_t> h_ix, h_iy;     // host copies (geometry classes of the condensed mass operators, gg_ebe.cu)
  // geometry classes: cells with the same (x-interval, y-interval) pair have the same chart box, hence the same intrinsic
  // element matrices (the metric depends on the chart point only; on the cubed sphere every class occurs on all 6 panels)
  bool classes_built = false;
  int64_t n_classes = 0;
  std::vector<int32_t> cls, cls_rep;   // cell -> class (numbered by first occurrence), class -> representative cell
  gg::DevBuf<int32_t> d_cls, d_cls_rep;
  struct TanTab { gg::DevBuf<double> tx, ty; };
  std::map<int, std::shared_ptr<TanTab>> tan_tabs;
  bool topology_built = false;
  // 3-D shell (dim == 3, src/Geometry/CubedSphereWithThicknessMesh.jl): hexahedra ordered gamma-fastest inside a panel; the
  // (alpha, beta) geometry of a cell is that of its column, a cell of the 2-D surface mesh `surface`
  int n_layers = 0;
  std::vector<int32_t> cell_nodes8;    // [n_cells][8], Gridap hex vertex order (x = gamma fastest)
  std::vector<int32_t> cell_col, cell_layer;
  std::vector<int32_t> cell_edges12, cell_faces6;  // hex topology by first encounter (built on demand)
  int64_t n_faces = 0;
  std::unique_ptr<gg_mesh> surface;
  gg::DevBuf<int32_t> d_col, d_layer;
  // per layer: int N'_i N'_j s^2 / h and int N_i N_j h; per column: scratch for the surface matrices (keyed by order, nq)
  struct GammaTab { gg::DevBuf<double> Dg, Ng, K2, M2; };
  std::map<int, std::shared_ptr<GammaTab>> gamma_tabs;
  // partitioned models (gg_dist.cu): the plan this rank-local sub-mesh (owned + ghost cells) was cut from, the peer window
  // used for halo exchanges, and the owned-cell mask (norms and integrals count owned cells only)
  gg_plan* plan = nullptr;
  gg_comm* comm = nullptr;
  std::vector<uint8_t> cell_owned;
  gg::DevBuf<uint8_t> d_cell_owned;
};

struct gg_space {
  gg_mesh* mesh = nullptr;
  int kind = 0, order = 0, constraint = 0;
  int ndofs_loc = 0;
  int64_t n_free = 0, n_dirichlet = 0;
  std::vector<int32_t> cell_dofs;  // [n_cells][ndofs_loc], 0-based, -1 = Dirichlet
  std::vector<int8_t> cell_signs;  // [n_cells][ndofs_loc] (RT) or empty
  gg::DevBuf<int32_t> d_cell_dofs; // same layout on the device
  gg::DevBuf<int8_t> d_cell_signs;
  // local dof -> tensor multi-index (Lagrangian): lex index ix + (order+1)*iy
  std::vector<int32_t> loc2lex;
  // vector H1 (GG_SPACE_CGV): per (cell, node) the master (cell, local node) across a panel interface (-1: none) and the
  // 2x2 change-of-basis blocks [n_cells][n_nodes][2][2] (gg_vec.cu)
  std::vector<int32_t> cob_master_cell;
  std::vector<int8_t> cob_master_node;
  // the same information as geometry, so that a rank's sub-model does not need the master cell itself (it may lie beyond the
  // ghost layer): panel of the master (-1: identity block) and the chart point of the node in the master's chart
  std::vector<int32_t> cob_master_panel;
  std::vector<double> cob_master_point;
  gg::DevBuf<double> d_cob;
  // groups of consecutively numbered free DOFs owned by one mesh entity: (first dof, number of groups, dofs per group);
  // empty for caller-numbered spaces (the preconditioner then degenerates to point Jacobi)
  struct BlockClass { int64_t start, count; int size; };
  std::vector<BlockClass> blocks;
  // DOF -> contributing (cell, local) lists for atomic-free vector assembly
  gg::DevBuf<int32_t> d_vptr;  // [n_free+1]
  gg::DevBuf<int32_t> d_vsrc;  // [n_entries] = local*n_cells + cell (entry-major element-vector buffer index)
  bool vec_map_built = false;
  // partitioned spaces (gg_dist.cu): local -> global DOF ids, owner rank of every local DOF, exchange lists grouped by
  // neighbour rank (ascending global id on both sides); send_remote = the receiver's local id of every sent DOF
  struct Dist {
    int64_t n_global = 0, n_owned = 0, slot_capacity = 0;
    std::vector<int64_t> l2g;
    std::vector<int32_t> owner;
    std::vector<int32_t> peers, send_ptr, send_idx, send_remote, recv_ptr, recv_idx;
    gg::DevBuf<uint8_t> d_owned;  // [n_free] 0 ghost, 1 owned, 2 owned and a ghost of some neighbour (device send lists are sorted by row)
    gg::DevBuf<int32_t> d_peers, d_send_ptr, d_send_idx, d_send_remote, d_send_peer, d_push_first, d_recv_ptr, d_recv_idx;
  };
  std::unique_ptr<Dist> dist;
};

struct gg_vec {
  gg_ctx* ctx = nullptr;
  gg::DevBuf<double> d;
};

namespace gg { struct HexTensor; }
// number of cell batches of the pipelined CSR fill (GG_PIPELINE_BATCHES, 1..8)
inline int pipeline_batches() {
  static const int n = [] {
    const char* e = getenv("GG_PIPELINE_BATCHES");
    int v = e ? atoi(e) : 2;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  return n;
}

struct gg_csr {
  gg_ctx* ctx = nullptr;
  // 3-D shell matrices: maps of the tensor-structured numeric phase (gg_forms.cu), built at first use
  std::shared_ptr<gg::HexTensor> tensor;
  gg_space* test = nullptr;
  gg_space* trial = nullptr;
  int64_t n_rows = 0, n_cols = 0, nnz = 0;
  bool skeleton = false;         // DG pattern with facet-neighbour coupling (gg_pattern_skeleton): filled by gg_assemble_advection only
  gg::DevBuf<int32_t> d_rowptr;  // [n_rows+1]
  gg::DevBuf<int32_t> d_colind;  // [nnz]
  gg::DevBuf<double> d_vals;     // [nnz]
  // numeric-phase maps: slot -> sources in the element-matrix buffer
  gg::DevBuf<int32_t> d_sptr;    // [nnz+1]
  gg::DevBuf<int32_t> d_ssrc;    // [n_pairs_kept] raw pair id = cell*(mt*mu) + (i*mu+j)
  int64_t n_pairs = 0;
  // chunked gather lists for one element-buffer layout (0 full / 1 symmetric packed, entry-major): sources of every
  // 2048 consecutive slots sorted by (round, buffer address); see gg_pattern.cu
  int cached_layout = -1;
  int n_rounds = 0;
  int64_t n_chunks = 0;
  gg::DevBuf<int32_t> d_lsrc;      // [n_pairs] element-buffer index
  gg::DevBuf<uint16_t> d_lslot;    // [n_pairs] slot index inside its chunk
  gg::DevBuf<int32_t> d_roundptr;  // [n_chunks][n_rounds+1]
  // pipelining of the numeric phase: chunks ordered by the cell batch that completes them (all of a chunk's sources
  // lie in batches <= its batch), so the CSR fill of batch b can run while the element kernel works on batch b+1
  static constexpr int N_BATCH = 8;   // upper bound; pipeline_batches() gives the count in use
  gg::DevBuf<int32_t> d_chunk_order;       // [n_chunks] chunk ids sorted by completing batch
  int64_t batch_chunk_ptr[N_BATCH + 1] = {0};
  // CSC view (gg_csr_export_csc), built at the first export: column pointers and row indices on the host, and the
  // CSC position -> CSR slot permutation on the device (values are gathered there and copied out in CSC order)
  std::vector<int64_t> h_colptr;
  std::vector<int32_t> h_rowval;
  gg::DevBuf<int32_t> d_csc_perm;
  gg::DevBuf<double> d_csc_vals;
  // fused tile assembly (gg_tile.cuh): plans per mode (0 per cell, 1 geometry classes), built at first use
  std::shared_ptr<gg::TilePlan> tiles[2];
  bool tiles_tried[2] = {false, false};
};

struct gg_solver {
  gg_csr* A = nullptr;
  double rtol = 1e-12;
  int max_iter = 1000;
  // block-Jacobi preconditioner: the DOFs owned by one mesh entity (facet / edge / cell interior) form a block;
  // per row: first row of its block, block size, offset of the row's slice of the block inverse
  gg::DevBuf<int32_t> bstart, rowoff;
  gg::DevBuf<uint8_t> bsize;
  gg::DevBuf<double> binv;
  int64_t n_blocks = 0;
  std::vector<int32_t> h_blk_start;  // per block: first row
  std::vector<uint8_t> h_blk_size;
  gg::DevBuf<int32_t> d_blk_start, d_blk_off;
  gg::DevBuf<uint8_t> d_blk_size;
  gg::DevBuf<double> r, r2, z, p, Ap, partials;
  gg::DevBuf<double> stats;  // [0] iterations of the last solve, [1] its relative residual
  int grid = 0, block = 256;
  // matrix-free operator (gg_mf.cu): when on, A only provides the space and the preconditioner blocks
  bool mf_on = false;
  int mf_nq = 0, mf_grid = 0;
  const double* mf_qcoef = nullptr;  // coefficient at the quadrature points [n_cells][Q] (scalar mass), not owned
  gg::DevBuf<double> mf_ev;          // [ndofs_loc][n_cells] element vectors of A p
};

namespace gg {
// device tabulation of a reference basis at the tensor quadrature points, cached per (kind, order, nq) on the context
struct FormTab {
  int kind, order, nq, Q, m;
  DevBuf<double> xi, w;  // 1-D rule on [0,1]
  DevBuf<double> val;    // scalar: [Q][m]      RT: [Q][m][2]
  DevBuf<double> der;    // scalar: [Q][m][2]   RT: [Q][m] (divergence)
};
std::shared_ptr<FormTab> get_tab(gg_ctx* ctx, int kind, int order, int nq);
// per-cell geometry handles for kernels that read the tangents of the 1-D abscissae from the per-(mesh, nq) tables
struct CellGeo {
  const double *hx, *hy;      // cell extents in the chart
  const int32_t *ix, *iy;     // interval ids of the cell in x and y
  const double *tx, *ty;      // tangent tables [interval][NQ]
  const int32_t* chart;       // 0-based panel
  double radius;
};
CellGeo cell_geo(gg_mesh* m, int nq);
void mesh_geometry_classes(gg_mesh* m);   // fills cls / cls_rep / n_classes (2-D meshes with a device)
// host-side construction shared with the partition plans (gg_dist.cu); ctx == nullptr gives a host-only mesh
gg_mesh* build_cubed_sphere(gg_ctx* ctx, int n, double radius, int ordering = 0);
void upload_geometry(gg_mesh* m);
gg_mesh* build_cubed_sphere_shell(gg_ctx* ctx, int n, int n_layers, double radius, double thickness, int ordering = 0);
std::vector<int32_t> hex_local_to_lex(int p);  // per-cell chart boxes, interval tables and the owned-cell mask -> device
void space_builtin_host(gg_mesh* m, int kind, int order, int constraint, gg_space* s);
void space_from_plan(gg_mesh* m, int kind, int order, gg_space* s);
void space_upload_dist(gg_space* s);
// internal entry points shared between translation units
void build_vec_map(gg_space* s);
void gather_vector(gg_space* s, const double* ebuf, double* out, double scale, int add);
void gather_matrix(gg_csr* A, int layout, const double* ebuf, double scale, int add);
void prepare_gather(gg_csr* A, int layout);
// CSR fill of the chunks completed by cell batch `batch`, enqueued on `st`
void gather_matrix_batch(gg_csr* A, int batch, const double* ebuf, double scale, int add, cudaStream_t st);
const double* element_matrices_device(gg_form form, const gg_form_args* args, int* layout);
void spmv(gg_csr* A, double alpha, const double* x, double beta, double* y);
void solver_update(gg_solver* s);
void solver_solve_async(gg_solver* s, const double* b, double scale_b, double* x);
bool mf_supported(int kind, int order, int nq);
void solver_set_matrix_free(gg_solver* s, int degree, const double* qcoef);
void solver_solve_mf_async(gg_solver* s, const double* b, double scale_b, double* x);
}  // namespace gg
