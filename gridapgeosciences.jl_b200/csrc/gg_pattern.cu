// Symbolic phase on the device: CSR sparsity pattern of a (test, trial) pair and the slot -> source lists that
// make the numeric phase an atomic-free, deterministic segmented sum.
//
// Replaces the symbolic loop + `sparse(I,J,V)` of Gridap's SparseMatrixAssembler (reference call sites:
// src/ODEs/ODEOpsFromTFEOps.jl:154-160 `allocate_matrix`, src/ODEs/DAE.jl:161-162).  Same result as
// `sparse(I,J,V)`: the union of all cell couplings, columns sorted inside each row, duplicates summed in
// cell order (radix sort is stable), structural zeros kept.
#include <cub/cub.cuh>
#include "gg_common.cuh"

namespace gg {

__global__ void k_make_pair_keys(int64_t nc, int mt, int mu, const int32_t* __restrict__ tdofs,
                                 const int32_t* __restrict__ udofs, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t tot = nc * mt * mu;
  if (t >= tot) return;
  int64_t c = t / (mt * mu);
  int ij = (int)(t - c * (mt * mu));
  int i = ij / mu, j = ij - i * mu;
  int32_t r = tdofs[c * mt + i], q = udofs[c * mu + j];
  keys[t] = (r < 0 || q < 0) ? ~0ull : (((uint64_t)(uint32_t)r << 32) | (uint32_t)q);
  vals[t] = (uint32_t)t;
}

__global__ void k_heads(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ head) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint64_t k = keys[t];
  head[t] = (k != ~0ull) && (t == 0 || keys[t - 1] != k);
}

// at every head position: column index, slot start and row of the slot; also counts valid pairs
__global__ void k_fill_slots(int64_t n, const uint64_t* __restrict__ keys, const int32_t* __restrict__ head,
                             const int32_t* __restrict__ slot_of, int32_t* __restrict__ colind,
                             int32_t* __restrict__ sptr, int32_t* __restrict__ slot_row) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (head[t]) {
    int32_t s = slot_of[t];
    uint64_t k = keys[t];
    colind[s] = (int32_t)(k & 0xffffffffu);
    slot_row[s] = (int32_t)(k >> 32);
    sptr[s] = (int32_t)t;
  }
}

// rowptr[r] = first slot whose row >= r (slot_row is sorted)
__global__ void k_rowptr(int64_t n_rows, int64_t nnz, const int32_t* __restrict__ slot_row, int32_t* __restrict__ rowptr) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (slot_row[mid] < r) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (int32_t)lo;
}

__global__ void k_first_invalid(int64_t n, const uint64_t* __restrict__ keys, int64_t* out) {
  // keys sorted; invalid keys (~0) are at the end: binary search by one thread
  if (blockIdx.x || threadIdx.x) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] != ~0ull) lo = mid + 1; else hi = mid;
  }
  *out = lo;
}

// numeric phase: vals[slot] (+)= scale * sum of the element-buffer entries of the slot, in cell order
__global__ void k_gather_slots(int64_t nnz, const int32_t* __restrict__ sptr, const int32_t* __restrict__ ssrc,
                               const double* __restrict__ ebuf, double* __restrict__ vals, double scale, int add) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  int32_t b = sptr[s], e = sptr[s + 1];
  double acc = 0.0;
  for (int32_t k = b; k < e; ++k) acc += ebuf[ssrc[k]];
  acc *= scale;
  vals[s] = add ? vals[s] + acc : acc;
}

// re-encode raw pair ids cell*(mt*mu)+ij into indices of an entry-major element buffer
//   layout 0: full       idx = ij * nc + cell
//   layout 1: symmetric  idx = tri(min(i,j), max(i,j)) * nc + cell, tri(a,b) = a*m - a(a-1)/2 + (b-a)
__global__ void k_encode_sources(int64_t n, int64_t nc, int mt, int mu, int layout, const int32_t* __restrict__ raw,
                                 int32_t* __restrict__ out) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t p = (uint32_t)raw[t];
  int mm = mt * mu;
  int64_t c = p / mm;
  int ij = (int)(p - c * mm);
  if (layout == 1) {
    int i = ij / mu, j = ij - i * mu;
    int a = min(i, j), b = max(i, j);
    ij = a * mu - (a * (a - 1)) / 2 + (b - a);
  }
  out[t] = (int32_t)((int64_t)ij * nc + c);
}

// ---- vector (DOF) source lists ------------------------------------------------------------------------
__global__ void k_make_dof_keys(int64_t nc, int m, const int32_t* __restrict__ dofs, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nc * m) return;
  int64_t c = t / m;
  int i = (int)(t - c * m);
  int32_t d = dofs[t];
  keys[t] = d < 0 ? 0xffffffffu : (uint32_t)d;
  vals[t] = (uint32_t)((int64_t)i * nc + c);  // entry-major element-vector buffer index
}

__global__ void k_dof_ptr(int64_t n_free, int64_t n, const uint32_t* __restrict__ keys, int32_t* __restrict__ vptr) {
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d > n_free) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < (uint32_t)d) lo = mid + 1; else hi = mid;
  }
  vptr[d] = (int32_t)lo;
}

__global__ void k_gather_dofs(int64_t n_free, const int32_t* __restrict__ vptr, const int32_t* __restrict__ vsrc,
                              const double* __restrict__ ebuf, double* __restrict__ out, double scale, int add) {
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n_free) return;
  double acc = 0.0;
  for (int32_t k = vptr[d]; k < vptr[d + 1]; ++k) acc += ebuf[vsrc[k]];
  acc *= scale;
  out[d] = add ? out[d] + acc : acc;
}

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

void build_vec_map(gg_space* s) {
  if (s->vec_map_built) return;
  gg_ctx* ctx = s->mesh->ctx;
  cudaStream_t st = ctx->stream;
  int64_t nc = s->mesh->n_cells;
  int m = s->ndofs_loc;
  int64_t n = nc * m;
  GG_REQUIRE(n < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "element-vector buffer exceeds Int32 indexing");
  DevBuf<uint32_t> k0(n), k1(n), v0(n), v1(n);
  if (n) {
    k_make_dof_keys<<<nblk(n, 256), 256, 0, st>>>(nc, m, s->d_cell_dofs.p, k0.p, v0.p);
    ctx->launches++;
    cub::DoubleBuffer<uint32_t> dk(k0.p, k1.p), dv(v0.p, v1.p);
    size_t tmp_bytes = 0;
    GG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, 32, st));
    DevBuf<uint8_t> tmp(tmp_bytes);
    GG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, (int)n, 0, 32, st));
    ctx->launches += 4;
    s->d_vptr.alloc(s->n_free + 1);
    k_dof_ptr<<<nblk(s->n_free + 1, 256), 256, 0, st>>>(s->n_free, n, dk.Current(), s->d_vptr.p);
    ctx->launches++;
    // valid entries are the first vptr[n_free]; keep the whole sorted value array (Dirichlet entries trail)
    s->d_vsrc.alloc(n);
    GG_CUDA(cudaMemcpyAsync(s->d_vsrc.p, dv.Current(), n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    GG_CUDA(cudaStreamSynchronize(st));
  } else {
    s->d_vptr.alloc(s->n_free + 1);
    GG_CUDA(cudaMemsetAsync(s->d_vptr.p, 0, (s->n_free + 1) * sizeof(int32_t), st));
  }
  s->vec_map_built = true;
}

void gather_vector(gg_space* s, const double* ebuf, double* out, double scale, int add) {
  build_vec_map(s);
  gg_ctx* ctx = s->mesh->ctx;
  if (s->n_free == 0) return;
  k_gather_dofs<<<nblk(s->n_free, 256), 256, 0, ctx->stream>>>(s->n_free, s->d_vptr.p, s->d_vsrc.p, ebuf, out, scale, add);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

const int32_t* sources_for_layout(gg_csr* A, int layout) {
  if (A->cached_layout == layout) return A->d_ssrc_layout.p;
  gg_ctx* ctx = A->ctx;
  int64_t n = A->n_pairs;
  A->d_ssrc_layout.alloc(n);
  if (n) {
    k_encode_sources<<<nblk(n, 256), 256, 0, ctx->stream>>>(n, A->test->mesh->n_cells, A->test->ndofs_loc,
                                                           A->trial->ndofs_loc, layout, A->d_ssrc.p, A->d_ssrc_layout.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
  }
  A->cached_layout = layout;
  return A->d_ssrc_layout.p;
}

void gather_matrix(gg_csr* A, int layout, const double* ebuf, double scale, int add) {
  const int32_t* src = sources_for_layout(A, layout);
  gg_ctx* ctx = A->ctx;
  if (A->nnz == 0) return;
  k_gather_slots<<<nblk(A->nnz, 256), 256, 0, ctx->stream>>>(A->nnz, A->d_sptr.p, src, ebuf, A->d_vals.p, scale, add);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_pattern(gg_space* test, gg_space* trial, gg_csr** out) {
  GG_API_BEGIN
  GG_REQUIRE(test && trial && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(test->mesh == trial->mesh, GG_ERR_INVALID, "test and trial spaces live on different meshes");
  gg_ctx* ctx = test->mesh->ctx;
  GG_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  int64_t nc = test->mesh->n_cells;
  int mt = test->ndofs_loc, mu = trial->ndofs_loc;
  int64_t n = nc * mt * mu;
  GG_REQUIRE(n < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "more than 2^31 cell couplings on one GPU");
  std::unique_ptr<gg_csr> A(new gg_csr);
  A->ctx = ctx; A->test = test; A->trial = trial;
  A->n_rows = test->n_free; A->n_cols = trial->n_free;
  int64_t nnz = 0, n_valid = 0;
  if (n > 0) {
    DevBuf<uint64_t> k0(n), k1(n);
    DevBuf<uint32_t> v0(n), v1(n);
    k_make_pair_keys<<<nblk(n, 256), 256, 0, st>>>(nc, mt, mu, test->d_cell_dofs.p, trial->d_cell_dofs.p, k0.p, v0.p);
    ctx->launches++;
    GG_CUDA(cudaGetLastError());
    cub::DoubleBuffer<uint64_t> dk(k0.p, k1.p);
    cub::DoubleBuffer<uint32_t> dv(v0.p, v1.p);
    size_t tmp_bytes = 0;
    GG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, 64, st));
    {
      DevBuf<uint8_t> tmp(tmp_bytes);
      GG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, (int)n, 0, 64, st));
      GG_CUDA(cudaStreamSynchronize(st));
    }
    ctx->launches += 8;
    const uint64_t* keys = dk.Current();
    DevBuf<int32_t> head(n), slot_of(n);
    DevBuf<int64_t> d_nvalid(1);
    k_heads<<<nblk(n, 256), 256, 0, st>>>(n, keys, head.p);
    k_first_invalid<<<1, 32, 0, st>>>(n, keys, d_nvalid.p);
    ctx->launches += 2;
    size_t scan_bytes = 0;
    GG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, head.p, slot_of.p, (int)n, st));
    {
      DevBuf<uint8_t> tmp(scan_bytes);
      GG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, scan_bytes, head.p, slot_of.p, (int)n, st));
      GG_CUDA(cudaStreamSynchronize(st));
    }
    ctx->launches += 2;
    int32_t last_head = 0, last_slot = 0;
    GG_CUDA(cudaMemcpyAsync(&last_head, head.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaMemcpyAsync(&last_slot, slot_of.p + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaMemcpyAsync(&n_valid, d_nvalid.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    GG_CUDA(cudaStreamSynchronize(st));
    nnz = (int64_t)last_slot + last_head;
    A->nnz = nnz;
    A->n_pairs = n_valid;
    A->d_colind.alloc(nnz);
    A->d_sptr.alloc(nnz + 1);
    DevBuf<int32_t> slot_row(nnz);
    k_fill_slots<<<nblk(n, 256), 256, 0, st>>>(n, keys, head.p, slot_of.p, A->d_colind.p, A->d_sptr.p, slot_row.p);
    ctx->launches++;
    int32_t nv32 = (int32_t)n_valid;
    GG_CUDA(cudaMemcpyAsync(A->d_sptr.p + nnz, &nv32, sizeof(int32_t), cudaMemcpyHostToDevice, st));
    A->d_rowptr.alloc(A->n_rows + 1);
    k_rowptr<<<nblk(A->n_rows + 1, 256), 256, 0, st>>>(A->n_rows, nnz, slot_row.p, A->d_rowptr.p);
    ctx->launches++;
    A->d_ssrc.alloc(n_valid);
    if (n_valid)
      GG_CUDA(cudaMemcpyAsync(A->d_ssrc.p, dv.Current(), n_valid * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    GG_CUDA(cudaGetLastError());
    GG_CUDA(cudaStreamSynchronize(st));
  } else {
    A->d_rowptr.alloc(A->n_rows + 1);
    GG_CUDA(cudaMemsetAsync(A->d_rowptr.p, 0, (A->n_rows + 1) * sizeof(int32_t), st));
    A->d_sptr.alloc(1);
    GG_CUDA(cudaMemsetAsync(A->d_sptr.p, 0, sizeof(int32_t), st));
    GG_CUDA(cudaStreamSynchronize(st));
  }
  A->d_vals.alloc(nnz);
  if (nnz) GG_CUDA(cudaMemsetAsync(A->d_vals.p, 0, nnz * sizeof(double), st));
  GG_CUDA(cudaStreamSynchronize(st));
  *out = A.release();
  GG_API_END
}

int gg_csr_info(gg_csr* A, int64_t* n_rows, int64_t* n_cols, int64_t* nnz) {
  GG_API_BEGIN
  GG_REQUIRE(A, GG_ERR_INVALID, "null matrix");
  if (n_rows) *n_rows = A->n_rows;
  if (n_cols) *n_cols = A->n_cols;
  if (nnz) *nnz = A->nnz;
  GG_API_END
}

int gg_csr_export(gg_csr* A, int64_t* rowptr, int64_t* colind, double* vals, int index_base) {
  GG_API_BEGIN
  GG_REQUIRE(A, GG_ERR_INVALID, "null matrix");
  GG_REQUIRE(index_base == 0 || index_base == 1, GG_ERR_INVALID, "index_base must be 0 or 1");
  GG_CUDA(cudaSetDevice(A->ctx->device));
  cudaStream_t st = A->ctx->stream;
  if (rowptr) {
    std::vector<int32_t> h(A->n_rows + 1);
    A->d_rowptr.download(h.data(), st);
    for (int64_t i = 0; i <= A->n_rows; ++i) rowptr[i] = (int64_t)h[i] + index_base;
  }
  if (colind && A->nnz) {
    std::vector<int32_t> h(A->nnz);
    A->d_colind.download(h.data(), st);
    for (int64_t i = 0; i < A->nnz; ++i) colind[i] = (int64_t)h[i] + index_base;
  }
  if (vals && A->nnz) A->d_vals.download(vals, st);
  GG_API_END
}

double* gg_csr_device_values(gg_csr* A) { return A ? A->d_vals.p : nullptr; }

void gg_csr_destroy(gg_csr* A) { delete A; }

}  // extern "C"
