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
  ProfScope ps(ctx, "k_gather_dofs");
  k_gather_dofs<<<nblk(s->n_free, 256), 256, 0, ctx->stream>>>(s->n_free, s->d_vptr.p, s->d_vsrc.p, ebuf, out, scale, add);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// ---- chunked numeric phase ---------------------------------------------------------------------------------
// The slot-by-slot gather above reads the entry-major element buffer with one 8-byte access per lane at 32
// unrelated addresses (consecutive slots of a CSR row live in different planes), which makes it L1-wavefront
// bound at ~20% of HBM bandwidth.  The chunked variant re-sorts, once per (matrix, buffer layout), the sources of
// every CHUNK consecutive slots by (round, buffer address): a CTA then streams its sources with coalesced loads
// (consecutive cells of one plane), accumulates them in shared memory and writes its slots with coalesced stores.
// "round" k holds the k-th contribution of each slot (cell order), rounds are separated by a barrier, so every slot
// is still summed in the same fixed order: deterministic and atomic-free.
constexpr int GATHER_CHUNK = 2048;

// layout 2: symmetric packed planes with one column per geometry CLASS (stride ncls, column cls[cell]) instead of per cell
__global__ void k_chunk_keys(int64_t n, int64_t nnz, int64_t nc, int mt, int mu, int layout, const int32_t* __restrict__ sptr,
                             const int32_t* __restrict__ raw, uint64_t* __restrict__ keys, uint16_t* __restrict__ lslot,
                             int64_t ncls, const int32_t* __restrict__ cls) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  // slot of source position t: last slot with sptr[slot] <= t
  int64_t lo = 0, hi = nnz;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (sptr[mid] <= t) lo = mid + 1; else hi = mid;
  }
  int64_t slot = lo - 1;
  int round = (int)(t - sptr[slot]);
  uint32_t p = (uint32_t)raw[t];
  int mm = mt * mu;
  int64_t c = p / mm;
  int ij = (int)(p - c * mm);
  if (layout >= 1) {
    int i = ij / mu, j = ij - i * mu;
    int a = min(i, j), b = max(i, j);
    ij = a * mu - (a * (a - 1)) / 2 + (b - a);
  }
  uint64_t src = layout == 2 ? (uint64_t)((int64_t)ij * ncls + cls[c]) : (uint64_t)((int64_t)ij * nc + c);
  uint64_t chunk = (uint64_t)(slot / GATHER_CHUNK);
  keys[t] = (chunk << 35) | ((uint64_t)min(round, 15) << 31) | src;
  lslot[t] = (uint16_t)(slot - (int64_t)chunk * GATHER_CHUNK);
}

__global__ void k_chunk_unpack(int64_t n, const uint64_t* __restrict__ keys, int32_t* __restrict__ src) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) src[t] = (int32_t)(keys[t] & 0x7fffffffull);
}

// roundptr[chunk*(R+1)+r] = first sorted position with (chunk, round) >= (chunk, r)
__global__ void k_round_ptr(int64_t n, int64_t nchunks, int R, const uint64_t* __restrict__ keys, int32_t* __restrict__ roundptr) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nchunks * (R + 1)) return;
  uint64_t chunk = (uint64_t)(t / (R + 1));
  uint64_t r = (uint64_t)(t - (int64_t)chunk * (R + 1));
  uint64_t target = (chunk << 35) | (r << 31);
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  roundptr[t] = (int32_t)lo;
}

__global__ void k_max_sources(int64_t nnz, const int32_t* __restrict__ sptr, int* __restrict__ out) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nnz) atomicMax(out, sptr[s + 1] - sptr[s]);
}

// last cell (completion batch) any source of the chunk comes from; buffer index = plane * nc + cell
__global__ void __launch_bounds__(256)
k_chunk_batch(int64_t nc, int R, int n_batch, const int32_t* __restrict__ roundptr, const int32_t* __restrict__ lsrc,
              int32_t* __restrict__ chunk_batch) {
  __shared__ int smax[8];
  const int64_t chunk = blockIdx.x;
  const int32_t b = roundptr[chunk * (R + 1)], e = roundptr[chunk * (R + 1) + R];
  int mx = 0;
  for (int32_t k = b + threadIdx.x; k < e; k += blockDim.x) mx = max(mx, (int)(lsrc[k] % nc));
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_down_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = max(mx, smax[w]);
    chunk_batch[chunk] = (int32_t)(((int64_t)mx * n_batch) / nc);
  }
}

// optional second job of a k_gather_chunks launch: CTAs [first_cta, gridDim.x) gather a DOF vector
struct VecGather {
  unsigned first_cta = 0xffffffffu;
  int64_t n_free = 0;
  const int32_t *vptr = nullptr, *vsrc = nullptr;
  const double* vbuf = nullptr;
  double* out = nullptr;
};

__global__ void __launch_bounds__(256)
k_gather_chunks(int64_t nnz, int R, const int32_t* __restrict__ chunk_order, const int32_t* __restrict__ roundptr,
                const int32_t* __restrict__ lsrc, const uint16_t* __restrict__ lslot, const double* __restrict__ ebuf,
                double* __restrict__ vals, double scale, int add, VecGather vg) {
  __shared__ double acc[GATHER_CHUNK];
  if (blockIdx.x >= vg.first_cta) {
    // the trailing CTAs of a fused launch gather the element vectors into the DOF vector (k_gather_dofs): on slices of a few
    // hundred CTAs a separate launch costs more than the rows themselves
    const int64_t d = (int64_t)(blockIdx.x - vg.first_cta) * blockDim.x + threadIdx.x;
    if (d < vg.n_free) {
      double a = 0.0;
      for (int32_t k = vg.vptr[d]; k < vg.vptr[d + 1]; ++k) a += vg.vbuf[vg.vsrc[k]];
      a *= scale;
      vg.out[d] = add ? vg.out[d] + a : a;
    }
    return;
  }
  const int64_t chunk = chunk_order[blockIdx.x];
  const int32_t* rp = roundptr + chunk * (R + 1);
  for (int r = 0; r < R; ++r) {
    const int32_t b = rp[r], e = rp[r + 1];
    if (r > 0 && b == e) break;  // uniform: rounds are nested (a slot has round r only if it has round r-1)
    // 4 independent index -> value load chains per thread in flight
    for (int32_t k0 = b + threadIdx.x; k0 < e; k0 += 4 * blockDim.x) {
      int32_t src[4];
      uint16_t ls[4];
      double v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int32_t k = k0 + j * blockDim.x;
        src[j] = k < e ? lsrc[k] : -1;
        ls[j] = k < e ? lslot[k] : 0;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = src[j] >= 0 ? ebuf[src[j]] : 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (src[j] >= 0) acc[ls[j]] = r == 0 ? v[j] : acc[ls[j]] + v[j];
    }
    __syncthreads();
  }
  const int64_t s0 = chunk * GATHER_CHUNK;
  for (int t = threadIdx.x; t < GATHER_CHUNK; t += blockDim.x) {
    int64_t s = s0 + t;
    if (s < nnz) {
      double v = scale * acc[t];
      vals[s] = add ? vals[s] + v : v;
    }
  }
}

static void build_chunk_lists(gg_csr* A, int layout) {
  if (A->cached_layout == layout) return;
  gg_ctx* ctx = A->ctx;
  cudaStream_t st = ctx->stream;
  int64_t n = A->n_pairs, nnz = A->nnz;
  int64_t nchunks = (nnz + GATHER_CHUNK - 1) / GATHER_CHUNK;
  GG_REQUIRE(nchunks < (1ll << 28), GG_ERR_UNSUPPORTED, "matrix too large for the chunked gather");
  // maximum number of contributions to one slot
  DevBuf<int> dmax(1);
  GG_CUDA(cudaMemsetAsync(dmax.p, 0, sizeof(int), st));
  k_max_sources<<<nblk(nnz, 256), 256, 0, st>>>(nnz, A->d_sptr.p, dmax.p);
  int R = 0;
  GG_CUDA(cudaMemcpyAsync(&R, dmax.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  GG_CUDA(cudaStreamSynchronize(st));
  GG_REQUIRE(R >= 1 && R <= 15, GG_ERR_UNSUPPORTED, "a matrix entry has more than 15 cell contributions");
  DevBuf<uint64_t> k0(n), k1(n);
  DevBuf<uint16_t> v0(n), v1(n);
  gg_mesh* mesh = A->test->mesh;
  if (layout == 2) mesh_geometry_classes(mesh);
  k_chunk_keys<<<nblk(n, 256), 256, 0, st>>>(n, nnz, mesh->n_cells, A->test->ndofs_loc, A->trial->ndofs_loc, layout,
                                             A->d_sptr.p, A->d_ssrc.p, k0.p, v0.p, layout == 2 ? mesh->n_classes : 0,
                                             layout == 2 ? mesh->d_cls.p : nullptr);
  cub::DoubleBuffer<uint64_t> dk(k0.p, k1.p);
  cub::DoubleBuffer<uint16_t> dv(v0.p, v1.p);
  size_t tmp_bytes = 0;
  GG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, 64, st));
  {
    DevBuf<uint8_t> tmp(tmp_bytes);
    GG_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, (int)n, 0, 64, st));
    GG_CUDA(cudaStreamSynchronize(st));
  }
  A->d_lsrc.alloc(n);
  A->d_lslot.alloc(n);
  k_chunk_unpack<<<nblk(n, 256), 256, 0, st>>>(n, dk.Current(), A->d_lsrc.p);
  GG_CUDA(cudaMemcpyAsync(A->d_lslot.p, dv.Current(), n * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st));
  A->d_roundptr.alloc(nchunks * (R + 1));
  k_round_ptr<<<nblk(nchunks * (R + 1), 256), 256, 0, st>>>(n, nchunks, R, dk.Current(), A->d_roundptr.p);
  ctx->launches += 12;
  GG_CUDA(cudaGetLastError());
  GG_CUDA(cudaStreamSynchronize(st));
  A->n_rounds = R;
  A->n_chunks = nchunks;
  // order the chunks by the cell batch that completes them
  {
    const int NBt = pipeline_batches();
    DevBuf<int32_t> d_cb(nchunks);
    k_chunk_batch<<<(unsigned)nchunks, 256, 0, st>>>(A->test->mesh->n_cells, R, NBt, A->d_roundptr.p, A->d_lsrc.p, d_cb.p);
    ctx->launches++;
    std::vector<int32_t> cb(nchunks), order(nchunks);
    d_cb.download(cb.data(), st);
    if (layout == 2) std::fill(cb.begin(), cb.end(), 0);   // class-indexed sources: no cell batches
    std::vector<int64_t> cnt(NBt + 1, 0);
    for (int64_t i = 0; i < nchunks; ++i) cnt[cb[i] + 1]++;
    for (int b2 = 0; b2 < NBt; ++b2) cnt[b2 + 1] += cnt[b2];
    for (int b2 = 0; b2 <= NBt; ++b2) A->batch_chunk_ptr[b2] = cnt[b2];
    std::vector<int64_t> pos(cnt.begin(), cnt.end() - 1);
    for (int64_t i = 0; i < nchunks; ++i) order[pos[cb[i]]++] = (int32_t)i;
    A->d_chunk_order.upload(order, st);
    GG_CUDA(cudaStreamSynchronize(st));
  }
  A->cached_layout = layout;
}

void prepare_gather(gg_csr* A, int layout) {
  if (A->nnz) build_chunk_lists(A, layout);
}

void gather_matrix(gg_csr* A, int layout, const double* ebuf, double scale, int add) {
  gg_ctx* ctx = A->ctx;
  if (A->nnz == 0) return;
  build_chunk_lists(A, layout);
  ProfScope ps(ctx, "k_gather_chunks");
  k_gather_chunks<<<(unsigned)A->n_chunks, 256, 0, ctx->stream>>>(A->nnz, A->n_rounds, A->d_chunk_order.p, A->d_roundptr.p,
                                                                  A->d_lsrc.p, A->d_lslot.p, ebuf, A->d_vals.p, scale, add,
                                                                  VecGather());
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

// CSR fill and DOF-vector gather of the test space in ONE launch (same scale / add for both)
void gather_matrix_and_vector(gg_csr* A, int layout, const double* ebuf, const double* vbuf, double* vec_out, double scale, int add) {
  gg_ctx* ctx = A->ctx;
  gg_space* s = A->test;
  if (A->nnz == 0 || s->n_free == 0) {
    gather_matrix(A, layout, ebuf, scale, add);
    gather_vector(s, vbuf, vec_out, scale, add);
    return;
  }
  build_chunk_lists(A, layout);
  build_vec_map(s);
  VecGather vg;
  vg.first_cta = (unsigned)A->n_chunks; vg.n_free = s->n_free; vg.vptr = s->d_vptr.p; vg.vsrc = s->d_vsrc.p; vg.vbuf = vbuf; vg.out = vec_out;
  ProfScope ps(ctx, "k_gather_chunks");
  k_gather_chunks<<<(unsigned)A->n_chunks + nblk(s->n_free, 256), 256, 0, ctx->stream>>>(
      A->nnz, A->n_rounds, A->d_chunk_order.p, A->d_roundptr.p, A->d_lsrc.p, A->d_lslot.p, ebuf, A->d_vals.p, scale, add, vg);
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

void gather_matrix_batch(gg_csr* A, int batch, const double* ebuf, double scale, int add, cudaStream_t st) {
  gg_ctx* ctx = A->ctx;
  int64_t c0 = A->batch_chunk_ptr[batch], c1 = A->batch_chunk_ptr[batch + 1];
  if (c1 == c0) return;
  k_gather_chunks<<<(unsigned)(c1 - c0), 256, 0, st>>>(A->nnz, A->n_rounds, A->d_chunk_order.p + c0, A->d_roundptr.p,
                                                       A->d_lsrc.p, A->d_lslot.p, ebuf, A->d_vals.p, scale, add, VecGather());
  ctx->launches++;
  GG_CUDA(cudaGetLastError());
}

}  // namespace gg

using namespace gg;

namespace gg {
// values in CSC order: out[k] = vals[perm[k]]
__global__ void k_permute_values(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ vals, double* __restrict__ out) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = vals[perm[k]];
}
}  // namespace gg

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

int gg_csr_export_csc(gg_csr* A, int64_t* colptr, int64_t* rowval, double* nzval, int index_base) {
  GG_API_BEGIN
  GG_REQUIRE(A, GG_ERR_INVALID, "null matrix");
  GG_REQUIRE(index_base == 0 || index_base == 1, GG_ERR_INVALID, "index_base must be 0 or 1");
  GG_CUDA(cudaSetDevice(A->ctx->device));
  cudaStream_t st = A->ctx->stream;
  if (A->h_colptr.empty()) {
    // one-off: stable counting sort of the CSR slots by column => rows ascending inside every column, as in
    // SparseMatrixCSC; the permutation stays on the device for the numeric exports
    std::vector<int32_t> rp(A->n_rows + 1), ci(A->nnz), perm(A->nnz);
    A->d_rowptr.download(rp.data(), st);
    if (A->nnz) A->d_colind.download(ci.data(), st);
    A->h_colptr.assign(A->n_cols + 1, 0);
    for (int64_t k = 0; k < A->nnz; ++k) A->h_colptr[ci[k] + 1]++;
    for (int64_t j = 0; j < A->n_cols; ++j) A->h_colptr[j + 1] += A->h_colptr[j];
    std::vector<int64_t> next(A->h_colptr.begin(), A->h_colptr.end() - 1);
    A->h_rowval.resize(A->nnz);
    for (int64_t i = 0; i < A->n_rows; ++i)
      for (int32_t k = rp[i]; k < rp[i + 1]; ++k) {
        const int64_t pos = next[ci[k]]++;
        A->h_rowval[pos] = (int32_t)i;
        perm[pos] = k;
      }
    A->d_csc_perm.upload(perm, st);
    GG_CUDA(cudaStreamSynchronize(st));
  }
  if (colptr) for (int64_t j = 0; j <= A->n_cols; ++j) colptr[j] = A->h_colptr[j] + index_base;
  if (rowval) for (int64_t k = 0; k < A->nnz; ++k) rowval[k] = (int64_t)A->h_rowval[k] + index_base;
  if (nzval && A->nnz) {
    if ((int64_t)A->d_csc_vals.n != A->nnz) A->d_csc_vals.alloc(A->nnz);
    k_permute_values<<<(unsigned)((A->nnz + 255) / 256), 256, 0, st>>>(A->nnz, A->d_csc_perm.p, A->d_vals.p, A->d_csc_vals.p);
    A->ctx->launches++;
    GG_CUDA(cudaGetLastError());
    A->d_csc_vals.download(nzval, st);
  }
  GG_API_END
}

double* gg_csr_device_values(gg_csr* A) { return A ? A->d_vals.p : nullptr; }

void gg_csr_destroy(gg_csr* A) { delete A; }

}  // extern "C"
