// Partitioned models: partition plan (host), rank-local sub-model, peer windows (CUDA IPC) and the owner -> ghost update.
//
// Replaces, on the cubed-sphere atlas, the distributed model of the reference
//   src/Distributed/AtlasDistributedDiscreteModels.jl:20-73   linear cell partition `div((i-1)*P, Nc)+1` (:31), ghost layer
//                                                             = cells sharing a vertex with an owned cell (:32-49),
//                                                             (the reference numbers a part's cells owned-first, then
//                                                             ghosts, :51-66; here the local cells are sorted by global
//                                                             id -- a local permutation only, gids are ABI outputs)
// and PartitionedArrays' PRange / `consistent!` on the DOF vectors (call sites src/ODEs/StageOperators.jl:109,134,
// src/ODEs/DAE.jl:291,314).  GridapDistributed's DOF ownership rule is not in the reference tree; here a DOF belongs to
// the rank that owns the lowest-numbered cell touching it (any consistent rule gives the same numbers; gid parity with
// GridapDistributed is an ABI input question).
//
// Design: one process per GPU.  The plan is a pure function of (n, nparts): every rank computes it redundantly on the host,
// including the lists of its neighbours, so set-up needs no communication at all except the exchange of the 64-byte IPC
// handles of the peer windows (done by the caller, e.g. with torch.distributed).  Rank-local spaces use the global builtin
// numbering restricted to the local cells (boundary DOFs first, cell-private interior DOFs after them in local cell order),
// which keeps the statically condensed solver applicable.  All run-time exchanges go through peer memory (gg_dist.cuh).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include "gg_common.cuh"
#include "gg_dist.cuh"
#include "gg_dist_host.cuh"

namespace cg = cooperative_groups;

namespace gg {

static inline unsigned nblk(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// vertex-adjacent cells of the cells owned by rank r that r does not own
static std::vector<int64_t> ghost_cells_of(const gg_mesh* m, const std::vector<int32_t>& cell_owner, int r) {
  const int nv = m->dim == 3 ? 8 : 4;
  const std::vector<int32_t>& cn = m->dim == 3 ? m->cell_nodes8 : m->cell_nodes;
  std::vector<uint8_t> touched(m->n_nodes, 0);
  for (int64_t c = 0; c < m->n_cells; ++c)
    if (cell_owner[c] == r)
      for (int v = 0; v < nv; ++v) touched[cn[c * nv + v]] = 1;
  std::vector<int64_t> out;
  for (int64_t c = 0; c < m->n_cells; ++c) {
    if (cell_owner[c] == r) continue;
    bool near = false;
    for (int v = 0; v < nv; ++v) near = near || touched[cn[c * nv + v]];
    if (near) out.push_back(c);
  }
  return out;
}

static void plan_partition(gg_plan* P) {
  const int64_t Nc = P->gmesh->n_cells;
  const int nparts = P->nparts;
  GG_REQUIRE(Nc >= nparts, GG_ERR_INVALID, "fewer cells than parts");
  P->first.assign(nparts, Nc); P->last.assign(nparts, 0); P->n_owned.assign(nparts, 0); P->cells.resize(nparts);
  P->cell_owner.resize(Nc);
  if (P->mode == 1) {
    // panel strips: the in-panel index (the cell order is panel-major, AtlasGrids.jl:277-280) is cut like the linear partition
    GG_REQUIRE(P->gmesh->dim == 2 && Nc % 6 == 0 && Nc / 6 >= nparts, GG_ERR_INVALID, "panel strips need the 2-D cubed sphere");
    const int64_t per_panel = Nc / 6;
    for (int64_t c = 0; c < Nc; ++c) P->cell_owner[c] = (int32_t)(((c % per_panel) * nparts) / per_panel);
  } else {
    // part(i) = floor(i * P / Nc) (0-based i): the same cut as gg_partition_range
    for (int64_t c = 0; c < Nc; ++c) P->cell_owner[c] = (int32_t)((c * nparts) / Nc);
  }
  for (int64_t c = 0; c < Nc; ++c) {
    const int r = P->cell_owner[c];
    P->n_owned[r]++;
    P->first[r] = std::min(P->first[r], c);
    P->last[r] = std::max(P->last[r], c + 1);
  }
  for (int r = 0; r < nparts; ++r) {
    std::vector<int64_t> ghosts = ghost_cells_of(P->gmesh.get(), P->cell_owner, r);
    auto& L = P->cells[r];
    L.reserve(P->n_owned[r] + ghosts.size());
    for (int64_t c = 0; c < Nc; ++c)
      if (P->cell_owner[c] == r) L.push_back(c);
    L.insert(L.end(), ghosts.begin(), ghosts.end());
    std::sort(L.begin(), L.end());
  }
}

SpacePlan& plan_space(gg_plan* P, int kind, int order) {
  const int key = kind * 16 + order;
  auto it = P->spaces.find(key);
  if (it != P->spaces.end()) return *it->second;
  GG_REQUIRE(kind == GG_SPACE_CG || kind == GG_SPACE_DG || kind == GG_SPACE_RT || kind == GG_SPACE_CGV, GG_ERR_INVALID, "unknown space kind");
  std::unique_ptr<SpacePlan> sp(new SpacePlan);
  gg_mesh* gm = P->gmesh.get();
  gg_space G;
  space_builtin_host(gm, kind, order, GG_CONSTRAINT_NONE, &G);
  const int m = G.ndofs_loc;
  const int64_t NG = G.n_free, Nc = gm->n_cells;
  const int np = P->nparts, me = P->part;
  sp->kind = kind; sp->order = order; sp->mloc = m; sp->n_global = NG;
  // owner of a DOF = owner of the lowest-numbered cell that touches it
  std::vector<int32_t> owner_g(NG, -1);
  for (int64_t c = 0; c < Nc; ++c) {
    const int32_t r = P->cell_owner[c];
    for (int j = 0; j < m; ++j) {
      int32_t g = G.cell_dofs[c * m + j];
      if (owner_g[g] < 0) owner_g[g] = r;
    }
  }
  // local numbering of every rank: ascending global id over the DOFs of its local cells
  std::vector<std::vector<int32_t>> l2g(np);
  std::vector<int32_t> stamp(NG, -1);
  for (int r = 0; r < np; ++r) {
    auto& L = l2g[r];
    for (int64_t c : P->cells[r])
      for (int j = 0; j < m; ++j) {
        int32_t g = G.cell_dofs[c * m + j];
        if (stamp[g] != r) { stamp[g] = r; L.push_back(g); }
      }
    std::sort(L.begin(), L.end());
    sp->slot_capacity = std::max<int64_t>(sp->slot_capacity, (int64_t)L.size());
  }
  const auto& mine = l2g[me];
  std::vector<int32_t> g2l(NG, -1);
  for (size_t i = 0; i < mine.size(); ++i) g2l[mine[i]] = (int32_t)i;
  sp->l2g.assign(mine.begin(), mine.end());
  sp->owner.resize(mine.size());
  for (size_t i = 0; i < mine.size(); ++i) {
    sp->owner[i] = owner_g[mine[i]];
    if (sp->owner[i] == me) sp->n_owned++;
  }
  const auto& lc = P->cells[me];
  sp->cell_dofs.resize(lc.size() * m);
  if (!G.cell_signs.empty()) sp->signs.resize(lc.size() * m);
  for (size_t i = 0; i < lc.size(); ++i)
    for (int j = 0; j < m; ++j) {
      sp->cell_dofs[i * m + j] = g2l[G.cell_dofs[lc[i] * m + j]];
      if (!G.cell_signs.empty()) sp->signs[i * m + j] = G.cell_signs[lc[i] * m + j];
    }
  // vector H1: the change of basis of a local cell needs the chart of the entity's master cell -- the surrounding cell with the
  // largest GLOBAL id (src/FESpaces/GradConformingFESpaces.jl:14-19), which may lie beyond the ghost layer.  The reference finds
  // local masters and repairs the ghost cells' matrices with a `consistent!` of the flattened per-cell matrices
  // (src/Distributed/GradConformingFESpaces.jl:7-35); here every rank holds the global mesh on the host (the plan is a pure
  // function of (n, nparts)), so the exact master geometry of every local (cell, node) is copied from the global space: same
  // blocks as the serial space, no exchange
  if (kind == GG_SPACE_CGV) {
    const int nn = m / 2;
    sp->cob_master_panel.resize(lc.size() * nn);
    sp->cob_master_point.resize(lc.size() * nn * 2);
    for (size_t i = 0; i < lc.size(); ++i)
      for (int a = 0; a < nn; ++a) {
        sp->cob_master_panel[i * nn + a] = G.cob_master_panel[lc[i] * nn + a];
        sp->cob_master_point[(i * nn + a) * 2] = G.cob_master_point[(lc[i] * nn + a) * 2];
        sp->cob_master_point[(i * nn + a) * 2 + 1] = G.cob_master_point[(lc[i] * nn + a) * 2 + 1];
      }
  }
  // entity block classes of the local numbering (the DOFs of one entity stay consecutive under the restriction)
  for (const auto& bc : G.blocks) {
    const int64_t lo = bc.start, hi = bc.start + bc.count * bc.size;
    const int64_t a = std::lower_bound(mine.begin(), mine.end(), (int32_t)std::min<int64_t>(lo, INT32_MAX)) - mine.begin();
    const int64_t b = std::lower_bound(mine.begin(), mine.end(), (int32_t)std::min<int64_t>(hi, INT32_MAX)) - mine.begin();
    GG_REQUIRE((b - a) % bc.size == 0, GG_ERR_INVALID, "partition cuts a DOF block");
    sp->blocks.push_back({a, (b - a) / bc.size, bc.size});
  }
  // exchange lists, ascending global id on both sides
  sp->send_ptr.push_back(0);
  sp->recv_ptr.push_back(0);
  for (int r = 0; r < np; ++r) {
    if (r == me) continue;
    const size_t s0 = sp->send_idx.size(), r0 = sp->recv_idx.size();
    const auto& theirs = l2g[r];
    for (size_t k = 0; k < theirs.size(); ++k)
      if (owner_g[theirs[k]] == me) { sp->send_idx.push_back(g2l[theirs[k]]); sp->send_remote.push_back((int32_t)k); }
    for (size_t i = 0; i < mine.size(); ++i)
      if (sp->owner[i] == r) sp->recv_idx.push_back((int32_t)i);
    if (sp->send_idx.size() == s0 && sp->recv_idx.size() == r0) continue;
    sp->peers.push_back(r);
    sp->send_ptr.push_back((int32_t)sp->send_idx.size());
    sp->recv_ptr.push_back((int32_t)sp->recv_idx.size());
  }
  for (int32_t v : sp->send_idx) GG_REQUIRE(v >= 0, GG_ERR_INVALID, "an owned DOF is missing from the local numbering");
  SpacePlan& ref = *sp;
  P->spaces[key] = std::move(sp);
  return ref;
}

// rank-local space of a partitioned sub-mesh: host arrays from the plan
void space_from_plan(gg_mesh* m, int kind, int order, gg_space* s) {
  GG_REQUIRE(m->plan, GG_ERR_INVALID, "mesh has no partition plan");
  SpacePlan& sp = plan_space(m->plan, kind, order);
  s->mesh = m; s->kind = kind; s->order = order; s->constraint = GG_CONSTRAINT_NONE;
  s->ndofs_loc = sp.mloc;
  s->n_free = (int64_t)sp.l2g.size();
  s->cell_dofs = sp.cell_dofs;
  s->cell_signs = sp.signs;
  s->cob_master_panel = sp.cob_master_panel;
  s->cob_master_point = sp.cob_master_point;
  s->blocks.clear();
  for (const auto& b : sp.blocks) s->blocks.push_back({b.start, b.count, b.size});
  s->dist.reset(new gg_space::Dist);
  auto& d = *s->dist;
  d.n_global = sp.n_global; d.n_owned = sp.n_owned; d.slot_capacity = sp.slot_capacity;
  d.l2g = sp.l2g; d.owner = sp.owner; d.peers = sp.peers;
  d.send_ptr = sp.send_ptr; d.send_idx = sp.send_idx; d.send_remote = sp.send_remote;
  d.recv_ptr = sp.recv_ptr; d.recv_idx = sp.recv_idx;
}

void space_upload_dist(gg_space* s) {
  auto& d = *s->dist;
  cudaStream_t st = s->mesh->ctx->stream;
  const int me = s->mesh->plan->part;
  // device copies of the send entries sorted by local row, so that the solver can push a row's value to its ghosts the
  // moment it is computed: owned[i] = 2 marks rows with entries, push_first[i] is the first of them
  const size_t ns = d.send_idx.size();
  std::vector<int32_t> peer_of(ns);
  for (size_t k = 0; k + 1 < d.send_ptr.size(); ++k)
    for (int32_t e = d.send_ptr[k]; e < d.send_ptr[k + 1]; ++e) peer_of[e] = d.peers[k];
  std::vector<int32_t> order(ns);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return d.send_idx[a] < d.send_idx[b]; });
  std::vector<int32_t> e_row(ns), e_peer(ns), e_remote(ns);
  for (size_t k = 0; k < ns; ++k) { e_row[k] = d.send_idx[order[k]]; e_peer[k] = peer_of[order[k]]; e_remote[k] = d.send_remote[order[k]]; }
  std::vector<uint8_t> owned(d.owner.size());
  std::vector<int32_t> push_first(d.owner.size(), -1);
  for (size_t i = 0; i < owned.size(); ++i) owned[i] = d.owner[i] == me;
  for (size_t k = ns; k-- > 0;) { push_first[e_row[k]] = (int32_t)k; owned[e_row[k]] = 2; }
  d.d_owned.upload(owned, st);
  d.d_peers.upload(d.peers, st);
  d.d_send_ptr.upload(d.send_ptr, st); d.d_send_idx.upload(e_row, st); d.d_send_remote.upload(e_remote, st);
  d.d_send_peer.upload(e_peer, st); d.d_push_first.upload(push_first, st);
  d.d_recv_ptr.upload(d.recv_ptr, st); d.d_recv_idx.upload(d.recv_idx, st);
  GG_CUDA(cudaStreamSynchronize(st));
}

// device view of the exchange data of a partitioned space
DistDev dist_dev(gg_space* s) {
  DistDev d;
  if (!s->dist || !s->mesh->comm) return d;
  gg_comm* c = s->mesh->comm;
  GG_REQUIRE(c->connected, GG_ERR_INVALID, "peer windows are not connected (gg_comm_connect)");
  GG_REQUIRE(s->dist->slot_capacity <= c->slot_doubles, GG_ERR_INVALID,
             "peer window too small: needs " + std::to_string(s->dist->slot_capacity) + " doubles per slot");
  auto& x = *s->dist;
  d.rank = c->rank; d.world = c->world;
  d.timing = getenv("GG_COMM_STATS") ? 1 : 0;
  d.win = (WinHeader* const*)c->d_win.p; d.slot = (double* const*)c->d_slot.p;
  d.slot_stride = std::max<int64_t>(c->slot_doubles, 1);
  d.owned = x.d_owned.p;
  d.send_idx = x.d_send_idx.p; d.send_remote = x.d_send_remote.p; d.send_peer = x.d_send_peer.p;
  d.push_first = x.d_push_first.p;
  d.n_send = (int64_t)x.send_idx.size();
  d.recv_idx = x.d_recv_idx.p; d.n_recv = (int64_t)x.recv_idx.size();
  return d;
}

// owner -> ghost update of one vector: push, barrier round, unpack
__global__ void __launch_bounds__(256) k_consistent(DistDev d, double* vec, int64_t n) {
  cg::grid_group grid = cg::this_grid();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  unsigned long long epoch = d.win[d.rank]->epoch;
  dist_push(d, epoch, vec, tid, nth, n);
  grid.sync();
  double dummy[1] = {0.0};
  dist_allreduce<1>(grid, d, epoch, dummy);
  dist_unpack(d, epoch, vec, tid, nth, n);
  grid.sync();
  if (tid == 0) d.win[d.rank]->epoch = epoch;
}

void comm_check(gg_comm* c) {
  if (!c || c->world == 1) return;
  unsigned long long err = 0;
  GG_CUDA(cudaMemcpy(&err, (char*)c->window + offsetof(WinHeader, error), sizeof(err), cudaMemcpyDeviceToHost));
  if (err) throw Error(GG_ERR_COMM, "a wait on a peer window timed out in round " + std::to_string(err) +
                                        " (a rank is lost or out of step); results since then are invalid");
}

void vec_consistent_async(gg_space* s, double* vec) {
  if (!s->dist || !s->mesh->comm || s->mesh->comm->world == 1) return;
  gg_ctx* ctx = s->mesh->ctx;
  DistDev d = dist_dev(s);
  int64_t n = s->n_free;
  int64_t work = std::max<int64_t>(std::max(d.n_send, d.n_recv), 1);
  int grid = (int)std::min<int64_t>((work + 255) / 256, ctx->sm_count);
  void* args[] = {&d, &vec, &n};
  GG_CUDA(cudaLaunchCooperativeKernel((void*)k_consistent, dim3(grid), dim3(256), args, 0, ctx->stream));
  ctx->launches++;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_plan_create(int n, double radius, int nparts, int part, gg_plan** out) {
  GG_API_BEGIN
  GG_REQUIRE(out, GG_ERR_INVALID, "null output");
  GG_REQUIRE(n >= 1 && radius > 0, GG_ERR_INVALID, "bad mesh parameters");
  GG_REQUIRE(nparts >= 1 && nparts <= GG_MAX_RANKS && part >= 0 && part < nparts, GG_ERR_INVALID, "bad partition request (at most 16 parts)");
  std::unique_ptr<gg_plan> P(new gg_plan);
  P->n = n; P->radius = radius; P->nparts = nparts; P->part = part;
  P->gmesh.reset(build_cubed_sphere(nullptr, n, radius));
  plan_partition(P.get());
  *out = P.release();
  GG_API_END
}

int gg_plan_create_strips(int n, double radius, int nparts, int part, gg_plan** out) {
  GG_API_BEGIN
  GG_REQUIRE(out, GG_ERR_INVALID, "null output");
  GG_REQUIRE(n >= 1 && radius > 0, GG_ERR_INVALID, "bad mesh parameters");
  GG_REQUIRE(nparts >= 1 && nparts <= GG_MAX_RANKS && part >= 0 && part < nparts, GG_ERR_INVALID, "bad partition request (at most 16 parts)");
  GG_REQUIRE((int64_t)n * n >= nparts, GG_ERR_INVALID, "fewer cells per panel than parts");
  std::unique_ptr<gg_plan> P(new gg_plan);
  P->n = n; P->radius = radius; P->nparts = nparts; P->part = part; P->mode = 1;
  P->gmesh.reset(build_cubed_sphere(nullptr, n, radius));
  plan_partition(P.get());
  *out = P.release();
  GG_API_END
}

int gg_plan_owned_cells(gg_plan* p, int64_t* n_owned) {
  GG_API_BEGIN
  GG_REQUIRE(p && n_owned, GG_ERR_INVALID, "null argument");
  *n_owned = p->n_owned[p->part];
  GG_API_END
}

int gg_plan_create_shell(int n, int n_layers, double radius, double thickness, int ordering, int nparts, int part, gg_plan** out) {
  GG_API_BEGIN
  GG_REQUIRE(out, GG_ERR_INVALID, "null output");
  GG_REQUIRE(n >= 1 && radius > 0 && thickness > 0 && n_layers >= 0, GG_ERR_INVALID, "bad mesh parameters");
  GG_REQUIRE(ordering == 0 || ordering == 2, GG_ERR_INVALID, "shell orderings: 0 (lexicographic) or 2 (p8est Morton)");
  if (ordering == 2) GG_REQUIRE((n & (n - 1)) == 0 && (n_layers == 0 || n_layers == n), GG_ERR_INVALID, "the p8est ordering needs n = n_layers = 2^l");
  GG_REQUIRE(nparts >= 1 && nparts <= GG_MAX_RANKS && part >= 0 && part < nparts, GG_ERR_INVALID, "bad partition request (at most 16 parts)");
  std::unique_ptr<gg_plan> P(new gg_plan);
  P->n = n; P->radius = radius; P->nparts = nparts; P->part = part;
  P->gmesh.reset(build_cubed_sphere_shell(nullptr, n, n_layers > 0 ? n_layers : n, radius, thickness, ordering));
  plan_partition(P.get());
  *out = P.release();
  GG_API_END
}

void gg_plan_destroy(gg_plan* p) { delete p; }

int gg_plan_info(gg_plan* p, int64_t* n_cells_global, int64_t* first_owned, int64_t* last_owned, int64_t* n_local) {
  GG_API_BEGIN
  GG_REQUIRE(p, GG_ERR_INVALID, "null plan");
  if (n_cells_global) *n_cells_global = p->gmesh->n_cells;
  if (first_owned) *first_owned = p->first[p->part];
  if (last_owned) *last_owned = p->last[p->part];
  if (n_local) *n_local = (int64_t)p->cells[p->part].size();
  GG_API_END
}

int gg_plan_local_cells(gg_plan* p, int64_t* out) {
  GG_API_BEGIN
  GG_REQUIRE(p && out, GG_ERR_INVALID, "null argument");
  const auto& L = p->cells[p->part];
  std::copy(L.begin(), L.end(), out);
  GG_API_END
}

int gg_plan_space_info(gg_plan* p, int kind, int order, int64_t* n_global, int64_t* n_local, int64_t* n_owned,
                       int32_t* n_neighbors, int64_t* slot_capacity) {
  GG_API_BEGIN
  GG_REQUIRE(p, GG_ERR_INVALID, "null plan");
  SpacePlan& sp = plan_space(p, kind, order);
  if (n_global) *n_global = sp.n_global;
  if (n_local) *n_local = (int64_t)sp.l2g.size();
  if (n_owned) *n_owned = sp.n_owned;
  if (n_neighbors) *n_neighbors = (int32_t)sp.peers.size();
  if (slot_capacity) *slot_capacity = sp.slot_capacity;
  GG_API_END
}

int gg_plan_space_maps(gg_plan* p, int kind, int order, int64_t* l2g, int32_t* owner) {
  GG_API_BEGIN
  GG_REQUIRE(p, GG_ERR_INVALID, "null plan");
  SpacePlan& sp = plan_space(p, kind, order);
  if (l2g) std::copy(sp.l2g.begin(), sp.l2g.end(), l2g);
  if (owner) std::copy(sp.owner.begin(), sp.owner.end(), owner);
  GG_API_END
}

int gg_plan_space_neighbor(gg_plan* p, int kind, int order, int k, int32_t* peer, int64_t* n_send, int64_t* n_recv,
                           int32_t* send_idx, int32_t* send_remote_idx, int32_t* recv_idx) {
  GG_API_BEGIN
  GG_REQUIRE(p, GG_ERR_INVALID, "null plan");
  SpacePlan& sp = plan_space(p, kind, order);
  GG_REQUIRE(k >= 0 && k < (int)sp.peers.size(), GG_ERR_INVALID, "neighbour index out of range");
  const int32_t s0 = sp.send_ptr[k], s1 = sp.send_ptr[k + 1], r0 = sp.recv_ptr[k], r1 = sp.recv_ptr[k + 1];
  if (peer) *peer = sp.peers[k];
  if (n_send) *n_send = s1 - s0;
  if (n_recv) *n_recv = r1 - r0;
  if (send_idx) std::copy(sp.send_idx.begin() + s0, sp.send_idx.begin() + s1, send_idx);
  if (send_remote_idx) std::copy(sp.send_remote.begin() + s0, sp.send_remote.begin() + s1, send_remote_idx);
  if (recv_idx) std::copy(sp.recv_idx.begin() + r0, sp.recv_idx.begin() + r1, recv_idx);
  GG_API_END
}

// ---- peer windows ------------------------------------------------------------------------------------------------------

int gg_comm_create(gg_ctx* ctx, int rank, int world, int64_t slot_doubles, gg_comm** out) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(world >= 1 && world <= GG_MAX_RANKS && rank >= 0 && rank < world && slot_doubles >= 0, GG_ERR_INVALID, "bad communicator");
  GG_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gg_comm> c(new gg_comm);
  c->ctx = ctx; c->rank = rank; c->world = world; c->slot_doubles = slot_doubles;
  c->bytes = sizeof(WinHeader) + 2 * (size_t)std::max<int64_t>(slot_doubles, 1) * sizeof(double);  // two halves (round parity)
  GG_CUDA(cudaMalloc(&c->window, c->bytes));
  GG_CUDA(cudaMemset(c->window, 0, c->bytes));
  c->peer.assign(world, nullptr);
  c->peer[rank] = c->window;
  if (world == 1) {
    std::vector<void*> w{c->window}, s{(char*)c->window + sizeof(WinHeader)};
    c->d_win.upload(w, ctx->stream); c->d_slot.upload(s, ctx->stream);
    GG_CUDA(cudaStreamSynchronize(ctx->stream));
    c->connected = true;
  }
  *out = c.release();
  GG_API_END
}

int gg_comm_handle(gg_comm* c, uint8_t* out64) {
  GG_API_BEGIN
  GG_REQUIRE(c && out64, GG_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  GG_CUDA(cudaSetDevice(c->ctx->device));
  cudaIpcMemHandle_t h;
  GG_CUDA(cudaIpcGetMemHandle(&h, c->window));
  std::memcpy(out64, &h, 64);
  GG_API_END
}

int gg_comm_connect(gg_comm* c, const uint8_t* handles) {
  GG_API_BEGIN
  GG_REQUIRE(c && handles, GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(c->ctx->device));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + (size_t)r * 64, 64);
    GG_CUDA(cudaIpcOpenMemHandle(&c->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  std::vector<void*> w(c->world), s(c->world);
  for (int r = 0; r < c->world; ++r) { w[r] = c->peer[r]; s[r] = (char*)c->peer[r] + sizeof(WinHeader); }
  c->d_win.upload(w, c->ctx->stream); c->d_slot.upload(s, c->ctx->stream);
  GG_CUDA(cudaStreamSynchronize(c->ctx->stream));
  c->connected = true;
  GG_API_END
}

int gg_comm_error(gg_comm* c, int64_t* error_epoch) {
  GG_API_BEGIN
  GG_REQUIRE(c && error_epoch, GG_ERR_INVALID, "null argument");
  GG_CUDA(cudaSetDevice(c->ctx->device));
  GG_CUDA(cudaStreamSynchronize(c->ctx->stream));
  WinHeader h;
  GG_CUDA(cudaMemcpy(&h, c->window, sizeof(h), cudaMemcpyDeviceToHost));
  *error_epoch = (int64_t)h.error;
  if (getenv("GG_COMM_STATS"))
    fprintf(stderr, "[gg_comm rank %d] rounds %llu  wait %.2f us/round  post %.2f us/round\n", c->rank, h.rounds,
            h.rounds ? 1e-3 * h.wait_ns / h.rounds : 0.0, h.rounds ? 1e-3 * h.post_ns / h.rounds : 0.0);
  GG_API_END
}

void gg_comm_destroy(gg_comm* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->window) cudaFree(c->window);
  delete c;
}

// ---- rank-local sub-model ------------------------------------------------------------------------------------------------

int gg_mesh_partitioned(gg_ctx* ctx, gg_plan* p, gg_comm* comm, gg_mesh** out) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && p && out, GG_ERR_INVALID, "null argument");
  if (comm) GG_REQUIRE(comm->rank == p->part && comm->world == p->nparts, GG_ERR_INVALID, "communicator and plan disagree");
  GG_CUDA(cudaSetDevice(ctx->device));
  const gg_mesh* g = p->gmesh.get();
  const auto& L = p->cells[p->part];
  if (g->dim == 3) {
    // rank-local sub-model of the shell: the listed hexahedra (global node ids kept; the DOF numbering of spaces comes from the plan)
    std::unique_ptr<gg_mesh> s3(new gg_mesh);
    s3->ctx = ctx; s3->dim = 3; s3->n = g->n; s3->n_layers = g->n_layers; s3->radius = g->radius; s3->thickness = g->thickness;
    s3->n_cells = (int64_t)L.size(); s3->n_nodes = g->n_nodes;
    s3->cell_nodes8.resize(L.size() * 8); s3->cell_to_chart.resize(L.size()); s3->cell_col.resize(L.size()); s3->cell_layer.resize(L.size());
    s3->cell_owned.resize(L.size());
    for (size_t i = 0; i < L.size(); ++i) {
      const int64_t c = L[i];
      std::copy(&g->cell_nodes8[c * 8], &g->cell_nodes8[c * 8] + 8, &s3->cell_nodes8[i * 8]);
      s3->cell_to_chart[i] = g->cell_to_chart[c]; s3->cell_col[i] = g->cell_col[c]; s3->cell_layer[i] = g->cell_layer[c];
      s3->cell_owned[i] = p->cell_owner[c] == p->part;
    }
    s3->surface.reset(build_cubed_sphere(ctx, g->n, g->radius, 0));
    s3->d_col.upload(s3->cell_col, ctx->stream); s3->d_layer.upload(s3->cell_layer, ctx->stream);
    s3->d_chart.upload(s3->cell_to_chart, ctx->stream); s3->d_cell_owned.upload(s3->cell_owned, ctx->stream);
    GG_CUDA(cudaStreamSynchronize(ctx->stream));
    s3->plan = p; s3->comm = comm;
    *out = s3.release();
    return GG_OK;
  }
  std::unique_ptr<gg_mesh> s(new gg_mesh);
  s->ctx = ctx; s->dim = 2; s->n = 0; s->radius = g->radius;
  s->n_cells = (int64_t)L.size(); s->n_nodes = g->n_nodes;
  s->cell_nodes.resize(L.size() * 4); s->cell_to_chart.resize(L.size()); s->chart_coords.resize(L.size() * 8);
  s->cell_owned.resize(L.size());
  for (size_t i = 0; i < L.size(); ++i) {
    const int64_t c = L[i];
    std::copy(&g->cell_nodes[c * 4], &g->cell_nodes[c * 4] + 4, &s->cell_nodes[i * 4]);
    s->cell_to_chart[i] = g->cell_to_chart[c];
    std::copy(&g->chart_coords[c * 8], &g->chart_coords[c * 8] + 8, &s->chart_coords[i * 8]);
    s->cell_owned[i] = p->cell_owner[c] == p->part;
  }
  s->plan = p; s->comm = comm;
  upload_geometry(s.get());
  *out = s.release();
  GG_API_END
}

int gg_space_global_dofs(gg_space* s, int64_t* l2g, int32_t* owner) {
  GG_API_BEGIN
  GG_REQUIRE(s, GG_ERR_INVALID, "null space");
  GG_REQUIRE(s->dist, GG_ERR_INVALID, "not a partitioned space");
  if (l2g) for (size_t i = 0; i < s->dist->l2g.size(); ++i) l2g[i] = s->dist->l2g[i] + 1;
  if (owner) std::copy(s->dist->owner.begin(), s->dist->owner.end(), owner);
  GG_API_END
}

int gg_vec_consistent(gg_space* s, gg_vec* v) {
  GG_API_BEGIN
  GG_REQUIRE(s && v, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(gg_vec_size(v) == s->n_free, GG_ERR_INVALID, "vector has the wrong length");
  GG_CUDA(cudaSetDevice(s->mesh->ctx->device));
  vec_consistent_async(s, v->d.p);
  GG_CUDA(cudaStreamSynchronize(s->mesh->ctx->stream));
  comm_check(s->mesh->comm);
  GG_API_END
}

}  // extern "C"
