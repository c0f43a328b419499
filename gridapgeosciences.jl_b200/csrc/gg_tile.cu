// Host-side construction of the tile plan (gg_tile.cuh): one-off, part of the symbolic phase of a matrix.
#include <algorithm>
#include <array>
#include <cstring>
#include <map>
#include <unordered_map>
#include "gg_common.cuh"
#include "gg_refel.cuh"
#include "gg_tile.cuh"

namespace gg {

namespace {

constexpr int TX = 16, TY = 8;

// byte-wise identity of a pattern: FNV hash + full comparison
struct Blob {
  std::vector<uint8_t> b;
  template <class T> void put(const std::vector<T>& v) {
    const uint64_t n = v.size();
    const uint8_t* q = (const uint8_t*)&n;
    b.insert(b.end(), q, q + 8);
    const uint8_t* p = (const uint8_t*)v.data();
    b.insert(b.end(), p, p + v.size() * sizeof(T));
  }
  uint64_t hash() const {
    uint64_t h = 1469598103934665603ull;
    for (uint8_t x : b) { h ^= x; h *= 1099511628211ull; }
    return h;
  }
};

struct PatStore {
  std::unordered_map<uint64_t, std::vector<std::pair<Blob, int32_t>>> map;
  int32_t find_or_add(Blob&& k, int32_t next_id, bool* is_new) {
    auto& bucket = map[k.hash()];
    for (auto& e : bucket)
      if (e.first.b == k.b) { *is_new = false; return e.second; }
    bucket.emplace_back(std::move(k), next_id);
    *is_new = true;
    return next_id;
  }
};

inline int plane_of(int M, int i, int j) {
  const int a = std::min(i, j), b = std::max(i, j);
  return a * M - (a * (a - 1)) / 2 + (b - a);
}

}  // namespace

TileDev TilePlan::dev() const {
  TileDev T;
  T.n_super = (int32_t)n_super; T.pad0 = 0;
  T.inst_ptr = inst_ptr.p; T.inst_tile = inst_tile.p;
  T.tile_cells = tile_cells.p; T.tile_pat = tile_pat.p; T.tile_eoff = tile_eoff.p; T.tile_voff = tile_voff.p;
  T.tile_row0 = tile_row0.p; T.trow_base = trow_base.p; T.trow_dof = trow_dof.p;
  T.pats = pats.p; T.pair_idx = pair_idx.p; T.vec_idx = vec_idx.p;
  T.tc_ix = tc_ix.p; T.tc_iy = tc_iy.p; T.tc_dofs = tc_dofs.p; T.tc_hx = tc_hx.p; T.tc_hy = tc_hy.p; T.bidx = bidx.p; T.types = types.p;
  T.grows = grows.p; T.n_grows = n_grows;
  T.ebuf_b = ebuf_b.p; T.vbuf_b = vbuf_b.p;
  return T;
}

std::shared_ptr<TilePlan> build_tile_plan(gg_csr* A, bool classes) {
  gg_space* S = A->test;
  gg_mesh* m = S->mesh;
  gg_ctx* ctx = A->ctx;
  cudaStream_t st = ctx->stream;
  const int M = S->ndofs_loc, NP = M * (M + 1) / 2;
  const int64_t nc = m->n_cells, nd = S->n_free;
  if (A->test != A->trial || m->dim != 2 || M > 16 || NP > 63 || nc == 0 || A->nnz == 0) return nullptr;
  if ((int64_t)m->h_ix.size() != nc || (int64_t)m->h_iy.size() != nc || (int64_t)S->cell_dofs.size() != nc * M) return nullptr;
  if (classes) {
    mesh_geometry_classes(m);
    if ((int64_t)m->cls.size() != nc) return nullptr;
  }
  std::vector<int32_t> rp(A->n_rows + 1), ci(A->nnz);
  A->d_rowptr.download(rp.data(), st);
  A->d_colind.download(ci.data(), st);
  const int32_t* dofs = S->cell_dofs.data();

  // ---- tiles: TX x TY blocks of the (x-interval, y-interval) grid of every panel; cells ascending inside a tile ----------
  std::vector<int64_t> order(nc);
  std::vector<uint64_t> key(nc);
  for (int64_t c = 0; c < nc; ++c)
    key[c] = ((uint64_t)m->cell_to_chart[c] << 48) | ((uint64_t)(m->h_iy[c] / TY) << 24) | (uint64_t)(m->h_ix[c] / TX);
  for (int64_t c = 0; c < nc; ++c) order[c] = c;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[a] < key[b]; });
  std::vector<int32_t> tile_cells, cell_tile(nc), cell_lc(nc);
  int64_t n_tiles = 0;
  for (int64_t k = 0; k < nc;) {
    int64_t e = k;
    while (e < nc && key[order[e]] == key[order[k]] && e - k < TILE_CELLS) ++e;   // an over-full block is split
    for (int64_t q = k; q < e; ++q) {
      cell_tile[order[q]] = (int32_t)n_tiles;
      cell_lc[order[q]] = (int32_t)(q - k);
      tile_cells.push_back((int32_t)order[q]);
    }
    tile_cells.resize((n_tiles + 1) * TILE_CELLS, -1);
    ++n_tiles;
    k = e;
  }

  // ---- DOF -> (cell, local dof) lists in ascending cell order, DOF -> set of tiles around it -----------------------------
  std::vector<int32_t> dptr(nd + 1, 0);
  for (int64_t k = 0; k < nc * M; ++k)
    if (dofs[k] >= 0) dptr[dofs[k] + 1]++;
  for (int64_t d = 0; d < nd; ++d) dptr[d + 1] += dptr[d];
  std::vector<int32_t> dsrc(dptr[nd]);
  {
    std::vector<int32_t> pos(dptr.begin(), dptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
      for (int i = 0; i < M; ++i) {
        const int32_t d = dofs[c * M + i];
        if (d >= 0) dsrc[pos[d]++] = (int32_t)(c * M + i);
      }
  }
  std::vector<int32_t> dtiles(nd * 4, -1);
  std::vector<uint8_t> dnt(nd, 0);
  for (int64_t d = 0; d < nd; ++d) {
    if (dptr[d + 1] - dptr[d] > 4 || rp[d + 1] - rp[d] > TILE_MAXLEN) return nullptr;   // more than four cells around a DOF / long rows
    for (int32_t k = dptr[d]; k < dptr[d + 1]; ++k) {
      const int32_t t = cell_tile[dsrc[k] / M];
      bool seen = false;
      for (int q = 0; q < dnt[d]; ++q) seen |= dtiles[d * 4 + q] == t;
      if (!seen) dtiles[d * 4 + dnt[d]++] = t;
    }
    std::sort(&dtiles[d * 4], &dtiles[d * 4] + dnt[d]);
  }

  // ---- row types ------------------------------------------------------------------------------------------------------------------
  std::vector<RowType> types;
  std::map<std::vector<uint8_t>, int32_t> type_id;
  std::vector<int32_t> dof_type(nd);
  for (int64_t d = 0; d < nd; ++d) {
    RowType R;
    std::memset(&R, 0xFF, sizeof(R));
    R.len = rp[d + 1] - rp[d];
    for (int32_t q = dptr[d]; q < dptr[d + 1]; ++q) {        // slot = position in the ascending cell list
      const int s_ = q - dptr[d];
      const int32_t c = dsrc[q] / M, i = dsrc[q] % M;
      R.vsrc[s_] = (uint8_t)((s_ << 4) | i);
      for (int j2 = 0; j2 < M; ++j2) {
        const int32_t col = dofs[(int64_t)c * M + j2];
        if (col < 0) continue;
        const int32_t* b = &ci[rp[d]];
        const int32_t* e = &ci[rp[d + 1]];
        const int32_t* it = std::lower_bound(b, e, col);
        if (it == e || *it != col) return nullptr;
        uint8_t* src = R.src[it - b];
        int w = 0;
        while (w < 4 && src[w] != 0xFF) ++w;
        if (w == 4) return nullptr;
        src[w] = (uint8_t)((s_ << 6) | plane_of(M, i, j2));
      }
    }
    std::vector<uint8_t> k((const uint8_t*)&R, (const uint8_t*)&R + sizeof(R));
    auto it = type_id.find(k);
    if (it == type_id.end()) {
      it = type_id.emplace(std::move(k), (int32_t)types.size()).first;
      types.push_back(R);
      if (types.size() > 60000) return nullptr;
    }
    dof_type[d] = it->second;
  }

  // position of column q in row d
  auto col_pos = [&](int32_t d, int32_t q) -> int {
    const int32_t* b = &ci[rp[d]];
    const int32_t* e = &ci[rp[d + 1]];
    const int32_t* it = std::lower_bound(b, e, q);
    return (it != e && *it == q) ? (int)(it - b) : -1;
  };

  auto plan = std::make_shared<TilePlan>();
  plan->classes = classes; plan->M = M; plan->NP = NP; plan->n_tiles = n_tiles;

  // ---- border cells of every tile (cells with a DOF shared with another tile) -------------------------------------------------
  std::vector<uint8_t> bidx_tile(n_tiles * TILE_CELLS, 0xFF);
  std::vector<int32_t> tile_eoff(n_tiles), tile_voff(n_tiles);
  for (int64_t t = 0; t < n_tiles; ++t)
    for (int lc = 0; lc < TILE_CELLS; ++lc) {
      const int32_t c = tile_cells[t * TILE_CELLS + lc];
      if (c < 0) continue;
      bool border = false;
      for (int i = 0; i < M; ++i) {
        const int32_t d = dofs[(int64_t)c * M + i];
        border |= d >= 0 && dnt[d] > 1;
      }
      if (border) { bidx_tile[t * TILE_CELLS + lc] = 1; plan->border_cells++; }
    }
  {
    // [tile * TILE_CELLS + local cell][plane] (rows padded to an even length); only border cells are ever written / read
    int64_t eo = 0, vo = 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
      tile_eoff[t] = (int32_t)eo; tile_voff[t] = (int32_t)vo;
      eo += (int64_t)TILE_CELLS * (NP + (NP & 1)); vo += (int64_t)TILE_CELLS * (M + (M & 1));
    }
    if (eo >= (int64_t)INT32_MAX) return nullptr;
    plan->ebuf_b.alloc((size_t)std::max<int64_t>(eo, 1));
    plan->vbuf_b.alloc((size_t)std::max<int64_t>(vo, 1));
  }

  // rows of one pattern: sorted by (length, DOF id), cut into classes of equal length
  auto sort_rows = [&](std::vector<int32_t>& rows) {
    std::sort(rows.begin(), rows.end(), [&](int32_t a, int32_t b) {
      const int la = rp[a + 1] - rp[a], lb = rp[b + 1] - rp[b];
      return la != lb ? la > lb : a < b;
    });
  };
  auto make_classes = [&](const std::vector<int32_t>& rows, TilePat& P) -> bool {
    P.nrows = (int32_t)rows.size(); P.ncls = 0; P.nslots = 0;
    std::memset(P.cls, 0, sizeof(P.cls));
    for (size_t k = 0; k < rows.size(); ++k) {
      const int len = rp[rows[k] + 1] - rp[rows[k]];
      if (P.ncls == 0 || P.cls[P.ncls - 1].len != len) {
        if (P.ncls == TILE_MAXCLS) return false;
        P.cls[P.ncls++] = RowClass{len, (int32_t)k, 0, 0};
      }
      P.cls[P.ncls - 1].count++;
    }
    return true;
  };

  // ---- complete rows of every tile -> patterns ----------------------------------------------------------------------------------
  std::vector<int32_t> l2x(M);                 // Gridap local dof -> lexicographic index (the order of the kernel's registers)
  {
    std::vector<int32_t> t = quad_local_to_lex(S->order);   // the same table the kernel's c_lex2loc is built from
    if ((int)t.size() != M) return nullptr;
    for (int i = 0; i < M; ++i) l2x[i] = t[i];
  }
  std::vector<TilePat> pats;
  std::vector<uint32_t> pair_idx, vec_idx;
  const int PW = (M * M + 1) / 2, VW = (M + 1) / 2;
  std::vector<int32_t> tile_row0(n_tiles), trow_base, trow_dof;
  std::vector<uint8_t> bidx;
  std::vector<int32_t> tile_pat(n_tiles);
  PatStore store;
  std::vector<std::vector<int32_t>> tile_rows(n_tiles);
  for (int64_t d = 0; d < nd; ++d)
    if (dnt[d] == 1) tile_rows[dtiles[d * 4]].push_back((int32_t)d);
  std::vector<int32_t> row_slot0(nd, -1), row_k(nd, -1);    // scratch: first shared-memory slot / row index of a complete row
  for (int64_t t = 0; t < n_tiles; ++t) {
    auto& rows = tile_rows[t];
    if ((int)rows.size() > TILE_ROWCAP) return nullptr;
    sort_rows(rows);
    TilePat P;
    P.row0 = 0;
    if (!make_classes(rows, P)) return nullptr;
    int nslots = 0;
    for (int q = 0; q < P.ncls; ++q) { P.cls[q].slot0 = nslots; nslots += P.cls[q].len * P.cls[q].count; }
    if (nslots + 32 > 16384) return nullptr;      // 14-bit slots
    P.nslots = nslots;
    tile_row0[t] = (int32_t)trow_base.size();
    {
      int q = 0;
      for (size_t k = 0; k < rows.size(); ++k) {
        const int32_t d = rows[k];
        while (q + 1 < P.ncls && (int)k >= P.cls[q + 1].first) ++q;
        row_slot0[d] = P.cls[q].slot0 + ((int)k - P.cls[q].first) * P.cls[q].len;
        row_k[d] = (int)k;
        trow_base.push_back(rp[d]);
        trow_dof.push_back(d);
        plan->complete_slots += rp[d + 1] - rp[d];
      }
    }
    // slot of every (local cell, I, J) in lexicographic local order, with the ROUND of the contribution in the two top bits: the
    // rank of the cell among the cells that feed the slot (ascending cell id).  Round 0 stores, rounds 1..3 add, with a barrier
    // between rounds: conflict-free without atomics, and every slot is summed in ascending cell order.  Entries of rows that are
    // not complete here (or Dirichlet) go, as round 0, to one of 32 dummy slots (one per lane) past the real ones
    std::vector<uint32_t> p_pair((size_t)PW * TILE_CELLS, 0), p_vec((size_t)VW * TILE_CELLS, 0);
    auto put16 = [](std::vector<uint32_t>& v, int idx, int lc, uint16_t val) {
      uint32_t& w = v[(size_t)(idx >> 1) * TILE_CELLS + lc];
      w = (idx & 1) ? (w & 0xFFFFu) | ((uint32_t)val << 16) : (w & 0xFFFF0000u) | val;
    };
    for (int lc = 0; lc < TILE_CELLS; ++lc) {
      const int32_t c = tile_cells[t * TILE_CELLS + lc];
      const uint16_t dummy = (uint16_t)(nslots + (lc & 31)), vdummy = (uint16_t)((int)rows.size() + (lc & 31));
      for (int i = 0; i < M; ++i) {
        const int I = l2x[i];
        const int32_t d = c >= 0 ? dofs[(int64_t)c * M + i] : -1;
        const bool complete = d >= 0 && dnt[d] == 1;
        int vround = 0;
        if (complete)
          for (int32_t q = dptr[d]; q < dptr[d + 1] && dsrc[q] / M != c; ++q) ++vround;   // cells before this one around the DOF
        if (vround > 0 && !tile_shared_dof(S->order, I)) return nullptr;
        put16(p_vec, I, lc, complete ? (uint16_t)(row_k[d] | (vround << 14)) : vdummy);
        for (int j2 = 0; j2 < M; ++j2) {
          const int J = l2x[j2];
          uint16_t slot = dummy;
          const int32_t col = c >= 0 ? dofs[(int64_t)c * M + j2] : -1;
          if (complete && col >= 0) {
            const int pos = col_pos(d, col);
            if (pos < 0) return nullptr;
            int round = 0;                     // cells before this one that hold both DOFs
            for (int32_t q = dptr[d]; q < dptr[d + 1] && dsrc[q] / M != c; ++q) {
              const int32_t c2 = dsrc[q] / M;
              bool has = false;
              for (int j3 = 0; j3 < M; ++j3) has |= dofs[(int64_t)c2 * M + j3] == col;
              round += has;
            }
            // the kernel only looks for later rounds where the local geometry allows them
            if (round > 0 && !tile_shared_pair(S->order, I, J)) return nullptr;
            if (round > 1 && !(I == J && tile_corner_dof(S->order, I))) return nullptr;
            slot = (uint16_t)((row_slot0[d] + pos) | (round << 14));
          }
          put16(p_pair, I * M + J, lc, slot);
        }
      }
    }
    std::vector<uint8_t> p_bidx(&bidx_tile[t * TILE_CELLS], &bidx_tile[t * TILE_CELLS] + TILE_CELLS);
    std::vector<RowClass> p_cls(P.cls, P.cls + TILE_MAXCLS);
    Blob key2;
    key2.put(p_pair); key2.put(p_vec); key2.put(p_bidx); key2.put(p_cls);
    bool is_new = false;
    const int32_t id = store.find_or_add(std::move(key2), (int32_t)pats.size(), &is_new);
    tile_pat[t] = id;
    if (is_new) {
      pats.push_back(P);
      pair_idx.insert(pair_idx.end(), p_pair.begin(), p_pair.end());
      vec_idx.insert(vec_idx.end(), p_vec.begin(), p_vec.end());
      bidx.insert(bidx.end(), p_bidx.begin(), p_bidx.end());
    }
    plan->max_rows = std::max(plan->max_rows, (int)rows.size());
    plan->max_slots = std::max(plan->max_slots, nslots);
  }

  // ---- incomplete rows (DOFs on tile edges and corners): one record per row, sorted by (length, DOF id) ------------------------
  std::vector<int32_t> inc;
  for (int64_t d = 0; d < nd; ++d)
    if (dnt[d] > 1) inc.push_back((int32_t)d);
  sort_rows(inc);
  std::vector<GroupRow> grows_all(inc.size());
  {
    TilePat P;
    if (!make_classes(inc, P)) return nullptr;
    GroupClasses& G = plan->gcls;
    std::memset(&G, 0, sizeof(G));
    G.ncls = P.ncls;
    int64_t e = 0;
    for (int q = 0; q < P.ncls; ++q) { G.cls[q] = P.cls[q]; G.ent0[q] = e; e += (int64_t)P.cls[q].len * P.cls[q].count; }
    for (int q = P.ncls; q <= TILE_MAXCLS; ++q) G.ent0[q] = e;
    plan->group_slots = e;
    if (e >= (int64_t)INT32_MAX) return nullptr;
  }
  for (size_t k = 0; k < inc.size(); ++k) {
    const int32_t d = inc[k];
    GroupRow R;
    std::memset(&R, 0, sizeof(R));
    R.base = rp[d]; R.dof = d; R.type = dof_type[d];
    for (int32_t q = dptr[d]; q < dptr[d + 1]; ++q) {
      const int32_t c = dsrc[q] / M;
      R.col[q - dptr[d]] = cell_tile[c] * TILE_CELLS + cell_lc[c];
    }
    grows_all[k] = R;
  }
  plan->n_grows = (int64_t)inc.size();

  // ---- supertiles: tiles that share their geometry (geometry classes) are served by one CTA --------------------------------------
  std::vector<int32_t> inst_ptr{0}, inst_tile;
  if (classes) {
    std::map<std::vector<int32_t>, std::vector<int32_t>> by_geom;
    for (int64_t t = 0; t < n_tiles; ++t) {
      std::vector<int32_t> k(TILE_CELLS);
      for (int lc = 0; lc < TILE_CELLS; ++lc) {
        const int32_t c = tile_cells[t * TILE_CELLS + lc];
        k[lc] = c < 0 ? -1 : m->cls[c];
      }
      by_geom[k].push_back((int32_t)t);
    }
    // keep the launch order close to the tile order: supertiles sorted by their first tile
    std::vector<const std::vector<int32_t>*> supers;
    for (auto& kv : by_geom) supers.push_back(&kv.second);
    std::sort(supers.begin(), supers.end(), [](const std::vector<int32_t>* a, const std::vector<int32_t>* b) { return (*a)[0] < (*b)[0]; });
    for (auto* sp : supers) {
      inst_tile.insert(inst_tile.end(), sp->begin(), sp->end());
      inst_ptr.push_back((int32_t)inst_tile.size());
    }
  } else {
    for (int64_t t = 0; t < n_tiles; ++t) { inst_tile.push_back((int32_t)t); inst_ptr.push_back((int32_t)t + 1); }
  }
  plan->n_super = (int64_t)inst_ptr.size() - 1;
  plan->n_pat = (int64_t)pats.size();
  plan->n_types = (int64_t)types.size();

  // ---- per (tile, local cell) copies of what a thread reads first ----------------------------------------------------------------------
  std::vector<int32_t> tc_ix(n_tiles * TILE_CELLS, 0), tc_iy(n_tiles * TILE_CELLS, 0), tc_dofs((size_t)n_tiles * M * TILE_CELLS, -1);
  std::vector<double> tc_hx(n_tiles * TILE_CELLS, 1.0), tc_hy(n_tiles * TILE_CELLS, 1.0);
  for (int64_t t = 0; t < n_tiles; ++t)
    for (int lc = 0; lc < TILE_CELLS; ++lc) {
      const int32_t c = tile_cells[t * TILE_CELLS + lc];
      if (c < 0) continue;
      const double* q = &m->chart_coords[(size_t)c * 8];
      tc_ix[t * TILE_CELLS + lc] = m->h_ix[c]; tc_iy[t * TILE_CELLS + lc] = m->h_iy[c];
      tc_hx[t * TILE_CELLS + lc] = q[2] - q[0]; tc_hy[t * TILE_CELLS + lc] = q[5] - q[1];
      for (int i = 0; i < M; ++i) tc_dofs[((size_t)t * M + l2x[i]) * TILE_CELLS + lc] = dofs[(int64_t)c * M + i];
    }
  plan->tc_ix.upload(tc_ix, st); plan->tc_iy.upload(tc_iy, st); plan->tc_dofs.upload(tc_dofs, st);
  plan->tc_hx.upload(tc_hx, st); plan->tc_hy.upload(tc_hy, st);

  // ---- upload ------------------------------------------------------------------------------------------------------------------------
  plan->inst_ptr.upload(inst_ptr, st); plan->inst_tile.upload(inst_tile, st);
  plan->tile_cells.upload(tile_cells, st); plan->tile_pat.upload(tile_pat, st);
  plan->tile_eoff.upload(tile_eoff, st); plan->tile_voff.upload(tile_voff, st);
  plan->tile_row0.upload(tile_row0, st); plan->trow_base.upload(trow_base, st); plan->trow_dof.upload(trow_dof, st);
  plan->pats.upload(pats, st); plan->pair_idx.upload(pair_idx, st); plan->vec_idx.upload(vec_idx, st); plan->bidx.upload(bidx, st);
  plan->types.upload(types, st);
  plan->grows.upload(grows_all, st);
  GG_CUDA(cudaStreamSynchronize(st));
  plan->pattern_bytes = (int64_t)((pair_idx.size() + vec_idx.size()) * sizeof(uint32_t) + grows_all.size() * sizeof(GroupRow) + bidx.size() +
                                  types.size() * sizeof(RowType) + pats.size() * sizeof(TilePat));
  if (getenv("GG_TILE_STATS"))
    fprintf(stderr, "[gg_tile] %lld tiles, %lld supertiles, %lld tile patterns, %lld rows on tile borders, %lld row types; slots: %lld "
                    "complete + %lld on borders of %lld; border cells %lld of %lld; types + patterns %.3f MB, row bases %.2f MB\n",
            (long long)n_tiles, (long long)plan->n_super, (long long)plan->n_pat, (long long)plan->n_grows,
            (long long)plan->n_types, (long long)plan->complete_slots, (long long)plan->group_slots, (long long)A->nnz,
            (long long)plan->border_cells, (long long)nc, plan->pattern_bytes / 1e6, trow_base.size() * 8 / 1e6);
  return plan;
}

}  // namespace gg
