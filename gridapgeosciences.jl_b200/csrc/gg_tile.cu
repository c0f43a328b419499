// Host-side construction of the tile plan (gg_tile.cuh): one-off, part of the symbolic phase of a matrix.
#include <algorithm>
#include <array>
#include <cstring>
#include <map>
#include <unordered_map>
#include "gg_common.cuh"
#include "gg_tile.cuh"

namespace gg {

namespace {

constexpr int TX = 16, TY = 8;

struct Src4 { uint16_t s[4]; };

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
  T.n_super = (int32_t)n_super;
  T.inst_ptr = inst_ptr.p; T.inst_tile = inst_tile.p;
  T.tile_cells = tile_cells.p; T.tile_pat = tile_pat.p; T.tile_eoff = tile_eoff.p; T.tile_voff = tile_voff.p;
  T.pats = pats.p; T.tile_row0 = tile_row0.p; T.trow_base = trow_base.p; T.trow_dof = trow_dof.p; T.rowvsrc = rowvsrc.p; T.emeta = emeta.p; T.esrc = esrc.p; T.bidx = bidx.p;
  T.tg_ptr = tg_ptr.p; T.tg_idx = tg_idx.p; T.groups = groups.p; T.gcount = gcount.p; T.gpats = gpats.p;
  T.grow_base = grow_base.p; T.grow_dof = grow_dof.p; T.g_rowvsrc = g_rowvsrc.p; T.g_emeta = g_emeta.p; T.g_esrc = g_esrc.p;
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
  if (A->test != A->trial || m->dim != 2 || M > 16 || NP > 64 || nc == 0 || A->nnz == 0) return nullptr;
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
    for (int32_t k = dptr[d]; k < dptr[d + 1]; ++k) {
      const int32_t t = cell_tile[dsrc[k] / M];
      bool seen = false;
      for (int q = 0; q < dnt[d]; ++q) seen |= dtiles[d * 4 + q] == t;
      if (seen) continue;
      if (dnt[d] == 4) return nullptr;   // more than four tiles around a DOF: not a quadrilateral-like tiling
      dtiles[d * 4 + dnt[d]++] = t;
    }
    std::sort(&dtiles[d * 4], &dtiles[d * 4] + dnt[d]);
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
  std::vector<int32_t> nborder(n_tiles, 0), tile_eoff(n_tiles), tile_voff(n_tiles);
  for (int64_t t = 0; t < n_tiles; ++t)
    for (int lc = 0; lc < TILE_CELLS; ++lc) {
      const int32_t c = tile_cells[t * TILE_CELLS + lc];
      if (c < 0) continue;
      bool border = false;
      for (int i = 0; i < M; ++i) {
        const int32_t d = dofs[(int64_t)c * M + i];
        border |= d >= 0 && dnt[d] > 1;
      }
      if (border) bidx_tile[t * TILE_CELLS + lc] = (uint8_t)nborder[t]++;
    }
  {
    int64_t eo = 0, vo = 0;
    for (int64_t t = 0; t < n_tiles; ++t) {
      tile_eoff[t] = (int32_t)eo; tile_voff[t] = (int32_t)vo;
      eo += (int64_t)nborder[t] * NP; vo += (int64_t)nborder[t] * M;
      plan->border_cells += nborder[t];
    }
    if (eo >= (int64_t)INT32_MAX) return nullptr;
    plan->ebuf_b.alloc((size_t)std::max<int64_t>(eo, 1));
    plan->vbuf_b.alloc((size_t)std::max<int64_t>(vo, 1));
  }

  // ---- complete rows of every tile -> patterns ----------------------------------------------------------------------------------
  std::vector<TilePat> pats;
  std::vector<int32_t> tile_row0(n_tiles), trow_base, trow_dof;
  std::vector<Src4> rowvsrc, esrc;
  std::vector<uint32_t> emeta;
  std::vector<uint8_t> bidx;
  std::vector<int32_t> tile_pat(n_tiles);
  PatStore store;
  std::vector<std::vector<int32_t>> tile_rows(n_tiles);
  for (int64_t d = 0; d < nd; ++d)
    if (dnt[d] == 1) tile_rows[dtiles[d * 4]].push_back((int32_t)d);   // ascending d
  for (int64_t t = 0; t < n_tiles; ++t) {
    const auto& rows = tile_rows[t];
    if ((int)rows.size() > TILE_ROWCAP) return nullptr;
    std::vector<Src4> p_rowvsrc, p_esrc;
    std::vector<uint32_t> p_emeta;
    tile_row0[t] = (int32_t)trow_base.size();
    for (size_t k = 0; k < rows.size(); ++k) {
      const int32_t d = rows[k];
      const int len = rp[d + 1] - rp[d];
      if (len > 255) return nullptr;
      trow_base.push_back(rp[d]);
      trow_dof.push_back(d);
      Src4 vs{{TILE_NONE, TILE_NONE, TILE_NONE, TILE_NONE}};
      if (dptr[d + 1] - dptr[d] > 4) return nullptr;
      for (int32_t q = dptr[d]; q < dptr[d + 1]; ++q)
        vs.s[q - dptr[d]] = (uint16_t)(((dsrc[q] % M) << 7) | cell_lc[dsrc[q] / M]);
      p_rowvsrc.push_back(vs);
      const size_t e0 = p_esrc.size();
      p_esrc.resize(e0 + len, Src4{{TILE_NONE, TILE_NONE, TILE_NONE, TILE_NONE}});
      for (int j = 0; j < len; ++j) p_emeta.push_back((uint32_t)((k << 8) | (uint32_t)j));
      for (int32_t q = dptr[d]; q < dptr[d + 1]; ++q) {        // ascending cell
        const int32_t c = dsrc[q] / M, i = dsrc[q] % M;
        for (int j2 = 0; j2 < M; ++j2) {
          const int32_t col = dofs[(int64_t)c * M + j2];
          if (col < 0) continue;
          const int pos = col_pos(d, col);
          if (pos < 0) return nullptr;
          Src4& e = p_esrc[e0 + pos];
          int w = 0;
          while (w < 4 && e.s[w] != TILE_NONE) ++w;
          if (w == 4) return nullptr;
          e.s[w] = (uint16_t)(plane_of(M, i, j2) * TILE_CELLS + cell_lc[c]);
        }
      }
    }
    std::vector<uint8_t> p_bidx(&bidx_tile[t * TILE_CELLS], &bidx_tile[t * TILE_CELLS] + TILE_CELLS);
    Blob key2;
    key2.put(p_rowvsrc); key2.put(p_emeta); key2.put(p_esrc); key2.put(p_bidx);
    bool is_new = false;
    const int32_t id = store.find_or_add(std::move(key2), (int32_t)pats.size(), &is_new);
    tile_pat[t] = id;
    if (is_new) {
      TilePat P{(int32_t)rowvsrc.size(), (int32_t)p_rowvsrc.size(), (int32_t)emeta.size(), (int32_t)p_emeta.size(), nborder[t], 0};
      pats.push_back(P);
      rowvsrc.insert(rowvsrc.end(), p_rowvsrc.begin(), p_rowvsrc.end());
      emeta.insert(emeta.end(), p_emeta.begin(), p_emeta.end());
      esrc.insert(esrc.end(), p_esrc.begin(), p_esrc.end());
      bidx.insert(bidx.end(), p_bidx.begin(), p_bidx.end());
    }
    plan->complete_slots += (int64_t)p_emeta.size();
    plan->max_rows = std::max(plan->max_rows, (int)p_rowvsrc.size());
  }

  // ---- incomplete rows: groups by the set of tiles around the row ------------------------------------------------------------
  std::map<std::array<int32_t, 4>, int32_t> gid;
  std::vector<std::array<int32_t, 4>> gtiles;
  std::vector<std::vector<int32_t>> grows;
  for (int64_t d = 0; d < nd; ++d) {
    if (dnt[d] <= 1) continue;
    std::array<int32_t, 4> k{dtiles[d * 4], dtiles[d * 4 + 1], dtiles[d * 4 + 2], dtiles[d * 4 + 3]};
    auto it = gid.find(k);
    if (it == gid.end()) {
      it = gid.emplace(k, (int32_t)gtiles.size()).first;
      gtiles.push_back(k);
      grows.emplace_back();
    }
    grows[it->second].push_back((int32_t)d);
  }
  const int64_t ng = (int64_t)gtiles.size();
  std::vector<GroupPat> gpats;
  std::vector<GroupInst> groups(ng);
  std::vector<int32_t> grow_base, grow_dof;
  std::vector<Src4> g_rowvsrc, g_esrc;
  std::vector<uint32_t> g_emeta;
  PatStore gstore;
  std::vector<std::vector<int32_t>> tile_groups(n_tiles);
  for (int64_t g = 0; g < ng; ++g) {
    const auto& tl = gtiles[g];
    const auto& rows = grows[g];
    if ((int)rows.size() > TILE_ROWCAP) return nullptr;
    int need = 0;
    for (int q = 0; q < 4; ++q)
      if (tl[q] >= 0) { ++need; tile_groups[tl[q]].push_back((int32_t)g); }
    auto slot_of = [&](int32_t tile) { for (int q = 0; q < 4; ++q) if (tl[q] == tile) return q; return -1; };
    std::vector<Src4> p_rowvsrc, p_esrc;
    std::vector<uint32_t> p_emeta;
    groups[g].row0 = (int32_t)grow_base.size();
    for (size_t k = 0; k < rows.size(); ++k) {
      const int32_t d = rows[k];
      const int len = rp[d + 1] - rp[d];
      if (len > 255) return nullptr;
      grow_base.push_back(rp[d]);
      grow_dof.push_back(d);
      if (dptr[d + 1] - dptr[d] > 4) return nullptr;
      Src4 vs{{TILE_NONE, TILE_NONE, TILE_NONE, TILE_NONE}};
      for (int32_t q = dptr[d]; q < dptr[d + 1]; ++q) {
        const int32_t c = dsrc[q] / M;
        const int b = bidx_tile[(int64_t)cell_tile[c] * TILE_CELLS + cell_lc[c]];
        vs.s[q - dptr[d]] = (uint16_t)((slot_of(cell_tile[c]) << 11) | (b << 4) | (dsrc[q] % M));
      }
      p_rowvsrc.push_back(vs);
      const size_t e0 = p_esrc.size();
      p_esrc.resize(e0 + len, Src4{{TILE_NONE, TILE_NONE, TILE_NONE, TILE_NONE}});
      for (int j = 0; j < len; ++j) p_emeta.push_back((uint32_t)((k << 8) | (uint32_t)j));
      for (int32_t q = dptr[d]; q < dptr[d + 1]; ++q) {
        const int32_t c = dsrc[q] / M, i = dsrc[q] % M;
        const int b = bidx_tile[(int64_t)cell_tile[c] * TILE_CELLS + cell_lc[c]];
        const int ts = slot_of(cell_tile[c]);
        for (int j2 = 0; j2 < M; ++j2) {
          const int32_t col = dofs[(int64_t)c * M + j2];
          if (col < 0) continue;
          const int pos = col_pos(d, col);
          if (pos < 0) return nullptr;
          Src4& e = p_esrc[e0 + pos];
          int w = 0;
          while (w < 4 && e.s[w] != TILE_NONE) ++w;
          if (w == 4) return nullptr;
          e.s[w] = (uint16_t)((ts << 13) | (b << 6) | plane_of(M, i, j2));
        }
      }
    }
    Blob key2;
    key2.put(p_rowvsrc); key2.put(p_emeta); key2.put(p_esrc);
    bool is_new = false;
    const int32_t id = gstore.find_or_add(std::move(key2), (int32_t)gpats.size(), &is_new);
    if (is_new) {
      gpats.push_back(GroupPat{(int32_t)g_rowvsrc.size(), (int32_t)p_rowvsrc.size(), (int32_t)g_emeta.size(), (int32_t)p_emeta.size()});
      g_rowvsrc.insert(g_rowvsrc.end(), p_rowvsrc.begin(), p_rowvsrc.end());
      g_emeta.insert(g_emeta.end(), p_emeta.begin(), p_emeta.end());
      g_esrc.insert(g_esrc.end(), p_esrc.begin(), p_esrc.end());
    }
    groups[g].pat = id; groups[g].need = need; groups[g].pad = 0;
    for (int q = 0; q < 4; ++q) {
      groups[g].eoff[q] = tl[q] >= 0 ? tile_eoff[tl[q]] : 0;
      groups[g].voff[q] = tl[q] >= 0 ? tile_voff[tl[q]] : 0;
      groups[g].nb[q] = tl[q] >= 0 ? nborder[tl[q]] : 0;
    }
    plan->group_slots += (int64_t)p_emeta.size();
  }
  std::vector<int32_t> tg_ptr(n_tiles + 1, 0), tg_idx;
  for (int64_t t = 0; t < n_tiles; ++t) {
    tg_idx.insert(tg_idx.end(), tile_groups[t].begin(), tile_groups[t].end());
    tg_ptr[t + 1] = (int32_t)tg_idx.size();
  }

  // ---- supertiles: tiles that share their geometry (geometry classes) are served by one CTA --------------------------------------
  std::vector<int32_t> inst_ptr{0}, inst_tile;
  if (classes) {
    std::map<std::vector<int32_t>, std::vector<int32_t>> by_geom;
    std::vector<const std::vector<int32_t>*> first_seen;
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
  plan->n_pat = (int64_t)pats.size(); plan->n_groups = ng; plan->n_gpat = (int64_t)gpats.size();

  // ---- upload ------------------------------------------------------------------------------------------------------------------------
  auto up4 = [&](DevBuf<ushort4>& d, const std::vector<Src4>& h) {
    static_assert(sizeof(Src4) == sizeof(ushort4), "source quad layout");
    d.upload((const ushort4*)h.data(), h.size(), st);
  };
  plan->inst_ptr.upload(inst_ptr, st); plan->inst_tile.upload(inst_tile, st);
  plan->tile_cells.upload(tile_cells, st); plan->tile_pat.upload(tile_pat, st);
  plan->tile_eoff.upload(tile_eoff, st); plan->tile_voff.upload(tile_voff, st);
  plan->pats.upload(pats, st); up4(plan->rowvsrc, rowvsrc);
  plan->tile_row0.upload(tile_row0, st); plan->trow_base.upload(trow_base, st); plan->trow_dof.upload(trow_dof, st);
  plan->grow_base.upload(grow_base, st); plan->grow_dof.upload(grow_dof, st);
  plan->emeta.upload(emeta, st); up4(plan->esrc, esrc); plan->bidx.upload(bidx, st);
  plan->tg_ptr.upload(tg_ptr, st); plan->tg_idx.upload(tg_idx, st);
  plan->groups.upload(groups, st); plan->gpats.upload(gpats, st);
  up4(plan->g_rowvsrc, g_rowvsrc);
  plan->g_emeta.upload(g_emeta, st); up4(plan->g_esrc, g_esrc);
  plan->gcount.alloc((size_t)std::max<int64_t>(ng, 1));
  GG_CUDA(cudaMemsetAsync(plan->gcount.p, 0, plan->gcount.n * sizeof(int32_t), st));
  GG_CUDA(cudaStreamSynchronize(st));
  plan->pattern_bytes = (int64_t)(rowvsrc.size() * 8 + emeta.size() * 4 + esrc.size() * 8 + bidx.size() +
                                  g_rowvsrc.size() * 8 + g_emeta.size() * 4 + g_esrc.size() * 8);
  if (getenv("GG_TILE_STATS"))
    fprintf(stderr, "[gg_tile] %lld tiles, %lld supertiles, %lld tile patterns, %lld groups (%lld patterns); slots: %lld complete + %lld in "
                    "groups of %lld; border cells %lld of %lld; patterns %.2f MB\n", (long long)n_tiles, (long long)plan->n_super,
            (long long)plan->n_pat, (long long)ng, (long long)plan->n_gpat, (long long)plan->complete_slots, (long long)plan->group_slots,
            (long long)A->nnz, (long long)plan->border_cells, (long long)nc, plan->pattern_bytes / 1e6);
  return plan;
}

}  // namespace gg
