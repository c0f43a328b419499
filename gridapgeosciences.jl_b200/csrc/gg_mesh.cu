// Atlas mesh of the cubed sphere, topology and FE-space numbering (host side, one-off set-up), uploaded
// to the device in the layouts the kernels want.
//
// Reference behaviour reproduced (paths relative to the reference checkout):
//   src/Geometry/CubedSphereMesh.jl:83-116  coarse cube: 8 nodes, cells [1,2,3,4],[3,4,5,6],[2,7,4,6],[8,5,7,6],[1,8,2,7],[1,3,8,5]
//   src/Geometry/AtlasGrids.jl:227-304      refinement by n: parent-major cell order, x-fastest inside a panel,
//                                           chart corners = bilinear panel map of the reference grid k/n
//   src/Distributed/AtlasDistributedDiscreteModels.jl:31-49   linear partition, vertex-adjacent ghost layer
// Gridap conventions restated (not verifiable without Gridap): edge ids by first
// encounter over cells and local edges (1,2),(3,4),(1,3),(2,4); conforming DOFs = vertices, edges, interiors;
// fine node ids = CG-Q_n numbering of the coarse cube; RT facet DOFs negated on the higher-id cell.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <unordered_map>
#include "gg_common.cuh"
#include "gg_vec.cuh"
#include "gg_refel.cuh"

namespace gg {

static const int QUAD_EDGE_V[4][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}};

struct EdgeTopo {
  std::vector<int32_t> cell_edges;  // [nc][4]
  std::vector<int8_t> flip;         // [nc][4]
  std::vector<int32_t> edge_cells;  // [ne][2]
  int64_t n_edges = 0;
};

static EdgeTopo build_edges(int64_t nc, const int32_t* cell_nodes) {
  EdgeTopo T;
  T.cell_edges.resize(nc * 4);
  T.flip.resize(nc * 4);
  struct Ref { uint64_t key; int64_t idx; };
  std::vector<Ref> refs(nc * 4);
  for (int64_t c = 0; c < nc; ++c)
    for (int e = 0; e < 4; ++e) {
      uint32_t a = cell_nodes[c * 4 + QUAD_EDGE_V[e][0]], b = cell_nodes[c * 4 + QUAD_EDGE_V[e][1]];
      uint64_t lo = std::min(a, b), hi = std::max(a, b);
      refs[c * 4 + e] = {(lo << 32) | hi, c * 4 + e};
    }
  std::sort(refs.begin(), refs.end(), [](const Ref& x, const Ref& y) { return x.key != y.key ? x.key < y.key : x.idx < y.idx; });
  // groups of equal key; first member (smallest idx) defines the encounter position
  std::vector<int64_t> group_first;  // encounter position of each group
  std::vector<int64_t> group_of(nc * 4);
  for (size_t i = 0; i < refs.size(); ++i) {
    if (i == 0 || refs[i].key != refs[i - 1].key) group_first.push_back(refs[i].idx);
    group_of[refs[i].idx] = (int64_t)group_first.size() - 1;
  }
  int64_t ne = (int64_t)group_first.size();
  std::vector<int64_t> order(ne);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return group_first[a] < group_first[b]; });
  std::vector<int32_t> rank(ne);
  for (int64_t r = 0; r < ne; ++r) rank[order[r]] = (int32_t)r;
  T.n_edges = ne;
  T.edge_cells.assign(ne * 2, -1);
  std::vector<int32_t> edge_first_node(ne);
  for (int64_t c = 0; c < nc; ++c)
    for (int e = 0; e < 4; ++e) {
      int32_t id = rank[group_of[c * 4 + e]];
      T.cell_edges[c * 4 + e] = id;
      int32_t a = cell_nodes[c * 4 + QUAD_EDGE_V[e][0]];
      if (T.edge_cells[id * 2] < 0) {
        T.edge_cells[id * 2] = (int32_t)c;
        edge_first_node[id] = a;
        T.flip[c * 4 + e] = 0;
      } else {
        GG_REQUIRE(T.edge_cells[id * 2 + 1] < 0, GG_ERR_INVALID, "non-manifold mesh: an edge has more than two cells");
        T.edge_cells[id * 2 + 1] = (int32_t)c;
        T.flip[c * 4 + e] = (a != edge_first_node[id]);
      }
    }
  return T;
}

// Conforming Q_p ids, Gridap local order (vertices, edge interiors edge by edge, cell interior)
static void cg_numbering(int64_t nc, const int32_t* cell_nodes, int64_t n_nodes, const EdgeTopo& T, int p,
                         std::vector<int32_t>& ids, int64_t& ndofs) {
  int m = (p + 1) * (p + 1);
  ids.resize(nc * m);
  int64_t ne = T.n_edges;
  int ni = (p - 1) * (p - 1);
  ndofs = n_nodes + ne * (p - 1) + nc * ni;
  GG_REQUIRE(ndofs < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "more than 2^31 DOFs in one space");
  for (int64_t c = 0; c < nc; ++c) {
    int32_t* o = &ids[c * m];
    int k = 0;
    for (int v = 0; v < 4; ++v) o[k++] = cell_nodes[c * 4 + v];
    for (int e = 0; e < 4; ++e)
      for (int t = 0; t < p - 1; ++t) {
        int tt = T.flip[c * 4 + e] ? (p - 2 - t) : t;
        o[k++] = (int32_t)(n_nodes + (int64_t)T.cell_edges[c * 4 + e] * (p - 1) + tt);
      }
    for (int t = 0; t < ni; ++t) o[k++] = (int32_t)(n_nodes + ne * (p - 1) + c * ni + t);
  }
}

static void ensure_topology(gg_mesh* m) {
  if (m->topology_built) return;
  EdgeTopo T = build_edges(m->n_cells, m->cell_nodes.data());
  m->cell_edges = std::move(T.cell_edges);
  m->cell_edge_flip = std::move(T.flip);
  m->edge_cells = std::move(T.edge_cells);
  m->n_edges = T.n_edges;
  m->topology_built = true;
}

void upload_geometry(gg_mesh* m) {
  if (!m->ctx) return;  // host-only mesh (partition plans)
  int64_t nc = m->n_cells;
  std::vector<double> x0(nc), y0(nc), hx(nc), hy(nc);
  for (int64_t c = 0; c < nc; ++c) {
    const double* q = &m->chart_coords[c * 8];  // BL, BR, TL, TR
    double tol = 1e-13 * (std::fabs(q[0]) + std::fabs(q[6]) + 1.0);
    bool ok = std::fabs(q[0] - q[4]) <= tol && std::fabs(q[2] - q[6]) <= tol &&   // x: BL==TL, BR==TR
              std::fabs(q[1] - q[3]) <= tol && std::fabs(q[5] - q[7]) <= tol;     // y: BL==BR, TL==TR
    GG_REQUIRE(ok, GG_ERR_UNSUPPORTED, "cell " + std::to_string(c) + " is not an axis-aligned rectangle in chart space");
    x0[c] = q[0]; y0[c] = q[1];
    hx[c] = q[2] - q[0]; hy[c] = q[5] - q[1];
    GG_REQUIRE(hx[c] > 0 && hy[c] > 0, GG_ERR_UNSUPPORTED, "cell " + std::to_string(c) + " has non-positive extent");
  }
  cudaStream_t s = m->ctx->stream;
  m->d_x0.upload(x0, s); m->d_y0.upload(y0, s); m->d_hx.upload(hx, s); m->d_hy.upload(hy, s);
  m->d_chart.upload(m->cell_to_chart, s);
  // distinct (lo, h) intervals per direction (bitwise comparison: cells of one column are produced by the same
  // expression; a caller-supplied mesh with last-bit differences just gets more intervals)
  auto intervals = [&](const std::vector<double>& lo, const std::vector<double>& h, std::vector<int32_t>& id,
                       std::vector<double>& ulo, std::vector<double>& uh) {
    std::vector<int64_t> order(nc);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return lo[a] != lo[b] ? lo[a] < lo[b] : h[a] < h[b]; });
    id.resize(nc);
    ulo.clear(); uh.clear();
    for (int64_t k = 0; k < nc; ++k) {
      int64_t c = order[k];
      if (k == 0 || lo[c] != ulo.back() || h[c] != uh.back()) { ulo.push_back(lo[c]); uh.push_back(h[c]); }
      id[c] = (int32_t)ulo.size() - 1;
    }
  };
  std::vector<int32_t> ix, iy;
  std::vector<double> xlo, xh, ylo, yh;
  intervals(x0, hx, ix, xlo, xh);
  intervals(y0, hy, iy, ylo, yh);
  m->n_xint = (int64_t)xlo.size(); m->n_yint = (int64_t)ylo.size();
  m->d_ix.upload(ix, s); m->d_iy.upload(iy, s);
  m->h_ix = ix; m->h_iy = iy;
  m->d_xlo.upload(xlo, s); m->d_xh.upload(xh, s); m->d_ylo.upload(ylo, s); m->d_yh.upload(yh, s);
  m->tan_tabs.clear();
  if (!m->cell_owned.empty()) m->d_cell_owned.upload(m->cell_owned, s);
  GG_CUDA(cudaStreamSynchronize(s));
}

static const int32_t COARSE_CELLS[6][4] = {{0, 1, 2, 3}, {2, 3, 4, 5}, {1, 6, 3, 5}, {7, 4, 6, 5}, {0, 7, 1, 6}, {0, 2, 7, 4}};

gg_mesh* build_cubed_sphere(gg_ctx* ctx, int n, double radius, int ordering) {
  std::unique_ptr<gg_mesh> m(new gg_mesh);
  m->ctx = ctx; m->dim = 2; m->n = n; m->radius = radius;
  int64_t nc = 6LL * n * n;
  GG_REQUIRE(nc * 25 < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "mesh too large for Int32 ids");
  m->n_cells = nc;
  // fine nodes = CG-Q_n DOFs of the coarse cube
  EdgeTopo CT = build_edges(6, &COARSE_CELLS[0][0]);
  std::vector<int32_t> coarse_ids;
  int64_t nn;
  cg_numbering(6, &COARSE_CELLS[0][0], 8, CT, n, coarse_ids, nn);
  m->n_nodes = nn;
  std::vector<int32_t> l2x = quad_local_to_lex(n);
  std::vector<int32_t> lut((n + 1) * (n + 1));
  for (size_t k = 0; k < l2x.size(); ++k) lut[l2x[k]] = (int32_t)k;
  int mloc = (n + 1) * (n + 1);
  m->cell_nodes.resize(nc * 4);
  m->cell_to_chart.resize(nc);
  m->chart_coords.resize(nc * 8);
  const double H = M_PI / 4;  // CUBE_HALF_EDGE, src/Geometry/CubedSphereMesh.jl:23
  auto psi = [&](double xi) { return (1.0 - xi) * (-H) + xi * H; };
  // ordering 0: x-fastest inside a panel (one-shot refinement, AtlasGrids.jl:254-280).  ordering 1 / 2: Morton (z-order, x bit
  // lowest) inside a panel = the cell order of successive `refine` calls (children BL, BR, TL, TR of each parent stay
  // together, test/AtlasDiscreteModels/seq/RefinementOrderingTests.jl:93-100) and of a uniformly refined p4est tree
  // (src/Distributed/AtlasOctreeDistributedDiscreteModels.jl:164-208); needs n = 2^l.
  auto morton = [&](int i, int j) {
    int64_t k = 0;
    for (int b = 0; (1 << b) < n; ++b) k |= ((int64_t)((i >> b) & 1) << (2 * b)) | ((int64_t)((j >> b) & 1) << (2 * b + 1));
    return k;
  };
  for (int p = 0; p < 6; ++p)
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        int64_t c = (int64_t)p * n * n + (ordering == 0 ? (int64_t)j * n + i : morton(i, j));
        const int32_t* cid = &coarse_ids[(size_t)p * mloc];
        m->cell_nodes[c * 4 + 0] = cid[lut[i + (n + 1) * j]];
        m->cell_nodes[c * 4 + 1] = cid[lut[i + 1 + (n + 1) * j]];
        m->cell_nodes[c * 4 + 2] = cid[lut[i + (n + 1) * (j + 1)]];
        m->cell_nodes[c * 4 + 3] = cid[lut[i + 1 + (n + 1) * (j + 1)]];
        m->cell_to_chart[c] = p;
        double xa = psi((double)i / n), xb = psi((double)(i + 1) / n);
        double ya = psi((double)j / n), yb = psi((double)(j + 1) / n);
        double* q = &m->chart_coords[c * 8];
        q[0] = xa; q[1] = ya; q[2] = xb; q[3] = ya; q[4] = xa; q[5] = yb; q[6] = xb; q[7] = yb;
      }
  upload_geometry(m.get());
  return m.release();
}

// ---- 3-D shell ------------------------------------------------------------------------------------------------------------
// local vertices of the 12 edges and 6 faces of the hexahedron, Gridap order
static const int HEX_EDGE_V[12][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}, {0, 2}, {1, 3}, {4, 6}, {5, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int HEX_FACE_V[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 2, 4, 6}, {1, 3, 5, 7}};
// panel permutation table of the gnomonic map (src/Fields/CubedSphereMap.jl:19-26), 0-based component, sign
static const int PERM_J[6][3] = {{0, 1, 2}, {2, 1, 0}, {1, 0, 2}, {0, 2, 1}, {1, 2, 0}, {2, 0, 1}};
static const int PERM_S[6][3] = {{1, 1, 1}, {-1, 1, 1}, {-1, 1, 1}, {-1, 1, 1}, {-1, 1, -1}, {-1, -1, 1}};

// ids of entities given by up to 4 vertex ids each, numbered by first encounter (same scheme as build_edges)
static int64_t number_entities(int64_t n_refs, int nv, const std::vector<int32_t>& verts, std::vector<int32_t>& ids) {
  struct Ref { int32_t v[4]; int64_t idx; };
  std::vector<Ref> refs(n_refs);
  for (int64_t i = 0; i < n_refs; ++i) {
    int32_t tmp[4] = {-1, -1, -1, -1};
    for (int k = 0; k < nv; ++k) tmp[k] = verts[i * nv + k];
    std::sort(tmp, tmp + nv);
    for (int k = 0; k < 4; ++k) refs[i].v[k] = tmp[k];
    refs[i].idx = i;
  }
  auto less = [](const Ref& a, const Ref& b) {
    for (int k = 0; k < 4; ++k) if (a.v[k] != b.v[k]) return a.v[k] < b.v[k];
    return a.idx < b.idx;
  };
  std::sort(refs.begin(), refs.end(), less);
  std::vector<int64_t> group_first, group_of(n_refs);
  for (int64_t i = 0; i < n_refs; ++i) {
    bool fresh = i == 0;
    if (!fresh) for (int k = 0; k < 4; ++k) fresh = fresh || refs[i].v[k] != refs[i - 1].v[k];
    if (fresh) group_first.push_back(refs[i].idx);
    group_of[refs[i].idx] = (int64_t)group_first.size() - 1;
  }
  int64_t ne = (int64_t)group_first.size();
  std::vector<int64_t> order(ne);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return group_first[a] < group_first[b]; });
  std::vector<int32_t> rank(ne);
  for (int64_t r = 0; r < ne; ++r) rank[order[r]] = (int32_t)r;
  ids.resize(n_refs);
  for (int64_t i = 0; i < n_refs; ++i) ids[i] = rank[group_of[i]];
  return ne;
}

static void ensure_hex_topology(gg_mesh* m) {
  if (m->topology_built) return;
  int64_t nc = m->n_cells;
  std::vector<int32_t> ev(nc * 12 * 2), fv(nc * 6 * 4);
  for (int64_t c = 0; c < nc; ++c) {
    for (int e = 0; e < 12; ++e)
      for (int k = 0; k < 2; ++k) ev[(c * 12 + e) * 2 + k] = m->cell_nodes8[c * 8 + HEX_EDGE_V[e][k]];
    for (int f = 0; f < 6; ++f)
      for (int k = 0; k < 4; ++k) fv[(c * 6 + f) * 4 + k] = m->cell_nodes8[c * 8 + HEX_FACE_V[f][k]];
  }
  m->n_edges = number_entities(nc * 12, 2, ev, m->cell_edges12);
  m->n_faces = number_entities(nc * 6, 4, fv, m->cell_faces6);
  m->topology_built = true;
}

// AtlasDiscreteModel(CubedSphereWithThicknessMesh(radius, thickness), l): n x n cells per panel, L cells through the shell
// (src/Geometry/CubedSphereWithThicknessMesh.jl:33-124, AtlasGrids.jl:227-304; the reference refines isotropically, L = n).
// Nodes are identified through their integer position on the cube-surface lattice and numbered by first encounter.
// ordering 0: gamma fastest, then alpha, then beta inside a panel.  ordering 2: the cell order of a uniformly refined p8est
// forest (src/Distributed/AtlasOctreeDistributedDiscreteModels.jl:164-208): Morton order inside each of the 6 trees, octant
// coordinates (x, y, z) = (gamma, alpha, beta), x bit lowest (n = L = 2^l)
gg_mesh* build_cubed_sphere_shell(gg_ctx* ctx, int n, int L, double radius, double thickness, int ordering) {
  std::unique_ptr<gg_mesh> m(new gg_mesh);
  m->ctx = ctx; m->dim = 3; m->n = n; m->n_layers = L; m->radius = radius; m->thickness = thickness;
  int64_t nc = 6LL * n * n * L;
  GG_REQUIRE(nc * 27 < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "shell mesh too large for Int32 ids");
  m->n_cells = nc;
  m->cell_nodes8.resize(nc * 8); m->cell_to_chart.resize(nc); m->cell_col.resize(nc); m->cell_layer.resize(nc);
  std::unordered_map<uint64_t, int32_t> nodes;
  nodes.reserve((size_t)(6LL * n * n + 2) * (L + 1) * 2);
  int64_t c = 0;
  const int64_t per_panel = (int64_t)n * n * L;
  for (int p = 0; p < 6; ++p)
    for (int64_t t = 0; t < per_panel; ++t) {
      int kg, ia, jb;
      if (ordering == 0) {
        kg = (int)(t % L); ia = (int)((t / L) % n); jb = (int)(t / ((int64_t)L * n));
      } else {
        kg = ia = jb = 0;
        for (int b = 0; (1 << b) < n; ++b) {
          kg |= (int)((t >> (3 * b)) & 1) << b;
          ia |= (int)((t >> (3 * b + 1)) & 1) << b;
          jb |= (int)((t >> (3 * b + 2)) & 1) << b;
        }
      }
        {
          for (int v = 0; v < 8; ++v) {
            int dg = v & 1, da = (v >> 1) & 1, db = (v >> 2) & 1;
            int h[3] = {n, 2 * (ia + da) - n, 2 * (jb + db) - n};
            uint64_t key = 0;
            for (int k = 0; k < 3; ++k) key = key * (uint64_t)(2 * n + 1) + (uint64_t)(PERM_S[p][k] * h[PERM_J[p][k]] + n);
            key = key * (uint64_t)(L + 1) + (uint64_t)(kg + dg);
            auto it = nodes.find(key);
            if (it == nodes.end()) it = nodes.emplace(key, (int32_t)nodes.size()).first;
            m->cell_nodes8[c * 8 + v] = it->second;
          }
          m->cell_to_chart[c] = p;
          m->cell_col[c] = (int32_t)((int64_t)p * n * n + (int64_t)jb * n + ia);
          m->cell_layer[c] = kg;
          ++c;
        }
    }
  m->n_nodes = (int64_t)nodes.size();
  m->surface.reset(build_cubed_sphere(ctx, n, radius, 0));
  if (ctx) {
    m->d_col.upload(m->cell_col, ctx->stream); m->d_layer.upload(m->cell_layer, ctx->stream);
    m->d_chart.upload(m->cell_to_chart, ctx->stream);
    GG_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return m.release();
}

// Gridap local node of Q_p on the hexahedron (vertices, edges, faces, interior) -> lexicographic i0 + nb (i1 + nb i2)
std::vector<int32_t> hex_local_to_lex(int p) {
  int nb = p + 1;
  if (p == 0) return {0};
  std::vector<int32_t> out;
  auto lex = [&](int i0, int i1, int i2) { return i0 + nb * (i1 + nb * i2); };
  for (int v = 0; v < 8; ++v) out.push_back(lex((v & 1) * p, ((v >> 1) & 1) * p, ((v >> 2) & 1) * p));
  // edges: spanning axis S, the fixed axes take the end-point values (first fixed axis fastest)
  for (int S = 0; S < 3; ++S)
    for (int bits = 0; bits < 4; ++bits) {
      int fa[2], k = 0;
      for (int a = 0; a < 3; ++a) if (a != S) fa[k++] = a;
      for (int t = 1; t < p; ++t) {
        int mi[3];
        mi[S] = t; mi[fa[0]] = (bits & 1) * p; mi[fa[1]] = ((bits >> 1) & 1) * p;
        out.push_back(lex(mi[0], mi[1], mi[2]));
      }
    }
  // faces: spanning axes (0,1) z fixed, (0,2) y fixed, (1,2) x fixed; first spanning axis fastest
  const int SP[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  const int FX[3] = {2, 1, 0};
  for (int f = 0; f < 3; ++f)
    for (int bit = 0; bit < 2; ++bit)
      for (int t1 = 1; t1 < p; ++t1)
        for (int t0 = 1; t0 < p; ++t0) {
          int mi[3];
          mi[SP[f][0]] = t0; mi[SP[f][1]] = t1; mi[FX[f]] = bit * p;
          out.push_back(lex(mi[0], mi[1], mi[2]));
        }
  for (int t2 = 1; t2 < p; ++t2)
    for (int t1 = 1; t1 < p; ++t1)
      for (int t0 = 1; t0 < p; ++t0) out.push_back(lex(t0, t1, t2));
  return out;
}

}  // namespace gg

using namespace gg;

extern "C" {

int gg_mesh_cubed_sphere(gg_ctx* ctx, int dim, int n, int n_layers, double radius, double thickness, int ordering,
                         gg_mesh** out) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && out, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(dim == 2 || dim == 3, GG_ERR_INVALID, "dim must be 2 (surface) or 3 (shell)");
  GG_REQUIRE(n >= 1, GG_ERR_INVALID, "n must be >= 1");
  GG_REQUIRE(radius > 0, GG_ERR_INVALID, "radius must be positive");
  GG_REQUIRE(ordering >= 0 && ordering <= 2, GG_ERR_INVALID, "ordering must be 0 (one-shot), 1 (recursive refinement) or 2 (Morton / p4est)");
  if (ordering != 0) {
    GG_REQUIRE(dim == 2 || ordering == 2, GG_ERR_UNSUPPORTED, "on the 3-D shell orderings 0 and 2 (p8est Morton) are implemented");
    GG_REQUIRE((n & (n - 1)) == 0, GG_ERR_INVALID, "recursive / Morton ordering needs n = 2^l cells per panel edge");
    if (dim == 3) GG_REQUIRE(n_layers == 0 || n_layers == n, GG_ERR_INVALID, "the p8est ordering refines isotropically: n_layers = n");
  }
  GG_CUDA(cudaSetDevice(ctx->device));
  if (dim == 3) {
    GG_REQUIRE(thickness > 0, GG_ERR_INVALID, "shell thickness must be positive");
    GG_REQUIRE(n_layers >= 0, GG_ERR_INVALID, "n_layers must be >= 0 (0: isotropic refinement, n layers)");
    *out = build_cubed_sphere_shell(ctx, n, n_layers > 0 ? n_layers : n, radius, thickness, ordering);
  } else {
    *out = build_cubed_sphere(ctx, n, radius, ordering);
  }
  GG_API_END
}

int gg_mesh_from_arrays(gg_ctx* ctx, int dim, int64_t n_cells, int64_t n_nodes, const int32_t* cell_to_chart,
                        const double* cell_chart_coords, const int32_t* cell_node_ids, double radius, double thickness,
                        gg_mesh** out) {
  GG_API_BEGIN
  GG_REQUIRE(ctx && out && cell_to_chart && cell_chart_coords && cell_node_ids, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(dim == 2, GG_ERR_UNSUPPORTED, "only dim=2 is implemented");
  GG_REQUIRE(n_cells >= 0 && n_nodes >= 0, GG_ERR_INVALID, "negative size");
  GG_REQUIRE(radius > 0, GG_ERR_INVALID, "radius must be positive");
  (void)thickness;
  GG_CUDA(cudaSetDevice(ctx->device));
  std::unique_ptr<gg_mesh> m(new gg_mesh);
  m->ctx = ctx; m->dim = 2; m->n = 0; m->radius = radius;
  m->n_cells = n_cells; m->n_nodes = n_nodes;
  m->cell_nodes.resize(n_cells * 4);
  m->cell_to_chart.resize(n_cells);
  for (int64_t i = 0; i < n_cells * 4; ++i) {
    GG_REQUIRE(cell_node_ids[i] >= 1 && cell_node_ids[i] <= n_nodes, GG_ERR_INVALID, "cell_node_ids out of range (1-based)");
    m->cell_nodes[i] = cell_node_ids[i] - 1;
  }
  for (int64_t i = 0; i < n_cells; ++i) {
    GG_REQUIRE(cell_to_chart[i] >= 1 && cell_to_chart[i] <= 6, GG_ERR_INVALID, "cell_to_chart must be in 1..6");
    m->cell_to_chart[i] = cell_to_chart[i] - 1;
  }
  m->chart_coords.assign(cell_chart_coords, cell_chart_coords + n_cells * 8);
  upload_geometry(m.get());
  *out = m.release();
  GG_API_END
}

int gg_mesh_info(gg_mesh* m, int64_t* n_cells, int64_t* n_nodes, int64_t* n_edges) {
  GG_API_BEGIN
  GG_REQUIRE(m, GG_ERR_INVALID, "null mesh");
  if (n_cells) *n_cells = m->n_cells;
  if (n_nodes) *n_nodes = m->n_nodes;
  if (n_edges) { if (m->dim == 3) ensure_hex_topology(m); else ensure_topology(m); *n_edges = m->n_edges; }
  GG_API_END
}

int gg_mesh_get_cell_nodes(gg_mesh* m, int32_t* out) {
  GG_API_BEGIN
  GG_REQUIRE(m && out, GG_ERR_INVALID, "null argument");
  if (m->dim == 3) {
    for (size_t i = 0; i < m->cell_nodes8.size(); ++i) out[i] = m->cell_nodes8[i] + 1;
    return GG_OK;
  }
  for (size_t i = 0; i < m->cell_nodes.size(); ++i) out[i] = m->cell_nodes[i] + 1;
  GG_API_END
}

int gg_mesh_get_cell_to_chart(gg_mesh* m, int32_t* out) {
  GG_API_BEGIN
  GG_REQUIRE(m && out, GG_ERR_INVALID, "null argument");
  for (size_t i = 0; i < m->cell_to_chart.size(); ++i) out[i] = m->cell_to_chart[i] + 1;
  GG_API_END
}

int gg_mesh_get_chart_coords(gg_mesh* m, double* out) {
  GG_API_BEGIN
  GG_REQUIRE(m && out, GG_ERR_INVALID, "null argument");
  if (m->dim == 3) {
    // [n_cells][8][3], coordinates (gamma, alpha, beta)
    const double H = M_PI / 4;
    for (int64_t c = 0; c < m->n_cells; ++c) {
      int col = m->cell_col[c] % (m->n * m->n), ia = col % m->n, jb = col / m->n, kg = m->cell_layer[c];
      for (int v = 0; v < 8; ++v) {
        out[(c * 8 + v) * 3 + 0] = (double)(kg + (v & 1)) / m->n_layers;
        out[(c * 8 + v) * 3 + 1] = -H + (ia + ((v >> 1) & 1)) * 2 * H / m->n;
        out[(c * 8 + v) * 3 + 2] = -H + (jb + ((v >> 2) & 1)) * 2 * H / m->n;
      }
    }
    return GG_OK;
  }
  std::copy(m->chart_coords.begin(), m->chart_coords.end(), out);
  GG_API_END
}

void gg_mesh_destroy(gg_mesh* m) { delete m; }

int gg_partition_range(int64_t n_cells, int nparts, int part, int64_t* first, int64_t* last) {
  GG_API_BEGIN
  GG_REQUIRE(first && last && nparts >= 1 && part >= 0 && part < nparts && n_cells >= 0, GG_ERR_INVALID, "bad partition request");
  // part(i) = floor(i * P / Nc) (0-based i)  =>  first cell of part p = ceil(p * Nc / P)
  auto lo = [&](int64_t p) { return (p * n_cells + nparts - 1) / nparts; };
  *first = lo(part);
  *last = lo(part + 1);
  GG_API_END
}

int gg_mesh_restrict(gg_mesh* m, int64_t n_sub, const int64_t* cells, gg_mesh** out) {
  GG_API_BEGIN
  GG_REQUIRE(m && out && (cells || n_sub == 0), GG_ERR_INVALID, "null argument");
  if (m->dim == 3) {
    // sub-mesh of the shell: the listed hexahedra with global node ids; the column geometry stays that of the full surface
    GG_CUDA(cudaSetDevice(m->ctx->device));
    std::unique_ptr<gg_mesh> s3(new gg_mesh);
    s3->ctx = m->ctx; s3->dim = 3; s3->n = m->n; s3->n_layers = m->n_layers; s3->radius = m->radius; s3->thickness = m->thickness;
    s3->n_cells = n_sub; s3->n_nodes = m->n_nodes;
    s3->cell_nodes8.resize(n_sub * 8); s3->cell_to_chart.resize(n_sub); s3->cell_col.resize(n_sub); s3->cell_layer.resize(n_sub);
    for (int64_t i = 0; i < n_sub; ++i) {
      int64_t c = cells[i];
      GG_REQUIRE(c >= 0 && c < m->n_cells, GG_ERR_INVALID, "cell id out of range");
      GG_REQUIRE(i == 0 || cells[i - 1] < c, GG_ERR_INVALID, "cells must be strictly ascending");
      std::copy(&m->cell_nodes8[c * 8], &m->cell_nodes8[c * 8] + 8, &s3->cell_nodes8[i * 8]);
      s3->cell_to_chart[i] = m->cell_to_chart[c]; s3->cell_col[i] = m->cell_col[c]; s3->cell_layer[i] = m->cell_layer[c];
    }
    s3->surface.reset(build_cubed_sphere(m->ctx, m->n, m->radius, 0));
    s3->d_col.upload(s3->cell_col, m->ctx->stream); s3->d_layer.upload(s3->cell_layer, m->ctx->stream);
    s3->d_chart.upload(s3->cell_to_chart, m->ctx->stream);
    GG_CUDA(cudaStreamSynchronize(m->ctx->stream));
    *out = s3.release();
    return GG_OK;
  }
  GG_CUDA(cudaSetDevice(m->ctx->device));
  std::unique_ptr<gg_mesh> s(new gg_mesh);
  s->ctx = m->ctx; s->dim = m->dim; s->n = 0; s->radius = m->radius; s->thickness = m->thickness;
  s->n_cells = n_sub; s->n_nodes = m->n_nodes;
  s->cell_nodes.resize(n_sub * 4); s->cell_to_chart.resize(n_sub); s->chart_coords.resize(n_sub * 8);
  for (int64_t i = 0; i < n_sub; ++i) {
    int64_t c = cells[i];
    GG_REQUIRE(c >= 0 && c < m->n_cells, GG_ERR_INVALID, "cell id out of range");
    GG_REQUIRE(i == 0 || cells[i - 1] < c, GG_ERR_INVALID, "cells must be strictly ascending");
    std::copy(&m->cell_nodes[c * 4], &m->cell_nodes[c * 4] + 4, &s->cell_nodes[i * 4]);
    s->cell_to_chart[i] = m->cell_to_chart[c];
    std::copy(&m->chart_coords[c * 8], &m->chart_coords[c * 8] + 8, &s->chart_coords[i * 8]);
  }
  upload_geometry(s.get());
  *out = s.release();
  GG_API_END
}

int gg_mesh_ghost_cells(gg_mesh* m, int64_t first, int64_t last, int64_t* n_ghost, int64_t* out) {
  GG_API_BEGIN
  GG_REQUIRE(m && n_ghost, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(first >= 0 && first <= last && last <= m->n_cells, GG_ERR_INVALID, "bad owned range");
  const int nv = m->dim == 3 ? 8 : 4;
  const std::vector<int32_t>& cn = m->dim == 3 ? m->cell_nodes8 : m->cell_nodes;
  std::vector<uint8_t> touched(m->n_nodes, 0);
  for (int64_t c = first; c < last; ++c)
    for (int v = 0; v < nv; ++v) touched[cn[c * nv + v]] = 1;
  int64_t cnt = 0;
  for (int64_t c = 0; c < m->n_cells; ++c) {
    if (c >= first && c < last) continue;
    bool near = false;
    for (int v = 0; v < nv; ++v) near = near || touched[cn[c * nv + v]];
    if (near) {
      if (out) out[cnt] = c;
      ++cnt;
    }
  }
  *n_ghost = cnt;
  GG_API_END
}

// ---- spaces ---------------------------------------------------------------------------------------

static void finish_space(gg_space* s) {
  gg_mesh* m = s->mesh;
  if (s->kind != GG_SPACE_RT) s->loc2lex = m->dim == 3 ? hex_local_to_lex(s->order) : quad_local_to_lex(s->order);
  if (!m->ctx) return;  // host-only space (partition plans)
  cudaStream_t st = m->ctx->stream;
  s->d_cell_dofs.upload(s->cell_dofs, st);
  if (!s->cell_signs.empty()) s->d_cell_signs.upload(s->cell_signs, st);
  GG_CUDA(cudaStreamSynchronize(st));
  gg::vec_space_finish(s);
}

}  // extern "C"

namespace gg {
void mesh_geometry_classes(gg_mesh* m) {
  if (m->classes_built) return;
  const int64_t nc = m->n_cells;
  m->cls.assign(nc, 0); m->cls_rep.clear();
  if ((int64_t)m->h_ix.size() == nc && m->dim == 2) {
    std::unordered_map<int64_t, int32_t> seen;
    seen.reserve((size_t)nc);
    for (int64_t c = 0; c < nc; ++c) {
      const int64_t key = (int64_t)m->h_iy[c] * m->n_xint + m->h_ix[c];
      auto it = seen.find(key);
      if (it == seen.end()) { it = seen.emplace(key, (int32_t)m->cls_rep.size()).first; m->cls_rep.push_back((int32_t)c); }
      m->cls[c] = it->second;
    }
  } else {
    m->cls_rep.resize(nc);
    for (int64_t c = 0; c < nc; ++c) { m->cls[c] = (int32_t)c; m->cls_rep[c] = (int32_t)c; }
  }
  m->n_classes = (int64_t)m->cls_rep.size();
  if (m->ctx) {
    m->d_cls.upload(m->cls, m->ctx->stream); m->d_cls_rep.upload(m->cls_rep, m->ctx->stream);
    GG_CUDA(cudaStreamSynchronize(m->ctx->stream));
  }
  m->classes_built = true;
}

// Builtin numbering (our restatement of Gridap's) on the host arrays of `m`
void space_builtin_host(gg_mesh* m, int kind, int order, int constraint, gg_space* s) {
  s->mesh = m; s->kind = kind; s->order = order; s->constraint = constraint;
  int64_t nc = m->n_cells;
  if (m->dim == 3) {
    // hexahedral Q_p: vertices, edges, faces, interior (one DOF per entity for p = 2)
    if (kind == GG_SPACE_DG) {
      GG_REQUIRE(order >= 0 && order <= 2 && constraint == GG_CONSTRAINT_NONE, GG_ERR_UNSUPPORTED, "hex DG order must be in 0..2");
      int mloc = (order + 1) * (order + 1) * (order + 1);
      s->ndofs_loc = mloc; s->n_free = nc * mloc;
      s->cell_dofs.resize(nc * mloc);
      std::iota(s->cell_dofs.begin(), s->cell_dofs.end(), 0);
      return;
    }
    if (kind == GG_SPACE_RT) {
      // hexahedral RT_k (test/Tutorials/LinearBoussinesq.jl:55-57): (k+1)^2 DOFs per face, 3 k (k+1)^2 per cell; the slave
      // (second) cell of a face sees them negated; GG_CONSTRAINT_NOFLUX drops the DOFs of the boundary faces -- on the closed
      // shell these are the bottom and top faces (dirichlet_tags = ["bottom_boundary", "top_boundary"]) -- and numbers the rest
      // in the order of the face ids
      GG_REQUIRE(order >= 0 && order <= 1, GG_ERR_UNSUPPORTED, "hexahedral RT order must be 0 or 1");
      GG_REQUIRE(constraint == GG_CONSTRAINT_NONE || constraint == GG_CONSTRAINT_NOFLUX, GG_ERR_INVALID, "unknown constraint for an RT space");
      ensure_hex_topology(m);
      const int nf = (order + 1) * (order + 1), ni = 3 * order * (order + 1) * (order + 1), mloc = 6 * nf + ni;
      std::vector<int32_t> first(m->n_faces, -1), count(m->n_faces, 0), fid(m->n_faces, -1);
      for (int64_t c = 0; c < nc; ++c)
        for (int f = 0; f < 6; ++f) {
          const int32_t g = m->cell_faces6[c * 6 + f];
          if (first[g] < 0) first[g] = (int32_t)c;
          count[g]++;
        }
      int64_t nfree_faces = 0;
      for (int64_t g = 0; g < m->n_faces; ++g)
        if (constraint == GG_CONSTRAINT_NONE || count[g] > 1) fid[g] = (int32_t)nfree_faces++;
      s->ndofs_loc = mloc;
      s->n_free = nfree_faces * nf + nc * ni;
      s->n_dirichlet = (m->n_faces - nfree_faces) * nf;
      GG_REQUIRE(s->n_free < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "too many DOFs");
      s->cell_dofs.resize(nc * mloc);
      s->cell_signs.assign(nc * mloc, 1);
      s->blocks.push_back({0, nfree_faces, nf});
      if (ni > 0) s->blocks.push_back({nfree_faces * nf, nc, ni});
      for (int64_t c = 0; c < nc; ++c) {
        int k = 0;
        for (int f = 0; f < 6; ++f) {
          const int32_t g = m->cell_faces6[c * 6 + f];
          for (int t = 0; t < nf; ++t) {
            s->cell_dofs[c * mloc + k] = fid[g] >= 0 ? fid[g] * nf + t : -1;
            s->cell_signs[c * mloc + k] = first[g] == (int32_t)c ? 1 : -1;
            ++k;
          }
        }
        for (int t = 0; t < ni; ++t) s->cell_dofs[c * mloc + k++] = (int32_t)(nfree_faces * nf + c * ni + t);
      }
      return;
    }
    GG_REQUIRE(kind == GG_SPACE_CG && (order == 1 || order == 2), GG_ERR_UNSUPPORTED, "on the 3-D shell CG Q1/Q2, DG and RT0/RT1 spaces are implemented");
    int mloc = (order + 1) * (order + 1) * (order + 1);
    s->ndofs_loc = mloc;
    s->cell_dofs.resize(nc * mloc);
    int64_t nd = m->n_nodes;
    if (order == 2) { ensure_hex_topology(m); nd += m->n_edges + m->n_faces + nc; }
    GG_REQUIRE(nd < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "too many DOFs");
    for (int64_t c = 0; c < nc; ++c) {
      int32_t* o = &s->cell_dofs[c * mloc];
      int k = 0;
      for (int v = 0; v < 8; ++v) o[k++] = m->cell_nodes8[c * 8 + v];
      if (order == 2) {
        for (int e = 0; e < 12; ++e) o[k++] = (int32_t)(m->n_nodes + m->cell_edges12[c * 12 + e]);
        for (int f = 0; f < 6; ++f) o[k++] = (int32_t)(m->n_nodes + m->n_edges + m->cell_faces6[c * 6 + f]);
        o[k++] = (int32_t)(m->n_nodes + m->n_edges + m->n_faces + c);
      }
    }
    if (constraint == GG_CONSTRAINT_ZEROMEAN) {
      for (auto& d : s->cell_dofs)
        if (d == nd - 1) d = -1;
      s->n_free = nd - 1; s->n_dirichlet = 1;
    } else {
      GG_REQUIRE(constraint == GG_CONSTRAINT_NONE, GG_ERR_INVALID, "unknown constraint");
      s->n_free = nd;
    }
    return;
  }
  if (kind == GG_SPACE_CG) {
    GG_REQUIRE(order >= 1 && order <= 4, GG_ERR_UNSUPPORTED, "CG order must be in 1..4");
    ensure_topology(m);
    EdgeTopo T;
    T.cell_edges = m->cell_edges; T.flip = m->cell_edge_flip; T.n_edges = m->n_edges;
    int64_t nd;
    cg_numbering(nc, m->cell_nodes.data(), m->n_nodes, T, order, s->cell_dofs, nd);
    s->ndofs_loc = (order + 1) * (order + 1);
    if (constraint == GG_CONSTRAINT_ZEROMEAN) {
      for (auto& d : s->cell_dofs)
        if (d == nd - 1) d = -1;
      s->n_free = nd - 1; s->n_dirichlet = 1;
    } else {
      GG_REQUIRE(constraint == GG_CONSTRAINT_NONE, GG_ERR_INVALID, "unknown constraint");
      s->n_free = nd;
    }
    s->blocks.push_back({0, m->n_nodes, 1});
    if (order > 1) {
      s->blocks.push_back({m->n_nodes, m->n_edges, order - 1});
      s->blocks.push_back({m->n_nodes + m->n_edges * (order - 1), nc, (order - 1) * (order - 1)});
    }
  } else if (kind == GG_SPACE_CGV) {
    // vector-valued H1 Lagrangian (ReferenceFE(lagrangian, VectorValue{2,Float64}, p), conformity=:H1) with the panel-interface
    // change of basis of src/FESpaces/GradConformingFESpaces.jl:1-113.  Local DOF = node + n_nodes * component; global ids
    // entity by entity (scalar Q_p order), component-major inside an entity (App. B hypothesis, unpinned).
    GG_REQUIRE(order >= 1 && order <= 2, GG_ERR_UNSUPPORTED, "vector H1 order must be 1 or 2");
    GG_REQUIRE(constraint == GG_CONSTRAINT_NONE, GG_ERR_UNSUPPORTED, "constraints on vector H1 spaces are not implemented");
    ensure_topology(m);
    EdgeTopo T;
    T.cell_edges = m->cell_edges; T.flip = m->cell_edge_flip; T.n_edges = m->n_edges;
    std::vector<int32_t> sid;
    int64_t nds;
    cg_numbering(nc, m->cell_nodes.data(), m->n_nodes, T, order, sid, nds);
    const int nn = (order + 1) * (order + 1), ne1 = order - 1, ni = ne1 * ne1;
    const int64_t Nv = m->n_nodes, Ne = m->n_edges;
    GG_REQUIRE(2 * nds < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "too many DOFs");
    s->ndofs_loc = 2 * nn; s->n_free = 2 * nds;
    s->cell_dofs.resize((size_t)nc * 2 * nn);
    for (int64_t c = 0; c < nc; ++c)
      for (int a = 0; a < nn; ++a) {
        const int64_t d = sid[c * nn + a];
        for (int comp = 0; comp < 2; ++comp) {
          int64_t g;
          if (d < Nv) g = 2 * d + comp;
          else if (d < Nv + Ne * ne1) { const int64_t e = (d - Nv) / ne1, k = (d - Nv) % ne1; g = 2 * Nv + e * 2 * ne1 + comp * ne1 + k; }
          else { const int64_t cc = (d - Nv - Ne * ne1) / ni, k = (d - Nv - Ne * ne1) % ni; g = 2 * Nv + 2 * Ne * ne1 + cc * 2 * ni + comp * ni + k; }
          s->cell_dofs[(size_t)c * 2 * nn + a + nn * comp] = (int32_t)g;
        }
      }
    // master (cell, local node) of every (cell, node) whose owning vertex / edge has its master cell -- the surrounding cell
    // with the largest id -- on another panel; -1 elsewhere
    std::vector<int32_t> vmaster(Nv, -1), emaster(Ne, -1);
    for (int64_t c = 0; c < nc; ++c)
      for (int k = 0; k < 4; ++k) {
        GG_REQUIRE(!m->cell_edge_flip[c * 4 + k], GG_ERR_UNSUPPORTED, "the change of basis needs shared edges with the same vertex order in both cells");
        vmaster[m->cell_nodes[c * 4 + k]] = std::max(vmaster[m->cell_nodes[c * 4 + k]], (int32_t)c);
        emaster[m->cell_edges[c * 4 + k]] = std::max(emaster[m->cell_edges[c * 4 + k]], (int32_t)c);
      }
    s->cob_master_cell.assign((size_t)nc * nn, -1);
    s->cob_master_node.assign((size_t)nc * nn, 0);
    s->cob_master_panel.assign((size_t)nc * nn, -1);
    s->cob_master_point.assign((size_t)nc * nn * 2, 0.0);
    std::vector<int32_t> l2x_v = quad_local_to_lex(order);
    for (int64_t c = 0; c < nc; ++c)
      for (int a = 0; a < 4 + 4 * ne1; ++a) {
        const bool vert = a < 4;
        const int f = vert ? a : (a - 4) / ne1, k = vert ? 0 : (a - 4) % ne1;
        const int32_t ent = vert ? m->cell_nodes[c * 4 + f] : m->cell_edges[c * 4 + f];
        const int32_t mc = vert ? vmaster[ent] : emaster[ent];
        if (m->cell_to_chart[mc] == m->cell_to_chart[c]) continue;
        int fm = -1;
        for (int j = 0; j < 4; ++j) if ((vert ? m->cell_nodes[mc * 4 + j] : m->cell_edges[mc * 4 + j]) == ent) fm = j;
        GG_REQUIRE(fm >= 0, GG_ERR_INVALID, "inconsistent topology");
        s->cob_master_cell[c * nn + a] = mc;
        const int am = vert ? fm : 4 + fm * ne1 + k;
        s->cob_master_node[c * nn + a] = (int8_t)am;
        // chart point of the node in the master cell (axis-aligned chart box, Gridap vertex order BL, BR, TL, TR)
        const double* q = &m->chart_coords[(size_t)mc * 8];
        const double rx = (double)(l2x_v[am] % (order + 1)) / order, ry = (double)(l2x_v[am] / (order + 1)) / order;
        s->cob_master_panel[c * nn + a] = m->cell_to_chart[mc];
        s->cob_master_point[(c * nn + a) * 2] = q[0] + rx * (q[2] - q[0]);
        s->cob_master_point[(c * nn + a) * 2 + 1] = q[1] + ry * (q[5] - q[1]);
      }
    s->blocks.push_back({0, Nv, 2});
    if (order > 1) {
      s->blocks.push_back({2 * Nv, Ne, 2 * ne1});
      s->blocks.push_back({2 * Nv + 2 * Ne * ne1, nc, 2 * ni});
    }
  } else if (kind == GG_SPACE_DG) {
    GG_REQUIRE(order >= 0 && order <= 4, GG_ERR_UNSUPPORTED, "DG order must be in 0..4");
    int mloc = (order + 1) * (order + 1);
    s->ndofs_loc = mloc;
    s->n_free = nc * mloc;
    GG_REQUIRE(s->n_free < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "too many DOFs");
    s->cell_dofs.resize(nc * mloc);
    std::iota(s->cell_dofs.begin(), s->cell_dofs.end(), 0);
    if (constraint == GG_CONSTRAINT_ZEROMEAN && nc > 0) {
      // the last DOF is fixed (test/Laplacian/HodgeLaplacian_scalar.jl:44: pressure space with constraint=:zeromean)
      s->cell_dofs.back() = -1;
      s->n_free -= 1; s->n_dirichlet = 1;
      if (nc > 1) s->blocks.push_back({0, nc - 1, mloc});
      if (mloc > 1) s->blocks.push_back({(nc - 1) * mloc, 1, mloc - 1});
    } else {
      GG_REQUIRE(constraint == GG_CONSTRAINT_NONE, GG_ERR_INVALID, "unknown constraint");
      s->blocks.push_back({0, nc, mloc});
    }
  } else if (kind == GG_SPACE_RT) {
    GG_REQUIRE(order >= 0 && order <= 3, GG_ERR_UNSUPPORTED, "RT order must be in 0..3");
    GG_REQUIRE(constraint == GG_CONSTRAINT_NONE, GG_ERR_UNSUPPORTED, "constraints on RT spaces are not implemented");
    ensure_topology(m);
    int nf = order + 1, ni = 2 * order * (order + 1);
    int mloc = 4 * nf + ni;
    s->ndofs_loc = mloc;
    s->n_free = m->n_edges * nf + nc * ni;
    GG_REQUIRE(s->n_free < (int64_t)INT32_MAX, GG_ERR_UNSUPPORTED, "too many DOFs");
    s->cell_dofs.resize(nc * mloc);
    s->cell_signs.assign(nc * mloc, 1);
    s->blocks.push_back({0, m->n_edges, nf});
    if (ni > 0) s->blocks.push_back({m->n_edges * nf, nc, ni});
    for (int64_t c = 0; c < nc; ++c) {
      int k = 0;
      for (int e = 0; e < 4; ++e) {
        int32_t ed = m->cell_edges[c * 4 + e];
        GG_REQUIRE(!m->cell_edge_flip[c * 4 + e], GG_ERR_UNSUPPORTED,
                   "RT spaces need shared edges with the same vertex order in both cells");
        bool slave = m->edge_cells[ed * 2 + 1] == (int32_t)c;
        for (int t = 0; t < nf; ++t) {
          s->cell_dofs[c * mloc + k] = ed * nf + t;
          s->cell_signs[c * mloc + k] = slave ? -1 : 1;
          ++k;
        }
      }
      for (int t = 0; t < ni; ++t) s->cell_dofs[c * mloc + k++] = (int32_t)(m->n_edges * nf + c * ni + t);
    }
  } else {
    throw Error(GG_ERR_INVALID, "unknown space kind");
  }
}
}  // namespace gg

extern "C" {

int gg_space_builtin(gg_mesh* m, int kind, int order, int constraint, gg_space** out) {
  GG_API_BEGIN
  GG_REQUIRE(m && out, GG_ERR_INVALID, "null argument");
  if (m->ctx) GG_CUDA(cudaSetDevice(m->ctx->device));
  std::unique_ptr<gg_space> s(new gg_space);
  if (m->plan) {
    // rank-local sub-mesh of a partitioned model: the global builtin numbering restricted to the local cells (gg_dist.cu)
    GG_REQUIRE(constraint == GG_CONSTRAINT_NONE, GG_ERR_UNSUPPORTED, "constraints on partitioned spaces are not implemented");
    space_from_plan(m, kind, order, s.get());
  } else {
    space_builtin_host(m, kind, order, constraint, s.get());
  }
  finish_space(s.get());
  if (m->plan && m->ctx) space_upload_dist(s.get());
  *out = s.release();
  GG_API_END
}

int gg_space_from_arrays(gg_mesh* m, int kind, int order, int64_t n_free, const int32_t* cell_dof_ids,
                         const int8_t* cell_dof_signs, gg_space** out) {
  GG_API_BEGIN
  GG_REQUIRE(m && out && cell_dof_ids, GG_ERR_INVALID, "null argument");
  GG_REQUIRE(n_free >= 0 && n_free < (int64_t)INT32_MAX, GG_ERR_INVALID, "bad n_free");
  GG_CUDA(cudaSetDevice(m->ctx->device));
  std::unique_ptr<gg_space> s(new gg_space);
  s->mesh = m; s->kind = kind; s->order = order; s->constraint = 0;
  if (kind == GG_SPACE_CG || kind == GG_SPACE_DG) {
    GG_REQUIRE(order >= 0 && order <= 4, GG_ERR_UNSUPPORTED, "Lagrangian order must be in 0..4");
    s->ndofs_loc = (order + 1) * (order + 1);
  } else if (kind == GG_SPACE_RT) {
    GG_REQUIRE(order >= 0 && order <= 3, GG_ERR_UNSUPPORTED, "RT order must be in 0..3");
    s->ndofs_loc = 2 * (order + 1) * (order + 2);
  } else {
    throw Error(GG_ERR_INVALID, "unknown space kind");
  }
  size_t tot = (size_t)m->n_cells * s->ndofs_loc;
  s->cell_dofs.resize(tot);
  int64_t ndir = 0;
  for (size_t i = 0; i < tot; ++i) {
    int32_t d = cell_dof_ids[i];
    GG_REQUIRE(d != 0 && d <= n_free, GG_ERR_INVALID, "cell_dof_ids must be 1-based (negative = Dirichlet) and <= n_free");
    if (d < 0) { s->cell_dofs[i] = -1; ndir = std::max<int64_t>(ndir, -d); }
    else s->cell_dofs[i] = d - 1;
  }
  s->n_free = n_free; s->n_dirichlet = ndir;
  if (cell_dof_signs) {
    s->cell_signs.assign(cell_dof_signs, cell_dof_signs + tot);
  } else if (kind == GG_SPACE_RT) {
    s->cell_signs.assign(tot, 1);
  }
  finish_space(s.get());
  *out = s.release();
  GG_API_END
}

int gg_space_info(gg_space* s, int64_t* n_free, int32_t* ndofs_loc, int64_t* n_dirichlet) {
  GG_API_BEGIN
  GG_REQUIRE(s, GG_ERR_INVALID, "null space");
  if (n_free) *n_free = s->n_free;
  if (ndofs_loc) *ndofs_loc = s->ndofs_loc;
  if (n_dirichlet) *n_dirichlet = s->n_dirichlet;
  GG_API_END
}

int gg_space_get_cell_dofs(gg_space* s, int32_t* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && out, GG_ERR_INVALID, "null argument");
  for (size_t i = 0; i < s->cell_dofs.size(); ++i) out[i] = s->cell_dofs[i] < 0 ? -1 : s->cell_dofs[i] + 1;
  GG_API_END
}

int gg_space_get_cell_signs(gg_space* s, int8_t* out) {
  GG_API_BEGIN
  GG_REQUIRE(s && out, GG_ERR_INVALID, "null argument");
  if (s->cell_signs.empty()) std::fill(out, out + s->cell_dofs.size(), (int8_t)1);
  else std::copy(s->cell_signs.begin(), s->cell_signs.end(), out);
  GG_API_END
}

void gg_space_destroy(gg_space* s) { delete s; }

}  // extern "C"
