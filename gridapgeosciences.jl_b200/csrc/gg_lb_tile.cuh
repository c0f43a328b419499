// Fused tile assembly of the Laplace-Beltrami Jacobian (+ residual): element matrices never leave the SM for rows that are
// complete inside a tile; see gg_tile.cuh for the scheme and gg_lb_fast.cuh for the element arithmetic (the same code path:
// one thread per cell, sum-factorised, FP64 registers).  Included by gg_forms.cu (shares its __constant__ tables).
#pragma once
#include "gg_lb_fast.cuh"
#include "gg_tile.cuh"

namespace gg {

struct TileArgs {
  const int32_t* cell_dofs;
  const double* u;        // trial DOF values for the fused residual (nullptr: matrix only)
  double dirichlet, scale;
  int add;
  double* vals;           // CSR values
  double* vec_out;        // residual (nullptr: matrix only)
};

__device__ __forceinline__ double tile_sum4(const double* __restrict__ s, ushort4 q) {
  double v = s[q.x];
  if (q.y != TILE_NONE) v += s[q.y];
  if (q.z != TILE_NONE) v += s[q.z];
  if (q.w != TILE_NONE) v += s[q.w];
  return v;
}

template <int NB, int NQ, bool EXTRINSIC>
__global__ void __launch_bounds__(TILE_CELLS, 3)
k_lb_tile(TileDev T, TileArgs a, CellGeo g) {
  constexpr int M = NB * NB, NS = NB * (NB + 1) / 2, NP = M * (M + 1) / 2;
  constexpr int ROW = 2 * NS + M;
  constexpr int ONN = 0, ODD = NS, OND = 2 * NS;
  extern __shared__ double sm[];
  double* Ks = sm;                                  // [NP][TILE_CELLS] packed element matrices of the tile
  double* Rs = Ks + NP * TILE_CELLS;                // [M][TILE_CELLS] element residuals (Gridap local order)
  int32_t* rowbase = (int32_t*)(Rs + M * TILE_CELLS);   // [TILE_ROWCAP] first CSR slot of the rows being written
  __shared__ int32_t s_ready[32];
  __shared__ int32_t s_geo[4], s_gvo[4], s_gnb[4];   // border-buffer offsets / border counts of the tiles of the group at hand
  const int lc = threadIdx.x;
  const int i0 = T.inst_ptr[blockIdx.x], i1 = T.inst_ptr[blockIdx.x + 1];
  const int32_t crep = T.tile_cells[(int64_t)T.inst_tile[i0] * TILE_CELLS + lc];

  // ---- element matrix of the thread's cell (the instances of a supertile have the same chart boxes) ----
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  if (crep >= 0) {
    const int64_t c = crep;
    const double HX = g.hx[c], HY = g.hy[c];
    const double rx = HY / HX, ry = HX / HY;
    const double* __restrict__ ta = g.tx + (int64_t)g.ix[c] * NQ;
    const double* __restrict__ tb = g.ty + (int64_t)g.iy[c] * NQ;
    int panel = 0;
    if (EXTRINSIC) panel = g.chart[c];
    // 1 + b^2 is recomputed at every point (one DFMA) instead of being kept for the whole cell: the 2 NQ registers it would
    // hold are what the scatter bookkeeping of this kernel needs to stay clear of spills in the hot loop
    double b[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) b[j] = tb[j];
#pragma unroll 1
    for (int q1 = 0; q1 < NQ; ++q1) {
      const double av = ta[q1];
      const double sa = fma(av, av, 1.0);
      double T11[NS], T22[NS], T12[M];
#pragma unroll
      for (int k = 0; k < NS; ++k) { T11[k] = 0.0; T22[k] = 0.0; }
#pragma unroll
      for (int k = 0; k < M; ++k) T12[k] = 0.0;
#pragma unroll
      for (int q2 = 0; q2 < NQ; ++q2) {
        double G11, G12, G22;
        if (!EXTRINSIC) {
          const double rinv = fast_rsqrt(fma(b[q2], b[q2], sa));
          const double ar = av * rinv;
          G11 = fma(b[q2], b[q2], 1.0) * rinv;
          G12 = ar * b[q2];
          G22 = sa * rinv;
        } else {
          Metric2 mt = extrinsic_metric_terms(panel, av, b[q2], g.radius);
          G11 = mt.i11 * mt.sqrtg; G12 = mt.i12 * mt.sqrtg; G22 = mt.i22 * mt.sqrtg;
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) T11[k] = fma(G11, c_lb[q2 * ROW + ONN + k], T11[k]);
#pragma unroll
        for (int k = 0; k < NS; ++k) T22[k] = fma(G22, c_lb[q2 * ROW + ODD + k], T22[k]);
#pragma unroll
        for (int k = 0; k < M; ++k) T12[k] = fma(G12, c_lb[q2 * ROW + OND + k], T12[k]);
      }
#pragma unroll
      for (int k = 0; k < NS; ++k) { T11[k] *= rx; T22[k] *= ry; }
      const double* __restrict__ row = c_lb + q1 * ROW;
      int p = 0;
#pragma unroll
      for (int i2 = 0; i2 < NB; ++i2)
#pragma unroll
        for (int i1 = 0; i1 < NB; ++i1)
#pragma unroll
          for (int j2 = i2; j2 < NB; ++j2)
#pragma unroll
            for (int j1 = 0; j1 < NB; ++j1) {
              if (j2 == i2 && j1 < i1) continue;
              const int s1 = i1 <= j1 ? sym_index<NB>(i1, j1) : sym_index<NB>(j1, i1);
              const int s2 = sym_index<NB>(i2, j2);
              double v = acc[p];
              v = fma(row[ODD + s1], T11[s2], v);
              v = fma(row[OND + j1 * NB + i1], T12[i2 * NB + j2], v);
              v = fma(row[OND + i1 * NB + j1], T12[j2 * NB + i2], v);
              v = fma(row[ONN + s1], T22[s2], v);
              acc[p] = v;
              ++p;
            }
    }
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) Ks[c_plane[p] * TILE_CELLS + lc] = acc[p];

  for (int inst = i0; inst < i1; ++inst) {
    const int32_t tile = T.inst_tile[inst];
    const int32_t patid = T.tile_pat[tile];
    const TilePat P = T.pats[patid];
    const int32_t c = T.tile_cells[(int64_t)tile * TILE_CELLS + lc];
    const int bi = T.bidx[(int64_t)patid * TILE_CELLS + lc];
    // the thread's own element matrix, back from shared memory (so that it is not live in registers across the scatter phases)
    double kv[NP];
    if (c >= 0 && (a.vec_out || bi != 0xFF)) {
#pragma unroll
      for (int p = 0; p < NP; ++p) kv[p] = Ks[c_plane[p] * TILE_CELLS + lc];
    }
    // ---- element residual r_e = K_e u_e of this instance's cell ----
    if (a.vec_out && c >= 0) {
      double ue[M], re[M];
#pragma unroll
      for (int I = 0; I < M; ++I) {
        const int32_t d = a.cell_dofs[(int64_t)c * M + c_lex2loc[I]];
        ue[I] = d >= 0 ? a.u[d] : a.dirichlet;
        re[I] = 0.0;
      }
      int p = 0;
#pragma unroll
      for (int I = 0; I < M; ++I)
#pragma unroll
        for (int J = I; J < M; ++J) {
          re[I] = fma(kv[p], ue[J], re[I]);
          if (J != I) re[J] = fma(kv[p], ue[I], re[J]);
          ++p;
        }
#pragma unroll
      for (int I = 0; I < M; ++I) Rs[c_lex2loc[I] * TILE_CELLS + lc] = re[I];
      if (bi != 0xFF) {
        double* vb = T.vbuf_b + T.tile_voff[tile];
#pragma unroll
        for (int I = 0; I < M; ++I) __stcg(vb + c_lex2loc[I] * P.nborder + bi, re[I]);
      }
    }
    // ---- border cells: element matrix to the global border buffer, for the rows shared with other tiles ----
    if (bi != 0xFF && c >= 0) {
      double* eb = T.ebuf_b + T.tile_eoff[tile];
#pragma unroll
      for (int p = 0; p < NP; ++p) __stcg(eb + c_plane[p] * P.nborder + bi, kv[p]);
    }
    __threadfence();
    __syncthreads();
    // ---- complete rows: first CSR slot of every row into shared memory; residual entries of the rows ----
    {
      const int32_t* __restrict__ tb = T.trow_base + T.tile_row0[tile];
      const int32_t* __restrict__ td = T.trow_dof + T.tile_row0[tile];
      const ushort4* __restrict__ vs = T.rowvsrc + P.row0;
      for (int k0 = lc; k0 < P.nrows; k0 += 4 * TILE_CELLS) {
        int32_t rb[4], d[4];
        ushort4 s4[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int k = k0 + w * TILE_CELLS;
          if (k < P.nrows) {
            rb[w] = tb[k];
            if (a.vec_out) { d[w] = td[k]; s4[w] = vs[k]; }
          }
        }
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          const int k = k0 + w * TILE_CELLS;
          if (k < P.nrows) {
            rowbase[k] = rb[w];
            if (a.vec_out) {
              const double v = a.scale * tile_sum4(Rs, s4[w]);
              a.vec_out[d[w]] = a.add ? a.vec_out[d[w]] + v : v;
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- complete rows: every CSR value summed from shared memory in ascending cell order, written once ----
    {
      const uint32_t* __restrict__ em = T.emeta + P.ent0;
      const ushort4* __restrict__ es = T.esrc + P.ent0;
      constexpr int U = 8;
      for (int e0 = lc; e0 < P.nent; e0 += U * TILE_CELLS) {
        uint32_t mt[U];
        ushort4 s4[U];
#pragma unroll
        for (int w = 0; w < U; ++w) {
          const int e = e0 + w * TILE_CELLS;
          if (e < P.nent) { mt[w] = em[e]; s4[w] = es[e]; }
        }
#pragma unroll
        for (int w = 0; w < U; ++w) {
          const int e = e0 + w * TILE_CELLS;
          if (e < P.nent) {
            const double v = a.scale * tile_sum4(Ks, s4[w]);
            double* dst = a.vals + rowbase[mt[w] >> 8] + (mt[w] & 255u);
            *dst = a.add ? *dst + v : v;
          }
        }
      }
    }
    // ---- groups of rows shared with other tiles: the last tile to arrive sums them from the border buffers ----
    const int gb = T.tg_ptr[tile], ge = T.tg_ptr[tile + 1];
    for (int g0 = gb; g0 < ge; g0 += 32) {
      __syncthreads();
      if (lc < 32) {
        int ready = -1;
        if (g0 + lc < ge) {
          const int32_t gidx = T.tg_idx[g0 + lc];
          const int old = atomicAdd(&T.gcount[gidx], 1);
          if (old == T.groups[gidx].need - 1) { ready = gidx; T.gcount[gidx] = 0; }
        }
        s_ready[lc] = ready;
      }
      __syncthreads();
      for (int q = 0; q < 32 && g0 + q < ge; ++q) {
        const int32_t gidx = s_ready[q];
        if (gidx < 0) continue;                 // uniform across the CTA
        __threadfence();
        const GroupInst* __restrict__ Gp = T.groups + gidx;
        const int32_t g_row0 = Gp->row0;
        const GroupPat GP = T.gpats[Gp->pat];
        __syncthreads();                        // rowbase and the group slots are free again
        if (lc < 4) { s_geo[lc] = Gp->eoff[lc]; s_gvo[lc] = Gp->voff[lc]; s_gnb[lc] = Gp->nb[lc]; }
        __syncthreads();
        {
          const int32_t* __restrict__ tb = T.grow_base + g_row0;
          const int32_t* __restrict__ td = T.grow_dof + g_row0;
          const ushort4* __restrict__ vs = T.g_rowvsrc + GP.row0;
          for (int k = lc; k < GP.nrows; k += TILE_CELLS) {
            rowbase[k] = tb[k];
            if (a.vec_out) {
              const ushort4 s4 = vs[k];
              const int32_t d = td[k];
              auto ld = [&](uint16_t code) {
                const int ts = code >> 11;
                return __ldcg(T.vbuf_b + s_gvo[ts] + (code & 15) * s_gnb[ts] + ((code >> 4) & 127));
              };
              const double v0 = ld(s4.x);
              const double v1 = s4.y != TILE_NONE ? ld(s4.y) : 0.0;
              const double v2 = s4.z != TILE_NONE ? ld(s4.z) : 0.0;
              const double v3 = s4.w != TILE_NONE ? ld(s4.w) : 0.0;
              double v = v0;
              if (s4.y != TILE_NONE) v += v1;
              if (s4.z != TILE_NONE) v += v2;
              if (s4.w != TILE_NONE) v += v3;
              v *= a.scale;
              a.vec_out[d] = a.add ? a.vec_out[d] + v : v;
            }
          }
        }
        __syncthreads();
        {
          const uint32_t* __restrict__ em = T.g_emeta + GP.ent0;
          const ushort4* __restrict__ es = T.g_esrc + GP.ent0;
          constexpr int U = 4;
          for (int e0 = lc; e0 < GP.nent; e0 += U * TILE_CELLS) {
            uint32_t mt[U];
            ushort4 s4[U];
            double v[U][4];
#pragma unroll
            for (int w = 0; w < U; ++w) {
              const int e = e0 + w * TILE_CELLS;
              if (e < GP.nent) { mt[w] = em[e]; s4[w] = es[e]; }
            }
            auto ld = [&](uint16_t code) {
              const int ts = code >> 13;
              return __ldcg(T.ebuf_b + s_geo[ts] + (code & 63) * s_gnb[ts] + ((code >> 6) & 127));
            };
#pragma unroll
            for (int w = 0; w < U; ++w) {
              const int e = e0 + w * TILE_CELLS;
              if (e < GP.nent) {
                v[w][0] = ld(s4[w].x);
                v[w][1] = s4[w].y != TILE_NONE ? ld(s4[w].y) : 0.0;
                v[w][2] = s4[w].z != TILE_NONE ? ld(s4[w].z) : 0.0;
                v[w][3] = s4[w].w != TILE_NONE ? ld(s4[w].w) : 0.0;
              }
            }
#pragma unroll
            for (int w = 0; w < U; ++w) {
              const int e = e0 + w * TILE_CELLS;
              if (e < GP.nent) {
                double t = v[w][0];                       // same order of additions as the sources are listed (ascending cell)
                if (s4[w].y != TILE_NONE) t += v[w][1];
                if (s4[w].z != TILE_NONE) t += v[w][2];
                if (s4[w].w != TILE_NONE) t += v[w][3];
                t *= a.scale;
                double* dst = a.vals + rowbase[mt[w] >> 8] + (mt[w] & 255u);
                *dst = a.add ? *dst + t : t;
              }
            }
          }
        }
      }
    }
    __syncthreads();   // Rs / rowbase are reused by the next instance
  }
}

}  // namespace gg
