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
  // shared-memory layout of k_lb_tile (sized by the plan): value accumulators, residual accumulators, rows staged
  int32_t n_acc, n_vacc, n_rows;
};

__device__ __forceinline__ void tile_cp_async8(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void tile_cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void tile_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory layout of k_lb_tile: double A_s[n_acc] value accumulators (incl. 32 dummies; the parked state values u_e
// [M][128] alias their head), double V_s[n_vacc] residual accumulators (incl. 32 dummies), int32 base_s[n_rows] / dof_s[n_rows]
// CSR bases and DOF ids of the tile's rows.  n_acc, n_vacc, n_rows are even.
inline size_t tile_smem_bytes(int M, int n_acc, int n_vacc, int n_rows) {
  const size_t acc = (size_t)(n_acc > M * TILE_CELLS ? n_acc : M * TILE_CELLS);
  return (acc + (size_t)n_vacc + (size_t)n_rows) * sizeof(double);
}

template <int M>
__host__ __device__ constexpr int tile_packed(int I, int J) {  // I <= J
  return I * M - (I * (I - 1)) / 2 + (J - I);
}

// MULTI: geometry-class mode, a CTA serves several tiles with the same chart boxes from one set of element matrices
template <int NB, int NQ, bool EXTRINSIC, bool MULTI>
__global__ void __launch_bounds__(TILE_CELLS, 3)
k_lb_tile(TileDev T, TileArgs a, CellGeo g) {
  constexpr int M = NB * NB, NS = NB * (NB + 1) / 2, NP = M * (M + 1) / 2;
  constexpr int ROW = 2 * NS + M;
  constexpr int ONN = 0, ODD = NS, OND = 2 * NS;
  extern __shared__ double sm[];
  double* A_s = sm;
  double* U_s = sm;                                  // parked u_e: consumed before the accumulators are zeroed
  double* V_s = sm + (a.n_acc > M * TILE_CELLS ? a.n_acc : M * TILE_CELLS);
  int32_t* base_s = (int32_t*)(V_s + a.n_vacc);
  int32_t* dof_s = base_s + a.n_rows;
  const int lc = threadIdx.x, lane = lc & 31, warp = lc >> 5;
  // without geometry classes supertile b is tile b
  const int i0 = MULTI ? T.inst_ptr[blockIdx.x] : (int)blockIdx.x, i1 = MULTI ? T.inst_ptr[blockIdx.x + 1] : i0 + 1;
  const int32_t tile0 = MULTI ? T.inst_tile[i0] : (int32_t)blockIdx.x;
  const int32_t crep = T.tile_cells[(int64_t)tile0 * TILE_CELLS + lc];

  // ---- prologue: the row bases the write-out will need start their way into shared memory now (cp.async); the state values of
  //      the thread's cell are gathered while the geometry loads are in flight and parked in shared memory ----
  auto prefetch_rows = [&](int32_t tile) {
    const int nrows = T.pats[T.tile_pat[tile]].nrows;
    const int32_t r0 = T.tile_row0[tile];
    for (int k = lc; k < nrows; k += TILE_CELLS) {
      tile_cp_async4(base_s + k, T.trow_base + r0 + k);
      if (a.vec_out) tile_cp_async4(dof_s + k, T.trow_dof + r0 + k);
    }
  };
  auto gather_state = [&](int32_t tile) {
    if (a.vec_out) {
      const int32_t* __restrict__ td = T.tc_dofs + (int64_t)tile * (M * TILE_CELLS) + lc;
      int32_t d[M];
#pragma unroll
      for (int I = 0; I < M; ++I) d[I] = td[I * TILE_CELLS];
#pragma unroll
      for (int I = 0; I < M; ++I) U_s[I * TILE_CELLS + lc] = d[I] >= 0 ? a.u[d[I]] : a.dirichlet;
    }
  };
  prefetch_rows(tile0);
  gather_state(tile0);
  const int32_t patid0 = T.tile_pat[tile0];
  const uint8_t bidx0 = T.bidx[(int64_t)patid0 * TILE_CELLS + lc];

  // ---- element matrix of the thread's cell (the instances of a supertile have the same chart boxes) ----
  double acc[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) acc[p] = 0.0;
  if (crep >= 0) {
    const int64_t tl = (int64_t)tile0 * TILE_CELLS + lc;
    const double HX = T.tc_hx[tl], HY = T.tc_hy[tl];
    const double rx = HY / HX, ry = HX / HY;
    const double* __restrict__ ta = g.tx + (int64_t)T.tc_ix[tl] * NQ;
    const double* __restrict__ tb = g.ty + (int64_t)T.tc_iy[tl] * NQ;
    int panel = 0;
    if (EXTRINSIC) panel = g.chart[crep];
    double b[NQ], sb[NQ];
#pragma unroll
    for (int j = 0; j < NQ; ++j) {
      b[j] = tb[j];
      sb[j] = fma(b[j], b[j], 1.0);
    }
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
          G11 = sb[q2] * rinv;
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


  for (int inst = i0; inst < i1; ++inst) {
    const int32_t tile = MULTI ? T.inst_tile[inst] : tile0;
    const int32_t patid = (!MULTI || inst == i0) ? patid0 : T.tile_pat[tile];
    const int32_t c = (!MULTI || inst == i0) ? crep : T.tile_cells[(int64_t)tile * TILE_CELLS + lc];
    const bool border = ((!MULTI || inst == i0) ? bidx0 : T.bidx[(int64_t)patid * TILE_CELLS + lc]) != 0xFF && c >= 0;
    const TilePat* __restrict__ P = T.pats + patid;
    const int nslots = P->nslots, nrows = P->nrows;
    if (MULTI && inst > i0) {
      __syncthreads();            // the previous instance has been written out
      prefetch_rows(tile);
      gather_state(tile);
    }
    // ---- element residual r_e = K_e u_e; border cells leave their element matrix / vector in the global border buffers ----
    double re[M], ue[M];
    if (a.vec_out && c >= 0) {
#pragma unroll
      for (int I = 0; I < M; ++I) ue[I] = U_s[I * TILE_CELLS + lc];
    }
    __syncthreads();              // everybody has its u_e: the area becomes the accumulators
    if (a.vec_out && c >= 0) {
#pragma unroll
      for (int I = 0; I < M; ++I) re[I] = 0.0;
#pragma unroll
      for (int I = 0; I < M; ++I)
#pragma unroll
        for (int J = I; J < M; ++J) {
          re[I] = fma(acc[tile_packed<M>(I, J)], ue[J], re[I]);
          if (J != I) re[J] = fma(acc[tile_packed<M>(I, J)], ue[I], re[J]);
        }
      if (border) {
        double* vb = T.vbuf_b + ((int64_t)tile * TILE_CELLS + lc) * (M + (M & 1));
#pragma unroll
        for (int I = 0; I < M; ++I) vb[c_lex2loc[I]] = re[I];
      }
    }
    if (border) {
      // cell-major: the NP entries of a border cell are contiguous, so the lanes that sum one row share their sectors
      double* eb = T.ebuf_b + ((int64_t)tile * TILE_CELLS + lc) * (NP + (NP & 1));
#pragma unroll
      for (int p = 0; p < NP; ++p) eb[c_plane[p]] = acc[p];
    }
    // ---- conflict-free accumulation in rounds.  Every (cell, entry) knows the shared-memory slot it feeds and its ROUND: its
    //      rank among the cells that feed the slot, in ascending cell order.  Round 0 stores, rounds 1..3 add, with a barrier in
    //      between: no atomics, no zeroing, and every value is summed in ascending cell order.  Only entries whose two DOFs lie on
    //      one edge of the cell can have a later round (tile_shared_pair), only a corner's diagonal entry a third or fourth ----
    constexpr int PW = (M * M + 1) / 2, VW = (M + 1) / 2, P_ = NB - 1;
    uint32_t pw[PW], vw[VW];
    {
      const uint32_t* __restrict__ pi = T.pair_idx + (size_t)patid * (PW * TILE_CELLS) + lc;
      const uint32_t* __restrict__ vi = T.vec_idx + (size_t)patid * (VW * TILE_CELLS) + lc;
#pragma unroll
      for (int q = 0; q < PW; ++q) pw[q] = __ldg(pi + q * TILE_CELLS);
      if (a.vec_out) {
#pragma unroll
        for (int q = 0; q < VW; ++q) vw[q] = __ldg(vi + q * TILE_CELLS);
      }
    }
    auto entry = [&](int e) { return (e & 1) ? (pw[e >> 1] >> 16) : (pw[e >> 1] & 0xFFFFu); };
    auto ventry = [&](int I) { return (I & 1) ? (vw[I >> 1] >> 16) : (vw[I >> 1] & 0xFFFFu); };
    if (c >= 0) {
#pragma unroll
      for (int I = 0; I < M; ++I)
#pragma unroll
        for (int J = 0; J < M; ++J) {
          const uint32_t w = entry(I * M + J);
          const double v = acc[I <= J ? tile_packed<M>(I, J) : tile_packed<M>(J, I)];
          if (!tile_shared_pair(P_, I, J) || (w >> 14) == 0u) A_s[w & 0x3FFFu] = v;
        }
      if (a.vec_out) {
#pragma unroll
        for (int I = 0; I < M; ++I) {
          const uint32_t w = ventry(I);
          if (!tile_shared_dof(P_, I) || (w >> 14) == 0u) V_s[w & 0x3FFFu] = re[I];
        }
      }
    }
#pragma unroll
    for (int round = 1; round < 4; ++round) {
      __syncthreads();
      if (c >= 0) {
#pragma unroll
        for (int I = 0; I < M; ++I)
#pragma unroll
          for (int J = 0; J < M; ++J) {
            if (!tile_shared_pair(P_, I, J)) continue;
            if (round > 1 && !(I == J && tile_corner_dof(P_, I))) continue;
            const uint32_t w = entry(I * M + J);
            if ((w >> 14) == (uint32_t)round) A_s[w & 0x3FFFu] += acc[I <= J ? tile_packed<M>(I, J) : tile_packed<M>(J, I)];
          }
        if (a.vec_out) {
#pragma unroll
          for (int I = 0; I < M; ++I) {
            if (!tile_shared_dof(P_, I)) continue;
            if (round > 1 && !tile_corner_dof(P_, I)) continue;
            const uint32_t w = ventry(I);
            if ((w >> 14) == (uint32_t)round) V_s[w & 0x3FFFu] += re[I];
          }
        }
      }
    }
    tile_cp_async_wait();
    __syncthreads();
    // ---- write-out: a warp takes floor(32 / len) rows at a time, one lane per CSR value (consecutive lanes -> consecutive
    //      addresses); the accumulators of a class are laid out row after row ----
    const int ncls = P->ncls;
    for (int q = 0; q < ncls; ++q) {
      const int len = P->cls[q].len, first = P->cls[q].first, count = P->cls[q].count, slot0 = P->cls[q].slot0;
      const int rpw = 32 / len;
      const int r = lane / len, jj = lane - r * len;
      if (r >= rpw) continue;
#pragma unroll 4
      for (int k = warp * rpw + r; k < count; k += 4 * rpw) {
        const double t = a.scale * A_s[slot0 + k * len + jj];
        double* dst = a.vals + base_s[first + k] + jj;
        *dst = a.add ? *dst + t : t;
      }
    }
    if (a.vec_out) {
      for (int k = lc; k < nrows; k += TILE_CELLS) {
        const double t = a.scale * V_s[k];
        const int32_t d = dof_s[k];
        a.vec_out[d] = a.add ? a.vec_out[d] + t : t;
      }
    }
  }
}

// Rows shared by several tiles (tile edges and corners): one thread per CSR value, summed from the border buffers the tile kernel
// has just written (mostly L2 hits) in ascending cell order.  The rows are sorted by length, so a thread finds its (row, column)
// by one division inside its length class.
template <int M>
__global__ void __launch_bounds__(256)
k_lb_tile_groups(TileDev T, TileArgs a, GroupClasses G) {
  constexpr int NP = M * (M + 1) / 2, U = 2;     // U values per thread, their load chains in flight together
  const uint32_t n_ent = (uint32_t)G.ent0[G.ncls];
  const uint32_t e0 = blockIdx.x * (blockDim.x * U) + threadIdx.x;
  const GroupRow* R[U];
  int jj[U];
  uint32_t s4[U];
  double v[U][4];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint32_t e = e0 + u * blockDim.x;
    R[u] = nullptr;
    if (e < n_ent) {
      // length class of the entry: compile-time indexed compares keep the class table in the parameter bank
      uint32_t len = (uint32_t)G.cls[0].len, first = (uint32_t)G.cls[0].first, start = 0;
#pragma unroll
      for (int q = 1; q < TILE_MAXCLS; ++q)
        if (q < G.ncls && e >= (uint32_t)G.ent0[q]) { len = (uint32_t)G.cls[q].len; first = (uint32_t)G.cls[q].first; start = (uint32_t)G.ent0[q]; }
      const uint32_t le = e - start, row = le / len;
      R[u] = T.grows + first + row;
      jj[u] = (int)(le - row * len);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (R[u]) s4[u] = __ldg((const uint32_t*)T.types[__ldg(&R[u]->type)].src[jj[u]]);
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      v[u][w] = 0.0;
      if (R[u]) {
        const uint32_t code = (s4[u] >> (8 * w)) & 255u;
        if (code != 255u)    // border buffer: [tile * 128 + local cell][plane]
          v[u][w] = __ldcg(T.ebuf_b + (int64_t)__ldg(&R[u]->col[code >> 6]) * (NP + (NP & 1)) + (int)(code & 63u));
      }
    }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!R[u]) continue;
    const double t = a.scale * (((v[u][0] + v[u][1]) + v[u][2]) + v[u][3]);
    double* dst = a.vals + __ldg(&R[u]->base) + jj[u];
    *dst = a.add ? *dst + t : t;
    if (jj[u] == 0 && a.vec_out) {
      const uint32_t s4v = __ldg((const uint32_t*)T.types[__ldg(&R[u]->type)].vsrc);
      double r[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t code = (s4v >> (8 * w)) & 255u;
        r[w] = 0.0;
        if (code != 255u) r[w] = __ldcg(T.vbuf_b + (int64_t)__ldg(&R[u]->col[(code >> 4) & 3u]) * (M + (M & 1)) + (int)(code & 15u));
      }
      const double tv = a.scale * (((r[0] + r[1]) + r[2]) + r[3]);
      const int32_t d = __ldg(&R[u]->dof);
      a.vec_out[d] = a.add ? a.vec_out[d] + tv : tv;
    }
  }
}

}  // namespace gg
