// 3-D shell driver pieces around the assembled operator (gg_hex_ops.cu): internal interface.
#pragma once
#include "gg_common.cuh"

namespace gg {
// points of the tensor Gauss rule with nq points per direction (order < 0) or the Lagrange nodes of `order` (nq ignored), first
// (gamma) index fastest: chart [n_cells][P][3] and ambient [n_cells][P][3] host arrays (either may be NULL)
void hex_points(gg_mesh* m, int order, int nq, double* chart_host, double* xyz_host);
void hex_source_vectors(gg_space* s, int nq, const double* fq, double* vbuf);                               // int f v sqrtg
void hex_boundary_flux_vectors(gg_space* s, int nq, const double* u, double dirichlet, double* vbuf);       // int_G (ginv grad u_h).n v sqrtg
// mode 0: int (u_h - f)^2 sqrtg, 1: int u_h (chart measure), 2: chart volume
double hex_cell_functional(gg_space* s, int mode, int nq, const double* u, double dirichlet, const double* fq);
}  // namespace gg
