// Vector-valued H1 space with the panel-interface change of basis (gg_vec.cu): internal interface.
#pragma once
#include "gg_common.cuh"

namespace gg {
void vec_space_finish(gg_space* s);   // change-of-basis blocks on the device
void vec_mass_element_matrices(gg_space* s, int nq, double* ebuf);   // full entry-major [i * m + j][cell]
void vec_laplacian_element_matrices(gg_space* s, int nq, double* ebuf);   // same layout
void vec_div_element_matrices(gg_space* test, gg_space* trial, int nq, double* ebuf);   // [I * m_trial + j][cell]
void vec_source_element_vectors(gg_space* s, int nq, const double* fq, double* vbuf);
void vec_interp_values(gg_space* s, const double* ref_dev, const double* v_dev, double* vals_dev);
double vec_h1_seminorm_sq(gg_space* s, const double* U, int nq);
double vec_error_sq(gg_space* s, const double* U, const double* fq, int nq);
}  // namespace gg
