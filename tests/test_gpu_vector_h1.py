"""GPU parity: vector-valued H1 space with the panel-interface change of basis (src/FESpaces/GradConformingFESpaces.jl) and the
L2 projection driver of test/Projection/L2_projection_Lagrangian_vector.jl against oracle/vector_h1.py."""
import numpy as np
import pytest

from __graft_entry__ import load_package
from oracle import forms
from oracle import vector_h1 as vh
from oracle.mesh import CubedSphereMesh2D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg():
    return load_package()


def uX(X):
    return np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], axis=-1)


def wX(X):
    return np.stack([X[..., 2] * X[..., 1], -X[..., 0] ** 2, X[..., 0] + 0.3], axis=-1)


@pytest.mark.parametrize("n,p", [(2, 1), (3, 2), (4, 1)])
def test_numbering_change_of_basis_and_forms_match_oracle(gg, n, p):
    r = 1.4
    mesh = CubedSphereMesh2D(n, r)
    Vo = vh.VectorCGSpace(mesh, p)
    model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(r), n=n)
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, p), conformity="H1")
    assert V.num_free_dofs == Vo.num_free and V.ndofs_loc == Vo.ndofs_loc
    assert (V.get_cell_dof_ids() - 1 == Vo.cell_dofs).all()
    assert np.abs(gg.change_of_basis(V) - Vo.cob).max() < 1e-13
    degree = 4 * (p + 1)
    dΩ = gg.Measure(gg.Triangulation(model), degree)
    q = forms.CellQuadrature(mesh, degree)
    Ko = vh.mass_element_matrices(Vo, q)
    Ke = gg.element_matrices(gg.VectorMass(dΩ), V, V)
    assert np.abs(Ke - Ko).max() <= 1e-12 * np.abs(Ko).max()
    fq = vh.contravariant(mesh, q, wX)
    fg = gg.contravariant_components(dΩ, wX).get().reshape(fq.shape)
    assert np.abs(fg - fq).max() <= 1e-12 * np.abs(fq).max()
    bo = forms.assemble_vector(Vo, vh.source_element_vectors(Vo, q, fq))
    b = gg.assemble_vector(gg.VectorSource(dΩ, fq), V).get()
    assert np.abs(b - bo).max() <= 1e-12 * np.abs(bo).max()
    Ao = forms.assemble_matrix(Vo, Vo, Ko)
    A = gg.assemble_matrix(gg.VectorMass(dΩ), V, V).to_scipy()
    assert np.abs((A - Ao).data).max() <= 1e-12 * np.abs(Ao.data).max()
    # interpolation and the error norm
    U = gg.interpolate(wX, V).get()
    Uo = vh.interpolate(Vo, wX)
    assert np.abs(U - Uo).max() <= 1e-12 * np.abs(Uo).max()
    assert abs(gg.l2_error(U, V, dΩ, fq=fq) - vh.l2_error(Vo, q, Uo, fq)) <= 1e-10 * vh.l2_error(Vo, q, Uo, fq)


@pytest.mark.parametrize("p", [1, 2])
def test_l2_projection_driver_converges(gg, p):
    # test/Projection/L2ProjectionTests.jl:13-15,36: vecX = (-y, x, 0), H1, slope >= p; PCG on the device-assembled mass matrix
    errs = []
    levels = (1, 2, 3, 4)
    for lvl in levels:
        model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=2**lvl)
        V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, gg.VectorValue, p), conformity="H1")
        Ω = gg.Triangulation(model)
        degree = 4 * (p + 1)
        dΩ, dΩe = gg.Measure(Ω, degree), gg.Measure(Ω, 2 * degree)
        op = gg.AffineFEOperator(gg.VectorMass(dΩ), gg.VectorSource(dΩ, gg.contravariant_components(dΩ, uX)), V, V)
        ns = gg.CGSolver(rtol=1e-13, maxiter=5000).numerical_setup(op.get_matrix())
        uh = ns.solve(op.get_vector())
        assert ns.rel_res < 1e-12
        errs.append(gg.l2_error(uh, V, dΩe, fq=gg.contravariant_components(dΩe, uX)))
        if lvl == 2:
            e_ref = vh.l2_projection(CubedSphereMesh2D(4), p, uX)[0]
            assert abs(errs[-1] - e_ref) <= 1e-8 * e_ref
    slope = np.polyfit(np.log10([1 / 2.0**l for l in levels]), np.log10(errs), 1)[0]
    assert slope >= p
