#!/usr/bin/env python
"""Smallest end-to-end run of every kernel family (for compute-sanitizer): assembly, shell, wave, shallow water, advection."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

gg = load_package()
ctx = gg.get_context(0)
ic = gg.initial_conditions
model = gg.AtlasDiscreteModel(gg.CubedSphereMesh(1.0), n=3)
for p, deg in ((1, 6), (2, 8)):
    V = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="H1")
    dΩ = gg.Measure(gg.Triangulation(model), deg)
    assem = gg.SparseMatrixAssembler(V, V)
    u = np.random.default_rng(1).standard_normal(V.num_free_dofs)
    A, b = gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dΩ, u=u), assem)
    gg.assemble_matrix(gg.Helmholtz(dΩ), assem)
    ns = gg.CGSolver(rtol=1e-10).numerical_setup(gg.assemble_matrix(gg.ScalarMass(dΩ), assem))
    ns.solve(b)
shell = gg.AtlasDiscreteModel(gg.CubedSphereWithThicknessMesh(1.0, 0.2), n=2)
Vs = gg.TestFESpace(shell, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1")
dS = gg.Measure(gg.Triangulation(shell), 8)
As = gg.SparseMatrixAssembler(Vs, Vs)
gg.assemble_matrix_and_vector(gg.LaplaceBeltrami(dS, u=np.ones(Vs.num_free_dofs)), As)
sub = shell.restrict(np.arange(0, 20))
Vq = gg.TestFESpace(sub, gg.ReferenceFE(gg.lagrangian, float, 2), conformity="H1")
gg.assemble_matrix(gg.LaplaceBeltrami(gg.Measure(gg.Triangulation(sub), 8)), Vq, Vq)
dS.integrate()
for p in (0, 1, 2):
    Vr = gg.TestFESpace(model, gg.ReferenceFE(gg.raviart_thomas, float, p), conformity="HDiv")
    Q = gg.TestFESpace(model, gg.ReferenceFE(gg.lagrangian, float, p), conformity="L2")
    x0 = np.concatenate([gg.interpolate(ic.williamson2_velocity, Vr).get(), gg.interpolate(ic.williamson2_depth, Q).get()])
    rk = gg.RungeKutta(gg.CGSolver(rtol=1e-12), None, 0.01, "EXRK_SSP_3_3")
    for op in (gg.WaveOperator(model, p), gg.ShallowWaterOperator(model, p, ic.coriolis, gravity=ic.GRAVITY),
               gg.ShallowWaterOperator(model, p, ic.coriolis, gravity=ic.GRAVITY, q0_mode="start_of_step")):
        st = gg.RKStepper(rk, op)
        st.set_state(x0)
        st.step(2)
        assert np.isfinite(st.get_state()).all()
    if p == 1:
        stt = gg.RKStepper(rk, gg.ThermalShallowWaterOperator(model, p, ic.coriolis))
        stt.set_state(np.concatenate([x0, 0.1 * x0[Vr.num_free_dofs:]]))
        stt.step(2)
        assert np.isfinite(stt.get_state()).all()
    wind = lambda X: np.stack([-X[..., 1], X[..., 0], 0 * X[..., 0]], -1)   # noqa: E731
    sa = gg.RKStepper(gg.RungeKutta(None, None, 0.01, "EXRK_SSP_2_2"), gg.AdvectionOperator(model, p, wind, degree=max(4 * p, 2)))
    sa.set_state(gg.interpolate(ic.williamson2_depth, Q).get())
    sa.step(2)
    gg.norm(sa.get_state(), Q, 6)
print("sanitize_small ok, launches", ctx.launches)
