"""Host-side mirror of the Gridap-facing API of GridapGeosciences.jl for the GPU hot path.

Julia is not available in this environment, so the host side above the C ABI (include/geosgpu.h) is
written in Python with the reference's own names and argument meaning
(`AtlasDiscreteModel(CubedSphereMesh(r), nref)`, `Triangulation`, `Measure`, `TestFESpace`,
`assemble_matrix`, `assemble_vector`, `AffineFEOperator`, `RungeKutta` / `solve`); INTEGRATION.md shows the
equivalent Julia `ccall` shim.  Every numeric call goes through libgeosgpu.so -- hand-written sm_100a
kernels -- and fails loudly when the library or a B200 is missing.  Nothing here imports `oracle/`.

Deviation from the reference: Gridap forms are arbitrary Julia closures; here a form is
a descriptor object (`LaplaceBeltrami(dΩ)`, `Helmholtz(dΩ)`, `ScalarMass(dΩ)`, `RTMass(dΩ)`, `Div(dΩ)`,
`Source(dΩ, f_q)`) that selects one of the library's kernels.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from . import initial_conditions  # noqa: F401
from ._lib import GGError  # noqa: F401

lagrangian = "lagrangian"
raviart_thomas = "raviart_thomas"


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _ip64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


# ---- context ----------------------------------------------------------------------------------------

class Context:
    """One per process: one GPU, one stream (`gg_init`)."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.gg_init(int(device), C.byref(h)))
        self.h = h
        self.device = device

    def synchronize(self):
        _lib.check(self.lib.gg_synchronize(self.h))

    @property
    def stream(self):
        return self.lib.gg_stream(self.h)

    @property
    def launches(self):
        return int(self.lib.gg_launch_count(self.h))

    def set_geometry_classes(self, enabled=True):
        """One Laplace-Beltrami element matrix per geometry class instead of per cell (same numbers; see geosgpu.h)."""
        _lib.check(self.lib.gg_set_geometry_classes(self.h, 1 if enabled else 0))

    def set_async(self, enabled=True):
        _lib.check(self.lib.gg_set_async(self.h, 1 if enabled else 0))

    def profile(self, enabled=True):
        _lib.check(self.lib.gg_profile_enable(self.h, 1 if enabled else 0))

    def profile_query(self, name):
        n, ms = C.c_int64(), C.c_double()
        _lib.check(self.lib.gg_profile_query(self.h, name.encode(), C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def measure_fp64_peak(self):
        out = C.c_double()
        _lib.check(self.lib.gg_measure_fp64_peak(self.h, C.byref(out)))
        return out.value

    def close(self):
        if self.h:
            self.lib.gg_finalize(self.h)
            self.h = None


_default_ctx = None


def get_context(device=None):
    global _default_ctx
    if _default_ctx is None:
        import os
        dev = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        _default_ctx = Context(dev)
    return _default_ctx


# ---- geometry / model ---------------------------------------------------------------------------------

class IntrinsicManifold:
    pass


class ExtrinsicManifold:
    pass


class CubedSphereMesh:
    """CubedSphereMesh(radius) -- src/Geometry/CubedSphereMesh.jl:32-35."""

    def __init__(self, radius=1.0):
        self.radius = float(radius)


class CubedSphereWithThicknessMesh:
    """CubedSphereWithThicknessMesh(radius, thickness) -- src/Geometry/CubedSphereWithThicknessMesh.jl:11-16: the 3-D shell
    between the spheres of radius `radius` and `radius + thickness`, chart coordinates (γ, α, β)."""

    def __init__(self, radius=1.0, thickness=0.1):
        self.radius, self.thickness = float(radius), float(thickness)


class AtlasDiscreteModel:
    """AtlasDiscreteModel(CubedSphereMesh(r), nref; manifold_style) -- src/Geometry/AtlasDiscreteModels.jl:222-229.

    `n` (cells per panel edge) may be given instead of `nref` (n = 2**nref); the reference only offers powers of two."""

    def __init__(self, mesh, nref=None, *, n=None, n_layers=None, ordering=0, manifold_style=None, ctx=None, _handle=None):
        self.ctx = ctx or get_context()
        self.lib = self.ctx.lib
        self.mesh = mesh
        self.manifold_style = manifold_style or IntrinsicManifold()
        self.dim = 3 if isinstance(mesh, CubedSphereWithThicknessMesh) else 2
        if _handle is not None:
            self.h = _handle
            self.n = 0
        else:
            if n is None:
                if nref is None:
                    raise ValueError("give nref or n")
                n = 2 ** int(nref)
            self.n = int(n)
            h = C.c_void_p()
            if self.dim == 3:
                # the reference refines the 6 coarse hexahedra isotropically: n cells through the shell unless n_layers is given
                _lib.check(self.lib.gg_mesh_cubed_sphere(self.ctx.h, 3, self.n, int(n_layers or 0), mesh.radius, mesh.thickness, int(ordering), C.byref(h)))
                self.n_layers = int(n_layers or self.n)
            else:
                _lib.check(self.lib.gg_mesh_cubed_sphere(self.ctx.h, 2, self.n, 1, mesh.radius, 0.0, int(ordering), C.byref(h)))
            self.h = h
        nc, nn, ne = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.gg_mesh_info(self.h, C.byref(nc), C.byref(nn), C.byref(ne)))
        self.num_cells, self.num_nodes, self.num_edges = nc.value, nn.value, ne.value

    @classmethod
    def from_arrays(cls, mesh, cell_to_chart, cell_chart_coords, cell_node_ids, n_nodes, ctx=None, manifold_style=None):
        """Connectivity owned by the caller (1-based ids, as Gridap Tables)."""
        ctx = ctx or get_context()
        c2c = np.ascontiguousarray(cell_to_chart, dtype=np.int32)
        xy = np.ascontiguousarray(cell_chart_coords, dtype=np.float64)
        cn = np.ascontiguousarray(cell_node_ids, dtype=np.int32)
        h = C.c_void_p()
        _lib.check(ctx.lib.gg_mesh_from_arrays(ctx.h, 2, len(c2c), int(n_nodes), _ip32(c2c), _dp(xy), _ip32(cn), mesh.radius, 0.0, C.byref(h)))
        return cls(mesh, ctx=ctx, manifold_style=manifold_style, _handle=h)

    def restrict(self, cells):
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        h = C.c_void_p()
        _lib.check(self.lib.gg_mesh_restrict(self.h, len(cells), _ip64(cells), C.byref(h)))
        sub = AtlasDiscreteModel(self.mesh, ctx=self.ctx, manifold_style=self.manifold_style, _handle=h)
        sub.n, sub.n_layers = self.n, getattr(self, "n_layers", 0)
        return sub

    def ghost_cells(self, first, last):
        n = C.c_int64()
        _lib.check(self.lib.gg_mesh_ghost_cells(self.h, first, last, C.byref(n), None))
        out = np.empty(n.value, dtype=np.int64)
        _lib.check(self.lib.gg_mesh_ghost_cells(self.h, first, last, C.byref(n), _ip64(out)))
        return out

    def get_cell_node_ids(self):
        out = np.empty((self.num_cells, 2 ** self.dim), dtype=np.int32)
        _lib.check(self.lib.gg_mesh_get_cell_nodes(self.h, _ip32(out)))
        return out

    def get_cell_to_chart(self):
        out = np.empty(self.num_cells, dtype=np.int32)
        _lib.check(self.lib.gg_mesh_get_cell_to_chart(self.h, _ip32(out)))
        return out

    def get_cell_coordinates(self):
        out = np.empty((self.num_cells, 2 ** self.dim, self.dim))
        _lib.check(self.lib.gg_mesh_get_chart_coords(self.h, _dp(out)))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.lib.gg_mesh_destroy(self.h)
        except Exception:
            pass


class PartitionPlan:
    """Partition of AtlasDiscreteModel(CubedSphereMesh(r), n) over `nparts` ranks as seen by rank `part`
    (src/Distributed/AtlasDistributedDiscreteModels.jl:20-73): owned cell range, ghost layer, and per FE space the
    local -> global DOF map, DOF owners and exchange lists.  Pure host computation (no GPU, no communication)."""

    def __init__(self, n, nparts, part, radius=1.0, thickness=None, ordering=0, n_layers=0, partition="linear"):
        """thickness given: the 3-D shell (AtlasOctreeDistributedDiscreteModel on CubedSphereWithThicknessMesh; ordering 2 = the
        p8est cell order, equal-count chunks of the space-filling curve).  partition="strips" (2-D sphere, opt-in): every rank owns
        the same rows of cells on all six panels instead of the reference's linear cut (include/geosgpu.h, gg_plan_create_strips)."""
        self.lib = _lib.load()
        self.n, self.nparts, self.part, self.radius = int(n), int(nparts), int(part), float(radius)
        if partition not in ("linear", "strips"):
            raise ValueError("partition must be 'linear' or 'strips'")
        self.partition = partition
        h = C.c_void_p()
        if partition == "strips":
            if thickness is not None:
                raise ValueError("panel strips are a partition of the 2-D sphere")
            _lib.check(self.lib.gg_plan_create_strips(self.n, self.radius, self.nparts, self.part, C.byref(h)))
        elif thickness is not None:
            _lib.check(self.lib.gg_plan_create_shell(self.n, int(n_layers), self.radius, float(thickness), int(ordering), self.nparts,
                                                     self.part, C.byref(h)))
        else:
            _lib.check(self.lib.gg_plan_create(self.n, self.radius, self.nparts, self.part, C.byref(h)))
        self.h = h
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.gg_plan_info(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        self.num_cells_global, self.first_owned, self.last_owned, self.num_local_cells = a.value, b.value, c.value, d.value
        no = C.c_int64()
        _lib.check(self.lib.gg_plan_owned_cells(self.h, C.byref(no)))
        self.num_owned_cells = no.value

    def local_cells(self):
        out = np.empty(self.num_local_cells, dtype=np.int64)
        _lib.check(self.lib.gg_plan_local_cells(self.h, _ip64(out)))
        return out

    def space(self, kind, order):
        """dict with n_global, n_local, n_owned, slot_capacity, l2g, owner and the neighbour lists of space (kind, order)."""
        ng, nl, no, nn, cap = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_int64()
        _lib.check(self.lib.gg_plan_space_info(self.h, kind, order, C.byref(ng), C.byref(nl), C.byref(no), C.byref(nn), C.byref(cap)))
        l2g, owner = np.empty(nl.value, dtype=np.int64), np.empty(nl.value, dtype=np.int32)
        _lib.check(self.lib.gg_plan_space_maps(self.h, kind, order, _ip64(l2g), _ip32(owner)))
        nbrs = {}
        for k in range(nn.value):
            peer, ns, nr = C.c_int32(), C.c_int64(), C.c_int64()
            _lib.check(self.lib.gg_plan_space_neighbor(self.h, kind, order, k, C.byref(peer), C.byref(ns), C.byref(nr), None, None, None))
            send, remote, recv = (np.empty(ns.value, dtype=np.int32), np.empty(ns.value, dtype=np.int32),
                                  np.empty(nr.value, dtype=np.int32))
            _lib.check(self.lib.gg_plan_space_neighbor(self.h, kind, order, k, C.byref(peer), C.byref(ns), C.byref(nr),
                                                       _ip32(send), _ip32(remote), _ip32(recv)))
            nbrs[peer.value] = {"send": send, "send_remote": remote, "recv": recv}
        return {"n_global": ng.value, "n_local": nl.value, "n_owned": no.value, "slot_capacity": cap.value,
                "l2g": l2g, "owner": owner, "neighbors": nbrs}

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.gg_plan_destroy(self.h)
        except Exception:
            pass


class PeerWindow:
    """This rank's peer-visible window (flags, reduction operands, one exchange slot) and the IPC mappings of the other
    ranks' windows.  `connect()` exchanges the 64-byte CUDA-IPC handles with torch.distributed (the only set-up
    communication of the multi-GPU path)."""

    def __init__(self, ctx, rank, world, slot_doubles):
        self.ctx, self.lib, self.rank, self.world = ctx, ctx.lib, int(rank), int(world)
        h = C.c_void_p()
        _lib.check(self.lib.gg_comm_create(ctx.h, self.rank, self.world, int(slot_doubles), C.byref(h)))
        self.h = h

    def handle(self):
        buf = (C.c_uint8 * 64)()
        _lib.check(self.lib.gg_comm_handle(self.h, buf))
        return bytes(buf)

    def connect(self, handles=None):
        if self.world == 1:
            return
        if handles is None:
            import torch
            import torch.distributed as dist
            mine = torch.tensor(list(self.handle()), dtype=torch.uint8)
            dev = torch.device("cuda", self.ctx.device) if dist.get_backend() == "nccl" else torch.device("cpu")
            mine = mine.to(dev)
            allh = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(allh, mine)
            handles = b"".join(bytes(t.cpu().tolist()) for t in allh)
        buf = (C.c_uint8 * (64 * self.world)).from_buffer_copy(handles)
        _lib.check(self.lib.gg_comm_connect(self.h, buf))

    def error_epoch(self):
        out = C.c_int64()
        _lib.check(self.lib.gg_comm_error(self.h, C.byref(out)))
        return out.value

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.lib.gg_comm_destroy(self.h)
        except Exception:
            pass


def window_doubles(plan, p_fe):
    """Exchange-slot size that fits every space of the RT_p x DG_p / CG_{p+1} steppers on this plan (and the vector H1 space
    of order max(p, 1))."""
    cap = 0
    spaces = [(_lib.SPACE_RT, p_fe), (_lib.SPACE_DG, p_fe), (_lib.SPACE_CG, p_fe + 1)]
    if max(p_fe, 1) <= 2:
        spaces.append((_lib.SPACE_CGV, max(p_fe, 1)))
    for kind, order in spaces:
        cap = max(cap, plan.space(kind, order)["slot_capacity"])
    return cap


def partitioned_model(mesh, plan, comm=None, ctx=None, manifold_style=None):
    """AtlasDiscreteModel(ranks, mesh, nref): this rank's sub-model (owned + ghost cells) on the device."""
    ctx = ctx or get_context()
    h = C.c_void_p()
    _lib.check(ctx.lib.gg_mesh_partitioned(ctx.h, plan.h, comm.h if comm is not None else None, C.byref(h)))
    m = AtlasDiscreteModel(mesh, ctx=ctx, manifold_style=manifold_style, _handle=h)
    m.plan, m.comm = plan, comm
    return m


def consistent(v, V):
    """consistent!(v): owner -> ghost update of a DOF vector of the partitioned space V (collective)."""
    _lib.check(V.lib.gg_vec_consistent(V.h, v.h))
    return v


def global_dof_ids(V):
    """(1-based global DOF id, owner rank) of every local DOF of a partitioned space."""
    l2g, owner = np.empty(V.num_free_dofs, dtype=np.int64), np.empty(V.num_free_dofs, dtype=np.int32)
    _lib.check(V.lib.gg_space_global_dofs(V.h, _ip64(l2g), _ip32(owner)))
    return l2g, owner


def partition_range(num_cells, nparts, part):
    """Owned cell range of the reference's linear partition (AtlasDistributedDiscreteModels.jl:31), 0-based [first,last)."""
    lib = _lib.load()
    a, b = C.c_int64(), C.c_int64()
    _lib.check(lib.gg_partition_range(num_cells, nparts, part, C.byref(a), C.byref(b)))
    return a.value, b.value


def num_cells(model):
    return model.num_cells


def get_sphere_radius(model):
    return model.mesh.radius


def dx(model):
    """src/ConvergenceTools/Tools.jl:228-232."""
    return math.sqrt(4 * math.pi * model.mesh.radius ** 2 / model.num_cells)


class Triangulation:
    def __init__(self, model):
        self.model = model


class Measure:
    """Measure(Ω, degree): tensor Gauss-Legendre rule with ceil((degree+1)/2) points per direction."""

    def __init__(self, trian, degree):
        self.trian = trian
        self.model = trian.model
        self.degree = int(degree)
        self.nq1 = (self.degree + 2) // 2
        self.num_points = self.nq1 ** self.model.dim

    def quadrature_points(self, chart=True, ambient=True):
        m = self.model
        ch = np.empty((m.num_cells, self.num_points, m.dim)) if chart else None
        xyz = np.empty((m.num_cells, self.num_points, 3)) if ambient else None
        _lib.check(m.lib.gg_quadrature_points(m.h, self.degree, _dp(ch) if chart else None, _dp(xyz) if ambient else None))
        return ch, xyz

    def integrate(self, fq=None):
        """sum(∫(f * meas_cf)dΩ) with f given at the quadrature points (None = 1)."""
        m = self.model
        out = C.c_double()
        v = fq
        if fq is not None and not isinstance(fq, DeviceVector):
            v = DeviceVector.from_numpy(m.ctx, np.ascontiguousarray(fq, dtype=np.float64).reshape(-1))
        _lib.check(m.lib.gg_integrate(m.h, self.degree, v.h if v is not None else None, C.byref(out)))
        return out.value


class VectorValue:
    """Marker for ReferenceFE(lagrangian, VectorValue, p): the vector-valued Lagrangian element (VectorValue{2,Float64})."""


def ReferenceFE(name, T, order):
    if T is VectorValue:
        return (name, int(order), "vector")
    return (name, int(order))


class FESpace:
    def __init__(self, model, kind, order, constraint=0, _handle=None):
        self.model = model
        self.lib = model.lib
        self.kind, self.order = kind, order
        if _handle is None:
            h = C.c_void_p()
            _lib.check(self.lib.gg_space_builtin(model.h, kind, order, constraint, C.byref(h)))
            _handle = h
        self.h = _handle
        nf, ml, nd = C.c_int64(), C.c_int32(), C.c_int64()
        _lib.check(self.lib.gg_space_info(self.h, C.byref(nf), C.byref(ml), C.byref(nd)))
        self.num_free_dofs, self.ndofs_loc, self.num_dirichlet_dofs = nf.value, ml.value, nd.value

    @classmethod
    def from_arrays(cls, model, kind, order, n_free, cell_dof_ids, cell_dof_signs=None):
        ids = np.ascontiguousarray(cell_dof_ids, dtype=np.int32)
        sg = None if cell_dof_signs is None else np.ascontiguousarray(cell_dof_signs, dtype=np.int8)
        h = C.c_void_p()
        _lib.check(model.lib.gg_space_from_arrays(model.h, kind, order, int(n_free), _ip32(ids),
                                                  None if sg is None else sg.ctypes.data_as(C.POINTER(C.c_int8)), C.byref(h)))
        return cls(model, kind, order, _handle=h)

    def get_cell_dof_ids(self):
        out = np.empty((self.model.num_cells, self.ndofs_loc), dtype=np.int32)
        _lib.check(self.lib.gg_space_get_cell_dofs(self.h, _ip32(out)))
        return out

    def get_cell_dof_signs(self):
        out = np.empty((self.model.num_cells, self.ndofs_loc), dtype=np.int8)
        _lib.check(self.lib.gg_space_get_cell_signs(self.h, out.ctypes.data_as(C.POINTER(C.c_int8))))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.model.ctx.h:
                self.lib.gg_space_destroy(self.h)
        except Exception:
            pass


def TestFESpace(model_or_trian, reffe, conformity="H1", constraint=None, dirichlet_tags=None):
    model = model_or_trian.model if isinstance(model_or_trian, Triangulation) else model_or_trian
    name, order = reffe[0], reffe[1]
    conformity = conformity.lstrip(":")
    if name == lagrangian and len(reffe) > 2:
        if conformity != "H1":
            raise ValueError("vector-valued Lagrangian spaces are implemented for conformity=:H1")
        kind = _lib.SPACE_CGV
    elif name == lagrangian:
        kind = {"H1": _lib.SPACE_CG, "L2": _lib.SPACE_DG}[conformity]
    elif name == raviart_thomas:
        kind = _lib.SPACE_RT
    else:
        raise ValueError(f"unknown reference FE {name}")
    cons = {None: 0, "zeromean": 1, ":zeromean": 1}[constraint]
    if dirichlet_tags is not None:
        # test/Tutorials/LinearBoussinesq.jl:55-57: RT space on the shell with no flux through the bottom and top boundaries
        if kind != _lib.SPACE_RT or model.dim != 3 or sorted(dirichlet_tags) != ["bottom_boundary", "top_boundary"] or cons:
            raise ValueError('dirichlet_tags=["bottom_boundary", "top_boundary"] on a 3-D Raviart-Thomas space is what is implemented')
        cons = _lib.CONSTRAINT_NOFLUX
    return FESpace(model, kind, order, cons)


def TrialFESpace(V):
    return V


def num_free_dofs(V):
    return V.num_free_dofs


# ---- device vectors / matrices --------------------------------------------------------------------------

class DeviceVector:
    def __init__(self, ctx, n):
        self.ctx, self.lib, self.n = ctx, ctx.lib, int(n)
        h = C.c_void_p()
        _lib.check(self.lib.gg_vec_create(ctx.h, self.n, C.byref(h)))
        self.h = h

    @classmethod
    def from_numpy(cls, ctx, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        v = cls(ctx, a.size)
        v.set(a)
        return v

    def set(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
        assert a.size == self.n
        _lib.check(self.lib.gg_vec_set(self.h, _dp(a)))

    def get(self, out=None):
        out = np.empty(self.n) if out is None else out
        _lib.check(self.lib.gg_vec_get(self.h, _dp(out)))
        return out

    def fill(self, v):
        _lib.check(self.lib.gg_vec_fill(self.h, float(v)))

    @property
    def device_ptr(self):
        return self.lib.gg_vec_device_ptr(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.lib.gg_vec_destroy(self.h)
        except Exception:
            pass


class CSRMatrix:
    """Device CSR matrix with a fixed pattern (`gg_pattern`); `to_scipy()` exports it."""

    def __init__(self, test, trial, skeleton=False):
        self.test, self.trial = test, trial
        self.ctx, self.lib = test.model.ctx, test.lib
        h = C.c_void_p()
        if skeleton:
            # SparseMatrixAssembler on a form with SkeletonTriangulation terms: facet-neighbour coupling (DG)
            assert test is trial
            _lib.check(self.lib.gg_pattern_skeleton(test.h, C.byref(h)))
        else:
            _lib.check(self.lib.gg_pattern(test.h, trial.h, C.byref(h)))
        self.h = h
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.gg_csr_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        self.shape, self.nnz = (a.value, b.value), c.value

    def pattern(self, index_base=0):
        rp = np.empty(self.shape[0] + 1, dtype=np.int64)
        ci = np.empty(self.nnz, dtype=np.int64)
        _lib.check(self.lib.gg_csr_export(self.h, _ip64(rp), _ip64(ci), None, index_base))
        return rp, ci

    def values(self, out=None):
        out = np.empty(self.nnz) if out is None else out
        _lib.check(self.lib.gg_csr_export(self.h, None, None, _dp(out), 0))
        return out

    def to_scipy(self):
        import scipy.sparse as sp
        rp, ci = self.pattern()
        return sp.csr_matrix((self.values(), ci, rp), shape=self.shape)

    def to_scipy_csc(self, index_base=0):
        """The matrix as Julia's SparseMatrixCSC sees it (`gg_csr_export_csc`): colptr, rowval, nzval."""
        import scipy.sparse as sp
        cp = np.empty(self.shape[1] + 1, dtype=np.int64)
        rv = np.empty(self.nnz, dtype=np.int64)
        nz = np.empty(self.nnz)
        _lib.check(self.lib.gg_csr_export_csc(self.h, _ip64(cp), _ip64(rv), _dp(nz), index_base))
        if index_base:
            return cp, rv, nz
        return sp.csc_matrix((nz, rv, cp), shape=self.shape)

    def csc_values(self, out=None):
        out = np.empty(self.nnz) if out is None else out
        _lib.check(self.lib.gg_csr_export_csc(self.h, None, None, _dp(out), 0))
        return out

    def spmv(self, x, y, alpha=1.0, beta=0.0):
        _lib.check(self.lib.gg_csr_spmv(self.h, alpha, x.h, beta, y.h))

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.lib.gg_csr_destroy(self.h)
        except Exception:
            pass


# ---- forms -----------------------------------------------------------------------------------------------

class Form:
    form_id = None
    bilinear = True

    def __init__(self, dΩ, u=None, qdata=None, dirichlet=0.0, scale=1.0):
        self.dΩ, self.u, self.qdata, self.dirichlet, self.scale = dΩ, u, qdata, dirichlet, scale
        self._keep = []

    def args(self, test, trial):
        ctx = test.model.ctx
        a = _lib.FormArgs()
        a.test = test.h
        a.trial = trial.h if trial is not None else None
        a.degree = self.dΩ.degree
        self._keep = []
        for name in ("u", "qdata"):
            v = getattr(self, name)
            if v is not None and not isinstance(v, DeviceVector):
                v = DeviceVector.from_numpy(ctx, v)
            self._keep.append(v)
            setattr(a, name, v.h if v is not None else None)
        a.dirichlet = self.dirichlet
        a.scale = self.scale
        return a


class LaplaceBeltrami(Form):
    """a(u,v) = ∫( ∇(v)⋅(inv_metric_cf⋅∇(u)) * meas_cf )dΩ -- test/Laplacian/LaplaceBeltrami.jl:47.
    With `extrinsic=True`: a(u,v) = ∫( ∇(v)⋅∇(u) )dΩ on the ExtrinsicManifold model (AmbientLaplaceBeltrami.jl:44)."""

    def __init__(self, dΩ, extrinsic=None, **kw):
        super().__init__(dΩ, **kw)
        if extrinsic is None:
            extrinsic = isinstance(dΩ.model.manifold_style, ExtrinsicManifold)
        self.form_id = _lib.FORM_LAPLACE_BELTRAMI_EXTRINSIC if extrinsic else _lib.FORM_LAPLACE_BELTRAMI


class Helmholtz(Form):
    """a(u,v) = ∫(u*v*meas_cf)dΩ - ∫( ∇(v)⋅(inv_metric_cf⋅∇(u)) * meas_cf )dΩ -- test/Laplacian/Helmholtz.jl:43."""
    form_id = _lib.FORM_HELMHOLTZ


class ScalarMass(Form):
    """m(p,q) = ∫( (q*p)*meas_cf )dΩ (optionally times a coefficient given at the quadrature points)."""
    form_id = _lib.FORM_MASS_SCALAR


class RTMass(Form):
    """m(u,v) = ∫( (v⋅(metric_cf⋅u))*(1.0/meas_cf) )dΩ -- test/Transient/TransientWaveEquation.jl:63."""
    form_id = _lib.FORM_MASS_RT


class Div(Form):
    """b(u,q) = ∫( q*(∇⋅u) )dΩ -- test/Transient/TransientWaveEquation.jl:64."""
    form_id = _lib.FORM_DIV


class VectorMass(Form):
    """a(u,v) = ∫( (u⋅(metric_cf⋅v))*meas_cf )dΩ on a vector H1 space -- test/Projection/L2_projection_Lagrangian_vector.jl:41."""
    form_id = _lib.FORM_MASS_VECTOR


class VectorLaplacian(Form):
    """a(u,v) = ∫( (divg_product_rule(u)*divg_product_rule(v))*(1/meas_cf) )dΩ + ∫( -1.0*(divergenceRgu(u)*divergenceRgu(v))*(1/meas_cf) )dΩ
    on a vector H1 space -- test/SurfaceStokes/VectorLaplacian2D.jl:71-73 (ν of the Stokes driver through `scale`)."""
    form_id = _lib.FORM_VECTOR_LAPLACIAN


class VectorDiv(Form):
    """b(u,q) = ∫( q*divg_product_rule(u) )dΩ, Lagrangian test x vector H1 trial -- test/SurfaceStokes/SurfaceStokes_TaylorHood.jl:106."""
    form_id = _lib.FORM_VECTOR_DIV


class VectorSource(Form):
    """l(v) = ∫( (f⋅(metric_cf⋅v))*meas_cf )dΩ, f = contravariant components (n_cells, Q, 2) at the quadrature points
    -- test/Projection/L2_projection_Lagrangian_vector.jl:42."""
    form_id = _lib.FORM_SOURCE_VECTOR
    bilinear = False

    def __init__(self, dΩ, fq, **kw):
        if not isinstance(fq, DeviceVector):
            fq = np.ascontiguousarray(fq, dtype=np.float64).reshape(-1)
        super().__init__(dΩ, qdata=fq, **kw)


class ShellBoundaryFlux(Form):
    """boundary(v) = ∫( ((inv_metric_cf⋅∇(f_int))⋅nΓ)*v*meas_cf )dΓ over the bottom and top boundaries of the shell, f_int = `u`
    -- the 3-D branch of test/Laplacian/LaplaceBeltrami.jl:58-64."""
    form_id = _lib.FORM_SHELL_BOUNDARY_FLUX
    bilinear = False


class Source(Form):
    """l(v) = ∫( (f*v)*meas_cf )dΩ with f at the quadrature points -- test/Laplacian/LaplaceBeltrami.jl:53."""
    form_id = _lib.FORM_SOURCE
    bilinear = False

    def __init__(self, dΩ, fq, **kw):
        if not isinstance(fq, DeviceVector):
            fq = np.ascontiguousarray(fq, dtype=np.float64).reshape(-1)
        super().__init__(dΩ, qdata=fq, **kw)


# ---- assembly ---------------------------------------------------------------------------------------------

class SparseMatrixAssembler:
    """SparseMatrixAssembler(U, V): owns the device CSR pattern (built once) and the numeric-phase maps."""

    def __init__(self, U, V):
        self.trial, self.test = U, V
        self.matrix = CSRMatrix(V, U)


def allocate_matrix(assem):
    return assem.matrix


def assemble_matrix(form, U_or_assem, V=None, add=False):
    """assemble_matrix(a, U, V) / assemble_matrix!(A, assem, ...): numeric phase on the device."""
    assem = U_or_assem if isinstance(U_or_assem, SparseMatrixAssembler) else SparseMatrixAssembler(U_or_assem, V)
    A = assem.matrix
    args = form.args(assem.test, assem.trial)
    _lib.check(A.lib.gg_assemble_matrix(form.form_id, C.byref(args), A.h, 1 if add else 0))
    A._assembler = assem
    return A


def assemble_vector(form, V, U=None, out=None, add=False):
    """assemble_vector(l, V).  For a bilinear form with `form.u` set this is the residual a(u_h, v)."""
    ctx = V.model.ctx
    b = out if out is not None else DeviceVector(ctx, V.num_free_dofs)
    trial = U if U is not None else (V if form.bilinear else None)
    args = form.args(V, trial)
    _lib.check(V.lib.gg_assemble_vector(form.form_id, C.byref(args), b.h, 1 if add else 0))
    return b


def assemble_matrix_and_vector(form, assem, b=None):
    """Jacobian and residual in one pass over the cells (A = a(.,.), b_i = a(u_h, φ_i))."""
    A = assem.matrix
    b = b if b is not None else DeviceVector(A.ctx, assem.test.num_free_dofs)
    args = form.args(assem.test, assem.trial)
    _lib.check(A.lib.gg_assemble_matrix_and_vector(form.form_id, C.byref(args), A.h, b.h))
    return A, b


def element_matrices(form, U, V, first=0, n=None):
    n = V.model.num_cells - first if n is None else n
    out = np.empty((n, V.ndofs_loc, U.ndofs_loc))
    args = form.args(V, U)
    _lib.check(V.lib.gg_element_matrices(form.form_id, C.byref(args), first, n, _dp(out)))
    return out


def element_vectors(form, V, U=None, first=0, n=None):
    n = V.model.num_cells - first if n is None else n
    out = np.empty((n, V.ndofs_loc))
    trial = U if U is not None else (V if form.bilinear else None)
    args = form.args(V, trial)
    _lib.check(V.lib.gg_element_vectors(form.form_id, C.byref(args), first, n, _dp(out)))
    return out


def eval_geometry(panel, radius, pts, ctx=None):
    """Device evaluation of CubedSphereMap / its Jacobian / metric / inverse metric / sqrt(det g) at chart points."""
    ctx = ctx or get_context()
    pts = np.ascontiguousarray(pts, dtype=np.float64).reshape(-1, 2)
    n = pts.shape[0]
    xyz, jac, g, gi, sq = np.empty((n, 3)), np.empty((n, 2, 3)), np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
    _lib.check(ctx.lib.gg_eval_geometry(ctx.h, panel, radius, n, _dp(pts), _dp(xyz), _dp(jac), _dp(g), _dp(gi), _dp(sq)))
    return xyz, jac, g, gi, sq


class AffineFEOperator:
    """AffineFEOperator(a, l, U, V): assembles A and b on the device (test/Laplacian/LaplaceBeltrami.jl:68)."""

    def __init__(self, a, l, U, V):
        self.assem = SparseMatrixAssembler(U, V)
        self.A = assemble_matrix(a, self.assem)
        self.b = assemble_vector(l, V)

    def get_matrix(self):
        return self.A

    def get_vector(self):
        return self.b


# ---- interpolation / norms ---------------------------------------------------------------------------------

def interpolation_points(V, chart=False, ambient=True):
    """Points at which `interpolate` samples its function: (n_cells, n_pts, 2|3) chart / ambient coordinates."""
    n = C.c_int32()
    _lib.check(V.lib.gg_space_interp_points(V.h, C.byref(n), None, None))
    nc = V.model.num_cells
    ch = np.empty((nc, n.value, V.model.dim)) if chart else None
    xyz = np.empty((nc, n.value, 3)) if ambient else None
    _lib.check(V.lib.gg_space_interp_points(V.h, C.byref(n), _dp(ch) if chart else None, _dp(xyz) if ambient else None))
    return ch, xyz


def interpolate(fun, V, out=None):
    """interpolate(fun∘ambient_map_cf, V).  Lagrangian V: `fun(xyz)` is scalar.  Raviart-Thomas V: `fun(xyz)` is an
    ambient 3-vector field and the interpolated chart field is meas_cf*((pinvJ∘transpose∘∇(ambient_map_cf))⋅fun)
    (test/Transient/TransientShallowWater.jl:88-95)."""
    _, xyz = interpolation_points(V)
    vals = np.asarray(fun(xyz), dtype=np.float64)
    shape = xyz.shape if V.kind in (_lib.SPACE_RT, _lib.SPACE_CGV) else xyz.shape[:2]
    vals = np.ascontiguousarray(np.broadcast_to(vals, shape))
    out = out if out is not None else DeviceVector(V.model.ctx, V.num_free_dofs)
    _lib.check(V.lib.gg_space_interpolate(V.h, _dp(vals), out.h))
    return out


def norm(e, V, degree):
    """sqrt(sum(∫(e*e*meas_cf)dΩ)) for Lagrangian V, sqrt(sum(∫(e⋅(metric_cf⋅e)*(1/meas_cf))dΩ)) for RT V."""
    if not isinstance(e, DeviceVector):
        e = DeviceVector.from_numpy(V.model.ctx, e)
    out = C.c_double()
    _lib.check(V.lib.gg_norm_sq(V.h, e.h, int(degree), C.byref(out)))
    return math.sqrt(max(out.value, 0.0))


def h1_seminorm(e, V, degree):
    """sqrt(sum(∫( (gradient(e)⊙(inv_metric_cf⋅gradient(e)))*meas_cf )dΩ)) of a vector H1 FE function
    (test/SurfaceStokes/SurfaceStokes_TaylorHood.jl:122)."""
    if not isinstance(e, DeviceVector):
        e = DeviceVector.from_numpy(V.model.ctx, e)
    out = C.c_double()
    _lib.check(V.lib.gg_h1_seminorm_sq(V.h, e.h, int(degree), C.byref(out)))
    return math.sqrt(max(out.value, 0.0))


def surface_laplacian(dΩ, grad_f, hess_f):
    """Δs(f, Ω) evaluated at the quadrature points of dΩ for an ambient function f given by its ambient gradient and Hessian
    (closed form of src/CellData/SurfaceDiffOps.jl:26-65).  Returns a DeviceVector usable as `fq` of Source."""
    m = dΩ.model
    _, xyz = dΩ.quadrature_points(chart=False)
    g = np.ascontiguousarray(np.broadcast_to(grad_f(xyz), xyz.shape), dtype=np.float64)
    h = np.ascontiguousarray(np.broadcast_to(hess_f(xyz), xyz.shape + (3,)), dtype=np.float64)
    out = DeviceVector(m.ctx, m.num_cells * dΩ.num_points)
    _lib.check(m.lib.gg_surface_laplacian(m.h, dΩ.degree, _dp(g), _dp(h), out.h))
    return out


def surface_gradient(dΩ, grad_f):
    """∇s(f, Ω): contravariant components (n_cells, Q, 2) at the quadrature points (src/CellData/SurfaceDiffOps.jl:107-133)."""
    m = dΩ.model
    _, xyz = dΩ.quadrature_points(chart=False)
    g = np.ascontiguousarray(np.broadcast_to(grad_f(xyz), xyz.shape), dtype=np.float64)
    out = DeviceVector(m.ctx, 2 * m.num_cells * dΩ.num_points)
    _lib.check(m.lib.gg_surface_gradient(m.h, dΩ.degree, _dp(g), out.h))
    return out


def surface_divergence(dΩ, F, grad_F):
    """divs(F, Ω) of an ambient vector field at the quadrature points (src/CellData/SurfaceDiffOps.jl:228-259)."""
    m = dΩ.model
    _, xyz = dΩ.quadrature_points(chart=False)
    f = np.ascontiguousarray(np.broadcast_to(F(xyz), xyz.shape), dtype=np.float64)
    g = np.ascontiguousarray(np.broadcast_to(grad_F(xyz), xyz.shape + (3,)), dtype=np.float64)
    out = DeviceVector(m.ctx, m.num_cells * dΩ.num_points)
    _lib.check(m.lib.gg_surface_divergence(m.h, dΩ.degree, _dp(f), _dp(g), out.h))
    return out


def contravariant_components(dΩ, vecX):
    """(pinvJ∘transpose∘∇(ambient_map_cf))⋅(vecX∘ambient_map_cf) at the quadrature points: (n_cells, Q, 2) DeviceVector
    (the contravariant surface gradient operator applied to the ambient vector itself, src/Helpers/Operators.jl:16-24)."""
    m = dΩ.model
    _, xyz = dΩ.quadrature_points(chart=False)
    v = np.ascontiguousarray(np.broadcast_to(vecX(xyz), xyz.shape), dtype=np.float64)
    out = DeviceVector(m.ctx, 2 * m.num_cells * dΩ.num_points)
    _lib.check(m.lib.gg_surface_gradient(m.h, dΩ.degree, _dp(v), out.h))   # ginv . (J v) = pinvJ . v
    return out


def change_of_basis(V):
    """(n_cells, n_nodes, 2, 2) blocks of a vector H1 space (src/FESpaces/GradConformingFESpaces.jl:60-113)."""
    out = np.empty((V.model.num_cells, V.ndofs_loc // 2, 2, 2))
    _lib.check(V.lib.gg_space_change_of_basis(V.h, _dp(out)))
    return out


def l2_error(uh, V, dΩ, fq=None, dirichlet=0.0):
    """sqrt(sum(∫((f - uh)*(f - uh)*meas_cf)dΩ)) with f at the quadrature points of dΩ (test/Laplacian/LaplaceBeltrami.jl:82-83)."""
    ctx = V.model.ctx
    if not isinstance(uh, DeviceVector):
        uh = DeviceVector.from_numpy(ctx, uh)
    if fq is not None and not isinstance(fq, DeviceVector):
        fq = DeviceVector.from_numpy(ctx, np.ascontiguousarray(fq, dtype=np.float64).reshape(-1))
    out = C.c_double()
    _lib.check(V.lib.gg_l2_error_sq(V.h, uh.h, float(dirichlet), fq.h if fq is not None else None, dΩ.degree, C.byref(out)))
    return math.sqrt(max(out.value, 0.0))


def chart_mean(uh, V, dirichlet=0.0):
    """(∫uh dξ)/(∫dξ): the constant Gridap's ZeroMeanFESpace removes from a solution (FEFunction of a zeromean space)."""
    if not isinstance(uh, DeviceVector):
        uh = DeviceVector.from_numpy(V.model.ctx, uh)
    out = C.c_double()
    _lib.check(V.lib.gg_chart_mean(V.h, uh.h, float(dirichlet), C.byref(out)))
    return out.value


def laplace_beltrami_solver(model, order, f, grad_f, hess_f, degree=None, ls=None):
    """The driver of test/Laplacian/LaplaceBeltrami.jl:19-98 on the device: zero-mean Q_order space, a = LaplaceBeltrami,
    l = Source(-Δs f) (+ the boundary term on the 3-D shell), Jacobi-PCG solve, zero-mean shift, L2 error with a
    degree-2*degree rule.  Returns (error, uh, V, numerical setup)."""
    degree = 6 * (order + 1) if degree is None else degree
    Ω = Triangulation(model)
    dΩ, dΩe = Measure(Ω, degree), Measure(Ω, min(2 * degree, 31))  # device rules stop at 16 points per direction
    V = TestFESpace(model, ReferenceFE(lagrangian, float, order), conformity="H1", constraint="zeromean")
    if model.dim == 3:
        # the shell is a flat 3-D domain in curvilinear coordinates: Δs f is the Euclidean Laplacian of f
        _, xyz = dΩ.quadrature_points(chart=False)
        rhs = -np.trace(np.broadcast_to(hess_f(xyz), xyz.shape + (3,)), axis1=-2, axis2=-1)
        op = AffineFEOperator(LaplaceBeltrami(dΩ), Source(dΩ, rhs), V, V)
        # f_int = interpolate(f, U): only its gradient enters, so the nodal value at the DOF fixed by the zero-mean space is kept
        Uf = interpolate(f, TestFESpace(model, ReferenceFE(lagrangian, float, order), conformity="H1")).get()
        assemble_vector(ShellBoundaryFlux(dΩ, u=Uf[:-1], dirichlet=float(Uf[-1])), V, out=op.get_vector(), add=True)
    else:
        rhs = surface_laplacian(dΩ, grad_f, hess_f)
        op = AffineFEOperator(LaplaceBeltrami(dΩ), Source(dΩ, rhs, scale=-1.0), V, V)
    ns = (ls or CGSolver(rtol=1e-13, maxiter=20000)).numerical_setup(op.get_matrix())
    uh = ns.solve(op.get_vector())
    mean = chart_mean(uh, V)
    _, xyz = dΩe.quadrature_points(chart=False)
    err = l2_error(uh, V, dΩe, fq=np.asarray(f(xyz), dtype=np.float64) + mean)
    return err, uh, V, ns


# ---- linear solver ------------------------------------------------------------------------------------------

class CGSolver:
    """CGSolver(JacobiLinearSolver(); rtol) on the device (persistent single-launch PCG)."""

    def __init__(self, rtol=1e-12, maxiter=2000):
        self.rtol, self.maxiter = rtol, maxiter

    def numerical_setup(self, A):
        return NumericalSetup(self, A)


class NumericalSetup:
    def __init__(self, ls, A):
        self.A, self.lib = A, A.lib
        h = C.c_void_p()
        _lib.check(self.lib.gg_solver_pcg_create(A.h, ls.rtol, ls.maxiter, C.byref(h)))
        self.h = h
        self.iterations, self.rel_res = 0, 0.0

    def update(self):
        _lib.check(self.lib.gg_solver_update(self.h))

    def solve(self, b, x=None):
        x = x if x is not None else DeviceVector(self.A.ctx, self.A.shape[0])
        it, rr = C.c_int32(), C.c_double()
        _lib.check(self.lib.gg_solver_solve(self.h, b.h, x.h, C.byref(it), C.byref(rr)))
        self.iterations, self.rel_res = it.value, rr.value
        return x

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.A.ctx.h:
                self.lib.gg_solver_destroy(self.h)
        except Exception:
            pass


# ---- explicit Runge-Kutta (linear wave) -------------------------------------------------------------------

TABLEAUS = {  # Gridap names
    "EXRK_SSP_3_3": ([[0, 0, 0], [1, 0, 0], [0.25, 0.25, 0]], [1 / 6, 1 / 6, 2 / 3], [0, 1, 0.5]),
    "EXRK_SSP_2_2": ([[0, 0], [1, 0]], [0.5, 0.5], [0, 1.0]),
    "EXRK_Euler_1_1": ([[0.0]], [1.0], [0.0]),
}


class RungeKutta:
    """RungeKutta(ls, ls, dt, :EXRK_SSP_3_3)."""

    def __init__(self, sysslvr, sysslvr2, dt, tableau):
        self.ls = sysslvr
        self.dt = float(dt)
        self.tableau = tableau.lstrip(":")


class WaveOperator:
    """TransientSemilinearFEOperator(mass, res, (jac, jac_t), X, Y, constant_mass=true) of the linear wave driver
    (test/Transient/TransientWaveEquation.jl:63-68) for X = RT_p x DG_p."""

    def __init__(self, model, p_fe, degree=None):
        self.model, self.p_fe = model, p_fe
        self.degree = 5 * (p_fe + 1) if degree is None else degree


class ShallowWaterOperator:
    """DAEFEOperator(opT, opFE, ls_diag) for the rotating shallow-water forms of
    test/Transient/TransientShallowWater.jl:113-145: prognostic (u, p) in RT_p x DG_p, diagnostic (q, F, Φ) in
    CG_{p+1} x RT_p x DG_p.  `f` (Coriolis) and `b` (topography) are functions of the ambient coordinates
    (`cor_cf = f∘ambient_map_cf`); `q0_mode` is "alias" (what the reference's stepper does) or
    "start_of_step"."""

    def __init__(self, model, p_fe, f, b=None, gravity=1.0, degree=None, q0_mode="alias", tau=None):
        self.model, self.p_fe = model, p_fe
        self.degree = 4 * (p_fe + 1) if degree is None else degree
        self.gravity, self.q0_mode, self.tau = float(gravity), q0_mode, tau
        dΩ = Measure(Triangulation(model), self.degree)
        _, xyz = dΩ.quadrature_points(chart=False)
        self.f_q = np.ascontiguousarray(np.broadcast_to(f(xyz), xyz.shape[:2]), dtype=np.float64)
        self.b_q = None if b is None else np.ascontiguousarray(np.broadcast_to(b(xyz), xyz.shape[:2]), dtype=np.float64)


class ThermalShallowWaterOperator(ShallowWaterOperator):
    """DAEFEOperator(opT, opFE, ls_diag) of test/Tutorials/ThermalShallowWater.jl:170-252: prognostic (u, p, B) in
    RT_p x DG_p x DG_p, diagnostic (q, F, Φ, b) in CG_{p+1} x RT_p x DG_p x DG_p; Measure(Ω, 4*order) and Measure(Λ, 4*order)."""

    def __init__(self, model, p_fe, f, degree=None, q0_mode="alias", tau=None):
        degree = (4 * p_fe if p_fe > 0 else 2) if degree is None else degree
        super().__init__(model, p_fe, f, gravity=0.0, degree=degree, q0_mode=q0_mode, tau=tau)


class BoussinesqOperator:
    """TransientSemilinearFEOperator(mass, res, (jac, jac_t), X, Y; constant_mass=true) of test/Tutorials/LinearBoussinesq.jl:112-134
    on AtlasDiscreteModel(CubedSphereWithThicknessMesh(r, t), l): RT_order x DG_order x DG_order, rotation rate `omega(xyz)`."""

    def __init__(self, model, order, omega, c=1.0, N=1.48, degree=None):
        self.model, self.p_fe = model, order
        self.degree = 4 * (order + 1) if degree is None else degree
        dΩ = Measure(Triangulation(model), self.degree)
        _, xyz = dΩ.quadrature_points(chart=False)
        self.omega_q = np.ascontiguousarray(np.broadcast_to(omega(xyz), xyz.shape[:2]), dtype=np.float64)
        self.c, self.N = float(c), float(N)


class AdvectionOperator:
    """TransientSemilinearFEOperator(mass, res, (jac, jac_t), P, Q) of the DG upwind advection tutorial
    (test/Tutorials/AdvectionUpwinding.jl:107-137): Q = DG-Q_order, wind `beta(xyz)` given in ambient components."""

    def __init__(self, model, order, beta, degree=None):
        self.model, self.p_fe = model, order
        self.degree = 4 * order if degree is None else degree
        dΩ = Measure(Triangulation(model), self.degree)
        _, xyz = dΩ.quadrature_points(chart=False)
        self.beta_cells = np.ascontiguousarray(np.broadcast_to(beta(xyz), xyz.shape), dtype=np.float64)
        n = C.c_int32()
        _lib.check(model.lib.gg_skeleton_points(model.h, self.degree, C.byref(n), None))
        fx = np.empty((model.num_cells, 4, n.value, 3))
        _lib.check(model.lib.gg_skeleton_points(model.h, self.degree, C.byref(n), _dp(fx)))
        self.beta_facets = np.ascontiguousarray(np.broadcast_to(beta(fx), fx.shape), dtype=np.float64)


def assemble_advection_matrix(op, Q, A=None):
    """jac(t,u,du,v) = a_Ω(du,v) + b_Ω(du,v) + s_Ω(du,v) of test/Tutorials/AdvectionUpwinding.jl:124-131 as a device CSR matrix
    (volume + skeleton terms; the tutorial hands it to SDIRK + LU)."""
    A = A if A is not None else CSRMatrix(Q, Q, skeleton=True)
    _lib.check(Q.lib.gg_assemble_advection(Q.h, op.degree, _dp(op.beta_cells), _dp(op.beta_facets), A.h))
    return A


class RKStepper:
    """Device-resident explicit RK march (`gg_rk_*`): Gridap.ODEs `ode_march!` for the wave operator, the DAE-aware
    `ode_march!` of src/ODEs/RungeKuttaEX.jl:48-134 for the shallow-water operator."""

    def __init__(self, solver, op):
        m = op.model
        self.lib, self.ctx = m.lib, m.ctx
        A, b, c = TABLEAUS[solver.tableau]
        s = len(b)
        self._A = (C.c_double * (s * s))(*[float(v) for row in A for v in row])
        self._b = (C.c_double * s)(*[float(v) for v in b])
        self._c = (C.c_double * s)(*[float(v) for v in c])
        a = _lib.RKArgs()
        a.order, a.degree, a.n_stages = op.p_fe, op.degree, s
        a.A, a.b, a.c = self._A, self._b, self._c
        a.dt = solver.dt
        a.rtol = getattr(solver.ls, "rtol", 1e-12) or 1e-12
        if isinstance(op, ShallowWaterOperator):
            a.problem = _lib.PROBLEM_THERMAL_SHALLOW_WATER if isinstance(op, ThermalShallowWaterOperator) else _lib.PROBLEM_SHALLOW_WATER
            a.gravity = op.gravity
            a.tau = -1.0 if op.tau is None else float(op.tau)
            a.q0_mode = {"alias": _lib.Q0_ALIAS, "start_of_step": _lib.Q0_START_OF_STEP}[op.q0_mode]
            a.coriolis = _dp(op.f_q)
            a.topography = _dp(op.b_q) if op.b_q is not None else None
        elif isinstance(op, BoussinesqOperator):
            a.problem = _lib.PROBLEM_BOUSSINESQ
            a.coriolis = _dp(op.omega_q)
            a.sound_speed, a.buoyancy_frequency = op.c, op.N
        elif isinstance(op, AdvectionOperator):
            a.problem = _lib.PROBLEM_ADVECTION
            a.velocity_cells, a.velocity_facets = _dp(op.beta_cells), _dp(op.beta_facets)
        else:
            a.problem = _lib.PROBLEM_WAVE
        h = C.c_void_p()
        _lib.check(self.lib.gg_rk_create(m.h, C.byref(a), C.byref(h)))
        self.h = h
        self._model = m
        nu, np_ = C.c_int64(), C.c_int64()
        _lib.check(self.lib.gg_rk_info(self.h, C.byref(nu), C.byref(np_)))
        self.nu, self.np_ = nu.value, np_.value
        nq, nF, nP = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.gg_rk_diag_info(self.h, C.byref(nq), C.byref(nF), C.byref(nP)))
        self.nq, self.ny = nq.value, nq.value + nF.value + nP.value
        self.dt = solver.dt

    def set_state(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.nu + self.np_
        _lib.check(self.lib.gg_rk_set_state(self.h, _dp(x)))

    def get_state(self, out=None):
        out = np.empty(self.nu + self.np_) if out is None else out
        _lib.check(self.lib.gg_rk_get_state(self.h, _dp(out)))
        return out

    def step(self, n=1):
        _lib.check(self.lib.gg_rk_step(self.h, int(n)))

    def residual(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.empty_like(x)
        _lib.check(self.lib.gg_rk_residual(self.h, _dp(x), _dp(r)))
        return r

    # -- shallow water only
    def get_diagnostics(self):
        y = np.empty(self.ny)
        _lib.check(self.lib.gg_rk_get_diagnostics(self.h, _dp(y)))
        return y

    def diagnostics(self, x):
        """y = (q, F, Φ) from the diagnostic solve at x (src/ODEs/DAE.jl:241-274)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.ny)
        _lib.check(self.lib.gg_rk_diagnostics(self.h, _dp(x), _dp(y)))
        return y

    def stage_residual(self, x, y, q0=None):
        """res_x(t,(u,p),(q,F,Φ),(v,r),(q0,..)) + mass(0) as the vector [r_u; r_p]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        q0 = np.ascontiguousarray(y[: self.nq] if q0 is None else q0, dtype=np.float64)
        r = np.empty_like(x)
        _lib.check(self.lib.gg_rk_swe_residual(self.h, _dp(x), _dp(y), _dp(q0), _dp(r)))
        return r

    def stats(self):
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self.lib.gg_rk_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def solver_phase_ns(self, which=0):
        """(cell phase, gather+dot phase, update+preconditioner phase, iterations) of the matrix-free PCG `which`."""
        out = np.zeros(4)
        _lib.check(self.lib.gg_rk_solver_phase_ns(self.h, int(which), _dp(out)))
        return out

    @property
    def state_device_ptr(self):
        return self.lib.gg_rk_state_device_ptr(self.h)

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.lib.gg_rk_destroy(self.h)
        except Exception:
            pass


WaveStepper = RKStepper


def gridap_num_steps(t0, tF, dt):
    """Steps taken by Gridap's ODE-solution iterator: stop when t >= tF - 100 eps, t accumulated as t += dt."""
    eps = 100 * np.finfo(np.float64).eps
    t, n = t0, 0
    while not (t >= tF - eps):
        t = t + dt
        n += 1
    return n


def solve(solver, op, t0, tF, x0):
    """solve(RungeKutta(...), op, t0, tF, xh0): marches to tF on the device, returns (tF, x_F, stepper)."""
    st = RKStepper(solver, op)
    st.set_state(x0)
    n = gridap_num_steps(t0, tF, solver.dt)
    st.step(n)
    return t0 + n * solver.dt, st.get_state(), st
