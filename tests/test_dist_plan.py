"""Multi-rank host logic of the partitioned path on CPU: two gloo ranks build their partition plans independently
(gg_plan_*: pure host code, no GPU) and check that they fit together -- the linear partition and ghost layer of
src/Distributed/AtlasDistributedDiscreteModels.jl:31-49, unique DOF ownership, matching send / receive lists, and an
emulated `consistent!` (owner -> ghost update) that reproduces the global vector on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from __graft_entry__ import load_package

SPACES = [(2, 0), (2, 1), (2, 2), (0, 1), (0, 3), (1, 2)]   # (kind, order): RT0-2, CG1, CG3, DG2
SPACES_SHELL = [(0, 1), (0, 2), (1, 1), (2, 0), (2, 1)]     # hexahedra: CG1, CG2, DG1, RT0, RT1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, shell=False, strips=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gg = load_package()
        if shell:
            # 3-D shell in p8est order (Morton per tree), equal-count chunks of the space-filling curve
            # (src/Distributed/AtlasOctreeDistributedDiscreteModels.jl:286-294)
            plan = gg.PartitionPlan(n, world, rank, thickness=0.19, ordering=2, n_layers=n)
            nc = 6 * n * n * n
            ranges = [None] * world
            dist.all_gather_object(ranges, (plan.first_owned, plan.last_owned))
            assert ranges[0][0] == 0 and ranges[-1][1] == nc and all(ranges[r][1] == ranges[r + 1][0] for r in range(world - 1))
            counts = [b - a for a, b in ranges]
            assert max(counts) - min(counts) <= 1
            mine = np.arange(plan.first_owned, plan.last_owned)
        elif strips:
            # opt-in panel strips: the same rows of cells on all six panels (include/geosgpu.h, gg_plan_create_strips)
            plan = gg.PartitionPlan(n, world, rank, partition="strips")
            nc = 6 * n * n
            cells = np.arange(nc)
            mine = cells[((cells % (n * n)) * world) // (n * n) == rank]
            assert len(mine) % 6 == 0 and len(np.unique(mine // (n * n))) == 6
        else:
            plan = gg.PartitionPlan(n, world, rank)
            nc = 6 * n * n
            # linear partition of the reference: part(i) = div((i-1)*P, Nc) + 1 (1-based i)
            cells = np.arange(nc)
            mine = cells[(cells * world) // nc == rank]
        assert plan.num_cells_global == nc and plan.num_owned_cells == len(mine)
        assert plan.first_owned == mine[0] and plan.last_owned == mine[-1] + 1
        lc = plan.local_cells()
        assert (np.diff(lc) > 0).all() and set(mine) <= set(lc) and len(lc) > len(mine)
        owned_cells = torch.tensor([plan.num_owned_cells])
        dist.all_reduce(owned_cells)
        assert owned_cells.item() == nc
        for kind, order in (SPACES_SHELL if shell else SPACES):
            sp = plan.space(kind, order)
            l2g, owner = sp["l2g"], sp["owner"]
            assert (np.diff(l2g) > 0).all() and owner.min() >= 0 and owner.max() < world
            # every global DOF has exactly one owner
            tot = torch.tensor([sp["n_owned"]])
            dist.all_reduce(tot)
            assert tot.item() == sp["n_global"], (kind, order)
            flags = torch.zeros(sp["n_global"], dtype=torch.int32)
            flags[torch.from_numpy(l2g[owner == rank])] = 1
            dist.all_reduce(flags)
            assert (flags == 1).all()
            # lists match pairwise: what I send to r is what r expects from me, in the same order, at the ids I was told
            others = [None] * world
            dist.all_gather_object(others, {p: {"recv_g": sp_["recv"].tolist(), "recv_gid": l2g[sp_["recv"]].tolist()}
                                            for p, sp_ in sp["neighbors"].items()})
            for peer, lists in sp["neighbors"].items():
                theirs = others[peer].get(rank, {"recv_g": [], "recv_gid": []})
                assert l2g[lists["send"]].tolist() == theirs["recv_gid"], (kind, order, peer)
                assert lists["send_remote"].tolist() == theirs["recv_g"], (kind, order, peer)
                assert (owner[lists["send"]] == rank).all() and (owner[lists["recv"]] == peer).all()
            ghosts = np.flatnonzero(owner != rank)
            recv_all = np.sort(np.concatenate([v["recv"] for v in sp["neighbors"].values()] or [np.empty(0, np.int32)]))
            assert (recv_all == ghosts).all()
            # emulated consistent!: owned values known, ghosts unknown -> exchange -> equals the global vector
            g = np.sin(0.37 * np.arange(sp["n_global"])) + 2.0
            v = np.where(owner == rank, g[l2g], np.nan)
            reqs, bufs = [], {}
            for peer, lists in sorted(sp["neighbors"].items()):
                if len(lists["send"]):
                    reqs.append(dist.isend(torch.from_numpy(v[lists["send"]].copy()), peer))
                if len(lists["recv"]):
                    bufs[peer] = torch.empty(len(lists["recv"]), dtype=torch.float64)
                    reqs.append(dist.irecv(bufs[peer], peer))
            for r in reqs:
                r.wait()
            for peer, b in bufs.items():
                v[sp["neighbors"][peer]["recv"]] = b.numpy()
            assert (v == g[l2g]).all(), (kind, order)
            assert sp["slot_capacity"] >= sp["n_local"]
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 4), (3, 5)])
def test_partition_plans_of_all_ranks_fit_together(world, n):
    mp.spawn(_worker, args=(world, _free_port(), n), nprocs=world, join=True)


@pytest.mark.parametrize("world,n", [(2, 2), (3, 4)])
def test_shell_plans_in_p8est_order_fit_together(world, n):
    mp.spawn(_worker, args=(world, _free_port(), n, True), nprocs=world, join=True)


@pytest.mark.parametrize("world,n", [(2, 4), (3, 5)])
def test_panel_strip_plans_fit_together(world, n):
    mp.spawn(_worker, args=(world, _free_port(), n, False, True), nprocs=world, join=True)


def test_single_rank_plan_is_the_whole_model():
    gg = load_package()
    plan = gg.PartitionPlan(3, 1, 0)
    assert plan.num_local_cells == plan.num_cells_global == 54 and plan.first_owned == 0
    sp = plan.space(2, 1)
    assert sp["n_owned"] == sp["n_local"] == sp["n_global"] and not sp["neighbors"]
    assert (sp["l2g"] == np.arange(sp["n_global"])).all()
