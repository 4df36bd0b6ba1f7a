"""Multi-rank host logic on CPU: world_size-2 and -3 `gloo` runs of ad4sm_b200.dist with the ORACLE standing in for
the CUDA kernels (per-element duals = what K1 stages; oracle assembly = what K2/K3 do).  Checks the partition
(ownership, halo, send/recv lists), the staging exchange and the ϕ reduction against a single-process evaluation of
the whole mesh: owned rows must be bitwise identical to the corresponding rows of the global result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ad4sm_b200 import dist as AD, mesh
from ad4sm_b200 import Elements as El, Materials as Ma
from oracle import cxx_binding as X


def _mesh():
    coords, conn = mesh.hex_block(3, 2, 5, jitter=0.1, seed=3)
    b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
    u = mesh.random_u(len(coords), 3, 1 / 5, seed=11)
    return b, conn, u


class OracleBackend:
    """CPU stand-in for GpuBackend: staging rows = (vectorised Hessian, gradient) of each element dual."""

    def __init__(self, b, conn, part, u_global):
        self.part = part
        self.n_own = len(part.own_elems)
        nloc = len(part.own_elems) + len(part.halo_elems)
        self.Ke = torch.zeros(nloc, 24 * 24, dtype=torch.float64)
        self.re = torch.zeros(nloc, 24, dtype=torch.float64)
        self.val = np.zeros(nloc)
        self.b, self.conn = b, conn
        self.u_local = np.asfortranarray(u_global[:, part.local_nodes - 1])
        allel = np.concatenate([part.own_elems, part.halo_elems])
        self.local_conn = part.local_conn(conn[allel])
        self.order = np.argsort(allel, kind="stable")           # assembly in global `elems` order

    def eval_elements(self):
        own = self.part.own_elems
        lc = self.local_conn[:self.n_own]
        val, g, H = X.elem_duals("CE", 3, self.b.P, 8, self.b.mat, lc, self.b.gradN[own], self.b.wgt[own], None, X.M_ELASTIC,
                                 self.u_local)
        self.Ke[:self.n_own] = torch.from_numpy(H.reshape(len(own), -1))
        self.re[:self.n_own] = torch.from_numpy(g)
        self.val[:self.n_own] = val

    def staging(self):
        return self.Ke, self.re

    def assemble(self):
        o = self.order
        H = self.Ke.numpy().reshape(-1, 24, 24)[o]
        self.result = X.assemble(self.local_conn[o], 3, self.val[o], self.re.numpy()[o], H, 3 * self.part.n_local_nodes)

    def phi_tensor(self):
        return torch.tensor([self.val[:self.n_own].sum()], dtype=torch.float64)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, conn, u = _mesh()
        part = AD.build_partition(conn, AD.contiguous_elem_ranks(len(conn), world), rank, world)
        be = OracleBackend(b, conn, part, u)
        # the product's step logic (element stage -> exchange + ϕ all-gather -> assembly) with the oracle as kernel stand-in
        phi_total = float(AD.DistPlan(be, part).step())
        phi_r, r, (cp, rv, nz) = be.result
        np.savez(os.path.join(out, f"rank{rank}.npz"), r=r, cp=cp, rv=rv, nz=nz, local_nodes=part.local_nodes, owned=part.owned,
                 phi=phi_total, n_halo=len(part.halo_elems),
                 n_send=sum(len(v) for v in part.send.values()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_assembly_matches_global(world, tmp_path):
    b, conn, u = _mesh()
    phi0, r0, (cp0, rv0, nz0) = X.make_phi_r_Kt([b], u)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    seen = np.zeros(u.shape[1], dtype=int)
    tot_halo = tot_send = 0
    for g in range(world):
        z = np.load(tmp_path / f"rank{g}.npz")
        ln, owned = z["local_nodes"], z["owned"]
        tot_halo += int(z["n_halo"])
        tot_send += int(z["n_send"])
        assert abs(float(z["phi"]) - phi0) <= 1e-13 * abs(phi0)
        for loc in np.where(owned)[0]:
            gn = ln[loc]
            seen[gn - 1] += 1
            for c in range(3):
                lcol, gcol = loc * 3 + c, (gn - 1) * 3 + c
                assert z["r"][lcol] == r0[gcol]
                ls, ge = slice(z["cp"][lcol] - 1, z["cp"][lcol + 1] - 1), slice(cp0[gcol] - 1, cp0[gcol + 1] - 1)
                lrows = z["rv"][ls] - 1
                grows = (ln[lrows // 3] - 1) * 3 + lrows % 3 + 1
                assert np.array_equal(grows, rv0[ge]), "row pattern of an owned node differs from the global one"
                assert np.array_equal(z["nz"][ls], nz0[ge]), "owned rows must be bitwise identical to the single-process result"
    assert np.all(seen == 1), "every node must be owned by exactly one rank"
    assert tot_halo == tot_send > 0


def test_partition_lists_are_consistent():
    _, conn, _ = _mesh()
    world = 3
    er = AD.contiguous_elem_ranks(len(conn), world)
    parts = [AD.build_partition(conn, er, g, world) for g in range(world)]
    for g, p in enumerate(parts):
        for peer, idx in p.send.items():
            s, c = parts[peer].recv[g]
            assert c == len(idx)
            assert np.array_equal(p.own_elems[idx], parts[peer].halo_elems[s:s + c])
        assert set(p.own_elems).isdisjoint(p.halo_elems)


def test_slab_workload_matches_generic_partition():
    """bench.py's rank-local slab builder == build_partition on the global mesh"""
    n, nz, world = 3, 2, 3
    coords, conn = mesh.hex_block(n, n, nz * world)
    er = AD.contiguous_elem_ranks(len(conn), world)
    for g in range(world):
        p = AD.build_partition(conn, er, g, world)
        ob, halo_conn, q, u = AD.slab_hex_workload(n, nz, g, world, Ma.NeoHooke(1000.0, 5000.0))
        assert np.array_equal(p.own_elems, q.own_elems) and np.array_equal(p.halo_elems, q.halo_elems)
        assert np.array_equal(p.local_nodes, q.local_nodes) and np.array_equal(p.owned, q.owned)
        assert np.array_equal(p.local_conn(conn[p.own_elems]), ob.nodes)
        assert np.array_equal(p.local_conn(conn[p.halo_elems]), halo_conn)
        assert set(p.send) == set(q.send) and set(p.recv) == set(q.recv)
        for k in p.send:
            assert np.array_equal(p.send[k], q.send[k])
        for k in p.recv:
            assert p.recv[k] == q.recv[k]
        assert u.shape == (3, len(q.local_nodes))
