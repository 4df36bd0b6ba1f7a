"""Real multi-GPU checks (need >= 2 GPUs; the driver's single-GPU `-m gpu` run skips them, `gpurun --gpus 2` runs them).
Two NCCL ranks evaluate a partitioned Hex8 mesh with (a) the NCCL send/recv interface exchange, (b) peer push (the copy
engine ships the interface rows into the owner's double-buffered HALO rows while K1 runs) and (c) K1 peer stores (K1
writes them into the owner's memory itself); owned rows of all must be bitwise equal to a single-GPU evaluation of the
whole mesh, after several evaluations with changing inputs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from ad4sm_b200 import dist as AD, mesh, Elements as El, Materials as Ma
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        coords, conn = mesh.hex_block(12, 10, 16, jitter=0.1, seed=31)
        b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
        u = mesh.random_u(len(coords), 3, 1 / 16, seed=32)
        p = AD.build_partition(conn, AD.contiguous_elem_ranks(len(conn), world), rank, world)
        res = {}
        for mode in ("nccl", "push", "peerk1"):
            own = El.ElemBatch("CE", 3, p.local_conn(conn[p.own_elems]), b.gradN[p.own_elems], b.wgt[p.own_elems], b.mat)
            be = AD.GpuBackend(own, p.local_conn(conn[p.halo_elems]), p, 3, device=rank)
            # no explicit be.set_stream: the backend binds the plan to torch's current stream by itself, also when that
            # changes AFTER construction (the NCCL collectives order on it; ADVICE round 1: a silent race otherwise)
            stream = torch.cuda.Stream()
            torch.cuda.set_stream(stream)
            if mode != "nccl":
                assert be.enable_peer_staging(p, push=(mode == "push")), "peer staging (CUDA IPC) not available"
                assert be.push == (mode == "push")
            dp = AD.DistPlan(be, p)
            ul = torch.from_numpy(np.ascontiguousarray(u[:, p.local_nodes - 1].T)).cuda()
            # evaluations with other inputs first: a stale or wrong-parity HALO row would show in the last one.  The push
            # mode alternates two receive buffers: 2 evaluations end on buffer B, 5 on buffer A.
            for tag, seq in ((mode, (0.5, 1.0)), (mode + "_b", (0.25, 0.5, 1.0))):
                for f in seq:
                    phi = dp.step(ul * f)
                torch.cuda.synchronize()
                Kt = be.plan.fetch_csc()
                res[tag] = (float(phi.item()), be.r_tensor().cpu().numpy(), Kt)
        np.savez(os.path.join(out, f"r{rank}.npz"), local_nodes=p.local_nodes, owned=p.owned,
                 **{f"{m}_{k}": v for m, (ph, r, Kt) in res.items()
                    for k, v in dict(phi=ph, r=r, cp=Kt.colptr, rv=Kt.rowval, nz=Kt.nzval).items()})
    finally:
        dist.destroy_process_group()


def test_two_gpu_partition_nccl_and_peer_staging(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from ad4sm_b200 import mesh, Elements as El, Materials as Ma
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    coords, conn = mesh.hex_block(12, 10, 16, jitter=0.1, seed=31)
    b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
    u = mesh.random_u(len(coords), 3, 1 / 16, seed=32)
    phi0, r0, Kt0 = El.Plan(b, u.shape[1], 3).makeϕrKt(u)
    for g in range(world):
        z = np.load(tmp_path / f"r{g}.npz")
        ln = z["local_nodes"]
        for mode in ("nccl", "nccl_b", "push", "push_b", "peerk1", "peerk1_b"):
            assert abs(float(z[f"{mode}_phi"]) - phi0) <= 1e-13 * abs(phi0)
            cp, rv, nz, r = z[f"{mode}_cp"], z[f"{mode}_rv"], z[f"{mode}_nz"], z[f"{mode}_r"]
            for loc in np.where(z["owned"])[0]:
                gn = ln[loc]
                for c in range(3):
                    lc, gc = loc * 3 + c, (gn - 1) * 3 + c
                    assert r[lc] == r0[gc], (mode, g, gn)
                    ls, gs = slice(cp[lc] - 1, cp[lc + 1] - 1), slice(Kt0.colptr[gc] - 1, Kt0.colptr[gc + 1] - 1)
                    assert np.array_equal(nz[ls], Kt0.nzval[gs]), (mode, g, gn)


# ---- the exchange inside the library (ad4sm_dist_*): what a host without torch.distributed calls ---------------------
def _checkerboard(nx, ny, nz, world):
    """2x2x2-element bricks dealt to the ranks like a checkerboard: every rank has several neighbours-in-pieces, send
    lists are non-contiguous, and elements at brick corners are HALO on the other rank through different nodes."""
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    r = ((i // 2 + j // 2 + k // 2) % world).transpose(2, 1, 0).reshape(-1)   # element index = i + nx*(j + ny*k)
    return r.astype(np.int64)


def _lib_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from ad4sm_b200 import dist as AD, mesh, Elements as El, Materials as Ma, _capi as capi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # bootstrap only: ships the NCCL id; no torch NCCL at all
    try:
        dims = (8, 6, 8)
        coords, conn = mesh.hex_block(*dims, jitter=0.1, seed=51)
        b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
        u = mesh.random_u(len(coords), 3, 1 / 8, seed=52)
        res = {}
        for name, er in (("slab", AD.contiguous_elem_ranks(len(conn), world)), ("checker", _checkerboard(*dims, world))):
            p = AD.build_partition(conn, er, rank, world)
            own = El.ElemBatch("CE", 3, p.local_conn(conn[p.own_elems]), b.gradN[p.own_elems], b.wgt[p.own_elems], b.mat)
            lp = AD.LibDistPlan(own, p.local_conn(conn[p.halo_elems]), p, 3, AD.torch_share_id, device=rank)
            assert lp.plan.info(capi.INFO_SORTED) == 2
            # slabs: contiguous interface ranges -> the copy-engine push transport set up inside ad4sm_dist_attach;
            # checkerboard: general partition -> packed ncclSend/ncclRecv
            assert lp.plan.info(capi.INFO_DIST_TRANSPORT) == (2 if name == "slab" else 1)
            ul = np.asfortranarray(u[:, p.local_nodes - 1])
            r_host = np.empty(ul.size)
            for f in (0.5, 1.0):                       # host entry point: u in, global ϕ and local r out
                phi = lp.step_host(np.asfortranarray(ul * f), r_host)
            Kt = lp.plan.fetch_csc()
            # device entry point must give the same bits
            ud = torch.from_numpy(np.ascontiguousarray(ul.T)).cuda()
            torch.cuda.synchronize()
            lp.step_device(ud.data_ptr())
            lp.plan.sync()
            Kt2 = lp.plan.fetch_csc()
            assert np.array_equal(Kt.nzval, Kt2.nzval)
            # ϕ, r only (makeϕr) through the same plan: not a push evaluation -> send/recv of the element-major rows
            capi.check(capi.lib().ad4sm_dist_eval_device(lp.plan._h, capi.MODE_PHI_R, ud.data_ptr(), None, None))
            lp.plan.sync()
            v = lp.plan.device_view()
            r3 = AD._wrap(torch, v.r_dev, (v.n_dof,)).cpu().numpy()
            own3 = np.repeat(p.owned, 3)
            assert np.array_equal(r3[own3], r_host[own3])
            res[name] = (phi, r_host.copy(), Kt, p)
            lp.close()
        np.savez(os.path.join(out, f"lib{rank}.npz"),
                 **{f"{m}_{k}": v for m, (ph, r, Kt, p) in res.items()
                    for k, v in dict(phi=ph, r=r, cp=Kt.colptr, rv=Kt.rowval, nz=Kt.nzval, ln=p.local_nodes, owned=p.owned).items()})
    finally:
        dist.destroy_process_group()


def test_two_gpu_partition_inside_the_library(tmp_path):
    """ad4sm_dist_init / attach / eval on 2 GPUs: a slab partition and a checkerboard (general, non-contiguous) partition
    must both reproduce the single-GPU Kt and r bit for bit on owned rows."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from ad4sm_b200 import mesh, Elements as El, Materials as Ma
    world = 2
    mp.spawn(_lib_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    coords, conn = mesh.hex_block(8, 6, 8, jitter=0.1, seed=51)
    b = El.Hex08_batch(conn, coords, Ma.NeoHooke(1000.0, 5000.0))
    u = mesh.random_u(len(coords), 3, 1 / 8, seed=52)
    phi0, r0, Kt0 = El.Plan(b, u.shape[1], 3).makeϕrKt(u)
    for name in ("slab", "checker"):
        seen = np.zeros(u.shape[1], dtype=int)
        for g in range(world):
            z = np.load(tmp_path / f"lib{g}.npz")
            ln = z[f"{name}_ln"]
            assert abs(float(z[f"{name}_phi"]) - phi0) <= 1e-13 * abs(phi0)
            cp, nz, r = z[f"{name}_cp"], z[f"{name}_nz"], z[f"{name}_r"]
            for loc in np.where(z[f"{name}_owned"])[0]:
                gn = ln[loc]
                seen[gn - 1] += 1
                for c in range(3):
                    lc, gc = loc * 3 + c, (gn - 1) * 3 + c
                    assert r[lc] == r0[gc], (name, g, gn)
                    ls, gs = slice(cp[lc] - 1, cp[lc + 1] - 1), slice(Kt0.colptr[gc] - 1, Kt0.colptr[gc + 1] - 1)
                    assert np.array_equal(nz[ls], Kt0.nzval[gs]), (name, g, gn)
        assert np.all(seen == 1), "every node is owned by exactly one rank"
