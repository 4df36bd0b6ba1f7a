"""Mesh-partitioned `makeϕrKt` across GPUs (one process per GPU, `torch.distributed` for the plumbing).

The reference has no multi-device path (its only parallelism on this path is `Threads.@threads` over elements,
/root/reference/src/elasticelements.jl:216); this module is the B200 answer to north_star's "mesh partitioned across
the 8 B200s, each GPU owns a block of CSR rows".

Partition (DESIGN.md "Multi-GPU"):
  * every element belongs to exactly one rank (contiguous ranges of `elems`; z-slabs for the structured configs) and
    is evaluated there, once;
  * a node -- i.e. its `dofs_per_node` CSR rows of Kt and entries of r -- is owned by the LOWEST rank that has an
    element touching it;
  * a rank's plan holds its own elements plus, as a HALO batch (AD4SM_BATCH_HALO), the other ranks' elements that
    touch its owned nodes.  HALO elements are not evaluated locally: after the element stage (K1) the ranks that
    evaluated them send their *staged element blocks and residuals* (the same arrays K1 writes for local elements)
    with NCCL send/recv over NVLink, straight into the receiver's staging slots; then every rank assembles (K2/K3).
    Because the owner sums exactly the contributions, in exactly the `elems` order, that a single GPU would, the
    distributed Kt and r are BITWISE identical to the single-GPU result -- not merely within 1e-12.
  * ϕ: every rank reduces the energies of its own elements (HALO slots hold 0); the partials are all-gathered and
    summed in rank order, so ϕ is run-to-run reproducible (it differs from the single-GPU fixed-tree sum by rounding).

Message sizes (Hex8, n x n interface): n² elements x (324 + 24) doubles = 27.8 MB at n = 100 -- about 36 µs at the
measured 770 GB/s/direction -- against ~3 ms of element + assembly work per 1M-element shard.

The transport is abstracted (`Backend`) so that the partition / exchange logic is exercised on CPU with `gloo` and the
oracle standing in for the kernels (tests/test_dist_cpu.py), and on one GPU with two in-process plans
(tests/test_gpu_parity.py::test_two_rank_halo_exchange_single_gpu).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


# ------------------------------------------------------------------------------------------------
# partition
# ------------------------------------------------------------------------------------------------
@dataclass
class Partition:
    rank: int
    world: int
    own_elems: np.ndarray      # global element ids (0-based, ascending) evaluated here
    halo_elems: np.ndarray     # global ids of the other ranks' elements touching owned nodes, sorted by (owner, id)
    halo_owner: np.ndarray     # owner rank of each halo element
    local_nodes: np.ndarray    # global node ids (1-based, ascending) of all local elements
    owned: np.ndarray          # bool per local node: its rows are complete on this rank
    send: dict = field(default_factory=dict)   # peer -> positions in own_elems, in the peer's halo order
    recv: dict = field(default_factory=dict)   # peer -> (start, count) inside the halo batch

    @property
    def n_local_nodes(self):
        return len(self.local_nodes)

    def local_conn(self, conn_global_rows):
        """global 1-based node ids -> local 1-based ids"""
        return (np.searchsorted(self.local_nodes, conn_global_rows) + 1).astype(np.int64)


def contiguous_elem_ranks(ne, world):
    """rank of each element for the default partition: contiguous, equal ranges of `elems`."""
    bounds = (np.arange(world + 1) * ne) // world
    r = np.empty(ne, dtype=np.int64)
    for g in range(world):
        r[bounds[g]:bounds[g + 1]] = g
    return r


def build_partition(conn, elem_rank, rank, world, n_nodes=None):
    """Partition from the GLOBAL connectivity (`conn` [ne, N], 1-based) and the element -> rank map.  Every rank
    computes the same global ownership, so send and receive lists agree without communication."""
    conn = np.asarray(conn, dtype=np.int64)
    elem_rank = np.asarray(elem_rank, dtype=np.int64)
    ne, N = conn.shape
    n_nodes = int(conn.max()) if n_nodes is None else n_nodes
    node_owner = np.full(n_nodes + 1, world, dtype=np.int64)
    np.minimum.at(node_owner, conn.reshape(-1), np.repeat(elem_rank, N))
    touches = node_owner[conn]                                   # [ne, N] owner of each node of each element
    parts = []
    for g in range(world):
        halo_mask = (elem_rank != g) & (touches == g).any(axis=1)
        h = np.where(halo_mask)[0]
        order = np.lexsort((h, elem_rank[h]))
        parts.append(h[order])
    own = np.where(elem_rank == rank)[0]
    halo = parts[rank]
    howner = elem_rank[halo]
    local_nodes = np.unique(np.concatenate([conn[own].reshape(-1), conn[halo].reshape(-1)]))
    p = Partition(rank, world, own, halo, howner, local_nodes, node_owner[local_nodes] == rank)
    for peer in range(world):
        if peer == rank:
            continue
        mine_in_peer_halo = parts[peer][elem_rank[parts[peer]] == rank]    # ascending global id
        if len(mine_in_peer_halo):
            p.send[peer] = np.searchsorted(own, mine_in_peer_halo).astype(np.int64)
        sel = np.where(howner == peer)[0]
        if len(sel):
            p.recv[peer] = (int(sel[0]), int(len(sel)))
    return p


# ------------------------------------------------------------------------------------------------
# backends: who evaluates elements / assembles, and where the staging lives
# ------------------------------------------------------------------------------------------------
class GpuBackend:
    """The product path: a plan of libad4sm_b200 on this rank's GPU."""

    def __init__(self, own_batch, halo_nodes_local, part, dofs_per_node, device=-1, mode=None, **plan_opts):
        import torch
        from . import Elements as El, _capi as capi
        self.capi, self.torch = capi, torch
        self.mode = capi.MODE_ELASTIC if mode is None else mode
        batches = [own_batch]
        own_batch.orig_index = np.ascontiguousarray(part.own_elems, dtype=np.int64)
        if len(part.halo_elems):
            batches.append(El.ElemBatch(own_batch.kind, own_batch.D, halo_nodes_local, None, None, own_batch.mat, None,
                                        orig_index=np.ascontiguousarray(part.halo_elems, dtype=np.int64), halo=True,
                                        halo_P=own_batch.P))
        self.plan = El.Plan(batches, part.n_local_nodes, dofs_per_node, device=device, **plan_opts)
        self.n_own = own_batch.ne
        self.peer = False
        self.push = False
        self._bound = None
        self._phi_t = self._r_t = None
        self.bind_current_stream()
        self._declare_exports(part)
        field_id = capi.FIELD_D if self.mode == capi.MODE_PF_D else capi.FIELD_U
        v = self.plan.staging_view(field_id)
        self.Ke = _wrap(torch, v.Ke_dev, (v.n_slots, v.ke_width))
        self.re = _wrap(torch, v.re_dev, (v.n_slots, v.re_width))
        self.field_id = field_id

    def _declare_exports(self, part):
        """Tell the library which slot ranges this rank reads between the stages (the rows it sends): everything else may
        then use the destination-sorted staging.  Non-contiguous send lists keep the all-element-major staging."""
        import ctypes as C
        rng = []
        for peer, ix in part.send.items():
            if len(ix) == 0 or int(ix[-1]) - int(ix[0]) + 1 != len(ix):
                return
            rng.append((int(ix[0]), len(ix)))
        if len(rng) > 4:
            return
        first = (C.c_int64 * max(1, len(rng)))(*[r[0] for r in rng])
        count = (C.c_int64 * max(1, len(rng)))(*[r[1] for r in rng])
        self.capi.check(self.capi.lib().ad4sm_export_ranges(self.plan._h, len(rng), first, count))

    def set_stream(self, stream):
        self.plan.set_stream(stream.cuda_stream or 1)
        self._bound = stream.cuda_stream

    def bind_current_stream(self):
        """The library's kernels and torch.distributed's collectives must order on ONE stream: K1 -> send/recv or
        barrier -> assembly -> ϕ all-gather.  A plan launches on a private non-blocking stream by default, which NCCL
        (ordering on torch's current stream) would race with; so the backend binds the plan to torch's current stream at
        construction and again whenever a step finds the current stream changed.  (Handle 0 is torch's legacy default
        stream: passed as cudaStreamLegacy = 1, since NULL means "the plan's own stream" in the C ABI.)"""
        cur = self.torch.cuda.current_stream().cuda_stream
        if cur != self._bound:
            self.plan.set_stream(cur or 1)
            self._bound = cur

    def enable_peer_staging(self, part, dist=None, push=True):
        """Interface rows through peer-mapped memory (CUDA IPC over NVLink) instead of NCCL send/recv; the exchange step
        shrinks to a barrier.  push=True (default): K1 stays on the destination-sorted path, evaluates the interface
        elements first and the COPY ENGINE pushes their rows into the owner's (double-buffered) HALO rows while the rest
        of K1 runs.  push=False: K1's bulk-store epilogue writes them into the owner's memory itself (element-major
        path).  Collective: every rank must call it; returns True only if ALL ranks attached their links (otherwise
        everybody keeps the NCCL send/recv exchange)."""
        import ctypes as C
        import torch
        import torch.distributed as D
        dist = dist or D
        capi = self.capi
        L = capi.lib()
        ok, links, keep = 1, [], []
        try:
            h = (C.c_ubyte * capi.PEER_HANDLE_BYTES)()
            capi.check(L.ad4sm_peer_export(self.plan._h, h))
            n_slots = int(self.plan.staging_view(self.field_id).n_slots)
            mine = {"handles": bytes(h), "n_own": self.n_own, "n_slots": n_slots, "recv": dict(part.recv)}
        except Exception:
            ok, mine = 0, {"handles": b"", "n_own": self.n_own, "n_slots": 0, "recv": dict(part.recv)}
        everyone = [None] * part.world
        dist.all_gather_object(everyone, mine)
        if ok:
            try:
                if len(part.send) > 4:
                    raise ValueError("more than AD4SM_MAX_PEER_LINKS peers")
                for peer, ix in part.send.items():
                    if int(ix[-1]) - int(ix[0]) + 1 != len(ix) or not everyone[peer]["handles"]:
                        raise ValueError("peer staging needs contiguous interface ranges")
                    s, c = everyone[peer]["recv"][part.rank]
                    buf = C.create_string_buffer(everyone[peer]["handles"], capi.PEER_HANDLE_BYTES)
                    keep.append(buf)
                    links.append(capi.PeerLink(int(ix[0]), len(ix), everyone[peer]["n_own"] + s, C.cast(buf, C.c_void_p),
                                               everyone[peer]["n_slots"] + s))
                arr = (capi.PeerLink * max(1, len(links)))(*links)
                self.plan.set_option(capi.OPT_PEER_PUSH, 1 if push else 0)
                capi.check(L.ad4sm_peer_attach(self.plan._h, len(links), arr))
            except Exception:
                ok = 0
        flag = torch.tensor([ok], device="cuda", dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        self.peer = bool(int(flag.item()))
        self.push = self.peer and push
        if self.peer and not push:
            # peer stores leave through K1's 2.6 KB-per-element bulk copies (TMA), which exist in the element-major
            # staging path; from the sorted epilogue they would be 8-byte stores over NVLink (measured: slower)
            self.plan.set_option(capi.OPT_SORTED, 0)
        if not self.peer:
            self.plan.set_option(capi.OPT_PEER_PUSH, 0)
            capi.check(L.ad4sm_peer_attach(self.plan._h, 0, None))   # drop any links this rank did attach
            self._declare_exports(part)   # back to the send/recv exchange of locally materialised rows
        return self.peer

    def eval_elements(self, u_dev, d_dev=None, hist_dev=None):
        self.plan.eval_elements_device(self.mode, u_dev.data_ptr(), d_dev.data_ptr() if d_dev is not None else None,
                                       hist_dev.data_ptr() if hist_dev is not None else None)

    def eval_elements_host(self, u_host, d_host=None):
        self.plan.eval_elements_host(self.mode, u_host, d_host)

    def fetch_r_async(self, r_host):
        self.plan.fetch_phi_r_async(r_host, None, self.field_id)

    def staging(self):
        return self.Ke, self.re

    def assemble(self):
        self.plan.eval_assemble_device(self.mode)

    def assemble_interior(self):
        self.plan.eval_assemble_interior_device(self.mode)

    def halo_prepare(self, cuda_stream):
        self.plan.eval_halo_prepare_device(self.mode, cuda_stream)

    def phi_tensor(self):
        if self._phi_t is None:   # library-owned device memory: the views are made once
            v = self.plan.device_view(self.field_id)
            self._phi_t = _wrap(self.torch, v.phi_dev, (1,))
        return self._phi_t

    def r_tensor(self):
        if self._r_t is None:
            v = self.plan.device_view(self.field_id)
            self._r_t = _wrap(self.torch, v.r_dev, (v.n_dof,))
        return self._r_t


class LibDistPlan:
    """Partitioned `makeϕrKt` with the whole exchange INSIDE the library (ad4sm_dist_*: NCCL bound by the library itself,
    no torch.distributed on the data path).  This is the path a Julia / C host uses; Python only bootstraps it:
    `share_id(bytes_or_None) -> bytes` must hand rank 0's 128-byte NCCL id to every rank (MPI bcast, a file, or -- here
    -- torch.distributed.broadcast_object_list)."""

    def __init__(self, own_batch, halo_nodes_local, part, dofs_per_node, share_id, device=-1, mode=None, **plan_opts):
        import ctypes as C
        from . import Elements as El, _capi as capi
        self.capi, self.part = capi, part
        self.mode = capi.MODE_ELASTIC if mode is None else mode
        L = capi.lib()
        batches = [own_batch]
        own_batch.orig_index = np.ascontiguousarray(part.own_elems, dtype=np.int64)
        if len(part.halo_elems):
            batches.append(El.ElemBatch(own_batch.kind, own_batch.D, halo_nodes_local, None, None, own_batch.mat, None,
                                        orig_index=np.ascontiguousarray(part.halo_elems, dtype=np.int64), halo=True,
                                        halo_P=own_batch.P))
        self.plan = El.Plan(batches, part.n_local_nodes, dofs_per_node, device=device, **plan_opts)
        self.n_own = own_batch.ne
        idb = None
        if part.rank == 0:
            buf = (C.c_ubyte * capi.NCCL_ID_BYTES)()
            capi.check(L.ad4sm_dist_unique_id(buf))
            idb = bytes(buf)
        idb = share_id(idb)
        h = C.c_void_p()
        capi.check(L.ad4sm_dist_init(part.world, part.rank, C.create_string_buffer(idb, capi.NCCL_ID_BYTES), device, C.byref(h)))
        self._comm = h
        peers = sorted(set(part.send) | set(part.recv))
        arr = (capi.DistPeer * max(1, len(peers)))()
        self._keep = []
        for i, peer in enumerate(peers):
            ix = np.ascontiguousarray(part.send.get(peer, np.zeros(0, dtype=np.int64)), dtype=np.int64)
            self._keep.append(ix)
            s, c = part.recv.get(peer, (0, 0))
            arr[i].rank, arr[i].n_send, arr[i].send_slots = peer, len(ix), ix.ctypes.data
            arr[i].recv_first_slot, arr[i].n_recv = self.n_own + s, c
        capi.check(L.ad4sm_dist_attach(self.plan._h, h, len(peers), arr))

    def close(self):
        if self._comm is not None:
            self.plan.sync()
            self.capi.lib().ad4sm_dist_destroy(self._comm)
            self._comm = None

    def step_device(self, u_ptr, d_ptr=None, hist_ptr=None):
        """asynchronous on the plan's stream; the global ϕ, this rank's r and Kt rows are then in the plan's device view"""
        import ctypes as C
        self.capi.check(self.capi.lib().ad4sm_dist_eval_device(self.plan._h, self.mode, C.c_void_p(u_ptr),
                                                               C.c_void_p(d_ptr) if d_ptr else None,
                                                               C.c_void_p(hist_ptr) if hist_ptr else None))

    def step_host(self, u_host, r_host, d_host=None):
        """u (column-major D x nLocalNodes, pinned for overlap) in; returns the global ϕ, fills this rank's r"""
        import ctypes as C
        phi = C.c_double()
        self.capi.check(self.capi.lib().ad4sm_dist_eval(self.plan._h, self.mode, C.c_void_p(u_host.ctypes.data),
                                                        C.c_void_p(d_host.ctypes.data) if d_host is not None else None,
                                                        C.byref(phi), C.c_void_p(r_host.ctypes.data)))
        return phi.value


def torch_share_id(idb):
    """bootstrap helper: broadcast rank 0's NCCL id through an initialised torch.distributed group"""
    import torch.distributed as D
    box = [idb]
    D.broadcast_object_list(box, src=0)
    return box[0]


class _DevArr:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _wrap(torch, ptr, shape):
    """torch view (no copy) of library-owned device memory"""
    return torch.as_tensor(_DevArr(ptr, shape), device="cuda")


# ------------------------------------------------------------------------------------------------
# the exchange step and the distributed evaluation
# ------------------------------------------------------------------------------------------------
def exchange_staging(part, n_own, arrays, dist=None):
    """Ship the staged rows of the own elements that are HALO on a peer, and receive this rank's HALO rows in place.
    `arrays`: 2-D tensors [n_slots, width] (Ke, re): rows [0, n_own) own, rows [n_own, ...) halo."""
    if not part.send and not part.recv:
        return
    import torch
    import torch.distributed as D
    dist = dist or D
    ops, keep = [], []
    for peer in sorted(set(part.send) | set(part.recv)):
        for a in arrays:
            if peer in part.send:
                ix = part.send[peer]
                if len(ix) and int(ix[-1]) - int(ix[0]) + 1 == len(ix):
                    buf = a[int(ix[0]):int(ix[0]) + len(ix)]   # contiguous rows (z-slabs): send in place, no pack
                else:
                    buf = a.index_select(0, torch.as_tensor(ix, device=a.device))   # pack on the current stream
                keep.append(buf)
                ops.append(dist.P2POp(dist.isend, buf, peer))
            if peer in part.recv:
                s, c = part.recv[peer]
                ops.append(dist.P2POp(dist.irecv, a[n_own + s:n_own + s + c], peer))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class DistPlan:
    """`makeϕrKt` on a mesh partition: element stage -> interface exchange -> assembly -> ϕ all-gather."""

    def __init__(self, backend, part):
        self.b, self.part = backend, part
        self._tok = None
        self._allphi = self._side = None

    def _prof_events(self):
        """developer tool (AD4SM_DIST_PROF=1): CUDA events after K1 / after the exchange (side stream) / after the interior
        assembly / at the end of every step; `prof_report()` averages the intervals"""
        import os
        import torch
        if not os.environ.get("AD4SM_DIST_PROF"):
            return None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self._prof_log = getattr(self, "_prof_log", [])
        self._prof_log.append(ev)
        return ev

    def prof_report(self, skip=5):
        import torch
        torch.cuda.synchronize()
        log = getattr(self, "_prof_log", [])[skip:]
        if not log:
            return None
        n = len(log)
        return {"steps": n, "k1_end_to_exchange_done_ms": sum(e[0].elapsed_time(e[1]) for e in log) / n,
                "k1_end_to_interior_done_ms": sum(e[0].elapsed_time(e[2]) for e in log) / n,
                "k1_end_to_step_end_ms": sum(e[0].elapsed_time(e[3]) for e in log) / n,
                "cpu_enqueue_ms_per_step": 1e3 * sum(self._prof_cpu[skip:]) / max(1, len(self._prof_cpu[skip:]))}

    def step_host(self, u_host, r_host, d_host=None):
        """One evaluation with host-resident u (in) and r (out, this rank's local dofs): the upload overlaps K1, the
        download of r overlaps the Kt assembly.  Returns the global ϕ as a 0-d tensor; r_host is valid after
        `backend.plan.sync()`."""
        return self.step(host=(u_host, r_host, d_host))

    def step(self, *eval_args, host=None):
        """One evaluation, everything left on the device.  Returns the global ϕ as a 0-d tensor.

        Stream plan (world > 1): the element stage and both halves of the assembly run on torch's current stream; the
        interface exchange (send/recv, or the barrier that follows the copy-engine pushes) and the all-gather of the
        per-rank ϕ -- which doubles as that barrier -- run on a side stream while the main stream assembles the interior
        rows and pairs (`ad4sm_eval_assemble_split_device`); only the interface part of the assembly waits for them."""
        import time
        import torch
        import torch.distributed as D
        t_cpu0 = time.perf_counter()
        if hasattr(self.b, "bind_current_stream"):
            self.b.bind_current_stream()   # no-op unless the caller switched torch streams since the last step
        peer = getattr(self.b, "peer", False)
        if peer and not getattr(self.b, "push", False):
            # K1 itself writes into the owners' (single-buffered) HALO rows: nobody may start before every rank has
            # finished assembling the previous evaluation
            if self._tok is None:
                self._tok = torch.zeros(1, device="cuda")
            D.all_reduce(self._tok)
        if host is not None:
            self.b.eval_elements_host(host[0], host[2])
        else:
            self.b.eval_elements(*eval_args)
        phi = self.b.phi_tensor()   # this rank's partial: reduced at the end of the element stage
        if self.part.world == 1:
            self.b.assemble()
            if host is not None:
                self.b.fetch_r_async(host[1])
            return phi[0]
        split = phi.is_cuda and hasattr(self.b, "assemble_interior")   # (the CPU stand-in of the gloo tests has no streams)
        if self._allphi is None:
            self._allphi = torch.empty(self.part.world, dtype=phi.dtype, device=phi.device)
            # high priority: the collective's few CTAs must not queue behind the thousands of CTAs of the interior
            # assembly launched right after it on the main stream
            self._side = torch.cuda.Stream(priority=-1) if split else None

        def exchange():
            if not peer:
                exchange_staging(self.part, self.b.n_own, self.b.staging())
            # every rank contributes after its element stage and its pushes: completion here means every peer's rows
            # have landed (the barrier of the peer-staging protocols) and carries the ϕ partials at the same time
            D.all_gather_into_tensor(self._allphi, phi.contiguous())

        if split:
            main = torch.cuda.current_stream()
            prof = self._prof_events()
            if prof:
                prof[0].record(main)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                exchange()
                if prof:
                    prof[1].record(self._side)
            self.b.assemble_interior()
            if prof:
                prof[2].record(main)
            # behind the exchange, on its stream: HALO rows into the sorted staging + the interface rows of r
            self.b.halo_prepare(self._side.cuda_stream)
            main.wait_stream(self._side)
        else:
            exchange()
        self.b.assemble()
        if split and prof:
            prof[3].record(main)
            self._prof_cpu = getattr(self, "_prof_cpu", [])
            self._prof_cpu.append(time.perf_counter() - t_cpu0)
        if host is not None:
            self.b.fetch_r_async(host[1])
        return self._allphi.sum()   # one fixed reduction over `world` values: run-to-run reproducible


# ------------------------------------------------------------------------------------------------
# structured z-slab workload for bench.py (weak scaling): built rank-locally, no global arrays
# ------------------------------------------------------------------------------------------------
def slab_hex_workload(n, nz_per_rank, rank, world, mat, seed=20261021, nz_total=None):
    """Rank-local pieces of an n x n x nzt Hex8 block partitioned into z-slabs.  Weak scaling: nzt = nz_per_rank*world
    (every rank owns nz_per_rank layers); strong scaling: `nz_total` layers split as evenly as integers allow (rank g
    owns layers [g*nzt//world, (g+1)*nzt//world)).
    Returns (own ElemBatch with LOCAL node ids, halo connectivity in local ids, Partition, u_local [3, n_local])."""
    from . import Elements as El, mesh
    h = 1.0 / n
    if nz_total is None:
        k0, k1 = rank * nz_per_rank, (rank + 1) * nz_per_rank     # own element layers [k0, k1)
    else:
        k0, k1 = (rank * nz_total) // world, ((rank + 1) * nz_total) // world
        if k1 - k0 < 1:
            raise ValueError("fewer element layers than ranks")
        nz_per_rank = k1 - k0
    # owned nodes: planes k0+1..k1 (plus plane 0 on rank 0); the plane k1 is shared with rank+1 and owned here (lower
    # rank), so the halo is the first element layer of rank+1, and this rank's first layer is halo on rank-1.
    has_halo = rank + 1 < world
    kz0, kz1 = k0, k1 + (1 if has_halo else 0)                   # local element layers
    npl = (n + 1) * (n + 1)
    node_lo = kz0 * npl                                           # first global node (0-based) of the local range
    n_local_nodes = (kz1 - kz0 + 1) * npl
    I, J, K = np.meshgrid(np.arange(n + 1), np.arange(n + 1), np.arange(kz0, kz1 + 1), indexing="ij")
    ids = I + (n + 1) * (J + (n + 1) * (K - kz0))                 # local 0-based ids, x fastest
    coords = np.empty((n_local_nodes, 3))
    coords[ids.ravel(), 0] = (I * h).ravel()
    coords[ids.ravel(), 1] = (J * h).ravel()
    coords[ids.ravel(), 2] = (K * h).ravel()
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(kz1 - kz0), indexing="ij")
    conn = np.stack([ids[i, j, k], ids[i + 1, j, k], ids[i + 1, j + 1, k], ids[i, j + 1, k],
                     ids[i, j, k + 1], ids[i + 1, j, k + 1], ids[i + 1, j + 1, k + 1], ids[i, j + 1, k + 1]], axis=-1)
    conn = conn.transpose(2, 1, 0, 3).reshape(-1, 8) + 1          # local element index = i + n*(j + n*k_local)
    ne_own = n * n * nz_per_rank
    own_batch = El.Hex08_batch(conn[:ne_own], coords, mat)
    halo_conn = conn[ne_own:]
    g0 = n * n * k0                                               # global id of the first own element
    own_elems = g0 + np.arange(ne_own, dtype=np.int64)
    halo_elems = g0 + ne_own + np.arange(len(halo_conn), dtype=np.int64)
    local_nodes = node_lo + 1 + np.arange(n_local_nodes, dtype=np.int64)
    node_plane = np.arange(n_local_nodes) // npl + kz0
    owned = (node_plane > k0) | (rank == 0)
    owned &= node_plane <= k1
    part = Partition(rank, world, own_elems, halo_elems, np.full(len(halo_elems), rank + 1, dtype=np.int64), local_nodes, owned)
    if has_halo:
        part.recv[rank + 1] = (0, len(halo_elems))
    if rank > 0:
        part.send[rank - 1] = np.arange(n * n, dtype=np.int64)   # this rank's first element layer
    rng = np.random.default_rng(seed + 2)
    # u is defined per GLOBAL node so that neighbouring ranks agree on the shared planes: draw plane by plane
    u = np.empty((3, n_local_nodes))
    for kk in range(kz0, kz1 + 1):
        pr = np.random.default_rng([seed + 2, kk])
        u[:, (kk - kz0) * npl:(kk - kz0 + 1) * npl] = 0.02 * h * pr.uniform(-1.0, 1.0, (3, npl))
    del rng
    return own_batch, halo_conn, part, np.asfortranarray(u)


class SlabRunner:
    """bench.py's multi-GPU arm: config C2 weak-scaled -- every rank owns an n x n x nz Hex8 NeoHooke slab of an
    n x n x (nz*world) block."""

    def __init__(self, n, nz_per_rank, rank, world, device=0, mat=None, peer_staging=False, nz_total=None):
        import torch
        from . import Materials as Ma
        self.torch = torch
        mat = mat or Ma.NeoHooke(1000.0, 5000.0)
        self.batch, halo_conn, self.part, self.u_local = slab_hex_workload(n, nz_per_rank, rank, world, mat, nz_total=nz_total)
        self.lib = None
        if peer_staging == "lib":   # the exchange runs inside the library (ad4sm_dist_*), torch only ships the NCCL id
            self.lib = LibDistPlan(self.batch, halo_conn, self.part, 3, torch_share_id, device=device)
            self.plan, self.peer = self.lib.plan, False
            self.backend = None
            return
        self.backend = GpuBackend(self.batch, halo_conn, self.part, 3, device=device)
        self.plan = self.backend.plan
        self.dplan = DistPlan(self.backend, self.part)
        self._u_dev = None
        self.peer = False
        if world > 1 and peer_staging:   # "push" (default for True) or "k1"
            self.peer = self.backend.enable_peer_staging(self.part, push=(peer_staging != "k1"))

    def step_device(self, u_dev):
        if self.lib is not None:
            return self.lib.step_device(u_dev.data_ptr())
        return self.dplan.step(u_dev)

    def step_host(self, u_host_np, r_host_np):
        """end to end: u from pinned host memory -> device, evaluation + exchange, ϕ and this rank's r back to the host"""
        if self.lib is not None:
            return self.lib.step_host(u_host_np, r_host_np)   # one C call; it returns when ϕ and r are on the host
        phi = self.dplan.step_host(u_host_np, r_host_np)
        val = float(phi.item())   # synchronises the main stream
        self.plan.sync()          # and the copy stream that carries r
        return val
