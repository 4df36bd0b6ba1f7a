"""Synthetic structured meshes for the benchmark / parity configurations.

The reference has no mesh generator (user scripts build `Vector{CElem}` by hand, README.md:118-167);
these definitions are ours.  Nodes are numbered x-fastest, then y, then z, 1-based; elements x-fastest.
Coordinates are h·(i,j,k) with h = 1/n, plus an optional seeded jitter U(-jitter·h, jitter·h) on
interior nodes (parity runs use jitter so no stiffness entry cancels exactly)."""
from __future__ import annotations

import itertools

import numpy as np


def _jitter(coords, interior, h, jitter, seed):
    if jitter:
        rng = np.random.default_rng(seed)
        coords = coords + interior[:, None] * rng.uniform(-jitter * h, jitter * h, coords.shape)
    return coords


def quad_grid(nx, ny, jitter=0.0, seed=0):
    """(coords [nn,2], conn [ne,4] 1-based, CCW node order of `Elements.Quad`, elasticelements.jl:57)."""
    h = 1.0 / max(nx, ny)
    I, J = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), indexing="ij")
    ids = (I + (nx + 1) * J)  # [i][j]
    coords = np.empty(((nx + 1) * (ny + 1), 2))
    coords[ids.ravel(), 0] = (I * h).ravel()
    coords[ids.ravel(), 1] = (J * h).ravel()
    interior = np.zeros(len(coords))
    interior[ids[1:-1, 1:-1].ravel()] = 1.0
    coords = _jitter(coords, interior, h, jitter, seed)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    conn = np.stack([ids[i, j], ids[i + 1, j], ids[i + 1, j + 1], ids[i, j + 1]], axis=-1)  # [i][j][4]
    conn = conn.transpose(1, 0, 2).reshape(-1, 4) + 1  # element index = i + nx*j
    return coords, conn.astype(np.int64)


def hex_block(nx, ny, nz, jitter=0.0, seed=0):
    """(coords [nn,3], conn [ne,8] 1-based, node order of `Elements.Hex08`, elasticelements.jl:134-137)."""
    h = 1.0 / max(nx, ny, nz)
    I, J, K = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    ids = I + (nx + 1) * (J + (ny + 1) * K)
    coords = np.empty(((nx + 1) * (ny + 1) * (nz + 1), 3))
    coords[ids.ravel(), 0] = (I * h).ravel()
    coords[ids.ravel(), 1] = (J * h).ravel()
    coords[ids.ravel(), 2] = (K * h).ravel()
    interior = np.zeros(len(coords))
    interior[ids[1:-1, 1:-1, 1:-1].ravel()] = 1.0
    coords = _jitter(coords, interior, h, jitter, seed)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    conn = np.stack([ids[i, j, k], ids[i + 1, j, k], ids[i + 1, j + 1, k], ids[i, j + 1, k],
                     ids[i, j, k + 1], ids[i + 1, j, k + 1], ids[i + 1, j + 1, k + 1], ids[i, j + 1, k + 1]], axis=-1)
    conn = conn.transpose(2, 1, 0, 3).reshape(-1, 8) + 1  # element index = i + nx*(j + ny*k)
    return coords, conn.astype(np.int64)


def tet_kuhn(nx, ny, nz, jitter=0.0, seed=0):
    """Kuhn (Freudenthal) split of every cube of an nx×ny×nz grid into 6 positively oriented tets
    (`Elements.Tet04` takes wgt = det([1 x y z])/6 signed, elasticelements.jl:92)."""
    coords, _ = hex_block(nx, ny, nz, jitter, seed)
    I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
    base = (I * sx + J * sy + K * sz).transpose(2, 1, 0).reshape(-1)  # cube index x-fastest
    steps = np.array([sx, sy, sz])
    tets = []
    for perm in itertools.permutations(range(3)):
        a = steps[perm[0]]
        b = a + steps[perm[1]]
        c = sx + sy + sz
        # parity of the permutation decides the orientation
        inv = sum(1 for x in range(3) for y in range(x + 1, 3) if perm[x] > perm[y])
        t = np.stack([base, base + a, base + b, base + c], axis=1)
        if inv % 2 == 1:
            t = t[:, [0, 2, 1, 3]]
        tets.append(t)
    conn = np.stack(tets, axis=1).reshape(-1, 4) + 1  # 6 consecutive tets per cube
    return coords, conn.astype(np.int64)


def tri_grid(nx, ny, jitter=0.0, seed=0):
    """Two CCW triangles per quad cell."""
    coords, q = quad_grid(nx, ny, jitter, seed)
    conn = np.stack([q[:, [0, 1, 2]], q[:, [0, 2, 3]]], axis=1).reshape(-1, 3)
    return coords, conn


def random_u(n_nodes, dpn, h, a=0.02, seed=0):
    """u = a·h·U(-1,1), shape (dpn, n_nodes) like the reference's `u::Matrix` (keeps det F > 0)."""
    rng = np.random.default_rng(seed)
    return np.asfortranarray(a * h * rng.uniform(-1.0, 1.0, (dpn, n_nodes)))
