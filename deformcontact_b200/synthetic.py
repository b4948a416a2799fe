"""Deterministic synthetic inputs of SURVEY.md section 8(d) / BASELINE.md section 2 (product side).

Same per-graph seeding (``1234 + g``), draw order and shapes as the oracle's generator
(tests/test_synthetic.py checks they agree), but edges and features are produced by the GPU
kernels (kNN bit-identical to the oracle by tests/test_gpu_kernels.py).  Soft graph:
``pos ~ U[-0.5,0.5)^3``, 21-d log-frequency features, kNN-k edges.  Collider: UV sphere of
Open3D's ``create_sphere(radius)`` topology (loaders/common.py:26; 762 vertices / 1520
triangles / 4560 directed edges) centred on a random soft vertex, 25-d features
``[force_dir(3) | force/force_max(1) | posenc(21)]`` (loaders/common.py:6-19).
"""
import functools
import math

import torch

from . import ops
from .data import Data, Batch

FORCE_MAX = 10000.0   # configs/everyday.json:10
SPHERE_R = 0.05       # configs/everyday.json:9


@functools.lru_cache(maxsize=4)
def uv_sphere(radius=SPHERE_R, resolution=20):
    """Open3D TriangleMesh::CreateSphere topology: (vertices fp64 [2r(r-1)+2, 3], triangles int64)."""
    res, step = resolution, math.pi / resolution
    V = [(0.0, 0.0, radius), (0.0, 0.0, -radius)]
    for i in range(1, res):
        a = step * i
        for j in range(2 * res):
            t = step * j
            V.append((math.sin(a) * math.cos(t) * radius, math.sin(a) * math.sin(t) * radius, math.cos(a) * radius))
    j = torch.arange(2 * res)
    j1 = (j + 1) % (2 * res)
    last = 2 + 2 * res * (res - 2)
    caps = torch.stack([torch.stack([torch.zeros_like(j), 2 + j, 2 + j1], 1),
                        torch.stack([torch.ones_like(j), last + j1, last + j], 1)], 1).reshape(-1, 3)
    mids = []
    for i in range(1, res - 1):
        b1 = 2 + 2 * res * (i - 1)
        b2 = b1 + 2 * res
        mids.append(torch.stack([torch.stack([b2 + j, b1 + j1, b1 + j], 1),
                                 torch.stack([b2 + j, b2 + j1, b1 + j1], 1)], 1).reshape(-1, 3))
    return torch.tensor(V, dtype=torch.float64), torch.cat([caps] + mids, 0)


def make_batch(n_graphs, n_nodes=2000, k=8, first=0, device="cuda"):
    """-> (soft_rest, rigid, soft_deformed) ``Batch``es on ``device`` (graphs first..first+n_graphs-1)."""
    sv, st = uv_sphere()
    nv = sv.shape[0]
    pos_l, def_l, rpos_l, rhead_l, centers, heads = [], [], [], [], [], []
    for g in range(first, first + n_graphs):
        gen = torch.Generator().manual_seed(1234 + g)
        pos = torch.rand(n_nodes, 3, generator=gen) - 0.5
        dpos = pos + 0.01 * torch.randn(pos.shape, generator=gen)
        ci = int(torch.randint(0, n_nodes, (1,), generator=gen))
        fdir = torch.randn(3, generator=gen)
        fdir = fdir / fdir.norm()
        force = torch.rand(1, generator=gen) * FORCE_MAX / FORCE_MAX
        pos_l.append(pos)
        def_l.append(dpos)
        rpos_l.append((sv + pos[ci].double()).float())
        rhead_l.append(torch.cat([fdir, force]).repeat(nv, 1))
        centers.append(pos[ci].double())
        heads.append(torch.cat([fdir, force]))
    pos = torch.cat(pos_l).to(device)
    dpos = torch.cat(def_l).to(device)
    rpos = torch.cat(rpos_l).to(device)
    ptr = torch.arange(n_graphs + 1, dtype=torch.long, device=device) * n_nodes
    rptr = torch.arange(n_graphs + 1, dtype=torch.long, device=device) * nv
    ei = ops.table_to_edge_index(ops.knn_table(pos, k, ptr=ptr))
    x = ops.posenc(pos)
    rx = torch.empty((rpos.shape[0], 25), dtype=torch.float32, device=device)
    rx[:, :4] = torch.cat(rhead_l).to(device)
    ops.posenc(rpos, out=rx, col0=4)
    tri = st.to(device)
    rei = torch.empty((2, 3 * tri.shape[0] * n_graphs), dtype=torch.long, device=device)
    for g in range(n_graphs):
        ops.mesh_edges(tri, offset=g * nv, out=rei, start=g * 3 * tri.shape[0])

    def batch(x_, ei_, pos_, ptr_, n_per, e_per):
        b = Batch(x=x_, edge_index=ei_, pos=pos_)
        b.ptr = ptr_
        b.batch = torch.arange(n_graphs, device=device).repeat_interleave(n_per)
        b._ptr_host = [i * n_per for i in range(n_graphs + 1)]
        b._edge_ptr = [i * e_per for i in range(n_graphs + 1)] if e_per is not None else None
        return b

    e_per = ei.shape[1] // n_graphs if ei.shape[1] == n_graphs * n_nodes * k else None
    rest = batch(x, ei, pos, ptr, n_nodes, e_per)
    deformed = batch(x, ei, dpos, ptr, n_nodes, e_per)
    rigid = batch(rx, rei, rpos, rptr, nv, 3 * tri.shape[0])
    # the raw per-sample collider parameters (contact point, force direction | magnitude): what assemble.collider_batch takes
    rigid._centers, rigid._head = torch.stack(centers), torch.stack(heads)
    return rest, rigid, deformed
