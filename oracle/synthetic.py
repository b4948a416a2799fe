"""Deterministic synthetic inputs of SURVEY.md section 8(d) (oracle side; CPU).

Graph g uses ``torch.Generator().manual_seed(1234 + g)``.  Soft graph: ``pos ~ U[-0.5, 0.5)^3``,
``x = to_log_freq(pos)`` (21-d), kNN-k edges (or a mesh-edge variant).  Collider: UV sphere
(762 vertices, 4560 directed edges, radius ``configs/everyday.json:9`` = 0.05) centred on a
random soft vertex, features ``[force_dir(3) | force/force_max(1) | posenc(21)]`` = 25-d
(``loaders/common.py:6-19``, ``loaders/everyday_deform.py:56``).  Target
``pos_def = pos + 0.01 * N(0, 1)``.
"""
import torch

from .data import Data, Batch
from .graphs import to_log_freq, knn_graph, mesh_to_graph, uv_sphere, grid_mesh

FORCE_MAX = 10000.0   # configs/everyday.json:10
SPHERE_R = 0.05       # configs/everyday.json:9


def soft_graph(g, n_nodes=2000, k=8, kind="knn"):
    gen = torch.Generator().manual_seed(1234 + g)
    if kind == "knn":
        pos = torch.rand(n_nodes, 3, generator=gen) - 0.5
        ei = knn_graph(pos, k)
    else:  # mesh-edge variant: open jittered sheet
        nx = max(2, int(round(n_nodes ** 0.5)))
        pos, tri = grid_mesh(nx, nx, jitter=0.002, generator=gen)
        ei = mesh_to_graph(pos, tri).edge_index
    rest = Data(x=to_log_freq(pos, 3, 1), edge_index=ei, pos=pos)
    deformed = Data(x=rest.x, edge_index=ei, pos=pos + 0.01 * torch.randn(pos.shape, generator=gen))
    return rest, deformed, gen


def rigid_graph(center, gen):
    v, t = uv_sphere(SPHERE_R, 20, center=tuple(float(c) for c in center))
    d = mesh_to_graph(v, t)
    fdir = torch.randn(3, generator=gen)
    fdir = fdir / fdir.norm()
    force = torch.rand(1, generator=gen) * FORCE_MAX / FORCE_MAX
    n = d.x.shape[0]
    d.x = torch.cat([fdir.repeat(n, 1), force.repeat(n, 1), d.x], dim=1)
    return d


def make_batch(n_graphs, n_nodes=2000, k=8, kind="knn", first=0):
    rests, defs, rigids = [], [], []
    for g in range(first, first + n_graphs):
        rest, deformed, gen = soft_graph(g, n_nodes, k, kind)
        ci = int(torch.randint(0, rest.pos.shape[0], (1,), generator=gen))
        rests.append(rest)
        defs.append(deformed)
        rigids.append(rigid_graph(rest.pos[ci], gen))
    return (Batch.from_data_list(rests), Batch.from_data_list(rigids), Batch.from_data_list(defs))
