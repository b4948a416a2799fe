"""Restated edge construction + input features (oracle; CPU; test infrastructure only).

* ``mesh_to_graph``  follows ``utils/graph_utils.py:7-20`` (takes vertex / triangle arrays
  instead of an Open3D mesh; Open3D is not installable here).
* ``to_log_freq``    follows ``utils/pos_encoding.py:6-44`` — PINNED against the reference
  function itself (importable) by ``tests/golden/make_golden.py``.
* ``knn_graph`` / ``radius_graph`` / ``construct_graph`` follow
  ``utils/pointcloud_utils.py:7-13`` -> torch_cluster [3P, unpinned version, restated from
  the published algorithm; PARITY UNPINNED].  Conventions the reference leaves open and
  this repo fixes:
    - distance = fp32 ``((dx*dx) + (dy*dy)) + (dz*dz)``, every operation rounded (no FMA);
    - kNN ties: the lower neighbour index wins; neighbours listed by ascending distance;
    - radius: strict ``d2 < fl(r*r)``, candidates taken in ascending index order, at most
      ``max_num_neighbors + 1`` *including* a possible self match, self dropped afterwards
      (torch_cluster's CUDA rule).
* ``uv_sphere`` restates Open3D 0.18 ``TriangleMesh::CreateSphere(radius, resolution=20)``
  (``loaders/common.py:26``): 762 vertices, 1520 triangles  [3P].
"""
import math

import torch

from .data import Data


def to_log_freq(x, N_freqs=3, dim=1):
    freq_bands = 2.0 ** torch.linspace(0.0, N_freqs - 1, steps=N_freqs)
    outs = [x]
    for f in freq_bands:
        outs.append(torch.sin(x * f))
        outs.append(torch.cos(x * f))
    return torch.cat(outs, -1)


def mesh_to_graph(vertices, triangles, encode=True):
    """Per triangle (a, b, c) emit (a->b), (b->c), (c->a) in that order; ``[2, 3T]`` int64."""
    pos = torch.as_tensor(vertices, dtype=torch.float32)
    tri = torch.as_tensor(triangles, dtype=torch.long).reshape(-1, 3)
    src = tri.reshape(-1)
    dst = tri[:, [1, 2, 0]].reshape(-1)
    edge_index = torch.stack([src, dst], 0).contiguous()
    x = to_log_freq(pos, 3, 1) if encode else pos
    return Data(x=x, edge_index=edge_index, pos=pos)


def _sqdist(q, p):
    """fp32 squared distance [Q, P], each op individually rounded, fixed association."""
    dx = q[:, None, 0] - p[None, :, 0]
    dy = q[:, None, 1] - p[None, :, 1]
    dz = q[:, None, 2] - p[None, :, 2]
    return ((dx * dx) + (dy * dy)) + (dz * dz)


def _segments(n, batch):
    if batch is None:
        return [(0, n)]
    counts = torch.bincount(batch)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
    return [(int(ptr[g]), int(ptr[g + 1])) for g in range(len(counts))]


def knn_graph(x, k, batch=None, loop=False, chunk=2048):
    """``torch_cluster.knn_graph(x, k, batch, loop, flow='source_to_target')``:
    search k (+1 if not loop) nearest of every point among the points of its own graph,
    self included; emit (row = neighbour, col = query); then drop row == col."""
    x = x.float()
    kk = k if loop else k + 1
    rows, cols = [], []
    for lo, hi in _segments(x.shape[0], batch):
        p = x[lo:hi]
        n = hi - lo
        kq = min(kk, n)
        for s in range(0, n, chunk):
            q = p[s:s + chunk]
            d2 = _sqdist(q, p)
            idx = torch.sort(d2, dim=1, stable=True).indices[:, :kq]
            qi = torch.arange(s, s + q.shape[0]).unsqueeze(1).expand_as(idx)
            rows.append(idx.reshape(-1) + lo)
            cols.append(qi.reshape(-1) + lo)
    row = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long)
    col = torch.cat(cols) if cols else torch.zeros(0, dtype=torch.long)
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], 0)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, chunk=2048):
    x = x.float()
    r2 = torch.tensor(r, dtype=torch.float32) * torch.tensor(r, dtype=torch.float32)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    rows, cols = [], []
    for lo, hi in _segments(x.shape[0], batch):
        p = x[lo:hi]
        n = hi - lo
        for s in range(0, n, chunk):
            q = p[s:s + chunk]
            hit = _sqdist(q, p) < r2
            rank = hit.long().cumsum(1)
            keep = hit & (rank <= cap)
            qi, pj = keep.nonzero(as_tuple=True)
            rows.append(pj + lo)
            cols.append(qi + s + lo)
    row = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long)
    col = torch.cat(cols) if cols else torch.zeros(0, dtype=torch.long)
    if not loop:
        m = row != col
        row, col = row[m], col[m]
    return torch.stack([row, col], 0)


def construct_graph(point_cloud, k=None, radius=None):
    """``utils/pointcloud_utils.py:7-13``."""
    if radius is not None:
        return radius_graph(point_cloud, radius, batch=None, loop=False)
    return knn_graph(point_cloud, k, batch=None, loop=False)


def canonical_sort(edge_index):
    """Lexicographic by (col, row) — the order parity tests compare in."""
    n = int(edge_index.max()) + 1 if edge_index.numel() else 1
    key = edge_index[1] * n + edge_index[0]
    return edge_index[:, torch.sort(key, stable=True).indices]


def uv_sphere(radius=0.05, resolution=20, center=(0.0, 0.0, 0.0)):
    """(vertices fp64 [762, 3], triangles int64 [1520, 3]) for the defaults."""
    res = resolution
    V = [(0.0, 0.0, radius), (0.0, 0.0, -radius)]
    step = math.pi / res
    for i in range(1, res):
        alpha = step * i
        for j in range(2 * res):
            theta = step * j
            V.append((math.sin(alpha) * math.cos(theta) * radius,
                      math.sin(alpha) * math.sin(theta) * radius,
                      math.cos(alpha) * radius))
    T = []
    for j in range(2 * res):
        j1 = (j + 1) % (2 * res)
        T.append((0, 2 + j, 2 + j1))
        base = 2 + 2 * res * (res - 2)
        T.append((1, base + j1, base + j))
    for i in range(1, res - 1):
        b1 = 2 + 2 * res * (i - 1)
        b2 = b1 + 2 * res
        for j in range(2 * res):
            j1 = (j + 1) % (2 * res)
            T.append((b2 + j, b1 + j1, b1 + j))
            T.append((b2 + j, b2 + j1, b1 + j1))
    v = torch.tensor(V, dtype=torch.float64) + torch.tensor(center, dtype=torch.float64)
    return v, torch.tensor(T, dtype=torch.long)


def grid_mesh(nx, ny, jitter=0.0, generator=None):
    """Open (non-closed) triangulated nx x ny sheet: an asymmetric, mesh-like test graph."""
    ii, jj = torch.meshgrid(torch.arange(nx), torch.arange(ny), indexing="ij")
    pos = torch.stack([ii.reshape(-1) / max(nx - 1, 1) - 0.5,
                       jj.reshape(-1) / max(ny - 1, 1) - 0.5,
                       torch.zeros(nx * ny)], 1).float()
    if jitter:
        pos = pos + jitter * torch.randn(pos.shape, generator=generator)
    vid = (ii * ny + jj)
    a, b, c, d = vid[:-1, :-1], vid[1:, :-1], vid[1:, 1:], vid[:-1, 1:]
    tri = torch.cat([torch.stack([a, b, c], -1).reshape(-1, 3),
                     torch.stack([a, c, d], -1).reshape(-1, 3)], 0)
    return pos, tri
