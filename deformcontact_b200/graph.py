"""Edge construction and input features on the GPU (drop-ins for utils/graph_utils.py:7-20,
utils/pointcloud_utils.py:7-13, utils/pos_encoding.py:6-44).  All arithmetic is in libdcb200."""
import torch

from . import ops
from . import torch_ops  # noqa: F401  (registers torch.ops.dcb200.*)
from .data import Data


def to_log_freq(x, N_freqs=3, dim=1):
    """utils/pos_encoding.py:to_log_freq for the only configuration the reference uses
    (``to_log_freq(pos, 3, 1)``, utils/graph_utils.py:16): [N,3] -> [N,21]."""
    if N_freqs != 3 or x.dim() != 2 or x.shape[1] != 3 or dim not in (1, -1):
        raise NotImplementedError("to_log_freq: only the reference configuration (N_freqs=3, [N,3], dim=1) is built")
    return ops.posenc(x)


def mesh_to_graph(vertices, triangles, encode=True, device="cuda"):
    """utils/graph_utils.py:7-20 with (vertices, triangles) arrays in place of an Open3D mesh:
    directed half-edges (a,b),(b,c),(c,a) per triangle, ``x = to_log_freq(pos)``."""
    pos = torch.as_tensor(vertices).to(device=device, dtype=torch.float32)
    tri = torch.as_tensor(triangles).to(device=device, dtype=torch.int64).reshape(-1, 3)
    edge_index = ops.mesh_edges(tri)
    x = ops.posenc(pos) if encode else pos
    return Data(x=x, edge_index=edge_index, pos=pos)


def _with_order_hint(edge_index, order, x, batch, ptr):
    """One large cloud: remember the grid-cell order of its points for this edge_index, so the layers can run their hops
    on spatially coherent node labels (ops.REORDER; results unchanged).  The grid search hands the order over for free;
    after a brute-force search it is computed by dc_cell_order."""
    single = batch is None and (ptr is None or ptr.numel() == 2)
    if ops.REORDER and single and x.dim() == 2 and x.shape[1] == 3 and x.shape[0] >= ops.REORDER_MIN_NODES:
        ops.register_order_hint(edge_index, order if order.numel() == x.shape[0] else ops.cell_order(x))
    return edge_index


def knn_graph(x, k, batch=None, loop=False, ptr=None):
    """torch_cluster.knn_graph(x, k, batch, loop, flow='source_to_target') -> int64 [2, E] (``torch.ops.dcb200.knn_graph``)."""
    edge_index, order = torch.ops.dcb200.knn_graph(x, int(k), batch, bool(loop), ptr)
    return _with_order_hint(edge_index, order, x, batch, ptr)


def radius_graph(x, r, batch=None, loop=False, max_num_neighbors=32, ptr=None):
    """torch_cluster.radius_graph(x, r, batch, loop, max_num_neighbors) -> int64 [2, E] (``torch.ops.dcb200.radius_graph``)."""
    edge_index, order = torch.ops.dcb200.radius_graph(x, float(r), batch, bool(loop), int(max_num_neighbors), ptr)
    return _with_order_hint(edge_index, order, x, batch, ptr)


def construct_graph(point_cloud, k=None, radius=None):
    """utils/pointcloud_utils.py:7-13."""
    if radius is not None:
        return radius_graph(point_cloud, radius, batch=None, loop=False)
    return knn_graph(point_cloud, k, batch=None, loop=False)
