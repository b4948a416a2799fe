"""Data parallelism over independent graphs.

A batched graph is block-diagonal (``Batch.from_data_list`` offsets indices per graph, no edge
crosses graphs), so every hop, the CSR build and kNN are independent per graph: ranks take
contiguous ranges of graphs and run the whole step on their shard with NO data-path collective.
Training adds exactly one exchange: a SUM all-reduce of the flat fp32 gradient buffer
(1,033,219 floats = 4.13 MB for everyday.json) over NCCL / NVLink, then the identical optimizer
step on every rank.  Inference needs no communication.
"""
import os

import torch
import torch.distributed as dist


def init(backend=None):
    """One process per GPU (torchrun env: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


def shard_range(n_graphs, rank, world, weights=None):
    """Contiguous range of graphs for ``rank``; balanced by ``weights`` (e.g. edges per graph)
    when given, else by count.  Returns (first, last_exclusive)."""
    if weights is None:
        base, rem = divmod(n_graphs, world)
        first = rank * base + min(rank, rem)
        return first, first + base + (1 if rank < rem else 0)
    w = torch.as_tensor(weights, dtype=torch.float64)
    c = torch.cat([torch.zeros(1, dtype=torch.float64), w.cumsum(0)])
    total = float(c[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(torch.searchsorted(c, torch.tensor(total * r / world, dtype=torch.float64), right=False)))
    cuts.append(n_graphs)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts[rank], cuts[rank + 1]


class FlatGrads:
    """All parameter gradients as views of ONE contiguous fp32 buffer, so the training exchange
    is a single all-reduce (latency-bound at 4 MB: one launch instead of ~30)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(total, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, scale=None):
        """SUM over ranks (losses are pre-scaled by each rank's share of nodes / edges so the sum
        is the reference's global mean)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if scale is not None:
            self.flat.mul_(scale)


def loss_shares(n_nodes_local, n_edges_local, device):
    """(node_share, edge_share) of this rank in the global batch: multiply the local L1 (mean over
    3*N_r) by node_share and the local consistency loss (mean over E_r) by edge_share; the
    SUM-all-reduced gradient is then that of the global-mean losses (train.py:52-58)."""
    t = torch.tensor([float(n_nodes_local), float(n_edges_local)], dtype=torch.float64, device=device)
    tot = t.clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    s = (t / tot).tolist()
    return s[0], s[1]
