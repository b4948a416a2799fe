"""Generate the committed golden fixtures by running the REFERENCE's own importable code.

Run in the build container only (needs /root/reference; the GPU box has no such path):
    python tests/golden/make_golden.py

What is pinned (reference code executed, not restated):
  * utils/pos_encoding.py:to_log_freq             -> posenc.pt
  * models/losses.py:GradientConsistencyLoss      -> gcl.pt
  * loaders/collate.py:collate_fn                 -> checked structurally (collate.pt)
  * models/model.py:GraphNet (ctor + forward wiring, MultiHeadAttention, decoder, 'res' mode),
    executed unmodified with the oracle's restated convs injected as ``torch_geometric.nn``
    (the real PyG is not installable here)        -> graphnet_{TAGConv,GCNConv,GATConv}.pt
What is NOT pinned by the reference (third-party arithmetic, no reference tests): the conv
layers themselves and knn/radius graphs.  For those the fixtures below are produced by the
oracle (self-pins, so later refactors of the oracle are themselves checked) -> layers.pt, graphs.pt
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import oracle  # noqa: E402
from oracle import synthetic  # noqa: E402


def _inject_pyg():
    tg = types.ModuleType("torch_geometric")
    tgnn = types.ModuleType("torch_geometric.nn")
    tgdata = types.ModuleType("torch_geometric.data")
    tgnn.GATConv, tgnn.GCNConv, tgnn.TAGConv, tgnn.knn = oracle.GATConv, oracle.GCNConv, oracle.TAGConv, None
    tgdata.Data, tgdata.Batch = oracle.Data, oracle.Batch
    tg.nn, tg.data = tgnn, tgdata
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tgnn, "torch_geometric.data": tgdata})


def main():
    sys.path.insert(0, REF)
    _inject_pyg()
    from utils.pos_encoding import to_log_freq as ref_posenc
    from models.losses import GradientConsistencyLoss as RefGCL
    from loaders.collate import collate_fn as ref_collate
    from models.model import GraphNet as RefGraphNet

    g = torch.Generator().manual_seed(7)
    pos = torch.rand(257, 3, generator=g) * 2 - 1
    torch.save({"pos": pos, "out": ref_posenc(pos, 3, 1)}, os.path.join(HERE, "posenc.pt"))

    rest, rigid, deformed = synthetic.make_batch(2, 120, 6)
    pred = rest.clone()
    pred.pos = rest.pos + 0.03 * torch.randn(rest.pos.shape, generator=g)
    torch.save({"pred_pos": pred.pos, "rest_pos": deformed.pos, "edge_index": rest.edge_index,
                "loss": RefGCL()(pred, deformed)}, os.path.join(HERE, "gcl.pt"))

    samples = []
    for i in range(3):
        r, d, gen = synthetic.soft_graph(i, 50, 4)
        samples.append((f"obj{i}", r, d, {"force_vector": torch.randn(3, generator=gen), "force": float(i), "flag": True},
                        synthetic.rigid_graph(r.pos[0], gen)))
    names, rs, ds, meta, rg = ref_collate(samples)
    torch.save({"names": names, "n_rest": len(rs), "meta_force_vector": meta["force_vector"], "meta_force": meta["force"],
                "meta_flag": meta["flag"]}, os.path.join(HERE, "collate.pt"))

    for backbone in ("TAGConv", "GCNConv", "GATConv"):
        torch.manual_seed(11)
        kw = dict(oracle.EVERYDAY, hidden_dim=32, backbone=backbone)
        ref_model = RefGraphNet(**kw)
        out = ref_model(rest, rigid)
        torch.save({"kw": kw, "state_dict": ref_model.state_dict(), "n_graphs": 2, "n_nodes": 120, "k": 6,
                    "out_pos": out.pos.detach()}, os.path.join(HERE, f"graphnet_{backbone}.pt"))

    # oracle self-pins for the third-party arithmetic (parity unpinned by the reference)
    layers = {}
    x, ei = rest.x, rest.edge_index
    for name, cls in (("TAGConv", oracle.TAGConv), ("GCNConv", oracle.GCNConv), ("GATConv", oracle.GATConv)):
        torch.manual_seed(3)
        layer = cls(21, 16)
        with torch.no_grad():
            layer.bias.uniform_(-0.1, 0.1)
        xx = x.clone().requires_grad_(True)
        out = layer(xx, ei)
        out.square().sum().backward()
        layers[name] = {"state_dict": layer.state_dict(), "out": out.detach(), "dx": xx.grad,
                        "grads": {k: p.grad for k, p in layer.named_parameters()}}
    torch.save({"x": x, "edge_index": ei, "layers": layers}, os.path.join(HERE, "layers.pt"))

    pts = torch.rand(300, 3, generator=g)
    batch = torch.arange(3).repeat_interleave(100)
    torch.save({"pts": pts, "batch": batch,
                "knn5": oracle.knn_graph(pts, 5), "knn5_b": oracle.knn_graph(pts, 5, batch),
                "rad": oracle.radius_graph(pts, 0.15), "rad_b": oracle.radius_graph(pts, 0.3, batch)},
               os.path.join(HERE, "graphs.pt"))
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
