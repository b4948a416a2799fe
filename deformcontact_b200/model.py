"""``GraphNet`` with the B200 message-passing layers in place of the PyG convs.

Mirrors models/model.py:23-97 (constructor arguments, module names — hence state-dict keys —
forward order) and models/model_loader.py:3-16.  The encoder loops (models/model.py:69-78) —
the hot path — run in libdcb200 with bias+ReLU fused into the layer epilogue; the cross
attention (models/model.py:7-21,82) runs on the same tcgen05 3xTF32 GEMMs
(attention.py), and so does the decoder MLP (:52-64,88; mlp.py, the concatenation of :84 passed as K-segments of
its first GEMM).  ``DCB200_ATTENTION / DCB200_DECODER / DCB200_LOSS = torch`` select plain fp32 torch for comparison.

``attn_group``: the reference attention is unmasked over the whole batch (a soft node attends
to the collider nodes of *every* sample, models/model.py:16-18).  ``attn_group=None``
reproduces that literally.  ``attn_group=G`` applies the same attention inside each consecutive
group of G graphs — numerically the reference run on mini-batches of G
(configs/everyday.json:26 has G=4) — which keeps large batches from needing an
O(sum Ns x sum Nr) score matrix (62 GB/head at batch 64).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

import os

from . import attention, layers, mlp

# "dcb200": the cross attention runs on libdcb200 (attention.py); "torch": plain fp32 torch / cuBLAS (round-1 path)
ATTENTION_IMPL = os.environ.get("DCB200_ATTENTION", "dcb200")
# same switch for the decoder MLP (mlp.py) and the fused training losses (dc_edge_loss)
DECODER_IMPL = os.environ.get("DCB200_DECODER", "dcb200")
LOSS_IMPL = os.environ.get("DCB200_LOSS", "dcb200")

EVERYDAY = dict(input_dims=[21, 25], hidden_dim=256, output_dim=3, encoder_layers=2, decoder_layers=3,
                dropout_rate=0.0, knn_k=7, backbone="TAGConv", use_mha=True, num_mha_heads=2,
                mode="res")  # configs/everyday.json:36-48


class MultiHeadAttention(nn.Module):
    """models/model.py:7-21: per head softmax(L(xr) L(xg)^T) xg with ONE shared Linear for Q and K,
    no 1/sqrt(d); heads concatenated."""

    def __init__(self, feature_dim, num_heads=8):
        super().__init__()
        self.num_heads = num_heads
        self.attention_heads = nn.ModuleList([nn.Linear(feature_dim, feature_dim) for _ in range(num_heads)])

    def forward(self, x_resting, x_rigid):
        """2-D inputs [Ns,F],[Nr,F] (reference form) or 3-D [G,ns,F],[G,nr,F] (equal-size groups)."""
        outs = []
        for head in self.attention_heads:
            q, k = head(x_resting), head(x_rigid)
            attn = F.softmax(torch.matmul(q, k.transpose(-1, -2)), dim=-1)
            outs.append(torch.matmul(attn, x_rigid))
        return torch.cat(outs, dim=-1)


def _host_ptr(graph):
    p = getattr(graph, "_ptr_host", None)
    if p is None:
        p = graph.ptr.tolist()
        try:
            graph._ptr_host = p
        except Exception:
            pass
    return p


BRANCH_STREAMS = os.environ.get("DCB200_BRANCH_STREAMS", "1") == "1"   # collider encoder branch on a second CUDA stream (0: one stream)
_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class GraphNet(nn.Module):
    def __init__(self, input_dims, hidden_dim, output_dim, encoder_layers, decoder_layers, dropout_rate, knn_k,
                 backbone, use_mha, num_mha_heads, mode, attn_group=None):
        super().__init__()
        self.encoder_layers, self.decoder_layers, self.backbone = encoder_layers, decoder_layers, backbone
        self.use_mha, self.dropout_rate, self.knn_k, self.mode = use_mha, dropout_rate, knn_k, mode
        self.attn_group = attn_group
        conv_layer = (layers.GATConv if backbone == "GATConv" else layers.GCNConv if backbone == "GCNConv"
                      else layers.MPNNLayer if backbone == "MPNN"   # extension (north_star's edge/node-MLP layer)
                      else layers.TAGConv)  # models/model.py:39 — any other string selects TAGConv
        d_rest, d_rigid = input_dims[0], input_dims[1]
        self.conv_layers_resting = nn.ModuleList()
        self.conv_layers_rigid = nn.ModuleList()
        for _ in range(encoder_layers):
            self.conv_layers_resting.append(conv_layer(d_rest, hidden_dim))
            d_rest = hidden_dim
        for _ in range(encoder_layers):
            self.conv_layers_rigid.append(conv_layer(d_rigid, hidden_dim))
            d_rigid = hidden_dim
        d = hidden_dim * (num_mha_heads + 1) if use_mha else hidden_dim * 2
        dec = []
        for _ in range(decoder_layers):
            dec += [nn.Linear(d, hidden_dim), nn.ReLU(), nn.Dropout(dropout_rate)]
            d = hidden_dim
        dec.append(nn.Linear(hidden_dim, output_dim))
        self.decoder = nn.Sequential(*dec)
        self.multihead_attention = MultiHeadAttention(hidden_dim, num_heads=num_mha_heads)

    def encode(self, graph_resting, graph_rigid):
        """models/model.py:69-78: x = dropout(relu(conv(x, edge_index))) per layer, both branches."""
        ptr_rest = _host_ptr(graph_resting) if getattr(graph_resting, "ptr", None) is not None else None
        ptr_rigid = _host_ptr(graph_rigid) if getattr(graph_rigid, "ptr", None) is not None else None

        drop = (lambda t: F.dropout(t, p=self.dropout_rate, training=self.training)) if self.dropout_rate > 0 else None

        def branch(convs, graph, ptr):
            return layers.layer_stack(list(convs), graph.x, graph.edge_index, relu=True, ptr=ptr, between=drop)

        if BRANCH_STREAMS and graph_rigid.x.is_cuda:
            # The two encoder branches are independent until the attention.  The collider branch is small (762-node graphs): on
            # a second stream its latency-bound kernels (1-3 waves of tiles at 32 graphs per GPU) run next to the resting branch
            # instead of in front of it; autograd replays each branch's backward on the stream of its forward.
            main = torch.cuda.current_stream()
            side = _side_stream(graph_rigid.x.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                x_rigid = branch(self.conv_layers_rigid, graph_rigid, ptr_rigid)
            x_resting = branch(self.conv_layers_resting, graph_resting, ptr_rest)
            main.wait_stream(side)
            x_rigid.record_stream(main)   # allocated under `side`, consumed by the attention on `main`
            return x_resting, x_rigid
        return branch(self.conv_layers_resting, graph_resting, ptr_rest), branch(self.conv_layers_rigid, graph_rigid, ptr_rigid)

    def attend(self, x_resting, x_rigid, graph_resting, graph_rigid):
        G = self.attn_group
        if ATTENTION_IMPL == "dcb200":   # N1: tcgen05 3xTF32 GEMMs + fused softmax kernels (attention.py)
            has_ptr = getattr(graph_resting, "ptr", None) is not None and getattr(graph_rigid, "ptr", None) is not None
            ps = _host_ptr(graph_resting) if has_ptr else [0, x_resting.shape[0]]
            pr = _host_ptr(graph_rigid) if has_ptr else [0, x_rigid.shape[0]]
            return attention.cross_attention(x_resting, x_rigid, self.multihead_attention.attention_heads, ps, pr, G,
                                             concat=DECODER_IMPL != "dcb200")
        if G is None:
            return self.multihead_attention(x_resting, x_rigid)
        ps, pr = _host_ptr(graph_resting), _host_ptr(graph_rigid)
        B = len(ps) - 1
        bounds = [(g0, min(g0 + G, B)) for g0 in range(0, B, G)]
        ns = {ps[b] - ps[a] for a, b in bounds}
        nr = {pr[b] - pr[a] for a, b in bounds}
        if len(ns) == 1 and len(nr) == 1:  # equal-size groups: one batched matmul per head
            F_ = x_resting.shape[1]
            out = self.multihead_attention(x_resting.view(len(bounds), -1, F_), x_rigid.view(len(bounds), -1, F_))
            return out.reshape(x_resting.shape[0], -1)
        return torch.cat([self.multihead_attention(x_resting[ps[a]:ps[b]], x_rigid[pr[a]:pr[b]]) for a, b in bounds], 0)

    def decode(self, segs):
        """models/model.py:52-64 (Linear, ReLU, Dropout)* + Linear, on dc_gemm with bias + ReLU fused (mlp.py)."""
        mods = list(self.decoder)
        x, i = None, 0
        while i < len(mods):
            lin = mods[i]
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            x = mlp.linear(segs if x is None else [x], lin.weight, lin.bias, relu)
            i += 2 if relu else 1
            if i < len(mods) and isinstance(mods[i], nn.Dropout):
                if mods[i].p > 0:
                    x = F.dropout(x, p=mods[i].p, training=self.training)
                i += 1
        return x

    def forward(self, graph_resting, graph_rigid):
        x_resting, x_rigid = self.encode(graph_resting, graph_rigid)
        pooled = self.attend(x_resting, x_rigid, graph_resting, graph_rigid)
        if DECODER_IMPL == "dcb200":   # the concatenation of models/model.py:84 becomes K-segments of the first GEMM
            x_out = self.decode([x_resting] + (list(pooled) if isinstance(pooled, (list, tuple)) else [pooled]))
        else:
            x_out = self.decoder(torch.cat([x_resting, pooled], dim=-1))
        deformed = graph_resting.clone()
        deformed._structure_of = graph_resting.edge_index   # same edges: lets fused_losses reuse the encoder's cached CSR pair
        if self.mode == "res":
            deformed.pos = deformed.pos + x_out  # models/model.py:93 (out-of-place: keeps autograd simple)
        elif self.mode == "rec":
            deformed.pos = x_out
        return deformed


def load_model(config=None, **overrides):
    """models/model_loader.py:3-16.  ``config`` is the reference ``Config`` object (reads
    ``config.network.*``), a plain dict of constructor arguments, or None for everyday.json."""
    if config is None:
        kw = dict(EVERYDAY)
    elif isinstance(config, dict):
        kw = dict(config)
    else:
        n = config.network
        kw = dict(input_dims=list(n.input_dims), hidden_dim=n.hidden_dim, output_dim=n.output_dim,
                  encoder_layers=n.encoder_layers, decoder_layers=n.decoder_layers, dropout_rate=n.dropout_rate,
                  knn_k=n.knn_k, use_mha=n.use_mha, num_mha_heads=n.num_mha_heads, backbone=n.backbone, mode=n.mode)
    kw.update(overrides)
    kw["input_dims"] = list(kw["input_dims"])
    return GraphNet(**kw)


class GradientConsistencyLoss(nn.Module):
    """models/losses.py:7-19."""

    def forward(self, pred, rest):
        er = rest.pos[rest.edge_index[1]] - rest.pos[rest.edge_index[0]]
        ep = pred.pos[pred.edge_index[1]] - pred.pos[pred.edge_index[0]]
        return (er - ep).norm(p=2, dim=-1).sum() / rest.edge_index.shape[1]


class _FusedLossFn(torch.autograd.Function):
    """N2: both losses and their gradients in one pass over the CSR pair of the soft graph (dc_edge_loss)."""

    @staticmethod
    def forward(ctx, pred_pos, tgt_pos, g):
        from . import ops, _abi
        pred_pos, tgt_pos = pred_pos.contiguous(), tgt_pos.contiguous()
        N = pred_pos.shape[0]
        partial = torch.empty((N, 2), dtype=torch.float32, device=pred_pos.device)
        gc, gl = torch.empty_like(pred_pos), torch.empty_like(pred_pos)
        rp_t, nb_t, _ = g.t
        _abi.call("dc_edge_loss", ops._ptr(g.rowptr), ops._ptr(g.nbr), ops._ptr(rp_t), ops._ptr(nb_t), ops._ptr(pred_pos),
                  ops._ptr(tgt_pos), N, ops._ptr(partial), ops._ptr(gc), ops._ptr(gl), ops._stream())
        sums = ops.colsum(partial)
        E = max(g.E, 1)
        ctx.save_for_backward(gc, gl)
        ctx.scales = (1.0 / (3 * max(N, 1)), 1.0 / E)
        return sums[1] * ctx.scales[0], sums[0] * ctx.scales[1]

    @staticmethod
    def backward(ctx, g_l1, g_c):
        gc, gl = ctx.saved_tensors
        return gl * (g_l1 * ctx.scales[0]) + gc * (g_c * ctx.scales[1]), None, None


def fused_losses(pred, tgt):
    """(L1 displacement loss, gradient consistency loss) = (F.l1_loss(pred.pos, tgt.pos), GradientConsistencyLoss()(pred, tgt))
    of train.py:47-58, computed by one libdcb200 kernel on the CSR pair the encoder built for ``pred.edge_index``."""
    from . import ops
    ptr = _host_ptr(pred) if getattr(pred, "ptr", None) is not None else None
    g = ops.graph_csr(getattr(pred, "_structure_of", pred.edge_index), pred.pos.shape[0], "tag", ptr,
                      reorder=False)   # dc_edge_loss reads rowptr / nbr with the positions in their original order
    return _FusedLossFn.apply(pred.pos, tgt.pos, g)


def train_step_loss(model, soft_rest, rigid, soft_def, lambda_gradient=1.0):
    """train.py:46-58 without the logging syncs: displacement L1 + lambda * consistency."""
    pred = model(soft_rest, rigid)
    pred.pos = pred.pos - soft_rest.pos
    tgt = soft_def.clone()
    tgt.pos = soft_def.pos - soft_rest.pos
    if LOSS_IMPL == "dcb200" and pred.pos.is_cuda:
        loss_l1, loss_c = fused_losses(pred, tgt)
    else:
        loss_l1 = F.l1_loss(pred.pos, tgt.pos)
        loss_c = GradientConsistencyLoss()(pred, tgt)
    return loss_l1 + lambda_gradient * loss_c, loss_l1, loss_c
