"""Restated ``GraphNet`` wiring, loss and train-step objective (oracle; CPU; test infra only).

Follows ``models/model.py:7-21`` (MultiHeadAttention), ``:23-65`` (ctor), ``:67-97`` (forward),
``models/model_loader.py:3-16``, ``models/losses.py:7-19`` and ``train.py:46-58``.
The *wiring* is PINNED: ``tests/golden/make_golden.py`` runs the reference's own
``models/model.py`` (with these oracle convs injected for the un-installable
``torch_geometric.nn``) and commits its outputs; ``tests/test_oracle.py`` compares.

One documented generalisation: ``attn_group``.  The reference attention is unmasked over the
whole batch (``models/model.py:16-18``); with ``attn_group=None`` that is reproduced
literally.  ``attn_group=G`` applies the same unmasked attention independently inside each
consecutive group of G graphs — numerically what the reference computes for mini-batches of
G (``configs/everyday.json:26``: G = 4) — so large batches do not need an
O(sum Ns x sum Nr) score matrix (SURVEY.md section 7 H4 option b).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import convs

EVERYDAY = dict(input_dims=[21, 25], hidden_dim=256, output_dim=3, encoder_layers=2,
                decoder_layers=3, dropout_rate=0.0, knn_k=7, backbone="TAGConv",
                use_mha=True, num_mha_heads=2, mode="res")  # configs/everyday.json:36-48


class MultiHeadAttention(nn.Module):
    def __init__(self, feature_dim, num_heads=8):
        super().__init__()
        self.num_heads = num_heads
        self.attention_heads = nn.ModuleList([nn.Linear(feature_dim, feature_dim) for _ in range(num_heads)])

    def forward(self, x_resting, x_rigid):
        outs = []
        for head in self.attention_heads:
            scores = torch.mm(head(x_resting), head(x_rigid).t())
            outs.append(torch.mm(F.softmax(scores, dim=-1), x_rigid))
        return torch.cat(outs, dim=-1)


class GraphNet(nn.Module):
    def __init__(self, input_dims, hidden_dim, output_dim, encoder_layers, decoder_layers, dropout_rate,
                 knn_k, backbone, use_mha, num_mha_heads, mode, attn_group=None):
        super().__init__()
        self.encoder_layers, self.decoder_layers, self.backbone = encoder_layers, decoder_layers, backbone
        self.use_mha, self.dropout_rate, self.knn_k, self.mode = use_mha, dropout_rate, knn_k, mode
        self.attn_group = attn_group
        conv_layer = (convs.GATConv if backbone == "GATConv" else convs.GCNConv if backbone == "GCNConv"
                      else convs.MPNNLayer if backbone == "MPNN" else convs.TAGConv)
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
        x_resting = graph_resting.x
        for conv in self.conv_layers_resting:
            x_resting = F.relu(conv(x_resting, graph_resting.edge_index))
            x_resting = F.dropout(x_resting, p=self.dropout_rate, training=self.training)
        x_rigid = graph_rigid.x
        for conv in self.conv_layers_rigid:
            x_rigid = F.relu(conv(x_rigid, graph_rigid.edge_index))
            x_rigid = F.dropout(x_rigid, p=self.dropout_rate, training=self.training)
        return x_resting, x_rigid

    def attend(self, x_resting, x_rigid, graph_resting, graph_rigid):
        G = self.attn_group
        if G is None:
            return self.multihead_attention(x_resting, x_rigid)
        ps, pr = graph_resting.ptr.tolist(), graph_rigid.ptr.tolist()
        B = len(ps) - 1
        outs = []
        for g0 in range(0, B, G):
            g1 = min(g0 + G, B)
            outs.append(self.multihead_attention(x_resting[ps[g0]:ps[g1]], x_rigid[pr[g0]:pr[g1]]))
        return torch.cat(outs, 0)

    def forward(self, graph_resting, graph_rigid):
        x_resting, x_rigid = self.encode(graph_resting, graph_rigid)
        pooled = self.attend(x_resting, x_rigid, graph_resting, graph_rigid)
        x_out = self.decoder(torch.cat([x_resting, pooled], dim=-1))
        deformed = graph_resting.clone()
        if self.mode == "res":
            deformed.pos = deformed.pos + x_out
        elif self.mode == "rec":
            deformed.pos = x_out
        return deformed


def load_model(cfg=None, **over):
    kw = dict(EVERYDAY if cfg is None else cfg)
    kw.update(over)
    return GraphNet(**kw)


class GradientConsistencyLoss(nn.Module):
    """``models/losses.py:7-19``: mean over edges of ||(p_rest[i]-p_rest[j]) - (p_pred[i]-p_pred[j])||_2."""

    def forward(self, pred, rest):
        er = rest.pos[rest.edge_index[1]] - rest.pos[rest.edge_index[0]]
        ep = pred.pos[pred.edge_index[1]] - pred.pos[pred.edge_index[0]]
        return (er - ep).norm(p=2, dim=-1).sum() / len(rest.edge_index[0])


def train_step_loss(model, soft_rest, rigid, soft_def, lambda_gradient=1.0):
    """``train.py:46-58``: predictions and targets are turned into displacement fields,
    loss = L1(mean) + lambda * GradientConsistencyLoss(pred_disp, target_disp)."""
    pred = model(soft_rest, rigid)
    pred.pos = pred.pos - soft_rest.pos
    tgt = soft_def.clone()
    tgt.pos = soft_def.pos - soft_rest.pos
    loss_l1 = F.l1_loss(pred.pos, tgt.pos)
    loss_c = GradientConsistencyLoss()(pred, tgt)
    return loss_l1 + lambda_gradient * loss_c, loss_l1, loss_c
