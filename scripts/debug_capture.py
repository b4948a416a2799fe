"""Isolate which part of the step misbehaves under CUDA-graph replay: python scripts/debug_capture.py <stage>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import ops, synthetic

stage = sys.argv[1]
rest, rigid, deformed = synthetic.make_batch(4, 300, 8)
torch.manual_seed(0)
model = dc.load_model(hidden_dim=64, attn_group=2).cuda()
x = torch.randn(1200, 64, device="cuda")
OPT = torch.optim.Adam(model.parameters(), lr=4e-4, capturable=True)
keep = []


def body():
    ops.clear_csr_cache()
    if stage == "csr":
        g = ops.GraphCSR(rest.edge_index, 1200, "tag", rest._ptr_host); _ = g.t
        return g.rowptr.sum()
    if stage == "hop":
        g = ops.graph_csr(rest.edge_index, 1200, "tag", rest._ptr_host)
        return g.propagate(x).sum()
    if stage == "chain":
        conv = model.conv_layers_resting[1]
        with torch.no_grad():
            return conv(x, rest.edge_index, relu=True, ptr=rest._ptr_host).sum()
    if stage == "gemm":
        return ops.gemm([(x, model.conv_layers_resting[1].lins[0].weight)], 1200, 64).sum()
    if stage == "gemmb":      # one batched launch (host-built problem table staged through pinned memory)
        A = [x[i * 300:(i + 1) * 300] for i in range(3)]
        outs = [torch.empty(300, 300, device="cuda") for _ in range(3)]
        ops.gemm_batched([(a, a, o) for a, o in zip(A, outs)], trans_b=True)
        return sum(o.sum() for o in outs)
    if stage == "attn":
        from deformcontact_b200 import attention
        xs, xr = x, x[:600] * 0.5
        with torch.no_grad():
            o = attention.cross_attention(xs, xr, model.multihead_attention.attention_heads, [0, 600, 1200], [0, 300, 600], 1)
        return o.sum()
    if stage == "adam":
        for p in model.parameters():
            p.grad = torch.ones_like(p) * 1e-3
        OPT.step()
        return sum(p.sum() for p in model.parameters())
    if stage == "encode":
        with torch.no_grad():
            a, b = model.encode(rest, rigid)
        return a.sum() + b.sum()
    if stage == "fwd":
        with torch.no_grad():
            return model(rest, rigid).pos.sum()
    if stage == "loss":
        with torch.no_grad():
            return dc.train_step_loss(model, rest, rigid, deformed)[0]
    if stage == "bwd":
        for p in model.parameters():
            if p.grad is not None:
                p.grad.zero_()
        l = dc.train_step_loss(model, rest, rigid, deformed)[0]
        l.backward()
        return l.detach()
    raise SystemExit("unknown stage")


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        ref = body()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
print(stage, "eager ok", float(ref), flush=True)
ops.CAPTURE_KEEPALIVE = keep
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = body()
ops.CAPTURE_KEEPALIVE = None
torch.cuda.synchronize()
print(stage, "captured", flush=True)
g.replay()
torch.cuda.synchronize()
print(stage, "replayed", float(out), "match" if float(out) == float(ref) else "MISMATCH", flush=True)
