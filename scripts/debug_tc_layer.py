import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import deformcontact_b200 as dc, oracle
from deformcontact_b200 import ops
from oracle import synthetic
from helpers import rel_err
rest, rigid, _ = synthetic.make_batch(3, 400, 8)
ei = rest.edge_index
g = torch.Generator().manual_seed(2)
x = torch.randn(rest.x.shape[0], 256, generator=g)
torch.manual_seed(3)
ref = oracle.TAGConv(256, 256)
with torch.no_grad(): ref.bias.uniform_(-0.2, 0.2)
for prec, name in ((ops.GEMM_FP32, "fp32"), (ops.GEMM_PREFER_TC, "prefer_tc"), (ops.GEMM_AUTO, "auto")):
    ours = dc.TAGConv(256, 256, precision=prec); ours.load_state_dict(ref.state_dict()); ours = ours.cuda()
    for relu in (False, True):
        ref.zero_grad(); ours.zero_grad()
        xr = x.clone().requires_grad_(True); xo = x.clone().cuda().requires_grad_(True)
        o_r = ref(xr, ei); o_r = torch.relu(o_r) if relu else o_r
        o_o = ours(xo, ei.cuda(), relu=relu)
        o_r.square().sum().backward(); o_o.square().sum().backward()
        errs = {"out": rel_err(o_o, o_r), "dx": rel_err(xo.grad, xr.grad)}
        for (k, pr), (_, po) in zip(ref.named_parameters(), ours.named_parameters()):
            errs["d" + k] = rel_err(po.grad, pr.grad)
        print(name, "relu" if relu else "lin ", " ".join(f"{k}={v:.1e}" for k, v in errs.items()), flush=True)
# direct gemm checks with strided A views and ragged M
torch.manual_seed(0)
M = 1200
buf = torch.randn(M, 768, device="cuda"); x0 = torch.randn(M, 256, device="cuda")
Ws = [torch.randn(256, 256, device="cuda") * 0.06 for _ in range(4)]
As = [x0] + [buf[:, k * 256:(k + 1) * 256] for k in range(3)]
refo = sum(a.double() @ w.double().t() for a, w in zip(As, Ws))
for prec, name in ((ops.GEMM_FP32, "fp32"), (ops.GEMM_TF32X3, "tf32x3")):
    o = ops.gemm(list(zip(As, Ws)), M, 256, False, True, precision=prec)
    print("strided 4-seg", name, rel_err(o, refo.float()))
dout = torch.randn(M, 256, device="cuda")
for prec, name in ((ops.GEMM_FP32, "fp32"), (ops.GEMM_TF32X3, "tf32x3")):
    o = ops.gemm([(dout, Ws[0])], M, 256, False, False, precision=prec)
    print("dH", name, rel_err(o, (dout.double() @ Ws[0].double()).float()))
