"""Which component limits whole-model gradient accuracy?  Runs the 4 x 300-node train step (tests/test_gpu_layers.py::
test_graphnet_train_step_vs_oracle[None]) and prints, per parameter, the error of the CUDA gradients against the fp64 oracle
next to the fp32 oracle's own error.  Switch components with the environment:
  DCB200_GEMM=fp32 (exact-fp32 FFMA for every product)  DCB200_ATTENTION=torch  DCB200_DECODER=torch  DCB200_LOSS=torch"""
import copy, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import oracle
from oracle import synthetic
from helpers import rel_err
import deformcontact_b200 as dc

B = 4
attn_group = None if os.environ.get("LAB_GROUP", "none") == "none" else int(os.environ["LAB_GROUP"])
NODES = int(os.environ.get("LAB_NODES", "300"))
rest, rigid, deformed = synthetic.make_batch(B, NODES, 8)
torch.manual_seed(0)
ref = oracle.load_model(attn_group=attn_group)
ours = dc.load_model(attn_group=attn_group)
ours.load_state_dict(ref.state_dict())
ours = ours.cuda()
ref64 = copy.deepcopy(ref).double()
oracle.train_step_loss(ref, rest, rigid, deformed)[0].backward()
to64 = lambda b: oracle.Batch.from_data_list([oracle.Data(x=b[i].x.double(), edge_index=b[i].edge_index, pos=b[i].pos.double()) for i in range(B)])
oracle.train_step_loss(ref64, to64(rest), to64(rigid), to64(deformed))[0].backward()
cu = lambda b: dc.Batch.from_data_list([dc.Data(x=d.x, edge_index=d.edge_index, pos=d.pos) for d in [b[i] for i in range(B)]]).to("cuda")
dc.train_step_loss(ours, cu(rest), cu(rigid), cu(deformed))[0].backward()
# the reference's own formulation on THIS GPU: the oracle port on CUDA tensors, stock torch kernels, true fp32 (TF32 off)
torch.backends.cuda.matmul.allow_tf32 = False
refg = copy.deepcopy(ref).cuda()
refg.zero_grad()
mv = lambda b: oracle.Batch.from_data_list([oracle.Data(x=b[i].x.cuda(), edge_index=b[i].edge_index.cuda(), pos=b[i].pos.cuda()) for i in range(B)])
oracle.train_step_loss(refg, mv(rest), mv(rigid), mv(deformed))[0].backward()
worst = (0, "")
worst_g = (0, "")
for (k, pg), (_, p64) in zip(refg.named_parameters(), ref64.named_parameters()):
    worst_g = max(worst_g, (rel_err(pg.grad, p64.grad), k))
for (k, pr), (_, po), (_, p64) in zip(ref.named_parameters(), ours.named_parameters(), ref64.named_parameters()):
    e_o, e_r = rel_err(po.grad, p64.grad), rel_err(pr.grad, p64.grad)
    worst = max(worst, (e_o, k))
    e_g = rel_err(dict(refg.named_parameters())[k].grad, p64.grad)
    if e_o > 3e-6:
        print(f"  {k:55s} ours-vs-fp64 {e_o:.2e}   fp32oracle(CPU)-vs-fp64 {e_r:.2e}   torch-CUDA-fp32-vs-fp64 {e_g:.2e}")
print("config", {k: v for k, v in os.environ.items() if k.startswith(("DCB200", "LAB"))}, "WORST ours", f"{worst[0]:.3e}", worst[1],
      "| WORST torch-CUDA-fp32", f"{worst_g[0]:.3e}", worst_g[1], flush=True)
