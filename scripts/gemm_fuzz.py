"""Random-shape check of the tcgen05 GEMM (CTA pairs by default) against fp64: every operand layout, ragged M / N / K (odd tile
counts, N < 64 so that the peer CTA's B half is empty, K not a multiple of 32), single launches, split-K and batched launches."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deformcontact_b200 as dc
from deformcontact_b200 import ops

random.seed(0)
g = torch.Generator(device="cuda").manual_seed(0)
worst = 0.0
n = 0
def check(out, A, B, ta, tb, what):
    global worst, n
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    err = ((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
    worst = max(worst, err); n += 1
    assert err < 5e-6, (what, err)
for trial in range(60):
    M = random.choice([1, 3, 64, 127, 128, 129, 255, 256, 257, 300, 385, 640, 1000])
    N = random.choice([4, 8, 24, 60, 64, 68, 128, 132, 256, 260, 520])
    K = random.choice([4, 8, 28, 32, 36, 100, 256, 260, 1000, 4100, 20000])
    for ta in (False, True):
        for tb in (False, True):
            Mp, Np, Kp = (M + 3) // 4 * 4, (N + 3) // 4 * 4, (K + 3) // 4 * 4      # 16-byte row pitches
            A = torch.randn((K, Mp) if ta else (M, Kp), generator=g, device="cuda")[:, :M if ta else K]
            B = torch.randn((N, Kp) if tb else (K, Np), generator=g, device="cuda")[:, :K if tb else N]
            out = ops.gemm([(A, B)], M, N, ta, tb, precision=ops.GEMM_TF32X3)
            check(out, A, B, ta, tb, ("single", M, N, K, ta, tb))
            if trial % 3 == 0:
                outs = [torch.empty(M, N, device="cuda") for _ in range(3)]
                ops.gemm_batched([(A, B, o) for o in outs], ta, tb)
                for o in outs:
                    check(o, A, B, ta, tb, ("batched", M, N, K, ta, tb))
torch.cuda.synchronize()
print(f"gemm fuzz ok: {n} products, worst error vs fp64 {worst:.2e} (DCB200_T2_PAIR={os.environ.get('DCB200_T2_PAIR', '1')})")
