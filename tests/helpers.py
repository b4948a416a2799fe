"""Parity metric of SURVEY.md section 8(d): fp32 tensors must satisfy BOTH
   rel = ||a - b||_inf / ||b||_inf <= tol   and   allclose(a, b, rtol=tol, atol=tol * ||b||_inf)
with tol = 1e-5 (north_star: "fp32 node outputs and gradients within 1e-5 relative")."""
import torch

TOL = 1e-5


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def assert_close(a, b, tol=TOL, what=""):
    a_, b_ = a.detach().float().cpu(), b.detach().float().cpu()
    assert a_.shape == b_.shape, f"{what}: shape {tuple(a_.shape)} vs {tuple(b_.shape)}"
    scale = b_.abs().max().item()
    r = rel_err(a_, b_)
    assert r <= tol, f"{what}: rel err {r:.3e} > {tol}"
    assert torch.allclose(a_, b_, rtol=tol, atol=tol * max(scale, 1e-30)), f"{what}: allclose failed (rel {r:.3e})"
    return r


def canonical(edge_index):
    ei = edge_index.detach().cpu()
    n = int(ei.max()) + 1 if ei.numel() else 1
    key = ei[1] * n + ei[0]
    return ei[:, torch.sort(key, stable=True).indices]


def assert_close_arbiter(a, ref32, ref64, tol=TOL, what=""):
    """fp64 arbiter: where the fp32 oracle itself is ill-conditioned
    (||oracle_fp32 - fp64|| > tol, e.g. softmax-gradient cancellation), the CUDA result must be
    as close to the fp64 evaluation as the fp32 oracle is (factor 2), else the plain 1e-5 bar."""
    r = rel_err(a, ref32)
    if r <= tol:
        return r
    e_ours, e_ref = rel_err(a, ref64), rel_err(ref32, ref64)
    assert e_ours <= max(tol, 2.0 * e_ref), (f"{what}: rel err vs fp32 oracle {r:.3e}; vs fp64 arbiter ours {e_ours:.3e}, "
                                             f"fp32 oracle {e_ref:.3e}")
    return e_ours
