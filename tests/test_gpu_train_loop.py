"""The CUDA path against (1) the reference's OWN training loop and (2) the oracle at BASELINE config C1's full size.

(1) tests/golden/train_loop.pt was produced by /root/reference/train.py:train(config) executed unmodified for one epoch over
three mini-batches (tests/golden/make_golden_train.py; train.py:15-135): the three losses it logs per step (train.py:60-69),
its validation loss (train.py:100-105) and the weights it saved (train.py:121-126).  tests/test_oracle.py checks the CPU
oracle against that fixture; here the product — ``dc.load_model`` + ``dc.train_step_loss`` + Adam on cuda:0, every kernel
reached through the C ABI — has to reproduce the same numbers to 1e-5.

(2) BASELINE.json configs[0]: everyday.json model, fwd + bwd on 4 synthetic object graphs of 2000 nodes (kNN k = 8) and
their 4 collider graphs, hidden 256.  Loss, predicted positions and EVERY parameter gradient against the fp32 oracle at
1e-5 (north_star), the fp64 oracle as arbiter where the fp32 oracle itself is further than that from fp64 — and, because the
step is not smooth (ReLU masks, the sign in the L1 loss: at 8000 nodes a single unit flipping at zero moves a weight gradient
by ~1e-4, and which units flip is luck — profiles/r02_accuracy_lab.txt), no tighter than the fp64 gradients themselves move
under a 1e-6 relative perturbation of the weights (helpers.gradient_conditioning).
"""
import copy

import pytest
import torch

import oracle
from oracle import synthetic
from helpers import TOL, assert_close, assert_close_arbiter, assert_close_conditioned, gradient_conditioning, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dc():
    import deformcontact_b200 as m
    return m


def _cu(dc, b, n):
    return dc.Batch.from_data_list([dc.Data(x=b[i].x, edge_index=b[i].edge_index, pos=b[i].pos) for i in range(n)]).to("cuda")


def _golden_batch(m, first):
    rests, defs, rigids = [], [], []
    for g in range(first, first + m["batch"]):
        rest, deformed, gen = synthetic.soft_graph(g, m["nodes"], m["k"])
        ci = int(torch.randint(0, m["nodes"], (1,), generator=gen))
        rests.append(rest); defs.append(deformed); rigids.append(synthetic.rigid_graph(rest.pos[ci], gen))
    return tuple(oracle.Batch.from_data_list(l) for l in (rests, rigids, defs))


@pytest.mark.parametrize("captured", [False, True])
def test_cuda_train_loop_vs_reference_training_loop(dc, golden_dir, captured):
    """train.py:35-73 (three optimisation steps), :100-105 (validation), :121-126 (saved weights) on the CUDA path.
    ``captured``: the same three steps through step.CapturedTrainStep (one CUDA graph per step shape, replayed)."""
    gold = torch.load(f"{golden_dir}/train_loop.pt")
    m = gold["meta"]
    torch.manual_seed(m["seed"])
    model = dc.load_model(hidden_dim=m["hidden"]).cuda()      # same constructor order => same initial weights as the reference
    opt = torch.optim.Adam(model.parameters(), lr=m["lr"], capturable=captured)
    model.train()
    runner = None
    for b, logged in enumerate(gold["steps"]):
        rest, rigid, deformed = (_cu(dc, x, m["batch"]) for x in _golden_batch(m, b * m["batch"]))
        if captured:
            from deformcontact_b200.step import CapturedTrainStep
            if runner is None:
                runner = CapturedTrainStep(model, opt, rest, rigid, deformed, lambda_gradient=m["lambda_gradient"])
            loss, l1, lc = runner.run(rest, rigid, deformed)
        else:
            loss, l1, lc = dc.train_step_loss(model, rest, rigid, deformed, lambda_gradient=m["lambda_gradient"])
            opt.zero_grad()
            loss.backward()
            opt.step()
        for got, key in ((loss, "tr_loss"), (l1, "tr_mse_loss"), (lc, "tr_consistency_loss")):
            assert abs(got.item() - logged[key]) <= TOL * abs(logged[key]), (b, key, got.item(), logged[key])
    model.eval()
    with torch.no_grad():
        rest, rigid, deformed = (_cu(dc, x, m["batch"]) for x in _golden_batch(m, 1000))
        pred = model(rest, rigid)
        val = torch.nn.functional.l1_loss(pred.pos, deformed.pos) + m["lambda_gradient"] * dc.GradientConsistencyLoss()(pred, deformed)
    assert abs(val.item() - gold["validation_loss"]) <= TOL * abs(gold["validation_loss"])
    sd = model.state_dict()
    assert set(sd) == set(gold["state_dict"])
    for k, v in gold["state_dict"].items():
        assert_close(sd[k], v, what=f"weights saved by the reference loop: {k}")


def test_whole_model_parity_at_baseline_c1_size(dc):
    """BASELINE.json configs[0] at its full size: 4 x 2000-node kNN-8 graphs + 4 x 762-node colliders, everyday.json widths."""
    B, n, k = 4, 2000, 8
    rest, rigid, deformed = synthetic.make_batch(B, n, k)
    torch.manual_seed(0)
    ref = oracle.load_model()
    ours = dc.load_model()
    ours.load_state_dict(ref.state_dict())
    ours = ours.cuda()
    ref64 = copy.deepcopy(ref).double()
    loss_r, l1_r, lc_r = oracle.train_step_loss(ref, rest, rigid, deformed)
    loss_r.backward()
    to64 = lambda b: oracle.Batch.from_data_list([oracle.Data(x=b[i].x.double(), edge_index=b[i].edge_index, pos=b[i].pos.double())
                                                  for i in range(B)])
    loss_64, _, _ = oracle.train_step_loss(ref64, to64(rest), to64(rigid), to64(deformed))
    loss_64.backward()
    loss_o, l1_o, lc_o = dc.train_step_loss(ours, _cu(dc, rest, B), _cu(dc, rigid, B), _cu(dc, deformed, B))
    loss_o.backward()
    assert_close(loss_o, loss_r, what="C1 loss")
    assert_close(l1_o, l1_r, what="C1 L1 loss")
    assert_close(lc_o, lc_r, what="C1 consistency loss")
    with torch.no_grad():
        ours.eval(); ref.eval()
        assert_close(ours(_cu(dc, rest, B), _cu(dc, rigid, B)).pos, ref(rest, rigid).pos, what="C1 predicted positions")
    r64, g64, d64 = to64(rest), to64(rigid), to64(deformed)
    cond = gradient_conditioning(ref64, lambda m: oracle.train_step_loss(m, r64, g64, d64)[0], eps=1e-6, samples=3)
    worst, widened = 0.0, []
    for (kk, pr), (_, po), (_, p64) in zip(ref.named_parameters(), ours.named_parameters(), ref64.named_parameters()):
        assert po.grad is not None, kk
        e = assert_close_conditioned(po.grad, pr.grad, p64.grad, cond[kk], what=f"C1 d{kk}")
        worst = max(worst, e)
        if e > max(TOL, 2.0 * rel_err(pr.grad, p64.grad)):
            widened.append((kk, e, cond[kk]))
    print(f"C1 whole-model parity: worst gradient error {worst:.2e}; parameters judged by the perturbation bound: {widened}")
