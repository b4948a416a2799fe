"""step.CapturedTrainStep: the optimisation step of train.py:35-73 captured in a CUDA graph must replay bit for bit what the
eager step computes — losses, gradients and the weights after Adam — for resident batches and for raw inputs assembled
inside the graph (N3)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dc():
    import deformcontact_b200 as m
    return m


def _eager(dc, model, opt, rest, rigid, deformed, lam):
    from deformcontact_b200 import ops
    ops.clear_csr_cache()
    for p in model.parameters():
        if p.grad is not None:
            p.grad.zero_()
    pred = model(rest, rigid)
    pred.pos = pred.pos - rest.pos
    tgt = deformed.clone()
    tgt.pos = deformed.pos - rest.pos
    l1, lc = dc.fused_losses(pred, tgt)
    loss = 1.0 * l1 + (lam * 1.0) * lc
    loss.backward()
    opt.step()
    return loss.detach().clone(), l1.detach().clone(), lc.detach().clone()


@pytest.mark.parametrize("hidden,attn_group", [(64, 2), (256, 4)])
def test_captured_step_replays_the_eager_step_bit_for_bit(dc, hidden, attn_group):
    from deformcontact_b200 import synthetic
    lam = 0.7
    batches = [synthetic.make_batch(4, 300, 8, first=4 * i) for i in range(3)]
    torch.manual_seed(0)
    m_e = dc.load_model(hidden_dim=hidden, attn_group=attn_group).cuda()
    m_g = copy.deepcopy(m_e)
    o_e = torch.optim.Adam(m_e.parameters(), lr=4e-4, capturable=True)
    o_g = torch.optim.Adam(m_g.parameters(), lr=4e-4, capturable=True)
    runner = dc.CapturedTrainStep(m_g, o_g, *batches[0], lambda_gradient=lam)
    assert len(runner.graphs) == 1 and runner.launches_per_step > 50
    for pe, pg in zip(m_e.parameters(), m_g.parameters()):      # capture (and its warm-up steps) left no trace
        assert torch.equal(pe, pg)
    for rest, rigid, deformed in batches:
        le = _eager(dc, m_e, o_e, rest, rigid, deformed, lam)
        lg = [t.clone() for t in runner.run(rest, rigid, deformed)]
        for a, b in zip(le, lg):
            assert torch.equal(a, b)
        for (k, pe), (_, pg) in zip(m_e.named_parameters(), m_g.named_parameters()):
            assert torch.equal(pe.grad, pg.grad), f"grad {k}"
            assert torch.equal(pe, pg), f"weight {k}"


def test_captured_step_rejects_other_shapes(dc):
    from deformcontact_b200 import synthetic, _abi
    torch.manual_seed(0)
    model = dc.load_model(hidden_dim=32, attn_group=2).cuda()
    opt = torch.optim.Adam(model.parameters(), lr=4e-4, capturable=True)
    runner = dc.CapturedTrainStep(model, opt, *synthetic.make_batch(2, 200, 8))
    with pytest.raises(_abi.DcError):
        runner.run(*synthetic.make_batch(2, 150, 8))
    with pytest.raises(_abi.DcError):
        dc.CapturedTrainStep(model, torch.optim.Adam(model.parameters(), lr=4e-4), *synthetic.make_batch(2, 200, 8))


def test_captured_step_with_batch_assembly_inside_the_graph(dc):
    """Raw form (bench.py's end-to-end arm): positions, graph-local edges and collider parameters go into static buffers; the
    batch assembly kernels (N3) run inside the captured graph."""
    from deformcontact_b200 import synthetic
    B, n = 4, 250

    def raw_of(first):
        rest, rigid, deformed = synthetic.make_batch(B, n, 8, first=first)
        local_e = rest.edge_index - rest.ptr[:-1][rest.batch[rest.edge_index[1]]]
        return (rest, rigid, deformed), {"rp": rest.pos, "dp": deformed.pos, "le": local_e, "nptr": rest.ptr,
                                         "eptr": torch.tensor(rest._edge_ptr, dtype=torch.long, device="cuda"),
                                         "centers": rigid._centers.cuda(), "head": rigid._head.cuda().float()}, rest._edge_ptr, rest._ptr_host

    (b0, raw0, eptr_l, nptr_l), (b1, raw1, _, _) = raw_of(0), raw_of(B)

    def assemble(raw):
        rb = dc.graph_batch_packed(raw["rp"], raw["le"], raw["nptr"], raw["eptr"], node_ptr_host=nptr_l, edge_ptr_host=eptr_l)
        db = dc.Batch(x=rb.x, edge_index=rb.edge_index, pos=raw["dp"]); db.ptr = rb.ptr
        return rb, dc.collider_batch_device(raw["centers"], raw["head"]), db

    torch.manual_seed(0)
    m_e = dc.load_model(hidden_dim=64, attn_group=2).cuda()
    m_g = copy.deepcopy(m_e)
    o_e = torch.optim.Adam(m_e.parameters(), lr=4e-4, capturable=True)
    o_g = torch.optim.Adam(m_g.parameters(), lr=4e-4, capturable=True)
    runner = dc.CapturedTrainStep(m_g, o_g, raw=raw0, assemble=assemble)
    for (rest, rigid, deformed), raw in ((b0, raw0), (b1, raw1)):
        host = {k: v.cpu().pin_memory() for k, v in raw.items()}      # as a loader would hand them over
        le = _eager(dc, m_e, o_e, *assemble(raw), 1.0)
        lg = [t.clone() for t in runner.run(raw=host)]
        assert torch.equal(le[0], lg[0])
        for pe, pg in zip(m_e.parameters(), m_g.parameters()):
            assert torch.equal(pe, pg)
    # the assembled batch equals the resident one (indices and positions bit for bit), so this is the same step
    rb, gb, _ = assemble(raw1)
    assert torch.equal(rb.edge_index, b1[0].edge_index) and torch.equal(gb.edge_index, b1[1].edge_index)
    assert torch.equal(rb.pos, b1[0].pos) and torch.equal(gb.pos, b1[1].pos)
