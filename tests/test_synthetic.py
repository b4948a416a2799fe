"""The product's synthetic generator (GPU kNN / features) reproduces the oracle's section-8(d) inputs."""
import pytest
import torch

from oracle import synthetic as osyn
from helpers import assert_close

pytestmark = pytest.mark.gpu


def test_make_batch_matches_oracle():
    from deformcontact_b200 import synthetic
    r, g, d = synthetic.make_batch(3, 300, 8, first=2)
    ro, go, do = osyn.make_batch(3, 300, 8, first=2)
    assert torch.equal(r.pos.cpu(), ro.pos) and torch.equal(d.pos.cpu(), do.pos) and torch.equal(g.pos.cpu(), go.pos)
    assert torch.equal(r.edge_index.cpu(), ro.edge_index) and torch.equal(g.edge_index.cpu(), go.edge_index)
    assert torch.equal(r.ptr.cpu(), ro.ptr) and torch.equal(g.ptr.cpu(), go.ptr) and torch.equal(r.batch.cpu(), ro.batch)
    assert_close(r.x, ro.x, tol=2e-6, what="soft x")
    assert_close(g.x, go.x, tol=2e-6, what="rigid x")
    assert torch.equal(g.x[:, :4].cpu(), go.x[:, :4])
    assert torch.equal(r[1].edge_index.cpu(), ro[1].edge_index)
