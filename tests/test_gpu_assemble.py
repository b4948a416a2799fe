"""N3 parity: on-GPU batch assembly vs the reference's own host code (golden fixture made by
tests/golden/make_golden_assemble.py from utils/graph_utils.py:mesh_to_graph and loaders/common.py:_feature_rigid)
and vs the oracle's Batch.from_data_list on seeded inputs.  Index and position outputs are bit-exact; the
sin/cos feature columns are held to 2e-6 (CUDA sinf vs the host libm, as for dc_posenc)."""
import pytest
import torch

import oracle
from oracle import synthetic as osyn
from helpers import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dc():
    import deformcontact_b200 as d
    assert torch.cuda.is_available()
    return d


def _same_batch(out, ref, exact_x=False, what=""):
    assert torch.equal(out.edge_index.cpu(), ref["edge_index"]), f"{what}: edge_index"
    assert out.edge_index.dtype == torch.int64
    assert torch.equal(out.pos.cpu(), ref["pos"]), f"{what}: pos"
    assert torch.equal(out.batch.cpu(), ref["batch"]), f"{what}: batch"
    assert torch.equal(out.ptr.cpu(), ref["ptr"]), f"{what}: ptr"
    if exact_x:
        assert torch.equal(out.x.cpu(), ref["x"]), f"{what}: x"
    else:
        assert_close(out.x, ref["x"], tol=2e-6, what=f"{what}: x")


def test_mesh_batch_vs_reference_golden(dc, golden_dir):
    gold = torch.load(f"{golden_dir}/assemble.pt")
    vs, ts = [m[0] for m in gold["meshes"]], [m[1] for m in gold["meshes"]]
    out = dc.mesh_batch(vs, ts)                        # fp64 vertices, int32 triangles: what Open3D hands over
    _same_batch(out, gold["soft"], what="mesh_batch")
    assert out._ptr_host == gold["soft"]["ptr"].tolist()
    out64 = dc.mesh_batch([v.numpy() for v in vs], [t.long().numpy() for t in ts], encode=False)
    _same_batch(out64, gold["soft_raw"], exact_x=True, what="mesh_batch(encode=False, int64)")
    # the first 3 feature columns are the position itself, bit for bit
    assert torch.equal(out.x[:, :3].cpu(), gold["soft"]["pos"])
    # per-graph access (eval.py:149,158) recovers local indices
    g1 = out[1]
    lo, hi = gold["soft"]["ptr"][1:3].tolist()
    assert g1.pos.shape[0] == hi - lo and int(g1.edge_index.max()) < hi - lo and int(g1.edge_index.min()) >= 0


def test_collider_batch_vs_reference_golden(dc, golden_dir):
    gold = torch.load(f"{golden_dir}/assemble.pt")
    out = dc.collider_batch(gold["centers"], gold["force_vec"], gold["force"], radius=0.05)
    _same_batch(out, gold["rigid"], what="collider_batch")
    assert torch.equal(out.x[:, :7].cpu(), gold["rigid"]["x"][:, :7])     # force vector, force, position: exact
    assert out.x.shape[1] == 25 and out.pos.shape[0] == 3 * 762 and out.edge_index.shape[1] == 3 * 4560


@pytest.mark.parametrize("index_dtype", [torch.int64, torch.int32])
def test_batch_from_data_list_vs_oracle(dc, index_dtype):
    datas, odatas = [], []
    for g, n in enumerate((50, 0, 1, 333, 64)):        # ragged, with an empty and a single-node graph
        gen = torch.Generator().manual_seed(100 + g)
        x = torch.randn(n, 21, generator=gen)
        pos = torch.randn(n, 3, generator=gen)
        e = 0 if n == 0 else int(torch.randint(0, 4 * n + 1, (1,), generator=gen))
        ei = torch.randint(0, max(n, 1), (2, e), generator=gen)
        odatas.append(oracle.Data(x=x, edge_index=ei, pos=pos))
        datas.append(dc.Data(x=x, edge_index=ei.to(index_dtype), pos=pos))
    ref = oracle.Batch.from_data_list(odatas)
    for out in (dc.batch_from_data_list(datas), dc.Batch.from_data_list(datas, device="cuda")):
        assert out.x.is_cuda
        _same_batch(out, dict(x=ref.x, pos=ref.pos, edge_index=ref.edge_index, batch=ref.batch, ptr=ref.ptr),
                    exact_x=True, what="batch_from_data_list")
        assert out._edge_ptr == ref._edge_ptr
        for i in (0, 3, 4):
            a, b = out[i], ref[i]
            assert torch.equal(a.edge_index.cpu(), b.edge_index) and torch.equal(a.x.cpu(), b.x)
    # identical to the host path of our own Batch followed by .to(device)
    host = dc.Batch.from_data_list([dc.Data(x=d.x, edge_index=d.edge_index.long(), pos=d.pos) for d in datas]).to("cuda")
    assert torch.equal(host.edge_index, out.edge_index) and torch.equal(host.batch, out.batch)


def test_batch_from_data_list_edgeless_and_pos_only(dc):
    datas = [dc.Data(pos=torch.rand(5, 3), edge_index=torch.zeros((2, 0), dtype=torch.long)) for _ in range(3)]
    out = dc.batch_from_data_list(datas)
    assert out.x is None and out.edge_index.shape == (2, 0) and out.batch.tolist() == [0] * 5 + [1] * 5 + [2] * 5
    assert out.ptr.tolist() == [0, 5, 10, 15]


def test_assembled_synthetic_batch_equals_oracle_make_batch(dc):
    """The C1-shaped batch assembled from raw per-sample inputs equals the oracle's host-built batch, and the model
    consumes it (edges bit-exact, so the CSR and every hop are the ones the parity tests cover)."""
    rest, rigid, _ = osyn.make_batch(3, 200, 6)
    softs, centers, fvs, fs = [], [], [], []
    for g in range(3):
        r, _, gen = osyn.soft_graph(g, 200, 6)
        ci = int(torch.randint(0, 200, (1,), generator=gen))
        fdir = torch.randn(3, generator=gen)
        fvs.append(fdir / fdir.norm())
        fs.append(torch.rand(1, generator=gen) * osyn.FORCE_MAX / osyn.FORCE_MAX)
        centers.append(r.pos[ci].double())
        softs.append(dc.Data(x=r.x, edge_index=r.edge_index, pos=r.pos))
    sb = dc.batch_from_data_list(softs)
    rb = dc.collider_batch(torch.stack(centers), torch.stack(fvs), torch.cat(fs))
    _same_batch(sb, dict(x=rest.x, pos=rest.pos, edge_index=rest.edge_index, batch=rest.batch, ptr=rest.ptr), exact_x=True,
                what="soft")
    _same_batch(rb, dict(x=rigid.x, pos=rigid.pos, edge_index=rigid.edge_index, batch=rigid.batch, ptr=rigid.ptr), what="rigid")
    torch.manual_seed(0)
    model = dc.load_model(hidden_dim=32).cuda()
    out = model(sb, rb)
    assert out.pos.shape == sb.pos.shape and torch.isfinite(out.pos).all()


def test_staging_reuse_is_safe(dc):
    """Back-to-back assemblies rotate pinned slots; results of earlier batches stay intact."""
    st = dc.Staging(slots=2)
    outs, refs = [], []
    for i in range(5):
        gen = torch.Generator().manual_seed(i)
        x = torch.randn(4000, 21, generator=gen)
        ei = torch.randint(0, 4000, (2, 30000), generator=gen)
        outs.append(dc.batch_from_data_list([dc.Data(x=x, edge_index=ei, pos=x[:, :3].clone())] * 2, staging=st))
        refs.append((torch.cat([x, x]), torch.cat([ei, ei + 4000], 1)))
    torch.cuda.synchronize()
    for o, (x, ei) in zip(outs, refs):
        assert torch.equal(o.x.cpu(), x) and torch.equal(o.edge_index.cpu(), ei)


def test_assemble_abi_errors(dc):
    from deformcontact_b200 import _abi, ops
    with pytest.raises(_abi.DcError):
        ops.edges_offset(torch.zeros((2, 4), dtype=torch.int64), torch.zeros(2, dtype=torch.int64).cuda(),
                         torch.zeros(2, dtype=torch.int64).cuda())                      # CPU tensor: no CPU path
    with pytest.raises(_abi.DcError):
        ops.edges_offset(torch.zeros((2, 4), dtype=torch.int16).cuda(), torch.zeros(2, dtype=torch.int64).cuda(),
                         torch.zeros(2, dtype=torch.int64).cuda())                      # unsupported index width
    with pytest.raises(_abi.DcError):
        dc.batch_from_data_list([dc.Data(x=torch.zeros(1, 2), edge_index=torch.zeros((2, 0), dtype=torch.long))], device="cpu")


def test_graph_batch_features_on_gpu(dc):
    rest, _, _ = osyn.make_batch(3, 150, 5)
    ref = dict(x=rest.x, pos=rest.pos, edge_index=rest.edge_index, batch=rest.batch, ptr=rest.ptr)
    parts = [rest[i] for i in range(3)]
    out = dc.graph_batch([p.pos for p in parts], [p.edge_index for p in parts])
    _same_batch(out, ref, what="graph_batch")
    out32 = dc.graph_batch([p.pos for p in parts], [p.edge_index.int() for p in parts])
    assert torch.equal(out32.edge_index, out.edge_index) and torch.equal(out32.x, out.x)
    # packed pinned form (what bench.py's end-to-end arm feeds)
    local = torch.cat([p.edge_index for p in parts], 1).pin_memory()
    nptr = rest.ptr.clone().pin_memory()
    eptr = torch.tensor(rest._edge_ptr, dtype=torch.long).pin_memory()
    outp = dc.graph_batch_packed(rest.pos.pin_memory(), local, nptr, eptr)
    assert torch.equal(outp.edge_index, out.edge_index) and torch.equal(outp.x, out.x) and torch.equal(outp.batch, out.batch)
    assert torch.equal(outp.x, dc.to_log_freq(outp.pos))      # same kernel arithmetic as dc_posenc, bit for bit
