"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, torch
sys.path.insert(0, ".")
import deformcontact_b200 as dc
from deformcontact_b200 import ops, synthetic
torch.manual_seed(0)
rest, rigid, deformed = synthetic.make_batch(2, 300, 8)
for backbone in ("TAGConv", "GCNConv", "GATConv", "MPNN"):
    m = dc.load_model(hidden_dim=64, backbone=backbone, attn_group=2).cuda()
    loss, _, _ = dc.train_step_loss(m, rest, rigid, deformed)
    loss.backward()
pos = torch.rand(700, 3, device="cuda")
ptr3 = torch.tensor([0, 100, 100, 700], device="cuda")
dc.knn_graph(pos, 40); dc.knn_graph(pos, 16); dc.knn_graph(pos, 70, loop=True); dc.knn_graph(pos, 8, ptr=ptr3)
dc.radius_graph(pos, 0.2); dc.radius_graph(pos, 0.3, ptr=ptr3, max_num_neighbors=5)
ops.KNN_MODE = "grid"   # K4g: uniform-grid search (bounding box, counting sort, ring walk), kNN and radius
dc.knn_graph(pos, 16); dc.knn_graph(pos, 70, loop=True); dc.radius_graph(pos, 0.2); dc.radius_graph(pos, 0.3, max_num_neighbors=5)
dc.knn_graph(torch.full((50, 3), 0.5, device="cuda"), 4)
dc.knn_graph(pos, 8, ptr=ptr3); dc.knn_graph(pos, 40, loop=True, ptr=ptr3)   # one grid per graph (dc_knn_grid_batched), empty graph in between
dc.radius_graph(pos, 0.3, ptr=ptr3, max_num_neighbors=5); dc.radius_graph(pos, 0.2, ptr=ptr3, loop=True)   # dc_radius_grid_batched
ops.KNN_MODE = "auto"
# relabelled large single graph (ops.REORDER): cell order from the grid search, permuted hops, TAGConv fwd + bwd
big = torch.rand(ops.REORDER_MIN_NODES + 100, 3, device="cuda")
ei_big = dc.knn_graph(big, 6)
xb = torch.randn(big.shape[0], 32, device="cuda", requires_grad=True)
dc.TAGConv(32, 32).cuda()(xb, ei_big, relu=True).sum().backward()
dc.GCNConv(32, 32).cuda()(xb.detach(), ei_big)
ops.cell_order(big)
# N3 batch assembly: every entry point, ragged inputs, int32 and int64 local indices
parts = [rest[i] for i in range(2)]
cpu = lambda t: t.cpu()
dc.batch_from_data_list([dc.Data(x=cpu(p.x), edge_index=cpu(p.edge_index), pos=cpu(p.pos)) for p in parts])
dc.graph_batch([cpu(p.pos) for p in parts], [cpu(p.edge_index).int() for p in parts])
sv, st = synthetic.uv_sphere()
dc.mesh_batch([sv.numpy(), sv[:400].numpy()], [st.int().numpy(), st[st.max(1).values < 400].numpy()])
dc.collider_batch(torch.rand(3, 3, dtype=torch.float64), torch.randn(3, 3), torch.rand(3))
# K1 v9 hop chain: forward into a strided buffer and transposed in place, ragged tiles
gc = ops.GraphCSR(rest.edge_index, rest.x.shape[0], "tag", rest._ptr_host)
hx = torch.randn(rest.x.shape[0], 64, device="cuda"); hb = torch.zeros(rest.x.shape[0], 192, device="cuda")
ops.propagate_chain(gc, [(hx, None, hb[:, :64]), (hb[:, :64], None, hb[:, 64:128]), (hb[:, 64:128], None, hb[:, 128:])])
ops.propagate_chain(gc, [(hx, hb[:, :64], hb[:, :64]), (hb[:, :64], hb[:, 64:128], hb[:, 64:128])], transpose=True)
# K1 v10 staged chain (TMA -> shared memory): 7-lane and 8-lane tiles, narrow layer-1 width, transposed in place
for Fs in (64, 24):
    hx = torch.randn(rest.x.shape[0], Fs, device="cuda"); hb = torch.zeros(rest.x.shape[0], 2 * Fs, device="cuda")
    ops.spmm_chain(gc.rowptr, gc.edges, None, [(hx, None, hb[:, :Fs]), (hb[:, :Fs], None, hb[:, Fs:])], tile_ptr=gc.tile_ptr,
                   n_tiles=gc.n_tiles, max_tile_rows=gc._max_tile)
    ops.spmm_chain(gc.t[0], gc._edges_t, None, [(hx, hb[:, :Fs], hb[:, :Fs])], tile_ptr=gc.tile_ptr, n_tiles=gc.n_tiles,
                   max_tile_rows=gc._max_tile)
# K1 v11 stream chain (cp.async record ring, two-batch gather window): forward and transposed in place, ragged tiles
ops.K1_STREAM = 1
hx = torch.randn(rest.x.shape[0], 64, device="cuda"); hb = torch.zeros(rest.x.shape[0], 192, device="cuda")
ops.spmm_chain(gc.rowptr, gc.edges, None, [(hx, None, hb[:, :64]), (hb[:, :64], None, hb[:, 64:128]), (hb[:, 64:128], None, hb[:, 128:])],
               tile_ptr=gc.tile_ptr, n_tiles=gc.n_tiles)
ops.spmm_chain(gc.t[0], gc._edges_t, None, [(hx, hb[:, :64], hb[:, :64]), (hb[:, :64], hb[:, 64:128], hb[:, 64:128])], tile_ptr=gc.tile_ptr,
               n_tiles=gc.n_tiles)
ops.K1_STREAM = 0
# float4 column sums
ops.colsum(torch.randn(3000, 256, device="cuda")); ops.relu_bwd_colsum(torch.randn(3000, 256, device="cuda").relu(), torch.randn(3000, 256, device="cuda"))
# fused ReLU backward + column sums; the captured train step (graph replay)
ops.relu_bwd_colsum(torch.randn(3000, 100, device="cuda").relu(), torch.randn(3000, 100, device="cuda"))
mg = dc.load_model(hidden_dim=64, attn_group=2).cuda()
og = torch.optim.Adam(mg.parameters(), lr=4e-4, capturable=True)
cs = dc.CapturedTrainStep(mg, og, rest, rigid, deformed, warmup=1)
cs.run(rest, rigid, deformed)
x = torch.randn(rest.x.shape[0], 256, device="cuda", requires_grad=True)
for variant in ("generic", "tiled", "tiled8", "tiled_prefetch", "smem", "lean", "blocks", "auto"):
    ops.K1_VARIANT = variant
    ops.clear_csr_cache()
    layer = dc.TAGConv(256, 256, precision=ops.GEMM_PREFER_TC).cuda()
    layer(x, rest.edge_index, relu=True, ptr=rest._ptr_host).sum().backward()
A = torch.randn(3000, 256, device="cuda"); B = torch.randn(3000, 64, device="cuda")
ops.gemm([(A, B)], 256, 64, True, False, precision=ops.GEMM_TF32X3)
# K2 v2: every layout, column chunks, ragged K, split-K slabs, batched launch
for ta in (False, True):
    for tb in (False, True):
        M, N, K = 300, 520, 200
        A = torch.randn((K, M) if ta else (M, K), device="cuda"); B = torch.randn((N, K) if tb else (K, N), device="cuda")
        ops.gemm([(A, B)], M, N, ta, tb, precision=ops.GEMM_TF32X3)
        ops.gemm_batched([(A, B, torch.empty(M, N, device="cuda")), (A, B, torch.empty(M, N, device="cuda"))], ta, tb)
ops.gemm([(torch.randn(130, 3048, device="cuda"), torch.randn(64, 3048, device="cuda"))], 130, 64, False, True, precision=ops.GEMM_TF32X3)
torch.cuda.synchronize()
print("sanitize run ok")
