"""N3: batch assembly on the GPU — the step before the message-passing path.

The reference builds every training batch on the host, single-threaded, per step:
``Batch.from_data_list`` three times (train.py:36-38: concat, per-graph index increment, ``batch`` / ``ptr``),
then three ``.to(device)`` (train.py:40-44); per sample ``mesh_to_graph`` walks the triangles in a Python
list comprehension (utils/graph_utils.py:12), ``to_log_freq`` / ``_feature_rigid`` build the features
(utils/graph_utils.py:16, loaders/common.py:6-19) and the collider sphere is created and translated by
Open3D (loaders/common.py:25-30).

Here the host only packs the per-sample arrays back to back (indices stay graph-local) into ONE pinned
staging buffer, issues ONE host-to-device copy, and libdcb200 kernels (``csrc/assemble.cu``) emit the
batched layout: offset ``edge_index``, ``batch``, mesh half-edges of all graphs in one launch, 21-d / 25-d
features, instanced collider spheres.  Three entry points, from drop-in to leanest:

* ``batch_from_data_list(data_list)``      = ``Batch.from_data_list(data_list).to(device)`` for CPU ``Data`` lists
* ``mesh_batch(vertices, triangles)``      = ``from_data_list([mesh_to_graph(m) for m in meshes]).to(device)``
* ``collider_batch(centers, vectors, forces)`` = the rigid branch of ``EverydayDeformDataset.__getitem__`` + collate

All results are bit-identical to the host formulation (tests/test_gpu_assemble.py).  No CPU fallback.
"""
import torch

from . import _abi, ops
from .data import Batch

_ALIGN = 256


class Staging:
    """Rotating pinned staging buffers.  A slot is reused only after the copy that last read it has finished
    (an event recorded behind that copy), so the host may pack batch n+1 while batch n is still in flight."""

    def __init__(self, slots=2):
        self._buf = [None] * slots
        self._evt = [None] * slots
        self._i = 0

    def take(self, nbytes):
        self._i = (self._i + 1) % len(self._buf)
        if self._evt[self._i] is not None:
            self._evt[self._i].synchronize()
        buf = self._buf[self._i]
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, pin_memory=True)
            self._buf[self._i] = buf
        return buf

    def copied(self):
        e = torch.cuda.Event()
        e.record()
        self._evt[self._i] = e


_default_staging = Staging()


class _Packer:
    """Lays typed arrays out in one byte buffer at 256-byte aligned offsets; the same plan carves the device copy."""

    def __init__(self):
        self.fields, self.total = [], 0

    def add(self, key, dtype, shape):
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        self.fields.append((key, dtype, tuple(int(s) for s in shape), self.total, nbytes))
        self.total = (self.total + nbytes + _ALIGN - 1) // _ALIGN * _ALIGN

    def views(self, buf):
        return {k: buf[off:off + nb].view(dt).view(shape) for k, dt, shape, off, nb in self.fields}


def _upload(packer, fill, device, staging):
    """fill(host_views) writes the staged arrays; returns the device views after ONE async copy."""
    if torch.device(device).type != "cuda":
        raise _abi.DcError("batch assembly runs on a CUDA device (libdcb200 has no CPU path)")
    staging = staging or _default_staging
    hbuf = staging.take(max(packer.total, 1))
    fill(packer.views(hbuf))
    dbuf = hbuf[:max(packer.total, 1)].to(device, non_blocking=True)
    staging.copied()   # on the current stream of the current device: callers run with ``device`` current, like all of ops
    return packer.views(dbuf), packer.total


def _cumsum0(sizes):
    out = [0]
    for s in sizes:
        out.append(out[-1] + int(s))
    return out


def _finish(batch, node_ptr_dev, node_ptr, edge_ptr, nbytes):
    batch.ptr = node_ptr_dev
    batch.batch = ops.batch_vector(node_ptr_dev, node_ptr[-1])
    batch._ptr_host = node_ptr
    batch._edge_ptr = edge_ptr
    batch._h2d_bytes = nbytes
    return batch


def batch_from_data_list(data_list, device="cuda", staging=None):
    """``Batch.from_data_list(data_list).to(device)`` (train.py:36-44) for a list of CPU ``Data`` (ours or PyG's;
    only ``.x / .pos / .edge_index`` are read): one packed copy, index increment and ``batch`` on the GPU."""
    data_list = list(data_list)
    if not data_list:
        raise ValueError("batch_from_data_list: empty list")
    has_x = all(d.x is not None for d in data_list)
    has_pos = all(d.pos is not None for d in data_list)
    sizes = [(d.x.shape[0] if d.x is not None else d.pos.shape[0] if d.pos is not None
              else (int(d.edge_index.max()) + 1 if d.edge_index.numel() else 0)) for d in data_list]
    node_ptr = _cumsum0(sizes)
    edge_ptr = _cumsum0(d.edge_index.shape[1] for d in data_list)
    B, N, E = len(data_list), node_ptr[-1], edge_ptr[-1]
    idt = torch.int32 if all(d.edge_index.dtype == torch.int32 for d in data_list) else torch.int64
    pk = _Packer()
    pk.add("node_ptr", torch.int64, (B + 1,))
    pk.add("edge_ptr", torch.int64, (B + 1,))
    if has_x:
        pk.add("x", data_list[0].x.dtype, (N,) + tuple(data_list[0].x.shape[1:]))
    if has_pos:
        pk.add("pos", data_list[0].pos.dtype, (N,) + tuple(data_list[0].pos.shape[1:]))
    pk.add("ei", idt, (2, E))

    def fill(v):
        v["node_ptr"].copy_(torch.tensor(node_ptr, dtype=torch.int64))
        v["edge_ptr"].copy_(torch.tensor(edge_ptr, dtype=torch.int64))
        if has_x and N:
            torch.cat([d.x for d in data_list], 0, out=v["x"])
        if has_pos and N:
            torch.cat([d.pos for d in data_list], 0, out=v["pos"])
        if E:
            torch.cat([d.edge_index.to(idt) for d in data_list], 1, out=v["ei"])

    dv, nbytes = _upload(pk, fill, device, staging)
    out = Batch(x=dv["x"] if has_x else None, pos=dv["pos"] if has_pos else None,
                edge_index=ops.edges_offset(dv["ei"], dv["edge_ptr"], dv["node_ptr"]))
    return _finish(out, dv["node_ptr"], node_ptr, edge_ptr, nbytes)


def mesh_batch(vertices_list, triangles_list, device="cuda", encode=True, staging=None):
    """Raw meshes -> device ``Batch``: ``from_data_list([mesh_to_graph(m, encode) for m in meshes]).to(device)``
    (utils/graph_utils.py:7-20 + train.py:36-44).  ``vertices``: [n_i, 3] float arrays (fp64 as Open3D holds them,
    or fp32; cast to fp32 like utils/graph_utils.py:10); ``triangles``: [t_i, 3] integer arrays with local indices."""
    verts = [torch.as_tensor(v).to(torch.float32).reshape(-1, 3) for v in vertices_list]
    tris = [torch.as_tensor(t).reshape(-1, 3) for t in triangles_list]
    if len(verts) != len(tris) or not verts:
        raise ValueError("mesh_batch: need one triangle array per vertex array")
    idt = torch.int32 if all(t.dtype == torch.int32 for t in tris) else torch.int64
    node_ptr = _cumsum0(v.shape[0] for v in verts)
    tri_ptr = _cumsum0(t.shape[0] for t in tris)
    B, N, T = len(verts), node_ptr[-1], tri_ptr[-1]
    pk = _Packer()
    pk.add("node_ptr", torch.int64, (B + 1,))
    pk.add("tri_ptr", torch.int64, (B + 1,))
    pk.add("pos", torch.float32, (N, 3))
    pk.add("tri", idt, (T, 3))

    def fill(v):
        v["node_ptr"].copy_(torch.tensor(node_ptr, dtype=torch.int64))
        v["tri_ptr"].copy_(torch.tensor(tri_ptr, dtype=torch.int64))
        if N:
            torch.cat(verts, 0, out=v["pos"])
        if T:
            torch.cat([t.to(idt) for t in tris], 0, out=v["tri"])

    dv, nbytes = _upload(pk, fill, device, staging)
    pos = dv["pos"]
    out = Batch(x=ops.node_features(pos) if encode else pos, pos=pos,
                edge_index=ops.mesh_edges_batched(dv["tri"], dv["tri_ptr"], dv["node_ptr"]))
    return _finish(out, dv["node_ptr"], node_ptr, [3 * t for t in tri_ptr], nbytes)


_template_cache = {}


def _sphere_template(radius, resolution, device):
    from .synthetic import uv_sphere
    key = (float(radius), int(resolution), str(torch.device(device)))
    if key not in _template_cache:
        v, t = uv_sphere(radius, resolution)
        _template_cache[key] = (v.to(device), t.to(device))
    return _template_cache[key]


def collider_batch(centers, force_vectors, forces, radius=0.05, resolution=20, device="cuda", staging=None):
    """The rigid branch of the reference loader for a whole batch (loaders/everyday_deform.py:52-56,
    loaders/common.py:6-30): a ``create_sphere(radius)`` mesh translated to each contact point,
    ``mesh_to_graph`` edges, 25-d features ``[force_vector | force | to_log_freq(pos)]``.
    ``centers`` [B, 3] (fp64 like Open3D, or anything castable), ``force_vectors`` [B, 3], ``forces`` [B]."""
    centers = torch.as_tensor(centers).to(torch.float64).reshape(-1, 3)
    B = centers.shape[0]
    head = torch.cat([torch.as_tensor(force_vectors).to(torch.float32).reshape(B, 3),
                      torch.as_tensor(forces).to(torch.float32).reshape(B, 1)], 1)
    tv, tt = _sphere_template(radius, resolution, device)
    V, T = tv.shape[0], tt.shape[0]
    node_ptr = [g * V for g in range(B + 1)]
    pk = _Packer()
    pk.add("node_ptr", torch.int64, (B + 1,))
    pk.add("centers", torch.float64, (B, 3))
    pk.add("head", torch.float32, (B, 4))

    def fill(v):
        v["node_ptr"].copy_(torch.tensor(node_ptr, dtype=torch.int64))
        v["centers"].copy_(centers)
        v["head"].copy_(head)

    dv, nbytes = _upload(pk, fill, device, staging)
    pos = ops.instance_points(tv, dv["centers"])
    out = Batch(x=ops.node_features(pos, dv["head"], dv["node_ptr"]), pos=pos,
                edge_index=ops.mesh_edges_batched(tt, num_graphs=B, nodes_per_graph=V))
    return _finish(out, dv["node_ptr"], node_ptr, [g * 3 * T for g in range(B + 1)], nbytes)


def graph_batch(positions, edge_indices, device="cuda", encode=True, staging=None):
    """Per-sample point sets + graph-local edge lists -> device ``Batch`` with the features computed on the GPU:
    ``from_data_list([Data(x=to_log_freq(p, 3, 1), edge_index=e, pos=p) ...]).to(device)`` without ever building ``x``
    on the host (21 of the 24 floats per node are derived from ``pos``)."""
    pos_l = [torch.as_tensor(p).to(torch.float32).reshape(-1, 3) for p in positions]
    ei_l = [torch.as_tensor(e).reshape(2, -1) for e in edge_indices]
    if len(pos_l) != len(ei_l) or not pos_l:
        raise ValueError("graph_batch: need one edge list per point set")
    idt = torch.int32 if all(e.dtype == torch.int32 for e in ei_l) else torch.int64
    node_ptr = _cumsum0(p.shape[0] for p in pos_l)
    edge_ptr = _cumsum0(e.shape[1] for e in ei_l)
    B, N, E = len(pos_l), node_ptr[-1], edge_ptr[-1]
    pk = _Packer()
    pk.add("node_ptr", torch.int64, (B + 1,))
    pk.add("edge_ptr", torch.int64, (B + 1,))
    pk.add("pos", torch.float32, (N, 3))
    pk.add("ei", idt, (2, E))

    def fill(v):
        v["node_ptr"].copy_(torch.tensor(node_ptr, dtype=torch.int64))
        v["edge_ptr"].copy_(torch.tensor(edge_ptr, dtype=torch.int64))
        if N:
            torch.cat(pos_l, 0, out=v["pos"])
        if E:
            torch.cat([e.to(idt) for e in ei_l], 1, out=v["ei"])

    dv, nbytes = _upload(pk, fill, device, staging)
    pos = dv["pos"]
    out = Batch(x=ops.node_features(pos) if encode else pos, pos=pos,
                edge_index=ops.edges_offset(dv["ei"], dv["edge_ptr"], dv["node_ptr"]))
    return _finish(out, dv["node_ptr"], node_ptr, edge_ptr, nbytes)


def graph_batch_packed(pos, local_edge_index, node_ptr, edge_ptr, device="cuda", encode=True, node_ptr_host=None,
                       edge_ptr_host=None):
    """Same as ``graph_batch`` for inputs a loader already packed: ``pos`` [N, 3] fp32 and ``local_edge_index``
    [2, E] (graph-local, int32 / int64) in PINNED host memory, ``node_ptr`` / ``edge_ptr`` int64 [B+1] pinned tensors.
    The arrays are copied as they are (no staging pass); offsets, ``batch`` and features are produced on the GPU.
    The four arrays may also already be on the device (static input buffers of a captured step): pass the two offset
    tables as host lists too (``node_ptr_host`` / ``edge_ptr_host``) so that nothing is read back."""
    if torch.device(device).type != "cuda":
        raise _abi.DcError("batch assembly runs on a CUDA device (libdcb200 has no CPU path)")
    d_pos, d_ei, d_np, d_ep = (t.to(device, non_blocking=True) for t in (pos, local_edge_index, node_ptr, edge_ptr))
    out = Batch(x=ops.node_features(d_pos) if encode else d_pos, pos=d_pos, edge_index=ops.edges_offset(d_ei, d_ep, d_np))
    nbytes = sum(t.numel() * t.element_size() for t in (pos, local_edge_index, node_ptr, edge_ptr))
    return _finish(out, d_np, node_ptr.tolist() if node_ptr_host is None else list(node_ptr_host),
                   edge_ptr.tolist() if edge_ptr_host is None else list(edge_ptr_host), nbytes)


def collider_batch_device(centers, head, radius=0.05, resolution=20):
    """``collider_batch`` for parameters that are already on the device (static input buffers of a captured step):
    ``centers`` fp64 [B, 3], ``head`` fp32 [B, 4] = force_vector | force.  No copy, no host read."""
    ops._need(centers, torch.float64, "centers"); ops._need(head, torch.float32, "head")
    dev = centers.device
    B = centers.shape[0]
    tv, tt = _sphere_template(radius, resolution, dev)
    V, T = tv.shape[0], tt.shape[0]
    node_ptr = [g * V for g in range(B + 1)]
    d_np = ops._device_table(("arange", B + 1, V), lambda: torch.tensor(node_ptr, dtype=torch.int64), dev)
    pos = ops.instance_points(tv, centers)
    out = Batch(x=ops.node_features(pos, head, d_np), pos=pos, edge_index=ops.mesh_edges_batched(tt, num_graphs=B, nodes_per_graph=V))
    return _finish(out, d_np, node_ptr, [g * 3 * T for g in range(B + 1)], 0)
