"""Golden fixture for N3 (on-GPU batch assembly), produced by running the REFERENCE's own code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_assemble.py

Executed unmodified from /root/reference:
  * utils/graph_utils.py:mesh_to_graph  (triangle walk, fp32 cast, to_log_freq) on duck-typed mesh objects
    (``.vertices`` / ``.triangles`` numpy arrays, what ``np.asarray(o3d_mesh.vertices)`` yields: fp64 / int32)
  * loaders/common.py:_feature_rigid    (imported with a stub ``open3d`` module; the function is pure torch)
``torch_geometric.data.Data`` / ``Batch`` are the oracle's restatement (real PyG is not installable here), so the
concatenation / index increment / ``batch`` / ``ptr`` part is pinned by the oracle, not by the reference.
-> assemble.pt
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

import oracle  # noqa: E402
from make_golden import _inject_pyg  # noqa: E402


def main():
    sys.path.insert(0, REF)
    _inject_pyg()
    sys.modules["open3d"] = types.ModuleType("open3d")
    from utils.graph_utils import mesh_to_graph as ref_mesh_to_graph
    from loaders.common import _feature_rigid as ref_feature_rigid

    gen = torch.Generator().manual_seed(21)
    meshes = []
    for nx, ny in ((7, 5), (2, 2), (12, 9)):                      # ragged open sheets
        v, t = oracle.grid_mesh(nx, ny, jitter=0.01, generator=gen)
        meshes.append((v.double().numpy() + 0.1234567891, t.to(torch.int32).numpy()))
    sv, st = oracle.uv_sphere(0.05, 20)
    centers = (torch.rand(3, 3, generator=gen, dtype=torch.float64) - 0.5)
    colliders = [((sv + c).numpy(), st.to(torch.int32).numpy()) for c in centers]   # create_sphere + translate (fp64)
    force_vec = torch.randn(3, 3, generator=gen)
    force = [float(f) for f in torch.rand(3, generator=gen, dtype=torch.float64)]

    soft = [ref_mesh_to_graph(types.SimpleNamespace(vertices=v, triangles=t)) for v, t in meshes]
    soft_raw = [ref_mesh_to_graph(types.SimpleNamespace(vertices=v, triangles=t), encode=False) for v, t in meshes]
    rigid = []
    for (v, t), fv, f in zip(colliders, force_vec, force):
        g = ref_mesh_to_graph(types.SimpleNamespace(vertices=v, triangles=t))
        g.x = ref_feature_rigid({"force_vector": fv, "force": f}, g.x)        # loaders/everyday_deform.py:56
        rigid.append(g)
    bs, bsr, br = (oracle.Batch.from_data_list(l) for l in (soft, soft_raw, rigid))
    pack = lambda b: {"x": b.x, "pos": b.pos, "edge_index": b.edge_index, "batch": b.batch, "ptr": b.ptr}
    torch.save({"meshes": [(torch.from_numpy(v), torch.from_numpy(t)) for v, t in meshes],
                "centers": centers, "force_vec": force_vec, "force": torch.tensor(force, dtype=torch.float64),
                "soft": pack(bs), "soft_raw": pack(bsr), "rigid": pack(br)}, os.path.join(HERE, "assemble.pt"))
    print("assemble.pt", os.path.getsize(os.path.join(HERE, "assemble.pt")), "soft", tuple(bs.x.shape), "rigid", tuple(br.x.shape))


if __name__ == "__main__":
    main()
