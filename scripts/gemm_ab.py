"""A/B of two builds of the K2 kernel on the same box, interleaved (controls for the clocks of the box):
python scripts/gemm_ab.py libA.so libB.so  -> TFLOP/s of dc_gemm (tcgen05 3xTF32) per shape and library, three rounds."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deformcontact_b200 import _abi

libs = []
for path in sys.argv[1:]:
    L = C.CDLL(os.path.abspath(path))
    L.dc_gemm.restype, L.dc_gemm.argtypes = _abi.PROTOTYPES["dc_gemm"]
    L.dc_gemm_workspace_bytes.restype, L.dc_gemm_workspace_bytes.argtypes = _abi.PROTOTYPES["dc_gemm_workspace_bytes"]
    libs.append((os.path.basename(path), L))
g = torch.Generator(device="cuda").manual_seed(0)
shapes = [(512000, 256, 1024, 0, 1, "layer fwd"), (512000, 1024, 256, 0, 0, "layer dX"), (256, 1024, 512000, 1, 0, "layer dW"),
          (8000, 3048, 256, 0, 1, "scores"), (8000, 256, 3048, 0, 0, "attn @ Xr")]
st = torch.cuda.current_stream().cuda_stream
for M, N, K, ta, tb, label in shapes:
    A = torch.randn((K, M) if ta else (M, K), generator=g, device="cuda")
    B = torch.randn((N, K) if tb else (K, N), generator=g, device="cuda")
    out = torch.empty(M, N, device="cuda")
    seg = (_abi.GemmSeg * 1)(_abi.GemmSeg(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), K))
    res = {n: [] for n, _ in libs}
    for rnd in range(3):
        for name, L in libs:
            nb = L.dc_gemm_workspace_bytes(M, N, K, ta, tb)
            ws = torch.empty(max(nb, 1), dtype=torch.uint8, device="cuda")
            run = lambda: L.dc_gemm(seg, 1, ta, tb, M, N, out.data_ptr(), out.stride(0), None, 0, 0, _abi.GEMM_TF32X3, ws.data_ptr(), nb, st)
            for _ in range(2):
                assert run() == 0
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                run()
            e1.record(); torch.cuda.synchronize()
            res[name].append(2.0 * M * N * K / (e0.elapsed_time(e1) / 5) / 1e9)
    print(f"{label:10s} {M}x{N}x{K}: " + " | ".join(f"{n} " + "/".join(f"{v:.1f}" for v in vs) for n, vs in res.items()) + " TF/s", flush=True)
