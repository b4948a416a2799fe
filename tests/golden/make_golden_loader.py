"""Golden fixture for A1: the model the REFERENCE builds from its own config.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_loader.py

Executed unmodified from /root/reference: configs/config.py:Config on configs/everyday.json and
models/model_loader.py:load_model (-> models/model.py:GraphNet.__init__), with the oracle's restated convs injected as
``torch_geometric.nn`` (the real PyG is not installable here; the conv parameter names / shapes are the oracle's restatement
of PyG 2.5.2).  Records the network section of the config, every state-dict key with its shape, and the parameter count.
-> loader.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from make_golden import _inject_pyg  # noqa: E402


def main():
    sys.path.insert(0, REF)
    _inject_pyg()
    from configs.config import Config
    from models.model_loader import load_model
    out = {}
    for backbone in ("TAGConv", "GCNConv", "GATConv"):
        cfg = Config(os.path.join(REF, "configs", "everyday.json"), updates={"network": {"backbone": backbone}})
        model = load_model(cfg)
        out[backbone] = {"network": dict(cfg.network.__dict__),
                         "state_dict": {k: list(v.shape) for k, v in model.state_dict().items()},
                         "num_parameters": sum(p.numel() for p in model.parameters()),
                         "decoder": [type(m).__name__ for m in model.decoder]}
    json.dump(out, open(os.path.join(HERE, "loader.json"), "w"), indent=1, sort_keys=True)
    print({k: (v["num_parameters"], len(v["state_dict"])) for k, v in out.items()})


if __name__ == "__main__":
    main()
