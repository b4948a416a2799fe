"""Golden fixture for the TRAIN STEP (SURVEY.md 8d metric (i)): the reference's own training loop, executed unmodified.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_train.py

/root/reference/train.py:train(config) runs as it is — Batch.from_data_list x3, model forward, displacement targets,
nn.L1Loss + lambda * GradientConsistencyLoss, backward, Adam(lr = config.training.learning_rate), the validation pass and
the checkpoint save — for one epoch over three synthetic mini-batches of the reference's batch size (4).  What is stubbed is
only what cannot exist here: ``wandb`` (records what the loop logs), ``loaders.dataset_loader.load_dataset`` (returns the
oracle's seeded synthetic samples, collated by the reference's own loaders/collate.py:collate_fn), and ``torch_geometric``
(the oracle's restated convs / Data / Batch; real PyG is not installable).  hidden_dim is reduced to 32 through the
reference's own ``Config(updates=...)`` to keep the fixture small.
-> train_loop.pt : the logged losses of every step, the validation loss, and the weights the loop saved.
"""
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import synthetic  # noqa: E402
from make_golden import _inject_pyg  # noqa: E402

N_TRAIN_BATCHES, N_VAL_BATCHES, BATCH, NODES, K, HIDDEN, SEED = 3, 1, 4, 60, 4, 32, 11


def samples(first, count):
    out = []
    for g in range(first, first + count):
        rest, deformed, gen = synthetic.soft_graph(g, NODES, K)
        ci = int(torch.randint(0, NODES, (1,), generator=gen))
        rigid = synthetic.rigid_graph(rest.pos[ci], gen)
        out.append((f"obj{g}", rest, deformed, {"force": float(g)}, rigid))
    return out


def main():
    sys.path.insert(0, REF)
    _inject_pyg()
    logs = []
    run_dir = tempfile.mkdtemp()
    wandb = types.ModuleType("wandb")
    wandb.log = lambda d: logs.append(dict(d))
    wandb.save = lambda *a, **k: None
    wandb.finish = lambda: None
    wandb.init = lambda *a, **k: None
    wandb.run = types.SimpleNamespace(dir=run_dir)
    sys.modules["wandb"] = wandb
    from loaders.collate import collate_fn
    train_batches = [collate_fn(samples(b * BATCH, BATCH)) for b in range(N_TRAIN_BATCHES)]
    val_batches = [collate_fn(samples(1000 + b * BATCH, BATCH)) for b in range(N_VAL_BATCHES)]
    dl = types.ModuleType("loaders.dataset_loader")
    dl.load_dataset = lambda config: (train_batches, val_batches)
    sys.modules["loaders.dataset_loader"] = dl
    from configs.config import Config
    import train as ref_train
    cfg = Config(os.path.join(REF, "configs", "everyday.json"),
                 updates={"network": {"hidden_dim": HIDDEN}, "training": {"n_epochs": 1}})
    torch.manual_seed(SEED)
    ref_train.train(cfg)
    weights = torch.load(os.path.join(run_dir, "model_weights.pth"))
    steps = [l for l in logs if "tr_loss" in l]
    val = [l["validation_loss"] for l in logs if "validation_loss" in l]
    assert len(steps) == N_TRAIN_BATCHES and len(val) == 1
    torch.save({"meta": dict(n_train_batches=N_TRAIN_BATCHES, n_val_batches=N_VAL_BATCHES, batch=BATCH, nodes=NODES, k=K,
                             hidden=HIDDEN, seed=SEED, lr=cfg.training.learning_rate, lambda_gradient=cfg.training.lambda_gradient),
                "steps": steps, "validation_loss": val[0], "state_dict": weights}, os.path.join(HERE, "train_loop.pt"))
    print("train_loop.pt", os.path.getsize(os.path.join(HERE, "train_loop.pt")), [round(s["tr_loss"], 6) for s in steps], val)


if __name__ == "__main__":
    main()
