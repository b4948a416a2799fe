"""``Data`` / ``Batch``: the batch layout the reference feeds its model.

The reference's ``collate_fn`` (loaders/collate.py:4-16) returns tuples of per-sample graphs;
the concatenated layout is produced by ``Batch.from_data_list`` in the loops (train.py:36-38,
eval.py:107-109) and moved with ``.to(device)`` (train.py:40-44).  Layout (PyG 2.5.2
data/batch.py): ``x``/``pos`` concatenated on dim 0, ``edge_index`` on dim 1 with per-graph node
offsets, ``batch`` int64 [sum N], ``ptr`` int64 [B+1].  A real PyG ``Batch`` is accepted anywhere
these are (only ``.x/.edge_index/.pos/.batch/.ptr`` are read).
"""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, pos=None, **kwargs):
        self.x, self.edge_index, self.pos = x, edge_index, pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        if self.x is not None:
            return self.x.shape[0]
        if self.pos is not None:
            return self.pos.shape[0]
        return int(self.edge_index.max()) + 1 if self.edge_index is not None and self.edge_index.numel() else 0

    @property
    def num_edges(self):
        return 0 if self.edge_index is None else self.edge_index.shape[1]

    def _apply(self, fn):
        for k, v in list(self.__dict__.items()):
            if isinstance(v, torch.Tensor):
                setattr(self, k, fn(v))
        return self

    def clone(self):
        out = object.__new__(type(self))
        out.__dict__ = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in self.__dict__.items()}
        return out

    def to(self, *args, **kwargs):
        return self._apply(lambda t: t.to(*args, **kwargs))

    def pin_memory(self):
        return self._apply(lambda t: t.pin_memory())

    def cuda(self, device=None, non_blocking=False):
        return self._apply(lambda t: t.cuda(device, non_blocking=non_blocking))

    def __repr__(self):
        items = ", ".join(f"{k}={list(v.shape)}" for k, v in self.__dict__.items() if isinstance(v, torch.Tensor))
        return f"{type(self).__name__}({items})"


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list, device=None):
        """PyG layout (data/batch.py).  With ``device`` a CUDA device and CPU inputs, the batch is assembled on the
        GPU from one packed copy (``assemble.batch_from_data_list`` = ``from_data_list(...).to(device)``)."""
        data_list = list(data_list)
        on_host = all(t is None or not t.is_cuda for d in data_list for t in (d.x, d.pos, d.edge_index))
        if device is not None and torch.device(device).type == "cuda" and on_host and data_list:
            from .assemble import batch_from_data_list
            return batch_from_data_list(data_list, device=device)
        if device is not None:                       # graphs already on a device (or a CPU target): concatenate, then move
            return cls.from_data_list(data_list).to(device)
        sizes = [d.num_nodes for d in data_list]
        esizes = [d.num_edges for d in data_list]
        ref = next((t for d in data_list for t in (d.x, d.pos, d.edge_index) if t is not None), None)
        dev = ref.device if ref is not None else "cpu"
        ptr = torch.zeros(len(sizes) + 1, dtype=torch.long)
        ptr[1:] = torch.tensor(sizes, dtype=torch.long).cumsum(0)
        out = cls()
        out.x = torch.cat([d.x for d in data_list], 0) if all(d.x is not None for d in data_list) else None
        out.pos = torch.cat([d.pos for d in data_list], 0) if all(d.pos is not None for d in data_list) else None
        offs = ptr[:-1].tolist()
        out.edge_index = torch.cat([d.edge_index + o for d, o in zip(data_list, offs)], 1)
        out.batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes, dtype=torch.long)).to(dev)
        out.ptr = ptr.to(dev)
        out._edge_ptr = [0]
        for e in esizes:
            out._edge_ptr.append(out._edge_ptr[-1] + e)
        return out

    @property
    def num_graphs(self):
        return self.ptr.numel() - 1

    def __getitem__(self, i):
        i = i if i >= 0 else self.num_graphs + i
        lo, hi = int(self.ptr[i]), int(self.ptr[i + 1])
        elo, ehi = self._edge_ptr[i], self._edge_ptr[i + 1]
        return Data(x=None if self.x is None else self.x[lo:hi], edge_index=self.edge_index[:, elo:ehi] - lo,
                    pos=None if self.pos is None else self.pos[lo:hi])

    def to_data_list(self):
        return [self[i] for i in range(self.num_graphs)]


def collate_fn(batch):
    """loaders/collate.py:4-16 — tuple-of-graphs collate; tensor-valued meta keys are stacked."""
    names, rest, deformed, metas, rigid = zip(*batch)
    meta = {}
    for key, v0 in metas[0].items():
        vals = [m[key] for m in metas]
        meta[key] = torch.stack(vals) if isinstance(v0, torch.Tensor) else vals
    return list(names), rest, deformed, meta, rigid
