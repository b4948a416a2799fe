"""Minimal ``Data`` / ``Batch`` restating the slice of PyG the reference uses.

Reference call sites: ``utils/graph_utils.py:20`` (``Data(x=, edge_index=, pos=)``),
``train.py:36-44`` (``Batch.from_data_list(...).to(device)``),
``models/model.py:69,75,91-93`` (``.x``, ``.edge_index``, ``.clone()``, ``.pos``),
``eval.py:149,158`` (``batch[i]``).  Semantics follow torch_geometric 2.5.2
``data/batch.py`` + ``data/collate.py``  [3P, restated from the published
algorithm]: node tensors are concatenated on dim 0, ``edge_index`` on dim 1 with
each graph's indices incremented by the cumulative node count, plus
``batch`` (int64 graph id per node) and ``ptr`` (int64 [B+1]).
"""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, pos=None, **kw):
        self.x, self.edge_index, self.pos = x, edge_index, pos
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def num_nodes(self):
        for t in (self.x, self.pos):
            if t is not None:
                return t.shape[0]
        return int(self.edge_index.max()) + 1 if self.edge_index.numel() else 0

    def _tensor_items(self):
        return [(k, v) for k, v in self.__dict__.items() if isinstance(v, torch.Tensor)]

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        out.__dict__ = {k: (v.clone() if isinstance(v, torch.Tensor) else v)
                        for k, v in self.__dict__.items()}
        return out

    def to(self, *a, **kw):
        for k, v in self._tensor_items():
            setattr(self, k, v.to(*a, **kw))
        return self


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        xs, poss, eis, bs, ptr = [], [], [], [], [0]
        for g, d in enumerate(data_list):
            n = d.num_nodes
            if d.x is not None:
                xs.append(d.x)
            if d.pos is not None:
                poss.append(d.pos)
            eis.append(d.edge_index + ptr[-1])
            bs.append(torch.full((n,), g, dtype=torch.long))
            ptr.append(ptr[-1] + n)
        out = cls(x=torch.cat(xs, 0) if xs else None,
                  edge_index=torch.cat(eis, 1),
                  pos=torch.cat(poss, 0) if poss else None)
        out.batch = torch.cat(bs, 0)
        out.ptr = torch.tensor(ptr, dtype=torch.long)
        out._edge_ptr = [0]
        for d in data_list:
            out._edge_ptr.append(out._edge_ptr[-1] + d.edge_index.shape[1])
        return out

    @property
    def num_graphs(self):
        return self.ptr.numel() - 1

    def __getitem__(self, i):
        lo, hi = int(self.ptr[i]), int(self.ptr[i + 1])
        elo, ehi = self._edge_ptr[i], self._edge_ptr[i + 1]
        return Data(x=None if self.x is None else self.x[lo:hi],
                    edge_index=self.edge_index[:, elo:ehi] - lo,
                    pos=None if self.pos is None else self.pos[lo:hi])
