"""torch_geometric.data subset: Data and Batch.from_data_list (node offset only on *index* keys)."""
import torch


class Data:
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if torch.is_tensor(v)]

    @property
    def num_nodes(self):
        return self.x.size(0)

    def to(self, device=None, non_blocking=False):
        out = type(self)()
        for k, v in self.__dict__.items():
            setattr(out, k, v.to(device=device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out


class Batch(Data):
    @staticmethod
    def from_data_list(data_list):
        keys = data_list[0].keys
        cat = {k: [] for k in keys}
        batch, offset = [], 0
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = getattr(d, k)
                if "index" in k or "face" in k:
                    v = v + offset
                cat[k].append(v)
            batch.append(torch.full((n,), g, dtype=torch.long))
            offset += n
        out = Batch()
        for k in keys:
            dim = -1 if ("index" in k or "face" in k) else 0
            setattr(out, k, torch.cat(cat[k], dim=dim))
        out.batch = torch.cat(batch)
        out.num_graphs = len(data_list)
        return out
