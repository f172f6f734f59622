import torch


class Data(object):
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kw):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos = x, edge_index, edge_attr, y, pos
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith("_")]

    def __getitem__(self, k):
        return getattr(self, k)

    def __setitem__(self, k, v):
        setattr(self, k, v)

    @property
    def num_nodes(self):
        return self.x.size(0)

    def cat_dim(self, key, item=None):
        return -1 if "index" in key and item is not None and item.dim() == 2 else 0

    def contiguous(self):
        return self

    def to(self, device):
        for k in self.keys:
            if torch.is_tensor(self[k]):
                self[k] = self[k].to(device)
        return self


class Batch(Data):
    pass


class InMemoryDataset(torch.utils.data.Dataset):
    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        self.root, self.transform = root, transform


class DataLoader(torch.utils.data.DataLoader):
    pass
