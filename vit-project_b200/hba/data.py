"""HBM-resident input store for the fixed THINGS image sets.

The reference decodes and resizes every JPEG on the training thread in every epoch
(``num_workers=0``, NEW:1123; ThingsDataset.__getitem__, NEW:196-204) although the images are never
augmented (NEW:183-188).  ``ResidentStore`` runs the dataset's own ``__getitem__`` once per image and
keeps the resulting tensors in HBM; ``ResidentLoader`` then iterates exactly like
``DataLoader(dataset, batch_size, shuffle, generator=...)`` — it drives a real DataLoader over the
*indices*, so batch order and the consumption of the shuffle generator (NEW:1119-1124, restored per
epoch by NEW:129-131) are identical — and yields ``(names, images, targets)`` batches that already live
on the device.  ``last_ids`` carries stable per-image integer ids for the frozen-trunk cache.
"""
from __future__ import annotations

import os

import torch
from torch.utils.data import DataLoader, Dataset

_NAME_IDS = {}


def image_id(name: str, namespace: str = "") -> int:
    """Process-wide stable integer id of an image (key of the frozen-trunk cache).  `namespace` is the image
    directory (or a per-store token for synthetic stores): two runs in one worker whose img_dir differs but
    whose file names collide must not share cache entries."""
    return _NAME_IDS.setdefault((namespace, name), len(_NAME_IDS))


class _Indices(Dataset):
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return i


class ResidentStore:
    """All items of `dataset` decoded once and stacked on `device`."""

    def __init__(self, dataset, device, names=None, images=None, targets=None, namespace=None):
        self.device = torch.device(device)
        if namespace is None:
            img_dir = getattr(dataset, "img_dir", None)
            namespace = os.path.abspath(img_dir) if isinstance(img_dir, str) else f"store@{id(self):x}"
        self.namespace = namespace
        if images is None:
            names, imgs, tgts = [], [], []
            for i in range(len(dataset)):
                item = dataset[i]
                names.append(item[0])
                imgs.append(item[1])
                if len(item) > 2:
                    tgts.append(item[2])
            images = torch.stack(imgs)
            targets = torch.stack(tgts) if tgts else None
        self.names = list(names)
        self.images = images.to(self.device, torch.float32).contiguous()
        self.targets = None if targets is None else targets.to(self.device, torch.float32).contiguous()
        self.ids = [image_id(n, namespace) for n in self.names]
        self.ids_dev = torch.tensor(self.ids, dtype=torch.int64, device=self.device)

    def __len__(self):
        return len(self.names)


class ResidentLoader:
    """Drop-in for ``DataLoader(dataset, batch_size=..., shuffle=..., generator=...)`` over a
    ResidentStore.  `index_map` selects a subset of the store (SubsetWithIndices, NEW:164-177)."""

    def __init__(self, store, batch_size, shuffle=False, generator=None, index_map=None, dataset=None):
        self.store = store
        self.index_map = None if index_map is None else torch.as_tensor(list(index_map), dtype=torch.int64)
        n = len(store) if self.index_map is None else len(self.index_map)
        self.dataset = dataset if dataset is not None else _Indices(n)
        self._index_loader = DataLoader(_Indices(n), batch_size=batch_size, shuffle=shuffle,
                                        generator=generator)
        self.batch_size = batch_size
        self.last_ids = None
        self.last_ids_dev = None   # the same ids as a device tensor (no host round trip)

    def __len__(self):
        return len(self._index_loader)

    def __iter__(self):
        st = self.store
        # The whole epoch's batches are drawn up front - the same DataLoader, hence the same consumption
        # of the shuffle generator as iterating lazily - so that the index tensors cross to the device in
        # ONE copy per epoch instead of one small pageable copy per step (the cached, graph-replayed step
        # leaves ~1 ms of host time per batch).
        batches = list(self._index_loader)
        if not batches:
            return
        if self.index_map is not None:
            batches = [self.index_map[idx] for idx in batches]
        all_dev = torch.cat(batches).to(st.device)
        off = 0
        for idx in batches:
            host = idx.tolist()
            dev = all_dev[off:off + len(host)]
            off += len(host)
            self.last_ids = [st.ids[i] for i in host]
            self.last_ids_dev = st.ids_dev.index_select(0, dev)
            names = [st.names[i] for i in host]
            images = st.images.index_select(0, dev)
            if st.targets is None:
                yield names, images
            else:
                yield names, images, st.targets.index_select(0, dev)
        self.last_ids = self.last_ids_dev = None
