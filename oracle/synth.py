"""ORACLE (test infrastructure only): seeded synthetic inputs shared by oracle/make_golden.py and the
tests (images are regenerated from the seed instead of being stored in the fixtures)."""
import numpy as np
import torch


class ListDataset(torch.utils.data.Dataset):
    def __init__(self, rows, **attrs):
        self.rows = rows
        for k, v in attrs.items():
            setattr(self, k, v)

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        return self.rows[i]


def synthetic_problem(seed=0, n_train=16, n_test=8, n_rsa=8, n_cls=6):
    g = torch.Generator().manual_seed(seed)
    def imgs(n):
        return torch.randn(n, 3, 224, 224, generator=g)
    tr, te, rs = imgs(n_train), imgs(n_test), imgs(n_rsa)
    ttr = torch.randn(n_train, n_cls, generator=g) * 9.5 + 5.75
    tte = torch.randn(n_test, n_cls, generator=g) * 9.5 + 5.75
    rdm = 1 - np.corrcoef(torch.randn(n_rsa, 12, generator=g).numpy())
    np.fill_diagonal(rdm, 0)
    return dict(train_images=tr, train_targets=ttr, test_images=te, test_targets=tte, rsa_images=rs,
                human_rdm=rdm)


PROMPTS = ["metallic; artificial", "food-related", "animal-related", "textile", "plant-related",
           "house-related; furnishing-related"]
