"""Stand-in for ogb==1.3.6 `AtomEncoder` (TEST INFRASTRUCTURE ONLY; reference requirements.txt:2).

Published behaviour: one nn.Embedding(dim_f, emb_dim) per categorical atom feature,
xavier_uniform_-initialised, summed over the 9 features.  Feature cardinalities are
ogb.utils.features.get_atom_feature_dims() = [119, 5, 12, 12, 10, 6, 6, 2, 2].
Call sites: reference mhnn.py:164, equihnn_egnn.py:121.
"""
import torch

FULL_ATOM_FEATURE_DIMS = [119, 5, 12, 12, 10, 6, 6, 2, 2]


class AtomEncoder(torch.nn.Module):
    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = torch.nn.ModuleList()
        for dim in FULL_ATOM_FEATURE_DIMS:
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)

    def forward(self, x):
        x_embedding = 0
        for i in range(x.shape[1]):
            x_embedding = x_embedding + self.atom_embedding_list[i](x[:, i])
        return x_embedding
