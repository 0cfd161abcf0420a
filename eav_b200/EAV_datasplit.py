"""Drop-in for the reference's EAV_datasplit.py (EAVDataSplit, EAV_datasplit.py:7-58).

Pure index logic, kept on the host in numpy with the reference's exact semantics
(bit-exact by construction, pinned by tests/golden/split_*.npz):
  * per class 0..4 (labels outside range(5) are silently dropped), samples keep
    their original order; the first h_idx of a class go to train, the rest to test;
  * features are np.squeeze'd (all singleton dims dropped).
In addition the split INDICES are exposed so a GPU-resident dataset can be gathered by
index without moving data (`get_split_indices`).
"""
import numpy as np
import torch
from torch.utils.data import DataLoader, TensorDataset


class EAVDataSplit:
    N_CLASSES = 5   # "Assuming there are 5 classes" (EAV_datasplit.py:16)

    def __init__(self, x, y, batch_size=32):
        self.x = np.array(x)
        self.y = np.array(y)
        self.batch_size = batch_size

    def _class_indices(self):
        return [np.where(self.y == class_idx)[0] for class_idx in range(self.N_CLASSES)]

    def _split_features_labels(self):
        idx = self._class_indices()
        return [self.x[i] for i in idx], [self.y[i] for i in idx]

    def get_split_indices(self, h_idx=40):
        """(train_idx, test_idx) int64 arrays into (x, y): class-major, time-ascending."""
        idx = self._class_indices()
        tr = np.concatenate([i[:h_idx] for i in idx], axis=0).astype(np.int64)
        te = np.concatenate([i[h_idx:] for i in idx], axis=0).astype(np.int64)
        return tr, te

    def get_split(self, h_idx=40):
        tr, te = self.get_split_indices(h_idx)
        return np.squeeze(self.x[tr]), self.y[tr], np.squeeze(self.x[te]), self.y[te]

    def get_loaders(self):
        train_features, train_labels, test_features, test_labels = self.get_split()
        train_features = torch.Tensor(np.squeeze(train_features))
        test_features = torch.Tensor(np.squeeze(test_features))
        train_labels = torch.Tensor(train_labels).long()
        test_labels = torch.Tensor(test_labels).long()
        loader_train = DataLoader(TensorDataset(train_features, train_labels), batch_size=self.batch_size, shuffle=True)
        loader_test = DataLoader(TensorDataset(test_features, test_labels), batch_size=self.batch_size, shuffle=False)
        return loader_train, loader_test
