"""Drop-in for the reference's CNN_torch/EEGNet_tor.py: `EEGNet_tor` (EEGNet_tor.py:15-67)
and `Trainer_uni` (EEGNet_tor.py:69-135) with unchanged constructor / forward / trainer
signatures, submodule names and state_dict keys, running on hand-written sm_100a kernels
(libeav_b200.so).  There is no CPU fallback: forward() on a CPU tensor raises.

Differences from the shipped file are only the ones SURVEY.md section 0 shows are required
for it to run at all: no import of the missing `Fusion` package (F1), DataLoader imported
(F2), max-norm "hooks" implemented with their intended semantics inside the kernels
(forward uses W_old, then W <- renorm(W); backward sees W_new: F3/F4), plus supersets:
3-D (B, Chans, Samples) input is accepted like CNN_EEG.EEGNet does (F10).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader, TensorDataset

from .. import _lib
from ..ops import EegnetDims, EegnetEngine
from .._lib import EAV_VARIANT_TOR
from .._module_base import ArenaModule, FusedTrainerMixin


class EEGNet_tor(ArenaModule):
    _VARIANT = EAV_VARIANT_TOR
    _BN_NAMES = ("firstBN", "depthwiseBN", "separableBN")

    def __init__(self, nb_classes, Chans=30, Samples=500, dropoutRate=0.5, kernLength=300, F1=8, D=8, F2=64,
                 norm_rate=1.0, dropoutType='Dropout'):
        super(EEGNet_tor, self).__init__()
        # same submodules, same construction order as the reference (EEGNet_tor.py:21-44), so
        # default initialisation consumes the RNG identically and state_dict keys match
        self.dropout = nn.Dropout(dropoutRate) if dropoutType == 'Dropout' else nn.Dropout2d(dropoutRate)
        self.firstConv = nn.Conv2d(1, F1, (1, kernLength), padding='same', bias=False)
        self.firstBN = nn.BatchNorm2d(F1)
        self.elu = nn.ELU()
        self.depthwiseConv = nn.Conv2d(F1, F1 * D, (Chans, 1), groups=F1, padding=0, bias=False)
        self.depthwiseBN = nn.BatchNorm2d(F1 * D)
        self.depthwisePool = nn.AvgPool2d((1, 4))
        self.separableConv = nn.Conv2d(F1 * D, F2, (1, 16), padding='same', bias=False)
        self.separableBN = nn.BatchNorm2d(F2)
        self.separablePool = nn.AvgPool2d((1, 8))
        self.flatten = nn.Flatten()
        self.dense = nn.Linear(F2 * ((Samples // 4 // 8)), nb_classes)
        self.softmax = nn.Softmax(dim=1)
        self._dims = EegnetDims(nb_classes, Chans=Chans, Samples=Samples, dropoutRate=dropoutRate,
                                kernLength=kernLength, F1=F1, D=D, F2=F2, norm_rate=norm_rate,
                                variant=EAV_VARIANT_TOR)
        self._dropout2d = dropoutType != 'Dropout'
        self._param_modules = ("firstConv.weight", "firstBN.weight", "firstBN.bias", "depthwiseConv.weight",
                               "depthwiseBN.weight", "depthwiseBN.bias", "separableConv.weight",
                               "separableBN.weight", "separableBN.bias", "dense.weight", "dense.bias")

    def forward(self, x):
        """x: (B, 1, Chans, Samples) or (B, Chans, Samples) float32 CUDA -> (B, nb_classes) probabilities."""
        return self._forward_cuda(x)


class Trainer_uni(FusedTrainerMixin):
    """EEGNet_tor.py:69-135.  Same attributes and control flow, including the reference's
    quirk that only epoch 1 runs in train mode (validate() leaves the model in eval mode,
    SURVEY F5).  The dataset stays resident on the GPU; the DataLoaders are kept as the
    index / permutation source so the RNG consumption order matches the reference."""

    def __init__(self, model, data, lr=1e-4, batch_size=32, num_epochs=10, device=None):
        self.lr = lr
        self.batch_size = batch_size
        self.num_epochs = num_epochs

        self.tr_x, self.tr_y, self.te_x, self.te_y = data
        self.train_dataloader = self._prepare_dataloader(self.tr_x, self.tr_y, shuffle=True)
        self.test_dataloader = self._prepare_dataloader(self.te_x, self.te_y, shuffle=False)

        self.model = model
        self.criterion = nn.CrossEntropyLoss()
        self.optimizer = optim.Adam(self.model.parameters(), lr=self.lr)

        self.device = torch.device(device) if device else torch.device("cuda" if torch.cuda.is_available() else "cpu")
        if self.device.type != "cuda":
            raise RuntimeError("eav_b200.Trainer_uni needs a CUDA (B200) device; there is no CPU fallback")
        # nn.DataParallel (EEGNet_tor.py:86-88) is deliberately NOT used: multi-GPU is by subject
        # sharding, one process per GPU (eav_b200.sharding), never by splitting one tiny batch.
        self.model.to(self.device)
        self._setup_fused(self.model, self.tr_x, self.tr_y, self.te_x, self.te_y)

    def _prepare_dataloader(self, x, y, shuffle=False):
        dataset = TensorDataset(torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32),
                                torch.as_tensor(np.asarray(y) if not torch.is_tensor(y) else y, dtype=torch.long))
        dataloader = DataLoader(dataset, batch_size=self.batch_size, shuffle=shuffle)
        return dataloader

    def train(self):
        self.model.train()  # once, outside the epoch loop -- as the reference (EEGNet_tor.py:97)
        for epoch in range(self.num_epochs):
            n_batches = len(self.train_dataloader)
            for batch_idx, rows in enumerate(self._index_batches(train=True)):
                loss = self._fused_train_step(rows)
                if batch_idx % 100 == 0:
                    print(f"Epoch [{epoch+1}/{self.num_epochs}], Step [{batch_idx}/{n_batches}], Loss: {loss.item():.4f}")
            if self.test_dataloader:
                self.validate()

    def validate(self):
        self.model.eval()
        total_loss, total_correct, n_batches = self._fused_validate()
        avg_loss = total_loss / n_batches
        accuracy = total_correct / len(self.test_dataloader.dataset)
        print(f"Validation - Loss: {avg_loss:.4f}, Accuracy: {accuracy:.4f}")
        return avg_loss, accuracy


def _prepare_dataloader(self, x, y, shuffle=False):   # module-level duplicate kept for signature parity (EEGNet_tor.py:138-141)
    dataset = TensorDataset(torch.tensor(x, dtype=torch.float32), torch.tensor(y, dtype=torch.long))
    return DataLoader(dataset, batch_size=self.batch_size, shuffle=shuffle)
