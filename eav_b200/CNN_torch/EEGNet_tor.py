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
from .._module_base import ArenaAdam, ArenaModule, FusedTrainerMixin


class EEGNet_tor(ArenaModule):
    _VARIANT = EAV_VARIANT_TOR
    _BN_NAMES = ("firstBN", "depthwiseBN", "separableBN")

    def __init__(self, nb_classes, Chans=30, Samples=500, dropoutRate=0.5, kernLength=300, F1=8, D=8, F2=64,
                 norm_rate=1.0, dropoutType='Dropout'):
        super().__init__()
        G, n_feat = F1 * D, F2 * (Samples // 4 // 8)
        # Attribute name -> module, registered in the reference's order (EEGNet_tor.py:21-44): the state_dict keys
        # are the attribute names, and the parameterised layers draw their default initialisation from torch's
        # global RNG in exactly this order.
        layers = (
            ("dropout", lambda: (nn.Dropout if dropoutType == 'Dropout' else nn.Dropout2d)(dropoutRate)),
            ("firstConv", lambda: nn.Conv2d(1, F1, (1, kernLength), padding="same", bias=False)),
            ("firstBN", lambda: nn.BatchNorm2d(F1)),
            ("elu", nn.ELU),
            ("depthwiseConv", lambda: nn.Conv2d(F1, G, (Chans, 1), groups=F1, bias=False)),
            ("depthwiseBN", lambda: nn.BatchNorm2d(G)),
            ("depthwisePool", lambda: nn.AvgPool2d((1, 4))),
            ("separableConv", lambda: nn.Conv2d(G, F2, (1, 16), padding="same", bias=False)),
            ("separableBN", lambda: nn.BatchNorm2d(F2)),
            ("separablePool", lambda: nn.AvgPool2d((1, 8))),
            ("flatten", nn.Flatten),
            ("dense", lambda: nn.Linear(n_feat, nb_classes)),
            ("softmax", lambda: nn.Softmax(dim=1)),
        )
        for name, ctor in layers:
            setattr(self, name, ctor())
        self._dims = EegnetDims(nb_classes, Chans=Chans, Samples=Samples, dropoutRate=dropoutRate,
                                kernLength=kernLength, F1=F1, D=D, F2=F2, norm_rate=norm_rate,
                                variant=EAV_VARIANT_TOR, dropout2d=dropoutType != 'Dropout')
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
        self.lr, self.batch_size, self.num_epochs = lr, batch_size, num_epochs
        self.tr_x, self.tr_y, self.te_x, self.te_y = data
        self.train_dataloader = self._prepare_dataloader(self.tr_x, self.tr_y, shuffle=True)
        self.test_dataloader = self._prepare_dataloader(self.te_x, self.te_y, shuffle=False)
        self.model = model
        self.criterion = nn.CrossEntropyLoss()
        self.optimizer = ArenaAdam(model.parameters(), lr=lr)       # optim.Adam whose state is the fused trainer's
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("eav_b200.Trainer_uni needs a CUDA (B200) device; there is no CPU fallback")
        # nn.DataParallel (EEGNet_tor.py:86-88) is deliberately NOT used: multi-GPU is by subject
        # sharding, one process per GPU (eav_b200.sharding), never by splitting one tiny batch.
        self.model.to(self.device)
        self._setup_fused(self.model, self.tr_x, self.tr_y, self.te_x, self.te_y)

    def _prepare_dataloader(self, x, y, shuffle=False):
        """float32 features / int64 labels -> DataLoader(batch_size, shuffle) (EEGNet_tor.py:91-94); numpy arrays
        and tensors are both accepted."""
        def as_t(a, dt):
            return a.to(dt) if torch.is_tensor(a) else torch.as_tensor(np.asarray(a), dtype=dt)
        return DataLoader(TensorDataset(as_t(x, torch.float32), as_t(y, torch.long)), batch_size=self.batch_size,
                          shuffle=shuffle)

    def train(self):
        self.model.train()  # once, outside the epoch loop -- as the reference (EEGNet_tor.py:97)
        steps_per_epoch = len(self.train_dataloader)
        for ep in range(1, self.num_epochs + 1):
            for step, rows in enumerate(self._index_batches(train=True)):
                loss = self._fused_train_step(rows)
                if step % 100 == 0:      # the reference's print cadence and format (EEGNet_tor.py:112-113)
                    print(f"Epoch [{ep}/{self.num_epochs}], Step [{step}/{steps_per_epoch}], Loss: {loss.item():.4f}")
            if self.test_dataloader:
                self.validate()

    def validate(self):
        """(mean loss, accuracy) over the test set; leaves the model in eval mode like the reference (F5)."""
        self.model.eval()
        loss_sum, n_correct, n_batches = self._fused_validate()
        mean_loss, acc = loss_sum / n_batches, n_correct / len(self.test_dataloader.dataset)
        print(f"Validation - Loss: {mean_loss:.4f}, Accuracy: {acc:.4f}")
        return mean_loss, acc


def _prepare_dataloader(self, x, y, shuffle=False):
    """Module-level twin of Trainer_uni._prepare_dataloader: the reference file defines one too (EEGNet_tor.py:138-141)."""
    return Trainer_uni._prepare_dataloader(self, x, y, shuffle)
