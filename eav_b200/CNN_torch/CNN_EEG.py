"""Drop-in for the reference's CNN_torch/CNN_EEG.py: `EEGNet` (CNN_EEG.py:7-67) and
`EEGNetTrainer` (CNN_EEG.py:70-162) on the same sm_100a kernels as EEGNet_tor, selected
with the variant flag (no ELU between BN1 and the depthwise conv, block 2 = depthwise
temporal + pointwise conv, logits out, no max-norm).  state_dict keys are the reference's
(`block1.{0,1,2,3}.*`, `block2.{0,1,2}.*`, `classifier.*`).  No CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader, TensorDataset

from ..ops import EegnetDims
from .._lib import EAV_VARIANT_CNN
from .._module_base import ArenaModule, FusedTrainerMixin


class EEGNet(ArenaModule):
    _VARIANT = EAV_VARIANT_CNN
    _BN_NAMES = ("block1.1", "block1.3", "block2.2")

    def __init__(self, nb_classes, Chans=64, Samples=128, dropoutRate=0.5,
                 kernLength=64, F1=8, D=2, F2=16, norm_rate=0.25):
        super(EEGNet, self).__init__()
        self.Chans = Chans
        self.Samples = Samples
        # identical containers / order as the reference so initialisation and keys match (CNN_EEG.py:20-55)
        self.block1 = nn.Sequential(
            nn.Conv2d(1, F1, (1, kernLength), padding='same', bias=False),
            nn.BatchNorm2d(F1),
            nn.Conv2d(F1, D * F1, (Chans, 1), groups=F1, bias=False),
            nn.BatchNorm2d(D * F1),
            nn.ELU(),
            nn.AvgPool2d((1, 4)),
            nn.Dropout(dropoutRate)
        )
        self.block2 = nn.Sequential(
            nn.Conv2d(D * F1, D * F1, (1, 16), padding='same', groups=D * F1, bias=False),
            nn.Conv2d(D * F1, F2, (1, 1), bias=False),
            nn.BatchNorm2d(F2),
            nn.ELU(),
            nn.AvgPool2d((1, 8)),
            nn.Dropout(dropoutRate)
        )
        self.flatten = nn.Flatten()
        # The reference sizes the classifier with a dry run through block1/block2 in TRAIN mode
        # (CNN_EEG.py:48-53).  Its side effects are part of the initial state and are reproduced
        # by doing the same at construction time (host, once): BatchNorm running_var -> 0.9 and
        # num_batches_tracked -> 1, and two dropout masks drawn from the global RNG BEFORE the
        # classifier is initialised.
        with torch.no_grad():
            n_flatten = self.flatten(self.block2(self.block1(torch.zeros(1, 1, Chans, Samples)))).shape[1]
        assert n_flatten == F2 * (Samples // 4 // 8)
        self.classifier = nn.Linear(n_flatten, nb_classes)
        self._dims = EegnetDims(nb_classes, Chans=Chans, Samples=Samples, dropoutRate=dropoutRate,
                                kernLength=kernLength, F1=F1, D=D, F2=F2, norm_rate=0.0, variant=EAV_VARIANT_CNN)
        self._dropout2d = False
        self._param_modules = ("block1.0.weight", "block1.1.weight", "block1.1.bias", "block1.2.weight",
                               "block1.3.weight", "block1.3.bias", "block2.0.weight", "block2.1.weight",
                               "block2.2.weight", "block2.2.bias", "classifier.weight", "classifier.bias")

    def forward(self, x):
        """x: (Batch, Chans, Samples) or (Batch, 1, Chans, Samples) CUDA float32 -> logits."""
        return self._forward_cuda(x)


def _dataset_tensors(ds):
    if isinstance(ds, TensorDataset):
        return ds.tensors[0], ds.tensors[1]
    xs, ys = zip(*[ds[i] for i in range(len(ds))])
    return torch.stack([torch.as_tensor(x) for x in xs]), torch.as_tensor(ys)


class EEGNetTrainer(FusedTrainerMixin):
    """CNN_EEG.py:70-162: train_epoch() / validate_epoch() / train() / predict() with the
    reference's signatures; model.train() is called every epoch here (no F5 quirk)."""

    def __init__(self, model, train_dataset, val_dataset, batch_size=32, epochs=100, lr=0.001):
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        print(f"Using device: {self.device}")
        if self.device.type != "cuda":
            raise RuntimeError("eav_b200.EEGNetTrainer needs a CUDA (B200) device; there is no CPU fallback")
        self.model = model.to(self.device)
        self.epochs = epochs
        self.batch_size = batch_size
        self.lr = lr
        self.train_loader = DataLoader(train_dataset, batch_size=batch_size, shuffle=True)
        self.test_loader = DataLoader(val_dataset, batch_size=batch_size, shuffle=False)
        self.criterion = nn.CrossEntropyLoss()
        self.optimizer = optim.Adam(model.parameters(), lr=lr)
        trx, try_ = _dataset_tensors(train_dataset)
        tex, tey = _dataset_tensors(val_dataset)
        self._setup_fused(self.model, trx, try_, tex, tey, lr=lr, batch_size=batch_size)

    def train_epoch(self):
        self.model.train()
        running = torch.zeros((), dtype=torch.float64, device=self.device)
        nb = 0
        for rows in self._index_batches(train=True):
            running += self._fused_train_step(rows).double()
            nb += 1
        return float(running.item()) / nb

    def validate_epoch(self):
        self.model.eval()
        total_loss, correct, nb = self._fused_validate()
        accuracy = 100 * correct / self._n_test
        return total_loss / nb, accuracy

    def train(self):
        print(f"Starting training for {self.epochs} epochs...")
        for epoch in range(self.epochs):
            train_loss = self.train_epoch()
            val_loss, accuracy = self.validate_epoch()
            print(f'Epoch {epoch + 1}/{self.epochs} | '
                  f'Train Loss: {train_loss:.4f} | '
                  f'Val Loss: {val_loss:.4f} | '
                  f'Val Acc: {accuracy:.2f}%')

    def predict(self, dataset=None):
        loader = self.test_loader
        if dataset is not None:
            loader = DataLoader(dataset, batch_size=32, shuffle=False)
        predictions = []
        self.model.eval()
        with torch.no_grad():
            for inputs, _ in loader:
                outputs = self.model(inputs.to(self.device))
                predictions.extend(outputs.argmax(1).cpu().tolist())
        return predictions
