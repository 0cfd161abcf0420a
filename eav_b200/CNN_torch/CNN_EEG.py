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
from .._module_base import ArenaAdam, ArenaModule, FusedTrainerMixin


class EEGNet(ArenaModule):
    _VARIANT = EAV_VARIANT_CNN
    _BN_NAMES = ("block1.1", "block1.3", "block2.2")

    def __init__(self, nb_classes, Chans=64, Samples=128, dropoutRate=0.5, kernLength=64, F1=8, D=2, F2=16,
                 norm_rate=0.25):
        super().__init__()
        self.Chans, self.Samples = Chans, Samples
        G = D * F1
        # Layer table of the two containers.  Position in the container == state_dict index of the reference
        # (CNN_EEG.py:20-47) and construction order == its order of draws from torch's global RNG, so a seeded
        # reference checkpoint loads and a seeded construction initialises identically.
        spec1 = (("conv", 1, F1, (1, kernLength), 1, "same"), ("bn", F1), ("conv", F1, G, (Chans, 1), F1, 0),
                 ("bn", G), ("elu",), ("pool", 4), ("drop",))
        spec2 = (("conv", G, G, (1, 16), G, "same"), ("conv", G, F2, (1, 1), 1, 0), ("bn", F2), ("elu",),
                 ("pool", 8), ("drop",))

        def make(item):
            kind = item[0]
            if kind == "conv":
                _, cin, cout, ksize, groups, pad = item
                return nn.Conv2d(cin, cout, ksize, padding=pad, groups=groups, bias=False)
            if kind == "bn":
                return nn.BatchNorm2d(item[1])
            if kind == "elu":
                return nn.ELU()
            if kind == "pool":
                return nn.AvgPool2d((1, item[1]))
            return nn.Dropout(dropoutRate)

        self.block1 = nn.Sequential(*[make(i) for i in spec1])
        self.block2 = nn.Sequential(*[make(i) for i in spec2])
        self.flatten = nn.Flatten()
        # The reference sizes the classifier with a dry run through block1/block2 in TRAIN mode
        # (CNN_EEG.py:48-53).  Its side effects are part of the initial state and are reproduced
        # by doing the same at construction time (host, once): BatchNorm running_var -> 0.9 and
        # num_batches_tracked -> 1, and two dropout masks drawn from the global RNG BEFORE the
        # classifier is initialised.
        probe = torch.zeros(1, 1, Chans, Samples)
        with torch.no_grad():
            n_flatten = int(self.flatten(self.block2(self.block1(probe))).shape[1])
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
        have_cuda = torch.cuda.is_available()
        self.device = torch.device("cuda" if have_cuda else "cpu")
        print(f"Using device: {self.device}")
        if not have_cuda:
            raise RuntimeError("eav_b200.EEGNetTrainer needs a CUDA (B200) device; there is no CPU fallback")
        self.model = model.to(self.device)
        self.epochs, self.batch_size, self.lr = epochs, batch_size, lr
        # the reference's public attributes (CNN_EEG.py:80-86); the fused step below does not iterate them
        self.train_loader, self.test_loader = (DataLoader(ds, batch_size=batch_size, shuffle=sh)
                                               for ds, sh in ((train_dataset, True), (val_dataset, False)))
        self.criterion = nn.CrossEntropyLoss()
        self.optimizer = ArenaAdam(model.parameters(), lr=lr)       # optim.Adam whose state is the fused trainer's
        trx, try_ = _dataset_tensors(train_dataset)
        tex, tey = _dataset_tensors(val_dataset)
        self._setup_fused(self.model, trx, try_, tex, tey, lr=lr, batch_size=batch_size)

    def train_epoch(self):
        """mean training loss of one epoch (CNN_EEG.py:88-108)"""
        self.model.train()
        running = torch.zeros((), dtype=torch.float64, device=self.device)
        nb = 0
        for rows in self._index_batches(train=True):
            running += self._fused_train_step(rows).double()
            nb += 1
        return float(running.item()) / nb

    def validate_epoch(self):
        """(mean validation loss, accuracy in percent) (CNN_EEG.py:110-133)"""
        self.model.eval()
        total_loss, correct, nb = self._fused_validate()
        return total_loss / nb, 100 * correct / self._n_test

    def train(self):
        """Same console output as CNN_EEG.py:135-146."""
        print(f"Starting training for {self.epochs} epochs...")
        for ep in range(1, self.epochs + 1):
            tl = self.train_epoch()
            vl, acc = self.validate_epoch()
            print(f"Epoch {ep}/{self.epochs} | Train Loss: {tl:.4f} | Val Loss: {vl:.4f} | Val Acc: {acc:.2f}%")

    def predict(self, dataset=None):
        """argmax class of every sample of `dataset` (default: the validation set), CNN_EEG.py:148-162."""
        loader = self.test_loader if dataset is None else DataLoader(dataset, batch_size=32, shuffle=False)
        self.model.eval()
        out = []
        with torch.no_grad():
            for xb, _ in loader:
                out += self.model(xb.to(self.device)).argmax(1).cpu().tolist()
        return out
