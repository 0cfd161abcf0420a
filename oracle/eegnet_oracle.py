"""CPU oracle (torch-CPU fp32 restatement) of the EEGNet training path of the
reference -- TEST INFRASTRUCTURE, never shipped (see eeg_oracle.py header).

Restates, as explicit functional arithmetic over named parameter dictionaries:
  * CNN_torch/EEGNet_tor.py:15-67   EEGNet_tor.__init__/forward      (variant "tor")
  * CNN_torch/CNN_EEG.py:7-67       EEGNet.__init__/forward          (variant "cnn")
  * CNN_torch/EEGNet_tor.py:33-34,47-48  max-norm forward hooks with their intended
    semantics: layer forward uses W_old, then W <- renorm(W) in
    place; backward's grad_input sees W_new.
  * nn.CrossEntropyLoss on the model output (probabilities for "tor": SURVEY F6).
  * torch.optim.Adam single-tensor update order (EEGNet_tor.py:82,108-110).
  * Trainer_uni.train()/validate() control flow incl. the train-mode-only-in-epoch-1
    quirk (EEGNet_tor.py:96-135) and EEGNetTrainer (CNN_EEG.py:70-162).

Parity status: PINNED against the shimmed reference executed in the build
container (tests/golden/eegnet_*.npz written by oracle/gen_golden.py; checked by
tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# parameter order == reference module construction order (EEGNet_tor.py:24-43) ==
# the order torch.optim.Adam sees them == the flat arena order of the CUDA path.
TOR_PARAMS = ("firstConv.weight", "firstBN.weight", "firstBN.bias",
              "depthwiseConv.weight", "depthwiseBN.weight", "depthwiseBN.bias",
              "separableConv.weight", "separableBN.weight", "separableBN.bias",
              "dense.weight", "dense.bias")
TOR_BN = ("firstBN", "depthwiseBN", "separableBN")
CNN_PARAMS = ("block1.0.weight", "block1.1.weight", "block1.1.bias",
              "block1.2.weight", "block1.3.weight", "block1.3.bias",
              "block2.0.weight", "block2.1.weight", "block2.2.weight", "block2.2.bias",
              "classifier.weight", "classifier.bias")
CNN_BN = ("block1.1", "block1.3", "block2.2")


def split_state(state_dict, variant="tor"):
    """(params requiring grad, buffers) as cloned fp32 leaf tensors."""
    names = TOR_PARAMS if variant == "tor" else CNN_PARAMS
    params = OrderedDict((k, state_dict[k].detach().clone().float().requires_grad_(True)) for k in names)
    buffers = OrderedDict((k, v.detach().clone()) for k, v in state_dict.items() if k not in names)
    return params, buffers


def _same_pad(k):
    total = k - 1  # torch padding='same', stride 1, dilation 1
    return total // 2, total - total // 2


def _bn(x, prefix, params, buffers, train, eps=1e-5, momentum=0.1):
    """nn.BatchNorm2d: batch statistics + running-stat update when train, else
    running statistics (frozen)."""
    rm, rv = buffers[prefix + ".running_mean"], buffers[prefix + ".running_var"]
    if train:
        buffers[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, params[prefix + ".weight"], params[prefix + ".bias"],
                        training=train, momentum=momentum, eps=eps)


def _dropout(x, p, train, masks, dropout2d=False):
    """nn.Dropout / nn.Dropout2d.  masks: None -> draw from the global CPU RNG in the
    reference's order; list -> pop(0) an explicit keep-mask (1 keep / 0 drop) and
    record nothing; a dict {'record': []} records the drawn masks."""
    if not train or p == 0.0:
        return x
    if isinstance(masks, list):
        m = masks.pop(0).to(x.dtype)
    else:
        shape = x.shape[:2] + (1, 1) if dropout2d else x.shape
        m = torch.empty(shape, dtype=x.dtype).bernoulli_(1 - p)
        if isinstance(masks, dict):
            masks["record"].append(m.clone())
    return x * m / (1 - p) if p < 1 else x * 0


def renorm_rows_(w, maxnorm):
    """torch.renorm(p=2, dim=0) in place on w.data: rows with L2 norm > maxnorm are
    scaled by maxnorm / (norm + 1e-7)."""
    with torch.no_grad():
        w.data.renorm_(p=2, dim=0, maxnorm=maxnorm)


def tor_forward(params, buffers, x, train, cfg=None, masks=None):
    """EEGNet_tor.forward (EEGNet_tor.py:50-67) -> probabilities [B, nb_classes].
    cfg: dict(dropoutRate=0.5, norm_rate=1.0, dropoutType='Dropout', F1, D ...)."""
    cfg = cfg or {}
    p = cfg.get("dropoutRate", 0.5)
    nr = cfg.get("norm_rate", 1.0)
    d2d = cfg.get("dropoutType", "Dropout") != "Dropout"
    W1, W2, W3 = params["firstConv.weight"], params["depthwiseConv.weight"], params["separableConv.weight"]
    F1 = W1.shape[0]
    l, r = _same_pad(W1.shape[-1])
    h = F.conv2d(F.pad(x, (l, r)), W1)
    h = F.elu(_bn(h, "firstBN", params, buffers, train))
    h = F.conv2d(h, W2, groups=F1)
    renorm_rows_(W2, nr)                      # hook, EEGNet_tor.py:33-34 (intended semantics)
    h = F.elu(_bn(h, "depthwiseBN", params, buffers, train))
    h = F.avg_pool2d(h, (1, 4))
    h = _dropout(h, p, train, masks, d2d)
    l, r = _same_pad(W3.shape[-1])
    h = F.conv2d(F.pad(h, (l, r)), W3)
    h = F.elu(_bn(h, "separableBN", params, buffers, train))
    h = F.avg_pool2d(h, (1, 8))
    h = _dropout(h, p, train, masks, d2d)
    h = h.flatten(1)
    z = F.linear(h, params["dense.weight"], params["dense.bias"])
    renorm_rows_(params["dense.weight"], nr)  # hook, EEGNet_tor.py:47-48
    return F.softmax(z, dim=1)


def cnn_forward(params, buffers, x, train, cfg=None, masks=None):
    """CNN_EEG.EEGNet.forward (CNN_EEG.py:57-67) -> logits."""
    cfg = cfg or {}
    p = cfg.get("dropoutRate", 0.5)
    if x.dim() == 3:
        x = x.unsqueeze(1)
    W1, W2 = params["block1.0.weight"], params["block1.2.weight"]
    W3d, W3p = params["block2.0.weight"], params["block2.1.weight"]
    F1 = W1.shape[0]
    l, r = _same_pad(W1.shape[-1])
    h = F.conv2d(F.pad(x, (l, r)), W1)
    h = _bn(h, "block1.1", params, buffers, train)          # no ELU here (CNN_EEG.py:22-26)
    h = F.conv2d(h, W2, groups=F1)
    h = F.elu(_bn(h, "block1.3", params, buffers, train))
    h = F.avg_pool2d(h, (1, 4))
    h = _dropout(h, p, train, masks)
    l, r = _same_pad(W3d.shape[-1])
    h = F.conv2d(F.pad(h, (l, r)), W3d, groups=W3d.shape[0])
    h = F.conv2d(h, W3p)
    h = F.elu(_bn(h, "block2.2", params, buffers, train))
    h = F.avg_pool2d(h, (1, 8))
    h = _dropout(h, p, train, masks)
    h = h.flatten(1)
    return F.linear(h, params["classifier.weight"], params["classifier.bias"])


def forward(variant, params, buffers, x, train, cfg=None, masks=None):
    return (tor_forward if variant == "tor" else cnn_forward)(params, buffers, x, train, cfg, masks)


def loss_fn(out, y):
    """nn.CrossEntropyLoss(): mean over the batch of -log_softmax(out)[y]."""
    return F.cross_entropy(out, y)


class Adam:
    """torch.optim.Adam(lr, betas=(0.9,0.999), eps=1e-8, weight_decay=0, amsgrad=False),
    single-tensor CPU update order: m.lerp_(g, 1-b1); v = b2*v + (1-b2)*g*g;
    denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) * m/denom."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        self.params = list(params.values()) if isinstance(params, dict) else list(params)
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.t = 0
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = self.lr / bc1
        bc2_sqrt = math.sqrt(bc2)
        for p, m, v in zip(self.params, self.m, self.v):
            if p.grad is None:
                continue
            g = p.grad
            m.lerp_(g, 1 - self.b1)
            v.mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (v.sqrt() / bc2_sqrt).add_(self.eps)
            p.addcdiv_(m, denom, value=-step_size)


def train_step(variant, params, buffers, opt, x, y, train, cfg=None, masks=None):
    """One Trainer step (EEGNet_tor.py:104-110 / CNN_EEG.py:95-104). Returns (loss, out)."""
    out = forward(variant, params, buffers, x, train, cfg, masks)
    loss = loss_fn(out, y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return loss.detach(), out.detach()


def trainer_uni_train(params, buffers, data, lr=1e-4, batch_size=32, num_epochs=10, cfg=None, log=None):
    """Trainer_uni.__init__ + train() + validate() (EEGNet_tor.py:70-135) restated.
    The torch DataLoader stays the index/permutation source so the global-RNG
    consumption order matches the reference.  Only epoch 1 runs in train mode
    (SURVEY F5).  Returns a log dict: per-step train losses, per-epoch validation
    (loss, accuracy)."""
    from torch.utils.data import DataLoader, TensorDataset
    tr_x, tr_y, te_x, te_y = data
    mk = lambda x, y, sh: DataLoader(TensorDataset(torch.tensor(x, dtype=torch.float32),
                                                   torch.tensor(y, dtype=torch.long)),
                                     batch_size=batch_size, shuffle=sh)
    train_dl, test_dl = mk(tr_x, tr_y, True), mk(te_x, te_y, False)
    opt = Adam(params, lr)
    log = log if log is not None else {}
    log.setdefault("step_loss", [])
    log.setdefault("val", [])
    training = True                      # self.model.train() once (EEGNet_tor.py:97)
    for _ in range(num_epochs):
        for xb, yb in train_dl:
            loss, _ = train_step("tor", params, buffers, opt, xb, yb, training, cfg)
            log["step_loss"].append(float(loss))
        training = False                 # validate() -> model.eval(), never undone (F5)
        tot, correct = 0.0, 0
        with torch.no_grad():
            for xb, yb in test_dl:
                out = tor_forward(params, buffers, xb, False, cfg)
                tot += float(loss_fn(out, yb))
                correct += int((out.argmax(1) == yb).sum())
        log["val"].append((tot / len(test_dl), correct / len(test_dl.dataset)))
    return log
