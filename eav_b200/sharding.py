"""Multi-GPU by SUBJECT (BASELINE.json configs[2]): the 42 per-subject models
share nothing, so subject s (1-based) lives on rank (s-1) % world and the data path needs NO
collective.  One process per GPU; torch.distributed is used only for the control plane
(gathering the 42 accuracies on rank 0).  The reference's own multi-GPU mechanism,
single-process nn.DataParallel over one tiny batch (EEGNet_tor.py:86-88), is not reproduced.
"""
from __future__ import annotations

import io
import contextlib
from typing import Callable, Dict, List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def subjects_for_rank(subjects: Sequence[int], rank: int, world: int) -> List[int]:
    """Round-robin assignment: subject s -> rank (s-1) % world (ids are 1-based like the dataset)."""
    return [s for s in subjects if (s - 1) % world == rank]


def shard_sizes(n_subjects: int, world: int) -> List[int]:
    return [len(subjects_for_rank(range(1, n_subjects + 1), r, world)) for r in range(world)]


def gather_results(local: Dict[int, float], group=None) -> Dict[int, float] | None:
    """Control-plane gather of {subject: metric} onto rank 0 (None elsewhere)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(dict(local), bucket, dst=0, group=group)
    if rank != 0:
        return None
    out: Dict[int, float] = {}
    for part in bucket:
        out.update(part)
    return dict(sorted(out.items()))


def train_subjects(subjects: Sequence[int], load_subject: Callable[[int], tuple], nb_classes=5, lr=1e-5,
                   batch_size=32, num_epochs=10, device=None, reference_eval_quirk=True, seed_base=0,
                   model_kwargs=None, verbose=False, use_graph=True, return_runner=False):
    """Train one EEGNet_tor per subject, all subjects of THIS rank in lock-step on one GPU.

    load_subject(s) -> (tr_x [N,Chans,Samples], tr_y [N], te_x, te_y) as numpy / tensors (all
    subjects must share shapes, as in the dataset: 280 / 120 epochs of 30 x 500).
    Follows Trainer_uni.train() (EEGNet_tor.py:96-135): per-epoch shuffled batches of
    `batch_size`, ragged last batch kept, then a validation pass; with reference_eval_quirk only
    epoch 1 runs BatchNorm/dropout in train mode.  Model s is initialised under
    torch.manual_seed(seed_base + s) exactly like a stand-alone EEGNet_tor(nb_classes).
    Returns ({subject: test accuracy}, {subject: [per-epoch mean train loss]}) (+ the EpochRunner
    when return_runner: its history holds per-epoch validation loss / accuracy as well).
    """
    from .CNN_torch.EEGNet_tor import EEGNet_tor
    from .ops import EegnetDims
    from .trainer_core import SubjectBatchTrainer
    subjects = list(subjects)
    M = len(subjects)
    if M == 0:
        return {}, {}
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    data = [load_subject(s) for s in subjects]
    as_t = lambda a, dt: (a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))).to(dt)
    n_tr, n_te = int(data[0][0].shape[0]), int(data[0][2].shape[0])
    rows = n_tr + n_te
    xs, ys = [], []
    for trx, try_, tex, tey in data:
        trx, tex = as_t(trx, torch.float32), as_t(tex, torch.float32)
        if trx.dim() == 4:
            trx, tex = trx[:, 0], tex[:, 0]
        if trx.shape[0] != n_tr or tex.shape[0] != n_te:
            raise ValueError("all subjects must have the same number of train/test epochs")
        xs += [trx, tex]
        ys += [as_t(try_, torch.int64), as_t(tey, torch.int64)]
    x = torch.cat(xs).to(dev).contiguous()
    y = torch.cat(ys).to(dev).contiguous()
    kw = dict(model_kwargs or {})
    sds, dims = [], None
    for s in subjects:
        torch.manual_seed(seed_base + s)
        mdl = EEGNet_tor(nb_classes, **kw)
        dims = mdl._dims
        sds.append(mdl.state_dict())
    core = SubjectBatchTrainer(dims, M, x, y, lr=lr, max_batch=batch_size, seed=seed_base)
    core.load_state_dicts(sds, EEGNet_tor._BN_NAMES)
    # One CUDA graph per epoch (trainer_core.EpochRunner): the per-epoch permutation is drawn ON THE DEVICE from a
    # stream keyed by the SUBJECT id, so a subject's result does not depend on which other subjects share its GPU
    # (i.e. on the number of ranks), and nothing touches the host between steps -- at 5-6 subjects per GPU
    # (42 subjects over 8 GPUs) a step is ~0.4 ms of kernels and any per-step host tensor op would dominate it.
    runner = core.epoch_runner(n_tr, n_te, batch_size, rows_per_model=rows, subject_ids=subjects,
                               max_epochs=max(1, num_epochs), seed=1000003 * (seed_base + 1))
    training = True
    for epoch in range(num_epochs):
        runner.run_epoch(bn_train=training, use_graph=use_graph)
        if reference_eval_quirk:
            training = False                                            # validate() -> model.eval(), never undone
    hist = runner.results()                                             # the only synchronisation
    losses = {s: [float(hist[e, i, 0]) for e in range(hist.shape[0])] for i, s in enumerate(subjects)}
    acc = {s: (float(hist[-1, i, 2]) if hist.shape[0] else float("nan")) for i, s in enumerate(subjects)}
    if verbose:
        for e in range(hist.shape[0]):
            print(f"epoch {e + 1}/{num_epochs}: mean train loss {float(hist[e, :, 0].mean()):.4f}, "
                  f"mean test acc {float(hist[e, :, 2].mean()):.4f}")
    if return_runner:
        return acc, losses, runner
    return acc, losses
