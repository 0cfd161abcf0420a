"""-m gpu tests of the whole-epoch CUDA graph (trainer_core.EpochRunner): the on-device batch schedule
is a per-model permutation keyed by the SUBJECT, and replaying one graph per epoch gives exactly what the
step-by-step path gives on the same schedule (Trainer_uni.train(), EEGNet_tor.py:96-135: 9 steps incl. the
ragged last batch, then the validation pass; epoch 1 in train mode, later epochs in eval mode)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _data(M, n_tr, n_te, seed=0):
    g = torch.Generator().manual_seed(seed)
    rows = n_tr + n_te
    w = torch.randn(5, 30, generator=g)
    y = torch.randint(0, 5, (M * rows,), generator=g)
    x = torch.randn(M * rows, 30, 500, generator=g) + 0.8 * w[y].unsqueeze(-1) * torch.sin(torch.arange(500) * 0.2)
    return x.cuda(), y.cuda()


def _core(M, x, y, batch, seed=5, lr=1e-3, dropout=0.5, weight_scale=1.0):
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    from eav_b200.trainer_core import SubjectBatchTrainer
    sds, dims = [], None
    for m in range(M):
        torch.manual_seed(10 + m)
        mdl = EEGNet_tor(5, dropoutRate=dropout)
        if weight_scale != 1.0:                 # rows above the max-norm: the forward hooks really clip
            with torch.no_grad():
                mdl.depthwiseConv.weight.mul_(weight_scale)
                mdl.dense.weight.mul_(weight_scale)
        dims = mdl._dims
        sds.append(mdl.state_dict())
    core = SubjectBatchTrainer(dims, M, x, y, lr=lr, max_batch=batch, seed=seed)
    core.load_state_dicts(sds, EEGNet_tor._BN_NAMES)
    return core


def test_schedule_is_a_permutation_keyed_by_subject():
    M, n_tr, n_te, B = 3, 280, 120, 32
    x, y = _data(1, 4, 4)                       # the schedule kernel never touches the data
    core = _core(M, x, y, B)
    r = core.epoch_runner(n_tr, n_te, B, subject_ids=[4, 9, 17], seed=77)
    e0, e1 = r.peek_schedule(0), r.peek_schedule(1)
    assert [t.numel() for t in e0] == [M * 32] * 8 + [M * 24]       # 9 steps, the last one ragged (280 = 8*32 + 24)
    for m in range(M):
        rows0 = torch.cat([t.reshape(M, -1)[m] for t in e0])
        rows1 = torch.cat([t.reshape(M, -1)[m] for t in e1])
        lo = m * (n_tr + n_te)
        assert sorted(rows0.tolist()) == list(range(lo, lo + n_tr))   # a permutation of the model's own train rows
        assert sorted(rows1.tolist()) == list(range(lo, lo + n_tr))
        assert rows0.tolist() != rows1.tolist()                       # a fresh one every epoch
    # the stream follows the subject id, not the slot or the number of co-resident models
    solo = _core(1, x, y, B).epoch_runner(n_tr, n_te, B, subject_ids=[9], seed=77).peek_schedule(0)
    mine = torch.cat([t.reshape(M, -1)[1] for t in e0]) - (n_tr + n_te)
    assert torch.equal(torch.cat([t.reshape(-1) for t in solo]), mine)


@pytest.mark.parametrize("weight_scale", [1.0, 4.0])
@pytest.mark.parametrize("pipeline", [True, False])
@pytest.mark.parametrize("M", [1, 3])
def test_epoch_graph_equals_step_by_step(M, pipeline, weight_scale):
    """weight_scale 4: the max-norm hooks clip in every forward, including validate()'s -- the pipelined validation (on a
    snapshot) must leave the live weights exactly as the in-between validation does."""
    n_tr, n_te, B = 88, 40, 32                 # 3 train steps (32, 32, 24) + 2 validation batches (32, 8)
    x, y = _data(M, n_tr, n_te, seed=3)
    lr = 1e-3 if weight_scale == 1.0 else 2e-2  # big steps push rows back over the norm between the forwards
    a, b = _core(M, x, y, B, lr=lr, weight_scale=weight_scale), _core(M, x, y, B, lr=lr, weight_scale=weight_scale)
    ra = a.epoch_runner(n_tr, n_te, B, seed=123, max_epochs=8, pipeline_validation=pipeline)   # validation of epoch e inside graph e+1, or in between
    rb = b.epoch_runner(n_tr, n_te, B, seed=123, max_epochs=8)
    n_epochs = 3
    for e in range(n_epochs):                  # A: one graph replay per epoch
        ra.run_epoch(bn_train=(e == 0))
    hist = ra.results().numpy()
    assert hist.shape == (n_epochs, M, 3)
    base = (torch.arange(M) * (n_tr + n_te) + n_tr).unsqueeze(1)
    for e in range(n_epochs):                  # B: the same schedule, one launch sequence per step
        sched = rb.peek_schedule(e)
        run = np.zeros(M)
        for idx in sched:
            run += b.train_step(idx.cuda(), bn_train=(e == 0)).double().cpu().numpy()
        vloss, corr, nb = np.zeros(M), np.zeros(M), 0
        for b0 in range(0, n_te, B):
            cols = torch.arange(b0, min(n_te, b0 + B)).unsqueeze(0)
            loss, nc, _ = b.eval_batch((base + cols).reshape(-1).int().cuda())
            vloss += loss.double().cpu().numpy()
            corr += nc.cpu().numpy()
            nb += 1
        assert np.allclose(hist[e, :, 0], run / len(sched), rtol=1e-6, atol=0), (e, hist[e, :, 0], run / len(sched))
        assert np.allclose(hist[e, :, 1], vloss / nb, rtol=1e-6, atol=0)
        assert np.allclose(hist[e, :, 2], corr / n_te, rtol=1e-6, atol=0)
    # same kernels on the same inputs in the same order: the trained state is bit-identical
    assert torch.equal(a.params, b.params)
    assert torch.equal(a.bn_state, b.bn_state)
    assert torch.equal(a.exp_avg_sq, b.exp_avg_sq)
    assert int(a.step_dev.item()) == n_epochs * 3 and ra.epochs_done() == n_epochs
    if weight_scale == 1.0:
        assert hist[-1, :, 0].mean() < hist[0, :, 0].mean()         # and it learns


def test_train_subjects_does_not_depend_on_the_sharding():
    """A subject trained alone (as on an 8-GPU shard) and among others sees the same batches."""
    from eav_b200.sharding import train_subjects

    def subject(s, n_tr=40, n_te=16):
        g = torch.Generator().manual_seed(100 + s)
        w = torch.randn(5, 30, generator=g)
        def make(n):
            yy = torch.randint(0, 5, (n,), generator=g)
            xx = torch.randn(n, 30, 500, generator=g) + 0.8 * w[yy].unsqueeze(-1) * torch.sin(torch.arange(500) * 0.2)
            return xx, yy
        return make(n_tr) + make(n_te)

    kw = dict(nb_classes=5, lr=1e-3, batch_size=16, num_epochs=2, model_kwargs=dict(dropoutRate=0.0))
    _, loss_all, run_all = train_subjects([2, 5, 11], subject, return_runner=True, **kw)
    _, loss_one, run_one = train_subjects([5], subject, return_runner=True, **kw)
    s_all = torch.cat([t.reshape(3, -1)[1] for t in run_all.peek_schedule(1)]) - 56
    s_one = torch.cat([t.reshape(-1) for t in run_one.peek_schedule(1)])
    assert torch.equal(s_all, s_one)
    assert np.allclose(loss_all[5], loss_one[5], rtol=1e-3)
