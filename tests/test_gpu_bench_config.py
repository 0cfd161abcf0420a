"""-m gpu parity on the BENCHMARK's own configuration: goldens written by the unmodified reference
for one full subject -- 280 / 120 epochs, batch 32 with the ragged last batch of 24, lr 1e-5, epoch 1 in train mode and
epoch 2 in eval mode -- checked through the drop-in Trainer_uni AND through the lock-step
SubjectBatchTrainer at M = 42 (what bench.py times), plus the EEGNetTrainer loop golden (CNN_EEG.py:88-144)."""
import contextlib
import io
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _init(g):
    return {k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("init::")}


def test_trainer_uni_on_the_bench_configuration_matches_reference(golden):
    import golden_inputs as GI
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor, Trainer_uni
    g = golden("trainer_uni_bench_2ep.npz")
    data = GI.bench_subject_inputs()
    assert np.allclose(GI.checksum(*data), g["input_checksum"], rtol=1e-12)
    model = EEGNet_tor(5)
    model.load_state_dict(_init(g))
    model.dropout_source = "torch_cpu"                       # masks from torch's CPU RNG in the reference's order
    trainer = Trainer_uni(model, list(data), lr=1e-5, batch_size=32, num_epochs=2)
    trainer.record_losses = True
    torch.manual_seed(78)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        trainer.train()
    got = torch.stack(trainer.loss_history).cpu().numpy()
    ref = g["train_step_loss"]
    assert got.shape == ref.shape == (18,)                   # 2 epochs x (8 x 32 + 24)
    assert np.abs(got - ref).max() < TOL * np.abs(ref).max(), (got, ref)
    # per-epoch mean loss (what a loss curve shows)
    assert np.allclose(got.reshape(2, 9).mean(1), ref.reshape(2, 9).mean(1), rtol=TOL)
    ref_lines = str(g["stdout"]).strip().splitlines()
    got_lines = buf.getvalue().strip().splitlines()
    assert len(ref_lines) == len(got_lines) == 4             # one step print + one validation line per epoch
    for a, b in zip(got_lines, ref_lines):
        assert a.split("Loss")[0] == b.split("Loss")[0]
        va = [float(t) for t in re.findall(r"\d+\.\d+", a)]
        vb = [float(t) for t in re.findall(r"\d+\.\d+", b)]
        assert len(va) == len(vb) and np.allclose(va, vb, atol=2e-4), (a, b)
    assert int(model.firstBN.num_batches_tracked) == 9 and not model.training
    final = model.state_dict()
    for k in g.files:
        if k.startswith("final::") and "num_batches" not in k:
            assert np.allclose(final[k[7:]].cpu().numpy(), g[k], rtol=1e-4, atol=1e-6), k


def test_lockstep_trainer_at_42_models_matches_reference_step_by_step(golden):
    """bench.py's engine: 42 models advanced by the same launches.  Every model replays the reference's batches and
    dropout masks; each of the 42 per-model losses must equal the reference's loss at every one of the 18 steps."""
    import golden_inputs as GI
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    from eav_b200.ops import EegnetDims
    from eav_b200.trainer_core import SubjectBatchTrainer
    g = golden("trainer_uni_bench_2ep.npz")
    trx, try_, tex, tey = GI.bench_subject_inputs()
    M = 42
    x = torch.from_numpy(np.concatenate([trx, tex])[:, 0]).cuda().contiguous()       # one resident copy, shared rows
    y = torch.from_numpy(np.concatenate([try_, tey])).long().cuda()
    core = SubjectBatchTrainer(EegnetDims(5), M, x, y, lr=1e-5, max_batch=32, use_graph=False)
    core.load_state_dicts([_init(g)] * M, EEGNet_tor._BN_NAMES)
    rows = g["batch_rows"]
    sizes = ([32] * 8 + [24]) * 2
    m1 = np.unpackbits(g["mask1_bits"])[:sum(sizes[:9]) * 64 * 125]
    m2 = np.unpackbits(g["mask2_bits"])[:sum(sizes[:9]) * 64 * 15]
    ref = g["train_step_loss"]
    r0 = o1 = o2 = 0
    for s, B in enumerate(sizes):
        idx = torch.from_numpy(np.tile(rows[r0:r0 + B], M)).int().cuda()
        r0 += B
        masks = None
        if s < 9:                                            # epoch 1: train-mode BN + the reference's own masks
            k1 = torch.from_numpy(m1[o1:o1 + B * 64 * 125].reshape(B, 64, 125)); o1 += B * 64 * 125
            k2 = torch.from_numpy(m2[o2:o2 + B * 64 * 15].reshape(B, 64, 15)); o2 += B * 64 * 15
            masks = (k1.repeat(M, 1, 1).cuda().contiguous(), k2.repeat(M, 1, 1).cuda().contiguous())
        loss = core.train_step(idx, bn_train=s < 9, masks=masks).cpu().numpy()
        assert loss.shape == (M,)
        assert np.abs(loss - ref[s]).max() < TOL * abs(ref[s]), (s, loss[:3], ref[s])
    # validation pass of epoch 2 through the eval programs: batch losses as the reference printed them
    vref = g["val_batch_loss"][4:]
    for v, b0 in enumerate(range(0, 120, 32)):
        cols = np.arange(b0, min(120, b0 + 32)) + 280
        loss, _, _ = core.eval_batch(torch.from_numpy(np.tile(cols, M)).int().cuda())
        assert np.abs(loss.cpu().numpy() - vref[v]).max() < TOL * abs(vref[v]), v


def test_eegnet_trainer_loop_matches_reference(golden):
    """CNN_EEG.EEGNetTrainer (CNN_EEG.py:70-162): train_epoch / validate_epoch / predict over two epochs."""
    import golden_inputs as GI
    from torch.utils.data import TensorDataset
    from eav_b200.CNN_torch.CNN_EEG import EEGNet, EEGNetTrainer
    g = golden("cnn_eeg_trainer_2ep.npz")
    trx, try_, tex, tey = GI.cnn_trainer_inputs()
    assert np.allclose(GI.checksum(trx.numpy(), try_.numpy(), tex.numpy(), tey.numpy()), g["input_checksum"], rtol=1e-12)
    model = EEGNet(nb_classes=4, Chans=64, Samples=128, dropoutRate=0.25)
    model.load_state_dict(_init(g))
    model.dropout_source = "torch_cpu"
    with contextlib.redirect_stdout(io.StringIO()):
        tr = EEGNetTrainer(model, TensorDataset(trx, try_), TensorDataset(tex, tey), batch_size=32, epochs=2, lr=1e-3)
        torch.manual_seed(79)
        tl1 = tr.train_epoch(); v1 = tr.validate_epoch()
        tl2 = tr.train_epoch(); v2 = tr.validate_epoch()
        pred = tr.predict()
    assert np.allclose([tl1, tl2], g["train_loss"], rtol=TOL), ([tl1, tl2], g["train_loss"])
    assert np.allclose([v1[0], v2[0]], g["val_loss"], rtol=TOL)
    assert np.allclose([v1[1], v2[1]], g["val_acc"], atol=100.0 / 40 + 1e-9)      # at most one borderline sample
    assert np.mean(np.array(pred) == g["predict"]) >= 0.95
    final = model.state_dict()
    for k in g.files:
        if k.startswith("final::") and "num_batches" not in k:
            if k.endswith("block1.3.running_mean"):
                # In this variant nothing non-linear sits between BatchNorm-1 and the depthwise conv, and train-mode
                # BatchNorm-2 removes any per-channel shift: d(loss)/d(block1.1.bias) is EXACTLY zero in exact arithmetic,
                # so what Adam normalises to +-lr steps is rounding noise ("Adam's first steps are
                # sign-like").  The bias, and with it this running mean (= W2 . bias), random-walks differently on every
                # implementation; no loss or prediction depends on it.
                continue
            assert np.allclose(final[k[7:]].cpu().numpy(), g[k], rtol=1e-4, atol=2e-5), k
