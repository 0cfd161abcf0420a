"""-m gpu tests of the drop-in classes (the reference's own Python surface) end to end."""
import contextlib
import io
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _load_init(model, g):
    model.load_state_dict({k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("init::")})


def test_eegnet_tor_module_autograd_matches_reference(golden):
    """Unmodified user code: criterion(model(x), y).backward() runs on the kernels."""
    import gpu_util as U
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    g = golden("eegnet_tor_b8.npz")
    model = EEGNet_tor(5)
    _load_init(model, g)
    model = model.cuda()
    model.dropout_source = "torch_cpu"
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    crit = torch.nn.CrossEntropyLoss()
    for mode in ("train", "eval"):
        _load_init(model, g)
        model.train(mode == "train")
        model.zero_grad()
        torch.manual_seed(100)
        p = model(x)
        loss = crit(p, y)
        loss.backward()
        assert p.shape == (8, 5)
        assert U.rel_max(p.detach().cpu().numpy(), g[f"{mode}::probs"]) < TOL
        assert abs(loss.item() - float(g[f"{mode}::loss"])) < TOL * float(g[f"{mode}::loss"])
        for k, prm in model.named_parameters():
            assert U.rel_l2(prm.grad.cpu().numpy(), g[f"{mode}::grad::{k}"]) < TOL, (mode, k)
        sd = model.state_dict()
        for k in g.files:
            if k.startswith(f"{mode}::after::"):
                name = k.split("::")[-1]
                assert np.allclose(sd[name].cpu().numpy(), g[k], rtol=1e-5, atol=1e-6), name
    # 3-D input is accepted (superset); torch optimizers update the arena in place
    model.eval()
    with torch.no_grad():
        p3 = model(x[:, 0])
        p4 = model(x)
    assert torch.equal(p3, p4)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    before = model._arena.clone()
    model.zero_grad(); crit(model(x), y).backward(); opt.step()
    assert model._arena_ok() and not torch.equal(before, model._arena)


def test_trainer_uni_loop_matches_reference(golden):
    import golden_inputs as GI
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor, Trainer_uni
    g = golden("trainer_uni_3ep.npz")
    data = GI.trainer_inputs()
    model = EEGNet_tor(5)
    _load_init(model, g)
    model.dropout_source = "torch_cpu"
    trainer = Trainer_uni(model, list(data), lr=1e-3, batch_size=16, num_epochs=3)
    trainer.record_losses = True
    torch.manual_seed(77)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        trainer.train()
    got = torch.stack(trainer.loss_history).cpu().numpy()
    ref = g["train_step_loss"]
    assert got.shape == ref.shape == (9,)
    assert np.abs(got - ref).max() < TOL * np.abs(ref).max(), (got, ref)
    # same print cadence / format as the reference, values equal to print precision
    ref_lines = str(g["stdout"]).strip().splitlines()
    got_lines = buf.getvalue().strip().splitlines()
    assert len(ref_lines) == len(got_lines) == 6
    for a, b in zip(got_lines, ref_lines):
        assert a.split("Loss")[0] == b.split("Loss")[0]
        va = [float(t) for t in re.findall(r"\d+\.\d+", a)]
        vb = [float(t) for t in re.findall(r"\d+\.\d+", b)]
        assert len(va) == len(vb) and np.allclose(va, vb, atol=2e-4), (a, b)
    assert int(model.firstBN.num_batches_tracked) == 3                # F5: BN trains only in epoch 1
    assert not model.training
    final = model.state_dict()
    for k in ("firstBN.running_mean", "separableBN.running_var"):
        assert np.allclose(final[k].cpu().numpy(), g[f"final::{k}"], rtol=1e-4, atol=1e-6), k


def test_dataload_eeg_dropin_vs_oracle():
    import eeg_oracle as O
    from eav_b200.Dataload_eeg import DataLoadEEG
    from eav_b200.EAV_datasplit import EAVDataSplit
    raw, label = O.synth_subject(3)
    d = DataLoadEEG(subject=3, band=[5, 30])
    d.set_raw(np.transpose(raw, (2, 1, 0)), label)          # (Time, Channels, Trials) like the .mat
    x_dev, y = d.prepare_data_device()
    xo, yo = O.prepare_data(raw, label, [5, 30])
    assert np.array_equal(y, yo) and x_dev.shape == (400, 30, 500)
    x = x_dev.cpu().numpy()
    rms = np.sqrt((xo ** 2).mean(axis=(0, 2)))
    assert (np.abs(x - xo).max(axis=(0, 2)) / rms).max() < 1e-5
    # shipped labels {1,3,5,7,9} through the shipped split -> 112/48; remap -> 280/120
    trx, try_, tex, tey = EAVDataSplit(x, y).get_split(h_idx=56)
    assert trx.shape == (112, 30, 500) and tex.shape == (48, 30, 500)
    trx, try_, tex, tey = EAVDataSplit(x, (y - 1) // 2).get_split(h_idx=56)
    assert trx.shape == (280, 30, 500) and tex.shape == (120, 30, 500)
    # staged API leaves the same attributes and agrees with the fused path
    d2 = DataLoadEEG(subject=3, band=[5, 30])
    d2.set_raw(np.transpose(raw, (2, 1, 0)), label)
    d2.downsampling()
    assert d2.seg.shape == (30, 2000, 200)
    d2.bandpass_filter()
    assert d2.seg_f.shape == (30, 2000, 200)
    d2.segment_and_select_classes()
    assert d2.seg_f_div.shape == (400, 30, 500) and np.array_equal(d2.label_div, y)
    assert (np.abs(d2.seg_f_div - xo).max(axis=(0, 2)) / rms).max() < 1e-5


def test_cnn_eeg_trainer_runs_and_learns():
    from torch.utils.data import TensorDataset
    from eav_b200.CNN_torch.CNN_EEG import EEGNet, EEGNetTrainer
    torch.manual_seed(0)
    X = torch.randn(96, 64, 128)
    y = (X[:, :8, :32].mean(dim=(1, 2)) > 0).long() + 2 * (X[:, 8:16, :32].mean(dim=(1, 2)) > 0).long()
    model = EEGNet(nb_classes=4, Chans=64, Samples=128, dropoutRate=0.25)
    with contextlib.redirect_stdout(io.StringIO()):
        tr = EEGNetTrainer(model, TensorDataset(X[:80], y[:80]), TensorDataset(X[80:], y[80:]), batch_size=16, epochs=3, lr=1e-2)
        first = tr.train_epoch()
        for _ in range(12):
            last = tr.train_epoch()
        vloss, acc = tr.validate_epoch()
    assert last < first and np.isfinite(vloss) and 0 <= acc <= 100
    preds = tr.predict()
    assert len(preds) == 16 and all(isinstance(p, int) for p in preds)


def test_dataload_eeg_mat_file_path_end_to_end(tmp_path):
    """P1: the .mat layout the reference reads (Datasets/EAV/subjectNN/EEG/subjectNN_eeg.mat with 'seg',
    subjectNN_eeg_label.mat with 'label'; Dataload_eeg.py:54-83) through prepare_data() and its legacy alias."""
    import scipy.io
    import eeg_oracle as O
    from eav_b200.Dataload_eeg import DataLoadEEG
    raw, label = O.synth_subject(5)
    folder = tmp_path / "subject05" / "EEG"
    folder.mkdir(parents=True)
    scipy.io.savemat(str(folder / "subject05_eeg.mat"), {"seg": np.transpose(raw, (2, 1, 0))})      # (10000, 30, 200)
    scipy.io.savemat(str(folder / "subject05_eeg_label.mat"), {"label": label})
    with contextlib.redirect_stdout(io.StringIO()) as out:
        d = DataLoadEEG(subject=5, band=[0.5, 45], fs_orig=500, fs_target=100, parent_directory=str(tmp_path))
        x, y = d.data_prepare()
    assert "[Info] Loaded EEG data for subject05" in out.getvalue()
    xo, yo = O.prepare_data(raw, label, [0.5, 45])
    assert x.shape == (400, 30, 500) and x.dtype == np.float32 and np.array_equal(y, yo)
    rms = np.sqrt((xo ** 2).mean(axis=(0, 2)))
    assert (np.abs(x - xo).max(axis=(0, 2)) / rms).max() < 1e-5
    assert d.seg_f_div_device.is_cuda and d.seg.shape == (30, 10000, 200)
    # a missing subject prints the reference's error line and returns the empty placeholders
    with contextlib.redirect_stdout(io.StringIO()) as out:
        assert DataLoadEEG(subject=6, parent_directory=str(tmp_path)).prepare_data() == (None, None)
    assert "[Error] EEG data not found for subject06" in out.getvalue()


def test_trainer_uni_validate_matches_oracle(golden):
    """Evaluation path: validate() = eval-mode forward over the test loader, mean of the
    per-batch CE losses and the accuracy, as EEGNet_tor.py:118-135 computes them."""
    import eegnet_oracle as EO
    import golden_inputs as GI
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor, Trainer_uni
    g = golden("trainer_uni_3ep.npz")
    trx, try_, tex, tey = GI.trainer_inputs()
    model = EEGNet_tor(5)
    _load_init(model, g)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with contextlib.redirect_stdout(io.StringIO()) as out:
        trainer = Trainer_uni(model, [trx, try_, tex, tey], lr=1e-3, batch_size=6, num_epochs=1)
        loss, acc = trainer.validate()
    params, buffers = EO.split_state(sd, "tor")
    tot, correct, nb = 0.0, 0, 0
    with torch.no_grad():
        for b0 in range(0, 16, 6):                       # batches of 6, 6, 4 (ragged)
            xb, yb = torch.from_numpy(tex[b0:b0 + 6]), torch.from_numpy(tey[b0:b0 + 6])
            o = EO.tor_forward(params, buffers, xb, False)
            tot += float(EO.loss_fn(o, yb)); correct += int((o.argmax(1) == yb).sum()); nb += 1
    assert abs(loss - tot / nb) < 1e-4 * (tot / nb) and acc == correct / 16
    assert f"Validation - Loss: {tot / nb:.4f}, Accuracy: {correct / 16:.4f}" in out.getvalue()


def test_mat_ingest_pipeline_matches_direct_path(tmp_path):
    """prepare_subjects (zero-copy MAT reader + prefetch thread + copy stream) gives the epochs the
    per-subject DataLoadEEG.prepare_data() gives, for both preprocessing orders, with slots being reused."""
    import scipy.io
    import eeg_oracle as O
    from eav_b200 import mat_ingest as MI
    from eav_b200.Dataload_eeg import DataLoadEEG
    subs = [1, 2, 3, 4]
    for s in subs:
        folder = tmp_path / f"subject{s:02d}" / "EEG"
        folder.mkdir(parents=True)
        raw, label = O.synth_subject(s, n_trials=20, trial_len=10000)
        scipy.io.savemat(str(folder / f"subject{s:02d}_eeg.mat"), {"seg": np.transpose(raw, (2, 1, 0))},
                         do_compression=(s % 2 == 0))
        scipy.io.savemat(str(folder / f"subject{s:02d}_eeg_label.mat"), {"label": label})
    for legacy in (False, True):
        got = {s: (x.clone(), y) for s, x, y in MI.prepare_subjects(str(tmp_path), subs, band=[3, 45], legacy_order=legacy)}
        torch.cuda.synchronize()
        for s in subs:
            D = DataLoadEEG(subject=s, band=[3, 45], parent_directory=str(tmp_path))
            D.load_mat_data()
            x, y = D.prepare_data_legacy_device(band=(3, 45)) if legacy else D.prepare_data_device()
            assert np.array_equal(got[s][1], y)
            assert torch.equal(got[s][0], x), (legacy, s)
