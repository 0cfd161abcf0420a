"""-m gpu tests of the round-2 robustness items: on-device Dropout2d, copy /
pickle of a drop-in module after a forward, the Adam state behind trainer.optimizer, learning-rate edits."""
import copy
import io

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _data(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 1, 30, 500, generator=g), torch.randint(0, 5, (n,), generator=g)


def test_dropout2d_on_device_drops_whole_channels_and_matches_the_oracle():
    """dropoutType != 'Dropout' -> nn.Dropout2d (EEGNet_tor.py:21): the Philox draw is per (sample, channel) row,
    forward and backward agree, and with the masks read back from the saved activations the result equals the
    CPU oracle."""
    import eegnet_oracle as EO
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    torch.manual_seed(3)
    model = EEGNet_tor(5, dropoutType='SpatialDropout2D')
    assert isinstance(model.dropout, torch.nn.Dropout2d) and model._dims.dropout2d
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    model = model.cuda().train()
    x, y = _data(8, 1)
    out = model(x.cuda())
    loss = torch.nn.functional.cross_entropy(out, y.cuda())
    loss.backward()
    eng = model._engine(8)
    d1, feat = eng.saved("d1").cpu(), eng.saved("feat").cpu().reshape(8, 64, 15)
    keep1, keep2 = (d1 != 0).any(-1), (feat != 0).any(-1)                   # (B, 64): whole rows kept or dropped
    assert torch.equal((d1 != 0).all(-1), keep1)                            # kept rows are non-zero everywhere (ELU mean != 0)
    assert 0.3 < keep1.float().mean() < 0.7 and 0.3 < keep2.float().mean() < 0.7
    m1 = keep1.reshape(8, 64, 1, 1).expand(8, 64, 1, 125).float()
    m2 = keep2.reshape(8, 64, 1, 1).expand(8, 64, 1, 15).float()
    params, buffers = EO.split_state(sd, "tor")
    o = EO.tor_forward(params, buffers, x, True, masks=[m1, m2])
    l = EO.loss_fn(o, y)
    l.backward()
    assert abs(loss.item() - l.item()) < TOL * abs(l.item())
    for k, p in model.named_parameters():
        a, b = p.grad.cpu().numpy(), params[k].grad.numpy()
        rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12)
        assert rel < TOL or np.abs(a - b).max() < 1e-7, (k, rel)


def test_module_survives_deepcopy_and_torch_save_after_a_forward():
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    torch.manual_seed(0)
    model = EEGNet_tor(5).cuda().eval()
    x, _ = _data(4)
    ref = model(x.cuda()).detach().clone()
    snap = copy.deepcopy(model)                        # best-model snapshot idiom
    buf = io.BytesIO()
    torch.save(model, buf)
    buf.seek(0)
    loaded = torch.load(buf, weights_only=False)
    with torch.no_grad():
        model.dense.bias.add_(1.0)                     # the copies must not alias the original's arena
    assert torch.equal(snap(x.cuda()), ref) and torch.equal(loaded(x.cuda()), ref)
    assert not torch.equal(model(x.cuda()), ref)


def test_trainer_optimizer_exposes_and_resumes_the_fused_adam_state():
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor, Trainer_uni
    trx, try_ = _data(48, 2)
    tex, tey = _data(16, 3)
    data = [trx.numpy(), try_.numpy(), tex.numpy(), tey.numpy()]

    def make(seed):
        torch.manual_seed(seed)
        return Trainer_uni(EEGNet_tor(5, dropoutRate=0.0), data, lr=1e-3, batch_size=16, num_epochs=1)

    a = make(5)
    assert isinstance(a.optimizer, torch.optim.Adam)
    rows = [torch.arange(16 * i, 16 * i + 16, dtype=torch.int32).cuda() for i in range(3)]
    a.model.train()
    a._fused_train_step(rows[0]); a._fused_train_step(rows[1])
    sd_opt, sd_model = a.optimizer.state_dict(), {k: v.clone() for k, v in a.model.state_dict().items()}
    st = sd_opt["state"]
    assert len(st) == 11 and all(float(s["step"]) == 2.0 for s in st.values())
    assert all(s["exp_avg"].abs().sum() > 0 and s["exp_avg_sq"].abs().sum() > 0 for s in st.values())
    sd_opt = copy.deepcopy(sd_opt)
    la = float(a._fused_train_step(rows[2]))
    # a fresh trainer resumed from (model, optimizer) checkpoints takes the SAME third step
    b = make(99)
    b.model.load_state_dict(sd_model)
    b.optimizer.load_state_dict(sd_opt)
    b.model.train()
    lb = float(b._fused_train_step(rows[2]))
    assert la == lb
    assert torch.equal(a._core.params, b._core.params) and torch.equal(a._core.exp_avg_sq, b._core.exp_avg_sq)
    # an LR edit (what a scheduler does) reaches the fused step
    before = a._core.params.clone()
    a.optimizer.param_groups[0]["lr"] = 0.0
    a._fused_train_step(rows[0])
    assert torch.equal(a._core.params, before) and a._core.lr == 0.0


def test_second_device_if_present():
    """Trainer_uni(device='cuda:1') in a process whose current device is 0 (needs 2 GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    torch.cuda.set_device(0)
    torch.manual_seed(0)
    m0 = EEGNet_tor(5).cuda(0).eval()
    torch.manual_seed(0)
    m1 = EEGNet_tor(5).to("cuda:1").eval()
    x, _ = _data(4)
    assert torch.equal(m0(x.cuda(0)).cpu(), m1(x.to("cuda:1")).cpu())
