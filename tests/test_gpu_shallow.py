"""-m gpu parity of the ShallowConvNet CUDA path (csrc/shallow.cu behind the drop-in
eav_b200.Transformer_torch.Transformer_EEG) against the fixture written by the UNMODIFIED reference
(Transformer_torch/Transformer_EEG.py:107-148): eval forward, train forward + loss + every recorded gradient with
the reference's own dropout masks, BatchNorm running statistics; then the reference's training loop end to end."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _golden_masks(g):
    parts, i = [], 0
    while f"train::mask{i:02d}" in g.files:
        shape = tuple(int(v) for v in g[f"train::mask{i:02d}_shape"])
        parts.append(np.unpackbits(g[f"train::mask{i:02d}"])[:int(np.prod(shape))])
        i += 1
    assert i == 37                                   # 12 layers x 3 dropouts + the head
    return torch.from_numpy(np.concatenate(parts).astype(np.uint8))


def test_shallowconvnet_forward_backward_vs_reference(golden):
    import gpu_util as U
    from eav_b200.Transformer_torch.Transformer_EEG import ShallowConvNet
    g = golden("shallowconvnet_b4.npz")
    model = ShallowConvNet(nb_classes=5)
    model.load_state_dict({k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("init::")})
    model = model.cuda()
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    model.eval()
    with torch.no_grad():
        pe = model(x)
    assert pe.shape == (4, 5)
    assert U.rel_max(pe.cpu().numpy(), g["eval::probs"]) < TOL
    # train mode with the masks the reference itself drew
    masks = _golden_masks(g).cuda()
    model._draw_masks = lambda B, dev: masks
    model.train()
    probs = model(x)
    loss = torch.nn.CrossEntropyLoss()(probs, y)
    loss.backward()
    assert U.rel_max(probs.detach().cpu().numpy(), g["train::probs"]) < TOL
    assert abs(float(loss) - float(g["train::loss"])) < TOL * float(g["train::loss"])
    grads = {n: p.grad.detach().cpu().numpy() for n, p in model.named_parameters()}
    for k in g.files:
        if k.startswith("train::grad::"):
            name = k[len("train::grad::"):]
            assert U.rel_l2(grads[name], g[k]) < TOL, (name, U.rel_l2(grads[name], g[k]))
    l2 = np.array([np.linalg.norm(grads[n]) for n, _ in model.named_parameters()])
    ref = g["train::grad_l2_all"]
    assert l2.shape == ref.shape
    assert np.abs(l2 - ref).max() < 2e-4 * ref.max() and np.allclose(l2, ref, rtol=5e-3, atol=1e-6 * ref.max())
    assert np.allclose(model.bn.running_mean.cpu().numpy(), g["train::bn_running_mean"], rtol=1e-4, atol=1e-6)
    assert np.allclose(model.bn.running_var.cpu().numpy(), g["train::bn_running_var"], rtol=1e-4, atol=1e-6)
    assert int(model.bn.num_batches_tracked) == 1


def test_shallowconvnet_trainer_learns(tmp_path):
    """The reference's loop (Transformer_EEG.py:182-204) unmodified in structure: Adam steps through autograd, the
    per-step max-norm on fc.weight, validate(), the results file."""
    from eav_b200.Transformer_torch.Transformer_EEG import ShallowConvNet, TrainerUni
    g = torch.Generator().manual_seed(0)
    w = torch.randn(3, 30, generator=g)

    def make(n):
        yy = torch.randint(0, 3, (n,), generator=g)
        xx = torch.randn(n, 1, 30, 500, generator=g) + 1.5 * w[yy].reshape(n, 1, 30, 1) * torch.sin(torch.arange(500) * 0.3)
        return xx, yy
    trx, try_ = make(48)
    tex, tey = make(24)
    torch.manual_seed(1)
    model = ShallowConvNet(nb_classes=3, num_layers=2, dropout=0.1)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with contextlib.redirect_stdout(io.StringIO()) as buf:
            tr = TrainerUni(model, [trx, try_, tex, tey], lr=2e-3, batch_size=16, epochs=6, subject=7)
            model.cuda().eval()
            with torch.no_grad():
                l0 = float(torch.nn.functional.cross_entropy(model(trx.cuda()), try_.cuda()))
            tr.train()
            model.eval()
            with torch.no_grad():
                l1 = float(torch.nn.functional.cross_entropy(model(trx.cuda()), try_.cuda()))
        assert l1 < l0
        assert buf.getvalue().count("Validation Accuracy") == 6
        assert "Subject 7 | Accuracy" in open("eeg_results_new_shallow_.txt").read()
        assert float(model.fc.weight.norm(dim=1).max()) <= 0.5 + 1e-5
    finally:
        os.chdir(cwd)


def test_shallowconvnet_has_no_cpu_path():
    from eav_b200.Transformer_torch.Transformer_EEG import ShallowConvNet
    with pytest.raises(RuntimeError):
        ShallowConvNet(5)(torch.zeros(2, 1, 30, 500))
