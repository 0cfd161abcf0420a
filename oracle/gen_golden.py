"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (through
oracle/ref_shim.py) in the build container.  The reference ships no golden vectors
(SURVEY.md section 4), so these files are the pins for oracle/ and for the CUDA path.

Run:  python oracle/gen_golden.py        (needs /root/reference; CPU only, ~1 min)

Every fixture stores the library versions it was produced with.
"""
import io
import os
import sys
import contextlib

import numpy as np
import scipy
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import eeg_oracle as O  # noqa: E402
import golden_inputs as GI  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
VERS = dict(scipy=scipy.__version__, numpy=np.__version__, torch=torch.__version__)


def save(name, **arrs):
    arrs["_versions"] = np.array(repr(VERS))
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.0f} KiB")


def ref_preproc(ns, raw, label, band, upto="segment"):
    """Run the reference's DataLoadEEG stages on raw [trials][ch][time] (values upcast
    to f64 exactly, presented in the .mat's (time, ch, trials) F-order view)."""
    D = ns.Dataload_eeg.DataLoadEEG(subject=1, band=band, fs_orig=500, fs_target=100)
    cnt = np.transpose(raw.astype(np.float64), (2, 1, 0))        # (time, ch, trials), F-contiguous
    D.label = label
    D.seg = np.transpose(cnt, [1, 0, 2])                          # Dataload_eeg.py:82
    D.downsampling()
    dec = D.seg.copy()
    D.bandpass_filter()
    if upto == "filter":
        return dec, D.seg_f
    D.segment_and_select_classes()
    return dec, D.seg_f_div, D.label_div


def gen_preproc(ns):
    # (1) small free-shape case: downsampling + bandpass only (segment hard-codes 30/500/4/200)
    rng = np.random.default_rng(7)
    raw = rng.standard_normal((6, 4, 1000)).astype(np.float32)
    raw += (3.0 * np.sin(2 * np.pi * 0.2 * np.arange(6000) / 500.0)).reshape(6, 1, 1000).astype(np.float32)
    out = {"raw": raw}
    for tag, band in (("b0545", [0.5, 45]), ("b0530", [5, 30])):
        dec, filt = ref_preproc(ns, raw, None, band, upto="filter")
        out["dec"] = dec                      # (4, 200, 6)  (ch, t, trials)
        out["filt_" + tag] = filt
    save("preproc_small.npz", **out)

    # (2) dataset-shaped subject 1: digest of the reference's prepare_data output
    raw, label = O.synth_subject(1)
    dec, x, y = ref_preproc(ns, raw, label, [0.5, 45])
    save("preproc_subject1_digest.npz",
         label=label.astype(np.uint8), y=y.astype(np.int64),
         x_sub=x[::25, ::7, ::20].copy(),              # strided sample of (400, 30, 500)
         x_epoch_sum=x.sum(axis=(1, 2)), x_chan_rms=np.sqrt((x ** 2).mean(axis=(0, 2))),
         raw_checksum=np.array([float(raw.astype(np.float64).sum()), float(np.abs(raw).astype(np.float64).sum())]),
         dec_sub=dec[::7, ::50, ::25].copy())
    # (3) index-coded run of segment_and_select_classes: proves epoch e=4k+q <- trial k, [500q,500q+500)
    D = ns.Dataload_eeg.DataLoadEEG()
    c, t, k = np.meshgrid(np.arange(30), np.arange(2000), np.arange(200), indexing="ij")
    D.seg_f = (c * 2000 * 200 + t * 200 + k).astype(np.float64)
    D.label = label
    D.segment_and_select_classes()
    code = D.seg_f_div.astype(np.int64)                 # (400, 30, 500) of source codes
    src_c, rem = code // (2000 * 200), code % (2000 * 200)
    src_t, src_k = rem // 200, rem % 200
    assert (src_c == np.arange(30)[None, :, None]).all()
    assert (src_k == src_k[:, :1, :1]).all() and ((src_t - np.arange(500)[None, None, :]) == (src_t[:, :1, :1])).all()
    save("segment_plan.npz", label=label.astype(np.uint8), y=D.label_div.astype(np.int64),
         src_trial=src_k[:, 0, 0].copy(), src_t0=src_t[:, 0, 0].copy())

    # (4) split: reference get_split on index-coded features, shipped labels and remapped labels
    for tag, yy in (("shipped", y), ("remap", (y - 1) // 2)):
        xs = np.arange(yy.size, dtype=np.float64).reshape(-1, 1, 1) * np.ones((1, 2, 3))
        outs = {}
        for h in (40, 56):
            sp = ns.EAV_datasplit.EAVDataSplit(xs, yy)
            trx, try_, tex, tey = sp.get_split(h_idx=h)
            outs[f"tr_idx_{h}"] = trx[:, 0, 0].astype(np.int64)
            outs[f"te_idx_{h}"] = tex[:, 0, 0].astype(np.int64)
            outs[f"tr_y_{h}"] = try_
            outs[f"te_y_{h}"] = tey
        save(f"split_{tag}.npz", y=yy, **outs)


def grads_of(model, names):
    sd = dict(model.named_parameters())
    return {f"grad::{k}": sd[k].grad.detach().numpy().copy() for k in names}


def gen_eegnet_tor(ns):
    import eegnet_oracle as EO
    M = ns.EEGNet_tor
    for tag, B, scale in (("b8", 8, 1.0), ("b8_renorm", 8, 4.0)):
        torch.manual_seed(3)
        model = ref_shim.make_eegnet_tor(ns, 5)
        if scale != 1.0:
            with torch.no_grad():
                model.depthwiseConv.weight.mul_(scale)
                model.dense.weight.mul_(scale)
        # non-trivial BN affine + running stats so eval mode is exercised properly
        g = torch.Generator().manual_seed(11)
        with torch.no_grad():
            for bn in (model.firstBN, model.depthwiseBN, model.separableBN):
                bn.weight.copy_(1 + 0.2 * torch.randn(bn.weight.shape, generator=g))
                bn.bias.copy_(0.1 * torch.randn(bn.bias.shape, generator=g))
                bn.running_mean.copy_(0.05 * torch.randn(bn.bias.shape, generator=g))
                bn.running_var.copy_(1 + 0.3 * torch.rand(bn.bias.shape, generator=g))
        init = {f"init::{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
        x = torch.randn(B, 1, 30, 500, generator=g)
        y = torch.randint(0, 5, (B,), generator=g)
        out = dict(init, x=x.numpy(), y=y.numpy())
        crit = torch.nn.CrossEntropyLoss()
        for mode in ("train", "eval"):
            model.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in init.items()})
            model.train(mode == "train")
            model.zero_grad()
            torch.manual_seed(100)
            # record the dropout masks the reference draws (order: mask1, mask2) by replaying the RNG
            if mode == "train":
                m1 = torch.empty(B, 64, 1, 125).bernoulli_(0.5)
                m2 = torch.empty(B, 64, 1, 15).bernoulli_(0.5)
                out["mask1"], out["mask2"] = m1.numpy().astype(np.uint8), m2.numpy().astype(np.uint8)
                torch.manual_seed(100)
            p = model(x)
            loss = crit(p, y)
            loss.backward()
            out[f"{mode}::probs"] = p.detach().numpy()
            out[f"{mode}::loss"] = np.array(loss.item(), dtype=np.float64)
            for k, v in grads_of(model, EO.TOR_PARAMS).items():
                out[f"{mode}::{k}"] = v
            for k, v in model.state_dict().items():
                if "running" in k or "num_batches" in k or k in ("depthwiseConv.weight", "dense.weight"):
                    out[f"{mode}::after::{k}"] = v.detach().numpy().copy()
        save(f"eegnet_tor_{tag}.npz", **out)

    # Adam trajectory: 4 steps in eval-mode BN (steady state, F5) + 2 steps train-mode, B=8
    torch.manual_seed(5)
    model = ref_shim.make_eegnet_tor(ns, 5)
    xs, ys = GI.adam6_inputs()
    out = {f"init::{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    out["input_checksum"] = GI.checksum(xs.numpy(), ys.numpy())
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()
    losses, masks1, masks2 = [], [], []
    for i in range(6):
        model.train(i < 2)
        if i < 2:
            torch.manual_seed(200 + i)
            masks1.append(torch.empty(8, 64, 1, 125).bernoulli_(0.5).numpy().astype(np.uint8))
            masks2.append(torch.empty(8, 64, 1, 15).bernoulli_(0.5).numpy().astype(np.uint8))
            torch.manual_seed(200 + i)
        loss = crit(model(xs[i]), ys[i])
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    out["losses"] = np.array(losses)
    out["masks1"], out["masks2"] = np.stack(masks1), np.stack(masks2)
    for k, v in model.state_dict().items():
        out[f"final::{k}"] = v.detach().numpy().copy()
    save("eegnet_tor_adam6.npz", **out)

    # Trainer_uni.train(): 3 epochs, N_train=40, N_test=16, bs=16 -- per-step loss + val loss/acc
    torch.manual_seed(9)
    model = ref_shim.make_eegnet_tor(ns, 5)
    trx, try_, tex, tey = GI.trainer_inputs()
    out = {f"init::{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    out["input_checksum"] = GI.checksum(trx, try_, tex, tey)
    trainer = M.Trainer_uni(model, [trx, try_, tex, tey], lr=1e-3, batch_size=16, num_epochs=3,
                            device=torch.device("cpu"))
    crit0 = trainer.criterion
    step_losses = []

    class Rec(torch.nn.Module):
        def forward(self, s, t):
            l = crit0(s, t)
            step_losses.append((float(l.detach()), bool(torch.is_grad_enabled())))
            return l
    trainer.criterion = Rec()
    torch.manual_seed(77)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        trainer.train()
    out["train_step_loss"] = np.array([l for l, ge in step_losses if ge])
    out["val_batch_loss"] = np.array([l for l, ge in step_losses if not ge])
    out["stdout"] = np.array(buf.getvalue())
    for k, v in model.state_dict().items():
        out[f"final::{k}"] = v.detach().numpy().copy()
    save("trainer_uni_3ep.npz", **out)


def gen_cnn_eeg(ns):
    import eegnet_oracle as EO
    C = ns.CNN_EEG
    for tag, kw, B in (("default", dict(nb_classes=4, Chans=64, Samples=128, dropoutRate=0.25), 8),
                       ("eav", dict(nb_classes=5, Chans=30, Samples=500, kernLength=300, D=8, F2=64), 4)):
        torch.manual_seed(4)
        model = C.EEGNet(**kw)
        g = torch.Generator().manual_seed(21)
        with torch.no_grad():
            for bn in (model.block1[1], model.block1[3], model.block2[2]):
                bn.weight.copy_(1 + 0.2 * torch.randn(bn.weight.shape, generator=g))
                bn.bias.copy_(0.1 * torch.randn(bn.bias.shape, generator=g))
                bn.running_mean.copy_(0.05 * torch.randn(bn.bias.shape, generator=g))
                bn.running_var.copy_(1 + 0.3 * torch.rand(bn.bias.shape, generator=g))
        init = {f"init::{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
        x = torch.randn(B, kw["Chans"], kw["Samples"], generator=g)    # 3-D input (CNN_EEG.py:60-61)
        y = torch.randint(0, kw["nb_classes"], (B,), generator=g)
        out = dict(init, x=x.numpy(), y=y.numpy())
        crit = torch.nn.CrossEntropyLoss()
        p_drop = kw.get("dropoutRate", 0.5)
        F2 = kw.get("F2", 16)
        DF1 = kw.get("D", 2) * 8
        T4 = kw["Samples"] // 4
        for mode in ("train", "eval"):
            model.load_state_dict({k[6:]: torch.from_numpy(v) for k, v in init.items()})
            model.train(mode == "train")
            model.zero_grad()
            torch.manual_seed(100)
            if mode == "train":
                out["mask1"] = torch.empty(B, DF1, 1, T4).bernoulli_(1 - p_drop).numpy().astype(np.uint8)
                out["mask2"] = torch.empty(B, F2, 1, T4 // 8).bernoulli_(1 - p_drop).numpy().astype(np.uint8)
                torch.manual_seed(100)
            o = model(x)
            loss = crit(o, y)
            loss.backward()
            out[f"{mode}::logits"] = o.detach().numpy()
            out[f"{mode}::loss"] = np.array(loss.item(), dtype=np.float64)
            for k, v in grads_of(model, EO.CNN_PARAMS).items():
                out[f"{mode}::{k}"] = v
            for k, v in model.state_dict().items():
                if "running" in k or "num_batches" in k:
                    out[f"{mode}::after::{k}"] = v.detach().numpy().copy()
        save(f"cnn_eeg_{tag}.npz", **out)


def ref_legacy_functions():
    """`Bandpass` and `mysplit` of CNN_tensorflow/CNN_EEG_tf.py, compiled from the reference file's own source.
    The module cannot be imported (it needs TensorFlow and runs its training script at import time), so only these
    two FunctionDef nodes are executed, unmodified, with numpy/scipy in their globals."""
    import ast
    path = "/root/reference/CNN_tensorflow/CNN_EEG_tf.py"
    src = open(path).read()
    tree = ast.parse(src)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("Bandpass", "mysplit")]
    assert len(wanted) == 2
    from scipy import signal
    from scipy.signal import butter
    g = {"np": np, "signal": signal, "butter": butter}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), g)
    return g["Bandpass"], g["mysplit"]


def gen_legacy():
    """Legacy order (band-pass at 500 Hz, then decimate) on dataset-shaped subject 1: the reference's own
    Bandpass()/mysplit() plus the module-level glue of CNN_EEG_tf.py:180-206, which is restated line by line."""
    from scipy import signal
    Bandpass, mysplit = ref_legacy_functions()
    raw, label = O.synth_subject(1)
    cnt_ = np.transpose(raw.astype(np.float64), (2, 1, 0))               # (10000, 30, 200) as loadmat returns it
    Label = label
    cnt_f = Bandpass(cnt_, freq=[3, 50], fs=500)                            # :180
    tm = np.transpose(cnt_f, [0, 2, 1]).reshape([10000 * 200, 30], order='F')      # :185
    tm2 = signal.resample_poly(tm, up=1, down=int(500 / 100), axis=0)      # :188
    cnt_f2 = np.reshape(tm2, [2000, 200, 30], order='F')                   # :189
    cnt_seg = np.transpose(cnt_f2, [1, 0, 2])                               # :191
    cnt_seg_split, Label_split = mysplit(cnt_seg, Label)                    # :194
    dat = np.transpose(cnt_seg_split, (0, 2, 1)).reshape((800, 30, 500, 1))  # :196
    selected_classes = [1, 3, 5, 7, 9]
    selected_indices = np.isin(np.argmax(Label_split, axis=0), selected_classes)   # :201
    data_5class = dat[selected_indices]                                     # :203
    aa = Label_split[:, selected_indices]
    label_5class = aa[selected_classes, :]                                  # :206
    x = data_5class[..., 0]                                                 # (400, 30, 500)
    save("preproc_legacy_subject1_digest.npz",
         label=label.astype(np.uint8), onehot=label_5class.astype(np.uint8),
         x_sub=x[::25, ::7, ::20].copy(), x_epoch_sum=x.sum(axis=(1, 2)),
         x_chan_rms=np.sqrt((x ** 2).mean(axis=(0, 2))),
         filt500_sub=cnt_f[::500, ::7, ::25].copy())


def gen_shallow():
    """Transformer_torch/Transformer_EEG.py ShallowConvNet: the unmodified reference model, imported
    from /root/reference, forward in eval mode and forward+backward in train mode with its dropouts replaced by
    recorded masks (torch's own RNG stream, captured with forward hooks)."""
    import importlib
    sys.path.insert(0, os.path.join(ref_shim.REF_ROOT, "Transformer_torch"))
    TE = importlib.import_module("Transformer_EEG")
    torch.manual_seed(11)
    model = TE.ShallowConvNet(nb_classes=5)
    with torch.no_grad():                      # move the parameters off their defaults so every term matters
        for n, p_ in model.named_parameters():
            if n.endswith("norm1.bias") or n.endswith("norm2.bias") or n == "bn.bias":
                p_.normal_(0, 0.1)
        model.bn.running_mean.normal_(0, 0.2)
        model.bn.running_var.uniform_(0.5, 1.5)
    init = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 1, 30, 500, generator=g)
    y = torch.randint(0, 5, (4,), generator=g)
    out = {"x": x.numpy(), "y": y.numpy()}
    for k, v in init.items():
        out["init::" + k] = v.numpy()
    model.eval()
    with torch.no_grad():
        out["eval::probs"] = model(x).numpy()
    # train mode: record the mask each nn.Dropout applies (output / input where input != 0)
    model.train()
    masks = []

    def hook(mod, inp, outp):
        i = inp[0]
        m = torch.where(i != 0, outp / torch.where(i != 0, i, torch.ones_like(i)), torch.full_like(i, 2.0) * (outp != 0))
        # where the input is exactly 0 the mask is unobservable and irrelevant (0 * m = 0); store 0 there
        masks.append(torch.where(i != 0, m, torch.zeros_like(i)).detach().clone())

    hs = [m.register_forward_hook(hook) for m in model.modules() if isinstance(m, torch.nn.Dropout)]
    torch.manual_seed(5)
    probs = model(x)
    loss = torch.nn.CrossEntropyLoss()(probs, y)
    loss.backward()
    for h in hs:
        h.remove()
    out["train::probs"] = probs.detach().numpy()
    out["train::loss"] = np.array(float(loss))
    for i, m in enumerate(masks):
        out[f"train::mask{i:02d}"] = np.packbits((m != 0).numpy().reshape(-1))
        out[f"train::mask{i:02d}_shape"] = np.array(m.shape)
    grads = {n: p_.grad for n, p_ in model.named_parameters()}
    keep = ["conv.weight", "embedding.value_proj.0.weight", "embedding.value_proj.39.weight", "fc.weight", "bn.weight",
            "bn.bias"] + [f"transformer.{l}.{n}" for l in (0, 11) for n in
                          ("attn.W_q.weight", "attn.W_k.weight", "attn.W_v.weight", "norm1.weight", "norm1.bias",
                           "ffn.net.0.weight", "ffn.net.0.bias", "ffn.net.3.weight", "ffn.net.3.bias", "norm2.weight")]
    for n in keep:
        out["train::grad::" + n] = grads[n].numpy()
    out["train::grad_l2_all"] = np.array([float(grads[n].norm()) for n, _ in model.named_parameters()])
    out["train::bn_running_mean"] = model.bn.running_mean.numpy()
    out["train::bn_running_var"] = model.bn.running_var.numpy()
    save("shallowconvnet_b4.npz", **out)


def gen_bench_config(ns):
    """The benchmark's own configuration through the UNMODIFIED reference: one full subject,
    280 / 120 epochs, batch 32 (9 steps per epoch, the last one ragged: 24), lr 1e-5, 2 epochs -- epoch 1 in train
    mode, epoch 2 in eval mode.  Records every step's batch rows, dropout masks and loss, the validation
    batch losses and the printed lines."""
    M = ns.EEGNet_tor
    torch.manual_seed(21)
    model = ref_shim.make_eegnet_tor(ns, 5)
    trx, try_, tex, tey = GI.bench_subject_inputs()
    out = {f"init::{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    out["input_checksum"] = GI.checksum(trx, try_, tex, tey)
    trainer = M.Trainer_uni(model, [trx, try_, tex, tey], lr=1e-5, batch_size=32, num_epochs=2,
                            device=torch.device("cpu"))
    crit0 = trainer.criterion
    step_losses, batch_rows, masks1, masks2 = [], [], [], []
    marker = {float(v): i for i, v in enumerate(trx[:, 0, 0, 0])}
    assert len(marker) == 280

    def pre_hook(mod, args):
        if torch.is_grad_enabled():
            batch_rows.append(np.array([marker[float(v)] for v in args[0][:, 0, 0, 0]], dtype=np.int32))
    model.register_forward_pre_hook(pre_hook)

    def drop_hook(mod, inp, outp):          # keep-mask of this call, recovered from the reference's own output
        if mod.training:
            (masks1 if outp.shape[-1] == 125 else masks2).append((outp != 0).numpy().reshape(outp.shape[0], 64, -1))
    model.dropout.register_forward_hook(drop_hook)

    class Rec(torch.nn.Module):
        def forward(self, s, t):
            l = crit0(s, t)
            step_losses.append((float(l.detach()), bool(torch.is_grad_enabled())))
            return l
    trainer.criterion = Rec()
    torch.manual_seed(78)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        trainer.train()
    out["train_step_loss"] = np.array([l for l, ge in step_losses if ge])
    out["val_batch_loss"] = np.array([l for l, ge in step_losses if not ge])
    assert out["train_step_loss"].shape == (18,) and out["val_batch_loss"].shape == (8,)
    assert [len(r) for r in batch_rows] == ([32] * 8 + [24]) * 2 and len(masks1) == 9 and len(masks2) == 9
    out["batch_rows"] = np.concatenate(batch_rows)
    # ELU'd + pooled activations are never exactly 0, so (output != 0) IS the keep-mask
    out["mask1_bits"] = np.packbits(np.concatenate([m.reshape(-1) for m in masks1]))
    out["mask2_bits"] = np.packbits(np.concatenate([m.reshape(-1) for m in masks2]))
    out["stdout"] = np.array(buf.getvalue())
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k or k in ("firstBN.weight", "dense.bias"):
            out[f"final::{k}"] = v.detach().numpy().copy()
    save("trainer_uni_bench_2ep.npz", **out)

    # EEGNetTrainer.train() (CNN_EEG.py:70-146): 2 epochs, model.train() every epoch, logits + CE, Adam lr 1e-3
    C = ns.CNN_EEG
    torch.manual_seed(22)
    model = C.EEGNet(nb_classes=4, Chans=64, Samples=128, dropoutRate=0.25)
    trx, try_, tex, tey = GI.cnn_trainer_inputs()
    out = {f"init::{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}
    out["input_checksum"] = GI.checksum(trx.numpy(), try_.numpy(), tex.numpy(), tey.numpy())
    from torch.utils.data import TensorDataset
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tr = C.EEGNetTrainer(model, TensorDataset(trx, try_), TensorDataset(tex, tey), batch_size=32, epochs=2, lr=1e-3)
        tr.device = torch.device("cpu")
        tr.model.to("cpu")
        torch.manual_seed(79)
        tl1 = tr.train_epoch(); v1 = tr.validate_epoch()
        tl2 = tr.train_epoch(); v2 = tr.validate_epoch()
        pred = tr.predict()
    out["train_loss"] = np.array([tl1, tl2])
    out["val_loss"] = np.array([v1[0], v2[0]])
    out["val_acc"] = np.array([v1[1], v2[1]])
    out["predict"] = np.array(pred, dtype=np.int64)
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            out[f"final::{k}"] = v.detach().numpy().copy()
    save("cnn_eeg_trainer_2ep.npz", **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if "--only-bench-config" in sys.argv:
        gen_bench_config(ref_shim.load())
        sys.exit(0)
    if "--only-legacy" in sys.argv:
        gen_legacy()
        sys.exit(0)
    if "--only-shallow" in sys.argv:
        gen_shallow()
        sys.exit(0)
    ns = ref_shim.load()
    gen_preproc(ns)
    gen_legacy()
    gen_eegnet_tor(ns)
    gen_cnn_eeg(ns)
    gen_shallow()
    gen_bench_config(ns)
