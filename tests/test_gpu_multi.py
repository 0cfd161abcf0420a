"""-m gpu tests of the two multi-GPU modes.  Subject sharding needs no collective, so its
kernel path is fully exercised on one GPU; the data-parallel mode is checked with dp_world=1
on one GPU and, when two GPUs are visible, with a real 2-rank NCCL group against the
single-device result at the global batch."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _subject(s, n_tr=40, n_te=16):
    g = torch.Generator().manual_seed(100 + s)
    w = torch.randn(5, 30, generator=g)
    def make(n):
        y = torch.randint(0, 5, (n,), generator=g)
        x = torch.randn(n, 30, 500, generator=g) + 0.8 * w[y].unsqueeze(-1) * torch.sin(torch.arange(500) * 0.2)
        return x, y
    trx, try_ = make(n_tr)
    tex, tey = make(n_te)
    return trx, try_, tex, tey


def test_lockstep_subjects_equal_individually_trained_subjects():
    from eav_b200.sharding import train_subjects
    kw = dict(nb_classes=5, lr=1e-3, batch_size=16, num_epochs=3, model_kwargs=dict(dropoutRate=0.0))
    acc_all, loss_all = train_subjects([1, 2, 3], _subject, **kw)
    for s in (1, 2, 3):
        acc_s, loss_s = train_subjects([s], _subject, **kw)
        # same kernels, but the split-K plans (hence fp32 summation orders) depend on M, and three epochs of
        # Adam amplify those rounding differences: compare the loss curves at 1e-3, not bitwise
        assert np.allclose(loss_all[s], loss_s[s], rtol=1e-3), (s, loss_all[s], loss_s[s])
        assert abs(acc_all[s] - acc_s[s]) <= 1 / 16 + 1e-9
    assert all(l[-1] < l[0] for l in loss_all.values())            # every subject's model learns


def test_lockstep_matches_dropin_trainer_losses(golden):
    """train_subjects (M models, own permutations) and Trainer_uni (M=1, DataLoader order) run the
    same kernels; with the same batches they must produce the same losses."""
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    from eav_b200.ops import EegnetDims
    from eav_b200.trainer_core import SubjectBatchTrainer
    trx, try_, _, _ = _subject(7)
    torch.manual_seed(7)
    model = EEGNet_tor(5, dropoutRate=0.0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    core = SubjectBatchTrainer(model._dims, 1, trx.cuda(), try_.cuda(), lr=1e-3, max_batch=16)
    core.load_state_dicts([sd], EEGNet_tor._BN_NAMES)
    model = model.cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()
    model.train()
    for step in range(3):
        rows = torch.arange(16 * step, 16 * step + 16) % 40
        l_core = core.train_step(rows.int().cuda(), bn_train=True)
        opt.zero_grad()
        l_mod = crit(model(trx[rows].unsqueeze(1).cuda()), try_[rows].cuda())
        l_mod.backward()
        opt.step()
        assert abs(float(l_core[0]) - float(l_mod)) < TOL * float(l_mod), step


def _state(golden):
    g = golden("eegnet_tor_b8.npz")
    return g, {k[6:]: torch.from_numpy(np.array(g[k])) for k in g.files if k.startswith("init::")}


@pytest.mark.parametrize("train", [True, False])
def test_data_parallel_driver_world1_equals_engine(golden, train):
    import gpu_util as U
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    from eav_b200.data_parallel import DataParallelEEGNet
    from eav_b200.ops import EegnetDims, EegnetEngine
    g, sd = _state(golden)
    dims = EegnetDims(5)
    x = torch.from_numpy(g["x"]).cuda().reshape(8, 30, 500).contiguous()
    y = torch.from_numpy(g["y"]).cuda()
    m1 = torch.from_numpy(g["mask1"]).cuda().reshape(8, 64, 125).contiguous()
    m2 = torch.from_numpy(g["mask2"]).cuda().reshape(8, 64, 15).contiguous()
    dp = DataParallelEEGNet(dims, 8, lr=1e-3, state_dict=sd, bn_names=EEGNet_tor._BN_NAMES)
    loss = dp.step(x, y, bn_train=train, masks=(m1, m2) if train else None, update=False)
    mode = "train" if train else "eval"
    assert abs(float(loss) - float(g[f"{mode}::loss"])) < TOL * float(g[f"{mode}::loss"])
    gd = U.unpack(dims, dp.grads[0])
    for k, v in gd.items():
        assert U.rel_l2(v.numpy(), g[f"{mode}::grad::{k}"]) < TOL, k


def _dp_worker(rank, world, port, sd, x, y, m1, m2, q, collective="auto"):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from eav_b200.CNN_torch.EEGNet_tor import EEGNet_tor
    from eav_b200.data_parallel import DataParallelEEGNet
    from eav_b200.ops import EegnetDims
    dims = EegnetDims(5)
    B = x.shape[0] // world
    sl = slice(rank * B, (rank + 1) * B)
    out = {}
    for train in (True, False):
        dp = DataParallelEEGNet(dims, x.shape[0], lr=1e-3, state_dict=sd, bn_names=EEGNet_tor._BN_NAMES,
                                collective=collective)
        assert dp.collective == collective, dp.collective
        masks = (m1[sl].cuda().contiguous(), m2[sl].cuda().contiguous()) if train else None
        loss = dp.step(x[sl].cuda().contiguous(), y[sl].cuda().contiguous(), bn_train=train, masks=masks, update=True)
        torch.cuda.synchronize()
        out[train] = (float(loss), dp.grads.cpu(), dp.params.cpu(), dp.bn_state.cpu())
        for _ in range(3):       # a few more steps: the exchange slots and call numbers are reused
            dp.step(x[sl].cuda().contiguous(), y[sl].cuda().contiguous(), bn_train=train, masks=masks, update=True)
        torch.cuda.synchronize()
        out[(train, "later")] = dp.params.cpu()
        if collective == "peer":     # whole step as a CUDA graph (peer all-reduce numbered by a device counter)
            xs, ys = x[sl].cuda().contiguous(), y[sl].cuda().contiguous()
            a = DataParallelEEGNet(dims, x.shape[0], lr=1e-3, state_dict=sd, bn_names=EEGNet_tor._BN_NAMES, collective="peer", seed=5)
            b = DataParallelEEGNet(dims, x.shape[0], lr=1e-3, state_dict=sd, bn_names=EEGNet_tor._BN_NAMES, collective="peer", seed=5)
            for _ in range(4):
                la = a.step(xs, ys, bn_train=train)
                lb = b.step(xs, ys, bn_train=train, graph=True)
            torch.cuda.synchronize()
            out[(train, "graph")] = (a.params.cpu(), b.params.cpu(), float(la), float(lb))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("collective", ["peer", "nccl"])
def test_data_parallel_two_ranks_equal_single_device_global_batch(golden, collective):
    import torch.multiprocessing as mp
    import gpu_util as U
    from eav_b200.ops import EegnetDims
    g, sd = _state(golden)
    dims = EegnetDims(5)
    x = torch.from_numpy(g["x"]).reshape(8, 30, 500).contiguous()
    y = torch.from_numpy(g["y"])
    m1 = torch.from_numpy(g["mask1"]).reshape(8, 64, 125).contiguous()
    m2 = torch.from_numpy(g["mask2"]).reshape(8, 64, 15).contiguous()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, sd, x, y, m1, m2, q, collective)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for train in (True, False):
        mode = "train" if train else "eval"
        l0, g0, p0, b0 = got[0][train]
        l1, g1, p1, b1 = got[1][train]
        assert l0 == l1 and torch.equal(g0, g1) and torch.equal(p0, p1) and torch.equal(b0, b1)   # replicas stay identical
        assert torch.equal(got[0][(train, "later")], got[1][(train, "later")])
        if collective == "peer":
            for r in (0, 1):
                pa, pb, la, lb = got[r][(train, "graph")]
                assert torch.equal(pa, pb) and la == lb          # graph replays == eager steps, bit for bit
            assert torch.equal(got[0][(train, "graph")][1], got[1][(train, "graph")][1])
        assert abs(l0 - float(g[f"{mode}::loss"])) < TOL * float(g[f"{mode}::loss"])
        gd = U.unpack(dims, g0[0])
        for k, v in gd.items():      # == the reference's single-device gradients at the global batch of 8
            assert U.rel_l2(v.numpy(), g[f"{mode}::grad::{k}"]) < TOL, (mode, k)
        if train:
            bn = U.unpack_bn(dims, b0[0])
            for k in g.files:
                if k.startswith("train::after::") and "running" in k:
                    assert np.allclose(bn[k.split("::")[-1]].numpy(), g[k], rtol=1e-5, atol=1e-6), k
