"""Host-side logic of the two multi-GPU modes, on CPU: subject sharding (with a real
world_size-2 gloo process group for the control-plane gather) and the data-parallel batch split."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_subject_assignment_matches_survey():
    from eav_b200.sharding import shard_sizes, subjects_for_rank
    subs = list(range(1, 43))
    assert shard_sizes(42, 8) == [6, 6, 5, 5, 5, 5, 5, 5]          # best speed-up 7.0x
    assert shard_sizes(42, 4) == [11, 11, 10, 10]
    assert shard_sizes(42, 2) == [21, 21]
    assert shard_sizes(42, 1) == [42]
    seen = sorted(s for r in range(8) for s in subjects_for_rank(subs, r, 8))
    assert seen == subs                                             # a partition: nobody lost, nobody twice
    assert subjects_for_rank(subs, 3, 8) == [4, 12, 20, 28, 36]


def test_split_batch():
    from eav_b200.data_parallel import split_batch
    assert split_batch(8192, 8) == 1024 and split_batch(32, 2) == 16
    with pytest.raises(ValueError):
        split_batch(30, 4)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eav_b200.sharding import gather_results, subjects_for_rank
    mine = subjects_for_rank(range(1, 12), rank, world)
    local = {s: 0.5 + 0.01 * s for s in mine}                       # stand-in for per-subject accuracy
    res = gather_results(local)
    # the data path has no collective; a barrier here only proves the group is healthy
    dist.barrier()
    q.put((rank, mine, res))
    dist.destroy_process_group()


def test_gather_results_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    (r0, mine0, res0), (r1, mine1, res1) = got
    assert mine0 == [1, 3, 5, 7, 9, 11] and mine1 == [2, 4, 6, 8, 10]
    assert res1 is None
    assert list(res0) == list(range(1, 12)) and res0[7] == pytest.approx(0.57)
