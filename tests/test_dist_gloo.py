"""CPU, world_size 2 over gloo: the multi-GPU plumbing of the path -- contiguous batch sharding,
ONE flat broadcast of the encoder weights + RoPE buffer, max/sum reductions used by bench.py.
(The data path itself has no collective: every image is independent.)"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import naf_b200
    from naf_b200 import dist as ndist

    r, w, _ = ndist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(100 + rank)          # different init on every rank
    model = naf_b200.NAF(dim=64, kernel_size=3).eval()
    before = torch.cat([t.reshape(-1) for t in model.state_dict().values()]).clone()
    n = ndist.broadcast_module_(model, src=0)
    after = torch.cat([t.reshape(-1) for _, t in sorted(model.state_dict().items())])
    gathered = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(gathered, after)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    start, stop = ndist.shard_range(13, world, rank)
    mx = ndist.max_over_ranks(float(rank + 1), torch.device("cpu"))
    sm = ndist.sum_over_ranks(float(stop - start), torch.device("cpu"))
    out[rank] = dict(n=n, same=same, changed=bool((before.sum() != after.sum()).item()), shard=(start, stop),
                     mx=mx, sm=sm, numel=int(after.numel()))
    dist.destroy_process_group()


def test_two_rank_weight_broadcast_and_sharding():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0]["same"] and res[1]["same"]
    assert res[0]["n"] == res[0]["numel"] == res[1]["numel"]
    assert not res[0]["changed"] and res[1]["changed"]     # rank 1 received rank 0's weights
    assert res[0]["shard"] == (0, 7) and res[1]["shard"] == (7, 13)
    assert res[0]["mx"] == res[1]["mx"] == 2.0
    assert res[0]["sm"] == res[1]["sm"] == 13.0


def test_shard_range_partitions_exactly():
    from naf_b200.dist import shard_range

    for total in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
