"""N > 1 plumbing on CPU: two gloo ranks take disjoint contiguous utterance shards, synthesise disjoint
utterances, and agree on the max-over-ranks time (the only communication the path has)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import seedvc_b200  # noqa: F401
from seedvc_b200 import synth
from seedvc_b200.sharding import barrier, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 32, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_utt, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = shard_range(n_utt, world, rank)
        mu, prompt, style, z = synth.synth_batch(count, 12, 4, 8, 16, first_id=start)
        # every rank learns every shard and a checksum of what each rank synthesised
        spans = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(spans, torch.tensor([start, count]))
        sums = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(sums, z.double().sum().reshape(1))
        barrier("cpu")
        t = max_over_ranks(10.0 + 5.0 * rank, "cpu")          # the slowest rank defines the job time
        q.put((rank, [tuple(int(v) for v in s) for s in spans], [float(s) for s in sums], t,
               float(z.double().sum())))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_shard_and_time():
    world, n_utt = 2, 7
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_utt, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spans = res[0][1]
    assert spans == res[1][1] == [(0, 4), (4, 3)]                 # contiguous, disjoint, complete
    assert res[0][3] == res[1][3] == 15.0                          # max over ranks, same on both
    assert res[0][2] == res[1][2]                                  # both saw both checksums
    assert abs(res[0][2][0] - res[0][4]) < 1e-9 and abs(res[0][2][1] - res[1][4]) < 1e-9
    assert res[0][4] != res[1][4]                                  # different utterances per rank
    # the shards are exactly the slices of the single-process batch
    _, _, _, z_all = synth.synth_batch(n_utt, 12, 4, 8, 16, first_id=0)
    assert abs(float(z_all[:4].double().sum()) - res[0][4]) < 1e-9
    assert abs(float(z_all[4:].double().sum()) - res[1][4]) < 1e-9
