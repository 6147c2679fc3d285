"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (the GPU box runs the same code over NCCL).
The path has no data-path collective (sets are independent): what is tested is that the shards
partition the global batch, that the timing reduction is a max and that counters come back in rank order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointdsc_b200.shard import gather_counters, max_over_ranks, output_checksum, shard_bounds
from pointdsc_b200.synth import make_batch


def test_shard_bounds_partition_every_batch():
    for total in (0, 1, 7, 256, 1024, 1031):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))           # contiguous, no overlap
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(total, rank, world)
        batch = make_batch(range(lo, hi), 16, "3dmatch", 0.5)                    # this rank's sets, by GLOBAL set index
        # stand-in for the engine's outputs (no GPU here): the ground-truth pose and labels of the rank's own sets
        cs = output_checksum(batch["gt_trans"], batch["gt_labels"])
        cs["rank"] = float(rank)
        cs["lo"], cs["hi"] = float(lo), float(hi)
        allc = gather_counters(cs)
        slow = max_over_ranks(1.0 + rank, "cpu")                                 # rank r "took" 1 + r seconds
        if rank == 0:
            out.put((allc, slow))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_shards_timing_and_counters():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world = 5, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    allc, slow = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert slow == 2.0                                                            # max over ranks, not mean or rank 0
    assert [c["rank"] for c in allc] == [0.0, 1.0]                                # rank order
    assert (allc[0]["lo"], allc[0]["hi"], allc[1]["lo"], allc[1]["hi"]) == (0.0, 3.0, 3.0, 5.0)
    whole = make_batch(range(total), 16, "3dmatch", 0.5)
    ref = output_checksum(whole["gt_trans"], whole["gt_labels"])
    for k in ("sets", "trans_sum", "trans_abs", "inliers"):                       # the shards are exactly the batch
        assert abs(sum(c[k] for c in allc) - ref[k]) < 1e-9 * max(1.0, abs(ref[k]))
