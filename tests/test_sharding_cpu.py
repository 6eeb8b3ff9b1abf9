"""N>1 host logic on CPU: streams are partitioned over ranks with no data-path collective; the only exchange is the
max-over-ranks of the device time (bench.py).  Runs a world-size-2 gloo group on 127.0.0.1."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_enhancement_mi_b200 import workload


def test_shard_streams_is_a_partition():
    for total in (1, 7, 1024, 16384, 16385):
        for world in (1, 2, 3, 4, 8):
            owned = []
            for r in range(world):
                start, count = workload.shard_streams(total, world, r)
                owned += list(range(start, start + count))
            assert owned == list(range(total))
            sizes = [workload.shard_streams(total, world, r)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = workload.shard_streams(total, world, rank)
    # every rank "processes" its own block of streams: stream s produces the value s (stands for the enhanced hop)
    mine = torch.arange(start, start + count, dtype=torch.float64)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), mine.numpy())
    # the bench's only collective: max over ranks of the measured time, then the whole-job aggregate on rank 0
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(count)], dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    if rank == 0:
        np.save(os.path.join(out_dir, "agg.npy"), np.array([t.item(), n.item()]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_cover_all_streams_without_exchange(tmp_path):
    world, total = 2, 1025
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"rank{r}.npy") for r in range(world)])
    assert np.array_equal(got, np.arange(total, dtype=np.float64))
    t_max, n_sum = np.load(tmp_path / "agg.npy")
    assert t_max == 2.0 and n_sum == total
    # whole-job throughput as bench.py reports it: all streams of all ranks / max time
    assert workload.AUDIO_SEC_PER_STEP * n_sum / t_max == 0.1 * total / 2.0
