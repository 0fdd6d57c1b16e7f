"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: disjoint batch slices, max-over-ranks timing."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gta_b200.shard import batch_slice, job_throughput, max_over_ranks


def test_batch_slice_partitions():
    for gb in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [batch_slice(gb, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == gb
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = batch_slice(65, world, rank)
        # each rank "processes" its slice; the job time is the slowest rank's
        ms = 10.0 + 5.0 * rank
        mx = max_over_ranks(ms)
        # the slices are disjoint and cover the batch: sum of sizes == 65
        n = torch.tensor([b - a], dtype=torch.int64)
        dist.all_reduce(n)
        dist.barrier()
        q.put((rank, a, b, mx, int(n.item())))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1:3] for r in res] == [(0, 33), (33, 65)]
    assert all(r[3] == 15.0 and r[4] == 65 for r in res)
    assert job_throughput(1000, 2, 15.0) == 2000 / 0.015


def test_packed_reps_batch_slice_is_a_view():
    """Batch chunking of the host-buffer pipeline (gta_b200/host.py): the packed rep tables slice along the batch."""
    import torch
    from gta_b200.ops import PackedReps
    r = PackedReps(se3_q=torch.arange(4 * 2 * 16.).reshape(4, 2, 16), se3_k=torch.zeros(4, 3, 16),
                   so2_q=torch.ones(4, 10, 6, 2), n_q_views=2, n_k_views=3)
    s = r.batch_slice(1, 3)
    assert s.se3_q.shape == (2, 2, 16) and s.se3_q.data_ptr() == r.se3_q[1:3].data_ptr()
    assert s.so3_q is None and s.so2_q.shape == (2, 10, 6, 2) and (s.n_q_views, s.n_k_views) == (2, 3)
    B, chunks = 64, 16
    bounds = [(B * i // chunks, B * (i + 1) // chunks) for i in range(chunks)]
    assert bounds[0] == (0, 4) and bounds[-1] == (60, 64) and all(b[0] == a[1] for a, b in zip(bounds, bounds[1:]))
