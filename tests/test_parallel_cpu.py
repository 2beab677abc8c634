"""CPU coverage of the N>1 path: window sharding rule, bucket construction, and the bucketed
gradient allreduce + 1/W scaling with the gloo backend, world_size 2 (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inferbiomechanics_b200 import parallel
from oracle import windows as ow


def test_shard_rule_matches_distributed_sampler():
    """rank r gets r::W of the first floor(N/W)*W windows (torch DistributedSampler(shuffle=False, drop_last=True),
    as constructed at /root/reference/src/cli/train.py:143)."""
    from torch.utils.data.distributed import DistributedSampler
    for n, w in ((10, 4), (17, 2), (3, 4), (64, 8), (100, 3)):
        for r in range(w):
            ds = DistributedSampler(range(n), num_replicas=w, rank=r, shuffle=False, drop_last=True)
            assert list(parallel.shard_indices(n, r, w)) == list(iter(ds)) == ow.sampler_indices(n, w, r)
    assert list(parallel.contiguous_shard(10, 3, 4)) == [9]
    assert sum(len(parallel.contiguous_shard(4096, r, 8)) for r in range(8)) == 4096


def test_make_buckets_covers_arena_in_backward_order():
    bounds = [0, 1000, 3000, 3500, 9000]          # 5 layer groups in a 10 000-element arena
    buckets = parallel.make_buckets(bounds, 10000, 2500)
    # fired from the last group to the first, contiguous, non-overlapping, covering everything
    assert buckets[0][1] == 10000 and buckets[-1][0] == 0
    for (a, b, g), (a2, b2, g2) in zip(buckets, buckets[1:]):
        assert a == b2 and g2 < g
    assert all(b - a >= 2500 or a == 0 for a, b, _ in buckets)
    assert [g for _, _, g in buckets] == sorted([g for _, _, g in buckets], reverse=True)
    one = parallel.make_buckets(bounds, 10000, 10 ** 9)
    assert one == [(0, 10000, 0)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 5000
        bounds = [0, 700, 2100, 4000]
        g = torch.Generator().manual_seed(100 + rank)
        grad = torch.randn(n, generator=g)
        mine = grad.clone()
        bucketer = parallel.GradBucketer(grad, parallel.make_buckets(bounds, n, 1500))
        bucketer.begin_step()
        fired = []
        for group in (3, 2, 1, 0):                 # backward finishes groups last-to-first
            bucketer.group_done(group)
            fired.append(bucketer._fired)
        bucketer.finish()
        # parameters broadcast from rank 0 (DDP constructor semantics)
        param = torch.full((16,), float(rank))
        bucketer.broadcast_(param)
        # reference: mean of all ranks' gradients
        others = [torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]
        want_sum = torch.stack(others).sum(0)
        ok = torch.allclose(grad, want_sum, atol=1e-6) and torch.allclose(grad / world, torch.stack(others).mean(0), atol=1e-6)
        out[rank] = (ok, fired, bucketer.collectives, float(param.sum()), torch.equal(mine, others[rank]))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        ok, fired, collectives, psum, same = out[r]
        assert ok and same
        assert fired == sorted(fired) and fired[-1] == collectives          # buckets fire progressively during "backward"
        assert collectives >= 2 and psum == 0.0                                # params equal rank 0's


def _tail_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 5000
        grad = torch.randn(n, generator=torch.Generator().manual_seed(100 + rank))
        bucketer = parallel.GradBucketer(grad, parallel.make_buckets([0, 700, 2100, 4000], n, 1500), mode="tail")
        for step in range(2):
            bucketer.begin_step()
            for group in (3, 2, 1, 0):
                bucketer.group_done(group)
            before = bucketer.collectives
            bucketer.finish()
            assert bucketer.collectives == before + 1
        want = torch.stack([torch.randn(n, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)]).sum(0)
        ok = torch.allclose(grad, want * world, atol=1e-5)                                      # summed twice: (a+b) then 2(a+b)
        # finish_hook (the trainer's graph capture cuts the step there): called INSTEAD of the collective, which the caller
        # then issues itself with allreduce_all()
        calls = []
        bucketer.finish_hook = lambda: calls.append(bucketer.collectives)
        bucketer.begin_step()
        bucketer.finish()
        ok = ok and calls == [2] and bucketer.collectives == 2 and torch.allclose(grad, want * world, atol=1e-5)
        bucketer.finish_hook = None
        bucketer.allreduce_all()
        ok = ok and bucketer.collectives == 3 and torch.allclose(grad, want * world * world, atol=1e-4)
        bucketer.collectives = 2
        out[rank] = (ok, bucketer.collectives)
    finally:
        dist.destroy_process_group()


def test_tail_allreduce_gloo_world2():
    """mode='tail': nothing fires during backward, finish() issues exactly one allreduce of the whole arena per step."""
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_tail_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        ok, collectives = out[r]
        assert ok and collectives == 2
    with pytest.raises(ValueError):
        parallel.GradBucketer(torch.zeros(4), [(0, 4, 0)], mode="sometimes")


def _agg_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import argparse
        import contextlib
        import io
        from inferbiomechanics_b200.cli.train import print_all_ranks_report
        from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
        ev = RegressionLossEvaluator(None, "dev")
        for b in range(3):                               # three "batches" per rank, result vector = rank*10 + batch everywhere
            ev._results.append(torch.full((40,), float(rank * 10 + b)))
        args = argparse.Namespace(predict_grf_components=[0], predict_cop_components=[], predict_moment_components=[],
                                  predict_wrench_components=[])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            print_all_ranks_report(ev, args, rank, world, torch.device("cpu"))
        out[rank] = (buf.getvalue(), len(ev._results), float(parallel.mean_over_ranks(torch.tensor([float(rank)]))[0]))
    finally:
        dist.destroy_process_group()


def test_rank0_metric_aggregate_gloo_world2():
    """rank 0 prints one aggregate over all ranks (mean of per-rank means of batch results) next to the
    per-rank reports of the reference; the per-rank lists are left untouched."""
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_agg_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    text0, n0, m0 = out[0]
    text1, n1, m1 = out[1]
    assert n0 == n1 == 3 and m0 == m1 == 0.5
    assert text1 == "" and "[all 2 ranks] dev set:" in text0
    assert "Force Avg Err: 6.0 N / kg" in text0          # mean over ranks {0,1} and batches {0,1,2} of rank*10 + batch
