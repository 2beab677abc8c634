"""Data-parallel equivalence on real NCCL ranks (needs >= 2 GPUs; skipped otherwise — run with `gpurun --gpus 2`):
the reference wraps the model in DistributedDataParallel and shards windows with DistributedSampler(shuffle=False,
drop_last=True) (/root/reference/src/cli/train.py:143-150,175,281).  Two properties of that contract are checked for the
native Trainer: (1) after every step all ranks hold IDENTICAL parameters although they started from different ones
(broadcast at construction + allreduce of the flat gradient arena), (2) W ranks with B windows each take the same step as one
rank with the union batch of W*B windows (mean-allreduce of per-rank mean losses == gradient of the union mean)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "workers", "dp_equivalence_worker.py")


def _spawn(tmp_path, kind, opt, steps, B, world=2, extra_env=None):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", WORKER, str(tmp_path), kind, opt, str(steps), str(B)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-3000:]
    return [torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(world)]


def _need_two():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")


@pytest.mark.parametrize("opt", ["sgd", "rmsprop"])
def test_two_ranks_equal_one_rank_on_the_union_batch(tmp_path, opt):
    _need_two()
    from inferbiomechanics_b200.data.window_store import WindowStore
    from inferbiomechanics_b200.models.FeedForwardRegressionBaseline import FeedForwardBaseline
    from inferbiomechanics_b200.trainer import Trainer
    steps, B = 3, 64
    ranks = _spawn(tmp_path, "feedforward", opt, steps, B)
    assert ranks[0]["world"] == 2 and ranks[0]["collectives"] >= steps
    for n, p in ranks[0]["params"].items():
        assert torch.equal(p, ranks[1]["params"][n]), f"{n} differs between ranks"
    # the 1-rank run on the union batches, from rank 0's initial weights (what the broadcast distributes)
    torch.manual_seed(100)
    model = FeedForwardBaseline(23, 2, 50, "all_frames", "sigmoid", 5, 10, hidden_dims=[64, 64]).cuda()
    store = WindowStore.synthetic(2048, 50, 5, 147, "all_frames", seed=3, device="cuda")
    tr = Trainer(model, opt_type=opt, lr=1e-2 if opt == "sgd" else 1e-3, seed=5)
    idx = store.shard(0, 1)
    for s in range(steps):
        loss = tr.train_step(store, idx[2 * s * B:2 * (s + 1) * B])[0].item()
        both = 0.5 * (ranks[0]["losses"][s] + ranks[1]["losses"][s])
        assert abs(loss - both) <= 1e-4 * abs(both), (s, loss, both)
    # parameters: fp32 sums in a different order (two partial sums + allreduce vs one split-K reduction).  SGD is linear in
    # the gradient: 1e-5 of max|p|.  RMSprop divides by sqrt(v) ~ |g|: near-zero gradients amplify the last bits, so its
    # bar is on the update, 2e-2 of the lr-sized steps taken
    for n, p in model.named_parameters():
        ref, got = p.detach().cpu(), ranks[0]["params"][n]
        tol = 1e-5 * ref.abs().max().item() if opt == "sgd" else 2e-2 * steps * 1e-3
        assert (ref - got).abs().max().item() <= tol, (n, (ref - got).abs().max().item(), tol)


def test_denoiser_ranks_stay_in_lockstep(tmp_path):
    """Denoiser training draws per-rank timesteps and noise (seed + rank), so there is no 1-rank twin; the DDP invariant is
    that parameters are bit-identical across ranks after every allreduced step, with finite, rank-specific losses."""
    _need_two()
    ranks = _spawn(tmp_path, "diffusion", "rmsprop", 3, 32)
    for n, p in ranks[0]["params"].items():
        assert torch.equal(p, ranks[1]["params"][n]), f"{n} differs between ranks"
        assert torch.isfinite(p).all()
    assert ranks[0]["losses"] != ranks[1]["losses"]


@pytest.mark.parametrize("kind", ["feedforward", "groundlink"])
def test_data_parallel_graph_replay_equals_eager(tmp_path, kind):
    """The small models' data-parallel step is replayed from two CUDA graphs cut at the gradient allreduce (trainer.py
    _capture_step; the reference's default is batch 64 under DDP, train.py:52,175): same parameters and losses as the eager
    launches, ranks in lockstep.  fp32 atomics (bias-gradient column sums) make the two runs differ in the last bits only."""
    _need_two()
    steps, B = 7, 32                                   # steps 1-2 eager, step 3 captures, 4-7 replay
    d_e, d_g = tmp_path / "eager", tmp_path / "graph"
    d_e.mkdir(), d_g.mkdir()
    eager = _spawn(str(d_e), kind, "rmsprop", steps, B, extra_env={"IBM_TRAIN_GRAPHS": "0"})
    graph = _spawn(str(d_g), kind, "rmsprop", steps, B)
    assert eager[0]["graphs"] == 0 and graph[0]["graphs"] == 1
    for n, p in graph[0]["params"].items():
        assert torch.equal(p, graph[1]["params"][n]), f"{n} differs between ranks"
        ref = eager[0]["params"][n]
        assert (p - ref).abs().max().item() <= 2e-3 * steps * 1e-3 + 1e-6 * ref.abs().max().item(), n     # << one lr-sized update
    for a, b in zip(eager[0]["losses"], graph[0]["losses"]):
        assert abs(a - b) <= 1e-4 * abs(a)
