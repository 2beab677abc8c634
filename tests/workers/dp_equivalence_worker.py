"""Worker of tests/test_gpu_multirank.py: one process per GPU (torchrun), NCCL.  Every rank trains a few native steps on its
DistributedSampler shard (r::W, train.py:143-150) and writes its post-step parameters and losses to <out>/rank<r>.pt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main(out_dir: str, kind: str, opt: str, steps: int, B: int, timed: int = 0):
    from inferbiomechanics_b200 import parallel
    from inferbiomechanics_b200.data.window_store import WindowStore
    from inferbiomechanics_b200.trainer import Trainer
    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(100 + rank)                      # DIFFERENT initial weights per rank: the broadcast must equalise them
    if kind == "feedforward":
        from inferbiomechanics_b200.models.FeedForwardRegressionBaseline import FeedForwardBaseline
        model = FeedForwardBaseline(23, 2, 50, "all_frames", "sigmoid", 5, 10, hidden_dims=[64, 64]).to(dev)
        store = WindowStore.synthetic(2048, 50, 5, 147, "all_frames", seed=3, device=dev)
    elif kind == "groundlink":
        from inferbiomechanics_b200.models.Groundlink import Groundlink
        model = Groundlink(23, 12, 10, "all_frames", fc_dropout=0.0).to(dev).train()
        store = WindowStore.synthetic(2048, 20, 1, 177, "all_frames", seed=3, device=dev, trial_len=120)
    else:
        from inferbiomechanics_b200.models.DiffusionDenoiser import DiffusionDenoiser
        model = DiffusionDenoiser(frames=10, d_model=128, num_heads=2, dim_feedforward=256, num_layers=2).to(dev)
        store = WindowStore.synthetic(2048, 10, 1, 177, "all_frames", seed=3, device=dev)
    tr = Trainer(model, opt_type=opt, lr=1e-2 if opt == "sgd" else 1e-3, seed=5)
    idx = store.shard(rank, world)
    losses = []
    for s in range(steps):
        losses.append(tr.train_step(store, idx[s * B:(s + 1) * B])[0].item())
    torch.cuda.synchronize()
    ms = None
    if timed > 0:                                      # steady-state step time (tools/dp_graph_probe.py): same batch, no host reads
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.distributed.barrier()
        torch.cuda.synchronize()
        e0.record()
        for s in range(timed):
            tr.train_step(store, idx[:B])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / timed
    torch.save({"ms_per_step": ms, "graphs": len(tr._graphs), "mode": tr.bucketer.mode, "params": {n: p.detach().cpu() for n, p in model.named_parameters()}, "losses": losses,
                "collectives": tr.bucketer.collectives, "world": world}, os.path.join(out_dir, f"rank{rank}.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]) if len(sys.argv) > 6 else 0)
