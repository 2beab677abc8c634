"""``analyze`` command: evaluates the latest checkpoint on the dev and train splits and prints the evaluator's
report, flags as in ``/root/reference/src/cli/analyze.py:20-242``.

The reference walks a ``DataLoader(batch_size=1)`` in one process (analyze.py:112,139-180); windows are
independent, so here each rank takes a CONTIGUOUS range of windows in large equal batches (no
collective during the pass), keeps the per-batch fp32[40] results on the device and merges them once at the end —
equal batch sizes keep the reference's mean-of-batch-means aggregate (RegressionLossEvaluator.py:383-387) exact.
Per-window plots (matplotlib) and the names CSV are outside the hot path.  ``--model-type diffusion`` runs the
1000-step reverse sampler per batch and scores the sampled trajectories with the same evaluator.
"""
from __future__ import annotations

import argparse
import logging
import os

import torch
import torch.distributed as dist

from .. import parallel
from ..diffusion import GaussianDiffusion
from ..keys import LOSS_QUANTITIES
from ..loss.RegressionLossEvaluator import RegressionLossEvaluator
from ..trainer import Trainer
from . import _data
from .abstract_command import AbstractCommand
from .train import _ResultLog, input_dict, label_dict


class AnalyzeCommand(AbstractCommand):
    def __init__(self):
        super().__init__()

    def register_subcommand(self, subparsers: argparse._SubParsersAction):
        sp = subparsers.add_parser('analyze', help='Evaluate the performance of a model on dataset.')
        sp.add_argument('--dataset-home', type=str, default='../data', help='The path to the AddBiomechanics dataset.')
        sp.add_argument('--model-type', type=str, default='feedforward', help='The model to train.')
        sp.add_argument('--no-wandb', action='store_true', default=False, help='Log this analysis to Weights and Biases.')
        sp.add_argument('--output-data-format', type=str, default='all_frames', choices=['all_frames', 'last_frame'],
                        help='Output for all frames in a window or only the last frame.')
        sp.add_argument('--checkpoint-dir', type=str, default='../checkpoints', help='Where the checkpoints of `train` live.')
        sp.add_argument('--geometry-folder', type=str, default=None, help='Path to the Geometry folder with bone mesh data.')
        sp.add_argument('--history-len', type=int, default=50,
                        help='The number of timesteps of context to show when constructing the inputs.')
        sp.add_argument('--stride', type=int, default=5,
                        help='The number of timesteps of context to show when constructing the inputs.')
        sp.add_argument('--hidden-dims', type=int, nargs='+', default=[512, 512], help='Hidden dims across different layers.')
        sp.add_argument('--activation', type=str, default='sigmoid', help='Which activation func?')
        sp.add_argument('--batchnorm', action='store_true', help='The checkpoint was trained with --batchnorm (else read off its keys).')
        sp.add_argument('--dropout', action='store_true', help='The checkpoint was trained with --dropout (else read off its keys).')
        sp.add_argument('--device', type=str, default='cuda', help='Accepted for compatibility; this path runs on the GPU only.')
        sp.add_argument('--short', type=bool, default=False, help='Use very short datasets to test without loading a bunch of data.')
        sp.add_argument('--data-loading-workers', type=int, default=3, help='Accepted for compatibility.')
        sp.add_argument('--predict-grf-components', type=int, nargs='+', default=[1], help='Which grf components to train.')
        sp.add_argument('--predict-cop-components', type=int, nargs='+', default=[], help='Which grf components to train.')
        sp.add_argument('--predict-moment-components', type=int, nargs='+', default=[], help='Which grf components to train.')
        sp.add_argument('--predict-wrench-components', type=int, nargs='+', default=[], help='Which grf components to train.')
        # additions of the B200 path
        sp.add_argument('--synthetic-windows', type=int, default=0, help='Analyze synthetic windows instead of *.ibmstore files.')
        sp.add_argument('--batch-size', type=int, default=4096, help='Windows per launch (the reference analyses one at a time).')
        sp.add_argument('--sampling-steps', type=int, default=1000, help='Reverse-diffusion steps for --model-type diffusion.')

    def run(self, args: argparse.Namespace):
        if 'command' in args and args.command != 'analyze':
            return False
        model_type: str = args.model_type
        checkpoint_dir: str = os.path.join(os.path.abspath(args.checkpoint_dir), model_type)
        root_history_len = 10
        rank, world, local = parallel.init_from_env("nccl")
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        # --batchnorm / --dropout shift the nn.Sequential positions of the checkpoint's keys (the reference's analyze has no
        # such flags and cannot load those checkpoints): taken from the flags, else read off the latest checkpoint's keys
        batchnorm, dropout = self.feedforward_layout(checkpoint_dir) if model_type == 'feedforward' else (False, False)
        model = self.get_model(_data.NUM_DOFS, 2, model_type, history_len=args.history_len, hidden_dims=args.hidden_dims,
                               activation=args.activation, stride=args.stride, batchnorm=args.batchnorm or batchnorm,
                               dropout=args.dropout or dropout, dropout_prob=0.0,
                               root_history_len=root_history_len, output_data_format=args.output_data_format,
                               device=str(device)).to(device)
        self.load_latest_checkpoint(model, checkpoint_dir=checkpoint_dir)
        model.eval()
        native = model_type == 'feedforward'
        trainer = Trainer(model, args=args) if native else None
        diffusion = GaussianDiffusion(device=device) if model_type == 'diffusion' else None
        if not args.no_wandb:
            import wandb
            wandb.init(project="addbiomechanics-baseline", config=dict(args.__dict__))

        reports = {}
        for split, seed in (('dev', 4321), ('train', 1234)):
            logging.info(f'## Loading {split} dataset:')
            store = _data.open_split(args, split, model_type, device, seed=seed)
            n = len(store)
            shard = parallel.contiguous_shard(n, rank, world)
            idx_all = torch.arange(shard.start, shard.stop, device=device)
            log, ev = _ResultLog(split), RegressionLossEvaluator(None, split, device=device)
            with torch.no_grad():
                for i in range(0, idx_all.numel(), args.batch_size):
                    idx = idx_all[i:i + args.batch_size]
                    if native:
                        log.add(trainer.eval_step(store, idx))
                        continue
                    if model_type == 'diffusion':
                        # condition = the packed kinematics in the engine's concat buffer; x_T ~ Philox; CUDA-graph replays
                        store.pack_rows(idx, model.engine().xc(idx.numel(), False), col0=30)
                        x0 = diffusion.sample(model, idx.numel(), steps=args.sampling_steps, seed=1234 + rank)
                        if store.Fo == 1:                     # --output-data-format last_frame: labels hold the last frame only
                            x0 = x0[:, -1:]
                        cuts = [(0, 6), (6, 12), (12, 18), (18, 30)]
                        inputs = {}
                        outputs = {k: x0[:, :, a:b] for k, (a, b) in zip(LOSS_QUANTITIES, cuts)}
                    else:
                        inputs = input_dict(store, idx, model_type, args.stride, root_history_len)
                        outputs = model(inputs)
                    ev(inputs, outputs, label_dict(store, idx), [], [], args)
            src = log.ev if native else ev
            # one merge of the per-rank batch results at the very end (≈ 40 floats per batch); no collective before
            if world > 1:
                mine = torch.stack(src._results) if src._results else torch.zeros(0, 40, device=device)
                counts = [torch.zeros(1, dtype=torch.long, device=device) for _ in range(world)]
                dist.all_gather(counts, torch.tensor([mine.shape[0]], device=device))
                mx = int(max(c.item() for c in counts))
                padded = torch.zeros(mx, 40, device=device)
                padded[:mine.shape[0]] = mine
                gathered = [torch.zeros_like(padded) for _ in range(world)]
                dist.all_gather(gathered, padded)
                merged = [g[:int(c.item())] for g, c in zip(gathered, counts)]
                src._results = [r for g in merged for r in g]
                src._wm_results = list(src._results)
            if rank == 0:
                print(f'{split} set evaluation ({n} windows):')
                reports[split] = src.aggregate()
                src.print_report(args, log_to_wandb=not args.no_wandb)
        if dist.is_initialized():
            dist.destroy_process_group()
        self.last_reports = reports
        return True
