"""Command-line entry point with the reference's dispatcher shape (``/root/reference/src/main.py:16-59``):

    python -m inferbiomechanics_b200.main train   --model-type feedforward --synthetic-windows 65536 --no-wandb …
    python -m inferbiomechanics_b200.main analyze --model-type feedforward --synthetic-windows 65536 --no-wandb …
    torchrun --nproc-per-node 8 -m inferbiomechanics_b200.main train …          # data parallel, one rank per GPU

Only the two commands on the hot path exist here; the GUI viewers, dataset splitting/pickling, plotting and
CSV export commands of the reference need nimblephysics / matplotlib and are out of scope.
"""
import argparse
import logging

from .cli.analyze import AnalyzeCommand
from .cli.train import TrainCommand


def main(argv=None):
    commands = [TrainCommand(), AnalyzeCommand()]
    parser = argparse.ArgumentParser(description='InferBiomechanics Command Line Interface (B200 hot path)')
    subparsers = parser.add_subparsers(dest="command")
    for command in commands:
        command.register_subcommand(subparsers)
    args = parser.parse_args(argv)
    for command in commands:
        if command.run(args):
            return command
    parser.print_help()
    return None


if __name__ == '__main__':
    logger = logging.getLogger()
    logger.addHandler(logging.StreamHandler())
    logger.setLevel(logging.INFO)
    main()
