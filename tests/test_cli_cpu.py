"""CLI surface (no GPU): the train / analyze sub-commands accept the reference's command lines and defaults
(/root/reference/src/cli/train.py:24-69, analyze.py:23-47), the model factory and checkpoint loader keep the reference's
signatures (abstract_command.py:44-120), and the pre-packed store header round-trips."""
import argparse
import inspect
import os

import pytest
import torch


def _parser():
    from inferbiomechanics_b200.cli.analyze import AnalyzeCommand
    from inferbiomechanics_b200.cli.train import TrainCommand
    p = argparse.ArgumentParser()
    sub = p.add_subparsers(dest="command")
    for c in (TrainCommand(), AnalyzeCommand()):
        c.register_subcommand(sub)
    return p


def test_train_flags_and_defaults_match_reference():
    a = _parser().parse_args(["train"])
    # defaults of the reference (train.py:26-69)
    assert (a.dataset_home, a.model_type, a.output_data_format, a.checkpoint_dir) == ("../data", "feedforward", "all_frames", "../checkpoints")
    assert (a.history_len, a.stride, a.learning_rate, a.dropout, a.dropout_prob) == (50, 5, 1e-4, False, 0.5)
    assert (a.hidden_dims, a.batchnorm, a.activation, a.epochs, a.opt_type, a.batch_size) == ([512, 512], False, "sigmoid", 10, "rmsprop", 64)
    assert a.predict_grf_components == list(range(6)) and a.predict_wrench_components == list(range(12))
    assert a.no_wandb is False and a.short is False and a.compute_report is False and a.data_loading_workers == 1
    # a reference-style command line parses unchanged
    b = _parser().parse_args("train --model-type groundlink --no-wandb --history-len 100 --stride 2 --opt-type adam --batch-size 128 "
                             "--hidden-dims 256 256 128 --activation relu --predict-grf-components 1 4 --trial-filter walk run".split())
    assert b.model_type == "groundlink" and b.hidden_dims == [256, 256, 128] and b.predict_grf_components == [1, 4]


def test_analyze_flags_and_defaults_match_reference():
    a = _parser().parse_args(["analyze"])
    assert (a.model_type, a.history_len, a.stride, a.hidden_dims, a.activation) == ("feedforward", 50, 5, [512, 512], "sigmoid")
    assert a.predict_grf_components == [1] and a.predict_cop_components == [] and a.predict_wrench_components == []


def test_factory_and_checkpoint_loader_signatures(tmp_path, capsys):
    from inferbiomechanics_b200.cli.abstract_command import AbstractCommand
    sig = inspect.signature(AbstractCommand.get_model)
    assert list(sig.parameters)[1:] == ["num_dofs", "num_contact_bodies", "model_type", "history_len", "stride", "hidden_dims",
                                        "activation", "batchnorm", "dropout", "dropout_prob", "root_history_len",
                                        "output_data_format", "device"]
    cmd = AbstractCommand()
    with pytest.raises(ValueError):
        cmd.get_model(23, 2, "analytical")
    # latest-checkpoint rule (epoch, batch) and the DDP `module.` prefix (train.py:276) on a plain torch module
    model = torch.nn.Linear(3, 2)
    d = tmp_path / "ck"
    assert cmd.load_latest_checkpoint(model, checkpoint_dir=str(d)) == (-1, 0)
    os.makedirs(d)
    assert cmd.load_latest_checkpoint(model, checkpoint_dir=str(d)) == (-1, 0)
    for e, b, val in ((0, 999, 1.0), (2, 5, 3.0), (1, 2000, 2.0)):
        sd = {"module." + k: torch.full_like(v, val) for k, v in model.state_dict().items()}
        torch.save({"epoch": e, "model_state_dict": sd, "optimizer_state_dict": {}}, d / f"epoch_{e}_batch_{b}.pt")
    assert cmd.load_latest_checkpoint(model, checkpoint_dir=str(d)) == (2, 5)
    assert float(model.weight.detach()[0, 0]) == 3.0
    assert "Loaded checkpoint from epoch 2, batch 5" in capsys.readouterr().out


def test_feedforward_state_dict_positions_with_batchnorm_and_dropout():
    """nn.Sequential positions shift with --dropout / --batchnorm exactly as in the reference (SURVEY §9.3,
    FeedForwardRegressionBaseline.py:68-75): per layer [Dropout, BatchNorm1d(h0), Linear, act], BatchNorm on the layer INPUT
    (the raw 1470-vector included).  Pure constructor check: no GPU needed."""
    from inferbiomechanics_b200.models.FeedForwardRegressionBaseline import FeedForwardBaseline
    m = FeedForwardBaseline(23, 2, 50, "all_frames", "sigmoid", 5, 10, hidden_dims=[512, 512], batchnorm=True, dropout=True,
                            dropout_prob=0.1)
    sd = m.state_dict()
    assert tuple(sd["net.1.running_mean"].shape) == (1470,) and "net.1.num_batches_tracked" in sd
    assert tuple(sd["net.2.weight"].shape) == (512, 1470)
    assert tuple(sd["net.5.weight"].shape) == (512,) and tuple(sd["net.6.weight"].shape) == (512, 512)
    assert tuple(sd["net.9.running_var"].shape) == (512,) and tuple(sd["net.10.weight"].shape) == (300, 512)
    plain = FeedForwardBaseline(23, 2, 50, "all_frames", "sigmoid", 5, 10).state_dict()
    assert sorted(plain) == ["net.0.bias", "net.0.weight", "net.2.bias", "net.2.weight", "net.4.bias", "net.4.weight"]
    only_bn = FeedForwardBaseline(23, 2, 50, "last_frame", "relu", 5, 10, hidden_dims=[64], batchnorm=True).state_dict()
    assert tuple(only_bn["net.1.weight"].shape) == (64, 1470) and tuple(only_bn["net.4.weight"].shape) == (30, 64)


def test_analyze_reads_batchnorm_dropout_layout_off_the_checkpoint(tmp_path):
    """analyze has no --batchnorm/--dropout in the reference and therefore cannot load such checkpoints; here the layout is
    read off the checkpoint's nn.Sequential positions (with and without the DDP `module.` prefix)."""
    from inferbiomechanics_b200.cli.abstract_command import AbstractCommand
    from inferbiomechanics_b200.models.FeedForwardRegressionBaseline import FeedForwardBaseline
    assert AbstractCommand.feedforward_layout(str(tmp_path / "none")) == (False, False)
    for i, (bn, do) in enumerate([(False, False), (True, False), (False, True), (True, True)]):
        d = tmp_path / f"ck{i}"
        os.makedirs(d)
        m = FeedForwardBaseline(23, 2, 50, "last_frame", "relu", 5, 10, hidden_dims=[16], batchnorm=bn, dropout=do, dropout_prob=0.1)
        sd = {("module." if i % 2 else "") + k: v for k, v in m.state_dict().items()}
        torch.save({"epoch": 0, "model_state_dict": sd, "optimizer_state_dict": {}}, d / "epoch_0_batch_7.pt")
        assert AbstractCommand.feedforward_layout(str(d)) == (bn, do)


def test_evaluator_wrench_moment_history_is_kept_as_host_floats():
    """The reference never resets wrench_moment_reported_metrics (RegressionLossEvaluator.py:412-426): the history survives
    print_report(reset=True), but as python floats rather than one retained device tensor per step."""
    from inferbiomechanics_b200.loss.RegressionLossEvaluator import RegressionLossEvaluator
    ev = RegressionLossEvaluator(None, "train")
    for v in (1.0, 2.0):
        r = torch.zeros(40)
        r[34] = v
        ev._results.append(r)
        ev._wm_results.append(r)
    assert ev.wrench_moment_reported_metrics == [1.0, 2.0]
    ev._reset_lists(keep_wrench_moment=True)
    assert ev._wm_results == [] and ev._results == [] and ev.wrench_moment_reported_metrics == [1.0, 2.0]
    r = torch.zeros(40)
    r[34] = 3.0
    ev._wm_results.append(r)
    assert ev.wrench_moment_reported_metrics == [1.0, 2.0, 3.0]
    ev._reset_lists()
    assert ev.wrench_moment_reported_metrics == []
