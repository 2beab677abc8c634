"""The train / analyze commands end to end on a B200 (synthetic AddBiomechanics-shaped windows, BASELINE configs[0]
shape: FeedForward, T=50, stride 5, batch 32): flags as in the reference, checkpoint written in the reference format
and picked up by analyze, loss decreasing, window-store file round trip bit-exact."""
import argparse
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(argv):
    from inferbiomechanics_b200.main import main
    return main(argv)


def test_train_then_analyze_feedforward(tmp_path, capsys):
    ck = str(tmp_path / "ck")
    common = ["--no-wandb", "--synthetic-windows", "4096", "--checkpoint-dir", ck, "--history-len", "50", "--stride", "5",
              "--hidden-dims", "512", "512", "--activation", "sigmoid"]
    cmd = _run(["train", *common, "--model-type", "feedforward", "--epochs", "2", "--batch-size", "32", "--learning-rate", "1e-3"])
    assert cmd is not None
    files = sorted(os.listdir(os.path.join(ck, "feedforward")))
    # reference naming (train.py:271): one checkpoint at the last batch of each epoch (synthetic stores round the window
    # count up to whole trials, so the batch count is read back rather than assumed)
    assert len(files) == 2 and files[0].startswith("epoch_0_batch_") and files[1].startswith("epoch_1_batch_")
    assert files[0].split("_batch_")[1] == files[1].split("_batch_")[1]
    sd = torch.load(os.path.join(ck, "feedforward", files[-1]), map_location="cpu")
    assert set(sd) == {"epoch", "model_state_dict", "optimizer_state_dict"} and sd["epoch"] == 1
    assert set(sd["model_state_dict"]) == {"net.0.weight", "net.0.bias", "net.2.weight", "net.2.bias", "net.4.weight", "net.4.bias"}
    out = capsys.readouterr().out
    assert "Evaluating Dev Set Before Epoch 1" in out and "Force Avg Err" in out
    an = _run(["analyze", *common, "--model-type", "feedforward", "--predict-grf-components", "0", "1", "2", "3", "4", "5"])
    rep = an.last_reports
    assert set(rep) == {"dev", "train"} and all(r is not None and r["loss"] > 0 for r in rep.values())
    # resuming: a third epoch starts from the epoch-1 checkpoint (abstract_command.load_latest_checkpoint)
    _run(["train", *common, "--model-type", "feedforward", "--epochs", "3", "--batch-size", "32", "--max-batches", "4"])
    assert "epoch_2_batch_3.pt" in os.listdir(os.path.join(ck, "feedforward"))
    assert "Loaded checkpoint from epoch 1" in capsys.readouterr().out


def test_train_groundlink_and_diffusion_smoke(tmp_path):
    ck = str(tmp_path / "ck")
    common = ["--no-wandb", "--synthetic-windows", "512", "--checkpoint-dir", ck, "--history-len", "50", "--stride", "1",
              "--epochs", "1", "--max-batches", "3"]
    assert _run(["train", *common, "--model-type", "groundlink", "--batch-size", "16"]) is not None
    assert _run(["train", *common, "--model-type", "diffusion", "--batch-size", "64"]) is not None
    an = _run(["analyze", "--no-wandb", "--synthetic-windows", "512", "--checkpoint-dir", ck, "--history-len", "50", "--stride", "1",
               "--model-type", "diffusion", "--batch-size", "64", "--sampling-steps", "20", "--predict-grf-components", "0", "1", "2"])
    assert an.last_reports["dev"]["loss"] > 0


def test_window_store_file_roundtrip(tmp_path):
    from inferbiomechanics_b200.data.window_store import WindowStore
    a = WindowStore.synthetic(3000, 50, 5, 147, "all_frames", seed=5, device="cuda")
    p = str(tmp_path / "train.ibmstore")
    a.save(p)
    b = WindowStore.load(p, 50, 5, "all_frames", device="cuda")
    assert len(a) == len(b) and torch.equal(a.win_row0, b.win_row0)
    idx = torch.arange(0, len(a), 7, device="cuda")
    assert torch.equal(a.pack_f32(idx), b.pack_f32(idx)) and torch.equal(a.labels(idx), b.labels(idx))
    with open(p, "r+b") as f:
        f.write(b"XX")
    with pytest.raises(ValueError):
        WindowStore.load(p, 50, 5, "all_frames", device="cuda")
