"""CPU restatement of the reference regression loss.  TEST INFRASTRUCTURE ONLY.

Follows ``/root/reference/src/loss/RegressionLossEvaluator.py``:

* ``squared_diff_mean_vector``  ← ``get_squared_diff_mean_vector`` (73-83)
* ``mask_by_threes``            ← ``get_mask_by_threes`` (85-108)
* ``mean_norm_error``           ← ``get_mean_norm_error`` (119-141)  (LAST FRAME ONLY, line 136)
* ``com_acc_error``             ← ``get_com_acc_error`` (143-158)
* ``regression_loss``           ← ``__call__`` step 1 (160-221) and step 2.2 (230-263)
* ``regression_loss_grad``      ← closed-form d loss / d outputs

Pinned by the reference's 24 unit tests (tests/test_oracle_loss.py restates each case) and by
tests/golden/loss_call_*.npz generated from the imported reference (oracle/gen_golden.py).

Written with plain torch CPU tensor arithmetic in closed form (one expression per quantity);
fp32 in, fp32 out, reductions in fp64 where noted so the oracle is the tighter side of every
tolerance.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

COP = "groundContactCenterOfPressureInRootFrame"
FORCE = "groundContactForceInRootFrame"
TORQUE = "groundContactTorqueInRootFrame"
WRENCH = "groundContactWrenchesInRootFrame"
LOSS_KEYS = (COP, FORCE, TORQUE, WRENCH)
COP_FORCE_THRESHOLD = 10.0  # RegressionLossEvaluator.py:207


def _check3(o: torch.Tensor, l: torch.Tensor) -> None:
    if o.shape != l.shape:
        raise ValueError("Output and label tensors must have the same shape")
    if o.dim() != 3:
        raise ValueError("Output and label tensors must be 3-dimensional")
    if o.numel() == 0:
        raise ValueError("Output and label tensors must not be empty")


def squared_diff_mean_vector(o: torch.Tensor, l: torch.Tensor) -> torch.Tensor:
    _check3(o, l)
    d = o - l
    return (d * d).sum(dim=(0, 1)) / float(o.shape[0] * o.shape[1])


def mask_by_threes(x: torch.Tensor, threshold: float = 0.0) -> torch.Tensor:
    if x.dim() != 3:
        raise ValueError("Mask tensor must be 3-dimensional")
    if x.numel() == 0:
        raise ValueError("Mask tensor must not be empty")
    if x.shape[-1] % 3 != 0:
        raise ValueError("Mask tensor must have a final dimension divisible by 3")
    B, F, C = x.shape
    g = x.detach().reshape(B, F, C // 3, 3)
    n = torch.sqrt((g * g).sum(-1))
    m = (n > threshold).to(x.dtype)
    return m.repeat_interleave(3, dim=-1).reshape(B, F, C)


def mean_norm_error(o: torch.Tensor, l: torch.Tensor, vec_size: int = 3) -> torch.Tensor:
    _check3(o, l)
    if o.shape[-1] % vec_size != 0:
        raise ValueError("Tensors must have a final dimension divisible by vec_size=" + str(vec_size))
    d = (o - l)[:, -1, :].reshape(o.shape[0], -1, vec_size)
    return torch.sqrt((d * d).sum(-1)).mean()


def com_acc_error(o: torch.Tensor, l: torch.Tensor) -> torch.Tensor:
    _check3(o, l)
    if o.shape[-1] != 6:
        raise ValueError("Output and label tensors must have a 6 dimensional final dimension")
    return mean_norm_error(o[:, :, :3] + o[:, :, 3:], l[:, :, :3] + l[:, :, 3:], 3)


def regression_loss(outputs: Dict[str, torch.Tensor], labels: Dict[str, torch.Tensor],
                    grf: Sequence[int], cop: Sequence[int], moment: Sequence[int],
                    wrench: Sequence[int]) -> Dict[str, torch.Tensor]:
    """Returns dict(loss, force, cop, moment, wrench vectors, and the 7 report scalars).

    Index lists may repeat or be empty (advanced indexing then sum; analyze.py:44-47)."""
    force_v = squared_diff_mean_vector(outputs[FORCE], labels[FORCE])
    moment_v = squared_diff_mean_vector(outputs[TORQUE], labels[TORQUE])
    wrench_v = squared_diff_mean_vector(outputs[WRENCH], labels[WRENCH])
    m = mask_by_threes(labels[FORCE], COP_FORCE_THRESHOLD)
    cop_v = squared_diff_mean_vector(outputs[COP] * m, labels[COP] * m)
    idx = lambda v, s: v[torch.as_tensor(list(s), dtype=torch.long)].sum()
    loss = idx(force_v, grf) + idx(cop_v, cop) + idx(moment_v, moment) + idx(wrench_v, wrench)
    with torch.no_grad():
        wm1 = mean_norm_error(outputs[WRENCH][:, :, :3], labels[WRENCH][:, :, :3], 3)
        wm2 = mean_norm_error(outputs[WRENCH][:, :, 6:9], labels[WRENCH][:, :, 6:9], 3)
        rep = dict(
            force_report=mean_norm_error(outputs[FORCE], labels[FORCE]),
            moment_report=mean_norm_error(outputs[TORQUE], labels[TORQUE]),
            cop_report=mean_norm_error(outputs[COP] * m, labels[COP] * m),
            wrench_moment_report=(wm1 + wm2) / 2.0,
            wrench_report=mean_norm_error(outputs[WRENCH], labels[WRENCH], 6),
            com_acc_report=com_acc_error(outputs[FORCE], labels[FORCE]),
        )
    return dict(loss=loss, force=force_v, cop=cop_v, moment=moment_v, wrench=wrench_v, **rep)


def regression_loss_grad(outputs: Dict[str, torch.Tensor], labels: Dict[str, torch.Tensor],
                         grf: Sequence[int], cop: Sequence[int], moment: Sequence[int],
                         wrench: Sequence[int]) -> Dict[str, torch.Tensor]:
    """Closed-form d loss/d outputs: 2 w_c m^2 (o-l)/N with w_c the multiplicity of c in the
    index list."""
    B, F, _ = outputs[FORCE].shape
    N = float(B * F)

    def weights(sel, C):
        w = torch.zeros(C, dtype=torch.float32)
        for c in sel:
            w[c] += 1.0
        return w

    m = mask_by_threes(labels[FORCE], COP_FORCE_THRESHOLD)
    return {
        FORCE: 2.0 * weights(grf, 6) * (outputs[FORCE] - labels[FORCE]) / N,
        TORQUE: 2.0 * weights(moment, 6) * (outputs[TORQUE] - labels[TORQUE]) / N,
        WRENCH: 2.0 * weights(wrench, 12) * (outputs[WRENCH] - labels[WRENCH]) / N,
        COP: 2.0 * weights(cop, 6) * m * (outputs[COP] - labels[COP]) / N,
    }
