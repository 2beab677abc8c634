"""GPU parity for the HBM-bound kernels, through the C ABI, against the CPU oracle (oracle/) and
the golden vectors generated from the imported reference.

Bars: bit-exact for window indexing, packing, masking; fp32 kernels within rtol 2e-6/atol 1e-7 of
the oracle (summation order only); bf16-I/O kernels within one bf16 rounding of an fp32
computation on the same bf16 inputs.
"""
import math

import numpy as np
import pytest
import torch

from oracle import ddpm as oddpm
from oracle import loss as ol
from oracle import windows as ow
from oracle.gen_golden import SELECTIONS, seeded_out_labels

pytestmark = pytest.mark.gpu
Q = (ol.COP, ol.FORCE, ol.TORQUE, ol.WRENCH)


def weights30(sel):
    grf, cop, moment, wrench = [list(x) for x in sel]
    w = [0.0] * 30
    for c in cop: w[c] += 1
    for c in grf: w[6 + c] += 1
    for c in moment: w[12 + c] += 1
    for c in wrench: w[18 + c] += 1
    return w


def unpack_result(r):
    r = r.cpu().double()
    return dict(loss=r[0], cop=r[1:7], force=r[7:13], moment=r[13:19], wrench=r[19:31], force_report=r[31],
                moment_report=r[32], cop_report=r[33], wrench_moment_report=r[34], wrench_report=r[35], com_acc_report=r[36])


# ---------------------------------------- loss ----------------------------------------------------
@pytest.mark.parametrize("case", ["b4f10", "b3f1", "b7f50"])
@pytest.mark.parametrize("sel", list(SELECTIONS))
def test_loss_matches_reference_golden(golden, case, sel):
    from inferbiomechanics_b200 import ops
    g = golden("loss_call.npz")
    B, F, seed = (int(v) for v in g[f"{case}/meta"])
    o, l = seeded_out_labels(B, F, seed)
    outs = [o[k].cuda() for k in Q]
    labs = [l[k].cuda() for k in Q]
    w = weights30(SELECTIONS[sel])
    res = unpack_result(ops.regression_loss_fwd(outs, labs, w))
    for k in ("loss", "force", "cop", "moment", "wrench", "force_report", "moment_report", "cop_report", "wrench_report",
              "wrench_moment_report", "com_acc_report"):
        np.testing.assert_allclose(res[k].numpy(), g[f"{case}/{sel}/{k}"], rtol=2e-6, atol=1e-7, err_msg=k)
    grads = [torch.full_like(t, float("nan")) for t in outs]
    ops.regression_loss_bwd(outs, labs, w, grads)
    for k, gt in zip(Q, grads):
        np.testing.assert_allclose(gt.cpu().numpy(), g[f"{case}/{sel}/grad/{k}"], rtol=2e-6, atol=1e-9, err_msg=k)


def test_loss_strided_layouts_and_bf16_grads():
    """FeedForward's quantity-blocked output (FeedForward…py:116-121), rows30 outputs with ld 32,
    upstream gradient scalar, bf16 gradient rows (what the head's dgrad GEMM consumes)."""
    from inferbiomechanics_b200 import ops
    B, F = 37, 10
    o, l = seeded_out_labels(B, F, 4242)
    w = weights30(SELECTIONS["repeat"])
    ref = ol.regression_loss(o, l, *[list(x) for x in SELECTIONS["repeat"]])
    gref = ol.regression_loss_grad(o, l, *[list(x) for x in SELECTIONS["repeat"]])
    # (a) FeedForward layout: x[:, 0:6F | 6F:12F | 12F:18F | 18F:30F]
    x = torch.cat([o[ol.COP].reshape(B, -1), o[ol.FORCE].reshape(B, -1), o[ol.TORQUE].reshape(B, -1),
                   o[ol.WRENCH].reshape(B, -1)], dim=1).cuda()
    views = [x[:, 0:6 * F].view(B, F, 6), x[:, 6 * F:12 * F].view(B, F, 6), x[:, 12 * F:18 * F].view(B, F, 6),
             x[:, 18 * F:30 * F].view(B, F, 12)]
    labs = [l[k].cuda() for k in Q]
    r = unpack_result(ops.regression_loss_fwd(views, labs, w))
    np.testing.assert_allclose(r["loss"].item(), ref["loss"].item(), rtol=2e-6)
    # (b) rows30, ld = 32
    rows = torch.zeros(B * F, 32, device="cuda")
    rows[:, :30] = torch.cat([o[k] for k in Q], dim=-1).reshape(B * F, 30).cuda()
    r3 = rows.view(B, F, 32)
    rv = [r3[:, :, 0:6], r3[:, :, 6:12], r3[:, :, 12:18], r3[:, :, 18:30]]
    lrows = torch.cat([l[k] for k in Q], dim=-1).cuda()
    lv = [lrows[:, :, 0:6], lrows[:, :, 6:12], lrows[:, :, 12:18], lrows[:, :, 18:30]]
    r = unpack_result(ops.regression_loss_fwd(rv, lv, w))
    np.testing.assert_allclose(r["loss"].item(), ref["loss"].item(), rtol=2e-6)
    np.testing.assert_allclose(r["wrench"].numpy(), ref["wrench"].numpy(), rtol=2e-6)
    for k in ("force_report", "cop_report", "wrench_moment_report", "com_acc_report"):
        np.testing.assert_allclose(r[k].item(), ref[k].item(), rtol=2e-6, err_msg=k)
    # bf16 grad rows with an upstream scalar
    g16 = torch.zeros(B * F, 32, dtype=torch.bfloat16, device="cuda").view(B, F, 32)
    gv = [g16[:, :, 0:6], g16[:, :, 6:12], g16[:, :, 12:18], g16[:, :, 18:30]]
    up = torch.tensor(0.5, device="cuda")
    ops.regression_loss_bwd(rv, lv, w, gv, upstream=up)
    want = torch.cat([gref[k] for k in Q], dim=-1) * 0.5
    got = g16[:, :, :30].float().cpu()
    assert torch.allclose(got, want.to(torch.bfloat16).float(), rtol=1e-2, atol=1e-12)
    assert torch.all(g16[:, :, 30:] == 0)


def test_loss_mask_is_bit_exact_at_threshold():
    """mask_by_threes uses strict > (…Evaluator.py:102): ‖(6,8,0)‖ = 10 exactly is NOT contact."""
    from inferbiomechanics_b200 import ops
    B, F = 2, 1
    lab_force = torch.tensor([[[6.0, 8.0, 0.0, 6.0, 8.0, 0.1]], [[0.0, 0.0, 10.0000009537, 0.0, 0.0, 0.0]]])
    o = {k: torch.ones(B, F, 12 if k == ol.WRENCH else 6) for k in Q}
    l = {k: torch.zeros(B, F, 12 if k == ol.WRENCH else 6) for k in Q}
    l[ol.FORCE] = lab_force
    res = unpack_result(ops.regression_loss_fwd([o[k].cuda() for k in Q], [l[k].cuda() for k in Q], [1.0] * 30))
    ref = ol.regression_loss(o, l, range(6), range(6), range(6), range(12))
    np.testing.assert_array_equal(res["cop"].float().numpy(), ref["cop"].numpy())
    assert res["cop"][0].item() == 0.5 and res["cop"][3].item() == 0.5     # one of two windows masked in each group


def test_loss_errors_like_reference():
    from inferbiomechanics_b200 import ops
    e = [torch.zeros(0, 1, 12 if i == 3 else 6, device="cuda") for i in range(4)]
    with pytest.raises(ValueError):       # "must not be empty" (…Evaluator.py:78-79)
        ops.regression_loss_fwd(e, e, [1.0] * 30)


def test_loss_large_property():
    """Full-size property: loss(o, l) == 0 when o == l, and loss is additive in the component weights."""
    from inferbiomechanics_b200 import ops
    B, F = 4096, 50
    g = torch.Generator(device="cuda").manual_seed(5)
    rows = torch.randn(B, F, 30, device="cuda", generator=g) * 5
    lab = torch.randn(B, F, 30, device="cuda", generator=g) * 5
    v = lambda r: [r[:, :, 0:6], r[:, :, 6:12], r[:, :, 12:18], r[:, :, 18:30]]
    assert ops.regression_loss_fwd(v(rows), v(rows), [1.0] * 30)[0].item() == 0.0
    w1 = [1.0 if i % 2 == 0 else 0.0 for i in range(30)]
    w2 = [0.0 if i % 2 == 0 else 1.0 for i in range(30)]
    a = ops.regression_loss_fwd(v(rows), v(lab), w1)[0].item()
    b = ops.regression_loss_fwd(v(rows), v(lab), w2)[0].item()
    c = ops.regression_loss_fwd(v(rows), v(lab), [1.0] * 30)[0].item()
    assert abs((a + b) - c) <= 1e-5 * abs(c)
    ref = ol.regression_loss({k: t.cpu() for k, t in zip(Q, v(rows))}, {k: t.cpu() for k, t in zip(Q, v(lab))},
                             range(6), range(6), range(6), range(12))
    assert abs(c - ref["loss"].item()) <= 2e-6 * abs(c)


# ---------------------------------------- DDPM ----------------------------------------------------
def test_q_sample_and_posterior_match_oracle():
    from inferbiomechanics_b200 import ops
    sched = oddpm.make_schedule()
    dev = {k: v.cuda() for k, v in sched.items()}
    for B, F in ((5, 10), (3, 1), (64, 50)):
        g = torch.Generator().manual_seed(B * 100 + F)
        x0 = torch.randn(B, F, 30, generator=g) * 3
        eps = torch.randn(B, F, 30, generator=g)
        t = torch.randint(0, 1000, (B,), generator=g)
        want = oddpm.q_sample(sched, x0, t, eps)
        xt = torch.empty(B, F, 30, device="cuda")
        xb = torch.zeros(B * F, 208, dtype=torch.bfloat16, device="cuda")
        ops.q_sample(x0.cuda(), eps.cuda(), t.int().cuda(), dev["sqrt_abar"], dev["sqrt_one_minus_abar"], xt_f32=xt,
                     xt_bf16=xb, bf16_ld=208)
        torch.testing.assert_close(xt.cpu(), want, rtol=2e-6, atol=1e-6)
        assert torch.equal(xb[:, :30].cpu(), xt.cpu().reshape(B * F, 30).to(torch.bfloat16))     # bit-exact RNE scatter
        assert torch.all(xb[:, 30:] == 0)
        # posterior step for t = 0 (no noise), 1 (clipped variance) and a middle step
        x0h = torch.randn(B * F, 32, generator=g)
        z = torch.randn(B * F, 30, generator=g)
        for tt in (0, 1, 500, 999):
            want = oddpm.posterior_step(sched, x0h[:, :30], xt.cpu().reshape(-1, 30), tt, z)
            tdev = torch.tensor([tt], dtype=torch.int32, device="cuda")
            tnext = torch.zeros(1, dtype=torch.int32, device="cuda")
            out = torch.empty(B * F, 30, device="cuda")
            ops.posterior_step(x0h.cuda(), 32, xt.view(-1, 30), z.cuda(), tdev, dev["coef_x0"], dev["coef_xt"], dev["sigma"],
                               B * F, x_prev=out, t_next=tnext)
            torch.testing.assert_close(out.cpu(), want, rtol=2e-6, atol=1e-6)
            assert tnext.item() == tt - 1


def test_on_device_noise_is_standard_normal_and_reproducible():
    from inferbiomechanics_b200 import ops
    sched = {k: v.cuda() for k, v in oddpm.make_schedule().items()}
    B, F = 2048, 50
    x0 = torch.zeros(B, F, 30, device="cuda")
    t = torch.full((B,), 999, dtype=torch.int32, device="cuda")
    e1 = torch.empty(B, F, 30, device="cuda")
    e2 = torch.empty_like(e1)
    xt = torch.empty_like(e1)
    ops.q_sample(x0, None, t, sched["sqrt_abar"], sched["sqrt_one_minus_abar"], xt_f32=xt, seed=7, offset=1, eps_out=e1)
    ops.q_sample(x0, None, t, sched["sqrt_abar"], sched["sqrt_one_minus_abar"], xt_f32=xt, seed=7, offset=1, eps_out=e2)
    assert torch.equal(e1, e2)
    ops.q_sample(x0, None, t, sched["sqrt_abar"], sched["sqrt_one_minus_abar"], xt_f32=xt, seed=7, offset=2, eps_out=e2)
    assert not torch.equal(e1, e2)
    assert abs(e1.mean().item()) < 5e-3 and abs(e1.std().item() - 1.0) < 5e-3
    assert abs((e1 ** 4).mean().item() - 3.0) < 0.05
    torch.testing.assert_close(xt, sched["sqrt_one_minus_abar"][999] * e2, rtol=1e-6, atol=1e-7)


def test_timestep_embedding_and_time_pos():
    from inferbiomechanics_b200 import ops
    from oracle.models import sinusoidal_embedding
    B, d, F = 33, 512, 10
    t = torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(1))
    out = torch.empty(B, d, dtype=torch.bfloat16, device="cuda")
    ops.timestep_embed(t.int().cuda(), out, d)
    want = sinusoidal_embedding(t, d)
    assert (out.float().cpu() - want).abs().max().item() <= 1.0 / 256 + 2e-3     # bf16 rounding of values in [-1, 1]
    g = torch.Generator().manual_seed(2)
    h = torch.randn(B * F, d, generator=g).to(torch.bfloat16)
    temb = torch.randn(B, d, generator=g).to(torch.bfloat16)
    pos = torch.randn(F, d, generator=g)
    hd = h.cuda()
    ops.add_time_pos(hd, temb.cuda(), pos.cuda(), B * F, F, d)
    want = (h.float().view(B, F, d) + temb.float().unsqueeze(1) + pos.unsqueeze(0)).view(B * F, d)
    assert torch.equal(hd.cpu(), want.to(torch.bfloat16))
    # table mode (reverse sampling): every window adds row t_row[0] of a per-timestep table
    hd2 = h.cuda()
    ops.add_time_pos(hd2, temb.cuda(), pos.cuda(), B * F, F, d, t_row=torch.tensor([17], dtype=torch.int32, device="cuda"))
    want2 = (h.float().view(B, F, d) + temb[17].float().view(1, 1, d) + pos.unsqueeze(0)).view(B * F, d)
    assert torch.equal(hd2.cpu(), want2.to(torch.bfloat16))
    dh = torch.randn(B * F, d, generator=g).to(torch.bfloat16)
    dtemb = torch.empty(B, d, dtype=torch.bfloat16, device="cuda")
    dpos = torch.zeros(F, d, device="cuda")
    ops.add_time_pos_bwd(dh.cuda(), dtemb, dpos, B * F, F, d)
    torch.testing.assert_close(dpos.cpu(), dh.float().view(B, F, d).sum(0), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(dtemb.float().cpu(), dh.float().view(B, F, d).sum(1).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)
    # the one-pass kernel over other shapes (ragged window groups, 50 / 60 frames, narrow and wide rows) with the bias-gradient
    # output; accumulates into dpos / dbias
    for B2, F2, d2 in ((37, 50, 512), (5, 60, 128), (130, 7, 1024), (16, 50, 72)):
        dh2 = torch.randn(B2 * F2, d2, generator=g).to(torch.bfloat16)
        dt2 = torch.empty(B2, d2, dtype=torch.bfloat16, device="cuda")
        dp2 = torch.ones(F2, d2, device="cuda")
        db2 = torch.full((d2,), 2.0, device="cuda")
        ops.add_time_pos_bwd(dh2.cuda(), dt2, dp2, B2 * F2, F2, d2, dbias=db2)
        v = dh2.float().view(B2, F2, d2)
        torch.testing.assert_close(dp2.cpu(), 1.0 + v.sum(0), rtol=1e-5, atol=2e-4)
        torch.testing.assert_close(db2.cpu(), 2.0 + v.sum((0, 1)), rtol=1e-5, atol=1e-3)
        torch.testing.assert_close(dt2.float().cpu(), v.sum(1).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-2)


# ---------------------------------------- LayerNorm ------------------------------------------------
@pytest.mark.parametrize("M,d,ld", [(1000, 512, 512), (300, 108, 112), (77, 128, 128), (50, 1024, 1024), (129, 64, 64),
                                    (5000, 256, 256), (3001, 768, 768), (70001, 512, 512), (9, 512, 512),
                                    (40999, 108, 112), (1001, 115, 120), (33, 40, 40), (5, 8, 8), (4096, 64, 72)])
def test_layernorm_fwd_bwd(M, d, ld):
    from inferbiomechanics_b200 import ops
    g = torch.Generator().manual_seed(M + d)
    s = torch.zeros(M, ld, dtype=torch.bfloat16)
    s[:, :d] = (torch.randn(M, d, generator=g) * 2 + 0.5).to(torch.bfloat16)
    s[:, d:] = 9.0                                          # garbage in the pad columns must be ignored
    gamma = 1 + 0.1 * torch.randn(d, generator=g)
    beta = 0.1 * torch.randn(d, generator=g)
    dy = torch.zeros(M, ld, dtype=torch.bfloat16)
    dy[:, :d] = torch.randn(M, d, generator=g).to(torch.bfloat16)
    x = s[:, :d].float().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y_ref = torch.nn.functional.layer_norm(x, (d,), gm, bt, 1e-5)
    y_ref.backward(dy[:, :d].float())
    y = torch.empty(M, ld, dtype=torch.bfloat16, device="cuda")
    mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    ops.layernorm_fwd(s.cuda(), y, gamma.cuda(), beta.cuda(), M, d, mean=mean, rstd=rstd)
    assert (y[:, :d].float().cpu() - y_ref.detach()).abs().max().item() <= 1.0 / 64
    assert torch.all(y[:, d:] == 0)
    ds = torch.empty(M, ld, dtype=torch.bfloat16, device="cuda")
    dg = torch.zeros(d, device="cuda"); db = torch.zeros(d, device="cuda"); dc = torch.zeros(d, device="cuda")
    ops.layernorm_bwd(dy.cuda(), s.cuda(), gamma.cuda(), mean, rstd, M, d, ds, dg, db, dc)
    assert (ds[:, :d].float().cpu() - x.grad).abs().max().item() <= 2e-2 * x.grad.abs().max().item() + 1e-3
    torch.testing.assert_close(dg.cpu(), gm.grad, rtol=1e-3, atol=2e-3 * math.sqrt(M))
    torch.testing.assert_close(db.cpu(), bt.grad, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(dc.cpu(), x.grad.sum(0), rtol=1e-3, atol=2e-3 * math.sqrt(M))    # fp32 sum of the unrounded ds


# ---------------------------------------- attention ------------------------------------------------
def _attn_ref(q, k, v, scale):
    s = (q @ k.transpose(-2, -1)) * scale
    return torch.softmax(s, dim=-1) @ v


@pytest.mark.parametrize("n_win,T,H,hd", [(3, 50, 8, 64), (2, 64, 2, 64), (2, 200, 3, 48), (5, 10, 4, 32), (1, 256, 1, 64), (4, 1, 2, 64),
                                          (900, 50, 8, 64), (2000, 23, 3, 48),
                                          # long windows (64 < T <= 256): persistent one-CTA-per-(window, head) kernel
                                          (3, 65, 2, 32), (2, 129, 4, 64), (700, 200, 3, 48), (2, 250, 1, 48), (5, 72, 3, 48)])
def test_attention_forward(n_win, T, H, hd):
    from inferbiomechanics_b200 import ops
    d = H * hd
    g = torch.Generator().manual_seed(T * 7 + hd)
    qkv = torch.randn(n_win * T, 3 * d, generator=g).to(torch.bfloat16)
    scale = 1.0 / math.sqrt(hd)
    o = torch.zeros(n_win * T, d, dtype=torch.bfloat16, device="cuda")
    ops.attention_fwd_fused(qkv.cuda(), d, o, n_win, T, H, hd, scale)
    x = qkv.double().view(n_win, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(x[0], x[1], x[2], scale).permute(0, 2, 1, 3).reshape(n_win * T, d)
    assert (o.double().cpu() - ref).abs().max().item() <= 2.5e-2


def test_attention_forward_tcgen05_variant():
    """The opt-in tcgen05 forward (IBM_ATTN_FWD=tc, read once per process) against the same fp64 reference, in a subprocess."""
    import os
    import subprocess
    import sys
    code = (
        "import torch, math, sys; sys.path.insert(0, '.');"
        "from inferbiomechanics_b200 import ops;"
        "n_win, T, H, hd = 701, 50, 8, 64; d = H * hd;"
        "g = torch.Generator().manual_seed(5);"
        "qkv = torch.randn(n_win * T, 3 * d, generator=g).to(torch.bfloat16);"
        "o = torch.zeros(n_win * T, d, dtype=torch.bfloat16, device='cuda');"
        "ops.attention_fwd_fused(qkv.cuda(), d, o, n_win, T, H, hd, 1.0 / math.sqrt(hd));"
        "x = qkv.double().view(n_win, T, 3, H, hd).permute(2, 0, 3, 1, 4);"
        "s = (x[0] @ x[1].transpose(-2, -1)) / math.sqrt(hd);"
        "ref = (torch.softmax(s, -1) @ x[2]).permute(0, 2, 1, 3).reshape(n_win * T, d);"
        "print((o.double().cpu() - ref).abs().max().item())")
    env = dict(os.environ, IBM_ATTN_FWD="tc")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=os.path.dirname(os.path.dirname(__file__)))
    assert r.returncode == 0, r.stderr[-800:]
    assert float(r.stdout.split()[-1]) <= 2.5e-2


@pytest.mark.parametrize("n_win,T,H,hd", [(3, 50, 8, 64), (2, 64, 2, 64), (5, 10, 4, 32), (2, 33, 3, 48), (700, 50, 8, 64),
                                          (1500, 17, 4, 32)])
def test_attention_backward(n_win, T, H, hd):
    from inferbiomechanics_b200 import ops
    d = H * hd
    g = torch.Generator().manual_seed(T * 11 + hd)
    qkv = (torch.randn(n_win * T, 3 * d, generator=g) * 0.8).to(torch.bfloat16)
    do = torch.randn(n_win * T, d, generator=g).to(torch.bfloat16)
    scale = 1.0 / math.sqrt(hd)
    x = qkv.double().view(n_win, T, 3, H, hd).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    out = _attn_ref(x[0], x[1], x[2], scale).permute(0, 2, 1, 3).reshape(n_win * T, d)
    out.backward(do.double())
    want = x.grad.permute(1, 3, 0, 2, 4).reshape(n_win * T, 3 * d)
    dqkv = torch.full((n_win * T, 3 * d), float("nan"), dtype=torch.bfloat16, device="cuda")
    dbias = torch.full((3 * d,), 0.5, device="cuda")
    ops.attention_bwd(qkv.cuda(), d, do.cuda(), dqkv, n_win, T, H, hd, scale, dbias=dbias)
    err = (dqkv.double().cpu() - want).abs().max().item()
    assert err <= 3e-2 * want.abs().max().item() + 1e-3, err
    # fused in_proj_bias gradient: 0.5 + fp32 column sums of the kernel's unrounded dqkv (fp32 atomics: order differs);
    # vs the fp64 reference the per-element error is the bf16-operand error above, accumulating like sqrt(rows)
    torch.testing.assert_close(dbias.double().cpu() - 0.5, want.sum(0), rtol=2e-2,
                               atol=2e-2 * want.abs().max().item() * math.sqrt(n_win * T))
    # and it stays within bf16 rounding of the column sums of what was written
    torch.testing.assert_close(dbias.double().cpu() - 0.5, dqkv.double().sum(0).cpu(), rtol=2e-2,
                               atol=4e-3 * want.abs().max().item() * math.sqrt(n_win * T))
    # without the accumulator the kernel must still run (NULL pointer path)
    dqkv2 = torch.empty_like(dqkv)
    ops.attention_bwd(qkv.cuda(), d, do.cuda(), dqkv2, n_win, T, H, hd, scale)
    assert torch.equal(dqkv, dqkv2)


@pytest.mark.parametrize("n_win,T", [(2, 200), (3, 50), (600, 200), (2, 77), (1, 256)])
def test_simple_attention_head(n_win, T):
    """SimpleAttention (TransformerBaseline.py:51-70): unscaled scores, value dim 3 (padded to 8).  T > 64 runs the
    persistent long-window kernel (one CTA per window, 112-wide q/k, 8-wide values), T <= 64 the one-shot kernel."""
    from inferbiomechanics_b200 import ops
    d = 112
    g = torch.Generator().manual_seed(9)
    q = torch.zeros(n_win * T, d, dtype=torch.bfloat16); k = torch.zeros_like(q)
    q[:, :108] = (torch.randn(n_win * T, 108, generator=g) * 0.3).to(torch.bfloat16)
    k[:, :108] = (torch.randn(n_win * T, 108, generator=g) * 0.3).to(torch.bfloat16)
    v = torch.zeros(n_win * T, 8, dtype=torch.bfloat16)
    v[:, :3] = torch.randn(n_win * T, 3, generator=g).to(torch.bfloat16)
    o = torch.zeros(n_win * T, 8, dtype=torch.bfloat16, device="cuda")
    ops.attention_fwd(q.cuda(), k.cuda(), v.cuda(), o, n_win, T, 1, 112, 8, 1.0)
    ref = _attn_ref(q.double().view(n_win, T, d), k.double().view(n_win, T, d), v.double().view(n_win, T, 8), 1.0)
    assert (o.double().cpu().view(n_win, T, 8) - ref).abs().max().item() <= 2e-2


# ---------------------------------------- window batcher -------------------------------------------
def _store(subjects, C_keys, nb):
    """Flatten synthetic subjects into the HBM frame-store arrays."""
    frames, raw, missing, base, mass, cidx = [], [], [], [], [], []
    off = 0
    for s in subjects:
        for tr in s["trials"]:
            L = len(tr["missing"])
            frames.append(np.concatenate([tr[k] for k in C_keys], axis=1).astype(np.float32))
            raw.append(np.concatenate([tr["groundContactCenterOfPressureInRootFrame"], tr["groundContactForceInRootFrame"],
                                       tr["groundContactTorqueInRootFrame"], tr["groundContactWrenchesInRootFrame"]],
                                      axis=1).astype(np.float32))
            missing.append(tr["missing"].astype(np.uint8))
            base.append(off)
            off += L
    return np.concatenate(frames), np.concatenate(raw), np.concatenate(missing), np.array(base, dtype=np.int64)


@pytest.mark.parametrize("T,s,hist", [(50, 5, 15), (20, 1, 30), (50, 7, 30)])
def test_window_index_and_packing_bit_exact(T, s, hist):
    from inferbiomechanics_b200 import ops
    D, nb = 23, 2
    subjects = ow.make_synthetic_subjects(11 + T, 6, T, num_dofs=D, hist_cols=hist, max_len=160)
    want_windows = ow.enumerate_windows(subjects, T, s)
    frames, raw, missing, base = _store(subjects, ow.INPUT_ORDER, nb)
    # candidates: every (trial, ws) with ws in range(max(L - T - 1, 0))   (Dataset.py:134)
    ct, cs, tri, info = [], [], 0, []
    for si, sub in enumerate(subjects):
        for ti, tr in enumerate(sub["trials"]):
            L = len(tr["missing"])
            n = max(L - T - 1, 0)
            ct += [tri] * n
            cs += list(range(n))
            info += [(si, ti)] * n
            tri += 1
    ct_t = torch.tensor(ct, dtype=torch.int32, device="cuda"); cs_t = torch.tensor(cs, dtype=torch.int32, device="cuda")
    valid = torch.empty(len(ct), dtype=torch.uint8, device="cuda")
    ops.window_valid_mask(torch.from_numpy(missing).cuda(), torch.from_numpy(base).cuda(), ct_t, cs_t, T, s, valid)
    keep = valid.cpu().numpy().astype(bool)
    got_windows = [(info[i][0], info[i][1], cs[i]) for i in range(len(ct)) if keep[i]]
    assert got_windows == want_windows                       # bit-exact index, same order
    # pack a ragged batch of windows (sampler rule for rank 1 of 3, last partial batch)
    idx = ow.sampler_indices(len(want_windows), 3, 1)
    batch = ow.batches(idx, 7)[-1]
    F = T // s
    C = frames.shape[1]
    ld = ops.round_up(C, 4)
    fr = torch.zeros(frames.shape[0], ld); fr[:, :C] = torch.from_numpy(frames)
    tmap = {}
    tri = 0
    for si, sub in enumerate(subjects):
        for ti in range(len(sub["trials"])):
            tmap[(si, ti)] = tri; tri += 1
    row0 = torch.tensor([base[tmap[(want_windows[i][0], want_windows[i][1])]] + want_windows[i][2] for i in batch], dtype=torch.int64)
    out32 = torch.empty(len(batch), F, C, device="cuda")
    ldk = ops.round_up(F * C, 8)
    ff16 = torch.zeros(len(batch), ldk, dtype=torch.bfloat16, device="cuda")
    ops.pack_windows(fr.cuda(), C, row0.cuda(), F, s, out_f32=out32, out_bf16=ff16, frame_stride=C, win_extra=ldk - F * C)
    rows16 = torch.zeros(len(batch) * F, 216, dtype=torch.bfloat16, device="cuda")
    ops.pack_windows(fr.cuda(), C, row0.cuda(), F, s, out_bf16=rows16, frame_stride=216, win_extra=0, col0=30)
    lab = torch.empty(len(batch) * F, 30, device="cuda")
    cidx = torch.tensor([subjects[want_windows[i][0]]["contact_indices"] for i in batch], dtype=torch.int32)
    mass = torch.tensor([subjects[want_windows[i][0]]["mass"] for i in batch], dtype=torch.float32)
    rw = torch.from_numpy(raw)
    ops.pack_labels(rw.cuda(), nb, row0.cuda(), cidx.cuda(), mass.cuda(), F, s, False, lab)
    lab1 = torch.empty(len(batch), 30, device="cuda")
    ops.pack_labels(rw.cuda(), nb, row0.cuda(), cidx.cuda(), mass.cuda(), F, s, True, lab1)
    for bi, wi in enumerate(batch):
        inputs, labels = ow.get_window(subjects, want_windows[wi], T, s, "all_frames", nb)
        x = ow.pack_inputs(inputs, flatten=False)
        assert np.array_equal(out32[bi].cpu().numpy(), x)                                   # bit-exact fp32 copy
        assert torch.equal(ff16[bi, :F * C].cpu(), torch.from_numpy(x.reshape(-1)).to(torch.bfloat16))
        assert torch.equal(rows16[bi * F:(bi + 1) * F, 30:30 + C].cpu(), torch.from_numpy(x).to(torch.bfloat16))
        assert np.array_equal(lab[bi * F:(bi + 1) * F].cpu().numpy(), ow.pack_labels30(labels))   # incl. fp32 /mass
        _, last = ow.get_window(subjects, want_windows[wi], T, s, "last_frame", nb)
        assert np.array_equal(lab1[bi:bi + 1].cpu().numpy(), ow.pack_labels30(last))
    assert torch.all(ff16[:, F * C:] == 0) and torch.all(rows16[:, :30] == 0)
    # empty batch is a no-op, not an error
    ops.pack_windows(fr.cuda(), C, row0[:0].cuda(), F, s, out_f32=out32)


# ---------------------------------------- optimizer / misc -----------------------------------------
@pytest.mark.parametrize("kind", ["rmsprop", "adam", "sgd", "adagrad", "adadelta", "adamax"])
def test_optimizer_matches_torch(kind):
    from inferbiomechanics_b200 import ops
    n = 10007
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    cls = {"rmsprop": torch.optim.RMSprop, "adam": torch.optim.Adam, "sgd": torch.optim.SGD, "adagrad": torch.optim.Adagrad,
           "adadelta": torch.optim.Adadelta, "adamax": torch.optim.Adamax}[kind]
    opt = cls([ref_p], lr=1e-2)
    pad = ops.round_up(n, 8)
    p = torch.zeros(pad, device="cuda"); p[:n] = p0.cuda()
    s0 = torch.zeros(pad, device="cuda"); s1 = torch.zeros(pad, device="cuda")
    pb = torch.zeros(pad, dtype=torch.bfloat16, device="cuda")
    for step in range(1, 6):
        grad = torch.randn(n, generator=g)
        ref_p.grad = grad.clone()
        opt.step()
        gd = torch.zeros(pad, device="cuda"); gd[:n] = grad.cuda() * 4.0
        ops.optimizer_step(kind, p, gd, s0, s1, pb, 1e-2, 0.25, step)        # grad_scale undoes the x4 (allreduce-sum / W)
    torch.testing.assert_close(p[:n].cpu(), ref_p.detach(), rtol=2e-5, atol=2e-6)
    assert torch.equal(pb[:n].cpu(), p[:n].cpu().to(torch.bfloat16))


def test_colsum_cast_act():
    from inferbiomechanics_b200 import ops
    g = torch.Generator().manual_seed(4)
    X = torch.randn(1234, 1536, generator=g).to(torch.bfloat16)
    out = torch.ones(1536, device="cuda")
    ops.colsum(X.cuda(), 1234, 1536, out)
    torch.testing.assert_close(out.cpu(), 1 + X.float().sum(0), rtol=1e-4, atol=1e-3)
    Xr = torch.randn(100, 304, generator=g).to(torch.bfloat16)
    out = torch.zeros(300, device="cuda")
    ops.colsum(Xr.cuda(), 100, 300, out)
    torch.testing.assert_close(out.cpu(), Xr[:, :300].float().sum(0), rtol=1e-4, atol=1e-3)
    src = torch.randn(77, 207, generator=g)
    dst = torch.full((77, 208), 5.0, dtype=torch.bfloat16, device="cuda")
    ops.cast_pad(src.cuda(), dst, 77, 207)
    assert torch.equal(dst[:, :207].cpu(), src.to(torch.bfloat16)) and torch.all(dst[:, 207] == 0)
    flat = torch.randn(4099 * 8, generator=g)
    d2 = torch.empty(4099 * 8, dtype=torch.bfloat16, device="cuda")
    ops.cast_f32_bf16(flat.cuda(), d2)
    assert torch.equal(d2.cpu(), flat.to(torch.bfloat16))
    x = torch.randn(5000, generator=g).to(torch.bfloat16)
    y = torch.empty(5000, dtype=torch.bfloat16, device="cuda")
    ops.act_fwd(x.cuda(), y, "silu")
    assert (y.float().cpu() - torch.nn.functional.silu(x.float())).abs().max().item() <= 2e-2
    xr = x.float().requires_grad_(True)
    torch.nn.functional.silu(xr).backward(torch.ones(5000))
    dx = torch.empty(5000, dtype=torch.bfloat16, device="cuda")
    ops.act_bwd(torch.ones(5000, dtype=torch.bfloat16, device="cuda"), x.cuda(), dx, "silu")
    assert (dx.float().cpu() - xr.grad).abs().max().item() <= 1e-2
    w = torch.randn(128, 177, 7, generator=g)
    wg = torch.empty(128, 7 * 192, dtype=torch.bfloat16, device="cuda")
    ops.conv_weight_to_gemm(w.cuda(), wg, 192)
    want = torch.zeros(128, 7, 192); want[:, :, :177] = w.permute(0, 2, 1)
    assert torch.equal(wg.cpu(), want.reshape(128, -1).to(torch.bfloat16))


# ---------------------------------------- BatchNorm1d ---------------------------------------------
@pytest.mark.parametrize("M,C", [(6, 10), (300, 1470), (4096, 512), (70001, 147)])
@pytest.mark.parametrize("training", [True, False])
def test_batchnorm_fwd_bwd_vs_torch(M, C, training):
    """ibm_batchnorm_fwd/bwd against torch.nn.functional.batch_norm in fp32 on the same bf16-rounded rows
    (nn.BatchNorm1d semantics, FeedForward…py:71-72): outputs within one bf16 rounding, statistics and parameter
    gradients rtol 2e-3 (fp32 sums over M rows in a different order), running stats with the unbiased variance.
    M = 300 and 70001 exercise ragged last chunks and the > 64-chunk regrouping; C = 1470 / 147 / 10 odd tails."""
    from inferbiomechanics_b200 import ops
    g = torch.Generator().manual_seed(M + C)
    ld = ops.round_up(C, 8)
    x32 = (torch.randn(M, C, generator=g) * (1 + torch.arange(C) % 5) + 3.0 * torch.randn(C, generator=g)).bfloat16().float()
    x = torch.zeros(M, ld, dtype=torch.bfloat16, device="cuda")
    x[:, :C] = x32.cuda()
    gamma, beta = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    rm, rv = torch.randn(C, generator=g).cuda(), (torch.rand(C, generator=g) + 0.5).cuda()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y = torch.full((M, ld), 7.0, dtype=torch.bfloat16, device="cuda")
    mean, rstd = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ws = ops.batchnorm_workspace(C, "cuda")
    ops.batchnorm_fwd(x, y, M, C, gamma, beta, rm, rv, mean, rstd, training, 0.1, 1e-5, ws)
    xr = x32.cuda().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = torch.nn.functional.batch_norm(xr, rm_ref, rv_ref, gr, br, training=training, momentum=0.1, eps=1e-5)
    err = (y[:, :C].float() - yr).abs().max().item()
    assert err <= 2 ** -7 * yr.abs().max().item() + 1e-6, err            # one bf16 rounding of the output
    assert (y[:, C:] == 0).all()
    if training:
        torch.testing.assert_close(rm, rm_ref, rtol=2e-4, atol=2e-5)
        torch.testing.assert_close(rv, rv_ref, rtol=2e-3, atol=1e-5)
        torch.testing.assert_close(mean, x32.cuda().mean(0), rtol=2e-4, atol=2e-5)
    else:
        assert torch.equal(rm, rm_ref) and torch.equal(rv, rv_ref)
    # backward
    dy32 = torch.randn(M, C, generator=g).bfloat16().float()
    dy = torch.zeros(M, ld, dtype=torch.bfloat16, device="cuda")
    dy[:, :C] = dy32.cuda()
    yr.backward(dy32.cuda())
    dgamma, dbeta = torch.ones(C, device="cuda"), torch.ones(C, device="cuda")     # accumulate (+=) semantics
    dx = torch.empty(M, ld, dtype=torch.bfloat16, device="cuda")
    m_, r_ = (mean, rstd) if training else (rm, rv)
    cs = torch.zeros(C, device="cuda")
    ops.batchnorm_bwd(dy, x, dx, M, C, gamma, m_, r_, training, 1e-5, dgamma, dbeta, ws, dx_colsum=cs)
    scale = xr.grad.abs().max().item()
    # fp32 column sums of dx (upstream bias gradient): ~0 in training mode, so measure against the sum of magnitudes
    assert ((cs - xr.grad.sum(0)).abs() <= 2e-4 * xr.grad.abs().sum(0) + 1e-6).all()
    assert (dx[:, :C].float() - xr.grad).abs().max().item() <= 2 ** -7 * scale + 2e-3 * scale
    torch.testing.assert_close(dgamma - 1, gr.grad, rtol=3e-3, atol=3e-3 * gr.grad.abs().max().item())
    torch.testing.assert_close(dbeta - 1, br.grad, rtol=3e-3, atol=3e-3 * br.grad.abs().max().item())
    # in place (dx aliases dy): same result (the column sums are atomics over row chunks, so the last bit may differ)
    ops.batchnorm_bwd(dy, x, dy, M, C, gamma, m_, r_, training, 1e-5, None, None, ws)
    assert (dy[:, :C].float() - dx[:, :C].float()).abs().max().item() <= 2 ** -7 * scale


def test_batchnorm_single_row_training_is_a_value_error():
    from inferbiomechanics_b200 import ops
    x = torch.zeros(1, 8, dtype=torch.bfloat16, device="cuda")
    o = torch.ones(8, device="cuda")
    with pytest.raises(ValueError):              # torch: "Expected more than 1 value per channel when training"
        ops.batchnorm_fwd(x, x.clone(), 1, 8, o, o, o.clone(), o.clone(), o.clone(), o.clone(), True, 0.1, 1e-5,
                          ops.batchnorm_workspace(8, "cuda"))


@pytest.mark.parametrize("B,T,E", [(3, 200, 30), (1, 20, 30), (5, 64, 0), (2, 65, 7), (300, 50, 30)])
def test_pack_channel_major_bit_exact(B, T, E):
    """TransformerBaseline input pack (TransformerBaseline.py:108-126: cat on the channel dim, transpose(1, 2), temporal
    embedding concatenated): bf16 rows equal torch's own cat/transpose/cat followed by an RNE cast, bit for bit; pad
    columns are zero."""
    from inferbiomechanics_b200 import ops
    g = torch.Generator().manual_seed(B * T + E)
    chans = [23, 23, 23, 3, 3, 3]
    srcs = [torch.randn(B, c, T, generator=g) for c in chans]
    emb = torch.randn(T + 3, E, generator=g) if E else None
    C = sum(chans)
    ld = ops.round_up(C + E, 8)
    out = torch.full((B * T, ld), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.pack_channel_major([s.cuda() for s in srcs], T, None if emb is None else emb.cuda(), out)
    want = torch.cat(srcs, dim=1).transpose(1, 2)
    if E:
        want = torch.cat([want, emb[:T].unsqueeze(0).expand(B, T, E)], dim=2)
    want = want.reshape(B * T, C + E).to(torch.bfloat16)
    assert torch.equal(out[:, :C + E].cpu(), want)
    assert (out[:, C + E:] == 0).all()


# ---------------------------------------- walk order -----------------------------------------------
def test_walk_order_does_not_change_results():
    """ibm_set_walk_order: GEMM work items, LayerNorm rows and attention windows taken in DESCENDING order (mode 2 alternates
    every launch, so two launches cover both directions) give the ascending order's results — bit for bit where the
    arithmetic has no cross-item sums (everything except split-K reduce-adds and fp32 atomics of column sums)."""
    from inferbiomechanics_b200 import ops
    g = torch.Generator().manual_seed(77)
    dev = "cuda"

    def run_all():
        out = {}
        M, N, K = 1111, 768, 512                       # ragged M: the last row block is partial in either direction
        A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(dev)
        W = (torch.randn(N, K, generator=g) * 0.05).to(torch.bfloat16).to(dev)
        bias = torch.randn(N, generator=g).to(dev)
        res = torch.randn(M, N, generator=g).to(torch.bfloat16).to(dev)
        y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        out["gemm_plain"] = ops.gemm(A, W, y, M, N, K, bias=bias, act="relu").clone()
        out["gemm_aux"] = ops.gemm(A, W, torch.empty_like(y), M, N, K, bias=bias, aux=res, aux_mode=1).clone()
        mask = torch.zeros(M, N // 8, dtype=torch.uint8, device=dev)
        out["gemm_mask_w"] = ops.gemm(A, W, torch.empty_like(y), M, N, K, bias=bias, act="relu", mask=mask, mask_mode=1).clone()
        out["mask"] = mask.clone()
        cs = torch.zeros(N, device=dev)
        out["gemm_mask_r"] = ops.gemm(A, W, torch.empty_like(y), M, N, K, mask=mask, mask_mode=2, colsum=cs).clone()
        out["~colsum"] = cs
        M2 = 2048 * 4                                  # supertiles (K >= 1024, enough row blocks) with the aux ring
        A2 = torch.randn(M2, 1024, generator=g).to(torch.bfloat16).to(dev)
        W2 = (torch.randn(512, 1024, generator=g) * 0.05).to(torch.bfloat16).to(dev)
        r2 = torch.randn(M2, 512, generator=g).to(torch.bfloat16).to(dev)
        out["gemm_super_aux"] = ops.gemm(A2, W2, torch.empty(M2, 512, dtype=torch.bfloat16, device=dev), M2, 512, 1024, aux=r2, aux_mode=1).clone()
        wg = torch.zeros(512, 1024, device=dev)
        d2 = torch.randn(M2, 512, generator=g).to(torch.bfloat16).to(dev)
        ops.gemm(d2, A2, wg, 512, 1024, M2, a_mn=True, b_mn=True, accumulate=True)
        out["~wgrad"] = wg
        Ml, d = 3001, 512
        s = (torch.randn(Ml, d, generator=g) * 2).to(torch.bfloat16).to(dev)
        gamma, beta = (1 + 0.1 * torch.randn(d, generator=g)).to(dev), (0.1 * torch.randn(d, generator=g)).to(dev)
        yl = torch.empty_like(s)
        mean, rstd = torch.empty(Ml, device=dev), torch.empty(Ml, device=dev)
        ops.layernorm_fwd(s, yl, gamma, beta, Ml, d, mean=mean, rstd=rstd)
        out["ln_y"], out["ln_mean"], out["ln_rstd"] = yl, mean, rstd
        dy = torch.randn(Ml, d, generator=g).to(torch.bfloat16).to(dev)
        ds = torch.empty_like(s)
        dg, db, dc = (torch.zeros(d, device=dev) for _ in range(3))
        ops.layernorm_bwd(dy, s, gamma, mean, rstd, Ml, d, ds, dg, db, dc)
        out["ln_ds"], out["~ln_dg"], out["~ln_db"], out["~ln_dc"] = ds, dg, db, dc
        n_win, T, H, hd = 301, 50, 8, 64               # odd window count: the last pair of the backward is half empty
        qkv = (torch.randn(n_win * T, 3 * H * hd, generator=g) * 0.8).to(torch.bfloat16).to(dev)
        o = torch.empty(n_win * T, H * hd, dtype=torch.bfloat16, device=dev)
        ops.attention_fwd_fused(qkv, H * hd, o, n_win, T, H, hd, 0.125)
        out["attn_o"] = o
        do = torch.randn(n_win * T, H * hd, generator=g).to(torch.bfloat16).to(dev)
        dqkv = torch.empty_like(qkv)
        dbias = torch.zeros(3 * H * hd, device=dev)
        ops.attention_bwd(qkv, H * hd, do, dqkv, n_win, T, H, hd, 0.125, dbias=dbias)
        out["attn_dqkv"], out["~attn_dbias"] = dqkv, dbias
        torch.cuda.synchronize()
        return out

    try:
        ops.set_walk_order(0)
        g.manual_seed(77)
        ref = run_all()
        for rep in range(2):                           # every launch alternates: the two passes swap directions per op
            ops.set_walk_order(2)
            if rep == 1:
                ops.layernorm_fwd(ref["ln_y"], torch.empty_like(ref["ln_y"]), torch.ones(512, device=dev), torch.zeros(512, device=dev), 8, 512)
            g.manual_seed(77)
            got = run_all()
            for k, v in ref.items():
                if k.startswith("~"):                  # sums across work items: order of fp32 additions differs
                    torch.testing.assert_close(got[k], v, rtol=1e-4, atol=1e-3 * (v.abs().max().item() + 1e-6), msg=k)
                else:
                    assert torch.equal(got[k], v), f"{k} differs with walk order 2, pass {rep}"
    finally:
        ops.set_walk_order(0)
    with pytest.raises(Exception):
        ops.set_walk_order(3)
