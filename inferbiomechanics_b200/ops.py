"""Tensor-level wrappers over the C ABI (include/ibm_b200.h).

PyTorch is used here only for device memory and streams: every function takes CUDA tensors,
passes raw device pointers + sizes + the current stream to libibm_b200.so, and returns.  No
function falls back to a torch implementation.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT, BF16, F32, call

_ws_cache = {}


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.IbmError("inferbiomechanics_b200 kernels need CUDA tensors (no CPU fallback)")


def workspace(device) -> torch.Tensor:
    dev = torch.device(device)
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    if key not in _ws_cache:
        _ws_cache[key] = torch.zeros(_lib.load().ibm_workspace_bytes(), dtype=torch.uint8, device=dev)
    return _ws_cache[key]


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def pack_channel_major(srcs, T, emb, out_bf16):
    """srcs: contiguous fp32 CUDA tensors (B, C_i, T); emb fp32 [>=T, E] or None; out bf16 [B*T, ld]."""
    n = len(srcs)
    B = srcs[0].shape[0]
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
    chans = (ctypes.c_int32 * n)(*[t.shape[1] for t in srcs])
    call("ibm_pack_channel_major", ptrs, chans, n, B, T, _p(emb), 0 if emb is None else emb.shape[1], _p(out_bf16), out_bf16.stride(0),
         stream_ptr())


def expand_rows_bf16(src_bf16, C, T, emb, out_bf16, v_out=None, v_col0=0):
    """src bf16 [B*T, ld_src] frame-major rows (pre-packed on the host) -> out bf16 [B*T, ld] with the temporal embedding appended."""
    n_rows = out_bf16.shape[0]
    call("ibm_expand_rows_bf16", _p(src_bf16), src_bf16.stride(0), C, n_rows, T, _p(emb), 0 if emb is None else emb.shape[1],
         _p(out_bf16), out_bf16.stride(0), _p(v_out), v_col0, stream_ptr())


# ---- GEMM ---------------------------------------------------------------------------------------
def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, M: int, N: int, K: int, *, lda=None, ldb=None, ldd=None,
         a_mn=False, b_mn=False, bias: Optional[torch.Tensor] = None, act="none", aux: Optional[torch.Tensor] = None,
         ldaux: int = 0, aux_mode: int = 0, accumulate=False, split_k: int = 0, taps: int = 1,
         colsum: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None, mask_mode: int = 0) -> torch.Tensor:
    """out[M,N] (+)= epilogue(A[M,K] · B[N,K]^T); see ibm_gemm_bf16.  Leading dims default to the
    tensors' row strides.  colsum (fp32 [N]): += column sums of the bf16 output (a bias gradient).
    mask (uint8 [M, N/8]) with mask_mode 1: sign bits of the output are written; mask_mode 2: outputs whose bit is clear
    are zeroed (ReLU derivative without re-reading the activation)."""
    _require_cuda(A, B, out)
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    lda = A.stride(0) if lda is None else lda
    ldb = B.stride(0) if ldb is None else ldb
    ldd = out.stride(0) if ldd is None else ldd
    if aux is not None and ldaux == 0:
        ldaux = aux.stride(0)
    od = F32 if out.dtype == torch.float32 else BF16
    call("ibm_gemm_bf16", _p(A), lda, int(a_mn), _p(B), ldb, int(b_mn), M, N, K, _p(bias), ACT[act], _p(aux), ldaux,
         aux_mode, _p(out), ldd, od, int(accumulate), split_k, taps, _p(colsum), _p(mask),
         0 if mask is None else mask.stride(0), mask_mode if mask is not None else 0, stream_ptr())
    return out


def set_walk_order(mode: int) -> None:
    """0: kernels walk rows / tiles in ascending order; 1: big launches (>= 48 MB streamed) alternate ascending /
    descending so a consumer starts on what its producer wrote last (still in L2); 2: every launch alternates (tests)."""
    call("ibm_set_walk_order", int(mode))


def colsum(X: torch.Tensor, M: int, N: int, out: torch.Tensor, ld=None) -> None:
    call("ibm_colsum_bf16", _p(X), X.stride(0) if ld is None else ld, M, N, _p(out), stream_ptr())


def act_fwd(x, y, act):
    call("ibm_act_fwd", _p(x), _p(y), x.numel(), ACT[act], stream_ptr())


def act_bwd(dy, x, dx, act):
    call("ibm_act_bwd", _p(dy), _p(x), _p(dx), x.numel(), ACT[act], stream_ptr())


def cast_f32_bf16(src, dst):
    call("ibm_cast_f32_bf16", _p(src), _p(dst), src.numel(), stream_ptr())


def cast_bf16_f32(src, dst):
    call("ibm_cast_bf16_f32", _p(src), _p(dst), src.numel(), stream_ptr())


def cast_pad(src: torch.Tensor, dst: torch.Tensor, rows: int, cols: int, ld_src=None, ld_dst=None):
    call("ibm_cast_pad_f32_bf16", _p(src), src.stride(0) if ld_src is None else ld_src, _p(dst),
         dst.stride(0) if ld_dst is None else ld_dst, rows, cols, stream_ptr())


def conv_weight_to_gemm(w: torch.Tensor, dst: torch.Tensor, cin_pad: int):
    cout, cin, kt = w.shape
    call("ibm_conv_weight_to_gemm", _p(w), cout, cin, kt, cin_pad, _p(dst), stream_ptr())


def conv_wgrad_from_gemm(g: torch.Tensor, dw: torch.Tensor, cin_pad: int, accumulate: bool):
    cout, cin, kt = dw.shape
    call("ibm_conv_wgrad_from_gemm", _p(g), cout, cin, kt, cin_pad, _p(dw), int(accumulate), stream_ptr())


def replicate_pad_rows(X, n_win, T, pad, cols):
    call("ibm_replicate_pad_rows", _p(X), X.stride(0), n_win, T, pad, cols, stream_ptr())


def fold_pad_rows(G, n_win, T, pad, cols):
    call("ibm_fold_pad_rows", _p(G), G.stride(0), n_win, T, pad, cols, stream_ptr())


def dropout(x, y, p, seed, offset, step_dev=None, step_mul=0):
    """Philox offset = offset (+ step_mul * step_dev[0], read on the device at execution time, when a counter is given)."""
    if step_dev is None:
        call("ibm_dropout_bf16", _p(x), _p(y), x.numel(), p, seed, offset, stream_ptr())
    else:
        call("ibm_dropout_bf16_dev", _p(x), _p(y), x.numel(), p, seed, offset, _p(step_dev), step_mul, stream_ptr())


def counter_add(counter_dev, inc=1):
    call("ibm_counter_add", _p(counter_dev), inc, stream_ptr())


def conv_weight_to_dgrad(w: torch.Tensor, dst: torch.Tensor, cout_pad: int):
    cout, cin, kt = w.shape
    call("ibm_conv_weight_to_dgrad", _p(w), cout, cin, kt, cout_pad, _p(dst), stream_ptr())


# ---- BatchNorm1d -----------------------------------------------------------------------------------
def batchnorm_workspace(C: int, device) -> torch.Tensor:
    return torch.zeros(_lib.load().ibm_batchnorm_workspace_floats(C), dtype=torch.float32, device=device)


def batchnorm_fwd(x, y, M, C, gamma, beta, running_mean, running_var, save_mean, save_rstd, training, momentum, eps, ws):
    call("ibm_batchnorm_fwd", _p(x), x.stride(0), _p(y), y.stride(0), M, C, _p(gamma), _p(beta), _p(running_mean), _p(running_var),
         _p(save_mean), _p(save_rstd), int(training), momentum, eps, _p(ws), stream_ptr())


def batchnorm_bwd(dy, x, dx, M, C, gamma, mean, rstd_or_var, training, eps, dgamma, dbeta, ws, act_out=None, act=None,
                  dx_colsum=None):
    call("ibm_batchnorm_bwd", _p(dy), dy.stride(0), _p(x), x.stride(0), _p(dx), 0 if dx is None else dx.stride(0), M, C, _p(gamma),
         _p(mean), _p(rstd_or_var), int(training), eps, _p(act_out), 0 if act_out is None else act_out.stride(0), ACT[act],
         _p(dgamma), _p(dbeta), _p(dx_colsum), _p(ws), stream_ptr())


# ---- LayerNorm ----------------------------------------------------------------------------------
def layernorm_fwd(s, y, gamma, beta, M, d, eps=1e-5, mean=None, rstd=None, ld=None):
    call("ibm_layernorm_fwd", _p(s), _p(y), s.stride(0) if ld is None else ld, _p(gamma), _p(beta), M, d, eps, _p(mean),
         _p(rstd), stream_ptr())


def layernorm_bwd(dy, s, gamma, mean, rstd, M, d, ds, dgamma, dbeta, dcolsum=None, ld=None):
    call("ibm_layernorm_bwd", _p(dy), _p(s), s.stride(0) if ld is None else ld, _p(gamma), _p(mean), _p(rstd), M, d,
         _p(ds), _p(dgamma), _p(dbeta), _p(dcolsum), stream_ptr())


# ---- attention ----------------------------------------------------------------------------------
def attention_fwd_fused(qkv, kv_off, o, n_win, T, H, hd, scale):
    es = qkv.element_size()
    base = qkv.data_ptr()
    call("ibm_attention_fwd", base, qkv.stride(0), base + kv_off * es, qkv.stride(0), base + 2 * kv_off * es,
         qkv.stride(0), _p(o), o.stride(0), n_win, T, H, hd, hd, scale, stream_ptr())


def attention_fwd(q, k, v, o, n_win, T, H, hd_qk, hd_v, scale):
    call("ibm_attention_fwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0), n_win, T, H,
         hd_qk, hd_v, scale, stream_ptr())


def attention_bwd(qkv, kv_off, d_o, dqkv, n_win, T, H, hd, scale, dbias=None):
    """dbias (fp32 [3*kv_off]): the column sums of dqkv are added to it (in_proj_bias gradient)."""
    call("ibm_attention_bwd", _p(qkv), qkv.stride(0), kv_off, _p(d_o), d_o.stride(0), _p(dqkv), n_win, T, H, hd, scale,
         _p(dbias), stream_ptr())


def attention_bwd_long(q, k, v, o, d_o, dq, dk, dv, n_win, T, H, hd_qk, hd_v, scale, dbq=None, dbk=None, dbv=None):
    """Backward for whole windows of up to 256 frames (ibm_attention_bwd_long): q/k/v/o/d_o/dq/dk/dv are bf16 row-major
    views (one row per (window, frame), head h at columns h*hd); dv None when the values are an input."""
    call("ibm_attention_bwd_long", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(o), o.stride(0), _p(d_o),
         d_o.stride(0), _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), 0 if dv is None else dv.stride(0), n_win, T, H,
         hd_qk, hd_v, scale, _p(dbq), _p(dbk), _p(dbv), stream_ptr())


# ---- regression loss ------------------------------------------------------------------------------
def _ptr_array(ts: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in ts])
    return arr


def _stride_array(ts: Sequence[torch.Tensor]):
    vals = []
    for t in ts:
        assert t.dim() == 3 and t.stride(2) == 1, "loss tensors must be (B,F,C) with unit channel stride"
        vals += [t.stride(0), t.stride(1)]
    return (ctypes.c_int64 * 8)(*vals)


def regression_loss_fwd(outs: Sequence[torch.Tensor], labs: Sequence[torch.Tensor], weights30: Sequence[float],
                        threshold: float = 10.0, result: Optional[torch.Tensor] = None) -> torch.Tensor:
    """outs/labs in quantity order (cop, force, torque, wrench), fp32 (B,F,C).  Returns fp32[40]."""
    _require_cuda(*outs, *labs)
    B, F = outs[0].shape[0], outs[0].shape[1]
    if result is None:
        result = torch.empty(40, dtype=torch.float32, device=outs[0].device)
    w = (ctypes.c_float * 30)(*weights30)
    call("ibm_regression_loss_fwd", _ptr_array(outs), _stride_array(outs), _ptr_array(labs), _stride_array(labs), B, F, w,
         threshold, _p(result), _p(workspace(outs[0].device)), stream_ptr())
    return result


def regression_loss_bwd(outs, labs, weights30, grads: Sequence[torch.Tensor], upstream: Optional[torch.Tensor] = None,
                        threshold: float = 10.0) -> None:
    B, F = outs[0].shape[0], outs[0].shape[1]
    w = (ctypes.c_float * 30)(*weights30)
    gd = F32 if grads[0].dtype == torch.float32 else BF16
    call("ibm_regression_loss_bwd", _ptr_array(outs), _stride_array(outs), _ptr_array(labs), _stride_array(labs), B, F, w,
         threshold, _p(upstream), _ptr_array(grads), _stride_array(grads), gd, stream_ptr())


# ---- the evaluator's static helpers, general contract (csrc/loss_helpers.cu) -----------------------------
def sqdiff_mean_vector(o: torch.Tensor, l: torch.Tensor) -> torch.Tensor:
    _require_cuda(o, l)
    B, F, C = o.shape
    out = torch.empty(C, dtype=torch.float32, device=o.device)
    call("ibm_sqdiff_mean_vector", _p(o), o.stride(0), o.stride(1), _p(l), l.stride(0), l.stride(1), B, F, C, _p(out),
         _p(workspace(o.device)), stream_ptr())
    return out


def sqdiff_mean_vector_bwd(o, l, upstream, grad_out=None, grad_lab=None) -> None:
    B, F, C = o.shape
    call("ibm_sqdiff_mean_vector_bwd", _p(o), o.stride(0), o.stride(1), _p(l), l.stride(0), l.stride(1), B, F, C, _p(upstream),
         _p(grad_out), _p(grad_lab), stream_ptr())


def mask_by_threes(x: torch.Tensor, threshold: float) -> torch.Tensor:
    _require_cuda(x)
    B, F, C = x.shape
    out = torch.empty(B, F, C, dtype=torch.float32, device=x.device)
    call("ibm_mask_by_threes", _p(x), x.stride(0), x.stride(1), B, F, C, float(threshold), _p(out), stream_ptr())
    return out


def mean_norm_error(o: torch.Tensor, l: torch.Tensor, vec_size: int, fold_halves: bool = False) -> torch.Tensor:
    _require_cuda(o, l)
    B, F, C = o.shape
    out = torch.empty(1, dtype=torch.float32, device=o.device)
    call("ibm_mean_norm_error", _p(o), o.stride(0), o.stride(1), _p(l), l.stride(0), l.stride(1), B, F, C, int(vec_size),
         1 if fold_halves else 0, _p(out), _p(workspace(o.device)), stream_ptr())
    return out[0]


# ---- DDPM -----------------------------------------------------------------------------------------
def q_sample(x0, eps, t, sqrt_abar, sqrt_1m_abar, xt_f32=None, xt_bf16=None, bf16_ld=0, seed=0, offset=0, eps_out=None):
    B = x0.shape[0]
    per_win = x0.numel() // B
    call("ibm_q_sample", _p(x0), _p(eps), _p(t), _p(sqrt_abar), _p(sqrt_1m_abar), B, per_win, _p(xt_f32), _p(xt_bf16), bf16_ld,
         seed, offset, _p(eps_out), stream_ptr())


def posterior_step(x0_hat, x0_ld, x_t, z, t_dev, coef_x0, coef_xt, sigma, M, x_prev=None, xprev_bf16=None, bf16_ld=0,
                   seed=0, offset=0, t_next=None):
    call("ibm_ddpm_posterior_step", _p(x0_hat), x0_ld, _p(x_t), _p(z), _p(t_dev), _p(coef_x0), _p(coef_xt), _p(sigma), M,
         _p(x_prev), _p(xprev_bf16), bf16_ld, seed, offset, _p(t_next), stream_ptr())


def timestep_embed(t, out_bf16, dim, scalar=False):
    B = out_bf16.shape[0]
    call("ibm_timestep_embed", _p(t), int(scalar), B, dim, _p(out_bf16), stream_ptr())


def add_time_pos(h, temb, pos, M, F, d, t_row=None):
    """t_row (device int32[1]): temb is a per-timestep table and row t_row[0] is added to every window."""
    call("ibm_add_time_pos", _p(h), h.stride(0), _p(temb), temb.stride(0), _p(pos), M, F, d, _p(t_row), stream_ptr())


def add_time_pos_bwd(dh, dtemb, dpos, M, F, d, dbias=None):
    """dbias (fp32 [d]): += the column sums of dh over all rows (bias gradient of the Linear that produced h), same pass."""
    call("ibm_add_time_pos_bwd", _p(dh), dh.stride(0), _p(dtemb), dtemb.stride(0), _p(dpos), M, F, d, _p(dbias), stream_ptr())


# ---- window batcher ---------------------------------------------------------------------------------
def window_valid_mask(missing, trial_base, cand_trial, cand_start, window_size, stride, valid):
    call("ibm_window_valid_mask", _p(missing), _p(trial_base), _p(cand_trial), _p(cand_start), cand_trial.numel(),
         window_size, stride, _p(valid), stream_ptr())


def pack_windows(frames, C, win_row0, F, stride, out_f32=None, out_bf16=None, frame_stride=0, win_extra=0, col0=0):
    call("ibm_pack_windows", _p(frames), frames.stride(0), C, _p(win_row0), win_row0.numel(), F, stride, _p(out_f32),
         _p(out_bf16), frame_stride, win_extra, col0, stream_ptr())


def pack_inputs(srcs, n_rows, F, out_f32=None, out_bf16=None, frame_stride=0, win_extra=0, col0=0):
    """srcs: list of contiguous fp32 CUDA tensors [n_rows, w_k] (any leading shape that flattens to n_rows)."""
    n = len(srcs)
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
    widths = (ctypes.c_int32 * n)(*[t.shape[-1] for t in srcs])
    call("ibm_pack_inputs", ptrs, widths, n, n_rows, F, _p(out_f32), _p(out_bf16), frame_stride, win_extra, col0, stream_ptr())


def pack_labels(raw, nb, win_row0, contact_idx, mass, F, stride, last_only, out_rows):
    call("ibm_pack_labels", _p(raw), raw.stride(0), nb, _p(win_row0), _p(contact_idx), _p(mass), win_row0.numel(), F, stride,
         int(last_only), _p(out_rows), out_rows.stride(0), stream_ptr())


# ---- optimizer ----------------------------------------------------------------------------------------
def optimizer_step(kind: str, param, grad, state0, state1, param_bf16, lr, grad_scale, step, step_dev=None):
    """``step_dev`` (device int64[1]) replaces the by-value 1-based step count (graph-replayed steps)."""
    if step_dev is not None:
        call("ibm_optimizer_step_dev", _lib.OPT_KIND[kind], _p(param), _p(grad), _p(state0), _p(state1), _p(param_bf16), param.numel(),
             lr, grad_scale, _p(step_dev), stream_ptr())
        return
    call("ibm_optimizer_step", _lib.OPT_KIND[kind], _p(param), _p(grad), _p(state0), _p(state1), _p(param_bf16), param.numel(),
         lr, grad_scale, step, stream_ptr())
