"""Drop-in ``TransformerBaseline`` (forward / analysis pass) on libibm_b200.

Constructor signature, sub-module names and ``state_dict`` keys follow
``/root/reference/src/models/TransformerBaseline.py:8-148`` (``temporal_embedding.embedding.weight``,
``transformer_layers.{l}.multihead_attention.in_proj_weight`` …, ``fc.*``,
``com_attention.{query,key}_linear.*``; fp64 parameters by default) so a reference checkpoint loads
unchanged.  ``forward`` reproduces lines 104-148: inputs are (B, C, T) tensors keyed ``pos, vel, acc,
comPos, comVel, comAcc`` (the three COM keys do not exist in the reference's ``InputDataKeys`` — SURVEY
§0.3 — and are defined in ``inferbiomechanics_b200.keys``), learned temporal embedding concatenated,
3 post-LN encoder layers, ``fc`` head, unscaled single-head "CoM blend" attention.

B200 mapping: d (108 for the dataset's 23 DOF) is padded to a multiple of 16 columns (112) and each head (36-wide) to
32 / 48 / 64 (48) so every row is a
16-byte multiple (TMA-legal) and head slices are 16-byte aligned; the pads carry exact zeros (zero
weight rows/columns), so results are those of the unpadded model.  GEMMs run in bf16 on tcgen05 with
fp32 accumulation; the reference computes in fp64 — tolerance stated in tests/test_gpu_transformer.py.
The whole sequence (T <= 256) of a (window, head) stays in shared memory; long streams shard by window.

Inference (BASELINE configs[4], the analyze pass) runs under ``torch.no_grad()`` over re-used buffers.  With autograd
enabled the outputs carry a graph (``_TransformerFunction``): the backward of the whole stack — both heads, the CoM blend
(``ibm_attention_bwd_long`` without dv: its values are a model input), every layer (attention backward for whole
windows of up to 256 frames, LayerNorm backward over the 108 valid of 112 columns, tcgen05 dgrad / wgrad GEMMs) and
the temporal-embedding gradient (a sum over windows of the last 30 input-gradient columns) — runs on libibm_b200 and
hands fp64 gradients back to the reference-shaped ``nn.Parameter``s, so ``loss.backward(); optimizer.step()`` works as
on the reference.  Dropout > 0 in training mode is not implemented (the reference default is 0.0).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib, ops
from ..keys import InputDataKeys, OutputDataKeys

BF16 = torch.bfloat16


class TransformerLayer(nn.Module):
    """Parameter container with the reference layer's module names (TransformerBaseline.py:8-22)."""

    def __init__(self, timestep_vector_dim: int, num_heads: int, dim_feedforward: int, dropout: float, dtype=torch.float64):
        super().__init__()
        self.multihead_attention = nn.MultiheadAttention(timestep_vector_dim, num_heads, dropout=dropout, batch_first=True, dtype=dtype)
        self.feedforward = nn.Sequential(nn.Linear(timestep_vector_dim, dim_feedforward, dtype=dtype), nn.ReLU(),
                                         nn.Linear(dim_feedforward, timestep_vector_dim, dtype=dtype))
        self.dropout1 = nn.Dropout(dropout)
        self.dropout2 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(timestep_vector_dim, dtype=dtype)
        self.norm2 = nn.LayerNorm(timestep_vector_dim, dtype=dtype)


class TemporalEmbedding(nn.Module):
    def __init__(self, window_size: int, embedding_dim: int, dtype=torch.float64):
        super().__init__()
        self.embedding = nn.Embedding(window_size, embedding_dim, dtype=dtype)


class SimpleAttention(nn.Module):
    def __init__(self, key_query_dim: int, dtype=torch.float64):
        super().__init__()
        self.query_linear = nn.Linear(key_query_dim, key_query_dim, dtype=dtype)
        self.key_linear = nn.Linear(key_query_dim, key_query_dim, dtype=dtype)


def _pad_head(n: int) -> int:
    for c in (32, 48, 64):
        if n <= c:
            return c
    raise NotImplementedError(f"head_dim {n} > 64 is not supported by the attention kernel")


class _TransformerFunction(torch.autograd.Function):
    """Autograd bridge for training: forward/backward are launch sequences over saved activations; the parameter
    tensors are passed only so that autograd (and DDP's hooks) see them."""

    @staticmethod
    def forward(ctx, model, x, *params):
        out, blend, saved = model._forward_launches(x, save=True)
        ctx.model, ctx.saved = model, saved
        return out.clone(), blend.float()

    @staticmethod
    def backward(ctx, g_out, g_blend):
        grads = ctx.model._backward_launches(ctx.saved, g_out, g_blend)
        ctx.saved = None
        return (None, None, *grads)


class TransformerBaseline(nn.Module):
    timestep_vector_dim: int
    window_size: int
    temporal_embedding_dim: int
    output_vector_dim: int

    def __init__(self, dofs: int, window_size: int, temporal_embedding_dim: int = 30, num_layers: int = 3, num_heads: int = 3,
                 dim_feedforward: int = 60, dropout: float = 0.0, dtype=torch.float64):
        super().__init__()
        self.timestep_vector_dim = (dofs * 3) + (3 * 3) + temporal_embedding_dim
        self.output_vector_dim = (2 + 3 + 6)
        self.window_size = window_size
        self.temporal_embedding_dim = temporal_embedding_dim
        self.num_heads, self.dim_feedforward, self.num_layers, self.dofs = num_heads, dim_feedforward, num_layers, dofs
        self.temporal_embedding = TemporalEmbedding(window_size, temporal_embedding_dim, dtype=dtype)
        self.transformer_layers = nn.ModuleList([
            TransformerLayer(self.timestep_vector_dim, num_heads, dim_feedforward, dropout, dtype=dtype) for _ in range(num_layers)])
        self.fc = nn.Linear(self.timestep_vector_dim, self.output_vector_dim, dtype=dtype)
        self.contact_sigmoid = nn.Sigmoid()
        self.com_attention = SimpleAttention(self.timestep_vector_dim)          # reference leaves this at its fp64 default
        self._prep = None
        self._prep_version = None
        self._bufs: Dict[int, Dict[str, torch.Tensor]] = {}

    # ---- padded bf16 weights --------------------------------------------------------------------------
    def _prepare(self, dev):
        ver = sum(p._version for p in self.parameters()) + sum(hash(p.data_ptr()) % 1000003 for p in self.parameters())
        if self._prep is not None and self._prep_version == ver:
            return self._prep
        d, H, ff = self.timestep_vector_dim, self.num_heads, self.dim_feedforward
        hd = d // H
        # rows are padded to a multiple of 16 columns: the CoM blend contracts over the whole row with m16n8k16 MMAs
        dp, hp, fp = ops.round_up(d, 16), _pad_head(hd), ops.round_up(ff, 8)
        if dp not in (64, 80, 96, 112, 128):
            raise NotImplementedError(f"timestep vector width {d} (= 3*dofs + 9 + temporal_embedding_dim) pads to {dp}: the CoM-blend "
                                      "attention kernels are instantiated for padded widths 64-128 (the dataset's 23 DOF give 112)")

        def padw(w, rows, cols):
            out = torch.zeros(rows, cols, dtype=BF16, device=dev)
            out[:w.shape[0], :w.shape[1]] = w.detach().to(dev, torch.float32).to(BF16)
            return out

        def padb(b, n):
            out = torch.zeros(n, dtype=torch.float32, device=dev)
            out[:b.shape[0]] = b.detach().to(dev, torch.float32)
            return out

        layers = []
        for L in self.transformer_layers:
            m = L.multihead_attention
            wi = m.in_proj_weight.detach().to(dev, torch.float32).view(3, H, hd, d)
            bi = m.in_proj_bias.detach().to(dev, torch.float32).view(3, H, hd)
            wqkv = torch.zeros(3, H, hp, dp, dtype=torch.float32, device=dev)
            wqkv[:, :, :hd, :d] = wi                                     # head h of q/k/v → 48-wide padded slot
            bqkv = torch.zeros(3, H, hp, dtype=torch.float32, device=dev)
            bqkv[:, :, :hd] = bi
            wo = torch.zeros(dp, H, hp, dtype=torch.float32, device=dev)
            wo[:d, :, :hd] = m.out_proj.weight.detach().to(dev, torch.float32).view(d, H, hd)
            layers.append(dict(
                wqkv=wqkv.view(3 * H * hp, dp).to(BF16), bqkv=bqkv.view(-1), wo=wo.view(dp, H * hp).to(BF16),
                bo=padb(m.out_proj.bias, dp), w1=padw(L.feedforward[0].weight, fp, dp), b1=padb(L.feedforward[0].bias, fp),
                w2=padw(L.feedforward[2].weight, dp, fp), b2=padb(L.feedforward[2].bias, dp),
                g1=L.norm1.weight.detach().to(dev, torch.float32).contiguous(), be1=L.norm1.bias.detach().to(dev, torch.float32).contiguous(),
                g2=L.norm2.weight.detach().to(dev, torch.float32).contiguous(), be2=L.norm2.bias.detach().to(dev, torch.float32).contiguous()))
        self._prep = dict(
            d=d, dp=dp, hd=hd, hp=hp, fp=fp, layers=layers,
            fc_w=padw(self.fc.weight, self.output_vector_dim, dp), fc_b=self.fc.bias.detach().to(dev, torch.float32).contiguous(),
            # the CoM blend's query and key projections (SimpleAttention, …:54-55,60-61) as ONE [2*dp, dp] GEMM: rows 0..dp-1 = Wq
            wqk=torch.cat([padw(self.com_attention.query_linear.weight, dp, dp), padw(self.com_attention.key_linear.weight, dp, dp)]),
            bqk=torch.cat([padb(self.com_attention.query_linear.bias, dp), padb(self.com_attention.key_linear.bias, dp)]),
            emb=self.temporal_embedding.embedding.weight.detach().to(dev, torch.float32).contiguous())
        self._prep_version = ver
        return self._prep

    def _act_buffers(self, M: int, dev, P) -> Dict[str, torch.Tensor]:
        if M not in self._bufs:
            if len(self._bufs) >= 4:
                self._bufs.pop(next(iter(self._bufs)))
            H, hp, dp, fp = self.num_heads, P["hp"], P["dp"], P["fp"]
            z = lambda c, dt=BF16: torch.zeros(M, c, dtype=dt, device=dev)
            self._bufs[M] = dict(xa=z(dp), xb=z(dp), qkv=z(3 * H * hp), o=z(H * hp), s=z(dp), x1=z(dp), h=z(fp), qk=z(2 * dp),
                                 v=z(8), blend=z(8), out=z(12, torch.float32))
            self._bufs[M]["s2"] = self._bufs[M]["s"]      # inference: both pre-LayerNorm sums share one buffer
        return self._bufs[M]

    # ---- host-fed stream (BASELINE configs[4]: the analysis pass over a long window stream) -----------------
    @torch.no_grad()
    def forward_stream(self, batches):
        """Generator over an iterable of host batches — input dicts (pinned CPU tensors keyed like ``forward``) or tensors
        pre-packed once with ``prepack`` (frame-major bf16: half the bytes over the host link) — yielding one output
        dict of pinned CPU tensors per batch, in order.  The H2D copies of batch i+1 run on a copy stream into the other
        half of a double-buffered staging area while batch i computes, and the D2H copies of batch i's three outputs
        run on a third stream, so neither PCIe direction leaves the GPU idle (analyze.py:112-156 moves one window at a
        time and synchronises on every ``.item()``).  Windows are independent: shard the stream across GPUs by window,
        no collective."""
        dev = next(self.parameters()).device
        main = torch.cuda.current_stream(dev)
        up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        slots = [dict(x=None, out=None, host=None, uploaded=torch.cuda.Event(), consumed=torch.cuda.Event(), done=torch.cuda.Event())
                 for _ in range(2)]

        def upload(batch, slot):
            with torch.cuda.stream(up):
                up.wait_event(slot["consumed"])
                if torch.is_tensor(batch):                  # pre-packed frame-major bf16 rows (``prepack``): half the link bytes
                    if not torch.is_tensor(slot["x"]) or slot["x"].shape != batch.shape:
                        slot["x"] = torch.empty(batch.shape, dtype=BF16, device=dev)
                    slot["x"].copy_(batch, non_blocking=True)
                else:
                    if not isinstance(slot["x"], dict) or any(slot["x"][k].shape != v.shape for k, v in batch.items()):
                        slot["x"] = {k: torch.empty(v.shape, dtype=torch.float32, device=dev) for k, v in batch.items()}
                    for k, v in batch.items():
                        slot["x"][k].copy_(v, non_blocking=True)
                slot["uploaded"].record(up)

        def compute(slot):
            main.wait_event(slot["uploaded"])
            out = self.forward(slot["x"])
            slot["consumed"].record(main)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(down):
                down.wait_event(ready)
                if slot["host"] is None or any(slot["host"][k].shape != v.shape for k, v in out.items()):
                    slot["host"] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out.items()}
                for k, v in out.items():
                    slot["host"][k].copy_(v, non_blocking=True)
                slot["done"].record(down)
            slot["out"] = out                          # keeps the device tensors alive until the copy has run

        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        for s_ in slots:
            s_["consumed"].record(main)
        upload(cur, slots[0])
        i, pending = 0, None
        while cur is not None:
            nxt = next(it, None)
            if nxt is not None:
                upload(nxt, slots[(i + 1) & 1])        # waits (on the copy stream) until batch i-1 has been consumed
            compute(slots[i & 1])                      # its pinned outputs replace batch i-2's, already handed out
            if pending is not None:
                pending["done"].synchronize()
                yield pending["host"]                  # valid until the next item is requested
            pending, cur, i = slots[i & 1], nxt, i + 1
        pending["done"].synchronize()
        yield pending["host"]

    # ---- pre-packed window streams (BASELINE configs[4]) -------------------------------------------------------------
    @staticmethod
    def prepack(x: Dict[str, torch.Tensor], pin: bool = True) -> torch.Tensor:
        """One-off HOST-side conversion of a batch of windows for a long analysis stream: the six (B, C, T) inputs of
        ``forward`` -> one frame-major bf16 tensor (B, T, round_up(3*dofs + 9, 8)), i.e. the cat(dim=1) + transpose(1, 2) of
        TransformerBaseline.py:108-116 done once where the data lives.  ``forward`` / ``forward_stream`` accept the result and
        produce bit-identical outputs (the device path rounds the same fp32 values to bf16 with the same RNE), while the
        host link carries 2 bytes per value instead of 4."""
        parts = [x[k].detach().to("cpu", torch.float32) for k in (InputDataKeys.POS, InputDataKeys.VEL, InputDataKeys.ACC,
                                                                     InputDataKeys.COM_POS, InputDataKeys.COM_VEL, InputDataKeys.COM_ACC)]
        rows = torch.cat(parts, dim=1).transpose(1, 2)                     # (B, T, C)
        B, T, C = rows.shape
        out = torch.zeros(B, T, ops.round_up(C, 8), dtype=BF16)
        out[:, :, :C] = rows.to(BF16)
        return out.pin_memory() if pin else out

    # ---- forward (TransformerBaseline.py:104-148) ---------------------------------------------------------
    def forward(self, x):
        """``x``: the reference's input dict of (B, C, T) tensors, or a tensor produced by ``prepack``."""
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise _lib.IbmError("TransformerBaseline runs only on a B200: move it to CUDA (no CPU fallback)")
        train = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if train:
            if self.training and any(m.p > 0.0 for m in self.modules() if isinstance(m, nn.Dropout)):
                raise NotImplementedError("TransformerBaseline training with dropout > 0 is not implemented on the B200 path")
            out, blend = _TransformerFunction.apply(self, x, *self.parameters())
        else:
            with torch.no_grad():
                out, blend, _ = self._forward_launches(x, save=False)
        batch_size, T = out.shape[0], out.shape[1]
        dt = self.fc.weight.dtype
        return {
            OutputDataKeys.CONTACT: torch.sigmoid(out[:, :, :2]).transpose(1, 2).to(dt),
            OutputDataKeys.COM_ACC: blend.view(batch_size, T, 8)[:, :, :3].transpose(1, 2).to(dt),
            OutputDataKeys.CONTACT_FORCES: out[:, :, 5:11].transpose(1, 2).to(dt),
        }

    def _forward_launches(self, x: Dict[str, torch.Tensor], save: bool):
        """Returns (out fp32 (B, T, 12), blend bf16 (B*T, 8), saved activations or None).  ``save`` keeps every layer's
        activations in fresh buffers for the backward; inference ping-pongs two."""
        dev = next(self.parameters()).device
        P = self._prepare(dev)
        d, dp, H, hp = P["d"], P["dp"], self.num_heads, P["hp"]
        packed = None
        if torch.is_tensor(x):                            # ``prepack`` output: (B, T, round_up(C, 8)) bf16 frame-major rows
            C = d - self.temporal_embedding_dim
            assert x.dim() == 3 and x.dtype == BF16 and x.size(2) == ops.round_up(C, 8), "expected a tensor made by prepack()"
            packed = x.to(dev, non_blocking=True).contiguous()
            batch_size, T = packed.size(0), packed.size(1)
        else:
            batch_size = x[InputDataKeys.POS].size(0)
            # (q, dq, ddq, com_pos, com_vel, com_acc) per timestep; inputs are (B, C, T) → (B, T, C)   (…:108-116)
            parts = [x[k].detach().to(dev, torch.float32).contiguous() for k in (InputDataKeys.POS, InputDataKeys.VEL, InputDataKeys.ACC,
                                                                                 InputDataKeys.COM_POS, InputDataKeys.COM_VEL, InputDataKeys.COM_ACC)]
            T = parts[0].size(2)
        assert T == self.window_size, "TemporalEmbedding.expand needs T == window_size (…:121-123)"
        M = batch_size * T
        b = self._train_buffers(M, dev, P) if save else self._act_buffers(M, dev, P)
        if packed is not None:
            # embedding columns appended and the 3 CoM-acceleration columns copied out as the blend's values: one kernel
            ops.expand_rows_bf16(packed.view(M, packed.size(2)), C, T, P["emb"], b["xa"], v_out=b["v"], v_col0=C - 3)
        else:
            # cat(dim=1) + transpose(1, 2) + temporal embedding CONCATENATED, not added (…:108-126): one kernel, (B, C, T) fp32
            # in, bf16 rows [B*T, 112] out
            ops.pack_channel_major(parts, T, P["emb"], b["xa"])
        cur, nxt = b["xa"], b["xb"]
        scale = 1.0 / math.sqrt(P["hd"])
        acts = []
        for li, L in enumerate(P["layers"]):
            a = b["layers"][li] if save else b
            if save:
                nxt = a["x2"]
            ops.gemm(cur, L["wqkv"], a["qkv"], M, 3 * H * hp, dp, bias=L["bqkv"])
            ops.attention_fwd_fused(a["qkv"], H * hp, a["o"], batch_size, T, H, hp, scale)
            ops.gemm(a["o"], L["wo"], a["s"], M, dp, H * hp, bias=L["bo"], aux=cur, aux_mode=1)
            ops.layernorm_fwd(a["s"], a["x1"], L["g1"], L["be1"], M, d, mean=a.get("mean1"), rstd=a.get("rstd1"))
            ops.gemm(a["x1"], L["w1"], a["h"], M, P["fp"], dp, bias=L["b1"], act="relu")
            ops.gemm(a["h"], L["w2"], a["s2"], M, dp, P["fp"], bias=L["b2"], aux=a["x1"], aux_mode=1)
            ops.layernorm_fwd(a["s2"], nxt, L["g2"], L["be2"], M, d, mean=a.get("mean2"), rstd=a.get("rstd2"))
            if save:
                acts.append(dict(a, x_in=cur))
                cur = nxt
            else:
                cur, nxt = nxt, cur
        ops.gemm(cur, P["fc_w"], b["out"], M, self.output_vector_dim, dp, bias=P["fc_b"])
        # CoM acceleration as an (unscaled) attention blend over the input CoM accelerations (…:51-70, 135-137)
        ops.gemm(cur, P["wqk"], b["qk"], M, 2 * dp, dp, bias=P["bqk"])
        if packed is None:
            ops.pack_channel_major(parts[5:6], T, None, b["v"])        # CoM accelerations as the (3 -> 8)-wide values of the blend
        ops.attention_fwd(b["qk"][:, :dp], b["qk"][:, dp:], b["v"], b["blend"], batch_size, T, 1, dp, 8, 1.0)
        saved = dict(acts=acts, last=cur, qk=b["qk"], v=b["v"], blend=b["blend"], M=M, B=batch_size, T=T, P=P) if save else None
        return b["out"].view(batch_size, T, 12), b["blend"], saved

    def _train_buffers(self, M: int, dev, P):
        """Fresh activation buffers for one training forward (kept alive by the autograd node until its backward)."""
        H, hp, dp, fp = self.num_heads, P["hp"], P["dp"], P["fp"]
        z = lambda c, dt=BF16: torch.zeros(M, c, dtype=dt, device=dev)
        f = lambda: torch.empty(M, dtype=torch.float32, device=dev)
        layers = [dict(qkv=z(3 * H * hp), o=z(H * hp), s=z(dp), x1=z(dp), h=z(fp), s2=z(dp), x2=z(dp), mean1=f(), rstd1=f(), mean2=f(),
                       rstd2=f()) for _ in range(self.num_layers)]
        return dict(xa=z(dp), xb=None, layers=layers, qk=z(2 * dp), v=z(8), blend=z(8), out=z(12, torch.float32))

    # ---- backward (autograd of TransformerBaseline.py:104-148 through the same kernels' transposes) -------------
    def _backward_launches(self, sv, g_out: torch.Tensor, g_blend: torch.Tensor):
        P, M, B, T = sv["P"], sv["M"], sv["B"], sv["T"]
        d, dp, H, hp, hd, fp = P["d"], P["dp"], self.num_heads, P["hp"], P["hd"], P["fp"]
        dev = sv["last"].device
        zb = lambda c: torch.zeros(M, c, dtype=BF16, device=dev)
        zf = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        scale = 1.0 / math.sqrt(hd)
        # --- heads: fc (d -> 11) and the CoM blend's query / key projections, all reading the last layer's output ---
        dout = zb(16)
        ops.cast_pad(g_out.contiguous().view(M, 12), dout, M, 11)
        g_fc_w, g_fc_b = zf(self.output_vector_dim, dp), zf(16)
        ops.gemm(dout, sv["last"], g_fc_w, self.output_vector_dim, dp, M, a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(dout, M, self.output_vector_dim, g_fc_b)
        dxa, dxb = zb(dp), zb(dp)
        ops.gemm(dout, P["fc_w"], dxa, M, dp, self.output_vector_dim, b_mn=True)
        dblend, dqk = zb(8), zb(2 * dp)
        ops.cast_pad(g_blend.contiguous().view(M, 8), dblend, M, 3)
        ops.attention_bwd_long(sv["qk"][:, :dp], sv["qk"][:, dp:], sv["v"], sv["blend"], dblend, dqk[:, :dp], dqk[:, dp:], None,
                               B, T, 1, dp, 8, 1.0)
        g_wqk, g_bqk = zf(2 * dp, dp), zf(2 * dp)
        ops.gemm(dqk, sv["last"], g_wqk, 2 * dp, dp, M, a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(dqk, M, 2 * dp, g_bqk)
        ops.gemm(dqk, P["wqk"], dxb, M, dp, 2 * dp, b_mn=True, aux=dxa, aux_mode=1)
        dx, other = dxb, dxa
        # --- layers, last to first (TransformerBaseline.py:24-38 reversed) ---
        ds, dh, dx1, do, dqkv = zb(dp), zb(fp), zb(dp), zb(H * hp), zb(3 * H * hp)
        lg = []
        for li in range(self.num_layers - 1, -1, -1):
            L, a = P["layers"][li], sv["acts"][li]
            g = dict(wqkv=zf(3 * H * hp, dp), bqkv=zf(3 * H * hp), wo=zf(dp, H * hp), bo=zf(dp), w1=zf(fp, dp), b1=zf(fp),
                     w2=zf(dp, fp), b2=zf(dp), g1=zf(dp), be1=zf(dp), g2=zf(dp), be2=zf(dp))
            ops.layernorm_bwd(dx, a["s2"], L["g2"], a["mean2"], a["rstd2"], M, d, ds, g["g2"], g["be2"], g["b2"])
            ops.gemm(ds, a["h"], g["w2"], dp, fp, M, a_mn=True, b_mn=True, accumulate=True)
            ops.gemm(ds, L["w2"], dh, M, fp, dp, b_mn=True, act="relu", aux=a["h"], aux_mode=2, colsum=g["b1"])
            ops.gemm(dh, a["x1"], g["w1"], fp, dp, M, a_mn=True, b_mn=True, accumulate=True)
            ops.gemm(dh, L["w1"], dx1, M, dp, fp, b_mn=True, aux=ds, aux_mode=1)
            ops.layernorm_bwd(dx1, a["s"], L["g1"], a["mean1"], a["rstd1"], M, d, ds, g["g1"], g["be1"], g["bo"])
            ops.gemm(ds, a["o"], g["wo"], dp, H * hp, M, a_mn=True, b_mn=True, accumulate=True)
            ops.gemm(ds, L["wo"], do, M, H * hp, dp, b_mn=True)
            qkv, w = a["qkv"], H * hp
            ops.attention_bwd_long(qkv[:, :w], qkv[:, w:2 * w], qkv[:, 2 * w:], a["o"], do, dqkv[:, :w], dqkv[:, w:2 * w],
                                   dqkv[:, 2 * w:], B, T, H, hp, hp, scale, dbq=g["bqkv"][:w], dbk=g["bqkv"][w:2 * w],
                                   dbv=g["bqkv"][2 * w:])
            ops.gemm(dqkv, a["x_in"], g["wqkv"], 3 * w, dp, M, a_mn=True, b_mn=True, accumulate=True)
            ops.gemm(dqkv, L["wqkv"], other, M, dp, 3 * w, b_mn=True, aux=ds, aux_mode=1)
            dx, other = other, dx
            lg.append(g)
        lg.reverse()
        # --- temporal embedding: rows of the table were concatenated as input columns [d - E, d) of every window (…:121-126) ---
        E = self.temporal_embedding_dim
        g_in = zf(T * dp)
        ops.colsum(dx.view(B, T * dp), B, T * dp, g_in)
        g_emb = torch.zeros_like(self.temporal_embedding.embedding.weight, dtype=torch.float32)
        g_emb[:T] = g_in.view(T, dp)[:, d - E:d]
        # --- un-pad into the reference's parameter shapes, in self.parameters() order ---
        by_name = {"temporal_embedding.embedding.weight": g_emb, "fc.weight": g_fc_w[:, :d], "fc.bias": g_fc_b[:self.output_vector_dim],
                   "com_attention.query_linear.weight": g_wqk[:d, :d], "com_attention.query_linear.bias": g_bqk[:d],
                   "com_attention.key_linear.weight": g_wqk[dp:dp + d, :d], "com_attention.key_linear.bias": g_bqk[dp:dp + d]}
        for li, g in enumerate(lg):
            pre = f"transformer_layers.{li}."
            by_name[pre + "multihead_attention.in_proj_weight"] = g["wqkv"].view(3, H, hp, dp)[:, :, :hd, :d].reshape(3 * d, d)
            by_name[pre + "multihead_attention.in_proj_bias"] = g["bqkv"].view(3, H, hp)[:, :, :hd].reshape(3 * d)
            by_name[pre + "multihead_attention.out_proj.weight"] = g["wo"].view(dp, H, hp)[:d, :, :hd].reshape(d, d)
            by_name[pre + "multihead_attention.out_proj.bias"] = g["bo"][:d]
            by_name[pre + "feedforward.0.weight"] = g["w1"][:self.dim_feedforward, :d]
            by_name[pre + "feedforward.0.bias"] = g["b1"][:self.dim_feedforward]
            by_name[pre + "feedforward.2.weight"] = g["w2"][:d, :self.dim_feedforward]
            by_name[pre + "feedforward.2.bias"] = g["b2"][:d]
            by_name[pre + "norm1.weight"], by_name[pre + "norm1.bias"] = g["g1"][:d], g["be1"][:d]
            by_name[pre + "norm2.weight"], by_name[pre + "norm2.bias"] = g["g2"][:d], g["be2"][:d]
        return [by_name[n].to(p.dtype).reshape(p.shape) if p.requires_grad else None for n, p in self.named_parameters()]
