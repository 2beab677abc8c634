"""Launch plan for the Groundlink CNN (/root/reference/src/models/Groundlink.py:20-77, 105-156):
4 x [Conv1d(k=7, padding=3, padding_mode="replicate") + ELU] over time, then per-frame
Linear(256,256)+ELU x2 and Linear(256,30, bias=False), with Dropout(0.2) before each Linear in training.

The temporal convolution is an implicit GEMM on tcgen05 (``ibm_gemm_bf16(taps=7)``): activations live in a
*padded row layout* — window b owns rows [b*Tp, (b+1)*Tp), Tp = T + 6, frame t at row b*Tp + 3 + t, the 3+3
pad rows holding copies of the first / last frame (replicate padding) — so tap j of output row r reads row
r + j and the whole convolution is one GEMM with K = 7 * C_in.  The GEMM's TMA store is pointed 3 rows down,
which drops each frame's output straight into its slot of the next layer's padded buffer; a tiny kernel then
refreshes the 6 pad rows per window.  Backward: dgrad is the same implicit GEMM with flipped taps and
transposed weights, wgrad is 7 MN-major split-K GEMMs (one per tap) accumulating into an fp32 scratch in GEMM
layout that is folded back into the (C_out, C_in, 7) parameter gradient.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import ops
from .engine import _Buffers, _r8
from .params import ParamArena

BF16, F32 = torch.bfloat16, torch.float32
PAD, KT = 3, 7


class GroundlinkEngine:
    CNN_DROPOUT_SEED = 0x636e6e64

    def __init__(self, arena: ParamArena, c_in: int, features: List[int], fc_dropout: float, cnn_dropout: float = 0.0):
        self.arena = arena
        self.ch = [c_in] + list(features)                 # [177, 128, 128, 256, 256]
        self.ld = [_r8(c) for c in self.ch]               # row pitch of each layer's activation buffer
        self.cin_pad = [ops.round_up(self.ld[i], 64) for i in range(4)]       # per-tap K padded to whole 64-wide k blocks
        self.cout_pad = [ops.round_up(self.ch[i + 1], 64) for i in range(4)]
        self.conv_pos = (1, 4, 7, 10)                     # nn.Sequential positions
        self.fc_pos = (2, 5, 8)
        self.fc_dropout = fc_dropout
        self.cnn_dropout = float(cnn_dropout)             # nn.Dropout before every Conv1d (Groundlink.py:41, default 0.0)
        self.buf = _Buffers(arena.device)
        self._wver = None
        self._w: Dict[str, torch.Tensor] = {}
        self.step = 0
        self.bucket_hook = None

    # ---- weights in GEMM layouts (refreshed when the fp32 masters change) ---------------------------------
    def _weights(self) -> Dict[str, torch.Tensor]:
        ver = self.arena._version_sum()
        if self._wver == ver and self._w:
            return self._w
        dev = self.arena.device
        for i, pos in enumerate(self.conv_pos):
            cout, cin = self.ch[i + 1], self.ch[i]
            w = self.arena.master_of(f"cnn.{pos}.weight", (cout, cin, KT))
            fw = self._w.get(f"f{i}")
            if fw is None:
                fw = self._w[f"f{i}"] = torch.empty(cout, KT * self.cin_pad[i], dtype=BF16, device=dev)
                self._w[f"d{i}"] = torch.empty(cin, KT * self.cout_pad[i], dtype=BF16, device=dev)
            ops.conv_weight_to_gemm(w, fw, self.cin_pad[i])
            ops.conv_weight_to_dgrad(w, self._w[f"d{i}"], self.cout_pad[i])
        self._wver = ver
        return self._w

    def weights_changed(self) -> None:
        """The fused optimizer writes the fp32 masters through raw pointers (no torch version bump): drop the cache."""
        self._wver = None

    # ---- buffers ---------------------------------------------------------------------------------------------
    def _state(self, B: int, T: int):
        st = self.buf.get((B, T))
        Tp = T + 2 * PAD
        Mp = B * Tp
        if "x0" not in st:
            slack = 8                                      # rows of zeros before/after: tap shifts (+6) and the +3 store shift
            for i in range(5):
                full = torch.zeros(Mp + 2 * slack, self.ld[i], dtype=BF16, device=self.arena.device)
                st[f"x{i}_full"] = full
                st[f"x{i}"] = full[slack:slack + Mp]
                gfull = torch.zeros(Mp + 2 * slack, self.ld[i], dtype=BF16, device=self.arena.device)
                st[f"g{i}_full"] = gfull
                st[f"g{i}"] = gfull[slack:slack + Mp]
            for k in ("h1", "h2", "h1d", "h2d", "y4d", "dh1", "dh2", "dtmp"):
                st[k] = torch.zeros(Mp, 256, dtype=BF16, device=self.arena.device)
            st["out"] = torch.zeros(Mp, 32, dtype=F32, device=self.arena.device)
            st["dout"] = torch.zeros(Mp, 32, dtype=BF16, device=self.arena.device)
            st["slack"] = slack
        return st, Tp, Mp

    def input_rows(self, B: int, T: int) -> Tuple[torch.Tensor, int, int, int]:
        """(buffer, frame_stride, win_extra, col0) for the packers: frame t of window b → row b*Tp + 3 + t."""
        st, Tp, Mp = self._state(B, T)
        ld = self.ld[0]
        return st["x0"], ld, 2 * PAD * ld, PAD * ld

    def out_view(self, B: int, T: int) -> torch.Tensor:
        st, Tp, Mp = self._state(B, T)
        return st["out"].view(B, Tp, 32)[:, PAD:PAD + T, :]

    def dout_view(self, B: int, T: int) -> torch.Tensor:
        st, Tp, Mp = self._state(B, T)
        return st["dout"].view(B, Tp, 32)[:, PAD:PAD + T, :]

    # ---- forward ---------------------------------------------------------------------------------------------
    def forward(self, B: int, T: int, train: bool) -> torch.Tensor:
        """Consumes input_rows (frames written at rows b*Tp+3+t); returns the (B, T, 32) fp32 output view."""
        A = self.arena
        W = self._weights()
        st, Tp, Mp = self._state(B, T)
        self.step += 1
        cdrop = train and self.cnn_dropout > 0.0
        st["cnn_dropped"] = cdrop
        ops.replicate_pad_rows(st["x0"], B, T, PAD, self.ld[0])
        for i, pos in enumerate(self.conv_pos):
            x, y_full = st[f"x{i}"], st[f"x{i + 1}_full"]
            slack = st["slack"]
            if cdrop:
                # Dropout on the conv INPUT, before the (replicate) padding the conv applies itself: mask the whole buffer with
                # Philox(seed, 4*step + i) (zeros stay zeros), then let the pad rows replicate the DROPPED edge frames
                xd_full = st.get(f"xd{i}_full")
                if xd_full is None:
                    xd_full = st[f"xd{i}_full"] = torch.zeros_like(st[f"x{i}_full"])
                    st[f"xd{i}"] = xd_full[slack:slack + Mp]
                ops.dropout(st[f"x{i}_full"], xd_full, self.cnn_dropout, self.CNN_DROPOUT_SEED, 4 * self.step + i)
                ops.replicate_pad_rows(st[f"xd{i}"], B, T, PAD, self.ld[i])
                x = st[f"xd{i}"]
            y_shift = y_full[slack + PAD: slack + PAD + Mp]             # the store lands 3 rows down: frame slots of layer i+1
            ops.gemm(x, W[f"f{i}"], y_shift, Mp, self.ch[i + 1], KT * self.ld[i], lda=self.ld[i], bias=A.master_of(f"cnn.{pos}.bias"),
                     act="elu", taps=KT)
            ops.replicate_pad_rows(st[f"x{i + 1}"], B, T, PAD, self.ld[i + 1])
        drop = train and self.fc_dropout > 0.0
        y4 = st["x4"]
        a = y4
        if drop:
            ops.dropout(y4, st["y4d"], self.fc_dropout, 0x6c696e6b, 3 * self.step)
            a = st["y4d"]
        ops.gemm(a, A.shadow_of("fc.2.weight", (256, 256)), st["h1"], Mp, 256, 256, bias=A.master_of("fc.2.bias"), act="elu")
        a = st["h1"]
        if drop:
            ops.dropout(st["h1"], st["h1d"], self.fc_dropout, 0x6c696e6b, 3 * self.step + 1)
            a = st["h1d"]
        ops.gemm(a, A.shadow_of("fc.5.weight", (256, 256)), st["h2"], Mp, 256, 256, bias=A.master_of("fc.5.bias"), act="elu")
        a = st["h2"]
        if drop:
            ops.dropout(st["h2"], st["h2d"], self.fc_dropout, 0x6c696e6b, 3 * self.step + 2)
            a = st["h2d"]
        ops.gemm(a, A.shadow_of("fc.8.weight", (30, 256)), st["out"], Mp, 30, 256)
        st["dropped"] = drop
        return self.out_view(B, T)

    # ---- backward ---------------------------------------------------------------------------------------------
    def backward(self, B: int, T: int) -> None:
        """Consumes dout (bf16, written through dout_view; all other rows zero); accumulates parameter grads."""
        A = self.arena
        W = self._weights()
        g = A.grad_of
        st, Tp, Mp = self._state(B, T)
        drop = st.get("dropped", False)
        cdrop = st.get("cnn_dropped", False)
        dout = st["dout"]
        p, s = self.fc_dropout, self.step
        # fc.8 (no bias)
        x3 = st["h2d"] if drop else st["h2"]
        ops.gemm(dout, x3, g("fc.8.weight", (30, 256)), 30, 256, Mp, a_mn=True, b_mn=True, accumulate=True)
        ops.gemm(dout, A.shadow_of("fc.8.weight", (30, 256)), st["dh2"], Mp, 256, 30, b_mn=True, act="elu", aux=st["h2"], aux_mode=2)
        if drop:
            ops.dropout(st["dh2"], st["dh2"], p, 0x6c696e6b, 3 * s + 2)
        # fc.5
        x2 = st["h1d"] if drop else st["h1"]
        ops.gemm(st["dh2"], x2, g("fc.5.weight", (256, 256)), 256, 256, Mp, a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(st["dh2"], Mp, 256, g("fc.5.bias"))
        ops.gemm(st["dh2"], A.shadow_of("fc.5.weight", (256, 256)), st["dh1"], Mp, 256, 256, b_mn=True, act="elu", aux=st["h1"], aux_mode=2)
        if drop:
            ops.dropout(st["dh1"], st["dh1"], p, 0x6c696e6b, 3 * s + 1)
        # fc.2
        x1 = st["y4d"] if drop else st["x4"]
        ops.gemm(st["dh1"], x1, g("fc.2.weight", (256, 256)), 256, 256, Mp, a_mn=True, b_mn=True, accumulate=True)
        ops.colsum(st["dh1"], Mp, 256, g("fc.2.bias"))
        # gradient w.r.t. the last conv layer's (padded) output, through its ELU
        ops.gemm(st["dh1"], A.shadow_of("fc.2.weight", (256, 256)), st["g4"], Mp, 256, 256, b_mn=True, act="elu", aux=st["x4"], aux_mode=2)
        if drop:
            ops.dropout(st["g4"], st["g4"], p, 0x6c696e6b, 3 * s)
        if self.bucket_hook is not None:
            self.bucket_hook(4)
        slack = st["slack"]
        for i in range(3, -1, -1):
            pos = self.conv_pos[i]
            cout, cin = self.ch[i + 1], self.ch[i]
            G = st[f"g{i + 1}"]                              # d loss / d (ELU output of conv i), padded layout, already times ELU'
            ops.fold_pad_rows(G, B, T, PAD, self.ld[i + 1])  # replicate-pad adjoint; pad rows become zero
            if cdrop and i + 1 <= 3:
                # adjoint of the dropout in front of conv i+1: the same Philox mask over the identically shaped buffer (the ELU
                # derivative fused into the producing dgrad commutes with it, and with the fold because pads replicate x)
                Gfull = st[f"g{i + 1}_full"]
                ops.dropout(Gfull, Gfull, self.cnn_dropout, self.CNN_DROPOUT_SEED, 4 * s + i + 1)
            ops.colsum(G, Mp, cout, g(f"cnn.{pos}.bias"))
            # wgrad: dW_j[co, ci] = sum_r G[r + 3, co] * Xp[r + j, ci]  (7 split-K MN-major GEMMs into GEMM-layout scratch)
            wg = self.buf.tensor(st, f"wg{i}", (cout, KT * self.cin_pad[i]), F32)
            wg.zero_()
            Gf, Xf = st[f"g{i + 1}_full"], st[f"xd{i}_full" if cdrop else f"x{i}_full"]
            for j in range(KT):
                ops.gemm(Gf[slack + PAD:], Xf[slack + j:], wg[:, j * self.cin_pad[i]:], cout, cin, Mp, lda=self.ld[i + 1], ldb=self.ld[i],
                         ldd=KT * self.cin_pad[i], a_mn=True, b_mn=True, accumulate=True)
            ops.conv_wgrad_from_gemm(wg, g(f"cnn.{pos}.weight", (cout, cin, KT)), self.cin_pad[i], accumulate=True)
            if i > 0:
                # dgrad: dXp[s] = sum_j G[s - j + 3] W_j^T  → taps GEMM over rows shifted up by 3, flipped/transposed weights,
                # times ELU'(x_i) in the epilogue (x_i = padded output of conv i-1, pad rows included)
                ops.gemm(Gf[slack - PAD:], W[f"d{i}"], st[f"g{i}"], Mp, cin, KT * self.ld[i + 1], lda=self.ld[i + 1], act="elu",
                         aux=st[f"x{i}"], aux_mode=2, taps=KT)
            if self.bucket_hook is not None:
                self.bucket_hook(i)
