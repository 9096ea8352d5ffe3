"""CPU emulation of the GPU path's rounding points (test infrastructure; torch CPU fp32).

Mirrors csrc/b200sr.cu::run_region for the RRDBNet: 16-bit activations/weights at exactly the places
the engine stores them, fp32 accumulation, residual stream stored as a bf16 hi + e5m2 lo pair.  Used to predict the parity gate
without a GPU and to tell rounding effects from bugs when a GPU result differs from the oracle.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F


def _q(t: torch.Tensor, dtype) -> torch.Tensor:
    return t if dtype is None else t.to(dtype).float()


def emulate_rrdb(sd: Dict[str, torch.Tensor], img_bgr_u8: np.ndarray, scale: int = 4, num_block: int = 23,
                 act_dtype=torch.bfloat16, w_dtype=torch.bfloat16, trunk_copy_dtype="same", tail_dtype="same",
                 first_dtype="same", tail_w_dtype="same", trunk_mode="hybrid") -> np.ndarray:
    trunk_copy_dtype = act_dtype if trunk_copy_dtype == "same" else trunk_copy_dtype
    tail_dtype = act_dtype if tail_dtype == "same" else tail_dtype
    first_dtype = act_dtype if first_dtype == "same" else first_dtype
    tail_w_dtype = w_dtype if tail_w_dtype == "same" else tail_w_dtype

    def tq(v, rrdb_end=True):
        """Residual-stream storage: "hybrid" (the engine, conv3x3_tc.cuh TrunkLo) = bf16 hi (the conv-input copy) +
        e5m2 lo at RRDB boundaries, hi alone inside an RRDB, and the lo part used by the RRDB-level skip only (option
        trunk_lo = 0); "hybrid1" = lo also in the first RDB's residual add (trunk_lo = 1); "hilo" = the pair after
        every RDB (trunk_lo = 2); "f32"; or "bf16" alone."""
        if trunk_mode == "f32":
            return v
        hi = _q(v, torch.bfloat16)
        if trunk_mode == "bf16" or (trunk_mode in ("hybrid", "hybrid1") and not rrdb_end):
            return hi
        return hi + (v - hi).to(torch.float8_e5m2).float()

    def conv(x, name, wq=True, wd="trunk"):
        w = sd[name + ".weight"].float()
        if wq:
            w = _q(w, w_dtype if wd == "trunk" else tail_w_dtype)
        return F.conv2d(x, w, sd[name + ".bias"].float(), padding=1)

    x = torch.from_numpy(img_bgr_u8[:, :, ::-1].copy().transpose(2, 0, 1)).float().unsqueeze(0) / 255.0
    if scale == 2:
        from oracle.oracle import pixel_unshuffle
        x = pixel_unshuffle(x, 2)
    feat = conv(x, "conv_first", wq=False)           # fp32 CUDA-core kernel
    trunk = tq(feat.clone())
    xq = _q(feat, first_dtype)
    for b in range(num_block):
        x0 = trunk.clone()
        for r in (1, 2, 3):
            p = f"body.{b}.rdb{r}."
            cur = _q(trunk, trunk_copy_dtype) if not (b == 0 and r == 1) else xq
            cat = cur
            for k in range(1, 5):
                xk = F.leaky_relu(conv(cat, p + f"conv{k}"), 0.2)
                cat = torch.cat((cat, _q(xk, act_dtype)), 1)
            x5 = conv(cat, p + "conv5")
            v = x5 * 0.2 + (_q(trunk, torch.bfloat16) if (trunk_mode == "hybrid" and r == 1) else trunk)
            # RRDB-level skip fused in the rdb3 epilogue: x <- ((acc+b)*0.2 + x)*0.2 + x0
            trunk = tq(v * 0.2 + x0) if r == 3 else tq(v, rrdb_end=False)
    body = conv(_q(trunk, trunk_copy_dtype), "conv_body")
    feat = _q(feat + body, tail_dtype)
    feat = _q(F.leaky_relu(conv(F.interpolate(feat, scale_factor=2, mode="nearest"), "conv_up1", wd="tail"), 0.2), tail_dtype)
    feat = _q(F.leaky_relu(conv(F.interpolate(feat, scale_factor=2, mode="nearest"), "conv_up2", wd="tail"), 0.2), tail_dtype)
    feat = _q(F.leaky_relu(conv(feat, "conv_hr", wd="tail"), 0.2), tail_dtype)
    out = conv(feat, "conv_last", wd="tail")
    out = out.squeeze(0).clamp_(0, 1).numpy()
    out = np.transpose(out[[2, 1, 0]], (1, 2, 0))
    return (out * 255.0).round().astype(np.uint8)


def emulate_srvgg(sd: Dict[str, torch.Tensor], img_bgr_u8: np.ndarray, num_conv: int = 32, upscale: int = 4,
                  act_dtype=torch.bfloat16, w_dtype=torch.bfloat16) -> np.ndarray:
    x = torch.from_numpy(img_bgr_u8[:, :, ::-1].copy().transpose(2, 0, 1)).float().unsqueeze(0) / 255.0
    out = F.conv2d(x, sd["body.0.weight"].float(), sd["body.0.bias"].float(), padding=1)
    out = _q(F.prelu(out, sd["body.1.weight"].float()), act_dtype)
    for i in range(num_conv):
        k = 2 * (i + 1)
        out = F.conv2d(out, _q(sd[f"body.{k}.weight"].float(), w_dtype), sd[f"body.{k}.bias"].float(), padding=1)
        out = _q(F.prelu(out, sd[f"body.{k + 1}.weight"].float()), act_dtype)
    k = 2 * (num_conv + 1)
    out = F.conv2d(out, _q(sd[f"body.{k}.weight"].float(), w_dtype), sd[f"body.{k}.bias"].float(), padding=1)
    out = F.pixel_shuffle(out, upscale) + F.interpolate(x, scale_factor=upscale, mode="nearest")
    out = out.squeeze(0).clamp_(0, 1).numpy()
    out = np.transpose(out[[2, 1, 0]], (1, 2, 0))
    return (out * 255.0).round().astype(np.uint8)
