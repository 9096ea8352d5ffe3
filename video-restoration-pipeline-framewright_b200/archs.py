"""Architecture descriptors and synthetic weights for the Real-ESRGAN upscaling hot path.

The reference (`/root/reference/src/framewright/processors/pytorch_realesrgan.py:103-129`) maps five
model names to `RRDBNet(...)` constructor calls.  Upstream Real-ESRGAN ships two of those names
(`realesr-general-x4v3`, `realesr-animevideov3`) as `SRVGGNetCompact` checkpoints, so this table keeps
the reference's *names and scales* and takes the *architectures* from upstream (SURVEY.md finding 3,
Appendix A).  State-dict key names are the upstream ones so real `.pth` files drop in.

Nothing here touches the GPU; descriptors are plain data consumed by `engine.py` (weight packing)
and by the CPU oracle under `oracle/` (test infrastructure only).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch


@dataclass(frozen=True)
class ArchDesc:
    """One network architecture on the hot path."""

    kind: str            # "rrdb" | "srvgg"
    scale: int           # network scale (netscale): 4 or 2
    num_in_ch: int = 3
    num_out_ch: int = 3
    num_feat: int = 64
    num_block: int = 23  # RRDB blocks (rrdb) / body convs (srvgg: num_conv)
    num_grow_ch: int = 32

    @property
    def first_in_ch(self) -> int:
        """Input channels seen by the first conv (pixel-unshuffle folds s*s into channels)."""
        if self.kind == "rrdb" and self.scale == 2:
            return self.num_in_ch * 4
        if self.kind == "rrdb" and self.scale == 1:
            return self.num_in_ch * 16
        return self.num_in_ch

    def macs_per_input_pixel(self) -> int:
        """Algorithmic multiply-accumulates per network-input pixel (true channel counts).

        For the x2 RRDBNet the unit is one *unshuffled* pixel.
        """
        nf, gc = self.num_feat, self.num_grow_ch
        if self.kind == "rrdb":
            rdb = 9 * (nf * gc + (nf + gc) * gc + (nf + 2 * gc) * gc + (nf + 3 * gc) * gc + (nf + 4 * gc) * nf)
            total = 9 * self.first_in_ch * nf + 3 * self.num_block * rdb + 9 * nf * nf
            total += 4 * 9 * nf * nf            # conv_up1 at 2x
            total += 16 * 9 * nf * nf * 2       # conv_up2 + conv_hr at 4x
            total += 16 * 9 * nf * self.num_out_ch
            return total
        up2 = self.scale * self.scale
        return 9 * self.num_in_ch * nf + self.num_block * 9 * nf * nf + 9 * nf * self.num_out_ch * up2


# model name -> architecture.  Names/scales: pytorch_realesrgan.py:103-129 (reference);
# architectures: upstream Real-ESRGAN inference table.
MODEL_ARCHS: Dict[str, ArchDesc] = {
    "RealESRGAN_x4plus": ArchDesc("rrdb", 4, num_block=23),
    "RealESRGAN_x4plus_anime_6B": ArchDesc("rrdb", 4, num_block=6),
    "RealESRGAN_x2plus": ArchDesc("rrdb", 2, num_block=23),
    "realesr-animevideov3": ArchDesc("srvgg", 4, num_block=16),
    "realesr-general-x4v3": ArchDesc("srvgg", 4, num_block=32),
}


def conv_layers(arch: ArchDesc) -> List[Tuple[str, int, int]]:
    """(state-dict prefix, Cin, Cout) for every 3x3 conv, in execution order."""
    nf, gc = arch.num_feat, arch.num_grow_ch
    out: List[Tuple[str, int, int]] = []
    if arch.kind == "rrdb":
        out.append(("conv_first", arch.first_in_ch, nf))
        for b in range(arch.num_block):
            for r in (1, 2, 3):
                for k in range(1, 5):
                    out.append((f"body.{b}.rdb{r}.conv{k}", nf + (k - 1) * gc, gc))
                out.append((f"body.{b}.rdb{r}.conv5", nf + 4 * gc, nf))
        out.append(("conv_body", nf, nf))
        out.append(("conv_up1", nf, nf))
        out.append(("conv_up2", nf, nf))
        out.append(("conv_hr", nf, nf))
        out.append(("conv_last", nf, arch.num_out_ch))
    else:
        out.append(("body.0", arch.num_in_ch, nf))
        for i in range(arch.num_block):
            out.append((f"body.{2 * (i + 1)}", nf, nf))
        out.append((f"body.{2 * (arch.num_block + 1)}", nf, arch.num_out_ch * arch.scale * arch.scale))
    return out


def prelu_layers(arch: ArchDesc) -> List[str]:
    """State-dict keys of the per-channel PReLU slopes (srvgg only), in execution order."""
    if arch.kind != "srvgg":
        return []
    return [f"body.{2 * i + 1}.weight" for i in range(arch.num_block + 1)]


def make_synthetic_state_dict(model_name: str, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic random-init weights of the named architecture (fp32, upstream key names).

    Recipe (committed so goldens are reproducible; SURVEY.md §7 hard part 3):
      * RDB convs: N(0, 2/fan_in) scaled by 0.1, zero bias   (upstream `default_init_weights(scale=0.1)`).
      * every other conv: U(-b, b) with b = 1/sqrt(fan_in) for weight and bias (torch Conv2d default);
        SRVGG body convs: N(0, 2 / (1.0625 fan_in)) weights (variance-preserving through PReLU).
      * PReLU slopes: 0.25 + U(-0.1, 0.1) per channel (torch default is a constant 0.25; jitter makes
        the per-channel path observable).
      * the last conv is re-centred so the pre-clamp output sits inside [0, 1] instead of saturating:
        weight scaled by LAST_GAIN[(kind, scale, num_block)], bias set to 0.5 (RRDBNet) / 0.0 (SRVGG, whose nearest-upsampled
        input residual already carries the image).
    Sampling order is the execution order of `conv_layers`, weight then bias, from one
    `torch.Generator(seed)`.
    """
    arch = MODEL_ARCHS[model_name]
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    sd: Dict[str, torch.Tensor] = {}
    layers = conv_layers(arch)
    for idx, (name, cin, cout) in enumerate(layers):
        fan_in = cin * 9
        is_last = idx == len(layers) - 1
        if ".rdb" in name:
            w = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / fan_in) * 0.1
            b = torch.zeros(cout)
        elif arch.kind == "srvgg" and not is_last:
            # variance-preserving through PReLU(0.25): N(0, 2 / ((1 + 0.25^2) fan_in)); otherwise the
            # 32-conv body decays to nothing and the output is just the nearest-upsampled input
            bound = 1.0 / math.sqrt(fan_in)
            w = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (1.0625 * fan_in))
            b = (torch.rand(cout, generator=g) * 2 - 1) * bound
        else:
            bound = 1.0 / math.sqrt(fan_in)
            w = (torch.rand(cout, cin, 3, 3, generator=g) * 2 - 1) * bound
            b = (torch.rand(cout, generator=g) * 2 - 1) * bound
        if is_last:
            w = w * LAST_GAIN[(arch.kind, arch.scale, arch.num_block)]
            b = torch.full((cout,), 0.5 if arch.kind == "rrdb" else 0.0)
        sd[name + ".weight"] = w.contiguous()
        sd[name + ".bias"] = b.contiguous()
    for key in prelu_layers(arch):
        sd[key] = 0.25 + (torch.rand(arch.num_feat, generator=g) * 2 - 1) * 0.1
    return sd


# Output-range gains for the synthetic recipe, chosen once with the fp32 oracle on uniform-noise
# frames so the pre-clamp output has sigma ~ 0.15 (few saturated pixels); see oracle/gen_golden.py.
LAST_GAIN = {("rrdb", 4, 23): 0.45, ("rrdb", 4, 6): 4.0, ("rrdb", 2, 23): 0.25, ("srvgg", 4, 16): 0.35, ("srvgg", 4, 32): 0.5}
