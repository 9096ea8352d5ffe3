"""CPU oracle for the Real-ESRGAN upscaling hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module; the product package never does (and fails loudly without its CUDA library).

PARITY: the RRDBNet NETWORK (and the frame conventions around it: channel order, / 255, layout, clip) is PINNED to
reference-made vectors; `RealESRGANer`'s final `round`, pre_pad / mod-pad, tile loop, gray / alpha / 16-bit branches
and `SRVGGNetCompact` are UNPINNED.  The reference (`/root/reference/src/framewright/processors/pytorch_realesrgan.py`)
does not contain the arithmetic of this path: it constructs and calls two third-party PyPI
packages -- `basicsr` (class `RRDBNet`; last release 1.4.2) and `realesrgan` (classes `RealESRGANer`,
`SRVGGNetCompact`; last release 0.3.0) -- which the reference neither vendors, pins nor declares
(`pyproject.toml:31-89`), which are not installed here and cannot be installed (no network).  The
reference's own tests hold no golden vector, known answer or tolerance for this path
(`tests/test_processors/test_pytorch_realesrgan.py:36` mocks `enhance` to return zeros).  This file
therefore restates the *published* upstream algorithms in plain PyTorch
fp32 -- which is what the reference executes on a CPU-only host
(`pytorch_realesrgan.py:168-169`: half = half_precision and torch.cuda.is_available()) -- and is
anchored on the reference's call sites:
  * architectures by name            pytorch_realesrgan.py:103-129  (RRDBNet(num_in_ch=3, num_out_ch=3,
                                     num_feat=64, num_block=23|6, num_grow_ch=32, scale=4|2))
  * upsampler construction           pytorch_realesrgan.py:160-170, cli.py:742-750
                                     (scale, model, tile, tile_pad, pre_pad, half, gpu_id)
  * the call                         pytorch_realesrgan.py:223  `upsampler.enhance(img, outscale=scale)`
  * PSNR definition                  metrics.py:433-458
What the reference DOES carry is one in-tree ESRGAN generator, `processors/aesrgan_face.py:171-268`
(`ResidualDenseBlock`, `RRDB`, `AESRGAN`: the same trunk and upsampling tail, upstream's parameter names; its
attention gate is the identity at the constructed gamma = 0).  `oracle/ref_pin.py` imports that file unmodified,
loads the same checkpoints into it and commits what it computes (`tests/golden/reference_made/*.npz`);
`tests/test_reference_pin.py` holds `RRDBNet` below to those vectors (x4plus, anime_6B, the x2plus network incl.
`pixel_unshuffle` against torch's own) -- in this container bit for bit against the module run live.
The same file's frame path around the network (`AESRGANFaceRestorer._enhance_face`, :516-542: BGR <-> RGB, / 255,
layout, clip, x 255) pins those conventions of `enhance` too, up to the last step (it truncates, upstream rounds).
The other committed fixtures, tests/golden/*.npz, are outputs of THIS oracle (generator:
oracle/gen_golden.py), i.e. regression vectors, not reference-made vectors.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# basicsr.archs.rrdbnet_arch (restated; constructed at pytorch_realesrgan.py:107,112,117,122,127)
# ----------------------------------------------------------------------------------------------
def pixel_unshuffle(x: torch.Tensor, scale: int) -> torch.Tensor:
    """(b,c,H,W) -> (b, c*s*s, H/s, W/s); output channel = c*s*s + dy*s + dx."""
    b, c, hh, hw = x.shape
    assert hh % scale == 0 and hw % scale == 0
    h, w = hh // scale, hw // scale
    v = x.view(b, c, h, scale, w, scale)
    return v.permute(0, 1, 3, 5, 2, 4).reshape(b, c * scale * scale, h, w)


class ResidualDenseBlock(nn.Module):
    def __init__(self, num_feat: int = 64, num_grow_ch: int = 32):
        super().__init__()
        self.conv1 = nn.Conv2d(num_feat, num_grow_ch, 3, 1, 1)
        self.conv2 = nn.Conv2d(num_feat + num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv3 = nn.Conv2d(num_feat + 2 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv4 = nn.Conv2d(num_feat + 3 * num_grow_ch, num_grow_ch, 3, 1, 1)
        self.conv5 = nn.Conv2d(num_feat + 4 * num_grow_ch, num_feat, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=False)

    def forward(self, x):
        x1 = self.lrelu(self.conv1(x))
        x2 = self.lrelu(self.conv2(torch.cat((x, x1), 1)))
        x3 = self.lrelu(self.conv3(torch.cat((x, x1, x2), 1)))
        x4 = self.lrelu(self.conv4(torch.cat((x, x1, x2, x3), 1)))
        x5 = self.conv5(torch.cat((x, x1, x2, x3, x4), 1))
        return x5 * 0.2 + x


class RRDB(nn.Module):
    def __init__(self, num_feat: int, num_grow_ch: int = 32):
        super().__init__()
        self.rdb1 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb2 = ResidualDenseBlock(num_feat, num_grow_ch)
        self.rdb3 = ResidualDenseBlock(num_feat, num_grow_ch)

    def forward(self, x):
        out = self.rdb3(self.rdb2(self.rdb1(x)))
        return out * 0.2 + x


class RRDBNet(nn.Module):
    def __init__(self, num_in_ch, num_out_ch, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
        super().__init__()
        self.scale = scale
        if scale == 2:
            num_in_ch = num_in_ch * 4
        elif scale == 1:
            num_in_ch = num_in_ch * 16
        self.conv_first = nn.Conv2d(num_in_ch, num_feat, 3, 1, 1)
        self.body = nn.Sequential(*[RRDB(num_feat, num_grow_ch) for _ in range(num_block)])
        self.conv_body = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up1 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_up2 = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_hr = nn.Conv2d(num_feat, num_feat, 3, 1, 1)
        self.conv_last = nn.Conv2d(num_feat, num_out_ch, 3, 1, 1)
        self.lrelu = nn.LeakyReLU(negative_slope=0.2, inplace=False)

    def forward(self, x):
        if self.scale == 2:
            feat = pixel_unshuffle(x, 2)
        elif self.scale == 1:
            feat = pixel_unshuffle(x, 4)
        else:
            feat = x
        feat = self.conv_first(feat)
        body_feat = self.conv_body(self.body(feat))
        feat = feat + body_feat
        feat = self.lrelu(self.conv_up1(F.interpolate(feat, scale_factor=2, mode="nearest")))
        feat = self.lrelu(self.conv_up2(F.interpolate(feat, scale_factor=2, mode="nearest")))
        return self.conv_last(self.lrelu(self.conv_hr(feat)))


# ----------------------------------------------------------------------------------------------
# realesrgan.archs.srvgg_arch (restated; named by BASELINE.json north_star, absent from reference)
# ----------------------------------------------------------------------------------------------
class SRVGGNetCompact(nn.Module):
    def __init__(self, num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4, act_type="prelu"):
        super().__init__()
        assert act_type == "prelu"
        self.upscale = upscale
        body = [nn.Conv2d(num_in_ch, num_feat, 3, 1, 1), nn.PReLU(num_parameters=num_feat)]
        for _ in range(num_conv):
            body += [nn.Conv2d(num_feat, num_feat, 3, 1, 1), nn.PReLU(num_parameters=num_feat)]
        body.append(nn.Conv2d(num_feat, num_out_ch * upscale * upscale, 3, 1, 1))
        self.body = nn.ModuleList(body)
        self.upsampler = nn.PixelShuffle(upscale)

    def forward(self, x):
        out = x
        for layer in self.body:
            out = layer(out)
        out = self.upsampler(out)
        base = F.interpolate(x, scale_factor=self.upscale, mode="nearest")
        return out + base


# model-name table: names and scales from pytorch_realesrgan.py:103-129; SRVGG for the two names
# upstream ships as SRVGGNetCompact checkpoints (SURVEY.md finding 3).
def build_model(model_name: str) -> Tuple[nn.Module, int]:
    table = {
        "RealESRGAN_x4plus": lambda: (RRDBNet(3, 3, scale=4, num_feat=64, num_block=23, num_grow_ch=32), 4),
        "RealESRGAN_x4plus_anime_6B": lambda: (RRDBNet(3, 3, scale=4, num_feat=64, num_block=6, num_grow_ch=32), 4),
        "RealESRGAN_x2plus": lambda: (RRDBNet(3, 3, scale=2, num_feat=64, num_block=23, num_grow_ch=32), 2),
        "realesr-animevideov3": lambda: (SRVGGNetCompact(3, 3, 64, 16, 4, "prelu"), 4),
        "realesr-general-x4v3": lambda: (SRVGGNetCompact(3, 3, 64, 32, 4, "prelu"), 4),
    }
    if model_name not in table:
        raise ValueError(f"Unknown model: {model_name}")
    return table[model_name]()


def load_model(model_name: str, state_dict: Dict[str, torch.Tensor]) -> Tuple[nn.Module, int]:
    """Upstream loader semantics: strict load, eval mode, fp32 on CPU."""
    model, netscale = build_model(model_name)
    model.load_state_dict({k: v.clone().float() for k, v in state_dict.items()}, strict=True)
    model.eval()
    return model, netscale


# ----------------------------------------------------------------------------------------------
# realesrgan.RealESRGANer (restated; constructed at pytorch_realesrgan.py:160-170, cli.py:742-750)
# ----------------------------------------------------------------------------------------------
class RealESRGANer:
    """fp32 CPU restatement of the upstream helper (RGB u8/u16, gray and RGBA branches)."""

    def __init__(self, scale, model_path=None, dni_weight=None, model=None, tile=0, tile_pad=10, pre_pad=10,
                 half=False, device=None, gpu_id=None):
        self.scale = scale
        self.tile_size = tile
        self.tile_pad = tile_pad
        self.pre_pad = pre_pad
        self.mod_scale = None
        self.half = False  # CPU oracle is always fp32
        self.device = torch.device("cpu")
        assert model is not None
        self.model = model.eval()

    def pre_process(self, img: np.ndarray) -> None:
        t = torch.from_numpy(np.transpose(img, (2, 0, 1))).float()
        self.img = t.unsqueeze(0)
        if self.pre_pad != 0:
            self.img = F.pad(self.img, (0, self.pre_pad, 0, self.pre_pad), "reflect")
        if self.scale == 2:
            self.mod_scale = 2
        elif self.scale == 1:
            self.mod_scale = 4
        if self.mod_scale is not None:
            self.mod_pad_h, self.mod_pad_w = 0, 0
            _, _, h, w = self.img.size()
            if h % self.mod_scale != 0:
                self.mod_pad_h = self.mod_scale - h % self.mod_scale
            if w % self.mod_scale != 0:
                self.mod_pad_w = self.mod_scale - w % self.mod_scale
            self.img = F.pad(self.img, (0, self.mod_pad_w, 0, self.mod_pad_h), "reflect")

    def process(self) -> None:
        self.output = self.model(self.img)

    def tile_process(self) -> None:
        batch, channel, height, width = self.img.shape
        s = self.scale
        self.output = self.img.new_zeros((batch, channel, height * s, width * s))
        tiles_x = math.ceil(width / self.tile_size)
        tiles_y = math.ceil(height / self.tile_size)
        for y in range(tiles_y):
            for x in range(tiles_x):
                in_x0 = x * self.tile_size
                in_y0 = y * self.tile_size
                in_x1 = min(in_x0 + self.tile_size, width)
                in_y1 = min(in_y0 + self.tile_size, height)
                pad_x0 = max(in_x0 - self.tile_pad, 0)
                pad_x1 = min(in_x1 + self.tile_pad, width)
                pad_y0 = max(in_y0 - self.tile_pad, 0)
                pad_y1 = min(in_y1 + self.tile_pad, height)
                tile_w = in_x1 - in_x0
                tile_h = in_y1 - in_y0
                in_tile = self.img[:, :, pad_y0:pad_y1, pad_x0:pad_x1]
                with torch.no_grad():
                    out_tile = self.model(in_tile)
                ox0 = (in_x0 - pad_x0) * s
                oy0 = (in_y0 - pad_y0) * s
                self.output[:, :, in_y0 * s:in_y1 * s, in_x0 * s:in_x1 * s] = out_tile[
                    :, :, oy0:oy0 + tile_h * s, ox0:ox0 + tile_w * s]

    def post_process(self) -> torch.Tensor:
        if self.mod_scale is not None:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.mod_pad_h * self.scale, 0:w - self.mod_pad_w * self.scale]
        if self.pre_pad != 0:
            _, _, h, w = self.output.size()
            self.output = self.output[:, :, 0:h - self.pre_pad * self.scale, 0:w - self.pre_pad * self.scale]
        return self.output

    def _run(self, img_rgb: np.ndarray) -> np.ndarray:
        self.pre_process(img_rgb)
        with torch.no_grad():
            if self.tile_size > 0:
                self.tile_process()
            else:
                self.process()
        out = self.post_process()
        return out.data.squeeze(0).float().cpu().clamp_(0, 1).numpy()

    @torch.no_grad()
    def enhance(self, img: np.ndarray, outscale: Optional[float] = None, alpha_upsampler: str = "realesrgan"):
        import cv2

        h_input, w_input = img.shape[0:2]
        img = img.astype(np.float32)
        if np.max(img) > 256:
            max_range = 65535
        else:
            max_range = 255
        img = img / max_range
        alpha = None
        if len(img.shape) == 2:
            img_mode = "L"
            img = cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
        elif img.shape[2] == 4:
            img_mode = "RGBA"
            alpha = img[:, :, 3]
            img = img[:, :, 0:3]
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
            if alpha_upsampler == "realesrgan":
                alpha = cv2.cvtColor(alpha, cv2.COLOR_GRAY2RGB)
        else:
            img_mode = "RGB"
            img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)

        output_img = self._run(img)
        output_img = np.transpose(output_img[[2, 1, 0], :, :], (1, 2, 0))
        if img_mode == "L":
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2GRAY)

        if img_mode == "RGBA":
            if alpha_upsampler == "realesrgan":
                output_alpha = self._run(alpha)
                output_alpha = np.transpose(output_alpha[[2, 1, 0], :, :], (1, 2, 0))
                output_alpha = cv2.cvtColor(output_alpha, cv2.COLOR_BGR2GRAY)
            else:
                h, w = alpha.shape[0:2]
                output_alpha = cv2.resize(alpha, (w * self.scale, h * self.scale), interpolation=cv2.INTER_LINEAR)
            output_img = cv2.cvtColor(output_img, cv2.COLOR_BGR2BGRA)
            output_img[:, :, 3] = output_alpha

        if max_range == 65535:
            output = (output_img * 65535.0).round().astype(np.uint16)
        else:
            output = (output_img * 255.0).round().astype(np.uint8)

        if outscale is not None and outscale != float(self.scale):
            output = cv2.resize(output, (int(w_input * outscale), int(h_input * outscale)),
                                interpolation=cv2.INTER_LANCZOS4)
        return output, img_mode


def make_upsampler(model_name: str, state_dict: Dict[str, torch.Tensor], tile: int = 0, tile_pad: int = 10,
                   pre_pad: int = 0) -> RealESRGANer:
    model, netscale = load_model(model_name, state_dict)
    return RealESRGANer(scale=netscale, model=model, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad, half=False)


# ----------------------------------------------------------------------------------------------
# parity metrics (PSNR per metrics.py:433-458: MSE over uint8, peak 255)
# ----------------------------------------------------------------------------------------------
def calculate_psnr(a: np.ndarray, b: np.ndarray) -> float:
    if a.shape != b.shape:
        return 0.0
    mse = np.mean((a.astype(float) - b.astype(float)) ** 2)
    if mse == 0:
        return float("inf")
    return float(20 * np.log10(255.0 / np.sqrt(mse)))


def parity_report(ref_u8: np.ndarray, got_u8: np.ndarray) -> Dict[str, float]:
    d = np.abs(ref_u8.astype(np.int32) - got_u8.astype(np.int32))
    return {
        "frac_within_1lsb": float(np.mean(d <= 1)),
        "frac_exact": float(np.mean(d == 0)),
        "max_abs": int(d.max()),
        "psnr_db": calculate_psnr(ref_u8, got_u8),
    }


# the gate BASELINE.json states for this path
GATE_FRAC_WITHIN_1LSB = 0.999
GATE_PSNR_DB = 45.0


def synthetic_frame(h: int, w: int, seed: int, kind: str = "mixed") -> np.ndarray:
    """Synthetic u8 BGR frame: uniform noise, or smooth gradients + texture + noise ("mixed")."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 3), np.float32)
    for c in range(3):
        fx, fy = rng.uniform(0.5, 3.0, 2)
        ph = rng.uniform(0, 6.28, 2)
        img[:, :, c] = 0.5 + 0.35 * np.sin(fx * 6.28 * xx / w + ph[0]) * np.cos(fy * 6.28 * yy / h + ph[1])
    img += 0.08 * np.sin(xx[..., None] * 0.9) * np.sin(yy[..., None] * 1.1)
    img += rng.normal(0, 0.04, size=img.shape).astype(np.float32)
    return np.clip(img * 255.0, 0, 255).round().astype(np.uint8)
