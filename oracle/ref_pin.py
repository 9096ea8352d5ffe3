#!/usr/bin/env python
"""Pins the oracle's RRDBNet arithmetic against the REFERENCE'S OWN code  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The path's network classes live in absent PyPI packages (see oracle/oracle.py), but the reference does carry one
in-tree implementation of the same ESRGAN generator: `processors/aesrgan_face.py` defines `ResidualDenseBlock`
(:171-189: five 3x3 convs over the growing concatenation, LeakyReLU 0.2, `x5 * 0.2 + x`), `RRDB` (:191-204: three of
them, `out * 0.2 + x`) and `AESRGAN` (:206-268: conv_first, the RRDB trunk, conv_body + skip, nearest x2 -> conv_up1 ->
lrelu, nearest x2 -> conv_up2 -> lrelu, conv_hr -> lrelu -> conv_last), with the same parameter names as upstream's
`RRDBNet` -- the reference loads ESRGAN checkpoints into it (`load_state_dict(checkpoint['params'], strict=False)`,
:477).  `AESRGAN` adds `AttentionBlock`s (:142-169) whose output is `gamma * attention + x` with `gamma` created as
`torch.zeros(1)`: at its constructed value the block is the identity, so

    AESRGAN(num_in_ch, 3, num_feat=64, num_block=B, scale=4, num_attention=1)   (one gate, after the first RRDB)

computes exactly what `RRDBNet(num_in_ch, 3, scale=4, num_feat=64, num_block=B, num_grow_ch=32)` computes, by the
reference's own lines.  This script imports that file UNMODIFIED from /root/reference (it needs only numpy, cv2 and
torch), loads the repo's synthetic checkpoints into it (RRDB k of the checkpoint -> the k-th RRDB of `body`, which
is what the index shift of the inserted gate amounts to), runs it on seeded frames and commits input + output as

    tests/golden/reference_made/*.npz      (frame uint8 BGR; `net_out` float32 (3, 4h', 4w') as the reference computed
                                            it; `frame_out_truncated` uint8 BGR: the same frame through the reference's
                                            own pre- / post-processing around the network, `_enhance_face` :516-542)

Covered: RealESRGAN_x4plus (23 blocks), RealESRGAN_x4plus_anime_6B (6 blocks) and the network of RealESRGAN_x2plus
after its pixel-unshuffle (12 input channels; the unshuffle here is torch's own `F.pixel_unshuffle`, so the oracle's
restated one is checked against an independent implementation too).  The reference's own frame path around the
network (`AESRGANFaceRestorer._enhance_face`: BGR <-> RGB, / 255, layout, clip, x 255) pins those conventions of
`RealESRGANer.enhance` as well, up to its final rounding (the reference truncates, upstream rounds).  NOT covered by
any reference code, hence still "unpinned": upstream's `round`, `pre_pad` / mod-pad, the tile loop, the 16-bit / gray /
alpha branches, and `SRVGGNetCompact`.

    python oracle/ref_pin.py          # rewrites tests/golden/reference_made/ (needs /root/reference)
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILE = "/root/reference/src/framewright/processors/aesrgan_face.py"
OUT_DIR = os.path.join(ROOT, "tests", "golden", "reference_made")

# model name, RRDB blocks, network input channels, frame h, w, frame kind, frame seed
CASES = [
    ("RealESRGAN_x4plus", 23, 3, 24, 28, "mixed", 31),
    ("RealESRGAN_x4plus", 23, 3, 20, 36, "noise", 32),
    ("RealESRGAN_x4plus_anime_6B", 6, 3, 28, 24, "mixed", 33),
    ("RealESRGAN_x2plus", 23, 12, 48, 56, "mixed", 34),
]


def case_name(c):
    name, nb, cin, h, w, kind, seed = c
    return f"{name}_{h}x{w}_{kind}{seed}"


def reference_available() -> bool:
    return os.path.isfile(REF_FILE)


def load_reference_module():
    """The reference file itself, executed from where it lies (nothing is copied)."""
    spec = importlib.util.spec_from_file_location("_ref_aesrgan_face", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod          # dataclasses in the file look their module up by name
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True        # /root/reference is read-only
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = dont
    return mod


def reference_network(mod, state_dict, num_block: int, num_in_ch: int):
    """`AESRGAN` as the reference constructs it (:462) with the checkpoint's RRDBs in its trunk."""
    net = mod.AESRGAN(num_in_ch=num_in_ch, num_out_ch=3, num_feat=64, num_block=num_block, scale=4, num_attention=1)
    rrdb_at = [i for i, layer in enumerate(net.body) if type(layer).__name__ == "RRDB"]
    gates = [layer for layer in net.body if type(layer).__name__ == "AttentionBlock"]
    assert len(rrdb_at) == num_block and len(gates) == 1
    assert float(gates[0].gamma.detach().abs().max()) == 0.0      # the constructed value: the gate is the identity
    remapped = {}
    for k, v in state_dict.items():
        parts = k.split(".")
        if parts[0] == "body":
            parts[1] = str(rrdb_at[int(parts[1])])
        remapped[".".join(parts)] = v
    missing, unexpected = net.load_state_dict(remapped, strict=False)     # strict=False as aesrgan_face.py:477
    assert not unexpected and all(m.split(".")[2] in ("gamma", "query", "key", "value") for m in missing), (missing, unexpected)
    return net.eval()


def network_input(img_bgr_u8: np.ndarray, num_in_ch: int):
    """uint8 BGR frame -> the network's input as `RealESRGANer.pre_process` + `RRDBNet.forward` define it: RGB, /255,
    NCHW float32; for the x2 model the 2x2 pixel-unshuffle (torch's own)."""
    import torch
    import torch.nn.functional as F

    x = torch.from_numpy(np.ascontiguousarray(img_bgr_u8[:, :, ::-1].transpose(2, 0, 1))).float().unsqueeze(0) / 255.0
    return F.pixel_unshuffle(x, 2) if num_in_ch == 12 else x


def quantise(net_out: np.ndarray) -> np.ndarray:
    """`RealESRGANer.post_process` + the uint8 branch of `enhance`: clamp to [0, 1], RGB -> BGR, HWC, round(x * 255)."""
    o = np.clip(net_out, 0.0, 1.0)[[2, 1, 0]].transpose(1, 2, 0)
    return (o * 255.0).round().astype(np.uint8)


def reference_frame_path(mod, net, img_bgr_u8: np.ndarray) -> np.ndarray:
    """uint8 BGR frame in -> uint8 BGR frame out through the reference's OWN pre- and post-processing around that
    network: `AESRGANFaceRestorer._enhance_face` (`aesrgan_face.py:516-542`, the method itself, unmodified) -- BGR ->
    RGB, `/ 255.0`, HWC -> NCHW, the model, NCHW -> HWC, `np.clip(x * 255.0, 0, 255).astype(np.uint8)`, RGB -> BGR.
    Same conventions as `RealESRGANer.enhance` except the last step: the reference TRUNCATES where upstream ROUNDS
    (`(x * 255.0).round()`), so upstream's result is this one, or this one + 1 where the fraction is >= 0.5."""
    import types

    import torch

    me = types.SimpleNamespace(_model=net, _device=torch.device("cpu"), config=types.SimpleNamespace(half_precision=False))
    return mod.AESRGANFaceRestorer._enhance_face(me, img_bgr_u8)


def run_case(mod, c):
    """-> (frame, the reference network's float output, the reference frame path's uint8 output or None)."""
    import torch

    sys.path.insert(0, ROOT)
    import framewright_b200  # noqa: F401
    from framewright_b200.archs import make_synthetic_state_dict
    from oracle import oracle

    name, nb, cin, h, w, kind, seed = c
    net = reference_network(mod, make_synthetic_state_dict(name, 0), nb, cin)
    img = oracle.synthetic_frame(h, w, seed=seed, kind=kind)
    with torch.no_grad():
        out = net(network_input(img, cin))
    frame_out = reference_frame_path(mod, net, img) if cin == 3 else None     # (no pixel-unshuffle in that path)
    return img, out.squeeze(0).numpy().astype(np.float32), frame_out


def main():
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    mod = load_reference_module()
    os.makedirs(OUT_DIR, exist_ok=True)
    for c in CASES:
        img, out, frame_out = run_case(mod, c)
        path = os.path.join(OUT_DIR, case_name(c) + ".npz")
        extra = {} if frame_out is None else {"frame_out_truncated": frame_out}
        np.savez_compressed(path, input=img, net_out=out, meta=np.array(c[1:5] + c[6:7]), **extra)
        print(path, out.shape, float(out.min()), float(out.max()), os.path.getsize(path))


if __name__ == "__main__":
    main()
