"""Pins that anchor the (otherwise unpinned) oracle and the engine's architecture tables to PUBLISHED facts about the
upstream networks the reference constructs (`pytorch_realesrgan.py:103-129`): parameter counts and the full
state-dict key list a strict `load_state_dict` of the released checkpoints requires."""
import os

import pytest
import torch


def _count(sd):
    return sum(int(v.numel()) for v in sd.values())


def test_parameter_counts_match_published_numbers():
    from framewright_b200.archs import MODEL_ARCHS, conv_layers, make_synthetic_state_dict, prelu_layers
    from oracle import oracle

    # RealESRGAN_x4plus: 16 697 987 parameters (the released .pth is 64 MB of fp32)
    expected = {"RealESRGAN_x4plus": 16_697_987}
    # RealESRGAN_x2plus differs from x4plus only in conv_first (12 instead of 3 input channels after pixel-unshuffle):
    # + 9 ch x 64 x 3 x 3 = 5 184 parameters (the released files differ by exactly 4 x 5 184 bytes)
    expected["RealESRGAN_x2plus"] = expected["RealESRGAN_x4plus"] + 9 * 64 * 9
    for name, arch in MODEL_ARCHS.items():
        n_tables = sum(ci * co * 9 + co for _, ci, co in conv_layers(arch)) + arch.num_feat * len(prelu_layers(arch))
        model, _ = oracle.build_model(name)
        n_oracle = sum(p.numel() for p in model.parameters())
        n_synth = _count(make_synthetic_state_dict(name, 0))
        assert n_tables == n_oracle == n_synth, name
        if name in expected:
            assert n_oracle == expected[name], (name, n_oracle)
    # RRDB trunk dominates: 23 blocks x 3 RDBs x 239 616 weights (+ biases)
    assert 69 * (239_616 + 4 * 32 + 64) == 16_533_504 + 69 * 192


def _upstream_rrdb_keys(num_block):
    keys = ["conv_first.weight", "conv_first.bias"]
    for b in range(num_block):
        for r in (1, 2, 3):
            for c in (1, 2, 3, 4, 5):
                keys += [f"body.{b}.rdb{r}.conv{c}.weight", f"body.{b}.rdb{r}.conv{c}.bias"]
    for n in ("conv_body", "conv_up1", "conv_up2", "conv_hr", "conv_last"):
        keys += [f"{n}.weight", f"{n}.bias"]
    return keys


def _upstream_srvgg_keys(num_conv):
    keys = []
    for i in range(num_conv + 2):          # first conv, num_conv body convs, last conv at even indices
        keys += [f"body.{2 * i}.weight", f"body.{2 * i}.bias"]
        if i <= num_conv:                  # a PReLU (one weight vector) after every conv but the last
            keys.append(f"body.{2 * i + 1}.weight")
    return keys


@pytest.mark.parametrize("name,keys", [
    ("RealESRGAN_x4plus", _upstream_rrdb_keys(23)),
    ("RealESRGAN_x2plus", _upstream_rrdb_keys(23)),
    ("RealESRGAN_x4plus_anime_6B", _upstream_rrdb_keys(6)),
    ("realesr-general-x4v3", _upstream_srvgg_keys(32)),
    ("realesr-animevideov3", _upstream_srvgg_keys(16)),
])
def test_state_dict_key_list_is_upstreams(name, keys):
    """The key list of the released checkpoints (basicsr RRDBNet: conv_first, body.{i}.rdb{1..3}.conv{1..5},
    conv_body, conv_up1, conv_up2, conv_hr, conv_last; SRVGGNetCompact: body.{k}) -- the oracle's modules, the
    engine's layer table and the synthetic weights all carry exactly these keys, so a strict load of a real file works."""
    from framewright_b200.archs import MODEL_ARCHS, conv_layers, make_synthetic_state_dict, prelu_layers
    from oracle import oracle

    model, _ = oracle.build_model(name)
    assert sorted(model.state_dict().keys()) == sorted(keys)
    arch = MODEL_ARCHS[name]
    table = [n + s for n, _, _ in conv_layers(arch) for s in (".weight", ".bias")] + prelu_layers(arch)
    assert sorted(table) == sorted(keys)
    sd = make_synthetic_state_dict(name, 0)
    assert sorted(sd.keys()) == sorted(keys)
    model.load_state_dict(sd, strict=True)        # shapes too
    if name == "RealESRGAN_x4plus":
        assert len(keys) == 702
        assert tuple(sd["conv_first.weight"].shape) == (64, 3, 3, 3)
        assert tuple(sd["body.22.rdb3.conv5.weight"].shape) == (64, 192, 3, 3)
        assert tuple(sd["conv_last.weight"].shape) == (3, 64, 3, 3)
    if name == "RealESRGAN_x2plus":
        assert tuple(sd["conv_first.weight"].shape) == (64, 12, 3, 3)


def test_checkpoint_resolution_never_invents_weights(tmp_path, monkeypatch):
    """`RealESRGANer` weight lookup: a file path, a URL resolved offline to the weights directory, a bare file name
    (face_restore.py:391) -- and FileNotFoundError otherwise (the constructor must not fall back to random weights)."""
    from framewright_b200 import upsampler as up
    from framewright_b200.archs import make_synthetic_state_dict

    wdir = tmp_path / "w"
    wdir.mkdir()
    monkeypatch.setenv("B200SR_WEIGHTS_DIR", str(wdir))
    monkeypatch.setenv("B200SR_DOWNLOAD_TIMEOUT", "1")
    with pytest.raises(FileNotFoundError):
        up._resolve_checkpoint("https://example.invalid/releases/RealESRGAN_x4plus.pth")
    with pytest.raises(FileNotFoundError):
        up._resolve_checkpoint("RealESRGAN_x4plus.pth")
    sd = make_synthetic_state_dict("realesr-animevideov3", 3)
    torch.save({"params": sd}, str(wdir / "realesr-animevideov3.pth"))
    p = up._resolve_checkpoint("https://example.invalid/releases/download/v0.2.5.0/realesr-animevideov3.pth")
    assert os.path.samefile(p, wdir / "realesr-animevideov3.pth")
    assert os.path.samefile(up._resolve_checkpoint("realesr-animevideov3.pth"), p)
    loaded = up._load_checkpoint(p)
    assert sorted(loaded.keys()) == sorted(sd.keys()) and torch.equal(loaded["body.0.weight"], sd["body.0.weight"])
    assert up._arch_for_model_path(p) == "realesr-animevideov3"
