"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/b200sr.h declares.
Host-only logic behind the ABI (tile planning, weight packing, tile-height choice) is checked against numpy /
oracle restatements.  No compute calls here (no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200sr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200sr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(native_lib):
    from framewright_b200 import _native

    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(native_lib, s), f"{s} declared in include/b200sr.h but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == syms
    assert b"sm_100a" in native_lib.b200sr_version()


def test_library_is_sm100a_tcgen05(native_lib):
    """The built library carries sm_100a SASS with tcgen05 (UTC*MMA), TMEM loads (LDTM) and TMA (UTMALDG)."""
    import shutil
    import subprocess

    from framewright_b200 import _native

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_create_fails_loudly_without_gpu(native_lib):
    from framewright_b200 import _native

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    desc = _native.ModelDesc(_native.ARCH_RRDB, 4, 64, 23, 32)
    h = ctypes.c_void_p()
    assert native_lib.b200sr_create(ctypes.byref(desc), 0, ctypes.byref(h)) == _native.ERR_CUDA
    assert not h.value


@pytest.mark.parametrize("arch,scale,h,w,tile,tile_pad,pre_pad", [
    (0, 4, 720, 1280, 512, 10, 0), (0, 4, 720, 1280, 512, 10, 10), (0, 2, 1080, 1920, 400, 10, 0),
    (0, 2, 51, 77, 32, 4, 0), (1, 4, 100, 90, 64, 10, 10), (0, 4, 64, 64, 0, 10, 0), (0, 2, 45, 63, 0, 10, 3),
])
def test_plan_regions_matches_upstream_tile_process(native_lib, arch, scale, h, w, tile, tile_pad, pre_pad):
    """Region list == the upstream tile loop's slices (oracle/oracle.py::tile_process), including the final
    post_process crop of pre_pad / mod_pad."""
    _check_region_plan(native_lib, arch, scale, h, w, tile, tile_pad, pre_pad)


def test_plan_regions_random_shapes(native_lib):
    """The same equality over 400 seeded random shapes: odd sizes (reflect mod-pad of the x2 network), tiles larger
    and smaller than the frame, tile_pad from 0 up to more than a tile, pre_pad, frames one pixel over a tile boundary."""
    rng = np.random.default_rng(2024)
    for _ in range(400):
        arch = int(rng.integers(0, 2))
        scale = int(rng.choice([2, 4])) if arch == 0 else 4
        h, w = int(rng.integers(8, 700)), int(rng.integers(8, 900))
        tile = int(rng.choice([0, 24, 32, 48, 64, 100, 128, 200, 256, 400, 512]))
        if rng.random() < 0.2 and tile:
            h, w = tile * int(rng.integers(1, 4)) + int(rng.integers(-1, 2)), tile * int(rng.integers(1, 4)) + int(rng.integers(-1, 2))
            h, w = max(h, 8), max(w, 8)
        _check_region_plan(native_lib, arch, scale, h, w, tile, int(rng.choice([0, 1, 4, 10, 16, 40])),
                           int(rng.choice([0, 0, 3, 10])))


def _check_region_plan(native_lib, arch, scale, h, w, tile, tile_pad, pre_pad):
    import math

    cap = 4096
    buf = (ctypes.c_int * (10 * cap))()
    n = native_lib.b200sr_debug_plan_regions(arch, scale, h, w, tile, tile_pad, pre_pad, buf, cap)
    assert 0 < n <= cap, (arch, scale, h, w, tile, tile_pad, pre_pad, n)
    got = [tuple(buf[i * 10:(i + 1) * 10]) for i in range(n)]
    mod = 2 if (arch == 0 and scale == 2) else 1
    Hp = -(-(h + pre_pad) // mod) * mod
    Wp = -(-(w + pre_pad) // mod) * mod
    want = []
    if tile == 0:
        want.append((0, 0, Hp, Wp, 0, 0, h * scale, w * scale, 0, 0))
    else:
        for y in range(math.ceil(Hp / tile)):
            for x in range(math.ceil(Wp / tile)):
                x0, y0 = x * tile, y * tile
                x1, y1 = min(x0 + tile, Wp), min(y0 + tile, Hp)
                px0, px1 = max(x0 - tile_pad, 0), min(x1 + tile_pad, Wp)
                py0, py1 = max(y0 - tile_pad, 0), min(y1 + tile_pad, Hp)
                ch = min((y1 - y0) * scale, h * scale - y0 * scale)
                cw = min((x1 - x0) * scale, w * scale - x0 * scale)
                if ch <= 0 or cw <= 0:
                    continue
                want.append((py0, px0, py1 - py0, px1 - px0, (y0 - py0) * scale, (x0 - px0) * scale, ch, cw,
                             y0 * scale, x0 * scale))
    assert got == want, (arch, scale, h, w, tile, tile_pad, pre_pad)
    # the kept windows tile the destination frame exactly once
    cover = np.zeros((h * scale, w * scale), np.int32)
    for r in got:
        cover[r[8]:r[8] + r[6], r[9]:r[9] + r[7]] += 1
    assert cover.min() == 1 and cover.max() == 1


@pytest.mark.parametrize("cout,cin,fp16", [(32, 64, 0), (32, 96, 0), (64, 192, 0), (3, 64, 1), (48, 64, 1), (64, 64, 1)])
def test_pack_weights_layout(native_lib, cout, cin, fp16):
    """Packed image == numpy restatement: [chunk][dx][blk*COUTP+co][64 ch], 128 B rows, 16 B chunks XOR (row & 7),
    blk <-> ky = 2 - blk, dx <-> kx; bf16/fp16 round-to-nearest-even; zero padding of channels and rows."""
    g = torch.Generator().manual_seed(cout * 1000 + cin)
    w = torch.randn(cout, cin, 3, 3, generator=g)
    wc = w.contiguous()
    nbytes = native_lib.b200sr_debug_pack_weights(ctypes.cast(wc.data_ptr(), ctypes.POINTER(ctypes.c_float)), cout,
                                                  cin, fp16, None, 0)
    coutp = 16 if cout <= 16 else 32 if cout <= 32 else 48 if cout <= 48 else 64
    nchunks = (cin + 63) // 64
    assert nbytes == nchunks * 9 * coutp * 128
    out = np.zeros(nbytes, np.uint8)
    native_lib.b200sr_debug_pack_weights(ctypes.cast(wc.data_ptr(), ctypes.POINTER(ctypes.c_float)), cout, cin, fp16,
                                         out.ctypes.data_as(ctypes.c_void_p), nbytes)
    w16 = w.to(torch.float16 if fp16 else torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)
    want = np.zeros(nbytes // 2, np.uint16)
    for c in range(nchunks):
        for dx in range(3):
            base = ((c * 3 + dx) * 3 * coutp * 128) // 2
            for blk in range(3):
                for co in range(cout):
                    r = blk * coutp + co
                    for j in range(64):
                        ci = c * 64 + j
                        if ci >= cin:
                            break
                        off = r * 128 + (((j // 8) ^ (r & 7)) * 16) + (j % 8) * 2
                        want[base + off // 2] = w16[co, ci, 2 - blk, dx]
    assert np.array_equal(out.view(np.uint16), want)


def test_choose_th_bounds(native_lib):
    for coutp in (16, 32, 48, 64):
        for (n, h, w) in [(1, 720, 1280), (4, 720, 1280), (1, 64, 64), (64, 480, 640), (1, 2880, 5120)]:
            th = native_lib.b200sr_debug_choose_th(coutp, n, h, w, 148)
            assert 1 <= th <= 512 // coutp


def test_fused_rdb_work_list_random_shapes(native_lib):
    """The same proof (coverage exactly once, dependencies produced by earlier items) over 40 seeded random shapes
    around the strip / counter-block / column-tile boundaries."""
    rng = np.random.default_rng(7)
    for _ in range(40):
        n = int(rng.integers(1, 4))
        h = int(rng.choice([int(rng.integers(1, 200)), 8 * int(rng.integers(1, 20)) + int(rng.integers(-1, 2))]))
        w = int(rng.choice([int(rng.integers(1, 600)), 128 * int(rng.integers(1, 5)) + int(rng.integers(-1, 2))]))
        test_fused_rdb_work_list(native_lib, n, max(h, 1), max(w, 1))


@pytest.mark.parametrize("n,h,w", [(1, 720, 1280), (2, 45, 300), (1, 16, 128), (1, 7, 50), (1, 333, 517)])
def test_fused_rdb_work_list(native_lib, n, h, w):
    """Work list of the fused-RDB kernel: every (conv, frame, row, column) is covered exactly once, item row
    ranges start on 8-row boundaries (a multiple of the completion-counter block), and every counter block an item
    reads from a lower conv (its own rows +-1) is produced entirely by EARLIER items of the list -- so in-order
    round-robin execution on co-resident CTAs cannot deadlock."""
    cap = 200000
    buf = (ctypes.c_int * (8 * cap))()
    cnt = native_lib.b200sr_debug_rdb_items(n, h, w, buf, cap)
    assert 0 < cnt <= cap
    items = np.frombuffer(buf, dtype=np.int32, count=cnt * 8).reshape(cnt, 8)
    xt = (w + 127) // 128
    fr = native_lib.b200sr_debug_rdb_flag_rows()
    assert fr in (4, 8)
    nblk = (h + fr - 1) // fr
    cover = np.zeros((5, n, h, xt), np.int32)
    last_writer = np.full((n, 4, nblk), -1, np.int64)     # index of the last item that stores into the block
    for i, (k, fn, y0, rows, tx, fb, d0, d1) in enumerate(items):
        assert 0 <= k < 5 and 1 <= rows <= (16 if k < 4 else 8) and y0 >= 0 and y0 + rows <= h
        assert y0 % fr == 0       # (the 2-CTAs-per-SM build: 8- / 4-row items on 4-row counter blocks)
        cover[k, fn, y0:y0 + rows, tx] += 1
        if k < 4:
            assert fb == (fn * 4 + k) * nblk
            for b in range(y0 // fr, (y0 + rows - 1) // fr + 1):
                last_writer[fn, k, b] = i
        else:
            assert fb == -1
    assert cover.min() == 1 and cover.max() == 1
    nchunks = [1, 2, 2, 3, 3]
    for i, (k, fn, y0, rows, tx, fb, d0, d1) in enumerate(items):
        for c, dbase in ((1, d0), (2, d1)):
            if c >= nchunks[k]:
                assert dbase == -1
                continue
            dep = min(2 * c - 1, k - 1)
            assert dbase == (fn * 4 + dep) * nblk
            for r in range(max(y0 - 1, 0), min(y0 + rows + 1, h)):
                assert 0 <= last_writer[fn, dep, r // fr] < i, (i, k, r)


def test_plan_regions_equals_the_oracles_executable_tile_loop(native_lib):
    """No restated loop in between: the oracle's own `pre_process` / `tile_process` / `post_process` run on a coordinate
    image with a spy network that reports which slice it was given and stamps its output with (call index, local
    row, local column).  From the stitched, cropped result one reads, per tile, the window that was kept and where it
    came from -- exactly the ten numbers of a device-side region -- for 60 seeded random shapes."""
    import framewright_b200  # noqa: F401
    from oracle import oracle

    class Spy(torch.nn.Module):
        def __init__(self, s):
            super().__init__()
            self.s, self.calls = s, []

        def forward(self, t):
            _, _, th, tw = t.shape
            self.calls.append((int(t[0, 0, 0, 0]), int(t[0, 1, 0, 0]), th, tw))
            s = self.s
            out = torch.empty(1, 3, th * s, tw * s)
            out[0, 0] = float(len(self.calls) - 1)
            out[0, 1] = torch.arange(th * s, dtype=torch.float32)[:, None]
            out[0, 2] = torch.arange(tw * s, dtype=torch.float32)[None, :]
            return out

    rng = np.random.default_rng(99)
    for _ in range(60):
        scale = int(rng.choice([2, 4]))
        h, w = int(rng.integers(8, 300)), int(rng.integers(8, 300))
        tile = int(rng.choice([16, 24, 32, 50, 64, 128, 256]))
        tile_pad, pre_pad = int(rng.choice([0, 2, 10, 20])), int(rng.choice([0, 0, 3, 10]))
        spy = Spy(scale)
        up = oracle.RealESRGANer(scale=scale, model=spy, tile=tile, tile_pad=tile_pad, pre_pad=pre_pad)
        up.pre_process(np.zeros((h, w, 3), np.float32))
        _, _, Hp, Wp = up.img.shape
        up.img = torch.stack([torch.arange(Hp, dtype=torch.float32)[:, None].expand(Hp, Wp),
                              torch.arange(Wp, dtype=torch.float32)[None, :].expand(Hp, Wp),
                              torch.zeros(Hp, Wp)])[None]
        up.tile_process()
        out = up.post_process()[0].numpy()
        assert out.shape == (3, h * scale, w * scale)
        want = []
        for r, (py0, px0, th, tw) in enumerate(spy.calls):
            ys, xs = np.nonzero(out[0] == r)
            if ys.size == 0:
                continue                                   # a tile that lies entirely in the cropped padding
            dy0, dx0, ch, cw = ys.min(), xs.min(), ys.max() - ys.min() + 1, xs.max() - xs.min() + 1
            assert ys.size == ch * cw                      # the kept window is a full rectangle
            want.append((py0, px0, th, tw, int(out[1, dy0, dx0]), int(out[2, dy0, dx0]), int(ch), int(cw),
                         int(dy0), int(dx0)))
        cap = 4096
        buf = (ctypes.c_int * (10 * cap))()
        n = native_lib.b200sr_debug_plan_regions(0, scale, h, w, tile, tile_pad, pre_pad, buf, cap)
        got = [tuple(buf[i * 10:(i + 1) * 10]) for i in range(n)]
        assert got == want, (scale, h, w, tile, tile_pad, pre_pad)
