"""`basicsr.archs.rrdbnet_arch` shim: RRDBNet(...) returns an architecture descriptor + state-dict holder
(the reference constructs it at processors/pytorch_realesrgan.py:107-127 and cli.py:715-723 and hands it to
RealESRGANer as `model=`)."""
from framewright_b200.archs import ArchDesc
from framewright_b200.upsampler import ArchSpec

__b200sr_shim__ = True


def RRDBNet(num_in_ch=3, num_out_ch=3, scale=4, num_feat=64, num_block=23, num_grow_ch=32):
    if (num_in_ch, num_out_ch, num_feat, num_grow_ch) != (3, 3, 64, 32) or scale not in (2, 4):
        raise NotImplementedError("the B200 engine supports RRDBNet(3, 3, scale in {2,4}, 64, num_block, 32)")
    return ArchSpec(ArchDesc("rrdb", int(scale), num_block=int(num_block)))
