"""`realesrgan.archs.srvgg_arch` shim: SRVGGNetCompact(...) returns an architecture descriptor."""
from framewright_b200.archs import ArchDesc
from framewright_b200.upsampler import ArchSpec

__b200sr_shim__ = True


def SRVGGNetCompact(num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4, act_type="prelu"):
    if (num_in_ch, num_out_ch, num_feat, act_type) != (3, 3, 64, "prelu") or upscale != 4:
        raise NotImplementedError("the B200 engine supports SRVGGNetCompact(3, 3, 64, num_conv, 4, 'prelu')")
    return ArchSpec(ArchDesc("srvgg", int(upscale), num_block=int(num_conv)))
