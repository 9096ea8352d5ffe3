"""`basicsr` shim (only what the reference's upscaling path imports)."""
__version__ = "1.4.2+b200sr"
__b200sr_shim__ = True
