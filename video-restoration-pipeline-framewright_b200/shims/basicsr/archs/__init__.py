__b200sr_shim__ = True
