"""One-file FrameWright plugin: Real-ESRGAN upscaling on the B200 engine.

Copy (or symlink) this file into a FrameWright plugin directory -- `~/.framewright/plugins/`, or any directory passed as
`PluginManager(plugin_dirs=[...])` (`/root/reference/src/framewright/plugins/manager.py:259-285`); the reference's
`PluginLoader.load_from_file` (`:172-205`) executes it and registers every `PluginBase` subclass it finds, i.e. the class
below.  Then

    mgr = PluginManager(plugin_dirs=[...]); mgr.set_device("cuda:0")
    up = mgr.get_processor("b200-realesrgan", {"model_name": "RealESRGAN_x4plus"})
    out = up.process_batch(frames, start_frame=0)

`framewright_b200` must be importable (the repository root on `sys.path`).
"""
from framewright_b200.plugin_adapters import make_processor_plugin

B200RealESRGANPlugin = make_processor_plugin()
