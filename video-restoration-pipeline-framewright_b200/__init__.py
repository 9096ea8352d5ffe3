"""B200-native Real-ESRGAN upscaling path (package root)."""
