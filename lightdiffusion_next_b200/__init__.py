"""B200-native SD1.5 sampling engine — drop-in backend for LightDiffusion-Next's sampler hot path."""
__version__ = "0.1.0"
