"""B200-native batched Stretch simulation engine."""
__version__ = "0.1.0"
