"""vg_b200 — B200-native Giraffe short-read hot path (libgiraffe_b200.so + thin Python plumbing)."""
__version__ = "0.1.0"
