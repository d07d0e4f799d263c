"""pota_b200 — B200-native implementation of the lentil (zpelgrims/pota) per-ray hot paths."""
__version__ = "0.4.0"
