"""emdr2_b200 — B200-native retrieve-and-read hot path of EMDR2."""
__version__ = "0.1.0"
