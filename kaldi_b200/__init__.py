"""kaldi_b200 — B200-native (sm_100a) implementation of Kaldi's online2
inference hot path behind the reference's own class surfaces.  See DESIGN.md."""
__version__ = "0.1.0"
