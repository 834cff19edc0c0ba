"""Builds oracle/_ref/ (the reference's OWN sources compiled where they lie under
/root/reference) — TEST INFRASTRUCTURE ONLY.  Filled in per stage; raises when
/root/reference is absent (the GPU box uses the prebuilt files)."""
from __future__ import annotations

import os

REF = "/root/reference"


def build_all(quiet: bool = False) -> None:
    if not os.path.isdir(REF):
        raise RuntimeError("/root/reference not present")
    from . import ref_decoder, ref_det, ref_feat, ref_nnet
    ref_feat.build(quiet=quiet)
    ref_nnet.build(quiet=quiet)
    ref_decoder.build(quiet=quiet)
    ref_det.build(quiet=quiet)
