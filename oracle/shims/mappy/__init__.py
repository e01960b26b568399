"""TEST INFRASTRUCTURE ONLY -- name-only stub so the reference's `from mappy import Aligner`
(`Assembler.py:17`) resolves.  The hot path never aligns (bubble popping is out of scope)."""


class Aligner:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("mappy is not available in this image; bubble popping is out of scope")
