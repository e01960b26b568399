"""TEST INFRASTRUCTURE ONLY -- name-only stub for `from speedict import Rdict, ...`
(`Assembler.py:18`).  `--lowmem` (RocksDB k-mer store) is never used by the oracle."""


class _Unavailable:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("speedict is not available in this image")


Rdict = Options = SliceTransform = PlainTableFactoryOptions = _Unavailable
