"""TEST INFRASTRUCTURE ONLY -- name-only stand-in for the `mOTUlizer` package, which `scripts/SuperPang.py:22-23`
imports unconditionally.  The stub predicts no core genome, so every contig gets the tag `noinfo`
(`scripts/SuperPang.py:241-244`); core / accessory calling is out of scope of the accelerated path."""
