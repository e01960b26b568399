"""TEST INFRASTRUCTURE ONLY -- see the package docstring."""


def parse_checkm(path):
    raise NotImplementedError("CheckM parsing is out of scope of the accelerated path; pass a *.tsv completeness file")
