"""Truncation strategies (MatrixAlgebraKit.TruncationStrategy as used by TNRKit).

Only `truncrank` runs on the device path (the north star keeps
`run!(scheme, truncrank(chi), maxiter(n))`); other strategies raise."""
from __future__ import annotations


class TruncationStrategy:
    def __and__(self, other):
        raise NotImplementedError(
            "composite truncation strategies are not on the B200 path; use truncrank(chi)")


class truncrank(TruncationStrategy):
    """Keep the `howmany` values of largest magnitude (MatrixAlgebraKit.truncrank)."""

    def __init__(self, howmany: int):
        if int(howmany) < 1:
            raise ValueError("truncrank: howmany must be >= 1")
        self.howmany = int(howmany)

    @property
    def chi(self):
        return self.howmany

    def __repr__(self):
        return f"truncrank({self.howmany})"


class trunctol(TruncationStrategy):
    def __init__(self, *a, **k):
        raise NotImplementedError("trunctol is not on the B200 path; use truncrank(chi)")
