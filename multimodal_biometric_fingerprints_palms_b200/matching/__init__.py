"""Mirror of the reference's `src/matching` package on the B200 matcher (include/fpb200_match.h)."""
from .match import MinutiaeMatcher, match_minutiae_pair, match_pairs  # noqa: F401
