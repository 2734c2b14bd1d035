"""Stand-in for the `keras` 2.1 symbols the reference's decoder path uses.  TEST INFRASTRUCTURE ONLY."""
from . import backend, layers  # noqa: F401
