"""Empty stand-in (reference lib/funcs_utils.py:12 imports it; the hot path never calls it)."""
