"""Empty stand-in so the reference's `funcs_utils.py:12` (`import matplotlib.pyplot`) imports.
TEST INFRASTRUCTURE ONLY."""
