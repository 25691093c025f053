"""Import-only stand-in for matplotlib (plotting is not on the hot path).

TEST INFRASTRUCTURE.  The reference imports ``matplotlib.pyplot`` at module
scope (vsc/metrics.py:14, vsc/baseline/sscd_baseline.py:25); matplotlib is not
installed in this image.  Nothing numerical depends on it.
"""
