"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (fastoverlap_b200/).

Import-time compatibility shim that makes the UNMODIFIED reference package at
/root/reference importable on numpy>=2 / scipy>=1.15.
It exists only in the build container: /root/reference does not travel to the GPU box,
so this module is used solely by oracle/make_golden.py to freeze golden vectors under
tests/golden/ and by the container-only cross-checks in tests/ (skipped when the
reference tree is absent).

Names injected (left = API the reference expects and that upstream removed):
  scipy.special.sph_harm(m, n, azimuth, polar)   reference sphericalAlignment.py:13,62
  scipy.median                                   reference periodicAlignment.py:11
  scipy.special.orthogonal.eval_genlaguerre      reference utils.py:14
  numpy.NaN                                      reference utils.py:395
  module `munkres` (Munkres().compute)           reference utils.py:35-36,64-65
"""
import os
import sys
import types

import numpy as np
import scipy
import scipy.special

REFERENCE_ROOT = os.environ.get("FASTOVERLAP_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "fastoverlap"))


def install():
    """Install the shim and return the imported reference package."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    from scipy.optimize import linear_sum_assignment
    import scipy.special.orthogonal as _orth

    if not hasattr(scipy.special, "sph_harm"):
        scipy.special.sph_harm = (
            lambda m, n, az, pol: scipy.special.sph_harm_y(n, m, pol, az))
    if not hasattr(scipy, "median"):
        scipy.median = np.median
    if not hasattr(_orth, "eval_genlaguerre"):
        _orth.eval_genlaguerre = scipy.special.eval_genlaguerre
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    if "munkres" not in sys.modules:
        m = types.ModuleType("munkres")

        class Munkres(object):
            def compute(self, cost):
                r, c = linear_sum_assignment(np.asarray(cost))
                return list(zip(r.tolist(), c.tolist()))

        m.Munkres = Munkres
        sys.modules["munkres"] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import fastoverlap  # noqa: the unmodified reference
    return fastoverlap


def periodic_align_pele(fo, *args, **kwargs):
    """PeriodicAlign whose cost_matrix follows the pele convention:
    the reference's cost_matrix returns cost[j(X2), i(X1)] (periodicAlignment.py:94-102)
    while the munkres fallback lap() (utils.py:34-41) assumes cost[i(X1), j(X2)].
    Transposing reproduces the documented 1.559 (periodicAlignment.py:609)."""
    base = fo.PeriodicAlign

    class PeriodicAlignPele(base):
        def cost_matrix(self, X1, X2):
            return base.cost_matrix(self, X1, X2).T

    return PeriodicAlignPele(*args, **kwargs)
