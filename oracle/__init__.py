"""CPU oracle for the MARTINI particle->datacube projection.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / reference
arm may import it, and only as the checker or as the timed CPU baseline.  The
product path (``martini_b200``) never imports this package and fails loudly when its
CUDA library is missing.

Parity status: PINNED.  ``oracle/martini_oracle.py`` restates the reference's numpy
arithmetic line by line (each function cites the ``/root/reference`` file:line it
follows).  It is pinned two ways:

* against golden vectors produced by executing the reference's OWN, unmodified source
  files (``martini/sph_kernels.py``, ``martini/spectral_models.py`` and the
  ``_BaseMartini`` hot loop of ``martini/martini.py``) under ``oracle/refshim.py`` --
  a units shim standing in for the missing ``astropy`` dependency in which every unit
  has scale 1, inputs being supplied already in (pix, km/s, Mpc, Msun, arcsec).  The
  generating script is ``tests/golden/make_golden.py``; the oracle reproduces those
  vectors bit-for-bit;
* against the analytic known-answer tests the reference's own test-suite holds for
  this path (``tests/test_sph_kernels.py``, ``tests/test_spectral_models.py``,
  ``tests/test_martini.py``), replayed astropy-free in ``tests/test_oracle_kats.py``.

What stays unpinned (astropy is not installable here): the ulp-level operation order
astropy's unit conversions impose (e.g. m/s channel edges minus km/s velocities,
``spectral_models.py:411``).  The effect is a few ulp on erf arguments, ~1e-16
relative on voxels, ten orders below the 1e-6 x peak tolerance.
"""
