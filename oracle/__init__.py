"""CPU oracle for the jaxoplanet hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
``jaxoplanet_b200`` never imports it and has no CPU fallback.

Contents
--------
``ref_numpy``   float64/float32 NumPy restatement of the reference formulas
                (operation order kept), each function citing the reference
                file:line it follows.  The same source also runs on torch-f64
                tensors (``ref_numpy.TH`` backend) so that autograd provides the
                derivative oracle with the reference's two custom JVP rules.
``truth_mp``    mpmath (50 digits) evaluation of the *same discretised*
                formulas and of the exact Kepler equation: tells rounding noise
                of the reference algorithm apart from errors of ours.
``ref_c.c``     C/OpenMP restatement (``make -C oracle``): the multi-threaded
                CPU baseline ``bench.py`` times.

Pinning status
--------------
The reference is pure JAX and JAX is not installable in this image (no wheel,
no network), so the oracle cannot be compared with reference outputs produced
here.  It IS pinned against every self-contained known-answer test the
reference's own suite holds for this path (tests/test_oracle_kats.py lists them
with reference file:line).  Against real ``jax[cpu]`` outputs: parity unpinned.
"""
