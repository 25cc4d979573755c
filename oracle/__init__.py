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
Pinned two ways.  (1) Against outputs of the reference's OWN source files: jax/jaxlib/equinox
are not installable in this image, but the reference's unmodified ``core/kepler.py``,
``core/limb_dark.py``, ``orbits/keplerian.py``, ``orbits/transit.py``, ``light_curves/*.py``
execute eagerly on a NumPy-backed ``jax`` shim (tests/golden/jax_shim); their outputs are committed
as tests/golden/reference_v1.npz (made by tests/golden/make_reference_vectors.py) and
tests/test_reference_vectors.py holds the oracle to them: bit-exact for Kepler, orbit constants,
positions; <= 4 eps for fluxes (summation order of the final dot product).  (2) Against every
self-contained known-answer test of the reference's suite (tests/test_oracle_kats.py).
What stays unpinned: XLA's own libm/fusion versus NumPy's (<= 1 ulp per call) and derivatives
produced by JAX tracing (the reference's two hand-written JVP rules ARE executed and compared).
"""
