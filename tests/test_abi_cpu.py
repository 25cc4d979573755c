"""CPU-only checks of the C-ABI boundary: the library loads without a GPU and exports every
symbol include/jaxoplanet_b200.h declares; argument validation fails loudly with a message;
the product package has no oracle import and no CPU fallback."""

import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "jaxoplanet_b200.h")


@pytest.fixture(scope="module")
def lib():
    from jaxoplanet_b200 import build

    build.build_library()
    from jaxoplanet_b200 import _lib

    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jxp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from jaxoplanet_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.jxp_version() == 1


def test_argument_validation_messages(lib):
    # no compute: every call below must be rejected before any launch
    rc = lib.jxp_kepler_f64(None, 1, None, 1, 10, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.jxp_last_error()
    rc = lib.jxp_kepler_f64(None, 1, None, 1, -5, None, None, None, None, None, None)
    assert rc == -2
    assert lib.jxp_kepler_f64(None, 1, None, 1, 0, None, None, None, None, None, None) == 0
    rc = lib.jxp_limb_dark_f64(None, 9, None, 1, None, 1, 10, 10, None, None, None, None, None)
    assert rc == -2 and b"n_u" in lib.jxp_last_error()
    rc = lib.jxp_limb_dark_f64(None, 2, None, 1, None, 1, 10, 0, None, None, None, None, None)
    assert rc == -3
    rc = lib.jxp_system_light_curve_f64(None, 4, 1, None, 2, 0, None, 100, 0.02, 3, 7, 10, 0,
                                        None, None, None, None, 0, None, None, 0, None)
    assert rc == -3 and b"must be 0, 1, or 2" in lib.jxp_last_error()
    assert lib.jxp_system_workspace_bytes(4096, 1, 100000) > 4096 * 100
    assert lib.jxp_system_workspace_bytes(0, 1, 10) == 0
    # the transit walk keeps nothing per sample: records, 8-byte segments (3 per transit; room for one transit
    # per 72 samples) and one 24-byte slot per chunk of in-window samples for 1/8 of the grid -- about half a
    # byte per (set, sample), against 1.5 bytes for round 1's materialised list; the per-set function tables
    # (3.5 KB per set), the transit items of the sets that use them (24 bytes per transit slot, room for 1/21 of
    # the grid) and the fix list add ~0.15 byte
    one = lib.jxp_system_workspace_bytes(4096, 1, 100000)
    assert 24 * (4096 * 100000 // 256) <= one < 0.75 * 4096 * 100000
    assert lib.jxp_system_walk_stats(None, 4, 1, 10, None, None) == -1
    # development options: one process-wide word, default 0
    assert lib.jxp_get_dev_options() == 0
    assert lib.jxp_set_dev_options(5) == 0 and lib.jxp_get_dev_options() == 5
    assert lib.jxp_set_dev_options(0) == 0
    # gradient step: 80 B per item, 1/16 of the grid + 65536 items
    assert lib.jxp_transit_workspace_bytes(16384, 50000) >= 80 * (16384 * 50000 // 16 + 65536)
    # the list statistics helpers reject null pointers before touching the device
    assert lib.jxp_system_list_stats(None, 4, 1, 10, None, None) == -1
    assert lib.jxp_transit_list_stats(None, 4, 10, None, None) == -1


def test_facade_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np

    from jaxoplanet_b200.core import kepler

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kepler(np.linspace(0, 1, 5), 0.1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "jaxoplanet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
    code = "import sys, jaxoplanet_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_body_validation_errors():
    # keplerian.py:176-238 error behaviour
    from jaxoplanet_b200.orbits.keplerian import Body, Central

    with pytest.raises(ValueError, match="Exactly one of period or semimajor"):
        Body(period=1.0, semimajor=2.0)
    with pytest.raises(ValueError, match="must be scalars"):
        Body(period=[1.0, 2.0])
    with pytest.raises(ValueError, match="Both or neither of eccentricity and omega_peri"):
        Body(period=1.0, eccentricity=0.1)
    with pytest.raises(ValueError, match="Only one of impact_param and inclination"):
        Body(period=1.0, impact_param=0.1, inclination=1.0)
    with pytest.raises(ValueError, match="Only one of time_transit or time_peri"):
        Body(period=1.0, time_transit=0.1, time_peri=1.0)
    with pytest.raises(ValueError, match="exactly two of mass, radius, and density"):
        Central(mass=1.0, radius=1.0, density=1.0)
    with pytest.raises(ValueError, match="must be scalars"):
        Central(mass=[1.0, 2.0], radius=1.0)


def test_orbital_constants_match_oracle():
    # host-side mirror of OrbitalBody.__init__ (keplerian.py:262-340) vs the oracle
    import numpy as np

    from jaxoplanet_b200.orbits.keplerian import Central, System
    from oracle import ref_numpy as ref

    kw = dict(period=3.456, time_transit=0.3, impact_param=0.3, radius=0.1, eccentricity=0.1,
              omega_peri=0.5)
    body = System(Central(mass=1.3, radius=1.1)).add_body(**kw).bodies[0]
    want = ref.orbital_body(central_mass=1.3, central_radius=1.1, **kw)
    for name in ("period", "semimajor", "time_transit", "time_ref", "sin_omega_peri",
                 "cos_omega_peri", "sin_inclination", "cos_inclination", "impact_param"):
        np.testing.assert_allclose(float(getattr(body, name)), want[name], rtol=2e-16, atol=1e-300,
                                   err_msg=name)
    # circular body, inclination given
    body = System().add_body(period=12.5, inclination=0.3, mass=0.1).bodies[0]
    want = ref.orbital_body(period=12.5, inclination=0.3, mass=0.1)
    for name in ("semimajor", "time_ref", "sin_inclination", "cos_inclination", "impact_param"):
        np.testing.assert_allclose(float(getattr(body, name)), want[name], rtol=4e-16, err_msg=name)


def test_pack_bodies_batch_axis_comes_from_any_field():
    """A fixed ephemeris with sampled impact parameter / radius (scalar period, batched b and r) is a batch
    of parameter sets: the set axis is the longest leading axis over every packed field."""
    import numpy as np

    import jaxoplanet_b200 as jxp
    from jaxoplanet_b200.light_curves.limb_dark import pack_bodies
    from jaxoplanet_b200.orbits.keplerian import Central, System

    rng = np.random.default_rng(0)
    b = rng.uniform(0, 1, 8)
    with jxp.batched():
        sysm = System(Central()).add_body(period=3.0, impact_param=b, radius=0.1)
    recs, n_sets, batched = pack_bodies(sysm)
    assert n_sets == 8 and batched and tuple(recs.shape) == (8, 1, 12)
    recs = recs.cpu().numpy()
    assert np.all(recs[:, 0, 0] == 3.0)
    # cos i = b R* / a differs per set
    assert len(np.unique(recs[:, 0, 6])) == 8
    with pytest.raises((ValueError, RuntimeError)):  # two different batch lengths: refused, never mis-stacked
        with jxp.batched():
            pack_bodies(System(Central()).add_body(period=np.full(4, 3.0), impact_param=b, radius=0.1))


def test_interpolate_rejects_batched_orbits_and_integrate_keeps_loglike_signature():
    import inspect

    import numpy as np

    import jaxoplanet_b200 as jxp
    from jaxoplanet_b200.light_curves import limb_dark_light_curve, transforms
    from jaxoplanet_b200.light_curves.limb_dark import log_likelihood
    from jaxoplanet_b200.orbits.keplerian import Central, System

    with jxp.batched():
        sysm = System(Central()).add_body(period=np.array([3.0, 4.0]), impact_param=0.2, radius=0.1)
    lc = limb_dark_light_curve(sysm, [0.3, 0.2])
    ll = log_likelihood(sysm, [0.3, 0.2])
    integ = transforms.integrate(ll, exposure_time=0.01, num_samples=3)
    assert list(inspect.signature(integ).parameters) == ["time", "y", "sigma"]
    assert list(inspect.signature(transforms.integrate(lc, exposure_time=0.01)).parameters) == ["time"]
    # interpolate() builds ONE table for ONE ephemeris: a batched orbit must be refused before any compute
    with pytest.raises((ValueError, RuntimeError)) as ei:
        transforms.interpolate(lambda tt: np.zeros((2, 16, 1)), period=3.0, time_transit=0.0, num_samples=16)(
            np.linspace(0, 1, 5))
    assert "grid axis" in str(ei.value) or "leading" in str(ei.value)
