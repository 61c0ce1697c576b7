"""Pin the oracle's restatement of the sky chain behind FaradaySky.galactic_rm_image (faraday_sky/faraday_sky.py:62-141):
HEALPix (Gorski et al. 2005) known answers and the pix2ang -> ang2pix round trip, the Galactic frame's defining points,
and the zenithal WCS at its reference pixel.  astropy / astropy_healpix are absent: parity with them is unpinned."""
import numpy as np
import pytest

from oracle import restatement as ob


@pytest.mark.parametrize("nside", [1, 2, 4, 32, 512])
def test_healpix_ring_round_trip(nside):
    pix = np.arange(12 * nside * nside) if nside <= 32 else np.random.RandomState(1).randint(0, 12 * nside * nside, 20000)
    theta, phi = ob.healpix_pix2ang_ring(nside, pix)
    np.testing.assert_array_equal(ob.healpix_ang2pix_ring(nside, theta, phi), pix)


def test_healpix_known_answers():
    # nside = 1: four pixels per ring at z = 2/3, 0, -2/3; polar rings start at phi = 45 deg, the equatorial one at 0
    theta, phi = ob.healpix_pix2ang_ring(1, np.arange(12))
    np.testing.assert_allclose(np.cos(theta), np.repeat([2 / 3, 0.0, -2 / 3], 4), atol=1e-15)
    np.testing.assert_allclose(np.degrees(phi), [45, 135, 225, 315, 0, 90, 180, 270, 45, 135, 225, 315], atol=1e-12)
    # poles and the equator at nside = 512 (npix = 3145728): first pixel, last pixel, first pixel of the middle ring
    ns = 512
    assert ob.healpix_ang2pix_ring(ns, 1e-9, 0.1) == 0
    assert ob.healpix_ang2pix_ring(ns, np.pi - 1e-9, 6.2) == 12 * ns * ns - 1
    assert ob.healpix_ang2pix_ring(ns, 0.5 * np.pi, 1e-9) == 2 * ns * (ns - 1) + ns * 4 * ns   # ring 2 nside = nside rings below the cap


def test_galactic_frame_and_wcs_reference_pixel():
    l, b = ob.icrs_to_galactic(np.radians(192.85948), np.radians(27.12825))       # north Galactic pole
    assert np.degrees(b) == pytest.approx(90.0, abs=1e-6)
    l, b = ob.icrs_to_galactic(np.radians(266.40499), np.radians(-28.93617))       # Sgr A*: l = b = 0 by definition (to ~0.1")
    assert abs(np.degrees(b)) < 1e-3 and min(np.degrees(l), 360 - np.degrees(l)) < 1e-3
    l, b = ob.icrs_to_galactic(0.0, np.radians(90.0))                              # north celestial pole: l = 122.93192
    assert np.degrees(l) == pytest.approx(122.93192, abs=1e-9) and np.degrees(b) == pytest.approx(27.12825, abs=1e-9)
    for proj in ("TAN", "SIN"):
        ra, dec = ob.wcs_pixel_to_sky(99.0, 49.0, (150.0, -30.0), (100.0, 50.0), [[-1e-3, 0.0], [0.0, 1e-3]], proj)
        assert np.degrees(ra) == pytest.approx(150.0, abs=1e-10) and np.degrees(dec) == pytest.approx(-30.0, abs=1e-10)
        # one pixel north: declination grows by CDELT2 (to second order), right ascension unchanged
        ra, dec = ob.wcs_pixel_to_sky(99.0, 50.0, (150.0, -30.0), (100.0, 50.0), [[-1e-3, 0.0], [0.0, 1e-3]], proj)
        assert np.degrees(dec) == pytest.approx(-30.0 + 1e-3, abs=1e-9) and np.degrees(ra) == pytest.approx(150.0, abs=1e-9)
        # one pixel to the right with negative CDELT1: right ascension decreases by CDELT1 / cos(dec)
        ra, dec = ob.wcs_pixel_to_sky(100.0, 49.0, (150.0, -30.0), (100.0, 50.0), [[-1e-3, 0.0], [0.0, 1e-3]], proj)
        assert np.degrees(ra) == pytest.approx(150.0 - 1e-3 / np.cos(np.radians(30.0)), abs=1e-7)
