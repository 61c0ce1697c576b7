"""``FaradaySky`` -- Galactic rotation-measure foreground at a position or for every pixel of an image
(reference: faraday_sky/faraday_sky.py:15-141).

The reference reads the Hutschenreuter et al. Faraday-sky HEALPix map (``faraday2020v2.hdf5``, absent from the reference
snapshot: ``.MISSING_LARGE_BLOBS``) with h5py and walks image pixels through astropy.wcs, astropy.coordinates and
astropy_healpix on the host.  Here the map lives on the GPU and the whole chain -- pixel -> sky (zenithal TAN / SIN WCS) ->
Galactic -> HEALPix index -> map value -- is one kernel (``csr_galactic_rm_image``); its output is the per-pixel
``phi_gal`` that ``Dataset.subtract_galacticrm`` / ``csr_derotate`` removes.

Differences, all forced by the environment (h5py, astropy, astropy_healpix are not installed):
  * the map is given as arrays (``data=(mean, std)``) or as an ``.npz`` / ``.npy`` file; ``.hdf5`` / ``.fits`` need h5py /
    astropy and raise a clear error without them;
  * angles are plain numbers in radians (objects with ``to_value`` -- astropy Quantities -- are converted);
  * a FITS header is any mapping with NAXIS1/2, CRVAL1/2, CRPIX1/2, CDELT1/2 or CDi_j, CTYPE1 (``...TAN`` / ``...SIN``);
  * nearest-pixel lookup only (the reference's default); ``use_bilinear_interpolation=True`` raises;
  * ``galactic_rm_image`` evaluates element ``[row, col]`` at pixel ``(x = col, y = row)``.  The reference passes
    ``(xx, yy)`` to ``array_index_to_world``, which takes NumPy index order, i.e. it evaluates the transposed position
    (faraday_sky.py:117-123); ``reference_axis_order=True`` reproduces that.
"""
import ctypes
import os
from dataclasses import dataclass, field

import numpy as np


def _radians(v):
    if hasattr(v, "to_value"):
        return np.asarray(v.to_value("rad"), dtype=np.float64)
    return np.asarray(v, dtype=np.float64)


@dataclass(init=True, repr=True)
class FaradaySky:
    filename: str = None
    nside: int = None
    ordering: str = None
    data: tuple = field(default=None, repr=False)

    def __post_init__(self):
        self.ordering = "ring" if self.ordering is None else str(self.ordering).lower()
        if self.ordering not in ("ring", "nested"):
            raise ValueError("ordering must be 'ring' or 'nested'")
        self.extension = None
        if self.data is None:
            if self.filename is None:
                raise ValueError("FaradaySky needs data=(mean, std) or a map file (the reference's faraday2020v2.hdf5 is not "
                                 "part of this package)")
            self.extension = os.path.splitext(str(self.filename))[1]
            if self.extension == ".npz":
                with np.load(self.filename) as f:
                    self.data = (np.asarray(f["faraday_sky_mean"]), np.asarray(f["faraday_sky_std"]))
            elif self.extension == ".npy":
                arr = np.load(self.filename)
                self.data = (arr[0], arr[1])
            elif self.extension == ".hdf5":
                try:
                    import h5py
                except ImportError as exc:
                    raise ImportError("reading %s needs h5py; convert the map to .npz (faraday_sky_mean, faraday_sky_std)"
                                      % self.filename) from exc
                with h5py.File(self.filename, "r") as hf:
                    self.data = (np.array(hf.get("faraday_sky_mean")), np.array(hf.get("faraday_sky_std")))
            else:
                raise ValueError("The extension is not HDF5, NPZ or NPY")
        mean, std = (np.ascontiguousarray(a, dtype=np.float32).reshape(-1) for a in self.data)
        npix = mean.shape[0]
        nside = int(round(np.sqrt(npix / 12.0)))
        if 12 * nside * nside != npix or std.shape[0] != npix:
            raise ValueError("a HEALPix map has 12 nside^2 pixels (got %d and %d)" % (npix, std.shape[0]))
        if self.nside is not None and int(self.nside) != nside:
            raise ValueError("nside=%s does not match the map (%d pixels -> nside %d)" % (self.nside, npix, nside))
        self.nside = nside
        self.data = (mean, std)
        self._dev = None

    def _device_maps(self):
        from .. import _device as dv
        if self._dev is None:
            self._dev = tuple(dv.to_device(a) for a in self.data)
        return self._dev

    def galactic_rm(self, ra=None, dec=None, frame="icrs", use_bilinear_interpolation=False):
        """(mean, std) of the Galactic RM at the given position(s); ``frame``: "icrs" / "fk5" (right ascension,
        declination) or "galactic" (l, b)"""
        from .. import _device as dv
        from .. import _lib
        if use_bilinear_interpolation:
            raise NotImplementedError("only the nearest-pixel lookup (the reference default) is implemented")
        if ra is None or dec is None:
            raise TypeError("Not valid type")
        lon, lat = np.atleast_1d(_radians(ra)), np.atleast_1d(_radians(dec))
        shape = np.broadcast(lon, lat).shape
        lon, lat = (np.ascontiguousarray(np.broadcast_to(a, shape)).reshape(-1) for a in (lon, lat))
        mean_d, std_d = self._device_maps()
        lon_d, lat_d = dv.to_device(lon), dv.to_device(lat)
        n = lon.shape[0]
        out_m = dv.torch.empty((n,), dtype=dv.torch.float32, device=mean_d.device)
        out_s = dv.torch.empty_like(out_m)
        _lib.check(_lib.load().csr_healpix_lookup(dv.tptr(mean_d), dv.tptr(std_d), self.nside, int(self.ordering == "nested"),
                                                  dv.tptr(lon_d), dv.tptr(lat_d), int(str(frame).lower() != "galactic"), n,
                                                  dv.tptr(out_m), dv.tptr(out_s), None, dv.stream_ptr()), "csr_healpix_lookup")
        return out_m.cpu().numpy().reshape(shape), out_s.cpu().numpy().reshape(shape)

    def galactic_rm_image(self, fitsfile=None, use_bilinear_interpolation=False, reference_axis_order=False, device_output=False):
        """(rm_mean, rm_std) ``[NAXIS2, NAXIS1]`` for every pixel of the image the header describes (NumPy arrays, or CUDA
        tensors with ``device_output=True`` -- ready for ``DevicePlan.derotate``)"""
        from .. import _device as dv
        from .. import _lib
        if use_bilinear_interpolation:
            raise NotImplementedError("only the nearest-pixel lookup (the reference default) is implemented")
        header = getattr(fitsfile, "header", fitsfile)
        if header is None or not hasattr(header, "__getitem__"):
            raise TypeError("galactic_rm_image needs a FITS header (a mapping)")

        def get(key, default=None):
            try:
                return header[key]
            except (KeyError, IndexError):
                return default
        nx, ny = int(header["NAXIS1"]), int(header["NAXIS2"])
        ctype = str(get("CTYPE1", "RA---TAN")).upper()
        if ctype.endswith("SIN"):
            proj = 1
        elif ctype.endswith("TAN"):
            proj = 0
        else:
            raise NotImplementedError("projection %r: only zenithal TAN and SIN are implemented" % ctype)
        wcs = _lib.Wcs()
        wcs.crval1, wcs.crval2 = float(header["CRVAL1"]), float(header["CRVAL2"])
        wcs.crpix1, wcs.crpix2 = float(header["CRPIX1"]), float(header["CRPIX2"])
        cdelt1, cdelt2 = float(get("CDELT1", 1.0)), float(get("CDELT2", 1.0))
        wcs.cd11 = float(get("CD1_1", cdelt1 * float(get("PC1_1", 1.0))))
        wcs.cd12 = float(get("CD1_2", cdelt1 * float(get("PC1_2", 0.0))))
        wcs.cd21 = float(get("CD2_1", cdelt2 * float(get("PC2_1", 0.0))))
        wcs.cd22 = float(get("CD2_2", cdelt2 * float(get("PC2_2", 1.0))))
        wcs.lonpole = float(get("LONPOLE", 180.0))
        wcs.proj = proj
        galactic = ctype.startswith("GLON")
        mean_d, std_d = self._device_maps()
        rm_mean = dv.torch.empty((ny, nx), dtype=dv.torch.float32, device=mean_d.device)
        rm_std = dv.torch.empty_like(rm_mean)
        _lib.check(_lib.load().csr_galactic_rm_image(dv.tptr(mean_d), dv.tptr(std_d), self.nside, int(self.ordering == "nested"),
                                                     ctypes.byref(wcs), nx, ny, int(bool(reference_axis_order)), int(galactic),
                                                     dv.tptr(rm_mean), dv.tptr(rm_std), dv.stream_ptr()), "csr_galactic_rm_image")
        print("The Galactic RM in the field is {0:.2f} ± {1:.2f}".format(float(rm_mean.mean()), float(rm_std.mean())))
        if device_output:
            return rm_mean, rm_std
        return rm_mean.cpu().numpy(), rm_std.cpu().numpy()
