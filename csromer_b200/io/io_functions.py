"""Cube I/O either side of the hot path (reference: ``io/io_functions.py``: ``filter_cubes`` :15-33, ``Reader`` :36-146,
``Writer`` :149-215).  Same class and method names; ``astropy.io.fits`` is replaced by the NumPy codec in
``fits_lite`` (astropy is absent here) and dask by plain arrays.  ``DeviceCube`` is the batch extension: it keeps the
frequency-first Q / U planes on the GPU and hands pixel slabs to the kernels through the transposes of
``csrc/cubeio.cu`` (``csr_pack_slab`` / ``csr_unpack_slab`` / ``csr_channel_stats``).
"""
import sys

import numpy as np
import torch

from .. import _device as dv
from .. import _lib
from . import fits_lite


def filter_cubes(data_I, data_Q, data_U, header, additional_outlier_idxs=None):
    """drop the channels whose I, Q and U images all sum to zero (io_functions.py:15-33)"""
    nfreqs = header["NAXIS3"]
    nu = header["CRVAL3"] + np.arange(0, nfreqs) * header["CDELT3"]
    sums = [np.nansum(c, axis=(1, 2)) for c in (data_I, data_Q, data_U)]
    keep = np.where((sums[0] != 0.0) | (sums[1] != 0.0) | (sums[2] != 0.0))[0]
    if additional_outlier_idxs:
        keep = np.setxor1d(keep, additional_outlier_idxs)
    print("Filtering {0:.2f}% of the total data".format(100.0 * (nfreqs - len(keep)) / nfreqs))
    return data_I[keep], data_Q[keep], data_U[keep], nu[keep]


class Reader:

    def __init__(self, stokes_I_name=None, stokes_Q_name=None, stokes_U_name=None, Q_cube_name=None, U_cube_name=None,
                 freq_file_name=None, numpy_file=None):
        self.stokes_I_name, self.stokes_Q_name, self.stokes_U_name = stokes_I_name, stokes_Q_name, stokes_U_name
        self.Q_cube_name, self.U_cube_name = Q_cube_name, U_cube_name
        self.freq_file_name, self.numpy_file = freq_file_name, numpy_file

    def readCube(self, file=None, stokes=None, memmap=True):
        if stokes is not None and file is None:
            file = self.Q_cube_name if stokes == "Q" else self.U_cube_name
        try:
            header, data = fits_lite.open_image(file, memmap=memmap)
        except FileNotFoundError:
            print("FileNotFoundError: The FITS file cannot be found")
            sys.exit(1)
        image = data.squeeze()
        print("FITS shape: ", image.shape)
        return header, image

    def readQU(self, memmap=True):
        header, Q = self.readCube(file=self.Q_cube_name, memmap=memmap)
        header, U = self.readCube(file=self.U_cube_name, memmap=memmap)
        return Q, U, header

    def readImage(self, name=None, stokes=None):
        if name is None and stokes is not None:
            name = {"I": self.stokes_I_name, "Q": self.stokes_Q_name}.get(stokes, self.stokes_U_name)
        header, data = fits_lite.open_image(name, memmap=False)
        return header, np.squeeze(data)

    def readNumpyFile(self):
        try:
            arr = np.load(self.numpy_file)
        except FileNotFoundError:
            print("FileNotFoundError: The numpy file cannot be found")
            sys.exit(1)
        return arr[:, :, 0], arr[:, :, 1]

    def readHeader(self, name=None):
        with open(self.Q_cube_name if name is None else name, "rb") as f:
            return fits_lite.read_header(f)[0]

    def getFileNFrequencies(self):
        try:
            with open(self.freq_file_name, "r") as f:
                return np.array([float(line.rstrip("\n")) for line in f.readlines()])
        except Exception:
            print("Cannot open file")
            sys.exit(1)

    def readFreqsNumpyFile(self):
        return np.load(self.freq_file_name)


class Writer:

    def __init__(self, output=""):
        self.output = output

    def writeFITSCube(self, cube, header, nphi, phi, dphi, output=None, overwrite=True):
        """Faraday-depth cube (n_phi, Y, X) -> FITS; complex cubes are stacked (2, n_phi, Y, X) = real, imaginary"""
        header = header.copy() if isinstance(header, fits_lite.Header) else fits_lite.Header(header)
        header["NAXIS"] = 4
        header["NAXIS3"] = (nphi, "Length of Faraday depth axis")
        header["NAXIS4"] = (2, "Real and imaginary")
        header["CTYPE3"] = "Phi"
        header["CDELT3"] = dphi
        header["CUNIT3"] = "rad/m/m"
        header["CRVAL3"] = phi[0]
        cube = np.asarray(cube)
        if np.iscomplexobj(cube):
            cube = np.stack([cube.real, cube.imag], axis=0)
        fits_lite.write_image(self.output if output is None else output, cube, header, overwrite=overwrite)

    def writeNPCube(self, cube, output=None):
        np.save(self.output if output is None else output, cube)

    def writeFITS(self, data=None, header=None, output=None, overwrite=True):
        fits_lite.write_image(self.output if output is None else output, data, header, overwrite=overwrite)


class DeviceCube:
    """Frequency-first Stokes Q / U cube ``(m, Y, X)`` resident on the GPU (fp32), the format of README.md:288-297.

    ``slab(p0, P)`` returns the pixel-major planar rows ``[P, ld2m]`` the kernels take plus the finite-sample mask
    ``[P, m]`` (non-finite samples read as 0 and must get weight 0); ``channel_stats`` gives each channel's nansum and
    nanstd (``calculate_noise``, ``filter_cubes``); ``store`` scatters planar Faraday spectra ``[P, ld2n]`` into
    ``(2, n, Y, X)`` real / imaginary planes on the device (the Writer's FITS layout).
    """

    def __init__(self, Q, U, device=None):
        dv.require_cuda()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.q = torch.as_tensor(np.ascontiguousarray(Q, dtype=np.float32) if not dv.is_torch(Q) else Q, dtype=torch.float32).to(self.device).contiguous()
        self.u = torch.as_tensor(np.ascontiguousarray(U, dtype=np.float32) if not dv.is_torch(U) else U, dtype=torch.float32).to(self.device).contiguous()
        if self.q.shape != self.u.shape or self.q.dim() < 2:
            raise ValueError("Q and U cubes must have the same (m, ...) shape")
        self.m = int(self.q.shape[0])
        self.pix = tuple(self.q.shape[1:])
        self.npix = int(np.prod(self.pix))

    def slab(self, p0, P, ld=None):
        ld = (2 * self.m + 7) // 8 * 8 if ld is None else ld
        out = torch.zeros((P, ld), dtype=torch.float32, device=self.device)
        valid = torch.empty((P, self.m), dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.csr_pack_slab(dv.tptr(self.q), dv.tptr(self.u), self.npix, self.m, int(p0), int(P), dv.tptr(out),
                                          ld, dv.tptr(valid), dv.stream_ptr()), "csr_pack_slab")
        return out, valid

    def channel_stats(self, which="Q", x0=None, xn=None, y0=None, yn=None):
        """(nansum[m], nanstd[m]) of the chosen plane stack over the image window (utils/utilities.py:68-110)"""
        if len(self.pix) != 2:
            raise ValueError("channel statistics need an (m, Y, X) cube")
        Y, X = self.pix
        x0, y0 = 0 if x0 is None else x0, 0 if y0 is None else y0
        xn, yn = X if xn is None else xn, Y if yn is None else yn
        stats = torch.empty((self.m, 2), dtype=torch.float32, device=self.device)
        img = self.q if which == "Q" else self.u
        _lib.check(self.lib.csr_channel_stats(dv.tptr(img), self.m, Y, X, x0, xn, y0, yn, dv.tptr(stats), dv.stream_ptr()),
                   "csr_channel_stats")
        return stats[:, 0], stats[:, 1]

    def new_output(self, n):
        """(2, n, *pix) fp32 zeros on the device: real and imaginary planes of a Faraday-depth cube"""
        return torch.zeros((2, n) + self.pix, dtype=torch.float32, device=self.device)

    def store(self, planar, n, p0, out):
        P = planar.shape[0]
        _lib.check(self.lib.csr_unpack_slab(dv.tptr(planar), planar.shape[1], int(n), self.npix, int(p0), P, dv.tptr(out[0]),
                                            dv.tptr(out[1]), dv.stream_ptr()), "csr_unpack_slab")
        return out
