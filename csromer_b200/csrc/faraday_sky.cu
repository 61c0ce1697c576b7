// faraday_sky.cu -- Galactic rotation-measure foreground of an image, on the device.
//
// Reference: faraday_sky/faraday_sky.py:62-141 -- FaradaySky.galactic_rm_image walks every image pixel through
// astropy.wcs (pixel -> sky), astropy.coordinates (ICRS -> Galactic) and astropy_healpix (sky -> HEALPix index) and gathers
// the Hutschenreuter et al. Faraday-sky map there; the per-pixel values are what Dataset.subtract_galacticrm
// (base/dataset.py:377-381; csr_derotate here) removes before the reconstruction.  One thread per pixel does the same chain
// in fp64:
//   pixel -> intermediate world coordinates (CRPIX / CD) -> native spherical (zenithal TAN or SIN deprojection) -> celestial
//   (CRVAL as the native pole, LONPOLE = 180 deg)  [FITS WCS papers I / II]
//   -> Galactic (l, b)  [rotation with the J2000 pole (192.85948, 27.12825) deg and l_NCP = 122.93192 deg]
//   -> HEALPix pixel, RING or NESTED scheme  [Gorski et al. 2005, ang2pix]  -> map value (nearest pixel, the reference default).
// astropy and astropy_healpix are third-party packages absent from the reference tree and from this image: parity with
// THEM is unpinned; the chain is pinned by known HEALPix answers and an independent pix2ang round trip (tests).
#include <math.h>

#include "plan.cuh"

namespace csr {

__device__ __forceinline__ long long healpix_ring(int nside, double z, double phi) {
    const double kPi = 3.14159265358979323846;
    const long long ns = nside, npix = 12 * ns * ns;
    const double za = fabs(z);
    double tt = fmod(phi, 2.0 * kPi);
    if (tt < 0.0) tt += 2.0 * kPi;
    tt /= 0.5 * kPi;  // in [0, 4)
    if (za <= 2.0 / 3.0) {  // equatorial belt
        const double t1 = ns * (0.5 + tt), t2 = ns * z * 0.75;
        const long long jp = (long long)floor(t1 - t2), jm = (long long)floor(t1 + t2);
        const long long ir = ns + 1 + jp - jm;  // ring counted from z = 2/3
        const long long kshift = 1 - (ir & 1);
        long long ip = (jp + jm - ns + kshift + 1) / 2;
        ip = ((ip % (4 * ns)) + 4 * ns) % (4 * ns);
        return 2 * ns * (ns - 1) + (ir - 1) * 4 * ns + ip;
    }
    const double tp = tt - floor(tt), tmp = ns * sqrt(3.0 * (1.0 - za));
    const long long jp = (long long)floor(tp * tmp), jm = (long long)floor((1.0 - tp) * tmp);
    const long long ir = jp + jm + 1;  // ring counted from the closest pole
    long long ip = (long long)floor(tt * ir);
    ip = ((ip % (4 * ir)) + 4 * ir) % (4 * ir);
    return z > 0.0 ? 2 * ir * (ir - 1) + ip : npix - 2 * ir * (ir + 1) + ip;
}

__device__ __forceinline__ unsigned long long spread_bits(unsigned int v) {  // bit i -> bit 2 i
    unsigned long long x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}

__device__ __forceinline__ long long healpix_nest(int nside, double z, double phi) {
    const double kPi = 3.14159265358979323846;
    const long long ns = nside;
    const double za = fabs(z);
    double tt = fmod(phi, 2.0 * kPi);
    if (tt < 0.0) tt += 2.0 * kPi;
    tt /= 0.5 * kPi;
    int face;
    long long ix, iy;
    if (za <= 2.0 / 3.0) {
        const double t1 = ns * (0.5 + tt), t2 = ns * z * 0.75;
        const long long jp = (long long)floor(t1 - t2), jm = (long long)floor(t1 + t2);
        const long long ifp = jp / ns, ifm = jm / ns;
        face = ifp == ifm ? (int)(ifp & 3) + 4 : (ifp < ifm ? (int)(ifp & 3) : (int)(ifm & 3) + 8);
        ix = jm & (ns - 1);
        iy = ns - (jp & (ns - 1)) - 1;
    } else {
        int ntt = (int)floor(tt);
        if (ntt >= 4) ntt = 3;
        const double tp = tt - ntt, tmp = ns * sqrt(3.0 * (1.0 - za));
        long long jp = (long long)floor(tp * tmp), jm = (long long)floor((1.0 - tp) * tmp);
        if (jp >= ns) jp = ns - 1;
        if (jm >= ns) jm = ns - 1;
        if (z >= 0.0) {
            face = ntt;
            ix = ns - jm - 1;
            iy = ns - jp - 1;
        } else {
            face = ntt + 8;
            ix = jp;
            iy = jm;
        }
    }
    return (long long)face * ns * ns + (long long)(spread_bits((unsigned)ix) | (spread_bits((unsigned)iy) << 1));
}

__device__ __forceinline__ void icrs_to_galactic(double ra, double dec, double& l, double& b) {
    const double kD2R = 3.14159265358979323846 / 180.0;
    const double ra_ngp = 192.85948 * kD2R, dec_ngp = 27.12825 * kD2R, l_ncp = 122.93192 * kD2R;
    double sd, cd, sn, cn, sa, ca;
    sincos(dec, &sd, &cd);
    sincos(dec_ngp, &sn, &cn);
    sincos(ra - ra_ngp, &sa, &ca);
    b = asin(fmin(1.0, fmax(-1.0, sd * sn + cd * cn * ca)));
    l = l_ncp - atan2(cd * sa, sd * cn - cd * sn * ca);
}

__device__ __forceinline__ void gather(const float* mean, const float* sd, long long pix, float* out_mean, float* out_sd, size_t at) {
    out_mean[at] = mean[pix];
    if (sd && out_sd) out_sd[at] = sd[pix];
}

__global__ void galactic_rm_image_kernel(const float* __restrict__ map_mean, const float* __restrict__ map_std, int nside,
                                         int nested, csr_wcs wcs, int nx, int ny, int swap_axes, int galactic_frame,
                                         float* __restrict__ rm_mean, float* __restrict__ rm_std) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= nx || r >= ny) return;
    const double kD2R = 3.14159265358979323846 / 180.0;
    // FITS pixel coordinates are 1-based; output element [r, c] is pixel (x = c, y = r) -- or, with swap_axes, pixel
    // (x = r, y = c), which is what faraday_sky.py:117-123 evaluates (array_index_to_world takes NumPy index order)
    const double px = (swap_axes ? r : c) + 1.0 - wcs.crpix1, py = (swap_axes ? c : r) + 1.0 - wcs.crpix2;
    const double xi = (wcs.cd11 * px + wcs.cd12 * py) * kD2R, eta = (wcs.cd21 * px + wcs.cd22 * py) * kD2R;
    const double rr = sqrt(xi * xi + eta * eta);
    const double phi = rr > 0.0 ? atan2(xi, -eta) : 0.0;
    const double theta = wcs.proj == 1 ? acos(fmin(1.0, rr)) : atan2(1.0, rr);  // SIN : TAN
    double st, ct, s0, c0, sp, cp;
    sincos(theta, &st, &ct);
    sincos(wcs.crval2 * kD2R, &s0, &c0);
    sincos(phi - wcs.lonpole * kD2R, &sp, &cp);
    const double dec = asin(fmin(1.0, fmax(-1.0, st * s0 + ct * c0 * cp)));
    const double ra = wcs.crval1 * kD2R + atan2(-ct * sp, st * c0 - ct * s0 * cp);
    double l = ra, b = dec;
    if (!galactic_frame) icrs_to_galactic(ra, dec, l, b);
    const double z = sin(b);
    const long long pix = nested ? healpix_nest(nside, z, l) : healpix_ring(nside, z, l);
    gather(map_mean, map_std, pix, rm_mean, rm_std, (size_t)r * nx + c);
}

__global__ void healpix_lookup_kernel(const float* __restrict__ map_mean, const float* __restrict__ map_std, int nside, int nested,
                                      const double* __restrict__ lon, const double* __restrict__ lat, int icrs, int n,
                                      float* __restrict__ out_mean, float* __restrict__ out_std, long long* __restrict__ out_pix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double l = lon[i], b = lat[i];
    if (icrs) icrs_to_galactic(lon[i], lat[i], l, b);
    const long long pix = nested ? healpix_nest(nside, sin(b), l) : healpix_ring(nside, sin(b), l);
    if (out_pix) out_pix[i] = pix;
    if (map_mean && out_mean) gather(map_mean, map_std, pix, out_mean, out_std, i);
}

}  // namespace csr

using namespace csr;

extern "C" int csr_galactic_rm_image(const float* map_mean, const float* map_std, int nside, int nested, const csr_wcs* wcs, int nx,
                                     int ny, int swap_axes, int galactic_frame, float* rm_mean, float* rm_std, void* stream) {
    CSR_REQUIRE(map_mean && wcs && rm_mean && nside > 0 && nx > 0 && ny > 0, "csr_galactic_rm_image: bad argument");
    CSR_REQUIRE(wcs->proj == 0 || wcs->proj == 1, "csr_galactic_rm_image: projection must be TAN (0) or SIN (1)");
    CSR_REQUIRE(!nested || (nside & (nside - 1)) == 0, "csr_galactic_rm_image: the NESTED scheme needs nside = 2^k");
    dim3 grid((nx + 127) / 128, ny);
    galactic_rm_image_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(map_mean, map_std, nside, nested, *wcs, nx, ny, swap_axes,
                                                                     galactic_frame, rm_mean, rm_std);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_healpix_lookup(const float* map_mean, const float* map_std, int nside, int nested, const double* lon,
                                  const double* lat, int icrs, int n, float* out_mean, float* out_std, long long* out_pix,
                                  void* stream) {
    CSR_REQUIRE(lon && lat && n > 0 && nside > 0 && (out_pix || (map_mean && out_mean)), "csr_healpix_lookup: bad argument");
    CSR_REQUIRE(!nested || (nside & (nside - 1)) == 0, "csr_healpix_lookup: the NESTED scheme needs nside = 2^k");
    healpix_lookup_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(map_mean, map_std, nside, nested, lon, lat, icrs, n,
                                                                            out_mean, out_std, out_pix);
    CSR_LAUNCHED();
    return CSR_OK;
}
