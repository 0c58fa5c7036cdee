// pm_fits.cuh — FITS image payload -> device floats: the data format on the INPUT side of the hot path.
// The reference reads every frame with astropy (`fits.getdata` + `byteswap().newbyteorder()`, Image.py:132-134) and
// hands the analysis `image.astype(np.float64)` (Image.py:148).  A FITS payload is big-endian BITPIX-typed pixels
// with physical value = BZERO + BSCALE * stored (the bundled SDSS frames are BITPIX 16 with BZERO = 32768 — the
// unsigned-16 convention — or BITPIX -32).  Decoding on the device lets the host ship the payload as it lies in the
// file: 2 bytes per pixel instead of 4 for the 16-bit frames, on a path that is PCIe bound end to end.
#pragma once
#include "pm_platform.cuh"

namespace pmk {

PM_HOSTDEV int fits_bytes_per_pixel(int bitpix) {
    return bitpix == 8 ? 1 : bitpix == 16 ? 2 : (bitpix == 32 || bitpix == -32) ? 4 : bitpix == -64 ? 8 : 0;
}
PM_HOSTDEV unsigned bswap32(unsigned v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }

// physical value of pixel i of the payload, float64: stored * bscale + bzero (exact for the unsigned-integer
// convention and for unscaled data; the multiply-add is done in float64)
PM_HOSTDEV double fits_pixel(const unsigned char* raw, size_t i, int bitpix, double bzero, double bscale) {
    double stored;
    if (bitpix == 8) stored = (double)raw[i];
    else if (bitpix == 16) stored = (double)(short)(unsigned short)(((unsigned)raw[2 * i] << 8) | raw[2 * i + 1]);
    else if (bitpix == 32 || bitpix == -32) {
        const unsigned char* p = raw + 4 * i;
        const unsigned u = ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3];
        if (bitpix == 32) stored = (double)(int)u;
        else { float f; memcpy(&f, &u, 4); stored = (double)f; }
    } else {
        const unsigned char* p = raw + 8 * i;
        unsigned long long u = 0;
        for (int k = 0; k < 8; k++) u = (u << 8) | p[k];
        memcpy(&stored, &u, 8);
    }
    if (bscale == 1.0 && bzero == 0.0) return stored;
    return stored * bscale + bzero;
}

#if !defined(PM_EMULATE)
// grid-stride; the two common layouts take 16-byte loads (8 x int16 / 4 x float32 per thread-step)
template <typename Out>
__global__ void pm_fits_decode_kernel(const unsigned char* __restrict__ raw, int bitpix, double bzero, double bscale,
                                      size_t count, Out* __restrict__ out) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    const bool plain = bscale == 1.0 && bzero == 0.0;
    if (bitpix == 16 && (((uintptr_t)raw) & 15) == 0) {
        const size_t nvec = count / 8;
        const uint4* v = (const uint4*)raw;
        for (size_t k = tid; k < nvec; k += nthreads) {
            const uint4 q = __ldg(v + k);
            const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // memory order b0 b1 b2 b3 = two big-endian int16: (b0 b1), (b2 b3); little-endian register = b3 b2 b1 b0
                const short s0 = (short)__byte_perm(w[j], 0, 0x4401), s1 = (short)__byte_perm(w[j], 0, 0x4423);
                const double d0 = plain ? (double)s0 : (double)s0 * bscale + bzero;
                const double d1 = plain ? (double)s1 : (double)s1 * bscale + bzero;
                out[8 * k + 2 * j] = (Out)d0;
                out[8 * k + 2 * j + 1] = (Out)d1;
            }
        }
        for (size_t i = nvec * 8 + tid; i < count; i += nthreads) out[i] = (Out)fits_pixel(raw, i, bitpix, bzero, bscale);
        return;
    }
    if (bitpix == -32 && (((uintptr_t)raw) & 15) == 0) {
        const size_t nvec = count / 4;
        const uint4* v = (const uint4*)raw;
        for (size_t k = tid; k < nvec; k += nthreads) {
            const uint4 q = __ldg(v + k);
            const unsigned w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const float f = __uint_as_float(__byte_perm(w[j], 0, 0x0123));
                out[4 * k + j] = plain ? (Out)f : (Out)((double)f * bscale + bzero);
            }
        }
        for (size_t i = nvec * 4 + tid; i < count; i += nthreads) out[i] = (Out)fits_pixel(raw, i, bitpix, bzero, bscale);
        return;
    }
    for (size_t i = tid; i < count; i += nthreads) out[i] = (Out)fits_pixel(raw, i, bitpix, bzero, bscale);
}
#endif

}  // namespace pmk
