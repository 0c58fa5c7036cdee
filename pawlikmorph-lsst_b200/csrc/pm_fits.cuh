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

// stored -> physical value: a rounded multiply, then a rounded add (numpy's `a * bscale + bzero`; never an FMA)
PM_HOSTDEV double fits_scale(double stored, double bzero, double bscale) {
    if (bscale == 1.0 && bzero == 0.0) return stored;
#if defined(__CUDA_ARCH__)
    return __dadd_rn(__dmul_rn(stored, bscale), bzero);
#else
    volatile double prod = stored * bscale;      // keeps -ffp-contract from fusing the two on the host build
    return prod + bzero;
#endif
}

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
    return fits_scale(stored, bzero, bscale);
}

#if !defined(PM_EMULATE)
template <int NB> struct RawWords;             // NB payload bytes (one output vector's worth) as 32-bit words
template <> struct RawWords<2>  { unsigned w[1]; __device__ void load(const unsigned char* p) { w[0] = __ldg((const unsigned short*)p); } };
template <> struct RawWords<4>  { unsigned w[1]; __device__ void load(const unsigned char* p) { w[0] = __ldg((const unsigned*)p); } };
template <> struct RawWords<8>  { unsigned w[2]; __device__ void load(const unsigned char* p) { const uint2 q = __ldg((const uint2*)p); w[0] = q.x; w[1] = q.y; } };
template <> struct RawWords<16> { unsigned w[4]; __device__ void load(const unsigned char* p) { const uint4 q = __ldg((const uint4*)p); w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; } };
template <> struct RawWords<32> { unsigned w[8]; __device__ void load(const unsigned char* p) {
    const uint4 q = __ldg((const uint4*)p), r = __ldg((const uint4*)p + 1);
    w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w; w[4] = r.x; w[5] = r.y; w[6] = r.z; w[7] = r.w; } };

// stored value of pixel j of a vector held as little-endian words of big-endian pixels (byte swaps are one PRMT each)
template <int BITPIX> __device__ __forceinline__ double stored_of(const unsigned* w, int j) {
    if (BITPIX == 8) return (double)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
    if (BITPIX == 16) return (double)(short)__byte_perm(w[j >> 1], 0, (j & 1) ? 0x4423 : 0x4401);
    if (BITPIX == 32) return (double)(int)__byte_perm(w[j], 0, 0x0123);
    if (BITPIX == -32) return (double)__uint_as_float(__byte_perm(w[j], 0, 0x0123));
    return __hiloint2double((int)__byte_perm(w[2 * j], 0, 0x0123), (int)__byte_perm(w[2 * j + 1], 0, 0x0123));
}

template <int BITPIX> __device__ __forceinline__ int stored_int_of(const unsigned* w, int j) {
    if (BITPIX == 8) return (int)((w[j >> 2] >> (8 * (j & 3))) & 0xffu);
    return (int)(short)__byte_perm(w[j >> 1], 0, (j & 1) ? 0x4423 : 0x4401);
}

// One output vector = 16 bytes (4 floats / 2 doubles) per thread-step, so that stores AND loads of a warp are contiguous;
// PM_FITS_UNROLL vectors per thread with all loads issued before the first conversion.  HBM-bound byte work: algorithmic
// bytes per pixel = bytes_per_pixel read + sizeof(Out) written (6 for the 16-bit SDSS frames into float32).
#ifndef PM_FITS_UNROLL
#define PM_FITS_UNROLL 4
#endif
template <int BITPIX, typename Out>
__global__ void __launch_bounds__(256) pm_fits_decode_kernel(const unsigned char* __restrict__ raw, double bzero, double bscale,
                                                              size_t count, Out* __restrict__ out) {
    constexpr int BPP = BITPIX == 8 ? 1 : BITPIX == 16 ? 2 : BITPIX == -64 ? 8 : 4;
    constexpr int P = 16 / (int)sizeof(Out), NB = P * BPP, U = PM_FITS_UNROLL;
    const bool plain = bscale == 1.0 && bzero == 0.0;
    // integer pixels with BSCALE 1 and an integer BZERO (the unsigned conventions, and unscaled data): the sum is exact in
    // int32 and one int -> Out conversion rounds like the float64 route does
    const bool ints = (BITPIX == 8 || BITPIX == 16) && bscale == 1.0 && bzero == rint(bzero) && fabs(bzero) <= 1073741824.0;
    const int ibias = ints ? (int)bzero : 0;
    const size_t nvec = ((((uintptr_t)raw) % NB) == 0 && (((uintptr_t)out) & 15) == 0) ? count / P : 0;   // unaligned views: scalar path
    const size_t tile = (size_t)blockDim.x * U, stride = (size_t)gridDim.x * tile;
    for (size_t base = (size_t)blockIdx.x * tile + threadIdx.x; base < nvec; base += stride) {
        RawWords<NB> v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t k = base + (size_t)u * blockDim.x;
            if (k < nvec) v[u].load(raw + k * NB);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t k = base + (size_t)u * blockDim.x;
            if (k >= nvec) break;
            Out o[P];
            if (ints) {
#pragma unroll
                for (int j = 0; j < P; j++) o[j] = (Out)(stored_int_of<BITPIX>(v[u].w, j) + ibias);
            } else {
#pragma unroll
                for (int j = 0; j < P; j++) {
                    const double st = stored_of<BITPIX>(v[u].w, j);
                    o[j] = (Out)(plain ? st : __dadd_rn(__dmul_rn(st, bscale), bzero));
                }
            }
            if constexpr (sizeof(Out) == 4) ((float4*)out)[k] = make_float4(o[0], o[1], o[2], o[3]);
            else ((double2*)out)[k] = make_double2(o[0], o[1]);
        }
    }
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t i = nvec * P + tid; i < count; i += nthreads) out[i] = (Out)fits_pixel(raw, i, BITPIX, bzero, bscale);
}

// Cut-out decode: out[b, r, c] = pixel (origin[2b] + r, origin[2b+1] + c) of frame b (H x W pixels), npix x npix per frame —
// `Cutout2D(...).data` of Image.py:175-186 fused with the decode.  V consecutive pixels of one window row per thread
// (V | npix): pixel-sized loads that a warp coalesces, one 16-byte (float) / two 16-byte (double) stores.
template <int BITPIX> __device__ __forceinline__ double stored_at(const unsigned char* raw, size_t i) {
    if (BITPIX == 8) return (double)__ldg(raw + i);
    if (BITPIX == 16) return (double)(short)__byte_perm((unsigned)__ldg((const unsigned short*)raw + i), 0, 0x4401);
    if (BITPIX == 32) return (double)(int)__byte_perm(__ldg((const unsigned*)raw + i), 0, 0x0123);
    if (BITPIX == -32) return (double)__uint_as_float(__byte_perm(__ldg((const unsigned*)raw + i), 0, 0x0123));
    const uint2 q = __ldg((const uint2*)raw + i);
    return __hiloint2double((int)__byte_perm(q.x, 0, 0x0123), (int)__byte_perm(q.y, 0, 0x0123));
}

// grid: x over the V-pixel vectors of one window (32-bit index arithmetic, one division per thread), y striding the frames
template <int BITPIX, typename Out, int V>
__global__ void __launch_bounds__(256) pm_fits_window_kernel(const unsigned char* __restrict__ raw, double bzero, double bscale,
                                                              long long B, int H, int W, const int* __restrict__ origin, int npix,
                                                              Out* __restrict__ out) {
    const bool plain = bscale == 1.0 && bzero == 0.0;
    const bool ints = (BITPIX == 8 || BITPIX == 16) && bscale == 1.0 && bzero == rint(bzero) && fabs(bzero) <= 1073741824.0;
    const int ibias = ints ? (int)bzero : 0;
    const unsigned rowv = (unsigned)(npix / V), per = (unsigned)npix * rowv;
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= per) return;
    const unsigned r = q / rowv, c = (q - r * rowv) * V;
    for (long long b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t src = ((size_t)b * H + (size_t)(__ldg(origin + 2 * b) + (int)r)) * W + (size_t)(__ldg(origin + 2 * b + 1) + (int)c);
        Out o[V];
        if (ints) {
            int si[V];
#pragma unroll
            for (int j = 0; j < V; j++) si[j] = BITPIX == 8 ? (int)__ldg(raw + src + j) : (int)(short)__byte_perm((unsigned)__ldg((const unsigned short*)raw + src + j), 0, 0x4401);
#pragma unroll
            for (int j = 0; j < V; j++) o[j] = (Out)(si[j] + ibias);
        } else {
            double st[V];
#pragma unroll
            for (int j = 0; j < V; j++) st[j] = stored_at<BITPIX>(raw, src + j);
#pragma unroll
            for (int j = 0; j < V; j++) o[j] = (Out)(plain ? st[j] : __dadd_rn(__dmul_rn(st[j], bscale), bzero));
        }
        Out* dst = out + ((size_t)b * per + q) * V;
        if constexpr (V == 4 && sizeof(Out) == 4) *(float4*)dst = make_float4(o[0], o[1], o[2], o[3]);
        else if constexpr (V == 4) { ((double2*)dst)[0] = make_double2(o[0], o[1]); ((double2*)dst)[1] = make_double2(o[2], o[3]); }
        else dst[0] = o[0];
    }
}

// BITPIX 16 with an integer BZERO (the SDSS frames): eight pixels per thread from five ALIGNED 32-bit words — a window row starts at
// any even byte, so the 16 bytes a thread needs are funnel-shifted out of the words around them — and two 16-byte stores.
template <typename Out>
__global__ void __launch_bounds__(256) pm_fits_window16_kernel(const unsigned char* __restrict__ raw, int ibias, long long B, int H, int W,
                                                                const int* __restrict__ origin, int npix, Out* __restrict__ out) {
    const unsigned rowv = (unsigned)(npix / 8), per = (unsigned)npix * rowv;
    const unsigned q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= per) return;
    const unsigned r = q / rowv, c = (q - r * rowv) * 8;
    const size_t total_bytes = (size_t)B * H * W * 2;
    for (long long b = blockIdx.y; b < B; b += gridDim.y) {
        const size_t src = ((size_t)b * H + (size_t)(__ldg(origin + 2 * b) + (int)r)) * W + (size_t)(__ldg(origin + 2 * b + 1) + (int)c);
        const size_t byte = 2 * src;
        const unsigned* wp = (const unsigned*)(raw + (byte & ~(size_t)3));
        const unsigned sh = (unsigned)(byte & 2) * 8u;
        unsigned w[5];
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = __ldg(wp + k);
        // the fifth word is only needed for a row piece that starts in the middle of a word, and then only its low half: it is not
        // read when it would start beyond the payload (a window ending on the last pixel of the last frame)
        w[4] = (sh && (byte & ~(size_t)3) + 16 < total_bytes) ? __ldg(wp + 4) : 0u;
        unsigned p[4];
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = __funnelshift_r(w[k], w[k + 1], sh);
        Out o[8];
#pragma unroll
        for (int j = 0; j < 8; j++) o[j] = (Out)(stored_int_of<16>(p, j) + ibias);
        Out* dst = out + ((size_t)b * per + q) * 8;
        if constexpr (sizeof(Out) == 4) {
            ((float4*)dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
            ((float4*)dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) ((double2*)dst)[j] = make_double2(o[2 * j], o[2 * j + 1]);
        }
    }
}

template <typename Out, int V>
inline void fits_window_launch(const unsigned char* raw, int bitpix, double bzero, double bscale, long long B, int H, int W,
                               const int* origin, int npix, Out* out, unsigned max_ctas, cudaStream_t st) {
    const unsigned per = (unsigned)((size_t)npix * npix / V), gx = (per + 255u) / 256u;
    unsigned gy = gx ? max_ctas / gx : 1u;
    gy = gy < 1u ? 1u : gy;
    gy = (long long)gy > B ? (unsigned)B : gy;
    gy = gy > 65535u ? 65535u : gy;
    const dim3 grid(gx, gy, 1);
    if (V == 4 && bitpix == 16 && npix % 8 == 0 && bscale == 1.0 && bzero == rint(bzero) && fabs(bzero) <= 1073741824.0 &&
        (((uintptr_t)raw) & 3) == 0) {
        const unsigned per8 = (unsigned)((size_t)npix * npix / 8), gx8 = (per8 + 255u) / 256u;
        unsigned gy8 = gx8 ? max_ctas / gx8 : 1u;
        gy8 = gy8 < 1u ? 1u : gy8;
        gy8 = (long long)gy8 > B ? (unsigned)B : gy8;
        gy8 = gy8 > 65535u ? 65535u : gy8;
        pm_fits_window16_kernel<Out><<<dim3(gx8, gy8, 1), 256, 0, st>>>(raw, (int)bzero, B, H, W, origin, npix, out);
        return;
    }
    switch (bitpix) {
        case 8:   pm_fits_window_kernel<8, Out, V><<<grid, 256, 0, st>>>(raw, bzero, bscale, B, H, W, origin, npix, out); break;
        case 16:  pm_fits_window_kernel<16, Out, V><<<grid, 256, 0, st>>>(raw, bzero, bscale, B, H, W, origin, npix, out); break;
        case 32:  pm_fits_window_kernel<32, Out, V><<<grid, 256, 0, st>>>(raw, bzero, bscale, B, H, W, origin, npix, out); break;
        case -32: pm_fits_window_kernel<-32, Out, V><<<grid, 256, 0, st>>>(raw, bzero, bscale, B, H, W, origin, npix, out); break;
        default:  pm_fits_window_kernel<-64, Out, V><<<grid, 256, 0, st>>>(raw, bzero, bscale, B, H, W, origin, npix, out); break;
    }
}

template <typename Out>
inline void fits_decode_launch(const unsigned char* raw, int bitpix, double bzero, double bscale, size_t count, Out* out,
                               unsigned max_ctas, cudaStream_t st) {
    constexpr int P = 16 / (int)sizeof(Out);
    const size_t want = (count / P + 256 * PM_FITS_UNROLL - 1) / (256 * PM_FITS_UNROLL);
    const unsigned grid = (unsigned)(want < max_ctas ? (want > 0 ? want : 1) : max_ctas);
    switch (bitpix) {
        case 8:   pm_fits_decode_kernel<8, Out><<<grid, 256, 0, st>>>(raw, bzero, bscale, count, out); break;
        case 16:  pm_fits_decode_kernel<16, Out><<<grid, 256, 0, st>>>(raw, bzero, bscale, count, out); break;
        case 32:  pm_fits_decode_kernel<32, Out><<<grid, 256, 0, st>>>(raw, bzero, bscale, count, out); break;
        case -32: pm_fits_decode_kernel<-32, Out><<<grid, 256, 0, st>>>(raw, bzero, bscale, count, out); break;
        default:  pm_fits_decode_kernel<-64, Out><<<grid, 256, 0, st>>>(raw, bzero, bscale, count, out); break;
    }
}
#endif

}  // namespace pmk
