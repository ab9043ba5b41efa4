// image_kernels.cuh -- device side of the image-domain ("Gabor") detector, pick compaction and the instantaneous
// frequency helper.  Reference arithmetic replaced (file:line under /root/reference/):
//   src/das4whales/improcess.py:23-63   scale_pixels / trace2image (min-max scaling of envelope / sigma)
//   src/das4whales/improcess.py:395-421 binning = torchvision Resize (bilinear, antialias) -> k_resize_aa_{h,v}
//   scripts/main_gabordetect.py:109,135 cv2.filter2D(img, CV_64F, gabor) (correlation, BORDER_REFLECT_101, centre anchor)
//                                       + the thresholds of :123-124,:136-137 -> k_filter2d
//   scripts/main_gabordetect.py:166-169 binning(mask, 10, 10) and apply_smooth_mask (improcess.py:424-454: array * mask)
//                                       -> k_mask_upsample_mul
//   src/das4whales/dsp.py:830-856       instant_freq: diff(unwrap(angle(hilbert))) / 2 pi * fs -> k_inst_freq
//   src/das4whales/detect.py:192,271    the index arrays find_peaks returns -> k_peaks_count / k_peaks_fill
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace d4w {

// ---------------------------------------------------------------------------------- global min / max, scaling
// order-preserving float <-> unsigned mapping so that atomicMin / atomicMax work on floats of either sign
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

static __global__ void k_minmax_init(unsigned* mm) { mm[0] = 0xffffffffu; mm[1] = 0u; }

static __global__ void __launch_bounds__(256)
k_minmax(const float* __restrict__ x, size_t n, unsigned* __restrict__ mm) {
    float lo = INFINITY, hi = -INFINITY;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ float s_lo[8], s_hi[8];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_lo[w] = lo; s_hi[w] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { lo = fminf(lo, s_lo[i]); hi = fmaxf(hi, s_hi[i]); }
        atomicMin(mm, f2ord(lo));
        atomicMax(mm + 1, f2ord(hi));
    }
}

// y = (x - min) / (max - min) * mul   (improcess.py:40, :62)
static __global__ void __launch_bounds__(256)
k_scale_pixels(const float* __restrict__ x, float* __restrict__ y, size_t n, const unsigned* __restrict__ mm, float mul) {
    const float lo = ord2f(mm[0]), hi = ord2f(mm[1]);
    const float inv = 1.0f / (hi - lo);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (x[i] - lo) * inv * mul;
}

// y = 20 log10(x / max(x))   (dsp.get_spectrogram, dsp.py:76); mm[1] holds the maximum in the ordered-uint encoding
static __global__ void __launch_bounds__(256)
k_db_re_max(const float* __restrict__ x, float* __restrict__ y, size_t n, const unsigned* __restrict__ mm) {
    const float inv = 1.0f / ord2f(mm[1]);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = 20.0f * log10f(x[i] * inv);
}

// ---------------------------------------------------------------------------------- antialiased bilinear resize
// One output index of ATen's separable `upsample_bilinear2d_aa` (align_corners = False, size given):
//   scale = in / out; support = max(scale, 1); center = scale * (i + 0.5); xmin = max(int(center - support + 0.5), 0);
//   xsize = min(int(center + support + 0.5), in) - xmin; w_j = tri((j + xmin - center + 0.5) / max(scale, 1)) / sum.
struct AaSpan { int xmin, xsize; float center, invscale; };
__host__ __device__ inline AaSpan aa_span(int i, int in_size, int out_size) {
    const float scale = (float)in_size / (float)out_size;
    const float support = scale >= 1.f ? scale : 1.f;
    AaSpan s;
    s.center = scale * ((float)i + 0.5f);
    s.invscale = scale >= 1.f ? 1.f / scale : 1.f;
    int xmin = (int)(s.center - support + 0.5f);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(s.center + support + 0.5f);
    if (xmax > in_size) xmax = in_size;
    s.xmin = xmin; s.xsize = xmax - xmin;
    return s;
}
__host__ __device__ inline float aa_tri(float x) { x = fabsf(x); return x < 1.f ? 1.f - x : 0.f; }

// horizontal pass: in [h][iw] -> out [h][ow]
static __global__ void __launch_bounds__(256)
k_resize_aa_h(const float* __restrict__ in, int h, int iw, float* __restrict__ out, int ow) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (ox >= ow) return;
    const AaSpan s = aa_span(ox, iw, ow);
    const float* r = in + (size_t)y * iw + s.xmin;
    float tot = 0.f, acc = 0.f;
    for (int j = 0; j < s.xsize; ++j) {
        const float w = aa_tri(((float)(j + s.xmin) - s.center + 0.5f) * s.invscale);
        tot += w; acc += w * r[j];
    }
    out[(size_t)y * ow + ox] = tot != 0.f ? acc / tot : 0.f;
}
// vertical pass: in [ih][w] -> out [oh][w]
static __global__ void __launch_bounds__(256)
k_resize_aa_v(const float* __restrict__ in, int ih, int w, float* __restrict__ out, int oh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (x >= w) return;
    const AaSpan s = aa_span(oy, ih, oh);
    const float* c = in + (size_t)s.xmin * w + x;
    float tot = 0.f, acc = 0.f;
    for (int j = 0; j < s.xsize; ++j) {
        const float wt = aa_tri(((float)(j + s.xmin) - s.center + 0.5f) * s.invscale);
        tot += wt; acc += wt * c[(size_t)j * w];
    }
    out[(size_t)oy * w + x] = tot != 0.f ? acc / tot : 0.f;
}

// ---------------------------------------------------------------------------------- filter2D (correlation)
// out[y][x] = sum_{ky,kx} src(y + ky - ay, x + kx - ax) * K[ky][kx], borders by BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba),
// anchor (ay, ax) = (kh/2, kw/2) as cv2.filter2D's default.  `in_thr` (if not NaN) binarises the source on the fly
// (src = in > in_thr ? 1 : 0, main_gabordetect.py:124,135); `out_thr` (if not NaN) binarises the result (:124, :137).
// Tile: FT_TY x FT_TX outputs per CTA, every thread FT_R rows x 4 consecutive columns with the source window sliding
// through registers; kernel rows are padded to a multiple of 4 taps with zeros.
constexpr int FT_TX = 128, FT_TY = 16, FT_R = 2;
constexpr int FT_THREADS = (FT_TX / 4) * (FT_TY / FT_R);      // 256

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

static __global__ void __launch_bounds__(FT_THREADS)
k_filter2d(const float* __restrict__ in, int h, int w, const float* __restrict__ K, int kh, int kw, float in_thr, float out_thr,
           int border, float* __restrict__ out) {
    extern __shared__ __align__(16) float fsm[];
    const int kwp = (kw + 3) & ~3;                      // padded taps per kernel row
    const int tw = FT_TX + kwp;                         // tile width (multiple of 4)
    const int th = FT_TY + kh - 1;
    float* sK = fsm;                                    // [kh][kwp]
    float* sI = fsm + (size_t)kh * kwp;                 // [th][tw]
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * FT_TX, y0 = blockIdx.y * FT_TY;
    const int ay = kh / 2, ax = kw / 2;
    for (int i = tid; i < kh * kwp; i += FT_THREADS) {
        const int ky = i / kwp, kx = i - ky * kwp;
        sK[i] = kx < kw ? K[ky * kw + kx] : 0.f;
    }
    const bool bin = !isnan(in_thr);
    for (int i = tid; i < th * tw; i += FT_THREADS) {
        const int ty = i / tw, tx = i - ty * tw;
        int sy = y0 + ty - ay, sx = x0 + tx - ax;
        float v = 0.f;
        if (border == 0) { sy = reflect101(sy, h); sx = reflect101(sx, w); }
        if (sy >= 0 && sy < h && sx >= 0 && sx < w) {          // border == 1: zeros outside (scipy.signal.correlate 'same')
            v = in[(size_t)sy * w + sx];
            if (bin) v = v > in_thr ? 1.f : 0.f;
        }
        sI[i] = v;
    }
    __syncthreads();
    const int cx = (tid % (FT_TX / 4)) * 4, cy = (tid / (FT_TX / 4)) * FT_R;
    float acc[FT_R][4];
#pragma unroll
    for (int r = 0; r < FT_R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
        const float* krow = sK + ky * kwp;
        float4 cur[FT_R];
#pragma unroll
        for (int r = 0; r < FT_R; ++r) cur[r] = *reinterpret_cast<const float4*>(sI + (size_t)(cy + r + ky) * tw + cx);
        for (int kx = 0; kx < kwp; kx += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(krow + kx);
#pragma unroll
            for (int r = 0; r < FT_R; ++r) {
                const float4 nxt = *reinterpret_cast<const float4*>(sI + (size_t)(cy + r + ky) * tw + cx + kx + 4);
                const float wv[8] = {cur[r].x, cur[r].y, cur[r].z, cur[r].w, nxt.x, nxt.y, nxt.z, nxt.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    acc[r][j] = fmaf(wv[j], kv.x, fmaf(wv[j + 1], kv.y, fmaf(wv[j + 2], kv.z, fmaf(wv[j + 3], kv.w, acc[r][j]))));
                cur[r] = nxt;
            }
        }
    }
    const bool obin = !isnan(out_thr);
#pragma unroll
    for (int r = 0; r < FT_R; ++r) {
        const int y = y0 + cy + r;
        if (y >= h) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = x0 + cx + j;
            if (x < w) out[(size_t)y * w + x] = obin ? (acc[r][j] > out_thr ? 1.f : 0.f) : acc[r][j];
        }
    }
}

// ---------------------------------------------------------------------------------- mask upsample + multiply
// out[c][t] = trace[c][t] * (bilinear_resize(mask)[c][t] != 0): improcess.binning of a boolean mask back to the trace
// size (torchvision resizes in float32 and casts back to bool, i.e. "any contributing pixel set") followed by
// improcess.apply_smooth_mask, which multiplies by the raw mask (improcess.py:452).  mask: float32 [mh][mw], 0 / 1.
static __global__ void __launch_bounds__(256)
k_mask_upsample_mul(const float* __restrict__ trace, int nx, int ns, const float* __restrict__ mask, int mh, int mw,
                    float* __restrict__ out, unsigned char* __restrict__ mask_out) {
    // the rows of the small mask that contribute to output row c (blockIdx.y) with a non-zero weight are the same for the
    // whole CTA: worked out once, then every thread only tests its <= 3 columns of those rows
    __shared__ int s_rows[8];
    __shared__ int s_nrows;
    const int c = blockIdx.y;
    if (threadIdx.x == 0) {
        const AaSpan sy = aa_span(c, mh, nx);
        int n = 0;
        for (int jy = 0; jy < sy.xsize && n < 8; ++jy)
            if (aa_tri(((float)(jy + sy.xmin) - sy.center + 0.5f) * sy.invscale) != 0.f) s_rows[n++] = sy.xmin + jy;
        s_nrows = n;
    }
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ns) return;
    const AaSpan sx = aa_span(t, mw, ns);
    bool on = false;
    for (int jx = 0; jx < sx.xsize; ++jx) {
        if (aa_tri(((float)(jx + sx.xmin) - sx.center + 0.5f) * sx.invscale) == 0.f) continue;
        for (int r = 0; r < s_nrows; ++r) on = on || (mask[(size_t)s_rows[r] * mw + sx.xmin + jx] != 0.f);
    }
    const size_t o = (size_t)c * ns + t;
    if (out) out[o] = on ? trace[o] : 0.f * trace[o];
    if (mask_out) mask_out[o] = on ? 1 : 0;
}

// ---------------------------------------------------------------------------------- instantaneous frequency
// fi[i] = wrap(angle(z[i+1]) - angle(z[i])) / (2 pi) * fs with numpy.unwrap's wrapping rule
// (dd = mod(d + pi, 2 pi) - pi, and +pi instead of -pi when d > 0), z = x + i H(x).
static __global__ void __launch_bounds__(256)
k_inst_freq(const float* __restrict__ x, const float* __restrict__ hx, int n, double fs, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const double pi = 3.14159265358979323846;
    const double a0 = atan2((double)hx[i], (double)x[i]), a1 = atan2((double)hx[i + 1], (double)x[i + 1]);
    const double d = a1 - a0;
    double dd = fmod(d + pi, 2.0 * pi);
    if (dd < 0) dd += 2.0 * pi;
    dd -= pi;
    if (dd == -pi && d > 0) dd = pi;
    // numpy.unwrap leaves differences smaller than the discontinuity (pi) untouched
    const double corrected = fabs(d) < pi ? d : dd;
    out[i] = (float)(corrected / (2.0 * pi) * fs);
}

// ---------------------------------------------------------------------------------- pick compaction
// flags [nx][ns] bytes (0 / 1) -> counts per row -> exclusive offsets -> ascending sample indices per row.
static __global__ void __launch_bounds__(256)
k_peaks_count(const unsigned char* __restrict__ flags, int ns, int* __restrict__ counts) {
    const size_t row = blockIdx.x;
    const unsigned char* r = flags + row * ns;
    int c = 0;
    const int nv = ((((uintptr_t)r) & 15) == 0) ? ns / 16 : 0;
    const uint4* rv = reinterpret_cast<const uint4*>(r);
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const uint4 u = rv[i];
        c += __popc(u.x & 0x01010101u) + __popc(u.y & 0x01010101u) + __popc(u.z & 0x01010101u) + __popc(u.w & 0x01010101u);
    }
    for (int i = nv * 16 + threadIdx.x; i < ns; i += blockDim.x) c += r[i] ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s[i]; counts[row] = t; }
}

// single-CTA exclusive scan of counts[n] into offsets[n + 1] (n rows <= a few 100 k)
static __global__ void __launch_bounds__(1024)
k_scan_offsets(const int* __restrict__ counts, int n, int* __restrict__ offsets) {
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n ? counts[i] : 0;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += u; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += s_w[w];
        const int carry = s_carry;
        if (i < n) offsets[i] = carry + wbase + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wbase + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = s_carry;
}

static __global__ void __launch_bounds__(256)
k_peaks_fill(const unsigned char* __restrict__ flags, int ns, const int* __restrict__ offsets, int* __restrict__ idx) {
    const size_t row = blockIdx.x;
    const unsigned char* r = flags + row * ns;
    int* dst = idx + offsets[row];
    if (offsets[row + 1] == offsets[row]) return;
    __shared__ int s_w[8];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    // tiles of 256 threads x 16 consecutive samples keep the per-row order
    for (int t0 = 0; t0 < ns; t0 += 256 * 16) {
        const int i0 = t0 + threadIdx.x * 16;
        unsigned bits = 0;
        for (int j = 0; j < 16; ++j) { const int i = i0 + j; if (i < ns && r[i]) bits |= 1u << j; }
        const int v = __popc(bits);
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += u; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        int wbase = 0;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wbase += s_w[w];
        int pos = s_base + wbase + inc - v;
        while (bits) { const int j = __ffs(bits) - 1; bits &= bits - 1; dst[pos++] = i0 + j; }
        __syncthreads();
        if (threadIdx.x == 255) s_base = pos;
        __syncthreads();
    }
}

}  // namespace d4w
