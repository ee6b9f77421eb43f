// HBM-bound DDM-const kernels: fused q_sample (K1), fused C/eps loss forward+backward (K2), sampler update (K3),
// EDM preconditioning edges, plus library bookkeeping (error string, launch counter).
// All are single-pass, 128-bit vectorised, grid sized as a multiple of the SM count.
#include "adm_internal.h"
#include <cuda_bf16.h>
#include <atomic>

namespace adm {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static inline int ew_grid(long long work_items, int threads, int per_sm) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = 1LL * num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

// ------------------------------------------------------------------------------------------------ K1 q_sample
// x_t = x0 + C*t + sqrt(t)*noise with C = -x0.  One float4 per thread per iteration; t is per-sample.
__global__ void __launch_bounds__(256) qsample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                      const float* __restrict__ t, float* __restrict__ xt,
                                                      long long batch, long long chw) {
    const long long total = batch * chw;
    const long long stride = 1LL * gridDim.x * blockDim.x;
    if ((chw & 3) == 0) {
        const long long n4 = total >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x0);
        const float4* e4 = reinterpret_cast<const float4*>(noise);
        float4* o4 = reinterpret_cast<float4*>(xt);
        for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n4; i += stride) {
            const long long b = (i << 2) / chw;
            const float tb = __ldg(t + b);
            const float st = sqrtf(tb);
            const float4 x = __ldcs(x4 + i), e = __ldcs(e4 + i);
            float4 o;
            // keep the reference's evaluation order: (x0 + C*t) + sqrt(t)*noise
            // (no FMA contraction: bit-identical to the reference's separate mul / add kernels)
            o.x = __fadd_rn(__fadd_rn(x.x, __fmul_rn(-x.x, tb)), __fmul_rn(st, e.x));
            o.y = __fadd_rn(__fadd_rn(x.y, __fmul_rn(-x.y, tb)), __fmul_rn(st, e.y));
            o.z = __fadd_rn(__fadd_rn(x.z, __fmul_rn(-x.z, tb)), __fmul_rn(st, e.z));
            o.w = __fadd_rn(__fadd_rn(x.w, __fmul_rn(-x.w, tb)), __fmul_rn(st, e.w));
            __stcs(o4 + i, o);
        }
    } else {
        for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += stride) {
            const float tb = __ldg(t + i / chw);
            const float x = x0[i];
            xt[i] = __fadd_rn(__fadd_rn(x, __fmul_rn(-x, tb)), __fmul_rn(sqrtf(tb), noise[i]));
        }
    }
}

// ------------------------------------------------------------------------------------------------ K2 loss fwd+bwd
// grid = (blocks_per_sample, B).  Each block reduces its slice of one sample, then one atomicAdd per block.
__global__ void __launch_bounds__(256) ddm_loss_kernel(const float* __restrict__ cp, const float* __restrict__ ep,
                                                       const float* __restrict__ x0, const float* __restrict__ noise,
                                                       const float* __restrict__ t, float eps, int weighting,
                                                       int use_l1, float grad_scale, float* __restrict__ loss,
                                                       float* __restrict__ dcp, float* __restrict__ dep,
                                                       long long batch, long long chw) {
    const long long b = blockIdx.y;
    const float tb = __ldg(t + b);
    float w1 = 1.f, w2 = 1.f;
    if (weighting) {
        const float q = tb * tb - tb + 1.f;
        w1 = q / tb;
        w2 = q / (1.f - tb + eps);
    }
    const float inv_b = grad_scale / static_cast<float>(batch);
    // use_l1 flags: 1 = + w * mean|d| (image space, ddm_const.py:345-348), 2 = + w * sum|d| (latent,
    // ddm_const_2.py:561-564), both then / 2;  4 = + rec_w * sum|x_rec - x0| with rec_w = -log(t)/2 and
    // x_rec - x0 = -(t*d1 + sqrt(t)*d2) (ddm_const_2.py:566-568 with the sqrt(t) schedule of ddm_const.py:290-293).
    const float l1c = (use_l1 & 1) ? 1.f / static_cast<float>(chw) : ((use_l1 & 2) ? 1.f : 0.f);
    const float half = (use_l1 & 3) ? 0.5f : 1.f;
    // The reference multiplies the per-sample [B] reconstruction sums by rec_weight of shape [B, 1]
    // (ddm_const_2.py:565-568): that broadcasts to a [B, B] outer product whose total is (sum_i a_i) * (sum_j w_j), so
    // every sample is effectively weighted by W = sum_j -log(t_j)/2.  Reproduced as is.
    const bool vlb = (use_l1 & 4) != 0;
    float rec_w = 0.f;
    if (vlb) {
        __shared__ float w_part[8];
        float w = 0.f;
        for (long long j = threadIdx.x; j < batch; j += blockDim.x) w -= 0.5f * logf(__ldg(t + j));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if ((threadIdx.x & 31) == 0) w_part[threadIdx.x >> 5] = w;
        __syncthreads();
        for (int i = 0; i < (blockDim.x >> 5); ++i) rec_w += w_part[i];
    }
    const float sq_t = sqrtf(tb);
    float svlb = 0.f;
    const float* cpb = cp + b * chw;
    const float* epb = ep + b * chw;
    const float* xb = x0 + b * chw;
    const float* nb = noise + b * chw;
    float sse1 = 0.f, sse2 = 0.f, sae1 = 0.f, sae2 = 0.f;
    const long long stride = 1LL * gridDim.x * blockDim.x;
    const bool vec = (chw & 3) == 0;
    const long long cnt = vec ? (chw >> 2) : chw;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < cnt; i += stride) {
        float c[4], e[4], x[4], n[4];
        int m = 1;
        if (vec) {
            const float4 c4 = __ldcs(reinterpret_cast<const float4*>(cpb) + i);
            const float4 e4 = __ldcs(reinterpret_cast<const float4*>(epb) + i);
            const float4 x4 = __ldcs(reinterpret_cast<const float4*>(xb) + i);
            const float4 n4 = __ldcs(reinterpret_cast<const float4*>(nb) + i);
            c[0] = c4.x; c[1] = c4.y; c[2] = c4.z; c[3] = c4.w;
            e[0] = e4.x; e[1] = e4.y; e[2] = e4.z; e[3] = e4.w;
            x[0] = x4.x; x[1] = x4.y; x[2] = x4.z; x[3] = x4.w;
            n[0] = n4.x; n[1] = n4.y; n[2] = n4.z; n[3] = n4.w;
            m = 4;
        } else {
            c[0] = cpb[i]; e[0] = epb[i]; x[0] = xb[i]; n[0] = nb[i];
        }
        float g1[4], g2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < m) {
                const float d1 = c[j] - (-x[j]);  // target1 = C = -x0
                const float d2 = e[j] - n[j];     // target2 = noise
                sse1 += d1 * d1;
                sse2 += d2 * d2;
                sae1 += fabsf(d1);
                sae2 += fabsf(d2);
                const float s1 = (d1 > 0.f) - (d1 < 0.f), s2 = (d2 > 0.f) - (d2 < 0.f);
                g1[j] = inv_b * half * w1 * (2.f * d1 + l1c * s1);
                g2[j] = inv_b * half * w2 * (2.f * d2 + l1c * s2);
                if (vlb) {
                    const float u = tb * d1 + sq_t * d2;
                    const float su = (u > 0.f) - (u < 0.f);
                    svlb += fabsf(u);
                    g1[j] += inv_b * rec_w * su * tb;
                    g2[j] += inv_b * rec_w * su * sq_t;
                }
            }
        }
        if (dcp != nullptr) {
            if (vec) {
                __stcs(reinterpret_cast<float4*>(dcp + b * chw) + i, make_float4(g1[0], g1[1], g1[2], g1[3]));
                __stcs(reinterpret_cast<float4*>(dep + b * chw) + i, make_float4(g2[0], g2[1], g2[2], g2[3]));
            } else {
                dcp[b * chw + i] = g1[0];
                dep[b * chw + i] = g2[0];
            }
        }
    }
    float pv = rec_w * svlb;
    float part = half * (w1 * (sse1 + l1c * sae1) + w2 * (sse2 + l1c * sae2)) + pv;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        part += __shfl_xor_sync(0xffffffffu, part, o);
        pv += __shfl_xor_sync(0xffffffffu, pv, o);
    }
    __shared__ float warp_sums[8], warp_sums_v[8];
    if ((threadIdx.x & 31) == 0) { warp_sums[threadIdx.x >> 5] = part; warp_sums_v[threadIdx.x >> 5] = pv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f, sv = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) { s += warp_sums[i]; sv += warp_sums_v[i]; }
        if (vlb) atomicAdd(loss + batch + b, sv);
        atomicAdd(loss + b, s);
    }
}

// ------------------------------------------------------------------------------------------------ K3 sampler step
template <typename S>
__global__ void __launch_bounds__(256) sampler_step_kernel(const S* __restrict__ x, const float* __restrict__ cp,
                                                           const float* __restrict__ ep, S* __restrict__ xn,
                                                           double t_cur, double t_next, double clip, int do_clip,
                                                           int last, double scale_input, long long numel) {
    const S tc = static_cast<S>(t_cur), tn = static_cast<S>(t_next);
    const S sc = static_cast<S>(sqrt(t_cur)), sn = static_cast<S>(sqrt(t_next));
    const S lo = static_cast<S>(-clip), hi = static_cast<S>(clip);
    const long long stride = 1LL * gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < numel; i += stride) {
        const S c = static_cast<S>(cp[i]), e = static_cast<S>(ep[i]);
        S x0 = x[i] - c * tc - e * sc;
        if (do_clip) x0 = x0 < lo ? lo : (x0 > hi ? hi : x0);
        S v = x0 + c * tn + e * sn;
        if (last) {
            v = v < lo ? lo : (v > hi ? hi : v);
            if (scale_input != 1.0) v = v / static_cast<S>(scale_input);
            v = (v + static_cast<S>(1)) * static_cast<S>(0.5);
        }
        xn[i] = v;
    }
}

__global__ void __launch_bounds__(256) sampler_step_stoch_kernel(const float* __restrict__ x,
                                                                 const float* __restrict__ cp,
                                                                 const float* __restrict__ ep,
                                                                 const float* __restrict__ z, float* __restrict__ xn,
                                                                 float t, float s, float clip, int do_clip,
                                                                 long long numel) {
    const float st = sqrtf(t);
    const float sigma = sqrtf(s * (t - s) / t);
    const long long stride = 1LL * gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < numel; i += stride) {
        const float xi = x[i], e = ep[i];
        float x0 = xi - cp[i] * t - st * e;
        if (do_clip) x0 = fminf(fmaxf(x0, -clip), clip);
        const float c = -x0;
        const float mean = xi + c * (t - s) - c * t - s / st * e;
        xn[i] = mean + sigma * z[i];
    }
}

// ------------------------------------------------------------------------------------------------ EDM edges
__device__ __forceinline__ float edm_c_in(float s) { return 1.f / sqrtf((1.f - s) * (1.f - s) + s); }

// NCHW fp32 -> NHWC bf16 (ld_out channels, zero padded), scaled by c_in(sigma_b).  One thread per pixel.
__global__ void __launch_bounds__(256) unet_input_kernel(const float* __restrict__ x, const float* __restrict__ sigma,
                                                         int scalar_sigma, __nv_bfloat16* __restrict__ out, int n,
                                                         int c, int hw, int ld_out) {
    const long long total = 1LL * n * hw;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int b = static_cast<int>(i / hw), p = static_cast<int>(i % hw);
        const float cin = edm_c_in(__ldg(sigma + (scalar_sigma ? 0 : b)));
        __nv_bfloat16* o = out + i * ld_out;
        for (int ch = 0; ch < ld_out; ++ch) {
            const float v = ch < c ? cin * x[(1LL * b * c + ch) * hw + p] : 0.f;
            o[ch] = __float2bfloat16(v);
        }
    }
}

__global__ void __launch_bounds__(256) unet_output_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                          int ldf, const float* __restrict__ x,
                                                          const float* __restrict__ sigma, int scalar_sigma,
                                                          float* __restrict__ d1, float* __restrict__ d2, int n, int c,
                                                          int hw) {
    const long long total = 1LL * n * c * hw;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int p = static_cast<int>(i % hw);
        const int ch = static_cast<int>((i / hw) % c);
        const int b = static_cast<int>(i / (1LL * hw * c));
        const float s = __ldg(sigma + (scalar_sigma ? 0 : b));
        const float q = s * s - s + 1.f;
        const float c_skip1 = (s - 1.f) / q, c_skip2 = sqrtf(s) / q;
        const float c_out1 = sqrtf(s / q), c_out2 = (1.f - s) / sqrtf(q);
        const long long fi = (1LL * b * hw + p) * ldf + ch;
        const float xv = x[i];
        d1[i] = c_skip1 * xv + c_out1 * f1[fi];
        d2[i] = c_skip2 * xv + c_out2 * f2[fi];
    }
}

__global__ void __launch_bounds__(256) unet_output_bwd_kernel(const float* __restrict__ dd1,
                                                              const float* __restrict__ dd2,
                                                              const float* __restrict__ sigma,
                                                              __nv_bfloat16* __restrict__ df1,
                                                              __nv_bfloat16* __restrict__ df2, int n, int c, int hw,
                                                              int ld_out) {
    const long long total = 1LL * n * hw;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const int b = static_cast<int>(i / hw), p = static_cast<int>(i % hw);
        const float s = __ldg(sigma + b);
        const float q = s * s - s + 1.f;
        const float c_out1 = sqrtf(s / q), c_out2 = (1.f - s) / sqrtf(q);
        for (int ch = 0; ch < ld_out; ++ch) {
            float a = 0.f, d = 0.f;
            if (ch < c) {
                const long long si = (1LL * b * c + ch) * hw + p;
                a = c_out1 * dd1[si];
                d = c_out2 * dd2[si];
            }
            df1[i * ld_out + ch] = __float2bfloat16(a);
            df2[i * ld_out + ch] = __float2bfloat16(d);
        }
    }
}

// ------------------------------------------------------------------------------------------------ AugmentPipe warp (f-4)
// The whole geometric augmentation of ddm/augment.py:153-328 as the reference's DDM module configures it
// (ddm_const.py:179-180) in ONE kernel: x / y flips, reflect padding by the batch-wide margins, 2x upsampling with the sym6
// low-pass, bilinear sampling through the per-sample inverse affine map, sym6 low-pass + 2x decimation, crop.
// grid (AUG_BANDS, N): a CTA produces a band of output rows of one sample and evaluates only the rows of the sampled grid
// that band's decimation filter reads.  The padded and the upsampled images are never materialised: the upsampling filter is
// separable, so a bilinear sample of the upsampled image is sum_iy sum_ix wY[iy] wX[ix] P[iy][ix] over an 8 x 8 window of
// the (reflect-indexed) flipped source image, where wY / wX fold the two bilinear weights into the zero-stuffed 12-tap
// filter (6 non-zero taps per upsampled row).  Shared memory: image [C][H][W], grid band [C][rows][Gw], x-decimated band.
// theta [N][6] is the matrix handed to affine_grid (all compositions done on the host: a few floats per sample).
__constant__ float c_sym6[12] = {0.015404109327027373f, 0.0034907120842174702f, -0.11799011114819057f,
                                 -0.048311742585633f,   0.4910559419267466f,    0.787641141030194f,
                                 0.3379294217276218f,   -0.07263752278646252f,  -0.021060292512300564f,
                                 0.04472490177066578f,  0.0017677118642428036f, -0.007800708325034148f};
constexpr int AUG_BANDS = 4;
constexpr int AUG_PAD4 = 3;  // len(sym6) / 4

__device__ __forceinline__ int aug_reflect(int i, int n) {  // one reflection is enough: margins are <= n - 1
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// Combined weights of one axis for a bilinear sample at upsampled coordinate p (p0 = floor(p), weights w0 / w1 on p0 / p0 + 1):
// wgt[k] multiplies padded-image index base + k (k < 8, base = floor(p0 / 2) - 3); src[k] = its reflect-mapped source index,
// or -1 when the padded index falls outside the image (zeros).  The upsampling conv is a cross-correlation of the
// zero-stuffed image with the FLIPPED filter and padding 6: U[P] = sum over a = (P mod 2) + 2h of sym6[11 - a] * pad[(P + a - 6) / 2],
// which puts tap h of row p0 at k = h + (p0 mod 2) and tap h of row p0 + 1 at k = h + 1.  Every array index below is a
// compile-time constant (registers, no local memory).
__device__ __forceinline__ void aug_axis(int p0, float w0, float w1, int size_up, int size_pad, int margin, int size,
                                         float (&wgt)[8], int (&src)[8]) {
    constexpr float E[6] = {-0.007800708325034148f, 0.04472490177066578f, -0.07263752278646252f, 0.787641141030194f,
                            -0.048311742585633f, 0.0034907120842174702f};  // sym6[11 - 2h]
    constexpr float O[6] = {0.0017677118642428036f, -0.021060292512300564f, 0.3379294217276218f, 0.4910559419267466f,
                            -0.11799011114819057f, 0.015404109327027373f};  // sym6[10 - 2h]
    const bool par = (p0 & 1) != 0;
    if (p0 < 0 || p0 >= size_up) w0 = 0.f;          // grid_sample padding_mode = zeros
    if (p0 + 1 < 0 || p0 + 1 >= size_up) w1 = 0.f;
    const int base = (p0 >> 1) - 3;  // arithmetic shift = floor for negative p0
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // row p0 (parity par): even rows use the even taps E, odd rows the odd taps O; row p0 + 1 the other set
        const float f0_same = (k <= 5) ? (par ? 0.f : E[k <= 5 ? k : 0]) : 0.f;               // par == 0: tap k
        const float f0_shift = (k >= 1 && k <= 6) ? (par ? O[(k >= 1 && k <= 6) ? k - 1 : 0] : 0.f) : 0.f;  // par == 1: tap k - 1
        const float f1 = (k >= 1 && k <= 6) ? (par ? E[(k >= 1 && k <= 6) ? k - 1 : 0] : O[(k >= 1 && k <= 6) ? k - 1 : 0]) : 0.f;
        wgt[k] = w0 * (f0_same + f0_shift) + w1 * f1;
        const int i = base + k;
        src[k] = (i < 0 || i >= size_pad) ? -1 : aug_reflect(i - margin, size);
    }
}

__global__ void __launch_bounds__(256) augment_warp_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           const float* __restrict__ theta,
                                                           const int* __restrict__ flips, int C, int H, int W, int mx0,
                                                           int mx1, int my0, int my1) {
    extern __shared__ float aug_sm[];
    const int n = blockIdx.y;
    const int Gh = (H + 2 * AUG_PAD4) * 2, Gw = (W + 2 * AUG_PAD4) * 2;
    const int Hp = H + my0 + my1, Wp = W + mx0 + mx1;     // reflect-padded image
    const int Hu = 2 * Hp, Wu = 2 * Wp;                   // upsampled image the sampler reads
    const int rows_out = (H + AUG_BANDS - 1) / AUG_BANDS;
    const int h_lo = blockIdx.x * rows_out, h_hi = min(H, h_lo + rows_out);
    if (h_lo >= h_hi) return;
    // grid rows the band's y-decimation reads: 2 (h + pad4) + b - 5, b = 0 .. 11
    const int g_lo = max(0, 2 * (h_lo + AUG_PAD4) - 5), g_hi = min(Gh, 2 * (h_hi - 1 + AUG_PAD4) + 7);
    const int R = g_hi - g_lo;
    float* img = aug_sm;                                  // [C][H][W], flips applied
    float* grid = img + C * H * W;                        // [C][R][Gw]
    float* dec = grid + C * R * Gw;                       // [C][R][W]
    const int fx = flips[2 * n], fy = flips[2 * n + 1];
    for (int i = threadIdx.x; i < C * H * W; i += blockDim.x) {
        const int w = i % W, h = (i / W) % H, c = i / (W * H);
        img[i] = x[(static_cast<long long>(n) * C + c) * H * W + (fy ? H - 1 - h : h) * W + (fx ? W - 1 - w : w)];
    }
    __syncthreads();
    const float t00 = theta[6 * n], t01 = theta[6 * n + 1], t02 = theta[6 * n + 2];
    const float t10 = theta[6 * n + 3], t11 = theta[6 * n + 4], t12 = theta[6 * n + 5];
    // bilinear sampling (affine_grid + grid_sample, align_corners = False, zeros outside) of the upsampled image
    for (int i = threadIdx.x; i < R * Gw; i += blockDim.x) {
        const int r = i / Gw, gj = i - r * Gw, gi = g_lo + r;
        const float bx = (2.f * gj + 1.f) / Gw - 1.f, by = (2.f * gi + 1.f) / Gh - 1.f;
        const float gx = t00 * bx + t01 * by + t02, gy = t10 * bx + t11 * by + t12;
        const float px = ((gx + 1.f) * Wu - 1.f) * 0.5f, py = ((gy + 1.f) * Hu - 1.f) * 0.5f;
        // far outside: nothing to sample (also keeps the float -> int conversion in range)
        if (!(px > -2.f && px < Wu + 1.f && py > -2.f && py < Hu + 1.f)) {
            for (int c = 0; c < C; ++c) grid[(c * R + r) * Gw + gj] = 0.f;
            continue;
        }
        const float fx0 = floorf(px), fy0 = floorf(py);
        float wX[8], wY[8];
        int sX[8], sY[8];
        aug_axis(static_cast<int>(fx0), 1.f - (px - fx0), px - fx0, Wu, Wp, mx0, W, wX, sX);
        aug_axis(static_cast<int>(fy0), 1.f - (py - fy0), py - fy0, Hu, Hp, my0, H, wY, sY);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 8; ++ky) {
            if (sY[ky] < 0 || wY[ky] == 0.f) continue;
            float row[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int kx = 0; kx < 8; ++kx) {
                if (sX[kx] < 0) continue;
                const float* pix = img + sY[ky] * W + sX[kx];
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < C) row[c] += wX[kx] * pix[c * H * W];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[c] += wY[ky] * row[c];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < C) grid[(c * R + r) * Gw + gj] = acc[c];
    }
    __syncthreads();
    // low-pass + decimate along x (stride 2, padding 5), cropped to the W centre columns
    for (int i = threadIdx.x; i < C * R * W; i += blockDim.x) {
        const int w = i % W, r = (i / W) % R, c = i / (W * R);
        const float* row = grid + (c * R + r) * Gw;
        float a = 0.f;
#pragma unroll
        for (int b = 0; b < 12; ++b) {
            const int j = 2 * (w + AUG_PAD4) + b - 5;
            if (j >= 0 && j < Gw) a += c_sym6[b] * row[j];
        }
        dec[i] = a;
    }
    __syncthreads();
    // ... and along y, cropped to the H centre rows (this CTA's band)
    const int rows = h_hi - h_lo;
    for (int i = threadIdx.x; i < C * rows * W; i += blockDim.x) {
        const int w = i % W, h = h_lo + (i / W) % rows, c = i / (W * rows);
        float a = 0.f;
#pragma unroll
        for (int b = 0; b < 12; ++b) {
            const int j = 2 * (h + AUG_PAD4) + b - 5;
            if (j >= 0 && j < Gh) a += c_sym6[b] * dec[(c * R + (j - g_lo)) * W + w];
        }
        y[(static_cast<long long>(n) * C + c) * H * W + h * W + w] = a;
    }
}

// ------------------------------------------------------------------------------------------------ weight packing
// w [cout][cin][k][k] fp32 -> bf16 [cout][k*k][kpad]; source 2 channels start at pad64(c1).
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ w,
                                                               __nv_bfloat16* __restrict__ o, int cout, int c1, int c2,
                                                               int kk, int p1, int kpad,
                                                               const int* __restrict__ row_perm) {
    const unsigned total = static_cast<unsigned>(cout) * kk * kpad;  // < 2^31 (checked by the launcher): 32-bit divisions
    const int cin = c1 + c2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kc = static_cast<int>(i % kpad);
        const int tap = static_cast<int>((i / kpad) % kk);
        const int co = static_cast<int>(i / (static_cast<unsigned>(kpad) * kk));
        int ci = -1;
        if (kc < c1) ci = kc;
        else if (kc >= p1 && kc - p1 < c2) ci = c1 + kc - p1;
        const int src = row_perm != nullptr ? row_perm[co] : co;  // packed row co holds reference row src
        const float v = ci >= 0 ? w[(1LL * src * cin + ci) * kk + tap] : 0.f;
        o[i] = __float2bfloat16(v);
    }
}
// Batched row gather: row i of the table copies `row_bytes` bytes from src + table[i].x to dst + table[i].y (byte offsets).
// Re-derives, in ONE launch per optimizer step, every operand that is a row permutation of arena-resident parameters:
// the qkv projections stored as (q | k | v) x head x d instead of the reference's interleaved (head, d, {q,k,v}) rows.
template <typename T>
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          const longlong2* __restrict__ table, long long n_rows,
                                                          int units) {  // units of sizeof(T) per row
    const long long total = n_rows * units;
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < total; i += 1LL * gridDim.x * blockDim.x) {
        const long long r = i / units;
        const int u = static_cast<int>(i - r * units);
        const longlong2 t = table[r];
        reinterpret_cast<T*>(dst + t.y)[u] = reinterpret_cast<const T*>(src + t.x)[u];
    }
}

// ------------------------------------------------------------------------------------------------ dgrad weight shadow
// Batched 64 x 64 tile transpose of packed bf16 conv weights: [cout][tap][cin] -> [cin][ntaps-1-tap][cout], i.e. the
// K-major operand of the DATA-gradient GEMM (dX = dY * W^T with the 3x3 taps mirrored), so that dgrad runs through
// the fprop kernels.  One CTA per tile; `tiles` holds {src offset, dst offset, src row stride, dst row stride} in
// elements for every tile of every conv (built once by the host), so ONE launch re-derives the shadow of a whole arena.
__global__ void __launch_bounds__(256) transpose_tiles_kernel(const uint16_t* __restrict__ src,
                                                              uint16_t* __restrict__ dst,
                                                              const longlong4* __restrict__ tiles) {
    __shared__ uint32_t t32[64 * 33];  // 64 rows of 66 bf16 (odd word pitch: the column gather below is 2-way at worst)
    const longlong4 d = tiles[blockIdx.x];
    const uint16_t* s = src + d.x;
    uint16_t* o = dst + d.y;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int i = threadIdx.x + it * 256;
        const int r = i >> 3, c = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(s + r * d.z + c * 8);
        uint32_t* row = t32 + r * 33 + c * 4;
        row[0] = v.x; row[1] = v.y; row[2] = v.z; row[3] = v.w;
    }
    __syncthreads();
    const uint16_t* t16 = reinterpret_cast<const uint16_t*>(t32);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int i = threadIdx.x + it * 256;
        const int r = i >> 3, c = i & 7;  // output row r (a source column), output chunk c (source rows 8c .. 8c+7)
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t lo = t16[(c * 8 + 2 * j) * 66 + r], hi = t16[(c * 8 + 2 * j + 1) * 66 + r];
            w[j] = lo | (hi << 16);
        }
        *reinterpret_cast<uint4*>(o + r * d.w + c * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}
__global__ void __launch_bounds__(256) unpack_conv_wgrad_kernel(const float* __restrict__ g, float* __restrict__ o,
                                                                int cout, int c1, int c2, int kk, int p1, int kpad,
                                                                int accumulate, const int* __restrict__ row_perm) {
    const int cin = c1 + c2;
    const unsigned total = static_cast<unsigned>(cout) * cin * kk;  // < 2^31 (checked by the launcher): 32-bit divisions
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int tap = static_cast<int>(i % kk);
        const int ci = static_cast<int>((i / kk) % cin);
        const int co = static_cast<int>(i / (static_cast<unsigned>(kk) * cin));
        const int kc = ci < c1 ? ci : p1 + (ci - c1);
        const float v = g[(1LL * co * kk + tap) * kpad + kc];
        const long long dst = row_perm != nullptr ? ((1LL * row_perm[co] * cin + ci) * kk + tap) : i;
        o[dst] = accumulate ? o[dst] + v : v;
    }
}
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ d,
                                                            long long n) {
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x)
        d[i] = __float2bfloat16(s[i]);
}

}  // namespace adm

using namespace adm;

extern "C" {

const char* adm_last_error(void) { return get_error(); }
int adm_version(void) { return 100; }
long long adm_launch_count(void) { return g_launches.load(); }

int adm_qsample(const float* x0, const float* noise, const float* t, float* x_t, long long batch, long long chw,
                void* stream) {
    if (batch <= 0 || chw <= 0) { set_error("qsample: empty input"); return ADM_ERR_SHAPE; }
    const long long items = (chw & 3) == 0 ? batch * chw / 4 : batch * chw;
    qsample_kernel<<<ew_grid(items, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x0, noise, t, x_t, batch,
                                                                                          chw);
    ADM_CHECK_LAUNCH("qsample");
    return 0;
}

int adm_ddm_loss(const float* c_pred, const float* eps_pred, const float* x0, const float* noise, const float* t,
                 float eps, int weighting, int use_l1, float grad_scale, float* loss_per_sample, float* d_c_pred,
                 float* d_eps_pred, long long batch, long long chw, void* stream) {
    if (batch <= 0 || chw <= 0 || batch > 65535) { set_error("ddm_loss: bad batch %lld", batch); return ADM_ERR_SHAPE; }
    if ((d_c_pred == nullptr) != (d_eps_pred == nullptr)) { set_error("ddm_loss: pass both gradients or neither"); return ADM_ERR_SHAPE; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaMemsetAsync(loss_per_sample, 0, sizeof(float) * batch * ((use_l1 & 4) ? 2 : 1), s);
    const long long items = (chw & 3) == 0 ? chw / 4 : chw;
    long long bx = (items + 255) / 256;
    const long long cap = (8LL * num_sms() + batch - 1) / batch;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(batch));
    ddm_loss_kernel<<<grid, 256, 0, s>>>(c_pred, eps_pred, x0, noise, t, eps, weighting, use_l1, grad_scale,
                                         loss_per_sample, d_c_pred, d_eps_pred, batch, chw);
    ADM_CHECK_LAUNCH("ddm_loss");
    return 0;
}

int adm_sampler_step(const void* x, const float* c_pred, const float* eps_pred, void* x_next, double t_cur,
                     double t_next, double clip, int do_clip, int last, double scale_input, int state_f64,
                     long long numel, void* stream) {
    if (numel <= 0) { set_error("sampler_step: empty input"); return ADM_ERR_SHAPE; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = ew_grid(numel, 256, 8);
    if (state_f64)
        sampler_step_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(x), c_pred, eps_pred,
                                                         static_cast<double*>(x_next), t_cur, t_next, clip, do_clip,
                                                         last, scale_input, numel);
    else
        sampler_step_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), c_pred, eps_pred,
                                                        static_cast<float*>(x_next), t_cur, t_next, clip, do_clip,
                                                        last, scale_input, numel);
    ADM_CHECK_LAUNCH("sampler_step");
    return 0;
}

int adm_sampler_step_stochastic(const float* x, const float* c_pred, const float* eps_pred, const float* z,
                                float* x_next, double t_cur, double s, double clip, int do_clip, long long numel,
                                void* stream) {
    if (numel <= 0) { set_error("sampler_step_stochastic: empty input"); return ADM_ERR_SHAPE; }
    sampler_step_stoch_kernel<<<ew_grid(numel, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, c_pred, eps_pred, z, x_next, static_cast<float>(t_cur), static_cast<float>(s), static_cast<float>(clip),
        do_clip, numel);
    ADM_CHECK_LAUNCH("sampler_step_stochastic");
    return 0;
}

int adm_unet_input(const float* x_nchw, const float* sigma, int sigma_is_scalar, void* x_nhwc, int n, int c, int h,
                   int w, int ld_out, void* stream) {
    if (ld_out < c || ld_out % 8) { set_error("unet_input: ld_out must be >= c and a multiple of 8"); return ADM_ERR_SHAPE; }
    unet_input_kernel<<<ew_grid(1LL * n * h * w, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x_nchw, sigma, sigma_is_scalar, static_cast<__nv_bfloat16*>(x_nhwc), n, c, h * w, ld_out);
    ADM_CHECK_LAUNCH("unet_input");
    return 0;
}

int adm_unet_output(const float* f1, const float* f2, int ldf, const float* x_nchw, const float* sigma,
                    int sigma_is_scalar, float* d1, float* d2, int n, int c, int h, int w, void* stream) {
    unet_output_kernel<<<ew_grid(1LL * n * c * h * w, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        f1, f2, ldf, x_nchw, sigma, sigma_is_scalar, d1, d2, n, c, h * w);
    ADM_CHECK_LAUNCH("unet_output");
    return 0;
}

int adm_unet_output_bwd(const float* dd1, const float* dd2, const float* sigma, void* df1, void* df2, int n, int c,
                        int h, int w, int ld_out, void* stream) {
    unet_output_bwd_kernel<<<ew_grid(1LL * n * h * w, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        dd1, dd2, sigma, static_cast<__nv_bfloat16*>(df1), static_cast<__nv_bfloat16*>(df2), n, c, h * w, ld_out);
    ADM_CHECK_LAUNCH("unet_output_bwd");
    return 0;
}

long long adm_augment_warp_smem(int c, int h, int w) {
    // worst-case band: rows_out output rows read 2 rows_out + 10 grid rows
    const long long rows_out = (h + AUG_BANDS - 1) / AUG_BANDS, gw = (w + 2 * AUG_PAD4) * 2;
    const long long r = 2 * rows_out + 10;
    return 4LL * (1LL * c * h * w + c * r * gw + c * r * w);
}

int adm_augment_warp(const float* x, float* y, const float* theta, const int* flips, int n, int c, int h, int w,
                     int mx0, int mx1, int my0, int my1, void* stream) {
    if (n <= 0 || n > 65535 || c <= 0 || c > 4 || h < 2 || w < 2) { set_error("augment_warp: bad shape (channels <= 4)"); return ADM_ERR_SHAPE; }
    if (mx0 < 0 || mx1 < 0 || my0 < 0 || my1 < 0 || mx0 >= w || mx1 >= w || my0 >= h || my1 >= h) {
        set_error("augment_warp: margins must lie in [0, size - 1]");
        return ADM_ERR_SHAPE;
    }
    const long long smem = adm_augment_warp_smem(c, h, w);
    if (smem > 200 * 1024) { set_error("augment_warp: image too large for one CTA (%lld B of shared memory)", smem); return ADM_ERR_SHAPE; }
    static long long attr_set = 0;
    if (smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(augment_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_error("augment_warp: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return ADM_ERR_CUDA; }
        attr_set = smem;
    }
    augment_warp_kernel<<<dim3(AUG_BANDS, n), 256, smem, static_cast<cudaStream_t>(stream)>>>(x, y, theta, flips, c, h, w, mx0,
                                                                                             mx1, my0, my1);
    ADM_CHECK_LAUNCH("augment_warp");
    return 0;
}

int adm_pack_conv_weight(const float* w, void* wpk, int cout, int c1, int c2, int ksize, const int* row_perm,
                         void* stream) {
    const int p1 = (c1 + 63) / 64 * 64;
    const int kpad = p1 + (c2 > 0 ? (c2 + 63) / 64 * 64 : 0);
    if (1LL * cout * ksize * ksize * kpad >= (1LL << 31)) { set_error("pack_conv_weight: weight too large"); return ADM_ERR_SHAPE; }
    pack_conv_weight_kernel<<<ew_grid(1LL * cout * ksize * ksize * kpad, 256, 8), 256, 0,
                              static_cast<cudaStream_t>(stream)>>>(w, static_cast<__nv_bfloat16*>(wpk), cout, c1, c2,
                                                                   ksize * ksize, p1, kpad, row_perm);
    ADM_CHECK_LAUNCH("pack_conv_weight");
    return 0;
}

int adm_unpack_conv_wgrad(const float* dw_packed, float* dw, int cout, int c1, int c2, int ksize, int accumulate,
                          const int* row_perm, void* stream) {
    const int p1 = (c1 + 63) / 64 * 64;
    const int kpad = p1 + (c2 > 0 ? (c2 + 63) / 64 * 64 : 0);
    if (1LL * cout * ksize * ksize * kpad >= (1LL << 31)) { set_error("unpack_conv_wgrad: weight too large"); return ADM_ERR_SHAPE; }
    unpack_conv_wgrad_kernel<<<ew_grid(1LL * cout * (c1 + c2) * ksize * ksize, 256, 8), 256, 0,
                               static_cast<cudaStream_t>(stream)>>>(dw_packed, dw, cout, c1, c2, ksize * ksize, p1,
                                                                    kpad, accumulate, row_perm);
    ADM_CHECK_LAUNCH("unpack_conv_wgrad");
    return 0;
}

int adm_gather_rows(const void* src, void* dst, const long long* table, long long n_rows, int row_bytes,
                    void* stream) {
    if (n_rows <= 0) return 0;
    if (row_bytes <= 0 || row_bytes % 4) { set_error("gather_rows: row_bytes must be a positive multiple of 4"); return ADM_ERR_SHAPE; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const uint8_t* sp = static_cast<const uint8_t*>(src);
    uint8_t* dp = static_cast<uint8_t*>(dst);
    const longlong2* tb = reinterpret_cast<const longlong2*>(table);
    // 16 B units need 16 B aligned offsets: the caller guarantees that whenever row_bytes is a multiple of 16
    if (row_bytes % 16 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0)
        gather_rows_kernel<uint4><<<ew_grid(n_rows * (row_bytes / 16), 256, 8), 256, 0, s>>>(sp, dp, tb, n_rows, row_bytes / 16);
    else
        gather_rows_kernel<uint32_t><<<ew_grid(n_rows * (row_bytes / 4), 256, 8), 256, 0, s>>>(sp, dp, tb, n_rows, row_bytes / 4);
    ADM_CHECK_LAUNCH("gather_rows");
    return 0;
}

int adm_transpose_weight_tiles(const void* src, void* dst, const long long* tiles, int num_tiles, void* stream) {
    if (num_tiles <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) {
        set_error("transpose_weight_tiles: buffers must be 16 B aligned");
        return ADM_ERR_SHAPE;
    }
    transpose_tiles_kernel<<<num_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), reinterpret_cast<const longlong4*>(tiles));
    ADM_CHECK_LAUNCH("transpose_weight_tiles");
    return 0;
}

int adm_cast_f32_bf16(const float* src, void* dst, long long numel, void* stream) {
    cast_f32_bf16_kernel<<<ew_grid(numel, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        src, static_cast<__nv_bfloat16*>(dst), numel);
    ADM_CHECK_LAUNCH("cast_f32_bf16");
    return 0;
}

}  // extern "C"
