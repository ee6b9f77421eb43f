// Full-resolution memory-side kernels of the RelationNet / BasicAttetnionLayer of the conditional UNet
// (reference unet/cond_unet.py:160-252).  A relation layer ends in
//     out = GroupNorm(x2 + concat_conv(cat[x1, x2])) + out_conv(bilinear_up(pooled))            (:236-251)
// The 1x1 out_conv and the bilinear resize are both linear per pixel and the resize weights sum to one, so
// out_conv(up(pooled)) == up(out_conv(pooled)): the conv runs on the few hundred pooled tokens and what is left at
// full resolution is ONE streaming pass
//     rel_gn_stats  : per-(sample, group) sum / sum of squares of pre = x2 + y, y = concat_conv output   (reads x2, y)
//     rel_gn_apply  : out = (pre - mean) * rstd * gamma + beta + bilinear(z)                   (reads x2, y; writes out)
// with the residual stream in fp32 registers (what the reference's fp32 arithmetic gives; no fp32 tensor in HBM), and
// the matching backward
//     rel_gn_bwd_sums  : per-(sample, channel) sum dout * xhat and sum dout                      (reads dout, x2, y)
//     rel_gn_bwd_apply : dpre = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dout * gamma   (same reads; writes dpre)
//     lerp_axis_bwd    : the transpose of the align_corners bilinear resize as two separable axis reductions.
// Also here: the general NHWC bilinear resize (align_corners=True, F.interpolate at :184 / :248) and the window average
// pool with zero padding at the far edges (F.pad + nn.AvgPool2d, :190-200), both directions.
// Layout: NHWC, 8 channels (16 B of bf16) per thread; C % 8 == 0, (C / groups) % 8 == 0, C / 8 divides 256.
#include "adm_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace adm {
typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void ld8(const bf16* p, float* f) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ void ld8(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x, f[1] = a.y, f[2] = a.z, f[3] = a.w, f[4] = b.x, f[5] = b.y, f[6] = b.z, f[7] = b.w;
}
__device__ __forceinline__ void st8(bf16* p, const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(f[2 * i + 1]), "f"(f[2 * i]));
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void st8(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// align_corners=True source coordinate of output index o (ATen area_pixel_compute_source_index): i0, i1, weight of i1
__device__ __forceinline__ void lerp_coord(int o, float scale, int n_in, int& i0, int& i1, float& l1) {
    const float s = scale * static_cast<float>(o);
    i0 = static_cast<int>(s);
    if (i0 > n_in - 1) i0 = n_in - 1;
    i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    l1 = s - static_cast<float>(i0);
}
static float lerp_scale(int n_in, int n_out) { return n_out > 1 ? static_cast<float>(n_in - 1) / (n_out - 1) : 0.f; }

constexpr int REL_THREADS = 256;
constexpr int REL_MAX_G = 64;
constexpr int REL_MAX_C = 1024;

// ------------------------------------------------------------------------------------------------ forward statistics
// grid (chunks, batch).  work[b][chunk][g] = (sum, sum of squares) of pre over the chunk's pixels and the group's channels.
__global__ void __launch_bounds__(REL_THREADS) rel_gn_stats_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y,
                                                                  int npix, int c, int groups, float* __restrict__ work) {
    __shared__ float sg[REL_MAX_G][2];
    const int lanes = c >> 3, ppi = REL_THREADS / lanes;
    const int cv = threadIdx.x % lanes, pl = threadIdx.x / lanes;
    const int chunk = blockIdx.x, chunks = gridDim.x, b = blockIdx.y;
    const int per = (npix + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(npix, p0 + per);
    if (threadIdx.x < groups) sg[threadIdx.x][0] = sg[threadIdx.x][1] = 0.f;
    __syncthreads();
    const size_t base = (static_cast<size_t>(b) * npix) * c + cv * 8;
    float s = 0.f, q = 0.f;
#pragma unroll 4
    for (int p = p0 + pl; p < p1; p += ppi) {
        float xv[8], yv[8];
        ld8(x + base + static_cast<size_t>(p) * c, xv);
        ld8(y + base + static_cast<size_t>(p) * c, yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float v = xv[i] + yv[i];
            s += v;
            q = fmaf(v, v, q);
        }
    }
    const int g = (cv * 8) / (c / groups);
    atomicAdd(&sg[g][0], s);
    atomicAdd(&sg[g][1], q);
    __syncthreads();
    if (threadIdx.x < groups) {
        float* o = work + ((static_cast<size_t>(b) * chunks + chunk) * groups + threadIdx.x) * 2;
        o[0] = sg[threadIdx.x][0];
        o[1] = sg[threadIdx.x][1];
    }
}

// ------------------------------------------------------------------------------------------------ forward apply
// grid (chunks, batch), the same chunk count as the statistics pass.  stats[b][g] = (mean, rstd) is written for backward.
template <typename TOut>
__global__ void __launch_bounds__(REL_THREADS)
rel_gn_apply_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const float* __restrict__ z, int h, int w,
                    int c, int hq, int wq, float sh, float sw, int groups, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, const float* __restrict__ work, float* __restrict__ stats,
                    TOut* __restrict__ out) {
    __shared__ float sm[REL_MAX_G][2];
    const int npix = h * w;
    const int lanes = c >> 3, ppi = REL_THREADS / lanes;
    const int cv = threadIdx.x % lanes, pl = threadIdx.x / lanes;
    const int chunk = blockIdx.x, chunks = gridDim.x, b = blockIdx.y;
    // chunk partials -> (mean, rstd): one warp per group, lanes over the chunks (every CTA of the sample repeats this, so it
    // must not be a serial walk over the chunks)
    for (int gg = threadIdx.x >> 5; gg < groups; gg += REL_THREADS / 32) {
        const int lane = threadIdx.x & 31;
        double s = 0.0, q = 0.0;
        const float* wk = work + (static_cast<size_t>(b) * chunks * groups + gg) * 2;
        for (int k = lane; k < chunks; k += 32) {
            s += wk[static_cast<size_t>(k) * groups * 2];
            q += wk[static_cast<size_t>(k) * groups * 2 + 1];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) {
            const double n = static_cast<double>(npix) * (c / groups);
            const double mean = s / n;
            double var = q / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const float rstd = rsqrtf(static_cast<float>(var) + eps);
            sm[gg][0] = static_cast<float>(mean);
            sm[gg][1] = rstd;
            if (chunk == 0) {
                stats[(static_cast<size_t>(b) * groups + gg) * 2] = static_cast<float>(mean);
                stats[(static_cast<size_t>(b) * groups + gg) * 2 + 1] = rstd;
            }
        }
    }
    __syncthreads();
    const int g = (cv * 8) / (c / groups);
    const float mean = sm[g][0], rstd = sm[g][1];
    float ga[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {  // out = pre * ga + be + resize(z)
        ga[i] = gamma[cv * 8 + i] * rstd;
        be[i] = fmaf(-mean, ga[i], beta[cv * 8 + i]);
    }
    const int per = (npix + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(npix, p0 + per);
    const size_t base = (static_cast<size_t>(b) * npix) * c + cv * 8;
    const float* zb = z + (static_cast<size_t>(b) * hq * wq) * c + cv * 8;
#pragma unroll 2
    for (int p = p0 + pl; p < p1; p += ppi) {
        float xv[8], yv[8], o[8];
        ld8(x + base + static_cast<size_t>(p) * c, xv);
        ld8(y + base + static_cast<size_t>(p) * c, yv);
        const int ph = p / w, pw = p - ph * w;  // (an incrementally advanced (row, column) pair measured 25 % slower: the
                                                //  loop-carried update keeps the compiler from batching the loads)
        int h0, h1, w0, w1;
        float lh, lw;
        lerp_coord(ph, sh, hq, h0, h1, lh);
        lerp_coord(pw, sw, wq, w0, w1, lw);
        float z00[8], z01[8], z10[8], z11[8];
        ld8(zb + (static_cast<size_t>(h0) * wq + w0) * c, z00);
        ld8(zb + (static_cast<size_t>(h0) * wq + w1) * c, z01);
        ld8(zb + (static_cast<size_t>(h1) * wq + w0) * c, z10);
        ld8(zb + (static_cast<size_t>(h1) * wq + w1) * c, z11);
        const float a00 = (1.f - lh) * (1.f - lw), a01 = (1.f - lh) * lw, a10 = lh * (1.f - lw), a11 = lh * lw;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float up = a00 * z00[i] + a01 * z01[i] + a10 * z10[i] + a11 * z11[i];
            o[i] = fmaf(xv[i] + yv[i], ga[i], be[i]) + up;
        }
        st8(out + base + static_cast<size_t>(p) * c, o);
    }
}

// ------------------------------------------------------------------------------------------------ backward sums
// grid (chunks, batch).  work[b][chunk][ch] = (sum dout * xhat, sum dout) over the chunk's pixels.
__global__ void __launch_bounds__(REL_THREADS)
rel_gn_bwd_sums_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ x, const bf16* __restrict__ y, int npix,
                       int c, int groups, const float* __restrict__ stats, float* __restrict__ work) {
    extern __shared__ float sc[];  // [c][2]
    const int lanes = c >> 3, ppi = REL_THREADS / lanes;
    const int cv = threadIdx.x % lanes, pl = threadIdx.x / lanes;
    const int chunk = blockIdx.x, chunks = gridDim.x, b = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * c; i += REL_THREADS) sc[i] = 0.f;
    __syncthreads();
    const int g = (cv * 8) / (c / groups);
    const float mean = stats[(static_cast<size_t>(b) * groups + g) * 2];
    const float rstd = stats[(static_cast<size_t>(b) * groups + g) * 2 + 1];
    const int per = (npix + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(npix, p0 + per);
    const size_t base = (static_cast<size_t>(b) * npix) * c + cv * 8;
    float A[8], B[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) A[i] = B[i] = 0.f;
#pragma unroll 2
    for (int p = p0 + pl; p < p1; p += ppi) {
        float xv[8], yv[8], dv[8];
        ld8(x + base + static_cast<size_t>(p) * c, xv);
        ld8(y + base + static_cast<size_t>(p) * c, yv);
        ld8(dout + base + static_cast<size_t>(p) * c, dv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xh = (xv[i] + yv[i] - mean) * rstd;
            A[i] = fmaf(dv[i], xh, A[i]);
            B[i] += dv[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        atomicAdd(&sc[(cv * 8 + i) * 2], A[i]);
        atomicAdd(&sc[(cv * 8 + i) * 2 + 1], B[i]);
    }
    __syncthreads();
    float* o = work + (static_cast<size_t>(b) * chunks + chunk) * c * 2;
    for (int i = threadIdx.x; i < 2 * c; i += REL_THREADS) o[i] = sc[i];
}

// ------------------------------------------------------------------------------------------------ backward apply
__global__ void __launch_bounds__(REL_THREADS)
rel_gn_bwd_apply_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ x, const bf16* __restrict__ y, int npix,
                        int c, int groups, const float* __restrict__ gamma, const float* __restrict__ stats,
                        const float* __restrict__ work, bf16* __restrict__ dpre) {
    extern __shared__ float sc[];  // [c][2] channel sums of this sample, then [groups][2] group sums
    float* sgp = sc + 2 * c;
    const int lanes = c >> 3, ppi = REL_THREADS / lanes;
    const int cv = threadIdx.x % lanes, pl = threadIdx.x / lanes;
    const int chunk = blockIdx.x, chunks = gridDim.x, b = blockIdx.y;
    const int cpg = c / groups;
    for (int i = threadIdx.x; i < 2 * c; i += REL_THREADS) {
        const float* wk = work + static_cast<size_t>(b) * chunks * c * 2 + i;
        float s = 0.f;
#pragma unroll 8
        for (int k = 0; k < chunks; ++k) s += wk[static_cast<size_t>(k) * c * 2];  // eight loads in flight
        sc[i] = s;
    }
    __syncthreads();
    if (threadIdx.x < groups) {
        float s1 = 0.f, s2 = 0.f;
        for (int j = 0; j < cpg; ++j) {
            const int ch = threadIdx.x * cpg + j;
            s1 = fmaf(gamma[ch], sc[ch * 2], s1);
            s2 = fmaf(gamma[ch], sc[ch * 2 + 1], s2);
        }
        const float inv_n = 1.f / (static_cast<float>(npix) * cpg);
        sgp[threadIdx.x * 2] = s1 * inv_n;
        sgp[threadIdx.x * 2 + 1] = s2 * inv_n;
    }
    __syncthreads();
    const int g = (cv * 8) / cpg;
    const float mean = stats[(static_cast<size_t>(b) * groups + g) * 2];
    const float rstd = stats[(static_cast<size_t>(b) * groups + g) * 2 + 1];
    const float m1 = sgp[g * 2], m2 = sgp[g * 2 + 1];
    float ga[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ga[i] = gamma[cv * 8 + i];
    const int per = (npix + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(npix, p0 + per);
    const size_t base = (static_cast<size_t>(b) * npix) * c + cv * 8;
#pragma unroll 2
    for (int p = p0 + pl; p < p1; p += ppi) {
        float xv[8], yv[8], dv[8], o[8];
        ld8(x + base + static_cast<size_t>(p) * c, xv);
        ld8(y + base + static_cast<size_t>(p) * c, yv);
        ld8(dout + base + static_cast<size_t>(p) * c, dv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float xh = (xv[i] + yv[i] - mean) * rstd;
            o[i] = rstd * (dv[i] * ga[i] - m2 - xh * m1);
        }
        st8(dpre + base + static_cast<size_t>(p) * c, o);
    }
}

// ------------------------------------------------------------------------------------------------ resize, forward
// y[b][oh][ow][:] = bilinear(x[b]) (align_corners=True).  One thread per 8 output channels.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) bilinear_fwd_kernel(const TIn* __restrict__ x, long long ldx, int hin, int win,
                                                           int c, TOut* __restrict__ y, long long ldy, int hout,
                                                           int wout, float sh, float sw, unsigned total) {
    const unsigned lanes = c >> 3;  // 32-bit index math: a 64-bit division costs ~100 instructions on the SM
    for (unsigned t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
        const int cv = static_cast<int>(t % lanes);
        unsigned p = t / lanes;
        const int ow = static_cast<int>(p % wout);
        p /= wout;
        const int oh = static_cast<int>(p % hout);
        const long long b = p / hout;
        int h0, h1, w0, w1;
        float lh, lw;
        lerp_coord(oh, sh, hin, h0, h1, lh);
        lerp_coord(ow, sw, win, w0, w1, lw);
        const TIn* xb = x + (b * hin * win) * ldx + cv * 8;
        float v00[8], v01[8], v10[8], v11[8], o[8];
        ld8(xb + (static_cast<long long>(h0) * win + w0) * ldx, v00);
        ld8(xb + (static_cast<long long>(h0) * win + w1) * ldx, v01);
        ld8(xb + (static_cast<long long>(h1) * win + w0) * ldx, v10);
        ld8(xb + (static_cast<long long>(h1) * win + w1) * ldx, v11);
        const float a00 = (1.f - lh) * (1.f - lw), a01 = (1.f - lh) * lw, a10 = lh * (1.f - lw), a11 = lh * lw;
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = a00 * v00[i] + a01 * v01[i] + a10 * v10[i] + a11 * v11[i];
        st8(y + ((b * hout + oh) * wout + ow) * ldy + cv * 8, o);
    }
}

// ------------------------------------------------------------------------------------------------ resize, backward
// One axis of the transposed resize: src [outer][n_out][inner] -> dst [outer][n_in][inner],
// dst[o][i][k] = sum_p weight(p -> i) * src[o][p][k]; the weights are the forward's, re-derived per p.  inner % 8 == 0.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) lerp_axis_bwd_kernel(const TIn* __restrict__ src, TOut* __restrict__ dst,
                                                            int n_out, int n_in, long long inner, float scale,
                                                            unsigned total) {
    const unsigned iv = static_cast<unsigned>(inner >> 3);
    for (unsigned t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
        const long long kv = t % iv;
        const unsigned q = t / iv;
        const int i = static_cast<int>(q % n_in);
        const long long o = q / n_in;
        int plo = 0, phi = n_out - 1;
        if (scale > 0.f) {
            plo = max(0, static_cast<int>(floorf((i - 1) / scale)) - 1);
            phi = min(n_out - 1, static_cast<int>(ceilf((i + 1) / scale)) + 1);
        }
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const TIn* sp = src + (o * n_out) * inner + kv * 8;
        for (int p = plo; p <= phi; ++p) {
            int i0, i1;
            float l1;
            lerp_coord(p, scale, n_in, i0, i1, l1);
            const float wgt = (i0 == i ? 1.f - l1 : 0.f) + (i1 == i ? l1 : 0.f);
            if (wgt == 0.f) continue;
            float v[8];
            ld8(sp + static_cast<long long>(p) * inner, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(wgt, v[j], acc[j]);
        }
        st8(dst + (o * n_in + i) * inner + kv * 8, acc);
    }
}

// ------------------------------------------------------------------------------------------------ window average pool
// out[b][i][j][:] = sum over the kh x kw window (pixels past the edge count as zero) / (kh * kw)
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const bf16* __restrict__ x, int h, int w, int c, int kh, int kw,
                                                          int ho, int wo, bf16* __restrict__ out, unsigned total) {
    const unsigned lanes = c >> 3;
    const float inv = 1.f / (kh * kw);
    for (unsigned t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
        const int cv = static_cast<int>(t % lanes);
        unsigned p = t / lanes;
        const int j = static_cast<int>(p % wo);
        p /= wo;
        const int i = static_cast<int>(p % ho);
        const long long b = p / ho;
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        const int r1 = min(h, (i + 1) * kh), c1 = min(w, (j + 1) * kw);
        for (int r = i * kh; r < r1; ++r) {
            const bf16* row = x + ((b * h + r) * w) * c + cv * 8;
#pragma unroll 4
            for (int cc = j * kw; cc < c1; ++cc) {
                float v[8];
                ld8(row + static_cast<long long>(cc) * c, v);
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] += v[q];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] *= inv;
        st8(out + ((b * ho + i) * wo + j) * c + cv * 8, acc);
    }
}

// dx[b][r][cc][:] = dy[b][r / kh][cc / kw][:] / (kh * kw)
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const bf16* __restrict__ dy, int h, int w, int c, int kh, int kw,
                                                          int ho, int wo, bf16* __restrict__ dx, unsigned total) {
    const unsigned lanes = c >> 3;
    const float inv = 1.f / (kh * kw);
    for (unsigned t = blockIdx.x * 256u + threadIdx.x; t < total; t += gridDim.x * 256u) {
        const int cv = static_cast<int>(t % lanes);
        unsigned p = t / lanes;
        const int cc = static_cast<int>(p % w);
        p /= w;
        const int r = static_cast<int>(p % h);
        const long long b = p / h;
        float v[8];
        ld8(dy + ((b * ho + r / kh) * wo + cc / kw) * c + cv * 8, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] *= inv;
        st8(dx + ((b * h + r) * w + cc) * c + cv * 8, v);
    }
}

static int rel_shape_ok(int batch, int h, int w, int c, int groups) {
    if (batch <= 0 || h <= 0 || w <= 0 || c <= 0 || groups <= 0) return 0;
    if (c % 8 || c > REL_MAX_C || groups > REL_MAX_G || c % groups || (c / groups) % 8) return 0;
    if (REL_THREADS % (c / 8)) return 0;
    if (1LL * h * w > (1LL << 30)) return 0;
    return 1;
}

static int rel_chunks(int batch, long long npix, int c) {
    const long long ppi = REL_THREADS / (c / 8);
    long long want = (4LL * num_sms() + batch - 1) / batch;  // (8 CTAs per SM measured slower: the chunk combine grows)
    const long long most = (npix + 4 * ppi - 1) / (4 * ppi);  // at least four iterations per thread
    if (want > most) want = most;
    if (want > 256) want = 256;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

static bool flat_ok(long long total, const char* what) {  // the flat kernels index with 32 bits
    if (total > 0 && total < (1LL << 31)) return true;
    set_error("%s: %lld vectors of 8 channels do not fit the 32-bit index space", what, total);
    return false;
}

static unsigned flat_grid(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = 32LL * num_sms();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<unsigned>(blocks);
}

}  // namespace adm

using namespace adm;

extern "C" {

int adm_rel_gn_ok(int batch, int h, int w, int c, int groups) { return rel_shape_ok(batch, h, w, c, groups); }

int adm_rel_gn_chunks(int batch, int h, int w, int c) {
    if (c <= 0 || c % 8 || REL_THREADS % (c / 8)) return 0;
    return rel_chunks(batch, 1LL * h * w, c);
}

int adm_rel_gn_fwd(const void* x, const void* y, const float* z, int batch, int h, int w, int c, int hq, int wq,
                   int groups, const float* gamma, const float* beta, float eps, void* out, int out_fp32, float* stats,
                   float* work, void* stream) {
    if (!rel_shape_ok(batch, h, w, c, groups) || hq <= 0 || wq <= 0) {
        set_error("rel_gn_fwd: unsupported shape [%d, %d, %d, %d] groups %d pooled %d x %d", batch, h, w, c, groups, hq, wq);
        return ADM_ERR_SHAPE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = rel_chunks(batch, 1LL * h * w, c);
    const bf16 *xp = static_cast<const bf16*>(x), *yp = static_cast<const bf16*>(y);
    rel_gn_stats_kernel<<<dim3(chunks, batch), REL_THREADS, 0, s>>>(xp, yp, h * w, c, groups, work);
    ADM_CHECK_LAUNCH("rel_gn_stats");
    const float sh = lerp_scale(hq, h), sw = lerp_scale(wq, w);
    if (out_fp32)
        rel_gn_apply_kernel<float><<<dim3(chunks, batch), REL_THREADS, 0, s>>>(
            xp, yp, z, h, w, c, hq, wq, sh, sw, groups, gamma, beta, eps, work, stats, static_cast<float*>(out));
    else
        rel_gn_apply_kernel<bf16><<<dim3(chunks, batch), REL_THREADS, 0, s>>>(
            xp, yp, z, h, w, c, hq, wq, sh, sw, groups, gamma, beta, eps, work, stats, static_cast<bf16*>(out));
    ADM_CHECK_LAUNCH("rel_gn_apply");
    return 0;
}

int adm_rel_gn_bwd(const void* dout, const void* x, const void* y, int batch, int h, int w, int c, int groups,
                   const float* gamma, const float* stats, void* dpre, float* work, void* stream) {
    if (!rel_shape_ok(batch, h, w, c, groups)) {
        set_error("rel_gn_bwd: unsupported shape [%d, %d, %d, %d] groups %d", batch, h, w, c, groups);
        return ADM_ERR_SHAPE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = rel_chunks(batch, 1LL * h * w, c);
    const bf16 *dp = static_cast<const bf16*>(dout), *xp = static_cast<const bf16*>(x), *yp = static_cast<const bf16*>(y);
    rel_gn_bwd_sums_kernel<<<dim3(chunks, batch), REL_THREADS, 2 * c * sizeof(float), s>>>(dp, xp, yp, h * w, c, groups,
                                                                                         stats, work);
    ADM_CHECK_LAUNCH("rel_gn_bwd_sums");
    rel_gn_bwd_apply_kernel<<<dim3(chunks, batch), REL_THREADS, (2 * c + 2 * groups) * sizeof(float), s>>>(
        dp, xp, yp, h * w, c, groups, gamma, stats, work, static_cast<bf16*>(dpre));
    ADM_CHECK_LAUNCH("rel_gn_bwd_apply");
    return 0;
}

int adm_bilinear_fwd(const void* x, long long ldx, int batch, int hin, int win, int c, void* y, long long ldy, int hout,
                     int wout, int in_fp32, int out_fp32, void* stream) {
    if (batch <= 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0 || c <= 0 || c % 8 || ldx % 8 || ldy % 8) {
        set_error("bilinear_fwd: bad shape (C and the pixel strides must be multiples of 8)");
        return ADM_ERR_SHAPE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long total = 1LL * batch * hout * wout * (c / 8);
    if (!flat_ok(total, "bilinear_fwd")) return ADM_ERR_SHAPE;
    const float sh = lerp_scale(hin, hout), sw = lerp_scale(win, wout);
    const unsigned grid = flat_grid(total);
    if (in_fp32 && out_fp32)
        bilinear_fwd_kernel<float, float><<<grid, 256, 0, s>>>(static_cast<const float*>(x), ldx, hin, win, c,
                                                               static_cast<float*>(y), ldy, hout, wout, sh, sw, static_cast<unsigned>(total));
    else if (in_fp32)
        bilinear_fwd_kernel<float, bf16><<<grid, 256, 0, s>>>(static_cast<const float*>(x), ldx, hin, win, c,
                                                              static_cast<bf16*>(y), ldy, hout, wout, sh, sw, static_cast<unsigned>(total));
    else if (out_fp32)
        bilinear_fwd_kernel<bf16, float><<<grid, 256, 0, s>>>(static_cast<const bf16*>(x), ldx, hin, win, c,
                                                              static_cast<float*>(y), ldy, hout, wout, sh, sw, static_cast<unsigned>(total));
    else
        bilinear_fwd_kernel<bf16, bf16><<<grid, 256, 0, s>>>(static_cast<const bf16*>(x), ldx, hin, win, c,
                                                             static_cast<bf16*>(y), ldy, hout, wout, sh, sw, static_cast<unsigned>(total));
    ADM_CHECK_LAUNCH("bilinear_fwd");
    return 0;
}

int adm_bilinear_bwd(const void* dy, int batch, int hout, int wout, int c, int hin, int win, float* tmp, float* dx,
                     void* stream) {
    if (batch <= 0 || hin <= 0 || win <= 0 || hout <= 0 || wout <= 0 || c <= 0 || c % 8) {
        set_error("bilinear_bwd: bad shape (C must be a multiple of 8)");
        return ADM_ERR_SHAPE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // rows first: [batch][hout][wout * c] -> tmp [batch][hin][wout * c] (fp32), then columns: [batch * hin][wout][c] -> dx
    const long long inner1 = 1LL * wout * c;
    const long long total1 = 1LL * batch * hin * (inner1 / 8);
    if (!flat_ok(total1, "bilinear_bwd") || !flat_ok(1LL * batch * hout * (inner1 / 8), "bilinear_bwd")) return ADM_ERR_SHAPE;
    lerp_axis_bwd_kernel<bf16, float><<<flat_grid(total1), 256, 0, s>>>(static_cast<const bf16*>(dy), tmp, hout, hin,
                                                                        inner1, lerp_scale(hin, hout), static_cast<unsigned>(total1));
    ADM_CHECK_LAUNCH("lerp_axis_bwd(rows)");
    const long long total2 = 1LL * batch * hin * win * (c / 8);
    lerp_axis_bwd_kernel<float, float><<<flat_grid(total2), 256, 0, s>>>(tmp, dx, wout, win, c, lerp_scale(win, wout),
                                                                         static_cast<unsigned>(total2));
    ADM_CHECK_LAUNCH("lerp_axis_bwd(cols)");
    return 0;
}

int adm_avgpool_fwd(const void* x, int batch, int h, int w, int c, int kh, int kw, void* out, void* stream) {
    if (batch <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 || kh <= 0 || kw <= 0) {
        set_error("avgpool_fwd: bad shape (C must be a multiple of 8)");
        return ADM_ERR_SHAPE;
    }
    const int ho = (h + kh - 1) / kh, wo = (w + kw - 1) / kw;
    const long long total = 1LL * batch * ho * wo * (c / 8);
    if (!flat_ok(total, "avgpool_fwd")) return ADM_ERR_SHAPE;
    avgpool_fwd_kernel<<<flat_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(x), h, w, c, kh, kw, ho, wo, static_cast<bf16*>(out), static_cast<unsigned>(total));
    ADM_CHECK_LAUNCH("avgpool_fwd");
    return 0;
}

int adm_avgpool_bwd(const void* dy, int batch, int h, int w, int c, int kh, int kw, void* dx, void* stream) {
    if (batch <= 0 || h <= 0 || w <= 0 || c <= 0 || c % 8 || kh <= 0 || kw <= 0) {
        set_error("avgpool_bwd: bad shape (C must be a multiple of 8)");
        return ADM_ERR_SHAPE;
    }
    const int ho = (h + kh - 1) / kh, wo = (w + kw - 1) / kw;
    const long long total = 1LL * batch * h * w * (c / 8);
    if (!flat_ok(total, "avgpool_bwd")) return ADM_ERR_SHAPE;
    avgpool_bwd_kernel<<<flat_grid(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(dy), h, w, c, kh, kw, ho, wo, static_cast<bf16*>(dx), static_cast<unsigned>(total));
    ADM_CHECK_LAUNCH("avgpool_bwd");
    return 0;
}

}  // extern "C"
