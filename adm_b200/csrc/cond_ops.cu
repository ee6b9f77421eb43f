// Kernels specific to the conditional two-decoder UNet (reference unet/cond_unet.py):
//   K11 ws_pack / ws_pack_bwd : weight standardisation of WeightStandardizedConv2d (:345-358) fused with the bf16
//                               [Cout][tap][Cin] re-pack the GEMM engine consumes, and its backward fused with the
//                               un-pack of the packed weight gradient.
//   K12 linattn_*             : LinearAttention (:503-531) without materialising the two softmaxes:
//         pass A  (reads k, v) per-(b, head) online softmax over pixels + context accumulation  -> partials
//         combine              partials -> ctx[d][e] = sum_n softmax_n(k)[d,n] v[e,n] / N, and (max, Z) per k channel
//         pass C  (reads q)    out[e,n] = sum_d ctx[d][e] * softmax_d(q)[d,n] * scale
//       backward: pass B1 (reads q, dout) -> dq and dctx partials; combine; pass B2 (reads k, v) -> dk, dv.
// Channel order of qkv is (which, head, d) exactly as `to_qkv(x).chunk(3, dim=1)` + 'b (h c) x y' produces, head dim 32.
#include "adm_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace adm {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {  // red: >= 32 floats of smem
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    if (w == 0) {
        t = warp_sum(t);
        if (l == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}

// ------------------------------------------------------------------------------------------------ K11 forward
// One block per output channel.  w fp32 [cout][cin][k][k]; wpk bf16 [cout][taps][kpad] (channels >= cin zero);
// stats fp32 [cout][2] = (mean, rstd) with var = mean((w - mean)^2) (unbiased=False) and rstd = rsqrt(var + eps).
__global__ void __launch_bounds__(256) ws_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpk,
                                                      float* __restrict__ stats, int cin, int taps, int kpad,
                                                      float eps) {
    __shared__ float red[32];
    const int o = blockIdx.x;
    const int K = cin * taps;
    const float* wr = w + 1LL * o * K;
    float s = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) s += wr[i];
    const float mean = block_sum(s, red) / K;
    float q = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const float d = wr[i] - mean;
        q += d * d;
    }
    const float var = block_sum(q, red) / K;
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
        stats[2 * o] = mean;
        stats[2 * o + 1] = rstd;
    }
    __nv_bfloat16* out = wpk + 1LL * o * taps * kpad;
    for (int i = threadIdx.x; i < taps * kpad; i += blockDim.x) {
        const int t = i / kpad, c = i - t * kpad;
        out[i] = __float2bfloat16(c < cin ? (wr[c * taps + t] - mean) * rstd : 0.f);
    }
}

// Backward: g = d(loss)/d(w_hat) packed fp32 [cout][taps][kpad];  dw[o][c][t] (+)= rstd * (g - mean(g) - w_hat * mean(g * w_hat)).
__global__ void __launch_bounds__(256) ws_pack_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                          const float* __restrict__ stats, float* __restrict__ dw,
                                                          int cin, int taps, int kpad, int accumulate) {
    __shared__ float red[32];
    const int o = blockIdx.x;
    const int K = cin * taps;
    const float* wr = w + 1LL * o * K;
    const float* gr = g + 1LL * o * taps * kpad;
    const float mean = stats[2 * o], rstd = stats[2 * o + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const int c = i / taps, t = i - c * taps;
        const float gi = gr[t * kpad + c];
        s1 += gi;
        s2 += gi * (wr[i] - mean) * rstd;
    }
    const float m1 = block_sum(s1, red) / K;
    const float m2 = block_sum(s2, red) / K;
    float* dr = dw + 1LL * o * K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        const int c = i / taps, t = i - c * taps;
        const float wh = (wr[i] - mean) * rstd;
        const float v = rstd * (gr[t * kpad + c] - m1 - wh * m2);
        dr[i] = accumulate ? dr[i] + v : v;
    }
}

// ------------------------------------------------------------------------------------------------ K12 LinearAttention
constexpr int LA_D = 32;          // head dim (reference default dim_head = 32)
constexpr int LA_WARPS = 8;       // warps per block in the streaming passes
constexpr int LA_U = 4;           // pixels per warp iteration in the per-pixel passes

__device__ __forceinline__ float ldbf(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// Pass A.  grid (chunks, B*heads), 256 threads.  Lane = k channel d; each warp strides over the chunk's pixels keeping an
// online softmax (m, Z) for its channel and the un-normalised context row acc[e] = sum_n exp(k[d,n] - m) v[e,n].
// part: [B*heads][chunks][32][34] = (acc[0..31], m, Z).
__global__ void __launch_bounds__(LA_WARPS * 32) linattn_ctx_partial_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                            long long ld, int N, int heads,
                                                                            float* __restrict__ part) {
    __shared__ float vsm[LA_WARPS][LA_D];
    __shared__ float comb[LA_WARPS][LA_D][LA_D + 2];
    const int bh = blockIdx.y, b = bh / heads, h = bh % heads;
    const int hidden = heads * LA_D;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per = (N + gridDim.x - 1) / gridDim.x;
    const int n0 = blockIdx.x * per, n1 = min(N, n0 + per);
    const __nv_bfloat16* kb = qkv + 1LL * b * N * ld + hidden + h * LA_D + lane;
    const __nv_bfloat16* vb = qkv + 1LL * b * N * ld + 2 * hidden + h * LA_D + lane;
    float m = -INFINITY, Z = 0.f, acc[LA_D];
#pragma unroll
    for (int e = 0; e < LA_D; ++e) acc[e] = 0.f;
    for (int n = n0 + warp; n < n1; n += LA_WARPS) {
        const float kd = ldbf(kb + 1LL * n * ld);
        vsm[warp][lane] = ldbf(vb + 1LL * n * ld);
        __syncwarp();
        if (kd > m) {  // rescale the running sums (rare after the first few pixels)
            const float f = __expf(m - kd);
            Z *= f;
#pragma unroll
            for (int e = 0; e < LA_D; ++e) acc[e] *= f;
            m = kd;
        }
        const float p = __expf(kd - m);
        Z += p;
#pragma unroll
        for (int e = 0; e < LA_D; e += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(&vsm[warp][e]);
            acc[e] = fmaf(p, v4.x, acc[e]);
            acc[e + 1] = fmaf(p, v4.y, acc[e + 1]);
            acc[e + 2] = fmaf(p, v4.z, acc[e + 2]);
            acc[e + 3] = fmaf(p, v4.w, acc[e + 3]);
        }
        __syncwarp();
    }
#pragma unroll
    for (int e = 0; e < LA_D; ++e) comb[warp][lane][e] = acc[e];
    comb[warp][lane][LA_D] = m;
    comb[warp][lane][LA_D + 1] = Z;
    __syncthreads();
    if (warp == 0) {
        float M = -INFINITY;
        for (int w = 0; w < LA_WARPS; ++w) M = fmaxf(M, comb[w][lane][LA_D]);
        float Zt = 0.f, out[LA_D];
#pragma unroll
        for (int e = 0; e < LA_D; ++e) out[e] = 0.f;
        for (int w = 0; w < LA_WARPS; ++w) {
            const float mw = comb[w][lane][LA_D];
            const float f = mw == -INFINITY ? 0.f : __expf(mw - M);
            Zt += f * comb[w][lane][LA_D + 1];
#pragma unroll
            for (int e = 0; e < LA_D; ++e) out[e] = fmaf(f, comb[w][lane][e], out[e]);
        }
        float* dst = part + ((1LL * bh * gridDim.x + blockIdx.x) * LA_D + lane) * (LA_D + 2);
#pragma unroll
        for (int e = 0; e < LA_D; ++e) dst[e] = out[e];
        dst[LA_D] = M;
        dst[LA_D + 1] = Zt;
    }
}

// Combine.  grid (B*heads), 32 threads (lane = d).  ctx [B*heads][32][32] (d-major), kstat [B*heads][32][2] = (max, Z).
__global__ void linattn_ctx_combine_kernel(const float* __restrict__ part, int chunks, int N, float* __restrict__ ctx,
                                           float* __restrict__ kstat) {
    const int bh = blockIdx.x, lane = threadIdx.x;
    float M = -INFINITY;
    for (int c = 0; c < chunks; ++c) M = fmaxf(M, part[((1LL * bh * chunks + c) * LA_D + lane) * (LA_D + 2) + LA_D]);
    float Z = 0.f, out[LA_D];
#pragma unroll
    for (int e = 0; e < LA_D; ++e) out[e] = 0.f;
    for (int c = 0; c < chunks; ++c) {
        const float* src = part + ((1LL * bh * chunks + c) * LA_D + lane) * (LA_D + 2);
        const float mw = src[LA_D];
        const float f = mw == -INFINITY ? 0.f : __expf(mw - M);
        Z += f * src[LA_D + 1];
#pragma unroll
        for (int e = 0; e < LA_D; ++e) out[e] = fmaf(f, src[e], out[e]);
    }
    const float inv = 1.f / (Z * static_cast<float>(N));
#pragma unroll
    for (int e = 0; e < LA_D; ++e) ctx[(1LL * bh * LA_D + lane) * LA_D + e] = out[e] * inv;
    kstat[(1LL * bh * LA_D + lane) * 2] = M;
    kstat[(1LL * bh * LA_D + lane) * 2 + 1] = Z;
}

// Pass C.  grid (pixel blocks, B), blockDim = heads*32: warp = head, lane = e (and = d while computing softmax_d(q)).
__global__ void linattn_apply_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int N, int heads,
                                     const float* __restrict__ ctx, float scale, __nv_bfloat16* __restrict__ out,
                                     long long ldo, int pix_per_block) {
    extern __shared__ float qsm[];  // [heads][32]
    const int b = blockIdx.y, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float col[LA_D];  // ctx[d][e = lane]
    const float* cb = ctx + (1LL * (b * heads + h) * LA_D) * LA_D;
#pragma unroll
    for (int d = 0; d < LA_D; ++d) col[d] = cb[d * LA_D + lane];
    const int n0 = blockIdx.x * pix_per_block, n1 = min(N, n0 + pix_per_block);
    const __nv_bfloat16* qb = qkv + 1LL * b * N * ld + h * LA_D + lane;
    __nv_bfloat16* ob = out + 1LL * b * N * ldo + h * LA_D + lane;
    float* qs = qsm + h * LA_U * LA_D;  // [LA_U][32] per warp
    for (int n = n0; n < n1; n += LA_U) {  // LA_U pixels per iteration: independent shuffle / FMA chains in flight
        float q[LA_U], mx[LA_U], pe[LA_U], sm_[LA_U];
#pragma unroll
        for (int u = 0; u < LA_U; ++u) q[u] = n + u < n1 ? ldbf(qb + 1LL * (n + u) * ld) : 0.f;
#pragma unroll
        for (int u = 0; u < LA_U; ++u) mx[u] = warp_max(q[u]);
#pragma unroll
        for (int u = 0; u < LA_U; ++u) pe[u] = __expf(q[u] - mx[u]);
#pragma unroll
        for (int u = 0; u < LA_U; ++u) sm_[u] = warp_sum(pe[u]);
#pragma unroll
        for (int u = 0; u < LA_U; ++u) qs[u * LA_D + lane] = pe[u] * (scale / sm_[u]);
        __syncwarp();
        float o[LA_U];
#pragma unroll
        for (int u = 0; u < LA_U; ++u) o[u] = 0.f;
#pragma unroll
        for (int d = 0; d < LA_D; d += 4) {
#pragma unroll
            for (int u = 0; u < LA_U; ++u) {
                const float4 q4 = *reinterpret_cast<const float4*>(&qs[u * LA_D + d]);
                o[u] = fmaf(col[d], q4.x, o[u]);
                o[u] = fmaf(col[d + 1], q4.y, o[u]);
                o[u] = fmaf(col[d + 2], q4.z, o[u]);
                o[u] = fmaf(col[d + 3], q4.w, o[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < LA_U; ++u)
            if (n + u < n1) ob[1LL * (n + u) * ldo] = __float2bfloat16(o[u]);
        __syncwarp();
    }
}

// Backward pass B1.  grid (chunks, B), blockDim = heads*32: warp = head, lane = d.  Reads q and dout; writes dq and the
// partial dctx[d][e] = sum_n softmax_d(q)[d,n]*scale * dout[e,n] of this chunk: dpart [B*heads][chunks][32][32].
__global__ void linattn_bwd_q_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int N, int heads,
                                     const float* __restrict__ ctx, float scale, const __nv_bfloat16* __restrict__ dout,
                                     long long ldd, __nv_bfloat16* __restrict__ dqkv, long long ldg,
                                     float* __restrict__ dpart) {
    extern __shared__ float dsm[];  // [heads][32] staged dout
    const int b = blockIdx.y, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float row[LA_D], acc[LA_D];  // ctx[d = lane][e], dctx partial row
    const float* cb = ctx + (1LL * (b * heads + h) * LA_D + lane) * LA_D;
#pragma unroll
    for (int e = 0; e < LA_D; ++e) { row[e] = cb[e]; acc[e] = 0.f; }
    const int per = (N + gridDim.x - 1) / gridDim.x;
    const int n0 = blockIdx.x * per, n1 = min(N, n0 + per);
    const __nv_bfloat16* qb = qkv + 1LL * b * N * ld + h * LA_D + lane;
    const __nv_bfloat16* db = dout + 1LL * b * N * ldd + h * LA_D + lane;
    __nv_bfloat16* gq = dqkv + 1LL * b * N * ldg + h * LA_D + lane;
    float* ds = dsm + h * LA_U * LA_D;  // [LA_U][32] staged dout rows per warp
    for (int n = n0; n < n1; n += LA_U) {
        float q[LA_U], mx[LA_U], pe[LA_U], sv[LA_U];
#pragma unroll
        for (int u = 0; u < LA_U; ++u) {
            const bool ok = n + u < n1;
            q[u] = ok ? ldbf(qb + 1LL * (n + u) * ld) : 0.f;
            ds[u * LA_D + lane] = ok ? ldbf(db + 1LL * (n + u) * ldd) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < LA_U; ++u) mx[u] = warp_max(q[u]);
#pragma unroll
        for (int u = 0; u < LA_U; ++u) pe[u] = __expf(q[u] - mx[u]);
#pragma unroll
        for (int u = 0; u < LA_U; ++u) sv[u] = pe[u] / warp_sum(pe[u]);  // softmax_d(q)[d]
        __syncwarp();
        float dqh[LA_U];  // d(loss)/d(q_hat[d]) = sum_e ctx[d][e] dout[e]
#pragma unroll
        for (int u = 0; u < LA_U; ++u) dqh[u] = 0.f;
#pragma unroll
        for (int e = 0; e < LA_D; e += 4) {
#pragma unroll
            for (int u = 0; u < LA_U; ++u) {
                const float4 d4 = *reinterpret_cast<const float4*>(&ds[u * LA_D + e]);
                const float qh = sv[u] * scale;
                dqh[u] = fmaf(row[e], d4.x, dqh[u]);
                dqh[u] = fmaf(row[e + 1], d4.y, dqh[u]);
                dqh[u] = fmaf(row[e + 2], d4.z, dqh[u]);
                dqh[u] = fmaf(row[e + 3], d4.w, dqh[u]);
                acc[e] = fmaf(qh, d4.x, acc[e]);
                acc[e + 1] = fmaf(qh, d4.y, acc[e + 1]);
                acc[e + 2] = fmaf(qh, d4.z, acc[e + 2]);
                acc[e + 3] = fmaf(qh, d4.w, acc[e + 3]);
            }
        }
        float dot[LA_U];
#pragma unroll
        for (int u = 0; u < LA_U; ++u) dot[u] = warp_sum(scale * dqh[u] * sv[u]);
#pragma unroll
        for (int u = 0; u < LA_U; ++u)
            if (n + u < n1) gq[1LL * (n + u) * ldg] = __float2bfloat16(sv[u] * (scale * dqh[u] - dot[u]));
        __syncwarp();
    }
    float* dst = dpart + ((1LL * (b * heads + h) * gridDim.x + blockIdx.x) * LA_D + lane) * LA_D;
#pragma unroll
    for (int e = 0; e < LA_D; ++e) dst[e] = acc[e];
}

// Combine for backward.  grid (B*heads), 1024 threads (warp = d, lane = e): dctx [B*heads][32][32] = sum over the chunk
// partials (coalesced: consecutive threads read consecutive floats of one partial); r[d] = sum_e dctx[d][e] ctx[d][e].
__global__ void __launch_bounds__(LA_D * LA_D) linattn_dctx_combine_kernel(const float* __restrict__ dpart, int chunks,
                                                                          const float* __restrict__ ctx,
                                                                          float* __restrict__ dctx, float* __restrict__ r) {
    const int bh = blockIdx.x, t = threadIdx.x;
    const float* src = dpart + 1LL * bh * chunks * (LA_D * LA_D) + t;
    float out = 0.f;
#pragma unroll 4
    for (int c = 0; c < chunks; ++c) out += src[1LL * c * (LA_D * LA_D)];
    dctx[1LL * bh * LA_D * LA_D + t] = out;
    const float rr = warp_sum(out * ctx[1LL * bh * LA_D * LA_D + t]);
    if ((t & 31) == 0) r[1LL * bh * LA_D + (t >> 5)] = rr;
}

// Backward pass B2.  grid (pixel blocks, B), blockDim = heads*32: warp = head, lane = d for dk and = e for dv.
//   k_hat[d] = exp(k[d] - max_d) / Z_d;  dv[e] = (1/N) sum_d k_hat[d] dctx[d][e];
//   dk[d] = k_hat[d] * ((1/N) sum_e dctx[d][e] v[e] - r[d]).
__global__ void linattn_bwd_kv_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int N, int heads,
                                      const float* __restrict__ kstat, const float* __restrict__ dctx,
                                      const float* __restrict__ r, __nv_bfloat16* __restrict__ dqkv, long long ldg,
                                      int pix_per_block) {
    extern __shared__ float sm2[];  // [heads][2][32]: k_hat and v staged per warp
    const int b = blockIdx.y, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hidden = heads * LA_D;
    const int bh = b * heads + h;
    float row[LA_D], col[LA_D];  // dctx[lane][e], dctx[d][lane]
#pragma unroll
    for (int i = 0; i < LA_D; ++i) {
        row[i] = dctx[(1LL * bh * LA_D + lane) * LA_D + i];
        col[i] = dctx[(1LL * bh * LA_D + i) * LA_D + lane];
    }
    const float mx = kstat[(1LL * bh * LA_D + lane) * 2], invZ = 1.f / kstat[(1LL * bh * LA_D + lane) * 2 + 1];
    const float rd = r[1LL * bh * LA_D + lane];
    const float invN = 1.f / static_cast<float>(N);
    const int n0 = blockIdx.x * pix_per_block, n1 = min(N, n0 + pix_per_block);
    const __nv_bfloat16* kb = qkv + 1LL * b * N * ld + hidden + h * LA_D + lane;
    const __nv_bfloat16* vb = qkv + 1LL * b * N * ld + 2 * hidden + h * LA_D + lane;
    __nv_bfloat16* gk = dqkv + 1LL * b * N * ldg + hidden + h * LA_D + lane;
    __nv_bfloat16* gv = dqkv + 1LL * b * N * ldg + 2 * hidden + h * LA_D + lane;
    float* ks = sm2 + h * 2 * LA_U * LA_D;  // [LA_U][32] k_hat rows, then [LA_U][32] v rows, per warp
    float* vs = ks + LA_U * LA_D;
    for (int n = n0; n < n1; n += LA_U) {
        float kh[LA_U];
#pragma unroll
        for (int u = 0; u < LA_U; ++u) {
            const bool ok = n + u < n1;
            kh[u] = ok ? __expf(ldbf(kb + 1LL * (n + u) * ld) - mx) * invZ : 0.f;
            ks[u * LA_D + lane] = kh[u];
            vs[u * LA_D + lane] = ok ? ldbf(vb + 1LL * (n + u) * ld) : 0.f;
        }
        __syncwarp();
        float dv[LA_U], t[LA_U];
#pragma unroll
        for (int u = 0; u < LA_U; ++u) dv[u] = t[u] = 0.f;
#pragma unroll
        for (int i = 0; i < LA_D; i += 4) {
#pragma unroll
            for (int u = 0; u < LA_U; ++u) {
                const float4 k4 = *reinterpret_cast<const float4*>(&ks[u * LA_D + i]);
                const float4 v4 = *reinterpret_cast<const float4*>(&vs[u * LA_D + i]);
                dv[u] = fmaf(k4.x, col[i], dv[u]);
                dv[u] = fmaf(k4.y, col[i + 1], dv[u]);
                dv[u] = fmaf(k4.z, col[i + 2], dv[u]);
                dv[u] = fmaf(k4.w, col[i + 3], dv[u]);
                t[u] = fmaf(row[i], v4.x, t[u]);
                t[u] = fmaf(row[i + 1], v4.y, t[u]);
                t[u] = fmaf(row[i + 2], v4.z, t[u]);
                t[u] = fmaf(row[i + 3], v4.w, t[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < LA_U; ++u)
            if (n + u < n1) {
                gv[1LL * (n + u) * ldg] = __float2bfloat16(dv[u] * invN);
                gk[1LL * (n + u) * ldg] = __float2bfloat16(kh[u] * (t[u] * invN - rd));
            }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------ channel LayerNorm
// cond_unet.py LayerNorm (:360-369): y[p][c] = (x[p][c] - mean_p) * rsqrt(var_p + eps) * g[c], statistics over the C channels
// of ONE pixel (biased variance), NHWC bf16 in and out.  HBM-bound: 2 B read + 2 B written per element forward; backward
// reads x and dy, writes dx (6 B) and re-derives mean / rstd from x instead of storing them.
// A pixel is held by gs = C / (8 * NV) lanes (a power of two <= 32), each with NV 16-byte vectors in registers; a warp works on
// 32 / gs pixels at a time and reduces with xor-shuffles inside each lane group.
template <int NV>
__device__ __forceinline__ void cln_load(const __nv_bfloat16* row, int li, int gs, float (&v)[NV][8]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const uint4 raw = *reinterpret_cast<const uint4*>(row + (li + k * gs) * 8);
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[k][2 * j] = __uint_as_float(w[j] << 16);
            v[k][2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
        }
    }
}
__device__ __forceinline__ float cln_group_sum(float a, int gs) {
    for (int o = gs >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    return a;
}
template <int NV>
__device__ __forceinline__ void cln_stats(const float (&v)[NV][8], int gs, float inv_c, float eps, float& mean,
                                          float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
    mean = cln_group_sum(s, gs) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; q += d * d; }
    rstd = rsqrtf(cln_group_sum(q, gs) * inv_c + eps);
}
__device__ __forceinline__ uint4 cln_pack(const float (&o)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(o[2 * j], o[2 * j + 1]);
        w[j] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int NV>
__global__ void __launch_bounds__(256) chan_ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                          long long rows, int C, const float* __restrict__ g, float eps,
                                                          __nv_bfloat16* __restrict__ y, long long ldy) {
    const int gs = (C >> 3) / NV, lane = threadIdx.x & 31;
    const int sub = lane / gs, li = lane % gs, ppw = 32 / gs;
    const long long warp0 = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (1LL * gridDim.x * blockDim.x) >> 5;
    const float inv_c = 1.f / static_cast<float>(C);
    float gg[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) gg[k][j] = g[(li + k * gs) * 8 + j];
    for (long long r0 = warp0 * ppw; r0 < rows; r0 += nwarps * ppw) {
        const long long r = r0 + sub;
        const bool ok = r < rows;  // lanes of a missing pixel still take part in the shuffles
        float v[NV][8];
        cln_load<NV>(x + (ok ? r : rows - 1) * ldx, li, gs, v);
        float mean, rstd;
        cln_stats<NV>(v, gs, inv_c, eps, mean, rstd);
        if (!ok) continue;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * gg[k][j];
            *reinterpret_cast<uint4*>(y + r * ldy + (li + k * gs) * 8) = cln_pack(o);
        }
    }
}

// dx = rstd * (g dy - mean_c(g dy) - xhat * mean_c(g dy xhat)); dg[c] += sum_p dy * xhat (per-thread partials over the
// pixels it visits, then one shared-memory reduction per CTA and C atomics).
template <int NV>
__global__ void __launch_bounds__(256) chan_ln_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long lddy,
                                                          const __nv_bfloat16* __restrict__ x, long long ldx,
                                                          long long rows, int C, const float* __restrict__ g, float eps,
                                                          __nv_bfloat16* __restrict__ dx, long long lddx,
                                                          float* __restrict__ dg) {
    extern __shared__ float cln_sm[];  // [C] dg partials of this CTA
    const int gs = (C >> 3) / NV, lane = threadIdx.x & 31;
    const int sub = lane / gs, li = lane % gs, ppw = 32 / gs;
    const long long warp0 = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (1LL * gridDim.x * blockDim.x) >> 5;
    const float inv_c = 1.f / static_cast<float>(C);
    for (int c = threadIdx.x; c < C; c += blockDim.x) cln_sm[c] = 0.f;
    __syncthreads();
    float gg[NV][8], acc[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) { gg[k][j] = g[(li + k * gs) * 8 + j]; acc[k][j] = 0.f; }
    for (long long r0 = warp0 * ppw; r0 < rows; r0 += nwarps * ppw) {
        const long long r = r0 + sub;
        const bool ok = r < rows;
        const long long rr = ok ? r : rows - 1;
        float v[NV][8], d[NV][8];
        cln_load<NV>(x + rr * ldx, li, gs, v);
        cln_load<NV>(dy + rr * lddy, li, gs, d);
        float mean, rstd;
        cln_stats<NV>(v, gs, inv_c, eps, mean, rstd);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[k][j] = (v[k][j] - mean) * rstd;  // xhat
                if (ok) acc[k][j] += d[k][j] * v[k][j];
                d[k][j] *= gg[k][j];                // g dy
                s1 += d[k][j];
                s2 += d[k][j] * v[k][j];
            }
        s1 = cln_group_sum(s1, gs) * inv_c;
        s2 = cln_group_sum(s2, gs) * inv_c;
        if (!ok) continue;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rstd * (d[k][j] - s1 - v[k][j] * s2);
            *reinterpret_cast<uint4*>(dx + r * lddx + (li + k * gs) * 8) = cln_pack(o);
        }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(cln_sm + (li + k * gs) * 8 + j, acc[k][j]);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dg + c, cln_sm[c]);
}

}  // namespace adm

using namespace adm;
typedef __nv_bfloat16 bf16;

extern "C" {

int adm_ws_pack(const float* w, void* wpk, float* stats, int cout, int cin, int ksize, float eps, void* stream) {
    if (cout <= 0 || cin <= 0 || ksize <= 0) { set_error("ws_pack: bad shape"); return ADM_ERR_SHAPE; }
    const int kpad = (cin + 63) / 64 * 64;
    ws_pack_kernel<<<cout, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<bf16*>(wpk), stats, cin,
                                                                       ksize * ksize, kpad, eps);
    ADM_CHECK_LAUNCH("ws_pack");
    return 0;
}

int adm_ws_pack_bwd(const float* dw_packed, const float* w, const float* stats, float* dw, int cout, int cin, int ksize,
                    int accumulate, void* stream) {
    if (cout <= 0 || cin <= 0 || ksize <= 0) { set_error("ws_pack_bwd: bad shape"); return ADM_ERR_SHAPE; }
    const int kpad = (cin + 63) / 64 * 64;
    ws_pack_bwd_kernel<<<cout, 256, 0, static_cast<cudaStream_t>(stream)>>>(dw_packed, w, stats, dw, cin, ksize * ksize,
                                                                           kpad, accumulate);
    ADM_CHECK_LAUNCH("ws_pack_bwd");
    return 0;
}

static int la_chunks(int n_pix, int batch_heads) {
    long long want = (4LL * num_sms() + batch_heads - 1) / batch_heads;
    long long max_chunks = (n_pix + 8 * LA_WARPS - 1) / (8 * LA_WARPS);
    if (want > max_chunks) want = max_chunks;
    if (want > 64) want = 64;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

// Backward pass B1 runs one CTA of `heads` warps per (pixel chunk, sample): enough chunks for ~8 CTAs per SM, at least
// 16 iterations of LA_U pixels each.
static int la_chunks_q(int n_pix, int batch) {
    long long want = (8LL * num_sms() + batch - 1) / batch;
    long long max_chunks = (n_pix + 16 * LA_U - 1) / (16 * LA_U);
    if (want > max_chunks) want = max_chunks;
    if (want > 256) want = 256;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

int adm_linattn_workspace(int batch, int heads, int n_pix, long long* floats) {
    const int chunks = la_chunks(n_pix, batch * heads);
    const long long fwd = 1LL * batch * heads * chunks * LA_D * (LA_D + 2);
    const long long bwd = 1LL * batch * heads * la_chunks_q(n_pix, batch) * LA_D * LA_D;
    *floats = fwd > bwd ? fwd : bwd;
    return chunks;
}

int adm_linattn_fwd(const void* qkv, long long ld, int batch, int n_pix, int heads, int dim_head, float scale, void* out,
                    long long ldo, float* ctx, float* kstat, float* work, void* stream) {
    if (dim_head != LA_D) { set_error("linattn: dim_head must be 32 (got %d)", dim_head); return ADM_ERR_SHAPE; }
    if (heads < 1 || heads > 32 || batch <= 0 || n_pix <= 0) { set_error("linattn: bad shape"); return ADM_ERR_SHAPE; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = la_chunks(n_pix, batch * heads);
    linattn_ctx_partial_kernel<<<dim3(chunks, batch * heads), LA_WARPS * 32, 0, s>>>(static_cast<const bf16*>(qkv), ld,
                                                                                     n_pix, heads, work);
    ADM_CHECK_LAUNCH("linattn_ctx_partial");
    linattn_ctx_combine_kernel<<<batch * heads, 32, 0, s>>>(work, chunks, n_pix, ctx, kstat);
    ADM_CHECK_LAUNCH("linattn_ctx_combine");
    if (out != nullptr) {
        int blocks = (8 * num_sms() + batch - 1) / batch;
        int ppb = (n_pix + blocks - 1) / blocks;
        if (ppb < 4) ppb = 4;
        blocks = (n_pix + ppb - 1) / ppb;
        linattn_apply_kernel<<<dim3(blocks, batch), heads * 32, heads * LA_U * LA_D * sizeof(float), s>>>(
            static_cast<const bf16*>(qkv), ld, n_pix, heads, ctx, scale, static_cast<bf16*>(out), ldo, ppb);
        ADM_CHECK_LAUNCH("linattn_apply");
    }
    return 0;
}

int adm_linattn_bwd(const void* qkv, long long ld, int batch, int n_pix, int heads, int dim_head, float scale,
                    const void* dout, long long ldd, const float* ctx, const float* kstat, float* dctx, float* r,
                    float* work, void* dqkv, long long ldg, void* stream) {
    if (dim_head != LA_D) { set_error("linattn: dim_head must be 32 (got %d)", dim_head); return ADM_ERR_SHAPE; }
    if (heads < 1 || heads > 32 || batch <= 0 || n_pix <= 0) { set_error("linattn: bad shape"); return ADM_ERR_SHAPE; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunks = la_chunks_q(n_pix, batch);
    linattn_bwd_q_kernel<<<dim3(chunks, batch), heads * 32, heads * LA_U * LA_D * sizeof(float), s>>>(
        static_cast<const bf16*>(qkv), ld, n_pix, heads, ctx, scale, static_cast<const bf16*>(dout), ldd,
        static_cast<bf16*>(dqkv), ldg, work);
    ADM_CHECK_LAUNCH("linattn_bwd_q");
    linattn_dctx_combine_kernel<<<batch * heads, LA_D * LA_D, 0, s>>>(work, chunks, ctx, dctx, r);
    ADM_CHECK_LAUNCH("linattn_dctx_combine");
    int blocks = (8 * num_sms() + batch - 1) / batch;
    int ppb = (n_pix + blocks - 1) / blocks;
    if (ppb < 4) ppb = 4;
    blocks = (n_pix + ppb - 1) / ppb;
    linattn_bwd_kv_kernel<<<dim3(blocks, batch), heads * 32, heads * 2 * LA_U * LA_D * sizeof(float), s>>>(
        static_cast<const bf16*>(qkv), ld, n_pix, heads, kstat, dctx, r, static_cast<bf16*>(dqkv), ldg, ppb);
    ADM_CHECK_LAUNCH("linattn_bwd_kv");
    return 0;
}

// vectors per lane for a channel count: C / 8 vectors spread over gs = C / (8 NV) lanes, gs a power of two <= 32
static int cln_nv(int c) {
    if (c <= 0 || c % 8) return 0;
    const int nvec = c / 8;
    for (int nv = 1; nv <= 4; nv *= 2) {
        if (nvec % nv) continue;
        const int gs = nvec / nv;
        if (gs <= 32 && (gs & (gs - 1)) == 0) return nv;
    }
    return 0;
}

int adm_chan_layernorm_ok(int c) { return cln_nv(c) != 0; }

int adm_chan_layernorm_fwd(const void* x, long long ldx, long long rows, int c, const float* g, float eps, void* y,
                           long long ldy, void* stream) {
    const int nv = cln_nv(c);
    if (nv == 0 || rows <= 0) { set_error("chan_layernorm: C = %d must be 8 * 2^k * {1,2,4} with 2^k <= 32", c); return ADM_ERR_SHAPE; }
    if (ldx % 8 || ldy % 8) { set_error("chan_layernorm: pixel strides must be multiples of 8"); return ADM_ERR_SHAPE; }
    const int ppw = 32 / ((c / 8) / nv);
    long long blocks = (rows + 8LL * ppw - 1) / (8LL * ppw);
    if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bf16* xp = static_cast<const bf16*>(x);
    bf16* yp = static_cast<bf16*>(y);
    if (nv == 1) chan_ln_fwd_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, s>>>(xp, ldx, rows, c, g, eps, yp, ldy);
    else if (nv == 2) chan_ln_fwd_kernel<2><<<static_cast<unsigned>(blocks), 256, 0, s>>>(xp, ldx, rows, c, g, eps, yp, ldy);
    else chan_ln_fwd_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, s>>>(xp, ldx, rows, c, g, eps, yp, ldy);
    ADM_CHECK_LAUNCH("chan_layernorm_fwd");
    return 0;
}

int adm_chan_layernorm_bwd(const void* dy, long long lddy, const void* x, long long ldx, long long rows, int c,
                           const float* g, float eps, void* dx, long long lddx, float* dg, void* stream) {
    const int nv = cln_nv(c);
    if (nv == 0 || rows <= 0) { set_error("chan_layernorm: C = %d must be 8 * 2^k * {1,2,4} with 2^k <= 32", c); return ADM_ERR_SHAPE; }
    if (ldx % 8 || lddy % 8 || lddx % 8) { set_error("chan_layernorm: pixel strides must be multiples of 8"); return ADM_ERR_SHAPE; }
    const int ppw = 32 / ((c / 8) / nv);
    long long blocks = (rows + 8LL * ppw - 1) / (8LL * ppw);
    if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const bf16 *dyp = static_cast<const bf16*>(dy), *xp = static_cast<const bf16*>(x);
    bf16* dxp = static_cast<bf16*>(dx);
    const size_t sm = sizeof(float) * c;
    const unsigned b = static_cast<unsigned>(blocks);
    if (nv == 1) chan_ln_bwd_kernel<1><<<b, 256, sm, s>>>(dyp, lddy, xp, ldx, rows, c, g, eps, dxp, lddx, dg);
    else if (nv == 2) chan_ln_bwd_kernel<2><<<b, 256, sm, s>>>(dyp, lddy, xp, ldx, rows, c, g, eps, dxp, lddx, dg);
    else chan_ln_bwd_kernel<4><<<b, 256, sm, s>>>(dyp, lddy, xp, ldx, rows, c, g, eps, dxp, lddx, dg);
    ADM_CHECK_LAUNCH("chan_layernorm_bwd");
    return 0;
}

}  // extern "C"
