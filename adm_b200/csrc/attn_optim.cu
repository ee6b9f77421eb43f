// Row softmax forward/backward for the attention products (unet/uncond_unet.py:207), the rank-1 SpatialAtt gate of the
// decouple branches (unet/uncond_unet.py:19-37), and the flat-arena optimizer kernels (gradient norm, clip + AdamW:
// train_uncond_dpm.py:179-180,292,296).
#include "adm_internal.h"
#include <cuda_bf16.h>
#include <stdint.h>

namespace adm {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// P[row] = softmax(S[row]) ; one warp per row, L <= 1024 (32 values per lane in registers).
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p,
                                                          long long rows, int L) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (1LL * gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < rows; r += nwarps) {
        const float* sr = s + r * L;
        float v[32];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int j = lane + 32 * i;
            v[i] = j < L ? sr[j] : -INFINITY;
            m = fmaxf(m, v[i]);
        }
        m = warp_max(m);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            v[i] = (lane + 32 * i) < L ? __expf(v[i] - m) : 0.f;
            sum += v[i];
        }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        __nv_bfloat16* pr = p + r * L;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int j = lane + 32 * i;
            if (j < L) pr[j] = __float2bfloat16(v[i] * inv);
        }
    }
}

// Rows longer than 1024 (the single-head mid attention of the frozen autoencoder, 4096 tokens): one CTA per row, three
// passes over the row (maximum, sum of exponentials, normalised write); the 16 KB row stays in L1 / L2 between passes.
__global__ void __launch_bounds__(256) softmax_fwd_long_kernel(const float* __restrict__ s,
                                                               __nv_bfloat16* __restrict__ p, long long rows, int L) {
    __shared__ float red[8];
    __shared__ float bcast;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const float4* sr = reinterpret_cast<const float4*>(s + r * L);
        const int L4 = L >> 2;  // L is a multiple of 4 (checked by the launcher)
        float m = -INFINITY;
        for (int j = threadIdx.x; j < L4; j += 256) {
            const float4 v = sr[j];
            m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
        }
        m = warp_max(m);
        if (lane == 0) red[warp] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = red[0];
            for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
            bcast = t;
        }
        __syncthreads();
        m = bcast;
        float sum = 0.f;
        for (int j = threadIdx.x; j < L4; j += 256) {
            const float4 v = sr[j];
            sum += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
        }
        sum = warp_sum(sum);
        __syncthreads();  // everyone has read bcast
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; ++i) t += red[i];
            bcast = 1.f / t;
        }
        __syncthreads();
        const float inv = bcast;
        uint2* pr = reinterpret_cast<uint2*>(p + r * L);
        for (int j = threadIdx.x; j < L4; j += 256) {
            const float4 v = sr[j];
            const __nv_bfloat162 a = __floats2bfloat162_rn(__expf(v.x - m) * inv, __expf(v.y - m) * inv);
            const __nv_bfloat162 b = __floats2bfloat162_rn(__expf(v.z - m) * inv, __expf(v.w - m) * inv);
            pr[j] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        }
        __syncthreads();  // red / bcast are re-used by the next row
    }
}

// dS[row] = scale * P * (dP - sum_j dP_j P_j)
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p,
                                                          const float* __restrict__ dp, __nv_bfloat16* __restrict__ ds,
                                                          float scale, long long rows, int L) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * 1LL * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = (1LL * gridDim.x * blockDim.x) >> 5;
    for (long long r = warp0; r < rows; r += nwarps) {
        float pv[32], gv[32];
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int j = lane + 32 * i;
            pv[i] = j < L ? __bfloat162float(p[r * L + j]) : 0.f;
            gv[i] = j < L ? dp[r * L + j] : 0.f;
            dot += pv[i] * gv[i];
        }
        dot = warp_sum(dot);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int j = lane + 32 * i;
            if (j < L) ds[r * L + j] = __float2bfloat16(scale * pv[i] * (gv[i] - dot));
        }
    }
}

// ------------------------------------------------------------------------------------------------ SpatialAtt
// One CTA per image, HW <= 1024 (dynamic smem: 6*HW floats).  pr = {b_map, wq, bq, wk, bk} read from device scalars.
// out = softsign(o) * h + res,  o_i = sum_j softmax_j(q_i k_j) att_j,  att = h . w_map + b_map.
__global__ void __launch_bounds__(256) spatial_att_fwd_kernel(const __nv_bfloat16* __restrict__ h, long long ldh,
                                                              const __nv_bfloat16* __restrict__ res, long long ldr,
                                                              const float* __restrict__ w_map,
                                                              const float* __restrict__ scal, int HW, int C,
                                                              __nv_bfloat16* __restrict__ out, long long ldo,
                                                              float* __restrict__ att_save,
                                                              float* __restrict__ o_save) {
    extern __shared__ float sm[];
    float* att = sm;
    float* ov = sm + HW;
    const int n = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float b_map = scal[0], wq = scal[1], bq = scal[2], wk = scal[3], bk = scal[4];
    for (int i = warp; i < HW; i += nw) {
        const __nv_bfloat16* hp = h + (1LL * n * HW + i) * ldh;
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a += __bfloat162float(hp[c]) * w_map[c];
        a = warp_sum(a);
        if (lane == 0) att[i] = a + b_map;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float q = wq * att[i] + bq;
        float m = -INFINITY;
        for (int j = 0; j < HW; ++j) m = fmaxf(m, q * (wk * att[j] + bk));
        float z = 0.f, acc = 0.f;
        for (int j = 0; j < HW; ++j) {
            const float e = __expf(q * (wk * att[j] + bk) - m);
            z += e;
            acc += e * att[j];
        }
        ov[i] = acc / z;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        att_save[1LL * n * HW + i] = att[i];
        o_save[1LL * n * HW + i] = ov[i];
    }
    for (int idx = threadIdx.x; idx < HW * C; idx += blockDim.x) {
        const int i = idx / C, c = idx % C;
        const float o = ov[i];
        const float s = o / (1.f + fabsf(o));
        const float v = s * __bfloat162float(h[(1LL * n * HW + i) * ldh + c]) +
                        __bfloat162float(res[(1LL * n * HW + i) * ldr + c]);
        out[(1LL * n * HW + i) * ldo + c] = __float2bfloat16(v);
    }
}

// dscal: {db_map, dwq, dbq, dwk, dbk} accumulated atomically; dw_map[C] accumulated atomically.
__global__ void __launch_bounds__(256) spatial_att_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long ldy,
                                                              const __nv_bfloat16* __restrict__ h, long long ldh,
                                                              const float* __restrict__ w_map,
                                                              const float* __restrict__ scal,
                                                              const float* __restrict__ att_save,
                                                              const float* __restrict__ o_save, int HW, int C,
                                                              __nv_bfloat16* __restrict__ dh, long long lddh,
                                                              float* __restrict__ dw_map, float* __restrict__ dscal) {
    extern __shared__ float sm[];
    float* att = sm;
    float* ov = att + HW;
    float* dov = ov + HW;   // d o_i
    float* mx = dov + HW;   // row max
    float* zz = mx + HW;    // row normaliser
    float* da = zz + HW;    // total d att_i
    __shared__ float red[5];
    const int n = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const float wq = scal[1], bq = scal[2], wk = scal[3], bk = scal[4];
    if (threadIdx.x < 5) red[threadIdx.x] = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        att[i] = att_save[1LL * n * HW + i];
        ov[i] = o_save[1LL * n * HW + i];
    }
    __syncthreads();
    // ds_i = sum_c dy_ic h_ic ; d o_i = ds_i / (1 + |o_i|)^2
    for (int i = warp; i < HW; i += nw) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32)
            a += __bfloat162float(dy[(1LL * n * HW + i) * ldy + c]) * __bfloat162float(h[(1LL * n * HW + i) * ldh + c]);
        a = warp_sum(a);
        if (lane == 0) {
            const float d = 1.f + fabsf(ov[i]);
            dov[i] = a / (d * d);
        }
    }
    __syncthreads();
    // phase A (thread i): softmax row stats, dq_i
    float l_dwq = 0.f, l_dbq = 0.f, l_dwk = 0.f, l_dbk = 0.f, l_dbm = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        const float q = wq * att[i] + bq;
        float m = -INFINITY;
        for (int j = 0; j < HW; ++j) m = fmaxf(m, q * (wk * att[j] + bk));
        float z = 0.f;
        for (int j = 0; j < HW; ++j) z += __expf(q * (wk * att[j] + bk) - m);
        mx[i] = m;
        zz[i] = z;
        float dq = 0.f;
        for (int j = 0; j < HW; ++j) {
            const float k = wk * att[j] + bk;
            const float pij = __expf(q * k - m) / z;
            dq += pij * dov[i] * (att[j] - ov[i]) * k;
        }
        da[i] = wq * dq;
        l_dwq += dq * att[i];
        l_dbq += dq;
    }
    __syncthreads();
    // phase B (thread j): dk_j, direct d att_j
    for (int j = threadIdx.x; j < HW; j += blockDim.x) {
        const float k = wk * att[j] + bk;
        float dk = 0.f, dadir = 0.f;
        for (int i = 0; i < HW; ++i) {
            const float q = wq * att[i] + bq;
            const float pij = __expf(q * k - mx[i]) / zz[i];
            dk += pij * dov[i] * (att[j] - ov[i]) * q;
            dadir += pij * dov[i];
        }
        const float tot = da[j] + wk * dk + dadir;
        da[j] = tot;
        l_dwk += dk * att[j];
        l_dbk += dk;
        l_dbm += tot;
    }
    atomicAdd(&red[0], l_dbm);
    atomicAdd(&red[1], l_dwq);
    atomicAdd(&red[2], l_dbq);
    atomicAdd(&red[3], l_dwk);
    atomicAdd(&red[4], l_dbk);
    __syncthreads();
    if (threadIdx.x < 5) atomicAdd(dscal + threadIdx.x, red[threadIdx.x]);
    // dh_ic = s_i dy_ic + da_i w_c ; dw_c += sum_i da_i h_ic
    for (int idx = threadIdx.x; idx < HW * C; idx += blockDim.x) {
        const int i = idx / C, c = idx % C;
        const float o = ov[i];
        const float s = o / (1.f + fabsf(o));
        const float v = s * __bfloat162float(dy[(1LL * n * HW + i) * ldy + c]) + da[i] * w_map[c];
        dh[(1LL * n * HW + i) * lddh + c] = __float2bfloat16(v);
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int i = 0; i < HW; ++i) a += da[i] * __bfloat162float(h[(1LL * n * HW + i) * ldh + c]);
        atomicAdd(dw_map + c, a);
    }
}

// ------------------------------------------------------------------------------------------------ optimizer
__global__ void __launch_bounds__(256) sq_norm_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
    float acc = 0.f;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n4; i += 1LL * gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (long long i = (n4 << 2) + blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x)
        acc += g[i] * g[i];
    acc = warp_sum(acc);
    __shared__ float ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (blockDim.x >> 5); ++i) s += ws[i];
        atomicAdd(out, s);
    }
}

// AdamW (decoupled weight decay, torch.optim.AdamW semantics) with the clip coefficient computed on the device:
// coef = min(1, max_norm / (sqrt(sqnorm * gscale^2) + 1e-6)); g = grad * gscale * coef.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                    float beta1, float beta2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, float gscale, float max_norm,
                                                    const float* __restrict__ sqnorm,
                                                    const float* __restrict__ hyper,
                                                    __nv_bfloat16* __restrict__ p_bf16) {
    if (hyper != nullptr) {  // CUDA-graph friendly: step-dependent scalars live in device memory
        lr = hyper[0];
        bc1 = hyper[1];
        bc2_sqrt = hyper[2];
    }
    float coef = gscale;
    if (sqnorm != nullptr && max_norm > 0.f) {
        const float norm = sqrtf(*sqnorm) * gscale;
        coef = gscale * fminf(1.f, max_norm / (norm + 1e-6f));
    }
    const float step = lr / bc1;
    const float decay = 1.f - lr * wd;
    const float inv_bc2 = 1.f / bc2_sqrt;
    auto upd = [&](float gi, float& pi, float& mi, float& vi) {
        gi *= coef;
        pi *= decay;
        mi = beta1 * mi + (1.f - beta1) * gi;
        vi = beta2 * vi + (1.f - beta2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi = pi - step * mi / denom;
    };
    (void)inv_bc2;
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0) &&
                     (reinterpret_cast<uintptr_t>(p_bf16) % 8 == 0);
    if (vec) {
        const long long n4 = n / 4;
        for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n4; i += 1LL * gridDim.x * blockDim.x) {
            const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g) + i);
            float4 p4 = reinterpret_cast<float4*>(p)[i];
            float4 m4 = reinterpret_cast<float4*>(m)[i];
            float4 v4 = reinterpret_cast<float4*>(v)[i];
            upd(g4.x, p4.x, m4.x, v4.x);
            upd(g4.y, p4.y, m4.y, v4.y);
            upd(g4.z, p4.z, m4.z, v4.z);
            upd(g4.w, p4.w, m4.w, v4.w);
            reinterpret_cast<float4*>(p)[i] = p4;
            reinterpret_cast<float4*>(m)[i] = m4;
            reinterpret_cast<float4*>(v)[i] = v4;
            if (p_bf16 != nullptr) {  // bf16 shadow of the parameters (the GEMM operands of the next step)
                const __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
                uint2 w;
                w.x = *reinterpret_cast<const uint32_t*>(&lo);
                w.y = *reinterpret_cast<const uint32_t*>(&hi);
                reinterpret_cast<uint2*>(p_bf16)[i] = w;
            }
        }
        return;
    }
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        upd(g[i], pi, mi, vi);
        p[i] = pi;
        m[i] = mi;
        v[i] = vi;
        if (p_bf16 != nullptr) p_bf16[i] = __float2bfloat16(pi);
    }
}

// EMA of the parameter arena (ddm/ema.py:141-156 as one pass): dst += w * (src - dst); 12 B per element, HBM-bound.
__global__ void __launch_bounds__(256) lerp_f32_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                       long long n, float w) {
    const long long n4 = n >> 2;
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (long long i = blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n4; i += 1LL * gridDim.x * blockDim.x) {
        float4 a = d4[i];
        const float4 b = __ldcs(s4 + i);
        a.x = fmaf(w, b.x - a.x, a.x);
        a.y = fmaf(w, b.y - a.y, a.y);
        a.z = fmaf(w, b.z - a.z, a.z);
        a.w = fmaf(w, b.w - a.w, a.w);
        d4[i] = a;
    }
    for (long long i = (n4 << 2) + blockIdx.x * 1LL * blockDim.x + threadIdx.x; i < n; i += 1LL * gridDim.x * blockDim.x)
        dst[i] = fmaf(w, src[i] - dst[i], dst[i]);
}

static int ew_blocks(long long work, int per_sm) {
    long long b = (work + 255) / 256;
    const long long cap = 1LL * num_sms() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

}  // namespace adm

using namespace adm;
typedef __nv_bfloat16 bf16;

extern "C" {

int adm_softmax_fwd(const float* s, void* p, long long rows, int len, void* stream) {
    if (len > 1024) {
        if (len % 4 || len > 65536) { set_error("softmax: long rows must be a multiple of 4 and <= 65536 (got %d)", len); return ADM_ERR_SHAPE; }
        const long long cap = 16LL * num_sms();
        softmax_fwd_long_kernel<<<static_cast<unsigned>(rows < cap ? rows : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            s, static_cast<bf16*>(p), rows, len);
        ADM_CHECK_LAUNCH("softmax_fwd_long");
        return 0;
    }
    if (len <= 0 || len > 1024) { set_error("softmax: row length %d not in [1, 1024]", len); return ADM_ERR_SHAPE; }
    softmax_fwd_kernel<<<ew_blocks(rows * 32, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        s, static_cast<bf16*>(p), rows, len);
    ADM_CHECK_LAUNCH("softmax_fwd");
    return 0;
}

int adm_softmax_bwd(const void* p, const float* dp, void* ds, float scale, long long rows, int len, void* stream) {
    if (len <= 0 || len > 1024) { set_error("softmax: row length %d not in [1, 1024]", len); return ADM_ERR_SHAPE; }
    softmax_bwd_kernel<<<ew_blocks(rows * 32, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(p), dp, static_cast<bf16*>(ds), scale, rows, len);
    ADM_CHECK_LAUNCH("softmax_bwd");
    return 0;
}

int adm_spatial_att_fwd(const void* h, long long ldh, const void* res, long long ldr, const float* w_map,
                        const float* scalars, int n, int hw, int c, void* out, long long ldo, float* att_save,
                        float* o_save, void* stream) {
    if (hw <= 0 || hw > 1024) { set_error("spatial_att: HW %d not in [1, 1024]", hw); return ADM_ERR_SHAPE; }
    spatial_att_fwd_kernel<<<n, 256, 2 * hw * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(h), ldh, static_cast<const bf16*>(res), ldr, w_map, scalars, hw, c,
        static_cast<bf16*>(out), ldo, att_save, o_save);
    ADM_CHECK_LAUNCH("spatial_att_fwd");
    return 0;
}

int adm_spatial_att_bwd(const void* dy, long long ldy, const void* h, long long ldh, const float* w_map,
                        const float* scalars, const float* att_save, const float* o_save, int n, int hw, int c,
                        void* dh, long long lddh, float* dw_map, float* dscalars, void* stream) {
    if (hw <= 0 || hw > 1024) { set_error("spatial_att: HW %d not in [1, 1024]", hw); return ADM_ERR_SHAPE; }
    spatial_att_bwd_kernel<<<n, 256, 6 * hw * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(dy), ldy, static_cast<const bf16*>(h), ldh, w_map, scalars, att_save, o_save, hw, c,
        static_cast<bf16*>(dh), lddh, dw_map, dscalars);
    ADM_CHECK_LAUNCH("spatial_att_bwd");
    return 0;
}

int adm_sq_norm(const float* g, long long numel, float* out, void* stream) {
    if ((reinterpret_cast<uintptr_t>(g) & 15) != 0) { set_error("sq_norm: pointer must be 16 B aligned"); return ADM_ERR_SHAPE; }
    sq_norm_kernel<<<ew_blocks(numel / 4 + 1, 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(g, numel, out);
    ADM_CHECK_LAUNCH("sq_norm");
    return 0;
}

int adm_adamw(float* p, const float* g, float* m, float* v, long long numel, float lr, float beta1, float beta2,
              float eps, float weight_decay, int step, float grad_scale, float max_norm, const float* sqnorm,
              const float* hyper_dev, void* p_bf16, void* stream) {
    if (step < 1) { set_error("adamw: step must be >= 1"); return ADM_ERR_SHAPE; }
    const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
    const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
    adamw_kernel<<<ew_blocks(numel, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, g, m, v, numel, lr, beta1, beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale, max_norm, sqnorm, hyper_dev,
        static_cast<__nv_bfloat16*>(p_bf16));
    ADM_CHECK_LAUNCH("adamw");
    return 0;
}

int adm_lerp_f32(float* dst, const float* src, long long numel, float weight, void* stream) {
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) != 0) {
        set_error("lerp_f32: pointers must be 16 B aligned");
        return ADM_ERR_SHAPE;
    }
    if (numel <= 0) return 0;
    lerp_f32_kernel<<<ew_blocks(numel / 4 + 1, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, numel, weight);
    ADM_CHECK_LAUNCH("lerp_f32");
    return 0;
}

}  // extern "C"
