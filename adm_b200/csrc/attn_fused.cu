// K9 forward: fused softmax attention over pixels for the low-resolution UNet levels (reference unet/uncond_unet.py:204-208:
// w = softmax_k(q^T k / sqrt(d)); a = v w^T), one CTA per (sample, head), everything on chip:
//   TMA        : Q [N x 64], K [N x 64], V [N x 64] (bf16, 128B-swizzled) straight out of the (q | k | v) x head x d
//                projection output;
//   tcgen05    : S = Q K^T  (M = 128 query rows per tile, N = keys <= 256, K = 64) -> fp32 in TMEM (2 x 256 columns);
//   softmax    : one thread per query row, three sweeps over its TMEM row (max, sum, normalise), P written as bf16 into
//                the K-major swizzled A-operand layout in shared memory (and to HBM when the backward needs it);
//   tcgen05    : O = P V    (V consumed MN-major), accumulating over the S columns it replaces;
//   epilogue   : TMEM -> bf16 -> [B, N, C] with this head's 64 channels.
// The same kernel in backward mode (MODE 1) fuses three of the five backward products: dP = dO V^T in TMEM, per row
// dS = scale * P o (dP - sum_j dP_j P_j) with P read back as bf16, dS written as the swizzled A operand (and to HBM for
// the dK product), and dQ = dS K as the second MMA — dP never exists in HBM.  dV = P^T dO and dK = dS^T Q stay batched
// GEMMs of the shared engine.
// Neither S nor P round-trips through HBM in inference; in training only the normalised P is stored (bf16) because the
// backward kernels consume it.  N (pixels) in {16, 64, 256}; head dim 64 (32 runs zero-padded, see cond_unet.Attention).
#include "adm_internal.h"
#include "ptx.cuh"

namespace adm {

constexpr int AF_THREADS = 256;           // warps 0-3: query tile 0, warps 4-7: query tile 1
constexpr int AF_TILE = 16384;            // 128 rows x 128 B
constexpr int AF_SMEM_Q = 0;              // 2 tiles
constexpr int AF_SMEM_K = 2 * AF_TILE;    // 256 rows x 128 B
constexpr int AF_SMEM_V = 4 * AF_TILE;
constexpr int AF_SMEM_P = 6 * AF_TILE;    // 2 query tiles x 4 key chunks x 16 KB
constexpr int AF_SMEM_BAR = 14 * AF_TILE;
constexpr int AF_SMEM_TOTAL = AF_SMEM_BAR + 128 + 1024;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnParams {
    int n_pix;       // queries = keys
    int heads, c;    // C = heads * 64
    float scale;     // softmax(scale * q.k)
    __nv_bfloat16* out;    // fwd: O [B, n_pix, ld_out] head slice h*64; bwd: dQ into dqkv (ld_out = 3C)
    long long ld_out;
    __nv_bfloat16* p_out;  // fwd: normalised P [B*heads, n_pix, n_pix] or null; bwd: dS (same shape)
    const __nv_bfloat16* p_in;  // bwd: the saved P
    int a_col, b1_col, b2_col;  // channel offsets (before + h*64) of the three operands in their tensors
};

template <int MODE>  // 0: forward (Q, K, V -> P, O);  1: backward (dO, V, K, P -> dS, dQ)
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                  const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + AF_SMEM_BAR);
    uint64_t* bar_s = bar_load + 1;  // [2]
    uint64_t* bar_o = bar_load + 3;  // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_load + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.x, b = bh / p.heads, h = bh % p.heads;
    const int N = p.n_pix;
    const int m_tiles = (N + 127) / 128;       // 1 or 2
    const int q_rows = m_tiles * 128;          // TMA box rows of Q (rows >= N are zero-filled)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmKV);
        mbar_init(bar_load, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_o[i], 1); }
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (threadIdx.x == 0) {
        // ---- loads: Q (q_rows x 64), K, V (N x 64) of this (sample, head)
        mbar_expect_tx(bar_load, (q_rows + 2 * N) * 128);
        tma_load_3d(smem + AF_SMEM_Q, &tmQ, bar_load, p.a_col + h * 64, 0, b);
        tma_load_3d(smem + AF_SMEM_K, &tmKV, bar_load, p.b1_col + h * 64, 0, b);
        tma_load_3d(smem + AF_SMEM_V, &tmKV, bar_load, p.b2_col + h * 64, 0, b);
        mbar_wait(bar_load, 0, 11);
        tc_fence_after();
        // ---- S[mt] = Q[mt] K^T  (K-major x K-major, 4 k-steps of 16)
        const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
        const uint32_t sq = smem_u32(smem + AF_SMEM_Q), sk = smem_u32(smem + AF_SMEM_K);
        for (int mt = 0; mt < m_tiles; ++mt) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem_base + mt * 256, make_smem_desc(sq + mt * AF_TILE + k * 32u, 16u, 1024),
                          make_smem_desc(sk + k * 32u, 16u, 1024), idesc, k > 0 ? 1u : 0u);
            umma_commit(&bar_s[mt]);
        }
    }

    // ---- softmax: thread -> one query row of tile mt (TMEM lane = row inside the tile)
    const int mt = warp >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;            // row inside the tile
    const int q = mt * 128 + row;                // query index
    const bool tile_ok = mt < m_tiles;
    const bool row_ok = tile_ok && q < N;
    const float c1 = p.scale * 1.4426950408889634f;  // exp(scale * s) = exp2(c1 * s)
    float inv_sum = 0.f;
    if (tile_ok) {
        mbar_wait(&bar_s[mt], 0, 12);
        tc_fence_after();
        const uint32_t taddr = tmem_base + mt * 256 + (static_cast<uint32_t>(quad * 32) << 16);
        if (MODE == 0) {
        float mx = -INFINITY;
        for (int c0 = 0; c0 < N; c0 += 64) {
            uint32_t v[4][16];
            const int nsub = min(4, (N - c0) >> 4);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) tmem_ld_x16(taddr + c0 + 16 * s, v[s]);
            tmem_ld_wait();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(v[s][j]));
                }
        }
        const float mc = mx * c1;
        float sum = 0.f;
        for (int c0 = 0; c0 < N; c0 += 64) {
            uint32_t v[4][16];
            const int nsub = min(4, (N - c0) >> 4);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) tmem_ld_x16(taddr + c0 + 16 * s, v[s]);
            tmem_ld_wait();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) sum += ex2(fmaf(__uint_as_float(v[s][j]), c1, -mc));
                }
        }
        inv_sum = 1.f / sum;
        uint8_t* sp = smem + AF_SMEM_P + mt * 4 * AF_TILE;
        __nv_bfloat16* pg = (p.p_out != nullptr && row_ok) ? p.p_out + (1LL * bh * N + q) * N : nullptr;
        for (int c0 = 0; c0 < N; c0 += 64) {
            uint32_t v[4][16];
            const int nsub = min(4, (N - c0) >> 4);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) tmem_ld_x16(taddr + c0 + 16 * s, v[s]);
            tmem_ld_wait();
            uint8_t* chunk = sp + (c0 >> 6) * AF_TILE + row * 128;  // this row inside key chunk c0/64
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) {
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float p0 = ex2(fmaf(__uint_as_float(v[s][2 * j]), c1, -mc)) * inv_sum;
                        const float p1 = ex2(fmaf(__uint_as_float(v[s][2 * j + 1]), c1, -mc)) * inv_sum;
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(p0, p1);
                        w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    // 16-byte units 2s, 2s+1 of the 128 B row, XOR-swizzled with the row index (SWIZZLE_128B)
                    *reinterpret_cast<uint4*>(chunk + (((2 * s) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(chunk + (((2 * s + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                    if (pg != nullptr) {
                        *reinterpret_cast<uint4*>(pg + c0 + 16 * s) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(pg + c0 + 16 * s + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
        }
        } else {
        // dS = scale * P o (dP - dot), dot = sum_j dP_j P_j; P of this row comes back from HBM as bf16
        const __nv_bfloat16* pr = p.p_in + (1LL * bh * N + (row_ok ? q : 0)) * N;
        float dot = 0.f;
        for (int c0 = 0; c0 < N; c0 += 64) {
            uint32_t v[4][16];
            uint4 pv[4][2];
            const int nsub = min(4, (N - c0) >> 4);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) {
                    tmem_ld_x16(taddr + c0 + 16 * s, v[s]);
                    pv[s][0] = *reinterpret_cast<const uint4*>(pr + c0 + 16 * s);
                    pv[s][1] = *reinterpret_cast<const uint4*>(pr + c0 + 16 * s + 8);
                }
            tmem_ld_wait();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) {
                    const uint32_t w[8] = {pv[s][0].x, pv[s][0].y, pv[s][0].z, pv[s][0].w,
                                           pv[s][1].x, pv[s][1].y, pv[s][1].z, pv[s][1].w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        dot = fmaf(__uint_as_float(v[s][2 * j]), __uint_as_float(w[j] << 16), dot);
                        dot = fmaf(__uint_as_float(v[s][2 * j + 1]), __uint_as_float(w[j] & 0xFFFF0000u), dot);
                    }
                }
        }
        uint8_t* sp = smem + AF_SMEM_P + mt * 4 * AF_TILE;
        __nv_bfloat16* pg = row_ok ? p.p_out + (1LL * bh * N + q) * N : nullptr;
        for (int c0 = 0; c0 < N; c0 += 64) {
            uint32_t v[4][16];
            uint4 pv[4][2];
            const int nsub = min(4, (N - c0) >> 4);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) {
                    tmem_ld_x16(taddr + c0 + 16 * s, v[s]);
                    pv[s][0] = *reinterpret_cast<const uint4*>(pr + c0 + 16 * s);
                    pv[s][1] = *reinterpret_cast<const uint4*>(pr + c0 + 16 * s + 8);
                }
            tmem_ld_wait();
            uint8_t* chunk = sp + (c0 >> 6) * AF_TILE + row * 128;
#pragma unroll
            for (int s = 0; s < 4; ++s)
                if (s < nsub) {
                    const uint32_t pw[8] = {pv[s][0].x, pv[s][0].y, pv[s][0].z, pv[s][0].w,
                                            pv[s][1].x, pv[s][1].y, pv[s][1].z, pv[s][1].w};
                    uint32_t w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float d0 = p.scale * __uint_as_float(pw[j] << 16) * (__uint_as_float(v[s][2 * j]) - dot);
                        const float d1 =
                            p.scale * __uint_as_float(pw[j] & 0xFFFF0000u) * (__uint_as_float(v[s][2 * j + 1]) - dot);
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(d0, d1);
                        w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    *reinterpret_cast<uint4*>(chunk + (((2 * s) ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                    *reinterpret_cast<uint4*>(chunk + (((2 * s + 1) ^ (row & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                    if (pg != nullptr) {
                        *reinterpret_cast<uint4*>(pg + c0 + 16 * s) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(pg + c0 + 16 * s + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
        }
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (threadIdx.x == 0) {
        // ---- O[mt] = P[mt] V : A = P (K-major over keys), B = V (MN-major: d contiguous), overwrites S[mt]'s columns
        const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);
        const uint32_t spu = smem_u32(smem + AF_SMEM_P), sv = smem_u32(smem + AF_SMEM_V);
        const int ksteps = N >> 4;
        for (int t = 0; t < m_tiles; ++t) {
            for (int k = 0; k < ksteps; ++k) {
                const uint32_t a_addr = spu + t * 4 * AF_TILE + (k >> 2) * AF_TILE + (k & 3) * 32u;
                const uint32_t b_addr = sv + k * 2048u;
                umma_bf16(tmem_base + t * 256, make_smem_desc(a_addr, 16u, 1024), make_smem_desc(b_addr, 8192u, 1024),
                          idesc, k > 0 ? 1u : 0u);
            }
            umma_commit(&bar_o[t]);
        }
    }

    if (tile_ok) {
        mbar_wait(&bar_o[mt], 0, 13);
        tc_fence_after();
        const uint32_t taddr = tmem_base + mt * 256 + (static_cast<uint32_t>(quad * 32) << 16);
        uint32_t v[4][16];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 4; ++s) tmem_ld_x16(taddr + 16 * s, v[s]);
        tmem_ld_wait();
        if (row_ok) {
            __nv_bfloat16* op = p.out + (1LL * b * N + q) * p.ld_out + h * 64;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __nv_bfloat162 b2 =
                        __floats2bfloat162_rn(__uint_as_float(v[s][2 * j]), __uint_as_float(v[s][2 * j + 1]));
                    w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                }
                *reinterpret_cast<uint4*>(op + 16 * s) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(op + 16 * s + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace adm

using namespace adm;

#include <cudaTypedefs.h>

static PFN_cuTensorMapEncodeTiled af_encode = nullptr;

static int af_map(CUtensorMap* m, const void* ptr, long long c3, int n_pix, int batch, int box_rows) {
    if (af_encode == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || fn == nullptr) {
            set_error("cuTensorMapEncodeTiled unavailable");
            return ADM_ERR_CUDA;
        }
        af_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    }
    cuuint64_t gd[3] = {static_cast<cuuint64_t>(c3), static_cast<cuuint64_t>(n_pix), static_cast<cuuint64_t>(batch)};
    cuuint64_t gs[2] = {static_cast<cuuint64_t>(c3) * 2, static_cast<cuuint64_t>(c3) * n_pix * 2};
    cuuint32_t bx[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = af_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gd, gs, bx, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("attn_fwd_fused: cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
        return ADM_ERR_CUDA;
    }
    return 0;
}

template <int MODE>
static int af_launch(const CUtensorMap& ma, const CUtensorMap& mb, const AttnParams& p, int blocks, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(attn_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, AF_SMEM_TOTAL);
        attr_set = true;
    }
    attn_fused_kernel<MODE><<<blocks, AF_THREADS, AF_SMEM_TOTAL, st>>>(ma, mb, p);
    ADM_CHECK_LAUNCH("attn_fused");
    return 0;
}

static int af_check(const char* what, int batch, int n_pix, int heads, const void* a, const void* b, const void* c) {
    if (n_pix != 16 && n_pix != 64 && n_pix != 256) {
        set_error("%s: n_pix must be 16, 64 or 256 (got %d)", what, n_pix);
        return ADM_ERR_SHAPE;
    }
    if (batch <= 0 || heads <= 0 || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                                       reinterpret_cast<uintptr_t>(c)) & 15)) {
        set_error("%s: bad arguments", what);
        return ADM_ERR_SHAPE;
    }
    return 0;
}

extern "C" int adm_attn_fwd_fused(const void* qkv, int batch, int n_pix, int heads, float scale, void* out, void* p_out,
                                  void* stream) {
    if (int e = af_check("attn_fwd_fused", batch, n_pix, heads, qkv, out, p_out)) return e;
    const int c = heads * 64;
    CUtensorMap mq, mkv;
    if (int e = af_map(&mq, qkv, 3LL * c, n_pix, batch, n_pix > 128 ? 256 : 128)) return e;
    if (int e = af_map(&mkv, qkv, 3LL * c, n_pix, batch, n_pix)) return e;
    AttnParams p;
    p.n_pix = n_pix; p.heads = heads; p.c = c; p.scale = scale;
    p.out = static_cast<__nv_bfloat16*>(out); p.ld_out = c;
    p.p_out = static_cast<__nv_bfloat16*>(p_out); p.p_in = nullptr;
    p.a_col = 0; p.b1_col = c; p.b2_col = 2 * c;
    return af_launch<0>(mq, mkv, p, batch * heads, static_cast<cudaStream_t>(stream));
}

extern "C" int adm_attn_bwd_fused(const void* da, const void* qkv, const void* p_saved, int batch, int n_pix, int heads,
                                  float scale, void* ds_out, void* dqkv, void* stream) {
    if (int e = af_check("attn_bwd_fused", batch, n_pix, heads, da, qkv, dqkv)) return e;
    if (p_saved == nullptr || ds_out == nullptr || ((reinterpret_cast<uintptr_t>(p_saved) |
                                                       reinterpret_cast<uintptr_t>(ds_out)) & 15)) {
        set_error("attn_bwd_fused: P and dS buffers are required (16 B aligned)");
        return ADM_ERR_SHAPE;
    }
    const int c = heads * 64;
    CUtensorMap mdo, mkv;
    if (int e = af_map(&mdo, da, c, n_pix, batch, n_pix > 128 ? 256 : 128)) return e;
    if (int e = af_map(&mkv, qkv, 3LL * c, n_pix, batch, n_pix)) return e;
    AttnParams p;
    p.n_pix = n_pix; p.heads = heads; p.c = c; p.scale = scale;
    p.out = static_cast<__nv_bfloat16*>(dqkv); p.ld_out = 3LL * c;  // dQ occupies channels [0, C) of dqkv
    p.p_out = static_cast<__nv_bfloat16*>(ds_out); p.p_in = static_cast<const __nv_bfloat16*>(p_saved);
    p.a_col = 0; p.b1_col = 2 * c; p.b2_col = c;                    // A = dO, B1 = V, B2 = K
    return af_launch<1>(mdo, mkv, p, batch * heads, static_cast<cudaStream_t>(stream));
}
