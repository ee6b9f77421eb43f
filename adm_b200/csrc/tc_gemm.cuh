// tcgen05 / TMEM / TMA GEMM engine shared by conv3x3 (implicit GEMM fprop + dgrad), conv1x1, wgrad, Linear and the
// batched attention products.  One persistent warp-specialised kernel:
//   warp 0      : TMA producer (one lane)        global -> 128B-swizzled smem ring
//   warp 1      : TMEM allocator + UMMA issuer   tcgen05.mma, fp32 accumulators in TMEM (2 x 256 columns)
//   warps 2..5  : epilogue                       tcgen05.ld -> +bias, +residual, *alpha -> bf16 / fp32 / fp32 atomics
// Tile = 128 (M) x BN (N, runtime, multiple of 16, <= 256) x 64 (K per stage, bf16 = one 128 B swizzle row).
#pragma once
#include "ptx.cuh"

namespace adm {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_A_STAGE = GEMM_BLOCK_M * 128;  // 16 KB
constexpr int GEMM_SMEM_RING = 200 * 1024;        // operand ring budget
constexpr int GEMM_SMEM_AUX = 1024;               // barriers + tmem ptr
constexpr int GEMM_SMEM_TOTAL = GEMM_SMEM_RING + GEMM_SMEM_AUX + 1024;  // + alignment slack
constexpr int GEMM_MAX_STAGES = 8;

enum GemmMode { GEMM_PLAIN = 0, GEMM_CONV = 1, GEMM_WGRAD = 2 };
enum GemmOut { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ATOMIC = 2 };

struct GemmParams {
    // ---- tile space: tile = (((bt * splits + sp) * m_tiles + mt) * n_tiles + nt)
    int m_tiles, n_tiles, batches, splits;
    int k_iters;  // K iterations (of 64) per split
    int k_total;  // total K iterations (last split may be shorter)
    int bn;       // N tile
    int a_mn, b_mn;
    int M, N;  // valid rows (per batch) / cols
    // ---- conv / wgrad geometry (NHWC tensors; box = bw x bh x bni pixels)
    int H, W, bw, bh, bni, tiles_w, tiles_h;
    int ntaps;     // 1 or 9
    int cchunks;   // 64-channel chunks per tap (both A sources)
    int cchunks1;  // chunks taken from A source 1; the rest come from A source 2 (fused channel concat)
    int n_split;   // wgrad: output columns (per tap) served by X source 1; the rest by X source 2
    // ---- plain / batched coordinates: bt -> (b_hi = bt / bdiv, b_lo = bt % bdiv)
    int bdiv;
    int a_c0, a_c0_lo, a_c1, a_c1_lo, a_bhi, a_blo;
    int b_c0, b_c0_lo, b_c1, b_c1_lo, b_bhi, b_blo;
    // ---- epilogue
    void* C;
    long long ldc, c_bhi, c_blo;  // element strides
    int c_col_lo;                 // extra column offset per b_lo
    const float* bias;            // [N] or null
    const __nv_bfloat16* residual;
    long long ldr;
    float alpha;
    int out_mode;
};

__device__ __forceinline__ void decode_pix(const GemmParams& p, int idx, int& n0, int& h0, int& w0) {
    const int per = p.tiles_w * p.tiles_h;
    n0 = (idx / per) * p.bni;
    const int r = idx % per;
    h0 = (r / p.tiles_w) * p.bh;
    w0 = (r % p.tiles_w) * p.bw;
}

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = GEMM_A_STAGE + p.bn * 128;
    int num_stages = GEMM_SMEM_RING / stage_bytes;
    if (num_stages > GEMM_MAX_STAGES) num_stages = GEMM_MAX_STAGES;

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_SMEM_RING);
    uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < num_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 4);
        }
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_tiles = p.batches * p.splits * p.m_tiles * p.n_tiles;

    if (warp == 0) {
        // =========================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                int r = tile / p.n_tiles;
                const int mt = r % p.m_tiles;
                r /= p.m_tiles;
                const int sp = r % p.splits;
                const int bt = r / p.splits;
                const int k_begin = sp * p.k_iters;
                const int k_end = min(p.k_total, k_begin + p.k_iters);
                int n0 = 0, h0 = 0, w0 = 0;
                if (MODE == GEMM_CONV) decode_pix(p, mt, n0, h0, w0);
                const int b_hi = bt / p.bdiv, b_lo = bt % p.bdiv;
                for (int ki = k_begin; ki < k_end; ++ki) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem + stage * stage_bytes;
                    uint8_t* sb = sa + GEMM_A_STAGE;
                    mbar_expect_tx(&full_bar[stage], stage_bytes);
                    if (MODE == GEMM_CONV) {
                        const int tap = ki / p.cchunks, kc = ki % p.cchunks;
                        int dh = 0, dw = 0;
                        if (p.ntaps == 9) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
                        if (kc < p.cchunks1)
                            tma_load_4d(sa, &tmA, &full_bar[stage], kc * 64, w0 + dw, h0 + dh, n0);
                        else
                            tma_load_4d(sa, &tmA2, &full_bar[stage], (kc - p.cchunks1) * 64, w0 + dw, h0 + dh, n0);
                        if (!p.b_mn) {
                            // weights [Nout][ntaps * Kpad], K-major
                            tma_load_2d(sb, &tmB, &full_bar[stage], ki * 64, nt * p.bn);
                        } else {
                            // dgrad on the fprop-packed weights viewed as (N_gemm = Cin, taps, K_gemm = Cout): the
                            // spatial flip is the tap reversal, the transpose is the MN-major operand mode.
                            for (int c = 0; c * 64 < p.bn; ++c)
                                tma_load_3d(sb + c * 8192, &tmB, &full_bar[stage], nt * p.bn + c * 64,
                                            p.ntaps - 1 - tap, kc * 64);
                        }
                    } else if (MODE == GEMM_WGRAD) {
                        // K runs over pixel blocks of 64; bt = tap.  A = dY (MN-major), B = shifted X (MN-major).
                        decode_pix(p, ki, n0, h0, w0);
                        int dh = 0, dw = 0;
                        if (p.ntaps == 9) { dh = bt / 3 - 1; dw = bt % 3 - 1; }
                        tma_load_4d(sa, &tmA, &full_bar[stage], mt * 128, w0, h0, n0);
                        tma_load_4d(sa + 8192, &tmA, &full_bar[stage], mt * 128 + 64, w0, h0, n0);
                        // fused channel concat on the X side: columns >= n_split come from the second source
                        const int ncol = nt * p.bn;
                        const bool src2 = ncol >= p.n_split;
                        const CUtensorMap* mb = src2 ? &tmA2 : &tmB;
                        const int nc = src2 ? ncol - p.n_split : ncol;
                        for (int c = 0; c * 64 < p.bn; ++c)
                            tma_load_4d(sb + c * 8192, mb, &full_bar[stage], nc + c * 64, w0 + dw, h0 + dh, n0);
                    } else {
                        const int a0 = p.a_c0 + b_lo * p.a_c0_lo, a1 = p.a_c1 + b_lo * p.a_c1_lo;
                        const int a2 = b_hi * p.a_bhi + b_lo * p.a_blo;
                        const int b0 = p.b_c0 + b_lo * p.b_c0_lo, b1 = p.b_c1 + b_lo * p.b_c1_lo;
                        const int b2 = b_hi * p.b_bhi + b_lo * p.b_blo;
                        if (!p.a_mn) {
                            tma_load_3d(sa, &tmA, &full_bar[stage], a0 + ki * 64, a1 + mt * 128, a2);
                        } else {
                            tma_load_3d(sa, &tmA, &full_bar[stage], a0 + mt * 128, a1 + ki * 64, a2);
                            tma_load_3d(sa + 8192, &tmA, &full_bar[stage], a0 + mt * 128 + 64, a1 + ki * 64, a2);
                        }
                        if (!p.b_mn) {
                            tma_load_3d(sb, &tmB, &full_bar[stage], b0 + ki * 64, b1 + nt * p.bn, b2);
                        } else {
                            for (int c = 0; c * 64 < p.bn; ++c)
                                tma_load_3d(sb + c * 8192, &tmB, &full_bar[stage], b0 + nt * p.bn + c * 64,
                                            b1 + ki * 64, b2);
                        }
                    }
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== UMMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, p.bn, p.a_mn, p.b_mn);
            const uint32_t a_lbo = p.a_mn ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
            const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int sp = (tile / (p.n_tiles * p.m_tiles)) % p.splits;
                const int k_begin = sp * p.k_iters;
                const int k_end = min(p.k_total, k_begin + p.k_iters);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * 256;
                for (int ki = k_begin; ki < k_end; ++ki) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                    const uint32_t sb = sa + GEMM_A_STAGE;
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
                        const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
                        const uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
                        umma_bf16(tmem_d, da, db, idesc, (ki > k_begin || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // =========================================================== epilogue (4 warps, one TMEM lane quadrant each)
        const int quad = warp & 3;
        const int m = quad * 32 + lane;  // row inside the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles;
            int r = tile / p.n_tiles;
            const int mt = r % p.m_tiles;
            r /= p.m_tiles;
            const int bt = r / p.splits;
            const int b_hi = bt / p.bdiv, b_lo = bt % p.bdiv;

            long long row_off;  // element offset of this thread's output row (before column)
            bool row_ok;
            if (MODE == GEMM_CONV) {
                int n0, h0, w0;
                decode_pix(p, mt, n0, h0, w0);
                const int wi = m % p.bw, hi = (m / p.bw) % p.bh, ni = m / (p.bw * p.bh);
                const long long pix = (static_cast<long long>(n0 + ni) * p.H + (h0 + hi)) * p.W + (w0 + wi);
                row_ok = pix < p.M;
                row_off = pix;
            } else {
                const int row = mt * 128 + m;
                row_ok = row < p.M;
                row_off = row;
            }
            const long long c_base = b_hi * p.c_bhi + b_lo * p.c_blo + row_off * p.ldc;
            const int col_base = nt * p.bn;           // column inside [0, N)
            const int col_shift = b_lo * p.c_col_lo;  // extra offset in the output row

            mbar_wait(&tfull_bar[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(quad * 32) << 16);
            for (int c = 0; c < p.bn; c += 16) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld_x16(taddr + c, v);
                tmem_ld_wait();
                const int col = col_base + c;
                if (col >= p.N) break;  // warp-uniform
                if (row_ok) {
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) * p.alpha;
                const bool full = (col + 16 <= p.N);
                if (p.bias != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (full || col + j < p.N) f[j] += __ldg(p.bias + col + j);
                }
                const long long off = c_base + col_shift + col;
                if (p.residual != nullptr) {
                    const __nv_bfloat16* rp = p.residual + row_off * p.ldr + col;
                    if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
                        const uint4 r0 = *reinterpret_cast<const uint4*>(rp);
                        const uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
                        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
                            f[2 * j] += __bfloat162float(b2.x);
                            f[2 * j + 1] += __bfloat162float(b2.y);
                        }
                    } else {
                        for (int j = 0; j < 16; ++j)
                            if (col + j < p.N) f[j] += __bfloat162float(rp[j]);
                    }
                }
                if (p.out_mode == OUT_BF16) {
                    __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + off;
                    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
                        uint32_t w[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                            w[j] = *reinterpret_cast<const uint32_t*>(&b2);
                        }
                        *reinterpret_cast<uint4*>(cp) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(cp + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                    } else {
                        for (int j = 0; j < 16; ++j)
                            if (col + j < p.N) cp[j] = __float2bfloat16(f[j]);
                    }
                } else if (p.out_mode == OUT_F32) {
                    float* cp = reinterpret_cast<float*>(p.C) + off;
                    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4*>(cp + 4 * j) =
                                make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                    } else {
                        for (int j = 0; j < 16; ++j)
                            if (col + j < p.N) cp[j] = f[j];
                    }
                } else {
                    float* cp = reinterpret_cast<float*>(p.C) + off;
                    if (full && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            atomicAdd(reinterpret_cast<float4*>(cp + 4 * j),
                                      make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                    } else {
                        for (int j = 0; j < 16; ++j)
                            if (col + j < p.N) atomicAdd(cp + j, f[j]);
                    }
                }
                }  // row_ok
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
