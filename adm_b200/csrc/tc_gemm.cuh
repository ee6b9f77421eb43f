// tcgen05 / TMEM / TMA GEMM engine shared by conv3x3 (implicit GEMM fprop + dgrad), conv1x1, wgrad, Linear and the
// batched attention products.  Persistent warp-specialised kernels — tc_gemm_kernel<PLAIN|CONV|WGRAD> (general),
// tc_conv_halo_kernel (3x3 fprop/dgrad with halo re-use of the pixel operand), tc_wgrad_rows_kernel (3x3 wgrad, three
// taps per MMA), tc_conv_pair_kernel (cta_group::2 experiment) — with the same roles:
//   warp 0      : TMA producer (one lane)        global -> 128B-swizzled smem ring
//   warp 1      : TMEM allocator + UMMA issuer   tcgen05.mma, fp32 accumulators in TMEM (2 x 256 columns)
//   warps 2..   : epilogue                       tcgen05.ld -> +bias, +residual, *alpha -> bf16 / fp32 / fp32 atomics
//                 (warp w drains TMEM lane quadrant w % 4; with GEMM_EPI_WARPS = 8 the two warps of a quadrant split
//                 the tile's columns)
// Tile = 128 (M) x BN (N, runtime, multiple of 16, <= 256) x 64 (K per stage, bf16 = one 128 B swizzle row).
#pragma once
#include "ptx.cuh"
#include "adm_internal.h"
#include "vecmath.cuh"
#ifdef ADM_GEMM_TIMING
#include <stdio.h>
#endif

namespace adm {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;
// Epilogue warps: 4 (one per TMEM lane quadrant) or 8 (two per quadrant, each draining half of the tile's columns).
// Measured on B200 (tools/gemm_accounting.py): 8 warps change nothing on the short-K 1x1 convs — they wait for operand
// bytes, not for the epilogue — and cost ~8 % on the small 8x8 / 4x4 tiles, so 4 it is.
constexpr int GEMM_EPI_WARPS = 4;
constexpr int EPI_HALVES = GEMM_EPI_WARPS / 4;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;   // TMA warp + MMA warp + epilogue warps
constexpr int PAIR_THREADS = 192;                        // the (experimental) CTA-pair kernel keeps 4 epilogue warps
constexpr int GEMM_A_STAGE = GEMM_BLOCK_M * 128;  // 16 KB
constexpr int GEMM_SMEM_RING = 200 * 1024;        // operand ring budget
constexpr int GEMM_SMEM_AUX = 4096 + 8192;        // barriers + tmem ptr (first 1 KB) + per-tile bias slices (2 x 1 KB) +
                                                  // 1 KB spare + the prologue's per-tile (A, B) table (1024 channels x 8 B)
// halo conv with the GroupNorm prologue: + 4 transform warps.  (8 warps halve the transform time but cap the kernel at 128
// registers per thread, which spills in the epilogue; measured in the step they are equal within noise — see
// profiles/r03_conv_gn_prologue.txt for all the variants that were timed.)
constexpr int GEMM_PRO_WARPS = 4;
constexpr int GEMM_PRO_THREADS = GEMM_THREADS + 32 * GEMM_PRO_WARPS;
constexpr int GEMM_SMEM_TOTAL = GEMM_SMEM_RING + GEMM_SMEM_AUX + 1024;  // + alignment slack
constexpr int GEMM_MAX_STAGES = 8;

enum GemmMode { GEMM_PLAIN = 0, GEMM_CONV = 1, GEMM_WGRAD = 2 };
enum GemmOut { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_ATOMIC = 2 };

struct GemmParams {
    // ---- tile space: tile = (((bt * splits + sp) * m_tiles + mt) * n_tiles + nt)
    int m_tiles, n_tiles, batches, splits;
    int k_iters;  // K iterations (of 64) per split
    int k_total;  // total K iterations (last split may be shorter)
    int ksplit;   // tc_conv_splitk_kernel: CTAs of a cluster that split the K range of one tile (reduced through DSMEM)
    int bn;       // N tile
    int a_mn, b_mn;
    int M, N;  // valid rows (per batch) / cols
    // ---- conv / wgrad geometry (NHWC tensors; box = bw x bh x bni pixels)
    int H, W, bw, bh, bni, tiles_w, tiles_h;
    int ntaps;     // kw * kw: 1, 9, 25 or 49 (odd square kernels, 'same' padding kw / 2)
    int kw;        // taps per kernel row (1, 3, 5, 7)
    int cchunks;   // 64-channel chunks per tap (both A sources)
    int cchunks1;  // chunks taken from A source 1; the rest come from A source 2 (fused channel concat)
    int n_split;   // wgrad: output columns (per tap) served by X source 1; the rest by X source 2
    int co_chunks; // tap-row wgrad: 64-channel chunks of dY (M chunks = 3 dy x co_chunks)
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
    const int* row_map;  // plain / wgrad tiles: GEMM row r is stored at output row row_map[r] (nullptr = identity)
    int debug;  // experiments only (ADM_GEMM_DEBUG): bit 0 = skip the MMAs, bit 1 = skip the TMA loads
    // ---- GroupNorm statistics of the OUTPUT, emitted by the conv epilogue: per (sample, slot, channel)
    // partial {sum, sum of squares} over the pixels one epilogue warp (or half-warp, 4x4 images) holds, plain stores into
    // stats[n][slot][N][2] — no atomics, no zeroing, deterministic.  The next norm derives its group statistics from these
    // instead of re-reading the activation (adm_gn_finalize).  nullptr = off.
    float* stats;
    int stats_slots;  // slots per sample: (H*W) / 32, at least 1
    // ---- GroupNorm(+ adaptive scale/shift) + SiLU (+ dropout) PROLOGUE of the halo conv kernel (north_star: "implicit-GEMM
    // conv3x3 ... with a GroupNorm+SiLU prologue"): the conv reads the RAW tensor; four transform warps apply
    // y = dropout(silu(x * A[n,c] + B[n,c])) to each halo tile in shared memory, once per 64-channel chunk (not once per
    // tap), between the TMA load and the MMAs.  pro_coef = the norm's coefficient table [n][C]{A, B, mean, rstd}.
    const float4* pro_coef;
    int pro_act, pro_c1, pro_c2;          // SiLU on/off; valid channels of A source 1 / 2 (C = c1 + c2)
    float pro_drop_p;
    unsigned long long pro_seed;
    const unsigned long long* pro_seed_dev;
    __nv_bfloat16* pro_out;               // optional: the activated tensor [N,H,W,C] (the wgrad operand), written once
    long long pro_ldo;
    const __nv_bfloat16* pro_x1;          // the raw A sources (the prologue warps read them with plain vector loads)
    const __nv_bfloat16* pro_x2;
    long long pro_ld1, pro_ld2;
};

// sm_100a SFU tanh; silu(v) = h * tanh(h) + h with h = v / 2
__device__ __forceinline__ float gemm_tanh_fast(float h) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return t;
}
// The dropout keep-scales are those of csrc/norm.cu (vecmath.cuh: same hash, same indexing — backward regenerates the
// masks there).

__device__ __forceinline__ void decode_pix(const GemmParams& p, int idx, int& n0, int& h0, int& w0) {
    const int per = p.tiles_w * p.tiles_h;
    n0 = (idx / per) * p.bni;
    const int r = idx % per;
    h0 = (r / p.tiles_w) * p.bh;
    w0 = (r % p.tiles_w) * p.bw;
}

// Pixel-box coordinates of consecutive k-iterations without per-iteration divisions (the producer is ONE thread: a few
// runtime integer divisions per k-iteration are a measurable share of its budget).
struct PixWalker {
    int n0, h0, w0;
    __device__ __forceinline__ PixWalker(const GemmParams& p, int idx) { decode_pix(p, idx, n0, h0, w0); }
    __device__ __forceinline__ void advance(const GemmParams& p) {
        w0 += p.bw;
        if (w0 >= p.W) {
            w0 = 0;
            h0 += p.bh;
            if (h0 >= p.H) { h0 = 0; n0 += p.bni; }
        }
    }
};

// The epilogue warps stage the tile's bias slice (bn <= 256 floats, zero beyond N) in shared memory BEFORE waiting for the
// accumulator: per-chunk __ldg of the bias missed L1 behind the streaming stores and cost ~30 us on the epilogue-bound
// 1x1 convs.  `t` = epilogue thread index; named barrier 1 covers exactly the epilogue threads.
__device__ __forceinline__ void stage_bias(const GemmParams& p, float* sbias, int col_base, int t) {
    if (p.bias == nullptr) return;
    for (int c = t; c < p.bn; c += 32 * GEMM_EPI_WARPS)
        sbias[c] = (col_base + c < p.N) ? __ldg(p.bias + col_base + c) : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory");
}
// pair kernel: 128 epilogue threads, two columns each
__device__ __forceinline__ void stage_bias_128(const GemmParams& p, float* sbias, int col_base, int m) {
    if (p.bias == nullptr) return;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = m + 128 * i;
        if (c < p.bn) sbias[c] = (col_base + c < p.N) ? __ldg(p.bias + col_base + c) : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Column sums of 16 values per lane over the lanes of a warp in 15 (+1) shuffles: at each halving step a lane keeps the
// half of its values selected by one lane bit and receives the partner's copy of that half.  Afterwards lane l holds in
// v[0] the total of column (l & 15) over the 16 lanes sharing its bit 4 — or, with whole_warp, over all 32 lanes.
__device__ __forceinline__ void warp_colsum16(float (&v)[16], int lane, bool whole_warp) {
#pragma unroll
    for (int step = 0; step < 4; ++step) {
        const int off = 8 >> step;
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < (8 >> step); ++i) {
            const float send = hi ? v[i] : v[i + off];
            const float keep = hi ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    if (whole_warp) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// Epilogue of one accumulator tile for one thread (= one output row): TMEM -> registers in groups of up to 64 columns
// (four 32x32b.x16 loads in flight behind ONE wait, with the residual row segment prefetched behind the same wait),
// then alpha * acc + bias + residual -> bf16 / fp32 / fp32 atomics.  c_off / r_off: element offsets of this row in the
// output and the residual (before the column); col_shift: extra output column offset (batched GEMMs); the tile columns
// drained are [c_begin, ncols) (multiples of 16).
// stat_row: this lane's base into the statistics table (sample and slot resolved, column 0), or nullptr; the column sums
// are taken over the bf16-ROUNDED outputs (what the next norm will read); rows that do not exist contribute zeros.
__device__ __forceinline__ void epilogue_row(const GemmParams& p, uint32_t taddr, int col_base, int col_shift,
                                             long long c_off, long long r_off, bool row_ok, const float* sbias,
                                             int c_begin, int ncols, float* stat_row = nullptr, bool stat_whole = true,
                                             int lane = 0) {
    for (int c0 = c_begin; c0 < ncols; c0 += 64) {
        if (col_base + c0 >= p.N) break;  // warp-uniform
        uint32_t v[4][16];
        uint4 rr[4][2];
        const int nsub = min(4, (ncols - c0) >> 4);
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s < nsub) tmem_ld_x16(taddr + c0 + 16 * s, v[s]);
        // residual prefetch (aligned full chunks only; the ragged path re-reads below)
        const bool res_vec = p.residual != nullptr && row_ok &&
                             ((reinterpret_cast<uintptr_t>(p.residual + r_off + col_base + c0) & 15) == 0);
        if (res_vec) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int col = col_base + c0 + 16 * s;
                if (s < nsub && col + 16 <= p.N) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + r_off + col);
                    rr[s][0] = rp[0];
                    rr[s][1] = rp[1];
                }
            }
        }
        tmem_ld_wait();
        if (!row_ok && stat_row == nullptr) continue;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s >= nsub) break;
            const int col = col_base + c0 + 16 * s;
            if (col >= p.N) break;
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[s][j]) * p.alpha;
            const bool full = (col + 16 <= p.N);
            if (p.bias != nullptr) {  // this tile's bias slice was staged in shared memory (zero beyond N)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + 16 * s + 4 * j);
                    f[4 * j] += b4.x; f[4 * j + 1] += b4.y; f[4 * j + 2] += b4.z; f[4 * j + 3] += b4.w;
                }
            }
            if (p.residual != nullptr) {
                if (res_vec && full) {
                    const uint32_t w[8] = {rr[s][0].x, rr[s][0].y, rr[s][0].z, rr[s][0].w,
                                           rr[s][1].x, rr[s][1].y, rr[s][1].z, rr[s][1].w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        f[2 * j] += __uint_as_float(w[j] << 16);
                        f[2 * j + 1] += __uint_as_float(w[j] & 0xFFFF0000u);
                    }
                } else if (row_ok) {
                    const __nv_bfloat16* rp = p.residual + r_off + col;
                    for (int j = 0; j < 16; ++j)
                        if (col + j < p.N) f[j] += __bfloat162float(rp[j]);
                }
            }
            const long long off = c_off + col_shift + col;
            if (stat_row != nullptr) {  // every lane of the warp takes part (rows that do not exist contribute zeros)
                float s1[16], s2[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float r = (row_ok && col + j < p.N) ? __bfloat162float(__float2bfloat16(f[j])) : 0.f;
                    s1[j] = r;
                    s2[j] = r * r;
                }
                warp_colsum16(s1, lane, stat_whole);
                warp_colsum16(s2, lane, stat_whole);
                if ((stat_whole ? lane < 16 : true) && col + (lane & 15) < p.N)
                    *reinterpret_cast<float2*>(stat_row + 2 * (col + (lane & 15))) = make_float2(s1[0], s2[0]);
                if (!row_ok) continue;
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
                    if ((reinterpret_cast<uintptr_t>(cp) & 31) == 0) {
                        // one 256-bit store = one full 32 B sector per thread (sm_100 st.global.v8): rows of a tile are
                        // scattered over 128 different lines, so two 16 B halves would each be a partial-sector write
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(cp), "r"(w[0]),
                                     "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                     : "memory");
                    } else {
                        *reinterpret_cast<uint4*>(cp) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(cp + 8) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
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
        }
    }
}

// The epilogue warps' persistent loop (4 warps, one TMEM lane quadrant each): per tile, stage the bias slice, wait for
// the accumulator, drain it through epilogue_row, release it.  Shared by tc_gemm_kernel and tc_conv_halo_kernel.
template <int MODE>
__device__ __forceinline__ void epilogue_warps(const GemmParams& p, uint8_t* smem, uint32_t tmem_base,
                                               uint64_t* tfull_bar, uint64_t* tempty_bar, int warp, int lane,
                                               int num_tiles) {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;   // which half of the tile's columns this warp drains
    const int m = quad * 32 + lane;     // row inside the tile
    const int c_split = EPI_HALVES == 2 ? (((p.bn >> 1) + 15) & ~15) : p.bn;
    const int c_begin = half ? c_split : 0, c_end = half ? p.bn : c_split;
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
        float* stat_row = nullptr;
        bool stat_whole = true;
        if (MODE == GEMM_CONV) {
            int n0, h0, w0;
            decode_pix(p, mt, n0, h0, w0);
            const int wi = m % p.bw, hi = (m / p.bw) % p.bh, ni = m / (p.bw * p.bh);
            const long long pix = (static_cast<long long>(n0 + ni) * p.H + (h0 + hi)) * p.W + (w0 + wi);
            row_ok = pix < p.M;
            row_off = pix;
            if (p.stats != nullptr) {
                // slot of this warp's 32 rows inside its sample: tiles of one image are (tile in image) * 4 + quadrant;
                // images smaller than a tile (8x8: 2 warps each, 4x4: half a warp each) count their own warps
                const int rows_per_img = p.bw * p.bh;  // rows of one image inside this tile
                int slot;
                if (p.bni == 1) {
                    slot = (mt % (p.tiles_w * p.tiles_h)) * 4 + quad;
                } else {
                    slot = (m % rows_per_img) >> 5;
                    stat_whole = rows_per_img >= 32;
                }
                // samples beyond the batch (ragged last tile of packed small images) write into one scratch sample row
                const long long nsmp = p.M / (p.H * p.W);
                const long long smp = (n0 + ni) < nsmp ? (n0 + ni) : nsmp;
                stat_row = p.stats + (smp * p.stats_slots + slot) * 2LL * p.N;
            }
        } else {
            const int row = mt * 128 + m;
            row_ok = row < p.M;
            row_off = (p.row_map != nullptr && row_ok) ? __ldg(p.row_map + row) : row;
        }
        const long long c_base = b_hi * p.c_bhi + b_lo * p.c_blo + row_off * p.ldc;
        const int col_base = nt * p.bn;           // column inside [0, N)
        const int col_shift = b_lo * p.c_col_lo;  // extra offset in the output row

        float* sbias = reinterpret_cast<float*>(smem + GEMM_SMEM_RING + 1024) + acc * 256;
        stage_bias(p, sbias, col_base, (warp - 2) * 32 + lane);
        mbar_wait(&tfull_bar[acc], acc_phase, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(quad * 32) << 16);
        epilogue_row(p, taddr, col_base, col_shift, c_base, row_off * p.ldr, row_ok, sbias, c_begin, c_end, stat_row,
                     stat_whole, lane);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
}

template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
    pdl_trigger();
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
            mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
        }
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();  // everything above touched only this kernel's own state; the predecessor's results are visible from here on

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
                int tap = 0, kc = 0;  // conv: ki = tap * cchunks + kc, walked without divisions
                if (MODE == GEMM_CONV) { tap = k_begin / p.cchunks; kc = k_begin % p.cchunks; }
                PixWalker px(p, MODE == GEMM_WGRAD ? k_begin : 0);
                for (int ki = k_begin; ki < k_end; ++ki) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem + stage * stage_bytes;
                    uint8_t* sb = sa + GEMM_A_STAGE;
                    if (p.debug & 2) {
                        mbar_arrive(&full_bar[stage]);
                        if (++stage == num_stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[stage], stage_bytes);
                    if (MODE == GEMM_CONV) {
                        int dh = 0, dw = 0;
                        if (p.kw == 3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
                        else if (p.kw > 1) { dh = tap / p.kw - (p.kw >> 1); dw = tap % p.kw - (p.kw >> 1); }
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
                        if (++kc == p.cchunks) { kc = 0; ++tap; }
                    } else if (MODE == GEMM_WGRAD) {
                        // K runs over pixel blocks of 64; bt = tap.  A = dY (MN-major), B = shifted X (MN-major).
                        n0 = px.n0; h0 = px.h0; w0 = px.w0;
                        px.advance(p);
                        int dh = 0, dw = 0;
                        if (p.kw == 3) { dh = bt / 3 - 1; dw = bt % 3 - 1; }
                        else if (p.kw > 1) { dh = bt / p.kw - (p.kw >> 1); dw = bt % p.kw - (p.kw >> 1); }
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
            // descriptor constants and k-step strides (in 16 B units) are hoisted: this thread's instruction stream is
            // the critical path of the main loop
            const uint32_t a_kstep = (p.a_mn ? 2048u : 32u) >> 4, b_kstep = (p.b_mn ? 2048u : 32u) >> 4;
            const uint64_t da0 = make_smem_desc(0, a_lbo, 1024), db0 = make_smem_desc(0, b_lbo, 1024);
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
#ifdef ADM_GEMM_TIMING
                long long t_wait = 0, t_issue = 0, t_commit = 0;
#endif
                for (int ki = k_begin; ki < k_end; ++ki) {
#ifdef ADM_GEMM_TIMING
                    const long long c0 = clock64();
#endif
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
#ifdef ADM_GEMM_TIMING
                    const long long c1 = clock64();
#endif
                    const uint32_t sa16 = (smem_u32(smem + stage * stage_bytes) & 0x3FFFFu) >> 4;
                    const uint64_t da = da0 + sa16, db = db0 + (sa16 + (GEMM_A_STAGE >> 4));
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
                        if (!(p.debug & 1))
                            umma_bf16(tmem_d, da + k * a_kstep, db + k * b_kstep, idesc, (ki > k_begin || k > 0) ? 1u : 0u);
#ifdef ADM_GEMM_TIMING
                    const long long c2 = clock64();
#endif
                    umma_commit(&empty_bar[stage]);
#ifdef ADM_GEMM_TIMING
                    const long long c3 = clock64();
                    t_wait += c1 - c0; t_issue += c2 - c1; t_commit += c3 - c2;
#endif
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
#ifdef ADM_GEMM_TIMING
                if (blockIdx.x == 0 && tile == blockIdx.x)
                    printf("[mma thread] k-iters %d: wait %lld issue %lld commit %lld clocks per k-iter\n", k_end - k_begin,
                           t_wait / (k_end - k_begin), t_issue / (k_end - k_begin), t_commit / (k_end - k_begin));
#endif
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // =========================================================== epilogue (4 warps, one TMEM lane quadrant each)
        epilogue_warps<MODE>(p, smem, tmem_base, tfull_bar, tempty_bar, warp, lane, num_tiles);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ cluster split-K conv
// Implicit-GEMM conv (fprop / dgrad, 3x3 or 1x1) for problems with too few 128 x bn output tiles to fill the GPU (the 4x4
// level: 2048 pixels x 384 channels = 48 tiles of 128 x 128, each with a K loop of 54-108 iterations).  A thread-block
// cluster of ksplit CTAs shares ONE output tile: CTA r multiplies K iterations [r * k_iters, (r + 1) * k_iters) into its own
// TMEM accumulator; the non-leaders then park their fp32 partial tiles in their (now idle) operand ring, signal the leader's
// mbarrier across the cluster, and the leader adds them through distributed shared memory (ld.shared::cluster) into its
// accumulator before the ordinary epilogue.  One tile per CTA (grid = tiles x ksplit); roles as in tc_gemm_kernel.
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t caddr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(caddr));
    return v;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {  // release at cluster scope
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int code) {  // acquire at cluster scope
    const long long t0 = clock64();
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!ok && clock64() - t0 > 8000000000LL) {
            g_device_error = code;
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_conv_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                      const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = GEMM_A_STAGE + p.bn * 128;
    int num_stages = GEMM_SMEM_RING / stage_bytes;
    if (num_stages > GEMM_MAX_STAGES) num_stages = GEMM_MAX_STAGES;

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_SMEM_RING);
    uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;
    uint64_t* peers_bar = tfull_bar + 1;  // leader: the non-leaders' partial tiles are parked
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(peers_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const int tile = blockIdx.x / p.ksplit;
    const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
    const int k_begin = rank * p.k_iters;
    const int k_end = min(p.k_total, k_begin + p.k_iters);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < num_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        mbar_init(peers_bar, (p.ksplit - 1) * GEMM_EPI_WARPS);
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 256);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the leader's peers_bar is initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();

    int n0, h0, w0;
    decode_pix(p, mt, n0, h0, w0);
    if (warp == 0) {
        // =========================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int tap = k_begin / p.cchunks, kc = k_begin % p.cchunks;
            for (int ki = k_begin; ki < k_end; ++ki) {
                mbar_wait(&empty_bar[stage], phase ^ 1, 41);
                uint8_t* sa = smem + stage * stage_bytes;
                uint8_t* sb = sa + GEMM_A_STAGE;
                mbar_expect_tx(&full_bar[stage], stage_bytes);
                int dh = 0, dw = 0;
                if (p.kw == 3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
                        else if (p.kw > 1) { dh = tap / p.kw - (p.kw >> 1); dw = tap % p.kw - (p.kw >> 1); }
                if (kc < p.cchunks1)
                    tma_load_4d(sa, &tmA, &full_bar[stage], kc * 64, w0 + dw, h0 + dh, n0);
                else
                    tma_load_4d(sa, &tmA2, &full_bar[stage], (kc - p.cchunks1) * 64, w0 + dw, h0 + dh, n0);
                if (!p.b_mn) {
                    tma_load_2d(sb, &tmB, &full_bar[stage], ki * 64, nt * p.bn);
                } else {
                    for (int c = 0; c * 64 < p.bn; ++c)
                        tma_load_3d(sb + c * 8192, &tmB, &full_bar[stage], nt * p.bn + c * 64, p.ntaps - 1 - tap, kc * 64);
                }
                if (++kc == p.cchunks) { kc = 0; ++tap; }
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // =========================================================== UMMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, p.bn, 0, p.b_mn);
            const uint32_t b_lbo = p.b_mn ? 8192u : 16u;
            const uint32_t b_kstep = (p.b_mn ? 2048u : 32u) >> 4;
            const uint64_t da0 = make_smem_desc(0, 16u, 1024), db0 = make_smem_desc(0, b_lbo, 1024);
            int stage = 0;
            uint32_t phase = 0;
            for (int ki = k_begin; ki < k_end; ++ki) {
                mbar_wait(&full_bar[stage], phase, 43);
                tc_fence_after();
                // In a non-leader CTA of a cluster the shared-window address carries bits above the 14-bit descriptor field:
                // unmasked they spill into the leading-byte-offset field (which only MN-major multi-chunk B reads).
                const uint32_t sa16 = (smem_u32(smem + stage * stage_bytes) & 0x3FFFFu) >> 4;
                const uint64_t da = da0 + sa16, db = db0 + (sa16 + (GEMM_A_STAGE >> 4));
#pragma unroll
                for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
                    umma_bf16(tmem_base, da + 2 * k, db + k * b_kstep, idesc, (ki > k_begin || k > 0) ? 1u : 0u);
                umma_commit(&empty_bar[stage]);
                if (++stage == num_stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tfull_bar);
        }
    } else {
        // =========================================================== epilogue (4 warps, one TMEM lane quadrant each)
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        const int wi = m % p.bw, hi = (m / p.bw) % p.bh, ni = m / (p.bw * p.bh);
        const long long pix = (static_cast<long long>(n0 + ni) * p.H + (h0 + hi)) * p.W + (w0 + wi);
        const bool row_ok = pix < p.M;
        const int col_base = nt * p.bn;
        float* sbias = reinterpret_cast<float*>(smem + GEMM_SMEM_RING + 1024);
        if (rank == 0) stage_bias(p, sbias, col_base, (warp - 2) * 32 + lane);
        mbar_wait(tfull_bar, 0, 44);  // this CTA's K range is in its accumulator; its operand ring is idle from here on
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        float4* park = reinterpret_cast<float4*>(smem);  // [bn / 4][128 rows] float4: lanes write consecutive 16 B
        if (rank != 0) {
            for (int c0 = 0; c0 < p.bn; c0 += 16) {
                uint32_t v[16];
                __syncwarp();
                tmem_ld_x16(taddr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    park[((c0 >> 2) + j) * 128 + m] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                   __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(peers_bar), 0));
        } else {
            if (p.ksplit > 1) {
                mbar_wait_cluster(peers_bar, 0, 45);
                for (int c0 = 0; c0 < p.bn; c0 += 16) {
                    uint32_t v[16];
                    __syncwarp();
                    tmem_ld_x16(taddr + c0, v);
                    tmem_ld_wait();
                    for (int r = 1; r < p.ksplit; ++r) {
                        const uint32_t peer = mapa_shared(smem_u32(park), static_cast<uint32_t>(r));
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 t = ld_dsmem_f4(peer + (((c0 >> 2) + j) * 128 + m) * 16);
                            v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + t.x);
                            v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + t.y);
                            v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + t.z);
                            v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + t.w);
                        }
                    }
                    __syncwarp();
                    tmem_st_x16(taddr + c0, v);  // the total goes back into the accumulator: the epilogue below is unchanged
                }
                tmem_st_wait();
            }
            epilogue_row(p, taddr, col_base, 0, pix * p.ldc, pix * p.ldr, row_ok, sbias, 0, p.bn);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // no CTA exits (or frees TMEM) while the leader may still read its shared memory
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// ------------------------------------------------------------------------------------------------ CTA-pair conv kernel
// Implicit-GEMM conv (K-major packed weights) with cta_group::2: a cluster of two CTAs computes a 256-pixel x BN tile.
// Each CTA TMA-loads its own 128-pixel A box and HALF of the weight rows (BN/2 x 64), so per MMA k-step an SM pulls
// 16 KB + BN*64 B through the L2 fabric instead of 16 KB + BN*128 B — the single-CTA kernel is bound by exactly that
// traffic (profiles/r01_conv_ncu_full.txt).  Roles per CTA are those of tc_gemm_kernel; differences:
//   * full barriers live in the leader (rank 0) and collect the TMA bytes of both CTAs;
//   * the leader's MMA thread issues tcgen05.mma.cta_group::2 (M = 256) and its commits arrive, multicast, on the
//     empty / tmem-full barriers of both CTAs;
//   * both epilogues (4 warps each) release the accumulator on the leader's tmem-empty barrier (count 8).
__global__ void __launch_bounds__(PAIR_THREADS, 1)
tc_conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int half_bn = p.bn >> 1;
    const int stage_bytes = GEMM_A_STAGE + half_bn * 128;
    int num_stages = GEMM_SMEM_RING / stage_bytes;
    if (num_stages > GEMM_MAX_STAGES) num_stages = GEMM_MAX_STAGES;

    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_SMEM_RING);
    uint64_t* empty_bar = full_bar + GEMM_MAX_STAGES;
    uint64_t* tfull_bar = empty_bar + GEMM_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());

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
            mbar_init(&tempty_bar[i], 8);
        }
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc_pair(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // peer barriers are initialised before any remote arrive / TMA credit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int m_pairs = (p.m_tiles + 1) >> 1;
    const int num_tiles = m_pairs * p.n_tiles;
    const int first = blockIdx.x >> 1, stride = gridDim.x >> 1;

    if (warp == 0) {
        // =========================================================== TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = first; tile < num_tiles; tile += stride) {
                const int nt = tile % p.n_tiles;
                const int mt = 2 * (tile / p.n_tiles) + rank;
                int n0, h0, w0;
                decode_pix(p, mt, n0, h0, w0);  // mt == m_tiles (odd tail): n0 is out of range -> zero-filled boxes
                for (int ki = 0; ki < p.k_total; ++ki) {
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem + stage * stage_bytes;
                    uint8_t* sb = sa + GEMM_A_STAGE;
                    if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);
                    const int tap = ki / p.cchunks, kc = ki % p.cchunks;
                    int dh = 0, dw = 0;
                    if (p.kw == 3) { dh = tap / 3 - 1; dw = tap % 3 - 1; }
                        else if (p.kw > 1) { dh = tap / p.kw - (p.kw >> 1); dw = tap % p.kw - (p.kw >> 1); }
                    if (kc < p.cchunks1)
                        tma_load_4d_pair(sa, &tmA, &full_bar[stage], kc * 64, w0 + dw, h0 + dh, n0);
                    else
                        tma_load_4d_pair(sa, &tmA2, &full_bar[stage], (kc - p.cchunks1) * 64, w0 + dw, h0 + dh, n0);
                    tma_load_2d_pair(sb, &tmB, &full_bar[stage], ki * 64, nt * p.bn + rank * half_bn);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== UMMA issuer (leader only)
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = make_idesc_bf16(2 * GEMM_BLOCK_M, p.bn, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = first; tile < num_tiles; tile += stride) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * 256;
#ifdef ADM_GEMM_TIMING
                long long t_wait = 0, t_issue = 0, t_commit = 0;
#endif
                for (int ki = 0; ki < p.k_total; ++ki) {
#ifdef ADM_GEMM_TIMING
                    const long long c0 = clock64();
#endif
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
#ifdef ADM_GEMM_TIMING
                    const long long c1 = clock64();
#endif
                    const uint32_t sa = smem_u32(smem + stage * stage_bytes);
                    const uint32_t sb = sa + GEMM_A_STAGE;
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
                        const uint64_t da = make_smem_desc(sa + k * 32u, 16u, 1024);
                        const uint64_t db = make_smem_desc(sb + k * 32u, 16u, 1024);
                        umma_bf16_pair(tmem_d, da, db, idesc, (ki > 0 || k > 0) ? 1u : 0u);
                    }
#ifdef ADM_GEMM_TIMING
                    const long long c2 = clock64();
#endif
                    umma_commit_pair(&empty_bar[stage]);
#ifdef ADM_GEMM_TIMING
                    const long long c3 = clock64();
                    t_wait += c1 - c0; t_issue += c2 - c1; t_commit += c3 - c2;
#endif
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
#ifdef ADM_GEMM_TIMING
                if (blockIdx.x == 0 && tile == first)
                    printf("[pair mma thread] k-iters %d stages %d: wait %lld issue %lld commit %lld clocks per k-iter\n",
                           p.k_total, num_stages, t_wait / p.k_total, t_issue / p.k_total, t_commit / p.k_total);
#endif
                umma_commit_pair(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // =========================================================== epilogue (4 warps per CTA, own 128 rows)
        const int quad = warp & 3;
        const int m = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = first; tile < num_tiles; tile += stride) {
            const int nt = tile % p.n_tiles;
            const int mt = 2 * (tile / p.n_tiles) + rank;
            int n0, h0, w0;
            decode_pix(p, mt, n0, h0, w0);
            const int wi = m % p.bw, hi = (m / p.bw) % p.bh, ni = m / (p.bw * p.bh);
            const long long pix = (static_cast<long long>(n0 + ni) * p.H + (h0 + hi)) * p.W + (w0 + wi);
            const bool row_ok = mt < p.m_tiles && pix < p.M;
            const long long c_base = pix * p.ldc;
            const int col_base = nt * p.bn;

            float* sbias = reinterpret_cast<float*>(smem + GEMM_SMEM_RING + 1024) + acc * 256;
            stage_bias_128(p, sbias, col_base, m);
            mbar_wait(&tfull_bar[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(quad * 32) << 16);
            epilogue_row(p, taddr, col_base, 0, c_base, pix * p.ldr, row_ok, sbias, 0, p.bn);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or remote arrives are in flight
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ halo conv kernel
// 3x3 implicit-GEMM conv whose pixel operand is loaded ONCE per 64-channel chunk instead of once per tap.  The pixel
// tile is 8 (w) x 16 (h); its 10 x 18 halo (zero-filled outside the image by TMA) lands in shared memory as 180 rows of
// 128 B.  A tap (dy, dx) is then just a shifted window of that halo: the UMMA A descriptor starts at row dy*10 + dx and
// walks its sixteen 8-row groups (= pixel rows of the tile) SBO = 10 rows apart.  The 128B-swizzle XOR is a function of
// the absolute shared-memory address on both the TMA and the UMMA side, so an unaligned start and a group stride that
// is not a multiple of 1024 B read back exactly what TMA wrote (tools/exp/umma_row_offset.cu, measured on B200).
// Per k-iteration (one tap x 64 channels) the SM ingests bn*128 B of weights + 23 KB / 9 of pixels instead of
// bn*128 B + 16 KB: -34 % for bn = 192.  K order is chunk-major (chunk, tap) — the weight TMA coordinates follow.
// Roles, accumulators and the epilogue are those of tc_gemm_kernel; two rings replace the single one:
//   halo ring  : HALO_BUFS buffers, filled one chunk step AHEAD (across tiles), released by a commit after the 9th tap
//   weight ring: as many bn*128 B stages as fit beside it
constexpr int HALO_W = 10, HALO_H = 18;
constexpr int HALO_TX_BYTES = HALO_W * HALO_H * 128;  // 23040
constexpr int HALO_BYTES = 23 * 1024;                 // buffer pitch, 1024-aligned
constexpr int HALO_BUFS = 3;

template <bool PRO>
__global__ void __launch_bounds__(PRO ? GEMM_PRO_THREADS : GEMM_THREADS, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = p.bn * 128;
    int nb = (GEMM_SMEM_RING - HALO_BUFS * HALO_BYTES) / b_bytes;
    if (nb > GEMM_MAX_STAGES) nb = GEMM_MAX_STAGES;
    uint8_t* b_ring = smem + HALO_BUFS * HALO_BYTES;

    uint64_t* bfull = reinterpret_cast<uint64_t*>(smem + GEMM_SMEM_RING);
    uint64_t* bempty = bfull + GEMM_MAX_STAGES;
    uint64_t* tfull_bar = bempty + GEMM_MAX_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* hfull = tempty_bar + 2;                    // halo buffer ready for the MMAs
    uint64_t* hempty = hfull + HALO_BUFS;
    uint64_t* hraw = PRO ? hempty + HALO_BUFS : hfull;   // halo buffer landed (PRO: still to be transformed)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(hempty + 2 * HALO_BUFS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < nb; ++i) {
            mbar_init(&bfull[i], 1);
            mbar_init(&bempty[i], 1);
        }
        for (int i = 0; i < HALO_BUFS; ++i) {
            mbar_init(&hfull[i], PRO ? GEMM_PRO_WARPS : 1);
            mbar_init(&hempty[i], 1);
            if (PRO) mbar_init(&hraw[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
        }
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();  // everything above touched only this kernel's own state; the predecessor's results are visible from here on

    const int num_tiles = p.m_tiles * p.n_tiles;

    if (warp == 0) {
        // =========================================================== TMA producer
        if (lane == 0) {
            int hs = 0, bs = 0;
            uint32_t hphase = 0, bphase = 0;
            auto load_halo = [&](int tile, int kc) {
                int n0, h0, w0;
                decode_pix(p, (tile / p.n_tiles) % p.m_tiles, n0, h0, w0);
                mbar_wait(&hempty[hs], hphase ^ 1, 5);
                mbar_expect_tx(&hraw[hs], HALO_TX_BYTES);
                uint8_t* dst = smem + hs * HALO_BYTES;
                if (kc < p.cchunks1)
                    tma_load_4d(dst, &tmA, &hraw[hs], kc * 64, w0 - 1, h0 - 1, n0);
                else
                    tma_load_4d(dst, &tmA2, &hraw[hs], (kc - p.cchunks1) * 64, w0 - 1, h0 - 1, n0);
                if (++hs == HALO_BUFS) { hs = 0; hphase ^= 1; }
            };
            // Halo loads run AHEAD of the weight tiles by `ahead` chunk steps (a cursor over (tile, chunk) that only this
            // thread advances): 1 step hides the TMA latency (and, with the GroupNorm prologue, the transform) under the nine
            // MMAs of the current chunk; three buffers: loading / transforming, multiplying, draining.
            int cur_tile = blockIdx.x, cur_kc = 0;
            auto next_halo = [&]() {
                if (cur_tile >= num_tiles) return;
                load_halo(cur_tile, cur_kc);
                if (++cur_kc == p.cchunks) { cur_kc = 0; cur_tile += gridDim.x; }
            };
            const int ahead = 1;  // (2 makes this thread wait for the buffer still being multiplied before it issues weights)
            for (int i = 0; i < ahead; ++i) next_halo();
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                for (int kc = 0; kc < p.cchunks; ++kc) {
                    next_halo();  // the halo `ahead` steps from now goes out before this step's nine weight tiles
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&bempty[bs], bphase ^ 1, 1);
                        uint8_t* sb = b_ring + bs * b_bytes;
                        if (p.debug & 2) {  // ADM_GEMM_DEBUG=2 (timing experiment, results are garbage): no weight traffic
                            mbar_arrive(&bfull[bs]);
                            if (++bs == nb) { bs = 0; bphase ^= 1; }
                            continue;
                        }
                        mbar_expect_tx(&bfull[bs], b_bytes);
                        if (!p.b_mn) {
                            tma_load_2d(sb, &tmB, &bfull[bs], (tap * p.cchunks + kc) * 64, nt * p.bn);
                        } else {
                            for (int c = 0; c * 64 < p.bn; ++c)
                                tma_load_3d(sb + c * 8192, &tmB, &bfull[bs], nt * p.bn + c * 64, 8 - tap, kc * 64);
                        }
                        if (++bs == nb) { bs = 0; bphase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== UMMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, p.bn, 0, p.b_mn);
            const uint32_t b_lbo = p.b_mn ? 8192u : 16u;
            const uint32_t b_kstep = (p.b_mn ? 2048u : 32u) >> 4;
            const uint64_t da0 = make_smem_desc(0, 16, HALO_W * 128), db0 = make_smem_desc(0, b_lbo, 1024);
            int hs = 0, bs = 0;
            uint32_t hphase = 0, bphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * 256;
                for (int kc = 0; kc < p.cchunks; ++kc) {
                    mbar_wait(&hfull[hs], hphase, 6);
                    tc_fence_after();
                    const uint64_t da_halo = da0 + (smem_u32(smem + hs * HALO_BYTES) >> 4);
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&bfull[bs], bphase, 3);
                        tc_fence_after();
                        const uint64_t da = da_halo + ((tap / 3) * HALO_W + tap % 3) * 8;  // (128 B rows) >> 4
                        const uint64_t db = db0 + (smem_u32(b_ring + bs * b_bytes) >> 4);
#pragma unroll
                        for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
                            umma_bf16(tmem_d, da + 2 * k, db + k * b_kstep, idesc, (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&bempty[bs]);
                        if (++bs == nb) { bs = 0; bphase ^= 1; }
                    }
                    umma_commit(&hempty[hs]);
                    if (++hs == HALO_BUFS) { hs = 0; hphase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp < 2 + GEMM_EPI_WARPS) {
        epilogue_warps<GEMM_CONV>(p, smem, tmem_base, tfull_bar, tempty_bar, warp, lane, num_tiles);
    } else if (PRO) {
        // =========================================================== GroupNorm + SiLU (+ dropout) prologue warps
        // Thread tt owns the 16-byte unit u = tt & 7 (8 channels) of the rows r = tt >> 3, + PT / 8, ... of each 180-row halo
        // tile.  A row is one halo pixel; rows outside the image were zero-filled by TMA and must stay zero (the conv pads
        // the ACTIVATED tensor), channels beyond the valid count get A = B = 0 (silu(0) = 0).
        constexpr int PT = 32 * GEMM_PRO_WARPS;
        const int tt = threadIdx.x - GEMM_THREADS;
        const int u = tt & 7, r0 = tt >> 3;
        float2* sco = reinterpret_cast<float2*>(smem + GEMM_SMEM_RING + 4096);  // [cchunks * 64] (A, B) of this tile's sample
        const int C = p.pro_c1 + p.pro_c2, V = C >> 3;
        unsigned long long seed = p.pro_seed;
        if (p.pro_seed_dev != nullptr) seed += *p.pro_seed_dev * 0x9E3779B97F4A7C15ull;
        const DropCtx dc = drop_ctx(seed, p.pro_drop_p);
        int hs = 0;
        uint32_t hphase = 0;
        int last_n = -1;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int n0, h0, w0;
            decode_pix(p, (tile / p.n_tiles) % p.m_tiles, n0, h0, w0);
            const bool writer = p.pro_out != nullptr && (tile % p.n_tiles) == 0;
            if (n0 != last_n) {  // (uniform over these warps) stage this sample's coefficients, padded per 64-channel chunk
                asm volatile("bar.sync 2, %0;" ::"n"(PT) : "memory");
                for (int i = tt; i < p.cchunks * 64; i += PT) {
                    const int kc = i >> 6, cc = i & 63;
                    const bool first = kc < p.cchunks1;
                    const int ch = first ? kc * 64 + cc : (kc - p.cchunks1) * 64 + cc;
                    const bool ok = ch < (first ? p.pro_c1 : p.pro_c2);
                    float2 ab = make_float2(0.f, 0.f);
                    if (ok) {
                        const float4 t = __ldg(p.pro_coef + 1LL * n0 * C + (first ? ch : p.pro_c1 + ch));
                        ab = p.pro_act ? make_float2(0.5f * t.x, 0.5f * t.y) : make_float2(t.x, t.y);
                    }
                    sco[i] = ab;
                }
                asm volatile("bar.sync 2, %0;" ::"n"(PT) : "memory");
                last_n = n0;
            }
            for (int kc = 0; kc < p.cchunks; ++kc) {
                mbar_wait(&hraw[hs], hphase, 7);
                uint8_t* buf = smem + hs * HALO_BYTES;
                float a[8], b[8];
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    const float4 t = *reinterpret_cast<const float4*>(sco + kc * 64 + u * 8 + j);
                    a[j] = t.x; b[j] = t.y; a[j + 1] = t.z; b[j + 1] = t.w;
                }
                const bool first = kc < p.cchunks1;
                const int cbase = first ? kc * 64 + u * 8 : p.pro_c1 + (kc - p.cchunks1) * 64 + u * 8;  // channel in the concat
                const bool ch_ok = (first ? kc * 64 + u * 8 : (kc - p.cchunks1) * 64 + u * 8) < (first ? p.pro_c1 : p.pro_c2);
#pragma unroll 1
                for (int r = r0; r < HALO_W * HALO_H; r += PT / 8) {
                    const int hy = r / HALO_W, hx = r - hy * HALO_W;
                    const int y = h0 - 1 + hy, x = w0 - 1 + hx;
                    if (y < 0 || y >= p.H || x < 0 || x >= p.W) continue;
                    uint4* cell = reinterpret_cast<uint4*>(buf + r * 128 + ((u ^ (r & 7)) << 4));
                    const uint4 raw = *cell;
                    const uint32_t w4[4] = {raw.x, raw.y, raw.z, raw.w};
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[2 * j] = __uint_as_float(w4[j] << 16);
                        v[2 * j + 1] = __uint_as_float(w4[j] & 0xFFFF0000u);
                    }
                    const long long pix = (static_cast<long long>(n0) * p.H + y) * p.W + x;
                    float ds[8];
                    if (p.pro_drop_p > 0.f)
                        dropout_scales(dc, static_cast<uint32_t>(static_cast<unsigned long long>(pix) * V + (cbase >> 3)), ds);
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        float y0 = fmaf(v[j], a[j], b[j]), y1 = fmaf(v[j + 1], a[j + 1], b[j + 1]);
                        if (p.pro_act) {
                            y0 = fmaf(y0, gemm_tanh_fast(y0), y0);
                            y1 = fmaf(y1, gemm_tanh_fast(y1), y1);
                        }
                        if (p.pro_drop_p > 0.f) { y0 *= ds[j]; y1 *= ds[j + 1]; }
                        const __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
                        o[j >> 1] = *reinterpret_cast<const uint32_t*>(&b2);
                    }
                    const uint4 res = make_uint4(o[0], o[1], o[2], o[3]);
                    *cell = res;
                    if (writer && ch_ok && hy >= 1 && hy <= HALO_H - 2 && hx >= 1 && hx <= HALO_W - 2)
                        *reinterpret_cast<uint4*>(p.pro_out + pix * p.pro_ldo + cbase) = res;
                }
                fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive(&hfull[hs]);
                if (++hs == HALO_BUFS) { hs = 0; hphase ^= 1; }
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

// ------------------------------------------------------------------------------------------------ tap-row wgrad kernel
// 3x3 weight gradient  dW[co][dy][dx][ci] = sum_q dY[q - (dy-1, 0)][co] * X[q + (0, dx-1)][ci]  with the vertical shift
// on the dY side and the horizontal shift on the X side, so that
//   * N = 192 = the THREE dx taps of one 64-channel X chunk, read as three windows of ONE w-halo tile (bw+2 pixels per
//     row): the MN-major B descriptor's chunk stride (LBO) is a single 128 B row — one MMA covers three taps;
//   * M = 128 = two (dy, 64-channel dY chunk) pairs, each an ordinary 64-pixel box loaded with its own vertical offset
//     (zero outside the image): Cout = 192 gives 9 M chunks = 4.5 tiles instead of 9 x 2 half-empty ones.
// Per k-iteration (64 pixels) an SM ingests 16 KB + ~9 KB for 128 x 192 x 64 MACs, against 16 KB + 24 KB per tap before.
// K = pixels (split across CTAs, fp32 atomics into the gradient arena), roles as in tc_gemm_kernel.
// K = pixels (split across CTAs, fp32 atomics into the gradient arena), roles as in tc_gemm_kernel.
__global__ void __launch_bounds__(GEMM_THREADS, 1)
tc_wgrad_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ GemmParams p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int halo_w = p.bw + 2;
    const int b_tx = p.bh * halo_w * 128;
    const int stage_bytes = GEMM_A_STAGE + ((b_tx + 1023) & ~1023);
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
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmB2);
        for (int i = 0; i < num_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
        }
        fence_barrier_init();
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();  // everything above touched only this kernel's own state; the predecessor's results are visible from here on

    // tile = (sp * m_tiles + mt) * n_tiles + nt : CTAs running together share a pixel range (split) — its dY / X boxes
    // stay L2-resident.  (A stream-K division of the (tile, k) space balanced the SMs but was 10 % slower: neighbouring
    // CTAs then stream different pixel ranges and the operand re-reads miss L2.)
    const int num_tiles = p.splits * p.m_tiles * p.n_tiles;
    const int m_chunks = 3 * p.co_chunks;

    if (warp == 0) {
        // =========================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nt = tile % p.n_tiles;
                const int mt = (tile / p.n_tiles) % p.m_tiles;
                const int sp = tile / (p.n_tiles * p.m_tiles);
                const int k_begin = sp * p.k_iters;
                const int k_end = min(p.k_total, k_begin + p.k_iters);
                const int mc0 = 2 * mt, mc1 = min(2 * mt + 1, m_chunks - 1);  // an odd tail repeats a chunk (masked)
                const int dy0 = mc0 / p.co_chunks, cc0 = mc0 % p.co_chunks;
                const int dy1 = mc1 / p.co_chunks, cc1 = mc1 % p.co_chunks;
                const bool src2 = nt >= p.cchunks1;
                const CUtensorMap* mb = src2 ? &tmB2 : &tmB;
                const int nc = (src2 ? nt - p.cchunks1 : nt) * 64;
                PixWalker px(p, k_begin);
                for (int ki = k_begin; ki < k_end; ++ki) {
                    const int n0 = px.n0, h0 = px.h0, w0 = px.w0;
                    px.advance(p);
                    mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    uint8_t* sa = smem + stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], GEMM_A_STAGE + b_tx);
                    tma_load_4d(sa, &tmA, &full_bar[stage], cc0 * 64, w0, h0 + 1 - dy0, n0);
                    tma_load_4d(sa + 8192, &tmA, &full_bar[stage], cc1 * 64, w0, h0 + 1 - dy1, n0);
                    tma_load_4d(sa + GEMM_A_STAGE, mb, &full_bar[stage], nc, w0 - 1, h0, n0);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== UMMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, 192, 1, 1);
            // a 16-pixel k-step is two 8-pixel runs of the halo tile: adjacent (one image row, bw >= 16) or one halo
            // row apart (bw == 8: two image rows)
            const uint32_t b_sbo = p.bw >= 16 ? 1024u : static_cast<uint32_t>(halo_w) * 128u;
            // everything that does not depend on the stage is hoisted: this one thread's issue loop is the critical
            // path (two runtime integer divisions per k-step cost 25 % of the kernel when they sat inside it)
            const uint64_t da0 = make_smem_desc(0, 8192, 1024), db0 = make_smem_desc(0, 128, b_sbo);
            uint32_t a_off[GEMM_BLOCK_K / 16], b_off[GEMM_BLOCK_K / 16];
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / 16; ++k) {
                const int px = k * 16;  // halo row of this k-step's first pixel at dx = 0
                a_off[k] = (k * 2048) >> 4;
                b_off[k] = (GEMM_A_STAGE + ((px / p.bw) * halo_w + px % p.bw) * 128) >> 4;
            }
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int sp = tile / (p.n_tiles * p.m_tiles);
                const int k_begin = sp * p.k_iters;
                const int k_end = min(p.k_total, k_begin + p.k_iters);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * 256;
                for (int ki = k_begin; ki < k_end; ++ki) {
                    mbar_wait(&full_bar[stage], phase, 3);
                    tc_fence_after();
                    const uint32_t sa16 = smem_u32(smem + stage * stage_bytes) >> 4;
#pragma unroll
                    for (int k = 0; k < GEMM_BLOCK_K / 16; ++k)
                        umma_bf16(tmem_d, da0 + (sa16 + a_off[k]), db0 + (sa16 + b_off[k]), idesc,
                                  (ki > k_begin || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == num_stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // =========================================================== epilogue: rows = (dy, co), three 64-column taps
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int m = quad * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int nt = tile % p.n_tiles;
            const int mt = (tile / p.n_tiles) % p.m_tiles;
            const int mc = 2 * mt + (m >> 6);
            const int dy = mc / p.co_chunks, co = (mc % p.co_chunks) * 64 + (m & 63);
            const bool row_ok = mc < m_chunks && co < p.M;
            const long long c_base = static_cast<long long>(co) * p.ldc;
            mbar_wait(&tfull_bar[acc], acc_phase, 4);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(quad * 32) << 16);
            // 3 taps x 64 columns (with two warps per quadrant: 96 columns each, taps {0, 1 lower half} / {1 upper, 2})
#pragma unroll 1
            for (int i = 0; i < (EPI_HALVES == 2 ? 2 : 3); ++i) {
                const int dx = half + i;
                const int cb = (half == 1 && i == 0) ? 32 : 0, ce = (EPI_HALVES == 2 && half == 0 && i == 1) ? 32 : 64;
                epilogue_row(p, taddr + dx * 64, nt * 64, (dy * 3 + dx) * p.c_col_lo, c_base, 0, row_ok, nullptr, cb, ce);
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
