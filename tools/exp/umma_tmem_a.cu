// Experiment for round 3 (attention K9 v2): the A operand of tcgen05.mma taken from TENSOR MEMORY, so that the softmax
// probabilities never touch shared memory.
//   1. S[128 x 128] = Q K^T by an ordinary smem x smem MMA (fp32, TMEM columns [0, 128)).
//   2. Each thread (lane = row) reads its S row, converts to bf16 and writes it back IN PLACE as packed pairs with
//      tcgen05.st.32x32b: element k of row m lands in lane m, column k / 2, half k % 2 (columns [0, 64)).
//   3. O[128 x 64] = P V with A = P from TMEM (K-major by construction), B = V from shared memory (MN-major), 8 k-steps,
//      the A address advancing 8 columns per k-step; accumulator in columns [128, 192).
// Host check: O == bf16(Q K^T) V exactly (small integers, exact in bf16/fp32).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I adm_b200/csrc -o gpurun_out/umma_tmem_a tools/exp/umma_tmem_a.cu -lcuda
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "ptx.cuh"
using namespace adm;

__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap mQ, const __grid_constant__ CUtensorMap mK,
                                            const __grid_constant__ CUtensorMap mV, float* out_o, float* out_s) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;            // 128 rows x 128 B
    uint8_t* sK = smem + 16384;    // 128 rows x 128 B
    uint8_t* sV = smem + 32768;    // 128 rows (keys) x 128 B (64 d)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
    uint64_t* bar_s = bar + 1;
    uint64_t* bar_o = bar + 2;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1); mbar_init(bar_s, 1); mbar_init(bar_o, 1);
        fence_barrier_init(); fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tptr, 256);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 3 * 16384);
        tma_load_2d(sQ, &mQ, bar, 0, 0);
        tma_load_2d(sK, &mK, bar, 0, 0);
        tma_load_2d(sV, &mV, bar, 0, 0);
        mbar_wait(bar, 0, 1);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
        for (int kk = 0; kk < 4; ++kk)
            umma_bf16(tbase, make_smem_desc(smem_u32(sQ) + kk * 32, 16, 1024), make_smem_desc(smem_u32(sK) + kk * 32, 16, 1024),
                      idesc, kk > 0);
        umma_commit(bar_s);
    }
    mbar_wait(bar_s, 0, 2);
    tc_fence_after();
    const uint32_t trow = tbase + (static_cast<uint32_t>(warp * 32) << 16);
    const int row = warp * 32 + lane;
    // S row -> registers (all 128 columns BEFORE any in-place write), dump for the host, pack to bf16 pairs
    uint32_t pk[64];
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(trow + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out_s[row * 128 + c0 + j] = __uint_as_float(v[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            pk[c0 / 2 + j] = *reinterpret_cast<const uint32_t*>(&b2);
        }
    }
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) w[j] = pk[c0 + j];
        tmem_st_x8(trow + c0, w);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, 64, 0, 1);  // A K-major (TMEM), B MN-major (V: d contiguous)
        for (int kk = 0; kk < 8; ++kk)  // 16 keys per k-step: A advances 8 columns, B advances 16 rows of 128 B
            umma_bf16_ts(tbase + 128, tbase + kk * 8, make_smem_desc(smem_u32(sV) + kk * 2048, 8192, 1024), idesc, kk > 0);
        umma_commit(bar_o);
    }
    mbar_wait(bar_o, 0, 3);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(trow + 128 + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out_o[row * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tbase, 256); }
}

static PFN_cuTensorMapEncodeTiled enc;
static void make_map(CUtensorMap* m, void* p, int rows) {
    cuuint64_t gd[2] = {64, (cuuint64_t)rows}; cuuint64_t gs[1] = {128}; cuuint32_t bx[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    enc = (PFN_cuTensorMapEncodeTiled)fn;
    const int N = 128, D = 64;
    std::vector<__nv_bfloat16> hQ(N * D), hK(N * D), hV(N * D);
    std::vector<float> fQ(N * D), fK(N * D), fV(N * D);
    srand(3);
    for (int i = 0; i < N * D; ++i) {
        fQ[i] = (rand() % 5) - 2; fK[i] = (rand() % 3) - 1; fV[i] = (rand() % 7) - 3;
        hQ[i] = __float2bfloat16(fQ[i]); hK[i] = __float2bfloat16(fK[i]); hV[i] = __float2bfloat16(fV[i]);
    }
    __nv_bfloat16 *dQ, *dK, *dV; float *dO, *dS;
    cudaMalloc(&dQ, N * D * 2); cudaMalloc(&dK, N * D * 2); cudaMalloc(&dV, N * D * 2);
    cudaMalloc(&dO, N * D * 4); cudaMalloc(&dS, N * N * 4);
    cudaMemcpy(dQ, hQ.data(), N * D * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dK, hK.data(), N * D * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, hV.data(), N * D * 2, cudaMemcpyHostToDevice);
    CUtensorMap mQ, mK, mV; make_map(&mQ, dQ, N); make_map(&mK, dK, N); make_map(&mV, dV, N);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 53248);
    k<<<1, 128, 53248>>>(mQ, mK, mV, dO, dS);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(N * D), s(N * N);
    cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(s.data(), dS, s.size() * 4, cudaMemcpyDeviceToHost);
    int bad_s = 0, bad_o = 0; double maxerr = 0;
    std::vector<float> S(N * N);
    for (int m = 0; m < N; ++m)
        for (int n = 0; n < N; ++n) {
            float r = 0;
            for (int kk = 0; kk < D; ++kk) r += fQ[m * D + kk] * fK[n * D + kk];
            S[m * N + n] = r;
            if (fabs(r - s[m * N + n]) > 0.5) ++bad_s;
        }
    for (int m = 0; m < N; ++m)
        for (int d = 0; d < D; ++d) {
            float r = 0;
            for (int n = 0; n < N; ++n) r += S[m * N + n] * fV[n * D + d];  // |S| <= 128: exact in bf16
            const double err = fabs(r - o[m * D + d]);
            if (err > maxerr) maxerr = err;
            if (err > 0.5) ++bad_o;
        }
    printf("S = Q K^T (smem x smem): wrong=%d/%d\n", bad_s, N * N);
    printf("O = P V with A = P from TMEM (in-place bf16 pairs, +8 columns per k-step): %s max|err|=%.1f wrong=%d/%d (%s)\n",
           bad_o == 0 ? "EXACT" : "WRONG", maxerr, bad_o, N * D, cudaGetErrorString(e));
    return bad_o == 0 && bad_s == 0 ? 0 : 1;
}
