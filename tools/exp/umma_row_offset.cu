// Experiment for round 2 (halo re-use of the conv A operand): can a K-major SWIZZLE_128B UMMA A-descriptor start at a
// row that is NOT a multiple of 8 (start address not 1024 B aligned), and can its 8-row groups be SBO = 10 rows apart
// (so that a 3x3 tap is just a shifted window of ONE halo tile in shared memory)?  One CTA: TMA-loads A [256 x 64] and
// B [16 x 64] (bf16, 128B-swizzled), issues D[128 x 16] = A_window * B^T with
//     row(m) = off + (m / 8) * sbo_rows + (m % 8)
// and compares with the host result, for several (off, sbo_rows) and both settings of the descriptor's base-offset field.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I adm_b200/csrc -o gpurun_out/umma_row_offset tools/exp/umma_row_offset.cu -lcuda
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "ptx.cuh"
using namespace adm;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;                                   // LBO (unused for K-major SW128)
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(base_off & 7) << 49;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                                            int off, int sbo_rows, int use_base_off, float* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;            // 256 rows x 128 B
    uint8_t* sB = smem + 32768;    // 16 rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 2048);
    uint64_t* bar2 = bar + 1;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); fence_proxy_async_smem(); }
    if (warp == 1) tmem_alloc(tptr, 32);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tbase = *tptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 32768 + 2048);
        tma_load_2d(sA, &mA, bar, 0, 0);
        tma_load_2d(sB, &mB, bar, 0, 0);
        mbar_wait(bar, 0, 1);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
        for (int kk = 0; kk < 4; ++kk) {
            const uint32_t a = smem_u32(sA) + off * 128 + kk * 32;
            const uint32_t b = smem_u32(sB) + kk * 32;
            umma_bf16(tbase, desc_sw128(a, sbo_rows * 128, use_base_off ? (a >> 7) & 7 : 0), desc_sw128(b, 1024, 0), idesc,
                      kk > 0);
        }
        umma_commit(bar2);
    }
    mbar_wait(bar2, 0, 2);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld_x16(tbase + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tbase, 32); }
}

static PFN_cuTensorMapEncodeTiled enc;
static void make_map(CUtensorMap* m, void* p, int rows, int box_rows) {
    cuuint64_t gd[2] = {64, (cuuint64_t)rows}; cuuint64_t gs[1] = {128}; cuuint32_t bx[2] = {64, (cuuint32_t)box_rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
    void* fn; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    enc = (PFN_cuTensorMapEncodeTiled)fn;
    std::vector<__nv_bfloat16> hA(256 * 64), hB(16 * 64);
    std::vector<float> fA(256 * 64), fB(16 * 64);
    srand(1);
    for (int i = 0; i < 256 * 64; ++i) { fA[i] = (rand() % 17) - 8; hA[i] = __float2bfloat16(fA[i]); }
    for (int i = 0; i < 16 * 64; ++i) { fB[i] = (rand() % 9) - 4; hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dO;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 16 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap mA, mB; make_map(&mA, dA, 256, 256); make_map(&mB, dB, 16, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    const int cfgs[][2] = {{0, 8}, {3, 8}, {8, 8}, {0, 10}, {1, 10}, {3, 10}, {11, 10}, {22, 10}, {5, 18}};
    for (auto& c : cfgs)
        for (int ub = 0; ub < 2; ++ub) {
            k<<<1, 128, 40960>>>(mA, mB, c[0], c[1], ub, dO);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> o(128 * 16);
            cudaMemcpy(o.data(), dO, o.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0; int bad = 0;
            for (int m = 0; m < 128; ++m) {
                const int row = c[0] + (m / 8) * c[1] + (m % 8);
                for (int n = 0; n < 16; ++n) {
                    float ref = 0;
                    if (row < 256) for (int kk = 0; kk < 64; ++kk) ref += fA[row * 64 + kk] * fB[n * 64 + kk];
                    const double d = fabs(ref - o[m * 16 + n]);
                    if (row < 256) { if (d > maxerr) maxerr = d; if (d > 0.5) ++bad; }
                }
            }
            printf("off=%2d sbo_rows=%2d base_offset_field=%d : %s max|err|=%.1f wrong=%d/2048 (%s)\n", c[0], c[1], ub,
                   bad == 0 ? "EXACT" : "WRONG", maxerr, bad, cudaGetErrorString(e));
        }
    return 0;
}
