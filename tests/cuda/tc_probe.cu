// Hardware probe for the tcgen05 building blocks in i2sdf_b200/csrc/tc_common.cuh:
// one 128 x 256 x 64 bf16 GEMM (fp32 accumulate in TMEM) with operands in the canonical K-major/no-swizzle layout,
// staged by cp.async.bulk + mbarrier, issued by one thread, read back with tcgen05.ld.  Compared with a CPU GEMM.
// variant 1 swaps the LBO/SBO roles in the descriptor (diagnostic only).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../i2sdf_b200/csrc/tc_common.cuh"
using namespace i2sdf::tc;

constexpr int M = 128, N = 256, K = 64;

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ Ag, const uint8_t* __restrict__ Bg, float* __restrict__ D,
                                                      int variant, int dbuf) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* As = smem;                 // M*K*2 = 16 KB
    uint8_t* Bs = smem + M * K * 2;     // N*K*2 = 32 KB
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_load, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_load, M * K * 2 + N * K * 2);
        bulk_g2s(As, Ag, M * K * 2, &bar_load);
        bulk_g2s(Bs, Bg, N * K * 2, &bar_load);
        mbar_wait(&bar_load, 0);
        tc_fence_after();
        const uint32_t idesc = instr_desc_bf16(M, N);
        const uint32_t lboA = M * 16, lboB = N * 16, sbo = 128;
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t da = variant == 0 ? smem_desc(smem_u32(As) + ks * 2 * lboA, lboA, sbo) : smem_desc(smem_u32(As) + ks * 2 * lboA, sbo, lboA);
            uint64_t db = variant == 0 ? smem_desc(smem_u32(Bs) + ks * 2 * lboB, lboB, sbo) : smem_desc(smem_u32(Bs) + ks * 2 * lboB, sbo, lboB);
            mma_bf16_ss(tmem_base + dbuf * 256, da, db, idesc, ks > 0);
        }
        mma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_base + dbuf * 256 + c * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); uint32_t r = u + 0x7FFF + ((u >> 16) & 1); return (uint16_t)(r >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    std::vector<float> A(M * K), B(N * K);
    srand(1);
    for (auto& x : A) x = bf2f(f2bf((rand() / (float)RAND_MAX) - 0.5f));
    for (auto& x : B) x = bf2f(f2bf((rand() / (float)RAND_MAX) - 0.5f));
    std::vector<uint8_t> Ap(M * K * 2), Bp(N * K * 2);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(A[r * K + k]); memcpy(&Ap[seg_off<M>(r, k / 8) + (k % 8) * 2], &h, 2); }
    for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(B[r * K + k]); memcpy(&Bp[seg_off<N>(r, k / 8) + (k % 8) * 2], &h, 2); }
    std::vector<float> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = (float)s; }
    uint8_t *dA, *dB; float* dD;
    cudaMalloc(&dA, Ap.size()); cudaMalloc(&dB, Bp.size()); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, Ap.data(), Ap.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, Bp.data(), Bp.size(), cudaMemcpyHostToDevice);
    const int smem = M * K * 2 + N * K * 2;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int ok_any = 0;
    for (int variant = 0; variant < 2; ++variant) for (int dbuf = 0; dbuf < 2; ++dbuf) {
        cudaMemset(dD, 0, M * N * 4);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, variant, dbuf);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d dbuf %d: CUDA error %s\n", variant, dbuf, cudaGetErrorString(e)); return 2; }
        std::vector<float> out(M * N);
        cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; for (int i = 0; i < M * N; ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
        printf("variant %d dbuf %d: max abs err %.3e  (D[0]=%f ref %f, D[last]=%f ref %f)\n", variant, dbuf, maxerr, out[0], ref[0], out[M * N - 1], ref[M * N - 1]);
        if (variant == 0 && maxerr < 1e-3) ok_any++;
    }
    printf(ok_any == 2 ? "TC_PROBE_OK\n" : "TC_PROBE_FAIL\n");
    return ok_any == 2 ? 0 : 1;
}
