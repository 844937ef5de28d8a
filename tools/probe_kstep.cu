// Stand-alone microbenchmark (development tool, NOT part of libi2sdf_b200.so): what paces one k step of the chain kernels' MMA
// issuer?  DESIGN.md §5.1 measures 462 clocks per k step (3 x tcgen05.mma 128x256x16 = 384 clocks of tensor pipe) and lists two
// suspects: shared-memory traffic (operand reads + weight-ring writes + epilogue operand stores ~ 60 KB per k step) and the issue
// loop itself.  This program runs the product's issue loop (same descriptors, same ring, same barriers, tc_common.cuh) on dummy
// operands with each traffic source switchable:
//     nmma   1 | 2 | 3      MMAs per k step (operand reads 12 | 24 | 36 KB)
//     stream 0 | 1          weight ring filled by cp.async.bulk (16 KB per k step) or just signalled
//     epi    0 | 1 | 2      16 epilogue warps idle | running the product's per-item sequence on registers (MUFU softplus, bf16 split,
//                           st.shared of the operand, fence.proxy.async) | the same plus the TMEM load of the accumulator
// and prints clocks per k step (issue loop only, and including the drain of the last MMAs) per variant.  The epilogue warps are
// NOT chained to the MMAs here (they free-run the same number of items), so the numbers isolate throughput interference from the
// hand-off latency the real kernels add on top.
//
// Build + run (on the GPU box):  bash tools/build_probe.sh && ./tools/probe_kstep
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace i2sdf::tc;

constexpr int TM = 128, NSTAGE = 4, STAGE = 16384, A_CHUNKS = 36, A_PART = A_CHUNKS * TM * 16;
constexpr int N_EPI = 16, NTHREADS = (2 + N_EPI) * 32;
constexpr uint32_t LBO_A = TM * 16, SBO = 128;
constexpr size_t kSmem = 1024 + 2 * (size_t)A_PART + NSTAGE * STAGE + 256;

struct Args {
    const uint8_t* w;     // 16 blocks of 16 KB (one layer's worth of packed weights; contents irrelevant)
    long long* out;       // [grid][4]: issue clocks, issue + drain clocks, epilogue clocks, -
    int nops, nmma, stream, epi;
    int half;             // 1 (with issue = 1): N = 128 MMAs, one ring stage = TWO k steps of one N half (6 MMAs = 384 tensor clocks per stage) - the
                          //    N-split schedule (half A of an op completes T/2 before half B, hiding the first-chunk latency of the next op)
    int issue;            // 0: the product's loop; 1: unrolled over the 4 ring stages (compile-time stage / descriptor offsets), no tcgen05 fence after the
                          //    weight barrier (the weights arrive through the async proxy; the fence is only needed behind the a_ready waits)
};

__device__ __forceinline__ void tmem_ld8p(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__global__ void __launch_bounds__(NTHREADS, 1) probe(const Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_align1024(smem_raw);
    uint8_t* A_hi = smem;
    uint8_t* A_lo = smem + A_PART;
    uint8_t* ring = smem + 2 * A_PART;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + NSTAGE * STAGE);
    uint64_t* empty = full + NSTAGE;
    uint64_t* done = empty + NSTAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < (2 * A_PART + NSTAGE * STAGE) / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // weight producer (chain_producer of tc_chain.cuh)
            uint32_t stage = 0, phase = 0;
            for (int op = 0; op < a.nops; ++op)
                for (int ks = 0; ks < 16; ++ks) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (a.stream) {
                        mbar_arrive_expect_tx(&full[stage], STAGE);
                        bulk_g2s(ring + stage * STAGE, a.w + (size_t)ks * STAGE, STAGE, &full[stage]);
                    } else {
                        mbar_arrive(&full[stage]);
                    }
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {                                // MMA issuer (chain_mma of tc_chain.cuh without the a_ready waits)
        const uint32_t a_hi_s = smem_u32(A_hi), a_lo_s = smem_u32(A_lo), ring_s = smem_u32(ring);
        const uint64_t dA_hi0 = smem_desc(a_hi_s, LBO_A, SBO), dA_lo0 = smem_desc(a_lo_s, LBO_A, SBO);
        const uint32_t idesc = instr_desc_bf16(TM, 256);
        const uint64_t dB0 = smem_desc(ring_s, 256u * 16u, SBO);
        const uint32_t lo_off = 256u * 32u;
        uint32_t stage = 0, phase = 0;
        const long long t0 = clock64();
        if (a.issue == 0) {
            for (int op = 0; op < a.nops; ++op) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(op & 1) * 256u;
                for (int ks = 0; ks < 16; ++ks) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da_hi = dA_hi0 + (uint64_t)(((uint32_t)ks * 2u * LBO_A) >> 4);
                    const uint64_t da_lo = dA_lo0 + (uint64_t)(((uint32_t)ks * 2u * LBO_A) >> 4);
                    const uint64_t db_hi = dB0 + (uint64_t)((stage * (uint32_t)STAGE) >> 4);
                    const uint64_t db_lo = db_hi + (uint64_t)(lo_off >> 4);
                    if (elect_one_sync()) {
                        mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, ks > 0 ? 1u : 0u);
                        if (a.nmma > 1) mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                        if (a.nmma > 2) mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                        mma_commit(&empty[stage]);
                    }
                    __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        } else {
            // 16 k steps = 4 passes over the 4-stage ring: the stage (barrier addresses, B descriptors) is a compile-time constant and
            // the A descriptors advance by a constant per k step
            for (int op = 0; op < a.nops; ++op) {
                const uint32_t d_tmem = tmem_base + (uint32_t)(op & 1) * 256u;
                uint64_t da_hi = dA_hi0, da_lo = dA_lo0;
#pragma unroll 1
                for (int pass = 0; pass < 4; ++pass) {
#pragma unroll
                    for (int st = 0; st < NSTAGE; ++st) {
                        mbar_wait(&full[st], phase);
                        const uint64_t db_hi = dB0 + (uint64_t)(((uint32_t)st * (uint32_t)STAGE) >> 4);
                        const uint64_t db_lo = db_hi + (uint64_t)(lo_off >> 4);
                        if (a.half) {
                            // stage = [k step a: hi 4 KB | lo 4 KB][k step b: hi | lo], N = 128: LBO 2048
                            const uint32_t id128 = instr_desc_bf16(TM, 128);
                            const uint64_t b0h = smem_desc(ring_s + (uint32_t)st * STAGE, 128u * 16u, SBO), b0l = b0h + (4096u >> 4), b1h = b0h + (8192u >> 4), b1l = b0h + (12288u >> 4);
                            const uint64_t da_hi2 = da_hi + (uint64_t)((2u * LBO_A) >> 4), da_lo2 = da_lo + (uint64_t)((2u * LBO_A) >> 4);
                            if (elect_one_sync()) {
                                mma_bf16_ss(d_tmem, da_hi, b0h, id128, (pass | st) ? 1u : 0u);
                                mma_bf16_ss(d_tmem, da_lo, b0h, id128, 1u);
                                mma_bf16_ss(d_tmem, da_hi, b0l, id128, 1u);
                                mma_bf16_ss(d_tmem, da_hi2, b1h, id128, 1u);
                                mma_bf16_ss(d_tmem, da_lo2, b1h, id128, 1u);
                                mma_bf16_ss(d_tmem, da_hi2, b1l, id128, 1u);
                                mma_commit(&empty[st]);
                            }
                            __syncwarp();
                            da_hi += (uint64_t)((2u * LBO_A) >> 4);        // (stays inside the 36-chunk operand: 4 passes x 4 stages x 1 chunk pair + 1)
                            da_lo += (uint64_t)((2u * LBO_A) >> 4);
                            continue;
                        }
                        if (elect_one_sync()) {
                            mma_bf16_ss(d_tmem, da_hi, db_hi, idesc, (pass | st) ? 1u : 0u);
                            if (a.nmma > 1) mma_bf16_ss(d_tmem, da_lo, db_hi, idesc, 1u);
                            if (a.nmma > 2) mma_bf16_ss(d_tmem, da_hi, db_lo, idesc, 1u);
                            mma_commit(&empty[st]);
                        }
                        __syncwarp();
                        da_hi += (uint64_t)((2u * LBO_A) >> 4);
                        da_lo += (uint64_t)((2u * LBO_A) >> 4);
                    }
                    phase ^= 1;
                }
            }
        }
        const long long t1 = clock64();
        if (elect_one_sync()) mma_commit(done);
        __syncwarp();
        mbar_wait(done, 0);
        const long long t2 = clock64();
        if (lane == 0) { a.out[blockIdx.x * 4 + 0] = t1 - t0; a.out[blockIdx.x * 4 + 1] = t2 - t0; }
    } else if (a.epi) {                                    // epilogue warps: the per-item sequence of tc_sdf8_kernel, free-running
        const int q = warp & 3, sub = (warp - 2) >> 2, row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        float seed = 0.001f * (float)(tid + 1);
        const long long t0 = clock64();
        for (int op = 0; op < a.nops; ++op) {
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
                const int col0 = it * 32 + sub * 8;
                uint32_t v[8];
                if (a.epi > 1) {
                    tmem_ld8p(tmem_base + lane_base + (uint32_t)(op & 1) * 256u + (uint32_t)col0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(seed + 0.01f * (float)j);
                }
                float hv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float x = __uint_as_float(v[j]) * 1e-3f + seed;
                    const float e = ex2a(-fabsf(x) * 144.26950408889634f);
                    hv[j] = fmaf(lg2a(1.0f + e), 0.0069314718055994531f, fmaxf(x, 0.0f));
                }
                seed = hv[3] * 0.5f + 0.001f;
                uint32_t h[4], lo[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) split_bf16x2(hv[2 * i], hv[2 * i + 1], h[i], lo[i]);
                const uint32_t off = seg_off<TM>(row, col0 >> 3);
                *reinterpret_cast<uint4*>(A_hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4*>(A_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
            }
        }
        const long long t1 = clock64();
        if (tid == 64) a.out[blockIdx.x * 4 + 2] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    int nops = argc > 1 ? atoi(argv[1]) : 64, grid = argc > 2 ? atoi(argv[2]) : 148;
    uint8_t* w;
    long long* out;
    CK(cudaMalloc(&w, 16 * STAGE));
    CK(cudaMemset(w, 0, 16 * STAGE));
    CK(cudaMalloc(&out, sizeof(long long) * 4 * grid));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    std::vector<long long> h(4 * grid);
    printf("grid %d CTAs, %d ops x 16 k steps per CTA; clocks per k step (median over CTAs)\n", grid, nops);
    printf("issue nmma stream epi |  issue loop   issue+drain   epilogue per op\n");
    for (int issue = 0; issue < 2; ++issue)
    for (int epi = 0; epi < 3; ++epi)
        for (int stream = 0; stream < 2; ++stream)
            for (int nmma = 1; nmma <= 3; ++nmma) {
                Args a{w, out, nops, nmma, stream, epi, 0, issue};
                for (int rep = 0; rep < 2; ++rep) {            // second launch is the measurement (weights L2-resident)
                    CK(cudaMemset(out, 0, sizeof(long long) * 4 * grid));
                    probe<<<grid, NTHREADS, kSmem>>>(a);
                    CK(cudaGetLastError());
                    CK(cudaDeviceSynchronize());
                }
                CK(cudaMemcpy(h.data(), out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost));
                std::vector<long long> c0, c1, c2;
                for (int b = 0; b < grid; ++b) { c0.push_back(h[b * 4]); c1.push_back(h[b * 4 + 1]); c2.push_back(h[b * 4 + 2]); }
                auto med = [](std::vector<long long>& v) { std::sort(v.begin(), v.end()); return (double)v[v.size() / 2]; };
                const double ks = (double)nops * 16.0;
                printf("  %d     %d     %d     %d  |  %9.1f   %9.1f   %12.1f\n", issue, nmma, stream, epi, med(c0) / ks, med(c1) / ks, med(c2) / (double)nops);
            }
    printf("N-split schedule: N = 128 MMAs, 6 per ring stage (= 2 k steps, 384 tensor clocks): clocks per STAGE\n");
    for (int epi = 0; epi < 3; ++epi)
        for (int stream = 0; stream < 2; ++stream) {
            Args a{w, out, nops, 3, stream, epi, 1, 1};
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaMemset(out, 0, sizeof(long long) * 4 * grid));
                probe<<<grid, NTHREADS, kSmem>>>(a);
                CK(cudaGetLastError());
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h.data(), out, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost));
            std::vector<long long> c0, c1, c2;
            for (int b = 0; b < grid; ++b) { c0.push_back(h[b * 4]); c1.push_back(h[b * 4 + 1]); c2.push_back(h[b * 4 + 2]); }
            auto med = [](std::vector<long long>& v) { std::sort(v.begin(), v.end()); return (double)v[v.size() / 2]; };
            const double ks = (double)nops * 16.0;
            printf("  half  stream %d  epi %d  |  %9.1f   %9.1f   %12.1f\n", stream, epi, med(c0) / ks, med(c1) / ks, med(c2) / (double)nops);
        }
    cudaFree(w);
    cudaFree(out);
    return 0;
}
