// Host-side job description of the plane-slot weight-gradient / column-sum kernels (wgrad_planes.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct i2sdf_handle;

namespace i2sdf {

constexpr int kWgMaxTerms = 28;
constexpr int kWgMaxJobs = 16;

struct WgTerm {
    const uint8_t* P;      // 256-column slot: rows of dW (out features)
    const uint8_t* X;      // slot with xchunks * 8 columns: columns of dW (in features)
    int job;               // terms of one job are consecutive
    int xchunks;           // 32, or 6 for the positional-encoding slots
    float* colsum;         // optional: colsum[j] += sum_m P[m][j]  (bias gradient), j < colsum_n
    int colsum_n;
    int p_planes, x_planes;   // planes per sub tile of each operand: 2 = hi + lo, 1 = hi only (adjoint slots); 0 is read as 2
};
struct WgJob {
    float* dW;             // fp32 [rows][ld], accumulated into
    int ld, rows, cols;
};
struct WgArgs {
    long long ntiles;
    int nterms, njobs, variant;
    WgTerm terms[kWgMaxTerms];
    WgJob jobs[kWgMaxJobs];
};

struct CsJob {
    const uint8_t* slot;   // 256-column slot
    const float* w;        // optional per-point weights w_k[m] = w[m * wstride + k]
    int wstride;
    float* out;            // out[k * ostride + j] += sum_m w_k[m] X[m][j], j < n, k < nw
    int n;
    int nw;                // weight vectors sharing ONE pass over the slot (1..3; w null: 1 = plain column sums)
    int ostride;
    int nplanes;           // planes of the slot (0 is read as 2)
    long long wrows;       // rows the weight array w covers (0: all M); points beyond it carry weight 0
};
struct CsArgs {
    long long ntiles, M;
    int njobs;
    CsJob jobs[kWgMaxJobs];
};

int wgrad_planes_launch(const i2sdf_handle* h, const WgArgs& args, cudaStream_t st);
int planes_colsum_launch(const i2sdf_handle* h, const CsArgs& args, cudaStream_t st);
int planes_pack_launch(const float* X, int ld, int width, long long M, uint8_t* slot, int chunks, cudaStream_t st);
int planes_unpack_launch(const uint8_t* slot, int chunks, long long M, float* X, int ld, int width, cudaStream_t st);

}  // namespace i2sdf
