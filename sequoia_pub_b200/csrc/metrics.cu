// Per-step training metrics on the device (reference: src/vit.py:167-168 — sklearn mean_absolute_error and
// he2rna.compute_correlations, src/he2rna.py:140-149, both on host copies of pred / labels every step; the per-gene
// numpy loop costs ~1.3 s per step on the CPU at 20530 genes).  One thread per gene walks the batch column in float64
// (np.corrcoef computes in float64); genes with constant labels are skipped, NaN correlations (constant predictions)
// dropped, the rest averaged in a fixed order.  SMAPE of evaluate() (src/vit.py:32-33,269): 100 / len(A) * sum(2|F-A| / (|A|+|F|))
// over the whole [B, G] array with len(A) = B, element terms in float32 like numpy, summed in float64.
#include "gemm.cuh"
#include "../../include/sequoia_b200.h"

namespace sq {

__global__ void __launch_bounds__(256) gene_metrics_kernel(const float* __restrict__ labels, const float* __restrict__ preds, int B, int G,
                                                           double* __restrict__ part /* [gridDim.x][4]: sum r, n valid, sum |err|, sum smape terms */) {
    __shared__ double red[4][8];
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    double r = 0.0, valid = 0.0, ae = 0.0, sm = 0.0;
    if (g < G) {
        double sy = 0.0, sp = 0.0;
        float ymin = INFINITY, ymax = -INFINITY;
        for (int b = 0; b < B; ++b) {
            const float y = labels[(size_t)b * G + g], p = preds[(size_t)b * G + g];
            sy += y; sp += p; ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
            ae += fabs((double)y - (double)p);
            sm += (double)__fdiv_rn(__fmul_rn(2.0f, fabsf(__fsub_rn(p, y))), __fadd_rn(fabsf(y), fabsf(p)));
        }
        if (ymax > ymin) {                                   // len(np.unique(y_true)) > 1
            const double my = sy / B, mp = sp / B;
            double cyy = 0.0, cpp = 0.0, cyp = 0.0;
            for (int b = 0; b < B; ++b) {
                const double dy = (double)labels[(size_t)b * G + g] - my, dp = (double)preds[(size_t)b * G + g] - mp;
                cyy += dy * dy; cpp += dp * dp; cyp += dy * dp;
            }
            const double c = cyp / sqrt(cyy * cpp);          // np.corrcoef: c / sqrt(d_i d_j); NaN when a variance is 0
            if (c == c) { r = fmin(1.0, fmax(-1.0, c)); valid = 1.0; }   // corrcoef clips to [-1, 1]; NaNs are dropped
        }
    }
    double v[4] = {r, valid, ae, sm};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        part[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
    }
}

__global__ void gene_metrics_final_kernel(const double* __restrict__ part, int nblk, int B, int G, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double sr = 0.0, nv = 0.0, sa = 0.0, ss = 0.0;
    for (int i = 0; i < nblk; ++i) { sr += part[4 * i]; nv += part[4 * i + 1]; sa += part[4 * i + 2]; ss += part[4 * i + 3]; }
    out[0] = (float)(sa / ((double)B * G));                  // mean absolute error
    out[1] = (float)(sr / nv);                               // mean per-gene Pearson r (NaN when no gene is valid, like np.mean([]))
    out[2] = (float)nv;
    out[3] = (float)(100.0 / (double)B * ss);                // smape(A, F): len(A) is the batch size
}

}  // namespace sq

using namespace sq;

extern "C" {

size_t sq_step_metrics_scratch_bytes(int num_outputs) { return (size_t)((num_outputs + 255) / 256) * 4 * sizeof(double); }

int sq_step_metrics(const float* labels, const float* preds, int batch, int num_outputs, float* out4, void* scratch, size_t scratch_bytes,
                    void* stream) {
    if (!labels || !preds || !out4 || !scratch) { set_error("step_metrics: null pointer"); return -1; }
    if (batch <= 0 || num_outputs <= 0) { set_error("step_metrics: empty batch"); return -1; }
    const int nblk = (num_outputs + 255) / 256;
    if (scratch_bytes < (size_t)nblk * 4 * sizeof(double)) { set_error("step_metrics: scratch too small"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    gene_metrics_kernel<<<nblk, 256, 0, st>>>(labels, preds, batch, num_outputs, (double*)scratch);
    gene_metrics_final_kernel<<<1, 32, 0, st>>>((const double*)scratch, nblk, batch, num_outputs, out4);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("step_metrics: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // extern "C"
