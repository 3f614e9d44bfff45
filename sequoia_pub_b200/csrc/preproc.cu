// Image preprocessing in front of the UNI extractor (reference: pre_processing/compute_features_hdf5.py:53-56,125-126):
// `transforms.Resize(224)` of a PIL image = Pillow's separable, antialiased bilinear resample (libImaging/Resample.c) in 22-bit
// fixed point, horizontal pass rounded to uint8 before the vertical pass.  All integer arithmetic, so the GPU result is
// bit-identical to Pillow's; the (data independent) coefficient tables are computed on the host in double exactly like
// precompute_coeffs / normalize_coeffs_8bpc do.  oracle/resize_oracle.py is the CPU restatement pinned against Pillow.
#include "gemm.cuh"
#include "../../include/sequoia_b200.h"

namespace sq {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

// One thread per output element (pixel, channel).  axis 1: along x (in [n,H,Win,3] -> out [n,H,Wout,3]);
// axis 0: along y (in [n,Hin,W,3] -> out [n,Hout,W,3]).
__global__ void resize_pass_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n, int in_len, int out_len, int other, int axis,
                                   const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
    const long long total = (long long)n * out_len * other * 3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % 3); long long t = i / 3;
        int o, q;                                    // o: index along the resized axis, q: index along the other axis
        if (axis == 1) { o = (int)(t % out_len); t /= out_len; q = (int)(t % other); }
        else { q = (int)(t % other); t /= other; o = (int)(t % out_len); }
        const int img = (int)(t / (axis == 1 ? other : out_len));
        const int x0 = bounds[2 * o], cnt = bounds[2 * o + 1];
        const int* k = kk + (size_t)o * ksize;
        int acc = 1 << (RS_PRECISION_BITS - 1);
        if (axis == 1) {
            const uint8_t* row = in + (((long long)img * other + q) * in_len + x0) * 3 + c;
            for (int x = 0; x < cnt; ++x) acc += (int)row[x * 3] * k[x];
        } else {
            const uint8_t* col = in + (((long long)img * in_len + x0) * other + q) * 3 + c;
            for (int x = 0; x < cnt; ++x) acc += (int)col[(long long)x * other * 3] * k[x];
        }
        acc >>= RS_PRECISION_BITS;                   // arithmetic shift, then clip8
        out[i] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
    }
}

}  // namespace sq

using namespace sq;

extern "C" {

int sq_resize_ksize(int in_size, int out_size) {
    if (in_size <= 0 || out_size <= 0) return 0;
    double filterscale = (double)in_size / (double)out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    return (int)ceil(1.0 * filterscale) * 2 + 1;
}

int sq_resize_coeffs(int in_size, int out_size, int* bounds_host, int* coeffs_host) {
    if (in_size <= 0 || out_size <= 0 || !bounds_host || !coeffs_host) { set_error("resize_coeffs: bad arguments"); return -1; }
    const double scale = (double)in_size / (double)out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale;        // bilinear (triangle) filter
    const int ksize = (int)ceil(support) * 2 + 1;
    const double ss = 1.0 / filterscale;
    double w[64];
    if (ksize > 64) { set_error("resize_coeffs: shrink factor too large"); return -1; }
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5); if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5); if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            double t = (x + xmin - center + 0.5) * ss;
            if (t < 0.0) t = -t;
            w[x] = t < 1.0 ? 1.0 - t : 0.0;
            ww += w[x];
        }
        for (int x = 0; x < ksize; ++x) {
            double v = 0.0;
            if (x < xmax) v = ww != 0.0 ? w[x] / ww : w[x];
            coeffs_host[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << RS_PRECISION_BITS)) : (int)(0.5 + v * (1 << RS_PRECISION_BITS));
        }
        bounds_host[2 * xx] = xmin; bounds_host[2 * xx + 1] = xmax;
    }
    return 0;
}

int sq_resize_bilinear_u8(const void* in, int n, int Hin, int Win, void* out, int Hout, int Wout, const int* xbounds, const int* xcoeffs,
                          int xksize, const int* ybounds, const int* ycoeffs, int yksize, void* tmp, void* stream) {
    if (n <= 0) return 0;
    if (!in || !out) { set_error("resize: null pointer"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const bool need_x = Wout != Win, need_y = Hout != Hin;
    if ((need_x && (!xbounds || !xcoeffs)) || (need_y && (!ybounds || !ycoeffs)) || (need_x && need_y && !tmp)) { set_error("resize: missing tables / scratch"); return -1; }
    auto blocks = [](long long total) { long long b = (total + 255) / 256; return (unsigned)(b > 148LL * 32 ? 148LL * 32 : b); };
    const uint8_t* src = (const uint8_t*)in;
    if (!need_x && !need_y) { cudaMemcpyAsync(out, in, (size_t)n * Hin * Win * 3, cudaMemcpyDeviceToDevice, st); }
    if (need_x) {                                    // horizontal pass first (Resample.c)
        uint8_t* dst = need_y ? (uint8_t*)tmp : (uint8_t*)out;
        resize_pass_kernel<<<blocks((long long)n * Hin * Wout * 3), 256, 0, st>>>(src, dst, n, Win, Wout, Hin, 1, xbounds, xcoeffs, xksize);
        src = dst;
    }
    if (need_y) resize_pass_kernel<<<blocks((long long)n * Hout * Wout * 3), 256, 0, st>>>(src, (uint8_t*)out, n, Hin, Hout, Wout, 0, ybounds, ycoeffs, yksize);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("resize: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // extern "C"
