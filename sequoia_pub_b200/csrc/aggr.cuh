// Kernels and launch helpers shared by the two aggregators (vis.cu: SummaryMixing ViS, vit.cu: softmax-attention ViT):
// bf16 hi/lo plane stores, row LayerNorm forward/backward, token means / sums, column sums (bias gradients), MSE, flat
// AdamW and the split-precision GEMM call builder.  Everything has internal linkage (one copy per translation unit).
#pragma once
#include "gemm.cuh"
#include <stdlib.h>

namespace sq {
namespace {

static inline size_t aup(size_t x) { return (x + 1023) / 1024 * 1024; }
constexpr int LN_RPB = 16;    // rows per block in the row-LayerNorm backward
// ------------------------------------------------------------------------------------------------ small kernels
__device__ __forceinline__ void store_planes4(bf16* hi, bf16* lo, size_t off, float4 v) {
    const bf16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    uint2 ph, pl;
    __nv_bfloat162 t;
    t = __halves2bfloat162(h0, h1); ph.x = *reinterpret_cast<uint32_t*>(&t);
    t = __halves2bfloat162(h2, h3); ph.y = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint2*>(hi + off) = ph;
    if (lo) {
        t = __halves2bfloat162(__float2bfloat16_rn(v.x - __bfloat162float(h0)), __float2bfloat16_rn(v.y - __bfloat162float(h1)));
        pl.x = *reinterpret_cast<uint32_t*>(&t);
        t = __halves2bfloat162(__float2bfloat16_rn(v.z - __bfloat162float(h2)), __float2bfloat16_rn(v.w - __bfloat162float(h3)));
        pl.y = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint2*>(lo + off) = pl;
    }
}

// x_in [B,N,D] + pos [N,D] -> x (fp32 + planes).  tformer_lin.py:100
__global__ void prep_input_kernel(const float* __restrict__ x_in, const float* __restrict__ pos, float* __restrict__ x,
                                  bf16* __restrict__ xh, bf16* __restrict__ xl, int N, int D, long long total4) {
    const int D4 = D / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / D4; const int c = (int)(i - row * D4) * 4;
        const float4 a = *reinterpret_cast<const float4*>(x_in + i * 4);
        const float4 p = *reinterpret_cast<const float4*>(pos + (size_t)(row % N) * D + c);
        const float4 v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
        *reinterpret_cast<float4*>(x + i * 4) = v;
        store_planes4(xh, xl, (size_t)i * 4, v);
    }
}

// mean over the N tokens of every slide: x [B*N, D] fp32 -> [B, D] (fp32 and/or planes).  tformer_lin.py:22,103
// block = (slide, 128-column slab); the 8 warps stride over the tokens, partial sums are combined in warp order.
__global__ void __launch_bounds__(256) group_mean_kernel(const float* __restrict__ x, int N, int D, float scale, float* __restrict__ out,
                                                         bf16* __restrict__ oh, bf16* __restrict__ ol) {
    __shared__ float4 red[8][32];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 128 + lane * 4;
    const int b = blockIdx.y;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < D) {
#pragma unroll 4
        for (int n = w; n < N; n += 8) {
            const float4 v = *reinterpret_cast<const float4*>(x + ((size_t)b * N + n) * D + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    red[w][lane] = acc;
    __syncthreads();
    if (w == 0 && c < D) {
        float4 t = red[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) { const float4 u = red[i][lane]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
        t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
        if (out) *reinterpret_cast<float4*>(out + (size_t)b * D + c) = t;
        if (oh) store_planes4(oh, ol, (size_t)b * D + c, t);
    }
}

// sum of hi+lo over groups of gs consecutive rows: planes [groups*gs, C] -> [groups, C] (fp32 and/or planes).
// block = (group, 256-column slab); 8 warps stride over the rows, 8 columns per lane.
__global__ void __launch_bounds__(256) group_sum_planes_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, long long ld, int gs, int C,
                                                               float* __restrict__ out, bf16* __restrict__ oh, bf16* __restrict__ ol) {
    __shared__ float red[8][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 256 + lane * 8;
    const int g = blockIdx.y;
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.f;
    if (c < C) {
#pragma unroll 2
        for (int r = w; r < gs; r += 8) {
            const size_t off = ((size_t)g * gs + r) * ld + c;
            const uint4 h = *reinterpret_cast<const uint4*>(hi + off);
            const uint4 l = *reinterpret_cast<const uint4*>(lo + off);
            const __nv_bfloat162* hp = reinterpret_cast<const __nv_bfloat162*>(&h);
            const __nv_bfloat162* lp = reinterpret_cast<const __nv_bfloat162*>(&l);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 hf = __bfloat1622float2(hp[j]), lf = __bfloat1622float2(lp[j]);
                a[2 * j] += hf.x + lf.x; a[2 * j + 1] += hf.y + lf.y;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[w][lane * 8 + j] = a[j];
    __syncthreads();
    const int col = blockIdx.x * 256 + threadIdx.x;
    if (col < C) {
        float t = red[0][threadIdx.x];
#pragma unroll
        for (int i = 1; i < 8; ++i) t += red[i][threadIdx.x];
        const size_t o = (size_t)g * C + col;
        if (out) out[o] = t;
        if (oh) {
            const bf16 h0 = __float2bfloat16_rn(t);
            oh[o] = h0; ol[o] = __float2bfloat16_rn(t - __bfloat162float(h0));
        }
    }
}

// out[c] = scale * sum_p in[p*ld + c]   (fixed order); block = 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ in, int P, int C, long long ld, float scale, float* __restrict__ out) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int p = ty; p < P; p += 8) a += in[(size_t)p * ld + c];
    }
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < C) {
        float t = red[0][tx];
#pragma unroll
        for (int i = 1; i < 8; ++i) t += red[i][tx];
        out[c] = t * scale;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of two values (256 threads); result broadcast to all threads
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* red) {
    a = warp_sum(a); b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) red[w] = make_float2(a, b);
    __syncthreads();
    float2 t = make_float2(0.f, 0.f);
    for (int i = 0; i < nw; ++i) { t.x += red[i].x; t.y += red[i].y; }
    return t;
}

// LayerNorm over the last dim of every row (eps 1e-5, biased variance): x [rows, D] -> planes (+ mean, rstd).
// tformer_lin.py:55,92 (nn.LayerNorm(dim))
__global__ void __launch_bounds__(256) ln_rows_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int D, float eps, float* __restrict__ mean_out,
                                                          float* __restrict__ rstd_out, bf16* __restrict__ oh, bf16* __restrict__ ol) {
    __shared__ float2 red[8];
    const size_t row = blockIdx.x;
    const float* xr = x + row * D;
    float s = 0.f;
    for (int c = threadIdx.x * 4; c < D; c += 1024) { const float4 v = *reinterpret_cast<const float4*>(xr + c); s += v.x + v.y + v.z + v.w; }
    const float mean = block_sum2(s, 0.f, red).x / (float)D;
    float q = 0.f;
    for (int c = threadIdx.x * 4; c < D; c += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
        q += a * a + b * b + cc * cc + d * d;
    }
    const float rstd = rsqrtf(block_sum2(q, 0.f, red).x / (float)D + eps);
    if (threadIdx.x == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
    for (int c = threadIdx.x * 4; c < D; c += 1024) {
        const float4 v = *reinterpret_cast<const float4*>(xr + c);
        const float4 g = *reinterpret_cast<const float4*>(gamma + c);
        const float4 b = *reinterpret_cast<const float4*>(beta + c);
        store_planes4(oh, ol, row * D + c, make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y,
                                                        (v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w));
    }
}

// Backward of the row LayerNorm: dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)) (+ res), and per-block partial
// sums of dgamma = dy*xhat, dbeta = dy.  Each thread owns V float4 column groups for all rows of its block.
template <int V>
__global__ void __launch_bounds__(256) ln_rows_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          const float* __restrict__ gamma, const float* __restrict__ res, int rows, int D,
                                                          float* __restrict__ dx, bf16* __restrict__ dxh, bf16* __restrict__ dxl,
                                                          float* __restrict__ part) {
    __shared__ float2 red[8];
    float4 dg[V], db[V], gm[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        dg[k] = make_float4(0.f, 0.f, 0.f, 0.f); db[k] = dg[k];
        const int c = (threadIdx.x + k * 256) * 4;
        gm[k] = c < D ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = 0; i < LN_RPB; ++i) {
        const size_t row = (size_t)blockIdx.x * LN_RPB + i;
        if (row >= (size_t)rows) break;
        const float mu = mean[row], rs = rstd[row];
        float4 a[V], xh[V];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = (threadIdx.x + k * 256) * 4;
            if (c < D) {
                const float4 d = *reinterpret_cast<const float4*>(dy + row * D + c);
                const float4 v = *reinterpret_cast<const float4*>(x + row * D + c);
                xh[k] = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
                a[k] = make_float4(d.x * gm[k].x, d.y * gm[k].y, d.z * gm[k].z, d.w * gm[k].w);
                s1 += a[k].x + a[k].y + a[k].z + a[k].w;
                s2 += a[k].x * xh[k].x + a[k].y * xh[k].y + a[k].z * xh[k].z + a[k].w * xh[k].w;
                dg[k].x += d.x * xh[k].x; dg[k].y += d.y * xh[k].y; dg[k].z += d.z * xh[k].z; dg[k].w += d.w * xh[k].w;
                db[k].x += d.x; db[k].y += d.y; db[k].z += d.z; db[k].w += d.w;
            }
        }
        const float2 t = block_sum2(s1, s2, red);
        const float c1 = t.x / (float)D, c2 = t.y / (float)D;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = (threadIdx.x + k * 256) * 4;
            if (c < D) {
                float4 o = make_float4(rs * (a[k].x - c1 - xh[k].x * c2), rs * (a[k].y - c1 - xh[k].y * c2),
                                       rs * (a[k].z - c1 - xh[k].z * c2), rs * (a[k].w - c1 - xh[k].w * c2));
                if (res) { const float4 r = *reinterpret_cast<const float4*>(res + row * D + c); o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }
                if (dx) *reinterpret_cast<float4*>(dx + row * D + c) = o;
                if (dxh) store_planes4(dxh, dxl, row * D + c, o);
            }
        }
    }
    float* pg = part + (size_t)blockIdx.x * 2 * D;
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = (threadIdx.x + k * 256) * 4;
        if (c < D) { *reinterpret_cast<float4*>(pg + c) = dg[k]; *reinterpret_cast<float4*>(pg + D + c) = db[k]; }
    }
}

__device__ __forceinline__ float half_sum(float v) {     // sum over the 16 lanes of a half-warp
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Backward of GELU(LayerNorm64(pre)) per head group of 64 columns (tformer_lin.py:20,22): din = dL/d(GELU output).
// A half-warp owns one (row, head) group, 4 columns per lane; block = 8 warps x 16 rows x 128 columns.
__global__ void __launch_bounds__(256) ln64_bwd_kernel(const float* __restrict__ din, const float* __restrict__ pre,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int HD, int rpw,
                                                       bf16* __restrict__ oh, bf16* __restrict__ ol, float* __restrict__ part) {
    __shared__ float red[8][2][128];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cl = (lane >> 4) * 64 + (lane & 15) * 4;      // column within the 128-wide slab
    const int col = blockIdx.x * 128 + cl;
    const bool col_ok = col < HD;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f), bt = g, dg = g, db = g;
    if (col_ok) { g = *reinterpret_cast<const float4*>(gamma + col); bt = *reinterpret_cast<const float4*>(beta + col); }
    for (int i = 0; i < rpw; ++i) {
        const size_t row = ((size_t)blockIdx.y * 8 + w) * rpw + i;
        const bool ok = col_ok && row < (size_t)rows;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f), d = p;
        if (ok) { p = *reinterpret_cast<const float4*>(pre + row * HD + col); d = *reinterpret_cast<const float4*>(din + row * HD + col); }
        const float mean = half_sum(p.x + p.y + p.z + p.w) * (1.0f / 64.0f);
        const float4 e = make_float4(p.x - mean, p.y - mean, p.z - mean, p.w - mean);
        const float rstd = rsqrtf(half_sum(e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w) * (1.0f / 64.0f) + 1e-5f);
        const float4 xh = make_float4(e.x * rstd, e.y * rstd, e.z * rstd, e.w * rstd);
        const float4 dy = make_float4(d.x * dgelu_f(xh.x * g.x + bt.x), d.y * dgelu_f(xh.y * g.y + bt.y),
                                      d.z * dgelu_f(xh.z * g.z + bt.z), d.w * dgelu_f(xh.w * g.w + bt.w));
        const float4 a = make_float4(dy.x * g.x, dy.y * g.y, dy.z * g.z, dy.w * g.w);
        const float c1 = half_sum(a.x + a.y + a.z + a.w) * (1.0f / 64.0f);
        const float c2 = half_sum(a.x * xh.x + a.y * xh.y + a.z * xh.z + a.w * xh.w) * (1.0f / 64.0f);
        if (ok) {
            store_planes4(oh, ol, row * HD + col, make_float4(rstd * (a.x - c1 - xh.x * c2), rstd * (a.y - c1 - xh.y * c2),
                                                               rstd * (a.z - c1 - xh.z * c2), rstd * (a.w - c1 - xh.w * c2)));
            dg.x += dy.x * xh.x; dg.y += dy.y * xh.y; dg.z += dy.z * xh.z; dg.w += dy.w * xh.w;
            db.x += dy.x; db.y += dy.y; db.z += dy.z; db.w += dy.w;
        }
    }
    *reinterpret_cast<float4*>(&red[w][0][cl]) = dg;
    *reinterpret_cast<float4*>(&red[w][1][cl]) = db;
    __syncthreads();
    const int t = threadIdx.x;          // 256 threads = 2 x 128 outputs
    const int which = t >> 7, c = t & 127;
    if (blockIdx.x * 128 + c < HD) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += red[i][which][c];
        part[((size_t)blockIdx.y * 2 + which) * HD + blockIdx.x * 128 + c] = s;
    }
}

// GELU(LayerNorm64(pre)) per head group, rows = slides (the summary branch after its split-K GEMM; tformer_lin.py:22).
// One warp per row and 128-column slab (two head groups).
__global__ void __launch_bounds__(256) ln64_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int rows, int HD, bf16* __restrict__ oh, bf16* __restrict__ ol) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int col = blockIdx.x * 128 + (lane >> 4) * 64 + (lane & 15) * 4;
    const size_t row = (size_t)blockIdx.y * 8 + w;
    const bool ok = col < HD && row < (size_t)rows;
    float4 p = make_float4(0.f, 0.f, 0.f, 0.f), g = p, bt = p;
    if (ok) { p = *reinterpret_cast<const float4*>(pre + row * HD + col); g = *reinterpret_cast<const float4*>(gamma + col); bt = *reinterpret_cast<const float4*>(beta + col); }
    const float mean = half_sum(p.x + p.y + p.z + p.w) * (1.0f / 64.0f);
    const float4 e = make_float4(p.x - mean, p.y - mean, p.z - mean, p.w - mean);
    const float rstd = rsqrtf(half_sum(e.x * e.x + e.y * e.y + e.z * e.z + e.w * e.w) * (1.0f / 64.0f) + 1e-5f);
    if (ok) store_planes4(oh, ol, row * HD + col, make_float4(gelu_f(e.x * rstd * g.x + bt.x), gelu_f(e.y * rstd * g.y + bt.y),
                                                               gelu_f(e.z * rstd * g.z + bt.z), gelu_f(e.w * rstd * g.w + bt.w)));
}

// g[b*N + n, :] = scale * src[b, :]   (backward of the token mean, tformer_lin.py:103)
__global__ void bcast_rows_kernel(const float* __restrict__ src, int N, int D, long long total4, float scale, float* __restrict__ out,
                                  bf16* __restrict__ oh, bf16* __restrict__ ol) {
    const int D4 = D / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / D4; const int c = (int)(i - row * D4) * 4;
        float4 v = *reinterpret_cast<const float4*>(src + (row / N) * D + c);
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
        *reinterpret_cast<float4*>(out + row * D + c) = v;
        store_planes4(oh, ol, (size_t)row * D + c, v);
    }
}

// dpos[n, :] = sum_b g[b, n, :]   (pos_emb1D broadcasts over the batch, tformer_lin.py:100)
__global__ void pos_grad_kernel(const float* __restrict__ g, int B, int N, int D, float* __restrict__ dpos) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * D / 4) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < B; ++b) {
        const float4 v = *reinterpret_cast<const float4*>(g + (size_t)b * N * D + i * 4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4*>(dpos + i * 4) = a;
}

// MSELoss (mean over all elements, src/vit.py:129,166) and d loss / d pred = 2 (pred - y) / n
constexpr int MSE_BLOCKS = 512;
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ y, long long n, float scale,
                                                  float* __restrict__ dpred, float* __restrict__ partial) {
    __shared__ float2 red[8];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float d = pred[i] - y[i];
        s += d * d;
        if (dpred) dpred[i] = d * scale;
    }
    const float2 t = block_sum2(s, 0.f, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t.x;
}
__global__ void __launch_bounds__(256) mse_final_kernel(const float* __restrict__ partial, int np, float inv_n, float* __restrict__ loss) {
    __shared__ float2 red[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < np; i += 256) s += partial[i];
    const float2 t = block_sum2(s, 0.f, red);
    if (threadIdx.x == 0) *loss = t.x * inv_n;
}

// torch.optim.AdamW single-tensor update order (amsgrad=False) on a flat buffer, plus the refreshed bf16 planes.
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, bf16* __restrict__ ph, bf16* __restrict__ pl, long long n4,
                                                    float decay_mul, float b1, float b2, float eps, float step_size, float bc2_sqrt, float gscale) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        float4 Mv = reinterpret_cast<float4*>(m)[i], Vv = reinterpret_cast<float4*>(v)[i];
        float* pp = &P.x; const float* gg = &G.x; float* mm = &Mv.x; float* vv = &Vv.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = gg[j] * gscale;
            const float pj = pp[j] * decay_mul;
            mm[j] = mm[j] + (1.0f - b1) * (gr - mm[j]);                 // exp_avg.lerp_(grad, 1 - beta1)
            vv[j] = vv[j] * b2 + ((1.0f - b2) * gr) * gr;               // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
            pp[j] = pj + (-step_size * mm[j]) / denom;                  // param.addcdiv_(exp_avg, denom, value=-step_size)
        }
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = Mv;
        reinterpret_cast<float4*>(v)[i] = Vv;
        if (ph) store_planes4(ph, pl, (size_t)i * 4, P);
    }
}

// ------------------------------------------------------------------------------------------------ GEMM call builder
struct GB {
    GemmArgs g;
    GB(int M, int N, int K) {
        memset(&g, 0, sizeof(g));
        g.M = M; g.N = N; g.K = K; g.nterms = 3; g.e.alpha = 1.0f; g.e.rowbias_div = 1;
    }
    GB& A(const void* hi, const void* lo, long long ld, int mn = 0) { g.A.hi = (const bf16*)hi; g.A.lo = (const bf16*)lo; g.A.ld = ld; g.A.mn_major = mn; return *this; }
    GB& B(const void* hi, const void* lo, long long ld, int mn = 0) { g.B.hi = (const bf16*)hi; g.B.lo = (const bf16*)lo; g.B.ld = ld; g.B.mn_major = mn; return *this; }
    GB& bias(const float* b) { g.e.bias = b; return *this; }
    GB& rowbias(const float* rb, int div, long long ld) { g.e.rowbias = rb; g.e.rowbias_div = div; g.e.ld_rowbias = ld; return *this; }
    GB& res(const float* r, long long ld) { g.e.res_f32 = r; g.e.ld_res = ld; return *this; }
    GB& act(int a) { g.e.act = a; return *this; }
    GB& ln64(const float* gm, const float* bt) { g.e.act = ACT_LN64_GELU; g.e.ln_gamma = gm; g.e.ln_beta = bt; return *this; }
    GB& dgelu(const float* aux, long long ld) { g.e.act = ACT_MUL_DGELU; g.e.aux = aux; g.e.ld_aux = ld; return *this; }
    GB& save_pre(float* p, long long ld) { g.e.save_pre = p; g.e.ld_pre = ld; return *this; }
    GB& out_f32(float* o, long long ld) { g.e.out_f32 = o; g.e.ld_f32 = ld; return *this; }
    GB& out_planes(void* hi, void* lo, long long ld) { g.e.out_hi = (bf16*)hi; g.e.out_lo = (bf16*)lo; g.e.ld_bf = ld; return *this; }
    GB& alpha(float a) { g.e.alpha = a; return *this; }
    GB& bn(int b) { g.block_n = b; return *this; }
    GB& akoff(int k) { g.a_koff_per_ntile = k; return *this; }
    GB& bdiag_dgrad(int map_mn, int map_k) { g.b_koff_per_ntile = 64; g.b_nadj_per_ntile = -64; g.b_map_mn = map_mn; g.b_map_k = map_k; return *this; }
    GB& diag64() { g.diag64 = 1; g.block_n = 64; return *this; }
    // workspace for stream-K scheduling of GEMMs whose tile count does not fill whole waves
    GB& sk(void* ws, size_t ws_bytes) { g.workspace = (float*)ws; g.workspace_bytes = ws_bytes; return *this; }
    // split-K for GEMMs with too few output tiles to occupy the machine
    GB& auto_split(void* ws, size_t ws_bytes) {
        const int bnn = g.block_n ? g.block_n : 128;
        const long long tiles = (long long)((g.M + 127) / 128) * ((g.N + bnn - 1) / bnn);
        const int kb = (g.K + 63) / 64 * g.nterms;
        int s = (int)(num_sms() / (tiles > 0 ? tiles : 1));
        if (s > kb / 6) s = kb / 6;
        if (s > 16) s = 16;
        while (s > 1 && (size_t)s * g.M * g.N * 4 > ws_bytes) --s;
        if (s > 1) { g.split_k = s; g.workspace = (float*)ws; g.workspace_bytes = ws_bytes; }
        return *this;
    }
    int run(cudaStream_t st) { return gemm_launch(g, st); }
};

#define SQ_TRY(x) do { if ((x) != 0) return -1; } while (0)

// The summary branch of a layer is a chain of tiny, latency-bound kernels that is independent of the big local-branch
// GEMMs next to it; it is enqueued on the library's side stream (common.cu: side_stream / side_fork / side_join) so that it
// soaks up the SMs the persistent GEMM kernels leave idle in their last, partially filled wave.

static int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(err)); return -1; }
    return 0;
}


static int launch_ln_rows_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, const float* res,
                              int rows, int D, float* dx, bf16* dxh, bf16* dxl, float* part, float* dgamma_dbeta, cudaStream_t st) {
    const int nblk = (rows + LN_RPB - 1) / LN_RPB;
    const int v = (D / 4 + 255) / 256;
    if (v <= 1) ln_rows_bwd_kernel<1><<<nblk, 256, 0, st>>>(dy, x, mean, rstd, gamma, res, rows, D, dx, dxh, dxl, part);
    else if (v <= 2) ln_rows_bwd_kernel<2><<<nblk, 256, 0, st>>>(dy, x, mean, rstd, gamma, res, rows, D, dx, dxh, dxl, part);
    else if (v <= 4) ln_rows_bwd_kernel<4><<<nblk, 256, 0, st>>>(dy, x, mean, rstd, gamma, res, rows, D, dx, dxh, dxl, part);
    else ln_rows_bwd_kernel<8><<<nblk, 256, 0, st>>>(dy, x, mean, rstd, gamma, res, rows, D, dx, dxh, dxl, part);
    colsum_kernel<<<(2 * D + 31) / 32, 256, 0, st>>>(part, nblk, 2 * D, 2LL * D, 1.0f, dgamma_dbeta);   // gamma and beta are adjacent
    return check_launch("ln_rows_bwd");
}

static int launch_ln64_bwd(const float* din, const float* pre, const float* gamma, const float* beta, int rows, int HD, bf16* oh, bf16* ol,
                           float* part, float* dgamma_dbeta, cudaStream_t st) {
    const int rpw = rows >= 1024 ? 16 : 1;
    const int nrb = (rows + 8 * rpw - 1) / (8 * rpw);
    ln64_bwd_kernel<<<dim3((HD + 127) / 128, nrb), 256, 0, st>>>(din, pre, gamma, beta, rows, HD, rpw, oh, ol, part);
    colsum_kernel<<<(2 * HD + 31) / 32, 256, 0, st>>>(part, nrb, 2 * HD, 2LL * HD, 1.0f, dgamma_dbeta);
    return check_launch("ln64_bwd");
}

// bias gradient: column sums of a [rows, C] planes matrix, rows = groups * gs; via per-slide sums (fixed order)
static int launch_bias_grad(const bf16* hi, const bf16* lo, int groups, int gs, int C, float* gsum, float* out, cudaStream_t st) {
    group_sum_planes_kernel<<<dim3((C + 255) / 256, groups), 256, 0, st>>>(hi, lo, C, gs, C, gsum, nullptr, nullptr);
    colsum_kernel<<<(C + 31) / 32, 256, 0, st>>>(gsum, groups, C, C, 1.0f, out);
    return check_launch("bias_grad");
}

}  // namespace
}  // namespace sq
