// ViS aggregator (reference: src/tformer_lin.py — SummaryMixing :18-26, MultiHeadSummary :39-48, FeedForward :51-61,
// SummaryTransformer :73-77, ViS :97-106) forward, hand-written backward, fused MSE and flat AdamW
// (callers: src/vit.py:163-166,175-180; src/main.py:180-183).
//
// Every Linear runs as a split-precision (bf16 hi/lo, 3 MMAs per tile, fp32 TMEM accumulation) tcgen05 GEMM from
// gemm.cuh with its bias / LayerNorm(64)+GELU / GELU / residual / GELU' epilogue fused; what is left are bandwidth
// kernels (row LayerNorm, token means, column sums) that use warp-shuffle reductions and fixed reduction orders, so
// the whole step is run-to-run deterministic (no float atomics).
//
// Algebra used (all exact identities):
//   * the 16 heads' f / s weights are contiguous in the flat parameter buffer -> one GEMM for all heads;
//   * mean_tokens(s(x)) = s(mean_tokens(x)) -> the summary branch is a [B,D] x [D,H*64] GEMM;
//   * c(cat[local, t]) = Wc[:, :64] local + (Wc[:, 64:] t + bc) -> per-head 64x64 block-diagonal GEMM + per-slide row bias.
#include "gemm.cuh"
#include "../../include/sequoia_b200.h"
#include <stdlib.h>

namespace sq {

constexpr int MAXL = 64;

struct VisDims { int D, L, H, N, G, HD; long long Gpad; };

struct LayerOff { long long lnl_g, lnl_b, lns_g, lns_b, ws, bs, wf, bf, wc, bc, wp, bp, fg, fb, w1, b1, w2, b2; };
struct VisLayout { long long pos; LayerOff lay[MAXL]; long long hg, hb, wh, bh, total; };

static int vis_dims(const sq_vis_config* c, VisDims* d) {
    if (!c) { set_error("vis: null config"); return -1; }
    if (c->input_dim <= 0 || c->input_dim % 64 != 0 || c->input_dim > 8192) { set_error("vis: input_dim %d must be a multiple of 64 in (0, 8192]", c->input_dim); return -1; }
    if (c->depth <= 0 || c->depth > MAXL) { set_error("vis: depth %d out of range", c->depth); return -1; }
    if (c->nheads <= 0 || c->nheads > 128) { set_error("vis: nheads %d out of range", c->nheads); return -1; }
    if (c->num_clusters <= 0 || c->num_outputs <= 0) { set_error("vis: num_clusters / num_outputs must be positive"); return -1; }
    d->D = c->input_dim; d->L = c->depth; d->H = c->nheads; d->N = c->num_clusters; d->G = c->num_outputs;
    d->HD = c->nheads * 64; d->Gpad = (c->num_outputs + 7) / 8 * 8;
    return 0;
}

// Flat parameter layout (fp32 elements; every tensor starts on a 64-element boundary so that the bf16 planes, which
// share the offsets, are 128-byte aligned for TMA).  Order: pos, layers 0..L-1, head — contiguous per backward stage.
static void vis_layout(const VisDims& d, VisLayout* L) {
    long long off = 0;
    auto take = [&](long long n) { long long o = off; off += (n + 63) / 64 * 64; return o; };
    const long long D = d.D, HD = d.HD;
    L->pos = take((long long)d.N * D);
    for (int l = 0; l < d.L; ++l) {
        LayerOff& o = L->lay[l];
        o.lnl_g = take(HD); o.lnl_b = take(HD); o.lns_g = take(HD); o.lns_b = take(HD);
        o.ws = take(HD * D); o.bs = take(HD); o.wf = take(HD * D); o.bf = take(HD);
        o.wc = take(HD * 128); o.bc = take(HD);
        o.wp = take(D * HD); o.bp = take(D);
        o.fg = take(D); o.fb = take(D);
        o.w1 = take(D * D); o.b1 = take(D); o.w2 = take(D * D); o.b2 = take(D);
    }
    L->hg = take(D); L->hb = take(D);
    L->wh = take((long long)d.G * D); L->bh = take(d.G);
    L->total = off;
}

static inline size_t aup(size_t x) { return (x + 1023) / 1024 * 1024; }

struct LayerAct { size_t x_f32, x_hi, x_lo, xm_hi, xm_lo, fpre, loc_hi, loc_lo, spre, t_hi, t_lo, rb, cpre, out_hi, out_lo, x1, ln_mean, ln_rstd, h_hi, h_lo, upre, u_hi, u_lo; };
struct VisAct { LayerAct lay[MAXL]; size_t xL, pooled, hmean, hrstd, z_hi, z_lo, splitk, splitk_bytes, splitk2, splitk2_bytes, total; };

static void vis_act_layout(const VisDims& d, int B, VisAct* A) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = aup(off + bytes); return o; };
    const size_t M = (size_t)B * d.N, D = d.D, HD = d.HD;
    for (int l = 0; l < d.L; ++l) {
        LayerAct& a = A->lay[l];
        a.x_f32 = take(M * D * 4); a.x_hi = take(M * D * 2); a.x_lo = take(M * D * 2);
        a.xm_hi = take((size_t)B * D * 2); a.xm_lo = take((size_t)B * D * 2);
        a.fpre = take(M * HD * 4); a.loc_hi = take(M * HD * 2); a.loc_lo = take(M * HD * 2);
        a.spre = take((size_t)B * HD * 4); a.t_hi = take((size_t)B * HD * 2); a.t_lo = take((size_t)B * HD * 2);
        a.rb = take((size_t)B * HD * 4);
        a.cpre = take(M * HD * 4); a.out_hi = take(M * HD * 2); a.out_lo = take(M * HD * 2);
        a.x1 = take(M * D * 4); a.ln_mean = take(M * 4); a.ln_rstd = take(M * 4);
        a.h_hi = take(M * D * 2); a.h_lo = take(M * D * 2);
        a.upre = take(M * D * 4); a.u_hi = take(M * D * 2); a.u_lo = take(M * D * 2);
    }
    A->xL = take(M * D * 4);
    A->pooled = take((size_t)B * D * 4); A->hmean = take((size_t)B * 4); A->hrstd = take((size_t)B * 4);
    A->z_hi = take((size_t)B * D * 2); A->z_lo = take((size_t)B * D * 2);
    A->splitk_bytes = (size_t)16 * 128 * (size_t)(D > HD ? D : HD) * 4;     // small-M split-K partials / stream-K partial tiles
    if (A->splitk_bytes < ((size_t)160 * 256 * 128 * 4 + 8192)) A->splitk_bytes = (size_t)160 * 256 * 128 * 4 + 8192;
    A->splitk = take(A->splitk_bytes);
    A->splitk2_bytes = (size_t)16 * 128 * (size_t)(D > HD ? D : HD) * 4;    // split-K partials of the side-stream (summary branch) GEMMs
    A->splitk2 = take(A->splitk2_bytes);
    A->total = off;
}

struct VisBwd { size_t dp_hi, dp_lo, splitk, splitk_bytes, dz, dpooled, g2_f32, g2_hi, g2_lo, g1_f32, g1_hi, g1_lo, du_hi, du_lo, dh,
                dc_hi, dc_lo, dlocal, df_hi, df_lo, drb, drb_hi, drb_lo, dt, ds_hi, ds_lo, dxm, part, part_bytes, gsum, splitk2, splitk2_bytes, part2, total; };

constexpr int LN_RPB = 16;    // rows per block in the row-LayerNorm backward

static void vis_bwd_layout(const VisDims& d, int B, VisBwd* S) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = aup(off + bytes); return o; };
    const size_t M = (size_t)B * d.N, D = d.D, HD = d.HD, W = D > HD ? D : HD;
    S->dp_hi = take((size_t)B * d.Gpad * 2); S->dp_lo = take((size_t)B * d.Gpad * 2);
    S->splitk_bytes = (size_t)16 * 128 * W * 4;
    if (S->splitk_bytes < ((size_t)160 * 256 * 128 * 4 + 8192)) S->splitk_bytes = (size_t)160 * 256 * 128 * 4 + 8192;
    S->splitk = take(S->splitk_bytes);
    S->dz = take((size_t)B * D * 4); S->dpooled = take((size_t)B * D * 4);
    S->g2_f32 = take(M * D * 4); S->g2_hi = take(M * D * 2); S->g2_lo = take(M * D * 2);
    S->g1_f32 = take(M * D * 4); S->g1_hi = take(M * D * 2); S->g1_lo = take(M * D * 2);
    S->du_hi = take(M * D * 2); S->du_lo = take(M * D * 2); S->dh = take(M * D * 4);
    S->dc_hi = take(M * HD * 2); S->dc_lo = take(M * HD * 2); S->dlocal = take(M * HD * 4);
    S->df_hi = take(M * HD * 2); S->df_lo = take(M * HD * 2);
    S->drb = take((size_t)B * HD * 4); S->drb_hi = take((size_t)B * HD * 2); S->drb_lo = take((size_t)B * HD * 2);
    S->dt = take((size_t)B * HD * 4); S->ds_hi = take((size_t)B * HD * 2); S->ds_lo = take((size_t)B * HD * 2);
    S->dxm = take((size_t)B * D * 4);
    const size_t nblk = (M + 7) / 8 + 128;
    S->part_bytes = nblk * 2 * W * 4; S->part = take(S->part_bytes);
    S->gsum = take((size_t)B * W * 4);
    S->splitk2_bytes = (size_t)16 * 128 * W * 4; S->splitk2 = take(S->splitk2_bytes);     // side-stream copies (summary branch)
    S->part2 = take((size_t)((B + 7) / 8 + 8) * 2 * W * 4);
    S->total = off;
}

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

// ------------------------------------------------------------------------------------------------ forward
static int vis_forward(const VisDims& d, const VisLayout& P, const float* prm, const bf16* wh, const bf16* wl, const float* x_in, int B,
                       float* pred, uint8_t* act, const VisAct& A, cudaStream_t st) {
    const int M = B * d.N, D = d.D, HD = d.HD, N = d.N;
    void* sk = act + A.splitk;
    for (int l = 0; l < d.L; ++l) {
        const LayerAct& a = A.lay[l];
        const LayerOff& o = P.lay[l];
        float* x = (float*)(act + a.x_f32);
        if (l == 0) {
            prep_input_kernel<<<148 * 8, 256, 0, st>>>(x_in, prm + P.pos, x, (bf16*)(act + a.x_hi), (bf16*)(act + a.x_lo), N, D, (long long)M * D / 4);
        }
        SQ_TRY(check_launch("vis prep"));
        // ---- summary branch on the token mean, on the side stream: GELU(LN64(mean(x) Ws^T + bs)), then the per-slide row
        //      bias Wc[:, 64:] t + bc (per head)                                    tformer_lin.py:21-24
        {
            SideStream* ss = side_stream();
            cudaStream_t s2 = side_fork(ss, st);
            group_mean_kernel<<<dim3((D + 127) / 128, B), 256, 0, s2>>>(x, N, D, 1.0f / (float)N, nullptr, (bf16*)(act + a.xm_hi), (bf16*)(act + a.xm_lo));
            SQ_TRY(GB(B, HD, D).A(act + a.xm_hi, act + a.xm_lo, D).B(wh + o.ws, wl + o.ws, D).bias(prm + o.bs).out_f32((float*)(act + a.spre), HD)
                       .bn(128).auto_split(act + A.splitk2, A.splitk2_bytes).run(s2));
            ln64_fwd_kernel<<<dim3((HD + 127) / 128, (B + 7) / 8), 256, 0, s2>>>((const float*)(act + a.spre), prm + o.lns_g, prm + o.lns_b, B, HD,
                                                                                (bf16*)(act + a.t_hi), (bf16*)(act + a.t_lo));
            SQ_TRY(check_launch("vis ln64 fwd"));
            SQ_TRY(GB(B, HD, 64).A(act + a.t_hi, act + a.t_lo, HD).akoff(64).B(wh + o.wc + 64, wl + o.wc + 64, 128).bn(64).bias(prm + o.bc)
                       .out_f32((float*)(act + a.rb), HD).run(s2));
            // ---- local branch, all heads, on the main stream: GELU(LN64(x Wf^T + bf))            tformer_lin.py:20
            SQ_TRY(GB(M, HD, D).A(act + a.x_hi, act + a.x_lo, D).B(wh + o.wf, wl + o.wf, D).bias(prm + o.bf).ln64(prm + o.lnl_g, prm + o.lnl_b)
                       .save_pre((float*)(act + a.fpre), HD).out_planes(act + a.loc_hi, act + a.loc_lo, HD).sk(sk, A.splitk_bytes).run(st));
            side_join(ss, st);
        }
        // combine: GELU(Wc[:, :64] local + rowbias) (per head)                      tformer_lin.py:24
        SQ_TRY(GB(M, HD, 64).A(act + a.loc_hi, act + a.loc_lo, HD).akoff(64).B(wh + o.wc, wl + o.wc, 128).bn(64)
                   .rowbias((const float*)(act + a.rb), N, HD).act(ACT_GELU).save_pre((float*)(act + a.cpre), HD)
                   .out_planes(act + a.out_hi, act + a.out_lo, HD).run(st));
        // projection + residual                                                     tformer_lin.py:45-46,75
        SQ_TRY(GB(M, D, HD).A(act + a.out_hi, act + a.out_lo, HD).B(wh + o.wp, wl + o.wp, HD).bias(prm + o.bp).res(x, D)
                   .out_f32((float*)(act + a.x1), D).sk(sk, A.splitk_bytes).run(st));
        // feed-forward + residual                                                   tformer_lin.py:54-59,76
        ln_rows_fwd_kernel<<<M, 256, 0, st>>>((const float*)(act + a.x1), prm + o.fg, prm + o.fb, D, 1e-5f, (float*)(act + a.ln_mean),
                                              (float*)(act + a.ln_rstd), (bf16*)(act + a.h_hi), (bf16*)(act + a.h_lo));
        SQ_TRY(check_launch("vis ln"));
        SQ_TRY(GB(M, D, D).A(act + a.h_hi, act + a.h_lo, D).B(wh + o.w1, wl + o.w1, D).bias(prm + o.b1).act(ACT_GELU)
                   .save_pre((float*)(act + a.upre), D).out_planes(act + a.u_hi, act + a.u_lo, D).sk(sk, A.splitk_bytes).run(st));
        const bool last = (l == d.L - 1);
        float* xn = (float*)(act + (last ? A.xL : A.lay[l + 1].x_f32));
        GB g2(M, D, D);
        g2.A(act + a.u_hi, act + a.u_lo, D).B(wh + o.w2, wl + o.w2, D).bias(prm + o.b2).res((const float*)(act + a.x1), D).out_f32(xn, D);
        if (!last) g2.out_planes(act + A.lay[l + 1].x_hi, act + A.lay[l + 1].x_lo, D);
        g2.sk(sk, A.splitk_bytes);
        SQ_TRY(g2.run(st));
    }
    // token mean, head LayerNorm, gene regression head                              tformer_lin.py:103-106
    group_mean_kernel<<<dim3((D + 127) / 128, B), 256, 0, st>>>((const float*)(act + A.xL), N, D, 1.0f / (float)N, (float*)(act + A.pooled), nullptr, nullptr);
    ln_rows_fwd_kernel<<<B, 256, 0, st>>>((const float*)(act + A.pooled), prm + P.hg, prm + P.hb, D, 1e-5f, (float*)(act + A.hmean),
                                          (float*)(act + A.hrstd), (bf16*)(act + A.z_hi), (bf16*)(act + A.z_lo));
    SQ_TRY(check_launch("vis head ln"));
    SQ_TRY(GB(B, d.G, D).A(act + A.z_hi, act + A.z_lo, D).B(wh + P.wh, wl + P.wh, D).bias(prm + P.bh).out_f32(pred, d.G).sk(sk, A.splitk_bytes).run(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ backward
static int vis_backward_head(const VisDims& d, const VisLayout& P, const float* prm, const bf16* wh, const bf16* wl, const float* dpred, int B,
                             uint8_t* act, const VisAct& A, float* grads, uint8_t* sc, const VisBwd& S, cudaStream_t st) {
    const int D = d.D, N = d.N, G = d.G;
    SQ_TRY(split_planes(dpred, (bf16*)(sc + S.dp_hi), (bf16*)(sc + S.dp_lo), B, G, G, d.Gpad, st));
    // dWh = dpred^T z ; dbh = colsum(dpred)
    SQ_TRY(GB(G, D, B).A(sc + S.dp_hi, sc + S.dp_lo, d.Gpad, 1).B(act + A.z_hi, act + A.z_lo, D, 1).out_f32(grads + P.wh, D).sk(sc + S.splitk, S.splitk_bytes).run(st));
    colsum_kernel<<<(G + 31) / 32, 256, 0, st>>>(dpred, B, G, G, 1.0f, grads + P.bh);
    // dz = dpred Wh
    SQ_TRY(GB(B, D, G).A(sc + S.dp_hi, sc + S.dp_lo, d.Gpad).B(wh + P.wh, wl + P.wh, D, 1).out_f32((float*)(sc + S.dz), D)
               .auto_split(sc + S.splitk, S.splitk_bytes).run(st));
    SQ_TRY(launch_ln_rows_bwd((const float*)(sc + S.dz), (const float*)(act + A.pooled), (const float*)(act + A.hmean), (const float*)(act + A.hrstd),
                              prm + P.hg, nullptr, B, D, (float*)(sc + S.dpooled), nullptr, nullptr, (float*)(sc + S.part), grads + P.hg, st));
    const long long total4 = (long long)B * N * D / 4;
    bcast_rows_kernel<<<148 * 8, 256, 0, st>>>((const float*)(sc + S.dpooled), N, D, total4, 1.0f / (float)N, (float*)(sc + S.g2_f32),
                                               (bf16*)(sc + S.g2_hi), (bf16*)(sc + S.g2_lo));
    return check_launch("vis head bwd");
}

static int vis_backward_layer(const VisDims& d, const VisLayout& P, int l, const float* prm, const bf16* wh, const bf16* wl, int B, uint8_t* act,
                              const VisAct& A, float* grads, float* dx_out, uint8_t* sc, const VisBwd& S, cudaStream_t st) {
    const int M = B * d.N, D = d.D, HD = d.HD, N = d.N;
    const LayerAct& a = A.lay[l];
    const LayerOff& o = P.lay[l];
    float* gsum = (float*)(sc + S.gsum);
    float* part = (float*)(sc + S.part);
    // Main stream = the dgrad chain (critical path to the next layer's gradient); side stream = every weight / bias gradient
    // and the summary branch.  Side work is forked right after the tensor it consumes is produced and joined before a
    // buffer it reads is overwritten; both streams run persistent GEMMs, so the side stream's CTAs fill the SMs that the
    // main stream's partially filled waves leave idle (and vice versa).
    SideStream* ss = side_stream();
    cudaStream_t s2 = side_fork(ss, st);                                       // g2 (and its planes) are complete on `st`
    // ---- feed-forward (x2 = W2 GELU(W1 LN(x1) + b1) + b2 + x1)
    SQ_TRY(GB(D, D, M).A(sc + S.g2_hi, sc + S.g2_lo, D, 1).B(act + a.u_hi, act + a.u_lo, D, 1).out_f32(grads + o.w2, D).run(s2));
    SQ_TRY(launch_bias_grad((bf16*)(sc + S.g2_hi), (bf16*)(sc + S.g2_lo), B, N, D, gsum, grads + o.b2, s2));
    SQ_TRY(GB(M, D, D).A(sc + S.g2_hi, sc + S.g2_lo, D).B(wh + o.w2, wl + o.w2, D, 1).dgelu((const float*)(act + a.upre), D)
               .out_planes(sc + S.du_hi, sc + S.du_lo, D).sk(sc + S.splitk, S.splitk_bytes).run(st));
    s2 = side_fork(ss, st);                                                    // dUpre planes
    SQ_TRY(GB(D, D, M).A(sc + S.du_hi, sc + S.du_lo, D, 1).B(act + a.h_hi, act + a.h_lo, D, 1).out_f32(grads + o.w1, D).run(s2));
    SQ_TRY(launch_bias_grad((bf16*)(sc + S.du_hi), (bf16*)(sc + S.du_lo), B, N, D, gsum, grads + o.b1, s2));
    SQ_TRY(GB(M, D, D).A(sc + S.du_hi, sc + S.du_lo, D).B(wh + o.w1, wl + o.w1, D, 1).out_f32((float*)(sc + S.dh), D).sk(sc + S.splitk, S.splitk_bytes).run(st));
    SQ_TRY(launch_ln_rows_bwd((const float*)(sc + S.dh), (const float*)(act + a.x1), (const float*)(act + a.ln_mean), (const float*)(act + a.ln_rstd),
                              prm + o.fg, (const float*)(sc + S.g2_f32), M, D, (float*)(sc + S.g1_f32), (bf16*)(sc + S.g1_hi), (bf16*)(sc + S.g1_lo),
                              part, grads + o.fg, st));
    // ---- mixer (x1 = Wp out + bp + x)
    s2 = side_fork(ss, st);                                                    // g1
    SQ_TRY(GB(D, HD, M).A(sc + S.g1_hi, sc + S.g1_lo, D, 1).B(act + a.out_hi, act + a.out_lo, HD, 1).out_f32(grads + o.wp, HD).run(s2));
    SQ_TRY(launch_bias_grad((bf16*)(sc + S.g1_hi), (bf16*)(sc + S.g1_lo), B, N, D, gsum, grads + o.bp, s2));
    SQ_TRY(GB(M, HD, D).A(sc + S.g1_hi, sc + S.g1_lo, D).B(wh + o.wp, wl + o.wp, HD, 1).dgelu((const float*)(act + a.cpre), HD)
               .out_planes(sc + S.dc_hi, sc + S.dc_lo, HD).sk(sc + S.splitk, S.splitk_bytes).run(st));
    s2 = side_fork(ss, st);                                                    // dCpre planes
    {
        // summary branch: per-slide sums of dCpre (gradient of the row bias and of bc), back through Wc[:, 64:],
        // LayerNorm64+GELU and Ws; ends with dxm, the per-slide term of the input gradient
        float* part2 = (float*)(sc + S.part2);
        group_sum_planes_kernel<<<dim3((HD + 255) / 256, B), 256, 0, s2>>>((bf16*)(sc + S.dc_hi), (bf16*)(sc + S.dc_lo), HD, N, HD, (float*)(sc + S.drb),
                                                                               (bf16*)(sc + S.drb_hi), (bf16*)(sc + S.drb_lo));
        colsum_kernel<<<(HD + 31) / 32, 256, 0, s2>>>((const float*)(sc + S.drb), B, HD, HD, 1.0f, grads + o.bc);
        SQ_TRY(check_launch("vis drb"));
        SQ_TRY(GB(B, HD, 64).A(sc + S.drb_hi, sc + S.drb_lo, HD).akoff(64).B(wh + o.wc + 64, wl + o.wc + 64, 128, 1).bdiag_dgrad(64, HD).bn(64)
                   .out_f32((float*)(sc + S.dt), HD).run(s2));
        SQ_TRY(GB(HD, HD, B).A(sc + S.drb_hi, sc + S.drb_lo, HD, 1).B(act + a.t_hi, act + a.t_lo, HD, 1).diag64().out_f32(grads + o.wc + 64, 128).run(s2));
        SQ_TRY(launch_ln64_bwd((const float*)(sc + S.dt), (const float*)(act + a.spre), prm + o.lns_g, prm + o.lns_b, B, HD, (bf16*)(sc + S.ds_hi),
                               (bf16*)(sc + S.ds_lo), part2, grads + o.lns_g, s2));
        SQ_TRY(GB(HD, D, B).A(sc + S.ds_hi, sc + S.ds_lo, HD, 1).B(act + a.xm_hi, act + a.xm_lo, D, 1).out_f32(grads + o.ws, D).run(s2));
        group_sum_planes_kernel<<<dim3((HD + 255) / 256, 1), 256, 0, s2>>>((bf16*)(sc + S.ds_hi), (bf16*)(sc + S.ds_lo), HD, B, HD, grads + o.bs, nullptr, nullptr);
        SQ_TRY(check_launch("vis dbs"));
        SQ_TRY(GB(B, D, HD).A(sc + S.ds_hi, sc + S.ds_lo, HD).B(wh + o.ws, wl + o.ws, D, 1).alpha(1.0f / (float)N).out_f32((float*)(sc + S.dxm), D)
                   .auto_split(sc + S.splitk2, S.splitk2_bytes).run(s2));
        // dWc[:, :64] = dCpre^T local per head (split-K partials in the side stream's workspace, in order after dxm's)
        SQ_TRY(GB(HD, HD, M).A(sc + S.dc_hi, sc + S.dc_lo, HD, 1).B(act + a.loc_hi, act + a.loc_lo, HD, 1).diag64().out_f32(grads + o.wc, 128)
                   .auto_split(sc + S.splitk2, S.splitk2_bytes).run(s2));
    }
    // local branch on the main stream: dlocal = dCpre Wc[:, :64] per head, back through LayerNorm64+GELU
    SQ_TRY(GB(M, HD, 64).A(sc + S.dc_hi, sc + S.dc_lo, HD).akoff(64).B(wh + o.wc, wl + o.wc, 128, 1).bdiag_dgrad(64, HD).bn(64)
               .out_f32((float*)(sc + S.dlocal), HD).run(st));
    SQ_TRY(launch_ln64_bwd((const float*)(sc + S.dlocal), (const float*)(act + a.fpre), prm + o.lnl_g, prm + o.lnl_b, M, HD, (bf16*)(sc + S.df_hi),
                           (bf16*)(sc + S.df_lo), part, grads + o.lnl_g, st));
    side_join(ss, st);                                                         // dxm is needed now; everything forked so far is done
    s2 = side_fork(ss, st);                                                    // dFpre planes
    SQ_TRY(GB(HD, D, M).A(sc + S.df_hi, sc + S.df_lo, HD, 1).B(act + a.x_hi, act + a.x_lo, D, 1).out_f32(grads + o.wf, D).run(s2));
    SQ_TRY(launch_bias_grad((bf16*)(sc + S.df_hi), (bf16*)(sc + S.df_lo), B, N, HD, gsum, grads + o.bf, s2));
    float* gout = (l == 0 && dx_out) ? dx_out : (float*)(sc + S.g2_f32);
    GB gx(M, D, HD);
    gx.A(sc + S.df_hi, sc + S.df_lo, HD).B(wh + o.wf, wl + o.wf, D, 1).res((const float*)(sc + S.g1_f32), D)
        .rowbias((const float*)(sc + S.dxm), N, D).out_f32(gout, D);
    if (l > 0) gx.out_planes(sc + S.g2_hi, sc + S.g2_lo, D);
    gx.sk(sc + S.splitk, S.splitk_bytes);
    SQ_TRY(gx.run(st));
    if (l == 0) {
        pos_grad_kernel<<<(int)(((long long)N * D / 4 + 255) / 256), 256, 0, st>>>(gout, B, N, D, grads + P.pos);
        SQ_TRY(check_launch("vis pos grad"));
    }
    side_join(ss, st);          // the layer's weight gradients are complete on `st` (all-reduce / AdamW / next layer follow)
    return 0;
}

}  // namespace sq

using namespace sq;

extern "C" {

int sq_vis_param_table_len(const sq_vis_config* cfg) {
    VisDims d; if (vis_dims(cfg, &d)) return -1;
    return 1 + 18 * d.L + 4;
}

int sq_vis_param_layout(const sq_vis_config* cfg, long long* offsets, int n, long long* total_elems) {
    VisDims d; if (vis_dims(cfg, &d)) return -1;
    const int need = 1 + 18 * d.L + 4;
    if (n < need || !offsets) { set_error("vis_param_layout: table too short (%d < %d)", n, need); return -1; }
    VisLayout* L = new VisLayout; vis_layout(d, L);
    int i = 0;
    offsets[i++] = L->pos;
    for (int l = 0; l < d.L; ++l) {
        const LayerOff& o = L->lay[l];
        const long long v[18] = {o.lnl_g, o.lnl_b, o.lns_g, o.lns_b, o.ws, o.bs, o.wf, o.bf, o.wc, o.bc, o.wp, o.bp, o.fg, o.fb, o.w1, o.b1, o.w2, o.b2};
        for (int j = 0; j < 18; ++j) offsets[i++] = v[j];
    }
    offsets[i++] = L->hg; offsets[i++] = L->hb; offsets[i++] = L->wh; offsets[i++] = L->bh;
    if (total_elems) *total_elems = L->total;
    delete L;
    return 0;
}

size_t sq_vis_act_bytes(const sq_vis_config* cfg, int batch) {
    VisDims d; if (vis_dims(cfg, &d) || batch <= 0) return 0;
    VisAct* A = new VisAct; vis_act_layout(d, batch, A);
    const size_t t = A->total; delete A; return t;
}

size_t sq_vis_bwd_bytes(const sq_vis_config* cfg, int batch) {
    VisDims d; if (vis_dims(cfg, &d) || batch <= 0) return 0;
    VisBwd S; vis_bwd_layout(d, batch, &S);
    return S.total;
}

int sq_vis_forward(const sq_vis_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* x, int batch, float* pred,
                   void* act, size_t act_bytes, void* stream) {
    VisDims d; if (vis_dims(cfg, &d)) return -1;
    if (batch <= 0) return 0;
    if (!params || !w_hi || !w_lo || !x || !pred) { set_error("vis_forward: null pointer"); return -1; }
    VisLayout* L = new VisLayout; vis_layout(d, L);
    VisAct* A = new VisAct; vis_act_layout(d, batch, A);
    int rc = -1;
    if (!act || act_bytes < A->total) set_error("vis_forward: activation buffer %zu < %zu", act_bytes, A->total);
    else rc = vis_forward(d, *L, params, (const bf16*)w_hi, (const bf16*)w_lo, x, batch, pred, (uint8_t*)act, *A, (cudaStream_t)stream);
    delete L; delete A;
    return rc;
}

int sq_vis_backward(const sq_vis_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* dpred, int batch, void* act,
                    size_t act_bytes, float* grads, float* dx, void* scratch, size_t scratch_bytes, int stage_hi, int stage_lo, void* stream) {
    VisDims d; if (vis_dims(cfg, &d)) return -1;
    if (batch <= 0) return 0;
    if (!params || !w_hi || !w_lo || !grads) { set_error("vis_backward: null pointer"); return -1; }
    if (stage_hi > d.L || stage_lo < 0 || stage_lo > stage_hi) { set_error("vis_backward: bad stage range [%d, %d]", stage_lo, stage_hi); return -1; }
    VisLayout* L = new VisLayout; vis_layout(d, L);
    VisAct* A = new VisAct; vis_act_layout(d, batch, A);
    VisBwd S; vis_bwd_layout(d, batch, &S);
    int rc = 0;
    if (!act || act_bytes < A->total) { set_error("vis_backward: activation buffer %zu < %zu", act_bytes, A->total); rc = -1; }
    else if (!scratch || scratch_bytes < S.total) { set_error("vis_backward: scratch %zu < %zu", scratch_bytes, S.total); rc = -1; }
    for (int s = stage_hi; rc == 0 && s >= stage_lo; --s) {
        if (s == d.L) {
            if (!dpred) { set_error("vis_backward: null dpred"); rc = -1; break; }
            rc = vis_backward_head(d, *L, params, (const bf16*)w_hi, (const bf16*)w_lo, dpred, batch, (uint8_t*)act, *A, grads, (uint8_t*)scratch, S,
                                   (cudaStream_t)stream);
        } else {
            rc = vis_backward_layer(d, *L, s, params, (const bf16*)w_hi, (const bf16*)w_lo, batch, (uint8_t*)act, *A, grads, dx, (uint8_t*)scratch, S,
                                    (cudaStream_t)stream);
        }
    }
    delete L; delete A;
    return rc;
}

int sq_mse_fwd_bwd(const float* pred, const float* target, int batch, int num_outputs, float* loss, float* dpred, float* scratch, void* stream) {
    if (!pred || !target || !loss || !scratch) { set_error("mse: null pointer"); return -1; }
    const long long n = (long long)batch * num_outputs;
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    mse_kernel<<<MSE_BLOCKS, 256, 0, st>>>(pred, target, n, 2.0f / (float)n, dpred, scratch);
    mse_final_kernel<<<1, 256, 0, st>>>(scratch, MSE_BLOCKS, 1.0f / (float)n, loss);
    return check_launch("mse");
}

int sq_adamw_flat(float* p, const float* g, float* m, float* v, void* p_hi, void* p_lo, long long n, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int step, float grad_scale, void* stream) {
    if (!p || !g || !m || !v) { set_error("adamw: null pointer"); return -1; }
    if (n % 4 != 0 || step < 1) { set_error("adamw: n must be a multiple of 4 and step >= 1"); return -1; }
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const long long n4 = n / 4;
    long long blocks = (n4 + 255) / 256; if (blocks > 148LL * 16) blocks = 148LL * 16;
    adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p_hi, (bf16*)p_lo, n4, 1.0f - lr * weight_decay, beta1, beta2, eps,
                                                                    step_size, bc2_sqrt, grad_scale);
    return check_launch("adamw");
}

}  // extern "C"
