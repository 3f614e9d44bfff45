// ViT softmax-attention aggregator, the reference's `--model_type vit` baseline (src/vit.py: FeedForward :39-49,
// Attention :51-77, Transformer :79-91, ViT :93-116; built at src/main.py:141-143,161-163 with dim_head 64, mlp_dim 2048)
// — forward and hand-written backward behind the same stage contract as the ViS aggregator (vis.cu), so the fused
// MSE / AdamW / data-parallel trainer serve both models.
//
// Linears are split-precision tcgen05 GEMMs (gemm.cuh); LayerNorm, token mean, bias gradients come from aggr.cuh.
// New here is softmax(q k^T / 8) v over the N <= 128 tokens of a slide: 2.56 MFLOP per (slide, head) forward — 1 % of the
// layer's GEMM work — kept in fp32 on the CUDA cores so that the 1e-4 parity bar holds without a split-precision
// formulation of the 100x100 score matrix: one CTA per (slide, head), q/k/v/dO staged in shared memory, scores in shared
// memory, warp-shuffle row reductions, fixed summation order (deterministic).
#include "aggr.cuh"
#include "../../include/sequoia_b200.h"

namespace sq {

constexpr int VIT_MAXL = 64;
constexpr int AT_LD = 68;        // smem row stride (floats) of a 64-wide operand: 16-byte aligned rows, conflict-free LDS.128
constexpr int AT_MAXN = 128;     // tokens per slide supported by the attention kernels (4 score columns per lane)

struct VitDims { int D, L, H, N, G, I, F; long long Gpad; };     // I = heads*64 (inner_dim), F = mlp_dim
struct VitLayerOff { long long ag, ab, wqkv, wo, fg, fb, w1, b1, w2, b2; };
struct VitLayout { long long pos; VitLayerOff lay[VIT_MAXL]; long long hg, hb, wh, bh, total; };

static int vit_dims(const sq_vit_config* c, VitDims* d) {
    if (!c) { set_error("vit: null config"); return -1; }
    if (c->dim <= 0 || c->dim % 64 != 0 || c->dim > 8192) { set_error("vit: dim %d must be a multiple of 64 in (0, 8192]", c->dim); return -1; }
    if (c->mlp_dim <= 0 || c->mlp_dim % 64 != 0 || c->mlp_dim > 8192) { set_error("vit: mlp_dim %d must be a multiple of 64 in (0, 8192]", c->mlp_dim); return -1; }
    if (c->depth <= 0 || c->depth > VIT_MAXL) { set_error("vit: depth %d out of range", c->depth); return -1; }
    if (c->heads <= 0 || c->heads > 128) { set_error("vit: heads %d out of range", c->heads); return -1; }
    if (c->num_clusters <= 0 || c->num_clusters > AT_MAXN) { set_error("vit: num_clusters %d must be in [1, %d]", c->num_clusters, AT_MAXN); return -1; }
    if (c->num_outputs <= 0) { set_error("vit: num_outputs must be positive"); return -1; }
    d->D = c->dim; d->L = c->depth; d->H = c->heads; d->N = c->num_clusters; d->G = c->num_outputs;
    d->I = c->heads * 64; d->F = c->mlp_dim; d->Gpad = (c->num_outputs + 7) / 8 * 8;
    return 0;
}

// Flat parameter layout, 64-element aligned tensors, contiguous per backward stage: pos, layers 0..L-1, head.
static void vit_layout(const VitDims& d, VitLayout* L) {
    long long off = 0;
    auto take = [&](long long n) { long long o = off; off += (n + 63) / 64 * 64; return o; };
    const long long D = d.D, I = d.I, F = d.F;
    L->pos = take((long long)d.N * D);
    for (int l = 0; l < d.L; ++l) {
        VitLayerOff& o = L->lay[l];
        o.ag = take(D); o.ab = take(D);
        o.wqkv = take(3 * I * D); o.wo = take(D * I);
        o.fg = take(D); o.fb = take(D);
        o.w1 = take(F * D); o.b1 = take(F); o.w2 = take(D * F); o.b2 = take(D);
    }
    L->hg = take(D); L->hb = take(D);
    L->wh = take((long long)d.G * D); L->bh = take(d.G);
    L->total = off;
}

struct VitLayerAct { size_t x, mean1, rstd1, h_hi, h_lo, qkv, o_hi, o_lo, x1, mean2, rstd2, h2_hi, h2_lo, upre, u_hi, u_lo; };
struct VitAct { VitLayerAct lay[VIT_MAXL]; size_t xL, pooled, hmean, hrstd, z_hi, z_lo, splitk, splitk_bytes, total; };

static void vit_act_layout(const VitDims& d, int B, VitAct* A) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = aup(off + bytes); return o; };
    const size_t M = (size_t)B * d.N, D = d.D, I = d.I, F = d.F;
    for (int l = 0; l < d.L; ++l) {
        VitLayerAct& a = A->lay[l];
        a.x = take(M * D * 4); a.mean1 = take(M * 4); a.rstd1 = take(M * 4);
        a.h_hi = take(M * D * 2); a.h_lo = take(M * D * 2);
        a.qkv = take(M * 3 * I * 4);
        a.o_hi = take(M * I * 2); a.o_lo = take(M * I * 2);
        a.x1 = take(M * D * 4); a.mean2 = take(M * 4); a.rstd2 = take(M * 4);
        a.h2_hi = take(M * D * 2); a.h2_lo = take(M * D * 2);
        a.upre = take(M * F * 4); a.u_hi = take(M * F * 2); a.u_lo = take(M * F * 2);
    }
    A->xL = take(M * D * 4);
    A->pooled = take((size_t)B * D * 4); A->hmean = take((size_t)B * 4); A->hrstd = take((size_t)B * 4);
    A->z_hi = take((size_t)B * D * 2); A->z_lo = take((size_t)B * D * 2);
    size_t W = D; if (3 * I > W) W = 3 * I; if (F > W) W = F;
    A->splitk_bytes = (size_t)16 * 128 * W * 4;
    if (A->splitk_bytes < ((size_t)160 * 256 * 128 * 4 + 8192)) A->splitk_bytes = (size_t)160 * 256 * 128 * 4 + 8192;
    A->splitk = take(A->splitk_bytes);
    A->total = off;
}

struct VitBwd { size_t dp_hi, dp_lo, splitk, splitk_bytes, dz, dpooled, g2_f32, g2_hi, g2_lo, g1_f32, g1_hi, g1_lo, du_hi, du_lo, dh, dout,
                dqkv_hi, dqkv_lo, part, gsum, total; };

static void vit_bwd_layout(const VitDims& d, int B, VitBwd* S) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = aup(off + bytes); return o; };
    const size_t M = (size_t)B * d.N, D = d.D, I = d.I, F = d.F;
    size_t W = D; if (3 * I > W) W = 3 * I; if (F > W) W = F;
    S->dp_hi = take((size_t)B * d.Gpad * 2); S->dp_lo = take((size_t)B * d.Gpad * 2);
    S->splitk_bytes = (size_t)16 * 128 * W * 4;
    if (S->splitk_bytes < ((size_t)160 * 256 * 128 * 4 + 8192)) S->splitk_bytes = (size_t)160 * 256 * 128 * 4 + 8192;
    S->splitk = take(S->splitk_bytes);
    S->dz = take((size_t)B * D * 4); S->dpooled = take((size_t)B * D * 4);
    S->g2_f32 = take(M * D * 4); S->g2_hi = take(M * D * 2); S->g2_lo = take(M * D * 2);
    S->g1_f32 = take(M * D * 4); S->g1_hi = take(M * D * 2); S->g1_lo = take(M * D * 2);
    S->du_hi = take(M * F * 2); S->du_lo = take(M * F * 2);
    S->dh = take(M * D * 4); S->dout = take(M * I * 4);
    S->dqkv_hi = take(M * 3 * I * 2); S->dqkv_lo = take(M * 3 * I * 2);
    S->part = take(((M + LN_RPB - 1) / LN_RPB + 8) * 2 * D * 4);
    S->gsum = take((size_t)B * W * 4);
    S->total = off;
}

// ------------------------------------------------------------------------------------------------ attention kernels
// 64 floats x N rows, global (row stride ld) -> shared (row stride AT_LD)
__device__ __forceinline__ void at_load(float* dst, const float* __restrict__ src, long long ld, int N) {
    for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
        const int r = i >> 4, c = (i & 15) * 4;
        *reinterpret_cast<float4*>(dst + r * AT_LD + c) = *reinterpret_cast<const float4*>(src + (size_t)r * ld + c);
    }
}

// P[i][j] = softmax_j(scale * q_i . k_j)   (src/vit.py:68-70).  One warp per row, lane owns columns lane + 32k.
__device__ __forceinline__ void at_softmax(const float* Qs, const float* Ks, float* Ps, int N, int pld, float scale) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int i = w; i < N; i += nw) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        const float* q = Qs + i * AT_LD;
#pragma unroll 4
        for (int dd = 0; dd < 64; dd += 4) {
            const float4 qv = *reinterpret_cast<const float4*>(q + dd);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = lane + 32 * k;
                if (j < N) {
                    const float4 kv = *reinterpret_cast<const float4*>(Ks + j * AT_LD + dd);
                    s[k] = fmaf(qv.x, kv.x, s[k]); s[k] = fmaf(qv.y, kv.y, s[k]); s[k] = fmaf(qv.z, kv.z, s[k]); s[k] = fmaf(qv.w, kv.w, s[k]);
                }
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[k] *= scale; if (lane + 32 * k < N) mx = fmaxf(mx, s[k]); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s[k] = (lane + 32 * k < N) ? expf(s[k] - mx) : 0.f; sum += s[k]; }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) if (lane + 32 * k < N) Ps[i * pld + lane + 32 * k] = s[k] * inv;
    }
}

// acc[m] = sum_t P(row, t) X[t][c..c+3] for the rows row = r + AT_RG m this thread owns (r = tid/16, c = 4 (tid%16)).
// TRANS = false: P(row, t) = Ps[row][t];  TRANS = true: P(row, t) = Ps[t][row].
constexpr int AT_THREADS = 512;                 // 16 warps per (slide, head)
constexpr int AT_RG = AT_THREADS / 16;          // row groups
constexpr int AT_NM = AT_MAXN / AT_RG;          // rows per thread
template <bool TRANS>
__device__ __forceinline__ void at_mix(const float* Ps, int pld, const float* Xs, int N, float4 (&acc)[AT_NM]) {
    const int r = threadIdx.x >> 4, c = (threadIdx.x & 15) * 4;
#pragma unroll
    for (int m = 0; m < AT_NM; ++m) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < N; ++t) {
        const float4 x = *reinterpret_cast<const float4*>(Xs + t * AT_LD + c);
#pragma unroll
        for (int m = 0; m < AT_NM; ++m) {
            const int row = r + AT_RG * m;
            if (row < N) {
                const float p = TRANS ? Ps[t * pld + row] : Ps[row * pld + t];
                acc[m].x = fmaf(p, x.x, acc[m].x); acc[m].y = fmaf(p, x.y, acc[m].y);
                acc[m].z = fmaf(p, x.z, acc[m].z); acc[m].w = fmaf(p, x.w, acc[m].w);
            }
        }
    }
}

__device__ __forceinline__ void at_store_planes(const float4 (&acc)[AT_NM], bf16* hi, bf16* lo, size_t row0, long long ld, int col0, int N) {
    const int r = threadIdx.x >> 4, c = (threadIdx.x & 15) * 4;
#pragma unroll
    for (int m = 0; m < AT_NM; ++m) {
        const int row = r + AT_RG * m;
        if (row < N) store_planes4(hi, lo, (row0 + row) * ld + col0 + c, acc[m]);
    }
}

// qkv fp32 [B*N, 3I] (q | k | v column blocks, head h at columns h*64, 'b n (h d) -> b h n d', src/vit.py:65-66)
// -> out planes [B*N, I] ('b h n d -> b n (h d)', :73).  grid (H, B).  Shared memory holds q, ONE k/v buffer and the
// probabilities (95 KB at N = 100: two CTAs per SM); v is prefetched into registers while the scores are computed and
// replaces k afterwards.
__global__ void __launch_bounds__(AT_THREADS, 2) vit_attn_fwd_kernel(const float* __restrict__ qkv, int N, int I, float scale, bf16* __restrict__ oh,
                                                                  bf16* __restrict__ ol) {
    extern __shared__ __align__(16) float at_smem[];
    const int h = blockIdx.x, b = blockIdx.y, pld = N + 1;
    float* Qs = at_smem; float* Ks = Qs + N * AT_LD; float* Ps = Ks + N * AT_LD;
    const size_t row0 = (size_t)b * N;
    const float* base = qkv + row0 * 3 * I + h * 64;
    at_load(Qs, base, 3LL * I, N); at_load(Ks, base + I, 3LL * I, N);
    constexpr int VPT = (AT_MAXN * 16 + AT_THREADS - 1) / AT_THREADS;      // float4 of v per thread
    float4 vreg[VPT];
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int i = threadIdx.x + k * AT_THREADS;
        vreg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < N * 16) vreg[k] = *reinterpret_cast<const float4*>(base + 2 * I + (size_t)(i >> 4) * 3 * I + (i & 15) * 4);
    }
    __syncthreads();
    at_softmax(Qs, Ks, Ps, N, pld, scale);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPT; ++k) {
        const int i = threadIdx.x + k * AT_THREADS;
        if (i < N * 16) *reinterpret_cast<float4*>(Ks + (i >> 4) * AT_LD + (i & 15) * 4) = vreg[k];
    }
    __syncthreads();
    float4 acc[AT_NM];
    at_mix<false>(Ps, pld, Ks, N, acc);
    at_store_planes(acc, oh, ol, row0, I, h * 64, N);
}

// Backward of the above: dqkv planes [B*N, 3I] from qkv (probabilities are recomputed) and dout fp32 [B*N, I].
//   dV = P^T dO;  dP = dO V^T;  dS = P o (dP - rowsum(P o dP)) * scale;  dQ = dS K;  dK = dS^T Q.
__global__ void __launch_bounds__(AT_THREADS) vit_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, int N, int I, float scale,
                                                                  bf16* __restrict__ gh, bf16* __restrict__ gl) {
    extern __shared__ __align__(16) float at_smem[];
    const int h = blockIdx.x, b = blockIdx.y, pld = N + 1;
    float* Qs = at_smem; float* Ks = Qs + N * AT_LD; float* Vs = Ks + N * AT_LD; float* Gs = Vs + N * AT_LD; float* Ps = Gs + N * AT_LD;
    const size_t row0 = (size_t)b * N;
    const float* base = qkv + row0 * 3 * I + h * 64;
    at_load(Qs, base, 3LL * I, N); at_load(Ks, base + I, 3LL * I, N); at_load(Vs, base + 2 * I, 3LL * I, N);
    at_load(Gs, dout + row0 * I + h * 64, I, N);
    __syncthreads();
    at_softmax(Qs, Ks, Ps, N, pld, scale);
    __syncthreads();
    float4 acc[AT_NM];
    at_mix<true>(Ps, pld, Gs, N, acc);                                   // dV[j] = sum_i P[i][j] dO[i]
    at_store_planes(acc, gh, gl, row0, 3LL * I, 2 * I + h * 64, N);
    __syncthreads();
    {   // dS in place of P, one warp per row
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int i = w; i < N; i += nw) {
            float dp[4] = {0.f, 0.f, 0.f, 0.f};
            const float* g = Gs + i * AT_LD;
#pragma unroll 4
            for (int dd = 0; dd < 64; dd += 4) {
                const float4 gv = *reinterpret_cast<const float4*>(g + dd);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int j = lane + 32 * k;
                    if (j < N) {
                        const float4 vv = *reinterpret_cast<const float4*>(Vs + j * AT_LD + dd);
                        dp[k] = fmaf(gv.x, vv.x, dp[k]); dp[k] = fmaf(gv.y, vv.y, dp[k]); dp[k] = fmaf(gv.z, vv.z, dp[k]); dp[k] = fmaf(gv.w, vv.w, dp[k]);
                    }
                }
            }
            float p[4], delta = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) { p[k] = (lane + 32 * k < N) ? Ps[i * pld + lane + 32 * k] : 0.f; delta = fmaf(p[k], dp[k], delta); }
            delta = warp_sum(delta);
#pragma unroll
            for (int k = 0; k < 4; ++k) if (lane + 32 * k < N) Ps[i * pld + lane + 32 * k] = p[k] * (dp[k] - delta) * scale;
        }
    }
    __syncthreads();
    at_mix<false>(Ps, pld, Ks, N, acc);                                  // dQ[i] = sum_j dS[i][j] K[j]
    at_store_planes(acc, gh, gl, row0, 3LL * I, h * 64, N);
    at_mix<true>(Ps, pld, Qs, N, acc);                                   // dK[j] = sum_i dS[i][j] Q[i]
    at_store_planes(acc, gh, gl, row0, 3LL * I, I + h * 64, N);
}

static size_t at_smem_bytes(int N, int operands) { return ((size_t)operands * N * AT_LD + (size_t)N * (N + 1)) * sizeof(float); }

static int launch_attn_fwd(const float* qkv, int B, int N, int H, bf16* oh, bf16* ol, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(vit_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem_bytes(AT_MAXN, 2)) != cudaSuccess) {
            set_error("vit attention: cannot raise the dynamic shared memory limit"); (void)cudaGetLastError(); return -1;
        }
        attr_set = true;
    }
    vit_attn_fwd_kernel<<<dim3(H, B), AT_THREADS, at_smem_bytes(N, 2), st>>>(qkv, N, H * 64, 0.125f, oh, ol);
    return check_launch("vit attention fwd");
}

static int launch_attn_bwd(const float* qkv, const float* dout, int B, int N, int H, bf16* gh, bf16* gl, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(vit_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)at_smem_bytes(AT_MAXN, 4)) != cudaSuccess) {
            set_error("vit attention: cannot raise the dynamic shared memory limit"); (void)cudaGetLastError(); return -1;
        }
        attr_set = true;
    }
    vit_attn_bwd_kernel<<<dim3(H, B), AT_THREADS, at_smem_bytes(N, 4), st>>>(qkv, dout, N, H * 64, 0.125f, gh, gl);
    return check_launch("vit attention bwd");
}

// x_in [B,N,D] + pos [N,D] -> x fp32 (src/vit.py:109)
__global__ void vit_prep_kernel(const float* __restrict__ x_in, const float* __restrict__ pos, float* __restrict__ x, int N, int D, long long total4) {
    const int D4 = D / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / D4; const int c = (int)(i - row * D4) * 4;
        const float4 a = *reinterpret_cast<const float4*>(x_in + i * 4);
        const float4 p = *reinterpret_cast<const float4*>(pos + (size_t)(row % N) * D + c);
        *reinterpret_cast<float4*>(x + i * 4) = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
}

// ------------------------------------------------------------------------------------------------ forward
static int vit_forward(const VitDims& d, const VitLayout& P, const float* prm, const bf16* wh, const bf16* wl, const float* x_in, int B,
                       float* pred, uint8_t* act, const VitAct& A, cudaStream_t st) {
    const int M = B * d.N, D = d.D, I = d.I, F = d.F, N = d.N;
    void* sk = act + A.splitk;
    vit_prep_kernel<<<148 * 8, 256, 0, st>>>(x_in, prm + P.pos, (float*)(act + A.lay[0].x), N, D, (long long)M * D / 4);
    SQ_TRY(check_launch("vit prep"));
    for (int l = 0; l < d.L; ++l) {
        const VitLayerAct& a = A.lay[l];
        const VitLayerOff& o = P.lay[l];
        const float* x = (const float*)(act + a.x);
        // ---- attention: x1 = to_out(softmax(q k^T / 8) v) + x, q|k|v = to_qkv(LN(x))      src/vit.py:63-74,88
        ln_rows_fwd_kernel<<<M, 256, 0, st>>>(x, prm + o.ag, prm + o.ab, D, 1e-5f, (float*)(act + a.mean1), (float*)(act + a.rstd1),
                                              (bf16*)(act + a.h_hi), (bf16*)(act + a.h_lo));
        SQ_TRY(check_launch("vit ln1"));
        SQ_TRY(GB(M, 3 * I, D).A(act + a.h_hi, act + a.h_lo, D).B(wh + o.wqkv, wl + o.wqkv, D).out_f32((float*)(act + a.qkv), 3LL * I)
                   .sk(sk, A.splitk_bytes).run(st));
        SQ_TRY(launch_attn_fwd((const float*)(act + a.qkv), B, N, d.H, (bf16*)(act + a.o_hi), (bf16*)(act + a.o_lo), st));
        SQ_TRY(GB(M, D, I).A(act + a.o_hi, act + a.o_lo, I).B(wh + o.wo, wl + o.wo, I).res(x, D).out_f32((float*)(act + a.x1), D)
                   .sk(sk, A.splitk_bytes).run(st));
        // ---- feed-forward: x2 = W2 GELU(W1 LN(x1) + b1) + b2 + x1                         src/vit.py:41-48,89
        ln_rows_fwd_kernel<<<M, 256, 0, st>>>((const float*)(act + a.x1), prm + o.fg, prm + o.fb, D, 1e-5f, (float*)(act + a.mean2),
                                              (float*)(act + a.rstd2), (bf16*)(act + a.h2_hi), (bf16*)(act + a.h2_lo));
        SQ_TRY(check_launch("vit ln2"));
        SQ_TRY(GB(M, F, D).A(act + a.h2_hi, act + a.h2_lo, D).B(wh + o.w1, wl + o.w1, D).bias(prm + o.b1).act(ACT_GELU)
                   .save_pre((float*)(act + a.upre), F).out_planes(act + a.u_hi, act + a.u_lo, F).sk(sk, A.splitk_bytes).run(st));
        float* xn = (float*)(act + (l == d.L - 1 ? A.xL : A.lay[l + 1].x));
        SQ_TRY(GB(M, D, F).A(act + a.u_hi, act + a.u_lo, F).B(wh + o.w2, wl + o.w2, F).bias(prm + o.b2).res((const float*)(act + a.x1), D)
                   .out_f32(xn, D).sk(sk, A.splitk_bytes).run(st));
    }
    // token mean, head LayerNorm, regression head                                          src/vit.py:112-116
    group_mean_kernel<<<dim3((D + 127) / 128, B), 256, 0, st>>>((const float*)(act + A.xL), N, D, 1.0f / (float)N, (float*)(act + A.pooled), nullptr, nullptr);
    ln_rows_fwd_kernel<<<B, 256, 0, st>>>((const float*)(act + A.pooled), prm + P.hg, prm + P.hb, D, 1e-5f, (float*)(act + A.hmean),
                                          (float*)(act + A.hrstd), (bf16*)(act + A.z_hi), (bf16*)(act + A.z_lo));
    SQ_TRY(check_launch("vit head ln"));
    SQ_TRY(GB(B, d.G, D).A(act + A.z_hi, act + A.z_lo, D).B(wh + P.wh, wl + P.wh, D).bias(prm + P.bh).out_f32(pred, d.G).sk(sk, A.splitk_bytes).run(st));
    return 0;
}

// ------------------------------------------------------------------------------------------------ backward
static int vit_backward_head(const VitDims& d, const VitLayout& P, const float* prm, const bf16* wh, const bf16* wl, const float* dpred, int B,
                             uint8_t* act, const VitAct& A, float* grads, uint8_t* sc, const VitBwd& S, cudaStream_t st) {
    const int D = d.D, N = d.N, G = d.G;
    SQ_TRY(split_planes(dpred, (bf16*)(sc + S.dp_hi), (bf16*)(sc + S.dp_lo), B, G, G, d.Gpad, st));
    SQ_TRY(GB(G, D, B).A(sc + S.dp_hi, sc + S.dp_lo, d.Gpad, 1).B(act + A.z_hi, act + A.z_lo, D, 1).out_f32(grads + P.wh, D).sk(sc + S.splitk, S.splitk_bytes).run(st));
    colsum_kernel<<<(G + 31) / 32, 256, 0, st>>>(dpred, B, G, G, 1.0f, grads + P.bh);
    SQ_TRY(GB(B, D, G).A(sc + S.dp_hi, sc + S.dp_lo, d.Gpad).B(wh + P.wh, wl + P.wh, D, 1).out_f32((float*)(sc + S.dz), D)
               .auto_split(sc + S.splitk, S.splitk_bytes).run(st));
    SQ_TRY(launch_ln_rows_bwd((const float*)(sc + S.dz), (const float*)(act + A.pooled), (const float*)(act + A.hmean), (const float*)(act + A.hrstd),
                              prm + P.hg, nullptr, B, D, (float*)(sc + S.dpooled), nullptr, nullptr, (float*)(sc + S.part), grads + P.hg, st));
    const long long total4 = (long long)B * N * D / 4;
    bcast_rows_kernel<<<148 * 8, 256, 0, st>>>((const float*)(sc + S.dpooled), N, D, total4, 1.0f / (float)N, (float*)(sc + S.g2_f32),
                                               (bf16*)(sc + S.g2_hi), (bf16*)(sc + S.g2_lo));
    return check_launch("vit head bwd");
}

static int vit_backward_layer(const VitDims& d, const VitLayout& P, int l, const float* prm, const bf16* wh, const bf16* wl, int B, uint8_t* act,
                              const VitAct& A, float* grads, float* dx_out, uint8_t* sc, const VitBwd& S, cudaStream_t st) {
    const int M = B * d.N, D = d.D, I = d.I, F = d.F, N = d.N;
    const VitLayerAct& a = A.lay[l];
    const VitLayerOff& o = P.lay[l];
    float* gsum = (float*)(sc + S.gsum);
    float* part = (float*)(sc + S.part);
    // main stream: the dgrad chain; side stream: weight / bias gradients (forked after their inputs exist, joined before a
    // buffer they read is overwritten)
    SideStream* ss = side_stream();
    cudaStream_t s2 = side_fork(ss, st);                                       // g2 = dL/d(layer output), fp32 + planes
    // ---- feed-forward
    SQ_TRY(GB(D, F, M).A(sc + S.g2_hi, sc + S.g2_lo, D, 1).B(act + a.u_hi, act + a.u_lo, F, 1).out_f32(grads + o.w2, F).run(s2));
    SQ_TRY(launch_bias_grad((bf16*)(sc + S.g2_hi), (bf16*)(sc + S.g2_lo), B, N, D, gsum, grads + o.b2, s2));
    SQ_TRY(GB(M, F, D).A(sc + S.g2_hi, sc + S.g2_lo, D).B(wh + o.w2, wl + o.w2, F, 1).dgelu((const float*)(act + a.upre), F)
               .out_planes(sc + S.du_hi, sc + S.du_lo, F).sk(sc + S.splitk, S.splitk_bytes).run(st));
    s2 = side_fork(ss, st);                                                    // dUpre planes
    SQ_TRY(GB(F, D, M).A(sc + S.du_hi, sc + S.du_lo, F, 1).B(act + a.h2_hi, act + a.h2_lo, D, 1).out_f32(grads + o.w1, D).run(s2));
    SQ_TRY(launch_bias_grad((bf16*)(sc + S.du_hi), (bf16*)(sc + S.du_lo), B, N, F, gsum, grads + o.b1, s2));
    SQ_TRY(GB(M, D, F).A(sc + S.du_hi, sc + S.du_lo, F).B(wh + o.w1, wl + o.w1, D, 1).out_f32((float*)(sc + S.dh), D).sk(sc + S.splitk, S.splitk_bytes).run(st));
    SQ_TRY(launch_ln_rows_bwd((const float*)(sc + S.dh), (const float*)(act + a.x1), (const float*)(act + a.mean2), (const float*)(act + a.rstd2),
                              prm + o.fg, (const float*)(sc + S.g2_f32), M, D, (float*)(sc + S.g1_f32), (bf16*)(sc + S.g1_hi), (bf16*)(sc + S.g1_lo),
                              part, grads + o.fg, st));
    // ---- attention
    s2 = side_fork(ss, st);                                                    // g1 = dL/dx1
    SQ_TRY(GB(D, I, M).A(sc + S.g1_hi, sc + S.g1_lo, D, 1).B(act + a.o_hi, act + a.o_lo, I, 1).out_f32(grads + o.wo, I).run(s2));
    SQ_TRY(GB(M, I, D).A(sc + S.g1_hi, sc + S.g1_lo, D).B(wh + o.wo, wl + o.wo, I, 1).out_f32((float*)(sc + S.dout), I).sk(sc + S.splitk, S.splitk_bytes).run(st));
    SQ_TRY(launch_attn_bwd((const float*)(act + a.qkv), (const float*)(sc + S.dout), B, N, d.H, (bf16*)(sc + S.dqkv_hi), (bf16*)(sc + S.dqkv_lo), st));
    s2 = side_fork(ss, st);                                                    // dqkv planes
    SQ_TRY(GB(3 * I, D, M).A(sc + S.dqkv_hi, sc + S.dqkv_lo, 3LL * I, 1).B(act + a.h_hi, act + a.h_lo, D, 1).out_f32(grads + o.wqkv, D).run(s2));
    SQ_TRY(GB(M, D, 3 * I).A(sc + S.dqkv_hi, sc + S.dqkv_lo, 3LL * I).B(wh + o.wqkv, wl + o.wqkv, D, 1).out_f32((float*)(sc + S.dh), D)
               .sk(sc + S.splitk, S.splitk_bytes).run(st));
    side_join(ss, st);                                                         // g2 planes are overwritten next
    float* gout = (l == 0 && dx_out) ? dx_out : (float*)(sc + S.g2_f32);
    SQ_TRY(launch_ln_rows_bwd((const float*)(sc + S.dh), (const float*)(act + a.x), (const float*)(act + a.mean1), (const float*)(act + a.rstd1),
                              prm + o.ag, (const float*)(sc + S.g1_f32), M, D, gout, l > 0 ? (bf16*)(sc + S.g2_hi) : nullptr,
                              l > 0 ? (bf16*)(sc + S.g2_lo) : nullptr, part, grads + o.ag, st));
    if (l == 0) {
        pos_grad_kernel<<<(int)(((long long)N * D / 4 + 255) / 256), 256, 0, st>>>(gout, B, N, D, grads + P.pos);
        SQ_TRY(check_launch("vit pos grad"));
    }
    return 0;
}

}  // namespace sq

using namespace sq;

extern "C" {

int sq_vit_param_table_len(const sq_vit_config* cfg) {
    VitDims d; if (vit_dims(cfg, &d)) return -1;
    return 1 + 10 * d.L + 4;
}

int sq_vit_param_layout(const sq_vit_config* cfg, long long* offsets, int n, long long* total_elems) {
    VitDims d; if (vit_dims(cfg, &d)) return -1;
    const int need = 1 + 10 * d.L + 4;
    if (n < need || !offsets) { set_error("vit_param_layout: table too short (%d < %d)", n, need); return -1; }
    VitLayout* L = new VitLayout; vit_layout(d, L);
    int i = 0;
    offsets[i++] = L->pos;
    for (int l = 0; l < d.L; ++l) {
        const VitLayerOff& o = L->lay[l];
        const long long v[10] = {o.ag, o.ab, o.wqkv, o.wo, o.fg, o.fb, o.w1, o.b1, o.w2, o.b2};
        for (int j = 0; j < 10; ++j) offsets[i++] = v[j];
    }
    offsets[i++] = L->hg; offsets[i++] = L->hb; offsets[i++] = L->wh; offsets[i++] = L->bh;
    if (total_elems) *total_elems = L->total;
    delete L;
    return 0;
}

size_t sq_vit_act_bytes(const sq_vit_config* cfg, int batch) {
    VitDims d; if (vit_dims(cfg, &d) || batch <= 0) return 0;
    VitAct* A = new VitAct; vit_act_layout(d, batch, A);
    const size_t t = A->total; delete A; return t;
}

size_t sq_vit_bwd_bytes(const sq_vit_config* cfg, int batch) {
    VitDims d; if (vit_dims(cfg, &d) || batch <= 0) return 0;
    VitBwd S; vit_bwd_layout(d, batch, &S);
    return S.total;
}

int sq_vit_forward(const sq_vit_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* x, int batch, float* pred,
                   void* act, size_t act_bytes, void* stream) {
    VitDims d; if (vit_dims(cfg, &d)) return -1;
    if (batch <= 0) return 0;
    if (!params || !w_hi || !w_lo || !x || !pred) { set_error("vit_forward: null pointer"); return -1; }
    VitLayout* L = new VitLayout; vit_layout(d, L);
    VitAct* A = new VitAct; vit_act_layout(d, batch, A);
    int rc = -1;
    if (!act || act_bytes < A->total) set_error("vit_forward: activation buffer %zu < %zu", act_bytes, A->total);
    else rc = vit_forward(d, *L, params, (const bf16*)w_hi, (const bf16*)w_lo, x, batch, pred, (uint8_t*)act, *A, (cudaStream_t)stream);
    delete L; delete A;
    return rc;
}

int sq_vit_backward(const sq_vit_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* dpred, int batch, void* act,
                    size_t act_bytes, float* grads, float* dx, void* scratch, size_t scratch_bytes, int stage_hi, int stage_lo, void* stream) {
    VitDims d; if (vit_dims(cfg, &d)) return -1;
    if (batch <= 0) return 0;
    if (!params || !w_hi || !w_lo || !grads) { set_error("vit_backward: null pointer"); return -1; }
    if (stage_hi > d.L || stage_lo < 0 || stage_lo > stage_hi) { set_error("vit_backward: bad stage range [%d, %d]", stage_lo, stage_hi); return -1; }
    VitLayout* L = new VitLayout; vit_layout(d, L);
    VitAct* A = new VitAct; vit_act_layout(d, batch, A);
    VitBwd S; vit_bwd_layout(d, batch, &S);
    int rc = 0;
    if (!act || act_bytes < A->total) { set_error("vit_backward: activation buffer %zu < %zu", act_bytes, A->total); rc = -1; }
    else if (!scratch || scratch_bytes < S.total) { set_error("vit_backward: scratch %zu < %zu", scratch_bytes, S.total); rc = -1; }
    for (int s = stage_hi; rc == 0 && s >= stage_lo; --s) {
        if (s == d.L) {
            if (!dpred) { set_error("vit_backward: null dpred"); rc = -1; break; }
            rc = vit_backward_head(d, *L, params, (const bf16*)w_hi, (const bf16*)w_lo, dpred, batch, (uint8_t*)act, *A, grads, (uint8_t*)scratch, S,
                                   (cudaStream_t)stream);
        } else {
            rc = vit_backward_layer(d, *L, s, params, (const bf16*)w_hi, (const bf16*)w_lo, batch, (uint8_t*)act, *A, grads, dx, (uint8_t*)scratch, S,
                                    (cudaStream_t)stream);
        }
    }
    delete L; delete A;
    return rc;
}

}  // extern "C"
