// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   C[M,N] = sum_terms A_t[M,K] * B_t[N,K]^T        (bf16 operands, fp32 accumulation in TMEM)
//
// * operands are staged by TMA into 128B-swizzled shared memory, either K-major ([MN][K] row-major)
//   or MN-major ([K][MN] row-major) so dgrad/wgrad need no explicit transposes;
// * nterms == 3 is the split-precision mode (A_hi*B_hi + A_hi*B_lo + A_lo*B_hi) that keeps fp32
//   parity on bf16 tensor cores;
// * conv mode loads A as a 4-D NHWC box per filter tap (TMA zero-fills the padding halo);
// * the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1;
// * the epilogue is fused: bias, per-slide row bias, residual, ReLU / exact GELU / per-head LayerNorm(64)+GELU /
//   GELU', and writes any of fp32, bf16-hi, bf16-lo planes.
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4..7 = epilogue (TMEM lane quadrant = warp % 4).
#pragma once
#include "ptx.cuh"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace sq {

typedef __nv_bfloat16 bf16;

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_LN64_GELU = 3, ACT_MUL_DGELU = 4 };

struct EpiParams {
    float* out_f32;     long long ld_f32;
    bf16* out_hi;       bf16* out_lo;  long long ld_bf;
    const float* bias;                       // [N], may be null
    const float* rowbias; int rowbias_div; long long ld_rowbias;   // rowbias[(row / div) * ld + col]
    const float* res_f32; const bf16* res_bf; long long ld_res;    // added before the activation
    float* save_pre;    long long ld_pre;    // value just before the activation
    const float* aux;   long long ld_aux;    // ACT_MUL_DGELU: out = v * gelu'(aux)
    const float* ln_gamma; const float* ln_beta;  // ACT_LN64_GELU: per-column affine, groups of 64 columns
    int act;
    float alpha;                             // scales the accumulator first
};

struct GemmKParams {
    int M, N;
    int num_m, num_n;
    int nk;        // 64-wide k-blocks per term
    int nterms;    // 1 or 3
    int split_k;   // >= 1
    int kb_per_split;
    int a_mn, b_mn;
    int a_koff_per_ntile;   // block-diagonal mode: extra A k-offset per n-tile
    int b_koff_per_ntile;   // block-diagonal dgrad: extra B k-offset per n-tile ...
    int b_nadj_per_ntile;   // ... and an adjustment of B's n coordinate per n-tile
    int diag64;             // block-diagonal wgrad: only the 64x64 diagonal blocks of C are produced (BN = 64)
    // conv mode
    int conv, cblocks, S, stride, pad, tiles_per_img, BH, BIMG;
    float* partial;         // split-K workspace [split][M][N]; stream-K: per-CTA partial tiles [cta][BN][128]
    int streamk;            // 1: the last, partially filled wave of tiles is split evenly over all CTAs (see WorkIter)
    int sk_t0;              // first tile scheduled stream-K (tiles below it are whole waves)
    int* sk_flags;          // stream-K: [cta][8] "partial published" flags, zeroed before the launch
    unsigned long long* prof;   // optional per-CTA cycle counters (sq_gemm_profile): [cta][16]
    EpiParams e;
};

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dgelu_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Epilogue classes: the kernel is instantiated per class so that each instantiation only carries the code of the
// options its class can use (one generic instantiation was ~15k SASS instructions and instruction-fetch bound on
// small-K launches).  Within a class the options are still runtime switches.
enum EpiClass {
    EPI_GENERIC = 0,   // everything (tests, rare combinations)
    EPI_CONV = 1,      // bias, optional bf16 residual, optional ReLU -> bf16 plane and/or fp32          (ResNet convolutions)
    EPI_F32 = 2,       // alpha, bias, row bias, fp32 residual -> fp32 and/or planes, no activation     (projections, dgrad, wgrad)
    EPI_GELU = 3,      // bias, row bias, save_pre, exact GELU -> planes / fp32                         (W1, combine)
    EPI_DGELU = 4,     // multiply by GELU'(aux) -> planes / fp32                                       (dgrad through a GELU)
    EPI_LN64 = 5,      // bias, save_pre, per-head LayerNorm(64) + GELU -> planes / fp32                (local branch)
    EPI_CONV_PF = 6,   // EPI_CONV with 32-column chunks and the residual tile prefetched one chunk ahead (opt-in, SQ_CONV_EPI_PF=1)
    EPI_NUM_CLASSES = 7
};
__host__ __device__ constexpr bool epi_has(int cls, int opt) {
    // opt: 0 alpha, 1 bias, 2 rowbias, 3 res_f32, 4 res_bf, 5 save_pre, 6 relu, 7 gelu, 8 dgelu, 9 ln64, 10 out_f32, 11 out_bf, 12 out_lo
    return cls == EPI_GENERIC ? true
         : (cls == EPI_CONV || cls == EPI_CONV_PF) ? (opt == 1 || opt == 4 || opt == 6 || opt == 7 || opt == 10 || opt == 11)
         : cls == EPI_F32 ? (opt == 0 || opt == 1 || opt == 2 || opt == 3 || opt == 10 || opt == 11 || opt == 12)
         : cls == EPI_GELU ? (opt == 1 || opt == 2 || opt == 5 || opt == 7 || opt == 10 || opt == 11 || opt == 12)
         : cls == EPI_DGELU ? (opt == 0 || opt == 8 || opt == 10 || opt == 11 || opt == 12)
         : cls == EPI_LN64 ? (opt == 1 || opt == 5 || opt == 9 || opt == 10 || opt == 11 || opt == 12)
         : false;
}

// Applies the fused epilogue to NC consecutive columns of one row and stores them.
template <int NC, int CLS = EPI_GENERIC>
__device__ __forceinline__ void epilogue_apply(float* v, long long row, int col0, int N, const EpiParams& e0) {
    // options outside the class are compiled out by nulling them in a local copy the optimiser can see through
    EpiParams e = e0;
    if constexpr (!epi_has(CLS, 0)) e.alpha = 1.0f;
    if constexpr (!epi_has(CLS, 1)) e.bias = nullptr;
    if constexpr (!epi_has(CLS, 2)) e.rowbias = nullptr;
    if constexpr (!epi_has(CLS, 3)) e.res_f32 = nullptr;
    if constexpr (!epi_has(CLS, 4)) e.res_bf = nullptr;
    if constexpr (!epi_has(CLS, 5)) e.save_pre = nullptr;
    if constexpr (!epi_has(CLS, 10)) e.out_f32 = nullptr;
    if constexpr (!epi_has(CLS, 11)) e.out_hi = nullptr;
    if constexpr (!epi_has(CLS, 12)) e.out_lo = nullptr;
    if constexpr (CLS != EPI_GENERIC) {
        if constexpr (CLS == EPI_CONV || CLS == EPI_CONV_PF) e.act = (e.act == ACT_RELU || e.act == ACT_GELU) ? e.act : ACT_NONE;
        else if constexpr (CLS == EPI_F32) e.act = ACT_NONE;
        else if constexpr (CLS == EPI_GELU) e.act = ACT_GELU;
        else if constexpr (CLS == EPI_DGELU) e.act = ACT_MUL_DGELU;
        else if constexpr (CLS == EPI_LN64) e.act = ACT_LN64_GELU;
    }
    const int nvalid = min(NC, N - col0);
    if (e.alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] *= e.alpha;
    }
    if (e.bias) {
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (i < nvalid) v[i] += __ldg(e.bias + col0 + i);
    }
    if (e.rowbias) {
        const float* rb = e.rowbias + (row / e.rowbias_div) * e.ld_rowbias + col0;
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (i < nvalid) v[i] += __ldg(rb + i);
    }
    if (e.res_f32) {
        const float* r = e.res_f32 + row * e.ld_res + col0;
        if (nvalid == NC && (e.ld_res & 3) == 0) {
#pragma unroll
            for (int i = 0; i < NC; i += 4) {
                const float4 t = *reinterpret_cast<const float4*>(r + i);
                v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < nvalid) v[i] += r[i];
        }
    }
    if (e.res_bf) {
        const bf16* r = e.res_bf + row * e.ld_res + col0;
        if (nvalid == NC && (e.ld_res & 7) == 0) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
                const uint4 t = *reinterpret_cast<const uint4*>(r + i);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h[j]);
                    v[i + 2 * j] += f.x; v[i + 2 * j + 1] += f.y;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < nvalid) v[i] += __bfloat162float(r[i]);
        }
    }
    if (e.save_pre) {
        float* s = e.save_pre + row * e.ld_pre + col0;
        if (nvalid == NC && (e.ld_pre & 3) == 0) {
#pragma unroll
            for (int i = 0; i < NC; i += 4) *reinterpret_cast<float4*>(s + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < nvalid) s[i] = v[i];
        }
    }
    if (e.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (e.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = gelu_f(v[i]);
    } else if (e.act == ACT_MUL_DGELU) {
        const float* a = e.aux + row * e.ld_aux + col0;
#pragma unroll
        for (int i = 0; i < NC; ++i)
            if (i < nvalid) v[i] *= dgelu_f(a[i]);
    } else if (e.act == ACT_LN64_GELU) {
        // per-head LayerNorm over groups of 64 columns (eps 1e-5, biased variance), then exact GELU
        if constexpr (NC % 64 == 0) {
#pragma unroll
            for (int g = 0; g < NC; g += 64) {
                float mean = 0.f;
#pragma unroll
                for (int i = 0; i < 64; ++i) mean += v[g + i];
                mean *= (1.0f / 64.0f);
                float var = 0.f;
#pragma unroll
                for (int i = 0; i < 64; ++i) { const float d = v[g + i] - mean; var += d * d; }
                const float rstd = rsqrtf(var * (1.0f / 64.0f) + 1e-5f);
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const float xh = (v[g + i] - mean) * rstd;
                    v[g + i] = gelu_f(xh * __ldg(e.ln_gamma + col0 + g + i) + __ldg(e.ln_beta + col0 + g + i));
                }
            }
        }
    }
    if (e.out_f32) {
        float* o = e.out_f32 + row * e.ld_f32 + col0;
        if (nvalid == NC && (e.ld_f32 & 3) == 0) {
#pragma unroll
            for (int i = 0; i < NC; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < nvalid) o[i] = v[i];
        }
    }
    if (e.out_hi) {
        bf16* oh = e.out_hi + row * e.ld_bf + col0;
        bf16* ol = e.out_lo ? e.out_lo + row * e.ld_bf + col0 : nullptr;
        if (nvalid == NC && (e.ld_bf & 7) == 0) {
#pragma unroll
            for (int i = 0; i < NC; i += 8) {
                uint4 ph, pl;
                __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&ph);
                __nv_bfloat162* ll = reinterpret_cast<__nv_bfloat162*>(&pl);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float a = v[i + 2 * j], b = v[i + 2 * j + 1];
                    const bf16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
                    hh[j] = __halves2bfloat162(ha, hb);
                    ll[j] = __halves2bfloat162(__float2bfloat16_rn(a - __bfloat162float(ha)),
                                               __float2bfloat16_rn(b - __bfloat162float(hb)));
                }
                *reinterpret_cast<uint4*>(oh + i) = ph;
                if (ol) *reinterpret_cast<uint4*>(ol + i) = pl;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i < nvalid) {
                    const bf16 h = __float2bfloat16_rn(v[i]);
                    oh[i] = h;
                    if (ol) ol[i] = __float2bfloat16_rn(v[i] - __bfloat162float(h));
                }
        }
    }
}

// ResNet convolution epilogue with COALESCED global accesses: one warp owns 32 accumulator rows x NC columns (lane = row
// after tcgen05.ld).  Row-per-lane 16-byte accesses touch 32 different 128-byte lines per request and are limited by L1
// tag throughput (measured: ~100 cycles per element on the 64->256 1x1 convolutions), so the residual tile is read and
// the output tile written in a lines-per-request pattern (32/SEGS rows x SEGS 16-byte segments) and transposed through a
// per-warp shared-memory staging buffer (32 rows x 144 B, conflict-free for both patterns).
constexpr int GEMM_STG_LD = 144;
constexpr int GEMM_STG_BYTES = 32 * GEMM_STG_LD;
template <int NC>
__device__ __forceinline__ void epilogue_conv_staged(float* v, long long row_base, int lane, int M, int col0, const EpiParams& e, uint8_t* stg) {
    constexpr int SEGS = NC / 8;            // 16-byte segments per row of NC bf16
    constexpr int RPR = 32 / SEGS;          // rows per warp-wide request
    constexpr int NREQ = 32 / RPR;
    const int r_sub = lane / SEGS, seg = lane % SEGS;
    uint4 rr[NREQ];
    if (e.res_bf) {
#pragma unroll
        for (int i = 0; i < NREQ; ++i) {
            const long long r = row_base + i * RPR + r_sub;
            rr[i] = make_uint4(0, 0, 0, 0);
            if (r < M) rr[i] = *reinterpret_cast<const uint4*>(e.res_bf + r * e.ld_res + col0 + seg * 8);
        }
    }
    if (e.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) {
            const float4 b = __ldg(b4 + i);
            v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
    }
    if (e.res_bf) {
#pragma unroll
        for (int i = 0; i < NREQ; ++i) *reinterpret_cast<uint4*>(stg + (i * RPR + r_sub) * GEMM_STG_LD + seg * 16) = rr[i];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < SEGS; ++j) {
            const uint4 t = *reinterpret_cast<const uint4*>(stg + lane * GEMM_STG_LD + j * 16);
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[8 * j + 2 * k] += __uint_as_float(w[k] << 16);
                v[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
            }
        }
        __syncwarp();
    }
    if (e.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (e.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] = gelu_f(v[i]);
    }
    if (e.out_hi) {
#pragma unroll
        for (int j = 0; j < SEGS; ++j) {
            uint4 pk;
            __nv_bfloat162 t;
            t = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
            *reinterpret_cast<uint4*>(stg + lane * GEMM_STG_LD + j * 16) = pk;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NREQ; ++i) {
            const long long r = row_base + i * RPR + r_sub;
            const uint4 t = *reinterpret_cast<const uint4*>(stg + (i * RPR + r_sub) * GEMM_STG_LD + seg * 16);
            if (r < M) *reinterpret_cast<uint4*>(e.out_hi + r * e.ld_bf + col0 + seg * 8) = t;
        }
        __syncwarp();
    }
    if (e.out_f32) {
        const long long r = row_base + lane;
        if (r < M) {
            float4* o4 = reinterpret_cast<float4*>(e.out_f32 + r * e.ld_f32 + col0);
#pragma unroll
            for (int i = 0; i < NC / 4; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    }
}

// ---- EPI_CONV_PF (opt-in, NOT yet validated on hardware: written at the end of round 1 after the GPU budget was spent).
// The EPI_CONV kernels sit at the 168-register cap with spills (64 accumulator values + 8 residual vectors per lane), which is
// why the residual loads cannot be hoisted there.  This variant walks the tile in 32-column chunks (32 values + 4 residual
// vectors per lane), so the residual of chunk c+1 is requested before chunk c is processed and the residual of a tile's first
// chunk before the warp waits for the accumulator — the K <= 512 1x1 expansions are bound by exactly that latency
// (profiles/r01_resnet_per_conv_efficiency.txt).  Staging: 32 rows x 64 B per warp, 16-byte chunk s of row r stored at
// chunk (s ^ ((r >> 1) & 3)): conflict-free for the row-per-lane and for the 8-rows-x-4-segments access patterns.
constexpr int GEMM_PF_STG_BYTES = 32 * 64;
__device__ __forceinline__ int pf_stg_off(int r, int s) { return r * 64 + ((s ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ void conv_pf_load_res(uint4 (&rr)[4], const EpiParams& e, long long row_base, int lane, int M, int col0) {
    const int r_sub = lane >> 2, seg = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = row_base + i * 8 + r_sub;
        rr[i] = make_uint4(0u, 0u, 0u, 0u);
        if (r < M) rr[i] = *reinterpret_cast<const uint4*>(e.res_bf + r * e.ld_res + col0 + seg * 8);
    }
}

// one 32-row x 32-column chunk: v = accumulator values of this lane's row, rr = the residual chunk in the coalesced pattern
__device__ __forceinline__ void conv_pf_chunk(float (&v)[32], const uint4 (&rr)[4], bool has_res, long long row_base, int lane, int M,
                                              int col0, const EpiParams& e, uint8_t* stg) {
    const int r_sub = lane >> 2, seg = lane & 3;
    if (e.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(e.bias + col0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
    }
    if (has_res) {
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(stg + pf_stg_off(i * 8 + r_sub, seg)) = rr[i];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 t = *reinterpret_cast<const uint4*>(stg + pf_stg_off(lane, j));
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[8 * j + 2 * k] += __uint_as_float(w[k] << 16);
                v[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
            }
        }
        __syncwarp();
    }
    if (e.act == ACT_RELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
    } else if (e.act == ACT_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_f(v[i]);
    }
    if (e.out_hi) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 pk;
            __nv_bfloat162 t;
            t = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
            *reinterpret_cast<uint4*>(stg + pf_stg_off(lane, j)) = pk;
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const long long r = row_base + i * 8 + r_sub;
            const uint4 t = *reinterpret_cast<const uint4*>(stg + pf_stg_off(i * 8 + r_sub, seg));
            if (r < M) *reinterpret_cast<uint4*>(e.out_hi + r * e.ld_bf + col0 + seg * 8) = t;
        }
        __syncwarp();
    }
    if (e.out_f32) {
        const long long r = row_base + lane;
        if (r < M) {
            float4* o4 = reinterpret_cast<float4*>(e.out_f32 + r * e.ld_f32 + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) o4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }
    }
}

constexpr int GEMM_THREADS = 384;      // warps 0-3: TMA producer, MMA issuer, TMEM allocator, idle; warps 4-11: epilogue
constexpr int GEMM_EPI_WARPS = 8;      // two warps per TMEM lane quadrant, each draining half of the tile's columns
constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB

// FUSE3: split-precision launches stage A_hi, A_lo, B_hi, B_lo of a k-block ONCE and issue the three MMA groups
// (hi*hi, hi*lo, lo*hi) from that stage: 4 tile loads per k-block instead of 6.  The big ViS GEMMs are bound by
// L2 -> shared-memory operand traffic (~14 TB/s measured), not by the tensor pipe, so this is worth ~1/3 of their time.
template <int BN, int CLS = 0, int FUSE3 = 0> struct GemmCfg {
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = (FUSE3 ? 2 : 1) * (GEMM_A_BYTES + B_BYTES);
    static constexpr int STAGING = (CLS == EPI_CONV) ? GEMM_EPI_WARPS * GEMM_STG_BYTES
                                 : (CLS == EPI_CONV_PF) ? GEMM_EPI_WARPS * GEMM_PF_STG_BYTES : 0;     // per-warp epilogue transposition buffers
    static constexpr int STAGES = FUSE3 ? (BN >= 192 ? 2 : (BN == 128 ? 3 : 4))
                                        : (BN == 256) ? (STAGING ? 3 : 4) : (BN == 128 ? (STAGING ? 5 : 6) : (STAGING ? 6 : 8));
    static_assert(STAGES * STAGE_BYTES + 1024 + 256 + STAGING <= 232448, "shared memory budget");
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + STAGING;
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
};

// Work assignment shared by the three warp roles.
//  classic : tile t = blockIdx.x + i * gridDim.x over [split][m][n] tiles, k-blocks of the tile's split.
//  stream-K: whole waves of tiles (t < sk_t0) are processed classic; the CTAs then split the tiles x k-blocks iteration
//            space of the LAST, partially filled wave evenly.  A CTA walks its share in DESCENDING order as segments
//            (tile, kb0, kb1).  A segment that does not reach its tile's last k-block is a partial: its accumulator is
//            published to the workspace (it is always the first stream-K segment of the CTA, so it is published early);
//            the CTA that owns the tile's last k-block adds the partials (fixed order) and runs the epilogue.
struct WorkIter {
    int mode, t, step, total_tiles, tiles_mn, total_kb, kb_per_split, t0;
    long long s, cur_end;
    __device__ __forceinline__ WorkIter(int streamk, int sk_t0, int tiles_mn_, int total_tiles_, int total_kb_, int kb_per_split_) {
        mode = streamk; tiles_mn = tiles_mn_; total_tiles = total_tiles_; total_kb = total_kb_; kb_per_split = kb_per_split_;
        t = blockIdx.x; step = gridDim.x; s = 0; cur_end = 0; t0 = sk_t0;
        if (mode) {
            const long long total = (long long)(tiles_mn_ - sk_t0) * total_kb_;
            s = total * blockIdx.x / gridDim.x; cur_end = total * (blockIdx.x + 1) / gridDim.x;
        }
    }
    __device__ __forceinline__ bool next(int& tile, int& kb0, int& kb1, bool& first, bool& last) {
        if (!mode || t < t0) {
            if (t >= total_tiles) return false;
            const int split = t / tiles_mn;
            tile = t; kb0 = split * kb_per_split; kb1 = min(kb0 + kb_per_split, total_kb); first = true; last = true;
            t += step;
            return true;
        }
        if (cur_end <= s) return false;
        const long long tt = (cur_end - 1) / total_kb, tb = tt * total_kb;
        const long long sb = s > tb ? s : tb;
        tile = t0 + (int)tt; kb0 = (int)(sb - tb); kb1 = (int)(cur_end - tb); first = kb0 == 0; last = kb1 == total_kb;
        cur_end = sb;
        return true;
    }
};
__device__ __forceinline__ long long streamk_range_start(int cta, int sk_tiles, int total_kb, int grid) {
    return (long long)sk_tiles * total_kb * cta / grid;
}

// adds the partial accumulators the `nprev` preceding CTAs published for this tile (fixed order; L1 bypassed)
template <int NC>
__device__ __forceinline__ void streamk_add(float* v, const float* partial, int nprev, int bn, int c, int r_in) {
    for (int j = 1; j <= nprev; ++j) {
        const float* ws = partial + ((size_t)(blockIdx.x - j) * bn + c) * 128 + r_in;
#pragma unroll
        for (int i = 0; i < NC; ++i) v[i] += __ldcg(ws + (size_t)i * 128);
    }
}

template <int BN, int CLS, int FUSE3>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
               const __grid_constant__ CUtensorMap mapB0, const __grid_constant__ CUtensorMap mapB1,
               const GemmKParams p) {
    using Cfg = GemmCfg<BN, CLS, FUSE3>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // stage layout: [A (hi)][A lo (FUSE3)][B (hi)][B lo (FUSE3)]
    constexpr int OFF_ALO = GEMM_A_BYTES, OFF_B = (FUSE3 ? 2 : 1) * GEMM_A_BYTES, OFF_BLO = OFF_B + Cfg::B_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // Programmatic dependent launch: let the next kernel of the stream become resident as soon as SMs free up, and run this
    // kernel's own prologue (barrier init, TMEM allocation, descriptor prefetch) while the previous kernel is still draining.
    asm volatile("griddepcontrol.launch_dependents;");
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA0); tma_prefetch_desc(&mapA1);
        tma_prefetch_desc(&mapB0); tma_prefetch_desc(&mapB1);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], GEMM_EPI_WARPS); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // everything below reads or writes global memory produced by earlier kernels of the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int tiles_mn = p.num_m * p.num_n;
    const int total_tiles = tiles_mn * p.split_k;
    const int total_kb = FUSE3 ? p.nk : p.nk * p.nterms;

    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer =====================
            int stage = 0; uint32_t phase = 0;
            long long pw = 0; const long long pt0 = clock64();
            WorkIter wi(p.streamk, p.sk_t0, tiles_mn, total_tiles, total_kb, p.kb_per_split);
            int t, kb0, kb1; bool seg_first, seg_last;
            while (wi.next(t, kb0, kb1, seg_first, seg_last)) {
                const int split = t / tiles_mn;
                const int mn = t - split * tiles_mn;
                const int mt = mn / p.num_n, nt = mn - mt * p.num_n;
                const int m0 = mt * GEMM_BM, n0 = p.diag64 ? (2 * mt + nt) * 64 : nt * BN;
                const int bn0 = n0 + nt * p.b_nadj_per_ntile, bkoff = nt * p.b_koff_per_ntile;
                int img = 0, hin0 = 0;
                if (p.conv) {
                    if (p.BIMG == 1) { img = mt / p.tiles_per_img; hin0 = (mt - img * p.tiles_per_img) * p.BH * p.stride - p.pad; }
                    else { img = mt * p.BIMG; hin0 = -p.pad; }
                }
                // running (k-block, term) and (filter row, filter column, channel block) counters: no divisions in the loop
                int kk = FUSE3 ? kb0 : kb0 / p.nterms, term = FUSE3 ? 0 : kb0 - kk * p.nterms;
                int tap_r = 0, tap_s = 0, cb = 0;
                if (p.conv) { const int tap = kk / p.cblocks; cb = kk - tap * p.cblocks; tap_r = tap / p.S; tap_s = tap - tap_r * p.S; }
                const bool prof_on = p.prof != nullptr;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (prof_on) {
                        const long long w0 = clock64();
                        mbar_wait(&empty[stage], phase ^ 1);
                        pw += clock64() - w0;
                    } else {
                        mbar_wait(&empty[stage], phase ^ 1);
                    }
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    uint8_t* sbase = smem + stage * Cfg::STAGE_BYTES;
                    if constexpr (FUSE3) {
                        const int k0 = kb * GEMM_BK;
                        if (p.a_mn) {
                            tma_load_2d(&mapA0, &full[stage], sbase, m0, k0);
                            tma_load_2d(&mapA0, &full[stage], sbase + 8192, m0 + 64, k0);
                            tma_load_2d(&mapA1, &full[stage], sbase + OFF_ALO, m0, k0);
                            tma_load_2d(&mapA1, &full[stage], sbase + OFF_ALO + 8192, m0 + 64, k0);
                        } else {
                            tma_load_2d(&mapA0, &full[stage], sbase, k0, m0);
                            tma_load_2d(&mapA1, &full[stage], sbase + OFF_ALO, k0, m0);
                        }
                        if (p.b_mn) {
#pragma unroll
                            for (int a = 0; a < BN / 64; ++a) {
                                tma_load_2d(&mapB0, &full[stage], sbase + OFF_B + a * 8192, n0 + a * 64, k0);
                                tma_load_2d(&mapB1, &full[stage], sbase + OFF_BLO + a * 8192, n0 + a * 64, k0);
                            }
                        } else {
                            tma_load_2d(&mapB0, &full[stage], sbase + OFF_B, k0, n0);
                            tma_load_2d(&mapB1, &full[stage], sbase + OFF_BLO, k0, n0);
                        }
                    } else {
                    const CUtensorMap* ma = (term == 2) ? &mapA1 : &mapA0;
                    const CUtensorMap* mb = (term == 1) ? &mapB1 : &mapB0;
                    uint8_t* sa = sbase;
                    uint8_t* sb = sbase + OFF_B;
                    const int k0 = kk * GEMM_BK;
                    if (p.conv) {
                        tma_load_4d(ma, &full[stage], sa, cb * GEMM_BK, tap_s - p.pad, hin0 + tap_r, img);
                    } else if (p.a_mn) {
                        tma_load_2d(ma, &full[stage], sa, m0, k0);
                        tma_load_2d(ma, &full[stage], sa + 8192, m0 + 64, k0);
                    } else {
                        tma_load_2d(ma, &full[stage], sa, k0 + nt * p.a_koff_per_ntile, m0);
                    }
                    if (p.b_mn) {
#pragma unroll
                        for (int a = 0; a < BN / 64; ++a) tma_load_2d(mb, &full[stage], sb + a * 8192, bn0 + a * 64, k0 + bkoff);
                    } else {
                        tma_load_2d(mb, &full[stage], sb, k0 + bkoff, bn0);
                    }
                    if (++term == p.nterms) {
                        term = 0; ++kk;
                        if (p.conv && ++cb == p.cblocks) { cb = 0; if (++tap_s == p.S) { tap_s = 0; ++tap_r; } }
                    }
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (p.prof) { p.prof[blockIdx.x * 16 + 0] = pw; p.prof[blockIdx.x * 16 + 1] = clock64() - pt0; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc_bf16(BN, p.a_mn, p.b_mn);
        const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
        const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
        int stage = 0; uint32_t phase = 0; int iter = 0;
        long long mw_full = 0, mw_tmem = 0; const long long mt0 = clock64();
        WorkIter wi(p.streamk, p.sk_t0, tiles_mn, total_tiles, total_kb, p.kb_per_split);
        int t, kb0, kb1; bool seg_first, seg_last;
        for (; wi.next(t, kb0, kb1, seg_first, seg_last); ++iter) {
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            const long long w1 = clock64();
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            mw_tmem += clock64() - w1;
            tc_fence_after();
            const uint32_t tacc = tmem_base + as * BN;
            for (int kb = kb0; kb < kb1; ++kb) {
                const long long w2 = clock64();
                mbar_wait(&full[stage], phase);
                mw_full += clock64() - w2;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sbase = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    if constexpr (FUSE3) {
                        // hi*hi, hi*lo, lo*hi from the same stage
                        const uint32_t abase[3] = {sbase, sbase, sbase + OFF_ALO};
                        const uint32_t bbase[3] = {sbase + OFF_B, sbase + OFF_BLO, sbase + OFF_B};
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
#pragma unroll
                            for (int k = 0; k < GEMM_BK / 16; ++k) {
                                const uint64_t da = make_smem_desc(abase[term] + k * a_kstep, 1024, a_lbo);
                                const uint64_t db = make_smem_desc(bbase[term] + k * b_kstep, 1024, b_lbo);
                                umma_bf16(tacc, da, db, idesc, (kb > kb0 || k > 0 || term > 0) ? 1u : 0u);
                            }
                        }
                    } else {
                        const uint32_t a_base = sbase;
                        const uint32_t b_base = sbase + OFF_B;
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            const uint64_t da = make_smem_desc(a_base + k * a_kstep, 1024, a_lbo);
                            const uint64_t db = make_smem_desc(b_base + k * b_kstep, 1024, b_lbo);
                            umma_bf16(tacc, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (kb == kb1 - 1) umma_commit(&tmem_full[as]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        if (p.prof && lane == 0) { p.prof[blockIdx.x * 16 + 2] = mw_full; p.prof[blockIdx.x * 16 + 3] = mw_tmem; p.prof[blockIdx.x * 16 + 4] = clock64() - mt0; }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;                       // TMEM lane quadrant this warp may read
        const int cbeg = ((warp - 4) >> 2) * (BN / 2), cend = cbeg + BN / 2;   // its half of the tile's columns
        int iter = 0;
        long long ew = 0; const long long et0 = clock64();
        WorkIter wi(p.streamk, p.sk_t0, tiles_mn, total_tiles, total_kb, p.kb_per_split);
        int t, kb0, kb1; bool seg_first, seg_last;
        for (; wi.next(t, kb0, kb1, seg_first, seg_last); ++iter) {
            const int split = t / tiles_mn;
            const int mn = t - split * tiles_mn;
            const int mt = mn / p.num_n, nt = mn - mt * p.num_n;
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            [[maybe_unused]] uint4 pf_rr[4];
            [[maybe_unused]] bool pf_staged = false;
            if constexpr (CLS == EPI_CONV_PF) {
                // the residual of the tile's first chunk does not depend on the MMAs: request it before waiting for them
                const int pn0 = nt * BN;
                pf_staged = !p.diag64 && !p.streamk && p.split_k <= 1 && pn0 + cend <= p.N && ((p.e.ld_bf | p.e.ld_f32 | p.e.ld_res) & 7) == 0 &&
                            ((reinterpret_cast<uintptr_t>(p.e.bias) | reinterpret_cast<uintptr_t>(p.e.out_f32) |
                              reinterpret_cast<uintptr_t>(p.e.out_hi) | reinterpret_cast<uintptr_t>(p.e.res_bf)) & 15) == 0;
                if (pf_staged && p.e.res_bf) conv_pf_load_res(pf_rr, p.e, (long long)mt * GEMM_BM + q * 32, lane, p.M, pn0 + cbeg);
            }
            const long long w3 = clock64();
            mbar_wait(&tmem_full[as], aphase);
            ew += clock64() - w3;
            tc_fence_after();
            const long long row = (long long)mt * GEMM_BM + q * 32 + lane;
            const int n0 = p.diag64 ? (2 * mt + nt) * 64 : nt * BN;
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            // diag64: a row belongs to head row/64 and only that head's 64 columns are kept, stored at columns 0..63
            const bool row_ok = row < p.M && (!p.diag64 || (row >> 6) == (long long)(n0 >> 6));
            const int ncols = p.diag64 ? 64 : p.N;
            const int c0 = p.diag64 ? 0 : n0;
            // ---- stream-K: publish a partial segment, or collect the partials of the tile this segment finishes
            const int r_in = q * 32 + lane;
            int nprev = 0;
            if (p.streamk && !seg_last) {
                float* ws = p.partial + (size_t)blockIdx.x * BN * 128;
                for (int c = cbeg; c < cend; c += 32) {
                    float v[32];
                    tmem_ld32(tacc + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) ws[(size_t)(c + i) * 128 + r_in] = v[i];
                }
                __threadfence();
                __syncwarp();
                if (lane == 0) *reinterpret_cast<volatile int*>(p.sk_flags + blockIdx.x * 8 + (warp - 4)) = 1;
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
                continue;
            }
            if (p.streamk && !seg_first) {
                const long long tb = (long long)(t - p.sk_t0) * total_kb;
                int c2 = blockIdx.x - 1;
                while (true) { ++nprev; if (c2 <= 0 || streamk_range_start(c2, tiles_mn - p.sk_t0, total_kb, gridDim.x) <= tb) break; --c2; }
                if (lane == 0)
                    for (int j = 1; j <= nprev; ++j)
                        while (*reinterpret_cast<volatile int*>(p.sk_flags + (blockIdx.x - j) * 8 + (warp - 4)) == 0) {}
                __syncwarp();
                __threadfence();
            }
            if (p.split_k > 1) {
                float* dst = p.partial + ((long long)split * p.M + row) * ncols + c0;
                for (int c = cbeg; c < cend && n0 + c < p.N; c += 32) {
                    float v[32];
                    tmem_ld32(tacc + c, v);
                    tmem_ld_wait();
                    if (row_ok) {
                        const int nvalid = min(32, p.N - n0 - c);
                        if (nvalid == 32 && (ncols & 3) == 0) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        } else {
                            for (int i = 0; i < nvalid; ++i) dst[c + i] = v[i];
                        }
                    }
                }
            } else if ((CLS == EPI_GENERIC || CLS == EPI_LN64) && p.e.act == ACT_LN64_GELU) {
                if constexpr (CLS == EPI_GENERIC || CLS == EPI_LN64) {
                    for (int c = cbeg; c < cend && n0 + c < p.N; c += 64) {
                        float v[64];
                        tmem_ld32(tacc + c, v);
                        tmem_ld32(tacc + c + 32, v + 32);
                        tmem_ld_wait();
                        if (nprev) streamk_add<64>(v, p.partial, nprev, BN, c, r_in);
                        if (row_ok) epilogue_apply<64, CLS>(v, row, n0 + c, p.N, p.e);
                    }
                }
            } else if constexpr (CLS == EPI_CONV) {
                // coalesced path: whole warp cooperates (rows beyond M are predicated inside)
                const bool staged = !p.diag64 && n0 + cend <= p.N && ((p.e.ld_bf | p.e.ld_f32 | p.e.ld_res) & 7) == 0 &&
                                    ((reinterpret_cast<uintptr_t>(p.e.bias) | reinterpret_cast<uintptr_t>(p.e.out_f32) |
                                      reinterpret_cast<uintptr_t>(p.e.out_hi) | reinterpret_cast<uintptr_t>(p.e.res_bf)) & 15) == 0;
                uint8_t* stg = smem + STAGES * Cfg::STAGE_BYTES + 256 + (warp - 4) * GEMM_STG_BYTES;
                const long long row_base = (long long)mt * GEMM_BM + q * 32;
                if (staged) {
                    if constexpr (BN >= 128) {
                        for (int c = cbeg; c < cend; c += 64) {
                            float v[64];
                            tmem_ld32(tacc + c, v);
                            tmem_ld32(tacc + c + 32, v + 32);
                            tmem_ld_wait();
                            if (nprev) streamk_add<64>(v, p.partial, nprev, BN, c, r_in);
                            epilogue_conv_staged<64>(v, row_base, lane, p.M, n0 + c, p.e, stg);
                        }
                    } else {
                        for (int c = cbeg; c < cend; c += 32) {
                            float v[32];
                            tmem_ld32(tacc + c, v);
                            tmem_ld_wait();
                            if (nprev) streamk_add<32>(v, p.partial, nprev, BN, c, r_in);
                            epilogue_conv_staged<32>(v, row_base, lane, p.M, n0 + c, p.e, stg);
                        }
                    }
                } else {
                    for (int c = cbeg; c < cend && n0 + c < p.N; c += 32) {
                        float v[32];
                        tmem_ld32(tacc + c, v);
                        tmem_ld_wait();
                        if (nprev) streamk_add<32>(v, p.partial, nprev, BN, c, r_in);
                        if (row_ok) epilogue_apply<32, CLS>(v, row, c0 + c, ncols, p.e);
                    }
                }
            } else if constexpr (CLS == EPI_CONV_PF) {
                uint8_t* stg = smem + STAGES * Cfg::STAGE_BYTES + 256 + (warp - 4) * GEMM_PF_STG_BYTES;
                const long long row_base = (long long)mt * GEMM_BM + q * 32;
                if (pf_staged) {
                    const bool has_res = p.e.res_bf != nullptr;
#pragma unroll
                    for (int ci = 0; ci < BN / 64; ++ci) {
                        const int c = cbeg + ci * 32;
                        float v[32];
                        uint4 rr_next[4];
                        tmem_ld32(tacc + c, v);
                        if (has_res && ci + 1 < BN / 64) conv_pf_load_res(rr_next, p.e, row_base, lane, p.M, n0 + c + 32);   // one chunk ahead
                        tmem_ld_wait();
                        conv_pf_chunk(v, pf_rr, has_res, row_base, lane, p.M, n0 + c, p.e, stg);
                        if (ci + 1 < BN / 64) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) pf_rr[i] = rr_next[i];
                        }
                    }
                } else {
                    for (int c = cbeg; c < cend && n0 + c < p.N; c += 32) {
                        float v[32];
                        tmem_ld32(tacc + c, v);
                        tmem_ld_wait();
                        if (nprev) streamk_add<32>(v, p.partial, nprev, BN, c, r_in);
                        if (row_ok) epilogue_apply<32, CLS>(v, row, c0 + c, ncols, p.e);
                    }
                }
            } else if constexpr (CLS != EPI_LN64) {
                for (int c = cbeg; c < cend && n0 + c < p.N; c += 32) {
                    float v[32];
                    tmem_ld32(tacc + c, v);
                    tmem_ld_wait();
                    if (nprev) streamk_add<32>(v, p.partial, nprev, BN, c, r_in);
                    if (row_ok) epilogue_apply<32, CLS>(v, row, c0 + c, ncols, p.e);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        if (p.prof && lane == 0) { p.prof[blockIdx.x * 16 + 5 + (warp - 4)] = ew; if (warp == 4) p.prof[blockIdx.x * 16 + 13] = clock64() - et0; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// Sums split-K partials and applies the fused epilogue. One thread per (row, 8-column chunk): a warp reads 1 KB
// of consecutive columns per split, in a fixed order (deterministic).
static __global__ void splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N, EpiParams e) {
    const int chunks = (N + 7) / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)M * chunks) return;
    const int chunk = (int)(idx % chunks);
    const long long row = idx / chunks;
    const int col0 = chunk * 8;
    const int nvalid = min(8, N - col0);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    const bool vec = (nvalid == 8) && ((N & 3) == 0);
    for (int s = 0; s < splits; ++s) {
        const float* src = partial + ((long long)s * M + row) * N + col0;
        if (vec) {
            const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (i < nvalid) v[i] += src[i];
        }
    }
    epilogue_apply<8>(v, row, col0, N, e);
}

// Same, for the per-head LayerNorm(64)+GELU epilogue (needs whole 64-column groups).
static __global__ void splitk_reduce_ln64_kernel(const float* __restrict__ partial, int splits, int M, int N, EpiParams e) {
    const int chunks = N / 64;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)M * chunks) return;
    const int chunk = (int)(idx % chunks);
    const long long row = idx / chunks;
    const int col0 = chunk * 64;
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0.f;
    for (int s = 0; s < splits; ++s) {
        const float* src = partial + ((long long)s * M + row) * N + col0;
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] += src[i];
    }
    epilogue_apply<64>(v, row, col0, N, e);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Operand {
    const bf16* hi; const bf16* lo;   // lo may be null when nterms == 1
    int mn_major;                     // 0: [MN][K] row-major, 1: [K][MN] row-major
    long long ld;                     // elements between consecutive rows of the stored matrix
};

struct ConvGeom {
    int enabled;
    int batch, H, W, C;      // NHWC input
    int Ho, Wo, R, S, stride, pad;
};

struct GemmArgs {
    int M, N, K;
    Operand A, B;
    int nterms;
    int split_k;             // 0/1 = off
    int a_koff_per_ntile;
    int b_koff_per_ntile, b_nadj_per_ntile;   // block-diagonal dgrad (see GemmKParams)
    int b_map_mn, b_map_k;   // explicit extents of B's tensor map (0 = N / K)
    int diag64;              // block-diagonal wgrad
    int epi_conv_pref;       // prefer the coalesced convolution epilogue class even for a plain GEMM (1x1 convolutions)
    int block_n;             // 0 = auto
    ConvGeom conv;
    float* workspace; size_t workspace_bytes;   // for split-K partials
    EpiParams e;
};

void set_error(const char* fmt, ...);
int num_sms();
int split_planes(const float* x, bf16* hi, bf16* lo, long long rows, int cols, long long ld_in, long long ld_out, cudaStream_t st);
// optional per-launch CUDA-event timing of the GEMM kernel (bench.py's roofline leg); no-ops unless enabled
void gemm_timing_begin(cudaStream_t st, double flops);
void gemm_timing_end(cudaStream_t st);
void gemm_timing_chain_begin(cudaStream_t st);   // one event pair around a chain of launches (per-launch events suppressed inside)
void gemm_timing_chain_end(cudaStream_t st);
unsigned long long* gemm_prof_buffer();
struct SideStream { cudaStream_t s; cudaEvent_t fork, join; bool ok; };
SideStream* side_stream();                                   // null when disabled (SQ_SIDE_STREAM=0) or unavailable
cudaStream_t side_fork(SideStream* ss, cudaStream_t st);     // side stream now waits for everything enqueued on st so far
void side_join(SideStream* ss, cudaStream_t st);             // st now waits for everything enqueued on the side stream so far   // device buffer for per-CTA role cycle counters, or null

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_fn();

inline int encode_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box, const cuuint32_t* estr) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return -1; }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] stride0 %llu box [%u,%u,%u,%u] base %p",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                  (unsigned long long)(rank > 1 ? strides[0] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
                  rank > 3 ? box[3] : 0, base);
        return -1;
    }
    return 0;
}

// 2-D operand map. rows_mn = extent along M or N, K = reduction extent, box_mn = tile extent along MN.
inline int make_operand_map(CUtensorMap* m, const bf16* base, int mn_major, long long ld, int rows_mn, int K, int box_mn) {
    cuuint64_t dims[2], strides[1]; cuuint32_t box[2], estr[2] = {1, 1};
    if (!mn_major) { dims[0] = K; dims[1] = rows_mn; box[0] = GEMM_BK; box[1] = box_mn; }
    else { dims[0] = rows_mn; dims[1] = K; box[0] = 64; box[1] = GEMM_BK; }
    strides[0] = (cuuint64_t)ld * 2;
    return encode_map(m, base, 2, dims, strides, box, estr);
}

template <int BN, int CLS, int FUSE3 = 0>
int launch_gemm_inst(const CUtensorMap* maps, const GemmKParams& kp, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(gemm_tc_kernel<BN, CLS, FUSE3>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN, CLS, FUSE3>::SMEM_BYTES);
        if (err != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return -1; }
        configured = true;
    }
    static const int pdl = getenv("SQ_PDL") ? atoi(getenv("SQ_PDL")) : 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = GemmCfg<BN, CLS, FUSE3>::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t err = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CLS, FUSE3>, maps[0], maps[1], maps[2], maps[3], kp);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("gemm launch: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

// Defined once in common.cu (explicit dispatch over the instantiated <BN, class> pairs).
int launch_gemm_dispatch(int bn, int cls, int fuse3, const CUtensorMap* maps, const GemmKParams& kp, int grid, cudaStream_t st);

// The narrowest epilogue class that covers the options a launch uses.
inline int epi_class_of(const EpiParams& e, bool split, bool conv_mode) {
    if (split) return EPI_F32;          // the kernel only writes partials; the reduce kernel applies the epilogue
    bool used[13] = {e.alpha != 1.0f, e.bias != nullptr, e.rowbias != nullptr, e.res_f32 != nullptr, e.res_bf != nullptr, e.save_pre != nullptr,
                     e.act == ACT_RELU, e.act == ACT_GELU, e.act == ACT_MUL_DGELU, e.act == ACT_LN64_GELU, e.out_f32 != nullptr,
                     e.out_hi != nullptr, e.out_lo != nullptr};
    const int order_conv[5] = {EPI_CONV, EPI_F32, EPI_GELU, EPI_DGELU, EPI_LN64};
    const int order_gemm[5] = {EPI_F32, EPI_CONV, EPI_GELU, EPI_DGELU, EPI_LN64};      // plain GEMMs keep the deeper pipeline of the f32 class
    for (int k = 0; k < 5; ++k) {
        const int cls = conv_mode ? order_conv[k] : order_gemm[k];
        bool ok = true;
        for (int o = 0; o < 13; ++o) if (used[o] && !epi_has(cls, o)) ok = false;
        if (cls == EPI_CONV && e.act == ACT_GELU && !conv_mode) ok = false;     // plain GEMMs keep the GELU class unless asked
        if (cls == EPI_GELU && e.act != ACT_GELU) ok = false;
        if (cls == EPI_DGELU && e.act != ACT_MUL_DGELU) ok = false;
        if (cls == EPI_LN64 && e.act != ACT_LN64_GELU) ok = false;
        if (ok) return cls;
    }
    return EPI_GENERIC;
}

inline int gemm_launch(const GemmArgs& g, cudaStream_t st) {
    GemmKParams kp;
    memset(&kp, 0, sizeof(kp));
    int bn = g.block_n;
    static const int env_bn = getenv("SQ_GEMM_BN") ? atoi(getenv("SQ_GEMM_BN")) : 0;     // tuning knob (experiments only)
    if (bn == 0 && env_bn && g.e.act != ACT_LN64_GELU && g.N >= env_bn) bn = env_bn;
    if (bn == 0) {
        if (g.e.act == ACT_LN64_GELU) bn = 128;
        else if (g.N <= 64) bn = 64;
        else {
            // prefer the widest tile that still gives every SM work
            const int mt = (g.M + GEMM_BM - 1) / GEMM_BM;
            const long long t256 = (long long)mt * ((g.N + 255) / 256);
            bn = (g.N >= 256 && t256 >= (long long)num_sms() * 2 / 3) ? 256 : 128;
        }
    }
    // split-precision GEMMs may also use 192-wide tiles: for M = 3200, N = 2048 that is 275 tiles = 2 waves of 192 columns
    // instead of 2 waves of 256 (tile time is proportional to the width)
    static const int bn192_enabled = getenv("SQ_BN192") ? atoi(getenv("SQ_BN192")) : 1;
    if (bn192_enabled && g.block_n == 0 && !env_bn && bn == 256 && g.nterms == 3 && !g.conv.enabled && !g.diag64 && g.split_k <= 1 &&
        g.e.act != ACT_LN64_GELU && !g.a_koff_per_ntile && !g.b_koff_per_ntile) {
        const long long mt = (g.M + GEMM_BM - 1) / GEMM_BM, G = num_sms();
        auto cost = [&](int w) { const long long tiles = mt * ((g.N + w - 1) / w); return ((tiles + G - 1) / G) * w; };
        if (cost(192) < cost(256) && epi_class_of(g.e, false, false) != EPI_GENERIC && epi_class_of(g.e, false, false) != EPI_CONV) bn = 192;
    }
    if (bn != 64 && bn != 128 && bn != 192 && bn != 256) { set_error("gemm: bad block_n %d", bn); return -1; }
    kp.M = g.M; kp.N = g.N;
    kp.num_m = (g.M + GEMM_BM - 1) / GEMM_BM;
    kp.num_n = (g.N + bn - 1) / bn;
    kp.nterms = g.nterms;
    kp.a_mn = g.A.mn_major; kp.b_mn = g.B.mn_major;
    kp.a_koff_per_ntile = g.a_koff_per_ntile;
    kp.b_koff_per_ntile = g.b_koff_per_ntile; kp.b_nadj_per_ntile = g.b_nadj_per_ntile;
    kp.diag64 = g.diag64;
    if (g.diag64) {
        if (bn != 64 || g.M != g.N || g.M % 64 != 0 || g.conv.enabled || g.e.act == ACT_LN64_GELU) { set_error("gemm: diag64 needs block_n 64, M == N, M %% 64 == 0"); return -1; }
        kp.num_n = 2;
    }
    kp.e = g.e;
    kp.prof = gemm_prof_buffer();
    if (kp.e.alpha == 0.0f) kp.e.alpha = 1.0f;
    if (g.nterms != 1 && g.nterms != 3) { set_error("gemm: nterms must be 1 or 3"); return -1; }
    if (g.nterms == 3 && (!g.A.lo || !g.B.lo)) { set_error("gemm: split precision needs lo planes"); return -1; }

    CUtensorMap maps[4];
    if (g.conv.enabled) {
        const ConvGeom& c = g.conv;
        if (c.C % 64 != 0) { set_error("conv: C=%d must be a multiple of 64", c.C); return -1; }
        if (g.A.mn_major) { set_error("conv: A must be NHWC"); return -1; }
        int BW = c.Wo, BH, BIMG = 1;
        if (BW > 128 || 128 % BW != 0) { set_error("conv: Wo=%d must divide 128", c.Wo); return -1; }
        BH = 128 / BW;
        if (BH > c.Ho) { BIMG = BH / c.Ho; BH = c.Ho; if (BIMG * BH * BW != 128) { set_error("conv: tile does not fit Ho=%d Wo=%d", c.Ho, c.Wo); return -1; } }
        if (c.Ho % BH != 0) { set_error("conv: Ho=%d not a multiple of %d", c.Ho, BH); return -1; }
        kp.conv = 1; kp.cblocks = c.C / 64; kp.S = c.S; kp.stride = c.stride; kp.pad = c.pad;
        kp.tiles_per_img = c.Ho / BH; kp.BH = BH; kp.BIMG = BIMG;
        kp.nk = c.R * c.S * kp.cblocks;
        if (g.M != c.batch * c.Ho * c.Wo) { set_error("conv: M mismatch"); return -1; }
        cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.batch};
        cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)(BW * c.stride), (cuuint32_t)(BH * c.stride), (cuuint32_t)BIMG};
        cuuint32_t estr[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
        if (encode_map(&maps[0], g.A.hi, 4, dims, strides, box, estr)) return -1;
        if (g.nterms == 3) { if (encode_map(&maps[1], g.A.lo, 4, dims, strides, box, estr)) return -1; }      // split precision: the lo plane of the NHWC input
        else maps[1] = maps[0];
    } else {
        kp.nk = (g.K + GEMM_BK - 1) / GEMM_BK;
        if (make_operand_map(&maps[0], g.A.hi, g.A.mn_major, g.A.ld, g.M, g.a_koff_per_ntile ? (int)g.A.ld : g.K, GEMM_BM)) return -1;
        if (g.nterms == 3) { if (make_operand_map(&maps[1], g.A.lo, g.A.mn_major, g.A.ld, g.M, g.a_koff_per_ntile ? (int)g.A.ld : g.K, GEMM_BM)) return -1; }
        else maps[1] = maps[0];
    }
    const int bmn = g.b_map_mn ? g.b_map_mn : g.N;
    const int bk = g.b_map_k ? g.b_map_k : (g.conv.enabled ? kp.nk * 64 : g.K);
    if (make_operand_map(&maps[2], g.B.hi, g.B.mn_major, g.B.ld, bmn, bk, bn)) return -1;
    if (g.nterms == 3) { if (make_operand_map(&maps[3], g.B.lo, g.B.mn_major, g.B.ld, bmn, bk, bn)) return -1; }
    else maps[3] = maps[2];

    // fused split-precision stages (see GemmCfg): plain 3-term GEMMs with wide tiles, when the epilogue class has a fused build
    static const int fuse_enabled = getenv("SQ_FUSE3") ? atoi(getenv("SQ_FUSE3")) : 1;
    int fuse3 = (fuse_enabled && g.nterms == 3 && bn >= 128 && !g.conv.enabled && !g.diag64 && !g.a_koff_per_ntile && !g.b_koff_per_ntile) ? 1 : 0;
    int total_kb = 0, split = 1, cls = EPI_GENERIC;
    for (int attempt = 0; attempt < 2; ++attempt) {
        total_kb = fuse3 ? kp.nk : kp.nk * kp.nterms;
        split = g.split_k > 1 ? g.split_k : 1;
        if (split > total_kb) split = total_kb;
        kp.kb_per_split = (total_kb + split - 1) / split;
        split = (total_kb + kp.kb_per_split - 1) / kp.kb_per_split;
        cls = epi_class_of(kp.e, split > 1, g.conv.enabled != 0 || g.epi_conv_pref != 0);
        static const int conv_pf = getenv("SQ_CONV_EPI_PF") ? atoi(getenv("SQ_CONV_EPI_PF")) : 0;   // opt-in until measured on hardware
        if (conv_pf && cls == EPI_CONV) cls = EPI_CONV_PF;
        if (fuse3 && (cls == EPI_GENERIC || cls == EPI_CONV || cls == EPI_CONV_PF)) { fuse3 = 0; continue; }
        break;
    }
    kp.split_k = split;
    if (split > 1) {
        const size_t need = (size_t)split * g.M * (g.diag64 ? 64 : g.N) * sizeof(float);
        if (!g.workspace || g.workspace_bytes < need) { set_error("gemm: split-K workspace too small (%zu < %zu)", g.workspace_bytes, need); return -1; }
        kp.partial = g.workspace;
    }
    const long long total_tiles = (long long)kp.num_m * kp.num_n * split;
    int grid = (int)(total_tiles < num_sms() ? total_tiles : num_sms());
    // stream-K when whole-tile scheduling would leave the last wave mostly empty (e.g. 200 tiles on 148 SMs)
    static const int sk_enabled = getenv("SQ_STREAMK") ? atoi(getenv("SQ_STREAMK")) : 0;   // correct but not faster yet (r01): opt-in
    if (sk_enabled && !g.conv.enabled && !g.diag64 && split == 1 && !g.a_koff_per_ntile && !g.b_koff_per_ntile && g.workspace) {
        const int G = num_sms();
        const long long rounds = (total_tiles + G - 1) / G;
        const double eff = (double)total_tiles / (double)(rounds * G);
        const size_t need = (size_t)G * bn * 128 * 4 + (size_t)G * 8 * 4;
        const long long t0 = (total_tiles / G) * G;                 // whole waves stay tile-scheduled
        const long long rem = total_tiles - t0;                     // tiles of the last wave
        // every CTA must get a non-empty share and a tile may be spread over a handful of CTAs at most
        if (eff < 0.9 && rem > 0 && rem * (long long)total_kb >= 2LL * G && rem * 8 >= G && total_kb >= 4 && g.workspace_bytes >= need) {
            kp.streamk = 1; kp.sk_t0 = (int)t0; kp.partial = g.workspace;
            kp.sk_flags = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(g.workspace) + (size_t)G * bn * 128 * 4);
            grid = G;
            cudaMemsetAsync(kp.sk_flags, 0, (size_t)G * 8 * 4, st);
        }
    }
    int rc;
    gemm_timing_begin(st, 2.0 * g.M * g.N * (g.conv.enabled ? (double)kp.nk * 64 : (double)g.K) * g.nterms);
    rc = launch_gemm_dispatch(bn, cls, fuse3, maps, kp, grid, st);
    gemm_timing_end(st);
    if (rc) return rc;
    if (split > 1) {
        EpiParams e = kp.e;
        e.alpha = kp.e.alpha;
        if (e.act == ACT_LN64_GELU) {
            const long long n = (long long)g.M * (g.N / 64);
            splitk_reduce_ln64_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kp.partial, split, g.M, g.N, e);
        } else {
            const int rn = g.diag64 ? 64 : g.N;      // diag64 partials are a dense [M, 64] matrix
            const long long n = (long long)g.M * ((rn + 7) / 8);
            splitk_reduce_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(kp.partial, split, g.M, rn, e);
        }
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) { set_error("splitk reduce launch: %s", cudaGetErrorString(err)); return -1; }
    }
    return 0;
}

}  // namespace sq
