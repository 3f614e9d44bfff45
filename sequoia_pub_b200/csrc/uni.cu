// UNI ViT-L/16 feature extractor (reference call sites: pre_processing/compute_features_hdf5.py:63-66 (timm
// "vit_large_patch16_224", init_values -> LayerScale, num_classes=0) and :125-129; the arithmetic is timm's, restated in
// oracle/uni_oracle.py — PARITY UNPINNED).
//
// Data layout: residual stream x fp32 [B*197, 1024]; GEMM operands bf16 (single-pass tensor-core GEMMs with fp32
// accumulation, like the ResNet path); weights pre-cast to bf16 with the LayerScale gammas folded into proj / fc2.
// Per block: LN -> qkv GEMM(+bias) -> fused softmax attention (mma.sync bf16, one CTA per (image, head)) ->
// proj GEMM(+bias, +residual) -> LN -> fc1 GEMM(+bias, GELU) -> fc2 GEMM(+bias, +residual).
#include "convgemm.cuh"
#include "../../include/sequoia_b200.h"

namespace sq {

constexpr int U_DIM = 1024, U_HEADS = 16, U_HD = 64, U_MLP = 4096, U_TOK = 197, U_GRID = 14, U_PK = 768;

struct UniW { long long pe, blk[64][4], total; };                     // bf16: patch-embed, per block {qkv, proj, fc1, fc2}
struct UniV { long long cls, pos, pb, blk[64][8], ng, nb, total; };   // fp32: per block {n1g, n1b, bqkv, bproj, n2g, n2b, b1, b2}

static void uni_layout(int depth, UniW* w, UniV* v) {
    long long o = 0;
    auto tw = [&](long long n) { long long r = o; o += (n + 63) / 64 * 64; return r; };
    w->pe = tw((long long)U_DIM * U_PK);
    for (int i = 0; i < depth; ++i) {
        w->blk[i][0] = tw(3LL * U_DIM * U_DIM); w->blk[i][1] = tw((long long)U_DIM * U_DIM);
        w->blk[i][2] = tw((long long)U_MLP * U_DIM); w->blk[i][3] = tw((long long)U_DIM * U_MLP);
    }
    w->total = o;
    o = 0;
    v->cls = tw(U_DIM); v->pos = tw((long long)U_TOK * U_DIM); v->pb = tw(U_DIM);
    for (int i = 0; i < depth; ++i) {
        v->blk[i][0] = tw(U_DIM); v->blk[i][1] = tw(U_DIM); v->blk[i][2] = tw(3 * U_DIM); v->blk[i][3] = tw(U_DIM);
        v->blk[i][4] = tw(U_DIM); v->blk[i][5] = tw(U_DIM); v->blk[i][6] = tw(U_MLP); v->blk[i][7] = tw(U_DIM);
    }
    v->ng = tw(U_DIM); v->nb = tw(U_DIM);
    v->total = o;
}

// W fp32 [rows, cols] -> bf16, row r scaled by scale[r] (LayerScale folded into the producing Linear)
__global__ void uni_cast_kernel(const float* __restrict__ w, const float* __restrict__ scale, long long rows, int cols, bf16* __restrict__ out) {
    const long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float s = scale ? scale[i / cols] : 1.0f;
        out[i] = __float2bfloat16_rn(w[i] * s);
    }
}
__global__ void uni_vec_kernel(const float* __restrict__ v, const float* __restrict__ scale, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = scale ? v[i] * scale[i] : v[i];
}

// 16x16 patches -> rows of the patch-embed GEMM: col[b*196 + p][c*256 + kh*16 + kw], bf16.
// kind 0: uint8 [B,224,224,3] (ToTensor + Normalize fused), kind 1: fp32 [B,3,224,224] already normalised.
__global__ void uni_patchify_kernel(const void* __restrict__ in, int kind, int batch, bf16* __restrict__ col) {
    const long long n = (long long)batch * 196 * (U_PK / 8);
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k8 = (int)(i % (U_PK / 8)) * 8;
        const long long bp = i / (U_PK / 8);
        const int p = (int)(bp % 196); const int b = (int)(bp / 196);
        const int c = k8 >> 8, kh = (k8 >> 4) & 15, kw = k8 & 15;
        const int y = (p / U_GRID) * 16 + kh, x0 = (p % U_GRID) * 16 + kw;
        uint4 pk; bf16* h = reinterpret_cast<bf16*>(&pk);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v;
            if (kind == 0) {
                const uint8_t u = reinterpret_cast<const uint8_t*>(in)[(((long long)b * 224 + y) * 224 + x0 + j) * 3 + c];
                v = (static_cast<float>(u) / 255.0f - mean[c]) / stdv[c];
            } else {
                v = reinterpret_cast<const float*>(in)[(((long long)b * 3 + c) * 224 + y) * 224 + x0 + j];
            }
            h[j] = __float2bfloat16_rn(v);
        }
        *reinterpret_cast<uint4*>(col + bp * U_PK + k8) = pk;
    }
}

// x[b, 0] = cls + pos[0]; x[b, 1+p] = patch[b*196+p] + pos[1+p]
__global__ void uni_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls, const float* __restrict__ pos, int batch,
                                    float* __restrict__ x) {
    const long long n = (long long)batch * U_TOK * (U_DIM / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % (U_DIM / 4)) * 4;
        const long long bt = i / (U_DIM / 4);
        const int t = (int)(bt % U_TOK); const long long b = bt / U_TOK;
        const float4 a = t == 0 ? *reinterpret_cast<const float4*>(cls + c) : *reinterpret_cast<const float4*>(patch + (b * 196 + t - 1) * U_DIM + c);
        const float4 q = *reinterpret_cast<const float4*>(pos + (long long)t * U_DIM + c);
        *reinterpret_cast<float4*>(x + bt * U_DIM + c) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
    }
}

// LayerNorm over 1024 columns, one warp per row (32 values per lane), eps 1e-6; bf16 and/or fp32 output.
__global__ void __launch_bounds__(256) uni_ln_kernel(const float* __restrict__ x, long long row_stride, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, long long rows, bf16* __restrict__ ob, float* __restrict__ of) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + row * row_stride;
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] = *reinterpret_cast<const float4*>(xr + (i * 32 + lane) * 4); s += v[i].x + v[i].y + v[i].z + v[i].w; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / U_DIM);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + c * c + d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / U_DIM) + 1e-6f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
        const float4 y = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                     (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        if (of) *reinterpret_cast<float4*>(of + row * U_DIM + c) = y;
        if (ob) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(y.x, y.y), p1 = __floats2bfloat162_rn(y.z, y.w);
            uint2 pk; pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(ob + row * U_DIM + c) = pk;
        }
    }
}

// ---------------------------------------------------------------- fused softmax attention, 197 tokens x 64 dims per head
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int AT_TP = 208;          // tokens padded to 13 tiles of 16
constexpr int AT_LD = 72;           // smem row stride (bf16): 144 B keeps ldmatrix conflict-free
constexpr int AT_WARPS = 7;         // 13 query tiles over 7 warps
constexpr int AT_SMEM = 3 * AT_TP * AT_LD * 2;

// One chunk of NT key tiles (8 keys each) starting at key tile NT0: scores, online-softmax update of (m, l, o), P V.
template <int NT0, int NT>
__device__ __forceinline__ void attn_chunk(const uint32_t (&qa)[4][4], const bf16* sk, const bf16* sv, int lane, int tq, float sl2,
                                           float& m0, float& m1, float& l0, float& l1, float (&o)[8][4]) {
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
            uint32_t kb[4];
            ldsm_x4(kb, sk + ((NT0 + nt) * 8 + (lane & 7)) * AT_LD + kp * 32 + (lane >> 3) * 8);
            mma_bf16_16816(s[nt], qa[kp * 2], kb[0], kb[1]);
            mma_bf16_16816(s[nt], qa[kp * 2 + 1], kb[2], kb[3]);
        }
    }
    float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int c = (NT0 + nt) * 8 + tq * 2;
        if (c >= U_TOK) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
        if (c + 1 >= U_TOK) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
        c0 = fmaxf(c0, fmaxf(s[nt][0], s[nt][1])); c1 = fmaxf(c1, fmaxf(s[nt][2], s[nt][3]));
    }
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1)); c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
    c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1)); c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
    const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);            // every chunk holds at least one real key, so n0/n1 are finite
    const float a0 = exp2f((m0 - n0) * sl2), a1 = exp2f((m1 - n1) * sl2);     // exp2(-inf) = 0 on the first chunk
    m0 = n0; m1 = n1;
    l0 *= a0; l1 *= a1;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= a0; o[nt][1] *= a0; o[nt][2] *= a1; o[nt][3] *= a1; }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        s[nt][0] = exp2f((s[nt][0] - m0) * sl2); s[nt][1] = exp2f((s[nt][1] - m0) * sl2);
        s[nt][2] = exp2f((s[nt][2] - m1) * sl2); s[nt][3] = exp2f((s[nt][3] - m1) * sl2);
        l0 += s[nt][0] + s[nt][1]; l1 += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) {
        uint32_t pa[4];
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(s[2 * j][0], s[2 * j][1]); pa[0] = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(s[2 * j][2], s[2 * j][3]); pa[1] = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(s[2 * j + 1][0], s[2 * j + 1][1]); pa[2] = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(s[2 * j + 1][2], s[2 * j + 1][3]); pa[3] = *reinterpret_cast<uint32_t*>(&t);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t vb[4];
            ldsm_x4_t(vb, sv + ((NT0 / 2 + j) * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * AT_LD + np * 16 + (lane >> 4) * 8);
            mma_bf16_16816(o[np * 2], pa, vb[0], vb[1]);
            mma_bf16_16816(o[np * 2 + 1], pa, vb[2], vb[3]);
        }
    }
}

// qkv: bf16 [B*197, 3*1024] = [q | k | v] x [head][64] per token (timm reshape (B,N,3,H,hd)); out: bf16 [B*197, 1024]
__global__ void __launch_bounds__(AT_WARPS * 32, 2) uni_attention_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t at_smem[];
    bf16* sq_ = reinterpret_cast<bf16*>(at_smem);
    bf16* sk = sq_ + AT_TP * AT_LD;
    bf16* sv = sk + AT_TP * AT_LD;
    const int b = blockIdx.x / U_HEADS, hd = blockIdx.x % U_HEADS;
    const bf16* base = qkv + (long long)b * U_TOK * 3 * U_DIM + hd * U_HD;
    for (int i = threadIdx.x; i < AT_TP * 8 * 3; i += blockDim.x) {
        const int which = i / (AT_TP * 8); const int r = (i / 8) % AT_TP; const int c8 = (i % 8) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < U_TOK) v = *reinterpret_cast<const uint4*>(base + (long long)r * 3 * U_DIM + which * U_DIM + c8);
        *reinterpret_cast<uint4*>((which == 0 ? sq_ : which == 1 ? sk : sv) + r * AT_LD + c8) = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;
    const float sl2 = 0.125f * 1.44269504088896340736f;        // softmax scale 64^-0.5 folded with log2(e)
    for (int tile = warp; tile < AT_TP / 16; tile += AT_WARPS) {
        const int r0 = tile * 16;
        uint32_t qa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) ldsm_x4(qa[ks], sq_ + (r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * AT_LD + ks * 16 + (lane >> 4) * 8);
        float o[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
        // keys in two chunks (112 + 96) with an online softmax: halves the score registers so that two CTAs fit on an SM
        attn_chunk<0, 14>(qa, sk, sv, lane, tq, sl2, m0, m1, l0, l1, o);
        attn_chunk<14, 12>(qa, sk, sv, lane, tq, sl2, m0, m1, l0, l1, o);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float i0 = 1.0f / l0, i1 = 1.0f / l1;
        bf16* ob = out + (long long)b * U_TOK * U_DIM + hd * U_HD;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const int c = nt * 8 + tq * 2;
            if (r0 + g < U_TOK) *reinterpret_cast<__nv_bfloat162*>(ob + (long long)(r0 + g) * U_DIM + c) = __floats2bfloat162_rn(o[nt][0] * i0, o[nt][1] * i0);
            if (r0 + g + 8 < U_TOK) *reinterpret_cast<__nv_bfloat162*>(ob + (long long)(r0 + g + 8) * U_DIM + c) = __floats2bfloat162_rn(o[nt][2] * i1, o[nt][3] * i1);
        }
    }
}

// ---------------------------------------------------------------- softmax attention on tcgen05 / TMEM (default)
// One work item = (image, head): S = Q K^T and O = P V on the 5th-generation tensor cores, the softmax between them in registers.
//   warp 16     TMA producer: the head's q / k / v slices of the qkv matrix, one 64 x 208-row box each (rows past the image belong
//               to the next image or are zero-filled past the end: finite, and masked below), double-buffered across items
//   warp 17     MMA issuer: per 128-query tile S[128 x 208] = Q K^T (4 k-steps, both operands K-major in shared memory), later
//               O[128 x 64] = P V (13 k-steps): P is read from TENSOR MEMORY (A operand in TMEM), V is the MN-major B operand
//   warps 0-15  softmax, 8 warps per query tile: a TMEM lane is a query row; the two warps of a lane quadrant split the row's key
//               columns (0..103 / 104..207) - four warps per scheduler hide the tcgen05.ld round trips.  Pass 1: partial row
//               maxima, exchanged through shared memory; pass 2: exp2, partial sums, bf16 pack into registers; once every warp
//               of the tile has read its scores, P overwrites the S columns (packed, 104 columns) and the second MMA starts;
//               afterwards each warp scales its 32 output columns by 1 / sum and stores 64 contiguous bytes per row
// TMEM: two 256-column regions (one per query tile): S in [0, 208), P in [0, 104), O in [128, 192).
constexpr int ATC_THREADS = 18 * 32;
constexpr int ATC_KP = 208;                                   // keys padded to 13 k-steps of 16
constexpr int ATC_Q_BYTES = 256 * 128, ATC_KV_BYTES = ATC_KP * 128;
constexpr int ATC_BUF_BYTES = ATC_Q_BYTES + 2 * ATC_KV_BYTES;
constexpr int ATC_SMEM = 2 * ATC_BUF_BYTES + 2 * 2 * 2 * 128 * 4 + 256;

__global__ void __launch_bounds__(ATC_THREADS, 1) uni_attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, bf16* __restrict__ out, int batch) {
    extern __shared__ __align__(1024) uint8_t atc_smem[];
    float* s_max = reinterpret_cast<float*>(atc_smem + 2 * ATC_BUF_BYTES);      // [tile][half][row]
    float* s_sum = s_max + 2 * 2 * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_sum + 2 * 2 * 128);
    uint64_t* kv_full = bars;          // [2] TMA bytes
    uint64_t* kv_empty = bars + 2;     // [2] tcgen05.commit after the item's last MMA
    uint64_t* s_full = bars + 4;       // [tile] scores ready
    uint64_t* p_full = bars + 6;       // [tile] probabilities written (the tile's active warps)
    uint64_t* o_full = bars + 8;       // [tile] output accumulator ready
    uint64_t* o_free = bars + 10;      // [tile] output read: the region may take the next item's scores
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nitems = batch * U_HEADS;
    // rows 197..255 of the second query tile are padding: its last lane quadrant (rows 224..255) has no work at all
    constexpr int ACTIVE_Q1 = (U_TOK - 128 + 31) / 32;                // 3 active quadrants in tile 1
    if (threadIdx.x == 0) {
        if (smem_u32(atc_smem) & 1023u) { printf("sequoia_b200: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        tma_prefetch_desc(&map_qkv);
        for (int i = 0; i < 2; ++i) {
            const uint32_t nw = i == 0 ? 8 : 2 * ACTIVE_Q1;
            mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); mbar_init(&s_full[i], 1); mbar_init(&p_full[i], nw);
            mbar_init(&o_full[i], 1); mbar_init(&o_free[i], nw);
        }
        mbar_fence_init();
    }
    if (warp == 17) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 16) {
        if (elect_one()) {
            int n = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
                const int buf = n & 1, b = item / U_HEADS, hd = item - b * U_HEADS;
                mbar_wait(&kv_empty[buf], ((n >> 1) & 1) ^ 1);
                uint8_t* sq = atc_smem + buf * ATC_BUF_BYTES;
                mbar_expect_tx(&kv_full[buf], 3 * ATC_KV_BYTES);
                tma_load_2d(&map_qkv, &kv_full[buf], sq, hd * U_HD, b * U_TOK);
                tma_load_2d(&map_qkv, &kv_full[buf], sq + ATC_Q_BYTES, U_DIM + hd * U_HD, b * U_TOK);
                tma_load_2d(&map_qkv, &kv_full[buf], sq + ATC_Q_BYTES + ATC_KV_BYTES, 2 * U_DIM + hd * U_HD, b * U_TOK);
            }
        }
        __syncwarp();
    } else if (warp == 17) {
        const uint32_t idesc_s = make_idesc_bf16(ATC_KP, 0, 0, 128), idesc_o = make_idesc_bf16(U_HD, 0, 1, 128);
        int n = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
            const int buf = n & 1;
            const uint32_t sq = smem_u32(atc_smem + buf * ATC_BUF_BYTES), sk = sq + ATC_Q_BYTES, sv = sk + ATC_KV_BYTES;
            mbar_wait(&kv_full[buf], (n >> 1) & 1);
            for (int t = 0; t < 2; ++t) {
                mbar_wait(&o_free[t], (n & 1) ^ 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + t * 256, make_smem_desc(sq + t * 16384 + k * 32, 1024, 0), make_smem_desc(sk + k * 32, 1024, 0), idesc_s, k ? 1u : 0u);
                    umma_commit(&s_full[t]);
                }
                __syncwarp();
            }
            for (int t = 0; t < 2; ++t) {
                mbar_wait(&p_full[t], n & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < ATC_KP / 16; ++k)
                        umma_bf16_ts(tmem_base + t * 256 + 128, tmem_base + t * 256 + k * 8, make_smem_desc(sv + k * 2048, 1024, 8192), idesc_o, k ? 1u : 0u);
                    umma_commit(&o_full[t]);
                    if (t == 1) umma_commit(&kv_empty[buf]);
                }
                __syncwarp();
            }
        }
    } else if ((warp >> 3) == 0 || (warp & 3) < ACTIVE_Q1) {
        const int t = warp >> 3, h = (warp >> 2) & 1, q = warp & 3, r = q * 32 + lane, row = t * 128 + r;
        const int bar_threads = (t == 0 ? 8 : 2 * ACTIVE_Q1) * 32;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 256;
        const uint32_t sbase = taddr + h * 104;                        // this warp's score columns
        const float sl2 = 0.125f * 1.44269504088896340736f;           // softmax scale 64^-0.5 folded with log2(e)
        float* my_max = s_max + (t * 2 + h) * 128 + r; const float* other_max = s_max + (t * 2 + (h ^ 1)) * 128 + r;
        float* my_sum = s_sum + (t * 2 + h) * 128 + r; const float* other_sum = s_sum + (t * 2 + (h ^ 1)) * 128 + r;
        constexpr int LASTV = U_TOK - 104 - 64;                       // valid columns in the last 32-column piece of the second half (29)
        int n = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x, ++n) {
            const int b = item / U_HEADS, hd = item - b * U_HEADS;
            mbar_wait(&s_full[t], n & 1);
            tc_fence_after();
            float v[32];
            // ---- pass 1: maximum of this warp's half of the row
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                tmem_ld32(sbase + 32 * c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) if (!(h == 1 && c == 2 && i >= LASTV)) mx = fmaxf(mx, v[i]);
            }
            if (h == 0) {
                tmem_ld8(sbase + 96, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) mx = fmaxf(mx, v[i]);
            }
            *my_max = mx;
            named_bar_sync(1 + t, bar_threads);
            const float msl = fmaxf(mx, *other_max) * sl2;
            // ---- pass 2: exp2, partial row sum, bf16 pack (kept in registers until every warp of the tile has read its scores)
            uint32_t pk[52];
            float l = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                tmem_ld32(sbase + 32 * c, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const bool m0 = h == 1 && c == 2 && 2 * i >= LASTV, m1 = h == 1 && c == 2 && 2 * i + 1 >= LASTV;
                    const float e0 = m0 ? 0.f : ex2_approx(fmaf(v[2 * i], sl2, -msl)), e1 = m1 ? 0.f : ex2_approx(fmaf(v[2 * i + 1], sl2, -msl));
                    l += e0 + e1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
                    pk[16 * c + i] = *reinterpret_cast<uint32_t*>(&h2);
                }
            }
            if (h == 0) {
                tmem_ld8(sbase + 96, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float e0 = ex2_approx(fmaf(v[2 * i], sl2, -msl)), e1 = ex2_approx(fmaf(v[2 * i + 1], sl2, -msl));
                    l += e0 + e1;
                    __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
                    pk[48 + i] = *reinterpret_cast<uint32_t*>(&h2);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) pk[48 + i] = 0u;               // keys 200..207: padding
            }
            *my_sum = l;
            tc_fence_before();
            named_bar_sync(1 + t, bar_threads);                            // all scores of the tile are in registers: P may overwrite S
            tc_fence_after();
            const uint32_t pbase = taddr + h * 52;
            tmem_st32(pbase, pk);
            tmem_st16(pbase + 32, pk + 32);
            tmem_st4(pbase + 48, pk + 48);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
            const float inv = 1.0f / (l + *other_sum);
            mbar_wait(&o_full[t], n & 1);
            tc_fence_after();
            tmem_ld32(taddr + 128 + 32 * h, v);
            tmem_ld_wait();
            if (row < U_TOK) {
                uint4* dst = reinterpret_cast<uint4*>(out + ((long long)b * U_TOK + row) * U_DIM + hd * U_HD + 32 * h);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint4 w; __nv_bfloat162 h2;
                    h2 = __floats2bfloat162_rn(v[8 * k] * inv, v[8 * k + 1] * inv); w.x = *reinterpret_cast<uint32_t*>(&h2);
                    h2 = __floats2bfloat162_rn(v[8 * k + 2] * inv, v[8 * k + 3] * inv); w.y = *reinterpret_cast<uint32_t*>(&h2);
                    h2 = __floats2bfloat162_rn(v[8 * k + 4] * inv, v[8 * k + 5] * inv); w.z = *reinterpret_cast<uint32_t*>(&h2);
                    h2 = __floats2bfloat162_rn(v[8 * k + 6] * inv, v[8 * k + 7] * inv); w.w = *reinterpret_cast<uint32_t*>(&h2);
                    dst[k] = w;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[t]);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 17) tmem_dealloc(tmem_base, 512);
}

static int uni_attention_tc(const bf16* qkv, bf16* attn, int batch, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(uni_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM) != cudaSuccess) {
            set_error("uni_attention: cannot raise the dynamic shared memory limit"); (void)cudaGetLastError(); return -1;
        }
        attr = true;
    }
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)3 * U_DIM, (cuuint64_t)batch * U_TOK}, strides[1] = {(cuuint64_t)3 * U_DIM * 2};
    cuuint32_t box[2] = {(cuuint32_t)U_HD, (cuuint32_t)ATC_KP}, estr[2] = {1, 1};
    if (encode_map(&map, qkv, 2, dims, strides, box, estr)) return -1;
    const int nitems = batch * U_HEADS;
    const int grid = nitems < num_sms() ? nitems : num_sms();
    uni_attention_tc_kernel<<<grid, ATC_THREADS, ATC_SMEM, st>>>(map, attn, batch);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("uni_attention: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

struct UniWs { size_t col, patch, x, h, qkv, attn, u, total; };
static void uni_ws_layout(int batch, UniWs* w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) / 1024 * 1024; return o; };
    const size_t M = (size_t)batch * U_TOK, P = (size_t)batch * 196;
    w->col = take(P * U_PK * 2); w->patch = take(P * U_DIM * 4); w->x = take(M * U_DIM * 4); w->h = take(M * U_DIM * 2);
    w->qkv = take(M * 3 * U_DIM * 2); w->attn = take(M * U_DIM * 2); w->u = take(M * U_MLP * 2);
    w->total = off;
}

static int uni_gemm(int M, int N, int K, const bf16* a, const bf16* w, const float* bias, const float* res, float* out_f32, bf16* out_bf, int act,
                    cudaStream_t st) {
    // bf16 -> bf16 GEMMs with a bias (qkv, fc1 + GELU: 7/12 of the model's FLOPs) run on the CTA-pair kernel with the TMA epilogue
    static const int use_cg = getenv("SQ_UNI_CONVGEMM") ? atoi(getenv("SQ_UNI_CONVGEMM")) : 1;
    if (use_cg && out_bf && !out_f32 && !res && bias && (act == ACT_NONE || (act == ACT_GELU && N % 256 == 0)) && N % 64 == 0 && K % 64 == 0) {
        ConvGemmArgs c; memset(&c, 0, sizeof(c));
        c.M = M; c.N = N; c.K = K; c.A = a; c.lda = K; c.W = w; c.bias = bias; c.out = out_bf; c.relu = act == ACT_GELU ? 2 : 0;
        if (convgemm_supported(c)) return convgemm_launch(c, st);
    }
    GemmArgs g; memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K; g.nterms = 1;
    g.A.hi = a; g.A.ld = K; g.B.hi = w; g.B.ld = K;
    g.e.bias = bias; g.e.res_f32 = res; g.e.ld_res = N; g.e.out_f32 = out_f32; g.e.ld_f32 = N; g.e.out_hi = out_bf; g.e.ld_bf = N;
    g.e.act = act; g.e.alpha = 1.0f; g.e.rowbias_div = 1;
    g.epi_conv_pref = out_bf != nullptr;       // bf16 outputs go through the coalesced (staged) epilogue
    return gemm_launch(g, st);
}

}  // namespace sq

using namespace sq;

extern "C" {

int sq_vitl16_num_tensors(int depth) { return 4 + 14 * depth + 2; }
long long sq_vitl16_packed_weight_elems(int depth) { if (depth < 1 || depth > 64) return -1; UniW w; UniV v; uni_layout(depth, &w, &v); return w.total; }
long long sq_vitl16_packed_vec_elems(int depth) { if (depth < 1 || depth > 64) return -1; UniW w; UniV v; uni_layout(depth, &w, &v); return v.total; }

int sq_vitl16_prepack(const void* const* tensors, int depth, void* packed_w, float* packed_v, void* stream) {
    if (depth < 1 || depth > 64) { set_error("vitl16: depth %d out of range", depth); return -1; }
    const int nt = 4 + 14 * depth + 2;
    for (int i = 0; i < nt; ++i) if (!tensors[i]) { set_error("vitl16 prepack: null tensor %d", i); return -1; }
    UniW W; UniV V; uni_layout(depth, &W, &V);
    cudaStream_t st = (cudaStream_t)stream;
    bf16* pw = (bf16*)packed_w;
    const float* const* t = reinterpret_cast<const float* const*>(tensors);
    auto cast = [&](const float* w, const float* scale, long long rows, int cols, long long off) {
        uni_cast_kernel<<<1184, 256, 0, st>>>(w, scale, rows, cols, pw + off);
    };
    auto vec = [&](const float* v, const float* scale, int n, long long off) { uni_vec_kernel<<<(n + 255) / 256, 256, 0, st>>>(v, scale, n, packed_v + off); };
    vec(t[0], nullptr, U_DIM, V.cls); vec(t[1], nullptr, U_TOK * U_DIM, V.pos);
    cast(t[2], nullptr, U_DIM, U_PK, W.pe); vec(t[3], nullptr, U_DIM, V.pb);
    for (int i = 0; i < depth; ++i) {
        const float* const* b = t + 4 + 14 * i;   // norm1.w, norm1.b, qkv.w, qkv.b, proj.w, proj.b, ls1, norm2.w, norm2.b, fc1.w, fc1.b, fc2.w, fc2.b, ls2
        vec(b[0], nullptr, U_DIM, V.blk[i][0]); vec(b[1], nullptr, U_DIM, V.blk[i][1]);
        cast(b[2], nullptr, 3 * U_DIM, U_DIM, W.blk[i][0]); vec(b[3], nullptr, 3 * U_DIM, V.blk[i][2]);
        cast(b[4], b[6], U_DIM, U_DIM, W.blk[i][1]); vec(b[5], b[6], U_DIM, V.blk[i][3]);          // ls1 folded
        vec(b[7], nullptr, U_DIM, V.blk[i][4]); vec(b[8], nullptr, U_DIM, V.blk[i][5]);
        cast(b[9], nullptr, U_MLP, U_DIM, W.blk[i][2]); vec(b[10], nullptr, U_MLP, V.blk[i][6]);
        cast(b[11], b[13], U_DIM, U_MLP, W.blk[i][3]); vec(b[12], b[13], U_DIM, V.blk[i][7]);      // ls2 folded
    }
    vec(t[nt - 2], nullptr, U_DIM, V.ng); vec(t[nt - 1], nullptr, U_DIM, V.nb);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("vitl16 prepack: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

size_t sq_vitl16_workspace_bytes(int batch) { if (batch <= 0) return 0; UniWs w; uni_ws_layout(batch, &w); return w.total; }

int sq_vitl16_extract(const void* input, int input_kind, int batch, int depth, const void* packed_w, const float* packed_v, float* features,
                      void* workspace, size_t workspace_bytes, void* stream) {
    if (batch <= 0) return 0;
    if (depth < 1 || depth > 64) { set_error("vitl16: depth %d out of range", depth); return -1; }
    if (!input || !packed_w || !packed_v || !features) { set_error("vitl16_extract: null pointer"); return -1; }
    UniWs L; uni_ws_layout(batch, &L);
    if (!workspace || workspace_bytes < L.total) { set_error("vitl16_extract: workspace %zu < %zu", workspace_bytes, L.total); return -1; }
    UniW W; UniV V; uni_layout(depth, &W, &V);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    const bf16* pw = (const bf16*)packed_w;
    bf16* col = (bf16*)(ws + L.col); float* patch = (float*)(ws + L.patch); float* x = (float*)(ws + L.x);
    bf16* h = (bf16*)(ws + L.h); bf16* qkv = (bf16*)(ws + L.qkv); bf16* attn = (bf16*)(ws + L.attn); bf16* u = (bf16*)(ws + L.u);
    const int M = batch * U_TOK, P = batch * 196;
    static const int attn_tc = getenv("SQ_UNI_ATTN_TC") ? atoi(getenv("SQ_UNI_ATTN_TC")) : 1;      // 0: the mma.sync kernel (A/B runs)
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(uni_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM); attr = true; }
    uni_patchify_kernel<<<1184, 256, 0, st>>>(input, input_kind, batch, col);
    if (uni_gemm(P, U_DIM, U_PK, col, pw + W.pe, packed_v + V.pb, nullptr, patch, nullptr, ACT_NONE, st)) return -1;
    uni_assemble_kernel<<<1184, 256, 0, st>>>(patch, packed_v + V.cls, packed_v + V.pos, batch, x);
    for (int i = 0; i < depth; ++i) {
        const long long* vb = V.blk[i];
        uni_ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, U_DIM, packed_v + vb[0], packed_v + vb[1], M, h, nullptr);
        if (uni_gemm(M, 3 * U_DIM, U_DIM, h, pw + W.blk[i][0], packed_v + vb[2], nullptr, nullptr, qkv, ACT_NONE, st)) return -1;
        if (attn_tc) { if (uni_attention_tc(qkv, attn, batch, st)) return -1; }
        else uni_attention_kernel<<<batch * U_HEADS, AT_WARPS * 32, AT_SMEM, st>>>(qkv, attn);
        if (uni_gemm(M, U_DIM, U_DIM, attn, pw + W.blk[i][1], packed_v + vb[3], x, x, nullptr, ACT_NONE, st)) return -1;
        uni_ln_kernel<<<(M + 7) / 8, 256, 0, st>>>(x, U_DIM, packed_v + vb[4], packed_v + vb[5], M, h, nullptr);
        if (uni_gemm(M, U_MLP, U_DIM, h, pw + W.blk[i][2], packed_v + vb[6], nullptr, nullptr, u, ACT_GELU, st)) return -1;
        if (uni_gemm(M, U_DIM, U_MLP, u, pw + W.blk[i][3], packed_v + vb[7], x, x, nullptr, ACT_NONE, st)) return -1;
    }
    // final LayerNorm of the cls token of every image (global_pool='token', num_classes=0)
    uni_ln_kernel<<<(batch + 7) / 8, 256, 0, st>>>(x, (long long)U_TOK * U_DIM, packed_v + V.ng, packed_v + V.nb, batch, nullptr, features);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("vitl16_extract: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // extern "C"
