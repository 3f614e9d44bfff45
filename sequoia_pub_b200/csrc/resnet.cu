// ResNet-50 feature extractor (reference: src/resnet.py:96-170 `ResNet`, `Bottleneck`, `forward_extract`;
// caller pre_processing/compute_features_hdf5.py:49-51,116-123).
//
// Data layout in HBM: activations NHWC bf16, weights [Cout][R][S][Cin] bf16 with the eval-mode BatchNorm scale
// folded in, one fp32 shift per output channel.  Per batch (256-px tiles): `stem_tc_kernel` (this file: uint8 -> /255 -> normalise,
// conv1 7x7/2, shift, ReLU, 3x3/2 max-pool on tcgen05; `stem_fused_kernel` for other tile sizes), then the 16 bottlenecks as
// launches of `convgemm_kernel` (convgemm.cuh: CTA-pair tcgen05 implicit GEMM, epilogue = shift + residual + ReLU by TMA) with
// layer 1's conv2 + conv3 as one `bneck_l1_kernel` (fusedconv.cuh); the top-left 7x7 average pool (fact 4) is fused into the last
// convolution's epilogue.  The round-1 chain (im2col stem + gemm.cuh kernels, SQ_STEM_FUSED=0 / SQ_CONVGEMM=0) is kept for A/B runs;
// the split-precision mode at the end of the file carries every tensor as bf16 hi / lo planes.
#include "convgemm.cuh"
#include "fusedconv.cuh"
#include "../../include/sequoia_b200.h"
#include <stdlib.h>

namespace sq {

struct ConvSpec { int cin, cout, k, stride, pad; long long w_off, s_off; };

struct ResNetPlan {
    ConvSpec conv[53];
    long long w_elems, s_elems;
};

static const int STEM_K = 192;   // 7*7*3 = 147 padded to 3 k-blocks

static const ResNetPlan& plan() {
    static ResNetPlan p;
    static bool init = false;
    if (!init) {
        int n = 0; long long w = 0, s = 0;
        auto add = [&](int cin, int cout, int k, int stride, int pad) {
            ConvSpec c; c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; c.pad = pad; c.w_off = w; c.s_off = s;
            w += (n == 0) ? (long long)cout * STEM_K : (long long)cout * k * k * cin;
            s += cout;
            p.conv[n++] = c;
        };
        add(3, 64, 7, 2, 3);
        const int planes[4] = {64, 128, 256, 512}, blocks[4] = {3, 4, 6, 3}, strides[4] = {1, 2, 2, 2};
        int inpl = 64;
        for (int st = 0; st < 4; ++st)
            for (int b = 0; b < blocks[st]; ++b) {
                const int stride = b == 0 ? strides[st] : 1;
                add(inpl, planes[st], 1, 1, 0);
                add(planes[st], planes[st], 3, stride, 1);
                add(planes[st], planes[st] * 4, 1, 1, 0);
                if (b == 0) add(inpl, planes[st] * 4, 1, stride, 0);
                inpl = planes[st] * 4;
            }
        p.w_elems = w; p.s_elems = s;
        init = true;
    }
    return p;
}

// OIHW fp32 + BN(gamma, beta, mean, var) -> [O][R][S][I] bf16 (scale folded) + fp32 shift
__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps, int cout, int cin, int k,
                               int kpad /*0 = dense*/, bf16* __restrict__ wp, float* __restrict__ shift, bf16* __restrict__ wlo = nullptr) {
    const int kk = k * k * cin;
    const int row = kpad ? kpad : kk;
    const long long n = (long long)cout * row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i / row); const int j = (int)(i - (long long)o * row);
        const float sc = gamma[o] / sqrtf(var[o] + eps);
        float v = 0.f;
        if (kpad) {
            // stem: K index = r * 24 + (s * 3 + c), 21 taps per filter row padded to 24 (keeps the im2col rows 4-byte aligned)
            const int r = j / 24, q = j - r * 24;
            if (r < k && q < k * cin) { const int s = q / cin, c = q - s * cin; v = w[(((long long)o * cin + c) * k + r) * k + s] * sc; }
        } else if (j < kk) {
            const int c = j % cin; const int rs = j / cin; const int s = rs % k; const int r = rs / k;
            v = w[(((long long)o * cin + c) * k + r) * k + s] * sc;
        }
        const bf16 hi = __float2bfloat16_rn(v);
        wp[i] = hi;
        if (wlo) wlo[i] = __float2bfloat16_rn(v - __bfloat162float(hi));       // split-precision mode: w = hi + lo
        if (j == 0) shift[o] = beta[o] - mean[o] * sc;
    }
}

// Stem im2col: one block = 32 consecutive output pixels of one output row.
// kind 0: uint8 NHWC raw patch, applies x/255 then (x-mean)/std  (compute_features_hdf5.py:49-51)
// kind 1: fp32 NCHW, already normalised (the tensor the reference hands to forward_extract)
// The 7 input rows x 69 columns x 3 channels the block needs are staged as bf16 in shared memory (row stride 208, even, so
// every 2-element group is 4-byte aligned); for a fixed filter row r the 21 (s, c) taps of output pixel p are CONTIGUOUS
// there (offset r*208 + 6p), so an im2col row is 7 segments of 24 = 21 taps + 3 zeros, then zeros up to K = 192.
__global__ void __launch_bounds__(256) stem_im2col_kernel(const void* __restrict__ in, int kind, int H, int W, int Ho, int Wo,
                                                          bf16* __restrict__ col) {
    constexpr int TW = 32, IW = TW * 2 + 5, ROW = 208;      // 69 input columns -> 207 values (+1 pad) per staged row
    __shared__ __align__(16) bf16 tile[7 * ROW + 8];
    const int tiles_w = Wo / TW;
    int b = blockIdx.x;
    const int tw = b % tiles_w; b /= tiles_w;
    const int oh = b % Ho; const int img = b / Ho;
    const int ow0 = tw * TW;
    const int ih0 = oh * 2 - 3, iw0 = ow0 * 2 - 3;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    for (int i = threadIdx.x; i < 7 * ROW; i += 256) {
        const int r = i / ROW; const int xc = i - r * ROW; const int x = xc / 3; const int c = xc - x * 3;
        const int ih = ih0 + r, iw = iw0 + x;
        float v = 0.f;
        if (xc < IW * 3 && ih >= 0 && ih < H && iw >= 0 && iw < W) {
            if (kind == 0) {
                const uint8_t u = reinterpret_cast<const uint8_t*>(in)[(((long long)img * H + ih) * W + iw) * 3 + c];
                v = (static_cast<float>(u) / 255.0f - mean[c]) / stdv[c];
            } else {
                v = reinterpret_cast<const float*>(in)[(((long long)img * 3 + c) * H + ih) * W + iw];
            }
        }
        tile[i] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    bf16* dst = col + (((long long)img * Ho + oh) * Wo + ow0) * STEM_K;
    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile);
    for (int i = threadIdx.x; i < TW * (STEM_K / 8); i += 256) {
        const int p = i / (STEM_K / 8); const int k8 = (i - p * (STEM_K / 8)) * 8;
        const int r = k8 / 24, q0 = k8 - r * 24;            // q0 in {0, 8, 16}
        uint4 pack = make_uint4(0u, 0u, 0u, 0u);
        if (r < 7) {
            const int e = (r * ROW + p * 6 + q0) >> 1;      // 32-bit word index (r*208 + 6p + q0 is even)
            pack.x = t32[e]; pack.y = t32[e + 1]; pack.z = t32[e + 2]; pack.w = t32[e + 3];
            if (q0 == 16) { pack.z &= 0x0000ffffu; pack.w = 0u; }      // taps 16..20 valid, 21..23 are zero padding
        }
        *reinterpret_cast<uint4*>(dst + (long long)p * STEM_K + k8) = pack;
    }
}

// 3x3 stride-2 pad-1 max pool, NHWC bf16, 8 channels per thread
__global__ void maxpool3x3s2_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int batch, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
    const long long n = (long long)batch * Ho * Wo * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8); long long t = i / C8;
        const int ow = (int)(t % Wo); t /= Wo; const int oh = (int)(t % Ho); const int img = (int)(t / Ho);
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int r = 0; r < 3; ++r) {
            const int ih = oh * 2 - 1 + r; if (ih < 0 || ih >= H) continue;
            for (int s = 0; s < 3; ++s) {
                const int iw = ow * 2 - 1 + s; if (iw < 0 || iw >= W) continue;
                const uint4 v = *reinterpret_cast<const uint4*>(in + (((long long)img * H + ih) * W + iw) * C + c8 * 8);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); m[2 * j] = fmaxf(m[2 * j], f.x); m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y); }
            }
        }
        uint4 o; __nv_bfloat162* oh2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh2[j] = __floats2bfloat162_rn(m[2 * j], m[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + (((long long)img * Ho + oh) * Wo + ow) * C + c8 * 8) = o;
    }
}

// ---------------------------------------------------------------------------------------------------- fused stem
// conv1 7x7/2 (+ folded BN shift, ReLU) + 3x3/2 max-pool in ONE kernel, straight from the uint8 tile (R4 fused too):
// no im2col buffer (403 MB written + read per batch of 64) and no full-resolution stem map (134 MB written + read).
// A CTA owns an 8x8 tile of the POOLED map = 17x17 conv outputs = 39x39 input pixels, staged in shared memory as
// normalised bf16 [row][x*3 + c].  For a fixed filter row r the 21 (s, c) taps of a conv pixel are contiguous there
// (offset 6*ox), exactly the K layout of the packed stem weights (K index = r*24 + s*3 + c, 3 zero weights per filter
// row), so the im2col matrix is never materialised: mma.sync A fragments are 32-bit loads from the staged tile, B
// fragments come from the weight slab with ldmatrix.  M = 289 conv pixels (19 m16 tiles over 10 warps), N = 64,
// K = 176 (11 k16 steps; taps 168..175 multiply zero weights).  The conv tile goes to shared memory as bf16 after shift +
// ReLU, then the 3x3/2 max-pool writes whole 128-byte channel rows.  ~20 GFLOP per batch: legacy mma.sync is enough
// here, the kernel is bound by staging and the pool, not by the tensor pipe.
constexpr int ST_THREADS = 320;
constexpr int ST_IROW = 120, ST_IROWS = 40;     // staged input: 39 rows (+1 over-read row) x (39*3 = 117 -> 120) bf16
constexpr int ST_WLD = 184;                     // weight slab [64][176] row stride: 368 B, conflict-free ldmatrix
constexpr int ST_CLD = 72;                      // conv tile [289][64] row stride: 144 B, conflict-free fragment stores
constexpr int ST_KSTEPS = 11;
constexpr int ST_CPIX = 17 * 17;
constexpr int ST_WORDS = 31;                    // aligned 32-bit words that cover the 117 (+3) bytes of a staged uint8 row
constexpr size_t ST_SMEM = (size_t)(ST_IROWS * ST_IROW + 64 * ST_WLD + ST_CPIX * ST_CLD + 3 * 256) * 2;

__device__ __forceinline__ void st_ldsm_x4(uint32_t (&r)[4], const void* p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void st_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// kind 0: uint8 NHWC; 1: fp32 NCHW normalised; 2: uint8 NHWC with a 4-byte aligned base and W % 4 == 0 (word loads + LUT)
__global__ void __launch_bounds__(ST_THREADS, 2) stem_fused_kernel(const void* __restrict__ in, int kind, int batch, int H, int W,
                                                                   const bf16* __restrict__ wpk, const float* __restrict__ shift,
                                                                   bf16* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    bf16* s_in = reinterpret_cast<bf16*>(st_smem);
    bf16* s_w = s_in + ST_IROWS * ST_IROW;
    bf16* s_c = s_w + 64 * ST_WLD;
    bf16* s_lut = s_c + ST_CPIX * ST_CLD;                         // normalised value of every (channel, byte): bit-identical to the formula
    for (int i = threadIdx.x; i < 3 * 256; i += ST_THREADS) {
        const int c = i >> 8;
        const float mu = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f), sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
        s_lut[i] = __float2bfloat16_rn((static_cast<float>(i & 255) / 255.0f - mu) / sd);
    }
    for (int i = threadIdx.x; i < ST_IROW; i += ST_THREADS) s_in[39 * ST_IROW + i] = __float2bfloat16_rn(0.f);   // over-read row
    // conv1 7x7/2 pad 3 and max-pool 3x3/2 pad 1 output sizes (any H, W; partial 8x8 pooled tiles are masked on store)
    const int Hc = (H - 1) / 2 + 1, Wc = (W - 1) / 2 + 1, Hp = (Hc - 1) / 2 + 1, Wp = (Wc - 1) / 2 + 1;
    const int tiles_x = (Wp + 7) / 8, tiles_y = (Hp + 7) / 8;
    const int ntiles = batch * tiles_y * tiles_x;
    for (int i = threadIdx.x; i < 64 * (ST_KSTEPS * 2); i += ST_THREADS) {          // [64][192] packed -> [64][176] slab
        const int n = i / (ST_KSTEPS * 2), q = i - n * (ST_KSTEPS * 2);
        *reinterpret_cast<uint4*>(s_w + n * ST_WLD + q * 8) = *reinterpret_cast<const uint4*>(wpk + n * STEM_K + q * 8);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    float sh[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { sh[nt][0] = shift[nt * 8 + 2 * t]; sh[nt][1] = shift[nt * 8 + 2 * t + 1]; }
    const uint32_t* in32 = reinterpret_cast<const uint32_t*>(s_in);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int tx = tile % tiles_x; const int rest = tile / tiles_x; const int ty = rest % tiles_y; const int img = rest / tiles_y;
        const int cy0 = 16 * ty - 1, cx0 = 16 * tx - 1;          // conv-map origin of the 17x17 tile
        const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;          // input origin of the 39x39 tile
        __syncthreads();                                         // the previous tile's readers of s_in / s_c are done
        if (kind == 2) {
            const int row_bytes = W * 3, xb0 = ix0 * 3;          // byte offset of tile column 0 inside an image row (may be < 0)
            const int a0 = (xb0 >> 2) << 2, mis = xb0 - a0;      // word-aligned start; rows start on word boundaries (W % 4 == 0)
            const uint8_t* img_base = reinterpret_cast<const uint8_t*>(in) + (long long)img * H * row_bytes;
            for (int i = threadIdx.x; i < 39 * ST_WORDS; i += ST_THREADS) {
                const int r = i / ST_WORDS, wi = i - r * ST_WORDS;
                const int ih = iy0 + r, wb = a0 + 4 * wi;
                const bool ok = ih >= 0 && ih < H && wb >= 0 && wb < row_bytes;     // a word is entirely inside or outside the row
                uint32_t word = 0u;
                if (ok) word = *reinterpret_cast<const uint32_t*>(img_base + (long long)ih * row_bytes + wb);
                const int c0 = ok ? wb % 3 : 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int xc = 4 * wi + b - mis;
                    if (xc >= 0 && xc < ST_IROW) {
                        int c = c0 + b; c = c >= 3 ? c - 3 : c; c = c >= 3 ? c - 3 : c;
                        s_in[r * ST_IROW + xc] = ok ? s_lut[(c << 8) + ((word >> (8 * b)) & 255u)] : __float2bfloat16_rn(0.f);
                    }
                }
            }
        } else
        for (int i = threadIdx.x; i < ST_IROWS * ST_IROW; i += ST_THREADS) {
            const int r = i / ST_IROW; const int xc = i - r * ST_IROW; const int x = xc / 3; const int c = xc - x * 3;
            const int ih = iy0 + r, iw = ix0 + x;
            float v = 0.f;
            if (r < 39 && x < 39 && ih >= 0 && ih < H && iw >= 0 && iw < W) {
                if (kind == 0) {
                    const uint8_t u = reinterpret_cast<const uint8_t*>(in)[(((long long)img * H + ih) * W + iw) * 3 + c];
                    const float mu = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f), sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
                    v = (static_cast<float>(u) / 255.0f - mu) / sd;
                } else {
                    v = reinterpret_cast<const float*>(in)[(((long long)img * 3 + c) * H + ih) * W + iw];
                }
            }
            s_in[i] = __float2bfloat16_rn(v);
        }
        __syncthreads();
        for (int mt = warp; mt < (ST_CPIX + 15) / 16; mt += ST_THREADS / 32) {
            const int p0 = mt * 16 + g, p1 = p0 + 8;
            const int q0 = p0 < ST_CPIX ? p0 : ST_CPIX - 1, q1 = p1 < ST_CPIX ? p1 : ST_CPIX - 1;
            const int oy0 = q0 / 17, ox0 = q0 - oy0 * 17, oy1 = q1 / 17, ox1 = q1 - oy1 * 17;
            const int b0 = oy0 * 2 * (ST_IROW / 2) + ox0 * 3 + t, b1 = oy1 * 2 * (ST_IROW / 2) + ox1 * 3 + t;     // 32-bit word offsets
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < ST_KSTEPS; ++ks) {
                const int kA = ks * 16, kB = ks * 16 + 8;        // each 8-wide half lies inside one 24-wide filter-row segment
                const int offA = (kA / 24) * (ST_IROW / 2) + (kA % 24) / 2, offB = (kB / 24) * (ST_IROW / 2) + (kB % 24) / 2;
                uint32_t a[4];
                a[0] = in32[b0 + offA]; a[1] = in32[b1 + offA]; a[2] = in32[b0 + offB]; a[3] = in32[b1 + offB];
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) {
                    uint32_t b[4];   // matrices: (n-tile 2jn, k-half 0), (2jn, 1), (2jn+1, 0), (2jn+1, 1)
                    st_ldsm_x4(b, s_w + ((jn * 2 + (lane >> 4)) * 8 + (lane & 7)) * ST_WLD + ks * 16 + ((lane >> 3) & 1) * 8);
                    st_mma(acc[2 * jn], a, b[0], b[1]);
                    st_mma(acc[2 * jn + 1], a, b[2], b[3]);
                }
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int p = half ? p1 : p0;
                if (p < ST_CPIX) {
                    const int oy = half ? oy1 : oy0, ox = half ? ox1 : ox0;
                    const int cy = cy0 + oy, cx = cx0 + ox;
                    const bool valid = cy >= 0 && cy < Hc && cx >= 0 && cx < Wc;      // outside the conv map: excluded from the max (ReLU output >= 0)
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
                        const float v0 = valid ? fmaxf(acc[nt][2 * half] + sh[nt][0], 0.f) : 0.f;
                        const float v1 = valid ? fmaxf(acc[nt][2 * half + 1] + sh[nt][1], 0.f) : 0.f;
                        *reinterpret_cast<__nv_bfloat162*>(s_c + p * ST_CLD + nt * 8 + 2 * t) = __floats2bfloat162_rn(v0, v1);
                    }
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 64 * 8; i += ST_THREADS) {
            const int c8 = i & 7, pp = i >> 3, ply = pp >> 3, plx = pp & 7;
            if (8 * ty + ply >= Hp || 8 * tx + plx >= Wp) continue;
            uint4 m = *reinterpret_cast<const uint4*>(s_c + ((2 * ply) * 17 + 2 * plx) * ST_CLD + c8 * 8);
            __nv_bfloat162* mh = reinterpret_cast<__nv_bfloat162*>(&m);
#pragma unroll
            for (int rs = 1; rs < 9; ++rs) {
                const uint4 v = *reinterpret_cast<const uint4*>(s_c + ((2 * ply + rs / 3) * 17 + 2 * plx + rs % 3) * ST_CLD + c8 * 8);
                const __nv_bfloat162* vh = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) mh[j] = __hmax2(mh[j], vh[j]);
            }
            *reinterpret_cast<uint4*>(out + ((((long long)img * Hp + 8 * ty + ply) * Wp + 8 * tx + plx) * 64 + c8 * 8)) = m;
        }
    }
}

static int stem_fused_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SQ_STEM_FUSED"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

static int launch_stem_fused(const void* input, int kind, int batch, int H, int W, const bf16* wpk, const float* shift, bf16* out, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(stem_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM) != cudaSuccess) {
            set_error("stem: cannot raise the dynamic shared memory limit"); (void)cudaGetLastError(); return -1;
        }
        attr_set = true;
    }
    const int Hp_ = ((H - 1) / 2) / 2 + 1, Wp_ = ((W - 1) / 2) / 2 + 1;
    const int ntiles = batch * ((Hp_ + 7) / 8) * ((Wp_ + 7) / 8);
    int grid = 2 * num_sms(); if (grid > ntiles) grid = ntiles;
    if (kind == 0 && (reinterpret_cast<uintptr_t>(input) & 3) == 0 && W % 4 == 0) kind = 2;
    stem_fused_kernel<<<grid, ST_THREADS, ST_SMEM, st>>>(input, kind, batch, H, W, wpk, shift, out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("stem_fused: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- tcgen05 stem (256 x 256 tiles)
// conv1 7x7/2 + BN shift + ReLU + 3x3/2 max-pool on the 5th-generation tensor cores, warp-specialised (no block-wide barrier in
// the steady state).  One MMA tile = ONE ROW of the conv map (128 pixels x 64 channels).  The K dimension is organised by INPUT
// ROW, not by filter row: k-block P_j ("pair j") holds, for every pixel of the row, the 24 (s, c) taps of input row 2j-1 in
// columns 0..23 and those of input row 2j in columns 32..55.  The taps of a pixel in one input row are 24 CONTIGUOUS values of
// the normalised row (6*ox values in; the last 3 meet zero weights), i.e. twelve 32-bit shared-memory loads per row.
// Conv row cy = P_{cy-1} x W[r0 r1] + P_cy x W[r2 r3] + P_{cy+1} x W[r4 r5] + (first half of) P_{cy+2} x W[r6]: every pair is
// built ONCE and used by four conv rows.
//  * The pairs live in TENSOR MEMORY (the A operand of tcgen05.mma is read from TMEM): a builder thread is a pixel = a TMEM lane,
//    its 2 x 12 words go to the pair's 32 columns with tcgen05.st; shared memory only feeds the 64-channel weight tile (2 KB per
//    MMA instead of 6 KB - with A in a shared-memory ring, the first version, the kernel was bound by shared-memory bandwidth).
//  * The BN shift rides in the MMA: columns 24, 25 of every pair are 1.0 and the weight block of filter row r6 carries
//    bf16 hi / lo of the shift there (the other blocks have zeros), so the epilogue is one cvt.rn.relu.bf16x2 per two values.
//  * Normalisation (byte / 255 - mean) / std is ONE fma per value with constants that reproduce the bf16 result of the reference
//    formula for all 256 byte values (exhaustive check; the byte-wise path keeps the table and the tests compare the bits).
//   warps 0-7   epilogue: tcgen05.ld (32 lanes x 32 channels each), ReLU + bf16 pack; the VERTICAL 3-max of the pool lives in
//               registers (a TMEM lane is a pixel column, the same thread sees it in every conv row); every second conv row the
//               vertical maxima go to shared memory and the horizontal stride-2 3-max writes the pooled row (8 KB, contiguous)
//   warps 8-15  two builder groups (pairs dealt round-robin): uint8 rows -> cp.async into a 4-deep ring (three pairs ahead) -> staged
//               normalised bf16 rows -> registers -> tcgen05.st -> mbarrier
//   warp 16     one elected thread issues 14 tcgen05.mma (M 128, N 64, K 16) per conv row into a double-buffered TMEM accumulator;
//               tcgen05.commit releases the oldest pair and publishes the accumulator
// A CTA walks bands of 4 pooled rows (9 conv rows, 12 pairs) of one image.
constexpr int S2_NG = 2;                          // builder groups of 128 threads, pairs dealt round-robin (3 groups measured: no gain, 80 registers + spills)
constexpr int S2_MMA_WARP = 8 + 4 * S2_NG;
constexpr int S2_THREADS = (S2_MMA_WARP + 1) * 32;
constexpr int S2_P = 4;                          // pooled rows per band
constexpr int S2_R = 8;                          // pair ring slots (32 TMEM columns each)
constexpr int S2_D = 4;                          // raw-row ring depth per builder group (pairs in flight from global memory)
constexpr int S2_B_BYTES = 4 * 8192;             // weights: 4 k-blocks x [64 channels][64 k]
constexpr int S2_STG_LD = 800;                   // staged row: 3 zero pixels + 256 pixels + 3 zero pixels = 786 values (+ pad)
constexpr int S2_V_LD = 144;                     // vertical-max row: 128 pixels x 64 bf16, row pitch 144 B (conflict-free 16-byte accesses)
constexpr int S2_V_BYTES = 128 * S2_V_LD;
constexpr int S2_RAW_BYTES = 2 * 768;            // the two uint8 rows of a pair
constexpr uint32_t S2_TMEM_COLS = 512;           // 2 x 64 accumulator columns + 8 pairs x 32 columns = 384 -> 512
constexpr size_t S2_SMEM = (size_t)S2_B_BYTES + 2 * S2_V_BYTES + S2_NG * 2 * 2 * S2_STG_LD * 2 + S2_NG * S2_D * S2_RAW_BYTES + 3 * 256 * 2 + 256;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// {lo, hi} -> bf16x2 with ReLU fused into the conversion
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

struct S2Item { int img, py0, cyb, nrows; };
__device__ __forceinline__ S2Item s2_item(int item) {
    S2Item t; t.img = item / (64 / S2_P); t.py0 = (item - t.img * (64 / S2_P)) * S2_P;
    t.cyb = t.py0 == 0 ? 0 : 2 * t.py0 - 1; t.nrows = 2 * t.py0 + 2 * S2_P - t.cyb;
    return t;
}

// kind 0: uint8 NHWC (any alignment, byte loads); 1: fp32 NCHW normalised; 2: uint8 NHWC with a 16-byte aligned base (cp.async)
__global__ void __launch_bounds__(S2_THREADS, 1) stem_tc_kernel(const void* __restrict__ in, int kind, int batch, const bf16* __restrict__ wpk,
                                                                const float* __restrict__ shift, bf16* __restrict__ out, int prof) {
    constexpr int H = 256, W = 256, HP = 64;
    long long pc[6] = {0, 0, 0, 0, 0, 0}; const long long pt0 = clock64();          // SQ_STEM_PROF=1: cycles each role of block 0 spends waiting
    extern __shared__ __align__(1024) uint8_t s2_smem[];
    uint8_t* sB = s2_smem;                                              // [4][64][128 B], 128B swizzle
    uint8_t* sV = sB + S2_B_BYTES;                                      // [2][128][144 B]
    bf16* sStg = reinterpret_cast<bf16*>(sV + 2 * S2_V_BYTES);          // [group][buffer][row][S2_STG_LD]
    uint8_t* sRaw = reinterpret_cast<uint8_t*>(sStg + S2_NG * 2 * 2 * S2_STG_LD);   // [group][S2_D][2 x 768 B]
    bf16* sLut = reinterpret_cast<bf16*>(sRaw + S2_NG * S2_D * S2_RAW_BYTES);       // [3][256] (byte-wise path only)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sLut + 3 * 256);
    uint64_t* pair_full = bars;                                         // [R] 4 builder warps
    uint64_t* pair_free = bars + S2_R;                                  // [R] tcgen05.commit
    uint64_t* tmem_full = bars + 2 * S2_R;                              // [2]
    uint64_t* tmem_empty = tmem_full + 2;                               // [2] 8 epilogue warps
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nitems = batch * (HP / S2_P);

    if (kind == 0)
        for (int i = tid; i < 3 * 256; i += S2_THREADS) {
            const int c = i >> 8;
            const float mu = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f), sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
            sLut[i] = __float2bfloat16_rn((static_cast<float>(i & 255) / 255.0f - mu) / sd);
        }
    // weights [64][7 x 24] -> four K-major 128B-swizzled k-blocks [r0 r1] [r2 r3] [r4 r5] [r6 0], each filter row padded to 32 columns;
    // columns 24, 25 of the r6 block: bf16 hi / lo of the BN shift (they meet the 1.0 columns of the pairs)
    for (int i = tid; i < 4 * 64 * 8; i += S2_THREADS) {
        const int ch = i & 7, n = (i >> 3) & 63, b = i >> 9, r = 2 * b + (ch >> 2), part = ch & 3;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < 7 && part < 3) v = *reinterpret_cast<const uint4*>(wpk + n * STEM_K + r * 24 + part * 8);
        if (r == 6 && part == 3) {
            const float sh = shift[n];
            const bf16 hi = __float2bfloat16_rn(sh), lo = __float2bfloat16_rn(sh - __bfloat162float(hi));
            v.x = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
        }
        *reinterpret_cast<uint4*>(sB + b * 8192 + n * 128 + ((ch ^ (n & 7)) << 4)) = v;
    }
    // the zero pixels left / right of a staged row are written here once and never again
    for (int i = tid; i < S2_NG * 2 * 2 * S2_STG_LD; i += S2_THREADS) sStg[i] = __float2bfloat16_rn(0.f);
    if (tid == 0) {
        for (int i = 0; i < S2_R; ++i) { mbar_init(&pair_full[i], 4); mbar_init(&pair_free[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
        mbar_fence_init();
    }
    if (warp == S2_MMA_WARP) tmem_alloc(tmem_ptr, S2_TMEM_COLS);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 8) {
        // ================================================================ epilogue + max-pool
        const int q = warp & 3, h = warp >> 2, m = q * 32 + lane, et = tid;          // et: 0..255
        const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * 32;
        int n = 0, vb = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const S2Item it = s2_item(item);
            uint32_t acc[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = 0u;                        // ReLU outputs are >= 0: 0 is the identity of the max
            for (int c = 0; c < it.nrows; ++c, ++n) {
                const int cy = it.cyb + c, ab = n & 1;
                if (prof) { const long long w0 = clock64(); mbar_wait(&tmem_full[ab], (n >> 1) & 1); pc[0] += clock64() - w0; }
                else mbar_wait(&tmem_full[ab], (n >> 1) & 1);
                tc_fence_after();
                float v[32];
                tmem_ld32(trow + ab * 64, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[ab]);
                uint32_t p[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) p[k] = pack_relu_bf16x2(v[2 * k], v[2 * k + 1]);
                if (cy & 1) {
                    const int py = (cy - 1) >> 1;                    // conv rows 2py-1, 2py, 2py+1 are complete
                    if (py >= it.py0) {
                        uint8_t* vrow = sV + (vb & 1) * S2_V_BYTES + m * S2_V_LD + h * 64;
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            *reinterpret_cast<uint4*>(vrow + 16 * k4) = make_uint4(hmax2_u32(acc[4 * k4], p[4 * k4]), hmax2_u32(acc[4 * k4 + 1], p[4 * k4 + 1]),
                                                                                   hmax2_u32(acc[4 * k4 + 2], p[4 * k4 + 2]), hmax2_u32(acc[4 * k4 + 3], p[4 * k4 + 3]));
                        if (prof) { const long long w0 = clock64(); named_bar_sync(1, 256); pc[1] += clock64() - w0; }
                        else named_bar_sync(1, 256);
                        const uint8_t* vbuf = sV + (vb & 1) * S2_V_BYTES;
                        bf16* orow = out + (((size_t)it.img * HP + py) * HP) * 64;
#pragma unroll
                        for (int rep = 0; rep < 2; ++rep) {
                            const int idx = et + rep * 256, c8 = idx & 7, px = idx >> 3;
                            uint4 t1 = *reinterpret_cast<const uint4*>(vbuf + (2 * px) * S2_V_LD + c8 * 16);
                            const uint4 t2 = *reinterpret_cast<const uint4*>(vbuf + (2 * px + 1) * S2_V_LD + c8 * 16);
                            uint4 t0 = t1;
                            if (px > 0) t0 = *reinterpret_cast<const uint4*>(vbuf + (2 * px - 1) * S2_V_LD + c8 * 16);
                            t1.x = hmax2_u32(hmax2_u32(t1.x, t0.x), t2.x); t1.y = hmax2_u32(hmax2_u32(t1.y, t0.y), t2.y);
                            t1.z = hmax2_u32(hmax2_u32(t1.z, t0.z), t2.z); t1.w = hmax2_u32(hmax2_u32(t1.w, t0.w), t2.w);
                            *reinterpret_cast<uint4*>(orow + px * 64 + c8 * 8) = t1;
                        }
                        ++vb;
                    }
#pragma unroll
                    for (int k = 0; k < 16; ++k) acc[k] = p[k];
                } else {
#pragma unroll
                    for (int k = 0; k < 16; ++k) acc[k] = hmax2_u32(acc[k], p[k]);
                }
            }
        }
    } else if (warp < S2_MMA_WARP) {
        // ================================================================ pair builders (S2_NG groups of 128 threads, pairs dealt round-robin)
        const int g = (warp - 8) >> 2, t = tid - 256 - g * 128;                  // t: pixel column of this thread = TMEM lane, 0..127
        bf16* stg_base = sStg + g * (2 * 2 * S2_STG_LD);
        uint8_t* raw_base = sRaw + g * (S2_D * S2_RAW_BYTES);
        struct Cur { int item, i, np, j0, img; int seq; };
        auto decode = [&](Cur& c) { const S2Item it = s2_item(c.item); c.np = it.nrows + 3; c.j0 = it.cyb - 1; c.img = it.img; };
        auto advance = [&](Cur& c) {
            ++c.seq;
            if (++c.i == c.np) { c.item += gridDim.x; c.i = 0; if (c.item < nitems) decode(c); }
        };
        auto advance_ng = [&](Cur& c) { for (int a = 0; a < S2_NG && c.item < nitems; ++a) advance(c); };    // this group's next pair
        // cp.async: thread t < 96 copies 16-byte chunk t of the pair's 2 x 768 raw bytes
        const int my_row = t >= 48 ? 1 : 0, my_chunk = t - my_row * 48;
        auto issue_copy = [&](const Cur& c, int rslot) {
            const int iy = 2 * (c.j0 + c.i) - 1 + my_row;
            if (t < 96 && iy >= 0 && iy < H)
                cp_async16(raw_base + rslot * S2_RAW_BYTES + t * 16, reinterpret_cast<const uint8_t*>(in) + ((size_t)c.img * H + iy) * (W * 3) + my_chunk * 16);
        };
        // (byte / 255 - mean) / std as ONE fma per value: for these three (a, b) pairs bf16(fma(byte, a, b)) equals bf16 of the
        // reference formula for all 256 byte values (checked exhaustively; the byte-wise path below keeps the table, and
        // tests/test_resnet_gpu.py asserts both paths give the same bits)
        const float na0 = __uint_as_float(0x3c8c48f6u), nb0 = __uint_as_float(0xc0078b9cu);
        const float na1 = __uint_as_float(0x3c8f6a98u), nb1 = __uint_as_float(0xc00248eeu);
        const float na2 = __uint_as_float(0x3c8ec798u), nb2 = __uint_as_float(0xbfe6f7ddu);
        Cur cur; cur.item = blockIdx.x; cur.i = 0; cur.seq = 0; cur.np = 0; cur.j0 = 0; cur.img = 0;
        if (cur.item < nitems) decode(cur);
        for (int a = 0; a < g && cur.item < nitems; ++a) advance(cur);
        Cur pf = cur;
        if (kind == 2) {
#pragma unroll
            for (int d = 0; d < S2_D - 1; ++d) {
                if (pf.item < nitems) { issue_copy(pf, d); advance_ng(pf); }
                cp_async_commit();
            }
            cp_async_wait<S2_D - 2>();                     // this thread's chunk of the first pair has landed ...
            named_bar_sync(2 + g, 128);                    // ... and so have the other threads' chunks
        }
        int itn = 0;
        while (cur.item < nitems) {
            bf16* stg = stg_base + (itn & 1) * (2 * S2_STG_LD);
            const int j = cur.j0 + cur.i;
            const long long tb0 = prof ? clock64() : 0;
            if (kind == 2) {
                if (pf.item < nitems) { issue_copy(pf, (itn + S2_D - 1) % S2_D); advance_ng(pf); }
                cp_async_commit();
                // thread t converts bytes 6t+1 .. 6t+6 of both rows: three aligned bf16 pairs per row (staged index = byte + 9), channels
                // (1,2) (0,1) (2,0) whatever t is; consecutive threads write words 3 apart (conflict-free).  byte -> float through the
                // 2^23 trick (one PRMT + one FADD on the FMA pipe instead of I2F on the quarter-rate conversion unit; same exact value),
                // loads unconditional and both rows interleaved (rows outside the image are zeroed by a select)
                const uint8_t* rp = raw_base + (itn % S2_D) * S2_RAW_BYTES + 6 * t;
                uint32_t* sw = reinterpret_cast<uint32_t*>(stg) + 5 + 3 * t;
                uint32_t b1[2], h2[2], h4[2], b6[2], b0[2];
#pragma unroll
                for (int row = 0; row < 2; ++row) {
                    const uint8_t* r8 = rp + row * 768;
                    b1[row] = r8[1]; h2[row] = *reinterpret_cast<const uint16_t*>(r8 + 2); h4[row] = *reinterpret_cast<const uint16_t*>(r8 + 4); b6[row] = r8[6];
                    b0[row] = t == 0 ? r8[0] : 0u;
                }
                auto b2f = [](uint32_t x, uint32_t sel) { return __uint_as_float(__byte_perm(x, 0x4B000000u, sel)) - 8388608.0f; };   // 2^23 + byte, exact
#pragma unroll
                for (int row = 0; row < 2; ++row) {
                    const int iy = 2 * j - 1 + row;
                    const bool valid = iy >= 0 && iy < H;
                    const float v1 = fmaf(b2f(b1[row], 0x7540u), na1, nb1), v2 = fmaf(b2f(h2[row], 0x7540u), na2, nb2), v3 = fmaf(b2f(h2[row], 0x7541u), na0, nb0);
                    const float v4 = fmaf(b2f(h4[row], 0x7540u), na1, nb1), v5 = fmaf(b2f(h4[row], 0x7541u), na2, nb2);
                    const float v6 = t == 127 ? 0.f : fmaf(b2f(b6[row], 0x7540u), na0, nb0);
                    __nv_bfloat162 q0 = __floats2bfloat162_rn(v1, v2), q1 = __floats2bfloat162_rn(v3, v4), q2 = __floats2bfloat162_rn(v5, v6);
                    uint32_t* d = sw + row * (S2_STG_LD / 2);
                    d[0] = valid ? *reinterpret_cast<uint32_t*>(&q0) : 0u;
                    d[1] = valid ? *reinterpret_cast<uint32_t*>(&q1) : 0u;
                    d[2] = valid ? *reinterpret_cast<uint32_t*>(&q2) : 0u;
                    if (t == 0) {                                            // staged values 8 (zero pad), 9 (byte 0)
                        __nv_bfloat162 qe = __floats2bfloat162_rn(0.f, fmaf(b2f(b0[row], 0x7540u), na0, nb0));
                        d[-1] = valid ? *reinterpret_cast<uint32_t*>(&qe) : 0u;
                    }
                }
                if (prof) { const long long tb1 = clock64(); cp_async_wait<S2_D - 2>(); const long long tb2 = clock64(); pc[2] += tb1 - tb0; pc[3] += tb2 - tb1; }
                else
                cp_async_wait<S2_D - 2>();                 // the NEXT pair's chunk of this thread has landed; the barrier below publishes it
            } else if (kind == 0) {
                // uint8 tiles whose base is not 16-byte aligned: byte loads, the 3 x 256 table of the reference formula, same staged values
                const uint8_t* src = reinterpret_cast<const uint8_t*>(in) + (size_t)cur.img * H * (W * 3);
                for (int i = t; i < 2 * 3 * W; i += 128) {
                    const int row = i / (3 * W), e = i - row * (3 * W), iy = 2 * j - 1 + row;
                    stg[row * S2_STG_LD + 9 + e] = (iy >= 0 && iy < H) ? sLut[((e % 3) << 8) + src[(size_t)iy * (W * 3) + e]] : __float2bfloat16_rn(0.f);
                }
            } else {
                // fp32 NCHW, already normalised (the tensor the reference hands to forward_extract)
                const float* src = reinterpret_cast<const float*>(in);
                for (int i = t; i < 2 * 3 * W; i += 128) {
                    const int row = i / (3 * W), e = i - row * (3 * W), c = e / W, x = e - c * W, iy = 2 * j - 1 + row;
                    float v = 0.f;
                    if (iy >= 0 && iy < H) v = src[(((size_t)cur.img * 3 + c) * H + iy) * W + x];
                    stg[row * S2_STG_LD + 9 + x * 3 + c] = __float2bfloat16_rn(v);
                }
            }
            const int slot = cur.seq % S2_R;
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(stg) + 3 * t;             // pixel t: values 6t .. 6t+23 of a staged row
            if (prof) { const long long w0 = clock64(); named_bar_sync(2 + g, 128); pc[0] += clock64() - w0; }
            else named_bar_sync(2 + g, 128);
            uint32_t r0[16], r1[16];
#pragma unroll
            for (int u = 0; u < 12; ++u) { r0[u] = s32[u]; r1[u] = s32[S2_STG_LD / 2 + u]; }
#pragma unroll
            for (int u = 12; u < 16; ++u) { r0[u] = 0u; r1[u] = 0u; }
            r0[12] = 0x3f803f80u;                          // columns 24, 25 = 1.0: they meet shift hi / lo in the r6 weight block, zeros elsewhere
            if (prof) { const long long w1 = clock64(); mbar_wait(&pair_free[slot], ((cur.seq / S2_R) & 1) ^ 1); pc[1] += clock64() - w1; }
            else mbar_wait(&pair_free[slot], ((cur.seq / S2_R) & 1) ^ 1);
            tc_fence_after();
            const long long tb3 = prof ? clock64() : 0;
            const uint32_t ta = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 128 + slot * 32;
            tmem_st16(ta, r0);
            tmem_st16(ta + 16, r1);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pair_full[slot]);
            advance_ng(cur);
            ++itn;
            if (prof) { const long long tb4 = clock64(); pc[4] += tb4 - tb3; pc[5] += tb4 - tb0; }
        }
    } else {
        // ================================================================ MMA issuer
        const uint32_t idesc = make_idesc_bf16(64, 0, 0, 128);
        const uint32_t b0 = smem_u32(sB);
        int n = 0, seq0 = 0, next_wait = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const S2Item it = s2_item(item);
            for (int c = 0; c < it.nrows; ++c, ++n) {
                const int ab = n & 1;
                const long long w0 = prof ? clock64() : 0;
                mbar_wait(&tmem_empty[ab], ((n >> 1) & 1) ^ 1);
                const long long w1 = prof ? clock64() : 0;
                while (next_wait <= seq0 + c + 3) { mbar_wait(&pair_full[next_wait % S2_R], (next_wait / S2_R) & 1); ++next_wait; }
                if (prof) { pc[0] += w1 - w0; pc[1] += clock64() - w1; }
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int slot = (seq0 + c + b) % S2_R;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (b == 3 && k >= 2) break;
                            umma_bf16_ts(tmem_base + ab * 64, tmem_base + 128 + slot * 32 + k * 8, make_smem_desc(b0 + b * 8192 + k * 32, 1024, 0), idesc,
                                         (b | k) ? 1u : 0u);
                        }
                    }
                    umma_commit(&tmem_full[ab]);
                    umma_commit(&pair_free[(seq0 + c) % S2_R]);
                    if (c == it.nrows - 1)
                        for (int b = 1; b < 4; ++b) umma_commit(&pair_free[(seq0 + c + b) % S2_R]);
                }
                __syncwarp();
            }
            seq0 += it.nrows + 3;
        }
    }
    if (prof && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 8 || warp == 12 || warp == S2_MMA_WARP))
        printf("stem_tc prof block 0 warp %2d (%s): total %lld cycles; waits: %s %lld, %s %lld\n", warp, warp == 0 ? "epilogue" : (warp == S2_MMA_WARP ? "mma" : "builder"),
               clock64() - pt0, warp == 0 ? "accumulator" : (warp == S2_MMA_WARP ? "tmem_empty" : "stage barrier"), pc[0],
               warp == 0 ? "pool barrier" : (warp == S2_MMA_WARP ? "pair_full" : "pair_free"), pc[1]);
    if (prof && blockIdx.x == 0 && lane == 0 && (warp == 8 || warp == 12))
        printf("stem_tc prof builder warp %d: issue+convert %lld, cp.async wait %lld, tcgen05.st+arrive+advance %lld, loop total %lld\n", warp, pc[2], pc[3], pc[4], pc[5]);
    tc_fence_before();
    __syncthreads();
    if (warp == S2_MMA_WARP) tmem_dealloc(tmem_base, S2_TMEM_COLS);
}

// SQ_STEM_TC=0 keeps the mma.sync fused stem for 256 x 256 tiles too (A/B runs)
static int stem_tc_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SQ_STEM_TC"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

// 256 x 256 tiles (uint8 NHWC or normalised fp32 NCHW); other sizes keep stem_fused_kernel
static bool stem_tc_supported(int kind, int H, int W) { return H == 256 && W == 256 && (kind == 0 || kind == 1); }

static int launch_stem_tc(const void* input, int kind, int batch, const bf16* wpk, const float* shift, bf16* out, cudaStream_t st) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_SMEM) != cudaSuccess) {
            set_error("stem_tc: cannot raise the dynamic shared memory limit"); (void)cudaGetLastError(); return -1;
        }
        attr_set = true;
    }
    const int nitems = batch * (64 / S2_P);
    int grid = num_sms(); if (grid > nitems) grid = nitems;
    if (kind == 0 && (reinterpret_cast<uintptr_t>(input) & 15) == 0) kind = 2;      // 16-byte cp.async chunks: every row is 768 B
    static const int prof = getenv("SQ_STEM_PROF") ? atoi(getenv("SQ_STEM_PROF")) : 0;
    stem_tc_kernel<<<grid, S2_THREADS, S2_SMEM, st>>>(input, kind, batch, wpk, shift, out, prof);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("stem_tc: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

// AvgPool2d(7) on an HfxWf map with Hf,Wf in [7,13]: mean of the top-left 7x7 window (src/resnet.py:110,166; fact 4)
__global__ void avgpool7_kernel(const float* __restrict__ in, float* __restrict__ out, int batch, int Hf, int Wf, int C) {
    const long long n = (long long)batch * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const int img = (int)(i / C);
        float acc = 0.f;
        for (int h = 0; h < 7; ++h)
            for (int w = 0; w < 7; ++w) acc += in[(((long long)img * Hf + h) * Wf + w) * C + c];
        out[i] = acc * (1.0f / 49.0f);
    }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct ResNetWs {
    size_t col, stem, big[3], small[2], fmap, im2col, total;
};

static inline int conv_out(int x, int k, int stride, int pad) { return (x + 2 * pad - k) / stride + 1; }

static ResNetWs ws_layout(int batch, int H, int W) {
    ResNetWs w; size_t off = 0;
    const size_t Ho = conv_out(H, 7, 2, 3), Wo = conv_out(W, 7, 2, 3), Hp = conv_out((int)Ho, 3, 2, 1), Wp = conv_out((int)Wo, 3, 2, 1);
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    const bool fused = stem_fused_enabled() != 0;
    w.col = take(fused ? 0 : (size_t)batch * Ho * Wo * STEM_K * 2);
    w.stem = take(fused ? 0 : (size_t)batch * Ho * Wo * 64 * 2);
    for (int i = 0; i < 3; ++i) w.big[i] = take((size_t)batch * Hp * Wp * 256 * 2);
    for (int i = 0; i < 2; ++i) w.small[i] = take((size_t)batch * Hp * Wp * 128 * 2);
    // final map (only materialised when the average pool is not fused) and the im2col scratch of the stride-2 convolutions on maps
    // the TMA boxes cannot tile (inputs other than 256x256)
    int h = (int)Hp, wd = (int)Wp; size_t im2col = 0;
    {
        const int planes[4] = {64, 128, 256, 512}, strides[4] = {1, 2, 2, 2};
        int inpl = 64;
        for (int st = 0; st < 4; ++st) {
            if (strides[st] == 2) {
                const int ho = conv_out(h, 3, 2, 1), wo = conv_out(wd, 3, 2, 1);
                ConvGeom g; memset(&g, 0, sizeof(g)); g.enabled = 1; g.Ho = ho; g.Wo = wo;
                if (!conv_pertap_geometry_ok(g)) {
                    const size_t a = (size_t)batch * ho * wo * 9 * planes[st] * 2, b = (size_t)batch * ho * wo * inpl * 2;
                    im2col = a > im2col ? a : im2col; im2col = b > im2col ? b : im2col;
                }
                h = ho; wd = wo;
            }
            inpl = planes[st] * 4;
        }
    }
    w.fmap = take((size_t)batch * h * wd * 2048 * 4);
    w.im2col = take(im2col);
    w.total = off;
    return w;
}

// one instantiation per (tile width, CTA group); ring depth 3
int convgemm_dispatch(int bn, int cg, int halo, const CUtensorMap* maps, const CgParams& kp, int grid, cudaStream_t st) {
    if (kp.relu == 2) return convgemm_launch_inst<256, 2, 3, 0, 1>(maps, kp, grid, st);       // GELU epilogue (UNI fc1): one instantiation
#define SQ_CG_CASE(BN, CG) if (bn == BN && cg == CG && !halo) return convgemm_launch_inst<BN, CG, 3, 0>(maps, kp, grid, st);
#define SQ_CG_HALO(BN, CG, D) if (bn == BN && cg == CG && halo) return convgemm_launch_inst<BN, CG, D, 1>(maps, kp, grid, st);
    SQ_CG_CASE(64, 2) SQ_CG_CASE(128, 2) SQ_CG_CASE(256, 2) SQ_CG_CASE(64, 1) SQ_CG_CASE(128, 1)
    SQ_CG_HALO(64, 2, 3) SQ_CG_HALO(128, 2, 3) SQ_CG_HALO(256, 2, 2) SQ_CG_HALO(64, 1, 3) SQ_CG_HALO(128, 1, 3)     // 256-wide: ring depth 2 leaves room for 5 weight stages
#undef SQ_CG_HALO
#undef SQ_CG_CASE
    set_error("convgemm: no kernel for block_n %d cta_group %d", bn, cg);
    return -1;
}

static int pool_fused_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SQ_POOL_FUSED"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

static int bneck_fuse_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SQ_BNECK_FUSE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

static int bneck_ds_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SQ_BNECK_DS"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

static int convgemm_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SQ_CONVGEMM"); v = (e && e[0] == '0') ? 0 : 1; }
    return v;
}

static int run_conv(const ConvSpec& c, const bf16* wbase, const float* sbase, const bf16* in, int batch, int H, int W, bf16* out_bf,
                    float* out_f32, const bf16* res, bool relu, cudaStream_t st, int* Ho_out, int* Wo_out, float* pool_out = nullptr,
                    bf16* scratch = nullptr, size_t scratch_bytes = 0) {
    const int Ho = (H + 2 * c.pad - c.k) / c.stride + 1, Wo = (W + 2 * c.pad - c.k) / c.stride + 1;
    if (convgemm_enabled() && ((out_bf && !out_f32) || pool_out)) {
        ConvGemmArgs a; memset(&a, 0, sizeof(a));
        a.pool_out = pool_out; a.pool_batch = batch; a.scratch = scratch; a.scratch_bytes = scratch_bytes;
        a.M = batch * Ho * Wo; a.N = c.cout; a.K = c.k * c.k * c.cin;
        a.A = in; a.lda = c.cin; a.W = wbase + c.w_off; a.bias = sbase + c.s_off; a.res = res; a.out = out_bf; a.relu = relu ? 1 : 0;
        a.conv.enabled = (c.k == 1 && c.stride == 1) ? 0 : 1; a.conv.batch = batch; a.conv.H = H; a.conv.W = W; a.conv.C = c.cin; a.conv.Ho = Ho; a.conv.Wo = Wo;
        a.conv.R = c.k; a.conv.S = c.k; a.conv.stride = c.stride; a.conv.pad = c.pad;
        if (convgemm_supported(a)) { *Ho_out = Ho; *Wo_out = Wo; return convgemm_launch(a, st); }
    }
    GemmArgs g; memset(&g, 0, sizeof(g));
    g.M = batch * Ho * Wo; g.N = c.cout; g.K = c.k * c.k * c.cin;
    g.A.hi = in; g.A.ld = c.cin;
    g.B.hi = wbase + c.w_off; g.B.ld = g.K;
    g.nterms = 1;
    // 1x1 stride-1 convolutions are plain GEMMs over [pixels, Cin] (2-D tensor maps: cheaper to issue than 4-D boxes)
    g.epi_conv_pref = 1;
    g.conv.enabled = (c.k == 1 && c.stride == 1) ? 0 : 1; g.conv.batch = batch; g.conv.H = H; g.conv.W = W; g.conv.C = c.cin; g.conv.Ho = Ho; g.conv.Wo = Wo;
    g.conv.R = c.k; g.conv.S = c.k; g.conv.stride = c.stride; g.conv.pad = c.pad;
    g.e.bias = sbase + c.s_off;
    g.e.out_hi = out_bf; g.e.ld_bf = c.cout;
    g.e.out_f32 = out_f32; g.e.ld_f32 = c.cout;
    g.e.res_bf = res; g.e.ld_res = c.cout;
    g.e.act = relu ? ACT_RELU : ACT_NONE;
    g.e.alpha = 1.0f;
    g.e.rowbias_div = 1;
    *Ho_out = Ho; *Wo_out = Wo;
    return gemm_launch(g, st);
}

}  // namespace sq

using namespace sq;

extern "C" {

int sq_resnet50_num_convs(void) { return 53; }
long long sq_resnet50_packed_weight_elems(void) { return plan().w_elems; }
long long sq_resnet50_shift_elems(void) { return plan().s_elems; }

int sq_resnet50_conv_info(int idx, int* cin, int* cout, int* k, int* stride, int* pad) {
    if (idx < 0 || idx >= 53) { set_error("conv index out of range"); return -1; }
    const ConvSpec& c = plan().conv[idx];
    *cin = c.cin; *cout = c.cout; *k = c.k; *stride = c.stride; *pad = c.pad;
    return 0;
}

int sq_resnet50_prepack_planes(const void* const* tensors, void* packed_w, void* packed_w_lo, float* shifts, float bn_eps, void* stream);

int sq_resnet50_prepack(const void* const* tensors, void* packed_w, float* shifts, float bn_eps, void* stream) {
    return sq_resnet50_prepack_planes(tensors, packed_w, nullptr, shifts, bn_eps, stream);
}

int sq_resnet50_prepack_planes(const void* const* tensors, void* packed_w, void* packed_w_lo, float* shifts, float bn_eps, void* stream) {
    const ResNetPlan& p = plan();
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < 53; ++i) {
        const ConvSpec& c = p.conv[i];
        const float* const* t = reinterpret_cast<const float* const*>(tensors + 5 * i);
        for (int j = 0; j < 5; ++j) if (!t[j]) { set_error("prepack: null tensor %d of conv %d", j, i); return -1; }
        const long long n = (long long)c.cout * (i == 0 ? STEM_K : c.k * c.k * c.cin);
        int blocks = (int)((n + 255) / 256); if (blocks > 4096) blocks = 4096;
        fold_bn_kernel<<<blocks, 256, 0, st>>>(t[0], t[1], t[2], t[3], t[4], bn_eps, c.cout, c.cin, c.k, i == 0 ? STEM_K : 0,
                                               (bf16*)packed_w + c.w_off, shifts + c.s_off, packed_w_lo ? (bf16*)packed_w_lo + c.w_off : nullptr);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("prepack: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

int sq_conv_bf16(const sq_conv_desc* d, void* stream) {
    if (!d || !d->in || !d->weight || !d->shift || !d->out) { set_error("conv: null pointer"); return -1; }
    if (d->batch <= 0 || d->Cin % 64 || d->Cout % 64 || d->R != d->S || d->stride < 1) { set_error("conv: unsupported geometry"); return -1; }
    const int Ho = (d->H + 2 * d->pad - d->R) / d->stride + 1, Wo = (d->W + 2 * d->pad - d->S) / d->stride + 1;
    ConvGemmArgs a; memset(&a, 0, sizeof(a));
    a.M = d->batch * Ho * Wo; a.N = d->Cout; a.K = d->R * d->S * d->Cin;
    a.A = (const bf16*)d->in; a.lda = d->Cin; a.W = (const bf16*)d->weight; a.bias = d->shift; a.res = (const bf16*)d->residual; a.out = (bf16*)d->out;
    a.relu = d->relu; a.block_n = d->block_n; a.cta_group = d->cta_group;
    a.conv.enabled = (d->R == 1 && d->stride == 1 && d->pad == 0) ? 0 : 1; a.conv.batch = d->batch; a.conv.H = d->H; a.conv.W = d->W; a.conv.C = d->Cin;
    a.conv.Ho = Ho; a.conv.Wo = Wo; a.conv.R = d->R; a.conv.S = d->S; a.conv.stride = d->stride; a.conv.pad = d->pad;
    if (!convgemm_supported(a)) { set_error("conv: operands must be 16-byte aligned, channels multiples of 64"); return -1; }
    return convgemm_launch(a, (cudaStream_t)stream);
}

int sq_bneck_l1_bf16(const void* in, const void* w2, const float* shift2, const void* w3, const float* shift3, const void* residual, void* out,
                     int batch, int H, int W, void* stream) {
    if (!in || !w2 || !shift2 || !w3 || !shift3 || !residual || !out) { set_error("bneck_l1: null pointer"); return -1; }
    if (batch <= 0 || !bneck_l1_supported(H, W)) { set_error("bneck_l1: H must be a multiple of %d and W of %d", HALO_TH, HALO_TW); return -1; }
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(w3) | reinterpret_cast<uintptr_t>(residual) |
         reinterpret_cast<uintptr_t>(out)) & 15) { set_error("bneck_l1: operands must be 16-byte aligned"); return -1; }
    return bneck_l1_launch((const bf16*)in, (const bf16*)w2, shift2, (const bf16*)w3, shift3, (const bf16*)residual, (bf16*)out, batch, H, W, (cudaStream_t)stream);
}

int sq_bneck_l1_ds_bf16(const void* in, const void* w2, const float* shift2, const void* w3, const float* shift3, const void* x, const void* wds,
                        const float* shiftds, void* out, int batch, int H, int W, void* stream) {
    if (!in || !w2 || !shift2 || !w3 || !shift3 || !x || !wds || !shiftds || !out) { set_error("bneck_l1_ds: null pointer"); return -1; }
    if (batch <= 0 || !bneck_l1_supported(H, W)) { set_error("bneck_l1_ds: H must be a multiple of %d and W of %d", HALO_TH, HALO_TW); return -1; }
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(w2) | reinterpret_cast<uintptr_t>(w3) | reinterpret_cast<uintptr_t>(x) |
         reinterpret_cast<uintptr_t>(wds) | reinterpret_cast<uintptr_t>(out)) & 15) { set_error("bneck_l1_ds: operands must be 16-byte aligned"); return -1; }
    return bneck_l1_ds_launch((const bf16*)in, (const bf16*)w2, shift2, (const bf16*)w3, shift3, (const bf16*)x, (const bf16*)wds, shiftds, (bf16*)out, batch, H, W,
                              (cudaStream_t)stream);
}

size_t sq_resnet50_workspace_bytes(int batch, int H, int W) { return ws_layout(batch, H, W).total; }

int sq_resnet50_extract(const void* input, int input_kind, int batch, int H, int W, const void* packed_w, const float* shifts,
                        float* features, void* workspace, size_t workspace_bytes, void* stream) {
    if (batch <= 0) return 0;
    // Any patch size whose final map is 7..13 pixels per side (AvgPool2d(7) then yields one 2048-vector, src/resnet.py:110,166-167):
    // 193..416 px.  256x256 - the tiles the reference pipeline produces (pre_processing/patch_gen_hdf5.py:119-120) - is the tuned
    // case; other sizes (224 px, spatial_vis/visualize.py's 256x265) run the same kernels with clipped edge tiles and an im2col
    // scratch for their stride-2 convolutions.
    {
        int hf = conv_out(conv_out(H, 7, 2, 3), 3, 2, 1), wf = conv_out(conv_out(W, 7, 2, 3), 3, 2, 1);
        for (int i = 0; i < 3; ++i) { hf = conv_out(hf, 3, 2, 1); wf = conv_out(wf, 3, 2, 1); }
        if (hf < 7 || hf > 13 || wf < 7 || wf > 13) { set_error("resnet50_extract: %dx%d patches give a %dx%d final map; AvgPool2d(7) needs 7..13 per side", H, W, hf, wf); return -1; }
        if ((H != 256 || W != 256) && (!convgemm_enabled() || !stem_fused_enabled())) { set_error("resnet50_extract: SQ_CONVGEMM=0 / SQ_STEM_FUSED=0 (round-1 kernels) support 256x256 patches only"); return -1; }
    }
    const ResNetWs L = ws_layout(batch, H, W);
    if (!workspace || workspace_bytes < L.total) { set_error("resnet50_extract: workspace %zu < %zu", workspace_bytes, L.total); return -1; }
    const ResNetPlan& p = plan();
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    const bf16* wp = (const bf16*)packed_w;
    bf16* col = (bf16*)(ws + L.col);
    bf16* stem = (bf16*)(ws + L.stem);
    bf16* big[3] = {(bf16*)(ws + L.big[0]), (bf16*)(ws + L.big[1]), (bf16*)(ws + L.big[2])};
    bf16* small_[2] = {(bf16*)(ws + L.small[0]), (bf16*)(ws + L.small[1])};
    float* fmap = (float*)(ws + L.fmap);

    // ---- stem: one fused kernel (preprocessing + conv1 + BN shift + ReLU + max-pool); SQ_STEM_FUSED=0 selects the older
    //      im2col -> tcgen05 GEMM -> max-pool chain (kept for A/B measurements)
    const int Ho = conv_out(H, 7, 2, 3), Wo = conv_out(W, 7, 2, 3);
    int h = conv_out(Ho, 3, 2, 1), w = conv_out(Wo, 3, 2, 1);
    bf16* im2col = (bf16*)(ws + L.im2col);
    const size_t im2col_bytes = L.total - L.im2col;
    if (stem_fused_enabled() && stem_tc_enabled() && stem_tc_supported(input_kind, H, W)) {
        if (launch_stem_tc(input, input_kind, batch, wp + p.conv[0].w_off, shifts + p.conv[0].s_off, big[0], st)) return -1;
    } else if (stem_fused_enabled()) {
        if (launch_stem_fused(input, input_kind, batch, H, W, wp + p.conv[0].w_off, shifts + p.conv[0].s_off, big[0], st)) return -1;
    } else {
    stem_im2col_kernel<<<batch * Ho * (Wo / 32), 256, 0, st>>>(input, input_kind, H, W, Ho, Wo, col);
    {
        GemmArgs g; memset(&g, 0, sizeof(g));
        g.M = batch * Ho * Wo; g.N = 64; g.K = STEM_K;
        g.A.hi = col; g.A.ld = STEM_K; g.B.hi = wp + p.conv[0].w_off; g.B.ld = STEM_K; g.nterms = 1;
        g.e.bias = shifts + p.conv[0].s_off; g.e.out_hi = stem; g.e.ld_bf = 64; g.e.act = ACT_RELU; g.e.alpha = 1.0f; g.e.rowbias_div = 1;
        if (gemm_launch(g, st)) return -1;
    }
    {
        const long long n = (long long)batch * h * w * 8;
        maxpool3x3s2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(stem, big[0], batch, Ho, Wo, 64);
    }
    }
    // ---- 16 bottlenecks (when bench.py times the kernels: ONE event pair around the chain, so the launches overlap as they do untimed)
    gemm_timing_chain_begin(st);
    const int blocks[4] = {3, 4, 6, 3};
    bool pooled = false;
    int ci = 1; int x = 0;   // big[x] holds the block input
    for (int stg = 0; stg < 4; ++stg)
        for (int b = 0; b < blocks[stg]; ++b) {
            const bool last = (stg == 3 && b == blocks[3] - 1);
            const bool down = (b == 0);
            const ConvSpec& c1 = p.conv[ci]; const ConvSpec& c2 = p.conv[ci + 1]; const ConvSpec& c3 = p.conv[ci + 2];
            int h1, w1, h2, w2, h3, w3, hd, wd;
            if (run_conv(c1, wp, shifts, big[x], batch, h, w, small_[0], nullptr, nullptr, true, st, &h1, &w1)) return -1;
            // layer 1: conv2 + conv3 as one kernel (the 64-channel intermediate stays in tensor memory); SQ_BNECK_FUSE=0 keeps the two launches
            const bool fuse_tail = stg == 0 && bneck_fuse_enabled() && convgemm_enabled() && !last && bneck_l1_supported(h1, w1);
            if (fuse_tail) { h2 = h1; w2 = w1; }
            else if (run_conv(c2, wp, shifts, small_[0], batch, h1, w1, small_[1], nullptr, nullptr, true, st, &h2, &w2, nullptr, im2col, im2col_bytes)) return -1;
            const bf16* res = big[x];
            const int y = (x + 1) % 3, d = (x + 2) % 3;
            // first block of layer 1: the downsample branch is computed inside the fused tail (no residual tensor at all); SQ_BNECK_DS=0 keeps its launch
            const bool fuse_ds = fuse_tail && down && bneck_ds_enabled();
            if (down && !fuse_ds) {
                if (run_conv(p.conv[ci + 3], wp, shifts, big[x], batch, h, w, big[d], nullptr, nullptr, false, st, &hd, &wd, nullptr, im2col, im2col_bytes)) return -1;
                res = big[d];
            }
            // the last convolution of an 8x8 final map feeds the fused average pool (no fp32 map, no pooling kernel)
            const bool fuse_pool = last && h2 == 8 && w2 == 8 && convgemm_enabled() && pool_fused_enabled();
            if (fuse_tail) {
                const ConvSpec& cd = p.conv[ci + 3];
                if (fuse_ds ? bneck_l1_ds_launch(small_[0], wp + c2.w_off, shifts + c2.s_off, wp + c3.w_off, shifts + c3.s_off, big[x], wp + cd.w_off, shifts + cd.s_off,
                                                 big[y], batch, h1, w1, st)
                            : bneck_l1_launch(small_[0], wp + c2.w_off, shifts + c2.s_off, wp + c3.w_off, shifts + c3.s_off, res, big[y], batch, h1, w1, st)) return -1;
                h3 = h2; w3 = w2;
            } else
            if (fuse_pool) {
                cudaMemsetAsync(features, 0, (size_t)batch * 2048 * sizeof(float), st);
                if (run_conv(c3, wp, shifts, small_[1], batch, h2, w2, nullptr, nullptr, res, true, st, &h3, &w3, features)) return -1;
                pooled = true;
            } else
            if (run_conv(c3, wp, shifts, small_[1], batch, h2, w2, last ? nullptr : big[y], last ? fmap : nullptr, res, true, st, &h3, &w3)) return -1;
            x = y; h = h3; w = w3;
            ci += down ? 4 : 3;
        }
    gemm_timing_chain_end(st);
    if (!pooled) {
        const long long n = (long long)batch * 2048;
        avgpool7_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fmap, features, batch, h, w, 2048);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("resnet50_extract: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ split-precision extraction
// Opt-in HIGH-PRECISION mode (quantifies what bf16 operands cost): every tensor is carried as fp32 plus bf16 hi / lo planes, every
// convolution is three tcgen05 MMA groups (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) through the generic kernel of gemm.cuh,
// residuals are added in fp32.  ~1e-5 from the fp64 reference (bf16 operands: 1.4e-3) at about a fifth of the throughput.
namespace sq {

// im2col of the stem as hi / lo planes: K index = r*24 + s*3 + c (21 taps + 3 zeros per filter row), zeros up to 192
__global__ void stem_im2col_planes_kernel(const void* __restrict__ in, int kind, int batch, int H, int W, int Ho, int Wo, bf16* __restrict__ col_hi,
                                          bf16* __restrict__ col_lo) {
    const long long total = (long long)batch * Ho * Wo * (STEM_K / 8);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k8 = (int)(i % (STEM_K / 8)) * 8; long long t = i / (STEM_K / 8);
        const int ow = (int)(t % Wo); t /= Wo; const int oh = (int)(t % Ho); const int img = (int)(t / Ho);
        __align__(16) bf16 hi[8]; __align__(16) bf16 lo[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = k8 + u, r = k / 24, q = k - r * 24;
            float v = 0.f;
            if (r < 7 && q < 21) {
                const int s_ = q / 3, c = q - s_ * 3, ih = oh * 2 - 3 + r, iw = ow * 2 - 3 + s_;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
                    if (kind == 0) {
                        const uint8_t b = reinterpret_cast<const uint8_t*>(in)[(((long long)img * H + ih) * W + iw) * 3 + c];
                        const float mu = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f), sd = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
                        v = (static_cast<float>(b) / 255.0f - mu) / sd;
                    } else {
                        v = reinterpret_cast<const float*>(in)[(((long long)img * 3 + c) * H + ih) * W + iw];
                    }
                }
            }
            hi[u] = __float2bfloat16_rn(v); lo[u] = __float2bfloat16_rn(v - __bfloat162float(hi[u]));
        }
        *reinterpret_cast<uint4*>(col_hi + i * 8) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(col_lo + i * 8) = *reinterpret_cast<const uint4*>(lo);
    }
}

// 3x3 stride-2 pad-1 max pool on the fp32 map -> fp32 + planes
__global__ void maxpool3x3s2_f32_kernel(const float* __restrict__ in, float* __restrict__ out, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo,
                                        int batch, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2;
    const long long n = (long long)batch * Ho * Wo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); long long t = i / C;
        const int ow = (int)(t % Wo); t /= Wo; const int oh = (int)(t % Ho); const int img = (int)(t / Ho);
        float m = -INFINITY;
        for (int r = 0; r < 3; ++r) {
            const int ih = oh * 2 - 1 + r; if (ih < 0 || ih >= H) continue;
            for (int s_ = 0; s_ < 3; ++s_) {
                const int iw = ow * 2 - 1 + s_; if (iw < 0 || iw >= W) continue;
                m = fmaxf(m, in[(((long long)img * H + ih) * W + iw) * C + c]);
            }
        }
        out[i] = m;
        const bf16 h = __float2bfloat16_rn(m);
        out_hi[i] = h; out_lo[i] = __float2bfloat16_rn(m - __bfloat162float(h));
    }
}

struct HpBuf { float* f; bf16* hi; bf16* lo; };
struct HpWs { size_t col_hi, col_lo, stem_f, stem_hi, stem_lo, big_f[3], big_hi[3], big_lo[3], small_hi[2], small_lo[2], fmap, total; };

static HpWs hp_layout(int batch, int H, int W) {
    HpWs w; size_t off = 0;
    const size_t Ho = H / 2, Wo = W / 2, Hp = H / 4, Wp = W / 4;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
    w.col_hi = take((size_t)batch * Ho * Wo * STEM_K * 2); w.col_lo = take((size_t)batch * Ho * Wo * STEM_K * 2);
    w.stem_f = take((size_t)batch * Ho * Wo * 64 * 4); w.stem_hi = take((size_t)batch * Ho * Wo * 64 * 2); w.stem_lo = take((size_t)batch * Ho * Wo * 64 * 2);
    for (int i = 0; i < 3; ++i) { w.big_f[i] = take((size_t)batch * Hp * Wp * 256 * 4); w.big_hi[i] = take((size_t)batch * Hp * Wp * 256 * 2); w.big_lo[i] = take((size_t)batch * Hp * Wp * 256 * 2); }
    for (int i = 0; i < 2; ++i) { w.small_hi[i] = take((size_t)batch * Hp * Wp * 128 * 2); w.small_lo[i] = take((size_t)batch * Hp * Wp * 128 * 2); }
    w.fmap = take((size_t)batch * (H / 32) * (W / 32) * 2048 * 4);
    w.total = off;
    return w;
}

// one split-precision convolution: planes in, fp32 (optional) + planes (optional) out
static int run_conv_hp(const ConvSpec& c, const bf16* w_hi, const bf16* w_lo, const float* sbase, HpBuf in, int batch, int H, int W, HpBuf out,
                       const float* res_f32, bool relu, cudaStream_t st, int* Ho_out, int* Wo_out) {
    const int Ho = (H + 2 * c.pad - c.k) / c.stride + 1, Wo = (W + 2 * c.pad - c.k) / c.stride + 1;
    GemmArgs g; memset(&g, 0, sizeof(g));
    g.M = batch * Ho * Wo; g.N = c.cout; g.K = c.k * c.k * c.cin;
    g.A.hi = in.hi; g.A.lo = in.lo; g.A.ld = c.cin;
    g.B.hi = w_hi + c.w_off; g.B.lo = w_lo + c.w_off; g.B.ld = g.K;
    g.nterms = 3;
    g.conv.enabled = (c.k == 1 && c.stride == 1) ? 0 : 1; g.conv.batch = batch; g.conv.H = H; g.conv.W = W; g.conv.C = c.cin; g.conv.Ho = Ho; g.conv.Wo = Wo;
    g.conv.R = c.k; g.conv.S = c.k; g.conv.stride = c.stride; g.conv.pad = c.pad;
    g.e.bias = sbase + c.s_off;
    g.e.out_f32 = out.f; g.e.ld_f32 = c.cout;
    g.e.out_hi = out.hi; g.e.out_lo = out.lo; g.e.ld_bf = c.cout;
    g.e.res_f32 = res_f32; g.e.ld_res = c.cout;
    g.e.act = relu ? ACT_RELU : ACT_NONE; g.e.alpha = 1.0f; g.e.rowbias_div = 1;
    *Ho_out = Ho; *Wo_out = Wo;
    return gemm_launch(g, st);
}

}  // namespace sq

extern "C" {

size_t sq_resnet50_hp_workspace_bytes(int batch, int H, int W) { return hp_layout(batch, H, W).total; }

int sq_resnet50_extract_hp(const void* input, int input_kind, int batch, int H, int W, const void* packed_w, const void* packed_w_lo,
                           const float* shifts, float* features, void* workspace, size_t workspace_bytes, void* stream) {
    if (batch <= 0) return 0;
    if (H != 256 || W != 256) { set_error("resnet50_extract_hp: the split-precision mode supports 256x256 patches only (got %dx%d)", H, W); return -1; }
    if (!packed_w_lo) { set_error("resnet50_extract_hp: needs the lo plane of the weights (sq_resnet50_prepack_planes)"); return -1; }
    const HpWs L = hp_layout(batch, H, W);
    if (!workspace || workspace_bytes < L.total) { set_error("resnet50_extract_hp: workspace %zu < %zu", workspace_bytes, L.total); return -1; }
    const ResNetPlan& p = plan();
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    const bf16* wh = (const bf16*)packed_w; const bf16* wl = (const bf16*)packed_w_lo;
    const int Ho = H / 2, Wo = W / 2;
    int h = Ho / 2, w = Wo / 2;
    HpBuf col = {nullptr, (bf16*)(ws + L.col_hi), (bf16*)(ws + L.col_lo)};
    HpBuf stem = {(float*)(ws + L.stem_f), (bf16*)(ws + L.stem_hi), (bf16*)(ws + L.stem_lo)};
    HpBuf big[3], small_[2];
    for (int i = 0; i < 3; ++i) big[i] = HpBuf{(float*)(ws + L.big_f[i]), (bf16*)(ws + L.big_hi[i]), (bf16*)(ws + L.big_lo[i])};
    for (int i = 0; i < 2; ++i) small_[i] = HpBuf{nullptr, (bf16*)(ws + L.small_hi[i]), (bf16*)(ws + L.small_lo[i])};
    float* fmap = (float*)(ws + L.fmap);
    {
        const long long total = (long long)batch * Ho * Wo * (STEM_K / 8);
        long long blocks = (total + 255) / 256; if (blocks > 148LL * 64) blocks = 148LL * 64;
        stem_im2col_planes_kernel<<<(unsigned)blocks, 256, 0, st>>>(input, input_kind, batch, H, W, Ho, Wo, col.hi, col.lo);
        GemmArgs g; memset(&g, 0, sizeof(g));
        g.M = batch * Ho * Wo; g.N = 64; g.K = STEM_K;
        g.A.hi = col.hi; g.A.lo = col.lo; g.A.ld = STEM_K; g.B.hi = wh + p.conv[0].w_off; g.B.lo = wl + p.conv[0].w_off; g.B.ld = STEM_K; g.nterms = 3;
        g.e.bias = shifts + p.conv[0].s_off; g.e.out_f32 = stem.f; g.e.ld_f32 = 64; g.e.act = ACT_RELU; g.e.alpha = 1.0f; g.e.rowbias_div = 1;
        if (gemm_launch(g, st)) return -1;
        const long long n = (long long)batch * h * w * 64;
        maxpool3x3s2_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(stem.f, big[0].f, big[0].hi, big[0].lo, batch, Ho, Wo, 64);
    }
    const int blocks[4] = {3, 4, 6, 3};
    int ci = 1; int x = 0;
    for (int stg = 0; stg < 4; ++stg)
        for (int b = 0; b < blocks[stg]; ++b) {
            const bool last = (stg == 3 && b == blocks[3] - 1);
            const bool down = (b == 0);
            int h1, w1, h2, w2, h3, w3, hd, wd;
            if (run_conv_hp(p.conv[ci], wh, wl, shifts, big[x], batch, h, w, small_[0], nullptr, true, st, &h1, &w1)) return -1;
            if (run_conv_hp(p.conv[ci + 1], wh, wl, shifts, small_[0], batch, h1, w1, small_[1], nullptr, true, st, &h2, &w2)) return -1;
            const float* res = big[x].f;
            const int y = (x + 1) % 3, d = (x + 2) % 3;
            if (down) {
                HpBuf dso = {big[d].f, nullptr, nullptr};
                if (run_conv_hp(p.conv[ci + 3], wh, wl, shifts, big[x], batch, h, w, dso, nullptr, false, st, &hd, &wd)) return -1;
                res = big[d].f;
            }
            HpBuf o3 = last ? HpBuf{fmap, nullptr, nullptr} : big[y];
            if (run_conv_hp(p.conv[ci + 2], wh, wl, shifts, small_[1], batch, h2, w2, o3, res, true, st, &h3, &w3)) return -1;
            x = y; h = h3; w = w3;
            ci += down ? 4 : 3;
        }
    {
        const long long n = (long long)batch * 2048;
        avgpool7_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fmap, features, batch, h, w, 2048);
    }
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("resnet50_extract_hp: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // extern "C"
