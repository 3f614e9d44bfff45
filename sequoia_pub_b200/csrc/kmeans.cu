// Per-slide k-means reduction (reference: pre_processing/kmean_features.py:96-105 —
// `KMeans(n_clusters=100, random_state=0).fit(features)` followed by the per-label mean of the raw features).
// The algorithm is scikit-learn's (sklearn/cluster/_kmeans.py: fit, _kmeans_plusplus, _kmeans_single_lloyd;
// _k_means_lloyd.pyx); this file re-expresses it for one B200 with every order-dependent float operation that can
// change a LABEL kept in sklearn's order:
//   * column mean / centring: sequential float32 over rows (numpy add.reduce over axis 0);
//   * k-means++ distances: float64 (sklearn upcasts float32 chunks), rounded to float32, clamped at 0;
//   * the potential sums `closest @ w`, `dist @ w`: the summation orders of the OpenBLAS 0.3.30 x86-64 sdot / sgemv_t
//     kernels numpy dispatches to (restated and pinned against numpy in oracle/kmeans_oracle.py);
//   * cumsum for the candidate draw: sequential float32; searchsorted(left) against float64 rand_vals;
//   * Lloyd: D = ||c||^2 - 2 x.c in float32 FMA, first-minimum argmin, centres = ascending-row sums * (1/count);
//   * cluster features: ascending-row float32 sums of the RAW features divided by the count (np.mean(axis=0)).
// The MT19937 draws (RandomState(0).choice / .uniform) are data independent and are produced on the host by the caller.
// Everything is deterministic (integer atomics only).
#include "gemm.cuh"
#include "../../include/sequoia_b200.h"

namespace sq {

constexpr int KM_NB = 4096;          // OpenBLAS sgemv_t block length
constexpr int KM_MAXT = 10;          // max local trials: 2 + int(ln k) for k < 2981

struct KmFlags { int n_changed; int tol_ok; int n_empty; int reloc_skip; float shift_tot; float tol; int iter; int done; int strict; int max_iter; };

// ---------------------------------------------------------------- preparation
// One thread per column: sequential float32 sum over rows (numpy's order), mean = sum / n; second pass: variance.
__global__ void km_colstats_kernel(const float* __restrict__ X, int n, int d, float* __restrict__ mean, float* __restrict__ var) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= d) return;
    float s = 0.f;
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = X[(size_t)(i + u) * d + j];
#pragma unroll
        for (int u = 0; u < 8; ++u) s = __fadd_rn(s, v[u]);
    }
    for (; i < n; ++i) s = __fadd_rn(s, X[(size_t)i * d + j]);
    const float m = __fdiv_rn(s, (float)n);
    mean[j] = m;
    float q = 0.f;
    for (i = 0; i + 8 <= n; i += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const float e = __fsub_rn(X[(size_t)(i + u) * d + j], m); v[u] = __fmul_rn(e, e); }
#pragma unroll
        for (int u = 0; u < 8; ++u) q = __fadd_rn(q, v[u]);
    }
    for (; i < n; ++i) { const float e = __fsub_rn(X[(size_t)i * d + j], m); q = __fadd_rn(q, __fmul_rn(e, e)); }
    var[j] = __fdiv_rn(q, (float)n);
}

// tol = mean(var) * tol_scale (sklearn _tolerance); single block
__global__ void km_tol_kernel(const float* __restrict__ var, int d, float tol_scale, KmFlags* flags) {
    __shared__ double red[256];
    double s = 0.0;
    for (int j = threadIdx.x; j < d; j += 256) s += (double)var[j];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) flags->tol = (float)(red[0] / d) * tol_scale;
}

// Xc = X - mean (float32), xx64[i] = sum_k (double)Xc[i,k]^2 ; one warp per row
__global__ void km_center_kernel(const float* __restrict__ X, const float* __restrict__ mean, int n, int d, float* __restrict__ Xc,
                                 double* __restrict__ xx64) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    double acc = 0.0;
    for (int c = lane * 4; c < d; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(X + (size_t)row * d + c);
        const float4 m = *reinterpret_cast<const float4*>(mean + c);
        const float4 e = make_float4(__fsub_rn(v.x, m.x), __fsub_rn(v.y, m.y), __fsub_rn(v.z, m.z), __fsub_rn(v.w, m.w));
        *reinterpret_cast<float4*>(Xc + (size_t)row * d + c) = e;
        acc += (double)e.x * e.x + (double)e.y * e.y + (double)e.z * e.z + (double)e.w * e.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) xx64[row] = acc;
}

// ---------------------------------------------------------------- k-means++ seeding
// out[t][i] = min(closest[i], float32(max(0, (-2 <x_cand_t, x_i> + ||x_cand_t||^2) + ||x_i||^2))) in float64.
// One warp per pair of rows; candidate rows are read through L1.
template <int T>
__global__ void __launch_bounds__(256) km_dist_kernel(const float* __restrict__ Xc, const double* __restrict__ xx64, const int* __restrict__ cand,
                                                      const float* __restrict__ closest, int n, int d, float* __restrict__ out) {
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int r0 = warp * 2;
    if (r0 >= n) return;
    const int r1 = min(r0 + 1, n - 1);
    int ci[T];
#pragma unroll
    for (int t = 0; t < T; ++t) ci[t] = cand[t];
    double acc0[T], acc1[T];
#pragma unroll
    for (int t = 0; t < T; ++t) { acc0[t] = 0.0; acc1[t] = 0.0; }
    for (int c = lane * 4; c < d; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(Xc + (size_t)r0 * d + c);
        const float4 b = *reinterpret_cast<const float4*>(Xc + (size_t)r1 * d + c);
        const double a0 = a.x, a1 = a.y, a2 = a.z, a3 = a.w, b0 = b.x, b1 = b.y, b2 = b.z, b3 = b.w;
#pragma unroll
        for (int t = 0; t < T; ++t) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(Xc + (size_t)ci[t] * d + c));
            const double q0 = q.x, q1 = q.y, q2 = q.z, q3 = q.w;
            acc0[t] = fma(a0, q0, fma(a1, q1, fma(a2, q2, fma(a3, q3, acc0[t]))));
            acc1[t] = fma(b0, q0, fma(b1, q1, fma(b2, q2, fma(b3, q3, acc1[t]))));
        }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc0[t] += __shfl_xor_sync(0xffffffffu, acc0[t], o);
            acc1[t] += __shfl_xor_sync(0xffffffffu, acc1[t], o);
        }
    }
    if (lane < 2 * T) {
        const int t = lane >> 1, which = lane & 1;
        const int r = which ? r1 : r0;
        if (which == 0 || r0 + 1 < n) {
            double dot = 0.0;
#pragma unroll
            for (int u = 0; u < T; ++u) if (u == t) dot = which ? acc1[u] : acc0[u];
            double dd = __dmul_rn(-2.0, dot);
            dd = __dadd_rn(dd, xx64[ci[t]]);
            dd = __dadd_rn(dd, xx64[r]);
            float f = (float)dd;                      // round to nearest even, like astype(float32)
            f = fmaxf(f, 0.0f);
            if (closest) f = fminf(closest[r], f);
            out[(size_t)t * n + r] = f;
        }
    }
}

// OpenBLAS sgemv_t summation order (oracle/kmeans_oracle.py: blas_order_gemv_row) of T rows of length n, executed by
// the whole block: the rows are streamed through shared memory in chunks of KM_CHUNK elements (coalesced loads by all
// threads), and 8 lanes per row (lane = element index mod 8, or mod 4 for the 2-column kernel) add their elements in
// ascending order.  Block boundaries of KM_NB elements fold the lanes: kind 0: (l + l+4) then (s0+s1)+(s2+s3).
constexpr int KM_CHUNK = 1024;        // divides KM_NB; multiple of 8
__device__ void km_gemv_order_block(const float* rows, int T, int n, float* stage /*[T][KM_CHUNK]*/,
                                    float* pots) {
    const int tid = threadIdx.x;
    const int t = tid >> 3, l = tid & 7;           // lane-threads: tid < 8*T
    const bool worker = t < T;
    const unsigned wmask = __ballot_sync(0xffffffffu, worker);
    // rows go to the 4-column kernel four at a time (kind 0), a remaining pair to the 2-column kernel (kind 1), a remaining
    // single row to the 1-column kernel, whose order equals kind 0 (oracle/kmeans_oracle.py: gemv_row_kind)
    const int rem = T & 3;
    int kind = 0;
    if (worker && t >= T - rem) { const int rr = t - (T - rem); kind = ((rem & 2) && rr < 2) ? 1 : 0; }
    const int m1 = n - (n & 3);
    float y = 0.f, acc = 0.f;
    for (int b0 = 0; b0 < m1; b0 += KM_NB) {
        const int len = min(KM_NB, m1 - b0);
        acc = 0.f;
        // a leading group of 4 goes to lanes 0-3 when the block length is not a multiple of 8 (for the 4-lane kernel this
        // is simply its first group); everything after it is whole groups of 8, so chunks never split a group
        const int lead = (len & 4) ? 4 : 0;
        if (worker && lead && l < 4) acc = rows[(size_t)t * n + b0 + l];
        for (int c0 = lead; c0 < len; c0 += KM_CHUNK) {
            const int clen = min(KM_CHUNK, len - c0);
            __syncthreads();
            for (int e = tid; e < T * clen; e += blockDim.x) {
                const int tt = e / clen, i = e - tt * clen;
                stage[tt * KM_CHUNK + i] = rows[(size_t)tt * n + b0 + c0 + i];
            }
            __syncthreads();
            if (worker) {
                const float* a = stage + t * KM_CHUNK;
                if (kind == 0) {
#pragma unroll 4
                    for (int i = 0; i < clen; i += 8) acc = __fadd_rn(acc, a[i + l]);
                } else if (l < 4) {
#pragma unroll 4
                    for (int i = 0; i < clen; i += 4) acc = __fadd_rn(acc, a[i + l]);
                }
            }
        }
        if (worker) {
            const unsigned mask = wmask;
            const int base = (tid & 31) & ~7;
            // (the shuffle is executed by every worker lane of the warp: rows of both kinds can share a warp, e.g. 7 trials)
            const float other = __shfl_sync(mask, acc, base + ((l + 4) & 7));
            const float sfold = kind == 0 ? __fadd_rn(acc, other) : acc;
            const float p01 = __fadd_rn(__shfl_sync(mask, sfold, base + 0), __shfl_sync(mask, sfold, base + 1));
            const float p23 = __fadd_rn(__shfl_sync(mask, sfold, base + 2), __shfl_sync(mask, sfold, base + 3));
            y = __fadd_rn(y, __fadd_rn(p01, p23));
        }
    }
    if (worker && l == 0) {
        if (n & 3) {
            const float* a = rows + (size_t)t * n;
            float tt = a[m1];
            for (int i = m1 + 1; i < n; ++i) tt = __fadd_rn(tt, a[i]);
            y = __fadd_rn(y, tt);
        }
        pots[t] = y;
    }
}

// One block (1024 threads) per seeding step: potentials in BLAS order -> best candidate -> closest := its row ->
// sequential float32 cumsum -> next candidates by searchsorted(left) of uniform * pot.
// step 0: `newc` holds the distances to the first centre (T_in = 1, potential through the sdot order).
// dynamic shared memory: n floats (closest / cumsum) + KM_MAXT * KM_CHUNK floats (staging)
__device__ void km_select_body(float* cs /* n floats: closest, then its cumsum */, float* stage /* KM_MAXT * KM_CHUNK floats */,
                               const float* newc, int T_in, int n, int step, int k, int T_next,
                               const double* __restrict__ uniforms, float* closest, int* cand,
                               int* chosen, float* pot_io) {
    __shared__ float pots[KM_MAXT];
    __shared__ float acc16[64];
    __shared__ int s_best;
    __shared__ float s_pot;
    __shared__ int counts[KM_MAXT];
    const int tid = threadIdx.x;
    if (step == 0) {
        // cblas_sdot order (see oracle/kmeans_oracle.py: blas_order_sdot)
        const int n32 = n & ~31, n64 = n32 & ~63;
        if (tid < 64) {
            float a = 0.f;
#pragma unroll 4
            for (int b = 0; b < n64; b += 64) a = __fadd_rn(a, newc[b + tid]);
            acc16[tid] = a;
        }
        __syncthreads();
        if (tid == 0) {
            float acc[4][8];
            for (int u = 0; u < 4; ++u)
                for (int l = 0; l < 8; ++l) acc[u][l] = __fadd_rn(acc16[u * 16 + l], acc16[u * 16 + l + 8]);
            if (n32 - n64 == 32)
                for (int u = 0; u < 4; ++u)
                    for (int l = 0; l < 8; ++l) acc[u][l] = __fadd_rn(acc[u][l], newc[n64 + u * 8 + l]);
            float t8[8];
            for (int l = 0; l < 8; ++l) t8[l] = __fadd_rn(__fadd_rn(__fadd_rn(acc[0][l], acc[1][l]), acc[2][l]), acc[3][l]);
            float s = 0.f;
            if (n32) {
                const float h0 = __fadd_rn(t8[0], t8[4]), h1 = __fadd_rn(t8[1], t8[5]), h2 = __fadd_rn(t8[2], t8[6]), h3 = __fadd_rn(t8[3], t8[7]);
                s = __fadd_rn(__fadd_rn(h0, h1), __fadd_rn(h2, h3));
            }
            double dsum = (double)s;
            for (int i = n32; i < n; ++i) dsum += (double)newc[i];
            s_pot = (float)dsum; s_best = 0;
        }
    } else {
        km_gemv_order_block(newc, T_in, n, stage, pots);
        __syncthreads();
        if (tid == 0) {
            int best = 0;
            for (int t2 = 1; t2 < T_in; ++t2) if (pots[t2] < pots[best]) best = t2;      // np.argmin: first minimum
            s_best = best; s_pot = pots[best];
        }
    }
    __syncthreads();
    const int best = s_best;
    const float pot = s_pot;
    if (tid == 0) { chosen[step] = cand[best]; *pot_io = pot; }
    for (int i = tid; i < n; i += blockDim.x) { const float v = newc[(size_t)best * n + i]; closest[i] = v; cs[i] = v; }
    if (tid < KM_MAXT) counts[tid] = 0;
    __syncthreads();
    if (step + 1 >= k) return;
    if (tid == 0) {                                  // np.cumsum(float32): strictly sequential adds; the loads run 16 elements ahead
        float a = 0.f;
        int i = 0;
        float4 nx[4];
        if (n >= 16) {
#pragma unroll
            for (int u = 0; u < 4; ++u) nx[u] = *reinterpret_cast<const float4*>(cs + 4 * u);
        }
        for (; i + 16 <= n; i += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) { v[4 * u] = nx[u].x; v[4 * u + 1] = nx[u].y; v[4 * u + 2] = nx[u].z; v[4 * u + 3] = nx[u].w; }
            if (i + 32 <= n) {
#pragma unroll
                for (int u = 0; u < 4; ++u) nx[u] = *reinterpret_cast<const float4*>(cs + i + 16 + 4 * u);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) { a = __fadd_rn(a, v[u]); v[u] = a; }
#pragma unroll
            for (int u = 0; u < 4; ++u) *reinterpret_cast<float4*>(cs + i + 4 * u) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
        }
        for (; i < n; ++i) { a = __fadd_rn(a, cs[i]); cs[i] = a; }
    }
    __syncthreads();
    for (int t = 0; t < T_next; ++t) {
        const double rv = __dmul_rn(uniforms[(size_t)step * T_next + t], (double)pot);
        int c = 0;
        for (int i = tid; i < n; i += blockDim.x) c += ((double)cs[i] < rv) ? 1 : 0;     // searchsorted side='left'
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((tid & 31) == 0 && c) atomicAdd(&counts[t], c);
    }
    __syncthreads();
    if (tid < T_next) cand[tid] = min(counts[tid], n - 1);
}

__global__ void __launch_bounds__(1024) km_select_kernel(const float* __restrict__ newc, int T_in, int n, int step, int k, int T_next,
                                                         const double* __restrict__ uniforms, float* __restrict__ closest, int* __restrict__ cand,
                                                         int* __restrict__ chosen, float* __restrict__ pot_io) {
    extern __shared__ float km_sel_smem[];
    km_select_body(km_sel_smem, km_sel_smem + ((n + 31) & ~31), newc, T_in, n, step, k, T_next, uniforms, closest, cand, chosen, pot_io);
}

// ---------------------------------------------------------------- selection step, second version (steps >= 1)
// The T candidate rows are staged in shared memory ONCE (the first version re-staged them chunk by chunk for the BLAS-order sums and
// then copied and accumulated the winner); the BLAS-order potentials (threads 0 .. 8T-1) run BESIDE the T sequential float32 cumsums
// (lanes 0 .. T-1 of one warp in lockstep, one row each, loads 16 elements ahead of the add chain, results streamed to global memory),
// so the serial part of a step is ONE 4096-add chain instead of potentials + copy + cumsum.  Same arithmetic, same orders.
__device__ void km_gemv_order_smem(const float* rows, int T, int n, int ns, float* pots) {
    const int tid = threadIdx.x;
    const int t = tid >> 3, l = tid & 7;
    const bool worker = t < T;
    const unsigned wmask = __ballot_sync(0xffffffffu, worker);      // called by whole warps
    if (!worker) return;
    const int rem = T & 3;
    int kind = 0;
    if (t >= T - rem) { const int rr = t - (T - rem); kind = ((rem & 2) && rr < 2) ? 1 : 0; }
    const int m1 = n - (n & 3);
    const float* a = rows + (size_t)t * ns;
    float y = 0.f;
    for (int b0 = 0; b0 < m1; b0 += KM_NB) {
        const int len = min(KM_NB, m1 - b0);
        float acc = 0.f;
        const int lead = (len & 4) ? 4 : 0;
        if (lead && l < 4) acc = a[b0 + l];
        if (kind == 0) {
#pragma unroll 8
            for (int i = lead; i < len; i += 8) acc = __fadd_rn(acc, a[b0 + i + l]);
        } else if (l < 4) {
#pragma unroll 8
            for (int i = lead; i < len; i += 4) acc = __fadd_rn(acc, a[b0 + i + l]);
        }
        const int base = (tid & 31) & ~7;
        const float other = __shfl_sync(wmask, acc, base + ((l + 4) & 7));
        const float sfold = kind == 0 ? __fadd_rn(acc, other) : acc;
        const float p01 = __fadd_rn(__shfl_sync(wmask, sfold, base + 0), __shfl_sync(wmask, sfold, base + 1));
        const float p23 = __fadd_rn(__shfl_sync(wmask, sfold, base + 2), __shfl_sync(wmask, sfold, base + 3));
        y = __fadd_rn(y, __fadd_rn(p01, p23));
    }
    if (l == 0) {
        if (n & 3) {
            float tt = a[m1];
            for (int i = m1 + 1; i < n; ++i) tt = __fadd_rn(tt, a[i]);
            y = __fadd_rn(y, tt);
        }
        pots[t] = y;
    }
}

__global__ void __launch_bounds__(1024) km_select2_kernel(const float* __restrict__ newc, int T, int n, int step, int k, const double* __restrict__ uniforms,
                                                          float* __restrict__ closest, float* cum /* [T][n4], n4 = n rounded up to 4 */,
                                                          int* __restrict__ cand, int* __restrict__ chosen, float* __restrict__ pot_io) {
    extern __shared__ __align__(16) float rows_s[];        // [T][ns], ns = n rounded up to 4, plus 4: 16-byte aligned rows, 4 banks apart
    __shared__ float pots[KM_MAXT];
    __shared__ int s_best;
    __shared__ float s_pot;
    __shared__ int counts[KM_MAXT];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n4 = (n + 3) & ~3, ns = n4 + 4;
    for (int e = tid; e < T * n; e += blockDim.x) { const int t = e / n, i = e - t * n; rows_s[(size_t)t * ns + i] = newc[e]; }
    if (tid < KM_MAXT) counts[tid] = 0;
    __syncthreads();
    const int pot_warps = (8 * T + 31) / 32;
    if (warp < pot_warps) {
        km_gemv_order_smem(rows_s, T, n, ns, pots);
    } else if (warp == pot_warps && lane < T && step + 1 < k) {
        const float* a = rows_s + (size_t)lane * ns;
        float* o = cum + (size_t)lane * n4;
        float run = 0.f;
        int i = 0;
        float4 nx[4];
        if (n >= 16) {
#pragma unroll
            for (int u = 0; u < 4; ++u) nx[u] = *reinterpret_cast<const float4*>(a + 4 * u);
        }
        for (; i + 16 <= n; i += 16) {
            float cur[16];
#pragma unroll
            for (int u = 0; u < 4; ++u) { cur[4 * u] = nx[u].x; cur[4 * u + 1] = nx[u].y; cur[4 * u + 2] = nx[u].z; cur[4 * u + 3] = nx[u].w; }
            if (i + 32 <= n) {
#pragma unroll
                for (int u = 0; u < 4; ++u) nx[u] = *reinterpret_cast<const float4*>(a + i + 16 + 4 * u);
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) { run = __fadd_rn(run, cur[u]); cur[u] = run; }
#pragma unroll
            for (int u = 0; u < 16; u += 4) *reinterpret_cast<float4*>(o + i + u) = make_float4(cur[u], cur[u + 1], cur[u + 2], cur[u + 3]);
        }
        for (; i < n; ++i) { run = __fadd_rn(run, a[i]); o[i] = run; }
    }
    __syncthreads();
    if (tid == 0) {
        int best = 0;
        for (int t2 = 1; t2 < T; ++t2) if (pots[t2] < pots[best]) best = t2;      // np.argmin: first minimum
        s_best = best; s_pot = pots[best];
        chosen[step] = cand[best]; *pot_io = pots[best];
    }
    __syncthreads();
    const int best = s_best;
    const float pot = s_pot;
    for (int i = tid; i < n; i += blockDim.x) closest[i] = rows_s[(size_t)best * ns + i];
    if (step + 1 >= k) return;
    const float* cb = cum + (size_t)best * n4;
    for (int t = 0; t < T; ++t) {
        const double rv = __dmul_rn(uniforms[(size_t)step * T + t], (double)pot);
        int c = 0;
        for (int i = tid; i < n; i += blockDim.x) c += ((double)cb[i] < rv) ? 1 : 0;     // searchsorted side='left'
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o2);
        if ((tid & 31) == 0 && c) atomicAdd(&counts[t], c);
    }
    __syncthreads();
    if (tid < T) cand[tid] = min(counts[tid], n - 1);
}

// (Measured and dropped in round 2: a distance kernel with the candidate rows staged as doubles in shared memory and four sample rows
//  per warp - 7.89 vs 7.74 ms per slide, no gain, the kernel is L2-latency bound - and a persistent cooperative kernel running all k
//  seeding steps with grid barriers - 9.4 ms, its selection step then runs on one 256-thread CTA.)

// ---------------------------------------------------------------- Lloyd iterations
__global__ void km_gather_centers_kernel(const float* __restrict__ Xc, const int* __restrict__ chosen, int k, int d, float* __restrict__ centers) {
    const int j = blockIdx.x;
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4)
        *reinterpret_cast<float4*>(centers + (size_t)j * d + c) = *reinterpret_cast<const float4*>(Xc + (size_t)chosen[j] * d + c);
}

// The Lloyd kernels take both centre buffers and pick "current" / "next" from the parity of the device-side iteration counter,
// so the same launches serve every iteration of the loop (host loop or CUDA-graph while node).
__device__ __forceinline__ const float* km_cur(const float* c0, const float* c1, const KmFlags* f) { return (f->iter & 1) ? c1 : c0; }
__device__ __forceinline__ float* km_nxt(float* c0, float* c1, const KmFlags* f) { return (f->iter & 1) ? c0 : c1; }

__global__ void km_center_norm_kernel(const float* __restrict__ c0, const float* __restrict__ c1, const KmFlags* __restrict__ flags, int final_pass,
                                      int k, int d, float* __restrict__ csq) {
    if (final_pass && flags->strict) return;
    const float* centers = km_cur(c0, c1, flags);
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (j >= k) return;
    float a = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = centers[(size_t)j * d + c]; a = fmaf(v, v, a); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) csq[j] = a;
}

// label[i] = first argmin_j (csq[j] - 2 <x_i, c_j>) in float32.  Block = 32 rows x 128 centres (looping over centre
// tiles when k > 128), 128 threads, each 4 rows x 8 centres (centres tx*4..+3 and 64+tx*4..+3); K is consumed in
// ascending chunks of 32 through shared memory, one FMA chain per (row, centre).
__global__ void __launch_bounds__(128) km_assign_kernel(const float* __restrict__ Xc, const float* __restrict__ c0, const float* __restrict__ c1,
                                                        const float* __restrict__ csq, int n, int d, int k, const int* __restrict__ labels_old_in,
                                                        int* __restrict__ labels, KmFlags* flags, int final_pass) {
    if (final_pass && flags->strict) return;               // labels of a strictly converged run are final (sklearn L741-753)
    const float* centers = km_cur(c0, c1, flags);
    const int* labels_old = (final_pass || flags->iter == 0) ? nullptr : labels_old_in;
    __shared__ __align__(16) float Xs[32][36];       // [k][row]    (+4 padding keeps float4 reads aligned)
    __shared__ __align__(16) float Cs[32][132];      // [k][centre]
    __shared__ float bestv[32][16];
    __shared__ int besti[32][16];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;          // tx: centre group, ty: row group (4 rows)
    const int row0 = blockIdx.x * 32;
    float rbest[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
    int ribest[4] = {0, 0, 0, 0};
    for (int j0 = 0; j0 < k; j0 += 128) {
        float acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
        for (int k0 = 0; k0 < d; k0 += 32) {
            // global reads are float4 along k (d % 4 == 0), stored transposed
            for (int e = tid; e < 32 * 8; e += 128) {
                const int r = e >> 3, kq = (e & 7) * 4;
                const int row = row0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < n && k0 + kq < d) v = *reinterpret_cast<const float4*>(Xc + (size_t)row * d + k0 + kq);
                Xs[kq][r] = v.x; Xs[kq + 1][r] = v.y; Xs[kq + 2][r] = v.z; Xs[kq + 3][r] = v.w;
            }
            for (int e = tid; e < 128 * 8; e += 128) {
                const int c = e >> 3, kq = (e & 7) * 4;
                const int j = j0 + c;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < k && k0 + kq < d) v = *reinterpret_cast<const float4*>(centers + (size_t)j * d + k0 + kq);
                Cs[kq][c] = v.x; Cs[kq + 1][c] = v.y; Cs[kq + 2][c] = v.z; Cs[kq + 3][c] = v.w;
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) {
                const float4 xv = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
                const float4 c0 = *reinterpret_cast<const float4*>(&Cs[kk][tx * 4]);
                const float4 c1 = *reinterpret_cast<const float4*>(&Cs[kk][64 + tx * 4]);
                const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
                const float cv[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(xr[r], cv[c], acc[r][c]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int j = j0 + (c < 4 ? tx * 4 + c : 64 + tx * 4 + (c - 4));
                if (j < k) {
                    const float dist = fmaf(-2.0f, acc[r][c], csq[j]);
                    if (dist < rbest[r] || (dist == rbest[r] && j < ribest[r])) { rbest[r] = dist; ribest[r] = j; }
                }
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) { bestv[ty * 4 + r][tx] = rbest[r]; besti[ty * 4 + r][tx] = ribest[r]; }
    __syncthreads();
    if (tid < 32) {
        const int row = row0 + tid;
        if (row < n) {
            float bv = bestv[tid][0]; int bi = besti[tid][0];
            for (int t = 1; t < 16; ++t) {
                const float v = bestv[tid][t]; const int i2 = besti[tid][t];
                if (v < bv || (v == bv && i2 < bi)) { bv = v; bi = i2; }
            }
            labels[row] = bi;
            if (labels_old && labels_old[row] != bi) atomicAdd(&flags->n_changed, 1);
        }
    }
}

// Stable bucketing of rows by label: members[offsets[j] .. offsets[j+1]) = rows with label j, ascending. One block.
__global__ void __launch_bounds__(1024) km_bucket_kernel(const int* __restrict__ labels, int n, int k, int* __restrict__ offsets,
                                                         int* __restrict__ members, KmFlags* flags) {
    extern __shared__ int sl[];                    // n labels, then k+1 offsets
    int* soff = sl + n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sl[i] = labels[i];
    for (int j = threadIdx.x; j <= k; j += blockDim.x) soff[j] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&soff[sl[i] + 1], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int empty = 0;
        for (int j = 0; j < k; ++j) { if (soff[j + 1] == 0) ++empty; soff[j + 1] += soff[j]; }
        flags->n_empty = empty;
    }
    __syncthreads();
    for (int j = threadIdx.x; j <= k; j += blockDim.x) offsets[j] = soff[j];
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        int p = soff[j];
        for (int i = 0; i < n; ++i) if (sl[i] == j) members[p++] = i;
    }
}

// ---------------------------------------------------------------- empty-cluster relocation (sklearn _relocate_empty_clusters_dense)
// numpy's float32 pairwise summation (loops_utils.h.src: blocks of <= 128 elements with 8 accumulators, halves split on
// multiples of 8) of (x[i] - c[i])^2 over i < n: the order of `((X - centers_old[labels])**2).sum(axis=1)`.
__device__ float km_pairwise_sqdist(const float* __restrict__ x, const float* __restrict__ c, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) { const float e = __fsub_rn(x[i], c[i]); res = __fadd_rn(res, __fmul_rn(e, e)); }
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const float e = __fsub_rn(x[u], c[u]); r[u] = __fmul_rn(e, e); }
        const int m = n - (n % 8);
        for (int i = 8; i < m; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { const float e = __fsub_rn(x[i + u], c[i + u]); r[u] = __fadd_rn(r[u], __fmul_rn(e, e)); }
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (int i = m; i < n; ++i) { const float e = __fsub_rn(x[i], c[i]); res = __fadd_rn(res, __fmul_rn(e, e)); }
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(km_pairwise_sqdist(x, c, n2), km_pairwise_sqdist(x + n2, c + n2, n - n2));
}

// distances[i] = ||x_i - centers_old[label_i]||^2 in numpy's order; only runs when the bucketing found an empty cluster
__global__ void km_reloc_dist_kernel(const float* __restrict__ Xc, const float* __restrict__ c0, const float* __restrict__ c1,
                                     const int* __restrict__ labels, int n, int d, const KmFlags* __restrict__ flags, float* __restrict__ dist) {
    if (flags->n_empty == 0) return;
    const float* centers = km_cur(c0, c1, flags);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dist[i] = km_pairwise_sqdist(Xc + (size_t)i * d, centers + (size_t)labels[i] * d, d);
}

// One block: empty cluster ids (ascending) and the n_empty samples farthest from their centres in descending distance (ties:
// lower row first) - the order numpy 2.x's argpartition(distances, -n_empty)[:-n_empty-1:-1] returns on x86-64 (see the oracle).
__global__ void __launch_bounds__(1024) km_reloc_select_kernel(float* __restrict__ dist, const int* __restrict__ offsets, int n, int k,
                                                               KmFlags* flags, int* __restrict__ empty_ids, int* __restrict__ far) {
    const int n_empty = flags->n_empty;
    if (n_empty == 0) return;
    __shared__ float bv[32];
    __shared__ int bi[32];
    __shared__ int s_pick;
    const int tid = threadIdx.x;
    if (tid == 0) { int e = 0; for (int j = 0; j < k; ++j) if (offsets[j + 1] == offsets[j]) empty_ids[e++] = j; }
    for (int e = 0; e < n_empty; ++e) {
        float v = -1.0f; int idx = 0x7fffffff;
        for (int i = tid; i < n; i += blockDim.x) { const float x = dist[i]; if (x > v || (x == v && i < idx)) { v = x; idx = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v2 = __shfl_xor_sync(0xffffffffu, v, o); const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
            if (v2 > v || (v2 == v && i2 < idx)) { v = v2; idx = i2; }
        }
        if ((tid & 31) == 0) { bv[tid >> 5] = v; bi[tid >> 5] = idx; }
        __syncthreads();
        if (tid == 0) {
            float m = bv[0]; int mi = bi[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) if (bv[w] > m || (bv[w] == m && bi[w] < mi)) { m = bv[w]; mi = bi[w]; }
            if (e == 0) flags->reloc_skip = (m == 0.0f) ? 1 : 0;       // np.max(distances) == 0: relocation is pointless, centres stay 0
            far[e] = mi; s_pick = mi;
        }
        __syncthreads();
        if (tid == 0) dist[s_pick] = -1.0f;                            // taken
        __syncthreads();
    }
}

// out[j, c] = (sum over members of cluster j, ascending rows, of src[row, c]) * (1/count)   [mode 0: sklearn _average_centers]
//                                                                          / count          [mode 1: np.mean]
// and, when `old` is given, per-(cluster, block) partial sums of (new - old)^2 for the centre shift.
// mode 0 also applies the relocation of empty clusters (reloc != null): an empty cluster takes its far sample, the cluster that
// sample was assigned to loses it (sum - x, count - 1, in the order of the empty cluster ids); a cluster left without weight keeps
// its sum (sklearn only scales clusters with weight > 0).
__global__ void __launch_bounds__(128) km_segment_mean_kernel(const float* __restrict__ src, const int* __restrict__ offsets,
                                                              const int* __restrict__ members, int d, int mode, float* __restrict__ out_in,
                                                              float* __restrict__ c0, float* __restrict__ c1, float* __restrict__ shift_part,
                                                              const KmFlags* __restrict__ flags, const int* __restrict__ labels,
                                                              const int* __restrict__ empty_ids, const int* __restrict__ far) {
    __shared__ float red[4];
    const int j = blockIdx.y;
    const int c = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int b = offsets[j], e = offsets[j + 1];
    const int n_reloc = (mode == 0 && flags->n_empty > 0 && !flags->reloc_skip) ? flags->n_empty : 0;
    float* out = mode == 0 ? km_nxt(c0, c1, flags) : out_in;
    const float* old = mode == 0 ? km_cur(c0, c1, flags) : nullptr;
    float sq = 0.f;
    if (c < d) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = b; p < e; ++p) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)members[p] * d + c);
            a.x = __fadd_rn(a.x, v.x); a.y = __fadd_rn(a.y, v.y); a.z = __fadd_rn(a.z, v.z); a.w = __fadd_rn(a.w, v.w);
        }
        float cnt = (float)(e - b);
        for (int r = 0; r < n_reloc; ++r) {
            const int fi = far[r];
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)fi * d + c);
            if (empty_ids[r] == j) { a = v; cnt = 1.0f; }
            else if (labels[fi] == j) { a.x = __fsub_rn(a.x, v.x); a.y = __fsub_rn(a.y, v.y); a.z = __fsub_rn(a.z, v.z); a.w = __fsub_rn(a.w, v.w); cnt -= 1.0f; }
        }
        if (mode == 0) { if (cnt > 0.0f) { const float inv = __fdiv_rn(1.0f, cnt); a.x = __fmul_rn(a.x, inv); a.y = __fmul_rn(a.y, inv); a.z = __fmul_rn(a.z, inv); a.w = __fmul_rn(a.w, inv); } }
        else { a.x = __fdiv_rn(a.x, cnt); a.y = __fdiv_rn(a.y, cnt); a.z = __fdiv_rn(a.z, cnt); a.w = __fdiv_rn(a.w, cnt); }
        *reinterpret_cast<float4*>(out + (size_t)j * d + c) = a;
        if (old) {
            const float4 o = *reinterpret_cast<const float4*>(old + (size_t)j * d + c);
            const float e0 = a.x - o.x, e1 = a.y - o.y, e2 = a.z - o.z, e3 = a.w - o.w;
            sq = e0 * e0 + e1 * e1 + e2 * e2 + e3 * e3;
        }
    }
    if (old) {
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o2);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) shift_part[(size_t)j * gridDim.x + blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
    }
}

// center_shift_tot = sum_j ||new_j - old_j||^2 ; tol test (sklearn L717-727)
// ... and the loop control of _kmeans_single_lloyd (L700-735): strict convergence (labels unchanged) is tested first, then the
// tolerance; advances the iteration counter (which also swaps the centre buffers) and, inside a CUDA-graph while node, clears the
// node's condition when the loop is over.
__global__ void km_converge_kernel(const float* __restrict__ shift_part, int k, int nparts, KmFlags* flags, int* __restrict__ n_iter_out,
                                   unsigned long long cond_handle) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float tot = 0.f;
    for (int j = 0; j < k; ++j) {
        float s = 0.f;
        for (int p = 0; p < nparts; ++p) s += shift_part[(size_t)j * nparts + p];
        const float sh = sqrtf(s);
        tot += sh * sh;
    }
    flags->shift_tot = tot;
    const int tol_ok = tot <= flags->tol ? 1 : 0;
    flags->tol_ok = tol_ok;
    const int it = flags->iter;
    const int strict = (it > 0 && flags->n_changed == 0) ? 1 : 0;
    const int done = (strict || tol_ok || it + 1 >= flags->max_iter) ? 1 : 0;
    flags->strict = strict; flags->done = done;
    flags->iter = it + 1;
    flags->n_changed = 0;
    if (n_iter_out) *n_iter_out = it + 1;
    if (done && cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, 0);
}

struct KmWs { size_t cum, xc, xx64, mean, var, closest, newc, cand, chosen, pot, cA, cB, csq, labels, labels_old, offsets, members, shift, flags, rdist, empty_ids, far, n_iter, total; };

static void km_ws_layout(int n, int d, int k, KmWs* w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    w->cum = take((size_t)KM_MAXT * ((n + 3) & ~3) * 4);
    w->xc = take((size_t)n * d * 4); w->xx64 = take((size_t)n * 8); w->mean = take((size_t)d * 4); w->var = take((size_t)d * 4);
    w->closest = take((size_t)n * 4); w->newc = take((size_t)KM_MAXT * n * 4); w->cand = take(KM_MAXT * 4); w->chosen = take((size_t)k * 4);
    w->pot = take(4); w->cA = take((size_t)k * d * 4); w->cB = take((size_t)k * d * 4); w->csq = take((size_t)k * 4);
    w->labels = take((size_t)n * 4); w->labels_old = take((size_t)n * 4); w->offsets = take((size_t)(k + 1) * 4); w->members = take((size_t)n * 4);
    w->shift = take((size_t)k * ((d + 511) / 512) * 4); w->flags = take(sizeof(KmFlags));
    w->rdist = take((size_t)n * 4); w->empty_ids = take((size_t)k * 4); w->far = take((size_t)k * 4); w->n_iter = take(4);
    w->total = off;
}

template <int T>
static void launch_dist(const float* Xc, const double* xx64, const int* cand, const float* closest, int n, int d, float* out, cudaStream_t st) {
    const int warps = (n + 1) / 2;
    km_dist_kernel<T><<<(warps + 7) / 8, 256, 0, st>>>(Xc, xx64, cand, closest, n, d, out);
}

}  // namespace sq

using namespace sq;

extern "C" {

size_t sq_kmeans_workspace_bytes(int n, int d, int k) {
    if (n <= 0 || d <= 0 || k <= 0) return 0;
    KmWs w; km_ws_layout(n, d, k, &w);
    return w.total;
}

}  // extern "C"

namespace sq {

struct KmPtrs {
    float *Xc, *cA, *cB, *csq, *shift, *rdist; int *labels, *labels_old, *offsets, *members, *empty_ids, *far, *n_iter; KmFlags* flags;
    int n, d, k;
};

// one Lloyd iteration (lloyd_iter_chunked_dense + relocation + averaging + centre shift + loop control)
static void km_enqueue_iteration(const KmPtrs& P, unsigned long long cond, cudaStream_t st) {
    const int n = P.n, d = P.d, k = P.k, nparts = (d + 511) / 512;
    const size_t bucket_smem = (size_t)n * 4 + (size_t)(k + 1) * 4;
    km_center_norm_kernel<<<(k + 7) / 8, 256, 0, st>>>(P.cA, P.cB, P.flags, 0, k, d, P.csq);
    km_assign_kernel<<<(n + 31) / 32, 128, 0, st>>>(P.Xc, P.cA, P.cB, P.csq, n, d, k, P.labels_old, P.labels, P.flags, 0);
    km_bucket_kernel<<<1, 1024, bucket_smem, st>>>(P.labels, n, k, P.offsets, P.members, P.flags);
    km_reloc_dist_kernel<<<(n + 127) / 128, 128, 0, st>>>(P.Xc, P.cA, P.cB, P.labels, n, d, P.flags, P.rdist);
    km_reloc_select_kernel<<<1, 1024, 0, st>>>(P.rdist, P.offsets, n, k, P.flags, P.empty_ids, P.far);
    km_segment_mean_kernel<<<dim3(nparts, k), 128, 0, st>>>(P.Xc, P.offsets, P.members, d, 0, nullptr, P.cA, P.cB, P.shift, P.flags, P.labels, P.empty_ids, P.far);
    km_converge_kernel<<<1, 32, 0, st>>>(P.shift, k, nparts, P.flags, P.n_iter, cond);
    cudaMemcpyAsync(P.labels_old, P.labels, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
}

// The Lloyd loop as a CUDA graph with a WHILE conditional node (CUDA >= 12.4): the body is one iteration, the condition is
// cleared on the device by km_converge_kernel - no host round trip per iteration.  Graphs are cached per (device, workspace, shape).
struct KmGraph { int dev; void* ws; int n, d, k; cudaGraphExec_t exec; cudaGraph_t graph; unsigned long long stamp; };
static KmGraph g_km_graphs[8];
static unsigned long long g_km_stamp = 0;
static cudaStream_t g_km_capture[16];

static cudaGraphExec_t km_lloyd_graph(const KmPtrs& P, void* ws) {
    int dev = 0; cudaGetDevice(&dev);
    for (auto& e : g_km_graphs)
        if (e.exec && e.dev == dev && e.ws == ws && e.n == P.n && e.d == P.d && e.k == P.k) { e.stamp = ++g_km_stamp; return e.exec; }
    if (dev < 0 || dev >= 16) return nullptr;
    if (!g_km_capture[dev] && cudaStreamCreateWithFlags(&g_km_capture[dev], cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    if (cudaGraphCreate(&graph, 0) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    cudaGraphConditionalHandle h;
    cudaGraphNodeParams np = {};
    bool ok = cudaGraphConditionalHandleCreate(&h, graph, 1, cudaGraphCondAssignDefault) == cudaSuccess;
    cudaGraphNode_t node;
    if (ok) {
        np.type = cudaGraphNodeTypeConditional; np.conditional.handle = h; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
        ok = cudaGraphAddNode(&node, graph, nullptr, 0, &np) == cudaSuccess;
    }
    if (ok) {
        cudaGraph_t body = np.conditional.phGraph_out[0];
        ok = cudaStreamBeginCaptureToGraph(g_km_capture[dev], body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) == cudaSuccess;
        if (ok) {
            km_enqueue_iteration(P, (unsigned long long)h, g_km_capture[dev]);
            ok = cudaStreamEndCapture(g_km_capture[dev], nullptr) == cudaSuccess;
        }
    }
    if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
    if (!ok) { (void)cudaGetLastError(); if (graph) cudaGraphDestroy(graph); return nullptr; }
    KmGraph* slot = &g_km_graphs[0];
    for (auto& e : g_km_graphs) { if (!e.exec) { slot = &e; break; } if (e.stamp < slot->stamp) slot = &e; }
    if (slot->exec) { cudaGraphExecDestroy(slot->exec); cudaGraphDestroy(slot->graph); }
    *slot = KmGraph{dev, ws, P.n, P.d, P.k, exec, graph, ++g_km_stamp};
    return exec;
}

}  // namespace sq

extern "C" {

int sq_kmeans_fit(const float* features, int n, int d, int k, int trials, int first_center, const double* uniforms, const int* init_rows,
                  int max_iter, float tol_scale, int* labels, float* cluster_means, int* chosen_out, int* n_iter_dev, void* workspace,
                  size_t workspace_bytes, void* stream) {
    if (!features || (!uniforms && !init_rows) || !labels || !cluster_means) { set_error("kmeans: null pointer"); return -1; }
    if (n < k || k < 1) { set_error("kmeans: need n >= k >= 1 (n=%d, k=%d)", n, k); return -1; }
    if (d % 4 != 0) { set_error("kmeans: feature dim %d must be a multiple of 4", d); return -1; }
    if (!init_rows && (trials < 2 || trials > KM_MAXT)) { set_error("kmeans: %d local trials unsupported (2 .. %d, i.e. k < 2981)", trials, KM_MAXT); return -1; }
    if (!init_rows && (first_center < 0 || first_center >= n)) { set_error("kmeans: first centre out of range"); return -1; }
    if (max_iter < 1) { set_error("kmeans: max_iter must be >= 1"); return -1; }
    // the selection / bucketing kernels keep one value per sample in the shared memory of a single block
    if ((size_t)n * 4 + (size_t)(k + 1) * 4 > 160 * 1024) { set_error("kmeans: n=%d exceeds the %d samples the single-block selection kernels hold", n, (160 * 1024 - (k + 1) * 4) / 4); return -1; }
    KmWs L; km_ws_layout(n, d, k, &L);
    if (!workspace || workspace_bytes < L.total) { set_error("kmeans: workspace %zu < %zu", workspace_bytes, L.total); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)workspace;
    float* Xc = (float*)(ws + L.xc); double* xx64 = (double*)(ws + L.xx64);
    float* mean = (float*)(ws + L.mean); float* var = (float*)(ws + L.var);
    float* closest = (float*)(ws + L.closest); float* newc = (float*)(ws + L.newc);
    int* cand = (int*)(ws + L.cand); int* chosen = (int*)(ws + L.chosen); float* pot = (float*)(ws + L.pot);
    KmPtrs P;
    P.Xc = Xc; P.cA = (float*)(ws + L.cA); P.cB = (float*)(ws + L.cB); P.csq = (float*)(ws + L.csq); P.shift = (float*)(ws + L.shift);
    P.rdist = (float*)(ws + L.rdist); P.labels = (int*)(ws + L.labels); P.labels_old = (int*)(ws + L.labels_old); P.offsets = (int*)(ws + L.offsets);
    P.members = (int*)(ws + L.members); P.empty_ids = (int*)(ws + L.empty_ids); P.far = (int*)(ws + L.far); P.n_iter = (int*)(ws + L.n_iter);
    P.flags = (KmFlags*)(ws + L.flags); P.n = n; P.d = d; P.k = k;
    KmFlags* flags = P.flags;

    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(km_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(km_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done = true;
    }
    const size_t sel_smem = (size_t)((n + 31) & ~31) * 4 + (size_t)KM_MAXT * KM_CHUNK * 4;
    KmFlags h0; memset(&h0, 0, sizeof(h0)); h0.max_iter = max_iter;
    cudaMemcpyAsync(flags, &h0, sizeof(KmFlags), cudaMemcpyHostToDevice, st);      // pageable source: copied before the call returns
    // ---- preparation (fit L1490-1500)
    km_colstats_kernel<<<(d + 63) / 64, 64, 0, st>>>(features, n, d, mean, var);
    km_tol_kernel<<<1, 256, 0, st>>>(var, d, tol_scale, flags);
    km_center_kernel<<<(n + 7) / 8, 256, 0, st>>>(features, mean, n, d, Xc, xx64);
    if (init_rows) {
        // explicit initial centres (sklearn's `init=X[rows]`): no seeding; duplicates are allowed and give empty clusters
        cudaMemcpyAsync(chosen, init_rows, (size_t)k * 4, cudaMemcpyDeviceToDevice, st);
    } else {
        // ---- k-means++ (L180-279)
        cudaMemcpyAsync(cand, &first_center, sizeof(int), cudaMemcpyHostToDevice, st);
        static const int sel2_on = getenv("SQ_KMEANS_SELECT_V2") ? atoi(getenv("SQ_KMEANS_SELECT_V2")) : 1;
        const size_t sel2_smem = (size_t)trials * (((n + 3) & ~3) + 4) * sizeof(float);
        const bool sel2 = sel2_on && sel2_smem <= 200 * 1024;
        float* cum = (float*)(ws + L.cum);
        static bool sel2_attr = false;
        if (sel2 && !sel2_attr) { cudaFuncSetAttribute(km_select2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); sel2_attr = true; }
        launch_dist<1>(Xc, xx64, cand, nullptr, n, d, newc, st);
        km_select_kernel<<<1, 1024, sel_smem, st>>>(newc, 1, n, 0, k, trials, uniforms, closest, cand, chosen, pot);
        for (int c = 1; c < k; ++c) {
            switch (trials) {
                case 2: launch_dist<2>(Xc, xx64, cand, closest, n, d, newc, st); break;
                case 3: launch_dist<3>(Xc, xx64, cand, closest, n, d, newc, st); break;
                case 4: launch_dist<4>(Xc, xx64, cand, closest, n, d, newc, st); break;
                case 5: launch_dist<5>(Xc, xx64, cand, closest, n, d, newc, st); break;
                case 6: launch_dist<6>(Xc, xx64, cand, closest, n, d, newc, st); break;
                case 7: launch_dist<7>(Xc, xx64, cand, closest, n, d, newc, st); break;
                case 8: launch_dist<8>(Xc, xx64, cand, closest, n, d, newc, st); break;
                case 9: launch_dist<9>(Xc, xx64, cand, closest, n, d, newc, st); break;
                default: launch_dist<10>(Xc, xx64, cand, closest, n, d, newc, st); break;
            }
            if (sel2) km_select2_kernel<<<1, 1024, sel2_smem, st>>>(newc, trials, n, c, k, uniforms, closest, cum, cand, chosen, pot);
            else km_select_kernel<<<1, 1024, sel_smem, st>>>(newc, trials, n, c, k, trials, uniforms, closest, cand, chosen, pot);
        }
    }
    km_gather_centers_kernel<<<k, 256, 0, st>>>(Xc, chosen, k, d, P.cA);
    if (chosen_out) cudaMemcpyAsync(chosen_out, chosen, (size_t)k * 4, cudaMemcpyDeviceToDevice, st);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("kmeans seeding: %s", cudaGetErrorString(err)); return -1; }
    // ---- Lloyd (L630-758)
    static const int use_graph = getenv("SQ_KMEANS_GRAPH") ? atoi(getenv("SQ_KMEANS_GRAPH")) : 1;
    cudaGraphExec_t exec = use_graph ? km_lloyd_graph(P, workspace) : nullptr;
    if (exec) {
        err = cudaGraphLaunch(exec, st);
        if (err != cudaSuccess) { set_error("kmeans lloyd graph: %s", cudaGetErrorString(err)); return -1; }
    } else {
        // no conditional graph nodes (old driver) or SQ_KMEANS_GRAPH=0: host loop, one 40-byte read + stream sync per iteration
        KmFlags h;
        for (int it = 0; it < max_iter; ++it) {
            km_enqueue_iteration(P, 0ull, st);
            cudaMemcpyAsync(&h, flags, sizeof(KmFlags), cudaMemcpyDeviceToHost, st);
            err = cudaStreamSynchronize(st);
            if (err != cudaSuccess) { set_error("kmeans lloyd: %s", cudaGetErrorString(err)); return -1; }
            if (h.done) break;
        }
    }
    // rerun the E-step with the final centres unless the run converged strictly (L741-753)
    const int nparts = (d + 511) / 512;
    const size_t bucket_smem = (size_t)n * 4 + (size_t)(k + 1) * 4;
    km_center_norm_kernel<<<(k + 7) / 8, 256, 0, st>>>(P.cA, P.cB, flags, 1, k, d, P.csq);
    km_assign_kernel<<<(n + 31) / 32, 128, 0, st>>>(Xc, P.cA, P.cB, P.csq, n, d, k, nullptr, P.labels, flags, 1);
    // ---- cluster features: per-label mean of the RAW features (kmean_features.py:99-105)
    km_bucket_kernel<<<1, 1024, bucket_smem, st>>>(P.labels, n, k, P.offsets, P.members, flags);
    km_segment_mean_kernel<<<dim3(nparts, k), 128, 0, st>>>(features, P.offsets, P.members, d, 1, cluster_means, nullptr, nullptr, nullptr, flags, nullptr, nullptr, nullptr);
    cudaMemcpyAsync(labels, P.labels, (size_t)n * 4, cudaMemcpyDeviceToDevice, st);
    if (n_iter_dev) cudaMemcpyAsync(n_iter_dev, P.n_iter, 4, cudaMemcpyDeviceToDevice, st);
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("kmeans: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // extern "C"
