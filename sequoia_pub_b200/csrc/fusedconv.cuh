// Layer-1 bottleneck tail in ONE kernel: conv2 (3x3, 64 -> 64, stride 1, + shift, ReLU) feeding conv3 (1x1, 64 -> 256, + shift,
// + residual, ReLU) of src/resnet.py:73-93 without the 64-channel intermediate ever leaving the SM.
//
// Why: an SS-mode tcgen05.mma fetches its shared-memory operands at ~64 B/clk, so the two N = 64 launches are bound by operand
// fetch and the 1x1 expansion by HBM; unfused they cost 40 + 49 us per block (floor of the pair: 46 us of HBM traffic).  Fused, the
// expansion's A operand (128 pixels x 64 channels, bf16) is written by the first epilogue straight into TENSOR MEMORY - an epilogue
// thread owns one pixel = one TMEM lane, so the tile is 32 packed columns stored with tcgen05.st - and the second MMA reads it from
// there (A in TMEM) against the 256 x 64 weight tile that stays resident in shared memory: 33.5 MB less written and read per block,
// one launch less, and no shared-memory fetch for A in the expansion.
//
//   warp 0      TMA producer: the 256 x 64 conv3 tile ONCE per CTA (resident); per 16 x 8 output patch one 18 x 10 halo box (64
//               channels, two slots) and the nine 64 x 64 conv2 tap tiles through a 4-stage ring
//   warp 1      MMA issuer: A(i) = 36 tcgen05.mma (nine taps from shifted views of the halo, N = 64) into a double-buffered
//               accumulator; B(i-1) = 4 tcgen05.mma (A from TMEM, N = 256) issued after A(i), so the tensor pipe never waits for
//               the first epilogue
//   warp 2      TMEM allocation (512 columns: 2 x 64 conv2 accumulators, 2 x 32 packed intermediates, 256 conv3 accumulator)
//   warps 4-7   epilogue A: tcgen05.ld, + shift, ReLU, bf16 pack (cvt.rn.relu), tcgen05.st of the intermediate - and, because that is
//               300 cycles of a 5000-cycle tile, channel chunks 2 and 3 of epilogue B of the previous tile
//   warps 8-11  epilogue B, channel chunks 0 and 1: the TMA epilogue of convgemm.cuh (residual sub-tiles prefetched by TMA into a
//               per-warp ring, shift + residual + ReLU in place, cp.async.bulk.tensor stores)
// Measured steps (per-role cycle counters, SQ_BNECK_PROF=1): with epilogue B on four warps the kernel took 67 us and the MMA warp
// waited 30 % of the time for the conv3 accumulator to drain (1800 cycles per 32 x 64 item); staging the halo as three shifted
// copies (every tap a plain K-major tile of whole swizzle atoms) changed nothing, so the shifted views stay.
// DS = 1 (first block of the layer): the downsample branch (1x1, 64 -> 256 on the BLOCK INPUT) is computed inside the kernel as four
// more MMAs into the same conv3 accumulator - A = the 16 x 8 input patch (one more TMA box per tile, issued by the otherwise idle
// warp 3), B = the resident 256 x 64 downsample tile - so the residual is never written (134 MB) nor read back (134 MB) and the
// downsample launch disappears: out = relu(T W3^T + X Wds^T + shift3 + shift_ds).  (The sum is rounded once instead of twice.)
// DS = 0: results are bit-identical to the two separate launches: the intermediate is rounded to bf16 exactly as when it was stored.
#pragma once
#include "convgemm.cuh"

namespace sq {

constexpr int FB_WS = 4;                              // conv2 tap-tile ring: [64 n][64 k] per stage
constexpr int FB_W2_BYTES = FB_WS * 8192;
constexpr int FB_W3_BYTES = 256 * 128;                // [256 n][64 k]
constexpr int FB_HS = 2;                              // halo slots (HALO_BYTES each: 18 x 10 pixels x 64 channels)
constexpr int FB_SUB_BYTES = 32 * 128;
constexpr int FB_TAB_BYTES = (64 + 256) * 4;
template <int DS> struct FbCfg {
    static constexpr int D = DS ? 2 : 3;                          // ring depth of an epilogue-B warp (no residual prefetch with DS)
    static constexpr int RING_BYTES = 8 * D * FB_SUB_BYTES;
    static constexpr int WDS_BYTES = DS ? FB_W3_BYTES : 0;        // resident downsample weights [256 n][64 k]
    static constexpr int X_BYTES = DS ? 128 * 128 : 0;            // the block-input patch of a tile: [128 pixels][64 channels]
    static constexpr int SMEM = FB_W2_BYTES + FB_W3_BYTES + WDS_BYTES + FB_HS * HALO_BYTES + X_BYTES + RING_BYTES + FB_TAB_BYTES + 512;
    static_assert(SMEM <= 232448, "shared memory budget");
};

struct FbParams {
    int batch, tiles_x, tiles_per_img, total_tiles;
    const float* shift2;   // [64]  folded BN shift of conv2
    const float* shift3;   // [256] folded BN shift of conv3
    const float* shiftds;  // [256] folded BN shift of the downsample (DS = 1)
    int prof;              // SQ_BNECK_PROF=1: block 0 prints the cycles each role spends waiting
};

template <int DS>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
bneck_l1_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW2, const __grid_constant__ CUtensorMap mapW3,
                const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapO, const __grid_constant__ CUtensorMap mapX,
                const __grid_constant__ CUtensorMap mapWds, const FbParams p) {
    using Cfg = FbCfg<DS>;
    constexpr int FB_D = Cfg::D;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sW2 = smem;
    uint8_t* sW3 = sW2 + FB_W2_BYTES;
    uint8_t* sWds = sW3 + FB_W3_BYTES;
    uint8_t* halo = sWds + Cfg::WDS_BYTES;
    uint8_t* sX = halo + FB_HS * HALO_BYTES;
    uint8_t* ring = sX + Cfg::X_BYTES;
    float* tab = reinterpret_cast<float*>(ring + Cfg::RING_BYTES);          // shift2[64], shift3[256] (+ shift_ds)
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Cfg::RING_BYTES + FB_TAB_BYTES);
    uint64_t* w_full = bars;                 // conv3 weights resident
    uint64_t* full_w = bars + 1;             // [WS] conv2 tap tile landed
    uint64_t* empty_w = full_w + FB_WS;      // [WS]
    uint64_t* full_h = empty_w + FB_WS;      // [HS]
    uint64_t* empty_h = full_h + FB_HS;      // [HS]
    uint64_t* acca_full = empty_h + FB_HS;   // [2] conv2 accumulator ready
    uint64_t* acca_free = acca_full + 2;     // [2] read by the four epilogue-A warps
    uint64_t* t_full = acca_free + 2;        // [2] intermediate written (four warps)
    uint64_t* t_free = t_full + 2;           // [2] intermediate consumed (tcgen05.commit)
    uint64_t* accb_full = t_free + 2;        // conv3 accumulator ready
    uint64_t* accb_free = accb_full + 1;     // read by the eight warps that run epilogue B
    uint64_t* x_full = accb_free + 1;        // DS: input patch landed
    uint64_t* x_empty = x_full + 1;          // DS: input patch consumed (tcgen05.commit)
    uint64_t* res_full = x_empty + 1;        // [8][D]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_full + 8 * FB_D);
    static_assert((1 + 2 * FB_WS + 2 * FB_HS + 8 + 2 + 2 + 8 * FB_D) * 8 + 4 <= 512, "barrier area");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    asm volatile("griddepcontrol.launch_dependents;");
    if (warp == 0 && lane == 0) {
        if (smem_u32(smem) & 1023u) { printf("sequoia_b200: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapW2); tma_prefetch_desc(&mapW3); tma_prefetch_desc(&mapR); tma_prefetch_desc(&mapO);
        if constexpr (DS) { tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapWds); }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(w_full, 1);
        for (int i = 0; i < FB_WS; ++i) { mbar_init(&full_w[i], 1); mbar_init(&empty_w[i], 1); }
        for (int i = 0; i < FB_HS; ++i) { mbar_init(&full_h[i], 1); mbar_init(&empty_h[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acca_full[i], 1); mbar_init(&acca_free[i], 4); mbar_init(&t_full[i], 4); mbar_init(&t_free[i], 1); }
        mbar_init(accb_full, 1); mbar_init(accb_free, 8); mbar_init(x_full, 1); mbar_init(x_empty, 1);
        for (int i = 0; i < 8 * FB_D; ++i) mbar_init(&res_full[i], 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    for (int i = threadIdx.x; i < 64 + 256; i += GEMM_THREADS) tab[i] = i < 64 ? p.shift2[i] : p.shift3[i - 64] + (DS ? p.shiftds[i - 64] : 0.0f);     // weights-side data: not produced by the previous kernel
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    long long pc[6] = {0, 0, 0, 0, 0, 0}; const long long pt0 = clock64();
    const bool prof = p.prof != 0;
    auto twait = [&](uint64_t* bar, uint32_t parity, int slot) {
        if (prof) { const long long w0 = clock64(); mbar_wait(bar, parity); pc[slot] += clock64() - w0; } else mbar_wait(bar, parity);
    };
    const int my_tiles = (int)blockIdx.x < p.total_tiles ? (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    constexpr uint32_t TM_ACCA = 0, TM_T = 128, TM_ACCB = 256;

    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer =====================
            mbar_expect_tx(w_full, FB_W3_BYTES + Cfg::WDS_BYTES);
            tma_load_2d(&mapW3, w_full, sW3, 0, 0);
            if constexpr (DS) tma_load_2d(&mapWds, w_full, sWds, 0, 0);
            int hslot = 0; uint32_t hphase = 0; int ws = 0; uint32_t wphase = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int t = blockIdx.x + ti * gridDim.x;
                const int img = t / p.tiles_per_img, rem = t - img * p.tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                mbar_wait(&empty_h[hslot], hphase ^ 1);
                mbar_expect_tx(&full_h[hslot], HALO_BYTES_RAW);
                tma_load_4d(&mapA, &full_h[hslot], halo + hslot * HALO_BYTES, 0, tx * HALO_TW - 1, ty * HALO_TH - 1, img);
                if (++hslot == FB_HS) { hslot = 0; hphase ^= 1; }
                for (int tap = 0; tap < 9; ++tap) {
                    mbar_wait(&empty_w[ws], wphase ^ 1);
                    mbar_expect_tx(&full_w[ws], 8192);
                    tma_load_2d(&mapW2, &full_w[ws], sW2 + ws * 8192, tap * 64, 0);
                    if (++ws == FB_WS) { ws = 0; wphase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (DS && warp == 3) {
        // ===================== DS: TMA producer of the block-input patches (its own warp: never blocks the halo / weight stream) =====================
        if (elect_one()) {
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int t = blockIdx.x + ti * gridDim.x;
                const int img = t / p.tiles_per_img, rem = t - img * p.tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                mbar_wait(x_empty, (ti & 1) ^ 1);
                mbar_expect_tx(x_full, Cfg::X_BYTES);
                tma_load_4d(&mapX, x_full, sX, 0, tx * HALO_TW, ty * HALO_TH, img);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_a = make_idesc_bf16(64, 0, 0, GEMM_BM), idesc_b = make_idesc_bf16(256, 0, 0, GEMM_BM);
        constexpr uint32_t DESC_HI = (1u << 14) | (2u << 29);
        constexpr uint32_t HI_K = DESC_HI | (1024u >> 4), HI_HALO = DESC_HI | ((HALO_PITCH * 128u) >> 4);
        const uint32_t w2_lo = smem_u32(sW2) >> 4, w3_lo = smem_u32(sW3) >> 4, halo_lo = smem_u32(halo) >> 4;
        [[maybe_unused]] const uint32_t wds_lo = smem_u32(sWds) >> 4, x_lo = smem_u32(sX) >> 4;
        auto desc = [](uint32_t lo, uint32_t hi) { return (static_cast<uint64_t>(hi) << 32) | lo; };
        auto issue_b = [&](int j) {          // conv3 of tile j: A = intermediate in TMEM, B = resident weight tile
            const int ts = j & 1;
            twait(&t_full[ts], (j >> 1) & 1, 3);
            twait(accb_free, (j & 1) ^ 1, 4);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem_base + TM_ACCB, tmem_base + TM_T + ts * 32 + k * 8, desc(w3_lo + 2 * k, HI_K), idesc_b, k ? 1u : 0u);
                if constexpr (!DS) umma_commit(accb_full);
                umma_commit(&t_free[ts]);
            }
            __syncwarp();
            if constexpr (DS) {              // + downsample(x): A = the input patch in shared memory, same accumulator
                twait(x_full, j & 1, 5);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + TM_ACCB, desc(x_lo + 2 * k, HI_K), desc(wds_lo + 2 * k, HI_K), idesc_b, 1u);
                    umma_commit(accb_full);
                    umma_commit(x_empty);
                }
                __syncwarp();
            }
        };
        mbar_wait(w_full, 0);
        int hslot = 0; uint32_t hphase = 0; int ws = 0; uint32_t wphase = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int as = ti & 1;
            twait(&acca_free[as], ((ti >> 1) & 1) ^ 1, 0);
            twait(&full_h[hslot], hphase, 1);
            const uint32_t h_lo = halo_lo + hslot * (HALO_BYTES >> 4);
#pragma unroll 1
            for (int tap = 0, tr = 0, ts = 0; tap < 9; ++tap) {
                twait(&full_w[ws], wphase, 2);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = h_lo + (tr * HALO_PITCH + ts) * 8, b_lo = w2_lo + ws * 512;     // tap (r, s): the halo view shifted by (r, s) pixels
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + TM_ACCA + as * 64, desc(a_lo + 2 * k, HI_HALO), desc(b_lo + 2 * k, HI_K), idesc_a, (tap | k) ? 1u : 0u);
                    umma_commit(&empty_w[ws]);
                    if (tap == 8) { umma_commit(&empty_h[hslot]); umma_commit(&acca_full[as]); }
                }
                __syncwarp();
                if (++ws == FB_WS) { ws = 0; wphase ^= 1; }
                if (++ts == 3) { ts = 0; ++tr; }
            }
            if (++hslot == FB_HS) { hslot = 0; hphase ^= 1; }
            if (ti >= 1) issue_b(ti - 1);
        }
        if (my_tiles > 0) issue_b(my_tiles - 1);
    } else if (warp >= 4) {
        // ===================== epilogues (warps 4-11) =====================
        const int q = warp & 3, grp = warp >= 8 ? 0 : 1, cc0 = grp * 2;          // epilogue-B channel chunks cc0, cc0 + 1 of every tile
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // ---- epilogue B state: the TMA epilogue of convgemm.cuh, two 64-channel items per tile and warp
        uint8_t* myring = ring + (grp * 4 + q) * FB_D * FB_SUB_BYTES;
        uint64_t* myfull = res_full + (grp * 4 + q) * FB_D;
        const float4* bias4 = reinterpret_cast<const float4*>(tab + 64);
        const int n_items = my_tiles * 2;
        const uint32_t rowoff = (uint32_t)lane * 128u, swz = (uint32_t)(lane & 7);
        auto coords = [&](int ti, int& img, int& ox0, int& row0) {
            const int t = blockIdx.x + ti * gridDim.x;
            img = t / p.tiles_per_img; const int rem = t - img * p.tiles_per_img;
            const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
            ox0 = tx * HALO_TW; row0 = ty * HALO_TH + q * 4;
        };
        int pf_j = 0, pf_c = 0, pf_slot = 0, pf_ti = 0;                           // residual prefetch cursor (lane 0)
        auto request_next = [&]() {
            int img, ox0, row0; coords(pf_ti, img, ox0, row0);
            mbar_expect_tx(&myfull[pf_slot], FB_SUB_BYTES);
            tma_load_4d(&mapR, &myfull[pf_slot], myring + pf_slot * FB_SUB_BYTES, (cc0 + pf_c) * 64, ox0, row0, img);
            ++pf_j; if (++pf_slot == FB_D) pf_slot = 0;
            if (++pf_c == 2) { pf_c = 0; ++pf_ti; }
        };
        if constexpr (!DS) {
            if (lane == 0)
                for (int i = 0; i < FB_D - 1 && i < n_items; ++i) request_next();
        }
        int j = 0, slot = 0; uint32_t sphase = 0;
        auto epilogue_b = [&](int ti) {
            int img, ox0, row0; coords(ti, img, ox0, row0);
            twait(accb_full, ti & 1, 2);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 2; ++c, ++j) {
                const int cc = cc0 + c;
                uint8_t* sl = myring + slot * FB_SUB_BYTES;
                uint8_t* rowp = sl + rowoff;
                float v[64];
                tmem_ld32(lane_base + TM_ACCB + cc * 64, v);
                tmem_ld32(lane_base + TM_ACCB + cc * 64 + 32, v + 32);
                uint4 r[8];
                if constexpr (!DS) {
                    twait(&myfull[slot], sphase, 3);
#pragma unroll
                    for (int k = 0; k < 8; ++k) r[k] = *reinterpret_cast<const uint4*>(rowp + ((k ^ swz) << 4));     // 128B swizzle: chunk ^ (row % 8)
                } else {
                    if (lane == 0) bulk_wait_group_read<FB_D - 1>();       // the store that last used this slot has read it
                    __syncwarp();
                }
                tmem_ld_wait();
                if (c == 1) {                       // this warp's share of the accumulator is in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(accb_free);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float* f = v + 8 * k;
                    const float4 b0 = bias4[cc * 16 + 2 * k], b1 = bias4[cc * 16 + 2 * k + 1];
                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                    if constexpr (!DS) {
                        const uint32_t w[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            f[2 * u] += __uint_as_float(w[u] << 16);
                            f[2 * u + 1] += __uint_as_float(w[u] & 0xffff0000u);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) f[u] = fmaxf(f[u], 0.0f);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float* f = v + 8 * k;
                    __nv_bfloat162 h2;
                    h2 = __floats2bfloat162_rn(f[0], f[1]); r[k].x = *reinterpret_cast<uint32_t*>(&h2);
                    h2 = __floats2bfloat162_rn(f[2], f[3]); r[k].y = *reinterpret_cast<uint32_t*>(&h2);
                    h2 = __floats2bfloat162_rn(f[4], f[5]); r[k].z = *reinterpret_cast<uint32_t*>(&h2);
                    h2 = __floats2bfloat162_rn(f[6], f[7]); r[k].w = *reinterpret_cast<uint32_t*>(&h2);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(rowp + ((k ^ swz) << 4)) = r[k];
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&mapO, sl, cc * 64, ox0, row0, img);      // 64 channels x 8 pixels x 4 output rows
                    bulk_commit_group();
                    if constexpr (!DS) {
                        if (pf_j < n_items) {
                            if (j >= 1) bulk_wait_group_read<1>();              // the store of item j-1 has read the slot being refilled
                            request_next();
                        }
                    }
                }
                __syncwarp();
                if (++slot == FB_D) { slot = 0; sphase ^= 1; }
            }
        };
        if (grp == 0) {
            for (int ti = 0; ti < my_tiles; ++ti) epilogue_b(ti);
        } else {
            // ---- epilogue A: conv2 accumulator -> (+ shift, ReLU, bf16) -> intermediate in TMEM; then chunks 2, 3 of the previous tile
            const float4* sh4 = reinterpret_cast<const float4*>(tab);
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int as = ti & 1;
                twait(&acca_full[as], (ti >> 1) & 1, 0);
                tc_fence_after();
                {
                    float v[64];
                    tmem_ld32(lane_base + TM_ACCA + as * 64, v);
                    tmem_ld32(lane_base + TM_ACCA + as * 64 + 32, v + 32);
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acca_free[as]);
                    uint32_t pk[32];
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float4 s4 = sh4[k];
                        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[2 * k]) : "f"(v[4 * k + 1] + s4.y), "f"(v[4 * k] + s4.x));
                        asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[2 * k + 1]) : "f"(v[4 * k + 3] + s4.w), "f"(v[4 * k + 2] + s4.z));
                    }
                    twait(&t_free[as], ((ti >> 1) & 1) ^ 1, 1);
                    tc_fence_after();
                    tmem_st32(lane_base + TM_T + as * 32, pk);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_full[as]);
                }
                if (ti >= 1) epilogue_b(ti - 1);
            }
            if (my_tiles > 0) epilogue_b(my_tiles - 1);
        }
        if (lane == 0) bulk_wait_group_read<0>();
    }
    if (prof && blockIdx.x == 0 && lane == 0) {
        if (warp == 1) printf("bneck prof mma: total %lld; waits acca_free %lld full_h %lld full_w %lld t_full %lld accb_free %lld x_full %lld\n", clock64() - pt0, pc[0], pc[1], pc[2], pc[3], pc[4], pc[5]);
        if (warp == 4) printf("bneck prof epiA+B(2,3): total %lld; waits acca_full %lld t_free %lld accb_full %lld residual %lld\n", clock64() - pt0, pc[0], pc[1], pc[2], pc[3]);
        if (warp == 8) printf("bneck prof epiB(0,1): total %lld; waits accb_full %lld residual %lld (tiles %d)\n", clock64() - pt0, pc[2], pc[3], my_tiles);
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// in: [batch, H, W, 64] (conv1's output), res / out: [batch, H, W, 256]; w2: [64][3][3][64], w3: [256][64] (BN scale folded)
inline bool bneck_l1_supported(int H, int W) { return H % HALO_TH == 0 && W % HALO_TW == 0; }

// res: residual [batch, H, W, 256] (DS = 0), or null with x / wds / shiftds: the block input [batch, H, W, 64] and the downsample weights [256][64] (DS = 1)
template <int DS>
int bneck_l1_launch_inst(const bf16* in, const bf16* w2, const float* shift2, const bf16* w3, const float* shift3, const bf16* res, bf16* out,
                         const bf16* x, const bf16* wds, const float* shiftds, int batch, int H, int W, cudaStream_t st) {
    using Cfg = FbCfg<DS>;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(bneck_l1_kernel<DS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (err != cudaSuccess) { set_error("bneck_l1: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return -1; }
        configured = true;
    }
    FbParams kp;
    kp.batch = batch; kp.tiles_x = W / HALO_TW; kp.tiles_per_img = (H / HALO_TH) * kp.tiles_x; kp.total_tiles = batch * kp.tiles_per_img;
    kp.shift2 = shift2; kp.shift3 = shift3; kp.shiftds = shiftds;
    static const int prof_env = getenv("SQ_BNECK_PROF") ? atoi(getenv("SQ_BNECK_PROF")) : 0;
    kp.prof = prof_env;
    CUtensorMap maps[7];
    {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
        cuuint64_t strides[3] = {64 * 2, (cuuint64_t)W * 64 * 2, (cuuint64_t)H * W * 64 * 2};
        cuuint32_t box[4] = {64, HALO_PITCH, HALO_TH + 2, 1}, estr[4] = {1, 1, 1, 1};
        if (encode_map(&maps[0], in, 4, dims, strides, box, estr)) return -1;
    }
    {
        cuuint64_t dims[2] = {576, 64}, strides[1] = {576 * 2};
        cuuint32_t box[2] = {64, 64}, estr[2] = {1, 1};
        if (encode_map(&maps[1], w2, 2, dims, strides, box, estr)) return -1;
    }
    {
        cuuint64_t dims[2] = {64, 256}, strides[1] = {64 * 2};
        cuuint32_t box[2] = {64, 256}, estr[2] = {1, 1};
        if (encode_map(&maps[2], w3, 2, dims, strides, box, estr)) return -1;
    }
    {
        cuuint64_t dims[4] = {256, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
        cuuint64_t strides[3] = {256 * 2, (cuuint64_t)W * 256 * 2, (cuuint64_t)H * W * 256 * 2};
        cuuint32_t box[4] = {64, HALO_TW, 4, 1}, estr[4] = {1, 1, 1, 1};
        if (encode_map(&maps[4], out, 4, dims, strides, box, estr)) return -1;
        if (DS) maps[3] = maps[4]; else if (encode_map(&maps[3], res, 4, dims, strides, box, estr)) return -1;
    }
    if (DS) {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)batch};
        cuuint64_t strides[3] = {64 * 2, (cuuint64_t)W * 64 * 2, (cuuint64_t)H * W * 64 * 2};
        cuuint32_t box[4] = {64, HALO_TW, HALO_TH, 1}, estr[4] = {1, 1, 1, 1};
        if (encode_map(&maps[5], x, 4, dims, strides, box, estr)) return -1;
        cuuint64_t wd[2] = {64, 256}, ws[1] = {64 * 2};
        cuuint32_t wb[2] = {64, 256}, we[2] = {1, 1};
        if (encode_map(&maps[6], wds, 2, wd, ws, wb, we)) return -1;
    } else { maps[5] = maps[0]; maps[6] = maps[2]; }
    const int grid = kp.total_tiles < num_sms() ? kp.total_tiles : num_sms();
    static const int pdl = getenv("SQ_PDL") ? atoi(getenv("SQ_PDL")) : 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    const double M = (double)batch * H * W;
    gemm_timing_begin(st, 2.0 * M * 64 * 576 + 2.0 * M * 256 * 64 * (DS ? 2 : 1));
    cudaError_t err = cudaLaunchKernelEx(&cfg, bneck_l1_kernel<DS>, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], kp);
    if (err == cudaSuccess) err = cudaGetLastError();
    gemm_timing_end(st);
    if (err != cudaSuccess) { set_error("bneck_l1 launch: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

inline int bneck_l1_launch(const bf16* in, const bf16* w2, const float* shift2, const bf16* w3, const float* shift3, const bf16* res, bf16* out,
                           int batch, int H, int W, cudaStream_t st) {
    return bneck_l1_launch_inst<0>(in, w2, shift2, w3, shift3, res, out, nullptr, nullptr, nullptr, batch, H, W, st);
}
// first block of the layer: out = relu(conv3(relu(conv2(in))) + downsample(x)), the downsample computed inside the kernel
inline int bneck_l1_ds_launch(const bf16* in, const bf16* w2, const float* shift2, const bf16* w3, const float* shift3, const bf16* x, const bf16* wds,
                              const float* shiftds, bf16* out, int batch, int H, int W, cudaStream_t st) {
    return bneck_l1_launch_inst<1>(in, w2, shift2, w3, shift3, nullptr, out, x, wds, shiftds, batch, H, W, st);
}

}  // namespace sq
