// bf16 -> bf16 convolution / GEMM kernel of the feature extractors (ResNet-50 bottleneck convolutions, src/resnet.py:73-93):
//
//   out[M, N] = act( A[M, K] * W[N, K]^T + shift[N] (+ residual[M, N]) )        bf16 operands, fp32 accumulation in TMEM
//
// Same warp-specialised tcgen05 skeleton as gemm.cuh, re-designed around the two things the round-1 profile showed to be
// the limits of these launches (profiles/r01_resnet_per_conv_efficiency.txt):
//
//  * CTA PAIRS (`cta_group::2`, cluster of two CTAs on one TPC): one tcgen05.mma covers a 256 x BN tile, each CTA stages
//    its own 128 rows of A and only HALF of the weight tile, so the L2 -> shared-memory operand traffic per FLOP drops by a
//    third and the pipeline gets deeper stages for the same shared memory.  Only the leader CTA issues MMAs; both CTAs'
//    TMA loads complete on the leader's "full" barrier; tcgen05.commit multicasts "stage free" / "accumulator ready" to both.
//  * TMA EPILOGUE: every epilogue warp owns a ring of D sub-tiles (32 rows x 64 channels, 4 KB, 128B-swizzled).  The residual
//    sub-tile of item j+D-1 is requested by TMA while item j is processed (the K <= 512 expansions were bound by the latency of
//    register loads of the residual), shift + residual + ReLU happen IN PLACE in that sub-tile, and the result leaves through
//    `cp.async.bulk.tensor` stores (UTMASTG) - no per-lane global loads / stores at all.
//
// A is either a plain [M, K] matrix (1x1 stride-1 convolutions) or an NHWC tensor walked as one 4-D TMA box per filter tap.
#pragma once
#include "gemm.cuh"

namespace sq {

struct CgParams {
    int M, N;
    int num_n;           // BN-wide column tiles
    int total_tiles;     // (pairs of) 128-row tiles x column tiles
    int nk;              // 64-wide k-blocks
    int conv, cblocks, S, stride, pad, tiles_per_img, BH, BIMG;
    int tiles_x, Ho, halo_bo;   // halo mode: 8-pixel-wide tiles per output row, output height, descriptor base-offset mode
    const float* bias;   // [N] folded BatchNorm shift
    int relu, has_res;   // relu: activation, 0 none, 1 ReLU, 2 exact GELU
    float* pool_out;     // non-null: instead of storing the tile, add mean over the top-left 7x7 of each 8x8 image map into pool_out[img][N] (src/resnet.py:110,166)
    int batch;
    int l2pf;            // prefetch residual sub-tiles into L2 two tiles ahead
    unsigned long long* prof;   // optional per-CTA cycle counters (sq_gemm_profile): [cta][16]
};

// HALO = 1 (3x3, stride 1, pad 1): the CTA's 128 output pixels are a 16 x 8 patch; its 18 x 10 input halo is staged ONCE per
// 64-channel block (one 4-D TMA box, 23 KB) and the nine taps are issued from shifted views of it: tap (r, s) reads pixel
// rows (oy + r, ox + s), i.e. a K-major operand whose 8-row groups (one output row = 8 consecutive halo pixels) are 10
// pixels = 1280 bytes apart (the descriptor's stride-byte-offset) and which starts (r * 10 + s) * 128 bytes into the halo.
// The stage ring then only carries weight tiles.  L2 -> shared-memory traffic for A drops from 9x to 1.4x the input.
constexpr int HALO_TH = 16, HALO_TW = 8, HALO_PITCH = HALO_TW + 2;
constexpr int HALO_BYTES_RAW = (HALO_TH + 2) * HALO_PITCH * 128;
constexpr int HALO_BYTES = (HALO_BYTES_RAW + 1023) / 1024 * 1024;
template <int BN, int CG, int D, int HALO = 0> struct CgCfg {
    static constexpr int B_ROWS = BN / CG;                        // rows of the weight tile this CTA stages
    static constexpr int A_BYTES = HALO ? 0 : GEMM_BM * GEMM_BK * 2;
    static constexpr int HS = HALO ? 3 : 0;                       // halo slots
    static constexpr int B_BYTES = B_ROWS * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NCHUNK = BN / 64;                        // 64-column epilogue items per tile
    static constexpr int CPH = NCHUNK >= 2 ? NCHUNK / 2 : 1;      // items per tile of one epilogue warp
    static constexpr int NWORK = NCHUNK >= 2 ? 8 : 4;             // epilogue warps with work (2 per TMEM lane quadrant when BN >= 128)
    static constexpr int SUB_BYTES = 32 * 128;
    static constexpr int RING_BYTES = NWORK * D * SUB_BYTES;
    static constexpr int BAR_BYTES = 512;
    static constexpr int BIAS_BYTES = 2 * BN * 4;                 // shift table, double-buffered like the accumulator
    static constexpr int BUDGET = 232448 - BAR_BYTES - BIAS_BYTES - RING_BYTES - HS * HALO_BYTES;
    static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
    static_assert(STAGES >= 2, "shared memory budget");
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + HS * HALO_BYTES + RING_BYTES + BIAS_BYTES + BAR_BYTES;
    static_assert((2 * STAGES + 4 + NWORK * D + 2 * HS) * 8 + 4 <= BAR_BYTES, "barrier area");
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
};

template <int BN, int CG, int D, int HALO, int GELU = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
convgemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapR, const __grid_constant__ CUtensorMap mapO, const CgParams p) {
    using Cfg = CgCfg<BN, CG, D, HALO>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];              // 128B-swizzled tiles need 1024-byte alignment (checked below)
    uint8_t* halo = smem + STAGES * Cfg::STAGE_BYTES;
    uint8_t* ring = halo + Cfg::HS * HALO_BYTES;
    float* bias_smem = reinterpret_cast<float*>(ring + Cfg::RING_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + Cfg::RING_BYTES + Cfg::BIAS_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint64_t* res_full = bars + 2 * STAGES + 4;                   // [NWORK][D]
    uint64_t* full_h = res_full + Cfg::NWORK * D;                 // [HS] halo slot filled / drained
    uint64_t* empty_h = full_h + Cfg::HS;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(empty_h + Cfg::HS);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;

    asm volatile("griddepcontrol.launch_dependents;");
    if (warp == 0 && lane == 0) {
        if (smem_u32(smem) & 1023u) { printf("sequoia_b200: dynamic shared memory is not 1024-byte aligned\n"); __trap(); }
        tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); tma_prefetch_desc(&mapR); tma_prefetch_desc(&mapO);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], CG * Cfg::NWORK); }
        for (int i = 0; i < Cfg::NWORK * D; ++i) mbar_init(&res_full[i], 1);
        for (int i = 0; i < Cfg::HS; ++i) { mbar_init(&full_h[i], 1); mbar_init(&empty_h[i], 1); }
        mbar_fence_init();
    }
    if (warp == 2) { if constexpr (CG == 2) tmem_alloc_pair(tmem_ptr, Cfg::TMEM_COLS); else tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); }
    __syncwarp();
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();      // the peer's barriers are initialised too
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const int my_tiles = cluster_id < p.total_tiles ? (p.total_tiles - cluster_id + num_clusters - 1) / num_clusters : 0;

    if (warp == 0) {
        if (elect_one()) {
            // ===================== TMA producer (both CTAs of a pair; completion on the leader's barrier) =====================
            int stage = 0; uint32_t phase = 0;
            [[maybe_unused]] int hslot = 0; [[maybe_unused]] uint32_t hphase = 0;
            long long pw = 0; const long long pt0 = clock64();
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int t = cluster_id + ti * num_clusters;
                const int pm = t / p.num_n, nt = t - pm * p.num_n;
                const int mt = pm * CG + (int)rank;
                const int m0 = mt * GEMM_BM, nb0 = nt * BN + (int)rank * Cfg::B_ROWS;
                int img = 0, hin0 = 0;
                if (p.conv) {
                    if (p.BIMG == 1) { img = mt / p.tiles_per_img; hin0 = (mt - img * p.tiles_per_img) * p.BH * p.stride - p.pad; }
                    else { img = mt * p.BIMG; hin0 = -p.pad; }
                }
                if constexpr (HALO) {
                    // tile = 16 x 8 output pixels of one image; per channel block: its halo once, then the nine weight tiles
                    const int img_h = mt / p.tiles_per_img, rem = mt - img_h * p.tiles_per_img;
                    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                    for (int cbh = 0; cbh < p.cblocks; ++cbh) {
                        mbar_wait(&empty_h[hslot], hphase ^ 1);
                        uint8_t* sh = halo + hslot * HALO_BYTES;
                        if constexpr (CG == 2) {
                            if (rank == 0) mbar_expect_tx(&full_h[hslot], 2 * HALO_BYTES_RAW);
                            tma_load_4d_pair(&mapA, mapa_u32(&full_h[hslot], 0), sh, cbh * GEMM_BK, tx * HALO_TW - 1, ty * HALO_TH - 1, img_h);
                        } else {
                            mbar_expect_tx(&full_h[hslot], HALO_BYTES_RAW);
                            tma_load_4d(&mapA, &full_h[hslot], sh, cbh * GEMM_BK, tx * HALO_TW - 1, ty * HALO_TH - 1, img_h);
                        }
                        if (++hslot == Cfg::HS) { hslot = 0; hphase ^= 1; }
                        for (int tap = 0; tap < 9; ++tap) {
                            if (p.prof) { const long long w0 = clock64(); mbar_wait(&empty[stage], phase ^ 1); pw += clock64() - w0; }
                            else mbar_wait(&empty[stage], phase ^ 1);
                            uint8_t* sb = smem + stage * Cfg::STAGE_BYTES;
                            const int kcol = (tap * p.cblocks + cbh) * GEMM_BK;
                            if constexpr (CG == 2) {
                                if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                                tma_load_2d_pair(&mapB, mapa_u32(&full[stage], 0), sb, kcol, nb0);
                            } else {
                                mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                                tma_load_2d(&mapB, &full[stage], sb, kcol, nb0);
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                } else {
                int tap_r = 0, tap_s = 0, cb = 0;
                for (int kb = 0; kb < p.nk; ++kb) {
                    if (p.prof) { const long long w0 = clock64(); mbar_wait(&empty[stage], phase ^ 1); pw += clock64() - w0; }
                    else mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if constexpr (CG == 2) {
                        if (rank == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                        const uint32_t bar = mapa_u32(&full[stage], 0);
                        if (p.conv) tma_load_4d_pair(&mapA, bar, sa, cb * GEMM_BK, tap_s - p.pad, hin0 + tap_r, img);
                        else tma_load_2d_pair(&mapA, bar, sa, kb * GEMM_BK, m0);
                        tma_load_2d_pair(&mapB, bar, sb, kb * GEMM_BK, nb0);
                    } else {
                        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        if (p.conv) tma_load_4d(&mapA, &full[stage], sa, cb * GEMM_BK, tap_s - p.pad, hin0 + tap_r, img);
                        else tma_load_2d(&mapA, &full[stage], sa, kb * GEMM_BK, m0);
                        tma_load_2d(&mapB, &full[stage], sb, kb * GEMM_BK, nb0);
                    }
                    if (p.conv && ++cb == p.cblocks) { cb = 0; if (++tap_s == p.S) { tap_s = 0; ++tap_r; } }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                }
            }
            if (p.prof) { p.prof[blockIdx.x * 16 + 0] = pw; p.prof[blockIdx.x * 16 + 1] = clock64() - pt0; }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===================== MMA issuer (leader CTA) =====================
            // A 64- or 128-wide MMA is 32 / 64 cycles of tensor-pipe time: the issuing thread must not spend more than that per
            // instruction, so descriptors are not rebuilt per k-step - only their low word (address >> 4) moves, by additions
            // from a running stage / tap offset, and the cycle counters are read only when profiling is on.
            const uint32_t idesc = make_idesc_bf16(BN, 0, 0, GEMM_BM * CG);
            constexpr uint32_t DESC_HI = (1u << 14) | (2u << 29);                              // descriptor version, SWIZZLE_128B
            constexpr uint32_t HI_K = DESC_HI | (1024u >> 4);                                  // K-major tile: 8-row groups 1024 B apart
            [[maybe_unused]] constexpr uint32_t HI_HALO = DESC_HI | ((HALO_PITCH * 128u) >> 4);    // halo view: 8-pixel groups one halo row apart
            const uint32_t smem_lo = smem_u32(smem) >> 4;
            [[maybe_unused]] const uint32_t halo_lo = smem_u32(halo) >> 4;
            auto desc = [](uint32_t lo, uint32_t hi) { return (static_cast<uint64_t>(hi) << 32) | lo; };
            const bool prof_on = p.prof != nullptr;
            int stage = 0; uint32_t phase = 0; uint32_t stage_lo = smem_lo;
            [[maybe_unused]] int hslot = 0; [[maybe_unused]] uint32_t hphase = 0;
            long long mw_full = 0, mw_tmem = 0; const long long mt0 = prof_on ? clock64() : 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const int as = ti & 1; const uint32_t aphase = (ti >> 1) & 1;
                if (prof_on) { const long long w1 = clock64(); mbar_wait(&tmem_empty[as], aphase ^ 1); mw_tmem += clock64() - w1; }
                else mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + as * BN;
                if constexpr (HALO) {
                    for (int cbh = 0; cbh < p.cblocks; ++cbh) {
                        if (prof_on) { const long long w2 = clock64(); mbar_wait(&full_h[hslot], hphase); mw_full += clock64() - w2; }
                        else mbar_wait(&full_h[hslot], hphase);
                        uint32_t a_lo = halo_lo + hslot * (HALO_BYTES >> 4);               // tap (r, s): + (r * HALO_PITCH + s) * 128 bytes
                        for (int tap = 0, ts = 0; tap < 9; ++tap) {
                            if (prof_on) { const long long w2 = clock64(); mbar_wait(&full[stage], phase); mw_full += clock64() - w2; }
                            else mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            if (elect_one()) {
                                if (p.halo_bo) {
                                    const uint32_t a_base = a_lo << 4, b_base = stage_lo << 4, bo = (a_base >> 7) & 7u;
#pragma unroll
                                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                                        const uint32_t acc = (cbh > 0 || tap > 0 || k > 0) ? 1u : 0u;
                                        const uint64_t da = make_smem_desc(a_base + k * 32, HALO_PITCH * 128, 0, bo), db = make_smem_desc(b_base + k * 32, 1024, 0);
                                        if constexpr (CG == 2) umma_bf16_pair(tacc, da, db, idesc, acc); else umma_bf16(tacc, da, db, idesc, acc);
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                                        const uint32_t acc = (cbh > 0 || tap > 0 || k > 0) ? 1u : 0u;
                                        if constexpr (CG == 2) umma_bf16_pair(tacc, desc(a_lo + 2 * k, HI_HALO), desc(stage_lo + 2 * k, HI_K), idesc, acc);
                                        else umma_bf16(tacc, desc(a_lo + 2 * k, HI_HALO), desc(stage_lo + 2 * k, HI_K), idesc, acc);
                                    }
                                }
                                const bool last = cbh == p.cblocks - 1 && tap == 8;
                                if constexpr (CG == 2) {
                                    umma_commit_pair(&empty[stage]);
                                    if (tap == 8) umma_commit_pair(&empty_h[hslot]);
                                    if (last) umma_commit_pair(&tmem_full[as]);
                                } else {
                                    umma_commit(&empty[stage]);
                                    if (tap == 8) umma_commit(&empty_h[hslot]);
                                    if (last) umma_commit(&tmem_full[as]);
                                }
                            }
                            __syncwarp();
                            a_lo += 8; if (++ts == 3) { ts = 0; a_lo += (HALO_PITCH - 3) * 8; }
                            stage_lo += Cfg::STAGE_BYTES >> 4;
                            if (++stage == STAGES) { stage = 0; phase ^= 1; stage_lo = smem_lo; }
                        }
                        if (++hslot == Cfg::HS) { hslot = 0; hphase ^= 1; }
                    }
                } else {
                for (int kb = 0; kb < p.nk; ++kb) {
                    if (prof_on) { const long long w2 = clock64(); mbar_wait(&full[stage], phase); mw_full += clock64() - w2; }
                    else mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b_lo = stage_lo + (Cfg::A_BYTES >> 4);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            if constexpr (CG == 2) umma_bf16_pair(tacc, desc(stage_lo + 2 * k, HI_K), desc(b_lo + 2 * k, HI_K), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                            else umma_bf16(tacc, desc(stage_lo + 2 * k, HI_K), desc(b_lo + 2 * k, HI_K), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        if constexpr (CG == 2) { umma_commit_pair(&empty[stage]); if (kb == p.nk - 1) umma_commit_pair(&tmem_full[as]); }
                        else { umma_commit(&empty[stage]); if (kb == p.nk - 1) umma_commit(&tmem_full[as]); }
                    }
                    __syncwarp();
                    stage_lo += Cfg::STAGE_BYTES >> 4;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; stage_lo = smem_lo; }
                }
                }
            }
            if (prof_on && lane == 0) { p.prof[blockIdx.x * 16 + 2] = mw_full; p.prof[blockIdx.x * 16 + 3] = mw_tmem; p.prof[blockIdx.x * 16 + 4] = clock64() - mt0; }
        }
    } else if (warp >= 4 && warp - 4 < Cfg::NWORK) {
        // ===================== epilogue: TMEM -> (+ shift, + residual, ReLU) in the warp's ring -> TMA store =====================
        // Two epilogue warps share a scheduler, so latency is hidden by instruction-level parallelism inside an item: all
        // residual reads of the item are issued together, the shift comes from a shared-memory table (broadcast reads), the
        // packed results are written together, and tile / item coordinates are tracked incrementally (no divisions per item).
        const int ew = warp - 4, q = warp & 3, half = ew >> 2;
        uint8_t* myring = ring + ew * D * Cfg::SUB_BYTES;
        uint64_t* myfull = res_full + ew * D;
        const bool has_res = p.has_res != 0;
        const float relu_lo = p.relu == 1 ? 0.0f : -INFINITY;
        const int n_items = my_tiles * Cfg::CPH;
        const uint32_t rowoff = (uint32_t)lane * 128u, swz = (uint32_t)(lane & 7);
        // residual prefetch cursor (lane 0): item pf_j = (tile pf_ti, chunk pf_cc), ring slot pf_slot
        int pf_j = 0, pf_cc = 0, pf_slot = 0, pf_pm = 0, pf_nt = 0, pf_t = cluster_id;
        if (my_tiles > 0) { pf_pm = pf_t / p.num_n; pf_nt = pf_t - pf_pm * p.num_n; }
        auto request_next = [&]() {
            const int row0 = (pf_pm * CG + (int)rank) * GEMM_BM + q * 32, col0 = pf_nt * BN + (half * Cfg::CPH + pf_cc) * 64;
            mbar_expect_tx(&myfull[pf_slot], Cfg::SUB_BYTES);
            tma_load_2d(&mapR, &myfull[pf_slot], myring + pf_slot * Cfg::SUB_BYTES, col0, row0);
            ++pf_j; if (++pf_slot == D) pf_slot = 0;
            if (++pf_cc == Cfg::CPH) { pf_cc = 0; pf_t += num_clusters; pf_pm = pf_t / p.num_n; pf_nt = pf_t - pf_pm * p.num_n; }
        };
        if (has_res && lane == 0)
            for (int i = 0; i < D - 1 && i < n_items; ++i) request_next();
        const uint32_t tmem_empty_leader = (CG == 2) ? mapa_u32(tmem_empty, 0) : 0u;
        int j = 0, slot = 0; uint32_t sphase = 0;
        long long ew_acc = 0, ew_res = 0, ew_grp = 0; const long long et0 = clock64();
        const bool prof_on = p.prof != nullptr;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int t = cluster_id + ti * num_clusters;
            const int pm = t / p.num_n, nt = t - pm * p.num_n;
            const int row0 = (pm * CG + (int)rank) * GEMM_BM + q * 32;
            [[maybe_unused]] int h_ox0 = 0, h_row0 = 0, h_img = 0;
            if constexpr (HALO) {
                const int mt = pm * CG + (int)rank, rem = mt - (mt / p.tiles_per_img) * p.tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                h_img = mt / p.tiles_per_img; h_ox0 = tx * HALO_TW; h_row0 = ty * HALO_TH + q * 4;
            }
            const int as = ti & 1; const uint32_t aphase = (ti >> 1) & 1;
            // this warp's CPH*64 shift values of the tile: requested before the accumulator wait, published after it (all warps
            // of the half write the same values; the buffer of tile ti was last read for tile ti-2, which every warp of the
            // pair had finished before this tile's MMAs could start)
            if (has_res && p.l2pf && lane == 0 && ti + 2 < my_tiles) {          // residual sub-tiles of the tile after next: HBM -> L2 now
                const int t2 = t + 2 * num_clusters, pm2 = t2 / p.num_n, nt2 = t2 - pm2 * p.num_n;
#pragma unroll
                for (int cc2 = 0; cc2 < Cfg::CPH; ++cc2)
                    tma_prefetch_l2_2d(&mapR, nt2 * BN + (half * Cfg::CPH + cc2) * 64, (pm2 * CG + (int)rank) * GEMM_BM + q * 32);
            }
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < Cfg::CPH * 16) bv = __ldg(reinterpret_cast<const float4*>(p.bias + nt * BN + half * Cfg::CPH * 64) + lane);
            if (prof_on) { const long long w3 = clock64(); mbar_wait(&tmem_full[as], aphase); ew_acc += clock64() - w3; }
            else mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            float4* bias4 = reinterpret_cast<float4*>(bias_smem) + (as * BN + half * Cfg::CPH * 64) / 4;
            if (lane < Cfg::CPH * 16) bias4[lane] = bv;
            __syncwarp();
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
            for (int cc = 0; cc < Cfg::CPH; ++cc, ++j) {
                const int c = half * Cfg::CPH + cc;
                uint8_t* sl = myring + slot * Cfg::SUB_BYTES;
                uint8_t* rowp = sl + rowoff;
                float v[64];
                tmem_ld32(tacc + c * 64, v);
                tmem_ld32(tacc + c * 64 + 32, v + 32);
                uint4 r[8];
                const long long w4 = prof_on ? clock64() : 0;
                if (has_res) {
                    mbar_wait(&myfull[slot], sphase);
#pragma unroll
                    for (int k = 0; k < 8; ++k) r[k] = *reinterpret_cast<const uint4*>(rowp + ((k ^ swz) << 4));     // 128B swizzle: chunk ^ (row % 8)
                } else {
                    if (lane == 0) bulk_wait_group_read<D - 1>();       // the store that last used this slot has read it
                    __syncwarp();
                }
                if (prof_on) ew_res += clock64() - w4;
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float* f = v + 8 * k;
                    const float4 b0 = bias4[cc * 16 + 2 * k], b1 = bias4[cc * 16 + 2 * k + 1];
                    f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
                    if (has_res) {
                        const uint32_t w[4] = {r[k].x, r[k].y, r[k].z, r[k].w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            f[2 * u] += __uint_as_float(w[u] << 16);
                            f[2 * u + 1] += __uint_as_float(w[u] & 0xffff0000u);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) f[u] = fmaxf(f[u], relu_lo);
                }
                if constexpr (GELU) {                        // exact GELU (UNI's fc1: timm nn.GELU): its own instantiation, the ReLU kernels carry no erff code
#pragma unroll
                    for (int u = 0; u < 64; ++u) v[u] = gelu_f(v[u]);
                }
                if (p.pool_out) {
                    // fused AvgPool2d(7) of the final 8x8 map: a 128-row tile is two images, this warp's 32 rows are image rows
                    // 4*(q&1) .. +3; masked recursive-halving reduction over the lanes leaves columns 2*lane, 2*lane+1 here; the two
                    // half-image partial means are combined with atomicAdd onto a zeroed output (two addends: order-independent)
                    const bool inside = ((q & 1) * 4 + (lane >> 3) < 7) && ((lane & 7) < 7);
#pragma unroll
                    for (int i = 0; i < 64; ++i) v[i] = inside ? v[i] : 0.0f;
#pragma unroll
                    for (int step = 0; step < 5; ++step) {
                        const int width = 32 >> step, mask = 16 >> step;
                        const bool upper = (lane & mask) != 0;
#pragma unroll
                        for (int i = 0; i < width; ++i) {
                            const float send = upper ? v[i] : v[i + width];
                            const float keep = upper ? v[i + width] : v[i];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
                        }
                    }
                    const int img = (pm * CG + (int)rank) * 2 + (q >> 1);
                    if (img < p.batch) {
                        float* o = p.pool_out + (size_t)img * p.N + nt * BN + c * 64 + 2 * lane;
                        atomicAdd(o, v[0] * (1.0f / 49.0f));
                        atomicAdd(o + 1, v[1] * (1.0f / 49.0f));
                    }
                    if (lane == 0 && has_res && pf_j < n_items) request_next();
                    __syncwarp();
                } else {
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
                    if constexpr (HALO) tma_store_4d(&mapO, sl, nt * BN + c * 64, h_ox0, h_row0, h_img);     // 4 output rows x 8 pixels (clipped at the map edge)
                    else tma_store_2d(&mapO, sl, nt * BN + c * 64, row0);
                    bulk_commit_group();
                    if (has_res && pf_j < n_items) {
                        const long long w5 = prof_on ? clock64() : 0;
                        if (j >= 1) bulk_wait_group_read<1>();          // the store of item j-1 has read the slot being refilled
                        if (prof_on) ew_grp += clock64() - w5;
                        request_next();
                    }
                }
                __syncwarp();
                }
                if (++slot == D) { slot = 0; sphase ^= 1; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CG == 2) mbar_arrive_cluster(tmem_empty_leader + as * 8); else mbar_arrive(&tmem_empty[as]); }
        }
        if (lane == 0) bulk_wait_group_read<0>();
        // counters of epilogue warp 0 (lane 0): accumulator wait, residual / slot wait, store-read wait, total
        if (prof_on && lane == 0 && ew == 0) {
            unsigned long long* o = p.prof + blockIdx.x * 16 + 5;
            o[0] = ew_acc; o[1] = ew_res; o[2] = ew_grp; o[3] = clock64() - et0;
        }
    }
    __syncwarp();                 // the single-lane roles rejoin their warps before the aligned barrier
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) { if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS); else tmem_dealloc(tmem_base, Cfg::TMEM_COLS); }
}

// ------------------------------------------------------------------------------------------------ host side
struct ConvGemmArgs {
    int M, N, K;                 // K = R*S*Cin
    const bf16* A; long long lda;
    const bf16* W;               // [N][K]
    const float* bias;
    const bf16* res;             // [M][N] or null
    bf16* out;                   // [M][N]
    int relu;                    // 0 none, 1 ReLU, 2 exact GELU
    ConvGeom conv;
    int block_n;                 // 0 = auto
    int cta_group;               // 0 = auto
    float* pool_out; int pool_batch;   // fused 7x7 average pool of an 8x8 final map (out may then be null)
    bf16* scratch; size_t scratch_bytes;   // im2col buffer for geometries neither TMA mode tiles (strided convolutions on odd-sized maps)
};

template <int BN, int CG, int D, int HALO, int GELU = 0>
int convgemm_launch_inst(const CUtensorMap* maps, const CgParams& kp, int grid, cudaStream_t st) {
    using Cfg = CgCfg<BN, CG, D, HALO>;
    static bool configured = false;
    if (!configured) {
        cudaError_t err = cudaFuncSetAttribute(convgemm_kernel<BN, CG, D, HALO, GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (err != cudaSuccess) { set_error("convgemm: cudaFuncSetAttribute: %s", cudaGetErrorString(err)); return -1; }
        configured = true;
    }
    static const int pdl = getenv("SQ_PDL") ? atoi(getenv("SQ_PDL")) : 1;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CG == 2) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na; }
    if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = attr; cfg.numAttrs = na;
    cudaError_t err = cudaLaunchKernelEx(&cfg, convgemm_kernel<BN, CG, D, HALO, GELU>, maps[0], maps[1], maps[2], maps[3], kp);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("convgemm launch: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

int convgemm_dispatch(int bn, int cg, int halo, const CUtensorMap* maps, const CgParams& kp, int grid, cudaStream_t st);   // resnet.cu

// per-tap mode tiles the output map with boxes of 128 pixels: whole rows (Wo divides 128) and whole row groups / images
inline bool conv_pertap_geometry_ok(const ConvGeom& c) {
    if (c.Wo > 128 || 128 % c.Wo != 0) return false;
    int BH = 128 / c.Wo;
    if (BH > c.Ho) { const int BIMG = BH / c.Ho; return BIMG * c.Ho * c.Wo == 128; }
    return c.Ho % BH == 0;
}

struct ConvGemmArgs;
inline int convgemm_im2col_fallback(const ConvGemmArgs& g, cudaStream_t st);

// true when the launch fits this kernel (N a multiple of 64, channels a multiple of 64, tile geometry of the 4-D boxes)
inline bool convgemm_supported(const ConvGemmArgs& g) {
    if (g.N % 64 != 0 || g.K % 64 != 0 || (!g.out && !g.pool_out) || !g.bias) return false;
    if ((reinterpret_cast<uintptr_t>(g.out) | reinterpret_cast<uintptr_t>(g.res) | reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.W)) & 15) return false;
    return true;
}

inline int convgemm_launch(const ConvGemmArgs& g, cudaStream_t st) {
    CgParams kp;
    memset(&kp, 0, sizeof(kp));
    static const int env_cg = getenv("SQ_CONV_CG") ? atoi(getenv("SQ_CONV_CG")) : 0;
    int cg = g.cta_group ? g.cta_group : (env_cg ? env_cg : 2);
    int bn = g.block_n;
    static const int env_bnsel = getenv("SQ_CONV_BNSEL") ? atoi(getenv("SQ_CONV_BNSEL")) : 0;   // wave-aware widths: -23 us per batch alone, but -3 % with two extractor lanes (fuller grids leave the other lane no SMs)
    static const int env_l2pf = getenv("SQ_CONV_L2PF") ? atoi(getenv("SQ_CONV_L2PF")) : 0;      // measured: +13 us per batch when on
    kp.l2pf = env_l2pf;
    if (bn == 0 && !env_bnsel) bn = g.N >= 256 ? 256 : (g.N >= 128 ? 128 : 64);
    if (bn == 0) {
        // widest tile whose waves are not mostly empty: cost = waves x tile width (64-wide MMAs run at ~half rate: x1.5)
        const long long mp = ((g.M + GEMM_BM - 1) / GEMM_BM + cg - 1) / cg, G = num_sms() / cg;
        long long best = -1;
        for (int w = 256; w >= 64; w >>= 1) {
            if (g.N % w != 0) continue;
            const long long tiles = mp * (g.N / w), cost = ((tiles + G - 1) / G) * w * (w == 64 ? 3 : 2);
            if (best < 0 || cost < best) { best = cost; bn = w; }
        }
        if (bn == 0) bn = 64;
    }
    if (cg == 1 && bn == 256) bn = 128;                     // a single CTA has no room for 256-wide stages beside the ring
    if (g.relu == 2) { bn = 256; cg = 2; if (g.N % 256 != 0 || g.conv.enabled) { set_error("convgemm: the GELU epilogue is built for plain GEMMs with N %% 256 == 0"); return -1; } }
    if (g.N % bn != 0) bn = 64;
    kp.M = g.M; kp.N = g.N;
    const int num_m = (g.M + GEMM_BM - 1) / GEMM_BM;
    kp.num_n = g.N / bn;
    kp.total_tiles = ((num_m + cg - 1) / cg) * kp.num_n;      // (halo mode recounts: one tile per 16 x 8 patch)
    kp.bias = g.bias; kp.relu = g.relu; kp.has_res = g.res != nullptr;
    kp.prof = gemm_prof_buffer();
    kp.pool_out = g.pool_out; kp.batch = g.pool_batch;

    CUtensorMap maps[4];
    // halo mode: 3x3 / stride 1 / pad 1 without a residual, output map tiled by 16 x 8 patches
    static const int env_halo = getenv("SQ_CONV_HALO") ? atoi(getenv("SQ_CONV_HALO")) : 1;     // 0 off, 1 on, 2 on with descriptor base offset
    const ConvGeom& cc = g.conv;
    // (measured: 49 -> 41 us on layer 1, 31.6 -> 29.4 us on layer 2; with 256-wide tiles the three weight stages that fit beside
    //  the halo slots starve the MMAs, 25.2 -> 26.6 us on layer 3, so those keep the per-tap boxes)
    // Patches that hang over the map edge are handled by TMA clipping, so ANY output size works in halo mode; maps that do not tile
    // by the per-tap boxes (Wo not a divisor of 128: 224-px inputs, odd sizes) always take it.
    const bool pertap_ok = cc.enabled && conv_pertap_geometry_ok(cc);
    const bool halo_shape = cc.enabled && cc.R == 3 && cc.S == 3 && cc.stride == 1 && cc.pad == 1 && !g.res && cc.C % 64 == 0;
    const bool even = cc.Wo % HALO_TW == 0 && cc.Ho % HALO_TH == 0;
    const int halo = (halo_shape && (!pertap_ok || (env_halo && even && (bn <= 128 || env_halo == 3)))) ? 1 : 0;
    if (cc.enabled && !halo && !pertap_ok) return convgemm_im2col_fallback(g, st);
    if (halo) {
        const ConvGeom& c = g.conv;
        kp.conv = 1; kp.cblocks = c.C / 64; kp.S = 3; kp.stride = 1; kp.pad = 1;
        kp.tiles_x = (c.Wo + HALO_TW - 1) / HALO_TW; kp.tiles_per_img = ((c.Ho + HALO_TH - 1) / HALO_TH) * kp.tiles_x; kp.Ho = c.Ho; kp.halo_bo = env_halo == 2;
        kp.nk = 9 * kp.cblocks;
        if (g.M != c.batch * c.Ho * c.Wo) { set_error("conv: M mismatch"); return -1; }
        // one 128-row tile per 16 x 8 patch (partial patches included)
        kp.total_tiles = ((c.batch * kp.tiles_per_img + cg - 1) / cg) * kp.num_n;
        cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.batch};
        cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
        cuuint32_t box[4] = {64, HALO_PITCH, HALO_TH + 2, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (encode_map(&maps[0], g.A, 4, dims, strides, box, estr)) return -1;
    } else if (g.conv.enabled) {
        const ConvGeom& c = g.conv;
        if (c.C % 64 != 0) { set_error("conv: C=%d must be a multiple of 64", c.C); return -1; }
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
        if (encode_map(&maps[0], g.A, 4, dims, strides, box, estr)) return -1;
    } else {
        kp.nk = g.K / GEMM_BK;
        if (make_operand_map(&maps[0], g.A, 0, g.lda, g.M, g.K, GEMM_BM)) return -1;
    }
    if (make_operand_map(&maps[1], g.W, 0, (long long)kp.nk * 64, g.N, kp.nk * 64, bn / cg)) return -1;
    if (halo) {
        // output sub-tiles: 64 channels x 8 pixels x 4 output rows (= the 32 accumulator rows of one epilogue warp); a 4-D map
        // so that patches hanging over the right / bottom edge of an image are clipped by the TMA unit
        const ConvGeom& c = g.conv;
        cuuint64_t dims[4] = {(cuuint64_t)g.N, (cuuint64_t)c.Wo, (cuuint64_t)c.Ho, (cuuint64_t)c.batch};
        cuuint64_t strides[3] = {(cuuint64_t)g.N * 2, (cuuint64_t)c.Wo * g.N * 2, (cuuint64_t)c.Ho * c.Wo * g.N * 2};
        cuuint32_t box[4] = {64, HALO_TW, 4, 1}, estr[4] = {1, 1, 1, 1};
        if (encode_map(&maps[3], g.out, 4, dims, strides, box, estr)) return -1;
        maps[2] = maps[3];
    } else
    // residual / output sub-tiles: 64 channels x 32 rows
    {
        cuuint64_t dims[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M}, strides[1] = {(cuuint64_t)g.N * 2};
        cuuint32_t box[2] = {64, 32}, estr[2] = {1, 1};
        if (g.res) { if (encode_map(&maps[2], g.res, 2, dims, strides, box, estr)) return -1; }
        if (g.out) { if (encode_map(&maps[3], g.out, 2, dims, strides, box, estr)) return -1; }
        else maps[3] = maps[2];
        if (!g.res) maps[2] = maps[3];
    }
    const int clusters_max = num_sms() / cg;
    const int clusters = kp.total_tiles < clusters_max ? kp.total_tiles : clusters_max;
    gemm_timing_begin(st, 2.0 * g.M * g.N * (double)kp.nk * 64);
    const int rc = convgemm_dispatch(bn, cg, halo, maps, kp, clusters * cg, st);
    gemm_timing_end(st);
    return rc;
}

// Generic fallback: explicit im2col (NHWC -> [M, R*S*C], zero padding) followed by the plain GEMM mode.  Only reached for
// geometries the TMA boxes cannot tile (e.g. the stride-2 convolutions of a 224-px or 256x265 input).
static __global__ void im2col_nhwc_kernel(const bf16* __restrict__ in, bf16* __restrict__ col, int batch, int H, int W, int C, int R, int S, int stride,
                                          int pad, int Ho, int Wo) {
    const int C8 = C / 8;
    const long long total = (long long)batch * Ho * Wo * R * S * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8); long long t = i / C8;
        const int s_ = (int)(t % S); t /= S; const int r = (int)(t % R); t /= R;
        const int ow = (int)(t % Wo); t /= Wo; const int oh = (int)(t % Ho); const int img = (int)(t / Ho);
        const int ih = oh * stride - pad + r, iw = ow * stride - pad + s_;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = *reinterpret_cast<const uint4*>(in + (((long long)img * H + ih) * W + iw) * C + c8 * 8);
        *reinterpret_cast<uint4*>(col + i * 8) = v;
    }
}

inline int convgemm_launch(const ConvGemmArgs& g, cudaStream_t st);
inline int convgemm_im2col_fallback(const ConvGemmArgs& g, cudaStream_t st) {
    const ConvGeom& c = g.conv;
    const size_t need = (size_t)g.M * g.K * sizeof(bf16);
    if (!g.scratch || g.scratch_bytes < need) { set_error("conv: geometry %dx%d -> %dx%d needs an im2col scratch of %zu bytes", c.H, c.W, c.Ho, c.Wo, need); return -1; }
    const long long total = (long long)g.M * (g.K / 8);
    long long blocks = (total + 255) / 256; if (blocks > 148LL * 32) blocks = 148LL * 32;
    im2col_nhwc_kernel<<<(unsigned)blocks, 256, 0, st>>>(g.A, g.scratch, c.batch, c.H, c.W, c.C, c.R, c.S, c.stride, c.pad, c.Ho, c.Wo);
    ConvGemmArgs p = g;
    p.A = g.scratch; p.lda = g.K; p.conv.enabled = 0;
    return convgemm_launch(p, st);
}

}  // namespace sq
