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
#include "aggr.cuh"
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
