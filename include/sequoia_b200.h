/* sequoia_b200 — C ABI of the B200-native SEQUOIA hot paths.
 *
 * The reference (gevaertlab/sequoia-pub) has no native code and no FFI: its "plugin API" for these
 * paths is the Python nn.Module surface.  This header is the boundary a replacement
 * sits behind: plain C types, raw device pointers, explicit sizes, a CUDA stream handle.  Every entry
 * point cites the reference interface it replaces.
 *
 * Conventions
 *   - return 0 on success, negative on error; sq_last_error() returns a thread-local message;
 *   - no C++ exception crosses the boundary, nothing here allocates memory the caller cannot see:
 *     workspaces are sized by the *_workspace_bytes functions and passed in;
 *   - every function ENQUEUES on `stream` (a cudaStream_t passed as void*) and never synchronises;
 *   - all pointers are device pointers unless the name says `host`;
 *   - "planes": an fp32 matrix split as hi = bf16(x), lo = bf16(x - hi); the tensor-core GEMMs use
 *     hi*hi + hi*lo + lo*hi to keep fp32 parity.
 */
#ifndef SEQUOIA_B200_H
#define SEQUOIA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SQ_API __attribute__((visibility("default")))
#else
#define SQ_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ library */
SQ_API int sq_version(void);
SQ_API const char* sq_last_error(void);
/* 0 when the current device is an sm_100 part; the product path refuses to run otherwise. */
SQ_API int sq_device_ok(void);

/* Measurement aid for bench.py: when enabled, launches of the tcgen05 GEMM/conv kernels are bracketed by CUDA events on
 * their stream - one pair per launch, except inside sq_resnet50_extract, where ONE pair brackets the chain of 52 back-to-back
 * convolution launches of a batch (so they overlap through programmatic dependent launch as in an untimed run; the chain counts
 * as 52 launches).  _read (after a stream synchronise) returns the summed time, the launch count and the tensor-core FLOPs those
 * launches issued, then resets the counters. */
SQ_API int sq_gemm_timing_enable(int on);
SQ_API int sq_gemm_timing_read(double* total_ms, long long* launches, double* mma_flops);

/* Independent sub-chains of a pass (ViS summary branch, weight gradients) are enqueued on a library-owned side stream so
 * that they fill SMs the main chain leaves idle (default on; environment SQ_SIDE_STREAM=0 disables).  Measurement code
 * turns it off to time kernels without overlap. */
SQ_API int sq_side_stream_enable(int on);

/* Development aid: when `device_buffer` (>= 148*16 uint64) is non-NULL every following GEMM launch stores per-CTA cycle
 * counters of its warp roles: [cta][0..1] producer {waiting for a free stage, total}; [2..4] MMA issuer {waiting for
 * data, waiting for a drained accumulator, total}; [5+2q, 6+2q] epilogue warp q {waiting for an accumulator, total}. */
SQ_API int sq_gemm_profile(void* device_buffer);

/* ------------------------------------------------------------------ building blocks (exposed for tests) */

/* fp32 [rows, cols] (row stride ld_in) -> bf16 hi / lo planes (row stride ld_out); lo may be NULL. */
SQ_API int sq_split_bf16(const float* x, void* hi, void* lo, long long rows, int cols, long long ld_in, long long ld_out,
                  void* stream);

enum { SQ_ACT_NONE = 0, SQ_ACT_RELU = 1, SQ_ACT_GELU = 2, SQ_ACT_LN64_GELU = 3, SQ_ACT_MUL_DGELU = 4 };

/* One tcgen05 GEMM / implicit-GEMM convolution launch with its fused epilogue.
 *   C[M,N] = alpha * sum_terms A[M,K] * B[N,K]^T, then bias / rowbias / residual / activation.
 * Replaces every torch.nn.Linear / nn.Conv2d library call on the hot path
 * (src/tformer_lin.py:14-16,37,55-58,93; src/resnet.py:60-68,101,124-128). */
typedef struct sq_gemm_desc {
    int M, N, K;
    const void* a_hi; const void* a_lo; int a_mn_major; long long lda;
    const void* b_hi; const void* b_lo; int b_mn_major; long long ldb;
    int nterms;            /* 1 = plain bf16, 3 = split precision */
    int split_k;           /* >1: partials go to workspace, a second kernel reduces + applies the epilogue */
    int block_n;           /* 0 = auto, else 64 / 128 / 256 */
    int a_koff_per_ntile;  /* block-diagonal mode (per-head 64x64 mixing weights) */
    void* workspace; size_t workspace_bytes;
    /* conv mode: A is an NHWC bf16 tensor, B is [Cout][R][S][Cin] */
    int conv_enabled, conv_batch, conv_H, conv_W, conv_C, conv_Ho, conv_Wo, conv_R, conv_S, conv_stride, conv_pad;
    /* epilogue */
    float* out_f32; long long ld_f32;
    void* out_hi; void* out_lo; long long ld_bf;
    const float* bias;
    const float* rowbias; int rowbias_div; long long ld_rowbias;
    const float* res_f32; const void* res_bf; long long ld_res;
    float* save_pre; long long ld_pre;
    const float* aux; long long ld_aux;
    const float* ln_gamma; const float* ln_beta;
    int act;
    float alpha;           /* 0 is read as 1 */
    /* block-diagonal backward modes (per-head 64x64 weights of SummaryMixing.c, src/tformer_lin.py:16,24) */
    int b_koff_per_ntile;  /* dgrad: B k-offset per n-tile */
    int b_nadj_per_ntile;  /* dgrad: B n-coordinate adjustment per n-tile */
    int b_map_mn, b_map_k; /* explicit extents of B's tensor map (0 = N, K) */
    int diag64;            /* wgrad: only the diagonal 64x64 blocks of C[M,M] are produced, stored at columns 0..63 */
} sq_gemm_desc;

SQ_API int sq_gemm_bf16(const sq_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------ ResNet-50 feature extractor
 * Replaces ResNet.forward_extract (src/resnet.py:155-170) and the CPU preprocessing in front of it
 * (pre_processing/compute_features_hdf5.py:49-51,117-123). */

/* Topology: 53 convolutions in module order: conv1, then per Bottleneck conv1, conv2, conv3[, downsample.0]
 * (src/resnet.py:60-68,101-109,123-128). */
SQ_API int sq_resnet50_num_convs(void);
SQ_API int sq_resnet50_conv_info(int idx, int* cin, int* cout, int* k, int* stride, int* pad);
SQ_API long long sq_resnet50_packed_weight_elems(void); /* bf16 elements */
SQ_API long long sq_resnet50_shift_elems(void);         /* fp32 elements */

/* Folds eval-mode BatchNorm (src/resnet.py:103, compute_features_hdf5.py:60) into the conv weights.
 * `tensors` is a HOST array of 53*5 DEVICE pointers: {conv.weight (OIHW fp32), bn.weight, bn.bias,
 * bn.running_mean, bn.running_var} per convolution.  Must be re-run whenever parameters change. */
SQ_API int sq_resnet50_prepack(const void* const* tensors, void* packed_w, float* shifts, float bn_eps, void* stream);

SQ_API size_t sq_resnet50_workspace_bytes(int batch, int H, int W);

/* input_kind 0: uint8 [batch,H,W,3] raw RGB tiles; x/255 and Normalize(mean,std) are fused (R4).
 * input_kind 1: fp32 [batch,3,H,W] already normalised — the tensor the reference passes to forward_extract.
 * features: fp32 [batch, 2048] = AvgPool2d(7) of layer4, i.e. the top-left 7x7 of the 8x8 map at 256 px. */
SQ_API int sq_resnet50_extract(const void* input, int input_kind, int batch, int H, int W, const void* packed_w,
                               const float* shifts, float* features, void* workspace, size_t workspace_bytes,
                               void* stream);

/* Opt-in split-precision ("bf16x3") extraction: same interface, every convolution computed as hi*hi + hi*lo + lo*hi with fp32
 * residuals (~1e-5 from the fp64 reference instead of 1.4e-3 for plain bf16 operands, at roughly a fifth of the throughput).
 * Exists to QUANTIFY what the bf16 operands of sq_resnet50_extract cost downstream; 256x256 patches only.
 * sq_resnet50_prepack_planes also writes the lo plane of the folded weights (same element count as packed_w). */
SQ_API int sq_resnet50_prepack_planes(const void* const* tensors, void* packed_w, void* packed_w_lo, float* shifts, float bn_eps,
                                      void* stream);
SQ_API size_t sq_resnet50_hp_workspace_bytes(int batch, int H, int W);
SQ_API int sq_resnet50_extract_hp(const void* input, int input_kind, int batch, int H, int W, const void* packed_w,
                                  const void* packed_w_lo, const float* shifts, float* features, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* One bottleneck convolution on the CTA-pair tcgen05 kernel (csrc/convgemm.cuh):
 *   out = act(conv(in, weight) + shift [+ residual]),  NHWC bf16 in / out, weight [Cout][R][S][Cin] bf16 (BatchNorm scale folded),
 *   shift fp32 [Cout], residual bf16 [batch*Ho*Wo, Cout] or NULL.  Cin and Cout multiples of 64.
 * Building block of sq_resnet50_extract, exposed for tests (src/resnet.py:73-93: conv -> bn -> (+= residual) -> relu).
 * block_n: 0 = auto, 64 / 128 / 256; cta_group: 0 = auto (2), 1, 2. */
typedef struct sq_conv_desc {
    int batch, H, W, Cin, Cout, R, S, stride, pad;
    const void* in; const void* weight; const float* shift; const void* residual; void* out;
    int relu, block_n, cta_group;
} sq_conv_desc;
SQ_API int sq_conv_bf16(const sq_conv_desc* desc, void* stream);

/* Tail of a layer-1 bottleneck (src/resnet.py:73-93) as ONE kernel (csrc/fusedconv.cuh):
 *   out = relu(conv1x1(relu(conv3x3(in, w2) + shift2), w3) + shift3 + residual)
 * in: NHWC bf16 [batch, H, W, 64] (conv1's output); w2: bf16 [64][3][3][64]; w3: bf16 [256][64]; shifts fp32; residual / out: bf16
 * [batch, H, W, 256].  H a multiple of 16, W of 8.  The 64-channel intermediate is rounded to bf16 and kept in tensor memory; results
 * are bit-identical to sq_conv_bf16 called twice.  Building block of sq_resnet50_extract, exposed for tests. */
SQ_API int sq_bneck_l1_bf16(const void* in, const void* w2, const float* shift2, const void* w3, const float* shift3, const void* residual,
                            void* out, int batch, int H, int W, void* stream);
/* The same tail for the FIRST block of the layer: the downsample branch is computed inside the kernel,
 *   out = relu(conv1x1(relu(conv3x3(in, w2) + shift2), w3) + shift3 + conv1x1(x, wds) + shiftds)
 * x: NHWC bf16 [batch, H, W, 64] (the block input), wds: bf16 [256][64].  The residual sum is formed in fp32 and rounded once. */
SQ_API int sq_bneck_l1_ds_bf16(const void* in, const void* w2, const float* shift2, const void* w3, const float* shift3, const void* x,
                               const void* wds, const float* shiftds, void* out, int batch, int H, int W, void* stream);

/* ------------------------------------------------------------------ ViS aggregator (SummaryMixing transformer)
 * Replaces ViS.forward (src/tformer_lin.py:97-106 and everything it calls, :18-26,39-48,60-61,73-77), the autograd
 * backward behind loss.backward() (src/vit.py:179), nn.MSELoss (src/vit.py:129,166) and the AdamW step
 * (src/main.py:180-183, src/vit.py:180).  dimensions_f = dimensions_s = dimensions_c = 64 (hard-coded at
 * src/main.py:147,167). */
typedef struct sq_vis_config {
    int input_dim;     /* D: 2048 (ResNet features) or 1024 (UNI); multiple of 64 */
    int depth;         /* L */
    int nheads;        /* H */
    int num_clusters;  /* N tokens per slide (100) */
    int num_outputs;   /* G genes */
} sq_vis_config;

/* Flat fp32 parameter buffer.  The table lists element offsets of, in order: pos_emb1D [N,D]; per layer 18 entries
 * {local_norm.weight [H*64], local_norm.bias, summary_norm.weight, summary_norm.bias, s.weight [H*64,D], s.bias,
 *  f.weight [H*64,D], f.bias, c.weight [H*64,128], c.bias, projection.weight [D,H*64], projection.bias,
 *  net.0.weight [D], net.0.bias, net.1.weight [D,D], net.1.bias, net.3.weight [D,D], net.3.bias}
 * (the per-mixer tensors of head h are rows h*64..h*64+63 of these); linear_head.0.weight, .0.bias, .1.weight [G,D], .1.bias.
 * Gradients, Adam moments and the bf16 hi/lo planes of the parameters use the same offsets. */
SQ_API int sq_vis_param_table_len(const sq_vis_config* cfg);
SQ_API int sq_vis_param_layout(const sq_vis_config* cfg, long long* offsets, int n, long long* total_elems);
SQ_API size_t sq_vis_act_bytes(const sq_vis_config* cfg, int batch);   /* activations kept for the backward pass */
SQ_API size_t sq_vis_bwd_bytes(const sq_vis_config* cfg, int batch);   /* scratch of the backward pass */

/* x: fp32 [batch, N, D] -> pred fp32 [batch, G].  w_hi / w_lo: bf16 planes of `params` (sq_split_bf16 or sq_adamw_flat). */
SQ_API int sq_vis_forward(const sq_vis_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* x,
                          int batch, float* pred, void* act, size_t act_bytes, void* stream);
/* Backward stages stage_hi .. stage_lo (inclusive, descending): stage `depth` = regression head (needs dpred fp32
 * [batch,G]), stage l < depth = transformer layer l (stage 0 also produces pos_emb1D's gradient and, if dx != NULL,
 * dL/dx [batch,N,D]).  Every gradient of a stage is OVERWRITTEN in `grads` (flat, same offsets as params); stages are
 * contiguous ranges of that buffer so a data-parallel caller can all-reduce a stage while the next one runs. */
SQ_API int sq_vis_backward(const sq_vis_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* dpred,
                           int batch, void* act, size_t act_bytes, float* grads, float* dx, void* scratch, size_t scratch_bytes,
                           int stage_hi, int stage_lo, void* stream);
/* loss = mean((pred - target)^2) over batch*G elements; dpred = 2 (pred - target) / (batch*G) (may be NULL).
 * scratch: >= 512 floats. */
SQ_API int sq_mse_fwd_bwd(const float* pred, const float* target, int batch, int num_outputs, float* loss, float* dpred,
                          float* scratch, void* stream);
/* torch.optim.AdamW(amsgrad=False) update of a flat buffer (n multiple of 4); gradients are multiplied by grad_scale
 * first (1/world_size after a sum all-reduce).  p_hi / p_lo (may be NULL) receive the refreshed bf16 planes. */
SQ_API int sq_adamw_flat(float* p, const float* g, float* m, float* v, void* p_hi, void* p_lo, long long n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ------------------------------------------------------------------ ViT softmax-attention aggregator
 * Replaces ViT.forward (src/vit.py:107-116 and everything it calls: Attention :62-74, FeedForward :47-48, Transformer
 * :87-91) and its autograd backward for the reference's `--model_type vit` baseline (src/main.py:141-143,161-163:
 * dim_head = 64, mlp_dim = 2048, to_qkv / to_out without bias).  Same stage contract as sq_vis_*: sq_mse_fwd_bwd and
 * sq_adamw_flat serve both models. */
typedef struct sq_vit_config {
    int dim;           /* D: feature dimension (2048 / 1024); multiple of 64 */
    int depth;         /* L */
    int heads;         /* H, dim_head = 64 */
    int num_clusters;  /* N tokens per slide, <= 128 */
    int num_outputs;   /* G genes */
    int mlp_dim;       /* hidden width of the feed-forward block; multiple of 64 */
} sq_vit_config;

/* Flat fp32 parameter buffer, element offsets in order: pos_emb1D [N,D]; per layer 10 entries {0.norm.weight [D],
 * 0.norm.bias, 0.to_qkv.weight [3*H*64, D], 0.to_out.weight [D, H*64], 1.net.0.weight [D], 1.net.0.bias,
 * 1.net.1.weight [mlp, D], 1.net.1.bias, 1.net.3.weight [D, mlp], 1.net.3.bias}; linear_head.0.weight, .0.bias,
 * .1.weight [G,D], .1.bias. */
SQ_API int sq_vit_param_table_len(const sq_vit_config* cfg);
SQ_API int sq_vit_param_layout(const sq_vit_config* cfg, long long* offsets, int n, long long* total_elems);
SQ_API size_t sq_vit_act_bytes(const sq_vit_config* cfg, int batch);
SQ_API size_t sq_vit_bwd_bytes(const sq_vit_config* cfg, int batch);
SQ_API int sq_vit_forward(const sq_vit_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* x,
                          int batch, float* pred, void* act, size_t act_bytes, void* stream);
/* Stages as in sq_vis_backward: `depth` = regression head, l < depth = transformer layer l (stage 0 also pos_emb1D, dx). */
SQ_API int sq_vit_backward(const sq_vit_config* cfg, const float* params, const void* w_hi, const void* w_lo, const float* dpred,
                           int batch, void* act, size_t act_bytes, float* grads, float* dx, void* scratch, size_t scratch_bytes,
                           int stage_hi, int stage_lo, void* stream);

/* Persistent GEMM kernels launch one CTA per SM and fill the register file, so nothing can co-reside with them.  While a
 * communication kernel has to run beside them (data-parallel backward pass) the caller limits the GEMM grids to `sms` SMs;
 * 0 restores the full device. */
SQ_API int sq_set_sm_budget(int sms);

/* ------------------------------------------------------------------ data-parallel gradient exchange
 * In-place sum all-reduce of grads[begin, begin + count) over NVSwitch multicast (NVLS): `multicast_base` is the MULTICAST address
 * of the symmetric flat gradient buffer (torch.distributed._symmetric_memory: hdl.multicast_ptr), `peer_flags` a HOST array of
 * `world` device pointers to each rank's copy of a zero-initialised symmetric flag buffer of sq_multimem_flag_bytes(ctas) bytes
 * (hdl.buffer_ptrs of that buffer).  `epoch` must grow by 2 per call on a flag buffer (the kernel uses epoch and epoch + 1).
 * Every rank must enqueue the same sequence of calls.  begin and count in elements, multiples of 4.
 * Replaces the one `loss.backward()` + DDP-style exchange the reference would need for multi-GPU training (it trains on one GPU,
 * src/main.py:177). */
SQ_API size_t sq_multimem_flag_bytes(int max_ctas);
SQ_API int sq_multimem_allreduce_f32(void* multicast_base, long long begin, long long count, const void* const* peer_flags, int rank,
                                     int world, unsigned int epoch, int ctas, void* stream);

/* ------------------------------------------------------------------ per-step training metrics
 * Replaces sklearn mean_absolute_error + he2rna.compute_correlations of the training loop (src/vit.py:167-168,
 * src/he2rna.py:140-149) and evaluate()'s smape (src/vit.py:32-33,269).  labels, preds: fp32 [batch, num_outputs] (device).
 * out4 (device): {mean absolute error, mean over genes of the Pearson correlation (genes with constant labels skipped, NaN
 * correlations dropped), number of genes that entered the mean, SMAPE = 100 / batch * sum(2|F-A| / (|A|+|F|))}. */
SQ_API size_t sq_step_metrics_scratch_bytes(int num_outputs);
SQ_API int sq_step_metrics(const float* labels, const float* preds, int batch, int num_outputs, float* out4, void* scratch,
                           size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------ UNI ViT-L/16 feature extractor
 * Replaces the timm model of pre_processing/compute_features_hdf5.py:63-66 and its call `model(image[None,:])` (:128) plus
 * the ToTensor/Normalize preprocessing (:53-56).  `tensors`: HOST array of sq_vitl16_num_tensors(depth) DEVICE pointers in
 * timm state_dict order: cls_token, pos_embed, patch_embed.proj.{weight,bias}, per block {norm1.weight, norm1.bias,
 * attn.qkv.weight, attn.qkv.bias, attn.proj.weight, attn.proj.bias, ls1.gamma, norm2.weight, norm2.bias, mlp.fc1.weight,
 * mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias, ls2.gamma}, norm.weight, norm.bias (all fp32).
 * input_kind 0: uint8 [batch,224,224,3]; 1: fp32 [batch,3,224,224] normalised.  features: fp32 [batch,1024]. */
SQ_API int sq_vitl16_num_tensors(int depth);
SQ_API long long sq_vitl16_packed_weight_elems(int depth);   /* bf16 elements */
SQ_API long long sq_vitl16_packed_vec_elems(int depth);      /* fp32 elements */
SQ_API int sq_vitl16_prepack(const void* const* tensors, int depth, void* packed_w, float* packed_v, void* stream);
SQ_API size_t sq_vitl16_workspace_bytes(int batch);
SQ_API int sq_vitl16_extract(const void* input, int input_kind, int batch, int depth, const void* packed_w, const float* packed_v,
                             float* features, void* workspace, size_t workspace_bytes, void* stream);

/* Image resize in front of UNI: `transforms.Resize(224)` of a PIL image (pre_processing/compute_features_hdf5.py:53-56,125-126)
 * = Pillow's antialiased bilinear resample in 22-bit fixed point (libImaging/Resample.c), bit-identical on the GPU.
 * sq_resize_ksize / sq_resize_coeffs are HOST helpers (no GPU): window bounds int32 [out, 2] = (first input index, count) and
 * weights int32 [out, ksize] for one axis.  sq_resize_bilinear_u8: uint8 [n, Hin, Win, 3] -> uint8 [n, Hout, Wout, 3]; the tables
 * are DEVICE copies of the helpers' output (NULL for an axis that keeps its size); tmp: n*Hin*Wout*3 bytes when both axes change. */
SQ_API int sq_resize_ksize(int in_size, int out_size);
SQ_API int sq_resize_coeffs(int in_size, int out_size, int* bounds_host, int* coeffs_host);
SQ_API int sq_resize_bilinear_u8(const void* in, int n, int Hin, int Win, void* out, int Hout, int Wout, const int* xbounds,
                                 const int* xcoeffs, int xksize, const int* ybounds, const int* ycoeffs, int yksize, void* tmp,
                                 void* stream);

/* ------------------------------------------------------------------ per-slide k-means reduction
 * Replaces `KMeans(n_clusters=100, random_state=0).fit(features)` and the per-label mean loop of
 * pre_processing/kmean_features.py:96-105 (the arithmetic is scikit-learn's: init='k-means++', n_init=1,
 * max_iter=300, tol=1e-4, algorithm='lloyd').  features: fp32 [n, d] (d % 4 == 0), device.
 * The MT19937 draws are data independent and come from the caller:
 *   first_center = RandomState(seed).choice(n, p=uniform);  uniforms = the next (k-1)*trials .uniform() doubles (device),
 *   trials = 2 + int(ln k).
 * init_rows (may be NULL): int32 [k] device, explicit initial centre rows instead of k-means++ (sklearn's `init=X[rows]`;
 *   duplicates are allowed and exercise the empty-cluster relocation); uniforms may then be NULL.
 * Outputs (device): labels int32 [n]; cluster_means fp32 [k, d] = mean of the RAW rows of every label, ascending row
 * order (a row of NaN for an empty label, like np.mean); chosen (may be NULL) int32 [k] = initial centre rows;
 * n_iter_dev (may be NULL) int32 [1] = sklearn's n_iter_.
 * Empty clusters are relocated like sklearn's _relocate_empty_clusters_dense (farthest samples in descending distance).
 * Like every other entry point this one only ENQUEUES: the Lloyd loop is a CUDA graph with a device-controlled WHILE node
 * (one graph per workspace and shape, cached); with SQ_KMEANS_GRAPH=0, or a driver without conditional graph nodes, it falls back
 * to a host loop that synchronises the stream once per iteration.
 * Limits: d % 4 == 0; 2 <= trials <= 10 (k < 2981); n <= ~40 000 samples (single-block selection kernels; the reference caps a
 * slide at max_patch_number = 4000 tiles, pre_processing/compute_features_hdf5.py:26). */
SQ_API size_t sq_kmeans_workspace_bytes(int n, int d, int k);
SQ_API int sq_kmeans_fit(const float* features, int n, int d, int k, int trials, int first_center, const double* uniforms,
                         const int* init_rows, int max_iter, float tol_scale, int* labels, float* cluster_means, int* chosen,
                         int* n_iter_dev, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEQUOIA_B200_H */
