// Library-wide plumbing: error string, SM count, driver entry point for tensor-map encoding,
// and the generic GEMM / plane-split entry points of the C ABI (include/sequoia_b200.h).
#include "gemm.cuh"
#include "../../include/sequoia_b200.h"
#include <stdarg.h>
#include <vector>

namespace sq {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static int g_sm_budget = 0;         // 0 = all SMs; set by sq_set_sm_budget while communication kernels need room beside the persistent GEMMs

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return (g_sm_budget > 0 && g_sm_budget < n) ? g_sm_budget : n;
}

PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

// Side stream for work that is independent of the main chain of a forward / backward pass (fork / join with events): its
// kernels soak up the SMs that the persistent GEMM kernels leave idle in partially filled waves.  One side stream + two
// events per device, created lazily (handles only; no device memory).
static int g_side_enabled = -1;      // -1: not decided yet (environment SQ_SIDE_STREAM, default on)
SideStream* side_stream() {
    static SideStream tab[16];
    static bool init[16] = {false};
    if (g_side_enabled < 0) g_side_enabled = getenv("SQ_SIDE_STREAM") ? atoi(getenv("SQ_SIDE_STREAM")) : 1;
    const int enabled = g_side_enabled;
    int dev = 0;
    if (!enabled || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    if (!init[dev]) {
        init[dev] = true;
        tab[dev].ok = cudaStreamCreateWithFlags(&tab[dev].s, cudaStreamNonBlocking) == cudaSuccess &&
                      cudaEventCreateWithFlags(&tab[dev].fork, cudaEventDisableTiming) == cudaSuccess &&
                      cudaEventCreateWithFlags(&tab[dev].join, cudaEventDisableTiming) == cudaSuccess;
    }
    return tab[dev].ok ? &tab[dev] : nullptr;
}
cudaStream_t side_fork(SideStream* ss, cudaStream_t st) {
    if (!ss) return st;
    cudaEventRecord(ss->fork, st);
    cudaStreamWaitEvent(ss->s, ss->fork, 0);
    return ss->s;
}
void side_join(SideStream* ss, cudaStream_t st) {
    if (!ss) return;
    cudaEventRecord(ss->join, ss->s);
    cudaStreamWaitEvent(st, ss->join, 0);
}

// One instantiation of the GEMM kernel per (tile width, epilogue class[, fused split-precision staging]).
int launch_gemm_dispatch(int bn, int cls, int fuse3, const CUtensorMap* maps, const GemmKParams& kp, int grid, cudaStream_t st) {
#define SQ_CASE(BN, CLS) if (!fuse3 && bn == BN && cls == CLS) return launch_gemm_inst<BN, CLS, 0>(maps, kp, grid, st);
#define SQ_CASE3(BN, CLS) if (fuse3 && bn == BN && cls == CLS) return launch_gemm_inst<BN, CLS, 1>(maps, kp, grid, st);
#define SQ_ALL_BN(CLS) SQ_CASE(64, CLS) SQ_CASE(128, CLS) SQ_CASE(256, CLS)
#define SQ_FUSED(CLS) SQ_CASE3(128, CLS) SQ_CASE3(192, CLS) SQ_CASE3(256, CLS)
    SQ_FUSED(EPI_F32) SQ_FUSED(EPI_GELU) SQ_FUSED(EPI_DGELU) SQ_FUSED(EPI_LN64)
    SQ_ALL_BN(EPI_CONV) SQ_ALL_BN(EPI_CONV_PF) SQ_ALL_BN(EPI_F32) SQ_ALL_BN(EPI_GELU) SQ_ALL_BN(EPI_DGELU) SQ_ALL_BN(EPI_LN64) SQ_ALL_BN(EPI_GENERIC)
#undef SQ_FUSED
#undef SQ_ALL_BN
#undef SQ_CASE3
#undef SQ_CASE
    set_error("gemm: no kernel for block_n %d class %d fuse3 %d", bn, cls, fuse3);
    return -1;
}

// ---- per-launch timing of the tensor-core kernel (used by bench.py to report the roofline of the dominant kernel)
static bool g_timing = false;
static std::vector<cudaEvent_t> g_ev;       // pairs: begin, end
static size_t g_ev_used = 0;
static double g_flops = 0.0;

static bool g_chain = false;          // inside a bracketed chain of launches: per-launch events are suppressed, work is still counted
static long long g_chain_launches = 0, g_extra_launches = 0;

// One event pair around a CHAIN of back-to-back launches (the 52 convolutions of a ResNet batch): the launches keep overlapping
// through programmatic dependent launch exactly as in an untimed run, and duration / launches is the average launch duration.
void gemm_timing_chain_begin(cudaStream_t st) {
    if (!g_timing || g_chain) return;
    if (g_ev_used + 2 > g_ev.size()) {
        for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); g_ev.push_back(e); }
    }
    cudaEventRecord(g_ev[g_ev_used], st);
    g_chain = true; g_chain_launches = 0;
}
void gemm_timing_chain_end(cudaStream_t st) {
    if (!g_timing || !g_chain) return;
    cudaEventRecord(g_ev[g_ev_used + 1], st);
    g_ev_used += 2;
    g_extra_launches += g_chain_launches - 1;      // the event pair counts as one launch in sq_gemm_timing_read
    g_chain = false;
}

void gemm_timing_begin(cudaStream_t st, double flops) {
    if (!g_timing) return;
    if (g_chain) { g_flops += flops; ++g_chain_launches; return; }
    if (g_ev_used + 2 > g_ev.size()) {
        for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); g_ev.push_back(e); }
    }
    g_flops += flops;
    cudaEventRecord(g_ev[g_ev_used], st);
}
void gemm_timing_end(cudaStream_t st) {
    if (!g_timing || g_chain) return;
    cudaEventRecord(g_ev[g_ev_used + 1], st);
    g_ev_used += 2;
}

static unsigned long long* g_prof = nullptr;
unsigned long long* gemm_prof_buffer() { return g_prof; }

// fp32 [rows, cols] (ld_in) -> bf16 hi / lo planes (ld_out); lo may be null
__global__ void split_planes_kernel(const float* __restrict__ x, bf16* __restrict__ hi, bf16* __restrict__ lo, long long rows,
                                    int cols, long long ld_in, long long ld_out) {
    const long long n = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols; const int c = (int)(i - r * cols);
        const float v = x[r * ld_in + c];
        const bf16 h = __float2bfloat16_rn(v);
        hi[r * ld_out + c] = h;
        if (lo) lo[r * ld_out + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

int split_planes(const float* x, bf16* hi, bf16* lo, long long rows, int cols, long long ld_in, long long ld_out, cudaStream_t st) {
    const long long n = rows * cols;
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    split_planes_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, hi, lo, rows, cols, ld_in, ld_out);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("split_planes: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // namespace sq

using namespace sq;

extern "C" {

int sq_version(void) { return 100; }

const char* sq_last_error(void) { return g_err; }

int sq_device_ok(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) { set_error("no CUDA device"); return -1; }
    if (major != 10) { set_error("sequoia_b200 needs an sm_100 device, found sm_%d%d", major, minor); return -2; }
    return 0;
}

int sq_gemm_timing_enable(int on) {
    g_timing = on != 0; g_ev_used = 0; g_flops = 0.0; g_chain = false; g_extra_launches = 0;
    return 0;
}

int sq_gemm_timing_read(double* total_ms, long long* launches, double* mma_flops) {
    double ms = 0.0;
    for (size_t i = 0; i + 1 < g_ev_used; i += 2) {
        float t = 0.f;
        cudaError_t err = cudaEventElapsedTime(&t, g_ev[i], g_ev[i + 1]);
        if (err != cudaSuccess) { set_error("gemm timing: %s (synchronise the stream first)", cudaGetErrorString(err)); return -1; }
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (long long)(g_ev_used / 2) + g_extra_launches;
    if (mma_flops) *mma_flops = g_flops;
    g_ev_used = 0; g_flops = 0.0; g_extra_launches = 0;
    return 0;
}

int sq_side_stream_enable(int on) { g_side_enabled = on ? 1 : 0; return 0; }

int sq_set_sm_budget(int sms) { g_sm_budget = sms > 0 ? sms : 0; return 0; }

int sq_gemm_profile(void* device_buffer) { g_prof = (unsigned long long*)device_buffer; return 0; }

int sq_split_bf16(const float* x, void* hi, void* lo, long long rows, int cols, long long ld_in, long long ld_out, void* stream) {
    return split_planes(x, (bf16*)hi, (bf16*)lo, rows, cols, ld_in, ld_out, (cudaStream_t)stream);
}

int sq_gemm_bf16(const sq_gemm_desc* d, void* stream) {
    if (!d) { set_error("null desc"); return -1; }
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.M = d->M; g.N = d->N; g.K = d->K;
    g.A.hi = (const bf16*)d->a_hi; g.A.lo = (const bf16*)d->a_lo; g.A.mn_major = d->a_mn_major; g.A.ld = d->lda;
    g.B.hi = (const bf16*)d->b_hi; g.B.lo = (const bf16*)d->b_lo; g.B.mn_major = d->b_mn_major; g.B.ld = d->ldb;
    g.nterms = d->nterms; g.split_k = d->split_k; g.block_n = d->block_n; g.a_koff_per_ntile = d->a_koff_per_ntile;
    g.workspace = (float*)d->workspace; g.workspace_bytes = d->workspace_bytes;
    if (d->conv_enabled) {
        g.conv.enabled = 1; g.conv.batch = d->conv_batch; g.conv.H = d->conv_H; g.conv.W = d->conv_W; g.conv.C = d->conv_C;
        g.conv.Ho = d->conv_Ho; g.conv.Wo = d->conv_Wo; g.conv.R = d->conv_R; g.conv.S = d->conv_S; g.conv.stride = d->conv_stride; g.conv.pad = d->conv_pad;
    }
    g.e.out_f32 = d->out_f32; g.e.ld_f32 = d->ld_f32;
    g.e.out_hi = (bf16*)d->out_hi; g.e.out_lo = (bf16*)d->out_lo; g.e.ld_bf = d->ld_bf;
    g.e.bias = d->bias;
    g.e.rowbias = d->rowbias; g.e.rowbias_div = d->rowbias_div > 0 ? d->rowbias_div : 1; g.e.ld_rowbias = d->ld_rowbias;
    g.e.res_f32 = d->res_f32; g.e.res_bf = (const bf16*)d->res_bf; g.e.ld_res = d->ld_res;
    g.e.save_pre = d->save_pre; g.e.ld_pre = d->ld_pre;
    g.e.aux = d->aux; g.e.ld_aux = d->ld_aux;
    g.e.ln_gamma = d->ln_gamma; g.e.ln_beta = d->ln_beta;
    g.e.act = d->act; g.e.alpha = d->alpha;
    g.b_koff_per_ntile = d->b_koff_per_ntile; g.b_nadj_per_ntile = d->b_nadj_per_ntile;
    g.b_map_mn = d->b_map_mn; g.b_map_k = d->b_map_k; g.diag64 = d->diag64;
    return gemm_launch(g, (cudaStream_t)stream);
}

}  // extern "C"
