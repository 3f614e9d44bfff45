// Gradient exchange of the data-parallel ViS / ViT train step over NVLink 5 / NVSwitch (one sum all-reduce of the
// flat fp32 gradient per step, issued per backward stage).
//
// sq_multimem_allreduce_f32 is an in-place all-reduce through the NVSwitch's multicast objects (NVLS): the gradient buffer is
// symmetric memory with a multicast mapping (allocated by the caller with torch.distributed._symmetric_memory, plumbing only);
// every rank owns 1/world of the slice, pulls the SUM of that part from all ranks with ONE `multimem.ld_reduce` per 16 bytes (the
// switch adds the eight copies), and pushes the result to all ranks with ONE `multimem.st` - 1x the slice out and 1x in per GPU
// instead of the 2 x 7/8 of a ring, from a handful of CTAs, so the persistent GEMMs of the backward pass keep (almost) all SMs.
// Cross-GPU ordering: an epoch-valued flag per (CTA, source rank) in a small symmetric flag buffer, written to the peers with
// st.release.sys before the data phase and again after it.
#include "gemm.cuh"
#include "../../include/sequoia_b200.h"

namespace sq {

constexpr int MM_MAX_WORLD = 16;
constexpr int MM_UNROLL = 8;

struct MmPeers { unsigned int* flags[MM_MAX_WORLD]; };

__device__ __forceinline__ void mm_barrier(const MmPeers& peers, int rank, int world, unsigned int epoch) {
    // every thread of the CTA has finished its part before the CTA signals (and all of them wait for the peers afterwards)
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        unsigned int* dst = peers.flags[peer] + (size_t)blockIdx.x * MM_MAX_WORLD + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
        const unsigned int* src = peers.flags[rank] + (size_t)blockIdx.x * MM_MAX_WORLD + peer;
        unsigned int v = 0;
        const long long t0 = clock64();
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
            if (clock64() - t0 > 4000000000LL) { printf("sequoia_b200: multimem all-reduce barrier timeout (rank %d waits for %d, epoch %u, saw %u)\n", rank, peer, epoch, v); __trap(); }
        } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) multimem_allreduce_kernel(float* __restrict__ mc, long long begin, long long count, MmPeers peers, int rank, int world,
                                                                  unsigned int epoch) {
    mm_barrier(peers, rank, world, epoch);                       // every rank's gradients of this slice are complete
    // my part of the slice, in 16-byte units
    const long long n4 = (count + 3) / 4, per = (n4 + world - 1) / world;
    const long long lo = per * rank, hi = (lo + per < n4) ? lo + per : n4;
    float* base = mc + begin;
    // MM_UNROLL independent 16-byte reductions in flight per thread (a multimem.ld_reduce is a round trip through the switch)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * MM_UNROLL) {
        float4 v[MM_UNROLL];
#pragma unroll
        for (int u = 0; u < MM_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(base + 4 * i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < MM_UNROLL; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi)
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + 4 * i), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
        }
    }
    __threadfence_system();                                      // my multimem stores are performed at every peer before I signal
    mm_barrier(peers, rank, world, epoch + 1);                   // every rank's part has landed here
}

}  // namespace sq

using namespace sq;

extern "C" {

size_t sq_multimem_flag_bytes(int max_ctas) { return (size_t)max_ctas * MM_MAX_WORLD * sizeof(unsigned int); }

int sq_multimem_allreduce_f32(void* multicast_base, long long begin, long long count, const void* const* peer_flags, int rank, int world,
                              unsigned int epoch, int ctas, void* stream) {
    if (!multicast_base || !peer_flags) { set_error("multimem_allreduce: null pointer"); return -1; }
    if (world < 1 || world > MM_MAX_WORLD || rank < 0 || rank >= world) { set_error("multimem_allreduce: bad rank / world"); return -1; }
    if (count <= 0) return 0;
    if ((begin & 3) || (reinterpret_cast<uintptr_t>(multicast_base) & 15)) { set_error("multimem_allreduce: slice must start on a 16-byte boundary"); return -1; }
    if (count & 3) { set_error("multimem_allreduce: slice length must be a multiple of 4 elements"); return -1; }
    if (ctas < 1) ctas = 8;
    MmPeers p;
    for (int i = 0; i < MM_MAX_WORLD; ++i) p.flags[i] = i < world ? (unsigned int*)peer_flags[i] : nullptr;
    multimem_allreduce_kernel<<<ctas, 1024, 0, (cudaStream_t)stream>>>((float*)multicast_base, begin, count, p, rank, world, epoch);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { set_error("multimem_allreduce: %s", cudaGetErrorString(err)); return -1; }
    return 0;
}

}  // extern "C"
