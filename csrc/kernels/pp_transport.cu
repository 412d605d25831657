// Pipeline-boundary transport over peer memory (see runtime/pp_context.h): push / wait / credit.
// These are small copy + flag kernels on purpose: the boundary tile of the reference workload is 16 KB, the cost
// that matters is the ~10-15 us of an NCCL send/recv pair, which becomes one NVLink store stream + one flag.
#include "kernels/ptx.cuh"
#include "runtime/pp_context.h"

namespace ssb {

__global__ void __launch_bounds__(256) pp_push_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t n4,
                                                      uint32_t* flag_peer, const uint32_t* credit_local, const uint32_t* epoch_ptr,
                                                      uint32_t* done_counter) {
    const uint32_t epoch = *epoch_ptr;
    // the consumer must have released the slots of the previous step (credit == epoch - 1 after its step epoch - 1)
    if (threadIdx.x == 0) wait_flag_ge(credit_local, epoch - 1u);
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(done_counter, 1u);
        if (old == gridDim.x - 1u) {                       // last CTA: every part of the tile is on its way
            *done_counter = 0u;
            __threadfence_system();
            st_release_sys(flag_peer, epoch);
        }
    }
}

// Boundary tiles of narrow stages are a few KB: one CTA, no cross-CTA completion counter - so any number of these may be
// in flight at once, each on the stream of the micro-batch that produced its tile.
__global__ void __launch_bounds__(256) pp_push_small_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int n4,
                                                            uint32_t* flag_peer, const uint32_t* credit_local, const uint32_t* epoch_ptr) {
    const uint32_t epoch = *epoch_ptr;
    if (threadIdx.x == 0) wait_flag_ge(credit_local, epoch - 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                            // cumulative: the CTA's stores are ordered before the flag
        st_relaxed_sys(flag_peer, epoch);
    }
}

__global__ void pp_wait_kernel(const uint32_t* flag_local, const uint32_t* epoch_ptr) {
    if (threadIdx.x == 0) wait_flag_ge(flag_local, *epoch_ptr);
}

__global__ void pp_credit_kernel(uint32_t* credit_a, uint32_t* credit_b, const uint32_t* epoch_ptr) {
    if (threadIdx.x == 0) {
        const uint32_t epoch = *epoch_ptr;
        __threadfence_system();
        if (credit_a != nullptr) st_release_sys(credit_a, epoch);
        if (credit_b != nullptr) st_release_sys(credit_b, epoch);
    }
}

cudaError_t launch_pp_push(const float* src, float* dst_peer, int64_t n, uint32_t* flag_peer, const uint32_t* credit_local,
                           const uint32_t* epoch, uint32_t* done_counter, cudaStream_t stream) {
    if (n % 4 != 0) return cudaErrorInvalidValue;
    const int64_t n4 = n / 4;
    if (n4 <= kPpSmallTileF4) {
        pp_push_small_kernel<<<1, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst_peer), (int)n4,
                                                    flag_peer, credit_local, epoch);
        return cudaGetLastError();
    }
    int64_t ctas = (n4 + 256 * 4 - 1) / (256 * 4);
    if (ctas < 1) ctas = 1;
    if (ctas > 32) ctas = 32;
    pp_push_kernel<<<(int)ctas, 256, 0, stream>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<float4*>(dst_peer), n4, flag_peer,
                                                   credit_local, epoch, done_counter);
    return cudaGetLastError();
}

cudaError_t launch_pp_wait(const uint32_t* flag_local, const uint32_t* epoch, cudaStream_t stream) {
    pp_wait_kernel<<<1, 32, 0, stream>>>(flag_local, epoch);
    return cudaGetLastError();
}

cudaError_t launch_pp_credit(uint32_t* credit_a, uint32_t* credit_b, const uint32_t* epoch, cudaStream_t stream) {
    pp_credit_kernel<<<1, 32, 0, stream>>>(credit_a, credit_b, epoch);
    return cudaGetLastError();
}

}  // namespace ssb
