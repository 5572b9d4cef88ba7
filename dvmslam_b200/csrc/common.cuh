// common.cuh -- shared host/device helpers for libdvmslam_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/dvmslam_b200.h"

#define DVM_STR2(x) #x
#define DVM_STR(x) DVM_STR2(x)

namespace dvm {

// ---- error plumbing (no exceptions cross the C-ABI) ----
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define DVM_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            ::dvm::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,            \
                             cudaGetErrorString(e__));                                       \
            return DVM_ERR_CUDA;                                                             \
        }                                                                                    \
    } while (0)

#define DVM_REQUIRE(cond, msg)                                                               \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ::dvm::set_error("%s (%s) at %s:%d", msg, #cond, __FILE__, __LINE__);            \
            return DVM_ERR_INVALID;                                                          \
        }                                                                                    \
    } while (0)

// every kernel launch goes through this so dvm_kernel_launch_count() is truthful
#define DVM_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
    do {                                                                                     \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                          \
        ::dvm::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
    } while (0)

// Programmatic dependent launch for chains of small dependent kernels on one stream: the kernel is scheduled
// while its predecessor still runs (the predecessor calls pdl_trigger() at its top) and blocks in pdl_wait() until
// the predecessor has completed and its memory is visible, which takes the launch latency off the dependency
// chain.  Both device calls are no-ops in a launch without the attribute.
#define DVM_LAUNCH_PDL(kernel, grid_, block_, smem_, stream_, ...)                              \
    do {                                                                                     \
        cudaLaunchConfig_t cfg__ = {};                                                       \
        cfg__.gridDim = dim3(grid_); cfg__.blockDim = dim3(block_);                        \
        cfg__.dynamicSmemBytes = (smem_); cfg__.stream = (stream_);                        \
        cudaLaunchAttribute at__[1];                                                         \
        at__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                     \
        at__[0].val.programmaticStreamSerializationAllowed = 1;                              \
        cfg__.attrs = at__; cfg__.numAttrs = 1;                                              \
        cudaLaunchKernelEx(&cfg__, kernel, __VA_ARGS__);                                     \
        ::dvm::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
    } while (0)
__device__ inline void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ inline void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int select_device(int device); // validates sm_100, returns dvm_status

constexpr int kNumSMs = 148; // B200

__host__ __device__ inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---- block-wide exclusive scan over one int per thread (blockDim.x multiple of 32, <= 1024) ----
// returns the exclusive prefix of v; *total receives the block sum.  `warp_sums` is 33 ints of smem.
__device__ inline int block_exclusive_scan(int v, int* warp_sums, int* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads(); // protect warp_sums from a previous call
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int s = lane < nw ? warp_sums[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < nw) warp_sums[lane] = si - s; // exclusive warp offsets
        if (lane == 31) warp_sums[32] = si;
    }
    __syncthreads();
    *total = warp_sums[32];
    return warp_sums[wid] + incl - v;
}

} // namespace dvm
