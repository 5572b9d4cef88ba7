// common.cu -- error text, launch counter, device selection.
#include "common.cuh"
#include <stdarg.h>

namespace dvm {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{ 0 };

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

int select_device(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device visible (%s); libdvmslam_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return DVM_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("device ordinal %d out of range (%d devices)", device, count);
        return DVM_ERR_INVALID;
    }
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e != cudaSuccess) {
        set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return DVM_ERR_CUDA;
    }
    if (major != 10) {
        set_error("device %d has compute capability %d.x; this library is built for sm_100a only", device, major);
        return DVM_ERR_NO_DEVICE;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return DVM_ERR_CUDA;
    }
    return DVM_OK;
}

} // namespace dvm

extern "C" {

const char* dvm_last_error(void) { return dvm::t_error; }
const char* dvm_version(void) { return "dvmslam_b200 0.1 (sm_100a, CUDA " DVM_STR(CUDART_VERSION) ")"; }
uint64_t dvm_kernel_launch_count(void) { return dvm::g_launches.load(); }

} // extern "C"
