// Library state: per-thread error string and device binding.
#include <mutex>

#include "kgb_internal.h"

namespace kgb {

static thread_local std::string t_error;
static thread_local int t_device = -1;
static int g_sm_count[64];
static bool g_ready[64];

void set_error(const char* fmt, ...) {
    char buf[8192];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}
int current_device() { return t_device; }
int sm_count() { return (t_device >= 0 && t_device < 64) ? g_sm_count[t_device] : 0; }

}  // namespace kgb

extern "C" int kgb_abi_version(void) { return KGB_ABI_VERSION; }
extern "C" const char* kgb_last_error(void) { return kgb::t_error.c_str(); }

extern "C" int kgb_init(int device) {
    // re-binding a thread to a device that was already initialised (one interpreter driving several GPUs switches per
    // launch, klongpy_b200/sharded.py): cudaSetDevice only
    if (device >= 0 && device < 64 && kgb::g_sm_count[device] > 0 && kgb::g_ready[device]) {
        KGB_CUDA_CHECK(cudaSetDevice(device));
        kgb::t_device = device;
        return KGB_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        kgb::set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return KGB_ENOGPU;
    }
    KGB_REQUIRE(device >= 0 && device < ndev && device < 64, "kgb_init: device %d out of range (%d devices)", device, ndev);
    KGB_CUDA_CHECK(cudaSetDevice(device));
    KGB_CUDA_CHECK(cudaFree(0));
    cudaDeviceProp prop;
    KGB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        kgb::set_error("device %d is sm_%d%d; libkgb200 is built for sm_100a only", device, prop.major, prop.minor);
        return KGB_ENOGPU;
    }
    kgb::g_sm_count[device] = prop.multiProcessorCount;
    kgb::g_ready[device] = true;
    kgb::t_device = device;
    return KGB_OK;
}

extern "C" int kgb_device_sm_count(int device, int* out) {
    KGB_REQUIRE(out && device >= 0 && device < 64, "kgb_device_sm_count: bad arguments");
    if (kgb::g_sm_count[device] == 0) {
        cudaDeviceProp prop;
        KGB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        kgb::g_sm_count[device] = prop.multiProcessorCount;
    }
    *out = kgb::g_sm_count[device];
    return KGB_OK;
}
