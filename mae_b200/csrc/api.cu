// Library-level entry points: ABI version, error strings, device query.
#include <cstdlib>

#include "common.cuh"

namespace mae {
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

namespace {
int g_pdl = -1;   // -1: not decided yet (MAE_NO_PDL is consulted on first use), 0 off, 1 on
}
bool pdl_enabled() {
    if (g_pdl < 0) {
        const char* e = getenv("MAE_NO_PDL");
        g_pdl = (e != nullptr && e[0] == '1') ? 0 : 1;
    }
    return g_pdl != 0;
}
}  // namespace mae

extern "C" int mae_set_pdl(int enabled) {
    const int before = mae::pdl_enabled() ? 1 : 0;
    mae::g_pdl = enabled ? 1 : 0;
    return before;
}

extern "C" int mae_abi_version(void) { return MAE_ABI_VERSION; }

extern "C" const char* mae_strerror(int code) {
    switch (code) {
        case MAE_OK: return "ok";
        case MAE_ERR_NULL: return "mae_b200: required pointer is NULL";
        case MAE_ERR_SHAPE: return "mae_b200: invalid or inconsistent size";
        case MAE_ERR_STRIDE: return "mae_b200: row stride smaller than row length";
        case MAE_ERR_WORKSPACE: return "mae_b200: workspace too small";
        case MAE_ERR_ALIGN: return "mae_b200: pointer not 16-byte aligned";
        case MAE_ERR_UNSUPPORTED: return "mae_b200: unsupported configuration";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "mae_b200: unknown error";
}

extern "C" int mae_device_info(int* sm, int* major, int* minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    int a = 0, b = 0, c = 0;
    if ((e = cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return static_cast<int>(e);
    if ((e = cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess) return static_cast<int>(e);
    if ((e = cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess) return static_cast<int>(e);
    if (sm) *sm = a;
    if (major) *major = b;
    if (minor) *minor = c;
    return MAE_OK;
}
