// C-ABI housekeeping: version, error strings, device check.
#include "hvae_common.cuh"

extern "C" int hvae_version(void) { return HVAE_VERSION; }

extern "C" const char* hvae_strerror(int code) {
    switch (code) {
        case HVAE_OK: return "ok";
        case HVAE_ESHAPE: return "unsupported or inconsistent shape";
        case HVAE_EALIGN: return "misaligned pointer";
        case HVAE_EARCH: return "device is not sm_100 (B200); hvae_b200 has no fallback path";
        case HVAE_ELAUNCH: return "CUDA kernel launch failed";
        case HVAE_EARG: return "null pointer, bad flag, or workspace too small";
        default: return "unknown hvae error";
    }
}

extern "C" int hvae_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return HVAE_EARCH;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return HVAE_EARCH;
    return major == 10 ? HVAE_OK : HVAE_EARCH;
}
