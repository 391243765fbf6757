// Library identification and status strings for libpem_b200.so.
#include "pem_common.cuh"

#ifndef PEM_BUILD_STAMP
#define PEM_BUILD_STAMP "unstamped"
#endif

extern "C" int pem_abi_version(void) { return 1; }

extern "C" const char *pem_build_info(void) {
    return "libpem_b200 abi=1 arch=sm_100a build=" PEM_BUILD_STAMP;
}

extern "C" const char *pem_status_string(int status) {
    switch (status) {
        case PEM_OK: return "ok";
        case PEM_E_NULL: return "required pointer is NULL";
        case PEM_E_SHAPE: return "invalid size argument";
        case PEM_E_UNSUPPORTED: return "configuration not supported by this build";
        case PEM_E_WORKSPACE: return "workspace missing or too small";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown status";
}

#ifdef PEM_DEV_KNOBS   // make DEV=1: programmatic dependent launches on/off for A/B runs (always on otherwise)
int g_pem_pdl = 1;
extern "C" void pem_dev_set_pdl(int on) { g_pem_pdl = on ? 1 : 0; }
#endif
