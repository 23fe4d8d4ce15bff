// cabi.cu -- library-level entry points of the C ABI (include/hmpc.h).
#include "common.cuh"

namespace hmpc {
thread_local char g_last_error[256] = "";
}

extern "C" int hmpc_version(void) { return 100; }  // 0.1.0

extern "C" const char* hmpc_last_cuda_error(void) { return hmpc::g_last_error; }

extern "C" int hmpc_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin_bytes) {
    int dev = 0;
    HMPC_CUDA_TRY(cudaGetDevice(&dev));
    cudaDeviceProp p;
    HMPC_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin_bytes) *smem_optin_bytes = p.sharedMemPerBlockOptin;
    return p.major == 10 ? HMPC_OK : HMPC_ERR_NO_DEVICE;
}
