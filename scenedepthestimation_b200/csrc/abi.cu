// Error channel, device queries and small ABI helpers.
#include "common.cuh"
#include <cstring>

namespace mccnn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    return dev;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace mccnn

extern "C" const char* mccnn_last_error(void) { return mccnn::g_err; }
extern "C" int mccnn_abi_version(void) { return MCCNN_ABI_VERSION; }
extern "C" int mccnn_disp_pitch(int D) { return mccnn::disp_pitch(D); }

extern "C" void mccnn_default_sgm_params(mccnn_sgm_params* p) {
    if (!p) return;
    // process_functional.py:1141-1144; the penalty kernel stores P/lambda computed in fp64 as fp32 (:141-142, 149)
    p->P1 = (float)2.3;
    p->P2 = (float)55.9;
    p->P1_red = (float)(2.3 / 4);
    p->P2_red = (float)(55.9 / 4);
    p->threshold = 30;
    p->subpixel = 0;
    p->bilateral = 0;
    p->cbca_iters = 0;  // the reference runs no aggregation
    p->cbca_L1 = 14;
    p->cbca_tau = 6;
}

extern "C" int mccnn_device_supported(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        mccnn::set_error("no CUDA device %d", device);
        return 0;
    }
    if (prop.major != 10) {
        mccnn::set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return 0;
    }
    return 1;
}
