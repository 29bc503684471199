// api.cu -- error plumbing and library-level entry points of librwr_b200.
#include <cstdarg>

#include "common.cuh"

static thread_local char g_last_error[1024] = "";
thread_local cudaStream_t tls_alloc_stream = nullptr;

void rwr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

extern "C" {

int rwr_abi_version(void) { return RWR_ABI_VERSION; }

int rwr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char* rwr_last_error(void) { return g_last_error; }

}  // extern "C"
