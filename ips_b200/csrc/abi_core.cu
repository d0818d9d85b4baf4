// Error plumbing + device probe for the C ABI (include/ips_b200.h).
#include "common.cuh"
#include "../../include/ips_b200.h"

namespace ipsb {
static thread_local char g_err[512] = "";
char* err_buf() { return g_err; }
int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
}  // namespace ipsb

extern "C" {
int ipsb_abi_version(void) { return IPSB_ABI_VERSION; }
const char* ipsb_last_error(void) { return ipsb::err_buf(); }
int ipsb_device_ok(void) {
    int dev = 0, major = 0;
    IPSB_CUDA(cudaGetDevice(&dev));
    IPSB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    IPSB_REQUIRE(major == 10, "device compute capability %d.x, this library is built for sm_100a only", major);
    return 0;
}
}
