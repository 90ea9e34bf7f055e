// Error reporting and device queries shared by all entry points.
#include "common.cuh"
#include <stdarg.h>

namespace gsb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;
    }
    return sms;
}

}  // namespace gsb

extern "C" int gsb_abi_version(void) { return GSB_ABI_VERSION; }
extern "C" const char *gsb_last_error(void) { return gsb::g_err; }
