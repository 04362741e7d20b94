// Error reporting, launch counter and misc. exports of librdm_b200.
#include "common.cuh"
#include <cstdlib>
#include "../../include/rdm_b200.h"

static thread_local char t_err[1024] = "";
unsigned long long g_rdm_launches = 0;
int g_rdm_use_pdl = getenv("RDM_PDL") ? atoi(getenv("RDM_PDL")) : 1;
int g_rdm_use_pdl_glue = getenv("RDM_PDL_GLUE") ? atoi(getenv("RDM_PDL_GLUE")) : 0;

void rdm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

int rdm_num_sms(int device) {
    static int cached[16] = {0};
    if (device >= 0 && device < 16 && cached[device]) return cached[device];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    if (device >= 0 && device < 16) cached[device] = n;
    return n;
}

extern "C" {
const char* rdm_last_error(void) { return t_err; }
int rdm_abi_version(void) { return 2; }
unsigned long long rdm_launch_count(void) { return g_rdm_launches; }
}
