// capi.cu -- library-wide C-ABI entry points and globals of libpointunet_b200.so
#include "common.cuh"

namespace pu {
thread_local int g_last_cuda_error = 0;
unsigned long long g_launch_count = 0;
}  // namespace pu

extern "C" {

const char *pu_version(void) { return "pointunet_b200 0.1 (sm_100a)"; }

int pu_last_cuda_error(void) { return pu::g_last_cuda_error; }

unsigned long long pu_launch_count(void) { return __atomic_load_n(&pu::g_launch_count, __ATOMIC_RELAXED); }

}  // extern "C"
