// Library-level entry points of libb200d.so: version, last error, device check.
#include "common.cuh"

namespace b200d {
thread_local char g_last_error[512] = "";
}

extern "C" const char* b200d_version(void) { return "b200d 0.1.0 (sm_100a)"; }
extern "C" const char* b200d_last_error(void) { return b200d::g_last_error; }

extern "C" int b200d_check_device(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess)
    return b200d::set_error(B200D_ELAUNCH, "%s: no CUDA device%s", "b200d_check_device");
  if (prop.major != 10) return b200d::set_error(B200D_EARCH, "%s: device is not sm_100 (Blackwell B200)%s", "b200d_check_device");
  return B200D_OK;
}
