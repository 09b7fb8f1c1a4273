#include "common.cuh"

#include <string.h>

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

int cvb_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int cvb_fail_cuda(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return CV_ERR_CUDA;
}

void cvb_reset_launches() { g_launches = 0; }
void cvb_count_launch() { g_launches++; }

extern "C" const char* cv_last_error(void) { return g_err; }
extern "C" const char* cv_version(void) { return "circuitvision_b200 0.1.0 sm_100a"; }
extern "C" int cv_last_launch_count(void) { return g_launches; }

extern "C" int cv_device_is_sm100(int device) {
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}
