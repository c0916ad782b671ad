// library identity / device probe
#include "common.cuh"
#include "usvm2_b200.h"

extern "C" int usvm_abi_version(void) { return USVM_ABI_VERSION; }

extern "C" int usvm_device_sm(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return USVM_ERR_CUDA;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return USVM_ERR_CUDA;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return USVM_ERR_CUDA;
  return major * 10 + minor;
}
