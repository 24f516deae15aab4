// api.cu -- library-level entry points: error slot, version, device check, driver entry points.
#include "common.cuh"

namespace mst {

char* error_buffer() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 1024, fmt, ap);
  va_end(ap);
  return 1;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

PFN_encodeTiled tensor_map_encoder() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr) {
    fail("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

}  // namespace mst

extern "C" {

const char* mst_last_error(void) { return mst::error_buffer(); }

int mst_version(void) { return 100; }

int mst_device_check(int device) {
  cudaDeviceProp prop;
  MST_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  MST_CHECK(prop.major == 10, "device %d is sm_%d%d; libmst_b200 is built for sm_100a only", device, prop.major,
            prop.minor);
  MST_CHECK(mst::tensor_map_encoder() != nullptr, "%s", mst::error_buffer());
  return 0;
}

}  // extern "C"
