// api.cu — library-level entry points: ABI version, error text, device capability check.
#include "common.cuh"
#include <mutex>

namespace bnn {

char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

namespace {
struct DevInfo {
  int known = 0;   // 0 = not probed, 1 = supported, 2 = unsupported
  int sms = 0;
};
DevInfo g_dev[64];
std::mutex g_dev_mutex;

int probe(int dev, DevInfo* out) {
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  if (dev < 0 || dev >= 64) return fail(BNN_ERR_BAD_ARGUMENT, "device index %d out of range", dev);
  if (!g_dev[dev].known) {
    int major = 0, sms = 0;
    BNN_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    BNN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    g_dev[dev].sms = sms;
    g_dev[dev].known = (major == 10) ? 1 : 2;
  }
  *out = g_dev[dev];
  return BNN_OK;
}
}  // namespace

int check_device() {
  int dev = 0;
  BNN_CUDA_OK(cudaGetDevice(&dev));
  DevInfo info;
  int rc = probe(dev, &info);
  if (rc != BNN_OK) return rc;
  if (info.known != 1)
    return fail(BNN_ERR_UNSUPPORTED_ARCH, "device %d is not compute capability 10.x (B200, sm_100a)", dev);
  return BNN_OK;
}

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  DevInfo info;
  if (probe(dev, &info) != BNN_OK || info.sms <= 0) return 148;
  return info.sms;
}

}  // namespace bnn

extern "C" {

int bnn_abi_version(void) { return BNN_B200_ABI_VERSION; }

const char* bnn_last_error_string(void) { return bnn::error_buffer(); }

int bnn_device_supported(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return bnn::fail(BNN_ERR_UNSUPPORTED_ARCH, "no CUDA device %d", device);
  int major = 0;
  BNN_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10)
    return bnn::fail(BNN_ERR_UNSUPPORTED_ARCH, "device %d has compute capability %d.x, need 10.x", device, major);
  return BNN_OK;
}

}  // extern "C"
