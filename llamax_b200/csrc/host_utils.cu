#include "host_utils.h"

#include <stdio.h>
#include <string.h>

#include <mutex>

#include "llamax_b200.h"

namespace lx {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
  return LLAMAX_ERR_CUDA;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int esz, const void* ptr, int64_t inner, int64_t outer,
                 int64_t ld, int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(LLAMAX_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(2d) failed: CUresult %d (inner=%lld outer=%lld ld=%lld)",
             (int)r, (long long)inner, (long long)outer, (long long)ld);
    return set_error(LLAMAX_ERR_CUDA, buf);
  }
  return 0;
}

int make_tmap_4d(CUtensorMap* map, CUtensorMapDataType dt, int esz, const void* ptr, const int64_t dims_[4],
                 const int64_t strides_[3], const int box_[4]) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(LLAMAX_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4], strides[3];
  cuuint32_t box[4], estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 4; ++i) {
    dims[i] = (cuuint64_t)dims_[i];
    box[i] = (cuuint32_t)box_[i];
  }
  for (int i = 0; i < 3; ++i) strides[i] = (cuuint64_t)strides_[i] * esz;
  CUresult r = fn(map, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(4d) failed: CUresult %d", (int)r);
    return set_error(LLAMAX_ERR_CUDA, buf);
  }
  return 0;
}

}  // namespace lx

extern "C" {

const char* llamax_last_error(void) { return lx::g_err; }

int llamax_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return lx::set_cuda_error(e, "llamax_set_device");
  return 0;
}

int llamax_version(void) { return LLAMAX_B200_VERSION; }

}  // extern "C"
