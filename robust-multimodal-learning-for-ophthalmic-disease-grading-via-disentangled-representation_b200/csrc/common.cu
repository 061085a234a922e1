#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include "../../include/edrl_b200.h"

namespace edrl {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
thread_local int g_target_device = -1;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int make_tmap_2d(CUtensorMap *out, CUtensorMapDataType dt, const void *base, uint64_t rows, uint64_t cols,
                        uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols);

int make_tmap_2d_f32(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                     uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rows, cols, row_pitch_bytes, box_rows, box_cols);
}
int make_tmap_2d_f16(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                     uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rows, cols, row_pitch_bytes, box_rows, box_cols);
}

static int make_tmap_2d(CUtensorMap *out, CUtensorMapDataType dt, const void *base, uint64_t rows, uint64_t cols,
                        uint64_t row_pitch_bytes, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return 4;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu pitch=%llu box=%ux%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)row_pitch_bytes, box_rows,
              box_cols);
    return 4;
  }
  return 0;
}

int device_sm_count() {
  static int sms = 0;            // (one node carries one GPU model: the first device asked answers for all)
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
  }
  return sms;
}

}  // namespace edrl

extern "C" {
int edrl_abi_version(void) { return EDRL_ABI_VERSION; }
const char *edrl_last_error(void) { return edrl::g_err; }
uint64_t edrl_launch_count(void) { return edrl::g_launches.load(std::memory_order_relaxed); }
int edrl_set_device(int device) {
  int count = 0;
  EDRL_CUDA_OK(cudaGetDeviceCount(&count));
  EDRL_CHECK_ARG(device >= 0 && device < count, "edrl_set_device: device %d outside [0, %d)", device, count);
  edrl::g_target_device = device;      // bound per call by DeviceGuard; the caller's current device is left alone
  return 0;
}
}
