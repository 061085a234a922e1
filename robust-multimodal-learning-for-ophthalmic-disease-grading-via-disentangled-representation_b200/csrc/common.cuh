// Host-side plumbing shared by the C-ABI translation units: error string, launch counter,
// TMA descriptor encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

namespace edrl {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define EDRL_CHECK_ARG(cond, ...)   \
  do {                              \
    if (!(cond)) {                  \
      ::edrl::set_error(__VA_ARGS__); \
      return 1;                     \
    }                               \
  } while (0)

#define EDRL_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::edrl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                              \
    }                                                                                        \
  } while (0)

// Count one kernel launch and surface launch-configuration errors immediately.
#define EDRL_LAUNCHED()                                                                         \
  do {                                                                                          \
    ::edrl::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      ::edrl::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 3;                                                                                 \
    }                                                                                           \
  } while (0)

// 2-D fp32 row-major tensor map: `cols` is the contiguous dimension, rows are `row_pitch_bytes`
// apart; the box is box_cols x box_rows and lands in shared memory with the 128-byte swizzle.
int make_tmap_2d_f32(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                     uint32_t box_rows, uint32_t box_cols);

// same for binary16 elements (cols / box_cols in elements; 64 halfs = one 128-byte swizzle row)
int make_tmap_2d_f16(CUtensorMap *out, const void *base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                     uint32_t box_rows, uint32_t box_cols);

int device_sm_count();

// The device the next C-ABI calls of this thread run on (edrl_set_device); -1 = whatever is current.
extern thread_local int g_target_device;

// Binds the calling thread to that device for the duration of ONE C-ABI call and restores the caller's current
// device afterwards: the library's (static) runtime shares the primary contexts with the host framework, so a
// lasting cudaSetDevice here would silently move the caller's later allocations to another GPU.
struct DeviceGuard {
  int prev = -1, dev = -1;
  DeviceGuard() {
    dev = g_target_device;
    if (dev < 0 || cudaGetDevice(&prev) != cudaSuccess) {
      prev = dev = -1;
      return;
    }
    // (also when dev == prev: cudaSetDevice binds the device's primary context to THIS thread.  A host framework's worker
    //  thread -- torch's autograd thread -- may have a device selected and no context current yet; driver entry points
    //  such as cuTensorMapEncodeTiled then fail with CUDA_ERROR_INVALID_CONTEXT when they are the thread's first call)
    if (cudaSetDevice(dev) != cudaSuccess) {
      (void)cudaGetLastError();
      prev = dev = -1;
    }
  }
  ~DeviceGuard() {
    if (prev >= 0 && dev != prev) (void)cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard &) = delete;
  DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define EDRL_DEVICE_GUARD() ::edrl::DeviceGuard _edrl_device_guard

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace edrl
