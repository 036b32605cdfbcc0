// Host-side helpers shared by the C-ABI translation units: thread-local error text, CUDA error checks,
// TMA tensor-map construction through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

namespace ldot {

enum : int {
  kOk = 0,
  kErrArg = -1,        // bad argument (shape, alignment, null pointer, unsupported k / d)
  kErrCuda = -2,       // a CUDA runtime / driver call failed
  kErrWorkspace = -3,  // caller-provided workspace too small
  kErrArch = -4,       // not an sm_100 device
};

char* error_buffer();  // thread-local, 512 bytes (defined in capi.cu)

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define LDOT_CUDA(expr)                                                                                   \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      return ::ldot::set_error(::ldot::kErrCuda, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                       \
  } while (0)

#define LDOT_CHECK_LAUNCH() LDOT_CUDA(cudaGetLastError())

#define LDOT_REQUIRE(cond, ...)                                                \
  do {                                                                         \
    if (!(cond)) return ::ldot::set_error(::ldot::kErrArg, __VA_ARGS__);  \
  } while (0)

// 2-D K-major tensor map: global [rows, cols] 16-bit elements, row pitch `row_stride_bytes`,
// box = [box_rows, 64 cols] (64 x 2 B = 128 B inner extent, SWIZZLE_128B).  Out-of-bounds reads are zero-filled,
// which is what pads the last M / N tile and a K that is not a multiple of 64.
int make_tmap_kmajor_16b(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t row_stride_bytes, uint32_t box_rows);

// 3-D view of a K-major 16-bit matrix whose row length is a multiple of 64: [64 (k in block), rows, cols / 64 (k block)],
// box = [64, box_rows, box_kb]: one TMA instruction lands box_kb consecutive [box_rows x 64] SWIZZLE_128B tiles.
int make_tmap_kblocks_16b(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                          uint32_t box_rows, uint32_t box_kb);

// 2-D tensor map of a row-major [rows, cols] matrix of `elt_bytes`-wide elements for TMA stores of
// [box_rows, box_cols] boxes whose inner extent (box_cols * elt_bytes) is 64 B (SWIZZLE_64B) or 128 B (SWIZZLE_128B).
// as_float: 4-byte elements are typed FLOAT32 (required by cp.reduce.async.bulk.tensor .add; plain stores do not care).
int make_tmap_store(CUtensorMap* out, const void* base, uint32_t elt_bytes, uint64_t rows, uint64_t cols,
                    uint64_t row_stride_bytes, uint32_t box_cols, uint32_t box_rows, bool as_float = false);

int device_sm_count(int* out);

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace ldot
