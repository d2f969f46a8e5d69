// Host-side helpers shared by the C-ABI translation units: error reporting, TMA tensor-map
// encoding through the driver entry point (no link-time dependency on libcuda, so the library
// loads on a box without a GPU), device properties.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/bpreveal_b200.h"

namespace bp {

void set_error(const char* fmt, ...);
int sm_count();
int max_smem_optin();
// Device pointer of the 2048-word host-mapped watchdog buffer (nullptr if it cannot be allocated).
unsigned* debug_buffer();

// 3-D bf16 tensor [d2][d1][d0] (d0 innermost, contiguous), box [1][box1][64], 128B swizzle.
bool make_tmap_3d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                  uint32_t box0, uint32_t box1);
// 3-D bf16 tensor with explicit byte strides (multiples of 16) for dims 1 and 2; rows may overlap
// (stride smaller than the row extent), which turns a channels-last [L][C] tensor into the
// implicit im2col matrix of a width-W window: row t = elements [t*C, t*C + W*C).
bool make_tmap_3d_sw64(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2);
bool make_tmap_3d_strided(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2,
                          uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                          uint32_t box1);
// 2-D bf16 tensor [d1][d0], box [box1][64], 128B swizzle.
bool make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t d0, uint64_t d1, uint32_t box0,
                  uint32_t box1);

// Launch with programmatic stream serialization (PDL) unless BPB_PDL=0: the kernel's prologue
// overlaps the tail of its predecessor; the kernel itself must call pdl_wait() before touching
// data of earlier kernels.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                       cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// The same for a kernel of CTA pairs (cluster of two along x).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_pairs(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                             cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define BP_CHECK_CUDA(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      bp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                    __LINE__);                                                           \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

#define BP_REQUIRE(cond, ...)      \
  do {                             \
    if (!(cond)) {                 \
      bp::set_error(__VA_ARGS__);  \
      return 2;                    \
    }                              \
  } while (0)

}  // namespace bp
