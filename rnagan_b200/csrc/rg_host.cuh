// rg_host.cuh -- host-side helpers shared by the C-ABI translation units: error reporting, device checks,
// TMA tensor-map encoding (through the driver entry point, so the library has no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/rnagan_b200.h"

namespace rg {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int num_sms();
void count_launch();

#define RG_CHECK_ARG(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      rg::set_error(__VA_ARGS__);      \
      return RG_EINVAL;                \
    }                                  \
  } while (0)

#define RG_CUDA(call)                                        \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return rg::cuda_fail(e__, #call); \
  } while (0)

// every kernel launch site goes through this macro: it also feeds rg_launch_count()
#define RG_LAUNCH_CHECK(name)                                   \
  do {                                                          \
    rg::count_launch();                                         \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return rg::cuda_fail(e__, name);    \
  } while (0)

// bf16 tensor map over an NHWC-like 4-D view: dims (C, W, H, B) with element strides (1, sw, sh, sb) and
// box (64, bw, bh, bb), SWIZZLE_128B, zero OOB fill.
int encode_map_4d(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint64_t sw,
                  uint64_t sh, uint64_t sb, uint32_t boxc, uint32_t bw, uint32_t bh, uint32_t bb);
// bf16 row-major matrix [rows][ld]: dims (cols, rows), box (64, box_rows).
int encode_map_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t boxc,
                  uint32_t box_rows);

// Launch with the programmatic-stream-serialisation attribute (RG_PDL=0: plain launch) and, for cluster > 1, the
// cluster dimension.  ONLY for kernels whose first global-memory access comes after griddep_wait().
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);
}

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace rg
