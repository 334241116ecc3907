// rg_ops.cu -- the HBM-bound kernels of the RNA-GAN hot path:
// BatchNorm statistics / apply / backward / double-backward fused with LeakyReLU, latent preparation, image
// im2col (with the gradient-penalty interpolation and the tanh backward fused in), the critic head + WGAN losses,
// the gradient-penalty scalar, and a multi-tensor Adam that re-emits the packed bf16 operands.
//
// All activations are bf16 NHWC viewed as [M = B*H*W rows][C channels]; every kernel moves 16-byte vectors
// (8 channels) with the channel index fastest so warps read/write whole 128-byte lines.  Per-channel reductions
// are two-stage and summed in a fixed order (deterministic; the reference sets cudnn.deterministic=True,
// src/histopathology_gan.py:289).
#include <algorithm>
#include "rg_host.cuh"
#include "rg_ptx.cuh"
#include <cuda_bf16.h>

namespace rg {

// ------------------------------------------------------------------------------------------------ vector helpers
struct Vec8 {
  float v[8];
};
__device__ __forceinline__ Vec8 ld8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  Vec8 r;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const Vec8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ uint32_t pack_bf16x2_ops(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ Vec8 ldf8(const float* p) {
  // per-channel parameter vectors are read-only for the kernel's lifetime: ld.global.nc lets the compiler hoist
  // these loads out of the row loops when the channel group is loop-invariant
  Vec8 r;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// ------------------------------------------------------------------------------------------------ column reduce
// Stage 1: partial[block][k][c] = sum over the block's rows of f_k(row, c).  Functor F: static K, a register struct
// F::In, `In load(row, c0)` (global loads only) and `void acc(const In&, c0, float (&acc)[K][8])` (math only).
// The split keeps FOUR rows of loads in flight per thread before any arithmetic: written as one call per row the
// compiler interleaved load / compute and the kernels ran latency-bound at 38-48 % of HBM peak.
__device__ __forceinline__ Vec8 unpack8(const uint4& u) {
  Vec8 r;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
// streaming 16-byte load; `asm volatile` keeps the batch of loads a thread issues ahead of its arithmetic in program
// order (ptxas otherwise sinks each load next to its first use, leaving one row in flight per thread)
__device__ __forceinline__ uint4 ldu4(const __nv_bfloat16* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kRedRows = 4;                       // rows per batch per thread

// Functor F: static K (sums), static NT (input tensors), `const bf16* ptr(int t)`, and
// `void acc(const uint4 (&d)[NT], int c0, float (&acc)[K][8])`.
// Every thread issues a batch of kRedRows x NT 16-byte streaming loads (ldu4: `asm volatile`, so ptxas keeps the whole
// batch ahead of the arithmetic instead of sinking each load next to its use) and only then accumulates; 2-4 CTAs per
// SM keep ~100 KiB per SM in flight.  (The previous version staged the rows in shared memory with cp.async: ncu showed
// L1TEX at 88 % of its peak with DRAM at 45 % -- the staging traffic, not HBM, was the limit; profiles/r2_ncu_hbm_summary.txt.)
template <class F>
__global__ void __launch_bounds__(256, F::NT >= 3 ? 2 : (F::NT == 2 ? 3 : 4)) colreduce_stage1(F f, int M, int C, int rows_per_block, float* __restrict__ partial) {
  griddep_wait();
  griddep_launch();
  constexpr int K = F::K;
  constexpr int NT = F::NT;
  constexpr int R = kRedRows;
  extern __shared__ __align__(16) uint8_t smraw[];
  float* sm = reinterpret_cast<float*>(smraw);   // [lanes][K][C] cross-lane reduction scratch
  const int cgs = C >> 3;
  const int lanes = cgs >= 256 ? 1 : 256 / cgs;            // row lanes per block iteration
  const int rl = cgs >= 256 ? 0 : threadIdx.x / cgs;
  const int cg0 = cgs >= 256 ? threadIdx.x : threadIdx.x % cgs;
  const bool active = cgs >= 256 ? true : (threadIdx.x < lanes * cgs);
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  for (int cg = cg0; cg < cgs; cg += 256) {
    float acc[K][8];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[k][e] = 0.0f;
    if (active) {
      const int c0 = cg * 8;
      const int step = R * lanes;
      int r = r0 + rl;
      for (; r + (R - 1) * lanes < r1; r += step) {          // full batches: R x NT loads, then the arithmetic
        uint4 d[R][NT];
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
          for (int t = 0; t < NT; ++t) d[j][t] = ldu4(f.ptr(t) + static_cast<size_t>(r + j * lanes) * C + c0);
#pragma unroll
        for (int j = 0; j < R; ++j) f.acc(d[j], c0, acc);
      }
      for (; r < r1; r += lanes) {                           // ragged tail, one row at a time
        uint4 d[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) d[t] = ldu4(f.ptr(t) + static_cast<size_t>(r) * C + c0);
        f.acc(d, c0, acc);
      }
    }
    if (lanes == 1) {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e)
          partial[(static_cast<size_t>(blockIdx.x) * K + k) * C + cg * 8 + e] = acc[k][e];
    } else {
      // sm[rl][k][c]
      if (active) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int e = 0; e < 8; ++e) sm[(rl * K + k) * C + cg * 8 + e] = acc[k][e];
      }
      __syncthreads();
      for (int i = threadIdx.x; i < K * C; i += 256) {
        float s = 0.0f;
        for (int l = 0; l < lanes; ++l) s += sm[l * K * C + i];
        partial[static_cast<size_t>(blockIdx.x) * K * C + i] = s;
      }
    }
  }
}

// Stage 2: out[k][c] = sum_blocks partial[block][k][c]; 32 columns x 32 block-lanes per CTA (narrow layers have few
// columns, so the parallelism has to come from the block dimension), four loads in flight per thread, lanes summed
// in a fixed order
__global__ void __launch_bounds__(1024) colreduce_stage2(const float* __restrict__ partial, int nblocks, int KC,
                                                         float* __restrict__ out) {
  griddep_wait();
  griddep_launch();
  __shared__ float sm[32][33];
  const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cx;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  if (i < KC) {
    int b = g;
    for (; b + 96 < nblocks; b += 128) {
      s0 += partial[static_cast<size_t>(b) * KC + i];
      s1 += partial[static_cast<size_t>(b + 32) * KC + i];
      s2 += partial[static_cast<size_t>(b + 64) * KC + i];
      s3 += partial[static_cast<size_t>(b + 96) * KC + i];
    }
    for (; b < nblocks; b += 32) s0 += partial[static_cast<size_t>(b) * KC + i];
  }
  sm[g][cx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (g == 0 && i < KC) {
    float t = sm[0][cx];
#pragma unroll
    for (int l = 1; l < 32; ++l) t += sm[l][cx];
    out[i] = t;
  }
}

// Few rows (BatchNorm1d over a batch of 128 in the betaVAE step, the top of the conv stacks in small jobs): the two-stage
// plan above would run 16 CTAs and a second launch for 1.5 MB.  Here one CTA owns a 128-channel column block over ALL rows
// -- 16 column groups x 16 row lanes, the same batched streaming loads -- and combines the lanes in a fixed order: one
// launch, no workspace, bit-reproducible.
constexpr int kSmallLanes = 16;
constexpr int kSmallM = 256;
template <class F>
__global__ void __launch_bounds__(256) colreduce_small(F f, int M, int C, float* __restrict__ out) {
  griddep_wait();
  griddep_launch();
  constexpr int K = F::K;
  constexpr int NT = F::NT;
  constexpr int R = kRedRows;
  __shared__ float sm[kSmallLanes][K][16 * 8];
  const int cgs = C >> 3;
  const int cgl = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int cg = blockIdx.x * 16 + cgl;
  float acc[K][8];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.0f;
  if (cg < cgs) {
    const int c0 = cg * 8;
    int r = rl;
    for (; r + (R - 1) * kSmallLanes < M; r += R * kSmallLanes) {      // full batches: R x NT loads, then the arithmetic
      uint4 d[R][NT];
#pragma unroll
      for (int j = 0; j < R; ++j)
#pragma unroll
        for (int t = 0; t < NT; ++t) d[j][t] = ldu4(f.ptr(t) + static_cast<size_t>(r + j * kSmallLanes) * C + c0);
#pragma unroll
      for (int j = 0; j < R; ++j) f.acc(d[j], c0, acc);
    }
    for (; r < M; r += kSmallLanes) {                                   // ragged tail, one row at a time
      uint4 d[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) d[t] = ldu4(f.ptr(t) + static_cast<size_t>(r) * C + c0);
      f.acc(d, c0, acc);
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) sm[rl][k][cgl * 8 + e] = acc[k][e];
  __syncthreads();
  for (int i = threadIdx.x; i < K * 128; i += 256) {
    const int k = i >> 7, col = i & 127;
    const int c = blockIdx.x * 128 + col;
    if (c < C) {
      float t = sm[0][k][col];
#pragma unroll
      for (int l = 1; l < kSmallLanes; ++l) t += sm[l][k][col];
      out[static_cast<size_t>(k) * C + c] = t;
    }
  }
}

struct ReducePlan {
  int blocks, rows_per_block;
  size_t smem;
};
static ReducePlan plan_reduce(int M, int C, int K, int NT = 1) {
  ReducePlan p;
  const int cgs = C / 8;
  const int lanes = cgs >= 256 ? 1 : 256 / cgs;
  // one wave of resident CTAs: 4 / 3 / 2 per SM for 1 / 2 / 3 input tensors (register budget of the load batches)
  // (NT = 1 yields the most blocks: the workspace query uses it as the upper bound.)
  int target = num_sms() * (NT <= 1 ? 4 : (NT == 2 ? 3 : 2));
  int rpb = std::max(lanes * 8, ceil_div(M, target));
  rpb = ceil_div(rpb, lanes) * lanes;
  p.rows_per_block = rpb;
  p.blocks = ceil_div(M, rpb);
  p.smem = lanes > 1 ? static_cast<size_t>(lanes) * K * C * sizeof(float) : 0;   // cross-lane scratch
  return p;
}
static size_t reduce_ws_floats(int M, int C, int K) {
  ReducePlan p = plan_reduce(M, C, K);
  return static_cast<size_t>(p.blocks) * K * C;
}

template <class F>
static int run_colreduce(F f, int M, int C, float* ws, size_t ws_bytes, float* out, cudaStream_t st, const char* name) {
  constexpr int K = F::K;
  if (C % 8 != 0 || C < 8 || C > 65536 || M <= 0) {
    set_error("%s: need C %% 8 == 0, 8 <= C <= 65536, M > 0 (M=%d C=%d)", name, M, C);
    return RG_EINVAL;
  }
  if (M <= kSmallM) {
    RG_CUDA(launch_pdl(colreduce_small<F>, dim3(ceil_div(C / 8, 16)), dim3(256), 0, st, 1, f, M, C, out));
    RG_LAUNCH_CHECK(name);
    return 0;
  }
  ReducePlan p = plan_reduce(M, C, K, F::NT);
  const size_t need = static_cast<size_t>(p.blocks) * K * C * sizeof(float);
  if (!ws || ws_bytes < need) {
    set_error("%s: reduction workspace too small (need %zu, have %zu)", name, need, ws_bytes);
    return RG_EWORKSPACE;
  }
  const size_t smem = p.smem;
  if (smem > 48 * 1024) {
    static bool attr_done = false;   // per instantiation
    if (!attr_done) {
      RG_CUDA(cudaFuncSetAttribute(colreduce_stage1<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_done = true;
    }
  }
  RG_CUDA(launch_pdl(colreduce_stage1<F>, dim3(p.blocks), dim3(256), smem, st, 1, f, M, C, p.rows_per_block, ws));
  RG_LAUNCH_CHECK(name);
  RG_CUDA(launch_pdl(colreduce_stage2, dim3(ceil_div(K * C, 32)), dim3(1024), 0, st, 1, ws, p.blocks, K * C, out));
  RG_LAUNCH_CHECK(name);
  return 0;
}

// ------------------------------------------------------------------------------------------------ elementwise driver
// Functor: __device__ void operator()(size_t row, int c0) const  -- processes 8 channels of one row
template <class F>
__global__ void __launch_bounds__(256) ew_kernel(F f, unsigned nvec, int cgs, int cg_shift) {
  griddep_wait();
  griddep_launch();
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg_shift >= 0) {
    // cgs is a power of two and divides the stride: this thread's channel group never changes, so the per-channel
    // parameter loads inside f are loop-invariant and there is no integer division
    const int c0 = static_cast<int>(i0 & static_cast<unsigned>(cgs - 1)) * 8;
    unsigned i = i0;
    for (; i + stride < nvec && i + stride > i; i += 2 * stride) {    // two independent rows in flight
      f(static_cast<size_t>(i >> cg_shift), c0);
      f(static_cast<size_t>((i + stride) >> cg_shift), c0);
    }
    if (i < nvec) f(static_cast<size_t>(i >> cg_shift), c0);
  } else {
    for (unsigned i = i0; i < nvec; i += stride) {
      const unsigned row = i / static_cast<unsigned>(cgs);
      f(static_cast<size_t>(row), static_cast<int>(i - row * cgs) * 8);
    }
  }
}
template <class F>
static int run_ew(F f, int M, int C, cudaStream_t st, const char* name) {
  if (C % 8 != 0 || M <= 0) {
    set_error("%s: need C %% 8 == 0 and M > 0 (M=%d C=%d)", name, M, C);
    return RG_EINVAL;
  }
  const size_t nvec = static_cast<size_t>(M) * (C / 8);
  if (nvec >= (1ull << 31)) {
    set_error("%s: tensor too large for 32-bit vector indexing (%zu vectors)", name, nvec);
    return RG_EINVAL;
  }
  const int cgs = C / 8;
  int grid = static_cast<int>(std::min<size_t>((nvec + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  int shift = -1;
  if (is_pow2(cgs)) {
    const int mult = std::max(1, cgs / 256);             // stride = grid*256 must be a multiple of cgs
    grid = std::max(mult, grid / mult * mult);
    shift = 0;
    while ((1 << shift) < cgs) ++shift;
  }
  RG_CUDA(launch_pdl(ew_kernel<F>, dim3(grid), dim3(256), 0, st, 1, f, static_cast<unsigned>(nvec), cgs, shift));
  RG_LAUNCH_CHECK(name);
  return 0;
}

// Same driver for functors with the load / apply split (`In load(row, c0)`, `void apply(const In&, row, c0)`):
// four rows of loads are issued before the first store.
template <class F, class = void>
struct ew_min_ctas { static constexpr int value = 1; };
template <class F>
struct ew_min_ctas<F, decltype(void(F::MIN_CTAS))> { static constexpr int value = F::MIN_CTAS; };


template <class F>
__global__ void __launch_bounds__(256, ew_min_ctas<F>::value) ew_split_kernel(F f, unsigned nvec, int cgs, int cg_shift) {
  griddep_wait();
  griddep_launch();
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg_shift >= 0) {
    const int c0 = static_cast<int>(i0 & static_cast<unsigned>(cgs - 1)) * 8;
    unsigned i = i0;
    for (; static_cast<unsigned long long>(i) + 3ull * stride < nvec; i += 4 * stride) {
      const size_t ra = i >> cg_shift, rb = (i + stride) >> cg_shift, rc = (i + 2 * stride) >> cg_shift,
                   rd = (i + 3 * stride) >> cg_shift;
      const typename F::In d0 = f.load(ra, c0);
      const typename F::In d1 = f.load(rb, c0);
      const typename F::In d2 = f.load(rc, c0);
      const typename F::In d3 = f.load(rd, c0);
      f.apply(d0, ra, c0);
      f.apply(d1, rb, c0);
      f.apply(d2, rc, c0);
      f.apply(d3, rd, c0);
    }
    for (; i < nvec && i >= i0; i += stride) {
      const size_t ra = i >> cg_shift;
      const typename F::In d0 = f.load(ra, c0);
      f.apply(d0, ra, c0);
      if (i + stride < i) break;   // 32-bit wrap
    }
  } else {
    for (unsigned i = i0; i < nvec; i += stride) {
      const unsigned row = i / static_cast<unsigned>(cgs);
      const int c0 = static_cast<int>(i - row * cgs) * 8;
      const typename F::In d0 = f.load(static_cast<size_t>(row), c0);
      f.apply(d0, static_cast<size_t>(row), c0);
      if (i + stride < i) break;
    }
  }
}
template <class F>
static int run_ew_split(F f, int M, int C, cudaStream_t st, const char* name) {
  if (C % 8 != 0 || M <= 0) {
    set_error("%s: need C %% 8 == 0 and M > 0 (M=%d C=%d)", name, M, C);
    return RG_EINVAL;
  }
  const size_t nvec = static_cast<size_t>(M) * (C / 8);
  if (nvec >= (1ull << 31)) {
    set_error("%s: tensor too large for 32-bit vector indexing (%zu vectors)", name, nvec);
    return RG_EINVAL;
  }
  const int cgs = C / 8;
  // 4 rows per thread per trip: a grid of ~4 CTAs per SM still covers the tensor in a few trips
  int grid = static_cast<int>(std::min<size_t>((nvec + 1023) / 1024, static_cast<size_t>(num_sms()) * 8));
  grid = std::max(grid, 1);
  int shift = -1;
  if (is_pow2(cgs)) {
    const int mult = std::max(1, cgs / 256);             // stride = grid*256 must be a multiple of cgs
    grid = std::max(mult, grid / mult * mult);
    shift = 0;
    while ((1 << shift) < cgs) ++shift;
  }
  RG_CUDA(launch_pdl(ew_split_kernel<F>, dim3(grid), dim3(256), 0, st, 1, f, static_cast<unsigned>(nvec), cgs, shift));
  RG_LAUNCH_CHECK(name);
  return 0;
}

// ------------------------------------------------------------------------------------------------ BN functors
struct StatsF {   // sum a, sum a^2
  static constexpr int K = 2;
  static constexpr int NT = 1;
  const __nv_bfloat16* a;
  int C;
  __device__ __forceinline__ const __nv_bfloat16* ptr(int) const { return a; }
  __device__ __forceinline__ void acc(const uint4 (&d)[1], int, float (&acc)[2][8]) const {
    const Vec8 x = unpack8(d[0]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += x.v[e];
      acc[1][e] += x.v[e] * x.v[e];
    }
  }
};

struct BnActF {   // h = lrelu(scale*a + shift)
  struct In { uint4 x; };
  static constexpr int MIN_CTAS = 6;
  const __nv_bfloat16* a;
  __nv_bfloat16* h;
  const float* scale;
  const float* shift;
  float slope;
  int C;
  __device__ __forceinline__ In load(size_t row, int c0) const { return In{ldu4(a + row * C + c0)}; }
  __device__ __forceinline__ void apply(const In& d, size_t row, int c0) const {
    const Vec8 x = unpack8(d.x);
    const Vec8 sc = ldf8(scale + c0), sh = ldf8(shift + c0);
    Vec8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float u = fmaf(x.v[e], sc.v[e], sh.v[e]);
      o.v[e] = u > 0.0f ? u : u * slope;
    }
    st8(h + row * C + c0, o);
  }
};

// du = dh * lrelu'(u), u = scale*a + shift, xhat = (a - mean) * rstd
struct BwdReduceF {   // S(du), S(du*xhat)
  static constexpr int K = 2;
  static constexpr int NT = 2;
  const __nv_bfloat16* dh;
  const __nv_bfloat16* a;
  const float *mean, *rstd, *scale, *shift;
  float slope;
  int C;
  __device__ __forceinline__ const __nv_bfloat16* ptr(int t) const { return t == 0 ? dh : a; }
  __device__ __forceinline__ void acc(const uint4 (&d)[2], int c0, float (&acc)[2][8]) const {
    const Vec8 g = unpack8(d[0]), x = unpack8(d[1]);
    const Vec8 mu = ldf8(mean + c0), rs = ldf8(rstd + c0), sc = ldf8(scale + c0), sh = ldf8(shift + c0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float u = fmaf(x.v[e], sc.v[e], sh.v[e]);
      const float du = u > 0.0f ? g.v[e] : g.v[e] * slope;
      acc[0][e] += du;
      acc[1][e] += du * (x.v[e] - mu.v[e]) * rs.v[e];
    }
  }
};

// da = scale*(du - s1/M - xhat*s2/M) (+ add); optional du output.  With xhat = (a - mean)*rstd this is
//   da = scale*du - k2*a + k0 (+ add),  k2 = scale*rstd*s2/M,  k0 = k2*mean - scale*s1/M:
// four per-channel vectors (scale, shift for the mask, k2, k0) stay live across the row loop instead of six, and the
// `add` operand is a template flag, so the kernel fits three CTAs per SM (ncu: it ran at 2 with 102 registers and 24 %
// active warps, profiles/r2_ncu_hbm_summary.txt).
template <bool HAS_ADD>
struct BwdApplyF {
  struct In { uint4 g, x, ad; };
  static constexpr int MIN_CTAS = 3;
  const __nv_bfloat16* dh;
  const __nv_bfloat16* a;
  const __nv_bfloat16* add;
  __nv_bfloat16* da;
  __nv_bfloat16* du_out;
  const float *mean, *rstd, *scale, *shift, *sums;   // sums[0][C] = S(du), sums[1][C] = S(du*xhat)
  float slope, invM;
  int C;
  __device__ __forceinline__ In load(size_t row, int c0) const {
    const size_t off = row * C + c0;
    In d;
    d.g = ldu4(dh + off);
    d.x = ldu4(a + off);
    if (HAS_ADD) d.ad = ldu4(add + off);
    return d;
  }
  __device__ __forceinline__ void apply(const In& in, size_t row, int c0) const {
    const size_t off = row * C + c0;
    const Vec8 g = unpack8(in.g), x = unpack8(in.x);
    const Vec8 sc = ldf8(scale + c0), sh = ldf8(shift + c0);
    Vec8 k2, k0;
    {   // loop-invariant for this thread (its channel group is fixed): hoisted out of the row loop
      const Vec8 mu = ldf8(mean + c0), rs = ldf8(rstd + c0), s1 = ldf8(sums + c0), s2 = ldf8(sums + C + c0);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        k2.v[e] = sc.v[e] * rs.v[e] * s2.v[e] * invM;
        k0.v[e] = k2.v[e] * mu.v[e] - sc.v[e] * s1.v[e] * invM;
      }
    }
    Vec8 o, d;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float u = fmaf(x.v[e], sc.v[e], sh.v[e]);
      const float du = u > 0.0f ? g.v[e] : g.v[e] * slope;
      d.v[e] = du;
      o.v[e] = fmaf(sc.v[e], du, fmaf(-k2.v[e], x.v[e], k0.v[e]));
    }
    if (HAS_ADD) {
      const Vec8 ad = unpack8(in.ad);
#pragma unroll
      for (int e = 0; e < 8; ++e) o.v[e] += ad.v[e];
    }
    st8(da + off, o);
    if (du_out) st8(du_out + off, d);
  }
};

struct LreluBwdF {   // da = dh * lrelu'(h)   (layer without BatchNorm: mask from the stored activation)
  struct In { uint4 g, x; };
  static constexpr int MIN_CTAS = 6;
  const __nv_bfloat16* dh;
  const __nv_bfloat16* h;
  __nv_bfloat16* da;
  float slope;
  int C;
  __device__ __forceinline__ In load(size_t row, int c0) const {
    return In{ldu4(dh + row * C + c0), ldu4(h + row * C + c0)};
  }
  __device__ __forceinline__ void apply(const In& in, size_t row, int c0) const {
    const Vec8 g = unpack8(in.g), x = unpack8(in.x);
    Vec8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = x.v[e] > 0.0f ? g.v[e] : g.v[e] * slope;
    st8(da + row * C + c0, o);
  }
};

struct ColSumF {   // S(x)
  static constexpr int K = 1;
  static constexpr int NT = 1;
  const __nv_bfloat16* x;
  int C;
  __device__ __forceinline__ const __nv_bfloat16* ptr(int) const { return x; }
  __device__ __forceinline__ void acc(const uint4 (&d)[1], int, float (&acc)[1][8]) const {
    const Vec8 v = unpack8(d[0]);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] += v.v[e];
  }
};

// gradient-penalty double backward through BatchNorm: ggI = adjoint of da, gO = du
struct GpReduceF {   // S(ggI), S(ggI*xhat), S(ggI*gO)
  static constexpr int K = 3;
  static constexpr int NT = 3;
  const __nv_bfloat16* ggI;
  const __nv_bfloat16* a;
  const __nv_bfloat16* gO;
  const float *mean, *rstd;
  int C;
  __device__ __forceinline__ const __nv_bfloat16* ptr(int t) const { return t == 0 ? ggI : (t == 1 ? a : gO); }
  __device__ __forceinline__ void acc(const uint4 (&d)[3], int c0, float (&acc)[3][8]) const {
    const Vec8 gi = unpack8(d[0]), x = unpack8(d[1]), go = unpack8(d[2]);
    const Vec8 mu = ldf8(mean + c0), rs = ldf8(rstd + c0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += gi.v[e];
      acc[1][e] += gi.v[e] * (x.v[e] - mu.v[e]) * rs.v[e];
      acc[2][e] += gi.v[e] * go.v[e];
    }
  }
};

struct GpApplyF {
  // A_dh = lrelu'(u) * gamma*r/M*(M*ggI - q1 - xhat*q2)
  // A_a  = gamma*r^2/M * [ xhat*(q1*s1/M - q3 + 3*s2*q2/M) + q2*(s1/M - gO) + s2*(q1/M - ggI) ]
  struct In { uint4 gi, x, go; };
  static constexpr int MIN_CTAS = 2;
  const __nv_bfloat16* ggI;
  const __nv_bfloat16* a;
  const __nv_bfloat16* gO;
  __nv_bfloat16* A_dh;
  __nv_bfloat16* A_a;
  const float *mean, *rstd, *gamma, *scale, *shift, *s, *q;   // s[2][C], q[3][C]
  float slope, invM;
  int C;
  __device__ __forceinline__ In load(size_t row, int c0) const {
    const size_t off = row * C + c0;
    return In{ldu4(ggI + off), ldu4(a + off), ldu4(gO + off)};
  }
  __device__ __forceinline__ void apply(const In& in, size_t row, int c0) const {
    const size_t off = row * C + c0;
    const Vec8 gi = unpack8(in.gi), x = unpack8(in.x), go = unpack8(in.go);
    const Vec8 mu = ldf8(mean + c0), rs = ldf8(rstd + c0), ga = ldf8(gamma + c0), sc = ldf8(scale + c0),
               sh = ldf8(shift + c0);
    const Vec8 s1 = ldf8(s + c0), s2 = ldf8(s + C + c0);
    const Vec8 q1 = ldf8(q + c0), q2 = ldf8(q + C + c0), q3 = ldf8(q + 2 * C + c0);
    Vec8 o1, o2;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (x.v[e] - mu.v[e]) * rs.v[e];
      const float u = fmaf(x.v[e], sc.v[e], sh.v[e]);
      const float gr = ga.v[e] * rs.v[e];
      const float adu = gr * (gi.v[e] - q1.v[e] * invM - xh * q2.v[e] * invM);
      o1.v[e] = u > 0.0f ? adu : adu * slope;
      const float k = gr * rs.v[e] * invM;
      o2.v[e] = k * (xh * (q1.v[e] * s1.v[e] * invM - q3.v[e] + 3.0f * s2.v[e] * q2.v[e] * invM) +
                     q2.v[e] * (s1.v[e] * invM - go.v[e]) + s2.v[e] * (q1.v[e] * invM - gi.v[e]));
    }
    st8(A_dh + off, o1);
    st8(A_a + off, o2);
  }
};

// ------------------------------------------------------------------------------------------------ small kernels
__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, int C, float invM, float unbias, float eps,
                                   float momentum, float* running_mean, float* running_var, long long* nbt,
                                   float* mean, float* rstd, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  const float m = sums[c] * invM;
  const float var = fmaxf(sums[C + c] * invM - m * m, 0.0f);
  const float r = rsqrtf(var + eps);
  mean[c] = m;
  rstd[c] = r;
  const float sc = gamma[c] * r;
  scale[c] = sc;
  shift[c] = beta[c] - m * sc;
  if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * m;
  if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var * unbias;
}

// Same as bn_finalize_kernel, starting from the per-CTA partial sums a convolution epilogue wrote
// (partial[part][2][C]): 32 channels x 8 part-lanes per block, lanes and then parts summed in a fixed order.
__global__ void __launch_bounds__(256) bn_finalize_partials_kernel(
    const float* __restrict__ partial, int nparts, const float* __restrict__ gamma, const float* __restrict__ beta,
    int C, float invM, float unbias, float eps, float momentum, float* running_mean, float* running_var,
    long long* nbt, float* sums_out, float* mean, float* rstd, float* scale, float* shift) {
  griddep_wait();
  griddep_launch();
  __shared__ float sm[2][8][33];
  const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float s0 = 0.0f, s1 = 0.0f;
  if (c < C) {
    for (int b = g; b < nparts; b += 8) {
      s0 += partial[(static_cast<size_t>(b) * 2 + 0) * C + c];
      s1 += partial[(static_cast<size_t>(b) * 2 + 1) * C + c];
    }
  }
  sm[0][g][cx] = s0;
  sm[1][g][cx] = s1;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0 && nbt) *nbt += 1;
  if (g != 0 || c >= C) return;
  float t0 = sm[0][0][cx], t1 = sm[1][0][cx];
#pragma unroll
  for (int l = 1; l < 8; ++l) {
    t0 += sm[0][l][cx];
    t1 += sm[1][l][cx];
  }
  if (sums_out) {
    sums_out[c] = t0;
    sums_out[C + c] = t1;
  }
  const float m = t0 * invM;
  const float var = fmaxf(t1 * invM - m * m, 0.0f);
  const float r = rsqrtf(var + eps);
  mean[c] = m;
  rstd[c] = r;
  const float sc = gamma[c] * r;
  scale[c] = sc;
  shift[c] = beta[c] - m * sc;
  if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * m;
  if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var * unbias;
}

// dgamma (+)= S(du*xhat); dbeta (+)= S(du)
__global__ void bn_param_grads_kernel(const float* __restrict__ sums, float* dgamma, float* dbeta, int C,
                                      float acc_gamma, float acc_beta) {
  griddep_wait();
  griddep_launch();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dgamma[c] = (acc_gamma != 0.0f ? acc_gamma * dgamma[c] : 0.0f) + sums[C + c];
  dbeta[c] = (acc_beta != 0.0f ? acc_beta * dbeta[c] : 0.0f) + sums[c];
}
// dgamma (+)= r/M*(M*q3 - q1*s1 - q2*s2)
__global__ void bn_gp_dgamma_kernel(const float* __restrict__ s, const float* __restrict__ q,
                                    const float* __restrict__ rstd, float* dgamma, int C, float invM, float acc) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float v = rstd[c] * (q[2 * C + c] - (q[c] * s[c] + q[C + c] * s[C + c]) * invM);
  dgamma[c] = (acc != 0.0f ? acc * dgamma[c] : 0.0f) + v;
}
__global__ void vec_axpby_kernel(const float* __restrict__ x, float* y, int n, float a, float b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (b != 0.0f ? b * y[i] : 0.0f) + a * x[i];
}

// latent = standardise_0(noise + z) with unbiased std (src/wgan_loss.py:105-106); one thread per feature column
// 32 feature columns x 8 row lanes per block: each lane walks rows lane, lane+8, ... (coalesced 128-byte rows), the
// lanes are combined in shared memory in a fixed order.  Three passes (mean, unbiased variance, write) like the
// reference's two-pass formula; large synthesis batches (B = 1024) no longer run one thread per column.
__global__ void __launch_bounds__(256) latent_prep_kernel(const float* __restrict__ noise, const float* __restrict__ z,
                                                          int B, int E, int zB, __nv_bfloat16* __restrict__ lat_bf16,
                                                          float* __restrict__ lat_f32) {
  __shared__ float sm[8][33];
  const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + cx;
  const bool ok = e < E;
  auto val = [&](int b) { return noise[static_cast<size_t>(b) * E + e] + z[static_cast<size_t>(zB == 1 ? 0 : b) * E + e]; };
  float s = 0.0f;
  if (ok) for (int b = g; b < B; b += 8) s += val(b);
  sm[g][cx] = s;
  __syncthreads();
  float tot = 0.0f;
#pragma unroll
  for (int l = 0; l < 8; ++l) tot += sm[l][cx];
  const float m = tot / B;
  __syncthreads();
  float ss = 0.0f;
  if (ok) for (int b = g; b < B; b += 8) { const float d = val(b) - m; ss += d * d; }
  sm[g][cx] = ss;
  __syncthreads();
  float tss = 0.0f;
#pragma unroll
  for (int l = 0; l < 8; ++l) tss += sm[l][cx];
  const float sd = sqrtf(tss / (B - 1));   // B == 1 -> NaN, like the reference
  if (!ok) return;
  for (int b = g; b < B; b += 8) {
    const float v = (val(b) - m) / sd;
    if (lat_bf16) lat_bf16[static_cast<size_t>(b) * E + e] = __float2bfloat16(v);
    if (lat_f32) lat_f32[static_cast<size_t>(b) * E + e] = v;
  }
}
__global__ void __launch_bounds__(256) im2col_img_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         int mode, const float* __restrict__ eps_dev,
                                                         const float* __restrict__ mul_dev, int B, int Cimg, int S,
                                                         int R, __nv_bfloat16* __restrict__ col,
                                                         float* __restrict__ mixed_out) {
  // One block stages the 2R + 2 input rows that R consecutive output rows need (1 + 1/R reads per image row instead
  // of 2) and emits R x (S/2) col rows of 128 bytes.
  extern __shared__ float rows[];   // [Cimg][2R + 2][S + 2], one zero column on each side
  const int Ho = S / 2, Wp = S + 2, NR = 2 * R + 2;
  const int groups = Ho / R;                    // R divides Ho (host)
  const int b = blockIdx.x / groups, ho0 = (blockIdx.x - b * groups) * R;
  const float eps = eps_dev ? __ldg(eps_dev) : 0.0f;
  const float mul = mul_dev ? __ldg(mul_dev) : 1.0f;
  const int ls = 31 - __clz(S);                 // S is a power of two (checked by the host wrapper)
  for (int idx = threadIdx.x; idx < Cimg * NR * S; idx += blockDim.x) {
    const int xx = idx & (S - 1);
    const int cr = idx >> ls;
    const int c = cr / NR, r = cr - c * NR;
    const int yy = 2 * ho0 - 1 + r;
    float t = 0.0f;
    if (yy >= 0 && yy < S) {
      const size_t o = ((static_cast<size_t>(b) * Cimg + c) * S + yy) * S + xx;
      t = __ldg(x + o);
      if (mode == 1) t = eps * t + (1.0f - eps) * __ldg(y + o);
      else if (mode == 2) { const float th = __ldg(y + o); t = t * (1.0f - th * th); }
      t *= mul;
      if (mixed_out && r >= 1 && r <= 2 * R) mixed_out[o] = t;   // rows 2ho0 .. 2ho0 + 2R - 1 are owned by this block
    }
    rows[(c * NR + r) * Wp + xx + 1] = t;
  }
  for (int idx = threadIdx.x; idx < Cimg * NR * 2; idx += blockDim.x)
    rows[(idx >> 1) * Wp + ((idx & 1) ? S + 1 : 0)] = 0.0f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * Ho * 8; idx += blockDim.x) {
    const int part = idx & 7;
    const int pw = idx >> 3;
    const int lr = pw / Ho, wo = pw - lr * Ho;
    const int kh = part >> 1, kw0 = (part & 1) * 2;
    uint32_t packed[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const int sx = 2 * wo + kw0 + k;            // smem column of input x = 2*wo - 1 + kw
      for (int c = 0; c < Cimg; ++c) v[c] = rows[(c * NR + 2 * lr + kh) * Wp + sx];
      packed[k * 2] = pack_bf16x2_ops(v[0], v[1]);
      packed[k * 2 + 1] = pack_bf16x2_ops(v[2], v[3]);
    }
    const size_t pix = (static_cast<size_t>(b) * Ho + ho0 + lr) * Ho + wo;
    *reinterpret_cast<uint4*>(col + pix * 64 + part * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

// per-channel sum over pixels of x (mode 0) or x*(1-y^2) (mode 2): bias gradient of the generator's last layer.
// stage 1: grid (nblk, Cimg) -> partial[c][blk]; stage 2: one block per channel, fixed-order tree (deterministic).
__global__ void __launch_bounds__(256) img_channel_sum_stage1(const float* __restrict__ x, const float* __restrict__ y,
                                                              int mode, int B, int Cimg, int S,
                                                              float* __restrict__ partial) {
  __shared__ float sm[256];
  const int c = blockIdx.y;
  const size_t plane = static_cast<size_t>(S) * S;
  const size_t total = static_cast<size_t>(B) * plane;
  float s = 0.0f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / plane;
    const size_t o = (b * Cimg + c) * plane + (i - b * plane);
    float t = x[o];
    if (mode == 2) { const float th = y[o]; t = t * (1.0f - th * th); }
    s += t;
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[c * gridDim.x + blockIdx.x] = sm[0];
}
__global__ void __launch_bounds__(256) img_channel_sum_stage2(const float* __restrict__ partial, int nblk,
                                                              float* __restrict__ out, float acc) {
  __shared__ float sm[256];
  const int c = blockIdx.x;
  float s = 0.0f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) s += partial[c * nblk + i];
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[c] = (acc != 0.0f ? acc * out[c] : 0.0f) + sm[0];
}

// col2im for the image-side transposed conv computed in "dgrad form": col[pix][n = tap*Cimg + c] (fp32) holds each
// low-res pixel's contribution to its 4x4 output patch; out[b][c][y][x] = act(bias[c] + sum of the 4 taps that hit it).
__global__ void __launch_bounds__(256) col2im_img_kernel(const float* __restrict__ col, int ldc,
                                                         const float* __restrict__ bias, int act_tanh, int B, int Cimg,
                                                         int H, int W, float* __restrict__ img) {
  const int OH = 2 * H, OW = 2 * W;
  const unsigned n = static_cast<unsigned>(B) * OH * OW;
  const int lw = 31 - __clz(OW), lh = 31 - __clz(OH);    // power-of-two image sides (checked by the host wrapper)
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int x = static_cast<int>(i & (OW - 1));
    const int y = static_cast<int>((i >> lw) & (OH - 1));
    const int b = static_cast<int>(i >> (lw + lh));
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    // y = 2*ii - 1 + kh  ->  kh has the parity of (y + 1); two candidates each for rows and columns
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kh = ((y + 1) & 1) + 2 * a;
      const int ii = (y + 1 - kh) >> 1;
      if (ii < 0 || ii >= H) continue;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int kw = ((x + 1) & 1) + 2 * e;
        const int jj = (x + 1 - kw) >> 1;
        if (jj < 0 || jj >= W) continue;
        const float* src = col + ((static_cast<size_t>(b) * H + ii) * W + jj) * ldc + (kh * 4 + kw) * Cimg;
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < Cimg) acc[c] += __ldg(src + c);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (c < Cimg) {
        float v = acc[c] + (bias ? __ldg(bias + c) : 0.0f);
        if (act_tanh & 1) v = tanhf(v);
        if (act_tanh & 4) {   // uint8 NHWC tile as the reference writes it to disk: trunc(255 * (x + 1) / 2)
                              // (src/generate_tissue_images.py:127-129; bit 3: channel order reversed, cv2's BGR)
          const float u = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), 255.0f);
          const int cc = (act_tanh & 8) ? Cimg - 1 - c : c;
          reinterpret_cast<uint8_t*>(img)[((static_cast<size_t>(b) * OH + y) * OW + x) * Cimg + cc] =
              static_cast<uint8_t>(__float2uint_rz(fminf(fmaxf(u, 0.0f), 255.0f)));
        } else if (act_tanh & 2)     // synthesis output (src/gan_utils.py:236-241): (x + 1) / 2, NHWC
          img[((static_cast<size_t>(b) * OH + y) * OW + x) * Cimg + c] = (v + 1.0f) * 0.5f;
        else
          img[((static_cast<size_t>(b) * Cimg + c) * OH + y) * OW + x] = v;
      }
    }
  }
}
// W[Cp][Cimg][16] -> w_colT[n = tap*Cimg + c][Cp] bf16 (rows >= 16*Cimg zero): B operand of the dgrad-form GEMM
__global__ void pack_edge_t_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int Cp, int Cimg,
                                   int rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * Cp) return;
  const int n = idx / Cp, p = idx - n * Cp;
  float v = 0.0f;
  if (n < 16 * Cimg) {
    const int tap = n / Cimg, c = n - tap * Cimg;
    v = W[(static_cast<size_t>(p) * Cimg + c) * 16 + tap];
  }
  out[idx] = __float2bfloat16(v);
}

// dW[p][c][kh][kw] (fp32 torch layout) (+)= dWcol[p][k = tap*4 + c]
__global__ void unpack_edge_grad_kernel(const float* __restrict__ dcol, float* dW, int Cp, int Cimg, float acc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Cp * Cimg * 16) return;
  const int tap = i & 15, c = (i >> 4) % Cimg, p = i / (16 * Cimg);
  dW[i] = (acc != 0.0f ? acc * dW[i] : 0.0f) + dcol[p * 64 + tap * 4 + c];
}

// critic head: a6[b] = dot(h5[b,:], w[:]); out[b] = lrelu(a6[b]).  One block per sample.
__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ h5, const float* __restrict__ w, int K,
                                float slope, float* __restrict__ a6, float* __restrict__ out) {
  __shared__ float sm[256];
  const int b = blockIdx.x;
  const __nv_bfloat16* x = h5 + static_cast<size_t>(b) * K;
  float s = 0.0f;
  for (int k = threadIdx.x * 8; k < K; k += blockDim.x * 8) {
    const Vec8 v = ld8(x + k);
    const Vec8 ww = ldf8(w + k);
#pragma unroll
    for (int e = 0; e < 8; ++e) s = fmaf(v.v[e], ww.v[e], s);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int wd = 128; wd > 0; wd >>= 1) {
    if (threadIdx.x < wd) sm[threadIdx.x] += sm[threadIdx.x + wd];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float a = sm[0];
    a6[b] = a;
    out[b] = a > 0.0f ? a : a * slope;
  }
}
// da6[b] = dout[b] * lrelu'(a6[b]);  dh5[b][k] = da6[b] * w[k]
__global__ void head_bwd_data_kernel(const float* __restrict__ a6, const float* __restrict__ dout, float dout_const,
                                     const float* __restrict__ w, int B, int K, float slope, float* __restrict__ da6,
                                     __nv_bfloat16* __restrict__ dh5) {
  const size_t nvec = static_cast<size_t>(B) * (K / 8);
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / (K / 8));
    const int k = static_cast<int>(i - static_cast<size_t>(b) * (K / 8)) * 8;
    const float go = dout ? dout[b] : dout_const;
    const float d = a6[b] > 0.0f ? go : go * slope;
    if (k == 0 && da6) da6[b] = d;
    const Vec8 ww = ldf8(w + k);
    Vec8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) o.v[e] = d * ww.v[e];
    st8(dh5 + static_cast<size_t>(b) * K + k, o);
  }
}
// dw[k] (+)= sum_b da6[b] * x[b][k]   (x = h5 or the adjoint A_dh5); writes the torch layout [1][C][4][4]
__global__ void head_wgrad_kernel(const float* __restrict__ da6, const __nv_bfloat16* __restrict__ x, int B, int K,
                                  int C, float* dW, float acc) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.0f;
  for (int b = 0; b < B; ++b) s = fmaf(da6[b], __bfloat162float(x[static_cast<size_t>(b) * K + k]), s);
  const int tap = k / C, c = k - tap * C;
  const int o = c * 16 + tap;
  dW[o] = (acc != 0.0f ? acc * dW[o] : 0.0f) + s;
}
// w_head[k = tap*C + c] (fp32) = W[0][c][tap]
__global__ void pack_head_kernel(const float* __restrict__ W, float* __restrict__ w, int C) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 16 * C) return;
  const int tap = k / C, c = k - tap * C;
  w[k] = W[c * 16 + tap];
}

// WGAN losses (src/wgan_loss.py:24-29): loss_out[0] = mean(sign_a * a) + mean(sign_b * b) (b optional)
__global__ void wgan_loss_kernel(const float* __restrict__ a, float sign_a, const float* __restrict__ b, float sign_b,
                                 int B, float* __restrict__ loss_out) {
  __shared__ float sm[256];
  float s = 0.0f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += sign_a * a[i] + (b ? sign_b * b[i] : 0.0f);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = sm[0] / B;
}

// sum of squares of a fp32 buffer -> partial[block]; then finalize the gradient penalty:
//   norm = sqrt(sum); P = (norm - 1)^2 (src/wgan_loss.py:43); seed scale = lambda * 2 * (norm - 1) / norm
__global__ void sumsq_stage1_kernel(const float* __restrict__ x, size_t n, float* __restrict__ partial) {
  __shared__ float sm[256];
  float s = 0.0f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    s = fmaf(v, v, s);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void gp_finalize_kernel(const float* __restrict__ partial, int n, float lambd, float* __restrict__ out) {
  // out[0] = penalty, out[1] = seed scale, out[2] = norm
  __shared__ double sm[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += static_cast<double>(partial[i]);
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = sqrtf(static_cast<float>(sm[0]));
    out[0] = (norm - 1.0f) * (norm - 1.0f);
    // torch.norm's backward gives the zero subgradient at norm == 0 (dead critic / all-zero input gradient)
    out[1] = norm > 0.0f ? lambd * 2.0f * (norm - 1.0f) / norm : 0.0f;
    out[2] = norm;
  }
}

// ------------------------------------------------------------------------------------------------ Adam
struct AdamChunk {
  float* p;
  const float* g;
  float* m;
  float* v;
  __nv_bfloat16* shadow;   // optional bf16 copy of the updated parameters in the same element order (the packed GEMM
                           // operand of a channels_last weight): the optimiser step re-emits it, no pack kernel
  int n;
  // pitched shadow (sh_cols > 0): the parameter is a dense [rows][sh_cols] matrix whose bf16 copy has rows of sh_pitch
  // elements (K padded for the 16-byte TMA stride rule, e.g. 19198 -> 19200 genes); `shadow` then points at the start of
  // the shadow row that holds this chunk's first element, which sits at column sh_col0 of it
  int sh_cols, sh_pitch, sh_col0;
};
// torch.optim.Adam (no amsgrad, no weight decay), src/histopathology_gan.py:252,257:
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// dyn (optional, device): {learning rate, step count as float} read at run time instead of the by-value arguments, so a
// CUDA graph that contains this launch stays valid from step to step (adam_tick_kernel advances the count)
__global__ void adam_tick_kernel(float* dyn) { dyn[1] += 1.0f; }
__global__ void __launch_bounds__(256) adam_kernel(const AdamChunk* __restrict__ chunks, float lr, float b1, float b2,
                                                   float eps, float bc1, float bc2_sqrt, float clamp_lo, float clamp_hi,
                                                   int do_clamp, float gscale, const float* __restrict__ dyn) {
  if (dyn != nullptr) {
    lr = dyn[0];
    const float t = dyn[1];
    bc1 = 1.0f - powf(b1, t);
    bc2_sqrt = sqrtf(1.0f - powf(b2, t));
  }
  const AdamChunk ch = chunks[blockIdx.x];
  const float step = lr / bc1;
  const float ob1 = 1.0f - b1, ob2 = 1.0f - b2;
  const int n4 = ((reinterpret_cast<uintptr_t>(ch.p) | reinterpret_cast<uintptr_t>(ch.g) |
                   reinterpret_cast<uintptr_t>(ch.m) | reinterpret_cast<uintptr_t>(ch.v)) & 15) == 0 ? (ch.n >> 2) : 0;
  float4* p4 = reinterpret_cast<float4*>(ch.p);
  const float4* g4 = reinterpret_cast<const float4*>(ch.g);
  float4* m4 = reinterpret_cast<float4*>(ch.m);
  float4* v4 = reinterpret_cast<float4*>(ch.v);
  const bool pitched = ch.shadow != nullptr && ch.sh_cols > 0;
  uint2* s4 = (ch.shadow != nullptr && !pitched && (reinterpret_cast<uintptr_t>(ch.shadow) & 7) == 0)
                  ? reinterpret_cast<uint2*>(ch.shadow) : nullptr;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    float4 p = p4[i], m = m4[i], v = v4[i];
    float4 g = g4[i];
    g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale;
    float* pp = reinterpret_cast<float*>(&p);
    float* mm = reinterpret_cast<float*>(&m);
    float* vv = reinterpret_cast<float*>(&v);
    const float* gg = reinterpret_cast<const float*>(&g);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mm[e] = b1 * mm[e] + ob1 * gg[e];
      vv[e] = b2 * vv[e] + ob2 * gg[e] * gg[e];
      float q = pp[e] - step * (mm[e] / (sqrtf(vv[e]) / bc2_sqrt + eps));
      if (do_clamp) q = fminf(fmaxf(q, clamp_lo), clamp_hi);
      pp[e] = q;
    }
    p4[i] = p; m4[i] = m; v4[i] = v;
    if (s4) s4[i] = make_uint2(pack_bf16x2_ops(p.x, p.y), pack_bf16x2_ops(p.z, p.w));
    if (pitched) {
      unsigned q = static_cast<unsigned>(ch.sh_col0) + 4u * static_cast<unsigned>(i);
      unsigned r = q / static_cast<unsigned>(ch.sh_cols);
      unsigned c = q - r * static_cast<unsigned>(ch.sh_cols);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ch.shadow[static_cast<size_t>(r) * ch.sh_pitch + c] = __float2bfloat16(pp[e]);
        if (++c == static_cast<unsigned>(ch.sh_cols)) { c = 0; ++r; }
      }
    }
  }
  if (ch.shadow != nullptr && s4 == nullptr && !pitched) {     // shadow not 8-byte aligned (offset views through the C ABI)
    __syncthreads();                                // the elements below were written by other threads' float4 stores
    for (int i = threadIdx.x; i < n4 * 4; i += blockDim.x) ch.shadow[i] = __float2bfloat16(ch.p[i]);
  }
  for (int i = n4 * 4 + threadIdx.x; i < ch.n; i += blockDim.x) {
    const float g = ch.g[i] * gscale;
    const float m = b1 * ch.m[i] + ob1 * g;
    const float v = b2 * ch.v[i] + ob2 * g * g;
    ch.m[i] = m;
    ch.v[i] = v;
    float p = ch.p[i] - step * (m / (sqrtf(v) / bc2_sqrt + eps));
    if (do_clamp) p = fminf(fmaxf(p, clamp_lo), clamp_hi);
    ch.p[i] = p;
    if (pitched) {
      const unsigned q = static_cast<unsigned>(ch.sh_col0) + static_cast<unsigned>(i);
      const unsigned r = q / static_cast<unsigned>(ch.sh_cols);
      ch.shadow[static_cast<size_t>(r) * ch.sh_pitch + (q - r * static_cast<unsigned>(ch.sh_cols))] = __float2bfloat16(p);
    } else if (ch.shadow) {
      ch.shadow[i] = __float2bfloat16(p);
    }
  }
}
__global__ void clamp_kernel(float* p, size_t n, float lo, float hi) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    p[i] = fminf(fmaxf(p[i], lo), hi);
}
// out[i] = sum_{r < nparts} src[r * stride + i], r ascending (fixed order: every rank of a data-parallel job reduces a
// slice with the same association, and a rerun reproduces it bit for bit).  float4 lanes, up to 8 parts in flight.
__global__ void __launch_bounds__(256) slices_sum_kernel(const float* __restrict__ src, int nparts, size_t stride, size_t n4,
                                                         float* __restrict__ out) {
  const size_t step = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += step) {
    float4 acc = reinterpret_cast<const float4*>(src)[i];
    for (int r0 = 1; r0 < nparts; r0 += 8) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < nparts) v[j] = reinterpret_cast<const float4*>(src + static_cast<size_t>(r0 + j) * stride)[i];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < nparts) {
          acc.x += v[j].x;
          acc.y += v[j].y;
          acc.z += v[j].z;
          acc.w += v[j].w;
        }
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}
// In-switch all-reduce of this rank's slice of a symmetric buffer through its MULTICAST address (NVLink SHARP):
// multimem.ld_reduce returns the fp32 sum over every GPU's copy (reduced inside the NVSwitch), multimem.st writes the
// result back into every GPU's copy.  Each element is reduced exactly once (by the rank that owns its slice) and then
// broadcast, so all ranks end with bit-identical values.  Needs a barrier before (all contributions written) and
// after (all stores landed); a handful of CTAs saturates the link, the rest of the GPU keeps computing.
// CTAs of 128 threads and <= 64 registers: small enough to CO-RESIDE with a persistent tile-engine CTA (224 threads x 236
// registers, ~200 KiB of shared memory) on the same SM, so the exchange does not displace GEMM CTAs into a second wave.
__global__ void __launch_bounds__(128, 4) nvls_allreduce_kernel(float* mc, size_t n4) {
  // a switch round trip costs microseconds: keep kU independent 16-byte reductions in flight per thread
  constexpr int kU = 8;
  const size_t step = static_cast<size_t>(gridDim.x) * blockDim.x;
  size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (kU - 1) * step < n4; i += kU * step) {
    float4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(mc + (i + u * step) * 4) : "memory");
#pragma unroll
    for (int u = 0; u < kU; ++u)
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                   :: "l"(mc + (i + u * step) * 4), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
  }
  for (; i < n4; i += step) {
    float* p = mc + i * 4;
    float x, y, z, w;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "l"(p) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
  }
}
__global__ void cast_f32_kernel(const float* __restrict__ src, float* __restrict__ dst32, __nv_bfloat16* dst16, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (dst32) dst32[i] = src[i];
    if (dst16) dst16[i] = __float2bfloat16(src[i]);
  }
}
// uint8 HWC tiles (as stored in the patch LMDBs, src/read_data.py:336-343) -> fp32 NCHW in [-1, 1]:
// cv2.cvtColor(BGR2RGB) + permute(2,0,1) + ConvertImageDtype(float) + Normalize(0.5, 0.5)
// (src/read_data.py:341-343, src/histopathology_gan.py:106-109) in one pass, same fp32 operations (x / 255, then
// (x - 0.5) / 0.5) so the result is bit-identical to the reference's CPU transforms.  One thread per 4 pixels of a row.
__global__ void __launch_bounds__(256) tiles_u8_to_nchw_kernel(const uint8_t* __restrict__ tiles, float* __restrict__ img,
                                                               int B, int C, int S, int swap_rb) {
  const size_t nquad = static_cast<size_t>(B) * S * (S / 4);
  for (size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; q < nquad;
       q += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int xq = static_cast<int>(q % (S / 4));
    const size_t by = q / (S / 4);                 // b * S + y
    const size_t b = by / S;
    const int y = static_cast<int>(by - b * S);
    const uint8_t* src = tiles + (by * S + static_cast<size_t>(xq) * 4) * C;
    for (int c = 0; c < C; ++c) {
      const int cs = (swap_rb && C >= 3 && c < 3) ? 2 - c : c;
      float4 o;
      o.x = (__fdiv_rn(static_cast<float>(src[0 * C + cs]), 255.0f) - 0.5f) / 0.5f;
      o.y = (__fdiv_rn(static_cast<float>(src[1 * C + cs]), 255.0f) - 0.5f) / 0.5f;
      o.z = (__fdiv_rn(static_cast<float>(src[2 * C + cs]), 255.0f) - 0.5f) / 0.5f;
      o.w = (__fdiv_rn(static_cast<float>(src[3 * C + cs]), 255.0f) - 0.5f) / 0.5f;
      *reinterpret_cast<float4*>(img + ((b * C + c) * S + y) * S + static_cast<size_t>(xq) * 4) = o;
    }
  }
}
__global__ void nchw_to_unit_nhwc_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int C, int S) {
  // (x+1)/2 and NCHW -> NHWC (src/gan_utils.py:236-241)
  const size_t n = static_cast<size_t>(B) * C * S * S;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const size_t pix = i / C;
    const size_t b = pix / (static_cast<size_t>(S) * S);
    const size_t yx = pix - b * S * S;
    out[i] = (img[(b * C + c) * S * S + yx] + 1.0f) * 0.5f;
  }
}


// ------------------------------------------------------------------------------------------------ betaVAE training
// dst[r][c] = bf16(src[r][c] * mul[r][c] * scale) for c < cols, 0 for the pad columns (Dropout(0.5) in train mode,
// src/betaVAE.py:26-27, with the caller's mask; mul == nullptr -> plain cast)
__global__ void mul_cast_pad_kernel(const float* __restrict__ src, const float* __restrict__ mul, float scale,
                                    __nv_bfloat16* __restrict__ dst, int rows, int cols, int cols_pad) {
  const size_t n = static_cast<size_t>(rows) * cols_pad;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / cols_pad;
    const int c = static_cast<int>(i - r * cols_pad);
    float v = 0.0f;
    if (c < cols) {
      v = src[r * cols + c] * scale;
      if (mul) v *= mul[r * cols + c];
    }
    dst[i] = __float2bfloat16(v);
  }
}
__device__ __forceinline__ float block_sum_256(float s, float* sm) {
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) sm[threadIdx.x] += sm[threadIdx.x + w];
    __syncthreads();
  }
  return sm[0];
}
// z = mu + eps*exp(lv/2) (src/betaVAE.py:96-100); partial[block] = sum(1 + lv - mu^2 - exp(lv)) (src/betaVAE.py:148)
__global__ void __launch_bounds__(256) vae_reparam_kernel(const float* __restrict__ mulv, const float* __restrict__ eps,
                                                          int B, int Z, __nv_bfloat16* __restrict__ z,
                                                          float* __restrict__ partial) {
  __shared__ float sm[256];
  const size_t n = static_cast<size_t>(B) * Z;
  float s = 0.0f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / Z;
    const int j = static_cast<int>(i - b * Z);
    const float mu = mulv[b * 2 * Z + j], lv = mulv[b * 2 * Z + Z + j];
    const float e = expf(lv);
    z[i] = __float2bfloat16(mu + eps[i] * expf(0.5f * lv));
    s += 1.0f + lv - mu * mu - e;
  }
  const float t = block_sum_256(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// out = tanh(pre); d_pre = gscale*(out - x)*(1 - out^2) (bf16, pad columns zero); partial[block] = sum (out - x)^2
__global__ void __launch_bounds__(256) vae_recon_kernel(const float* __restrict__ pre, int ldp,
                                                        const float* __restrict__ x, int B, int F, float gscale,
                                                        __nv_bfloat16* __restrict__ dpre, float* __restrict__ partial) {
  __shared__ float sm[256];
  const size_t n = static_cast<size_t>(B) * ldp;
  float s = 0.0f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / ldp;
    const int j = static_cast<int>(i - b * ldp);
    float d = 0.0f;
    if (j < F) {
      const float o = tanhf(pre[i]);
      const float e = o - x[b * F + j];
      s += e * e;
      d = gscale * e * (1.0f - o * o);
    }
    dpre[i] = __float2bfloat16(d);
  }
  const float t = block_sum_256(s, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
// d_mu = dz + kscale*mu ; d_lv = dz*eps*0.5*exp(lv/2) - 0.5*kscale*(1 - exp(lv)), kscale = beta/B
__global__ void vae_latent_grad_kernel(const __nv_bfloat16* __restrict__ dz, const float* __restrict__ mulv,
                                       const float* __restrict__ eps, int B, int Z, float kscale,
                                       __nv_bfloat16* __restrict__ dcat) {
  const size_t n = static_cast<size_t>(B) * Z;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / Z;
    const int j = static_cast<int>(i - b * Z);
    const float mu = mulv[b * 2 * Z + j], lv = mulv[b * 2 * Z + Z + j];
    const float g = __bfloat162float(dz[i]);
    dcat[b * 2 * Z + j] = __float2bfloat16(g + kscale * mu);
    dcat[b * 2 * Z + Z + j] = __float2bfloat16(g * eps[i] * 0.5f * expf(0.5f * lv) - 0.5f * kscale * (1.0f - expf(lv)));
  }
}
// out3 = {total, recon, kld}: recon = sse/(B*F); kld = -0.5*sum/B; total = recon + beta*kld (src/betaVAE.py:145-162)
__global__ void __launch_bounds__(256) vae_loss_finalize_kernel(const float* __restrict__ p_sse, int n1,
                                                                const float* __restrict__ p_kld, int n2, float inv_bf,
                                                                float inv_b, float beta, float* __restrict__ out3) {
  __shared__ float sm[256];
  float a = 0.0f, b = 0.0f;
  for (int i = threadIdx.x; i < n1; i += blockDim.x) a += p_sse[i];
  for (int i = threadIdx.x; i < n2; i += blockDim.x) b += p_kld[i];
  const float sse = block_sum_256(a, sm);
  __syncthreads();
  const float ks = block_sum_256(b, sm);
  if (threadIdx.x == 0) {
    const float recon = sse * inv_bf, kld = -0.5f * ks * inv_b;
    out3[0] = recon + beta * kld;
    out3[1] = recon;
    out3[2] = kld;
  }
}


// ------------------------------------------------------------------------------------------------ resize-conv generator
// nn.Upsample(scale_factor=2, mode='bilinear') [align_corners=False] followed by nn.ReflectionPad2d(1)
// (src/dcgan.py:48-49,78-79; index rules of SURVEY.md Appendix B.11), bf16 NHWC [B,H,W,C] -> [B,2H+2,2W+2,C].
__device__ __forceinline__ void up_taps(int d, int n, int& i0, int& i1, float& w1) {
  // source coordinate of up-sampled index d: max(0, (d + 0.5)/2 - 0.5); i1 clamped to n-1
  const float src = fmaxf(0.0f, (d + 0.5f) * 0.5f - 0.5f);
  i0 = static_cast<int>(src);
  w1 = src - i0;
  i1 = min(i0 + 1, n - 1);
}
__device__ __forceinline__ int reflect_idx(int p, int n) {   // padded index p in [-1, n] -> [0, n)
  return p < 0 ? -p : (p >= n ? 2 * n - 2 - p : p);
}
struct UpPadF {
  const __nv_bfloat16* h;
  __nv_bfloat16* u;
  int H, W, C;
  __device__ void operator()(size_t row, int c0) const {
    const int Wp = 2 * W + 2, Hp = 2 * H + 2;
    const int px = static_cast<int>(row % Wp);
    const int py = static_cast<int>((row / Wp) % Hp);
    const size_t b = row / (static_cast<size_t>(Wp) * Hp);
    const int yy = reflect_idx(py - 1, 2 * H), xx = reflect_idx(px - 1, 2 * W);
    int y0, y1, x0, x1;
    float wy, wx;
    up_taps(yy, H, y0, y1, wy);
    up_taps(xx, W, x0, x1, wx);
    const __nv_bfloat16* base = h + b * H * W * C + c0;
    const Vec8 a = ld8(base + (static_cast<size_t>(y0) * W + x0) * C), bq = ld8(base + (static_cast<size_t>(y0) * W + x1) * C);
    const Vec8 c = ld8(base + (static_cast<size_t>(y1) * W + x0) * C), d = ld8(base + (static_cast<size_t>(y1) * W + x1) * C);
    Vec8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float top = a.v[e] + wx * (bq.v[e] - a.v[e]);
      const float bot = c.v[e] + wx * (d.v[e] - c.v[e]);
      o.v[e] = top + wy * (bot - top);
    }
    st8(u + row * C + c0, o);
  }
};
// adjoint: dh[b,i,j,:] = sum over the padded up-sampled pixels that read (i,j), weights as in the forward
struct UpPadBwdF {
  const __nv_bfloat16* du;
  __nv_bfloat16* dh;
  int H, W, C;
  // gradient of the un-padded up-sampled pixel (yy,xx): its own padded position plus the reflected border copies
  __device__ __forceinline__ void gather(size_t b, int yy, int xx, int c0, float wgt, Vec8& acc) const {
    const int Wp = 2 * W + 2, Hp = 2 * H + 2, H2 = 2 * H, W2 = 2 * W;
    int ys[2], xs[2], ny = 1, nx = 1;
    ys[0] = yy + 1; xs[0] = xx + 1;
    if (yy == 1) ys[ny++] = 0; else if (yy == H2 - 2) ys[ny++] = Hp - 1;
    if (xx == 1) xs[nx++] = 0; else if (xx == W2 - 2) xs[nx++] = Wp - 1;
    for (int a = 0; a < ny; ++a)
      for (int q = 0; q < nx; ++q) {
        const Vec8 v = ld8(du + ((b * Hp + ys[a]) * Wp + xs[q]) * C + c0);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc.v[e] += wgt * v.v[e];
      }
  }
  __device__ void operator()(size_t row, int c0) const {
    const int j = static_cast<int>(row % W);
    const int i = static_cast<int>((row / W) % H);
    const size_t b = row / (static_cast<size_t>(W) * H);
    Vec8 acc;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc.v[e] = 0.0f;
    // up-sampled rows yy in [2i-1, 2i+2] can read input row i; recompute their taps to get the exact weights
    for (int yy = max(0, 2 * i - 1); yy <= min(2 * H - 1, 2 * i + 2); ++yy) {
      int y0, y1; float wy;
      up_taps(yy, H, y0, y1, wy);
      const float cy = (y0 == i ? 1.0f - wy : 0.0f) + (y1 == i ? wy : 0.0f);
      if (cy == 0.0f) continue;
      for (int xx = max(0, 2 * j - 1); xx <= min(2 * W - 1, 2 * j + 2); ++xx) {
        int x0, x1; float wx;
        up_taps(xx, W, x0, x1, wx);
        const float cx = (x0 == j ? 1.0f - wx : 0.0f) + (x1 == j ? wx : 0.0f);
        if (cx == 0.0f) continue;
        gather(b, yy, xx, c0, cy * cx, acc);
      }
    }
    st8(dh + row * C + c0, acc);
  }
};

// last layer of the resize-conv generator (Conv2d(C,3,3) on the padded tensor): backward on CUDA cores.
// dW[co][ci][kh][kw] partials: one block per pixel chunk, thread = (ci) x (co,tap split); deterministic 2-stage.
__global__ void __launch_bounds__(256) upg_last_wgrad_stage1(const __nv_bfloat16* __restrict__ u,
                                                             const float* __restrict__ dout, int B, int S, int C,
                                                             int Cimg, int pix_per_block, float* __restrict__ partial) {
  // partial[block][co*9+tap][ci]; requires C == 64 (4 groups of 64 threads split the 9 taps)
  extern __shared__ float sd[];   // dout values of the current pixel batch: [Cimg][batch]
  const int ci = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int Sp = S + 2;
  const size_t npix = static_cast<size_t>(B) * S * S;
  const size_t p0 = static_cast<size_t>(blockIdx.x) * pix_per_block;
  const size_t p1 = min(npix, p0 + pix_per_block);
  float acc[3][3];    // [co][tap slot]: this group handles taps grp, grp+4, grp+8
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int t = 0; t < 3; ++t) acc[a][t] = 0.0f;
  for (size_t pb = p0; pb < p1; pb += 64) {
    const int nb = static_cast<int>((p1 - pb) < 64 ? (p1 - pb) : 64);
    __syncthreads();
    for (int e = threadIdx.x; e < Cimg * nb; e += blockDim.x) {
      const int co = e / nb, q = e - co * nb;
      const size_t pix = pb + q;
      const size_t b = pix / (static_cast<size_t>(S) * S);
      const size_t yx = pix - b * S * S;
      sd[co * 64 + q] = dout[(b * Cimg + co) * S * S + yx];
    }
    __syncthreads();
    for (int q = 0; q < nb; ++q) {
      const size_t pix = pb + q;
      const int x = static_cast<int>(pix % S), y = static_cast<int>((pix / S) % S);
      const size_t b = pix / (static_cast<size_t>(S) * S);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int tap = grp + 4 * t;
        if (tap < 9) {
          const int kh = tap / 3, kw = tap - kh * 3;
          const float uv = __bfloat162float(u[((b * Sp + y + kh) * Sp + x + kw) * C + ci]);
#pragma unroll
          for (int co = 0; co < 3; ++co)
            if (co < Cimg) acc[co][t] = fmaf(sd[co * 64 + q], uv, acc[co][t]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int tap = grp + 4 * t;
    if (tap < 9) {
#pragma unroll
      for (int co = 0; co < 3; ++co)
        if (co < Cimg) partial[(static_cast<size_t>(blockIdx.x) * (Cimg * 9) + co * 9 + tap) * C + ci] = acc[co][t];
    }
  }
}
__global__ void upg_last_wgrad_stage2(const float* __restrict__ partial, int nblocks, int C, int Cimg,
                                      float* __restrict__ dW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;    // over (co, tap, ci)
  if (i >= Cimg * 9 * C) return;
  float s = 0.0f;
  for (int b = 0; b < nblocks; ++b) s += partial[static_cast<size_t>(b) * Cimg * 9 * C + i];
  const int ci = i % C, ct = i / C, tap = ct % 9, co = ct / 9;
  dW[(static_cast<size_t>(co) * C + ci) * 9 + tap] = s;
}
// du[b,y',x',ci] = sum_{co,kh,kw} dout[b,co,y'-kh,x'-kw] * W[co][ci][kh][kw]  on the padded grid
__global__ void __launch_bounds__(256) upg_last_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ W,
                                                             int B, int S, int C, int Cimg,
                                                             __nv_bfloat16* __restrict__ du) {
  extern __shared__ float sw[];   // W as [co][tap][ci]
  for (int e = threadIdx.x; e < Cimg * 9 * C; e += blockDim.x) {
    const int ci = e % C, ct = e / C, tap = ct % 9, co = ct / 9;
    sw[e] = W[(static_cast<size_t>(co) * C + ci) * 9 + tap];
  }
  __syncthreads();
  const int Sp = S + 2, cgs = C / 8;
  const size_t nvec = static_cast<size_t>(B) * Sp * Sp * cgs;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % cgs);
    const size_t pix = i / cgs;
    const int px = static_cast<int>(pix % Sp), py = static_cast<int>((pix / Sp) % Sp);
    const size_t b = pix / (static_cast<size_t>(Sp) * Sp);
    Vec8 acc;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc.v[e] = 0.0f;
    for (int kh = 0; kh < 3; ++kh) {
      const int y = py - kh;
      if (y < 0 || y >= S) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int x = px - kw;
        if (x < 0 || x >= S) continue;
        for (int co = 0; co < Cimg; ++co) {
          const float g = __ldg(dout + ((b * Cimg + co) * S + y) * S + x);
          const float* wr = sw + (co * 9 + kh * 3 + kw) * C + cg * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc.v[e] = fmaf(g, wr[e], acc.v[e]);
        }
      }
    }
    st8(du + pix * C + cg * 8, acc);
  }
}

}  // namespace rg

using namespace rg;
typedef __nv_bfloat16 bf16;

extern "C" {

size_t rg_reduce_ws_bytes(int M, int C) { return reduce_ws_floats(M, C, 3) * sizeof(float); }

int rg_bn_stats(const void* a, int M, int C, void* ws, size_t ws_bytes, float* sums, rg_stream_t st) {
  StatsF f{static_cast<const bf16*>(a), C};
  return run_colreduce(f, M, C, static_cast<float*>(ws), ws_bytes, sums, static_cast<cudaStream_t>(st), "rg_bn_stats");
}

int rg_bn_finalize(const float* sums, const float* gamma, const float* beta, int M, int C, float eps, float momentum,
                   float* running_mean, float* running_var, int64_t* num_batches_tracked, float* mean, float* rstd,
                   float* scale, float* shift, rg_stream_t st) {
  RG_CHECK_ARG(sums && gamma && beta && mean && rstd && scale && shift && M > 0 && C > 0, "rg_bn_finalize: bad arguments");
  const float unbias = M > 1 ? static_cast<float>(M) / static_cast<float>(M - 1) : 1.0f;
  bn_finalize_kernel<<<ceil_div(C, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      sums, gamma, beta, C, 1.0f / M, unbias, eps, momentum, running_mean, running_var,
      reinterpret_cast<long long*>(num_batches_tracked), mean, rstd, scale, shift);
  RG_LAUNCH_CHECK("rg_bn_finalize");
  return 0;
}

int rg_bn_finalize_partials(const float* stats_ws, const float* gamma, const float* beta, int M, int C, float eps,
                            float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                            float* sums_out, float* mean, float* rstd, float* scale, float* shift, rg_stream_t st) {
  RG_CHECK_ARG(stats_ws && gamma && beta && mean && rstd && scale && shift && M > 0 && C > 0,
               "rg_bn_finalize_partials: bad arguments");
  const float unbias = M > 1 ? static_cast<float>(M) / static_cast<float>(M - 1) : 1.0f;
  RG_CUDA(launch_pdl(bn_finalize_partials_kernel, dim3(ceil_div(C, 32)), dim3(256), 0, static_cast<cudaStream_t>(st), 1,
                     stats_ws, num_sms(), gamma, beta, C, 1.0f / M, unbias, eps, momentum, running_mean, running_var,
                     reinterpret_cast<long long*>(num_batches_tracked), sums_out, mean, rstd, scale, shift));
  RG_LAUNCH_CHECK("rg_bn_finalize_partials");
  return 0;
}

int rg_reduce_partials(const float* stats_ws, int KC, float* out, rg_stream_t st) {
  RG_CHECK_ARG(stats_ws && out && KC > 0, "rg_reduce_partials: bad arguments");
  colreduce_stage2<<<ceil_div(KC, 32), 1024, 0, static_cast<cudaStream_t>(st)>>>(stats_ws, num_sms(), KC, out);
  RG_LAUNCH_CHECK("rg_reduce_partials");
  return 0;
}

int rg_bn_act(const void* a, const float* scale, const float* shift, float slope, void* h, int M, int C,
              rg_stream_t st) {
  RG_CHECK_ARG(a && scale && shift && h, "rg_bn_act: null pointer");
  BnActF f{static_cast<const bf16*>(a), static_cast<bf16*>(h), scale, shift, slope, C};
  return run_ew_split(f, M, C, static_cast<cudaStream_t>(st), "rg_bn_act");
}

int rg_bn_bwd_reduce(const void* dh, const void* a, const float* mean, const float* rstd, const float* scale,
                     const float* shift, float slope, int M, int C, void* ws, size_t ws_bytes, float* sums,
                     rg_stream_t st) {
  RG_CHECK_ARG(dh && a && mean && rstd && scale && shift && sums, "rg_bn_bwd_reduce: null pointer");
  BwdReduceF f{static_cast<const bf16*>(dh), static_cast<const bf16*>(a), mean, rstd, scale, shift, slope, C};
  return run_colreduce(f, M, C, static_cast<float*>(ws), ws_bytes, sums, static_cast<cudaStream_t>(st),
                       "rg_bn_bwd_reduce");
}

int rg_bn_bwd_apply(const void* dh, const void* a, const void* add, const float* mean, const float* rstd,
                    const float* scale, const float* shift, float slope, const float* sums, int M, int C, void* da,
                    void* du_out, rg_stream_t st) {
  RG_CHECK_ARG(dh && a && mean && rstd && scale && shift && sums && da, "rg_bn_bwd_apply: null pointer");
  if (add != nullptr) {
    BwdApplyF<true> f{static_cast<const bf16*>(dh), static_cast<const bf16*>(a), static_cast<const bf16*>(add),
                      static_cast<bf16*>(da), static_cast<bf16*>(du_out), mean, rstd, scale, shift, sums, slope,
                      1.0f / M, C};
    return run_ew_split(f, M, C, static_cast<cudaStream_t>(st), "rg_bn_bwd_apply");
  }
  BwdApplyF<false> f{static_cast<const bf16*>(dh), static_cast<const bf16*>(a), nullptr, static_cast<bf16*>(da),
                     static_cast<bf16*>(du_out), mean, rstd, scale, shift, sums, slope, 1.0f / M, C};
  return run_ew_split(f, M, C, static_cast<cudaStream_t>(st), "rg_bn_bwd_apply");
}

int rg_bn_param_grads(const float* sums, float* dgamma, float* dbeta, int C, float acc_gamma, float acc_beta,
                      rg_stream_t st) {
  RG_CHECK_ARG(sums && dgamma && dbeta && C > 0, "rg_bn_param_grads: bad arguments");
  RG_CUDA(launch_pdl(bn_param_grads_kernel, dim3(ceil_div(C, 256)), dim3(256), 0, static_cast<cudaStream_t>(st), 1, sums,
                     dgamma, dbeta, C, acc_gamma, acc_beta));
  RG_LAUNCH_CHECK("rg_bn_param_grads");
  return 0;
}

int rg_lrelu_bwd(const void* dh, const void* h, float slope, void* da, int M, int C, rg_stream_t st) {
  RG_CHECK_ARG(dh && h && da, "rg_lrelu_bwd: null pointer");
  LreluBwdF f{static_cast<const bf16*>(dh), static_cast<const bf16*>(h), static_cast<bf16*>(da), slope, C};
  return run_ew_split(f, M, C, static_cast<cudaStream_t>(st), "rg_lrelu_bwd");
}

int rg_col_sum(const void* x, int M, int C, void* ws, size_t ws_bytes, float* tmp, float* out, float acc,
               rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(x && tmp && out, "rg_col_sum: null pointer");
  ColSumF f{static_cast<const bf16*>(x), C};
  int rc = run_colreduce(f, M, C, static_cast<float*>(ws), ws_bytes, tmp, st, "rg_col_sum");
  if (rc) return rc;
  vec_axpby_kernel<<<ceil_div(C, 256), 256, 0, st>>>(tmp, out, C, 1.0f, acc);
  RG_LAUNCH_CHECK("rg_col_sum");
  return 0;
}

int rg_bn_gp_reduce(const void* ggI, const void* a, const void* gO, const float* mean, const float* rstd, int M, int C,
                    void* ws, size_t ws_bytes, float* q, rg_stream_t st) {
  RG_CHECK_ARG(ggI && a && gO && mean && rstd && q, "rg_bn_gp_reduce: null pointer");
  GpReduceF f{static_cast<const bf16*>(ggI), static_cast<const bf16*>(a), static_cast<const bf16*>(gO), mean, rstd, C};
  return run_colreduce(f, M, C, static_cast<float*>(ws), ws_bytes, q, static_cast<cudaStream_t>(st), "rg_bn_gp_reduce");
}

int rg_bn_gp_apply(const void* ggI, const void* a, const void* gO, const float* mean, const float* rstd,
                   const float* gamma, const float* scale, const float* shift, float slope, const float* s,
                   const float* q, int M, int C, void* A_dh, void* A_a, float* dgamma, float dgamma_acc,
                   rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(ggI && a && gO && mean && rstd && gamma && scale && shift && s && q && A_dh && A_a,
               "rg_bn_gp_apply: null pointer");
  GpApplyF f{static_cast<const bf16*>(ggI), static_cast<const bf16*>(a), static_cast<const bf16*>(gO),
             static_cast<bf16*>(A_dh), static_cast<bf16*>(A_a), mean, rstd, gamma, scale, shift, s, q, slope,
             1.0f / M, C};
  int rc = run_ew_split(f, M, C, st, "rg_bn_gp_apply");
  if (rc) return rc;
  if (dgamma) {
    bn_gp_dgamma_kernel<<<ceil_div(C, 256), 256, 0, st>>>(s, q, rstd, dgamma, C, 1.0f / M, dgamma_acc);
    RG_LAUNCH_CHECK("rg_bn_gp_apply(dgamma)");
  }
  return 0;
}

int rg_latent_prep(const float* noise, const float* z, int B, int E, int z_rows, void* lat_bf16, float* lat_f32,
                   rg_stream_t st) {
  RG_CHECK_ARG(noise && z && B > 0 && E > 0 && (z_rows == B || z_rows == 1), "rg_latent_prep: bad arguments");
  latent_prep_kernel<<<ceil_div(E, 32), 256, 0, static_cast<cudaStream_t>(st)>>>(noise, z, B, E, z_rows,
                                                                                  static_cast<bf16*>(lat_bf16), lat_f32);
  RG_LAUNCH_CHECK("rg_latent_prep");
  return 0;
}

int rg_im2col_img(const float* x, const float* y, int mode, const float* eps_dev, const float* mul_dev, int B,
                  int Cimg, int S, void* col, float* mixed_out, rg_stream_t st) {
  RG_CHECK_ARG(x && col && B > 0 && Cimg >= 1 && Cimg <= 4 && S >= 4 && is_pow2(S),
               "rg_im2col_img: need 1..4 channels and a power-of-two image side (S=%d)", S);
  RG_CHECK_ARG(mode == 0 || y, "rg_im2col_img: mode %d needs a second image", mode);
  // output rows per block.  Measured on B200: R = 4 (1.25 instead of 2 reads per image row, 4x fewer blocks) is SLOWER
  // (92 vs 86 us at B = 64, 256x256) -- the kernel is bound by its shared-memory gather, not by the image reads.
  int R = 1;
  while (R > 1 && ((S / 2) % R != 0 || static_cast<size_t>(Cimg) * (2 * R + 2) * (S + 2) * sizeof(float) > 48 * 1024))
    R >>= 1;
  const size_t smem = static_cast<size_t>(Cimg) * (2 * R + 2) * (S + 2) * sizeof(float);
  RG_CHECK_ARG(smem <= 48 * 1024, "rg_im2col_img: image side %d too large for the row-staging buffer", S);
  im2col_img_kernel<<<B * (S / 2 / R), 256, smem, static_cast<cudaStream_t>(st)>>>(x, y, mode, eps_dev, mul_dev, B, Cimg,
                                                                                   S, R, static_cast<bf16*>(col), mixed_out);
  RG_LAUNCH_CHECK("rg_im2col_img");
  return 0;
}

int rg_img_channel_sum(const float* x, const float* y, int mode, int B, int Cimg, int S, float* partial_ws,
                       int partial_len, float* out, float acc, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(x && out && partial_ws && (mode == 0 || y), "rg_img_channel_sum: bad arguments");
  const int nblk = std::min(256, partial_len / std::max(Cimg, 1));
  RG_CHECK_ARG(nblk >= 1, "rg_img_channel_sum: partial workspace too small");
  img_channel_sum_stage1<<<dim3(nblk, Cimg), 256, 0, st>>>(x, y, mode, B, Cimg, S, partial_ws);
  RG_LAUNCH_CHECK("rg_img_channel_sum(stage1)");
  img_channel_sum_stage2<<<Cimg, 256, 0, st>>>(partial_ws, nblk, out, acc);
  RG_LAUNCH_CHECK("rg_img_channel_sum(stage2)");
  return 0;
}

int rg_col2im_img(const float* col, int ldc, const float* bias, int act_tanh, int B, int Cimg, int H, int W, float* img,
                  rg_stream_t st) {
  RG_CHECK_ARG(col && img && Cimg >= 1 && Cimg <= 4 && ldc >= 16 * Cimg && is_pow2(H) && is_pow2(W),
               "rg_col2im_img: need power-of-two H, W and ldc >= 16*Cimg");
  const size_t n = static_cast<size_t>(B) * 4 * H * W;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  col2im_img_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(col, ldc, bias, act_tanh, B, Cimg, H, W, img);
  RG_LAUNCH_CHECK("rg_col2im_img");
  return 0;
}

int rg_pack_edge_t(const float* W, void* w_colT, int Cp, int Cimg, int rows, rg_stream_t st) {
  RG_CHECK_ARG(W && w_colT && Cimg >= 1 && Cimg <= 4 && rows >= 16 * Cimg, "rg_pack_edge_t: bad arguments");
  pack_edge_t_kernel<<<ceil_div(rows * Cp, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(
      W, static_cast<bf16*>(w_colT), Cp, Cimg, rows);
  RG_LAUNCH_CHECK("rg_pack_edge_t");
  return 0;
}

int rg_unpack_edge_grad(const float* dcol, float* dW, int Cp, int Cimg, float acc, rg_stream_t st) {
  RG_CHECK_ARG(dcol && dW && Cimg >= 1 && Cimg <= 4, "rg_unpack_edge_grad: bad arguments");
  unpack_edge_grad_kernel<<<ceil_div(Cp * Cimg * 16, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(dcol, dW, Cp,
                                                                                                    Cimg, acc);
  RG_LAUNCH_CHECK("rg_unpack_edge_grad");
  return 0;
}

int rg_pack_head(const float* W, float* w_head, int C, rg_stream_t st) {
  RG_CHECK_ARG(W && w_head && C > 0, "rg_pack_head: bad arguments");
  pack_head_kernel<<<ceil_div(16 * C, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(W, w_head, C);
  RG_LAUNCH_CHECK("rg_pack_head");
  return 0;
}

int rg_head_fwd(const void* h5, const float* w_head, int B, int K, float slope, float* a6, float* out,
                rg_stream_t st) {
  RG_CHECK_ARG(h5 && w_head && a6 && out && K % 8 == 0, "rg_head_fwd: bad arguments");
  head_fwd_kernel<<<B, 256, 0, static_cast<cudaStream_t>(st)>>>(static_cast<const bf16*>(h5), w_head, K, slope, a6,
                                                               out);
  RG_LAUNCH_CHECK("rg_head_fwd");
  return 0;
}

int rg_head_bwd_data(const float* a6, const float* dout, float dout_const, const float* w_head, int B, int K,
                     float slope, float* da6, void* dh5, rg_stream_t st) {
  RG_CHECK_ARG(a6 && w_head && dh5 && K % 8 == 0, "rg_head_bwd_data: bad arguments");
  const size_t nvec = static_cast<size_t>(B) * (K / 8);
  const int grid = static_cast<int>(std::min<size_t>((nvec + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  head_bwd_data_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(a6, dout, dout_const, w_head, B, K, slope,
                                                                        da6, static_cast<bf16*>(dh5));
  RG_LAUNCH_CHECK("rg_head_bwd_data");
  return 0;
}

int rg_head_wgrad(const float* da6, const void* x, int B, int K, int C, float* dW, float acc, rg_stream_t st) {
  RG_CHECK_ARG(da6 && x && dW && K == 16 * C, "rg_head_wgrad: bad arguments");
  head_wgrad_kernel<<<ceil_div(K, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(da6, static_cast<const bf16*>(x), B,
                                                                                 K, C, dW, acc);
  RG_LAUNCH_CHECK("rg_head_wgrad");
  return 0;
}

int rg_wgan_loss(const float* a, float sign_a, const float* b, float sign_b, int B, float* loss_out, rg_stream_t st) {
  RG_CHECK_ARG(a && loss_out && B > 0, "rg_wgan_loss: bad arguments");
  wgan_loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(st)>>>(a, sign_a, b, sign_b, B, loss_out);
  RG_LAUNCH_CHECK("rg_wgan_loss");
  return 0;
}

int rg_gp_norm(const float* g, size_t n, float lambd, float* partial_ws, int partial_len, float* out3,
               rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(g && partial_ws && out3 && partial_len >= 1, "rg_gp_norm: bad arguments");
  const int blocks = static_cast<int>(std::min<size_t>(static_cast<size_t>(partial_len), (n + 255) / 256));
  sumsq_stage1_kernel<<<blocks, 256, 0, st>>>(g, n, partial_ws);
  RG_LAUNCH_CHECK("rg_gp_norm(stage1)");
  gp_finalize_kernel<<<1, 256, 0, st>>>(partial_ws, blocks, lambd, out3);
  RG_LAUNCH_CHECK("rg_gp_norm(finalize)");
  return 0;
}

int rg_adam_table_bytes(int num_chunks) { return static_cast<int>(sizeof(AdamChunk)) * num_chunks; }

// chunks_host: arrays of num_tensors pointers / sizes; table_dev: device buffer for the chunk table (filled here with
// a synchronous-with-stream async copy from the caller's pinned or pageable host staging buffer table_host).
int rg_adam_build_table(void* const* params, void* const* grads, void* const* ms, void* const* vs,
                        void* const* shadows, const int64_t* sizes, int num_tensors, int chunk_elems, void* table_host,
                        int max_chunks) {
  return rg_adam_build_table_pitched(params, grads, ms, vs, shadows, nullptr, nullptr, sizes, num_tensors, chunk_elems,
                                     table_host, max_chunks);
}

int rg_adam_build_table_pitched(void* const* params, void* const* grads, void* const* ms, void* const* vs,
                                void* const* shadows, const int* shadow_cols, const int* shadow_pitch,
                                const int64_t* sizes, int num_tensors, int chunk_elems, void* table_host,
                                int max_chunks) {
  RG_CHECK_ARG(params && grads && ms && vs && sizes && table_host && chunk_elems > 0, "rg_adam_build_table: bad arguments");
  AdamChunk* t = static_cast<AdamChunk*>(table_host);
  int n = 0;
  for (int i = 0; i < num_tensors; ++i) {
    for (int64_t off = 0; off < sizes[i]; off += chunk_elems) {
      if (n >= max_chunks) {
        set_error("rg_adam_build_table: table too small (%d chunks)", max_chunks);
        return RG_EWORKSPACE;
      }
      t[n].p = static_cast<float*>(params[i]) + off;
      t[n].g = static_cast<const float*>(grads[i]) + off;
      t[n].m = static_cast<float*>(ms[i]) + off;
      t[n].v = static_cast<float*>(vs[i]) + off;
      t[n].shadow = (shadows && shadows[i]) ? static_cast<__nv_bfloat16*>(shadows[i]) + off : nullptr;
      t[n].sh_cols = t[n].sh_pitch = t[n].sh_col0 = 0;
      if (shadows && shadows[i] && shadow_cols && shadow_pitch && shadow_pitch[i] > shadow_cols[i] && shadow_cols[i] > 0) {
        const int64_t row = off / shadow_cols[i];
        t[n].shadow = static_cast<__nv_bfloat16*>(shadows[i]) + row * shadow_pitch[i];
        t[n].sh_cols = shadow_cols[i];
        t[n].sh_pitch = shadow_pitch[i];
        t[n].sh_col0 = static_cast<int>(off - row * shadow_cols[i]);
      }
      t[n].n = static_cast<int>(std::min<int64_t>(chunk_elems, sizes[i] - off));
      ++n;
    }
  }
  return n;   // number of chunks (>= 0)
}

int rg_adam_step(const void* table_dev, int num_chunks, float lr, float beta1, float beta2, float eps, int step,
                 int do_clamp, float clamp_lo, float clamp_hi, float grad_scale, rg_stream_t st) {
  RG_CHECK_ARG(table_dev && num_chunks > 0 && step >= 1, "rg_adam_step: bad arguments");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adam_kernel<<<num_chunks, 256, 0, static_cast<cudaStream_t>(st)>>>(static_cast<const AdamChunk*>(table_dev), lr, beta1,
                                                                     beta2, eps, bc1, sqrtf(bc2), clamp_lo, clamp_hi,
                                                                     do_clamp, grad_scale, nullptr);
  RG_LAUNCH_CHECK("rg_adam_step");
  return 0;
}

int rg_adam_step_dyn(const void* table_dev, int num_chunks, float* dyn, float beta1, float beta2, float eps,
                     int do_clamp, float clamp_lo, float clamp_hi, float grad_scale, rg_stream_t st) {
  RG_CHECK_ARG(table_dev && num_chunks > 0 && dyn, "rg_adam_step_dyn: bad arguments");
  adam_tick_kernel<<<1, 1, 0, static_cast<cudaStream_t>(st)>>>(dyn);
  RG_LAUNCH_CHECK("rg_adam_step_dyn(tick)");
  adam_kernel<<<num_chunks, 256, 0, static_cast<cudaStream_t>(st)>>>(static_cast<const AdamChunk*>(table_dev), 0.0f, beta1,
                                                                     beta2, eps, 1.0f, 1.0f, clamp_lo, clamp_hi, do_clamp,
                                                                     grad_scale, dyn);
  RG_LAUNCH_CHECK("rg_adam_step_dyn");
  return 0;
}

int rg_clamp(float* p, size_t n, float lo, float hi, rg_stream_t st) {
  RG_CHECK_ARG(p && n > 0, "rg_clamp: bad arguments");
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  clamp_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(p, n, lo, hi);
  RG_LAUNCH_CHECK("rg_clamp");
  return 0;
}

int rg_nvls_allreduce(float* mc, size_t offset, size_t n, int max_ctas, rg_stream_t st) {
  RG_CHECK_ARG(mc && n > 0 && n % 4 == 0 && offset % 4 == 0 && reinterpret_cast<uintptr_t>(mc) % 16 == 0,
               "rg_nvls_allreduce: need a 16-byte aligned multicast pointer and offset / n multiples of 4 floats");
  const size_t n4 = n / 4;
  const int cap = max_ctas > 0 ? max_ctas : 128;
  const int grid = static_cast<int>(std::min<size_t>((n4 + 8 * 128 - 1) / (8 * 128), static_cast<size_t>(cap)));
  nvls_allreduce_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(st)>>>(mc + offset, n4);
  RG_LAUNCH_CHECK("rg_nvls_allreduce");
  return 0;
}

int rg_slices_sum(const float* src, int nparts, size_t stride, size_t n, float* out, rg_stream_t st) {
  RG_CHECK_ARG(src && out && nparts >= 1 && n > 0 && n % 4 == 0 && stride % 4 == 0 && stride >= n &&
                   reinterpret_cast<uintptr_t>(src) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0,
               "rg_slices_sum: need n and stride multiples of 4 floats, stride >= n, 16-byte aligned pointers");
  const size_t n4 = n / 4;
  const int grid = static_cast<int>(std::min<size_t>((n4 + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  slices_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(src, nparts, stride, n4, out);
  RG_LAUNCH_CHECK("rg_slices_sum");
  return 0;
}

int rg_tiles_u8_to_nchw(const void* tiles, float* img, int B, int C, int S, int swap_rb, rg_stream_t st) {
  RG_CHECK_ARG(tiles && img && B > 0 && C >= 1 && C <= 4 && S >= 4 && S % 4 == 0,
               "rg_tiles_u8_to_nchw: need 1..4 channels and an image side that is a multiple of 4 (S=%d)", S);
  const size_t nquad = static_cast<size_t>(B) * S * (S / 4);
  const int grid = static_cast<int>(std::min<size_t>((nquad + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  tiles_u8_to_nchw_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(static_cast<const uint8_t*>(tiles), img, B, C,
                                                                           S, swap_rb);
  RG_LAUNCH_CHECK("rg_tiles_u8_to_nchw");
  return 0;
}

int rg_tiles_to_unit_nhwc(const float* img, float* out, int B, int C, int S, rg_stream_t st) {
  RG_CHECK_ARG(img && out && B > 0, "rg_tiles_to_unit_nhwc: bad arguments");
  const size_t n = static_cast<size_t>(B) * C * S * S;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  nchw_to_unit_nhwc_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(img, out, B, C, S);
  RG_LAUNCH_CHECK("rg_tiles_to_unit_nhwc");
  return 0;
}


int rg_mul_cast_pad_bf16(const float* src, const float* mul, float scale, void* dst, int rows, int cols, int cols_pad,
                         rg_stream_t st) {
  RG_CHECK_ARG(src && dst && rows > 0 && cols > 0 && cols_pad >= cols, "rg_mul_cast_pad_bf16: bad arguments");
  const size_t n = static_cast<size_t>(rows) * cols_pad;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  mul_cast_pad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(src, mul, scale, static_cast<bf16*>(dst), rows,
                                                                       cols, cols_pad);
  RG_LAUNCH_CHECK("rg_mul_cast_pad_bf16");
  return 0;
}

int rg_vae_reparam(const float* mulv, const float* eps, int B, int Z, void* z, float* partial, int partial_len,
                   rg_stream_t st) {
  RG_CHECK_ARG(mulv && eps && z && partial && partial_len >= 1, "rg_vae_reparam: bad arguments");
  const int grid = std::min(partial_len, 256);
  vae_reparam_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(mulv, eps, B, Z, static_cast<bf16*>(z), partial);
  RG_LAUNCH_CHECK("rg_vae_reparam");
  return grid;
}

int rg_vae_recon(const float* pre, int ldp, const float* x, int B, int F, float gscale, void* dpre, float* partial,
                 int partial_len, rg_stream_t st) {
  RG_CHECK_ARG(pre && x && dpre && partial && partial_len >= 1 && ldp >= F, "rg_vae_recon: bad arguments");
  const int grid = std::min(partial_len, 256);
  vae_recon_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(pre, ldp, x, B, F, gscale, static_cast<bf16*>(dpre),
                                                                    partial);
  RG_LAUNCH_CHECK("rg_vae_recon");
  return grid;
}

int rg_vae_latent_grad(const void* dz, const float* mulv, const float* eps, int B, int Z, float kscale, void* dcat,
                       rg_stream_t st) {
  RG_CHECK_ARG(dz && mulv && eps && dcat, "rg_vae_latent_grad: bad arguments");
  const size_t n = static_cast<size_t>(B) * Z;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 8));
  vae_latent_grad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(st)>>>(static_cast<const bf16*>(dz), mulv, eps, B, Z,
                                                                          kscale, static_cast<bf16*>(dcat));
  RG_LAUNCH_CHECK("rg_vae_latent_grad");
  return 0;
}

int rg_vae_loss_finalize(const float* p_sse, int n1, const float* p_kld, int n2, int B, int F, float beta, float* out3,
                         rg_stream_t st) {
  RG_CHECK_ARG(p_sse && p_kld && out3 && B > 0 && F > 0, "rg_vae_loss_finalize: bad arguments");
  vae_loss_finalize_kernel<<<1, 256, 0, static_cast<cudaStream_t>(st)>>>(
      p_sse, n1, p_kld, n2, 1.0f / (static_cast<float>(B) * static_cast<float>(F)), 1.0f / B, beta, out3);
  RG_LAUNCH_CHECK("rg_vae_loss_finalize");
  return 0;
}


int rg_upsample2x_reflectpad(const void* h, void* u, int B, int H, int W, int C, rg_stream_t st) {
  RG_CHECK_ARG(h && u && H >= 2 && W >= 2, "rg_upsample2x_reflectpad: bad arguments");
  UpPadF f{static_cast<const bf16*>(h), static_cast<bf16*>(u), H, W, C};
  return run_ew(f, B * (2 * H + 2) * (2 * W + 2), C, static_cast<cudaStream_t>(st), "rg_upsample2x_reflectpad");
}

int rg_upsample2x_reflectpad_bwd(const void* du, void* dh, int B, int H, int W, int C, rg_stream_t st) {
  RG_CHECK_ARG(du && dh && H >= 2 && W >= 2, "rg_upsample2x_reflectpad_bwd: bad arguments");
  UpPadBwdF f{static_cast<const bf16*>(du), static_cast<bf16*>(dh), H, W, C};
  return run_ew(f, B * H * W, C, static_cast<cudaStream_t>(st), "rg_upsample2x_reflectpad_bwd");
}

size_t rg_upg_last_ws_bytes(int B, int S, int C, int Cimg) {
  const size_t npix = static_cast<size_t>(B) * S * S;
  const int blocks = static_cast<int>(std::min<size_t>((npix + 1023) / 1024, static_cast<size_t>(num_sms()) * 4));
  return static_cast<size_t>(blocks) * Cimg * 9 * C * sizeof(float);
}

int rg_upg_last_bwd(const void* u, const float* dout, const float* W, int B, int S, int C, int Cimg, float* dW, void* du,
                    void* ws, size_t ws_bytes, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(u && dout && W && dW && C == 64 && Cimg >= 1 && Cimg <= 3, "rg_upg_last_bwd: needs C == 64, Cimg <= 3");
  const size_t npix = static_cast<size_t>(B) * S * S;
  const int blocks = static_cast<int>(std::min<size_t>((npix + 1023) / 1024, static_cast<size_t>(num_sms()) * 4));
  const int ppb = static_cast<int>((npix + blocks - 1) / blocks);
  const size_t need = static_cast<size_t>(blocks) * Cimg * 9 * C * sizeof(float);
  if (!ws || ws_bytes < need) {
    set_error("rg_upg_last_bwd: workspace too small (need %zu, have %zu)", need, ws_bytes);
    return RG_EWORKSPACE;
  }
  upg_last_wgrad_stage1<<<blocks, 256, 3 * 64 * sizeof(float), st>>>(static_cast<const bf16*>(u), dout, B, S, C, Cimg,
                                                                     ppb, static_cast<float*>(ws));
  RG_LAUNCH_CHECK("rg_upg_last_bwd(wgrad1)");
  upg_last_wgrad_stage2<<<ceil_div(Cimg * 9 * C, 256), 256, 0, st>>>(static_cast<const float*>(ws), blocks, C, Cimg, dW);
  RG_LAUNCH_CHECK("rg_upg_last_bwd(wgrad2)");
  if (du) {
    const size_t nvec = static_cast<size_t>(B) * (S + 2) * (S + 2) * (C / 8);
    const int grid = static_cast<int>(std::min<size_t>((nvec + 255) / 256, static_cast<size_t>(num_sms()) * 8));
    upg_last_dgrad_kernel<<<grid, 256, Cimg * 9 * C * sizeof(float), st>>>(dout, W, B, S, C, Cimg,
                                                                           static_cast<bf16*>(du));
    RG_LAUNCH_CHECK("rg_upg_last_bwd(dgrad)");
  }
  return 0;
}

}  // extern "C"
