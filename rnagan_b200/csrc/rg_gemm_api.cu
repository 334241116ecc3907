// rg_gemm_api.cu -- C-ABI entry points for the tcgen05 tile engine (rg_gemm.cuh) and the weight packers.
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include "rg_gemm.cuh"
#include "rg_host.cuh"

namespace rg {

// ------------------------------------------------------------------------------------------------ errors / device
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return static_cast<int>(e);
}
static std::atomic<long long> g_launches{0};
static long long* g_prof = nullptr;   // tools/gemm_prof.py: per-CTA role clocks of the next forward launches
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    // RG_NUM_SMS caps the persistent grids (even, >= 2): data-parallel runs can leave a few SMs to NCCL's all-reduce
    // kernels, which cannot co-reside with a tile-engine CTA that owns the whole shared memory of its SM
    const char* e = getenv("RG_NUM_SMS");
    if (e && atoi(e) >= 2 && atoi(e) < sms) sms = atoi(e) & ~1;
  }
  return sms;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RG_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// ------------------------------------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                  const cuuint32_t* box) {
  EncodeTiledFn fn = get_encode();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return RG_EDRIVER;
  }
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), dims,
                  strides_b, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] base %p",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
    return RG_EDRIVER;
  }
  return 0;
}

int encode_map_4d(CUtensorMap* m, const void* base, uint64_t C, uint64_t W, uint64_t H, uint64_t B, uint64_t sw,
                  uint64_t sh, uint64_t sb, uint32_t boxc, uint32_t bw, uint32_t bh, uint32_t bb) {
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {sw * 2, sh * 2, sb * 2};
  cuuint32_t box[4] = {boxc, bw, bh, bb};
  return encode(m, base, 4, dims, strides, box);
}
int encode_map_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t boxc,
                  uint32_t box_rows) {
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {boxc, box_rows};
  return encode(m, base, 2, dims, strides, box);
}

// ------------------------------------------------------------------------------------------------ geometry
struct Boxing {
  int bw, bh, bb, tw, th, tb;
};
static Boxing make_boxing(int B, int H, int W, int rows) {
  Boxing g;
  if (is_pow2(W) && is_pow2(H)) {
    g.bw = std::min(W, rows);
    g.bh = std::min(H, rows / g.bw);
    g.bb = rows / (g.bw * g.bh);
  } else {
    // general grids (the reflect-padded 2H+2 x 2W+2 grid of the resize-conv generator): widest divisor of W that
    // fits, as many rows / images as fit; the tile then has bw*bh*bb <= rows valid rows (the rest is masked)
    g.bw = W;
    if (W > rows)
      for (int d = rows; d >= 1; --d)
        if (W % d == 0) { g.bw = d; break; }
    g.bh = std::max(1, std::min(H, rows / g.bw));
    g.bb = std::max(1, std::min(B, rows / (g.bw * g.bh)));
  }
  g.tw = ceil_div(W, g.bw);
  g.th = ceil_div(H, g.bh);
  g.tb = ceil_div(B, g.bb);
  return g;
}

// kh (or kw) -> (offset in the low-res grid, parity of the high-res row) for y = 2i - 1 + kh
static const int kDownOff[4] = {-1, 0, 0, 1};
static const int kDownPar[4] = {1, 0, 1, 0};
// output parity r, tap index t -> (low-res offset, kernel index) for the transposed form
static const int kUpOff[2][2] = {{0, -1}, {0, 1}};
static const int kUpK[2][2] = {{1, 3}, {2, 0}};

static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (!e || !e[0]) return dflt;
  return e[0] != '0';
}

// CTA pairs (tcgen05 cta_group::2): the two SMs of a TPC share one 256-row tile, each fetching half of the B rows, so
// an SM ingests 1.5x fewer operand bytes per FLOP than with two independent 128 x 256 tiles -- the L2 -> SM path, not
// the tensor pipe, bounds the single-CTA kernel (profiles/r1_ncu_gemm.txt: 52 B/clk/SM at 55 % tensor peak).
static bool pair_allowed(const FwdArgs& a, int out_kind) {
  static const bool allow = env_flag("RG_CG2", true);
  return allow && out_kind == OUT_BF16_NHWC && a.m_tiles >= 2;
}
static int pick_block_n(int n_total, int tiles_per_n, int cg = 1) {
  int bn = 256;
  while ((bn >> 1) >= n_total && bn > 16) bn >>= 1;   // smallest power of two >= n_total (rows beyond zero-fill)
  const int target = (num_sms() / cg * 4) / 5;
  const int units_per_n = ceil_div(tiles_per_n, cg);
  while (bn > 64 && ceil_div(n_total, bn) * units_per_n < target) bn >>= 1;
  return bn;
}
static int pick_cg(const FwdArgs& a, int out_kind) {
  if (!pair_allowed(a, out_kind)) return 1;
  if (a.block_n < 32 || a.block_n % 16 != 0) return 1;
  if (a.b_mn && a.block_n % 128 != 0) return 1;       // MN-major B arrives in 64-column slabs: each CTA needs >= 1
  return 2;
}

static int ensure_attrs() {
  static bool done = false;
  if (!done) {
    RG_CUDA(cudaFuncSetAttribute(gemm_fwd_kernel<OUT_BF16_NHWC, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemMaxBytes));
    RG_CUDA(cudaFuncSetAttribute(gemm_fwd_kernel<OUT_BF16_NHWC, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemMaxBytes));
    RG_CUDA(cudaFuncSetAttribute(gemm_fwd_kernel<OUT_BF16_NHWC, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemMaxBytes));
    RG_CUDA(cudaFuncSetAttribute(gemm_fwd_kernel<OUT_BF16_NHWC, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemMaxBytes));
    RG_CUDA(cudaFuncSetAttribute(gemm_fwd_kernel<OUT_F32_NHWC, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemMaxBytes));
    RG_CUDA(cudaFuncSetAttribute(gemm_fwd_kernel<OUT_F32_NCHW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kSmemMaxBytes));
    RG_CUDA(cudaFuncSetAttribute(gemm_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgradSmemBytes));
    done = true;
  }
  return 0;
}

template <int OUT, int CG, bool AUX = false>
static int launch_fwd_t(const GemmMaps& maps, const FwdArgs& a, int grid, size_t smem, cudaStream_t st) {
  RG_CUDA(launch_pdl(gemm_fwd_kernel<OUT, CG, AUX>, dim3(grid), dim3(kGemmThreads), smem, st, CG, maps, a));
  RG_LAUNCH_CHECK("gemm_fwd_kernel");
  return 0;
}

size_t stats_ws_floats(int C) { return static_cast<size_t>(num_sms()) * 2 * C; }

// cg: the CTA-group size the caller encoded the B map for (pick_cg).  stats_ws: optional [num_sms][2][n_total] fp32.
static int launch_fwd(GemmMaps& maps, FwdArgs& a, int out_kind, int cg, float* stats_ws, cudaStream_t st,
                      const rg_epilogue_aux* aux = nullptr) {
  int rc = ensure_attrs();
  if (rc) return rc;
  if (aux && aux->aux && aux->mode != 0) {
    if (aux->mode != 1 && aux->mode != 2) {
      set_error("rg_epilogue_aux: mode must be 0, 1 or 2 (got %d)", aux->mode);
      return RG_EINVAL;
    }
    if (aux->mode == 2 && !(aux->mean && aux->rstd && aux->scale && aux->shift)) {
      set_error("rg_epilogue_aux: mode 2 needs mean, rstd, scale and shift");
      return RG_EINVAL;
    }
    if ((reinterpret_cast<uintptr_t>(aux->aux) & 15) != 0 || a.col_scale || a.col_shift || a.slope != 1.0f) {
      set_error("rg_epilogue_aux: aux must be 16-byte aligned and excludes the affine / activation epilogue");
      return RG_EINVAL;
    }
    a.aux = static_cast<const __nv_bfloat16*>(aux->aux);
    a.aux_mode = aux->mode;
    a.aux_mean = aux->mean; a.aux_rstd = aux->rstd; a.aux_scale = aux->scale; a.aux_shift = aux->shift;
    a.aux_slope = aux->slope;
  }
  static const bool allow_tma_store = env_flag("RG_TMA_STORE", true);
  a.tma_store = (allow_tma_store && out_kind == OUT_BF16_NHWC && a.block_n % 64 == 0 && a.OC % 8 == 0 &&
                 (reinterpret_cast<uintptr_t>(a.out) & 15) == 0) ? 1 : 0;
  if (a.aux_mode != 0 && !(a.tma_store && a.n_total % 64 == 0 && a.n_valid == a.n_total)) {
    set_error("the fused elementwise backward needs a bf16 output with a multiple of 64 channels (n=%d)", a.n_total);
    return RG_EINVAL;
  }
  if (stats_ws && !(a.tma_store && a.n_total % 64 == 0)) {
    set_error("fused statistics need a bf16 output with a multiple of 64 channels (n_total=%d block_n=%d)", a.n_total,
              a.block_n);
    return RG_EINVAL;
  }
  a.stats = stats_ws;
  {
    const char* e = getenv("RG_WHATIF");    // read per call: the profiling tool flips it between launches
    a.whatif = e ? atoi(e) : 0;
  }
  a.prof = g_prof;
  // shared-memory plan: ring of (A 16 KiB + this CTA's share of B) stages, then the epilogue staging slabs
  a.b_stage_bytes = (a.block_n / cg) * 128;
  const int stage = kAStageBytes + a.b_stage_bytes;
  const int avail = kSmemMaxBytes - kSmemFixedBytes;
  a.nbuf = a.tma_store ? 2 : 0;
  a.naux = a.aux_mode != 0 ? 2 : 0;
  int ns = (avail - (a.nbuf + a.naux) * kStagingBytes) / stage;
  if (a.tma_store && ns < 4 && a.naux == 0) {
    a.nbuf = 1;
    ns = (avail - kStagingBytes) / stage;
  }
  if (ns < 2) {
    set_error("internal: shared-memory plan leaves %d stages (block_n=%d cg=%d aux=%d)", ns, a.block_n, cg, a.aux_mode);
    return RG_EINVAL;
  }
  a.nstages = std::min(ns, kMaxStages);
  size_t smem = static_cast<size_t>(a.nstages) * stage + (a.nbuf + a.naux) * kStagingBytes + kSmemFixedBytes;
  if (a.tma_store) {
    // per-phase output views: pixel (b, i*sy + oy, j*sx + ox), channels [0, n_valid); TMA clips what lies outside
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(a.out);
    const int out_maps = a.merged ? 4 : a.num_phases;        // merged tiles store all four phase views
    for (int ph = 0; ph < out_maps; ++ph) {
      const __nv_bfloat16* b0 = base + (static_cast<size_t>(a.oy[ph]) * a.OW + a.ox[ph]) * a.OC;
      rc = encode_map_4d(&maps.o[ph], b0, a.n_valid, a.W, a.H, a.nB, static_cast<uint64_t>(a.sx) * a.OC,
                         static_cast<uint64_t>(a.sy) * a.OW * a.OC, static_cast<uint64_t>(a.OH) * a.OW * a.OC, 64, a.bw,
                         a.bh, a.bb);
      if (rc) return rc;
      if (a.aux_mode != 0) {       // the aux tensor has the layout of the output: same views, other base
        const __nv_bfloat16* x0 = a.aux + (static_cast<size_t>(a.oy[ph]) * a.OW + a.ox[ph]) * a.OC;
        rc = encode_map_4d(&maps.x[ph], x0, a.n_valid, a.W, a.H, a.nB, static_cast<uint64_t>(a.sx) * a.OC,
                           static_cast<uint64_t>(a.sy) * a.OW * a.OC, static_cast<uint64_t>(a.OH) * a.OW * a.OC, 64,
                           a.bw, a.bh, a.bb);
        if (rc) return rc;
      }
    }
  }
  const int slots = ceil_div(a.m_tiles, cg) * a.n_tiles * a.num_phases;
  const int grid = std::min(slots, num_sms() / cg) * cg;
  if (stats_ws && grid < num_sms()) {
    const size_t row = 2ull * a.n_total * sizeof(float);
    RG_CUDA(cudaMemsetAsync(stats_ws + static_cast<size_t>(grid) * 2 * a.n_total, 0, (num_sms() - grid) * row, st));
  }
  if (out_kind == OUT_BF16_NHWC) {
    if (a.aux_mode != 0) {
      if (cg == 2) return launch_fwd_t<OUT_BF16_NHWC, 2, true>(maps, a, grid, smem, st);
      return launch_fwd_t<OUT_BF16_NHWC, 1, true>(maps, a, grid, smem, st);
    }
    if (cg == 2) return launch_fwd_t<OUT_BF16_NHWC, 2>(maps, a, grid, smem, st);
    return launch_fwd_t<OUT_BF16_NHWC, 1>(maps, a, grid, smem, st);
  }
  if (cg != 1) {
    set_error("internal: CTA pairs are only instantiated for bf16 outputs");
    return RG_EINVAL;
  }
  if (out_kind == OUT_F32_NHWC) return launch_fwd_t<OUT_F32_NHWC, 1>(maps, a, grid, smem, st);
  return launch_fwd_t<OUT_F32_NCHW, 1>(maps, a, grid, smem, st);
}

static void fill_common(FwdArgs& a, int B, int H, int W) {
  memset(&a, 0, sizeof(a));
  Boxing g = make_boxing(B, H, W, kBlockM);
  a.nB = B; a.H = H; a.W = W;
  a.bw = g.bw; a.bh = g.bh; a.bb = g.bb;
  a.rows_valid = g.bw * g.bh * g.bb;
  a.tw = g.tw; a.th = g.th; a.tb = g.tb;
  a.m_tiles = g.tw * g.th * g.tb;
  a.slope = 1.0f;
  a.sy = a.sx = 1;
  a.num_phases = 1;
}

// four parity views of hi[B,2H,2W,C]: view (ph,pw) holds pixels (2i+ph, 2j+pw)
static int encode_parity_maps(GemmMaps& maps, const void* hi, int B, int H, int W, int C, int bw, int bh, int bb) {
  const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(hi);
  for (int ph = 0; ph < 2; ++ph)
    for (int pw = 0; pw < 2; ++pw) {
      const __nv_bfloat16* b0 = base + (static_cast<size_t>(ph) * 2 * W + pw) * C;
      int rc = encode_map_4d(&maps.a[ph * 2 + pw], b0, C, W, H, B, 2ull * C, 4ull * W * C, 4ull * H * W * C, 64, bw,
                             bh, bb);
      if (rc) return rc;
    }
  return 0;
}

// ------------------------------------------------------------------------------------------------ reduce kernel
template <int TAPS, bool PADDED>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dW,
                                                           int splits, int Cp, int Cs, int Cs_out, float alpha,
                                                           const float* __restrict__ alpha_dev, float beta) {
  // ws: [splits][TAPS][Cp][Cs] partials; dW: [Cp][Cs_out][TAPS] (PADDED: Cs_out < Cs, the ragged plain-GEMM case)
  const unsigned n_out = static_cast<unsigned>(Cp) * Cs_out;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_out) return;
  const size_t n = static_cast<size_t>(Cp) * Cs;
  unsigned src = idx;
  if (PADDED) {
    const unsigned r = idx / static_cast<unsigned>(Cs_out);
    src = r * Cs + (idx - r * Cs_out);
  }
  const float a = alpha * (alpha_dev ? __ldg(alpha_dev) : 1.0f);
  float acc[TAPS];
#pragma unroll
  for (int t = 0; t < TAPS; ++t) acc[t] = 0.0f;
  for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
    for (int t = 0; t < TAPS; ++t) acc[t] += ws[(static_cast<size_t>(sp) * TAPS + t) * n + src];
  }
  float* o = dW + static_cast<size_t>(idx) * TAPS;
  if (TAPS % 4 == 0) {
#pragma unroll
    for (int t = 0; t < TAPS; t += 4) {
      float4 v = make_float4(a * acc[t], a * acc[t + 1], a * acc[t + 2], a * acc[t + 3]);
      if (beta != 0.0f) {
        const float4 old = *reinterpret_cast<const float4*>(o + t);
        v.x += beta * old.x; v.y += beta * old.y; v.z += beta * old.z; v.w += beta * old.w;
      }
      *reinterpret_cast<float4*>(o + t) = v;
    }
  } else {
#pragma unroll
    for (int t = 0; t < TAPS; ++t) o[t] = (beta != 0.0f ? beta * o[t] : 0.0f) + a * acc[t];
  }
}

// ws: [splits][taps][Cp][Cs] partials -> NATIVE dW[p][tap][s] (channels_last weight gradient): float4 in, float4 out
__global__ void __launch_bounds__(256) wgrad_reduce_native_kernel(const float* __restrict__ ws, float* __restrict__ dW,
                                                                  int splits, int taps, int Cp, int Cs, float alpha,
                                                                  const float* __restrict__ alpha_dev, float beta) {
  const int cs4 = Cs >> 2;
  const size_t n4 = static_cast<size_t>(taps) * Cp * cs4;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // over [p][tap][s4]
  if (idx >= n4) return;
  const int s4 = static_cast<int>(idx % cs4);
  const size_t r = idx / cs4;
  const int tap = static_cast<int>(r % taps);
  const int prow = static_cast<int>(r / taps);
  const float a = alpha * (alpha_dev ? __ldg(alpha_dev) : 1.0f);
  const size_t plane = static_cast<size_t>(Cp) * Cs;
  const float* src = ws + (static_cast<size_t>(tap) * Cp + prow) * Cs + s4 * 4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = 0; sp < splits; ++sp) {
    const float4 v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(sp) * taps * plane);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float4* dst = reinterpret_cast<float4*>(dW) + idx;
  float4 o = make_float4(a * acc.x, a * acc.y, a * acc.z, a * acc.w);
  if (beta != 0.0f) {
    const float4 old = *dst;
    o.x += beta * old.x; o.y += beta * old.y; o.z += beta * old.z; o.w += beta * old.w;
  }
  *dst = o;
}

// Split-K factor: minimise waves x (k-blocks per unit + per-unit epilogue cost) + the cost of moving the fp32 partials
// (every extra split writes and re-reads the whole gradient once).  Favours unit counts that fill whole waves of the
// persistent grid without drowning a small-K layer in partial traffic (L3: 9 splits -> 151 MB of partials for an
// 8 MB gradient; 4 splits run the MMAs 3 % longer and move 67 MB).
static void choose_splits(int units, int num_pb, double out_bytes, int msub, int& splits, int& pb_per_split) {
  const int sms = num_sms();
  // one pixel block = msub x 4 MMAs of 128x256x16 ~ msub x 0.34 us at the power-capped clock; partials stream at ~5 TB/s
  const double pb_us = 0.34 * msub;
  double best = 1e30;
  int best_s = 1;
  for (int sp = 1; sp <= std::min(num_pb, 4 * sms); ++sp) {
    const int per = ceil_div(num_pb, sp);
    const int eff = ceil_div(num_pb, per);
    if (eff != sp) continue;
    const double traffic_us = sp > 1 ? (2.0 * sp + 1.0) * out_bytes / 5.0e6 : 0.0;
    const double cost = static_cast<double>(ceil_div(units * sp, sms)) * (per + 4) + traffic_us / pb_us + 0.05 * sp;
    if (cost < best) { best = cost; best_s = sp; }
  }
  splits = best_s;
  pb_per_split = ceil_div(num_pb, splits);
}

struct WgradGeom {
  Boxing g;
  int num_pb, m_tiles, chunks_s, slabs_per_tile, n_tiles, splits, pb_per_split, taps, msub;
};
static WgradGeom wgrad_geom(int B, int H, int W, int Cp, int Cs, int taps) {
  WgradGeom w;
  w.g = make_boxing(B, H, W, 64);
  w.num_pb = w.g.tw * w.g.th * w.g.tb;
  static const bool allow256 = [] { const char* e = getenv("RG_WGRAD_M256"); return !(e && e[0] == '0'); }();
  // 256-row units give up the accumulator double-buffering: only worth it when a unit runs many pixel blocks
  w.msub = (allow256 && Cp % 256 == 0 && w.num_pb >= 64) ? 2 : 1;
  w.m_tiles = ceil_div(Cp, 128 * w.msub);
  w.chunks_s = ceil_div(Cs, 64);
  w.taps = taps;
  const int num_slabs = taps * w.chunks_s;
  w.slabs_per_tile = (num_slabs % 4 == 0) ? 4 : ((num_slabs % 2 == 0) ? 2 : 1);
  w.n_tiles = num_slabs / w.slabs_per_tile;
  choose_splits(w.m_tiles * w.n_tiles, w.num_pb, static_cast<double>(taps) * Cp * Cs * sizeof(float), w.msub, w.splits,
                w.pb_per_split);
  return w;
}

static int launch_wgrad(const GemmMaps& maps, const WgradGeom& w, const Tap* taps, int B, int H, int W, int Cp, int Cs,
                        float* dW, void* ws, size_t ws_bytes, float alpha, const float* alpha_dev, float beta,
                        cudaStream_t st, int Cs_out = -1, bool native = false) {
  if (Cs_out < 0) Cs_out = Cs;
  // native: rows of Cs_out floats per (p, tap).  Column padding (Cs_out < Cs) is only meaningful for one tap (the
  // ragged plain GEMM), needs an even Cs_out (8-byte aligned rows) and the direct epilogue (the native reduce kernel
  // walks unpadded float4 rows).
  if (native && w.taps != 1 && (Cs_out != Cs || Cs % 4 != 0)) {
    set_error("native weight-gradient layout needs Cs %% 4 == 0 and no column padding (Cs=%d)", Cs);
    return RG_EINVAL;
  }
  if (native && w.taps == 1) {     // one tap: native == dense row-major, so falling back to the reduce path is safe
    if (Cs_out != Cs && (Cs_out % 2 != 0 || w.splits != 1)) native = false;
    if (Cs_out == Cs && Cs % 4 != 0) native = false;
  }
  int rc = ensure_attrs();
  if (rc) return rc;
  // measured on B200: inside the full training step the direct epilogue LOSES ~1.2 ms/step to the split-K + reduce
  // path (16-byte read-modify-write pieces at a 64-byte stride when accumulating), so it is opt-in for experiments
  static const bool allow_direct = [] { const char* e = getenv("RG_WGRAD_DIRECT"); return e && e[0] == '1'; }();
  // direct 16-byte stores at a 64-byte stride pay off while dW stays L2-sized; the 268 MB projection gradient does not
  const bool native_direct = native && w.splits == 1;
  const bool direct = native_direct ||
                      (!native && allow_direct && (w.splits == 1 && w.taps == 16 && w.slabs_per_tile == 4 && Cs_out == Cs) &&
                       static_cast<size_t>(Cp) * Cs * 64 <= (160u << 20));
  const size_t need = static_cast<size_t>(w.splits) * w.taps * Cp * Cs * sizeof(float);
  if (!direct && (ws_bytes < need || ws == nullptr)) {
    set_error("wgrad workspace too small: need %zu bytes, have %zu", need, ws_bytes);
    return RG_EWORKSPACE;
  }
  WgradArgs a;
  memset(&a, 0, sizeof(a));
  a.nB = B; a.H = H; a.W = W;
  a.bw = w.g.bw; a.bh = w.g.bh; a.bb = w.g.bb;
  a.tw = w.g.tw; a.th = w.g.th; a.tb = w.g.tb;
  a.num_pb = w.num_pb; a.splits = w.splits; a.pb_per_split = w.pb_per_split;
  a.m_tiles = w.m_tiles; a.n_tiles = w.n_tiles; a.slabs_per_tile = w.slabs_per_tile; a.chunks_s = w.chunks_s;
  a.Cp = Cp; a.Cs = Cs; a.num_taps = w.taps;
  for (int t = 0; t < w.taps; ++t) a.taps[t] = taps[t];
  a.ws = static_cast<float*>(ws);
  a.msub = w.msub;
  a.ld_out = Cs_out;
  a.n_out = Cs_out;
  if (direct) {
    if (native_direct) a.native_out = dW;
    else a.direct_out = dW;
    a.alpha = alpha;
    a.alpha_dev = alpha_dev;
    a.beta = beta;
  }
  const int units = w.m_tiles * w.n_tiles * w.splits;
  const int grid = std::min(units, num_sms());
  RG_CUDA(launch_pdl(gemm_wgrad_kernel, dim3(grid), dim3(kGemmThreads), kWgradSmemBytes, st, 1, maps, a));
  RG_LAUNCH_CHECK("gemm_wgrad_kernel");
  if (direct) return 0;
  if (native) {
    const size_t n4 = static_cast<size_t>(w.taps) * Cp * (Cs / 4);
    wgrad_reduce_native_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, st>>>(a.ws, dW, w.splits, w.taps, Cp,
                                                                                         Cs, alpha, alpha_dev, beta);
    RG_LAUNCH_CHECK("wgrad_reduce_native_kernel");
    return 0;
  }
  const size_t n = static_cast<size_t>(Cp) * Cs_out;
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  if (w.taps == 16)
    wgrad_reduce_kernel<16, false><<<blocks, 256, 0, st>>>(a.ws, dW, w.splits, Cp, Cs, Cs_out, alpha, alpha_dev, beta);
  else if (w.taps == 9)
    wgrad_reduce_kernel<9, false><<<blocks, 256, 0, st>>>(a.ws, dW, w.splits, Cp, Cs, Cs_out, alpha, alpha_dev, beta);
  else if (Cs_out != Cs)
    wgrad_reduce_kernel<1, true><<<blocks, 256, 0, st>>>(a.ws, dW, w.splits, Cp, Cs, Cs_out, alpha, alpha_dev, beta);
  else
    wgrad_reduce_kernel<1, false><<<blocks, 256, 0, st>>>(a.ws, dW, w.splits, Cp, Cs, Cs_out, alpha, alpha_dev, beta);
  RG_LAUNCH_CHECK("wgrad_reduce_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------ pack kernels
// W[p][s][16] fp32 -> w_down[p][tap*Cs + s] bf16.  One block per (p, 64-wide s chunk).
__global__ void pack_down_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int Cp, int Cs) {
  __shared__ float sm[64 * 17];
  const int p = blockIdx.y;
  const int s0 = blockIdx.x * 64;
  const int ns = min(64, Cs - s0);
  const float* src = W + (static_cast<size_t>(p) * Cs + s0) * 16;
  for (int e = threadIdx.x; e < ns * 16; e += blockDim.x) sm[(e >> 4) * 17 + (e & 15)] = src[e];
  __syncthreads();
  for (int e = threadIdx.x; e < ns * 16; e += blockDim.x) {
    const int tap = e / ns, sl = e - tap * ns;
    out[static_cast<size_t>(p) * 16 * Cs + static_cast<size_t>(tap) * Cs + s0 + sl] = __float2bfloat16(sm[sl * 17 + tap]);
  }
}
// W[p][s][16] fp32 -> w_up[phase][s][t*Cp + p] bf16 (rows s >= Cs are left untouched: caller zero-fills once).
// One block per (32-wide p chunk, 16-wide s chunk).
__global__ void pack_up_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int Cp, int Cs,
                               int Cs_pad) {
  __shared__ float sm[32][257];
  const int p0 = blockIdx.y * 32, s0 = blockIdx.x * 16;
  const int np = min(32, Cp - p0), ns = min(16, Cs - s0);
  for (int e = threadIdx.x; e < np * ns * 16; e += blockDim.x) {
    const int pl = e / (ns * 16), off = e - pl * ns * 16;
    sm[pl][off] = W[(static_cast<size_t>(p0 + pl) * Cs + s0) * 16 + off];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < ns * 16 * 32; e += blockDim.x) {
    const int pl = e & 31;
    const int so = e >> 5;              // (s_local, tap16)
    if (pl >= np) continue;
    const int sl = so >> 4, tap16 = so & 15;
    const int kh = tap16 >> 2, kw = tap16 & 3;
    // y = 2i - 1 + kh: kh=1,3 feed even rows (taps 0,1); kh=2,0 feed odd rows (taps 0,1)
    const int rh = (kh == 1 || kh == 3) ? 0 : 1, th = (kh == 1 || kh == 2) ? 0 : 1;
    const int rw = (kw == 1 || kw == 3) ? 0 : 1, tw = (kw == 1 || kw == 2) ? 0 : 1;
    const int phase = rh * 2 + rw, t = th * 2 + tw;
    out[(static_cast<size_t>(phase) * Cs_pad + s0 + sl) * (4 * static_cast<size_t>(Cp)) + static_cast<size_t>(t) * Cp + p0 + pl] =
        __float2bfloat16(sm[pl][sl * 16 + tap16]);
  }
}
// w_down[p][tap16*Cs + s] bf16 -> w_up[phase][s][t*Cp + p] bf16: per kernel position a [p][s] -> [s][p] transpose.
// grid (ceil(Cs/32), ceil(Cp/32), 16), block 256.
__global__ void pack_up_from_down_kernel(const __nv_bfloat16* __restrict__ wd, __nv_bfloat16* __restrict__ out, int Cp,
                                         int Cs, int Cs_pad) {
  __shared__ __nv_bfloat16 sm[32][34];
  const int tap16 = blockIdx.z;
  const int p0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (p0 + r < Cp && s0 + tx < Cs)
      sm[r][tx] = wd[static_cast<size_t>(p0 + r) * 16 * Cs + static_cast<size_t>(tap16) * Cs + s0 + tx];
  __syncthreads();
  const int kh = tap16 >> 2, kw = tap16 & 3;
  const int rh = (kh == 1 || kh == 3) ? 0 : 1, th = (kh == 1 || kh == 2) ? 0 : 1;
  const int rw = (kw == 1 || kw == 3) ? 0 : 1, tw = (kw == 1 || kw == 2) ? 0 : 1;
  const int phase = rh * 2 + rw, t = th * 2 + tw;
  for (int r = ty; r < 32; r += 8) {
    const int sidx = s0 + r;
    if (sidx < Cs && p0 + tx < Cp)
      out[(static_cast<size_t>(phase) * Cs_pad + sidx) * (4 * static_cast<size_t>(Cp)) + static_cast<size_t>(t) * Cp + p0 + tx] =
          sm[tx][r];
  }
}
// Merged-phase operand of rg_conv_up (Cs == 64, CTA pairs).  The nine input shifts, centre first; shift i feeds
// `nslots` consecutive phase slabs starting at phase col0 (phase = 2*rh + rw of the output pixel parity); a slot the shift
// does not feed holds zeros so one MMA of N = 64 * nslots covers the span.
struct Up9Tap {
  int8_t dh, dw, nslots, col0;
  int8_t phase[4];          // phase of each slot, -1 = zero slab
};
struct Up9Table {
  Up9Tap t[9];
};
static Up9Table up9_table() {
  static const Up9Table tab = {{
      {0, 0, 4, 0, {0, 1, 2, 3}},
      {-1, 0, 2, 0, {0, 1, -1, -1}},      // even output rows, both column parities
      {1, 0, 2, 2, {2, 3, -1, -1}},       // odd output rows
      {0, -1, 3, 0, {0, -1, 2, -1}},      // even output columns: phases 0 and 2, zero slab between
      {0, 1, 3, 1, {1, -1, 3, -1}},       // odd output columns: phases 1 and 3
      {-1, -1, 1, 0, {0, -1, -1, -1}},
      {-1, 1, 1, 1, {1, -1, -1, -1}},
      {1, -1, 1, 2, {2, -1, -1, -1}},
      {1, 1, 1, 3, {3, -1, -1, -1}},
  }};
  return tab;
}
// out[shift i][chunk][CTA half][128 rows][64 k] bf16 from w_down[p][tap16*Cs + s] (K-major rows [column][p]; the first
// 32 * nslots rows of a half are used)
__global__ void pack_up9_kernel(const __nv_bfloat16* __restrict__ wd, __nv_bfloat16* __restrict__ out, int Cp, int Cs,
                                const Up9Table tab) {
  const int chunks = Cp / 64;
  const size_t total = static_cast<size_t>(9) * chunks * 2 * 4 * 32 * 64;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = static_cast<int>(idx & 63);
  const int n = static_cast<int>((idx >> 6) & 31);
  const int q = static_cast<int>((idx >> 11) & 3);
  const int half = static_cast<int>((idx >> 13) & 1);
  const int rest = static_cast<int>(idx >> 14);
  const int chunk = rest % chunks, ti = rest / chunks;
  const Up9Tap t = tab.t[ti];
  // cta_group::2 splits the N columns of ONE instruction between the CTAs: with N = 64 * nslots the leader supplies
  // columns [0, N/2), the peer [N/2, N).  Row r of this CTA's box is therefore column c = half * N/2 + r of the span:
  // slot c / 64 (a phase or a zero slab), output channel c % 64.
  const int r = q * 32 + n;
  const int half_cols = t.nslots * 32;
  __nv_bfloat16 o = __float2bfloat16(0.0f);
  if (r < half_cols) {
    const int c = half * half_cols + r;
    const int ph = t.phase[c >> 6];
    if (ph >= 0) {
      const int rh = ph >> 1, rw = ph & 1;
      // kernel index for (parity r, shift d): r = 0: d = 0 -> 1, d = -1 -> 3;  r = 1: d = 0 -> 2, d = +1 -> 0
      const int kh = rh == 0 ? (t.dh == 0 ? 1 : 3) : (t.dh == 0 ? 2 : 0);
      const int kw = rw == 0 ? (t.dw == 0 ? 1 : 3) : (t.dw == 0 ? 2 : 0);
      const int prow = chunk * 64 + k, sidx = c & 63;
      o = wd[static_cast<size_t>(prow) * 16 * Cs + static_cast<size_t>(kh * 4 + kw) * Cs + sidx];
    }
  }
  out[idx] = o;
}
// W[E][C0*16] fp32 -> out[(tap*C0 + co)][E] bf16 : 32x32 tile transpose
__global__ void pack_proj_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int E, int C0) {
  __shared__ float sm[32][33];
  const int e0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int N = C0 * 16;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 256 threads: 8 rows per pass
  for (int r = ty; r < 32; r += 8)
    if (e0 + r < E && n0 + tx < N) sm[r][tx] = W[static_cast<size_t>(e0 + r) * N + n0 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r;
    if (n < N && e0 + tx < E) {
      const int co = n >> 4, tap = n & 15;
      out[(static_cast<size_t>(tap) * C0 + co) * E + e0 + tx] = __float2bfloat16(sm[tx][r]);
    }
  }
}
// W[Cp][Cimg][16] -> w_col[Cp][64], k = tap*4 + c
__global__ void pack_edge_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int Cp, int Cimg) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cp * 64) return;
  const int p = idx >> 6, k = idx & 63, tap = k >> 2, c = k & 3;
  const float v = c < Cimg ? W[(static_cast<size_t>(p) * Cimg + c) * 16 + tap] : 0.0f;
  out[idx] = __float2bfloat16(v);
}
__global__ void pack_conv3_kernel(const float* __restrict__ W, __nv_bfloat16* __restrict__ out, int Cout, int Cin,
                                  int rows) {
  const size_t n = static_cast<size_t>(rows) * 9 * Cin;
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int co = static_cast<int>(i / (9 * Cin));
  const int k = static_cast<int>(i - static_cast<size_t>(co) * 9 * Cin);
  const int tap = k / Cin, c = k - tap * Cin;
  out[i] = __float2bfloat16(co < Cout ? W[(static_cast<size_t>(co) * Cin + c) * 9 + tap] : 0.0f);
}
__global__ void cast_pad_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows, int cols,
                                int cols_pad) {
  const size_t n = static_cast<size_t>(rows) * cols_pad;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / cols_pad;
    const int c = static_cast<int>(i - r * cols_pad);
    dst[i] = __float2bfloat16(c < cols ? src[r * cols + c] : 0.0f);
  }
}

}  // namespace rg

using namespace rg;

extern "C" {

int rg_version(void) { return 100; }
long long rg_launch_count(void) { return rg::launch_count(); }
const char* rg_last_error(void) { return g_err; }

int rg_check_device(void) {
  int dev = 0;
  RG_CUDA(cudaGetDevice(&dev));
  int major = 0;
  RG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("device compute capability %d.x is not sm_100 (B200)", major);
    return RG_EARCH;
  }
  return 0;
}

void rg_debug_set_prof(long long* buf) { g_prof = buf; }
size_t rg_stats_ws_bytes(int C) { return C > 0 ? stats_ws_floats(C) * sizeof(float) : 0; }
int rg_stats_parts(void) { return num_sms(); }

int rg_pack_link(const float* W, void* w_down, void* w_up, int Cp, int Cs, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(W && Cp > 0 && Cs > 0, "rg_pack_link: bad arguments");
  if (w_down) {
    dim3 grid(ceil_div(Cs, 64), Cp);
    pack_down_kernel<<<grid, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(w_down), Cp, Cs);
    RG_LAUNCH_CHECK("pack_down_kernel");
  }
  if (w_up) {
    const int Cs_pad = std::max(16, (Cs + 15) / 16 * 16);
    dim3 grid(ceil_div(Cs, 16), ceil_div(Cp, 32));
    pack_up_kernel<<<grid, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(w_up), Cp, Cs, Cs_pad);
    RG_LAUNCH_CHECK("pack_up_kernel");
  }
  return 0;
}

int rg_pack_up_from_down(const void* w_down, void* w_up, int Cp, int Cs, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(w_down && w_up && Cp > 0 && Cs > 0, "rg_pack_up_from_down: bad arguments");
  const int Cs_pad = std::max(16, (Cs + 15) / 16 * 16);
  dim3 grid(ceil_div(Cs, 32), ceil_div(Cp, 32), 16);
  pack_up_from_down_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(w_down),
                                                 static_cast<__nv_bfloat16*>(w_up), Cp, Cs, Cs_pad);
  RG_LAUNCH_CHECK("pack_up_from_down_kernel");
  return 0;
}

size_t rg_up9_elems(int Cp) { return Cp > 0 && Cp % 64 == 0 ? static_cast<size_t>(9) * (Cp / 64) * 2 * 4 * 32 * 64 : 0; }

int rg_pack_up9_from_down(const void* w_down, void* w_up9, int Cp, int Cs, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(w_down && w_up9 && Cp > 0 && Cp % 64 == 0 && Cs == 64, "rg_pack_up9_from_down: need Cs == 64, Cp %% 64 == 0");
  const size_t total = rg_up9_elems(Cp);
  pack_up9_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(w_down), static_cast<__nv_bfloat16*>(w_up9), Cp, Cs, up9_table());
  RG_LAUNCH_CHECK("pack_up9_kernel");
  return 0;
}

int rg_pack_proj(const float* W, void* w_proj, int E, int C0, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(W && w_proj && E > 0 && C0 > 0, "rg_pack_proj: bad arguments");
  dim3 grid(ceil_div(C0 * 16, 32), ceil_div(E, 32));
  pack_proj_kernel<<<grid, 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(w_proj), E, C0);
  RG_LAUNCH_CHECK("pack_proj_kernel");
  return 0;
}

int rg_pack_edge(const float* W, void* w_col, int Cp, int Cimg, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(W && w_col && Cp > 0 && Cimg > 0 && Cimg <= 4, "rg_pack_edge: bad arguments");
  pack_edge_kernel<<<ceil_div(Cp * 64, 256), 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(w_col), Cp, Cimg);
  RG_LAUNCH_CHECK("pack_edge_kernel");
  return 0;
}

int rg_cast_pad_bf16(const float* src, void* dst, int rows, int cols, int cols_pad, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(src && dst && rows > 0 && cols > 0 && cols_pad >= cols, "rg_cast_pad_bf16: bad arguments");
  const size_t n = static_cast<size_t>(rows) * cols_pad;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(num_sms()) * 16));
  cast_pad_kernel<<<grid, 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), rows, cols, cols_pad);
  RG_LAUNCH_CHECK("cast_pad_kernel");
  return 0;
}

int rg_conv_down(const void* hi, const void* w_down, void* lo, int B, int H, int W, int Cs, int Cp, float* stats_ws,
                 const rg_epilogue_aux* aux, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(hi && w_down && lo, "rg_conv_down: null pointer");
  RG_CHECK_ARG(B > 0 && is_pow2(H) && is_pow2(W), "rg_conv_down: H, W must be powers of two (got %d x %d)", H, W);
  RG_CHECK_ARG(Cs % 64 == 0 && Cp % 8 == 0 && Cp >= 16, "rg_conv_down: need Cs %% 64 == 0, Cp %% 8 == 0 (Cs=%d Cp=%d)", Cs, Cp);
  GemmMaps maps;
  FwdArgs a;
  fill_common(a, B, H, W);
  int rc = encode_parity_maps(maps, hi, B, H, W, Cs, a.bw, a.bh, a.bb);
  if (rc) return rc;
  a.num_taps = 16;
  a.chunks = Cs / 64;
  for (int kh = 0; kh < 4; ++kh)
    for (int kw = 0; kw < 4; ++kw) {
      Tap t;
      t.map = static_cast<int8_t>(kDownPar[kh] * 2 + kDownPar[kw]);
      t.dh = static_cast<int8_t>(kDownOff[kh]);
      t.dw = static_cast<int8_t>(kDownOff[kw]);
      t.wtap = static_cast<int8_t>(kh * 4 + kw);
      a.taps[0][kh * 4 + kw] = t;
    }
  a.n_total = Cp;
  a.block_n = pick_block_n(Cp, a.m_tiles, pair_allowed(a, OUT_BF16_NHWC) ? 2 : 1);
  a.n_tiles = ceil_div(Cp, a.block_n);
  a.b_phase_rows = 0;
  const int cg = pick_cg(a, OUT_BF16_NHWC);
  rc = encode_map_2d(&maps.b, w_down, 16ull * Cs, Cp, 16ull * Cs, 64, a.block_n / cg);
  if (rc) return rc;
  a.out = lo;
  a.OH = H; a.OW = W; a.OC = Cp;
  a.n_valid = Cp;
  return launch_fwd(maps, a, OUT_BF16_NHWC, cg, stats_ws, st, aux);
}

// all four output phases of an M tile in one tile (weights packed by rg_pack_up9_from_down)
static int conv_up_merged(const void* lo, const void* w9, void* out, int B, int H, int W, int Cp, int Cs,
                          float* stats_ws, const rg_epilogue_aux* aux, cudaStream_t st) {
  RG_CHECK_ARG(Cs == 64 && Cp % 64 == 0, "rg_conv_up (merged phases): need Cs == 64, Cp %% 64 == 0");
  GemmMaps maps;
  FwdArgs a;
  fill_common(a, B, H, W);
  RG_CHECK_ARG(a.m_tiles >= 2 && pair_allowed(a, OUT_BF16_NHWC),
               "rg_conv_up (merged phases) needs at least two M tiles (CTA pairs); use the w_up operand for tiny inputs");
  int rc = encode_map_4d(&maps.a[0], lo, Cp, W, H, B, Cp, 1ull * W * Cp, 1ull * H * W * Cp, 64, a.bw, a.bh, a.bb);
  if (rc) return rc;
  const int chunks = Cp / 64;
  const uint64_t rows9 = 9ull * chunks * 2 * 4 * 32;      // w_up9 as a [rows][64] matrix
  rc = encode_map_2d(&maps.b, w9, 64, rows9, 64, 64, 4 * 32);          // box heights: 4, 2, 1, 3 slots of 32 rows
  if (rc) return rc;
  rc = encode_map_2d(&maps.a[1], w9, 64, rows9, 64, 64, 2 * 32);
  if (rc) return rc;
  rc = encode_map_2d(&maps.a[2], w9, 64, rows9, 64, 64, 1 * 32);
  if (rc) return rc;
  rc = encode_map_2d(&maps.a[3], w9, 64, rows9, 64, 64, 3 * 32);
  if (rc) return rc;
  a.merged = 1;
  a.num_taps = 9;
  a.chunks = chunks;
  a.num_phases = 1;
  const Up9Table tab = up9_table();
  for (int i = 0; i < 9; ++i) {
    Tap t = {0, tab.t[i].dh, tab.t[i].dw, 0};
    a.taps[0][i] = t;
    a.mg_nph[i] = tab.t[i].nslots;
    a.mg_col[i] = tab.t[i].col0;
  }
  for (int ph = 0; ph < 4; ++ph) {
    a.oy[ph] = static_cast<int8_t>(ph >> 1);
    a.ox[ph] = static_cast<int8_t>(ph & 1);
  }
  a.n_total = Cs;
  a.block_n = 256;               // TMEM columns of one tile: four 64-column phase slabs
  a.n_tiles = 1;
  a.out = out;
  a.OH = 2 * H; a.OW = 2 * W; a.OC = Cs;
  a.sy = a.sx = 2;
  a.n_valid = Cs;
  return launch_fwd(maps, a, OUT_BF16_NHWC, 2, stats_ws, st, aux);
}

static int conv_up_common(const void* lo, const void* w, void* out, const float* bias, int act_tanh, int B, int H,
                          int W, int Cp, int Cs, int out_kind, bool w_is_down, float* stats_ws, cudaStream_t st,
                          const rg_epilogue_aux* aux = nullptr) {
  RG_CHECK_ARG(lo && w && out, "rg_conv_up: null pointer");
  RG_CHECK_ARG(B > 0 && is_pow2(H) && is_pow2(W), "rg_conv_up: H, W must be powers of two (got %d x %d)", H, W);
  RG_CHECK_ARG(Cp % 64 == 0, "rg_conv_up: need Cp %% 64 == 0 (Cp=%d)", Cp);
  const int Cs_pad = std::max(16, (Cs + 15) / 16 * 16);
  GemmMaps maps;
  FwdArgs a;
  fill_common(a, B, H, W);
  int rc = encode_map_4d(&maps.a[0], lo, Cp, W, H, B, Cp, 1ull * W * Cp, 1ull * H * W * Cp, 64, a.bw, a.bh, a.bb);
  if (rc) return rc;
  maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
  a.num_taps = 4;
  a.chunks = Cp / 64;
  a.num_phases = 4;
  for (int rh = 0; rh < 2; ++rh)
    for (int rw = 0; rw < 2; ++rw) {
      const int ph = rh * 2 + rw;
      a.oy[ph] = static_cast<int8_t>(rh);
      a.ox[ph] = static_cast<int8_t>(rw);
      for (int th = 0; th < 2; ++th)
        for (int tw = 0; tw < 2; ++tw) {
          Tap t;
          t.map = 0;
          t.dh = static_cast<int8_t>(kUpOff[rh][th]);
          t.dw = static_cast<int8_t>(kUpOff[rw][tw]);
          t.wtap = static_cast<int8_t>(kUpK[rh][th] * 4 + kUpK[rw][tw]);
          a.taps[ph][th * 2 + tw] = t;
        }
    }
  a.n_total = Cs_pad;
  a.block_n = pick_block_n(Cs_pad, a.m_tiles * 4, pair_allowed(a, out_kind) ? 2 : 1);
  a.n_tiles = ceil_div(Cs_pad, a.block_n);
  a.b_phase_rows = Cs_pad;
  a.b_mn = w_is_down ? 1 : 0;
  const int cg = pick_cg(a, out_kind);
  if (w_is_down) {
    // B is read MN-major straight from w_down[Cp][16*Cs]: 64x64 slabs at column tap*Cs + n, row p
    a.b_tap_cols = Cs;
    rc = encode_map_2d(&maps.b, w, 16ull * Cs, Cp, 16ull * Cs, 64, 64);
  } else {
    rc = encode_map_2d(&maps.b, w, 4ull * Cp, 4ull * Cs_pad, 4ull * Cp, 64, a.block_n / cg);
  }
  if (rc) return rc;
  a.out = out;
  a.OH = 2 * H; a.OW = 2 * W; a.OC = Cs;
  a.sy = a.sx = 2;
  a.n_valid = Cs;
  a.col_shift = bias;
  a.act_tanh = act_tanh;
  return launch_fwd(maps, a, out_kind, cg, stats_ws, st, aux);
}

int rg_conv_up(const void* lo, const void* w, int w_is_down, void* hi, int B, int H, int W, int Cp, int Cs,
               float* stats_ws, const rg_epilogue_aux* aux, rg_stream_t st_) {
  if (w_is_down == 2) {
    RG_CHECK_ARG(lo && w && hi && B > 0 && is_pow2(H) && is_pow2(W), "rg_conv_up: bad arguments");
    return conv_up_merged(lo, w, hi, B, H, W, Cp, Cs, stats_ws, aux, static_cast<cudaStream_t>(st_));
  }
  RG_CHECK_ARG(Cs % (w_is_down ? 64 : 16) == 0,
               "rg_conv_up: need Cs %% 64 == 0 with w_down, %% 16 with w_up (Cs=%d); use rg_conv_up_img for images", Cs);
  return conv_up_common(lo, w, hi, nullptr, 0, B, H, W, Cp, Cs, OUT_BF16_NHWC, w_is_down != 0, stats_ws,
                        static_cast<cudaStream_t>(st_), aux);
}

int rg_conv_up_img(const void* lo, const void* w_up, float* img, const float* bias, int act_tanh, int B, int H, int W,
                   int Cp, int Cimg, rg_stream_t st_) {
  RG_CHECK_ARG(Cimg >= 1 && Cimg <= 8, "rg_conv_up_img: 1..8 image channels supported (got %d)", Cimg);
  return conv_up_common(lo, w_up, img, bias, act_tanh, B, H, W, Cp, Cimg, OUT_F32_NCHW, false, nullptr,
                        static_cast<cudaStream_t>(st_));
}

static int gemm_plain(const void* A, int lda, const void* Bw, int ldb, bool b_is_kn, void* C, int M, int N, int K,
                      int ldc, const float* col_scale, const float* col_shift, float slope, int out_f32,
                      cudaStream_t st, const char* name, float* stats_ws = nullptr,
                      const rg_epilogue_aux* aux = nullptr) {
  RG_CHECK_ARG(A && Bw && C, "%s: null pointer", name);
  RG_CHECK_ARG(M > 0 && N > 0 && K > 0 && lda >= K && lda % 8 == 0 && ldb % 8 == 0,
               "%s: need lda >= K and lda, ldb multiples of 8 elements (M=%d N=%d K=%d lda=%d ldb=%d)", name, M, N, K,
               lda, ldb);
  RG_CHECK_ARG(ldc >= N && (out_f32 ? ldc % 4 == 0 : ldc % 8 == 0), "%s: bad ldc %d", name, ldc);
  GemmMaps maps;
  FwdArgs a;
  fill_common(a, M, 1, 1);
  // K need not be a multiple of 64: the tensor maps carry the true K and TMA zero-fills the tail of the last k-block
  static const bool a2d = [] { const char* e = getenv("RG_A2D"); return !(e && e[0] == '0'); }();
  int rc;
  if (a2d) {
    rc = encode_map_2d(&maps.a[0], A, K, M, lda, 64, kBlockM);
    a.a_2d = 1;
  } else {
    rc = encode_map_4d(&maps.a[0], A, K, 1, 1, M, lda, lda, lda, 64, 1, 1, kBlockM);
  }
  if (rc) return rc;
  maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
  a.num_taps = 1;
  a.chunks = ceil_div(K, 64);
  Tap t = {0, 0, 0, 0};
  a.taps[0][0] = t;
  const int out_tanh = (out_f32 & 2) ? 1 : 0;     // out_f32 is a flag word: bit 0 fp32 output, bit 1 tanh (fp32 only)
  out_f32 &= 1;
  RG_CHECK_ARG(!out_tanh || out_f32, "%s: the tanh epilogue needs an fp32 output", name);
  const int out_kind = out_f32 ? OUT_F32_NHWC : OUT_BF16_NHWC;
  a.n_total = N;
  a.block_n = pick_block_n(N, a.m_tiles, pair_allowed(a, out_kind) ? 2 : 1);
  if (b_is_kn && a.block_n < 64) a.block_n = 64;
  a.n_tiles = ceil_div(N, a.block_n);
  a.b_mn = b_is_kn ? 1 : 0;
  const int cg = pick_cg(a, out_kind);
  if (b_is_kn) {
    // Bw is [K][N] row-major (e.g. an nn.Linear weight [out=K][in=N] used for its input gradient): MN-major B slabs
    a.b_tap_cols = 0;
    rc = encode_map_2d(&maps.b, Bw, N, K, ldb, 64, 64);
  } else {
    rc = encode_map_2d(&maps.b, Bw, K, N, ldb, 64, a.block_n / cg);
  }
  if (rc) return rc;
  a.out = C;
  a.OH = 1; a.OW = 1; a.OC = ldc;
  a.n_valid = N;
  a.col_scale = col_scale;
  a.col_shift = col_shift;
  a.slope = slope;
  a.act_tanh = out_tanh;
  return launch_fwd(maps, a, out_kind, cg, stats_ws, st, aux);
}

int rg_gemm_nt(const void* A, const void* Bw, void* C, int M, int N, int K, int ldc, const float* col_scale,
               const float* col_shift, float slope, int out_f32, rg_stream_t st_) {
  return gemm_plain(A, K, Bw, K, false, C, M, N, K, ldc, col_scale, col_shift, slope, out_f32,
                    static_cast<cudaStream_t>(st_), "rg_gemm_nt");
}

int rg_gemm_nt_ld(const void* A, int lda, const void* Bw, int ldb, void* C, int M, int N, int K, int ldc,
                  const float* col_scale, const float* col_shift, float slope, int out_f32, rg_stream_t st_) {
  return gemm_plain(A, lda, Bw, ldb, false, C, M, N, K, ldc, col_scale, col_shift, slope, out_f32,
                    static_cast<cudaStream_t>(st_), "rg_gemm_nt_ld");
}

int rg_gemm_nt_bwd(const void* A, int lda, const void* Bw, int ldb, void* C, int M, int N, int K, int ldc,
                   float* stats_ws, const rg_epilogue_aux* aux, rg_stream_t st_) {
  return gemm_plain(A, lda, Bw, ldb, false, C, M, N, K, ldc, nullptr, nullptr, 1.0f, 0, static_cast<cudaStream_t>(st_),
                    "rg_gemm_nt_bwd", stats_ws, aux);
}

int rg_gemm_nn(const void* A, int lda, const void* Bw, int ldb, void* C, int M, int N, int K, int ldc,
               const float* col_scale, const float* col_shift, float slope, int out_f32, rg_stream_t st_) {
  return gemm_plain(A, lda, Bw, ldb, true, C, M, N, K, ldc, col_scale, col_shift, slope, out_f32,
                    static_cast<cudaStream_t>(st_), "rg_gemm_nn");
}

// ------------------------------------------------------------------------------------------------ 3x3 stride-1 conv
// u: reflect-padded, 2x-upsampled activation bf16 [B][Ho+2][Wo+2][Cin]; w3: bf16 [Cout_pad][9*Cin], k = tap*Cin + c.
static int conv3_common(const void* u, const void* w3, void* out, const float* bias, int B, int Ho, int Wo, int Cin,
                        int Cout, int out_kind, float* stats_ws, cudaStream_t st) {
  RG_CHECK_ARG(u && w3 && out, "rg_conv3x3: null pointer");
  RG_CHECK_ARG(B > 0 && is_pow2(Ho) && is_pow2(Wo) && Cin % 64 == 0, "rg_conv3x3: need power-of-two Ho, Wo and Cin %% 64 == 0");
  const int Cout_pad = std::max(16, (Cout + 15) / 16 * 16);
  GemmMaps maps;
  FwdArgs a;
  fill_common(a, B, Ho, Wo);
  const uint64_t Wp = Wo + 2, Hp = Ho + 2;
  int rc = encode_map_4d(&maps.a[0], u, Cin, Wp, Hp, B, Cin, Wp * Cin, Hp * Wp * Cin, 64, a.bw, a.bh, a.bb);
  if (rc) return rc;
  maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
  a.num_taps = 9;
  a.chunks = Cin / 64;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      Tap t = {0, static_cast<int8_t>(kh), static_cast<int8_t>(kw), static_cast<int8_t>(kh * 3 + kw)};
      a.taps[0][kh * 3 + kw] = t;
    }
  a.n_total = Cout_pad;
  a.block_n = pick_block_n(Cout_pad, a.m_tiles, pair_allowed(a, out_kind) ? 2 : 1);
  a.n_tiles = ceil_div(Cout_pad, a.block_n);
  const int cg = pick_cg(a, out_kind);
  rc = encode_map_2d(&maps.b, w3, 9ull * Cin, Cout_pad, 9ull * Cin, 64, a.block_n / cg);
  if (rc) return rc;
  a.out = out;
  a.OH = Ho; a.OW = Wo; a.OC = Cout;
  a.n_valid = Cout;
  a.col_shift = bias;
  return launch_fwd(maps, a, out_kind, cg, stats_ws, st);
}

int rg_conv3x3(const void* u, const void* w3, void* out, const float* bias, int B, int Ho, int Wo, int Cin, int Cout,
               float* stats_ws, rg_stream_t st) {
  RG_CHECK_ARG(Cout % 8 == 0, "rg_conv3x3: need Cout %% 8 == 0 (use rg_conv3x3_img for image channels)");
  return conv3_common(u, w3, out, bias, B, Ho, Wo, Cin, Cout, OUT_BF16_NHWC, stats_ws, static_cast<cudaStream_t>(st));
}

int rg_conv3x3_img(const void* u, const void* w3, float* img, const float* bias, int B, int Ho, int Wo, int Cin,
                   int Cimg, rg_stream_t st) {
  RG_CHECK_ARG(Cimg >= 1 && Cimg <= 8, "rg_conv3x3_img: 1..8 image channels supported");
  return conv3_common(u, w3, img, bias, B, Ho, Wo, Cin, Cimg, OUT_F32_NCHW, nullptr, static_cast<cudaStream_t>(st));
}

// du[b,y,x,ci] = sum_{kh,kw,co} da[b,y-kh,x-kw,co] * W[co,ci,kh,kw] on the padded (Ho+2)x(Wo+2) grid; the weights
// are read MN-major from the same w3 buffer.
int rg_conv3x3_dgrad(const void* da, const void* w3, void* du, int B, int Ho, int Wo, int Cin, int Cout,
                     rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(da && w3 && du, "rg_conv3x3_dgrad: null pointer");
  RG_CHECK_ARG(B > 0 && Cout % 64 == 0 && Cin % 64 == 0, "rg_conv3x3_dgrad: need Cin, Cout %% 64 == 0");
  GemmMaps maps;
  FwdArgs a;
  fill_common(a, B, Ho + 2, Wo + 2);
  int rc = encode_map_4d(&maps.a[0], da, Cout, Wo, Ho, B, Cout, 1ull * Wo * Cout, 1ull * Ho * Wo * Cout, 64, a.bw, a.bh,
                         a.bb);
  if (rc) return rc;
  maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
  a.num_taps = 9;
  a.chunks = Cout / 64;
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      Tap t = {0, static_cast<int8_t>(-kh), static_cast<int8_t>(-kw), static_cast<int8_t>(kh * 3 + kw)};
      a.taps[0][kh * 3 + kw] = t;
    }
  a.n_total = Cin;
  a.block_n = std::max(64, pick_block_n(Cin, a.m_tiles, pair_allowed(a, OUT_BF16_NHWC) ? 2 : 1));
  a.n_tiles = ceil_div(Cin, a.block_n);
  a.b_mn = 1;
  a.b_tap_cols = Cin;
  const int cg = pick_cg(a, OUT_BF16_NHWC);
  rc = encode_map_2d(&maps.b, w3, 9ull * Cin, Cout, 9ull * Cin, 64, 64);
  if (rc) return rc;
  a.out = du;
  a.OH = Ho + 2; a.OW = Wo + 2; a.OC = Cin;
  a.n_valid = Cin;
  return launch_fwd(maps, a, OUT_BF16_NHWC, cg, nullptr, st);
}

size_t rg_conv3x3_wgrad_ws_bytes(int B, int Ho, int Wo, int Cin, int Cout) {
  if (B <= 0 || Ho <= 0 || Wo <= 0 || Cin < 64 || Cout <= 0) return 0;
  WgradGeom w = wgrad_geom(B, Ho, Wo, Cout, Cin, 9);
  return static_cast<size_t>(w.splits) * 9 * Cout * Cin * sizeof(float);
}

// dW[co,ci,kh,kw] = sum_{b,y,x} da[b,y,x,co] * u[b,y+kh,x+kw,ci]
int rg_conv3x3_wgrad(const void* da, const void* u, float* dW, void* ws, size_t ws_bytes, int B, int Ho, int Wo,
                     int Cin, int Cout, float beta, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(da && u && dW, "rg_conv3x3_wgrad: null pointer");
  RG_CHECK_ARG(B > 0 && is_pow2(Ho) && is_pow2(Wo) && Cin % 64 == 0 && Cout % 8 == 0,
               "rg_conv3x3_wgrad: need power-of-two Ho, Wo, Cin %% 64 == 0, Cout %% 8 == 0");
  WgradGeom w = wgrad_geom(B, Ho, Wo, Cout, Cin, 9);
  GemmMaps maps;
  const uint64_t Wp = Wo + 2, Hp = Ho + 2;
  int rc = encode_map_4d(&maps.a[0], u, Cin, Wp, Hp, B, Cin, Wp * Cin, Hp * Wp * Cin, 64, w.g.bw, w.g.bh, w.g.bb);
  if (rc) return rc;
  maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
  rc = encode_map_4d(&maps.b, da, Cout, Wo, Ho, B, Cout, 1ull * Wo * Cout, 1ull * Ho * Wo * Cout, 64, w.g.bw, w.g.bh,
                     w.g.bb);
  if (rc) return rc;
  Tap taps[16];
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw) {
      Tap t = {0, static_cast<int8_t>(kh), static_cast<int8_t>(kw), static_cast<int8_t>(kh * 3 + kw)};
      taps[kh * 3 + kw] = t;
    }
  return launch_wgrad(maps, w, taps, B, Ho, Wo, Cout, Cin, dW, ws, ws_bytes, 1.0f, nullptr, beta, st);
}

// W[Cout][Cin][3][3] fp32 -> w3[Cout_pad][9*Cin] bf16 (rows >= Cout zero)
int rg_pack_conv3(const float* W, void* w3, int Cout, int Cin, int rows, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(W && w3 && rows >= Cout, "rg_pack_conv3: bad arguments");
  const size_t n = static_cast<size_t>(rows) * 9 * Cin;
  pack_conv3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(W, static_cast<__nv_bfloat16*>(w3), Cout,
                                                                            Cin, rows);
  RG_LAUNCH_CHECK("rg_pack_conv3");
  return 0;
}

size_t rg_conv_wgrad_ws_bytes(int B, int H, int W, int Cp, int Cs) {
  if (B <= 0 || H <= 0 || W <= 0 || Cp <= 0 || Cs < 64) return 0;
  WgradGeom w = wgrad_geom(B, H, W, Cp, Cs, 16);
  return static_cast<size_t>(w.splits) * 16 * Cp * Cs * sizeof(float);
}

int rg_conv_wgrad(const void* lo, const void* hi, float* dW, void* ws, size_t ws_bytes, int B, int H, int W, int Cp,
                  int Cs, float alpha, const float* alpha_dev, float beta, int native_layout, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(lo && hi && dW, "rg_conv_wgrad: null pointer");
  RG_CHECK_ARG(B > 0 && is_pow2(H) && is_pow2(W), "rg_conv_wgrad: H, W must be powers of two (got %d x %d)", H, W);
  RG_CHECK_ARG(Cs % 64 == 0 && Cp % 8 == 0, "rg_conv_wgrad: need Cs %% 64 == 0 and Cp %% 8 == 0 (Cs=%d Cp=%d)", Cs, Cp);
  WgradGeom w = wgrad_geom(B, H, W, Cp, Cs, 16);
  GemmMaps maps;
  int rc = encode_parity_maps(maps, hi, B, H, W, Cs, w.g.bw, w.g.bh, w.g.bb);
  if (rc) return rc;
  rc = encode_map_4d(&maps.b, lo, Cp, W, H, B, Cp, 1ull * W * Cp, 1ull * H * W * Cp, 64, w.g.bw, w.g.bh, w.g.bb);
  if (rc) return rc;
  Tap taps[16];
  for (int kh = 0; kh < 4; ++kh)
    for (int kw = 0; kw < 4; ++kw) {
      Tap t;
      t.map = static_cast<int8_t>(kDownPar[kh] * 2 + kDownPar[kw]);
      t.dh = static_cast<int8_t>(kDownOff[kh]);
      t.dw = static_cast<int8_t>(kDownOff[kw]);
      t.wtap = static_cast<int8_t>(kh * 4 + kw);
      taps[kh * 4 + kw] = t;
    }
  return launch_wgrad(maps, w, taps, B, H, W, Cp, Cs, dW, ws, ws_bytes, alpha, alpha_dev, beta, st, -1,
                      native_layout != 0);
}

size_t rg_proj_wgrad_ws_bytes(int B, int E, int C0) {
  if (B <= 0 || E <= 0 || C0 < 64) return 0;
  WgradGeom w = wgrad_geom(B, 1, 1, E, C0, 16);
  return static_cast<size_t>(w.splits) * 16 * E * C0 * sizeof(float);
}

int rg_proj_wgrad(const void* z, const void* da0, float* dW, void* ws, size_t ws_bytes, int B, int E, int C0,
                  float alpha, const float* alpha_dev, float beta, int native_layout, rg_stream_t st_) {
  cudaStream_t st = static_cast<cudaStream_t>(st_);
  RG_CHECK_ARG(z && da0 && dW, "rg_proj_wgrad: null pointer");
  RG_CHECK_ARG(B > 0 && E % 8 == 0 && C0 % 64 == 0, "rg_proj_wgrad: need E %% 8 == 0, C0 %% 64 == 0 (E=%d C0=%d)", E, C0);
  WgradGeom w = wgrad_geom(B, 1, 1, E, C0, 16);
  GemmMaps maps;
  // hi = da0[B][4][4][C0]; tap (kh,kw) reads pixel (kh,kw) of every sample
  int rc = encode_map_4d(&maps.a[0], da0, C0, 4, 4, B, C0, 4ull * C0, 16ull * C0, 64, 1, 1, 64);
  if (rc) return rc;
  maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
  rc = encode_map_4d(&maps.b, z, E, 1, 1, B, E, E, E, 64, 1, 1, 64);
  if (rc) return rc;
  Tap taps[16];
  for (int kh = 0; kh < 4; ++kh)
    for (int kw = 0; kw < 4; ++kw) {
      Tap t = {0, static_cast<int8_t>(kh), static_cast<int8_t>(kw), 0};
      taps[kh * 4 + kw] = t;
    }
  return launch_wgrad(maps, w, taps, B, 1, 1, E, C0, dW, ws, ws_bytes, alpha, alpha_dev, beta, st, -1,
                      native_layout != 0);
}

size_t rg_gemm_tn_ws_bytes(int R, int M, int N) {
  if (R <= 0 || M <= 0 || N < 1) return 0;
  const int Nw = (N + 3) / 4 * 4;
  WgradGeom w = wgrad_geom(R, 1, 1, M, Nw, 1);
  return static_cast<size_t>(w.splits) * M * Nw * sizeof(float);
}

static int gemm_tn_impl(const void* A, int lda, const void* Bm, int ldb, float* C, void* ws, size_t ws_bytes, int R,
                        int M, int N, float alpha, const float* alpha_dev, float beta, cudaStream_t st) {
  RG_CHECK_ARG(A && Bm && C, "rg_gemm_tn: null pointer");
  RG_CHECK_ARG(R > 0 && M > 0 && N > 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= M && ldb >= N,
               "rg_gemm_tn: need lda >= M, ldb >= N, both multiples of 8 (M=%d N=%d lda=%d ldb=%d)", M, N, lda, ldb);
  const int Nw = (N + 3) / 4 * 4;     // partials use a 16-byte aligned row stride; the final reduce writes N columns
  WgradGeom w = wgrad_geom(R, 1, 1, M, Nw, 1);
  GemmMaps maps;
  int rc = encode_map_4d(&maps.a[0], Bm, N, 1, 1, R, ldb, ldb, ldb, 64, 1, 1, 64);
  if (rc) return rc;
  maps.a[1] = maps.a[0]; maps.a[2] = maps.a[0]; maps.a[3] = maps.a[0];
  rc = encode_map_4d(&maps.b, A, M, 1, 1, R, lda, lda, lda, 64, 1, 1, 64);
  if (rc) return rc;
  Tap taps[1] = {{0, 0, 0, 0}};
  // With a single tap the native gradient layout [p][tap][s] IS the dense row-major C[M][N]: when N needs no column
  // padding the epilogue stores C directly (alpha, beta in place) whenever one unit covers all R rows -- every
  // nn.Linear weight gradient of the betaVAE step except the ragged 19198-column one -- instead of writing fp32
  // partials and copying them (1.2 ms of the 5.0 ms step, tools/vae_profile.py).
  return launch_wgrad(maps, w, taps, R, 1, 1, M, Nw, C, ws, ws_bytes, alpha, alpha_dev, beta, st, N, true);
}

int rg_gemm_tn(const void* A, const void* Bm, float* C, void* ws, size_t ws_bytes, int R, int M, int N, float alpha,
               const float* alpha_dev, float beta, rg_stream_t st_) {
  return gemm_tn_impl(A, M, Bm, N, C, ws, ws_bytes, R, M, N, alpha, alpha_dev, beta, static_cast<cudaStream_t>(st_));
}

int rg_gemm_tn_ld(const void* A, int lda, const void* Bm, int ldb, float* C, void* ws, size_t ws_bytes, int R, int M,
                  int N, float alpha, const float* alpha_dev, float beta, rg_stream_t st_) {
  return gemm_tn_impl(A, lda, Bm, ldb, C, ws, ws_bytes, R, M, N, alpha, alpha_dev, beta,
                      static_cast<cudaStream_t>(st_));
}

}  // extern "C"
