// rg_gemm.cuh -- the tcgen05/TMEM/TMA implicit-GEMM tile engine shared by every dense contraction on the
// RNA-GAN hot path:
//
//   * gemm_fwd_kernel   : D[M=128 pixels, N<=256] += A[pixels, K] * B[N, K]^T, both operands K-major.
//                         The A tile of one k-block is a 4-D TMA box (64 channels x bw x bh x bb pixels) read at a
//                         per-tap pixel offset from one of up to four strided views ("parity maps") of an NHWC
//                         activation, so a stride-2 4x4 convolution (16 taps), its transposed form (4 output phases
//                         x 4 taps) and a plain GEMM (1 tap, H=W=1) are the same kernel with different tap tables.
//                         Out-of-image taps are zero-filled by TMA: no padding buffers, no im2col in HBM.
//   * gemm_wgrad_kernel : dW[tap][p, s] = sum_pixels lo[pixel, p] * hi[pixel@tap, s]; both operands are the raw NHWC
//                         tiles used as MN-major UMMA operands (reduction over pixel rows), split-K over pixel blocks
//                         with fp32 partials reduced in a fixed order (deterministic, like cudnn.deterministic=True in
//                         the reference, src/histopathology_gan.py:289).
//
// Both are persistent, warp-specialised kernels: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner),
// warps 2-5 = epilogue (TMEM -> registers -> HBM) overlapping the next tile's MMAs through a double-buffered
// accumulator (2 x 256 TMEM columns).
#pragma once
#include "rg_ptx.cuh"

namespace rg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                  // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kStages = 4;
constexpr int kAStageBytes = 128 * 128;      // 16 KiB
constexpr int kBStageBytes = 256 * 128;      // 32 KiB
constexpr int kStageBytes = kAStageBytes + kBStageBytes;
constexpr int kGemmSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 2048 /*epilogue affine*/;
constexpr int kGemmThreads = 192;
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;              // TMEM columns per accumulator stage

struct Tap {
  int8_t map;   // which A-view (parity map) this tap reads
  int8_t dh;    // row offset added to the tile's first row
  int8_t dw;    // column offset
  int8_t wtap;  // kernel position kh*4+kw this tap multiplies (used when B is read MN-major from w_down)
};

struct GemmMaps {
  CUtensorMap a[4];
  CUtensorMap b;
};

enum OutKind { OUT_BF16_NHWC = 0, OUT_F32_NHWC = 1, OUT_F32_NCHW = 2 };

struct FwdArgs {
  int nB, H, W;        // M index space: pixels of the low-resolution grid (plain GEMM: H=W=1, nB=M)
  int bw, bh, bb;      // TMA box (pixels) per M tile; bw*bh*bb <= 128 (== 128 for power-of-two grids)
  int rows_valid;      // bw*bh*bb: accumulator rows beyond it are never stored
  int tw, th, tb;      // boxes per dimension
  int m_tiles;
  int num_taps, chunks;   // k-blocks per tile = num_taps * chunks (chunks = C_A / 64)
  int num_phases;         // 1, or 4 for the transposed (upsampling) form
  int n_total, block_n, n_tiles;
  int b_phase_rows;       // row offset between phases in the packed weight matrix
  int a_2d;               // 1: maps.a[0] is a rank-2 [rows][K] view (plain GEMMs): cheaper for the TMA unit than 4-D
  int mc;                 // 1: launched as 2-CTA clusters; the two CTAs take neighbouring M tiles of the same N tile
                          //    and each fetches half of the B tile with TMA multicast (halves the L2->SM B traffic)
  int b_mn;               // 1: B operand is read MN-major from w_down[Cp][16*Cs] (K rows = p, N contiguous = s)
  int b_tap_cols;         // b_mn: column stride between kernel positions in w_down (= Cs)
  Tap taps[4][16];
  void* out;
  const float* col_scale;   // optional per-output-column scale (folded eval BatchNorm1d)
  const float* col_shift;   // optional per-output-column shift / bias
  float slope;              // LeakyReLU slope in [0,1] applied after scale/shift (1.0f = identity)
  int act_tanh;             // OUT_F32_NCHW only: tanh instead of LeakyReLU
  int OH, OW, OC;           // output tensor dims
  int sy, sx;               // output pixel = (i*sy + oy[phase], j*sx + ox[phase])
  int n_valid;              // columns >= n_valid are not stored
  int8_t oy[4], ox[4];
};

struct WgradArgs {
  int nB, H, W;
  int bw, bh, bb;      // pixel box per k-block; bw*bh*bb == 64
  int tw, th, tb;
  int num_pb, splits, pb_per_split;
  int m_tiles, n_tiles, slabs_per_tile, chunks_s;
  int Cp, Cs, num_taps;
  Tap taps[16];
  float* ws;           // [splits][num_taps][Cp][Cs] fp32 partials
  // splits == 1 with 16 taps: the epilogue writes the torch layout dW[p][s][kh][kw] directly (no partials, no reduce)
  float* direct_out;
  const float* alpha_dev;
  float alpha, beta;
  int msub;            // 128-row accumulators per unit: 1, or 2 (256-row M tile sharing every B slab)
};

// slab sl of N-tile nt -> (kernel position, 64-channel chunk).  For 4x4 kernels a tile holds the four kw of one kh
// and one chunk, so an epilogue thread owns 4 consecutive floats (16 B) of dW[p][s][kh][0..3].
__device__ __forceinline__ void wgrad_slab(const WgradArgs& p, int nt, int sl, int& tap, int& chunk) {
  if (p.num_taps == 16 && p.slabs_per_tile == 4) {
    chunk = nt >> 2;
    tap = (nt & 3) * 4 + sl;
  } else {
    const int qd = nt * p.slabs_per_tile + sl;
    tap = qd / p.chunks_s;
    chunk = qd - tap * p.chunks_s;
  }
}

// ------------------------------------------------------------------------------------------------ shared setup
struct PipeSmem {
  uint8_t* stages;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tfull;
  uint64_t* tempty;
  uint32_t* tmem_slot;
  float* s_scale;   // [256] per-column epilogue scale of the current tile
  float* s_shift;   // [256]
};

__device__ __forceinline__ PipeSmem carve_smem(uint8_t* raw) {
  PipeSmem s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  s.stages = base;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + kStages * kStageBytes);
  s.full = bars;
  s.empty = bars + kStages;
  s.tfull = bars + 2 * kStages;
  s.tempty = bars + 2 * kStages + 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  s.s_scale = reinterpret_cast<float*>(base + kStages * kStageBytes + 256);
  s.s_shift = s.s_scale + 256;
  return s;
}

__device__ __forceinline__ uint32_t pipeline_prologue(const PipeSmem& s, int warp, int cluster_size = 1) {
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < kStages; ++i) {
        mbar_init(&s.full[i], 1);
        mbar_init(&s.empty[i], cluster_size);   // a stage is free when EVERY CTA that multicasts into it released it
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s.tfull[i], 1);
        mbar_init(&s.tempty[i], 4);   // one arrival per epilogue warp
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(s.tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (cluster_size > 1) cluster_sync_all();     // peers' barriers exist before any remote arrive / multicast
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(s.tmem_slot);
}

// ------------------------------------------------------------------------------------------------ forward / dgrad
template <int OUT>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_fwd_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ FwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const PipeSmem s = carve_smem(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
  }
  const int csize = p.mc ? 2 : 1;
  const int crank = p.mc ? static_cast<int>(cluster_ctarank()) : 0;
  const uint32_t tmem_base = pipeline_prologue(s, warp, csize);

  const int num_kb = p.num_taps * p.chunks;
  // tile loop: one "slot" = csize neighbouring M tiles (one per CTA of the cluster) of the same (N tile, phase)
  const int mslots = (p.m_tiles + csize - 1) / csize;
  const int total_tiles = mslots * p.n_tiles * p.num_phases;
  const int tile0 = blockIdx.x / csize, tile_step = gridDim.x / csize;
  const uint32_t stage_tx = static_cast<uint32_t>(p.rows_valid) * 128u + static_cast<uint32_t>(p.block_n) * 128u;

  if (warp == 0) {
    // ===================================================== TMA producer (one elected lane)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int m_tile = (tile % mslots) * csize + crank;    // may be >= m_tiles for the odd tail: loads zero-fill
        const int rest = tile / mslots;
        const int n_tile = rest % p.n_tiles;
        const int ph = rest / p.n_tiles;
        const int jt = m_tile % p.tw;
        const int it = (m_tile / p.tw) % p.th;
        const int bt = m_tile / (p.tw * p.th);
        const int j0 = jt * p.bw, i0 = it * p.bh, b0 = bt * p.bb;
        const int brow = ph * p.b_phase_rows + n_tile * p.block_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.chunks;
          const int chunk = kb - tap * p.chunks;
          const Tap t = p.taps[ph][tap];
          mbar_wait(&s.empty[stage], phase ^ 1u);
          uint8_t* sa = s.stages + stage * kStageBytes;
          uint8_t* sb = sa + kAStageBytes;
          mbar_expect_tx(&s.full[stage], stage_tx);
          if (p.a_2d) tma_load_2d(&maps.a[0], &s.full[stage], sa, chunk * kBlockK, b0);
          else tma_load_4d(&maps.a[t.map], &s.full[stage], sa, chunk * kBlockK, j0 + t.dw, i0 + t.dh, b0);
          if (p.b_mn) {
            // B^T slabs [64 k-rows = p][64 n = s] straight out of w_down: no second packed copy of the weights
            const int col0 = t.wtap * p.b_tap_cols + n_tile * p.block_n;
            const int ns = p.block_n >> 6;
            if (p.mc) {       // this CTA fetches half of the slabs for both CTAs
              for (int sl = crank * (ns >> 1); sl < (crank + 1) * (ns >> 1); ++sl)
                tma_load_2d_mc(&maps.b, &s.full[stage], sb + sl * 8192, col0 + sl * 64, chunk * kBlockK, 3);
            } else {
              for (int sl = 0; sl < ns; ++sl)
                tma_load_2d(&maps.b, &s.full[stage], sb + sl * 8192, col0 + sl * 64, chunk * kBlockK);
            }
          } else if (p.mc) {  // half of the B rows (box = block_n/2 rows) for both CTAs
            const int half = p.block_n >> 1;
            tma_load_2d_mc(&maps.b, &s.full[stage], sb + crank * half * 128, kb * kBlockK, brow + crank * half, 3);
          } else {
            tma_load_2d(&maps.b, &s.full[stage], sb, kb * kBlockK, brow);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one elected lane)
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, p.block_n, 0, p.b_mn);
      const uint32_t b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t b_kstep = p.b_mn ? 128u : 2u;   // (>>4) address advance per UMMA_K: 16 rows x 128 B, or 32 B
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1u;
        mbar_wait(&s.tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(s.stages + stage * kStageBytes);
          const uint32_t sb = sa + kAStageBytes;
          const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(sb, b_lbo, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // +32 bytes per UMMA_K inside the 128-byte swizzle row => +2 in the (>>4) address field
            umma_bf16(tmem_d, da + 2u * k, db + b_kstep * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (p.mc) umma_commit_mc(&s.empty[stage], 3);   // both producers write into this stage of my smem
          else umma_commit(&s.empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&s.tfull[acc]);
      }
    }
  } else {
    // ===================================================== epilogue warps (TMEM lanes 32*(warp%4) ...)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int jj = row % p.bw;
    const int ii = (row / p.bw) % p.bh;
    const int bbi = row / (p.bw * p.bh);
    int iter = 0;
    int staged_n_tile = -1;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      const int m_tile = (tile % mslots) * csize + crank;
      const int rest = tile / mslots;
      const int n_tile = rest % p.n_tiles;
      const int ph = rest / p.n_tiles;
      const int jt = m_tile % p.tw;
      const int it = (m_tile / p.tw) % p.th;
      const int bt = m_tile / (p.tw * p.th);
      const int b = bt * p.bb + bbi, i = it * p.bh + ii, j = jt * p.bw + jj;
      const bool row_ok = (m_tile < p.m_tiles) && (row < p.rows_valid) && (b < p.nB) && (i < p.H) && (j < p.W);
      const int y = i * p.sy + p.oy[ph], x = j * p.sx + p.ox[ph];
      const int n0 = n_tile * p.block_n;

      // per-column scale / shift of this tile -> smem (one broadcast read per element instead of global loads)
      const bool affine = (p.col_scale != nullptr) || (p.col_shift != nullptr);
      if (affine && n_tile != staged_n_tile) {                 // reload only when the column range changes
        const int et = threadIdx.x - 64;                       // 0..127 among the epilogue warps
        asm volatile("bar.sync 1, 128;" ::: "memory");        // previous tile's readers are done
        for (int cc = et; cc < p.block_n; cc += 128) {
          const int col = n0 + cc;
          const bool ok = col < p.n_valid;
          s.s_scale[cc] = (ok && p.col_scale) ? __ldg(p.col_scale + col) : 1.0f;
          s.s_shift[cc] = (ok && p.col_shift) ? __ldg(p.col_shift + col) : 0.0f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        staged_n_tile = n_tile;
      }
      const bool lrelu = p.slope != 1.0f;

      mbar_wait(&s.tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(q * 32) << 16);

      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t v[32];
        if (p.block_n >= 32) {
          tmem_ld_32x32(taddr + c, v);
        } else {
          uint32_t w[16];
          tmem_ld_32x16(taddr + c, w);
#pragma unroll
          for (int e = 0; e < 16; ++e) { v[e] = w[e]; v[16 + e] = 0u; }
        }
        tmem_ld_wait();
        const int ncols = min(32, p.block_n - c);
        if (affine) {
          const float4* sc4 = reinterpret_cast<const float4*>(s.s_scale + (c & 255));
          const float4* sh4 = reinterpret_cast<const float4*>(s.s_shift + (c & 255));
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4) {
            const float4 a4 = sc4[g4], b4 = sh4[g4];
            v[g4 * 4 + 0] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 0]), a4.x, b4.x));
            v[g4 * 4 + 1] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 1]), a4.y, b4.y));
            v[g4 * 4 + 2] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 2]), a4.z, b4.z));
            v[g4 * 4 + 3] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 3]), a4.w, b4.w));
          }
        }
        if (OUT == OUT_F32_NCHW && p.act_tanh) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(tanhf(__uint_as_float(v[e])));
        } else if (lrelu) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float f = __uint_as_float(v[e]);
            v[e] = __float_as_uint(fmaxf(f, f * p.slope));
          }
        }
        if (row_ok) {
          if (OUT == OUT_BF16_NHWC) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                               (static_cast<size_t>(b * p.OH + y) * p.OW + x) * p.OC + n0 + c;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int col = n0 + c + g * 8;
              if (g * 8 < ncols && col + 8 <= p.n_valid) {
                uint4 u;
                u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
                u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
                u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
                u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
                *reinterpret_cast<uint4*>(o + g * 8) = u;
              } else if (g * 8 < ncols) {
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (col + e < p.n_valid) o[g * 8 + e] = __float2bfloat16(__uint_as_float(v[g * 8 + e]));
              }
            }
          } else if (OUT == OUT_F32_NHWC) {
            float* o = reinterpret_cast<float*>(p.out) + (static_cast<size_t>(b * p.OH + y) * p.OW + x) * p.OC + n0 + c;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const int col = n0 + c + g * 4;
              if (g * 4 < ncols && col + 4 <= p.n_valid) {
                float4 u = make_float4(__uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1]),
                                       __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
                *reinterpret_cast<float4*>(o + g * 4) = u;
              } else if (g * 4 < ncols) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (col + e < p.n_valid) o[g * 4 + e] = __uint_as_float(v[g * 4 + e]);
              }
            }
          } else {   // OUT_F32_NCHW: a handful of image channels, planar fp32
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = n0 + c + e;
              if (e < ncols && col < p.n_valid)
                o[(static_cast<size_t>(b * p.OC + col) * p.OH + y) * p.OW + x] = __uint_as_float(v[e]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.mc) cluster_sync_all();   // the peer may still multicast into this CTA's smem / arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// Measured on B200 (profiles/r1_wgrad_whatif.txt): with both operands streamed once from HBM this kernel is bound by
// bytes in flight per SM (smem ring / DRAM latency), not by the tensor pipe -- removing every MMA changes its time by
// < 10 %.  msub = 2 therefore processes a 256-row M tile per unit (two 128-row accumulators sharing every B slab):
// 1.5x the FLOPs per byte of smem ring at the price of the accumulator double-buffering (all 512 TMEM columns).
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ WgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const PipeSmem s = carve_smem(smem_raw);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
  }
  const uint32_t tmem_base = pipeline_prologue(s, warp);

  const int msub = p.msub;                                    // 128-row accumulators per unit (1 or 2)
  const int a_bytes = msub * 16384;                           // msub * 2 slabs of [64 px][64 ch]
  const int stage_bytes = a_bytes + 32768;
  const int nstages = msub == 2 ? 3 : 4;                      // 3 x 64 KiB or 4 x 48 KiB
  const int total_units = p.m_tiles * p.n_tiles * p.splits;
  const uint32_t stage_tx = static_cast<uint32_t>(2 * msub + p.slabs_per_tile) * 8192u;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int m_tile = u % p.m_tiles;
        const int rest = u / p.m_tiles;
        const int n_tile = rest % p.n_tiles;
        const int split = rest / p.n_tiles;
        const int pb0 = split * p.pb_per_split;
        const int pb1 = min(p.num_pb, pb0 + p.pb_per_split);
        for (int pb = pb0; pb < pb1; ++pb) {
          const int jt = pb % p.tw;
          const int it = (pb / p.tw) % p.th;
          const int bt = pb / (p.tw * p.th);
          const int j0 = jt * p.bw, i0 = it * p.bh, b0 = bt * p.bb;
          mbar_wait(&s.empty[stage], phase ^ 1u);
          uint8_t* sa = s.stages + stage * stage_bytes;
          uint8_t* sb = sa + a_bytes;
          mbar_expect_tx(&s.full[stage], stage_tx);
          // MMA-A operand: 64-channel slabs of the low-resolution tensor (channels beyond Cp zero-fill)
          for (int sl = 0; sl < 2 * msub; ++sl)
            tma_load_4d(&maps.b, &s.full[stage], sa + sl * 8192, m_tile * 128 * msub + sl * 64, j0, i0, b0);
          // MMA-B operand: one slab per (tap, 64-channel chunk) of the high-resolution tensor
          for (int sl = 0; sl < p.slabs_per_tile; ++sl) {
            int tap, chunk;
            wgrad_slab(p, n_tile, sl, tap, chunk);
            const Tap t = p.taps[tap];
            tma_load_4d(&maps.a[t.map], &s.full[stage], sb + sl * 8192, chunk * 64, j0 + t.dw, i0 + t.dh, b0);
          }
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, 64 * p.slabs_per_tile, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++iter) {
        const int split = u / (p.m_tiles * p.n_tiles);
        const int pb0 = split * p.pb_per_split;
        const int pb1 = min(p.num_pb, pb0 + p.pb_per_split);
        // msub == 2 owns all 512 TMEM columns: a single accumulator stage
        const int acc = msub == 2 ? 0 : (iter & 1);
        const uint32_t acc_phase = msub == 2 ? (iter & 1u) : ((iter >> 1) & 1u);
        mbar_wait(&s.tempty[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        for (int pb = pb0; pb < pb1; ++pb) {
          mbar_wait(&s.full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(s.stages + stage * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          // MN-major SWIZZLE_128B: LBO = stride between 64-element MN slabs (8 KiB),
          // SBO = stride between 8-row K groups (1 KiB); one UMMA_K = 16 pixel rows = 2 KiB.
          const uint64_t db = make_smem_desc_sw128(sb, 8192, 1024);
          for (int ms = 0; ms < msub; ++ms) {
            const uint64_t da = make_smem_desc_sw128(sa + ms * 16384, 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_d + ms * kAccStride, da + 128u * k, db + 128u * k, idesc, (pb > pb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&s.empty[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&s.tfull[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int iter = 0;
    for (int u = blockIdx.x; u < total_units; u += gridDim.x, ++iter) {
      const int acc = msub == 2 ? 0 : (iter & 1);
      const uint32_t acc_phase = msub == 2 ? (iter & 1u) : ((iter >> 1) & 1u);
      const int m_tile = u % p.m_tiles;
      const int rest = u / p.m_tiles;
      const int n_tile = rest % p.n_tiles;
      const int split = rest / p.n_tiles;
      mbar_wait(&s.tfull[acc], acc_phase);
      tc_fence_after();
      for (int ms = 0; ms < msub; ++ms) {
        const int prow = (m_tile * msub + ms) * 128 + row;
        const bool row_ok = prow < p.Cp;
        const uint32_t taddr = tmem_base + (acc + ms) * kAccStride + (static_cast<uint32_t>(q * 32) << 16);
        if (p.direct_out != nullptr) {
          // direct write: dW[p][s][kh][kw] = beta*dW + alpha * acc; this tile = (kh, chunk), slabs = kw 0..3
          const int chunk = n_tile >> 2, kh = n_tile & 3;
          const float a = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f);
#pragma unroll 1
          for (int h = 0; h < 4; ++h) {
            uint32_t v0[16], v1[16], v2[16], v3[16];
            tmem_ld_32x16(taddr + 0 * 64 + h * 16, v0);
            tmem_ld_32x16(taddr + 1 * 64 + h * 16, v1);
            tmem_ld_32x16(taddr + 2 * 64 + h * 16, v2);
            tmem_ld_32x16(taddr + 3 * 64 + h * 16, v3);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const int sidx = chunk * 64 + h * 16 + e;
                if (sidx < p.Cs) {
                  float* o = p.direct_out + (static_cast<size_t>(prow) * p.Cs + sidx) * 16 + kh * 4;
                  float4 f = make_float4(a * __uint_as_float(v0[e]), a * __uint_as_float(v1[e]),
                                         a * __uint_as_float(v2[e]), a * __uint_as_float(v3[e]));
                  if (p.beta != 0.0f) {
                    const float4 old = *reinterpret_cast<const float4*>(o);
                    f.x += p.beta * old.x; f.y += p.beta * old.y; f.z += p.beta * old.z; f.w += p.beta * old.w;
                  }
                  *reinterpret_cast<float4*>(o) = f;
                }
              }
            }
          }
        } else {
          for (int sl = 0; sl < p.slabs_per_tile; ++sl) {
            int tap, chunk;
            wgrad_slab(p, n_tile, sl, tap, chunk);
            float* o = p.ws + ((static_cast<size_t>(split) * p.num_taps + tap) * p.Cp + prow) * p.Cs + chunk * 64;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32];
              tmem_ld_32x32(taddr + sl * 64 + h * 32, v);
              tmem_ld_wait();
              if (row_ok) {
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                  if (chunk * 64 + h * 32 + g * 4 + 4 <= p.Cs) {
                    float4 f = make_float4(__uint_as_float(v[g * 4 + 0]), __uint_as_float(v[g * 4 + 1]),
                                           __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3]));
                    *reinterpret_cast<float4*>(o + h * 32 + g * 4) = f;
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace rg
