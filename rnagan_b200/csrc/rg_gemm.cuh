// rg_gemm.cuh -- the tcgen05/TMEM/TMA implicit-GEMM tile engine shared by every dense contraction on the
// RNA-GAN hot path:
//
//   * gemm_fwd_kernel   : D[M pixels, N<=256] += A[pixels, K] * B[N, K]^T.  The A tile of one k-block is a 4-D TMA box
//                         (64 channels x bw x bh x bb pixels) read at a per-tap pixel offset from one of up to four
//                         strided views ("parity maps") of an NHWC activation, so a stride-2 4x4 convolution (16
//                         taps), its transposed form (4 output phases x 4 taps, or all four phases merged in one tile
//                         for Cs = 64), a 3x3 convolution and a plain GEMM (1 tap, H=W=1) are the same kernel with
//                         different tap tables.  Out-of-image taps are zero-filled by TMA: no padding buffers, no
//                         im2col in HBM.  B is K-major or, straight from w_down, MN-major.
//                         CG = 2: the two SMs of a TPC run one tcgen05.mma.cta_group::2 (M = 256) per k-step, each CTA
//                         fetching its own 128 A rows and half of the B rows.
//   * gemm_wgrad_kernel : dW[tap][p, s] = sum_pixels lo[pixel, p] * hi[pixel@tap, s]; both operands are the raw NHWC
//                         tiles used as MN-major UMMA operands (reduction over pixel rows); output in the native
//                         [p][tap][s] layout directly (one unit covers all pixels) or split-K over pixel blocks with
//                         fp32 partials reduced in a fixed order (deterministic, like cudnn.deterministic=True in the
//                         reference, src/histopathology_gan.py:289).
//
// Both are persistent, warp-specialised kernels of 224 threads: warp 0 = TMA producer of the A operand (+ expect_tx),
// warp 6 = TMA producer of the B operand, warp 1 = single-thread MMA issuer (+ TMEM owner), warps 2-5 = epilogue
// (TMEM -> registers -> swizzled smem slab -> TMA store, with optional per-column affine / activation, fused BatchNorm
// statistics and the opt-in fused elementwise backward) overlapping the next tile's MMAs through a double-buffered
// accumulator (2 x 256 TMEM columns).
#pragma once
#include "rg_ptx.cuh"

namespace rg {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                  // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kMaxStages = 8;
constexpr int kAStageBytes = 128 * 128;      // 16 KiB: 128 pixel rows x 64 channels
constexpr int kStagingBytes = 128 * 128;     // one 64-column bf16 slab of the output tile (epilogue -> TMA store)
constexpr int kBarBytes = 256;
constexpr int kAffineBytes = 4096;           // per-column vectors of the current tile: scale, shift (+ mean, rstd for the
                                             // fused BatchNorm backward), 4 x 256 floats
constexpr int kStatBytes = 4096;             // per-warp column sums of one slab (4 warps x 2 x 64 floats), two slab parities
constexpr int kSmemFixedBytes = kBarBytes + kAffineBytes + kStatBytes + 1024 /*alignment slack*/;
constexpr int kSmemMaxBytes = 232448;        // 227 KiB: the sm_100 per-CTA dynamic shared memory limit
constexpr int kGemmThreads = 224;            // warps: 0 A-producer, 1 MMA issuer, 2-5 epilogue, 6 B-producer
constexpr int kTmemCols = 512;
constexpr int kAccStride = 256;              // TMEM columns per accumulator stage
// weight-gradient kernel: fixed ring (4 x 48 KiB, or 3 x 64 KiB for 256-row units)
constexpr int kWgradStageT = 4 * 4096;         // epilogue transpose slabs: 4 warps x (32 rows x 32 fp32)
constexpr int kWgradSmemBytes = 4 * (kAStageBytes + 256 * 128) + kWgradStageT + kSmemFixedBytes;

struct Tap {
  int8_t map;   // which A-view (parity map) this tap reads
  int8_t dh;    // row offset added to the tile's first row
  int8_t dw;    // column offset
  int8_t wtap;  // kernel position kh*4+kw this tap multiplies (used when B is read MN-major from w_down)
};

struct GemmMaps {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap o[4];   // per-phase output views (bf16 NHWC through the TMA-store epilogue)
  CUtensorMap x[4];   // the same views of the aux tensor of the fused elementwise backward (TMA-staged into smem)
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
  int b_mn;               // 1: B operand is read MN-major from w_down[Cp][16*Cs] (K rows = p, N contiguous = s)
  int b_tap_cols;         // b_mn: column stride between kernel positions in w_down (= Cs)
  int nstages;            // smem ring depth (host-planned from block_n and the CTA-pair mode)
  int b_stage_bytes;      // bytes of the B operand one CTA holds per stage ((block_n / CG) * 128)
  int tma_store;          // OUT_BF16_NHWC: tile leaves through swizzled smem slabs + TMA stores (whole 128-byte lines)
  int nbuf;               // staging slabs (1 or 2)
  int merged;             // transposed form with ALL FOUR output phases in one tile (Cs == 64, CTA pairs): the 9 distinct
                          // shifted A tiles of an M tile are fetched once instead of 16 times; TMEM slab q = phase q
  // merged: tap i of the list (centre shift first) feeds mg_nph[i] consecutive 64-column phase slabs starting at phase
  // mg_col[i] with ONE MMA of N = 64 * mg_nph[i] per k-step (a phase the shift does not feed gets a zero B slab): an
  // N = 64 tcgen05.mma costs ~94 cycles against 128 for N = 256 (measured), so wide instructions are what counts
  int8_t mg_nph[9];
  int8_t mg_col[9];
  // Fused elementwise backward of the layer this GEMM's output is the input gradient of (TMA-store epilogue only):
  // aux has the layout of `out`.  mode 1 (no BatchNorm): out = acc * lrelu'(aux), aux = the stored activation h.
  // mode 2 (BatchNorm + LeakyReLU): aux = the pre-BN activation a; out = du = acc * lrelu'(scale*a + shift) and, with
  // `stats`, the per-CTA partial sums are S(du) and S(du * xhat), xhat = (a - mean) * rstd -- rg_bn_bwd_reduce without
  // a pass over dh and a.
  const __nv_bfloat16* aux;
  int aux_mode;
  int naux;                      // aux staging slabs in smem (2 when aux_mode != 0)
  const float* aux_mean;
  const float* aux_rstd;
  const float* aux_scale;
  const float* aux_shift;
  float aux_slope;
  int whatif;             // profiling experiments (RG_WHATIF): 1 no MMA, 2 no epilogue work, 4 no A loads, 8 no B loads
  long long* prof;        // optional [gridDim.x][12] clock64 totals per role (tools/gemm_prof.py); null in production
  float* stats;           // optional [gridDim.x][2][n_total]: per-CTA column sums / sums of squares of the STORED
                          // (bf16-rounded) outputs, accumulated in a fixed order -> BatchNorm statistics without
                          // re-reading the activation (finalised by rg_bn_finalize_partials)
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
  // splits == 1: the epilogue writes the NATIVE layout dW[p][tap][s] (= a channels_last [Cp][Cs][kh][kw] tensor): each
  // thread stores whole 128-byte lines, beta/alpha applied in place, no partials and no reduce pass
  float* native_out;
  int ld_out, n_out;   // native_out row pitch and valid columns per (row, tap): Cs, or the ragged N of a plain GEMM
                       // (then rows are only 8-byte aligned and the epilogue stores float2)
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
  uint8_t* staging;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tfull;
  uint64_t* tempty;
  uint32_t* tmem_slot;
  uint64_t* auxbar;   // [2] "aux slab landed" barriers (fused elementwise backward)
  uint8_t* auxstg;    // [2][16 KiB] aux slabs, same swizzled layout as the output staging slabs
  float* s_scale;   // [256] per-column epilogue scale of the current tile
  float* s_shift;   // [256]
  float* s_mean;    // [256] fused BatchNorm backward only
  float* s_rstd;    // [256]
  float* s_stat;    // [2 slab parities][4 warps][2][64]
};

// layout: [ring: ring_bytes][staging: staging_bytes][barriers 256][affine 2048][stat 2048], ring 1024-byte aligned
__device__ __forceinline__ PipeSmem carve_smem(uint8_t* raw, int ring_bytes, int staging_bytes, int aux_bytes = 0) {
  PipeSmem s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  s.stages = base;
  s.staging = base + ring_bytes;
  s.auxstg = base + ring_bytes + staging_bytes;
  staging_bytes += aux_bytes;          // the fixed block follows the aux slabs
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + ring_bytes + staging_bytes);
  s.full = bars;
  s.empty = bars + kMaxStages;
  s.tfull = bars + 2 * kMaxStages;
  s.tempty = bars + 2 * kMaxStages + 2;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  s.auxbar = bars + 2 * kMaxStages + 6;
  s.s_scale = reinterpret_cast<float*>(base + ring_bytes + staging_bytes + kBarBytes);
  s.s_shift = s.s_scale + 256;
  s.s_mean = s.s_scale + 512;
  s.s_rstd = s.s_scale + 768;
  s.s_stat = s.s_scale + 1024;
  uint32_t dyn;
  asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn));
  if (reinterpret_cast<uint8_t*>(s.s_stat) + kStatBytes > raw + dyn) __trap();   // host under-provisioned the launch
  return s;
}

// CG = CTAs per MMA group (1, or 2 = tcgen05 cta_group::2 pair launched as a 2-CTA cluster)
template <int CG>
__device__ __forceinline__ uint32_t pipeline_prologue(const PipeSmem& s, int warp, int nstages) {
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < nstages; ++i) {
        mbar_init(&s.full[i], 1);
        mbar_init(&s.empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&s.tfull[i], 1);
        mbar_init(&s.tempty[i], 4 * CG);   // one arrival per epilogue warp of every CTA in the group
        mbar_init(&s.auxbar[i], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_cg<CG>(s.tmem_slot, kTmemCols);
    tmem_relinquish_cg<CG>();
  }
  tc_fence_before();
  __syncthreads();
  if (CG > 1) cluster_sync_all();     // the peer's barriers exist before any remote arrive / paired TMA load
  tc_fence_after();
  return *reinterpret_cast<volatile uint32_t*>(s.tmem_slot);
}

// Column sums over the 32 lanes of a warp of a 32-register row fragment: lane l returns sum_lanes x[l].
// Butterfly transpose-reduce: 31 shuffles for 32 columns, fixed summation order (deterministic).
__device__ __forceinline__ float warp_colsum32(const float (&x)[32], int lane) {
  float a16[16], a8[8], a4[4], a2[2];
  const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0, h2 = (lane & 2) != 0,
             h1 = (lane & 1) != 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float keep = h16 ? x[16 + i] : x[i], send = h16 ? x[i] : x[16 + i];
    a16[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float keep = h8 ? a16[8 + i] : a16[i], send = h8 ? a16[i] : a16[8 + i];
    a8[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float keep = h4 ? a8[4 + i] : a8[i], send = h4 ? a8[i] : a8[4 + i];
    a4[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float keep = h2 ? a4[2 + i] : a4[i], send = h2 ? a4[i] : a4[2 + i];
    a2[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, 2);
  }
  const float keep = h1 ? a2[1] : a2[0], send = h1 ? a2[0] : a2[1];
  return keep + __shfl_xor_sync(0xFFFFFFFFu, send, 1);
}

__device__ __forceinline__ float bf16_round(float f) { return __bfloat162float(__float2bfloat16(f)); }

// ------------------------------------------------------------------------------------------------ forward / dgrad
// Tile order: phase fastest, then M slot, then N tile -- CTAs running at the same time share the A tiles of an M slot
// across the four output phases of the transposed form (L2 hits) and a persistent CTA keeps its N tile for long runs
// (statistics accumulate in registers and reach memory only when the N tile changes).
// AUX: compile the fused elementwise-backward epilogue (rg_epilogue_aux) in; the plain instantiation carries none of it
template <int OUT, int CG, bool AUX = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_fwd_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ FwdArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const int stage_bytes = kAStageBytes + p.b_stage_bytes;
  const PipeSmem s = carve_smem(smem_raw, p.nstages * stage_bytes, p.nbuf * kStagingBytes,
                                p.naux * kStagingBytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    if (OUT == OUT_BF16_NHWC && p.tma_store) tma_prefetch_desc(&maps.o[0]);
  }
  const int crank = CG > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const uint32_t tmem_base = pipeline_prologue<CG>(s, warp, p.nstages);
  griddep_wait();        // PDL: everything above overlapped the previous kernel's tail; no global access before this
  griddep_launch();

  const int num_kb = p.num_taps * p.chunks;
  const int nstages = p.nstages;
  const int mslots = (p.m_tiles + CG - 1) / CG;     // one slot = the M tiles of one MMA group
  const int total_tiles = mslots * p.n_tiles * p.num_phases;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;
  // bytes landing per stage on the (leader's) full barrier: A box + this CTA's share of B, from every CTA of the group
  const uint32_t stage_tx = (((p.whatif & 4) ? 0u : static_cast<uint32_t>(p.rows_valid) * 128u) +
                             ((p.whatif & 8) ? 0u : static_cast<uint32_t>(p.b_stage_bytes))) * CG;
  const int bn_cta = p.block_n / CG;                // B rows (N columns) this CTA fetches
  const int mg_rows = 64 / CG;                      // merged: rows of one phase's B slab held by this CTA

  if (warp == 0 || warp == 6) {
    // ===================================================== TMA producers (one elected lane each, in every CTA of the
    // group): warp 0 announces the stage's bytes and fetches the A tile, warp 6 fetches the B tile.  Issuing a TMA
    // costs its thread ~100 cycles, so two issuers halve the per-k-block cost that bounds the narrow layers.
    // This single thread paces the whole pipeline: everything loop-invariant lives in registers (the asm memory
    // clobbers would otherwise make the compiler re-read kernel parameters every k-block), there is no integer
    // division in the k loop (taps outer, channel chunks inner) and shared addresses are plain 32-bit integers.
    if (elect_one()) {
      const int chunks = p.chunks, num_taps = p.num_taps, num_phases = p.num_phases;
      const int tw = p.tw, th = p.th, bw = p.bw, bh = p.bh, bb = p.bb;
      const int block_n = p.block_n, b_phase_rows = p.b_phase_rows, b_tap_cols = p.b_tap_cols;
      const bool a_2d = p.a_2d != 0, b_mn = p.b_mn != 0, merged = p.merged != 0;
      const uint32_t a_tx = (p.whatif & 4) ? 0u : static_cast<uint32_t>(p.rows_valid) * 128u;
      const bool do_a = warp == 0 && (p.whatif & 4) == 0, do_b = warp == 6 && (p.whatif & 8) == 0;
      const bool prof = p.prof != nullptr && warp == 0;
      const int ns = bn_cta >> 6;
      const uint32_t ring0 = smem_u32(s.stages);
      const uint32_t full0 = smem_u32(&s.full[0]), empty0 = smem_u32(&s.empty[0]);
      const uint32_t full0_tx = CG > 1 ? (full0 & kPeerBitMask) : full0;     // the barrier the TMA bytes count on
      const uint64_t desc_b = reinterpret_cast<uint64_t>(&maps.b);
      const uint64_t desc_a0 = reinterpret_cast<uint64_t>(&maps.a[0]);
      int stage = 0;
      uint32_t phase = 0;
      long long t_wait = 0, t_issue = 0;
      const long long t_begin = clock64();
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int ph = tile % num_phases;
        const int rest = tile / num_phases;
        const int m_tile = (rest % mslots) * CG + crank;      // may be >= m_tiles for an odd tail: loads zero-fill
        const int n_tile = rest / mslots;
        const int jt = m_tile % tw;
        const int it = (m_tile / tw) % th;
        const int bt = m_tile / (tw * th);
        const int j0 = jt * bw, i0 = it * bh, b0 = bt * bb;
        const int ncol0 = n_tile * block_n + crank * bn_cta;
        const int brow = ph * b_phase_rows + ncol0;
        int kcol = 0;                                          // kb * 64: K coordinate of a K-major B
        for (int tap = 0; tap < num_taps; ++tap) {
          const Tap t = p.taps[ph][tap];
          const int mg_n = merged ? p.mg_nph[tap] : 0;
          const uint64_t desc_a = desc_a0 + static_cast<uint64_t>(t.map) * sizeof(CUtensorMap);
          const int cj = j0 + t.dw, ci = i0 + t.dh;
          const int bcol = t.wtap * b_tap_cols + ncol0;
          for (int chunk = 0; chunk < chunks; ++chunk, kcol += kBlockK) {
            const long long c0 = prof ? clock64() : 0;
            mbar_wait_raw(empty0 + stage * 8, phase ^ 1u);
            const long long c1 = prof ? clock64() : 0;
            const uint32_t sa = ring0 + stage * stage_bytes;
            const uint32_t sb = sa + kAStageBytes;
            const uint32_t fb = full0_tx + stage * 8;
            if (warp == 0 && crank == 0)
              mbar_expect_tx_raw(full0 + stage * 8,
                                 merged ? (a_tx + static_cast<uint32_t>(mg_n * mg_rows) * 128u) * CG : stage_tx);
            if (do_a) {
              if (a_2d) tma_ld_2d_raw<CG>(desc_a0, fb, sa, chunk * kBlockK, b0);
              else tma_ld_4d_raw<CG>(desc_a, fb, sa, chunk * kBlockK, cj, ci, b0);
            }
            if (do_b && merged) {
              // one box = the slabs of every phase this tap feeds (this CTA's half of their rows), contiguous in w_up9
              const uint64_t d9 = mg_n == 4 ? desc_b : desc_a0 + static_cast<uint64_t>(mg_n == 2 ? 1 : (mg_n == 1 ? 2 : 3)) * sizeof(CUtensorMap);
              tma_ld_2d_raw<CG>(d9, fb, sb, 0, ((tap * chunks + chunk) * CG + crank) * (4 * mg_rows));
            } else if (do_b) {
              if (b_mn) {
                // B^T slabs [64 k-rows = p][64 n = s] straight out of w_down: no second packed copy of the weights
                for (int sl = 0; sl < ns; ++sl)
                  tma_ld_2d_raw<CG>(desc_b, fb, sb + sl * 8192, bcol + sl * 64, chunk * kBlockK);
              } else {
                tma_ld_2d_raw<CG>(desc_b, fb, sb, kcol, brow);
              }
            }
            if (prof) { t_wait += c1 - c0; t_issue += clock64() - c1; }
            if (++stage == nstages) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (prof) {
        long long* o = p.prof + static_cast<size_t>(blockIdx.x) * 12;
        o[0] = clock64() - t_begin; o[1] = t_wait; o[2] = t_issue;
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one elected lane of the leader CTA)
    if (crank == 0 && elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kBlockM * CG, p.block_n, 0, p.b_mn);
      const uint32_t b_lbo = p.b_mn ? 8192u : 16u;
      const uint32_t b_kstep = p.b_mn ? 128u : 2u;   // (>>4) address advance per UMMA_K: 16 rows x 128 B, or 32 B
      const bool skip_mma = (p.whatif & 1) != 0, prof = p.prof != nullptr;
      const uint32_t ring0 = smem_u32(s.stages);
      const uint32_t full0 = smem_u32(&s.full[0]);
      // descriptor templates: only the 14-bit start-address field changes per stage / k step
      const uint64_t da0 = make_smem_desc_sw128(0, 16, 1024);
      const uint64_t db0 = make_smem_desc_sw128(0, b_lbo, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      long long t_wfull = 0, t_wtempty = 0;
      const long long t_begin = clock64();
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1u;
        const long long c0 = prof ? clock64() : 0;
        mbar_wait(&s.tempty[acc], acc_phase ^ 1u);
        if (prof) t_wtempty += clock64() - c0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kAccStride;
        if (p.merged) {
          // 9 shifts x chunks stages; one MMA per k-step covers every phase slab the shift feeds (N = 64..256).
          // The centre shift comes first and feeds all four phases: its first MMA initialises the whole accumulator.
          for (int tap = 0; tap < p.num_taps; ++tap) {
            const int nsl = p.mg_nph[tap];
            const uint32_t idesc_n = make_idesc_bf16(kBlockM * CG, 64 * nsl, 0, 0);
            const uint32_t tmem_t = tmem_d + p.mg_col[tap] * 64;
            for (int chunk = 0; chunk < p.chunks; ++chunk) {
              const long long c1 = prof ? clock64() : 0;
              mbar_wait_raw(full0 + stage * 8, phase);
              if (prof) t_wfull += clock64() - c1;
              tc_fence_after();
              const uint32_t sa = ring0 + stage * stage_bytes;
              const uint64_t da = da0 | static_cast<uint64_t>((sa >> 4) & 0x3FFF);
              const uint32_t sbk = sa + kAStageBytes;
              const uint64_t db = db0 | static_cast<uint64_t>((sbk >> 4) & 0x3FFF);
              if (!skip_mma) {
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k)
                  umma_bf16_cg<CG>(tmem_t, da + 2u * k, db + 2u * k, idesc_n, (tap | chunk | k) != 0 ? 1u : 0u);
              }
              umma_commit_cg<CG>(&s.empty[stage]);
              if (++stage == nstages) { stage = 0; phase ^= 1u; }
            }
          }
          umma_commit_cg<CG>(&s.tfull[acc]);
          continue;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          const long long c1 = prof ? clock64() : 0;
          mbar_wait_raw(full0 + stage * 8, phase);
          if (prof) t_wfull += clock64() - c1;
          tc_fence_after();
          const uint32_t sa = ring0 + stage * stage_bytes;
          const uint64_t da = da0 | static_cast<uint64_t>((sa >> 4) & 0x3FFF);
          const uint32_t sbk = sa + kAStageBytes;
          const uint64_t db = db0 | static_cast<uint64_t>((sbk >> 4) & 0x3FFF);
          if (!skip_mma) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              // +32 bytes per UMMA_K inside the 128-byte swizzle row => +2 in the (>>4) address field
              umma_bf16_cg<CG>(tmem_d, da + 2u * k, db + b_kstep * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit_cg<CG>(&s.empty[stage]);        // frees this stage in every CTA of the group
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_cg<CG>(&s.tfull[acc]);
      }
      if (prof) {
        long long* o = p.prof + static_cast<size_t>(blockIdx.x) * 12;
        o[3] = clock64() - t_begin; o[4] = t_wfull; o[5] = t_wtempty;
      }
    }
  } else {
    // ===================================================== epilogue warps (TMEM lanes 32*(warp%4) ...)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;                         // 0..127 among the epilogue warps
    const int jj = row % p.bw;
    const int ii = (row / p.bw) % p.bh;
    const int bbi = row / (p.bw * p.bh);
    const bool affine = (p.col_scale != nullptr) || (p.col_shift != nullptr);
    const bool lrelu = p.slope != 1.0f;
    const bool use_tma = (OUT == OUT_BF16_NHWC) && p.tma_store;
    const bool do_stats = use_tma && (p.stats != nullptr);
    const bool aux_on = AUX && use_tma && p.aux != nullptr && p.aux_mode != 0;
    const bool aux_bn = aux_on && p.aux_mode == 2;
    // the tempty barrier lives in the leader CTA
    const uint32_t tempty_addr0 = CG > 1 ? mapa_rank(smem_u32(&s.tempty[0]), 0) : smem_u32(&s.tempty[0]);
    // statistics: thread et owns (quantity k = et >> 6, column (et & 63) of every slab) of this CTA's partial row
    const int st_k = et >> 6, st_c = et & 63;
    float st_acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float* st_row = do_stats ? p.stats + (static_cast<size_t>(blockIdx.x) * 2 + st_k) * p.n_total : nullptr;
    if (do_stats)
      for (int c = st_c; c < p.n_total; c += 64) st_row[c] = 0.0f;
    int st_ntile = -1;
    int store_count = 0;
    int iter = 0;
    int staged_n_tile = -1;
    long long t_wtfull = 0, t_wstore = 0;
    const long long t_begin = clock64();
    // fused elementwise backward: the leader streams the aux slabs two slabs ahead of their use, in the (tile, slab)
    // order the epilogue consumes them
    int pf_tile = tile0, pf_sl = 0, pf_count = 0;
    auto aux_prefetch = [&]() {
      if (pf_tile >= total_tiles) return;
      const int ph_ = pf_tile % p.num_phases;
      const int rest_ = pf_tile / p.num_phases;
      const int m_ = (rest_ % mslots) * CG + crank;
      const int nt_ = rest_ / mslots;
      const int j0_ = (m_ % p.tw) * p.bw, i0_ = ((m_ / p.tw) % p.th) * p.bh, b0_ = (m_ / (p.tw * p.th)) * p.bb;
      uint64_t* bar = &s.auxbar[pf_count & 1];
      mbar_expect_tx(bar, static_cast<uint32_t>(p.rows_valid) * 128u);
      tma_load_4d(p.merged ? &maps.x[pf_sl] : &maps.x[ph_], bar, s.auxstg + (pf_count & 1) * kStagingBytes,
                  p.merged ? 0 : nt_ * p.block_n + pf_sl * 64, j0_, i0_, b0_);
      ++pf_count;
      if (++pf_sl == (p.block_n >> 6)) { pf_sl = 0; pf_tile += tile_step; }
    };
    if (aux_on && et == 0) { aux_prefetch(); aux_prefetch(); }
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1u;
      const int ph = tile % p.num_phases;
      const int rest = tile / p.num_phases;
      const int m_tile = (rest % mslots) * CG + crank;
      const int n_tile = rest / mslots;
      const int jt = m_tile % p.tw;
      const int it = (m_tile / p.tw) % p.th;
      const int bt = m_tile / (p.tw * p.th);
      const int b = bt * p.bb + bbi, i = it * p.bh + ii, j = jt * p.bw + jj;
      const bool tile_ok = m_tile < p.m_tiles;
      const bool row_ok = tile_ok && (row < p.rows_valid) && (b < p.nB) && (i < p.H) && (j < p.W);
      const int y = i * p.sy + p.oy[ph], x = j * p.sx + p.ox[ph];
      const int n0 = n_tile * p.block_n;

      // per-column scale / shift of this tile -> smem (one broadcast read per element instead of global loads)
      if ((affine || aux_bn) && n_tile != staged_n_tile) {     // reload only when the column range changes
        asm volatile("bar.sync 1, 128;" ::: "memory");        // previous tile's readers are done
        for (int cc = et; cc < p.block_n; cc += 128) {
          const int col = n0 + cc;
          const bool ok = col < p.n_valid;
          if (aux_bn) {
            s.s_scale[cc] = ok ? __ldg(p.aux_scale + col) : 1.0f;
            s.s_shift[cc] = ok ? __ldg(p.aux_shift + col) : 0.0f;
            s.s_mean[cc] = ok ? __ldg(p.aux_mean + col) : 0.0f;
            s.s_rstd[cc] = ok ? __ldg(p.aux_rstd + col) : 0.0f;
          } else {
            s.s_scale[cc] = (ok && p.col_scale) ? __ldg(p.col_scale + col) : 1.0f;
            s.s_shift[cc] = (ok && p.col_shift) ? __ldg(p.col_shift + col) : 0.0f;
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        staged_n_tile = n_tile;
      }
      if (do_stats && n_tile != st_ntile) {                    // flush the register accumulators of the old N tile
        if (st_ntile >= 0) {
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            const int c = st_ntile * p.block_n + sl * 64 + st_c;
            if (sl * 64 < p.block_n && c < p.n_total) st_row[c] += st_acc[sl];
            st_acc[sl] = 0.0f;
          }
        }
        st_ntile = n_tile;
      }

      const long long cw0 = p.prof ? clock64() : 0;
      mbar_wait(&s.tfull[acc], acc_phase);
      if (p.prof) t_wtfull += clock64() - cw0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * kAccStride + (static_cast<uint32_t>(q * 32) << 16);

      if (p.whatif & 2) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG > 1) mbar_arrive_cluster(tempty_addr0 + acc * 8);
          else mbar_arrive(&s.tempty[acc]);
        }
        continue;
      }
      if (use_tma) {
        // ------------------------------------------------- bf16 NHWC through swizzled smem slabs + TMA stores
        const int nslabs = p.block_n >> 6;
        const int j0 = jt * p.bw, i0 = it * p.bh, b0 = bt * p.bb;
#pragma unroll 1
        for (int sl = 0; sl < nslabs; ++sl, ++store_count) {
          uint8_t* buf = s.staging + (p.nbuf == 2 ? (store_count & 1) : 0) * kStagingBytes;
          float* stat_buf = s.s_stat + (store_count & 1) * 512;
          // Two staging slabs: ONE barrier per slab.  Before barrier B(s) the leader waits until store(s-1) has finished
          // reading the other slab, so passing B(s) tells every thread both "slab s is written" (the leader may store
          // it) and "the buffer of slab s+1 is free".  One staging slab (single-CTA 256-column tiles): wait + barrier
          // before writing as well.
          if (p.nbuf != 2) {
            if (et == 0) bulk_wait_read<0>();
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
          float csum[2] = {0.0f, 0.0f}, csq[2] = {0.0f, 0.0f};
          uint4 ax[8];                                      // this row's 64 aux values of the slab (fused backward)
          if (aux_on) {
            // the slab was TMA-loaded two slabs ago into auxstg[count & 1] (same swizzle as the output slab)
            mbar_wait(&s.auxbar[store_count & 1], (store_count >> 1) & 1u);
            const uint8_t* arow = s.auxstg + (store_count & 1) * kStagingBytes + row * 128;
#pragma unroll
            for (int g = 0; g < 8; ++g)
              ax[g] = row_ok ? *reinterpret_cast<const uint4*>(arow + ((g ^ (row & 7)) << 4)) : make_uint4(0u, 0u, 0u, 0u);
          }
          uint32_t vv[2][32];
          tmem_ld_32x32(taddr + sl * 64, vv[0]);            // both halves in flight before the single wait
          tmem_ld_32x32(taddr + sl * 64 + 32, vv[1]);
          tmem_ld_wait();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int c = sl * 64 + hh * 32;
            uint32_t (&v)[32] = vv[hh];
            if (affine) {
              const float4* sc4 = reinterpret_cast<const float4*>(s.s_scale + c);
              const float4* sh4 = reinterpret_cast<const float4*>(s.s_shift + c);
#pragma unroll
              for (int g4 = 0; g4 < 8; ++g4) {
                const float4 a4 = sc4[g4], b4 = sh4[g4];
                v[g4 * 4 + 0] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 0]), a4.x, b4.x));
                v[g4 * 4 + 1] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 1]), a4.y, b4.y));
                v[g4 * 4 + 2] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 2]), a4.z, b4.z));
                v[g4 * 4 + 3] = __float_as_uint(fmaf(__uint_as_float(v[g4 * 4 + 3]), a4.w, b4.w));
              }
            }
            if (lrelu) {
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const float f = __uint_as_float(v[e]);
                v[e] = __float_as_uint(fmaxf(f, f * p.slope));
              }
            }
            float xh[32];                                   // mode 2: normalised activation of each element
            if (aux_on) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&ax[hh * 4 + g]);
#pragma unroll
                for (int e2 = 0; e2 < 4; ++e2) {
                  const float2 av = __bfloat1622float2(h2[e2]);
                  const int e = g * 8 + e2 * 2;
                  if (aux_bn) {
                    const int pc = (p.merged ? hh * 32 : c) + e;   // merged: every slab holds the SAME 64 channels
                    const float u0 = fmaf(av.x, s.s_scale[pc], s.s_shift[pc]);
                    const float u1 = fmaf(av.y, s.s_scale[pc + 1], s.s_shift[pc + 1]);
                    if (!(u0 > 0.0f)) v[e] = __float_as_uint(__uint_as_float(v[e]) * p.aux_slope);
                    if (!(u1 > 0.0f)) v[e + 1] = __float_as_uint(__uint_as_float(v[e + 1]) * p.aux_slope);
                    xh[e] = (av.x - s.s_mean[pc]) * s.s_rstd[pc];
                    xh[e + 1] = (av.y - s.s_mean[pc + 1]) * s.s_rstd[pc + 1];
                  } else {
                    if (!(av.x > 0.0f)) v[e] = __float_as_uint(__uint_as_float(v[e]) * p.aux_slope);
                    if (!(av.y > 0.0f)) v[e + 1] = __float_as_uint(__uint_as_float(v[e + 1]) * p.aux_slope);
                  }
                }
              }
            }
            // pack to bf16 and write this row's 64 bytes: 16-byte chunk index XOR (row & 7) = the SWIZZLE_128B
            // pattern the output tensor map expects (and conflict-free: 8 rows cover all 8 chunk positions)
            uint8_t* rowp = buf + row * 128;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
              u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
              u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
              u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
              const int chunk = hh * 4 + g;
              *reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4)) = u;
            }
            if (do_stats) {
              float xs[32];
#pragma unroll
              for (int e = 0; e < 32; ++e) xs[e] = row_ok ? bf16_round(__uint_as_float(v[e])) : 0.0f;
              csum[hh] = warp_colsum32(xs, lane);
              if (aux_bn) {
#pragma unroll
                for (int e = 0; e < 32; ++e) xs[e] *= xh[e];          // S(du * xhat)
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) xs[e] *= xs[e];          // S(a^2)
              }
              csq[hh] = warp_colsum32(xs, lane);
            }
          }
          if (sl == nslabs - 1) {          // accumulator drained: hand it back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG > 1) mbar_arrive_cluster(tempty_addr0 + acc * 8);
              else mbar_arrive(&s.tempty[acc]);
            }
          }
          if (do_stats) {
            float* ss = stat_buf + q * 128;
            ss[lane] = csum[0]; ss[32 + lane] = csum[1];
            ss[64 + lane] = csq[0]; ss[96 + lane] = csq[1];
          }
          fence_proxy_async();             // generic-proxy smem writes -> visible to the TMA (async proxy)
          if (p.nbuf == 2 && et == 0) {
            const long long cs0 = p.prof ? clock64() : 0;
            bulk_wait_read<0>();           // store(s-1) no longer reads the slab that slab s+1 will overwrite
            if (p.prof) t_wstore += clock64() - cs0;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0 && tile_ok) {
            if (p.merged) tma_store_4d(&maps.o[sl], buf, 0, j0, i0, b0);      // slab = output phase, all 64 channels
            else tma_store_4d(&maps.o[ph], buf, n0 + sl * 64, j0, i0, b0);
            bulk_commit();
          }
          if (aux_on && et == 0) aux_prefetch();   // every thread is past its reads of auxstg[count & 1]: refill it
          if (do_stats) {                  // fixed-order sum of the four warps' partial column sums
            const float* ss = stat_buf + st_k * 64 + st_c;
            const float t = ((ss[0] + ss[128]) + ss[256]) + ss[384];
            if (sl == 0 || p.merged) st_acc[0] += t;     // merged: the four slabs are four phases of the SAME channels
            else if (sl == 1) st_acc[1] += t;
            else if (sl == 2) st_acc[2] += t;
            else st_acc[3] += t;
          }
        }
        continue;
      }

      // ------------------------------------------------- direct stores (fp32 outputs, narrow bf16 tiles)
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
        } else if (OUT == OUT_F32_NHWC && p.act_tanh) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(tanhf(__uint_as_float(v[e])));
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
      if (lane == 0) {
        if (CG > 1) mbar_arrive_cluster(tempty_addr0 + acc * 8);
        else mbar_arrive(&s.tempty[acc]);
      }
    }
    if (do_stats && st_ntile >= 0) {
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const int c = st_ntile * p.block_n + sl * 64 + st_c;
        if (sl * 64 < p.block_n && c < p.n_total) st_row[c] += st_acc[sl];
      }
    }
    if (use_tma && et == 0) bulk_wait_read<0>();     // smem must outlive the last store's reads
    if (p.prof && et == 0) {
      long long* o = p.prof + static_cast<size_t>(blockIdx.x) * 12;
      o[6] = clock64() - t_begin; o[7] = t_wtfull; o[8] = t_wstore; o[9] = iter;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG > 1) cluster_sync_all();   // the peer may still arrive on this CTA's barriers / its MMAs read this smem
  if (warp == 1) tmem_dealloc_cg<CG>(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// Measured on B200 (profiles/r1_wgrad_whatif.txt): with both operands streamed once from HBM this kernel is bound by
// bytes in flight per SM (smem ring / DRAM latency), not by the tensor pipe -- removing every MMA changes its time by
// < 10 %.  msub = 2 therefore processes a 256-row M tile per unit (two 128-row accumulators sharing every B slab):
// 1.5x the FLOPs per byte of smem ring at the price of the accumulator double-buffering (all 512 TMEM columns).
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_wgrad_kernel(const __grid_constant__ GemmMaps maps, const __grid_constant__ WgradArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const PipeSmem s = carve_smem(smem_raw, 4 * (kAStageBytes + 256 * 128), kWgradStageT);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
  }
  const uint32_t tmem_base = pipeline_prologue<1>(s, warp, 4);
  griddep_wait();
  griddep_launch();

  const int msub = p.msub;                                    // 128-row accumulators per unit (1 or 2)
  const int a_bytes = msub * 16384;                           // msub * 2 slabs of [64 px][64 ch]
  const int stage_bytes = a_bytes + 32768;
  const int nstages = msub == 2 ? 3 : 4;                      // 3 x 64 KiB or 4 x 48 KiB
  const int total_units = p.m_tiles * p.n_tiles * p.splits;
  const uint32_t stage_tx = static_cast<uint32_t>(2 * msub + p.slabs_per_tile) * 8192u;

  if (warp == 0 || warp == 6) {
    // Producers (warp 0: stage bytes + the low-resolution slabs, warp 6: the high-resolution slabs): one thread each paces the pipeline, so the pixel-block loop holds no division, no parameter re-reads and
    // no per-slab tap decoding (all hoisted per unit into registers).
    if (elect_one()) {
      const int m_tiles = p.m_tiles, n_tiles = p.n_tiles, pb_per_split = p.pb_per_split, num_pb = p.num_pb;
      const int tw = p.tw, th = p.th, bw = p.bw, bh = p.bh, bb = p.bb;
      const int spt = p.slabs_per_tile;
      const uint32_t ring0 = smem_u32(s.stages);
      const uint32_t full0 = smem_u32(&s.full[0]), empty0 = smem_u32(&s.empty[0]);
      const uint64_t desc_lo = reinterpret_cast<uint64_t>(&maps.b);
      const uint64_t desc_hi0 = reinterpret_cast<uint64_t>(&maps.a[0]);
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < total_units; u += gridDim.x) {
        const int m_tile = u % m_tiles;
        const int rest = u / m_tiles;
        const int n_tile = rest % n_tiles;
        const int split = rest / n_tiles;
        const int pb0 = split * pb_per_split;
        const int pb1 = min(num_pb, pb0 + pb_per_split);
        uint64_t sdesc[4];
        int sdw[4], sdh[4], scol[4];
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          sdesc[sl] = desc_hi0; sdw[sl] = 0; sdh[sl] = 0; scol[sl] = 0;
          if (sl < spt) {
            int tap, chunk;
            wgrad_slab(p, n_tile, sl, tap, chunk);
            const Tap t = p.taps[tap];
            sdesc[sl] = desc_hi0 + static_cast<uint64_t>(t.map) * sizeof(CUtensorMap);
            sdw[sl] = t.dw; sdh[sl] = t.dh; scol[sl] = chunk * 64;
          }
        }
        const int mcol = m_tile * 128 * msub;
        int jt = pb0 % tw, it = (pb0 / tw) % th, bt = pb0 / (tw * th);
        for (int pb = pb0; pb < pb1; ++pb) {
          const int j0 = jt * bw, i0 = it * bh, b0 = bt * bb;
          mbar_wait_raw(empty0 + stage * 8, phase ^ 1u);
          const uint32_t sa = ring0 + stage * stage_bytes;
          const uint32_t sb = sa + a_bytes;
          const uint32_t fb = full0 + stage * 8;
          if (warp == 0) mbar_expect_tx_raw(fb, stage_tx);
          // MMA-A operand: 64-channel slabs of the low-resolution tensor (channels beyond Cp zero-fill)
#pragma unroll
          for (int sl = 0; sl < 4; ++sl)
            if (warp == 0 && sl < 2 * msub) tma_ld_4d_raw<1>(desc_lo, fb, sa + sl * 8192, mcol + sl * 64, j0, i0, b0);
          // MMA-B operand: one slab per (tap, 64-channel chunk) of the high-resolution tensor
#pragma unroll
          for (int sl = 0; sl < 4; ++sl)
            if (warp == 6 && sl < spt) tma_ld_4d_raw<1>(sdesc[sl], fb, sb + sl * 8192, scol[sl], j0 + sdw[sl], i0 + sdh[sl], b0);
          if (++stage == nstages) { stage = 0; phase ^= 1u; }
          if (++jt == tw) { jt = 0; if (++it == th) { it = 0; ++bt; } }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(kBlockM, 64 * p.slabs_per_tile, 1, 1);
      const uint32_t ring0 = smem_u32(s.stages);
      const uint32_t full0 = smem_u32(&s.full[0]);
      const uint64_t d0 = make_smem_desc_sw128(0, 8192, 1024);
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
          mbar_wait_raw(full0 + stage * 8, phase);
          tc_fence_after();
          const uint32_t sa = ring0 + stage * stage_bytes;
          const uint32_t sb = sa + a_bytes;
          // MN-major SWIZZLE_128B: LBO = stride between 64-element MN slabs (8 KiB),
          // SBO = stride between 8-row K groups (1 KiB); one UMMA_K = 16 pixel rows = 2 KiB.
          const uint64_t db = d0 | static_cast<uint64_t>((sb >> 4) & 0x3FFF);
          for (int ms = 0; ms < msub; ++ms) {
            const uint64_t da = d0 | static_cast<uint64_t>(((sa + ms * 16384) >> 4) & 0x3FFF);
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
        if (p.native_out != nullptr) {
          const float a = p.alpha * (p.alpha_dev ? __ldg(p.alpha_dev) : 1.0f);
          const float beta = p.beta;
          for (int sl = 0; sl < p.slabs_per_tile; ++sl) {
            int tap, chunk;
            wgrad_slab(p, n_tile, sl, tap, chunk);
            // A thread owns one accumulator ROW (TMEM lane); stored as is, every store instruction would touch 32 rows
            // x 16 bytes.  Each 32 x 32 fp32 block therefore goes through a per-warp shared-memory slab (16-byte chunks,
            // chunk index XOR row: conflict-free both ways) and leaves as whole 128-byte row segments -- 8 lanes per row
            // with 16-byte stores, or 16 lanes per row with 8-byte stores when the row pitch is only 8-byte aligned
            // (ragged N of an nn.Linear, e.g. 19198).  These launches are store-bound (K = batch rows): the gradient
            // of betaVAE's 6000 x 19198 layer is 460 MB of fp32.
            const size_t col0 = static_cast<size_t>(tap) * p.ld_out + chunk * 64;
            const bool vec4 = (p.ld_out & 3) == 0;
            float* slab = reinterpret_cast<float*>(s.staging) + q * 1024;
            const int prow0 = (m_tile * msub + ms) * 128 + q * 32;           // first row of this warp's 32
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32];
              tmem_ld_32x32(taddr + sl * 64 + h * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4)
                *reinterpret_cast<float4*>(slab + lane * 32 + ((c4 ^ (lane & 7)) << 2)) =
                    make_float4(a * __uint_as_float(v[c4 * 4 + 0]), a * __uint_as_float(v[c4 * 4 + 1]),
                                a * __uint_as_float(v[c4 * 4 + 2]), a * __uint_as_float(v[c4 * 4 + 3]));
              __syncwarp();
              const int cbase = chunk * 64 + h * 32;                          // first column of this block in the row
              if (vec4) {
                const int c4 = lane & 7;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const int rr = it * 4 + (lane >> 3);
                  if (prow0 + rr < p.Cp && cbase + c4 * 4 + 4 <= p.n_out) {
                    float4 f = *reinterpret_cast<const float4*>(slab + rr * 32 + ((c4 ^ (rr & 7)) << 2));
                    float4* dst = reinterpret_cast<float4*>(p.native_out + static_cast<size_t>(prow0 + rr) * p.num_taps * p.ld_out +
                                                            col0 + h * 32 + c4 * 4);
                    if (beta != 0.0f) {
                      const float4 old = *dst;
                      f.x += beta * old.x; f.y += beta * old.y; f.z += beta * old.z; f.w += beta * old.w;
                    }
                    *dst = f;
                  }
                }
              } else {
                const int c2 = lane & 15;
#pragma unroll
                for (int it = 0; it < 16; ++it) {
                  const int rr = it * 2 + (lane >> 4);
                  if (prow0 + rr < p.Cp && cbase + c2 * 2 + 2 <= p.n_out) {
                    float2 f = *reinterpret_cast<const float2*>(slab + rr * 32 + (((c2 >> 1) ^ (rr & 7)) << 2) + (c2 & 1) * 2);
                    float2* dst = reinterpret_cast<float2*>(p.native_out + static_cast<size_t>(prow0 + rr) * p.num_taps * p.ld_out +
                                                            col0 + h * 32 + c2 * 2);
                    if (beta != 0.0f) {
                      const float2 old = *dst;
                      f.x += beta * old.x; f.y += beta * old.y;
                    }
                    *dst = f;
                  }
                }
              }
              __syncwarp();
            }
          }
        } else if (p.direct_out != nullptr) {
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
