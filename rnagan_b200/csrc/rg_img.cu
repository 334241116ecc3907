// rg_img.cu -- the 3-channel image side of the DCGAN pair without a materialised im2col / col2im buffer:
//
//   rg_img_conv_up     ConvTranspose2d(64, Cimg, 4, 2, 1) (+bias, tanh): generator output layer (src/dcgan.py:82 /
//                      torchgan DCGANGenerator last block) and the critic layer-0 input gradient
//   rg_img_conv_down   Conv2d(Cimg, 64, 4, 2, 1) (+bias, LeakyReLU): critic layer 0 (torchgan DCGANDiscriminator first
//                      block) on the fp32 NCHW image, with the gradient-penalty interpolation (src/wgan_loss.py:376-380)
//                      or the tanh backward fused into the patch load; also the generator output layer's input gradient
//   rg_img_conv_wgrad  the weight (and bias) gradient of either layer
//
// These layers carry 6 GFLOP per pass against 134 MB of activations: HBM is the floor (28 us at B = 64).  Each CTA stages a
// halo'd tile in shared memory ONCE (the old path wrote and re-read a 134-201 MB `col` matrix per pass) -- the bf16
// activation tile through one TMA box load (SWIZZLE_128B, zero fill outside the image), the fp32 image patch through
// 16-byte cp.async, double-buffered in the persistent kernels -- and contracts it with warp-level mma.sync.m16n8k16 bf16
// fragments (fp32 accumulate).  N is 6..12 useful columns here, so tcgen05 would not help: a UMMA re-reads the 4 KB A
// slice from shared memory for every shift, while ldmatrix fragments are reused across the shifts that share a tile row.
// What bounds them (ncu, profiles/r2_ncu_hbm_summary.txt): conv_up the legacy tensor pipe (57-60 % busy) and L1/shared
// bandwidth (77-82 %); conv_down / wgrad instruction issue and 2-way bank conflicts of the stride-2 fp32 tap gather.
#include <algorithm>
#include <type_traits>
#include "rg_host.cuh"
#include "rg_ptx.cuh"
#include <cuda_bf16.h>

namespace rg {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
// D (16x8 fp32) += A (16x16 bf16, row) * B (16x8 bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// hardware tanh (MUFU.TANH): max relative error 2^-11, two orders below the bf16 operand rounding of the contraction
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t bf2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// 16-byte global -> shared copy; bytes == 0 zero-fills (out-of-image halo)
__device__ __forceinline__ void cp16_zfill(uint32_t sdst, const void* gsrc, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

constexpr int kTH = 8, kTW = 32;                  // low-resolution pixels per CTA tile (the 64-channel side)
constexpr int kWgTH = 4;                          // the weight-gradient kernel walks 4 x 32 tiles (two stages fit four CTAs per SM)
constexpr int kThreads = 128;                     // 4 warps: warp w owns tile rows 2w, 2w+1
constexpr int kCtasPerSm = 4;                     // small CTAs, several per SM: one CTA's tile load overlaps the others' math
constexpr int kC = 64;                            // channels of the 64-channel side (step_channels)

__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}

// ------------------------------------------------------------------------------------------------ image patch (fp32)
// One stage holds `planes` planes (the CIMG channels of x, then -- for the modes that need it -- the CIMG channels of y) of
// R = 2*TH + 2 image rows 2*y0-1 .. 2*y0+2*TH; a row is kPitch = 72 floats = 18 aligned float4 covering the image columns
// gxl = gx - 2*x0 in [-4, 67] at kPX + gxl (the taps use [-1, 64]; zero outside the image).  The stage is filled with
// 16-byte cp.async (zero-fill for out-of-image vectors): no registers are staged, so the NEXT tile's patch is in flight
// while the current one is contracted.  Thread t < 126 owns column vector v = t % 18 of rows rr, rr+7, rr+14 (rr = t / 18)
// of EVERY plane -- the thread that copied an x vector also copied the matching y vector, so the transform below needs no
// barrier between the copy and the arithmetic, only the thread's own cp.async.wait_group.
constexpr int kPX = 4, kPitch = 72, kPVec = kPitch / 4;
constexpr int kMaxCimg = 4;

template <int TH, int CIMG>
__device__ __forceinline__ void patch_prefetch(uint32_t stage, const float* __restrict__ x, const float* __restrict__ y,
                                               int b, int S, int y0, int x0, int v, int rr) {
  constexpr int R = 2 * TH + 2;
  if (rr >= 7) return;                                          // threads 126, 127
  const int gx = 2 * x0 - 4 + 4 * v;
  const bool okx = gx >= 0 && gx < S;                           // S and gx are multiples of 4: a vector is all in or all out
  const size_t img0 = static_cast<size_t>(b) * CIMG * S * S;
  const float* xb = x + img0;
  const float* yb = y != nullptr ? y + img0 : nullptr;
#pragma unroll
  for (int k = 0; k < (R + 6) / 7; ++k) {
    const int r = rr + 7 * k;
    if (r < R) {
      const int gy = 2 * y0 - 1 + r;
      const bool ok = okx && gy >= 0 && gy < S;
      const int off = ok ? gy * S + gx : 0;
      const int nbytes = ok ? 16 : 0;
      const uint32_t d = stage + static_cast<uint32_t>((r * kPitch + 4 * v) * 4);
#pragma unroll
      for (int c = 0; c < CIMG; ++c) {
        cp16_zfill(d + c * (R * kPitch * 4), xb + c * S * S + off, nbytes);
        if (yb != nullptr) cp16_zfill(d + (CIMG + c) * (R * kPitch * 4), yb + c * S * S + off, nbytes);
      }
    }
  }
}

__device__ __forceinline__ float img_xform(float t, float y, int mode, float eps, float mul) {
  if (mode == 1) t = eps * t + (1.0f - eps) * y;
  else if (mode == 2) t = t * (1.0f - y * y);
  return t * mul;
}

// mode 0: x * mul;  1: (eps*x + (1-eps)*y) * mul (gradient-penalty interpolate);  2: x * (1 - y^2) * mul (tanh backward,
// y = tanh).  In place on the x planes, by the thread that copied the vectors (after its own cp.async.wait_group);
// zero-filled vectors stay zero.
template <int TH, int CIMG>
__device__ __forceinline__ void patch_transform(float* stage, int mode, float eps, float mul, int v, int rr) {
  constexpr int R = 2 * TH + 2;
  if (rr >= 7) return;
#pragma unroll
  for (int k = 0; k < (R + 6) / 7; ++k) {
    const int r = rr + 7 * k;
    if (r < R) {
#pragma unroll
      for (int c = 0; c < CIMG; ++c) {
        float4* px = reinterpret_cast<float4*>(stage + (c * R + r) * kPitch + 4 * v);
        float4 t = *px;
        float4 yv = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (mode != 0) yv = *reinterpret_cast<const float4*>(stage + ((CIMG + c) * R + r) * kPitch + 4 * v);
        t.x = img_xform(t.x, yv.x, mode, eps, mul);
        t.y = img_xform(t.y, yv.y, mode, eps, mul);
        t.z = img_xform(t.z, yv.z, mode, eps, mul);
        t.w = img_xform(t.w, yv.w, mode, eps, mul);
        *px = t;
      }
    }
  }
}

// =================================================================================================== conv_up (K2)
// out[b, c, 2i+py, 2j+px] = bias[c] + sum_{dy,dx,p} lo[b, i+dy, j+dx, p] * W[p, c, py+1-2dy, px+1-2dx]
// Per 16-pixel M tile and channel chunk: one A fragment per shift (dy,dx), two accumulator tiles (py = 0 / 1) whose 8
// columns are (px, c) pairs; a shift with dy = -1 feeds only py = 0, dy = +1 only py = 1, dy = 0 both: 12 MMAs per chunk.
// The halo'd tile (10 x 34 pixels x 128 B) arrives through ONE TMA box load (SWIZZLE_128B = the chunk ^ (P & 7) layout
// ldmatrix wants; out-of-image pixels are zero-filled by the TMA unit): the per-thread cp.async loop it replaces was a
// third of the kernel's instructions, and the kernel is issue-bound (ncu: 53 % issue-active at 24 % occupancy).
constexpr int kUpPH = kTH + 2, kUpPW = kTW + 2;
constexpr int kUpTileBytes = kUpPH * kUpPW * 128;            // 43520
constexpr int kUpFragBytes = 12 * 4 * 32 * 8;                // (py, shift) x channel chunk x lane x {b0, b1}
constexpr int kUpSmem = 1024 + kUpTileBytes + kUpFragBytes + 16;   // alignment slack + tile + B fragments + mbarrier

// B fragments of conv_up, ready for mma.sync: entry ((combo*4 + kc)*32 + lane) = {b0, b1} with
// combo = py*6 + (dy - dymin(py))*3 + (dx+1) and B[k = channel][n = px*Cimg + c] = W[k][c][py+1-2dy][px+1-2dx]
__global__ void img_up_pack_kernel(const float* __restrict__ Wt, int Cimg, uint2* __restrict__ bfrag) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 12 * 4 * 32) return;
  const int l = idx & 31, kc = (idx >> 5) & 3, combo = idx >> 7;
  const int py = combo / 6, rem = combo - py * 6;
  const int dy = rem / 3 + (py == 0 ? -1 : 0), dx = rem % 3 - 1;
  const int n = l >> 2, kq = l & 3;
  const int px = n / Cimg, c = n - px * Cimg;
  const int kh = py + 1 - 2 * dy, kw = px + 1 - 2 * dx;
  float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (n < 2 * Cimg && kw >= 0 && kw <= 3) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int p = kc * 16 + kq * 2 + (e & 1) + (e >> 1) * 8;
      v[e] = Wt[((static_cast<size_t>(p) * Cimg + c) * 4 + kh) * 4 + kw];
    }
  }
  bfrag[idx] = make_uint2(bf2(v[0], v[1]), bf2(v[2], v[3]));
}

template <int Cimg>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
img_conv_up_kernel(const __grid_constant__ CUtensorMap lomap, const uint2* __restrict__ bfrag_g, const float* __restrict__ bias, void* __restrict__ out, int B, int H,
                   int W, int flags, const float* __restrict__ bn_scale, const float* __restrict__ bn_shift,
                   float bn_slope) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* tile = smem;
  uint2* bfrag = reinterpret_cast<uint2*>(smem + kUpTileBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + kUpTileBytes + kUpFragBytes);
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  const int txs = 31 - __clz(tiles_x), tys = 31 - __clz(tiles_y);          // H, W are powers of two: so are the tile counts
  const int tx = blockIdx.x & (tiles_x - 1), ty = (blockIdx.x >> txs) & (tiles_y - 1), b = blockIdx.x >> (txs + tys);
  const int y0 = ty * kTH, x0 = tx * kTW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
    mbar_expect_tx(bar, kUpTileBytes);
    tma_load_4d(&lomap, bar, tile, 0, x0 - 1, y0 - 1, b);
  }
  {
    const uint32_t fb = smem_u32(bfrag);
    for (int i = threadIdx.x; i < kUpFragBytes / 16; i += kThreads)
      cp16_zfill(fb + i * 16, reinterpret_cast<const uint8_t*>(bfrag_g) + i * 16, 16);
  }
  // fused BatchNorm: this thread's 8-channel chunk is fixed (128 threads step by 16 pixels); fetch its scale / shift
  // while the tile is still in flight
  float sc[8], sh[8];
  if (bn_scale != nullptr) {
    const int ch0 = (threadIdx.x & 7) * 8;
    const float4 sa = __ldg(reinterpret_cast<const float4*>(bn_scale + ch0));
    const float4 sb = __ldg(reinterpret_cast<const float4*>(bn_scale + ch0) + 1);
    const float4 ha = __ldg(reinterpret_cast<const float4*>(bn_shift + ch0));
    const float4 hb = __ldg(reinterpret_cast<const float4*>(bn_shift + ch0) + 1);
    sc[0] = sa.x; sc[1] = sa.y; sc[2] = sa.z; sc[3] = sa.w; sc[4] = sb.x; sc[5] = sb.y; sc[6] = sb.z; sc[7] = sb.w;
    sh[0] = ha.x; sh[1] = ha.y; sh[2] = ha.z; sh[3] = ha.w; sh[4] = hb.x; sh[5] = hb.y; sh[6] = hb.z; sh[7] = hb.w;
  }
  cp_commit_wait_all();
  __syncthreads();                       // B fragments landed; the mbarrier init is visible
  mbar_wait(bar, 0);
  if (bn_scale != nullptr) {
    // `lo` is the PRE-BatchNorm activation a: h = lrelu(scale * a + shift) is applied to the staged tile in place (same
    // arithmetic and bf16 rounding as rg_bn_act, so h is never written to HBM); halo pixels outside the image stay zero
    // the pixel's (row, column) in the tile advances without a division
    const int ch = threadIdx.x & 7;
    int r = 0, c = threadIdx.x >> 3;                       // P = r * kUpPW + c, c < 16 < kUpPW
    for (int P = threadIdx.x >> 3; P < kUpPH * kUpPW; P += kThreads / 8) {
      const int gy = y0 - 1 + r, gx = x0 - 1 + c;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        uint4* qp = reinterpret_cast<uint4*>(tile + P * 128 + ((ch ^ (P & 7)) << 4));
        uint4 v = *qp;
        __nv_bfloat162* vh = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(vh[e]);
          float u0 = fmaf(f.x, sc[2 * e], sh[2 * e]), u1 = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
          u0 = u0 > 0.0f ? u0 : u0 * bn_slope;
          u1 = u1 > 0.0f ? u1 : u1 * bn_slope;
          vh[e] = __floats2bfloat162_rn(u0, u1);
        }
        *qp = v;
      }
      c += kThreads / 8;
      if (c >= kUpPW) { c -= kUpPW; ++r; }
    }
    __syncthreads();
  }

  // Warp w owns the four pixel rows 4*(w >> 1) .. +3 of one 16-pixel half (w & 1): M tile m = row 4*(w >> 1) + m.  Shift
  // (dy, dx) of row m reads halo'd tile row 4*(w >> 1) + m + 1 + dy, so the four rows share SIX tile rows: one ldmatrix
  // per (tile row, dx, channel chunk) feeds every (m, dy) pair that lands on it -- 72 ldmatrix per tile instead of 144
  // (the kernel is bound by shared-memory bandwidth: 512 B per ldmatrix against 1-2 N = 8 MMAs).
  float acc[4][2][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][t][e] = 0.0f;
  // ldmatrix row of this lane inside an M tile: pixel i = (lane & 7) + 8 * ((lane >> 3) & 1), k half = lane >> 4
  const int li = (lane & 7) + ((lane >> 3) & 1) * 8, lk = lane >> 4;
  const uint32_t tile_u32 = smem_u32(tile);
  const int rg4 = 4 * (warp >> 1), xh = (warp & 1) * 16;
  const int Pw = rg4 * kUpPW + xh + 1 + li;            // halo'd tile row rg4, shift dx = 0
#pragma unroll 1
  for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      // B fragments of this (kc, dx): py = 0 takes dy = -1, 0; py = 1 takes dy = 0, +1
      const uint2 b0m = bfrag[((0 * 3 + dx + 1) * 4 + kc) * 32 + lane];     // py 0, dy -1
      const uint2 b0z = bfrag[((1 * 3 + dx + 1) * 4 + kc) * 32 + lane];     // py 0, dy  0
      const uint2 b1z = bfrag[((6 + dx + 1) * 4 + kc) * 32 + lane];         // py 1, dy  0
      const uint2 b1p = bfrag[((6 + 3 + dx + 1) * 4 + kc) * 32 + lane];     // py 1, dy +1
#pragma unroll
      for (int tr = 0; tr < 6; ++tr) {
        const int P = Pw + tr * kUpPW + dx;
        uint32_t a[4];
        ldsm_x4(a, tile_u32 + P * 128 + (((kc * 2 + lk) ^ (P & 7)) << 4));
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int dy = tr - 1 - m;
          if (dy == -1) mma16816(acc[m][0], a, b0m.x, b0m.y);
          if (dy == 0) {
            mma16816(acc[m][0], a, b0z.x, b0z.y);
            mma16816(acc[m][1], a, b1z.x, b1z.y);
          }
          if (dy == 1) mma16816(acc[m][1], a, b1p.x, b1p.y);
        }
      }
    }
  }
  __syncthreads();                       // the activation tile is dead: reuse it as the output staging buffer

  // ---- epilogue: bias, tanh, layout / dtype of the output, staged so that global stores are whole 16-byte vectors.
  // This lane's two accumulator columns n = 2q, 2q+1 are (px, c) pairs with n = px*Cimg + c: both valid iff q < Cimg, and
  // in the NHWC formats they are ADJACENT elements of the staged row (offset 2*Cimg*xl + n), so a lane stores pairs.
  const int OH = 2 * H, OW = 2 * W;
  const bool do_tanh = (flags & 1) != 0, unit = (flags & 2) != 0, u8 = (flags & 4) != 0, bgr = (flags & 8) != 0;
  float* stf = reinterpret_cast<float*>(smem);
  uint8_t* stb = smem;
  if (q < Cimg) {
    const int n0 = 2 * q, n1 = 2 * q + 1;
    const int px0 = n0 / Cimg, c0 = n0 - px0 * Cimg, px1 = n1 / Cimg, c1 = n1 - px1 * Cimg;
    const float nb0 = bias != nullptr ? __ldg(bias + c0) : 0.0f, nb1 = bias != nullptr ? __ldg(bias + c1) : 0.0f;
    // element offsets of this lane for (m = 0, t = 0, h = 0); the loop adds compile-time constants
    const int xg = xh + g;
    int off0, off1;                                   // per-format lane bases of the two columns
    if (u8 || unit) {
      const bool rev = u8 && bgr;
      const int o0 = rev ? px0 * Cimg + (Cimg - 1 - c0) : n0, o1 = rev ? px1 * Cimg + (Cimg - 1 - c1) : n1;
      off0 = (2 * rg4 * 64 + 2 * xg) * Cimg + o0;
      off1 = (2 * rg4 * 64 + 2 * xg) * Cimg + o1;
    } else {
      off0 = (c0 * (2 * kTH) + 2 * rg4) * 64 + 2 * xg + px0;
      off1 = (c1 * (2 * kTH) + 2 * rg4) * 64 + 2 * xg + px1;
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // staged pixel (Yl, Xl) = (2*yl + t, 2*xl + px), yl = 4*(warp >> 1) + m, xl = (warp & 1)*16 + g + 8*h
          constexpr int kRowNhwc = 64 * Cimg;
          const int dY = 2 * m + t, dXl = 8 * h;
          float v0 = acc[m][t][2 * h] + nb0, v1 = acc[m][t][2 * h + 1] + nb1;
          if (do_tanh) { v0 = tanh_fast(v0); v1 = tanh_fast(v1); }
          if (u8) {
            const float u0 = __fmul_rn(__fmul_rn(__fadd_rn(v0, 1.0f), 0.5f), 255.0f);
            const float u1 = __fmul_rn(__fmul_rn(__fadd_rn(v1, 1.0f), 0.5f), 255.0f);
            const uint32_t b0 = __float2uint_rz(fminf(fmaxf(u0, 0.0f), 255.0f));
            const uint32_t b1 = __float2uint_rz(fminf(fmaxf(u1, 0.0f), 255.0f));
            const int d = dY * kRowNhwc + 2 * Cimg * dXl;
            if (bgr) {
              stb[off0 + d] = static_cast<uint8_t>(b0);
              stb[off1 + d] = static_cast<uint8_t>(b1);
            } else {
              *reinterpret_cast<uint16_t*>(stb + off0 + d) = static_cast<uint16_t>(b0 | (b1 << 8));
            }
          } else if (unit) {
            const int d = dY * kRowNhwc + 2 * Cimg * dXl;
            *reinterpret_cast<float2*>(stf + off0 + d) = make_float2((v0 + 1.0f) * 0.5f, (v1 + 1.0f) * 0.5f);
          } else {
            const int d = dY * 64 + 2 * dXl;
            stf[off0 + d] = v0;
            stf[off1 + d] = v1;
          }
        }
      }
    }
  }
  __syncthreads();
  const int Y0 = 2 * y0, X0 = 2 * x0;
  const int xvalid = min(64, OW - X0);           // multiple of 16 (W is a power of two >= 8)
  constexpr int kOutRows = 2 * kTH;
  if (u8) {
    uint8_t* o = static_cast<uint8_t*>(out);
    const int vec_row = 4 * Cimg;                // 16-byte vectors per staged row of 64*Cimg bytes
    for (int Yl = warp; Yl < kOutRows; Yl += kThreads / 32) {
      if (Y0 + Yl >= OH) break;
      for (int v = lane; v < vec_row && v * 16 < xvalid * Cimg; v += 32)
        *reinterpret_cast<uint4*>(o + (((static_cast<size_t>(b) * OH + Y0 + Yl) * OW + X0) * Cimg) + v * 16) =
            *reinterpret_cast<const uint4*>(stb + Yl * 64 * Cimg + v * 16);
    }
  } else if (unit) {
    float* o = static_cast<float*>(out);
    const int vec_row = 16 * Cimg;
    for (int Yl = warp; Yl < kOutRows; Yl += kThreads / 32) {
      if (Y0 + Yl >= OH) break;
      for (int v = lane; v < vec_row && v * 4 < xvalid * Cimg; v += 32)
        *reinterpret_cast<float4*>(o + (((static_cast<size_t>(b) * OH + Y0 + Yl) * OW + X0) * Cimg) + v * 4) =
            *reinterpret_cast<const float4*>(stf + Yl * 64 * Cimg + v * 4);
    }
  } else {
    float* o = static_cast<float*>(out);
    for (int idx = threadIdx.x; idx < Cimg * kOutRows * 16; idx += kThreads) {
      const int v = idx & 15, Yl = (idx >> 4) & (kOutRows - 1), c = idx / (16 * kOutRows);
      if (Y0 + Yl >= OH || v * 4 >= xvalid) continue;
      *reinterpret_cast<float4*>(o + ((static_cast<size_t>(b) * Cimg + c) * OH + Y0 + Yl) * OW + X0 + v * 4) =
          *reinterpret_cast<const float4*>(stf + (c * kOutRows + Yl) * 64 + v * 4);
    }
  }
}

// =================================================================================================== conv_down (K1)
// out[b, y, x, p] = act(bias[p] + sum_{c,kh,kw} img[b, c, 2y-1+kh, 2x-1+kw] * W[p, c, kh, kw])   (bf16 NHWC, 64 channels)
// GEMM view: M = pixels, N = 64, K = CIMG*16 with one k16 step per image channel (k = kh*4 + kw): the A fragment of a lane
// is four float2 loads of horizontally adjacent taps from the fp32 patch, the B fragments (the whole weight) live in
// registers for the CTA's life.  Persistent CTAs, two patch stages: the next tile's cp.async traffic is in flight while
// the current tile is contracted and stored.
constexpr int kDownR = 2 * kTH + 2;
constexpr int kDownStageBytes = (kThreads / 32) * 16 * 128;            // one 16-pixel x 64-channel bf16 tile per warp
__host__ __device__ constexpr int down_patch_bytes(int cimg, int mode) {
  return (mode != 0 ? 2 : 1) * cimg * kDownR * kPitch * 4;
}
__host__ __device__ constexpr int down_smem_bytes(int cimg, int mode) {
  return 2 * down_patch_bytes(cimg, mode) + kDownStageBytes + kC * 4;
}

template <int CIMG>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
img_conv_down_kernel(const float* __restrict__ x, const float* __restrict__ yimg, int mode,
                     const float* __restrict__ eps_dev, const float* __restrict__ mul_dev, const float* __restrict__ Wt,
                     const float* __restrict__ bias, float slope, const __nv_bfloat16* __restrict__ mask_src,
                     float mask_slope, __nv_bfloat16* __restrict__ out, int B, int S) {
  extern __shared__ __align__(16) uint8_t smem_dyn[];
  const int pbytes = down_patch_bytes(CIMG, mode);
  const int H = S / 2, W = S / 2;
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  const int ntiles = B * tiles_x * tiles_y;
  const int txs = 31 - __clz(tiles_x), tys = 31 - __clz(tiles_y);          // S is a power of two: so are the tile counts
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int pv = threadIdx.x % kPVec, prr = threadIdx.x / kPVec;      // this thread's patch vectors (patch_prefetch)
  const float eps = eps_dev ? __ldg(eps_dev) : 0.0f;
  const float mul = mul_dev ? __ldg(mul_dev) : 1.0f;
  const bool xform = mode != 0 || mul != 1.0f;
  const float* ysrc = mode != 0 ? yimg : nullptr;
  uint8_t* stage = smem_dyn + 2 * pbytes + warp * (16 * 128);
  float* sbias = reinterpret_cast<float*>(smem_dyn + 2 * pbytes + kDownStageBytes);
  if (threadIdx.x < kC) sbias[threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.0f;   // read after the first barrier
  // B fragments: B[k = kh*4 + kw (channel c)][n = p]; lane holds k = 2q, 2q+1 (kh = q/2) and k + 8 (kh + 2), n = g
  uint32_t breg[CIMG][8][2];
  const int kh0 = q >> 1, kw0 = (q & 1) * 2;
#pragma unroll
  for (int c = 0; c < CIMG; ++c)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float* wp = Wt + ((static_cast<size_t>(nt * 8 + g) * CIMG + c) * 4 + kh0) * 4 + kw0;
      breg[c][nt][0] = bf2(__ldg(wp), __ldg(wp + 1));
      breg[c][nt][1] = bf2(__ldg(wp + 8), __ldg(wp + 9));
    }
  const uint32_t patch_u32 = smem_u32(smem_dyn);
  auto prefetch = [&](int tile, int s) {
    const int tx = tile & (tiles_x - 1), ty = (tile >> txs) & (tiles_y - 1), b = tile >> (txs + tys);
    patch_prefetch<kTH, CIMG>(patch_u32 + s * pbytes, x, ysrc, b, S, ty * kTH, tx * kTW, pv, prr);
  };
  int it = 0;
  if (static_cast<int>(blockIdx.x) < ntiles) prefetch(blockIdx.x, 0);
  cp_commit();
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int cur = it & 1;
    const int nxt = tile + gridDim.x;
    if (nxt < ntiles) prefetch(nxt, cur ^ 1);            // stage cur^1 was released by the barrier that ended tile it-1
    cp_commit();                                         // (possibly empty) group: the count per iteration is uniform
    cp_wait<1>();                                        // this thread's copies of the current stage have landed
    float* patch = reinterpret_cast<float*>(smem_dyn + cur * pbytes);
    if (xform) patch_transform<kTH, CIMG>(patch, mode, eps, mul, pv, prr);
    __syncthreads();
    const int tx = tile & (tiles_x - 1), ty = (tile >> txs) & (tiles_y - 1), b = tile >> (txs + tys);
    const int y0 = ty * kTH, x0 = tx * kTW;
    // this lane's 16-byte store slot: pixel (lane >> 3) of a group of four, channel chunk lane & 7
    const size_t obase = ((static_cast<size_t>(b) * H + y0) * W + x0 + (lane >> 3)) * kC + (lane & 7) * 8;
    const int xlim = W - x0 - (lane >> 3), ylim = H - y0;
#pragma unroll 1
    for (int m = 0; m < 4; ++m) {
      const int yl = 2 * warp + (m >> 1), xl0 = (m & 1) * 16;
      float acc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
#pragma unroll
      for (int c = 0; c < CIMG; ++c) {
        // tap (kh, kw) of output pixel (yl, xl) sits at patch row 2*yl + kh, column gxl = 2*xl - 1 + kw
        const float* pr = patch + (c * kDownR + 2 * yl + kh0) * kPitch + kPX - 1 + 2 * (xl0 + g) + kw0;
        const uint32_t a[4] = {bf2(pr[0], pr[1]),                                     // row g,     kh0
                               bf2(pr[16], pr[17]),                                   // row g + 8, kh0
                               bf2(pr[2 * kPitch], pr[2 * kPitch + 1]),               // row g,     kh0 + 2
                               bf2(pr[2 * kPitch + 16], pr[2 * kPitch + 17])};        // row g + 8, kh0 + 2
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mma16816(acc[nt], a, breg[c][nt][0], breg[c][nt][1]);
      }
      // epilogue: bias + LeakyReLU, bf16, swizzled per-warp staging (16 pixels x 128 B), then 16-byte global stores
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 bb = *reinterpret_cast<const float2*>(sbias + nt * 8 + 2 * q);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v0 = acc[nt][2 * h] + bb.x, v1 = acc[nt][2 * h + 1] + bb.y;
          v0 = v0 > 0.0f ? v0 : v0 * slope;
          v1 = v1 > 0.0f ? v1 : v1 * slope;
          const int row = g + 8 * h;
          *reinterpret_cast<uint32_t*>(stage + row * 128 + ((nt ^ (row & 7)) << 4) + q * 4) = bf2(v0, v1);
        }
      }
      __syncwarp();
      if (yl < ylim) {
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int row = i4 * 4 + (lane >> 3), ch = lane & 7;
          if (xl0 + i4 * 4 < xlim) {
            uint4 v = *reinterpret_cast<const uint4*>(stage + row * 128 + ((ch ^ (row & 7)) << 4));
            const size_t o = obase + static_cast<size_t>((yl * W + xl0 + i4 * 4) * kC);
            if (mask_src != nullptr) {        // LeakyReLU backward mask from a stored activation: out *= (m > 0 ? 1 : slope)
              const uint4 mk = *reinterpret_cast<const uint4*>(mask_src + o);
              const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&mk);
              __nv_bfloat162* vh = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 mf = __bfloat1622float2(mh[e]);
                float2 vf = __bfloat1622float2(vh[e]);
                vf.x *= mf.x > 0.0f ? 1.0f : mask_slope;
                vf.y *= mf.y > 0.0f ? 1.0f : mask_slope;
                vh[e] = __floats2bfloat162_rn(vf.x, vf.y);
              }
            }
            *reinterpret_cast<uint4*>(out + o) = v;
          }
        }
      }
      __syncwarp();
    }
    __syncthreads();                                            // every warp is done with this stage's patch
  }
  cp_wait<0>();
}

// =================================================================================================== wgrad (K3)
// part[cta][p][n] = sum over the CTA's tiles and pixels of act[b, y, x, p] * img'[b, c, 2y-1+kh, 2x-1+kw],
// n = (c*2 + kh/2)*8 + (kh%2)*4 + kw; column 2*CIMG*8 holds sum act (the bias gradient of the 64-channel side).
// GEMM view: M = p (64), N = taps, K = pixels.  A (p x pixel) comes transposed out of the activation tile with
// ldmatrix.trans; B (pixel x tap) is gathered from the fp32 patch.  Warp w: pixel rows {w>>1, (w>>1)+2} of the 4 x 32
// tile, n tiles (w&1)*4 .. +3.  Fixed order everywhere: bit-reproducible.  Two stages: the activation tile of the next
// tile arrives through one TMA box load (SWIZZLE_128B = the layout ldmatrix wants), its image patch through cp.async,
// both in flight while the current tile is contracted.
constexpr int kWgR = 2 * kWgTH + 2;
constexpr int kWgActBytes = kWgTH * kTW * 128;                 // 16384 per stage; the two stages later hold the 32 KiB reduction
constexpr int kWgN = 64;                                       // padded number of output columns
__host__ __device__ constexpr int wg_patch_bytes(int cimg, int mode) {
  return (mode != 0 ? 2 : 1) * cimg * kWgR * kPitch * 4;
}
__host__ __device__ constexpr int wg_smem_bytes(int cimg, int mode) {
  return 1024 + 2 * kWgActBytes + 2 * wg_patch_bytes(cimg, mode) + 16;
}

template <int CIMG>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
img_conv_wgrad_kernel(const __grid_constant__ CUtensorMap amap, const float* __restrict__ x, const float* __restrict__ yimg, int mode,
                      const float* __restrict__ eps_dev, const float* __restrict__ mul_dev, float* __restrict__ part,
                      int B, int S) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  const int pbytes = wg_patch_bytes(CIMG, mode);
  uint8_t* patch_base = smem + 2 * kWgActBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(patch_base + 2 * pbytes);
  const int H = S / 2, W = S / 2;
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kWgTH - 1) / kWgTH;
  const int ntiles = B * tiles_x * tiles_y;
  const int txs = 31 - __clz(tiles_x), tys = 31 - __clz(tiles_y);          // S is a power of two: so are the tile counts
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int nh = warp & 1, kg = warp >> 1;
  const int pv = threadIdx.x % kPVec, prr = threadIdx.x / kPVec;
  const float eps = eps_dev ? __ldg(eps_dev) : 0.0f;
  const float mul = mul_dev ? __ldg(mul_dev) : 1.0f;
  const bool xform = mode != 0 || mul != 1.0f;
  const float* ysrc = mode != 0 ? yimg : nullptr;
  constexpr int bias_nt = 2 * CIMG;                             // n tile whose column 0 is the all-ones tap
  float acc[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.0f;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const uint32_t patch_u32 = smem_u32(patch_base);
  auto issue = [&](int tile, int s) {
    const int tx = tile & (tiles_x - 1), ty = (tile >> txs) & (tiles_y - 1), b = tile >> (txs + tys);
    const int y0 = ty * kWgTH, x0 = tx * kTW;
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar[s], kWgActBytes);
      tma_load_4d(&amap, &bar[s], smem + s * kWgActBytes, 0, x0, y0, b);
    }
    patch_prefetch<kWgTH, CIMG>(patch_u32 + s * pbytes, x, ysrc, b, S, y0, x0, pv, prr);
  };
  // ldmatrix.trans addressing: lanes 0-7 / 8-15 -> pixels 0-7, channel chunk 2m / 2m+1; lanes 16-31 -> pixels 8-15
  const int lpx = (lane & 7) + (lane >> 4) * 8, lch = (lane >> 3) & 1;
  const uint32_t one_bf2 = bf2(1.0f, 1.0f);
  int it = 0;
  if (static_cast<int>(blockIdx.x) < ntiles) issue(blockIdx.x, 0);
  cp_commit();
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int cur = it & 1;
    const int nxt = tile + gridDim.x;
    if (nxt < ntiles) issue(nxt, cur ^ 1);               // stage cur^1 was released by the barrier that ended tile it-1
    cp_commit();
    cp_wait<1>();
    float* patch = reinterpret_cast<float*>(patch_base + cur * pbytes);
    if (xform) patch_transform<kWgTH, CIMG>(patch, mode, eps, mul, pv, prr);
    mbar_wait(&bar[cur], (it >> 1) & 1);
    __syncthreads();
    const uint32_t atile_u32 = smem_u32(smem + cur * kWgActBytes);
#pragma unroll 1
    for (int yl = kg; yl < kWgTH; yl += kThreads / 64) {
#pragma unroll 1
      for (int xs = 0; xs < 2; ++xs) {
        const int xl0 = xs * 16;
        uint32_t a[4][4];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int P = yl * kTW + xl0 + lpx;
          ldsm_x4_trans(a[m], atile_u32 + P * 128 + (((2 * m + lch) ^ (P & 7)) << 4));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int nt = nh * 4 + j;
          uint32_t b0, b1;
          if (nt < bias_nt) {
            const int c = nt >> 1, kh = (nt & 1) * 2 + (g >> 2), kw = g & 3;
            const float* pr = patch + (c * kWgR + 2 * yl + kh) * kPitch + kPX - 1 + 2 * (xl0 + 2 * q) + kw;
            b0 = bf2(pr[0], pr[2]);                             // pixels xl0+2q, xl0+2q+1
            b1 = bf2(pr[16], pr[18]);                           // pixels +8
          } else if (nt == bias_nt) {
            b0 = b1 = (g == 0) ? one_bf2 : 0u;
          } else {
            continue;
          }
#pragma unroll
          for (int m = 0; m < 4; ++m) mma16816(acc[m][j], a[m], b0, b1);
        }
      }
    }
    __syncthreads();                                            // every warp is done with this stage
  }
  cp_wait<0>();
  // cross-warp reduction over the two pixel groups (fixed order), then one [64][64] partial per CTA
  float* red = reinterpret_cast<float*>(smem);                  // [2 kg][64 p][64 n] = 32 KiB (the two activation stages)
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int p = m * 16 + g + (e >> 1) * 8, n = (nh * 4 + j) * 8 + 2 * q + (e & 1);
        red[(kg * 64 + p) * kWgN + n] = acc[m][j][e];
      }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * kWgN; i += kThreads)
    part[static_cast<size_t>(blockIdx.x) * 64 * kWgN + i] = red[i] + red[64 * kWgN + i];
}

// dW[p][c][kh][kw] = acc*dW + sum_cta part;  dbias[p] = acc_b*dbias + sum_cta part[..][p][2*Cimg*8]
// One block per output row p: thread (sub, n) sums every 16th partial of column n (256-byte coalesced reads across n, two
// loads in flight), the sixteen sub-sums are combined in a fixed order -- the ~600 partials of 16 KiB are read in a few
// microseconds instead of by one serial strided loop per output element.
constexpr int kFinSubs = 16;
__global__ void __launch_bounds__(kFinSubs * kWgN) img_wgrad_finish_kernel(const float* __restrict__ part, int nparts,
                                                                           int Cimg, float* __restrict__ dW, float acc,
                                                                           float* __restrict__ dbias, float acc_b) {
  __shared__ float sm[kFinSubs][kWgN];
  const int p = blockIdx.x, n = threadIdx.x & (kWgN - 1), sub = threadIdx.x >> 6;
  float s0 = 0.0f, s1 = 0.0f;
  int r = sub;
  for (; r + kFinSubs < nparts; r += 2 * kFinSubs) {
    s0 += part[(static_cast<size_t>(r) * 64 + p) * kWgN + n];
    s1 += part[(static_cast<size_t>(r + kFinSubs) * 64 + p) * kWgN + n];
  }
  if (r < nparts) s0 += part[(static_cast<size_t>(r) * 64 + p) * kWgN + n];
  sm[sub][n] = s0 + s1;
  __syncthreads();
  if (sub != 0) return;
  float s = 0.0f;
#pragma unroll
  for (int l = 0; l < kFinSubs; ++l) s += sm[l][n];
  const int nt = n >> 3, e = n & 7;
  if (nt < 2 * Cimg) {
    const int c = nt >> 1, kh = (nt & 1) * 2 + (e >> 2), kw = e & 3;
    float* o = dW + (static_cast<size_t>(p) * Cimg + c) * 16 + kh * 4 + kw;
    *o = (acc != 0.0f ? acc * *o : 0.0f) + s;
  } else if (nt == 2 * Cimg && e == 0 && dbias != nullptr) {
    dbias[p] = (acc_b != 0.0f ? acc_b * dbias[p] : 0.0f) + s;
  }
}

// ------------------------------------------------------------------------------------------------ host side
// Per (kernel instantiation, patch-mode class): opt into the dynamic shared-memory size once and ask the runtime how many
// CTAs fit per SM (the modes that stage y as well have twice the patch bytes: three CTAs per SM instead of four).
struct ImgLaunch {
  int ctas_per_sm = 0;      // 0: not initialised
};

template <typename K>
static int img_prepare(ImgLaunch& L, K kernel, int smem_max, int smem) {
  if (L.ctas_per_sm > 0) return 0;
  RG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
  int n = 0;
  RG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem));
  if (n < 1) {
    set_error("image-side kernel does not fit an SM with %d bytes of shared memory", smem);
    return RG_EINVAL;
  }
  L.ctas_per_sm = std::min(n, kCtasPerSm);
  return 0;
}

// TMA view of the bf16 NHWC activation [B][H][W][64] with a (64, bw, bh, 1) box, SWIZZLE_128B, zero fill outside the
// tensor (boxes larger than a small test image are fine: the whole overhang is zero-filled).
static int img_act_map(CUtensorMap* m, const void* act, int B, int H, int W, int bw, int bh) {
  memset(m, 0, sizeof(*m));
  return encode_map_4d(m, act, kC, W, H, B, kC, static_cast<uint64_t>(W) * kC, static_cast<uint64_t>(H) * W * kC, kC, bw,
                       bh, 1);
}

template <int CIMG>
static int launch_up(const void* lo, const void* wfrag, const float* bias, int flags, int B, int H, int Wd, void* out,
                     const float* bn_scale, const float* bn_shift, float bn_slope, cudaStream_t s) {
  static ImgLaunch L;
  if (int rc = img_prepare(L, img_conv_up_kernel<CIMG>, kUpSmem, kUpSmem)) return rc;
  CUtensorMap map;
  if (int rc = img_act_map(&map, lo, B, H, Wd, kUpPW, kUpPH)) return rc;
  const int grid = B * ceil_div(H, kTH) * ceil_div(Wd, kTW);
  img_conv_up_kernel<CIMG><<<grid, kThreads, kUpSmem, s>>>(map, static_cast<const uint2*>(wfrag), bias, out, B, H, Wd,
                                                           flags, bn_scale, bn_shift, bn_slope);
  RG_LAUNCH_CHECK("rg_img_conv_up");
  return 0;
}

template <int CIMG>
static int launch_down(const float* x, const float* y, int mode, const float* eps_dev, const float* mul_dev,
                       const float* W, const float* bias, float slope, const void* mask_src, float mask_slope, int B,
                       int S, void* out, cudaStream_t s) {
  static ImgLaunch L[2];
  const int smem = down_smem_bytes(CIMG, mode);
  ImgLaunch& l = L[mode != 0];
  if (int rc = img_prepare(l, img_conv_down_kernel<CIMG>, down_smem_bytes(CIMG, 1), smem)) return rc;
  const int grid = std::min(B * ceil_div(S / 2, kTH) * ceil_div(S / 2, kTW), l.ctas_per_sm * num_sms());
  img_conv_down_kernel<CIMG><<<grid, kThreads, smem, s>>>(x, y, mode, eps_dev, mul_dev, W, bias, slope,
                                                          static_cast<const __nv_bfloat16*>(mask_src), mask_slope,
                                                          static_cast<__nv_bfloat16*>(out), B, S);
  RG_LAUNCH_CHECK("rg_img_conv_down");
  return 0;
}

template <int CIMG>
static int launch_wgrad(const void* act, const float* x, const float* y, int mode, const float* eps_dev,
                        const float* mul_dev, int B, int S, void* ws, size_t ws_bytes, int* grid_out, cudaStream_t s) {
  static ImgLaunch L[2];
  const int smem = wg_smem_bytes(CIMG, mode);
  ImgLaunch& l = L[mode != 0];
  if (int rc = img_prepare(l, img_conv_wgrad_kernel<CIMG>, wg_smem_bytes(CIMG, 1), smem)) return rc;
  const int H = S / 2;
  CUtensorMap map;
  if (int rc = img_act_map(&map, act, B, H, H, kTW, kWgTH)) return rc;
  const int ntiles = B * ceil_div(H, kWgTH) * ceil_div(H, kTW);
  const int grid = std::min(ntiles, l.ctas_per_sm * num_sms());
  if (ws_bytes < static_cast<size_t>(grid) * 64 * kWgN * sizeof(float)) {
    set_error("rg_img_conv_wgrad: workspace too small (need %zu bytes)", static_cast<size_t>(grid) * 64 * kWgN * 4);
    return RG_EWORKSPACE;
  }
  img_conv_wgrad_kernel<CIMG><<<grid, kThreads, smem, s>>>(map, x, y, mode, eps_dev, mul_dev, static_cast<float*>(ws), B,
                                                           S);
  RG_LAUNCH_CHECK("rg_img_conv_wgrad");
  *grid_out = grid;
  return 0;
}

}  // namespace rg

using namespace rg;

extern "C" {

size_t rg_img_conv_up_pack_bytes(void) { return kUpFragBytes; }

int rg_img_conv_up_pack(const float* W, int Cp, int Cimg, void* wfrag, rg_stream_t st) {
  RG_CHECK_ARG(W && wfrag && Cp == kC && Cimg >= 1 && Cimg <= kMaxCimg,
               "rg_img_conv_up_pack: need 64 input channels and 1..4 image channels (Cp=%d Cimg=%d)", Cp, Cimg);
  img_up_pack_kernel<<<ceil_div(12 * 4 * 32, 256), 256, 0, static_cast<cudaStream_t>(st)>>>(W, Cimg,
                                                                                          static_cast<uint2*>(wfrag));
  RG_LAUNCH_CHECK("rg_img_conv_up_pack");
  return 0;
}

int rg_img_conv_up(const void* lo, const void* W, const float* bias, int flags, int B, int H, int Wd, int Cp, int Cimg,
                   void* out, const float* bn_scale, const float* bn_shift, float bn_slope, rg_stream_t st) {
  RG_CHECK_ARG(lo && W && out && B > 0 && Cp == kC && Cimg >= 1 && Cimg <= kMaxCimg && is_pow2(H) && is_pow2(Wd) &&
                   H >= 8 && Wd >= 8 && ((bn_scale == nullptr) == (bn_shift == nullptr)),
               "rg_img_conv_up: need 64 input channels, 1..4 image channels and power-of-two H, W >= 8 (Cp=%d Cimg=%d H=%d W=%d)",
               Cp, Cimg, H, Wd);
  cudaStream_t s = static_cast<cudaStream_t>(st);
  switch (Cimg) {
    case 1: return launch_up<1>(lo, W, bias, flags, B, H, Wd, out, bn_scale, bn_shift, bn_slope, s);
    case 2: return launch_up<2>(lo, W, bias, flags, B, H, Wd, out, bn_scale, bn_shift, bn_slope, s);
    case 3: return launch_up<3>(lo, W, bias, flags, B, H, Wd, out, bn_scale, bn_shift, bn_slope, s);
    default: return launch_up<4>(lo, W, bias, flags, B, H, Wd, out, bn_scale, bn_shift, bn_slope, s);
  }
}

int rg_img_conv_down(const float* x, const float* y, int mode, const float* eps_dev, const float* mul_dev,
                     const float* W, const float* bias, float slope, const void* mask_src, float mask_slope, int B,
                     int Cimg, int S, int Cp, void* out, rg_stream_t st) {
  RG_CHECK_ARG(x && W && out && B > 0 && Cp == kC && Cimg >= 1 && Cimg <= kMaxCimg && is_pow2(S) && S >= 16 &&
                   mode >= 0 && mode <= 2 && (mode == 0 || y != nullptr) && (mode != 1 || eps_dev != nullptr),
               "rg_img_conv_down: need 64 output channels, 1..4 image channels, a power-of-two side >= 16, mode 0..2 "
               "(Cp=%d Cimg=%d S=%d mode=%d)", Cp, Cimg, S, mode);
  cudaStream_t s = static_cast<cudaStream_t>(st);
  switch (Cimg) {
    case 1: return launch_down<1>(x, y, mode, eps_dev, mul_dev, W, bias, slope, mask_src, mask_slope, B, S, out, s);
    case 2: return launch_down<2>(x, y, mode, eps_dev, mul_dev, W, bias, slope, mask_src, mask_slope, B, S, out, s);
    case 3: return launch_down<3>(x, y, mode, eps_dev, mul_dev, W, bias, slope, mask_src, mask_slope, B, S, out, s);
    default: return launch_down<4>(x, y, mode, eps_dev, mul_dev, W, bias, slope, mask_src, mask_slope, B, S, out, s);
  }
}

size_t rg_img_conv_wgrad_ws_bytes(void) {
  return static_cast<size_t>(kCtasPerSm * num_sms()) * 64 * kWgN * sizeof(float);
}

int rg_img_conv_wgrad(const void* act, const float* x, const float* y, int mode, const float* eps_dev,
                      const float* mul_dev, int B, int Cimg, int S, int Cp, void* ws, size_t ws_bytes, float* dW,
                      float acc, float* dbias, float acc_bias, rg_stream_t st) {
  RG_CHECK_ARG(act && x && dW && ws && B > 0 && Cp == kC && Cimg >= 1 && Cimg <= kMaxCimg && is_pow2(S) && S >= 16 &&
                   mode >= 0 && mode <= 2 && (mode == 0 || y != nullptr) && (mode != 1 || eps_dev != nullptr) &&
                   (dbias == nullptr || Cimg <= 3),
               "rg_img_conv_wgrad: need 64 channels, 1..4 image channels (<= 3 with a fused bias gradient), a power-of-two "
               "side >= 16, mode 0..2 (Cp=%d Cimg=%d S=%d mode=%d)", Cp, Cimg, S, mode);
  cudaStream_t s = static_cast<cudaStream_t>(st);
  int grid = 0, rc = 0;
  switch (Cimg) {
    case 1: rc = launch_wgrad<1>(act, x, y, mode, eps_dev, mul_dev, B, S, ws, ws_bytes, &grid, s); break;
    case 2: rc = launch_wgrad<2>(act, x, y, mode, eps_dev, mul_dev, B, S, ws, ws_bytes, &grid, s); break;
    case 3: rc = launch_wgrad<3>(act, x, y, mode, eps_dev, mul_dev, B, S, ws, ws_bytes, &grid, s); break;
    default: rc = launch_wgrad<4>(act, x, y, mode, eps_dev, mul_dev, B, S, ws, ws_bytes, &grid, s); break;
  }
  if (rc) return rc;
  img_wgrad_finish_kernel<<<64, kFinSubs * kWgN, 0, s>>>(static_cast<const float*>(ws), grid, Cimg, dW, acc, dbias, acc_bias);
  RG_LAUNCH_CHECK("rg_img_conv_wgrad(finish)");
  return 0;
}

}  // extern "C"
