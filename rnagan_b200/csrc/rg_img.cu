// rg_img.cu -- the 3-channel image side of the DCGAN pair without a materialised im2col / col2im buffer:
//
//   rg_img_conv_up     ConvTranspose2d(64, Cimg, 4, 2, 1) (+bias, tanh): generator output layer (src/dcgan.py:82 /
//                      torchgan DCGANGenerator last block) and the critic layer-0 input gradient
//   rg_img_conv_down   Conv2d(Cimg, 64, 4, 2, 1) (+bias, LeakyReLU): critic layer 0 (torchgan DCGANDiscriminator first
//                      block) on the fp32 NCHW image, with the gradient-penalty interpolation (src/wgan_loss.py:376-380)
//                      or the tanh backward fused into the patch load; also the generator output layer's input gradient
//   rg_img_conv_wgrad  the weight (and bias) gradient of either layer
//
// These layers carry 6 GFLOP per pass against 134 MB of activations: HBM-bound.  Each CTA stages a halo'd tile in shared
// memory ONCE (the old path wrote and re-read a 134-201 MB `col` matrix per pass) and contracts it with warp-level
// mma.sync.m16n8k16 bf16 fragments (fp32 accumulate) -- the tensor pipe is idle >90 % of the time here either way, so the
// point of the MMA is only to keep the arithmetic off the critical path; tcgen05 / TMEM would buy nothing.
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
constexpr int kThreads = 128;                     // 4 warps: warp w owns tile rows 2w, 2w+1
constexpr int kCtasPerSm = 4;                     // small CTAs, several per SM: one CTA's tile load overlaps the others' math
constexpr int kC = 64;                            // channels of the 64-channel side (step_channels)

// ------------------------------------------------------------------------------------------------ image patch (fp32)
// patch[c][r][kPX + gxl]: image rows 2*y0-1 .. 2*y0+2*kTH (r = 0 .. kPR-1), columns gxl = gx - 2*x0 in [-1, 2*kTW] (zero
// outside the image), optionally transformed while loading.  The 64 interior columns of a row are 16 aligned float4 loads
// (all of a thread's loads are issued before the first use), the two halo columns are scalars.
constexpr int kPR = 2 * kTH + 2, kPX = 4, kPitch = 72;
constexpr int kMaxCimg = 4;

__device__ __forceinline__ float img_xform(float t, float y, int mode, float eps, float mul) {
  if (mode == 1) t = eps * t + (1.0f - eps) * y;
  else if (mode == 2) t = t * (1.0f - y * y);
  return t * mul;
}

// mode 0: x * mul;  1: eps*x + (1-eps)*y (gradient-penalty interpolate);  2: x * (1 - y^2) (tanh backward, y = tanh)
__device__ __forceinline__ void load_img_patch(float* patch, const float* __restrict__ x, const float* __restrict__ y,
                                               int mode, float eps, float mul, int b, int Cimg, int S, int y0, int x0) {
  const int nrows = Cimg * kPR;
  const int nvec = nrows * 16;
  const size_t img0 = static_cast<size_t>(b) * Cimg * S * S;
  for (int base = 0; base < nvec; base += kThreads * 4) {
    float4 xv[4], yv[4];
    int dst[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = base + j * kThreads + threadIdx.x;
      xv[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      yv[j] = xv[j];
      dst[j] = -1;
      if (i < nvec) {
        const int row = i >> 4, v = i & 15;
        const int c = row / kPR, r = row - c * kPR;
        const int gy = 2 * y0 - 1 + r, gx = 2 * x0 + 4 * v;
        dst[j] = row * kPitch + kPX + 4 * v;
        if (gy >= 0 && gy < S && gx < S) {
          const size_t o = img0 + (static_cast<size_t>(c) * S + gy) * S + gx;
          xv[j] = __ldg(reinterpret_cast<const float4*>(x + o));
          if (mode != 0) yv[j] = __ldg(reinterpret_cast<const float4*>(y + o));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (dst[j] >= 0) {
        float4 t;
        t.x = img_xform(xv[j].x, yv[j].x, mode, eps, mul);
        t.y = img_xform(xv[j].y, yv[j].y, mode, eps, mul);
        t.z = img_xform(xv[j].z, yv[j].z, mode, eps, mul);
        t.w = img_xform(xv[j].w, yv[j].w, mode, eps, mul);
        *reinterpret_cast<float4*>(patch + dst[j]) = t;
      }
    }
  }
  for (int i = threadIdx.x; i < nrows * 2; i += kThreads) {
    const int row = i >> 1, gxl = (i & 1) ? 2 * kTW : -1;
    const int c = row / kPR, r = row - c * kPR;
    const int gy = 2 * y0 - 1 + r, gx = 2 * x0 + gxl;
    float t = 0.0f;
    if (gy >= 0 && gy < S && gx >= 0 && gx < S) {
      const size_t o = img0 + (static_cast<size_t>(c) * S + gy) * S + gx;
      t = img_xform(__ldg(x + o), mode != 0 ? __ldg(y + o) : 0.0f, mode, eps, mul);
    }
    patch[row * kPitch + kPX + gxl] = t;
  }
}

// ------------------------------------------------------------------------------------------------ activation tile (bf16)
// tile[P][64 channels] with P = r * pw + c over a (ph x pw) pixel window whose top-left pixel is (ya, xa); 128-byte rows,
// 16-byte chunk index XOR (P & 7): ldmatrix over 8 consecutive pixels is conflict-free.  Out-of-image pixels are zero.
__device__ __forceinline__ void load_act_tile(uint8_t* tile, const __nv_bfloat16* __restrict__ act, int b, int H, int W,
                                              int ya, int xa, int ph, int pw) {
  const uint32_t base = smem_u32(tile);
  for (int idx = threadIdx.x; idx < ph * pw * 8; idx += kThreads) {
    const int P = idx >> 3, ch = idx & 7;
    const int r = P / pw, c = P - r * pw;
    const int gy = ya + r, gx = xa + c;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const __nv_bfloat16* src = act + ((static_cast<size_t>(b) * H + (ok ? gy : 0)) * W + (ok ? gx : 0)) * kC + ch * 8;
    cp16_zfill(base + P * 128 + ((ch ^ (P & 7)) << 4), src, ok ? 16 : 0);
  }
}

// =================================================================================================== conv_up (K2)
// out[b, c, 2i+py, 2j+px] = bias[c] + sum_{dy,dx,p} lo[b, i+dy, j+dx, p] * W[p, c, py+1-2dy, px+1-2dx]
// Per 16-pixel M tile and channel chunk: one A fragment per shift (dy,dx), two accumulator tiles (py = 0 / 1) whose 8
// columns are (px, c) pairs; a shift with dy = -1 feeds only py = 0, dy = +1 only py = 1, dy = 0 both: 12 MMAs per chunk.
constexpr int kUpPH = kTH + 2, kUpPW = kTW + 2;
constexpr int kUpTileBytes = kUpPH * kUpPW * 128;            // 43520
constexpr int kUpFragBytes = 12 * 4 * 32 * 8;                // (py, shift) x channel chunk x lane x {b0, b1}
constexpr int kUpSmem = kUpTileBytes + kUpFragBytes;

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
img_conv_up_kernel(const __nv_bfloat16* __restrict__ lo, const uint2* __restrict__ bfrag_g, const float* __restrict__ bias,
                   void* __restrict__ out, int B, int H, int W, int flags, const float* __restrict__ bn_scale,
                   const float* __restrict__ bn_shift, float bn_slope) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* tile = smem;
  uint2* bfrag = reinterpret_cast<uint2*>(smem + kUpTileBytes);
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  const int tx = blockIdx.x % tiles_x, ty = (blockIdx.x / tiles_x) % tiles_y, b = blockIdx.x / (tiles_x * tiles_y);
  const int y0 = ty * kTH, x0 = tx * kTW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;

  load_act_tile(tile, lo, b, H, W, y0 - 1, x0 - 1, kUpPH, kUpPW);
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
  __syncthreads();
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
  int P0[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) P0[m] = (2 * warp + (m >> 1) + 1) * kUpPW + (m & 1) * 16 + 1 + li;
#pragma unroll 1
  for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        uint2 b0 = make_uint2(0u, 0u), b1 = make_uint2(0u, 0u);
        if (dy <= 0) b0 = bfrag[(((dy + 1) * 3 + dx + 1) * 4 + kc) * 32 + lane];
        if (dy >= 0) b1 = bfrag[((6 + dy * 3 + dx + 1) * 4 + kc) * 32 + lane];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int P = P0[m] + dy * kUpPW + dx;
          uint32_t a[4];
          ldsm_x4(a, tile_u32 + P * 128 + (((kc * 2 + lk) ^ (P & 7)) << 4));
          if (dy <= 0) mma16816(acc[m][0], a, b0.x, b0.y);
          if (dy >= 0) mma16816(acc[m][1], a, b1.x, b1.y);
        }
      }
    }
  }
  __syncthreads();                       // the activation tile is dead: reuse it as the output staging buffer

  // ---- epilogue: bias, tanh, layout / dtype of the output, staged so that global stores are whole 16-byte vectors
  const int OH = 2 * H, OW = 2 * W;
  const bool do_tanh = (flags & 1) != 0, unit = (flags & 2) != 0, u8 = (flags & 4) != 0, bgr = (flags & 8) != 0;
  float* stf = reinterpret_cast<float*>(smem);
  uint8_t* stb = smem;
  // this lane's two accumulator columns n = 2q, 2q+1 are fixed: (px, c) and the bias are per-lane constants
  int npx[2], nc[2];
  float nb[2];
  bool nvalid[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int n = 2 * q + e;
    nvalid[e] = n < 2 * Cimg;
    npx[e] = n / Cimg;
    nc[e] = n - npx[e] * Cimg;
    nb[e] = (bias != nullptr && nvalid[e]) ? __ldg(bias + nc[e]) : 0.0f;
  }
  // mode: 0 fp32 NCHW, 1 fp32 NHWC unit range, 2 uint8 NHWC -- one specialised staging loop each
  auto stage_out = [&](auto mode_tag) {
    constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int yl = 2 * warp + (m >> 1), xl0 = (m & 1) * 16;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (!nvalid[e & 1]) continue;
          const int px = npx[e & 1], c = nc[e & 1];
          const int Yl = 2 * yl + t, Xl = 2 * (xl0 + g + (e >> 1) * 8) + px;
          float v = acc[m][t][e] + nb[e & 1];
          if (do_tanh) v = tanh_fast(v);
          if (MODE == 2) {
            const float u = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), 255.0f);
            stb[(Yl * 64 + Xl) * Cimg + (bgr ? Cimg - 1 - c : c)] =
                static_cast<uint8_t>(__float2uint_rz(fminf(fmaxf(u, 0.0f), 255.0f)));
          } else if (MODE == 1) {
            stf[(Yl * 64 + Xl) * Cimg + c] = (v + 1.0f) * 0.5f;
          } else {
            stf[(c * (2 * kTH) + Yl) * 64 + Xl] = v;
          }
        }
      }
    }
  };
  if (u8) stage_out(std::integral_constant<int, 2>{});
  else if (unit) stage_out(std::integral_constant<int, 1>{});
  else stage_out(std::integral_constant<int, 0>{});
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
// GEMM view: M = pixels, N = 64, K = Cimg*16 with one k16 step per image channel (k = kh*4 + kw): the A fragment of a lane
// is four float2 loads of horizontally adjacent taps from the fp32 patch, the B fragments (the whole weight) live in
// registers for the CTA's life.
constexpr int kDownPatchBytes = kMaxCimg * kPR * kPitch * 4;          // 43520
constexpr int kDownStageBytes = (kThreads / 32) * 16 * 128;            // one 16-pixel x 64-channel bf16 tile per warp
constexpr int kDownSmem = kDownPatchBytes + kDownStageBytes;

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
img_conv_down_kernel(const float* __restrict__ x, const float* __restrict__ yimg, int mode,
                     const float* __restrict__ eps_dev, const float* __restrict__ mul_dev, const float* __restrict__ Wt,
                     const float* __restrict__ bias, float slope, const __nv_bfloat16* __restrict__ mask_src,
                     float mask_slope, __nv_bfloat16* __restrict__ out, int B, int Cimg, int S) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* patch = reinterpret_cast<float*>(smem);
  const int H = S / 2, W = S / 2;
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  const int ntiles = B * tiles_x * tiles_y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const float eps = eps_dev ? __ldg(eps_dev) : 0.0f;
  const float mul = mul_dev ? __ldg(mul_dev) : 1.0f;
  // B fragments: B[k = kh*4 + kw (channel c)][n = p]; lane holds k = 2q, 2q+1 (kh = q/2) and k + 8 (kh + 2), n = g
  uint32_t breg[kMaxCimg][8][2];
  const int kh0 = q >> 1, kw0 = (q & 1) * 2;
#pragma unroll
  for (int c = 0; c < kMaxCimg; ++c)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float w0 = 0.0f, w1 = 0.0f, w2 = 0.0f, w3 = 0.0f;
      if (c < Cimg) {
        const float* wp = Wt + ((static_cast<size_t>(nt * 8 + g) * Cimg + c) * 4 + kh0) * 4 + kw0;
        w0 = __ldg(wp); w1 = __ldg(wp + 1); w2 = __ldg(wp + 8); w3 = __ldg(wp + 9);
      }
      breg[c][nt][0] = bf2(w0, w1);
      breg[c][nt][1] = bf2(w2, w3);
    }
  float bv[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    bv[nt][0] = bias ? __ldg(bias + nt * 8 + 2 * q) : 0.0f;
    bv[nt][1] = bias ? __ldg(bias + nt * 8 + 2 * q + 1) : 0.0f;
  }
  uint8_t* stage = smem + kDownPatchBytes + warp * (16 * 128);
#pragma unroll 1
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
  const int y0 = ty * kTH, x0 = tx * kTW;
  __syncthreads();                                              // previous patch fully consumed
  load_img_patch(patch, x, yimg, mode, eps, mul, b, Cimg, S, y0, x0);
  __syncthreads();
#pragma unroll 1
  for (int m = 0; m < 4; ++m) {
    const int yl = 2 * warp + (m >> 1), xl0 = (m & 1) * 16;
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.0f;
#pragma unroll
    for (int c = 0; c < kMaxCimg; ++c) {
      if (c < Cimg) {
        // tap (kh, kw) of output pixel (yl, xl) sits at patch row 2*yl + kh, column gxl = 2*xl - 1 + kw
        const float* pr = patch + (c * kPR + 2 * yl + kh0) * kPitch + kPX - 1 + 2 * (xl0 + g) + kw0;
        const uint32_t a[4] = {bf2(pr[0], pr[1]),                                     // row g,     kh0
                               bf2(pr[16], pr[17]),                                   // row g + 8, kh0
                               bf2(pr[2 * kPitch], pr[2 * kPitch + 1]),               // row g,     kh0 + 2
                               bf2(pr[2 * kPitch + 16], pr[2 * kPitch + 17])};        // row g + 8, kh0 + 2
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) mma16816(acc[nt], a, breg[c][nt][0], breg[c][nt][1]);
      }
    }
    // epilogue: bias + LeakyReLU, bf16, swizzled per-warp staging (16 pixels x 128 B), then 16-byte global stores
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v0 = acc[nt][2 * h] + bv[nt][0], v1 = acc[nt][2 * h + 1] + bv[nt][1];
        v0 = v0 > 0.0f ? v0 : v0 * slope;
        v1 = v1 > 0.0f ? v1 : v1 * slope;
        const int row = g + 8 * h;
        *reinterpret_cast<uint32_t*>(stage + row * 128 + ((nt ^ (row & 7)) << 4) + q * 4) = bf2(v0, v1);
      }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int vi = it * 32 + lane, row = vi >> 3, ch = vi & 7;
      const int gy = y0 + yl, gx = x0 + xl0 + row;
      if (gy < H && gx < W) {
        uint4 v = *reinterpret_cast<const uint4*>(stage + row * 128 + ((ch ^ (row & 7)) << 4));
        const size_t o = ((static_cast<size_t>(b) * H + gy) * W + gx) * kC + ch * 8;
        if (mask_src != nullptr) {          // LeakyReLU backward mask from a stored activation: out *= (m > 0 ? 1 : slope)
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
    __syncwarp();
  }
  }
}

// =================================================================================================== wgrad (K3)
// part[cta][p][n] = sum over the CTA's tiles and pixels of act[b, y, x, p] * img'[b, c, 2y-1+kh, 2x-1+kw],
// n = (c*2 + kh/2)*8 + (kh%2)*4 + kw; column 2*Cimg*8 holds sum act (the bias gradient of the 64-channel side).
// GEMM view: M = p (64), N = taps, K = pixels.  A (p x pixel) comes transposed out of the activation tile with
// ldmatrix.trans; B (pixel x tap) is gathered from the fp32 patch.  Warp w: pixel rows {w>>1, (w>>1)+2, ...} of the
// tile, n tiles (w&1)*4 .. +3.  Fixed order everywhere: bit-reproducible.
constexpr int kWgActBytes = kTH * kTW * 128;                   // 32768
constexpr int kWgSmem = kWgActBytes + kDownPatchBytes;
constexpr int kWgN = 64;                                       // padded number of output columns

__global__ void __launch_bounds__(kThreads, kCtasPerSm)
img_conv_wgrad_kernel(const __nv_bfloat16* __restrict__ act, const float* __restrict__ x,
                      const float* __restrict__ yimg, int mode, const float* __restrict__ eps_dev,
                      const float* __restrict__ mul_dev, float* __restrict__ part, int B, int Cimg, int S) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* atile = smem;
  float* patch = reinterpret_cast<float*>(smem + kWgActBytes);
  const int H = S / 2, W = S / 2;
  const int tiles_x = (W + kTW - 1) / kTW, tiles_y = (H + kTH - 1) / kTH;
  const int ntiles = B * tiles_x * tiles_y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int nh = warp & 1, kg = warp >> 1;
  const float eps = eps_dev ? __ldg(eps_dev) : 0.0f;
  const float mul = mul_dev ? __ldg(mul_dev) : 1.0f;
  const int bias_nt = 2 * Cimg;                                 // n tile whose column 0 is the all-ones tap
  float acc[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[m][j][e] = 0.0f;
  const uint32_t atile_u32 = smem_u32(atile);
  // ldmatrix.trans addressing: lanes 0-7 / 8-15 -> pixels 0-7, channel chunk 2m / 2m+1; lanes 16-31 -> pixels 8-15
  const int lpx = (lane & 7) + (lane >> 4) * 8, lch = (lane >> 3) & 1;
  const uint32_t one_bf2 = bf2(1.0f, 1.0f);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int y0 = ty * kTH, x0 = tx * kTW;
    __syncthreads();                                            // previous tile fully consumed
    load_act_tile(atile, act, b, H, W, y0, x0, kTH, kTW);
    load_img_patch(patch, x, yimg, mode, eps, mul, b, Cimg, S, y0, x0);
    cp_commit_wait_all();
    __syncthreads();
#pragma unroll 1
    for (int yl = kg; yl < kTH; yl += kThreads / 64) {
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
            const float* pr = patch + (c * kPR + 2 * yl + kh) * kPitch + kPX - 1 + 2 * (xl0 + 2 * q) + kw;
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
  }
  // cross-warp reduction over the four pixel groups (fixed order), then one [64][64] partial per CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(smem);                  // [2 kg][64 p][64 n] = 32 KiB (the activation tile)
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
// One block per output row p: thread (sub, n) sums every 4th partial of column n (256-byte coalesced reads across n), the
// four sub-sums are combined in a fixed order -- the ~600 partials of 16 KiB are read in a few microseconds instead of by
// one serial strided loop per output element.
__global__ void __launch_bounds__(256) img_wgrad_finish_kernel(const float* __restrict__ part, int nparts, int Cimg,
                                                               float* __restrict__ dW, float acc,
                                                               float* __restrict__ dbias, float acc_b) {
  __shared__ float sm[4][kWgN];
  const int p = blockIdx.x, n = threadIdx.x & (kWgN - 1), sub = threadIdx.x >> 6;
  float s0 = 0.0f, s1 = 0.0f;
  int r = sub;
  for (; r + 4 < nparts; r += 8) {
    s0 += part[(static_cast<size_t>(r) * 64 + p) * kWgN + n];
    s1 += part[(static_cast<size_t>(r + 4) * 64 + p) * kWgN + n];
  }
  if (r < nparts) s0 += part[(static_cast<size_t>(r) * 64 + p) * kWgN + n];
  sm[sub][n] = s0 + s1;
  __syncthreads();
  if (sub != 0) return;
  const float s = (sm[0][n] + sm[1][n]) + (sm[2][n] + sm[3][n]);
  const int nt = n >> 3, e = n & 7;
  if (nt < 2 * Cimg) {
    const int c = nt >> 1, kh = (nt & 1) * 2 + (e >> 2), kw = e & 3;
    float* o = dW + (static_cast<size_t>(p) * Cimg + c) * 16 + kh * 4 + kw;
    *o = (acc != 0.0f ? acc * *o : 0.0f) + s;
  } else if (nt == 2 * Cimg && e == 0 && dbias != nullptr) {
    dbias[p] = (acc_b != 0.0f ? acc_b * dbias[p] : 0.0f) + s;
  }
}

int ensure_img_attrs() {
  static bool done = false;
  if (!done) {
    RG_CUDA(cudaFuncSetAttribute(img_conv_up_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem));
    RG_CUDA(cudaFuncSetAttribute(img_conv_up_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem));
    RG_CUDA(cudaFuncSetAttribute(img_conv_up_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem));
    RG_CUDA(cudaFuncSetAttribute(img_conv_up_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUpSmem));
    RG_CUDA(cudaFuncSetAttribute(img_conv_down_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDownSmem));
    RG_CUDA(cudaFuncSetAttribute(img_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
    done = true;
  }
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
  if (int rc = ensure_img_attrs()) return rc;
  const int grid = B * ceil_div(H, kTH) * ceil_div(Wd, kTW);
  const __nv_bfloat16* lop = static_cast<const __nv_bfloat16*>(lo);
  const uint2* wf = static_cast<const uint2*>(W);
  cudaStream_t s = static_cast<cudaStream_t>(st);
  switch (Cimg) {
    case 1: img_conv_up_kernel<1><<<grid, kThreads, kUpSmem, s>>>(lop, wf, bias, out, B, H, Wd, flags, bn_scale, bn_shift, bn_slope); break;
    case 2: img_conv_up_kernel<2><<<grid, kThreads, kUpSmem, s>>>(lop, wf, bias, out, B, H, Wd, flags, bn_scale, bn_shift, bn_slope); break;
    case 3: img_conv_up_kernel<3><<<grid, kThreads, kUpSmem, s>>>(lop, wf, bias, out, B, H, Wd, flags, bn_scale, bn_shift, bn_slope); break;
    default: img_conv_up_kernel<4><<<grid, kThreads, kUpSmem, s>>>(lop, wf, bias, out, B, H, Wd, flags, bn_scale, bn_shift, bn_slope); break;
  }
  RG_LAUNCH_CHECK("rg_img_conv_up");
  return 0;
}

int rg_img_conv_down(const float* x, const float* y, int mode, const float* eps_dev, const float* mul_dev,
                     const float* W, const float* bias, float slope, const void* mask_src, float mask_slope, int B,
                     int Cimg, int S, int Cp, void* out, rg_stream_t st) {
  RG_CHECK_ARG(x && W && out && B > 0 && Cp == kC && Cimg >= 1 && Cimg <= kMaxCimg && is_pow2(S) && S >= 16 &&
                   mode >= 0 && mode <= 2 && (mode == 0 || y != nullptr) && (mode != 1 || eps_dev != nullptr),
               "rg_img_conv_down: need 64 output channels, 1..4 image channels, a power-of-two side >= 16, mode 0..2 "
               "(Cp=%d Cimg=%d S=%d mode=%d)", Cp, Cimg, S, mode);
  if (int rc = ensure_img_attrs()) return rc;
  const int grid = std::min(B * ceil_div(S / 2, kTH) * ceil_div(S / 2, kTW), kCtasPerSm * num_sms());
  img_conv_down_kernel<<<grid, kThreads, kDownSmem, static_cast<cudaStream_t>(st)>>>(
      x, y, mode, eps_dev, mul_dev, W, bias, slope, static_cast<const __nv_bfloat16*>(mask_src), mask_slope,
      static_cast<__nv_bfloat16*>(out), B, Cimg, S);
  RG_LAUNCH_CHECK("rg_img_conv_down");
  return 0;
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
  if (int rc = ensure_img_attrs()) return rc;
  const int ntiles = B * ceil_div(S / 2, kTH) * ceil_div(S / 2, kTW);
  const int grid = std::min(ntiles, kCtasPerSm * num_sms());
  if (ws_bytes < static_cast<size_t>(grid) * 64 * kWgN * sizeof(float)) {
    set_error("rg_img_conv_wgrad: workspace too small (need %zu bytes)", static_cast<size_t>(grid) * 64 * kWgN * 4);
    return RG_EWORKSPACE;
  }
  cudaStream_t s = static_cast<cudaStream_t>(st);
  img_conv_wgrad_kernel<<<grid, kThreads, kWgSmem, s>>>(static_cast<const __nv_bfloat16*>(act), x, y, mode, eps_dev,
                                                        mul_dev, static_cast<float*>(ws), B, Cimg, S);
  RG_LAUNCH_CHECK("rg_img_conv_wgrad");
  img_wgrad_finish_kernel<<<64, 256, 0, s>>>(static_cast<const float*>(ws), grid, Cimg, dW, acc, dbias, acc_bias);
  RG_LAUNCH_CHECK("rg_img_conv_wgrad(finish)");
  return 0;
}

}  // extern "C"
