// aux_kernels.cu — the HBM-bound kernels of the hot path (NHWC bf16, 16-byte vector accesses, 8 channels/thread):
// InstanceNorm statistics (with the SPADE noise, virtual up-sampling and concat), InstanceNorm apply, layout
// converters with nearest resampling, space-to-depth, 3x3/s2 average pool, bilinear x2 (+add), and the fused
// appearance-flow warp.  Citations: see include/hrviton_sm100.h.
#include "hrv_host.h"
#include "hrv_ptx.cuh"

namespace hrv {

struct View {
  const void* ptr;
  int n, h, w, c, pitch, dtype;
};
static View mk(const hrv_tensor* t) {
  View v;
  if (t) { v.ptr = t->ptr; v.n = t->n; v.h = t->h; v.w = t->w; v.c = t->c; v.pitch = t->pitch; v.dtype = t->dtype; }
  else { memset(&v, 0, sizeof(v)); }
  return v;
}

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16_lo(u.x); f[1] = bf16_hi(u.x); f[2] = bf16_lo(u.y); f[3] = bf16_hi(u.y);
  f[4] = bf16_lo(u.z); f[5] = bf16_hi(u.z); f[6] = bf16_lo(u.w); f[7] = bf16_hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// ------------------------------------------------------------------------------------------------ IN statistics
// grid (chunks, N); thread = (pixel lane pl, channel group g of 8).  Each thread accumulates fp32 partial sums over its
// pixels (4 independent 16-byte loads in flight), partials meet in shared memory (one slot per thread, no atomics),
// then one fp64 global atomic per channel per block.
struct StatsSrc {
  const __nv_bfloat16* p0;
  const __nv_bfloat16* p1;
  const float* noise;
  long long pitch0, pitch1;
  int W, W0, sh, from0;
};
__device__ __forceinline__ void stats_load(const StatsSrc& S, long long p, uint4& v, float& nz) {
  const int y = (int)(p / S.W), x = (int)(p - (long long)y * S.W);
  const __nv_bfloat16* src = S.from0 ? S.p0 + ((long long)(y >> S.sh) * S.W0 + (x >> S.sh)) * S.pitch0 : S.p1 + p * S.pitch1;
  v = ldg16(src);
  nz = S.noise ? __ldg(S.noise + p) : 0.f;
}
__global__ void __launch_bounds__(256) instnorm_stats_kernel(View x0, int x0_shift, View x1, int H, int W, int G, int PL,
                                                            int chunk, const float* __restrict__ noise,
                                                            const float* __restrict__ ns, double* __restrict__ acc) {
  extern __shared__ float shf[];  // [PL][C][2]
  const int n = blockIdx.y;
  const int C = G * 8;
  const int g = threadIdx.x % G;
  const int pl = threadIdx.x / G;
  if (pl < PL) {
    const int c0 = g * 8;
    const long long HW = (long long)H * W;
    const long long p_begin = (long long)blockIdx.x * chunk;
    long long p_end = p_begin + chunk;
    if (p_end > HW) p_end = HW;
    float s[8], q[8], nsv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; nsv[i] = ns ? __ldg(ns + c0 + i) : 0.f; }
    StatsSrc S;
    S.from0 = c0 < x0.c;
    S.sh = x0_shift; S.W = W; S.W0 = W >> x0_shift;
    S.pitch0 = x0.pitch; S.pitch1 = x1.pitch;
    S.p0 = reinterpret_cast<const __nv_bfloat16*>(x0.ptr) + (long long)n * (H >> x0_shift) * S.W0 * x0.pitch + c0;
    S.p1 = reinterpret_cast<const __nv_bfloat16*>(x1.ptr) + (long long)n * HW * x1.pitch + (c0 - x0.c);
    S.noise = noise ? noise + (long long)n * HW : nullptr;
    long long p = p_begin + pl;
    for (; p + 3LL * PL < p_end; p += 4LL * PL) {
      uint4 v[4];
      float nz[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) stats_load(S, p + (long long)u * PL, v[u], nz[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = fmaf(nz[u], nsv[i], f[i]);
          s[i] += t;
          q[i] = fmaf(t, t, q[i]);
        }
      }
    }
    for (; p < p_end; p += PL) {
      uint4 v;
      float nz, f[8];
      stats_load(S, p, v, nz);
      unpack8(v, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = fmaf(nz, nsv[i], f[i]);
        s[i] += t;
        q[i] = fmaf(t, t, q[i]);
      }
    }
    float* dst = shf + ((long long)pl * C + c0) * 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dst[2 * i] = s[i]; dst[2 * i + 1] = q[i]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double S1 = 0.0, S2 = 0.0;
    for (int l = 0; l < PL; ++l) {
      S1 += (double)shf[((long long)l * C + c) * 2];
      S2 += (double)shf[((long long)l * C + c) * 2 + 1];
    }
    atomicAdd(&acc[((long long)n * C + c) * 2], S1);
    atomicAdd(&acc[((long long)n * C + c) * 2 + 1], S2);
  }
}

// ------------------------------------------------------------------------------------------------ IN statistics, fused form
// Statistics of up to TWO SPADE norms that share one input (norm_s and norm_0 of a block both normalise x = cat(up2(x0), x1), each with
// its own noise draw and noise scale, network_generator.py:101-113,160-170) from ONE pass over the SOURCE tensors:
//   sum_P (x_c[P] + nz_j[P] ns_jc)   = S1_c + ns_jc * NZ1_j               S1_c = sum x_c,   NZ1_j = sum nz_j
//   sum_P (x_c[P] + nz_j[P] ns_jc)^2 = S2_c + 2 ns_jc X_jc + ns_jc^2 NZ2_j   S2_c = sum x_c^2, X_jc = sum x_c nz_j, NZ2_j = sum nz_j^2
// x0 is read at ITS OWN resolution: with x0_shift = 1 every source pixel stands for its 2x2 children, so S1/S2 count it four times and
// X_jc pairs it with the sum of the children's noise — 4x fewer bytes than walking the virtual up-sampled tensor, and the second
// norm costs no second pass.  Segment 0 of the grid (blockIdx.z) covers x0, segment 1 covers x1.  acc: [N][C][4] doubles
// {S1, S2, X_0, X_1}; nzacc: [N][2][2] doubles {NZ1_j, NZ2_j}.
__global__ void __launch_bounds__(256) instnorm_stats2_kernel(View xs, int shift, int c_off, int C, int W_full, int G, int PL, int chunk,
                                                             const float* __restrict__ nz0, const float* __restrict__ nz1,
                                                             double* __restrict__ acc, double* __restrict__ nzacc, int do_noise_sums) {
  extern __shared__ float shf[];  // [PL][G*8][4]
  __shared__ float nzred[8][4];
  const int n = blockIdx.y;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int Cs = G * 8;
  const long long HWs = (long long)xs.h * xs.w;           // source pixels
  const long long HWf = HWs << (2 * shift);               // full-resolution pixels
  float s[8], q[8], x0a[8], x1a[8], nsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; x0a[i] = 0.f; x1a[i] = 0.f; }
  if (pl < PL) {
    const long long p_begin = (long long)blockIdx.x * chunk;
    long long p_end = p_begin + chunk;
    if (p_end > HWs) p_end = HWs;
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(xs.ptr) + (long long)n * HWs * xs.pitch + g * 8;
    const float* n0 = nz0 ? nz0 + (long long)n * HWf : nullptr;
    const float* n1 = nz1 ? nz1 + (long long)n * HWf : nullptr;
    for (long long p = p_begin + pl; p < p_end; p += PL) {
      const uint4 v = ldg16(base + p * xs.pitch);
      float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;  // noise sum / sum of squares over the pixel's footprint, per norm
      if (shift) {
        const int y = (int)(p / xs.w), x = (int)(p - (long long)y * xs.w);
        const long long f = (long long)(2 * y) * W_full + 2 * x;
        if (n0) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(n0 + f)), u = __ldg(reinterpret_cast<const float2*>(n0 + f + W_full));
          a0 = (t.x + t.y) + (u.x + u.y);
          b0 = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(u.x, u.x, u.y * u.y)));
        }
        if (n1) {
          const float2 t = __ldg(reinterpret_cast<const float2*>(n1 + f)), u = __ldg(reinterpret_cast<const float2*>(n1 + f + W_full));
          a1 = (t.x + t.y) + (u.x + u.y);
          b1 = fmaf(t.x, t.x, fmaf(t.y, t.y, fmaf(u.x, u.x, u.y * u.y)));
        }
      } else {
        if (n0) { a0 = __ldg(n0 + p); b0 = a0 * a0; }
        if (n1) { a1 = __ldg(n1 + p); b1 = a1 * a1; }
      }
      float f8[8];
      unpack8(v, f8);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s[i] += f8[i];
        q[i] = fmaf(f8[i], f8[i], q[i]);
        x0a[i] = fmaf(f8[i], a0, x0a[i]);
        x1a[i] = fmaf(f8[i], a1, x1a[i]);
      }
      if (g == 0) { nsum[0] += a0; nsum[1] += b0; nsum[2] += a1; nsum[3] += b1; }
    }
    float* dst = shf + ((long long)pl * Cs + g * 8) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dst[4 * i] = s[i]; dst[4 * i + 1] = q[i]; dst[4 * i + 2] = x0a[i]; dst[4 * i + 3] = x1a[i]; }
  }
  __syncthreads();
  const double mult = (double)(1 << (2 * shift));  // every source pixel stands for 4^shift full-resolution pixels
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int l = 0; l < PL; ++l) {
      const float* e = shf + ((long long)l * Cs + c) * 4;
      t[0] += (double)e[0]; t[1] += (double)e[1]; t[2] += (double)e[2]; t[3] += (double)e[3];
    }
    double* a = acc + ((long long)n * C + c_off + c) * 4;
    atomicAdd(a, t[0] * mult);
    atomicAdd(a + 1, t[1] * mult);
    atomicAdd(a + 2, t[2]);
    atomicAdd(a + 3, t[3]);
  }
  if (do_noise_sums) {  // NZ1/NZ2 of both norms: only the g == 0 threads hold contributions; warp reduce, then one atomic per block
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = nsum[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) nzred[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0.0;
      for (int wv = 0; wv < 8; ++wv) t += (double)nzred[wv][threadIdx.x];
      atomicAdd(nzacc + (long long)n * 4 + threadIdx.x, t);
    }
  }
}

__global__ void instnorm_finalize2_kernel(const double* __restrict__ acc, const double* __restrict__ nzacc, int N, int C, double inv_hw, float eps,
                                          const float* __restrict__ ns0, const float* __restrict__ ns1, float* __restrict__ mean0,
                                          float* __restrict__ rstd0, float* __restrict__ mean1, float* __restrict__ rstd1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  const double S1 = acc[4 * (long long)i], S2 = acc[4 * (long long)i + 1];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    float* mean = j ? mean1 : mean0;
    float* rstd = j ? rstd1 : rstd0;
    if (!mean) continue;
    const float* nsp = j ? ns1 : ns0;
    const double ns = nsp ? (double)nsp[c] : 0.0;
    const double X = acc[4 * (long long)i + 2 + j], NZ1 = nzacc[n * 4 + 2 * j], NZ2 = nzacc[n * 4 + 2 * j + 1];
    const double m = (S1 + ns * NZ1) * inv_hw;
    double var = (S2 + 2.0 * ns * X + ns * ns * NZ2) * inv_hw - m * m;
    if (var < 0.0) var = 0.0;
    mean[i] = (float)m;
    rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

__global__ void instnorm_finalize_kernel(const double* __restrict__ acc, int NC, double inv_hw, float eps,
                                         float* __restrict__ mean, float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NC) return;
  const double m = acc[2 * i] * inv_hw;
  double var = acc[2 * i + 1] * inv_hw - m * m;
  if (var < 0.0) var = 0.0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

__global__ void instnorm_apply_kernel(View x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                      const float* __restrict__ gamma, const float* __restrict__ beta, View res, int act, View y,
                                      long long total, int G) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int n = (int)(pix / ((long long)x.h * x.w));
  float f[8], r[8];
  unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(x.ptr) + pix * x.pitch + g * 8), f);
  if (res.ptr) unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(res.ptr) + pix * res.pitch + g * 8), r);
  const float* mp = mean + (long long)n * x.c + g * 8;
  const float* rp = rstd + (long long)n * x.c + g * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    if (c < x.c) {
      float v = (f[i] - __ldg(mp + i)) * __ldg(rp + i);
      if (gamma) v = fmaf(v, __ldg(gamma + c), beta ? __ldg(beta + c) : 0.f);
      if (res.ptr) v += r[i];
      f[i] = apply_act(v, act);
    } else {
      f[i] = 0.f;
    }
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(y.ptr)) + pix * y.pitch + g * 8) = pack8(f);
}

// ------------------------------------------------------------------------------------------------ layout converters
// thread = (dst pixel, 8-channel group); src fp32 NCHW with nearest resampling.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int C, int SH, int SW, View dst, int G, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int x = (int)(pix % dst.w);
  const int y = (int)((pix / dst.w) % dst.h);
  const int n = (int)(pix / ((long long)dst.w * dst.h));
  // src = floor(dst * in / out): exact in integers
  const int sy = (int)(((long long)y * SH) / dst.h);
  const int sx = (int)(((long long)x * SW) / dst.w);
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = g * 8 + i;
    f[i] = c < C ? __ldg(src + (((long long)n * C + c) * SH + sy) * SW + sx) : 0.f;
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(dst.ptr)) + pix * dst.pitch + g * 8) = pack8(f);
}

// thread = (pixel); loops channels; writes are coalesced per plane.
__global__ void nhwc_to_nchw_kernel(View src, float* __restrict__ dst, long long npix) {
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const long long hw = (long long)src.h * src.w;
  const int n = (int)(pix / hw);
  const long long rem = pix - (long long)n * hw;
  float* d = dst + (long long)n * src.c * hw + rem;
  if (src.dtype == 0) {
    const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(src.ptr) + pix * src.pitch;
    for (int c = 0; c < src.c; ++c) d[(long long)c * hw] = __bfloat162float(s[c]);
  } else {
    const float* s = reinterpret_cast<const float*>(src.ptr) + pix * src.pitch;
    for (int c = 0; c < src.c; ++c) d[(long long)c * hw] = s[c];
  }
}

// thread = (dst pixel, sub-pixel 0..3, 8-channel group of src)
__global__ void space_to_depth_kernel(View src, View dst, int G, int Gd, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gd = (int)(idx % Gd);  // dst channel group
  const long long pix = idx / Gd;
  const int X = (int)(pix % dst.w);
  const int Y = (int)((pix / dst.w) % dst.h);
  const int n = (int)(pix / ((long long)dst.w * dst.h));
  uint4 v = make_uint4(0, 0, 0, 0);
  const int sub = gd / G, g = gd % G;
  if (sub < 4) {
    const int sy = 2 * Y + (sub >> 1), sx = 2 * X + (sub & 1);
    if (sy < src.h && sx < src.w)
      v = ldg16(reinterpret_cast<const __nv_bfloat16*>(src.ptr) + (((long long)n * src.h + sy) * src.w + sx) * src.pitch + g * 8);
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(dst.ptr)) + pix * dst.pitch + gd * 8) = v;
}

__global__ void avgpool3s2_kernel(View src, View dst, int G, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int X = (int)(pix % dst.w);
  const int Y = (int)((pix / dst.w) % dst.h);
  const int n = (int)(pix / ((long long)dst.w * dst.h));
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cnt = 0;
  for (int dy = -1; dy <= 1; ++dy) {
    const int y = 2 * Y + dy;
    if (y < 0 || y >= src.h) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = 2 * X + dx;
      if (x < 0 || x >= src.w) continue;
      float f[8];
      unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(src.ptr) + (((long long)n * src.h + y) * src.w + x) * src.pitch + g * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += f[i];
      ++cnt;
    }
  }
  const float inv = 1.f / (float)cnt;
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] *= inv;
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(dst.ptr)) + pix * dst.pitch + g * 8) = pack8(s);
}

// F.interpolate(bilinear, x2, align_corners=False): s = max((d+0.5)/2-0.5, 0); i0=floor(s); i1=min(i0+1,n-1)
__device__ __forceinline__ void up2_taps(int d, int n, int& i0, int& i1, float& w0, float& w1) {
  float s = __fsub_rn(__fmul_rn(__fadd_rn((float)d, 0.5f), 0.5f), 0.5f);
  s = fmaxf(s, 0.f);
  const float fl = floorf(s);
  i0 = (int)fl;
  i1 = min(i0 + 1, n - 1);
  w1 = __fsub_rn(s, fl);
  w0 = __fsub_rn(1.f, w1);
}

__global__ void bilinear_up2_add_kernel(View a, View b, View dst, int G, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int x = (int)(pix % dst.w);
  const int y = (int)((pix / dst.w) % dst.h);
  const int n = (int)(pix / ((long long)dst.w * dst.h));
  int x0, x1, y0, y1;
  float wx0, wx1, wy0, wy1;
  up2_taps(x, a.w, x0, x1, wx0, wx1);
  up2_taps(y, a.h, y0, y1, wy0, wy1);
  const __nv_bfloat16* ap = reinterpret_cast<const __nv_bfloat16*>(a.ptr) + (long long)n * a.h * a.w * a.pitch + g * 8;
  float f00[8], f01[8], f10[8], f11[8], o[8];
  unpack8(ldg16(ap + ((long long)y0 * a.w + x0) * a.pitch), f00);
  unpack8(ldg16(ap + ((long long)y0 * a.w + x1) * a.pitch), f01);
  unpack8(ldg16(ap + ((long long)y1 * a.w + x0) * a.pitch), f10);
  unpack8(ldg16(ap + ((long long)y1 * a.w + x1) * a.pitch), f11);
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = wy0 * (wx0 * f00[i] + wx1 * f01[i]) + wy1 * (wx0 * f10[i] + wx1 * f11[i]);
  if (b.ptr) {
    float fb[8];
    unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(b.ptr) + pix * b.pitch + g * 8), fb);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] += fb[i];
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(dst.ptr)) + pix * dst.pitch + g * 8) = pack8(o);
}

// ------------------------------------------------------------------------------------------------ flow warp
// thread = (dst pixel, 8-channel group).  The coordinate chain uses explicitly rounded fp32 operations (no FMA
// contraction) in the exact order of oracle/hrviton_oracle.py:np_flow_warp_coords so that the gather indices
// are bit-exact.
__global__ void flow_warp_kernel(const float* __restrict__ flow_lo, const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                                 View src, View dst, float* __restrict__ flow_up, int* __restrict__ idx_out, float sx, float sy,
                                 int G, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int W = dst.w, H = dst.h;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  const int hl = H >> 1, wl = W >> 1;
  int xa, xb, ya, yb;
  float wxa, wxb, wya, wyb;
  up2_taps(x, wl, xa, xb, wxa, wxb);
  up2_taps(y, hl, ya, yb, wya, wyb);
  const float2* fl = reinterpret_cast<const float2*>(flow_lo) + (long long)n * hl * wl;
  const float2 f00 = __ldg(fl + (long long)ya * wl + xa), f01 = __ldg(fl + (long long)ya * wl + xb);
  const float2 f10 = __ldg(fl + (long long)yb * wl + xa), f11 = __ldg(fl + (long long)yb * wl + xb);
  // lerp along x inside each source row, then along y (torch's CPU kernel order)
  const float ux0 = __fadd_rn(__fmul_rn(f00.x, wxa), __fmul_rn(f01.x, wxb));
  const float ux1 = __fadd_rn(__fmul_rn(f10.x, wxa), __fmul_rn(f11.x, wxb));
  const float uy0 = __fadd_rn(__fmul_rn(f00.y, wxa), __fmul_rn(f01.y, wxb));
  const float uy1 = __fadd_rn(__fmul_rn(f10.y, wxa), __fmul_rn(f11.y, wxb));
  const float fx = __fadd_rn(__fmul_rn(ux0, wya), __fmul_rn(ux1, wyb));
  const float fy = __fadd_rn(__fmul_rn(uy0, wya), __fmul_rn(uy1, wyb));
  const float gx = __fadd_rn(__fdiv_rn(fx, sx), __ldg(lin_x + x));
  const float gy = __fadd_rn(__fdiv_rn(fy, sy), __ldg(lin_y + y));
  float ix = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)src.w), 1.f), 2.f);
  float iy = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)src.h), 1.f), 2.f);
  ix = fminf(fmaxf(ix, 0.f), (float)(src.w - 1));
  iy = fminf(fmaxf(iy, 0.f), (float)(src.h - 1));
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const float tx = __fsub_rn(ix, fx0), ty = __fsub_rn(iy, fy0);
  if (g == 0) {
    if (flow_up) reinterpret_cast<float2*>(flow_up)[pix] = make_float2(fx, fy);
    if (idx_out) reinterpret_cast<int2*>(idx_out)[pix] = make_int2(x0, y0);
  }
  if (g * 8 >= ((src.c + 7) & ~7)) return;
  const int x1 = min(x0 + 1, src.w - 1), y1 = min(y0 + 1, src.h - 1);
  const float wx1 = (x0 + 1 <= src.w - 1) ? tx : 0.f, wy1 = (y0 + 1 <= src.h - 1) ? ty : 0.f;
  const float wx0 = 1.f - tx, wy0 = 1.f - ty;
  const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
  float o[8];
  const long long ib = (long long)n * src.h * src.w;
  if (src.dtype == 0) {
    const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(src.ptr) + g * 8;
    float a[8], b[8], c[8], d[8];
    unpack8(ldg16(sp + (ib + (long long)y0 * src.w + x0) * src.pitch), a);
    unpack8(ldg16(sp + (ib + (long long)y0 * src.w + x1) * src.pitch), b);
    unpack8(ldg16(sp + (ib + (long long)y1 * src.w + x0) * src.pitch), c);
    unpack8(ldg16(sp + (ib + (long long)y1 * src.w + x1) * src.pitch), d);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = a[i] * w00 + b[i] * w01 + c[i] * w10 + d[i] * w11;
  } else {
    const float* sp = reinterpret_cast<const float*>(src.ptr) + g * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (g * 8 + i < src.c) {
        o[i] = __ldg(sp + (ib + (long long)y0 * src.w + x0) * src.pitch + i) * w00 + __ldg(sp + (ib + (long long)y0 * src.w + x1) * src.pitch + i) * w01 +
               __ldg(sp + (ib + (long long)y1 * src.w + x0) * src.pitch + i) * w10 + __ldg(sp + (ib + (long long)y1 * src.w + x1) * src.pitch + i) * w11;
      } else {
        o[i] = 0.f;
      }
    }
  }
  if (dst.dtype == 0) {
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(dst.ptr)) + pix * dst.pitch + g * 8) = pack8(o);
  } else {
    float* dp = reinterpret_cast<float*>(const_cast<void*>(dst.ptr)) + pix * dst.pitch + g * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (g * 8 + i < dst.c) dp[i] = o[i];
  }
}

static int check_bf16_vec(const hrv_tensor* t, const char* what) {
  if (!t || !t->ptr) return set_error(HRV_EINVAL, "%s: null tensor", what);
  if (t->dtype != HRV_BF16) return set_error(HRV_EINVAL, "%s: must be bf16", what);
  if (((uintptr_t)t->ptr & 15) || (t->pitch % 8)) return set_error(HRV_EINVAL, "%s: needs 16-byte alignment and pitch %% 8 == 0", what);
  return 0;
}
static int launch_ok(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(HRV_ECUDA, "%s launch: %s", what, cudaGetErrorString(e));
  return HRV_OK;
}
static inline unsigned blocks_for(long long total, int bs) { return (unsigned)((total + bs - 1) / bs); }

}  // namespace hrv

using namespace hrv;

extern "C" int hrv_instnorm_stats(const hrv_tensor* x0, int32_t x0_shift, const hrv_tensor* x1, int32_t h, int32_t w,
                                  const float* noise, const float* noise_scale, float eps, float* mean, float* rstd,
                                  void* workspace, size_t workspace_bytes, hrv_stream stream) {
  int rc = check_bf16_vec(x0, "instnorm_stats x0");
  if (rc) return rc;
  const bool has1 = x1 && x1->ptr;
  if (has1 && (rc = check_bf16_vec(x1, "instnorm_stats x1"))) return rc;
  const int C = x0->c + (has1 ? x1->c : 0);
  if ((x0->c % 8) || (C % 8)) return set_error(HRV_EINVAL, "instnorm_stats: channel counts must be multiples of 8");
  if (C / 8 > 256) return set_error(HRV_EUNSUPPORTED, "instnorm_stats: more than 2048 channels");
  if ((x0->h << x0_shift) != h || (x0->w << x0_shift) != w) return set_error(HRV_EINVAL, "instnorm_stats: x0 extent mismatch");
  if (has1 && (x1->h != h || x1->w != w || x1->n != x0->n)) return set_error(HRV_EINVAL, "instnorm_stats: x1 extent mismatch");
  const int N = x0->n;
  const size_t need = (size_t)N * C * 2 * sizeof(double);
  if (!workspace || workspace_bytes < need) return set_error(HRV_EINVAL, "instnorm_stats: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(workspace, 0, need, st);
  const int G = C / 8;
  const int PL = 256 / G;
  const long long HW = (long long)h * w;
  long long target_blocks = (long long)sm_count() * 8 / (N > 0 ? N : 1);
  if (target_blocks < 1) target_blocks = 1;
  long long chunk = (HW + target_blocks - 1) / target_blocks;
  const long long min_chunk = (long long)PL * 16;
  if (chunk < min_chunk) chunk = min_chunk;
  const unsigned gx = (unsigned)((HW + chunk - 1) / chunk);
  instnorm_stats_kernel<<<dim3(gx, N), 256, (size_t)PL * C * 2 * sizeof(float), st>>>(mk(x0), x0_shift, mk(has1 ? x1 : nullptr), h, w, G, PL,
                                                                                 (int)chunk, noise, noise_scale, (double*)workspace);
  if ((rc = launch_ok("instnorm_stats"))) return rc;
  instnorm_finalize_kernel<<<blocks_for(N * C, 256), 256, 0, st>>>((const double*)workspace, N * C, 1.0 / (double)HW, eps, mean, rstd);
  return launch_ok("instnorm_finalize");
}

extern "C" int hrv_instnorm_stats2(const hrv_tensor* x0, int32_t x0_shift, const hrv_tensor* x1, int32_t h, int32_t w, const float* noise0,
                                   const float* noise_scale0, const float* noise1, const float* noise_scale1, float eps, float* mean0,
                                   float* rstd0, float* mean1, float* rstd1, void* workspace, size_t workspace_bytes, hrv_stream stream) {
  int rc = check_bf16_vec(x0, "instnorm_stats2 x0");
  if (rc) return rc;
  const bool has1 = x1 && x1->ptr;
  if (has1 && (rc = check_bf16_vec(x1, "instnorm_stats2 x1"))) return rc;
  const int C = x0->c + (has1 ? x1->c : 0);
  if ((x0->c % 8) || (C % 8)) return set_error(HRV_EINVAL, "instnorm_stats2: channel counts must be multiples of 8");
  if (x0->c / 8 > 256 || (has1 && x1->c / 8 > 256)) return set_error(HRV_EUNSUPPORTED, "instnorm_stats2: more than 2048 channels per source");
  if (x0_shift < 0 || x0_shift > 1) return set_error(HRV_EINVAL, "instnorm_stats2: x0_shift must be 0 or 1");
  if ((x0->h << x0_shift) != h || (x0->w << x0_shift) != w) return set_error(HRV_EINVAL, "instnorm_stats2: x0 extent mismatch");
  if (x0_shift && (w & 1)) return set_error(HRV_EINVAL, "instnorm_stats2: up-sampled width must be even");
  if (has1 && (x1->h != h || x1->w != w || x1->n != x0->n)) return set_error(HRV_EINVAL, "instnorm_stats2: x1 extent mismatch");
  if (!mean0 || !rstd0 || ((mean1 == nullptr) != (rstd1 == nullptr))) return set_error(HRV_EINVAL, "instnorm_stats2: outputs");
  if ((noise0 && !noise_scale0) || (noise1 && !noise_scale1)) return set_error(HRV_EINVAL, "instnorm_stats2: noise without noise_scale");
  const int N = x0->n;
  const size_t need = ((size_t)N * C * 4 + (size_t)N * 4) * sizeof(double);
  if (!workspace || workspace_bytes < need) return set_error(HRV_EINVAL, "instnorm_stats2: workspace too small (%zu < %zu)", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(workspace, 0, need, st);
  double* acc = (double*)workspace;
  double* nzacc = acc + (size_t)N * C * 4;
  for (int seg = 0; seg < (has1 ? 2 : 1); ++seg) {
    const hrv_tensor* xs = seg ? x1 : x0;
    const int shift = seg ? 0 : x0_shift;
    const int G = xs->c / 8, PL = 256 / G;
    const long long HWs = (long long)xs->h * xs->w;
    long long target_blocks = (long long)sm_count() * 8 / (N > 0 ? N : 1);
    if (target_blocks < 1) target_blocks = 1;
    long long chunk = (HWs + target_blocks - 1) / target_blocks;
    const long long min_chunk = (long long)PL * 16;
    if (chunk < min_chunk) chunk = min_chunk;
    const unsigned gx = (unsigned)((HWs + chunk - 1) / chunk);
    // the noise-only sums are taken once, by the segment that walks the full-resolution grid exactly once per pixel: x0's walk covers
    // every full-resolution pixel through its footprint, so segment 0 always does it
    instnorm_stats2_kernel<<<dim3(gx, N), 256, (size_t)PL * G * 8 * 4 * sizeof(float), st>>>(mk(xs), shift, seg ? x0->c : 0, C, w, G, PL, (int)chunk,
                                                                                             noise0, noise1, acc, nzacc, seg == 0);
    if ((rc = launch_ok("instnorm_stats2"))) return rc;
  }
  instnorm_finalize2_kernel<<<blocks_for(N * C, 256), 256, 0, st>>>(acc, nzacc, N, C, 1.0 / ((double)h * w), eps, noise0 ? noise_scale0 : nullptr,
                                                                   noise1 ? noise_scale1 : nullptr, mean0, rstd0, mean1, rstd1);
  return launch_ok("instnorm_finalize2");
}

extern "C" int hrv_norm_apply_affine(const hrv_tensor* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                                     const hrv_tensor* res, int32_t act, const hrv_tensor* y, hrv_stream stream) {
  int rc;
  if ((rc = check_bf16_vec(x, "norm_apply x")) || (rc = check_bf16_vec(y, "norm_apply y"))) return rc;
  const bool hr = res && res->ptr;
  if (hr && (rc = check_bf16_vec(res, "norm_apply res"))) return rc;
  if (!mean || !rstd) return set_error(HRV_EINVAL, "norm_apply: mean/rstd required");
  const int G = (x->c + 7) / 8;
  const long long total = (long long)x->n * x->h * x->w * G;
  instnorm_apply_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mk(x), mean, rstd, gamma, beta, mk(hr ? res : nullptr), act,
                                                                                 mk(y), total, G);
  return launch_ok("norm_apply");
}

extern "C" int hrv_instnorm_apply(const hrv_tensor* x, const float* mean, const float* rstd, int32_t act, const hrv_tensor* y,
                                  hrv_stream stream) {
  return hrv_norm_apply_affine(x, mean, rstd, nullptr, nullptr, nullptr, act, y, stream);
}

extern "C" int hrv_nchw_to_nhwc(const float* src, int32_t c, int32_t src_h, int32_t src_w, const hrv_tensor* dst, hrv_stream stream) {
  int rc = check_bf16_vec(dst, "nchw_to_nhwc dst");
  if (rc) return rc;
  const int G = (dst->c + 7) / 8;
  if (G * 8 > dst->pitch) return set_error(HRV_EINVAL, "nchw_to_nhwc: dst pitch too small for padded channels");
  const long long total = (long long)dst->n * dst->h * dst->w * G;
  nchw_to_nhwc_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, c, src_h, src_w, mk(dst), G, total);
  return launch_ok("nchw_to_nhwc");
}

extern "C" int hrv_nhwc_to_nchw(const hrv_tensor* src, float* dst, hrv_stream stream) {
  if (!src || !src->ptr || !dst) return set_error(HRV_EINVAL, "nhwc_to_nchw: null");
  const long long npix = (long long)src->n * src->h * src->w;
  nhwc_to_nchw_kernel<<<blocks_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(mk(src), dst, npix);
  return launch_ok("nhwc_to_nchw");
}

extern "C" int hrv_space_to_depth(const hrv_tensor* src, const hrv_tensor* dst, hrv_stream stream) {
  int rc;
  if ((rc = check_bf16_vec(src, "s2d src")) || (rc = check_bf16_vec(dst, "s2d dst"))) return rc;
  const int G = (src->c + 7) / 8;
  if (dst->h != (src->h + 1) / 2 || dst->w != (src->w + 1) / 2 || dst->c < 4 * G * 8 || (dst->c % 8))
    return set_error(HRV_EINVAL, "s2d: dst must be ceil(h/2) x ceil(w/2) x (>= 4*roundup8(c))");
  const int Gd = dst->c / 8;
  const long long total = (long long)dst->n * dst->h * dst->w * Gd;
  space_to_depth_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mk(src), mk(dst), G, Gd, total);
  return launch_ok("space_to_depth");
}

extern "C" int hrv_avgpool3s2(const hrv_tensor* src, const hrv_tensor* dst, hrv_stream stream) {
  int rc;
  if ((rc = check_bf16_vec(src, "avgpool src")) || (rc = check_bf16_vec(dst, "avgpool dst"))) return rc;
  if (dst->h != (src->h - 1) / 2 + 1 || dst->w != (src->w - 1) / 2 + 1) return set_error(HRV_EINVAL, "avgpool: dst extent");
  const int G = (src->c + 7) / 8;
  const long long total = (long long)dst->n * dst->h * dst->w * G;
  avgpool3s2_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mk(src), mk(dst), G, total);
  return launch_ok("avgpool3s2");
}

extern "C" int hrv_bilinear_up2_add(const hrv_tensor* a, const hrv_tensor* b, const hrv_tensor* dst, hrv_stream stream) {
  int rc;
  if ((rc = check_bf16_vec(a, "up2 a")) || (rc = check_bf16_vec(dst, "up2 dst"))) return rc;
  const bool hb = b && b->ptr;
  if (hb && (rc = check_bf16_vec(b, "up2 b"))) return rc;
  if (dst->h != 2 * a->h || dst->w != 2 * a->w) return set_error(HRV_EINVAL, "up2: dst must be 2x");
  const int G = (a->c + 7) / 8;
  const long long total = (long long)dst->n * dst->h * dst->w * G;
  bilinear_up2_add_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mk(a), mk(hb ? b : nullptr), mk(dst), G, total);
  return launch_ok("bilinear_up2_add");
}

extern "C" int hrv_flow_warp(const float* flow_lo, const float* lin_x, const float* lin_y, const hrv_tensor* src, const hrv_tensor* dst,
                             float* flow_up, int32_t* idx_out, hrv_stream stream) {
  if (!flow_lo || !lin_x || !lin_y || !src || !dst || !src->ptr || !dst->ptr) return set_error(HRV_EINVAL, "flow_warp: null argument");
  if ((dst->h & 1) || (dst->w & 1)) return set_error(HRV_EINVAL, "flow_warp: output extent must be even");
  if (src->dtype == HRV_BF16 && (((uintptr_t)src->ptr & 15) || (src->pitch % 8))) return set_error(HRV_EINVAL, "flow_warp: src alignment");
  if (dst->dtype == HRV_BF16 && (((uintptr_t)dst->ptr & 15) || (dst->pitch % 8))) return set_error(HRV_EINVAL, "flow_warp: dst alignment");
  if (src->n != dst->n || src->c != dst->c) return set_error(HRV_EINVAL, "flow_warp: src/dst mismatch");
  const int G = (src->c + 7) / 8;
  const long long total = (long long)dst->n * dst->h * dst->w * G;
  // Python-float arithmetic of the reference ((iW/2 - 1.0)/2.0), rounded once to fp32 as torch does for a scalar divisor
  const float sx = (float)((dst->w / 2.0 - 1.0) / 2.0);
  const float sy = (float)((dst->h / 2.0 - 1.0) / 2.0);
  flow_warp_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(flow_lo, lin_x, lin_y, mk(src), mk(dst), flow_up, idx_out, sx, sy, G, total);
  return launch_ok("flow_warp");
}

// ------------------------------------------------------------------------------------------------ backward of the resampling ops
namespace hrv {

// Adjoint of F.interpolate(bilinear, x2, align_corners=False) along one axis: low-res index i receives from high-res
// d in {2i-1, 2i, 2i+1, 2i+2} with weights {0.25, 0.75, 0.75, 0.25}; the clamped border taps fold onto i (weights 1.0).
__device__ __forceinline__ void up2_adj(int i, int n, float (&wt)[4]) {
  wt[0] = i >= 1 ? 0.25f : 0.f;
  wt[1] = i == 0 ? 1.0f : 0.75f;
  wt[2] = i == n - 1 ? 1.0f : 0.75f;
  wt[3] = i <= n - 2 ? 0.25f : 0.f;
}

// thread = (low-res pixel, 8-channel group): da[n,y,x,:] = sum over the 4x4 high-res neighbourhood of wy*wx*dout
__global__ void bilinear_up2_bwd_kernel(View dout, View da, int G, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int x = (int)(pix % da.w);
  const int y = (int)((pix / da.w) % da.h);
  const int n = (int)(pix / ((long long)da.w * da.h));
  float wy[4], wx[4], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  up2_adj(y, da.h, wy);
  up2_adj(x, da.w, wx);
  const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(dout.ptr) + (long long)n * dout.h * dout.w * dout.pitch + g * 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int Y = 2 * y - 1 + j;
    if (wy[j] == 0.f || Y < 0 || Y >= dout.h) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int X = 2 * x - 1 + i;
      if (wx[i] == 0.f || X < 0 || X >= dout.w) continue;
      float f[8];
      unpack8(ldg16(base + ((long long)Y * dout.w + X) * dout.pitch), f);
      const float w = wy[j] * wx[i];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(w, f[k], acc[k]);
    }
  }
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(da.ptr)) + pix * da.pitch + g * 8) = pack8(acc);
}

// Backward of flow_warp_kernel.  thread = (dst pixel, 8-channel group): recomputes the sampling coordinates exactly as the
// forward, scatter-adds d_dst * w into d_src (fp32 accumulation buffer, atomics: the gather is irregular) and accumulates the
// analytic coordinate gradient (grid_sample border semantics: zero where the coordinate was clamped) into d_flow_up (fp32).
__global__ void flow_warp_bwd_kernel(const float* __restrict__ flow_lo, const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                                     View src, View ddst, float* __restrict__ dsrc32, float* __restrict__ dflow_up, float sx, float sy,
                                     int G, long long total) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % G);
  const long long pix = idx / G;
  const int W = ddst.w, H = ddst.h;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const int n = (int)(pix / ((long long)W * H));
  const int hl = H >> 1, wl = W >> 1;
  int xa, xb, ya, yb;
  float wxa, wxb, wya, wyb;
  up2_taps(x, wl, xa, xb, wxa, wxb);
  up2_taps(y, hl, ya, yb, wya, wyb);
  const float2* fl = reinterpret_cast<const float2*>(flow_lo) + (long long)n * hl * wl;
  const float2 f00 = __ldg(fl + (long long)ya * wl + xa), f01 = __ldg(fl + (long long)ya * wl + xb);
  const float2 f10 = __ldg(fl + (long long)yb * wl + xa), f11 = __ldg(fl + (long long)yb * wl + xb);
  const float fx = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(f00.x, wxa), __fmul_rn(f01.x, wxb)), wya), __fmul_rn(__fadd_rn(__fmul_rn(f10.x, wxa), __fmul_rn(f11.x, wxb)), wyb));
  const float fy = __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(f00.y, wxa), __fmul_rn(f01.y, wxb)), wya), __fmul_rn(__fadd_rn(__fmul_rn(f10.y, wxa), __fmul_rn(f11.y, wxb)), wyb));
  const float gx = __fadd_rn(__fdiv_rn(fx, sx), __ldg(lin_x + x));
  const float gy = __fadd_rn(__fdiv_rn(fy, sy), __ldg(lin_y + y));
  const float ixr = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gx, 1.f), (float)src.w), 1.f), 2.f);
  const float iyr = __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(gy, 1.f), (float)src.h), 1.f), 2.f);
  const float ix = fminf(fmaxf(ixr, 0.f), (float)(src.w - 1)), iy = fminf(fmaxf(iyr, 0.f), (float)(src.h - 1));
  const float mx = (ixr < 0.f || ixr > (float)(src.w - 1)) ? 0.f : 1.f;  // clip_coordinates_set_grad
  const float my = (iyr < 0.f || iyr > (float)(src.h - 1)) ? 0.f : 1.f;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const float tx = ix - fx0, ty = iy - fy0;
  const bool vx = x0 + 1 <= src.w - 1, vy = y0 + 1 <= src.h - 1;
  const int x1 = vx ? x0 + 1 : x0, y1 = vy ? y0 + 1 : y0;
  const float wx1 = vx ? tx : 0.f, wy1 = vy ? ty : 0.f, wx0 = 1.f - tx, wy0 = 1.f - ty;
  if (g * 8 >= ((src.c + 7) & ~7)) return;
  float d[8], a[8], b[8], c[8], e[8];
  unpack8(ldg16(reinterpret_cast<const __nv_bfloat16*>(ddst.ptr) + pix * ddst.pitch + g * 8), d);
  const long long ib = (long long)n * src.h * src.w;
  const __nv_bfloat16* sp = reinterpret_cast<const __nv_bfloat16*>(src.ptr) + g * 8;
  unpack8(ldg16(sp + (ib + (long long)y0 * src.w + x0) * src.pitch), a);
  unpack8(ldg16(sp + (ib + (long long)y0 * src.w + x1) * src.pitch), b);
  unpack8(ldg16(sp + (ib + (long long)y1 * src.w + x0) * src.pitch), c);
  unpack8(ldg16(sp + (ib + (long long)y1 * src.w + x1) * src.pitch), e);
  float dix = 0.f, diy = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    // value = a*wy0*wx0 + b*wy0*wx1 + c*wy1*wx0 + e*wy1*wx1 ; taps outside the image carry zero weight AND zero derivative
    const float bb = vx ? b[k] : 0.f, cc = vy ? c[k] : 0.f, ee = (vx && vy) ? e[k] : 0.f;
    dix += d[k] * ((bb - a[k]) * wy0 + (ee - cc) * (vy ? ty : 0.f));
    diy += d[k] * ((cc - a[k]) * wx0 + (ee - bb) * (vx ? tx : 0.f));
  }
  if (dflow_up) {
    // d ix / d gx = src.w/2 ; d gx / d flow_up.x = 1/sx
    atomicAdd(dflow_up + pix * 2, dix * mx * (0.5f * (float)src.w) / sx);
    atomicAdd(dflow_up + pix * 2 + 1, diy * my * (0.5f * (float)src.h) / sy);
  }
  if (dsrc32) {
    const int cs = (src.c + 7) & ~7;
    float* o00 = dsrc32 + (ib + (long long)y0 * src.w + x0) * cs + g * 8;
    float* o01 = dsrc32 + (ib + (long long)y0 * src.w + x1) * cs + g * 8;
    float* o10 = dsrc32 + (ib + (long long)y1 * src.w + x0) * cs + g * 8;
    float* o11 = dsrc32 + (ib + (long long)y1 * src.w + x1) * cs + g * 8;
    const float w00 = wy0 * wx0, w01 = wy0 * wx1, w10 = wy1 * wx0, w11 = wy1 * wx1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      atomicAdd(o00 + k, d[k] * w00);
      if (w01 != 0.f) atomicAdd(o01 + k, d[k] * w01);
      if (w10 != 0.f) atomicAdd(o10 + k, d[k] * w10);
      if (w11 != 0.f) atomicAdd(o11 + k, d[k] * w11);
    }
  }
}

// d_flow_lo[n,y,x,:] = adjoint of the x2 bilinear up-sampling applied to d_flow_up (2 fp32 channels)
__global__ void flow_up2_bwd_kernel(const float* __restrict__ dup, float* __restrict__ dlo, int N, int hl, int wl) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * hl * wl;
  if (idx >= total) return;
  const int x = (int)(idx % wl), y = (int)((idx / wl) % hl), n = (int)(idx / ((long long)wl * hl));
  float wy[4], wx[4], ax = 0.f, ay = 0.f;
  up2_adj(y, hl, wy);
  up2_adj(x, wl, wx);
  const int H = 2 * hl, W = 2 * wl;
  const float2* src = reinterpret_cast<const float2*>(dup) + (long long)n * H * W;
  for (int j = 0; j < 4; ++j) {
    const int Y = 2 * y - 1 + j;
    if (wy[j] == 0.f || Y < 0 || Y >= H) continue;
    for (int i = 0; i < 4; ++i) {
      const int X = 2 * x - 1 + i;
      if (wx[i] == 0.f || X < 0 || X >= W) continue;
      const float2 v = __ldg(src + (long long)Y * W + X);
      ax = fmaf(wy[j] * wx[i], v.x, ax);
      ay = fmaf(wy[j] * wx[i], v.y, ay);
    }
  }
  reinterpret_cast<float2*>(dlo)[idx] = make_float2(ax, ay);
}

}  // namespace hrv

extern "C" int hrv_bilinear_up2_bwd(const hrv_tensor* dout, const hrv_tensor* da, hrv_stream stream) {
  int rc;
  if ((rc = check_bf16_vec(dout, "up2_bwd dout")) || (rc = check_bf16_vec(da, "up2_bwd da"))) return rc;
  if (dout->h != 2 * da->h || dout->w != 2 * da->w) return set_error(HRV_EINVAL, "up2_bwd: dout must be 2x da");
  const int G = (da->c + 7) / 8;
  const long long total = (long long)da->n * da->h * da->w * G;
  bilinear_up2_bwd_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mk(dout), mk(da), G, total);
  return launch_ok("bilinear_up2_bwd");
}

extern "C" int hrv_flow_warp_bwd(const float* flow_lo, const float* lin_x, const float* lin_y, const hrv_tensor* src, const hrv_tensor* ddst,
                                 float* dsrc32, float* dflow_up, float* dflow_lo, hrv_stream stream) {
  int rc;
  if (!flow_lo || !lin_x || !lin_y) return set_error(HRV_EINVAL, "flow_warp_bwd: null argument");
  if ((rc = check_bf16_vec(src, "flow_warp_bwd src")) || (rc = check_bf16_vec(ddst, "flow_warp_bwd ddst"))) return rc;
  if ((ddst->h & 1) || (ddst->w & 1) || src->n != ddst->n || src->c != ddst->c) return set_error(HRV_EINVAL, "flow_warp_bwd: shape mismatch");
  if (dflow_lo && !dflow_up) return set_error(HRV_EINVAL, "flow_warp_bwd: dflow_lo needs the dflow_up scratch buffer");
  const int G = (src->c + 7) / 8;
  const long long total = (long long)ddst->n * ddst->h * ddst->w * G;
  const float sx = (float)((ddst->w / 2.0 - 1.0) / 2.0), sy = (float)((ddst->h / 2.0 - 1.0) / 2.0);
  cudaStream_t st = (cudaStream_t)stream;
  flow_warp_bwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(flow_lo, lin_x, lin_y, mk(src), mk(ddst), dsrc32, dflow_up, sx, sy, G, total);
  if ((rc = launch_ok("flow_warp_bwd"))) return rc;
  if (dflow_lo) {
    const long long tl = (long long)ddst->n * (ddst->h / 2) * (ddst->w / 2);
    flow_up2_bwd_kernel<<<blocks_for(tl, 256), 256, 0, st>>>(dflow_up, dflow_lo, ddst->n, ddst->h / 2, ddst->w / 2);
    return launch_ok("flow_up2_bwd");
  }
  return HRV_OK;
}

// ------------------------------------------------------------------------------------------------ weight packing
namespace hrv {
// One pass from the fp32 parameter layout (cout,cin,kh,kw) to the bf16 GEMM layout [taps][n_pad][cin_k] (zero padded), with
// optional 1/sigma scaling, (gamma,beta) row interleave of two parameters and the flip+transpose of the data-gradient convolution.
__global__ void pack_conv_weight_kernel(const float* __restrict__ w0, const float* __restrict__ w1, int cout, int cin, int kh, int kw,
                                        int transpose_flip, const float* __restrict__ inv_scale, __nv_bfloat16* __restrict__ dst,
                                        int n_pad, int cin_k, long long plane) {
  // thread = one (row n, K index c) of the GEMM operand; it walks the kh*kw taps, i.e. reads the taps of one (co,ci) pair from
  // CONTIGUOUS fp32 memory and writes one bf16 per tap plane (coalesced across the threads of a warp, which vary c).
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= plane) return;
  const int c = (int)(idx % cin_k);
  const int n = (int)(idx / cin_k);
  const int inter = w1 ? 2 : 1;
  const int taps = kh * kw;
  const float* src = nullptr;
  if (!transpose_flip) {
    // GEMM row n = output channel (interleaved pair index), K index c = input channel
    if (n < cout * inter && c < cin) src = ((w1 && (n & 1)) ? w1 : w0) + ((long long)(n / inter) * cin + c) * taps;
  } else {
    // data-gradient conv: GEMM row n = original INPUT channel, K index c = original OUTPUT channel (interleaved), taps mirrored
    if (n < cin && c < cout * inter) src = ((w1 && (c & 1)) ? w1 : w0) + ((long long)(c / inter) * cin + n) * taps;
  }
  const float sc = inv_scale ? __ldg(inv_scale) : 1.f;
  for (int tap = 0; tap < taps; ++tap) {
    const float v = src ? __ldg(src + (transpose_flip ? taps - 1 - tap : tap)) * sc : 0.f;
    dst[(long long)tap * plane + idx] = __float2bfloat16(v);
  }
}
}  // namespace hrv

extern "C" int hrv_pack_conv_weight(const float* w0, const float* w1, int32_t cout, int32_t cin, int32_t kh, int32_t kw,
                                    int32_t transpose_flip, const float* inv_scale, void* dst, int32_t n_pad, int32_t cin_k,
                                    hrv_stream stream) {
  if (!w0 || !dst || cout < 1 || cin < 1 || kh < 1 || kw < 1) return set_error(HRV_EINVAL, "pack_conv_weight: bad arguments");
  const int inter = w1 ? 2 : 1;
  const int rows = transpose_flip ? cin : cout * inter, cols = transpose_flip ? cout * inter : cin;
  if (n_pad < rows || cin_k < cols) return set_error(HRV_EINVAL, "pack_conv_weight: destination too small (%d x %d for %d x %d)", n_pad, cin_k, rows, cols);
  const long long plane = (long long)n_pad * cin_k;
  pack_conv_weight_kernel<<<blocks_for(plane, 256), 256, 0, (cudaStream_t)stream>>>(w0, w1, cout, cin, kh, kw, transpose_flip, inv_scale,
                                                                                   (__nv_bfloat16*)dst, n_pad, cin_k, plane);
  return launch_ok("pack_conv_weight");
}
