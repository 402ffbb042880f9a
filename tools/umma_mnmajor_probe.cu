// umma_mnmajor_probe.cu — hardware probe (not part of the product): MN-major (K-strided) SWIZZLE_128B operands for tcgen05.mma,
// as needed by the convolution weight-gradient GEMM  dW[co,ci] = sum_pixels dY[p,co] * X[p,ci]  where both operands are
// channel-contiguous in NHWC memory (K = pixels is the strided index).
//   A: global [K=64][M=128] bf16 (M contiguous), loaded as two TMA boxes {64 m, 64 k} -> smem blocks of 64 rows x 128 B
//   B: global [K=64][N=64]  bf16 (N contiguous), one box.
//   D[m][n] = sum_k A[k][m] * B[k][n]   (M=128, N=64, K=64 = 4 MMAs of K=16)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_mnmajor_probe tools/umma_mnmajor_probe.cu -I hr-viton_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "hrv_ptx.cuh"
using namespace hrv;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                uint32_t lbo, uint32_t sbo, uint32_t kstep, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar = base, bar2 = base + 8, slot = base + 16;
  const uint32_t sa = base + 1024, sb = sa + 2 * 8192;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 3 * 8192);
    tma_load_2d(sa, &tmA, bar, 0, 0);          // m 0..63
    tma_load_2d(sa + 8192, &tmA, bar, 64, 0);  // m 64..127
    tma_load_2d(sb, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);  // both operands MN-major
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t da = make_smem_desc(sa + kk * kstep, sbo, 2, lbo);
      uint64_t db = make_smem_desc(sb + kk * kstep, sbo, 2, lbo);
      umma_f16(tmem, da, db, idesc, kk != 0);
    }
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  for (int col = 0; col < 64; col += 16) {
    uint32_t v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + col, v);
    tmem_wait_ld();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + col + i] = __uint_as_float(v[i]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
int main() {
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  const int K = 64, M = 128, N = 64;
  std::vector<uint16_t> hA(K * M), hB(K * N);
  std::vector<float> fA(K * M), fB(K * N), ref(M * N, 0.f), out(M * N);
  for (int k = 0; k < K; ++k) for (int m = 0; m < M; ++m) { fA[k * M + m] = (float)(((k * 7 + m * 3) % 5) - 2); hA[k * M + m] = f2bf(fA[k * M + m]); }
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) { fB[k * N + n] = (float)(((k * 5 + n * 11) % 7) - 3); hB[k * N + n] = f2bf(fB[k * N + n]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += fA[k * M + m] * fB[k * N + n]; ref[m * N + n] = s; }
  uint16_t *dA, *dB; float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  { cuuint64_t d[2] = {(cuuint64_t)M, (cuuint64_t)K}; cuuint64_t s[1] = {(cuuint64_t)M * 2}; cuuint32_t b[2] = {64, 64}; cuuint32_t e[2] = {1, 1};
    enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  { cuuint64_t d[2] = {(cuuint64_t)N, (cuuint64_t)K}; cuuint64_t s[1] = {(cuuint64_t)N * 2}; cuuint32_t b[2] = {64, 64}; cuuint32_t e[2] = {1, 1};
    enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct V { uint32_t lbo, sbo, kstep; const char* what; } vars[] = {
      {8192, 1024, 2048, "LBO=M-block stride(8192) SBO=K-atom stride(1024) kstep=16 rows"},
      {1024, 8192, 2048, "swapped: LBO=1024 SBO=8192"},
      {8192, 1024, 1024, "kstep = 8 rows (wrong on purpose)"},
      {16, 1024, 2048, "LBO=16 (ignored?)"}};
  for (auto v : vars) {
    cudaMemset(dOut, 0, M * N * 4);
    probe<<<1, 128, 64 * 1024>>>(tmA, tmB, v.lbo, v.sbo, v.kstep, dOut);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.what, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dOut, M * N * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_lo = 0; for (int i = 0; i < M * N; ++i) { if (out[i] != ref[i]) { ++bad; if (i / N < 64) ++bad_lo; } }
    printf("%-70s mismatches %5d / %d (of which rows<64: %d)  D[0][0..3]= %g %g %g %g  ref %g %g %g %g\n", v.what, bad, M * N, bad_lo, out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
