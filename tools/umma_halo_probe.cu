// umma_halo_probe.cu — hardware probe (not part of the product): how does tcgen05.mma address a K-major SWIZZLE_128B
// operand whose descriptor start is NOT 1024-byte aligned / whose SBO is not a multiple of 1024?  Decides whether a
// 3x3 convolution can read its 9 taps as shifted views of ONE halo tile in shared memory.
//   A[r][k] in global: value encodes r (run 0) or k (run 1); B = identity (64x64) so D[m][n] = A_read(m, n).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_halo_probe tools/umma_halo_probe.cu -I hr-viton_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "hrv_ptx.cuh"
using namespace hrv;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                int r0, int sbo, int base_off, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar = base, bar2 = base + 8, slot = base + 16;
  const uint32_t sa = base + 1024, sb = sa + 256 * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(sa, &tmA, bar, 0, 0);
    tma_load_2d(sb, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 64);
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t da = make_smem_desc(sa + r0 * 128 + kk * 32, sbo, 2) | ((uint64_t)(base_off & 7) << 49);
      uint64_t db = make_smem_desc(sb + kk * 32, 1024, 2);
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
  const int R = 256;
  std::vector<uint16_t> hA(R * 64), hB(64 * 64, 0);
  for (int n = 0; n < 64; ++n) hB[n * 64 + n] = f2bf(1.f);
  uint16_t *dA, *dB; float* dOut;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dOut, 128 * 64 * 4);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  { cuuint64_t d[2] = {64, (cuuint64_t)R}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, (cuuint32_t)R}; cuuint32_t e[2] = {1, 1};
    enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  { cuuint64_t d[2] = {64, 64}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, 64}; cuuint32_t e[2] = {1, 1};
    enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct V { int r0, sbo, bo; } vars[] = {{0, 1024, 0}, {1, 1024, 0}, {1, 1024, 1}, {3, 1024, 0}, {3, 1024, 3}, {0, 1280, 0}, {1, 1280, 0}, {1, 1280, 1},
                                          {2, 2048, 0}, {2, 2048, 2}, {11, 1280, 0}, {11, 1280, 3}, {8, 1280, 0}, {16, 2048, 0}, {5, 2048, 0}, {5, 2048, 5}};
  std::vector<float> out(128 * 64), outk(128 * 64);
  for (auto v : vars) {
    for (int run = 0; run < 2; ++run) {
      for (int r = 0; r < R; ++r) for (int k = 0; k < 64; ++k) hA[r * 64 + k] = f2bf(run == 0 ? (float)r : (float)k);
      cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
      probe<<<1, 128, 64 * 1024>>>(tmA, tmB, v.r0, v.sbo, v.bo, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant r0=%d sbo=%d bo=%d: CUDA error %s\n", v.r0, v.sbo, v.bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(run == 0 ? out.data() : outk.data(), dOut, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    }
    int bad_row = 0, bad_k = 0;
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) {
      const int exp_r = v.r0 + (m / 8) * (v.sbo / 128) + (m % 8);
      if ((int)out[m * 64 + k] != exp_r) ++bad_row;
      if ((int)outk[m * 64 + k] != k) ++bad_k;
    }
    printf("r0=%2d sbo=%4d base_off=%d : wrong-row=%4d wrong-k=%4d  %s\n", v.r0, v.sbo, v.bo, bad_row, bad_k, (bad_row | bad_k) ? "MISMATCH" : "OK");
    if (bad_row | bad_k) {
      for (int m = 0; m < 12; ++m) { printf("   m=%2d rows:", m); for (int k = 0; k < 64; k += 8) printf(" %3d", (int)out[m * 64 + k]); printf("  k:"); for (int k = 0; k < 64; k += 8) printf(" %2d", (int)outk[m * 64 + k]); printf("\n"); }
    }
  }
  return 0;
}
