// umma_rate_probe.cu — hardware probe (not part of the product): issue rate of tcgen05.mma kind::f16 (K=16) as a function of the
// tile shape (M, N), the operand majorness and the swizzle width, with operands resident in shared memory (no TMA, zeros).
// Answers: what is the per-instruction floor that bounds the small-N convolutions, and would swapping the operand roles help?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate_probe tools/umma_rate_probe.cu -I hr-viton_b200/csrc
#include <cstdio>
#include <cstdlib>
#include "hrv_ptx.cuh"
using namespace hrv;

struct Variant {
  int M, N, a_mn, b_mn, layout;  // layout: 2=SW128 4=SW64 6=SW32
  int kstep_a, kstep_b;          // descriptor start advance per K=16 step (bytes)
  int sbo_a, sbo_b, lbo_a, lbo_b;
  int mmas_per_stage;            // K=16 steps per smem stage
  const char* what;
};

__global__ void __launch_bounds__(128, 1) probe(Variant v, int stages, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar = base, slot = base + 16;
  const uint32_t ops = base + 1024;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (ops - raw))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(v.M, v.N, v.a_mn, v.b_mn);
    const uint32_t a_bytes = 16384, b_bytes = 32768;  // per stage regions (A: 128 rows x 128 B; B: 256 rows x 128 B)
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const int s = it % stages;
      const uint32_t sa = ops + s * (a_bytes + b_bytes), sb = sa + a_bytes;
      for (int kk = 0; kk < v.mmas_per_stage; ++kk) {
        const uint64_t da = make_smem_desc(sa + kk * v.kstep_a, v.sbo_a, v.layout, v.lbo_a);
        const uint64_t db = make_smem_desc(sb + kk * v.kstep_b, v.sbo_b, v.layout, v.lbo_b);
        umma_f16(tmem + (it & 1) * 256, da, db, idesc, 1);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2048;
  Variant vars[] = {
      // K-major SW128 (bk64): 4 K-steps of 32 B inside one 128-B swizzle row; SBO = 8 rows * 128 B
      {128, 16, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N16  K-major SW128"},
      {128, 32, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N32  K-major SW128"},
      {128, 64, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N64  K-major SW128"},
      {128, 96, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N96  K-major SW128"},
      {128, 128, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N128 K-major SW128"},
      {128, 160, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N160 K-major SW128"},
      {128, 192, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N192 K-major SW128"},
      {128, 256, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M128 N256 K-major SW128"},
      {64, 32, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M64  N32  K-major SW128"},
      {64, 64, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M64  N64  K-major SW128"},
      {64, 128, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M64  N128 K-major SW128"},
      {64, 256, 0, 0, 2, 32, 32, 1024, 1024, 16, 16, 4, "M64  N256 K-major SW128"},
      // narrower swizzles (bk32 / bk16 stages)
      {128, 32, 0, 0, 4, 32, 32, 512, 512, 16, 16, 2, "M128 N32  K-major SW64"},
      {128, 128, 0, 0, 4, 32, 32, 512, 512, 16, 16, 2, "M128 N128 K-major SW64"},
      {128, 32, 0, 0, 6, 32, 32, 256, 256, 16, 16, 1, "M128 N32  K-major SW32"},
      {128, 128, 0, 0, 6, 32, 32, 256, 256, 16, 16, 1, "M128 N128 K-major SW32"},
      // MN-major SW128 (weight-gradient GEMM): 64-element blocks of 128 B, K advances by 16 rows = 2048 B
      {128, 64, 1, 1, 2, 2048, 2048, 1024, 1024, 8192, 8192, 4, "M128 N64  MN-major SW128"},
      {128, 128, 1, 1, 2, 2048, 2048, 1024, 1024, 8192, 8192, 4, "M128 N128 MN-major SW128"},
      {128, 256, 1, 1, 2, 2048, 2048, 1024, 1024, 8192, 8192, 4, "M128 N256 MN-major SW128"},
      {128, 128, 1, 0, 2, 2048, 32, 1024, 1024, 8192, 16, 4, "M128 N128 A MN-major, B K-major"},
      {128, 128, 0, 1, 2, 32, 2048, 1024, 1024, 16, 8192, 4, "M128 N128 A K-major, B MN-major"},
  };
  printf("%-36s %10s %12s %10s\n", "variant", "cyc/MMA", "ideal(cyc)", "pipe%%");
  for (const Variant& v : vars) {
    for (int stages = 1; stages <= 4; stages += 3) {
      probe<<<148, 128, 200 * 1024>>>(v, stages, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", v.what, cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per = (double)c / (iters * v.mmas_per_stage);
      const double ideal = (double)v.M * v.N * 16 / 4096.0;  // 4096 bf16 MAC/clk/SM (dense) = 8192 FLOP/clk/SM
      printf("%-36s %10.1f %12.1f %9.1f%%   (stages=%d)\n", v.what, per, ideal, 100.0 * ideal / per, stages);
    }
  }
  return 0;
}
