// umma_rate_probe.cu (v2) — hardware probe (not part of the product): sustained issue/execution rate of tcgen05.mma kind::f16
// (K = 16) as a function of the tile shape (M = 128, N = 16..256) and the swizzle width, operands resident in shared memory.
//
// v1 of this probe (round 1) measured its own scalar loop (divergent thread-0 issue, descriptors rebuilt per MMA, a runtime modulo
// per iteration): its "153 cycles per MMA whatever N" was software, as the round-1 review showed.  This version issues from a
// CONVERGED warp through elect.sync, keeps all descriptors in registers (precomputed), and executes fully unrolled chains of 64
// MMAs per loop iteration inside ONE asm block (no per-MMA predicate set-up), alternating between `nacc` accumulators.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate_probe tools/umma_rate_probe.cu -I hr-viton_b200/csrc
#include <cstdio>
#include <cstdlib>
#include "hrv_ptx.cuh"
using namespace hrv;

#define MMA1(D, A, B) "tcgen05.mma.cta_group::1.kind::f16 [" D "], " A ", " B ", %10, p;\n\t"
// 8 MMAs: the four K=16 steps of one 64-wide K block on accumulator %0, then the same on accumulator %1
#define MMA8 MMA1("%0", "%2", "%6") MMA1("%0", "%3", "%7") MMA1("%0", "%4", "%8") MMA1("%0", "%5", "%9") \
             MMA1("%1", "%2", "%6") MMA1("%1", "%3", "%7") MMA1("%1", "%4", "%8") MMA1("%1", "%5", "%9")
#define MMA64 MMA8 MMA8 MMA8 MMA8 MMA8 MMA8 MMA8 MMA8

__global__ void __launch_bounds__(128, 1) probe(int M, int N, int layout, int kstep, int sbo, int nacc, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t bar = base, slot = base + 16;
  const uint32_t ops = base + 1024;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (ops - raw))[i] = 0;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  if (warp == 0) {  // whole warp, converged; one elected lane issues
    const uint32_t idesc = make_idesc_bf16(M, N);
    const uint32_t sa = ops, sb = ops + 16384;
    uint64_t da[4], db[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      da[k] = make_smem_desc(sa + k * kstep, sbo, layout);
      db[k] = make_smem_desc(sb + k * kstep, sbo, layout);
    }
    const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? 256 : 0);
    __syncwarp();
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t" MMA64 "}"
                     ::"r"(d0), "r"(d1), "l"(da[0]), "l"(da[1]), "l"(da[2]), "l"(da[3]), "l"(db[0]), "l"(db[1]), "l"(db[2]), "l"(db[3]), "r"(idesc)
                     : "memory");
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 256;  // x 64 MMAs
  struct { int layout, kstep, sbo; const char* name; } sw[] = {{2, 32, 1024, "SW128"}, {4, 32, 512, "SW64"}, {6, 32, 256, "SW32"}};
  const int Ns[] = {16, 32, 64, 96, 128, 160, 192, 224, 256};
  printf("%-28s %10s %12s %8s\n", "variant", "cyc/MMA", "ideal(cyc)", "pipe%");
  for (int grid : {1, 148}) {
    for (auto& s : sw) {
      for (int N : Ns) {
        for (int nacc = 1; nacc <= 2; ++nacc) {
          if ((nacc == 1) != (N == 256 || N == 128)) { if (nacc == 1) continue; }
          probe<<<grid, 128, 100 * 1024>>>(128, N, s.layout, s.kstep, s.sbo, nacc, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("N=%d %s: %s\n", N, s.name, cudaGetErrorString(e)); return 1; }
          long long c;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          const double per = (double)c / (iters * 64.0);
          const double ideal = 128.0 * N * 16 / 4096.0;  // 4096 bf16 MAC/clk/SM dense = 8192 FLOP/clk/SM
          printf("M128 N%-3d %-6s nacc=%d grid=%-3d %10.1f %12.1f %7.1f%%\n", N, s.name, nacc, grid, per, ideal, 100.0 * ideal / per);
        }
      }
    }
  }
  return 0;
}
