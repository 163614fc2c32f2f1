// Microbenchmark: dependent-issue latency and throughput of DFMA on one SM sub-partition.
// usage: ./dfma   (prints cycles per DFMA for ILP = 1..8 chains and 1..8 warps per SMSP)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000, threads = 128 * warps_per_smsp;   // 4 SMSPs
  k<ILP><<<1, threads>>>(out, cyc, iters, 0.999, 1e-3);
  k<ILP><<<1, threads>>>(out, cyc, iters, 0.999, 1e-3);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per_warp_instr = (double)iters * 8 * ILP;
  printf("ILP %d warps/SMSP %d: %.2f cycles per DFMA per warp, %.2f cycles per warp-DFMA per SMSP\n", ILP, warps_per_smsp,
         h / per_warp_instr, h / (per_warp_instr * warps_per_smsp));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4, 5, 6, 8}) { run<1>(w); run<2>(w); run<4>(w); run<8>(w); }
  return 0;
}
