// Dependent-issue latencies on one warp (clock64 around a chain of n dependent instructions).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
  __syncthreads();
  double x = a + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  double y = x;
  for (int i = 0; i < n; ++i) y = y * a;
  long long t2 = clock64();
  double z = y;
  for (int i = 0; i < n; ++i) z = z + b;
  long long t3 = clock64();
  // two independent chains
  double p = z, q = z + 1;
  for (int i = 0; i < n; ++i) { p = fma(p, a, b); q = fma(q, a, b); }
  long long t4 = clock64();
  double r0 = p, r1 = q, r2 = p + 1, r3 = q + 1;
  for (int i = 0; i < n; ++i) { r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b); }
  long long t5 = clock64();
  // LDS pointer chase
  int idx = threadIdx.x & 31;
  volatile double* vs = sm;
  double acc = r0 + r1 + r2 + r3;
  for (int i = 0; i < n; ++i) { double v = vs[idx]; idx = (int)v & 31; }
  long long t6 = clock64();
  float f = (float)acc;
  for (int i = 0; i < n; ++i) f = fmaf(f, 1.0001f, 0.5f);
  long long t7 = clock64();
  double w = acc;
  for (int i = 0; i < n; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w)); w = r; }
  long long t8 = clock64();
  if (threadIdx.x == 0) { cyc[0]=t1-t0; cyc[1]=t2-t1; cyc[2]=t3-t2; cyc[3]=t4-t3; cyc[4]=t5-t4; cyc[5]=t6-t5; cyc[6]=t7-t6; cyc[7]=t8-t7; }
  out[threadIdx.x] = acc + idx + f + w;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 64);
  const int n = 4096;
  for (int threads : {1, 32}) {
    k<<<1, threads>>>(out, cyc, 0.0, 0.0, n); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, cyc, 0.0, 0.0, n); cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    const char* nm[8] = {"DFMA dep", "DMUL dep", "DADD dep", "DFMA x2 indep (per pair)", "DFMA x4 indep (per quad)", "LDS.64 chase", "FFMA dep", "MUFU.RCP64H dep"};
    for (int i = 0; i < 8; ++i) printf("threads %2d  %-28s %.2f cycles\n", threads, nm[i], (double)h[i] / n);
  }
  return 0;
}
