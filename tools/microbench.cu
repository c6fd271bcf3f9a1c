// Design-constant microbenchmarks for the compare kernels (B200, sm_100a):
// L2-resident and DRAM read bandwidth, FP64 FMA / divide rate, SHFL and
// LDS.64 rate.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void read_kernel(const double2 *__restrict__ p, size_t n2, int reps, double *sink) {
  double acc = 0;
  for (int r = 0; r < reps; ++r)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
      double2 v = p[i];
      acc += v.x + v.y;
    }
  if (acc == 1234.5) *sink = acc;
}
__global__ void read64_kernel(const double *__restrict__ p, size_t n, int reps, double *sink) {
  double acc = 0;
  for (int r = 0; r < reps; ++r)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += p[i];
  if (acc == 1234.5) *sink = acc;
}
template <int ILP>
__global__ void dfma_kernel(int iters, double *sink, double a, double b) {
  double v[ILP];
  for (int k = 0; k < ILP; ++k) v[k] = threadIdx.x + k;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = fma(v[k], a, b);
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += v[k];
  if (s == 1234.5) *sink = s;
}
template <int ILP>
__global__ void ddiv_kernel(int iters, double *sink, double a) {
  double v[ILP];
  for (int k = 0; k < ILP; ++k) v[k] = 1.0 + threadIdx.x + k;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = a / v[k] + 1.0;
  double s = 0;
  for (int k = 0; k < ILP; ++k) s += v[k];
  if (s == 1234.5) *sink = s;
}
__global__ void shfl_kernel(int iters, double *sink) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3;
  for (int i = 0; i < iters; ++i) {
    v0 = __shfl_down_sync(0xffffffffu, v0, 1); v1 = __shfl_down_sync(0xffffffffu, v1, 2);
    v2 = __shfl_down_sync(0xffffffffu, v2, 3); v3 = __shfl_down_sync(0xffffffffu, v3, 4);
  }
  if (v0 + v1 + v2 + v3 == 1234.5) *sink = v0;
}
__global__ void lds_kernel(int iters, double *sink) {
  __shared__ double s[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = i;
  __syncthreads();
  double a = 0;
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    a += s[idx & 2047] + s[(idx + 256) & 2047] + s[(idx + 512) & 2047] + s[(idx + 768) & 2047];
    idx += 1;
  }
  if (a == 1234.5) *sink = a;
}
__global__ void rcp_err_kernel(double *out) {
  // max relative error of rcp.approx.ftz.f64 raw / after one / after two Newton steps
  double e0 = 0, e1 = 0, e2 = 0;
  unsigned long long st = 88172645463325252ull + threadIdx.x * 7919ull + blockIdx.x * 104729ull;
  for (int i = 0; i < 20000; ++i) {
    st ^= st << 13; st ^= st >> 7; st ^= st << 17;
    const double m = 1.0 + (double)(st >> 11) * (1.0 / 9007199254740992.0);   // [1,2)
    const int ex = (int)((st >> 3) % 600) - 300;
    const double d = ldexp(m, ex);
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = 1.0 / d;
    e0 = fmax(e0, fabs(y - t) / t);
    double e = fma(-d, y, 1.0); y = fma(y, e, y);
    e1 = fmax(e1, fabs(y - t) / t);
    e = fma(-d, y, 1.0); y = fma(y, e, y);
    e2 = fmax(e2, fabs(y - t) / t);
  }
  atomicMax((unsigned long long *)&out[0], (unsigned long long)__double_as_longlong(e0));
  atomicMax((unsigned long long *)&out[1], (unsigned long long)__double_as_longlong(e1));
  atomicMax((unsigned long long *)&out[2], (unsigned long long)__double_as_longlong(e2));
}
template <typename F> float timeit(F f, int n = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  float best = 1e30f;
  for (int i = 0; i < n; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
  return best;
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %d MB, smem/block optin %zu\n", prop.name, sms, prop.l2CacheSize >> 20, prop.sharedMemPerBlockOptin);
  double *sink; CK(cudaMalloc(&sink, 8));
  size_t big = (size_t)2 << 30; double *buf; CK(cudaMalloc(&buf, big)); CK(cudaMemset(buf, 0, big));
  for (size_t mb : {8, 24, 48, 96}) {
    size_t bytes = mb << 20; int reps = 50;
    float ms = timeit([&] { read_kernel<<<sms * 8, 512>>>((double2 *)buf, bytes / 16, reps, sink); });
    printf("L2-resident read %3zu MB (LDG.128): %.0f GB/s\n", mb, bytes * (double)reps / ms / 1e6);
    ms = timeit([&] { read64_kernel<<<sms * 8, 512>>>(buf, bytes / 8, reps, sink); });
    printf("L2-resident read %3zu MB (LDG.64):  %.0f GB/s\n", mb, bytes * (double)reps / ms / 1e6);
  }
  { float ms = timeit([&] { read_kernel<<<sms * 8, 512>>>((double2 *)buf, big / 16, 1, sink); });
    printf("DRAM read 2 GB (LDG.128): %.0f GB/s\n", big / ms / 1e6); }
  { int iters = 4096; float ms = timeit([&] { dfma_kernel<8><<<sms * 4, 512>>>(iters, sink, 1.000001, 1e-9); });
    double flops = 2.0 * iters * 8 * (double)sms * 4 * 512; printf("FP64 FMA: %.2f TFLOP/s (%.1f DFMA/clk/SM at %.0f MHz nominal)\n", flops / ms / 1e9, flops / 2 / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1e3); }
  { int iters = 1024; float ms = timeit([&] { ddiv_kernel<4><<<sms * 4, 512>>>(iters, sink, 3.0); });
    double divs = (double)iters * 4 * sms * 4 * 512; printf("FP64 divide (+1 add): %.1f Gdiv/s = %.2f div/clk/SM nominal\n", divs / ms / 1e6, divs / (ms * 1e-3) / sms / (prop.clockRate * 1e3)); }
  { int iters = 4096; float ms = timeit([&] { shfl_kernel<<<sms * 4, 512>>>(iters, sink); });
    double sh = (double)iters * 8 * sms * 4 * 512; printf("SHFL.32: %.1f G lane-shuffles/s = %.1f lanes/clk/SM nominal\n", sh / ms / 1e6, sh / (ms * 1e-3) / sms / (prop.clockRate * 1e3)); }
  { int iters = 4096; float ms = timeit([&] { lds_kernel<<<sms * 4, 512>>>(iters, sink); });
    double by = (double)iters * 4 * 8 * sms * 4 * 512; printf("LDS.64: %.0f GB/s = %.1f B/clk/SM nominal\n", by / ms / 1e6, by / (ms * 1e-3) / sms / (prop.clockRate * 1e3)); }
  { double *errbuf; CK(cudaMalloc(&errbuf, 24)); CK(cudaMemset(errbuf, 0, 24)); rcp_err_kernel<<<64, 128>>>(errbuf); double h[3]; CK(cudaMemcpy(h, errbuf, 24, cudaMemcpyDeviceToHost));
    printf("rcp.approx.ftz.f64 max rel error: raw %.3e, +1 Newton %.3e, +2 Newton %.3e\n", h[0], h[1], h[2]); }
  return 0;
}
