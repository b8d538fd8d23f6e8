// FP32 math throughput on sm_100a: scalar FFMA / FADD against the packed f32x2 forms (FFMA2 / FADD2), and f64 DADD / F2F.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096, ILP = 8;
__global__ void k_ffma(float* out, float a, float b) {
  float x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
  float s = 0; for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[ILP], aa, bb;
  float2 t = make_float2(a, a), u = make_float2(b, b);
  aa = *reinterpret_cast<unsigned long long*>(&t); bb = *reinterpret_cast<unsigned long long*>(&u);
  for (int i = 0; i < ILP; i++) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
  for (int it = 0; it < ITER; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
  float s = 0; for (int i = 0; i < ILP; i++) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd(float* out, float a) {
  float x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = __fadd_rn(x[i], a);
  float s = 0; for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float* out, float a) {
  unsigned long long x[ILP], aa;
  float2 t = make_float2(a, a);
  aa = *reinterpret_cast<unsigned long long*>(&t);
  for (int i = 0; i < ILP; i++) { float2 v = make_float2(threadIdx.x + i, threadIdx.x - i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
  for (int it = 0; it < ITER; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(aa));
  float s = 0; for (int i = 0; i < ILP; i++) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dadd(float* out, double a) {
  double x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = __dadd_rn(x[i], a);
  double s = 0; for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
__global__ void k_f2f(float* out, double a) {  // the vfield chain's step: f32(f64(x) + v)
  float x[ILP];
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITER; it++)
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = (float)__dadd_rn((double)x[i], a);
  float s = 0; for (int i = 0; i < ILP; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class K, class... A> static float run(K k, A... a) {
  float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<148 * 8, 256>>>(out, a...); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<<<148 * 8, 256>>>(out, a...); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); cudaFree(out); return ms;
}
int main() {
  const double ops = 148.0 * 8 * 256 * ITER * ILP;  // thread-instructions
  float t;
  t = run(k_ffma, 1.0001f, 0.5f);  printf("FFMA   %.3f ms  %.2f T thread-instr/s  %.1f TFLOP/s\n", t, ops / t / 1e9, 2 * ops / t / 1e9);
  t = run(k_ffma2, 1.0001f, 0.5f); printf("FFMA2  %.3f ms  %.2f T thread-instr/s  %.1f TFLOP/s\n", t, ops / t / 1e9, 4 * ops / t / 1e9);
  t = run(k_fadd, 0.5f);           printf("FADD   %.3f ms  %.2f T thread-instr/s\n", t, ops / t / 1e9);
  t = run(k_fadd2, 0.5f);          printf("FADD2  %.3f ms  %.2f T thread-instr/s (x2 adds)\n", t, ops / t / 1e9);
  t = run(k_dadd, 0.5);            printf("DADD   %.3f ms  %.2f T thread-instr/s\n", t, ops / t / 1e9);
  t = run(k_f2f, 0.5);             printf("F2F+DADD+F2F %.3f ms  %.2f T chain-steps/s\n", t, ops / t / 1e9);
  return 0;
}
