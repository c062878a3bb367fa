// Throughput of the legacy tensor path on this part: mma.sync.m16n8k8 TF32 (fp32 accumulate) and m8n8k4 FP64, as
// multiply-adds per clock per SM, for a given number of resident warps per SM (each warp keeps 8 independent accumulator
// chains).  Decides whether the mid-size Riccati products are worth moving from FFMA/DFMA to mma.sync.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_rate tools/mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f64(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int MODE>
__global__ void rate_kernel(float* out, int iters, long long* clocks) {
  const int lane = threadIdx.x & 31;
  float acc = 0.f;
  long long t0 = clock64();
  if (MODE == 0) {
    float c[8][4] = {};
    unsigned a[4] = {0x3f800000u + lane, 0x3f800000u, 0x3f900000u, 0x3fa00000u}, b[2] = {0x3f800000u, 0x3f800000u + lane};
    for (int i = 0; i < iters; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) mma_tf32(c[j], a, b);
    for (int j = 0; j < 8; j++) acc += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  } else if (MODE == 1) {
    double c[8][2] = {};
    double a = 1.0 + lane * 1e-3, b = 1.0 - lane * 1e-3;
    for (int i = 0; i < iters; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) mma_f64(c[j], a, b);
    for (int j = 0; j < 8; j++) acc += (float)(c[j][0] + c[j][1]);
  } else if (MODE == 3) {  // packed FP32: 16 independent float2 chains (FFMA2)
    float2 c[16];
    for (int j = 0; j < 16; j++) c[j] = make_float2(lane + j, lane - j);
    const float2 a = make_float2(1.0001f, 0.9999f), b = make_float2(0.5f, 0.25f);
    for (int i = 0; i < iters; i++)
#pragma unroll
      for (int j = 0; j < 16; j++) c[j] = __ffma2_rn(c[j], a, b);
    for (int j = 0; j < 16; j++) acc += c[j].x + c[j].y;
  } else {  // FFMA reference: 8 x 4 independent chains
    float c[32];
    for (int j = 0; j < 32; j++) c[j] = lane + j;
    float a = 1.0001f, b = 0.5f;
    for (int i = 0; i < iters; i++)
#pragma unroll
      for (int j = 0; j < 32; j++) c[j] = fmaf(c[j], a, b);
    for (int j = 0; j < 32; j++) acc += c[j];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* clk;
  cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
  cudaMalloc(&clk, sizeof(long long) * sms * 4);
  const int iters = 4096;
  const char* names[4] = {"mma.sync m16n8k8 tf32", "mma.sync m8n8k4 f64", "ffma", "ffma2 (packed fp32)"};
  const double fma_per_inst[4] = {16.0 * 8 * 8, 8.0 * 8 * 4, 32.0, 64.0};
  for (int mode = 0; mode < 4; mode++)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; rep++) {
        if (mode == 0) rate_kernel<0><<<sms, warps * 32>>>(out, iters, clk);
        if (mode == 1) rate_kernel<1><<<sms, warps * 32>>>(out, iters, clk);
        if (mode == 2) rate_kernel<2><<<sms, warps * 32>>>(out, iters, clk);
        if (mode == 3) rate_kernel<3><<<sms, warps * 32>>>(out, iters, clk);
        cudaDeviceSynchronize();
      }
      long long h[1024];
      cudaMemcpy(h, clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      double mean = 0;
      for (int i = 0; i < sms; i++) mean += h[i];
      mean /= sms;
      const double insts = (double)iters * (mode == 2 ? 32 : (mode == 3 ? 16 : 8)) * warps;
      printf("%-24s warps/SM %2d: %8.1f multiply-adds / clk / SM  (%.2f warp-instructions / clk / SM)\n", names[mode], warps,
             insts * fma_per_inst[mode] / mean, insts / mean);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
