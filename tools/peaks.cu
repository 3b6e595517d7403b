// peaks.cu -- measured SIMT / tensor peaks of the pipes the BA and KLT kernels run on (B200):
// fp64 FMA (DFMA), fp64 tensor (DMMA m8n8k4), fp32 FMA (FFMA).  Register-only kernels, 148 x 8 CTAs
// of 256 threads, best of 5, CUDA events.  Prints one JSON object; bench.py reads the committed copy
// profiles/r2_peaks.json for the fp64 / fp32 roofline denominators ("measured").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/peaks tools/peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double* out, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fma_rn(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

// dependent-chain latencies, one warp, cycles per operation
__global__ void k_lat(long long* out, double a, double b, float fa) {
  __shared__ double sm[64];
  sm[threadIdx.x] = a;
  sm[threadIdx.x + 32] = b;
  __syncwarp();
  const int N = 2048;
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) x = __fma_rn(x, a, b);
  long long t1 = clock64();
  out[0] = t1 - t0;
  double y = x;
  t0 = clock64();
  for (int i = 0; i < N; ++i) y = __dmul_rn(y, a);
  t1 = clock64();
  out[1] = t1 - t0;
  float f = fa + threadIdx.x;
  t0 = clock64();
  for (int i = 0; i < N; ++i) f = __fmaf_rn(f, fa, 1e-9f);
  t1 = clock64();
  out[2] = t1 - t0;
  double z = y + 2.0;
  t0 = clock64();
  for (int i = 0; i < N; ++i) z = (double)rsqrtf((float)z) + 1.5;  // cvt + MUFU.RSQ + cvt + DADD
  t1 = clock64();
  out[3] = t1 - t0;
  double w = z;
  t0 = clock64();
  for (int i = 0; i < N; ++i) w = __shfl_xor_sync(0xffffffffu, w, 1);
  t1 = clock64();
  out[4] = t1 - t0;
  int idx = threadIdx.x & 31;
  double v = w;
  t0 = clock64();
  for (int i = 0; i < N; ++i) {
    v = sm[idx];
    idx = (idx + (int)v) & 31;
  }
  t1 = clock64();
  out[5] = t1 - t0;
  double c0 = x, c1 = y;
  t0 = clock64();
  for (int i = 0; i < N; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
  t1 = clock64();
  out[6] = t1 - t0;
  out[7] = N;
  if (x + y + f + z + w + v + c0 + c1 == 12345.0) out[8] = 1;
}

template <typename F>
static double best_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) {
    std::printf("{\"error\": \"no CUDA device\"}\n");
    return 1;
  }
  void* buf;
  cudaMalloc(&buf, 64);
  const int grid = prop.multiProcessorCount * 8, block = 256;
  const double nthr = (double)grid * block;
  const double msD = best_ms([&] { k_dfma<<<grid, block>>>((double*)buf, 1.0000001, 1e-9); });
  const double msF = best_ms([&] { k_ffma<<<grid, block>>>((float*)buf, 1.0000001f, 1e-9f); });
  const double msM = best_ms([&] { k_dmma<<<grid, block>>>((double*)buf, 1.0000001, 1e-9); });
  const double dfma = 2.0 * 8 * ITERS * nthr / (msD * 1e-3) / 1e12;
  const double ffma = 2.0 * 8 * ITERS * nthr / (msF * 1e-3) / 1e12;
  // one DMMA m8n8k4 per warp = 8*8*4 multiply-adds
  const double dmma = 2.0 * 256.0 * 8 * (ITERS / 4) * (nthr / 32) / (msM * 1e-3) / 1e12;
  long long* lat;
  cudaMallocManaged(&lat, 16 * sizeof(long long));
  k_lat<<<1, 32>>>(lat, 1.0000001, 1e-9, 1.0000001f);
  cudaDeviceSynchronize();
  k_lat<<<1, 32>>>(lat, 1.0000001, 1e-9, 1.0000001f);
  cudaDeviceSynchronize();
  const double N = (double)lat[7];
  std::printf("{\"latency_cycles\": {\"dfma\": %.1f, \"dmul\": %.1f, \"ffma\": %.1f, \"cvt_rsqrtf_cvt_dadd\": %.1f, "
              "\"shfl_f64\": %.1f, \"lds_f64_dependent\": %.1f, \"dmma884\": %.1f}, ",
              lat[0] / N, lat[1] / N, lat[2] / N, lat[3] / N, lat[4] / N, lat[5] / N, lat[6] / N);
  std::printf("\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.2f, \"fp64_dmma_tflops\": %.2f, "
              "\"fp32_fma_tflops\": %.2f, \"how\": \"register-only FMA / DMMA.8x8x4 chains, 8 independent "
              "accumulators per thread, %d CTAs x 256 threads, best of 5, CUDA events\"}\n",
              prop.name, prop.multiProcessorCount, dfma, dmma, ffma, grid);
  return 0;
}
