// tools/peaks.cu -- measured instruction-pipe peaks of the device the bench runs on.
//
// MEASURED_PEAKS.json only has HBM GB/s and bf16 TFLOP/s; the DP level kernels are bound by the
// FP32/FP64 CUDA-core issue rate and the MUFU (SFU) pipe instead, so bench.py runs this
// binary once per round and uses its numbers as roofline denominators ("of measured").
// Each test keeps 8 independent dependency chains per thread, 1024 threads x 2 CTAs per SM, and is
// timed with CUDA events after a warm-up launch.  Output: one JSON object, ops are per-lane
// operations per second for the whole device (an FMA counts as ONE op here; x2 for FLOP/s).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CHECK(x)                                                                   \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e));                      \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int ITERS = 4096;
constexpr int CH = 8;

enum Kind { K_FADD, K_FMUL, K_FFMA, K_FFMA2, K_FADD2, K_DADD, K_DFMA, K_RCP, K_LG2, K_F2D, K_D2F, K_I2F, K_FSEL,
            K_IADD, K_MIX };

template <int KIND>
__global__ void __launch_bounds__(1024) bench_kernel(float* out, float seed) {
  float x[CH];
  double d[CH];
  float2 p[CH];
  int q[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    x[c] = seed + threadIdx.x * 1e-3f + c;
    d[c] = (double)x[c];
    p[c] = make_float2(x[c], x[c] + 0.5f);
    q[c] = threadIdx.x + c;
  }
  const float m = 1.0000001f, ad = 1e-7f;
  const double dm = 1.0000000001, dad = 1e-9;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (KIND == K_FADD) x[c] = __fadd_rn(x[c], ad);
      if (KIND == K_FMUL) x[c] = __fmul_rn(x[c], m);
      if (KIND == K_FFMA) x[c] = __fmaf_rn(x[c], m, ad);
      if (KIND == K_FFMA2) p[c] = __ffma2_rn(p[c], make_float2(m, m), make_float2(ad, ad));
      if (KIND == K_FADD2) p[c] = __fadd2_rn(p[c], make_float2(ad, ad));
      if (KIND == K_DADD) d[c] = __dadd_rn(d[c], dad);
      if (KIND == K_DFMA) d[c] = __fma_rn(d[c], dm, dad);
      if (KIND == K_RCP) {
        // rcp(rcp(x)) back to back is folded away by ptxas (round 1 reported 1.6e15/s, 350x the pipe); the add between two
        // reciprocals (FP32 pipe, 8x the MUFU rate) keeps every MUFU.RCP in the loop
        asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
        x[c] = __fadd_rn(x[c], ad);
      }
      if (KIND == K_LG2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(x[c]));
      if (KIND == K_F2D) {
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[c]) : "f"(x[c]));
        asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(x[c]) : "d"(d[c]));
      }
      if (KIND == K_I2F) {
        asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(x[c]) : "r"(q[c]));
        asm volatile("mov.b32 %0, %1;" : "=r"(q[c]) : "f"(x[c]));
      }
      if (KIND == K_FSEL) x[c] = (x[c] > seed) ? x[c] : p[c].x;
      if (KIND == K_IADD) q[c] = q[c] * 3 + 1;
    }
  }
  float acc = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) acc += x[c] + (float)d[c] + p[c].x + p[c].y + (float)q[c];
  if (acc == 12345.678f) out[0] = acc;
}

template <int KIND>
double run(int sms, float* d_out, double ops_per_iter_chain) {
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  dim3 grid(sms * 2), block(1024);
  bench_kernel<KIND><<<grid, block>>>(d_out, 1.0f);
  CHECK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    CHECK(cudaEventRecord(e0));
    bench_kernel<KIND><<<grid, block>>>(d_out, 1.0f);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    float ms;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  double ops = (double)grid.x * block.x * ITERS * CH * ops_per_iter_chain;
  return ops / (best * 1e-3);
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  float* d_out;
  CHECK(cudaMalloc(&d_out, 64));
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d", prop.name, sms, clk);
  printf(", \"fadd_ops\": %.4e", run<K_FADD>(sms, d_out, 1));
  printf(", \"fmul_ops\": %.4e", run<K_FMUL>(sms, d_out, 1));
  printf(", \"ffma_ops\": %.4e", run<K_FFMA>(sms, d_out, 1));
  printf(", \"ffma2_lane_ops\": %.4e", run<K_FFMA2>(sms, d_out, 2));
  printf(", \"fadd2_lane_ops\": %.4e", run<K_FADD2>(sms, d_out, 2));
  printf(", \"dadd_ops\": %.4e", run<K_DADD>(sms, d_out, 1));
  printf(", \"dfma_ops\": %.4e", run<K_DFMA>(sms, d_out, 1));
  printf(", \"mufu_rcp_ops\": %.4e", run<K_RCP>(sms, d_out, 1));
  printf(", \"mufu_lg2_ops\": %.4e", run<K_LG2>(sms, d_out, 1));
  printf(", \"cvt_f32_f64_roundtrip_ops\": %.4e", run<K_F2D>(sms, d_out, 2));
  printf(", \"i2f_ops\": %.4e", run<K_I2F>(sms, d_out, 1));
  printf(", \"fsel_ops\": %.4e", run<K_FSEL>(sms, d_out, 1));
  printf(", \"imad_ops\": %.4e", run<K_IADD>(sms, d_out, 1));
  printf("}\n");
  return 0;
}
