// fp32_peak.cu -- measured FP32 issue rates on the CUDA cores of this GPU (roofline denominator
// for the PairHMM recurrence; MEASURED_PEAKS.json has no fp32 entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_peak fp32_peak.cu && ./fp32_peak
// Every kernel runs N_ACC independent accumulator chains per thread so that latency never binds;
// variants differ in how many distinct register operands each instruction reads.
#include <cuda_runtime.h>
#include <stdio.h>

#define N_ACC 16
#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float* in, int iters) {
  float acc[N_ACC], a[N_ACC], b[N_ACC];
#pragma unroll
  for (int i = 0; i < N_ACC; i++) {
    acc[i] = in[threadIdx.x + i];
    a[i] = in[threadIdx.x + 64 + i];
    b[i] = in[threadIdx.x + 128 + i];
  }
  float2 acc2[N_ACC / 2], a2[N_ACC / 2], b2[N_ACC / 2];
#pragma unroll
  for (int i = 0; i < N_ACC / 2; i++) {
    acc2[i] = make_float2(acc[2 * i], acc[2 * i + 1]);
    a2[i] = make_float2(a[2 * i], a[2 * i + 1]);
    b2[i] = make_float2(b[2 * i], b[2 * i + 1]);
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < N_ACC; i++) {
      if (MODE == 0) acc[i] = fmaf(acc[i], a[0], b[0]);          // FFMA, 1 varying + 2 shared operands
      if (MODE == 1) acc[i] = fmaf(acc[i], a[i], b[i]);          // FFMA, 3 distinct operands per instruction
      if (MODE == 2) acc[i] = acc[i] * a[i];                     // FMUL, 2 distinct
      if (MODE == 3) acc[i] = acc[i] + a[i];                     // FADD, 2 distinct
      if (MODE == 4) acc[i] = fmaf(a[i], b[(i + 1) % N_ACC], acc[i]);  // FFMA, accumulate form, 3 distinct
      if (MODE == 7) acc[i] = fmaf(acc[i], 1.0001f, 0.5f);      // FFMA immediates
    }
    if (MODE == 5 || MODE == 6) {
#pragma unroll
      for (int i = 0; i < N_ACC / 2; i++) {
        if (MODE == 5) acc2[i] = __ffma2_rn(acc2[i], a2[i], b2[i]);  // FFMA2, 3 distinct pairs
        if (MODE == 6) acc2[i] = __ffma2_rn(acc2[i], a2[0], b2[0]);  // FFMA2, shared operands
      }
    }
    if (MODE == 8 || MODE == 9 || MODE == 10) {
#pragma unroll
      for (int i = 0; i < N_ACC / 2; i++) {
        // FFMA2 whose first operand is a 32-bit scalar broadcast to both halves (SASS R.F32), as in the
        // two-haplotypes-per-lane sweep: a distinct scalar per instruction
        if (MODE == 8) acc2[i] = __ffma2_rn(make_float2(a[i], a[i]), acc2[i], b2[i]);
        if (MODE == 9) acc2[i] = __fmul2_rn(make_float2(a[i], a[i]), acc2[i]);   // FMUL2, scalar x pair
        if (MODE == 10) acc2[i] = __fmul2_rn(a2[i], acc2[i]);                    // FMUL2, pair x pair
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < N_ACC; i++) s += acc[i];
#pragma unroll
  for (int i = 0; i < N_ACC / 2; i++) s += acc2[i].x + acc2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flops_per_inst, int insts_per_iter, int sms, float* out, float* in) {
  const int grid = sms * 8, block = 256;
  k<MODE><<<grid, block>>>(out, in, 64);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k<MODE><<<grid, block>>>(out, in, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double insts = (double)grid * block * ITERS * insts_per_iter;  // thread-instructions
  const double tflops = insts * flops_per_inst / (best * 1e-3) / 1e12;
  const double warp_inst_per_clk_sm = insts / 32.0 / (best * 1e-3) / sms;  // per second per SM
  printf("{\"mode\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"warp_inst_per_s_per_sm\": %.4e}\n", name, best, tflops,
         warp_inst_per_clk_sm);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  float *out, *in;
  cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
  cudaMalloc(&in, sizeof(float) * 1024);
  cudaMemset(in, 0, sizeof(float) * 1024);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("ffma_shared_operands", 2, N_ACC, p.multiProcessorCount, out, in);
  run<1>("ffma_3_distinct", 2, N_ACC, p.multiProcessorCount, out, in);
  run<4>("ffma_3_distinct_acc", 2, N_ACC, p.multiProcessorCount, out, in);
  run<2>("fmul_2_distinct", 1, N_ACC, p.multiProcessorCount, out, in);
  run<3>("fadd_2_distinct", 1, N_ACC, p.multiProcessorCount, out, in);
  run<7>("ffma_imm", 2, N_ACC, p.multiProcessorCount, out, in);
  run<5>("ffma2_3_distinct", 4, N_ACC / 2, p.multiProcessorCount, out, in);
  run<6>("ffma2_shared_operands", 4, N_ACC / 2, p.multiProcessorCount, out, in);
  run<8>("ffma2_scalar_a_2_pairs", 4, N_ACC / 2, p.multiProcessorCount, out, in);
  run<9>("fmul2_scalar_x_pair", 2, N_ACC / 2, p.multiProcessorCount, out, in);
  run<10>("fmul2_pair_x_pair", 2, N_ACC / 2, p.multiProcessorCount, out, in);
  return 0;
}
