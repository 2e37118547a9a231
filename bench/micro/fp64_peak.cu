// fp64_peak.cu -- measured FP64 FMA rate of this GPU (roofline denominator of the PDHMM kernels and of
// the PairHMM fp64 rerun; MEASURED_PEAKS.json has no fp64 entry).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
// N_ACC independent accumulator chains per thread so that latency never binds.
#include <cuda_runtime.h>
#include <stdio.h>

#define N_ACC 16
#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, const double* in, int iters) {
  double acc[N_ACC], a[N_ACC], b[N_ACC];
#pragma unroll
  for (int i = 0; i < N_ACC; i++) {
    acc[i] = in[threadIdx.x + i];
    a[i] = in[threadIdx.x + 64 + i];
    b[i] = in[threadIdx.x + 128 + i];
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < N_ACC; i++) {
      if (MODE == 0) acc[i] = fma(acc[i], a[0], b[0]);   // DFMA, shared operands
      if (MODE == 1) acc[i] = fma(acc[i], a[i], b[i]);   // DFMA, 3 distinct operands
      if (MODE == 2) acc[i] = acc[i] * a[i];             // DMUL
      if (MODE == 3) acc[i] = acc[i] + a[i];             // DADD
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < N_ACC; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double flops_per_inst, int sms, double* out, double* in) {
  const int grid = sms * 8, block = 256;
  k<MODE><<<grid, block>>>(out, in, 16);
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
  const double insts = (double)grid * block * ITERS * N_ACC;
  printf("{\"mode\": \"%s\", \"ms\": %.4f, \"tflops\": %.3f, \"warp_inst_per_s_per_sm\": %.4e}\n", name, best,
         insts * flops_per_inst / (best * 1e-3) / 1e12, insts / 32.0 / (best * 1e-3) / sms);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double *out, *in;
  cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 256);
  cudaMalloc(&in, sizeof(double) * 1024);
  cudaMemset(in, 0, sizeof(double) * 1024);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("dfma_shared_operands", 2, p.multiProcessorCount, out, in);
  run<1>("dfma_3_distinct", 2, p.multiProcessorCount, out, in);
  run<2>("dmul_2_distinct", 1, p.multiProcessorCount, out, in);
  run<3>("dadd_2_distinct", 1, p.multiProcessorCount, out, in);
  return 0;
}
