// microbench_fp64_lds.cu -- the two per-SM ceilings the STRIP kernels sit under, measured on the B200 itself:
//   (1) FP64 pipe: DFMA latency of a dependent chain, and DFMA throughput per SM as a function of resident warps per
//       scheduler and independent chains per thread (the staged kernels run 4 warps per scheduler);
//   (2) shared-memory data pipe: LDS.64 / LDS.128 wavefronts per clock per SM;
//   (3) a loop with the instruction mix of staged_momentum_kernel's body (per 3 strip entries: 270 FP64, 14 LDS.128,
//       7 LDS.64, 3 STS.64) and no dependencies on memory: what the mix could do with perfect latency hiding.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp64_lds microbench_fp64_lds.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      printf("cuda error %s at line %d\n", cudaGetErrorString(e_), __LINE__);   \
      exit(1);                                                                  \
    }                                                                           \
  } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += x[i];
  if (s == 123.456) out[0] = s;
}

template <int VEC>  // 1: LDS.64, 2: LDS.128
__global__ void lds_kernel(double* out, int iters) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
  __syncthreads();
  // conflict-free: lane l reads 8 * VEC bytes at l * 8 * VEC; the address moves with the iteration (no hoisting) and
  // the consumer is an integer XOR (the FP64 pipe must not be what is measured)
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x % 32 * (VEC * 8) + (threadIdx.x / 32 % 4) * 512;
  int s = 0;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const unsigned a = base + (unsigned)(((u + it) & 15) * 2048);
      if (VEC == 2) {
        double2 v;
        asm volatile("ld.volatile.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory");
        s ^= __double2hiint(v.x) ^ __double2loint(v.x) ^ __double2hiint(v.y) ^ __double2loint(v.y);
      } else {
        double v;
        asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
        s ^= __double2hiint(v) ^ __double2loint(v);
      }
    }
  }
  if (s == 123456) out[0] = s;
}

// the momentum loop's mix: per "entry" 90 FP64 in CH independent chains, 5 LDS.128 (4 records + oldu xy), 2 LDS.64,
// 1 STS.64; loads feed the chains (as the records do) so nothing can be dropped
template <int CH>
__global__ void mix_kernel(double* out, int iters, double a, double b) {
  extern __shared__ __align__(16) double sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = 1e-3 * i;
  __syncthreads();
  const unsigned lane16 = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x % 128) * 16;
  const unsigned lane8 = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x % 128) * 8;
  double x[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; it++) {
    double2 r[5];
    double q[2];
#pragma unroll
    for (int u = 0; u < 5; u++)
      asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(r[u].x), "=d"(r[u].y) : "r"(lane16 + u * 2048 + (it & 3) * 16));
#pragma unroll
    for (int u = 0; u < 2; u++)
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(q[u]) : "r"(lane8 + 32768 + u * 1024 + (it & 3) * 8));
#pragma unroll
    for (int i = 0; i < CH; i++) x[i] += (i < 5 ? r[i % 5].x + r[i % 5].y : q[i & 1]);  // CH extra DADDs: part of the 90
#pragma unroll
    for (int u = 0; u < (90 - 2 * CH) / CH; u++) {
#pragma unroll
      for (int i = 0; i < CH; i++) x[i] = fma(x[i], a, b);
    }
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(lane8 + 40960), "d"(x[0]));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += x[i];
  if (s == 123.456) out[0] = s;
}

static float time_ms(void (*launch)(void*), void* ctx) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  launch(ctx);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  launch(ctx);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

struct Ctx {
  int kind, ilp, warps_per_smsp, iters, blocks;
  double* out;
};

static void launch(void* p) {
  Ctx* c = (Ctx*)p;
  const int threads = c->warps_per_smsp * 4 * 32;  // one block per SM
  if (c->kind == 0) {
    switch (c->ilp) {
      case 1: dfma_kernel<1><<<c->blocks, threads>>>(c->out, c->iters, 1.0000001, 1e-9); break;
      case 2: dfma_kernel<2><<<c->blocks, threads>>>(c->out, c->iters, 1.0000001, 1e-9); break;
      case 3: dfma_kernel<3><<<c->blocks, threads>>>(c->out, c->iters, 1.0000001, 1e-9); break;
      case 4: dfma_kernel<4><<<c->blocks, threads>>>(c->out, c->iters, 1.0000001, 1e-9); break;
      default: dfma_kernel<8><<<c->blocks, threads>>>(c->out, c->iters, 1.0000001, 1e-9); break;
    }
  } else if (c->kind == 1) {
    lds_kernel<1><<<c->blocks, threads, 65536>>>(c->out, c->iters);
  } else if (c->kind == 2) {
    lds_kernel<2><<<c->blocks, threads, 65536>>>(c->out, c->iters);
  } else {
    switch (c->ilp) {
      case 1: mix_kernel<1><<<c->blocks, threads, 98304>>>(c->out, c->iters, 1.0000001, 1e-9); break;
      case 2: mix_kernel<2><<<c->blocks, threads, 98304>>>(c->out, c->iters, 1.0000001, 1e-9); break;
      case 3: mix_kernel<3><<<c->blocks, threads, 98304>>>(c->out, c->iters, 1.0000001, 1e-9); break;
      default: mix_kernel<5><<<c->blocks, threads, 98304>>>(c->out, c->iters, 1.0000001, 1e-9); break;
    }
  }
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz * 1e-6;
  const int sms = prop.multiProcessorCount;
  printf("# %s, %d SMs, %.3f GHz (nominal max; cycles below assume it)\n", prop.name, sms, ghz);
  double* out;
  CK(cudaMalloc(&out, 64));
  CK(cudaFuncSetAttribute(lds_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(lds_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(mix_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  CK(cudaFuncSetAttribute(mix_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  CK(cudaFuncSetAttribute(mix_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  CK(cudaFuncSetAttribute(mix_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  // (1) latency: one warp on one SM, one chain
  {
    Ctx c{0, 1, 1, 20000, 1, out};
    // one block of 128 threads = one warp per scheduler: each scheduler sees a pure dependent chain
    const float ms = time_ms(launch, &c);
    printf("dfma dependent chain: %.2f cycles per DFMA (1 warp per scheduler)\n", ms * 1e-3 * ghz * 1e9 / (c.iters * 8.0));
  }
  // (2) throughput table
  printf("dfma throughput, warp-instructions per clock per SM (pipe peak = 2.0 if 64 lanes/SM):\n  warps/sched:");
  const int ws[] = {1, 2, 3, 4, 6, 8};
  for (int w : ws) printf(" %6d", w);
  printf("\n");
  for (int ilp : {1, 2, 3, 4, 8}) {
    printf("  ILP %d      :", ilp);
    for (int w : ws) {
      Ctx c{0, ilp, w, 4000, sms, out};
      const float ms = time_ms(launch, &c);
      const double instr = (double)c.iters * 8 * ilp * w * 4;  // warp instructions per SM
      printf(" %6.3f", instr / (ms * 1e-3 * ghz * 1e9));
    }
    printf("\n");
  }
  // (3) LDS wavefronts
  for (int kind : {1, 2}) {
    printf("LDS.%d, wavefronts (128 B) per clock per SM:", kind == 1 ? 64 : 128);
    for (int w : {1, 2, 4, 8}) {
      Ctx c{kind, 1, w, 4000, sms, out};
      const float ms = time_ms(launch, &c);
      const double wf = (double)c.iters * 8 * w * 4 * (kind == 1 ? 2 : 4);
      printf("  %dw: %.3f", w, wf / (ms * 1e-3 * ghz * 1e9));
    }
    printf("\n");
  }
  // (4) the loop's mix
  printf("momentum-loop mix (90 FP64 + 5 LDS.128 + 2 LDS.64 + 1 STS.64 per entry), cycles per entry per scheduler\n"
         "(FP64 floor 180 = 90 x 2; shared-memory floor 26 wavefronts x 4 schedulers = 104):\n  warps/sched:");
  for (int w : {1, 2, 3, 4, 6}) printf(" %6d", w);
  printf("\n");
  for (int ch : {1, 2, 3, 5}) {
    printf("  chains %d   :", ch);
    for (int w : {1, 2, 3, 4, 6}) {
      Ctx c{3, ch, w, 3000, sms, out};
      const float ms = time_ms(launch, &c);
      printf(" %6.1f", ms * 1e-3 * ghz * 1e9 / ((double)c.iters * w));
    }
    printf("\n");
  }
  return 0;
}
