// Pipe-throughput microbenchmark for sm_100a (B200): measures warp-instructions
// per clock per SM for the instruction classes the NLM kernel is built from.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_microbench pipe_microbench.cu
// Output: one line per test: name, warp-instr/clk/SM, lane-ops/clk/SM, measured SM MHz.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__);exit(1);} }while(0)

constexpr int ITERS = 4096;
constexpr int UNROLL = 16;   // independent chains per thread

__global__ void k_ffma(float* out, float a, float b, long long* cyc) {
  float r[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r[i]) : "f"(a), "f"(b));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// 3 distinct register sources per FFMA (acc = x*y+acc with rotating x,y)
__global__ void k_ffma3(float* out, float a, float b, long long* cyc) {
  float r[UNROLL], x[4], y[4];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 4; i++) { x[i] = a + i * 1e-3f * threadIdx.x; y[i] = b - i * 1e-3f * threadIdx.x; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(r[i]) : "f"(x[i & 3]), "f"(y[(i >> 2) & 3]));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ffma2(float* out, float a, float b, long long* cyc) {
  unsigned long long r[UNROLL], x[4], y[4];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) { float2 v = make_float2(threadIdx.x * 0.001f + i, i); r[i] = *reinterpret_cast<unsigned long long*>(&v); }
#pragma unroll
  for (int i = 0; i < 4; i++) { float2 v = make_float2(a + i * 1e-3f * threadIdx.x, a); x[i] = *reinterpret_cast<unsigned long long*>(&v);
    float2 w = make_float2(b - i * 1e-3f * threadIdx.x, b); y[i] = *reinterpret_cast<unsigned long long*>(&w); }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(r[i]) : "l"(x[i & 3]), "l"(y[(i >> 2) & 3]));
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int i = 0; i < UNROLL; i++) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_fadd(float* out, float a, float b, long long* cyc) {
  float r[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(a));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_fadd2(float* out, float a, float b, long long* cyc) {
  unsigned long long r[UNROLL], x;
#pragma unroll
  for (int i = 0; i < UNROLL; i++) { float2 v = make_float2(threadIdx.x * 0.001f + i, i); r[i] = *reinterpret_cast<unsigned long long*>(&v); }
  { float2 v = make_float2(a, b); x = *reinterpret_cast<unsigned long long*>(&v); }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r[i]) : "l"(x));
  }
  long long t1 = clock64();
  unsigned long long s = 0; for (int i = 0; i < UNROLL; i++) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_fmnmx(float* out, float a, float b, long long* cyc) {
  float r[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("max.f32 %0, %0, %1;" : "+f"(r[i]) : "f"(a));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_ex2(float* out, float a, float b, long long* cyc) {
  float r[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = -threadIdx.x * 0.001f - i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i]));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(float* out, float a, float b, long long* cyc) {
  float r[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("shfl.sync.up.b32 %0, %0, 1, 0, 0xffffffff;" : "+f"(r[i]));
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dadd(float* out, float a, float b, long long* cyc) {
  double r[UNROLL]; const double x = a;
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001 + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(r[i]) : "d"(x));
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_dfma(float* out, float a, float b, long long* cyc) {
  double r[UNROLL]; const double x = a, y = b;
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001 + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(r[i]) : "d"(x), "d"(y));
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_f2d(float* out, float a, float b, long long* cyc) {
  float r[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) r[i] = threadIdx.x * 0.001f + i;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(r[i])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(r[i]) : "d"(d)); }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int VEC>
__global__ void k_lds(float* out, float a, float b, long long* cyc) {
  extern __shared__ float4 sm[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, a, b, 1.f);
  __syncthreads();
  float acc[UNROLL];
#pragma unroll
  for (int i = 0; i < UNROLL; i++) acc[i] = 0;
  unsigned base = (unsigned)__cvta_generic_to_shared(sm) + threadIdx.x * (VEC * 4);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < UNROLL; i++) {
      unsigned addr = base + ((it + i * 37) & 31) * 1024 * (VEC == 4 ? 1 : 1);
      if (VEC == 4) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr)); acc[i] += x; }
      else { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr)); acc[i] += x; }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < UNROLL; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// NLM-like mix per "voxel-offset pair": 4 FADD2/FMUL2/FFMA2-class + 3 FADD + 2 SHFL + FFMA + FMNMX + EX2 + 2 FFMA2 + FADD + FMNMX
__global__ void k_mix(float* out, float a, float b, long long* cyc) {
  unsigned long long p[4], q[4], acc0[4], acc1[4];
  float sw[4], mw[4], t[4];
#pragma unroll
  for (int i = 0; i < 4; i++) { float2 v = make_float2(threadIdx.x * 0.001f + i, i); p[i] = *reinterpret_cast<unsigned long long*>(&v); q[i] = p[i] + 12345; acc0[i] = 0; acc1[i] = 0; sw[i] = 0; mw[i] = 0; t[i] = i; }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      unsigned long long d0, d1, s; float lo, hi, u, w;
      asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d0) : "l"(p[i]), "l"(q[i]));
      asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d1) : "l"(q[i]), "l"(p[(i + 1) & 3]));
      asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(s) : "l"(d0));
      asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(s) : "l"(d1));
      asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(s));
      asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(u) : "f"(lo), "f"(hi));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(u) : "f"(t[i]));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(u) : "f"(t[(i + 1) & 3]));
      float v1, v2;
      asm volatile("shfl.sync.up.b32 %0, %1, 1, 0, 0xffffffff;" : "=f"(v1) : "f"(u));
      asm volatile("shfl.sync.down.b32 %0, %1, 1, 31, 0xffffffff;" : "=f"(v2) : "f"(u));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(u) : "f"(v1));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(u) : "f"(v2));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(u) : "f"(a), "f"(b));
      asm volatile("min.f32 %0, %0, 0f00000000;" : "+f"(u));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(u));
      unsigned long long w2;
      asm volatile("mov.b64 %0, {%1,%1};" : "=l"(w2) : "f"(w));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc0[i]) : "l"(w2), "l"(q[i]));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc1[i]) : "l"(w2), "l"(p[i]));
      asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(sw[i]) : "f"(w));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(mw[i]) : "f"(w));
      t[i] = u;
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 4; i++) s += sw[i] + mw[i] + (float)(acc0[i] ^ acc1[i]) + t[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

struct Test { const char* name; void (*fn)(float*, float, float, long long*); double instr_per_iter; double lanes_mult; size_t smem; };

int main() {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
  int sms = pr.multiProcessorCount;
  printf("device %s sms %d clockRate(kHz) %d\n", pr.name, sms, pr.clockRate);
  float* out; long long* cyc; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024)); CK(cudaMalloc(&cyc, 8));
  CK(cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  std::vector<Test> tests = {
    {"FFMA(r,imm-ish 2 uniform srcs)", k_ffma, UNROLL, 1, 0},
    {"FFMA 3-reg", k_ffma3, UNROLL, 1, 0},
    {"FFMA2 3-reg", k_ffma2, UNROLL, 2, 0},
    {"FADD", k_fadd, UNROLL, 1, 0},
    {"FADD2", k_fadd2, UNROLL, 2, 0},
    {"FMNMX", k_fmnmx, UNROLL, 1, 0},
    {"MUFU.EX2", k_ex2, UNROLL, 1, 0},
    {"SHFL.UP", k_shfl, UNROLL, 1, 0},
    {"DADD", k_dadd, UNROLL, 1, 0},
    {"DFMA", k_dfma, UNROLL, 1, 0},
    {"F2F f32->f64->f32 (2 instr)", k_f2d, 2 * UNROLL, 1, 0},
    {"LDS.128", k_lds<4>, UNROLL, 1, 65536},
    {"LDS.32", k_lds<1>, UNROLL, 1, 65536},
    {"NLM-mix (22 instr / pair-step)", k_mix, 4 * 22, 1, 0},
  };
  for (int threads : {1024}) {
    for (auto& t : tests) {
      int blocks = sms * (t.smem ? 1 : (1024 / threads));
      if (t.smem && threads != 1024) continue;
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      t.fn<<<blocks, threads, t.smem>>>(out, 1.0001f, 0.5f, cyc); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      t.fn<<<blocks, threads, t.smem>>>(out, 1.0001f, 0.5f, cyc);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      long long c; CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
      double warps_per_sm = (double)blocks * threads / 32 / sms;
      double winstr = warps_per_sm * t.instr_per_iter * ITERS;     // warp-instr per SM
      double mhz = c / (ms * 1e3);
      printf("thr/SM=%4d %-34s cycles=%9lld  warp-instr/clk/SM=%6.3f  lane-ops/clk/SM=%7.2f  ms=%.3f  ~SM MHz(clock64/event)=%.0f\n",
             (int)(warps_per_sm * 32), t.name, c, winstr / c, winstr / c * 32 * t.lanes_mult, ms, mhz);
    }
  }
  return 0;
}
