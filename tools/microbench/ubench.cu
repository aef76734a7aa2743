// Integer-pipe microbenchmarks for sm_100a: which SASS ops cost what, so the sample chain can be
// balanced across the FMA and ALU pipes.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 ubench.cu -o ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int OP>
__global__ void __launch_bounds__(512) k(int *out, int a0, int b0, unsigned long long *cycles)
{
  __shared__ int sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (i * 2654435761u) >> 20 & 4095;
  __syncthreads();
  int r[8];
#pragma unroll
  for (int j = 0; j < 8; j++) r[j] = a0 + threadIdx.x * (j + 1);
  int b = b0, c = b0 ^ 0x55;
  long long w = a0;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (OP == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c));
      if (OP == 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[j]) : "r"(b), "r"(c));
      if (OP == 2) asm volatile("shr.s32 %0, %0, 3;" : "+r"(r[j]));
      if (OP == 3) asm volatile("add.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(b));
      if (OP == 4) asm volatile("min.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(b));
      if (OP == 5) r[j] = __viaddmin_s32(r[j], b, c);
      if (OP == 6) r[j] = __viaddmin_s32_relu(r[j], b, c);
      if (OP == 7) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c));
      if (OP == 8) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(b));
      if (OP == 9) { long long x; asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(x) : "r"(r[j]), "r"(b), "l"(w)); r[j] = (int)(x >> 32); }
      if (OP == 10) r[j] = sm[(r[j] & 4095)];                          // dependent LDS, random bank
      if (OP == 11) r[j] = sm[((r[j] & 127) << 5) | (threadIdx.x & 31)]; // LDS conflict-free
      if (OP == 12) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(j + 4) & 7]) : "r"(b), "r"(c)); }
      if (OP == 13) r[j] = __vimax3_s32(r[j], b, c);
      if (OP == 14) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c));
      if (OP == 15) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c));
      if (OP == 16) { unsigned long long x; asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(x) : "r"(r[j]), "r"(b), "l"((unsigned long long)w)); w = (long long)x; }
      if (OP == 17) asm volatile("bfe.u32 %0, %0, 4, 4;" : "+r"(r[j]));
      if (OP == 18) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.s32 %0, %1, %2, p;}" : "+r"(r[j]) : "r"(b), "r"(c));
      if (OP == 19) asm volatile("shl.b32 %0, %0, 3;" : "+r"(r[j]));
      if (OP == 20) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[j]) : "r"(b), "r"(c)); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[(j + 3) & 7]) : "r"(b), "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(j + 5) & 7]) : "r"(b), "r"(c)); }
      if (OP == 21) asm volatile("sub.s32 %0, %1, %0;" : "+r"(r[j]) : "r"(b));
      if (OP == 22) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r[j]));
      if (OP == 23) asm volatile("abs.s32 %0, %0;" : "+r"(r[j]));
      if (OP == 24) asm volatile("{.reg .s16 h; cvt.sat.s16.s32 h, %0; cvt.s32.s16 %0, h;}" : "+r"(r[j]));          // 16-bit clip in one op?
      if (OP == 25) { asm volatile("{.reg .s16 h; cvt.sat.s16.s32 h, %0; cvt.s32.s16 %0, h;}" : "+r"(r[j])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[(j + 4) & 7]) : "r"(b), "r"(c)); }
      if (OP == 26) { asm volatile("{.reg .s16 h; cvt.sat.s16.s32 h, %0; cvt.s32.s16 %0, h;}" : "+r"(r[j])); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[(j + 4) & 7]) : "r"(b), "r"(c)); }
      if (OP == 27) { asm volatile("min.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(b)); asm volatile("max.s32 %0, %0, %1;" : "+r"(r[j]) : "r"(c)); }
    }
  }
  long long t1 = clock64();
  int acc = (int)w;
#pragma unroll
  for (int j = 0; j < 8; j++) acc ^= r[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = (unsigned long long)(t1 - t0);
}

template <int OP>
int run(const char *name, int per_iter, int threads)
{
  int *out; unsigned long long *cyc, h = 0;
  CHECK(cudaMalloc(&out, 148 * 1024 * 4)); CHECK(cudaMalloc(&cyc, 8));
  k<OP><<<148, threads>>>(out, 3, 7, cyc);
  CHECK(cudaDeviceSynchronize());
  k<OP><<<148, threads>>>(out, 3, 7, cyc);
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  double warp_instr = (double)ITERS * 8 * per_iter * (threads / 32);
  printf("%-34s threads/SM=%4d  cycles=%9llu  warp-instr/clk/SM=%.3f  (clk per warp-instr per SMSP=%.3f)\n", name, threads, h,
         warp_instr / h, h / (warp_instr / 4));
  cudaFree(out); cudaFree(cyc);
  return 0;
}

int main()
{
  for (int threads : {128, 512}) {
    run<0>("IMAD (mad.lo)", 1, threads);
    run<1>("LOP3", 1, threads);
    run<2>("SHR.S32 imm", 1, threads);
    run<19>("SHL imm", 1, threads);
    run<3>("IADD", 1, threads);
    run<21>("SUB", 1, threads);
    run<4>("IMNMX (min)", 1, threads);
    run<5>("viaddmin", 1, threads);
    run<6>("viaddmin_relu", 1, threads);
    run<13>("vimax3", 1, threads);
    run<7>("PRMT", 1, threads);
    run<14>("SHF.R funnel", 1, threads);
    run<17>("BFE.U32", 1, threads);
    run<23>("ABS", 1, threads);
    run<18>("SETP+SELP", 2, threads);
    run<8>("MUL.HI", 1, threads);
    run<15>("MAD.HI", 1, threads);
    run<9>("MAD.WIDE.S32 (hi used)", 1, threads);
    run<16>("MAD.WIDE.U32 accumulate (dep)", 1, threads);
    run<22>("I2F", 1, threads);
    run<10>("LDS random (dependent)", 1, threads);
    run<11>("LDS conflict-free (dependent)", 1, threads);
    run<24>("cvt.sat.s16.s32 (clip16)", 1, threads);
    run<25>("clip16 cvt + LOP3 interleaved", 2, threads);
    run<26>("clip16 cvt + IMAD interleaved", 2, threads);
    run<27>("min + max (clip16 as today)", 2, threads);
    run<12>("IMAD + LOP3 interleaved", 2, threads);
    run<20>("2 IMAD + 1 LOP3 interleaved", 3, threads);
  }
  return 0;
}
