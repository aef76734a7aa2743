// pcie.cu -- host<->device copy ceilings for the e2e pipeline (pinned memory, one GPU).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o pcie pcie.cu ; run on the GPU box.
// Prints GB/s for: 1-D H2D, 1-D D2H, both directions at once, and the 2-D (row-sliced) copies
// AADGpu_EncodeBatch / AADGpu_DecodeBatch issue (12,500 rows of ~55 KiB at a 882,000-byte pitch).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static float timed(cudaStream_t s0, cudaStream_t s1, void (*fn)(void *), void *arg)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a, s0));
  fn(arg);
  if (s1) { cudaEvent_t j; CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming)); CK(cudaEventRecord(j, s1)); CK(cudaStreamWaitEvent(s0, j, 0)); }
  CK(cudaEventRecord(b, s0));
  CK(cudaEventSynchronize(b));
  float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

struct Ctx { char *h, *h2, *d, *d2; size_t bytes; cudaStream_t s0, s1; size_t rows, width, pitch; };

static void h2d(void *p) { Ctx *c = (Ctx *)p; CK(cudaMemcpyAsync(c->d, c->h, c->bytes, cudaMemcpyHostToDevice, c->s0)); }
static void d2h(void *p) { Ctx *c = (Ctx *)p; CK(cudaMemcpyAsync(c->h, c->d, c->bytes, cudaMemcpyDeviceToHost, c->s0)); }
static void both(void *p) { Ctx *c = (Ctx *)p; CK(cudaMemcpyAsync(c->d, c->h, c->bytes, cudaMemcpyHostToDevice, c->s0)); CK(cudaMemcpyAsync(c->h2, c->d2, c->bytes, cudaMemcpyDeviceToHost, c->s1)); }
static void h2d_2d(void *p) { Ctx *c = (Ctx *)p; CK(cudaMemcpy2DAsync(c->d, c->pitch, c->h, c->pitch, c->width, c->rows, cudaMemcpyHostToDevice, c->s0)); }
static void d2h_2d(void *p) { Ctx *c = (Ctx *)p; CK(cudaMemcpy2DAsync(c->h, c->pitch, c->d, c->pitch, c->width, c->rows, cudaMemcpyDeviceToHost, c->s0)); }

int main()
{
  Ctx c;
  c.bytes = (size_t)2 << 30;
  CK(cudaMallocHost(&c.h, c.bytes)); CK(cudaMallocHost(&c.h2, c.bytes));
  CK(cudaMalloc(&c.d, c.bytes)); CK(cudaMalloc(&c.d2, c.bytes));
  memset(c.h, 1, c.bytes); memset(c.h2, 2, c.bytes);
  CK(cudaStreamCreateWithFlags(&c.s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&c.s1, cudaStreamNonBlocking));
  for (int rep = 0; rep < 2; rep++) {
    float a = timed(c.s0, 0, h2d, &c), b = timed(c.s0, 0, d2h, &c), d = timed(c.s0, c.s1, both, &c);
    printf("1-D 2 GiB: H2D %.1f GB/s  D2H %.1f GB/s  both at once %.1f GB/s aggregate\n", c.bytes / a / 1e6, c.bytes / b / 1e6,
           2.0 * c.bytes / d / 1e6);
  }
  c.pitch = 882000;
  for (size_t width = 8192; width <= 131072; width *= 4) {
    c.width = width;
    c.rows = c.bytes / c.pitch;
    float a = timed(c.s0, 0, h2d_2d, &c), b = timed(c.s0, 0, d2h_2d, &c);
    printf("2-D %zu rows x %zu B (pitch %zu): H2D %.1f GB/s  D2H %.1f GB/s\n", c.rows, c.width, c.pitch,
           c.rows * c.width / a / 1e6, c.rows * c.width / b / 1e6);
  }
  return 0;
}
