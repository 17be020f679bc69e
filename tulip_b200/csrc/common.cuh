// Shared device helpers for the tulip_b200 sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define TULIP_OK 0
#define TULIP_ERR_ARG 1
#define TULIP_ERR_CUDA 2
#define TULIP_ERR_UNSUPPORTED 3

// Launch-error check used by every host launcher: returns a TULIP_ERR_* code, never exits.
#define TULIP_CHECK_LAUNCH()                                   \
  do {                                                         \
    cudaError_t e__ = cudaGetLastError();                      \
    if (e__ != cudaSuccess) { tulip_set_error(cudaGetErrorString(e__)); return TULIP_ERR_CUDA; } \
  } while (0)

#define TULIP_CUDA(call)                                       \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) { tulip_set_error(cudaGetErrorString(e__)); return TULIP_ERR_CUDA; } \
  } while (0)

#define TULIP_REQUIRE(cond, msg)                               \
  do {                                                         \
    if (!(cond)) { tulip_set_error(msg); return TULIP_ERR_ARG; } \
  } while (0)

void tulip_set_error(const char* msg);
// L2 eviction-priority hints (bit mask, env TULIP_B200_HINTS, default 127 = all): which operand streams are marked evict_first /
// evict_last.  Stage-0/1 tensors are 25-100 MB each and L2 holds 126 MB: loads of data that is read once (activations saved by
// the forward pass, operands of the weight-gradient GEMMs, residual / LayerNorm-input tiles) leave the tensors the NEXT kernel
// re-reads, and the weights every CTA shares, in L2.  Measured on the training step, same box: 4.92 ms without -> 4.80 ms.
//   1 weight-gradient GEMM operands first | 2 NT GEMM A operand first (resident-weight schedule) | 4 NT GEMM A first (tile-major)
//   8 NT GEMM auxiliary tiles (residual, LayerNorm inputs) first | 16 fused MLP / head-backward kernel: input tile and by-product
//   stores first | 32 attention loads first | 64 NT GEMM / fused-kernel weight loads last
// (a bit that is off selects the evict_normal policy on the same instruction form: the NT GEMM's TMA issue path has no branches)
int tulip_hints();
int tulip_num_sms();
bool tulip_pdl_enabled();          // env TULIP_B200_NO_PDL=1 turns programmatic dependent launch off

// Programmatic dependent launch (PDL): every kernel of the step is launched with the stream-serialization attribute, does
// its private set-up (barrier init, TMEM allocation, shared-memory tables), then waits for the preceding kernel's memory
// with griddepcontrol.wait BEFORE its first global access, and immediately lets its own successor start scheduling.
// The step is ~330 short kernels; overlapping each prologue with the predecessor's tail removes a few us per launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_sync() { pdl_trigger(); pdl_wait(); }

template <class... KArgs, class... Args>
inline void tulip_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tulip_pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);       // errors surface through cudaGetLastError (TULIP_CHECK_LAUNCH)
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Column sums over the 32 lanes of a warp for 32 per-lane values: recursive halving (31 shuffles), lane i returns the sum of
// v[i] over all lanes.  Fixed tree, so the result does not depend on timing.  Destroys v.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
// two-term bf16 split of a pair: hi = bf16(x), lo = bf16(x - hi); hi + lo carries 16 significand bits of x
__device__ __forceinline__ void split_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);
  const float2 h = unpack_bf16(hi);
  lo = pack_bf16(a - h.x, b - h.y);
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// erf-form GELU (nn.GELU() default, reference tulip.py:183,196) and its derivative.
// erf via Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below a bf16 ulp), evaluated without cancellation:
//   Phi(x) = 1 - h (x >= 0),  h (x < 0),   h = 0.5 * poly(t) * exp(-x^2/2),  t = 1 / (1 + 0.3275911 |x| / sqrt 2)
// Two MUFU ops (rcp, ex2) + ~12 FMA-class instructions; the library erff costs ~3x that and made the GELU GEMM
// epilogues issue-bound.
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// h(x) = 0.5 * poly(t) * t * exp(-x^2/2) (the 0.5 is folded into the coefficients) and ex = exp(-x^2/2)
__device__ __forceinline__ float gelu_tail(float x, float& ex) {
  const float t = fast_rcp(fmaf(0.231641888f, fabsf(x), 1.0f));            // 0.3275911 / sqrt(2)
  ex = fast_ex2((x * -0.72134752f) * x);                                    // exp(-x^2/2)
  float p = fmaf(0.5307027145f, t, -0.7265760135f);
  p = fmaf(p, t, 0.7107068705f);
  p = fmaf(p, t, -0.142248368f);
  p = fmaf(p, t, 0.127414796f);
  return (p * t) * ex;
}
// gelu(x) = x * Phi(x) = max(x, 0) - |x| h   (x >= 0: x (1 - h);  x < 0: x h)
__device__ __forceinline__ float gelu_erf(float x) {
  float ex;
  const float h = gelu_tail(x, ex);
  return fmaf(-fabsf(x), h, fmaxf(x, 0.f));
}
// gelu'(x) = Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float ex;
  const float h = gelu_tail(x, ex);
  const float cdf = x >= 0.f ? 1.0f - h : h;
  return fmaf(x * 0.39894228040143268f, ex, cdf);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 16-byte async copy global->shared; src_bytes == 0 zero-fills (used for row tails)
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes));
}
// the same with an L2 eviction-priority policy (createpolicy) on the global read
__device__ __forceinline__ void cp_async16_pol(void* smem, const void* gmem, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "l"(policy));
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t (&r)[2], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(smem)));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n"
               : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(smem)));
}

// D(16x8,f32) += A(16x16,bf16,row) * B(16x8,bf16,col)
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
