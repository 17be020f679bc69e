// Bring-up probe for the fused W-MSA kernel's assumptions (run on the B200 box; prints PASS / FAIL per item):
//  1. UMMA with SWIZZLE_64B K-major operands written by threads (32-column K blocks, SBO = 512 B)
//  2. tcgen05.ld.16x256b.x4 register <-> (row, column) mapping, at lane offsets 0 and 16 inside a warp's TMEM quarter
//  3. movmatrix.m8n8.trans on packed bf16 fragments
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I tulip_b200/csrc -o /tmp/probe scripts/probes/probe_tmem_frag.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc05.cuh"

void tulip_set_error(const char*) {}
int tulip_num_sms() { return 148; }
bool tulip_pdl_enabled() { return false; }

constexpr int M = 128, N = 96, K = 96, KB = 32;     // three 32-column K blocks

__device__ __forceinline__ uint64_t desc_k_sw64(const void* smem) {
  const uint64_t addr = (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  return addr | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ int sw64_off(int r, int c16) { return r * 64 + ((c16 ^ ((r >> 1) & 3)) << 4); }

__device__ __forceinline__ void ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

__global__ void __launch_bounds__(128, 1) probe(const bf16* A, const bf16* B, float* d32, float* d16, uint32_t* mv) {
  extern __shared__ unsigned char raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;                      // 3 blocks x [128 x 32] bf16 = 3 x 8 KB
  unsigned char* sB = smem + 3 * 8192;           // 3 blocks x [96 x 32] = 3 x 6 KB
  uint64_t* bar = (uint64_t*)(smem + 3 * 8192 + 3 * 6144);
  uint32_t* slot = (uint32_t*)(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int kb = 0; kb < 3; ++kb)
    for (int c = 0; c < 4; ++c) {
      *(uint4*)(sA + kb * 8192 + sw64_off(tid, c)) = *(const uint4*)(A + tid * K + kb * KB + c * 8);
      if (tid < N) *(uint4*)(sB + kb * 6144 + sw64_off(tid, c)) = *(const uint4*)(B + tid * K + kb * KB + c * 8);
    }
  if (tid == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<128>(slot);
  tc::fence_proxy_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    if (tc::elect_one_sync()) {
      constexpr uint32_t idesc = tc::make_idesc(M, N, 0, 0);
      for (int kb = 0; kb < 3; ++kb)
        for (int ks = 0; ks < 2; ++ks)
          tc::umma_bf16(tmem, desc_k_sw64(sA + kb * 8192) + 2 * ks, desc_k_sw64(sB + kb * 6144) + 2 * ks, idesc, (kb | ks) ? 1u : 0u);
      tc::umma_commit(bar);
    }
    __syncwarp();
  }
  tc::mbar_wait(bar, 0);
  tc::fence_after_sync();
  // control read: 32x32b, thread = row
  for (int j = 0; j < 3; ++j) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + j * 32, v);
    for (int i = 0; i < 32; ++i) d32[(warp * 32 + lane) * N + j * 32 + i] = v[i];
  }
  // fragment read: 16x256b.x4 = 16 lanes x 32 columns; reg 4j+{0,1}: (row g, cols 8j+2t,+1); reg 4j+{2,3}: (row g+8, same cols)
  const int g = lane >> 2, t = lane & 3;
  for (int half = 0; half < 2; ++half)
    for (int j32 = 0; j32 < 3; ++j32) {
      uint32_t r[16];
      ld_16x256b_x4(tmem + ((uint32_t)(warp * 32 + half * 16) << 16) + j32 * 32, r);
      for (int j = 0; j < 4; ++j)
        for (int e = 0; e < 4; ++e) {
          const int row = warp * 32 + half * 16 + g + (e >> 1) * 8, col = j32 * 32 + 8 * j + 2 * t + (e & 1);
          d16[row * N + col] = __uint_as_float(r[4 * j + e]);
        }
    }
  // movmatrix: x[i][j] = 16 i + j as bf16; this thread holds (row g, cols 2t, 2t+1)
  if (warp == 0) {
    const uint32_t x = pack_bf16((float)(16 * g + 2 * t), (float)(16 * g + 2 * t + 1));
    mv[lane] = movm_t(x);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

int main() {
  std::vector<bf16> hA(M * K), hB(N * K);
  std::vector<float> ref(M * N);
  for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) hA[r * K + k] = __float2bfloat16((float)(((r * 7 + k * 3) % 11) - 5));
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[n * K + k] = __float2bfloat16((float)(((n * 5 + k * 2) % 9) - 4));
  for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
    float s = 0;
    for (int k = 0; k < K; ++k) s += __bfloat162float(hA[r * K + k]) * __bfloat162float(hB[n * K + k]);
    ref[r * N + n] = s;
  }
  bf16 *dA, *dB; float *d32, *d16; uint32_t* mv;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&d32, M * N * 4); cudaMalloc(&d16, M * N * 4); cudaMalloc(&mv, 128);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(d32, 0xff, M * N * 4); cudaMemset(d16, 0xff, M * N * 4);
  const int smem = 3 * 8192 + 3 * 6144 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dB, d32, d16, mv);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("FAIL launch: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> o32(M * N), o16(M * N); uint32_t hm[32];
  cudaMemcpy(o32.data(), d32, M * N * 4, cudaMemcpyDeviceToHost); cudaMemcpy(o16.data(), d16, M * N * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(hm, mv, 128, cudaMemcpyDeviceToHost);
  int bad32 = 0, bad16 = 0, bad16hi = 0;
  for (int i = 0; i < M * N; ++i) {
    if (o32[i] != ref[i]) ++bad32;
    if (o16[i] != ref[i]) { ++bad16; if (((i / N) % 32) >= 16) ++bad16hi; }
  }
  printf("%s umma sw64 k-major + 32x32b read: %d mismatches\n", bad32 ? "FAIL" : "PASS", bad32);
  printf("%s 16x256b.x4 fragment mapping: %d mismatches (%d in upper lane halves)\n", bad16 ? "FAIL" : "PASS", bad16, bad16hi);
  if (bad32) for (int i = 0, n = 0; i < M * N && n < 8; ++i) if (o32[i] != ref[i]) { printf("   32x32b [%d][%d] got %g want %g\n", i / N, i % N, o32[i], ref[i]); ++n; }
  if (bad16) for (int i = 0, n = 0; i < M * N && n < 8; ++i) if (o16[i] != ref[i]) { printf("   16x256b [%d][%d] got %g want %g\n", i / N, i % N, o16[i], ref[i]); ++n; }
  int badm = 0;
  for (int l = 0; l < 32; ++l) {
    const int g = l >> 2, t = l & 3;
    const float lo = __bfloat162float(__ushort_as_bfloat16((unsigned short)(hm[l] & 0xffff)));
    const float hi = __bfloat162float(__ushort_as_bfloat16((unsigned short)(hm[l] >> 16)));
    if (lo != (float)(16 * (2 * t) + g) || hi != (float)(16 * (2 * t + 1) + g)) ++badm;
  }
  printf("%s movmatrix.trans: %d lanes off\n", badm ? "FAIL" : "PASS", badm);
  return (bad32 || bad16 || badm) ? 2 : 0;
}
