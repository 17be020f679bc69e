// Building blocks shared by the fused half-block kernels (wmsa.cu, mlp.cu): 64B-swizzled K-major operand tiles written by
// threads, TMEM loads that hand out MMA fragments, register re-balancing between warpgroups.
#pragma once
#include "tc05.cuh"

namespace fused {

// n / d for n * d < 2^32 with mul = ceil(2^32 / d) (d == 1: mul = 0 and the quotient is n)
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint32_t mul) { return mul ? __umulhi(n, mul) : n; }

__device__ __forceinline__ uint64_t desc_k_sw64(const void* smem) {
  // K-major operand, 64-byte rows, SWIZZLE_64B (layout type 4): 8-row groups 512 B apart (SBO), LBO unused
  const uint64_t addr = (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
  return addr | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// byte offset of 16-byte chunk c (0..3) of row r inside a [rows x 32] bf16 SWIZZLE_64B block
__device__ __forceinline__ int sw64_off(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }

// 16 TMEM lanes x 32 columns as m16n8 C fragments: v[4j + {0,1}] = (row g, cols 8j + 2t, +1), v[4j + {2,3}] = (row g + 8, same)
// Issue only; tmem_ld_wait() then tmem_ld_use() on every destination array before the values are read.
__device__ __forceinline__ void tmem_ld_frag32(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
        "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr));
}
// one accumulator row per thread, 8 consecutive columns; issue only
__device__ __forceinline__ void tmem_ld_row8(uint32_t taddr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_use8(float (&v)[8]) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// empty volatile statement that "rewrites" the loaded registers: volatile statements keep their order, so every use of
// v[] is scheduled after the wait (the compiler does not know tcgen05.wait::ld guards the registers of earlier loads)
__device__ __forceinline__ void tmem_ld_use(float (&v)[16]) {
  asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
               "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]));
}
// branch-free pick of one of four values by a per-lane index (selp; a ternary chain compiled to divergent branches)
__device__ __forceinline__ float sel4(int k, float a, float b, float c, float d) {
  float lo, hi, r;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f32 %0, %2, %1, p;\n\t}" : "=f"(lo) : "f"(a), "f"(b), "r"(k & 1));
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f32 %0, %2, %1, p;\n\t}" : "=f"(hi) : "f"(c), "f"(d), "r"(k & 1));
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %3, 0;\n\tselp.f32 %0, %2, %1, p;\n\t}" : "=f"(r) : "f"(lo), "f"(hi), "r"(k & 2));
  return r;
}
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ uint32_t movm_trans(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}


}  // namespace fused
