// TMEM (Blackwell tensor memory) used as a per-lane FP64 scratchpad.
//
// TMEM is 128 lanes x 512 32-bit columns per SM; warp w of a CTA can access lanes 32*(w%4)..+31 with the
// 32x32b shape: every thread reads/writes consecutive columns of its own lane.  A double occupies two
// columns (low word in the even column).  That gives each thread of the four "consumer" warps a private
// 256-double store with its own datapath (LDTM/STTM), which is where the adjoint sweep keeps its running
// sums W_j, Lambda, Mu — data that would otherwise pin 55 KB of shared memory per cell.
// tools/probes/tmem_probe.cu checks this usage on the target (correctness + cost).
#pragma once
#include <stdint.h>

namespace fatode {
namespace tmem {

__device__ __forceinline__ uint32_t alloc512(uint32_t* smem_slot) {   // executed by one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_slot)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  return 0;
}
__device__ __forceinline__ void dealloc512(uint32_t base) {           // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(base) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// one double at column `col` (even) of the calling thread's lane
__device__ __forceinline__ void st1(uint32_t taddr, double v) {
  uint32_t lo = (uint32_t)__double2loint(v), hi = (uint32_t)__double2hiint(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};\n" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double ld1(uint32_t taddr) {
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
  wait_ld();
  return __hiloint2double((int)hi, (int)lo);
}

// 16 consecutive doubles (32 columns) in one instruction; the caller issues wait_ld() before using them
__device__ __forceinline__ void ld16_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, double (&d)[16]) {
  uint32_t r[32];
  ld16_nowait(taddr, r);
  wait_ld();
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void st16(uint32_t taddr, const double (&d)[16]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) { r[2 * i] = (uint32_t)__double2loint(d[i]); r[2 * i + 1] = (uint32_t)__double2hiint(d[i]); }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// 32 consecutive doubles (64 columns) in one instruction; the caller issues wait_ld() before using them
__device__ __forceinline__ void ld32_nowait(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, double (&d)[32]) {
  uint32_t r[64];
  ld32_nowait(taddr, r);
  wait_ld();
#pragma unroll
  for (int i = 0; i < 32; ++i) d[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void st32(uint32_t taddr, const double (&d)[32]) {
  uint32_t r[64];
#pragma unroll
  for (int i = 0; i < 32; ++i) { r[2 * i] = (uint32_t)__double2loint(d[i]); r[2 * i + 1] = (uint32_t)__double2hiint(d[i]); }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x64.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, %64};\n"
               :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]), "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]), "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]), "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]), "r"(r[61]), "r"(r[62]), "r"(r[63]) : "memory");
}
}  // namespace tmem
}  // namespace fatode
