// tmem.cuh -- Blackwell tensor memory (TMEM) used as thread-private scratch:
// allocation by one warp (power-of-two columns), a warp reaches only the lanes
// of its quadrant 32*(warp%4), shape 32x32b = one 32-bit word per thread per
// column.  Shared by tmem_kernels.cu and cluster_kernels.cu.
#ifndef FFTWPP_TMEM_CUH
#define FFTWPP_TMEM_CUH

namespace fftwpp_gpu {
namespace {

// ---- tensor memory (TMEM) as thread-private scratch ----
__device__ __forceinline__ void tmemAlloc(unsigned *slot, int cols)
{
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"((unsigned) __cvta_generic_to_shared(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmemFree(unsigned taddr, int cols)
{
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(cols) : "memory");
}
// 4 complex doubles (16 x 32 bit) of this thread's lane at columns [taddr, +16)
__device__ __forceinline__ void tmemSt4(unsigned taddr, const double2 *v)
{
  unsigned r[16];
#pragma unroll
  for(int i=0; i < 4; ++i) {
    r[4*i]=__double2loint(v[i].x); r[4*i+1]=__double2hiint(v[i].x);
    r[4*i+2]=__double2loint(v[i].y); r[4*i+3]=__double2hiint(v[i].y);
  }
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),
                  "r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmemLd4(unsigned taddr, double2 *v)
{
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),
                 "=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for(int i=0; i < 4; ++i)
    v[i]=make_double2(__hiloint2double(r[4*i+1],r[4*i]),__hiloint2double(r[4*i+3],r[4*i+2]));
}
__device__ __forceinline__ void tmemWaitSt()
{
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

} // namespace
} // namespace fftwpp_gpu

#endif
