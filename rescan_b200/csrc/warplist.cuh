// Sorted list of 32*EPL 64-bit keys spread over the registers of one warp (element j lives in lane j / EPL, slot
// j % EPL): the k-best structure of the radius / k-NN searches (replacing the reference's per-query heap and final
// sort, msh_hash_grid.h:796-824, :579-703) and of the per-object top-k of the pose proposals.
#pragma once
#include "rsgpu_internal.cuh"

#ifdef __CUDACC__
constexpr unsigned long long KEY_INF = 0xffffffffffffffffull;

// sorted list of 32*EPL keys, element j lives in lane j / EPL, slot j % EPL
template <int EPL>
struct WarpList
{
  unsigned long long v[EPL];
  __device__ __forceinline__ void init()
  {
#pragma unroll
    for( int s = 0; s < EPL; ++s ) { v[s] = KEY_INF; }
  }
  // insert warp-uniform key x (drops the largest element)
  __device__ __forceinline__ void insert( unsigned long long x, int lane )
  {
    unsigned long long up = __shfl_up_sync( RS_FULL, v[EPL - 1], 1 );
    if( lane == 0 ) { up = 0ull; } // nothing precedes element 0: "previous <= x" always holds
#pragma unroll
    for( int s = EPL - 1; s >= 0; --s )
    {
      unsigned long long prev = s > 0 ? v[s - 1] : up;
      v[s] = ( v[s] <= x ) ? v[s] : ( prev <= x ? x : prev );
    }
  }
  __device__ __forceinline__ unsigned long long get( int j ) const
  {
    unsigned long long r = KEY_INF;
    int src = j / EPL, slot = j % EPL;
#pragma unroll
    for( int s = 0; s < EPL; ++s ) { if( s == slot ) { r = v[s]; } }
    return __shfl_sync( RS_FULL, r, src );
  }
};
#endif // __CUDACC__
