// Batched point-to-plane ICP: replaces icp_align (reference lib/rs/icp.h:416-500) with its per-iteration
// icp_find_corrs (:306-412) and icp_estimate_rigid_xform_pt2pl (:210-298, LDL^T from lineqn.h:153-218).
//
// Every iteration: (A) correspondences with the nearest-compatible-neighbour query (k = 16 rank rule, acosf gate folded
// into a dot threshold); (B) distance statistics for the 2.5 sigma rejection, weighted centroids and the 6x6 normal
// equations, accumulated SEQUENTIALLY IN FLOAT in the reference's point order (one accumulator per column walked down
// the rows of shared tiles; msh_std.h:1778-1824, icp.h:137-148, 226-252) - the outlier cut and the |d err| < 1e-5 stop
// are discontinuous in those sums, so only the reference's order reproduces its iteration counts; (C) the 6x6 system
// factored and solved in fp64 exactly like trimesh::ldltdc/ldltsl and the update composed in the reference's float
// msh_translate / msh_rotate / msh_mat4_mul order.  Refined poses, errors and iteration counts are bit-identical to the
// reference's (tests/test_gpu_parity.py, tests/test_gpu_golden.py).  `icp_sums=fp64` selects deterministic fp64
// warp-shuffle reductions instead (poses then agree to ~1e-3 m only; profiles/icp_parity_r01.md).
// The scan grid is built ONCE by the caller and shared by all poses: the search is exact, so the reference's per-call
// grid rebuild (:434-437) is not reproduced.  Batch structure: see icp_run below.
#include "rsgpu_internal.cuh"
#include "nearest.cuh"
#include "nearest_group.cuh"
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <mutex>

using namespace rs;

#ifndef RS_ICP_THREADS
#define RS_ICP_THREADS 256
#endif
namespace
{
// Block size of the ICP kernels (-DRS_ICP_THREADS=128 builds the lean variant: <= 20 KB of shared memory per block, which
// fits on an SM next to the three resident blocks of the dense pose search; measured on C2 it changes nothing at eight
// lanes (33.9 vs 33.4 ms per step) and costs the solve 10 % alone, because a tile of 64 rows pays the stage barrier three
// times as often - DESIGN.md 8b).
constexpr int ICP_THREADS = RS_ICP_THREADS;
constexpr int ICP_WARPS = ICP_THREADS / 32;
constexpr int ICP_MAX_NV = 32;
constexpr int ICP_G = 4;       // lanes per correspondence search (nearest_group.cuh)

struct IcpShared
{
  float T[16];        // current T1
  float M[16];        // T2i applied after T1 is done point by point, kept separately
  float max_dist, prev_err, err;
  int stop, steps;
  double red[ICP_WARPS][ICP_MAX_NV];
  double out[ICP_MAX_NV];
};

template <int NV>
__device__ __forceinline__ void block_reduce( double* v, IcpShared& sh )
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for( int i = 0; i < NV; ++i )
  {
    double x = v[i];
#pragma unroll
    for( int o = 16; o > 0; o >>= 1 ) { x += __shfl_down_sync( RS_FULL, x, o ); }
    if( lane == 0 ) { sh.red[warp][i] = x; }
  }
  __syncthreads();
  if( threadIdx.x < NV )
  {
    double s = 0.0;
    for( int w = 0; w < ICP_WARPS; ++w ) { s += sh.red[w][threadIdx.x]; }
    sh.out[threadIdx.x] = s;
  }
  __syncthreads();
}

// msh_translate (msh_vec_math.h:2064-2073): col3 = (col0*tx + col1*ty) + (col2*tz + col3)
__device__ void xf_translate( float* m, float tx, float ty, float tz )
{
  for( int r = 0; r < 4; ++r )
  {
    m[12 + r] = __fadd_rn( __fadd_rn( __fmul_rn( m[r], tx ), __fmul_rn( m[4 + r], ty ) ), __fadd_rn( __fmul_rn( m[8 + r], tz ), m[12 + r] ) );
  }
}

// msh_rotate (msh_vec_math.h:2089-2136) about a unit coordinate axis
__device__ void xf_rotate( float* m, float angle, float ax, float ay, float az )
{
  // cosf/sinf of the host libm are (almost always) correctly rounded; fp64 evaluation rounded to float matches them
  float c = (float)cos( (double)angle ), s = (float)sin( (double)angle ), t = __fsub_rn( 1.0f, c );
  float inv = __fdiv_rn( 1.0f, sqrtf( __fadd_rn( __fadd_rn( __fmul_rn( ax, ax ), __fmul_rn( ay, ay ) ), __fmul_rn( az, az ) ) ) );
  ax = __fmul_rn( ax, inv ); ay = __fmul_rn( ay, inv ); az = __fmul_rn( az, inv );
  float R[16];
  for( int i = 0; i < 16; ++i ) { R[i] = 0.f; }
  R[0] = __fadd_rn( c, __fmul_rn( __fmul_rn( ax, ax ), t ) );
  R[5] = __fadd_rn( c, __fmul_rn( __fmul_rn( ay, ay ), t ) );
  R[10] = __fadd_rn( c, __fmul_rn( __fmul_rn( az, az ), t ) );
  float a = __fmul_rn( __fmul_rn( ax, ay ), t ), b = __fmul_rn( az, s );
  R[1] = __fadd_rn( a, b ); R[4] = __fsub_rn( a, b );
  a = __fmul_rn( __fmul_rn( ax, az ), t ); b = __fmul_rn( ay, s );
  R[2] = __fsub_rn( a, b ); R[8] = __fadd_rn( a, b );
  a = __fmul_rn( __fmul_rn( ay, az ), t ); b = __fmul_rn( ax, s );
  R[6] = __fadd_rn( a, b ); R[9] = __fsub_rn( a, b );
  float o[16];
  for( int i = 0; i < 16; ++i ) { o[i] = m[i]; }
  for( int j = 0; j < 3; ++j )
    for( int r = 0; r < 4; ++r )
    {
      o[4 * j + r] = __fadd_rn( __fmul_rn( m[r], R[4 * j] ), __fadd_rn( __fmul_rn( m[4 + r], R[4 * j + 1] ), __fmul_rn( m[8 + r], R[4 * j + 2] ) ) );
    }
  for( int i = 0; i < 16; ++i ) { m[i] = o[i]; }
}

// msh_mat4_mul (msh_vec_math.h:1441-1480): o = a*b, each entry a 4-term left-to-right sum
__device__ void xf_mul( const float* a, const float* b, float* o )
{
  float t[16];
  for( int c = 0; c < 4; ++c )
    for( int r = 0; r < 4; ++r )
    {
      t[4 * c + r] = __fadd_rn( __fadd_rn( __fadd_rn( __fmul_rn( b[4 * c], a[r] ), __fmul_rn( b[4 * c + 1], a[4 + r] ) ),
                                           __fmul_rn( b[4 * c + 2], a[8 + r] ) ), __fmul_rn( b[4 * c + 3], a[12 + r] ) );
    }
  for( int i = 0; i < 16; ++i ) { o[i] = t[i]; }
}

// trimesh::ldltdc + ldltsl for N = 6 (lineqn.h:177-193, 206-217).  A zero pivot aborts the factorisation but
// the caller ignores that (icp.h:276-277) and still runs the sweeps with the remaining reciprocal pivots 0.
__device__ void ldlt_solve6( double A[6][6], const double* b, double* x )
{
  double rd[6] = { 0, 0, 0, 0, 0, 0 }, v[5];
  bool ok = true;
  for( int i = 0; i < 6 && ok; ++i )
  {
    for( int k = 0; k < i; ++k ) { v[k] = __dmul_rn( A[i][k], rd[k] ); }
    for( int j = i; j < 6; ++j )
    {
      double s = A[i][j];
      for( int k = 0; k < i; ++k ) { s = __dsub_rn( s, __dmul_rn( v[k], A[j][k] ) ); }
      if( i == j ) { if( s == 0 ) { ok = false; break; } rd[i] = __ddiv_rn( 1.0, s ); }
      else { A[j][i] = s; }
    }
  }
  for( int i = 0; i < 6; ++i )
  {
    double s = b[i];
    for( int k = 0; k < i; ++k ) { s = __dsub_rn( s, __dmul_rn( A[i][k], x[k] ) ); }
    x[i] = __dmul_rn( s, rd[i] );
  }
  for( int i = 5; i >= 0; --i )
  {
    double s = 0;
    for( int k = i + 1; k < 6; ++k ) { s = __dadd_rn( s, __dmul_rn( A[k][i], x[k] ) ); }
    x[i] = __dsub_rn( x[i], __dmul_rn( s, rd[i] ) );
  }
}

// Sums in the reference's own order.  The reference accumulates every sum of an ICP step sequentially in
// float over the correspondences in point order (msh_std.h:1778-1808, icp.h:137-148, 226-252); since the 2.5
// sigma cut and the stopping rule are discontinuous in those sums, matching its poses to 1e-5 needs the same
// rounding.  All threads form the per-point terms in parallel into a padded shared tile (one row per thread),
// then the 32 lanes of warp 0 each run ONE accumulator down the rows in point order: the dependent chain is a
// single FADD (and a DADD for the two fp64 sums) per row, loads are conflict-free.  Rows of points without a
// correspondence hold zeros, which leave a float sum unchanged.
constexpr int TILE_LD = 33;

// acc + col[0] + col[TILE_LD] + ... in row order, one dependent add per row; the loads of the next eight rows are
// issued before the eight adds of the current ones, so the chain never waits on shared memory
template <typename A>
__device__ __forceinline__ A chain_add( A acc, float x );
template <> __device__ __forceinline__ float chain_add<float>( float acc, float x ) { return __fadd_rn( acc, x ); }
template <> __device__ __forceinline__ double chain_add<double>( double acc, float x ) { return __dadd_rn( acc, (double)x ); }
template <typename A>
__device__ __forceinline__ A chain_sum( A acc, const float* __restrict__ col, int rows )
{
  float x[8], y[8];
  const int nchunk = rows >> 3;
  int c = 0;
  if( nchunk > 0 )
  {
#pragma unroll
    for( int j = 0; j < 8; ++j ) { x[j] = col[j * TILE_LD]; }
  }
  for( ; c + 2 <= nchunk; c += 2 )
  {
#pragma unroll
    for( int j = 0; j < 8; ++j ) { y[j] = col[( 8 * ( c + 1 ) + j ) * TILE_LD]; }
#pragma unroll
    for( int j = 0; j < 8; ++j ) { acc = chain_add<A>( acc, x[j] ); }
    if( c + 2 < nchunk )
    {
#pragma unroll
      for( int j = 0; j < 8; ++j ) { x[j] = col[( 8 * ( c + 2 ) + j ) * TILE_LD]; }
    }
#pragma unroll
    for( int j = 0; j < 8; ++j ) { acc = chain_add<A>( acc, y[j] ); }
  }
  if( c < nchunk ) // one chunk left, already in x
  {
#pragma unroll
    for( int j = 0; j < 8; ++j ) { acc = chain_add<A>( acc, x[j] ); }
    ++c;
  }
  for( int r = 8 * c; r < rows; ++r ) { acc = chain_add<A>( acc, col[r * TILE_LD] ); }
  return acc;
}

template <int NV, bool WITH_F64, class LoadFn, class TermFn>
__device__ __forceinline__ void ordered_sums( int n, LoadFn load_fn, TermFn term_fn, float* tile, float* fout, double* dout )
{
  // Three roles, pipelined over tiles of TR rows: warp 0 runs the float sums down the rows of tile s - 1, warp 1 (when
  // asked) the two fp64 sums of the error term (icp.h:250-253: columns NV-2, NV-1), and every other thread owns one
  // row per tile: it turns the inputs it loaded during the previous tile into the terms of tile s and issues the loads
  // of tile s + 1, so a whole tile time hides the memory latency.
  constexpr int FIRST_FILL = 2;                     // warp 1 is idle in the passes without fp64 sums (the tile size is fixed)
  constexpr int TR = ICP_THREADS - 32 * FIRST_FILL; // rows per tile = fill threads
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ft = tid - 32 * FIRST_FILL;
  float fa = 0.0f; double da = 0.0;
  const int stages = ( n + TR - 1 ) / TR;
  // one pipeline step: stage s of the fill (terms from `mine`, loads of stage s + 1 into `other`) beside the sums of stage s - 1.
  // Two input registers sets are used alternately (no copy between them: a copy would wait for the loads just issued).
  decltype( load_fn( 0 ) ) in_a = {}, in_b = {};
  auto step = [&]( int s, decltype( load_fn( 0 ) )& mine, decltype( load_fn( 0 ) )& other ) {
    if( ft >= 0 )
    {
      if( s < stages )
      {
        const int inext = ( s + 1 ) * TR + ft;
        if( inext < n ) { other = load_fn( inext ); }
        float* buf = tile + ( s & 1 ) * ( TR * TILE_LD );
        float t[NV];
#pragma unroll
        for( int v = 0; v < NV; ++v ) { t[v] = 0.0f; }
        if( s * TR + ft < n ) { term_fn( mine, t ); }
#pragma unroll
        for( int v = 0; v < NV; ++v ) { buf[ft * TILE_LD + v] = t[v]; }
      }
    }
    else if( s >= 1 )
    {
      const float* buf = tile + ( ( s - 1 ) & 1 ) * ( TR * TILE_LD );
      const int rows = min( TR, n - ( s - 1 ) * TR );
      if( warp == 0 )
      {
        if( lane < NV ) { fa = chain_sum<float>( fa, buf + lane, rows ); }
      }
      else if( WITH_F64 && lane < 2 ) { da = chain_sum<double>( da, buf + ( NV - 2 + lane ), rows ); } // warp 1
    }
    __syncthreads();
  };
  if( ft >= 0 && ft < n ) { in_a = load_fn( ft ); }
  for( int s = 0; s <= stages; s += 2 )
  {
    step( s, in_a, in_b );
    if( s + 1 <= stages ) { step( s + 1, in_b, in_a ); }
  }
  if( warp == 0 ) { fout[lane] = fa; }
  if( WITH_F64 && warp == 1 && lane < 2 ) { dout[NV - 2 + lane] = da; }
  __syncthreads();
}

// inputs of one correspondence as the sums need them (all direct loads from the per-alignment scratch)
struct CorrIn1 { uint32_t pos; float d; };
struct CorrIn2 { uint2 m; float4 q, p2; };
struct CorrIn3 { uint2 m; float4 q, p2, nn; };

// what one alignment needs: which object it aligns and where its scratch lives
struct IcpBlock
{
  const float* p1;
  const float* n1;
  int n;
  unsigned long long scratch_off;
};

// (A) correspondences of one batch of 32 object points (icp.h:339-391), searched 8 at a time by 4-lane groups
__device__ __forceinline__ void icp_correspond_batch( const GridView& g, const float* __restrict__ T, const float* __restrict__ M,
                                                      const float* __restrict__ p1, const float* __restrict__ n1, int c1n, int ib, int n_batch, bool warm,
                                                      double radius, float r2f, float dot_thr, float4* __restrict__ cq, uint2* __restrict__ cm,
                                                      float4* __restrict__ cp, float4* __restrict__ cn,
                                                      uint4* __restrict__ cand, unsigned char* __restrict__ slot )
{
  const int lane = threadIdx.x & 31;
  const int i = ib + lane;
  const bool valid = i < c1n && lane < n_batch;
  LaneQuery q;
  q.px = q.py = q.pz = q.nx = q.ny = q.nz = 0.f;
  if( valid )
  {
    float ax, ay, az, bx, by, bz;
    xf_apply( T, __ldg( p1 + 3 * (size_t)i ), __ldg( p1 + 3 * (size_t)i + 1 ), __ldg( p1 + 3 * (size_t)i + 2 ), 1.0f, ax, ay, az );
    xf_apply( T, __ldg( n1 + 3 * (size_t)i ), __ldg( n1 + 3 * (size_t)i + 1 ), __ldg( n1 + 3 * (size_t)i + 2 ), 0.0f, bx, by, bz );
    xf_apply( M, ax, ay, az, 1.0f, q.px, q.py, q.pz );
    xf_apply( M, bx, by, bz, 0.0f, q.nx, q.ny, q.nz );
  }
  const rsg::Stage1 s1 = rsg::stage1_test( g, radius, dot_thr, true, q.px, q.py, q.pz, q.nx, q.ny, q.nz, valid );
  const bool fastq = s1.active && s1.fast, slowq = s1.active && !s1.fast;
  // warm start: the previous iteration's correspondent, if it is still compatible and inside the (shrunken) radius,
  // bounds the search from its first cell on (the scratch starts out as "no correspondence")
  unsigned long long seedkey = ~0ull; float seeddot = 0.f;
  if( fastq && warm )
  {
    const uint32_t prev = __ldcg( &cm[i] ).x; // written by another block in the previous iteration: never through L1
    if( prev != 0xffffffffu )
    {
      const float4 rec = __ldg( g.recs + prev ), mm = __ldg( g.nrm + prev );
      const float d2 = dist2_exact( rec, q.px, q.py, q.pz );
      const float dot = dot3_exact( mm.x, mm.y, mm.z, q.nx, q.ny, q.nz );
      if( d2 < r2f && dot >= dot_thr && dot <= 1.0f ) { seedkey = ( (unsigned long long)__float_as_uint( d2 ) << 32 ) | prev; seeddot = dot; }
    }
  }
  const unsigned fastm = __ballot_sync( RS_FULL, fastq );
  const int rank = __popc( fastm & ( ( 1u << lane ) - 1u ) );
  if( fastq ) { slot[rank] = (unsigned char)lane; }
  __syncwarp();
  auto query_of = [&]( int r, float& px, float& py, float& pz, float& nx, float& ny, float& nz, unsigned long long& sk, float& sd ) -> bool {
    const int src = slot[r];
    px = __shfl_sync( RS_FULL, q.px, src ); py = __shfl_sync( RS_FULL, q.py, src ); pz = __shfl_sync( RS_FULL, q.pz, src );
    nx = __shfl_sync( RS_FULL, q.nx, src ); ny = __shfl_sync( RS_FULL, q.ny, src ); nz = __shfl_sync( RS_FULL, q.nz, src );
    sk = __shfl_sync( RS_FULL, seedkey, src ); sd = __shfl_sync( RS_FULL, seeddot, src );
    return true;
  };
  const NearestHit hr = rsg::group_round<ICP_G>( g, __popc( fastm ), query_of, radius, r2f, dot_thr, 16, cand );
  // lane L of the round holds the result of the query with rank L
  NearestHit h;
  h.d2 = __shfl_sync( RS_FULL, hr.d2, rank ); h.dot = __shfl_sync( RS_FULL, hr.dot, rank );
  h.pos = __shfl_sync( RS_FULL, hr.pos, rank ); h.found = __shfl_sync( RS_FULL, (int)hr.found, rank ) != 0 && fastq;
  if( __any_sync( RS_FULL, slowq ) )
  {
    NearestHit hs = nearest_compatible_batch<false>( g, q, slowq, radius, r2f, dot_thr, 16, nullptr );
    if( slowq ) { h = hs; }
  }
  if( valid )
  {
    cq[i] = make_float4( q.px, q.py, q.pz, h.d2 );
    float dot = h.dot > 0.0f ? h.dot : 0.0f;
    cm[i] = make_uint2( h.found ? h.pos : 0xffffffffu, __float_as_uint( dot ) );
    // the correspondent's point and normal travel with the correspondence: the sums then read four plain streams
    if( h.found ) { cp[i] = __ldg( g.recs + h.pos ); cn[i] = __ldg( g.nrm + h.pos ); }
  }
  __syncwarp();
}

// (B) + (C) of one iteration for the block's alignment: statistics, centroids, normal equations, solve, compose.
// Returns false when the reference leaves its loop before the update (no correspondences / no weight, icp.h:453-468);
// otherwise sh.T, sh.err, sh.steps, sh.max_dist are updated and sh.stop is set when the stopping rule fires.
template <bool EXACT>
__device__ __forceinline__ bool icp_update( const GridView& g, int c1n, const float4* __restrict__ cq, const uint2* __restrict__ cm,
                                            const float4* __restrict__ cp, const float4* __restrict__ cn, int it,
                                            IcpShared& sh, float* tile, float* fout, double* dout )
{
  const int tid = threadIdx.x;
  const float max_dist = sh.max_dist;
  // ---- (B1) statistics of the squared distances (icp.h:394-396; msh_std.h:1778-1824)
  int nc; float sum_d, sum_dd;
  if( EXACT )
  {
    int mine = 0;
    for( int i = tid; i < c1n; i += ICP_THREADS ) { mine += __ldcg( &cm[i] ).x != 0xffffffffu; }
    {
      double v[1] = { (double)mine };
      block_reduce<1>( v, sh );
      nc = (int)sh.out[0];
    }
    ordered_sums<2, false>( c1n,
      [&]( int i ) { CorrIn1 in; in.pos = __ldcg( &cm[i] ).x; in.d = __ldcg( &cq[i] ).w; return in; },
      [&]( const CorrIn1& in, float* t ) { if( in.pos == 0xffffffffu ) { return false; } t[0] = in.d; t[1] = __fmul_rn( in.d, in.d ); return true; },
      tile, fout, dout );
    sum_d = fout[0]; sum_dd = fout[1];
  }
  else
  {
    double v[3] = { 0, 0, 0 };
    for( int i = tid; i < c1n; i += ICP_THREADS )
    {
      if( __ldcg( &cm[i] ).x != 0xffffffffu ) { float d = __ldcg( &cq[i] ).w; v[0] += 1.0; v[1] += (double)d; v[2] += (double)__fmul_rn( d, d ); }
    }
    block_reduce<3>( v, sh );
    nc = (int)sh.out[0]; sum_d = (float)sh.out[1]; sum_dd = (float)sh.out[2];
  }
  if( nc == 0 ) { return false; } // (icp.h:453-457)
  const float mean = __fdiv_rn( sum_d, (float)nc );
  const float sd = (float)sqrt( (double)__fsub_rn( __fdiv_rn( sum_dd, (float)nc ), __fmul_rn( mean, mean ) ) );
  const bool reject = (double)sd > 0.000001;
  const float cut = __fmul_rn( 2.5f, sd );
  __syncthreads();
  // weight of correspondence i (icp.h:387, 397-402)
  auto weight = [&]( const float4& q, const uint2& m ) {
    float w = __fmul_rn( __fsub_rn( 1.0f, __fdiv_rn( q.w, max_dist ) ), __uint_as_float( m.y ) );
    if( reject && q.w > cut ) { w = 0.0f; }
    return w;
  };
  // ---- (B2) weighted centroids (icp.h:137-148)
  float s7[7];
  if( EXACT )
  {
    ordered_sums<7, false>( c1n,
      [&]( int i ) { CorrIn2 in; in.m = __ldcg( &cm[i] ); in.q = __ldcg( &cq[i] ); in.p2 = __ldcg( &cp[i] ); return in; },
      [&]( const CorrIn2& in, float* t ) {
        if( in.m.x == 0xffffffffu ) { return false; }
        const float w = weight( in.q, in.m );
        t[0] = w;
        t[1] = __fmul_rn( in.q.x, w ); t[2] = __fmul_rn( in.q.y, w ); t[3] = __fmul_rn( in.q.z, w );
        t[4] = __fmul_rn( in.p2.x, w ); t[5] = __fmul_rn( in.p2.y, w ); t[6] = __fmul_rn( in.p2.z, w );
        return true;
      }, tile, fout, dout );
#pragma unroll
    for( int j = 0; j < 7; ++j ) { s7[j] = fout[j]; }
  }
  else
  {
    double v[7] = { 0, 0, 0, 0, 0, 0, 0 };
    for( int i = tid; i < c1n; i += ICP_THREADS )
    {
      uint2 m = __ldcg( &cm[i] );
      if( m.x == 0xffffffffu ) { continue; }
      float4 q = __ldcg( &cq[i] );
      float w = weight( q, m );
      float4 p2 = __ldcg( &cp[i] );
      v[0] += (double)w;
      v[1] += (double)__fmul_rn( q.x, w ); v[2] += (double)__fmul_rn( q.y, w ); v[3] += (double)__fmul_rn( q.z, w );
      v[4] += (double)__fmul_rn( p2.x, w ); v[5] += (double)__fmul_rn( p2.y, w ); v[6] += (double)__fmul_rn( p2.z, w );
    }
    block_reduce<7>( v, sh );
#pragma unroll
    for( int j = 0; j < 7; ++j ) { s7[j] = (float)sh.out[j]; }
  }
  const float tw = s7[0];
  if( (double)tw <= 1e-7 ) { return false; } // (icp.h:459-468)
  const float itw = __fdiv_rn( 1.0f, tw ); // msh_vec3_scalar_div multiplies by the reciprocal (msh_vec_math.h:754-758)
  const float c1x = __fmul_rn( s7[1], itw ), c1y = __fmul_rn( s7[2], itw ), c1z = __fmul_rn( s7[3], itw );
  const float c2x = __fmul_rn( s7[4], itw ), c2y = __fmul_rn( s7[5], itw ), c2z = __fmul_rn( s7[6], itw );
  __syncthreads();
  // ---- (B3) normal equations (icp.h:226-252): TL = sum w c c^T, TR = sum w c n^T, BR = sum w n n^T, b = sum w [c;n] (d.n)
  // 29 terms per correspondence: TL (6 unique), TR (9), BR (6 unique), b (6), w (d.n)^2, w
  auto load29 = [&]( int i ) { CorrIn3 in; in.m = __ldcg( &cm[i] ); in.q = __ldcg( &cq[i] ); in.p2 = __ldcg( &cp[i] ); in.nn = __ldcg( &cn[i] ); return in; };
  auto terms29 = [&]( const CorrIn3& in, float* t ) -> bool {
    const uint2 m = in.m;
    if( m.x == 0xffffffffu ) { return false; }
    const float4 q4 = in.q;
    float w = weight( q4, m );
    const float4 p2 = in.p2, nn = in.nn;
    float px = __fsub_rn( q4.x, c1x ), py = __fsub_rn( q4.y, c1y ), pz = __fsub_rn( q4.z, c1z );
    float qx = __fsub_rn( p2.x, c2x ), qy = __fsub_rn( p2.y, c2y ), qz = __fsub_rn( p2.z, c2z );
    float dx = __fsub_rn( px, qx ), dy = __fsub_rn( py, qy ), dz = __fsub_rn( pz, qz );
    float c[3], n[3] = { nn.x, nn.y, nn.z };
    c[0] = __fsub_rn( __fmul_rn( py, nn.z ), __fmul_rn( pz, nn.y ) );
    c[1] = __fsub_rn( __fmul_rn( pz, nn.x ), __fmul_rn( px, nn.z ) );
    c[2] = __fsub_rn( __fmul_rn( px, nn.y ), __fmul_rn( py, nn.x ) );
    float dn = dot3_exact( dx, dy, dz, nn.x, nn.y, nn.z );
    int o = 0;
#pragma unroll
    for( int col = 0; col < 3; ++col )
#pragma unroll
      for( int row = col; row < 3; ++row ) { t[o++] = __fmul_rn( __fmul_rn( c[row], c[col] ), w ); }
#pragma unroll
    for( int col = 0; col < 3; ++col )
#pragma unroll
      for( int row = 0; row < 3; ++row ) { t[o++] = __fmul_rn( __fmul_rn( c[row], n[col] ), w ); }
#pragma unroll
    for( int col = 0; col < 3; ++col )
#pragma unroll
      for( int row = col; row < 3; ++row ) { t[o++] = __fmul_rn( __fmul_rn( n[row], n[col] ), w ); }
#pragma unroll
    for( int a = 0; a < 3; ++a ) { t[21 + a] = __fmul_rn( __fmul_rn( w, c[a] ), dn ); t[24 + a] = __fmul_rn( __fmul_rn( w, n[a] ), dn ); }
    t[27] = __fmul_rn( __fmul_rn( w, dn ), dn );
    t[28] = w;
    return true;
  };
  if( EXACT )
  {
    ordered_sums<29, true>( c1n, load29, [&]( const CorrIn3& in, float* t ) { return terms29( in, t ); }, tile, fout, dout );
    if( tid < 27 ) { sh.out[tid] = (double)fout[tid]; }
    if( tid == 27 || tid == 28 ) { sh.out[tid] = dout[tid]; }
    __syncthreads();
  }
  else
  {
    double v[29];
#pragma unroll
    for( int j = 0; j < 29; ++j ) { v[j] = 0.0; }
    for( int i = tid; i < c1n; i += ICP_THREADS )
    {
      float t[29];
      if( terms29( load29( i ), t ) )
      {
#pragma unroll
        for( int j = 0; j < 29; ++j ) { v[j] += (double)t[j]; }
      }
    }
    block_reduce<29>( v, sh );
  }
  // ---- (C) solve + compose (icp.h:253-295), one thread
  if( tid == 0 )
  {
    const double* s = sh.out;
    float TL[3][3], TR[3][3], BR[3][3]; // [col][row]
    int o = 0;
    for( int col = 0; col < 3; ++col ) for( int row = col; row < 3; ++row ) { TL[col][row] = TL[row][col] = (float)s[o++]; }
    for( int col = 0; col < 3; ++col ) for( int row = 0; row < 3; ++row ) { TR[col][row] = (float)s[o++]; }
    for( int col = 0; col < 3; ++col ) for( int row = col; row < 3; ++row ) { BR[col][row] = BR[row][col] = (float)s[o++]; }
    float err = (float)sqrt( __ddiv_rn( s[27], s[28] ) );
    double A[6][6], rhs[6], x[6] = { 0, 0, 0, 0, 0, 0 };
    for( int r = 0; r < 3; ++r )
      for( int c = 0; c < 3; ++c )
      {
        A[r][c] = TL[c][r]; A[r][3 + c] = TR[c][r];
        A[3 + r][c] = TR[r][c]; A[3 + r][3 + c] = BR[c][r];
      }
    for( int a = 0; a < 6; ++a ) { rhs[a] = -(double)(float)s[21 + a]; }
    ldlt_solve6( A, rhs, x );
    float T[16];
    for( int i = 0; i < 16; ++i ) { T[i] = ( i % 5 == 0 ) ? 1.0f : 0.0f; }
    xf_translate( T, c1x, c1y, c1z );
    xf_translate( T, (float)x[3], (float)x[4], (float)x[5] );
    xf_rotate( T, (float)x[0], 1.0f, 0.0f, 0.0f );
    xf_rotate( T, (float)x[1], 0.0f, 1.0f, 0.0f );
    xf_rotate( T, (float)x[2], 0.0f, 0.0f, 1.0f );
    xf_translate( T, -c1x, -c1y, -c1z );
    float Tn[16];
    xf_mul( T, sh.T, Tn );
    for( int i = 0; i < 16; ++i ) { sh.T[i] = Tn[i]; }
    sh.err = err; sh.steps += 1;
    float delta = fabsf( __fsub_rn( sh.prev_err, err ) );
    if( it > 5 && (double)delta < 1e-5 ) { sh.stop = 1; }
    double shrunk = __dmul_rn( (double)max_dist, 0.95 );
    sh.max_dist = (float)( shrunk > 0.05 ? shrunk : 0.05 );
  }
  __syncthreads();
  return true;
}

// ---- variant 1: one thread block per alignment, resident for all its iterations (RSGPU_ICP_IMPL=block)
template <bool EXACT>
__global__ void __launch_bounds__( ICP_THREADS ) icp_kernel( GridView g, const IcpBlock* __restrict__ blocks,
                                                             float* __restrict__ T1_io, const float* __restrict__ T2i, float max_dist0,
                                                             float dot_thr, int max_iter, float4* __restrict__ scratch_q,
                                                             uint2* __restrict__ scratch_m, float4* __restrict__ scratch_p,
                                                             float4* __restrict__ scratch_n, float* __restrict__ errs, int* __restrict__ iters )
{
  const IcpBlock blk = blocks[blockIdx.x];
  __shared__ IcpShared sh;
  extern __shared__ float tile[]; // EXACT: 2 * ( ICP_THREADS - 64 ) * TILE_LD floats
  __shared__ float fout[32];
  __shared__ uint4 s_cand[ICP_WARPS][rsg::GroupCfg<ICP_G>::CAND_WORDS];
  __shared__ unsigned char s_slot[ICP_WARPS][32];
  __shared__ double dout[32];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5;
  float4* cq = scratch_q + blk.scratch_off;  // {q, d2}
  uint2* cm = scratch_m + blk.scratch_off;   // {recs position or ~0, dot bits}
  float4* cp = scratch_p + blk.scratch_off;  // the correspondent's point ...
  float4* cn = scratch_n + blk.scratch_off;  // ... and normal
  if( tid < 16 ) { sh.T[tid] = T1_io[16 * (size_t)b + tid]; sh.M[tid] = T2i[tid]; }
  if( tid == 0 ) { sh.max_dist = max_dist0; sh.prev_err = 1e6f; sh.err = 1e6f; sh.stop = 0; sh.steps = 0; }
  __syncthreads();
  for( int it = 0; it < max_iter; ++it )
  {
    if( tid == 0 ) { sh.prev_err = sh.err; }
    const double radius = (double)sh.max_dist;
    const float r2f = (float)__dmul_rn( radius, radius );
    for( int ib = warp * 32; ib < blk.n; ib += ICP_WARPS * 32 )
    {
      icp_correspond_batch( g, sh.T, sh.M, blk.p1, blk.n1, blk.n, ib, 32, it > 0, radius, r2f, dot_thr, cq, cm, cp, cn, s_cand[warp], s_slot[warp] );
    }
    __syncthreads();
    if( !icp_update<EXACT>( g, blk.n, cq, cm, cp, cn, it, sh, tile, fout, dout ) ) { break; }
    if( sh.stop ) { break; }
  }
  __syncthreads();
  if( tid < 16 ) { T1_io[16 * (size_t)b + tid] = sh.T[tid]; }
  if( tid == 0 ) { errs[b] = sh.err; if( iters ) { iters[b] = sh.steps; } }
}

// ---- variant 2 (default): iteration-synchronous over the whole batch.  Alignments differ 20x in work (points x
// iterations), so resident blocks leave most of the GPU idle behind the slowest one.  Here every iteration is two
// launches: icp_search_kernel spreads the correspondence searches of ALL still-running alignments over the whole
// GPU (one warp task = 32 consecutive points of one alignment), icp_solve_kernel runs (B) + (C) with one block per
// running alignment.  State lives in global memory between launches.
struct IcpState
{
  float T[16];
  float max_dist, prev_err, err;
  int active, steps;
};

__global__ void __launch_bounds__( ICP_THREADS ) icp_search_kernel( GridView g, const IcpBlock* __restrict__ blocks, const IcpState* __restrict__ state,
                                                                    const int* __restrict__ ids /* alignments of this partition */,
                                                                    const unsigned* __restrict__ task_start /* n_align + 1 */, int n_align, int pts_per_task,
                                                                    const float* __restrict__ T2i, float dot_thr, float4* __restrict__ scratch_q,
                                                                    uint2* __restrict__ scratch_m, float4* __restrict__ scratch_p,
                                                                    float4* __restrict__ scratch_n )
{
  __shared__ uint4 s_cand[ICP_WARPS][rsg::GroupCfg<ICP_G>::CAND_WORDS];
  __shared__ unsigned char s_slot[ICP_WARPS][32];
  __shared__ float s_T[ICP_WARPS][16], s_M[16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if( threadIdx.x < 16 ) { s_M[threadIdx.x] = T2i[threadIdx.x]; }
  __syncthreads();
  const unsigned n_tasks = task_start[n_align];
  for( unsigned task = blockIdx.x * ICP_WARPS + warp; task < n_tasks; task += gridDim.x * ICP_WARPS )
  {
    // alignment of this task: last a with task_start[a] <= task
    int lo = 0, hi = n_align;
    while( hi - lo > 1 ) { int mid = ( lo + hi ) >> 1; if( __ldg( task_start + mid ) <= task ) { lo = mid; } else { hi = mid; } }
    const int a = ids[lo];
    if( !state[a].active ) { continue; }
    const IcpBlock blk = blocks[a];
    if( lane < 16 ) { s_T[warp][lane] = state[a].T[lane]; }
    __syncwarp();
    const double radius = (double)state[a].max_dist;
    const float r2f = (float)__dmul_rn( radius, radius );
    icp_correspond_batch( g, s_T[warp], s_M, blk.p1, blk.n1, blk.n, (int)( task - __ldg( task_start + lo ) ) * pts_per_task, pts_per_task, state[a].steps > 0, radius, r2f, dot_thr,
                          scratch_q + blk.scratch_off, scratch_m + blk.scratch_off, scratch_p + blk.scratch_off, scratch_n + blk.scratch_off,
                          s_cand[warp], s_slot[warp] );
  }
}

template <bool EXACT>
__global__ void __launch_bounds__( ICP_THREADS ) icp_solve_kernel( GridView g, const IcpBlock* __restrict__ blocks, IcpState* __restrict__ state,
                                                                   const int* __restrict__ ids, int it, const float4* __restrict__ scratch_q,
                                                                   const uint2* __restrict__ scratch_m, const float4* __restrict__ scratch_p,
                                                                   const float4* __restrict__ scratch_n, int* __restrict__ n_active )
{
  const int a = ids[blockIdx.x], tid = threadIdx.x;
  if( !state[a].active ) { return; }
  const IcpBlock blk = blocks[a];
  __shared__ IcpShared sh;
  extern __shared__ float tile[];
  __shared__ float fout[32];
  __shared__ double dout[32];
  if( tid < 16 ) { sh.T[tid] = state[a].T[tid]; }
  if( tid == 0 ) { sh.max_dist = state[a].max_dist; sh.prev_err = state[a].err; sh.err = state[a].err; sh.stop = 0; sh.steps = state[a].steps; }
  __syncthreads();
  const bool updated = icp_update<EXACT>( g, blk.n, scratch_q + blk.scratch_off, scratch_m + blk.scratch_off, scratch_p + blk.scratch_off,
                                              scratch_n + blk.scratch_off, it, sh, tile, fout, dout );
  __syncthreads();
  if( tid < 16 ) { state[a].T[tid] = sh.T[tid]; }
  if( tid == 0 )
  {
    const int active = ( updated && !sh.stop ) ? 1 : 0;
    state[a].max_dist = sh.max_dist; state[a].prev_err = sh.prev_err; state[a].err = sh.err; state[a].steps = sh.steps;
    state[a].active = active;
    if( active ) { atomicAdd( n_active, 1 ); }
  }
}

// ---- the split variant as a CUDA graph: the two launches of an iteration are the same every time once their arguments live
// in device memory (a descriptor per lane and partition; the iteration number the stopping rule needs is the alignment's
// own step count), so four iterations - eight kernel nodes - are captured once per lane and partition and replayed with ONE
// driver call.  With eight objects' chains in flight the host threads issue ~1 300 ICP launches per C2 step; funnelled
// through the driver they, not the device, set the pace of the chains (eight batches side by side took 8-16 ms each, 3.6 ms
// alone).  Grids are fixed at capture (grid-stride loops inside), finished alignments are counted in the descriptor.
struct IcpDesc
{
  GridView g;
  const IcpBlock* blocks; IcpState* state; const int* ids; const unsigned* task_start;
  int n_align, pts_per_task, max_iter;
  const float* T2i; float dot_thr;
  float4* sq; uint2* sm; float4* sp; float4* sn;
  int* n_done;
};

__global__ void __launch_bounds__( ICP_THREADS ) icp_search_desc_kernel( const IcpDesc* __restrict__ dp )
{
  __shared__ uint4 s_cand[ICP_WARPS][rsg::GroupCfg<ICP_G>::CAND_WORDS];
  __shared__ unsigned char s_slot[ICP_WARPS][32];
  __shared__ float s_T[ICP_WARPS][16], s_M[16];
  __shared__ IcpDesc d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for( int i = threadIdx.x; i < (int)( sizeof( IcpDesc ) / 4 ); i += ICP_THREADS ) { ( (uint32_t*)&d )[i] = ( (const uint32_t*)dp )[i]; }
  __syncthreads();
  if( threadIdx.x < 16 ) { s_M[threadIdx.x] = d.T2i[threadIdx.x]; }
  __syncthreads();
  const unsigned n_tasks = d.task_start[d.n_align];
  for( unsigned task = blockIdx.x * ICP_WARPS + warp; task < n_tasks; task += gridDim.x * ICP_WARPS )
  {
    int lo = 0, hi = d.n_align;
    while( hi - lo > 1 ) { int mid = ( lo + hi ) >> 1; if( __ldg( d.task_start + mid ) <= task ) { lo = mid; } else { hi = mid; } }
    const int a = d.ids[lo];
    if( !d.state[a].active ) { continue; }
    const IcpBlock blk = d.blocks[a];
    if( lane < 16 ) { s_T[warp][lane] = d.state[a].T[lane]; }
    __syncwarp();
    const double radius = (double)d.state[a].max_dist;
    const float r2f = (float)__dmul_rn( radius, radius );
    icp_correspond_batch( d.g, s_T[warp], s_M, blk.p1, blk.n1, blk.n, (int)( task - __ldg( d.task_start + lo ) ) * d.pts_per_task, d.pts_per_task,
                          d.state[a].steps > 0, radius, r2f, d.dot_thr, d.sq + blk.scratch_off, d.sm + blk.scratch_off, d.sp + blk.scratch_off,
                          d.sn + blk.scratch_off, s_cand[warp], s_slot[warp] );
  }
}

template <bool EXACT>
__global__ void __launch_bounds__( ICP_THREADS ) icp_solve_desc_kernel( const IcpDesc* __restrict__ dp )
{
  __shared__ IcpShared sh;
  extern __shared__ float tile[];
  __shared__ float fout[32];
  __shared__ double dout[32];
  __shared__ IcpDesc d;
  const int tid = threadIdx.x;
  for( int i = tid; i < (int)( sizeof( IcpDesc ) / 4 ); i += ICP_THREADS ) { ( (uint32_t*)&d )[i] = ( (const uint32_t*)dp )[i]; }
  __syncthreads();
  for( int bi = blockIdx.x; bi < d.n_align; bi += gridDim.x )
  {
    const int a = d.ids[bi];
    if( !d.state[a].active ) { continue; } // block-uniform
    const IcpBlock blk = d.blocks[a];
    const int it = d.state[a].steps; // a running alignment has made one step per iteration
    if( tid < 16 ) { sh.T[tid] = d.state[a].T[tid]; }
    if( tid == 0 ) { sh.max_dist = d.state[a].max_dist; sh.prev_err = d.state[a].err; sh.err = d.state[a].err; sh.stop = 0; sh.steps = it; }
    __syncthreads();
    const bool updated = icp_update<EXACT>( d.g, blk.n, d.sq + blk.scratch_off, d.sm + blk.scratch_off, d.sp + blk.scratch_off, d.sn + blk.scratch_off,
                                                it, sh, tile, fout, dout );
    __syncthreads();
    if( tid < 16 ) { d.state[a].T[tid] = sh.T[tid]; }
    if( tid == 0 )
    {
      const int active = ( updated && !sh.stop && it + 1 < d.max_iter ) ? 1 : 0;
      d.state[a].max_dist = sh.max_dist; d.state[a].prev_err = sh.prev_err; d.state[a].err = sh.err; d.state[a].steps = sh.steps;
      d.state[a].active = active;
      if( !active ) { atomicAdd( d.n_done, 1 ); }
    }
    __syncthreads();
  }
}

// per calling thread (= lane): the device descriptors and the instantiated graphs of its partitions
struct IcpGraphCache
{
  IcpDesc* d_desc[4] = { nullptr, nullptr, nullptr, nullptr };
  IcpDesc* h_desc = nullptr;   // pinned staging, 4 descriptors
  int* d_done[4] = { nullptr, nullptr, nullptr, nullptr };
  int* h_done = nullptr;       // pinned, 4 counters
  cudaGraphExec_t exec[4][2] = { { nullptr, nullptr }, { nullptr, nullptr }, { nullptr, nullptr }, { nullptr, nullptr } };
};
constexpr int ICP_GRAPH_ITERS = 4;

// ---- variant 3 (default): ONE persistent launch per batch, driven by a device-side work queue.  The iteration-
// synchronous split needs two launches per iteration and a host look at the running count every four: with eight objects'
// batches in flight that is ~1 300 launches per C2 step funnelled through the driver from eight host threads, and every
// one of them queues for an SM slot behind the dense search's resident blocks - the eight ICP batches of a step then take
// as long side by side as one after the other (measured: 12-16 ms each in flight together, 3 ms alone).  Here nothing
// returns to the host between the first search and the last solve:
//   * a work item = (alignment, chunk of 8 x pts_per_task consecutive object points); the queue is a plain array that is
//     only ever appended to (capacity = every chunk of every possible iteration), `tail` reserves, `head` hands out
//     tickets, a consumer waits until ITS slot is published;
//   * a block takes an item and its eight warps search 8 x pts_per_task correspondences (icp_correspond_batch);
//   * the block that completes an alignment's last chunk of the iteration (per-alignment arrival counter) runs the
//     reference-order sums, the 6x6 solve and the update for it (icp_update), and - unless the stopping rule fired -
//     appends the alignment's chunks for the next iteration; alignments advance independently of one another;
//   * the block that retires the last alignment appends one TERMINATE item per block of the grid.
// No block ever waits for a block that is not running (an item exists only after it has been published, and whoever holds
// one is resident), so the grid needs no co-residency and can be smaller or larger than the machine.  Per-alignment
// scratch written by one block and read by another goes through L2 only (__ldcg / volatile; plain stores are write-through).
struct IcpQueue
{
  unsigned head, tail, remaining, pad;
};
constexpr unsigned ICPQ_EMPTY = 0xffffffffu, ICPQ_TERMINATE = 0xfffffffeu;
constexpr int ICPQ_CHUNK_BITS = 13; // chunks per alignment < 8192 (1 M points at 16 per task x 8 warps), alignments < 2^19 - 1

template <bool EXACT>
__global__ void __launch_bounds__( ICP_THREADS, 512 / ICP_THREADS ) icp_persistent_kernel( GridView g, const IcpBlock* __restrict__ blocks, IcpState* state,
                                                                         const unsigned* __restrict__ n_chunks, unsigned* arrived, IcpQueue* q,
                                                                         unsigned* items, int pts_per_task, const float* __restrict__ T2i, float dot_thr,
                                                                         int max_iter, float4* scratch_q, uint2* scratch_m, float4* scratch_p, float4* scratch_n )
{
  extern __shared__ __align__( 16 ) float tile[]; // solve: two tiles of the ordered sums; search: the warps' cell tables and slots
  __shared__ IcpShared sh;
  __shared__ float fout[32];
  __shared__ double dout[32];
  __shared__ float s_T[16], s_M[16];
  __shared__ unsigned s_item, s_last, s_base;
  uint4* s_cand = (uint4*)tile;                                                                          // [ICP_WARPS][CAND_WORDS]
  unsigned char* s_slot = (unsigned char*)( s_cand + ICP_WARPS * rsg::GroupCfg<ICP_G>::CAND_WORDS );      // [ICP_WARPS][32]
  const int tid = threadIdx.x, warp = tid >> 5;
  if( tid < 16 ) { s_M[tid] = T2i[tid]; }
  volatile unsigned* vitems = items;
  for( ;; )
  {
    if( tid == 0 )
    {
      const unsigned ticket = atomicAdd( &q->head, 1u );
      unsigned it;
      while( ( it = vitems[ticket] ) == ICPQ_EMPTY ) { __nanosleep( 100 ); }
      __threadfence(); // the state the publisher wrote before the item
      s_item = it;
    }
    __syncthreads();
    const unsigned item = s_item;
    if( item == ICPQ_TERMINATE ) { break; }
    const int a = (int)( item >> ICPQ_CHUNK_BITS ), chunk = (int)( item & ( ( 1u << ICPQ_CHUNK_BITS ) - 1u ) );
    const IcpBlock blk = blocks[a];
    const volatile IcpState* vs = state + a;
    if( tid < 16 ) { s_T[tid] = vs->T[tid]; }
    const float cur_max_dist = vs->max_dist;
    const int cur_steps = vs->steps;
    __syncthreads();
    // ---- (A) this chunk's correspondences
    {
      const double radius = (double)cur_max_dist;
      const float r2f = (float)__dmul_rn( radius, radius );
      const int ib = ( chunk * ICP_WARPS + warp ) * pts_per_task;
      if( ib < blk.n )
      {
        icp_correspond_batch( g, s_T, s_M, blk.p1, blk.n1, blk.n, ib, pts_per_task, cur_steps > 0, radius, r2f, dot_thr,
                              scratch_q + blk.scratch_off, scratch_m + blk.scratch_off, scratch_p + blk.scratch_off, scratch_n + blk.scratch_off,
                              s_cand + warp * rsg::GroupCfg<ICP_G>::CAND_WORDS, s_slot + warp * 32 );
      }
    }
    __threadfence(); // every thread's correspondences are in L2 before the arrival is counted
    __syncthreads();
    if( tid == 0 )
    {
      const unsigned old = atomicAdd( arrived + a, 1u );
      s_last = old + 1u == n_chunks[a] ? 1u : 0u;
      if( s_last ) { arrived[a] = 0u; }
      __threadfence(); // the other blocks' correspondences, counted before ours
    }
    __syncthreads();
    if( !s_last ) { continue; }
    // ---- (B) + (C): this block completed the alignment's iteration
    if( tid < 16 ) { sh.T[tid] = vs->T[tid]; }
    if( tid == 0 ) { sh.max_dist = cur_max_dist; sh.prev_err = vs->err; sh.err = vs->err; sh.stop = 0; sh.steps = cur_steps; }
    __syncthreads();
    const bool updated = icp_update<EXACT>( g, blk.n, scratch_q + blk.scratch_off, scratch_m + blk.scratch_off, scratch_p + blk.scratch_off,
                                                scratch_n + blk.scratch_off, cur_steps, sh, tile, fout, dout );
    __syncthreads();
    const bool more = updated && !sh.stop && cur_steps + 1 < max_iter;
    if( tid < 16 ) { state[a].T[tid] = sh.T[tid]; }
    if( tid == 0 )
    {
      state[a].max_dist = sh.max_dist; state[a].prev_err = sh.prev_err; state[a].err = sh.err; state[a].steps = sh.steps;
      state[a].active = more ? 1 : 0;
    }
    __threadfence(); // the new state before the items that announce it
    __syncthreads();
    if( more )
    {
      const unsigned nch = n_chunks[a];
      if( tid == 0 ) { s_base = atomicAdd( &q->tail, nch ); }
      __syncthreads();
      for( unsigned c = tid; c < nch; c += ICP_THREADS ) { vitems[s_base + c] = ( (unsigned)a << ICPQ_CHUNK_BITS ) | c; }
    }
    else if( tid == 0 )
    {
      if( atomicSub( &q->remaining, 1u ) == 1u )
      {
        const unsigned base = atomicAdd( &q->tail, gridDim.x );
        for( unsigned c = 0; c < gridDim.x; ++c ) { vitems[base + c] = ICPQ_TERMINATE; }
      }
    }
    __syncthreads();
  }
}

// msh_mat4_inverse (msh_vec_math.h:1818-1917): cofactor expansion over 2x2 minors, all float, each cofactor
// a three-term expression evaluated left to right, scaled by 1.0f/det
float tri( float a, float x, float b, float y, float c, float z, int s2, int s3 )
{
  volatile float r = a * x;
  volatile float t = b * y;
  r = s2 > 0 ? r + t : r - t;
  t = c * z;
  r = s3 > 0 ? r + t : r - t;
  return r;
}
float det2( float a, float b, float c, float d )
{
  volatile float x = a * b, y = c * d;
  volatile float r = x - y;
  return r;
}
} // namespace

namespace rs
{
void mat4_inverse_ref( const float* m, float* o )
{
  float C[16], d[6];
  d[0] = det2( m[10], m[15], m[14], m[11] ); d[1] = det2( m[6], m[11], m[10], m[7] ); d[2] = det2( m[2], m[7], m[6], m[3] );
  d[3] = det2( m[6], m[15], m[14], m[7] );   d[4] = det2( m[2], m[11], m[10], m[3] ); d[5] = det2( m[2], m[15], m[14], m[3] );
  C[0] = tri( m[5], d[0], m[9], d[3], m[13], d[1], -1, +1 );
  C[1] = tri( m[9], d[5], m[1], d[0], m[13], d[4], -1, -1 );
  C[2] = tri( m[1], d[3], m[5], d[5], m[13], d[2], -1, +1 );
  C[3] = tri( m[5], d[4], m[9], d[2], m[1], d[1], -1, -1 );
  C[4] = tri( m[8], d[3], m[4], d[0], m[12], d[1], -1, -1 );
  C[5] = tri( m[0], d[0], m[8], d[5], m[12], d[4], -1, +1 );
  C[6] = tri( m[4], d[5], m[0], d[3], m[12], d[2], -1, -1 );
  C[7] = tri( m[0], d[1], m[4], d[4], m[8], d[2], -1, +1 );
  d[0] = det2( m[8], m[13], m[12], m[9] ); d[1] = det2( m[4], m[9], m[8], m[5] );  d[2] = det2( m[0], m[5], m[4], m[1] );
  d[3] = det2( m[4], m[13], m[12], m[5] ); d[4] = det2( m[0], m[9], m[8], m[1] );  d[5] = det2( m[0], m[13], m[12], m[1] );
  C[8]  = tri( m[7], d[0], m[11], d[3], m[15], d[1], -1, +1 );
  C[9]  = tri( m[11], d[5], m[3], d[0], m[15], d[4], -1, -1 );
  C[10] = tri( m[3], d[3], m[7], d[5], m[15], d[2], -1, +1 );
  C[11] = tri( m[7], d[4], m[3], d[1], m[11], d[2], -1, -1 );
  C[12] = tri( m[10], d[3], m[6], d[0], m[14], d[1], -1, -1 );
  C[13] = tri( m[2], d[0], m[10], d[5], m[14], d[4], -1, +1 );
  C[14] = tri( m[6], d[5], m[2], d[3], m[14], d[2], -1, -1 );
  C[15] = tri( m[2], d[1], m[6], d[4], m[10], d[2], -1, +1 );
  volatile float det = m[0] * C[0];
  volatile float t = m[4] * C[1]; det = det + t;
  t = m[8] * C[2]; det = det + t;
  t = m[12] * C[3]; det = det + t;
  float s = 1.0f / det;
  for( int i = 0; i < 16; ++i ) { o[i] = s * C[i]; }
}

// smallest float dot with acosf(dot) < max_angle (icp.h:373-374); -inf when even dot = 0 passes, because the
// reference clamps negative dots to 0 before the test
float compat_threshold_acosf( float max_angle )
{
  auto ok = [&]( float d ) { return acosf( d ) < max_angle; };
  if( ok( 0.0f ) ) { return -INFINITY; }
  if( !ok( 1.0f ) ) { return 2.0f; }
  uint32_t lo, hi; float f0 = 0.0f, f1 = 1.0f;
  memcpy( &lo, &f0, 4 ); memcpy( &hi, &f1, 4 );
  while( hi - lo > 1 )
  {
    uint32_t mid = lo + ( hi - lo ) / 2; float fm; memcpy( &fm, &mid, 4 );
    if( ok( fm ) ) { hi = mid; } else { lo = mid; }
  }
  float out; memcpy( &out, &hi, 4 );
  return out;
}
} // namespace rs

namespace
{
int icp_run( const rsgpu_icp_job_t* jobs, int32_t n_jobs, const rsgpu_grid_t* scan, const float* T2, float max_dist, float max_angle,
             int32_t max_iter )
{
  if( n_jobs < 0 || ( n_jobs > 0 && !jobs ) || !scan ) { return fail( RSGPU_ERR_INVALID, "rsgpu_icp_align: bad argument" ); }
  RS_TRY( ensure_device() );
  size_t total = 0, scratch = 0;
  for( int j = 0; j < n_jobs; ++j )
  {
    const rsgpu_icp_job_t& J = jobs[j];
    if( !J.object || J.n_batch < 0 || ( J.n_batch > 0 && ( !J.T1 || !J.errs ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_icp_align: bad job" ); }
    total += (size_t)J.n_batch; scratch += (size_t)J.n_batch * (size_t)( J.object->n > 0 ? J.object->n : 1 );
  }
  if( total == 0 ) { return RSGPU_OK; }
  if( total > 2147483647u ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_icp_align: too many poses" ); }
  if( !scan->has_normals ) { return fail( RSGPU_ERR_INVALID, "rsgpu_icp_align: the scan grid has no normals" ); }
  if( !( max_dist > 0.f ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_icp_align: max_dist must be > 0" ); }
  if( max_iter <= 0 ) { max_iter = 100; }
  // RSGPU_ICP_SUMS=fp64 selects block-wide fp64 shuffle reductions instead of the reference-order float sums
  const bool exact = option( "icp_sums" ) != "fp64";
  // "icp_impl": default = two launches per iteration over all running alignments, four iterations replayed as one captured CUDA
  // graph; "split" = the same launches issued one by one (round 1); "persistent" = one launch per
  // batch with a device-side work queue (no host in the loop; measured slower on one GPU - its hand-offs cost more than the
  // launches they replace, and resident blocks that mostly wait take issue slots from the dense search - kept for hosts
  // whose cores are oversubscribed by ranks x lanes); "block" = one resident block per alignment
  const std::string impl = option( "icp_impl" );
  const bool split = impl != "block";
  bool persistent = impl == "persistent";
  cudaStream_t st = rt().stream;
  float ident[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 }, T2i[16];
  mat4_inverse_ref( T2 ? T2 : ident, T2i );
  std::vector<IcpBlock> hb( total );
  std::vector<float> hT( total * 16 );
  std::vector<IcpState> hs( split ? total : 0 );
  // the split variant runs the alignments as NPART interleaved partitions on separate streams, so that the solve
  // launches of one partition (a few busy warps per alignment) overlap the search launches of the others
  int NPART = 2;
  { const std::string o = option( "icp_parts" ); if( !o.empty() ) { NPART = std::min( 4, std::max( 1, atoi( o.c_str() ) ) ); } }
  if( (size_t)NPART > total ) { NPART = (int)total; }
  // object points per warp task of the search launches: small tasks shorten the late iterations (few alignments left,
  // the launch is then pure latency), at the price of idle lanes in the cell-window pass
  // the launch is then pure latency), at the price of idle lanes in the cell-window pass.  A small batch (one object's
  // survivors inside its lane: a few thousand tasks) does not fill the GPU at 16 points per task: 8 there (measured
  // 52.2 -> 50.9 ms per C2 step with 8 lanes), 16 for the big batches.
  size_t total_points = 0;
  for( int j = 0; j < n_jobs; ++j ) { total_points += (size_t)jobs[j].n_batch * (size_t)( jobs[j].object->n > 0 ? jobs[j].object->n : 0 ); }
  int TPT = total_points / 16 < 8192 ? 8 : 16;
  { const std::string o = option( "icp_tpt" ); if( !o.empty() ) { const int v = atoi( o.c_str() ); if( v == 8 || v == 16 || v == 32 ) { TPT = v; } } }
  std::vector<std::vector<int>> part_ids( NPART );
  std::vector<std::vector<unsigned>> part_task( NPART );
  size_t bi = 0, off = 0;
  for( int j = 0; j < n_jobs; ++j )
  {
    const rsgpu_icp_job_t& J = jobs[j];
    for( int b = 0; b < J.n_batch; ++b, ++bi )
    {
      hb[bi].p1 = J.object->pos.p; hb[bi].n1 = J.object->nor.p; hb[bi].n = J.object->n; hb[bi].scratch_off = off;
      memcpy( &hT[bi * 16], J.T1 + 16 * (size_t)b, 64 );
      off += (size_t)( J.object->n > 0 ? J.object->n : 1 );
      if( split )
      {
        memcpy( hs[bi].T, J.T1 + 16 * (size_t)b, 64 );
        hs[bi].max_dist = max_dist; hs[bi].prev_err = 1e6f; hs[bi].err = 1e6f; hs[bi].active = 1; hs[bi].steps = 0;
        const int part = (int)( bi % (size_t)NPART );
        if( part_task[part].empty() ) { part_task[part].push_back( 0u ); }
        const unsigned long long next = (unsigned long long)part_task[part].back() + (unsigned long long)( ( J.object->n + TPT - 1 ) / TPT );
        if( next > 0xffffffffull ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_icp_align: too many points" ); }
        part_ids[part].push_back( (int)bi );
        part_task[part].push_back( (unsigned)next );
      }
    }
  }
  DevBuf<IcpBlock> dB; DevBuf<float> dT, dT2i, derr; DevBuf<int> dit; DevBuf<float4> sq, sp, sn; DevBuf<uint2> sm;
  RS_CUDA( dB.alloc( total ) ); RS_CUDA( dT.alloc( total * 16 ) ); RS_CUDA( dT2i.alloc( 16 ) ); RS_CUDA( derr.alloc( total ) ); RS_CUDA( dit.alloc( total ) );
  RS_CUDA( sq.alloc( scratch ) ); RS_CUDA( sm.alloc( scratch ) ); RS_CUDA( sp.alloc( scratch ) ); RS_CUDA( sn.alloc( scratch ) );
  RS_CUDA( cudaMemcpyAsync( dB.p, hb.data(), sizeof( IcpBlock ) * total, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( dT2i.p, T2i, 64, cudaMemcpyHostToDevice, st ) );
  const size_t tile_bytes = 2 * sizeof( float ) * ( ICP_THREADS - 64 ) * TILE_LD; // two tiles of ICP_THREADS - 64 rows (ordered_sums: warps 0 and 1 sum, the others fill)
  const float dot_thr = compat_threshold_acosf( max_angle );
  std::vector<float> herr( total ); std::vector<int> hit( total );
  // the persistent variant's work list: chunks of ICP_WARPS x TPT points per alignment, every chunk of every possible iteration
  std::vector<unsigned> h_chunks( persistent ? total : 0 );
  size_t chunks_total = 0, live = 0;
  if( persistent )
  {
    for( size_t i = 0; i < total; ++i )
    {
      const size_t c = ( (size_t)( hb[i].n > 0 ? hb[i].n : 0 ) + (size_t)ICP_WARPS * TPT - 1 ) / ( (size_t)ICP_WARPS * TPT );
      if( c >= ( (size_t)1 << ICPQ_CHUNK_BITS ) || total >= ( (size_t)1 << ( 32 - ICPQ_CHUNK_BITS ) ) - 2 ) { persistent = false; break; }
      h_chunks[i] = (unsigned)c; chunks_total += c; live += c > 0;
    }
    if( chunks_total * (size_t)max_iter + 4096 > ( (size_t)1 << 28 ) ) { persistent = false; } // queue above 1 GB: use the split variant
  }
  if( persistent && live > 0 )
  {
    DevBuf<IcpState> dS; DevBuf<unsigned> dchunks, darrived, ditems; DevBuf<IcpQueue> dq;
    int n_sm = 148;
    cudaDeviceGetAttribute( &n_sm, cudaDevAttrMultiProcessorCount, rt().device );
    // blocks of the launch: enough to search every alignment's chunks of one iteration at once, capped - the batches of
    // other objects and the dense search want the SMs too ("icp_ctas")
    unsigned n_ctas = 96 * ( 256 / ICP_THREADS );
    { const std::string o = option( "icp_ctas" ); if( !o.empty() ) { n_ctas = (unsigned)std::max( 1, atoi( o.c_str() ) ); } }
    n_ctas = (unsigned)std::min<size_t>( n_ctas, chunks_total );
    const size_t q_cap = chunks_total * (size_t)max_iter + n_ctas;
    RS_CUDA( dS.alloc( total ) ); RS_CUDA( dchunks.alloc( total ) ); RS_CUDA( darrived.alloc( total ) ); RS_CUDA( ditems.alloc( q_cap ) ); RS_CUDA( dq.alloc( 1 ) );
    RS_CUDA( cudaMemcpyAsync( dS.p, hs.data(), sizeof( IcpState ) * total, cudaMemcpyHostToDevice, st ) );
    RS_CUDA( cudaMemcpyAsync( dchunks.p, h_chunks.data(), sizeof( unsigned ) * total, cudaMemcpyHostToDevice, st ) );
    RS_CUDA( cudaMemsetAsync( darrived.p, 0, sizeof( unsigned ) * total, st ) );
    RS_CUDA( cudaMemsetAsync( ditems.p, 0xff, sizeof( unsigned ) * q_cap, st ) ); // ICPQ_EMPTY
    std::vector<unsigned> first; first.reserve( chunks_total );
    for( size_t i = 0; i < total; ++i ) { for( unsigned c = 0; c < h_chunks[i]; ++c ) { first.push_back( ( (unsigned)i << ICPQ_CHUNK_BITS ) | c ); } }
    RS_CUDA( cudaMemcpyAsync( ditems.p, first.data(), sizeof( unsigned ) * first.size(), cudaMemcpyHostToDevice, st ) );
    IcpQueue hq; hq.head = 0; hq.tail = (unsigned)first.size(); hq.remaining = (unsigned)live; hq.pad = 0;
    RS_CUDA( cudaMemcpyAsync( dq.p, &hq, sizeof( hq ), cudaMemcpyHostToDevice, st ) );
    const size_t search_bytes = sizeof( uint4 ) * ICP_WARPS * rsg::GroupCfg<ICP_G>::CAND_WORDS + ICP_WARPS * 32;
    const size_t smem = exact ? std::max( tile_bytes, search_bytes ) : search_bytes;
    {
      static std::once_flag once;
      cudaError_t ae = cudaSuccess;
      std::call_once( once, [&]() {
        ae = cudaFuncSetAttribute( icp_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max( tile_bytes, search_bytes ) );
        if( ae == cudaSuccess ) { ae = cudaFuncSetAttribute( icp_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)search_bytes ); }
      } );
      RS_CUDA( ae );
    }
    {
      ProfScope prof( "icp" );
      if( exact )
      {
        icp_persistent_kernel<true><<<n_ctas, ICP_THREADS, smem, st>>>( scan->view(), dB.p, dS.p, dchunks.p, darrived.p, dq.p, ditems.p, TPT, dT2i.p, dot_thr, max_iter,
                                                                          sq.p, sm.p, sp.p, sn.p );
      }
      else
      {
        icp_persistent_kernel<false><<<n_ctas, ICP_THREADS, smem, st>>>( scan->view(), dB.p, dS.p, dchunks.p, darrived.p, dq.p, ditems.p, TPT, dT2i.p, dot_thr, max_iter,
                                                                           sq.p, sm.p, sp.p, sn.p );
      }
      RS_CHECK_LAUNCH();
    }
    RS_CUDA( cudaMemcpyAsync( hs.data(), dS.p, sizeof( IcpState ) * total, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( rs::stream_sync( st, true ) );
    for( size_t i = 0; i < total; ++i ) { memcpy( &hT[i * 16], hs[i].T, 64 ); herr[i] = hs[i].err; hit[i] = hs[i].steps; }
  }
  else if( persistent ) // nothing but empty clouds: the reference leaves every one with err = 1e6 and its pose untouched (icp.h:444-456)
  {
    for( size_t i = 0; i < total; ++i ) { herr[i] = 1e6f; hit[i] = 0; }
  }
  else if( split )
  {
    DevBuf<IcpState> dS;
    RS_CUDA( dS.alloc( total ) );
    RS_CUDA( cudaMemcpyAsync( dS.p, hs.data(), sizeof( IcpState ) * total, cudaMemcpyHostToDevice, st ) );
    std::vector<DevBuf<int>> dids( NPART ), dact( NPART ); std::vector<DevBuf<unsigned>> dtask( NPART );
    for( int p = 0; p < NPART; ++p )
    {
      const size_t np = part_ids[p].size();
      RS_CUDA( dids[p].alloc( np ) ); RS_CUDA( dtask[p].alloc( np + 1 ) ); RS_CUDA( dact[p].alloc( (size_t)max_iter ) );
      RS_CUDA( cudaMemcpyAsync( dids[p].p, part_ids[p].data(), sizeof( int ) * np, cudaMemcpyHostToDevice, st ) );
      RS_CUDA( cudaMemcpyAsync( dtask[p].p, part_task[p].data(), sizeof( unsigned ) * ( np + 1 ), cudaMemcpyHostToDevice, st ) );
      RS_CUDA( cudaMemsetAsync( dact[p].p, 0, sizeof( int ) * (size_t)max_iter, st ) );
    }
    if( exact )
    {
      RS_CUDA( cudaFuncSetAttribute( icp_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes ) );
      RS_CUDA( cudaFuncSetAttribute( icp_solve_desc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes ) );
    }
    int n_sm = 148;
    cudaDeviceGetAttribute( &n_sm, cudaDevAttrMultiProcessorCount, rt().device );
    cudaStream_t* aux = nullptr;
    RS_TRY( aux_streams( NPART, &aux ) );
    cudaEvent_t fork = nullptr, join[4] = { nullptr, nullptr, nullptr, nullptr };
    RS_CUDA( cudaEventCreateWithFlags( &fork, cudaEventDisableTiming ) );
    for( int p = 0; p < NPART; ++p ) { RS_CUDA( cudaEventCreateWithFlags( &join[p], cudaEventDisableTiming ) ); }
    int status = RSGPU_OK;
    {
      ProfScope prof( "icp" );
      cudaEventRecord( fork, st );
      for( int p = 0; p < NPART; ++p ) { cudaStreamWaitEvent( aux[p], fork, 0 ); }
      // the host looks at the number of still-running alignments every CHECK iterations (one 4-byte copy per partition)
      const int CHECK = 4;
      std::vector<char> done( NPART, 0 );
      const bool phase_prof = rt().profile && option( "icp_phases" ) == "1";
      const bool use_graph = impl != "split" && !phase_prof;
      if( use_graph )
      {
        // four iterations per driver call: the captured graph of this lane and partition, arguments through the descriptor
        static thread_local IcpGraphCache gc;
        if( !gc.h_desc )
        {
          if( cudaHostAlloc( (void**)&gc.h_desc, sizeof( IcpDesc ) * 4, cudaHostAllocDefault ) != cudaSuccess ||
              cudaHostAlloc( (void**)&gc.h_done, sizeof( int ) * 4, cudaHostAllocDefault ) != cudaSuccess )
          {
            gc.h_desc = nullptr; status = cuda_fail( cudaGetLastError(), "icp graph staging", __FILE__, __LINE__ );
          }
          for( int p = 0; p < 4 && status == RSGPU_OK; ++p )
          {
            if( cudaMalloc( (void**)&gc.d_desc[p], sizeof( IcpDesc ) ) != cudaSuccess || cudaMalloc( (void**)&gc.d_done[p], sizeof( int ) ) != cudaSuccess )
            {
              status = cuda_fail( cudaGetLastError(), "icp graph descriptors", __FILE__, __LINE__ );
            }
          }
        }
        const int ex = exact ? 1 : 0;
        for( int p = 0; p < NPART && status == RSGPU_OK; ++p )
        {
          IcpDesc& D = gc.h_desc[p];
          D.g = scan->view(); D.blocks = dB.p; D.state = dS.p; D.ids = dids[p].p; D.task_start = dtask[p].p;
          D.n_align = (int)part_ids[p].size(); D.pts_per_task = TPT; D.max_iter = max_iter; D.T2i = dT2i.p; D.dot_thr = dot_thr;
          D.sq = sq.p; D.sm = sm.p; D.sp = sp.p; D.sn = sn.p; D.n_done = gc.d_done[p];
          cudaMemcpyAsync( gc.d_desc[p], &D, sizeof( IcpDesc ), cudaMemcpyHostToDevice, aux[p] );
          cudaMemsetAsync( gc.d_done[p], 0, sizeof( int ), aux[p] );
          if( !gc.exec[p][ex] )
          {
            cudaGraph_t graph = nullptr;
            cudaError_t ce = cudaStreamBeginCapture( aux[p], cudaStreamCaptureModeThreadLocal );
            for( int i = 0; i < ICP_GRAPH_ITERS && ce == cudaSuccess; ++i )
            {
              icp_search_desc_kernel<<<(unsigned)n_sm * 4, ICP_THREADS, 0, aux[p]>>>( gc.d_desc[p] );
              if( exact ) { icp_solve_desc_kernel<true><<<64, ICP_THREADS, tile_bytes, aux[p]>>>( gc.d_desc[p] ); }
              else { icp_solve_desc_kernel<false><<<64, ICP_THREADS, 0, aux[p]>>>( gc.d_desc[p] ); }
            }
            if( ce == cudaSuccess ) { ce = cudaStreamEndCapture( aux[p], &graph ); }
            if( ce == cudaSuccess ) { ce = cudaGraphInstantiate( &gc.exec[p][ex], graph, 0 ); }
            if( graph ) { cudaGraphDestroy( graph ); }
            if( ce != cudaSuccess ) { gc.exec[p][ex] = nullptr; status = cuda_fail( ce, "icp graph capture", __FILE__, __LINE__ ); }
          }
        }
        const int rounds = ( max_iter + ICP_GRAPH_ITERS - 1 ) / ICP_GRAPH_ITERS;
        for( int r = 0; r < rounds && status == RSGPU_OK; ++r )
        {
          for( int p = 0; p < NPART; ++p )
          {
            if( done[p] ) { continue; }
            if( cudaGraphLaunch( gc.exec[p][ex], aux[p] ) != cudaSuccess ) { status = cuda_fail( cudaGetLastError(), "icp graph launch", __FILE__, __LINE__ ); break; }
            for( int i = 0; i < 2 * ICP_GRAPH_ITERS; ++i ) { count_launch(); }
            cudaMemcpyAsync( &gc.h_done[p], gc.d_done[p], sizeof( int ), cudaMemcpyDeviceToHost, aux[p] );
          }
          bool all_done = true;
          for( int p = 0; p < NPART && status == RSGPU_OK; ++p )
          {
            if( done[p] ) { continue; }
            if( rs::stream_sync( aux[p] ) != cudaSuccess ) { status = cuda_fail( cudaGetLastError(), "icp partition", __FILE__, __LINE__ ); }
            if( gc.h_done[p] >= (int)part_ids[p].size() ) { done[p] = 1; } else { all_done = false; }
          }
          if( all_done ) { break; }
        }
      }
      for( int it = 0; !use_graph && it < max_iter && status == RSGPU_OK; ++it )
      {
        for( int p = 0; p < NPART; ++p )
        {
          if( done[p] ) { continue; }
          const int np = (int)part_ids[p].size();
          const unsigned n_tasks = part_task[p].back();
          const unsigned search_blocks = (unsigned)std::min<size_t>( ( (size_t)n_tasks + ICP_WARPS - 1 ) / ICP_WARPS, (size_t)n_sm * 16 );
          // "icp_phases" = 1: time the two kinds of launches separately (rsgpu_profile_get "icp_search" / "icp_solve")
          cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
          if( phase_prof ) { cudaEventCreate( &e0 ); cudaEventCreate( &e1 ); cudaEventCreate( &e2 ); cudaEventRecord( e0, aux[p] ); }
          // every alignment of the partition has an empty object cloud: nothing to search (a 0-block launch is invalid); the
          // solve marks them inactive through nc == 0 and they return err = 1e6 like the reference (icp.h:444-456)
          if( search_blocks > 0 )
          {
            icp_search_kernel<<<search_blocks, ICP_THREADS, 0, aux[p]>>>( scan->view(), dB.p, dS.p, dids[p].p, dtask[p].p, np, TPT, dT2i.p, dot_thr, sq.p, sm.p, sp.p, sn.p );
            count_launch();
          }
          if( phase_prof ) { cudaEventRecord( e1, aux[p] ); }
          if( exact ) { icp_solve_kernel<true><<<(unsigned)np, ICP_THREADS, tile_bytes, aux[p]>>>( scan->view(), dB.p, dS.p, dids[p].p, it, sq.p, sm.p, sp.p, sn.p, dact[p].p + it ); }
          else { icp_solve_kernel<false><<<(unsigned)np, ICP_THREADS, 0, aux[p]>>>( scan->view(), dB.p, dS.p, dids[p].p, it, sq.p, sm.p, sp.p, sn.p, dact[p].p + it ); }
          count_launch();
          if( phase_prof ) { cudaEventRecord( e2, aux[p] ); prof_add_pending( "icp_search", e0, e1, true, false ); prof_add_pending( "icp_solve", e1, e2, true, true ); }
        }
        if( cudaGetLastError() != cudaSuccess ) { status = fail( RSGPU_ERR_CUDA, "rsgpu_icp_align: kernel launch failed" ); break; }
        if( it % CHECK == CHECK - 1 || it == max_iter - 1 )
        {
          int running[4] = { 0, 0, 0, 0 };
          for( int p = 0; p < NPART; ++p ) { if( !done[p] ) { cudaMemcpyAsync( &running[p], dact[p].p + it, sizeof( int ), cudaMemcpyDeviceToHost, aux[p] ); } }
          bool all_done = true;
          for( int p = 0; p < NPART; ++p )
          {
            if( done[p] ) { continue; }
            if( rs::stream_sync( aux[p] ) != cudaSuccess ) { status = cuda_fail( cudaGetLastError(), "icp partition", __FILE__, __LINE__ ); }
            if( running[p] == 0 ) { done[p] = 1; } else { all_done = false; }
          }
          if( all_done ) { break; }
        }
      }
      for( int p = 0; p < NPART; ++p ) { cudaEventRecord( join[p], aux[p] ); cudaStreamWaitEvent( st, join[p], 0 ); }
    }
    cudaEventDestroy( fork );
    for( int p = 0; p < NPART; ++p ) { cudaEventDestroy( join[p] ); }
    RS_TRY( status );
    RS_CUDA( cudaMemcpyAsync( hs.data(), dS.p, sizeof( IcpState ) * total, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( rs::stream_sync( st ) );
    for( size_t i = 0; i < total; ++i ) { memcpy( &hT[i * 16], hs[i].T, 64 ); herr[i] = hs[i].err; hit[i] = hs[i].steps; }
  }
  else
  {
    RS_CUDA( cudaMemcpyAsync( dT.p, hT.data(), sizeof( float ) * 16 * total, cudaMemcpyHostToDevice, st ) );
    {
      ProfScope prof( "icp" );
      if( exact )
      {
        RS_CUDA( cudaFuncSetAttribute( icp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes ) );
        icp_kernel<true><<<(unsigned)total, ICP_THREADS, tile_bytes, st>>>( scan->view(), dB.p, dT.p, dT2i.p, max_dist, dot_thr, max_iter, sq.p, sm.p, sp.p, sn.p, derr.p, dit.p );
      }
      else { icp_kernel<false><<<(unsigned)total, ICP_THREADS, 0, st>>>( scan->view(), dB.p, dT.p, dT2i.p, max_dist, dot_thr, max_iter, sq.p, sm.p, sp.p, sn.p, derr.p, dit.p ); }
      RS_CHECK_LAUNCH();
    }
    RS_CUDA( cudaMemcpyAsync( hT.data(), dT.p, sizeof( float ) * 16 * total, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( cudaMemcpyAsync( herr.data(), derr.p, sizeof( float ) * total, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( cudaMemcpyAsync( hit.data(), dit.p, sizeof( int ) * total, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( rs::stream_sync( st ) );
  }
  bi = 0;
  for( int j = 0; j < n_jobs; ++j )
  {
    const rsgpu_icp_job_t& J = jobs[j];
    for( int b = 0; b < J.n_batch; ++b, ++bi )
    {
      memcpy( J.T1 + 16 * (size_t)b, &hT[bi * 16], 64 );
      J.errs[b] = herr[bi];
      if( J.iters ) { J.iters[b] = hit[bi]; }
    }
  }
  return RSGPU_OK;
}
} // namespace

extern "C" int rsgpu_icp_align_multi( const rsgpu_icp_job_t* jobs, int32_t n_jobs, const rsgpu_grid_t* scan, const float* T2,
                                      float max_dist, float max_angle )
{
  return icp_run( jobs, n_jobs, scan, T2, max_dist, max_angle, 100 );
}

extern "C" int rsgpu_icp_align_batch( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scan, float* T1, int32_t n_batch, const float* T2,
                                      float max_dist, float max_angle, float* errs, int32_t* iters )
{
  rsgpu_icp_job_t job = { obj, T1, n_batch, errs, iters };
  return icp_run( &job, 1, scan, T2, max_dist, max_angle, 100 );
}

extern "C" int rsgpu_icp_align_batch_ex( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scan, float* T1, int32_t n_batch, const float* T2,
                                         float max_dist, float max_angle, int32_t max_iter, float* errs, int32_t* iters )
{
  rsgpu_icp_job_t job = { obj, T1, n_batch, errs, iters };
  return icp_run( &job, 1, scan, T2, max_dist, max_angle, max_iter );
}
