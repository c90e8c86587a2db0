// Level building on the GPU: the greedy Poisson-disk subsampling that derives levels 1-4 of an
// rs_pointcloud_t from level 0.
//
// Replaces rs_pointcloud__compute_level_poisson (reference lib/rs/rs_pointcloud.h:984-1037): "in ascending index order,
// the first point not yet marked becomes a sample and marks every point within r of it".  With a symmetric
// neighbour relation that is the lexicographically-first maximal independent set of the r-disk graph, which has an
// exact parallel form: a point is a sample iff none of its EARLIER neighbours is one.  So instead of the reference's
// one-sample-at-a-time loop (millions of dependent single-point searches) the state of every point is resolved by
// propagation:
//
//   init      cnt[i] = number of earlier neighbours; cnt == 0 -> IN (a sample), queued
//   phase A   every newly IN point marks its undecided neighbours OUT and queues them
//   phase B   every newly OUT point decrements cnt of its later undecided neighbours; a count reaching 0 means all
//             earlier neighbours are OUT -> IN, queued for the next phase A
//
// Every point is queued exactly once, so the total work is three neighbourhood scans per point; the number of rounds
// is the depth of the dependency chains (hundreds to a few thousand), which is why the rounds run inside ONE
// cooperative launch with grid-wide barriers instead of two launches per round.
//
// Exactness: the neighbour test is the reference's own, dist^2 = (vx*vx + vy*vy) + vz*vz < (float)((double)r*r)
// (msh_hash_grid.h:852-857), symmetric in its two points.  The reference's search returns at most max_n_neigh points
// (:994-995, 1017); a sample whose ball holds more would mark only the nearest ones there, which this formulation does
// not model: such inputs fail loudly (RSGPU_ERR_UNSUPPORTED) instead of returning a different level.  The cell size of
// the grid used here is our own choice (about r instead of the reference's 5 r): the set of points within r does not
// depend on it (the window is taken with a 1e-4 relative margin, the distance test is exact).
#include "rsgpu_internal.cuh"
#include <cooperative_groups.h>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <algorithm>
#include <cmath>
#include <vector>

using namespace rs;
namespace cg = cooperative_groups;

namespace
{
enum : int { UNDECIDED = 0, IN = 1, OUT = 2 };

struct LevelCtx
{
  int* state;            // per grid position
  int* cnt;              // undecided-or-OUT earlier neighbours still to hear from
  uint32_t* in_list;     // grid positions in the order they became IN
  uint32_t* out_list;    // ... OUT
  unsigned* counters;    // [0] in_tail, [1] out_tail, [2] k-cap violated, [3] rounds
};

// f( position, is_earlier ) for every point within r of the point at grid position p (itself included), the 32 lanes of
// the calling warp striding over the contiguous x-rows of the window
template <class F>
__device__ __forceinline__ void for_each_neighbor( const GridView& g, uint32_t p, double radius_w, float r2f, F f )
{
  const int lane = threadIdx.x & 31;
  const float4 q = __ldg( g.recs + p );
  const uint32_t my_idx = __float_as_uint( q.w );
  const CellWindow w = make_window( g, q.x, q.y, q.z, radius_w );
  for( int iz = 0; iz < w.nz; ++iz )
  {
    for( int iy = 0; iy < w.ny; ++iy )
    {
      const size_t row = ( (size_t)( w.loz + iz ) * g.H + ( w.loy + iy ) ) * g.W + w.lox;
      const uint32_t s = __ldg( g.cell_start + row ), t = __ldg( g.cell_start + row + w.nx );
      for( uint32_t j = s + lane; j < t; j += 32 )
      {
        const float4 rec = __ldg( g.recs + j );
        if( dist2_exact( rec, q.x, q.y, q.z ) < r2f ) { f( j, __float_as_uint( rec.w ) < my_idx ); }
      }
    }
  }
}

__global__ void __launch_bounds__( 256 ) level_init_kernel( GridView g, double radius_w, float r2f, LevelCtx c )
{
  const int lane = threadIdx.x & 31;
  const uint32_t warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5, n_warps = ( gridDim.x * blockDim.x ) >> 5;
  for( uint32_t p = warp; p < (uint32_t)g.n_pts; p += n_warps )
  {
    int earlier = 0;
    for_each_neighbor( g, p, radius_w, r2f, [&]( uint32_t, bool is_earlier ) { earlier += is_earlier; } );
    for( int o = 16; o > 0; o >>= 1 ) { earlier += __shfl_xor_sync( RS_FULL, earlier, o ); }
    if( lane == 0 )
    {
      c.cnt[p] = earlier;
      c.state[p] = earlier == 0 ? IN : UNDECIDED;
      if( earlier == 0 ) { c.in_list[atomicAdd( c.counters + 0, 1u )] = p; }
    }
  }
}

__global__ void __launch_bounds__( 256 ) level_propagate_kernel( GridView g, double radius_w, float r2f, int max_n_neigh, LevelCtx c )
{
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const uint32_t warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5, n_warps = ( gridDim.x * blockDim.x ) >> 5;
  volatile unsigned* counters = c.counters;
  uint32_t in_begin = 0, in_end = counters[0], out_begin = 0;
  unsigned rounds = 0;
  while( in_begin < in_end )
  {
    // ---- phase A: the new samples mark their undecided neighbours OUT
    for( uint32_t f = in_begin + warp; f < in_end; f += n_warps )
    {
      const uint32_t p = c.in_list[f];
      int ball = 0;
      for_each_neighbor( g, p, radius_w, r2f, [&]( uint32_t j, bool ) {
        ++ball;
        if( j != p && atomicExch( c.state + j, OUT ) == UNDECIDED ) { c.out_list[atomicAdd( c.counters + 1, 1u )] = j; }
      } );
      for( int o = 16; o > 0; o >>= 1 ) { ball += __shfl_xor_sync( RS_FULL, ball, o ); }
      if( lane == 0 && ball > max_n_neigh ) { c.counters[2] = 1u; }
    }
    grid.sync();
    const uint32_t out_end = counters[1];
    // ---- phase B: the newly OUT points report to their later undecided neighbours
    for( uint32_t f = out_begin + warp; f < out_end; f += n_warps )
    {
      const uint32_t p = c.out_list[f];
      for_each_neighbor( g, p, radius_w, r2f, [&]( uint32_t j, bool is_earlier ) {
        if( is_earlier || j == p ) { return; }
        if( ( (volatile int*)c.state )[j] != UNDECIDED ) { return; }
        if( atomicSub( c.cnt + j, 1 ) == 1 )
        {
          c.state[j] = IN; // no OUT mark can race with this: those are only made in phase A
          c.in_list[atomicAdd( c.counters + 0, 1u )] = j;
        }
      } );
    }
    grid.sync();
    in_begin = in_end; in_end = counters[0]; out_begin = out_end;
    ++rounds;
  }
  if( blockIdx.x == 0 && threadIdx.x == 0 ) { c.counters[3] = rounds; }
}

__global__ void level_flags_kernel( const float4* __restrict__ recs, const int* __restrict__ state, int n, unsigned char* __restrict__ flag, unsigned* __restrict__ undecided )
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if( p >= n ) { return; }
  const int s = state[p];
  flag[__float_as_uint( recs[p].w )] = s == IN;
  if( s == UNDECIDED ) { atomicAdd( undecided, 1u ); }
}
} // namespace

extern "C" {

int rsgpu_poisson_level( const float* pts, int32_t n, float voxel, int32_t max_n_neigh, int32_t* out_indices, int32_t* n_out, int32_t* n_rounds )
{
  if( n < 0 || !n_out || ( n > 0 && ( !pts || !out_indices ) ) || !( voxel > 0.f ) || max_n_neigh <= 0 )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_poisson_level: bad argument" );
  }
  RS_TRY( ensure_device() );
  *n_out = 0;
  if( n_rounds ) { *n_rounds = 0; }
  if( n == 0 ) { return RSGPU_OK; }
  cudaStream_t st = rt().stream;
  // cell size: just above r (a window is then at most 3 x 3 x 3 cells), coarser when the extent would need more than
  // 2^25 cells of that size
  double mn[3] = { 1e300, 1e300, 1e300 }, mx[3] = { -1e300, -1e300, -1e300 };
  for( int32_t i = 0; i < n; ++i )
  {
    for( int a = 0; a < 3; ++a )
    {
      const double v = pts[3 * (size_t)i + a];
      if( !( v == v ) || fabs( v ) > 1e8 ) { return fail( RSGPU_ERR_INVALID, "rsgpu_poisson_level: non-finite or huge coordinate" ); }
      mn[a] = v < mn[a] ? v : mn[a]; mx[a] = v > mx[a] ? v : mx[a];
    }
  }
  const double r = (double)voxel;
  double cell = 1.001 * r;
  const double vol = ( mx[0] - mn[0] + cell ) * ( mx[1] - mn[1] + cell ) * ( mx[2] - mn[2] + cell );
  const double floor_cell = cbrt( vol / 33554432.0 );
  if( cell < floor_cell ) { cell = floor_cell; }
  const double radius_w = r * 1.0001; // the window's radius: a margin over r so that no in-radius point can fall outside it
  const float r2f = (float)( r * r ); // the reference's radius_sq (msh_hash_grid.h:1111)
  DevBuf<float> d_pts;
  RS_CUDA( d_pts.alloc( (size_t)n * 3 ) );
  RS_CUDA( cudaMemcpyAsync( d_pts.p, pts, sizeof( float ) * 3 * (size_t)n, cudaMemcpyHostToDevice, st ) );
  rsgpu_grid_t* grid = nullptr;
  RS_TRY( rsgpu_grid_create_dev( d_pts.p, n, (float)( 0.5 * cell ), &grid ) );
  struct Guard { rsgpu_grid_t* g; ~Guard() { rsgpu_grid_destroy( g ); } } guard{ grid };
  if( !( grid->info.cell_size > radius_w ) ) { return fail( RSGPU_ERR_CUDA, "rsgpu_poisson_level: internal: cell size not above the radius" ); }
  const GridView g = grid->view();

  DevBuf<int> state, cnt; DevBuf<uint32_t> in_list, out_list; DevBuf<unsigned> counters; DevBuf<unsigned char> flag; DevBuf<int32_t> d_out, d_nsel;
  RS_CUDA( state.alloc( n ) ); RS_CUDA( cnt.alloc( n ) ); RS_CUDA( in_list.alloc( n ) ); RS_CUDA( out_list.alloc( n ) );
  RS_CUDA( counters.alloc( 8 ) ); RS_CUDA( flag.alloc( n ) ); RS_CUDA( d_out.alloc( n ) ); RS_CUDA( d_nsel.alloc( 1 ) );
  RS_CUDA( cudaMemsetAsync( counters.p, 0, sizeof( unsigned ) * 8, st ) );
  LevelCtx c; c.state = state.p; c.cnt = cnt.p; c.in_list = in_list.p; c.out_list = out_list.p; c.counters = counters.p;
  {
    ProfScope prof( "levels" );
    int n_sm = 148;
    cudaDeviceGetAttribute( &n_sm, cudaDevAttrMultiProcessorCount, rt().device );
    const unsigned init_blocks = (unsigned)std::min<long long>( ( (long long)n + 7 ) / 8, (long long)n_sm * 32 );
    level_init_kernel<<<init_blocks, 256, 0, st>>>( g, radius_w, r2f, c );
    RS_CHECK_LAUNCH();
    int per_sm = 0;
    RS_CUDA( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &per_sm, level_propagate_kernel, 256, 0 ) );
    if( per_sm < 1 ) { return fail( RSGPU_ERR_CUDA, "rsgpu_poisson_level: the propagation kernel does not fit an SM" ); }
    const unsigned coop_blocks = (unsigned)n_sm * (unsigned)std::min( per_sm, 2 );
    GridView gv = g; double rw = radius_w; float r2 = r2f; int mk = max_n_neigh;
    void* args[] = { &gv, &rw, &r2, &mk, &c };
    RS_CUDA( cudaLaunchCooperativeKernel( (const void*)level_propagate_kernel, dim3( coop_blocks ), dim3( 256 ), args, 0, st ) );
    count_launch();
    level_flags_kernel<<<( n + 255 ) / 256, 256, 0, st>>>( g.recs, state.p, n, flag.p, counters.p + 4 );
    RS_CHECK_LAUNCH();
  }
  // ascending level-0 indices of the samples (the level's arrays are copies of those rows, rs_pointcloud.h:1077-1086)
  thrust::counting_iterator<int32_t> iota( 0 );
  size_t tmp_bytes = 0;
  RS_CUDA( cub::DeviceSelect::Flagged( nullptr, tmp_bytes, iota, flag.p, d_out.p, d_nsel.p, n, st ) );
  DevBuf<unsigned char> tmp;
  RS_CUDA( tmp.alloc( tmp_bytes ) );
  RS_CUDA( cub::DeviceSelect::Flagged( tmp.p, tmp_bytes, iota, flag.p, d_out.p, d_nsel.p, n, st ) );
  unsigned hc[8]; int32_t nsel = 0;
  RS_CUDA( cudaMemcpyAsync( hc, counters.p, sizeof( hc ), cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( cudaMemcpyAsync( &nsel, d_nsel.p, sizeof( int32_t ), cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  if( hc[2] )
  {
    return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_poisson_level: a sample's ball holds more than max_n_neigh points; the reference then marks only the "
                                        "nearest max_n_neigh (rs_pointcloud.h:994-1017), which this entry point does not reproduce" );
  }
  if( hc[4] ) { return fail( RSGPU_ERR_CUDA, "rsgpu_poisson_level: internal: undecided points left after propagation" ); }
  RS_CUDA( cudaMemcpyAsync( out_indices, d_out.p, sizeof( int32_t ) * (size_t)nsel, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  *n_out = nsel;
  if( n_rounds ) { *n_rounds = (int32_t)hc[3]; }
  return RSGPU_OK;
}

} // extern "C"
