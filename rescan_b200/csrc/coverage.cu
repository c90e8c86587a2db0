// Coverage term of the arrangement optimiser on the GPU.
//
// segment_transfer scores an arrangement (a set of placed objects) partly by how much of the scan it explains: the scan
// and the placed dynamic objects are rasterised into the same 5 cm grid and the score is |cells lit by both| / |cells lit
// by the scan| (reference apps/segment_transfer/arrangement_optimization.cpp:343-373, rasterisation :1064-1106, grid
// lib/rs/intersect.h:57-116).  The reference re-rasterises the whole arrangement for every one of the 25 000 simulated
// annealing moves.  Here every candidate placement is rasterised ONCE, into a bit mask over the scan's lit cells; the
// coverage of any arrangement is then popcount( OR of its placements' masks ) / n_lit - a few hundred 32-bit words per
// placement, which the (sequential, host-side) optimiser can combine at memory speed.  Integer work: masks and counts
// are bit-identical to the reference's grids.
//
//   rsgpu_rasterize_points   points (optionally posed) -> byte grid        rsao_rasterize_scene_to_grid :1064-1080
//   rsgpu_coverage_masks     (object, pose) list -> bit masks over the scan's lit cells   rsao__rasterize_arrangement_to_grid :1082-1106
#include "rsgpu_internal.cuh"
#include <cub/cub.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <cmath>
#include <cstring>
#include <vector>

using namespace rs;

namespace
{
struct CovGrid
{
  float ox, oy, oz, inv_voxel; // isect_grid3d_cell_from_world_space: (pt - origin) * (1.0f / voxel_size), floorf (intersect.h:102-116)
  int xr, yr, zr;
};

// cell of a world-space point, -1 outside the grid; layout y * (xr * zr) + z * xr + x (:115)
__device__ __forceinline__ long long cov_cell( const CovGrid& g, float px, float py, float pz )
{
  const float fx = floorf( __fmul_rn( __fsub_rn( px, g.ox ), g.inv_voxel ) );
  const float fy = floorf( __fmul_rn( __fsub_rn( py, g.oy ), g.inv_voxel ) );
  const float fz = floorf( __fmul_rn( __fsub_rn( pz, g.oz ), g.inv_voxel ) );
  // non-finite coordinates: the reference's x86 (int) cast gives INT_MIN (rejected below); CUDA's gives 0 for NaN
  if( !( fx == fx ) || !( fy == fy ) || !( fz == fz ) ) { return -1; }
  const int x = (int)fx, y = (int)fy, z = (int)fz;
  if( x < 0 || x >= g.xr || y < 0 || y >= g.yr || z < 0 || z >= g.zr ) { return -1; }
  return ( (long long)y * g.zr + z ) * g.xr + x;
}

__global__ void rasterize_kernel( const float* __restrict__ pts, int n, const float* __restrict__ pose, CovGrid g, unsigned char* __restrict__ grid )
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) { return; }
  float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
  if( pose ) { float qx, qy, qz; xf_apply( pose, px, py, pz, 1.0f, qx, qy, qz ); px = qx; py = qy; pz = qz; }
  const long long c = cov_cell( g, px, py, pz );
  if( c >= 0 ) { grid[c] = 1; }
}

struct CovJob
{
  const float* pos; // the object's level-2 points (arrangement_optimization.cpp:1089: lvl = 2)
  int n;
};

struct LitFlag
{
  __host__ __device__ int operator()( unsigned char v ) const { return v > 0 ? 1 : 0; }
};

// one block per placement: its points under its pose light bits of the placement's mask (bit = rank of the cell among
// the scan's lit cells = exclusive prefix count of lit cells; cells the scan does not light cannot count towards the
// score and have no bit)
__global__ void __launch_bounds__( 256 ) coverage_mask_kernel( const CovJob* __restrict__ jobs, const float* __restrict__ poses, CovGrid g,
                                                               const unsigned char* __restrict__ scene, const int* __restrict__ rank_of_cell,
                                                               int n_words, unsigned* __restrict__ masks )
{
  const CovJob job = jobs[blockIdx.x];
  const float* m = poses + 16 * (size_t)blockIdx.x;
  unsigned* mask = masks + (size_t)blockIdx.x * n_words;
  for( int i = threadIdx.x; i < job.n; i += blockDim.x )
  {
    float px, py, pz;
    xf_apply( m, job.pos[3 * (size_t)i], job.pos[3 * (size_t)i + 1], job.pos[3 * (size_t)i + 2], 1.0f, px, py, pz );
    const long long c = cov_cell( g, px, py, pz );
    if( c < 0 ) { continue; }
    if( scene[c] == 0 ) { continue; }
    const int b = rank_of_cell[c];
    atomicOr( mask + ( b >> 5 ), 1u << ( b & 31 ) );
  }
}

int make_grid( const float* origin, const int32_t* res, float voxel, CovGrid& g, long long& n_cells )
{
  if( !origin || !res || !( voxel > 0.f ) || res[0] <= 0 || res[1] <= 0 || res[2] <= 0 ) { return fail( RSGPU_ERR_INVALID, "rsgpu coverage: bad grid" ); }
  n_cells = (long long)res[0] * res[1] * res[2];
  if( n_cells > ( 1ll << 31 ) - 1 ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu coverage: more than 2^31 cells" ); }
  g.ox = origin[0]; g.oy = origin[1]; g.oz = origin[2];
  volatile float inv = 1.0f / voxel; // float division, like the reference's inv_voxel_size
  g.inv_voxel = inv;
  g.xr = res[0]; g.yr = res[1]; g.zr = res[2];
  return RSGPU_OK;
}
} // namespace

extern "C" {

int rsgpu_rasterize_points( const float* pts, int32_t n, const float* pose, const float origin[3], const int32_t res[3], float voxel,
                            uint8_t* grid )
{
  if( n < 0 || ( n > 0 && !pts ) || !grid ) { return fail( RSGPU_ERR_INVALID, "rsgpu_rasterize_points: bad argument" ); }
  RS_TRY( ensure_device() );
  CovGrid g; long long n_cells = 0;
  RS_TRY( make_grid( origin, res, voxel, g, n_cells ) );
  if( n == 0 ) { return RSGPU_OK; }
  cudaStream_t st = rt().stream;
  DevBuf<float> d_pts, d_pose; DevBuf<unsigned char> d_grid;
  RS_CUDA( d_pts.alloc( (size_t)n * 3 ) ); RS_CUDA( d_pose.alloc( 16 ) ); RS_CUDA( d_grid.alloc( (size_t)n_cells ) );
  RS_CUDA( cudaMemcpyAsync( d_pts.p, pts, sizeof( float ) * 3 * (size_t)n, cudaMemcpyHostToDevice, st ) );
  if( pose ) { RS_CUDA( cudaMemcpyAsync( d_pose.p, pose, 64, cudaMemcpyHostToDevice, st ) ); }
  RS_CUDA( cudaMemcpyAsync( d_grid.p, grid, (size_t)n_cells, cudaMemcpyHostToDevice, st ) ); // the caller's grid is OR-ed into, not cleared
  {
    ProfScope prof( "coverage" );
    rasterize_kernel<<<( n + 255 ) / 256, 256, 0, st>>>( d_pts.p, n, pose ? d_pose.p : nullptr, g, d_grid.p );
    RS_CHECK_LAUNCH();
  }
  RS_CUDA( cudaMemcpyAsync( grid, d_grid.p, (size_t)n_cells, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  return RSGPU_OK;
}

int rsgpu_coverage_masks( const rsgpu_cloud_t* const* objects, const float* poses, int32_t n_poses, const float origin[3],
                          const int32_t res[3], float voxel, const uint8_t* scene_grid, uint32_t* out_masks, int32_t n_words,
                          int32_t* n_lit )
{
  if( n_poses < 0 || !scene_grid || !n_lit || ( n_poses > 0 && ( !objects || !poses ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_coverage_masks: bad argument" ); }
  RS_TRY( ensure_device() );
  CovGrid g; long long n_cells = 0;
  RS_TRY( make_grid( origin, res, voxel, g, n_cells ) );
  // bit of every cell = its rank among the scan's lit cells in ascending cell index: an exclusive prefix count on the device
  cudaStream_t st = rt().stream;
  DevBuf<unsigned char> d_scene; DevBuf<int> d_rank; DevBuf<unsigned char> d_tmp;
  RS_CUDA( d_scene.alloc( (size_t)n_cells ) ); RS_CUDA( d_rank.alloc( (size_t)n_cells ) );
  RS_CUDA( cudaMemcpyAsync( d_scene.p, scene_grid, (size_t)n_cells, cudaMemcpyHostToDevice, st ) );
  int last_rank = 0;
  {
    ProfScope prof( "coverage" );
    auto flags = thrust::make_transform_iterator( (const unsigned char*)d_scene.p, LitFlag() );
    size_t tmp_bytes = 0;
    RS_CUDA( cub::DeviceScan::ExclusiveSum( nullptr, tmp_bytes, flags, d_rank.p, (int)n_cells, st ) );
    RS_CUDA( d_tmp.alloc( tmp_bytes ) );
    RS_CUDA( cub::DeviceScan::ExclusiveSum( d_tmp.p, tmp_bytes, flags, d_rank.p, (int)n_cells, st ) );
  }
  RS_CUDA( cudaMemcpyAsync( &last_rank, d_rank.p + ( n_cells - 1 ), sizeof( int ), cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  const int lit = last_rank + ( scene_grid[n_cells - 1] > 0 ? 1 : 0 );
  *n_lit = lit;
  const int need_words = ( lit + 31 ) / 32;
  if( n_poses == 0 || !out_masks ) { return RSGPU_OK; } // a caller sizing its buffer: n_words = (*n_lit + 31) / 32
  if( n_words < need_words ) { return fail( RSGPU_ERR_INVALID, "rsgpu_coverage_masks: n_words is smaller than ceil( lit cells / 32 )" ); }
  if( n_words == 0 ) { return RSGPU_OK; }
  std::vector<CovJob> jobs( n_poses );
  for( int32_t i = 0; i < n_poses; ++i )
  {
    if( !objects[i] ) { return fail( RSGPU_ERR_INVALID, "rsgpu_coverage_masks: NULL object" ); }
    jobs[i].pos = objects[i]->pos.p; jobs[i].n = objects[i]->n;
  }
  DevBuf<CovJob> d_jobs; DevBuf<float> d_poses; DevBuf<unsigned> d_masks;
  RS_CUDA( d_jobs.alloc( n_poses ) ); RS_CUDA( d_poses.alloc( (size_t)n_poses * 16 ) );
  RS_CUDA( d_masks.alloc( (size_t)n_poses * n_words ) );
  RS_CUDA( cudaMemcpyAsync( d_jobs.p, jobs.data(), sizeof( CovJob ) * (size_t)n_poses, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( d_poses.p, poses, sizeof( float ) * 16 * (size_t)n_poses, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemsetAsync( d_masks.p, 0, sizeof( unsigned ) * (size_t)n_poses * n_words, st ) );
  {
    ProfScope prof( "coverage" );
    coverage_mask_kernel<<<(unsigned)n_poses, 256, 0, st>>>( d_jobs.p, d_poses.p, g, d_scene.p, d_rank.p, n_words, d_masks.p );
    RS_CHECK_LAUNCH();
  }
  RS_CUDA( cudaMemcpyAsync( out_masks, d_masks.p, sizeof( unsigned ) * (size_t)n_poses * n_words, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  return RSGPU_OK;
}

} // extern "C"
