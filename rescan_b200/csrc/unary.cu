// Per-vertex label transfer and unary data terms for the gco graph cut (which itself stays reference host
// code): replaces rspf__assign_temporary_labels (reference lib/rs/rs_pointcloud_filters.cpp:738-778) and the
// data_cost block of rspf_smooth_labels (:926-939).
//
// Label transfer: one thread per scan vertex walks the placements in the reference's order and keeps the
// running arg-min — the V x A radius searches with k = 1 of the reference collapse into one kernel with
// a per-thread nearest-neighbour scan of each object's grid (objects are small and cache resident; almost
// all (vertex, placement) pairs are rejected by the clipped cell window before any point is read).
// Unary terms: a pure streaming write of V x L int32 (HBM-write bound).
#include "rsgpu_internal.cuh"
#include <cmath>
#include <cstring>
#include <vector>

using namespace rs;

namespace rs
{
void mat4_inverse_ref( const float* m, float* o );
}

namespace
{
struct PlacementDev
{
  GridView grid;   // the object's level-1 grid (with normals)
  float inv[16];   // msh_mat4_inverse(pose)            (:750)
  float nmat[16];  // msh_mat4_transpose(pose)          (:751)
};

__global__ void __launch_bounds__( 256 ) assign_labels_kernel( const float* __restrict__ pos, const float* __restrict__ nor, int n,
                                                               const PlacementDev* __restrict__ plc, int first, int last, double radius,
                                                               float r2f, float dot_thr, int8_t* __restrict__ labels,
                                                               float* __restrict__ min_d )
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if( j >= n ) { return; }
  const float sx = pos[3 * (size_t)j], sy = pos[3 * (size_t)j + 1], sz = pos[3 * (size_t)j + 2];
  const float tx = nor[3 * (size_t)j], ty = nor[3 * (size_t)j + 1], tz = nor[3 * (size_t)j + 2];
  float best = min_d[j];
  int8_t label = labels[j];
  for( int i = first; i < last; ++i )
  {
    const PlacementDev& P = plc[i];
    float qx, qy, qz;
    xf_apply( P.inv, sx, sy, sz, 1.0f, qx, qy, qz );
    CellWindow w = make_window( P.grid, qx, qy, qz, radius );
    // 1-NN inside the radius: smallest (d2, original index)
    uint32_t bd = 0xffffffffu, bpos = 0, bidx = 0xffffffffu;
    for( int e = 0; e < w.n_cells; ++e )
    {
      uint32_t s, t; float gap2;
      window_cell( P.grid, w, e, s, t, gap2 );
      if( s >= t || !( gap2 < r2f ) || __float_as_uint( gap2 ) > bd ) { continue; }
      for( uint32_t p = s; p < t; ++p )
      {
        float4 rec = __ldg( P.grid.recs + p );
        float d2 = dist2_exact( rec, qx, qy, qz );
        uint32_t db = __float_as_uint( d2 ), id = __float_as_uint( rec.w );
        if( d2 < r2f && ( db < bd || ( db == bd && id < bidx ) ) ) { bd = db; bpos = p; bidx = id; }
      }
    }
    if( bd == 0xffffffffu ) { continue; }               // n_neighbors == 0 (:760)
    float d2 = __uint_as_float( bd );
    if( !( d2 < best ) ) { continue; }                   // strict: the earlier placement keeps ties (:761)
    float ax, ay, az;
    xf_apply( P.nmat, tx, ty, tz, 0.0f, ax, ay, az );
    float4 m = __ldg( P.grid.nrm + bpos );
    // msh_vec3_normalize (msh_vec_math.h:868-872): multiply by 1.0f / sqrtf(dot)
    float ia = __fdiv_rn( 1.0f, sqrtf( dot3_exact( ax, ay, az, ax, ay, az ) ) );
    float ib = __fdiv_rn( 1.0f, sqrtf( dot3_exact( m.x, m.y, m.z, m.x, m.y, m.z ) ) );
    float dot = dot3_exact( __fmul_rn( ax, ia ), __fmul_rn( ay, ia ), __fmul_rn( az, ia ), __fmul_rn( m.x, ib ), __fmul_rn( m.y, ib ), __fmul_rn( m.z, ib ) );
    float ad = fabsf( dot );
    if( ad >= dot_thr && ad <= 1.0f ) { best = d2; label = (int8_t)( i + 1 ); } // acos(|dot|) < 70 deg (:767-768)
  }
  min_d[j] = best; labels[j] = label;
}

// data_cost (:926-939): one warp-coalesced row sweep; thread per output element
__global__ void __launch_bounds__( 256 ) unary_costs_kernel( const int32_t* __restrict__ labels, const uint8_t* __restrict__ is_static, size_t total,
                                                             int L, int32_t* __restrict__ cost )
{
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for( size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += stride )
  {
    size_t v = e / (size_t)L; int l = (int)( e - v * (size_t)L );
    int lab = __ldg( labels + v );
    int c = 30;
    if( __ldg( is_static + lab ) ) { c = 15; }
    if( lab == 0 ) { c = 1; }
    cost[e] = ( l == lab ) ? 0 : c;
  }
}

// smallest float d in [0,1] with (double)acosf(d) < max_angle
float label_threshold( double max_angle )
{
  auto ok = [&]( float d ) { return (double)acosf( d ) < max_angle; };
  if( ok( 0.0f ) ) { return 0.0f; }
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
} // namespace

extern "C" {

int rsgpu_assign_labels( const float* scan_pos, const float* scan_nor, int32_t n, const float* poses, const rsgpu_grid_t* const* grids,
                         int32_t first, int32_t last, float radius, int8_t* labels, float* min_dists )
{
  if( n < 0 || first < 0 || last < first || ( n > 0 && ( !scan_pos || !scan_nor || !labels || !min_dists ) ) ||
      ( last > first && ( !poses || !grids ) ) )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_assign_labels: bad argument" );
  }
  if( last > 127 ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_assign_labels: more than 127 placements do not fit the reference's int8 labels" ); }
  RS_TRY( ensure_device() );
  if( n == 0 || last == first ) { return RSGPU_OK; }
  if( !( radius > 0.f ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_assign_labels: radius must be > 0" ); }
  std::vector<PlacementDev> h( last );
  for( int i = first; i < last; ++i )
  {
    if( !grids[i] || !grids[i]->has_normals ) { return fail( RSGPU_ERR_INVALID, "rsgpu_assign_labels: object grid missing or without normals" ); }
    h[i].grid = grids[i]->view();
    const float* M = poses + 16 * (size_t)i;
    mat4_inverse_ref( M, h[i].inv );
    for( int c = 0; c < 4; ++c ) for( int r = 0; r < 4; ++r ) { h[i].nmat[4 * c + r] = M[4 * r + c]; }
  }
  for( int i = 0; i < first; ++i ) { memset( &h[i], 0, sizeof( PlacementDev ) ); }
  cudaStream_t st = rt().stream;
  DevBuf<PlacementDev> dp; DevBuf<float> dpos, dnor, dmin; DevBuf<int8_t> dlab;
  RS_CUDA( dp.alloc( last ) ); RS_CUDA( dpos.alloc( (size_t)n * 3 ) ); RS_CUDA( dnor.alloc( (size_t)n * 3 ) ); RS_CUDA( dmin.alloc( n ) ); RS_CUDA( dlab.alloc( n ) );
  RS_CUDA( cudaMemcpyAsync( dp.p, h.data(), sizeof( PlacementDev ) * (size_t)last, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( dpos.p, scan_pos, sizeof( float ) * 3 * (size_t)n, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( dnor.p, scan_nor, sizeof( float ) * 3 * (size_t)n, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( dmin.p, min_dists, sizeof( float ) * (size_t)n, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( dlab.p, labels, (size_t)n, cudaMemcpyHostToDevice, st ) );
  double r = radius;
  float r2f = (float)( r * r );
  double max_angle = 70.0 * 0.005555555556 * 3.1415926535897932384626433832; // msh_deg2rad(70.0)
  {
    ProfScope prof( "labels" );
    assign_labels_kernel<<<( n + 255 ) / 256, 256, 0, st>>>( dpos.p, dnor.p, n, dp.p, first, last, r, r2f, label_threshold( max_angle ), dlab.p, dmin.p );
    RS_CHECK_LAUNCH();
  }
  RS_CUDA( cudaMemcpyAsync( min_dists, dmin.p, sizeof( float ) * (size_t)n, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( cudaMemcpyAsync( labels, dlab.p, (size_t)n, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  return RSGPU_OK;
}

int rsgpu_unary_costs( const int32_t* labels, const uint8_t* label_is_static, int32_t n, int32_t L, int32_t* data_cost )
{
  if( n < 0 || L <= 0 || ( n > 0 && ( !labels || !label_is_static || !data_cost ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_unary_costs: bad argument" ); }
  RS_TRY( ensure_device() );
  if( n == 0 ) { return RSGPU_OK; }
  for( int i = 0; i < n; ++i )
  {
    if( labels[i] < 0 || labels[i] >= L ) { return fail( RSGPU_ERR_INVALID, "rsgpu_unary_costs: label out of [0, n_labels)" ); }
  }
  cudaStream_t st = rt().stream;
  size_t total = (size_t)n * L;
  DevBuf<int32_t> dl, dc; DevBuf<uint8_t> ds;
  RS_CUDA( dl.alloc( n ) ); RS_CUDA( ds.alloc( L ) ); RS_CUDA( dc.alloc( total ) );
  RS_CUDA( cudaMemcpyAsync( dl.p, labels, sizeof( int32_t ) * (size_t)n, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( ds.p, label_is_static, (size_t)L, cudaMemcpyHostToDevice, st ) );
  {
    ProfScope prof( "unary" );
    size_t blocks = ( total + 255 ) / 256; if( blocks > 148 * 32 ) { blocks = 148 * 32; }
    unary_costs_kernel<<<(unsigned)blocks, 256, 0, st>>>( dl.p, ds.p, total, L, dc.p );
    RS_CHECK_LAUNCH();
  }
  RS_CUDA( cudaMemcpyAsync( data_cost, dc.p, sizeof( int32_t ) * total, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  return RSGPU_OK;
}

} // extern "C"
