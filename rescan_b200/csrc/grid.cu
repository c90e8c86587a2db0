// GPU spatial hash grid build: replaces msh_hash_grid__init (reference lib/msh/msh_hash_grid.h:388-541).
//
// The reference inserts points one by one into an open-addressing map cell -> bin and then concatenates the
// bins in ascending cell id, each bin in insertion (= original index) order (:501-532).  Here the same layout
// falls out of ONE stable radix sort of (cell id, original index) pairs, and the hash map is replaced by a
// dense cell_start[n_cells + 1] table — the map probes are 56 % of the reference's search time.
#include "rsgpu_internal.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <cstring>
#include <cmath>
#include <vector>

using namespace rs;

namespace
{
// order-preserving float <-> uint mapping for atomic min/max
__device__ __forceinline__ uint32_t f2ord( float f )
{
  uint32_t u = __float_as_uint( f );
  return ( u & 0x80000000u ) ? ~u : ( u | 0x80000000u );
}
inline float ord2f( uint32_t o )
{
  uint32_t u = ( o & 0x80000000u ) ? ( o & 0x7fffffffu ) : ~o;
  float f; memcpy( &f, &u, 4 ); return f;
}

// bounding box (:413-434 minus the padding, which the host applies in float)
__global__ void bbox_kernel( const float* __restrict__ pts, int n, uint32_t* __restrict__ out /*[6] min xyz, max xyz*/ )
{
  float mn[3] = { 1e9f, 1e9f, 1e9f }, mx[3] = { -1e9f, -1e9f, -1e9f };
  for( int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x )
  {
#pragma unroll
    for( int a = 0; a < 3; ++a ) { float v = pts[3 * (size_t)i + a]; mn[a] = fminf( mn[a], v ); mx[a] = fmaxf( mx[a], v ); }
  }
#pragma unroll
  for( int a = 0; a < 3; ++a )
  {
    uint32_t lo = __reduce_min_sync( RS_FULL, f2ord( mn[a] ) ), hi = __reduce_max_sync( RS_FULL, f2ord( mx[a] ) );
    if( ( threadIdx.x & 31 ) == 0 ) { atomicMin( out + a, lo ); atomicMax( out + 3 + a, hi ); }
  }
}

// cell id of every point: float subtraction, then times the double inverse cell size, truncated (:471-475)
__global__ void cell_key_kernel( const float* __restrict__ pts, int n, float mnx, float mny, float mnz, double inv_cell,
                                 int W, int H, int D, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals )
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) { return; }
  float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
  long long cx = __double2ll_rz( __dmul_rn( (double)__fsub_rn( x, mnx ), inv_cell ) );
  long long cy = __double2ll_rz( __dmul_rn( (double)__fsub_rn( y, mny ), inv_cell ) );
  long long cz = __double2ll_rz( __dmul_rn( (double)__fsub_rn( z, mnz ), inv_cell ) );
  // points lie inside the padded box by construction; the clamp only guards NaN / inf input
  cx = cx < 0 ? 0 : ( cx >= W ? W - 1 : cx ); cy = cy < 0 ? 0 : ( cy >= H ? H - 1 : cy ); cz = cz < 0 ? 0 : ( cz >= D ? D - 1 : cz );
  keys[i] = (uint32_t)( ( cz * H + cy ) * W + cx );
  vals[i] = (uint32_t)i;
}

// re-lay the points as 16-byte records in sorted order and histogram the cells (count of cell c at c + 1,
// so that an inclusive scan of the table turns it into cell_start)
__global__ void relay_kernel( const float* __restrict__ pts, int n, const uint32_t* __restrict__ keys,
                              const uint32_t* __restrict__ vals, float4* __restrict__ recs,
                              uint32_t* __restrict__ cell_start, uint32_t* __restrict__ stats )
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) { return; }
  uint32_t src = vals[i];
  float4 r;
  r.x = pts[3 * (size_t)src]; r.y = pts[3 * (size_t)src + 1]; r.z = pts[3 * (size_t)src + 2]; r.w = __uint_as_float( src );
  recs[i] = r;
  uint32_t key = keys[i];
  if( i == 0 || key != keys[i - 1] )
  {
    // first point of a non-empty cell: find the run length (runs are short: points per cell)
    int j = i + 1;
    while( j < n && keys[j] == key ) { ++j; }
    cell_start[(size_t)key + 1] = (uint32_t)( j - i );
    atomicAdd( stats + 0, 1u );           // non-empty cells
    atomicMax( stats + 1, (uint32_t)( j - i ) ); // max_n_pts_in_bin
  }
}

// occ27[c] = number of points in the (clipped) 3x3x3 block of cells around c; rows along x are contiguous
// in the dense table, so a block is 9 range lookups
__global__ void occ27_kernel( const uint32_t* __restrict__ cell_start, int W, int H, int D, uint32_t* __restrict__ occ )
{
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t n = (size_t)W * H * D;
  if( c >= n ) { return; }
  int x = (int)( c % W ); size_t r = c / W; int y = (int)( r % H ); int z = (int)( r / H );
  int x0 = x > 0 ? x - 1 : 0, x1 = x < W - 1 ? x + 1 : W - 1;
  uint32_t total = 0;
  for( int zz = ( z > 0 ? z - 1 : 0 ); zz <= ( z < D - 1 ? z + 1 : D - 1 ); ++zz )
    for( int yy = ( y > 0 ? y - 1 : 0 ); yy <= ( y < H - 1 ? y + 1 : H - 1 ); ++yy )
    {
      size_t rowbase = ( (size_t)zz * H + yy ) * W;
      total += cell_start[rowbase + x1 + 1] - cell_start[rowbase + x0];
    }
  occ[c] = total;
}

struct NonZero
{
  __host__ __device__ uint32_t operator()( uint32_t v ) const { return v != 0u ? 1u : 0u; }
};
__global__ void active_cells_kernel( const uint32_t* __restrict__ occ27, const uint32_t* __restrict__ crank, size_t n_cells, uint32_t* __restrict__ acells )
{
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if( c < n_cells && occ27[c] != 0u ) { acells[crank[c]] = (uint32_t)c; }
}

// bounding box of the points of every cell (cbox) and of every 3x3x3 block of cells (nbox): {lo, hi} pairs, empty = {+inf, -inf}
__global__ void cbox_kernel( const float4* __restrict__ recs, const uint32_t* __restrict__ cell_start, size_t n_cells, float4* __restrict__ cbox )
{
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if( c >= n_cells ) { return; }
  const float inf = __int_as_float( 0x7f800000 );
  float4 lo = make_float4( inf, inf, inf, 0.f ), hi = make_float4( -inf, -inf, -inf, 0.f );
  for( uint32_t p = cell_start[c]; p < cell_start[c + 1]; ++p )
  {
    const float4 r = recs[p];
    lo.x = fminf( lo.x, r.x ); lo.y = fminf( lo.y, r.y ); lo.z = fminf( lo.z, r.z );
    hi.x = fmaxf( hi.x, r.x ); hi.y = fmaxf( hi.y, r.y ); hi.z = fmaxf( hi.z, r.z );
  }
  cbox[2 * c] = lo; cbox[2 * c + 1] = hi;
}
__global__ void nbox_kernel( const float4* __restrict__ cbox, int W, int H, int D, float4* __restrict__ nbox )
{
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t n = (size_t)W * H * D;
  if( c >= n ) { return; }
  int x = (int)( c % W ); size_t r = c / W; int y = (int)( r % H ); int z = (int)( r / H );
  const float inf = __int_as_float( 0x7f800000 );
  float4 lo = make_float4( inf, inf, inf, 0.f ), hi = make_float4( -inf, -inf, -inf, 0.f );
  for( int zz = max( z - 1, 0 ); zz <= min( z + 1, D - 1 ); ++zz )
    for( int yy = max( y - 1, 0 ); yy <= min( y + 1, H - 1 ); ++yy )
      for( int xx = max( x - 1, 0 ); xx <= min( x + 1, W - 1 ); ++xx )
      {
        size_t id = ( (size_t)zz * H + yy ) * W + xx;
        const float4 a = cbox[2 * id], b = cbox[2 * id + 1];
        lo.x = fminf( lo.x, a.x ); lo.y = fminf( lo.y, a.y ); lo.z = fminf( lo.z, a.z );
        hi.x = fmaxf( hi.x, b.x ); hi.y = fmaxf( hi.y, b.y ); hi.z = fmaxf( hi.z, b.z );
      }
  nbox[2 * c] = lo; nbox[2 * c + 1] = hi;
}

__global__ void relay_normals_kernel( const float* __restrict__ nor, int n, const float4* __restrict__ recs, float4* __restrict__ out )
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) { return; }
  uint32_t src = __float_as_uint( recs[i].w );
  out[i] = make_float4( nor[3 * (size_t)src], nor[3 * (size_t)src + 1], nor[3 * (size_t)src + 2], 0.f );
}

// per-cell normal cone (see ConeCull in nearest.cuh): unit mean normal and the cosine of the widest angle to it,
// shrunk by a safety margin; flags[0] is raised when some normal is not unit length to 1e-4
__global__ void cone_kernel( const float4* __restrict__ nrm, const uint32_t* __restrict__ cell_start, size_t n_cells,
                             float4* __restrict__ cone, uint32_t* __restrict__ flags )
{
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if( c >= n_cells ) { return; }
  uint32_t s = cell_start[c], t = cell_start[c + 1];
  float4 out = make_float4( 0.f, 0.f, 0.f, -1.f );
  if( s < t )
  {
    float sx = 0.f, sy = 0.f, sz = 0.f; bool unit = true;
    for( uint32_t p = s; p < t; ++p )
    {
      float4 m = nrm[p];
      sx += m.x; sy += m.y; sz += m.z;
      float l2 = m.x * m.x + m.y * m.y + m.z * m.z;
      if( !( fabsf( l2 - 1.0f ) < 2e-4f ) ) { unit = false; }
    }
    if( !unit ) { atomicOr( flags, 1u ); }
    float len = sqrtf( sx * sx + sy * sy + sz * sz );
    if( unit && len > 1e-3f )
    {
      float ux = sx / len, uy = sy / len, uz = sz / len, cmin = 1.0f;
      for( uint32_t p = s; p < t; ++p ) { float4 m = nrm[p]; cmin = fminf( cmin, m.x * ux + m.y * uy + m.z * uz ); }
      out = make_float4( ux, uy, uz, cmin - 1e-4f );
    }
  }
  cone[c] = out;
}

// cone enclosing every normal of the 3x3x3 block of cells around c: axis = count-weighted mean of the cells' axes,
// half-angle = max over cells of (angle between axes + the cell's own half-angle), all rounded outwards
__global__ void ncone_kernel( const float4* __restrict__ cone, const uint32_t* __restrict__ cell_start, int W, int H, int D,
                              float4* __restrict__ ncone )
{
  size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t n = (size_t)W * H * D;
  if( c >= n ) { return; }
  int x = (int)( c % W ); size_t r = c / W; int y = (int)( r % H ); int z = (int)( r / H );
  float sx = 0.f, sy = 0.f, sz = 0.f; bool usable = true; int filled = 0;
  for( int zz = max( z - 1, 0 ); zz <= min( z + 1, D - 1 ); ++zz )
    for( int yy = max( y - 1, 0 ); yy <= min( y + 1, H - 1 ); ++yy )
      for( int xx = max( x - 1, 0 ); xx <= min( x + 1, W - 1 ); ++xx )
      {
        size_t id = ( (size_t)zz * H + yy ) * W + xx;
        uint32_t cnt = cell_start[id + 1] - cell_start[id];
        if( !cnt ) { continue; }
        float4 u = cone[id];
        if( !( u.w > 0.0f ) ) { usable = false; }
        sx += u.x * (float)cnt; sy += u.y * (float)cnt; sz += u.z * (float)cnt; ++filled;
      }
  float4 out = make_float4( 0.f, 0.f, 0.f, filled ? -1.f : RS_NCONE_EMPTY ); // an empty block is marked: one load tells a query so
  float len = sqrtf( sx * sx + sy * sy + sz * sz );
  if( usable && filled && len > 1e-3f )
  {
    float ux = sx / len, uy = sy / len, uz = sz / len, worst = 0.0f; // worst = largest (axis angle + half-angle), radians
    for( int zz = max( z - 1, 0 ); zz <= min( z + 1, D - 1 ); ++zz )
      for( int yy = max( y - 1, 0 ); yy <= min( y + 1, H - 1 ); ++yy )
        for( int xx = max( x - 1, 0 ); xx <= min( x + 1, W - 1 ); ++xx )
        {
          size_t id = ( (size_t)zz * H + yy ) * W + xx;
          if( cell_start[id + 1] == cell_start[id] ) { continue; }
          float4 u = cone[id];
          float a = acosf( fminf( 1.0f, fmaxf( -1.0f, u.x * ux + u.y * uy + u.z * uz ) ) ) + acosf( fminf( 1.0f, u.w ) );
          worst = fmaxf( worst, a );
        }
    worst += 2e-3f;
    if( worst < 1.5f ) { out = make_float4( ux, uy, uz, cosf( worst ) ); }
  }
  ncone[c] = out;
}

__global__ void unpack_recs_kernel( const float4* __restrict__ recs, int n, float* __restrict__ xyz, int32_t* __restrict__ idx )
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) { return; }
  float4 r = recs[i];
  xyz[3 * (size_t)i] = r.x; xyz[3 * (size_t)i + 1] = r.y; xyz[3 * (size_t)i + 2] = r.z; idx[i] = (int32_t)__float_as_uint( r.w );
}

int build_from_device( const float* d_pts, int32_t n, float radius, rsgpu_grid_t** out )
{
  cudaStream_t st = rt().stream;
  ProfScope prof( "grid_build" );
  rsgpu_grid* g = new rsgpu_grid();
  struct Guard { rsgpu_grid* g; ~Guard() { delete g; } } guard{ g };
  memset( &g->info, 0, sizeof( g->info ) );

  // 1. bounding box, padded by 1e-4f in float (:413-434)
  DevBuf<uint32_t> d_box;
  RS_CUDA( d_box.alloc( 8 ) );
  uint32_t init[8] = { 0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 0u, 0u };
  RS_CUDA( cudaMemcpyAsync( d_box.p, init, sizeof( init ), cudaMemcpyHostToDevice, st ) );
  if( n > 0 )
  {
    int blocks = ( n + 255 ) / 256; blocks = blocks > 1184 ? 1184 : blocks;
    bbox_kernel<<<blocks, 256, 0, st>>>( d_pts, n, d_box.p );
    RS_CHECK_LAUNCH();
  }
  uint32_t hb[8];
  RS_CUDA( cudaMemcpyAsync( hb, d_box.p, sizeof( hb ), cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  float mn[3], mx[3], ext[3];
  for( int a = 0; a < 3; ++a )
  {
    mn[a] = n > 0 ? ord2f( hb[a] ) : 1e9f; mx[a] = n > 0 ? ord2f( hb[3 + a] ) : -1e9f;
    // the reference starts from +-1e9 stored as float, so points beyond that never move the box
    if( mn[a] > 1e9f ) { mn[a] = 1e9f; }
    if( mx[a] < -1e9f ) { mx[a] = -1e9f; }
    volatile float pmx = mx[a] + 0.0001f, pmn = mn[a] - 0.0001f;
    mx[a] = pmx; mn[a] = pmn;
    volatile float e = mx[a] - mn[a];
    ext[a] = e;
  }
  float max_ext = ext[0] > ext[1] ? ext[0] : ext[1];
  max_ext = max_ext > ext[2] ? max_ext : ext[2];
  // 2. cell size and dimensions (:443-450)
  double cell;
  if( radius > 0.0 ) { cell = 2.0 * radius; }
  else { cell = max_ext / ( 32 * sqrtf( 3.0f ) ); }
  long long dim[3];
  for( int a = 0; a < 3; ++a ) { dim[a] = (int)( ext[a] / cell + 1.0 ); }
  double inv_cell = 1.0f / cell;
  if( n <= 0 ) { dim[0] = dim[1] = dim[2] = 1; }
  for( int a = 0; a < 3; ++a )
  {
    if( dim[a] < 1 || !( cell > 0 ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_create: degenerate bounding box / cell size" ); }
  }
  double n_cells_d = (double)dim[0] * (double)dim[1] * (double)dim[2];
  if( n_cells_d > 1073741824.0 )
  {
    return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_grid_create: more than 2^30 cells; the dense cell table does not cover this radius/extent" );
  }
  size_t n_cells = (size_t)( dim[0] * dim[1] * dim[2] );
  g->info.width = dim[0]; g->info.height = dim[1]; g->info.depth = dim[2];
  g->info.cell_size = cell; g->info.inv_cell_size = inv_cell;
  for( int a = 0; a < 3; ++a ) { g->info.min_pt[a] = mn[a]; g->info.max_pt[a] = mx[a]; }
  g->info.n_pts = n > 0 ? n : 0;

  RS_CUDA( g->recs.alloc( (size_t)( n > 0 ? n : 0 ) ) );
  RS_CUDA( g->cell_start.alloc( n_cells + 1 ) );
  DevBuf<uint32_t> stats;
  RS_CUDA( stats.alloc( 2 ) );
  RS_CUDA( cudaMemsetAsync( stats.p, 0, 8, st ) );
  if( n <= 0 )
  {
    RS_CUDA( cudaMemsetAsync( g->cell_start.p, 0, sizeof( uint32_t ) * ( n_cells + 1 ), st ) );
    RS_CUDA( g->occ27.alloc( n_cells ) );
    RS_CUDA( cudaMemsetAsync( g->occ27.p, 0, sizeof( uint32_t ) * n_cells, st ) );
  }
  else
  {
    // 3. (cell id, original index) pairs, stable radix sort over just the bits a cell id needs
    DevBuf<uint32_t> k0, k1, v0, v1;
    RS_CUDA( k0.alloc( n ) ); RS_CUDA( k1.alloc( n ) ); RS_CUDA( v0.alloc( n ) ); RS_CUDA( v1.alloc( n ) );
    int blocks = ( n + 255 ) / 256;
    cell_key_kernel<<<blocks, 256, 0, st>>>( d_pts, n, mn[0], mn[1], mn[2], inv_cell, (int)dim[0], (int)dim[1], (int)dim[2], k0.p, v0.p );
    RS_CHECK_LAUNCH();
    int end_bit = 1;
    while( end_bit < 32 && ( (size_t)1 << end_bit ) < n_cells ) { ++end_bit; }
    size_t tmp_bytes = 0;
    RS_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, tmp_bytes, k0.p, k1.p, v0.p, v1.p, n, 0, end_bit, st ) );
    DevBuf<unsigned char> tmp;
    RS_CUDA( tmp.alloc( tmp_bytes ) );
    RS_CUDA( cub::DeviceRadixSort::SortPairs( tmp.p, tmp_bytes, k0.p, k1.p, v0.p, v1.p, n, 0, end_bit, st ) );
    // 4. records + cell histogram, then one in-place inclusive scan -> cell_start
    RS_CUDA( cudaMemsetAsync( g->cell_start.p, 0, sizeof( uint32_t ) * ( n_cells + 1 ), st ) );
    relay_kernel<<<blocks, 256, 0, st>>>( d_pts, n, k1.p, v1.p, g->recs.p, g->cell_start.p, stats.p );
    RS_CHECK_LAUNCH();
    size_t scan_bytes = 0;
    RS_CUDA( cub::DeviceScan::InclusiveSum( nullptr, scan_bytes, g->cell_start.p, g->cell_start.p, (int64_t)( n_cells + 1 ), st ) );
    DevBuf<unsigned char> tmp2;
    RS_CUDA( tmp2.alloc( scan_bytes ) );
    RS_CUDA( cub::DeviceScan::InclusiveSum( tmp2.p, scan_bytes, g->cell_start.p, g->cell_start.p, (int64_t)( n_cells + 1 ), st ) );
    RS_CUDA( g->occ27.alloc( n_cells ) );
    occ27_kernel<<<(unsigned)( ( n_cells + 255 ) / 256 ), 256, 0, st>>>( g->cell_start.p, (int)dim[0], (int)dim[1], (int)dim[2], g->occ27.p );
    RS_CHECK_LAUNCH();
    // ranks of the cells with a non-empty 3x3x3 block (the bins of the dense pose search are indexed by them: 64 sub-bins
    // per ACTIVE cell instead of 8 per cell of the whole table)
    {
      RS_CUDA( g->crank.alloc( n_cells + 1 ) );
      size_t rank_bytes = 0;
      auto flags = thrust::make_transform_iterator( (const uint32_t*)g->occ27.p, NonZero() );
      RS_CUDA( cub::DeviceScan::ExclusiveSum( nullptr, rank_bytes, flags, g->crank.p, (int64_t)n_cells, st ) );
      DevBuf<unsigned char> tmp3;
      RS_CUDA( tmp3.alloc( rank_bytes ) );
      RS_CUDA( cub::DeviceScan::ExclusiveSum( tmp3.p, rank_bytes, flags, g->crank.p, (int64_t)n_cells, st ) );
      uint32_t last_rank = 0, last_occ = 0;
      RS_CUDA( cudaMemcpyAsync( &last_rank, g->crank.p + ( n_cells - 1 ), 4, cudaMemcpyDeviceToHost, st ) );
      RS_CUDA( cudaMemcpyAsync( &last_occ, g->occ27.p + ( n_cells - 1 ), 4, cudaMemcpyDeviceToHost, st ) );
      RS_CUDA( rs::stream_sync( st ) );
      g->n_active = last_rank + ( last_occ != 0 ? 1u : 0u );
      RS_CUDA( g->acells.alloc( g->n_active ) );
      active_cells_kernel<<<(unsigned)( ( n_cells + 255 ) / 256 ), 256, 0, st>>>( g->occ27.p, g->crank.p, n_cells, g->acells.p );
      RS_CHECK_LAUNCH();
    }
    // point bounding boxes per cell / per 3x3x3 block (distance culling of the dense pose search); skipped for very large
    // tables (64 B per cell)
    if( n_cells <= ( (size_t)1 << 25 ) )
    {
      RS_CUDA( g->cbox.alloc( 2 * n_cells ) ); RS_CUDA( g->nbox.alloc( 2 * n_cells ) );
      cbox_kernel<<<(unsigned)( ( n_cells + 255 ) / 256 ), 256, 0, st>>>( g->recs.p, g->cell_start.p, n_cells, g->cbox.p );
      RS_CHECK_LAUNCH();
      nbox_kernel<<<(unsigned)( ( n_cells + 255 ) / 256 ), 256, 0, st>>>( g->cbox.p, (int)dim[0], (int)dim[1], (int)dim[2], g->nbox.p );
      RS_CHECK_LAUNCH();
      g->has_boxes = true;
    }
    RS_CUDA( rs::stream_sync( st ) ); // temporaries die here
  }
  uint32_t hs[2] = { 0, 0 };
  RS_CUDA( cudaMemcpyAsync( hs, stats.p, 8, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  g->info.n_bins = hs[0]; g->info.max_n_pts_in_bin = hs[1];
  guard.g = nullptr;
  *out = g;
  return RSGPU_OK;
}
} // namespace

extern "C" {

int rsgpu_grid_create_dev( const float* d_pts, int32_t n, float radius, rsgpu_grid_t** out )
{
  if( !out || ( n > 0 && !d_pts ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_create: bad argument" ); }
  RS_TRY( ensure_device() );
  return build_from_device( d_pts, n, radius, out );
}

int rsgpu_grid_create( const float* pts, int32_t n, float radius, rsgpu_grid_t** out )
{
  if( !out || ( n > 0 && !pts ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_create: bad argument" ); }
  RS_TRY( ensure_device() );
  DevBuf<float> d;
  RS_CUDA( d.alloc( (size_t)( n > 0 ? n : 0 ) * 3 ) );
  if( n > 0 ) { RS_CUDA( cudaMemcpyAsync( d.p, pts, sizeof( float ) * 3 * (size_t)n, cudaMemcpyHostToDevice, rt().stream ) ); }
  int s = build_from_device( d.p, n, radius, out );
  rs::stream_sync( rt().stream );
  return s;
}

void rsgpu_grid_destroy( rsgpu_grid_t* g ) { delete g; }

int rsgpu_grid_set_normals_dev( rsgpu_grid_t* g, const float* d_nor )
{
  if( !g || ( g->info.n_pts > 0 && !d_nor ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_set_normals: bad argument" ); }
  RS_TRY( ensure_device() );
  int n = (int)g->info.n_pts;
  RS_CUDA( g->nrm.alloc( n ) );
  if( n > 0 )
  {
    relay_normals_kernel<<<( n + 255 ) / 256, 256, 0, rt().stream>>>( d_nor, n, g->recs.p, g->nrm.p );
    RS_CHECK_LAUNCH();
  }
  g->has_normals = true;
  // normal cones per cell for the compatible-neighbour searches
  size_t n_cells = (size_t)( g->info.width * g->info.height * g->info.depth );
  g->has_cone = false;
  if( n > 0 )
  {
    DevBuf<uint32_t> flag;
    RS_CUDA( flag.alloc( 1 ) );
    RS_CUDA( cudaMemsetAsync( flag.p, 0, 4, rt().stream ) );
    RS_CUDA( g->cone.alloc( n_cells ) );
    cone_kernel<<<(unsigned)( ( n_cells + 127 ) / 128 ), 128, 0, rt().stream>>>( g->nrm.p, g->cell_start.p, n_cells, g->cone.p, flag.p );
    RS_CHECK_LAUNCH();
    RS_CUDA( g->ncone.alloc( n_cells ) );
    ncone_kernel<<<(unsigned)( ( n_cells + 127 ) / 128 ), 128, 0, rt().stream>>>( g->cone.p, g->cell_start.p, (int)g->info.width, (int)g->info.height,
                                                                                   (int)g->info.depth, g->ncone.p );
    RS_CHECK_LAUNCH();
    uint32_t h = 0;
    RS_CUDA( cudaMemcpyAsync( &h, flag.p, 4, cudaMemcpyDeviceToHost, rt().stream ) );
    RS_CUDA( rs::stream_sync( rt().stream ) );
    g->has_cone = h == 0; // any non-unit normal disables the culling for this grid
  }
  return RSGPU_OK;
}

int rsgpu_grid_set_normals( rsgpu_grid_t* g, const float* nor )
{
  if( !g || ( g->info.n_pts > 0 && !nor ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_set_normals: bad argument" ); }
  RS_TRY( ensure_device() );
  size_t n = (size_t)g->info.n_pts;
  DevBuf<float> d;
  RS_CUDA( d.alloc( n * 3 ) );
  if( n ) { RS_CUDA( cudaMemcpyAsync( d.p, nor, sizeof( float ) * 3 * n, cudaMemcpyHostToDevice, rt().stream ) ); }
  int s = rsgpu_grid_set_normals_dev( g, d.p );
  rs::stream_sync( rt().stream );
  return s;
}

int rsgpu_grid_get_info( const rsgpu_grid_t* g, rsgpu_grid_info_t* info )
{
  if( !g || !info ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_get_info: bad argument" ); }
  *info = g->info;
  return RSGPU_OK;
}

int rsgpu_grid_get_data( const rsgpu_grid_t* g, float* xyz, int32_t* idx )
{
  if( !g || !xyz || !idx ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_get_data: bad argument" ); }
  RS_TRY( ensure_device() );
  int n = (int)g->info.n_pts;
  if( n == 0 ) { return RSGPU_OK; }
  DevBuf<float> dx; DevBuf<int32_t> di;
  RS_CUDA( dx.alloc( (size_t)n * 3 ) ); RS_CUDA( di.alloc( n ) );
  unpack_recs_kernel<<<( n + 255 ) / 256, 256, 0, rt().stream>>>( g->recs.p, n, dx.p, di.p );
  RS_CHECK_LAUNCH();
  RS_CUDA( cudaMemcpyAsync( xyz, dx.p, sizeof( float ) * 3 * (size_t)n, cudaMemcpyDeviceToHost, rt().stream ) );
  RS_CUDA( cudaMemcpyAsync( idx, di.p, sizeof( int32_t ) * (size_t)n, cudaMemcpyDeviceToHost, rt().stream ) );
  RS_CUDA( rs::stream_sync( rt().stream ) );
  return RSGPU_OK;
}

} // extern "C"
