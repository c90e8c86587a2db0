// Internal declarations shared by the rsgpu translation units (sm_100a only).
//
// Exactness contract (SURVEY.md §7 "hard parts" 1): the reference is x86-64 gcc -O3 without FMA, so every
// float/double expression a result depends on is written here with the non-contracting intrinsics
// (__fmul_rn/__fadd_rn/__dmul_rn/...) in the reference's evaluation order; the library is additionally built
// with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <mutex>
#include "rsgpu.h"

#define RS_FULL 0xffffffffu
#define RS_INF_BITS 0x7f800000u
#define RS_NCONE_EMPTY 2.0f /* GridView::ncone[c].w of a 3x3x3 block without points (a cosine never exceeds 1) */

// ---------------------------------------------------------------------------------------------- host runtime
namespace rs
{
struct Runtime
{
  cudaStream_t stream = 0;
  int device = 0;
  bool profile = false;
};
Runtime& rt();
int fail( int code, const std::string& msg );
int cuda_fail( cudaError_t e, const char* what, const char* file, int line );
int ensure_device();
std::string option( const char* key );   // value set by rsgpu_set_option, else environment RSGPU_<KEY>, else ""
void set_option( const char* key, const char* value );
int aux_streams( int n, cudaStream_t** out ); // helper streams (non-blocking) for internally overlapped work
cudaError_t stream_sync( cudaStream_t s, bool long_wait = false ); // host wait; long waits sleep on a blocking event (runtime.cu)
cudaStream_t bulk_stream();                   // where long throughput launches go: low priority inside a lane, else rt().stream
void prof_add_pending( const char* name, cudaEvent_t a, cudaEvent_t b, bool own_a, bool own_b );
void count_launch(); // one of OUR kernels was launched (library kernels such as CUB are not counted)

// RAII device buffer from the stream-ordered pool (cudaMallocAsync): allocation and release are enqueued on the
// library stream, so per-call temporaries cost no device synchronisation (cudaFree would)
template <typename T>
struct DevBuf
{
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf( const DevBuf& ) = delete;
  DevBuf& operator=( const DevBuf& ) = delete;
  ~DevBuf() { release(); }
  void release()
  {
    if( p ) { cudaFreeAsync( p, rt().stream ); }
    p = nullptr; n = 0;
  }
  cudaError_t alloc( size_t count )
  {
    release();
    n = count;
    return cudaMallocAsync( (void**)&p, sizeof( T ) * ( count ? count : 1 ), rt().stream );
  }
  // grow-only variant for scratch that a thread keeps between calls: reallocates only when `count` exceeds the capacity
  cudaError_t reserve( size_t count )
  {
    if( p && count <= n ) { return cudaSuccess; }
    return alloc( count );
  }
};

// scoped per-kernel event timer (active only when profiling is enabled)
struct ProfScope
{
  const char* name;
  cudaStream_t st = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
  explicit ProfScope( const char* n );          // on the calling thread's stream
  ProfScope( const char* n, cudaStream_t s );
  ~ProfScope();
};
} // namespace rs

#define RS_CUDA( call )                                                                     \
  do {                                                                                      \
    cudaError_t e__ = ( call );                                                             \
    if( e__ != cudaSuccess ) { return rs::cuda_fail( e__, #call, __FILE__, __LINE__ ); }    \
  } while( 0 )
#define RS_CHECK_LAUNCH() do { rs::count_launch(); RS_CUDA( cudaGetLastError() ); } while( 0 )
#define RS_TRY( call )                                                                      \
  do { int s__ = ( call ); if( s__ != RSGPU_OK ) { return s__; } } while( 0 )

// ---------------------------------------------------------------------------------------------- device views
// Grid as kernels see it.  Dense layout: cell (x,y,z) -> id = (z*H + y)*W + x (x fastest, like
// msh_hash_grid.h:375-379), points of cell id are recs[cell_start[id] .. cell_start[id+1]) in ascending
// original index (msh_hash_grid.h:501-532), one 16-byte record {x, y, z, bitcast(idx)} per point like the
// reference's msh_hg_v3i_t, normals (if set) in the same order as 16-byte {nx, ny, nz, 0}.
struct GridView
{
  const float4* __restrict__ recs;
  const float4* __restrict__ nrm;
  const uint32_t* __restrict__ cell_start;
  const uint32_t* __restrict__ occ27; // points in the 3x3x3 block of cells centred on each cell (0 = a query there sees nothing)
  const float4* __restrict__ cone;    // per cell {unit mean normal, cos(max angle to it)}; nullptr when normals are not unit
  const float4* __restrict__ ncone;   // the same for all normals in the 3x3x3 block around each cell
  const float4* __restrict__ cbox;    // per cell: bounding box of ITS POINTS, {lo.xyz, -} at 2c, {hi.xyz, -} at 2c + 1 (lo > hi: empty); may be nullptr
  const float4* __restrict__ nbox;    // the same for the points of the 3x3x3 block around each cell; may be nullptr
  const uint32_t* __restrict__ crank; // per cell: its index among the cells with a non-empty 3x3x3 block ("active": the only cells a
                                      // query with any neighbour can call home), undefined for the others; may be nullptr
  const uint32_t* __restrict__ acells; // active rank -> cell id
  uint32_t n_active;
  float mnx, mny, mnz;
  int W, H, D;
  double cell, inv_cell;
  int n_pts;
};

struct rsgpu_grid
{
  rs::DevBuf<float4> recs;
  rs::DevBuf<float4> nrm;
  rs::DevBuf<uint32_t> cell_start;
  rs::DevBuf<uint32_t> occ27;
  rs::DevBuf<float4> cone;
  rs::DevBuf<float4> ncone;
  rs::DevBuf<float4> cbox, nbox;
  rs::DevBuf<uint32_t> crank, acells;
  uint32_t n_active = 0;
  // second copy of the records for radius searches on grids with hundreds of points per cell (csrc/search.cu): every cell's
  // records ordered by 4x4x4 sub-cell (then original index), sub_off[c * 64 + s] = first record of sub-cell s of cell c.
  // Built on the first such search (the handle is logically const: the result of a search does not depend on it).
  mutable rs::DevBuf<float4> sub_recs;
  mutable rs::DevBuf<uint32_t> sub_off;
  mutable int sub_state = 0; // 0 not built, 1 built, -1 not worth it / too large
  mutable std::mutex sub_mu;
  bool has_boxes = false;
  bool has_cone = false;
  bool has_normals = false;
  rsgpu_grid_info_t info;
  GridView view() const
  {
    GridView v;
    v.recs = recs.p; v.nrm = has_normals ? nrm.p : nullptr; v.cell_start = cell_start.p; v.occ27 = occ27.p; v.cone = ( has_normals && has_cone ) ? cone.p : nullptr; v.ncone = v.cone ? ncone.p : nullptr;
    v.cbox = has_boxes ? cbox.p : nullptr; v.nbox = has_boxes ? nbox.p : nullptr;
    v.crank = crank.p; v.acells = acells.p; v.n_active = n_active;
    v.mnx = info.min_pt[0]; v.mny = info.min_pt[1]; v.mnz = info.min_pt[2];
    v.W = (int)info.width; v.H = (int)info.height; v.D = (int)info.depth;
    v.cell = info.cell_size; v.inv_cell = info.inv_cell_size; v.n_pts = (int)info.n_pts;
    return v;
  }
};

struct rsgpu_cloud
{
  rs::DevBuf<float> pos; // n x 3
  rs::DevBuf<float> nor; // n x 3
  int32_t n = 0;
};

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------- exact math
// msh_mat4_vec3_mul (msh_vec_math.h:1554-1561): ((m0*x + m4*y) + m8*z) + w*m12, float, left to right
__device__ __forceinline__ void xf_apply( const float* __restrict__ m, float x, float y, float z, float w,
                                          float& ox, float& oy, float& oz )
{
  ox = __fadd_rn( __fadd_rn( __fadd_rn( __fmul_rn( m[0], x ), __fmul_rn( m[4], y ) ), __fmul_rn( m[8], z ) ), __fmul_rn( w, m[12] ) );
  oy = __fadd_rn( __fadd_rn( __fadd_rn( __fmul_rn( m[1], x ), __fmul_rn( m[5], y ) ), __fmul_rn( m[9], z ) ), __fmul_rn( w, m[13] ) );
  oz = __fadd_rn( __fadd_rn( __fadd_rn( __fmul_rn( m[2], x ), __fmul_rn( m[6], y ) ), __fmul_rn( m[10], z ) ), __fmul_rn( w, m[14] ) );
}

// squared distance of msh_hash_grid__find_neighbors_in_bin (msh_hash_grid.h:852-855): (vx*vx + vy*vy) + vz*vz
__device__ __forceinline__ float dist2_exact( const float4& r, float px, float py, float pz )
{
  float vx = __fsub_rn( r.x, px ), vy = __fsub_rn( r.y, py ), vz = __fsub_rn( r.z, pz );
  return __fadd_rn( __fadd_rn( __fmul_rn( vx, vx ), __fmul_rn( vy, vy ) ), __fmul_rn( vz, vz ) );
}

// Conservative squared distance between a query and the axis-aligned bounding box of a set of points (cbox / nbox), as
// ordered bits: never above the dist2_exact of any point of the set.  lo <= p <= hi per axis for every point p, rounding
// is monotone and the operations and their order are dist2_exact's (no FMA: -fmad=false), so already the plain sum is a
// lower bound; 5 mantissa bits are dropped on top of that.  An empty set (lo = +inf, hi = -inf) gives +inf.
__device__ __forceinline__ uint32_t box_gap_bits( const float4& lo, const float4& hi, float px, float py, float pz )
{
  const float dx = fmaxf( fmaxf( __fsub_rn( lo.x, px ), __fsub_rn( px, hi.x ) ), 0.0f );
  const float dy = fmaxf( fmaxf( __fsub_rn( lo.y, py ), __fsub_rn( py, hi.y ) ), 0.0f );
  const float dz = fmaxf( fmaxf( __fsub_rn( lo.z, pz ), __fsub_rn( pz, hi.z ) ), 0.0f );
  const float s = __fadd_rn( __fadd_rn( __fmul_rn( dx, dx ), __fmul_rn( dy, dy ) ), __fmul_rn( dz, dz ) );
  return __float_as_uint( s ) & 0xffffffe0u;
}

__device__ __forceinline__ float dot3_exact( float ax, float ay, float az, float bx, float by, float bz )
{
  return __fadd_rn( __fadd_rn( __fmul_rn( ax, bx ), __fmul_rn( ay, by ) ), __fmul_rn( az, bz ) );
}

// ---------------------------------------------------------------------------------------------- cell window
// The block of cells one radius query examines (msh_hash_grid.h:1150-1225), warp-uniform.
struct CellWindow
{
  float qx, qy, qz;          // query minus grid min, float (:1159-1161)
  int c0x, c0y, c0z;         // the query's own cell (:1165-1167), may lie outside the grid
  int lox, loy, loz;         // clipped inclusive range
  int nx, ny, nz;            // clipped extents (0 when empty)
  int n_cells;               // min(nx*ny*nz, RSGPU_MAX_CELLS_PER_QUERY)  (:1213 cap)
};

__device__ __forceinline__ int clamp_ll( long long v, long long lo, long long hi )
{
  return (int)( v < lo ? lo : ( v > hi ? hi : v ) );
}

__device__ __forceinline__ CellWindow make_window( const GridView& g, float px, float py, float pz, double radius )
{
  CellWindow w;
  w.qx = __fsub_rn( px, g.mnx ); w.qy = __fsub_rn( py, g.mny ); w.qz = __fsub_rn( pz, g.mnz );
  const double dq[3] = { (double)w.qx, (double)w.qy, (double)w.qz };
  int c0[3], lo[3], hi[3];
#pragma unroll
  for( int a = 0; a < 3; ++a )
  {
    // (int64)( q * inv_cell ) of the reference (:1165-1183); the saturating 32-bit conversion followed by the clamps
    // below gives the same cell numbers as the 64-bit one for every finite input
    c0[a] = __double2int_rz( __dmul_rn( dq[a], g.inv_cell ) );
    hi[a] = __double2int_rz( __dmul_rn( __dadd_rn( dq[a], radius ), g.inv_cell ) );
    lo[a] = __double2int_rz( __dmul_rn( __dsub_rn( dq[a], radius ), g.inv_cell ) );
  }
  const int big = 1 << 28;
  w.c0x = min( max( c0[0], -big ), big ); w.c0y = min( max( c0[1], -big ), big ); w.c0z = min( max( c0[2], -big ), big );
  int hx, hy, hz;
  w.lox = min( max( lo[0], 0 ), g.W ); hx = min( max( hi[0], -1 ), g.W - 1 );
  w.loy = min( max( lo[1], 0 ), g.H ); hy = min( max( hi[1], -1 ), g.H - 1 );
  w.loz = min( max( lo[2], 0 ), g.D ); hz = min( max( hi[2], -1 ), g.D - 1 );
  w.nx = hx - w.lox + 1; w.ny = hy - w.loy + 1; w.nz = hz - w.loz + 1;
  if( w.nx <= 0 || w.ny <= 0 || w.nz <= 0 ) { w.nx = w.ny = w.nz = 0; }
  long long n = (long long)w.nx * w.ny * w.nz;
  w.n_cells = n > RSGPU_MAX_CELLS_PER_QUERY ? RSGPU_MAX_CELLS_PER_QUERY : (int)n;
  return w;
}

// one axis of the cell-to-query gap (:1196-1198): float of a double expression, 0 for the query's own slab
__device__ __forceinline__ float axis_gap( int c, int c0, float q, double cell )
{
  if( c < c0 ) { return (float)__dsub_rn( (double)q, __dmul_rn( (double)( c + 1 ), cell ) ); }
  if( c > c0 ) { return (float)__dsub_rn( __dmul_rn( (double)c, cell ), (double)q ); }
  return 0.0f;
}

// Cell number `e` (enumeration order z-outer, y, x-inner like the reference) of a window: its point range
// and the squared gap between the query and the cell (:1221).  Empty / out-of-window cells give s == e.
__device__ __forceinline__ void window_cell( const GridView& g, const CellWindow& w, int e, uint32_t& s,
                                             uint32_t& t, float& gap2 )
{
  s = t = 0; gap2 = __int_as_float( RS_INF_BITS );
  if( e >= w.n_cells ) { return; }
  int ix = e % w.nx; int r = e / w.nx; int iy = r % w.ny; int iz = r / w.ny;
  int cx = w.lox + ix, cy = w.loy + iy, cz = w.loz + iz;
  size_t id = ( (size_t)cz * g.H + cy ) * g.W + cx;
  s = __ldg( g.cell_start + id ); t = __ldg( g.cell_start + id + 1 );
  float gz = axis_gap( cz, w.c0z, w.qz, g.cell ), gy = axis_gap( cy, w.c0y, w.qy, g.cell ), gx = axis_gap( cx, w.c0x, w.qx, g.cell );
  gap2 = __fadd_rn( __fadd_rn( __fmul_rn( gz, gz ), __fmul_rn( gy, gy ) ), __fmul_rn( gx, gx ) );
}

#endif // __CUDACC__
