// Batched radius / k-NN search on the GPU hash grid: replaces msh_hash_grid_radius_search
// (reference lib/msh/msh_hash_grid.h:1090-1259) and msh_hash_grid_knn_search (:1294-1450).
//
// One warp per query.  The 32 lanes first resolve the (up to 512, usually 27) cells of the query's window in
// parallel — one coalesced pair of cell_start loads each instead of the reference's per-cell hash probe —
// then sweep the cells nearest-first reading 32 consecutive 16-byte records per step (one 512-byte coalesced
// request).  The k best are kept as a sorted list of 64-bit keys (dist^2 bits << 32 | original index) spread
// over the warp's registers, so the reference's per-query heap and final sort (:796-824, :579-703) disappear:
// rows come out ascending and ties are broken by the lower original index, deterministically.
#include "rsgpu_internal.cuh"
#include "warplist.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace rs;

namespace
{
template <int EPL>
__device__ __forceinline__ void sweep_cell( const GridView& g, uint32_t cs, uint32_t ce, float px, float py, float pz,
                                            bool use_radius, float r2f, int k, int lane, WarpList<EPL>& list,
                                            unsigned long long& thr, uint32_t& seen )
{
  for( uint32_t p0 = cs; p0 < ce; p0 += 32 )
  {
    uint32_t p = p0 + lane;
    unsigned long long key = KEY_INF;
    bool ok = false;
    if( p < ce )
    {
      float4 rec = __ldg( g.recs + p );
      float d2 = dist2_exact( rec, px, py, pz );
      ok = !use_radius || d2 < r2f;
      key = ( (unsigned long long)__float_as_uint( d2 ) << 32 ) | __float_as_uint( rec.w );
    }
    unsigned in_range = __ballot_sync( RS_FULL, ok );
    seen += __popc( in_range );
    unsigned m = __ballot_sync( RS_FULL, ok && key < thr );
    while( m )
    {
      // every candidate left in the mask beats the current k-th best: insert the first, then drop the ones the new k-th best
      // rules out in one vote (half of the candidates that pass the first vote never make the list - profiles/kernels_r02.md)
      int src = __ffs( m ) - 1; m &= m - 1;
      unsigned long long x = __shfl_sync( RS_FULL, key, src );
      list.insert( x, lane );
      thr = list.get( k - 1 );
      m &= __ballot_sync( RS_FULL, key < thr );
    }
  }
}

template <int EPL>
__device__ __forceinline__ void write_row( const WarpList<EPL>& list, int lane, uint32_t count, size_t row, int k,
                                           float* __restrict__ out_d2, int32_t* __restrict__ out_idx )
{
#pragma unroll
  for( int s = 0; s < EPL; ++s )
  {
    uint32_t j = lane * EPL + s;
    if( j < count )
    {
      out_d2[row * k + j] = __uint_as_float( (uint32_t)( list.v[s] >> 32 ) );
      out_idx[row * k + j] = (int32_t)(uint32_t)( list.v[s] & 0xffffffffull );
    }
  }
}

template <int EPL>
__global__ void __launch_bounds__( 128 ) radius_search_kernel( GridView g, const float* __restrict__ q, size_t nq, double radius,
                                                               float r2f, int k, float* __restrict__ out_d2,
                                                               int32_t* __restrict__ out_idx, unsigned long long* __restrict__ out_nn,
                                                               unsigned long long* __restrict__ total )
{
  const int lane = threadIdx.x & 31;
  size_t warp = ( blockIdx.x * (size_t)blockDim.x + threadIdx.x ) >> 5;
  size_t n_warps = ( gridDim.x * (size_t)blockDim.x ) >> 5;
  unsigned long long local_total = 0;
  for( size_t qi = warp; qi < nq; qi += n_warps )
  {
    float px = __ldg( q + 3 * qi ), py = __ldg( q + 3 * qi + 1 ), pz = __ldg( q + 3 * qi + 2 );
    CellWindow w = make_window( g, px, py, pz, radius );
    WarpList<EPL> list; list.init();
    unsigned long long thr = KEY_INF;
    uint32_t seen = 0;
    for( int base = 0; base < w.n_cells; base += 32 )
    {
      uint32_t s, t; float gap2;
      window_cell( g, w, base + lane, s, t, gap2 );
      uint32_t gbits = ( s < t && gap2 < r2f ) ? __float_as_uint( gap2 ) : RS_INF_BITS;
      while( true )
      {
        uint32_t gmin = __reduce_min_sync( RS_FULL, gbits );
        // nothing left, or the list is full and no remaining cell can beat its last entry (:1232-1236)
        if( gmin == RS_INF_BITS || ( thr != KEY_INF && gmin >= (uint32_t)( thr >> 32 ) ) ) { break; }
        int src = __ffs( __ballot_sync( RS_FULL, gbits == gmin ) ) - 1;
        uint32_t cs = __shfl_sync( RS_FULL, s, src ), ce = __shfl_sync( RS_FULL, t, src );
        if( lane == src ) { gbits = RS_INF_BITS; }
        sweep_cell<EPL>( g, cs, ce, px, py, pz, true, r2f, k, lane, list, thr, seen );
      }
    }
    uint32_t count = seen < (uint32_t)k ? seen : (uint32_t)k;
    write_row<EPL>( list, lane, count, qi, k, out_d2, out_idx );
    if( lane == 0 && out_nn ) { out_nn[qi] = count; }
    local_total += count;
  }
  if( lane == 0 && local_total ) { atomicAdd( total, local_total ); }
}

// ---- sub-cell variant: grids whose cells hold hundreds of points (the reference fixes the cell edge at 2 x the build radius,
// so a 10 M-point scan searched with r = 0.10 m has ~500 points per cell and ~2 000 candidates per query, of which the
// k = 64 nearest lie within ~2 cm).  The warp-per-query kernel above is then bound by instruction issue on candidates that
// cannot make the list (profiles/kernels_r02.md: 8 160 warp instructions per query, 72 % issue-active, 14 % of the DRAM peak).
// Here every cell's records are kept a second time ordered by 4x4x4 sub-cell; a query ranks the 64 sub-cells of a window
// cell by their gap (two per lane), sweeps them nearest first and stops at the first one that cannot beat the list's
// last entry - the reference's own pruning rule (msh_hash_grid.h:1232-1236) one level down.  Same rows as the flat kernel:
// every pruned sub-cell holds only points that are out of range or behind the k-th best (its box is shrunk by a
// margin far above the float error of the sub-cell assignment, so the gap is a true lower bound).
constexpr int SUB_N = 4;                    // sub-cells per axis
constexpr int SUB_CELLS = SUB_N * SUB_N * SUB_N;

__global__ void sub_key_kernel( const float4* __restrict__ recs, const uint32_t* __restrict__ cell_start, size_t n_cells, int W, int H,
                                float mnx, float mny, float mnz, float inv_cell, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                uint32_t* __restrict__ counts )
{
  // one thread per cell walks its records (cells hold at most a few thousand points; the build runs once per grid)
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if( c >= n_cells ) { return; }
  const uint32_t s = cell_start[c], t = cell_start[c + 1];
  if( s == t ) { return; }
  const int cx = (int)( c % (size_t)W ), cy = (int)( ( c / (size_t)W ) % (size_t)H ), cz = (int)( c / ( (size_t)W * H ) );
  for( uint32_t p = s; p < t; ++p )
  {
    const float4 r = recs[p];
    const int sx = min( max( (int)floorf( ( ( r.x - mnx ) * inv_cell - (float)cx ) * (float)SUB_N ), 0 ), SUB_N - 1 );
    const int sy = min( max( (int)floorf( ( ( r.y - mny ) * inv_cell - (float)cy ) * (float)SUB_N ), 0 ), SUB_N - 1 );
    const int sz = min( max( (int)floorf( ( ( r.z - mnz ) * inv_cell - (float)cz ) * (float)SUB_N ), 0 ), SUB_N - 1 );
    const uint32_t sub = (uint32_t)( ( sz * SUB_N + sy ) * SUB_N + sx );
    keys[p] = (uint32_t)c * SUB_CELLS + sub;
    vals[p] = p;
    atomicAdd( counts + c * SUB_CELLS + sub, 1u );
  }
}
__global__ void sub_gather_kernel( const float4* __restrict__ recs, const uint32_t* __restrict__ order, int n, float4* __restrict__ out )
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i < n ) { out[i] = recs[order[i]]; }
}

template <int EPL>
__global__ void __launch_bounds__( 128 ) radius_search_sub_kernel( GridView g, const float4* __restrict__ sub_recs, const uint32_t* __restrict__ sub_off,
                                                                   const float* __restrict__ q, size_t nq, double radius, float r2f, int k,
                                                                   float* __restrict__ out_d2, int32_t* __restrict__ out_idx,
                                                                   unsigned long long* __restrict__ out_nn, unsigned long long* __restrict__ total )
{
  const int lane = threadIdx.x & 31;
  size_t warp = ( blockIdx.x * (size_t)blockDim.x + threadIdx.x ) >> 5;
  size_t n_warps = ( gridDim.x * (size_t)blockDim.x ) >> 5;
  unsigned long long local_total = 0;
  GridView gs = g; gs.recs = sub_recs; // sweep_cell reads g.recs
  const float cellf = (float)g.cell, subf = cellf * ( 1.0f / (float)SUB_N );
  // float error of the sub-cell assignment grows with the cell index (cancellation in q / cell - c): 1e-3 cells covers 2 000 cells per axis
  const float margin = cellf * fmaxf( 1e-3f, 5e-7f * (float)max( g.W, max( g.H, g.D ) ) );
  const uint32_t r2bits = __float_as_uint( r2f );
  for( size_t qi = warp; qi < nq; qi += n_warps )
  {
    float px = __ldg( q + 3 * qi ), py = __ldg( q + 3 * qi + 1 ), pz = __ldg( q + 3 * qi + 2 );
    CellWindow w = make_window( g, px, py, pz, radius );
    WarpList<EPL> list; list.init();
    unsigned long long thr = KEY_INF;
    uint32_t seen = 0;
    for( int base = 0; base < w.n_cells; base += 32 )
    {
      uint32_t s, t; float gap2;
      window_cell( g, w, base + lane, s, t, gap2 );
      uint32_t gbits = ( s < t && gap2 < r2f ) ? __float_as_uint( gap2 ) : RS_INF_BITS;
      while( true )
      {
        uint32_t gmin = __reduce_min_sync( RS_FULL, gbits );
        if( gmin == RS_INF_BITS || ( thr != KEY_INF && gmin >= (uint32_t)( thr >> 32 ) ) ) { break; }
        int src = __ffs( __ballot_sync( RS_FULL, gbits == gmin ) ) - 1;
        if( lane == src ) { gbits = RS_INF_BITS; }
        // the chosen cell: its index in the window -> grid coordinates (enumeration of window_cell: x fastest)
        const int e = base + src;
        const int ix = e % w.nx, rr = e / w.nx, iy = rr % w.ny, iz = rr / w.ny;
        const int cx = w.lox + ix, cy = w.loy + iy, cz = w.loz + iz;
        const size_t cid = ( (size_t)cz * g.H + cy ) * g.W + cx;
        // the cell's 64 sub-cells, two per lane: range and conservative squared gap
        uint32_t ss[2], se[2], sg[2];
#pragma unroll
        for( int h = 0; h < 2; ++h )
        {
          const int sub = lane + 32 * h;
          ss[h] = __ldg( sub_off + cid * SUB_CELLS + sub ); se[h] = __ldg( sub_off + cid * SUB_CELLS + sub + 1 );
          const int sx = sub % SUB_N, sy = ( sub / SUB_N ) % SUB_N, sz = sub / ( SUB_N * SUB_N );
          const float lx = (float)cx * cellf + (float)sx * subf, ly = (float)cy * cellf + (float)sy * subf, lz = (float)cz * cellf + (float)sz * subf;
          const float dx = fmaxf( fmaxf( lx - w.qx, w.qx - ( lx + subf ) ) - margin, 0.0f );
          const float dy = fmaxf( fmaxf( ly - w.qy, w.qy - ( ly + subf ) ) - margin, 0.0f );
          const float dz = fmaxf( fmaxf( lz - w.qz, w.qz - ( lz + subf ) ) - margin, 0.0f );
          const uint32_t gb = __float_as_uint( dx * dx + dy * dy + dz * dz ) & 0xffffff00u; // rounded down: still a lower bound
          sg[h] = ( ss[h] < se[h] && gb < r2bits ) ? gb : RS_INF_BITS;
        }
        while( true )
        {
          const uint32_t m2 = __reduce_min_sync( RS_FULL, min( sg[0], sg[1] ) );
          if( m2 == RS_INF_BITS || ( thr != KEY_INF && m2 >= (uint32_t)( thr >> 32 ) ) ) { break; }
          const unsigned b0 = __ballot_sync( RS_FULL, sg[0] == m2 ), b1 = __ballot_sync( RS_FULL, sg[1] == m2 );
          const int h = b0 ? 0 : 1;
          const int sl = __ffs( b0 ? b0 : b1 ) - 1;
          const uint32_t cs = __shfl_sync( RS_FULL, h == 0 ? ss[0] : ss[1], sl ), ce = __shfl_sync( RS_FULL, h == 0 ? se[0] : se[1], sl );
          if( lane == sl ) { if( h == 0 ) { sg[0] = RS_INF_BITS; } else { sg[1] = RS_INF_BITS; } }
          sweep_cell<EPL>( gs, cs, ce, px, py, pz, true, r2f, k, lane, list, thr, seen );
        }
      }
    }
    uint32_t count = seen < (uint32_t)k ? seen : (uint32_t)k;
    write_row<EPL>( list, lane, count, qi, k, out_d2, out_idx );
    if( lane == 0 && out_nn ) { out_nn[qi] = count; }
    local_total += count;
  }
  if( lane == 0 && local_total ) { atomicAdd( total, local_total ); }
}

// One query per THREAD, for small k on sparse windows (a few dozen candidate points per query): the warp-per-query
// kernel above spends ~300 instructions of election / reduction / shuffle per query there, while the whole job is a
// handful of distance tests.  Same result definition: the k nearest points with dist^2 < r^2, ascending by
// (dist^2, original index); the sorted list lives in K registers of the thread (insertion = K compare-selects).
template <int K>
__global__ void __launch_bounds__( 128 ) radius_search_lane_kernel( GridView g, const float* __restrict__ q, size_t nq, double radius,
                                                                    float r2f, int k, float* __restrict__ out_d2,
                                                                    int32_t* __restrict__ out_idx, unsigned long long* __restrict__ out_nn,
                                                                    unsigned long long* __restrict__ total )
{
  unsigned long long local_total = 0;
  const uint32_t r2bits = __float_as_uint( r2f );
  for( size_t qi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; qi < nq; qi += gridDim.x * (size_t)blockDim.x )
  {
    const float px = __ldg( q + 3 * qi ), py = __ldg( q + 3 * qi + 1 ), pz = __ldg( q + 3 * qi + 2 );
    const CellWindow w = make_window( g, px, py, pz, radius );
    unsigned long long keys[K];
#pragma unroll
    for( int s = 0; s < K; ++s ) { keys[s] = KEY_INF; }
    unsigned long long thr = KEY_INF; // keys[K - 1]
    uint32_t seen = 0;
    for( int e = 0; e < w.n_cells; ++e )
    {
      uint32_t cs, ce; float gap2;
      window_cell( g, w, e, cs, ce, gap2 );
      const uint32_t gbits = __float_as_uint( gap2 );
      // empty / out of range, or the list is full and the cell cannot beat its last entry (:1232-1236)
      if( cs >= ce || !( gbits < r2bits ) || ( thr != KEY_INF && gbits >= (uint32_t)( thr >> 32 ) ) ) { continue; }
      for( uint32_t p = cs; p < ce; ++p )
      {
        const float4 rec = __ldg( g.recs + p );
        const uint32_t db = __float_as_uint( dist2_exact( rec, px, py, pz ) );
        if( db < r2bits )
        {
          ++seen;
          const unsigned long long x = ( (unsigned long long)db << 32 ) | __float_as_uint( rec.w );
          if( x < thr )
          {
#pragma unroll
            for( int s = K - 1; s >= 0; --s )
            {
              const unsigned long long prev = s > 0 ? keys[s - 1] : 0ull;
              keys[s] = ( keys[s] <= x ) ? keys[s] : ( prev <= x ? x : prev );
            }
            thr = keys[K - 1]; // K >= k: pruning by the K-th best is merely a little weaker than by the k-th
          }
        }
      }
    }
    const uint32_t count = seen < (uint32_t)k ? seen : (uint32_t)k;
#pragma unroll
    for( int s = 0; s < K; ++s )
    {
      if( (uint32_t)s < count )
      {
        out_d2[qi * k + s] = __uint_as_float( (uint32_t)( keys[s] >> 32 ) );
        out_idx[qi * k + s] = (int32_t)(uint32_t)( keys[s] & 0xffffffffull );
      }
    }
    if( out_nn ) { out_nn[qi] = count; }
    local_total += count;
  }
  for( int o = 16; o > 0; o >>= 1 ) { local_total += __shfl_down_sync( RS_FULL, local_total, o ); }
  if( ( threadIdx.x & 31 ) == 0 && local_total ) { atomicAdd( total, local_total ); }
}

template <int K>
int launch_search_lane( const GridView& g, const float* d_q, size_t nq, double radius, float r2f, int k, float* d_d2, int32_t* d_idx,
                        unsigned long long* d_nn, unsigned long long* d_total )
{
  size_t blocks = ( nq + 127 ) / 128;
  const size_t max_blocks = 148 * 32;
  if( blocks > max_blocks ) { blocks = max_blocks; }
  radius_search_lane_kernel<K><<<(unsigned)blocks, 128, 0, rt().stream>>>( g, d_q, nq, radius, r2f, k, d_d2, d_idx, d_nn, d_total );
  RS_CHECK_LAUNCH();
  return RSGPU_OK;
}

// msh_hash_grid_knn_search: shells of cells around the query's cell are opened layer by layer and the search
// stops after the layer FOLLOWING the one that first filled the list (:1427-1429); cells pruned inside a
// layer cannot hold a better point, so the result is the k nearest points of the cube of half-width L + 1
// cells, L = first layer at which the cube holds >= k points.
template <int EPL>
__global__ void __launch_bounds__( 128 ) knn_search_kernel( GridView g, const float* __restrict__ q, size_t nq, int k,
                                                            float* __restrict__ out_d2, int32_t* __restrict__ out_idx,
                                                            unsigned long long* __restrict__ out_nn,
                                                            unsigned long long* __restrict__ total )
{
  const int lane = threadIdx.x & 31;
  size_t warp = ( blockIdx.x * (size_t)blockDim.x + threadIdx.x ) >> 5;
  size_t n_warps = ( gridDim.x * (size_t)blockDim.x ) >> 5;
  unsigned long long local_total = 0;
  const int max_layer = g.W + g.H + g.D;
  for( size_t qi = warp; qi < nq; qi += n_warps )
  {
    float px = __ldg( q + 3 * qi ), py = __ldg( q + 3 * qi + 1 ), pz = __ldg( q + 3 * qi + 2 );
    long long c0[3];
    c0[0] = __double2ll_rz( __dmul_rn( (double)__fsub_rn( px, g.mnx ), g.inv_cell ) );
    c0[1] = __double2ll_rz( __dmul_rn( (double)__fsub_rn( py, g.mny ), g.inv_cell ) );
    c0[2] = __double2ll_rz( __dmul_rn( (double)__fsub_rn( pz, g.mnz ), g.inv_cell ) );
    const long long big = 1 << 28;
    int cx = clamp_ll( c0[0], -big, big ), cy = clamp_ll( c0[1], -big, big ), cz = clamp_ll( c0[2], -big, big );
    // find L: grow the cube until it holds k points (row ranges are contiguous in the dense table)
    int L = 0;
    for( ; L <= max_layer; ++L )
    {
      int x0 = max( cx - L, 0 ), x1 = min( cx + L, g.W - 1 ), y0 = max( cy - L, 0 ), y1 = min( cy + L, g.H - 1 ),
          z0 = max( cz - L, 0 ), z1 = min( cz + L, g.D - 1 );
      uint32_t cnt = 0;
      if( x0 <= x1 && y0 <= y1 && z0 <= z1 )
      {
        int ny = y1 - y0 + 1, rows = ny * ( z1 - z0 + 1 );
        for( int r = lane; r < rows; r += 32 )
        {
          size_t rowbase = ( (size_t)( z0 + r / ny ) * g.H + ( y0 + r % ny ) ) * g.W;
          cnt += __ldg( g.cell_start + rowbase + x1 + 1 ) - __ldg( g.cell_start + rowbase + x0 );
        }
      }
      cnt = __reduce_add_sync( RS_FULL, cnt );
      if( cnt >= (uint32_t)k ) { break; }
      if( x0 == 0 && y0 == 0 && z0 == 0 && x1 == g.W - 1 && y1 == g.H - 1 && z1 == g.D - 1 ) { break; } // whole grid seen
    }
    int R = L + 1; // the cube actually searched
    WarpList<EPL> list; list.init();
    unsigned long long thr = KEY_INF;
    uint32_t seen = 0;
    int x0 = max( cx - R, 0 ), x1 = min( cx + R, g.W - 1 ), y0 = max( cy - R, 0 ), y1 = min( cy + R, g.H - 1 ),
        z0 = max( cz - R, 0 ), z1 = min( cz + R, g.D - 1 );
    if( x0 <= x1 && y0 <= y1 && z0 <= z1 )
    {
      int ny = y1 - y0 + 1, rows = ny * ( z1 - z0 + 1 );
      for( int r = 0; r < rows; ++r )
      {
        size_t rowbase = ( (size_t)( z0 + r / ny ) * g.H + ( y0 + r % ny ) ) * g.W;
        uint32_t cs = __ldg( g.cell_start + rowbase + x0 ), ce = __ldg( g.cell_start + rowbase + x1 + 1 );
        sweep_cell<EPL>( g, cs, ce, px, py, pz, false, 0.f, k, lane, list, thr, seen );
      }
    }
    uint32_t count = seen < (uint32_t)k ? seen : (uint32_t)k;
    write_row<EPL>( list, lane, count, qi, k, out_d2, out_idx );
    if( lane == 0 && out_nn ) { out_nn[qi] = count; }
    local_total += count;
  }
  if( lane == 0 && local_total ) { atomicAdd( total, local_total ); }
}

// rspf_compute_neighborhood's candidate edges (rs_pointcloud_filters.cpp:693-708): k <= 32 nearest within the
// radius of every vertex plus the edge weight, one warp per vertex
__global__ void __launch_bounds__( 128 ) neighborhood_kernel( GridView g, const float* __restrict__ pos, const float* __restrict__ nor, int n,
                                                              double radius, float r2f, int k, float radius_sq, float dist_exp,
                                                              float angle_exp, int32_t* __restrict__ nbr, float* __restrict__ wgt )
{
  const int lane = threadIdx.x & 31;
  size_t warp = ( blockIdx.x * (size_t)blockDim.x + threadIdx.x ) >> 5;
  size_t n_warps = ( gridDim.x * (size_t)blockDim.x ) >> 5;
  for( size_t qi = warp; qi < (size_t)n; qi += n_warps )
  {
    float px = __ldg( pos + 3 * qi ), py = __ldg( pos + 3 * qi + 1 ), pz = __ldg( pos + 3 * qi + 2 );
    CellWindow w = make_window( g, px, py, pz, radius );
    WarpList<1> list; list.init();
    unsigned long long thr = KEY_INF;
    uint32_t seen = 0;
    for( int base = 0; base < w.n_cells; base += 32 )
    {
      uint32_t s, t; float gap2;
      window_cell( g, w, base + lane, s, t, gap2 );
      uint32_t gbits = ( s < t && gap2 < r2f ) ? __float_as_uint( gap2 ) : RS_INF_BITS;
      while( true )
      {
        uint32_t gmin = __reduce_min_sync( RS_FULL, gbits );
        if( gmin == RS_INF_BITS || ( thr != KEY_INF && gmin >= (uint32_t)( thr >> 32 ) ) ) { break; }
        int src = __ffs( __ballot_sync( RS_FULL, gbits == gmin ) ) - 1;
        uint32_t cs = __shfl_sync( RS_FULL, s, src ), ce = __shfl_sync( RS_FULL, t, src );
        if( lane == src ) { gbits = RS_INF_BITS; }
        sweep_cell<1>( g, cs, ce, px, py, pz, true, r2f, k, lane, list, thr, seen );
      }
    }
    uint32_t count = seen < (uint32_t)k ? seen : (uint32_t)k;
    if( lane < k )
    {
      int32_t id = -1; float wv = 0.f;
      if( (uint32_t)lane < count )
      {
        float d2 = __uint_as_float( (uint32_t)( list.v[0] >> 32 ) );
        id = (int32_t)(uint32_t)( list.v[0] & 0xffffffffull );
        float dot = dot3_exact( __ldg( nor + 3 * qi ), __ldg( nor + 3 * qi + 1 ), __ldg( nor + 3 * qi + 2 ),
                                __ldg( nor + 3 * (size_t)id ), __ldg( nor + 3 * (size_t)id + 1 ), __ldg( nor + 3 * (size_t)id + 2 ) );
        dot = dot < 0.0f ? 0.0f : ( dot > 1.0f ? 1.0f : dot ); // msh_clamp (:707)
        float dist_cost = (float)( 1.0 - pow( (double)d2 / ( 4.0 * (double)radius_sq ), (double)dist_exp ) ); // (:706)
        float norm_cost = (float)pow( (double)dot, (double)angle_exp ); // float pow overload (:707)
        wv = __fmul_rn( dist_cost, norm_cost );
      }
      nbr[qi * k + lane] = id; wgt[qi * k + lane] = wv;
    }
  }
}

// census of what a radius search reads by the reference's data layout: per query the non-empty cells
// overlapping its window (B) and the points stored in them (C), no early-out credit
__global__ void __launch_bounds__( 256 ) search_census_kernel( GridView g, const float* __restrict__ q, size_t nq, double radius,
                                                               unsigned long long* __restrict__ counts )
{
  unsigned long long nB = 0, nC = 0;
  for( size_t qi = blockIdx.x * (size_t)blockDim.x + threadIdx.x; qi < nq; qi += gridDim.x * (size_t)blockDim.x )
  {
    CellWindow w = make_window( g, __ldg( q + 3 * qi ), __ldg( q + 3 * qi + 1 ), __ldg( q + 3 * qi + 2 ), radius );
    for( int e = 0; e < w.n_cells; ++e )
    {
      uint32_t s, t; float gap2;
      window_cell( g, w, e, s, t, gap2 );
      if( s < t ) { nB += 1; nC += t - s; }
    }
  }
  for( int o = 16; o > 0; o >>= 1 ) { nB += __shfl_down_sync( RS_FULL, nB, o ); nC += __shfl_down_sync( RS_FULL, nC, o ); }
  if( ( threadIdx.x & 31 ) == 0 ) { atomicAdd( counts + 0, nB ); atomicAdd( counts + 1, nC ); }
}

template <int EPL>
int launch_search( bool knn, const GridView& g, const float* d_q, size_t nq, double radius, float r2f, int k, float* d_d2,
                   int32_t* d_idx, unsigned long long* d_nn, unsigned long long* d_total )
{
  size_t warps = nq;
  size_t blocks = ( warps + 3 ) / 4;
  const size_t max_blocks = 148 * 64;
  if( blocks > max_blocks ) { blocks = max_blocks; }
  if( knn ) { knn_search_kernel<EPL><<<(unsigned)blocks, 128, 0, rt().stream>>>( g, d_q, nq, k, d_d2, d_idx, d_nn, d_total ); }
  else { radius_search_kernel<EPL><<<(unsigned)blocks, 128, 0, rt().stream>>>( g, d_q, nq, radius, r2f, k, d_d2, d_idx, d_nn, d_total ); }
  RS_CHECK_LAUNCH();
  return RSGPU_OK;
}

// the sub-cell ordered copy of a grid's records (see radius_search_sub_kernel); 1 = available, -1 = not for this grid
int ensure_sub_cells( const rsgpu_grid_t* grid )
{
  std::lock_guard<std::mutex> lk( grid->sub_mu );
  if( grid->sub_state != 0 ) { return grid->sub_state; }
  const size_t n_cells = (size_t)grid->info.width * grid->info.height * grid->info.depth;
  const size_t n = (size_t)grid->info.n_pts;
  if( n == 0 || n_cells * SUB_CELLS + 1 > ( (size_t)1 << 28 ) ) { grid->sub_state = -1; return -1; }
  cudaStream_t st = rt().stream;
  const GridView g = grid->view();
  DevBuf<uint32_t> k0, k1, v0, v1, counts;
  if( k0.alloc( n ) != cudaSuccess || k1.alloc( n ) != cudaSuccess || v0.alloc( n ) != cudaSuccess || v1.alloc( n ) != cudaSuccess ||
      counts.alloc( n_cells * SUB_CELLS + 1 ) != cudaSuccess || grid->sub_off.alloc( n_cells * SUB_CELLS + 1 ) != cudaSuccess ||
      grid->sub_recs.alloc( n ) != cudaSuccess )
  {
    cudaGetLastError(); grid->sub_recs.release(); grid->sub_off.release(); grid->sub_state = -1; return -1; // no room: the flat kernel serves
  }
  cudaMemsetAsync( counts.p, 0, sizeof( uint32_t ) * ( n_cells * SUB_CELLS + 1 ), st );
  sub_key_kernel<<<(unsigned)( ( n_cells + 127 ) / 128 ), 128, 0, st>>>( g.recs, g.cell_start, n_cells, g.W, g.H, g.mnx, g.mny, g.mnz, (float)g.inv_cell, k0.p, v0.p, counts.p );
  count_launch();
  size_t scan_bytes = 0, sort_bytes = 0;
  int end_bit = 1;
  while( end_bit < 32 && ( (size_t)1 << end_bit ) < n_cells * SUB_CELLS ) { ++end_bit; }
  cub::DeviceScan::ExclusiveSum( nullptr, scan_bytes, counts.p, grid->sub_off.p, (int64_t)( n_cells * SUB_CELLS + 1 ), st );
  cub::DeviceRadixSort::SortPairs( nullptr, sort_bytes, k0.p, k1.p, v0.p, v1.p, (int64_t)n, 0, end_bit, st );
  DevBuf<unsigned char> tmp;
  if( tmp.alloc( std::max( scan_bytes, sort_bytes ) ) != cudaSuccess ) { cudaGetLastError(); grid->sub_recs.release(); grid->sub_off.release(); grid->sub_state = -1; return -1; }
  cub::DeviceScan::ExclusiveSum( tmp.p, scan_bytes, counts.p, grid->sub_off.p, (int64_t)( n_cells * SUB_CELLS + 1 ), st );
  // stable: records with equal (cell, sub-cell) keep their order = ascending original index (the reference's bin order)
  cub::DeviceRadixSort::SortPairs( tmp.p, sort_bytes, k0.p, k1.p, v0.p, v1.p, (int64_t)n, 0, end_bit, st );
  sub_gather_kernel<<<(unsigned)( ( n + 255 ) / 256 ), 256, 0, st>>>( g.recs, v1.p, (int)n, grid->sub_recs.p );
  count_launch();
  const cudaError_t e = cudaStreamSynchronize( st ); // other lanes / streams may search this grid next
  if( e != cudaSuccess || cudaGetLastError() != cudaSuccess ) { grid->sub_recs.release(); grid->sub_off.release(); grid->sub_state = -1; return -1; }
  grid->sub_state = 1;
  return 1;
}

template <int EPL>
int launch_search_sub( const rsgpu_grid_t* grid, const GridView& g, const float* d_q, size_t nq, double radius, float r2f, int k, float* d_d2,
                       int32_t* d_idx, unsigned long long* d_nn, unsigned long long* d_total )
{
  size_t blocks = ( nq + 3 ) / 4;
  const size_t max_blocks = 148 * 64;
  if( blocks > max_blocks ) { blocks = max_blocks; }
  radius_search_sub_kernel<EPL><<<(unsigned)blocks, 128, 0, rt().stream>>>( g, grid->sub_recs.p, grid->sub_off.p, d_q, nq, radius, r2f, k, d_d2, d_idx, d_nn, d_total );
  RS_CHECK_LAUNCH();
  return RSGPU_OK;
}

int search_dev( bool knn, const rsgpu_grid_t* grid, const float* d_q, size_t nq, float radius, size_t k, float* d_d2,
                int32_t* d_idx, unsigned long long* d_nn, size_t* total )
{
  if( k == 0 || k > RSGPU_MAX_K )
  {
    return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu search: k must be in [1, RSGPU_MAX_K]" );
  }
  if( !knn && !( radius > 0.f ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_radius_search: radius must be > 0" ); }
  DevBuf<unsigned long long> d_total;
  RS_CUDA( d_total.alloc( 1 ) );
  RS_CUDA( cudaMemsetAsync( d_total.p, 0, 8, rt().stream ) );
  if( nq > 0 && grid->info.n_pts > 0 )
  {
    ProfScope prof( "search" );
    GridView g = grid->view();
    double r = radius;
    float r2f = (float)( r * r ); // double product narrowed when handed down (:1111, 828)
    int kk = (int)k, s;
    // thread-per-query kernel for small k when a window holds few candidate points (sparse cells); the warp-per-query
    // kernel otherwise.  RSGPU_SEARCH_IMPL=lane|warp forces one of them.
    const std::string oimpl = option( "search_impl" );
    const int impl_env = oimpl == "lane" ? 1 : ( oimpl == "warp" ? 2 : 0 );
    const double pts_per_bin = grid->info.n_bins > 0 ? (double)grid->info.n_pts / (double)grid->info.n_bins : 0.0;
    const double cells_axis = 2.0 * r / grid->info.cell_size + 1.0; // expected cells per axis under the window
    const double est_candidates = pts_per_bin * cells_axis * cells_axis; // surfaces: ~2-D occupancy
    // measured crossovers on B200 (profiles/nn_sweep_r01.md): ~300 candidate points per query for k <= 4, ~60 for k <= 16
    const bool lane = !knn && kk <= 16 && ( impl_env == 1 || ( impl_env == 0 && est_candidates <= ( kk <= 4 ? 320.0 : 64.0 ) ) );
    // cells with hundreds of points (cell edge = 2 x the build radius, fixed by the reference's API): rank sub-cells first.
    // "search_sub" = "0" / "1" forces the choice; otherwise from 256 points per non-empty cell on (surfaces fill ~16 of a cell's
    // 64 sub-cells: below that a sub-cell holds fewer than 16 points and its sweep step runs with most lanes idle - measured
    // on the 10 M-point sweep: 1.5 - 2.2 x faster at 500 points per cell, 0.8 - 0.97 x at 126), once per grid.
    const std::string osub = option( "search_sub" );
    const bool want_sub = !knn && !lane && kk <= 512 && ( osub == "1" || ( osub != "0" && pts_per_bin >= 256.0 && nq >= 4096 ) );
    const bool use_sub = want_sub && ensure_sub_cells( grid ) == 1;
    if( use_sub )
    {
      if( kk <= 32 ) { s = launch_search_sub<1>( grid, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
      else if( kk <= 64 ) { s = launch_search_sub<2>( grid, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
      else if( kk <= 128 ) { s = launch_search_sub<4>( grid, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
      else if( kk <= 256 ) { s = launch_search_sub<8>( grid, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
      else { s = launch_search_sub<16>( grid, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
    }
    else if( lane )
    {
      if( kk <= 1 ) { s = launch_search_lane<1>( g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
      else if( kk <= 4 ) { s = launch_search_lane<4>( g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
      else if( kk <= 8 ) { s = launch_search_lane<8>( g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
      else { s = launch_search_lane<16>( g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
    }
    else if( kk <= 32 ) { s = launch_search<1>( knn, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
    else if( kk <= 64 ) { s = launch_search<2>( knn, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
    else if( kk <= 128 ) { s = launch_search<4>( knn, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
    else if( kk <= 256 ) { s = launch_search<8>( knn, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
    else { s = launch_search<16>( knn, g, d_q, nq, r, r2f, kk, d_d2, d_idx, d_nn, d_total.p ); }
    RS_TRY( s );
  }
  else if( d_nn && nq > 0 ) { RS_CUDA( cudaMemsetAsync( d_nn, 0, sizeof( unsigned long long ) * nq, rt().stream ) ); }
  unsigned long long h = 0;
  RS_CUDA( cudaMemcpyAsync( &h, d_total.p, 8, cudaMemcpyDeviceToHost, rt().stream ) );
  RS_CUDA( rs::stream_sync( rt().stream ) );
  if( total ) { *total = (size_t)h; }
  return RSGPU_OK;
}

int search_host( bool knn, const rsgpu_grid_t* grid, rsgpu_search_desc_t* d, size_t* total )
{
  if( !grid || !d || !d->query_pts || !d->distances_sq || !d->indices )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu search: NULL grid / descriptor / buffer (the caller allocates every buffer)" );
  }
  RS_TRY( ensure_device() );
  size_t nq = d->n_query_pts, k = d->k;
  if( k == 0 || k > RSGPU_MAX_K ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu search: k must be in [1, RSGPU_MAX_K]" ); }
  DevBuf<float> dq, dd; DevBuf<int32_t> di; DevBuf<unsigned long long> dn;
  RS_CUDA( dq.alloc( nq * 3 ) ); RS_CUDA( dd.alloc( nq * k ) ); RS_CUDA( di.alloc( nq * k ) ); RS_CUDA( dn.alloc( nq ) );
  cudaStream_t st = rt().stream;
  if( nq ) { RS_CUDA( cudaMemcpyAsync( dq.p, d->query_pts, sizeof( float ) * 3 * nq, cudaMemcpyHostToDevice, st ) ); }
  RS_TRY( search_dev( knn, grid, dq.p, nq, d->radius, k, dd.p, di.p, dn.p, total ) );
  if( nq )
  {
    // rows are copied whole; entries past each row's count are unspecified in the reference as well
    RS_CUDA( cudaMemcpyAsync( d->distances_sq, dd.p, sizeof( float ) * nq * k, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( cudaMemcpyAsync( d->indices, di.p, sizeof( int32_t ) * nq * k, cudaMemcpyDeviceToHost, st ) );
    if( d->n_neighbors )
    {
      static_assert( sizeof( size_t ) == sizeof( unsigned long long ), "size_t must be 64-bit" );
      RS_CUDA( cudaMemcpyAsync( d->n_neighbors, dn.p, sizeof( size_t ) * nq, cudaMemcpyDeviceToHost, st ) );
    }
    RS_CUDA( rs::stream_sync( st ) );
  }
  return RSGPU_OK;
}
} // namespace

extern "C" {

int rsgpu_grid_radius_search( const rsgpu_grid_t* g, rsgpu_search_desc_t* d, size_t* total ) { return search_host( false, g, d, total ); }
int rsgpu_grid_knn_search( const rsgpu_grid_t* g, rsgpu_search_desc_t* d, size_t* total ) { return search_host( true, g, d, total ); }

int rsgpu_grid_radius_search_dev( const rsgpu_grid_t* g, rsgpu_search_desc_t* d, size_t* total )
{
  if( !g || !d || !d->query_pts || !d->distances_sq || !d->indices ) { return fail( RSGPU_ERR_INVALID, "rsgpu search: NULL argument" ); }
  RS_TRY( ensure_device() );
  return search_dev( false, g, d->query_pts, d->n_query_pts, d->radius, d->k, d->distances_sq, d->indices,
                     (unsigned long long*)d->n_neighbors, total );
}
int rsgpu_grid_knn_search_dev( const rsgpu_grid_t* g, rsgpu_search_desc_t* d, size_t* total )
{
  if( !g || !d || !d->query_pts || !d->distances_sq || !d->indices ) { return fail( RSGPU_ERR_INVALID, "rsgpu search: NULL argument" ); }
  RS_TRY( ensure_device() );
  return search_dev( true, g, d->query_pts, d->n_query_pts, d->radius, d->k, d->distances_sq, d->indices,
                     (unsigned long long*)d->n_neighbors, total );
}

int rsgpu_grid_search_census_dev( const rsgpu_grid_t* g, const float* d_query_pts, size_t n_query_pts, float radius, int64_t counts[2] )
{
  if( !g || !counts || ( n_query_pts > 0 && !d_query_pts ) || !( radius > 0.f ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_grid_search_census: bad argument" ); }
  RS_TRY( ensure_device() );
  counts[0] = counts[1] = 0;
  if( n_query_pts == 0 || g->info.n_pts == 0 ) { return RSGPU_OK; }
  DevBuf<unsigned long long> dc;
  RS_CUDA( dc.alloc( 2 ) );
  RS_CUDA( cudaMemsetAsync( dc.p, 0, 16, rt().stream ) );
  size_t blocks = ( n_query_pts + 255 ) / 256; if( blocks > 148 * 16 ) { blocks = 148 * 16; }
  search_census_kernel<<<(unsigned)blocks, 256, 0, rt().stream>>>( g->view(), d_query_pts, n_query_pts, (double)radius, dc.p );
  RS_CHECK_LAUNCH();
  unsigned long long h[2];
  RS_CUDA( cudaMemcpyAsync( h, dc.p, 16, cudaMemcpyDeviceToHost, rt().stream ) );
  RS_CUDA( rs::stream_sync( rt().stream ) );
  counts[0] = (int64_t)h[0]; counts[1] = (int64_t)h[1];
  return RSGPU_OK;
}

int rsgpu_neighborhood( const rsgpu_grid_t* grid, const float* pos, const float* nor, int32_t n, int32_t max_nn, float radius_sq,
                        float dist_exp, float angle_exp, int32_t* neighbors, float* weights )
{
  if( !grid || n < 0 || max_nn <= 0 || ( n > 0 && ( !pos || !nor || !neighbors || !weights ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_neighborhood: bad argument" ); }
  if( max_nn > 32 ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_neighborhood: max_nn > 32" ); }
  RS_TRY( ensure_device() );
  if( n == 0 ) { return RSGPU_OK; }
  cudaStream_t st = rt().stream;
  DevBuf<float> dp, dn, dw; DevBuf<int32_t> di;
  RS_CUDA( dp.alloc( (size_t)n * 3 ) ); RS_CUDA( dn.alloc( (size_t)n * 3 ) ); RS_CUDA( dw.alloc( (size_t)n * max_nn ) ); RS_CUDA( di.alloc( (size_t)n * max_nn ) );
  RS_CUDA( cudaMemcpyAsync( dp.p, pos, sizeof( float ) * 3 * (size_t)n, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( dn.p, nor, sizeof( float ) * 3 * (size_t)n, cudaMemcpyHostToDevice, st ) );
  float radius = (float)sqrt( radius_sq ); // search_opts.radius = sqrt(radius_sq) narrowed to float (:690)
  double r = radius;
  float r2f = (float)( r * r );
  {
    ProfScope prof( "edges" );
    size_t blocks = ( (size_t)n + 3 ) / 4; if( blocks > 148 * 64 ) { blocks = 148 * 64; }
    neighborhood_kernel<<<(unsigned)blocks, 128, 0, st>>>( grid->view(), dp.p, dn.p, n, r, r2f, max_nn, radius_sq, dist_exp, angle_exp, di.p, dw.p );
    RS_CHECK_LAUNCH();
  }
  RS_CUDA( cudaMemcpyAsync( neighbors, di.p, sizeof( int32_t ) * (size_t)n * max_nn, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( cudaMemcpyAsync( weights, dw.p, sizeof( float ) * (size_t)n * max_nn, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  return RSGPU_OK;
}

} // extern "C"
