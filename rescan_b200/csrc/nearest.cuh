// Nearest normal-compatible neighbour under the reference's k-nearest cap — the query both
// mgs_compute_object_alignment_score (reference apps/pose_proposal/pose_proposal.cpp:124-148) and icp_find_corrs
// (lib/rs/icp.h:349-380) put to the grid — without materialising the k-list:
//
//   accepted point = the nearest point within `radius` whose normal is compatible (dot in [dot_thr, 1]),
//   provided fewer than k points are strictly closer (= it lies inside the reference's sorted k-nearest list).
//
// Work split of one warp over a batch of 32 queries:
//   stage 1 (lane-parallel, one query per lane)  cell window with the reference's fp64 cell arithmetic, the
//            per-axis cell gaps, and a one-load occupancy test against the grid's 3x3x3 block-count table:
//            queries whose window holds no point at all — most candidate poses put most object points in free
//            space — stop here;
//   stage 2 (warp-cooperative, one query at a time)  lane l owns cell l of the (<= 27 cell) window: two
//            cell_start loads resolve its point range; cells are swept nearest-first, 32 consecutive 16-byte
//            records per step (one coalesced 512-byte request), pruned by the running best distance;
//            cells whose normal cone cannot contain a compatible normal are not swept at all (ConeCull below);
//   the rank test (fewer than k points strictly closer) is a second, counting-only sweep over the cells nearer
//            than the winner that stops at k; windows holding fewer than k points skip it (the cap cannot bind).
// All pruning is conservative, so results equal the brute-force definition above (exact distance ties excepted).
#pragma once
#include "rsgpu_internal.cuh"

#ifdef __CUDACC__
struct NearestHit
{
  float d2, dot;
  uint32_t pos;
  bool found;
};

// per-lane query state produced by stage 1
struct LaneQuery
{
  float px, py, pz, nx, ny, nz;  // query point and normal (filled by the caller)
  CellWindow w;
  float glx, ghx, gly, ghy, glz, ghz; // fast path: gap to the lower / upper neighbour cell per axis
  bool active, fast;
};

// Normal-cone culling.  Each non-empty cell stores the unit mean u of its normals and cos(alpha), alpha = the
// largest angle between u and a normal of the cell.  A point of the cell can only be compatible with the query
// normal n (dot >= dot_thr = cos(beta)) if angle(n, u) <= alpha + beta, so a cell with
//   dot(n, u) < cos(alpha + beta')      (beta' = beta widened by a safety margin)
// is skipped when looking for the nearest compatible point.  Only used when all normals involved are unit
// length to 1e-4 (checked at rsgpu_grid_set_normals / per query); planar regions (floors, walls) — where the
// reference wastes its whole k-list on incompatible points — have cones of a few degrees.
struct ConeCull
{
  bool on;
  float cb, sb; // cos / sin of beta'
};
__device__ __forceinline__ ConeCull make_cull( const GridView& g, float dot_thr, float nx, float ny, float nz )
{
  ConeCull c; c.on = false; c.cb = 0.f; c.sb = 1.f;
  float n2 = nx * nx + ny * ny + nz * nz;
  if( g.cone && dot_thr >= 1e-3f && dot_thr <= 1.0f && fabsf( n2 - 1.0f ) < 2e-4f )
  {
    c.on = true;
    c.cb = dot_thr - 1e-3f;
    c.sb = sqrtf( fmaxf( 0.0f, 1.0f - c.cb * c.cb ) ) ;
  }
  return c;
}
// the same test on a cone already loaded: u = {ux, uy, uz, cos(alpha)} (cos(alpha) <= 0: no usable cone)
__device__ __forceinline__ bool cone_possible_loaded( const float4& u, const ConeCull& c, float nx, float ny, float nz )
{
  if( !c.on ) { return true; }
  if( !( u.w > 0.0f ) ) { return true; }
  const float s2 = fmaxf( 1e-12f, 1.0f - u.w * u.w );
  float sa = s2 * rsqrtf( s2 ) + 2e-5f;
  float ct = u.x * nx + u.y * ny + u.z * nz;
  return !( ct < u.w * c.cb - sa * c.sb );
}
__device__ __forceinline__ bool cone_possible( const float4* __restrict__ cones, const ConeCull& c, size_t cell_id, float nx, float ny, float nz )
{
  if( !c.on ) { return true; }
  float4 u = __ldg( cones + cell_id ); // {ux, uy, uz, cos(alpha)} (cos(alpha) <= 0: no usable cone)
  if( !( u.w > 0.0f ) ) { return true; }
  const float s2 = fmaxf( 1e-12f, 1.0f - u.w * u.w );
  float sa = s2 * rsqrtf( s2 ) + 2e-5f; // sin(alpha) from the hardware reciprocal square root, rounded up by the margin
  float ct = u.x * nx + u.y * ny + u.z * nz;
  return !( ct < u.w * c.cb - sa * c.sb );
}

// stage 1: window + occupancy filter for this lane's own query
__device__ __forceinline__ void lane_query_setup( const GridView& g, double radius, float dot_thr, bool allow_cull, LaneQuery& q, bool valid )
{
  q.active = false; q.fast = false;
  q.glx = q.ghx = q.gly = q.ghy = q.glz = q.ghz = 0.f;
  if( !valid ) { q.w.n_cells = 0; return; }
  q.w = make_window( g, q.px, q.py, q.pz, radius );
  if( q.w.n_cells == 0 ) { return; }
  const bool inside = q.w.c0x >= 0 && q.w.c0x < g.W && q.w.c0y >= 0 && q.w.c0y < g.H && q.w.c0z >= 0 && q.w.c0z < g.D;
  const bool within1 = q.w.lox >= q.w.c0x - 1 && q.w.lox + q.w.nx <= q.w.c0x + 2 && q.w.loy >= q.w.c0y - 1 &&
                       q.w.loy + q.w.ny <= q.w.c0y + 2 && q.w.loz >= q.w.c0z - 1 && q.w.loz + q.w.nz <= q.w.c0z + 2;
  if( inside && within1 && g.occ27 )
  {
    // the window is a subset of the 3x3x3 block around the query's own cell: one load decides emptiness
    const size_t c0id = ( (size_t)q.w.c0z * g.H + q.w.c0y ) * g.W + q.w.c0x;
    q.active = __ldg( g.occ27 + c0id ) != 0;
    q.fast = true;
    if( q.active && allow_cull && g.ncone )
    {
      // one more load: the cone of ALL normals in that block; most surface-adjacent queries of a wrong pose face a
      // single plane whose normals cannot be compatible, and end here as well
      ConeCull cull = make_cull( g, dot_thr, q.nx, q.ny, q.nz );
      q.active = cone_possible( g.ncone, cull, c0id, q.nx, q.ny, q.nz );
    }
    if( q.active )
    {
      // gaps of msh_hash_grid.h:1196-1198 for the cells below (c0 - 1) and above (c0 + 1) the query's own
      q.glx = (float)__dsub_rn( (double)q.w.qx, __dmul_rn( (double)q.w.c0x, g.cell ) );
      q.ghx = (float)__dsub_rn( __dmul_rn( (double)( q.w.c0x + 1 ), g.cell ), (double)q.w.qx );
      q.gly = (float)__dsub_rn( (double)q.w.qy, __dmul_rn( (double)q.w.c0y, g.cell ) );
      q.ghy = (float)__dsub_rn( __dmul_rn( (double)( q.w.c0y + 1 ), g.cell ), (double)q.w.qy );
      q.glz = (float)__dsub_rn( (double)q.w.qz, __dmul_rn( (double)q.w.c0z, g.cell ) );
      q.ghz = (float)__dsub_rn( __dmul_rn( (double)( q.w.c0z + 1 ), g.cell ), (double)q.w.qz );
    }
  }
  else { q.active = true; }
}

__device__ __forceinline__ CellWindow shfl_window( const CellWindow& w, int src )
{
  CellWindow o;
  o.qx = __shfl_sync( RS_FULL, w.qx, src ); o.qy = __shfl_sync( RS_FULL, w.qy, src ); o.qz = __shfl_sync( RS_FULL, w.qz, src );
  o.c0x = __shfl_sync( RS_FULL, w.c0x, src ); o.c0y = __shfl_sync( RS_FULL, w.c0y, src ); o.c0z = __shfl_sync( RS_FULL, w.c0z, src );
  o.lox = __shfl_sync( RS_FULL, w.lox, src ); o.loy = __shfl_sync( RS_FULL, w.loy, src ); o.loz = __shfl_sync( RS_FULL, w.loz, src );
  int packed = __shfl_sync( RS_FULL, w.nx | ( w.ny << 10 ) | ( w.nz << 20 ), src ); // extents <= 512 each
  o.nx = packed & 1023; o.ny = ( packed >> 10 ) & 1023; o.nz = ( packed >> 20 ) & 1023;
  o.n_cells = __shfl_sync( RS_FULL, w.n_cells, src );
  return o;
}

// running state of one cooperative query
struct SweepState
{
  uint32_t best, best_pos; float best_dot; // lane-local nearest compatible so far (d2 as ordered bits)
  uint32_t dc;                              // warp-uniform bound = min over lanes of best
  unsigned nHits;                           // COUNT only
};

// sweep one cell's records [cs, ce): 32 consecutive 16-byte records per step
template <bool COUNT>
__device__ __forceinline__ void sweep_cell_nc( const GridView& g, uint32_t cs, uint32_t ce, int lane, float px, float py, float pz,
                                               float nx, float ny, float nz, float dot_thr, uint32_t r2bits, SweepState& st )
{
  uint32_t lim = st.best < st.dc ? st.best : st.dc;
  for( uint32_t p = cs + lane; p < ce; p += 32 )
  {
    float4 rec = __ldg( g.recs + p );
    uint32_t db = __float_as_uint( dist2_exact( rec, px, py, pz ) );
    if( COUNT ) { st.nHits += db < r2bits; }
    if( db < lim )
    {
      float4 m = __ldg( g.nrm + p );
      float dot = dot3_exact( m.x, m.y, m.z, nx, ny, nz );
      if( dot >= dot_thr && dot <= 1.0f ) { st.best = db; lim = db; st.best_pos = p; st.best_dot = dot; }
    }
  }
}

// phase 1 over one chunk of <= 32 cells (lane l holds cell l: range [s, t), squared gap gap2; `possible` false
// when the cell's normal cone proves that none of its points can be compatible)
template <bool COUNT>
__device__ __forceinline__ void phase1_chunk( const GridView& g, uint32_t s, uint32_t t, float gap2, bool possible, int lane, float px,
                                              float py, float pz, float nx, float ny, float nz, float dot_thr, uint32_t r2bits,
                                              SweepState& st )
{
  // key = gap bits with the lane in the low 5 bits: one redux gives the nearest unvisited cell and its owner
  // (dropping 5 mantissa bits only makes the pruning test marginally more conservative)
  uint32_t key = ( s < t && possible ) ? ( ( __float_as_uint( gap2 ) & 0xffffffe0u ) | (uint32_t)lane ) : 0xffffffffu;
  while( true )
  {
    uint32_t kmin = __reduce_min_sync( RS_FULL, key );
    if( kmin == 0xffffffffu ) { break; }
    if( !COUNT && ( kmin & 0xffffffe0u ) >= st.dc ) { break; }
    int src = (int)( kmin & 31u );
    uint32_t cs = __shfl_sync( RS_FULL, s, src ), ce = __shfl_sync( RS_FULL, t, src );
    if( lane == src ) { key = 0xffffffffu; }
    sweep_cell_nc<COUNT>( g, cs, ce, lane, px, py, pz, nx, ny, nz, dot_thr, r2bits, st );
    st.dc = __reduce_min_sync( RS_FULL, st.best );
  }
}

// phase 2 over one chunk: count points strictly closer than dcf, stop at k
__device__ __forceinline__ void phase2_chunk( const GridView& g, uint32_t s, uint32_t t, float gap2, int lane, float px, float py, float pz,
                                              float dcf, uint32_t uk, uint32_t& cnt )
{
  unsigned todo = __ballot_sync( RS_FULL, s < t && gap2 < dcf );
  while( todo && cnt < uk )
  {
    int src = __ffs( todo ) - 1; todo &= todo - 1;
    uint32_t cs = __shfl_sync( RS_FULL, s, src ), ce = __shfl_sync( RS_FULL, t, src );
    for( uint32_t p0 = cs; p0 < ce && cnt < uk; p0 += 32 )
    {
      uint32_t p = p0 + lane;
      bool closer = false;
      if( p < ce ) { float4 rec = __ldg( g.recs + p ); closer = dist2_exact( rec, px, py, pz ) < dcf; }
      cnt += __popc( __ballot_sync( RS_FULL, closer ) );
    }
  }
}

// stage 2: all 32 lanes call this with the same (broadcast) query.  FAST: the window is a subset of the
// 3x3x3 block around the query's own cell (always the case when radius <= cell size) and lane l < 27 owns block
// cell (l % 3, l / 3 % 3, l / 9) relative to (lox, loy, loz); otherwise the generic enumeration in chunks of 32.
template <bool COUNT>
__device__ __forceinline__ NearestHit nearest_compatible_w( const GridView& g, const CellWindow& w, bool fast, float glx, float ghx,
                                                            float gly, float ghy, float glz, float ghz, float px, float py, float pz,
                                                            float nx, float ny, float nz, float r2f, float dot_thr, int k,
                                                            unsigned long long* counts )
{
  const int lane = threadIdx.x & 31;
  NearestHit hit; hit.found = false; hit.d2 = 0.f; hit.dot = 0.f; hit.pos = 0;
  if( w.n_cells == 0 ) { return hit; }
  const uint32_t r2bits = __float_as_uint( r2f );
  const uint32_t uk = (uint32_t)k;
  SweepState st; st.best = r2bits; st.best_pos = 0xffffffffu; st.best_dot = 0.f; st.dc = r2bits; st.nHits = 0;
  unsigned long long nB = 0, nC = 0;
  uint32_t s0 = 0, t0 = 0; float gap0 = __int_as_float( RS_INF_BITS );
  bool possible0 = true;
  ConeCull cull = make_cull( g, dot_thr, nx, ny, nz );
  if( COUNT ) { cull.on = false; }
  if( fast )
  {
    const int ix = lane % 3, iy = ( lane / 3 ) % 3, iz = lane / 9;
    if( lane < 27 && ix < w.nx && iy < w.ny && iz < w.nz )
    {
      const int cx = w.lox + ix, cy = w.loy + iy, cz = w.loz + iz;
      const size_t id = ( (size_t)cz * g.H + cy ) * g.W + cx;
      s0 = __ldg( g.cell_start + id ); t0 = __ldg( g.cell_start + id + 1 );
      const float gx = cx < w.c0x ? glx : ( cx > w.c0x ? ghx : 0.0f );
      const float gy = cy < w.c0y ? gly : ( cy > w.c0y ? ghy : 0.0f );
      const float gz = cz < w.c0z ? glz : ( cz > w.c0z ? ghz : 0.0f );
      gap0 = __fadd_rn( __fadd_rn( __fmul_rn( gz, gz ), __fmul_rn( gy, gy ) ), __fmul_rn( gx, gx ) );
      if( s0 < t0 && gap0 < r2f ) { possible0 = cone_possible( g.cone, cull, id, nx, ny, nz ); }
    }
  }
  else { window_cell( g, w, lane, s0, t0, gap0 ); }
  // the k-cap can only bind when the window holds at least k points
  const uint32_t npts = __reduce_add_sync( RS_FULL, t0 - s0 );
  const bool cap = w.n_cells > 32 || npts >= uk;
  if( COUNT ) { nB += __popc( __ballot_sync( RS_FULL, s0 < t0 ) ); nC += npts; }
  // ---- phase 1: nearest compatible point, cells visited nearest-first and pruned by the running best
  phase1_chunk<COUNT>( g, s0, t0, gap0, possible0, lane, px, py, pz, nx, ny, nz, dot_thr, r2bits, st );
  for( int base = 32; base < w.n_cells; base += 32 )
  {
    uint32_t s, t; float gap2;
    window_cell( g, w, base + lane, s, t, gap2 );
    if( COUNT ) { nB += __popc( __ballot_sync( RS_FULL, s < t ) ); nC += __reduce_add_sync( RS_FULL, t - s ); }
    phase1_chunk<COUNT>( g, s, t, gap2, true, lane, px, py, pz, nx, ny, nz, dot_thr, r2bits, st );
  }
  uint32_t cnt = 0;
  if( st.dc < r2bits )
  {
    // winner: smallest recs position among the lanes holding dc (deterministic tie-break)
    uint32_t wpos = __reduce_min_sync( RS_FULL, st.best == st.dc ? st.best_pos : 0xffffffffu );
    int src = __ffs( __ballot_sync( RS_FULL, st.best == st.dc && st.best_pos == wpos ) ) - 1;
    hit.d2 = __uint_as_float( st.dc ); hit.pos = wpos; hit.dot = __shfl_sync( RS_FULL, st.best_dot, src );
    const float dcf = hit.d2;
    if( !cap && !COUNT ) { hit.found = true; } // fewer than k points in the whole window: the rank cannot reach k
    else
    {
      // ---- phase 2: rank of the winner = number of points strictly closer; k or more => it is not in the k-list
      phase2_chunk( g, s0, t0, gap0, lane, px, py, pz, dcf, uk, cnt );
      for( int base = 32; base < w.n_cells && cnt < uk; base += 32 )
      {
        uint32_t s, t; float gap2;
        window_cell( g, w, base + lane, s, t, gap2 );
        phase2_chunk( g, s, t, gap2, lane, px, py, pz, dcf, uk, cnt );
      }
      hit.found = cnt < uk;
    }
  }
  if( COUNT && counts )
  {
    // normals the reference fetches: up to and including the accepted neighbour, else the whole k-list
    unsigned long long nHits = __reduce_add_sync( RS_FULL, st.nHits );
    unsigned long long kk = (unsigned long long)k;
    unsigned long long T = hit.found ? (unsigned long long)cnt + 1 : ( nHits < kk ? nHits : kk );
    if( lane == 0 ) { atomicAdd( counts + 0, 1ull ); atomicAdd( counts + 1, nB ); atomicAdd( counts + 2, nC ); atomicAdd( counts + 3, T ); }
  }
  return hit;
}

// One batch of up to 32 queries, one per lane (q.px .. q.nz filled by the caller, `valid` false for padding lanes).
// On return each lane holds the result of ITS query.
template <bool COUNT>
__device__ __forceinline__ NearestHit nearest_compatible_batch( const GridView& g, LaneQuery& q, bool valid, double radius, float r2f,
                                                                float dot_thr, int k, unsigned long long* counts )
{
  const int lane = threadIdx.x & 31;
  lane_query_setup( g, radius, dot_thr, !COUNT, q, valid );
  NearestHit mine; mine.found = false; mine.d2 = 0.f; mine.dot = 0.f; mine.pos = 0;
  if( COUNT )
  {
    unsigned nq = __popc( __ballot_sync( RS_FULL, valid && !q.active ) ); // queries that end in stage 1 still count
    if( lane == 0 && nq && counts ) { atomicAdd( counts + 0, (unsigned long long)nq ); }
  }
  unsigned todo = __ballot_sync( RS_FULL, q.active );
  const unsigned fastmask = __ballot_sync( RS_FULL, q.fast );
  while( todo )
  {
    int src = __ffs( todo ) - 1; todo &= todo - 1;
    const bool fast = ( fastmask >> src ) & 1u;
    float px = __shfl_sync( RS_FULL, q.px, src ), py = __shfl_sync( RS_FULL, q.py, src ), pz = __shfl_sync( RS_FULL, q.pz, src );
    float nx = __shfl_sync( RS_FULL, q.nx, src ), ny = __shfl_sync( RS_FULL, q.ny, src ), nz = __shfl_sync( RS_FULL, q.nz, src );
    NearestHit h;
    if( fast )
    {
      CellWindow w;
      w.lox = __shfl_sync( RS_FULL, q.w.lox, src ); w.loy = __shfl_sync( RS_FULL, q.w.loy, src ); w.loz = __shfl_sync( RS_FULL, q.w.loz, src );
      // extents (1..3) and the own cell's offset inside the window (0..1) travel in one word
      int packed = __shfl_sync( RS_FULL, q.w.nx | ( q.w.ny << 2 ) | ( q.w.nz << 4 ) | ( ( q.w.c0x - q.w.lox ) << 6 ) |
                                         ( ( q.w.c0y - q.w.loy ) << 8 ) | ( ( q.w.c0z - q.w.loz ) << 10 ), src );
      w.nx = packed & 3; w.ny = ( packed >> 2 ) & 3; w.nz = ( packed >> 4 ) & 3;
      w.c0x = w.lox + ( ( packed >> 6 ) & 3 ); w.c0y = w.loy + ( ( packed >> 8 ) & 3 ); w.c0z = w.loz + ( ( packed >> 10 ) & 3 );
      w.n_cells = w.nx * w.ny * w.nz; w.qx = w.qy = w.qz = 0.f;
      float glx = __shfl_sync( RS_FULL, q.glx, src ), ghx = __shfl_sync( RS_FULL, q.ghx, src );
      float gly = __shfl_sync( RS_FULL, q.gly, src ), ghy = __shfl_sync( RS_FULL, q.ghy, src );
      float glz = __shfl_sync( RS_FULL, q.glz, src ), ghz = __shfl_sync( RS_FULL, q.ghz, src );
      h = nearest_compatible_w<COUNT>( g, w, true, glx, ghx, gly, ghy, glz, ghz, px, py, pz, nx, ny, nz, r2f, dot_thr, k, counts );
    }
    else
    {
      CellWindow w = shfl_window( q.w, src );
      h = nearest_compatible_w<COUNT>( g, w, false, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, px, py, pz, nx, ny, nz, r2f, dot_thr, k, counts );
    }
    if( lane == src ) { mine = h; }
  }
  return mine;
}
#endif // __CUDACC__
