// Sub-warp ("group") version of the nearest normal-compatible neighbour query of nearest.cuh: G lanes per query,
// 32 / G queries per warp at a time.  Same definition of the result, same exact arithmetic:
//
//   accepted point = the nearest point within `radius` whose normal is compatible (dot in [dot_thr, 1]),
//   provided fewer than k points are strictly closer (= it lies inside the reference's sorted k-nearest list;
//   reference apps/pose_proposal/pose_proposal.cpp:124-148, lib/rs/icp.h:349-380, lib/msh/msh_hash_grid.h:1090-1259).
//
// Why groups: level-1 cells hold a few dozen points, so a 32-lane sweep of one cell is mostly fixed cost (cell
// election, shuffles, reductions) paid once per QUERY; with G = 4 the same instructions serve 8 queries, a
// group still reads 64 contiguous bytes per step, and the dependent load chain of 8 queries overlaps.
//
//   phase 1 (warp-uniform)   the <= 27 cells of the query's window are evaluated G at a time by the group's lanes
//                            (own cell first, then face / edge / corner neighbours): clipped? gap < r? non-empty?
//                            survivors are parked in a per-warp shared table {start, end, gap bits, cell id};
//   phase 2 (per group)      cells popped in that order, skipped when their gap is not below the running best or
//                            their normal cone (nearest.cuh) excludes a compatible point, else swept G records a step;
//   phase 3 (per group)      rank of the winner: count points strictly closer, stop at k; skipped when the 3x3x3
//                            block around the query holds fewer than k points.
//
// Preconditions (checked per query by stage1_test): the window lies inside the 3x3x3 block around the query's own
// cell — always true for radius <= cell size up to last-bit cases; other queries take the generic warp-cooperative
// path of nearest.cuh.
#pragma once
#include "nearest.cuh"

#ifdef __CUDACC__
namespace rsg
{
// e-th cell of the visiting order as offsets (0, +1, +2 mod 3) from the query's own cell, 2 bits per axis
constexpr unsigned cand_code( int e )
{
  int n = 0;
  for( int cls = 0; cls <= 3; ++cls )
    for( int dz = 0; dz < 3; ++dz )
      for( int dy = 0; dy < 3; ++dy )
        for( int dx = 0; dx < 3; ++dx )
          if( ( dx != 0 ) + ( dy != 0 ) + ( dz != 0 ) == cls )
          {
            if( n == e ) { return (unsigned)( dx | ( dy << 2 ) | ( dz << 4 ) ); }
            ++n;
          }
  return 0;
}
template <int G>
constexpr unsigned long long cand_word( int j )
{
  unsigned long long w = 0;
  for( int s = 0; s < G; ++s )
  {
    int e = G * j + s;
    w |= (unsigned long long)( e < 27 ? cand_code( e ) : 0u ) << ( 6 * s );
  }
  return w;
}
static __constant__ unsigned long long kCandW4[7] = { cand_word<4>( 0 ), cand_word<4>( 1 ), cand_word<4>( 2 ), cand_word<4>( 3 ),
                                                      cand_word<4>( 4 ), cand_word<4>( 5 ), cand_word<4>( 6 ) };
static __constant__ unsigned long long kCandW8[4] = { cand_word<8>( 0 ), cand_word<8>( 1 ), cand_word<8>( 2 ), cand_word<8>( 3 ) };

template <int G> __device__ __forceinline__ unsigned long long cand_words( int j );
template <> __device__ __forceinline__ unsigned long long cand_words<4>( int j ) { return kCandW4[j]; }
template <> __device__ __forceinline__ unsigned long long cand_words<8>( int j ) { return kCandW8[j]; }

template <int G> struct GroupCfg
{
  static constexpr int NC = ( 27 + G - 1 ) / G;      // candidate cells per lane
  static constexpr int NG = 32 / G;                  // queries per warp step
  static constexpr int CAND_WORDS = NC * 32;         // uint4 entries of the per-warp cell table
};

// stage 1, one query per lane: is there anything to search at all, and may the group path take it
struct Stage1
{
  bool active, fast;
};
__device__ __forceinline__ Stage1 stage1_test( const GridView& g, double radius, float dot_thr, bool allow_cull, float px, float py, float pz,
                                               float nx, float ny, float nz, bool valid )
{
  Stage1 r; r.active = false; r.fast = false;
  if( !valid ) { return r; }
  const CellWindow w = make_window( g, px, py, pz, radius );
  if( w.n_cells == 0 ) { return r; }
  const bool inside = w.c0x >= 0 && w.c0x < g.W && w.c0y >= 0 && w.c0y < g.H && w.c0z >= 0 && w.c0z < g.D;
  const bool within1 = w.lox >= w.c0x - 1 && w.lox + w.nx <= w.c0x + 2 && w.loy >= w.c0y - 1 && w.loy + w.ny <= w.c0y + 2 &&
                       w.loz >= w.c0z - 1 && w.loz + w.nz <= w.c0z + 2;
  r.fast = within1;
  r.active = true;
  if( inside && within1 && g.occ27 )
  {
    // the window is a subset of the 3x3x3 block around the query's own cell: one load decides emptiness, a second
    // one whether any normal of the block can be compatible (nearest.cuh ConeCull)
    const size_t c0id = ( (size_t)w.c0z * g.H + w.c0y ) * g.W + w.c0x;
    if( allow_cull && g.ncone )
    {
      // ONE load answers both questions: an empty block is stored as cos = RS_NCONE_EMPTY (grid.cu ncone_kernel), so the
      // occupancy count does not have to be fetched first
      const float4 u = __ldg( g.ncone + c0id );
      if( u.w > 1.5f ) { r.active = false; }
      else
      {
        const ConeCull cull = make_cull( g, dot_thr, nx, ny, nz );
        r.active = cone_possible_loaded( u, cull, nx, ny, nz );
      }
    }
    else { r.active = __ldg( g.occ27 + c0id ) != 0; }
  }
  return r;
}

static __device__ __noinline__ float normal_dot_call( const float4* __restrict__ nrm, uint32_t p, float nx, float ny, float nz )
{
  const float4 mm = __ldg( nrm + p );
  return dot3_exact( mm.x, mm.y, mm.z, nx, ny, nz );
}

template <int G>
__device__ __forceinline__ unsigned long long group_min64( unsigned long long v, unsigned gmask )
{
#pragma unroll
  for( int o = G / 2; o > 0; o >>= 1 )
  {
    unsigned long long u = __shfl_xor_sync( gmask, v, o );
    v = u < v ? u : v;
  }
  return v;
}
template <int G>
__device__ __forceinline__ unsigned group_sum32( unsigned v, unsigned gmask )
{
#pragma unroll
  for( int o = G / 2; o > 0; o >>= 1 ) { v += __shfl_xor_sync( gmask, v, o ); }
  return v;
}

// All 32 lanes call this; the G lanes of a group pass the same query (qv false: the group idles).  `cand` is
// this warp's table of GroupCfg<G>::CAND_WORDS uint4 in shared memory.  The result is replicated in the group.
template <int G>
__device__ __forceinline__ NearestHit group_search( const GridView& g, bool qv, float px, float py, float pz, float nx, float ny, float nz,
                                                    double radius, float r2f, float dot_thr, int k, uint4* __restrict__ cand,
                                                    unsigned long long seedkey = ~0ull, float seeddot = 0.f )
{
  // seedkey (optional): key (d2 bits << 32 | recs position) of a point already known to be compatible and inside the
  // radius — e.g. the previous ICP iteration's correspondent; it only tightens the pruning bound from the start
  constexpr int NC = GroupCfg<G>::NC;
  const int lane = threadIdx.x & 31, sub = lane & ( G - 1 ), gbase = lane & ~( G - 1 );
  const unsigned gmask = ( ( G == 32 ) ? 0xffffffffu : ( ( 1u << G ) - 1u ) ) << gbase;
  NearestHit hit; hit.found = false; hit.d2 = 0.f; hit.dot = 0.f; hit.pos = 0;
  const uint32_t r2bits = __float_as_uint( r2f );

  // ---- the window and the squared per-axis gaps to the neighbour cells (msh_hash_grid.h:1150-1198), replicated
  CellWindow w; w.n_cells = 0;
  if( qv ) { w = make_window( g, px, py, pz, radius ); qv = w.n_cells != 0; }
  float glx2 = 0.f, ghx2 = 0.f, gly2 = 0.f, ghy2 = 0.f, glz2 = 0.f, ghz2 = 0.f;
  int oxc = 0, oyc = 0, ozc = 0;
  bool cap = true;
  if( qv )
  {
    float a;
    a = (float)__dsub_rn( (double)w.qx, __dmul_rn( (double)w.c0x, g.cell ) ); glx2 = __fmul_rn( a, a );
    a = (float)__dsub_rn( __dmul_rn( (double)( w.c0x + 1 ), g.cell ), (double)w.qx ); ghx2 = __fmul_rn( a, a );
    a = (float)__dsub_rn( (double)w.qy, __dmul_rn( (double)w.c0y, g.cell ) ); gly2 = __fmul_rn( a, a );
    a = (float)__dsub_rn( __dmul_rn( (double)( w.c0y + 1 ), g.cell ), (double)w.qy ); ghy2 = __fmul_rn( a, a );
    a = (float)__dsub_rn( (double)w.qz, __dmul_rn( (double)w.c0z, g.cell ) ); glz2 = __fmul_rn( a, a );
    a = (float)__dsub_rn( __dmul_rn( (double)( w.c0z + 1 ), g.cell ), (double)w.qz ); ghz2 = __fmul_rn( a, a );
    oxc = min( max( w.c0x - w.lox, 0 ), 2 ); oyc = min( max( w.c0y - w.loy, 0 ), 2 ); ozc = min( max( w.c0z - w.loz, 0 ), 2 );
    const bool inside = w.c0x >= 0 && w.c0x < g.W && w.c0y >= 0 && w.c0y < g.H && w.c0z >= 0 && w.c0z < g.D;
    if( inside && g.occ27 )
    {
      // the k-cap can only bind when the 3x3x3 block (a superset of the window) holds at least k points
      cap = __ldg( g.occ27 + ( ( (size_t)w.c0z * g.H + w.c0y ) * g.W + w.c0x ) ) >= (uint32_t)k;
    }
  }

  // a seed bounds everything: no cell at or beyond its distance can hold a closer compatible point or a point of its rank
  const uint32_t lim1 = ( qv && (uint32_t)( seedkey >> 32 ) < r2bits ) ? (uint32_t)( seedkey >> 32 ) : r2bits;
  // ---- phase 1: evaluate the candidate cells, G at a time (loads of all candidates are in flight together)
  const ConeCull cull = make_cull( g, dot_thr, nx, ny, nz );
  unsigned cmask = 0, smask = 0; // bit e: cell e is in range and non-empty / ... and its normal cone admits a compatible point
#pragma unroll
  for( int j = 0; j < NC; ++j )
  {
    const int e = G * j + sub;
    const unsigned code = (unsigned)( cand_words<G>( j ) >> ( 6 * sub ) ) & 63u;
    int ix = oxc + (int)( code & 3u ), iy = oyc + (int)( ( code >> 2 ) & 3u ), iz = ozc + (int)( code >> 4 );
    ix -= ix >= 3 ? 3 : 0; iy -= iy >= 3 ? 3 : 0; iz -= iz >= 3 ? 3 : 0;
    if( qv && e < 27 && ix < w.nx && iy < w.ny && iz < w.nz )
    {
      const int cx = w.lox + ix, cy = w.loy + iy, cz = w.loz + iz;
      const float gx2 = cx < w.c0x ? glx2 : ( cx > w.c0x ? ghx2 : 0.0f );
      const float gy2 = cy < w.c0y ? gly2 : ( cy > w.c0y ? ghy2 : 0.0f );
      const float gz2 = cz < w.c0z ? glz2 : ( cz > w.c0z ? ghz2 : 0.0f );
      // (gz*gz + gy*gy) + gx*gx (:1221); dropping 5 mantissa bits only makes every pruning test more conservative
      const uint32_t gapc = __float_as_uint( __fadd_rn( __fadd_rn( gz2, gy2 ), gx2 ) ) & 0xffffffe0u;
      if( gapc < lim1 )
      {
        const uint32_t id = (uint32_t)( ( cz * g.H + cy ) * g.W + cx ); // n_cells <= 2^30 (grid.cu)
        const uint32_t s = __ldg( g.cell_start + id ), t = __ldg( g.cell_start + id + 1 );
        if( s < t )
        {
          cand[j * 32 + lane] = make_uint4( s, t, gapc, id );
          cmask |= 1u << e;
          if( cone_possible( g.cone, cull, (size_t)id, nx, ny, nz ) ) { smask |= 1u << e; }
        }
      }
    }
  }
#pragma unroll
  for( int o = G / 2; o > 0; o >>= 1 )
  {
    cmask |= __shfl_xor_sync( RS_FULL, cmask, o );
    smask |= __shfl_xor_sync( RS_FULL, smask, o );
  }
  __syncwarp();

  // ---- phase 2: nearest compatible point; key = (d2 bits, recs position), smaller wins
  unsigned long long limkey = (unsigned long long)r2bits << 32, mykey = ~0ull;
  float mydot = 0.f;
  if( qv && seedkey < limkey )
  {
    limkey = seedkey;
    if( sub == 0 ) { mykey = seedkey; mydot = seeddot; }
  }
  for( unsigned m = smask; m; )
  {
    const int e = __ffs( m ) - 1; m &= m - 1;
    const uint4 c = cand[( e / G ) * 32 + gbase + ( e % G )];
    if( c.z >= (uint32_t)( limkey >> 32 ) ) { continue; }
    unsigned long long lk = limkey;
    for( uint32_t p = c.x + sub; p < c.y; p += 2 * G )
    {
      // two records per step, both loads issued before either is used
      const bool two = p + G < c.y;
      const float4 rec0 = __ldg( g.recs + p );
      const float4 rec1 = __ldg( g.recs + ( two ? p + G : p ) );
      const unsigned long long key0 = ( (unsigned long long)__float_as_uint( dist2_exact( rec0, px, py, pz ) ) << 32 ) | p;
      const unsigned long long key1 = ( (unsigned long long)__float_as_uint( dist2_exact( rec1, px, py, pz ) ) << 32 ) | ( p + G );
      // the normal test is rare once a near point is known: keep it a real branch (an out-of-line call stops the
      // compiler from predicating its ~15 instructions into every step)
      if( key0 < lk )
      {
        const float dot = normal_dot_call( g.nrm, p, nx, ny, nz );
        if( dot >= dot_thr && dot <= 1.0f ) { lk = key0; mykey = key0; mydot = dot; }
      }
      if( two && key1 < lk )
      {
        const float dot = normal_dot_call( g.nrm, p + G, nx, ny, nz );
        if( dot >= dot_thr && dot <= 1.0f ) { lk = key1; mykey = key1; mydot = dot; }
      }
    }
    limkey = group_min64<G>( lk, gmask );
  }
  const uint32_t dcb = (uint32_t)( limkey >> 32 );
  if( dcb < r2bits ) // group-uniform
  {
    const unsigned own = __ballot_sync( gmask, mykey == limkey );
    hit.dot = __shfl_sync( gmask, mydot, __ffs( own ) - 1 );
    hit.d2 = __uint_as_float( dcb ); hit.pos = (uint32_t)limkey;
    hit.found = true;
    if( cap )
    {
      // ---- phase 3: rank of the winner = number of points strictly closer; k or more => it is not in the k-list
      const float dcf = hit.d2;
      const uint32_t uk = (uint32_t)k;
      uint32_t cnt = 0;
      for( unsigned m = cmask; m && cnt < uk; )
      {
        const int e = __ffs( m ) - 1; m &= m - 1;
        const uint4 c = cand[( e / G ) * 32 + gbase + ( e % G )];
        if( c.z >= dcb ) { continue; }
        unsigned local = 0;
#pragma unroll 4
        for( uint32_t p = c.x + sub; p < c.y; p += G )
        {
          const float4 rec = __ldg( g.recs + p );
          local += dist2_exact( rec, px, py, pz ) < dcf;
        }
        cnt += group_sum32<G>( local, gmask );
      }
      hit.found = cnt < uk;
    }
  }
  __syncwarp(); // the table is reused by the next step
  return hit;
}

// One round = up to 32 queries, G warp steps of 32 / G queries each.  query_of(r, px, py, pz, nx, ny, nz, seedkey, seeddot) must be
// callable by all lanes with any r in [0, n_round) and give the r-th query of the round (the lanes of one group
// ask for the same r); it returns false for a query the group path must not take (left unfound here).  On return
// lane L holds the result of query L of the round.
template <int G, class QueryOf>
__device__ __forceinline__ NearestHit group_round( const GridView& g, int n_round, QueryOf query_of, double radius, float r2f, float dot_thr,
                                                   int k, uint4* __restrict__ cand )
{
  constexpr int NG = GroupCfg<G>::NG;
  const int lane = threadIdx.x & 31;
  NearestHit mine; mine.found = false; mine.d2 = 0.f; mine.dot = 0.f; mine.pos = 0;
  for( int step = 0; step * NG < n_round; ++step )
  {
    const int r = step * NG + lane / G;
    float px, py, pz, nx, ny, nz, seeddot = 0.f;
    unsigned long long seedkey = ~0ull;
    const bool ok = query_of( r < n_round ? r : n_round - 1, px, py, pz, nx, ny, nz, seedkey, seeddot );
    const NearestHit h = group_search<G>( g, ok && r < n_round, px, py, pz, nx, ny, nz, radius, r2f, dot_thr, k, cand, seedkey, seeddot );
    // query L of the round was served in step L / NG by group L % NG
    const int src = ( lane % NG ) * G;
    const float d2 = __shfl_sync( RS_FULL, h.d2, src ), dot = __shfl_sync( RS_FULL, h.dot, src );
    const uint32_t pos = __shfl_sync( RS_FULL, h.pos, src );
    const int found = __shfl_sync( RS_FULL, (int)h.found, src );
    if( lane / NG == step ) { mine.d2 = d2; mine.dot = dot; mine.pos = pos; mine.found = found != 0; }
  }
  return mine;
}
} // namespace rsg
#endif // __CUDACC__

#ifdef __CUDACC__
namespace rsg
{
static __constant__ unsigned char kCandCode[28] = {
  (unsigned char)cand_code( 0 ),  (unsigned char)cand_code( 1 ),  (unsigned char)cand_code( 2 ),  (unsigned char)cand_code( 3 ),
  (unsigned char)cand_code( 4 ),  (unsigned char)cand_code( 5 ),  (unsigned char)cand_code( 6 ),  (unsigned char)cand_code( 7 ),
  (unsigned char)cand_code( 8 ),  (unsigned char)cand_code( 9 ),  (unsigned char)cand_code( 10 ), (unsigned char)cand_code( 11 ),
  (unsigned char)cand_code( 12 ), (unsigned char)cand_code( 13 ), (unsigned char)cand_code( 14 ), (unsigned char)cand_code( 15 ),
  (unsigned char)cand_code( 16 ), (unsigned char)cand_code( 17 ), (unsigned char)cand_code( 18 ), (unsigned char)cand_code( 19 ),
  (unsigned char)cand_code( 20 ), (unsigned char)cand_code( 21 ), (unsigned char)cand_code( 22 ), (unsigned char)cand_code( 23 ),
  (unsigned char)cand_code( 24 ), (unsigned char)cand_code( 25 ), (unsigned char)cand_code( 26 ), 0 };

// One query per THREAD (same definition and arithmetic as group_search): the thread walks the <= 27 cells of its
// window itself, own cell first.  Meant for warps whose 32 queries are spatial neighbours (object points in their
// stored order at levels 1-3, queries sorted by cell): the lanes then read the same cache lines at the same time
// and run similar trip counts, and nothing is shuffled or staged.  Same precondition as group_search.
__device__ __forceinline__ NearestHit lane_search( const GridView& g, bool qv, float px, float py, float pz, float nx, float ny, float nz,
                                                   double radius, float r2f, float dot_thr, int k )
{
  NearestHit hit; hit.found = false; hit.d2 = 0.f; hit.dot = 0.f; hit.pos = 0;
  const uint32_t r2bits = __float_as_uint( r2f );
  CellWindow w; w.n_cells = 0;
  if( qv ) { w = make_window( g, px, py, pz, radius ); qv = w.n_cells != 0; }
  if( !qv ) { return hit; }
  float a;
  a = (float)__dsub_rn( (double)w.qx, __dmul_rn( (double)w.c0x, g.cell ) ); const float glx2 = __fmul_rn( a, a );
  a = (float)__dsub_rn( __dmul_rn( (double)( w.c0x + 1 ), g.cell ), (double)w.qx ); const float ghx2 = __fmul_rn( a, a );
  a = (float)__dsub_rn( (double)w.qy, __dmul_rn( (double)w.c0y, g.cell ) ); const float gly2 = __fmul_rn( a, a );
  a = (float)__dsub_rn( __dmul_rn( (double)( w.c0y + 1 ), g.cell ), (double)w.qy ); const float ghy2 = __fmul_rn( a, a );
  a = (float)__dsub_rn( (double)w.qz, __dmul_rn( (double)w.c0z, g.cell ) ); const float glz2 = __fmul_rn( a, a );
  a = (float)__dsub_rn( __dmul_rn( (double)( w.c0z + 1 ), g.cell ), (double)w.qz ); const float ghz2 = __fmul_rn( a, a );
  const int oxc = min( max( w.c0x - w.lox, 0 ), 2 ), oyc = min( max( w.c0y - w.loy, 0 ), 2 ), ozc = min( max( w.c0z - w.loz, 0 ), 2 );
  const bool inside = w.c0x >= 0 && w.c0x < g.W && w.c0y >= 0 && w.c0y < g.H && w.c0z >= 0 && w.c0z < g.D;
  bool cap = true;
  if( inside && g.occ27 ) { cap = __ldg( g.occ27 + ( ( (size_t)w.c0z * g.H + w.c0y ) * g.W + w.c0x ) ) >= (uint32_t)k; }
  const ConeCull cull = make_cull( g, dot_thr, nx, ny, nz );

  // cell e of the visiting order: point range and conservative gap bits (false: not in the window / out of range / empty)
  auto cell_of = [&]( int e, uint32_t lim, uint32_t& s, uint32_t& t, uint32_t& id ) -> bool {
    const unsigned code = kCandCode[e];
    int ix = oxc + (int)( code & 3u ), iy = oyc + (int)( ( code >> 2 ) & 3u ), iz = ozc + (int)( code >> 4 );
    ix -= ix >= 3 ? 3 : 0; iy -= iy >= 3 ? 3 : 0; iz -= iz >= 3 ? 3 : 0;
    if( ix >= w.nx || iy >= w.ny || iz >= w.nz ) { return false; }
    const int cx = w.lox + ix, cy = w.loy + iy, cz = w.loz + iz;
    const float gx2 = cx < w.c0x ? glx2 : ( cx > w.c0x ? ghx2 : 0.0f );
    const float gy2 = cy < w.c0y ? gly2 : ( cy > w.c0y ? ghy2 : 0.0f );
    const float gz2 = cz < w.c0z ? glz2 : ( cz > w.c0z ? ghz2 : 0.0f );
    const uint32_t gapc = __float_as_uint( __fadd_rn( __fadd_rn( gz2, gy2 ), gx2 ) ) & 0xffffffe0u;
    if( gapc >= lim ) { return false; }
    id = (uint32_t)( ( cz * g.H + cy ) * g.W + cx );
    s = __ldg( g.cell_start + id ); t = __ldg( g.cell_start + id + 1 );
    return s < t;
  };

  unsigned long long limkey = (unsigned long long)r2bits << 32;
  float bestdot = 0.f;
  for( int e = 0; e < 27; ++e )
  {
    uint32_t s, t, id;
    if( !cell_of( e, (uint32_t)( limkey >> 32 ), s, t, id ) ) { continue; }
    if( !cone_possible( g.cone, cull, (size_t)id, nx, ny, nz ) ) { continue; }
    for( uint32_t p = s; p < t; ++p )
    {
      const float4 rec = __ldg( g.recs + p );
      const unsigned long long key = ( (unsigned long long)__float_as_uint( dist2_exact( rec, px, py, pz ) ) << 32 ) | p;
      if( key < limkey )
      {
        const float4 mm = __ldg( g.nrm + p );
        const float dot = dot3_exact( mm.x, mm.y, mm.z, nx, ny, nz );
        if( dot >= dot_thr && dot <= 1.0f ) { limkey = key; bestdot = dot; }
      }
    }
  }
  const uint32_t dcb = (uint32_t)( limkey >> 32 );
  if( dcb >= r2bits ) { return hit; }
  hit.d2 = __uint_as_float( dcb ); hit.pos = (uint32_t)limkey; hit.dot = bestdot; hit.found = true;
  if( cap )
  {
    const float dcf = hit.d2;
    const uint32_t uk = (uint32_t)k;
    uint32_t cnt = 0;
    for( int e = 0; e < 27 && cnt < uk; ++e )
    {
      uint32_t s, t, id;
      if( !cell_of( e, dcb, s, t, id ) ) { continue; }
      for( uint32_t p = s; p < t; ++p )
      {
        const float4 rec = __ldg( g.recs + p );
        cnt += dist2_exact( rec, px, py, pz ) < dcf;
      }
    }
    hit.found = cnt < uk;
  }
  return hit;
}
} // namespace rsg
#endif // __CUDACC__
