// Dense pose-grid scoring, second design: queries binned by the scan cell they fall into, searched against the 3x3x3 block
// of cells STAGED IN SHARED MEMORY by bulk-async copies.  Included by score.cu (shares its ScoreParams / PoseSource).
//
// Replaces the loop of mgs__initial_pose_proposals over poses x object points (reference apps/pose_proposal/
// pose_proposal.cpp:213-243 calling mgs_compute_object_alignment_score :93-158) and, inside it, the per-query cell walk of
// msh_hash_grid_radius_search (lib/msh/msh_hash_grid.h:1187-1236) with msh_hash_grid__find_neighbors_in_bin (:826-862).
//
// Why: the first design (score_kernel_g) gives every (pose, point) query its own dependent chain of L1/L2 loads
// (block cone -> cell ranges -> records -> normals) and is latency-bound at a quarter of the SM's issue rate
// (profiles/score_kernel_r01.md).  Queries that fall into the same scan cell read the same <= 27 cells, but they belong to
// unrelated poses, so nothing in a pose-major kernel can share those reads.  Here the work is re-ordered:
//
//   P  prepare (once per object): the rotated clouds R_r p_i and R_r n_i of every rotation, so that the point of pose
//      (t, r) is ONE float4 load plus the translation - bit-identical to msh_mat4_vec3_mul by construction (the last
//      addition of ((m0 x + m4 y) + m8 z) + 1 m12 is the only one that involves the translation);
//   A  prefilter, one warp per pose: cell window + one load of the block normal cone per point (rsg::stage1_test); the
//      pose's survivors are counted, poses that cannot reach the level threshold are dropped (bound pruning, as before),
//      the others append their survivors to a queue and count them into their HOME CELL's bin (one atomic each);
//   S  exclusive scan of the bins, work items (cell, <= 512 queries) for the non-empty ones, scatter of the queue into
//      cell order (no atomics: position in the bin was returned by the count);
//   B  search, persistent blocks of 4 compute warps + 1 producer warp: the producer resolves an item's 27 cell ranges and
//      copies their records and normals (x-adjacent cells are contiguous: 9 + 9 bulk copies, cp.async.bulk + mbarrier
//      complete_tx) into one of two shared-memory buffers while the compute warps sweep the other; a compute warp takes
//      32 queries of the cell, ONE PER LANE, and walks the staged cells in a warp-uniform order - every record is one
//      broadcast shared-memory load serving 32 queries, per-lane pruning is predication;
//   C  per pose, the fp64 terms are summed in the reference's point order (:149-157).
//
// The definition of a query's result is unchanged (nearest_group.cuh): the nearest point within the radius whose normal
// is compatible, ties by record position, accepted iff fewer than k points are strictly closer; so scores are
// bit-identical to the first design's (tests/test_gpu_variants.py holds the two against each other).
#pragma once

namespace
{
#ifndef RS_DB_WARPS
#define RS_DB_WARPS 8
#endif
constexpr int DB_WARPS = RS_DB_WARPS;             // compute warps of the search kernel
constexpr int DB_THREADS = 32 * ( DB_WARPS + 1 ); // + the producer warp
#ifndef RS_DB_CAP
#define RS_DB_CAP 1024
#endif
#ifndef RS_DB_BPS
#define RS_DB_BPS 3
#endif
constexpr int DB_CAP = RS_DB_CAP;                 // points staged per buffer (records + normals: 32 B each)
#ifndef RS_DB_QCHUNK
#define RS_DB_QCHUNK 512
#endif
constexpr int DB_QCHUNK = RS_DB_QCHUNK;           // queries per work item
constexpr int DB_PTS_BITS = 12;                   // object points per cloud on this path: <= 4096
constexpr int DB_MAX_PTS = 1 << DB_PTS_BITS;
constexpr uint32_t DB_POSE_MAX = 1u << ( 32 - DB_PTS_BITS ); // poses per chunk
constexpr uint32_t DB_DONE = 0xffffffffu;
constexpr int DB_NCLS = 64;                       // sub-bins per ACTIVE cell: octant of the query inside the cell x class of its normal

// ------------------------------------------------------------------------------------------------ P: rotated clouds
__global__ void db_prepare_kernel( const float* __restrict__ pos, const float* __restrict__ nor, int n, const float* __restrict__ rots, int n_rot,
                                   float4* __restrict__ upos, float4* __restrict__ unor )
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if( j >= n * n_rot ) { return; }
  const int r = j / n, i = j - r * n;
  const float* m = rots + 16 * (size_t)r;
  const float x = pos[3 * (size_t)i], y = pos[3 * (size_t)i + 1], z = pos[3 * (size_t)i + 2];
  float4 u;
  // ((m0*x + m4*y) + m8*z): xf_apply up to, not including, the translation term
  u.x = __fadd_rn( __fadd_rn( __fmul_rn( m[0], x ), __fmul_rn( m[4], y ) ), __fmul_rn( m[8], z ) );
  u.y = __fadd_rn( __fadd_rn( __fmul_rn( m[1], x ), __fmul_rn( m[5], y ) ), __fmul_rn( m[9], z ) );
  u.z = __fadd_rn( __fadd_rn( __fmul_rn( m[2], x ), __fmul_rn( m[6], y ) ), __fmul_rn( m[10], z ) );
  u.w = 0.f;
  upos[j] = u;
  const float a = nor[3 * (size_t)i], b = nor[3 * (size_t)i + 1], c = nor[3 * (size_t)i + 2];
  float4 v;
  v.x = __fadd_rn( __fadd_rn( __fmul_rn( m[0], a ), __fmul_rn( m[4], b ) ), __fmul_rn( m[8], c ) );
  v.y = __fadd_rn( __fadd_rn( __fmul_rn( m[1], a ), __fmul_rn( m[5], b ) ), __fmul_rn( m[9], c ) );
  v.z = __fadd_rn( __fadd_rn( __fmul_rn( m[2], a ), __fmul_rn( m[6], b ) ), __fmul_rn( m[10], c ) );
  v.w = 0.f;
  unor[j] = v;
}

// the query of (pose = (t, r), point i): position = rotated point + 1.0f * translation, normal = rotated normal + 0.0f *
// translation - the last addition of msh_mat4_vec3_mul (msh_vec_math.h:1554-1561), so the bits are xf_apply's
struct DbPoseGrid
{
  const float4* __restrict__ upos; // n_rot x n
  const float4* __restrict__ unor;
  const float* __restrict__ trans; // n_trans x 3 (this launch's translations)
  int n_rot, n;
};
__device__ __forceinline__ void db_query( const DbPoseGrid& pg, long long t, int r, int i, float& px, float& py, float& pz, float& nx, float& ny, float& nz )
{
  const float tx = __ldg( pg.trans + 3 * t ), ty = __ldg( pg.trans + 3 * t + 1 ), tz = __ldg( pg.trans + 3 * t + 2 );
  const float4 u = __ldg( pg.upos + (size_t)r * pg.n + i ), v = __ldg( pg.unor + (size_t)r * pg.n + i );
  px = __fadd_rn( u.x, __fmul_rn( 1.0f, tx ) ); py = __fadd_rn( u.y, __fmul_rn( 1.0f, ty ) ); pz = __fadd_rn( u.z, __fmul_rn( 1.0f, tz ) );
  nx = __fadd_rn( v.x, __fmul_rn( 0.0f, tx ) ); ny = __fadd_rn( v.y, __fmul_rn( 0.0f, ty ) ); nz = __fadd_rn( v.z, __fmul_rn( 0.0f, tz ) );
}

// device-side counters of one chunk
struct DbCounters
{
  unsigned int n_queue;    // entries appended to the queue
  unsigned int n_items;    // work items of the search kernel
  unsigned int item_cursor;
  unsigned int pad;
};

// ------------------------------------------------------------------------------------------------ A: prefilter
// One warp per pose of the chunk.  Dynamic shared memory: per warp n_pad x (uint32 cell + uint16 point).
// bins: n_cells x DB_NCLS + 1 counters (cell-major, so a cell's queries are contiguous after the scan, ordered by class);
// the last bin takes the queries the staged search does not handle (a window that is not a subset of the 3x3x3 block
// around the clamped home cell: last-bit cases) - they go to the generic search.
template <int WARPS>
__global__ void __launch_bounds__( 32 * WARPS ) db_prefilter_kernel( GridView g, DbPoseGrid pg, long long pose0, unsigned n_chunk_poses, ScoreParams sp,
                                                                      double prune_cnt, int n_pad, int sub_mode, DbCounters* __restrict__ ctr,
                                                                      uint32_t* __restrict__ bins, uint32_t* __restrict__ pose_base,
                                                                      uint32_t* __restrict__ pose_cnt, uint2* __restrict__ queue,
                                                                      uint2* __restrict__ qbin )
{
  extern __shared__ __align__( 16 ) unsigned char db_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const unsigned pl = blockIdx.x * WARPS + wib; // pose within the chunk
  if( pl >= n_chunk_poses ) { return; }
  uint32_t* lcell = (uint32_t*)db_smem + (size_t)wib * n_pad;
  uint16_t* lidx = (uint16_t*)( db_smem + (size_t)WARPS * n_pad * 4 ) + (size_t)wib * n_pad;
  const long long pose = pose0 + pl;
  const long long t = pose / pg.n_rot; const int r = (int)( pose - t * pg.n_rot );
  const float tx = __ldg( pg.trans + 3 * t ), ty = __ldg( pg.trans + 3 * t + 1 ), tz = __ldg( pg.trans + 3 * t + 2 );
  const float4* __restrict__ up = pg.upos + (size_t)r * pg.n;
  const float4* __restrict__ un = pg.unor + (size_t)r * pg.n;
  const uint32_t n_cells = g.n_active; // bins are indexed by the rank of the home cell among the active cells
  int n_list = 0;
  for( int ib = 0; ib < pg.n; ib += 32 )
  {
    const int i = ib + lane;
    const bool valid = i < pg.n;
    float px = 0.f, py = 0.f, pz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    if( valid )
    {
      const float4 u = __ldg( up + i ), v = __ldg( un + i );
      px = __fadd_rn( u.x, __fmul_rn( 1.0f, tx ) ); py = __fadd_rn( u.y, __fmul_rn( 1.0f, ty ) ); pz = __fadd_rn( u.z, __fmul_rn( 1.0f, tz ) );
      nx = __fadd_rn( v.x, __fmul_rn( 0.0f, tx ) ); ny = __fadd_rn( v.y, __fmul_rn( 0.0f, ty ) ); nz = __fadd_rn( v.z, __fmul_rn( 0.0f, tz ) );
    }
    // the first design's test (window, block occupancy + block normal cone in one load), on the 3x3x3 block around the
    // query's own cell CLAMPED into the grid: a query up to one cell outside the scan's box still has a window of in-grid
    // cells, and that window is a subset of the clamped cell's block
    bool active = false; uint32_t bin = n_cells * DB_NCLS;
    if( valid )
    {
      const CellWindow w = make_window( g, px, py, pz, sp.radius );
      if( w.n_cells != 0 )
      {
        const int ccx = min( max( w.c0x, 0 ), g.W - 1 ), ccy = min( max( w.c0y, 0 ), g.H - 1 ), ccz = min( max( w.c0z, 0 ), g.D - 1 );
        const bool in_block = w.lox >= ccx - 1 && w.lox + w.nx <= ccx + 2 && w.loy >= ccy - 1 && w.loy + w.ny <= ccy + 2 &&
                              w.loz >= ccz - 1 && w.loz + w.nz <= ccz + 2;
        active = true;
        if( in_block )
        {
          const uint32_t ccid = ( (uint32_t)ccz * g.H + ccy ) * g.W + ccx;
          if( g.ncone )
          {
            const float4 u = __ldg( g.ncone + ccid );
            if( u.w > 1.5f ) { active = false; }
            else
            {
              const ConeCull cull = make_cull( g, sp.dot_thr, nx, ny, nz );
              active = cone_possible_loaded( u, cull, nx, ny, nz );
            }
          }
          else { active = __ldg( g.occ27 + ccid ) != 0; }
          // the points of the block (a superset of the window's) all lie in nbox: farther than the radius from it = no neighbour
          if( active && g.nbox )
          {
            const float4 lo = __ldg( g.nbox + 2 * (size_t)ccid ), hi = __ldg( g.nbox + 2 * (size_t)ccid + 1 );
            active = box_gap_bits( lo, hi, px, py, pz ) < __float_as_uint( sp.r2f );
          }
          // sub-bin = (octant of the query inside its home cell, class of its normal).  The 32 queries a warp of the search
          // takes are consecutive in this order: same octant = the same three faces / three edges / one corner of the
          // block are near, so the lanes agree on which cells are worth sweeping; same dominant axis and sign of the
          // normal = the per-cell normal-cone test skips the same cells.  (Order only: results do not depend on it.)
          const float ax = fabsf( nx ), ay = fabsf( ny ), az = fabsf( nz );
          const int cls = ay >= ax && ay >= az ? ( ny < 0.f ? 1 : 0 ) : ( ax >= az ? ( nx < 0.f ? 3 : 2 ) : ( nz < 0.f ? 5 : 4 ) );
          const float icell = (float)g.inv_cell;
          const int oct = ( w.qx * icell - (float)ccx >= 0.5f ? 1 : 0 ) | ( w.qy * icell - (float)ccy >= 0.5f ? 2 : 0 ) | ( w.qz * icell - (float)ccz >= 0.5f ? 4 : 0 );
          const int sub = sub_mode == 1 ? cls : ( sub_mode == 2 ? oct * 8 : oct * 8 + cls );
          if( active ) { bin = __ldg( g.crank + ccid ) * DB_NCLS + (uint32_t)sub; }
        }
      }
    }
    const unsigned bal = __ballot_sync( RS_FULL, active );
    if( active )
    {
      const int o = n_list + __popc( bal & ( ( 1u << lane ) - 1u ) );
      lcell[o] = bin; lidx[o] = (uint16_t)i;
    }
    n_list += __popc( bal );
  }
  __syncwarp();
  // bound pruning: every term is <= 1 (pose_proposal.cpp:149-152), so fewer survivors than threshold * N cannot pass
  const bool pruned = prune_cnt >= 0.0 && (double)n_list < prune_cnt;
  unsigned base = 0;
  if( !pruned && n_list > 0 && lane == 0 ) { base = atomicAdd( &ctr->n_queue, (unsigned)n_list ); }
  base = __shfl_sync( RS_FULL, base, 0 );
  if( lane == 0 ) { pose_base[pl] = base; pose_cnt[pl] = pruned ? 0u : (unsigned)n_list; }
  if( pruned ) { return; }
  for( int j = lane; j < n_list; j += 32 )
  {
    const uint32_t bin = lcell[j];
    const uint32_t at = atomicAdd( bins + bin, 1u );
    queue[base + j] = make_uint2( base + (unsigned)j, ( pl << DB_PTS_BITS ) | (uint32_t)lidx[j] );
    qbin[base + j] = make_uint2( bin, at );
  }
}

// ------------------------------------------------------------------------------------------------ S: items + scatter
// bins has been scanned (exclusive) into offs; one thread per cell appends the cell's work items
__global__ void db_items_kernel( const uint32_t* __restrict__ offs, uint32_t n_cells /* active */, DbCounters* __restrict__ ctr, uint2* __restrict__ items )
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; // rank of an active cell
  if( c >= n_cells ) { return; }
  const uint32_t a = offs[(size_t)c * DB_NCLS], b = offs[(size_t)( c + 1 ) * DB_NCLS];
  if( a == b ) { return; }
  const uint32_t n = ( b - a + DB_QCHUNK - 1 ) / DB_QCHUNK;
  const uint32_t at = atomicAdd( &ctr->n_items, n );
  for( uint32_t k = 0; k < n; ++k ) { items[at + k] = make_uint2( c, a + k * DB_QCHUNK ); }
}

__global__ void db_scatter_kernel( const DbCounters* __restrict__ ctr, const uint32_t* __restrict__ offs, const uint2* __restrict__ queue,
                                   const uint2* __restrict__ qbin, uint2* __restrict__ sorted )
{
  const unsigned n = ctr->n_queue;
  for( unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x )
  {
    const uint2 b = qbin[e];
    sorted[offs[b.x] + b.y] = queue[e];
  }
}

// -DRS_DB_STATS: work census of the search kernel (diagnostic builds only; printed by dense_binned_run)
#ifdef RS_DB_STATS
__device__ unsigned long long g_db_stats[8];
#define DB_STAT( i, v ) atomicAdd( &g_db_stats[i], (unsigned long long)( v ) )
#else
#define DB_STAT( i, v ) do { } while( 0 )
#endif

// ------------------------------------------------------------------------------------------------ B: staged search
__device__ __forceinline__ uint32_t db_smem_u32( const void* p ) { return (uint32_t)__cvta_generic_to_shared( p ); }
__device__ __forceinline__ void db_mbar_init( uint64_t* bar, unsigned count )
{
  asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( db_smem_u32( bar ) ), "r"( count ) );
}
__device__ __forceinline__ void db_mbar_expect_tx( uint64_t* bar, unsigned bytes )
{
  asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( db_smem_u32( bar ) ), "r"( bytes ) : "memory" );
}
__device__ __forceinline__ void db_mbar_arrive( uint64_t* bar )
{
  asm volatile( "mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"( db_smem_u32( bar ) ) : "memory" );
}
#ifdef RS_DB_WAIT_HINT
// consumers: the probe carries a suspend-time hint as well, so a warp that arrives before its block is staged sleeps in the
// barrier unit instead of spinning through issue slots the other warps could use
__device__ __forceinline__ void db_mbar_wait( uint64_t* bar, unsigned parity )
{
  unsigned ok = 0;
  while( !ok )
  {
    asm volatile( "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                  : "=r"( ok ) : "r"( db_smem_u32( bar ) ), "r"( parity ), "r"( (unsigned)RS_DB_WAIT_HINT ) : "memory" );
  }
}
#else
__device__ __forceinline__ void db_mbar_wait( uint64_t* bar, unsigned parity )
{
  asm volatile( "{\n\t.reg .pred p;\n\tDB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DB_DONE;\n\tbra DB_WAIT;\n\tDB_DONE:\n\t}"
                ::"r"( db_smem_u32( bar ) ), "r"( parity ) : "memory" );
}
#endif
// the producer warp waits for whole items to be consumed (tens of microseconds): the probe carries a suspend-time hint, so
// the hardware parks the warp instead of letting it spin through issue slots; one lane probes, the warp follows
__device__ __forceinline__ void db_mbar_wait_parked( uint64_t* bar, unsigned parity )
{
  if( ( threadIdx.x & 31 ) == 0 )
  {
    unsigned ok = 0;
    while( !ok )
    {
      asm volatile( "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"( ok ) : "r"( db_smem_u32( bar ) ), "r"( parity ), "r"( 20000u ) : "memory" );
    }
  }
  __syncwarp();
}
// 1-D bulk-async copy global -> shared (bytes a multiple of 16, both addresses 16-byte aligned), completion on the mbarrier
__device__ __forceinline__ void db_bulk_g2s( void* dst, const void* src, unsigned bytes, uint64_t* bar )
{
  asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( db_smem_u32( dst ) ), "l"( src ),
                "r"( bytes ), "r"( db_smem_u32( bar ) ) : "memory" );
}

struct DbBlock // one staged 3x3x3 block (per buffer)
{
  uint32_t cell;       // home cell id, DB_DONE = no more items
  uint32_t q_begin, q_end;
  uint32_t staged;     // 1: records / normals of the block are in shared memory, 0: block larger than DB_CAP, read from global
  uint32_t gs[27];     // per cell (dz, dy, dx order): first record (position in recs)
  uint32_t cn[27];     // number of records
  uint32_t so[27];     // offset in the staged arrays
  float4 cone[27];     // per cell {unit mean normal, cos(widest deviation)} (w <= 0: no usable cone), nearest.cuh ConeCull
  float4 blo[27], bhi[27]; // per cell: bounding box of its points (rsgpu_internal.cuh box_gap_bits)
};

struct DbSmem
{
  float4 recs[2][DB_CAP];
  float4 nrm[2][DB_CAP];
  DbBlock blk[2];
  uint64_t full[2], empty[2];
};

// visiting order of the 27 cells (index (dz*3 + dy)*3 + dx): own cell, faces, edges, corners
__constant__ unsigned char kDbOrder[27] = { 13, 12, 14, 10, 16, 4, 22, 9, 11, 15, 17, 3, 5, 21, 23, 1, 7, 19, 25, 0, 2, 6, 8, 18, 20, 24, 26 };

__global__ void __launch_bounds__( DB_THREADS, RS_DB_BPS ) db_search_kernel( GridView g, DbPoseGrid pg, long long pose0, ScoreParams sp, DbCounters* __restrict__ ctr,
                                                                  const uint2* __restrict__ items, const uint32_t* __restrict__ offs,
                                                                  const uint2* __restrict__ sorted, double* __restrict__ terms )
{
  extern __shared__ __align__( 128 ) unsigned char db_smem_raw[];
  DbSmem& S = *(DbSmem*)db_smem_raw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if( threadIdx.x == 0 )
  {
    for( int b = 0; b < 2; ++b ) { db_mbar_init( &S.full[b], 1 ); db_mbar_init( &S.empty[b], DB_WARPS ); }
    asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
  }
  __syncthreads();
  const unsigned n_items = ctr->n_items;

  if( warp == DB_WARPS )
  {
    // ---------------- producer warp: claim items, resolve the block's cell ranges, start the bulk copies
    for( unsigned it = 0;; ++it )
    {
      const int b = it & 1;
      if( it >= 2 ) { db_mbar_wait_parked( &S.empty[b], ( ( it >> 1 ) - 1 ) & 1 ); } // the compute warps left this buffer
      unsigned item = 0;
      if( lane == 0 ) { item = atomicAdd( &ctr->item_cursor, 1u ); }
      item = __shfl_sync( RS_FULL, item, 0 );
      DbBlock& B = S.blk[b];
      if( item >= n_items )
      {
        if( lane == 0 ) { B.cell = DB_DONE; db_mbar_arrive( &S.full[b] ); }
        break;
      }
      const uint2 im = __ldg( items + item );
      const uint32_t c0 = __ldg( g.acells + im.x ); // item = (rank of the active cell, first query)
      const int c0x = (int)( c0 % (uint32_t)g.W ); const uint32_t rr = c0 / (uint32_t)g.W;
      const int c0y = (int)( rr % (uint32_t)g.H ), c0z = (int)( rr / (uint32_t)g.H );
      uint32_t s = 0, n = 0;
      float4 cone = make_float4( 0.f, 0.f, 0.f, -1.f );
      const float ninf = __int_as_float( 0xff800000 );
      float4 blo = make_float4( ninf, ninf, ninf, 0.f ), bhi = make_float4( -ninf, -ninf, -ninf, 0.f ); // no table: everything is "inside"
      if( lane < 27 )
      {
        const int cx = c0x + lane % 3 - 1, cy = c0y + ( lane / 3 ) % 3 - 1, cz = c0z + lane / 9 - 1;
        if( cx >= 0 && cx < g.W && cy >= 0 && cy < g.H && cz >= 0 && cz < g.D )
        {
          const uint32_t id = ( (uint32_t)cz * g.H + cy ) * g.W + cx;
          s = __ldg( g.cell_start + id ); n = __ldg( g.cell_start + id + 1 ) - s;
          if( g.cone && n ) { cone = __ldg( g.cone + id ); }
          if( g.cbox && n ) { blo = __ldg( g.cbox + 2 * (size_t)id ); bhi = __ldg( g.cbox + 2 * (size_t)id + 1 ); }
        }
      }
      // compact layout in lane (= cell) order: exclusive prefix of the counts
      uint32_t incl = n;
#pragma unroll
      for( int o = 1; o < 32; o <<= 1 ) { const uint32_t v = __shfl_up_sync( RS_FULL, incl, o ); if( lane >= o ) { incl += v; } }
      const uint32_t total = __shfl_sync( RS_FULL, incl, 31 );
      const uint32_t so = incl - n;
      const bool staged = total <= (uint32_t)DB_CAP;
      if( lane < 27 ) { B.gs[lane] = s; B.cn[lane] = n; B.so[lane] = so; B.cone[lane] = cone; B.blo[lane] = blo; B.bhi[lane] = bhi; }
      if( lane == 0 )
      {
        B.cell = c0; B.q_begin = im.y; const uint32_t qe = __ldg( offs + (size_t)( im.x + 1 ) * DB_NCLS ); B.q_end = min( im.y + (uint32_t)DB_QCHUNK, qe );
        B.staged = staged ? 1u : 0u;
      }
      __syncwarp(); // the block description of all lanes is ordered before lane 0's (releasing) arrive
      if( lane == 0 )
      {
        if( staged && total > 0 ) { db_mbar_expect_tx( &S.full[b], total * 32u ); } else { db_mbar_arrive( &S.full[b] ); }
      }
      __syncwarp();
      if( staged && total > 0 )
      {
        // an x-row of the block (3 cells, fewer at the grid's x faces) is one contiguous range of records: cells inside a
        // row are consecutive in the dense table and the compact layout keeps their order.  Lanes 0, 3, 6, ... (dx = 0)
        // copy their row.  Rows clipped at a y / z face have n = 0 for all three lanes.
        const uint32_t n1 = __shfl_down_sync( RS_FULL, n, 1 ), n2 = __shfl_down_sync( RS_FULL, n, 2 );
        const uint32_t s1 = __shfl_down_sync( RS_FULL, s, 1 ), s2 = __shfl_down_sync( RS_FULL, s, 2 );
        if( lane < 27 && lane % 3 == 0 )
        {
          const uint32_t rn = n + n1 + n2;
          if( rn )
          {
            const uint32_t rs = n ? s : ( n1 ? s1 : s2 ); // first non-empty cell of the row (empty cells have s = 0 when clipped)
            db_bulk_g2s( &S.recs[b][so], g.recs + rs, rn * 16u, &S.full[b] );
            db_bulk_g2s( &S.nrm[b][so], g.nrm + rs, rn * 16u, &S.full[b] );
          }
        }
      }
    }
    return;
  }

  // ---------------- compute warps
  const uint32_t r2bits = __float_as_uint( sp.r2f );
  const uint32_t uk = (uint32_t)sp.k;
  for( unsigned it = 0;; ++it )
  {
    const int b = it & 1;
    db_mbar_wait( &S.full[b], ( it >> 1 ) & 1 );
    const DbBlock& B = S.blk[b];
    if( B.cell == DB_DONE ) { break; }
    const uint32_t c0 = B.cell;
    const int c0x = (int)( c0 % (uint32_t)g.W ); const uint32_t rr = c0 / (uint32_t)g.W;
    const int c0y = (int)( rr % (uint32_t)g.H ), c0z = (int)( rr / (uint32_t)g.H );
    const float4* __restrict__ recs = B.staged ? S.recs[b] : g.recs;
    const float4* __restrict__ nrm = B.staged ? S.nrm[b] : g.nrm;
    const bool staged = B.staged != 0;
    for( uint32_t q0 = B.q_begin + warp * 32; q0 < B.q_end; q0 += DB_WARPS * 32 )
    {
      const uint32_t qi = q0 + lane;
      const bool qv = qi < B.q_end;
      uint2 ent = make_uint2( 0u, 0u );
      float px = 0.f, py = 0.f, pz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
      unsigned inwin = 0; // bit e: cell e of the block lies in this query's window
      float gx[3] = { 0.f, 0.f, 0.f }, gy[3] = { 0.f, 0.f, 0.f }, gz[3] = { 0.f, 0.f, 0.f }; // squared per-axis gap to block column d
      ConeCull cull; cull.on = false; cull.cb = 0.f; cull.sb = 1.f;
      if( qv )
      {
        ent = __ldg( sorted + qi );
        const long long pose = pose0 + ( ent.y >> DB_PTS_BITS );
        const long long t = pose / pg.n_rot; const int r = (int)( pose - t * pg.n_rot );
        db_query( pg, t, r, (int)( ent.y & ( DB_MAX_PTS - 1 ) ), px, py, pz, nx, ny, nz );
        // window and per-axis squared gaps exactly as rsg::group_search (msh_hash_grid.h:1150-1198): a cell below the
        // query's own cell gets the gap to the own cell's lower face, a cell above it the gap to the upper face (for cells
        // further out - only possible for queries outside the grid - an underestimate, which is conservative)
        const CellWindow w = make_window( g, px, py, pz, sp.radius );
        float a;
        a = (float)__dsub_rn( (double)w.qx, __dmul_rn( (double)w.c0x, g.cell ) ); const float glx2 = __fmul_rn( a, a );
        a = (float)__dsub_rn( __dmul_rn( (double)( w.c0x + 1 ), g.cell ), (double)w.qx ); const float ghx2 = __fmul_rn( a, a );
        a = (float)__dsub_rn( (double)w.qy, __dmul_rn( (double)w.c0y, g.cell ) ); const float gly2 = __fmul_rn( a, a );
        a = (float)__dsub_rn( __dmul_rn( (double)( w.c0y + 1 ), g.cell ), (double)w.qy ); const float ghy2 = __fmul_rn( a, a );
        a = (float)__dsub_rn( (double)w.qz, __dmul_rn( (double)w.c0z, g.cell ) ); const float glz2 = __fmul_rn( a, a );
        a = (float)__dsub_rn( __dmul_rn( (double)( w.c0z + 1 ), g.cell ), (double)w.qz ); const float ghz2 = __fmul_rn( a, a );
        // block column d in {0,1,2} is grid column c0 + d - 1 (c0 = the block's centre = the query's own cell clamped into the grid)
        unsigned mx = 0, my = 0, mz = 0;
#pragma unroll
        for( int d = 0; d < 3; ++d )
        {
          const int cx = c0x + d - 1, cy = c0y + d - 1, cz = c0z + d - 1;
          if( cx >= w.lox && cx < w.lox + w.nx ) { mx |= 1u << d; }
          if( cy >= w.loy && cy < w.loy + w.ny ) { my |= 1u << d; }
          if( cz >= w.loz && cz < w.loz + w.nz ) { mz |= 1u << d; }
          gx[d] = cx < w.c0x ? glx2 : ( cx > w.c0x ? ghx2 : 0.0f );
          gy[d] = cy < w.c0y ? gly2 : ( cy > w.c0y ? ghy2 : 0.0f );
          gz[d] = cz < w.c0z ? glz2 : ( cz > w.c0z ? ghz2 : 0.0f );
        }
        // 27-bit mask = outer product of the three 3-bit masks
        const unsigned row = mx * ( ( my & 1u ) | ( ( my & 2u ) << 2 ) | ( ( my & 4u ) << 4 ) ); // 9 bits: (dy, dx)
        inwin = row * ( ( mz & 1u ) | ( ( mz & 2u ) << 8 ) | ( ( mz & 4u ) << 16 ) );
        cull = make_cull( g, sp.dot_thr, nx, ny, nz );
      }
      // ---- sweep: nearest compatible point, key = (d2 bits, record position).  closer = an upper bound of the number of
      // points strictly closer than the final winner: points met at or below the running best, plus whole cells that were
      // skipped for their normals only
      unsigned long long best = (unsigned long long)r2bits << 32;
      float bestdot = 0.f;
      uint32_t closer = 0;
#pragma unroll 1
      for( int eo = 0; eo < 27; ++eo )
      {
        const int e = kDbOrder[eo];
        const uint32_t cn = B.cn[e];
        if( cn == 0 ) { continue; }
        const int dx = e % 3, dy = ( e / 3 ) % 3, dz = e / 9;
        // (gz*gz + gy*gy) + gx*gx (:1221), 5 mantissa bits dropped: conservative
        const uint32_t gapc = __float_as_uint( __fadd_rn( __fadd_rn( dz == 0 ? gz[0] : ( dz == 1 ? gz[1] : gz[2] ), dy == 0 ? gy[0] : ( dy == 1 ? gy[1] : gy[2] ) ),
                                                          dx == 0 ? gx[0] : ( dx == 1 ? gx[1] : gx[2] ) ) ) & 0xffffffe0u;
        bool want = ( ( inwin >> e ) & 1u ) && gapc < (uint32_t)( best >> 32 );
        if( want ) { want = box_gap_bits( B.blo[e], B.bhi[e], px, py, pz ) < (uint32_t)( best >> 32 ); } // all its points are farther
        if( want && !cone_possible_loaded( B.cone[e], cull, nx, ny, nz ) ) { want = false; closer += cn; } // none can be compatible
        if( !__any_sync( RS_FULL, want ) ) { continue; }
#ifdef RS_DB_STATS
        if( want ) { DB_STAT( 2, 1 ); DB_STAT( 3, cn ); }
        if( lane == 0 ) { DB_STAT( 4, 1 ); DB_STAT( 5, cn ); }
#endif
        const uint32_t gs = B.gs[e];
        const float4* __restrict__ rp = recs + ( staged ? B.so[e] : gs );
        const float4* __restrict__ np = nrm + ( staged ? B.so[e] : gs );
        if( want )
        {
#pragma unroll 4
          for( uint32_t j = 0; j < cn; ++j )
          {
            const float4 rec = rp[j]; // all lanes read the same address: one broadcast
            const uint32_t db = __float_as_uint( dist2_exact( rec, px, py, pz ) );
            if( db <= (uint32_t)( best >> 32 ) )
            {
              ++closer;
              const unsigned long long key = ( (unsigned long long)db << 32 ) | ( gs + j );
              if( key < best )
              {
                const float4 mm = np[j];
                const float dot = dot3_exact( mm.x, mm.y, mm.z, nx, ny, nz );
                if( dot >= sp.dot_thr && dot <= 1.0f ) { best = key; bestdot = dot; }
              }
            }
          }
        }
      }
      const uint32_t dcb = (uint32_t)( best >> 32 );
      bool found = qv && dcb < r2bits;
#ifdef RS_DB_STATS
      if( qv ) { DB_STAT( 0, 1 ); DB_STAT( 1, found ? 1 : 0 ); }
      if( lane == 0 ) { DB_STAT( 6, 1 ); }
      if( found && closer >= uk ) { DB_STAT( 7, 1 ); }
#endif
      // ---- rank of the winner under the k-cap (fewer than k points strictly closer): exact count only where the bound allows k
      const bool need = found && closer >= uk;
      if( __any_sync( RS_FULL, need ) )
      {
        const float dcf = __uint_as_float( dcb );
        uint32_t cnt = 0;
#pragma unroll 1
        for( int e = 0; e < 27; ++e )
        {
          const uint32_t cn = B.cn[e];
          if( cn == 0 ) { continue; }
          const int dx = e % 3, dy = ( e / 3 ) % 3, dz = e / 9;
          const uint32_t gapc = __float_as_uint( __fadd_rn( __fadd_rn( dz == 0 ? gz[0] : ( dz == 1 ? gz[1] : gz[2] ), dy == 0 ? gy[0] : ( dy == 1 ? gy[1] : gy[2] ) ),
                                                            dx == 0 ? gx[0] : ( dx == 1 ? gx[1] : gx[2] ) ) ) & 0xffffffe0u;
          bool want = need && ( ( inwin >> e ) & 1u ) && gapc < dcb;
          if( want ) { want = box_gap_bits( B.blo[e], B.bhi[e], px, py, pz ) < dcb; }
          if( !__any_sync( RS_FULL, want ) ) { continue; }
          const float4* __restrict__ rp = recs + ( staged ? B.so[e] : B.gs[e] );
          if( want )
          {
#pragma unroll 4
            for( uint32_t j = 0; j < cn; ++j ) { cnt += dist2_exact( rp[j], px, py, pz ) < dcf; }
          }
        }
        if( need ) { found = cnt < uk; }
      }
      // ---- the reference's per-point term (pose_proposal.cpp:149-152) in fp64
      if( qv )
      {
        double term = 0.0;
        if( found )
        {
          const double angle = acos( (double)fmaxf( bestdot, 0.0f ) );
          const double nc = exp( -( angle * angle ) / ( 2.0 * 0.5 * 0.5 ) );
          const double dc = exp( -(double)__uint_as_float( dcb ) / sp.inv_two_sigma_sq_den );
          term = 0.05 * nc + ( 1.0 - 0.05 ) * dc;
        }
        terms[ent.x] = term;
      }
    }
    __syncwarp();
    if( lane == 0 ) { db_mbar_arrive( &S.empty[b] ); }
  }
}

// queries of the fallback bin (home cell outside the grid / window not inside the 3x3x3 block): generic warp-cooperative search
__global__ void __launch_bounds__( 128 ) db_fallback_kernel( GridView g, DbPoseGrid pg, long long pose0, ScoreParams sp, const uint32_t* __restrict__ offs,
                                                             uint32_t n_cells, const uint2* __restrict__ sorted, double* __restrict__ terms )
{
  const int lane = threadIdx.x & 31;
  const uint32_t a = offs[(size_t)n_cells * DB_NCLS], b = offs[(size_t)n_cells * DB_NCLS + 1];
  const unsigned warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5, n_warps = ( gridDim.x * blockDim.x ) >> 5;
  for( uint32_t q0 = a + warp * 32; q0 < b; q0 += n_warps * 32 )
  {
    const uint32_t qi = q0 + lane;
    const bool qv = qi < b;
    uint2 ent = make_uint2( 0u, 0u );
    LaneQuery q;
    q.px = q.py = q.pz = q.nx = q.ny = q.nz = 0.f;
    if( qv )
    {
      ent = __ldg( sorted + qi );
      const long long pose = pose0 + ( ent.y >> DB_PTS_BITS );
      const long long t = pose / pg.n_rot; const int r = (int)( pose - t * pg.n_rot );
      db_query( pg, t, r, (int)( ent.y & ( DB_MAX_PTS - 1 ) ), q.px, q.py, q.pz, q.nx, q.ny, q.nz );
    }
    const NearestHit h = nearest_compatible_batch<false>( g, q, qv, sp.radius, sp.r2f, sp.dot_thr, sp.k, nullptr );
    if( qv )
    {
      double term = 0.0;
      if( h.found )
      {
        const double angle = acos( (double)fmaxf( h.dot, 0.0f ) );
        const double nc = exp( -( angle * angle ) / ( 2.0 * 0.5 * 0.5 ) );
        const double dc = exp( -(double)h.d2 / sp.inv_two_sigma_sq_den );
        term = 0.05 * nc + ( 1.0 - 0.05 ) * dc;
      }
      terms[ent.x] = term;
    }
  }
}

// ------------------------------------------------------------------------------------------------ C: ordered sums
// one warp per pose: the terms of the pose's survivors, in point order, summed like the reference's loop (:127-157)
__global__ void __launch_bounds__( 128 ) db_reduce_kernel( const uint32_t* __restrict__ pose_base, const uint32_t* __restrict__ pose_cnt,
                                                           const double* __restrict__ terms, unsigned n_chunk_poses, int n_obj,
                                                           float* __restrict__ scores )
{
  const int lane = threadIdx.x & 31;
  const unsigned pl = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5;
  if( pl >= n_chunk_poses ) { return; }
  const unsigned base = pose_base[pl], n = pose_cnt[pl];
  double sum = 0.0;
  for( unsigned j0 = 0; j0 < n; j0 += 32 )
  {
    const unsigned j = j0 + lane;
    const double term = j < n ? terms[base + j] : 0.0;
    unsigned mask = __ballot_sync( RS_FULL, term != 0.0 );
    while( mask )
    {
      const int src = __ffs( mask ) - 1; mask &= mask - 1;
      sum += __shfl_sync( RS_FULL, term, src );
    }
  }
  if( lane == 0 ) { scores[pl] = (float)( sum / (double)n_obj ); } // (:156)
}

// scratch of the dense launches of one calling thread (= lane).  Kept between calls and only ever grown: ~1.4 GB per lane at
// C3 size, and handing blocks of that size back to the stream-ordered pool after every launch makes the pool serve one
// lane's next allocation from another lane's freed block - which it orders behind that lane's pending work
// (cudaMemPoolReuseAllowInternalDependencies), i.e. a cross-lane dependency nobody asked for.  "dense_scratch" = "pool"
// restores per-launch allocation (A/B).
struct DbScratch
{
  DevBuf<float4> upos, unor;
  DevBuf<uint32_t> bins, offs, pose_base, pose_cnt;
  DevBuf<uint2> queue, qbin, sorted, items;
  DevBuf<double> terms;
  DevBuf<DbCounters> ctr;
  DevBuf<unsigned char> scan_tmp;
};

// Sizes of one dense launch: the pose grid is processed in chunks of whole translations such that the worst case of a
// chunk (every point of every pose survives the prefilter) fits the queue - nothing can overflow, nothing is re-tried.
struct DbPlan
{
  size_t n_cells = 0, n_active = 0, n_bins = 0, entries_max = 0, items_max = 0, chunk_poses_max = 0, scan_bytes = 0;
  long long trans_per_chunk = 0, n_trans = 0;
};

bool dense_binned_supported( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const PoseSource& ps )
{
  const size_t n_cells = (size_t)scene->info.width * scene->info.height * scene->info.depth;
  (void)n_cells;
  return obj->n >= 1 && obj->n <= DB_MAX_PTS && ps.n_rot >= 1 && ps.n_rot <= 4096 && (size_t)scene->n_active * DB_NCLS + 2 < ( (size_t)1 << 28 ) &&
         scene->has_normals && scene->info.n_pts > 0 && scene->crank.p && scene->n_active > 0;
}

// allocations of a launch (on the calling thread's stream: BEFORE the launch is forked to the bulk stream, so that the
// stream-ordered allocator sees allocation -> use -> release in one order)
int dense_binned_alloc( DbScratch& S, DbPlan& P, const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const PoseSource& ps, long long n_poses )
{
  const int n = obj->n, n_rot = ps.n_rot;
  P.n_cells = (size_t)scene->info.width * scene->info.height * scene->info.depth;
  P.n_trans = n_poses / n_rot;
  P.n_active = scene->n_active;
  P.n_bins = P.n_active * DB_NCLS + 2; // + the fallback bin + the end of the scan
  size_t cap = (size_t)64 << 20; // queue entries per chunk (32 B of scratch each, per lane): C2 objects and an eighth of C3 go through in one chunk
  {
    const std::string o = option( "dense_cap" );
    if( !o.empty() ) { cap = (size_t)atoll( o.c_str() ); }
    if( cap < (size_t)n * n_rot ) { cap = (size_t)n * n_rot; }
  }
  long long tpc = (long long)( cap / ( (size_t)n * n_rot ) );
  if( tpc * n_rot > (long long)DB_POSE_MAX ) { tpc = DB_POSE_MAX / n_rot; }
  if( tpc < 1 ) { tpc = 1; }
  if( tpc > P.n_trans ) { tpc = P.n_trans; }
  P.trans_per_chunk = tpc;
  P.chunk_poses_max = (size_t)tpc * n_rot;
  P.entries_max = P.chunk_poses_max * (size_t)n;
  P.items_max = P.entries_max / DB_QCHUNK + P.n_active + 2;
  // capacities that do not depend on the object, so that a lane's scratch is allocated ONCE per scan size: every object's
  // worst case is just under `cap` entries, each by a different margin - growing the buffers by those margins meant
  // freeing and re-allocating a gigabyte per lane whenever a slightly bigger object came along (steps of 100-500 ms
  // instead of 33 on C2 until every lane had met the biggest one)
  const size_t entries_alloc = std::max( cap, P.entries_max );
  const size_t poses_alloc = std::max( (size_t)DB_POSE_MAX, P.chunk_poses_max );
  const size_t items_alloc = entries_alloc / DB_QCHUNK + P.n_active + 2;
  const size_t cloud_alloc = std::max( (size_t)n * n_rot, (size_t)DB_MAX_PTS * 128 );
  RS_CUDA( S.upos.reserve( cloud_alloc ) ); RS_CUDA( S.unor.reserve( cloud_alloc ) );
  RS_CUDA( S.bins.reserve( P.n_bins ) ); RS_CUDA( S.offs.reserve( P.n_bins ) );
  RS_CUDA( S.pose_base.reserve( poses_alloc ) ); RS_CUDA( S.pose_cnt.reserve( poses_alloc ) );
  RS_CUDA( S.queue.reserve( entries_alloc ) ); RS_CUDA( S.qbin.reserve( entries_alloc ) ); RS_CUDA( S.sorted.reserve( entries_alloc ) );
  RS_CUDA( S.items.reserve( items_alloc ) ); RS_CUDA( S.terms.reserve( entries_alloc ) ); RS_CUDA( S.ctr.reserve( 1 ) );
  RS_CUDA( cub::DeviceScan::ExclusiveSum( nullptr, P.scan_bytes, S.bins.p, S.offs.p, (int64_t)P.n_bins, rt().stream ) );
  RS_CUDA( S.scan_tmp.reserve( P.scan_bytes ) );
  return RSGPU_OK;
}

// scores (device) [n_rot * n_trans] of the dense pose grid, every launch on `st`
int dense_binned_run( DbScratch& S, const DbPlan& P, const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const PoseSource& ps, const ScoreParams& sp,
                      double prune_thr, float* d_scores, cudaStream_t st )
{
  const GridView g = scene->view();
  const size_t n_cells = P.n_active; // everything below indexes bins by active rank
  int sub_mode = 0; // "dense_sub": default octant x normal class; "n" = normal class only; "o" = octant only (A/B)
  { const std::string o = option( "dense_sub" ); sub_mode = o == "n" ? 1 : ( o == "o" ? 2 : 0 ); }
  const int n = obj->n, n_rot = ps.n_rot;
  size_t scan_bytes = P.scan_bytes;
  db_prepare_kernel<<<( n * n_rot + 255 ) / 256, 256, 0, st>>>( obj->pos.p, obj->nor.p, n, ps.rots, n_rot, S.upos.p, S.unor.p );
  RS_CHECK_LAUNCH();
  DbPoseGrid pg; pg.upos = S.upos.p; pg.unor = S.unor.p; pg.trans = ps.trans; pg.n_rot = n_rot; pg.n = n;
  const double prune_cnt = prune_thr > 0.0 ? prune_thr * (double)n / 1.000001 : -1.0;
  const int n_pad = ( n + 31 ) / 32 * 32;
  const int pf_warps = n_pad <= 2048 ? 4 : 2;
  const size_t pf_smem = (size_t)pf_warps * n_pad * 6;
  int dev_sms = 148;
  cudaDeviceGetAttribute( &dev_sms, cudaDevAttrMultiProcessorCount, rt().device );
  {
    static std::once_flag once; // function attributes are per process
    cudaError_t ae = cudaSuccess;
    std::call_once( once, [&]() {
      ae = cudaFuncSetAttribute( db_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof( DbSmem ) );
      if( ae == cudaSuccess ) { ae = cudaFuncSetAttribute( db_prefilter_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 2048 * 6 ); }
      if( ae == cudaSuccess ) { ae = cudaFuncSetAttribute( db_prefilter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * DB_MAX_PTS * 6 ); }
    } );
    RS_CUDA( ae );
  }
  int search_blocks_per_sm = RS_DB_BPS;
  {
    const std::string o = option( "dense_bps" );
    if( !o.empty() ) { search_blocks_per_sm = std::max( 1, atoi( o.c_str() ) ); }
  }
  for( long long t0 = 0; t0 < P.n_trans; t0 += P.trans_per_chunk )
  {
    const long long nt = std::min( P.trans_per_chunk, P.n_trans - t0 );
    const long long pose0 = t0 * n_rot;
    const unsigned n_chunk = (unsigned)( nt * n_rot );
    RS_CUDA( cudaMemsetAsync( S.bins.p, 0, sizeof( uint32_t ) * P.n_bins, st ) );
    RS_CUDA( cudaMemsetAsync( S.ctr.p, 0, sizeof( DbCounters ), st ) );
    {
      ProfScope prof( "dense_prefilter", st );
      const unsigned blocks = ( n_chunk + pf_warps - 1 ) / pf_warps;
      if( pf_warps == 4 )
      {
        db_prefilter_kernel<4><<<blocks, 128, pf_smem, st>>>( g, pg, pose0, n_chunk, sp, prune_cnt, n_pad, sub_mode, S.ctr.p, S.bins.p, S.pose_base.p, S.pose_cnt.p,
                                                              S.queue.p, S.qbin.p );
      }
      else
      {
        db_prefilter_kernel<2><<<blocks, 64, pf_smem, st>>>( g, pg, pose0, n_chunk, sp, prune_cnt, n_pad, sub_mode, S.ctr.p, S.bins.p, S.pose_base.p, S.pose_cnt.p,
                                                             S.queue.p, S.qbin.p );
      }
      RS_CHECK_LAUNCH();
    }
    {
      ProfScope prof( "dense_bin", st );
      RS_CUDA( cub::DeviceScan::ExclusiveSum( S.scan_tmp.p, scan_bytes, S.bins.p, S.offs.p, (int64_t)P.n_bins, st ) );
      db_items_kernel<<<(unsigned)( ( n_cells + 255 ) / 256 ), 256, 0, st>>>( S.offs.p, (uint32_t)n_cells, S.ctr.p, S.items.p );
      RS_CHECK_LAUNCH();
      db_scatter_kernel<<<dev_sms * 8, 256, 0, st>>>( S.ctr.p, S.offs.p, S.queue.p, S.qbin.p, S.sorted.p );
      RS_CHECK_LAUNCH();
    }
    {
      ProfScope prof( "dense_search", st );
      db_search_kernel<<<dev_sms * search_blocks_per_sm, DB_THREADS, sizeof( DbSmem ), st>>>( g, pg, pose0, sp, S.ctr.p, S.items.p, S.offs.p, S.sorted.p, S.terms.p );
      RS_CHECK_LAUNCH();
      db_fallback_kernel<<<dev_sms * 2, 128, 0, st>>>( g, pg, pose0, sp, S.offs.p, (uint32_t)n_cells, S.sorted.p, S.terms.p );
      RS_CHECK_LAUNCH();
    }
    if( option( "dense_stats" ) == "1" )
    {
      // diagnostic (synchronises): queue length, items, how full the items' warps are, fallback queries, blocks above the staging capacity
      DbCounters hc; std::vector<uint32_t> hoffs( P.n_bins );
      cudaStreamSynchronize( st );
      cudaMemcpy( &hc, S.ctr.p, sizeof( hc ), cudaMemcpyDeviceToHost );
      cudaMemcpy( hoffs.data(), S.offs.p, sizeof( uint32_t ) * P.n_bins, cudaMemcpyDeviceToHost );
      size_t used = 0, warps = 0, big = 0, mx = 0;
      for( size_t c = 0; c < n_cells; ++c )
      {
        const size_t q = hoffs[( c + 1 ) * DB_NCLS] - hoffs[c * DB_NCLS];
        if( !q ) { continue; }
        ++used; warps += ( q + 31 ) / 32; mx = std::max( mx, q ); big += q > 4096;
      }
      fprintf( stderr, "dense chunk: poses %u points %d queue %u (%.1f%% of worst case) items %u cells_with_queries %zu warp_rounds %zu (fill %.1f%%) "
                       "max_per_cell %zu cells>4096 %zu fallback %u\n", n_chunk, n, hc.n_queue, 100.0 * hc.n_queue / ( (double)n_chunk * n ), hc.n_items, used, warps,
               warps ? 100.0 * ( hoffs[n_cells * DB_NCLS] ) / ( 32.0 * warps ) : 0.0, mx, big, hoffs[n_cells * DB_NCLS + 1] - hoffs[n_cells * DB_NCLS] );
    }
    {
      ProfScope prof( "dense_reduce", st );
      db_reduce_kernel<<<( n_chunk + 3 ) / 4, 128, 0, st>>>( S.pose_base.p, S.pose_cnt.p, S.terms.p, n_chunk, n, d_scores + pose0 );
      RS_CHECK_LAUNCH();
    }
  }
#ifdef RS_DB_STATS
  {
    unsigned long long h[8];
    cudaStreamSynchronize( st );
    cudaMemcpyFromSymbol( h, g_db_stats, sizeof( h ) );
    fprintf( stderr, "dense search stats (cumulative): queries %llu found %llu (%.1f%%) | per query: cells swept %.2f records %.1f | per warp round (%llu rounds): cells %.2f records %.1f | "
                     "lane efficiency %.1f%% | rank counts needed %llu\n", h[0], h[1], 100.0 * h[1] / (double)std::max( h[0], 1ull ), h[2] / (double)std::max( h[0], 1ull ),
             h[3] / (double)std::max( h[0], 1ull ), h[6], h[4] / (double)std::max( h[6], 1ull ), h[5] / (double)std::max( h[6], 1ull ),
             100.0 * h[3] / ( 32.0 * std::max( h[5], 1ull ) ), h[7] );
  }
#endif
  return RSGPU_OK;
}
} // namespace
