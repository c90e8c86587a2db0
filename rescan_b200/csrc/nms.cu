// Voxel-occupancy overlap of one object under two poses and the greedy non-maxima suppression built on it: replaces
// isect_get_overlap_factor (reference lib/rs/intersect.h:309-368) and mgs_non_maxima_suppresion
// (apps/pose_proposal/pose_proposal.cpp:371-452) — SURVEY.md §8 row f1.
//
// The reference rasterises BOTH posed clouds into a grid whose origin and size depend on the pair (union of the two
// posed bounding boxes, fattened by 0.3 m, intersect.h:317-336), so occupancy cannot be cached per pose: every
// (kept pose, candidate) pair is voxelised afresh, exactly as the reference does, but all candidates of one greedy round
// in ONE launch, one thread block per pair:
//   1. both level-1 clouds are transformed and marked into two byte grids (global scratch, L2 resident);
//   2. the reference's scan-line fill (a free cell is inside iff, along its x-row and along its z-row, an odd number
//      of boundary->free transitions lies before it from BOTH ends, intersect.h:129-175) runs one thread per row,
//      first all x-rows, then all z-rows;
//   3. occupied / overlapping cells are counted with one block reduction.
// The greedy loop (arg-max, centroid distance, marks) stays on the host: it is sequential by definition and tiny.
#include "rsgpu_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

using namespace rs;

namespace
{
struct PairDesc
{
  int a, b;              // pose indices
  float origin[3];       // grid min corner (bbox of both poses, minus the fat factor)
  int xr, yr, zr;
  unsigned long long off; // scratch offset of grid A; grid B follows at off + n_cells
};

// bounding box of the level-3 points under each pose (intersect.h:119-130): one block per pose
__global__ void __launch_bounds__( 128 ) posed_bbox_kernel( const float* __restrict__ pos3, int n3, const float* __restrict__ poses, int stride,
                                                            float* __restrict__ boxes )
{
  const float* m = poses + (size_t)blockIdx.x * stride;
  float mn[3] = { 1e9f, 1e9f, 1e9f }, mx[3] = { -1e9f, -1e9f, -1e9f };
  for( int i = threadIdx.x; i < n3; i += blockDim.x )
  {
    float p[3];
    xf_apply( m, pos3[3 * (size_t)i], pos3[3 * (size_t)i + 1], pos3[3 * (size_t)i + 2], 1.0f, p[0], p[1], p[2] );
#pragma unroll
    for( int a = 0; a < 3; ++a ) { mn[a] = fminf( mn[a], p[a] ); mx[a] = fmaxf( mx[a], p[a] ); }
  }
  __shared__ float s_mn[4][3], s_mx[4][3];
#pragma unroll
  for( int a = 0; a < 3; ++a )
  {
    for( int o = 16; o > 0; o >>= 1 )
    {
      mn[a] = fminf( mn[a], __shfl_down_sync( RS_FULL, mn[a], o ) );
      mx[a] = fmaxf( mx[a], __shfl_down_sync( RS_FULL, mx[a], o ) );
    }
    if( ( threadIdx.x & 31 ) == 0 ) { s_mn[threadIdx.x >> 5][a] = mn[a]; s_mx[threadIdx.x >> 5][a] = mx[a]; }
  }
  __syncthreads();
  if( threadIdx.x < 3 )
  {
    const int a = threadIdx.x;
    float lo = s_mn[0][a], hi = s_mx[0][a];
    for( int w = 1; w < 4; ++w ) { lo = fminf( lo, s_mn[w][a] ); hi = fmaxf( hi, s_mx[w][a] ); }
    boxes[6 * (size_t)blockIdx.x + a] = lo; boxes[6 * (size_t)blockIdx.x + 3 + a] = hi;
  }
}

// cell flags of the byte grids
constexpr uint8_t OCC_BOUNDARY = 1, OCC_IN_X = 2, OCC_IN_Z = 4, OCC_TMP = 8;

__global__ void __launch_bounds__( 256 ) overlap_kernel( const float* __restrict__ pos1, int n1, const float* __restrict__ poses, int stride,
                                                         const PairDesc* __restrict__ pairs, float voxel, int inside, int normalize_by_smaller,
                                                         uint8_t* __restrict__ scratch, float* __restrict__ out )
{
  const PairDesc pd = pairs[blockIdx.x];
  const int xr = pd.xr, yr = pd.yr, zr = pd.zr;
  const size_t n_cells = (size_t)xr * yr * zr;
  uint8_t* grid[2] = { scratch + pd.off, scratch + pd.off + n_cells };
  const int tid = threadIdx.x, nt = blockDim.x;
  for( size_t c = tid; c < 2 * n_cells; c += nt ) { grid[0][c] = 0; }
  __syncthreads();
  // ---- 1. boundary cells (intersect.h:221-235): cell (x, y, z) lives at (y * zr + z) * xr + x (:108)
  for( int g = 0; g < 2; ++g )
  {
    const float* m = poses + (size_t)( g == 0 ? pd.a : pd.b ) * stride;
    for( int i = tid; i < n1; i += nt )
    {
      float px, py, pz;
      xf_apply( m, pos1[3 * (size_t)i], pos1[3 * (size_t)i + 1], pos1[3 * (size_t)i + 2], 1.0f, px, py, pz );
      const int x = (int)floorf( __fdiv_rn( __fsub_rn( px, pd.origin[0] ), voxel ) );
      const int y = (int)floorf( __fdiv_rn( __fsub_rn( py, pd.origin[1] ), voxel ) );
      const int z = (int)floorf( __fdiv_rn( __fsub_rn( pz, pd.origin[2] ), voxel ) );
      if( x >= 0 && x < xr && y >= 0 && y < yr && z >= 0 && z < zr ) { grid[g][( (size_t)y * zr + z ) * xr + x] = OCC_BOUNDARY; }
    }
  }
  __syncthreads();
  // ---- 2. scan-line fill, one thread per row; x-rows first, z-rows after the barrier (every cell then has one writer per phase)
  if( inside )
  {
    for( int dir = 0; dir < 2; ++dir )
    {
      const int n_rows = dir == 0 ? zr : xr, len = dir == 0 ? xr : zr;
      const uint8_t flag = dir == 0 ? OCC_IN_X : OCC_IN_Z;
      const size_t step = dir == 0 ? 1 : (size_t)xr;
      const int total = 2 * yr * n_rows;
      for( int job = tid; job < total; job += nt )
      {
        const int g = job / ( yr * n_rows ), rem = job % ( yr * n_rows ), y = rem / n_rows, r = rem % n_rows;
        uint8_t* row = grid[g] + (size_t)y * zr * xr + ( dir == 0 ? (size_t)r * xr : (size_t)r );
        int fill = 0, prev = 0;
        for( int t = 0; t < len; ++t )
        {
          uint8_t v = row[t * step];
          const int b = v & OCC_BOUNDARY;
          if( !b && prev ) { fill += 1; }
          row[t * step] = ( fill & 1 ) ? (uint8_t)( v | OCC_TMP ) : (uint8_t)( v & ~OCC_TMP );
          prev = b;
        }
        fill = 0; prev = 0;
        for( int t = len - 1; t >= 0; --t )
        {
          uint8_t v = row[t * step];
          const int b = v & OCC_BOUNDARY;
          if( !b && prev ) { fill += 1; }
          const bool in = ( v & OCC_TMP ) && ( fill & 1 ) && !b;
          v = (uint8_t)( v & ~OCC_TMP );
          row[t * step] = in ? (uint8_t)( v | flag ) : v;
          prev = b;
        }
      }
      __syncthreads();
    }
  }
  // ---- 3. counts (intersect.h:286-306, 289-305): occupied = boundary, or inside along both directions
  int ca = 0, cb = 0, both = 0;
  for( size_t c = tid; c < n_cells; c += nt )
  {
    const uint8_t va = grid[0][c], vb = grid[1][c];
    const bool oa = ( va & OCC_BOUNDARY ) || ( inside && ( va & OCC_IN_X ) && ( va & OCC_IN_Z ) );
    const bool ob = ( vb & OCC_BOUNDARY ) || ( inside && ( vb & OCC_IN_X ) && ( vb & OCC_IN_Z ) );
    ca += oa; cb += ob; both += oa && ob;
  }
  __shared__ int s_cnt[3];
  if( tid < 3 ) { s_cnt[tid] = 0; }
  __syncthreads();
  for( int o = 16; o > 0; o >>= 1 )
  {
    ca += __shfl_down_sync( RS_FULL, ca, o ); cb += __shfl_down_sync( RS_FULL, cb, o ); both += __shfl_down_sync( RS_FULL, both, o );
  }
  if( ( tid & 31 ) == 0 ) { atomicAdd( &s_cnt[0], ca ); atomicAdd( &s_cnt[1], cb ); atomicAdd( &s_cnt[2], both ); }
  __syncthreads();
  if( tid == 0 )
  {
    const int denom = normalize_by_smaller ? min( s_cnt[0], s_cnt[1] ) : max( s_cnt[0], s_cnt[1] );
    out[blockIdx.x] = denom > 0 ? __fdiv_rn( (float)s_cnt[2], (float)denom ) : 1.0f; // (:350-357)
  }
}

// msh_mat4_vec3_mul on the host (msh_vec_math.h:1554-1561); this TU is compiled with -ffp-contract=off
void host_xf_point( const float* m, const float* v, float* o )
{
  for( int r = 0; r < 3; ++r )
  {
    volatile float s = m[r] * v[0];
    volatile float t = m[4 + r] * v[1]; s = s + t;
    t = m[8 + r] * v[2]; s = s + t;
    t = 1.0f * m[12 + r]; s = s + t;
    o[r] = s;
  }
}

struct OverlapBatch
{
  const rsgpu_cloud_t* lvl3; const rsgpu_cloud_t* lvl1;
  DevBuf<float> d_poses; int stride = 16; int n_poses = 0;
  std::vector<float> boxes; // n_poses x 6

  int init( const rsgpu_cloud_t* l3, const rsgpu_cloud_t* l1, const float* poses, int n, int stride_floats )
  {
    lvl3 = l3; lvl1 = l1; n_poses = n; stride = stride_floats;
    cudaStream_t st = rt().stream;
    RS_CUDA( d_poses.alloc( (size_t)n * stride ) );
    RS_CUDA( cudaMemcpyAsync( d_poses.p, poses, sizeof( float ) * (size_t)n * stride, cudaMemcpyHostToDevice, st ) );
    DevBuf<float> d_boxes;
    RS_CUDA( d_boxes.alloc( (size_t)n * 6 ) );
    posed_bbox_kernel<<<(unsigned)n, 128, 0, st>>>( lvl3->pos.p, lvl3->n, d_poses.p, stride, d_boxes.p );
    RS_CHECK_LAUNCH();
    boxes.resize( (size_t)n * 6 );
    RS_CUDA( cudaMemcpyAsync( boxes.data(), d_boxes.p, sizeof( float ) * 6 * (size_t)n, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( rs::stream_sync( st ) );
    return RSGPU_OK;
  }

  // mshgeo_bbox_intersect (msh_geometry.h:1010-1015) of the posed level-3 boxes: without it the overlap is 0 (intersect.h:363-366)
  bool boxes_intersect( int a, int b ) const
  {
    const float* ba = &boxes[6 * (size_t)a]; const float* bb = &boxes[6 * (size_t)b];
    for( int k = 0; k < 3; ++k ) { if( !( ba[3 + k] >= bb[k] && bb[3 + k] >= ba[k] ) ) { return false; } }
    return true;
  }
  // bytes of voxel scratch the pair needs (0: boxes do not intersect)
  unsigned long long pair_scratch( int a, int b, float voxel ) const
  {
    if( !boxes_intersect( a, b ) ) { return 0; }
    const float* ba = &boxes[6 * (size_t)a]; const float* bb = &boxes[6 * (size_t)b];
    unsigned long long cells = 2;
    for( int k = 0; k < 3; ++k )
    {
      volatile float mn = std::min( ba[k], bb[k] ), mx = std::max( ba[3 + k], bb[3 + k] );
      mn = mn - 0.3f; mx = mx + 0.3f;
      volatile float w = mx - mn;
      volatile float q = w / voxel;
      const int r = (int)ceilf( q ) + 1;
      cells *= (unsigned long long)( r > 0 ? r : 0 );
    }
    return cells;
  }

  // overlap factors of pose `ref` against the poses listed in `others` -> out[j]
  int run( int ref, const std::vector<int>& others, float voxel, int inside, int normalize_by_smaller, std::vector<float>& out )
  {
    std::vector<int> as( others.size(), ref );
    return run_pairs( as, others, voxel, inside, normalize_by_smaller, out );
  }

  // overlap factors of the pairs (as[j], bs[j]) -> out[j], all in ONE launch (one block per pair whose boxes intersect)
  int run_pairs( const std::vector<int>& as, const std::vector<int>& bs, float voxel, int inside, int normalize_by_smaller, std::vector<float>& out )
  {
    const std::vector<int>& others = bs;
    out.assign( others.size(), 0.0f );
    std::vector<PairDesc> pairs; std::vector<int> slot;
    unsigned long long off = 0;
    for( size_t j = 0; j < others.size(); ++j )
    {
      const int ref = as[j];
      const float* ba = &boxes[6 * (size_t)ref];
      const float* bb = &boxes[6 * (size_t)others[j]];
      if( !boxes_intersect( ref, others[j] ) ) { continue; } // overlap = 0 (:363-366)
      PairDesc pd; pd.a = ref; pd.b = others[j];
      int res[3];
      for( int a = 0; a < 3; ++a )
      {
        volatile float mn = std::min( ba[a], bb[a] ), mx = std::max( ba[3 + a], bb[3 + a] );
        mn = mn - 0.3f; mx = mx + 0.3f; // fat_factor (intersect.h:61-65)
        volatile float w = mx - mn;
        volatile float q = w / voxel;
        res[a] = (int)ceilf( q ) + 1;  // (:67-69)
        pd.origin[a] = mn;
      }
      pd.xr = res[0]; pd.yr = res[1]; pd.zr = res[2];
      const unsigned long long n_cells = (unsigned long long)pd.xr * pd.yr * pd.zr;
      if( pd.xr <= 0 || pd.yr <= 0 || pd.zr <= 0 || n_cells > ( 1ull << 28 ) ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu overlap: pair grid too large" ); }
      pd.off = off; off += 2 * n_cells;
      pairs.push_back( pd ); slot.push_back( (int)j );
    }
    if( pairs.empty() ) { return RSGPU_OK; }
    cudaStream_t st = rt().stream;
    DevBuf<PairDesc> d_pairs; DevBuf<uint8_t> d_scratch; DevBuf<float> d_out;
    RS_CUDA( d_pairs.alloc( pairs.size() ) ); RS_CUDA( d_scratch.alloc( (size_t)off ) ); RS_CUDA( d_out.alloc( pairs.size() ) );
    RS_CUDA( cudaMemcpyAsync( d_pairs.p, pairs.data(), sizeof( PairDesc ) * pairs.size(), cudaMemcpyHostToDevice, st ) );
    {
      ProfScope prof( "overlap" );
      overlap_kernel<<<(unsigned)pairs.size(), 256, 0, st>>>( lvl1->pos.p, lvl1->n, d_poses.p, stride, d_pairs.p, voxel, inside, normalize_by_smaller,
                                                              d_scratch.p, d_out.p );
      RS_CHECK_LAUNCH();
    }
    std::vector<float> h( pairs.size() );
    RS_CUDA( cudaMemcpyAsync( h.data(), d_out.p, sizeof( float ) * pairs.size(), cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( rs::stream_sync( st ) );
    for( size_t p = 0; p < pairs.size(); ++p ) { out[slot[p]] = h[p]; }
    return RSGPU_OK;
  }
};
} // namespace

extern "C" {

int rsgpu_overlap_factors( const rsgpu_cloud_t* lvl3, const rsgpu_cloud_t* lvl1, const float* pose_ref, const float* poses, int32_t n,
                           float voxel_size, int32_t voxelize_inside, int32_t normalize_by_smaller, float* out )
{
  if( !lvl3 || !lvl1 || !pose_ref || n < 0 || ( n > 0 && ( !poses || !out ) ) || !( voxel_size > 0.f ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_overlap_factors: bad argument" ); }
  RS_TRY( ensure_device() );
  if( n == 0 ) { return RSGPU_OK; }
  std::vector<float> all( (size_t)( n + 1 ) * 16 );
  memcpy( all.data(), pose_ref, 64 );
  memcpy( all.data() + 16, poses, sizeof( float ) * 16 * (size_t)n );
  OverlapBatch ob;
  RS_TRY( ob.init( lvl3, lvl1, all.data(), n + 1, 16 ) );
  std::vector<int> others( n ); for( int i = 0; i < n; ++i ) { others[i] = i + 1; }
  std::vector<float> res;
  RS_TRY( ob.run( 0, others, voxel_size, voxelize_inside, normalize_by_smaller, res ) );
  memcpy( out, res.data(), sizeof( float ) * (size_t)n );
  return RSGPU_OK;
}

int rsgpu_nms( const rsgpu_cloud_t* lvl3, const rsgpu_cloud_t* lvl1, const float centroid[3], const float* proposals, int32_t n,
               float dist_threshold, uint8_t* keep )
{
  if( !lvl3 || !lvl1 || !centroid || n < 0 || ( n > 0 && ( !proposals || !keep ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_nms: bad argument" ); }
  RS_TRY( ensure_device() );
  if( n == 0 ) { return RSGPU_OK; }
  OverlapBatch ob;
  RS_TRY( ob.init( lvl3, lvl1, proposals, n, RSGPU_POSE_FLOATS ) );
  std::vector<uint8_t> mark( n, 0 ); // 0 unmarked, 1 keep, 2 discard (pose_proposal.cpp:384-389)
  std::vector<float> cpos( (size_t)n * 3 );
  for( int i = 0; i < n; ++i ) { host_xf_point( proposals + RSGPU_POSE_FLOATS * (size_t)i, centroid, &cpos[3 * (size_t)i] ); }
  auto centroid_dist = [&]( int a, int b ) {
    volatile float dx = cpos[3 * (size_t)a] - cpos[3 * (size_t)b], dy = cpos[3 * (size_t)a + 1] - cpos[3 * (size_t)b + 1],
                   dz = cpos[3 * (size_t)a + 2] - cpos[3 * (size_t)b + 2];
    volatile float s = dx * dx; volatile float t = dy * dy; s = s + t; t = dz * dz; s = s + t;
    return (float)sqrt( (double)s ); // msh_vec3_norm (msh_vec_math.h:988-991)
  };
  // The greedy loop asks for overlap( best, i ) only where i survives the distance and score tests against `best`
  // (pose_proposal.cpp:410-424).  Which pose becomes `best` in which round is not known up front, but every pair the loop can
  // ever ask for is among {i < j : both scores >= 0.01, centroids >= dist_threshold apart, posed boxes intersect}; the factor
  // is symmetric in the pair (union box, max of the two counts), so all of them go into ONE launch and the loop below
  // runs on the host without touching the device again: two host waits per call instead of one per kept pose.  Lists whose
  // pair set is too big for that (top_k = 0: thousands of survivors) take the round-by-round path.
  constexpr size_t NMS_MAX_PAIRS = 8192;
  constexpr unsigned long long NMS_MAX_SCRATCH = 1ull << 29;
  std::vector<int> pa, pb;
  bool all_pairs = option( "nms_impl" ) != "rounds" && (size_t)n * ( (size_t)n - 1 ) / 2 <= 4 * NMS_MAX_PAIRS;
  if( all_pairs )
  {
    unsigned long long scratch = 0;
    for( int i = 0; i < n && all_pairs; ++i )
    {
      if( proposals[RSGPU_POSE_FLOATS * (size_t)i + 16] < 0.01f ) { continue; }
      for( int j = i + 1; j < n; ++j )
      {
        if( proposals[RSGPU_POSE_FLOATS * (size_t)j + 16] < 0.01f || centroid_dist( i, j ) < dist_threshold ) { continue; }
        const unsigned long long need = ob.pair_scratch( i, j, 0.1f );
        if( !need ) { continue; }
        pa.push_back( i ); pb.push_back( j ); scratch += need;
        if( pa.size() > NMS_MAX_PAIRS || scratch > NMS_MAX_SCRATCH ) { all_pairs = false; break; }
      }
    }
  }
  std::vector<float> pair_ov;
  std::vector<int> pair_at; // [i * n + j], i < j -> index into pair_ov, -1 = not needed / boxes apart (overlap 0)
  if( all_pairs )
  {
    RS_TRY( ob.run_pairs( pa, pb, 0.1f, 1, 0, pair_ov ) );
    pair_at.assign( (size_t)n * n, -1 );
    for( size_t p = 0; p < pa.size(); ++p ) { pair_at[(size_t)pa[p] * n + pb[p]] = (int)p; }
  }
  int marked = 0;
  std::vector<int> others; std::vector<float> overlap;
  while( marked != n )
  {
    int best = -1; float best_score = -1e9f;
    for( int i = 0; i < n; ++i )
    {
      const float sc = proposals[RSGPU_POSE_FLOATS * (size_t)i + 16];
      if( mark[i] == 0 && sc > best_score ) { best_score = sc; best = i; } // strict: the first maximum wins (:394-401)
    }
    if( best < 0 ) { break; }
    mark[best] = 1; marked++;
    // candidates that are not already discarded by distance or score need the voxel overlap (:410-424: the three
    // conditions are OR-ed, so the overlap of the others cannot change the outcome)
    others.clear();
    for( int i = 0; i < n; ++i )
    {
      if( mark[i] != 0 ) { continue; }
      const float dist = centroid_dist( best, i );
      if( dist < dist_threshold || proposals[RSGPU_POSE_FLOATS * (size_t)i + 16] < 0.01f ) { mark[i] = 2; marked++; }
      else { others.push_back( i ); }
    }
    if( all_pairs )
    {
      overlap.assign( others.size(), 0.0f );
      for( size_t j = 0; j < others.size(); ++j )
      {
        const int lo = std::min( best, others[j] ), hi = std::max( best, others[j] );
        const int at = pair_at[(size_t)lo * n + hi];
        if( at >= 0 ) { overlap[j] = pair_ov[at]; }
      }
    }
    else { RS_TRY( ob.run( best, others, 0.1f, 1, 0, overlap ) ); }
    for( size_t j = 0; j < others.size(); ++j )
    {
      if( overlap[j] > 0.5f ) { mark[others[j]] = 2; marked++; }
    }
  }
  for( int i = 0; i < n; ++i ) { keep[i] = mark[i] == 1; }
  return RSGPU_OK;
}

} // extern "C"
