// Fused batched pose scoring and the propose pipeline built on it.
//
// Replaces mgs_compute_object_alignment_score (reference apps/pose_proposal/pose_proposal.cpp:93-158),
// the dense search loop of mgs__initial_pose_proposals (:170-254), mgs__pose_verification (:256-303) and the
// survivor copy of mgs_propose_poses (:325-369).
//
// One warp per candidate pose.  For every object point the warp transforms it in registers (same float
// evaluation order as msh_mat4_vec3_mul), asks the grid for the nearest normal-compatible scan point inside
// the k-nearest list (nearest_compatible, rsgpu_internal.cuh) and accumulates the reference's score term in
// fp64 in the reference's point order, so no k x N neighbour lists, no per-query sort and no scratch
// storage exist on the GPU.  The fp64 exp/acos of 32 consecutive points are evaluated lane-parallel.
#include "rsgpu_internal.cuh"
#include "nearest.cuh"
#include "nearest_group.cuh"
#include "warplist.cuh"
#include <cub/device/device_scan.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <mutex>
#include <string>

using namespace rs;

namespace
{
struct PoseSource
{
  const float* __restrict__ xforms;  // explicit: n x 16
  const float* __restrict__ rots;    // grid: n_rot x 16
  const float* __restrict__ trans;   // grid: n_trans x 3
  const float* __restrict__ gate;    // optional per-pose gate: poses with gate <= 0 are skipped (score untouched)
  int n_rot;
};

struct ScoreParams
{
  double radius;     // search radius = sigma (double of the float literal)
  float r2f;         // (float)(radius*radius)
  float dot_thr;     // smallest float dot with acos((double)dot) - max_angle < 1e-6
  double inv_two_sigma_sq_den; // 2.0*sigma*sigma
  int k;
};

template <bool GRID, bool COUNT>
__global__ void __launch_bounds__( 128 ) score_kernel( GridView g, const float* __restrict__ obj_pos, const float* __restrict__ obj_nor,
                                                       int n_obj, PoseSource ps, long long n_poses, int n_split, int chunk,
                                                       ScoreParams sp, double* __restrict__ partial,
                                                       unsigned long long* __restrict__ counts )
{
  const int lane = threadIdx.x & 31;
  long long warp = ( blockIdx.x * (long long)blockDim.x + threadIdx.x ) >> 5;
  if( warp >= n_poses * n_split ) { return; }
  long long pose = warp / n_split;
  int split = (int)( warp % n_split );
  if( ps.gate && !( __ldg( ps.gate + pose ) > 0.0f ) ) { return; }

  float m[16];
  if( GRID )
  {
    long long t = pose / ps.n_rot; int r = (int)( pose % ps.n_rot );
#pragma unroll
    for( int i = 0; i < 12; ++i ) { m[i] = __ldg( ps.rots + 16 * (size_t)r + i ); }
    m[12] = __ldg( ps.trans + 3 * t ); m[13] = __ldg( ps.trans + 3 * t + 1 ); m[14] = __ldg( ps.trans + 3 * t + 2 ); m[15] = 1.0f;
  }
  else
  {
#pragma unroll
    for( int i = 0; i < 16; ++i ) { m[i] = __ldg( ps.xforms + 16 * (size_t)pose + i ); }
  }

  const int i0 = split * chunk, i1 = min( n_obj, i0 + chunk );
  double sum = 0.0;
  for( int ib = i0; ib < i1; ib += 32 )
  {
    // 32 object points per batch, one per lane: transform in registers (pose_proposal.cpp:106-112) ...
    const int i = ib + lane;
    const bool valid = i < i1;
    LaneQuery q;
    if( valid )
    {
      xf_apply( m, __ldg( obj_pos + 3 * (size_t)i ), __ldg( obj_pos + 3 * (size_t)i + 1 ), __ldg( obj_pos + 3 * (size_t)i + 2 ), 1.0f, q.px, q.py, q.pz );
      xf_apply( m, __ldg( obj_nor + 3 * (size_t)i ), __ldg( obj_nor + 3 * (size_t)i + 1 ), __ldg( obj_nor + 3 * (size_t)i + 2 ), 0.0f, q.nx, q.ny, q.nz );
    }
    // ... search (:115-148) ...
    NearestHit h = nearest_compatible_batch<COUNT>( g, q, valid, sp.radius, sp.r2f, sp.dot_thr, sp.k, counts );
    // ... and the reference's per-point term (:149-152), lane-parallel in fp64, summed in point order
    double term = 0.0;
    if( h.found )
    {
      double angle = acos( (double)fmaxf( h.dot, 0.0f ) );
      double nc = exp( -( angle * angle ) / ( 2.0 * 0.5 * 0.5 ) );
      double dc = exp( -(double)h.d2 / sp.inv_two_sigma_sq_den );
      term = 0.05 * nc + ( 1.0 - 0.05 ) * dc;
    }
    unsigned mask = __ballot_sync( RS_FULL, h.found );
    while( mask )
    {
      int src = __ffs( mask ) - 1; mask &= mask - 1;
      sum += __shfl_sync( RS_FULL, term, src );
    }
  }
  if( lane == 0 ) { partial[pose * n_split + split] = sum; }
}

// ---------------------------------------------------------------------------------------------- group kernel
// The production scorer.  One warp per (pose, point chunk), three passes over its points:
//   A  lane-parallel: transform, cell window, 3x3x3 occupancy + block normal-cone test (rsg::stage1_test); the points
//      that can have a compatible neighbour at all are compacted, in point order, into a per-warp list in shared memory;
//   B  the listed points are searched 8 at a time by 4-lane groups (nearest_group.cuh), 32 per round;
//   C  per round the fp64 acos / exp terms of the 32 results are evaluated lane-parallel and summed in point order.
// Bound pruning (prune_cnt >= 0, propose only): every term is <= 1, so a pose whose number of still-possible
// contributions is below threshold * N can never be emitted by mgs__initial_pose_proposals (its score is reported
// as 0, which never wins the per-translation arg-max and never passes the threshold: pose_proposal.cpp:217-243).
constexpr int SC_LIST_CAP = 1024;

// -DRS_SCORE_STATS: work census of the group kernel (diagnostic builds only; printed by rsgpu_propose_poses)
#ifdef RS_SCORE_STATS
__device__ unsigned long long g_score_stats[12];
#define RS_STAT( i, v ) do { if( lane == 0 ) { atomicAdd( &g_score_stats[i], (unsigned long long)( v ) ); } } while( 0 )
#else
#define RS_STAT( i, v ) do { } while( 0 )
#endif

// SC_WARPS warps (poses) per block.  The poses of a block differ a lot in work (free space vs on a surface), but block
// size turned out not to matter: 4, 2 and 1 warps per block run the C2 dense launches in 28.8 / 28.5 / 28.6 ms per step
// (knob "score_warps", tests/test_gpu_variants.py)
template <bool GRID, int SC_G, int MINB, bool LANE, int SC_WARPS>
__global__ void __launch_bounds__( 32 * SC_WARPS, MINB ) score_kernel_g( GridView g, const float* __restrict__ obj_pos, const float* __restrict__ obj_nor,
                                                                   int n_obj, PoseSource ps, long long n_poses, int n_split, int chunk,
                                                                   ScoreParams sp, double prune_cnt, double* __restrict__ partial )
{
  __shared__ uint4 s_cand[SC_WARPS][rsg::GroupCfg<SC_G>::CAND_WORDS];
  __shared__ uint16_t s_list[SC_WARPS][SC_LIST_CAP];
  __shared__ float s_m[SC_WARPS][16];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  long long warp = ( blockIdx.x * (long long)blockDim.x + threadIdx.x ) >> 5;
  if( warp >= n_poses * n_split ) { return; }
  long long pose = warp / n_split;
  int split = (int)( warp % n_split );
  if( ps.gate && !( __ldg( ps.gate + pose ) > 0.0f ) ) { return; }

  // the pose lives in shared memory (16 registers less per thread; reads are warp broadcasts)
  float* m = s_m[wib];
  if( lane < 16 )
  {
    if( GRID )
    {
      long long t = pose / ps.n_rot; int r = (int)( pose % ps.n_rot );
      m[lane] = lane < 12 ? __ldg( ps.rots + 16 * (size_t)r + lane ) : ( lane < 15 ? __ldg( ps.trans + 3 * t + ( lane - 12 ) ) : 1.0f );
    }
    else { m[lane] = __ldg( ps.xforms + 16 * (size_t)pose + lane ); }
  }
  __syncwarp();
  uint16_t* list = s_list[wib];
  uint4* cand = s_cand[wib];
  const int i0 = split * chunk, i1 = min( n_obj, i0 + chunk );

  // ---- pass A: which points can contribute at all (pose_proposal.cpp:106-112 + the search's cell window)
  int n_list = 0;
  for( int ib = i0; ib < i1; ib += 32 )
  {
    const int i = ib + lane;
    const bool valid = i < i1;
    float px = 0.f, py = 0.f, pz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
    if( valid )
    {
      xf_apply( m, __ldg( obj_pos + 3 * (size_t)i ), __ldg( obj_pos + 3 * (size_t)i + 1 ), __ldg( obj_pos + 3 * (size_t)i + 2 ), 1.0f, px, py, pz );
      xf_apply( m, __ldg( obj_nor + 3 * (size_t)i ), __ldg( obj_nor + 3 * (size_t)i + 1 ), __ldg( obj_nor + 3 * (size_t)i + 2 ), 0.0f, nx, ny, nz );
    }
    const rsg::Stage1 s1 = rsg::stage1_test( g, sp.radius, sp.dot_thr, true, px, py, pz, nx, ny, nz, valid );
    const unsigned bal = __ballot_sync( RS_FULL, s1.active );
    if( s1.active ) { list[n_list + __popc( bal & ( ( 1u << lane ) - 1u ) )] = (uint16_t)( ( i - i0 ) | ( s1.fast ? 0 : 0x8000 ) ); }
    n_list += __popc( bal );
  }
  __syncwarp();

  double sum = 0.0;
  bool pruned = prune_cnt >= 0.0 && (double)n_list < prune_cnt;
  int n_found = 0;
  RS_STAT( 0, 1 ); RS_STAT( 1, i1 - i0 ); RS_STAT( 2, n_list ); RS_STAT( 3, pruned ? 1 : 0 ); RS_STAT( 4, pruned ? 0 : n_list );
  for( int base = 0; base < n_list && !pruned; base += 32 )
  {
    // the terms found so far are known, every remaining listed point contributes at most 1
    if( prune_cnt >= 0.0 && sum + (double)( n_list - base ) < prune_cnt ) { pruned = true; RS_STAT( 5, 1 ); RS_STAT( 6, n_list - base ); break; }
    const int n_round = min( 32, n_list - base );
    // ---- pass B: the searches of this round (:115-148)
    auto query_of = [&]( int r, float& px, float& py, float& pz, float& nx, float& ny, float& nz, unsigned long long& seedkey, float& seeddot ) -> bool {
      const int e = list[base + r];
      const int i = i0 + ( e & 0x7fff );
      xf_apply( m, __ldg( obj_pos + 3 * (size_t)i ), __ldg( obj_pos + 3 * (size_t)i + 1 ), __ldg( obj_pos + 3 * (size_t)i + 2 ), 1.0f, px, py, pz );
      xf_apply( m, __ldg( obj_nor + 3 * (size_t)i ), __ldg( obj_nor + 3 * (size_t)i + 1 ), __ldg( obj_nor + 3 * (size_t)i + 2 ), 0.0f, nx, ny, nz );
      return ( e & 0x8000 ) == 0;
    };
    NearestHit h;
    if( LANE )
    {
      float px = 0.f, py = 0.f, pz = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
      unsigned long long sk = ~0ull; float sd = 0.f;
      const bool qv = lane < n_round && query_of( lane, px, py, pz, nx, ny, nz, sk, sd );
      h = rsg::lane_search( g, qv, px, py, pz, nx, ny, nz, sp.radius, sp.r2f, sp.dot_thr, sp.k );
    }
    else { h = rsg::group_round<SC_G>( g, n_round, query_of, sp.radius, sp.r2f, sp.dot_thr, sp.k, cand ); }
    // windows the group path cannot take (last-bit cases, radius > cell size): generic warp-cooperative search
    const bool slow = lane < n_round && ( list[base + lane] & 0x8000 ) != 0;
    if( __any_sync( RS_FULL, slow ) )
    {
      LaneQuery q;
      unsigned long long sk2 = ~0ull; float sd2 = 0.f;
      if( slow ) { query_of( lane, q.px, q.py, q.pz, q.nx, q.ny, q.nz, sk2, sd2 ); }
      NearestHit hs = nearest_compatible_batch<false>( g, q, slow, sp.radius, sp.r2f, sp.dot_thr, sp.k, nullptr );
      if( slow ) { h = hs; }
    }
    // ---- pass C: the reference's per-point term (:149-152), lane-parallel in fp64, summed in point order
    double term = 0.0;
    if( h.found )
    {
      double angle = acos( (double)fmaxf( h.dot, 0.0f ) );
      double nc = exp( -( angle * angle ) / ( 2.0 * 0.5 * 0.5 ) );
      double dc = exp( -(double)h.d2 / sp.inv_two_sigma_sq_den );
      term = 0.05 * nc + ( 1.0 - 0.05 ) * dc;
    }
    unsigned mask = __ballot_sync( RS_FULL, h.found );
    n_found += __popc( mask );
    RS_STAT( 7, n_round ); RS_STAT( 8, __popc( mask ) );
    while( mask )
    {
      int src = __ffs( mask ) - 1; mask &= mask - 1;
      sum += __shfl_sync( RS_FULL, term, src );
    }
  }
  if( lane == 0 ) { partial[pose * n_split + split] = pruned ? 0.0 : sum; }
}

__global__ void finalize_kernel( const double* __restrict__ partial, long long n_poses, int n_split, int n_obj,
                                 const float* __restrict__ gate, float* __restrict__ scores )
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if( p >= n_poses ) { return; }
  if( gate && !( gate[p] > 0.0f ) ) { return; }
  double s = 0.0;
  for( int j = 0; j < n_split; ++j ) { s += partial[p * n_split + j]; }
  s /= (double)n_obj;  // (:156)
  scores[p] = (float)s;
}

// smallest float dot in [0, 1] the reference accepts: acos((double)dot) - max_angle < 1e-6 (:141-142), evaluated
// with the host libm the reference itself would use; acos is monotone, so the set of accepted dots is [thr, 1]
float compat_threshold( double max_angle )
{
  auto ok = [&]( float d ) { return acos( (double)d ) - max_angle < 0.000001; };
  if( ok( 0.0f ) ) { return -INFINITY; } // negative dots are clamped to 0 before the test (:139), so everything <= 1 passes
  if( !ok( 1.0f ) ) { return 2.0f; } // nothing is ever accepted
  uint32_t lo, hi; float f0 = 0.0f, f1 = 1.0f;
  memcpy( &lo, &f0, 4 ); memcpy( &hi, &f1, 4 ); // !ok(lo), ok(hi); positive floats order like their bits
  while( hi - lo > 1 )
  {
    uint32_t mid = lo + ( hi - lo ) / 2; float fm; memcpy( &fm, &mid, 4 );
    if( ok( fm ) ) { hi = mid; } else { lo = mid; }
  }
  float out; memcpy( &out, &hi, 4 );
  return out;
}

ScoreParams make_params( float radius, int k )
{
  ScoreParams sp;
  double max_angle = 35.0 * 0.005555555556 * 3.1415926535897932384626433832; // msh_deg2rad(35.0), msh_std.h:618,625
  double sigma = radius;
  sp.radius = radius;
  sp.r2f = (float)( sp.radius * sp.radius );
  sp.dot_thr = compat_threshold( max_angle );
  sp.inv_two_sigma_sq_den = 2.0 * sigma * sigma;
  sp.k = k;
  return sp;
}

} // namespace
#include "dense_binned.cuh"
namespace
{
// core launcher: scores (device) [n_poses]; gate may alias scores
// prune_thr > 0 (propose only): poses that provably cannot score above prune_thr are reported as 0
int score_launch( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const PoseSource& ps, bool grid_mode, long long n_poses,
                  int k, float radius, float* d_scores, unsigned long long* d_counts, float prune_thr = 0.0f )
{
  if( n_poses <= 0 ) { return RSGPU_OK; }
  if( !scene->has_normals ) { return fail( RSGPU_ERR_INVALID, "rsgpu score: the scene grid has no normals (rsgpu_grid_set_normals)" ); }
  if( obj->n <= 0 ) { return fail( RSGPU_ERR_INVALID, "rsgpu score: empty object cloud" ); }
  if( k <= 0 || !( radius > 0.f ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu score: max_n_neigh and radius must be positive" ); }
  cudaStream_t st = rt().stream;
  ScoreParams sp = make_params( radius, k );
  // few poses over many points (verification / rescoring): split the point range over several warps
  int n_split = 1;
  const long long want_warps = 148ll * 64;
  if( n_poses < want_warps )
  {
    long long s = ( want_warps + n_poses - 1 ) / n_poses;
    long long max_s = ( obj->n + 63 ) / 64;
    n_split = (int)( s < max_s ? s : max_s );
    if( n_split < 1 ) { n_split = 1; }
  }
  // RSGPU_SCORE_IMPL=coop selects the warp-per-query kernel (the census pass always uses it)
  const bool coop_env = option( "score_impl" ) == "coop";
  const bool group_impl = !d_counts && !coop_env;
  // the dense pose grid goes through the cell-binned, shared-memory-staged search (dense_binned.cuh) unless
  // "dense_impl" = "warp" asks for the first design (one warp per pose, score_kernel_g) or the input is outside its envelope
  if( grid_mode && group_impl && option( "dense_impl" ) != "warp" && dense_binned_supported( obj, scene, ps ) )
  {
    static thread_local DbScratch tl_scratch; // released with the thread (after the context is gone the free fails harmlessly)
    DbScratch pool_scratch;
    DbScratch& S = option( "dense_scratch" ) == "pool" ? pool_scratch : tl_scratch;
    DbPlan P;
    RS_TRY( dense_binned_alloc( S, P, obj, scene, ps, n_poses ) );
    const cudaStream_t lane_st2 = st;
    if( bulk_stream() != st )
    {
      cudaEvent_t fork = nullptr;
      RS_CUDA( cudaEventCreateWithFlags( &fork, cudaEventDisableTiming ) );
      cudaEventRecord( fork, lane_st2 );
      st = bulk_stream();
      cudaStreamWaitEvent( st, fork, 0 );
      cudaEventDestroy( fork );
    }
    // One dense search at a time: each one fills the GPU on its own, so launches of several lanes side by side only
    // interleave their blocks and ALL finish late (C2, 8 lanes: every search done at 12-17 ms instead of one every 1.6 ms),
    // and the latency-bound chains behind them (verification, NMS, ICP) then all start together and fight for launch
    // slots.  Taken in turn, the first object's chain runs beside the second object's search, and so on.
    // "dense_serial" = "0" restores the free-for-all (A/B).
    // The turn is taken ON THE DEVICE: the lock only orders the enqueueing, and every search waits for the event the previous
    // one recorded behind its last kernel - so consecutive searches follow one another without the host wake-up (~0.2-0.3 ms)
    // that handing the lock over after completion would put between them.
    static std::mutex dense_mu;
    static cudaEvent_t last_done = nullptr;            // recorded behind the most recently enqueued dense search
    static thread_local cudaEvent_t my_done = nullptr; // this lane's event (re-recorded per search)
    const bool serial = option( "dense_serial" ) != "0";
    std::unique_lock<std::mutex> dense_lock( dense_mu, std::defer_lock );
    if( serial )
    {
      if( !my_done ) { RS_CUDA( cudaEventCreateWithFlags( &my_done, cudaEventDisableTiming ) ); }
      dense_lock.lock();
      if( last_done && last_done != my_done ) { cudaStreamWaitEvent( st, last_done, 0 ); }
    }
    int status;
    {
      ProfScope prof( "score_dense", st );
      status = dense_binned_run( S, P, obj, scene, ps, sp, (double)prune_thr, d_scores, st );
    }
    if( serial )
    {
      cudaEventRecord( my_done, st );
      last_done = my_done;
      dense_lock.unlock();
    }
    if( st != lane_st2 )
    {
      cudaEvent_t join = nullptr;
      RS_CUDA( cudaEventCreateWithFlags( &join, cudaEventDisableTiming ) );
      cudaEventRecord( join, st );
      cudaStreamWaitEvent( lane_st2, join, 0 );
      cudaEventDestroy( join );
    }
    RS_CUDA( rs::stream_sync( lane_st2, true ) ); // the scratch is free for this thread's next launch
    return status;
  }
  if( group_impl && n_split < ( obj->n + SC_LIST_CAP - 1 ) / SC_LIST_CAP ) { n_split = ( obj->n + SC_LIST_CAP - 1 ) / SC_LIST_CAP; }
  int chunk = ( ( obj->n + n_split - 1 ) / n_split + 31 ) / 32 * 32;
  n_split = ( obj->n + chunk - 1 ) / chunk;
  const double prune_cnt = ( prune_thr > 0.0f && n_split == 1 ) ? (double)prune_thr * (double)obj->n / 1.000001 : -1.0;
  DevBuf<double> partial;
  RS_CUDA( partial.alloc( (size_t)n_poses * n_split ) );
  long long warps = n_poses * n_split;
  long long blocks = ( warps + 3 ) / 4;
  if( blocks > 2147483647ll ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu score: too many poses for one launch" ); }
  GridView g = scene->view();
  // the dense search is the throughput launch of the path: inside a lane it runs on the lane's low-priority stream
  // (runtime.cu), so the short launches of other objects' chains are dispatched ahead of its pending blocks
  const cudaStream_t lane_st = st;
  if( grid_mode && bulk_stream() != st )
  {
    cudaEvent_t fork = nullptr;
    RS_CUDA( cudaEventCreateWithFlags( &fork, cudaEventDisableTiming ) );
    cudaEventRecord( fork, lane_st );
    st = bulk_stream();
    cudaStreamWaitEvent( st, fork, 0 );
    cudaEventDestroy( fork );
  }
  {
    ProfScope prof( grid_mode ? "score_dense" : "score", st );
    if( group_impl )
    {
      // lanes per query, warps per block and resident blocks per SM (register cap) of the group kernel; the defaults
      // are the measured best
      const std::string og = option( "score_g" ), ob = option( "score_minb" ), ow = option( "score_warps" );
      const int cfg_g = og.empty() ? 4 : atoi( og.c_str() ), cfg_w = ow.empty() ? 4 : atoi( ow.c_str() );
      const int cfg_b = ob.empty() ? ( cfg_w == 1 ? 24 : ( cfg_w == 2 ? 12 : 6 ) ) : atoi( ob.c_str() );
      const bool lane_env = option( "search" ) == "lane";
      blocks = ( warps + cfg_w - 1 ) / cfg_w;
      if( cfg_w != 1 && cfg_w != 2 ) { blocks = ( warps + 3 ) / 4; }
      if( blocks > 2147483647ll ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu score: too many poses for one launch" ); }
#define RS_SCORE_G_LAUNCH( GRIDM, GG, MB, LN, WW ) \
      score_kernel_g<GRIDM, GG, MB, LN, WW><<<(unsigned)blocks, 32 * WW, 0, st>>>( g, obj->pos.p, obj->nor.p, obj->n, ps, n_poses, n_split, chunk, sp, prune_cnt, partial.p )
#define RS_SCORE_G_PICK( GRIDM ) \
      do { \
        if( lane_env ) { RS_SCORE_G_LAUNCH( GRIDM, 4, 6, true, 4 ); } \
        else if( cfg_w == 1 ) { if( cfg_g == 8 ) { RS_SCORE_G_LAUNCH( GRIDM, 8, 24, false, 1 ); } else if( cfg_b >= 24 ) { RS_SCORE_G_LAUNCH( GRIDM, 4, 24, false, 1 ); } else { RS_SCORE_G_LAUNCH( GRIDM, 4, 16, false, 1 ); } } \
        else if( cfg_w == 2 ) { if( cfg_g == 8 ) { RS_SCORE_G_LAUNCH( GRIDM, 8, 12, false, 2 ); } else { RS_SCORE_G_LAUNCH( GRIDM, 4, 12, false, 2 ); } } \
        else if( cfg_g == 8 ) { if( cfg_b >= 8 ) { RS_SCORE_G_LAUNCH( GRIDM, 8, 8, false, 4 ); } else if( cfg_b >= 6 ) { RS_SCORE_G_LAUNCH( GRIDM, 8, 6, false, 4 ); } else { RS_SCORE_G_LAUNCH( GRIDM, 8, 4, false, 4 ); } } \
        else { if( cfg_b >= 8 ) { RS_SCORE_G_LAUNCH( GRIDM, 4, 8, false, 4 ); } else if( cfg_b >= 6 ) { RS_SCORE_G_LAUNCH( GRIDM, 4, 6, false, 4 ); } else { RS_SCORE_G_LAUNCH( GRIDM, 4, 4, false, 4 ); } } \
      } while( 0 )
      if( grid_mode ) { RS_SCORE_G_PICK( true ); } else { RS_SCORE_G_PICK( false ); }
#undef RS_SCORE_G_PICK
#undef RS_SCORE_G_LAUNCH
    }
    else if( d_counts )
    {
      if( grid_mode ) { score_kernel<true, true><<<(unsigned)blocks, 128, 0, st>>>( g, obj->pos.p, obj->nor.p, obj->n, ps, n_poses, n_split, chunk, sp, partial.p, d_counts ); }
      else { score_kernel<false, true><<<(unsigned)blocks, 128, 0, st>>>( g, obj->pos.p, obj->nor.p, obj->n, ps, n_poses, n_split, chunk, sp, partial.p, d_counts ); }
    }
    else
    {
      if( grid_mode ) { score_kernel<true, false><<<(unsigned)blocks, 128, 0, st>>>( g, obj->pos.p, obj->nor.p, obj->n, ps, n_poses, n_split, chunk, sp, partial.p, nullptr ); }
      else { score_kernel<false, false><<<(unsigned)blocks, 128, 0, st>>>( g, obj->pos.p, obj->nor.p, obj->n, ps, n_poses, n_split, chunk, sp, partial.p, nullptr ); }
    }
    RS_CHECK_LAUNCH();
  }
  if( st != lane_st )
  {
    cudaEvent_t join = nullptr;
    RS_CUDA( cudaEventCreateWithFlags( &join, cudaEventDisableTiming ) );
    cudaEventRecord( join, st );
    st = lane_st;
    cudaStreamWaitEvent( st, join, 0 );
    cudaEventDestroy( join );
  }
  finalize_kernel<<<(unsigned)( ( n_poses + 255 ) / 256 ), 256, 0, st>>>( partial.p, n_poses, n_split, obj->n, ps.gate, d_scores );
  RS_CHECK_LAUNCH();
  RS_CUDA( rs::stream_sync( st, grid_mode ) ); // partial dies here; the dense launch runs for milliseconds: sleep, do not spin
  return RSGPU_OK;
}

// ---------------------------------------------------------------------------------------------- propose
// per translation: first strict maximum over rotations starting from 0, flagged iff > threshold (:217-243)
__global__ void select_kernel( const float* __restrict__ scores, long long n_trans, int n_rot, float thr,
                               int* __restrict__ flag, int* __restrict__ best_r, float* __restrict__ best_s )
{
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if( t >= n_trans ) { return; }
  float best = 0.f; int br = -1;
  for( int r = 0; r < n_rot; ++r )
  {
    float s = scores[t * n_rot + r];
    if( s > best ) { best = s; br = r; }
  }
  // br < 0: no rotation scored above 0; with a negative caller threshold the reference would emit a zero matrix with
  // score 0 here, which its final |score| > 1e-6 copy drops again (:348-359) - never flagged
  flag[t] = ( br >= 0 && best > thr ) ? 1 : 0;
  best_r[t] = br; best_s[t] = best;
}

// compact the flagged translations, in translation order, into pose_proposal_t records (16 + 1 floats)
__global__ void emit_kernel( const int* __restrict__ flag, const int* __restrict__ offs, const int* __restrict__ best_r,
                             const float* __restrict__ best_s, const float* __restrict__ rots, const float* __restrict__ trans,
                             long long n_trans, int n_rot, float* __restrict__ xforms, float* __restrict__ scores,
                             long long* __restrict__ pose_id, const long long* __restrict__ trans_ids )
{
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if( t >= n_trans || !flag[t] ) { return; }
  int o = offs[t], r = best_r[t];
  for( int i = 0; i < 12; ++i ) { xforms[16 * (size_t)o + i] = rots[16 * (size_t)r + i]; }
  xforms[16 * (size_t)o + 12] = trans[3 * t]; xforms[16 * (size_t)o + 13] = trans[3 * t + 1];
  xforms[16 * (size_t)o + 14] = trans[3 * t + 2]; xforms[16 * (size_t)o + 15] = 1.0f;
  scores[o] = best_s[t];
  pose_id[o] = ( trans_ids ? trans_ids[t] : t ) * n_rot + r; // the caller's numbering of the translations
}

// proposals re-ordered by ascending pose id (translations passed in another order than the caller numbers them)
__global__ void reorder_kernel( const int* __restrict__ src, int n, const float* __restrict__ xf_in, const float* __restrict__ sc_in,
                                float* __restrict__ xf_out, float* __restrict__ sc_out )
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if( j >= n ) { return; }
  const int i = src[j];
  for( int c = 0; c < 16; ++c ) { xf_out[16 * (size_t)j + c] = xf_in[16 * (size_t)i + c]; }
  sc_out[j] = sc_in[i];
}
__global__ void iota_kernel( int* __restrict__ v, int n )
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if( j < n ) { v[j] = j; }
}

// verification outcome (:289-292): new score if above the level's threshold, else -1; gate <= 0 untouched
__global__ void verify_kernel( float* __restrict__ scores, const float* __restrict__ fresh, int n, float thr )
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if( i >= n ) { return; }
  if( !( scores[i] > 0.0f ) ) { return; }
  scores[i] = fresh[i] > thr ? fresh[i] : -1.0f;
}
// Warp-cooperative top-k of one object's verified proposals: descending score, ties by ascending pose id (= emission
// order).  One warp streams the scores through a sorted list of k keys held in its registers (warplist.cuh); only
// survivors of the reference's final copy, |score| > 1e-6 (:352), enter.  sel[j] = index of the j-th best.
template <int EPL>
__global__ void __launch_bounds__( 32 ) topk_kernel( const float* __restrict__ scores, int n, int k, int* __restrict__ sel, int* __restrict__ n_sel )
{
  const int lane = threadIdx.x;
  WarpList<EPL> list; list.init();
  unsigned long long thr = KEY_INF;
  int seen = 0;
  for( int base = 0; base < n; base += 32 )
  {
    const int i = base + lane;
    unsigned long long key = KEY_INF;
    bool ok = false;
    if( i < n )
    {
      const float sc = scores[i];
      ok = fabsf( sc ) > 0.000001f;
      uint32_t u = __float_as_uint( sc );
      u = ( u & 0x80000000u ) ? ~u : ( u | 0x80000000u ); // ascending order of floats as unsigned
      key = ( (unsigned long long)( ~u ) << 32 ) | (uint32_t)i; // descending score, then ascending index
    }
    seen += __popc( __ballot_sync( RS_FULL, ok ) );
    unsigned m = __ballot_sync( RS_FULL, ok && key < thr );
    while( m )
    {
      const int src = __ffs( m ) - 1; m &= m - 1;
      const unsigned long long x = __shfl_sync( RS_FULL, key, src );
      if( x < thr ) { list.insert( x, lane ); thr = list.get( k - 1 ); }
    }
  }
  const int count = seen < k ? seen : k;
#pragma unroll
  for( int s = 0; s < EPL; ++s )
  {
    const int j = lane * EPL + s;
    if( j < count ) { sel[j] = (int)(uint32_t)( list.v[s] & 0xffffffffull ); }
  }
  if( lane == 0 ) { *n_sel = count; }
}

// all survivors in emission order (top_k = 0): sel = indices with |score| > 1e-6, ascending
__global__ void survivors_kernel( const float* __restrict__ scores, int n, int* __restrict__ sel, int* __restrict__ n_sel )
{
  // one warp, ordered compaction by ballot
  const int lane = threadIdx.x;
  int w = 0;
  for( int base = 0; base < n; base += 32 )
  {
    const int i = base + lane;
    const bool ok = i < n && fabsf( scores[i] ) > 0.000001f;
    const unsigned b = __ballot_sync( RS_FULL, ok );
    if( ok ) { sel[w + __popc( b & ( ( 1u << lane ) - 1u ) )] = i; }
    w += __popc( b );
  }
  if( lane == 0 ) { *n_sel = w; }
}

// out[j] = {xform, score} of proposal sel[j], ids[j] = its dense pose id
__global__ void gather_kernel( const int* __restrict__ sel, const int* __restrict__ n_sel, const float* __restrict__ xforms,
                               const float* __restrict__ scores, const long long* __restrict__ pose_id, float* __restrict__ out,
                               long long* __restrict__ ids )
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if( j >= *n_sel ) { return; }
  const int i = sel[j];
  for( int c = 0; c < 16; ++c ) { out[RSGPU_POSE_FLOATS * (size_t)j + c] = xforms[16 * (size_t)i + c]; }
  out[RSGPU_POSE_FLOATS * (size_t)j + 16] = scores[i];
  ids[j] = pose_id[i];
}
} // namespace

extern "C" {

int rsgpu_score_poses_dev( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const float* d_xforms, int64_t n_poses,
                           int32_t k, float radius, float* d_scores )
{
  if( !obj || !scene || ( n_poses > 0 && ( !d_xforms || !d_scores ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_score_poses: NULL argument" ); }
  RS_TRY( ensure_device() );
  PoseSource ps; memset( &ps, 0, sizeof( ps ) ); ps.xforms = d_xforms;
  return score_launch( obj, scene, ps, false, n_poses, k, radius, d_scores, nullptr );
}

int rsgpu_score_poses( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const float* xforms, int64_t n_poses, int32_t k,
                       float radius, float* scores )
{
  if( !obj || !scene || ( n_poses > 0 && ( !xforms || !scores ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_score_poses: NULL argument" ); }
  RS_TRY( ensure_device() );
  if( n_poses <= 0 ) { return RSGPU_OK; }
  DevBuf<float> dx, ds;
  RS_CUDA( dx.alloc( (size_t)n_poses * 16 ) ); RS_CUDA( ds.alloc( (size_t)n_poses ) );
  RS_CUDA( cudaMemcpyAsync( dx.p, xforms, sizeof( float ) * 16 * (size_t)n_poses, cudaMemcpyHostToDevice, rt().stream ) );
  RS_TRY( rsgpu_score_poses_dev( obj, scene, dx.p, n_poses, k, radius, ds.p ) );
  RS_CUDA( cudaMemcpyAsync( scores, ds.p, sizeof( float ) * (size_t)n_poses, cudaMemcpyDeviceToHost, rt().stream ) );
  RS_CUDA( rs::stream_sync( rt().stream ) );
  return RSGPU_OK;
}

int rsgpu_score_pose_grid_dev( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const float* d_rots, int32_t n_rot,
                               const float* d_trans, int64_t n_trans, int32_t k, float radius, float* d_scores )
{
  if( !obj || !scene || n_rot < 0 || n_trans < 0 ) { return fail( RSGPU_ERR_INVALID, "rsgpu_score_pose_grid: bad argument" ); }
  if( (long long)n_rot * n_trans > 0 && ( !d_rots || !d_trans || !d_scores ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_score_pose_grid: NULL argument" ); }
  RS_TRY( ensure_device() );
  PoseSource ps; memset( &ps, 0, sizeof( ps ) ); ps.rots = d_rots; ps.trans = d_trans; ps.n_rot = n_rot;
  return score_launch( obj, scene, ps, true, (long long)n_rot * n_trans, k, radius, d_scores, nullptr );
}

static int upload_pose_grid( const float* rots, int32_t n_rot, const float* trans, int64_t n_trans, DevBuf<float>& dr, DevBuf<float>& dt )
{
  RS_CUDA( dr.alloc( (size_t)n_rot * 16 ) ); RS_CUDA( dt.alloc( (size_t)n_trans * 3 ) );
  if( n_rot ) { RS_CUDA( cudaMemcpyAsync( dr.p, rots, sizeof( float ) * 16 * (size_t)n_rot, cudaMemcpyHostToDevice, rt().stream ) ); }
  if( n_trans ) { RS_CUDA( cudaMemcpyAsync( dt.p, trans, sizeof( float ) * 3 * (size_t)n_trans, cudaMemcpyHostToDevice, rt().stream ) ); }
  return RSGPU_OK;
}

int rsgpu_score_pose_grid( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const float* rots, int32_t n_rot, const float* trans,
                           int64_t n_trans, int32_t k, float radius, float* scores )
{
  if( !obj || !scene || n_rot < 0 || n_trans < 0 ) { return fail( RSGPU_ERR_INVALID, "rsgpu_score_pose_grid: bad argument" ); }
  long long n = (long long)n_rot * n_trans;
  if( n > 0 && ( !rots || !trans || !scores ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_score_pose_grid: NULL argument" ); }
  RS_TRY( ensure_device() );
  if( n == 0 ) { return RSGPU_OK; }
  DevBuf<float> dr, dt, ds;
  RS_TRY( upload_pose_grid( rots, n_rot, trans, n_trans, dr, dt ) );
  RS_CUDA( ds.alloc( (size_t)n ) );
  RS_TRY( rsgpu_score_pose_grid_dev( obj, scene, dr.p, n_rot, dt.p, n_trans, k, radius, ds.p ) );
  RS_CUDA( cudaMemcpyAsync( scores, ds.p, sizeof( float ) * (size_t)n, cudaMemcpyDeviceToHost, rt().stream ) );
  RS_CUDA( rs::stream_sync( rt().stream ) );
  return RSGPU_OK;
}

int rsgpu_score_pose_grid_count( const rsgpu_cloud_t* obj, const rsgpu_grid_t* scene, const float* rots, int32_t n_rot,
                                 const float* trans, int64_t n_trans, int32_t k, float radius, int64_t counts[4] )
{
  if( !obj || !scene || !counts || n_rot < 0 || n_trans < 0 ) { return fail( RSGPU_ERR_INVALID, "rsgpu_score_pose_grid_count: bad argument" ); }
  long long n = (long long)n_rot * n_trans;
  RS_TRY( ensure_device() );
  counts[0] = counts[1] = counts[2] = counts[3] = 0;
  if( n == 0 ) { return RSGPU_OK; }
  DevBuf<float> dr, dt, ds; DevBuf<unsigned long long> dc;
  RS_TRY( upload_pose_grid( rots, n_rot, trans, n_trans, dr, dt ) );
  RS_CUDA( ds.alloc( (size_t)n ) ); RS_CUDA( dc.alloc( 4 ) );
  RS_CUDA( cudaMemsetAsync( dc.p, 0, 32, rt().stream ) );
  PoseSource ps; memset( &ps, 0, sizeof( ps ) ); ps.rots = dr.p; ps.trans = dt.p; ps.n_rot = n_rot;
  bool prof = rt().profile; rt().profile = false; // the counting pass is not a timed launch
  int s = score_launch( obj, scene, ps, true, n, k, radius, ds.p, dc.p );
  rt().profile = prof;
  RS_TRY( s );
  unsigned long long h[4];
  RS_CUDA( cudaMemcpyAsync( h, dc.p, 32, cudaMemcpyDeviceToHost, rt().stream ) );
  RS_CUDA( rs::stream_sync( rt().stream ) );
  for( int i = 0; i < 4; ++i ) { counts[i] = (int64_t)h[i]; }
  return RSGPU_OK;
}

void rsgpu_propose_default_opts( rsgpu_propose_opts_t* o )
{
  if( !o ) { return; }
  o->max_n_neigh = 64; o->radius = 0.1f;
  o->thresholds[0] = 0.25f; o->thresholds[1] = 0.35f; o->thresholds[2] = 0.40f;
  o->top_k = 0;
  o->translation_ids = nullptr;
}

int rsgpu_propose_poses( const rsgpu_cloud_t* o4, const rsgpu_cloud_t* o3, const rsgpu_cloud_t* o2, const rsgpu_grid_t* scene,
                         const float* rots, int32_t n_rot, const float* trans, int64_t n_trans, const rsgpu_propose_opts_t* opts_in,
                         float* out, int64_t* out_pose_id, int64_t out_cap, int64_t* n_out )
{
  if( !o4 || !o3 || !o2 || !scene || !n_out || n_rot < 0 || n_trans < 0 || out_cap < 0 || ( out_cap > 0 && !out ) )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_propose_poses: bad argument" );
  }
  RS_TRY( ensure_device() );
  *n_out = 0;
  rsgpu_propose_opts_t opts;
  if( opts_in ) { opts = *opts_in; } else { rsgpu_propose_default_opts( &opts ); }
  long long n = (long long)n_rot * n_trans;
  if( opts.top_k > 512 ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_propose_poses: top_k > 512" ); }
  if( n == 0 ) { return RSGPU_OK; }
  if( n_trans > 2147483647ll ) { return fail( RSGPU_ERR_UNSUPPORTED, "rsgpu_propose_poses: more than 2^31 translations" ); }
  cudaStream_t st = rt().stream;
  DevBuf<float> dr, dt, ds;
  RS_TRY( upload_pose_grid( rots, n_rot, trans, n_trans, dr, dt ) );
  RS_CUDA( ds.alloc( (size_t)n ) );
  // level 4: dense search.  Scores that provably cannot exceed the level's threshold are not resolved (reported as
  // 0): they can neither be emitted nor change which rotation is emitted (RSGPU_PRUNE=0 resolves every score).
  {
    const bool no_prune = option( "prune" ) == "0";
    PoseSource ps; memset( &ps, 0, sizeof( ps ) ); ps.rots = dr.p; ps.trans = dt.p; ps.n_rot = n_rot;
    RS_TRY( score_launch( o4, scene, ps, true, n, opts.max_n_neigh, opts.radius, ds.p, nullptr, no_prune ? 0.0f : opts.thresholds[0] ) );
  }
  DevBuf<int> flag, offs, best_r; DevBuf<float> best_s;
  RS_CUDA( flag.alloc( n_trans ) ); RS_CUDA( offs.alloc( n_trans ) ); RS_CUDA( best_r.alloc( n_trans ) ); RS_CUDA( best_s.alloc( n_trans ) );
  unsigned tb = (unsigned)( ( n_trans + 255 ) / 256 );
  select_kernel<<<tb, 256, 0, st>>>( ds.p, n_trans, n_rot, opts.thresholds[0], flag.p, best_r.p, best_s.p );
  RS_CHECK_LAUNCH();
  size_t scan_bytes = 0;
  RS_CUDA( cub::DeviceScan::ExclusiveSum( nullptr, scan_bytes, flag.p, offs.p, (int)n_trans, st ) );
  DevBuf<unsigned char> tmp;
  RS_CUDA( tmp.alloc( scan_bytes ) );
  RS_CUDA( cub::DeviceScan::ExclusiveSum( tmp.p, scan_bytes, flag.p, offs.p, (int)n_trans, st ) );
  int last_off = 0, last_flag = 0;
  RS_CUDA( cudaMemcpyAsync( &last_off, offs.p + ( n_trans - 1 ), 4, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( cudaMemcpyAsync( &last_flag, flag.p + ( n_trans - 1 ), 4, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  int n_prop = last_off + last_flag;
  if( n_prop == 0 ) { return RSGPU_OK; }
  DevBuf<float> px, psc, fresh; DevBuf<long long> pid;
  RS_CUDA( px.alloc( (size_t)n_prop * 16 ) ); RS_CUDA( psc.alloc( n_prop ) ); RS_CUDA( fresh.alloc( n_prop ) ); RS_CUDA( pid.alloc( n_prop ) );
  DevBuf<long long> d_tids;
  if( opts.translation_ids )
  {
    RS_CUDA( d_tids.alloc( n_trans ) );
    RS_CUDA( cudaMemcpyAsync( d_tids.p, opts.translation_ids, sizeof( long long ) * (size_t)n_trans, cudaMemcpyHostToDevice, st ) );
  }
  emit_kernel<<<tb, 256, 0, st>>>( flag.p, offs.p, best_r.p, best_s.p, dr.p, dt.p, n_trans, n_rot, px.p, psc.p, pid.p, opts.translation_ids ? d_tids.p : nullptr );
  RS_CHECK_LAUNCH();
  DevBuf<float> px2, psc2; DevBuf<long long> pid2; DevBuf<int> src0, src1; DevBuf<unsigned char> sort_tmp;
  if( opts.translation_ids && n_prop > 1 )
  {
    // survivors in the caller's translation order (= ascending pose id: one rotation per translation), so that the
    // verification, the emission order and every tie below are what they would be for the caller's own order
    RS_CUDA( px2.alloc( (size_t)n_prop * 16 ) ); RS_CUDA( psc2.alloc( n_prop ) ); RS_CUDA( pid2.alloc( n_prop ) );
    RS_CUDA( src0.alloc( n_prop ) ); RS_CUDA( src1.alloc( n_prop ) );
    iota_kernel<<<( n_prop + 255 ) / 256, 256, 0, st>>>( src0.p, n_prop );
    RS_CHECK_LAUNCH();
    size_t sort_bytes = 0;
    RS_CUDA( cub::DeviceRadixSort::SortPairs( nullptr, sort_bytes, (const unsigned long long*)pid.p, (unsigned long long*)pid2.p, src0.p, src1.p, n_prop, 0, 64, st ) );
    RS_CUDA( sort_tmp.alloc( sort_bytes ) );
    RS_CUDA( cub::DeviceRadixSort::SortPairs( sort_tmp.p, sort_bytes, (const unsigned long long*)pid.p, (unsigned long long*)pid2.p, src0.p, src1.p, n_prop, 0, 64, st ) );
    reorder_kernel<<<( n_prop + 255 ) / 256, 256, 0, st>>>( src1.p, n_prop, px.p, psc.p, px2.p, psc2.p );
    RS_CHECK_LAUNCH();
    std::swap( px.p, px2.p ); std::swap( psc.p, psc2.p ); std::swap( pid.p, pid2.p );
  }
  // levels 3 and 2: verification of the survivors
  const rsgpu_cloud_t* lv[2] = { o3, o2 };
  for( int l = 0; l < 2; ++l )
  {
    PoseSource ps; memset( &ps, 0, sizeof( ps ) ); ps.xforms = px.p; ps.gate = psc.p;
    RS_TRY( score_launch( lv[l], scene, ps, false, n_prop, opts.max_n_neigh, opts.radius, fresh.p, nullptr ) );
    verify_kernel<<<( n_prop + 255 ) / 256, 256, 0, st>>>( psc.p, fresh.p, n_prop, opts.thresholds[1 + l] );
    RS_CHECK_LAUNCH();
  }
  // survivor copy: |score| > 1e-6 (:352) — every entry is either > threshold or -1, so all survive — and the
  // per-object top-k (descending score, ties by pose id), selected on the device; only the selected rows travel
  DevBuf<int> sel, nsel; DevBuf<float> dout; DevBuf<long long> dids;
  const int cap_sel = opts.top_k > 0 ? std::min( opts.top_k, n_prop ) : n_prop;
  RS_CUDA( sel.alloc( n_prop ) ); RS_CUDA( nsel.alloc( 1 ) ); RS_CUDA( dout.alloc( (size_t)cap_sel * RSGPU_POSE_FLOATS ) ); RS_CUDA( dids.alloc( cap_sel ) );
  if( opts.top_k > 0 )
  {
    const int k = opts.top_k;
    if( k <= 32 ) { topk_kernel<1><<<1, 32, 0, st>>>( psc.p, n_prop, k, sel.p, nsel.p ); }
    else if( k <= 64 ) { topk_kernel<2><<<1, 32, 0, st>>>( psc.p, n_prop, k, sel.p, nsel.p ); }
    else if( k <= 128 ) { topk_kernel<4><<<1, 32, 0, st>>>( psc.p, n_prop, k, sel.p, nsel.p ); }
    else if( k <= 256 ) { topk_kernel<8><<<1, 32, 0, st>>>( psc.p, n_prop, k, sel.p, nsel.p ); }
    else { topk_kernel<16><<<1, 32, 0, st>>>( psc.p, n_prop, k, sel.p, nsel.p ); }
  }
  else { survivors_kernel<<<1, 32, 0, st>>>( psc.p, n_prop, sel.p, nsel.p ); }
  RS_CHECK_LAUNCH();
  gather_kernel<<<( cap_sel + 127 ) / 128, 128, 0, st>>>( sel.p, nsel.p, px.p, psc.p, pid.p, dout.p, dids.p );
  RS_CHECK_LAUNCH();
  int n_sel = 0;
  RS_CUDA( cudaMemcpyAsync( &n_sel, nsel.p, sizeof( int ), cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  const int64_t n_copy = std::min<int64_t>( n_sel, out_cap );
  if( n_copy > 0 )
  {
    RS_CUDA( cudaMemcpyAsync( out, dout.p, sizeof( float ) * RSGPU_POSE_FLOATS * (size_t)n_copy, cudaMemcpyDeviceToHost, st ) );
    if( out_pose_id )
    {
      static_assert( sizeof( long long ) == sizeof( int64_t ), "pose ids are 64-bit" );
      RS_CUDA( cudaMemcpyAsync( out_pose_id, dids.p, sizeof( int64_t ) * (size_t)n_copy, cudaMemcpyDeviceToHost, st ) );
    }
    RS_CUDA( rs::stream_sync( st ) );
  }
  *n_out = (int64_t)n_sel; // may exceed out_cap: the caller then knows how much room a retry needs
#ifdef RS_SCORE_STATS
  {
    unsigned long long h[12];
    cudaMemcpyFromSymbol( h, g_score_stats, sizeof( h ) );
    fprintf( stderr, "score stats (cumulative): warps %llu points %llu stage1_survivors %llu pruned_after_A %llu listed_in_unpruned %llu pruned_midway %llu "
                     "skipped_by_midway %llu searched %llu found %llu\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8] );
  }
#endif
  return RSGPU_OK;
}

} // extern "C"
