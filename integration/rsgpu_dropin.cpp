// Drop-in replacement of the hot-path symbols of mhalber/Rescan's pose_proposal executable, on top of the
// rsgpu C ABI (include/rsgpu.h).  Linked together with the reference's UNMODIFIED apps/pose_proposal/main.cpp
// (see integration/Makefile and INTEGRATION.md) it gives a `pose_proposal` binary whose dense pose search,
// verification, non-maxima suppression, ICP refinement and rescoring run on the GPU while loading, sorting and saving
// stay reference host code.  Nothing here is copied from the reference: the reference headers are included in place
// for their TYPES only (no *_IMPLEMENTATION define), exactly like apps/pose_proposal/pose_proposal.cpp:1-17 does.
//
//   replaces                                   (reference)                               with
//   mgs_propose_poses                          apps/pose_proposal/pose_proposal.cpp:325  rsgpu_propose_poses
//   mgs_compute_object_alignment_score         apps/pose_proposal/pose_proposal.cpp:93   rsgpu_score_poses
//   mgs_non_maxima_suppresion                  apps/pose_proposal/pose_proposal.cpp:371  rsgpu_nms
//   icp_align                                  lib/rs/icp.h:416                          rsgpu_icp_align_batch
//
// The reference loops over the objects serially (pose_proposal.cpp:190-250, :377); the objects are independent, so the
// two per-object loops replaced here hand every object to a host thread bound to its own rsgpu lane (stream): the
// latency-bound stages of one object run next to the dense search of another.  Results do not depend on that.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cassert>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "msh/msh_std.h"
#include "msh/msh_vec_math.h"
#include "msh/msh_geometry.h"
#include "msh/msh_hash_grid.h"
#include "mg/hashtable.h"
#include "msh/msh_ply.h"
#include "rs_pointcloud.h"
#include "rs_distance_function.h"
#include "rs_database.h"
#include "pose_proposal.h"

#include "rsgpu.h"

namespace
{
void die( const char* what )
{
  // the reference reports errors with printf + exit(-1) (main.cpp:141, 153); there is no CPU fallback to take
  fprintf( stderr, "rsgpu drop-in: %s failed: %s\n", what, rsgpu_last_error() );
  exit( -1 );
}
#define RSGPU_OR_DIE( call ) do { if( ( call ) != RSGPU_OK ) { die( #call ); } } while( 0 )

typedef std::pair<const void*, size_t> key_t;
std::map<key_t, rsgpu_grid_t*> g_grids;    // grids over (positions pointer, count): the scan levels never change in a run
std::map<key_t, rsgpu_cloud_t*> g_clouds;
std::mutex g_cache_mu;                     // the caches are filled from the lane threads too

// fn( i ) for i in [0, n) on up to rsgpu_lane_count() host threads, each bound to its own lane
template <class F>
void for_each_object_on_lanes( int32_t n, F fn )
{
  const int n_threads = std::min<int>( rsgpu_lane_count(), n );
  if( n_threads <= 1 ) { for( int32_t i = 0; i < n; ++i ) { fn( i ); } return; }
  std::atomic<int32_t> next( 0 );
  std::vector<std::thread> pool;
  for( int t = 0; t < n_threads; ++t )
  {
    pool.emplace_back( [&, t]() {
      RSGPU_OR_DIE( rsgpu_thread_attach( t ) );
      for( int32_t i = next.fetch_add( 1 ); i < n; i = next.fetch_add( 1 ) ) { fn( i ); }
    } );
  }
  for( size_t t = 0; t < pool.size(); ++t ) { pool[t].join(); }
}

rsgpu_grid_t* grid_for( const msh_vec3_t* pos, const msh_vec3_t* nor, size_t n )
{
  std::lock_guard<std::mutex> lk( g_cache_mu );
  key_t k( pos, n );
  std::map<key_t, rsgpu_grid_t*>::iterator it = g_grids.find( k );
  if( it != g_grids.end() ) { return it->second; }
  rsgpu_grid_t* g = NULL;
  // rs_pointcloud_compute_search_grid builds every level's grid with radius 0.05 (rs_pointcloud.h:849-863)
  RSGPU_OR_DIE( rsgpu_grid_create( &pos[0].x, (int32_t)n, 0.05f, &g ) );
  RSGPU_OR_DIE( rsgpu_grid_set_normals( g, &nor[0].x ) );
  g_grids[k] = g;
  return g;
}

// ---- look-ahead over main.cpp's refinement loop (apps/pose_proposal/main.cpp:175-204) ---------------------------------
// The unmodified main calls icp_align and mgs_compute_object_alignment_score once per proposal, one after the other; served
// one at a time each is a batch of 1 on the device (a chain of ~70 dependent iterations per call).  The lists main walks
// are the ones this shim handed back (mgs_non_maxima_suppresion keeps their address) plus the previous placements main
// appended, so at the FIRST icp_align call every refinement the loop is going to ask for is known: all of them are run
// then - one batch per object, the objects side by side on their lanes - and the calls that follow are answered from the
// table, each checked against the pose it was computed for (anything else - another caller, other parameters, a pose that
// is not in the lists - takes the direct path).  The scores are batched the same way at the first score call of the loop.
// Results are those of the one-by-one calls (batches are bit-identical to singles, tests/test_gpu_parity.py).
struct RefineEntry { float in_T[16], out_T[16], err, score; };
struct RefineList
{
  const void* lvl2_pos = NULL; const void* lvl1_pos = NULL;
  rs_pointcloud_t* shape = NULL;
  std::vector<RefineEntry> e;
  size_t next_icp = 0, next_score = 0;
  bool scored = false;
};
struct Lookahead
{
  rsdb_t* rsdb = NULL;
  msh_array( msh_array( pose_proposal_t ) ) * poses = NULL; // main's variable
  int nms_calls = 0;
  bool built = false, scores_built = false;
  std::vector<RefineList> lists;
  float max_dist = 0.f, max_angle = 0.f; float T2[16]; const void* scan_lvl2 = NULL;
} g_look;

bool lookahead_enabled()
{
  const char* e = getenv( "RSGPU_DROPIN_LOOKAHEAD" );
  return !( e && strcmp( e, "0" ) == 0 );
}

rsgpu_cloud_t* cloud_for( const msh_vec3_t* pos, const msh_vec3_t* nor, size_t n )
{
  std::lock_guard<std::mutex> lk( g_cache_mu );
  key_t k( pos, n );
  std::map<key_t, rsgpu_cloud_t*>::iterator it = g_clouds.find( k );
  if( it != g_clouds.end() ) { return it->second; }
  rsgpu_cloud_t* c = NULL;
  RSGPU_OR_DIE( rsgpu_cloud_create( &pos[0].x, &nor[0].x, (int32_t)n, &c ) );
  g_clouds[k] = c;
  return c;
}
} // namespace

float
mgs_compute_object_alignment_score( rs_pointcloud_t* object, rs_pointcloud_t* scene, int search_lvl, int query_lvl,
                                    msh_mat4_t xform, tmp_score_calc_storage_t* storage )
{
  static const float search_radii[5] = { 0.05f, 0.1f, 0.15f, 0.2f, 0.25f }; // pose_proposal.cpp:98
  rsgpu_grid_t* scn = grid_for( scene->positions[search_lvl], scene->normals[search_lvl], scene->n_pts[search_lvl] );
  if( g_look.built && search_lvl == 1 && query_lvl == 1 && storage->max_n_neigh == 32 )
  {
    if( !g_look.scores_built )
    {
      // first score call of main's loop: the scores of every refined pose of every object, one batch per object on the lanes
      for_each_object_on_lanes( (int32_t)g_look.lists.size(), [&]( int32_t li ) {
        RefineList& L = g_look.lists[li];
        if( L.e.empty() ) { return; }
        rsgpu_cloud_t* o1 = cloud_for( L.shape->positions[1], L.shape->normals[1], L.shape->n_pts[1] );
        std::vector<float> T( L.e.size() * 16 ), sc( L.e.size() );
        for( size_t j = 0; j < L.e.size(); ++j ) { memcpy( &T[16 * j], L.e[j].out_T, 64 ); }
        RSGPU_OR_DIE( rsgpu_score_poses( o1, scn, T.data(), (int64_t)L.e.size(), 32, search_radii[1], sc.data() ) );
        for( size_t j = 0; j < L.e.size(); ++j ) { L.e[j].score = sc[j]; }
        L.scored = true;
      } );
      g_look.scores_built = true;
    }
    for( size_t li = 0; li < g_look.lists.size(); ++li )
    {
      RefineList& L = g_look.lists[li];
      if( L.lvl1_pos != (const void*)object->positions[1] || !L.scored ) { continue; }
      for( size_t probe = 0; probe < L.e.size(); ++probe )
      {
        const size_t j = ( L.next_score + probe ) % L.e.size();
        if( memcmp( L.e[j].out_T, xform.data, 64 ) == 0 ) { L.next_score = j + 1; return L.e[j].score; }
      }
    }
  }
  rsgpu_cloud_t* obj = cloud_for( object->positions[query_lvl], object->normals[query_lvl], object->n_pts[query_lvl] );
  float score = 0.0f;
  RSGPU_OR_DIE( rsgpu_score_poses( obj, scn, xform.data, 1, storage->max_n_neigh, search_radii[search_lvl], &score ) );
  return score;
}

void
mgs_propose_poses( rsdb_t* rsdb, rs_pointcloud_t* input_scan, msh_array( msh_array( pose_proposal_t ) ) * proposed_poses,
                   const mgs_opts_t* opts, int verbose )
{
  uint64_t t0 = msh_time_now();
  const int32_t search_lvl = 1; // pose_proposal.cpp:178
  rsgpu_grid_t* scn = grid_for( input_scan->positions[search_lvl], input_scan->normals[search_lvl], input_scan->n_pts[search_lvl] );

  // the candidate lattice of mgs__initial_pose_proposals (pose_proposal.cpp:203-222), same float loop counters
  float spacing = opts->search_grid_spacing;
  float y_angle_inc = opts->search_grid_angle_delta;
  msh_vec3_t origin = input_scan->bbox.min_p;
  float length_x = input_scan->bbox.max_p.x - input_scan->bbox.min_p.x;
  float length_z = input_scan->bbox.max_p.z - input_scan->bbox.min_p.z;
  float height = 0.0f;
  std::vector<float> rotations, translations;
  for( float y_angle = 0.0f; y_angle < MSH_TWO_PI; y_angle += y_angle_inc )
  {
    msh_mat4_t r = msh_rotate( msh_mat4_identity(), y_angle, msh_vec3( 0.0f, 1.0f, 0.0f ) );
    rotations.insert( rotations.end(), r.data, r.data + 16 );
  }
  for( float ox = -spacing; ox < length_x + spacing; ox += spacing )
  {
    for( float oz = -spacing; oz < length_z + spacing; oz += spacing )
    {
      translations.push_back( origin.x + ox ); translations.push_back( height ); translations.push_back( origin.z + oz );
    }
  }
  const int32_t n_rot = (int32_t)( rotations.size() / 16 );
  const int64_t n_trans = (int64_t)( translations.size() / 3 );

  rsgpu_propose_opts_t po;
  rsgpu_propose_default_opts( &po ); // k = 64, r = 0.10, thresholds 0.25 / 0.35 / 0.40, every survivor kept
  int32_t n_objects = (int32_t)msh_array_len( rsdb->objects );
  std::vector<std::vector<float> > found( n_objects );
  std::vector<char> is_static( n_objects );
  for( int32_t i = 0; i < n_objects; ++i ) { is_static[i] = (char)rsdb_is_object_static( rsdb, i ); } // caches class ids in function statics: keep it on this thread
  for_each_object_on_lanes( n_objects, [&]( int32_t i ) {
    if( is_static[i] ) { return; } // pose_proposal.cpp:198
    rs_pointcloud_t* shape = rsdb->objects[i].shape;
    rsgpu_cloud_t* lv[3];
    for( int l = 0; l < 3; ++l ) { lv[l] = cloud_for( shape->positions[4 - l], shape->normals[4 - l], shape->n_pts[4 - l] ); }
    std::vector<float> out( (size_t)( n_trans > 0 ? n_trans : 1 ) * RSGPU_POSE_FLOATS );
    int64_t n_out = 0;
    RSGPU_OR_DIE( rsgpu_propose_poses( lv[0], lv[1], lv[2], scn, rotations.data(), n_rot, translations.data(), n_trans, &po,
                                       out.data(), NULL, n_trans, &n_out ) );
    out.resize( (size_t)n_out * RSGPU_POSE_FLOATS );
    found[i].swap( out );
  } );
  for( int32_t i = 0; i < n_objects; ++i )
  {
    msh_array( pose_proposal_t ) cur = NULL;
    const size_t n_out = found[i].size() / RSGPU_POSE_FLOATS;
    for( size_t j = 0; j < n_out; ++j )
    {
      pose_proposal_t p;
      memcpy( p.xform.data, &found[i][j * RSGPU_POSE_FLOATS], 64 );
      p.score = found[i][j * RSGPU_POSE_FLOATS + 16];
      msh_array_push( cur, p );
    }
    if( !is_static[i] ) { msh_cprintf( verbose, "POSE_PROPOSAL:      object %d: %d potential poses (GPU)\n", i, (int)n_out ); }
    msh_array_push( *proposed_poses, cur );
  }
  msh_cprintf( verbose, "POSE PROPOSAL: Done in %fs (rsgpu: %d rotations x %d translations per object)\n",
               msh_time_diff_sec( msh_time_now(), t0 ), (int)n_rot, (int)n_trans );
}

// mgs_non_maxima_suppresion (pose_proposal.cpp:371-452): per object, greedy keep-the-best / discard by voxel overlap > 0.5,
// centroid distance < dist_threshold or score < 0.01; survivors keep their order (:440-447)
void
mgs_non_maxima_suppresion( rsdb_t* rsdb, msh_array( msh_array( pose_proposal_t ) ) * proposed_poses, int32_t verbose, float dist_threshold )
{
  const int32_t n_objects = (int32_t)msh_array_len( *proposed_poses );
  // the first call precedes main's refinement loop (main.cpp:161), the second follows it (:205): the table of the look-ahead
  // is only valid in between
  g_look.rsdb = rsdb; g_look.poses = proposed_poses; g_look.nms_calls += 1;
  g_look.built = false; g_look.scores_built = false; g_look.lists.clear();
  std::vector<msh_vec3_t> centroid( n_objects );
  for( int32_t i = 0; i < n_objects; ++i ) { centroid[i] = rs_pointcloud_centroid( rsdb->objects[i].shape, 0 ); } // cached in the cloud (:1321)
  std::vector<std::vector<uint8_t> > keep( n_objects );
  for_each_object_on_lanes( n_objects, [&]( int32_t i ) {
    msh_array( pose_proposal_t ) cur = ( *proposed_poses )[i];
    const int32_t n = (int32_t)msh_array_len( cur );
    if( n == 0 ) { return; }
    rs_pointcloud_t* shape = rsdb->objects[i].shape;
    rsgpu_cloud_t* l3 = cloud_for( shape->positions[3], shape->normals[3], shape->n_pts[3] );
    rsgpu_cloud_t* l1 = cloud_for( shape->positions[1], shape->normals[1], shape->n_pts[1] );
    static_assert( sizeof( pose_proposal_t ) == RSGPU_POSE_FLOATS * sizeof( float ), "pose_proposal_t is 16 + 1 floats" );
    keep[i].resize( (size_t)n );
    RSGPU_OR_DIE( rsgpu_nms( l3, l1, &centroid[i].x, (const float*)cur, n, dist_threshold, keep[i].data() ) );
  } );
  for( int32_t i = 0; i < n_objects; ++i )
  {
    msh_array( pose_proposal_t ) cur = ( *proposed_poses )[i];
    const int32_t n = (int32_t)msh_array_len( cur );
    if( n == 0 ) { continue; }
    msh_array( pose_proposal_t ) kept = NULL;
    int32_t n_keep = 0;
    for( int32_t j = 0; j < n; ++j ) { if( keep[i][j] ) { msh_array_push( kept, cur[j] ); ++n_keep; } }
    msh_cprintf( verbose, "POSE_PROPOSAL: Non-max suppress. --> object %d keep: %5d discard: %5d (GPU)\n", i, n_keep, n - n_keep );
    ( *proposed_poses )[i] = kept;
    msh_array_free( cur );
  }
}

extern "C" float
icp_align( msh_vec3_t* pts1, msh_vec3_t* nor1, int32_t n_pts1, msh_vec3_t* pts2, msh_vec3_t* nor2, int32_t n_pts2,
           msh_mat4_t* T1, msh_mat4_t T2, float max_dist, float max_angle, bool verbose )
{
  (void)verbose;
  rsgpu_grid_t* scn = grid_for( pts2, nor2, (size_t)n_pts2 ); // any cell size gives the same (exact) correspondences
  if( lookahead_enabled() && g_look.poses && g_look.nms_calls == 1 && !g_look.built )
  {
    // first refinement main asks for: run them all (see the note at struct Lookahead)
    g_look.built = true;
    g_look.max_dist = max_dist; g_look.max_angle = max_angle; memcpy( g_look.T2, T2.data, 64 ); g_look.scan_lvl2 = pts2;
    const int32_t n_objects = (int32_t)msh_array_len( *g_look.poses );
    g_look.lists.assign( (size_t)n_objects, RefineList() );
    for( int32_t i = 0; i < n_objects; ++i )
    {
      if( rsdb_is_object_static( g_look.rsdb, i ) ) { continue; } // main.cpp:180 (function statics inside: this thread only)
      RefineList& L = g_look.lists[i];
      L.shape = g_look.rsdb->objects[i].shape;
      L.lvl2_pos = L.shape->positions[2]; L.lvl1_pos = L.shape->positions[1];
      msh_array( pose_proposal_t ) cur = ( *g_look.poses )[i];
      L.e.resize( msh_array_len( cur ) );
      for( size_t j = 0; j < L.e.size(); ++j ) { memcpy( L.e[j].in_T, cur[j].xform.data, 64 ); }
    }
    for_each_object_on_lanes( n_objects, [&]( int32_t i ) {
      RefineList& L = g_look.lists[i];
      if( L.e.empty() ) { return; }
      rsgpu_cloud_t* o2 = cloud_for( L.shape->positions[2], L.shape->normals[2], L.shape->n_pts[2] );
      std::vector<float> T( L.e.size() * 16 ), errs( L.e.size() );
      for( size_t j = 0; j < L.e.size(); ++j ) { memcpy( &T[16 * j], L.e[j].in_T, 64 ); }
      RSGPU_OR_DIE( rsgpu_icp_align_batch( o2, scn, T.data(), (int32_t)L.e.size(), g_look.T2, max_dist, max_angle, errs.data(), NULL ) );
      for( size_t j = 0; j < L.e.size(); ++j ) { memcpy( L.e[j].out_T, &T[16 * j], 64 ); L.e[j].err = errs[j]; }
    } );
  }
  if( g_look.built && pts2 == g_look.scan_lvl2 && max_dist == g_look.max_dist && max_angle == g_look.max_angle && memcmp( T2.data, g_look.T2, 64 ) == 0 )
  {
    for( size_t li = 0; li < g_look.lists.size(); ++li )
    {
      RefineList& L = g_look.lists[li];
      if( L.lvl2_pos != (const void*)pts1 ) { continue; }
      for( size_t probe = 0; probe < L.e.size(); ++probe )
      {
        const size_t j = ( L.next_icp + probe ) % L.e.size();
        if( memcmp( L.e[j].in_T, T1->data, 64 ) == 0 ) { L.next_icp = j + 1; memcpy( T1->data, L.e[j].out_T, 64 ); return L.e[j].err; }
      }
    }
  }
  rsgpu_cloud_t* obj = cloud_for( pts1, nor1, (size_t)n_pts1 );
  float err = 1e6f;
  RSGPU_OR_DIE( rsgpu_icp_align_batch( obj, scn, T1->data, 1, T2.data, max_dist, max_angle, &err, NULL ) );
  return err;
}
