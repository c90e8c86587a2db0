// Drop-in replacement of the hot-path symbols of mhalber/Rescan's segment_transfer executable, on top of the rsgpu C ABI
// (include/rsgpu.h).  Linked FIRST, together with the reference's UNMODIFIED apps/segment_transfer/{main,
// arrangement_optimization,database_update}.cpp and lib/rs/rs_pointcloud_filters.cpp (all compiled -fPIC so that their
// calls to these symbols stay interposable; see integration/Makefile and INTEGRATION.md), it gives a `segment_transfer`
// binary whose scan rasterisation, coverage term, ICP refinement, label transfer, unary data terms and neighbourhood
// weights run on the GPU while loading, plane detection, the greedy / simulated-annealing drivers, the graph cut (gco)
// and saving stay reference host code.  Nothing here is copied from the reference: its headers are included in place for
// their TYPES only (no *_IMPLEMENTATION define).
//
//   replaces                                   (reference)                                           with
//   rsao_rasterize_scene_to_grid               apps/segment_transfer/arrangement_optimization.cpp:1064  rsgpu_rasterize_points
//   rsao__compute_scene_coverage_score         apps/segment_transfer/arrangement_optimization.cpp:343   rsgpu_coverage_masks (once per placement) + OR / popcount
//   icp_align                                  lib/rs/icp.h:416                                         rsgpu_icp_align_batch
//   rspf_arrangement_to_labels                 lib/rs/rs_pointcloud_filters.cpp:780                     rsgpu_assign_labels
//   rspf_compute_neighborhood                  lib/rs/rs_pointcloud_filters.cpp:674                     rsgpu_neighborhood
//   rspf_smooth_labels                         lib/rs/rs_pointcloud_filters.cpp:881                     rsgpu_unary_costs + the two above, then gco as before
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cassert>
#include <map>
#include <sys/mman.h>
#include <unordered_set>
#include <utility>
#include <thread>
#include <vector>

#include "msh/msh_std.h"
#include "msh/msh_vec_math.h"
#include "msh/msh_geometry.h"
#include "msh/msh_hash_grid.h"
#include "mg/hashtable.h"
#include "rs_pointcloud.h"
#include "rs_distance_function.h"
#include "rs_database.h"
#include "intersect.h"
#include "GCoptimization.h"
#include "rs_pointcloud_filters.h"
#include "arrangement_optimization.h"

#include "rsgpu.h"

namespace
{
void die( const char* what )
{
  // the reference reports errors with printf + exit; there is no CPU fallback to take
  fprintf( stderr, "rsgpu drop-in: %s failed: %s\n", what, rsgpu_last_error() );
  exit( -1 );
}
#define RSGPU_OR_DIE( call ) do { if( ( call ) != RSGPU_OK ) { die( #call ); } } while( 0 )

const int32_t LABEL_LVL = 1;         // RSPF_POINTCLOUD_LEVEL (rs_pointcloud_filters.cpp:21)
const int32_t MAX_INSTANCES = 1024;  // RSPF_MAX_INSTANCES (:20)

// ---- coverage term --------------------------------------------------------------------------------------------------
// One bit mask per distinct placement (object, pose) over the scan's lit cells; valid for the scan grid it was made for.
struct PlacementKey
{
  int32_t object_idx;
  float pose[16];
  bool operator<( const PlacementKey& o ) const
  {
    if( object_idx != o.object_idx ) { return object_idx < o.object_idx; }
    return memcmp( pose, o.pose, sizeof( pose ) ) < 0;
  }
};
struct CoverageCache
{
  const uint8_t* grid_data = NULL;   // the scan grid the masks belong to
  int32_t n_cells = 0, n_lit = 0, n_words = 0;
  std::map<PlacementKey, std::vector<uint32_t> > masks;
  std::map<int32_t, rsgpu_cloud_t*> clouds; // level-2 cloud of an object, by object index
  size_t n_evaluations = 0, n_rasterised = 0;
} g_cov;

void coverage_reset( const isect_grid3d_t* grd )
{
  g_cov.grid_data = grd->data; g_cov.n_cells = grd->n_cells;
  g_cov.masks.clear();
  for( std::map<int32_t, rsgpu_cloud_t*>::iterator it = g_cov.clouds.begin(); it != g_cov.clouds.end(); ++it ) { rsgpu_cloud_destroy( it->second ); }
  g_cov.clouds.clear();
  g_cov.n_lit = -1; g_cov.n_words = 0;
}

rsgpu_cloud_t* coverage_cloud( rsdb_t* rsdb, int32_t object_idx )
{
  std::map<int32_t, rsgpu_cloud_t*>::iterator it = g_cov.clouds.find( object_idx );
  if( it != g_cov.clouds.end() ) { return it->second; }
  const rs_pointcloud_t* pc = rsdb->objects[object_idx].shape;
  const int32_t lvl = 2; // arrangement_optimization.cpp:1089
  rsgpu_cloud_t* c = NULL;
  RSGPU_OR_DIE( rsgpu_cloud_create( &pc->positions[lvl][0].x, &pc->normals[lvl][0].x, (int32_t)pc->n_pts[lvl], &c ) );
  g_cov.clouds[object_idx] = c;
  return c;
}

void grid_geometry( const isect_grid3d_t* grd, float origin[3], int32_t res[3] )
{
  origin[0] = grd->bbox.min_p.x; origin[1] = grd->bbox.min_p.y; origin[2] = grd->bbox.min_p.z;
  res[0] = grd->x_res; res[1] = grd->y_res; res[2] = grd->z_res;
}

int32_t placement_cmp( const void* a, const void* b, void* rsdb_ptr )
{
  // the ordering of rsfp__static_plcmnt_cmp (rs_pointcloud_filters.cpp:725-736): dynamic before static, then by class
  const rsdb_t* rsdb = (const rsdb_t*)rsdb_ptr;
  const rs_obj_plcmnt_t* pa = (const rs_obj_plcmnt_t*)a;
  const rs_obj_plcmnt_t* pb = (const rs_obj_plcmnt_t*)b;
  const int32_t key_a = ( rsdb_is_object_static( rsdb, pa->object_idx ) ? 1 : 0 ) << 10 | rsdb->objects[pa->object_idx].class_idx;
  const int32_t key_b = ( rsdb_is_object_static( rsdb, pb->object_idx ) ? 1 : 0 ) << 10 | rsdb->objects[pb->object_idx].class_idx;
  return key_a - key_b;
}
// A large buffer the device will fill: 2 MB-aligned and advised for transparent huge pages, so that the copy back touches
// (and the driver pins) tens of pages instead of tens of thousands.  Released with free().  [applied after the last GPU
// session of round 1: functionally covered by the tests, its effect on the copy time is not measured yet - DESIGN.md 9]
void* result_buffer( size_t bytes )
{
  void* p = NULL;
  const size_t huge = (size_t)2 << 20;
  if( bytes < 4 * huge ) { return malloc( bytes ? bytes : 1 ); }
  if( posix_memalign( &p, huge, ( bytes + huge - 1 ) / huge * huge ) != 0 ) { return malloc( bytes ); }
#ifdef MADV_HUGEPAGE
  madvise( p, ( bytes + huge - 1 ) / huge * huge, MADV_HUGEPAGE );
#endif
  return p;
}

rsdb_t* g_sort_rsdb = NULL;
int placement_cmp_qsort( const void* a, const void* b ) { return placement_cmp( a, b, g_sort_rsdb ); }
} // namespace

void
rsao_rasterize_scene_to_grid( rs_scene_t* scn, isect_grid3d_t* grd, float quality_threshold )
{
  const int32_t lvl = 2;
  const rs_pointcloud_t* pc = scn->shape;
  std::vector<float> pts;
  pts.reserve( 3 * pc->n_pts[lvl] );
  for( size_t i = 0; i < pc->n_pts[lvl]; ++i )
  {
    if( pc->qualities[lvl][i] < quality_threshold ) { continue; }
    pts.push_back( pc->positions[lvl][i].x ); pts.push_back( pc->positions[lvl][i].y ); pts.push_back( pc->positions[lvl][i].z );
  }
  float origin[3]; int32_t res[3];
  grid_geometry( grd, origin, res );
  memset( grd->data, 0, grd->n_cells * sizeof( grd->data[0] ) );
  RSGPU_OR_DIE( rsgpu_rasterize_points( pts.data(), (int32_t)( pts.size() / 3 ), NULL, origin, res, grd->voxel_size, grd->data ) );
  coverage_reset( grd );
  printf( "ARRANGEMENT_OPTIMIZATION: scan rasterised into %d x %d x %d cells (GPU)\n", res[0], res[1], res[2] );
}

float
rsao__compute_scene_coverage_score( rsdb_t* rsdb, msh_array( rs_obj_plcmnt_t ) arrangement, rsao_opts_t* opts, int32_t verbose )
{
  const isect_grid3d_t* grd = opts->scn_grd;
  if( g_cov.grid_data != grd->data || g_cov.n_cells != grd->n_cells ) { coverage_reset( grd ); }
  float origin[3]; int32_t res[3];
  grid_geometry( grd, origin, res );
  if( g_cov.n_lit < 0 )
  {
    int32_t n_lit = 0;
    RSGPU_OR_DIE( rsgpu_coverage_masks( NULL, NULL, 0, origin, res, grd->voxel_size, grd->data, NULL, 0, &n_lit ) );
    g_cov.n_lit = n_lit; g_cov.n_words = ( n_lit + 31 ) / 32;
  }
  g_cov.n_evaluations++;
  if( g_cov.n_lit == 0 ) { return 0.0f; } // scn_grd_valid_cells == 0 (:368)
  // masks of the placements not seen before, in one call
  const size_t n = msh_array_len( arrangement );
  std::vector<PlacementKey> keys; std::vector<const rsgpu_cloud_t*> new_clouds; std::vector<float> new_poses; std::vector<PlacementKey> new_keys;
  for( size_t i = 0; i < n; ++i )
  {
    const rs_obj_plcmnt_t* p = &arrangement[i];
    if( rsdb_is_object_static( rsdb, p->object_idx ) ) { continue; } // :1094
    PlacementKey k; k.object_idx = p->object_idx; memcpy( k.pose, p->pose.data, sizeof( k.pose ) );
    keys.push_back( k );
    if( g_cov.masks.find( k ) == g_cov.masks.end() && std::find_if( new_keys.begin(), new_keys.end(), [&]( const PlacementKey& o ) { return !( o < k ) && !( k < o ); } ) == new_keys.end() )
    {
      new_keys.push_back( k ); new_clouds.push_back( coverage_cloud( rsdb, p->object_idx ) );
      new_poses.insert( new_poses.end(), k.pose, k.pose + 16 );
    }
  }
  if( !new_keys.empty() )
  {
    std::vector<uint32_t> out( new_keys.size() * (size_t)g_cov.n_words );
    int32_t n_lit = 0;
    RSGPU_OR_DIE( rsgpu_coverage_masks( new_clouds.data(), new_poses.data(), (int32_t)new_keys.size(), origin, res, grd->voxel_size, grd->data,
                                        out.data(), g_cov.n_words, &n_lit ) );
    for( size_t i = 0; i < new_keys.size(); ++i )
    {
      g_cov.masks[new_keys[i]].assign( out.begin() + i * g_cov.n_words, out.begin() + ( i + 1 ) * g_cov.n_words );
    }
    g_cov.n_rasterised += new_keys.size();
  }
  // |cells lit by the scan and by the arrangement| = popcount( OR of the placements' masks )
  std::vector<uint32_t> u( g_cov.n_words, 0u );
  for( size_t i = 0; i < keys.size(); ++i )
  {
    const std::vector<uint32_t>& m = g_cov.masks[keys[i]];
    for( int32_t w = 0; w < g_cov.n_words; ++w ) { u[w] |= m[w]; }
  }
  int32_t agreement = 0;
  for( int32_t w = 0; w < g_cov.n_words; ++w ) { agreement += __builtin_popcount( u[w] ); }
  const float score = (float)agreement / (float)g_cov.n_lit;
  msh_cprintf( verbose, "Coverage score: %f | %d %d | %zu placements rasterised so far (GPU)\n", score, g_cov.n_lit, agreement, g_cov.n_rasterised );
  return score;
}

extern "C" float
icp_align( msh_vec3_t* pts1, msh_vec3_t* nor1, int32_t n_pts1, msh_vec3_t* pts2, msh_vec3_t* nor2, int32_t n_pts2,
           msh_mat4_t* T1, msh_mat4_t T2, float max_dist, float max_angle, bool verbose )
{
  (void)verbose;
  // no handle cache here: database_update.cpp frees and re-allocates the shapes it aligns, so a pointer does not name a cloud
  rsgpu_cloud_t* obj = NULL; rsgpu_grid_t* scn = NULL;
  RSGPU_OR_DIE( rsgpu_cloud_create( &pts1[0].x, &nor1[0].x, n_pts1, &obj ) );
  RSGPU_OR_DIE( rsgpu_grid_create( &pts2[0].x, n_pts2, max_dist, &scn ) ); // icp.h:434-437; any cell size gives the same (exact) correspondences
  RSGPU_OR_DIE( rsgpu_grid_set_normals( scn, &nor2[0].x ) );
  float err = 1e6f;
  RSGPU_OR_DIE( rsgpu_icp_align_batch( obj, scn, T1->data, 1, T2.data, max_dist, max_angle, &err, NULL ) );
  rsgpu_grid_destroy( scn ); rsgpu_cloud_destroy( obj );
  return err;
}

void
rspf_arrangement_to_labels( rsdb_t* rsdb, rs_pointcloud_t* in_pc, msh_array( rs_obj_plcmnt_t ) arrangement, float radius, bool prioritize_static )
{
  printf( "LABEL_TRANSFER:   Starting copying labels from arrangement... (GPU)\n" );
  const int32_t lvl = LABEL_LVL;
  const int32_t n_pts = (int32_t)in_pc->n_pts[lvl];
  const size_t n_placements = msh_array_len( arrangement );
  uint64_t t1 = msh_time_now();
  if( n_placements > 127 ) { fprintf( stderr, "rsgpu drop-in: more than 127 placements do not fit the reference's int8 labels\n" ); exit( -1 ); }

  // the reference's order: a qsort of a copy, dynamic placements first (:826-835)
  std::vector<rs_obj_plcmnt_t> sorted( arrangement, arrangement + n_placements );
  g_sort_rsdb = rsdb;
  qsort( sorted.data(), n_placements, sizeof( rs_obj_plcmnt_t ), placement_cmp_qsort );
  g_sort_rsdb = NULL;
  size_t first_static = 0; // stays 0 when there is no static placement, like the reference (:830-835)
  for( size_t i = 0; i < n_placements; ++i ) { if( rsdb_is_object_static( rsdb, sorted[i].object_idx ) ) { first_static = i; break; } }

  std::vector<float> poses( 16 * n_placements );
  std::vector<rsgpu_grid_t*> grids( n_placements, (rsgpu_grid_t*)NULL );
  std::map<int32_t, rsgpu_grid_t*> by_object;
  for( size_t i = 0; i < n_placements; ++i )
  {
    memcpy( &poses[16 * i], sorted[i].pose.data, 64 );
    std::map<int32_t, rsgpu_grid_t*>::iterator it = by_object.find( sorted[i].object_idx );
    if( it == by_object.end() )
    {
      const rs_pointcloud_t* shape = rsdb->objects[sorted[i].object_idx].shape;
      rsgpu_grid_t* g = NULL;
      // rs_pointcloud_compute_search_grid builds every level's grid with radius 0.05 (rs_pointcloud.h:849-863)
      RSGPU_OR_DIE( rsgpu_grid_create( &shape->positions[lvl][0].x, (int32_t)shape->n_pts[lvl], 0.05f, &g ) );
      RSGPU_OR_DIE( rsgpu_grid_set_normals( g, &shape->normals[lvl][0].x ) );
      it = by_object.insert( std::make_pair( sorted[i].object_idx, g ) ).first;
    }
    grids[i] = it->second;
  }
  std::vector<int8_t> labels( n_pts, 0 );
  std::vector<float> min_dists( n_pts, 1e9f );
  const float* scan_pos = &in_pc->positions[lvl][0].x;
  const float* scan_nor = &in_pc->normals[lvl][0].x;
  if( n_pts > 0 && n_placements > 0 )
  {
    RSGPU_OR_DIE( rsgpu_assign_labels( scan_pos, scan_nor, n_pts, poses.data(), grids.data(), 0, (int32_t)first_static, radius, labels.data(), min_dists.data() ) );
    if( prioritize_static ) { std::fill( min_dists.begin(), min_dists.end(), 1e9f ); }
    const float radius2 = prioritize_static ? radius : 1.5f * radius; // :843-847
    RSGPU_OR_DIE( rsgpu_assign_labels( scan_pos, scan_nor, n_pts, poses.data(), grids.data(), (int32_t)first_static, (int32_t)n_placements, radius2, labels.data(),
                                       min_dists.data() ) );
  }
  for( std::map<int32_t, rsgpu_grid_t*>::iterator it = by_object.begin(); it != by_object.end(); ++it ) { rsgpu_grid_destroy( it->second ); }

  // temporary labels -> class / instance ids (:851-869)
  const int32_t unlabelled = rsdb_get_class_idx( rsdb, "unlabelled" );
  for( int32_t i = 0; i < n_pts; ++i )
  {
    int32_t class_idx = unlabelled, instance_idx = MAX_INSTANCES;
    if( labels[i] != 0 )
    {
      const rs_obj_plcmnt_t* p = &sorted[labels[i] - 1];
      class_idx = rsdb->objects[p->object_idx].class_idx;
      instance_idx = p->uidx;
    }
    in_pc->class_ids[lvl][i] = class_idx;
    in_pc->instance_ids[lvl][i] = instance_idx;
  }
  printf( "LABEL_TRANSFER:   Done in %fms\n", msh_time_diff_ms( msh_time_now(), t1 ) );
}

msh_array( rspf_edge_t )
rspf_compute_neighborhood( const rs_pointcloud_t* pc, int32_t lvl, int32_t max_nn, float radius_sq, float dist_exp, float angle_exp )
{
  const int32_t n = (int32_t)pc->n_pts[lvl];
  msh_array( rspf_edge_t ) edges = NULL;
  if( n == 0 ) { return edges; }
  rsgpu_grid_t* g = NULL;
  RSGPU_OR_DIE( rsgpu_grid_create( &pc->positions[lvl][0].x, n, 0.05f, &g ) ); // the cloud's own search grid (rs_pointcloud.h:849-863)
  RSGPU_OR_DIE( rsgpu_grid_set_normals( g, &pc->normals[lvl][0].x ) );
  std::vector<int32_t> nbr( (size_t)n * max_nn );
  std::vector<float> wgt( (size_t)n * max_nn );
  RSGPU_OR_DIE( rsgpu_neighborhood( g, &pc->positions[lvl][0].x, &pc->normals[lvl][0].x, n, max_nn, radius_sq, dist_exp, angle_exp, nbr.data(), wgt.data() ) );
  rsgpu_grid_destroy( g );
  // the reference's de-duplication: first edge seen per key max * n + min, the key computed in int32 (:73-78, 709-712).
  // Weights are symmetric in (i, j), so which of the two directions is seen first does not matter; for n > 46 340 the int32
  // key wraps and distinct edges can collide - then the surviving set depends on the visiting order inside a query, which
  // the reference leaves to its unsorted search
  std::unordered_set<int32_t> seen;
  seen.reserve( (size_t)n * 4 );
  for( int32_t i = 0; i < n; ++i )
  {
    for( int32_t j = 0; j < max_nn; ++j )
    {
      const int32_t k = nbr[(size_t)i * max_nn + j];
      if( k < 0 ) { continue; }
      const int32_t hi = std::max( i, k ), lo = std::min( i, k );
      const int32_t key = (int32_t)( (uint32_t)hi * (uint32_t)n + (uint32_t)lo );
      if( !seen.insert( key ).second ) { continue; }
      rspf_edge_t e = { i, k, wgt[(size_t)i * max_nn + j] };
      msh_array_push( edges, e );
    }
  }
  return edges;
}

void
rspf_smooth_labels( rsdb_t* rsdb, rs_pointcloud_t* in_pc )
{
  printf( "LABEL_TRANSFER:   Performing label smoothing... (GPU unary terms and neighbourhood)\n" );
  const int32_t lvl = LABEL_LVL;
  const int32_t n_pts = (int32_t)in_pc->n_pts[lvl];
  uint64_t gt1 = msh_time_now(), t1 = msh_time_now();
  const float radius = 0.05f;
  const float radius_sq = radius * radius;

  // labels = instance + 1, 0 for the class "unlabelled"; n_labels = largest instance id + 5 (:896-916)
  int32_t max_uidx = -1;
  for( int32_t i = 0; i < n_pts; ++i )
  {
    if( in_pc->instance_ids[lvl][i] < MAX_INSTANCES ) { max_uidx = msh_max( max_uidx, in_pc->instance_ids[lvl][i] ); }
  }
  const int32_t n_labels = max_uidx + 5;
  std::vector<int32_t> labels( n_pts ), label_to_class( n_labels, 0 ), label_to_instance( n_labels, 0 );
  const int32_t unlabelled = rsdb_get_class_idx( rsdb, (const char*)"unlabelled" );
  for( int32_t i = 0; i < n_pts; ++i )
  {
    const int32_t instance_idx = in_pc->instance_ids[lvl][i], class_idx = in_pc->class_ids[lvl][i];
    int32_t label = instance_idx + 1;
    if( class_idx == unlabelled ) { label = 0; }
    if( label < 0 || label >= n_labels ) { fprintf( stderr, "rsgpu drop-in: label %d outside [0, %d) (the reference writes out of bounds here)\n", label, n_labels ); exit( -1 ); }
    labels[i] = label;
    label_to_class[label] = class_idx;
    label_to_instance[label] = instance_idx;
  }

  t1 = msh_time_now();
  msh_array( rspf_edge_t ) edges = rspf_compute_neighborhood( in_pc, lvl, 8, radius_sq, 15.0f, 16.0f );
  printf( "LABEL_TRANSFER:      Neighborhood compatibility computation took %fms\n", msh_time_diff_ms( msh_time_now(), t1 ) );

  t1 = msh_time_now();
  // unary term (:926-939): 0 for the vertex's own label, else 30, 15 when its class is static, 1 when it is unlabelled
  std::vector<uint8_t> label_is_static( n_labels, 0 );
  for( int32_t i = 0; i < n_pts; ++i ) { label_is_static[labels[i]] = rsdb_is_class_static( rsdb, label_to_class[labels[i]] ) ? 1 : 0; }
  int32_t* data_cost = (int32_t*)result_buffer( (size_t)n_pts * n_labels * sizeof( int32_t ) );
  // gco reads the V x L table from HOST memory.  Filling it on the device (rsgpu_unary_costs: 0.04 ms of kernel) and copying
  // 4 V L bytes back costs six times the reference's own host loop inside this executable (114 vs 19 ms at C2 size,
  // profiles/dropin_r01.md: first touch of 61 MB of fresh pages behind a PCIe copy), so at THIS call site the table is
  // written where it is consumed - by all host cores, rows in blocks, from the labels the GPU produced;
  // RSGPU_DROPIN_UNARY=gpu takes the device path (same bytes: tests/test_gpu_dropin.py compares the labels gco returns).
  const char* unary_env = getenv( "RSGPU_DROPIN_UNARY" );
  if( n_pts > 0 && unary_env && strcmp( unary_env, "gpu" ) == 0 )
  {
    RSGPU_OR_DIE( rsgpu_unary_costs( labels.data(), label_is_static.data(), n_pts, n_labels, data_cost ) );
  }
  else if( n_pts > 0 )
  {
    const int n_threads = (int)std::max( 1u, std::min( 16u, std::thread::hardware_concurrency() ) );
    std::vector<std::thread> pool;
    for( int t = 0; t < n_threads; ++t )
    {
      pool.emplace_back( [&, t]() {
        const int32_t lo = (int32_t)( (int64_t)n_pts * t / n_threads ), hi = (int32_t)( (int64_t)n_pts * ( t + 1 ) / n_threads );
        for( int32_t i = lo; i < hi; ++i )
        {
          const int32_t own = labels[i];
          const int32_t cost = own == 0 ? 1 : ( label_is_static[own] ? 15 : 30 ); // (:930-933: unlabelled wins over static)
          int32_t* row = data_cost + (size_t)i * n_labels;
          for( int32_t l = 0; l < n_labels; ++l ) { row[l] = cost; }
          row[own] = 0;
        }
      } );
    }
    for( size_t t = 0; t < pool.size(); ++t ) { pool[t].join(); }
  }
  // Potts pairwise term (:941-950)
  const int32_t edge_cost = 10;
  int32_t* smooth_cost = (int32_t*)malloc( (size_t)n_labels * n_labels * sizeof( int32_t ) );
  for( int32_t l1 = 0; l1 < n_labels; l1++ ) { for( int32_t l2 = 0; l2 < n_labels; l2++ ) { smooth_cost[l1 + l2 * n_labels] = ( l1 == l2 ) ? 0 : edge_cost; } }
  printf( "LABEL_TRANSFER:      Data and smoothness terms setting took %fms\n", msh_time_diff_ms( msh_time_now(), t1 ) );

  t1 = msh_time_now();
  // the graph cut itself is unchanged reference host code (gco, :955-971)
  GCoptimizationGeneralGraph* gc = new GCoptimizationGeneralGraph( n_pts, n_labels );
  gc->setDataCost( data_cost );
  gc->setSmoothCost( smooth_cost );
  for( int32_t i = 0; i < n_pts; i++ ) { gc->setLabel( i, labels[i] ); }
  for( size_t i = 0; i < msh_array_len( edges ); ++i ) { gc->setNeighbors( edges[i].idx1, edges[i].idx2, (int32_t)( edges[i].weight * edge_cost ) ); }
  printf( "LABEL_TRANSFER:      Optimizing over %d pts. and %d labels\n", n_pts, n_labels );
  gc->swap( 2 );
  for( int32_t i = 0; i < n_pts; i++ ) { labels[i] = gc->whatLabel( i ); }
  delete gc;
  for( int32_t i = 0; i < n_pts; ++i )
  {
    in_pc->class_ids[lvl][i] = label_to_class[labels[i]];
    in_pc->instance_ids[lvl][i] = label_to_instance[labels[i]];
  }
  printf( "LABEL_TRANSFER:      Label optimization took: %fs.\n", msh_time_diff_sec( msh_time_now(), t1 ) );
  printf( "LABEL_TRANSFER:   Label smoothing took: %fs.\n", msh_time_diff_sec( msh_time_now(), gt1 ) );
  msh_array_free( edges );
  free( data_cost );
  free( smooth_cost );
}
