/* TEST INFRASTRUCTURE ONLY - never shipped, never linked by anything under rescan_b200/ or integration/_build/*_rsgpu.
 *
 * A host stand-in for the handful of rsgpu entry points that integration/rsgpu_dropin_st.cpp calls, implemented with the
 * CPU oracle (oracle/rescan_oracle.c).  Linking the segment_transfer shim against it instead of librsgpu.so lets the CPU
 * tier (`pytest -m "not gpu"`) exercise the shim's HOST logic - the placement order, the two labelling passes, the mask
 * cache of the coverage term, the edge de-duplication, the label maps - against what the pure-CPU reference build decided
 * (tests/golden/dropin_st.npz), the same way tests/fake_api.py stands in for the device in the multi-rank step tests.
 * It says nothing about the CUDA kernels; those are compared with the oracle in the -m gpu tests. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "rsgpu.h"

/* the oracle's entry points (oracle/rescan_oracle.c) */
typedef struct orc_grid orc_grid_t;
orc_grid_t* orc_grid_build( const float* pts, int32_t n, float radius );
void orc_grid_free( orc_grid_t* g );
float orc_icp_align( const float* p1, const float* n1, int32_t c1, const float* p2, const float* n2, int32_t c2, float* T1, const float* T2,
                     float max_dist, float max_angle, int32_t* n_iters_out );
void orc_assign_labels( const float* scan_pos, const float* scan_nor, int32_t V, const float* poses, const orc_grid_t* const* obj_grid,
                        const float* const* obj_nor, int32_t first, int32_t last, float radius, int8_t* labels, float* min_d );
void orc_unary_costs( const int32_t* labels, const uint8_t* label_is_static, int32_t V, int32_t L, int32_t* cost );
void orc_neighborhood( const orc_grid_t* grid, const float* pos, const float* nor, int32_t V, int32_t max_nn, float radius_sq, float dist_exp,
                       float angle_exp, int32_t* nbr, float* weight );
void orc_cov_rasterize( const float* pts, int32_t n, const float* pose, const float* origin, const int32_t* res, float voxel, uint8_t* grid );
int32_t orc_poisson_level( const float* pos0, int32_t n, float voxel, int32_t level, int32_t* out_idx );
void orc_plane_inlier_counts( const float* pts, const uint8_t* active, int32_t n, const float* planes, int32_t n_planes, float dist_threshold,
                              int32_t* counts );

struct rsgpu_grid { orc_grid_t* g; float* pts; float* nor; int32_t n; };
struct rsgpu_cloud { float* pos; float* nor; int32_t n; };

static int64_t g_calls = 0;

static float* dup_floats( const float* p, size_t n )
{
  float* q = (float*)malloc( sizeof( float ) * ( n ? n : 1 ) );
  if( n ) { memcpy( q, p, sizeof( float ) * n ); }
  return q;
}

const char* rsgpu_last_error( void ) { return "fake rsgpu (oracle-backed test stand-in)"; }
int64_t rsgpu_launch_count( void ) { return g_calls; }

int rsgpu_grid_create( const float* pts, int32_t n_pts, float radius, rsgpu_grid_t** out )
{
  if( !pts || n_pts <= 0 || !out ) { return RSGPU_ERR_INVALID; }
  rsgpu_grid_t* g = (rsgpu_grid_t*)calloc( 1, sizeof( *g ) );
  g->g = orc_grid_build( pts, n_pts, radius ); g->pts = dup_floats( pts, 3 * (size_t)n_pts ); g->n = n_pts;
  *out = g; g_calls++;
  return RSGPU_OK;
}
int rsgpu_grid_set_normals( rsgpu_grid_t* g, const float* normals )
{
  if( !g || !normals ) { return RSGPU_ERR_INVALID; }
  free( g->nor ); g->nor = dup_floats( normals, 3 * (size_t)g->n );
  return RSGPU_OK;
}
void rsgpu_grid_destroy( rsgpu_grid_t* g )
{
  if( !g ) { return; }
  orc_grid_free( g->g ); free( g->pts ); free( g->nor ); free( g );
}
int rsgpu_cloud_create( const float* pos, const float* nor, int32_t n_pts, rsgpu_cloud_t** out )
{
  if( n_pts < 0 || !out || ( n_pts > 0 && ( !pos || !nor ) ) ) { return RSGPU_ERR_INVALID; }
  rsgpu_cloud_t* c = (rsgpu_cloud_t*)calloc( 1, sizeof( *c ) );
  c->pos = dup_floats( pos, 3 * (size_t)n_pts ); c->nor = dup_floats( nor, 3 * (size_t)n_pts ); c->n = n_pts;
  *out = c;
  return RSGPU_OK;
}
void rsgpu_cloud_destroy( rsgpu_cloud_t* c )
{
  if( !c ) { return; }
  free( c->pos ); free( c->nor ); free( c );
}

int rsgpu_icp_align_batch( const rsgpu_cloud_t* object, const rsgpu_grid_t* scan, float* T1, int32_t n_batch, const float* T2, float max_dist,
                           float max_angle, float* errs, int32_t* iters )
{
  if( !object || !scan || !scan->nor || !T1 || !T2 || !errs ) { return RSGPU_ERR_INVALID; }
  for( int32_t b = 0; b < n_batch; ++b )
  {
    int32_t it = 0;
    errs[b] = orc_icp_align( object->pos, object->nor, object->n, scan->pts, scan->nor, scan->n, T1 + 16 * b, T2, max_dist, max_angle, &it );
    if( iters ) { iters[b] = it; }
  }
  g_calls++;
  return RSGPU_OK;
}

int rsgpu_assign_labels( const float* scan_pos, const float* scan_nor, int32_t n_vertices, const float* poses, const rsgpu_grid_t* const* object_grids,
                         int32_t first, int32_t last, float radius, int8_t* labels, float* min_dists )
{
  if( first > last ) { return RSGPU_ERR_INVALID; }
  if( first == last ) { return RSGPU_OK; }
  const orc_grid_t** grids = (const orc_grid_t**)calloc( (size_t)last, sizeof( *grids ) );
  const float** nors = (const float**)calloc( (size_t)last, sizeof( *nors ) );
  for( int32_t i = first; i < last; ++i ) { grids[i] = object_grids[i]->g; nors[i] = object_grids[i]->nor; }
  orc_assign_labels( scan_pos, scan_nor, n_vertices, poses, grids, nors, first, last, radius, labels, min_dists );
  free( grids ); free( nors ); g_calls++;
  return RSGPU_OK;
}

int rsgpu_unary_costs( const int32_t* labels, const uint8_t* label_is_static, int32_t n_vertices, int32_t n_labels, int32_t* data_cost )
{
  orc_unary_costs( labels, label_is_static, n_vertices, n_labels, data_cost ); g_calls++;
  return RSGPU_OK;
}

int rsgpu_neighborhood( const rsgpu_grid_t* grid, const float* pos, const float* nor, int32_t n_vertices, int32_t max_nn, float radius_sq,
                        float dist_exp, float angle_exp, int32_t* neighbors, float* weights )
{
  if( !grid ) { return RSGPU_ERR_INVALID; }
  orc_neighborhood( grid->g, pos, nor, n_vertices, max_nn, radius_sq, dist_exp, angle_exp, neighbors, weights ); g_calls++;
  return RSGPU_OK;
}

int rsgpu_rasterize_points( const float* pts, int32_t n, const float* pose, const float origin[3], const int32_t res[3], float voxel, uint8_t* grid )
{
  orc_cov_rasterize( pts, n, pose, origin, res, voxel, grid ); g_calls++;
  return RSGPU_OK;
}

int rsgpu_coverage_masks( const rsgpu_cloud_t* const* objects, const float* poses, int32_t n_poses, const float origin[3], const int32_t res[3],
                          float voxel, const uint8_t* scene_grid, uint32_t* out_masks, int32_t n_words, int32_t* n_lit )
{
  const size_t n_cells = (size_t)res[0] * res[1] * res[2];
  int32_t* rank = (int32_t*)malloc( sizeof( int32_t ) * n_cells );
  int32_t lit = 0;
  for( size_t c = 0; c < n_cells; ++c ) { rank[c] = scene_grid[c] > 0 ? lit++ : -1; }
  *n_lit = lit;
  if( n_poses == 0 || !out_masks ) { free( rank ); return RSGPU_OK; }
  if( n_words < ( lit + 31 ) / 32 ) { free( rank ); return RSGPU_ERR_INVALID; }
  uint8_t* tmp = (uint8_t*)malloc( n_cells );
  memset( out_masks, 0, sizeof( uint32_t ) * (size_t)n_poses * n_words );
  for( int32_t i = 0; i < n_poses; ++i )
  {
    memset( tmp, 0, n_cells );
    orc_cov_rasterize( objects[i]->pos, objects[i]->n, poses + 16 * (size_t)i, origin, res, voxel, tmp );
    for( size_t c = 0; c < n_cells; ++c )
    {
      if( tmp[c] && rank[c] >= 0 ) { out_masks[(size_t)i * n_words + ( rank[c] >> 5 )] |= 1u << ( rank[c] & 31 ); }
    }
  }
  free( tmp ); free( rank ); g_calls++;
  return RSGPU_OK;
}

int rsgpu_plane_inlier_counts( const float* pts, const uint8_t* active, int32_t n_pts, const float* planes, int32_t n_planes, float dist_threshold,
                               int32_t* counts )
{
  orc_plane_inlier_counts( pts, active, n_pts, planes, n_planes, dist_threshold, counts ); g_calls++;
  return RSGPU_OK;
}

int rsgpu_poisson_level( const float* pts, int32_t n, float voxel, int32_t max_n_neigh, int32_t* out_indices, int32_t* n_out, int32_t* n_rounds )
{
  /* the oracle takes the level and derives max_n_neigh = 1024 * level / 4 itself (rs_pointcloud.h:994-995) */
  const int32_t level = max_n_neigh / 256;
  if( level < 1 || level > 4 || level * 256 != max_n_neigh ) { return RSGPU_ERR_INVALID; }
  *n_out = orc_poisson_level( pts, n, voxel, level, out_indices );
  if( n_rounds ) { *n_rounds = 0; }
  g_calls++;
  return RSGPU_OK;
}
