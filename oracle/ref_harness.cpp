/* TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
 *
 * Function-level harness around the UNMODIFIED reference (mhalber/Rescan), compiled in place from
 * /root/reference by oracle/Makefile into oracle/_ref/librescan_ref.so.  No reference source is
 * copied: this TU instantiates the reference's single-header libraries exactly the way
 * apps/pose_proposal/main.cpp:1-45 does and wraps the hot-path entry points in a flat C ABI that
 * Python (ctypes) can drive with raw arrays:
 *
 *   msh_hash_grid_init_3d / _radius_search / _knn_search      (lib/msh/msh_hash_grid.h:218-230)
 *   rs_pointcloud_compute_levels                              (lib/rs/rs_pointcloud.h:1305)
 *   mgs_compute_object_alignment_score / mgs_propose_poses    (apps/pose_proposal/pose_proposal.h:36-61)
 *   icp_find_corrs / icp_estimate_rigid_xform_pt2pl / icp_align (lib/rs/icp.h:84-115)
 *   rspf_arrangement_to_labels / rspf_compute_neighborhood / rspf_smooth_labels
 *                                                             (lib/rs/rs_pointcloud_filters.h:60-74)
 *
 * It is used (a) to pin oracle/rescan_oracle.c and to write tests/golden/, (b) as the
 * "reference" CPU baseline of bench.py.
 */
#define MSH_STD_IMPLEMENTATION
#define MSH_PLY_IMPLEMENTATION
#define MSH_ARGPARSE_IMPLEMENTATION
#define MSH_VEC_MATH_IMPLEMENTATION
#define MSH_GEOMETRY_IMPLEMENTATION
#define MSH_HASH_GRID_IMPLEMENTATION
#define RS_POINTCLOUD_IMPLEMENTATION
#define RS_DISTANCE_FUNCTION_IMPLEMENTATION
#define RS_DATABASE_IMPLEMENTATION
#define FILEPATH_HELPERS_IMPLEMENTATION
#define HASHTABLE_IMPLEMENTATION
#define ICP_IMPLEMENTATION

#include <cassert>
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdarg>
#include <cstddef>
#include <cstdbool>
#include <cstdio>
#include <cstdlib>
#include <cfloat>
#include <cctype>

#include "msh/msh_std.h"
#include "msh/msh_argparse.h"
#include "msh/msh_vec_math.h"
#include "msh/msh_geometry.h"
#include "msh/msh_ply.h"
#include "msh/msh_hash_grid.h"
#include "mg/hashtable.h"
#include "icp.h"
#include "filepath_helpers.h"
#include "rs_pointcloud.h"
#include "rs_database.h"
#include "rs_distance_function.h"
#include "pose_proposal.h"
#include "intersect.h" /* declarations only: the implementation is instantiated in pose_proposal.cpp's TU */
#include "GCoptimization.h"
#include "rs_pointcloud_filters.h"

#include <vector>
#include <chrono>
#if defined(_OPENMP)
#include <omp.h>
#endif

/* explicit instantiations for every element type pushed into an msh_array by a TU that only sees the
   template declaration (pose_proposal.cpp, rs_pointcloud_filters.cpp) */
template int* msh_array__grow<int>(int*, unsigned long long, unsigned long long);
template rs_object_placement* msh_array__grow<rs_object_placement>(rs_object_placement*, unsigned long long, unsigned long long);
template pose_proposal* msh_array__grow<pose_proposal>(pose_proposal*, unsigned long long, unsigned long long);
template pose_proposal** msh_array__grow<pose_proposal*>(pose_proposal**, unsigned long long, unsigned long long);
template mark* msh_array__grow<mark>(mark*, unsigned long long, unsigned long long);
template rspf_edge_t* msh_array__grow<rspf_edge_t>(rspf_edge_t*, unsigned long long, unsigned long long);
template rspf_plane_model_t* msh_array__grow<rspf_plane_model_t>(rspf_plane_model_t*, unsigned long long, unsigned long long);
template unsigned long* msh_array__grow<unsigned long>(unsigned long*, unsigned long long, unsigned long long);
template msh_vec3_t* msh_array__grow<msh_vec3_t>(msh_vec3_t*, unsigned long long, unsigned long long);
template float* msh_array__grow<float>(float*, unsigned long long, unsigned long long);

gco_capture_t& gco_last_capture() { static gco_capture_t c; return c; }

static double now_s()
{
  return std::chrono::duration<double>( std::chrono::steady_clock::now().time_since_epoch() ).count();
}

static msh_mat4_t mat_from( const float* m ) { msh_mat4_t o; memcpy( o.data, m, 64 ); return o; }

extern "C" {

int ref_sizeof_hash_grid( void ) { return (int)sizeof(msh_hash_grid_t); }
int ref_openmp_threads( void )
{
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- hash grid */
void* ref_grid_build( const float* pts, int32_t n, float radius )
{
  msh_hash_grid_t* g = (msh_hash_grid_t*)calloc( 1, sizeof(msh_hash_grid_t) );
  msh_hash_grid_init_3d( g, pts, n, radius );
  return g;
}
void ref_grid_free( void* h )
{
  msh_hash_grid_t* g = (msh_hash_grid_t*)h;
  if( !g ) { return; }
  if( g->bin_table ) { msh_hg_map_free( g->bin_table ); }
  msh_hash_grid_term( g );
  free( g );
}
/* dims[3], cell_size, inv_cell_size, min[3], max[3], n_pts, n_bins, max_n_pts_in_bin */
void ref_grid_info( void* h, int64_t* dims, double* cell, float* minmax, int64_t* counts )
{
  msh_hash_grid_t* g = (msh_hash_grid_t*)h;
  dims[0] = g->width; dims[1] = g->height; dims[2] = g->depth;
  cell[0] = g->cell_size; cell[1] = g->_inv_cell_size;
  minmax[0] = g->min_pt.x; minmax[1] = g->min_pt.y; minmax[2] = g->min_pt.z;
  minmax[3] = g->max_pt.x; minmax[4] = g->max_pt.y; minmax[5] = g->max_pt.z;
  counts[0] = g->_n_pts; counts[1] = msh_hg_map_len( g->bin_table ); counts[2] = g->max_n_pts_in_bin;
}
/* the re-laid data buffer (16-B records, msh_hash_grid.h:238-242) */
void ref_grid_data( void* h, float* xyz, int32_t* idx )
{
  msh_hash_grid_t* g = (msh_hash_grid_t*)h;
  for( size_t i = 0; i < g->_n_pts; ++i )
  {
    xyz[3*i+0] = g->data_buffer[i].x; xyz[3*i+1] = g->data_buffer[i].y; xyz[3*i+2] = g->data_buffer[i].z;
    idx[i] = g->data_buffer[i].i;
  }
}
void ref_grid_set_threads( void* h, int n )
{
  msh_hash_grid_t* g = (msh_hash_grid_t*)h;
  g->_num_threads = (uint16_t)n; g->_dont_use_omp = ( n <= 1 );
}
/* offset/length of one cell, -1 when the cell is empty */
void ref_grid_cell( void* h, int64_t cell, int64_t* off_len )
{
  msh_hash_grid_t* g = (msh_hash_grid_t*)h;
  uint64_t* b = msh_hg_map_get( g->bin_table, (uint64_t)cell );
  if( !b ) { off_len[0] = -1; off_len[1] = 0; return; }
  off_len[0] = g->offsets[*b].offset; off_len[1] = g->offsets[*b].length;
}

static size_t ref_search( void* h, int knn, const float* q, size_t nq, float radius, size_t k, int sort,
                          float* d2, int32_t* idx, size_t* nn )
{
  msh_hash_grid_search_desc_t sd = {0};
  sd.query_pts = (float*)q; sd.n_query_pts = nq;
  sd.distances_sq = d2; sd.indices = idx; sd.n_neighbors = nn;
  sd.radius = radius; sd.max_n_neigh = k; sd.sort = sort;
  return knn ? msh_hash_grid_knn_search( (msh_hash_grid_t*)h, &sd )
             : msh_hash_grid_radius_search( (msh_hash_grid_t*)h, &sd );
}
size_t ref_radius_search( void* h, const float* q, size_t nq, float radius, size_t k, int sort,
                          float* d2, int32_t* idx, size_t* nn )
{ return ref_search( h, 0, q, nq, radius, k, sort, d2, idx, nn ); }
size_t ref_knn_search( void* h, const float* q, size_t nq, size_t k, int sort,
                       float* d2, int32_t* idx, size_t* nn )
{ return ref_search( h, 1, q, nq, 0.0f, k, sort, d2, idx, nn ); }

/* ---------------------------------------------------------------- clouds */
static void cloud_alloc_level( rs_pointcloud_t* pc, int lvl, int n )
{
  rs_pointcloud__allocate_level( pc, lvl, n );
}

/* level-0 arrays in; reference builds levels 1-4 and all five search grids (rs_pointcloud.h:1305) */
void* ref_cloud_from_level0( const float* pos, const float* nor, const int32_t* class_ids,
                             const int32_t* instance_ids, int32_t n )
{
  rs_pointcloud_t* pc = rs_pointcloud_init( 1 );
  cloud_alloc_level( pc, 0, n );
  for( int i = 0; i < n; ++i )
  {
    pc->positions[0][i] = msh_vec3( pos[3*i], pos[3*i+1], pos[3*i+2] );
    pc->normals[0][i]   = msh_vec3( nor[3*i], nor[3*i+1], nor[3*i+2] );
    pc->colors[0][i]    = msh_vec3( 0.5f, 0.5f, 0.5f );
    pc->radii[0][i]     = 0.01f;
    pc->qualities[0][i] = 1.0f;
    pc->class_ids[0][i]    = class_ids ? class_ids[i] : 0;
    pc->instance_ids[0][i] = instance_ids ? instance_ids[i] : 0;
  }
  rs_pointcloud_compute_levels( pc );
  return pc;
}

/* explicit levels in (any subset); bbox taken from level `bbox_lvl`; grids built like rs_pointcloud.h:849 */
void* ref_cloud_create( void ) { return rs_pointcloud_init( 1 ); }
void ref_cloud_set_level( void* h, int lvl, const float* pos, const float* nor, int32_t n, int update_bbox )
{
  rs_pointcloud_t* pc = (rs_pointcloud_t*)h;
  cloud_alloc_level( pc, lvl, n );
  if( update_bbox ) { mshgeo_bbox_reset( &pc->bbox ); }
  for( int i = 0; i < n; ++i )
  {
    pc->positions[lvl][i] = msh_vec3( pos[3*i], pos[3*i+1], pos[3*i+2] );
    pc->normals[lvl][i]   = msh_vec3( nor[3*i], nor[3*i+1], nor[3*i+2] );
    pc->colors[lvl][i]    = msh_vec3( 0.5f, 0.5f, 0.5f );
    pc->radii[lvl][i] = 0.01f; pc->qualities[lvl][i] = 1.0f;
    pc->class_ids[lvl][i] = 0; pc->instance_ids[lvl][i] = 0;
    if( update_bbox ) { mshgeo_bbox_union( &pc->bbox, pc->positions[lvl][i] ); }
  }
  rs_pointcloud_compute_search_grid( pc, lvl );
}
int32_t ref_cloud_n( void* h, int lvl ) { return (int32_t)((rs_pointcloud_t*)h)->n_pts[lvl]; }
void ref_cloud_get_level( void* h, int lvl, float* pos, float* nor )
{
  rs_pointcloud_t* pc = (rs_pointcloud_t*)h;
  memcpy( pos, pc->positions[lvl], pc->n_pts[lvl] * 12 );
  memcpy( nor, pc->normals[lvl], pc->n_pts[lvl] * 12 );
}
void ref_cloud_get_ids( void* h, int lvl, int32_t* class_ids, int32_t* instance_ids )
{
  rs_pointcloud_t* pc = (rs_pointcloud_t*)h;
  memcpy( class_ids, pc->class_ids[lvl], pc->n_pts[lvl] * 4 );
  memcpy( instance_ids, pc->instance_ids[lvl], pc->n_pts[lvl] * 4 );
}
void ref_cloud_bbox( void* h, float* mn_mx )
{
  rs_pointcloud_t* pc = (rs_pointcloud_t*)h;
  memcpy( mn_mx, &pc->bbox.min_p, 12 ); memcpy( mn_mx + 3, &pc->bbox.max_p, 12 );
}
void* ref_cloud_grid( void* h, int lvl ) { return ((rs_pointcloud_t*)h)->search_grids[lvl]; }
void ref_cloud_free( void* h ) { rs_pointcloud_free( (rs_pointcloud_t*)h, 1 ); }

/* ---------------------------------------------------------------- pose scoring */
float ref_score( void* obj, void* scene, int search_lvl, int query_lvl, const float* xform, int max_n_neigh )
{
  rs_pointcloud_t* o = (rs_pointcloud_t*)obj; rs_pointcloud_t* s = (rs_pointcloud_t*)scene;
  tmp_score_calc_storage_t st = allocate_tmp_calc_storage( o->n_pts[query_lvl], s->n_pts[query_lvl], max_n_neigh );
  float r = mgs_compute_object_alignment_score( o, s, search_lvl, query_lvl, mat_from( xform ), &st );
  free_tmp_calc_storage( &st );
  return r;
}

/* scores P poses; n_threads > 1 = OpenMP over poses in THIS harness (the scoring function is re-entrant
   given its own scratch).  Returns elapsed seconds of the scoring loop only. */
double ref_score_batch( void* obj, void* scene, int search_lvl, int query_lvl, const float* xforms,
                        int64_t n_poses, int max_n_neigh, float* out, int n_threads )
{
  rs_pointcloud_t* o = (rs_pointcloud_t*)obj; rs_pointcloud_t* s = (rs_pointcloud_t*)scene;
  if( n_threads < 1 ) { n_threads = 1; }
  /* the library's own OpenMP loop must not nest inside ours: it sizes its slices from the thread count
     captured at grid init (msh_hash_grid.h:395-408, 1122-1133) and would silently skip queries */
  msh_hash_grid_t* sg = s->search_grids[search_lvl];
  uint16_t saved_nt = sg->_num_threads; int32_t saved_no = sg->_dont_use_omp;
  if( n_threads > 1 ) { sg->_num_threads = 1; sg->_dont_use_omp = 1; }
  double t0 = now_s();
#if defined(_OPENMP)
  #pragma omp parallel num_threads(n_threads)
#endif
  {
    tmp_score_calc_storage_t st = allocate_tmp_calc_storage( o->n_pts[query_lvl], s->n_pts[query_lvl], max_n_neigh );
#if defined(_OPENMP)
    #pragma omp for schedule(dynamic, 16)
#endif
    for( int64_t p = 0; p < n_poses; ++p )
    {
      out[p] = mgs_compute_object_alignment_score( o, s, search_lvl, query_lvl, mat_from( xforms + 16 * p ), &st );
    }
    free_tmp_calc_storage( &st );
  }
  double dt = now_s() - t0;
  sg->_num_threads = saved_nt; sg->_dont_use_omp = saved_no;
  return dt;
}

/* ---------------------------------------------------------------- database + propose */
static const char* k_class_names[] = { "unlabelled", "wall", "floor", "ceiling", "cabinet", "chair", "table", "sofa", "box" };
enum { K_N_CLASSES = 9 };

void* ref_db_create( void )
{
  rsdb_t* db = rsdb_init();
  for( int i = 0; i < K_N_CLASSES; ++i )
  {
    char name[512] = {0};
    strncpy( name, k_class_names[i], 511 );
    rsdb_add_class( db, name, i );
  }
  return db;
}
int ref_db_n_classes( void ) { return K_N_CLASSES; }
const char* ref_db_class_name( int i ) { return k_class_names[i]; }
int ref_db_is_class_static( void* db, int class_idx ) { return rsdb_is_class_static( (rsdb_t*)db, class_idx ); }
int ref_db_add_object( void* db, void* cloud, int uidx, int class_idx )
{
  rs_object_t o = rsdb_object_init();
  o.uidx = uidx; o.class_idx = class_idx; o.shape = (rs_pointcloud_t*)cloud; o.filename = NULL;
  return rsdb_add_object( (rsdb_t*)db, &o );
}

/* mgs_propose_poses (pose_proposal.cpp:325).  Output: counts[n_objects], then a malloc'ed flat array of
   17 floats per proposal (16 xform column-major + score), object-major; caller frees with ref_free. */
int ref_propose_poses( void* db, void* scan, float spacing, float angle_delta, int32_t* counts, float** flat )
{
  mgs_opts_t opts; mgs_init_opts( &opts );
  if( spacing > 0 ) { opts.search_grid_spacing = spacing; }
  if( angle_delta > 0 ) { opts.search_grid_angle_delta = angle_delta; }
  msh_array( msh_array( pose_proposal_t ) ) pp = NULL;
  mgs_propose_poses( (rsdb_t*)db, (rs_pointcloud_t*)scan, &pp, &opts, 0 );
  size_t total = 0;
  for( size_t i = 0; i < msh_array_len( pp ); ++i ) { counts[i] = msh_array_len( pp[i] ); total += counts[i]; }
  float* out = (float*)malloc( (total ? total : 1) * 17 * sizeof(float) );
  size_t w = 0;
  for( size_t i = 0; i < msh_array_len( pp ); ++i )
    for( size_t j = 0; j < msh_array_len( pp[i] ); ++j )
    {
      memcpy( out + 17 * w, pp[i][j].xform.data, 64 ); out[17 * w + 16] = pp[i][j].score; w++;
    }
  *flat = out;
  int n = (int)msh_array_len( pp );
  for( size_t i = 0; i < msh_array_len( pp ); ++i ) { if( pp[i] ) { msh_array_free( pp[i] ); } }
  if( pp ) { msh_array_free( pp ); }
  return n;
}
void ref_free( void* p ) { free( p ); }

/* the exact pose xform the reference builds at pose_proposal.cpp:221-222 */
void ref_make_pose( float y_angle, float tx, float ty, float tz, float* out16 )
{
  msh_mat4_t xform = msh_rotate( msh_mat4_identity(), y_angle, msh_vec3( 0.0f, 1.0f, 0.0f ) );
  xform.col[3] = msh_vec4( tx, ty, tz, 1.0f );
  memcpy( out16, xform.data, 64 );
}

/* ---------------------------------------------------------------- ICP */
float ref_icp_align( const float* p1, const float* n1, int32_t c1, const float* p2, const float* n2, int32_t c2,
                     float* T1_inout, const float* T2, float max_dist, float max_angle )
{
  msh_mat4_t T1 = mat_from( T1_inout );
  float err = icp_align( (msh_vec3_t*)p1, (msh_vec3_t*)n1, c1, (msh_vec3_t*)p2, (msh_vec3_t*)n2, c2,
                         &T1, mat_from( T2 ), max_dist, max_angle, false );
  memcpy( T1_inout, T1.data, 64 );
  return err;
}

/* one icp_find_corrs (icp.h:306) against a grid built with `grid_radius`; outputs are n1-sized */
int32_t ref_icp_find_corrs( const float* p1, const float* n1, int32_t c1, const float* p2, const float* n2, int32_t c2,
                            const float* T1, const float* T2, float grid_radius, float max_dist, float max_angle,
                            float* cp1, float* cn1, float* cp2, float* cn2, float* w )
{
  msh_hash_grid_t i1 = {0}, i2 = {0};
  msh_hash_grid_init_3d( &i1, p1, c1, grid_radius );
  msh_hash_grid_init_3d( &i2, p2, c2, grid_radius );
  msh_vec3_t *a = NULL, *b = NULL, *c = NULL, *d = NULL; float* ww = NULL; int32_t n = 0;
  icp_find_corrs( (msh_vec3_t*)p1, (msh_vec3_t*)n1, c1, &i1, (msh_vec3_t*)p2, (msh_vec3_t*)n2, c2, &i2,
                  mat_from( T1 ), mat_from( T2 ), &a, &b, &c, &d, &ww, &n, max_dist, max_angle );
  memcpy( cp1, a, n * 12 ); memcpy( cn1, b, n * 12 ); memcpy( cp2, c, n * 12 ); memcpy( cn2, d, n * 12 );
  memcpy( w, ww, n * 4 );
  free( a ); free( b ); free( c ); free( d ); free( ww );
  msh_hg_map_free( i1.bin_table ); msh_hg_map_free( i2.bin_table );
  msh_hash_grid_term( &i1 ); msh_hash_grid_term( &i2 );
  return n;
}

float ref_icp_pt2pl( const float* cp1, const float* cp2, const float* cn2, const float* w, int32_t n, float* T1_inout )
{
  msh_mat4_t T1 = mat_from( T1_inout );
  float err = icp_estimate_rigid_xform_pt2pl( (msh_vec3_t*)cp1, (msh_vec3_t*)cp2, (msh_vec3_t*)cn2, (float*)w, n, &T1 );
  memcpy( T1_inout, T1.data, 64 );
  return err;
}

/* ---------------------------------------------------------------- labels / unary terms */
/* placements: object_idx[A], uidx[A], poses[A*16].  Writes class/instance ids into scan level 1. */
void ref_arrangement_to_labels( void* db, void* scan, const int32_t* object_idx, const int32_t* uidx,
                                const float* poses, int32_t n_plc, float radius, int prioritize_static )
{
  msh_array( rs_obj_plcmnt_t ) arr = NULL;
  for( int i = 0; i < n_plc; ++i )
  {
    rs_obj_plcmnt_t p; memset( &p, 0, sizeof(p) );
    p.uidx = uidx[i]; p.object_idx = object_idx[i]; p.pose = mat_from( poses + 16 * i ); p.score = 1.0f;
    msh_array_push( arr, p );
  }
  rspf_arrangement_to_labels( (rsdb_t*)db, (rs_pointcloud_t*)scan, arr, radius, prioritize_static != 0 );
  msh_array_free( arr );
}

/* rspf_smooth_labels with the recording gco stub; returns n_labels, fills capture getters below */
int ref_smooth_labels_capture( void* db, void* scan )
{
  rspf_smooth_labels( (rsdb_t*)db, (rs_pointcloud_t*)scan );
  return gco_last_capture().n_labels;
}
int64_t ref_capture_n_edges( void ) { return (int64_t)gco_last_capture().edge_a.size(); }
void ref_capture_get( int32_t* data_cost, int32_t* smooth_cost, int32_t* init_labels, int32_t* ea, int32_t* eb, int32_t* ew )
{
  gco_capture_t& c = gco_last_capture();
  if( data_cost )   memcpy( data_cost, c.data_cost.data(), c.data_cost.size() * 4 );
  if( smooth_cost ) memcpy( smooth_cost, c.smooth_cost.data(), c.smooth_cost.size() * 4 );
  if( init_labels ) memcpy( init_labels, c.init_labels.data(), c.init_labels.size() * 4 );
  if( ea ) memcpy( ea, c.edge_a.data(), c.edge_a.size() * 4 );
  if( eb ) memcpy( eb, c.edge_b.data(), c.edge_b.size() * 4 );
  if( ew ) memcpy( ew, c.edge_w.data(), c.edge_w.size() * 4 );
}

/* rspf_compute_neighborhood (rs_pointcloud_filters.cpp:674); returns count, malloc'ed arrays */
int64_t ref_compute_neighborhood( void* cloud, int lvl, int max_nn, float radius_sq, float dist_exp, float angle_exp,
                                  int32_t** a, int32_t** b, float** w )
{
  msh_array( rspf_edge_t ) e = rspf_compute_neighborhood( (rs_pointcloud_t*)cloud, lvl, max_nn, radius_sq, dist_exp, angle_exp );
  int64_t n = msh_array_len( e );
  *a = (int32_t*)malloc( (n ? n : 1) * 4 ); *b = (int32_t*)malloc( (n ? n : 1) * 4 ); *w = (float*)malloc( (n ? n : 1) * 4 );
  for( int64_t i = 0; i < n; ++i ) { (*a)[i] = e[i].idx1; (*b)[i] = e[i].idx2; (*w)[i] = e[i].weight; }
  msh_array_free( e );
  return n;
}

/* ---------------------------------------------------------------- coverage grids
   The arrangement optimiser's coverage term (apps/segment_transfer/arrangement_optimization.cpp:343-373) rasterises the
   scan (rsao_rasterize_scene_to_grid :1064-1080) and the placed dynamic objects (rsao__rasterize_arrangement_to_grid
   :1082-1106) into 5 cm grids built by isect_grid3d_init over the scan's bbox (apps/segment_transfer/main.cpp:323-325).
   Those two loops live in a translation unit that is not part of this harness, so they are re-issued here on the
   reference's own primitives: isect_grid3d_init, msh_mat4_vec3_mul, isect_grid3d_cell_from_world_space (intersect.h:57-116,
   defined in pose_proposal.cpp's TU). */
} /* extern "C" */
extern "C" {
/* res[3] = x/y/z resolution, origin[3] = padded min corner; returns n_cells */
int32_t ref_cov_grid( const float* bbox_min, const float* bbox_max, float voxel, int32_t* res, float* origin )
{
  msh_bbox_t bb; bb.min_p = msh_vec3( bbox_min[0], bbox_min[1], bbox_min[2] ); bb.max_p = msh_vec3( bbox_max[0], bbox_max[1], bbox_max[2] );
  isect_grid3d_t g; memset( &g, 0, sizeof( g ) );
  isect_grid3d_init( &g, &bb, voxel );
  res[0] = g.x_res; res[1] = g.y_res; res[2] = g.z_res;
  origin[0] = g.bbox.min_p.x; origin[1] = g.bbox.min_p.y; origin[2] = g.bbox.min_p.z;
  int32_t n = g.n_cells;
  isect_grid3d_term( &g );
  return n;
}
/* cells of the points (under `pose` when given) set to 1 in `grid` (n_cells bytes, not cleared) */
void ref_cov_rasterize( const float* bbox_min, const float* bbox_max, float voxel, const float* pts, int32_t n, const float* pose, uint8_t* grid )
{
  msh_bbox_t bb; bb.min_p = msh_vec3( bbox_min[0], bbox_min[1], bbox_min[2] ); bb.max_p = msh_vec3( bbox_max[0], bbox_max[1], bbox_max[2] );
  isect_grid3d_t g; memset( &g, 0, sizeof( g ) );
  isect_grid3d_init( &g, &bb, voxel );
  for( int32_t i = 0; i < n; ++i )
  {
    msh_vec3_t p = msh_vec3( pts[3 * i], pts[3 * i + 1], pts[3 * i + 2] );
    if( pose ) { p = msh_mat4_vec3_mul( mat_from( pose ), p, 1 ); }
    uint8_t* cell = isect_grid3d_cell_from_world_space( &g, p );
    if( cell ) { grid[cell - g.data] = 1; }
  }
  isect_grid3d_term( &g );
}

/* ---------------------------------------------------------------- NMS (pose_proposal.cpp:371-452, intersect.h:309-368) */
} /* extern "C" */
float isect_get_overlap_factor( const rs_pointcloud_t* pc_a, const msh_mat4_t pose_a, const rs_pointcloud_t* pc_b, const msh_mat4_t pose_b,
                                const float voxel_size, const int voxelize_inside, const int normalize_by_smaller ); /* defined in pose_proposal.cpp's TU */
extern "C" {
float ref_overlap_factor( void* cloud, const float* pose_a, const float* pose_b, float voxel, int inside, int norm_smaller )
{
  return isect_get_overlap_factor( (rs_pointcloud_t*)cloud, mat_from( pose_a ), (rs_pointcloud_t*)cloud, mat_from( pose_b ), voxel, inside, norm_smaller );
}
void ref_cloud_centroid( void* cloud, float* c )
{
  msh_vec3_t v = rs_pointcloud_centroid( (rs_pointcloud_t*)cloud, 0 );
  c[0] = v.x; c[1] = v.y; c[2] = v.z;
}
/* mgs_non_maxima_suppresion restricted to one object of the database: proposals in (n x 17 floats), survivors out
   (same layout, reference order); returns their number */
int32_t ref_nms( void* db, int object_idx, const float* proposals, int32_t n, float dist_threshold, float* kept )
{
  rsdb_t* rsdb = (rsdb_t*)db;
  msh_array( msh_array( pose_proposal_t ) ) pp = NULL;
  for( size_t i = 0; i < msh_array_len( rsdb->objects ); ++i )
  {
    msh_array( pose_proposal_t ) cur = NULL;
    if( (int)i == object_idx )
    {
      for( int32_t j = 0; j < n; ++j )
      {
        pose_proposal_t p; memcpy( p.xform.data, proposals + 17 * (size_t)j, 64 ); p.score = proposals[17 * (size_t)j + 16];
        msh_array_push( cur, p );
      }
    }
    msh_array_push( pp, cur );
  }
  mgs_non_maxima_suppresion( rsdb, &pp, 0, dist_threshold );
  int32_t m = (int32_t)msh_array_len( pp[object_idx] );
  for( int32_t j = 0; j < m; ++j ) { memcpy( kept + 17 * (size_t)j, pp[object_idx][j].xform.data, 64 ); kept[17 * (size_t)j + 16] = pp[object_idx][j].score; }
  for( size_t i = 0; i < msh_array_len( pp ); ++i ) { if( pp[i] ) { msh_array_free( pp[i] ); } }
  msh_array_free( pp );
  return m;
}

} /* extern "C" */
