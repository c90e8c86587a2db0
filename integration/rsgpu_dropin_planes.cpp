// Optional second object of the segment_transfer drop-in: the RANSAC rounds of the scan's wall / floor detector with all
// candidates of a round counted in ONE rsgpu_plane_inlier_counts call (include/rsgpu.h) instead of one
// evaluate_plane_model pass per candidate.  Linked in front of the reference's unmodified objects like
// integration/rsgpu_dropin_st.cpp (integration/Makefile: segment_transfer_rsgpu_planes); everything else of
// rspf_detect_planes (inlier gathering, connected components, refinement) stays reference code.
//
//   replaces                  (reference)                                  with
//   rspf__detect_floor        lib/rs/rs_pointcloud_filters.cpp:205-253     2 500 triples drawn with the reference's own sampler -> one count launch
//   rspf__detect_walls        lib/rs/rs_pointcloud_filters.cpp:137-203     per wall: 5 000 triples -> one count launch, first strict maximum, remove_inliers
//
// The triples never depend on the counts (the sampler is advanced the same number of times whatever a candidate scores), so
// drawing them first and counting afterwards visits the same candidates in the same order; counts are integers computed with
// the reference's float expression, so the chosen planes are the reference's.
// STATUS: host logic verified on the CPU tier against the pure-CPU build (tests/test_host_logic.py, oracle-backed stand-in);
// the CUDA entry point matches the reference's golden vectors and the oracle on a B200 (tests/test_gpu_zplanes.py); this
// object itself has not run on a GPU yet.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <vector>

#include "msh/msh_std.h"
#include "msh/msh_vec_math.h"
#include "msh/msh_geometry.h"
#include "msh/msh_hash_grid.h"
#include "mg/hashtable.h"
#include "rs_pointcloud.h"
#include "rs_distance_function.h"
#include "rs_database.h"
#include "rs_pointcloud_filters.h"

#include "rsgpu.h"

// the detector's parameter block is private to rs_pointcloud_filters.cpp (:82-92); the two replaced functions take it by
// pointer, so the same tag and layout are declared here
typedef struct rspf_detector_params
{
  msh_vec3_t* pts;
  msh_vec3_t* nrmls;
  size_t n_pts;
  float dot_threshold;
  float dist_threshold;
  size_t count_threshold;
  bool check_validity;
  bool check_extends;
} rspf_detector_params_t;

void remove_inliers( rspf_plane_model_t* model, double* weights, msh_vec3_t* pts, size_t n_pts, float dist_threshold ); // reference, :96-114

namespace
{
void die( const char* what )
{
  fprintf( stderr, "rsgpu drop-in: %s failed: %s\n", what, rsgpu_last_error() );
  exit( -1 );
}
#define RSGPU_OR_DIE( call ) do { if( ( call ) != RSGPU_OK ) { die( #call ); } } while( 0 )

struct Round
{
  std::vector<float> planes;   // candidates that are counted: {center, normal}
  std::vector<int32_t> counts;
  void clear() { planes.clear(); }
  void add( msh_vec3_t c, msh_vec3_t n )
  {
    const float v[6] = { c.x, c.y, c.z, n.x, n.y, n.z };
    planes.insert( planes.end(), v, v + 6 );
  }
  // counts of all candidates over the points with weight > 0.01 (evaluate_plane_model's test, :127)
  void count( const rspf_detector_params_t* params, const double* weights )
  {
    std::vector<uint8_t> active( params->n_pts );
    for( size_t i = 0; i < params->n_pts; ++i ) { active[i] = weights[i] > 0.01 ? 1 : 0; }
    const int32_t n = (int32_t)( planes.size() / 6 );
    counts.assign( n, 0 );
    if( n == 0 ) { return; }
    RSGPU_OR_DIE( rsgpu_plane_inlier_counts( &params->pts[0].x, active.data(), (int32_t)params->n_pts, planes.data(), n, params->dist_threshold, counts.data() ) );
  }
};

msh_vec3_t triple_normal( msh_vec3_t p_a, msh_vec3_t p_b, msh_vec3_t p_c )
{
  const msh_vec3_t v_a = msh_vec3_sub( p_b, p_a );
  const msh_vec3_t v_b = msh_vec3_sub( p_c, p_a );
  return msh_vec3_normalize( msh_vec3_cross( v_a, v_b ) );
}
} // namespace

int32_t
rspf__detect_floor( const rspf_detector_params_t* params, msh_array( rspf_plane_model_t ) * models )
{
  double* weights = (double*)malloc( params->n_pts * sizeof( double ) );
  const msh_vec3_t up = msh_vec3_posy();
  for( size_t i = 0; i < params->n_pts; ++i )
  {
    const float dot = msh_vec3_dot( params->nrmls[i], up );
    weights[i] = dot > params->dot_threshold ? 1.0 : 0.0;
  }
  msh_discrete_distrib_t dist = { 0 };
  msh_discrete_distribution_init( &dist, weights, params->n_pts, 12346ULL );
  const uint32_t max_ransac_iter = 2500;
  Round round;
  for( uint32_t i = 0; i < max_ransac_iter; ++i )
  {
    const int32_t idx_a = msh_discrete_distribution_sample( &dist );
    const int32_t idx_b = msh_discrete_distribution_sample( &dist );
    const int32_t idx_c = msh_discrete_distribution_sample( &dist );
    round.add( params->pts[idx_a], triple_normal( params->pts[idx_a], params->pts[idx_b], params->pts[idx_c] ) );
  }
  round.count( params, weights );
  rspf_plane_model_t best = { 0 };
  int32_t floor_count = 0;
  for( uint32_t i = 0; i < max_ransac_iter; ++i )
  {
    if( (size_t)round.counts[i] > best.n_inliers ) // strict: the first of equal candidates stays (:242)
    {
      best.plane.center = msh_vec3( round.planes[6 * i], round.planes[6 * i + 1], round.planes[6 * i + 2] );
      best.plane.normal = msh_vec3( round.planes[6 * i + 3], round.planes[6 * i + 4], round.planes[6 * i + 5] );
      best.n_inliers = (size_t)round.counts[i];
      floor_count = 1;
    }
  }
  free( weights );
  msh_discrete_distribution_free( &dist );
  if( floor_count ) { msh_array_push( ( *models ), best ); }
  return floor_count;
}

int32_t
rspf__detect_walls( const rspf_detector_params_t* params, msh_array( rspf_plane_model_t ) * models )
{
  double* weights = (double*)malloc( params->n_pts * sizeof( double ) );
  const msh_vec3_t up = msh_vec3_posy();
  for( size_t i = 0; i < params->n_pts; ++i )
  {
    const float dot = msh_vec3_dot( params->nrmls[i], up );
    weights[i] = msh_abs( dot ) < ( 1 - params->dot_threshold ) ? 1.0 : 0.0;
  }
  int32_t wall_count = 0;
  const uint32_t max_ransac_iter = 5000;
  rspf_plane_model_t best = { 0 };
  Round round;
  do
  {
    msh_discrete_distrib_t dist = { 0 };
    msh_discrete_distribution_init( &dist, weights, params->n_pts, 12346ULL );
    best.n_inliers = 0; // the reference resets only the count: a round without a better candidate keeps the previous plane (:161)
    int32_t wall_detected = 0;
    round.clear();
    for( uint32_t i = 0; i < max_ransac_iter; ++i )
    {
      int32_t idx_a, idx_b, idx_c;
      idx_a = msh_discrete_distribution_sample( &dist );
      do { idx_b = msh_discrete_distribution_sample( &dist ); } while( idx_a == idx_b );
      do { idx_c = msh_discrete_distribution_sample( &dist ); } while( idx_b == idx_c );
      const msh_vec3_t n = triple_normal( params->pts[idx_a], params->pts[idx_b], params->pts[idx_c] );
      // only near-vertical candidates are counted (:179); the others cannot win
      if( msh_abs( msh_vec3_dot( n, up ) ) < ( 1 - params->dot_threshold ) ) { round.add( params->pts[idx_a], n ); }
    }
    round.count( params, weights );
    for( size_t i = 0; i < round.counts.size(); ++i )
    {
      if( (size_t)round.counts[i] > best.n_inliers )
      {
        best.plane.center = msh_vec3( round.planes[6 * i], round.planes[6 * i + 1], round.planes[6 * i + 2] );
        best.plane.normal = msh_vec3( round.planes[6 * i + 3], round.planes[6 * i + 4], round.planes[6 * i + 5] );
        best.n_inliers = (size_t)round.counts[i];
        wall_detected = 1;
      }
    }
    msh_discrete_distribution_free( &dist );
    if( wall_detected ) { msh_array_push( ( *models ), best ); }
    remove_inliers( &best, weights, params->pts, params->n_pts, params->dist_threshold );
    wall_count++;
  } while( best.n_inliers > params->count_threshold );
  msh_array_pop( ( *models ) );
  wall_count--;
  free( weights );
  printf( "RSPF_PLANE_DETECTOR: %d wall rounds of %u candidates counted in one call each (GPU)\n", wall_count + 1, max_ransac_iter );
  return wall_count;
}
