// Optional object of both drop-in executables: level building at its call site.  Replaces
// rs_pointcloud__compute_level_poisson (reference lib/rs/rs_pointcloud.h:984-1106), which rs_pointcloud_compute_levels runs
// for levels 1-4 of every cloud at load time (:1305-1316), by rsgpu_poisson_level (the greedy Poisson-disk selection in its
// exact parallel form) followed by the reference's own row copy.  Linked in front of the reference's unmodified
// objects (integration/Makefile: pose_proposal_rsgpu_levels, segment_transfer_rsgpu_all).
//
// An input outside the exact envelope of rsgpu_poisson_level (a disk holding more than max_n_neigh points, where the
// reference marks only the nearest max_n_neigh) ends the run with the library's message: once this object is linked the
// reference's loop is no longer in the executable, and nothing here falls back to a CPU path silently.
// STATUS: host logic verified on the CPU tier against the pure-CPU build (tests/test_host_logic.py, oracle-backed stand-in);
// rsgpu_poisson_level itself is verified on the GPU (tests/test_gpu_levels.py); this glue has not run on a GPU yet.
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

#include "rsgpu.h"

// reference helpers of the same translation unit (rs_pointcloud.h:866-900), C++ linkage
void rs_pointcloud__allocate_level( rs_pointcloud_t* pc, int32_t level, int32_t n_pts );
void rs_pointcloud__free_level( rs_pointcloud_t* pc, int32_t level );

void
rs_pointcloud__compute_level_poisson( rs_pointcloud_t* pc, int32_t level )
{
  if( level <= 0 || level >= RSPC_N_LEVELS )
  {
    fprintf( stderr, "rsgpu drop-in: level %d is not built from level 0 by the Poisson-disk selection\n", level ); // the reference only calls it for 1-4 (:1313)
    exit( -1 );
  }
  static bool announced = false;
  if( !announced ) { announced = true; printf( "IO: levels 1-4 of every cloud selected by rsgpu_poisson_level (GPU)\n" ); }
  const int32_t n = (int32_t)pc->n_pts[0];
  size_t max_n_neigh = 1024 * ( ( level ) / (float)( RSPC_N_LEVELS - 1 ) ); // :994-995
  if( !max_n_neigh ) { max_n_neigh = 256; }
  std::vector<int32_t> sample_ind( n > 0 ? n : 1 );
  int32_t n_samples = 0;
  if( n > 0 && rsgpu_poisson_level( &pc->positions[0][0].x, n, pc->voxel_size[level], (int32_t)max_n_neigh, sample_ind.data(), &n_samples, NULL ) != RSGPU_OK )
  {
    fprintf( stderr, "rsgpu drop-in: rsgpu_poisson_level( level %d, %d points ) failed: %s\n", level, n, rsgpu_last_error() );
    exit( -1 );
  }
  // the level's arrays are copies of the selected level-0 rows (:1045-1086)
  const msh_vec3_t* positions = pc->positions[0];
  const msh_vec3_t* normals = pc->normals[0];
  const msh_vec3_t* colors = pc->colors[0];
  const float* radii = pc->radii[0];
  const float* qualities = pc->qualities[0];
  const int32_t* classes = pc->class_ids[0];
  const int32_t* instances = pc->instance_ids[0];
  rs_pointcloud__free_level( pc, level );
  rs_pointcloud__allocate_level( pc, level, n_samples );
  for( int32_t i = 0; i < n_samples; ++i )
  {
    const int32_t s = sample_ind[i];
    if( positions ) { pc->positions[level][i] = positions[s]; }
    if( normals ) { pc->normals[level][i] = normals[s]; }
    if( colors ) { pc->colors[level][i] = colors[s]; }
    if( radii ) { pc->radii[level][i] = radii[s]; }
    if( qualities ) { pc->qualities[level][i] = qualities[s]; }
    if( classes ) { pc->class_ids[level][i] = classes[s]; }
    if( instances ) { pc->instance_ids[level][i] = instances[s]; }
  }
}
