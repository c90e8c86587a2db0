/* rsgpu.h — C ABI of the B200 (sm_100a) implementation of Rescan's pose_proposal hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no CUDA / torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the mhalber/Rescan tree).
 * Host-buffer calls take and return exactly the layouts the reference's callers already hold
 * (AoS xyz floats, column-major 4x4 msh_mat4_t, row-major [n_query][k] result rows); the `_dev`
 * variants take device pointers of the same layouts for callers that keep data resident in HBM.
 *
 * Conventions
 *   - every function returns RSGPU_OK (0) or a negative rsgpu_status; rsgpu_last_error() gives the text.
 *     The reference itself reports nothing (assert-only) — callers that ignore the status
 *     get reference behaviour on success and untouched outputs on failure.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with RSGPU_ERR_NO_DEVICE.
 *   - handles are opaque and own device memory; destroy them with the matching *_destroy.
 *   - all work is enqueued on one stream per process (legacy default stream unless rsgpu_set_stream
 *     was called); host-buffer calls synchronise that stream before returning.  A host thread that called
 *     rsgpu_thread_attach( lane ) uses that lane's own stream instead, so independent call chains issued from
 *     different threads (one per object, say) overlap on the device.  Handles may be shared between threads
 *     for reading; the library keeps no other cross-call state.
 */
#ifndef RSGPU_H
#define RSGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rsgpu_status
{
  RSGPU_OK = 0,
  RSGPU_ERR_NO_DEVICE = -1,   /* no CUDA device / driver */
  RSGPU_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed */
  RSGPU_ERR_INVALID = -3,     /* bad argument */
  RSGPU_ERR_UNSUPPORTED = -4, /* outside the supported envelope (e.g. k > RSGPU_MAX_K) */
  RSGPU_ERR_OOM = -5
} rsgpu_status;

#define RSGPU_MAX_K 512           /* largest k / max_n_neigh of the search API */
#define RSGPU_MAX_CELLS_PER_QUERY 512 /* msh_hash_grid.h:1101 MAX_BIN_COUNT: cells examined per query */

typedef struct rsgpu_grid rsgpu_grid_t;   /* device hash grid      <-> msh_hash_grid_t   (msh_hash_grid.h:262-283) */
typedef struct rsgpu_cloud rsgpu_cloud_t; /* device point set      <-> one level of rs_pointcloud_t (rs_pointcloud.h:77-97) */

/* ------------------------------------------------------------------------------------------------ runtime */
int rsgpu_device_count( void );
int rsgpu_set_device( int device );       /* cudaSetDevice for this process; default 0 */
int rsgpu_set_stream( void* cuda_stream );/* cudaStream_t to enqueue on; NULL = legacy default stream */
int rsgpu_synchronize( void );
/* lanes: rsgpu_thread_attach( lane ), 0 <= lane < rsgpu_lane_count(), binds the CALLING host thread to the lane's
   stream (created on first use, non-blocking) and to the process's device; lane < 0 detaches.  The reference is
   single-threaded (SURVEY.md 8b "Threading"); lanes are how a caller overlaps its per-object loops
   (apps/pose_proposal/main.cpp:175-204, pose_proposal.cpp:190-250) on one GPU. */
int rsgpu_lane_count( void );
int rsgpu_thread_attach( int lane );
const char* rsgpu_last_error( void );
const char* rsgpu_version( void );

/* Per-kernel device timing with CUDA events on the launch stream (used by bench.py for the roofline line).
   names: "grid_build", "search", "score_dense" (pose-grid launches), "score" (explicit pose lists), "icp",
   "labels", "unary", "edges".  ms = summed event time. */
/* tuning / A-B knobs ("search_impl" = lane|warp, "score_impl" = coop, "prune" = 0, "icp_impl" = block, "icp_sums" = fp64,
   ...); value NULL or "" restores the default.  The same knobs are read from the environment as RSGPU_<NAME>. */
int rsgpu_set_option( const char* name, const char* value );
int rsgpu_profile_enable( int on );
int rsgpu_profile_reset( void );
int rsgpu_profile_get( const char* name, double* ms, int64_t* launches );
/* number of this library's own kernels launched so far by this process (CUB / memcpy / memset not counted) */
int64_t rsgpu_launch_count( void );

/* ------------------------------------------------------------------------------------------------ hash grid */
/* replaces msh_hash_grid_init_3d (msh_hash_grid.h:222, impl :388-541, 550-555).  `pts` = n_pts x {x,y,z}
   floats; radius > 0 gives cell = 2*radius, radius <= 0 the reference's automatic cell size. */
int rsgpu_grid_create( const float* pts, int32_t n_pts, float radius, rsgpu_grid_t** out );
int rsgpu_grid_create_dev( const float* d_pts, int32_t n_pts, float radius, rsgpu_grid_t** out );
/* replaces msh_hash_grid_term (msh_hash_grid.h:225) */
void rsgpu_grid_destroy( rsgpu_grid_t* grid );

/* Per-point unit normals in ORIGINAL point order (the scan level's normals the reference indexes with the
   returned point index: pose_proposal.cpp:137, icp.h:372); stored re-laid in cell order next to the points. */
int rsgpu_grid_set_normals( rsgpu_grid_t* grid, const float* normals );
int rsgpu_grid_set_normals_dev( rsgpu_grid_t* grid, const float* d_normals );

typedef struct rsgpu_grid_info
{
  int64_t width, height, depth;   /* msh_hash_grid_t::width/height/depth */
  double cell_size, inv_cell_size;
  float min_pt[3], max_pt[3];
  int64_t n_pts;
  int64_t n_bins;                 /* non-empty cells */
  int64_t max_n_pts_in_bin;
} rsgpu_grid_info_t;
int rsgpu_grid_get_info( const rsgpu_grid_t* grid, rsgpu_grid_info_t* info );
/* the re-laid point records = msh_hash_grid_t::data_buffer (msh_hash_grid.h:238-242, order :501-532) */
int rsgpu_grid_get_data( const rsgpu_grid_t* grid, float* xyz, int32_t* idx );

/* field-for-field msh_hash_grid_search_desc_t (msh_hash_grid.h:196-216) */
typedef struct rsgpu_search_desc
{
  float* query_pts;       /* n_query_pts x {x,y,z} */
  size_t n_query_pts;
  float* distances_sq;    /* [n_query_pts][k] */
  int32_t* indices;       /* [n_query_pts][k] */
  size_t* n_neighbors;    /* [n_query_pts], may be NULL */
  float radius;
  union { size_t k; size_t max_n_neigh; };
  int sort;               /* rows are always written ascending, which satisfies both sort = 0 and 1 */
} rsgpu_search_desc_t;

/* replace msh_hash_grid_radius_search / _knn_search (msh_hash_grid.h:227-230; impl :1090-1259, :1294-1450).
   *total = the reference's return value (sum of per-query counts). */
int rsgpu_grid_radius_search( const rsgpu_grid_t* grid, rsgpu_search_desc_t* desc, size_t* total );
int rsgpu_grid_knn_search( const rsgpu_grid_t* grid, rsgpu_search_desc_t* desc, size_t* total );
/* same with every buffer of `desc` in device memory (n_neighbors then is uint64 on the device) */
int rsgpu_grid_radius_search_dev( const rsgpu_grid_t* grid, rsgpu_search_desc_t* desc, size_t* total );
int rsgpu_grid_knn_search_dev( const rsgpu_grid_t* grid, rsgpu_search_desc_t* desc, size_t* total );
/* measurement aid: what msh_hash_grid_radius_search reads for these (device-resident) queries by the reference's
   data layout — counts[0] = non-empty cells overlapping the query windows, counts[1] = points stored in them
   (msh_hash_grid.h:1187-1225, :826-862); no early-out credit */
int rsgpu_grid_search_census_dev( const rsgpu_grid_t* grid, const float* d_query_pts, size_t n_query_pts, float radius, int64_t counts[2] );

/* ------------------------------------------------------------------------------------------------ clouds */
/* positions + normals of one sampling level of an object model (rs_pointcloud_t::positions[lvl] / normals[lvl]) */
int rsgpu_cloud_create( const float* pos, const float* nor, int32_t n_pts, rsgpu_cloud_t** out );
void rsgpu_cloud_destroy( rsgpu_cloud_t* cloud );
int32_t rsgpu_cloud_size( const rsgpu_cloud_t* cloud );

/* ------------------------------------------------------------------------------------------------ pose scoring */
/* replaces mgs_compute_object_alignment_score (pose_proposal.h:46-49, impl pose_proposal.cpp:93-158) for a
   batch of poses.  `scene` is the scan's level-`search_lvl` grid WITH normals set; radius = sigma =
   search_radii[search_lvl] (0.10 for level 1, pose_proposal.cpp:98); max_n_neigh = storage->max_n_neigh.
   xforms: n_poses x 16 floats, column-major msh_mat4_t.  scores: n_poses floats. */
int rsgpu_score_poses( const rsgpu_cloud_t* object, const rsgpu_grid_t* scene, const float* xforms,
                       int64_t n_poses, int32_t max_n_neigh, float radius, float* scores );
int rsgpu_score_poses_dev( const rsgpu_cloud_t* object, const rsgpu_grid_t* scene, const float* d_xforms,
                           int64_t n_poses, int32_t max_n_neigh, float radius, float* d_scores );

/* Dense pose grid = rotations x translations, pose (t, r) = rotation r with its translation column replaced
   by translation t (pose_proposal.cpp:221-222).  rotations: n_rot x 16 floats (column-major);
   translations: n_trans x 3 floats; scores: [n_trans][n_rot]. */
int rsgpu_score_pose_grid( const rsgpu_cloud_t* object, const rsgpu_grid_t* scene, const float* rotations,
                           int32_t n_rot, const float* translations, int64_t n_trans, int32_t max_n_neigh,
                           float radius, float* scores );
int rsgpu_score_pose_grid_dev( const rsgpu_cloud_t* object, const rsgpu_grid_t* scene, const float* d_rotations,
                               int32_t n_rot, const float* d_translations, int64_t n_trans,
                               int32_t max_n_neigh, float radius, float* d_scores );

/* Algorithmic-byte accounting of one scoring call: per query 8*B (non-empty cells overlapping
   the +-radius box) + 16*C (points stored in them) + 12*T (scene normals the reference fetches), per pose
   +68.  Counting pass only (slow, exact); results in counts[4] = {queries, B, C, T}. */
int rsgpu_score_pose_grid_count( const rsgpu_cloud_t* object, const rsgpu_grid_t* scene, const float* rotations,
                                 int32_t n_rot, const float* translations, int64_t n_trans,
                                 int32_t max_n_neigh, float radius, int64_t counts[4] );

/* pose_proposal_t (pose_proposal.h:6-10): 16 floats column-major xform + score = 17 floats */
#define RSGPU_POSE_FLOATS 17

typedef struct rsgpu_propose_opts
{
  int32_t max_n_neigh;      /* 64  (pose_proposal.cpp:179) */
  float radius;             /* 0.10 = search_radii[search_lvl = 1] (pose_proposal.cpp:98, 178) */
  float thresholds[3];      /* 0.25, 0.35, 0.40 for object levels 4, 3, 2 (pose_proposal.cpp:160-168) */
  int32_t top_k;            /* <= 0: keep every survivor in translation order (reference behaviour);
                               > 0: keep the top_k best per object, descending score, ties by pose id */
  const int64_t* translation_ids; /* NULL, or n_trans distinct ids: translation j of the call IS translation
                               translation_ids[j] of the caller's list.  Pose ids, the order of the survivors and every tie
                               then follow the caller's numbering, whatever order the translations are passed in - a caller
                               may hand them over sorted along a space-filling curve (neighbouring launches then touch
                               neighbouring scan cells: -9 % on the dense search) without changing a single output bit. */
} rsgpu_propose_opts_t;
void rsgpu_propose_default_opts( rsgpu_propose_opts_t* opts );

/* replaces mgs_propose_poses for ONE object (pose_proposal.cpp:325-369 = levels 4 -> 3 -> 2:
   mgs__initial_pose_proposals :170-254 then mgs__pose_verification :256-303 twice).
   object_lvl4/3/2: the object's level-4/3/2 point sets.  On return *n_out proposals are written to
   `out` (capacity out_cap x 17 floats) and their dense pose ids t*n_rot + r to `out_pose_id` (nullable).
   Survivors include verification failures with score -1, as in the reference (:348-359). */
int rsgpu_propose_poses( const rsgpu_cloud_t* object_lvl4, const rsgpu_cloud_t* object_lvl3,
                         const rsgpu_cloud_t* object_lvl2, const rsgpu_grid_t* scene, const float* rotations,
                         int32_t n_rot, const float* translations, int64_t n_trans,
                         const rsgpu_propose_opts_t* opts, float* out, int64_t* out_pose_id, int64_t out_cap,
                         int64_t* n_out );

/* ------------------------------------------------------------------------------------------------ ICP */
/* replaces icp_align (icp.h:84-88, impl :416-500) for a batch of starting poses of one object against one
   scan level.  `object` = pts1/nor1, `scan` = grid over pts2 WITH normals (any cell size: the search is
   exact, so results do not depend on it).  T1: n_batch x 16 floats column-major, updated in place;
   T2: 16 floats (the reference's callers pass identity).  errs: the return values; iters (nullable):
   estimation steps taken. */
int rsgpu_icp_align_batch( const rsgpu_cloud_t* object, const rsgpu_grid_t* scan, float* T1, int32_t n_batch,
                           const float* T2, float max_dist, float max_angle, float* errs, int32_t* iters );
/* several objects against the same scan in ONE launch (one thread block per starting pose), e.g. all the
   proposals main.cpp:175-204 refines for one scan */
typedef struct rsgpu_icp_job
{
  const rsgpu_cloud_t* object; /* pts1 / nor1 */
  float* T1;                   /* n_batch x 16, in/out */
  int32_t n_batch;
  float* errs;                 /* n_batch */
  int32_t* iters;              /* n_batch, may be NULL */
} rsgpu_icp_job_t;
int rsgpu_icp_align_multi( const rsgpu_icp_job_t* jobs, int32_t n_jobs, const rsgpu_grid_t* scan, const float* T2,
                           float max_dist, float max_angle );
/* same with the iteration cap exposed (the reference hard-codes max_iter = 100, icp.h:443); <= 0 means 100 */
int rsgpu_icp_align_batch_ex( const rsgpu_cloud_t* object, const rsgpu_grid_t* scan, float* T1, int32_t n_batch,
                              const float* T2, float max_dist, float max_angle, int32_t max_iter, float* errs,
                              int32_t* iters );

/* ------------------------------------------------------------------------------------------------ labels / unary terms */
/* replaces rspf__assign_temporary_labels (rs_pointcloud_filters.cpp:738-778) over placements
   [first, last): scan level-1 positions/normals (n_vertices), one pose + object level-1 grid (with normals)
   per placement.  labels (int8, value = placement index + 1) and min_dists (float, squared) are read and
   updated in place, exactly like the reference's running arg-min. */
int rsgpu_assign_labels( const float* scan_pos, const float* scan_nor, int32_t n_vertices, const float* poses,
                         const rsgpu_grid_t* const* object_grids, int32_t first, int32_t last, float radius,
                         int8_t* labels, float* min_dists );

/* replaces the data_cost block of rspf_smooth_labels (rs_pointcloud_filters.cpp:926-939):
   cost[v*n_labels + l] = (l == labels[v]) ? 0 : c, c = 30, 15 when label_is_static[labels[v]], 1 when label 0. */
int rsgpu_unary_costs( const int32_t* labels, const uint8_t* label_is_static, int32_t n_vertices,
                       int32_t n_labels, int32_t* data_cost );

/* candidate edges of rspf_compute_neighborhood (rs_pointcloud_filters.cpp:674-722) before its hashtable
   de-duplication: per vertex its <= max_nn nearest neighbours within sqrt(radius_sq) (self included) and
   weight (1-(d2/(4 r^2))^dist_exp) * clamp(n.m,0,1)^angle_exp.  neighbors/weights: [n][max_nn], -1 / 0 where absent.
   `grid` = the cloud's own grid with normals set. */
int rsgpu_neighborhood( const rsgpu_grid_t* grid, const float* pos, const float* nor, int32_t n_vertices,
                        int32_t max_nn, float radius_sq, float dist_exp, float angle_exp, int32_t* neighbors,
                        float* weights );

/* ---- non-maxima suppression of one object's proposals ------------------------------------------
   rsgpu_overlap_factors replaces isect_get_overlap_factor (lib/rs/intersect.h:309-368) for ONE cloud under a reference
   pose against n other poses (column-major 4x4 each): lvl3 = the object's level-3 cloud (posed bounding boxes,
   :119-130), lvl1 = its level-1 cloud (occupancy, :177-306).  out[i] is bit-identical to the reference's float.
   rsgpu_nms replaces mgs_non_maxima_suppresion for one object (apps/pose_proposal/pose_proposal.cpp:371-452):
   proposals = n x RSGPU_POSE_FLOATS, centroid = rs_pointcloud_centroid( shape, 0 ) (computed by the caller),
   keep[i] = 1 for the survivors (the caller copies them in their original order, :440-447). */
int rsgpu_overlap_factors( const rsgpu_cloud_t* lvl3, const rsgpu_cloud_t* lvl1, const float* pose_ref, const float* poses, int32_t n,
                           float voxel_size, int32_t voxelize_inside, int32_t normalize_by_smaller, float* out );
int rsgpu_nms( const rsgpu_cloud_t* lvl3, const rsgpu_cloud_t* lvl1, const float centroid[3], const float* proposals, int32_t n,
               float dist_threshold, uint8_t* keep );

/* ---- level building ---------------------------------------------------------------------------
   rsgpu_poisson_level replaces the sampling loop of rs_pointcloud__compute_level_poisson (lib/rs/rs_pointcloud.h:984-1037):
   pts = the n level-0 points, voxel = pc->voxel_size[level] (the disk radius), max_n_neigh = the k of the reference's
   search, (size_t)(1024 * (level / 4.0f)) or 256 when that is 0 (:994-995).  out_indices (room for n) receives the
   ascending level-0 indices of the samples - the level's arrays are copies of those rows (:1077-1086) - and *n_out
   their number; *n_rounds (nullable) the number of propagation rounds.  The result equals the reference's sequential
   greedy selection exactly as long as no sample has more than max_n_neigh points inside its disk (the reference then
   marks only the nearest max_n_neigh); such an input fails with RSGPU_ERR_UNSUPPORTED. */
int rsgpu_poisson_level( const float* pts, int32_t n, float voxel, int32_t max_n_neigh, int32_t* out_indices, int32_t* n_out,
                         int32_t* n_rounds );

/* ---- coverage term of the arrangement optimiser -------------------------------------------------
   Grids are isect_grid3d_t's (lib/rs/intersect.h:20-30, 57-116): `origin` = the padded bbox min corner, `res` = x/y/z
   resolution (isect_grid3d_init :57-75), cell of a point = floorf( (p - origin) * (1.0f / voxel) ) per axis, stored at
   (y * res[2] + z) * res[0] + x, one byte per cell.
   rsgpu_rasterize_points replaces the loop of rsao_rasterize_scene_to_grid (apps/segment_transfer/
   arrangement_optimization.cpp:1064-1080; the caller applies its quality filter to the points first): cells of the n
   points - moved by `pose` (column-major 4x4) when not NULL - are set to 1 in `grid`, which is not cleared.
   rsgpu_coverage_masks replaces rsao__rasterize_arrangement_to_grid (:1082-1106) + the counting loop of
   rsao__compute_scene_coverage_score (:343-373) for a whole LIST of candidate placements: placement i = objects[i] (its
   level-2 points) under poses[16 i ..].  *n_lit = number of lit cells of scene_grid; out_masks[i * n_words + w], n_words >=
   ceil( *n_lit / 32 ), holds one bit per lit scan cell (bit b = the b-th lit cell in ascending cell index), set iff the
   placement lights that cell.  The reference's coverage score of any arrangement is then
   popcount( OR of its placements' masks ) / *n_lit (0 when *n_lit is 0).  out_masks == NULL only counts the lit cells. */
int rsgpu_rasterize_points( const float* pts, int32_t n, const float* pose, const float origin[3], const int32_t res[3], float voxel,
                            uint8_t* grid );
int rsgpu_coverage_masks( const rsgpu_cloud_t* const* objects, const float* poses, int32_t n_poses, const float origin[3],
                          const int32_t res[3], float voxel, const uint8_t* scene_grid, uint32_t* out_masks, int32_t n_words,
                          int32_t* n_lit );

/* ---- plane detection of the scan (the stage that bounds segment_transfer once SURVEY.md 8 is on the device) ---------
   rsgpu_plane_inlier_counts replaces evaluate_plane_model (lib/rs/rs_pointcloud_filters.cpp:117-134) for a whole LIST of
   candidate planes, i.e. one RANSAC round of rspf__detect_walls / rspf__detect_floor (:137-253): planes = n_planes x
   {center xyz, normal xyz}; active[i] != 0 <-> weights[i] > 0.01 (points no earlier plane explained); counts[p] = number of
   active points with |normal . (pt - center)| < dist_threshold in the reference's float arithmetic.  The caller keeps the
   reference's choice (first candidate with the strictly largest count) and its remove_inliers pass. */
int rsgpu_plane_inlier_counts( const float* pts, const uint8_t* active, int32_t n_pts, const float* planes, int32_t n_planes, float dist_threshold,
                               int32_t* counts );

/* ---- multi-GPU exchange of the pose-sharded search ---------------------------------------------------
   The reference's loop over translations (apps/pose_proposal/pose_proposal.cpp:213-243) is split over one process per GPU;
   what the ranks exchange is the survivor list of :348-359 (pose_proposal_t records + pose ids) and, after the refinement
   of apps/pose_proposal/main.cpp:175-204, the refined rows.  rsgpu_peer_* is that all-gather over NVLink without an SM-resident
   collective kernel: every rank's receive area is mapped by all peers (CUDA IPC), payloads and round flags are written by
   copy engines, one warp waits for the flags.
     rsgpu_peer_init      allocates this rank's area: n_slots independent exchange slots (one per object chain in flight; each
                          double-buffered), slot_bytes = largest payload of one rank in one exchange; returns the area's IPC
                          handle (rsgpu_peer_handle_bytes() bytes) for the caller to distribute (set-up, any transport);
     rsgpu_peer_open      handles = world x rsgpu_peer_handle_bytes() bytes, rank-major;
     rsgpu_peer_allgather all-gather on slot `slot`: `nbytes` (equal on every rank) from host `send` -> host `recv`
                          [world][nbytes].  seq = 1, 2, 3, ... counts the uses of THIS slot (identical on every rank).  Exchanges
                          on different slots are independent: they may be issued from different host threads in any order,
                          every rank in its own order (no global collective order).  Runs on the calling thread's lane stream.
                          timeout_s <= 0 means 30 s; a peer that never sends gives RSGPU_ERR_CUDA, not a hang;
     rsgpu_peer_close     every rank must have left its last rsgpu_peer_allgather (caller barrier) before any rank closes. */
int rsgpu_peer_handle_bytes( void );
int rsgpu_peer_init( int32_t rank, int32_t world, int32_t n_slots, int64_t slot_bytes, void* handle_out );
int rsgpu_peer_open( const void* handles );
int rsgpu_peer_allgather( int32_t slot, uint32_t seq, const void* send, int64_t nbytes, void* recv, double timeout_s );
/* the two halves of an exchange, for rooted patterns (gather to the rank that owns an object's chain, broadcast of its result):
     rsgpu_peer_put  use `seq` of `slot`: `nbytes` (may differ between ranks) from host `send` into this rank's row of the slot
                     on every rank whose bit is set in dst_mask; returns when the payload has landed - it never waits for a peer;
     rsgpu_peer_get  waits until every rank in src_mask has put use `seq` of `slot`, then copies the rows to host `recv`
                     [world][row_bytes] and their lengths to nbytes_out[world] (0 for ranks outside the mask).
   Every rank in a use's src_mask must be a sender of that use to this rank; a slot's uses are numbered identically on all ranks. */
int rsgpu_peer_put( int32_t slot, uint32_t seq, uint32_t dst_mask, const void* send, int64_t nbytes );
int rsgpu_peer_get( int32_t slot, uint32_t seq, uint32_t src_mask, void* recv, int64_t row_bytes, int64_t* nbytes_out, double timeout_s );
int rsgpu_peer_close( void );

#ifdef __cplusplus
}
#endif
#endif /* RSGPU_H */
