/* TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of Rescan's pose_proposal hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker.  The product (rescan_b200/csrc, librsgpu.so) never links or calls it.
 *
 * Parity status: PINNED.  The reference has no tests or golden vectors of its own, so this
 * restatement is pinned differentially against the unmodified reference compiled in place
 * (oracle/_ref/librescan_ref.so, built by oracle/Makefile from /root/reference): tests/test_oracle_vs_ref.py
 * when that library is present, and the committed fixtures under tests/golden/ (written from the same
 * library by tests/golden/make_golden.py) everywhere else.
 *
 * This is a restatement of WHAT the reference computes, written from its observable semantics, not a copy
 * of how it computes it: the open-addressing cell map, the heap/quick-sort k-list machinery and the
 * stretchy buffers of the reference are replaced by a counting sort, a sorted unique-cell table and a
 * bounded insertion list.  What is kept bit-for-bit is the arithmetic the results depend on (operand types,
 * evaluation order, no FMA contraction: build with -ffp-contract=off), each place citing the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(_OPENMP)
#include <omp.h>
#endif

#define ORC_MAX_CELLS_PER_QUERY 512 /* msh_hash_grid.h:1101 MAX_BIN_COUNT */

typedef struct orc_grid
{
  int64_t dim[3];      /* width, height, depth  (msh_hash_grid.h:446-448) */
  double cell, inv_cell;
  float mn[3], mx[3];
  int32_t n_pts;
  float* xyz;          /* n_pts*3, re-laid by ascending cell id, ascending original index inside a cell */
  int32_t* idx;        /* original index of each re-laid point */
  int64_t n_cells;     /* non-empty cells */
  int64_t* cell_key;   /* ascending linear ids of the non-empty cells */
  int32_t* cell_start; /* n_cells+1 */
} orc_grid_t;

/* ------------------------------------------------------------------------------------------------ grid */

static int cmp_i64( const void* a, const void* b )
{
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return ( x > y ) - ( x < y );
}

/* msh_hash_grid__init, dim = 3 (msh_hash_grid.h:388-541) */
orc_grid_t* orc_grid_build( const float* pts, int32_t n, float radius )
{
  orc_grid_t* g = (orc_grid_t*)calloc( 1, sizeof( orc_grid_t ) );
  if( n < 0 ) { n = 0; }
  /* bbox starts at +-1e9 stored as float and is padded by 1e-4f (:413-434) */
  for( int a = 0; a < 3; ++a ) { g->mn[a] = 1e9; g->mx[a] = -1e9; }
  for( int32_t i = 0; i < n; ++i )
    for( int a = 0; a < 3; ++a )
    {
      float v = pts[3 * i + a];
      if( g->mn[a] > v ) { g->mn[a] = v; }
      if( g->mx[a] < v ) { g->mx[a] = v; }
    }
  float ext[3], max_ext = 0;
  for( int a = 0; a < 3; ++a )
  {
    g->mx[a] += 0.0001f; g->mn[a] -= 0.0001f;
    ext[a] = g->mx[a] - g->mn[a];
  }
  max_ext = ext[0] > ext[1] ? ext[0] : ext[1];
  max_ext = max_ext > ext[2] ? max_ext : ext[2];
  /* cell = 2*radius in double, or max_dim / (32*sqrtf(3)) evaluated in float (:443-444) */
  if( radius > 0.0 ) { g->cell = 2.0 * radius; }
  else               { g->cell = max_ext / ( 32 * sqrtf( 3.0f ) ); }
  for( int a = 0; a < 3; ++a ) { g->dim[a] = (int)( ext[a] / g->cell + 1.0 ); }
  g->inv_cell = 1.0f / g->cell;
  int32_t slab = (int32_t)( g->dim[1] * g->dim[0] ); /* _slab_size is int32 (:262, 450) */
  g->n_pts = n;

  /* cell id of every point: float subtraction, then times the double inverse (:471-475) */
  int64_t* key = (int64_t*)malloc( sizeof( int64_t ) * ( n ? n : 1 ) );
  for( int32_t i = 0; i < n; ++i )
  {
    uint64_t c[3];
    for( int a = 0; a < 3; ++a ) { c[a] = (uint64_t)( ( pts[3 * i + a] - g->mn[a] ) * g->inv_cell ); }
    key[i] = (int64_t)( c[2] * (uint64_t)(int64_t)slab + c[1] * (uint64_t)g->dim[0] + c[0] );
  }
  /* unique sorted cell ids */
  int64_t* sorted = (int64_t*)malloc( sizeof( int64_t ) * ( n ? n : 1 ) );
  memcpy( sorted, key, sizeof( int64_t ) * n );
  qsort( sorted, n, sizeof( int64_t ), cmp_i64 );
  int64_t nc = 0;
  for( int32_t i = 0; i < n; ++i ) { if( i == 0 || sorted[i] != sorted[i - 1] ) { sorted[nc++] = sorted[i]; } }
  g->n_cells = nc;
  g->cell_key = (int64_t*)malloc( sizeof( int64_t ) * ( nc ? nc : 1 ) );
  memcpy( g->cell_key, sorted, sizeof( int64_t ) * nc );
  free( sorted );
  /* counting sort: bins in ascending cell id, insertion (= original index) order inside (:501-532) */
  g->cell_start = (int32_t*)calloc( nc + 2, sizeof( int32_t ) );
  int32_t* slot = (int32_t*)malloc( sizeof( int32_t ) * ( n ? n : 1 ) );
  for( int32_t i = 0; i < n; ++i )
  {
    int64_t* hit = (int64_t*)bsearch( &key[i], g->cell_key, nc, sizeof( int64_t ), cmp_i64 );
    slot[i] = (int32_t)( hit - g->cell_key );
    g->cell_start[slot[i] + 1]++;
  }
  for( int64_t c = 0; c < nc; ++c ) { g->cell_start[c + 1] += g->cell_start[c]; }
  int32_t* fill = (int32_t*)malloc( sizeof( int32_t ) * ( nc ? nc : 1 ) );
  memcpy( fill, g->cell_start, sizeof( int32_t ) * nc );
  g->xyz = (float*)malloc( sizeof( float ) * 3 * ( n ? n : 1 ) );
  g->idx = (int32_t*)malloc( sizeof( int32_t ) * ( n ? n : 1 ) );
  for( int32_t i = 0; i < n; ++i )
  {
    int32_t w = fill[slot[i]]++;
    memcpy( g->xyz + 3 * w, pts + 3 * i, 12 );
    g->idx[w] = i;
  }
  free( fill ); free( slot ); free( key );
  return g;
}

void orc_grid_free( orc_grid_t* g )
{
  if( !g ) { return; }
  free( g->xyz ); free( g->idx ); free( g->cell_key ); free( g->cell_start ); free( g );
}

void orc_grid_info( const orc_grid_t* g, int64_t* dims, double* cell, float* minmax, int64_t* counts )
{
  memcpy( dims, g->dim, 24 );
  cell[0] = g->cell; cell[1] = g->inv_cell;
  memcpy( minmax, g->mn, 12 ); memcpy( minmax + 3, g->mx, 12 );
  int32_t mb = 0;
  for( int64_t c = 0; c < g->n_cells; ++c ) { int32_t l = g->cell_start[c + 1] - g->cell_start[c]; if( l > mb ) { mb = l; } }
  counts[0] = g->n_pts; counts[1] = g->n_cells; counts[2] = mb;
}
void orc_grid_data( const orc_grid_t* g, float* xyz, int32_t* idx )
{
  memcpy( xyz, g->xyz, 12 * (size_t)g->n_pts ); memcpy( idx, g->idx, 4 * (size_t)g->n_pts );
}

static int64_t orc_find_cell( const orc_grid_t* g, int64_t key )
{
  int64_t lo = 0, hi = g->n_cells - 1;
  while( lo <= hi )
  {
    int64_t mid = ( lo + hi ) >> 1;
    if( g->cell_key[mid] == key ) { return mid; }
    if( g->cell_key[mid] < key ) { lo = mid + 1; } else { hi = mid - 1; }
  }
  return -1;
}

/* bounded best-k list ordered by (d2, arrival); a later arrival never displaces an equal distance,
   which is what the reference's `dist >= max_dist -> return` does once full (:800) */
typedef struct { float* d2; int32_t* id; size_t cap, len; } orc_klist_t;

static void klist_offer( orc_klist_t* L, float d, int32_t id )
{
  if( L->len == L->cap && !( d < L->d2[L->len - 1] ) ) { return; }
  size_t p = L->len < L->cap ? L->len : L->cap - 1;
  while( p > 0 && d < L->d2[p - 1] ) { L->d2[p] = L->d2[p - 1]; L->id[p] = L->id[p - 1]; --p; }
  L->d2[p] = d; L->id[p] = id;
  if( L->len < L->cap ) { L->len++; }
}

typedef struct { float m; int64_t key; int32_t seq; } orc_cellref_t;
static int cmp_cellref( const void* a, const void* b )
{
  const orc_cellref_t *x = (const orc_cellref_t*)a, *y = (const orc_cellref_t*)b;
  if( x->m < y->m ) { return -1; }
  if( x->m > y->m ) { return 1; }
  return ( x->seq > y->seq ) - ( x->seq < y->seq );
}

/* One query of msh_hash_grid_radius_search (:1144-1249): the k nearest points with dist^2 < (float)(r*r),
   ascending.  Cells are enumerated z-outer / x-inner over the int64-truncated range of (q -+ r)*inv_cell
   and capped at 512 (:1165-1225), visited nearest-cell-first; results do not depend on that order except
   between exactly equal distances. */
static size_t orc_radius_query( const orc_grid_t* g, const float* qp, double radius, float r2f,
                                size_t k, float* d2, int32_t* id )
{
  float q[3];
  int64_t c0[3], lo[3], hi[3];
  for( int a = 0; a < 3; ++a )
  {
    q[a] = qp[a] - g->mn[a];                               /* float (:1159-1161) */
    c0[a] = (int64_t)( q[a] * g->inv_cell );               /* double product, truncation (:1165-1167) */
    hi[a] = (int64_t)( ( q[a] + radius ) * g->inv_cell );
    lo[a] = (int64_t)( ( q[a] - radius ) * g->inv_cell );
  }
  orc_cellref_t cells[ORC_MAX_CELLS_PER_QUERY];
  int32_t nc = 0;
  int64_t slab = (int64_t)(int32_t)( g->dim[1] * g->dim[0] );
  for( int64_t z = lo[2]; z <= hi[2] && nc < ORC_MAX_CELLS_PER_QUERY; ++z )
  {
    if( z < 0 || z >= g->dim[2] ) { continue; }
    float gz = z < c0[2] ? (float)( q[2] - ( z + 1 ) * g->cell ) : ( z > c0[2] ? (float)( z * g->cell - q[2] ) : 0.0f );
    for( int64_t y = lo[1]; y <= hi[1] && nc < ORC_MAX_CELLS_PER_QUERY; ++y )
    {
      if( y < 0 || y >= g->dim[1] ) { continue; }
      float gy = y < c0[1] ? (float)( q[1] - ( y + 1 ) * g->cell ) : ( y > c0[1] ? (float)( y * g->cell - q[1] ) : 0.0f );
      for( int64_t x = lo[0]; x <= hi[0]; ++x )
      {
        if( x < 0 || x >= g->dim[0] ) { continue; }
        if( nc >= ORC_MAX_CELLS_PER_QUERY ) { break; }
        float gx = x < c0[0] ? (float)( q[0] - ( x + 1 ) * g->cell ) : ( x > c0[0] ? (float)( x * g->cell - q[0] ) : 0.0f );
        cells[nc].m = gz * gz + gy * gy + gx * gx;           /* (:1221) */
        cells[nc].key = (int64_t)(int32_t)( z * slab + y * g->dim[0] + x ); /* bin_indices is int32 (:1140) */
        cells[nc].seq = nc;
        nc++;
      }
    }
  }
  qsort( cells, nc, sizeof( orc_cellref_t ), cmp_cellref );
  orc_klist_t L = { d2, id, k, 0 };
  for( int32_t c = 0; c < nc; ++c )
  {
    if( L.len == L.cap && L.d2[L.len - 1] <= cells[c].m ) { break; } /* (:1232-1236) */
    int64_t s = orc_find_cell( g, cells[c].key );
    if( s < 0 ) { continue; }
    for( int32_t p = g->cell_start[s]; p < g->cell_start[s + 1]; ++p )
    {
      float vx = g->xyz[3 * p + 0] - qp[0], vy = g->xyz[3 * p + 1] - qp[1], vz = g->xyz[3 * p + 2] - qp[2];
      float dd = vx * vx + vy * vy + vz * vz;                /* (:852-855) */
      if( dd < r2f ) { klist_offer( &L, dd, g->idx[p] ); }
    }
  }
  return L.len;
}

size_t orc_radius_search( const orc_grid_t* g, const float* q, size_t nq, float radius, size_t k, int sort,
                          float* d2, int32_t* idx, uint64_t* nn )
{
  (void)sort; /* rows are always returned ascending; the reference's unsorted order is unspecified */
  double r = radius;
  float r2f = (float)( r * r ); /* double product narrowed when passed down (:1111, 828) */
  uint32_t total = 0;
#if defined(_OPENMP)
  #pragma omp parallel for schedule(dynamic, 64) reduction(+:total)
#endif
  for( int64_t i = 0; i < (int64_t)nq; ++i )
  {
    size_t c = orc_radius_query( g, q + 3 * i, r, r2f, k, d2 + i * k, idx + i * k );
    if( nn ) { nn[i] = c; }
    total += (uint32_t)c;
  }
  return total;
}

/* msh_hash_grid_knn_search (:1294-1450): shells of cells around the query's cell are opened one layer at a
   time; the search stops after the layer FOLLOWING the one that first filled the list.  Pruned shell cells
   (min distance > current k-th, :1409) cannot hold a better point, so the result is the k nearest points of
   the cube of half-width L+1 cells, L = first layer at which >= k points were seen. */
size_t orc_knn_search( const orc_grid_t* g, const float* qs, size_t nq, size_t k, int sort,
                       float* d2, int32_t* idx, uint64_t* nn )
{
  (void)sort;
  uint32_t total = 0;
  int64_t slab = (int64_t)(int32_t)( g->dim[1] * g->dim[0] );
  int64_t max_layer = g->dim[0] + g->dim[1] + g->dim[2];
  for( size_t i = 0; i < nq; ++i )
  {
    const float* qp = qs + 3 * i;
    int64_t c0[3];
    for( int a = 0; a < 3; ++a ) { c0[a] = (int64_t)(uint64_t)( ( qp[a] - g->mn[a] ) * g->inv_cell ); }
    orc_klist_t L = { d2 + i * k, idx + i * k, k, 0 };
    int stop_next = 0;
    for( int64_t layer = 0; layer <= max_layer; ++layer )
    {
      for( int64_t z = c0[2] - layer; z <= c0[2] + layer; ++z )
        for( int64_t y = c0[1] - layer; y <= c0[1] + layer; ++y )
          for( int64_t x = c0[0] - layer; x <= c0[0] + layer; ++x )
          {
            int64_t m = llabs( x - c0[0] ) > llabs( y - c0[1] ) ? llabs( x - c0[0] ) : llabs( y - c0[1] );
            m = m > llabs( z - c0[2] ) ? m : llabs( z - c0[2] );
            if( m != layer ) { continue; } /* shell surface only (:1395-1398) */
            if( x < 0 || y < 0 || z < 0 || x >= g->dim[0] || y >= g->dim[1] || z >= g->dim[2] ) { continue; }
            int64_t s = orc_find_cell( g, z * slab + y * g->dim[0] + x );
            if( s < 0 ) { continue; }
            for( int32_t p = g->cell_start[s]; p < g->cell_start[s + 1]; ++p )
            {
              float vx = g->xyz[3 * p + 0] - qp[0], vy = g->xyz[3 * p + 1] - qp[1], vz = g->xyz[3 * p + 2] - qp[2];
              klist_offer( &L, vx * vx + vy * vy + vz * vz, g->idx[p] ); /* (:1281-1289) */
            }
          }
      if( stop_next ) { break; }                 /* (:1427-1429) */
      if( L.len >= k ) { stop_next = 1; }
    }
    if( nn ) { nn[i] = L.len; }
    total += (uint32_t)L.len;
  }
  return total;
}

/* ------------------------------------------------------------------------------------------ small math */

/* msh_mat4_vec3_mul (msh_vec_math.h:1554-1561): left-to-right float sum, translation times (float)is_point */
static void xf_apply( const float* m, const float* v, int is_point, float* o )
{
  float w = (float)is_point;
  o[0] = m[0] * v[0] + m[4] * v[1] + m[8] * v[2] + w * m[12];
  o[1] = m[1] * v[0] + m[5] * v[1] + m[9] * v[2] + w * m[13];
  o[2] = m[2] * v[0] + m[6] * v[1] + m[10] * v[2] + w * m[14];
}
void orc_xf_apply( const float* m, const float* v, int is_point, float* o ) { xf_apply( m, v, is_point, o ); }

/* msh_mat4_mul (msh_vec_math.h:1441-1480): o = a*b, column-major, each entry a 4-term left-to-right sum */
static void xf_mul( const float* a, const float* b, float* o )
{
  float t[16];
  for( int c = 0; c < 4; ++c )
    for( int r = 0; r < 4; ++r )
      t[4 * c + r] = b[4 * c + 0] * a[r] + b[4 * c + 1] * a[4 + r] + b[4 * c + 2] * a[8 + r] + b[4 * c + 3] * a[12 + r];
  memcpy( o, t, 64 );
}
void orc_xf_mul( const float* a, const float* b, float* o ) { xf_mul( a, b, o ); }

static void xf_identity( float* m ) { memset( m, 0, 64 ); m[0] = m[5] = m[10] = m[15] = 1.0f; }

/* msh_translate (msh_vec_math.h:2064-2073): col3 = (col0*tx + col1*ty) + (col2*tz + col3) */
static void xf_translate( float* m, float tx, float ty, float tz )
{
  for( int r = 0; r < 4; ++r ) { m[12 + r] = ( m[r] * tx + m[4 + r] * ty ) + ( m[8 + r] * tz + m[12 + r] ); }
}

/* msh_rotate (msh_vec_math.h:2089-2136): axis-angle matrix in float from cosf/sinf, then
   new col_j = col0*R[4j] + (col1*R[4j+1] + col2*R[4j+2]) */
static void xf_rotate( float* m, float angle, float ax, float ay, float az )
{
  float c = cosf( angle ), s = sinf( angle ), t = 1.0f - c;
  float inv = 1.0f / sqrtf( ax * ax + ay * ay + az * az );
  ax = ax * inv; ay = ay * inv; az = az * inv;
  float R[16] = { 0 };
  R[0] = c + ax * ax * t; R[5] = c + ay * ay * t; R[10] = c + az * az * t;
  float a = ax * ay * t, b = az * s;
  R[1] = a + b; R[4] = a - b;
  a = ax * az * t; b = ay * s;
  R[2] = a - b; R[8] = a + b;
  a = ay * az * t; b = ax * s;
  R[6] = a + b; R[9] = a - b;
  float o[16];
  memcpy( o, m, 64 );
  for( int j = 0; j < 3; ++j )
    for( int r = 0; r < 4; ++r )
      o[4 * j + r] = m[r] * R[4 * j] + ( m[4 + r] * R[4 * j + 1] + m[8 + r] * R[4 * j + 2] );
  memcpy( m, o, 64 );
}

/* the candidate pose of pose_proposal.cpp:221-222 */
void orc_make_pose( float y_angle, float tx, float ty, float tz, float* out )
{
  xf_identity( out );
  xf_rotate( out, y_angle, 0.0f, 1.0f, 0.0f );
  out[12] = tx; out[13] = ty; out[14] = tz; out[15] = 1.0f;
}

/* msh_mat4_inverse (msh_vec_math.h:1818-1917): cofactor expansion over 2x2 minors, all float, each
   cofactor a three-term expression evaluated left to right, scaled by 1.0f/det */
static float tri( float a, float x, float b, float y, float c, float z, int s2, int s3 )
{
  float r = a * x;
  r = s2 > 0 ? r + b * y : r - b * y;
  r = s3 > 0 ? r + c * z : r - c * z;
  return r;
}
static void xf_inverse( const float* m, float* o )
{
  float C[16], d[6];
  d[0] = m[10] * m[15] - m[14] * m[11]; d[1] = m[6] * m[11] - m[10] * m[7]; d[2] = m[2] * m[7] - m[6] * m[3];
  d[3] = m[6] * m[15] - m[14] * m[7];   d[4] = m[2] * m[11] - m[10] * m[3]; d[5] = m[2] * m[15] - m[14] * m[3];
  C[0] = tri( m[5], d[0], m[9], d[3], m[13], d[1], -1, +1 );
  C[1] = tri( m[9], d[5], m[1], d[0], m[13], d[4], -1, -1 );
  C[2] = tri( m[1], d[3], m[5], d[5], m[13], d[2], -1, +1 );
  C[3] = tri( m[5], d[4], m[9], d[2], m[1], d[1], -1, -1 );
  C[4] = tri( m[8], d[3], m[4], d[0], m[12], d[1], -1, -1 );
  C[5] = tri( m[0], d[0], m[8], d[5], m[12], d[4], -1, +1 );
  C[6] = tri( m[4], d[5], m[0], d[3], m[12], d[2], -1, -1 );
  C[7] = tri( m[0], d[1], m[4], d[4], m[8], d[2], -1, +1 );
  d[0] = m[8] * m[13] - m[12] * m[9]; d[1] = m[4] * m[9] - m[8] * m[5];  d[2] = m[0] * m[5] - m[4] * m[1];
  d[3] = m[4] * m[13] - m[12] * m[5]; d[4] = m[0] * m[9] - m[8] * m[1];  d[5] = m[0] * m[13] - m[12] * m[1];
  C[8]  = tri( m[7], d[0], m[11], d[3], m[15], d[1], -1, +1 );
  C[9]  = tri( m[11], d[5], m[3], d[0], m[15], d[4], -1, -1 );
  C[10] = tri( m[3], d[3], m[7], d[5], m[15], d[2], -1, +1 );
  C[11] = tri( m[7], d[4], m[3], d[1], m[11], d[2], -1, -1 );
  C[12] = tri( m[10], d[3], m[6], d[0], m[14], d[1], -1, -1 );
  C[13] = tri( m[2], d[0], m[10], d[5], m[14], d[4], -1, +1 );
  C[14] = tri( m[6], d[5], m[2], d[3], m[14], d[2], -1, -1 );
  C[15] = tri( m[2], d[1], m[6], d[4], m[10], d[2], -1, +1 );
  float det = m[0] * C[0] + m[4] * C[1] + m[8] * C[2] + m[12] * C[3];
  float s = 1.0f / det;
  for( int i = 0; i < 16; ++i ) { o[i] = s * C[i]; }
}
void orc_xf_inverse( const float* m, float* o ) { xf_inverse( m, o ); }

/* ------------------------------------------------------------------------------------------- scoring */

/* mgs_compute_object_alignment_score (pose_proposal.cpp:93-158).  `grid`/`scan_nor` are the scan's level
   `search_lvl` grid and normals, radius = sigma = search_radii[search_lvl] (0.10 for level 1, :98). */
float orc_score_pose( const float* obj_pos, const float* obj_nor, int32_t n_obj, const orc_grid_t* grid,
                      const float* scan_nor, const float* xform, int32_t k, float radius )
{
  double max_angle = 35.0 * 0.005555555556 * 3.1415926535897932384626433832; /* msh_deg2rad (msh_std.h:618,625) */
  double sigma = radius, alpha = 0.05, beta = 1.0 - alpha;
  float* d2 = (float*)malloc( sizeof( float ) * k );
  int32_t* id = (int32_t*)malloc( sizeof( int32_t ) * k );
  double r = radius;
  float r2f = (float)( r * r );
  double sum = 0.0;
  for( int32_t i = 0; i < n_obj; ++i )
  {
    float p[3], n[3];
    xf_apply( xform, obj_pos + 3 * i, 1, p );
    xf_apply( xform, obj_nor + 3 * i, 0, n );
    size_t cnt = orc_radius_query( grid, p, r, r2f, (size_t)k, d2, id );
    double hit_d2 = -1.0, hit_angle = 0.0;
    for( size_t j = 0; j < cnt; ++j )
    {
      const float* m = scan_nor + 3 * (size_t)id[j];
      float dotf = m[0] * n[0] + m[1] * n[1] + m[2] * n[2];   /* msh_vec3_dot(m, n) */
      double dot = dotf > 0.0f ? dotf : 0.0f;
      double angle = acos( dot );                            /* NaN when dot > 1: never accepted */
      if( angle - max_angle < 0.000001 ) { hit_d2 = d2[j]; hit_angle = angle; break; }
    }
    if( hit_d2 < -0.0001 ) { continue; }
    double nc = exp( -( hit_angle * hit_angle ) / ( 2.0 * 0.5 * 0.5 ) );
    double dc = exp( -hit_d2 / ( 2.0 * sigma * sigma ) );
    sum += alpha * nc + beta * dc;
  }
  free( d2 ); free( id );
  sum /= (double)n_obj;
  return (float)sum;
}

double orc_score_batch( const float* obj_pos, const float* obj_nor, int32_t n_obj, const orc_grid_t* grid,
                        const float* scan_nor, const float* xforms, int64_t n_poses, int32_t k, float radius,
                        float* out, int n_threads )
{
  double t0 = 0, t1 = 0;
#if defined(_OPENMP)
  t0 = omp_get_wtime();
  #pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads > 0 ? n_threads : 1)
#endif
  for( int64_t p = 0; p < n_poses; ++p )
  {
    out[p] = orc_score_pose( obj_pos, obj_nor, n_obj, grid, scan_nor, xforms + 16 * p, k, radius );
  }
#if defined(_OPENMP)
  t1 = omp_get_wtime();
#endif
  return t1 - t0;
}

/* mgs__score_threshold (pose_proposal.cpp:160-168) */
float orc_score_threshold( int lvl )
{
  if( lvl == 4 ) { return 0.25f; }
  if( lvl == 3 ) { return 0.35f; }
  if( lvl == 2 ) { return 0.40f; }
  return 0.50f;
}

/* The selection of mgs__initial_pose_proposals (pose_proposal.cpp:213-243) over a scores[T][R] table:
   per translation the first strict maximum over rotations starting from 0, emitted iff > threshold, in
   translation order.  Returns the number emitted; out_t/out_r/out_s sized T. */
int32_t orc_select_proposals( const float* scores, int32_t T, int32_t R, float threshold,
                              int32_t* out_t, int32_t* out_r, float* out_s )
{
  int32_t n = 0;
  for( int32_t t = 0; t < T; ++t )
  {
    float best = 0; int32_t br = -1;
    for( int32_t r = 0; r < R; ++r ) { float s = scores[(size_t)t * R + r]; if( s > best ) { best = s; br = r; } }
    if( best > threshold ) { out_t[n] = t; out_r[n] = br; out_s[n] = best; n++; }
  }
  return n;
}

/* ----------------------------------------------------------------------------------------------- ICP */

/* icp_find_corrs (icp.h:306-412) with T2 = identity-or-any; grid is the index over pts2 built with the
   INITIAL max_dist (icp.h:437).  Outputs are compacted in pts1 order; returns n_corrs. */
int32_t orc_icp_find_corrs( const float* p1, const float* n1, int32_t c1, const orc_grid_t* grid2,
                            const float* p2, const float* n2, const float* T1, const float* T2,
                            float max_dist, float max_angle,
                            float* cp1, float* cn1, float* cp2, float* cn2, float* w )
{
  float T2i[16];
  xf_inverse( T2, T2i );
  enum { KNN = 16 };
  float d2[KNN]; int32_t id[KNN];
  double r = max_dist;
  float r2f = (float)( r * r );
  float* kept = (float*)malloc( sizeof( float ) * ( c1 ? c1 : 1 ) );
  int32_t nc = 0;
  for( int32_t i = 0; i < c1; ++i )
  {
    float a[3], b[3], q[3], qn[3];
    xf_apply( T1, p1 + 3 * i, 1, a ); xf_apply( T1, n1 + 3 * i, 0, b );
    xf_apply( T2i, a, 1, q );         xf_apply( T2i, b, 0, qn );
    size_t cnt = orc_radius_query( grid2, q, r, r2f, KNN, d2, id );
    for( size_t j = 0; j < cnt; ++j )
    {
      const float* m = n2 + 3 * (size_t)id[j];
      float dot = m[0] * qn[0] + m[1] * qn[1] + m[2] * qn[2];
      dot = dot > 0.0f ? dot : 0.0f;
      if( acosf( dot ) < max_angle )
      {
        memcpy( cp1 + 3 * nc, q, 12 ); memcpy( cn1 + 3 * nc, qn, 12 );
        memcpy( cp2 + 3 * nc, p2 + 3 * (size_t)id[j], 12 ); memcpy( cn2 + 3 * nc, m, 12 );
        w[nc] = ( 1.0f - d2[j] / max_dist ) * dot;          /* squared distance over metres (:387) */
        kept[nc] = d2[j];
        nc++;
        break;
      }
    }
  }
  /* outlier pass over the SQUARED distances, threshold 2.5*stddev without the mean (:394-402;
     msh_compute_mean / msh_compute_stddev, msh_std.h:1811-1824: float sums, sqrt in double) */
  float s1 = 0, s2 = 0;
  for( int32_t i = 0; i < nc; ++i ) { s1 += kept[i]; }
  float mean = s1 / (float)nc;
  for( int32_t i = 0; i < nc; ++i ) { s2 += kept[i] * kept[i]; }
  float sd = (float)sqrt( s2 / (float)nc - mean * mean );
  if( sd > 0.000001 )
    for( int32_t i = 0; i < nc; ++i ) { if( kept[i] > 2.5f * sd ) { w[i] = 0.0; } }
  free( kept );
  return nc;
}

/* LDL^T without pivoting on the upper triangle, then two triangular sweeps
   (trimesh::ldltdc / ldltsl for N = 6, lineqn.h:177-193, 206-217).  A zero pivot aborts the factorisation
   (lineqn.h:185-186) but the caller ignores that (icp.h:276-277) and still runs the sweeps over the
   half-factored matrix with the remaining reciprocal pivots left at 0 — kept, because exactly singular
   systems do occur (all correspondences on one axis-aligned plane). */
static void ldlt_solve6( double A[6][6], const double* b, double* x )
{
  double rd[6] = { 0, 0, 0, 0, 0, 0 }, v[5];
  int ok = 1;
  for( int i = 0; i < 6 && ok; ++i )
  {
    for( int k = 0; k < i; ++k ) { v[k] = A[i][k] * rd[k]; }
    for( int j = i; j < 6; ++j )
    {
      double s = A[i][j];
      for( int k = 0; k < i; ++k ) { s -= v[k] * A[j][k]; }
      if( i == j ) { if( s == 0 ) { ok = 0; break; } rd[i] = 1 / s; } else { A[j][i] = s; }
    }
  }
  for( int i = 0; i < 6; ++i )
  {
    double s = b[i];
    for( int k = 0; k < i; ++k ) { s -= A[i][k] * x[k]; }
    x[i] = s * rd[i];
  }
  for( int i = 5; i >= 0; --i )
  {
    double s = 0;
    for( int k = i + 1; k < 6; ++k ) { s += A[k][i] * x[k]; }
    x[i] -= s * rd[i];
  }
}

/* icp_estimate_rigid_xform_pt2pl (icp.h:210-298) */
float orc_icp_pt2pl( const float* cp1, const float* cp2, const float* cn2, const float* w, int32_t n, float* T1 )
{
  /* weighted centroids, float accumulation (icp.h:137-148) */
  float c1[3] = { 0, 0, 0 }, c2[3] = { 0, 0, 0 }, tw = 0.0f;
  for( int32_t i = 0; i < n; ++i ) { tw += w[i]; for( int a = 0; a < 3; ++a ) { c1[a] = c1[a] + cp1[3 * i + a] * w[i]; } }
  { float inv = 1.0f / tw; for( int a = 0; a < 3; ++a ) { c1[a] = c1[a] * inv; } } /* msh_vec3_scalar_div multiplies by 1/s */
  tw = 0.0f;
  for( int32_t i = 0; i < n; ++i ) { tw += w[i]; for( int a = 0; a < 3; ++a ) { c2[a] = c2[a] + cp2[3 * i + a] * w[i]; } }
  { float inv = 1.0f / tw; for( int a = 0; a < 3; ++a ) { c2[a] = c2[a] * inv; } }

  float TL[3][3] = { { 0 } }, TR[3][3] = { { 0 } }, BR[3][3] = { { 0 } }, rhs[6] = { 0 }; /* [col][row] */
  double sum = 0.0, wsum = 0.0;
  for( int32_t i = 0; i < n; ++i )
  {
    float p[3], q[3], nn[3], d[3], c[3];
    for( int a = 0; a < 3; ++a ) { p[a] = cp1[3 * i + a] - c1[a]; q[a] = cp2[3 * i + a] - c2[a]; nn[a] = cn2[3 * i + a]; d[a] = p[a] - q[a]; }
    c[0] = p[1] * nn[2] - p[2] * nn[1]; c[1] = p[2] * nn[0] - p[0] * nn[2]; c[2] = p[0] * nn[1] - p[1] * nn[0];
    float wi = w[i];
    float dn = d[0] * nn[0] + d[1] * nn[1] + d[2] * nn[2];
    for( int col = 0; col < 3; ++col )
      for( int row = 0; row < 3; ++row )
      {
        TL[col][row] = TL[col][row] + ( c[row] * c[col] ) * wi;   /* outer(a,b)[col][row] = a[row]*b[col] */
        TR[col][row] = TR[col][row] + ( c[row] * nn[col] ) * wi;
        BR[col][row] = BR[col][row] + ( nn[row] * nn[col] ) * wi;
      }
    for( int a = 0; a < 3; ++a ) { rhs[a] += wi * c[a] * dn; rhs[3 + a] += wi * nn[a] * dn; }
    sum += wi * dn * dn;   /* float product promoted (icp.h:250) */
    wsum += wi;
  }
  float err = (float)sqrt( sum / wsum );

  double A[6][6], b[6], x[6] = { 0 };
  for( int r = 0; r < 3; ++r )
    for( int c = 0; c < 3; ++c )
    {
      A[r][c] = TL[c][r]; A[r][3 + c] = TR[c][r];
      A[3 + r][c] = TR[r][c]; A[3 + r][3 + c] = BR[c][r];    /* (icp.h:267-272) */
    }
  for( int a = 0; a < 6; ++a ) { b[a] = -rhs[a]; }
  ldlt_solve6( A, b, x );

  float T[16];
  xf_identity( T );
  xf_translate( T, c1[0], c1[1], c1[2] );
  xf_translate( T, (float)x[3], (float)x[4], (float)x[5] );
  xf_rotate( T, (float)x[0], 1.0f, 0.0f, 0.0f );
  xf_rotate( T, (float)x[1], 0.0f, 1.0f, 0.0f );
  xf_rotate( T, (float)x[2], 0.0f, 0.0f, 1.0f );
  xf_translate( T, -c1[0], -c1[1], -c1[2] );
  xf_mul( T, T1, T1 );
  return err;
}

/* icp_align (icp.h:416-500).  n_iters_out receives the number of estimation steps taken. */
float orc_icp_align( const float* p1, const float* n1, int32_t c1, const float* p2, const float* n2, int32_t c2,
                     float* T1, const float* T2, float max_dist, float max_angle, int32_t* n_iters_out )
{
  orc_grid_t* g2 = orc_grid_build( p2, c2, max_dist );
  size_t m = c1 ? c1 : 1;
  float *cp1 = (float*)malloc( 12 * m ), *cn1 = (float*)malloc( 12 * m ), *cp2 = (float*)malloc( 12 * m ),
        *cn2 = (float*)malloc( 12 * m ), *w = (float*)malloc( 4 * m );
  float prev = 1e6, err = 1e6;
  int32_t steps = 0;
  for( int i = 0; i < 100; ++i )
  {
    prev = err;
    int32_t nc = orc_icp_find_corrs( p1, n1, c1, g2, p2, n2, T1, T2, max_dist, max_angle, cp1, cn1, cp2, cn2, w );
    if( nc == 0 ) { break; }
    float tw = 0.0;
    for( int32_t j = 0; j < nc; ++j ) { tw += w[j]; }
    if( tw <= 1e-7 ) { break; }
    err = orc_icp_pt2pl( cp1, cp2, cn2, w, nc, T1 );
    steps++;
    float delta = fabsf( prev - err );
    if( i > 5 && delta < 1e-5 ) { break; }
    double shrunk = max_dist * 0.95;
    max_dist = (float)( shrunk > 0.05 ? shrunk : 0.05 );
  }
  if( n_iters_out ) { *n_iters_out = steps; }
  free( cp1 ); free( cn1 ); free( cp2 ); free( cn2 ); free( w );
  orc_grid_free( g2 );
  return err;
}

/* ------------------------------------------------------------------------------- labels / unary terms */

/* rspf__assign_temporary_labels over placements [first, last) (rs_pointcloud_filters.cpp:738-778).
   obj_grid[i] / obj_nor[i] are the level-1 grid and normals of placement i's object. */
void orc_assign_labels( const float* scan_pos, const float* scan_nor, int32_t V, const float* poses,
                        const orc_grid_t* const* obj_grid, const float* const* obj_nor,
                        int32_t first, int32_t last, float radius, int8_t* labels, float* min_d )
{
  double max_angle = 70.0 * 0.005555555556 * 3.1415926535897932384626433832;
  double r = radius;
  float r2f = (float)( r * r );
  for( int32_t i = first; i < last; ++i )
  {
    const float* M = poses + 16 * i;
    float Mi[16], Mt[16];
    xf_inverse( M, Mi );
    for( int c = 0; c < 4; ++c ) for( int rr = 0; rr < 4; ++rr ) { Mt[4 * c + rr] = M[4 * rr + c]; }
    for( int32_t j = 0; j < V; ++j )
    {
      float q[3], d2; int32_t id;
      xf_apply( Mi, scan_pos + 3 * j, 1, q );
      if( orc_radius_query( obj_grid[i], q, r, r2f, 1, &d2, &id ) == 0 ) { continue; }
      if( !( d2 < min_d[j] ) ) { continue; }
      float a[3];
      xf_apply( Mt, scan_nor + 3 * j, 0, a );
      const float* b = obj_nor[i] + 3 * (size_t)id;
      float ia = 1.0f / sqrtf( a[0] * a[0] + a[1] * a[1] + a[2] * a[2] );
      float ib = 1.0f / sqrtf( b[0] * b[0] + b[1] * b[1] + b[2] * b[2] );
      float an[3] = { a[0] * ia, a[1] * ia, a[2] * ia }, bn[3] = { b[0] * ib, b[1] * ib, b[2] * ib };
      float dot = an[0] * bn[0] + an[1] * bn[1] + an[2] * bn[2];
      float angle = acosf( fabsf( dot ) );                     /* C++ overloads pick the float versions (:767) */
      if( angle < max_angle ) { min_d[j] = d2; labels[j] = (int8_t)( i + 1 ); }
    }
  }
}

/* data_cost block of rspf_smooth_labels (rs_pointcloud_filters.cpp:926-939): cost[v][l] = 0 for the
   vertex's own label, else 30 (15 when its class is static, 1 when the label is 0) */
void orc_unary_costs( const int32_t* labels, const uint8_t* label_is_static, int32_t V, int32_t L, int32_t* cost )
{
  for( int32_t v = 0; v < V; ++v )
  {
    int32_t c = 30;
    if( label_is_static[labels[v]] ) { c = 15; }
    if( labels[v] == 0 ) { c = 1; }
    for( int32_t l = 0; l < L; ++l ) { cost[(size_t)v * L + l] = ( l == labels[v] ) ? 0 : c; }
  }
}

/* candidate edges of rspf_compute_neighborhood (rs_pointcloud_filters.cpp:674-722) BEFORE the hashtable
   de-duplication: for every vertex its <= max_nn nearest neighbours within sqrt(radius_sq) (self included)
   and the weight (1-(d2/(4 r^2))^dist_exp) * clamp(n.m,0,1)^angle_exp.  Output rows are V x max_nn,
   idx -1 where absent. */
void orc_neighborhood( const orc_grid_t* grid, const float* pos, const float* nor, int32_t V, int32_t max_nn,
                       float radius_sq, float dist_exp, float angle_exp, int32_t* nbr, float* weight )
{
  float radius = (float)sqrt( radius_sq );
  double r = radius;
  float r2f = (float)( r * r );
  float* d2 = (float*)malloc( 4 * (size_t)max_nn );
  int32_t* id = (int32_t*)malloc( 4 * (size_t)max_nn );
  for( int32_t i = 0; i < V; ++i )
  {
    size_t cnt = orc_radius_query( grid, pos + 3 * i, r, r2f, (size_t)max_nn, d2, id );
    for( int32_t j = 0; j < max_nn; ++j ) { nbr[(size_t)i * max_nn + j] = -1; weight[(size_t)i * max_nn + j] = 0; }
    for( size_t j = 0; j < cnt; ++j )
    {
      const float *n = nor + 3 * i, *m = nor + 3 * (size_t)id[j];
      float dot = n[0] * m[0] + n[1] * m[1] + n[2] * m[2];
      dot = dot > 0.0f ? dot : 0.0f; dot = dot < 1.0f ? dot : 1.0f;
      float dist_cost = 1.0f - pow( d2[j] / ( 4.0 * radius_sq ), dist_exp );
      float norm_cost = powf( dot, angle_exp );                /* float overload of pow in C++ (:707) */
      nbr[(size_t)i * max_nn + j] = id[j];
      weight[(size_t)i * max_nn + j] = dist_cost * norm_cost;
    }
  }
  free( d2 ); free( id );
}

/* ------------------------------------------------------------------------------------------------
 * Voxel-occupancy overlap of one object under two poses, and the greedy non-maxima suppression built on
 * it (reference lib/rs/intersect.h:309-368, apps/pose_proposal/pose_proposal.cpp:371-452).  Restated as
 * three flat passes over one byte grid per pose (mark, fill, count) instead of the reference's slice
 * copies and scan-line scratch arrays; the grid geometry, the per-row "odd number of boundary exits seen
 * from both ends" fill rule and every float expression are the reference's.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { float mn[3], mx[3]; } orc_bbox_t;

/* bounding box of the level-3 points under a pose (intersect.h:119-130, msh_geometry.h:945-969) */
void orc_posed_bbox( const float* pos3, int32_t n3, const float* pose, float* mn_mx )
{
  float mn[3] = { 1e9f, 1e9f, 1e9f }, mx[3] = { -1e9f, -1e9f, -1e9f };
  for( int32_t i = 0; i < n3; ++i )
  {
    float p[3];
    xf_apply( pose, pos3 + 3 * (size_t)i, 1, p );
    for( int a = 0; a < 3; ++a ) { if( p[a] < mn[a] ) { mn[a] = p[a]; } if( p[a] > mx[a] ) { mx[a] = p[a]; } }
  }
  memcpy( mn_mx, mn, 12 ); memcpy( mn_mx + 3, mx, 12 );
}

/* occupancy of the level-1 points under `pose` on the pair grid: 1 = boundary, 2 = inside (voxelize_inside);
   returns the number of non-free cells (intersect.h:177-306).  Cell (x,y,z) lives at (y*zr + z)*xr + x (:108). */
static int32_t orc_occupancy( const float* pos1, int32_t n1, const float* pose, const float* origin, float voxel,
                              int xr, int yr, int zr, int inside, uint8_t* grid )
{
  const size_t n_cells = (size_t)xr * yr * zr;
  memset( grid, 0, n_cells );
  for( int32_t i = 0; i < n1; ++i )
  {
    float p[3];
    xf_apply( pose, pos1 + 3 * (size_t)i, 1, p );
    const int x = (int)floorf( ( p[0] - origin[0] ) / voxel ), y = (int)floorf( ( p[1] - origin[1] ) / voxel ),
              z = (int)floorf( ( p[2] - origin[2] ) / voxel ); /* (:226-230); in range by construction of the grid */
    if( x < 0 || x >= xr || y < 0 || y >= yr || z < 0 || z >= zr ) { continue; }
    grid[( (size_t)y * zr + z ) * xr + x] = 1;
  }
  int32_t count = 0;
  if( !inside )
  {
    for( size_t c = 0; c < n_cells; ++c ) { count += grid[c] == 1; }
    return count;
  }
  /* per y layer, a free cell becomes "inside" iff along its x-row AND along its z-row the number of boundary->free
     transitions passed is odd when walking from the row's start and odd when walking from its end (:129-175, 253-284) */
  uint8_t* in_x = (uint8_t*)malloc( (size_t)xr * zr );
  uint8_t* in_z = (uint8_t*)malloc( (size_t)xr * zr );
  for( int y = 0; y < yr; ++y )
  {
    uint8_t* layer = grid + (size_t)y * zr * xr;
    memset( in_x, 0, (size_t)xr * zr ); memset( in_z, 0, (size_t)xr * zr );
    for( int dir = 0; dir < 2; ++dir )
    {
      /* dir 0: rows of constant z, walking x (the reference's dir = 0: r1 = z, r2 = x); dir 1: constant x, walking z */
      const int n_rows = dir == 0 ? zr : xr, len = dir == 0 ? xr : zr;
      uint8_t* out = dir == 0 ? in_x : in_z;
      for( int r = 0; r < n_rows; ++r )
      {
        #define ORC_CELL( t ) ( dir == 0 ? (size_t)r * xr + ( t ) : (size_t)( t ) * xr + r )
        int fill = 0, prev = 0;
        for( int t = 0; t < len; ++t )
        {
          const int v = layer[ORC_CELL( t )];
          if( v == 0 && prev == 1 ) { fill += 1; }
          if( fill % 2 == 1 ) { out[ORC_CELL( t )] = 1; } /* forward mark */
          prev = v;
        }
        fill = 0; prev = 0;
        for( int t = len - 1; t >= 0; --t )
        {
          const int v = layer[ORC_CELL( t )];
          if( v == 0 && prev == 1 ) { fill += 1; }
          /* inside along this direction = forward AND backward odd AND not a boundary cell */
          out[ORC_CELL( t )] = (uint8_t)( out[ORC_CELL( t )] && ( fill % 2 == 1 ) && v != 1 );
          prev = v;
        }
        #undef ORC_CELL
      }
    }
    for( size_t c = 0; c < (size_t)xr * zr; ++c )
    {
      if( layer[c] != 1 && in_x[c] && in_z[c] ) { layer[c] = 2; }
    }
  }
  free( in_x ); free( in_z );
  for( size_t c = 0; c < n_cells; ++c ) { count += grid[c] > 0; }
  return count;
}

float orc_overlap_factor( const float* pos3, int32_t n3, const float* pos1, int32_t n1, const float* pose_a, const float* pose_b,
                          float voxel, int inside, int normalize_by_smaller )
{
  float ba[6], bb[6];
  orc_posed_bbox( pos3, n3, pose_a, ba );
  orc_posed_bbox( pos3, n3, pose_b, bb );
  /* mshgeo_bbox_intersect (msh_geometry.h:1010-1015) */
  for( int a = 0; a < 3; ++a ) { if( !( ba[3 + a] >= bb[a] && bb[3 + a] >= ba[a] ) ) { return 0.0f; } }
  float mn[3], mx[3];
  for( int a = 0; a < 3; ++a )
  {
    mn[a] = ba[a] < bb[a] ? ba[a] : bb[a];
    mx[a] = ba[3 + a] > bb[3 + a] ? ba[3 + a] : bb[3 + a];
    mn[a] = mn[a] - 0.3f; mx[a] = mx[a] + 0.3f; /* fat_factor (intersect.h:61-65) */
  }
  const int xr = (int)ceilf( ( mx[0] - mn[0] ) / voxel ) + 1, yr = (int)ceilf( ( mx[1] - mn[1] ) / voxel ) + 1,
            zr = (int)ceilf( ( mx[2] - mn[2] ) / voxel ) + 1; /* (:67-69) */
  const size_t n_cells = (size_t)xr * yr * zr;
  uint8_t* ga = (uint8_t*)malloc( n_cells ), *gb = (uint8_t*)malloc( n_cells );
  const int32_t ca = orc_occupancy( pos1, n1, pose_a, mn, voxel, xr, yr, zr, inside, ga );
  const int32_t cb = orc_occupancy( pos1, n1, pose_b, mn, voxel, xr, yr, zr, inside, gb );
  int32_t both = 0;
  for( size_t c = 0; c < n_cells; ++c ) { both += ( ga[c] == 1 || ga[c] == 2 ) && ( gb[c] == 1 || gb[c] == 2 ); }
  free( ga ); free( gb );
  const int32_t denom = normalize_by_smaller ? ( ca < cb ? ca : cb ) : ( ca > cb ? ca : cb );
  return denom > 0 ? (float)both / (float)denom : 1.0f; /* (:350-357) */
}

/* mgs_non_maxima_suppresion for one object (pose_proposal.cpp:371-452): keep[i] = 1 for the survivors.
   proposals: n x 17 floats (xform column-major + score); centroid = rs_pointcloud_centroid( shape, 0 ). */
void orc_nms( const float* pos3, int32_t n3, const float* pos1, int32_t n1, const float* centroid, const float* proposals, int32_t n,
              float dist_threshold, uint8_t* keep )
{
  uint8_t* mark = (uint8_t*)calloc( n > 0 ? n : 1, 1 ); /* 0 unmarked, 1 keep, 2 discard */
  int32_t marked = 0;
  while( marked != n )
  {
    int32_t best = -1; float best_score = -1e9f;
    for( int32_t i = 0; i < n; ++i )
    {
      if( mark[i] == 0 && proposals[17 * (size_t)i + 16] > best_score ) { best_score = proposals[17 * (size_t)i + 16]; best = i; }
    }
    if( best < 0 ) { break; } /* only NaN scores left: the reference would index [-1]; nothing sensible to keep */
    mark[best] = 1; marked++;
    float p1[3];
    xf_apply( proposals + 17 * (size_t)best, centroid, 1, p1 );
    for( int32_t i = 0; i < n; ++i )
    {
      if( mark[i] != 0 ) { continue; }
      float p2[3];
      xf_apply( proposals + 17 * (size_t)i, centroid, 1, p2 );
      const float dx = p1[0] - p2[0], dy = p1[1] - p2[1], dz = p1[2] - p2[2];
      const float dist = (float)sqrt( (double)( dx * dx + dy * dy + dz * dz ) ); /* msh_vec3_norm (msh_vec_math.h:988-991) */
      const float overlap = orc_overlap_factor( pos3, n3, pos1, n1, proposals + 17 * (size_t)best, proposals + 17 * (size_t)i, 0.1f, 1, 0 );
      if( overlap > 0.5f || dist < dist_threshold || proposals[17 * (size_t)i + 16] < 0.01f ) { mark[i] = 2; marked++; }
    }
  }
  for( int32_t i = 0; i < n; ++i ) { keep[i] = mark[i] == 1; }
  free( mark );
}

/* ------------------------------------------------------------------------------------------------ level building */
/* rs_pointcloud__compute_level_poisson (lib/rs/rs_pointcloud.h:984-1037): greedy Poisson-disk subsampling of level 0.
   A grid is built over level 0 with radius 2.5 * voxel (:989-990); then, in ascending index order, the first point not
   yet marked becomes a sample and every point its radius search returns (r = voxel, the max_n_neigh NEAREST points
   with dist^2 < r^2, the sample itself included, :1017-1035) is marked.  max_n_neigh = (size_t)(1024 * (level / 4.0f)),
   256 if that is 0 (:994-995).  out_idx: the samples' level-0 indices, ascending (the level's arrays are copies of those
   rows, :1077-1086).  Returns the number of samples. */
int32_t orc_poisson_level( const float* pos0, int32_t n, float voxel, int32_t level, int32_t* out_idx )
{
  if( n <= 0 ) { return 0; }
  orc_grid_t* g = orc_grid_build( pos0, n, 2.5f * voxel );
  size_t k = (size_t)( 1024 * ( ( level ) / (float)( 5 - 1 ) ) );
  if( !k ) { k = 256; }
  uint8_t* unmarked = (uint8_t*)malloc( (size_t)n );
  float* d2 = (float*)malloc( k * sizeof( float ) );
  int32_t* id = (int32_t*)malloc( k * sizeof( int32_t ) );
  memset( unmarked, 1, (size_t)n );
  double r = voxel;
  float r2f = (float)( r * r );
  size_t n_marked = 0;
  int32_t n_out = 0, cur = 0;
  while( n_marked < (size_t)n )
  {
    while( !unmarked[cur] ) { cur++; }
    out_idx[n_out++] = cur;
    size_t c = orc_radius_query( g, pos0 + 3 * (size_t)cur, r, r2f, k, d2, id );
    for( size_t i = 0; i < c; ++i )
    {
      if( unmarked[id[i]] ) { n_marked++; }
      unmarked[id[i]] = 0;
    }
  }
  free( unmarked ); free( d2 ); free( id );
  orc_grid_free( g );
  return n_out;
}


/* ------------------------------------------------------------------------------------------------ coverage term (8 f3) */
/* isect_grid3d_init (lib/rs/intersect.h:57-75): bbox fattened by 0.3, resolution ceilf( extent / voxel ) + 1 per axis */
int32_t orc_cov_grid( const float* bbox_min, const float* bbox_max, float voxel, int32_t* res, float* origin )
{
  for( int a = 0; a < 3; ++a )
  {
    const float mn = bbox_min[a] - 0.3f, mx = bbox_max[a] + 0.3f;
    origin[a] = mn;
    res[a] = (int32_t)ceilf( ( mx - mn ) / voxel ) + 1;
  }
  return res[0] * res[1] * res[2];
}

/* rsao_rasterize_scene_to_grid / rsao__rasterize_arrangement_to_grid (apps/segment_transfer/arrangement_optimization.cpp:
   1064-1106): every point (moved by `pose` first when given, msh_mat4_vec3_mul) lights the cell
   floorf( (p - origin) * (1.0f / voxel) ) per axis, stored at (y * zr + z) * xr + x (intersect.h:102-116); points outside
   the grid are ignored.  `grid` is not cleared. */
void orc_cov_rasterize( const float* pts, int32_t n, const float* pose, const float* origin, const int32_t* res, float voxel, uint8_t* grid )
{
  const float inv = 1.0f / voxel;
  for( int32_t i = 0; i < n; ++i )
  {
    float p[3] = { pts[3 * (size_t)i], pts[3 * (size_t)i + 1], pts[3 * (size_t)i + 2] };
    if( pose ) { float q[3]; xf_apply( pose, p, 1, q ); p[0] = q[0]; p[1] = q[1]; p[2] = q[2]; }
    const int32_t x = (int32_t)floorf( ( p[0] - origin[0] ) * inv ), y = (int32_t)floorf( ( p[1] - origin[1] ) * inv ),
                  z = (int32_t)floorf( ( p[2] - origin[2] ) * inv );
    if( x < 0 || x >= res[0] || y < 0 || y >= res[1] || z < 0 || z >= res[2] ) { continue; }
    grid[( (size_t)y * res[2] + z ) * res[0] + x] = 1;
  }
}

/* rsao__compute_scene_coverage_score (:343-373): cells lit in both grids / cells lit in the scan's grid, 0 when none */
float orc_cov_score( const uint8_t* scn, const uint8_t* arr, int32_t n_cells )
{
  int32_t agree = 0, valid = 0;
  for( int32_t i = 0; i < n_cells; ++i )
  {
    if( scn[i] > 0 ) { valid++; }
    if( scn[i] > 0 && arr[i] > 0 ) { agree++; }
  }
  return valid == 0 ? 0.0f : (float)agree / (float)valid;
}

/* ------------------------------------------------------------------------------------------------ plane detection */
/* evaluate_plane_model (lib/rs/rs_pointcloud_filters.cpp:117-134) for a list of candidate planes {center xyz, normal xyz}:
   counts[p] = number of points with weights[i] > 0.01 (active[i] != 0 here) and |n . (pt - center)| < dist_threshold, the
   distance as the float expression n.x*d.x + n.y*d.y + n.z*d.z of msh_vec3_dot( n, msh_vec3_sub( pt, center ) ) */
void orc_plane_inlier_counts( const float* pts, const uint8_t* active, int32_t n, const float* planes, int32_t n_planes, float dist_threshold,
                              int32_t* counts )
{
  for( int32_t p = 0; p < n_planes; ++p )
  {
    const float* c = planes + 6 * (size_t)p;
    const float* nr = c + 3;
    int32_t cnt = 0;
    for( int32_t i = 0; i < n; ++i )
    {
      if( !active[i] ) { continue; }
      const float dx = pts[3 * (size_t)i] - c[0], dy = pts[3 * (size_t)i + 1] - c[1], dz = pts[3 * (size_t)i + 2] - c[2];
      float d = nr[0] * dx + nr[1] * dy + nr[2] * dz;
      d = d < 0 ? -d : d;
      if( d < dist_threshold ) { cnt++; }
    }
    counts[p] = cnt;
  }
}
