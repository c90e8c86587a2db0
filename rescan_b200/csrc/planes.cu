// RANSAC plane scoring of the scan's wall / floor detector on the GPU (the stage that bounds the segment_transfer drop-in
// once the path of SURVEY.md 8 is on the device).
//
// rspf__detect_walls / rspf__detect_floor (reference lib/rs/rs_pointcloud_filters.cpp:137-253) draw 5 000 / 2 500 point
// triples per round and count, for every candidate plane, the still-unexplained scan points closer than dist_threshold
// (evaluate_plane_model :117-134): P x N point-plane tests per round, on one thread.  The triples do not depend on the
// counts, so a round is ONE launch over all its candidates; the caller then takes the first maximum, as the reference's
// strict '>' does.  Integer results; the distance is the reference's float expression
// |n.x*d.x + n.y*d.y + n.z*d.z|, d = pt - center, unfused.
//
//   rsgpu_plane_inlier_counts   candidate planes x points -> inlier counts      evaluate_plane_model :117-134
#include "rsgpu_internal.cuh"

using namespace rs;

namespace
{
constexpr int TP = 16; // planes per block: a point is loaded once and tested against 16 planes (held in registers after the first read)

__global__ void __launch_bounds__( 256, 2 ) plane_inlier_kernel( const float* __restrict__ pts, const unsigned char* __restrict__ active, int n,
                                                              const float* __restrict__ planes, int n_planes, float thr, int* __restrict__ counts )
{
  __shared__ float sp[TP][6];
  __shared__ int sc[TP];
  const int p0 = blockIdx.x * TP;
  for( int i = threadIdx.x; i < TP * 6; i += blockDim.x )
  {
    const int p = p0 + i / 6;
    sp[i / 6][i % 6] = p < n_planes ? planes[(size_t)p * 6 + i % 6] : 0.f;
  }
  if( threadIdx.x < TP ) { sc[threadIdx.x] = 0; }
  __syncthreads();
  int cnt[TP];
#pragma unroll
  for( int k = 0; k < TP; ++k ) { cnt[k] = 0; }
  for( int i = blockIdx.y * blockDim.x + threadIdx.x; i < n; i += gridDim.y * blockDim.x )
  {
    if( !active[i] ) { continue; } // weights[i] > 0.01 (:127)
    const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
#pragma unroll
    for( int k = 0; k < TP; ++k )
    {
      const float dx = __fsub_rn( x, sp[k][0] ), dy = __fsub_rn( y, sp[k][1] ), dz = __fsub_rn( z, sp[k][2] );
      const float d = __fadd_rn( __fadd_rn( __fmul_rn( sp[k][3], dx ), __fmul_rn( sp[k][4], dy ) ), __fmul_rn( sp[k][5], dz ) );
      cnt[k] += ( ( d < 0.f ? -d : d ) < thr ) ? 1 : 0; // msh_abs, then '<': a NaN distance (degenerate triple) counts nothing
    }
  }
#pragma unroll
  for( int k = 0; k < TP; ++k )
  {
    int c = cnt[k];
    for( int o = 16; o > 0; o >>= 1 ) { c += __shfl_down_sync( 0xffffffffu, c, o ); }
    if( ( threadIdx.x & 31 ) == 0 && c ) { atomicAdd( &sc[k], c ); }
  }
  __syncthreads();
  if( threadIdx.x < TP && p0 + threadIdx.x < n_planes && sc[threadIdx.x] ) { atomicAdd( counts + p0 + threadIdx.x, sc[threadIdx.x] ); }
}
} // namespace

extern "C" {

int rsgpu_plane_inlier_counts( const float* pts, const uint8_t* active, int32_t n_pts, const float* planes, int32_t n_planes, float dist_threshold,
                               int32_t* counts )
{
  if( n_pts < 0 || n_planes < 0 || ( n_pts > 0 && ( !pts || !active ) ) || ( n_planes > 0 && ( !planes || !counts ) ) )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_plane_inlier_counts: bad argument" );
  }
  RS_TRY( ensure_device() );
  if( n_planes == 0 ) { return RSGPU_OK; }
  if( n_pts == 0 ) { for( int32_t i = 0; i < n_planes; ++i ) { counts[i] = 0; } return RSGPU_OK; }
  cudaStream_t st = rt().stream;
  DevBuf<float> d_pts, d_planes; DevBuf<unsigned char> d_active; DevBuf<int> d_counts;
  RS_CUDA( d_pts.alloc( 3 * (size_t)n_pts ) ); RS_CUDA( d_active.alloc( n_pts ) ); RS_CUDA( d_planes.alloc( 6 * (size_t)n_planes ) );
  RS_CUDA( d_counts.alloc( n_planes ) );
  RS_CUDA( cudaMemcpyAsync( d_pts.p, pts, sizeof( float ) * 3 * (size_t)n_pts, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( d_active.p, active, (size_t)n_pts, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemcpyAsync( d_planes.p, planes, sizeof( float ) * 6 * (size_t)n_planes, cudaMemcpyHostToDevice, st ) );
  RS_CUDA( cudaMemsetAsync( d_counts.p, 0, sizeof( int ) * (size_t)n_planes, st ) );
  {
    ProfScope prof( "planes" );
    const int gx = ( n_planes + TP - 1 ) / TP;
    // enough point slices for a few blocks per SM (148 SMs), each slice at least one block-width of points
    int gy = ( 148 * 4 + gx - 1 ) / gx;
    const int max_gy = ( n_pts + 255 ) / 256;
    if( gy > max_gy ) { gy = max_gy; }
    if( gy < 1 ) { gy = 1; }
    plane_inlier_kernel<<<dim3( (unsigned)gx, (unsigned)gy ), 256, 0, st>>>( d_pts.p, d_active.p, n_pts, d_planes.p, n_planes, dist_threshold, d_counts.p );
    RS_CHECK_LAUNCH();
  }
  RS_CUDA( cudaMemcpyAsync( counts, d_counts.p, sizeof( int ) * (size_t)n_planes, cudaMemcpyDeviceToHost, st ) );
  RS_CUDA( rs::stream_sync( st ) );
  return RSGPU_OK;
}

} // extern "C"
