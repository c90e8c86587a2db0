// Peer exchange of the pose-sharded search over NVLink / NVSwitch without an SM-resident collective kernel.
//
// What travels: the per-object survivor lists of mgs_propose_poses (reference apps/pose_proposal/pose_proposal.cpp:348-359:
// {xform, score} records + their dense pose ids), and later the ICP-refined rows (apps/pose_proposal/main.cpp:195-201) - a few KB
// per step.  NCCL's all-gather for them needs most of an SM per channel and, issued mid-step, waits behind the thousands of
// pending blocks of the dense search (14-18 ms, profiles/step_trace_r01_n2.txt).  Here every rank owns a receive area in
// its HBM that all peers map through CUDA IPC; an all-gather is then
//     my bytes -> slot [round % RING][my rank] of EVERY rank's area     (cudaMemcpyAsync peer copies: copy engines over NVLink)
//     the round number -> flag [round % RING][my rank] of every rank     (a second, stream-ordered 4-byte peer copy)
//     one 32-thread kernel on the receiver that spins until all flags of the round carry its number (with a time-out)
// so the data path uses copy engines only and one warp for the wait.  One process per GPU; the IPC handles are exchanged
// once at start-up by the caller (torch.distributed all_gather_object in rescan_b200/peerx.py) - set-up, not data path.
//
// Ring: a rank can be at most one round ahead of the slowest rank when it SENDS (it cannot finish round s + 1 without
// that rank's round-s + 1 data, which is sent after that rank read round s), so two buffers would do; RING = 4.
#include "rsgpu_internal.cuh"
#include <cstring>
#include <vector>
#include <mutex>

using namespace rs;

namespace
{
constexpr int RING = 4;
constexpr int MAX_WORLD = 16;

struct PeerState
{
  int rank = -1, world = 0;
  size_t slot_bytes = 0;            // capacity of one rank's slot in one ring buffer
  unsigned char* area = nullptr;    // [RING][world][slot_bytes] received payloads   (cudaMalloc: IPC needs a plain allocation)
  uint32_t* flags = nullptr;        // [RING][world] round numbers
  unsigned char* send = nullptr;    // [RING][slot_bytes] my payload staged in HBM
  uint32_t* seqs = nullptr;         // [RING] the round number as device words (source of the flag copies)
  int* status = nullptr;            // wait kernel outcome
  unsigned char* peer_area[MAX_WORLD] = {};
  uint32_t* peer_flags[MAX_WORLD] = {};
  void* peer_base[MAX_WORLD] = {};  // what cudaIpcOpenMemHandle returned (to close it)
  unsigned char* base = nullptr;    // one allocation: area | flags
  size_t flags_off = 0;
  cudaStream_t st = nullptr;
  unsigned char* h_send = nullptr;  // pinned staging [RING][slot_bytes]
  unsigned char* h_recv = nullptr;  // pinned [world][slot_bytes]
  uint32_t* h_seq = nullptr;        // pinned [RING]
  int* h_status = nullptr;
  uint32_t round = 0;
  bool open = false;
};
PeerState g_peer;
std::mutex g_peer_mu;

// lane r < world waits until flags[r] == want; status 0 = ok, 1 = timed out (peer died / never sent)
__global__ void __launch_bounds__( 32 ) peer_wait_kernel( const volatile uint32_t* flags, uint32_t want, int world, long long timeout_cycles,
                                                          int* __restrict__ status )
{
  const int lane = threadIdx.x;
  bool ok = true;
  if( lane < world )
  {
    const long long t0 = clock64();
    while( flags[lane] != want )
    {
      if( clock64() - t0 > timeout_cycles ) { ok = false; break; }
      __nanosleep( 200 );
    }
  }
  __threadfence_system();
  const unsigned all = __ballot_sync( RS_FULL, ok );
  if( lane == 0 ) { *status = all == RS_FULL ? 0 : 1; }
}
} // namespace

extern "C" {

int rsgpu_peer_handle_bytes( void ) { return (int)sizeof( cudaIpcMemHandle_t ); }

/* allocate this rank's receive area and return its IPC handle (rsgpu_peer_handle_bytes() bytes) */
int rsgpu_peer_init( int32_t rank, int32_t world, int64_t slot_bytes, void* handle_out )
{
  if( rank < 0 || world < 1 || world > MAX_WORLD || rank >= world || slot_bytes <= 0 || !handle_out )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_peer_init: bad argument (world <= 16)" );
  }
  RS_TRY( ensure_device() );
  std::lock_guard<std::mutex> lk( g_peer_mu );
  PeerState& p = g_peer;
  if( p.base ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_init: already initialised (rsgpu_peer_close first)" ); }
  p.rank = rank; p.world = world;
  p.slot_bytes = ( (size_t)slot_bytes + 255 ) / 256 * 256;
  const size_t area_bytes = (size_t)RING * world * p.slot_bytes;
  p.flags_off = area_bytes;
  RS_CUDA( cudaMalloc( (void**)&p.base, area_bytes + sizeof( uint32_t ) * RING * MAX_WORLD ) );
  RS_CUDA( cudaMemset( p.base, 0, area_bytes + sizeof( uint32_t ) * RING * MAX_WORLD ) );
  p.area = p.base; p.flags = (uint32_t*)( p.base + p.flags_off );
  RS_CUDA( cudaMalloc( (void**)&p.send, (size_t)RING * p.slot_bytes ) );
  RS_CUDA( cudaMalloc( (void**)&p.seqs, sizeof( uint32_t ) * RING ) );
  RS_CUDA( cudaMalloc( (void**)&p.status, sizeof( int ) ) );
  RS_CUDA( cudaHostAlloc( (void**)&p.h_send, (size_t)RING * p.slot_bytes, cudaHostAllocDefault ) );
  RS_CUDA( cudaHostAlloc( (void**)&p.h_recv, (size_t)world * p.slot_bytes, cudaHostAllocDefault ) );
  RS_CUDA( cudaHostAlloc( (void**)&p.h_seq, sizeof( uint32_t ) * RING, cudaHostAllocDefault ) );
  RS_CUDA( cudaHostAlloc( (void**)&p.h_status, sizeof( int ), cudaHostAllocDefault ) );
  int least = 0, greatest = 0;
  if( cudaDeviceGetStreamPriorityRange( &least, &greatest ) != cudaSuccess ) { cudaGetLastError(); least = greatest = 0; }
  RS_CUDA( cudaStreamCreateWithPriority( &p.st, cudaStreamNonBlocking, greatest ) );
  cudaIpcMemHandle_t h;
  RS_CUDA( cudaIpcGetMemHandle( &h, p.base ) );
  memcpy( handle_out, &h, sizeof( h ) );
  RS_CUDA( cudaDeviceSynchronize() );
  p.round = 0;
  return RSGPU_OK;
}

/* handles = world x rsgpu_peer_handle_bytes() bytes, rank-major (this rank's own entry is ignored) */
int rsgpu_peer_open( const void* handles )
{
  RS_TRY( ensure_device() );
  std::lock_guard<std::mutex> lk( g_peer_mu );
  PeerState& p = g_peer;
  if( !p.base || !handles ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_open: rsgpu_peer_init first" ); }
  for( int r = 0; r < p.world; ++r )
  {
    if( r == p.rank ) { p.peer_area[r] = p.area; p.peer_flags[r] = p.flags; continue; }
    cudaIpcMemHandle_t h;
    memcpy( &h, (const unsigned char*)handles + sizeof( h ) * (size_t)r, sizeof( h ) );
    void* base = nullptr;
    RS_CUDA( cudaIpcOpenMemHandle( &base, h, cudaIpcMemLazyEnablePeerAccess ) );
    p.peer_base[r] = base;
    p.peer_area[r] = (unsigned char*)base;
    p.peer_flags[r] = (uint32_t*)( (unsigned char*)base + p.flags_off );
  }
  p.open = true;
  return RSGPU_OK;
}

/* all-gather of `nbytes` (the same on every rank, <= the slot size) from host memory `send` into host memory `recv`
   (world x nbytes, rank-major).  Collective: every rank calls it the same number of times in the same order. */
int rsgpu_peer_allgather( const void* send, int64_t nbytes, void* recv, double timeout_s )
{
  RS_TRY( ensure_device() );
  std::lock_guard<std::mutex> lk( g_peer_mu );
  PeerState& p = g_peer;
  if( !p.open ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_allgather: rsgpu_peer_open first" ); }
  if( nbytes < 0 || (size_t)nbytes > p.slot_bytes || ( nbytes > 0 && ( !send || !recv ) ) )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_peer_allgather: payload larger than the slot size given to rsgpu_peer_init" );
  }
  RS_CUDA( cudaSetDevice( rt().device ) );
  const uint32_t round = ++p.round;
  const int buf = (int)( round % RING );
  const size_t nb = (size_t)nbytes;
  unsigned char* hs = p.h_send + (size_t)buf * p.slot_bytes;
  unsigned char* ds = p.send + (size_t)buf * p.slot_bytes;
  if( nb ) { memcpy( hs, send, nb ); }
  p.h_seq[buf] = round;
  if( nb ) { RS_CUDA( cudaMemcpyAsync( ds, hs, nb, cudaMemcpyHostToDevice, p.st ) ); }
  RS_CUDA( cudaMemcpyAsync( p.seqs + buf, p.h_seq + buf, sizeof( uint32_t ), cudaMemcpyHostToDevice, p.st ) );
  // payload to every rank's slot [buf][my rank] (NVLink, copy engines), then - stream-ordered behind it - the flag
  for( int k = 0; k < p.world; ++k )
  {
    const int r = ( p.rank + k ) % p.world; // start with my own area, spread the peers
    unsigned char* dst = p.peer_area[r] + ( (size_t)buf * p.world + p.rank ) * p.slot_bytes;
    if( nb ) { RS_CUDA( cudaMemcpyAsync( dst, ds, nb, cudaMemcpyDeviceToDevice, p.st ) ); }
  }
  for( int k = 0; k < p.world; ++k )
  {
    const int r = ( p.rank + k ) % p.world;
    RS_CUDA( cudaMemcpyAsync( p.peer_flags[r] + (size_t)buf * MAX_WORLD + p.rank, p.seqs + buf, sizeof( uint32_t ), cudaMemcpyDeviceToDevice, p.st ) );
  }
  int clock_khz = 1965000;
  cudaDeviceGetAttribute( &clock_khz, cudaDevAttrClockRate, rt().device );
  const long long timeout_cycles = (long long)( ( timeout_s > 0 ? timeout_s : 30.0 ) * 1e3 * (double)clock_khz );
  peer_wait_kernel<<<1, 32, 0, p.st>>>( p.flags + (size_t)buf * MAX_WORLD, round, p.world, timeout_cycles, p.status );
  RS_CHECK_LAUNCH();
  RS_CUDA( cudaMemcpyAsync( p.h_status, p.status, sizeof( int ), cudaMemcpyDeviceToHost, p.st ) );
  for( int r = 0; r < p.world && nb; ++r )
  {
    RS_CUDA( cudaMemcpyAsync( p.h_recv + (size_t)r * nb, p.area + ( (size_t)buf * p.world + r ) * p.slot_bytes, nb, cudaMemcpyDeviceToHost, p.st ) );
  }
  RS_CUDA( rs::stream_sync( p.st ) );
  if( *p.h_status != 0 ) { return fail( RSGPU_ERR_CUDA, "rsgpu_peer_allgather: timed out waiting for a peer's payload" ); }
  if( nb ) { memcpy( recv, p.h_recv, nb * (size_t)p.world ); }
  return RSGPU_OK;
}

int rsgpu_peer_close( void )
{
  std::lock_guard<std::mutex> lk( g_peer_mu );
  PeerState& p = g_peer;
  if( !p.base ) { return RSGPU_OK; }
  cudaSetDevice( rt().device );
  if( p.st ) { cudaStreamSynchronize( p.st ); }
  for( int r = 0; r < p.world; ++r ) { if( p.peer_base[r] ) { cudaIpcCloseMemHandle( p.peer_base[r] ); } }
  cudaFree( p.base ); cudaFree( p.send ); cudaFree( p.seqs ); cudaFree( p.status );
  cudaFreeHost( p.h_send ); cudaFreeHost( p.h_recv ); cudaFreeHost( p.h_seq ); cudaFreeHost( p.h_status );
  if( p.st ) { cudaStreamDestroy( p.st ); }
  cudaGetLastError();
  p = PeerState();
  return RSGPU_OK;
}

} // extern "C"
