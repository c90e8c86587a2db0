// Peer exchange of the pose-sharded search over NVLink / NVSwitch without an SM-resident collective kernel.
//
// What travels: the per-object survivor lists of mgs_propose_poses (reference apps/pose_proposal/pose_proposal.cpp:348-359:
// {xform, score} records + their dense pose ids), and later the ICP-refined rows (apps/pose_proposal/main.cpp:195-201) - a few KB
// per object.  NCCL's all-gather for them needs most of an SM per channel and, issued mid-step, waits behind the thousands
// of pending blocks of the dense search (14-18 ms, profiles/step_trace_r01_n2.txt); and a collective imposes ONE order of
// exchanges on all ranks, which serialises the objects' chains.  Here every rank owns a receive area in its HBM that all
// peers map through CUDA IPC, addressed by (slot, rank): an all-gather of slot s is
//     my bytes -> [s][my rank] of EVERY rank's area       (cudaMemcpyAsync peer copies: copy engines over NVLink, no SM)
//     host wait for those copies, then the sequence number -> flag [s][my rank] of every rank, in SHARED HOST MEMORY
//     the receiver's host thread polls its own flags of slot s, then copies the slot out of its HBM
// Exchanges of DIFFERENT slots share nothing: every object's chain (search -> exchange -> NMS -> refine -> exchange)
// runs on its own host thread / stream and meets its peers whenever they get there.
//
// Why the flags are not in HBM with a kernel waiting on them (the first version): with one waiting kernel per chain in
// flight, two ranks dead-lock through the GPU's hardware queues - the streams of a process are multiplexed onto a few
// channels (CUDA_DEVICE_MAX_CONNECTIONS), so rank A's copy for object j can sit in a channel behind A's spinning wait for
// object k while rank B's copy for object k sits behind B's wait for object j (observed: both ranks time out in the row
// exchange of C3 with 8 lanes).  With the flags in host memory NOTHING ever waits on the device: the data plane is copy
// engines over NVLink, the control plane is one 4-byte store per peer.
//
// One process per GPU on one host; the handles (CUDA IPC handle of the area + name of the flag segment) are exchanged once
// at start-up by the caller (torch.distributed all_gather_object in rescan_b200/peerx.py) - set-up, not data path.
//
// Re-use of a slot: the caller passes a sequence number that grows by one per use of the slot.  Slots are double-buffered
// by the parity of that number: a rank can send use q + 2 only after it received every peer's part of use q + 1, which a
// peer sends only after it has read use q out of its area - so a write never lands on data still being read.
#include "rsgpu_internal.cuh"
#include <algorithm>
#include <atomic>
#include <string>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

using namespace rs;

namespace
{
constexpr int MAX_WORLD = 16;
constexpr int NAME_BYTES = 64;

struct PeerHandle // what the ranks hand round at set-up
{
  cudaIpcMemHandle_t mem;
  char flags_name[NAME_BYTES];
};

struct PeerState
{
  int rank = -1, world = 0, n_slots = 0;
  size_t slot_bytes = 0;            // capacity of one rank's payload in one slot
  unsigned char* base = nullptr;    // one IPC allocation: area [2][n_slots][world][slot_bytes]
  unsigned char* send = nullptr;    // [2][n_slots][slot_bytes] my payload staged in HBM
  unsigned char* peer_area[MAX_WORLD] = {};
  void* peer_base[MAX_WORLD] = {};  // what cudaIpcOpenMemHandle returned (to close it)
  uint32_t* peer_flags[MAX_WORLD] = {}; // every rank's flag segment [2][n_slots][MAX_WORLD], mapped from shared host memory
  size_t flag_bytes = 0;
  char flags_name[NAME_BYTES] = {};
  unsigned char* h_send = nullptr;  // pinned [2][n_slots][slot_bytes]
  unsigned char* h_recv = nullptr;  // pinned [2][n_slots][world][slot_bytes]
  cudaStream_t st = nullptr;        // for callers without a lane
  bool open = false;
};
PeerState g_peer;
std::mutex g_peer_mu; // set-up / tear-down only; the exchanges of different slots share nothing

void unmap_flags( PeerState& p )
{
  for( int r = 0; r < MAX_WORLD; ++r )
  {
    if( p.peer_flags[r] ) { munmap( p.peer_flags[r], p.flag_bytes ); p.peer_flags[r] = nullptr; }
  }
  if( p.flags_name[0] ) { shm_unlink( p.flags_name ); p.flags_name[0] = 0; }
}
} // namespace

extern "C" {

int rsgpu_peer_handle_bytes( void ) { return (int)sizeof( PeerHandle ); }

/* allocate this rank's receive area + flag segment and return their handle (rsgpu_peer_handle_bytes() bytes) */
int rsgpu_peer_init( int32_t rank, int32_t world, int32_t n_slots, int64_t slot_bytes, void* handle_out )
{
  if( rank < 0 || world < 1 || world > MAX_WORLD || rank >= world || slot_bytes <= 0 || n_slots < 1 || !handle_out )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_peer_init: bad argument (world <= 16)" );
  }
  RS_TRY( ensure_device() );
  std::lock_guard<std::mutex> lk( g_peer_mu );
  PeerState& p = g_peer;
  if( p.base ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_init: already initialised (rsgpu_peer_close first)" ); }
  p.rank = rank; p.world = world; p.n_slots = n_slots;
  p.slot_bytes = ( (size_t)slot_bytes + 255 ) / 256 * 256;
  const size_t area_bytes = (size_t)2 * n_slots * world * p.slot_bytes;
  p.flag_bytes = sizeof( uint32_t ) * 2 * ( 2 * (size_t)n_slots * MAX_WORLD ); // flags [2][n_slots][MAX_WORLD] | lengths, same shape
  // the flag segment: POSIX shared memory, zero-filled, mapped by every rank of the host
  snprintf( p.flags_name, NAME_BYTES, "/rsgpu_peer_%d_%d_%llx", (int)getpid(), (int)rank,
            (unsigned long long)std::chrono::steady_clock::now().time_since_epoch().count() );
  const int fd = shm_open( p.flags_name, O_CREAT | O_EXCL | O_RDWR, 0600 );
  if( fd < 0 ) { p.flags_name[0] = 0; return fail( RSGPU_ERR_INVALID, "rsgpu_peer_init: shm_open failed (no /dev/shm?)" ); }
  if( ftruncate( fd, (off_t)p.flag_bytes ) != 0 ) { close( fd ); unmap_flags( p ); return fail( RSGPU_ERR_OOM, "rsgpu_peer_init: ftruncate of the flag segment failed" ); }
  void* fl = mmap( nullptr, p.flag_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0 );
  close( fd );
  if( fl == MAP_FAILED ) { unmap_flags( p ); return fail( RSGPU_ERR_OOM, "rsgpu_peer_init: mmap of the flag segment failed" ); }
  p.peer_flags[rank] = (uint32_t*)fl;
  RS_CUDA( cudaMalloc( (void**)&p.base, area_bytes ) ); // cudaMalloc: IPC needs a plain allocation
  RS_CUDA( cudaMemset( p.base, 0, area_bytes ) );
  RS_CUDA( cudaMalloc( (void**)&p.send, (size_t)2 * n_slots * p.slot_bytes ) );
  RS_CUDA( cudaHostAlloc( (void**)&p.h_send, (size_t)2 * n_slots * p.slot_bytes, cudaHostAllocDefault ) );
  RS_CUDA( cudaHostAlloc( (void**)&p.h_recv, (size_t)2 * n_slots * world * p.slot_bytes, cudaHostAllocDefault ) );
  int least = 0, greatest = 0;
  if( cudaDeviceGetStreamPriorityRange( &least, &greatest ) != cudaSuccess ) { cudaGetLastError(); least = greatest = 0; }
  RS_CUDA( cudaStreamCreateWithPriority( &p.st, cudaStreamNonBlocking, greatest ) );
  PeerHandle h;
  memset( &h, 0, sizeof( h ) );
  RS_CUDA( cudaIpcGetMemHandle( &h.mem, p.base ) );
  memcpy( h.flags_name, p.flags_name, NAME_BYTES );
  memcpy( handle_out, &h, sizeof( h ) );
  RS_CUDA( cudaDeviceSynchronize() );
  return RSGPU_OK;
}

/* map every peer's area and flag segment; handles = world x rsgpu_peer_handle_bytes() bytes in rank order */
int rsgpu_peer_open( const void* handles )
{
  RS_TRY( ensure_device() );
  std::lock_guard<std::mutex> lk( g_peer_mu );
  PeerState& p = g_peer;
  if( !p.base || !handles ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_open: rsgpu_peer_init first" ); }
  for( int r = 0; r < p.world; ++r )
  {
    if( r == p.rank ) { p.peer_area[r] = p.base; continue; }
    PeerHandle h;
    memcpy( &h, (const unsigned char*)handles + sizeof( h ) * (size_t)r, sizeof( h ) );
    h.flags_name[NAME_BYTES - 1] = 0;
    void* base = nullptr;
    RS_CUDA( cudaIpcOpenMemHandle( &base, h.mem, cudaIpcMemLazyEnablePeerAccess ) );
    p.peer_base[r] = base;
    p.peer_area[r] = (unsigned char*)base;
    const int fd = shm_open( h.flags_name, O_RDWR, 0600 );
    if( fd < 0 ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_open: a peer's flag segment cannot be opened (ranks on different hosts?)" ); }
    void* fl = mmap( nullptr, p.flag_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0 );
    close( fd );
    if( fl == MAP_FAILED ) { return fail( RSGPU_ERR_OOM, "rsgpu_peer_open: mmap of a peer's flag segment failed" ); }
    p.peer_flags[r] = (uint32_t*)fl;
  }
  p.open = true;
  return RSGPU_OK;
}

namespace
{
inline uint32_t* flag_of( PeerState& p, int owner_rank, size_t s, int writer ) { return p.peer_flags[owner_rank] + s * MAX_WORLD + writer; }
inline uint32_t* len_of( PeerState& p, int owner_rank, size_t s, int writer )
{
  return p.peer_flags[owner_rank] + (size_t)2 * p.n_slots * MAX_WORLD + s * MAX_WORLD + writer;
}
inline int check_slot( PeerState& p, const char* what, int32_t slot, uint32_t seq )
{
  if( !p.open ) { return fail( RSGPU_ERR_INVALID, std::string( what ) + ": rsgpu_peer_open first" ); }
  if( slot < 0 || slot >= p.n_slots || seq == 0 ) { return fail( RSGPU_ERR_INVALID, std::string( what ) + ": slot out of range / seq 0" ); }
  return RSGPU_OK;
}
} // namespace

/* use `seq` of slot `slot`, sending side: `nbytes` (<= the slot size; may differ from rank to rank) from host memory `send`
   into this rank's row of the slot in the area of every rank whose bit is set in dst_mask, then - once the copies have
   landed - the length and the sequence number into those ranks' flag segments.  Does not wait for anybody. */
int rsgpu_peer_put( int32_t slot, uint32_t seq, uint32_t dst_mask, const void* send, int64_t nbytes )
{
  RS_TRY( ensure_device() );
  PeerState& p = g_peer;
  RS_TRY( check_slot( p, "rsgpu_peer_put", slot, seq ) );
  if( nbytes < 0 || (size_t)nbytes > p.slot_bytes || ( nbytes > 0 && !send ) )
  {
    return fail( RSGPU_ERR_INVALID, "rsgpu_peer_put: payload larger than the slot size given to rsgpu_peer_init" );
  }
  // the calling thread's own stream when it is attached to a lane (so concurrent exchanges do not queue behind one another)
  cudaStream_t st = rt().stream ? rt().stream : p.st;
  const size_t s = (size_t)( seq & 1u ) * p.n_slots + (size_t)slot; // double-buffered by the parity of seq
  const size_t nb = (size_t)nbytes;
  if( nb )
  {
    unsigned char* hs = p.h_send + s * p.slot_bytes;
    unsigned char* ds = p.send + s * p.slot_bytes;
    memcpy( hs, send, nb );
    RS_CUDA( cudaMemcpyAsync( ds, hs, nb, cudaMemcpyHostToDevice, st ) );
    for( int k = 0; k < p.world; ++k ) // payload to [s][my rank] of the destinations' areas (NVLink, copy engines)
    {
      const int r = ( p.rank + k ) % p.world; // start with my own area, spread the peers
      if( !( ( dst_mask >> r ) & 1u ) ) { continue; }
      RS_CUDA( cudaMemcpyAsync( p.peer_area[r] + ( s * p.world + p.rank ) * p.slot_bytes, ds, nb, cudaMemcpyDeviceToDevice, st ) );
    }
    RS_CUDA( rs::stream_sync( st ) ); // the copies have landed ...
  }
  for( int k = 0; k < p.world; ++k )   // ... before any peer can see the flag
  {
    const int r = ( p.rank + k ) % p.world;
    if( !( ( dst_mask >> r ) & 1u ) ) { continue; }
    __atomic_store_n( len_of( p, r, s, p.rank ), (uint32_t)nb, __ATOMIC_RELAXED );
    __atomic_store_n( flag_of( p, r, s, p.rank ), seq, __ATOMIC_RELEASE );
  }
  return RSGPU_OK;
}

/* use `seq` of slot `slot`, receiving side: waits until every rank whose bit is set in src_mask has put its payload, then
   copies the rows into host memory `recv` [world][row_bytes] (rows of ranks outside the mask are left untouched) and their
   lengths into nbytes_out[world] (0 outside the mask).  timeout_s <= 0 means 30 s. */
int rsgpu_peer_get( int32_t slot, uint32_t seq, uint32_t src_mask, void* recv, int64_t row_bytes, int64_t* nbytes_out, double timeout_s )
{
  RS_TRY( ensure_device() );
  PeerState& p = g_peer;
  RS_TRY( check_slot( p, "rsgpu_peer_get", slot, seq ) );
  if( row_bytes < 0 || !nbytes_out || ( row_bytes > 0 && !recv ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_get: bad argument" ); }
  cudaStream_t st = rt().stream ? rt().stream : p.st;
  const size_t s = (size_t)( seq & 1u ) * p.n_slots + (size_t)slot;
  // wait for the flags of this use: spin briefly, then sleep (a peer may still be searching)
  const auto t0 = std::chrono::steady_clock::now();
  const double limit = timeout_s > 0 ? timeout_s : 30.0;
  int have = 0;
  for( ;; )
  {
    while( have < p.world && ( !( ( src_mask >> have ) & 1u ) || __atomic_load_n( flag_of( p, p.rank, s, have ), __ATOMIC_ACQUIRE ) == seq ) ) { ++have; }
    if( have == p.world ) { break; }
    const double waited = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();
    if( waited > limit ) { return fail( RSGPU_ERR_CUDA, "rsgpu_peer_get: timed out waiting for a peer's payload" ); }
    if( waited > 50e-6 ) { std::this_thread::sleep_for( std::chrono::microseconds( waited > 2e-3 ? 100 : 20 ) ); }
  }
  size_t widest = 0;
  for( int r = 0; r < p.world; ++r )
  {
    const size_t len = ( ( src_mask >> r ) & 1u ) ? (size_t)__atomic_load_n( len_of( p, p.rank, s, r ), __ATOMIC_RELAXED ) : 0;
    if( len > (size_t)row_bytes ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_get: a peer's payload is longer than row_bytes" ); }
    nbytes_out[r] = (int64_t)len;
    widest = std::max( widest, len );
  }
  if( widest )
  {
    unsigned char* hr = p.h_recv + s * p.world * p.slot_bytes;
    // the used bytes of every rank's row of the slot, packed, in one copy
    RS_CUDA( cudaMemcpy2DAsync( hr, widest, p.base + s * p.world * p.slot_bytes, p.slot_bytes, widest, (size_t)p.world, cudaMemcpyDeviceToHost, st ) );
    RS_CUDA( rs::stream_sync( st ) );
    for( int r = 0; r < p.world; ++r )
    {
      if( nbytes_out[r] > 0 ) { memcpy( (unsigned char*)recv + (size_t)r * (size_t)row_bytes, hr + (size_t)r * widest, (size_t)nbytes_out[r] ); }
    }
  }
  return RSGPU_OK;
}

/* all-gather of `nbytes` (the same on every rank, <= the slot size) of slot `slot` from host memory `send` into host memory
   `recv` (world x nbytes, rank-major): a put to every rank followed by a get from every rank.  seq = how many times this
   slot has been used before, plus one (identical on every rank).  Exchanges of different slots may run concurrently from
   different host threads, in any order. */
int rsgpu_peer_allgather( int32_t slot, uint32_t seq, const void* send, int64_t nbytes, void* recv, double timeout_s )
{
  PeerState& p = g_peer;
  RS_TRY( check_slot( p, "rsgpu_peer_allgather", slot, seq ) );
  const uint32_t all = p.world >= 32 ? 0xffffffffu : ( ( 1u << p.world ) - 1u );
  RS_TRY( rsgpu_peer_put( slot, seq, all, send, nbytes ) );
  int64_t lens[MAX_WORLD];
  RS_TRY( rsgpu_peer_get( slot, seq, all, recv, nbytes, lens, timeout_s ) );
  for( int r = 0; r < p.world; ++r )
  {
    if( lens[r] != nbytes ) { return fail( RSGPU_ERR_INVALID, "rsgpu_peer_allgather: the ranks sent payloads of different sizes" ); }
  }
  return RSGPU_OK;
}

int rsgpu_peer_close( void )
{
  std::lock_guard<std::mutex> lk( g_peer_mu );
  PeerState& p = g_peer;
  if( !p.base ) { unmap_flags( p ); return RSGPU_OK; }
  cudaSetDevice( rt().device );
  cudaDeviceSynchronize();
  for( int r = 0; r < p.world; ++r ) { if( p.peer_base[r] ) { cudaIpcCloseMemHandle( p.peer_base[r] ); } }
  cudaFree( p.base ); cudaFree( p.send );
  cudaFreeHost( p.h_send ); cudaFreeHost( p.h_recv );
  if( p.st ) { cudaStreamDestroy( p.st ); }
  cudaGetLastError();
  unmap_flags( p );
  p = PeerState();
  return RSGPU_OK;
}

} // extern "C"
