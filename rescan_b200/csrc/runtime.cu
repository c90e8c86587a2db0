// Runtime glue of librsgpu: device/stream selection, error text, per-kernel event profiling, clouds.
#include "rsgpu_internal.cuh"
#include <map>
#include <vector>
#include <mutex>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cctype>
#include <chrono>

namespace rs
{
static thread_local std::string g_err;
static std::mutex g_prof_mu;
struct ProfEntry { double ms = 0; int64_t launches = 0; };
static std::map<std::string, ProfEntry> g_prof;
struct Pending { std::string name; cudaEvent_t a, b; bool own_a = true, own_b = true; };
static std::vector<Pending>* g_pending = nullptr;

static std::atomic<long long> g_launches{ 0 };
void count_launch() { g_launches.fetch_add( 1, std::memory_order_relaxed ); }
long long launches() { return g_launches.load(); }

// tuning / A-B knobs: rsgpu_set_option( "search_impl", "lane" ) or the environment variable RSGPU_SEARCH_IMPL
static std::mutex g_opt_mu;
static std::map<std::string, std::string> g_opts;
std::string option( const char* key )
{
  {
    std::lock_guard<std::mutex> lk( g_opt_mu );
    auto it = g_opts.find( key );
    if( it != g_opts.end() ) { return it->second; }
  }
  std::string env = "RSGPU_";
  for( const char* c = key; *c; ++c ) { env += (char)toupper( (unsigned char)*c ); }
  const char* e = getenv( env.c_str() );
  return e ? std::string( e ) : std::string();
}
void set_option( const char* key, const char* value )
{
  std::lock_guard<std::mutex> lk( g_opt_mu );
  if( value && *value ) { g_opts[key] = value; } else { g_opts.erase( key ); }
}

// Lanes: a host thread that called rsgpu_thread_attach( lane ) enqueues everything on that lane's own stream (and
// helper streams), so independent call chains - one object's propose -> NMS -> ICP -> rescoring next to another's -
// overlap on the device.  Threads that never attached share the process-wide runtime (lane -1).
constexpr int N_LANES = 8;
static Runtime g_rt;
static Runtime g_lane_rt[N_LANES];
static cudaStream_t g_lane_stream[N_LANES];
static cudaStream_t g_lane_bulk[N_LANES];
static thread_local int tl_lane = -1;
static std::mutex g_lane_mu;

Runtime& rt() { return tl_lane >= 0 ? g_lane_rt[tl_lane] : g_rt; }

// Stream priorities inside a lane: the lane's own stream and its helper streams carry the latency-bound launches
// (verification, NMS rounds, ICP iterations) at the highest priority; the long throughput launches (dense pose
// search) go to the lane's bulk stream at the lowest one, so a small launch of one object never queues behind the
// pending blocks of another object's dense search.
static void priority_range( int* least, int* greatest )
{
  *least = 0; *greatest = 0;
  if( cudaDeviceGetStreamPriorityRange( least, greatest ) != cudaSuccess ) { cudaGetLastError(); *least = 0; *greatest = 0; }
}
cudaStream_t bulk_stream() { return tl_lane >= 0 && g_lane_bulk[tl_lane] ? g_lane_bulk[tl_lane] : rt().stream; }

// helper streams for work the library overlaps internally (ICP partitions); created once per lane
int aux_streams( int n, cudaStream_t** out )
{
  static cudaStream_t streams[N_LANES + 1][4];
  static std::mutex mu;
  if( n > 4 ) { return fail( RSGPU_ERR_INVALID, "rsgpu: at most 4 helper streams" ); }
  std::lock_guard<std::mutex> lk( mu );
  cudaStream_t* mine = streams[tl_lane + 1];
  int least, greatest;
  priority_range( &least, &greatest );
  for( int i = 0; i < n; ++i )
  {
    if( !mine[i] ) { RS_CUDA( cudaStreamCreateWithPriority( &mine[i], cudaStreamNonBlocking, greatest ) ); }
  }
  *out = mine;
  return RSGPU_OK;
}

// Host wait for a stream.  cudaStreamSynchronize spins on a CPU core for as long as the stream is busy; with one host
// thread per lane and one process per GPU that is (lanes + 1) x GPUs spinning threads, and once they outnumber the
// cores the threads that hold the next launch get descheduled (measured: 100 ms freezes of a whole rank at 2 ranks x 9
// threads on 16 cores).  Sleeping on a blocking event instead costs a wake-up per wait, which the latency-bound chains
// (an ICP batch syncs every four iterations) feel: + 4.5 ms per C2 step when every wait sleeps.  So a wait the caller
// knows to be long (the dense search, milliseconds) always sleeps, and the short ones spin unless the option
// "sync" = "block" says the host is oversubscribed (rescan_b200/pipeline.py sets it from the core count).
cudaError_t stream_sync( cudaStream_t s, bool long_wait )
{
  static std::atomic<int> mode{ -1 }; // 0 spin on short waits, 1 always sleep
  int m = mode.load( std::memory_order_relaxed );
  const std::string o = option( "sync" );
  const int want = o == "block" ? 1 : 0;
  if( m != want ) { mode.store( want ); m = want; }
  if( !long_wait && m == 0 ) { return cudaStreamSynchronize( s ); }
  static thread_local cudaEvent_t ev = nullptr;
  cudaError_t e;
  if( !ev )
  {
    e = cudaEventCreateWithFlags( &ev, cudaEventBlockingSync | cudaEventDisableTiming );
    if( e != cudaSuccess ) { ev = nullptr; return e; }
  }
  e = cudaEventRecord( ev, s );
  if( e != cudaSuccess ) { return e; }
  const auto t0 = std::chrono::steady_clock::now();
  for( ;; )
  {
    e = cudaEventQuery( ev );
    if( e != cudaErrorNotReady ) { return e; }
    if( std::chrono::steady_clock::now() - t0 > std::chrono::microseconds( 50 ) ) { break; }
  }
  return cudaEventSynchronize( ev );
}

int fail( int code, const std::string& msg )
{
  g_err = msg;
  return code;
}

int cuda_fail( cudaError_t e, const char* what, const char* file, int line )
{
  char buf[1024];
  snprintf( buf, sizeof( buf ), "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString( e ), file, line, what );
  cudaGetLastError();
  int code = RSGPU_ERR_CUDA;
  if( e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ) { code = RSGPU_ERR_NO_DEVICE; }
  if( e == cudaErrorMemoryAllocation ) { code = RSGPU_ERR_OOM; }
  return fail( code, buf );
}

// keep freed blocks in the stream-ordered pool instead of returning them to the driver at every sync
static void configure_pool( int device )
{
  cudaMemPool_t pool;
  if( cudaDeviceGetDefaultMemPool( &pool, device ) == cudaSuccess )
  {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute( pool, cudaMemPoolAttrReleaseThreshold, &keep );
  }
  cudaGetLastError();
}

// The CUDA current device is per HOST THREAD: every thread that enters the library is bound to the process's device once
// (thread_local flag), whether or not it ever attached to a lane; the process-wide part (device count, pool set-up) runs
// once under call_once.  rsgpu_set_device() bumps the generation, so threads re-bind after a device change.
static std::atomic<int> g_dev_generation{ 1 };
static std::once_flag g_dev_once;
static std::atomic<int> g_dev_count{ -1 };
int ensure_device()
{
  static thread_local int bound_generation = 0;
  const int gen = g_dev_generation.load( std::memory_order_acquire );
  if( bound_generation == gen ) { return RSGPU_OK; }
  std::call_once( g_dev_once, []() {
    // One object chain per lane = tens of streams per process; the default of 8 hardware channels makes unrelated streams
    // queue behind one another (a lane's short high-priority launches behind another lane's dense search).  Only effective
    // when this runs before the process creates its CUDA context; never overrides the user's own setting.
    setenv( "CUDA_DEVICE_MAX_CONNECTIONS", "32", 0 );
    int n = 0;
    if( cudaGetDeviceCount( &n ) != cudaSuccess ) { cudaGetLastError(); n = 0; }
    g_dev_count.store( n );
  } );
  if( g_dev_count.load() <= 0 ) { return fail( RSGPU_ERR_NO_DEVICE, "rsgpu: no CUDA device available (there is no CPU fallback)" ); }
  RS_CUDA( cudaSetDevice( g_rt.device ) );
  static std::mutex pool_mu;
  static int pool_device = -1;
  {
    std::lock_guard<std::mutex> lk( pool_mu );
    if( pool_device != g_rt.device ) { configure_pool( g_rt.device ); pool_device = g_rt.device; }
  }
  bound_generation = gen;
  return RSGPU_OK;
}

static void drain_pending()
{
  if( !g_pending ) { return; }
  for( auto& p : *g_pending )
  {
    float ms = 0;
    if( cudaEventSynchronize( p.b ) == cudaSuccess && cudaEventElapsedTime( &ms, p.a, p.b ) == cudaSuccess )
    {
      g_prof[p.name].ms += ms;
      g_prof[p.name].launches += 1;
    }
    if( p.own_a ) { cudaEventDestroy( p.a ); }
    if( p.own_b ) { cudaEventDestroy( p.b ); }
  }
  g_pending->clear();
}

// a - b interval measured by the caller's own events; the events are destroyed when the interval is read
// (own_x false: that event also bounds a LATER pending interval, which destroys it)
void prof_add_pending( const char* name, cudaEvent_t a, cudaEvent_t b, bool own_a, bool own_b )
{
  std::lock_guard<std::mutex> lk( g_prof_mu );
  if( !g_pending ) { g_pending = new std::vector<Pending>(); }
  g_pending->push_back( Pending{ name, a, b, own_a, own_b } );
}

ProfScope::ProfScope( const char* n ) : ProfScope( n, rt().stream ) {}
ProfScope::ProfScope( const char* n, cudaStream_t s ) : name( n ), st( s )
{
  if( !rt().profile ) { return; }
  cudaEventCreate( &a ); cudaEventCreate( &b );
  cudaEventRecord( a, st );
}
ProfScope::~ProfScope()
{
  if( !a ) { return; }
  cudaEventRecord( b, st );
  std::lock_guard<std::mutex> lk( g_prof_mu );
  if( !g_pending ) { g_pending = new std::vector<Pending>(); }
  g_pending->push_back( Pending{ name, a, b, true, true } );
}
} // namespace rs

using namespace rs;

extern "C" {

int rsgpu_set_option( const char* name, const char* value )
{
  if( !name || !*name ) { return fail( RSGPU_ERR_INVALID, "rsgpu_set_option: empty name" ); }
  set_option( name, value );
  return RSGPU_OK;
}

int rsgpu_device_count( void )
{
  int n = 0;
  if( cudaGetDeviceCount( &n ) != cudaSuccess ) { cudaGetLastError(); return 0; }
  return n;
}

int rsgpu_set_device( int device )
{
  int n = rsgpu_device_count();
  if( n <= 0 ) { return fail( RSGPU_ERR_NO_DEVICE, "rsgpu: no CUDA device available (there is no CPU fallback)" ); }
  if( device < 0 || device >= n ) { return fail( RSGPU_ERR_INVALID, "rsgpu_set_device: device index out of range" ); }
  g_rt.device = device; // ONE device per process: lanes and unattached threads all follow it
  for( int i = 0; i < N_LANES; ++i ) { g_lane_rt[i].device = device; }
  g_dev_generation.fetch_add( 1, std::memory_order_acq_rel ); // every thread re-binds on its next call
  return ensure_device();
}

int rsgpu_lane_count( void ) { return N_LANES; }

int rsgpu_thread_attach( int lane )
{
  if( lane >= N_LANES ) { return fail( RSGPU_ERR_INVALID, "rsgpu_thread_attach: lane out of range" ); }
  if( lane < 0 ) { tl_lane = -1; return RSGPU_OK; }
  tl_lane = -1;
  RS_TRY( ensure_device() );
  RS_CUDA( cudaSetDevice( g_rt.device ) ); // the current device is per host thread
  {
    std::lock_guard<std::mutex> lk( g_lane_mu );
    if( !g_lane_stream[lane] )
    {
      int least, greatest;
      priority_range( &least, &greatest );
      RS_CUDA( cudaStreamCreateWithPriority( &g_lane_stream[lane], cudaStreamNonBlocking, greatest ) );
      RS_CUDA( cudaStreamCreateWithPriority( &g_lane_bulk[lane], cudaStreamNonBlocking, least ) );
    }
    g_lane_rt[lane].stream = g_lane_stream[lane];
    g_lane_rt[lane].device = g_rt.device;
    g_lane_rt[lane].profile = g_rt.profile;
  }
  tl_lane = lane;
  return RSGPU_OK;
}

int rsgpu_set_stream( void* s )
{
  rt().stream = (cudaStream_t)s;
  return RSGPU_OK;
}

int rsgpu_synchronize( void )
{
  RS_TRY( ensure_device() );
  RS_CUDA( rs::stream_sync( rt().stream ) );
  return RSGPU_OK;
}

const char* rsgpu_last_error( void ) { return g_err.c_str(); }
const char* rsgpu_version( void ) { return "rsgpu 0.1 (sm_100a)"; }

int64_t rsgpu_launch_count( void ) { return (int64_t)rs::launches(); }

int rsgpu_profile_enable( int on )
{
  g_rt.profile = on != 0;
  for( int i = 0; i < N_LANES; ++i ) { g_lane_rt[i].profile = on != 0; }
  return RSGPU_OK;
}

int rsgpu_profile_reset( void )
{
  std::lock_guard<std::mutex> lk( g_prof_mu );
  drain_pending();
  g_prof.clear();
  return RSGPU_OK;
}

int rsgpu_profile_get( const char* name, double* ms, int64_t* launches )
{
  std::lock_guard<std::mutex> lk( g_prof_mu );
  drain_pending();
  auto it = g_prof.find( name );
  if( ms ) { *ms = it == g_prof.end() ? 0.0 : it->second.ms; }
  if( launches ) { *launches = it == g_prof.end() ? 0 : it->second.launches; }
  return RSGPU_OK;
}

// ---------------------------------------------------------------------------------------------- clouds
int rsgpu_cloud_create( const float* pos, const float* nor, int32_t n, rsgpu_cloud_t** out )
{
  if( !out || n < 0 || ( n > 0 && ( !pos || !nor ) ) ) { return fail( RSGPU_ERR_INVALID, "rsgpu_cloud_create: bad argument" ); }
  RS_TRY( ensure_device() );
  rsgpu_cloud* c = new rsgpu_cloud();
  c->n = n;
  cudaError_t e = c->pos.alloc( (size_t)n * 3 );
  if( e == cudaSuccess ) { e = c->nor.alloc( (size_t)n * 3 ); }
  if( e == cudaSuccess && n ) { e = cudaMemcpyAsync( c->pos.p, pos, sizeof( float ) * 3 * n, cudaMemcpyHostToDevice, rt().stream ); }
  if( e == cudaSuccess && n ) { e = cudaMemcpyAsync( c->nor.p, nor, sizeof( float ) * 3 * n, cudaMemcpyHostToDevice, rt().stream ); }
  if( e == cudaSuccess ) { e = rs::stream_sync( rt().stream ); }
  if( e != cudaSuccess ) { delete c; return cuda_fail( e, "rsgpu_cloud_create", __FILE__, __LINE__ ); }
  *out = c;
  return RSGPU_OK;
}

void rsgpu_cloud_destroy( rsgpu_cloud_t* c ) { delete c; }
int32_t rsgpu_cloud_size( const rsgpu_cloud_t* c ) { return c ? c->n : 0; }

} // extern "C"
