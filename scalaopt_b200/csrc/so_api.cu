// so_api.cu — the C ABI of libscalaopt_cuda.so (include/scalaopt_cuda.h): contexts, HBM-resident data
// sets, the four evaluation entry points and the cross-GPU sum.
//
// Per evaluation and per GPU: ONE kernel launch (parameters travel in the kernel argument buffer, the
// per-block partials are folded by the last block of the same launch).  That block also sums the result
// across the GPUs / ranks over NVLink peer memory and writes the total into mapped pinned host memory
// (so_finish, so_device.cuh): no collective launch, no device->host copy, no stream synchronisation.
// Only results above kMailCap doubles take ncclAllReduce + a copy.  The reference does the same sum as
// `rdd.aggregate` on the Spark driver (spark-apps/.../SparkDataSetConverter.scala:42-45).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <nccl.h>   // types only: the library is dlopen'ed on first multi-GPU use (see NcclApi)

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "so_internal.h"

using namespace so;

namespace {

thread_local std::string g_init_error;

// NCCL is bound at run time, not link time: a host process that also imports torch must end up with ONE
// libnccl.so.2 (torch bundles its own), and a single-GPU context needs none at all.  Search order:
// $SCALAOPT_NCCL_LIB, an already-loaded libnccl.so.2, then the system library.
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
  bool load() {
    if (handle) return true;
    const char* env = getenv("SCALAOPT_NCCL_LIB");
    const char* names[] = { env, "libnccl.so.2", "libnccl.so" };
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (handle) break;
    }
    if (!handle) { error = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SO_BIND(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(handle, sym)); \
    if (!field) { error = std::string("libnccl lacks ") + sym; handle = nullptr; return false; }
    SO_BIND(GetUniqueId, "ncclGetUniqueId") SO_BIND(CommInitAll, "ncclCommInitAll")
    SO_BIND(CommInitRank, "ncclCommInitRank") SO_BIND(CommDestroy, "ncclCommDestroy")
    SO_BIND(AllReduce, "ncclAllReduce") SO_BIND(AllGather, "ncclAllGather") SO_BIND(GroupStart, "ncclGroupStart") SO_BIND(GroupEnd, "ncclGroupEnd")
    SO_BIND(GetErrorString, "ncclGetErrorString")
#undef SO_BIND
    return true;
  }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
bool nccl_ready(std::string* why) {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.load()) return true;
  if (why) *why = g_nccl.error;
  return false;
}

// timing events of one evaluation; two sets alternate so that the previous call's times are read while the current
// call's kernel runs
struct EvSet { cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr; bool pending = false; bool finished_on_device = false; };

struct Device {
  int id = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  double* partials = nullptr; size_t partials_cap = 0;
  double* out = nullptr;      size_t out_cap = 0;
  double* hout = nullptr;     // pinned, out_cap doubles (results too large for the mailboxes)
  // control blocks (so_internal.h): ctl0 = leave the sums on the device, ctl1 = sum across GPUs + deliver to the host
  Ctl* ctl0 = nullptr; Ctl* ctl1 = nullptr;
  unsigned long long* seq_dev = nullptr;          // device: exchanges finished
  unsigned long long seq_host = 0;                // exchanges issued by the host
  char* mailbox = nullptr;                        // own mailbox: flags [2][kMaxRanks] u64 (256 B) + slots [2][kMaxRanks][kMailCap] doubles
  void* peer_open[kMaxRanks] = {};                // rank mode: peers' mailboxes opened with cudaIpcOpenMemHandle
  double* hres = nullptr;                         // mapped pinned: kMailCap + 4 doubles, then flag and error words
  unsigned long long* hflag = nullptr;
  unsigned long long* herr = nullptr;
  EvSet ev[2];
  int ev_turn = 0;
  ncclComm_t comm = nullptr;
  // ingest: two pinned staging buffers (allocated on first use) and the events that mark them free again
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_free[2] = {nullptr, nullptr};
};
constexpr size_t kMailFlagBytes = 256;
constexpr size_t kMailBytes = kMailFlagBytes + sizeof(double) * 2 * kMaxRanks * kMailCap;

}  // namespace

struct so_ctx {
  std::vector<Device> devs;
  int rank = 0, world = 1;
  bool rank_mode = false;
  bool use_nccl = false;
  std::mutex mu;
  std::string err;
  so_call_stats stats{};
  bool finish_multi = false;      // mailboxes of all GPUs / ranks are wired: results up to kMailCap doubles are summed by so_finish
  bool poisoned = false;          // a peer timed out inside an exchange: sequence numbers can no longer be trusted
  uint64_t data_epoch = 1;        // bumped by everything that writes observations (append, load, clear, generate)
  // owning data sets of this context and their generation (a new one after so_dataset_clear): what a view checks before it
  // touches its parent's memory, so that a freed or cleared parent is SO_ESTATE instead of a stale read
  std::unordered_map<const so_dataset*, uint64_t> live;
  uint64_t next_gen = 1;
};

struct ShardMem {
  double* X = nullptr; double* y = nullptr;
  int64_t cap = 0;      // rows reserved on this device
  int64_t filled = 0;   // rows written so far
  // f == 1: cached input range of the first `range_rows` rows as of data epoch `range_epoch` (views alias their
  // parent's memory, hence a context-wide epoch rather than a per-data-set flag)
  uint64_t range_epoch = 0;
  int64_t range_rows = -1;
  bool range_ok = false;
  double tmin = 0.0, tmax = 0.0;
};

struct so_dataset {
  so_ctx* ctx = nullptr;
  int64_t m = 0;        // rows reserved in this process
  int f = 0;
  int ny = 1;           // outputs per observation (row-major y[m x ny]); > 1 only for the network families
  std::vector<ShardMem> shards;   // one per device of the context
  bool view = false;
  const so_dataset* parent = nullptr;   // views: the owning data set and its generation when the view was taken
  uint64_t parent_gen = 0;
  int64_t global_m = -1;          // rank mode: sum of filled rows over ranks (cached)
};

namespace {

int fail(so_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (ctx) ctx->err = buf; else g_init_error = buf;
  return code;
}

// a view is only as good as its parent: freed or cleared since the view was taken -> SO_ESTATE (context mutex held)
int check_view(so_dataset* ds) {
  if (!ds->view) return SO_OK;
  const auto it = ds->ctx->live.find(ds->parent);
  if (it == ds->ctx->live.end()) return fail(ds->ctx, SO_ESTATE, "view of a data set that has been freed");
  if (it->second != ds->parent_gen) return fail(ds->ctx, SO_ESTATE, "view of a data set that has been cleared since the view was taken");
  return SO_OK;
}

#define SO_CUDA(ctx, call)                                                                            \
  do { cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail((ctx), e_ == cudaErrorMemoryAllocation ? SO_ENOMEM : SO_ECUDA, \
                                       "%s failed: %s", #call, cudaGetErrorString(e_)); } while (0)
#define SO_NCCL(ctx, call)                                                                            \
  do { ncclResult_t r_ = (call);                                                                      \
    if (r_ != ncclSuccess) return fail((ctx), SO_ENCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); } while (0)

// Host memory that a GPU reads over PCIe should live on the NUMA node that GPU hangs off: with the pinned buffers of
// 4-8 ranks all on node 0 the copies of the GPUs of the other socket cross the inter-socket link and round 1's ingest
// stopped scaling at ~110 GB/s per box.  While one of these is alive the calling thread runs on (and therefore allocates
// from) the node of `device`: preferred-node memory policy where the container allows the syscall, CPU affinity + the
// kernel's local-allocation default where it does not.  Best effort: a box without NUMA information is left alone.
struct NumaNear {
  cpu_set_t old_set;
  bool moved = false, policy = false;
  int node = -1;
  static int node_of(int device) {
    char bus[32] = {};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return -1; }
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    int nd = -1;
    if (FILE* fp = fopen(path.c_str(), "r")) { if (fscanf(fp, "%d", &nd) != 1) nd = -1; fclose(fp); }
    return nd;
  }
  explicit NumaNear(int device) {
    node = node_of(device);
    if (node < 0 || node >= 64) return;
    unsigned long mask = 1ul << node;
#ifdef SYS_set_mempolicy
    policy = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, &mask, 65ul) == 0;
#endif
    char path[96];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE* fp = fopen(path, "r");
    if (!fp) return;
    char buf[4096] = {};
    const bool got = fgets(buf, sizeof buf, fp) != nullptr;
    fclose(fp);
    if (!got || sched_getaffinity(0, sizeof old_set, &old_set) != 0) return;
    cpu_set_t want;
    CPU_ZERO(&want);
    int any = 0;
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
      int a = 0, b = 0;
      const int k = sscanf(tok, "%d-%d", &a, &b);
      if (k == 1) b = a;
      if (k < 1) continue;
      for (int c = a; c <= b && c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &old_set)) { CPU_SET(c, &want); ++any; }
    }
    if (any && sched_setaffinity(0, sizeof want, &want) == 0) moved = true;
  }
  ~NumaNear() {
    if (moved) sched_setaffinity(0, sizeof old_set, &old_set);
#ifdef SYS_set_mempolicy
    if (policy) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0ul);
#endif
  }
};

int setup_device(so_ctx* ctx, Device& d) {
  SO_CUDA(ctx, cudaSetDevice(d.id));
  cudaDeviceProp prop;
  SO_CUDA(ctx, cudaGetDeviceProperties(&prop, d.id));
  if (prop.major < 10) return fail(ctx, SO_EUNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", d.id, prop.major, prop.minor);
  d.sm_count = prop.multiProcessorCount;
  SO_CUDA(ctx, cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
  for (EvSet& es : d.ev) {
    SO_CUDA(ctx, cudaEventCreate(&es.e0));
    SO_CUDA(ctx, cudaEventCreate(&es.e1));
    SO_CUDA(ctx, cudaEventCreate(&es.e2));
  }
  // control blocks + sequence word in one allocation; own mailbox; host-visible result block
  char* cb = nullptr;
  SO_CUDA(ctx, cudaMalloc(&cb, 2 * 512 + 256));
  SO_CUDA(ctx, cudaMemset(cb, 0, 2 * 512 + 256));
  static_assert(sizeof(Ctl) <= 512, "Ctl size");
  d.ctl0 = reinterpret_cast<Ctl*>(cb);
  d.ctl1 = reinterpret_cast<Ctl*>(cb + 512);
  d.seq_dev = reinterpret_cast<unsigned long long*>(cb + 1024);
  SO_CUDA(ctx, cudaMalloc(&d.mailbox, kMailBytes));
  SO_CUDA(ctx, cudaMemset(d.mailbox, 0, kMailBytes));
  void* hp = nullptr;
  NumaNear near(d.id);
  SO_CUDA(ctx, cudaHostAlloc(&hp, sizeof(double) * (kMailCap + 4) + 128, cudaHostAllocMapped | cudaHostAllocPortable));
  memset(hp, 0, sizeof(double) * (kMailCap + 4) + 128);
  d.hres = static_cast<double*>(hp);
  d.hflag = reinterpret_cast<unsigned long long*>(d.hres + kMailCap + 4);
  d.herr = d.hflag + 8;
  SO_CUDA(ctx, cudaDeviceSynchronize());
  return SO_OK;
}

unsigned long long peer_timeout_ns() {
  const char* env = getenv("SCALAOPT_PEER_TIMEOUT_MS");
  const double ms = env && *env ? atof(env) : 20000.0;
  return (unsigned long long)(std::max(1.0, ms) * 1.0e6);
}

// Write the two control blocks of device `d`: rank `rank` of `world`, mailboxes[r] = rank r's mailbox as mapped HERE.
int write_ctl(so_ctx* ctx, Device& d, int rank, int world, char* const* mailboxes) {
  SO_CUDA(ctx, cudaSetDevice(d.id));
  Ctl c{};
  c.world = world; c.rank = rank; c.cap = kMailCap;
  c.seq = d.seq_dev;
  void* dp = nullptr;
  SO_CUDA(ctx, cudaHostGetDevicePointer(&dp, d.hres, 0));
  c.host_out = static_cast<double*>(dp);
  c.host_flag = reinterpret_cast<unsigned long long*>(c.host_out + kMailCap + 4);
  c.host_err = c.host_flag + 8;
  c.timeout_ns = peer_timeout_ns();
  for (int r = 0; r < world; ++r) {
    c.flags[r] = reinterpret_cast<unsigned long long*>(mailboxes[r]);
    c.slots[r] = reinterpret_cast<double*>(mailboxes[r] + kMailFlagBytes);
  }
  c.mode = 0;
  SO_CUDA(ctx, cudaMemcpy(d.ctl0, &c, sizeof c, cudaMemcpyHostToDevice));
  c.mode = 1;
  SO_CUDA(ctx, cudaMemcpy(d.ctl1, &c, sizeof c, cudaMemcpyHostToDevice));
  return SO_OK;
}

// single-process mode: peer access between every pair of the context's GPUs
int wire_local(so_ctx* ctx) {
  const int G = (int)ctx->devs.size();
  bool all = G <= kMaxRanks;
  for (int a = 0; a < G && all; ++a)
    for (int b = 0; b < G && all; ++b) {
      if (a == b) continue;
      if (ctx->devs[a].id == ctx->devs[b].id) { all = false; break; }   // two shards on one GPU must never wait on each other
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctx->devs[a].id, ctx->devs[b].id) != cudaSuccess || !can) { all = false; break; }
      cudaSetDevice(ctx->devs[a].id);
      cudaError_t e = cudaDeviceEnablePeerAccess(ctx->devs[b].id, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) { cudaGetLastError(); all = false; }
    }
  ctx->finish_multi = all && G > 1;
  for (int g = 0; g < G; ++g) {
    char* boxes[kMaxRanks] = {};
    if (ctx->finish_multi) { for (int r = 0; r < G; ++r) boxes[r] = ctx->devs[r].mailbox; }
    else boxes[0] = ctx->devs[g].mailbox;
    // without peer access every device is a world of its own for so_finish; the cross-GPU sum then stays with NCCL
    int rc = ctx->finish_multi ? write_ctl(ctx, ctx->devs[g], g, G, boxes) : write_ctl(ctx, ctx->devs[g], 0, 1, boxes);
    if (rc) return rc;
  }
  return SO_OK;
}

// rank mode: exchange the mailboxes' IPC handles through the NCCL communicator that was just created
struct PeerCard { cudaIpcMemHandle_t handle; unsigned long long host_hash; char bus[32]; int ok; int pad; };
int wire_ranks(so_ctx* ctx) {
  Device& d = ctx->devs[0];
  const int W = ctx->world;
  char* boxes[kMaxRanks] = {};
  boxes[0] = d.mailbox;
  if (W == 1) return write_ctl(ctx, d, 0, 1, boxes);
  bool ok = W <= kMaxRanks;
  if (const char* env = getenv("SCALAOPT_NO_PEER_FINISH")) if (*env && *env != '0') ok = false;
  PeerCard mine{};
  mine.ok = ok ? 1 : 0;
  if (ok && cudaIpcGetMemHandle(&mine.handle, d.mailbox) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
  {
    char host[256] = {};
    gethostname(host, sizeof host - 1);
    std::string key(host);
    if (FILE* fp = fopen("/proc/sys/kernel/random/boot_id", "r")) { char b[64] = {}; if (fgets(b, sizeof b, fp)) key += b; fclose(fp); }
    unsigned long long h = 1469598103934665603ull;
    for (unsigned char ch : key) { h ^= ch; h *= 1099511628211ull; }
    mine.host_hash = h;
    cudaDeviceGetPCIBusId(mine.bus, sizeof mine.bus, d.id);
  }
  static_assert(sizeof(PeerCard) % 8 == 0, "PeerCard size");
  std::vector<PeerCard> cards(W);
  char* dbuf = nullptr;
  SO_CUDA(ctx, cudaSetDevice(d.id));
  SO_CUDA(ctx, cudaMalloc(&dbuf, sizeof(PeerCard) * (size_t)(W + 1)));
  SO_CUDA(ctx, cudaMemcpy(dbuf + sizeof(PeerCard) * (size_t)W, &mine, sizeof mine, cudaMemcpyHostToDevice));
  SO_NCCL(ctx, g_nccl.AllGather(dbuf + sizeof(PeerCard) * (size_t)W, dbuf, sizeof(PeerCard), ncclChar, d.comm, d.stream));
  SO_CUDA(ctx, cudaStreamSynchronize(d.stream));
  SO_CUDA(ctx, cudaMemcpy(cards.data(), dbuf, sizeof(PeerCard) * (size_t)W, cudaMemcpyDeviceToHost));
  for (int r = 0; r < W && ok; ++r) {
    if (!cards[r].ok || cards[r].host_hash != mine.host_hash) ok = false;
    for (int q = 0; q < r && ok; ++q) if (strncmp(cards[q].bus, cards[r].bus, sizeof mine.bus) == 0) ok = false;   // two ranks on one GPU
  }
  for (int r = 0; r < W && ok; ++r) {
    if (r == ctx->rank) { boxes[r] = d.mailbox; continue; }
    void* ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, cards[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
    d.peer_open[r] = ptr;
    boxes[r] = static_cast<char*>(ptr);
  }
  // everybody or nobody: a rank that could not map a peer takes all ranks back to the NCCL sum
  double* flag = reinterpret_cast<double*>(dbuf);
  const double v = ok ? 0.0 : 1.0;
  SO_CUDA(ctx, cudaMemcpy(flag, &v, sizeof v, cudaMemcpyHostToDevice));
  SO_NCCL(ctx, g_nccl.AllReduce(flag, flag, 1, ncclDouble, ncclSum, d.comm, d.stream));
  SO_CUDA(ctx, cudaStreamSynchronize(d.stream));
  double bad = 0.0;
  SO_CUDA(ctx, cudaMemcpy(&bad, flag, sizeof bad, cudaMemcpyDeviceToHost));
  cudaFree(dbuf);
  ctx->finish_multi = bad == 0.0;
  if (!ctx->finish_multi) {
    for (int r = 0; r < W; ++r) if (d.peer_open[r]) { cudaIpcCloseMemHandle(d.peer_open[r]); d.peer_open[r] = nullptr; }
    char* own[kMaxRanks] = {};
    own[0] = d.mailbox;
    return write_ctl(ctx, d, 0, 1, own);
  }
  return write_ctl(ctx, d, ctx->rank, W, boxes);
}

int ensure_workspace(so_ctx* ctx, Device& d, size_t partials, size_t out) {
  SO_CUDA(ctx, cudaSetDevice(d.id));
  // some kernels reduce more accumulators than they emit (K4: two sets of n, folded on the way out): never size
  // the result buffer below what any catalogue kernel's block reduction needs
  out = std::max<size_t>(out, 2048);
  partials = std::max<size_t>(partials, (size_t)d.sm_count * 8 * 64);
  if (partials > d.partials_cap) {
    if (d.partials) cudaFree(d.partials);
    d.partials = nullptr; d.partials_cap = 0;
    SO_CUDA(ctx, cudaMalloc(&d.partials, partials * sizeof(double)));
    d.partials_cap = partials;
  }
  if (out > d.out_cap) {
    if (d.out) cudaFree(d.out);
    if (d.hout) cudaFreeHost(d.hout);
    d.out = nullptr; d.hout = nullptr; d.out_cap = 0;
    SO_CUDA(ctx, cudaMalloc(&d.out, out * sizeof(double)));
    SO_CUDA(ctx, cudaHostAlloc(&d.hout, out * sizeof(double), cudaHostAllocDefault));
    d.out_cap = out;
  }
  return SO_OK;
}

// rank mode: nobody unmaps or frees a mailbox while a peer may still be using it
void rank_barrier(so_ctx* ctx) {
  if (!(ctx->rank_mode && ctx->use_nccl && ctx->finish_multi && !ctx->poisoned)) return;
  Device& d = ctx->devs[0];
  if (!d.comm || !d.out) return;
  cudaSetDevice(d.id);
  if (g_nccl.AllReduce(d.out, d.out, 1, ncclDouble, ncclSum, d.comm, d.stream) == ncclSuccess) cudaStreamSynchronize(d.stream);
}

void teardown(so_ctx* ctx) {
  rank_barrier(ctx);
  for (Device& d : ctx->devs) {
    cudaSetDevice(d.id);
    if (d.stream) cudaStreamSynchronize(d.stream);
    for (void*& pp : d.peer_open) if (pp) { cudaIpcCloseMemHandle(pp); pp = nullptr; }
  }
  rank_barrier(ctx);
  for (Device& d : ctx->devs) {
    cudaSetDevice(d.id);
    if (d.comm) g_nccl.CommDestroy(d.comm);
    if (d.partials) cudaFree(d.partials);
    if (d.out) cudaFree(d.out);
    if (d.hout) cudaFreeHost(d.hout);
    if (d.ctl0) cudaFree(d.ctl0);
    if (d.mailbox) cudaFree(d.mailbox);
    if (d.hres) cudaFreeHost(d.hres);
    for (EvSet& es : d.ev) {
      if (es.e0) cudaEventDestroy(es.e0);
      if (es.e1) cudaEventDestroy(es.e1);
      if (es.e2) cudaEventDestroy(es.e2);
    }
    for (int k = 0; k < 2; ++k) { if (d.stage[k]) cudaFreeHost(d.stage[k]); if (d.stage_free[k]) cudaEventDestroy(d.stage_free[k]); }
    if (d.stream) cudaStreamDestroy(d.stream);
  }
}

int check_family(so_ctx* ctx, int family, int f, int n, int ny = 1) {
  if (family < 0 || family >= SO_FAMILY_COUNT) return fail(ctx, SO_EINVAL, "unknown family %d", family);
  if (ny != 1 && !family_is_net(family))
    return fail(ctx, SO_EINVAL, "family %d takes one output per observation, data set has %d", family, ny);
  if (family_is_net(family)) {
    if (net_hidden_units(family, f, ny, n) < 0)
      return fail(ctx, SO_EINVAL, "network family %d on f=%d inputs and %d outputs: n=%d is not h (f + 1) + K (h + 1) (softmax: K (f + 1), K >= 2)", family, f, ny, n);
    return SO_OK;
  }
  if (family_is_scalar_t(family) && f != 1) return fail(ctx, SO_EINVAL, "family %d takes one input per observation, data set has f=%d", family, f);
  if (family == SO_FAMILY_POLY) {
    if (n < 1 || n > SO_MAX_POLY_N) return fail(ctx, SO_EUNSUPPORTED, "polynomial family supports 1..%d coefficients, got %d", SO_MAX_POLY_N, n);
    return SO_OK;
  }
  if (f > SO_MAX_FEATURES) return fail(ctx, SO_EUNSUPPORTED, "f=%d exceeds SO_MAX_FEATURES=%d", f, SO_MAX_FEATURES);
  const int want = family_nparams(family, f, n);
  if (n != want) return fail(ctx, SO_EINVAL, "family %d on f=%d inputs takes n=%d parameters, got %d", family, f, want, n);
  return SO_OK;
}

int global_rows(so_dataset* ds, int64_t* m) {
  so_ctx* ctx = ds->ctx;
  int64_t filled = 0;
  for (const ShardMem& sm : ds->shards) filled += sm.filled;
  if (ctx->rank_mode && ctx->world > 1) {
    if (ds->global_m < 0) {
      Device& dv = ctx->devs[0];
      int rc = ensure_workspace(ctx, dv, 1, 2);
      if (rc) return rc;
      SO_CUDA(ctx, cudaSetDevice(dv.id));
      dv.hout[1] = (double)filled;
      SO_CUDA(ctx, cudaMemcpyAsync(dv.out, dv.hout + 1, sizeof(double), cudaMemcpyHostToDevice, dv.stream));
      SO_NCCL(ctx, g_nccl.AllReduce(dv.out, dv.out, 1, ncclDouble, ncclSum, dv.comm, dv.stream));
      SO_CUDA(ctx, cudaMemcpyAsync(dv.hout, dv.out, sizeof(double), cudaMemcpyDeviceToHost, dv.stream));
      SO_CUDA(ctx, cudaStreamSynchronize(dv.stream));
      ds->global_m = (int64_t)dv.hout[0];
    }
    *m = ds->global_m;
  } else {
    *m = filled;
  }
  return SO_OK;
}

// Times of an earlier evaluation whose events have completed by now: read them and add them to the context's counters.
void resolve_events(so_ctx* ctx, Device& dv, EvSet& es, double* kernel_ms, double* device_ms) {
  if (!es.pending) return;
  es.pending = false;
  cudaSetDevice(dv.id);
  cudaEventSynchronize(es.e2);
  float a = 0.f, b = 0.f;
  cudaEventElapsedTime(&a, es.e0, es.e1);
  cudaEventElapsedTime(&b, es.e0, es.e2);
  *kernel_ms = std::max(*kernel_ms, (double)a);
  *device_ms = std::max(*device_ms, (double)b);
}

// Per-call times are known once the call's events have completed; so_stats() and the next evaluation collect them.
void collect_stats(so_ctx* ctx, int turn) {
  double k = 0.0, d = 0.0;
  bool any = false;
  for (Device& dv : ctx->devs) { any = any || dv.ev[turn].pending; resolve_events(ctx, dv, dv.ev[turn], &k, &d); }
  if (!any) return;
  ctx->stats.kernel_ms = k;
  ctx->stats.device_ms = d;
  ctx->stats.sum_kernel_ms += k;
  ctx->stats.sum_device_ms += d;
  ctx->stats.min_kernel_ms = ctx->stats.min_kernel_ms > 0.0 ? std::min(ctx->stats.min_kernel_ms, k) : k;
}

// the times of the most recent evaluation (waits for its events: the kernel has delivered, they complete within microseconds)
void settle_last(so_ctx* ctx) {
  if (ctx->devs.empty()) return;
  const int last = ctx->devs[0].ev_turn ^ 1;
  collect_stats(ctx, last ^ 1);
  collect_stats(ctx, last);
}

// Wait until device `dv` has delivered result number `seq` into its host block.
int wait_delivery(so_ctx* ctx, Device& dv, unsigned long long seq) {
  const auto t0 = std::chrono::steady_clock::now();
  unsigned long long spins = 0;
  while (__atomic_load_n(dv.hflag, __ATOMIC_ACQUIRE) < seq) {
    __builtin_ia32_pause();
    if ((++spins & 0x3fffull) == 0) {
      cudaError_t e = cudaStreamQuery(dv.stream);
      if (e == cudaSuccess) {                      // the stream drained: the result must be there
        if (__atomic_load_n(dv.hflag, __ATOMIC_ACQUIRE) >= seq) break;
        ctx->poisoned = true;
        return fail(ctx, SO_ECUDA, "evaluation kernel on device %d finished without delivering result %llu", dv.id, seq);
      }
      if (e != cudaErrorNotReady) { ctx->poisoned = true; return fail(ctx, SO_ECUDA, "evaluation on device %d failed: %s", dv.id, cudaGetErrorString(e)); }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(600)) { ctx->poisoned = true; return fail(ctx, SO_ECUDA, "evaluation on device %d did not finish within 600 s", dv.id); }
    }
  }
  if (__atomic_load_n(dv.herr, __ATOMIC_ACQUIRE) != 0) {
    ctx->poisoned = true;
    return fail(ctx, SO_ENCCL, "device %d: a peer GPU did not post its sums within the exchange timeout (result %llu)", dv.id, seq);
  }
  return SO_OK;
}

// Run one evaluation: launch per device, sum across devices/ranks, result in host memory.
// On return `res` holds the summed raw sums and res[Q] the global number of rows.
struct NllCall { int pdf; const double* consts; int nconsts; };

int run_eval(so_dataset* ds, Kind kind, int family, const double* p, const double* d, int n, const double* alpha,
             int k, int want_grad, int Q, std::vector<double>& res, const NllCall* nll = nullptr) {
  so_ctx* ctx = ds->ctx;
  if (ctx->poisoned) return fail(ctx, SO_ESTATE, "context unusable after a failed cross-GPU exchange: %s", std::string(ctx->err).c_str());
  if (int rcv = check_view(ds)) return rcv;
  const bool scalar = nll != nullptr || family_is_scalar_t(family);
  const int G = (int)ctx->devs.size();
  int64_t rows_local = 0, launches = 0;
  // data.size (DataSet.scala:126) is known on the host; in rank mode it is summed over ranks once and cached
  int64_t m_global = 0;
  {
    int rc0 = global_rows(ds, &m_global);
    if (rc0) return rc0;
    if (m_global < 1) return fail(ctx, SO_ESTATE, "data set is empty");
  }
  // the two-exponential family's K1-K4 and the likelihood kernel drop their per-row exp range check when the shard's inputs allow it: one
  // 8 B/row pass per data-set state (not per evaluation) finds their range
  const bool wants_range = ds->f == 1 && (nll ? true : family == SO_FAMILY_GAUSS_EXPBKG &&
                           (kind == kNormalEq || kind == kLossGrad || kind == kHessVec || kind == kLine));
  for (int g = 0; g < G && wants_range; ++g) {
    Device& dv = ctx->devs[g];
    ShardMem& sm = ds->shards[g];
    if (sm.filled < 1 || (sm.range_epoch == ctx->data_epoch && sm.range_rows == sm.filled)) continue;
    int rc = ensure_workspace(ctx, dv, 1, 8);
    if (rc) return rc;
    SO_CUDA(ctx, cudaSetDevice(dv.id));
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(dv.out);
    SO_CUDA(ctx, cudaMemsetAsync(buf, 0xFF, 8, dv.stream));
    SO_CUDA(ctx, cudaMemsetAsync(buf + 1, 0, 16, dv.stream));
    Shard sh; sh.X = sm.X; sh.rows = sm.filled; sh.f = 1; sh.stream = dv.stream; sh.sm_count = dv.sm_count;
    rc = launch_range(sh, buf);
    if (rc < 0) return fail(ctx, rc, "input-range kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->stats.total_launches += rc;
    SO_CUDA(ctx, cudaMemcpyAsync(dv.hout, buf, 24, cudaMemcpyDeviceToHost, dv.stream));
    SO_CUDA(ctx, cudaStreamSynchronize(dv.stream));
    sm.range_ok = range_decode(reinterpret_cast<const unsigned long long*>(dv.hout), &sm.tmin, &sm.tmax);
    sm.range_epoch = ctx->data_epoch; sm.range_rows = sm.filled;
  }
  // Who sums across GPUs?  Results that fit a mailbox slot: the kernel itself (so_finish), one launch per GPU and nothing
  // else.  Larger ones (the Gram of several hundred parameters; the TSQR triangles of several GPUs, which are gathered, not
  // summed): NCCL all-reduce + device-to-host copy.
  const int ranks = ctx->rank_mode ? ctx->world : G;
  const bool fin = Q <= kMailCap && (ranks == 1 || (ctx->finish_multi && kind != kQr));
  const int turn = ctx->devs[0].ev_turn;
  // on any failure below: leave the streams idle and the counters of the control blocks at zero
  auto bail = [&](int rc) {
    for (Device& dv : ctx->devs) {
      cudaSetDevice(dv.id);
      cudaStreamSynchronize(dv.stream);
      cudaMemsetAsync(dv.ctl0, 0, 4 * sizeof(unsigned int), dv.stream);
      cudaMemsetAsync(dv.ctl1, 0, 4 * sizeof(unsigned int), dv.stream);
      cudaStreamSynchronize(dv.stream);
      dv.ev[turn].pending = false;
    }
    cudaGetLastError();
    if (fin && ranks > 1) ctx->poisoned = true;       // peers may be waiting for sums that will never come
    return rc;
  };
  for (int g = 0; g < G; ++g) {
    Device& dv = ctx->devs[g];
    const ShardMem& sm = ds->shards[g];
    const size_t need_partials = family_is_net(family) && !nll ? mlp_partials_needed(kind, n, dv.sm_count)
                                 : scalar ? scalar_partials_needed(kind, n, k, dv.sm_count)
                                          : wide_partials_needed(kind, family, n, ds->f, k, dv.sm_count);
    int rc = ensure_workspace(ctx, dv, need_partials, (size_t)Q + 2);
    if (rc) return bail(rc);
    EvSet& es = dv.ev[turn];
    if (cudaSetDevice(dv.id) != cudaSuccess || cudaEventRecord(es.e0, dv.stream) != cudaSuccess)
      return bail(fail(ctx, SO_ECUDA, "device %d: %s", dv.id, cudaGetErrorString(cudaGetLastError())));
    Shard sh;
    sh.X = sm.X; sh.y = sm.y; sh.rows = sm.filled; sh.f = ds->f; sh.ny = ds->ny;
    sh.partials = dv.partials; sh.partials_doubles = dv.partials_cap;
    sh.out = dv.out; sh.out_doubles = dv.out_cap;
    sh.ticket = reinterpret_cast<unsigned int*>(fin ? dv.ctl1 : dv.ctl0);
    sh.stream = dv.stream; sh.sm_count = dv.sm_count;
    if (sm.filled > 0) {
      if (wants_range && sm.range_ok && sm.range_epoch == ctx->data_epoch && sm.range_rows == sm.filled) {
        sh.has_range = true; sh.tmin = sm.tmin; sh.tmax = sm.tmax;
      }
      if (kind == kQr && ranks > 1) {      // every device writes its triangle into its own slot of a zeroed buffer: the sum is a gather
        const int slot = ctx->rank_mode ? ctx->rank : g, per = (n + 1) * (n + 1);
        cudaMemsetAsync(dv.out, 0, sizeof(double) * (size_t)Q, dv.stream);
        sh.out = dv.out + (size_t)slot * per;
        sh.out_doubles = (size_t)per;
      }
      rc = nll ? launch_nll(sh, nll->pdf, p, n, nll->consts, nll->nconsts)
               : family_is_net(family) ? launch_mlp(kind, sh, family, p, n, want_grad)
               : scalar ? launch_scalar(kind, sh, family, p, d, n, alpha, k, want_grad)
                        : launch_wide(kind, sh, family, p, d, n, alpha, k, want_grad);
      if (rc < 0) {
        if (rc == SO_EUNSUPPORTED) return bail(fail(ctx, rc, "family %d with n=%d, f=%d, k=%d is outside the kernel catalogue", family, n, ds->f, k));
        cudaError_t e = cudaGetLastError();
        return bail(fail(ctx, rc, "kernel launch failed (%d): %s", rc, cudaGetErrorString(e)));
      }
      launches += rc;
    } else {
      cudaMemsetAsync(dv.out, 0, sizeof(double) * (size_t)Q, dv.stream);      // an empty shard contributes zeros
      sh.device_finish = false;
    }
    cudaEventRecord(es.e1, dv.stream);
    if (fin && !sh.device_finish) {
      sh.out = dv.out;
      rc = launch_finish(sh, Q);
      if (rc < 0) return bail(fail(ctx, rc, "k_finish launch failed: %s", cudaGetErrorString(cudaGetLastError())));
      launches += rc;
    }
    if (fin) {
      ++dv.seq_host;
      cudaEventRecord(es.e2, dv.stream);
      es.pending = true;
    }
    rows_local += sm.filled;
  }
  if (fin) {
    // while the kernels run: the times of the previous evaluation
    collect_stats(ctx, turn ^ 1);
    for (int g = 0; g < G; ++g) {
      int rc = wait_delivery(ctx, ctx->devs[g], ctx->devs[g].seq_host);
      if (rc) return rc;
    }
    const Device& d0 = ctx->devs[0];
    res.assign(d0.hres, d0.hres + Q);
    ctx->stats.finish_ms = (d0.hres[kMailCap + 1] - d0.hres[kMailCap]) * 1.0e-6;
    ctx->stats.sum_finish_ms += ctx->stats.finish_ms;
  } else {
    if (ctx->use_nccl) {
      ncclResult_t first = ncclSuccess;
      g_nccl.GroupStart();
      for (int g = 0; g < G; ++g) {
        Device& dv = ctx->devs[g];
        ncclResult_t r = g_nccl.AllReduce(dv.out, dv.out, (size_t)Q, ncclDouble, ncclSum, dv.comm, dv.stream);
        if (r != ncclSuccess && first == ncclSuccess) first = r;
      }
      ncclResult_t r = g_nccl.GroupEnd();               // always close the group, also after a failed enqueue
      if (first == ncclSuccess) first = r;
      if (first != ncclSuccess) return bail(fail(ctx, SO_ENCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString(first)));
    }
    {
      Device& d0 = ctx->devs[0];
      cudaSetDevice(d0.id);
      cudaMemcpyAsync(d0.hout, d0.out, sizeof(double) * (size_t)Q, cudaMemcpyDeviceToHost, d0.stream);
    }
    for (int g = 0; g < G; ++g) {
      Device& dv = ctx->devs[g];
      cudaSetDevice(dv.id);
      cudaEventRecord(dv.ev[turn].e2, dv.stream);
      dv.ev[turn].pending = true;
    }
    collect_stats(ctx, turn ^ 1);
    for (int g = 0; g < G; ++g) {
      Device& dv = ctx->devs[g];
      cudaSetDevice(dv.id);
      cudaError_t e = cudaStreamSynchronize(dv.stream);
      if (e != cudaSuccess) return bail(fail(ctx, SO_ECUDA, "evaluation on device %d failed: %s", dv.id, cudaGetErrorString(e)));
    }
    res.assign(ctx->devs[0].hout, ctx->devs[0].hout + Q);
    ctx->stats.finish_ms = 0.0;
  }
  for (Device& dv : ctx->devs) dv.ev_turn = turn ^ 1;
  res.push_back((double)m_global);
  ctx->stats.rows = rows_local;
  ctx->stats.bytes = rows_local * 8 * (int64_t)(nll ? 1 : ds->f + ds->ny);
  ctx->stats.launches = launches;
  ctx->stats.total_launches += launches;
  ctx->stats.total_evals += 1;
  ctx->stats.finished_on_device = fin ? 1 : 0;
  return SO_OK;
}

}  // namespace

// =================================================================================================
// Context
// =================================================================================================
extern "C" int so_abi_version(void) { return SO_ABI_VERSION; }

extern "C" const char* so_last_error(so_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

extern "C" int so_init(int ngpus, const int* device_ids, so_ctx** out) {
  if (!out || ngpus < 1) return fail(nullptr, SO_EINVAL, "so_init: ngpus must be >= 1 and out non-null");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1)
    return fail(nullptr, SO_ECUDA, "so_init: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  if (ngpus > count) return fail(nullptr, SO_EINVAL, "so_init: %d GPUs requested, %d visible", ngpus, count);
  so_ctx* ctx = new so_ctx();
  ctx->devs.resize(ngpus);
  for (int g = 0; g < ngpus; ++g) {
    ctx->devs[g].id = device_ids ? device_ids[g] : g;
    int rc = setup_device(ctx, ctx->devs[g]);
    if (rc) { g_init_error = ctx->err; teardown(ctx); delete ctx; return rc; }
  }
  if (ngpus > 1) {
    std::vector<ncclComm_t> comms(ngpus);
    std::vector<int> ids(ngpus);
    for (int g = 0; g < ngpus; ++g) ids[g] = ctx->devs[g].id;
    std::string why;
    if (!nccl_ready(&why)) { int rc = fail(nullptr, SO_ENCCL, "%s", why.c_str()); teardown(ctx); delete ctx; return rc; }
    ncclResult_t r = g_nccl.CommInitAll(comms.data(), ngpus, ids.data());
    if (r != ncclSuccess) {
      int rc = fail(nullptr, SO_ENCCL, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
      teardown(ctx); delete ctx; return rc;
    }
    for (int g = 0; g < ngpus; ++g) ctx->devs[g].comm = comms[g];
    ctx->use_nccl = true;
  }
  if (int rc = wire_local(ctx)) { g_init_error = ctx->err; teardown(ctx); delete ctx; return rc; }
  *out = ctx;
  return SO_OK;
}

extern "C" int so_nccl_unique_id(void* out_uid, size_t bytes) {
  if (!out_uid || bytes < sizeof(ncclUniqueId)) return fail(nullptr, SO_EINVAL, "so_nccl_unique_id: need %zu bytes", sizeof(ncclUniqueId));
  std::string why;
  if (!nccl_ready(&why)) return fail(nullptr, SO_ENCCL, "%s", why.c_str());
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return fail(nullptr, SO_ENCCL, "ncclGetUniqueId failed: %s", g_nccl.GetErrorString(r));
  memset(out_uid, 0, bytes);
  memcpy(out_uid, &id, sizeof id);
  return SO_OK;
}

extern "C" int so_init_rank(int device_id, int rank, int world, const void* nccl_uid, size_t uid_bytes, so_ctx** out) {
  static_assert(sizeof(ncclUniqueId) <= SO_NCCL_UID_BYTES, "uid size");
  if (!out || world < 1 || rank < 0 || rank >= world) return fail(nullptr, SO_EINVAL, "so_init_rank: bad rank/world");
  if (world > 1 && (!nccl_uid || uid_bytes < sizeof(ncclUniqueId))) return fail(nullptr, SO_EINVAL, "so_init_rank: world > 1 needs the NCCL unique id");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count < 1)
    return fail(nullptr, SO_ECUDA, "so_init_rank: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
  so_ctx* ctx = new so_ctx();
  ctx->devs.resize(1);
  ctx->devs[0].id = device_id;
  ctx->rank = rank; ctx->world = world; ctx->rank_mode = true;
  int rc = setup_device(ctx, ctx->devs[0]);
  if (rc) { g_init_error = ctx->err; teardown(ctx); delete ctx; return rc; }
  if (world > 1) {
    ncclUniqueId id;
    memcpy(&id, nccl_uid, sizeof id);
    std::string why;
    if (!nccl_ready(&why)) { rc = fail(nullptr, SO_ENCCL, "%s", why.c_str()); teardown(ctx); delete ctx; return rc; }
    ncclResult_t r = g_nccl.CommInitRank(&ctx->devs[0].comm, world, id, rank);
    if (r != ncclSuccess) {
      rc = fail(nullptr, SO_ENCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
      teardown(ctx); delete ctx; return rc;
    }
    ctx->use_nccl = true;
  }
  if ((rc = wire_ranks(ctx))) { g_init_error = ctx->err; teardown(ctx); delete ctx; return rc; }
  *out = ctx;
  return SO_OK;
}

extern "C" void so_shutdown(so_ctx* ctx) {
  if (!ctx) return;
  teardown(ctx);
  delete ctx;
}

extern "C" int so_stats(so_ctx* ctx, so_call_stats* out) {
  if (!ctx || !out) return SO_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  settle_last(ctx);
  *out = ctx->stats;
  return SO_OK;
}

extern "C" int so_stats_reset(so_ctx* ctx) {
  if (!ctx) return SO_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  settle_last(ctx);
  ctx->stats.sum_kernel_ms = ctx->stats.sum_device_ms = ctx->stats.sum_finish_ms = ctx->stats.min_kernel_ms = 0.0;
  return SO_OK;
}

// pinned host memory for callers that want full-speed H2D (e.g. a JNI direct buffer); placed on the NUMA node of the
// calling thread's current device
extern "C" int so_host_alloc(size_t bytes, void** out) {
  if (!out) return SO_EINVAL;
  int dev = 0;
  cudaGetDevice(&dev);
  NumaNear near(dev);
  cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(nullptr, SO_ENOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  return SO_OK;
}
extern "C" void so_host_free(void* p) { if (p) cudaFreeHost(p); }

// =================================================================================================
// Data set
// =================================================================================================
// contiguous row blocks of ceil(m/G) rows; pure host arithmetic, no GPU needed
extern "C" int so_shard_plan(int64_t m, int f, int G, int g, int64_t* begin, int64_t* rows) {
  if (m < 0 || f < 1 || G < 1 || g < 0 || g >= G || !begin || !rows) return SO_EINVAL;
  int64_t mg = (m + G - 1) / G;
  if (f == 1 && (mg & 1) && G > 1) ++mg;    // keep every shard's first row 16-byte aligned for 128-bit loads
  const int64_t b = std::min<int64_t>((int64_t)g * mg, m), e = std::min<int64_t>((int64_t)(g + 1) * mg, m);
  *begin = b; *rows = e - b;
  return SO_OK;
}

extern "C" int so_dataset_create(so_ctx* ctx, int64_t m, int f, so_dataset** out) { return so_dataset_create_multi(ctx, m, f, 1, out); }

extern "C" int so_dataset_create_multi(so_ctx* ctx, int64_t m, int f, int ny, so_dataset** out) {
  if (!ctx || !out) return SO_EINVAL;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (m < 0 || f < 1 || ny < 1 || ny > SO_MAX_OUTPUTS) return fail(ctx, SO_EINVAL, "so_dataset_create: m=%lld f=%d ny=%d", (long long)m, f, ny);
  so_dataset* ds = new so_dataset();
  ds->ctx = ctx; ds->m = m; ds->f = f; ds->ny = ny;
  const int G = (int)ctx->devs.size();
  ds->shards.resize(G);
  for (int g = 0; g < G; ++g) {
    int64_t b = 0, cnt = 0;
    so_shard_plan(m, f, G, g, &b, &cnt);
    ShardMem& sm = ds->shards[g];
    sm.cap = cnt;
    if (sm.cap > 0) {
      cudaError_t ce = cudaSetDevice(ctx->devs[g].id);
      // + 16 bytes: the odd-f kernels stream rows as aligned 16-byte units and may touch one element past the last row
      if (ce == cudaSuccess) ce = cudaMalloc(&sm.X, sizeof(double) * ((size_t)sm.cap * f + 2));
      if (ce == cudaSuccess) ce = cudaMalloc(&sm.y, sizeof(double) * (size_t)sm.cap * ny);
      if (ce != cudaSuccess) {
        int rc = fail(ctx, ce == cudaErrorMemoryAllocation ? SO_ENOMEM : SO_ECUDA,
                      "so_dataset_create: %lld rows x %d on device %d: %s", (long long)sm.cap, f, ctx->devs[g].id, cudaGetErrorString(ce));
        cudaGetLastError();
        for (ShardMem& s2 : ds->shards) { if (s2.X) cudaFree(s2.X); if (s2.y) cudaFree(s2.y); }
        delete ds;
        return rc;
      }
    }
  }
  ctx->live[ds] = ctx->next_gen++;
  *out = ds;
  return SO_OK;
}

extern "C" void so_dataset_free(so_dataset* ds) {
  if (!ds) return;
  if (!ds->view) {
    { std::lock_guard<std::mutex> lk(ds->ctx->mu); ds->ctx->live.erase(ds); }
    for (size_t g = 0; g < ds->shards.size(); ++g) {
      cudaSetDevice(ds->ctx->devs[g].id);
      if (ds->shards[g].X) cudaFree(ds->shards[g].X);
      if (ds->shards[g].y) cudaFree(ds->shards[g].y);
    }
  }
  delete ds;
}

namespace {

constexpr size_t kStageBytes = (size_t)64 << 20;      // per pinned staging buffer

int ensure_staging(so_ctx* ctx, Device& dv) {
  SO_CUDA(ctx, cudaSetDevice(dv.id));
  NumaNear near(dv.id);
  for (int k = 0; k < 2; ++k) {
    if (!dv.stage[k]) SO_CUDA(ctx, cudaHostAlloc(&dv.stage[k], kStageBytes, cudaHostAllocDefault));
    if (!dv.stage_free[k]) SO_CUDA(ctx, cudaEventCreateWithFlags(&dv.stage_free[k], cudaEventDisableTiming));
  }
  return SO_OK;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// A source of rows: fills `dst` with `count` doubles starting `elem_offset` doubles into the X or y stream.
struct RowSource {
  virtual ~RowSource() = default;
  virtual int fetch(bool is_y, int64_t elem_offset, int64_t count, double* dst) = 0;
  virtual const double* direct(bool is_y, int64_t elem_offset) { (void)is_y; (void)elem_offset; return nullptr; }   // pinned: DMA straight from the source
};
// host-side part of a staged copy, split over a few threads: one core's memcpy / pread (~10-15 GB/s) is
// well below what one PCIe 5 x16 link takes (~55 GB/s)
constexpr int kIngestThreads = 6;
template <class F>
int parallel_chunks(int64_t count, F&& body) {      // body(begin, end) -> SO_* code
  const int64_t grain = (int64_t)1 << 20;             // doubles
  const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(kIngestThreads, count / grain));
  if (nt == 1) return body(0, count);
  std::vector<std::thread> th;
  std::vector<int> rc(nt, SO_OK);
  for (int t = 0; t < nt; ++t) {
    const int64_t b = count * t / nt, e = count * (t + 1) / nt;
    th.emplace_back([&, t, b, e] { rc[t] = body(b, e); });
  }
  for (auto& x : th) x.join();
  for (int r : rc) if (r != SO_OK) return r;
  return SO_OK;
}

struct MemSource : RowSource {
  const double* X; const double* y; bool pinned;
  MemSource(const double* X_, const double* y_) : X(X_), y(y_), pinned(is_pinned(X_) && is_pinned(y_)) {}
  int fetch(bool is_y, int64_t off, int64_t count, double* dst) override {
    const double* src = (is_y ? y : X) + off;
    return parallel_chunks(count, [&](int64_t b, int64_t e) { memcpy(dst + b, src + b, sizeof(double) * (size_t)(e - b)); return (int)SO_OK; });
  }
  const double* direct(bool is_y, int64_t off) override { return pinned ? (is_y ? y : X) + off : nullptr; }
};
struct FileSource : RowSource {
  int fdx = -1, fdy = -1; int64_t base_rows = 0; int f = 1, ny = 1;
  ~FileSource() override { if (fdx >= 0) close(fdx); if (fdy >= 0) close(fdy); }
  int fetch(bool is_y, int64_t off, int64_t count, double* dst) override {
    const int fd = is_y ? fdy : fdx;
    const int64_t base = (int64_t)sizeof(double) * ((is_y ? base_rows * ny : base_rows * f) + off);
    return parallel_chunks(count, [&](int64_t b, int64_t e) {
      char* out = reinterpret_cast<char*>(dst + b);
      int64_t pos = base + (int64_t)sizeof(double) * b, left = (int64_t)sizeof(double) * (e - b);
      while (left > 0) {
        const ssize_t got = pread(fd, out, (size_t)left, (off_t)pos);
        if (got <= 0) return (int)SO_EINVAL;           // error or end of file: the file is shorter than promised
        out += got; pos += got; left -= got;
      }
      return (int)SO_OK;
    });
  }
};

// Copy `rows` rows from `src` into the data set, filling the shards in order.  Staged copies alternate between the
// two pinned buffers of the destination device: while buffer k is in flight to the GPU, buffer k^1 is being filled.
int ingest(so_dataset* ds, RowSource& src, int64_t rows) {
  so_ctx* ctx = ds->ctx;
  int64_t filled = 0;
  for (const ShardMem& sm : ds->shards) filled += sm.filled;
  if (filled + rows > ds->m) return fail(ctx, SO_EINVAL, "ingest: %lld rows do not fit (%lld of %lld used)", (long long)rows, (long long)filled, (long long)ds->m);
  const int f = ds->f, ny = ds->ny;
  int64_t done = 0;
  int turn = 0;
  for (size_t g = 0; g < ds->shards.size() && done < rows; ++g) {
    ShardMem& sm = ds->shards[g];
    const int64_t take = std::min<int64_t>(sm.cap - sm.filled, rows - done);
    if (take <= 0) continue;
    Device& dv = ctx->devs[g];
    SO_CUDA(ctx, cudaSetDevice(dv.id));
    for (int pass = 0; pass < 2; ++pass) {               // X stream, then y stream
      const bool is_y = pass == 1;
      const int64_t width = is_y ? ny : f;
      double* dst = is_y ? sm.y + (size_t)sm.filled * ny : sm.X + (size_t)sm.filled * f;
      const int64_t total = take * width, src_off = done * width;
      if (const double* direct = src.direct(is_y, src_off)) {
        SO_CUDA(ctx, cudaMemcpyAsync(dst, direct, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice, dv.stream));
        continue;
      }
      int rc = ensure_staging(ctx, dv);
      if (rc) return rc;
      const int64_t chunk = (int64_t)(kStageBytes / sizeof(double));
      for (int64_t o = 0; o < total; o += chunk, turn ^= 1) {
        const int64_t cnt = std::min(chunk, total - o);
        SO_CUDA(ctx, cudaEventSynchronize(dv.stage_free[turn]));           // buffer `turn` has left for the GPU
        rc = src.fetch(is_y, src_off + o, cnt, (double*)dv.stage[turn]);
        if (rc) return fail(ctx, rc, "ingest: short read from the source");
        SO_CUDA(ctx, cudaMemcpyAsync(dst + o, dv.stage[turn], sizeof(double) * (size_t)cnt, cudaMemcpyHostToDevice, dv.stream));
        SO_CUDA(ctx, cudaEventRecord(dv.stage_free[turn], dv.stream));
      }
    }
    sm.filled += take;
    done += take;
  }
  for (Device& dv : ctx->devs) { SO_CUDA(ctx, cudaSetDevice(dv.id)); SO_CUDA(ctx, cudaStreamSynchronize(dv.stream)); }
  ds->global_m = -1;
  return SO_OK;
}

}  // namespace

extern "C" int so_dataset_append(so_dataset* ds, const double* X, const double* y, int64_t rows) {
  if (!ds) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ds->view) return fail(ctx, SO_EINVAL, "so_dataset_append: cannot append to a view");
  ++ctx->data_epoch;
  if (rows < 0 || (rows > 0 && (!X || !y))) return fail(ctx, SO_EINVAL, "so_dataset_append: bad arguments");
  MemSource src(X, y);
  return ingest(ds, src, rows);
}

extern "C" int so_dataset_load_raw(so_dataset* ds, const char* path_X, const char* path_y, int64_t rows, int64_t file_row_offset) {
  if (!ds || !path_X || !path_y) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ds->view) return fail(ctx, SO_EINVAL, "so_dataset_load_raw: cannot load into a view");
  ++ctx->data_epoch;
  if (rows < 0 || file_row_offset < 0) return fail(ctx, SO_EINVAL, "so_dataset_load_raw: bad arguments");
  FileSource src;
  src.fdx = open(path_X, O_RDONLY);
  src.fdy = open(path_y, O_RDONLY);
  src.base_rows = file_row_offset; src.f = ds->f; src.ny = ds->ny;
  if (src.fdx < 0 || src.fdy < 0) return fail(ctx, SO_EINVAL, "so_dataset_load_raw: cannot open %s / %s", path_X, path_y);
  return ingest(ds, src, rows);
}

// empty the data set without releasing HBM (lets a caller re-ingest into the same reservation)
extern "C" int so_dataset_clear(so_dataset* ds) {
  if (!ds) return SO_EINVAL;
  std::lock_guard<std::mutex> lk(ds->ctx->mu);
  if (ds->view) return fail(ds->ctx, SO_EINVAL, "so_dataset_clear: cannot clear a view");
  ++ds->ctx->data_epoch;
  ds->ctx->live[ds] = ds->ctx->next_gen++;        // views taken before this point are stale
  for (ShardMem& sm : ds->shards) sm.filled = 0;
  ds->global_m = -1;
  return SO_OK;
}

extern "C" int so_dataset_generate(so_dataset* ds, int family, const double* p_true, int n, uint64_t seed,
                                   double noise_sigma, int64_t row_offset) {
  if (!ds || !p_true) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ds->view) return fail(ctx, SO_EINVAL, "so_dataset_generate: cannot fill a view");
  ++ctx->data_epoch;
  if (ds->ny != 1) return fail(ctx, SO_EUNSUPPORTED, "so_dataset_generate: synthetic data exist for one output per observation only");
  int rc = check_family(ctx, family, ds->f, n);
  if (rc) return rc;
  // the t grid of the scalar families spans the global data set: in rank mode every rank holds m rows
  const int64_t total = ctx->rank_mode ? ds->m * ctx->world : ds->m;
  int64_t off = row_offset;
  for (size_t g = 0; g < ds->shards.size(); ++g) {
    ShardMem& sm = ds->shards[g];
    Device& dv = ctx->devs[g];
    SO_CUDA(ctx, cudaSetDevice(dv.id));
    Shard sh; sh.rows = sm.cap; sh.f = ds->f; sh.stream = dv.stream; sh.sm_count = dv.sm_count;
    rc = launch_generate(sh, sm.X, sm.y, family, p_true, n, seed, noise_sigma, off, total);
    if (rc < 0) return fail(ctx, rc, "so_dataset_generate: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ctx->stats.total_launches += rc;
    sm.filled = sm.cap;
    off += sm.cap;
  }
  for (Device& dv : ctx->devs) { SO_CUDA(ctx, cudaSetDevice(dv.id)); SO_CUDA(ctx, cudaStreamSynchronize(dv.stream)); }
  ds->global_m = -1;
  return SO_OK;
}

extern "C" int so_dataset_read(so_dataset* ds, int64_t row_begin, int64_t rows, double* X, double* y) {
  if (!ds) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (int rcv = check_view(ds)) return rcv;
  int64_t filled = 0;
  for (const ShardMem& sm : ds->shards) filled += sm.filled;
  if (row_begin < 0 || rows < 0 || row_begin + rows > filled) return fail(ctx, SO_EINVAL, "so_dataset_read: rows [%lld, %lld) outside [0, %lld)", (long long)row_begin, (long long)(row_begin + rows), (long long)filled);
  int64_t base = 0, done = 0;
  for (size_t g = 0; g < ds->shards.size() && done < rows; ++g) {
    const ShardMem& sm = ds->shards[g];
    const int64_t lo = std::max<int64_t>(row_begin + done, base), hi = std::min<int64_t>(row_begin + rows, base + sm.filled);
    if (hi > lo) {
      SO_CUDA(ctx, cudaSetDevice(ctx->devs[g].id));
      if (X) SO_CUDA(ctx, cudaMemcpy(X + (size_t)done * ds->f, sm.X + (size_t)(lo - base) * ds->f, sizeof(double) * (size_t)(hi - lo) * ds->f, cudaMemcpyDeviceToHost));
      if (y) SO_CUDA(ctx, cudaMemcpy(y + (size_t)done * ds->ny, sm.y + (size_t)(lo - base) * ds->ny, sizeof(double) * (size_t)(hi - lo) * ds->ny, cudaMemcpyDeviceToHost));
      done += hi - lo;
    }
    base += sm.filled;
  }
  return SO_OK;
}

extern "C" int so_dataset_size(so_dataset* ds, int64_t* m) {
  if (!ds || !m) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  return global_rows(ds, m);
}

extern "C" int so_dataset_head(so_dataset* ds, double* x, double* y) {
  if (!ds) return SO_EINVAL;
  int64_t filled = 0;
  for (const ShardMem& sm : ds->shards) filled += sm.filled;
  if (filled < 1) return fail(ds->ctx, SO_ESTATE, "so_dataset_head: empty data set");   // Seq.head on Nil throws
  return so_dataset_read(ds, 0, 1, x, y);
}

extern "C" int so_dataset_slice(so_dataset* ds, int64_t begin, int64_t end, so_dataset** view) {
  if (!ds || !view) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (int rcv = check_view(ds)) return rcv;
  int64_t filled = 0;
  for (const ShardMem& sm : ds->shards) filled += sm.filled;
  if (begin < 0 || end < begin || end > filled) return fail(ctx, SO_EINVAL, "so_dataset_slice: [%lld, %lld) outside [0, %lld)", (long long)begin, (long long)end, (long long)filled);
  so_dataset* v = new so_dataset();
  v->ctx = ctx; v->f = ds->f; v->ny = ds->ny; v->view = true; v->m = end - begin;
  v->parent = ds->view ? ds->parent : ds;
  v->parent_gen = ds->view ? ds->parent_gen : ctx->live[ds];
  v->shards.resize(ds->shards.size());
  int64_t base = 0;
  for (size_t g = 0; g < ds->shards.size(); ++g) {
    const ShardMem& sm = ds->shards[g];
    const int64_t lo = std::max<int64_t>(begin, base), hi = std::min<int64_t>(end, base + sm.filled);
    if (hi > lo) {
      v->shards[g].X = sm.X + (size_t)(lo - base) * ds->f;
      v->shards[g].y = sm.y + (size_t)(lo - base) * ds->ny;
      v->shards[g].cap = v->shards[g].filled = hi - lo;
    }
    base += sm.filled;
  }
  *view = v;
  return SO_OK;
}

// =================================================================================================
// Evaluation
// =================================================================================================
// loss and gradient at x with the context mutex already held (used by the composed MLP evaluations below)
static int loss_grad_unlocked(so_dataset* ds, int family, const double* x, int n, double* loss, double* grad, so_call_stats* acc) {
  so_ctx* ctx = ds->ctx;
  const int Q = grad ? 1 + n : 1;
  std::vector<double> res;
  int rc = run_eval(ds, kLossGrad, family, x, nullptr, n, nullptr, 0, grad ? 1 : 0, Q, res);
  if (rc) return rc;
  const double m = res[Q];
  *loss = res[0] / m;
  if (grad) for (int i = 0; i < n; ++i) grad[i] = res[1 + i] / m;
  if (acc) {
    settle_last(ctx);
    acc->kernel_ms += ctx->stats.kernel_ms; acc->device_ms += ctx->stats.device_ms; acc->launches += ctx->stats.launches;
    acc->rows += ctx->stats.rows; acc->bytes += ctx->stats.bytes;
  }
  return SO_OK;
}

extern "C" int so_eval_loss_grad(so_dataset* ds, int family, const double* p, int n, double* loss, double* grad) {
  if (!ds || !p || !loss) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = check_family(ctx, family, ds->f, n, ds->ny);
  if (rc) return rc;
  const int Q = grad ? 1 + n : 1;
  std::vector<double> res;
  rc = run_eval(ds, kLossGrad, family, p, nullptr, n, nullptr, 0, grad ? 1 : 0, Q, res);
  if (rc) return rc;
  const double m = res[Q];
  *loss = res[0] / m;                                   // ... / data.size, MSEFunction.scala:35
  if (grad) for (int i = 0; i < n; ++i) grad[i] = res[1 + i] / m;
  return SO_OK;
}

extern "C" int so_eval_normal_eq(so_dataset* ds, int family, const double* p, int n, double* JtJ, double* Jtr, double* rtr) {
  if (!ds || !p || !JtJ || !Jtr || !rtr) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = check_family(ctx, family, ds->f, n, ds->ny);
  if (rc) return rc;
  const int T = n * (n + 1) / 2, Q = T + n + 1;
  std::vector<double> res;
  rc = run_eval(ds, kNormalEq, family, p, nullptr, n, nullptr, 0, 0, Q, res);
  if (rc) return rc;
  int q = 0;
  for (int a = 0; a < n; ++a)
    for (int b = a; b < n; ++b) { JtJ[a * n + b] = res[q]; JtJ[b * n + a] = res[q]; ++q; }
  for (int a = 0; a < n; ++a) Jtr[a] = res[T + a];
  *rtr = res[T + n];
  return SO_OK;
}

// R factor of [R_0; R_1; ...]: Householder on the stacked triangles, rank order (host, O(G n^3))
static void merge_triangles(const double* tri, int count, int C, double* R) {
  std::vector<double> A((size_t)count * C * C);
  std::copy(tri, tri + (size_t)count * C * C, A.begin());
  const int M = count * C;
  for (int j = 0; j < C; ++j) {
    double sigma = 0.0;
    for (int i = j + 1; i < M; ++i) sigma += A[(size_t)i * C + j] * A[(size_t)i * C + j];
    const double ajj = A[(size_t)j * C + j];
    const double nrm = std::sqrt(ajj * ajj + sigma);
    if (nrm == 0.0) continue;
    const double alpha = ajj > 0.0 ? -nrm : nrm, vp = ajj - alpha, beta = -1.0 / (alpha * vp);
    A[(size_t)j * C + j] = alpha;
    for (int k = j + 1; k < C; ++k) {
      double sdot = vp * A[(size_t)j * C + k];
      for (int i = j + 1; i < M; ++i) sdot += A[(size_t)i * C + j] * A[(size_t)i * C + k];
      sdot *= beta;
      A[(size_t)j * C + k] -= sdot * vp;
      for (int i = j + 1; i < M; ++i) A[(size_t)i * C + k] -= sdot * A[(size_t)i * C + j];
    }
    for (int i = j + 1; i < M; ++i) A[(size_t)i * C + j] = 0.0;
  }
  for (int j = 0; j < C; ++j)
    for (int k = 0; k < C; ++k) R[j * C + k] = k >= j ? A[(size_t)j * C + k] : 0.0;
}

extern "C" int so_eval_qr(so_dataset* ds, int family, const double* p, int n, double* R) {
  if (!ds || !p || !R) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = check_family(ctx, family, ds->f, n, ds->ny);
  if (rc) return rc;
  const bool narrow = (family == SO_FAMILY_LINEAR || family == SO_FAMILY_LINEAR_NOINT || family == SO_FAMILY_LOGISTIC) && n <= SO_QR_COLS_MAX_N;
  if (!family_is_scalar_t(family) && !narrow)
    return fail(ctx, SO_EUNSUPPORTED, "so_eval_qr: family %d with n=%d is outside the TSQR catalogue (single-input families; linear / logistic with n <= %d)", family, n, SO_QR_COLS_MAX_N);
  const int C = n + 1, per = C * C;
  const int slots = ctx->rank_mode ? ctx->world : (int)ctx->devs.size();
  std::vector<double> res;
  rc = run_eval(ds, kQr, family, p, nullptr, n, nullptr, 0, 0, slots * per, res);
  if (rc) return rc;
  if (slots == 1) std::copy(res.begin(), res.begin() + per, R);
  else merge_triangles(res.data(), slots, C, R);
  return SO_OK;
}

extern "C" int so_eval_line(so_dataset* ds, int family, const double* p, const double* d, int n,
                            const double* alpha, int k, double* phi, double* dphi) {
  if (!ds || !p || !d || !alpha || !phi || !dphi) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = check_family(ctx, family, ds->f, n, ds->ny);
  if (rc) return rc;
  if (k < 1 || k > SO_MAX_ALPHAS) return fail(ctx, SO_EINVAL, "so_eval_line: k=%d outside [1, %d]", k, SO_MAX_ALPHAS);
  if (family_is_net(family)) {
    // LineSearchPoint.fx / dfx as the reference evaluates them: f(x + a d) and gradient(x + a d) . d (Point.scala:54-57),
    // one fused loss+gradient launch per step size
    so_call_stats acc{};
    std::vector<double> x(n), g(n);
    for (int j = 0; j < k; ++j) {
      for (int i = 0; i < n; ++i) x[i] = p[i] + d[i] * alpha[j];
      rc = loss_grad_unlocked(ds, family, x.data(), n, &phi[j], g.data(), &acc);
      if (rc) return rc;
      double s2 = 0.0;
      for (int i = 0; i < n; ++i) s2 += g[i] * d[i];
      dphi[j] = s2;
    }
    ctx->stats.kernel_ms = acc.kernel_ms; ctx->stats.device_ms = acc.device_ms; ctx->stats.launches = acc.launches;
    ctx->stats.rows = acc.rows; ctx->stats.bytes = acc.bytes;
    return SO_OK;
  }
  const bool scalar = family_is_scalar_t(family);
  // one pass scores up to `chunk` step sizes; longer schedules take ceil(k/chunk) passes
  const int chunk = 8;
  so_call_stats acc{};
  for (int k0 = 0; k0 < k; k0 += chunk) {
    const int kk = std::min(chunk, k - k0);
    const int KB = scalar ? scalar_line_block(kk) : kk;
    std::vector<double> res;
    rc = run_eval(ds, kLine, family, p, d, n, alpha + k0, kk, 0, 2 * KB, res);
    if (rc) return rc;
    const double m = res[2 * KB];
    for (int j = 0; j < kk; ++j) { phi[k0 + j] = res[j] / m; dphi[k0 + j] = res[KB + j] / m; }
    if (k > chunk) settle_last(ctx);
    acc.kernel_ms += ctx->stats.kernel_ms; acc.device_ms += ctx->stats.device_ms; acc.launches += ctx->stats.launches;
    acc.rows += ctx->stats.rows; acc.bytes += ctx->stats.bytes;
  }
  if (k > chunk) {        // several passes: their times add up (a single pass keeps the lazily collected per-call times)
    ctx->stats.kernel_ms = acc.kernel_ms; ctx->stats.device_ms = acc.device_ms; ctx->stats.launches = acc.launches;
    ctx->stats.rows = acc.rows; ctx->stats.bytes = acc.bytes;
  }
  return SO_OK;
}

extern "C" int so_eval_nll(so_dataset* ds, int pdf, const double* p, int n, const double* consts, int nconsts, double* nll) {
  if (!ds || !p || !nll) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (ds->f != 1) return fail(ctx, SO_EINVAL, "so_eval_nll: the data set must have one input per observation, has f=%d", ds->f);
  if (pdf < 0 || pdf >= SO_PDF_COUNT) return fail(ctx, SO_EUNSUPPORTED, "so_eval_nll: unknown pdf %d", pdf);
  NllCall call{pdf, consts, nconsts};
  std::vector<double> res;
  int rc = run_eval(ds, kNll, 0, p, nullptr, n, nullptr, 0, 0, 1, res, &call);
  if (rc) return rc;
  *nll = res[0];
  return SO_OK;
}

extern "C" int so_eval_hess_vec(so_dataset* ds, int family, const double* p, const double* d, int n, double* Hd) {
  if (!ds || !p || !d || !Hd) return SO_EINVAL;
  so_ctx* ctx = ds->ctx;
  std::lock_guard<std::mutex> lk(ctx->mu);
  int rc = check_family(ctx, family, ds->f, n, ds->ny);
  if (rc) return rc;
  if (family_is_net(family)) {
    // FFNeuralNetworkTrainer.dirHessian: (gradient(w + d * eps) - gradient(w)) / eps, eps = 1e-8 (Trainer.scala:46,76-80)
    const double eps = 1.0e-8;
    so_call_stats acc{};
    std::vector<double> x(n), g0(n), g1(n);
    double l0 = 0.0, l1 = 0.0;
    rc = loss_grad_unlocked(ds, family, p, n, &l0, g0.data(), &acc);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) x[i] = p[i] + d[i] * eps;
    rc = loss_grad_unlocked(ds, family, x.data(), n, &l1, g1.data(), &acc);
    if (rc) return rc;
    for (int i = 0; i < n; ++i) Hd[i] = (g1[i] - g0[i]) / eps;
    ctx->stats.kernel_ms = acc.kernel_ms; ctx->stats.device_ms = acc.device_ms; ctx->stats.launches = acc.launches;
    ctx->stats.rows = acc.rows; ctx->stats.bytes = acc.bytes;
    return SO_OK;
  }
  std::vector<double> res;
  rc = run_eval(ds, kHessVec, family, p, d, n, nullptr, 0, 0, n, res);
  if (rc) return rc;
  const double m = res[n];
  for (int i = 0; i < n; ++i) Hd[i] = res[i] / m;
  return SO_OK;
}
