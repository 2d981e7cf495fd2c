// Multi-GPU plumbing: one rank per GPU, env shards are independent, the only exchange is the gradient
// (+ info sums) all-reduce before Adam.  The reference has no distributed path at all.
//
// NCCL is resolved at run time with dlopen so that libcrux_cuda.so shares the libnccl.so.2 already mapped
// by the host process (torch bundles 2.28.9) instead of linking a second copy.
//
// Besides NCCL there is a one-shot peer all-reduce for the latency-bound 44 KB gradient: every rank stores
// its vector into a slot of every peer's receive buffer over NVLink (CUDA IPC mapped pointers), bumps a
// sequence flag, then each rank sums the slots in rank order (bit-identical on all ranks).  Receive slots and flags are
// double-buffered by the parity of the sequence number: rank A can only start all-reduce N+2 after every peer pushed N+1,
// which a peer does (stream order) only after its reduce of N has finished reading the buffer of parity N%2.
#include "common.cuh"
#include <dlfcn.h>

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm_t;
typedef int (*fn_getuid)(nccl_uid *);
typedef int (*fn_initrank)(nccl_comm_t *, int, nccl_uid, int);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_destroy)(nccl_comm_t);
typedef int (*fn_split)(nccl_comm_t, int, int, nccl_comm_t *, void *);
typedef const char *(*fn_errstr)(int);

struct NcclApi {
  void *lib = nullptr;
  fn_getuid get_uid = nullptr;
  fn_initrank init_rank = nullptr;
  fn_allreduce all_reduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_split split = nullptr;
  fn_errstr err = nullptr;
  std::string why;
} g_nccl;

bool load_nccl() {
  if (g_nccl.lib) return true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) { g_nccl.why = dlerror() ? dlerror() : "dlopen(libnccl.so.2) failed"; return false; }
  g_nccl.get_uid = (fn_getuid)dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_initrank)dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.all_reduce = (fn_allreduce)dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.destroy = (fn_destroy)dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.split = (fn_split)dlsym(g_nccl.lib, "ncclCommSplit");   // NCCL >= 2.18; optional
  g_nccl.err = (fn_errstr)dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.get_uid || !g_nccl.init_rank || !g_nccl.all_reduce) { g_nccl.why = "libnccl is missing symbols"; g_nccl.lib = nullptr; return false; }
  return true;
}

// ---- one-shot peer all-reduce ----------------------------------------------------------------------
struct PeerPtrs { float *recv[16]; unsigned long long *flag[16]; };

// each block pushes a slice of `src` into slot `rank` of every peer, then (last block) publishes the sequence number
__global__ void peer_push_kernel(const float *__restrict__ src, int64_t n, PeerPtrs peers, int rank, int world, int64_t cap,
                                 unsigned long long *__restrict__ seq_dev, unsigned long long *__restrict__ done_ctr) {
  const unsigned long long seq = *(volatile unsigned long long *)seq_dev + 1ULL;  // read by every CTA before the last one advances it
  const int par = (int)(seq & 1ULL);
  for (int p = 0; p < world; ++p) {
    float *dst = peers.recv[p] + ((int64_t)par * 16 + rank) * cap;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(done_ctr, 1ULL) == (unsigned long long)gridDim.x - 1ULL;
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *done_ctr = 0ULL;
    __threadfence_system();
    for (int p = 0; p < world; ++p) {
      volatile unsigned long long *f = peers.flag[p] + par * 16 + rank;
      *f = seq;
    }
    *seq_dev = seq;
    __threadfence_system();
  }
}
__global__ void peer_reduce_kernel(float *__restrict__ dst, int64_t n, const float *__restrict__ recv, volatile unsigned long long *flags,
                                   int world, int64_t cap, const unsigned long long *__restrict__ seq_dev) {
  const unsigned long long seq = *(volatile const unsigned long long *)seq_dev;
  const int par = (int)(seq & 1ULL);
  if (threadIdx.x < world) {
    while (flags[par * 16 + threadIdx.x] < seq) { __nanosleep(100); }
  }
  __syncthreads();
  __threadfence_system();
  recv += (int64_t)par * 16 * cap;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int p = 0; p < world; ++p) s += __ldcv(recv + (int64_t)p * cap + i);
    dst[i] = s;
  }
}

}  // namespace

extern "C" {

int32_t crux_nccl_unique_id(uint8_t *id_out_host) {
  if (!id_out_host) return CRUX_ERR_INVALID;
  if (!load_nccl()) return crux_set_err(nullptr, CRUX_ERR_NCCL, "NCCL unavailable: %s", g_nccl.why.c_str());
  nccl_uid id;
  const int rc = g_nccl.get_uid(&id);
  if (rc != 0) return crux_set_err(nullptr, CRUX_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.err ? g_nccl.err(rc) : "error");
  memcpy(id_out_host, id.internal, 128);
  return CRUX_OK;
}

int32_t crux_nccl_init(crux_ctx *ctx, int32_t rank, int32_t world, const uint8_t *id_host) {
  if (!ctx) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world && id_host, "crux_nccl_init: bad rank/world/id");
  if (!load_nccl()) return crux_set_err(ctx, CRUX_ERR_NCCL, "NCCL unavailable: %s", g_nccl.why.c_str());
  CRUX_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
  nccl_uid id;
  memcpy(id.internal, id_host, 128);
  nccl_comm_t comm = nullptr;
  const int rc = g_nccl.init_rank(&comm, world, id, rank);
  if (rc != 0) return crux_set_err(ctx, CRUX_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.err ? g_nccl.err(rc) : "error");
  ctx->nccl_comm = comm; ctx->rank = rank; ctx->world = world;
  // A duplicate communicator for the side stream: the critic epochs of the fused PPO update then keep running concurrently with the
  // actor epochs on several GPUs too (each communicator sees its own collectives in the same order on every rank).
  if (world > 1 && g_nccl.split && !getenv("CRUX_NO_SIDE_COMM")) {
    nccl_comm_t side = nullptr;
    if (g_nccl.split(comm, 0, rank, &side, nullptr) == 0) ctx->nccl_comm_side = side;
  }
  return CRUX_OK;
}

int32_t crux_peer_allreduce(crux_ctx *ctx, float *buf, int64_t n);

int32_t crux_nccl_allreduce_f32(crux_ctx *ctx, float *buf, int64_t n) {
  if (!ctx) return CRUX_ERR_INVALID;
  if (ctx->world <= 1 || n <= 0) return CRUX_OK;
  if (ctx->peer_ready && n <= ctx->peer_cap) return crux_peer_allreduce(ctx, buf, n);
  CRUX_REQUIRE(ctx, ctx->nccl_comm, "crux_nccl_allreduce_f32: NCCL not initialised");
  nccl_comm_t comm = (ctx->stream == ctx->side_stream && ctx->nccl_comm_side) ? (nccl_comm_t)ctx->nccl_comm_side : (nccl_comm_t)ctx->nccl_comm;
  const int rc = g_nccl.all_reduce(buf, buf, (size_t)n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, ctx->stream);
  if (rc != 0) return crux_set_err(ctx, CRUX_ERR_NCCL, "ncclAllReduce: %s", g_nccl.err ? g_nccl.err(rc) : "error");
  return CRUX_OK;
}

// ---- peer path -----------------------------------------------------------------------------------
// handle layout (64 bytes used of 2 x cudaIpcMemHandle_t = 128): we return both handles -> 128 bytes
int32_t crux_peer_handle(crux_ctx *ctx, uint8_t *handle_out_host, int64_t max_floats) {
  if (!ctx || !handle_out_host) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, max_floats >= 1, "crux_peer_handle: max_floats < 1");
  CRUX_REQUIRE(ctx, ctx->world >= 1 && ctx->world <= 16, "crux_peer_handle: world must be set by crux_nccl_init (<= 16 ranks)");
  if (!ctx->peer_recv) {
    ctx->peer_cap = (max_floats + 31) / 32 * 32;
    // one allocation: [2 parities][16 ranks][cap] floats, then [2][16] flags + [1] block counter
    // ... then the LL region of the fused gradient exchange: [2 networks][2 parities][16 ranks][cap] 8-byte words
    const size_t base_bytes = (size_t)32 * ctx->peer_cap * sizeof(float) + 64 * sizeof(unsigned long long);
    const size_t bytes = base_bytes + (size_t)64 * ctx->peer_cap * sizeof(unsigned long long);
    CRUX_CHECK_CUDA(ctx, cudaMalloc((void **)&ctx->peer_recv, bytes));
    CRUX_CHECK_CUDA(ctx, cudaMemset(ctx->peer_recv, 0, bytes));
    ctx->peer_flags = (unsigned long long *)((char *)ctx->peer_recv + (size_t)32 * ctx->peer_cap * sizeof(float));
    ctx->peer_ll = (unsigned long long *)((char *)ctx->peer_recv + base_bytes);
  }
  cudaIpcMemHandle_t h;
  CRUX_CHECK_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->peer_recv));
  memset(handle_out_host, 0, 64);
  memcpy(handle_out_host, &h, sizeof(h) <= 64 ? sizeof(h) : 64);
  return CRUX_OK;
}

int32_t crux_peer_init(crux_ctx *ctx, int32_t rank, int32_t world, const uint8_t *handles_host) {
  if (!ctx || !handles_host) return CRUX_ERR_INVALID;
  CRUX_REQUIRE(ctx, world >= 1 && world <= 16 && rank >= 0 && rank < world, "crux_peer_init: bad rank/world");
  CRUX_REQUIRE(ctx, ctx->peer_recv, "crux_peer_init: call crux_peer_handle first");
  ctx->rank = rank; ctx->world = world;
  for (int p = 0; p < world; ++p) {
    if (p == rank) {
      ctx->peer_recv_remote[p] = ctx->peer_recv;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles_host + (size_t)p * 64, sizeof(h));
      void *ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) return crux_set_err(ctx, CRUX_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e));
      ctx->peer_recv_remote[p] = (float *)ptr;
    }
    ctx->peer_flags_remote[p] = (unsigned long long *)((char *)ctx->peer_recv_remote[p] + (size_t)32 * ctx->peer_cap * sizeof(float));
    ctx->peer_ll_remote[p] = (unsigned long long *)((char *)ctx->peer_recv_remote[p] + (size_t)32 * ctx->peer_cap * sizeof(float) + 64 * sizeof(unsigned long long));
  }
  ctx->peer_seq_dev = ctx->peer_flags + 33;
  ctx->peer_ready = true;
  return CRUX_OK;
}

// Turns the peer paths off again (the mapped buffers stay allocated): every rank must call it when ANY rank failed to map its
// peers, so that all of them fall back to NCCL together.
int32_t crux_peer_disable(crux_ctx *ctx) {
  if (!ctx) return CRUX_ERR_INVALID;
  ctx->peer_ready = false;
  return CRUX_OK;
}

int32_t crux_peer_allreduce(crux_ctx *ctx, float *buf, int64_t n) {
  CRUX_REQUIRE(ctx, ctx->peer_ready && n <= ctx->peer_cap, "crux_peer_allreduce: not initialised or vector too long");
  PeerPtrs pp;
  for (int p = 0; p < ctx->world; ++p) { pp.recv[p] = ctx->peer_recv_remote[p]; pp.flag[p] = ctx->peer_flags_remote[p]; }
  const int blocks = (int)i64max(1, i64min(cdiv(n, 1024), 16));
  peer_push_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, n, pp, ctx->rank, ctx->world, ctx->peer_cap, ctx->peer_seq_dev, ctx->peer_flags + 32);
  CRUX_LAUNCHED(ctx);
  peer_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(buf, n, ctx->peer_recv, ctx->peer_flags, ctx->world, ctx->peer_cap, ctx->peer_seq_dev);
  CRUX_LAUNCHED(ctx);
  return CRUX_OK;
}

}  // extern "C"
