// Trajectory-sharded ensembles behind the C ABI (SURVEY.md §8b "ensemble / multi-GPU", §8e).
//
// An ensemble of B_total independent trajectories is cut into contiguous blocks, one per rank; every
// rank holds the full operators and a [N][B_local] batched state and steps it with the batched
// kernels -- NO communication inside the time loop.  The only collectives are the final gathers of
// per-trajectory numbers (expectation values) and of the states, and they live here, inside the
// library, so that a Julia host (no torch.distributed) can run an ensemble on all GPUs of a box:
//
//   single process, many GPUs   qp_ens_create(devices, n)      one context per entry, ncclCommInitAll
//   one process per GPU         qp_ens_unique_id + qp_ens_create_rank   (id distributed by the launcher)
//
// Transport: NCCL over NVLink (grouped ncclBroadcast, one per rank -- blocks may be ragged), loaded
// with dlopen at first use (libnccl.so.2: the copy a host such as PyTorch already loaded, else the
// system one), or plain device copies when all ranks of a single-process ensemble sit on the SAME
// device ("fake ranks": the layout logic of the gathers runs on a 1-GPU box).
//
// The hook this serves in the reference: one propagator per trajectory over a shared spectral
// envelope (`control_ranges`, src/cheby_propagator.jl:59-66), results collected by the caller.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "qprop_internal.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  const char* names[] = {getenv("QPROP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm || !*nm) continue;
    api.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    api.err = dlerror() ? dlerror() : "libnccl.so.2 not found";
    return nullptr;
  }
#define QP_SYM(field, name)                                         \
  *(void**)(&api.field) = dlsym(api.handle, name);                  \
  if (!api.field) {                                                 \
    api.err = std::string("missing NCCL symbol ") + name;           \
    dlclose(api.handle);                                            \
    api.handle = nullptr;                                           \
    return nullptr;                                                 \
  }
  QP_SYM(GetUniqueId, "ncclGetUniqueId")
  QP_SYM(CommInitRank, "ncclCommInitRank")
  QP_SYM(CommInitAll, "ncclCommInitAll")
  QP_SYM(CommDestroy, "ncclCommDestroy")
  QP_SYM(Broadcast, "ncclBroadcast")
  QP_SYM(GroupStart, "ncclGroupStart")
  QP_SYM(GroupEnd, "ncclGroupEnd")
  QP_SYM(GetErrorString, "ncclGetErrorString")
#undef QP_SYM
  return &api;
}

}  // namespace

struct qp_ens_s {
  int n_ranks = 0;                 // world size
  std::vector<int> local_rank;     // world rank of local member i
  std::vector<CtxHandle> ctx;      // context of local member i (holds a reference)
  std::vector<bool> own_ctx;       // created by qp_ens_create (destroyed with the ensemble)
  std::vector<ncclComm_t> comm;    // empty: local-copy transport
  bool use_nccl = false;
  std::vector<double2*> d_stage;   // per local member: staging of the concatenated rank blocks
  std::vector<size_t> stage_elems;
};

#define QP_NCCL(ctx, api, call)                                                                           \
  do {                                                                                                    \
    ncclResult_t r__ = (call);                                                                            \
    if (r__ != ncclSuccess)                                                                               \
      return qp_fail((ctx), QP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

// NCCL sets up its channels and peer connections lazily at the first collective (~1 s at 8 ranks):
// do that here, at creation, so that the first gather costs what every gather costs.
static int32_t ens_warm_up(qp_ens_t E);

extern "C" int32_t qp_ens_shard(int64_t n_total, int32_t rank, int32_t n_ranks, int64_t* b0, int64_t* b1) {
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks || n_total < 0 || !b0 || !b1)
    return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_shard: rank %d outside a world of %d", rank, n_ranks);
  const int64_t base = n_total / n_ranks, extra = n_total % n_ranks;
  *b0 = rank * base + std::min<int64_t>(rank, extra);
  *b1 = *b0 + base + (rank < extra ? 1 : 0);
  return QP_OK;
}

extern "C" int32_t qp_ens_create(const int32_t* devices, int32_t n_ranks, qp_ens_t* out) {
  if (!devices || !out || n_ranks < 1) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_create: bad arguments");
  *out = nullptr;
  qp_ens_t E = new qp_ens_s();
  E->n_ranks = n_ranks;
  bool distinct = true;
  for (int i = 0; i < n_ranks; ++i)
    for (int j = 0; j < i; ++j) distinct &= devices[i] != devices[j];
  for (int i = 0; i < n_ranks; ++i) {
    qp_ctx_t c = nullptr;
    int32_t rc = qp_ctx_create(devices[i], &c);
    if (rc != QP_OK) {
      for (qp_ctx_t cc : E->ctx) qp_ctx_destroy(cc);  /* deferred until E releases them */
      delete E;
      return rc;
    }
    E->ctx.emplace_back();
    E->ctx.back() = c;
    E->own_ctx.push_back(true);
    E->local_rank.push_back(i);
  }
  E->d_stage.assign(n_ranks, nullptr);
  E->stage_elems.assign(n_ranks, 0);
  if (n_ranks > 1 && distinct) {
    NcclApi* api = nccl_api();
    if (!api) {
      int32_t rc = qp_fail(E->ctx[0], QP_ERR_UNSUPPORTED, "qp_ens_create: NCCL is not available (libnccl.so.2 could not be loaded)");
      for (qp_ctx_t cc : E->ctx) qp_ctx_destroy(cc);  /* deferred until E releases them */
      delete E;
      return rc;
    }
    E->comm.resize(n_ranks);
    std::vector<int> devs(devices, devices + n_ranks);
    ncclResult_t r = api->CommInitAll(E->comm.data(), n_ranks, devs.data());
    if (r != ncclSuccess) {
      int32_t rc = qp_fail(E->ctx[0], QP_ERR_CUDA, "ncclCommInitAll failed: %s", api->GetErrorString(r));
      for (qp_ctx_t cc : E->ctx) qp_ctx_destroy(cc);  /* deferred until E releases them */
      delete E;
      return rc;
    }
    E->use_nccl = true;
    int32_t wrc = ens_warm_up(E);
    if (wrc != QP_OK) {
      qp_ens_destroy(E);
      return wrc;
    }
  } else if (n_ranks > 1 && !distinct) {
    for (int i = 1; i < n_ranks; ++i)
      if (devices[i] != devices[0]) {
        for (qp_ctx_t cc : E->ctx) qp_ctx_destroy(cc);  /* deferred until E releases them */
        delete E;
        return qp_fail(nullptr, QP_ERR_UNSUPPORTED,
                       "qp_ens_create: ranks must sit on distinct devices (NCCL) or all on the same device (fake ranks)");
      }
  }
  *out = E;
  return QP_OK;
}

extern "C" int32_t qp_ens_unique_id(uint8_t* id /*[128]*/) {
  if (!id) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_unique_id: null pointer");
  NcclApi* api = nccl_api();
  if (!api) return qp_fail(nullptr, QP_ERR_UNSUPPORTED, "qp_ens_unique_id: NCCL is not available");
  ncclUniqueId uid;
  ncclResult_t r = api->GetUniqueId(&uid);
  if (r != ncclSuccess) return qp_fail(nullptr, QP_ERR_CUDA, "ncclGetUniqueId failed: %s", api->GetErrorString(r));
  static_assert(sizeof(uid) == 128, "ncclUniqueId is 128 bytes");
  memcpy(id, &uid, 128);
  return QP_OK;
}

extern "C" int32_t qp_ens_create_rank(qp_ctx_t ctx, int32_t rank, int32_t n_ranks, const uint8_t* id, qp_ens_t* out) {
  if (!ctx) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_create_rank: null context");
  QP_REQUIRE(ctx, out != nullptr && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "qp_ens_create_rank: rank %d outside a world of %d", rank, n_ranks);
  *out = nullptr;
  QP_CHECK(qp_ctx_bind(ctx));
  qp_ens_t E = new qp_ens_s();
  E->n_ranks = n_ranks;
  E->ctx.emplace_back();
  E->ctx.back() = ctx;
  E->own_ctx.push_back(false);
  E->local_rank.push_back(rank);
  E->d_stage.assign(1, nullptr);
  E->stage_elems.assign(1, 0);
  if (n_ranks > 1) {
    NcclApi* api = nccl_api();
    if (!api || !id) {
      delete E;
      return qp_fail(ctx, QP_ERR_UNSUPPORTED, "qp_ens_create_rank: NCCL is not available or no id given");
    }
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    E->comm.resize(1);
    ncclResult_t r = api->CommInitRank(&E->comm[0], n_ranks, uid, rank);
    if (r != ncclSuccess) {
      delete E;
      return qp_fail(ctx, QP_ERR_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r));
    }
    E->use_nccl = true;
    int32_t wrc = ens_warm_up(E);
    if (wrc != QP_OK) {
      qp_ens_destroy(E);
      return wrc;
    }
  }
  *out = E;
  return QP_OK;
}

extern "C" int32_t qp_ens_destroy(qp_ens_t E) {
  if (!E) return QP_OK;
  for (size_t i = 0; i < E->ctx.size(); ++i) {
    cudaSetDevice(E->ctx[i]->device);
    cudaStreamSynchronize(E->ctx[i]->stream);
    cudaFree(E->d_stage[i]);
  }
  if (E->use_nccl)
    if (NcclApi* api = nccl_api())
      for (ncclComm_t c : E->comm) api->CommDestroy(c);
  for (size_t i = 0; i < E->ctx.size(); ++i)
    if (E->own_ctx[i]) qp_ctx_destroy(E->ctx[i]);  // deferred while objects (or this ensemble) still hold it
  delete E;
  return QP_OK;
}

extern "C" int32_t qp_ens_info(qp_ens_t E, int32_t* n_ranks, int32_t* n_local, int32_t* transport) {
  if (!E) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_info: null ensemble");
  if (n_ranks) *n_ranks = E->n_ranks;
  if (n_local) *n_local = (int32_t)E->ctx.size();
  if (transport) *transport = E->use_nccl ? 1 : 0;
  return QP_OK;
}

extern "C" int32_t qp_ens_ctx(qp_ens_t E, int32_t local_index, qp_ctx_t* ctx, int32_t* rank) {
  if (!E || local_index < 0 || local_index >= (int)E->ctx.size() || !ctx)
    return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_ctx: bad local index %d", local_index);
  *ctx = E->ctx[local_index];
  if (rank) *rank = E->local_rank[local_index];
  return QP_OK;
}

// full[row][off_r + b] = stage[block_r][row][b]: the rank blocks side by side in every row
__global__ void k_ens_unshard(const double2* __restrict__ stage, double2* __restrict__ full, int64_t n, int64_t b_total,
                              int n_ranks, const int64_t* __restrict__ off /*[n_ranks+1] column offsets*/) {
  const int64_t total = n * b_total;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / b_total, col = i - row * b_total;
    int r = 0;
    while (r + 1 < n_ranks && col >= off[r + 1]) ++r;
    const int64_t br = off[r + 1] - off[r];
    full[i] = stage[off[r] * n + row * br + (col - off[r])];
  }
}

// Gathers `width` complex numbers per (item, trajectory): every rank contributes a contiguous device
// buffer of items * count_r elements laid out [items][count_r]; on return stage_i (device, per
// local member) holds the blocks of all ranks back to back: block r at element offset items * off[r].
static int32_t ens_exchange(qp_ens_t E, const std::vector<const double2*>& src, int64_t items,
                            const std::vector<int64_t>& off /*[n_ranks+1]*/) {
  const int nl = (int)E->ctx.size();
  const size_t total = (size_t)items * (size_t)off[E->n_ranks];
  for (int i = 0; i < nl; ++i) {
    qp_ctx_t ctx = E->ctx[i];
    QP_CHECK(qp_ctx_bind(ctx));
    if (E->stage_elems[i] < total) {
      QP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(E->d_stage[i]);
      E->d_stage[i] = nullptr;
      E->stage_elems[i] = 0;
      QP_CUDA(ctx, cudaMalloc(&E->d_stage[i], sizeof(double2) * std::max<size_t>(total, 1)));
      E->stage_elems[i] = total;
    }
  }
  if (E->use_nccl) {
    NcclApi* api = nccl_api();
    QP_NCCL(E->ctx[0], api, api->GroupStart());
    for (int i = 0; i < nl; ++i) {
      qp_ctx_t ctx = E->ctx[i];
      QP_CHECK(qp_ctx_bind(ctx));
      for (int r = 0; r < E->n_ranks; ++r) {
        const size_t cnt = (size_t)items * (size_t)(off[r + 1] - off[r]) * 2;  // doubles
        if (cnt == 0) continue;
        double2* dst = E->d_stage[i] + (size_t)items * (size_t)off[r];
        const void* snd = (E->local_rank[i] == r) ? (const void*)src[i] : (const void*)dst;
        QP_NCCL(ctx, api, api->Broadcast(snd, dst, cnt, ncclDouble, r, E->comm[i], ctx->stream));
      }
    }
    QP_NCCL(E->ctx[0], api, api->GroupEnd());
  } else {
    // all members in this process: plain copies (same device for fake ranks)
    for (int i = 0; i < nl; ++i) QP_CUDA(E->ctx[i], cudaStreamSynchronize(E->ctx[i]->stream));  // sources complete
    for (int i = 0; i < nl; ++i) {
      qp_ctx_t ctx = E->ctx[i];
      QP_CHECK(qp_ctx_bind(ctx));
      for (int j = 0; j < nl; ++j) {
        const int r = E->local_rank[j];
        const size_t bytes = sizeof(double2) * (size_t)items * (size_t)(off[r + 1] - off[r]);
        if (bytes == 0) continue;
        QP_CUDA(ctx, cudaMemcpyAsync(E->d_stage[i] + (size_t)items * (size_t)off[r], src[j], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      }
    }
  }
  return QP_OK;
}

static int32_t ens_warm_up(qp_ens_t E) {
  // one trajectory per rank, one number each: the same grouped broadcasts a gather issues
  const int nl = (int)E->ctx.size();
  std::vector<int64_t> off((size_t)E->n_ranks + 1);
  for (int r = 0; r <= E->n_ranks; ++r) off[r] = r;
  std::vector<double2*> d_src(nl, nullptr);
  std::vector<const double2*> src(nl);
  int32_t rc = QP_OK;
  for (int i = 0; i < nl && rc == QP_OK; ++i) {
    rc = qp_ctx_bind(E->ctx[i]);
    if (rc == QP_OK && cudaMalloc(&d_src[i], sizeof(double2)) != cudaSuccess) rc = qp_fail(E->ctx[i], QP_ERR_OOM, "qp_ens: cudaMalloc failed");
    if (rc == QP_OK) cudaMemsetAsync(d_src[i], 0, sizeof(double2), E->ctx[i]->stream);
    src[i] = d_src[i];
  }
  if (rc == QP_OK) rc = ens_exchange(E, src, 1, off);
  for (int i = 0; i < nl; ++i) {
    cudaSetDevice(E->ctx[i]->device);
    cudaStreamSynchronize(E->ctx[i]->stream);
    cudaFree(d_src[i]);
  }
  return rc;
}

static int32_t ens_offsets(qp_ens_t E, int64_t n_total, std::vector<int64_t>& off) {
  off.assign((size_t)E->n_ranks + 1, 0);
  for (int r = 0; r < E->n_ranks; ++r) {
    int64_t b0, b1;
    QP_CHECK(qp_ens_shard(n_total, r, E->n_ranks, &b0, &b1));
    off[r] = b0;
    off[r + 1] = b1;
  }
  return QP_OK;
}

extern "C" int32_t qp_ens_gather_states(qp_ens_t E, const qp_state_t* local_states, int64_t n_total,
                                        const qp_state_t* full_states, qp_c128* host_out) {
  if (!E || !local_states) return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_gather_states: null argument");
  const int nl = (int)E->ctx.size();
  std::vector<int64_t> off;
  QP_CHECK(ens_offsets(E, n_total, off));
  const int64_t n = local_states[0] ? local_states[0]->n : 0;
  std::vector<const double2*> src(nl);
  for (int i = 0; i < nl; ++i) {
    qp_state_t s = local_states[i];
    const int r = E->local_rank[i];
    QP_REQUIRE(E->ctx[i], s != nullptr && s->ctx == E->ctx[i] && s->n == n && s->batch == off[r + 1] - off[r],
               "qp_ens_gather_states: local state %d must be a %lld x %lld state of its rank's context", i, (long long)n,
               (long long)(off[r + 1] - off[r]));
    if (full_states)
      QP_REQUIRE(E->ctx[i], full_states[i] && full_states[i]->ctx == E->ctx[i] && full_states[i]->n == n && full_states[i]->batch == n_total,
                 "qp_ens_gather_states: full state %d must be %lld x %lld", i, (long long)n, (long long)n_total);
    src[i] = s->d;
  }
  QP_CHECK(ens_exchange(E, src, n, off));
  // rank blocks [n][b_r] -> rows of the full [n][B_total] state
  for (int i = 0; i < nl; ++i) {
    if (!full_states && !(host_out && i == 0)) continue;
    qp_ctx_t ctx = E->ctx[i];
    QP_CHECK(qp_ctx_bind(ctx));
    int64_t* d_off = nullptr;
    QP_CUDA(ctx, cudaMalloc(&d_off, sizeof(int64_t) * off.size()));
    QP_CUDA(ctx, cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice, ctx->stream));
    double2* dst = nullptr;
    double2* tmp = nullptr;
    if (full_states) dst = full_states[i]->d;
    else {
      QP_CUDA(ctx, cudaMalloc(&tmp, sizeof(double2) * (size_t)n * (size_t)n_total));
      dst = tmp;
    }
    const int64_t total = n * n_total;
    const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_ens_unshard<<<blocks, 256, 0, ctx->stream>>>(E->d_stage[i], dst, n, n_total, E->n_ranks, d_off);
    ctx->launches++;
    cudaError_t le = cudaGetLastError();
    if (le == cudaSuccess && host_out && i == 0)
      le = cudaMemcpyAsync(host_out, dst, sizeof(double2) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream);
    if (le == cudaSuccess) le = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_off);
    cudaFree(tmp);
    QP_CUDA(ctx, le);
  }
  for (int i = 0; i < nl; ++i) QP_CUDA(E->ctx[i], cudaStreamSynchronize(E->ctx[i]->stream));
  return QP_OK;
}

extern "C" int32_t qp_ens_gather_expvals(qp_ens_t E, const qp_c128* const* local_values, int32_t n_values,
                                         int64_t n_total, qp_c128* out) {
  if (!E || !local_values || !out || n_values < 1)
    return qp_fail(nullptr, QP_ERR_INVALID_ARG, "qp_ens_gather_expvals: bad arguments");
  const int nl = (int)E->ctx.size();
  std::vector<int64_t> off;
  QP_CHECK(ens_offsets(E, n_total, off));
  // upload every member's [n_values][b_local] block, exchange, download the concatenation
  std::vector<double2*> d_src(nl, nullptr);
  std::vector<const double2*> src(nl);
  int32_t rc = QP_OK;
  for (int i = 0; i < nl && rc == QP_OK; ++i) {
    qp_ctx_t ctx = E->ctx[i];
    rc = qp_ctx_bind(ctx);
    const int r = E->local_rank[i];
    const size_t elems = (size_t)n_values * (size_t)(off[r + 1] - off[r]);
    if (rc == QP_OK && cudaMalloc(&d_src[i], sizeof(double2) * std::max<size_t>(elems, 1)) != cudaSuccess)
      rc = qp_fail(ctx, QP_ERR_OOM, "qp_ens_gather_expvals: cudaMalloc failed");
    if (rc == QP_OK && elems &&
        cudaMemcpyAsync(d_src[i], local_values[i], sizeof(double2) * elems, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      rc = qp_fail(ctx, QP_ERR_CUDA, "qp_ens_gather_expvals: upload failed");
    src[i] = d_src[i];
  }
  if (rc == QP_OK) rc = ens_exchange(E, src, n_values, off);
  std::vector<qp_c128> blocks((size_t)n_values * (size_t)n_total);
  if (rc == QP_OK) {
    qp_ctx_t ctx = E->ctx[0];
    rc = qp_ctx_bind(ctx);
    if (rc == QP_OK && !blocks.empty() &&
        cudaMemcpyAsync(blocks.data(), E->d_stage[0], sizeof(double2) * blocks.size(), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
      rc = qp_fail(ctx, QP_ERR_CUDA, "qp_ens_gather_expvals: download failed");
  }
  for (int i = 0; i < nl; ++i) {
    cudaSetDevice(E->ctx[i]->device);
    cudaStreamSynchronize(E->ctx[i]->stream);
    cudaFree(d_src[i]);
  }
  if (rc != QP_OK) return rc;
  // blocks: rank r at offset n_values * off[r], laid out [n_values][b_r]  ->  out [n_values][B_total]
  for (int r = 0; r < E->n_ranks; ++r) {
    const int64_t br = off[r + 1] - off[r];
    for (int v = 0; v < n_values; ++v)
      for (int64_t b = 0; b < br; ++b)
        out[(size_t)v * n_total + off[r] + b] = blocks[(size_t)n_values * off[r] + (size_t)v * br + b];
  }
  return QP_OK;
}
