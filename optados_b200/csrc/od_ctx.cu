// od_ctx.cu -- context lifetime, device buffers, module-data setters and the NCCL plumbing that
// replaces comms_reduce (comms.F90:556-602).  NCCL is resolved at run time (dlopen) so that the
// library loads on single-GPU hosts without NCCL and shares the copy a host program already loaded.
#include <dlfcn.h>
#include <nccl.h>

#include <cfloat>
#include <cmath>

#include "od_internal.h"

// ------------------------------------------------------------------ errors / buffers
int od_fail(od_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  else fprintf(stderr, "optados_b200: %s\n", buf);
  return code;
}

namespace {
__global__ void poison_kernel(int *p, size_t n, int v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
}  // namespace

int od_buf_ensure(od_ctx *ctx, DevBuf &b, size_t bytes) {
  if (b.owned && b.bytes >= bytes && b.p) return OD_OK;
  if (b.owned && b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  b.owned = true;
  if (bytes == 0) bytes = 8;
  OD_CUDA(ctx, cudaMalloc(&b.p, bytes));
  b.bytes = bytes;
  // tests only: OD_DEBUG_POISON=<int> fills every fresh allocation with that 32-bit pattern, so that code which reads a
  // buffer it never wrote (or that a re-allocation emptied) fails loudly instead of passing on a recycled address
  if (const char *poison = getenv("OD_DEBUG_POISON")) {
    poison_kernel<<<256, 256, 0, ctx ? ctx->stream : nullptr>>>(static_cast<int *>(b.p), bytes / sizeof(int), atoi(poison));
    OD_CUDA(ctx, cudaStreamSynchronize(ctx ? ctx->stream : nullptr));
  }
  return OD_OK;
}

void od_buf_release(DevBuf &b) {
  if (b.owned && b.p) cudaFree(b.p);
  b.p = nullptr;
  b.bytes = 0;
  b.owned = false;
}

int od_stage_in(od_ctx *ctx, const void *src, size_t bytes, int mem, DevBuf &stage, const void **dptr) {
  if (!src) return od_fail(ctx, OD_ERR_ARG, "null input array");
  if (mem == OD_MEM_DEVICE) {
    *dptr = src;
    return OD_OK;
  }
  if (mem != OD_MEM_HOST) return od_fail(ctx, OD_ERR_ARG, "unknown memory kind %d", mem);
  OD_CHECK(od_buf_ensure(ctx, stage, bytes));
  OD_CUDA(ctx, cudaMemcpyAsync(stage.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dptr = stage.p;
  return OD_OK;
}

// ------------------------------------------------------------------ NCCL (run-time bound)
namespace {
struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;

bool nccl_load() {
  if (g_nccl.ok) return true;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // reuse an already loaded copy (e.g. torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_nccl.h = h;
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(h, "ncclAllReduce"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.GetErrorString;
  return g_nccl.ok;
}
}  // namespace

// sizeof of the public structures: a binding in another language (the ISO_C_BINDING derived types of the Fortran shim,
// the ctypes mirrors of the tests) checks its own layout against these at start-up
extern "C" int od_abi_sizeof(const char *name) {
  if (!name) return -1;
#define OD_SZ(T) if (!strcmp(name, #T)) return (int)sizeof(T)
  OD_SZ(od_broadening); OD_SZ(od_dos_params); OD_SZ(od_dos_result); OD_SZ(od_jdos_params); OD_SZ(od_jdos_result);
  OD_SZ(od_optics_params); OD_SZ(od_optics_result); OD_SZ(od_core_params); OD_SZ(od_core_result); OD_SZ(od_bandgap_result);
  OD_SZ(od_odi); OD_SZ(od_bands_info); OD_SZ(od_pdos_info); OD_SZ(od_elnes_info);
#undef OD_SZ
  return -1;
}

extern "C" const char *od_version(void) { return "optados_b200 0.1 (sm_100a, FP64)"; }

extern "C" int od_comm_get_unique_id(void *id128) {
  if (!id128) return OD_ERR_ARG;
  if (!nccl_load()) return od_fail(nullptr, OD_ERR_NCCL, "cannot load libnccl.so.2");
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id is 128 bytes");
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return od_fail(nullptr, OD_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
  memcpy(id128, &id, 128);
  return OD_OK;
}

// debugging aid (OD_DEBUG_BACKTRACE=1): raw return addresses on SIGSEGV, to be resolved with addr2line against the .so
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void od_segv_handler(int sig) {
  void *frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "optados_b200: SIGSEGV, backtrace:\n";
  if (write(2, msg, sizeof(msg) - 1) < 0) {}
  backtrace_symbols_fd(frames, n, 2);
  signal(sig, SIG_DFL);
  raise(sig);
}

static int ctx_create_common(int device, od_ctx **out) {
  if (!out) return OD_ERR_ARG;
  if (getenv("OD_DEBUG_BACKTRACE")) signal(SIGSEGV, od_segv_handler);
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return od_fail(nullptr, OD_ERR_CUDA, "no CUDA device available (%s); optados_b200 has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= ndev) return od_fail(nullptr, OD_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
  od_ctx *ctx = new od_ctx();
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
    int rc = od_fail(nullptr, OD_ERR_CUDA, "context setup failed: %s", cudaGetErrorString(e));
    delete ctx;
    return rc;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  *out = ctx;
  return OD_OK;
}

extern "C" int od_ctx_create(int device, od_ctx **out) { return ctx_create_common(device, out); }

extern "C" int od_ctx_create_rank(int device, int rank, int nranks, const void *id128, od_ctx **out) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return od_fail(nullptr, OD_ERR_ARG, "bad rank %d of %d", rank, nranks);
  OD_CHECK(ctx_create_common(device, out));
  od_ctx *ctx = *out;
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks > 1) {
    if (!id128) return od_fail(ctx, OD_ERR_ARG, "nranks > 1 needs an NCCL unique id");
    if (!nccl_load()) return od_fail(ctx, OD_ERR_NCCL, "cannot load libnccl.so.2");
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return od_fail(ctx, OD_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    ctx->comm = comm;
  }
  return OD_OK;
}

extern "C" int od_ctx_destroy(od_ctx *ctx) {
  if (!ctx) return OD_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm && g_nccl.ok) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(ctx->comm));
  DevBuf *bufs[] = {&ctx->kweight, &ctx->bands, &ctx->grads, &ctx->partial, &ctx->accum, &ctx->small, &ctx->recs,
                    &ctx->keys, &ctx->scratch_a, &ctx->scratch_b, &ctx->scratch_c, &ctx->scratch_d, &ctx->mw_tmp, &ctx->in_tmp, &ctx->in_tmp2, &ctx->moments, &ctx->task_mw, &ctx->flags, &ctx->stats_dev, &ctx->dest_dev};
  for (DevBuf *b : bufs) od_buf_release(*b);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  for (const od_ctx::PendingTime &t : ctx->pending_times) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->t0) cudaEventDestroy(ctx->t0);
  if (ctx->t1) cudaEventDestroy(ctx->t1);
  for (int i = 0; i < 2; ++i) {
    od_buf_release(ctx->gslab[i]);
    if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
  }
  if (ctx->host_stage) cudaFreeHost(ctx->host_stage);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return OD_OK;
}

extern "C" const char *od_last_error(const od_ctx *ctx) { return ctx ? ctx->err.c_str() : ""; }
extern "C" int od_ctx_rank(const od_ctx *ctx) { return ctx ? ctx->rank : 0; }
extern "C" int od_ctx_nranks(const od_ctx *ctx) { return ctx ? ctx->nranks : 1; }
extern "C" long long od_kernel_launches(const od_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int od_set_engine(od_ctx *ctx, int engine) {
  if (!ctx) return OD_ERR_ARG;
  if (engine != OD_ENGINE_AUTO && engine != OD_ENGINE_DENSE) return od_fail(ctx, OD_ERR_ARG, "unknown engine %d", engine);
  ctx->engine = engine;
  return OD_OK;
}

extern "C" int od_synchronize(od_ctx *ctx) {
  if (!ctx) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

int od_flags_ready(od_ctx *ctx) {
  if (ctx->flags.p && ctx->stats_dev.p) return OD_OK;
  OD_CHECK(od_buf_ensure(ctx, ctx->flags, sizeof(int) * 4));
  OD_CHECK(od_buf_ensure(ctx, ctx->stats_dev, sizeof(unsigned long long) * 2));
  OD_CUDA(ctx, cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * 4, ctx->stream));
  OD_CUDA(ctx, cudaMemsetAsync(ctx->stats_dev.p, 0, sizeof(unsigned long long) * 2, ctx->stream));
  return OD_OK;
}

int od_check_deferred(od_ctx *ctx) {
  if (!ctx->flags.p) return OD_OK;
  int h[4] = {0, 0, 0, 0};
  OD_CUDA(ctx, cudaMemcpyAsync(h, ctx->flags.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (h[0]) {
    OD_CUDA(ctx, cudaMemsetAsync(ctx->flags.p, 0, sizeof(int) * 4, ctx->stream));
    return od_fail(ctx, OD_ERR_DOSLIN, "doslin: input arguments incorrect (dos_utils.f90:1431/1514)");
  }
  return OD_OK;
}

int od_time_begin(od_ctx *ctx, int kind, size_t *slot) {
  od_ctx::PendingTime t;
  for (cudaEvent_t *e : {&t.e0, &t.e1}) {
    if (!ctx->event_pool.empty()) { *e = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
    else OD_CUDA(ctx, cudaEventCreate(e));
  }
  t.kind = kind;
  OD_CUDA(ctx, cudaEventRecord(t.e0, ctx->stream));
  ctx->pending_times.push_back(t);
  *slot = ctx->pending_times.size() - 1;
  return OD_OK;
}

int od_time_end(od_ctx *ctx, size_t slot) {
  OD_CUDA(ctx, cudaEventRecord(ctx->pending_times[slot].e1, ctx->stream));
  return OD_OK;
}

int od_times_resolve(od_ctx *ctx) {
  for (const od_ctx::PendingTime &t : ctx->pending_times) {
    float ms = 0.f;
    if (cudaEventSynchronize(t.e1) == cudaSuccess && cudaEventElapsedTime(&ms, t.e0, t.e1) == cudaSuccess)
      (t.kind == 0 ? ctx->prof.prepare_ms : ctx->prof.accumulate_ms) += ms;
    ctx->event_pool.push_back(t.e0);
    ctx->event_pool.push_back(t.e1);
  }
  ctx->pending_times.clear();
  return OD_OK;
}

extern "C" int od_profile_reset(od_ctx *ctx) {
  if (!ctx) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OD_CHECK(od_times_resolve(ctx));
  ctx->prof = Profile();
  if (ctx->stats_dev.p) OD_CUDA(ctx, cudaMemsetAsync(ctx->stats_dev.p, 0, sizeof(unsigned long long) * 2, ctx->stream));
  return OD_OK;
}

extern "C" int od_profile_get(od_ctx *ctx, double *accumulate_ms, double *prepare_ms, double *contributions, double *lane_steps,
                              long long *accumulate_launches) {
  if (!ctx) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OD_CHECK(od_times_resolve(ctx));
  if (ctx->stats_dev.p) {
    unsigned long long h[2] = {0, 0};
    OD_CUDA(ctx, cudaMemcpyAsync(h, ctx->stats_dev.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->prof.lane_steps = (double)h[0];
    ctx->prof.contributions = (double)h[1];
  }
  if (accumulate_ms) *accumulate_ms = ctx->prof.accumulate_ms;
  if (prepare_ms) *prepare_ms = ctx->prof.prepare_ms;
  if (contributions) *contributions = ctx->prof.contributions;
  if (lane_steps) *lane_steps = ctx->prof.lane_steps;
  if (accumulate_launches) *accumulate_launches = ctx->prof.accumulate_launches;
  return OD_OK;
}

extern "C" int od_device_malloc(od_ctx *ctx, size_t nbytes, void **dptr) {
  if (!ctx || !dptr) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OD_CUDA(ctx, cudaMalloc(dptr, nbytes ? nbytes : 8));
  return OD_OK;
}
extern "C" int od_device_free(od_ctx *ctx, void *dptr) {
  if (!ctx) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  OD_CUDA(ctx, cudaFree(dptr));
  return OD_OK;
}
extern "C" int od_memcpy_h2d(od_ctx *ctx, void *dst, const void *src, size_t n) {
  if (!ctx) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}
extern "C" int od_memcpy_d2h(od_ctx *ctx, void *dst, const void *src, size_t n) {
  if (!ctx) return OD_ERR_ARG;
  OD_CUDA(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

// algorithms.f90:309-340
extern "C" int od_dist_array(int num_elements, int num_nodes, int *per) {
  if (num_nodes < 1 || !per) return OD_ERR_ARG;
  for (int i = 0; i < num_nodes; ++i) per[i] = num_elements / num_nodes;
  if (num_elements < num_nodes) return OD_ERR_REF;  // 'Fewer kpoints than nodes' (:329)
  const int rem = num_elements - per[0] * num_nodes;
  for (int i = 0; i < rem; ++i) per[i] += 1;
  return OD_OK;
}

int od_allreduce_device(od_ctx *ctx, double *dbuf, size_t n, char op) {
  if (ctx->nranks == 1 || n == 0) return OD_OK;
  if (!ctx->comm || !g_nccl.ok) return od_fail(ctx, OD_ERR_NCCL, "context has no communicator");
  ncclRedOp_t rop = (op == 's') ? ncclSum : (op == 'm') ? ncclMin : (op == 'M') ? ncclMax : ncclNumOps;
  if (rop == ncclNumOps) return od_fail(ctx, OD_ERR_ARG, "unknown reduction '%c'", op);
  ncclResult_t r = g_nccl.AllReduce(dbuf, dbuf, n, ncclDouble, rop, reinterpret_cast<ncclComm_t>(ctx->comm), ctx->stream);
  if (r != ncclSuccess) return od_fail(ctx, OD_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString(r));
  return OD_OK;
}

extern "C" int od_allreduce(od_ctx *ctx, double *host_buf, size_t n, char op) {
  if (!ctx || (!host_buf && n)) return OD_ERR_ARG;
  if (ctx->nranks == 1 || n == 0) return OD_OK;
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  OD_CHECK(od_buf_ensure(ctx, ctx->small, n * sizeof(double)));
  OD_CUDA(ctx, cudaMemcpyAsync(ctx->small.p, host_buf, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  OD_CHECK(od_allreduce_device(ctx, ctx->small.as<double>(), n, op));
  OD_CUDA(ctx, cudaMemcpyAsync(host_buf, ctx->small.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}

// ------------------------------------------------------------------ module data
extern "C" void od_broadening_defaults(od_broadening *b) {
  if (!b) return;
  b->adaptive_smearing = 0.4;       // parameters.f90:246
  b->fixed_smearing = 0.3;          // parameters.f90:249
  b->linear_smearing = 0.0;         // parameters.f90:252
  b->finite_bin_correction = 1;     // parameters.f90:275
  b->numerical_intdos = 0;          // parameters.f90:280
  b->hybrid_linear = 0;             // parameters.f90:286
  b->hybrid_linear_grad_tol = 0.01; // parameters.f90:289
}

static int set_array(od_ctx *ctx, DevBuf &dst, const double *src, size_t n, int mem) {
  if (!src) return od_fail(ctx, OD_ERR_ARG, "null array");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  if (mem == OD_MEM_DEVICE) {
    od_buf_release(dst);
    dst.p = const_cast<double *>(src);
    dst.bytes = n * sizeof(double);
    dst.owned = false;
    return OD_OK;
  }
  if (mem != OD_MEM_HOST) return od_fail(ctx, OD_ERR_ARG, "unknown memory kind %d", mem);
  OD_CHECK(od_buf_ensure(ctx, dst, n * sizeof(double)));
  OD_CUDA(ctx, cudaMemcpyAsync(dst.p, src, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  return OD_OK;
}

extern "C" int od_set_system(od_ctx *ctx, int nk_local, int ns, int nb, const double recip_lattice[9],
                             const int kpoint_grid_dim[3], const double *kpoint_weight, double electrons_per_state) {
  if (!ctx) return OD_ERR_ARG;
  if (nk_local < 1 || nb < 1 || ns < 1 || ns > 2) return od_fail(ctx, OD_ERR_ARG, "bad dimensions nk=%d ns=%d nb=%d", nk_local, ns, nb);
  if (!recip_lattice || !kpoint_grid_dim) return od_fail(ctx, OD_ERR_ARG, "null lattice / grid");
  for (int i = 0; i < 3; ++i)
    if (kpoint_grid_dim[i] < 1) return od_fail(ctx, OD_ERR_ARG, "kpoint_grid_dim(%d) = %d", i + 1, kpoint_grid_dim[i]);
  ctx->nk = nk_local;
  ctx->ns = ns;
  ctx->nb = nb;
  memcpy(ctx->recip, recip_lattice, sizeof(double) * 9);
  memcpy(ctx->kgrid, kpoint_grid_dim, sizeof(int) * 3);
  ctx->eps = electrons_per_state;
  if (kpoint_weight) OD_CHECK(set_array(ctx, ctx->kweight, kpoint_weight, (size_t)nk_local, OD_MEM_HOST));
  ctx->have_system = true;
  return OD_OK;
}

extern "C" int od_set_bands(od_ctx *ctx, const double *band_energy, int mem) {
  if (!ctx) return OD_ERR_ARG;
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system must be called first");
  return set_array(ctx, ctx->bands, band_energy, (size_t)ctx->nb * ctx->ns * ctx->nk, mem);
}

extern "C" int od_set_gradients(od_ctx *ctx, const double *band_gradient, int mem) {
  if (!ctx) return OD_ERR_ARG;
  if (!ctx->have_system) return od_fail(ctx, OD_ERR_STATE, "od_set_system must be called first");
  ctx->grads_host = nullptr;
  if (mem == OD_MEM_HOST_DEFERRED) {
    if (!band_gradient) return od_fail(ctx, OD_ERR_ARG, "null array");
    od_buf_release(ctx->grads);
    ctx->grads_host = band_gradient;
    return OD_OK;
  }
  return set_array(ctx, ctx->grads, band_gradient, (size_t)ctx->nb * 3 * ctx->nk * ctx->ns, mem);
}

int od_grads_resident(od_ctx *ctx) {
  if (!ctx->grads_host) return OD_OK;
  const double *h = ctx->grads_host;
  ctx->grads_host = nullptr;
  return set_array(ctx, ctx->grads, h, (size_t)ctx->nb * 3 * ctx->nk * ctx->ns, OD_MEM_HOST);
}

static int get_array(od_ctx *ctx, const DevBuf &src, double *dst, size_t n, const char *what) {
  if (!ctx || !dst) return OD_ERR_ARG;
  if (!src.p) return od_fail(ctx, OD_ERR_STATE, "%s not set", what);
  OD_CUDA(ctx, cudaMemcpyAsync(dst, src.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return OD_OK;
}
extern "C" int od_get_bands(od_ctx *ctx, double *h) { return get_array(ctx, ctx->bands, h, (size_t)ctx->nb * ctx->ns * ctx->nk, "band_energy"); }
extern "C" int od_get_gradients(od_ctx *ctx, double *h) { if (ctx) { int rc = od_grads_resident(ctx); if (rc) return rc; } return get_array(ctx, ctx->grads, h, (size_t)ctx->nb * 3 * ctx->nk * ctx->ns, "band_gradient"); }
extern "C" int od_get_kweights(od_ctx *ctx, double *h) { return get_array(ctx, ctx->kweight, h, (size_t)ctx->nk, "kpoint_weight"); }

// dos_utils.f90:1210-1216
GridGeom od_grid_geom(const od_ctx *ctx, double adaptive_smearing) {
  GridGeom g;
  double sub[3];
  for (int i = 0; i < 3; ++i) g.step[i] = 1.0 / (double)ctx->kgrid[i] / 2.0;
  for (int i = 0; i < 3; ++i) {
    const double a = ctx->recip[i + 0], b = ctx->recip[i + 3], c = ctx->recip[i + 6];
    sub[i] = std::sqrt(a * a + b * b + c * c) * g.step[i];
  }
  g.adaptive_smearing_temp = adaptive_smearing * (sub[0] + sub[1] + sub[2]) / 3.0;
  memcpy(g.recip, ctx->recip, sizeof(double) * 9);
  return g;
}

// ------------------------------------------------------------------ min / max of band_energy
namespace {
__global__ void minmax_kernel(const double *__restrict__ a, size_t n, double *__restrict__ out /* [2*gridDim.x] */) {
  double lo = DBL_MAX, hi = -DBL_MAX;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = a[i];
    lo = fmin(lo, v);
    hi = fmax(hi, v);
  }
  __shared__ double slo[32], shi[32];
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { slo[w] = lo; shi[w] = hi; }
  __syncthreads();
  if (w == 0) {
    const int nw = blockDim.x >> 5;
    lo = (l < nw) ? slo[l] : DBL_MAX;
    hi = (l < nw) ? shi[l] : -DBL_MAX;
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (l == 0) { out[2 * blockIdx.x] = lo; out[2 * blockIdx.x + 1] = hi; }
  }
}

// (min, -max) of the per-CTA pairs, so that one MIN allreduce serves both (dos_utils.f90:968-977)
__global__ void __launch_bounds__(256) minmax_final_kernel(const double *__restrict__ part, int nparts, double *__restrict__ out2) {
  double lo = DBL_MAX, hi = -DBL_MAX;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) { lo = fmin(lo, part[2 * i]); hi = fmax(hi, part[2 * i + 1]); }
  __shared__ double slo[256], shi[256];
  slo[threadIdx.x] = lo; shi[threadIdx.x] = hi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { slo[threadIdx.x] = fmin(slo[threadIdx.x], slo[threadIdx.x + o]); shi[threadIdx.x] = fmax(shi[threadIdx.x], shi[threadIdx.x + o]); }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = slo[0]; out2[1] = -shi[0]; }
}
}  // namespace

extern "C" int od_band_minmax(od_ctx *ctx, double *mn, double *mx) {
  if (!ctx || !mn || !mx) return OD_ERR_ARG;
  if (!ctx->bands.p) return od_fail(ctx, OD_ERR_STATE, "band_energy not set");
  OD_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = (size_t)ctx->nb * ctx->ns * ctx->nk;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
  OD_CHECK(od_buf_ensure(ctx, ctx->small, sizeof(double) * 2 * (size_t)blocks + 64));
  double *part = ctx->small.as<double>(), *out2 = part + 2 * (size_t)blocks;
  minmax_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bands.as<double>(), n, part);
  OD_LAUNCH_CHECK(ctx);
  minmax_final_kernel<<<1, 256, 0, ctx->stream>>>(part, blocks, out2);
  OD_LAUNCH_CHECK(ctx);
  // comms_reduce MIN / MAX + bcast (dos_utils.f90:968-977): one MIN allreduce of (min, -max) on the device
  OD_CHECK(od_allreduce_device(ctx, out2, 2, 'm'));
  double v[2] = {0.0, 0.0};
  OD_CUDA(ctx, cudaMemcpyAsync(v, out2, sizeof(v), cudaMemcpyDeviceToHost, ctx->stream));
  OD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *mn = v[0];
  *mx = -v[1];
  return OD_OK;
}
