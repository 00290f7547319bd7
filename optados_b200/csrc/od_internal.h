// od_internal.h -- context, device buffers and launch bookkeeping of liboptados_b200.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/optados_b200.h"

// ------------------------------------------------------------------ device buffer
struct DevBuf {
  void *p = nullptr;
  size_t bytes = 0;
  bool owned = false;
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Profile {
  double accumulate_ms = 0.0, prepare_ms = 0.0, contributions = 0.0, lane_steps = 0.0;
  long long accumulate_launches = 0;
};

// results of the task drivers (the reference's module-global dos_*/intdos_*/jdos_* arrays)
struct Spectra {
  int nbins = 0, ns = 0, norb = 0;
  double emin = 0.0, delta = 0.0;         // E(i) = emin + (i-1)*delta
  std::vector<double> dos[3], intdos[3];  // index 0 fixed, 1 adaptive, 2 linear
  bool have[3] = {false, false, false};
  std::vector<double> weighted;
  bool have_weighted = false;
};

// results of the task mirrors (od_tasks.cu)
struct OpticsRes {
  int nbins = 0, n_geom = 0;
  double delta = 0.0;
  std::vector<double> E, epsilon, conduct, refract, loss_fn, absorp, reflect;
};
struct CoreRes {
  int nbins = 0, ns = 0, norb = 0;
  double efermi = 0.0;
  std::vector<double> E, E_shift, wdos, wdos_broadened;
};

struct od_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t t0 = nullptr, t1 = nullptr;  // od_timer_start / od_timer_stop
  int rank = 0, nranks = 1;
  void *comm = nullptr;  // ncclComm_t
  std::string err;
  long long launches = 0;
  Profile prof;
  int sm_count = 148;
  int engine = 0;  // OD_ENGINE_*

  // module data (od_cell / od_electronic)
  bool have_system = false;
  int nk = 0, ns = 0, nb = 0;
  double recip[9] = {0};
  int kgrid[3] = {1, 1, 1};
  double eps = 2.0;  // electrons_per_state
  DevBuf kweight, bands, grads;
  const double *grads_host = nullptr;  // OD_MEM_HOST_DEFERRED: uploaded slab-wise by the DOS drivers
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  DevBuf gslab[2];

  // scratch
  DevBuf flags;      // int[4]: [0] doslin `stop` conditions met by kernels that are not followed by a host synchronisation
  DevBuf stats_dev;  // unsigned long long[2]: lane-steps / in-window pairs of the narrow engine since the last profile reset
  DevBuf dest_dev;
  double *merge_status_dev = nullptr;  // status word of the last merged JDOS call (od_merge_status_check)
  void *host_stage = nullptr;  // pinned staging for the results of the task drivers
  size_t host_stage_bytes = 0;
  struct PendingTime { cudaEvent_t e0, e1; int kind; };  // kind 0: prepare, 1: accumulate
  std::vector<PendingTime> pending_times;
  std::vector<cudaEvent_t> event_pool;
  DevBuf partial, accum, small, recs, keys, scratch_a, scratch_b, scratch_c, scratch_d, mw_tmp, in_tmp, in_tmp2, moments, task_mw;

  Spectra dos_res, jdos_res;
  double vbm_energy = 0.0, cbm_energy = 0.0;  // od_dos_utils module variables set by compute_bandgap (dos_utils.f90:55-56)
  OpticsRes optics_res;
  CoreRes core_res;
};

int od_fail(od_ctx *ctx, int code, const char *fmt, ...);
int od_buf_ensure(od_ctx *ctx, DevBuf &b, size_t bytes);
void od_buf_release(DevBuf &b);
// make `src` (host or device) available on the device; returns device pointer (borrowed or staged in `stage`)
int od_stage_in(od_ctx *ctx, const void *src, size_t bytes, int mem, DevBuf &stage, const void **dptr);

#define OD_CUDA(ctx, call)                                                                          \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return od_fail((ctx), OD_ERR_CUDA, "%s:%d CUDA error %s: %s", __FILE__, __LINE__, #call,      \
                     cudaGetErrorString(e__));                                                      \
  } while (0)

#define OD_CHECK(expr)        \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != 0) return rc__; \
  } while (0)

#define OD_LAUNCH_CHECK(ctx)                                                                       \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return od_fail((ctx), OD_ERR_CUDA, "%s:%d kernel launch failed: %s", __FILE__, __LINE__,     \
                     cudaGetErrorString(e__));                                                     \
  } while (0)

// adaptive prefactor shared by the hot loops (dos_utils.f90:1210-1216)
struct GridGeom {
  double step[3];
  double adaptive_smearing_temp;
  // columns of recip_lattice scaled by step: corner(c) = sum_j sigma_j * rs[.][j]  (dos_utils.f90:1359-1364)
  double recip[9];
};
GridGeom od_grid_geom(const od_ctx *ctx, double adaptive_smearing);

// allreduce on a DEVICE buffer (no-op when nranks == 1)
int od_allreduce_device(od_ctx *ctx, double *dbuf, size_t n, char op);
// ctx->flags / ctx->stats_dev allocated and zeroed
int od_flags_ready(od_ctx *ctx);
// after a host synchronisation: report (and clear) what kernels flagged since the last check
int od_check_deferred(od_ctx *ctx);
int od_merge_status_check(od_ctx *ctx);
// device stopwatch without a synchronisation: bracket a region with od_time_begin / od_time_end; the elapsed times are
// added to ctx->prof when od_profile_get (or od_times_resolve) is called
int od_time_begin(od_ctx *ctx, int kind, size_t *slot);
int od_time_end(od_ctx *ctx, size_t slot);
int od_times_resolve(od_ctx *ctx);
// make deferred gradients resident (paths that do not stream)
int od_grads_resident(od_ctx *ctx);
