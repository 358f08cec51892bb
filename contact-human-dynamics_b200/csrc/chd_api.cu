// C ABI of the phys-optim path (see include/chd.h).  Host driver: device memory, stage loops, launches.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/chd.h"
#include "chd_dev.h"

// kernels (chd_kernels.cu)
__global__ void chd_k_stage_begin(ChdDev D);
__global__ void chd_k_eval(ChdDev D);
__global__ void chd_k_init(ChdDev D);
__global__ void chd_k_kkt(ChdDev D);
__global__ void chd_k_kkt_gwin(ChdDev D);
__global__ void chd_k_kcopy(ChdDev D);
__global__ void chd_k_curv(ChdDev D);
__global__ void chd_k_asm(ChdDev D);
__global__ void chd_k_fp64_peak(int mode, int iters, double* sink);
__global__ void chd_k_hess_base(ChdDev D);
__global__ void chd_k_hess_dur(ChdDev D);
__global__ void chd_k_hess_zero(ChdDev D);
__global__ void chd_k_hess_fin(ChdDev D, int mode);
__global__ void chd_k_tables(ChdDev D);
__global__ void chd_k_clear_dyn(ChdDev D);
__global__ void chd_k_linesearch(ChdDev D);
__global__ void chd_k_sample(ChdDev D, double* out, int* frames_out);
__global__ void chd_k_snapshot(ChdDev D, int* frames_out);
__global__ void chd_k_sched_reset(ChdDev D);

#define CHD_CUDA(x)                                                                          \
  do {                                                                                       \
    cudaError_t e_ = (x);                                                                    \
    if (e_ != cudaSuccess) {                                                                 \
      fprintf(stderr, "libchd: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -100 - (int)e_;                                                                 \
    }                                                                                        \
  } while (0)

enum { KT_EVAL = 0, KT_KKT = 1, KT_LS = 2, KT_INIT = 3, KT_SAMPLE = 4, KT_N = 8 };

struct chd_phys_batch {
  ChdHostBatch hb;
  ChdDev D;
  std::vector<void*> allocs;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_kkt = nullptr, ev_ls = nullptr, ev_copy = nullptr;
  int64_t launches = 0;
  int timing = 0;
  bool host_only = false;
  double kt_ms[KT_N] = {0};
  int64_t kt_n[KT_N] = {0};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int* d_frames = nullptr;
  double* d_samples = nullptr;
  size_t smem_eval = 0, smem_kkt = 0, smem_ls = 0;
  int kcopy_blocks = 1;
  ChdIpm* h_ipm = nullptr;  // host copy of the per-sequence solver state
  double* d_x0 = nullptr;
  int64_t h2d_bytes = 0;
  ChdStageDev* d_stages = nullptr;
  int sched_max_iter = 0;
  int slots = 1 << 30;          // sequences iterating at once (CTAs that can be resident); the rest of a large batch queues up
  bool sched_has_dur = false;   // the running schedule contains stage 3 (its cost Hessian is rebuilt every iteration)
  // pristine copies of the tables stage 3 rewrites (chd_phys_reset)
  double *poly_T0 = nullptr, *poly_tend0 = nullptr, *phase_tend0 = nullptr;
  int* ent_col0 = nullptr;
};

namespace {

template <class T>
int dev_upload(chd_phys_batch* b, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CHD_CUDA(cudaMallocAsync(&p, bytes, b->stream));      // stream-ordered pool: freed blocks are reused by the next batch
  b->allocs.push_back(p);
  if (!v.empty()) CHD_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, b->stream));
  b->h2d_bytes += (int64_t)(v.size() * sizeof(T));
  *out = (const T*)p;
  return 0;
}
template <class T>
int dev_alloc(chd_phys_batch* b, size_t count, T** out) {
  void* p = nullptr;
  CHD_CUDA(cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), b->stream));
  CHD_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(count, 1) * sizeof(T), b->stream));
  b->allocs.push_back(p);
  *out = (T*)p;
  return 0;
}

ChdStageDev stage_dev(const ChdStageCfg& c, int stage) {
  ChdStageDev s;
  s.set_mask = c.set_mask;
  s.max_iter = c.max_iter;
  s.snap_after = stage == CHD_STAGE_12 ? 0 : (stage == CHD_STAGE_22 ? 1 : ((stage == CHD_STAGE_4 || stage == CHD_STAGE_3) ? 2 : -1));
  s.opt_dur = stage == CHD_STAGE_3;
  for (int i = 0; i < 3; ++i) s.w_data[i] = c.w_data[i], s.w_vel[i] = c.w_vel[i], s.w_acc[i] = c.w_acc[i];
  s.w_dur = c.w_dur;
  return s;
}

struct Timer {
  chd_phys_batch* b;
  int id;
  Timer(chd_phys_batch* bb, int i) : b(bb), id(i) {
    if (b->timing) cudaEventRecord(b->ev0, b->stream);
  }
  ~Timer() {
    b->launches++;
    b->kt_n[id]++;
    if (b->timing) {
      cudaEventRecord(b->ev1, b->stream);
      cudaEventSynchronize(b->ev1);
      float ms = 0;
      cudaEventElapsedTime(&ms, b->ev0, b->ev1);
      b->kt_ms[id] += ms;
    }
  }
};

void launch_eval(chd_phys_batch* b) {
  Timer t(b, KT_EVAL);
  chd_k_eval<<<b->hb.B, CHD_THREADS, b->smem_eval, b->stream>>>(b->D);
}

// uploads the stage table (optionally with an iteration-cap override for one stage) and the schedule
int set_schedule(chd_phys_batch* b, const int* sched, int nsched, int override_stage, int override_max_iter) {
  ChdStageDev tab[6];
  for (int s = 0; s < 6; ++s) tab[s] = stage_dev(b->hb.stage[s], s);
  if (override_stage >= 0 && override_max_iter > 0) tab[override_stage].max_iter = override_max_iter;
  CHD_CUDA(cudaMemcpyAsync(b->d_stages, tab, sizeof(tab), cudaMemcpyHostToDevice, b->stream));
  b->D.stages = b->d_stages;
  b->D.nsched = nsched;
  for (int i = 0; i < nsched; ++i) b->D.sched[i] = sched[i];
  b->sched_max_iter = 0;
  b->sched_has_dur = false;
  for (int i = 0; i < nsched; ++i) b->sched_max_iter += tab[sched[i]].max_iter + 2, b->sched_has_dur |= sched[i] == CHD_STAGE_3;
  {
    int q[2] = {0, b->slots};
    CHD_CUDA(cudaMemcpyAsync(b->D.queue, q, sizeof(q), cudaMemcpyHostToDevice, b->stream));
  }
  chd_k_sched_reset<<<(b->hb.B + 127) / 128, 128, 0, b->stream>>>(b->D);
  b->launches++;
  return 0;
}

// runs the uploaded schedule to completion: every sequence walks through its stages at its own pace
int run_schedule(chd_phys_batch* b) {
  const int B = b->hb.B;
  const int check_every = 8;
  for (int it = 0; it <= b->sched_max_iter; ++it) {
    {
      Timer t(b, KT_INIT);
      chd_k_stage_begin<<<B, CHD_THREADS, 0, b->stream>>>(b->D);
    }
    launch_eval(b);
    {
      Timer t(b, KT_INIT);
      chd_k_init<<<B, CHD_THREADS, 0, b->stream>>>(b->D);
    }
    {
      Timer t(b, KT_INIT);
      chd_k_hess_zero<<<dim3(8, B), 256, 0, b->stream>>>(b->D);
      chd_k_hess_base<<<dim3(8, B), CHD_THREADS, 0, b->stream>>>(b->D);
      chd_k_hess_fin<<<B, 128, 0, b->stream>>>(b->D, 0);
      b->launches += 2;
    }
    if (it > 0) {
      CHD_CUDA(cudaStreamWaitEvent(b->stream, b->ev_copy, 0));   // Kwork refreshed by the side stream
      chd_k_asm<<<dim3(8, B), 256, 0, b->stream>>>(b->D);     // matrix entries of the sequences that continue in their stage
      b->launches++;
    }
    {
      Timer t(b, KT_KKT);
      if (b->D.win_smem) chd_k_kkt<<<B, CHD_KKT_THREADS, b->smem_kkt, b->stream>>>(b->D);
      else chd_k_kkt_gwin<<<B, CHD_KKT_THREADS, b->smem_kkt, b->stream>>>(b->D);
    }
    // Kwork <- Kbase for the next iteration, overlapped with the line search / evaluation kernels (sequences in stage 3
    // are skipped: their cost Hessian moves with the durations and is rebuilt into Kwork after the line search)
    CHD_CUDA(cudaEventRecord(b->ev_kkt, b->stream));
    CHD_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_kkt, 0));
    chd_k_kcopy<<<dim3(b->kcopy_blocks, B), 256, 0, b->copy_stream>>>(b->D);
    b->launches++;
    {
      Timer t(b, KT_LS);
      chd_k_linesearch<<<B, CHD_THREADS, b->smem_ls, b->stream>>>(b->D);
    }
    // distance-row curvature of the next iteration needs the accepted iterate: after the line search, on the side stream
    CHD_CUDA(cudaEventRecord(b->ev_ls, b->stream));
    CHD_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_ls, 0));
    if (b->sched_has_dur) {
      chd_k_hess_dur<<<dim3(8, B), CHD_THREADS, 0, b->copy_stream>>>(b->D);
      chd_k_hess_fin<<<B, 128, 0, b->copy_stream>>>(b->D, 1);
      b->launches += 2;
    }
    chd_k_curv<<<dim3(8, B), 256, 0, b->copy_stream>>>(b->D);
    b->launches++;
    CHD_CUDA(cudaEventRecord(b->ev_copy, b->copy_stream));
    {
      Timer t(b, KT_SAMPLE);
      chd_k_snapshot<<<B, 128, 0, b->stream>>>(b->D, b->d_frames);
    }
    if ((it % check_every) == check_every - 1) {
      CHD_CUDA(cudaMemcpyAsync(b->h_ipm, b->D.ipm, B * sizeof(ChdIpm), cudaMemcpyDeviceToHost, b->stream));
      CHD_CUDA(cudaStreamSynchronize(b->stream));
      bool any = false;
      for (int i = 0; i < B; ++i) any |= (b->h_ipm[i].phase != CHD_PH_FINISHED);
      if (!any) break;
    }
  }
  CHD_CUDA(cudaMemcpyAsync(b->h_ipm, b->D.ipm, B * sizeof(ChdIpm), cudaMemcpyDeviceToHost, b->stream));
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  CHD_CUDA(cudaStreamSynchronize(b->copy_stream));
  CHD_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" {

const char* chd_version(void) { return "libchd 0.1 (sm_100a)"; }

int chd_measure_fp64_peak(double* dfma_gflops, double* dmma_gflops) {
  int dev = 0, sms = 0;
  CHD_CUDA(cudaGetDevice(&dev));
  CHD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  CHD_CUDA(cudaMalloc((void**)&sink, sizeof(double) * 1024));
  cudaEvent_t e0, e1;
  CHD_CUDA(cudaEventCreate(&e0));
  CHD_CUDA(cudaEventCreate(&e1));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      CHD_CUDA(cudaEventRecord(e0));
      chd_k_fp64_peak<<<blocks, threads>>>(mode, iters, sink);
      CHD_CUDA(cudaEventRecord(e1));
      CHD_CUDA(cudaEventSynchronize(e1));
      float ms = 0;
      CHD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best) best = ms;
    }
    // mode 0: 8 independent FMA chains per thread; mode 1: 4 independent m8n8k4 accumulators per warp (512 flop each)
    const double flop = mode == 0 ? (double)blocks * threads * iters * 8 * 2 : (double)blocks * (threads / 32) * iters * 4 * 512;
    double* out = mode == 0 ? dfma_gflops : dmma_gflops;
    if (out) *out = flop / (best * 1e-3) / 1e9;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(sink);
  return 0;
}

static int batch_create_impl(const chd_phys_problem* problems, int32_t batch, const chd_phys_weights* weights, int32_t device,
                             chd_phys_batch* b);

int chd_phys_batch_create(const chd_phys_problem* problems, int32_t batch, const chd_phys_weights* weights, int32_t device,
                          chd_phys_batch** out) {
  if (!problems || batch <= 0 || !out) return -1;
  chd_phys_batch* b = new chd_phys_batch();
  std::memset(&b->D, 0, sizeof(b->D));
  const int rc = batch_create_impl(problems, batch, weights, device, b);
  if (rc) {
    chd_phys_batch_destroy(b);   // single cleanup path: streams, events, pooled allocations, host buffers
    return rc;
  }
  *out = b;
  return 0;
}

static int batch_create_impl(const chd_phys_problem* problems, int32_t batch, const chd_phys_weights* weights, int32_t device,
                             chd_phys_batch* b) {
  const bool host_only = device == -2;  // layout tables only, no CUDA call (CPU-side tests of the host logic)
  if (device >= 0) CHD_CUDA(cudaSetDevice(device));
  chd_phys_weights w = {0.4, 1.7, 0.3, 0.1, 0.1};  // phys_optim.cpp:27-31
  if (weights) w = *weights;
  int rc = chd_build_layout(problems, batch, w, b->hb);
  if (rc) return rc;
  ChdHostBatch& hb = b->hb;
  ChdDev& D = b->D;
  std::memset(&D, 0, sizeof(D));
  b->host_only = host_only;
  if (host_only) return 0;
  D.B = hb.B, D.S = hb.S, D.Pmax = hb.Pmax, D.n_max = hb.n_max, D.m_max = hb.m_max, D.slots_max = hb.slots_max;
  D.sets_max = hb.sets_max, D.tab_max = hb.tab_max, D.F_max = hb.F_max, D.Kd_max = hb.Kd_max, D.Kr_max = hb.Kr_max;
  D.Na_max = hb.Na_max, D.nb_max = hb.nb_max, D.w_max = hb.w_max, D.par_stride = hb.par_stride(), D.n_ee_max = hb.n_ee_max;
  D.fo_max = hb.fo_max, D.Ph_max = hb.Ph_max;
  CHD_CUDA(cudaStreamCreate(&b->stream));
  {
    // keep freed device memory in the default pool of this device (a batch object is created per solve by callers
    // that mirror the reference's one-process-per-clip flow)
    int dev_id = 0;
    cudaMemPool_t pool;
    CHD_CUDA(cudaGetDevice(&dev_id));
    CHD_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev_id));
    unsigned long long keep = ~0ull;
    CHD_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  CHD_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
  CHD_CUDA(cudaEventCreateWithFlags(&b->ev_kkt, cudaEventDisableTiming));
  CHD_CUDA(cudaEventCreateWithFlags(&b->ev_copy, cudaEventDisableTiming));
  CHD_CUDA(cudaEventCreateWithFlags(&b->ev_ls, cudaEventDisableTiming));
  CHD_CUDA(cudaEventCreate(&b->ev0));
  CHD_CUDA(cudaEventCreate(&b->ev1));
#define UP(field) if ((rc = dev_upload(b, hb.field, &D.field))) return rc;
#define UPM(field, T) if ((rc = dev_upload(b, hb.field, (const T**)&D.field))) return rc;
  UP(seq) UPM(poly_T, double) UPM(poly_tend, double) UP(node_const) UP(par) UP(t_dyn) UP(t_rom) UP(t_data) UP(row_lo) UP(row_hi) UP(node_var)
  UP(itab) UP(ent_ptr) UPM(ent_col, int) UP(ent_row) UP(col_ptr) UP(col_ent) UP(var_kkt) UP(row_kkt) UP(row_set) UP(sets) UPM(phase_tend, double)
  UP(dur0) UP(poly_ph)
  // pristine copies of what stage 3 rewrites + the line search's trial tables
  if ((rc = dev_upload(b, hb.poly_T, (const double**)&b->poly_T0))) return rc;
  if ((rc = dev_upload(b, hb.poly_tend, (const double**)&b->poly_tend0))) return rc;
  if ((rc = dev_upload(b, hb.phase_tend, (const double**)&b->phase_tend0))) return rc;
  if ((rc = dev_upload(b, hb.ent_col, (const int**)&b->ent_col0))) return rc;
  if ((rc = dev_upload(b, hb.poly_T, (const double**)&D.poly_Tt))) return rc;
  if ((rc = dev_upload(b, hb.poly_tend, (const double**)&D.poly_tendt))) return rc;
#undef UPM
#undef UP
  const size_t B = hb.B, nm = B * hb.n_max, mm = B * hb.m_max;
#define AL(field, cnt) if ((rc = dev_alloc(b, (cnt), &D.field))) return rc;
  AL(x, nm) AL(xt, nm) AL(jty, nm) AL(unobs, nm) AL(dx, nm) AL(grad, nm) AL(g, mm) AL(gt, mm) AL(Jv, B * hb.slots_max) AL(rflag, mm)
  AL(sc, mm) AL(dL, mm) AL(dU, mm) AL(s, mm) AL(y, mm) AL(zL, mm) AL(zU, mm) AL(ds, mm) AL(dy, mm) AL(dzL, mm) AL(dzU, mm)
  D.nbc_max = (hb.Na_max + 7) / 8;
  D.Q = (hb.w_max + 7) / 8 + 1;
  D.Qfix = (hb.w_fix_max + 7) / 8 + 1;
  D.tma = getenv("CHD_TMA") ? atoi(getenv("CHD_TMA")) : 0;   // band tiles the fixed-duration stages work with (storage strides follow Q)
  D.nbt = (hb.nb_max + 1 + 7) / 8;
  D.win_tiles = std::max(D.Q * (D.Q + 1) / 2, 2 * D.Q);
  D.kstride = (size_t)D.nbc_max * D.Q * 64 + (size_t)D.nbc_max * D.nbt * 64 + (size_t)64 * D.nbt * D.nbt;
  b->kcopy_blocks = (int)std::min<size_t>(512, std::max<size_t>(std::max<size_t>(1, 592 / B), D.kstride * sizeof(double) / 131072));
  AL(cost, B * 2) AL(Kwork, B * D.kstride) AL(Kbase, B * D.kstride) AL(sol, B * (size_t)(hb.Na_max + hb.nb_max)) AL(ipm, B)
  AL(rhs0, B * (size_t)(hb.Na_max + hb.nb_max)) AL(rhs1, B * (size_t)(hb.Na_max + hb.nb_max))
#undef AL
  if ((rc = dev_upload(b, hb.x0, (const double**)&b->d_x0))) return rc;
  CHD_CUDA(cudaMemcpyAsync(D.x, b->d_x0, nm * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  // shared-memory budgets
  const size_t nbp8 = 8 * (size_t)D.nbt;
  b->smem_eval = (2 * (size_t)hb.n_max + CHD_THREADS) * sizeof(double);
  b->smem_ls = ((size_t)hb.n_max + CHD_THREADS) * sizeof(double);
  D.pan_doubles = std::max(2 * (D.Q + D.nbt) * 64, 8 * D.nbc_max);
  cudaFuncAttributes fa;
  CHD_CUDA(cudaFuncGetAttributes(&fa, chd_k_kkt));
  const size_t kkt_static = fa.sharedSizeBytes;
  const size_t n_even = (size_t)((hb.n_max + 1) & ~1), xs_len = 8 * (size_t)D.nbc_max + nbp8;
  // the per-unknown vectors alias the tail of the window region (chd_kkt_body)
  const size_t kkt_fixed = (CHD_KKT_THREADS + nbp8 * nbp8 + (size_t)D.pan_doubles + 16) * sizeof(double);
  const size_t kkt_win = ((size_t)D.win_tiles * 64 + (size_t)D.Q * D.nbt * 64) * sizeof(double);
  int dev = 0, smem_max = 0;
  CHD_CUDA(cudaGetDevice(&dev));
  CHD_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const bool vectors_fit = (size_t)D.win_tiles * 64 + (size_t)D.Q * D.nbt * 64 >= n_even + xs_len + 2 * (size_t)D.Q * 64 + 3072;
  if (vectors_fit && kkt_fixed + kkt_win + kkt_static + 256 <= (size_t)smem_max) {
    D.win_smem = 1;
    b->smem_kkt = kkt_fixed + kkt_win;
  } else {
    // long horizons / wide bands: vectors and window in a global scratch area (chd_k_kkt_gwin)
    D.win_smem = 0;
    D.pan_doubles = 2 * (D.Q + D.nbt) * 64;
    b->smem_kkt = (CHD_KKT_THREADS + nbp8 * nbp8 + (size_t)D.pan_doubles + 16) * sizeof(double);
    D.scratch_stride = n_even + xs_len + 8 * (size_t)D.nbc_max + (size_t)D.win_tiles * 64 + (size_t)D.Q * D.nbt * 64;
    if ((rc = dev_alloc(b, B * D.scratch_stride, &D.scratch))) return rc;
    if (D.Q - 1 + D.nbt > 76) {
      fprintf(stderr, "libchd: band + border too wide for the pair tables (Q=%d nbt=%d)\n", D.Q, D.nbt);
      return -5;
    }
  }
  {
    int sms = 0;
    CHD_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // Admission queue (continuous batching): CHD_SLOTS=n lets at most n sequences iterate at once, the others wait for a
    // finished one to hand over (chd_stage_advance).  Off by default: measured on a 1024-sequence job (148 slots = one
    // CTA per SM) 35.7 k frames/s against 39.4 k without the queue -- finished sequences already cost nothing but an
    // early-exit CTA, the cost of a KKT launch grows with the number of live sequences either way (16 us per live
    // sequence once their band storage exceeds the L2), and late admission only delays the slow sequences.
    (void)sms;
    b->slots = getenv("CHD_SLOTS") ? atoi(getenv("CHD_SLOTS")) : (1 << 30);
    if (b->slots < 1) b->slots = 1;
    if ((rc = dev_alloc(b, 2, &D.queue))) return rc;
  }
  if (b->smem_eval + 1024 > (size_t)smem_max || b->smem_kkt + kkt_static + 256 > (size_t)smem_max) {
    fprintf(stderr, "libchd: problem too large for the shared-memory staged kernels (n_max=%d)\n", hb.n_max);
    return -5;
  }
  CHD_CUDA(cudaFuncSetAttribute(chd_k_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_eval));
  CHD_CUDA(cudaFuncSetAttribute(chd_k_kkt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_kkt));
  CHD_CUDA(cudaFuncSetAttribute(chd_k_kkt_gwin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_kkt));
  CHD_CUDA(cudaFuncSetAttribute(chd_k_linesearch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem_ls));
  const size_t stride = 6 + 7 * (size_t)hb.n_ee_max;
  CHD_CUDA(cudaMallocAsync((void**)&b->d_samples, B * hb.fo_max * stride * sizeof(double), b->stream));
  CHD_CUDA(cudaMallocAsync((void**)&b->d_frames, B * sizeof(int), b->stream));
  b->allocs.push_back(b->d_samples);
  b->allocs.push_back(b->d_frames);
  b->h_ipm = (ChdIpm*)std::malloc(B * sizeof(ChdIpm));   // pageable: pinned allocation / release cost up to 0.3 s per batch
  if (!b->h_ipm) return -3;
  CHD_CUDA(cudaMallocAsync((void**)&b->d_stages, 6 * sizeof(ChdStageDev), b->stream));
  b->allocs.push_back(b->d_stages);
  if ((rc = dev_alloc(b, 3 * B * hb.fo_max * stride, &D.snapshots))) return rc;
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

void chd_phys_batch_destroy(chd_phys_batch* b) {
  if (!b) return;
  if (b->stream) {
    for (void* p : b->allocs) cudaFreeAsync(p, b->stream);   // back to the pool, not to the driver (cudaFree cost up to 350 ms per batch)
    cudaStreamSynchronize(b->stream);
  }
  std::free(b->h_ipm);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev_kkt) cudaEventDestroy(b->ev_kkt);
  if (b->ev_copy) cudaEventDestroy(b->ev_copy);
  if (b->ev_ls) cudaEventDestroy(b->ev_ls);
  if (b->copy_stream) cudaStreamDestroy(b->copy_stream);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int chd_phys_get_dims(const chd_phys_batch* b, chd_phys_dims* d) {
  if (!b || !d) return -1;
  const ChdHostBatch& hb = b->hb;
  d->batch = hb.B, d->n_max = hb.n_max, d->m_max = hb.m_max, d->slots_max = hb.slots_max, d->n_splines = hb.S;
  d->p_max = hb.Pmax, d->sets_max = hb.sets_max, d->na_max = hb.Na_max, d->nb_max = hb.nb_max, d->w_max = hb.w_max;
  d->frames_out_max = hb.fo_max;
  return 0;
}
int chd_phys_get_sizes(const chd_phys_batch* b, int32_t* s) {
  if (!b || !s) return -1;
  for (int i = 0; i < b->hb.B; ++i) {
    const ChdSeq& h = b->hb.seq[i];
    s[6 * i + 0] = h.n, s[6 * i + 1] = h.m, s[6 * i + 2] = h.nslots, s[6 * i + 3] = h.Na, s[6 * i + 4] = h.nb, s[6 * i + 5] = h.w;
  }
  return 0;
}
int chd_phys_get_sizes_fixed(const chd_phys_batch* b, int32_t* s) {
  if (!b || !s) return -1;
  for (int i = 0; i < b->hb.B; ++i) s[3 * i] = b->hb.seq[i].nb_fix, s[3 * i + 1] = b->hb.seq[i].w_fix, s[3 * i + 2] = b->hb.seq[i].n_dur;
  return 0;
}
int chd_phys_get_x(const chd_phys_batch* b, double* x) {
  if (!b || !x) return -1;
  if (b->host_only) {
    std::memcpy(x, b->hb.x0.data(), b->hb.x0.size() * sizeof(double));
    return 0;
  }
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  CHD_CUDA(cudaMemcpy(x, b->D.x, (size_t)b->hb.B * b->hb.n_max * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}
int chd_phys_set_x(chd_phys_batch* b, const double* x) {
  if (!b || !x || b->host_only) return -1;
  CHD_CUDA(cudaMemcpy(b->D.x, x, (size_t)b->hb.B * b->hb.n_max * sizeof(double), cudaMemcpyHostToDevice));
  chd_k_tables<<<b->hb.B, 64, 0, b->stream>>>(b->D);   // spline tables follow the durations held in x
  b->launches++;
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

int chd_phys_eval(chd_phys_batch* b, int32_t stage, double* cost, double* grad, double* g, double* jac_vals) {
  if (!b || stage < 0 || stage > 5 || b->host_only) return -1;
  const ChdHostBatch& hb = b->hb;
  int sched[1] = {stage};
  const int slots_keep = b->slots;
  b->slots = 1 << 30;                       // a function evaluation touches every sequence of the batch at once (no queue)
  int rc = set_schedule(b, sched, 1, -1, 0);
  b->slots = slots_keep;
  if (rc) return rc;
  {
    Timer t(b, KT_INIT);
    chd_k_stage_begin<<<hb.B, CHD_THREADS, 0, b->stream>>>(b->D);
  }
  launch_eval(b);
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  CHD_CUDA(cudaGetLastError());
  const size_t B = hb.B;
  if (cost) {
    std::vector<double> c2(2 * B);
    CHD_CUDA(cudaMemcpy(c2.data(), b->D.cost, 2 * B * sizeof(double), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < B; ++i) cost[i] = c2[2 * i];
  }
  if (grad) CHD_CUDA(cudaMemcpy(grad, b->D.grad, B * hb.n_max * sizeof(double), cudaMemcpyDeviceToHost));
  if (g) CHD_CUDA(cudaMemcpy(g, b->D.g, B * hb.m_max * sizeof(double), cudaMemcpyDeviceToHost));
  if (jac_vals) CHD_CUDA(cudaMemcpy(jac_vals, b->D.Jv, B * hb.slots_max * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

int chd_phys_get_layout(const chd_phys_batch* b, int32_t* ent_ptr, int32_t* ent_col, double* row_lo, double* row_hi,
                        int32_t* row_set, int32_t* var_kkt, int32_t* row_kkt) {
  if (!b) return -1;
  const ChdHostBatch& hb = b->hb;
  if (ent_ptr) std::memcpy(ent_ptr, hb.ent_ptr.data(), hb.ent_ptr.size() * sizeof(int));
  if (ent_col) std::memcpy(ent_col, hb.ent_col.data(), hb.ent_col.size() * sizeof(int));
  if (row_lo) std::memcpy(row_lo, hb.row_lo.data(), hb.row_lo.size() * sizeof(double));
  if (row_hi) std::memcpy(row_hi, hb.row_hi.data(), hb.row_hi.size() * sizeof(double));
  if (row_set) std::memcpy(row_set, hb.row_set.data(), hb.row_set.size() * sizeof(int));
  if (var_kkt) std::memcpy(var_kkt, hb.var_kkt.data(), hb.var_kkt.size() * sizeof(int));
  if (row_kkt) std::memcpy(row_kkt, hb.row_kkt.data(), hb.row_kkt.size() * sizeof(int));
  return 0;
}

int chd_phys_get_slot_index(const chd_phys_batch* b, int32_t* ent_row, int32_t* col_ptr, int32_t* col_ent) {
  if (!b) return -1;
  const ChdHostBatch& hb = b->hb;
  if (ent_row) std::memcpy(ent_row, hb.ent_row.data(), hb.ent_row.size() * sizeof(int));
  if (col_ptr) std::memcpy(col_ptr, hb.col_ptr.data(), hb.col_ptr.size() * sizeof(int));
  if (col_ent) std::memcpy(col_ent, hb.col_ent.data(), hb.col_ent.size() * sizeof(int));
  return 0;
}

int chd_phys_get_ent_col(const chd_phys_batch* b, int32_t* ent_col) {
  if (!b || !ent_col) return -1;
  if (b->host_only) {
    std::memcpy(ent_col, b->hb.ent_col.data(), b->hb.ent_col.size() * sizeof(int));
    return 0;
  }
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  CHD_CUDA(cudaMemcpy(ent_col, b->D.ent_col, b->hb.ent_col.size() * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

int chd_phys_get_duals(const chd_phys_batch* b, double* y, double* zL, double* zU, double* s, double* row_scale, double* obj_scale) {
  if (!b || b->host_only) return -1;
  const size_t B = b->hb.B, mm = B * b->hb.m_max;
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  if (y) CHD_CUDA(cudaMemcpy(y, b->D.y, mm * sizeof(double), cudaMemcpyDeviceToHost));
  if (zL) CHD_CUDA(cudaMemcpy(zL, b->D.zL, mm * sizeof(double), cudaMemcpyDeviceToHost));
  if (zU) CHD_CUDA(cudaMemcpy(zU, b->D.zU, mm * sizeof(double), cudaMemcpyDeviceToHost));
  if (s) CHD_CUDA(cudaMemcpy(s, b->D.s, mm * sizeof(double), cudaMemcpyDeviceToHost));
  if (row_scale) CHD_CUDA(cudaMemcpy(row_scale, b->D.sc, mm * sizeof(double), cudaMemcpyDeviceToHost));
  if (obj_scale) {
    std::vector<ChdIpm> ipm(B);
    CHD_CUDA(cudaMemcpy(ipm.data(), b->D.ipm, B * sizeof(ChdIpm), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < B; ++i) obj_scale[i] = ipm[i].sf;
  }
  return 0;
}

/* diagnostic: line-search inputs of the last KKT solve of sequence i (a_pr, a_du, dphi, phi0, theta0, theta_ref, mu, delta_w) */
int chd_phys_debug_ipm(const chd_phys_batch* b, int32_t i, double* out8) {
  if (!b || !out8 || b->host_only || !b->h_ipm || i < 0 || i >= b->hb.B) return -1;
  const ChdIpm& I = b->h_ipm[i];
  out8[0] = I.a_pr, out8[1] = I.a_du, out8[2] = I.dphi, out8[3] = I.phi0, out8[4] = I.theta0, out8[5] = I.theta_ref, out8[6] = I.mu, out8[7] = I.delta_w;
  for (int q = 0; q < 8; ++q) out8[8 + q] = I.dbg[q];
  return 0;
}

int chd_phys_stage_stats(const chd_phys_batch* b, double* stats) {
  if (!b || !stats || b->host_only || !b->h_ipm) return -1;
  const int B = b->hb.B;
  for (int s = 0; s < 6; ++s)
    for (int i = 0; i < B; ++i)
      for (int q = 0; q < 4; ++q) stats[((size_t)s * B + i) * 4 + q] = b->h_ipm[i].st_stat[s][q];
  return 0;
}

int chd_phys_solve_stage(chd_phys_batch* b, int32_t stage, int32_t max_iter, int32_t* status, int32_t* iters, double* stats) {
  if (!b || stage < 0 || stage > 5 || b->host_only) return -1;
  const int B = b->hb.B;
  int sched[1] = {stage};
  int rc = set_schedule(b, sched, 1, stage, max_iter);
  if (rc) return rc;
  if ((rc = run_schedule(b))) return rc;
  for (int i = 0; i < B; ++i) {
    const ChdIpm& I = b->h_ipm[i];
    if (status) status[i] = I.st_status[stage];
    if (iters) iters[i] = I.st_iters[stage];
    if (stats) {
      double* s = stats + 8 * i;
      s[0] = I.f, s[1] = I.E0, s[2] = I.viol_u, s[3] = I.dual_u, s[4] = I.compl_u, s[5] = I.mu, s[6] = I.delta_w, s[7] = I.ls_fail;
    }
    if (i == 0 && getenv("CHD_PROF")) fprintf(stderr, "chd prof factor loop (Mcycles, accumulated since batch creation) warp1: panel+barrier %.2f updates %.2f wait+barrier %.2f | warp0: panel+barrier %.2f diagonal %.2f wait+barrier %.2f\n", I.dbg[0]/1e6, I.dbg[1]/1e6, I.dbg[2]/1e6, I.dbg[3]/1e6, I.dbg[4]/1e6, I.dbg[5]/1e6);
    if (i == 0 && getenv("CHD_PROF")) fprintf(stderr, "chd prof (Mcycles) seq0 stage %d: err %.2f jasm %.2f hasm %.2f factor %.2f border %.2f back %.2f rec %.2f\n", stage, I.prof[0]/1e6, I.prof[1]/1e6, I.prof[2]/1e6, I.prof[3]/1e6, I.prof[4]/1e6, I.prof[5]/1e6, I.prof[6]/1e6);
  }
  return 0;
}

int chd_phys_sample_device(chd_phys_batch* b, double* out_device, void* stream) {
  if (!b || !out_device || b->host_only) return -1;
  cudaStream_t st = stream ? (cudaStream_t)stream : b->stream;
  if (st == b->stream) {
    Timer t(b, KT_SAMPLE);
    chd_k_sample<<<b->hb.B, 128, 0, st>>>(b->D, out_device, b->d_frames);
    return 0;
  }
  // caller's stream: order the launch behind everything pending on the batch's own stream (reset / upload copies, the
  // solve) and make the batch's stream wait for it in turn (it writes d_frames and reads x)
  cudaEvent_t ev;
  CHD_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  CHD_CUDA(cudaEventRecord(ev, b->stream));
  CHD_CUDA(cudaStreamWaitEvent(st, ev, 0));
  chd_k_sample<<<b->hb.B, 128, 0, st>>>(b->D, out_device, b->d_frames);
  CHD_CUDA(cudaEventRecord(ev, st));
  CHD_CUDA(cudaStreamWaitEvent(b->stream, ev, 0));
  CHD_CUDA(cudaEventDestroy(ev));
  return 0;
}

int chd_phys_sample(chd_phys_batch* b, double* out, int32_t* frames_out) {
  if (!b || !out || b->host_only) return -1;
  const ChdHostBatch& hb = b->hb;
  const size_t stride = 6 + 7 * (size_t)hb.n_ee_max, cnt = (size_t)hb.B * hb.fo_max * stride;
  CHD_CUDA(cudaMemsetAsync(b->d_samples, 0, cnt * sizeof(double), b->stream));
  int rc = chd_phys_sample_device(b, b->d_samples, nullptr);
  if (rc) return rc;
  CHD_CUDA(cudaMemcpyAsync(out, b->d_samples, cnt * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (frames_out) CHD_CUDA(cudaMemcpyAsync(frames_out, b->d_frames, hb.B * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

// Staged schedule of phys_optim.cpp:554-749.  Every sequence walks through 1.1, 1.2, 2.1, 2.2, 3 and -- only when its
// stage 3 did not succeed (:713-749) -- 4 at its own pace (converged sequences do not wait for the slowest one of
// their stage).  Stage status -9 = stage not run (stage 4 after a successful stage 3), -3 = stage 3 not attempted
// (more phase durations than CHD_MAX_DUR).
int chd_phys_solve(chd_phys_batch* b, double* samples, int32_t* frames_out, int32_t* success, int32_t* stage_status,
                   int32_t* stage_iters) {
  if (!b || b->host_only) return -1;
  const ChdHostBatch& hb = b->hb;
  const int B = hb.B;
  const size_t stride = 6 + 7 * (size_t)hb.n_ee_max, snap = (size_t)B * hb.fo_max * stride;
  int sched[6] = {CHD_STAGE_11, CHD_STAGE_12, CHD_STAGE_21, CHD_STAGE_22, CHD_STAGE_3, CHD_STAGE_4};
  int rc = set_schedule(b, sched, 6, -1, 0);
  if (rc) return rc;
  if (samples) CHD_CUDA(cudaMemsetAsync(b->D.snapshots, 0, 3 * snap * sizeof(double), b->stream));
  if ((rc = run_schedule(b))) return rc;
  for (int i = 0; i < B; ++i) {
    const ChdIpm& I = b->h_ipm[i];
    // dynamics_succeed (:655); durations_succeed = stage 3 (:709), overwritten by stage 4 when that had to run (:746)
    if (success) success[2 * i] = I.st_status[CHD_STAGE_22] == 0, success[2 * i + 1] = I.st_status[CHD_STAGE_3] == 0 || I.st_status[CHD_STAGE_4] == 0;
    for (int s = 0; s < 6; ++s) {
      if (stage_status) stage_status[(size_t)s * B + i] = I.st_status[s];
      if (stage_iters) stage_iters[(size_t)s * B + i] = I.st_iters[s];
    }
  }
  if (samples) CHD_CUDA(cudaMemcpyAsync(samples, b->D.snapshots, 3 * snap * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (frames_out) CHD_CUDA(cudaMemcpyAsync(frames_out, b->d_frames, B * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  CHD_CUDA(cudaStreamSynchronize(b->stream));
  return 0;
}

int64_t chd_phys_launch_count(const chd_phys_batch* b) { return b ? b->launches : 0; }
int64_t chd_phys_h2d_bytes(const chd_phys_batch* b) { return b ? b->h2d_bytes : 0; }
int chd_phys_reset(chd_phys_batch* b) {
  if (!b || b->host_only) return -1;
  CHD_CUDA(cudaMemcpyAsync(b->D.x, b->d_x0, (size_t)b->hb.B * b->hb.n_max * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  // the input durations are back: host-built spline tables and Jacobian columns
  const ChdHostBatch& hb = b->hb;
  CHD_CUDA(cudaMemcpyAsync(b->D.poly_T, b->poly_T0, hb.poly_T.size() * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  CHD_CUDA(cudaMemcpyAsync(b->D.poly_tend, b->poly_tend0, hb.poly_tend.size() * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  CHD_CUDA(cudaMemcpyAsync(b->D.phase_tend, b->phase_tend0, hb.phase_tend.size() * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  CHD_CUDA(cudaMemcpyAsync(b->D.ent_col, b->ent_col0, hb.ent_col.size() * sizeof(int), cudaMemcpyDeviceToDevice, b->stream));
  chd_k_clear_dyn<<<(hb.B + 127) / 128, 128, 0, b->stream>>>(b->D);
  b->launches++;
  return 0;
}
int chd_phys_set_timing(chd_phys_batch* b, int enable) {
  if (!b) return -1;
  b->timing = enable;
  return 0;
}
int chd_phys_kernel_times(chd_phys_batch* b, double* ms8, int64_t* launches8, int reset) {
  if (!b) return -1;
  for (int i = 0; i < KT_N; ++i) {
    if (ms8) ms8[i] = b->kt_ms[i];
    if (launches8) launches8[i] = b->kt_n[i];
    if (reset) b->kt_ms[i] = 0, b->kt_n[i] = 0;
  }
  return 0;
}

}  // extern "C"
