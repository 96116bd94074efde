// Reverse-process sampler: the loop of Diffusion.sample (src/DDPM.py:98-130) as one CUDA graph per
// timestep, replayed T times with a device-resident step counter.
//
// Per step (all on the caller's stream, no host synchronisation, no H2D of `t`, no .item()):
//   UNet forward on 2B rows (rows [0,B) conditional, [B,2B) unconditional; both read the same B-row x_t)
//   -> fused CFG lerp + p_sample update in place on x_t (in-kernel Philox noise keyed by the global sample
//      index, or injected noise for fixed-noise parity)
//   -> t -= 1
// The reference does, per step: 1 H2D (t), 2 UNet passes of ~270 kernels, ~12 element-wise kernels and a
// D2H sync (src/DDPM.py:85,116).
#include <new>

#include "../../include/ldm_b200.h"
#include "kernels.h"
#include "unet_internal.h"

struct ldm_sampler {
  ldm_unet* unet;
  ldm_sampler_desc d;
  int ub;             // rows per UNet pass (2B with guidance)
  int64_t n;          // elements per sample
  int64_t off_eps, off_t, off_unet, total;
  cudaGraphExec_t exec = nullptr;
  // The legacy default stream cannot be captured: when the caller hands us stream 0 the work is forked onto
  // this private stream (event-ordered after the caller's stream, and joined back before returning).
  cudaStream_t own = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  long long step_kernels = 0;  // kernels inside one captured step (for ldm_launch_count under graph replay)
  // capture key: everything baked into the graph's kernel arguments.  The Philox seed and the global sample offset are
  // NOT part of it: they live in device memory next to the step counter, so one captured step serves every call
  // (generate_images.py makes ten batch-1 calls with fresh seeds: one capture instead of ten)
  struct Key {
    const void *x, *y, *coef, *noise, *ws;
    bool operator==(const Key& o) const {
      return x == o.x && y == o.y && coef == o.coef && noise == o.noise && ws == o.ws;
    }
  } key{};
};

extern "C" int ldm_sampler_create(ldm_unet* unet, const ldm_sampler_desc* desc, ldm_sampler** out) {
  LDM_REQUIRE(unet && desc && out, "ldm_sampler_create: null argument");
  LDM_REQUIRE(desc->batch > 0 && desc->n_steps > 0, "sampler: batch and n_steps must be positive");
  LDM_REQUIRE(desc->y_len == 0 || desc->y_len == 1 || desc->y_len == desc->batch,
              "sampler: classes must have length 1 or batch (got %d for batch %d)", desc->y_len, desc->batch);
  LDM_REQUIRE(ldm_unet_in_channels(unet) == ldm_unet_out_channels(unet), "sampler: eps-model must map C -> C channels");
  ldm_sampler* s = new (std::nothrow) ldm_sampler();
  LDM_REQUIRE(s, "out of host memory");
  s->unet = unet;
  s->d = *desc;
  s->ub = desc->cfg_scale > 0.f ? 2 * desc->batch : desc->batch;
  const int S = ldm_unet_image_size(unet);
  s->n = (int64_t)ldm_unet_in_channels(unet) * S * S;
  int64_t off = 0;
  s->off_eps = off; off += align_up64((int64_t)s->ub * s->n * 4, 1024);
  s->off_t = off; off += 1024;
  s->off_unet = off; off += ldm_unet_workspace_bytes(unet, s->ub);
  s->total = off;
  if (cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_in, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_out, cudaEventDisableTiming) != cudaSuccess) {
    ldm_sampler_destroy(s);
    return ldm_set_error("sampler: could not create its stream/events: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = s;
  return 0;
}
extern "C" void ldm_sampler_destroy(ldm_sampler* s) {
  if (!s) return;
  if (s->exec) cudaGraphExecDestroy(s->exec);
  if (s->ev_in) cudaEventDestroy(s->ev_in);
  if (s->ev_out) cudaEventDestroy(s->ev_out);
  if (s->own) cudaStreamDestroy(s->own);
  delete s;
}
extern "C" int64_t ldm_sampler_workspace_bytes(const ldm_sampler* s) { return s ? s->total : -1; }

static int sampler_step(ldm_sampler* s, float* x, const int64_t* y, const float* coef, const float* noise,
                        uint64_t seed, uint64_t sample_offset, uint8_t* ws, cudaStream_t st) {
  const int B = s->d.batch;
  float* eps = (float*)(ws + s->off_eps);
  int64_t* tdev = (int64_t*)(ws + s->off_t);
  const bool cfg = s->d.cfg_scale > 0.f;
  int rc = ldm_unet_forward_ex(s->unet, x, B, nullptr, tdev, s->d.y_len > 0 ? y : nullptr, s->d.y_len, B, s->ub, eps,
                               ws + s->off_unet, s->total - s->off_unet, st);
  if (rc) return rc;
  rc = k_p_sample(x, eps, cfg ? eps + (int64_t)B * s->n : nullptr, s->d.cfg_scale, tdev, 0, coef, s->d.n_steps, noise,
                  noise ? (int64_t)B * s->n : 0, seed, sample_offset, x, B, s->n, st, (const uint64_t*)(tdev + 1));
  if (rc) return rc;
  return k_add_i64(tdev, -1, st);
}

static int sampler_run_on(ldm_sampler* s, float* x, int x_is_init, const int64_t* y, const float* coef,
                          const float* noise, uint64_t seed, uint64_t sample_offset, int first_step, int num_steps,
                          void* workspace, int64_t workspace_bytes, cudaStream_t stream);

extern "C" int ldm_sampler_run(ldm_sampler* s, float* x, int x_is_init, const int64_t* y, const float* coef,
                               const float* noise, uint64_t seed, uint64_t sample_offset, int first_step,
                               int num_steps, void* workspace, int64_t workspace_bytes, void* stream) {
  LDM_REQUIRE(s, "ldm_sampler_run: null sampler");
  cudaStream_t user = (cudaStream_t)stream;
  const bool fork = s->d.use_graph && (user == nullptr || user == cudaStreamLegacy || user == cudaStreamPerThread);
  if (!fork)
    return sampler_run_on(s, x, x_is_init, y, coef, noise, seed, sample_offset, first_step, num_steps, workspace,
                          workspace_bytes, user);
  LDM_CUDA(cudaEventRecord(s->ev_in, user));
  LDM_CUDA(cudaStreamWaitEvent(s->own, s->ev_in, 0));
  int rc = sampler_run_on(s, x, x_is_init, y, coef, noise, seed, sample_offset, first_step, num_steps, workspace,
                          workspace_bytes, s->own);
  LDM_CUDA(cudaEventRecord(s->ev_out, s->own));
  LDM_CUDA(cudaStreamWaitEvent(user, s->ev_out, 0));
  return rc;
}

static int sampler_run_on(ldm_sampler* s, float* x, int x_is_init, const int64_t* y, const float* coef,
                          const float* noise, uint64_t seed, uint64_t sample_offset, int first_step, int num_steps,
                          void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  LDM_REQUIRE(s && x && coef, "ldm_sampler_run: null argument");
  LDM_REQUIRE(workspace && workspace_bytes >= s->total, "sampler workspace too small: %lld < %lld bytes",
              (long long)workspace_bytes, (long long)s->total);
  LDM_REQUIRE(((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");
  LDM_REQUIRE(first_step >= 0 && first_step < s->d.n_steps && num_steps >= 0 && num_steps <= first_step + 1,
              "sampler: steps [%d down %d) outside the schedule of %d", first_step, num_steps, s->d.n_steps);
  LDM_REQUIRE(s->d.y_len == 0 || y != nullptr, "sampler: classes pointer required");
  cudaStream_t st = stream;
  uint8_t* ws = (uint8_t*)workspace;
  if (!x_is_init) {
    int rc = k_randn(x, s->d.batch, s->n, seed, sample_offset, 0x17u, st);
    if (rc) return rc;
  }
  if (num_steps == 0) return 0;
  int rc = k_set_i64((int64_t*)(ws + s->off_t), first_step, st);
  if (rc) return rc;
  if ((rc = k_set_i64((int64_t*)(ws + s->off_t) + 1, (int64_t)seed, st))) return rc;
  if ((rc = k_set_i64((int64_t*)(ws + s->off_t) + 2, (int64_t)sample_offset, st))) return rc;
  if (!s->d.use_graph) {
    for (int i = 0; i < num_steps; ++i) {
      rc = sampler_step(s, x, y, coef, noise, seed, sample_offset, ws, st);
      if (rc) return rc;
    }
    return 0;
  }
  ldm_sampler::Key key{x, y, coef, noise, workspace};
  int done = 0;
  if (!s->exec || !(key == s->key)) {
    if (s->exec) { cudaGraphExecDestroy(s->exec); s->exec = nullptr; }
    // The first step runs eagerly: it loads every kernel (lazy module loading is not capturable)
    // and does real work, so nothing is wasted.
    rc = sampler_step(s, x, y, coef, noise, seed, sample_offset, ws, st);
    if (rc) return rc;
    done = 1;
    if (num_steps > 1) {
      cudaGraph_t graph = nullptr;
      LDM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
      const long long before = g_ldm_launches.load();
      rc = sampler_step(s, x, y, coef, noise, seed, sample_offset, ws, st);
      cudaError_t e = cudaStreamEndCapture(st, &graph);
      s->step_kernels = g_ldm_launches.load() - before;
      g_ldm_launches.fetch_sub(s->step_kernels);  // captured, not executed
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      LDM_REQUIRE(e == cudaSuccess && graph, "stream capture of the sampling step failed: %s", cudaGetErrorString(e));
      e = cudaGraphInstantiate(&s->exec, graph, 0);
      cudaGraphDestroy(graph);
      LDM_REQUIRE(e == cudaSuccess, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
      s->key = key;
    }
  }
  for (int i = done; i < num_steps; ++i) {
    LDM_CUDA(cudaGraphLaunch(s->exec, st));
    g_ldm_launches.fetch_add(s->step_kernels, std::memory_order_relaxed);
  }
  return 0;
}
