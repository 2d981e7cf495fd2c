// Vectorised CPU implementation of the synthetic "LinQuad" MDP for the host-env (e2e) leg: N independent
// env streams stepped by a small pool of spinning worker threads.  This plays the role of the user's POMDPs.jl model
// (@gen(:sp,:r), isterminal, initialstate; src/sampler.jl:39-50,89-97); it is NOT part of the reference.
// The noise streams are the same counter-based Philox4x32-10 / Box-Muller streams as the device env (csrc/env.cu), keyed by
// (seed, tick, env), so host and device rollouts can be cross-checked (libm vs CUDA math: last-ulp differences only).
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#define CPU_PAUSE() _mm_pause()
#else
#define CPU_PAUSE() std::this_thread::yield()
#endif

namespace {

struct Philox4 { uint32_t x, y, z, w; };
inline Philox4 philox4x32_10(uint64_t seed, uint64_t ctr_hi, uint64_t ctr_lo) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
inline float u32_to_unit_open(uint32_t x) { return ((float)(x >> 8) + 1.0f) * (1.0f / 16777216.0f); }
inline void box_muller(uint32_t a, uint32_t b, float &n0, float &n1) {
  const float u1 = u32_to_unit_open(a), u2 = u32_to_unit_open(b);
  const float rad = sqrtf(-2.0f * logf(u1));
  const float ang = 6.283185307179586f * u2;
  n0 = rad * cosf(ang); n1 = rad * sinf(ang);
}

struct Env {
  int n = 0, sdim = 0, adim = 0, n_threads = 1;
  uint64_t seed = 0, tick = 0;
  std::vector<float> A, B, state;
  // worker pool
  std::vector<std::thread> workers;
  std::atomic<uint64_t> gen{0};
  std::atomic<int> pending{0};
  std::atomic<bool> quit{false};
  std::mutex mu;
  std::condition_variable cv;
  // current job
  int job = 0;  // 1 = step
  const float *a = nullptr;
  float *sp = nullptr, *r = nullptr;
  uint8_t *done = nullptr;
  uint64_t job_tick = 0;
  int job_e0 = 0, job_e1 = 0;
  uint64_t last_step_tick = 0;
};

void s0_row(const Env &E, uint64_t tick, int64_t e, float *out) {
  for (int k = 0; k < E.sdim; k += 4) {
    const Philox4 p = philox4x32_10(E.seed ^ 0x5851F42D4C957F2DULL, tick, (uint64_t)e * 16 + (k >> 2));
    const uint32_t u[4] = {p.x, p.y, p.z, p.w};
    for (int q = 0; q < 4 && k + q < E.sdim; ++q) out[k + q] = (u32_to_unit_open(u[q]) * 2.f - 1.f) * 0.1f;
  }
}

// Envs are processed in blocks of VB streams in structure-of-arrays form so that the compiler vectorises every stage over
// the streams (Philox integer rounds, libmvec logf/sinf/cosf/tanhf under -ffast-math, the small mat-vecs).
constexpr int VB = 16;

inline void philox_block(uint64_t seed, uint64_t tick, int64_t e0, int g, uint32_t out[4][VB]) {
#pragma omp simd
  for (int l = 0; l < VB; ++l) {
    const uint64_t ctr_lo = (uint64_t)(e0 + l) * 16 + (uint64_t)g;
    uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)tick, c3 = (uint32_t)(tick >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
      const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
      const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0][l] = c0; out[1][l] = c1; out[2][l] = c2; out[3][l] = c3;
  }
}

void step_range(Env &E, int e0, int e1) {
  const int S = E.sdim, Ad = E.adim;
  const int G = (S + 3) / 4;
  alignas(64) float sT[32][VB], taT[16][VB], xi[32][VB], spT[32][VB], a2[VB], n2[VB];
  alignas(64) uint32_t u[4][VB];
  for (int b0 = e0; b0 < e1; b0 += VB) {
    const int nb = std::min(VB, e1 - b0);
    // gather the block (rows beyond nb replicate the last stream; their results are dropped)
    for (int l = 0; l < VB; ++l) {
      const int e = b0 + std::min(l, nb - 1);
      const float *s = &E.state[(size_t)e * S];
      const float *a = E.a + (size_t)e * Ad;
      for (int j = 0; j < S; ++j) sT[j][l] = s[j];
      for (int j = 0; j < Ad; ++j) taT[j][l] = a[j];
    }
#pragma omp simd
    for (int l = 0; l < VB; ++l) a2[l] = 0.f;
    for (int j = 0; j < Ad; ++j) {
#pragma omp simd
      for (int l = 0; l < VB; ++l) { const float a = taT[j][l]; a2[l] += a * a; taT[j][l] = tanhf(a); }
    }
    // noise: one Philox draw per (stream, group of 4 dims) -> two Box-Muller pairs
    for (int g = 0; g < G; ++g) {
      philox_block(E.seed, E.job_tick, b0, g, u);
      // note: lanes >= nb use counters of streams b0+l (not the replicated stream): harmless, results dropped
      for (int h = 0; h < 2; ++h) {
#pragma omp simd
        for (int l = 0; l < VB; ++l) {
          const float u1 = ((float)(u[2 * h][l] >> 8) + 1.0f) * (1.0f / 16777216.0f);
          const float u2 = ((float)(u[2 * h + 1][l] >> 8) + 1.0f) * (1.0f / 16777216.0f);
          const float rad = sqrtf(-2.0f * logf(u1));
          const float ang = 6.283185307179586f * u2;
          xi[(4 * g + 2 * h) & 31][l] = rad * cosf(ang);
          xi[(4 * g + 2 * h + 1) & 31][l] = rad * sinf(ang);
        }
      }
    }
#pragma omp simd
    for (int l = 0; l < VB; ++l) n2[l] = 0.f;
    for (int kk = 0; kk < S; ++kk) {
      const float *Ar = &E.A[(size_t)kk * S], *Br = &E.B[(size_t)kk * Ad];
      alignas(64) float v[VB];
#pragma omp simd
      for (int l = 0; l < VB; ++l) v[l] = 0.f;
      for (int j = 0; j < S; ++j) {
        const float c = Ar[j];
#pragma omp simd
        for (int l = 0; l < VB; ++l) v[l] += c * sT[j][l];
      }
      for (int j = 0; j < Ad; ++j) {
        const float c = Br[j];
#pragma omp simd
        for (int l = 0; l < VB; ++l) v[l] += c * taT[j][l];
      }
#pragma omp simd
      for (int l = 0; l < VB; ++l) {
        float x = v[l] + 0.01f * xi[kk][l];
        x = fminf(fmaxf(x, -10.f), 10.f);
        spT[kk][l] = x; n2[l] += x * x;
      }
    }
    for (int l = 0; l < nb; ++l) {
      const int e = b0 + l;
      E.r[e] = 1.f - n2[l] / (float)S - 0.1f * a2[l] / (float)Ad;
      E.done[e] = fabsf(spT[0][l]) > 5.f ? 1 : 0;
      float *o = E.sp + (size_t)e * S;
      for (int k = 0; k < S; ++k) o[k] = spT[k][l];
    }
  }
  // the new state becomes current only after every thread is done reading the old one: written by the caller
}

void worker(Env *E, int id) {
  uint64_t seen = 0;
  while (true) {
    // spin briefly (rollouts call step every few hundred microseconds), then sleep on the condition variable
    int spins = 0;
    while (E->gen.load(std::memory_order_acquire) == seen && !E->quit.load(std::memory_order_relaxed)) {
      if (++spins < 20000) CPU_PAUSE();
      else {
        std::unique_lock<std::mutex> lk(E->mu);
        E->cv.wait_for(lk, std::chrono::milliseconds(2), [&] { return E->gen.load() != seen || E->quit.load(); });
      }
    }
    if (E->quit.load()) return;
    seen = E->gen.load(std::memory_order_acquire);
    const int cnt = E->job_e1 - E->job_e0;
    const int per = (cnt + E->n_threads - 1) / E->n_threads;
    const int e0 = E->job_e0 + id * per, e1 = std::min(E->job_e1, e0 + per);
    if (e0 < e1) step_range(*E, e0, e1);
    E->pending.fetch_sub(1, std::memory_order_acq_rel);
  }
}

}  // namespace

extern "C" {

void *crux_hostenv_create(int n_envs, int sdim, int adim, const float *A, const float *B, uint64_t seed, int n_threads) {
  if (n_envs < 1 || sdim < 1 || sdim > 32 || adim < 1 || adim > 16 || !A || !B) return nullptr;
  Env *E = new Env();
  E->n = n_envs; E->sdim = sdim; E->adim = adim; E->seed = seed;
  E->A.assign(A, A + (size_t)sdim * sdim);
  E->B.assign(B, B + (size_t)sdim * adim);
  E->state.assign((size_t)n_envs * sdim, 0.f);
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  if (n_threads <= 0) n_threads = hw;
  n_threads = std::max(1, std::min(n_threads, std::min(hw, 32)));
  n_threads = std::min(n_threads, std::max(1, n_envs / 64));
  E->n_threads = n_threads;
  for (int i = 1; i < n_threads; ++i) E->workers.emplace_back(worker, E, i);  // the caller's thread is worker 0
  return E;
}

void crux_hostenv_destroy(void *h) {
  Env *E = (Env *)h;
  if (!E) return;
  E->quit.store(true);
  { std::lock_guard<std::mutex> lk(E->mu); }
  E->cv.notify_all();
  for (auto &t : E->workers) t.join();
  delete E;
}

int crux_hostenv_threads(void *h) { return h ? ((Env *)h)->n_threads : 0; }

// reset the streams idx[0..n_idx) (all when idx == NULL); writes their new observations to obs_out [n_idx][sdim]
void crux_hostenv_reset(void *h, const int32_t *idx, int n_idx, float *obs_out) {
  Env *E = (Env *)h;
  const int S = E->sdim;
  if (!idx) {
    for (int e = 0; e < E->n; ++e) s0_row(*E, E->tick, e, &E->state[(size_t)e * S]);
    if (obs_out) memcpy(obs_out, E->state.data(), sizeof(float) * (size_t)E->n * S);
    E->tick += 1;
  } else {
    // after a step: same stream position the device env uses for in-step resets (tick of that step + 2^32)
    const uint64_t t = E->last_step_tick + 0x100000000ULL;
    for (int q = 0; q < n_idx; ++q) {
      s0_row(*E, t, idx[q], &E->state[(size_t)idx[q] * S]);
      if (obs_out) memcpy(obs_out + (size_t)q * S, &E->state[(size_t)idx[q] * S], sizeof(float) * S);
    }
  }
}

// @gen(:sp,:r)(mdp, s, a) + isterminal for the streams [e0, e1) (pointers address the FULL arrays; rows e0..e1-1 are
// read / written); the stream state advances to sp (the caller resets ended streams).  A vector step may be issued as several
// consecutive ranges: the noise position (tick) advances when the range that ends at n_envs has been stepped.
void crux_hostenv_step_range(void *h, int32_t e0, int32_t e1, const float *a, float *sp, float *r, uint8_t *done) {
  Env *E = (Env *)h;
  if (e0 < 0) e0 = 0;
  if (e1 > E->n) e1 = E->n;
  if (e0 >= e1) return;
  E->a = a; E->sp = sp; E->r = r; E->done = done; E->job_tick = E->tick; E->job_e0 = e0; E->job_e1 = e1;
  E->last_step_tick = E->tick;
  const int nt = E->n_threads;
  if (nt > 1) {
    E->pending.store(nt - 1, std::memory_order_release);
    E->gen.fetch_add(1, std::memory_order_acq_rel);
    E->cv.notify_all();
  }
  const int per = (e1 - e0 + nt - 1) / nt;
  step_range(*E, e0, std::min(e1, e0 + per));
  while (E->pending.load(std::memory_order_acquire) > 0) CPU_PAUSE();
  memcpy(E->state.data() + (size_t)e0 * E->sdim, sp + (size_t)e0 * E->sdim, sizeof(float) * (size_t)(e1 - e0) * E->sdim);
  if (e1 == E->n) E->tick += 1;
}

void crux_hostenv_step(void *h, const float *a, float *sp, float *r, uint8_t *done) {
  crux_hostenv_step_range(h, 0, ((Env *)h)->n, a, sp, r, done);
}

}  // extern "C"
