/* CPU oracle (plain C) for the advantage / return / PER-index part of the Crux.jl
 * hot path.  TEST INFRASTRUCTURE ONLY: loaded by tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Never linked into libcrux_cuda.so.
 *
 * Parity status: the recurrences follow src/sampler.jl:262-281 literally (float32,
 * no FMA contraction: build with -ffp-contract=off).  The reference's tests do not pin
 * GAE values (test/gym/sampler_tests.jl:75-81 asserts nothing) -> "parity unpinned";
 * the known-answer vector derived from that test's inputs is checked in
 * tests/test_oracle_golden.py.
 *
 * Layout: rollout columns are [T][N] (row t*N+e), each env stream e is one
 * reference "sampler"; episode_end[t][e] closes the episode range at row t
 * (terminate_episode! sampler.jl:53-57).
 */
#include <stdint.h>
#include <stddef.h>
#include <pthread.h>

/* fill_gae! (sampler.jl:262-273) + fill_returns! (:275-281) for env streams [e0, e1). */
static void gae_range(const float *r, const uint8_t *done, const uint8_t *episode_end,
                      const float *v_s, const float *v_sp, int64_t T, int64_t N,
                      float gamma, float lambda, float *adv, float *ret, int64_t e0, int64_t e1)
{
    const float c = lambda * gamma;
    for (int64_t e = e0; e < e1; ++e) {
        float A = 0.0f, R = 0.0f;
        for (int64_t t = T - 1; t >= 0; --t) {
            const int64_t i = t * N + e;
            if (episode_end[i]) { A = 0.0f; R = 0.0f; }
            /* A = c*A + r + (1 - done)*gamma*V(sp) - V(s)   (left-to-right like Julia) */
            float x = c * A;
            x = x + r[i];
            float nd = (1.0f - (float)done[i]) * gamma;
            x = x + nd * v_sp[i];
            A = x - v_s[i];
            R = r[i] + gamma * R;
            if (adv) adv[i] = A;
            if (ret) ret[i] = R;
        }
    }
}

typedef struct {
    const float *r; const uint8_t *done, *ee; const float *v_s, *v_sp; int64_t T, N;
    float gamma, lambda; float *adv, *ret; int64_t e0, e1;
} gae_job;

static void *gae_worker(void *p)
{
    gae_job *j = (gae_job *)p;
    gae_range(j->r, j->done, j->ee, j->v_s, j->v_sp, j->T, j->N, j->gamma, j->lambda, j->adv, j->ret, j->e0, j->e1);
    return NULL;
}

static int g_threads = 1;
void oracle_set_threads(int n) { g_threads = n < 1 ? 1 : (n > 256 ? 256 : n); }
int oracle_get_threads(void) { return g_threads; }

/* The reference is single-threaded; threads only split the independent env streams
 * (used by bench.py's "all host threads" baseline leg). */
void oracle_gae_returns(const float *r, const uint8_t *done, const uint8_t *episode_end,
                        const float *v_s, const float *v_sp, int64_t T, int64_t N,
                        float gamma, float lambda, float *adv, float *ret)
{
    int nt = g_threads;
    if (nt > N) nt = (int)(N > 0 ? N : 1);
    if (nt <= 1) { gae_range(r, done, episode_end, v_s, v_sp, T, N, gamma, lambda, adv, ret, 0, N); return; }
    pthread_t th[256];
    gae_job jobs[256];
    for (int k = 0; k < nt; ++k) {
        gae_job j = {r, done, episode_end, v_s, v_sp, T, N, gamma, lambda, adv, ret, N * k / nt, N * (k + 1) / nt};
        jobs[k] = j;
        pthread_create(&th[k], NULL, gae_worker, &jobs[k]);
    }
    for (int k = 0; k < nt; ++k) pthread_join(th[k], NULL);
}

/* searchsortedfirst over a float32 prefix array with Float64 thresholds
 * (experience_buffer.jl:333-341).  ids are 0-based; N means "past the end". */
void oracle_per_indices(const float *cumsum, int64_t N, int64_t B, const double *rands, int64_t *ids)
{
    const float ptot = cumsum[N - 1];
    const float dp = ptot / (float)B;
    for (int64_t j = 0; j < B; ++j) {
        const double x = ((double)(j + 1) + rands[j] - 1.0) * (double)dp;
        int64_t lo = 0, hi = N;
        while (lo < hi) {
            int64_t mid = lo + (hi - lo) / 2;
            if ((double)cumsum[mid] < x) lo = mid + 1; else hi = mid;
        }
        ids[j] = lo;
    }
}

/* mod1.(next_ind : next_ind+n-1, C) with 0-based in/out (experience_buffer.jl:236). */
void oracle_ring_indices(int64_t next0, int64_t n, int64_t C, int64_t *out)
{
    for (int64_t j = 0; j < n; ++j) out[j] = (next0 + j) % C;
}
