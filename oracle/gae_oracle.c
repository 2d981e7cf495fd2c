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

/* fill_gae! (sampler.jl:262-273) + fill_returns! (:275-281) for every env stream. */
void oracle_gae_returns(const float *r, const uint8_t *done, const uint8_t *episode_end,
                        const float *v_s, const float *v_sp, int64_t T, int64_t N,
                        float gamma, float lambda, float *adv, float *ret)
{
    const float c = lambda * gamma;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < N; ++e) {
        float A = 0.0f, R = 0.0f;
        for (int64_t t = T - 1; t >= 0; --t) {
            const int64_t i = t * N + e;
            if (episode_end[i]) { A = 0.0f; R = 0.0f; }
            /* A = c*A + r + (1 - done)*γ*V(sp) - V(s)   (left-to-right like Julia) */
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
