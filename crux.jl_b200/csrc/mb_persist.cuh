// Persistent epoch kernel for SMALL minibatches -- the reference-default TrainingParams (training.jl:1-11: batch_size = 128,
// epochs = 80): batch_train! (training.jl:28-55) of one network as ONE launch of ONE thread-block cluster.
//
// With 128-row minibatches a PPO iteration of BASELINE config[1] is up to 80 x 1024 DEPENDENT train! steps per network; launched
// step by step (minibatch kernel + tail, 4 launches) that is 27 us per step of pure launch / dependency latency.  Here a cluster of
// 8 CTAs keeps the whole optimiser state on chip and walks the minibatches without ever leaving the kernel:
//   * every CTA holds the full parameter vector (+ the transposed W2 / W3 of the data-backward GEMMs) in shared memory; the CTA that
//     OWNS a slice of the parameters keeps their fp32 master copy and Adam moments in registers,
//   * CTA c trains on rows [16 c, 16 c + 16) of the minibatch: gather (cp.async, one minibatch ahead), forward, loss head, backward
//     on 16-row tiles with the FFMA building blocks of the fused kernels -> its partial gradient in shared memory,
//   * cluster barrier; the owner of entry i sums the 8 partials through distributed shared memory (fixed order, double), applies Flux
//     Adam and stores the new weight into EVERY CTA's copies; cluster barrier; CTA 0 writes the info record; all CTAs take the same
//     KL early-stop decision from the summed head terms (rl/ppo.jl:59 via training.jl:46,49).
// Same head arithmetic, gradient layout and record fields as fused_minibatch_kernel / adam_body.  A non-finite gradient raises the
// sticky NaN flag and ends the loop (training.jl:20; unlike the step-by-step path the failing step's update has been applied by then --
// the host raises either way).  Single rank; minibatches of up to 128 rows.
// Included by ppo_fused.cu.
#pragma once
// (<cooperative_groups.h> is included at the top of ppo_fused.cu: this file is included inside its anonymous namespace)

namespace mbp {
namespace cg = cooperative_groups;

constexpr int C = 8;            // CTAs per cluster (portable maximum)
constexpr int TR = 16;          // rows per CTA and minibatch
constexpr int MAXB = C * TR;    // 128
constexpr int GMAX = P_MAX + 16;

struct Map {   // floats; P / W2T / W3T sit where stage_params / build_transposes expect them (SmemMap)
  static constexpr int P = SmemMap::P, W2T = SmemMap::W2T, W3T = SmemMap::W3T;
  static constexpr int G = W3T + MAX_O * H;                 // this CTA's partial gradient: [n_params] | dlogΣ[8] | obj, kl, clip, adv, ret, -, -, -
  static constexpr int XT = (G + GMAX + 3) / 4 * 4;         // [2][32][LD16] gathered observations, transposed (double-buffered)
  static constexpr int AT = XT + 2 * MAX_I * LD16;          // [2][8][LD16]  stored actions
  static constexpr int HD = AT + 2 * MAX_O * LD16;          // [2][3][16]    logp_old | advantage | return
  static constexpr int H1T = HD + 2 * 3 * TR;               // [64][LD16]
  static constexpr int H2T = H1T + H * LD16;                // [64][LD16]
  static constexpr int OT = H2T + H * LD16;                 // [8][LD16]
  static constexpr int LS = OT + MAX_O * LD16;              // logΣ[8] (every CTA's copy)
  static constexpr int LSP = LS + 8;                        // logΣ the current minibatch's gradient was taken at (the record's entropy)
  static constexpr int RED = LSP + 8;                       // [8 warps][16] head reduction scratch
  static constexpr int SLOT = RED + 8 * 16;                 // cluster mailboxes: (C unused ints) | norm2[C] (doubles, 8-byte aligned)
  static constexpr int TOT = SLOT + C + 2 * C;              // head sums of the whole minibatch (obj, kl, clip, adv, ret) + pad
  static constexpr int IDX = TOT + 8;                       // [2][16] ints: source rows of this CTA's tile, one minibatch ahead
  static constexpr int NEWV = (IDX + 2 * TR + 3) / 4 * 4;   // the owner's freshly updated slice (pulled by every CTA after the barrier)
  static constexpr int PER_MAX = ((GMAX + C - 1) / C + 3) / 4 * 4;
  static constexpr int MBAR = (NEWV + PER_MAX + 3) / 4 * 4;
  static constexpr int TOTAL = MBAR + 2;
  static constexpr size_t BYTES = (size_t)TOTAL * sizeof(float);
  static_assert(XT % 4 == 0 && H1T % 4 == 0 && H2T % 4 == 0 && OT % 4 == 0 && MBAR % 2 == 0 && (SLOT + C) % 2 == 0 && NEWV % 4 == 0, "alignment");
};

struct Args {
  NetDesc net;                 // net.params = the network's global parameter vector (read at the start, written at the end)
  const float *s, *act, *logp_old, *adv, *ret;
  const int32_t *order;        // [epochs][n] row orders
  int64_t n; int batch, epochs; int64_t max_batches;
  float *ls, *ls_m, *ls_v;     // actor: logΣ vector and its Adam moments (global, in / out); NULL for a critic
  float *m, *v;                // Adam moments of the network (global, in / out)
  int *step_dev;               // Adam step counter (in / out)
  double *beta_cache;          // {t, β1^t, β2^t} cache of the step-by-step path: left consistent
  double eta, b1, b2, eps;
  float eps_clip, lambda_p, lambda_e, target_kl; int a2c;
  float *info;                 // [epochs * ceil(n / batch)][8] records (zeroed by the caller)
  int *ctl;                    // actor: ctl[1] = 1 + index of the minibatch after which training stopped
  unsigned int *err_flags;
  int n_params;
};

// dA^T[k][r] = act'(A^T[k][r]) * sum_{j<J} dC^T[j][r] WT[j][k], in place over A^T; thread = (row t >> 4, 4 columns)
__device__ __forceinline__ void layer_bwd_data16(const float *__restrict__ DCT, int J, const float *__restrict__ WT, float *__restrict__ AT_, int act) {
  const int r = threadIdx.x >> 4, kg = threadIdx.x & 15;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float *dp = DCT + r, *wp = WT + 4 * kg;
#pragma unroll 8
  for (int j = 0; j < J; ++j) {
    const float d = dp[j * LD16];
    const float4 w = *reinterpret_cast<const float4 *>(wp + j * H);
    acc[0] = fmaf(d, w.x, acc[0]); acc[1] = fmaf(d, w.y, acc[1]); acc[2] = fmaf(d, w.z, acc[2]); acc[3] = fmaf(d, w.w, acc[3]);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float *p = AT_ + (4 * kg + c) * LD16 + r;
    *p = acc[c] * act_bwd_from_out(act, *p);
  }
}

template <int HEAD>
__global__ void __cluster_dims__(C, 1, 1) __launch_bounds__(NT, 1) epoch_kernel(Args a) {
  extern __shared__ __align__(16) float sm[];
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const NetDesc nd = a.net;
  const int I = nd.I, O = nd.O, act = nd.act, t = threadIdx.x, NP = a.n_params;
  const int A = HEAD == 0 ? O : 0;                 // logΣ entries trained with the actor
  stage_params(sm, nd, Map::MBAR);
  build_transposes(sm, I, O);
  float *P = sm + Map::P, *G = sm + Map::G, *H1T = sm + Map::H1T, *H2T = sm + Map::H2T, *OT = sm + Map::OT;
  float *LS = P + NP;   // logΣ sits right behind the parameters (P_SMEM reserves 8 floats there): (parameters | logΣ) is ONE vector
  double *norm_slot = reinterpret_cast<double *>(sm + Map::SLOT + C);
  __syncthreads();      // (stage_params' bulk copy may pad up to 16 bytes behind the parameters)
  if (t < 8) LS[t] = (HEAD == 0 && t < O) ? a.ls[t] : 0.f;
  for (int e = t; e < 2 * MAX_I * LD16 + 2 * MAX_O * LD16 + 2 * 3 * TR; e += NT) sm[Map::XT + e] = 0.f;

  // ---- the slice of (parameters | logΣ) this CTA owns: entries lo + t + NT u; master copy and Adam moments in registers
  const int NE = NP + A, per = ((NE + C - 1) / C + 3) / 4 * 4, lo = c * per, hi = min(NE, lo + per);   // slices of a multiple of 4 entries
  constexpr int NU = (GMAX / C + NT) / NT;          // entries per thread (<= 4)
  float pw[NU], pm[NU], pv[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int i = lo + t + NT * u;
    pw[u] = pm[u] = pv[u] = 0.f;
    if (i < hi) {
      if (i < NP) { pw[u] = nd.params[i]; pm[u] = a.m[i]; pv[u] = a.v[i]; }
      else { pw[u] = a.ls[i - NP]; pm[u] = a.ls_m[i - NP]; pv[u] = a.ls_v[i - NP]; }
    }
  }
  const int t0 = *a.step_dev;
  double p1 = pow(a.b1, (double)t0), p2 = pow(a.b2, (double)t0);   // Flux keeps the running products β^t

  const int64_t nmb = (a.n + a.batch - 1) / a.batch;
  const int64_t maxb = a.max_batches > 0 ? a.max_batches : INT64_MAX;
  const int64_t total_mb = min((int64_t)a.epochs * nmb, maxb);
  // minibatch q -> (epoch, index) -> first row in the order array, rows in it
  auto mb_rows = [&](int64_t q, int64_t &off, int &bm) {
    const int64_t e = q / nmb, k = q - e * nmb;
    off = e * a.n + k * a.batch;
    bm = (int)min((int64_t)a.batch, a.n - k * a.batch);
  };
  // gather of this CTA's 16 rows of minibatch q into buffer b (cp.async, 4 bytes per element, transposed on the fly)
  int *IDXs = reinterpret_cast<int *>(sm + Map::IDX);
  auto load_idx = [&](int64_t q) -> int {   // source row of this CTA's tile row t (t < 16) in minibatch q, -1 beyond the minibatch
    int64_t off; int bm;
    mb_rows(q, off, bm);
    const int row_in_mb = c * TR + t;
    return row_in_mb < bm ? __ldg(a.order + off + row_in_mb) : -1;
  };
  auto gather = [&](int64_t q, int b) {   // rows from IDXs[b] (filled before the preceding block barrier)
    const int *idx = IDXs + b * TR;
    float *XTb = sm + Map::XT + b * MAX_I * LD16, *ATb = sm + Map::AT + b * MAX_O * LD16, *HDb = sm + Map::HD + b * 3 * TR;
    for (int e = t; e < TR * I; e += NT) {
      const int r = e / I, i = e - r * I;
      const int row = idx[r];
      cp_async4(XTb + i * LD16 + r, a.s + (int64_t)(row >= 0 ? row : 0) * I + i, row >= 0);
    }
    if (HEAD == 0)
      for (int e = t; e < TR * O; e += NT) {
        const int r = e / O, o = e - r * O;
        const int row = idx[r];
        cp_async4(ATb + o * LD16 + r, a.act + (int64_t)(row >= 0 ? row : 0) * O + o, row >= 0);
      }
    if (t < TR) {
      const int rowi = idx[t];
      const bool live = rowi >= 0;
      const int row = live ? rowi : 0;
      if (HEAD == 0) {
        cp_async4(HDb + t, a.logp_old + row, live);
        cp_async4(HDb + TR + t, a.adv + row, live);
      }
      cp_async4(HDb + 2 * TR + t, a.ret ? a.ret + row : a.s, live && a.ret != nullptr);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int nidx = -1;                                   // index of tile row t for the minibatch AFTER the next one (threads 0..15)
  if (t < TR && total_mb > 0) IDXs[t] = load_idx(0);
  if (t < TR && total_mb > 1) nidx = load_idx(1);
  __syncthreads();
  if (total_mb > 0) gather(0, 0);
  cluster.sync();

  const int kg = t >> 4, jg = t & 15;
  bool stop = false;
  int64_t q = 0;
  for (; q < total_mb && !stop; ++q) {
    const int b = (int)(q & 1);
    int64_t off; int bm;
    mb_rows(q, off, bm);
    const float inv_bg = 1.0f / (float)bm;
    float *XT = sm + Map::XT + b * MAX_I * LD16, *AT_ = sm + Map::AT + b * MAX_O * LD16, *HD = sm + Map::HD + b * 3 * TR;
    // indices of minibatch q + 1 (requested a whole minibatch ago) -> shared; request those of q + 2; then gather q + 1 one barrier later
    if (t < TR) { IDXs[(b ^ 1) * TR + t] = nidx; nidx = q + 2 < total_mb ? load_idx(q + 2) : -1; }
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // the rows of minibatch q have landed (their gather was issued a minibatch ago)
    __syncthreads();
    if (q + 1 < total_mb) gather(q + 1, b ^ 1);
    if (t < 8) sm[Map::LSP + t] = LS[t];
    // ---------------- forward (16-row tile)
    layer_fwd16(XT, I, P, P + off_b1(I), H1T, act);
    __syncthreads();
    layer_fwd16(H1T, H, P + off_W2(I), P + off_b2(I), H2T, act);
    __syncthreads();
    layer_out16(H2T, P + off_W3(I), P + off_b3(I, O), O, OT);
    __syncthreads();
    // ---------------- loss head: dL/dout (scaled by 1/B) replaces out^T; head sums and dlogΣ of this CTA's rows
    float s_obj = 0.f, s_kl = 0.f, s_clip = 0.f, s_adv = 0.f, s_ret = 0.f, dls[MAX_O];
#pragma unroll
    for (int j = 0; j < MAX_O; ++j) dls[j] = 0.f;
    if (t < TR) {
      const bool live = c * TR + t < bm;
      if (HEAD == 0) {
        float logp = 0.f;
        for (int j = 0; j < O; ++j) {
          const float sg = expf(LS[j]);
          const float d = AT_[j * LD16 + t] - OT[j * LD16 + t];
          logp += -(d * d) / (2.f * (sg * sg)) - LOG_SQRT_2PI - LS[j];
        }
        const float Ai = HD[TR + t], old = HD[t];
        float dlogp = 0.f;
        if (live) {
          if (a.a2c) {
            s_obj += logp * Ai;
            dlogp = -a.lambda_p * inv_bg * Ai;
          } else {
            const float rt = expf(logp - old);
            const float lo_ = 1.f - a.eps_clip, hi_ = 1.f + a.eps_clip;
            const float x = rt * Ai, y = fminf(fmaxf(rt, lo_), hi_) * Ai;
            const bool first = !(y < x);  // min(x, y) keeps x on ties
            s_obj += first ? x : y;
            dlogp = first ? -a.lambda_p * inv_bg * x : 0.f;
            s_clip += (rt > hi_ || rt < lo_) ? 1.f : 0.f;
          }
          s_kl += old - logp; s_adv += Ai; s_ret += HD[2 * TR + t];
        }
        for (int j = 0; j < O; ++j) {
          const float sg = expf(LS[j]);
          const float var = sg * sg;
          const float d = AT_[j * LD16 + t] - OT[j * LD16 + t];
          OT[j * LD16 + t] = dlogp * d / var;
          dls[j] += dlogp * (d * d / var - 1.f);
        }
      } else {
        const float d = OT[t] - HD[2 * TR + t];
        if (live) s_obj += d * d;
        OT[t] = live ? 2.f * d * inv_bg : 0.f;
      }
    }
    if (t < 32) {   // the 16 head threads are the first half of warp 0
      float v;
      v = warp_sum(s_obj); if (t == 0) G[NP + 8 + 0] = v;
      v = warp_sum(s_kl); if (t == 0) G[NP + 8 + 1] = v;
      v = warp_sum(s_clip); if (t == 0) G[NP + 8 + 2] = v;
      v = warp_sum(s_adv); if (t == 0) G[NP + 8 + 3] = v;
      v = warp_sum(s_ret); if (t == 0) G[NP + 8 + 4] = v;
#pragma unroll
      for (int j = 0; j < MAX_O; ++j) { v = warp_sum(dls[j]); if (t == 0) G[NP + j] = v; }
    }
    __syncthreads();
    // ---------------- dW3 = h2^T dOut ; db3
    {
      const int k = t & 63, og = t >> 6;
      if (og < O) {
        float a0 = 0.f, a1 = 0.f;
        const bool v1 = og + 4 < O;
#pragma unroll
        for (int r4 = 0; r4 < TR / 4; ++r4) {
          const float4 h = *reinterpret_cast<const float4 *>(H2T + k * LD16 + 4 * r4);
          a0 = dot4(h, *reinterpret_cast<const float4 *>(OT + og * LD16 + 4 * r4), a0);
          if (v1) a1 = dot4(h, *reinterpret_cast<const float4 *>(OT + (og + 4) * LD16 + 4 * r4), a1);
        }
        G[off_W3(I) + k * O + og] = a0;
        if (v1) G[off_W3(I) + k * O + og + 4] = a1;
      }
      if (t >= 128 && t < 128 + O) {
        const int o = t - 128;
        float s3 = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < TR / 4; ++r4) {
          const float4 d = *reinterpret_cast<const float4 *>(OT + o * LD16 + 4 * r4);
          s3 += (d.x + d.y) + (d.z + d.w);
        }
        G[off_b3(I, O) + o] = s3;
      }
    }
    __syncthreads();
    layer_bwd_data16(OT, O, sm + Map::W3T, H2T, act);   // dz2^T in place over h2^T
    __syncthreads();
    // ---------------- dW2 = h1^T dz2 ; db2
    {
      float acc2[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = 0.f;
#pragma unroll
      for (int r4 = 0; r4 < TR / 4; ++r4) {
        float4 hv[4], zv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          hv[x] = *reinterpret_cast<const float4 *>(H1T + (kg + 16 * x) * LD16 + 4 * r4);
          zv[x] = *reinterpret_cast<const float4 *>(H2T + (jg + 16 * x) * LD16 + 4 * r4);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc2[i][j] = dot4(hv[i], zv[j], acc2[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) G[off_W2(I) + (kg + 16 * i) * H + jg + 16 * j] = acc2[i][j];
      if (t < 64) {
        float s2 = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < TR / 4; ++r4) {
          const float4 d = *reinterpret_cast<const float4 *>(H2T + t * LD16 + 4 * r4);
          s2 += (d.x + d.y) + (d.z + d.w);
        }
        G[off_b2(I) + t] = s2;
      }
    }
    __syncthreads();
    layer_bwd_data16(H2T, H, sm + Map::W2T, H1T, act);   // dz1^T in place over h1^T
    __syncthreads();
    // ---------------- dW1 = x^T dz1 ; db1
    {
      float acc1[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc1[i][j] = 0.f;
#pragma unroll
      for (int r4 = 0; r4 < TR / 4; ++r4) {
        float4 zv[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) zv[x] = *reinterpret_cast<const float4 *>(H1T + (jg + 16 * x) * LD16 + 4 * r4);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int ii = kg + 16 * i;
          if (ii < I) {
            const float4 xv = *reinterpret_cast<const float4 *>(XT + ii * LD16 + 4 * r4);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc1[i][j] = dot4(xv, zv[j], acc1[i][j]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int ii = kg + 16 * i;
        if (ii < I)
#pragma unroll
          for (int j = 0; j < 4; ++j) G[ii * H + jg + 16 * j] = acc1[i][j];
      }
      if (t >= 64 && t < 128) {
        float s1 = 0.f;
#pragma unroll
        for (int r4 = 0; r4 < TR / 4; ++r4) {
          const float4 d = *reinterpret_cast<const float4 *>(H1T + (t - 64) * LD16 + 4 * r4);
          s1 += (d.x + d.y) + (d.z + d.w);
        }
        G[off_b1(I) + (t - 64)] = s1;
      }
    }
    __syncthreads();
    cluster.sync();
    // A non-finite gradient (training.jl:20) shows in the squared norm every CTA reads after the second barrier: the flag is raised and the
    // loop ends there; that step's Adam update has already been applied (the host raises on the flag either way).
    // head sums of the whole minibatch: threads 0..4 add the C tails in rank order (every CTA computes the same totals)
    if (t < 5) {
      float s5 = 0.f;
#pragma unroll
      for (int x = 0; x < C; ++x) s5 += cluster.map_shared_rank(G, x)[NP + 8 + t];
      sm[Map::TOT + t] = s5;
    }
    const float cnt = (float)bm;
    // ---------------- the owner of entry i: sum of the C partials (fixed order, double), norm term, Flux Adam, broadcast
    p1 *= a.b1; p2 *= a.b2;
    const double c1 = 1.0 - p1, c2 = 1.0 - p2;
    double sq = 0.0;
    float gsum[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int i = lo + t + NT * u;
      gsum[u] = 0.f;
      if (i < hi) {
        double acc = 0.0;
#pragma unroll
        for (int x = 0; x < C; ++x) acc += (double)cluster.map_shared_rank(G, x)[i];
        float g = (float)acc;
        if (i >= NP) g += -a.lambda_e;   // d(λe·e_loss)/dlogΣ = -λe (policies.jl:348: entropy = const + sum(logΣ))
        gsum[u] = g;
        sq += (double)g * (double)g;
      }
    }
    {   // ||g||^2 of the slice -> every CTA's mailbox
      __shared__ double shn[8];
      sq = warp_sum_d(sq);
      if ((t & 31) == 0) shn[t >> 5] = sq;
      __syncthreads();
      if (t < C) {
        double s8 = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) s8 += shn[w];
        cluster.map_shared_rank(norm_slot, t)[c] = s8;
      }
    }
    float tot[5];   // (the barrier above also published TOT)
#pragma unroll
    for (int y = 0; y < 5; ++y) tot[y] = sm[Map::TOT + y];
    const float kl = tot[1] / cnt;
    {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int i = lo + t + NT * u;
        if (i < hi) {
          const double g = (double)gsum[u];
          const float mt = (float)(a.b1 * (double)pm[u] + (1.0 - a.b1) * g);
          const float vt = (float)(a.b2 * (double)pv[u] + (1.0 - a.b2) * g * g);
          pm[u] = mt; pv[u] = vt;
          const float pn = pw[u] - (float)((double)mt / c1 / (sqrt((double)vt / c2) + a.eps) * a.eta);
          pw[u] = pn;
          sm[Map::NEWV + (i - lo)] = pn;   // parked locally: every CTA pulls the slices with 128-bit DSMEM loads after the barrier
        }
      }
    }
    cluster.sync();
    {
      // (parameters | logΣ) <- the owners' slices; then the transposed W2 / W3 of the data-backward GEMMs are rebuilt locally
      constexpr int NV4 = (GMAX / 4 + NT - 1) / NT;   // 128-bit chunks per thread: all remote loads are issued before the first store
      float4 nv[NV4];
#pragma unroll
      for (int u = 0; u < NV4; ++u) {
        const int i4 = 4 * (t + NT * u);
        if (i4 < NE) { const int x = i4 / per; nv[u] = *reinterpret_cast<const float4 *>(cluster.map_shared_rank(sm + Map::NEWV, x) + (i4 - x * per)); }
      }
#pragma unroll
      for (int u = 0; u < NV4; ++u) {
        const int i4 = 4 * (t + NT * u);
        if (i4 + 3 < NE) *reinterpret_cast<float4 *>(P + i4) = nv[u];
        else if (i4 < NE) { const float e4[4] = {nv[u].x, nv[u].y, nv[u].z, nv[u].w}; for (int j = 0; i4 + j < NE; ++j) P[i4 + j] = e4[j]; }
      }
      __syncthreads();
      build_transposes(sm, I, O);
    }
    // ---------------- info record (CTA 0), early stop (every CTA decides alike)
    if (c == 0 && t == 0) {
      double n2 = 0.0;
#pragma unroll
      for (int x = 0; x < C; ++x) n2 += norm_slot[x];
      float *rec = a.info + q * CRUX_PPO_INFO_STRIDE;
      if (HEAD == 0) {
        float sls = 0.f;   // the loss is logged as computed BEFORE the update: logΣ of this minibatch's forward pass
        for (int j = 0; j < O; ++j) sls += sm[Map::LSP + j];
        const float entropy = 1.4189385332046727f + sls;
        const float p_loss = -(tot[0] / cnt);
        rec[CRUX_PPO_LOSS] = a.lambda_p * p_loss + a.lambda_e * (-entropy);
        rec[CRUX_PPO_ENTROPY] = entropy;
        rec[CRUX_PPO_KL] = kl;
        rec[CRUX_PPO_CLIP_FRAC] = a.a2c ? 0.f : tot[2] / cnt;
        rec[CRUX_PPO_AVG_ADV] = tot[3] / cnt;
        rec[CRUX_PPO_AVG_RET] = tot[4] / cnt;
        if (a.ctl && kl > a.target_kl) a.ctl[1] = (int)q + 1;
      } else {
        rec[CRUX_PPO_LOSS] = tot[0] / cnt;
      }
      rec[CRUX_PPO_GRAD_NORM] = (float)sqrt(n2);
      rec[CRUX_PPO_VALID] = 1.f;
      if (!isfinite(n2)) atomicOr(a.err_flags, CRUX_FLAG_NAN);
    }
    {
      double n2a = 0.0;
#pragma unroll
      for (int x = 0; x < C; ++x) n2a += norm_slot[x];
      if (!isfinite(n2a)) stop = true;
    }
    if (HEAD == 0 && kl > a.target_kl) stop = true;   // this minibatch was applied; later ones are skipped (rl/ppo.jl:59)
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // ---------------- state back to global memory: parameters, moments, logΣ, step counter
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int i = lo + t + NT * u;
    if (i < hi) {
      if (i < NP) { const_cast<float *>(nd.params)[i] = pw[u]; a.m[i] = pm[u]; a.v[i] = pv[u]; }
      else { a.ls[i - NP] = pw[u]; a.ls_m[i - NP] = pm[u]; a.ls_v[i - NP] = pv[u]; }
    }
  }
  if (c == 0 && t == 0) {
    const int tf = t0 + (int)q;
    *a.step_dev = tf;
    if (a.beta_cache) { a.beta_cache[1] = p1; a.beta_cache[2] = p2; a.beta_cache[0] = (double)tf; }
  }
  cluster.sync();   // no CTA may exit while its shared memory can still be addressed by a peer
}

}  // namespace mbp
