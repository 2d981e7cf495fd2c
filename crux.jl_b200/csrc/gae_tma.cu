// Hot path (ii), large rollouts: fill_gae! / fill_returns! (src/sampler.jl:262-281) as a TMA-fed streaming scan.
//
// Same recurrences as gae.cu,
//      A_t = delta_t + c_t A_{t+1},  c_t = episode_end_t ? 0 : lambda*gamma,   delta_t = r_t + (1-done_t) gamma V(sp_t) - V(s_t)
//      R_t = r_t     + g_t R_{t+1},  g_t = episode_end_t ? 0 : gamma
// over a [T][N] rollout (row t*N+e), 22 B per transition read/written once.  gae.cu keeps a whole chunk in registers and pays a
// CTA-wide load -> barrier -> look-back -> store sequence per chunk (no loads in flight for half of a CTA's life: 55 % of the
// copy roofline).  Here the memory system is never idle and no LSU instruction touches global memory:
//
//   tile = 128 adjacent env streams, chunk = CS time steps, stage = one [CS][128] box of each of the five input columns in a
//   shared-memory ring.  Four warp roles per CTA, chained by mbarriers (full -> prepped -> chained -> empty):
//     loader  (1 thread) : 2-D TMA tensor loads (cp.async.bulk.tensor.2d, complete_tx) of the next free stage; rows / streams
//                          outside the rollout are zero-filled by the TMA unit (zeros are the identity of both recurrences).
//     prep    (2 warps)  : delta_t for every element of the stage, row-parallel, 128-bit shared-memory accesses, in place.
//     chain   (2 warps)  : thread = 2 adjacent streams; walks the stage backwards in time with the carry (A, R) in registers --
//                          the plain sequential recurrence, bit-identical for every tiling -- and leaves A, R in the stage.
//     storer  (1 thread) : 2-D TMA tensor stores of the advantage / return boxes (clipped to the rollout by the TMA unit),
//                          frees the stage when the stores have read it.
//   Persistent CTAs take work items (time segment, tile) from an atomic counter, latest segment first.  Inside a segment the
//   carry never leaves the registers; between segments it crosses global memory once per stream (value + release flag, acquired
//   by the chain warp that starts the earlier segment).  An item is only waited on by items taken later from the counter, i.e.
//   by CTAs that were resident after its owner: deadlock-free for any residency.
//   The sequential chains are the critical resource (one chain warp per tile column group): the tile is as narrow as it can be while
//   every tile still gets its own SM -- 32-stream tiles (one stream per chain lane) up to 32 x SM-count streams, 64-stream tiles up
//   to 64 x SM-count, 128-stream tiles above.  The path is used for rollouts wider than 2048 streams and at least 64 steps long;
//   narrower or shorter ones stay on the time-parallel scan of gae.cu (profiles/r1_gae_sweep.log: [1024,4096] 3.6 TB/s against
//   2.65 for the scan, [2048,4096] 4.5 against 3.1, [2048,2560] 2.9 against 2.75).
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

namespace {

// streams per tile TS = 64 * NCW, NCW = chain warps (64 streams each, 2 per lane): 128-stream tiles for wide rollouts, 64-stream
// tiles (twice as many sequential chains / CTAs) for narrower ones
constexpr int NPW = 2;             // prep warps
constexpr int W_LOAD = 0, W_CHAIN = 1;   // then NCW chain warps, NPW prep warps, the storer warp
constexpr int n_threads(int ncw) { return 32 * (1 + ncw + NPW + 1); }
constexpr int MAX_STAGES = 8;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.b32 %0, 1, 0, p; }"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0),
               "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct TmaArgs {
  int64_t T, N;
  float gamma, lambda;
  int has_adv, has_ret;
  int n_chunks;      // ceil(T / CS)
  int seg_chunks;    // chunks per segment
  int n_segs;        // ceil(n_chunks / seg_chunks)
  int tiles;         // ceil(N / TS)
  int n_items;       // n_segs * tiles
  int stages;
  unsigned int *counter;   // work-item counter: never reset, this launch owns [counter_base, counter_base + n_items + grid)
  unsigned int counter_base;
  unsigned int epoch;      // launch number: a carry flag is valid when it equals the epoch (no per-launch memset)
  unsigned int *flags;     // [tiles][n_segs][NCW]
  float4 *carry;           // [tiles][n_segs][NCW][32]   (A0, R0, A1, R1) of a lane's two streams at the first step of the segment
  unsigned int *err_flags;
};

// stage layout (bytes): r [CS][TS] f32 (-> returns) | V(s) [CS][TS] f32 (-> delta -> advantages) | V(sp) [CS][TS] f32 |
//                       done [CS][TS] u8 | episode_end [CS][TS] u8
template <int CS, int TS>
struct StageMap {
  static constexpr int F = CS * TS * 4, B = CS * TS;
  static constexpr int R = 0, VS = F, VSP = 2 * F, DN = 3 * F, EE = 3 * F + B;
  static constexpr int BYTES = 3 * F + 2 * B;
  static_assert(BYTES % 128 == 0, "TMA boxes must stay 128-byte aligned");
};

template <int CS, int NCW, int SPL>
__global__ void __launch_bounds__(n_threads(NCW)) gae_tma_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_vs,
                                                           const __grid_constant__ CUtensorMap tm_vsp, const __grid_constant__ CUtensorMap tm_dn,
                                                           const __grid_constant__ CUtensorMap tm_ee, const __grid_constant__ CUtensorMap tm_adv,
                                                           const __grid_constant__ CUtensorMap tm_ret, const TmaArgs a) {
  constexpr int TS = 32 * SPL * NCW, W_PREP = W_CHAIN + NCW, W_STORE = W_PREP + NPW;   // SPL = streams per chain lane (2, or 1 for 32-stream tiles)
  using SM = StageMap<CS, TS>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES], prep_bar[MAX_STAGES], chain_bar[MAX_STAGES], empty_bar[MAX_STAGES];
  __shared__ int4 meta[MAX_STAGES];   // x: tile (-1: no more work), y: chunk, z: segment, w: bit0 latest chunk of the segment, bit1 earliest
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int NST = a.stages;
  if (t == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&prep_bar[s]), NPW);
      mbar_init(smem_u32(&chain_bar[s]), NCW); mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (w == W_LOAD) {
    // ================================================================ loader (one thread)
    if (lane != 0) return;
    int s = 0; uint32_t ph = 0;   // ring position + phase parity of the current pass over the ring
    for (;;) {
      const unsigned int item_u = atomicAdd(a.counter, 1u) - a.counter_base;   // the counter is never reset: launches own consecutive ranges
      const bool more = item_u < (unsigned int)a.n_items;
      const int item = (int)item_u;
      const int q = more ? item / a.tiles : 0, tile = more ? item - q * a.tiles : -1;
      const int seg = a.n_segs - 1 - q;                                   // latest segment first
      const int c_lo = seg * a.seg_chunks;
      const int c_hi = min(a.n_chunks, c_lo + a.seg_chunks) - 1;
      for (int c = c_hi; c >= (more ? c_lo : c_hi); --c) {
        mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);                       // the storer has drained this slot
        meta[s] = make_int4(tile, c, seg, (c == c_hi ? 1 : 0) | (c == c_lo ? 2 : 0));
        const uint32_t bar = smem_u32(&full_bar[s]);
        if (!more) { mbar_arrive(bar); break; }
        mbar_expect_tx(bar, (uint32_t)SM::BYTES);
        const uint32_t dst = smem_u32(smem + (size_t)s * SM::BYTES);
        const int x = tile * TS, y = c * CS;
        tma_load_2d(dst + SM::R, &tm_r, x, y, bar);
        tma_load_2d(dst + SM::VS, &tm_vs, x, y, bar);
        tma_load_2d(dst + SM::VSP, &tm_vsp, x, y, bar);
        tma_load_2d(dst + SM::DN, &tm_dn, x, y, bar);
        tma_load_2d(dst + SM::EE, &tm_ee, x, y, bar);
        if (++s == NST) { s = 0; ph ^= 1; }
      }
      if (!more) break;
    }
    return;
  }

  if (w >= W_PREP && w < W_PREP + NPW) {
    // ================================================================ prep warps: delta = (r + (1-done) gamma V(sp)) - V(s), in place over V(s)
    const int pw = w - W_PREP;
    const float gamma = a.gamma;
    constexpr int PER = CS * TS / 4 / (NPW * 32), PB = 4;   // float4 elements per lane (the stage arrays are dense: flat indexing)
    static_assert(PER % PB == 0, "prep batches");
    int s = 0; uint32_t ph = 0;
    for (;;) {
      mbar_wait(smem_u32(&full_bar[s]), ph);
      const bool stop = meta[s].x < 0;
      if (!stop) {
        unsigned char *st = smem + (size_t)s * SM::BYTES;
        float4 *pr = reinterpret_cast<float4 *>(st + SM::R), *pv = reinterpret_cast<float4 *>(st + SM::VS), *pp = reinterpret_cast<float4 *>(st + SM::VSP);
        const uchar4 *pdn = reinterpret_cast<const uchar4 *>(st + SM::DN);
#pragma unroll
        for (int b = 0; b < PER; b += PB) {
          float4 r4[PB], va[PB], vb[PB];
          uchar4 dn[PB];
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int f = (pw * PER + b + j) * 32 + lane;
            r4[j] = pr[f]; va[j] = pv[f]; vb[j] = pp[f]; dn[j] = pdn[f];
          }
#pragma unroll
          for (int j = 0; j < PB; ++j) {
            const int f = (pw * PER + b + j) * 32 + lane;
            float4 d;
            d.x = (r4[j].x + (dn[j].x ? 0.f : gamma) * vb[j].x) - va[j].x;
            d.y = (r4[j].y + (dn[j].y ? 0.f : gamma) * vb[j].y) - va[j].y;
            d.z = (r4[j].z + (dn[j].z ? 0.f : gamma) * vb[j].z) - va[j].z;
            d.w = (r4[j].w + (dn[j].w ? 0.f : gamma) * vb[j].w) - va[j].w;
            pv[f] = d;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&prep_bar[s]));
      if (stop) break;
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    return;
  }

  if (w >= W_CHAIN && w < W_CHAIN + NCW) {
    // ================================================================ chain warps: lane = streams (32*cw + lane) * SPL + {0, SPL-1} of the tile
    const int cw = w - W_CHAIN;
    const float gamma = a.gamma, cl = a.lambda * a.gamma;
    float A0 = 0.f, R0 = 0.f, A1 = 0.f, R1 = 0.f;
    bool bad = false;
    const int col = SPL * (32 * cw + lane);
    int s = 0; uint32_t ph = 0;
    for (;;) {
      mbar_wait(smem_u32(&prep_bar[s]), ph);
      const int4 m = meta[s];
      if (m.x < 0) {
        if (lane == 0) mbar_arrive(smem_u32(&chain_bar[s]));
        break;
      }
      const int tile = m.x, seg = m.z;
      unsigned char *st = smem + (size_t)s * SM::BYTES;
      if (m.w & 1) {   // carry into the latest chunk of this segment: zero at the end of the rollout, else published by the later segment
        A0 = R0 = A1 = R1 = 0.f;
        if (seg + 1 < a.n_segs) {
          const size_t slot = ((size_t)tile * a.n_segs + (seg + 1)) * NCW + cw;
          const unsigned int *f = a.flags + slot;
          while (ld_acquire(f) != a.epoch) { __nanosleep(40); }
          const float4 cin = __ldcg(a.carry + slot * 32 + lane);
          A0 = cin.x; R0 = cin.y; A1 = cin.z; R1 = cin.w;
        }
      }
      if constexpr (SPL == 2) {
        float2 *pr = reinterpret_cast<float2 *>(st + SM::R + col * 4);     // r     -> returns
        float2 *pd = reinterpret_cast<float2 *>(st + SM::VS + col * 4);    // delta -> advantages
        const uchar2 *pe = reinterpret_cast<const uchar2 *>(st + SM::EE + col);
        // the whole stage into registers first (independent shared-memory loads), then the dependent FMA chain, then the stores
        float2 r2[CS], d2[CS];
        uchar2 e2[CS];
#pragma unroll
        for (int i = 0; i < CS; ++i) { r2[i] = pr[i * (TS / 2)]; d2[i] = pd[i * (TS / 2)]; e2[i] = pe[i * (TS / 2)]; }
#pragma unroll
        for (int i = CS - 1; i >= 0; --i) {
          // an episode end cuts the trace: A = delta, R = r (select AFTER the FMA so that a NaN/Inf never crosses an episode boundary)
          const float a0 = fmaf(cl, A0, d2[i].x), a1 = fmaf(cl, A1, d2[i].y);
          const float q0 = fmaf(gamma, R0, r2[i].x), q1 = fmaf(gamma, R1, r2[i].y);
          A0 = e2[i].x ? d2[i].x : a0; A1 = e2[i].y ? d2[i].y : a1;
          R0 = e2[i].x ? r2[i].x : q0; R1 = e2[i].y ? r2[i].y : q1;
          d2[i] = make_float2(A0, A1);
          r2[i] = make_float2(R0, R1);
          bad |= (A0 != A0) | (A1 != A1);
        }
#pragma unroll
        for (int i = 0; i < CS; ++i) { pd[i * (TS / 2)] = d2[i]; pr[i * (TS / 2)] = r2[i]; }
      } else {   // one stream per lane: 32-stream tiles, twice as many sequential chains for a rollout of the same width
        float *pr = reinterpret_cast<float *>(st + SM::R + col * 4);
        float *pd = reinterpret_cast<float *>(st + SM::VS + col * 4);
        const unsigned char *pe = st + SM::EE + col;
        float r1[CS], d1[CS];
        unsigned char e1[CS];
#pragma unroll
        for (int i = 0; i < CS; ++i) { r1[i] = pr[i * TS]; d1[i] = pd[i * TS]; e1[i] = pe[i * TS]; }
#pragma unroll
        for (int i = CS - 1; i >= 0; --i) {
          const float a0 = fmaf(cl, A0, d1[i]);
          const float q0 = fmaf(gamma, R0, r1[i]);
          A0 = e1[i] ? d1[i] : a0;
          R0 = e1[i] ? r1[i] : q0;
          d1[i] = A0; r1[i] = R0;
          bad |= (A0 != A0);
        }
#pragma unroll
        for (int i = 0; i < CS; ++i) { pd[i * TS] = d1[i]; pr[i * TS] = r1[i]; }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the TMA store (async proxy) reads what this thread wrote
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&chain_bar[s]));
      if ((m.w & 2) && seg > 0) {   // hand the carry to the earlier segment of this tile
        const size_t slot = ((size_t)tile * a.n_segs + seg) * NCW + cw;
        __stcg(a.carry + slot * 32 + lane, make_float4(A0, R0, A1, R1));
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(a.flags + slot, a.epoch);
      }
      if (++s == NST) { s = 0; ph ^= 1; }
    }
    if (bad) atomicOr(a.err_flags, CRUX_FLAG_NAN);  // sampler.jl:270 @assert !isnan(A)
    return;
  }

  // ================================================================== storer (one thread)
  if (lane != 0) return;
  // Up to LAG store groups stay in flight: stage k is handed back to the loader once the group committed LAG stages later shows
  // that the TMA unit has finished READING it (waiting for every group right away would serialise one store latency per stage).
  constexpr int LAG = 2;
  int s = 0, s_free = 0, pending = 0; uint32_t ph = 0;
  for (;;) {
    mbar_wait(smem_u32(&chain_bar[s]), ph);
    const int4 m = meta[s];
    if (m.x < 0) break;
    const uint32_t src = smem_u32(smem + (size_t)s * SM::BYTES);
    const int x = m.x * TS, y = m.y * CS;
    if (a.has_adv) tma_store_2d(&tm_adv, x, y, src + SM::VS);
    if (a.has_ret) tma_store_2d(&tm_ret, x, y, src + SM::R);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (++pending > LAG) {
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(LAG) : "memory");
      mbar_arrive(smem_u32(&empty_bar[s_free]));
      if (++s_free == NST) s_free = 0;
      --pending;
    }
    if (++s == NST) { s = 0; ph ^= 1; }
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  for (; pending > 0; --pending) {   // the loader may still be waiting for a free slot to post the end-of-work marker
    mbar_arrive(smem_u32(&empty_bar[s_free]));
    if (++s_free == NST) s_free = 0;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

bool make_map(CUtensorMap *m, const void *base, bool is_u8, int64_t T, int64_t N, int cs, int ts) {
  const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)T};
  const cuuint64_t strides[1] = {(cuuint64_t)N * (is_u8 ? 1 : 4)};
  const cuuint32_t box[2] = {(cuuint32_t)ts, (cuuint32_t)cs};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_fn()(m, is_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

// Returns CRUX_OK with *handled = 0 when the shape is not eligible (the register-resident scan of gae.cu runs instead).
// Eligibility: T >= 64, N > 2048 and a multiple of 16 (TMA row pitch of the u8 columns), 16-byte aligned columns, TMA encode available.
// CRUX_GAE=scan forces the gae.cu kernel, CRUX_GAE=tma forces this one for any eligible T >= 1;
// CRUX_GAE_CFG="CS,STAGES,SEG_CHUNKS,CTAS_PER_SM,CHAIN_WARPS,STREAMS_PER_LANE" overrides the tuning (0 = default; tests use it to exercise the
// segment hand-off and both tile widths).
int gae_tma_launch(crux_ctx *ctx, const float *r, const uint8_t *done, const uint8_t *ee, const float *vs, const float *vsp, int64_t T, int64_t N,
                   float gamma, float lambda, float *adv, float *ret, int *handled) {
  *handled = 0;
  const char *mode = getenv("CRUX_GAE");
  if (mode && !strcmp(mode, "scan")) return CRUX_OK;
  const bool forced = mode && !strcmp(mode, "tma");
  // every tile is one sequential chain: the narrowest tile that still leaves one tile per SM (32, 64 or 128 streams); rollouts of at
  // most 2048 streams or fewer than 64 steps stay on the time-parallel scan of gae.cu (measured cross-overs, profiles/r1_gae_sweep.log:
  // [2048,8192] 5.4 TB/s with 64-stream tiles against 4.0 with 128 and 3.2 for the scan; [2048,4096] 4.5 with 32-stream tiles against
  // 3.5 with 64 and 3.1 for the scan)
  if (!forced && (T < 64 || N <= 2048)) return CRUX_OK;
  if (N % 16 != 0 || N >= ((int64_t)1 << 31) - 128 || T >= ((int64_t)1 << 31) - 64) return CRUX_OK;
  const uintptr_t al = (uintptr_t)r | (uintptr_t)done | (uintptr_t)ee | (uintptr_t)vs | (uintptr_t)vsp | (uintptr_t)adv | (uintptr_t)ret;
  if (al & 15) return CRUX_OK;
  if (!encode_fn()) return CRUX_OK;
  {  // the launch arguments carry per-launch state (epoch, counter base): a captured graph would replay stale values
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return CRUX_OK;
  }

  // measured best on B200 (profiles/r1_gae_sweep.log): 16-step stages, 4 stages loading + 2 draining per CTA, one CTA per SM;
  // with more tiles than SMs two smaller CTAs per SM so that every tile still is a single segment
  // 64-stream tiles while every tile still gets its own SM (N <= 64 x SMs = 9472 on B200), 128-stream tiles beyond
  // ... and 32-stream tiles (one stream per chain lane) while even those all get their own SM (N <= 32 x SMs = 4736)
  int ncw = cdiv(N, 64) <= (int64_t)ctx->num_sms ? 1 : 2, cs = 0, stages = 0, seg_chunks = 0, per_sm = 0;
  int spl = cdiv(N, 32) <= (int64_t)ctx->num_sms ? 1 : 2;
  if (const char *cfg = getenv("CRUX_GAE_CFG")) sscanf(cfg, "%d,%d,%d,%d,%d,%d", &cs, &stages, &seg_chunks, &per_sm, &ncw, &spl);
  if (ncw != 1 && ncw != 2) ncw = 2;
  if (spl != 1 && spl != 2) spl = 2;
  if (spl == 1) ncw = 1;
  const int TS = 32 * spl * ncw;
  const int tiles_ = (int)cdiv(N, TS);
  if (spl == 1) { if (cs != 32 && cs != 64) cs = 64; }                    // 28 KB stages at 64 steps
  else if (cs != 16 && cs != 32) cs = ncw == 2 ? 16 : 32;                 // 28 KB stages either way
  if (stages <= 0) stages = tiles_ > ctx->num_sms ? 4 : (spl == 1 ? 7 : 6);
  if (per_sm <= 0) per_sm = tiles_ > ctx->num_sms ? 2 : 1;
  if (stages < 3) stages = 3;   // the storer keeps two stages in flight
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const size_t stage_bytes = (size_t)cs * TS * 14;
  while (stages > 3 && stages * stage_bytes > 226 * 1024) --stages;   // 227 KB per CTA minus the static barriers
  const int n_chunks = (int)cdiv(T, cs);
  const int tiles = (int)cdiv(N, TS);
  const int G_max = ctx->num_sms * per_sm;
  if (seg_chunks <= 0) {
    // every tile is one sequential chain: with tiles <= resident CTAs a single segment per tile (no hand-off) is optimal;
    // with more tiles than CTAs, time segments (>= 128 steps) let the persistent CTAs share the tiles evenly
    if (tiles <= G_max) seg_chunks = n_chunks;
    else {
      const int want_segs = (int)i64max(1, cdiv((int64_t)4 * G_max, tiles));
      seg_chunks = (int)i64max(cdiv(128, cs), n_chunks / want_segs);
    }
  }
  if (seg_chunks > n_chunks) seg_chunks = n_chunks;
  const int n_segs = (int)cdiv(n_chunks, seg_chunks);
  const int64_t n_items64 = (int64_t)n_segs * tiles;
  if (n_items64 >= ((int64_t)1 << 30)) return CRUX_OK;

  TmaArgs a;
  memset(&a, 0, sizeof(a));
  a.T = T; a.N = N; a.gamma = gamma; a.lambda = lambda; a.has_adv = adv != nullptr; a.has_ret = ret != nullptr;
  a.n_chunks = n_chunks; a.seg_chunks = seg_chunks; a.n_segs = n_segs; a.tiles = tiles; a.n_items = (int)n_items64; a.stages = stages;
  // scratch slot 7 belongs to this path and survives between launches: [counter + flags: flag_cap words][carry records].
  // The flag region keeps a fixed capacity so that carry bits of one launch can never be read as flags by a later one.
  size_t flag_cap = ctx->gae_flag_cap ? ctx->gae_flag_cap : 4096;
  while (flag_cap < (size_t)n_items64 * ncw + 4) flag_cap *= 2;
  const size_t flag_al = flag_cap * sizeof(unsigned int);
  const size_t carry_bytes = (size_t)n_items64 * ncw * 32 * sizeof(float4);
  char *p = (char *)crux_scratch(ctx, 7, flag_al + carry_bytes);
  if (!p) return CRUX_ERR_OOM;
  if (p != ctx->gae_scratch_seen || flag_cap != ctx->gae_flag_cap || ctx->gae_epoch == 0xFFFFFFFFu) {   // fresh / regrown block or epoch wrap
    CRUX_CHECK_CUDA(ctx, cudaMemsetAsync(p, 0, flag_al, ctx->stream));
    ctx->gae_scratch_seen = p; ctx->gae_flag_cap = flag_cap; ctx->gae_epoch = 0; ctx->gae_ctr_base = 0;
  }
  a.counter = (unsigned int *)p;
  a.flags = (unsigned int *)p + 4;
  a.carry = (float4 *)(p + flag_al);
  a.err_flags = ctx->flags_dev;
  a.epoch = ++ctx->gae_epoch;
  a.counter_base = ctx->gae_ctr_base;

  CUtensorMap m[7];
  const void *cols[7] = {r, vs, vsp, done, ee, adv ? adv : ret, ret ? ret : adv};
  for (int i = 0; i < 7; ++i)
    if (!make_map(&m[i], cols[i], i == 3 || i == 4, T, N, cs, TS)) return CRUX_OK;   // not encodable (exotic pitch): fall back to the scan kernel

  const size_t smem = (size_t)stages * stage_bytes;
  const int grid = (int)i64min(n_items64, (int64_t)G_max);
  {
    CruxTimed timed(ctx, CRUX_T_GAE);
#define GAE_TMA_LAUNCH(CS_, NCW_, SPL_)                                                                                                        \
  do {                                                                                                                                         \
    CRUX_CHECK_CUDA(ctx, cudaFuncSetAttribute(gae_tma_kernel<CS_, NCW_, SPL_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    gae_tma_kernel<CS_, NCW_, SPL_><<<grid, n_threads(NCW_), smem, ctx->stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], a);                \
  } while (0)
    if (spl == 1)      { if (cs == 32) GAE_TMA_LAUNCH(32, 1, 1); else GAE_TMA_LAUNCH(64, 1, 1); }
    else if (ncw == 2) { if (cs == 16) GAE_TMA_LAUNCH(16, 2, 2); else GAE_TMA_LAUNCH(32, 2, 2); }
    else               { if (cs == 16) GAE_TMA_LAUNCH(16, 1, 2); else GAE_TMA_LAUNCH(32, 1, 2); }
#undef GAE_TMA_LAUNCH
  }
  CRUX_LAUNCHED(ctx);
  ctx->gae_ctr_base += (unsigned int)n_items64 + (unsigned int)grid;   // every CTA takes exactly one index past the end
  *handled = 1;
  return CRUX_OK;
}
