// Channels-last GroupNorm(+SiLU) (K6 of the hot path; reference ops: pnp_utils.py:48-49, 92, 104 and every other
// GroupNorm -> SiLU pair of the UNet / VAE).  HBM-bound: 4 B per element (read x once, write y once).
//
//   x, y : [n_samples][rows][C] fp16;  statistics per (sample, group) over rows x (C / groups) elements
//
// ONE persistent kernel, one CTA per SM, built around the 126 MB L2 (round 1 ran two kernels = three HBM passes, 6 B/element):
//   * the samples are cut into L2-sized CHUNKS (<= kChunkBytes of x).  For each chunk the whole grid first streams the chunk
//     once for the statistics (phase A: HBM -> L2 -> smem), meets at a grid-wide barrier, and then streams the SAME rows again
//     for the normalisation (phase B): that second read hits in L2, so x crosses HBM once.  A clip-level sample of the 64 x 64
//     level (65 536 rows x 320 channels = 42 MB) is one chunk; the per-frame norms (48 samples of 2.6 MB) go 16 frames at a time.
//     Samples larger than L2 (128-frame clips) still work — phase B then walks the chunk back to front, so its first reads
//     hit the part of x that phase A touched last.
//   * rows are contiguous in the channels-last layout, so a slice of rows is ONE byte range: a producer warp moves it with
//     1-D bulk TMA copies (cp.async.bulk, kStages x ~20 KB in flight per SM, completion on mbarriers); the consumer threads
//     only ever read shared memory.  No per-thread global loads, no address arithmetic in the inner loop.
//   * deterministic: per-(sample, slice, group) partial sums in a fixed order, folded in double; no float atomics.
//     The grid barrier is one integer atomic per CTA and chunk.
#include "host_util.cuh"
#include "ptx.cuh"

namespace av2v {
namespace {

constexpr int kGnMaxSlices = 512;   // slices per sample (partial-sum slots)
constexpr int kGnMaxGroups = 64;
constexpr int kStages = 4;
constexpr int kStageBytes = 24 * 1024;                 // upper bound of one stage (rows_per_stage * C * 2 <= this)
constexpr long long kChunkBytes = 48ll << 20;          // x + y of a chunk (2 x 48 MB) stay inside the 126 MB L2
constexpr int kMaxThreads = 512;                      // consumer threads (+ one producer warp)

// grid barrier state: [0] arrivals of the running launch, [1] exits.  Zero at module load; the last CTA to exit resets both,
// so consecutive (stream-ordered) launches start from zero.  One process drives one GPU, launches are stream-ordered.
__device__ unsigned int g_gn_sync[2];

struct GnParams {
  const __half* x;
  __half* y;
  const __half* gamma;
  const __half* beta;
  float* partial;  // [n][slices][groups][2]
  int n, rows, C, groups, cpg, vpr;
  int rp;              // row lanes: consumer threads = vpr * rp
  int k;               // rows per thread and stage
  int stage_rows;      // rp * k
  int slices;          // slices per sample
  int rows_per_slice;
  int chunk_samples, n_chunks;
  float eps;
  int silu;
};

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

// Work items of a chunk: (sample, slice) pairs, item j = s_local * slices + slice; CTA b takes j = b, b + G, ...
struct ItemIter {
  int chunk_first_sample, items, j;
};

__global__ void __launch_bounds__(kMaxThreads + 32, 1)
gn_persistent_kernel(const GnParams p) {
  extern __shared__ __align__(128) uint8_t gsm[];
  uint8_t* stage_buf = gsm;                                                      // [kStages][kStageBytes]
  float* red = reinterpret_cast<float*>(gsm + kStages * kStageBytes);            // [rp][C][2] fold scratch, then [groups][2]
  const int T = p.vpr * p.rp;                                                    // consumer threads (multiple of 32)
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + static_cast<size_t>(p.rp) * p.C * 2 + 2 * kGnMaxGroups);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  const int t = threadIdx.x;
  const int G = gridDim.x;
  const bool is_producer = t >= T;

  if (t == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], T / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const long long row_bytes = static_cast<long long>(p.C) * 2;
  // per item: rows [rbeg, rend) of sample s
  auto item_rows = [&](int slice, int& rbeg, int& rend) {
    rbeg = slice * p.rows_per_slice;
    rend = min(p.rows, rbeg + p.rows_per_slice);
  };

  if (is_producer) {
    // ================================================================== producer warp: the same stage sequence as the consumers
    if (t == T) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        const int s0 = ch * p.chunk_samples;
        const int ns = min(p.chunk_samples, p.n - s0);
        const int items = ns * p.slices;
        for (int pass = 0; pass < 2; ++pass) {
          // phase A: items ascending; phase B: items descending (what phase A read last is what L2 still holds for sure)
          const int my = (items - static_cast<int>(blockIdx.x) + G - 1) / G;  // number of items of this CTA (may be <= 0)
          for (int ii = 0; ii < my; ++ii) {
            const int j = static_cast<int>(blockIdx.x) + (pass == 0 ? ii : my - 1 - ii) * G;
            const int s = s0 + j / p.slices, slice = j % p.slices;
            int rbeg, rend;
            item_rows(slice, rbeg, rend);
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.x) + (static_cast<long long>(s) * p.rows + rbeg) * row_bytes;
            for (int r = rbeg; r < rend; r += p.stage_rows) {
              const int nr = min(p.stage_rows, rend - r);
              mbar_wait(&empty[stage], phase ^ 1u);
              const uint32_t bytes = static_cast<uint32_t>(nr * row_bytes);
              mbar_arrive_expect_tx(&full[stage], bytes);
              bulk_load_1d(smem_u32(stage_buf + stage * kStageBytes), src, bytes, smem_u32(&full[stage]));
              src += bytes;
              if (++stage == kStages) {
                stage = 0;
                phase ^= 1u;
              }
            }
          }
        }
      }
    }
    return;
  }

  // ==================================================================== consumers: thread = (8-channel vector v, row lane r0)
  const int v = t % p.vpr, r0 = t / p.vpr;
  const int lane = t & 31;
  int stage = 0;
  uint32_t phase = 0;
  unsigned int barrier_no = 0;
  auto consumer_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"r"(T) : "memory"); };
  auto release_stage = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++stage == kStages) {
      stage = 0;
      phase ^= 1u;
    }
  };

  for (int ch = 0; ch < p.n_chunks; ++ch) {
    const int s0 = ch * p.chunk_samples;
    const int ns = min(p.chunk_samples, p.n - s0);
    const int items = ns * p.slices;
    const int my = (items - static_cast<int>(blockIdx.x) + G - 1) / G;

    // ------------------------------------------------------------------ phase A: statistics
    for (int ii = 0; ii < my; ++ii) {
      const int j = static_cast<int>(blockIdx.x) + ii * G;
      const int s = s0 + j / p.slices, slice = j % p.slices;
      int rbeg, rend;
      item_rows(slice, rbeg, rend);
      float sm_[8], sq_[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) sm_[e] = sq_[e] = 0.f;
      for (int r = rbeg; r < rend; r += p.stage_rows) {
        const int nr = min(p.stage_rows, rend - r);
        mbar_wait(&full[stage], phase);
        const uint4* sb = reinterpret_cast<const uint4*>(stage_buf + stage * kStageBytes) + v;
        for (int i0 = 0; i0 < p.k; i0 += 4) {  // up to four shared-memory vectors in flight per thread
          uint4 a4[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int rr = r0 + (i0 + u) * p.rp;
            ok[u] = (i0 + u < p.k) && (rr < nr);
            a4[u] = ok[u] ? sb[rr * p.vpr] : make_uint4(0u, 0u, 0u, 0u);  // zeros add nothing to either sum
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const __half2* ah = reinterpret_cast<const __half2*>(&a4[u]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __half22float2(ah[e]);
              sm_[2 * e] += f.x;
              sm_[2 * e + 1] += f.y;
              sq_[2 * e] = fmaf(f.x, f.x, sq_[2 * e]);
              sq_[2 * e + 1] = fmaf(f.y, f.y, sq_[2 * e + 1]);
            }
          }
        }
        release_stage();
      }
      // fold: row lanes -> channel totals -> group totals (fixed order), one partial per (sample, slice, group)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[(r0 * p.C + v * 8 + e) * 2] = sm_[e];
        red[(r0 * p.C + v * 8 + e) * 2 + 1] = sq_[e];
      }
      consumer_sync();
      for (int c = t; c < p.C; c += T) {
        float ss = 0.f, qq = 0.f;
        for (int kk = 0; kk < p.rp; ++kk) {
          ss += red[(kk * p.C + c) * 2];
          qq += red[(kk * p.C + c) * 2 + 1];
        }
        red[c * 2] = ss;  // row-lane 0 now holds the channel totals (each thread only overwrites what it alone read at kk = 0)
        red[c * 2 + 1] = qq;
      }
      consumer_sync();
      if (t < p.groups) {
        float ss = 0.f, qq = 0.f;
        for (int c = t * p.cpg; c < (t + 1) * p.cpg; ++c) {
          ss += red[c * 2];
          qq += red[c * 2 + 1];
        }
        float* dst = p.partial + ((static_cast<long long>(s) * p.slices + slice) * p.groups + t) * 2;
        __stcg(dst, ss);
        __stcg(dst + 1, qq);
      }
      consumer_sync();  // `red` is reused by the next item
    }

    // ------------------------------------------------------------------ grid barrier: every partial of the chunk is written
    ++barrier_no;
    consumer_sync();
    if (t == 0) {
      __threadfence();
      atomicAdd(&g_gn_sync[0], 1u);
      const unsigned int target = barrier_no * static_cast<unsigned int>(G);
      const long long t0 = clock64();
      while (ld_acquire_gpu(&g_gn_sync[0]) < target) {
        __nanosleep(64);
        if (clock64() - t0 > AV2V_WAIT_TIMEOUT_CYCLES) {
          printf("av2v: groupnorm grid barrier timeout (block %d, barrier %u)\n", blockIdx.x, barrier_no);
          __trap();
        }
      }
      __threadfence();
    }
    consumer_sync();

    // ------------------------------------------------------------------ phase B: normalise (+SiLU), items descending
    float* stat = red + static_cast<size_t>(p.rp) * p.C * 2;  // [groups][2] = mean, rstd of the current sample
    int cur_sample = -1;
    float a[8], b[8];
    for (int ii = my - 1; ii >= 0; --ii) {
      const int j = static_cast<int>(blockIdx.x) + ii * G;
      const int s = s0 + j / p.slices, slice = j % p.slices;
      int rbeg, rend;
      item_rows(slice, rbeg, rend);
      if (s != cur_sample) {
        cur_sample = s;
        consumer_sync();  // everyone is done with the previous sample's `stat`
        // one warp per group (round robin): lanes stride over the slices, fixed-order shuffle tree, all in double
        const int warp = t >> 5, nwarps = T >> 5;
        for (int g = warp; g < p.groups; g += nwarps) {
          double ss = 0.0, qq = 0.0;
          for (int sl = lane; sl < p.slices; sl += 32) {
            const float* src = p.partial + ((static_cast<long long>(s) * p.slices + sl) * p.groups + g) * 2;
            ss += static_cast<double>(__ldcg(src));
            qq += static_cast<double>(__ldcg(src + 1));
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
            qq += __shfl_xor_sync(0xffffffffu, qq, o);
          }
          if (lane == 0) {
            const double cnt = static_cast<double>(p.rows) * p.cpg;
            const double mean = ss / cnt;
            double var = qq / cnt - mean * mean;
            if (var < 0.0) var = 0.0;
            stat[2 * g] = static_cast<float>(mean);
            stat[2 * g + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.eps)));
          }
        }
        consumer_sync();
        const uint4 gv = __ldg(reinterpret_cast<const uint4*>(p.gamma) + v);
        const uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.beta) + v);
        const __half* gh = reinterpret_cast<const __half*>(&gv);
        const __half* bh = reinterpret_cast<const __half*>(&bv);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int g = (v * 8 + e) / p.cpg;
          const float mean = stat[2 * g], rstd = stat[2 * g + 1];
          a[e] = rstd * __half2float(gh[e]);
          b[e] = __half2float(bh[e]) - mean * a[e];
        }
      }
      uint4* dst_base = reinterpret_cast<uint4*>(p.y) + (static_cast<long long>(s) * p.rows) * p.vpr + v;
      for (int r = rbeg; r < rend; r += p.stage_rows) {
        const int nr = min(p.stage_rows, rend - r);
        mbar_wait(&full[stage], phase);
        const uint4* sb = reinterpret_cast<const uint4*>(stage_buf + stage * kStageBytes) + v;
#pragma unroll 2
        for (int i = 0; i < p.k; ++i) {
          const int rr = r0 + i * p.rp;
          if (rr < nr) {
            const uint4 xv = sb[rr * p.vpr];
            const __half* xh = reinterpret_cast<const __half*>(&xv);
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = fmaf(__half2float(xh[e]), a[e], b[e]);
            if (p.silu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                // the reference rounds the GroupNorm output to fp16 before SiLU (two separate ops); silu(f) = f / (1 + 2^(-f log2 e))
                const float g = r16(f[e]);
                float rc;
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(1.0f + ex2_approx(g * -1.4426950408889634f)));
                f[e] = g * rc;
              }
            }
            uint4 ov;
            ov.x = pack_half2(f[0], f[1]);
            ov.y = pack_half2(f[2], f[3]);
            ov.z = pack_half2(f[4], f[5]);
            ov.w = pack_half2(f[6], f[7]);
            dst_base[static_cast<long long>(r + rr) * p.vpr] = ov;
          }
        }
        release_stage();
      }
    }
  }

  // last CTA out resets the barrier state for the next launch
  consumer_sync();
  if (t == 0) {
    __threadfence();
    const unsigned int old = atomicAdd(&g_gn_sync[1], 1u);
    if (old == static_cast<unsigned int>(G) - 1u) {
      g_gn_sync[0] = 0u;
      g_gn_sync[1] = 0u;
      __threadfence();
    }
  }
}

int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

}  // namespace
}  // namespace av2v

using namespace av2v;

extern "C" int av2v_groupnorm_workspace_floats(int n_samples, int C) {
  (void)C;  // partial sums are kept per (sample, slice, group): independent of the channel count
  return n_samples * kGnMaxSlices * kGnMaxGroups * 2;
}

extern "C" int av2v_groupnorm_silu_f16(const av2v_groupnorm_args* a, av2v_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AV2V_REQUIRE(a != nullptr, AV2V_EINVAL, "groupnorm: null args");
  AV2V_REQUIRE(a->x && a->y && a->gamma && a->beta && a->workspace, AV2V_EINVAL, "groupnorm: null pointer");
  AV2V_REQUIRE(a->n_samples > 0 && a->rows > 0 && a->C > 0 && a->groups > 0, AV2V_EINVAL, "groupnorm: bad shape");
  AV2V_REQUIRE(a->C % a->groups == 0, AV2V_EINVAL, "groupnorm: C %% groups != 0");
  AV2V_REQUIRE(a->C % 8 == 0 && a->C <= 8192, AV2V_ENOSUP, "groupnorm: C must be a multiple of 8 and <= 8192");
  AV2V_REQUIRE(a->groups <= kGnMaxGroups, AV2V_ENOSUP, "groupnorm: at most 64 groups");
  AV2V_REQUIRE(aligned16(a->x) && aligned16(a->y) && aligned16(a->gamma) && aligned16(a->beta), AV2V_EALIGN,
               "groupnorm: pointers must be 16-byte aligned");

  GnParams p{};
  p.x = static_cast<const __half*>(a->x);
  p.y = static_cast<__half*>(a->y);
  p.gamma = static_cast<const __half*>(a->gamma);
  p.beta = static_cast<const __half*>(a->beta);
  p.partial = a->workspace;
  p.n = a->n_samples;
  p.rows = a->rows;
  p.C = a->C;
  p.groups = a->groups;
  p.cpg = a->C / a->groups;
  p.vpr = a->C / 8;
  p.eps = a->eps;
  p.silu = a->silu;
  // consumer threads: vpr * rp, a multiple of 32, 384 ... 512 where the width allows it
  const int rp0 = 32 / gcd_int(p.vpr, 32);
  int sets = (384 + p.vpr * rp0 - 1) / (p.vpr * rp0);
  while (sets > 1 && p.vpr * rp0 * sets > kMaxThreads) --sets;
  const int rp = rp0 * sets;
  AV2V_REQUIRE(p.vpr * rp <= kMaxThreads, AV2V_ENOSUP, "groupnorm: C = %d needs %d threads (max %d)", a->C, p.vpr * rp, kMaxThreads);
  p.rp = rp;
  const long long row_bytes = static_cast<long long>(a->C) * 2;
  int k = static_cast<int>(kStageBytes / (row_bytes * rp));
  AV2V_REQUIRE(k >= 1, AV2V_ENOSUP, "groupnorm: C = %d: one row lane set (%lld B) exceeds a stage", a->C, row_bytes * rp);
  if (k > 8) k = 8;
  p.k = k;
  p.stage_rows = rp * k;

  // chunks of whole samples with at most kChunkBytes of x
  const long long sample_bytes = static_cast<long long>(a->rows) * row_bytes;
  long long cs = kChunkBytes / sample_bytes;
  if (cs < 1) cs = 1;
  if (cs > a->n_samples) cs = a->n_samples;
  // spread the samples evenly over the chunks (48 frames of 2.6 MB: 3 x 16, not 18 + 18 + 12)
  const int n_chunks = static_cast<int>((a->n_samples + cs - 1) / cs);
  p.chunk_samples = (a->n_samples + n_chunks - 1) / n_chunks;
  p.n_chunks = (a->n_samples + p.chunk_samples - 1) / p.chunk_samples;

  // slices per sample: ~2 items per CTA and chunk when an item then still has >= 2 stages, else 1 item per CTA
  const int sms = sm_count_cached();
  const long long chunk_rows = static_cast<long long>(p.chunk_samples) * a->rows;
  const long long max_items = (chunk_rows + p.stage_rows - 1) / p.stage_rows;  // >= one stage per item
  long long want_items = 2ll * sms;
  if (max_items < 4ll * sms) want_items = sms;
  if (want_items > max_items) want_items = max_items;
  if (want_items < 1) want_items = 1;
  int slices = static_cast<int>((want_items + p.chunk_samples - 1) / p.chunk_samples);
  if (slices > kGnMaxSlices) slices = kGnMaxSlices;
  const int max_slices_by_rows = (a->rows + p.stage_rows - 1) / p.stage_rows;
  if (slices > max_slices_by_rows) slices = max_slices_by_rows;
  if (slices < 1) slices = 1;
  p.rows_per_slice = (a->rows + slices - 1) / slices;
  p.slices = (a->rows + p.rows_per_slice - 1) / p.rows_per_slice;  // no empty slices

  const long long items = static_cast<long long>(p.chunk_samples) * p.slices;
  const int grid = static_cast<int>(items < sms ? items : sms);  // <= one CTA per SM: all CTAs are co-resident (grid barrier)
  const size_t smem = static_cast<size_t>(kStages) * kStageBytes + (static_cast<size_t>(rp) * a->C * 2 + 2 * kGnMaxGroups) * sizeof(float) +
                      2 * kStages * sizeof(uint64_t) + 128;
  AV2V_REQUIRE(smem <= 227 * 1024, AV2V_ENOSUP, "groupnorm: shared memory budget exceeded (%zu B)", smem);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    AV2V_CHECK_CUDA(cudaFuncSetAttribute(gn_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_smem = 227 * 1024;
  }
  gn_persistent_kernel<<<grid, p.vpr * rp + 32, smem, stream>>>(p);
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}
