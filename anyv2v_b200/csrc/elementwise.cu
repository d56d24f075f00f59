// HBM-bound kernels of the hot path: fused CFG + DDIM step (K7) and channels-last GroupNorm(+SiLU) (K6).
#include "host_util.cuh"
#include "ptx.cuh"

namespace av2v {
namespace {

// ---------------------------------------------------------------------------------------------------- K7
// Every arithmetic result is rounded to fp16 separately, in the order the reference's chain of PyTorch ops
// produces them (pipeline_i2vgen_xl.py:1162, consisti2v/ddim_inverse_scheduler.py:346-369): fp32 multiply by the
// fp32 scalar, round; fp32 add of two fp16 values, round.  __fmul_rn/__fadd_rn forbid FMA contraction.
__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ float ddim_one(float x, float vn, float ve, bool cfg, float g, float ca, float cb,
                                          float cc, float cd) {
  float v = vn;
  if (cfg) {
    const float d0 = r16(__fsub_rn(ve, vn));
    const float d1 = r16(__fmul_rn(g, d0));
    v = r16(__fadd_rn(vn, d1));
  }
  const float x0 = r16(__fsub_rn(r16(__fmul_rn(ca, x)), r16(__fmul_rn(cb, v))));
  const float ep = r16(__fadd_rn(r16(__fmul_rn(ca, v)), r16(__fmul_rn(cb, x))));
  const float dir = r16(__fmul_rn(cd, ep));
  return r16(__fadd_rn(r16(__fmul_rn(cc, x0)), dir));
}

__global__ void __launch_bounds__(256)
ddim_step_kernel(const __half* __restrict__ x, const __half* __restrict__ vn, const __half* __restrict__ ve,
                 __half* __restrict__ out, long long n, float g, float ca, float cb, float cc, float cd,
                 const float* __restrict__ coef_dev, int pdl) {
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  if (coef_dev != nullptr) {
    ca = coef_dev[0];
    cb = coef_dev[1];
    cc = coef_dev[2];
    cd = coef_dev[3];
    g = coef_dev[4];
  }
  const bool cfg = ve != nullptr;
  const long long nvec = n >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 xv = reinterpret_cast<const uint4*>(x)[i];
    const uint4 nv = reinterpret_cast<const uint4*>(vn)[i];
    uint4 ev = nv;
    if (cfg) ev = reinterpret_cast<const uint4*>(ve)[i];
    const __half* xh = reinterpret_cast<const __half*>(&xv);
    const __half* nh = reinterpret_cast<const __half*>(&nv);
    const __half* eh = reinterpret_cast<const __half*>(&ev);
    uint4 ov;
    __half* oh = reinterpret_cast<__half*>(&ov);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      oh[e] = __float2half_rn(ddim_one(__half2float(xh[e]), __half2float(nh[e]), __half2float(eh[e]), cfg, g, ca,
                                       cb, cc, cd));
    reinterpret_cast<uint4*>(out)[i] = ov;
  }
  // tail (n not a multiple of 8)
  const long long tail0 = nvec << 3;
  for (long long i = tail0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2half_rn(ddim_one(__half2float(x[i]), __half2float(vn[i]), cfg ? __half2float(ve[i]) : 0.f, cfg,
                                      g, ca, cb, cc, cd));
}

int ddim_launch(const av2v_ddim_args* a, cudaStream_t stream) {
  AV2V_REQUIRE(a != nullptr, AV2V_EINVAL, "ddim: null args");
  AV2V_REQUIRE(a->n >= 0, AV2V_EINVAL, "ddim: negative element count");
  if (a->n == 0) return AV2V_OK;  // empty latents: nothing to do (pointers may be null)
  AV2V_REQUIRE(a->x && a->v_neg && a->out, AV2V_EINVAL, "ddim: null x / v_neg / out");
  AV2V_REQUIRE(aligned16(a->x) && aligned16(a->v_neg) && aligned16(a->out) && (!a->v_edit || aligned16(a->v_edit)),
               AV2V_EALIGN, "ddim: pointers must be 16-byte aligned");
  const long long nvec = (a->n + 7) >> 3;
  long long blocks = (nvec + 255) / 256;
  const long long cap = static_cast<long long>(sm_count_cached()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const int pdl = pdl_enabled();
  if (pdl)
    AV2V_CHECK_CUDA(launch_ex(ddim_step_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, 1, 1,
                              static_cast<const __half*>(a->x), static_cast<const __half*>(a->v_neg),
                              static_cast<const __half*>(a->v_edit), static_cast<__half*>(a->out), a->n, a->guidance, a->ca,
                              a->cb, a->cc, a->cd, a->coef_dev, 1));
  else
    ddim_step_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const __half*>(a->x), static_cast<const __half*>(a->v_neg), static_cast<const __half*>(a->v_edit),
        static_cast<__half*>(a->out), a->n, a->guidance, a->ca, a->cb, a->cc, a->cd, a->coef_dev, 0);
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}

// ---------------------------------------------------------------------------------------------------- K6
// Channels-last GroupNorm: x[n][row][C].  Thread t owns a fixed 8-channel vector column v = t % VPR (VPR = C/8)
// and walks rows r = t / VPR, += rows_par.  Pass 1 writes per-(sample, slice, channel) partial sum / sum-of-
// squares (deterministic, no atomics); pass 2 folds them per group in double, then streams x -> y.
constexpr int kGnMaxSlices = 256;
constexpr int kGnMaxGroups = 64;
constexpr int kGnFoldParts = 16;

// A batch of U 16-byte read-only loads.  U = 4 (shipped): plain __ldg — cuobjdump shows that the compiler interleaves
// each load with the arithmetic on the previous one (LDG, use, LDG, use ...), i.e. about ONE load in flight per thread,
// which is why the statistics pass sits at 35-50 % of the HBM roofline.  U = 8 (AV2V_GN_V2): all addresses are formed
// first and the eight loads are issued from ONE asm statement, so no use can be scheduled between them.
template <int U>
__device__ __forceinline__ void gn_load_batch(const __half* p0, long long step, uint4 (&a)[U]) {
  if constexpr (U == 8) {
    const __half* q1 = p0 + step;
    const __half* q2 = q1 + step;
    const __half* q3 = q2 + step;
    const __half* q4 = q3 + step;
    const __half* q5 = q4 + step;
    const __half* q6 = q5 + step;
    const __half* q7 = q6 + step;
    asm volatile(
        "ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%32];\n"
        "ld.global.nc.v4.u32 {%4, %5, %6, %7}, [%33];\n"
        "ld.global.nc.v4.u32 {%8, %9, %10, %11}, [%34];\n"
        "ld.global.nc.v4.u32 {%12, %13, %14, %15}, [%35];\n"
        "ld.global.nc.v4.u32 {%16, %17, %18, %19}, [%36];\n"
        "ld.global.nc.v4.u32 {%20, %21, %22, %23}, [%37];\n"
        "ld.global.nc.v4.u32 {%24, %25, %26, %27}, [%38];\n"
        "ld.global.nc.v4.u32 {%28, %29, %30, %31}, [%39];\n"
        : "=r"(a[0].x), "=r"(a[0].y), "=r"(a[0].z), "=r"(a[0].w), "=r"(a[1].x), "=r"(a[1].y), "=r"(a[1].z), "=r"(a[1].w),
          "=r"(a[2].x), "=r"(a[2].y), "=r"(a[2].z), "=r"(a[2].w), "=r"(a[3].x), "=r"(a[3].y), "=r"(a[3].z), "=r"(a[3].w),
          "=r"(a[4].x), "=r"(a[4].y), "=r"(a[4].z), "=r"(a[4].w), "=r"(a[5].x), "=r"(a[5].y), "=r"(a[5].z), "=r"(a[5].w),
          "=r"(a[6].x), "=r"(a[6].y), "=r"(a[6].z), "=r"(a[6].w), "=r"(a[7].x), "=r"(a[7].y), "=r"(a[7].z), "=r"(a[7].w)
        : "l"(p0), "l"(q1), "l"(q2), "l"(q3), "l"(q4), "l"(q5), "l"(q6), "l"(q7));
  } else {
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = __ldg(reinterpret_cast<const uint4*>(p0 + u * step));
  }
}

template <int U>  // U loads per loop iteration (4 = shipped).  NOTE (cuobjdump, round 1 end): ptxas does NOT keep them in flight
                   // together — see gn_stats_async_kernel below for the cp.async version (AV2V_GN_V2=1)
__global__ void gn_stats_kernel(const __half* __restrict__ x, float* __restrict__ partial, int rows, int C,
                                int groups, int vpr, int rows_par, int slices, int pdl, int rev) {
  extern __shared__ float sm[];  // [rows_par][C][2]
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  // rev (AV2V_PINGPONG): samples / slices walked back to front; the partial sums are indexed by (n, slice): same result
  const int n = rev ? static_cast<int>(gridDim.y) - 1 - static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.y);
  const int slice = rev ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
  const int t = threadIdx.x;
  const int v = t % vpr, r0 = t / vpr;
  const int rows_per_slice = (rows + slices - 1) / slices;
  const int rbeg = slice * rows_per_slice;
  const int rend = min(rows, rbeg + rows_per_slice);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  const __half* base = x + (static_cast<long long>(n) * rows) * C + v * 8;
  if (r0 < rows_par) {
    int r = rbeg + r0;
    // U independent 16-byte loads in flight per thread (memory-level parallelism); rows are accumulated in the same order
    // for every U, so the statistics do not depend on it
    for (; r + (U - 1) * rows_par < rend; r += U * rows_par) {
      uint4 a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) a[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r + u * rows_par) * C));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const __half2* ah = reinterpret_cast<const __half2*>(&a[u]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 fa = __half22float2(ah[e]);
          s[2 * e] += fa.x;
          s[2 * e + 1] += fa.y;
          q[2 * e] += fa.x * fa.x;
          q[2 * e + 1] += fa.y * fa.y;
        }
      }
    }
    for (; r < rend; r += rows_par) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * C));
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = __half22float2(ah[e]);
        s[2 * e] += fa.x;
        s[2 * e + 1] += fa.y;
        q[2 * e] += fa.x * fa.x;
        q[2 * e + 1] += fa.y * fa.y;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sm[(r0 * C + v * 8 + e) * 2] = s[e];
      sm[(r0 * C + v * 8 + e) * 2 + 1] = q[e];
    }
  }
  __syncthreads();
  // fold the rows_par partials per channel (fixed order) ...
  for (int c = t; c < C; c += blockDim.x) {
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < rows_par; ++k) {
      ss += sm[(k * C + c) * 2];
      qq += sm[(k * C + c) * 2 + 1];
    }
    sm[c * 2] = ss;  // row 0 of the staging buffer now holds the per-channel totals of this slice
    sm[c * 2 + 1] = qq;
  }
  __syncthreads();
  // ... then the channels of each group (fixed order), write [n][slice][group][2]
  const int cpg = C / groups;
  if (t < groups) {
    float ss = 0.f, qq = 0.f;
    for (int c = t * cpg; c < (t + 1) * cpg; ++c) {
      ss += sm[c * 2];
      qq += sm[c * 2 + 1];
    }
    float* dst = partial + ((static_cast<long long>(n) * slices + slice) * groups + t) * 2;
    dst[0] = ss;
    dst[1] = qq;
  }
}

// Statistics pass, round-2 candidate (AV2V_GN_V2=1).  Same partial sums in the same order as gn_stats_kernel (bit-identical
// statistics), but the rows travel global -> shared through cp.async (LDGSTS): two stages of four 16-byte copies per
// thread are in flight regardless of how ptxas schedules the arithmetic.  (With plain loads ptxas puts the FADD / FFMA on
// load k between loads k+1 and k+2 — even when the loads come from one asm statement or are fenced with a warp barrier —
// so only ~2 x 16 B per thread are in flight and the pass runs at 35-50 % of the HBM roofline.)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kGnAsyncU = 4;       // copies per thread and stage
constexpr int kGnAsyncStages = 2;  // stages in flight

__global__ void gn_stats_async_kernel(const __half* __restrict__ x, float* __restrict__ partial, int rows, int C,
                                      int groups, int vpr, int rows_par, int slices, int pdl, int rev) {
  extern __shared__ float sm[];  // max([rows_par][C][2] floats, [stages][U][threads] uint4): staging first, then reduction
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  constexpr int U = kGnAsyncU;
  // REVERSED traversal (last sample / last slice first): the producer wrote x front to back, so for a tensor about the size
  // of L2 (126 MB at B = 3 on the 64 x 64 level) the TAIL is what is still resident.  Reading front to back would miss on the
  // head and, under LRU, evict the tail before it is reached; back to front hits on the tail and leaves the HEAD in L2 for the
  // apply pass, which walks forward.  The partial sums are indexed by (n, slice), so the result does not change.
  // (rev = 1 is this kernel's default; with AV2V_PINGPONG the host alternates it from launch to launch)
  const int n = rev ? static_cast<int>(gridDim.y) - 1 - static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.y);
  const int slice = rev ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
  const int t = threadIdx.x;
  const int v = t % vpr, r0 = t / vpr;
  const int rows_per_slice = (rows + slices - 1) / slices;
  const int rbeg = slice * rows_per_slice;
  const int rend = min(rows, rbeg + rows_per_slice);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  const __half* base = x + (static_cast<long long>(n) * rows) * C + v * 8;
  uint4* stage = reinterpret_cast<uint4*>(sm);
  auto accumulate = [&](const uint4& a) {
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 fa = __half22float2(ah[e]);
      s[2 * e] += fa.x;
      s[2 * e + 1] += fa.y;
      q[2 * e] += fa.x * fa.x;
      q[2 * e + 1] += fa.y * fa.y;
    }
  };
  if (r0 < rows_par) {
    // this thread's rows: first + k * rows_par, k = 0 .. cnt-1: `full` un-predicated batches of U through cp.async, then
    // the < U remaining rows with plain loads (same order of accumulation as gn_stats_kernel)
    const int first = rbeg + r0;
    const int cnt = first < rend ? (rend - first + rows_par - 1) / rows_par : 0;
    const int full = cnt / U;
    const long long step = static_cast<long long>(rows_par) * C;
    const __half* src = base + static_cast<long long>(first) * C;  // next row to copy
    uint4* const my = stage + t;
    const int sstride = blockDim.x;  // uint4 elements between two slots of this thread
    auto issue = [&](int st) {
#pragma unroll
      for (int u = 0; u < U; ++u) cp_async16(my + (st * U + u) * sstride, src + u * step);
      src += U * step;
    };
    if (full > 0) issue(0);
    cp_async_commit();
    for (int b = 0; b < full; ++b) {
      if (b + 1 < full) issue((b + 1) % kGnAsyncStages);
      cp_async_commit();   // one group per iteration, possibly empty: keeps the wait_group arithmetic uniform
      cp_async_wait<1>();  // batch b has landed (batch b + 1 may still be in flight)
      const uint4* got = my + (b % kGnAsyncStages) * U * sstride;
#pragma unroll
      for (int u = 0; u < U; ++u) accumulate(got[u * sstride]);
    }
    cp_async_wait<0>();
    for (int k = full * U; k < cnt; ++k, src += step) accumulate(__ldg(reinterpret_cast<const uint4*>(src)));
  }
  __syncthreads();  // the staging area becomes the reduction buffer
  if (r0 < rows_par) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      sm[(r0 * C + v * 8 + e) * 2] = s[e];
      sm[(r0 * C + v * 8 + e) * 2 + 1] = q[e];
    }
  }
  __syncthreads();
  for (int c = t; c < C; c += blockDim.x) {
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < rows_par; ++k) {
      ss += sm[(k * C + c) * 2];
      qq += sm[(k * C + c) * 2 + 1];
    }
    sm[c * 2] = ss;
    sm[c * 2 + 1] = qq;
  }
  __syncthreads();
  const int cpg = C / groups;
  if (t < groups) {
    float ss = 0.f, qq = 0.f;
    for (int c = t * cpg; c < (t + 1) * cpg; ++c) {
      ss += sm[c * 2];
      qq += sm[c * 2 + 1];
    }
    float* dst = partial + ((static_cast<long long>(n) * slices + slice) * groups + t) * 2;
    dst[0] = ss;
    dst[1] = qq;
  }
}

template <int U>
__global__ void gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                const float* __restrict__ partial, int rows, int C, int groups, int vpr,
                                int rows_par, int stat_slices, int slices, float eps, int silu, int pdl, int rev) {
  extern __shared__ float sm[];  // [groups][2] = mean, rstd ; then [8][groups][2] doubles for the slice fold
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  const int n = rev ? static_cast<int>(gridDim.y) - 1 - static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.y);
  const int slice = rev ? static_cast<int>(gridDim.x) - 1 - static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x);
  const int t = threadIdx.x;
  const int cpg = C / groups;
  double* red = reinterpret_cast<double*>(sm + 2 * groups + (2 * groups & 1));  // 8-byte aligned
  // fold the per-slice partials: kGnFoldParts strided sub-sums per group in parallel, then a fixed-order final sum
  for (int i = t; i < groups * kGnFoldParts; i += blockDim.x) {
    const int g = i % groups, part = i / groups;
    double s = 0.0, q = 0.0;
    for (int sl = part; sl < stat_slices; sl += kGnFoldParts) {
      const float* src = partial + ((static_cast<long long>(n) * stat_slices + sl) * groups + g) * 2;
      s += static_cast<double>(src[0]);
      q += static_cast<double>(src[1]);
    }
    red[(part * groups + g) * 2] = s;
    red[(part * groups + g) * 2 + 1] = q;
  }
  __syncthreads();
  if (t < groups) {
    double s = 0.0, q = 0.0;
    for (int part = 0; part < kGnFoldParts; ++part) {
      s += red[(part * groups + t) * 2];
      q += red[(part * groups + t) * 2 + 1];
    }
    const double cnt = static_cast<double>(rows) * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    sm[2 * t] = static_cast<float>(mean);
    sm[2 * t + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  const int v = t % vpr, r0 = t / vpr;
  if (r0 >= rows_par) return;
  float a[8], b[8];
  {
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma + v * 8));
    const uint4 bv = __ldg(reinterpret_cast<const uint4*>(beta + v * 8));
    const __half* gh = reinterpret_cast<const __half*>(&gv);
    const __half* bh = reinterpret_cast<const __half*>(&bv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (v * 8 + e) / cpg;
      const float mean = sm[2 * g], rstd = sm[2 * g + 1];
      a[e] = rstd * __half2float(gh[e]);
      b[e] = __half2float(bh[e]) - mean * a[e];
    }
  }
  const int rows_per_slice = (rows + slices - 1) / slices;
  const int rbeg = slice * rows_per_slice;
  const int rend = min(rows, rbeg + rows_per_slice);
  const long long off = (static_cast<long long>(n) * rows) * C + v * 8;
  auto emit = [&](const uint4& xv, int r) {
    const __half* xh = reinterpret_cast<const __half*>(&xv);
    uint4 ov;
    __half* oh = reinterpret_cast<__half*>(&ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float f = fmaf(__half2float(xh[e]), a[e], b[e]);
      if (silu) {
        f = r16(f);  // the reference rounds the GroupNorm output to fp16 before SiLU (two separate ops)
        if constexpr (U == 8) {
          // v2: f * rcp.approx(1 + ex2.approx(-f log2 e)) (1 ulp fp32 each) instead of the IEEE division.  Static count
          // (tools/sass_loop_stats.py): the v1 SiLU path issues ~187 instructions per 16-byte vector where the HBM roofline
          // leaves 175 (23.4 B / clk / SM, 128 thread-instructions / clk / SM) — it is instruction-issue bound; this
          // path needs ~119 (68 %) and 16 MUFU per vector (73 % of the MUFU rate).  Putting half of the exponentials on the
          // FMA pipe (ex2_poly) was tried on paper: 145 instructions — worse, MUFU is not the limiter here.
          float r;
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex2_approx(f * -1.4426950408889634f)));
          f *= r;
        } else {
          f = f / (1.0f + __expf(-f));
        }
      }
      oh[e] = __float2half_rn(f);
    }
    *reinterpret_cast<uint4*>(y + off + static_cast<long long>(r) * C) = ov;
  };
  int r = rbeg + r0;
  for (; r + (U - 1) * rows_par < rend; r += U * rows_par) {
    uint4 xv[U];
    if constexpr (U == 8) gn_load_batch<U>(x + off + static_cast<long long>(r) * C, static_cast<long long>(rows_par) * C, xv);
    else {
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u] = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r + u * rows_par) * C));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) emit(xv[u], r + u * rows_par);
  }
  for (; r < rend; r += rows_par) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r) * C));
    emit(xv, r);
  }
}


// ---------------------------------------------------------------------------------------------------- GroupNorm, one pass
// Round-2 candidate (AV2V_GN_CLUSTER=1, default off) for the PER-FRAME norms (resnet norm1 / norm2, Transformer2DModel.norm:
// 60 of the 166 GroupNorm calls of a step).  A sample of one frame is small — 4096 rows x 320 channels = 2.6 MB at the
// finest level — so a thread-block CLUSTER can hold a (sample, G-group channel block) slab in shared memory: CS CTAs each load
// rows / CS rows x (G * cpg) channels ONCE (<= 96 KB), reduce their partial sums, exchange them through distributed shared
// memory, and normalise straight from shared memory: x is read once and y written once (two passes over HBM instead of
// three, one launch instead of two).  The clip-level norms (65 536 rows per sample) do not fit and keep the two-kernel path.
__device__ __forceinline__ float ld_dsmem_f32(uint32_t cluster_addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(cluster_addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

constexpr int kGnClMaxG = 8;  // groups per channel block

__global__ void __launch_bounds__(256)
gn_cluster_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                  const __half* __restrict__ beta, int rows, int C, int cpg, int G, int cs, int rows_par, float eps, int silu,
                  int pdl, int rev) {
  extern __shared__ __align__(16) uint8_t gsm[];
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  const int CB = G * cpg;   // channels of this block
  const int VB = CB >> 3;   // 16-byte vectors per row segment
  const int rank = static_cast<int>(cluster_ctarank());
  const int cb = blockIdx.x / cs;  // channel block (gridDim.x = blocks * cs, clusters are consecutive CTAs)
  const int n = rev ? static_cast<int>(gridDim.y) - 1 - static_cast<int>(blockIdx.y) : static_cast<int>(blockIdx.y);  // AV2V_PINGPONG
  const int rows_cta = rows / cs;
  const int t = threadIdx.x;
  const int v = t % VB, r0 = t / VB;
  // smem: [rows_cta][VB] uint4 slab | [rows_par][CB][2] float reduction | part[kGnClMaxG][2] float | stat[kGnClMaxG][2] float
  uint4* slab = reinterpret_cast<uint4*>(gsm);
  float* red = reinterpret_cast<float*>(gsm + static_cast<size_t>(rows_cta) * VB * 16);
  float* part = red + static_cast<size_t>(rows_par) * CB * 2;
  float* stat = part + 2 * kGnClMaxG;
  const long long base = (static_cast<long long>(n) * rows + static_cast<long long>(rank) * rows_cta) * C + cb * CB + v * 8;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  if (r0 < rows_par) {
    // slab fill through cp.async: all of this thread's rows are in flight at once (plain loads would be serialised by ptxas,
    // see gn_stats_async_kernel); every thread later reads back only the slots it filled itself
    for (int r = r0; r < rows_cta; r += rows_par) cp_async16(&slab[r * VB + v], x + base + static_cast<long long>(r) * C);
    cp_async_commit();
    cp_async_wait<0>();
    for (int r = r0; r < rows_cta; r += rows_par) {
      const uint4 a = slab[r * VB + v];
      const __half2* ah = reinterpret_cast<const __half2*>(&a);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = __half22float2(ah[e]);
        s[2 * e] += fa.x;
        s[2 * e + 1] += fa.y;
        q[2 * e] += fa.x * fa.x;
        q[2 * e + 1] += fa.y * fa.y;
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      red[(r0 * CB + v * 8 + e) * 2] = s[e];
      red[(r0 * CB + v * 8 + e) * 2 + 1] = q[e];
    }
  }
  __syncthreads();
  for (int c = t; c < CB; c += blockDim.x) {  // per channel over the rows_par partial rows (fixed order)
    float ss = 0.f, qq = 0.f;
    for (int k = 0; k < rows_par; ++k) {
      ss += red[(k * CB + c) * 2];
      qq += red[(k * CB + c) * 2 + 1];
    }
    red[c * 2] = ss;
    red[c * 2 + 1] = qq;
  }
  __syncthreads();
  if (t < G) {  // per group of this CTA's rows
    float ss = 0.f, qq = 0.f;
    for (int c = t * cpg; c < (t + 1) * cpg; ++c) {
      ss += red[c * 2];
      qq += red[c * 2 + 1];
    }
    part[2 * t] = ss;
    part[2 * t + 1] = qq;
  }
  cluster_sync();  // every CTA's partials are in its shared memory (a block-level barrier too)
  if (t < G) {
    double ss = 0.0, qq = 0.0;
    for (int k = 0; k < cs; ++k) {  // fixed order over the cluster
      ss += static_cast<double>(ld_dsmem_f32(mapa_u32(smem_u32(&part[2 * t]), static_cast<uint32_t>(k))));
      qq += static_cast<double>(ld_dsmem_f32(mapa_u32(smem_u32(&part[2 * t + 1]), static_cast<uint32_t>(k))));
    }
    const double cnt = static_cast<double>(rows) * cpg;
    const double mean = ss / cnt;
    double var = qq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    stat[2 * t] = static_cast<float>(mean);
    stat[2 * t + 1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  cluster_sync();  // nobody leaves (or overwrites `part`) while a peer may still read its shared memory; stat is visible
  if (r0 >= rows_par) return;
  float a[8], b[8];
  {
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma + cb * CB + v * 8));
    const uint4 bv = __ldg(reinterpret_cast<const uint4*>(beta + cb * CB + v * 8));
    const __half* gh = reinterpret_cast<const __half*>(&gv);
    const __half* bh = reinterpret_cast<const __half*>(&bv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (v * 8 + e) / cpg;
      const float mean = stat[2 * g], rstd = stat[2 * g + 1];
      a[e] = rstd * __half2float(gh[e]);
      b[e] = __half2float(bh[e]) - mean * a[e];
    }
  }
  for (int r = r0; r < rows_cta; r += rows_par) {
    const uint4 xv = slab[r * VB + v];
    const __half* xh = reinterpret_cast<const __half*>(&xv);
    uint4 ov;
    __half* oh = reinterpret_cast<__half*>(&ov);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float f = fmaf(__half2float(xh[e]), a[e], b[e]);
      if (silu) {
        f = r16(f);  // the reference rounds the GroupNorm output to fp16 before SiLU (two separate ops)
        float rr;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr) : "f"(1.0f + ex2_approx(f * -1.4426950408889634f)));
        f *= rr;
      }
      oh[e] = __float2half_rn(f);
    }
    *reinterpret_cast<uint4*>(y + base + static_cast<long long>(r) * C) = ov;
  }
}

// -> AV2V_OK if the cluster kernel took the call, 1 if the shape does not fit (caller falls back), < 0 on error
int gn_cluster_try(const av2v_groupnorm_args* a, cudaStream_t stream) {
  const int cpg = a->C / a->groups;
  // channel block of G groups x cluster of cs CTAs: the largest block whose per-CTA slab fits 96 KB (two CTAs per SM), else 128 KB
  int G = 0, cs = 0;
  for (int limit_kb = 96; limit_kb <= 128 && cs == 0; limit_kb += 32) {
    for (int g = kGnClMaxG; g >= 1 && cs == 0; g >>= 1) {
      if (a->groups % g != 0 || (g * cpg) % 8 != 0 || (g * cpg) / 8 > 256) continue;
      const long long slab_total = static_cast<long long>(a->rows) * g * cpg * 2;
      for (int c = 1; c <= 8; c <<= 1) {
        if (a->rows % c == 0 && slab_total / c <= limit_kb * 1024) {
          G = g;
          cs = c;
          break;
        }
      }
    }
  }
  if (cs == 0) return 1;
  const int CB = G * cpg, VB = CB / 8;
  const int blocks = a->groups / G;
  // enough CTAs to fill the machine — or a tensor so small (<= 16 MB) that the call is launch-latency bound anyway (the B = 1
  // norms of the coarse levels: 22-27 us for 2.6-10 MB on the two-kernel path, profiles/r01_loss_ranking.txt)
  const long long total_bytes = static_cast<long long>(a->n_samples) * a->rows * a->C * 2;
  if (static_cast<long long>(blocks) * cs * a->n_samples < sm_count_cached() && total_bytes > (16ll << 20)) return 1;
  const int rows_cta = a->rows / cs;
  int rows_par = 256 / VB;
  if (rows_par > rows_cta) rows_par = rows_cta;
  if (rows_par < 1) return 1;
  const size_t smem = static_cast<size_t>(rows_cta) * VB * 16 + static_cast<size_t>(rows_par) * CB * 2 * sizeof(float) +
                      4 * kGnClMaxG * sizeof(float);
  if (smem > 200 * 1024) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    AV2V_CHECK_CUDA(cudaFuncSetAttribute(gn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int pdl = pdl_enabled();
  AV2V_CHECK_CUDA(launch_ex(gn_cluster_kernel, dim3(static_cast<unsigned>(blocks * cs), static_cast<unsigned>(a->n_samples)), dim3(256),
                            smem, stream, pdl, cs, static_cast<const __half*>(a->x), static_cast<__half*>(a->y),
                            static_cast<const __half*>(a->gamma), static_cast<const __half*>(a->beta), a->rows, a->C, cpg, G, cs,
                            rows_par, a->eps, a->silu, pdl, pick_direction(a->x, a->y, 1)));
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}

// ---------------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row (C <= 2048) lives in registers: sum -> mean, centred sum of squares -> rstd, normalise.
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                 const __half* __restrict__ beta, long long rows, int C, float eps, int pdl, int rev) {
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  const int lane = threadIdx.x & 31;
  long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  if (rev) row = rows - 1 - row;  // AV2V_PINGPONG: rows walked back to front
  const int vpr = C >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * C);
  uint4 v[kVecPerLane];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int vi = lane + i * 32;
    if (vi < vpr) {
      v[i] = __ldg(xr + vi);
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        s += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int vi = lane + i * 32;
    if (vi < vpr) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        q += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / static_cast<float>(C) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + row * C);
#pragma unroll
  for (int i = 0; i < kVecPerLane; ++i) {
    const int vi = lane + i * 32;
    if (vi < vpr) {
      const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma) + vi);
      const uint4 bv = __ldg(reinterpret_cast<const uint4*>(beta) + vi);
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[i]);
      const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
      const __half2* b2 = reinterpret_cast<const __half2*>(&bv);
      uint4 ov;
      uint32_t* ow = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        const float2 g = __half22float2(g2[e]);
        const float2 b = __half22float2(b2[e]);
        ow[e] = pack_half2((f.x - mean) * rstd * g.x + b.x, (f.y - mean) * rstd * g.y + b.y);
      }
      yr[vi] = ov;
    }
  }
}


// ---------------------------------------------------------------------------------------------------- LayerNorm v2
// Round-2 candidate (AV2V_LN_V2=1, default off).  The v1 kernel launches one warp per row (24 576 CTAs for the
// 196 608 x 320 token matrix of the finest level) and, at C = 320, uses 40 of a warp's 64 vector slots: it runs at
// ~52 % of the HBM roofline (profiles/r01_step_profile.txt).  Every I2VGen-XL width is a multiple of 320 = 40 vectors,
// so here LPR = C / 40 lanes (8, 16 or 32) share a row with exactly FIVE 16-byte vectors each: 32 / LPR rows per warp
// per iteration, all lanes busy; warps are persistent (grid-stride over rows), gamma / beta are staged in shared memory once
// per CTA, and the next iteration's vectors are loaded before the current ones are reduced.
template <int LPR>
__global__ void __launch_bounds__(256, 3)  // <= 85 registers: three CTAs (24 warps, 10 vector loads each in flight) per SM
layernorm5_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                  const __half* __restrict__ beta, long long rows, int C, float eps, int pdl, int rev) {
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  constexpr int RPW = 32 / LPR;  // rows per warp iteration
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;     // which of the warp's rows
  const int l = lane % LPR;       // lane inside the row group
  const long long warp_g = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long stride = static_cast<long long>(gridDim.x) * (blockDim.x >> 5) * RPW;
  __shared__ uint4 gb[2][5 * LPR];  // gamma / beta staged once per CTA (kept out of the register file)
  for (int i = threadIdx.x; i < 5 * LPR; i += blockDim.x) {
    gb[0][i] = __ldg(reinterpret_cast<const uint4*>(gamma) + i);
    gb[1][i] = __ldg(reinterpret_cast<const uint4*>(beta) + i);
  }
  __syncthreads();
  const float inv_c = 1.0f / static_cast<float>(C);
  long long row = warp_g * RPW + sub;
  uint4 v[5], vn[5];
  // rev = 1 (this kernel's default; alternated by the host under AV2V_PINGPONG): rows are walked BACK TO FRONT (logical row r
  // -> physical row rows-1-r): the producing GEMM wrote x front to back, so the
  // tail is what L2 still holds when x is about L2-sized, and the head of y — written last here — is what the next GEMM,
  // which reads front to back, finds resident
  auto load = [&](long long r, uint4 (&dst)[5]) {
    if (r < rows) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + (rev ? rows - 1 - r : r) * C);
#pragma unroll
      for (int i = 0; i < 5; ++i) dst[i] = __ldg(xr + l + i * LPR);
    } else {
#pragma unroll
      for (int i = 0; i < 5; ++i) dst[i] = make_uint4(0, 0, 0, 0);
    }
  };
  load(row, v);
  // the loop bound is warp-uniform (row - sub): every lane takes part in the shuffles of every iteration
  for (; row - sub < rows; row += stride) {
    load(row + stride, vn);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        s += f.x + f.y;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        q += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + eps);
    if (row < rows) {
      uint4* yr = reinterpret_cast<uint4*>(y + (rev ? rows - 1 - row : row) * C);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&v[i]);
        const uint4 gvi = gb[0][l + i * LPR], bvi = gb[1][l + i * LPR];
        const __half2* g2 = reinterpret_cast<const __half2*>(&gvi);
        const __half2* b2 = reinterpret_cast<const __half2*>(&bvi);
        uint4 ov;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __half22float2(h2[e]);
          const float2 g = __half22float2(g2[e]);
          const float2 b = __half22float2(b2[e]);
          ow[e] = pack_half2((f.x - mean) * rstd * g.x + b.x, (f.y - mean) * rstd * g.y + b.y);
        }
        yr[l + i * LPR] = ov;
      }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = vn[i];
  }
}

template <int LPR>
int layernorm5_launch(const av2v_layernorm_args* a, cudaStream_t stream) {
  constexpr int RPW = 32 / LPR;
  const int warps = 8;
  long long blocks = (a->rows + warps * RPW - 1) / (warps * RPW);
  const long long cap = static_cast<long long>(sm_count_cached()) * 3;
  if (blocks > cap) blocks = cap;
  AV2V_CHECK_CUDA(launch_ex(layernorm5_kernel<LPR>, dim3(static_cast<unsigned>(blocks)), dim3(warps * 32), 0, stream,
                            pdl_enabled(), 1, static_cast<const __half*>(a->x), static_cast<__half*>(a->y),
                            static_cast<const __half*>(a->gamma), static_cast<const __half*>(a->beta), a->rows, a->C, a->eps,
                            pdl_enabled(), pick_direction(a->x, a->y, 1)));
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}
}  // namespace
}  // namespace av2v

using namespace av2v;

extern "C" int av2v_layernorm_f16(const av2v_layernorm_args* a, av2v_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AV2V_REQUIRE(a != nullptr, AV2V_EINVAL, "layernorm: null args");
  AV2V_REQUIRE(a->rows >= 0 && a->C > 0, AV2V_EINVAL, "layernorm: bad shape");
  if (a->rows == 0) return AV2V_OK;
  AV2V_REQUIRE(a->x && a->y && a->gamma && a->beta, AV2V_EINVAL, "layernorm: null pointer");
  AV2V_REQUIRE(a->C % 8 == 0 && a->C <= 2048, AV2V_ENOSUP, "layernorm: C must be a multiple of 8 and <= 2048 (got %d)", a->C);
  AV2V_REQUIRE(aligned16(a->x) && aligned16(a->y) && aligned16(a->gamma) && aligned16(a->beta), AV2V_EALIGN,
               "layernorm: pointers must be 16-byte aligned");
  if (a->C % 40 == 0 && env_int("AV2V_LN_V2")) {  // round-2 candidate (default off), see layernorm5_kernel
    const int lpr = a->C / 40;
    if (lpr == 8) return layernorm5_launch<8>(a, stream);
    if (lpr == 16) return layernorm5_launch<16>(a, stream);
    if (lpr == 32) return layernorm5_launch<32>(a, stream);
  }
  const int warps = 8;
  const long long blocks = (a->rows + warps - 1) / warps;
  AV2V_REQUIRE(blocks <= 0x7fffffffll, AV2V_EINVAL, "layernorm: too many rows");
  const int vpl = (a->C / 8 + 31) / 32;
  const __half* x = static_cast<const __half*>(a->x);
  __half* y = static_cast<__half*>(a->y);
  const __half* g = static_cast<const __half*>(a->gamma);
  const __half* b = static_cast<const __half*>(a->beta);
  const unsigned grid = static_cast<unsigned>(blocks);
  const int pdl = pdl_enabled();
  const int rev = pick_direction(a->x, a->y);
#define AV2V_LN_LAUNCH(V)                                                                                              \
  do {                                                                                                                 \
    if (pdl) AV2V_CHECK_CUDA(launch_ex(layernorm_kernel<V>, dim3(grid), dim3(warps * 32), 0, stream, 1, 1, x, y, g, b,  \
                                       a->rows, a->C, a->eps, 1, rev));                                                \
    else layernorm_kernel<V><<<grid, warps * 32, 0, stream>>>(x, y, g, b, a->rows, a->C, a->eps, 0, rev);              \
  } while (0)
  if (vpl <= 1) AV2V_LN_LAUNCH(1);
  else if (vpl <= 2) AV2V_LN_LAUNCH(2);
  else if (vpl <= 3) AV2V_LN_LAUNCH(3);
  else if (vpl <= 5) AV2V_LN_LAUNCH(5);
  else AV2V_LN_LAUNCH(8);
#undef AV2V_LN_LAUNCH
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}

extern "C" int av2v_ddim_step_cfg_f16(const av2v_ddim_args* a, av2v_stream_t stream) {
  return ddim_launch(a, static_cast<cudaStream_t>(stream));
}
extern "C" int av2v_ddim_inverse_step_f16(const av2v_ddim_args* a, av2v_stream_t stream) {
  return ddim_launch(a, static_cast<cudaStream_t>(stream));
}

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
  if (env_int("AV2V_GN_CLUSTER") == 1) {  // round-2 candidate (default off): one-pass cluster kernel for the per-frame norms
    const int rc = gn_cluster_try(a, stream);
    if (rc <= 0) return rc;  // taken (0) or error (< 0); 1 = shape does not fit -> two-kernel path below
  }
  const int vpr = a->C / 8;
  int rows_par = 256 / vpr;
  if (rows_par < 1) rows_par = 1;
  if (rows_par > a->rows) rows_par = a->rows;
  int threads = vpr * rows_par;
  threads = (threads + 31) / 32 * 32;
  if (threads < a->groups) threads = (a->groups + 31) / 32 * 32;
  const size_t sm1 = static_cast<size_t>(rows_par) * a->C * 2 * sizeof(float);
  AV2V_REQUIRE(sm1 <= 48 * 1024, AV2V_ENOSUP, "groupnorm: C too large for the stats staging buffer");
  const size_t sm2 = (2 * a->groups + 2) * sizeof(float) + kGnFoldParts * a->groups * 2 * sizeof(double);
  const long long sample_bytes = static_cast<long long>(a->rows) * a->C * 2;
  (void)sample_bytes;
  const int pdl = pdl_enabled();
  // traversal directions (AV2V_PINGPONG): the statistics pass walks opposite to x's producer, the apply pass opposite to the
  // statistics pass (it re-reads x); y is recorded as written in the apply direction.  Without the switch: shipped kernels
  // forward, the v2 statistics kernel back to front.
  const int pingpong = env_int("AV2V_PINGPONG") == 1;
  const int dir_stats = pick_direction(a->x, nullptr, 0);
  const int dir_apply = pingpong ? !dir_stats : 0;
  record_direction(a->y, dir_apply);
  const int gn_v2 = env_int("AV2V_GN_V2") ? 1 : 0;  // round-2 candidate (default off): 8 loads in flight per thread
  const int chunk = a->n_samples;  // L2-sized chunks (stats+apply per <= 32 MB) measured SLOWER (fewer CTAs per launch)
  const __half* xh = static_cast<const __half*>(a->x);
  __half* yh = static_cast<__half*>(a->y);
  for (int s0 = 0; s0 < a->n_samples; s0 += chunk) {
    const int ns = (a->n_samples - s0 < chunk) ? (a->n_samples - s0) : chunk;
    const long long off = static_cast<long long>(s0) * a->rows * a->C;
    const int target_ctas = sm_count_cached() * 4;
    int slices = (target_ctas + ns - 1) / ns;
    const int max_by_rows = (a->rows + rows_par * 8 - 1) / (rows_par * 8);
    if (slices > max_by_rows) slices = max_by_rows;
    if (slices > kGnMaxSlices) slices = kGnMaxSlices;
    if (slices < 1) slices = 1;
    dim3 grid1(slices, ns);
    float* ws = a->workspace + static_cast<long long>(s0) * kGnMaxSlices * kGnMaxGroups * 2;
    const size_t sm_async = static_cast<size_t>(kGnAsyncStages) * kGnAsyncU * threads * sizeof(uint4);
    if (gn_v2 && sm_async <= 48 * 1024)
      AV2V_CHECK_CUDA(launch_ex(gn_stats_async_kernel, grid1, dim3(threads), sm1 > sm_async ? sm1 : sm_async, stream, pdl, 1,
                                xh + off, ws, a->rows, a->C, a->groups, vpr, rows_par, slices, pdl, pingpong ? dir_stats : 1));
    else if (pdl)
      AV2V_CHECK_CUDA(launch_ex(gn_stats_kernel<4>, grid1, dim3(threads), sm1, stream, 1, 1, xh + off, ws, a->rows, a->C, a->groups,
                                vpr, rows_par, slices, 1, dir_stats));
    else
      gn_stats_kernel<4><<<grid1, threads, sm1, stream>>>(xh + off, ws, a->rows, a->C, a->groups, vpr, rows_par, slices, 0,
                                                         dir_stats);
    AV2V_CHECK_CUDA(cudaGetLastError());
    int slices2 = (target_ctas * 2 + ns - 1) / ns;
    const int max2 = (a->rows + rows_par * 8 - 1) / (rows_par * 8);
    if (slices2 > max2) slices2 = max2;
    if (slices2 > 65535) slices2 = 65535;
    if (slices2 < 1) slices2 = 1;
    dim3 grid2(slices2, ns);
    if (gn_v2)
      AV2V_CHECK_CUDA(launch_ex(gn_apply_kernel<8>, grid2, dim3(threads), sm2, stream, pdl, 1, xh + off, yh + off,
                                static_cast<const __half*>(a->gamma), static_cast<const __half*>(a->beta),
                                static_cast<const float*>(ws), a->rows, a->C, a->groups, vpr, rows_par, slices, slices2, a->eps,
                                a->silu, pdl, dir_apply));
    else if (pdl)
      AV2V_CHECK_CUDA(launch_ex(gn_apply_kernel<4>, grid2, dim3(threads), sm2, stream, 1, 1, xh + off, yh + off,
                                static_cast<const __half*>(a->gamma), static_cast<const __half*>(a->beta),
                                static_cast<const float*>(ws), a->rows, a->C, a->groups, vpr, rows_par, slices, slices2, a->eps,
                                a->silu, 1, dir_apply));
    else
      gn_apply_kernel<4><<<grid2, threads, sm2, stream>>>(xh + off, yh + off, static_cast<const __half*>(a->gamma),
                                                       static_cast<const __half*>(a->beta), ws, a->rows, a->C, a->groups,
                                                       vpr, rows_par, slices, slices2, a->eps, a->silu, 0, dir_apply);
  }
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}
