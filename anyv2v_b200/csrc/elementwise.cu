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

__global__ void gn_stats_kernel(const __half* __restrict__ x, float* __restrict__ partial, int rows, int C,
                                int groups, int vpr, int rows_par, int slices, int pdl) {
  extern __shared__ float sm[];  // [rows_par][C][2]
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  const int n = blockIdx.y, slice = blockIdx.x;
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
    // 4 independent 16-byte loads in flight per thread (memory-level parallelism)
    for (; r + 3 * rows_par < rend; r += 4 * rows_par) {
      uint4 a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(r + u * rows_par) * C));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
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

__global__ void gn_apply_kernel(const __half* __restrict__ x, __half* __restrict__ y,
                                const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                const float* __restrict__ partial, int rows, int C, int groups, int vpr,
                                int rows_par, int stat_slices, int slices, float eps, int silu, int pdl) {
  extern __shared__ float sm[];  // [groups][2] = mean, rstd ; then [8][groups][2] doubles for the slice fold
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  const int n = blockIdx.y, slice = blockIdx.x;
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
        f = f / (1.0f + __expf(-f));
      }
      oh[e] = __float2half_rn(f);
    }
    *reinterpret_cast<uint4*>(y + off + static_cast<long long>(r) * C) = ov;
  };
  int r = rbeg + r0;
  for (; r + 3 * rows_par < rend; r += 4 * rows_par) {
    uint4 xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xv[u] = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r + u * rows_par) * C));
#pragma unroll
    for (int u = 0; u < 4; ++u) emit(xv[u], r + u * rows_par);
  }
  for (; r < rend; r += rows_par) {
    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + off + static_cast<long long>(r) * C));
    emit(xv, r);
  }
}

// ---------------------------------------------------------------------------------------------------- LayerNorm
// One warp per row; the row (C <= 2048) lives in registers: sum -> mean, centred sum of squares -> rstd, normalise.
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                 const __half* __restrict__ beta, long long rows, int C, float eps, int pdl) {
  pdl_launch_dependents(pdl);
  pdl_wait(pdl);
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
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
                  const __half* __restrict__ beta, long long rows, int C, float eps, int pdl) {
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
  auto load = [&](long long r, uint4 (&dst)[5]) {
    if (r < rows) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + r * C);
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
      uint4* yr = reinterpret_cast<uint4*>(y + row * C);
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
                            pdl_enabled()));
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
#define AV2V_LN_LAUNCH(V)                                                                                              \
  do {                                                                                                                 \
    if (pdl) AV2V_CHECK_CUDA(launch_ex(layernorm_kernel<V>, dim3(grid), dim3(warps * 32), 0, stream, 1, 1, x, y, g, b,  \
                                       a->rows, a->C, a->eps, 1));                                                     \
    else layernorm_kernel<V><<<grid, warps * 32, 0, stream>>>(x, y, g, b, a->rows, a->C, a->eps, 0);                   \
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
    if (pdl)
      AV2V_CHECK_CUDA(launch_ex(gn_stats_kernel, grid1, dim3(threads), sm1, stream, 1, 1, xh + off, ws, a->rows, a->C, a->groups,
                                vpr, rows_par, slices, 1));
    else
      gn_stats_kernel<<<grid1, threads, sm1, stream>>>(xh + off, ws, a->rows, a->C, a->groups, vpr, rows_par, slices, 0);
    AV2V_CHECK_CUDA(cudaGetLastError());
    int slices2 = (target_ctas * 2 + ns - 1) / ns;
    const int max2 = (a->rows + rows_par * 8 - 1) / (rows_par * 8);
    if (slices2 > max2) slices2 = max2;
    if (slices2 > 65535) slices2 = 65535;
    if (slices2 < 1) slices2 = 1;
    dim3 grid2(slices2, ns);
    if (pdl)
      AV2V_CHECK_CUDA(launch_ex(gn_apply_kernel, grid2, dim3(threads), sm2, stream, 1, 1, xh + off, yh + off,
                                static_cast<const __half*>(a->gamma), static_cast<const __half*>(a->beta),
                                static_cast<const float*>(ws), a->rows, a->C, a->groups, vpr, rows_par, slices, slices2, a->eps,
                                a->silu, 1));
    else
      gn_apply_kernel<<<grid2, threads, sm2, stream>>>(xh + off, yh + off, static_cast<const __half*>(a->gamma),
                                                       static_cast<const __half*>(a->beta), ws, a->rows, a->C, a->groups,
                                                       vpr, rows_par, slices, slices2, a->eps, a->silu, 0);
  }
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}
