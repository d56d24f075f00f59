// HBM-bound row kernels of the hot path: fused CFG + DDIM step (K7) and LayerNorm.  GroupNorm(+SiLU) (K6) is groupnorm.cu.
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
                 const float* __restrict__ coef_dev) {
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
  ddim_step_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      static_cast<const __half*>(a->x), static_cast<const __half*>(a->v_neg), static_cast<const __half*>(a->v_edit),
      static_cast<__half*>(a->out), a->n, a->guidance, a->ca, a->cb, a->cc, a->cd, a->coef_dev);
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}

// ---------------------------------------------------------------------------------------------------- LayerNorm
// Fallback for widths that are not a multiple of 320 (none in I2VGen-XL; tiny test configs): one warp per row; the row
// (C <= 2048) lives in registers: sum -> mean, centred sum of squares -> rstd, normalise.
template <int kVecPerLane>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                 const __half* __restrict__ beta, long long rows, int C, float eps) {
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


// ---------------------------------------------------------------------------------------------------- LayerNorm, C = 40 * LPR vectors
// The product path (measured on B200, profiles/r02_probe.txt: 49.1 us = 5.1 TB/s on the 196 608 x 320 token matrix of the finest
// level against 70.8 us for one-warp-per-row).  Every I2VGen-XL width is a multiple of 320 = 40 vectors, so LPR = C / 40 lanes (8, 16 or 32) share a row with exactly FIVE 16-byte vectors each: 32 / LPR rows per warp
// per iteration, all lanes busy; warps are persistent (grid-stride over rows), gamma / beta are staged in shared memory once
// per CTA, and the next iteration's vectors are loaded before the current ones are reduced.
template <int LPR>
__global__ void __launch_bounds__(256, 3)  // <= 85 registers: three CTAs (24 warps, 10 vector loads each in flight) per SM
layernorm5_kernel(const __half* __restrict__ x, __half* __restrict__ y, const __half* __restrict__ gamma,
                  const __half* __restrict__ beta, long long rows, int C, float eps) {
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
  // rows are walked BACK TO FRONT (logical row r -> physical row rows-1-r): the producing GEMM wrote x front to back, so the
  // tail is what L2 still holds when x is about L2-sized, and the head of y — written last here — is what the next GEMM,
  // which reads front to back, finds resident
  auto load = [&](long long r, uint4 (&dst)[5]) {
    if (r < rows) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + (rows - 1 - r) * C);
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
      uint4* yr = reinterpret_cast<uint4*>(y + (rows - 1 - row) * C);
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
  layernorm5_kernel<LPR><<<static_cast<unsigned>(blocks), warps * 32, 0, stream>>>(
      static_cast<const __half*>(a->x), static_cast<__half*>(a->y), static_cast<const __half*>(a->gamma),
      static_cast<const __half*>(a->beta), a->rows, a->C, a->eps);
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
  if (a->C % 40 == 0) {
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
#define AV2V_LN_LAUNCH(V) layernorm_kernel<V><<<grid, warps * 32, 0, stream>>>(x, y, g, b, a->rows, a->C, a->eps)
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
