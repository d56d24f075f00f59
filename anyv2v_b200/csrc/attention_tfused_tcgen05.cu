// Temporal self-attention with the Q/K/V projection fused in (sm_100a, head_dim 64, F <= 128 frames per sequence).
//
//   per (clip, pixel) sequence of F tokens:   O_h = softmax((X Wq_h^T)(X Wk_h^T)^T * scale) (X Wv_h^T)        h = 0 .. heads-1
//
// The product path for every temporal self-attention (unet_i2vgen_xl.AttnProcessor).  BASELINE.json's north_star asks for the temporal self-attention as a fused QKV-project + SDPA kernel.  At F = 16 a
// 128-row tile — 128 / F pixels x F frames, gathered from the frame-major channels-last activation by the same 4-D TMA
// box the attention kernels use — holds COMPLETE sequences, so the projection can live in the attention kernel and Q, K, V
// never reach HBM (today: QKV GEMM writes 3 x 126 MB at the 64 x 64 level and the attention kernel reads them back:
// 161 + 185 us per call against ~40 us of X-read + O-write).
//
// Work item = (clip, head, pixel tile).  Persistent CTA, 8 warps:
//   warp 0    : TMA producer — per 64-channel k-block: X box (128 tokens x 64) + the head's rows of Wq, Wk, Wv (3 x 64 x 64,
//               stacked into ONE 192-row K-major tile), ring of kStages
//   warp 1    : MMA sequencer —  QKV(i): [Q|K|V] (128 x 192, fp32 in TMEM) += X_kb W_kb^T over the k-blocks
//                                S(i)  : S = Q K^T   (A = Q re-packed to fp16 IN TMEM, B = K tile in smem)
//                                PV(i) : O = P V     (A = P fp16 in TMEM over S, B = V tile in smem, MN-major)
//               issue order  QKV(0) | conv(0)? S(0) QKV(1) p(0)? PV(0) | conv(1)? S(1) QKV(2) p(1)? PV(1) ...  so that the
//               projection of the NEXT item runs under the softmax of the current one (the accumulators are separate)
//   warp 2    : TMEM allocator
//   warps 4-7 : one thread per token row: convert (Q -> fp16 in TMEM, K / V -> fp16 SWIZZLE_128B smem tiles, the layout
//               TMA would have produced), block-diagonal softmax (the 128 / F sequences of a tile are kept apart by the
//               strided mask of the frames mode), epilogue O / l -> global
// TMEM columns: [Q K V] fp32 [0,192) | Q fp16 [192,224) | S / P [256,384) | O [384,448).
// Rounding points are those of the unfused path: Q, K, V rounded to fp16 (what the QKV GEMM stores), P to fp16.
//
// PnP-injected steps (n_v = 3; pnp_utils.py:295-302 overwrites q, k of the uncond / cond chunks with the source chunk's): the
// batch holds [source | uncond | cond] clips; the item of an edit-branch clip b projects Q and K from the SOURCE clip
// (b mod clips-per-branch) and V from clip b itself: its projection is two operand streams through the same ring
// ([Q | K] += X_src [Wq;Wk]^T over the k-blocks, then V += X_b Wv^T) instead of one.  The injection is the choice of the TMA
// coordinate; no q / k tensors, no copies.
//
// A W-RESIDENT build (every CTA serves one head and keeps its [Wq;Wk;Wv] slice in shared memory, Cx <= 320) was measured in round 2:
// L2 -> SM traffic 1.54 -> 0.65 GB per call, the same time (the one-slot kernel is chain-bound, not feed-bound) — removed when the
// two-slot kernel below took over the shapes it served; see profiles/r02_tattn_two_slots.txt for what it would take to bring it back.
//
// Replaces (reference): to_q / to_k / to_v + F.scaled_dot_product_attention of the temporal transformers' attn1 / attn2
// (pnp_utils.py:247-334 is the reference's restatement of that processor), injected and non-injected steps.
#include <cstdlib>

#include "host_util.cuh"
#include "ptx.cuh"

namespace av2v {
namespace {

constexpr int kThreads = 256;
constexpr int TQ = 128;  // tokens per tile
constexpr int HD = 64;
constexpr int BK = 64;
constexpr int kXBytes = TQ * BK * 2;       // 16 KB
constexpr int kWBytes = 3 * HD * BK * 2;   // 24 KB: rows [Wq_h ; Wk_h ; Wv_h] of one k-block
constexpr int kTileBytes = TQ * HD * 2;    // K / V tiles, 16 KB each
struct TCfg {
  static constexpr int kStageBytes = kXBytes + kWBytes;
  static constexpr int kStages = 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kTileBytes + 1024 /*align*/ + 1024 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "smem budget");
};
constexpr uint32_t kColQKV = 0, kColQ16 = 192, kColS = 256, kColO = 384, kTmemCols = 512;

struct TFusedParams {
  int clips, F, HW, heads, Cx;  // clips: ALL clips of the batch (3 x branch_clips when injected)
  int branch_clips;             // injected: clips per branch; Q / K come from clip (b mod branch_clips)
  int num_kb;
  int ppt, pix_tiles, total_items;
  __half* o;
  int ldo;
  float scale_log2;
};

__global__ void __launch_bounds__(kThreads, 1)
tattn_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const TFusedParams p) {
  constexpr int S = TCfg::kStages;
  constexpr int kStageBytes = TCfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_ring = smem;                              // [S][X 16 KB | W 24 KB]
  uint8_t* smem_k = smem_ring + S * kStageBytes;          // K_h tile, 128 keys x 64, K-major SWIZZLE_128B
  uint8_t* smem_v = smem_k + kTileBytes;                  // V_h tile, same image (consumed MN-major)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kTileBytes);
  uint64_t* full = bars;            // S
  uint64_t* empty = bars + S;       // S
  uint64_t* qkv_full = bars + 2 * S;  // projection of an item complete (tcgen05.commit)
  uint64_t* conv_done = qkv_full + 1; // 4 warps: Q16 in TMEM, K / V tiles in smem; the fp32 accumulators are free
  uint64_t* s_full = conv_done + 1;
  uint64_t* p_ready = s_full + 1;     // 4 warps
  uint64_t* o_full = p_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(qkv_full, 1);
    mbar_init(conv_done, 4);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // Work item = ((clip * heads + h) * pix_tiles + pt); CTA c takes items c, c + G, ...
  // src = the clip Q / K are projected from; two = the item needs two operand streams (injected step, edit-branch clip).
  auto get_item = [&](int i, int& h, int& pix, int& b, int& src, bool& two) -> bool {
    const int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
    if (item >= p.total_items) return false;
    const int pt = item % p.pix_tiles;
    const int r = item / p.pix_tiles;
    h = r % p.heads;
    b = r / p.heads;
    pix = pt * p.ppt;
    src = b % p.branch_clips;
    two = b >= p.branch_clips;
    return true;
  };

  if (warp == 0) {
    // ================================================================== TMA producer
    const uint32_t lead = elect_one() ? 1u : 0u;
    int stage = 0;
    uint32_t phase = 0;
    const int inner = p.heads * HD;
    int h, pix, b, src;
    bool two;
    // one operand stream = num_kb stages of X (clip `xc`) + rows [w0, w0 + 64 * nw) of the head's W slice
    auto stream = [&](int xc, int w0, int nw) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* sx = smem_ring + stage * kStageBytes;
        mbar_arrive_expect_tx_w(lead, &full[stage], kXBytes + nw * HD * BK * 2);
        tma_load_4d_w(lead, sx, &tmap_x, &full[stage], kb * BK, pix, 0, xc);
        for (int j = 0; j < nw; ++j)  // Wq / Wk / Wv rows of head h, stacked from the start of the stage's W region
          tma_load_2d_w(lead, sx + kXBytes + j * HD * BK * 2, &tmap_w, &full[stage], kb * BK, (w0 + j) * inner + h * HD);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    };
    for (int i = 0; get_item(i, h, pix, b, src, two); ++i) {
      if (!two) {
        stream(b, 0, 3);    // [Q | K | V] from the item's own clip (on injected steps: a source clip)
      } else {
        stream(src, 0, 2);  // [Q | K] from the source clip
        stream(b, 2, 1);    // V from the item's own clip
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA sequencer (whole warp, `lead` issues)
    const uint32_t lead = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc_qkv = make_idesc_f16(TQ, 3 * HD, 0, 0);  // 128 x 192, both operands K-major
    constexpr uint32_t idesc_qk = make_idesc_f16(TQ, 2 * HD, 0, 0);   // injected: [Q | K] from the source clip's tile ...
    constexpr uint32_t idesc_v = make_idesc_f16(TQ, HD, 0, 0);        // ... and V from the item's own
    constexpr uint32_t idesc_s = make_idesc_f16(TQ, TQ, 0, 0);        // S = Q K^T
    constexpr uint32_t idesc_o = make_idesc_f16(TQ, HD, 0, 1);        // O = P V, B = V MN-major
    int stage = 0;
    uint32_t phase = 0;
    // projection of item i: one stream (128 x 192) or two ([Q | K] 128 x 128 from the source clip, then V 128 x 64)
    auto issue_stream = [&](uint32_t d_col, uint32_t idesc) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t xdesc = make_sdesc(smem_u32(smem_ring + stage * kStageBytes), 16, 1024);
        const uint64_t wdesc = make_sdesc(smem_u32(smem_ring + stage * kStageBytes + kXBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_ss_w(lead, tmem_base + d_col, xdesc + 2 * k, wdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit_w(lead, &empty[stage]);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    };
    auto issue_qkv = [&](int i) {
      int h, pix, b, src;
      bool two;
      get_item(i, h, pix, b, src, two);
      if (!two) {
        issue_stream(kColQKV, idesc_qkv);
      } else {
        issue_stream(kColQKV, idesc_qk);
        issue_stream(kColQKV + 2 * HD, idesc_v);
      }
      umma_commit_w(lead, qkv_full);
    };
    uint32_t it = 0;
    int my_items = 0;
    {
      int h, pix, b, src;
      bool two;
      while (get_item(my_items, h, pix, b, src, two)) ++my_items;
    }
    if (my_items > 0) issue_qkv(0);
    for (int i = 0; i < my_items; ++i, ++it) {
      mbar_wait(conv_done, it & 1u);  // Q16 / K / V of item i are in place, the fp32 accumulators are free
      tc_fence_after();
      const uint64_t kdesc = make_sdesc(smem_u32(smem_k), 16, 1024);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_ts_w(lead, tmem_base + kColS, tmem_base + kColQ16 + k * 8, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
      umma_commit_w(lead, s_full);
      if (i + 1 < my_items) issue_qkv(i + 1);  // runs under the softmax of item i
      mbar_wait(p_ready, it & 1u);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(smem_v);
#pragma unroll
      for (int k = 0; k < TQ / 16; ++k) {
        const uint64_t vdesc = make_sdesc(v_addr + k * 2048, kTileBytes, 1024);
        umma_ts_w(lead, tmem_base + kColO, tmem_base + kColS + k * 8, vdesc, idesc_o, k != 0 ? 1u : 0u);
      }
      umma_commit_w(lead, o_full);
    }
  } else if (warp >= 4) {
    // ================================================================== convert / softmax / epilogue, one thread per token
    const int qd = warp & 3;
    const int r = qd * 32 + lane;  // tile row == TMEM lane; frame-major: r = f * ppt + p
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tb = tmem_base + lane_off;
    const int ppt_mask = p.ppt - 1;
    const int mine = r & ppt_mask;
    uint32_t it = 0;
    int h, pix, b, src_;
    bool two_;
    for (int i = 0; get_item(i, h, pix, b, src_, two_); ++i, ++it) {
      // ---- convert: fp32 accumulators -> fp16 operands (the rounding the QKV GEMM's store would have done)
      mbar_wait(qkv_full, it & 1u);
      tc_fence_after();
      {
        uint32_t v[32], w[32];
        // Q -> TMEM, two fp16 per column
        tmem_ld32(tb + kColQKV + 0, v);
        tmem_ld32(tb + kColQKV + 32, w);
        tmem_ld_wait();
        uint32_t q16[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          q16[e] = pack_half2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          q16[16 + e] = pack_half2(__uint_as_float(w[2 * e]), __uint_as_float(w[2 * e + 1]));
        }
        tmem_st32(tb + kColQ16, q16);
        // K, V -> smem rows of 128 B, 16-byte chunk j stored at j ^ (row & 7) (SWIZZLE_128B, 8-row / 1024-B atoms)
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          uint8_t* row = (m == 0 ? smem_k : smem_v) + r * 128;
          tmem_ld32(tb + kColQKV + 64 + m * 64, v);
          tmem_ld32(tb + kColQKV + 64 + m * 64 + 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 lo, hi;
            lo.x = pack_half2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            lo.y = pack_half2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            lo.z = pack_half2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            lo.w = pack_half2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            hi.x = pack_half2(__uint_as_float(w[8 * j + 0]), __uint_as_float(w[8 * j + 1]));
            hi.y = pack_half2(__uint_as_float(w[8 * j + 2]), __uint_as_float(w[8 * j + 3]));
            hi.z = pack_half2(__uint_as_float(w[8 * j + 4]), __uint_as_float(w[8 * j + 5]));
            hi.w = pack_half2(__uint_as_float(w[8 * j + 6]), __uint_as_float(w[8 * j + 7]));
            *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = lo;        // channels 8j .. 8j+7
            *reinterpret_cast<uint4*>(row + (((j + 4) ^ (r & 7)) << 4)) = hi;  // channels 32+8j .. 32+8j+7
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core's operand reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(conv_done);
      }
      // ---- softmax over the row's own sequence (columns of the same pixel), single key tile
      mbar_wait(s_full, it & 1u);
      tc_fence_after();
      float s[128];
      {
        uint32_t* su = reinterpret_cast<uint32_t*>(s);
        tmem_ld32(tb + kColS + 0, *reinterpret_cast<uint32_t(*)[32]>(su + 0));
        tmem_ld32(tb + kColS + 32, *reinterpret_cast<uint32_t(*)[32]>(su + 32));
        tmem_ld32(tb + kColS + 64, *reinterpret_cast<uint32_t(*)[32]>(su + 64));
        tmem_ld32(tb + kColS + 96, *reinterpret_cast<uint32_t(*)[32]>(su + 96));
        tmem_ld_wait();
      }
      if (p.ppt > 1) {
#pragma unroll
        for (int c = 0; c < 128; ++c) s[c] = ((c & ppt_mask) == mine) ? s[c] : -INFINITY;
      } else if (p.F < TQ) {
#pragma unroll
        for (int c = 0; c < 128; ++c) s[c] = c < p.F ? s[c] : -INFINITY;
      }
      float mx0[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        mx0[0] = fmaxf(mx0[0], s[c]);
        mx0[1] = fmaxf(mx0[1], s[c + 1]);
        mx0[2] = fmaxf(mx0[2], s[c + 2]);
        mx0[3] = fmaxf(mx0[3], s[c + 3]);
      }
      const float m = fmaxf(fmaxf(mx0[0], mx0[1]), fmaxf(mx0[2], mx0[3])) * p.scale_log2;  // own key always valid -> finite
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(s[c0 + 2 * e], p.scale_log2, -m));
          const float p1 = ex2_approx(fmaf(s[c0 + 2 * e + 1], p.scale_log2, -m));
          ls[e & 3] += p0 + p1;
          pk[e] = pack_half2(p0, p1);
        }
        tmem_st16(tb + kColS + (c0 >> 1), pk);  // P over S (all scores are in registers)
      }
      const float l = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // ---- epilogue
      mbar_wait(o_full, it & 1u);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int f = r / p.ppt;
      const int px = pix + (r & ppt_mask);
      const bool valid = (f < p.F) && (px < p.HW);
      __half* dst = p.o + ((static_cast<long long>(b) * p.F + f) * p.HW + px) * p.ldo + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t o[32];
        tmem_ld32(tb + kColO + c, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 ov;
            ov.x = pack_half2(__uint_as_float(o[v4 * 8 + 0]) * inv_l, __uint_as_float(o[v4 * 8 + 1]) * inv_l);
            ov.y = pack_half2(__uint_as_float(o[v4 * 8 + 2]) * inv_l, __uint_as_float(o[v4 * 8 + 3]) * inv_l);
            ov.z = pack_half2(__uint_as_float(o[v4 * 8 + 4]) * inv_l, __uint_as_float(o[v4 * 8 + 5]) * inv_l);
            ov.w = pack_half2(__uint_as_float(o[v4 * 8 + 6]) * inv_l, __uint_as_float(o[v4 * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c + v4 * 8) = ov;
          }
        }
      }
      tc_fence_before();  // O / S / Q16 reads of this item precede the next item's tcgen05 writes (ordered by conv_done / p_ready)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Two items in flight (tattn_fused2_kernel).  The single-warpgroup kernel above is bound by its per-item chain
//   projection -> convert -> S -> softmax -> PV -> epilogue -> (same warps) convert of the next item
// (6400 cycles per item of which 2432 are tensor-pipe work, profiles/r02_tattn_two_slots.txt).  Here TWO convert / softmax warpgroups
// take alternate items, each with its own TMEM slot and K / V tiles, so one item's softmax runs under the other's convert / MMAs:
//   TMEM slot s (256 columns): [Q K V] fp32 accumulators [0,192) — re-used in place: Q fp16 over the Q accumulator [0,32) (every
//   thread packs its own lane after reading it), S = Q K^T over the dead K / V accumulators [64,192), P fp16 over S [64,128) —
//   and O [192,256).  MMA issue order:  QKV(0) QKV(1) | S(i)  PV(i-1)  QKV(i+1) | ...   (QKV(i+1) goes into the slot PV(i-1) just
//   finished with; the tensor pipe is in order, the waits are conv_done(i) before S(i) and p_ready(i-1) before PV(i-1)).
// Streamed operands only (the K / V tiles of the second slot take the room of the resident weight slice).
constexpr int kThreads2 = 384;
constexpr int kStages2 = 4;
constexpr int kStageBytes2 = kXBytes + kWBytes;
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 4 * kTileBytes + 1024 /*align*/ + 1024 /*barriers*/;
static_assert(kSmemBytes2 <= 232448, "smem budget");
constexpr uint32_t kSlotCols = 256, kRelQ16 = 0, kRelS = 64, kRelO = 192;

__global__ void __launch_bounds__(kThreads2, 1)
tattn_fused2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const TFusedParams p) {
  constexpr int S = kStages2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_ring = smem;                               // [S][X 16 KB | W 24 KB]
  uint8_t* smem_kv = smem_ring + S * kStageBytes2;         // [slot][K tile | V tile], 16 KB each
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + 4 * kTileBytes);
  uint64_t* full = bars;               // S
  uint64_t* empty = bars + S;          // S
  uint64_t* qkv_full = bars + 2 * S;   // [2] projection of the slot's item complete
  uint64_t* conv_done = qkv_full + 2;  // [2] 4 warps: Q16 in TMEM, K / V tiles in smem
  uint64_t* s_full = conv_done + 2;    // [2]
  uint64_t* p_ready = s_full + 2;      // [2] 4 warps
  uint64_t* o_full = p_ready + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qkv_full[s], 1);
      mbar_init(&conv_done[s], 4);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);
      mbar_init(&o_full[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // item = ((clip * heads + h) * pix_tiles + pt), CTA c takes items c, c + G, ...
  auto get_item = [&](int i, int& h, int& pix, int& b, int& src, bool& two) -> bool {
    const int item = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
    if (item >= p.total_items) return false;
    const int pt = item % p.pix_tiles;
    const int r = item / p.pix_tiles;
    h = r % p.heads;
    b = r / p.heads;
    pix = pt * p.ppt;
    src = b % p.branch_clips;
    two = b >= p.branch_clips;
    return true;
  };

  if (warp == 0) {
    // ================================================================== TMA producer (as in the one-slot kernel, streamed build)
    const uint32_t lead = elect_one() ? 1u : 0u;
    int stage = 0;
    uint32_t phase = 0;
    const int inner = p.heads * HD;
    int h, pix, b, src;
    bool two;
    auto stream = [&](int xc, int w0, int nw) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* sx = smem_ring + stage * kStageBytes2;
        mbar_arrive_expect_tx_w(lead, &full[stage], kXBytes + nw * HD * BK * 2);
        tma_load_4d_w(lead, sx, &tmap_x, &full[stage], kb * BK, pix, 0, xc);
        for (int j = 0; j < nw; ++j)
          tma_load_2d_w(lead, sx + kXBytes + j * HD * BK * 2, &tmap_w, &full[stage], kb * BK, (w0 + j) * inner + h * HD);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    };
    for (int i = 0; get_item(i, h, pix, b, src, two); ++i) {
      if (!two) {
        stream(b, 0, 3);
      } else {
        stream(src, 0, 2);
        stream(b, 2, 1);
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA sequencer (whole warp, `lead` issues)
    const uint32_t lead = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc_qkv = make_idesc_f16(TQ, 3 * HD, 0, 0);
    constexpr uint32_t idesc_qk = make_idesc_f16(TQ, 2 * HD, 0, 0);
    constexpr uint32_t idesc_v = make_idesc_f16(TQ, HD, 0, 0);
    constexpr uint32_t idesc_s = make_idesc_f16(TQ, TQ, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(TQ, HD, 0, 1);
    int stage = 0;
    uint32_t phase = 0;
    auto issue_stream = [&](uint32_t d_col, uint32_t idesc) {
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint64_t xdesc = make_sdesc(smem_u32(smem_ring + stage * kStageBytes2), 16, 1024);
        const uint64_t wdesc = make_sdesc(smem_u32(smem_ring + stage * kStageBytes2 + kXBytes), 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          umma_ss_w(lead, tmem_base + d_col, xdesc + 2 * k, wdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        umma_commit_w(lead, &empty[stage]);
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    };
    auto issue_qkv = [&](int i) {
      int h, pix, b, src;
      bool two;
      get_item(i, h, pix, b, src, two);
      const uint32_t col = static_cast<uint32_t>(i & 1) * kSlotCols;
      if (!two) {
        issue_stream(col, idesc_qkv);
      } else {
        issue_stream(col, idesc_qk);
        issue_stream(col + 2 * HD, idesc_v);
      }
      umma_commit_w(lead, &qkv_full[i & 1]);
    };
    auto issue_pv = [&](int i) {  // O(i) = P(i) V(i)
      const int s = i & 1;
      mbar_wait(&p_ready[s], static_cast<uint32_t>(i >> 1) & 1u);
      tc_fence_after();
      const uint32_t v_addr = smem_u32(smem_kv + (2 * s + 1) * kTileBytes);
      const uint32_t tb = tmem_base + s * kSlotCols;
#pragma unroll
      for (int k = 0; k < TQ / 16; ++k) {
        const uint64_t vdesc = make_sdesc(v_addr + k * 2048, kTileBytes, 1024);
        umma_ts_w(lead, tb + kRelO, tb + kRelS + k * 8, vdesc, idesc_o, k != 0 ? 1u : 0u);
      }
      umma_commit_w(lead, &o_full[s]);
    };
    int my_items = 0;
    {
      int h, pix, b, src;
      bool two;
      while (get_item(my_items, h, pix, b, src, two)) ++my_items;
    }
    if (my_items > 0) issue_qkv(0);
    if (my_items > 1) issue_qkv(1);
    for (int i = 0; i < my_items; ++i) {
      const int s = i & 1;
      mbar_wait(&conv_done[s], static_cast<uint32_t>(i >> 1) & 1u);  // Q16 / K / V of item i in place, its K / V accumulators dead
      tc_fence_after();
      const uint32_t tb = tmem_base + s * kSlotCols;
      const uint64_t kdesc = make_sdesc(smem_u32(smem_kv + 2 * s * kTileBytes), 16, 1024);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k)
        umma_ts_w(lead, tb + kRelS, tb + kRelQ16 + k * 8, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
      umma_commit_w(lead, &s_full[s]);
      if (i >= 1) {
        issue_pv(i - 1);
        if (i + 1 < my_items) issue_qkv(i + 1);  // into the slot PV(i-1) has just finished with
      }
    }
    if (my_items > 0) issue_pv(my_items - 1);
  } else if (warp >= 4) {
    // ================================================================== convert / softmax / epilogue: warpgroup g takes items g, g + 2, ...
    const int qd = warp & 3;
    const int g = (warp - 4) >> 2;
    const int r = qd * 32 + lane;  // tile row == TMEM lane; frame-major: r = f * ppt + p
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t tb = tmem_base + g * kSlotCols + lane_off;
    uint8_t* smem_k = smem_kv + 2 * g * kTileBytes;
    uint8_t* smem_v = smem_k + kTileBytes;
    const int ppt_mask = p.ppt - 1;
    const int mine = r & ppt_mask;
    uint32_t it = 0;
    int h, pix, b, src_;
    bool two_;
    for (int i = g; get_item(i, h, pix, b, src_, two_); i += 2, ++it) {
      // ---- convert: fp32 accumulators -> fp16 operands (the rounding the QKV GEMM's store would have done)
      mbar_wait(&qkv_full[g], it & 1u);
      tc_fence_after();
      {
        uint32_t v[32], w[32];
        tmem_ld32(tb + 0, v);
        tmem_ld32(tb + 32, w);
        tmem_ld_wait();
        uint32_t q16[32];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          q16[e] = pack_half2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          q16[16 + e] = pack_half2(__uint_as_float(w[2 * e]), __uint_as_float(w[2 * e + 1]));
        }
        tmem_st32(tb + kRelQ16, q16);  // over this lane's Q accumulator, which this thread has just read
#pragma unroll
        for (int m = 0; m < 2; ++m) {
          uint8_t* row = (m == 0 ? smem_k : smem_v) + r * 128;
          tmem_ld32(tb + 64 + m * 64, v);
          tmem_ld32(tb + 64 + m * 64 + 32, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 lo, hi;
            lo.x = pack_half2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            lo.y = pack_half2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            lo.z = pack_half2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            lo.w = pack_half2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            hi.x = pack_half2(__uint_as_float(w[8 * j + 0]), __uint_as_float(w[8 * j + 1]));
            hi.y = pack_half2(__uint_as_float(w[8 * j + 2]), __uint_as_float(w[8 * j + 3]));
            hi.z = pack_half2(__uint_as_float(w[8 * j + 4]), __uint_as_float(w[8 * j + 5]));
            hi.w = pack_half2(__uint_as_float(w[8 * j + 6]), __uint_as_float(w[8 * j + 7]));
            *reinterpret_cast<uint4*>(row + ((j ^ (r & 7)) << 4)) = lo;
            *reinterpret_cast<uint4*>(row + (((j + 4) ^ (r & 7)) << 4)) = hi;
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_done[g]);
      }
      // ---- softmax over the row's own sequence (columns of the same pixel), single key tile
      mbar_wait(&s_full[g], it & 1u);
      tc_fence_after();
      float s[128];
      {
        uint32_t* su = reinterpret_cast<uint32_t*>(s);
        tmem_ld32(tb + kRelS + 0, *reinterpret_cast<uint32_t(*)[32]>(su + 0));
        tmem_ld32(tb + kRelS + 32, *reinterpret_cast<uint32_t(*)[32]>(su + 32));
        tmem_ld32(tb + kRelS + 64, *reinterpret_cast<uint32_t(*)[32]>(su + 64));
        tmem_ld32(tb + kRelS + 96, *reinterpret_cast<uint32_t(*)[32]>(su + 96));
        tmem_ld_wait();
      }
      if (p.ppt > 1) {
#pragma unroll
        for (int c = 0; c < 128; ++c) s[c] = ((c & ppt_mask) == mine) ? s[c] : -INFINITY;
      } else if (p.F < TQ) {
#pragma unroll
        for (int c = 0; c < 128; ++c) s[c] = c < p.F ? s[c] : -INFINITY;
      }
      float mx0[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 128; c += 4) {
        mx0[0] = fmaxf(mx0[0], s[c]);
        mx0[1] = fmaxf(mx0[1], s[c + 1]);
        mx0[2] = fmaxf(mx0[2], s[c + 2]);
        mx0[3] = fmaxf(mx0[3], s[c + 3]);
      }
      const float m = fmaxf(fmaxf(mx0[0], mx0[1]), fmaxf(mx0[2], mx0[3])) * p.scale_log2;
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float p0 = ex2_approx(fmaf(s[c0 + 2 * e], p.scale_log2, -m));
          const float p1 = ex2_approx(fmaf(s[c0 + 2 * e + 1], p.scale_log2, -m));
          ls[e & 3] += p0 + p1;
          pk[e] = pack_half2(p0, p1);
        }
        tmem_st16(tb + kRelS + (c0 >> 1), pk);  // P over S (all scores are in registers)
      }
      const float l = (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g]);
      // ---- epilogue
      mbar_wait(&o_full[g], it & 1u);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int f = r / p.ppt;
      const int px = pix + (r & ppt_mask);
      const bool valid = (f < p.F) && (px < p.HW);
      __half* dst = p.o + ((static_cast<long long>(b) * p.F + f) * p.HW + px) * p.ldo + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t o[32];
        tmem_ld32(tb + kRelO + c, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 ov;
            ov.x = pack_half2(__uint_as_float(o[v4 * 8 + 0]) * inv_l, __uint_as_float(o[v4 * 8 + 1]) * inv_l);
            ov.y = pack_half2(__uint_as_float(o[v4 * 8 + 2]) * inv_l, __uint_as_float(o[v4 * 8 + 3]) * inv_l);
            ov.z = pack_half2(__uint_as_float(o[v4 * 8 + 4]) * inv_l, __uint_as_float(o[v4 * 8 + 5]) * inv_l);
            ov.w = pack_half2(__uint_as_float(o[v4 * 8 + 6]) * inv_l, __uint_as_float(o[v4 * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + c + v4 * 8) = ov;
          }
        }
      }
      tc_fence_before();  // O / S / Q16 reads of this item precede the slot's next tcgen05 writes (ordered by conv_done / p_ready)
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace
}  // namespace av2v

using namespace av2v;

extern "C" int av2v_tattn_fused_f16(const av2v_tattn_fused_args* a, av2v_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AV2V_REQUIRE(a != nullptr, AV2V_EINVAL, "tattn_fused: null args");
  AV2V_REQUIRE(a->x && a->wqkv && a->o, AV2V_EINVAL, "tattn_fused: null x / wqkv / o");
  AV2V_REQUIRE(a->clips > 0 && a->F > 0 && a->HW > 0 && a->heads > 0 && a->Cx > 0, AV2V_EINVAL, "tattn_fused: bad shape");
  AV2V_REQUIRE(a->F <= 128 && 128 % a->F == 0, AV2V_ENOSUP, "tattn_fused: F must divide 128 (got %d)", a->F);
  AV2V_REQUIRE(a->Cx % BK == 0, AV2V_ENOSUP, "tattn_fused: the input width must be a multiple of 64 (got %d)", a->Cx);
  AV2V_REQUIRE(a->ldx % 8 == 0 && a->ldx >= a->Cx && a->ldo % 8 == 0 && a->ldo >= a->heads * HD, AV2V_EINVAL,
               "tattn_fused: row strides must be multiples of 8 and cover the row");
  AV2V_REQUIRE(aligned16(a->x) && aligned16(a->wqkv) && aligned16(a->o), AV2V_EALIGN, "tattn_fused: pointers must be 16-byte aligned");
  AV2V_REQUIRE(a->scale > 0.f, AV2V_EINVAL, "tattn_fused: scale must be positive");
  AV2V_REQUIRE(a->n_v == 1 || a->n_v == 3, AV2V_EINVAL, "tattn_fused: n_v must be 1 or 3 (got %d)", a->n_v);
  AV2V_REQUIRE(a->n_v == 1 || a->clips % 3 == 0, AV2V_EINVAL, "tattn_fused: n_v = 3 needs clips = 3 x clips-per-branch (got %d)", a->clips);

  TFusedParams p{};
  p.clips = a->clips;
  p.branch_clips = a->n_v == 3 ? a->clips / 3 : a->clips;
  p.F = a->F;
  p.HW = a->HW;
  p.heads = a->heads;
  p.Cx = a->Cx;
  p.num_kb = a->Cx / BK;
  p.ppt = 128 / a->F;
  p.pix_tiles = (a->HW + p.ppt - 1) / p.ppt;
  p.total_items = a->clips * a->heads * p.pix_tiles;
  p.o = static_cast<__half*>(a->o);
  p.ldo = a->ldo;
  p.scale_log2 = a->scale * 1.4426950408889634f;

  CUtensorMap tx, tw;
  int rc;
  {
    const uint64_t d[4] = {static_cast<uint64_t>(a->Cx), static_cast<uint64_t>(a->HW), static_cast<uint64_t>(a->F),
                           static_cast<uint64_t>(a->clips)};
    const uint64_t s[3] = {static_cast<uint64_t>(a->ldx) * 2, static_cast<uint64_t>(a->ldx) * 2 * a->HW,
                           static_cast<uint64_t>(a->ldx) * 2 * a->HW * a->F};
    const uint32_t box[4] = {BK, static_cast<uint32_t>(p.ppt), static_cast<uint32_t>(a->F), 1};
    if ((rc = make_tmap_f16(&tx, a->x, 4, d, s, box)) != AV2V_OK) return rc;
  }
  {
    const uint64_t d[2] = {static_cast<uint64_t>(a->Cx), 3ull * a->heads * HD};
    const uint64_t s[1] = {static_cast<uint64_t>(a->Cx) * 2};
    const uint32_t box[2] = {BK, HD};
    if ((rc = make_tmap_f16(&tw, a->wqkv, 2, d, s, box)) != AV2V_OK) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    AV2V_CHECK_CUDA(cudaFuncSetAttribute(tattn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCfg::kSmemBytes));
    attr_set = true;
  }
  const int sms = sm_count_cached();
  // two items in flight: at least two items per CTA (else nothing overlaps) and Cx <= 640 (measured: -11 ... -18 % at Cx = 320,
  // -5 % at 640, none / worse at 1280 where the projection dominates; profiles/r02_tattn_two_slots.txt)
  bool two_slots = p.total_items >= 2 * sms && p.num_kb <= 10;
#ifdef AV2V_GEMM_BRINGUP
  if (const char* e = getenv("AV2V_TATTN_SLOTS")) two_slots = atoi(e) == 2;
#endif
  if (two_slots) {
    static bool attr2 = false;
    if (!attr2) {
      AV2V_CHECK_CUDA(cudaFuncSetAttribute(tattn_fused2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
      attr2 = true;
    }
    const int grid = p.total_items < sms ? p.total_items : sms;
    tattn_fused2_kernel<<<grid, kThreads2, kSmemBytes2, stream>>>(tx, tw, p);
    AV2V_CHECK_CUDA(cudaGetLastError());
    return AV2V_OK;
  }
  const int grid = p.total_items < sms ? p.total_items : sms;
  tattn_fused_kernel<<<grid, kThreads, TCfg::kSmemBytes, stream>>>(tx, tw, p);
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}
