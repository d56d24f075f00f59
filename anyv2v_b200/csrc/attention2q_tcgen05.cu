// Two-query-tile attention on tcgen05 (sm_100a), head_dim 64, rows mode, one V branch (NV = 1).
//
//   O = softmax(Q K^T * scale) V           (self-attention, or cross-attention with seq_kv / kv_batch_div)
//
// The product kernel for plain (NV = 1) attention in rows mode; attention_tcgen05.cu keeps the PnP-injected (NV = 3, shared
// probabilities) and the frames-mode paths.  Measured on B200 (profiles/r02_probe.txt, 48 x 5 heads x 4096^2): 1554 us against
// 1843 us for the one-tile kernel.  Why a second kernel: the one-tile kernel gives one 128-row query tile to two softmax groups that take alternate key tiles
// and hand the row's running max from tile to tile; every key tile is a serial chain
//   S ready -> tcgen05.ld -> max -> hand-over -> ex2 -> tcgen05.st -> P ready -> PV -> S(j+2)
// of ~2000 cycles against the 1024-cycle MUFU bound (profiles/README.md).  At d = 64 and NV = 1 TMEM has room for a
// different layout: TWO query tiles per CTA (256 query rows of one (batch, head)), each owned by one softmax group
// from its first to its last key tile — no hand-over — and P kept in its OWN columns, so that S(j+1) of a query
// tile is issued as soon as its group has pulled S(j) into registers, long before PV(j):
//
//   TMEM columns: S_A [0,128) S_B [128,256) | P_A [256,320) P_B [320,384) (fp16 pairs) | O_A [384,448) O_B [448,512)
//
// Persistent CTA, warp-specialised (all tcgen05 / TMA issue is warp-convergent, see ptx.cuh):
//   warp 0     : TMA producer — Q_A, Q_B (128 x 64 each), K ring, V ring (128 x 64 tiles, SWIZZLE_128B); K/V tiles are
//                shared by the two query tiles (half the K/V smem traffic per score of the one-tile kernel)
//   warp 1     : MMA sequencer — S_x(j+1) = Q_x K(j+1)^T as soon as group x released S_x, O_x += P_x(j) V(j) when P is ready
//   warp 2     : TMEM allocator
//   warps 4-7  : softmax group A (one thread per query row of tile A);  warps 8-11: group B
// The running max is kept per thread (raised only when a tile's max exceeds it by > 2^8; O is then rescaled in TMEM by
// the same thread after PV(j-1)), one ex2 pass per tile.  kPoly > 0 evaluates 25 % / 50 % of the exponentials with a
// Cody-Waite + degree-3 polynomial on the FMA pipe (max rel. error 7.5e-5, six times below fp16 resolution) to get
// under the MUFU bound.
//
// Replaces (reference = library call inside PyTorch): F.scaled_dot_product_attention at pnp_utils.py:208-210 / 314-316
// for the non-injected steps and sites, and attn2 (cross-attention) of the spatial transformers.
#include <cstdlib>

#include "host_util.cuh"
#include "ptx.cuh"

namespace av2v {
namespace {

constexpr int kThreads = 384;
constexpr int kThreadsSplit = 576;  // producer + MMA warp + 16 softmax warps (two threads per query row); registers are allocated as for 640 threads: 96 each
constexpr int TQ = 128;  // query rows per tile (two tiles per CTA)
constexpr int TK = 128;  // keys per tile
constexpr int HD = 64;
constexpr int kTileBytes = TQ * HD * 2;  // 16 KB
constexpr int kStages = 4;
constexpr int kSmemBytes = 2 * kTileBytes /*Q_A, Q_B*/ + 2 * kStages * kTileBytes /*K, V rings*/ + 1024 /*align*/ +
                           1024 /*barriers*/;
constexpr int kSmemBytesSplit = kSmemBytes + 6144;  // + row-maximum / row-sum exchange between the two halves of a row
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: P <= 2^8 fits fp16 comfortably
constexpr uint32_t kColS = 0, kColP = 256, kColO = 384;  // + x * 128 / 64 / 64 for query tile x
constexpr uint32_t kTmemCols = 512;
static_assert(kSmemBytes <= 232448, "smem budget");

struct Attn2qParams {
  int batch, seq, seq_kv, kv_div, heads;
  int q_pairs;  // pairs of query tiles per (batch, head)
  int n_kv;     // key tiles per work item
  int total_items;
  __half* o;
  int ldo;
  float scale_log2;
};

template <int kPoly>
__global__ void __launch_bounds__(kThreads, 1)
attn2q_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const Attn2qParams p) {
  constexpr int S = kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                       // [2][128 x 64]
  uint8_t* smem_k = smem + 2 * kTileBytes;      // [S][128 x 64]
  uint8_t* smem_v = smem_k + S * kTileBytes;    // [S][128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + S * kTileBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* q_empty = bars + 1;     // 1
  uint64_t* k_full = bars + 2;      // S
  uint64_t* k_empty = k_full + S;   // S
  uint64_t* v_full = k_empty + S;   // S
  uint64_t* v_empty = v_full + S;   // S
  uint64_t* s_full = v_empty + S;   // 2: S_x(j) is in TMEM
  uint64_t* s_free = s_full + 2;    // 2: group x has S_x(j) in registers (4 warp arrivals)
  uint64_t* p_ready = s_free + 2;   // 2: P_x(j) is in TMEM (4 warp arrivals)
  uint64_t* pv_done = p_ready + 2;  // 2: PV_x(j) has completed (P_x and O_x may be touched again)
  uint64_t* o_empty = pv_done + 2;  // 2: group x has read the item's O_x (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < S; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_ready[i], 4);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_empty[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // item = (b * heads + h) * q_pairs + qp
  auto decode = [&](int item, int& h, int& b, int& qp) {
    qp = item % p.q_pairs;
    const int bh = item / p.q_pairs;
    h = bh % p.heads;
    b = bh / p.heads;
  };

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // ================================================================== TMA producer
    const uint32_t lead = elect_one() ? 1u : 0u;
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int h, b, qp;
      decode(item, h, b, qp);
      const int q_row_a = b * p.seq + qp * (2 * TQ);
      // tile B of a ragged last pair: re-load tile A's rows (results are masked on store)
      const int q_row_b = (qp * (2 * TQ) + TQ < p.seq) ? q_row_a + TQ : q_row_a;
      const int kv_row0 = (b / p.kv_div) * p.seq_kv;
      mbar_wait(q_empty, (it & 1u) ^ 1u);
      mbar_arrive_expect_tx_w(lead, q_full, 2 * kTileBytes);
      tma_load_2d_w(lead, smem_q, &tmap_q, q_full, h * HD, q_row_a);
      tma_load_2d_w(lead, smem_q + kTileBytes, &tmap_q, q_full, h * HD, q_row_b);
      for (int j = 0; j < p.n_kv; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1u);
        mbar_arrive_expect_tx_w(lead, &k_full[ks], kTileBytes);
        tma_load_2d_w(lead, smem_k + ks * kTileBytes, &tmap_k, &k_full[ks], h * HD, kv_row0 + j * TK);
        if (++ks == S) { ks = 0; kph ^= 1u; }
        mbar_wait(&v_empty[vs], vph ^ 1u);
        mbar_arrive_expect_tx_w(lead, &v_full[vs], kTileBytes);
        tma_load_2d_w(lead, smem_v + vs * kTileBytes, &tmap_v, &v_full[vs], h * HD, kv_row0 + j * TK);
        if (++vs == S) { vs = 0; vph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA sequencer (whole warp, `lead` issues)
    // Order per key tile j of an item (g = global tile counter):
    //   [j == 0: S_A(g), S_B(g)]   S_A(g+1), S_B(g+1)   PV_A(g), PV_B(g)
    // S_x(g+1) only needs group x to have loaded S_x(g) (s_free); PV_x(g) needs P_x(g) (p_ready).  Every wait depends
    // on softmax progress that itself depends only on MMAs issued EARLIER in this order -> no cycle.
    const uint32_t lead = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc_s = make_idesc_f16(TQ, TK, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(TQ, HD, 0, 1);  // B = V, MN-major
    const uint64_t qdesc0 = make_sdesc(smem_u32(smem_q), 16, 1024);
    const uint64_t qdesc1 = make_sdesc(smem_u32(smem_q + kTileBytes), 16, 1024);
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    uint32_t g = 0, it = 0;
    auto issue_s = [&](uint32_t gg, bool last_of_item) {
      mbar_wait(&k_full[ks], kph);
      const uint64_t kdesc = make_sdesc(smem_u32(smem_k + ks * kTileBytes), 16, 1024);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (gg > 0) mbar_wait(&s_free[x], (gg - 1u) & 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + kColS + x * 128;
        const uint64_t qd = x ? qdesc1 : qdesc0;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss_w(lead, d, qd + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit_w(lead, &s_full[x]);
      }
      umma_commit_w(lead, &k_empty[ks]);
      if (last_of_item) umma_commit_w(lead, q_empty);
      if (++ks == S) { ks = 0; kph ^= 1u; }
    };
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      mbar_wait(q_full, it & 1u);
      tc_fence_after();
      for (int j = 0; j < p.n_kv; ++j, ++g) {
        if (j == 0) issue_s(g, p.n_kv == 1);
        if (j + 1 < p.n_kv) issue_s(g + 1, j + 2 == p.n_kv);
        mbar_wait(&v_full[vs], vph);
        const uint32_t v_addr = smem_u32(smem_v + vs * kTileBytes);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          mbar_wait(&p_ready[x], g & 1u);
          if (j == 0 && it > 0) mbar_wait(&o_empty[x], (it - 1u) & 1u);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < TK / 16; ++k) {
            // B: 16 keys = two 8-row groups (SBO 1024 B) of the MN-major V tile; A: 16 keys = 8 TMEM columns of P
            const uint64_t vdesc = make_sdesc(v_addr + k * 2048, kTileBytes, 1024);
            umma_ts_w(lead, tmem_base + kColO + x * 64, tmem_base + kColP + x * 64 + k * 8, vdesc, idesc_o,
                      (j | k) != 0 ? 1u : 0u);
          }
          umma_commit_w(lead, &pv_done[x]);
        }
        umma_commit_w(lead, &v_empty[vs]);
        if (++vs == S) { vs = 0; vph ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================== softmax + epilogue, one thread per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int qd = warp & 3;          // TMEM lane quarter this warp may access
    const int x = (warp - 4) >> 2;    // query tile (0 = A, 1 = B) == softmax group
    const int r = qd * 32 + lane;     // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t sb = tmem_base + kColS + x * 128 + lane_off;
    const uint32_t pb = tmem_base + kColP + x * 64 + lane_off;
    const uint32_t ob = tmem_base + kColO + x * 64 + lane_off;
    uint32_t n = 0;  // key tiles this group has processed (all items)
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int h, b, qp;
      decode(item, h, b, qp);
      float m = 0.f, l = 0.f;
      for (int j = 0; j < p.n_kv; ++j, ++n) {
        mbar_wait(&s_full[x], n & 1u);
        tc_fence_after();
        float s[128];
        {
          uint32_t* su = reinterpret_cast<uint32_t*>(s);
          tmem_ld32(sb + 0, *reinterpret_cast<uint32_t(*)[32]>(su + 0));
          tmem_ld32(sb + 32, *reinterpret_cast<uint32_t(*)[32]>(su + 32));
          tmem_ld32(sb + 64, *reinterpret_cast<uint32_t(*)[32]>(su + 64));
          tmem_ld32(sb + 96, *reinterpret_cast<uint32_t(*)[32]>(su + 96));
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);  // S_x may be overwritten by S_x(j+1)
        const int kv_valid = p.seq_kv - j * TK;
        if (kv_valid < TK) {
#pragma unroll
          for (int c = 0; c < 128; ++c) s[c] = c < kv_valid ? s[c] : -INFINITY;
        }
        float mx0[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 128; c += 4) {
          mx0[0] = fmaxf(mx0[0], s[c]);
          mx0[1] = fmaxf(mx0[1], s[c + 1]);
          mx0[2] = fmaxf(mx0[2], s[c + 2]);
          mx0[3] = fmaxf(mx0[3], s[c + 3]);
        }
        const float rmax = fmaxf(fmaxf(mx0[0], mx0[1]), fmaxf(mx0[2], mx0[3])) * p.scale_log2;  // scale > 0
        bool need = false;
        float m_new;
        if (j == 0) {
          m_new = (rmax == -INFINITY) ? 0.f : rmax;
        } else {
          need = rmax > m + kRescaleThreshold;
          m_new = need ? rmax : m;
        }
        // PV_x(n-1) must have completed before P_x is overwritten and before O_x is rescaled
        if (n > 0) mbar_wait(&pv_done[x], (n - 1u) & 1u);
        tc_fence_after();
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          const float f = need ? ex2_approx(m - m_new) : 1.0f;
          l *= f;
#pragma unroll
          for (int c = 0; c < HD; c += 32) {
            uint32_t o[32];
            tmem_ld32(ob + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st32(ob + c, o);
          }
        }
        m = m_new;
        // P = exp2(s * scale_log2 - m) (fp16, two keys per TMEM column)
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (kPoly == 3) {
          // packed fp32x2 arithmetic (FFMA2 / FADD2: two keys per issue slot) for the scale-subtract and the row sum, and
          // three of every eight key pairs through the packed FMA-pipe polynomial: ~690 issue slots and 640 MUFU cycles per
          // key tile and SM sub-partition instead of ~600 / 1024 (profiles/r01_static_sass_analysis.txt)
          const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m, -m);
          float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float2 a2 = ffma2(make_float2(s[c0 + 2 * e], s[c0 + 2 * e + 1]), sc2, nm2);
              const bool poly = ((e & 7) == 1) || ((e & 7) == 4) || ((e & 7) == 6);
              const float2 p2 = poly ? ex2_poly2(a2) : make_float2(ex2_approx(a2.x), ex2_approx(a2.y));
              if (e & 1) lb = fadd2(lb, p2);
              else la = fadd2(la, p2);
              pk[e] = pack_half2(p2.x, p2.y);
            }
            tmem_st16(pb + (c0 >> 1), pk);
          }
          ls[0] = la.x + la.y;
          ls[1] = lb.x + lb.y;
        } else {
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float a0 = fmaf(s[c0 + 2 * e], p.scale_log2, -m);
              const float a1 = fmaf(s[c0 + 2 * e + 1], p.scale_log2, -m);
              const bool poly = (kPoly == 1) ? ((e & 3) == 3) : (kPoly == 2) ? ((e & 1) == 1) : false;
              const float p0 = poly ? ex2_poly(a0) : ex2_approx(a0);
              const float p1 = poly ? ex2_poly(a1) : ex2_approx(a1);
              ls[e & 3] += p0 + p1;
              pk[e] = pack_half2(p0, p1);
            }
            tmem_st16(pb + (c0 >> 1), pk);
          }
        }
        l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[x]);
      }
      // ---- epilogue: O_x / l -> global (one 128-byte row segment per thread)
      mbar_wait(&pv_done[x], (n - 1u) & 1u);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const int q_in_seq = qp * (2 * TQ) + x * TQ + r;
      const bool valid = q_in_seq < p.seq;
      __half* dst = p.o + (static_cast<long long>(b) * p.seq + q_in_seq) * p.ldo + h * HD;
#pragma unroll
      for (int c = 0; c < HD; c += 32) {
        uint32_t o[32];
        tmem_ld32(ob + c, o);
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[x]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Two threads per query row (each owns 64 of the key tile's 128 columns): sixteen softmax warps, four per scheduler instead of
// two.  The softmax loop is bound by exposed instruction latency (issue slots 44 %, MUFU 37 % busy with two warps per scheduler:
// profiles/r02_attention_split.txt); the two halves of a row agree on the row maximum through shared memory and a 64-thread named
// barrier per key tile, keep separate partial row sums (added once per item) and each rescale / store 32 of O's 64 columns.
__global__ void __launch_bounds__(kThreadsSplit, 1)
attn2q_split_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const Attn2qParams p) {
  constexpr int S = kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                       // [2][128 x 64]
  uint8_t* smem_k = smem + 2 * kTileBytes;      // [S][128 x 64]
  uint8_t* smem_v = smem_k + S * kTileBytes;    // [S][128 x 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + S * kTileBytes);
  uint64_t* q_full = bars;          // 1
  uint64_t* q_empty = bars + 1;     // 1
  uint64_t* k_full = bars + 2;      // S
  uint64_t* k_empty = k_full + S;   // S
  uint64_t* v_full = k_empty + S;   // S
  uint64_t* v_empty = v_full + S;   // S
  uint64_t* s_full = v_empty + S;   // 2: S_x(j) is in TMEM
  uint64_t* s_free = s_full + 2;    // 2: group x has S_x(j) in registers (4 warp arrivals)
  uint64_t* p_ready = s_free + 2;   // 2: P_x(j) is in TMEM (4 warp arrivals)
  uint64_t* pv_done = p_ready + 2;  // 2: PV_x(j) has completed (P_x and O_x may be touched again)
  uint64_t* o_empty = pv_done + 2;  // 2: group x has read the item's O_x (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);
  float* xchg = reinterpret_cast<float*>(bars + 64);  // [parity 2 + row-sum 1][tile 2][half 2][128 rows]: 6 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < S; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 8);
      mbar_init(&p_ready[i], 8);
      mbar_init(&pv_done[i], 1);
      mbar_init(&o_empty[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // item = (b * heads + h) * q_pairs + qp
  auto decode = [&](int item, int& h, int& b, int& qp) {
    qp = item % p.q_pairs;
    const int bh = item / p.q_pairs;
    h = bh % p.heads;
    b = bh / p.heads;
  };

  if (warp == 0) {
    // ================================================================== TMA producer
    const uint32_t lead = elect_one() ? 1u : 0u;
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int h, b, qp;
      decode(item, h, b, qp);
      const int q_row_a = b * p.seq + qp * (2 * TQ);
      // tile B of a ragged last pair: re-load tile A's rows (results are masked on store)
      const int q_row_b = (qp * (2 * TQ) + TQ < p.seq) ? q_row_a + TQ : q_row_a;
      const int kv_row0 = (b / p.kv_div) * p.seq_kv;
      mbar_wait(q_empty, (it & 1u) ^ 1u);
      mbar_arrive_expect_tx_w(lead, q_full, 2 * kTileBytes);
      tma_load_2d_w(lead, smem_q, &tmap_q, q_full, h * HD, q_row_a);
      tma_load_2d_w(lead, smem_q + kTileBytes, &tmap_q, q_full, h * HD, q_row_b);
      for (int j = 0; j < p.n_kv; ++j) {
        mbar_wait(&k_empty[ks], kph ^ 1u);
        mbar_arrive_expect_tx_w(lead, &k_full[ks], kTileBytes);
        tma_load_2d_w(lead, smem_k + ks * kTileBytes, &tmap_k, &k_full[ks], h * HD, kv_row0 + j * TK);
        if (++ks == S) { ks = 0; kph ^= 1u; }
        mbar_wait(&v_empty[vs], vph ^ 1u);
        mbar_arrive_expect_tx_w(lead, &v_full[vs], kTileBytes);
        tma_load_2d_w(lead, smem_v + vs * kTileBytes, &tmap_v, &v_full[vs], h * HD, kv_row0 + j * TK);
        if (++vs == S) { vs = 0; vph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA sequencer (whole warp, `lead` issues)
    // Order per key tile j of an item (g = global tile counter):
    //   [j == 0: S_A(g), S_B(g)]   S_A(g+1), S_B(g+1)   PV_A(g), PV_B(g)
    // S_x(g+1) only needs group x to have loaded S_x(g) (s_free); PV_x(g) needs P_x(g) (p_ready).  Every wait depends
    // on softmax progress that itself depends only on MMAs issued EARLIER in this order -> no cycle.
    const uint32_t lead = elect_one() ? 1u : 0u;
    constexpr uint32_t idesc_s = make_idesc_f16(TQ, TK, 0, 0);
    constexpr uint32_t idesc_o = make_idesc_f16(TQ, HD, 0, 1);  // B = V, MN-major
    const uint64_t qdesc0 = make_sdesc(smem_u32(smem_q), 16, 1024);
    const uint64_t qdesc1 = make_sdesc(smem_u32(smem_q + kTileBytes), 16, 1024);
    int ks = 0, vs = 0;
    uint32_t kph = 0, vph = 0;
    uint32_t g = 0, it = 0;
    auto issue_s = [&](uint32_t gg, bool last_of_item) {
      mbar_wait(&k_full[ks], kph);
      const uint64_t kdesc = make_sdesc(smem_u32(smem_k + ks * kTileBytes), 16, 1024);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (gg > 0) mbar_wait(&s_free[x], (gg - 1u) & 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + kColS + x * 128;
        const uint64_t qd = x ? qdesc1 : qdesc0;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss_w(lead, d, qd + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit_w(lead, &s_full[x]);
      }
      umma_commit_w(lead, &k_empty[ks]);
      if (last_of_item) umma_commit_w(lead, q_empty);
      if (++ks == S) { ks = 0; kph ^= 1u; }
    };
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      mbar_wait(q_full, it & 1u);
      tc_fence_after();
      for (int j = 0; j < p.n_kv; ++j, ++g) {
        if (j == 0) issue_s(g, p.n_kv == 1);
        if (j + 1 < p.n_kv) issue_s(g + 1, j + 2 == p.n_kv);
        mbar_wait(&v_full[vs], vph);
        const uint32_t v_addr = smem_u32(smem_v + vs * kTileBytes);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          mbar_wait(&p_ready[x], g & 1u);
          if (j == 0 && it > 0) mbar_wait(&o_empty[x], (it - 1u) & 1u);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < TK / 16; ++k) {
            // B: 16 keys = two 8-row groups (SBO 1024 B) of the MN-major V tile; A: 16 keys = 8 TMEM columns of P
            const uint64_t vdesc = make_sdesc(v_addr + k * 2048, kTileBytes, 1024);
            umma_ts_w(lead, tmem_base + kColO + x * 64, tmem_base + kColP + x * 64 + k * 8, vdesc, idesc_o,
                      (j | k) != 0 ? 1u : 0u);
          }
          umma_commit_w(lead, &pv_done[x]);
        }
        umma_commit_w(lead, &v_empty[vs]);
        if (++vs == S) { vs = 0; vph ^= 1u; }
      }
    }
  } else if (warp >= 2) {
    // ================================================================== softmax + epilogue, two threads per query row
    const int qd = warp & 3;                 // TMEM lane quarter this warp may access
    const int x = (warp - 2) >> 3;           // query tile (0 = A, 1 = B) == softmax group
    const int hf = ((warp - 2) >> 2) & 1;    // which 64 of the key tile's 128 columns this thread owns (warps 2-5 / 6-9 / ...: all four quarters each)
    const int r = qd * 32 + lane;            // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t sb = tmem_base + kColS + x * 128 + hf * 64 + lane_off;
    const uint32_t pb = tmem_base + kColP + x * 64 + hf * 32 + lane_off;
    const uint32_t ob = tmem_base + kColO + x * 64 + hf * 32 + lane_off;
    const int bar_id = 1 + x * 4 + qd;       // the two warps that share this tile's lane quarter
    float* my_x = xchg + (x * 2 + hf) * 128 + r;
    float* peer_x = xchg + (x * 2 + (hf ^ 1)) * 128 + r;
    uint32_t n = 0;  // key tiles this group has processed (all items)
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      int h, b, qp;
      decode(item, h, b, qp);
      float m = 0.f, l = 0.f;
      for (int j = 0; j < p.n_kv; ++j, ++n) {
        mbar_wait(&s_full[x], n & 1u);
        tc_fence_after();
        float s[64];
        {
          uint32_t* su = reinterpret_cast<uint32_t*>(s);
          tmem_ld32(sb + 0, *reinterpret_cast<uint32_t(*)[32]>(su + 0));
          tmem_ld32(sb + 32, *reinterpret_cast<uint32_t(*)[32]>(su + 32));
          tmem_ld_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[x]);  // S_x may be overwritten by S_x(j+1) once all eight warps have it
        const int kv_valid = p.seq_kv - j * TK - hf * 64;
        if (kv_valid < 64) {
#pragma unroll
          for (int c = 0; c < 64; ++c) s[c] = c < kv_valid ? s[c] : -INFINITY;
        }
        float mx0[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          mx0[0] = fmaxf(mx0[0], s[c]);
          mx0[1] = fmaxf(mx0[1], s[c + 1]);
          mx0[2] = fmaxf(mx0[2], s[c + 2]);
          mx0[3] = fmaxf(mx0[3], s[c + 3]);
        }
        const float hmax = fmaxf(fmaxf(mx0[0], mx0[1]), fmaxf(mx0[2], mx0[3]));
        // the row's maximum = max over both halves: exchanged through shared memory (double-buffered by tile parity)
        my_x[(n & 1u) * 512] = hmax;
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float rmax = fmaxf(hmax, peer_x[(n & 1u) * 512]) * p.scale_log2;  // scale > 0
        bool need = false;
        float m_new;
        if (j == 0) {
          m_new = (rmax == -INFINITY) ? 0.f : rmax;
        } else {
          need = rmax > m + kRescaleThreshold;
          m_new = need ? rmax : m;
        }
        // PV_x(n-1) must have completed before P_x is overwritten and before O_x is rescaled
        if (n > 0) mbar_wait(&pv_done[x], (n - 1u) & 1u);
        tc_fence_after();
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          const float f = need ? ex2_approx(m - m_new) : 1.0f;
          l *= f;
#pragma unroll 1
          for (int c = 0; c < 32; c += 8) {  // this half's 32 of the 64 output columns, 8 at a time (s[64] is live)
            uint32_t o[8];
            tmem_ld8(ob + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st8(ob + c, o);
          }
        }
        m = m_new;
        // P = exp2(s * scale_log2 - m) (fp16, two keys per TMEM column); packed fp32x2 arithmetic, three of every eight key
        // pairs through the FMA-pipe polynomial
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m, -m);
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {  // 8 key pairs at a time: few live registers next to s[64] (96 per thread)
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float2 a2 = ffma2(make_float2(s[c0 + 2 * e], s[c0 + 2 * e + 1]), sc2, nm2);
            const bool poly = (e == 1) || (e == 4) || (e == 6);
            const float2 p2 = poly ? ex2_poly2(a2) : make_float2(ex2_approx(a2.x), ex2_approx(a2.y));
            if (e & 1) lb = fadd2(lb, p2);
            else la = fadd2(la, p2);
            pk[e] = pack_half2(p2.x, p2.y);
          }
          tmem_st8(pb + (c0 >> 1), pk);
        }
        l += (la.x + la.y) + (lb.x + lb.y);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[x]);
      }
      // ---- epilogue: O_x / l -> global; l = this half's partial row sum + the other half's
      my_x[1024] = l;
      asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
      const float inv_l = 1.0f / (l + peer_x[1024]);
      mbar_wait(&pv_done[x], (n - 1u) & 1u);
      tc_fence_after();
      const int q_in_seq = qp * (2 * TQ) + x * TQ + r;
      const bool valid = q_in_seq < p.seq;
      __half* dst = p.o + (static_cast<long long>(b) * p.seq + q_in_seq) * p.ldo + h * HD + hf * 32;
      {
        uint32_t o[32];
        tmem_ld32(ob, o);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            uint4 ov;
            ov.x = pack_half2(__uint_as_float(o[v4 * 8 + 0]) * inv_l, __uint_as_float(o[v4 * 8 + 1]) * inv_l);
            ov.y = pack_half2(__uint_as_float(o[v4 * 8 + 2]) * inv_l, __uint_as_float(o[v4 * 8 + 3]) * inv_l);
            ov.z = pack_half2(__uint_as_float(o[v4 * 8 + 4]) * inv_l, __uint_as_float(o[v4 * 8 + 5]) * inv_l);
            ov.w = pack_half2(__uint_as_float(o[v4 * 8 + 6]) * inv_l, __uint_as_float(o[v4 * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(dst + v4 * 8) = ov;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_empty[x]);
      // the row-sum slot is rewritten at the end of the NEXT item, after at least one more 64-thread barrier: no hazard
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

template <int kPoly>
int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Attn2qParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AV2V_CHECK_CUDA(cudaFuncSetAttribute(attn2q_kernel<kPoly>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const int sms = sm_count_cached();
  attn2q_kernel<kPoly><<<p.total_items < sms ? p.total_items : sms, kThreads, kSmemBytes, stream>>>(tq, tk, tv, p);
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}

int launch_split(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const Attn2qParams& p, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    AV2V_CHECK_CUDA(cudaFuncSetAttribute(attn2q_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesSplit));
    attr_set = true;
  }
  const int sms = sm_count_cached();
  attn2q_split_kernel<<<p.total_items < sms ? p.total_items : sms, kThreadsSplit, kSmemBytesSplit, stream>>>(tq, tk, tv, p);
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}

}  // namespace

// Called by av2v_attn_pnp_f16 (attention_tcgen05.cu) after it validated `a`.  Exponentials: packed fp32x2 arithmetic with 3/8 of
// them on the FMA pipe (kPoly = 3) — the fastest of the four variants measured (all on MUFU 1690 us, 25 % scalar polynomial 1640,
// 50 % scalar 1783, packed 3/8 1554; profiles/r02_probe.txt).
int attn2q_launch(const av2v_attn_args* a, cudaStream_t stream) {
  AV2V_REQUIRE(a->seq_mode == AV2V_SEQ_ROWS && a->n_v == 1, AV2V_ENOSUP, "attn2q: rows mode, n_v = 1 only");
  Attn2qParams p{};
  p.batch = a->batch;
  p.seq = a->seq;
  p.seq_kv = a->seq_kv > 0 ? a->seq_kv : a->seq;
  p.kv_div = a->kv_batch_div > 0 ? a->kv_batch_div : 1;
  p.heads = a->heads;
  p.o = static_cast<__half*>(a->o);
  p.ldo = a->ldo;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  AV2V_REQUIRE(a->batch % p.kv_div == 0, AV2V_EINVAL, "attn: batch must be a multiple of kv_batch_div");
  const uint64_t cols = static_cast<uint64_t>(a->heads) * HD;
  const uint64_t rows = static_cast<uint64_t>(a->batch) * a->seq;
  const uint64_t krows = static_cast<uint64_t>(a->batch / p.kv_div) * p.seq_kv;
  const uint32_t box[2] = {HD, TQ};
  const uint64_t dq[2] = {cols, rows}, sq[1] = {static_cast<uint64_t>(a->ldq) * 2};
  const uint64_t dk[2] = {cols, krows}, sk[1] = {static_cast<uint64_t>(a->ldk) * 2};
  const uint64_t dv[2] = {cols, krows}, sv[1] = {static_cast<uint64_t>(a->ldv) * 2};
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_tmap_f16(&tq, a->q, 2, dq, sq, box)) != AV2V_OK) return rc;
  if ((rc = make_tmap_f16(&tk, a->k, 2, dk, sk, box)) != AV2V_OK) return rc;
  if ((rc = make_tmap_f16(&tv, a->v, 2, dv, sv, box)) != AV2V_OK) return rc;
  p.q_pairs = (a->seq + 2 * TQ - 1) / (2 * TQ);
  p.n_kv = (p.seq_kv + TK - 1) / TK;
  p.total_items = a->batch * a->heads * p.q_pairs;
  // two threads per query row once the key loop is long (measured on B200, profiles/r02_attention_split.txt: 4096 keys 1557 ->
  // 1409 us, 1024 keys 229 -> 228, 256 keys 45.8 -> 48.1, 145 keys 168 -> 176): the per-tile exchange of the row maximum
  // only pays when the softmax warps' exposed latency dominates
  bool split = p.n_kv >= 16;
#ifdef AV2V_GEMM_BRINGUP
  if (const char* e = getenv("AV2V_ATTN_SPLIT")) split = atoi(e) != 0;
#endif
  return split ? launch_split(tq, tk, tv, p, stream) : launch<3>(tq, tk, tv, p, stream);
}

}  // namespace av2v
