// PnP self-attention core on tcgen05 (sm_100a), head_dim 64.
//
//   O_j = softmax(Q K^T * scale) V_j      j = 0..NV-1
//
// NV = 1 is ordinary attention (self, or cross with seq_kv / kv_batch_div).  NV = 3 is the PnP-injected step (pnp_utils.py:189-196 / 295-302): the reference
// overwrites q,k of the uncond and cond chunks with the source chunk's, so the probabilities of the three
// branches are identical — they are computed ONCE from the source Q,K and applied to [V_src | V_unc | V_cond] in a
// single 128 x 192 x 128 MMA.  No injection copy, no redundant QK^T / softmax.
//
// Persistent CTA (one per SM), warp-specialised:
//   warp 0   : TMA producer  — Q tile (128 x 64), K tiles (128 x 64, ring), V tiles (NV x 128 x 64, ring), SWIZZLE_128B
//   warp 1   : MMA issuer    — S_b = Q K^T (SS, K-major x K-major) into TMEM, double-buffered;
//                              O += P V (A = P from TMEM, B = V MN-major from smem)
//   warp 2   : TMEM allocator
//   warps 4-11: softmax      — two threads per query row (64 keys each; the two warps of a row share an SM
//                              sub-partition so their streams interleave): tcgen05.ld the scores -> ONE fused pass (ex2
//                              against the running max + this tile's max, independent pipes) -> P (fp16) written back
//                              over S in TMEM; lazy rescale / recompute only when the max grew > 2^8; final O / l -> global
// Temporal attention (AV2V_SEQ_FRAMES) gathers its (pixel, frame) tokens straight from the frame-major
// channels-last activation with a 4-D TMA box [64 ch x PPT pixels x F frames]; 128/F pixels share one 128-row tile
// and a strided mask keeps the sequences apart — no [B,C,F,h,w] -> [B*hw,F,C] transpose is ever materialised.
#include "host_util.cuh"
#include "ptx.cuh"

namespace av2v {
namespace {

constexpr int kThreads = 384;
constexpr int kSoftmaxThreads = 256;  // warps 4-11: two threads per query row (one per 64-key half of a tile)
constexpr int TQ = 128;   // query rows per tile
constexpr int TK = 128;   // keys per tile
constexpr int HD = 64;    // head dim
constexpr int kTileBytes = TQ * HD * 2;  // 16 KB
constexpr float kRescaleThreshold = 8.0f;  // log2 domain: P <= 2^8 fits fp16 comfortably

template <int NV>
struct AttnCfg {
  static constexpr int kStages = (NV == 1) ? 4 : 3;
  static constexpr int kSmemBytes = kTileBytes /*Q*/ + kStages * kTileBytes /*K*/ + kStages * NV * kTileBytes /*V*/ +
                                    1024 /*align*/ + 4096 /*barriers + row-state exchange*/;
  static constexpr int kOCols = 64 * NV;
  static constexpr int kTmemCols = 512;
  static constexpr int kSCol0 = 0, kSCol1 = 128, kOCol = 256;
  static_assert(256 + kOCols <= 512, "TMEM budget");
  static_assert(kSmemBytes <= 232448, "smem budget");
};

struct AttnKParams {
  int seq_mode;
  int batch, seq, seq_kv, kv_div, heads;
  int q_tiles;      // query tiles per (batch, head) [rows mode] or per clip-head [frames mode: pixel tiles * frame tiles]
  int n_kv;         // key tiles per work item
  int total_items;
  // frames mode
  int HW, F, ppt, pix_tiles, f_tiles, box_f;
  int clips;                 // clips per branch
  // branches
  int v_branch_rows;         // rows mode: row offset between V branches; frames mode: clip offset
  long long o_branch_stride; // elements
  __half* o;
  int ldo;
  float scale_log2;
};

#ifdef AV2V_ATTN_TIMERS  // bring-up build only (tools/attn_timer_probe.py): cycles one softmax warp of CTA 0 spends per phase
__device__ unsigned long long g_attn_timers[3][8];  // [softmax half 0 | half 1 | MMA thread]
#define AT_DECL() long long at_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long at_t = clock64(); const long long at_start = at_t
#define AT_MARK(i) do { const long long n_ = clock64(); at_[i] += n_ - at_t; at_t = n_; } while (0)
#define AT_FLUSH() do { if (blockIdx.x == 0 && qd == 0 && lane == 0) { at_[7] = clock64() - at_start; \
    for (int i_ = 0; i_ < 8; ++i_) g_attn_timers[half][i_] = at_[i_]; } } while (0)
#else
#define AT_DECL() do {} while (0)
#define AT_MARK(i) do {} while (0)
#define AT_FLUSH() do {} while (0)
#endif

template <int NV>
__global__ void __launch_bounds__(kThreads, 1)
attn_pnp_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnKParams p) {
  using Cfg = AttnCfg<NV>;
  constexpr int S = Cfg::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + kTileBytes;
  uint8_t* smem_v = smem_k + S * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + S * NV * kTileBytes);
  uint64_t* q_full = bars;           // 1
  uint64_t* q_empty = bars + 1;      // 1
  uint64_t* k_full = bars + 2;       // S
  uint64_t* k_empty = k_full + S;    // S
  uint64_t* v_full = k_empty + S;    // S
  uint64_t* v_empty = v_full + S;    // S
  uint64_t* s_full = v_empty + S;    // 4: [S buffer][key half]
  uint64_t* p_ready = s_full + 4;    // 4: [S buffer][key half]
  uint64_t* pv_done = p_ready + 4;   // 2
  uint64_t* o_empty = pv_done + 2;   // 1
  uint64_t* xbar = o_empty + 1;      // 16: [parity][half][lane quarter] max / sum hand-over between a row's two threads
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 16);
  float* xch = reinterpret_cast<float*>(xbar + 18);  // [2 slots][2 halves][128 rows] max / sum exchange between a row's two threads

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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], kSoftmaxThreads / 2);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&pv_done[i], 1);
    mbar_init(o_empty, kSoftmaxThreads);
    for (int i = 0; i < 16; ++i) mbar_init(&xbar[i], 32);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // provably warp-uniform (uniform registers)

  // work item -> coordinates.  rows mode: item = (b * heads + h) * q_tiles + qt.
  // frames mode: item = ((clip * heads + h) * pix_tiles + pt) * f_tiles + ft.
  auto decode = [&](int item, int& h, int& c_q_row, int& c_pix, int& c_f, int& c_b) {
    if (p.seq_mode == AV2V_SEQ_ROWS) {
      const int qt = item % p.q_tiles;
      const int bh = item / p.q_tiles;
      h = bh % p.heads;
      c_b = bh / p.heads;
      c_q_row = c_b * p.seq + qt * TQ;
      c_pix = 0;
      c_f = qt * TQ;  // query offset inside the sequence (rows mode reuses c_f for masking the tail)
    } else {
      const int ft = item % p.f_tiles;
      int r = item / p.f_tiles;
      const int pt = r % p.pix_tiles;
      r /= p.pix_tiles;
      h = r % p.heads;
      c_b = r / p.heads;
      c_pix = pt * p.ppt;
      c_f = ft * p.box_f;
      c_q_row = 0;
    }
  };

  if (warp == 0) {
    // ================================================================== TMA producer
    {  // warp-convergent issue (uniform operands, one elected lane issues) — see ptx.cuh
      const uint32_t lead = elect_one() ? 1u : 0u;
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        int h, q_row, pix, f0, b;
        decode(item, h, q_row, pix, f0, b);
        mbar_wait(q_empty, (it & 1u) ^ 1u);
        mbar_arrive_expect_tx_w(lead, q_full, kTileBytes);
        if (p.seq_mode == AV2V_SEQ_ROWS) tma_load_2d_w(lead, smem_q, &tmap_q, q_full, h * HD, q_row);
        else tma_load_4d_w(lead, smem_q, &tmap_q, q_full, h * HD, pix, f0, b);
        for (int j = 0; j < p.n_kv; ++j) {
          mbar_wait(&k_empty[ks], kph ^ 1u);
          mbar_arrive_expect_tx_w(lead, &k_full[ks], kTileBytes);
          if (p.seq_mode == AV2V_SEQ_ROWS) tma_load_2d_w(lead, smem_k + ks * kTileBytes, &tmap_k, &k_full[ks], h * HD, (b / p.kv_div) * p.seq_kv + j * TK);
          else tma_load_4d_w(lead, smem_k + ks * kTileBytes, &tmap_k, &k_full[ks], h * HD, pix, j * p.box_f, b);
          if (++ks == S) { ks = 0; kph ^= 1u; }

          mbar_wait(&v_empty[vs], vph ^ 1u);
          mbar_arrive_expect_tx_w(lead, &v_full[vs], NV * kTileBytes);
#pragma unroll
          for (int br = 0; br < NV; ++br) {
            uint8_t* dst = smem_v + (vs * NV + br) * kTileBytes;
            if (p.seq_mode == AV2V_SEQ_ROWS)
              tma_load_2d_w(lead, dst, &tmap_v, &v_full[vs], h * HD, br * p.v_branch_rows + (b / p.kv_div) * p.seq_kv + j * TK);
            else
              tma_load_4d_w(lead, dst, &tmap_v, &v_full[vs], h * HD, pix, j * p.box_f, br * p.v_branch_rows + b);
          }
          if (++vs == S) { vs = 0; vph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    // Every key tile is handled as two independent 64-key halves (h = 0, 1), each with its own S / P sub-buffer and
    // its own barriers:  S_h(j) -> softmax_h(j) -> PV_h(j) -> S_h(j+2) (same sub-buffer).  Nothing a half does waits
    // for the other half's softmax warps, so the two softmax warps of an SM sub-partition can run out of phase (one in
    // its ex2 pass while the other reads TMEM / exchanges the row max).  Both PV halves accumulate into the same O.
    // The whole warp runs this loop convergently (uniform operands -> uniform registers); `lead` issues.
    {
      const uint32_t lead = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc_s = make_idesc_f16(TQ, TK / 2, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(TQ, 64 * NV, 0, 1);  // B = V, MN-major
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t g = 0;  // global key-tile counter (S / P buffer = g & 1)
      uint32_t it = 0;
      const uint64_t qdesc = make_sdesc(smem_u32(smem_q), 16, 1024);
      // S_h(tile gg) = Q K_h^T into S buffer gg & 1, columns [64 h, 64 h + 64); K tile `ks` must have landed
      auto issue_s_half = [&](uint32_t gg, int h) {
        const uint64_t kdesc = make_sdesc(smem_u32(smem_k + ks * kTileBytes + h * (kTileBytes / 2)), 16, 1024);
        const uint32_t d = tmem_base + ((gg & 1u) ? Cfg::kSCol1 : Cfg::kSCol0) + h * 64;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss_w(lead, d, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit_w(lead, &s_full[(gg & 1u) * 2 + h]);
      };
#ifdef AV2V_ATTN_TIMERS
      long long mt_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      long long mt_t = clock64();
      const long long mt_start = mt_t;
#define MT_MARK(i) do { const long long n_ = clock64(); mt_[i] += n_ - mt_t; mt_t = n_; } while (0)
#else
#define MT_MARK(i) do {} while (0)
#endif
      auto k_wait = [&]() {
        MT_MARK(7);
        mbar_wait(&k_full[ks], kph);
        tc_fence_after();
        MT_MARK(2);
      };
      auto k_release = [&]() {
        umma_commit_w(lead, &k_empty[ks]);
        if (++ks == S) { ks = 0; kph ^= 1u; }
      };
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        MT_MARK(7);
        mbar_wait(q_full, it & 1u);
        tc_fence_after();
        MT_MARK(4);
        // prologue: S(0) and S(1) (both sub-buffers are free: the previous item's PVs precede them in the pipe)
        const int pre = p.n_kv < 2 ? p.n_kv : 2;
        for (int j = 0; j < pre; ++j) {
          k_wait();
          issue_s_half(g + j, 0);
          issue_s_half(g + j, 1);
          k_release();
        }
        if (p.n_kv <= 2) umma_commit_w(lead, q_empty);
        for (int j = 0; j < p.n_kv; ++j, ++g) {
          const uint32_t p_tmem = tmem_base + ((g & 1u) ? Cfg::kSCol1 : Cfg::kSCol0);
          const bool more = j + 2 < p.n_kv;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            MT_MARK(7);
            mbar_wait(&p_ready[(g & 1u) * 2 + h], (g >> 1) & 1u);
            MT_MARK(h);
            if (h == 0) {
              if (j == 0) mbar_wait(o_empty, (it & 1u) ^ 1u);
              mbar_wait(&v_full[vs], vph);
              MT_MARK(3);
            }
            tc_fence_after();
            const uint32_t v_addr = smem_u32(smem_v + vs * NV * kTileBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // B: 16 keys = two 8-row groups (SBO 1024 B); branches = 64-wide N atoms 16 KB apart (LBO)
              const uint64_t vdesc = make_sdesc(v_addr + (4 * h + k) * 2048, kTileBytes, 1024);
              umma_ts_w(lead, tmem_base + Cfg::kOCol, p_tmem + h * 64 + k * 8, vdesc, idesc_o, (j | h | k) != 0 ? 1u : 0u);
            }
            if (more) {  // S_h(j+2) re-uses this half's sub-buffer right behind PV_h(j) in the (in-order) tensor pipe
              if (h == 0) k_wait();
              issue_s_half(g + 2, h);
            }
          }
          umma_commit_w(lead, &v_empty[vs]);
          umma_commit_w(lead, &pv_done[g & 1u]);
          if (++vs == S) { vs = 0; vph ^= 1u; }
          if (more) {
            k_release();
            if (j + 3 == p.n_kv) umma_commit_w(lead, q_empty);
          }
        }
      }
#ifdef AV2V_ATTN_TIMERS
      MT_MARK(7);
      if (blockIdx.x == 0 && lead) {
        mt_[6] = clock64() - mt_start;
        for (int i_ = 0; i_ < 8; ++i_) g_attn_timers[2][i_] = mt_[i_];
      }
#endif
    }
  } else if (warp >= 4) {
    // ================================================================== softmax / correction / epilogue
    // TWO threads per query row: warps w and w+4 share a TMEM lane quarter (and an SM sub-partition, so the scheduler
    // interleaves their instruction streams — one softmax warp per scheduler issues at IPC ~0.3, profiles/README.md);
    // `half` picks the 64 keys of the tile a thread owns.  Single fused pass per key tile: the exponentials are taken
    // against the running row max of the PREVIOUS tiles while this tile's (half-)row max is accumulated alongside; the
    // two halves exchange their max through smem.  Only when the row max grew by more than 2^8 — rare after the first
    // tiles — are this tile's probabilities recomputed from the registers and O / l rescaled (lazy rescale).  The row
    // sum is kept per half (both halves use the same max) and added once, in the epilogue.
    const int qd = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;    // key half of the tile: keys [64 half, 64 half + 64)
    const int r = qd * 32 + lane;        // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t ob = tmem_base + Cfg::kOCol + lane_off;
    // max / sum exchange between the two threads of a row WITHOUT a rendezvous: a value is published (smem slot + arrive
    // on this warp's mbarrier) as early as it is known and collected by the partner only when it needs it, so the two
    // warps of an SM sub-partition drift out of phase — one reads TMEM (64 B/clk/SM) while the other runs its ex2's
    // (MUFU) instead of both queueing on the same pipe.  Slots / barriers are double-buffered by exchange parity.
    uint32_t xc = 0;
    auto publish = [&](float v) {
      xch[((xc & 1u) * 2 + half) * TQ + r] = v;
      mbar_arrive(&xbar[((xc & 1u) * 2 + half) * 4 + qd]);
    };
    auto collect = [&]() -> float {
      mbar_wait(&xbar[((xc & 1u) * 2 + (half ^ 1)) * 4 + qd], (xc >> 1) & 1u);
      const float v = xch[((xc & 1u) * 2 + (half ^ 1)) * TQ + r];
      ++xc;
      return v;
    };
    uint32_t g = 0;
    uint32_t it = 0;
    const int ppt_mask = p.ppt - 1;
    const int mine = r & ppt_mask;
    const bool strided_mask = (p.seq_mode == AV2V_SEQ_FRAMES) && (p.ppt > 1);
    AT_DECL();
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      int h, q_row, pix, f0, b;
      decode(item, h, q_row, pix, f0, b);
      float m = 0.f, l = 0.f;
      AT_MARK(6);
      for (int j = 0; j < p.n_kv; ++j, ++g) {
        const uint32_t sbuf = tmem_base + ((g & 1u) ? Cfg::kSCol1 : Cfg::kSCol0) + lane_off;
        const uint32_t sb = sbuf + half * 64;   // this half's scores; its P (fp16 pairs) goes over their first 32 columns
        mbar_wait(&s_full[(g & 1u) * 2 + half], (g >> 1) & 1u);
        tc_fence_after();
        AT_MARK(0);
        float s[64];
        {
          uint32_t* su = reinterpret_cast<uint32_t*>(s);
          tmem_ld32(sb + 0, *reinterpret_cast<uint32_t(*)[32]>(su + 0));
          tmem_ld32(sb + 32, *reinterpret_cast<uint32_t(*)[32]>(su + 32));
          tmem_ld_wait();
        }
        AT_MARK(1);
        // masking: key tail (rows mode / long-F frames mode), sequence separation (packed frames mode)
        if (strided_mask) {
#pragma unroll
          for (int c = 0; c < 64; ++c) s[c] = (((half * 64 + c) & ppt_mask) == mine) ? s[c] : -INFINITY;
        } else {
          const int kv_valid = p.F - j * TK - half * 64;
          if (kv_valid < 64) {
#pragma unroll
            for (int c = 0; c < 64; ++c) s[c] = c < kv_valid ? s[c] : -INFINITY;
          }
        }
        // (half-)row max of this tile: published at once, needed by the partner only after its own exponentials
        float mx0[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
          mx0[0] = fmaxf(mx0[0], s[c]);
          mx0[1] = fmaxf(mx0[1], s[c + 1]);
          mx0[2] = fmaxf(mx0[2], s[c + 2]);
          mx0[3] = fmaxf(mx0[3], s[c + 3]);
        }
        const float mine = fmaxf(fmaxf(mx0[0], mx0[1]), fmaxf(mx0[2], mx0[3]));
        publish(mine);
        AT_MARK(2);
        if (j == 0) {  // the first tile needs the true row max before any exponential
          const float r0 = fmaxf(mine, collect()) * p.scale_log2;
          m = (r0 == -INFINITY) ? 0.f : r0;
        }
        // P = exp2(s * scale_log2 - m) against the running max (fp16, two keys per TMEM column, written over S)
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float p0 = ex2_approx(fmaf(s[c0 + 2 * e], p.scale_log2, -m));
            const float p1 = ex2_approx(fmaf(s[c0 + 2 * e + 1], p.scale_log2, -m));
            ls[e & 3] += p0 + p1;
            pk[e] = pack_half2(p0, p1);
          }
          tmem_st16(sb + (c0 >> 1), pk);  // this half's 64 scores are already in registers: safe to overwrite them
        }
        float lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
        AT_MARK(3);
        if (j > 0) {
          const float rmax = fmaxf(mine, collect()) * p.scale_log2;  // scale > 0
          AT_MARK(4);
          const bool need = rmax > m + kRescaleThreshold;
          if (__any_sync(0xffffffffu, need)) {  // both warps of the pair see the same rows -> take the same branch
            // O holds contributions of tiles < j: wait for PV_{g-1}, rescale this row, redo this tile's P
            mbar_wait(&pv_done[(g - 1) & 1u], ((g - 1) >> 1) & 1u);
            tc_fence_after();
            const float f = need ? ex2_approx(m - rmax) : 1.0f;
            if (need) {
              m = rmax;
              l *= f;
            }
#pragma unroll
            for (int c = 0; c < Cfg::kOCols / 2; c += 32) {  // each half rescales its half of the O columns
              uint32_t o[32];
              tmem_ld32(ob + half * (Cfg::kOCols / 2) + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
              tmem_st32(ob + half * (Cfg::kOCols / 2) + c, o);
            }
            // recompute P against the new max (warp-uniform control flow around the collective tcgen05.st;
            // rows that did not need it reproduce the same values)
            float ls2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
              uint32_t pk[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) {
                const float p0 = ex2_approx(fmaf(s[c0 + 2 * e], p.scale_log2, -m));
                const float p1 = ex2_approx(fmaf(s[c0 + 2 * e + 1], p.scale_log2, -m));
                ls2[e & 3] += p0 + p1;
                pk[e] = pack_half2(p0, p1);
              }
              tmem_st16(sb + (c0 >> 1), pk);
            }
            lsum = (ls2[0] + ls2[1]) + (ls2[2] + ls2[3]);
            // PV_h(g) of EITHER half accumulates into all O columns: neither may be issued before both threads of the row
            // have rescaled their O columns -> rendezvous (both took this branch: same rows, same rmax, same m)
            tmem_st_wait();
            tc_fence_before();
            publish(0.f);
            (void)collect();
            tc_fence_after();
          }
        }
        l += lsum;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[(g & 1u) * 2 + half]);
        AT_MARK(5);
      }
      // ---- epilogue: O / l -> global (each half writes 32 of every branch's 64 columns)
      publish(l);
      const float inv_l = 1.0f / (l + collect());
      const uint32_t gl = g - 1;
      mbar_wait(&pv_done[gl & 1u], (gl >> 1) & 1u);
      tc_fence_after();
      long long row;
      bool valid;
      if (p.seq_mode == AV2V_SEQ_ROWS) {
        valid = (f0 + r) < p.seq;
        row = static_cast<long long>(q_row) + r;
      } else {
        const int f = f0 + (r / p.ppt);
        const int px = pix + (r & ppt_mask);
        valid = (f < p.F) && (px < p.HW);
        row = (static_cast<long long>(b) * p.F + f) * p.HW + px;
      }
#pragma unroll
      for (int br = 0; br < NV; ++br) {
        __half* dst = p.o + br * p.o_branch_stride + row * p.ldo + h * HD + half * 32;
        uint32_t o[32];
        tmem_ld32(ob + br * 64 + half * 32, o);
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
      mbar_arrive(o_empty);
    }
    AT_FLUSH();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

template <int NV>
int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnKParams& p,
                cudaStream_t stream) {
  using Cfg = AttnCfg<NV>;
  static bool attr_set = false;
  if (!attr_set) {
    AV2V_CHECK_CUDA(cudaFuncSetAttribute(attn_pnp_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes));
    attr_set = true;
  }
  const int sms = sm_count_cached();
  const int grid = p.total_items < sms ? p.total_items : sms;
  attn_pnp_kernel<NV><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(tq, tk, tv, p);
  AV2V_CHECK_CUDA(cudaGetLastError());
  return AV2V_OK;
}

}  // namespace
}  // namespace av2v

using namespace av2v;

#ifdef AV2V_ATTN_TIMERS
extern "C" int av2v_attn_debug_timers(unsigned long long* out16) {
  AV2V_CHECK_CUDA(cudaMemcpyFromSymbol(out16, av2v::g_attn_timers, sizeof(unsigned long long) * 24));
  return AV2V_OK;
}
#endif

extern "C" int av2v_attn_pnp_f16(const av2v_attn_args* a, av2v_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AV2V_REQUIRE(a != nullptr, AV2V_EINVAL, "attn: null args");
  AV2V_REQUIRE(a->q && a->k && a->v && a->o, AV2V_EINVAL, "attn: null q/k/v/o");
  AV2V_REQUIRE(a->batch > 0 && a->seq > 0 && a->heads > 0, AV2V_EINVAL, "attn: batch/seq/heads must be positive");
  AV2V_REQUIRE(a->n_v == 1 || a->n_v == 3, AV2V_EINVAL, "attn: n_v must be 1 or 3 (got %d)", a->n_v);
  AV2V_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, AV2V_EALIGN,
               "attn: row strides must be multiples of 8 elements");
  AV2V_REQUIRE(a->ldq >= a->heads * HD && a->ldk >= a->heads * HD && a->ldv >= a->heads * HD &&
                   a->ldo >= a->heads * HD,
               AV2V_EINVAL, "attn: row strides must cover heads*64 columns");
  AV2V_REQUIRE(aligned16(a->q) && aligned16(a->k) && aligned16(a->v) && aligned16(a->o), AV2V_EALIGN,
               "attn: q/k/v/o must be 16-byte aligned");
  AV2V_REQUIRE(a->scale > 0.f, AV2V_EINVAL, "attn: scale must be positive");

  AttnKParams p{};
  p.seq_mode = a->seq_mode;
  p.batch = a->batch;
  p.seq = a->seq;
  p.seq_kv = a->seq_kv > 0 ? a->seq_kv : a->seq;
  p.kv_div = a->kv_batch_div > 0 ? a->kv_batch_div : 1;
  p.heads = a->heads;
  p.o = static_cast<__half*>(a->o);
  p.ldo = a->ldo;
  p.o_branch_stride = a->o_branch_stride;
  p.scale_log2 = a->scale * 1.4426950408889634f;
  p.ppt = 1;

  CUtensorMap tq, tk, tv;
  int rc;
  const uint64_t cols = static_cast<uint64_t>(a->heads) * HD;
  if (a->seq_mode == AV2V_SEQ_ROWS) {
    AV2V_REQUIRE(a->batch % p.kv_div == 0, AV2V_EINVAL, "attn: batch must be a multiple of kv_batch_div");
    const uint64_t rows = static_cast<uint64_t>(a->batch) * a->seq;
    const uint64_t krows = static_cast<uint64_t>(a->batch / p.kv_div) * p.seq_kv;
    if (a->n_v == 3) {
      AV2V_REQUIRE(a->v_branch_stride % a->ldv == 0, AV2V_EINVAL, "attn: v_branch_stride must be a multiple of ldv");
      p.v_branch_rows = static_cast<int>(a->v_branch_stride / a->ldv);
    }
    const uint64_t vrows = (a->n_v == 3) ? (2ull * p.v_branch_rows + krows) : krows;
    const uint32_t box[2] = {HD, TQ};
    const uint64_t dq[2] = {cols, rows}, sq[1] = {static_cast<uint64_t>(a->ldq) * 2};
    const uint64_t dk[2] = {cols, krows}, sk[1] = {static_cast<uint64_t>(a->ldk) * 2};
    const uint64_t dv[2] = {cols, vrows}, sv[1] = {static_cast<uint64_t>(a->ldv) * 2};
    if ((rc = make_tmap_f16(&tq, a->q, 2, dq, sq, box)) != AV2V_OK) return rc;
    if ((rc = make_tmap_f16(&tk, a->k, 2, dk, sk, box)) != AV2V_OK) return rc;
    if ((rc = make_tmap_f16(&tv, a->v, 2, dv, sv, box)) != AV2V_OK) return rc;
    p.q_tiles = (a->seq + TQ - 1) / TQ;
    p.n_kv = (p.seq_kv + TK - 1) / TK;
    p.total_items = a->batch * a->heads * p.q_tiles;
    p.F = p.seq_kv;  // key-tail masking uses the key/value sequence length
  } else if (a->seq_mode == AV2V_SEQ_FRAMES) {
    const int F = a->seq, HW = a->HW;
    AV2V_REQUIRE(HW > 0 && a->batch % HW == 0, AV2V_EINVAL, "attn/frames: batch must be clips*HW");
    AV2V_REQUIRE((a->seq_kv <= 0 || a->seq_kv == a->seq) && p.kv_div == 1, AV2V_ENOSUP,
                 "attn/frames: self-attention only (seq_kv / kv_batch_div are rows-mode options)");
    const int clips = a->batch / HW;
    AV2V_REQUIRE((F <= 128 && 128 % F == 0) || (F % 128 == 0), AV2V_ENOSUP,
                 "attn/frames: F must divide 128 or be a multiple of 128 (got %d)", F);
    p.F = F;
    p.HW = HW;
    p.clips = clips;
    p.box_f = F < 128 ? F : 128;
    p.ppt = 128 / p.box_f;
    AV2V_REQUIRE(p.ppt <= 256, AV2V_ENOSUP, "attn/frames: F too small");
    p.pix_tiles = (HW + p.ppt - 1) / p.ppt;  // a ragged last pixel tile is zero-filled by TMA and masked on store
    p.f_tiles = (F + 127) / 128;
    p.n_kv = p.f_tiles;
    p.q_tiles = p.pix_tiles * p.f_tiles;
    p.total_items = clips * a->heads * p.q_tiles;
    if (a->n_v == 3) {
      const long long clip_elems = static_cast<long long>(F) * HW * a->ldv;
      AV2V_REQUIRE(a->v_branch_stride % clip_elems == 0, AV2V_EINVAL,
                   "attn/frames: v_branch_stride must be a whole number of clips");
      p.v_branch_rows = static_cast<int>(a->v_branch_stride / clip_elems);
    }
    const uint64_t vclips = (a->n_v == 3) ? (2ull * p.v_branch_rows + clips) : clips;
    const uint32_t box[4] = {HD, static_cast<uint32_t>(p.ppt), static_cast<uint32_t>(p.box_f), 1};
    auto mk = [&](CUtensorMap* m, const void* base, int ld, uint64_t nclips) {
      const uint64_t d[4] = {cols, static_cast<uint64_t>(HW), static_cast<uint64_t>(F), nclips};
      const uint64_t s[3] = {static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(ld) * 2 * HW,
                             static_cast<uint64_t>(ld) * 2 * HW * F};
      return make_tmap_f16(m, base, 4, d, s, box);
    };
    if ((rc = mk(&tq, a->q, a->ldq, clips)) != AV2V_OK) return rc;
    if ((rc = mk(&tk, a->k, a->ldk, clips)) != AV2V_OK) return rc;
    if ((rc = mk(&tv, a->v, a->ldv, vclips)) != AV2V_OK) return rc;
  } else {
    return fail(AV2V_EINVAL, "attn: unknown seq_mode %d", a->seq_mode);
  }
  return a->n_v == 3 ? launch_attn<3>(tq, tk, tv, p, stream) : launch_attn<1>(tq, tk, tv, p, stream);
}
