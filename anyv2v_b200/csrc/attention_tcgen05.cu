// PnP self-attention core on tcgen05 (sm_100a), head_dim 64.
//
//   O_j = softmax(Q K^T * scale) V_j      j = 0..NV-1
//
// NV = 1 is ordinary attention (self, or cross with seq_kv / kv_batch_div).  NV = 3 is the PnP-injected step (pnp_utils.py:189-196 / 295-302): the reference
// overwrites q,k of the uncond and cond chunks with the source chunk's, so the probabilities of the three
// branches are identical — they are computed ONCE from the source Q,K and applied to [V_src | V_unc | V_cond] in a
// single 128 x 192 x 128 MMA.  No injection copy, no redundant QK^T / softmax.
//
// Persistent CTA (one per SM), warp-specialised (all tcgen05 / TMA issue is warp-convergent, see ptx.cuh):
//   warp 0   : TMA producer  — Q tile (128 x 64), K tiles (128 x 64, ring), V tiles (NV x 128 x 64, ring), SWIZZLE_128B
//   warp 1   : MMA issuer    — S_b = Q K^T (SS, K-major x K-major) into TMEM, double-buffered;
//                              O += P V (A = P from TMEM, B = V MN-major from smem)
//   warp 2   : TMEM allocator
//   warps 4-11: softmax      — TWO groups of 128 threads (one thread per query row each) that take ALTERNATE key tiles:
//                              group b owns S/P buffer b.  While one group runs its ex2 pass (MUFU) the other waits for
//                              its S tile / reads TMEM / stores P, so the MUFU pipe — the bound at d = 64 — stays busy.
//                              Per tile: tcgen05.ld the 128 scores -> row max -> running max decided BEFORE the
//                              exponentials (handed from tile to tile between the row's two threads through smem; raised
//                              only when it grew > 2^8, then O is rescaled) -> ONE ex2 pass -> P (fp16) over S in TMEM.
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
  uint64_t* s_full = v_empty + S;    // 2: [S buffer]
  uint64_t* p_ready = s_full + 2;    // 2: [S buffer] (the 128 threads of the buffer's softmax group)
  uint64_t* pv_done = p_ready + 2;   // 2
  uint64_t* o_empty = pv_done + 2;   // 1
  uint64_t* xbar = o_empty + 1;      // 16: [slot][sending group][lane quarter] running-max / row-sum hand-over
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 16);
  float* xch = reinterpret_cast<float*>(xbar + 18);  // [2 slots][2 groups][128 rows] outboxes

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
      mbar_init(&p_ready[i], kSoftmaxThreads / 2);
      mbar_init(&pv_done[i], 1);
    }
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

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");  // control warpgroup: registers go to the softmax warps
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
    // The whole warp runs this loop convergently (uniform operands -> uniform registers); `lead` issues.
    // Per key tile j (S / P buffer j & 1):  S(j) -> softmax group (j & 1) -> PV(j) -> S(j+2) into the same buffer.
    {
      const uint32_t lead = elect_one() ? 1u : 0u;
      constexpr uint32_t idesc_s = make_idesc_f16(TQ, TK, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_f16(TQ, 64 * NV, 0, 1);  // B = V, MN-major
      int ks = 0, vs = 0;
      uint32_t kph = 0, vph = 0;
      uint32_t g = 0;  // global key-tile counter (S / P buffer = g & 1)
      uint32_t it = 0;
      const uint64_t qdesc = make_sdesc(smem_u32(smem_q), 16, 1024);
#ifdef AV2V_ATTN_TIMERS
      long long mt_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      long long mt_t = clock64();
      const long long mt_start = mt_t;
#define MT_MARK(i) do { const long long n_ = clock64(); mt_[i] += n_ - mt_t; mt_t = n_; } while (0)
#else
#define MT_MARK(i) do {} while (0)
#endif
      auto issue_s = [&](uint32_t gg) {  // S(tile gg) = Q K^T into buffer gg & 1; waits for / releases K stage `ks`
        MT_MARK(7);
        mbar_wait(&k_full[ks], kph);
        tc_fence_after();
        MT_MARK(2);
        const uint64_t kdesc = make_sdesc(smem_u32(smem_k + ks * kTileBytes), 16, 1024);
        const uint32_t d = tmem_base + ((gg & 1u) ? Cfg::kSCol1 : Cfg::kSCol0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_ss_w(lead, d, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit_w(lead, &k_empty[ks]);
        umma_commit_w(lead, &s_full[gg & 1u]);
        if (++ks == S) { ks = 0; kph ^= 1u; }
      };
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        MT_MARK(7);
        mbar_wait(q_full, it & 1u);
        tc_fence_after();
        MT_MARK(4);
        // prologue: S(0) and S(1) (both buffers are free: the previous item's PVs precede them in the in-order pipe)
        issue_s(g);
        if (p.n_kv > 1) issue_s(g + 1);
        if (p.n_kv <= 2) umma_commit_w(lead, q_empty);
        for (int j = 0; j < p.n_kv; ++j, ++g) {
          MT_MARK(7);
          mbar_wait(&p_ready[g & 1u], (g >> 1) & 1u);
          MT_MARK(g & 1u);
          if (j == 0) mbar_wait(o_empty, (it & 1u) ^ 1u);
          mbar_wait(&v_full[vs], vph);
          MT_MARK(3);
          tc_fence_after();
          const uint32_t p_tmem = tmem_base + ((g & 1u) ? Cfg::kSCol1 : Cfg::kSCol0);
          const uint32_t v_addr = smem_u32(smem_v + vs * NV * kTileBytes);
#pragma unroll
          for (int k = 0; k < TK / 16; ++k) {
            // B: 16 keys = two 8-row groups (SBO 1024 B); branches = 64-wide N atoms 16 KB apart (LBO)
            const uint64_t vdesc = make_sdesc(v_addr + k * 2048, kTileBytes, 1024);
            umma_ts_w(lead, tmem_base + Cfg::kOCol, p_tmem + k * 8, vdesc, idesc_o, (j | k) != 0 ? 1u : 0u);
          }
          umma_commit_w(lead, &v_empty[vs]);
          umma_commit_w(lead, &pv_done[g & 1u]);
          if (++vs == S) { vs = 0; vph ^= 1u; }
          if (j + 2 < p.n_kv) {  // S(j+2) re-uses this buffer right behind PV(j) in the (in-order) tensor pipe
            issue_s(g + 2);
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
    // Two groups (grp = 0: warps 4-7, grp = 1: warps 8-11) of one thread per query row; group b processes the key
    // tiles that live in S / P buffer b (global tile counter g with (g & 1) == b), so consecutive tiles of a row are
    // handled by two different threads that run half a tile period apart.  The row's running max travels with the
    // tiles: the thread of tile j receives m_{j-1} from the thread of tile j-1, decides m_j BEFORE its exponentials
    // (raised only when this tile's max exceeds m_{j-1} by more than 2^8; O is then rescaled after PV(j-1)) and
    // sends m_j on.  Each thread keeps the partial row sum of its own tiles in the scale of the latest max it knows;
    // the two partial sums are added in the epilogue.
    // the 128 scores of a row live in registers: take the registers the control warps gave up (168 -> 224 per thread)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int qd = warp & 3;             // TMEM lane quarter this warp may access
    const int grp = (warp - 4) >> 2;     // softmax group == S / P buffer it serves
    [[maybe_unused]] const int half = grp;  // (timers)
    const int r = qd * 32 + lane;        // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t ob = tmem_base + Cfg::kOCol + lane_off;
    const uint32_t sb = tmem_base + (grp ? Cfg::kSCol1 : Cfg::kSCol0) + lane_off;
    // hand-over channel between the two threads of a row: per-group outboxes, double-buffered by the sender's message
    // count; strictly alternating protocol (m_0, m_1, ..., m_last, then the row sums both ways)
    uint32_t n_sent = 0, n_rcvd = 0;
    auto send = [&](float v) {
      xch[((n_sent & 1u) * 2 + grp) * TQ + r] = v;
      mbar_arrive(&xbar[((n_sent & 1u) * 2 + grp) * 4 + qd]);
      ++n_sent;
    };
    auto recv = [&]() -> float {
      mbar_wait(&xbar[((n_rcvd & 1u) * 2 + (grp ^ 1)) * 4 + qd], (n_rcvd >> 1) & 1u);
      const float v = xch[((n_rcvd & 1u) * 2 + (grp ^ 1)) * TQ + r];
      ++n_rcvd;
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
      float m_known = 0.f, l = 0.f;  // latest running max this thread knows / partial row sum of its own tiles
      AT_MARK(6);
      for (int j = 0; j < p.n_kv; ++j, ++g) {
        if ((g & 1u) != static_cast<uint32_t>(grp)) continue;  // the other group's tile
        mbar_wait(&s_full[grp], (g >> 1) & 1u);
        tc_fence_after();
        AT_MARK(0);
        float s[128];
        {
          uint32_t* su = reinterpret_cast<uint32_t*>(s);
          tmem_ld32(sb + 0, *reinterpret_cast<uint32_t(*)[32]>(su + 0));
          tmem_ld32(sb + 32, *reinterpret_cast<uint32_t(*)[32]>(su + 32));
          tmem_ld32(sb + 64, *reinterpret_cast<uint32_t(*)[32]>(su + 64));
          tmem_ld32(sb + 96, *reinterpret_cast<uint32_t(*)[32]>(su + 96));
          tmem_ld_wait();
        }
        AT_MARK(1);
        // masking: key tail (rows mode / long-F frames mode), sequence separation (packed frames mode)
        if (strided_mask) {
#pragma unroll
          for (int c = 0; c < 128; ++c) s[c] = ((c & ppt_mask) == mine) ? s[c] : -INFINITY;
        } else {
          const int kv_valid = p.F - j * TK;
          if (kv_valid < TK) {
#pragma unroll
            for (int c = 0; c < 128; ++c) s[c] = c < kv_valid ? s[c] : -INFINITY;
          }
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
        AT_MARK(2);
        // running max of this tile, decided before the exponentials and handed on at once
        float m, m_prev = 0.f;
        bool need = false;
        if (j == 0) {
          m = (rmax == -INFINITY) ? 0.f : rmax;
        } else {
          m_prev = recv();
          need = rmax > m_prev + kRescaleThreshold;
          m = need ? rmax : m_prev;
        }
        send(m);
        AT_MARK(4);
        if (l != 0.f && m != m_known) l *= ex2_approx(m_known - m);  // own partial sum follows the running max
        m_known = m;
        if (j > 0 && __any_sync(0xffffffffu, need)) {
          // O holds the contributions of tiles < j in the scale of m_prev: wait for PV(j-1), rescale this row
          mbar_wait(&pv_done[grp ^ 1], ((g - 1) >> 1) & 1u);
          tc_fence_after();
          const float f = need ? ex2_approx(m_prev - m) : 1.0f;
#pragma unroll
          for (int c = 0; c < Cfg::kOCols; c += 32) {
            uint32_t o[32];
            tmem_ld32(ob + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st32(ob + c, o);
          }
        }
        // P = exp2(s * scale_log2 - m) (fp16, two keys per TMEM column, written over S)
        // packed fp32x2 arithmetic (FFMA2 / FADD2: two keys per issue slot) for the scale-subtract and the row sum, and three of
        // every eight key pairs through the packed FMA-pipe polynomial (ex2_poly2, max rel. error 7.5e-5): 640 instead of 1024
        // MUFU cycles per key tile and sub-partition — the same exponential path as the two-query-tile kernel (attention2q)
        {
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
            tmem_st16(sb + (c0 >> 1), pk);  // all 128 scores are already in registers: safe to overwrite S
          }
          l += (la.x + la.y) + (lb.x + lb.y);
        }
        AT_MARK(3);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_ready[grp]);
        AT_MARK(5);
      }
      // ---- epilogue: final max to the thread that did not see the last tile, row sums both ways, O / l -> global
      const uint32_t gl = g - 1;  // last tile of the item
      if ((gl & 1u) != static_cast<uint32_t>(grp)) {
        const float m_final = recv();
        if (l != 0.f && m_final != m_known) l *= ex2_approx(m_known - m_final);
        m_known = m_final;
      }
      send(l);
      const float inv_l = 1.0f / (l + recv());
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
      for (int br = 0; br < NV; ++br) {  // each of the row's two threads writes 32 of every branch's 64 columns
        __half* dst = p.o + br * p.o_branch_stride + row * p.ldo + h * HD + grp * 32;
        uint32_t o[32];
        tmem_ld32(ob + br * 64 + grp * 32, o);
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

  // plain attention in rows mode (spatial self-attention of the non-injected steps / sites, cross-attention): the two-query-
  // tile kernel of attention2q_tcgen05.cu (measured 1.19x ... 1.25x this file's NV = 1 kernel, profiles/r02_probe.txt)
  if (a->seq_mode == AV2V_SEQ_ROWS && a->n_v == 1) return attn2q_launch(a, stream);

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
