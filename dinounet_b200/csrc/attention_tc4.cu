// tcgen05 / TMEM flash attention for the DINOv3 ViT, fourth generation (head_dim 64 / 128, non-causal, ntok = 1029 at 512^2).
// Replaces F.scaled_dot_product_attention at dinounet/dinov3/layers/attention.py:116 (math: :106-118).
//
// The third generation (attention_tc3.cu) showed in its ncu source view that the softmax warps spend 21 % of their stall
// samples waiting for S(j+1) and 8 % for PV(j): with ONE S buffer (TMEM) and ONE P buffer (smem) per query tile the
// tensor pipe and the softmax warps hand a single token back and forth.  Here the key chunk is 64 wide instead of 128, so
// that BOTH hand-over buffers fit twice:
//   TMEM per query-tile group: S0 [0,64) | S1 [64,128) | O [128,128+HD)        (2 groups x 192 = 384 of 512 columns)
//   smem per group           : P0, P1 = 2 x [128 rows x 64 keys] 16-bit = 2 x 16 KB (same 32 KB as one 128-key tile)
// and every softmax warp owns 32 score columns of its 32 rows: one tcgen05.ld, 32 live scores, no second pass.
//   S(j+2) is issued as soon as the softmax warps have LOADED S(j)   -> S runs two chunks ahead
//   P(j)   may be written as soon as PV(j-2) has retired             -> the exponentials of chunk j overlap PV(j-1)
// Kept from the third generation: O accumulates in TMEM over the whole item with a lazily updated reference maximum
// (rescale only when it grows by more than 2^8; the rare rescale additionally waits for PV(j-1)), the last key chunk is
// only as wide as it must be (ntok = 1029 -> N = 16), fully out-of-range query-row quarters only run the barrier
// protocol, items are ordered long-first.
// Warps: 0 = TMA producer, 1 / 2 = MMA issuer of group A / B, 3 idle, 4.. = softmax (2 per (group, TMEM lane quarter)).
#include <type_traits>

#include "common.cuh"
#include "../../include/dinounet_b200.h"
#include "host_util.h"
#include "gemm_common.h"
#include "attention_common.h"

namespace b2u {

template <int HD> struct At4Cfg {
  static constexpr int kGroups = HD == 64 ? 2 : 1;
  static constexpr int kSplit = 2;
  static constexpr int kKC = 64;                            // keys per chunk
  static constexpr int kStages = HD == 64 ? 5 : 4;          // K/V ring
  static constexpr int kKB = HD / 64;                       // 64-wide K blocks of the head dim
  static constexpr int kQBytes = 128 * HD * 2;              // kKB blocks of [128 rows x 64]
  static constexpr int kPBytes = 128 * kKC * 2;             // one P buffer: [128 rows x 64 keys]
  static constexpr int kKBytes = kKC * HD * 2;              // kKB blocks of [64 keys x 64]
  static constexpr int kVBytes = HD * kKC * 2;              // [HD rows x 64 keys]
  static constexpr int kCtrlWarps = 4;                      // TMA producer, MMA issuer A, MMA issuer B, (idle)
  static constexpr int kThreads = kCtrlWarps * 32 + kGroups * 128 * kSplit;
  static constexpr int kXchgBytes = 2 * 128 * 4 * 4;        // [2 groups][128 rows][4] fp32 exchange slots
  static constexpr int kGroupCols = 128 + HD;               // S0 | S1 | O
  static constexpr int kBarBytes = 512;
  static constexpr int kSmem = kGroups * (kQBytes + 2 * kPBytes) + kStages * (kKBytes + kVBytes) + 1024 + kBarBytes + kXchgBytes;
};

__device__ __forceinline__ void tmem_st32_4(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait4() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// item -> (bh, pair, number of query tiles in it).  Full pairs first, leftover single tiles last.
struct ItemDec4 {
  int bh, pair, nq;
};
template <int NG>
__device__ __forceinline__ ItemDec4 decode_item4(long long item, const AttnArgs& a) {
  ItemDec4 d;
  if (NG == 1) {
    d.bh = static_cast<int>(item / a.npairs);
    d.pair = static_cast<int>(item - static_cast<long long>(d.bh) * a.npairs);
    d.nq = 1;
    return d;
  }
  const long long nfull = static_cast<long long>(a.BH) * a.pairs_full;
  if (item < nfull) {
    d.bh = static_cast<int>(item / a.pairs_full);
    d.pair = static_cast<int>(item - static_cast<long long>(d.bh) * a.pairs_full);
    d.nq = 2;
  } else {
    d.bh = static_cast<int>(item - nfull);
    d.pair = a.pairs_full;
    d.nq = 1;
  }
  return d;
}

// Row maximum over one 32-column group of raw scores; only the last key chunk (kTail) has columns >= lim to skip.
// Four independent accumulators instead of one 32-deep dependent FMNMX chain.
template <bool kTail>
__device__ __forceinline__ float row_max4(const uint32_t (&v)[32], int lim) {
  float a[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    if (kTail) {
      if (c < lim) a[c & 3] = fmaxf(a[c & 3], __uint_as_float(v[c]));
    } else {
      a[c & 3] = fmaxf(a[c & 3], __uint_as_float(v[c]));
    }
  }
  return fmaxf(fmaxf(a[0], a[1]), fmaxf(a[2], a[3]));
}

// 32 columns: P = exp2(s*scale - m) -> 16-bit -> 128B-swizzled smem row (A operand of the PV MMA); returns the fp32 row
// sum of the un-rounded probabilities (as flash-attention).  `half` = which 64-byte half of the 128-byte row.
template <typename TT, bool kTail>
__device__ __forceinline__ float exp_store4(const uint32_t (&v)[32], int lim, float sl2, float m_new, int half,
                                            uint32_t sP_row, int row) {
  float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {       // 8 probabilities -> one 16-byte store
    uint32_t pk[4];
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      const int cc = c4 * 8 + c;
      float a = ex2(fmaf(__uint_as_float(v[cc]), sl2, -m_new));
      float b = ex2(fmaf(__uint_as_float(v[cc + 1]), sl2, -m_new));
      if constexpr (kTail) {
        if (cc >= lim) a = 0.f;
        if (cc + 1 >= lim) b = 0.f;
      }
      rs4[c >> 1] += a + b;
      pk[c >> 1] = TT::pack2(a, b);
    }
    const int chunk = half * 4 + c4;
    sts128a(sP_row + ((chunk ^ (row & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
  }
  return (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
}

template <typename T, int HD>
__global__ void __launch_bounds__(At4Cfg<HD>::kThreads, 1) attn_tc4_kernel(const __grid_constant__ AttnMaps maps, const AttnArgs args) {
  using TT = T16<T>;
  using CF = At4Cfg<HD>;
  constexpr int NG = CF::kGroups, NST = CF::kStages, KB = CF::kKB, SPLIT = CF::kSplit, KC = CF::kKC;
  constexpr int kArr = 4 * SPLIT;                    // one arrival per softmax warp of a group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // [NG][kQBytes]
  uint8_t* sP = sQ + NG * CF::kQBytes;              // [NG][2][kPBytes]
  uint8_t* sK = sP + NG * 2 * CF::kPBytes;          // [stages][kKBytes]
  uint8_t* sV = sK + NST * CF::kKBytes;             // [stages][kVBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NST * CF::kVBytes);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_free = q_full + 2;           // [2]
  uint64_t* kv_full = q_free + 2;          // [8]
  uint64_t* kv_empty = kv_full + 8;        // [8]
  uint64_t* s_full = kv_empty + 8;         // [group][buffer]
  uint64_t* s_free = s_full + 4;           // [group][buffer]
  uint64_t* p_full = s_free + 4;           // [group][buffer]
  uint64_t* o_full = p_full + 4;           // [group][buffer]  PV(j) retired (j & 1 = buffer)
  uint64_t* o_free = o_full + 4;           // [group]          the item's O has been read out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + CF::kBarBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.vt);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&q_full[g], 1); mbar_init(&q_free[g], 1); mbar_init(&o_free[g], kArr);
      for (int b = 0; b < 2; ++b) {
        mbar_init(&s_full[2 * g + b], 1); mbar_init(&s_free[2 * g + b], kArr);
        mbar_init(&p_full[2 * g + b], kArr); mbar_init(&o_full[2 * g + b], 1);
      }
    }
    for (int s = 0; s < 8; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], NG); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int J = args.nchunks;                                // 64-key chunks
  const int tail_valid = args.ntok - (J - 1) * KC;           // valid keys of the last chunk, 1..64
  const int tail_n = (tail_valid + 15) & ~15;                // MMA width of the last chunk (N of S, K of PV)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t kv_phase = 0;
      uint32_t qfree_cnt[2] = {0, 0};
      for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
        const ItemDec4 d = decode_item4<NG>(item, args);
        for (int g = 0; g < d.nq; ++g) {
          const int q0 = args.q_begin + (d.pair * NG + g) * 128;
          mbar_wait(&q_free[g], (qfree_cnt[g] & 1) ^ 1);
          ++qfree_cnt[g];
          mbar_expect_tx(&q_full[g], CF::kQBytes);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(sQ + g * CF::kQBytes + kb * (128 * 128), &maps.q, &q_full[g], kb * 64, q0, d.bh);
        }
        for (int j = 0; j < J; ++j) {
          mbar_wait(&kv_empty[stage], kv_phase ^ 1);
          mbar_expect_tx(&kv_full[stage], CF::kKBytes + CF::kVBytes);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(sK + stage * CF::kKBytes + kb * (KC * 128), &maps.k, &kv_full[stage], kb * 64, j * KC, d.bh);
          tma_load_3d(sV + stage * CF::kVBytes, &maps.vt, &kv_full[stage], j * KC, 0, d.bh);
          if (++stage == NST) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
  } else if (warp <= NG) {
    // ===================== MMA issuer of group g (one thread per group, warps 1 and 2) =====================
    // In-order stream per group: S(0), S(1), then per chunk j: S(j+2) (its conditions - K chunk j+2 landed, S(j) loaded by
    // the softmax warps - always come true before P(j) is complete), PV(j).  A single issuer thread polling both groups
    // was measured to be the bottleneck at 64-key chunks (35 % of the softmax stall samples waiting for S).
    // Chunk j uses S / P buffer j & 1; every per-buffer barrier advances one phase per use and cannot run two phases ahead
    // of its waiter (a buffer is re-armed only after its previous use has been consumed): parity waits are unambiguous.
    if (elect_one()) {
      const int g = warp - 1;
      constexpr uint32_t idesc_s = make_idesc_f16(TT::kFmt, 128, KC);
      constexpr uint32_t idesc_o = make_idesc_f16(TT::kFmt, 128, HD);
      const uint32_t idesc_s_tail = make_idesc_f16(TT::kFmt, 128, tail_n);
      const uint32_t tG = tmem_base + g * CF::kGroupCols;
      int st_s = 0, st_p = 0;                 // ring position of the next S chunk / the next PV chunk
      uint32_t ph_s = 0, ph_p = 0;
      uint32_t use0 = 0, use1 = 0;            // uses of buffer 0 / 1 (S and P alike: once per chunk) before this item
      uint32_t n_items = 0, qcnt = 0;
      uint64_t dq[KB];
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) dq[kb] = make_desc_k128(smem_u32(sQ + g * CF::kQBytes + kb * (128 * 128)));
      auto issue_s = [&](int j) {
        const int b = j & 1;
        const uint32_t n = (b ? use1 : use0) + (j >> 1);
        mbar_wait(&kv_full[st_s], ph_s);
        mbar_wait(&s_free[2 * g + b], (n & 1) ^ 1);           // the softmax warps have loaded the previous occupant
        tc_fence_after();
        const bool last = j + 1 == J;
        const uint32_t idesc = last ? idesc_s_tail : idesc_s;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t db = make_desc_k128(smem_u32(sK + st_s * CF::kKBytes + kb * (KC * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tG + b * KC, dq[kb] + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&s_full[2 * g + b]);
        if (last) tc_commit(&q_free[g]);                       // no later MMA of this item reads Q
        if (++st_s == NST) { st_s = 0; ph_s ^= 1; }
      };
      for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
        const ItemDec4 d = decode_item4<NG>(item, args);
        if (g >= d.nq) {
          // this group sits the item out, but the K/V ring counts one release per group and stage
          for (int j = 0; j < J; ++j) {
            mbar_wait(&kv_full[st_p], ph_p);
            tc_commit(&kv_empty[st_p]);
            if (++st_p == NST) { st_p = 0; ph_p ^= 1; }
          }
          st_s = st_p; ph_s = ph_p;
          continue;
        }
        mbar_wait(&q_full[g], qcnt & 1);
        ++qcnt;
        issue_s(0);
        if (J > 1) issue_s(1);
        for (int j = 0; j < J; ++j) {
          if (j + 2 < J) issue_s(j + 2);
          const int b = j & 1;
          const uint32_t n = (b ? use1 : use0) + (j >> 1);
          mbar_wait(&p_full[2 * g + b], n & 1);                // P_g(j) is in smem
          if (j == 0) mbar_wait(&o_free[g], (n_items & 1) ^ 1);   // the previous item's O has been read out
          tc_fence_after();
          const int nk = (j + 1 == J) ? (tail_n >> 4) : KC / 16;  // K = 16 steps of this chunk
          const uint64_t da = make_desc_k128(smem_u32(sP + (g * 2 + b) * CF::kPBytes));
          const uint64_t db = make_desc_k128(smem_u32(sV + st_p * CF::kVBytes));
          for (int k = 0; k < nk; ++k)
            tc_mma_f16(tG + 128, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc_o, (j | k) != 0 ? 1u : 0u);
          tc_commit(&o_full[2 * g + b]);
          tc_commit(&kv_empty[st_p]);                          // K_j / V_j are free once both groups' MMAs have retired
          if (++st_p == NST) { st_p = 0; ph_p ^= 1; }
        }
        ++n_items;
        use0 += (J + 1) >> 1;
        use1 += J >> 1;
      }
    }
  } else if (warp >= CF::kCtrlWarps) {
    // ===================== softmax / output warps =====================
    const int set = (warp - CF::kCtrlWarps) >> 2;
    const int g = set / SPLIT;              // query tile of the pair
    const int part = set % SPLIT;           // column / head-dim slice of this warp
    const int q4 = warp & 3;                // TMEM lane quarter this warp may access
    const int row = q4 * 32 + lane;
    constexpr int OW = HD / SPLIT;          // O columns (head dims) per warp
    const uint32_t tS = tmem_base + g * CF::kGroupCols + (static_cast<uint32_t>(q4 * 32) << 16) + part * 32;
    const uint32_t tO = tmem_base + g * CF::kGroupCols + 128 + (static_cast<uint32_t>(q4 * 32) << 16) + part * OW;
    const uint32_t sP_g = smem_u32(sP + g * 2 * CF::kPBytes);
    const uint32_t sP_row = sP_g + row * 128;
    float* xm = xchg + (g * 128 + row) * 4;  // [2 buffers][2 parts] row maxima; the l exchange reuses the slots
    const int bar_id = 1 + g * 4 + q4;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };
    uint64_t* const sfb = &s_full[2 * g];
    uint64_t* const sfr = &s_free[2 * g];
    uint64_t* const pfb = &p_full[2 * g];
    uint64_t* const ofb = &o_full[2 * g];
    // uses of S buffer b / PV completions on buffer b before this item (scalars: a dynamically indexed array would live in
    // local memory)
    uint32_t use_s0 = 0, use_s1 = 0, use_o0 = 0, use_o1 = 0;
    const float sl2 = args.scale_log2e;
    for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
      const ItemDec4 d = decode_item4<NG>(item, args);
      if (g >= d.nq) continue;              // this group has no query tile in this item (warp-uniform)
      const int q0 = args.q_begin + (d.pair * NG + g) * 128;
      const bool dead = q0 + q4 * 32 >= args.ntok;   // none of this warp's 32 rows exists: barrier protocol only
      int waited = -1;                      // PV(0..waited) of this item are known to have retired
      auto wait_pv = [&](int c) {           // in order, one phase per chunk and buffer
        while (waited < c) {
          ++waited;
          mbar_wait(&ofb[waited & 1], (((waited & 1) ? use_o1 : use_o0) + (waited >> 1)) & 1);
        }
      };
      if (dead) {
        // keep the barrier protocol in lock step (arrival counts include every warp of the group), touch no data; the
        // P rows of these query rows hold garbage, which only reaches O rows that are never stored
        for (int j = 0; j < J; ++j) {
          const int b = j & 1;
          mbar_wait(&sfb[b], ((b ? use_s1 : use_s0) + (j >> 1)) & 1);
          if (lane == 0) mbar_arrive(&sfr[b]);
          if (j >= 2) wait_pv(j - 2);
          if (lane == 0) mbar_arrive(&pfb[b]);
        }
        wait_pv(J - 1);
        if (lane == 0) mbar_arrive(&o_free[g]);
      } else {
        float m = -INFINITY, l = 0.f;
        for (int j = 0; j < J; ++j) {
          const int b = j & 1;
          const bool tail = j + 1 == J;
          const int lim = tail_valid - part * 32;                // (tail) columns [0, lim) of this warp's group are real keys
          const bool mine = !tail || tail_n > part * 32;         // this warp owns columns of this chunk (warp-uniform)
          mbar_wait(&sfb[b], ((b ? use_s1 : use_s0) + (j >> 1)) & 1);
          tc_fence_after();
          uint32_t v[32];
          tmem_ld32(tS + b * KC, v);                             // (tail: columns >= tail_n are stale, masked below)
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sfr[b]);                   // S(j) is in registers: S(j+2) may be produced
          float mx = -INFINITY;
          if (mine) mx = tail ? row_max4<true>(v, lim) : row_max4<false>(v, 32);
          xm[b * 2 + part] = mx;
          pair_sync();
          mx = fmaxf(mx, xm[b * 2 + (part ^ 1)]);
          const float m_cand = fmaxf(m, mx * sl2);               // chunk 0 always has valid keys -> finite
          const bool need = (m_cand - m) > 8.0f;                 // lazy: keep the old reference while exp2 stays <= 2^8
          float m_new = m, corr = 1.f;
          if (need) { m_new = m_cand; corr = ex2(m - m_new); }
          if (j >= 2) wait_pv(j - 2);                            // PV(j-2) retired: P buffer b is free
          float rs = 0.f;
          if (mine)
            rs = tail ? exp_store4<TT, true>(v, lim, sl2, m_new, part, sP_row + b * CF::kPBytes, row)
                      : exp_store4<TT, false>(v, 32, sl2, m_new, part, sP_row + b * CF::kPBytes, row);
          // ---- rare: the reference maximum moved -> rescale this warp's slice of O in TMEM.  No PV MMA may be in
          // flight: PV(j-1) must have retired (PV(j) waits for this warp's p_full arrival).
          if (j > 0 && __any_sync(0xffffffffu, need)) {
            wait_pv(j - 1);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < OW / 32; ++h) {
              uint32_t o[32];
              tmem_ld32(tO + h * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * corr);
              tmem_st32_4(tO + h * 32, o);
            }
            tmem_st_wait4();
          }
          l = l * corr + rs;
          m = m_new;
          fence_proxy_async();                     // make the generic-proxy P writes visible to the MMA (async proxy)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&pfb[b]);
        }
        // ---- all PVs retired: read this warp's O slice, release the accumulator for the next item
        wait_pv(J - 1);
        tc_fence_after();
        uint32_t o[OW / 32][32];
#pragma unroll
        for (int h = 0; h < OW / 32; ++h) tmem_ld32(tO + h * 32, o[h]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[g]);
        // ---- row sums of the column slices add up (same reference maximum in both warps of the row)
        pair_sync();                                   // the partner has consumed the last row-max slots
        xm[part] = l;
        pair_sync();
        l += xm[part ^ 1];
        // ---- normalise, stage this warp's head dims of its 32 rows through (now free) P smem: [128 rows x 64 dims] per
        // 16 KB block = P buffer hb of this group
        const float inv = 1.f / l;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < OW / 8; ++c) {
          const int dim0 = part * OW + 8 * c;          // first head dim of this 16-byte chunk
#define B2U_O(k_) (__uint_as_float(o[(8 * c + (k_)) >> 5][(8 * c + (k_)) & 31]) * inv)
          sts128a(sP_row + (dim0 >> 6) * CF::kPBytes + ((((dim0 >> 3) & 7) ^ (row & 7)) << 4), TT::pack2(B2U_O(0), B2U_O(1)),
                  TT::pack2(B2U_O(2), B2U_O(3)), TT::pack2(B2U_O(4), B2U_O(5)), TT::pack2(B2U_O(6), B2U_O(7)));
#undef B2U_O
        }
        __syncwarp();
        pair_sync();                                   // both dim slices of these 32 rows are staged
        const int bb = d.bh / args.heads, hd = d.bh - bb * args.heads;
        const int D = args.heads * HD;
        T* outp = reinterpret_cast<T*>(args.out);
#pragma unroll
        for (int hb = 0; hb < HD / 64; ++hb)
#pragma unroll
          for (int ii = 0; ii < 8 / SPLIT; ++ii) {     // the warps of a row quarter split its 32 rows
            const int i = part * (8 / SPLIT) + ii;
            const int rr = q4 * 32 + i * 4 + (lane >> 3);
            const uint4 val = lds128a(sP_g + hb * CF::kPBytes + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            const int t = q0 + rr;
            if (t < args.ntok)
              *reinterpret_cast<uint4*>(outp + (static_cast<long long>(bb) * args.ntok + t) * D + hd * HD + hb * 64 + (lane & 7) * 8) = val;
          }
        __syncwarp();
        pair_sync();                                   // the partner has read my staged chunks: P smem may be rewritten
      }
      use_s0 += (J + 1) >> 1; use_s1 += J >> 1;
      use_o0 += (J + 1) >> 1; use_o1 += J >> 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <typename T, int HD>
static int launch_attn_tc4(const AttnMaps& maps, const AttnArgs& a, cudaStream_t stream) {
  auto kern = attn_tc4_kernel<T, HD>;
  using CF = At4Cfg<HD>;
  static_assert(CF::kSmem <= 227 * 1024, "attention smem budget");
  static_assert(CF::kStages <= 8, "barrier arrays");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kSmem);
    if (e != cudaSuccess) return set_error(-2, "cudaFuncSetAttribute(attn_tc4): %s", cudaGetErrorString(e));
    configured = true;
  }
  const int sms = num_sms();
  const int grid = static_cast<int>(a.items < sms ? a.items : sms);
  kern<<<grid, CF::kThreads, CF::kSmem, stream>>>(maps, a);
  return check_launch("attention_tc4");
}

int attention_tc4_dispatch(const AttnMaps& maps, const AttnArgs& a, int head_dim, int dtype, cudaStream_t stream) {
  if (head_dim == 64)
    return dtype == B2U_BF16 ? launch_attn_tc4<__nv_bfloat16, 64>(maps, a, stream) : launch_attn_tc4<__half, 64>(maps, a, stream);
  return dtype == B2U_BF16 ? launch_attn_tc4<__nv_bfloat16, 128>(maps, a, stream) : launch_attn_tc4<__half, 128>(maps, a, stream);
}

}  // namespace b2u
