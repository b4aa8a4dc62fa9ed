// tcgen05 / TMEM flash attention for the DINOv3 ViT (head_dim 64, non-causal, ntok = 1029 at 512^2).
//
// Persistent CTAs (1 per SM, 320 threads); one work item = (batch*head, pair of 128-row query tiles):
//   warp 0    : TMA producer — Q tiles (once per item), K chunk [128 keys x 64] + V^T chunk [64 x 128 keys], 3-stage ring
//   warp 1    : MMA issuer   — S_g(j) = Q_g K_j^T (M128 N128 K64, 4 x tcgen05.mma) into TMEM buffer X_g[j&1],
//                              O_g(j) = P_g V_j   (M128 N64 K128, 8 x tcgen05.mma) into columns [0,64) of the same buffer
//                              (dead once the softmax has read S_g(j)), so S(j+1) is produced while softmax(j) runs
//   warps 2-5 : softmax group A (query tile 0), warps 6-9: softmax group B (query tile 1), one thread per query row:
//               pass 1 tcgen05.ld S -> row max; pass 2 tcgen05.ld S -> exp2 -> 16-bit P into 128B-swizzled smem
//               (the A operand of the PV MMA); then absorb the previous chunk's P V product from TMEM into fp32
//               registers with the online-softmax correction.  While group A does softmax the tensor core works for B.
// Every MMA operand is K-major SW128 (the layout the GEMM kernel already uses): V is consumed as V^T [B,H,64,npad]
// (written transposed by the QKV epilogue), so no MN-major descriptors are needed.  TMEM: 2 groups x 2 buffers x 128 = 512 cols.
// Replaces F.scaled_dot_product_attention at dinounet/dinov3/layers/attention.py:116.
#include <type_traits>

#include "common.cuh"
#include "../../include/dinounet_b200.h"
#include "host_util.h"
#include "gemm_common.h"
#include "attention_common.h"

namespace b2u {

// Per-head-dim configuration.  HD = 64 (ViT-S/B/L): two query-tile groups per CTA, 4-stage K/V ring.
// HD = 128 (ViT-7B): one group (the fp32 O accumulator needs 128 registers per thread), 2-stage ring.
template <int HD, int SPLIT = 1> struct AtCfg {
  static constexpr int kGroups = HD == 64 ? 2 : 1;
  static constexpr int kStages = HD == 64 ? (SPLIT > 1 ? 3 : 4) : 2;   // SPLIT > 1 pays for its exchange slots with a stage
  static constexpr int kKB = HD / 64;                       // 64-wide K blocks of the head dim
  static constexpr int kQBytes = 128 * HD * 2;              // kKB blocks of [128 rows x 64]
  static constexpr int kPBytes = 2 * 128 * 128;             // two key blocks of [128 rows x 64 keys]
  static constexpr int kKBytes = 128 * HD * 2;              // kKB blocks of [128 keys x 64]
  static constexpr int kVBytes = 2 * HD * 128;              // two key blocks of [HD rows x 64 keys]
  static constexpr int kThreads = 64 + kGroups * 128 * SPLIT;
  static constexpr int kXchgBytes = SPLIT > 1 ? 2 * 128 * 4 * 4 : 0;   // [2 groups][128 rows][4] fp32 exchange slots
  static constexpr int kSmem = kGroups * (kQBytes + kPBytes) + kStages * (kKBytes + kVBytes) + 1024 + 256 + kXchgBytes;
};

template <typename T, int HD, int SPLIT>
__global__ void __launch_bounds__(AtCfg<HD, SPLIT>::kThreads, 1) attn_tc_kernel(const __grid_constant__ AttnMaps maps, const AttnArgs args) {
  using TT = T16<T>;
  using CF = AtCfg<HD, SPLIT>;
  constexpr int NG = CF::kGroups, NST = CF::kStages, KB = CF::kKB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                               // [NG][kQBytes]
  uint8_t* sP = sQ + NG * CF::kQBytes;              // [NG][32 KB]
  uint8_t* sK = sP + NG * CF::kPBytes;              // [stages][kKBytes]
  uint8_t* sV = sK + NST * CF::kKBytes;             // [stages][kVBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NST * CF::kVBytes);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_free = q_full + 2;           // [2]
  uint64_t* kv_full = q_free + 2;          // [4]
  uint64_t* kv_empty = kv_full + 4;
  uint64_t* s_full = kv_empty + 4;         // [2 groups][2 buffers]
  uint64_t* x_free = s_full + 4;           // [2][2]
  uint64_t* p_full = x_free + 4;           // [2]
  uint64_t* o_full = p_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  float* xchg = reinterpret_cast<float*>(bars + 32);   // 256 B of barriers, then the exchange slots

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.q);
    tma_prefetch_desc(&maps.k);
    tma_prefetch_desc(&maps.vt);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&q_full[g], 1); mbar_init(&q_free[g], 1);
      mbar_init(&s_full[2 * g], 1); mbar_init(&s_full[2 * g + 1], 1);
      mbar_init(&x_free[2 * g], 128 * SPLIT); mbar_init(&x_free[2 * g + 1], 128 * SPLIT);
      mbar_init(&p_full[g], 128 * SPLIT); mbar_init(&o_full[g], 1);
    }
    for (int s = 0; s < 4; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: X_g[b] at g*256 + b*128 (S_g(j) for j&1 == b, then O_g(j) in its first HD columns)
  const int J = args.nchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t kv_phase = 0;
      uint32_t qfree_cnt[2] = {0, 0};
      for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
        const int bh = static_cast<int>(item / args.npairs);
        const int pair = static_cast<int>(item - static_cast<long long>(bh) * args.npairs);
        for (int g = 0; g < NG; ++g) {
          const int q0 = args.q_begin + (pair * NG + g) * 128;
          if (q0 >= args.ntok) continue;
          mbar_wait(&q_free[g], (qfree_cnt[g] & 1) ^ 1);
          ++qfree_cnt[g];
          mbar_expect_tx(&q_full[g], CF::kQBytes);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(sQ + g * CF::kQBytes + kb * (128 * 128), &maps.q, &q_full[g], kb * 64, q0, bh);
        }
        for (int j = 0; j < J; ++j) {
          mbar_wait(&kv_empty[stage], kv_phase ^ 1);
          mbar_expect_tx(&kv_full[stage], CF::kKBytes + CF::kVBytes);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_3d(sK + stage * CF::kKBytes + kb * (128 * 128), &maps.k, &kv_full[stage], kb * 64, j * 128, bh);
          tma_load_3d(sV + stage * CF::kVBytes, &maps.vt, &kv_full[stage], j * 128, 0, bh);
          tma_load_3d(sV + stage * CF::kVBytes + HD * 128, &maps.vt, &kv_full[stage], j * 128 + 64, 0, bh);
          if (++stage == NST) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(TT::kFmt, 128, 128);
      constexpr uint32_t idesc_o = make_idesc_f16(TT::kFmt, 128, HD);
      int stage = 0;
      uint32_t kv_phase = 0;
      uint32_t qfull_cnt[2] = {0, 0}, pfull_cnt[2] = {0, 0};
      uint32_t xfree_cnt[2][2] = {{0, 0}, {0, 0}};
      auto issue_s = [&](int g, int jj, int st) {     // S_g(jj) into X_g[jj & 1]
        const int bfr = jj & 1;
        mbar_wait(&x_free[2 * g + bfr], (xfree_cnt[g][bfr] & 1) ^ 1);   // softmax g has absorbed the previous occupant
        ++xfree_cnt[g][bfr];
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t da = make_desc_k128(smem_u32(sQ + g * CF::kQBytes + kb * (128 * 128)));
          const uint64_t db = make_desc_k128(smem_u32(sK + st * CF::kKBytes + kb * (128 * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_f16(tmem_base + g * 256 + bfr * 128, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc_s,
                       (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&s_full[2 * g + bfr]);
      };
      auto next_stage = [&](int st, uint32_t ph, int& nst, uint32_t& nph) {
        nst = st + 1; nph = ph;
        if (nst == NST) { nst = 0; nph ^= 1; }
      };
      for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
        const int bh = static_cast<int>(item / args.npairs);
        const int pair = static_cast<int>(item - static_cast<long long>(bh) * args.npairs);
        const int nq = (NG == 2 && args.q_begin + (pair * 2 + 1) * 128 < args.ntok) ? 2 : 1;
        for (int g = 0; g < nq; ++g) { mbar_wait(&q_full[g], qfull_cnt[g] & 1); ++qfull_cnt[g]; }
        // `stage`/`kv_phase` track chunk j (whose V the PV MMA uses); S runs one chunk ahead when the ring is deep enough.
        int st1; uint32_t ph1;
        next_stage(stage, kv_phase, st1, ph1);
        mbar_wait(&kv_full[stage], kv_phase);
        tc_fence_after();
        for (int g = 0; g < nq; ++g) issue_s(g, 0, stage);
        if (J > 1) {
          mbar_wait(&kv_full[st1], ph1);
          tc_fence_after();
          for (int g = 0; g < nq; ++g) issue_s(g, 1, st1);
        }
        for (int j = 0; j < J; ++j) {
          int st2; uint32_t ph2;
          next_stage(st1, ph1, st2, ph2);
          for (int g = 0; g < nq; ++g) {
            mbar_wait(&p_full[g], pfull_cnt[g] & 1);   // P_g(j) is in smem (and S_g(j) has been fully read)
            ++pfull_cnt[g];
            tc_fence_after();
            const uint32_t tO = tmem_base + g * 256 + (j & 1) * 128;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t da = make_desc_k128(smem_u32(sP + g * CF::kPBytes + kb * 128 * 128));
              const uint64_t db = make_desc_k128(smem_u32(sV + stage * CF::kVBytes + kb * HD * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_f16(tO, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc_o, (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit(&o_full[g]);
          }
          tc_commit(&kv_empty[stage]);   // K_j / V_j are free once every MMA issued so far has retired
          if (j + 2 < J) {               // S(j+2) reuses buffer j&1 once O(j) has been absorbed by the softmax
            mbar_wait(&kv_full[st2], ph2);   // (2-stage ring: st2 == stage, refilled after the commit above)
            tc_fence_after();
            for (int g = 0; g < nq; ++g) issue_s(g, j + 2, st2);
          }
          stage = st1; kv_phase = ph1;
          st1 = st2; ph1 = ph2;
        }
        for (int g = 0; g < nq; ++g) tc_commit(&q_free[g]);
      }
    }
  } else {
    // ===================== softmax / output warps =====================
    // SPLIT warps share each (query tile, TMEM lane quarter): warp `part` owns key columns [part*128/SPLIT, ...) of
    // every score tile and head dims [part*HD/SPLIT, ...) of O.  More resident warps per scheduler hide the TMEM-load /
    // MUFU / dependency latencies this loop is bound by (2 warps per scheduler stalled on "wait" + "long scoreboard"
    // 70 % of the time); the price is one 64-thread named barrier per key tile to combine the row maxima.
    const int sw = warp - 2;
    const int g = (sw >> 2) / SPLIT;        // query tile of the pair
    const int part = (sw >> 2) % SPLIT;     // column / head-dim slice of this warp
    const int q4 = warp & 3;                // TMEM lane quarter
    const int row = q4 * 32 + lane;
    constexpr int CW = 128 / SPLIT;         // score columns per warp
    constexpr int OW = HD / SPLIT;          // O columns (head dims) per warp
    const uint32_t tX = tmem_base + g * 256 + (static_cast<uint32_t>(q4 * 32) << 16);
    const uint32_t sP_row = smem_u32(sP + g * CF::kPBytes) + row * 128;
    const uint32_t sP_base = smem_u32(sP + g * CF::kPBytes);
    float* xm = xchg + (g * 128 + row) * 4;  // [2 buffers][2 parts] row maxima; the l exchange reuses the slots
    auto pair_sync = [&]() {
      if constexpr (SPLIT > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + g * 4 + q4), "n"(32 * SPLIT) : "memory");
    };
    uint32_t sfull_cnt[2] = {0, 0}, ofull_cnt = 0;
    for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
      const int bh = static_cast<int>(item / args.npairs);
      const int pair = static_cast<int>(item - static_cast<long long>(bh) * args.npairs);
      const int q0 = args.q_begin + (pair * NG + g) * 128;
      if (q0 >= args.ntok) continue;        // this group has no query tile in this item (warp-uniform)
      float m = -INFINITY, l = 0.f, corr_prev = 0.f;
      float o[OW];
#pragma unroll
      for (int i = 0; i < OW; ++i) o[i] = 0.f;
      auto absorb = [&](uint32_t t) {       // O = O * corr_prev + (P V)(chunk) read from TMEM (this warp's head dims)
#pragma unroll
        for (int h = 0; h < OW / 32; ++h) {
          uint32_t v[32];
          tmem_ld32(t + part * OW + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) o[h * 32 + c] = fmaf(o[h * 32 + c], corr_prev, __uint_as_float(v[c]));
        }
      };
      for (int j = 0; j < J; ++j) {
        const uint32_t tS = tX + (j & 1) * 128 + part * CW;
        mbar_wait(&s_full[2 * g + (j & 1)], sfull_cnt[j & 1] & 1);
        ++sfull_cnt[j & 1];
        tc_fence_after();
        const int kbase = j * 128 + part * CW;
        const bool tail = j * 128 + 128 > args.ntok;
        // ---- pass 1: row max of the raw scores over this warp's columns (scale > 0, applied once)
        float mx = -INFINITY;
#pragma unroll
        for (int pc = 0; pc < CW / 32; ++pc) {
          uint32_t v[32];
          tmem_ld32(tS + pc * 32, v);
          tmem_ld_wait();
          if (!tail) {
#pragma unroll
            for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(v[c]));
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (kbase + pc * 32 + c < args.ntok) mx = fmaxf(mx, __uint_as_float(v[c]));
          }
        }
        if constexpr (SPLIT > 1) {            // combine with the partner's columns (double-buffered slots, see above)
          xm[(j & 1) * 2 + part] = mx;
          pair_sync();
          mx = fmaxf(mx, xm[(j & 1) * 2 + (part ^ 1)]);
        }
        const float m_new = fmaxf(m, mx * args.scale_log2e);   // chunk 0 always has valid keys -> finite
        const float corr = ex2(m - m_new);
        m = m_new;
        if (j > 0) {                                // PV_{j-1} retired: P smem is free, O(j-1) sits in X[(j-1)&1][0:HD)
          mbar_wait(&o_full[g], ofull_cnt & 1);
          ++ofull_cnt;
          tc_fence_after();
          absorb(tX + ((j - 1) & 1) * 128);
          tc_fence_before();
          mbar_arrive(&x_free[2 * g + ((j - 1) & 1)]);   // buffer (j-1)&1 may now receive S(j+1)
        }
        // ---- pass 2: P = exp2(s*scale - m) -> 16-bit -> swizzled smem (A operand of the PV MMA).
        // Two compiled bodies: only the last key chunk has columns >= ntok to zero; left as a runtime test inside the
        // element loop the compiler if-converts it into an index add + compare + select PER ELEMENT of every chunk.
        float rs = 0.f;
        auto pass2 = [&](auto tail_c) {
          constexpr bool kTail = decltype(tail_c)::value;
#pragma unroll
          for (int pc = 0; pc < CW / 32; ++pc) {
            uint32_t v[32];
            tmem_ld32(tS + pc * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              float a = ex2(fmaf(__uint_as_float(v[c]), args.scale_log2e, -m_new));
              float b = ex2(fmaf(__uint_as_float(v[c + 1]), args.scale_log2e, -m_new));
              if constexpr (kTail) {
                if (kbase + pc * 32 + c >= args.ntok) a = 0.f;
                if (kbase + pc * 32 + c + 1 >= args.ntok) b = 0.f;
              }
              rs += a + b;                          // fp32 row sum of the un-rounded probabilities (as flash-attention)
              pk[c >> 1] = TT::pack2(a, b);
            }
            // tile columns gc*32 .. +31 -> K-block gc>>1, 16-byte chunks (gc&1)*4 .. +3 of this row
            const int gc = part * (CW / 32) + pc;
            const uint32_t base = sP_row + (gc >> 1) * (128 * 128);
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int chunk = (gc & 1) * 4 + c4;
              sts128a(base + ((chunk ^ (row & 7)) << 4), pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
            }
          }
        };
        if (tail) pass2(std::true_type{});
        else pass2(std::false_type{});
        tc_fence_before();                         // all reads of S_g(j) done: PV(j) may overwrite X[j&1][0:HD)
        fence_proxy_async();                       // make the generic-proxy P writes visible to the MMA (async proxy)
        mbar_arrive(&p_full[g]);
        l = l * corr + rs;
        corr_prev = corr;
      }
      // ---- last chunk
      mbar_wait(&o_full[g], ofull_cnt & 1);
      ++ofull_cnt;
      tc_fence_after();
      absorb(tX + ((J - 1) & 1) * 128);
      tc_fence_before();
      mbar_arrive(&x_free[2 * g + ((J - 1) & 1)]);
      // ---- row sums of the column slices add up (same running max in every warp of the row)
      if constexpr (SPLIT > 1) {
        pair_sync();                               // the partner has consumed the last row-max slots
        xm[part] = l;
        pair_sync();
        l += xm[part ^ 1];
      }
      // ---- normalise, stage this warp's head dims of its 32 rows through (now free) P smem (64 dims per 16 KB block)
      const float inv = 1.f / l;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < OW / 8; ++c) {
        const int dim0 = part * OW + 8 * c;          // first head dim of this 16-byte chunk
        const float* oo = o + 8 * c;
        sts128a(sP_row + (dim0 >> 6) * (128 * 128) + ((((dim0 >> 3) & 7) ^ (row & 7)) << 4), TT::pack2(oo[0] * inv, oo[1] * inv),
                TT::pack2(oo[2] * inv, oo[3] * inv), TT::pack2(oo[4] * inv, oo[5] * inv), TT::pack2(oo[6] * inv, oo[7] * inv));
      }
      __syncwarp();
      pair_sync();                                   // both dim slices of these 32 rows are staged
      const int b = bh / args.heads, hd = bh - b * args.heads;
      const int D = args.heads * HD;
      T* outp = reinterpret_cast<T*>(args.out);
#pragma unroll
      for (int hb = 0; hb < HD / 64; ++hb)
#pragma unroll
        for (int ii = 0; ii < 8 / SPLIT; ++ii) {     // the warps of a row quarter split its 32 rows
          const int i = part * (8 / SPLIT) + ii;
          const int rr = q4 * 32 + i * 4 + (lane >> 3);
          const uint4 val = lds128a(sP_base + hb * (128 * 128) + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
          const int t = q0 + rr;
          if (t < args.ntok)
            *reinterpret_cast<uint4*>(outp + (static_cast<long long>(b) * args.ntok + t) * D + hd * HD + hb * 64 + (lane & 7) * 8) = val;
        }
      __syncwarp();
      pair_sync();                                   // the partner has read my staged chunks: P smem may be rewritten
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int make_map_3d(CUtensorMap* map, const void* base, int dtype, uint64_t d0, uint64_t d1, uint64_t d2,
                       uint64_t stride1_elems, uint64_t stride2_elems, uint32_t b0, uint32_t b1) {
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return encode_tensor_map(map, dtype, 3, base, dims, strides, box, estr);
}

// ---- few-row companion (the cls/storage prefix rows): one CTA per (batch, head), up to 8 query rows -----------------
// phase 1: thread <-> key: each K row is read once (8 x 16 B) and dotted with all query rows (smem broadcast);
// phase 2: exact fp32 softmax per row (block reductions); phase 3: thread <-> (d, key quarter) streams V^T rows
// (keys contiguous) against the probabilities in smem; the 4 key quarters are reduced through smem.
constexpr int AR_MAXROWS = 8;
template <typename T>
__global__ void __launch_bounds__(256) attn_rows_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                        const T* __restrict__ vt, T* __restrict__ out, int heads, int ntok,
                                                        int npad, int row_begin, int nrows, float scale_log2e) {
  extern __shared__ float smf[];
  float* sq = smf;                                   // [8][64]
  float* sp = sq + AR_MAXROWS * 64;                  // [8][npad]
  float* red = sp + AR_MAXROWS * npad;               // [8][8]  per-warp partials, then [8] results in red[0..7]
  float* res = red + AR_MAXROWS * 8;                 // [8] row max, later row sum
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x;
  const T* kb = k + static_cast<size_t>(bh) * ntok * 64;
  const T* vb = vt + static_cast<size_t>(bh) * 64 * npad;
  for (int i = tid; i < nrows * 64; i += 256)
    sq[i] = T16<T>::to_f(q[(static_cast<size_t>(bh) * ntok + row_begin + (i >> 6)) * 64 + (i & 63)]);
  __syncthreads();
  // ---- phase 1: scores.  lane = (key-in-group-of-4, 16-byte chunk): every warp load is 512 contiguous bytes of K;
  // the 8 lanes of a key reduce their partial dot products with shuffles.
  float mx[AR_MAXROWS];
#pragma unroll
  for (int r = 0; r < AR_MAXROWS; ++r) mx[r] = -INFINITY;
  {
    const int kq = lane >> 3, ch = lane & 7;
    float qv[AR_MAXROWS][8];
#pragma unroll
    for (int r = 0; r < AR_MAXROWS; ++r)
#pragma unroll
      for (int e = 0; e < 8; ++e) qv[r][e] = r < nrows ? sq[r * 64 + ch * 8 + e] : 0.f;
    for (int k0 = warp * 4; k0 < ntok; k0 += 32) {
      const int key = k0 + kq;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (key < ntok) u = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<size_t>(key) * 64) + ch);
      const float2 a = T16<T>::unpack2(u.x), b = T16<T>::unpack2(u.y), c = T16<T>::unpack2(u.z), d = T16<T>::unpack2(u.w);
      const float kv[8] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int r = 0; r < AR_MAXROWS; ++r) {
        if (r < nrows) {
          float sc = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) sc = fmaf(qv[r][e], kv[e], sc);
          sc += __shfl_xor_sync(0xffffffffu, sc, 1);
          sc += __shfl_xor_sync(0xffffffffu, sc, 2);
          sc += __shfl_xor_sync(0xffffffffu, sc, 4);
          if (ch == 0 && key < ntok) {
            const float v = sc * scale_log2e;
            sp[r * npad + key] = v;
            mx[r] = fmaxf(mx[r], v);
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < AR_MAXROWS; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
    if (lane == 0) red[r * 8 + warp] = mx[r];
  }
  __syncthreads();
  if (tid < AR_MAXROWS) {
    float m = red[tid * 8];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[tid * 8 + w]);
    res[tid] = m;
  }
  __syncthreads();
  // ---- phase 2: probabilities (16-bit rounded for the PV product, like the MMA path) and fp32 row sums
  float sm[AR_MAXROWS];
#pragma unroll
  for (int r = 0; r < AR_MAXROWS; ++r) sm[r] = 0.f;
  for (int key = tid; key < npad; key += 256) {
#pragma unroll
    for (int r = 0; r < AR_MAXROWS; ++r) {
      if (r < nrows) {
        float e = 0.f;
        if (key < ntok) { e = ex2(sp[r * npad + key] - res[r]); sm[r] += e; }
        sp[r * npad + key] = T16<T>::to_f(T16<T>::from_f(e));
      }
    }
  }
  __syncthreads();   // everyone has read res[] (row max) before it is reused for the sums
#pragma unroll
  for (int r = 0; r < AR_MAXROWS; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sm[r] += __shfl_xor_sync(0xffffffffu, sm[r], o);
    if (lane == 0) red[r * 8 + warp] = sm[r];
  }
  __syncthreads();
  if (tid < AR_MAXROWS) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[tid * 8 + w];
    res[tid] = t;
  }
  __syncthreads();
  // ---- phase 3: O[r][d] = sum_key p[r][key] * V^T[d][key].  One warp per d row at a time, lane = 8-key chunk: every
  // warp load is 512 contiguous bytes of the V^T row; per-row partials are reduced with shuffles.
  const int b = bh / heads, hd = bh - b * heads;
  for (int d = warp; d < 64; d += 8) {
    const T* vr = vb + static_cast<size_t>(d) * npad;
    float acc[AR_MAXROWS];
#pragma unroll
    for (int r = 0; r < AR_MAXROWS; ++r) acc[r] = 0.f;
    for (int key = lane * 8; key < npad; key += 256) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(vr + key));
      const float2 a = T16<T>::unpack2(u.x), bb = T16<T>::unpack2(u.y), c = T16<T>::unpack2(u.z), dd = T16<T>::unpack2(u.w);
      const float vv[8] = {a.x, a.y, bb.x, bb.y, c.x, c.y, dd.x, dd.y};
#pragma unroll
      for (int r = 0; r < AR_MAXROWS; ++r) {
        if (r < nrows) {
          const float4 p0 = *reinterpret_cast<const float4*>(sp + r * npad + key);
          const float4 p1 = *reinterpret_cast<const float4*>(sp + r * npad + key + 4);
          acc[r] = fmaf(p0.x, vv[0], acc[r]); acc[r] = fmaf(p0.y, vv[1], acc[r]); acc[r] = fmaf(p0.z, vv[2], acc[r]);
          acc[r] = fmaf(p0.w, vv[3], acc[r]); acc[r] = fmaf(p1.x, vv[4], acc[r]); acc[r] = fmaf(p1.y, vv[5], acc[r]);
          acc[r] = fmaf(p1.z, vv[6], acc[r]); acc[r] = fmaf(p1.w, vv[7], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < AR_MAXROWS; ++r) {
      if (r < nrows) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        if (lane == 0)
          out[(static_cast<size_t>(b) * ntok + row_begin + r) * (heads * 64) + hd * 64 + d] = T16<T>::from_f(acc[r] / res[r]);
      }
    }
  }
}

extern "C" int b2u_attention_rows(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t heads,
                                  int32_t ntok, int32_t npad, int32_t row_begin, int32_t nrows, float scale, int32_t dtype,
                                  b2u_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!q || !k || !vt || !out) return set_error(-1, "b2u_attention_rows: null pointer");
  if (nrows <= 0) return 0;
  if (npad % 8 || npad < ntok || row_begin < 0 || row_begin + nrows > ntok) return set_error(-1, "b2u_attention_rows: bad range");
  if (nrows > AR_MAXROWS) return set_error(-1, "b2u_attention_rows: at most 8 rows");
  const size_t smem = (static_cast<size_t>(AR_MAXROWS) * (64 + npad + 8 + 1) + 4 * AR_MAXROWS * 64) * sizeof(float);
  if (smem > 200 * 1024) return set_error(-1, "b2u_attention_rows: ntok too large for the few-row kernel");
  static bool configured[2] = {false, false};
  const int di = dtype == B2U_BF16 ? 1 : 0;
  if (!configured[di]) {
    cudaError_t e = dtype == B2U_BF16
                        ? cudaFuncSetAttribute(attn_rows_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                        : cudaFuncSetAttribute(attn_rows_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return set_error(-2, "cudaFuncSetAttribute(attn_rows): %s", cudaGetErrorString(e));
    configured[di] = true;
  }
  const float sl2 = scale * 1.4426950408889634f;
  if (dtype == B2U_BF16)
    attn_rows_kernel<__nv_bfloat16><<<B * heads, 256, smem, stream>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(vt),
        static_cast<__nv_bfloat16*>(out), heads, ntok, npad, row_begin, nrows, sl2);
  else
    attn_rows_kernel<__half><<<B * heads, 256, smem, stream>>>(
        static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(vt),
        static_cast<__half*>(out), heads, ntok, npad, row_begin, nrows, sl2);
  return check_launch("attention_rows");
}

template <typename T, int HD, int SPLIT>
static int launch_attn_tc(const AttnMaps& maps, const AttnArgs& a, cudaStream_t stream) {
  auto kern = attn_tc_kernel<T, HD, SPLIT>;
  using CF = AtCfg<HD, SPLIT>;
  static_assert(CF::kSmem <= 227 * 1024, "attention smem budget");
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::kSmem);
    if (e != cudaSuccess) return set_error(-2, "cudaFuncSetAttribute(attn_tc): %s", cudaGetErrorString(e));
    configured = true;
  }
  const int sms = num_sms();
  const int grid = static_cast<int>(a.items < sms ? a.items : sms);
  kern<<<grid, CF::kThreads, CF::kSmem, stream>>>(maps, a);
  return check_launch("attention_tc");
}

static int attention_tc_impl(const void* q, const void* k, const void* vt, void* out, int B, int heads, int ntok, int npad,
                             int q_begin, int head_dim, float scale, int dtype, cudaStream_t stream) {
  if (!q || !k || !vt || !out) return set_error(-1, "b2u_attention_tc: null pointer");
  if (npad % 8 || npad < ntok) return set_error(-1, "b2u_attention_tc: npad must be a multiple of 8 and >= ntok");
  if (head_dim != 64 && head_dim != 128) return set_error(-1, "b2u_attention_tc: head_dim must be 64 or 128");
  if (q_begin < 0 || q_begin >= ntok) return set_error(-1, "b2u_attention_tc: bad q_begin");
  AttnMaps maps;
  AttnArgs a{};
  a.BH = B * heads;
  a.heads = heads;
  a.ntok = ntok;
  a.q_begin = q_begin;
  a.nchunks = (ntok + 127) / 128;
  const int groups = head_dim == 64 ? 2 : 1;
  const int ntiles = (ntok - q_begin + 127) / 128;
  a.npairs = (ntiles + groups - 1) / groups;
  a.pairs_full = ntiles / groups;
  a.items = static_cast<long long>(a.BH) * a.npairs;
  a.scale_log2e = scale * 1.4426950408889634f;
  a.out = out;
  int rc;
  const uint64_t BH = static_cast<uint64_t>(a.BH), hd = static_cast<uint64_t>(head_dim);
  // option 4 (A/B switch): 0 = third generation (default; 128-key chunks, 1.5-pass softmax), 4 = its single-pass softmax,
  // 5 = fourth generation (64-key chunks, double-buffered S and P: measured slower), 2 / 1 = second generation
  const int gen = get_option(4);
  const uint32_t kc = gen == 5 ? 64u : 128u;
  if (gen == 5) a.nchunks = (ntok + 63) / 64;
  if ((rc = make_map_3d(&maps.q, q, dtype, hd, ntok, BH, hd, static_cast<uint64_t>(ntok) * hd, 64, 128))) return rc;
  if ((rc = make_map_3d(&maps.k, k, dtype, hd, ntok, BH, hd, static_cast<uint64_t>(ntok) * hd, 64, kc))) return rc;
  if ((rc = make_map_3d(&maps.vt, vt, dtype, npad, hd, BH, npad, static_cast<uint64_t>(npad) * hd, 64, static_cast<uint32_t>(head_dim)))) return rc;
  // (option 4 = 1: second generation with one softmax warp per (query tile, TMEM lane quarter))
  if (gen == 5) return attention_tc4_dispatch(maps, a, head_dim, dtype, stream);
  if (gen == 0 || gen == 4) return attention_tc3_dispatch(maps, a, head_dim, dtype, gen == 4, stream);
  if (get_option(4) == 1) {
    if (head_dim == 64)
      return dtype == B2U_BF16 ? launch_attn_tc<__nv_bfloat16, 64, 1>(maps, a, stream) : launch_attn_tc<__half, 64, 1>(maps, a, stream);
    return dtype == B2U_BF16 ? launch_attn_tc<__nv_bfloat16, 128, 1>(maps, a, stream) : launch_attn_tc<__half, 128, 1>(maps, a, stream);
  }
  if (head_dim == 64)
    return dtype == B2U_BF16 ? launch_attn_tc<__nv_bfloat16, 64, 2>(maps, a, stream) : launch_attn_tc<__half, 64, 2>(maps, a, stream);
  return dtype == B2U_BF16 ? launch_attn_tc<__nv_bfloat16, 128, 2>(maps, a, stream) : launch_attn_tc<__half, 128, 2>(maps, a, stream);
}

extern "C" int b2u_attention_tc(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t heads,
                                int32_t ntok, int32_t npad, int32_t q_begin, float scale, int32_t dtype,
                                b2u_stream_t stream_) {
  return attention_tc_impl(q, k, vt, out, B, heads, ntok, npad, q_begin, 64, scale, dtype, static_cast<cudaStream_t>(stream_));
}

extern "C" int b2u_attention_tc_hd(const void* q, const void* k, const void* vt, void* out, int32_t B, int32_t heads,
                                   int32_t ntok, int32_t npad, int32_t head_dim, float scale, int32_t dtype,
                                   b2u_stream_t stream_) {
  return attention_tc_impl(q, k, vt, out, B, heads, ntok, npad, 0, head_dim, scale, dtype, static_cast<cudaStream_t>(stream_));
}

}  // namespace b2u
