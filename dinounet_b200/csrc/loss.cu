// Dice + cross-entropy loss and the online-validation statistics, fused (reference: training/loss/compound_losses.py:31-56,
// dice.py:72-119, robust_ce_loss.py:12-16, nnUNetTrainer.py:363-365 [build], :961-1005 [validation_step]).
//
// The reference runs, on the logits [B,C,H,W] the forward path produces: softmax, a one-hot scatter, three masked
// reductions, cross_entropy, argmax, another scatter and the tp/fp/fn products — about a dozen full passes over
// B*C*H*W.  Here ONE pass reads the logits and labels once and produces every sum (block partials in fp64, fixed
// reduction order -> deterministic), a one-block epilogue forms the loss, and a second pass writes dLoss/dlogits.
// HBM-bound: (4C + label) B/pixel forward, (8C + label) B/pixel backward.
#include "common.cuh"
#include "host_util.h"
#include "gemm_common.h"
#include "../../include/dinounet_b200.h"

#include <algorithm>

namespace b2u {

constexpr int kStat = 6;  // per (sample, class): intersect, sum_pred, sum_gt, tp, fp, fn

template <typename L> __device__ __forceinline__ int load_label(const void* t, long long i) {
  return static_cast<int>(static_cast<const L*>(t)[i]);
}
__device__ __forceinline__ int label_at(const void* t, int kind, long long i) {
  switch (kind) {
    case 0: return load_label<uint8_t>(t, i);
    case 1: return load_label<int32_t>(t, i);
    case 2: return load_label<int64_t>(t, i);
    default: return load_label<float>(t, i);
  }
}

// softmax of one pixel (class stride = plane); returns log-sum-exp pieces
template <int C>
__device__ __forceinline__ void pixel_softmax(const float* __restrict__ z, long long plane, float (&p)[C], float& mx,
                                              float& lse, int& amax) {
  float v[C];
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = z[c * plane];
  mx = v[0];
  amax = 0;
#pragma unroll
  for (int c = 1; c < C; ++c)
    if (v[c] > mx) { mx = v[c]; amax = c; }  // first maximum wins, like torch.argmax
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) { p[c] = expf(v[c] - mx); s += p[c]; }
  const float inv = 1.f / s;
#pragma unroll
  for (int c = 0; c < C; ++c) p[c] *= inv;
  lse = logf(s);
}

template <int C>
__global__ void __launch_bounds__(256) loss_stats_kernel(const float* __restrict__ logits, const void* __restrict__ target,
                                                         int tkind, double* __restrict__ partial, int* __restrict__ bad,
                                                         long long plane) {
  const int b = blockIdx.y;
  const float* zb = logits + static_cast<long long>(b) * C * plane;
  float acc[C * kStat];
  float ce = 0.f;
#pragma unroll
  for (int k = 0; k < C * kStat; ++k) acc[k] = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float p[C], mx, lse;
    int amax;
    pixel_softmax<C>(zb + i, plane, p, mx, lse, amax);
    int t = label_at(target, tkind, static_cast<long long>(b) * plane + i);
    if (t < 0 || t >= C) { *bad = 1; t = 0; }
    ce += lse - (zb[t * plane + i] - mx);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float y = t == c ? 1.f : 0.f, h = amax == c ? 1.f : 0.f;
      acc[c * kStat + 0] += p[c] * y;
      acc[c * kStat + 1] += p[c];
      acc[c * kStat + 2] += y;
      acc[c * kStat + 3] += h * y;
      acc[c * kStat + 4] += h * (1.f - y);
      acc[c * kStat + 5] += (1.f - h) * y;
    }
  }
  // block reduction in fp64, fixed order
  __shared__ double red[8][C * kStat + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k <= C * kStat; ++k) {
    double v = static_cast<double>(k < C * kStat ? acc[k < C * kStat ? k : 0] : ce);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x <= C * kStat) {
    double v = 0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    partial[(static_cast<long long>(b) * gridDim.x + blockIdx.x) * (C * kStat + 1) + threadIdx.x] = v;
  }
}

// stats: double [B][C*6 + 1] (the +1 is the sample's CE sum).  out: float[3] = loss, ce, dice term.
__global__ void loss_finalize_kernel(const double* __restrict__ partial, double* __restrict__ stats,
                                     float* __restrict__ out, int64_t* __restrict__ tpfpfn, int B, int C, int nblk,
                                     long long plane, float w_ce, float w_dice, int batch_dice, int do_bg, float smooth) {
  const int K = C * kStat + 1;
  for (int j = threadIdx.x; j < B * K; j += blockDim.x) {
    const int b = j / K, k = j - b * K;
    double v = 0;
    for (int n = 0; n < nblk; ++n) v += partial[(static_cast<long long>(b) * nblk + n) * K + k];
    stats[j] = v;
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  double ce = 0;
  for (int b = 0; b < B; ++b) ce += stats[b * K + C * kStat];
  ce /= static_cast<double>(B) * static_cast<double>(plane);
  const int c0 = do_bg ? 0 : 1;
  double dc = 0;
  if (batch_dice) {
    for (int c = c0; c < C; ++c) {
      double I = 0, P = 0, G = 0;
      for (int b = 0; b < B; ++b) {
        I += stats[b * K + c * kStat];
        P += stats[b * K + c * kStat + 1];
        G += stats[b * K + c * kStat + 2];
      }
      dc += (2 * I + smooth) / fmax(G + P + smooth, 1e-8);
    }
    dc /= (C - c0);
  } else {
    for (int b = 0; b < B; ++b)
      for (int c = c0; c < C; ++c) {
        const double I = stats[b * K + c * kStat], P = stats[b * K + c * kStat + 1], G = stats[b * K + c * kStat + 2];
        dc += (2 * I + smooth) / fmax(G + P + smooth, 1e-8);
      }
    dc /= static_cast<double>(B) * (C - c0);
  }
  out[1] = static_cast<float>(ce);
  out[2] = static_cast<float>(-dc);
  out[0] = static_cast<float>(w_ce * ce - w_dice * dc);
  if (tpfpfn)
    for (int c = 0; c < C; ++c)
      for (int s = 0; s < 3; ++s) {
        double v = 0;
        for (int b = 0; b < B; ++b) v += stats[b * K + c * kStat + 3 + s];
        tpfpfn[s * C + c] = static_cast<int64_t>(v + 0.5);
      }
}

template <int C>
__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ logits, const void* __restrict__ target,
                                                        int tkind, const double* __restrict__ stats,
                                                        float* __restrict__ grad, int B, long long plane, float w_ce,
                                                        float w_dice, int batch_dice, int do_bg, float smooth,
                                                        float grad_scale) {
  const int b = blockIdx.y;
  constexpr int K = C * kStat + 1;
  // per-class dice coefficients a_c, b_c with d(dice term)/dp_c = a_c * y_c + b_c
  __shared__ float sa[C], sb[C];
  if (threadIdx.x < C) {
    const int c = threadIdx.x, c0 = do_bg ? 0 : 1;
    float a = 0.f, bb = 0.f;
    if (c >= c0) {
      double I = 0, P = 0, G = 0;
      for (int n = batch_dice ? 0 : b; n < (batch_dice ? B : b + 1); ++n) {
        I += stats[n * K + c * kStat];
        P += stats[n * K + c * kStat + 1];
        G += stats[n * K + c * kStat + 2];
      }
      const double num = 2 * I + smooth, raw = G + P + smooth, den = fmax(raw, 1e-8);
      const double cnt = batch_dice ? static_cast<double>(C - c0) : static_cast<double>(B) * (C - c0);
      a = static_cast<float>(-2.0 / (den * cnt));                              // through the intersection
      bb = raw > 1e-8 ? static_cast<float>(num / (den * den * cnt)) : 0.f;      // through sum_pred (clip passes no grad)
    }
    sa[c] = a;
    sb[c] = bb;
  }
  __syncthreads();
  const float* zb = logits + static_cast<long long>(b) * C * plane;
  float* gb = grad + static_cast<long long>(b) * C * plane;
  const float ce_w = w_ce / (static_cast<float>(B) * static_cast<float>(plane));
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float p[C], mx, lse;
    int amax;
    pixel_softmax<C>(zb + i, plane, p, mx, lse, amax);
    int t = label_at(target, tkind, static_cast<long long>(b) * plane + i);
    if (t < 0 || t >= C) t = 0;
    float g[C], dot = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      g[c] = w_dice * (sa[c] * (t == c ? 1.f : 0.f) + sb[c]);
      dot += g[c] * p[c];
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
      gb[c * plane + i] = grad_scale * (ce_w * (p[c] - (t == c ? 1.f : 0.f)) + p[c] * (g[c] - dot));
  }
}

static int stats_blocks(long long plane) {
  return static_cast<int>(std::max<long long>(1, std::min<long long>((plane + 256 * 16 - 1) / (256 * 16), 1024)));
}

extern "C" int64_t b2u_dice_ce_work_doubles(int32_t B, int32_t C, int64_t plane) {
  return static_cast<int64_t>(B) * stats_blocks(plane) * (C * kStat + 1) + static_cast<int64_t>(B) * (C * kStat + 1);
}

// one instantiation per class count: every per-class sum lives in a named register (large counts spill, which is fine
// for a pass that is run once per step)
#define B2U_LOSS_SWITCH(C, BODY)                                                               \
  switch (C) {                                                                                 \
    case 2: { constexpr int CC = 2; BODY } break;                                        \
    case 3: { constexpr int CC = 3; BODY } break;                                        \
    case 4: { constexpr int CC = 4; BODY } break;                                        \
    case 5: { constexpr int CC = 5; BODY } break;                                        \
    case 6: { constexpr int CC = 6; BODY } break;                                        \
    case 7: { constexpr int CC = 7; BODY } break;                                        \
    case 8: { constexpr int CC = 8; BODY } break;                                        \
    case 9: { constexpr int CC = 9; BODY } break;                                        \
    case 10: { constexpr int CC = 10; BODY } break;                                        \
    case 11: { constexpr int CC = 11; BODY } break;                                        \
    case 12: { constexpr int CC = 12; BODY } break;                                        \
    case 13: { constexpr int CC = 13; BODY } break;                                        \
    case 14: { constexpr int CC = 14; BODY } break;                                        \
    case 15: { constexpr int CC = 15; BODY } break;                                        \
    case 16: { constexpr int CC = 16; BODY } break;                                        \
    case 17: { constexpr int CC = 17; BODY } break;                                        \
    case 18: { constexpr int CC = 18; BODY } break;                                        \
    case 19: { constexpr int CC = 19; BODY } break;                                        \
    case 20: { constexpr int CC = 20; BODY } break;                                        \
    case 21: { constexpr int CC = 21; BODY } break;                                        \
    case 22: { constexpr int CC = 22; BODY } break;                                        \
    case 23: { constexpr int CC = 23; BODY } break;                                        \
    case 24: { constexpr int CC = 24; BODY } break;                                        \
    case 25: { constexpr int CC = 25; BODY } break;                                        \
    case 26: { constexpr int CC = 26; BODY } break;                                        \
    case 27: { constexpr int CC = 27; BODY } break;                                        \
    case 28: { constexpr int CC = 28; BODY } break;                                        \
    case 29: { constexpr int CC = 29; BODY } break;                                        \
    case 30: { constexpr int CC = 30; BODY } break;                                        \
    case 31: { constexpr int CC = 31; BODY } break;                                        \
    case 32: { constexpr int CC = 32; BODY } break;                                        \
    default: return set_error(-1, "dice_ce: %d classes not in [2, 32]", C);                    \
  }

extern "C" int b2u_dice_ce_forward(const float* logits, const void* target, int32_t target_kind, double* work,
                                   float* out3, int64_t* tp_fp_fn, int32_t* bad_label, int32_t B, int32_t C,
                                   int64_t plane, float weight_ce, float weight_dice, int32_t batch_dice, int32_t do_bg,
                                   float smooth, b2u_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !work || !out3 || !bad_label) return set_error(-1, "b2u_dice_ce_forward: null pointer");
  if (B <= 0 || plane <= 0 || target_kind < 0 || target_kind > 3) return set_error(-1, "b2u_dice_ce_forward: bad argument");
  if (!do_bg && C < 2) return set_error(-1, "b2u_dice_ce_forward: do_bg=0 needs >= 2 classes");
  const int nblk = stats_blocks(plane);
  double* partial = work;
  double* stats = work + static_cast<int64_t>(B) * nblk * (C * kStat + 1);
  dim3 grid(nblk, B);
  B2U_LOSS_SWITCH(C, (loss_stats_kernel<CC><<<grid, 256, 0, stream>>>(logits, target, target_kind, partial, bad_label,
                                                                      plane));)
  if (int rc = check_launch("dice_ce_stats")) return rc;
  loss_finalize_kernel<<<1, 256, 0, stream>>>(partial, stats, out3, tp_fp_fn, B, C, nblk, plane, weight_ce, weight_dice,
                                              batch_dice, do_bg, smooth);
  return check_launch("dice_ce_finalize");
}

extern "C" int b2u_dice_ce_backward(const float* logits, const void* target, int32_t target_kind, const double* work,
                                    float* grad_logits, int32_t B, int32_t C, int64_t plane, float weight_ce,
                                    float weight_dice, int32_t batch_dice, int32_t do_bg, float smooth,
                                    float grad_scale, b2u_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !work || !grad_logits) return set_error(-1, "b2u_dice_ce_backward: null pointer");
  if (target_kind < 0 || target_kind > 3) return set_error(-1, "b2u_dice_ce_backward: bad target kind");
  const int nblk = stats_blocks(plane);
  const double* stats = work + static_cast<int64_t>(B) * nblk * (C * kStat + 1);
  const int gblk = static_cast<int>(std::max<long long>(1, std::min<long long>((plane + 1023) / 1024, 4096)));
  dim3 grid(gblk, B);
  B2U_LOSS_SWITCH(C, (loss_grad_kernel<CC><<<grid, 256, 0, stream>>>(logits, target, target_kind, stats, grad_logits, B,
                                                                     plane, weight_ce, weight_dice, batch_dice, do_bg,
                                                                     smooth, grad_scale));)
  return check_launch("dice_ce_backward");
}

}  // namespace b2u
