// tcgen05 tensor-core GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   acc[128 x BN] (fp32, TMEM) = A tile (TMA -> 128B-swizzled smem, K-major) x W tile (TMA, K-major)
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma issuer,
// warps 2..5 = epilogue (each owns one 32-lane TMEM sub-partition: tcgen05.ld -> registers -> fused epilogue -> HBM).
// One output tile per CTA; two CTAs are co-resident per SM (<=97 KB smem, <=256 TMEM columns each) so one CTA's
// epilogue overlaps the other's main loop.  Convolution mode walks the 9 taps with 4-D TMA halo boxes
// (out-of-bounds = zero fill = padding), so no im2col buffer ever exists in HBM.
#include "common.cuh"
#include "../../include/dinounet_b200.h"
#include "host_util.h"
#include "gemm_common.h"

namespace b2u {

template <int BN> struct Cfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 3 : 4);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == B2U_ACT_GELU) return gelu_erf(v);
  if (act == B2U_ACT_RELU) return fmaxf(v, 0.f);
  if (act == B2U_ACT_LRELU) return v > 0.f ? v : 0.01f * v;
  return v;
}

// ------------------------------------------------------------------------------------------------
template <int BN, bool QKV, typename T>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ GemmMaps maps, const GemmArgs args) {
  using C = Cfg<BN>;
  using TT = T16<T>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int nt = tile % args.n_tiles;
  const int mt = tile / args.n_tiles;
  const int n0 = nt * BN;

  // spatial tile decode (conv)
  int img = 0, y0 = 0, x0 = 0;
  if (args.conv) {
    const int per_img = args.tiles_x * args.tiles_y;
    img = mt / per_img;
    const int r = mt % per_img;
    y0 = (r / args.tiles_x) * args.TH;
    x0 = (r % args.tiles_x) * args.TW;
  }

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&maps.a[0]);
    tma_prefetch_desc(&maps.b);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < args.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sA = smem + stage * C::kStageBytes;
        uint8_t* sB = sA + C::kABytes;
        mbar_expect_tx(&full_bar[stage], C::kStageBytes);
        if (args.conv == 0) {
          tma_load_2d(sA, &maps.a[0], &full_bar[stage], kb * BK, mt * BM);
        } else {
          const int tap = kb / args.cb;
          const int c0 = (kb - tap * args.cb) * BK;
          const int dy = tap / 3, dx = tap - dy * 3;
          if (args.conv == 1) {
            tma_load_4d(sA, &maps.a[0], &full_bar[stage], c0, x0 + dx - 1, y0 + dy - 1, img);
          } else {
            // stride 2: input row 2y+dy-1 -> parity map ((dy+1)&1), coordinate y + (dy==0 ? -1 : 0)
            const int py = (dy + 1) & 1, px = (dx + 1) & 1;
            tma_load_4d(sA, &maps.a[py * 2 + px], &full_bar[stage], c0, x0 + (dx == 0 ? -1 : 0),
                        y0 + (dy == 0 ? -1 : 0), img);
          }
        }
        tma_load_2d(sB, &maps.b, &full_bar[stage], kb * BK, n0);
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(TT::kFmt, BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < args.num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sA = smem_u32(smem + stage * C::kStageBytes);
        const uint32_t sB = sA + C::kABytes;
        const uint64_t da = make_desc_k128(sA);
        const uint64_t db = make_desc_k128(sB);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements (32 B) along K inside the 128B swizzle row: +2 in the (addr>>4) field
          tc_mma_f16(tmem_base, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc,
                     (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int q4 = warp & 3;          // TMEM lane quarter this warp may access
    const int r = q4 * 32 + lane;     // row within the tile
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);

    // row addressing
    bool valid;
    long long m;  // logical GEMM row (token / pixel index)
    if (args.conv == 0) {
      m = static_cast<long long>(mt) * BM + r;
      valid = m < args.M;
    } else {
      const int ty = r / args.TW, tx = r - ty * args.TW;
      const int y = y0 + ty, x = x0 + tx;
      valid = (y < args.Ho) && (x < args.Wo);
      m = (static_cast<long long>(img) * args.Ho + y) * args.Wo + x;
    }

    if constexpr (QKV) {
      // ---- masked-bias + round + RoPE + head split (attention.py:30-40,66-92)
      const int b = static_cast<int>(m / args.ntok);
      const int t = static_cast<int>(m - static_cast<long long>(b) * args.ntok);
      const bool rot = t >= args.prefix;
      const float* sinr = args.rope_sin + static_cast<long long>(rot ? t - args.prefix : 0) * 64;
      const float* cosr = args.rope_cos + static_cast<long long>(rot ? t - args.prefix : 0) * 64;
#pragma unroll 1
      for (int g = 0; g < BN / 64; ++g) {
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + g * 64, v0);
        tmem_ld32(taddr + g * 64 + 32, v1);
        tmem_ld_wait();
        const int n = n0 + g * 64;
        if (!valid || n >= args.N) continue;
        const int which = n / args.D;
        const int head = (n - which * args.D) >> 6;
        float x[64];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = __uint_as_float(v0[j]), c = __uint_as_float(v1[j]);
          if (args.epi.bias) { a += __ldg(args.epi.bias + n + j); c += __ldg(args.epi.bias + n + 32 + j); }
          x[j] = TT::to_f(TT::from_f(a));
          x[32 + j] = TT::to_f(TT::from_f(c));
        }
        T* dst = reinterpret_cast<T*>(which == 0 ? args.q : (which == 1 ? args.k : args.v)) +
                 ((static_cast<long long>(b) * args.heads + head) * args.ntok + t) * 64;
        uint32_t packed[32];
        if (which < 2 && rot) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float2 c_lo = *reinterpret_cast<const float2*>(cosr + j);
            const float2 s_lo = *reinterpret_cast<const float2*>(sinr + j);
            const float2 c_hi = *reinterpret_cast<const float2*>(cosr + 32 + j);
            const float2 s_hi = *reinterpret_cast<const float2*>(sinr + 32 + j);
            // out[j] = x[j]*cos[j] - x[j+32]*sin[j] ; out[j+32] = x[j+32]*cos[j+32] + x[j]*sin[j+32]
            packed[j / 2] = TT::pack2(x[j] * c_lo.x - x[j + 32] * s_lo.x, x[j + 1] * c_lo.y - x[j + 33] * s_lo.y);
            packed[16 + j / 2] =
                TT::pack2(x[j + 32] * c_hi.x + x[j] * s_hi.x, x[j + 33] * c_hi.y + x[j + 1] * s_hi.y);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 64; j += 2) packed[j / 2] = TT::pack2(x[j], x[j + 1]);
        }
        uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
        for (int j = 0; j < 8; ++j) d4[j] = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
      }
    } else {
      const b2u_epilogue& e = args.epi;
      long long orow_base = m;
      if (e.rows_in > 0) orow_base = (m / e.rows_in) * e.rows_out + e.row_off + (m % e.rows_in);
      int ps_b = 0, ps_i = 0, ps_j = 0;
      if (e.ps_cout > 0) {
        const long long hw = static_cast<long long>(e.ps_h) * e.ps_w;
        ps_b = static_cast<int>(m / hw);
        const int rem = static_cast<int>(m - ps_b * hw);
        ps_i = rem / e.ps_w;
        ps_j = rem - ps_i * e.ps_w;
      }
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
        const int n = n0 + ch * 32;
        if (!valid || n >= args.N) continue;
        long long orow = orow_base;
        int ocol = n;
        if (e.ps_cout > 0) {
          const int qd = n / e.ps_cout;
          ocol = n - qd * e.ps_cout;
          orow = (static_cast<long long>(ps_b) * (2 * e.ps_h) + 2 * ps_i + (qd >> 1)) * (2 * e.ps_w) + 2 * ps_j + (qd & 1);
        }
        ocol += e.col_off;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float a = __uint_as_float(v[j]);
          if (e.bias) a += __ldg(e.bias + n + j);
          if (e.round16) a = TT::to_f(TT::from_f(a));
          a = apply_act(a, e.act1);
          if (e.scale) a *= __ldg(e.scale + n + j);
          if (e.shift) a += __ldg(e.shift + n + j);
          a = apply_act(a, e.act2);
          f[j] = a;
        }
        const int ncols = (args.N - n) < 32 ? (args.N - n) : 32;  // N tail (multiple of 8 by contract)
        if (e.residual) {
          const float4* rp = reinterpret_cast<const float4*>(e.residual + orow * e.ldres + ocol);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j * 4 < ncols) {
              const float4 rv = rp[j];
              f[4 * j] += rv.x; f[4 * j + 1] += rv.y; f[4 * j + 2] += rv.z; f[4 * j + 3] += rv.w;
            }
          }
        }
        if (e.add16) {
          const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(e.add16) + orow * e.ldadd + ocol);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j * 8 < ncols) {
              const uint4 av = ap[j];
              float2 t0 = TT::unpack2(av.x), t1 = TT::unpack2(av.y), t2 = TT::unpack2(av.z), t3 = TT::unpack2(av.w);
              f[8 * j] += t0.x; f[8 * j + 1] += t0.y; f[8 * j + 2] += t1.x; f[8 * j + 3] += t1.y;
              f[8 * j + 4] += t2.x; f[8 * j + 5] += t2.y; f[8 * j + 6] += t3.x; f[8 * j + 7] += t3.y;
            }
          }
        }
        if (e.out_fp32) {
          float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + orow * e.ldc + ocol);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j * 4 < ncols) op[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<T*>(e.out) + orow * e.ldc + ocol);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j * 8 < ncols)
              op[j] = make_uint4(TT::pack2(f[8 * j], f[8 * j + 1]), TT::pack2(f[8 * j + 2], f[8 * j + 3]),
                                 TT::pack2(f[8 * j + 4], f[8 * j + 5]), TT::pack2(f[8 * j + 6], f[8 * j + 7]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <int BN, bool QKV, typename T>
static int launch_variant(const GemmMaps& maps, const GemmArgs& args, int grid, cudaStream_t stream) {
  auto kern = gemm_tc_kernel<BN, QKV, T>;
  static bool configured = false;  // per template instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::kSmem);
    if (e != cudaSuccess) return set_error(-2, "cudaFuncSetAttribute(gemm_tc): %s", cudaGetErrorString(e));
    configured = true;
  }
  kern<<<grid, 192, Cfg<BN>::kSmem, stream>>>(maps, args);
  return check_launch("gemm_tc");
}

int gemm_v1_dispatch(bool qkv, int bn, int dtype, const GemmMaps& maps, const GemmArgs& args, cudaStream_t stream) {
  const long long grid_ll = static_cast<long long>(args.m_tiles) * args.n_tiles;
  if (grid_ll <= 0 || grid_ll > 0x7FFFFFFFLL) return set_error(-1, "gemm_tc: bad grid");
  const int grid = static_cast<int>(grid_ll);
#define B2U_CASE(Q_, BN_)                                                                          \
  case BN_:                                                                                        \
    return dtype == B2U_BF16 ? launch_variant<BN_, Q_, __nv_bfloat16>(maps, args, grid, stream)   \
                             : launch_variant<BN_, Q_, __half>(maps, args, grid, stream);
  if (qkv) {
    switch (bn) { B2U_CASE(true, 128) default: break; }
  } else {
    switch (bn) { B2U_CASE(false, 32) B2U_CASE(false, 64) B2U_CASE(false, 128) default: break; }
  }
#undef B2U_CASE
  return set_error(-3, "gemm_tc: unsupported BLOCK_N %d", bn);
}

// v2 (persistent, 128x256 tiles) is the default; option 0 (B2U_OPT_GEMM_IMPL) = 1 selects the v1 kernel for A/B runs
static int pick_bn(int N, bool v2) {
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (!v2 || N <= 128) return 128;
  const int waste = (N + 255) / 256 * 256 - N;
  return (waste == 0 || waste * 8 <= N) ? 256 : 128;
}

static int run_gemm(bool qkv, int bn, int dtype, const GemmMaps& maps, const GemmArgs& args, cudaStream_t stream) {
  return get_option(0) == 1 ? gemm_v1_dispatch(qkv, bn, dtype, maps, args, stream)
                            : gemm_v2_dispatch(qkv, bn, dtype, maps, args, stream);
}

// 2-D K-major operand map: dims {K, rows}, box {64, box_rows}, 128B swizzle, OOB zero fill.
static int make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows,
                       int dtype) {
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  return encode_tensor_map(map, dtype, 2, base, dims, strides, box, estr);
}

extern "C" int b2u_gemm(const b2u_gemm_params* p, b2u_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->A || !p->Wp || !p->epi.out) return set_error(-1, "b2u_gemm: null pointer");
  if (p->N % 32 || p->lda % 8 || p->ldw % 8) return set_error(-1, "b2u_gemm: N must be a multiple of 32; lda, ldw multiples of 8");
  const b2u_epilogue& e = p->epi;
  if ((e.ldc % 8) || (e.col_off % 8)) return set_error(-1, "b2u_gemm: ldc/col_off must be multiples of 8");
  if (e.ps_cout > 0 && (e.ps_cout % 32)) return set_error(-1, "b2u_gemm: ps_cout must be a multiple of 32");
  GemmMaps maps;
  GemmArgs a{};
  a.N = p->N;
  a.epi = p->epi;
  a.conv = p->conv;
  const bool v2 = get_option(0) != 1;
  if (e.act1 == B2U_ACT_SWIGLU && (!v2 || p->N % 64 || p->conv != B2U_CONV_NONE))
    return set_error(-1, "b2u_gemm: SwiGLU epilogue needs the v2 kernel, N %% 64 == 0 and a plain GEMM");
  const int bn = (e.act1 == B2U_ACT_SWIGLU && p->N <= 128) ? 128 : pick_bn(p->N, v2);
  a.n_tiles = (p->N + bn - 1) / bn;
  long long m_tiles;
  int rc;
  if (p->conv == B2U_CONV_NONE) {
    a.M = p->M;
    a.num_kb = (p->K + BK - 1) / BK;
    m_tiles = (static_cast<long long>(p->M) + BM - 1) / BM;
    // CTA pairs (cta_group::2) for the 256-wide tiles of plain GEMMs (option 3 = 1 disables, for A/B runs)
    a.pair = (v2 && bn == 256 && m_tiles >= 2 && get_option(3) != 1) ? 1 : 0;
    if ((rc = make_map_2d(&maps.a[0], p->A, p->M, p->K, p->lda, BM, p->dtype))) return rc;
    if ((rc = make_map_2d(&maps.b, p->Wp, p->N, p->K, p->ldw, a.pair ? bn / 2 : bn, p->dtype))) return rc;
  } else {
    const int stride = p->conv == B2U_CONV3X3_S2 ? 2 : 1;
    if (p->C % 8) return set_error(-1, "b2u_gemm(conv): C must be a multiple of 8");
    if (stride == 2 && ((p->Hin | p->Win) & 1)) return set_error(-1, "b2u_gemm(conv s2): odd image size");
    a.Ho = p->Hin / stride;
    a.Wo = p->Win / stride;
    a.cb = (p->C + BK - 1) / BK;
    a.num_kb = 9 * a.cb;
    a.TW = a.Wo >= 128 ? 128 : a.Wo;
    if (a.TW & (a.TW - 1)) return set_error(-1, "b2u_gemm(conv): output width must be a power of two (<128) or >=128");
    if (a.Wo % a.TW) return set_error(-1, "b2u_gemm(conv): output width must be a multiple of the tile width");
    a.TH = BM / a.TW;
    a.tiles_x = a.Wo / a.TW;
    a.tiles_y = (a.Ho + a.TH - 1) / a.TH;
    m_tiles = static_cast<long long>(p->B) * a.tiles_x * a.tiles_y;
    a.M = p->B * a.Ho * a.Wo;
    const int64_t C = p->C;
    // halo-reuse mode (option 2 != 0 disables): stride 1, <= 64 input channels, <= 64 output channels.
    // One [18 rows x 10 px] halo box per 16 x 8 output tile feeds all 9 taps (1.4x instead of 9x L2 -> SM traffic).
    const int halo_opt = get_option(2);
    if (v2 && halo_opt != 1 && stride == 1 && p->C <= 64 && p->N <= 64 && a.Wo % 8 == 0 && a.Ho >= 16) {
      a.conv = 3;
      a.TW = 8; a.TH = 16; a.tiles_x = a.Wo / 8; a.tiles_y = (a.Ho + 15) / 16;
      m_tiles = static_cast<long long>(p->B) * a.tiles_x * a.tiles_y;
      a.halo_stages = bn <= 32 ? 6 : 5;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p->Win, (cuuint64_t)p->Hin, (cuuint64_t)p->B};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * p->Win * 2, (cuuint64_t)C * p->Win * p->Hin * 2};
      cuuint32_t box[4] = {BK, 10, 18, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      if ((rc = encode_tensor_map(&maps.a[0], p->dtype, 4, p->A, dims, strides, box, estr))) return rc;
    } else if (stride == 1) {
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p->Win, (cuuint64_t)p->Hin, (cuuint64_t)p->B};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)C * p->Win * 2, (cuuint64_t)C * p->Win * p->Hin * 2};
      cuuint32_t box[4] = {BK, (cuuint32_t)a.TW, (cuuint32_t)a.TH, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      if ((rc = encode_tensor_map(&maps.a[0], p->dtype, 4, p->A, dims, strides, box, estr))) return rc;
    } else {
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px) {
          const char* base = static_cast<const char*>(p->A) + (static_cast<int64_t>(py) * p->Win + px) * C * 2;
          cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)(p->Win / 2), (cuuint64_t)(p->Hin / 2), (cuuint64_t)p->B};
          cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)C * p->Win * 4, (cuuint64_t)C * p->Win * p->Hin * 2};
          cuuint32_t box[4] = {BK, (cuuint32_t)a.TW, (cuuint32_t)a.TH, 1};
          cuuint32_t estr[4] = {1, 1, 1, 1};
          if ((rc = encode_tensor_map(&maps.a[py * 2 + px], p->dtype, 4, base, dims, strides, box, estr))) return rc;
        }
    }
    if ((rc = make_map_2d(&maps.b, p->Wp, p->N, static_cast<int64_t>(a.num_kb) * BK, p->ldw, bn, p->dtype))) return rc;
  }
  if (m_tiles <= 0 || m_tiles > 0x7FFFFFFFLL) return set_error(-1, "b2u_gemm: bad grid");
  a.m_tiles = static_cast<int>(m_tiles);
  return run_gemm(false, bn, p->dtype, maps, a, stream);
}

extern "C" int b2u_qkv_rope(const b2u_qkv_params* p, b2u_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!p || !p->A || !p->Wp || !p->q || !p->k || !p->v) return set_error(-1, "b2u_qkv_rope: null pointer");
  const int head_dim = p->heads > 0 ? p->D / p->heads : 0;
  if (p->D != p->heads * head_dim || (head_dim != 64 && head_dim != 128))
    return set_error(-1, "b2u_qkv_rope: head_dim must be 64 or 128");
  GemmMaps maps;
  GemmArgs a{};
  a.M = p->B * p->ntok;
  a.N = 3 * p->D;
  a.num_kb = (p->D + BK - 1) / BK;
  const bool v2 = get_option(0) != 1;
  const int bn = v2 ? pick_bn(a.N, true) : 128;
  a.n_tiles = (a.N + bn - 1) / bn;
  a.conv = 0;
  a.epi.bias = p->bias;
  a.ntok = p->ntok; a.D = p->D; a.heads = p->heads; a.prefix = p->prefix; a.head_dim = head_dim;
  if (head_dim == 128 && !v2) return set_error(-1, "b2u_qkv_rope: head_dim 128 needs the v2 GEMM kernel");
  a.rope_sin = p->rope_sin; a.rope_cos = p->rope_cos;
  a.q = p->q; a.k = p->k; a.v = p->v;
  a.npad = p->v_transposed ? p->npad : 0;
  if (p->rope_w > 0 && v2) {
    a.rope_w = p->rope_w;
    a.rope_h = (p->ntok - p->prefix) / p->rope_w;
    if (a.rope_h * a.rope_w != p->ntok - p->prefix || a.rope_h + a.rope_w > 128)
      return set_error(-1, "b2u_qkv_rope: rope grid %d x %d does not match ntok/prefix or is too large", a.rope_h, a.rope_w);
  }
  if (a.npad && (a.npad % 8 || a.npad < p->ntok)) return set_error(-1, "b2u_qkv_rope: bad npad");
  int rc;
  a.m_tiles = static_cast<int>((static_cast<long long>(a.M) + BM - 1) / BM);
  a.pair = (v2 && bn == 256 && a.m_tiles >= 2 && get_option(3) != 1) ? 1 : 0;
  if ((rc = make_map_2d(&maps.a[0], p->A, a.M, p->D, p->lda, BM, p->dtype))) return rc;
  if ((rc = make_map_2d(&maps.b, p->Wp, a.N, p->D, p->ldw, a.pair ? bn / 2 : bn, p->dtype))) return rc;
  if (a.npad && !v2) return set_error(-1, "b2u_qkv_rope: V^T output needs the v2 GEMM kernel");
  return run_gemm(true, bn, p->dtype, maps, a, stream);
}

}  // namespace b2u
