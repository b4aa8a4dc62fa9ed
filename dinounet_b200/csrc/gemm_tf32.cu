// tcgen05 kind::tf32 GEMM for the TRAINING path: same parameter block, addressing modes and epilogue as b2u_f32_gemm
// (csrc/fp32_tier.cu) - plain / transposed rows, the 3x3 window of an NHWC image on the A side (conv forward, conv data
// gradient) or on the W side (conv weight gradient, K = output pixels), row remap, 2x2 pixel shuffle, split-K atomics -
// but the products run on the 5th-generation tensor cores: fp32 operands rounded to TF32 (cvt.rna: 10 mantissa bits, the
// mantissa width of the fp16 autocast the reference trains under, nnUNetTrainer.py:899-929, with fp32's exponent range),
// fp32 accumulation in tensor memory.  The fp32 SIMT kernel stays the parity tier (1e-5); this one is the fast tier of
// the train step (the SIMT GEMM was 87 % of it, profiles/r02_train_step_kernels_dinounet_b_b64.md).
//
// Why register-path loaders instead of TMA: five of the eight operand modes are gathers that no tensor map expresses
// without a transposed copy in HBM (A[k][m], W[k][n], flipped 3x3 weights, the per-pixel window with K = pixel index).
// Every mode is therefore loaded with ordinary global loads - 16-byte ones where the mode is K-contiguous; for the
// convolution gradients' row-contiguous operands 4 x 4 blocks (four LDG.128 along the rows, transposed in registers:
// tf32::Roles in gemm_tf32_addr.h) - rounded to TF32 in registers and written to shared memory in exactly the layout a SWIZZLE_128B K-major tensor map would produce
// (row r of a k-block = 128 B = 32 floats; 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) << 4)), so the
// UMMA descriptors are byte-identical to the 16-bit GEMM's (make_desc_k128; +32 B per K = 8 step).
//
// CTA = one 128 x BN output tile of one K slice (grid = tiles (n fastest) x ksplit), 256 threads, 3-stage smem ring:
//   all threads: wait stage free (tcgen05.commit -> mbarrier) -> st.shared the k-block fetched kPF iterations earlier ->
//   fence.proxy.async -> issue the global loads of k-block kb + kPF into the freed register set (in flight across kPF
//   barriers) -> __syncthreads -> one elected thread issues 4 x tcgen05.mma (M128, N = BN, K8) + commit.
// Epilogue: tcgen05.ld, thread = accumulator row.  Two CTAs fit an SM's shared memory (60-96 KB each, BN TMEM columns).
#include <math.h>

#include "common.cuh"
#include "../../include/dinounet_b200.h"
#include "gemm_tf32_addr.h"
#include "host_util.h"

namespace b2u {
namespace {

using namespace tf32;

// A/B knobs (python csrc/build.py with B2U_EXTRA_FLAGS): register prefetch depth in k-blocks, resident CTAs per SM
#ifndef B2U_TF32_PF
#define B2U_TF32_PF 1
#endif
#ifndef B2U_TF32_MINB
#define B2U_TF32_MINB 2
#endif
constexpr int kPF = B2U_TF32_PF;
static_assert(kPF >= 1 && kPF <= kTStages, "a k-block is stored into the ring before its register set is refilled");

template <int BN> struct TfCfg {
  static constexpr int kABytes = kTM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kBarOff = kTStages * kStage;
  static constexpr int kSmem = kBarOff + 64 + 1024;   // + barriers / TMEM slot + alignment slack
};

__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
struct SmemPut {            // tf32::stage's sink: round to TF32, one 16-byte shared-memory store
  uint32_t base;
  __device__ __forceinline__ void operator()(uint32_t off, float4 v) const {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + off), "r"(to_tf32(v.x)), "r"(to_tf32(v.y)), "r"(to_tf32(v.z)),
                 "r"(to_tf32(v.w))
                 : "memory");
  }
};
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ABLK / WBLK: staging shape of the two operands (tf32::pick_modes), compile-time so that each instantiation carries only
// its own loader (a runtime switch cost the row-mode shapes 6-15 %)
template <int BN, bool ABLK, bool WBLK>
__global__ void __launch_bounds__(kTThreads, B2U_TF32_MINB) gemm_tf32_kernel(const b2u_f32_gemm_params p) {
  using C = TfCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  int k_lo, k_hi;
  k_slice(p, static_cast<int>(blockIdx.z), k_lo, k_hi);
  if (k_lo >= k_hi) return;                            // empty K slice (uniform per CTA): nothing to add
  const int nkb = (k_hi - k_lo + kTK - 1) / kTK;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* empty_bar = reinterpret_cast<uint64_t*>(smem + C::kBarOff);   // [kTStages]
  uint64_t* done_bar = empty_bar + kTStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long m0;
  int n0;
  tile_origin(p, BN, static_cast<long long>(blockIdx.x), m0, n0);

  if (tid == 0) {
    for (int s = 0; s < kTStages; ++s) mbar_init(&empty_bar[s], 1);
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const Roles R = make_roles(p, tid, m0, n0, BN, ABLK, WBLK);   // this thread's rows / chunks of the two operand tiles
  constexpr int kWN = BN / 32;
  constexpr uint32_t idesc = make_idesc_f16(2 /* TF32 */, kTM, BN);

  // kPF register sets: k-block kb lives in set kb % kPF from the moment its loads are issued (kPF iterations ahead) until it
  // is stored to the ring.  MEASURED (same box, 15 shapes, profiles/r02_tf32_gemm_micro.md): kPF 1 / 2 / 3 at two CTAs per SM
  // = 32.9 / 37.4 / 42.6 ms in total - deeper prefetch LOSES: the kernel is bound by L2->SM operand bytes (3.7 TB/s at the
  // 768-wide linears = 32 FLOP per operand byte of a 128 x 128 fp32 tile), not by load latency; default 1.
  float4 ra[kPF][4], rw[kPF][4];
#pragma unroll
  for (int u = 0; u < kPF; ++u) {
    const int kh = u < nkb ? k_hi : 0;                 // beyond the slice: fetch_* return zeros without touching memory
    fetch_a<ABLK>(p, R, k_lo + u * kTK, kh, ra[u]);
    fetch_w<WBLK, kWN>(p, R, k_lo + u * kTK, kh, rw[u]);
  }

  for (int kb0 = 0; kb0 < nkb; kb0 += kPF) {
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int kb = kb0 + u;
      if (kb >= nkb) break;                            // uniform over the CTA
      const int s = kb % kTStages;
      if (kb >= kTStages) mbar_wait(&empty_bar[s], static_cast<uint32_t>((kb / kTStages - 1) & 1));   // MMAs of k-block kb - 3 have read the stage
      const uint32_t sbase = smem_base + static_cast<uint32_t>(s) * C::kStage;
      stage<ABLK, 4>(R.a_r, R.a_c0, true, ra[u], SmemPut{sbase});
      stage<WBLK, kWN>(R.w_r, R.w_c0, R.w_n != 0, rw[u], SmemPut{sbase + C::kABytes});
      fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core's async proxy
      if (kb + kPF < nkb) {                            // refill this register set: in flight across kPF barriers / MMA issues
        fetch_a<ABLK>(p, R, k_lo + (kb + kPF) * kTK, k_hi, ra[u]);
        fetch_w<WBLK, kWN>(p, R, k_lo + (kb + kPF) * kTK, k_hi, rw[u]);
      }
      __syncthreads();
      if (warp == 0) {
        if (elect_one()) {
          tc_fence_after();
          const uint64_t da = make_desc_k128(sbase);
          const uint64_t db = make_desc_k128(sbase + C::kABytes);
#pragma unroll
          for (int k = 0; k < kTK / 8; ++k)
            tc_mma_tf32(tmem_base, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(&empty_bar[s]);
          if (kb == nkb - 1) tc_commit(done_bar);      // everything issued so far has completed when this one arrives
        }
        __syncwarp();
      }
    }
  }

  // ---------------------------------------------------------------- epilogue: thread = accumulator row
  mbar_wait(done_bar, 0);
  tc_fence_after();
  {
    const int q = warp & 3, half = warp >> 2;          // TMEM lane quarter (fixed by warp % 4), column half
    constexpr int kCols = BN / 2;
    const EpiRow e = make_epi_row(p, m0 + q * 32 + lane);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(half * kCols);
#pragma unroll 1
    for (int c0 = 0; c0 < kCols; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + static_cast<uint32_t>(c0), v);   // .sync.aligned: every lane, also rows beyond M
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float acc[4] = {__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]), __uint_as_float(v[g * 4 + 3])};
        emit4(p, e, n0 + half * kCols + c0 + g * 4, acc);
      }
      __syncwarp();                                      // reconverge before the next warp-collective tcgen05.ld
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

template <int BN, bool ABLK, bool WBLK>
int launch_tf32(const b2u_f32_gemm_params& p, cudaStream_t stream) {
  auto kern = gemm_tf32_kernel<BN, ABLK, WBLK>;
  static bool configured_dev[64] = {};
  bool& configured = configured_dev[current_device_index()];
  constexpr int kSmem = TfCfg<BN>::kSmem;
  static_assert(kSmem <= 113 * 1024, "two CTAs per SM must fit in shared memory");
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return set_error(-2, "cudaFuncSetAttribute(gemm_tf32): %s", cudaGetErrorString(e));
    configured = true;
  }
  const long long tiles = ((p.M + kTM - 1) / kTM) * ((p.N + BN - 1) / BN);
  if (tiles > 0x7fffffffLL) return set_error(-1, "b2u_tf32_gemm: too many output tiles");
  dim3 grid(static_cast<unsigned>(tiles), 1, p.ksplit > 1 ? p.ksplit : 1);
  kern<<<grid, kTThreads, kSmem, stream>>>(p);
  return check_launch("tf32_gemm");
}

}  // namespace

extern "C" int b2u_tf32_gemm(const b2u_f32_gemm_params* p, b2u_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (const char* why = tf32::validate(p)) return set_error(-1, "b2u_tf32_gemm: %s", why);
  bool a_blk, w_blk;
  tf32::pick_modes(*p, a_blk, w_blk);
  const int mode = a_blk ? 2 : (w_blk ? 1 : 0);        // a_blk implies w_blk
  switch (tf32::pick_bn(p->N) + mode) {
    case 32: return launch_tf32<32, false, false>(*p, stream);
    case 33: return launch_tf32<32, false, true>(*p, stream);
    case 34: return launch_tf32<32, true, true>(*p, stream);
    case 64: return launch_tf32<64, false, false>(*p, stream);
    case 65: return launch_tf32<64, false, true>(*p, stream);
    case 66: return launch_tf32<64, true, true>(*p, stream);
    case 128: return launch_tf32<128, false, false>(*p, stream);
    case 129: return launch_tf32<128, false, true>(*p, stream);
    default: return launch_tf32<128, true, true>(*p, stream);
  }
}

}  // namespace b2u
