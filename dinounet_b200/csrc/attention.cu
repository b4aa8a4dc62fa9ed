// Non-causal softmax attention for the DINOv3 ViT (head_dim 64, ntok = 1029 at 512^2): flash-style online softmax,
// Q tile 64 rows / CTA (4 warps x 16 rows), K/V streamed in 64-key chunks through a cp.async double buffer,
// XOR-swizzled shared memory + ldmatrix, fp32 softmax state, 16-bit P for the second MMA (as SDPA's flash backend).
// v1 uses warp-level mma.sync tensor-core tiles; the tcgen05/TMEM version replaces the two MMAs behind the same ABI.
// Replaces F.scaled_dot_product_attention at dinounet/dinov3/layers/attention.py:116.
#include "common.cuh"
#include "../../include/dinounet_b200.h"
#include "host_util.h"

namespace b2u {

template <typename T> struct MmaOp;
template <> struct MmaOp<__nv_bfloat16> {
  __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};
template <> struct MmaOp<__half> {
  __device__ static __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// 64 rows x 64 elems (128 B per row, 8 x 16 B chunks, chunk ^= row & 7)
template <typename T>
__device__ __forceinline__ void load_tile64(uint32_t smem_base, const T* gbase, int row0, int nrows, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 128;
    const int row = idx >> 3, ch = idx & 7;
    const int grow = row0 + row;
    const bool ok = grow < nrows;
    cp_async16(smem_base + row * 128 + ((ch ^ (row & 7)) << 4), gbase + static_cast<size_t>(ok ? grow : 0) * 64 + ch * 8,
               ok ? 16 : 0);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                       const T* __restrict__ v, T* __restrict__ out, int heads,
                                                       int ntok, float scale_log2e) {
  using TT = T16<T>;
  __shared__ __align__(1024) uint8_t sQ[64 * 128];
  __shared__ __align__(1024) uint8_t sK[2][64 * 128];
  __shared__ __align__(1024) uint8_t sV[2][64 * 128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * 64;
  const T* qb = q + static_cast<size_t>(bh) * ntok * 64;
  const T* kb = k + static_cast<size_t>(bh) * ntok * 64;
  const T* vb = v + static_cast<size_t>(bh) * ntok * 64;
  const uint32_t sQa = smem_u32(sQ);
  const uint32_t sKa[2] = {smem_u32(sK[0]), smem_u32(sK[1])};
  const uint32_t sVa[2] = {smem_u32(sV[0]), smem_u32(sV[1])};

  const int nchunks = (ntok + 63) / 64;
  load_tile64(sQa, qb, q0, ntok, tid);
  load_tile64(sKa[0], kb, 0, ntok, tid);
  load_tile64(sVa[0], vb, 0, ntok, tid);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
  float mrow[2] = {-INFINITY, -INFINITY};
  float lrow[2] = {0.f, 0.f};

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      load_tile64(sKa[buf ^ 1], kb, (c + 1) * 64, ntok, tid);
      load_tile64(sVa[buf ^ 1], vb, (c + 1) * 64, ntok, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (c == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = 2 * ks + (lane >> 4);
        ldsm_x4(sQa + row * 128 + ((ch ^ (row & 7)) << 4), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        const int row = jp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
        const int ch = 2 * ks + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sKa[buf] + row * 128 + ((ch ^ (row & 7)) << 4), b0, b1, b2, b3);
        MmaOp<T>::mma(s[2 * jp], qf[ks], b0, b1);
        MmaOp<T>::mma(s[2 * jp + 1], qf[ks], b2, b3);
      }
    }
    // ---- scale, mask the key tail, online softmax (fp32)
    const int kbase = c * 64;
    const bool tail = kbase + 64 > ntok;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float val = s[j][e] * scale_log2e;
        if (tail) {
          const int key = kbase + j * 8 + (lane & 3) * 2 + (e & 1);
          if (key >= ntok) val = -INFINITY;
        }
        s[j][e] = val;
        mx[e >> 1] = fmaxf(mx[e >> 1], val);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mnew[h] = fmaxf(mrow[h], mx[h]);  // chunk 0 always holds >= 1 valid key, so mnew is finite
      corr[h] = exp2f(mrow[h] - mnew[h]);
      mrow[h] = mnew[h];
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] - mnew[0]);
      const float p1 = exp2f(s[j][1] - mnew[0]);
      const float p2 = exp2f(s[j][2] - mnew[1]);
      const float p3 = exp2f(s[j][3] - mnew[1]);
      // P is rounded to 16 bits for the second MMA; the row sum uses the same rounded values
      const uint32_t lo = TT::pack2(p0, p1), hi = TT::pack2(p2, p3);
      const float2 flo = TT::unpack2(lo), fhi = TT::unpack2(hi);
      rs[0] += flo.x + flo.y;
      rs[1] += fhi.x + fhi.y;
      pf[j >> 1][(j & 1) * 2] = lo;
      pf[j >> 1][(j & 1) * 2 + 1] = hi;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) lrow[h] = lrow[h] * corr[h] + rs[h];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0];
      o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int ch = 2 * jp + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sVa[buf] + row * 128 + ((ch ^ (row & 7)) << 4), b0, b1, b2, b3);
        MmaOp<T>::mma(o[2 * jp], pf[ks], b0, b1);
        MmaOp<T>::mma(o[2 * jp + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }
  // ---- finalize: O /= l, write [B, ntok, heads*64]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 1);
    lrow[h] += __shfl_xor_sync(0xffffffffu, lrow[h], 2);
  }
  const int b = bh / heads, hd = bh - b * heads;
  const int D = heads * 64;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = q0 + warp * 16 + (lane >> 2) + h * 8;
    if (t < ntok) {
      const float inv = 1.f / lrow[h];
      T* dst = out + (static_cast<size_t>(b) * ntok + t) * D + hd * 64 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint32_t*>(dst + j * 8) = TT::pack2(o[j][2 * h] * inv, o[j][2 * h + 1] * inv);
    }
  }
}

extern "C" int b2u_attention(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t heads,
                             int32_t ntok, float scale, int32_t dtype, b2u_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!q || !k || !v || !out) return set_error(-1, "b2u_attention: null pointer");
  if (B * heads > 65535) return set_error(-1, "b2u_attention: B*heads > 65535");
  dim3 grid((ntok + 63) / 64, B * heads);
  const float sl2 = scale * 1.4426950408889634f;
  if (dtype == B2U_BF16)
    attn_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
        static_cast<const __nv_bfloat16*>(v), static_cast<__nv_bfloat16*>(out), heads, ntok, sl2);
  else
    attn_fwd_kernel<__half><<<grid, 128, 0, stream>>>(static_cast<const __half*>(q), static_cast<const __half*>(k),
                                                      static_cast<const __half*>(v), static_cast<__half*>(out), heads,
                                                      ntok, sl2);
  return check_launch("attention");
}

}  // namespace b2u
