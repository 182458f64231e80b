// swin_mma.cu — window attention on the tensor cores for fp16 activations (the reference's --amp path): the same
// operator as swin.cu's CUDA-core kernels (WindowAttention.forward swin_unetr.py:467-490 inside forward_part1 :554-606,
// with padding / roll / partition / mask / relative-position bias as index arithmetic), but S = QK^T, PV, dP = dO V^T,
// dQ = dS K, dV = P^T dO and dK = dS^T Q run as m16n8k16 tensor-core tiles (fp16 operands, fp32 accumulation).
//
// One CTA per (window, head), 11 warps; a warp owns 16-row tiles of the (padded) n x n score matrix and walks the other
// dimension in blocks of 16: two QK^T MMAs give a 16x16 block of scores in the accumulator layout, bias / mask / softmax
// run on those registers, and the SAME registers re-packed to fp16 are the A operand of the following PV (or dS K)
// MMAs — scores and probabilities never touch shared memory.  With head dimension 16 (every SwinUNETR stage at
// feature_size 48: 48/3 = 96/6 = ... = 16) QK^T is exactly one k-step.  The per-element work (bias lookup, mask,
// exp) is what remains on the CUDA cores: the relative-position index is  L_i - L_j + K0  with one precomputed
// integer per token (the 3-D offset (da, db, dc) is linear in the token coordinates).
// Why mma.sync and not tcgen05 here: the contraction is 16 deep and every score needs per-element work between the
// two GEMMs; a tcgen05 pipeline would move each 128 x 352 score tile TMEM -> registers -> shared memory -> MMA for one
// k-step of work.  The convolutions, where K is in the hundreds, are the tcgen05 kernels (conv_tc.cu, wgrad_tc.cu).
// Numerics: q is scaled then rounded to fp16 like the reference's `q * self.scale`; probabilities are rounded to fp16
// before PV like `attn.to(v.dtype) @ v`; the softmax statistics are fp32.
#include "swin_geom.cuh"
#include <math.h>
#include <stdlib.h>

namespace {

using namespace swin;

constexpr int kWarps = kWinThreads / 32;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float x, float y) {
  __half2 h = __floats2half2_rn(x, y);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t lds32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }

// Compile-time shape of one head: DH real channels, DK = contraction extent of QK^T (zero-padded to a k-step),
// RS = row stride of the row-major operand arrays (conflict-free B-fragment reads), NT = 8-wide output column tiles.
template <int DH> struct Shape {
  static constexpr int DK = DH < 16 ? 16 : DH, RS = DK + 8, KS = DK / 16, NT = DH / 8;
};

struct Smem {
  __half* row[4];      // row-major [NP][RS] operand arrays
  __half* tr[2];       // transposed [DH][VS] operand arrays
  float* table; float* dtable; float* lse; float* delta;
  int* L; int* rid; int* vox;
};

struct Cta {
  int n, NP, VS, head, b, wd, wh, ww, TBL, K0, S1, S2;
  int64_t V;
};

__device__ __forceinline__ Cta cta_setup(const WinGeom& g) {
  Cta c;
  c.n = g.n; c.NP = (g.n + 15) & ~15; c.VS = c.NP + 8;
  c.head = blockIdx.y;
  int wlin = blockIdx.x;
  c.ww = wlin % g.nw[2]; wlin /= g.nw[2];
  c.wh = wlin % g.nw[1]; wlin /= g.nw[1];
  c.wd = wlin % g.nw[0]; c.b = wlin / g.nw[0];
  c.V = (int64_t)g.D * g.H * g.W;
  c.S2 = 2 * g.full[2] - 1; c.S1 = (2 * g.full[1] - 1) * c.S2;
  c.TBL = (2 * g.full[0] - 1) * c.S1;
  c.K0 = (g.full[0] - 1) * c.S1 + (g.full[1] - 1) * c.S2 + (g.full[2] - 1);
  return c;
}

// 8 / 16 / 32 halves of one token's q, k or v (16-byte loads); a padding token's vector is the qkv bias (the reference
// pads after norm1 and before the Linear)
template <int DH>
__device__ __forceinline__ void load_token(const __half* p, const float* bias, bool valid, float mul, float (&v)[DH]) {
  if (valid) {
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      float f[8];
      ld8<__half>(p + c * 8, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[c * 8 + i] = f[i] * mul;
    }
  } else {
#pragma unroll
    for (int d = 0; d < DH; ++d) v[d] = (bias ? __half2float(__float2half_rn(bias[d])) : 0.f) * mul;
  }
}
template <int DH>
__device__ __forceinline__ void store_row(__half* dst, const float (&v)[DH]) {      // dst row of Shape<DH>::DK halves (zero-padded)
#pragma unroll
  for (int d = 0; d < DH; d += 2) *reinterpret_cast<__half2*>(dst + d) = __floats2half2_rn(v[d], v[d + 1]);
#pragma unroll
  for (int d = DH; d < Shape<DH>::DK; d += 2) *reinterpret_cast<__half2*>(dst + d) = __floats2half2_rn(0.f, 0.f);
}
template <int DH>
__device__ __forceinline__ void zero_row(__half* dst) {
#pragma unroll
  for (int d = 0; d < Shape<DH>::DK; d += 2) *reinterpret_cast<__half2*>(dst + d) = __floats2half2_rn(0.f, 0.f);
}

// A fragments (16 rows r0.. x DK) of a row-major array
template <int DH>
__device__ __forceinline__ void load_a(const __half* arr, int r0, int g, int t, uint32_t (&a)[Shape<DH>::KS][4]) {
  constexpr int RS = Shape<DH>::RS;
#pragma unroll
  for (int ks = 0; ks < Shape<DH>::KS; ++ks) {
    a[ks][0] = lds32(arr + (r0 + g) * RS + ks * 16 + 2 * t);
    a[ks][1] = lds32(arr + (r0 + g + 8) * RS + ks * 16 + 2 * t);
    a[ks][2] = lds32(arr + (r0 + g) * RS + ks * 16 + 2 * t + 8);
    a[ks][3] = lds32(arr + (r0 + g + 8) * RS + ks * 16 + 2 * t + 8);
  }
}
// acc(16 x 8) += A(16 x DK) * arr[c0 .. c0+8)^T   (arr row-major [.][RS]: B[k = d][n = row] = arr[c0 + n][d])
template <int DH>
__device__ __forceinline__ void mma_rowmajor(float (&acc)[4], const uint32_t (&a)[Shape<DH>::KS][4], const __half* arr, int c0, int g, int t) {
  constexpr int RS = Shape<DH>::RS;
#pragma unroll
  for (int ks = 0; ks < Shape<DH>::KS; ++ks)
    mma16816(acc, a[ks], lds32(arr + (c0 + g) * RS + ks * 16 + 2 * t), lds32(arr + (c0 + g) * RS + ks * 16 + 2 * t + 8));
}
// acc[nt](16 x 8) += P(16 x 16, accumulator registers re-packed) * X[k0 .. k0+16)(16 x DH)  with X given transposed [DH][VS]
template <int DH>
__device__ __forceinline__ void mma_transposed(float (&acc)[Shape<DH>::NT][4], const uint32_t (&p)[4], const __half* tr, int VS, int k0, int g, int t) {
#pragma unroll
  for (int nt = 0; nt < Shape<DH>::NT; ++nt)
    mma16816(acc[nt], p, lds32(tr + (nt * 8 + g) * VS + k0 + 2 * t), lds32(tr + (nt * 8 + g) * VS + k0 + 2 * t + 8));
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// token bookkeeping written by the thread that owns token t
__device__ __forceinline__ void stage_token_meta(const WinGeom& g, const Cta& c, const Smem& s, int t, TokenInfo& ti) {
  ti.valid = false; ti.vox = 0; ti.rid = 0; ti.rc = 0;
  int L = 0;
  if (t < c.n) {
    ti = token_info(g, c.wd, c.wh, c.ww, t);
    L = (ti.rc >> 16) * c.S1 + ((ti.rc >> 8) & 255) * c.S2 + (ti.rc & 255);
  }
  if (t < c.NP) { s.L[t] = L; s.rid[t] = ti.rid; s.vox[t] = (t < c.n && ti.valid) ? ti.vox : -1; }
}

// ------------------------------------------------------------------------------------------------ forward
template <int DH>
__global__ void __launch_bounds__(kWinThreads)
win_attn_fwd_mma_kernel(const __half* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ table,
                        __half* __restrict__ out, float* __restrict__ lse, WinGeom g, float scale) {
  using S = Shape<DH>;
  extern __shared__ __align__(16) unsigned char smraw[];
  const Cta c = cta_setup(g);
  const int C = g.heads * DH, tid = threadIdx.x;
  Smem s;
  s.row[0] = reinterpret_cast<__half*>(smraw);                     // Q (scaled)
  s.row[1] = s.row[0] + c.NP * S::RS;                              // K
  s.tr[0] = s.row[1] + c.NP * S::RS;                               // V^T
  s.table = reinterpret_cast<float*>(s.tr[0] + DH * c.VS);
  s.L = reinterpret_cast<int*>(s.table + c.TBL); s.rid = s.L + c.NP; s.vox = s.rid + c.NP;
  for (int i = tid; i < c.TBL; i += kWinThreads) s.table[i] = table[(int64_t)i * g.heads + c.head];
  {
    TokenInfo ti;
    stage_token_meta(g, c, s, tid, ti);
    if (tid < c.n) {
      const __half* base = qkv + ((int64_t)c.b * c.V + ti.vox) * (3 * C) + c.head * DH;
      float v[DH];
      load_token<DH>(base, qkv_bias ? qkv_bias + c.head * DH : nullptr, ti.valid, scale, v);
      store_row<DH>(s.row[0] + tid * S::RS, v);
      load_token<DH>(base + C, qkv_bias ? qkv_bias + C + c.head * DH : nullptr, ti.valid, 1.f, v);
      store_row<DH>(s.row[1] + tid * S::RS, v);
      load_token<DH>(base + 2 * C, qkv_bias ? qkv_bias + 2 * C + c.head * DH : nullptr, ti.valid, 1.f, v);
#pragma unroll
      for (int d = 0; d < DH; ++d) s.tr[0][d * c.VS + tid] = __float2half_rn(v[d]);
    } else if (tid < c.NP) {
      zero_row<DH>(s.row[0] + tid * S::RS); zero_row<DH>(s.row[1] + tid * S::RS);
#pragma unroll
      for (int d = 0; d < DH; ++d) s.tr[0][d * c.VS + tid] = __float2half_rn(0.f);
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t4 = lane & 3;
  for (int r0 = warp * 16; r0 < c.NP; r0 += kWarps * 16) {
    uint32_t aq[S::KS][4];
    load_a<DH>(s.row[0], r0, gq, t4, aq);
    const int ra = r0 + gq, rb = ra + 8;
    const int La = s.L[ra] + c.K0, Lb = s.L[rb] + c.K0, rida = s.rid[ra], ridb = s.rid[rb];
    float ma = -INFINITY, mb = -INFINITY, la = 0.f, lb = 0.f;
    float o[S::NT][4];
#pragma unroll
    for (int nt = 0; nt < S::NT; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
    for (int kb = 0; kb < c.NP; kb += 16) {
      float sc[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        sc[h][0] = sc[h][1] = sc[h][2] = sc[h][3] = 0.f;
        mma_rowmajor<DH>(sc[h], aq, s.row[1], kb + h * 8, gq, t4);
      }
      float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = kb + h * 8 + 2 * t4 + e;
          const int Lj = s.L[j], ridj = s.rid[j];
          float va = sc[h][e] + s.table[La - Lj], vb = sc[h][2 + e] + s.table[Lb - Lj];
          if (g.masked) { if (ridj != rida) va -= 100.f; if (ridj != ridb) vb -= 100.f; }
          if (j >= c.n) { va = -INFINITY; vb = -INFINITY; }
          sc[h][e] = va; sc[h][2 + e] = vb;
          mxa = fmaxf(mxa, va); mxb = fmaxf(mxb, vb);
        }
      }
      const float mna = fmaxf(ma, quad_max(mxa)), mnb = fmaxf(mb, quad_max(mxb));
      const float ca = __expf(ma - mna), cb = __expf(mb - mnb);
      ma = mna; mb = mnb;
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[h][e] = __expf(sc[h][e] - mna); sc[h][2 + e] = __expf(sc[h][2 + e] - mnb);
          sa += sc[h][e]; sb += sc[h][2 + e];
        }
      }
      la = la * ca + sa; lb = lb * cb + sb;
#pragma unroll
      for (int nt = 0; nt < S::NT; ++nt) { o[nt][0] *= ca; o[nt][1] *= ca; o[nt][2] *= cb; o[nt][3] *= cb; }
      const uint32_t p[4] = {pack2(sc[0][0], sc[0][1]), pack2(sc[0][2], sc[0][3]), pack2(sc[1][0], sc[1][1]), pack2(sc[1][2], sc[1][3])};
      mma_transposed<DH>(o, p, s.tr[0], c.VS, kb, gq, t4);
    }
    la = quad_sum(la); lb = quad_sum(lb);
    const float ia = 1.f / la, ib = 1.f / lb;
    const int voxa = s.vox[ra], voxb = s.vox[rb];
    if (voxa >= 0) {
      __half* op = out + ((int64_t)c.b * c.V + voxa) * C + c.head * DH + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < S::NT; ++nt) *reinterpret_cast<__half2*>(op + nt * 8) = __floats2half2_rn(o[nt][0] * ia, o[nt][1] * ia);
    }
    if (voxb >= 0) {
      __half* op = out + ((int64_t)c.b * c.V + voxb) * C + c.head * DH + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < S::NT; ++nt) *reinterpret_cast<__half2*>(op + nt * 8) = __floats2half2_rn(o[nt][2] * ib, o[nt][3] * ib);
    }
    if (t4 == 0) {
      float* lp = lse + ((int64_t)blockIdx.x * g.heads + c.head) * c.n;
      if (ra < c.n) lp[ra] = ma + __logf(la);
      if (rb < c.n) lp[rb] = mb + __logf(lb);
    }
  }
}

// staging shared by the two backward passes: Q (scaled), K, V, dO row-major; lse, delta, token bookkeeping
template <int DH>
__device__ __forceinline__ void stage_backward(const WinGeom& g, const Cta& c, const Smem& s, const __half* qkv, const float* qkv_bias,
                                               const __half* out, const __half* dout, const float* lse, float* delta_out,
                                               const float* delta_in, float scale, __half* kt, __half* qt, __half* dot) {
  using S = Shape<DH>;
  const int C = g.heads * DH, tid = threadIdx.x;
  TokenInfo ti;
  stage_token_meta(g, c, s, tid, ti);
  if (tid < c.n) {
    const __half* base = qkv + ((int64_t)c.b * c.V + ti.vox) * (3 * C) + c.head * DH;
    float v[DH];
    load_token<DH>(base, qkv_bias ? qkv_bias + c.head * DH : nullptr, ti.valid, scale, v);
    store_row<DH>(s.row[0] + tid * S::RS, v);
    if (qt) {
#pragma unroll
      for (int d = 0; d < DH; ++d) qt[d * c.VS + tid] = __float2half_rn(v[d]);
    }
    load_token<DH>(base + C, qkv_bias ? qkv_bias + C + c.head * DH : nullptr, ti.valid, 1.f, v);
    store_row<DH>(s.row[1] + tid * S::RS, v);
    if (kt) {
#pragma unroll
      for (int d = 0; d < DH; ++d) kt[d * c.VS + tid] = __float2half_rn(v[d]);
    }
    load_token<DH>(base + 2 * C, qkv_bias ? qkv_bias + 2 * C + c.head * DH : nullptr, ti.valid, 1.f, v);
    store_row<DH>(s.row[2] + tid * S::RS, v);
    // a padding query's output is cropped away by the reference (:600-601): its upstream gradient is zero
    const int64_t oo = ((int64_t)c.b * c.V + ti.vox) * C + c.head * DH;
    float dO[DH];
    load_token<DH>(dout + oo, nullptr, ti.valid, 1.f, dO);
    store_row<DH>(s.row[3] + tid * S::RS, dO);
    if (dot) {
#pragma unroll
      for (int d = 0; d < DH; ++d) dot[d * c.VS + tid] = __float2half_rn(dO[d]);
    }
    const int64_t so = ((int64_t)blockIdx.x * g.heads + c.head) * c.n + tid;
    float dl;
    if (delta_out) {
      float O[DH];
      load_token<DH>(out + oo, nullptr, ti.valid, 1.f, O);
      dl = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) dl += dO[d] * O[d];
      delta_out[so] = dl;
    } else {
      dl = delta_in[so];
    }
    s.delta[tid] = dl;
    s.lse[tid] = lse[so];
  } else if (tid < c.NP) {
#pragma unroll
    for (int a = 0; a < 4; ++a) zero_row<DH>(s.row[a] + tid * S::RS);
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      if (kt) kt[d * c.VS + tid] = __float2half_rn(0.f);
      if (qt) qt[d * c.VS + tid] = __float2half_rn(0.f);
      if (dot) dot[d * c.VS + tid] = __float2half_rn(0.f);
    }
    s.delta[tid] = 0.f; s.lse[tid] = 0.f;
  }
}

template <int DH>
__device__ __forceinline__ void carve_backward(unsigned char* smraw, const Cta& c, Smem& s, bool two_tables) {
  using S = Shape<DH>;
  __half* h = reinterpret_cast<__half*>(smraw);
  for (int a = 0; a < 4; ++a) { s.row[a] = h; h += c.NP * S::RS; }
  s.tr[0] = h; h += DH * c.VS;
  s.tr[1] = h; h += DH * c.VS;
  s.table = reinterpret_cast<float*>(h);
  s.dtable = s.table + c.TBL;
  float* f = s.dtable + (two_tables ? c.TBL : 0);
  s.lse = f; s.delta = f + c.NP;
  s.L = reinterpret_cast<int*>(s.delta + c.NP); s.rid = s.L + c.NP; s.vox = s.rid + c.NP;
}

// ------------------------------------------------------------------------------------------------ backward, pass A
// query-stationary: delta_i = <dO_i, O_i>, dQ_i = scale * sum_j dS_ij K_j, d(bias table)[idx(i,j)] += dS_ij
template <int DH>
__global__ void __launch_bounds__(kWinThreads)
win_attn_bwd_q_mma_kernel(const __half* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ table,
                          const __half* __restrict__ out, const __half* __restrict__ dout, const float* __restrict__ lse,
                          float* __restrict__ delta, __half* __restrict__ dqkv, float* __restrict__ dtable, WinGeom g, float scale) {
  using S = Shape<DH>;
  extern __shared__ __align__(16) unsigned char smraw[];
  const Cta c = cta_setup(g);
  const int C = g.heads * DH, tid = threadIdx.x;
  Smem s;
  carve_backward<DH>(smraw, c, s, true);
  for (int i = tid; i < c.TBL; i += kWinThreads) { s.table[i] = table[(int64_t)i * g.heads + c.head]; s.dtable[i] = 0.f; }
  stage_backward<DH>(g, c, s, qkv, qkv_bias, out, dout, lse, delta, nullptr, scale, s.tr[0], nullptr, nullptr);     // tr[0] = K^T
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t4 = lane & 3;
  for (int r0 = warp * 16; r0 < c.NP; r0 += kWarps * 16) {
    const int ra = r0 + gq, rb = ra + 8;
    const int voxa = s.vox[ra], voxb = s.vox[rb];
    if (__all_sync(0xffffffffu, voxa < 0 && voxb < 0)) continue;          // a tile of padding queries: nothing flows back
    uint32_t aq[S::KS][4], ado[S::KS][4];
    load_a<DH>(s.row[0], r0, gq, t4, aq);
    load_a<DH>(s.row[3], r0, gq, t4, ado);
    const int La = s.L[ra] + c.K0, Lb = s.L[rb] + c.K0, rida = s.rid[ra], ridb = s.rid[rb];
    const float lsa = s.lse[ra], lsb = s.lse[rb], dla = s.delta[ra], dlb = s.delta[rb];
    float dq[S::NT][4];
#pragma unroll
    for (int nt = 0; nt < S::NT; ++nt) { dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f; }
    for (int kb = 0; kb < c.NP; kb += 16) {
      float sc[2][4], dp[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        sc[h][0] = sc[h][1] = sc[h][2] = sc[h][3] = 0.f;
        dp[h][0] = dp[h][1] = dp[h][2] = dp[h][3] = 0.f;
        mma_rowmajor<DH>(sc[h], aq, s.row[1], kb + h * 8, gq, t4);       // S = Q K^T
        mma_rowmajor<DH>(dp[h], ado, s.row[2], kb + h * 8, gq, t4);      // dP = dO V^T
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = kb + h * 8 + 2 * t4 + e;
          const int Lj = s.L[j], ridj = s.rid[j];
          const int ia = La - Lj, ib = Lb - Lj;
          float va = sc[h][e] + s.table[ia], vb = sc[h][2 + e] + s.table[ib];
          if (g.masked) { if (ridj != rida) va -= 100.f; if (ridj != ridb) vb -= 100.f; }
          const bool jin = j < c.n;
          const float dsa = (jin && voxa >= 0) ? __expf(va - lsa) * (dp[h][e] - dla) : 0.f;
          const float dsb = (jin && voxb >= 0) ? __expf(vb - lsb) * (dp[h][2 + e] - dlb) : 0.f;
          if (jin && voxa >= 0) atomicAdd(&s.dtable[ia], dsa);
          if (jin && voxb >= 0) atomicAdd(&s.dtable[ib], dsb);
          sc[h][e] = dsa; sc[h][2 + e] = dsb;
        }
      }
      const uint32_t p[4] = {pack2(sc[0][0], sc[0][1]), pack2(sc[0][2], sc[0][3]), pack2(sc[1][0], sc[1][1]), pack2(sc[1][2], sc[1][3])};
      mma_transposed<DH>(dq, p, s.tr[0], c.VS, kb, gq, t4);              // dQ += dS K
    }
    if (voxa >= 0) {
      __half* op = dqkv + ((int64_t)c.b * c.V + voxa) * (3 * C) + c.head * DH + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < S::NT; ++nt) *reinterpret_cast<__half2*>(op + nt * 8) = __floats2half2_rn(dq[nt][0] * scale, dq[nt][1] * scale);
    }
    if (voxb >= 0) {
      __half* op = dqkv + ((int64_t)c.b * c.V + voxb) * (3 * C) + c.head * DH + 2 * t4;
#pragma unroll
      for (int nt = 0; nt < S::NT; ++nt) *reinterpret_cast<__half2*>(op + nt * 8) = __floats2half2_rn(dq[nt][2] * scale, dq[nt][3] * scale);
    }
  }
  __syncthreads();
  for (int i = tid; i < c.TBL; i += kWinThreads) {
    const float v = s.dtable[i];
    if (v != 0.f) atomicAdd(&dtable[(int64_t)i * g.heads + c.head], v);
  }
}

// ------------------------------------------------------------------------------------------------ backward, pass B
// key-stationary: dV_j = sum_i P_ij dO_i, dK_j = sum_i dS_ij (scale q_i); the padding keys' gradients go to the qkv bias
template <int DH>
__global__ void __launch_bounds__(kWinThreads)
win_attn_bwd_kv_mma_kernel(const __half* __restrict__ qkv, const float* __restrict__ qkv_bias, const float* __restrict__ table,
                           const __half* __restrict__ dout, const float* __restrict__ lse, const float* __restrict__ delta,
                           __half* __restrict__ dqkv, float* __restrict__ dbias_pad, WinGeom g, float scale) {
  using S = Shape<DH>;
  extern __shared__ __align__(16) unsigned char smraw[];
  const Cta c = cta_setup(g);
  const int C = g.heads * DH, tid = threadIdx.x;
  Smem s;
  carve_backward<DH>(smraw, c, s, false);
  for (int i = tid; i < c.TBL; i += kWinThreads) s.table[i] = table[(int64_t)i * g.heads + c.head];
  stage_backward<DH>(g, c, s, qkv, qkv_bias, nullptr, dout, lse, nullptr, delta, scale, nullptr, s.tr[0], s.tr[1]);   // Q^T, dO^T
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, t4 = lane & 3;
  for (int r0 = warp * 16; r0 < c.NP; r0 += kWarps * 16) {                // rows = keys
    uint32_t ak[S::KS][4], av[S::KS][4];
    load_a<DH>(s.row[1], r0, gq, t4, ak);
    load_a<DH>(s.row[2], r0, gq, t4, av);
    const int ja = r0 + gq, jb = ja + 8;
    const int Lja = s.L[ja] - c.K0, Ljb = s.L[jb] - c.K0, rida = s.rid[ja], ridb = s.rid[jb];
    float dk[S::NT][4], dv[S::NT][4];
#pragma unroll
    for (int nt = 0; nt < S::NT; ++nt) { dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f; dv[nt][0] = dv[nt][1] = dv[nt][2] = dv[nt][3] = 0.f; }
    for (int qb = 0; qb < c.NP; qb += 16) {                               // columns = queries
      float sc[2][4], dp[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        sc[h][0] = sc[h][1] = sc[h][2] = sc[h][3] = 0.f;
        dp[h][0] = dp[h][1] = dp[h][2] = dp[h][3] = 0.f;
        mma_rowmajor<DH>(sc[h], ak, s.row[0], qb + h * 8, gq, t4);       // S^T = K Q^T
        mma_rowmajor<DH>(dp[h], av, s.row[3], qb + h * 8, gq, t4);       // dP^T = V dO^T
      }
      float pr[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int i = qb + h * 8 + 2 * t4 + e;                          // query
          const int Li = s.L[i], ridi = s.rid[i];
          const bool live = s.vox[i] >= 0;                                // padding queries carry no gradient
          const float ls = s.lse[i], dl = s.delta[i];
          float va = sc[h][e] + s.table[Li - Lja], vb = sc[h][2 + e] + s.table[Li - Ljb];
          if (g.masked) { if (ridi != rida) va -= 100.f; if (ridi != ridb) vb -= 100.f; }
          const float pa = live ? __expf(va - ls) : 0.f, pb = live ? __expf(vb - ls) : 0.f;
          pr[h][e] = pa; pr[h][2 + e] = pb;
          sc[h][e] = pa * (dp[h][e] - dl); sc[h][2 + e] = pb * (dp[h][2 + e] - dl);
        }
      }
      const uint32_t pp[4] = {pack2(pr[0][0], pr[0][1]), pack2(pr[0][2], pr[0][3]), pack2(pr[1][0], pr[1][1]), pack2(pr[1][2], pr[1][3])};
      const uint32_t ps[4] = {pack2(sc[0][0], sc[0][1]), pack2(sc[0][2], sc[0][3]), pack2(sc[1][0], sc[1][1]), pack2(sc[1][2], sc[1][3])};
      mma_transposed<DH>(dv, pp, s.tr[1], c.VS, qb, gq, t4);             // dV += P^T dO
      mma_transposed<DH>(dk, ps, s.tr[0], c.VS, qb, gq, t4);             // dK += dS^T (scale Q)
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = half ? jb : ja;
      if (j >= c.n) continue;
      const int vox = s.vox[j];
      if (vox >= 0) {
        __half* op = dqkv + ((int64_t)c.b * c.V + vox) * (3 * C) + c.head * DH + 2 * t4;
#pragma unroll
        for (int nt = 0; nt < S::NT; ++nt) {
          *reinterpret_cast<__half2*>(op + C + nt * 8) = __floats2half2_rn(dk[nt][2 * half], dk[nt][2 * half + 1]);
          *reinterpret_cast<__half2*>(op + 2 * C + nt * 8) = __floats2half2_rn(dv[nt][2 * half], dv[nt][2 * half + 1]);
        }
      } else if (dbias_pad) {
#pragma unroll
        for (int nt = 0; nt < S::NT; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            atomicAdd(&dbias_pad[C + c.head * DH + nt * 8 + 2 * t4 + e], dk[nt][2 * half + e]);
            atomicAdd(&dbias_pad[2 * C + c.head * DH + nt * 8 + 2 * t4 + e], dv[nt][2 * half + e]);
          }
        }
      }
    }
  }
}

size_t smem_fwd(const WinGeom& g, int DH) {
  const int NP = (g.n + 15) & ~15, VS = NP + 8, RS = (DH < 16 ? 16 : DH) + 8;
  const int TBL = (2 * g.full[0] - 1) * (2 * g.full[1] - 1) * (2 * g.full[2] - 1);
  return (size_t)2 * (2 * NP * RS + DH * VS) + sizeof(float) * TBL + sizeof(int) * 3 * NP + 16;
}
size_t smem_bwd(const WinGeom& g, int DH, bool two_tables) {
  const int NP = (g.n + 15) & ~15, VS = NP + 8, RS = (DH < 16 ? 16 : DH) + 8;
  const int TBL = (2 * g.full[0] - 1) * (2 * g.full[1] - 1) * (2 * g.full[2] - 1);
  return (size_t)2 * (4 * NP * RS + 2 * DH * VS) + sizeof(float) * ((two_tables ? 2 : 1) * TBL + 2 * NP) + sizeof(int) * 3 * NP + 16;
}

template <int DH>
int launch_fwd(const WinGeom& g, const void* qkv, const float* qkv_bias, const float* table, void* out, float* lse, float scale, cudaStream_t st) {
  const size_t smem = smem_fwd(g, DH);
  B200_CUDA(cudaFuncSetAttribute(win_attn_fwd_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(g.B * g.nw[0] * g.nw[1] * g.nw[2], g.heads);
  win_attn_fwd_mma_kernel<DH><<<grid, kWinThreads, smem, st>>>((const __half*)qkv, qkv_bias, table, (__half*)out, lse, g, scale);
  B200_CHECK_LAUNCH("win_attn_fwd_mma_kernel");
  return B200SEG_OK;
}
template <int DH>
int launch_bwd(const WinGeom& g, const void* qkv, const float* qkv_bias, const float* table, const void* out, const void* dout, const float* lse,
               float* delta, void* dqkv, float* dtable, float* dbias_pad, float scale, cudaStream_t st) {
  const size_t sq = smem_bwd(g, DH, true), skv = smem_bwd(g, DH, false);
  B200_CUDA(cudaFuncSetAttribute(win_attn_bwd_q_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sq));
  B200_CUDA(cudaFuncSetAttribute(win_attn_bwd_kv_mma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)skv));
  dim3 grid(g.B * g.nw[0] * g.nw[1] * g.nw[2], g.heads);
  win_attn_bwd_q_mma_kernel<DH><<<grid, kWinThreads, sq, st>>>((const __half*)qkv, qkv_bias, table, (const __half*)out, (const __half*)dout, lse,
                                                                delta, (__half*)dqkv, dtable, g, scale);
  B200_CHECK_LAUNCH("win_attn_bwd_q_mma_kernel");
  win_attn_bwd_kv_mma_kernel<DH><<<grid, kWinThreads, skv, st>>>((const __half*)qkv, qkv_bias, table, (const __half*)dout, lse, delta,
                                                                  (__half*)dqkv, dbias_pad, g, scale);
  B200_CHECK_LAUNCH("win_attn_bwd_kv_mma_kernel");
  return B200SEG_OK;
}

}  // namespace

// Used by swin.cu's entry points: true when the tensor-core path takes this call (fp16, head dimension 8 / 16 / 32,
// 16-byte aligned rows; B200SEG_WINATTN_MMA=0 forces the CUDA-core kernels for A/B runs and the cross-check test).
bool b200seg_winattn_mma_applies(const void* qkv, int heads, int dh, int dtype) {
  if (dtype != B200SEG_F16 || (dh != 8 && dh != 16 && dh != 32)) return false;
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return false;
  (void)heads;
  const char* e = getenv("B200SEG_WINATTN_MMA");
  return !(e && e[0] == '0');
}

int b200seg_winattn_mma_fwd(const swin::WinGeom& g, const void* qkv, const float* qkv_bias, const float* table, void* out, float* lse,
                            float scale, cudaStream_t st) {
  switch (g.dh) {
    case 8: return launch_fwd<8>(g, qkv, qkv_bias, table, out, lse, scale, st);
    case 16: return launch_fwd<16>(g, qkv, qkv_bias, table, out, lse, scale, st);
    case 32: return launch_fwd<32>(g, qkv, qkv_bias, table, out, lse, scale, st);
  }
  return B200SEG_EUNSUPPORTED;
}

int b200seg_winattn_mma_bwd(const swin::WinGeom& g, const void* qkv, const float* qkv_bias, const float* table, const void* out,
                            const void* dout, const float* lse, float* delta, void* dqkv, float* dtable, float* dbias_pad, float scale,
                            cudaStream_t st) {
  switch (g.dh) {
    case 8: return launch_bwd<8>(g, qkv, qkv_bias, table, out, dout, lse, delta, dqkv, dtable, dbias_pad, scale, st);
    case 16: return launch_bwd<16>(g, qkv, qkv_bias, table, out, dout, lse, delta, dqkv, dtable, dbias_pad, scale, st);
    case 32: return launch_bwd<32>(g, qkv, qkv_bias, table, out, dout, lse, delta, dqkv, dtable, dbias_pad, scale, st);
  }
  return B200SEG_EUNSUPPORTED;
}
