// biattn.cu — MedFormer's bidirectional multi-head attention core (B-MHA), fused, forward and backward.
// Reference: BidirectionAttention.forward, model/dim3/medformer_utils.py:63-97 — the part between the q/v
// projections and the output projections:
//     S  = scale * einsum('bhid,bhjd->bhij', feat_q, map_q)            (:77-78)    i in N voxels, j in M map tokens
//     A1 = softmax(S, dim=-1) ; A2 = softmax(S, dim=-2)                (:80,82)
//     feat_out = A1 @ map_v ; map_out = A2^T @ feat_v                  (:84,89)
// The reference materialises S, A1, A2 as [B,h,N,M] fp32 tensors plus six relayout copies; here every voxel is
// visited ONCE per direction: the M<=32 map tokens live in shared memory, the row softmax is thread-local, the
// column softmax (over up to 55k voxels) is an online max/sum with per-block partials merged by a tiny kernel,
// and in the backward the column term  c_j = sum_i A2_ij dA2_ij  collapses to  <dmap_out_j, map_out_j>, so the
// backward is a single pass over N as well.  HBM-bound: fwd reads q_f, v_f and writes out_f (3*B*N*inner*s bytes).
// Channel convention (rearrange1, :43-51): channel c of the `inner` block = d * heads + h.
#include "common.cuh"

namespace {

constexpr int kT = 128;           // voxels per block (one per thread)
constexpr int MAXM_CAP = 64;      // map tokens: 27 (BCV 3x3x3) run the 32-row build, 64 (4x4x4 maps) the 64-row one
constexpr int DH = 32;            // head dimension (all BASELINE MedFormer levels use 32)

struct BiArgs {
  const void* fq; int fq_ld, fq_coff;
  const void* fv; int fv_ld, fv_coff;
  const void* mq; const void* mv; int m_ld;        // [B][M][m_ld], q at +0.., channel = d*heads + h
  int mq_coff, mv_coff;
  void* fo; int fo_ld, fo_coff;
  void* mo; int mo_ld, mo_coff;
  float* colstat;                                  // [B][heads][M][2] = {max, sum} of the column softmax
  float* partial;                                  // fwd: [B][heads][nblk][M][2+DH]; bwd: [B][heads][nblk][M][2*DH]
  // backward only
  const void* dfo; int dfo_ld, dfo_coff;
  const void* dmo; int dmo_ld, dmo_coff;
  void* dfq; int dfq_ld, dfq_coff;
  void* dfv; int dfv_ld, dfv_coff;
  void* dmq; void* dmv; int dm_ld, dmq_coff, dmv_coff;
  int B, M, heads; int64_t N; float scale;
};

constexpr int VFS = DH + 4;        // row stride of the staged value tile: 16-byte aligned rows

__device__ __forceinline__ float dot4(float q0, float q1, float q2, float q3, const float4& w, float acc) {
  return fmaf(q0, w.x, fmaf(q1, w.y, fmaf(q2, w.z, fmaf(q3, w.w, acc))));
}

// The map-side operands live in shared memory zero-padded to MAXM rows (MAXM = 32 or 64: BCV's 27 tokens / the
// 64 of the 4x4x4 maps), so the token loops run unpredicated with one broadcast LDS.128 per four FMAs; padded
// tokens get S = -inf after the contraction.
template <typename T, int MAXM>
__global__ void __launch_bounds__(kT)
biattn_fwd_kernel(BiArgs a) {
  extern __shared__ float sm[];
  float* s_qm = sm;                       // [MAXM][DH] (pre-scaled, rows >= M zero)
  float* s_vm = s_qm + MAXM * DH;         // [MAXM][DH]
  float* s_e = s_vm + MAXM * DH;          // [kT][MAXM+1]
  float* s_vf = s_e + kT * (MAXM + 1);    // [kT][VFS]
  float* s_cmax = s_vf + kT * VFS;        // [4 warps][MAXM] then [MAXM]
  const int h = blockIdx.y, b = blockIdx.z, M = a.M, heads = a.heads;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int o = tid; o < MAXM * DH; o += kT) {
    const int j = o / DH, d = o % DH;
    float q = 0.f, v = 0.f;
    if (j < M) {
      const int64_t off = ((int64_t)b * M + j) * a.m_ld + d * heads + h;
      q = Elem<T>::ld((const T*)a.mq + off + a.mq_coff) * a.scale;
      v = Elem<T>::ld((const T*)a.mv + off + a.mv_coff);
    }
    s_qm[o] = q; s_vm[o] = v;
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kT + tid;
  const bool valid = i < a.N;
  float S[MAXM];
#pragma unroll
  for (int j = 0; j < MAXM; ++j) S[j] = 0.f;
  float* my_v = s_vf + tid * VFS;
  if (valid) {
    const T* qp = (const T*)a.fq + ((int64_t)b * a.N + i) * a.fq_ld + a.fq_coff + h;
    const T* vp = (const T*)a.fv + ((int64_t)b * a.N + i) * a.fv_ld + a.fv_coff + h;
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) {
      const float q0 = Elem<T>::ld(qp + d0 * heads), q1 = Elem<T>::ld(qp + (d0 + 1) * heads);
      const float q2 = Elem<T>::ld(qp + (d0 + 2) * heads), q3 = Elem<T>::ld(qp + (d0 + 3) * heads);
      *reinterpret_cast<float4*>(my_v + d0) = make_float4(Elem<T>::ld(vp + d0 * heads), Elem<T>::ld(vp + (d0 + 1) * heads),
                                                          Elem<T>::ld(vp + (d0 + 2) * heads), Elem<T>::ld(vp + (d0 + 3) * heads));
#pragma unroll
      for (int j = 0; j < MAXM; ++j) S[j] = dot4(q0, q1, q2, q3, *reinterpret_cast<const float4*>(s_qm + j * DH + d0), S[j]);
    }
#pragma unroll
    for (int j = 0; j < MAXM; ++j) if (j >= M) S[j] = -INFINITY;
  } else {
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) *reinterpret_cast<float4*>(my_v + d0) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MAXM; ++j) S[j] = -INFINITY;
  }
  // ---- column direction first (it needs the raw scores): block max per token, exp into shared memory
#pragma unroll
  for (int j = 0; j < MAXM; ++j) {
    float mj = S[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mj = fmaxf(mj, __shfl_xor_sync(0xffffffffu, mj, o));
    if (lane == 0) s_cmax[wid * MAXM + j] = mj;
  }
  __syncthreads();
  if (tid < MAXM) {
    float mj = s_cmax[tid];
    for (int w = 1; w < kT / 32; ++w) mj = fmaxf(mj, s_cmax[w * MAXM + tid]);
    s_cmax[4 * MAXM + tid] = mj;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < MAXM; ++j) {
    const float cm = s_cmax[4 * MAXM + j];
    s_e[tid * (MAXM + 1) + j] = (cm == -INFINITY) ? 0.f : __expf(S[j] - cm);      // padded tokens / empty blocks
  }
  // ---- row softmax over the map tokens (probabilities overwrite the scores) + feat_out
  if (valid) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAXM; ++j) m = fmaxf(m, S[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXM; ++j) { S[j] = __expf(S[j] - m); sum += S[j]; }
    const float inv = 1.f / sum;
    T* op = (T*)a.fo + ((int64_t)b * a.N + i) * a.fo_ld + a.fo_coff + h;
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < MAXM; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(s_vm + j * DH + d0);
        o.x = fmaf(S[j], w.x, o.x); o.y = fmaf(S[j], w.y, o.y); o.z = fmaf(S[j], w.z, o.z); o.w = fmaf(S[j], w.w, o.w);
      }
      Elem<T>::st(op + d0 * heads, o.x * inv); Elem<T>::st(op + (d0 + 1) * heads, o.y * inv);
      Elem<T>::st(op + (d0 + 2) * heads, o.z * inv); Elem<T>::st(op + (d0 + 3) * heads, o.w * inv);
    }
  }
  __syncthreads();
  // 128 threads = 32 tokens x 4 channel octets; each accumulates its 1x8 patch of E^T V over the block's voxels
  const int nblk = gridDim.x;
  float* pb = a.partial + ((((int64_t)b * heads + h) * nblk + blockIdx.x) * M) * (2 + DH);
  const int dg = (tid & 3) * 8;
#pragma unroll
  for (int jh = 0; jh < MAXM; jh += 32) {
    const int j = jh + (tid >> 2);
    float acc[8], esum = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll 4
    for (int r = 0; r < kT; ++r) {
      const float e = s_e[r * (MAXM + 1) + j];
      const float4 v0 = *reinterpret_cast<const float4*>(s_vf + r * VFS + dg), v1 = *reinterpret_cast<const float4*>(s_vf + r * VFS + dg + 4);
      esum += e;
      acc[0] = fmaf(e, v0.x, acc[0]); acc[1] = fmaf(e, v0.y, acc[1]); acc[2] = fmaf(e, v0.z, acc[2]); acc[3] = fmaf(e, v0.w, acc[3]);
      acc[4] = fmaf(e, v1.x, acc[4]); acc[5] = fmaf(e, v1.y, acc[5]); acc[6] = fmaf(e, v1.z, acc[6]); acc[7] = fmaf(e, v1.w, acc[7]);
    }
    if (j < M) {
#pragma unroll
      for (int c = 0; c < 8; ++c) pb[j * (2 + DH) + 2 + dg + c] = acc[c];
      if (dg == 0) { pb[j * (2 + DH)] = s_cmax[4 * MAXM + j]; pb[j * (2 + DH) + 1] = esum; }
    }
  }
}

// merge the per-block column partials: map_out[j][:] and the {max, sum} the backward needs.
// grid (M, heads, B), 128 threads = 4 partitions of the block axis x 32 head channels, online-softmax combine.
template <typename T>
__global__ void biattn_fwd_merge_kernel(BiArgs a, int nblk) {
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z, M = a.M, heads = a.heads;
  const int d = threadIdx.x & 31, part = threadIdx.x >> 5;
  const float* pb = a.partial + ((((int64_t)b * heads + h) * nblk) * M + j) * (2 + DH);
  float m = -INFINITY, sum = 0.f, acc = 0.f;
  for (int k = part; k < nblk; k += 4) {
    const float* q = pb + (int64_t)k * M * (2 + DH);
    const float mk = q[0];
    if (mk > m) { const float sc = __expf(m - mk); sum *= sc; acc *= sc; m = mk; }
    const float e = __expf(mk - m);
    sum = fmaf(q[1], e, sum); acc = fmaf(q[2 + d], e, acc);
  }
  __shared__ float s_m[4], s_s[4], s_a[4][DH];
  if (d == 0) { s_m[part] = m; s_s[part] = sum; }
  s_a[part][d] = acc;
  __syncthreads();
  if (part == 0) {
    const float gm = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    float gs = 0.f, ga = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) { const float e = __expf(s_m[p] - gm); gs = fmaf(s_s[p], e, gs); ga = fmaf(s_a[p][d], e, ga); }
    Elem<T>::st((T*)a.mo + ((int64_t)b * M + j) * a.mo_ld + a.mo_coff + d * heads + h, ga / gs);
    if (d == 0) { float* cs = a.colstat + (((int64_t)b * heads + h) * M + j) * 2; cs[0] = gm; cs[1] = gs; }
  }
}

template <typename T, int MAXM>
__global__ void __launch_bounds__(kT)
biattn_bwd_kernel(BiArgs a) {
  extern __shared__ float sm[];
  float* s_qm = sm;                        // [MAXM][DH]  (unscaled; rows >= M zero)
  float* s_vm = s_qm + MAXM * DH;
  float* s_dmo = s_vm + MAXM * DH;         // [MAXM][DH]
  float* s_col = s_dmo + MAXM * DH;        // [MAXM][4] = {gmax, 1/gsum, c_j, -}
  float* s_p1 = s_col + MAXM * 4;          // [kT][MAXM+1]
  float* s_ds = s_p1 + kT * (MAXM + 1);    // [kT][MAXM+1]
  float* s_do = s_ds + kT * (MAXM + 1);    // [kT][VFS]
  float* s_q = s_do + kT * VFS;            // [kT][VFS]
  const int h = blockIdx.y, b = blockIdx.z, M = a.M, heads = a.heads;
  const int tid = threadIdx.x;
  for (int o = tid; o < MAXM * DH; o += kT) {
    const int j = o / DH, d = o % DH;
    float q = 0.f, v = 0.f, g = 0.f;
    if (j < M) {
      const int64_t off = ((int64_t)b * M + j) * a.m_ld + d * heads + h;
      q = Elem<T>::ld((const T*)a.mq + off + a.mq_coff);
      v = Elem<T>::ld((const T*)a.mv + off + a.mv_coff);
      g = Elem<T>::ld((const T*)a.dmo + ((int64_t)b * M + j) * a.dmo_ld + a.dmo_coff + d * heads + h);
    }
    s_qm[o] = q; s_vm[o] = v; s_dmo[o] = g;
  }
  __syncthreads();
  if (tid < MAXM) {
    float g0 = 0.f, g1 = 0.f, c = 0.f;
    if (tid < M) {
      const float* cs = a.colstat + (((int64_t)b * heads + h) * M + tid) * 2;
      for (int d = 0; d < DH; ++d)
        c = fmaf(s_dmo[tid * DH + d], Elem<T>::ld((const T*)a.mo + ((int64_t)b * M + tid) * a.mo_ld + a.mo_coff + d * heads + h), c);
      g0 = cs[0]; g1 = 1.f / cs[1];
    }
    s_col[tid * 4] = g0; s_col[tid * 4 + 1] = g1; s_col[tid * 4 + 2] = c;
  }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * kT + tid;
  const bool valid = i < a.N;
  float S[MAXM], v[DH];
  float* my_q = s_q + tid * VFS;           // own rows double as register relief; re-read below without a barrier
  float* my_do = s_do + tid * VFS;
#pragma unroll
  for (int j = 0; j < MAXM; ++j) S[j] = 0.f;
  if (valid) {
    const int64_t row = (int64_t)b * a.N + i;
    const T* qp = (const T*)a.fq + row * a.fq_ld + a.fq_coff + h;
    const T* vp = (const T*)a.fv + row * a.fv_ld + a.fv_coff + h;
    const T* gp = (const T*)a.dfo + row * a.dfo_ld + a.dfo_coff + h;
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) {
      const float q0 = Elem<T>::ld(qp + d0 * heads), q1 = Elem<T>::ld(qp + (d0 + 1) * heads);
      const float q2 = Elem<T>::ld(qp + (d0 + 2) * heads), q3 = Elem<T>::ld(qp + (d0 + 3) * heads);
#pragma unroll
      for (int u = 0; u < 4; ++u) v[d0 + u] = Elem<T>::ld(vp + (d0 + u) * heads);
      *reinterpret_cast<float4*>(my_q + d0) = make_float4(q0, q1, q2, q3);
      *reinterpret_cast<float4*>(my_do + d0) = make_float4(Elem<T>::ld(gp + d0 * heads), Elem<T>::ld(gp + (d0 + 1) * heads),
                                                           Elem<T>::ld(gp + (d0 + 2) * heads), Elem<T>::ld(gp + (d0 + 3) * heads));
#pragma unroll
      for (int j = 0; j < MAXM; ++j) S[j] = dot4(q0, q1, q2, q3, *reinterpret_cast<const float4*>(s_qm + j * DH + d0), S[j]);
    }
  } else {
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) {
      *reinterpret_cast<float4*>(my_q + d0) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(my_do + d0) = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) v[d0 + u] = 0.f;
    }
  }
  // row softmax: p1_j = exp(S_j - m) * inv is recomputed where needed (keeps MAXM registers free for 64 tokens)
  float dS[MAXM];
  float m = -INFINITY, inv = 0.f;
  {
#pragma unroll
    for (int j = 0; j < MAXM; ++j) { S[j] = (j < M && valid) ? S[j] * a.scale : -INFINITY; m = fmaxf(m, S[j]); dS[j] = 0.f; }
    if (!valid) m = 0.f;                   // keeps exp(S - m) = exp(-inf) = 0 instead of NaN for padding threads
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAXM; ++j) sum += __expf(S[j] - m);
    inv = valid ? 1.f / sum : 0.f;
    // dA1_j = <dO, Vm_j>
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) {
      const float4 g = *reinterpret_cast<const float4*>(my_do + d0);
#pragma unroll
      for (int j = 0; j < MAXM; ++j) dS[j] = dot4(g.x, g.y, g.z, g.w, *reinterpret_cast<const float4*>(s_vm + j * DH + d0), dS[j]);
    }
    float t1 = 0.f;
#pragma unroll
    for (int j = 0; j < MAXM; ++j) t1 = fmaf(__expf(S[j] - m) * inv, dS[j], t1);
#pragma unroll
    for (int j = 0; j < MAXM; ++j) dS[j] = __expf(S[j] - m) * inv * (dS[j] - t1);
  }
  // column direction: p2_ij = exp(S_ij - gmax_j) / gsum_j ; dS += p2 (dA2 - c_j) ; dVf = sum_j p2 dmo_j
  float dv[DH];
#pragma unroll
  for (int d = 0; d < DH; ++d) dv[d] = 0.f;
#pragma unroll
  for (int j = 0; j < MAXM; ++j) {
    const float p2 = __expf(S[j] - s_col[j * 4]) * s_col[j * 4 + 1];       // 0 for padded tokens / invalid voxels
    float dA2 = 0.f;
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) {
      const float4 w = *reinterpret_cast<const float4*>(s_dmo + j * DH + d0);
      dA2 = dot4(v[d0], v[d0 + 1], v[d0 + 2], v[d0 + 3], w, dA2);
      dv[d0] = fmaf(p2, w.x, dv[d0]); dv[d0 + 1] = fmaf(p2, w.y, dv[d0 + 1]);
      dv[d0 + 2] = fmaf(p2, w.z, dv[d0 + 2]); dv[d0 + 3] = fmaf(p2, w.w, dv[d0 + 3]);
    }
    dS[j] += p2 * (dA2 - s_col[j * 4 + 2]);
  }
  if (valid) {
    const int64_t row = (int64_t)b * a.N + i;
    T* dqp = (T*)a.dfq + row * a.dfq_ld + a.dfq_coff + h;
    T* dvp = (T*)a.dfv + row * a.dfv_ld + a.dfv_coff + h;
#pragma unroll
    for (int d0 = 0; d0 < DH; d0 += 4) {
      float4 dq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < MAXM; ++j) {
        const float4 w = *reinterpret_cast<const float4*>(s_qm + j * DH + d0);
        dq.x = fmaf(dS[j], w.x, dq.x); dq.y = fmaf(dS[j], w.y, dq.y); dq.z = fmaf(dS[j], w.z, dq.z); dq.w = fmaf(dS[j], w.w, dq.w);
      }
      Elem<T>::st(dqp + d0 * heads, dq.x * a.scale); Elem<T>::st(dqp + (d0 + 1) * heads, dq.y * a.scale);
      Elem<T>::st(dqp + (d0 + 2) * heads, dq.z * a.scale); Elem<T>::st(dqp + (d0 + 3) * heads, dq.w * a.scale);
#pragma unroll
      for (int u = 0; u < 4; ++u) Elem<T>::st(dvp + (d0 + u) * heads, dv[d0 + u]);
    }
  }
  // block partials of the map-side gradients: dVm[j][d] = sum_i p1_ij dO_i[d] ; dQm[j][d] = scale sum_i dS_ij q_i[d]
#pragma unroll
  for (int j = 0; j < MAXM; ++j) { s_p1[tid * (MAXM + 1) + j] = __expf(S[j] - m) * inv; s_ds[tid * (MAXM + 1) + j] = dS[j]; }
  __syncthreads();
  const int nblk = gridDim.x;
  float* pb = a.partial + ((((int64_t)b * heads + h) * nblk + blockIdx.x) * M) * (2 * DH);
  const int dg = (tid & 3) * 8;                        // 32 tokens x 4 channel octets per pass, 1x8 register patch each
#pragma unroll
  for (int jh = 0; jh < MAXM; jh += 32) {
    const int j = jh + (tid >> 2);
    float av[8], aq[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { av[c] = 0.f; aq[c] = 0.f; }
#pragma unroll 2
    for (int r = 0; r < kT; ++r) {
      const float pp = s_p1[r * (MAXM + 1) + j], dd = s_ds[r * (MAXM + 1) + j];
      const float4 g0 = *reinterpret_cast<const float4*>(s_do + r * VFS + dg), g1 = *reinterpret_cast<const float4*>(s_do + r * VFS + dg + 4);
      const float4 q0 = *reinterpret_cast<const float4*>(s_q + r * VFS + dg), q1 = *reinterpret_cast<const float4*>(s_q + r * VFS + dg + 4);
      av[0] = fmaf(pp, g0.x, av[0]); av[1] = fmaf(pp, g0.y, av[1]); av[2] = fmaf(pp, g0.z, av[2]); av[3] = fmaf(pp, g0.w, av[3]);
      av[4] = fmaf(pp, g1.x, av[4]); av[5] = fmaf(pp, g1.y, av[5]); av[6] = fmaf(pp, g1.z, av[6]); av[7] = fmaf(pp, g1.w, av[7]);
      aq[0] = fmaf(dd, q0.x, aq[0]); aq[1] = fmaf(dd, q0.y, aq[1]); aq[2] = fmaf(dd, q0.z, aq[2]); aq[3] = fmaf(dd, q0.w, aq[3]);
      aq[4] = fmaf(dd, q1.x, aq[4]); aq[5] = fmaf(dd, q1.y, aq[5]); aq[6] = fmaf(dd, q1.z, aq[6]); aq[7] = fmaf(dd, q1.w, aq[7]);
    }
    if (j < M) {
#pragma unroll
      for (int c = 0; c < 8; ++c) { pb[j * 2 * DH + dg + c] = aq[c] * a.scale; pb[j * 2 * DH + DH + dg + c] = av[c]; }
    }
  }
}

template <typename T>
__global__ void biattn_bwd_merge_kernel(BiArgs a, int nblk) {
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z, M = a.M, heads = a.heads;
  const int c = threadIdx.x & 63, part = threadIdx.x >> 6;            // 64 values (dq | dv) x 4 partitions
  const float* pb = a.partial + ((((int64_t)b * heads + h) * nblk) * M + j) * (2 * DH);
  float acc = 0.f;
  for (int k = part; k < nblk; k += 4) acc += pb[(int64_t)k * M * (2 * DH) + c];
  __shared__ float s_a[4][2 * DH];
  s_a[part][c] = acc;
  __syncthreads();
  if (part == 0) {
    const float v = s_a[0][c] + s_a[1][c] + s_a[2][c] + s_a[3][c];
    const int d = c & 31;
    const int64_t off = ((int64_t)b * M + j) * a.dm_ld + d * heads + h;
    if (c < DH) Elem<T>::st((T*)a.dmq + off + a.dmq_coff, v);
    else Elem<T>::st((T*)a.dmv + off + a.dmv_coff, v);
  }
}

}  // namespace

extern "C" size_t b200seg_biattn_workspace(int B, int64_t N, int M, int heads) {
  const int64_t nblk = (N + kT - 1) / kT;
  return (size_t)B * heads * nblk * M * (2 * DH) * sizeof(float);
}

static int check_args(int B, int64_t N, int M, int heads, int dim_head, int dtype) {
  if (B <= 0 || N <= 0 || M <= 0 || heads <= 0) return B200SEG_EINVAL;
  if (dim_head != DH || M > MAXM_CAP) return B200SEG_EUNSUPPORTED;
  if (dtype != B200SEG_F16 && dtype != B200SEG_F32) return B200SEG_EINVAL;
  if ((N + kT - 1) / kT > 65535 * 1024) return B200SEG_EUNSUPPORTED;
  return B200SEG_OK;
}

extern "C" int b200seg_biattn_fwd(const void* fq, int fq_ld, int fq_coff, const void* fv, int fv_ld, int fv_coff,
                                  const void* mq, int mq_coff, const void* mv, int mv_coff, int m_ld,
                                  void* fo, int fo_ld, int fo_coff, void* mo, int mo_ld, int mo_coff,
                                  float* colstat, float* workspace, int B, int64_t N, int M, int heads, int dim_head,
                                  float scale, int dtype, void* stream) {
  int rc = check_args(B, N, M, heads, dim_head, dtype);
  if (rc) return rc;
  if (!fq || !fv || !mq || !mv || !fo || !mo || !colstat || !workspace) return B200SEG_EINVAL;
  BiArgs a; memset(&a, 0, sizeof(a));
  a.fq = fq; a.fq_ld = fq_ld; a.fq_coff = fq_coff; a.fv = fv; a.fv_ld = fv_ld; a.fv_coff = fv_coff;
  a.mq = mq; a.mv = mv; a.m_ld = m_ld; a.mq_coff = mq_coff; a.mv_coff = mv_coff;
  a.fo = fo; a.fo_ld = fo_ld; a.fo_coff = fo_coff; a.mo = mo; a.mo_ld = mo_ld; a.mo_coff = mo_coff;
  a.colstat = colstat; a.partial = workspace; a.B = B; a.N = N; a.M = M; a.heads = heads; a.scale = scale;
  cudaStream_t st = as_stream(stream);
  const int nblk = (int)((N + kT - 1) / kT);
  dim3 grid(nblk, heads, B);
  const int MM = M <= 32 ? 32 : 64;
  const size_t sm = sizeof(float) * (2 * MM * DH + kT * (MM + 1) + kT * VFS + 5 * MM);
#define B200_BIATTN_FWD(TT, MMM)                                                                              \
  do {                                                                                                        \
    B200_CUDA(cudaFuncSetAttribute(biattn_fwd_kernel<TT, MMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    biattn_fwd_kernel<TT, MMM><<<grid, kT, sm, st>>>(a);                                                      \
    biattn_fwd_merge_kernel<TT><<<dim3(M, heads, B), 128, 0, st>>>(a, nblk);                                  \
  } while (0)
  if (dtype == B200SEG_F16) { if (MM == 32) B200_BIATTN_FWD(__half, 32); else B200_BIATTN_FWD(__half, 64); }
  else { if (MM == 32) B200_BIATTN_FWD(float, 32); else B200_BIATTN_FWD(float, 64); }
#undef B200_BIATTN_FWD
  B200_CHECK_LAUNCH("biattn_fwd");
  return B200SEG_OK;
}

extern "C" int b200seg_biattn_bwd(const void* fq, int fq_ld, int fq_coff, const void* fv, int fv_ld, int fv_coff,
                                  const void* mq, int mq_coff, const void* mv, int mv_coff, int m_ld,
                                  const void* mo, int mo_ld, int mo_coff, const float* colstat,
                                  const void* dfo, int dfo_ld, int dfo_coff, const void* dmo, int dmo_ld, int dmo_coff,
                                  void* dfq, int dfq_ld, int dfq_coff, void* dfv, int dfv_ld, int dfv_coff,
                                  void* dmq, int dmq_coff, void* dmv, int dmv_coff, int dm_ld,
                                  float* workspace, int B, int64_t N, int M, int heads, int dim_head, float scale,
                                  int dtype, void* stream) {
  int rc = check_args(B, N, M, heads, dim_head, dtype);
  if (rc) return rc;
  if (!fq || !fv || !mq || !mv || !mo || !colstat || !dfo || !dmo || !dfq || !dfv || !dmq || !dmv || !workspace) return B200SEG_EINVAL;
  BiArgs a; memset(&a, 0, sizeof(a));
  a.fq = fq; a.fq_ld = fq_ld; a.fq_coff = fq_coff; a.fv = fv; a.fv_ld = fv_ld; a.fv_coff = fv_coff;
  a.mq = mq; a.mv = mv; a.m_ld = m_ld; a.mq_coff = mq_coff; a.mv_coff = mv_coff;
  a.mo = const_cast<void*>(mo); a.mo_ld = mo_ld; a.mo_coff = mo_coff; a.colstat = const_cast<float*>(colstat);
  a.dfo = dfo; a.dfo_ld = dfo_ld; a.dfo_coff = dfo_coff; a.dmo = dmo; a.dmo_ld = dmo_ld; a.dmo_coff = dmo_coff;
  a.dfq = dfq; a.dfq_ld = dfq_ld; a.dfq_coff = dfq_coff; a.dfv = dfv; a.dfv_ld = dfv_ld; a.dfv_coff = dfv_coff;
  a.dmq = dmq; a.dmv = dmv; a.dm_ld = dm_ld; a.dmq_coff = dmq_coff; a.dmv_coff = dmv_coff;
  a.partial = workspace; a.B = B; a.N = N; a.M = M; a.heads = heads; a.scale = scale;
  cudaStream_t st = as_stream(stream);
  const int nblk = (int)((N + kT - 1) / kT);
  dim3 grid(nblk, heads, B);
  const int MM = M <= 32 ? 32 : 64;
  const size_t sm = sizeof(float) * (3 * MM * DH + 4 * MM + 2 * kT * (MM + 1) + 2 * kT * VFS);
#define B200_BIATTN_BWD(TT, MMM)                                                                              \
  do {                                                                                                        \
    B200_CUDA(cudaFuncSetAttribute(biattn_bwd_kernel<TT, MMM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    biattn_bwd_kernel<TT, MMM><<<grid, kT, sm, st>>>(a);                                                      \
    biattn_bwd_merge_kernel<TT><<<dim3(M, heads, B), 256, 0, st>>>(a, nblk);                                  \
  } while (0)
  if (dtype == B200SEG_F16) { if (MM == 32) B200_BIATTN_BWD(__half, 32); else B200_BIATTN_BWD(__half, 64); }
  else { if (MM == 32) B200_BIATTN_BWD(float, 32); else B200_BIATTN_BWD(float, 64); }
#undef B200_BIATTN_BWD
  B200_CHECK_LAUNCH("biattn_bwd");
  return B200SEG_OK;
}
