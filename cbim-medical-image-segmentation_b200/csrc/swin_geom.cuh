// swin_geom.cuh — window geometry shared by the window-attention kernels (swin.cu: CUDA-core fp32 / generic path,
// swin_mma.cu: tensor-core fp16 path).  Reference: SwinTransformerBlock.forward_part1 swin_unetr.py:554-606,
// get_window_size :358-381, compute_mask :737-773, relative_position_index :417-459.
#pragma once
#include "common.cuh"

namespace swin {

struct WinGeom {
  int B, D, H, W;            // real token grid
  int ws[3], ss[3];          // effective window / shift (get_window_size, :358-381)
  int P[3], nw[3];           // padded extents, windows per axis
  int heads, dh, n;          // n = ws0*ws1*ws2 tokens per window
  int full[3];               // the module's nominal window (7,7,7): bias-table geometry
  int masked;                // any shift > 0
};

constexpr int kWinThreads = 352;     // >= 343 tokens, 11 warps
constexpr int kMaxDh = 32;

struct TokenInfo { int vox; int rid; int rc; bool valid; };   // vox: linear index into the real grid (b excluded)

__device__ __forceinline__ TokenInfo token_info(const WinGeom& g, int wd, int wh, int ww, int t) {
  TokenInfo ti;
  const int k = t % g.ws[2], j = (t / g.ws[2]) % g.ws[1], i = t / (g.ws[2] * g.ws[1]);
  const int sd = wd * g.ws[0] + i, sh = wh * g.ws[1] + j, sw = ww * g.ws[2] + k;          // shifted (rolled) frame
  int od = sd + g.ss[0]; if (od >= g.P[0]) od -= g.P[0];                                 // torch.roll(x, -shift): x'[s] = x[s+shift]
  int oh = sh + g.ss[1]; if (oh >= g.P[1]) oh -= g.P[1];
  int ow = sw + g.ss[2]; if (ow >= g.P[2]) ow -= g.P[2];
  ti.valid = od < g.D && oh < g.H && ow < g.W;
  ti.vox = (od * g.H + oh) * g.W + ow;
  // region id of compute_mask: along each axis [0, P-ws) -> 0, [P-ws, P-shift) -> 1, [P-shift, P) -> 2
  auto reg = [](int s, int P, int ws, int sft) { return sft == 0 ? 0 : (s < P - ws ? 0 : (s < P - sft ? 1 : 2)); };
  ti.rid = (reg(sd, g.P[0], g.ws[0], g.ss[0]) * 3 + reg(sh, g.P[1], g.ws[1], g.ss[1])) * 3 + reg(sw, g.P[2], g.ws[2], g.ss[2]);
  // relative-position coordinates: the token's LINEAR index decoded in the nominal (7,7,7) window (the [:n,:n] slice)
  const int c = t % g.full[2], b = (t / g.full[2]) % g.full[1], a = t / (g.full[2] * g.full[1]);
  ti.rc = (a << 16) | (b << 8) | c;
  return ti;
}

__device__ __forceinline__ int rel_index(const WinGeom& g, int rci, int rcj) {
  const int da = (rci >> 16) - (rcj >> 16) + g.full[0] - 1;
  const int db = ((rci >> 8) & 255) - ((rcj >> 8) & 255) + g.full[1] - 1;
  const int dc = (rci & 255) - (rcj & 255) + g.full[2] - 1;
  return (da * (2 * g.full[1] - 1) + db) * (2 * g.full[2] - 1) + dc;
}


inline int fill_geom(WinGeom& g, int B, int D, int H, int W, int heads, int dh, const int* window, const int* shift) {
  g.B = B; g.D = D; g.H = H; g.W = W; g.heads = heads; g.dh = dh;
  const int dims[3] = {D, H, W};
  g.masked = 0; g.n = 1;
  for (int i = 0; i < 3; ++i) {
    g.full[i] = window[i];
    g.ws[i] = window[i]; g.ss[i] = shift[i];
    if (dims[i] <= window[i]) { g.ws[i] = dims[i]; g.ss[i] = 0; }          // get_window_size, swin_unetr.py:372-377
    g.P[i] = (dims[i] + g.ws[i] - 1) / g.ws[i] * g.ws[i];
    g.nw[i] = g.P[i] / g.ws[i];
    g.n *= g.ws[i];
    if (g.ss[i] > 0) g.masked = 1;
  }
  if (g.n > kWinThreads || dh > kMaxDh || dh < 1 || window[0] > 127 || window[1] > 127 || window[2] > 127) return B200SEG_EUNSUPPORTED;
  return B200SEG_OK;
}

}  // namespace swin
