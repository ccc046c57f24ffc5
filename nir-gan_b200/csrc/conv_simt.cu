// CUDA-core implicit-GEMM convolution (fp32 accumulate) over the tap-table geometry.
// This is the fp32 "verification mode" of the hot path (north-star: <=1e-4 vs the reference) and the
// independent on-device checker for the tcgen05 kernel; it accepts f32 / f16 / bf16 operands.
#include "common.cuh"

namespace ng {

constexpr int ST_PX = 64;   // virtual pixels per tile (8 x 8 patch)
constexpr int ST_CO = 64;   // output channels per tile
constexpr int ST_K = 16;    // channels per smem chunk

template <typename T>
__global__ void __launch_bounds__(256)
conv_simt_kernel(const __grid_constant__ ConvGeom g, const T* __restrict__ x, const T* __restrict__ w,
                 const float* __restrict__ bias, void* __restrict__ yv, int epilogue, int act, float slope,
                 int crop, int patches_x, int patches_y, int co_tiles) {
  __shared__ float As[ST_K][ST_PX + 4];
  __shared__ float Bs[ST_K][ST_CO + 4];

  int tile = blockIdx.x;
  const int cot = tile % co_tiles; tile /= co_tiles;
  const int px_ = tile % patches_x; tile /= patches_x;
  const int py_ = tile % patches_y; tile /= patches_y;
  const int n = tile % g.B;
  const int phase = tile / g.B;

  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  // loader mapping: 4 threads per row, 4 consecutive channels each
  const int lrow = t >> 2, lk = (t & 3) * 4;
  const int li = py_ * 8 + (lrow >> 3), lj = px_ * 8 + (lrow & 7);   // virtual pixel of the A row
  const int lco = cot * ST_CO + lrow;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tp = g.phase_tap0[phase]; tp < g.phase_tap0[phase + 1]; ++tp) {
    const int by = g.S * li + g.taps[tp].dy, bx = g.S * lj + g.taps[tp].dx;
    const bool a_ok = by >= 0 && by < g.Hb && bx >= 0 && bx < g.Wb;
    const T* ap = x + (((size_t)n * g.Hb + (a_ok ? by : 0)) * g.Wb + (a_ok ? bx : 0)) * g.Cin;
    const bool b_ok = lco < g.Cout;
    const T* bp = w + ((size_t)g.taps[tp].wrow + (b_ok ? lco : 0)) * g.Cin;
    for (int c0 = 0; c0 < g.Cin; c0 += ST_K) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        As[lk + e][lrow] = a_ok ? to_f32<T>(ap[c0 + lk + e]) : 0.f;
        Bs[lk + e][lrow] = b_ok ? to_f32<T>(bp[c0 + lk + e]) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < ST_K; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = ty * 4 + i;
    const int vi = py_ * 8 + (row >> 3), vj = px_ * 8 + (row & 7);
    const int oy = g.OS * vi + g.phase_oy[phase], ox = g.OS * vj + g.phase_ox[phase];
    if (oy >= g.Hout || ox >= g.Wout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = cot * ST_CO + tx * 4 + j;
      if (co >= g.Cout) continue;
      float v = acc[i][j];
      if (epilogue == NG_EPI_RAW) {
        reinterpret_cast<T*>(yv)[(((size_t)n * g.Hout + oy) * g.Wout + ox) * g.Cout + co] = from_f32<T>(v);
      } else if (epilogue == NG_EPI_BIAS_ACT) {
        v = apply_act(v + (bias ? bias[co] : 0.f), act, slope);
        reinterpret_cast<T*>(yv)[(((size_t)n * g.Hout + oy) * g.Wout + ox) * g.Cout + co] = from_f32<T>(v);
      } else {  // NG_EPI_HEAD
        if (co != 0) continue;
        const int hy = oy - crop, hx = ox - crop, HH = g.Hout - 2 * crop, WW = g.Wout - 2 * crop;
        if (hy < 0 || hx < 0 || hy >= HH || hx >= WW) continue;
        v = apply_act(v + (bias ? bias[0] : 0.f), act, slope);
        reinterpret_cast<float*>(yv)[((size_t)n * HH + hy) * WW + hx] = v;
      }
    }
  }
}

template <typename T>
static int launch_simt(const ng_conv_args& a, const ConvGeom& g, cudaStream_t st) {
  const int patches_y = (g.VH + 7) / 8, patches_x = (g.VW + 7) / 8, co_tiles = (g.Cout + ST_CO - 1) / ST_CO;
  const long long tiles = (long long)g.nphase * g.B * patches_y * patches_x * co_tiles;
  NG_REQUIRE(tiles < (1ll << 31), NG_E_SHAPE, "conv_simt: too many tiles");
  conv_simt_kernel<T><<<(unsigned)tiles, 256, 0, st>>>(g, (const T*)a.x, (const T*)a.w, a.bias, a.y, a.epilogue,
                                                      a.act, a.slope, a.crop, patches_x, patches_y, co_tiles);
  NG_LAUNCH_CHECK("conv_simt_kernel");
  return NG_OK;
}

int conv_simt(const ng_conv_args& a, const ConvGeom& g, cudaStream_t st) {
  NG_REQUIRE(a.Cin % ST_K == 0, NG_E_SHAPE, "conv_simt: Cin %d must be a multiple of %d", a.Cin, ST_K);
  switch (a.dtype) {
    case NG_F32:  return launch_simt<float>(a, g, st);
    case NG_F16:  return launch_simt<__half>(a, g, st);
    case NG_BF16: return launch_simt<__nv_bfloat16>(a, g, st);
  }
  set_error("conv_simt: bad dtype %d", a.dtype);
  return NG_E_ARG;
}

// ---- weight gradient ------------------------------------------------------------------------------
// dw[tap][n][k] += sum over virtual pixels of dy[.., n] * x[.. shifted .., k].  Tile: 64 n x 64 k,
// pixels split across blocks, fp32 atomics into a zero-initialised dw.
constexpr int WG_PIX = 16;

template <typename T>
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const __grid_constant__ ConvGeom g, const T* __restrict__ x, const T* __restrict__ dy,
                  float* __restrict__ dw, int n_tiles, int k_tiles, int splits, int pix_per_split) {
  __shared__ float Ys[WG_PIX][64 + 4];
  __shared__ float Xs[WG_PIX][64 + 4];
  int id = blockIdx.x;
  const int sp = id % splits; id /= splits;
  const int kt = id % k_tiles; id /= k_tiles;
  const int nt = id % n_tiles; id /= n_tiles;
  const int tp = id;
  int phase = 0;
  while (tp >= g.phase_tap0[phase + 1]) ++phase;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lp = t >> 4, lc = (t & 15) * 4;   // loader: 16 pixels x 16 threads x 4 channels
  const long long npix = (long long)g.B * g.VH * g.VW;
  const long long p0 = (long long)sp * pix_per_split, p1 = min(npix, p0 + pix_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long pb = p0; pb < p1; pb += WG_PIX) {
    const long long p = pb + lp;
    bool ok = p < p1;
    int n = 0, vi = 0, vj = 0;
    if (ok) { vj = (int)(p % g.VW); long long r = p / g.VW; vi = (int)(r % g.VH); n = (int)(r / g.VH); }
    const int oy = g.OS * vi + g.phase_oy[phase], ox = g.OS * vj + g.phase_ox[phase];
    const bool y_ok = ok && oy < g.Hout && ox < g.Wout;
    const int by = g.S * vi + g.taps[tp].dy, bx = g.S * vj + g.taps[tp].dx;
    const bool x_ok = y_ok && by >= 0 && by < g.Hb && bx >= 0 && bx < g.Wb;
    const T* yp = dy + (((size_t)n * g.Hout + (y_ok ? oy : 0)) * g.Wout + (y_ok ? ox : 0)) * g.Cout;
    const T* xp = x + (((size_t)n * g.Hb + (x_ok ? by : 0)) * g.Wb + (x_ok ? bx : 0)) * g.Cin;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int cn = nt * 64 + lc + e, ck = kt * 64 + lc + e;
      Ys[lp][lc + e] = (x_ok && cn < g.Cout) ? to_f32<T>(yp[cn]) : 0.f;
      Xs[lp][lc + e] = (x_ok && ck < g.Cin) ? to_f32<T>(xp[ck]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < WG_PIX; ++q) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = Ys[q][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Xs[q][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int wbase = g.taps[tp].wrow;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cn = nt * 64 + ty * 4 + i;
    if (cn >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ck = kt * 64 + tx * 4 + j;
      if (ck >= g.Cin) continue;
      atomicAdd(&dw[((size_t)wbase + cn) * g.Cin + ck], acc[i][j]);
    }
  }
}

// Weight gradient of a single-real-output-channel convolution (PatchGAN last layer, stored Cout = 16 with channel 0
// real): dw[tap][0][k] = sum_px dy[px][0] * x[px + tap][k].  grid = (taps, pixel splits).  Thread t owns the 8-channel
// group cg = t mod Cin/8 (one 16-byte load per pixel) and the pixel lane pl = t / (Cin/8); lanes walk the split's
// pixels with stride `lanes`, four loads in flight (32-bit index arithmetic), and are summed through shared memory so
// the block issues one atomic per output.  Latency-bound: many short splits.
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_cout1_kernel(const __grid_constant__ ConvGeom g, const T* __restrict__ x, const T* __restrict__ dy,
                   float* __restrict__ dw, int pix_per_split) {
  __shared__ float red[256 * 8];
  const int tp = blockIdx.x, sp = blockIdx.y;
  int phase = 0;
  while (tp >= g.phase_tap0[phase + 1]) ++phase;
  const int C8 = g.Cin >> 3;
  const int lanes = 256 / C8;                       // C8 divides 256 (checked by the launcher)
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8;
  const unsigned npix = (unsigned)g.B * g.VH * g.VW;      // < 2^31 (checked by the launcher)
  const unsigned p0 = (unsigned)sp * pix_per_split, p1 = min(npix, p0 + pix_per_split);
  const int tdy = g.taps[tp].dy, tdx = g.taps[tp].dx, poy = g.phase_oy[phase], pox = g.phase_ox[phase];
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  constexpr int U = 4;
  for (unsigned pb = p0 + pl; pb < p1; pb += lanes * U) {
    float d[U];
    uint4 raw[U];
    const T* xp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned p = pb + u * lanes;
      d[u] = 0.f;
      xp[u] = nullptr;
      if (p < p1) {
        const unsigned r = p / (unsigned)g.VW;
        const int vj = (int)(p - r * g.VW);
        const int n = (int)(r / (unsigned)g.VH);
        const int vi = (int)(r - (unsigned)n * g.VH);
        const int oy = g.OS * vi + poy, ox = g.OS * vj + pox;
        const int by = g.S * vi + tdy, bx = g.S * vj + tdx;
        if (oy < g.Hout && ox < g.Wout && by >= 0 && by < g.Hb && bx >= 0 && bx < g.Wb) {
          d[u] = to_f32<T>(dy[(((size_t)n * g.Hout + oy) * g.Wout + ox) * g.Cout]);
          xp[u] = x + (((size_t)n * g.Hb + by) * g.Wb + bx) * g.Cin + cg * 8;
          if constexpr (sizeof(T) == 2) raw[u] = *reinterpret_cast<const uint4*>(xp[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (xp[u] == nullptr || d[u] == 0.f) continue;
      if constexpr (sizeof(T) == 2) {
        const float2 a0 = unpack2<T>(raw[u].x), a1 = unpack2<T>(raw[u].y), a2 = unpack2<T>(raw[u].z), a3 = unpack2<T>(raw[u].w);
        acc[0] = fmaf(d[u], a0.x, acc[0]); acc[1] = fmaf(d[u], a0.y, acc[1]);
        acc[2] = fmaf(d[u], a1.x, acc[2]); acc[3] = fmaf(d[u], a1.y, acc[3]);
        acc[4] = fmaf(d[u], a2.x, acc[4]); acc[5] = fmaf(d[u], a2.y, acc[5]);
        acc[6] = fmaf(d[u], a3.x, acc[6]); acc[7] = fmaf(d[u], a3.y, acc[7]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(d[u], to_f32<T>(xp[u][k]), acc[k]);
      }
    }
  }
  // sum the pixel lanes of each channel group: red[k][thread], then thread (cg, 0) adds up its `lanes` partners
#pragma unroll
  for (int k = 0; k < 8; ++k) red[k * 256 + threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = threadIdx.x; o < C8 * 8; o += 256) {
    const int c = o >> 3, k = o & 7;
    float t = 0.f;
    for (int L = 0; L < lanes; ++L) t += red[k * 256 + L * C8 + c];
    atomicAdd(&dw[(size_t)g.taps[tp].wrow * g.Cin + c * 8 + k], t);
  }
}

template <typename T>
__global__ void colsum_kernel(const T* __restrict__ dy, long long rows, int C, float* __restrict__ out) {
  // grid.x blocks stride over rows; thread handles channel threadIdx.x (C <= 1024 handled by loop)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) s += to_f32<T>(dy[r * C + c]);
    atomicAdd(&out[c], s);
  }
}

// Column sums of a [rows][C] 16-bit matrix, C a multiple of 8 with C/8 dividing 256 (bias gradients of the wide layers:
// up to 2.5 M rows).  Thread t owns the 8-channel group t mod C/8 (one 16-byte load per row) and the row lane t / (C/8);
// four rows in flight per thread, lanes summed through shared memory, one atomic per channel and block.
template <typename T>
__global__ void __launch_bounds__(256)
colsum_vec_kernel(const T* __restrict__ dy, long long rows, int C, float* __restrict__ out) {
  __shared__ float red[8 * 256];
  const int C8 = C >> 3, lanes = 256 / C8;
  const int cg = threadIdx.x % C8, pl = threadIdx.x / C8;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const long long stride = (long long)gridDim.x * lanes;
  constexpr int U = 4;
  for (long long r0 = (long long)blockIdx.x * lanes + pl; r0 < rows; r0 += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long r = r0 + u * stride;
      v[u] = r < rows ? __ldg(reinterpret_cast<const uint4*>(dy + r * C + cg * 8)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float2 a0 = unpack2<T>(v[u].x), a1 = unpack2<T>(v[u].y), a2 = unpack2<T>(v[u].z), a3 = unpack2<T>(v[u].w);
      acc[0] += a0.x; acc[1] += a0.y; acc[2] += a1.x; acc[3] += a1.y;
      acc[4] += a2.x; acc[5] += a2.y; acc[6] += a3.x; acc[7] += a3.y;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[k * 256 + threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = threadIdx.x; o < C; o += 256) {
    const int c = o >> 3, k = o & 7;
    float t = 0.f;
    for (int L = 0; L < lanes; ++L) t += red[k * 256 + L * C8 + c];
    atomicAdd(&out[o], t);
  }
}

// Same gradient with every input pixel read ONCE (the kernel above walks the 71 MB activation once per tap: 16 times
// for the 4x4 PatchGAN layer, 0.34 ms).  Stride-1 geometries: the block walks a contiguous range of positions of the
// haloed input buffer; thread (cg, pl) keeps all ntaps x 8 accumulators of its channel group in registers and, per input
// position, multiplies its 8 channels by the <= 16 output gradients that reach it (dy[n][by - dy_t][bx - dx_t], an L1 hit
// shared by the 64 channel groups).  Per-block partials go to the workspace and a second kernel adds them in block order:
// deterministic, no atomics.
constexpr int WG1_TAPS = 16, WG1_TPT = 8;            // taps supported / taps per thread (two thread halves share a position)
template <typename T>
__global__ void __launch_bounds__(512, 1)
wgrad_cout1_once_kernel(const __grid_constant__ ConvGeom g, const T* __restrict__ x, const T* __restrict__ dy,
                        float* __restrict__ partial, int pos_per_block) {
  static_assert(sizeof(T) == 2, "16-bit storage only");
  __shared__ float red[512 * 8];
  const int C8 = g.Cin >> 3, lanes = 256 / C8;
  const int half = threadIdx.x >> 8, tl = threadIdx.x & 255;         // taps [8 * half, 8 * half + 8)
  const int cg = tl % C8, pl = tl / C8;
  const unsigned npos = (unsigned)g.B * g.Hb * g.Wb;
  const unsigned q0 = blockIdx.x * (unsigned)pos_per_block, q1 = min(npos, q0 + (unsigned)pos_per_block);
  float acc[WG1_TPT][8];
  int tdy[WG1_TPT], tdx[WG1_TPT];
#pragma unroll
  for (int t = 0; t < WG1_TPT; ++t) {
    const int tp = half * WG1_TPT + t;
    // a tap beyond ntaps gets an offset no position can satisfy
    tdy[t] = tp < g.ntaps ? g.taps[tp].dy : (1 << 20);
    tdx[t] = tp < g.ntaps ? g.taps[tp].dx : (1 << 20);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[t][k] = 0.f;
  }
  const unsigned per_img = (unsigned)g.Hb * g.Wb;
  unsigned q = q0 + pl;
  uint4 raw = make_uint4(0u, 0u, 0u, 0u);
  if (q < q1) raw = *reinterpret_cast<const uint4*>(x + (size_t)q * g.Cin + cg * 8);
  while (q < q1) {
    const unsigned qn = q + lanes;
    uint4 nxt = make_uint4(0u, 0u, 0u, 0u);
    if (qn < q1) nxt = *reinterpret_cast<const uint4*>(x + (size_t)qn * g.Cin + cg * 8);     // next position's load in flight
    const int n = (int)(q / per_img);
    const unsigned r = q - (unsigned)n * per_img;
    const int by = (int)(r / (unsigned)g.Wb), bx = (int)(r - (unsigned)by * g.Wb);
    const T* dyn = dy + (size_t)n * g.Hout * g.Wout * g.Cout;
    float d[WG1_TPT];
#pragma unroll
    for (int t = 0; t < WG1_TPT; ++t) {                    // the gradients that reach this position: loads first
      const int vi = by - tdy[t], vj = bx - tdx[t];
      d[t] = 0.f;
      if ((unsigned)vi < (unsigned)g.Hout && (unsigned)vj < (unsigned)g.Wout)
        d[t] = to_f32<T>(dyn[((size_t)vi * g.Wout + vj) * g.Cout]);
    }
    const float2 a0 = unpack2<T>(raw.x), a1 = unpack2<T>(raw.y), a2 = unpack2<T>(raw.z), a3 = unpack2<T>(raw.w);
#pragma unroll
    for (int t = 0; t < WG1_TPT; ++t) {
      acc[t][0] = fmaf(d[t], a0.x, acc[t][0]); acc[t][1] = fmaf(d[t], a0.y, acc[t][1]);
      acc[t][2] = fmaf(d[t], a1.x, acc[t][2]); acc[t][3] = fmaf(d[t], a1.y, acc[t][3]);
      acc[t][4] = fmaf(d[t], a2.x, acc[t][4]); acc[t][5] = fmaf(d[t], a2.y, acc[t][5]);
      acc[t][6] = fmaf(d[t], a3.x, acc[t][6]); acc[t][7] = fmaf(d[t], a3.y, acc[t][7]);
    }
    raw = nxt;
    q = qn;
  }
  // per tap: sum the pixel lanes of each channel group in lane order, one partial row [tap][Cin] per block
  float* dst = partial + (size_t)blockIdx.x * g.ntaps * g.Cin;
#pragma unroll
  for (int t = 0; t < WG1_TPT; ++t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k * 512 + threadIdx.x] = acc[t][k];
    __syncthreads();
    const int tp = half * WG1_TPT + t;
    if (tp < g.ntaps) {
      for (int o = tl; o < C8 * 8; o += 256) {
        const int c = o >> 3, k = o & 7;
        float v = 0.f;
        for (int L = 0; L < lanes; ++L) v += red[k * 512 + half * 256 + L * C8 + c];
        dst[(size_t)tp * g.Cin + o] = v;
      }
    }
    __syncthreads();
  }
}

// dw[taps[t].wrow][c] = sum over blocks (in block order) of partial[b][t][c]
__global__ void __launch_bounds__(256)
wgrad_cout1_reduce_kernel(const __grid_constant__ ConvGeom g, const float* __restrict__ partial, int blocks,
                          float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.ntaps * g.Cin) return;
  const int t = i / g.Cin, c = i - t * g.Cin;
  const size_t stride = (size_t)g.ntaps * g.Cin;
  float v = 0.f;
  int b = 0;
  for (; b + 4 <= blocks; b += 4) {
    const float p0 = partial[(size_t)b * stride + i], p1 = partial[(size_t)(b + 1) * stride + i],
                p2 = partial[(size_t)(b + 2) * stride + i], p3 = partial[(size_t)(b + 3) * stride + i];
    v += p0; v += p1; v += p2; v += p3;
  }
  for (; b < blocks; ++b) v += partial[(size_t)b * stride + i];
  dw[(size_t)g.taps[t].wrow * g.Cin + c] = v;
}

static bool wgrad_cout1_once_ok(const ng_conv_args& a, const ConvGeom& g, int* blocks, int* pos_per_block) {
  static const bool on = [] { const char* e = getenv("NIRGAN_B200_WGRAD_COUT1_ONCE"); return !(e && e[0] == '0'); }();
  const int c8 = g.Cin / 8;
  if (!on || (a.dtype != NG_F16 && a.dtype != NG_BF16) || g.Cout > 16 || g.nphase != 1 || g.S != 1 || g.OS != 1 ||
      g.ntaps > WG1_TAPS || g.Cin % 8 != 0 || c8 > 256 || 256 % c8 != 0)
    return false;
  const long long npos = (long long)g.B * g.Hb * g.Wb;
  if (npos >= (1ll << 31)) return false;
  const int lanes = 256 / c8;
  long long nb = num_sms();
  if (nb * lanes * 4 > npos) nb = (npos + lanes * 4 - 1) / (lanes * 4);     // at least four positions per lane
  if (nb < 1) nb = 1;
  const long long ppb = ((npos + nb - 1) / nb + lanes - 1) / lanes * lanes;
  *pos_per_block = (int)ppb;
  *blocks = (int)((npos + ppb - 1) / ppb);
  return true;
}

long long wgrad_simt_workspace_bytes(const ng_conv_args& a, const ConvGeom& g) {
  int blocks = 0, ppb = 0;
  if (!wgrad_cout1_once_ok(a, g, &blocks, &ppb)) return 0;
  return (long long)blocks * g.ntaps * g.Cin * (long long)sizeof(float);
}

template <typename T>
static int launch_wgrad(const ng_conv_args& a, const ConvGeom& g, float* dw, void* workspace, long long workspace_bytes,
                        cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    int blocks = 0, ppb = 0;
    if (a.epilogue == NG_EPI_HEAD && wgrad_cout1_once_ok(a, g, &blocks, &ppb) && workspace != nullptr &&
        workspace_bytes >= (long long)blocks * g.ntaps * g.Cin * (long long)sizeof(float)) {
      // the packed gradient has stored Cout (16) rows per tap; only row 0 is real: clear, then write row 0 of every tap
      int e = check_cuda(cudaMemsetAsync(dw, 0, (size_t)a.KH * a.KW * g.Cout * g.Cin * sizeof(float), st), "wgrad memset");
      if (e) return e;
      wgrad_cout1_once_kernel<T><<<blocks, 512, 0, st>>>(g, (const T*)a.x, (const T*)a.y, (float*)workspace, ppb);
      NG_LAUNCH_CHECK("wgrad_cout1_once_kernel");
      wgrad_cout1_reduce_kernel<<<(g.ntaps * g.Cin + 255) / 256, 256, 0, st>>>(g, (const float*)workspace, blocks, dw);
      NG_LAUNCH_CHECK("wgrad_cout1_reduce_kernel");
      return NG_OK;
    }
  }
  const int n_tiles = (g.Cout + 63) / 64, k_tiles = (g.Cin + 63) / 64;
  const long long npix = (long long)g.B * g.VH * g.VW;
  long long sp_ = npix / 2048; if (sp_ < 1) sp_ = 1; if (sp_ > 64) sp_ = 64;
  int splits = (int)sp_;
  long long pps = (npix + splits - 1) / splits;
  pps = (pps + WG_PIX - 1) / WG_PIX * WG_PIX;
  splits = (int)((npix + pps - 1) / pps);
  const size_t wbytes = (size_t)a.KH * a.KW * g.Cout * g.Cin * sizeof(float);
  int e = check_cuda(cudaMemsetAsync(dw, 0, wbytes, st), "wgrad memset");
  if (e) return e;
  const int c8 = g.Cin / 8;
  if (g.Cout <= 16 && a.epilogue == NG_EPI_HEAD && g.Cin % 8 == 0 && c8 <= 256 && 256 % c8 == 0) {
    // single real output channel (the caller's dY holds zeros in the padding channels)
    // short splits: ~8 blocks per SM over all taps, at least 4 trips of the 4-deep load batch per lane
    const int lanes1 = 256 / c8;
    long long want = (8ll * num_sms() + g.ntaps - 1) / g.ntaps;
    const long long most = npix / (16ll * lanes1);
    if (want > most) want = most;
    if (want < 1) want = 1;
    if (npix >= (1ll << 31)) { set_error("wgrad: %lld pixels exceed the 32-bit index range", npix); return NG_E_SHAPE; }
    const int pps1 = (int)((npix + want - 1) / want);
    dim3 grid(g.ntaps, (unsigned)((npix + pps1 - 1) / pps1));
    wgrad_cout1_kernel<T><<<grid, 256, 0, st>>>(g, (const T*)a.x, (const T*)a.y, dw, pps1);
    NG_LAUNCH_CHECK("wgrad_cout1_kernel");
    return NG_OK;
  }
  const long long blocks = (long long)g.ntaps * n_tiles * k_tiles * splits;
  wgrad_simt_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(g, (const T*)a.x, (const T*)a.y, dw, n_tiles, k_tiles,
                                                        splits, (int)pps);
  NG_LAUNCH_CHECK("wgrad_simt_kernel");
  return NG_OK;
}

template <typename T>
static int launch_bias_grad(const ng_conv_args& a, const ConvGeom& g, float* dbias, cudaStream_t st) {
  int e = check_cuda(cudaMemsetAsync(dbias, 0, (size_t)g.Cout * sizeof(float), st), "dbias memset");
  if (e) return e;
  const long long rows = (long long)g.B * g.Hout * g.Wout;
  if constexpr (sizeof(T) == 2) {
    const int c8 = g.Cout / 8;
    if (g.Cout % 8 == 0 && c8 >= 1 && c8 <= 256 && 256 % c8 == 0 && rows >= 4096) {
      const int lanes = 256 / c8;
      long long blocks = (rows + 4ll * lanes - 1) / (4ll * lanes);
      if (blocks > 8ll * num_sms()) blocks = 8ll * num_sms();
      colsum_vec_kernel<T><<<(unsigned)blocks, 256, 0, st>>>((const T*)a.y, rows, g.Cout, dbias);
      NG_LAUNCH_CHECK("colsum_vec_kernel");
      return NG_OK;
    }
  }
  const unsigned cs_blocks = (unsigned)(rows < 1024 ? rows : 1024);
  const unsigned cs_threads = (unsigned)(g.Cout < 256 ? g.Cout : 256);
  colsum_kernel<T><<<cs_blocks, cs_threads, 0, st>>>((const T*)a.y, rows, g.Cout, dbias);
  NG_LAUNCH_CHECK("colsum_kernel");
  return NG_OK;
}

int bias_grad(const ng_conv_args& a, const ConvGeom& g, float* dbias, cudaStream_t st) {
  switch (a.dtype) {
    case NG_F32:  return launch_bias_grad<float>(a, g, dbias, st);
    case NG_F16:  return launch_bias_grad<__half>(a, g, dbias, st);
    case NG_BF16: return launch_bias_grad<__nv_bfloat16>(a, g, dbias, st);
  }
  set_error("bias_grad: bad dtype %d", a.dtype);
  return NG_E_ARG;
}

int wgrad_simt(const ng_conv_args& a, const ConvGeom& g, float* dw, void* workspace, long long workspace_bytes,
               cudaStream_t st) {
  switch (a.dtype) {
    case NG_F32:  return launch_wgrad<float>(a, g, dw, workspace, workspace_bytes, st);
    case NG_F16:  return launch_wgrad<__half>(a, g, dw, workspace, workspace_bytes, st);
    case NG_BF16: return launch_wgrad<__nv_bfloat16>(a, g, dw, workspace, workspace_bytes, st);
  }
  set_error("wgrad_simt: bad dtype %d", a.dtype);
  return NG_E_ARG;
}

}  // namespace ng
