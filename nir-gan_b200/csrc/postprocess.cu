// Post-processing that follows the generator in the reference's inference loop (create_synthetic_dataset.py:34-52,
// 111-118): nearest / bilinear plane resize and per-tile histogram matching
// (skimage.exposure.match_histograms, channel_axis=None = _match_cumulative_cdf) on the device, with optional fp16 output
// (the reference stores float16).
//
// Histogram matching of one tile (N source pixels, M reference pixels):
//   q(v)   = #{source pixels <= v} / N                      (np.unique counts + cumsum)
//   out(v) = np.interp(q(v), tmpl_quantiles, tmpl_values)   with tmpl_* from np.unique(reference)
// Both arrays are sorted per tile by a segmented LSD radix sort (8-bit digits, 4 passes, stable scatter built on
// warp match_any ranks); each pixel then needs three binary searches (its own rank, and the bracket of its quantile among
// the reference's unique values) -- the sorted arrays of a batch stay in the 126 MB L2.  Integer cross-multiplication
// replaces the float comparison of quantiles (cnt/N vs e/M), which is exact; the interpolation itself runs in fp64 like
// np.interp.
#include "common.cuh"

namespace ng {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_PER_WARP = 512;
constexpr int RS_TILE = RS_WARPS * RS_PER_WARP;      // 4096 keys per block per pass

__device__ __forceinline__ uint32_t float_to_key(float f) {
  if (f == 0.f) f = 0.f;                              // -0.0 and +0.0 are one value for np.unique
  const uint32_t b = __float_as_uint(f);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

template <bool FROM_FLOAT>
__device__ __forceinline__ uint32_t load_key(const void* src, size_t i) {
  if constexpr (FROM_FLOAT) return float_to_key(reinterpret_cast<const float*>(src)[i]);
  else return reinterpret_cast<const uint32_t*>(src)[i];
}

// pass kernel 1: per-(tile, segment) digit histogram
template <bool FROM_FLOAT>
__global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const void* __restrict__ src, int n, int shift, int ntiles, uint32_t* __restrict__ ghist) {
  __shared__ uint32_t h[256];
  const int tile = blockIdx.x, seg = blockIdx.y;
  h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)seg * n;
  const int i0 = tile * RS_TILE, i1 = min(n, i0 + RS_TILE);
  for (int i = i0 + threadIdx.x; i < i1; i += RS_THREADS)
    atomicAdd(&h[(load_key<FROM_FLOAT>(src, base + i) >> shift) & 255u], 1u);
  __syncthreads();
  ghist[((size_t)seg * ntiles + tile) * 256 + threadIdx.x] = h[threadIdx.x];
}

// pass kernel 2: per segment, turn the histograms into global destination offsets [tile][digit]
__global__ void __launch_bounds__(256)
rs_scan_kernel(uint32_t* __restrict__ ghist, int ntiles) {
  __shared__ uint32_t tot[256];
  uint32_t* h = ghist + (size_t)blockIdx.x * ntiles * 256;
  const int d = threadIdx.x;
  uint32_t s = 0;
  for (int t = 0; t < ntiles; ++t) s += h[(size_t)t * 256 + d];
  tot[d] = s;
  __syncthreads();
  if (d == 0) {                                       // 256-entry exclusive scan: negligible
    uint32_t run = 0;
    for (int k = 0; k < 256; ++k) { const uint32_t v = tot[k]; tot[k] = run; run += v; }
  }
  __syncthreads();
  uint32_t run = tot[d];
  for (int t = 0; t < ntiles; ++t) {
    const uint32_t v = h[(size_t)t * 256 + d];
    h[(size_t)t * 256 + d] = run;
    run += v;
  }
}

// pass kernel 3: stable scatter.  Warp w owns keys [tile*4096 + w*512, +512) and walks them 32 at a time in order.
template <bool FROM_FLOAT>
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const void* __restrict__ src, uint32_t* __restrict__ dst, int n, int shift, int ntiles,
                  const uint32_t* __restrict__ goff) {
  __shared__ uint32_t wh[RS_WARPS][256];
  const int tile = blockIdx.x, seg = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wh[0][0])[i] = 0;
  __syncthreads();
  const size_t base = (size_t)seg * n;
  const int w0 = tile * RS_TILE + warp * RS_PER_WARP;
  // A: per-warp digit counts
  for (int c = 0; c < RS_PER_WARP; c += 32) {
    const int i = w0 + c + lane;
    if (i < n) atomicAdd(&wh[warp][(load_key<FROM_FLOAT>(src, base + i) >> shift) & 255u], 1u);
  }
  __syncthreads();
  // B: destination base of (warp, digit) = tile offset of the digit + counts of the lower warps
  {
    const int d = threadIdx.x;
    uint32_t run = goff[((size_t)seg * ntiles + tile) * 256 + d];
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) { const uint32_t v = wh[w][d]; wh[w][d] = run; run += v; }
  }
  __syncthreads();
  // C: ordered placement
  uint32_t* out = dst + base;
  for (int c = 0; c < RS_PER_WARP; c += 32) {
    const int i = w0 + c + lane;
    const bool act = i < n;
    const uint32_t key = act ? load_key<FROM_FLOAT>(src, base + i) : 0u;
    const uint32_t d = act ? ((key >> shift) & 255u) : (256u + lane);        // inactive lanes match nobody
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int leader = __ffs(peers) - 1;
    uint32_t b = 0;
    if (act && lane == leader) { b = wh[warp][d]; wh[warp][d] = b + __popc(peers); }
    b = __shfl_sync(0xffffffffu, b, leader);
    if (act) out[b + rank] = key;
    __syncwarp();
  }
}

__device__ __forceinline__ int count_le(const uint32_t* __restrict__ a, int n, uint32_t key) {   // upper_bound
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] <= key) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int count_lt(const uint32_t* __restrict__ a, int n, uint32_t key) {   // lower_bound
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
  return lo;
}

template <typename TO>
__global__ void __launch_bounds__(256)
hist_match_kernel(const float* __restrict__ img, const uint32_t* __restrict__ ssrc, const uint32_t* __restrict__ sref,
                  int B, int N, int M, TO* __restrict__ out) {
  const long long total = (long long)B * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / N);
    const uint32_t* s = ssrc + (size_t)b * N;
    const uint32_t* r = sref + (size_t)b * M;
    const uint32_t key = float_to_key(img[i]);
    const long long cnt = count_le(s, N, key);                 // source quantile = cnt / N
    long long i0 = (cnt * M + N - 1) / N - 1;                   // smallest i with (i+1)/M >= cnt/N
    i0 = i0 < 0 ? 0 : (i0 > M - 1 ? M - 1 : i0);
    const uint32_t rk = r[i0];
    const int first = count_lt(r, M, rk);                      // previous unique value ends at first - 1
    float res = key_to_float(rk);
    if (first > 0) {
      const int last1 = count_le(r, M, rk);                    // this unique value's quantile = last1 / M
      const double q = (double)cnt / (double)N, qk = (double)last1 / (double)M, qp = (double)first / (double)M;
      const double vk = (double)key_to_float(rk), vp = (double)key_to_float(r[first - 1]);
      res = (float)(vp + (q - qp) * ((vk - vp) / (qk - qp)));
    }
    if constexpr (sizeof(TO) == 2) out[i] = __float2half_rn(res);
    else out[i] = res;
  }
}

// plane resize, PyTorch semantics: mode 0 = 'nearest' (src = floor(dst * in/out)), mode 1 = 'bilinear',
// align_corners=False (src = (dst + 0.5) * in/out - 0.5, clamped at 0; neighbours clamped to the last index)
__global__ void __launch_bounds__(256)
resize_plane_kernel(const float* __restrict__ src, int P, int h, int w, int H, int W, int mode, float* __restrict__ dst) {
  const long long total = (long long)P * H * W;
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long p = i / ((long long)W * H);
    const float* s = src + p * h * w;
    if (mode == 0) {
      const int yy = min((int)floorf(y * sy), h - 1), xx = min((int)floorf(x * sx), w - 1);
      dst[i] = s[yy * w + xx];
    } else {
      const float fy = fmaxf((y + 0.5f) * sy - 0.5f, 0.f), fx = fmaxf((x + 0.5f) * sx - 0.5f, 0.f);
      const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
      const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
      const float ly = fy - y0, lx = fx - x0;
      dst[i] = (1.f - ly) * ((1.f - lx) * s[y0 * w + x0] + lx * s[y0 * w + x1]) +
               ly * ((1.f - lx) * s[y1 * w + x0] + lx * s[y1 * w + x1]);
    }
  }
}

__global__ void __launch_bounds__(256)
keys_to_float_kernel(const uint32_t* __restrict__ k, long long n, float* __restrict__ o) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    o[i] = key_to_float(k[i]);
}

static inline int rs_tiles(int n) { return (n + RS_TILE - 1) / RS_TILE; }

// sorts `segs` segments of n floats each into sorted (ordered-uint keys); tmp and ghist are scratch
static int segmented_sort(const float* src, int segs, int n, uint32_t* sorted, uint32_t* tmp, uint32_t* ghist,
                          cudaStream_t st) {
  const int nt = rs_tiles(n);
  dim3 grid(nt, segs);
  const void* in = src;
  uint32_t* bufs[2] = {tmp, sorted};                 // pass 0 -> tmp, 1 -> sorted, 2 -> tmp, 3 -> sorted
  for (int pass = 0; pass < 4; ++pass) {
    uint32_t* outb = bufs[pass & 1];
    if (pass == 0) rs_hist_kernel<true><<<grid, RS_THREADS, 0, st>>>(in, n, 0, nt, ghist);
    else rs_hist_kernel<false><<<grid, RS_THREADS, 0, st>>>(in, n, 8 * pass, nt, ghist);
    NG_LAUNCH_CHECK("rs_hist_kernel");
    rs_scan_kernel<<<segs, 256, 0, st>>>(ghist, nt);
    NG_LAUNCH_CHECK("rs_scan_kernel");
    if (pass == 0) rs_scatter_kernel<true><<<grid, RS_THREADS, 0, st>>>(in, outb, n, 0, nt, ghist);
    else rs_scatter_kernel<false><<<grid, RS_THREADS, 0, st>>>(in, outb, n, 8 * pass, nt, ghist);
    NG_LAUNCH_CHECK("rs_scatter_kernel");
    in = outb;
  }
  return NG_OK;
}

}  // namespace ng

using namespace ng;

extern "C" int64_t ng_hist_match_workspace_bytes(int32_t B, int32_t N, int32_t M) {
  if (B <= 0 || N <= 0 || M <= 0) return NG_E_ARG;
  const int64_t big = N > M ? N : M;
  return ((int64_t)B * N + (int64_t)B * M + (int64_t)B * big) * 4 + (int64_t)B * rs_tiles((int)big) * 256 * 4 + 256;
}

extern "C" int ng_sort_segments(const float* src, int32_t segs, int32_t n, float* sorted_out, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(src && sorted_out && workspace && segs > 0 && n > 0, NG_E_ARG, "sort_segments: bad arguments");
  const int64_t need = (int64_t)segs * n * 8 + (int64_t)segs * rs_tiles(n) * 1024;
  NG_REQUIRE(workspace_bytes >= need, NG_E_ARG, "sort_segments: workspace of %lld bytes needed", (long long)need);
  uint32_t* sorted = reinterpret_cast<uint32_t*>(workspace);
  uint32_t* tmp = sorted + (size_t)segs * n;
  uint32_t* ghist = tmp + (size_t)segs * n;
  r = segmented_sort(src, segs, n, sorted, tmp, ghist, (cudaStream_t)stream);
  if (r) return r;
  const long long total = (long long)segs * n;
  long long blocks = (total + 255) / 256; if (blocks > 148 * 8) blocks = 148 * 8;
  keys_to_float_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sorted, total, sorted_out);
  NG_LAUNCH_CHECK("key_to_float kernel");
  return NG_OK;
}

extern "C" int ng_hist_match(const float* image, const float* reference, int32_t B, int32_t N, int32_t M,
                             int32_t out_dtype, void* out, void* workspace, int64_t workspace_bytes, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(image && reference && out && workspace && B > 0 && N > 0 && M > 0, NG_E_ARG, "hist_match: bad arguments");
  NG_REQUIRE(out_dtype == NG_F32 || out_dtype == NG_F16, NG_E_ARG, "hist_match: output must be f32 or f16");
  NG_REQUIRE(workspace_bytes >= ng_hist_match_workspace_bytes(B, N, M), NG_E_ARG, "hist_match: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* ssrc = reinterpret_cast<uint32_t*>(workspace);
  uint32_t* sref = ssrc + (size_t)B * N;
  uint32_t* tmp = sref + (size_t)B * M;
  const int big = N > M ? N : M;
  uint32_t* ghist = tmp + (size_t)B * big;
  r = segmented_sort(image, B, N, ssrc, tmp, ghist, st);
  if (r) return r;
  r = segmented_sort(reference, B, M, sref, tmp, ghist, st);
  if (r) return r;
  const long long total = (long long)B * N;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (out_dtype == NG_F16)
    hist_match_kernel<__half><<<(unsigned)blocks, 256, 0, st>>>(image, ssrc, sref, B, N, M, (__half*)out);
  else
    hist_match_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(image, ssrc, sref, B, N, M, (float*)out);
  NG_LAUNCH_CHECK("hist_match_kernel");
  return NG_OK;
}

extern "C" int ng_resize_plane(const float* src, int32_t planes, int32_t h, int32_t w, int32_t H, int32_t W, int32_t mode,
                               float* dst, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(src && dst && planes > 0 && h > 0 && w > 0 && H > 0 && W > 0 && (mode == 0 || mode == 1), NG_E_ARG,
             "resize_plane: bad arguments");
  const long long total = (long long)planes * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  resize_plane_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, planes, h, w, H, W, mode, dst);
  NG_LAUNCH_CHECK("resize_plane_kernel");
  return NG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Validation metrics on the device (utils/calculate_metrics.py:6-37): L1, L2, PSNR and the mean
// SSIM map of kornia.metrics.ssim (kornia==0.7.3, requirements.txt:9): Gaussian window (sigma 1.5), 'same' padding with
// a reflect border, C1 = (0.01 max)^2, C2 = (0.03 max)^2, eps 1e-12:
//   num / (den + eps),  num = (2 mu1 mu2 + C1)(2 s12 + C2),  den = (mu1^2 + mu2^2 + C1)(s1 + s2 + C2)
// One block = a 32 x 8 output tile of one plane; the (32+w-1) x (8+w-1) haloed tiles of both images are staged in shared
// memory and the five windowed moments are accumulated directly (window 5 or 11).
// ---------------------------------------------------------------------------------------------------------------------
namespace ng {

constexpr int MT_W = 32, MT_H = 8, MT_MAXWIN = 11;

// GRAD: also writes, per pixel, the partial derivatives of the SSIM map w.r.t. the windowed moments of image a
// (d/d mu_a with the variance terms chained in, d/d E[a^2], d/d E[ab]) for ssim_bwd_kernel.
template <bool GRAD>
__global__ void __launch_bounds__(256)
image_metrics_kernel(const float* __restrict__ a, const float* __restrict__ b, int H, int W, int win, float c1, float c2,
                     const float* __restrict__ gauss, float* __restrict__ partial, int tiles_x, int tiles_y,
                     float4* __restrict__ dmaps) {
  __shared__ float sa[(MT_H + MT_MAXWIN - 1) * (MT_W + MT_MAXWIN - 1)];
  __shared__ float sb[(MT_H + MT_MAXWIN - 1) * (MT_W + MT_MAXWIN - 1)];
  __shared__ float gk[MT_MAXWIN];
  __shared__ float red[3][8];
  const int half = win / 2, PW = MT_W + win - 1, PH = MT_H + win - 1;
  int t = blockIdx.x;
  const int tx0 = (t % tiles_x) * MT_W; t /= tiles_x;
  const int ty0 = (t % tiles_y) * MT_H;
  const size_t plane = (size_t)(t / tiles_y) * H * W;
  if (threadIdx.x < win) gk[threadIdx.x] = gauss[threadIdx.x];
  for (int i = threadIdx.x; i < PH * PW; i += 256) {
    const int py = i / PW, px = i - py * PW;
    const int gy = reflect_idx(min(max(ty0 + py - half, -(H - 1)), 2 * (H - 1)), H);
    const int gx = reflect_idx(min(max(tx0 + px - half, -(W - 1)), 2 * (W - 1)), W);
    sa[i] = a[plane + (size_t)gy * W + gx];
    sb[i] = b[plane + (size_t)gy * W + gx];
  }
  __syncthreads();
  const int lx = threadIdx.x % MT_W, ly = threadIdx.x / MT_W;
  const int ox = tx0 + lx, oy = ty0 + ly;
  float l1 = 0.f, l2 = 0.f, ss = 0.f;
  if (ox < W && oy < H) {
    float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
    for (int ky = 0; ky < win; ++ky) {
      float r1 = 0.f, r2 = 0.f, r11 = 0.f, r22 = 0.f, r12 = 0.f;
      for (int kx = 0; kx < win; ++kx) {
        const float x = sa[(ly + ky) * PW + lx + kx], y = sb[(ly + ky) * PW + lx + kx], g = gk[kx];
        r1 = fmaf(g, x, r1); r2 = fmaf(g, y, r2);
        r11 = fmaf(g, x * x, r11); r22 = fmaf(g, y * y, r22); r12 = fmaf(g, x * y, r12);
      }
      const float g = gk[ky];
      m1 = fmaf(g, r1, m1); m2 = fmaf(g, r2, m2);
      e11 = fmaf(g, r11, e11); e22 = fmaf(g, r22, e22); e12 = fmaf(g, r12, e12);
    }
    const float s1 = e11 - m1 * m1, s2 = e22 - m2 * m2, s12 = e12 - m1 * m2;
    const float num = (2.f * m1 * m2 + c1) * (2.f * s12 + c2);
    const float den = (m1 * m1 + m2 * m2 + c1) * (s1 + s2 + c2);
    ss = num / (den + 1e-12f);
    if (GRAD) {
      const float A1 = 2.f * m1 * m2 + c1, A2 = 2.f * s12 + c2, B1 = m1 * m1 + m2 * m2 + c1, B2 = s1 + s2 + c2;
      const float inv = 1.f / (den + 1e-12f);
      const float d_e11 = -ss * B1 * inv;                         // via s1 in B2
      const float d_e12 = 2.f * A1 * inv;                         // via s12 in A2
      const float d_m1 = 2.f * m2 * (A2 - A1) * inv - ss * 2.f * m1 * (B2 - B1) * inv;
      dmaps[plane + (size_t)oy * W + ox] = make_float4(d_m1, d_e11, d_e12, 0.f);
    }
    const float d = sa[(ly + half) * PW + lx + half] - sb[(ly + half) * PW + lx + half];
    l1 = fabsf(d);
    l2 = d * d;
  }
  l1 = warp_sum(l1); l2 = warp_sum(l2); ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = l2; red[2][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[threadIdx.x][i];
    partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
  }
  if (threadIdx.x == 3) partial[(size_t)blockIdx.x * 4 + 3] = 0.f;
}

// out[0] = L1, out[1] = L2 (MSE), out[2] = PSNR = 10 log10(max^2 / MSE), out[3] = mean SSIM   (fixed-order sums, fp64)
__global__ void image_metrics_finalize_kernel(const float* __restrict__ partial, int nblocks, double inv_n, float max_val,
                                              float* __restrict__ out) {
  const int j = threadIdx.x;
  if (j >= 3) return;
  double s = 0.0;
  for (int i = 0; i < nblocks; ++i) s += partial[(size_t)i * 4 + j];
  const double mean = s * inv_n;
  if (j == 0) out[0] = (float)mean;
  if (j == 1) { out[1] = (float)mean; out[2] = (float)(10.0 * log10((double)max_val * max_val / mean)); }
  if (j == 2) out[3] = (float)mean;
}

}  // namespace ng

extern "C" int64_t ng_image_metrics_scratch_floats(int32_t planes, int32_t H, int32_t W) {
  if (planes <= 0 || H <= 0 || W <= 0) return NG_E_ARG;
  return (int64_t)planes * ((H + MT_H - 1) / MT_H) * ((W + MT_W - 1) / MT_W) * 4 + MT_MAXWIN;
}

extern "C" int ng_image_metrics(const float* pred, const float* target, int32_t planes, int32_t H, int32_t W,
                                int32_t window, float max_val, float* out4, float* scratch, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(pred && target && out4 && scratch && planes > 0 && H > 0 && W > 0, NG_E_ARG, "image_metrics: bad arguments");
  NG_REQUIRE(window % 2 == 1 && window >= 1 && window <= MT_MAXWIN, NG_E_UNSUPPORTED, "image_metrics: odd window <= 11");
  NG_REQUIRE(window / 2 < H && window / 2 < W, NG_E_SHAPE, "image_metrics: window larger than the image");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = (W + MT_W - 1) / MT_W, tiles_y = (H + MT_H - 1) / MT_H;
  const long long blocks = (long long)planes * tiles_x * tiles_y;
  NG_REQUIRE(blocks < (1ll << 31), NG_E_SHAPE, "image_metrics: too many tiles");
  // normalised Gaussian window, sigma 1.5 (kornia get_gaussian_kernel1d), computed on the host in fp32 like torch does
  float g[MT_MAXWIN];
  float sum = 0.f;
  for (int i = 0; i < window; ++i) { const float x = (float)(i - window / 2); g[i] = expf(-(x * x) / (2.f * 1.5f * 1.5f)); sum += g[i]; }
  for (int i = 0; i < window; ++i) g[i] /= sum;
  float* gdev = scratch + blocks * 4;
  int e = check_cuda(cudaMemcpyAsync(gdev, g, sizeof(float) * window, cudaMemcpyHostToDevice, st), "image_metrics window");
  if (e) return e;
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  image_metrics_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(pred, target, H, W, window, c1, c2, gdev, scratch, tiles_x,
                                                                tiles_y, nullptr);
  NG_LAUNCH_CHECK("image_metrics_kernel");
  image_metrics_finalize_kernel<<<1, 32, 0, st>>>(scratch, (int)blocks, 1.0 / ((double)planes * H * W), max_val, out4);
  NG_LAUNCH_CHECK("image_metrics_finalize_kernel");
  return NG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Optional generator losses (utils/losses.py:10-29 ssim_loss, :64-78 emd_loss = pix2pix.py's hist_loss; SURVEY.md 8f
// rank 4), forward and d/dpred.
//   ssim_loss = 1 - mean(kornia.metrics.ssim(pred, target, 11)).  The forward is image_metrics_kernel<true>; the
//   backward is the adjoint of the reflect-bordered Gaussian filter applied to the three derivative maps:
//     d/dx_j = sum_q in fold(j) sum_k g[k] (Dm + 2 x_j De11 + y_j De12)[q - half + k]     (maps zero outside the image)
//   where fold(j) = {j} plus the out-of-image positions the reflect border maps onto j.
//   emd_loss = mean |cumsum(softmax(pred_b)) - cumsum(softmax(target_b))| over all B * N entries; with s = sign of the
//   CDF difference, E_k = sum_{i<k} s_i and A = sum_j p_j E_j:  d/dx_k = p_k (A - E_k) / (B N).
// ---------------------------------------------------------------------------------------------------------------------
namespace ng {

__global__ void ssim_loss_finalize_kernel(const float* __restrict__ partial, int nblocks, double inv_n,
                                          float* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 1024) s += partial[(size_t)i * 4 + 2];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 32; ++i) t += red[i];
    out[0] = (float)(1.0 - t * inv_n);
  }
}

__device__ __forceinline__ int fold_list(int j, int n, int half, int* q) {
  int c = 0;
  q[c++] = j;
  if (j >= 1 && j <= half) q[c++] = -j;
  if (j <= n - 2 && j >= n - 1 - half) q[c++] = 2 * (n - 1) - j;
  return c;
}

__global__ void __launch_bounds__(256)
ssim_bwd_kernel(const float4* __restrict__ dmaps, const float* __restrict__ x, const float* __restrict__ y, int H, int W,
                int win, const float* __restrict__ gauss, float scale, long long total, float* __restrict__ dx) {
  __shared__ float gk[MT_MAXWIN];
  if (threadIdx.x < win) gk[threadIdx.x] = gauss[threadIdx.x];
  __syncthreads();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int half = win / 2;
  const int jx = (int)(idx % W), jy = (int)((idx / W) % H);
  const size_t plane = (size_t)(idx / ((long long)W * H)) * H * W;
  const float cx = 2.f * x[idx], cy = y[idx];
  int qx[3], qy[3];
  const int nx = fold_list(jx, W, half, qx), ny = fold_list(jy, H, half, qy);
  float acc = 0.f;
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      const int y0 = qy[a] - half, x0 = qx[b] - half;
      for (int ky = max(0, -y0); ky < win && y0 + ky < H; ++ky) {
        const float4* row = dmaps + plane + (size_t)(y0 + ky) * W;
        float r = 0.f;
        for (int kx = max(0, -x0); kx < win && x0 + kx < W; ++kx) {
          const float4 d = __ldg(row + x0 + kx);
          r = fmaf(gk[kx], fmaf(cy, d.z, fmaf(cx, d.y, d.x)), r);
        }
        acc = fmaf(gk[ky], r, acc);
      }
    }
  dx[idx] = scale * acc;
}

constexpr int EMD_T = 1024;

// inclusive scan over the 1024 threads of a block; `total` = the block sum.  ws: 32 entries of shared memory.
template <typename T>
__device__ __forceinline__ T emd_block_scan(T v, T* ws, T& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const T n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  __syncthreads();                       // previous use of ws is finished
  if (lane == 31) ws[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = ws[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const T n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    ws[lane] = w;
  }
  __syncthreads();
  total = ws[31];
  return v + (warp ? ws[warp - 1] : (T)0);
}

template <typename T, typename Op>
__device__ __forceinline__ T emd_block_reduce(T v, T* ws, Op op) {
  for (int o = 16; o; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = ws[0];
  for (int i = 1; i < 32; ++i) r = op(r, ws[i]);
  return r;
}

__global__ void __launch_bounds__(EMD_T)
emd_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target, int N, double* __restrict__ partial,
                float* __restrict__ grad, float gscale) {
  __shared__ float wf[32];
  __shared__ int wi[32];
  __shared__ double wd[32];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* x = pred + (size_t)b * N;
  const float* y = target + (size_t)b * N;
  float mx = -INFINITY, my = -INFINITY;
  for (int i = t; i < N; i += EMD_T) { mx = fmaxf(mx, x[i]); my = fmaxf(my, y[i]); }
  auto fmx = [](float a, float c) { return fmaxf(a, c); };
  auto dsum = [](double a, double c) { return a + c; };
  mx = emd_block_reduce(mx, wf, fmx);
  my = emd_block_reduce(my, wf, fmx);
  double sx = 0.0, sy = 0.0;
  for (int i = t; i < N; i += EMD_T) { sx += expf(x[i] - mx); sy += expf(y[i] - my); }
  sx = emd_block_reduce(sx, wd, dsum);
  sy = emd_block_reduce(sy, wd, dsum);
  const float inv_sx = (float)(1.0 / sx), inv_sy = (float)(1.0 / sy);
  double carry = 0.0, loss = 0.0, A = 0.0;
  long long carry_s = 0;
  for (int base = 0; base < N; base += EMD_T) {
    const int i = base + t;
    const bool valid = i < N;
    const float p = valid ? expf(x[i] - mx) * inv_sx : 0.f;
    const float q = valid ? expf(y[i] - my) * inv_sy : 0.f;
    float tot;
    const float incl = emd_block_scan(p - q, wf, tot);
    const float diff = (float)(carry + (double)incl);
    const int s = valid ? ((diff > 0.f) - (diff < 0.f)) : 0;
    int stot;
    const int sincl = emd_block_scan(s, wi, stot);
    if (valid) {
      const long long E = carry_s + sincl - s;
      loss += fabsf(diff);
      A += (double)p * (double)E;
      if (grad) grad[(size_t)b * N + i] = (float)E;
    }
    carry += (double)tot;
    carry_s += stot;
  }
  loss = emd_block_reduce(loss, wd, dsum);
  A = emd_block_reduce(A, wd, dsum);
  if (t == 0) partial[b] = loss;
  if (grad) {
    const float Af = (float)A;
    for (int i = t; i < N; i += EMD_T) {
      const float p = expf(x[i] - mx) * inv_sx;
      grad[(size_t)b * N + i] = p * (Af - grad[(size_t)b * N + i]) * gscale;
    }
  }
}

__global__ void emd_finalize_kernel(const double* __restrict__ partial, int B, double inv_n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < B; ++i) s += partial[i];
    out[0] = (float)(s * inv_n);
  }
}

}  // namespace ng

extern "C" int64_t ng_ssim_loss_scratch_floats(int32_t planes, int32_t H, int32_t W) {
  if (planes <= 0 || H <= 0 || W <= 0) return NG_E_ARG;
  // tile partials + window (image_metrics layout, rounded to 16 bytes) + float4 derivative maps
  const int64_t head = (ng_image_metrics_scratch_floats(planes, H, W) + 3) / 4 * 4;
  return head + (int64_t)planes * H * W * 4;
}

extern "C" int ng_ssim_loss(const float* pred, const float* target, int32_t planes, int32_t H, int32_t W, int32_t window,
                            float max_val, float* out1, float* dpred, float* scratch, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(pred && target && out1 && scratch && planes > 0 && H > 0 && W > 0, NG_E_ARG, "ssim_loss: bad arguments");
  NG_REQUIRE(window % 2 == 1 && window >= 1 && window <= MT_MAXWIN, NG_E_UNSUPPORTED, "ssim_loss: odd window <= 11");
  NG_REQUIRE(window / 2 < H && window / 2 < W, NG_E_SHAPE, "ssim_loss: window larger than the image");
  NG_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, NG_E_ALIGN, "ssim_loss: scratch must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_x = (W + MT_W - 1) / MT_W, tiles_y = (H + MT_H - 1) / MT_H;
  const long long blocks = (long long)planes * tiles_x * tiles_y;
  NG_REQUIRE(blocks < (1ll << 31), NG_E_SHAPE, "ssim_loss: too many tiles");
  float g[MT_MAXWIN];
  float sum = 0.f;
  for (int i = 0; i < window; ++i) { const float x = (float)(i - window / 2); g[i] = expf(-(x * x) / (2.f * 1.5f * 1.5f)); sum += g[i]; }
  for (int i = 0; i < window; ++i) g[i] /= sum;
  float* gdev = scratch + blocks * 4;
  int e = check_cuda(cudaMemcpyAsync(gdev, g, sizeof(float) * window, cudaMemcpyHostToDevice, st), "ssim_loss window");
  if (e) return e;
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  const int64_t head = (ng_image_metrics_scratch_floats(planes, H, W) + 3) / 4 * 4;
  float4* dmaps = reinterpret_cast<float4*>(scratch + head);
  const long long total = (long long)planes * H * W;
  if (dpred)
    image_metrics_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(pred, target, H, W, window, c1, c2, gdev, scratch, tiles_x,
                                                                 tiles_y, dmaps);
  else
    image_metrics_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(pred, target, H, W, window, c1, c2, gdev, scratch, tiles_x,
                                                                  tiles_y, nullptr);
  NG_LAUNCH_CHECK("image_metrics_kernel");
  ssim_loss_finalize_kernel<<<1, 1024, 0, st>>>(scratch, (int)blocks, 1.0 / (double)total, out1);
  NG_LAUNCH_CHECK("ssim_loss_finalize_kernel");
  if (dpred) {
    ssim_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dmaps, pred, target, H, W, window, gdev,
                                                                     (float)(-1.0 / (double)total), total, dpred);
    NG_LAUNCH_CHECK("ssim_bwd_kernel");
  }
  return NG_OK;
}

extern "C" int ng_emd_loss(const float* pred, const float* target, int32_t B, int32_t N, float* out1, float* dpred,
                           double* scratch, void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(pred && target && out1 && scratch && B > 0 && N > 0, NG_E_ARG, "emd_loss: bad arguments");
  NG_REQUIRE(N < (1 << 24), NG_E_SHAPE, "emd_loss: at most 2^24 - 1 elements per sample");
  cudaStream_t st = (cudaStream_t)stream;
  const double inv_n = 1.0 / ((double)B * (double)N);
  emd_loss_kernel<<<(unsigned)B, EMD_T, 0, st>>>(pred, target, N, scratch, dpred, (float)inv_n);
  NG_LAUNCH_CHECK("emd_loss_kernel");
  emd_finalize_kernel<<<1, 32, 0, st>>>(scratch, B, inv_n, out1);
  NG_LAUNCH_CHECK("emd_finalize_kernel");
  return NG_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// SatCLIP location encoder: the step right before the injected generator,
// SatClIP_wrapper.predict (model/satclip/satclip_wrapper.py:29-34) = LocationEncoder(posenc, nnet)
// (model/satclip/location_encoder.py:267-275), float64 like the reference (`x.double()`), float32 result.
//   posenc: SphericalHarmonics(L) (positional_encoding/spherical_harmonics.py:27-42): phi = deg2rad(lon + 180),
//           theta = deg2rad(lat + 90), the L*L real harmonics from the closed-form recursion
//           (positional_encoding/spherical_harmonics_closed_form.py:8-40), order l = 0..L-1, m = -l..l
//   nnet:   SirenNet (location_encoder.py:73-151), eval mode: hidden layers sin(w0_i * (W x + b)) with w0 = w0_initial
//           for the first layer and w0 for the others, linear last layer.
// One block per coordinate pair; thread t evaluates harmonic t, then output j = t (+256 ...) of every layer; activations
// ping-pong between two shared-memory vectors.  Weights are passed transposed ([in][out]) so a layer's reads coalesce.
// ---------------------------------------------------------------------------------------------------------------------
namespace ng {

constexpr int SC_MAXDIM = 1024;

__device__ double sc_assoc_legendre(int l, int m, double x) {
  double pmm = 1.0;
  if (m > 0) {
    const double somx2 = sqrt((1.0 - x) * (1.0 + x));
    double fact = 1.0;
    for (int i = 1; i <= m; ++i) { pmm = pmm * (-fact) * somx2; fact += 2.0; }
  }
  if (l == m) return pmm;
  double pmmp1 = x * (2.0 * m + 1.0) * pmm;
  if (l == m + 1) return pmmp1;
  double pll = 0.0;
  for (int ll = m + 2; ll <= l; ++ll) {
    pll = ((2.0 * ll - 1.0) * x * pmmp1 - (ll + m - 1.0) * pmm) / (double)(ll - m);
    pmm = pmmp1;
    pmmp1 = pll;
  }
  return pll;
}

__device__ double sc_renorm(int l, int m) {      // sqrt((2l+1) (l-m)! / (4 pi (l+m)!))
  double ratio = 1.0;
  for (int k = l - m + 1; k <= l + m; ++k) ratio /= (double)k;
  return sqrt((2.0 * l + 1.0) * ratio / (4.0 * 3.14159265358979323846));
}

__global__ void __launch_bounds__(256)
satclip_encode_kernel(const double* __restrict__ lonlat, int L, const double* __restrict__ params, int hidden,
                      int num_layers, int dim_out, double w0_initial, double w0, float* __restrict__ out) {
  __shared__ double xa[SC_MAXDIM], xb[SC_MAXDIM];
  const int b = blockIdx.x, dim_in = L * L;
  const double lon = lonlat[2 * b], lat = lonlat[2 * b + 1];
  const double deg = 3.14159265358979323846 / 180.0;
  const double phi = (lon + 180.0) * deg, theta = (lat + 90.0) * deg;
  const double ct = cos(theta);
  for (int t = threadIdx.x; t < dim_in; t += 256) {
    const int l = (int)floor(sqrt((double)t) + 1e-9);
    const int m = t - l * l - l;
    double y;
    if (m == 0) y = sc_renorm(l, 0) * sc_assoc_legendre(l, 0, ct);
    else if (m > 0) y = 1.4142135623730951 * sc_renorm(l, m) * cos(m * phi) * sc_assoc_legendre(l, m, ct);
    else y = 1.4142135623730951 * sc_renorm(l, -m) * sin(-m * phi) * sc_assoc_legendre(l, -m, ct);
    xa[t] = y;
  }
  __syncthreads();
  double* xin = xa;
  double* xout = xb;
  const double* w = params;
  int d = dim_in;
  for (int layer = 0; layer < num_layers; ++layer) {
    const double* bias = w + (size_t)d * hidden;
    const double scale = layer == 0 ? w0_initial : w0;
    for (int j = threadIdx.x; j < hidden; j += 256) {
      double acc = bias[j];
      for (int i = 0; i < d; ++i) acc = fma(w[(size_t)i * hidden + j], xin[i], acc);
      xout[j] = sin(scale * acc);
    }
    __syncthreads();
    w = bias + hidden;
    d = hidden;
    double* tmp = xin; xin = xout; xout = tmp;
  }
  const double* bias = w + (size_t)d * dim_out;
  for (int j = threadIdx.x; j < dim_out; j += 256) {
    double acc = bias[j];
    for (int i = 0; i < d; ++i) acc = fma(w[(size_t)i * dim_out + j], xin[i], acc);
    out[(size_t)b * dim_out + j] = (float)acc;
  }
}

}  // namespace ng

extern "C" int ng_satclip_encode(const double* lonlat, int32_t B, int32_t L, const double* params_t, int32_t hidden,
                                 int32_t num_layers, int32_t dim_out, double w0_initial, double w0, float* out,
                                 void* stream) {
  int r = require_sm100(); if (r) return r;
  NG_REQUIRE(lonlat && params_t && out && B > 0, NG_E_ARG, "satclip_encode: bad arguments");
  NG_REQUIRE(L >= 1 && L * L <= SC_MAXDIM && hidden >= 1 && hidden <= SC_MAXDIM && dim_out >= 1 && num_layers >= 1,
             NG_E_SHAPE, "satclip_encode: L*L and hidden must be <= %d", SC_MAXDIM);
  satclip_encode_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(lonlat, L, params_t, hidden, num_layers, dim_out,
                                                                      w0_initial, w0, out);
  NG_LAUNCH_CHECK("satclip_encode_kernel");
  return NG_OK;
}
