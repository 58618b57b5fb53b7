// hg_disc.cu — the bandwidth-bound ends of the discriminator stacks (src/models.py:128-248):
//   * first layers (Cin = 1): DiscriminatorS Conv1d(1,128,15,1,pad 7) :196 and DiscriminatorP
//     Conv2d(1,32,(5,1),(3,1),pad (2,0)) :134, the latter with the right-side reflect pad and the
//     [B,1,T] -> [B,1,H,p] view of :146-151 folded into the address arithmetic (no padded copy);
//   * last layers (Cout = 1): conv_post Conv1d(1024,1,3,pad 1) :204 / Conv2d(1024,1,(3,1)) :141;
//   * AvgPool1d(4,2,padding=2) between MSD scales :227-230 (count_include_pad);
//   * export of an internal bf16 [B*p][rows][C] feature map to the reference's fp32 [B][C][H][p].
// The wide middle layers run on the tcgen05 kernel (hg_conv1d_general_fwd).
#include "hg_common.cuh"

#include <atomic>

#include "../../include/hifigan_b200.h"

extern std::atomic<int64_t> g_hg_launches;

namespace {

constexpr int kFirstThreads = 128;
constexpr int kFirstMaxK = 16;

// one thread = one output position of one (batch, period-column) sequence, all cout channels
__global__ void __launch_bounds__(kFirstThreads)
disc_first_conv_kernel(const float* __restrict__ y, const float* __restrict__ w,
                       const float* __restrict__ bias, int t, int period, int h_in, int h_out,
                       int h_rows_out, int k, int stride, int pad, int cout, float slope,
                       __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];  // [k][cout] weights, then [cout] bias
  float* ws = sm;
  float* bs = sm + k * cout;
  for (int i = threadIdx.x; i < k * cout; i += blockDim.x) {
    const int j = i / cout, co = i % cout;
    ws[i] = w[co * k + j];
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int seq = blockIdx.y;            // b * period + wcol
  const int b = seq / period, wcol = seq % period;
  const int ho = blockIdx.x * blockDim.x + threadIdx.x;
  if (ho >= h_out) return;
  float xin[kFirstMaxK];
#pragma unroll
  for (int j = 0; j < kFirstMaxK; ++j) {
    float v = 0.f;
    if (j < k) {
      const int h = ho * stride + j - pad;
      if (h >= 0 && h < h_in) {
        int i = h * period + wcol;               // index into the (virtually) reflect-padded signal
        if (i >= t) i = 2 * (t - 1) - i;         // right-side reflect, models.py:147-149
        v = y[static_cast<size_t>(b) * t + i];
      }
    }
    xin[j] = v;
  }
  __nv_bfloat16* orow = out + (static_cast<size_t>(seq) * h_rows_out + ho) * cout;
  for (int c0 = 0; c0 < cout; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = bs[c0 + e];
#pragma unroll
    for (int j = 0; j < kFirstMaxK; ++j) {
      if (j < k) {
        const float xv = xin[j];
        const float4* wj = reinterpret_cast<const float4*>(ws + j * cout + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 wq = wj[q];
          acc[4 * q] += xv * wq.x; acc[4 * q + 1] += xv * wq.y;
          acc[4 * q + 2] += xv * wq.z; acc[4 * q + 3] += xv * wq.w;
        }
      }
    }
    hg::U8 o;
#pragma unroll
    for (int q = 0; q < 8; ++q)
      o.v[q] = hg::pack_bf16x2(hg::lrelu(acc[2 * q], slope), hg::lrelu(acc[2 * q + 1], slope));
    hg::stg256(orow + c0, o);
  }
}

// one warp = one output position: y[s,h] = bias + sum_{j,c} x[s, h + j - k/2, c] * w[c][j]
__global__ void __launch_bounds__(256)
disc_last_conv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, int nseq, int h, int h_rows, int c, int k,
                      float* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nseq * h) return;
  const int s = warp / h, ho = warp % h;
  const int half = k / 2;
  float acc = 0.f;
  for (int j = 0; j < k; ++j) {
    const int hi = ho + j - half;
    if (hi < 0 || hi >= h) continue;
    const __nv_bfloat16* row = x + (static_cast<size_t>(s) * h_rows + hi) * c;
    for (int c0 = lane * 8; c0 < c; c0 += 256) {
      const uint4 r = *reinterpret_cast<const uint4*>(row + c0);
      const float2 a = hg::unpack_bf16x2(r.x), b = hg::unpack_bf16x2(r.y), cc = hg::unpack_bf16x2(r.z),
                   d = hg::unpack_bf16x2(r.w);
      const float* wc = w + static_cast<size_t>(c0) * k + j;
      acc += a.x * wc[0] + a.y * wc[k] + b.x * wc[2 * k] + b.y * wc[3 * k] + cc.x * wc[4 * k] +
             cc.y * wc[5 * k] + d.x * wc[6 * k] + d.y * wc[7 * k];
    }
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[static_cast<size_t>(s) * h + ho] = acc + (bias ? bias[0] : 0.f);
}

__global__ void avgpool_4_2_2_kernel(const float* __restrict__ x, int t, int t_out, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= t_out) return;
  const float* xb = x + static_cast<size_t>(b) * t;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = 2 * o - 2 + j;
    if (i >= 0 && i < t) acc += xb[i];
  }
  out[static_cast<size_t>(b) * t_out + o] = acc * 0.25f;  // zero padding counts (count_include_pad=True)
}

// bf16 [B*p][h_rows][C] -> fp32 [B][C][H][p]
__global__ void disc_export_fmap_kernel(const __nv_bfloat16* __restrict__ x, int period, int h, int h_rows,
                                        int c, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int seq = blockIdx.z;
  const int b = seq / period, wcol = seq % period;
  const int c0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int hh = h0 + i, cc = c0 + tx;
    tile[i][tx] = (hh < h && cc < c)
                      ? __bfloat162float(x[(static_cast<size_t>(seq) * h_rows + hh) * c + cc]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int cc = c0 + i, hh = h0 + tx;
    if (cc < c && hh < h)
      out[((static_cast<size_t>(b) * c + cc) * h + hh) * period + wcol] = tile[tx][i];
  }
}

// fp32 [B][C][H][p] -> bf16 [B*p][h_rows][C] (rows >= H are not written): the inverse of disc_export_fmap_kernel,
// for a gradient that arrives at an exported feature map
__global__ void disc_import_fmap_kernel(const float* __restrict__ in, int period, int h, int h_rows, int c,
                                        __nv_bfloat16* __restrict__ x) {
  __shared__ float tile[32][33];
  const int seq = blockIdx.z;
  const int b = seq / period, wcol = seq % period;
  const int c0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int cc = c0 + i, hh = h0 + tx;
    tile[i][tx] = (cc < c && hh < h) ? in[((static_cast<size_t>(b) * c + cc) * h + hh) * period + wcol] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int hh = h0 + i, cc = c0 + tx;
    if (hh < h && cc < c) x[(static_cast<size_t>(seq) * h_rows + hh) * c + cc] = __float2bfloat16(tile[tx][i]);
  }
}

}  // namespace

extern "C" int hg_disc_import_fmap(const float* in, int batch, int period, int h, int h_rows, int c, void* x,
                                   void* stream) {
  HG_REQUIRE(x && in && batch > 0 && period > 0 && h > 0 && h_rows >= h && c > 0, "hg_disc_import_fmap: bad arguments");
  HG_REQUIRE(batch * period <= 65535 && (h + 31) / 32 <= 65535, "hg_disc_import_fmap: grid too large");
  dim3 grid((c + 31) / 32, (h + 31) / 32, batch * period), block(32, 8);
  disc_import_fmap_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(in, period, h, h_rows, c,
                                                                                 static_cast<__nv_bfloat16*>(x));
  HG_CHECK_CUDA(cudaGetLastError());
  g_hg_launches.fetch_add(1, std::memory_order_relaxed);
  return HG_OK;
}

extern "C" int hg_disc_first_conv_fwd(const float* y, const float* w, const float* bias, int batch, int t,
                                      int period, int k, int stride, int pad, int cout, int h_rows_out,
                                      void* out, float slope, void* stream) {
  HG_REQUIRE(y && w && out, "hg_disc_first_conv_fwd: null pointer");
  HG_REQUIRE(batch > 0 && t > 1 && period >= 1 && k >= 1 && k <= kFirstMaxK && stride >= 1,
             "hg_disc_first_conv_fwd: bad shape");
  HG_REQUIRE(cout % 16 == 0 && cout <= 512, "hg_disc_first_conv_fwd: cout must be a multiple of 16");
  const int t_pad = (t + period - 1) / period * period;
  HG_REQUIRE(t_pad - t < t, "hg_disc_first_conv_fwd: reflect pad needs n_pad < t");
  const int h_in = t_pad / period;
  const int h_out = (h_in + 2 * pad - k) / stride + 1;
  HG_REQUIRE(h_out > 0 && h_rows_out >= h_out, "hg_disc_first_conv_fwd: output rows %d < %d", h_rows_out, h_out);
  HG_REQUIRE(batch * period <= 65535, "hg_disc_first_conv_fwd: too many sequences");
  dim3 grid((h_out + kFirstThreads - 1) / kFirstThreads, batch * period);
  const size_t smem = static_cast<size_t>(k + 1) * cout * sizeof(float);
  disc_first_conv_kernel<<<grid, kFirstThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      y, w, bias, t, period, h_in, h_out, h_rows_out, k, stride, pad, cout, slope,
      static_cast<__nv_bfloat16*>(out));
  HG_CHECK_CUDA(cudaGetLastError());
  g_hg_launches.fetch_add(1, std::memory_order_relaxed);
  return HG_OK;
}

extern "C" int hg_disc_last_conv_fwd(const void* x, const float* w, const float* bias, int nseq, int h,
                                     int h_rows, int c, int k, float* out, void* stream) {
  HG_REQUIRE(x && w && out, "hg_disc_last_conv_fwd: null pointer");
  HG_REQUIRE(nseq > 0 && h > 0 && h_rows >= h && c % 8 == 0 && (k & 1), "hg_disc_last_conv_fwd: bad shape");
  const long long warps = static_cast<long long>(nseq) * h;
  const int blocks = static_cast<int>((warps * 32 + 255) / 256);
  disc_last_conv_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), w, bias, nseq, h, h_rows, c, k, out);
  HG_CHECK_CUDA(cudaGetLastError());
  g_hg_launches.fetch_add(1, std::memory_order_relaxed);
  return HG_OK;
}

extern "C" int hg_avgpool_4_2_2_fwd(const float* x, int batch, int t, float* out, void* stream) {
  HG_REQUIRE(x && out && batch > 0 && t > 0 && batch <= 65535, "hg_avgpool_4_2_2_fwd: bad arguments");
  const int t_out = t / 2 + 1;  // floor((t + 4 - 4) / 2) + 1
  dim3 grid((t_out + 255) / 256, batch);
  avgpool_4_2_2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, t, t_out, out);
  HG_CHECK_CUDA(cudaGetLastError());
  g_hg_launches.fetch_add(1, std::memory_order_relaxed);
  return HG_OK;
}

extern "C" int hg_disc_export_fmap(const void* x, int batch, int period, int h, int h_rows, int c,
                                   float* out, void* stream) {
  HG_REQUIRE(x && out && batch > 0 && period > 0 && h > 0 && h_rows >= h && c > 0,
             "hg_disc_export_fmap: bad arguments");
  HG_REQUIRE(batch * period <= 65535 && (h + 31) / 32 <= 65535, "hg_disc_export_fmap: grid too large");
  dim3 grid((c + 31) / 32, (h + 31) / 32, batch * period), block(32, 8);
  disc_export_fmap_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), period, h, h_rows, c, out);
  HG_CHECK_CUDA(cudaGetLastError());
  g_hg_launches.fetch_add(1, std::memory_order_relaxed);
  return HG_OK;
}
